# KrotovResult -- the 16 fields of the reference's result type, same names, order and meaning (src/result.jl:34-51 of
# JuliaQuantumControl/Krotov.jl), so that `check_convergence` functions, callbacks and `continue_from` written for the
# reference keep working.  NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image).
using QuantumControl.QuantumPropagators.Controls: get_controls, discretize
using QuantumControl: AbstractOptimizationResult
using Dates
using Printf

mutable struct KrotovResult{STST} <: AbstractOptimizationResult
    tlist::Vector{Float64}
    iter_start::Int64                            # the starting iteration number
    iter_stop::Int64                             # the maximum iteration number
    iter::Int64                                  # the current iteration number
    secs::Float64                                # seconds that the last iteration took
    tau_vals::Vector{ComplexF64}                 # overlaps of the propagated states with the target states
    J_T::Float64                                 # current value of the final-time functional
    J_T_prev::Float64                            # previous value of J_T
    guess_controls::Vector{Vector{Float64}}      # on the points of tlist
    optimized_controls::Vector{Vector{Float64}}  # on the points of tlist
    states::Vector{STST}                         # the forward-propagated states after each iteration
    start_local_time::DateTime
    end_local_time::DateTime
    records::Vector{Tuple}                       # what the callbacks returned, one tuple per iteration
    converged::Bool
    message::String
end

# A fresh result for `problem`: iteration counters from the `iter_start` / `iter_stop` keywords, controls sampled on
# the time grid, everything else zeroed.
function KrotovResult(problem)
    kw = problem.kwargs
    grid = Vector{Float64}(problem.tlist)
    first_iter = Int64(get(kw, :iter_start, 0))
    guesses = Vector{Float64}[discretize(c, grid) for c in get_controls(problem.trajectories)]
    psis = [similar(t.initial_state) for t in problem.trajectories]
    t_now = now()
    KrotovResult{eltype(psis)}(
        grid, first_iter, Int64(get(kw, :iter_stop, 5000)), first_iter, 0.0,
        zeros(ComplexF64, length(psis)), 0.0, 0.0,
        guesses, map(copy, guesses), psis,
        t_now, t_now, Tuple[], false, "in progress",
    )
end

Base.show(io::IO, r::KrotovResult) = print(io, "KrotovResult<", r.message, ">")

function Base.show(io::IO, ::MIME"text/plain", r::KrotovResult)
    elapsed = Dates.canonicalize(Dates.CompoundPeriod(r.end_local_time - r.start_local_time))
    println(io, "Krotov Optimization Result")
    println(io, "--------------------------")
    println(io, "- Started at ", r.start_local_time)
    println(io, "- Number of trajectories: ", length(r.states))
    println(io, "- Number of iterations: ", max(r.iter - r.iter_start, 0))
    println(io, "- Value of functional: ", @sprintf("%.5e", r.J_T))
    println(io, "- Reason for termination: ", r.message)
    println(io, "- Ended at ", r.end_local_time, " (", elapsed, ")")
end

# Results of other optimisers (`continue_from = res_grape`, test/test_tls_optimization.jl:100-130 of the reference):
# every field this type shares with the foreign result is taken over, the rest keeps its fresh value.
function Base.convert(::Type{KrotovResult}, other::AbstractOptimizationResult)
    other isa KrotovResult && return other
    STST = eltype(other.states)
    t_now = now()
    fresh = KrotovResult{STST}(
        Float64[], 0, 5000, 0, 0.0, ComplexF64[], 0.0, 0.0, Vector{Float64}[], Vector{Float64}[], STST[],
        t_now, t_now, Tuple[], false, "in progress",
    )
    for name in fieldnames(KrotovResult)
        if hasproperty(other, name)
            setfield!(fresh, name, convert(fieldtype(KrotovResult{STST}, name), getproperty(other, name)))
        end
    end
    fresh
end
