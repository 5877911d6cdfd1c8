# `optimize(problem; method=Krotov)` with the hot path in libkrotov_cuda.
#
# Behaviour follows the driver of JuliaQuantumControl/Krotov.jl (src/optimize.jl:155-235, 374-496): same keyword
# arguments, same callback protocol, same convergence / exception / atexit handling, same iteration table.  The two
# functions that ARE the hot path there -- `krotov_initial_fw_prop!` (:247-265, a loop over `prop_step!`) and
# `krotov_iteration` (:279-371, backward sweep + sequential update + forward sweep) -- are one `ccall` each here.
# NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image); the same call sequence is what
# krotov.jl_b200/optimize.py drives through ctypes and what the GPU tests exercise.
using QuantumControl.QuantumPropagators.Controls: discretize
using QuantumControl: set_atexit_save_optimization
using Dates: now, value
using LinearAlgebra: ⋅, ishermitian
using Printf

import QuantumControl: optimize, make_print_iters

import .LibKrotovCuda

optimize(problem, method::Val{:Krotov}) = optimize_krotov(problem)
optimize(problem, method::Val{:krotov}) = optimize_krotov(problem)

# The range hook of the reference (src/optimize.jl:238-244): the spectral envelope of a propagator is re-derived when
# twice the current amplitude range leaves the stored range, and is then derived for five times the current range.
function transform_control_ranges(c, ϵ_min, ϵ_max, check)
    f = check ? 2 : 5
    (min(ϵ_min, f * ϵ_min), max(ϵ_max, f * ϵ_max))
end

pulse_matrix(pulses) = reduce(hcat, pulses)   # Matrix(N_T, L) == [L][N_T] on the wire

# ---- hot path -------------------------------------------------------------------------------------------------------------
# All trajectories at once (the reference loops over k): range check of `reinit_prop!`, then one device sweep.
function krotov_initial_fw_prop!(ϵ⁽⁰⁾, wrk)
    if reinit!(wrk.fw_cheby, ϵ⁽⁰⁾, transform_control_ranges)
        push!(wrk.fw_cheby, wrk.handle, LibKrotovCuda.KROTOV_FORWARD)
    end
    LibKrotovCuda.forward(wrk.handle, pulse_matrix(ϵ⁽⁰⁾))
    wrk.states_cache = nothing
end

function krotov_iteration(wrk, ϵ⁽ⁱ⁾, ϵ⁽ⁱ⁺¹⁾)
    # boundary condition chi_k(T): built-in functionals form it on the device from tau; anything else is the user's
    # (or make_chi's) Julia function, evaluated on the host on the final states
    if wrk.chi_kind == LibKrotovCuda.KROTOV_CHI_HOST
        chi = wrk.kwargs[:chi]
        Ψ = final_states(wrk)
        χ = wrk.chi_takes_tau ? chi(Ψ, wrk.trajectories; tau = wrk.result.tau_vals) : chi(Ψ, wrk.trajectories)
        χT = reduce(hcat, [Vector{ComplexF64}(x) for x in χ])
        if wrk.sigma !== nothing
            # second order (the TODO at src/optimize.jl:350 of the reference).  For Hermitian generators and a sigma that is
            # constant over the time grid, chi(t_n) + sigma/2 (Psi^(i+1)(t_n) - Psi^(i)(t_n)) acts in the update like the
            # backward-propagated chi(T) - sigma/2 Psi^(i)(T): Im<Psi|mu|Psi> = 0 and the backward propagator under the
            # guess pulses inverts the forward propagator that produced Psi^(i).  No second forward storage is needed.
            σ = sigma_value(wrk.sigma, wrk.result.tlist)
            wrk.sigma_info = (forward_states0 = map(copy, Ψ), chi_states = [Vector{ComplexF64}(x) for x in χ])
            χT = χT .- (σ / 2) .* reduce(hcat, Ψ)
        end
        LibKrotovCuda.set_chi(wrk.handle, χT)
    end
    # `reinit_prop!` of the backward propagators under the guess pulses, then of the forward propagators, whose check
    # sees the update buffers as they are NOW (the reference's aliased arrays)
    if reinit!(wrk.bw_cheby, ϵ⁽ⁱ⁾, transform_control_ranges)
        push!(wrk.bw_cheby, wrk.handle, LibKrotovCuda.KROTOV_BACKWARD)
    end
    if reinit!(wrk.fw_cheby, ϵ⁽ⁱ⁺¹⁾, transform_control_ranges)
        push!(wrk.fw_cheby, wrk.handle, LibKrotovCuda.KROTOV_FORWARD)
    end
    fresh = Matrix{Float64}(undef, length(ϵ⁽ⁱ⁾[1]), length(ϵ⁽ⁱ⁾))
    LibKrotovCuda.iterate!(wrk.handle, pulse_matrix(ϵ⁽ⁱ⁾), fresh, wrk.g_a_int)
    for l in eachindex(ϵ⁽ⁱ⁺¹⁾)
        ϵ⁽ⁱ⁺¹⁾[l] .= @view fresh[:, l]   # in place: callbacks hold references to these arrays
    end
    wrk.states_cache = nothing
end

# ---- host bookkeeping -----------------------------------------------------------------------------------------------------
function update_result!(wrk::KrotovWrk, i::Int64)
    res = wrk.result
    J_T = wrk.kwargs[:J_T]
    res.J_T_prev = res.J_T
    Ψ = final_states(wrk)
    for k in eachindex(Ψ)
        res.states[k] = Ψ[k]
    end
    LibKrotovCuda.get_tau!(wrk.handle, res.tau_vals)   # taus! on the device (zero where there is no target)
    res.J_T = wrk.J_T_takes_tau ? J_T(res.states, wrk.trajectories; tau = res.tau_vals) : J_T(res.states, wrk.trajectories)
    i > 0 && (res.iter = i)
    if i >= res.iter_stop
        res.converged = true
        res.message = "Reached maximum number of iterations"
    end
    before = res.end_local_time
    res.end_local_time = now()
    res.secs = value(res.end_local_time - before) / 1000.0
end

# The "update sigma" step (TODO at src/optimize.jl:369 of the reference), once the iteration's J_T is known: a sigma that
# defines `refresh!` is handed the final-time states of this and of the previous iteration, chi(T), J_T.
function update_sigma!(wrk::KrotovWrk, ϵ⁽ⁱ⁺¹⁾, ϵ⁽ⁱ⁾)
    applicable(refresh!, wrk.sigma) || return
    res = wrk.result
    refresh!(wrk.sigma; forward_states = map(copy, res.states), J_T = res.J_T, J_T_prev = res.J_T_prev,
             optimized_pulses = ϵ⁽ⁱ⁺¹⁾, guess_pulses = ϵ⁽ⁱ⁾, trajectories = wrk.trajectories, result = res, wrk.sigma_info...)
end

"""Hook for second-order functions: `refresh!(sigma; forward_states, forward_states0, chi_states, J_T, J_T_prev, ...)`."""
function refresh! end

# Estimate of the second-order constant A from one iteration (Reich, Ndong, Koch, J. Chem. Phys. 136, 104103 (2012)):
#   A = [ sum_k 2 Re<chi_k(T)|dPsi_k(T)> + dJ_T ] / sum_k |dPsi_k(T)|^2 ,   dPsi = Psi^(i+1)(T) - Psi^(i)(T)
function numerical_estimate_A(forward_states, forward_states0, chi_states, ΔJ_T)
    ΔΨ = [a - b for (a, b) in zip(forward_states, forward_states0)]
    den = sum(real(x ⋅ x) for x in ΔΨ)
    den <= 1e-30 && return 0.0
    (sum(2 * real(c ⋅ x) for (c, x) in zip(chi_states, ΔΨ)) + ΔJ_T) / den
end

# sigma(t) = -max(eps_A, 2A + eps_A), A re-estimated after every iteration
mutable struct NumericalSigma
    A::Float64
    eps_A::Float64
end
(s::NumericalSigma)(t) = -max(s.eps_A, 2 * s.A + s.eps_A)
function refresh!(s::NumericalSigma; forward_states, forward_states0, chi_states, J_T, J_T_prev, kwargs...)
    s.A = numerical_estimate_A(forward_states, forward_states0, chi_states, J_T - J_T_prev)
end

function finalize_result!(ϵ_opt, wrk::KrotovWrk)
    res = wrk.result
    res.end_local_time = now()
    for l in eachindex(ϵ_opt)
        res.optimized_controls[l] = discretize(ϵ_opt[l], res.tlist)
    end
end

function optimize_krotov(problem)
    kw = problem.kwargs
    if haskey(kw, :update_hook) || haskey(kw, :info_hook)
        throw(ArgumentError("The `update_hook` and `info_hook` arguments have been superseded by the `callback` argument"))
    end
    callback = get(kw, :callback, (args...) -> nothing)
    check_convergence! = get(kw, :check_convergence, res -> res)   # "maximum number of iterations" is always checked
    wrk = KrotovWrk(problem; verbose = get(kw, :verbose, false))
    ϵ⁽ⁱ⁾, ϵ⁽ⁱ⁺¹⁾ = wrk.pulses0, wrk.pulses1

    record!(info) = (info === nothing || isempty(info)) || push!(wrk.result.records, info)

    if get(kw, :skip_initial_forward_propagation, false)
        @info "Skipping initial forward propagation"   # the device holds the initial states and their tau
    else
        krotov_initial_fw_prop!(ϵ⁽ⁱ⁾, wrk)
    end
    update_result!(wrk, 0)
    record!(callback(wrk, 0, ϵ⁽ⁱ⁺¹⁾, ϵ⁽ⁱ⁾))

    i = wrk.result.iter   # 0 unless continuing a previous optimisation
    atexit_filename = get(kw, :atexit_filename, nothing)
    if atexit_filename !== nothing
        set_atexit_save_optimization(atexit_filename, wrk.result)
        isinteractive() || @info "Set callback to store result in $(relpath(atexit_filename)) on unexpected exit."
    end
    try
        while !wrk.result.converged
            i += 1
            krotov_iteration(wrk, ϵ⁽ⁱ⁾, ϵ⁽ⁱ⁺¹⁾)   # a non-zero status of the library is an ErrorException raised here
            update_result!(wrk, i)
            wrk.sigma === nothing || update_sigma!(wrk, ϵ⁽ⁱ⁺¹⁾, ϵ⁽ⁱ⁾)
            record!(callback(wrk, i, ϵ⁽ⁱ⁺¹⁾, ϵ⁽ⁱ⁾))
            check_convergence!(wrk.result)
            ϵ⁽ⁱ⁾, ϵ⁽ⁱ⁺¹⁾ = ϵ⁽ⁱ⁺¹⁾, ϵ⁽ⁱ⁾
        end
    catch exc
        get(kw, :rethrow_exceptions, false) && rethrow()
        wrk.result.message = "Exception: " * sprint(showerror, exc)   # e.g. Ctrl-C in an interactive session
    end
    finalize_result!(ϵ⁽ⁱ⁾, wrk)
    atexit_filename === nothing || popfirst!(Base.atexit_hooks)
    LibKrotovCuda.destroy(wrk.handle)
    wrk.result
end

# ---- iteration table --------------------------------------------------------------------------------------------------------
make_print_iters(::Val{:Krotov}; kwargs...) = make_krotov_print_iters(; kwargs...)
make_print_iters(::Val{:krotov}; kwargs...) = make_krotov_print_iters(; kwargs...)

const TABLE_COLUMNS = ["iter.", "J_T", "∫gₐ(t)dt", "J", "ΔJ_T", "ΔJ", "secs"]

function make_krotov_print_iters(; kwargs...)
    wanted = Set(get(kwargs, :store_iter_info, Set()))
    for item in wanted
        item in TABLE_COLUMNS ||
            throw(ArgumentError("Item $(repr(item)) in `store_iter_info` is not one of $(repr(TABLE_COLUMNS)))"))
    end
    keep = [c in wanted for c in TABLE_COLUMNS]

    function print_table(wrk, iteration, args...)
        J_T = wrk.result.J_T
        running = sum(wrk.g_a_int)
        ΔJ_T = J_T - wrk.result.J_T_prev
        numbers = (iteration, J_T, running, J_T + running, ΔJ_T, ΔJ_T + running, wrk.result.secs)
        widths = (max(length(string(get(wrk.kwargs, :iter_stop, 5000))), 6), 11, 11, 11, 11, 11, 8)
        if iteration == 0
            println(join(lpad(h, w) for (h, w) in zip(TABLE_COLUMNS, widths)))
        end
        sci(x) = @sprintf("%.2e", x)
        cells = (string(iteration), sci(numbers[2]), sci(numbers[3]), sci(numbers[4]),
                 iteration > 0 ? sci(numbers[5]) : "n/a", iteration > 0 ? sci(numbers[6]) : "n/a",
                 @sprintf("%.1f", numbers[7]))
        println(join(lpad(c, w) for (c, w) in zip(cells, widths)))
        flush(stdout)
        Tuple(v for (v, k) in zip(numbers, keep) if k)
    end
    print_table
end
