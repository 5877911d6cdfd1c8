# Golden vectors from the UNMODIFIED reference.  NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image): this is
# the one step that turns "parity unpinned" (DESIGN.md section 2) into a pinned comparison.  On any machine with Julia,
#
#     julia -e 'using Pkg; Pkg.add(["Krotov", "QuantumControl", "QuantumPropagators", "JSON"])'
#     julia julia/reference_vectors.jl tests/golden/export/c1_tls tests/golden/julia/c1_tls.json
#
# reads a problem written by tools/export_problem.py, optimises it with stock Krotov.jl (`prop_method = Cheby`, the
# settings in problem.json) and writes the J_T history, the running costs and the optimised pulses on the midpoints.
# tests/test_oracle.py compares the oracles, and tests/test_parity_gpu.py the CUDA path, with every file found in
# tests/golden/julia/ at the BASELINE tolerances (1e-10 relative on J_T, 1e-9 on the pulses).
using JSON
using LinearAlgebra
using QuantumControl
using QuantumControl: hamiltonian, Trajectory, ControlProblem, optimize
using QuantumControl.Functionals: J_T_sm, J_T_ss, J_T_re
using QuantumPropagators: Cheby
using Krotov

folder, outfile = ARGS[1], ARGS[2]
meta = JSON.parsefile(joinpath(folder, "problem.json"))
d, N, L, N_T, n_gen = meta["d"], meta["N"], meta["L"], meta["N_T"], meta["n_gen"]

# the files are row-major: read into the reversed shape and permute
function readarray(name, T, dims...)
    raw = Array{T}(undef, reverse(dims)...)
    read!(joinpath(folder, name), raw)
    length(dims) == 1 ? raw : permutedims(raw, reverse(1:length(dims)))
end
tlist = readarray("tlist.f64", Float64, N_T + 1)
gen_of = readarray("gen_of_traj.i32", Int32, N) .+ 1
H0 = readarray("H0.c128", ComplexF64, n_gen, d, d)
Hc = readarray("Hc.c128", ComplexF64, n_gen, L, d, d)
psi0 = readarray("psi0.c128", ComplexF64, N, d)
target = readarray("target.c128", ComplexF64, N, d)
pulses = readarray("pulses.f64", Float64, L, N_T)
shape = readarray("shape.f64", Float64, L, N_T)
missing_terms = Set((m[1] + 1, m[2] + 1) for m in meta["missing"])

# controls are the midpoint pulse vectors themselves (length N_T: `discretize_on_midpoints` takes them as they are,
# test/test_pulse_optimization.jl:42); all generators share the same control objects, as an ensemble does
controls = [pulses[l, :] for l in 1:L]
generators = [hamiltonian(H0[g, :, :], [(Hc[g, l, :, :], controls[l]) for l in 1:L if !((g, l) in missing_terms)]...) for g in 1:n_gen]
trajectories = [Trajectory(psi0[k, :], generators[gen_of[k]]; target_state = target[k, :]) for k in 1:N]
J_T = Dict("sm" => J_T_sm, "ss" => J_T_ss, "re" => J_T_re)[meta["functional"]]
pulse_options = IdDict(controls[l] => Dict(:lambda_a => meta["lambda_a"][l], :update_shape => shape[l, :]) for l in 1:L)

history = Dict("J_T" => Float64[], "g_a_int" => Vector{Float64}[], "tau_re" => Float64[], "tau_im" => Float64[])
final_pulses = Ref{Any}(nothing)
function record(wrk, iter, ϵ_new, ϵ_old)
    push!(history["J_T"], wrk.result.J_T)
    iter > 0 && push!(history["g_a_int"], copy(wrk.g_a_int))
    final_pulses[] = [copy(ϵ) for ϵ in ϵ_new]
    history["tau_re"] = real.(wrk.result.tau_vals)
    history["tau_im"] = imag.(wrk.result.tau_vals)
    nothing
end

prop = Dict{Symbol,Any}(:prop_method => Cheby, :prop_cheby_coeffs_limit => meta["cheby_coeffs_limit"],
                        :prop_specrange_buffer => meta["specrange_buffer"])
if meta["specrange"] !== nothing
    prop[:prop_E_min], prop[:prop_E_max] = meta["specrange"]
end
problem = ControlProblem(trajectories, tlist; J_T = J_T, pulse_options = pulse_options, iter_stop = meta["iters"],
                         print_iters = false, callback = record, prop...)
result = optimize(problem; method = Krotov)
@assert result.message == "Reached maximum number of iterations" result.message

open(outfile, "w") do io
    JSON.print(io, Dict("source" => "Krotov.jl $(pkgversion(Krotov)), QuantumControl $(pkgversion(QuantumControl)), Julia $(VERSION)",
                        "problem" => meta["name"], "iters" => meta["iters"], "J_T" => history["J_T"],
                        "g_a_int" => history["g_a_int"], "pulses" => final_pulses[],
                        "tau_re" => history["tau_re"], "tau_im" => history["tau_im"]))
end
println("wrote ", outfile, ": J_T = ", history["J_T"])
