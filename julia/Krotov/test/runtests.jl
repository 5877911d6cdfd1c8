# Smoke tests of the Julia `Krotov` module over libkrotov_cuda.  NOT EXECUTED IN THIS REPOSITORY (no Julia in the build
# image); a maintainer with Julia, QuantumControl.jl and a B200 runs
#
#     KROTOV_CUDA_LIB=/path/to/libkrotov_cuda.so julia --project=julia/Krotov julia/Krotov/test/runtests.jl
#
# and, for the full behaviour, the reference's own test suite against this module (it is a drop-in: same module name,
# same `optimize(problem; method=Krotov)`).  The same scenarios run here through the Python mirror in tests/.
using Test
using LinearAlgebra
using QuantumControl
using QuantumControl: hamiltonian, Trajectory, ControlProblem, optimize
using QuantumControl.Functionals: J_T_sm
using QuantumControl.Shapes: flattop
using QuantumPropagators: Cheby
using Krotov

# the two-level system of the reference's TLS test (test/test_tls_optimization.jl), with the Chebyshev propagator
function tls_problem(; kwargs...)
    σz = ComplexF64[1 0; 0 -1]
    σx = ComplexF64[0 1; 1 0]
    T = 5.0
    ϵ(t) = 0.2 * flattop(t; T = T, t_rise = 0.3, func = :blackman)
    H = hamiltonian(-0.5 * σz, (σx, ϵ))
    tlist = collect(range(0, T, length = 501))
    trajectories = [Trajectory(ComplexF64[1, 0], H; target_state = ComplexF64[0, 1])]
    ControlProblem(trajectories, tlist; prop_method = Cheby, J_T = J_T_sm, lambda_a = 1.0, update_shape = t -> 1.0,
                   iter_stop = 5, print_iters = false, kwargs...)
end

@testset "two-level system, first order" begin
    res = optimize(tls_problem(); method = Krotov)
    @test res.converged && res.message == "Reached maximum number of iterations"
    @test res.J_T < 1e-3
    @test 1.0 < maximum(abs.(res.optimized_controls[1])) < 1.2
    # the committed 50-digit exact-propagator history (tests/golden/c1_tls_exact50.json): 1.7369162512568997e-05 after 5
    @test abs(res.J_T - 1.7369162512568997e-05) < 1e-12
end

@testset "continue_from reproduces the uninterrupted run" begin
    a = optimize(tls_problem(iter_stop = 2); method = Krotov)
    b = optimize(tls_problem(iter_stop = 5, continue_from = a); method = Krotov)
    c = optimize(tls_problem(iter_stop = 5); method = Krotov)
    @test b.iter == 5
    @test abs(b.J_T - c.J_T) < 1e-12
end

@testset "second order: constant sigma as a boundary condition" begin
    res = optimize(tls_problem(iter_stop = 3, sigma = -2.0); method = Krotov)
    # tests/golden/c1_tls_sigma_exact50.json (50-digit exact-propagator optimisation with the general second-order update)
    @test abs(res.J_T - 0.40083066640282965) < 1e-12
    @test_throws ArgumentError optimize(tls_problem(sigma = t -> -1.0 - t); method = Krotov)
    s = Krotov.NumericalSigma(1.0, 0.1)
    optimize(tls_problem(iter_stop = 2, sigma = s); method = Krotov)
    @test s(0.0) != -2.1   # refresh! re-estimated A after the first iteration
end
