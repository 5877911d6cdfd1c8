"""The oracle against everything that can pin it (CPU only).

PARITY UNPINNED: the reference's tests hold no numeric golden vector for the Krotov path; what they
do hold is asserted here on the restatement: the two TLS inequalities
(test/test_tls_optimization.jl:66-67).  Beyond that: internal invariants, the NumPy <-> C cross-check,
the committed golden vectors, and SURVEY.md's independent cross-check history."""
import json
import os

import numpy as np
import pytest

import workloads as W
from oracle import c_oracle as C
from oracle import krotov_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    with open(os.path.join(GOLD, name + ".json")) as fh:
        return json.load(fh)


# provisional history of SURVEY.md section 4 (exact 2x2 propagator, independent throw-away code)
SURVEY_JT = [0.9514590189717334, 0.7236195837372406, 0.18590337379475952, 0.010549341051677374,
             0.0004287760060682766, 1.736916251138254e-05]
SURVEY_GA = [0.06710525941726, 0.19617485112125, 0.08143906188287, 0.00484817505855, 0.00019737219929]


def test_tls_reference_inequalities_expm():
    """test/test_tls_optimization.jl:47-70 -- same problem, same propagator (ExpProp), same assertions."""
    h = O.optimize_krotov(W.to_oracle(W.c1_tls()), 5, "expm")
    assert h["J_T"][-1] < 1e-3
    assert 1.0 < np.abs(h["optimized_controls"][0]).max() < 1.2
    assert len(h["optimized_controls"][0]) == 501  # optimized_controls are ON tlist (test_pulse_optimization.jl:28)


def test_tls_matches_survey_crosscheck():
    h = O.optimize_krotov(W.to_oracle(W.c1_tls()), 5, "expm")
    assert np.allclose(h["J_T"], SURVEY_JT, rtol=1e-9, atol=0)
    assert np.allclose([g[0] for g in h["g_a_int"]], SURVEY_GA, rtol=1e-9, atol=0)


def test_tls_cheby_close_to_expm_and_monotonic():
    p = W.to_oracle(W.c1_tls())
    a = O.optimize_krotov(p, 5, "cheby")
    b = O.optimize_krotov(p, 5, "expm")
    assert np.allclose(a["J_T"], b["J_T"], rtol=1e-7, atol=1e-12)
    assert np.abs(a["pulses"] - b["pulses"]).max() < 1e-11
    assert a["J_T"][-1] < 1e-3 and 1.0 < np.abs(a["optimized_controls"][0]).max() < 1.2
    # Krotov's monotonic convergence: Delta J = Delta J_T + sum_l int g_a dt <= 0
    for i in range(1, 6):
        assert a["J_T"][i] - a["J_T"][i - 1] + float(np.sum(a["g_a_int"][i - 1])) < 0.0


@pytest.mark.parametrize("name,make,iters,method", [
    ("c1_tls_cheby", W.c1_tls, 5, "cheby"),
    ("c1_tls_expm", W.c1_tls, 5, "expm"),
    ("c2_transmon_x", W.c2_transmon_x, 4, "cheby"),
    ("dummy_d10", lambda: W.dummy_dense(d=10, n_traj=2, n_controls=2), 3, "cheby"),
])
def test_numpy_oracle_reproduces_golden(name, make, iters, method):
    g = gold(name)
    h = O.optimize_krotov(W.to_oracle(make()), iters, method)
    assert np.allclose(h["J_T"], g["J_T"], rtol=1e-12, atol=1e-15)
    assert np.abs(h["pulses"] - np.array(g["pulses"])).max() < 1e-13


@pytest.mark.parametrize("make,iters", [(W.c1_tls, 5), (W.c2_transmon_x, 4),
                                        (lambda: W.c4_ensemble(n_samples=2, n_grid=101), 2)])
def test_c_oracle_matches_numpy_oracle(make, iters):
    """Two independent restatements (different Bessel and eigenvalue code): agreement is at the
    coefficient-rounding floor, ~1e-13 absolute in J_T ('parity floor')."""
    p = W.to_oracle(make())
    a = O.optimize_krotov(p, iters)
    b = C.optimize_krotov_c(p, iters)
    assert np.abs(np.array(a["J_T"]) - np.array(b["J_T"])).max() < 5e-13
    assert np.abs(a["pulses"] - b["pulses"]).max() < 1e-11
    assert np.abs(np.array(a["g_a_int"]) - np.array(b["g_a_int"])).max() < 1e-12
    assert a["m_fw"][-1][0] == b["m"][0] and a["m_bw"][-1][0] == b["m"][1]


def test_c_oracle_reproduces_c3_golden():
    g = gold("c3_two_transmon")
    b = C.optimize_krotov_c(W.to_oracle(W.c3_two_transmon()), 2)
    assert np.abs(np.array(b["J_T"]) - np.array(g["J_T"])).max() < 5e-12
    assert np.abs(b["pulses"] - np.array(g["pulses"])).max() < 1e-11


def test_chebyshev_step_matches_expm_and_is_unitary():
    """Appendix A.1: one Chebyshev step agrees with the matrix exponential to ~1e-13, both directions."""
    from scipy.linalg import expm

    w = W.c3_two_transmon(n_grid=11)
    p = W.to_oracle(w)
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(25) + 1j * rng.standard_normal(25)
    psi /= np.linalg.norm(psi)
    for backward in (False, True):
        pr = O.ChebyPropagator(p.H0[0], p.Hc[0], p.tlist, list(p.pulses), backward)
        pr.reinit_prop(psi, O.transform_control_ranges)
        out = pr.prop_step()
        n = 9 if backward else 0
        H = p.H0[0] + p.pulses[0][n] * p.Hc[0][0] + p.pulses[1][n] * p.Hc[0][1]
        dt = p.tlist[1] - p.tlist[0]
        ref = expm((+1j if backward else -1j) * H * dt) @ psi
        assert np.abs(out - ref).max() < 5e-13
        assert abs(np.linalg.norm(out) - 1.0) < 1e-13


def test_range_logic_factor_2_and_5():
    """src/optimize.jl:238-244 and its effect on the spectral envelope (Appendix A.1)."""
    assert O.transform_control_ranges(None, -1.0, 2.0, True) == (-2.0, 4.0)
    assert O.transform_control_ranges(None, -1.0, 2.0, False) == (-5.0, 10.0)
    assert O.transform_control_ranges(None, 0.5, 2.0, False) == (0.5, 10.0)
    p = W.to_oracle(W.c2_transmon_x())
    pr = O.ChebyPropagator(p.H0[0], p.Hc[0], p.tlist, list(p.pulses), False)
    d0 = pr.Delta
    pr.reinit_prop(p.psi0[0], O.transform_control_ranges)  # first reinit always widens (init used raw ranges)
    assert pr.n_range_updates == 1 and pr.Delta > d0
    pr.reinit_prop(p.psi0[0], O.transform_control_ranges)  # same pulses: 2x check stays inside 5x range
    assert pr.n_range_updates == 1


def test_discretize_roundtrip_and_copy():
    t = np.linspace(0, 1, 11)
    v = np.arange(10.0)
    assert O.discretize_on_midpoints(v, t) is not v and np.array_equal(O.discretize_on_midpoints(v, t), v)
    on_grid = O.discretize(v, t)
    assert len(on_grid) == 11 and on_grid[0] == v[0] and on_grid[-1] == v[-1] and on_grid[3] == 0.5 * (v[2] + v[3])
    f = lambda x: 2.0 * x  # noqa: E731
    m = O.discretize_on_midpoints(f, t)
    assert m[0] == 0.0 and m[-1] == 2.0 and abs(m[4] - 2.0 * 0.45) < 1e-15
    assert abs(O.flattop(0.15, T=5, t_rise=0.3) - O.blackman(0.15, 0, 0.6)) < 1e-16
    assert O.flattop(2.5, T=5, t_rise=0.3) == 1.0 and O.flattop(0.0, T=5, t_rise=0.3) == 0.0


def test_oracles_against_50_digit_exact_propagator_optimisation():
    """An oracle-independent pin of the loop of src/optimize.jl:279-371: the TLS optimisation in 50-digit arithmetic
    with the closed-form 2x2 propagator (tests/mp_reference.py) depends on none of the recalled Chebyshev
    conventions.  The ExpProp oracle must agree to rounding, the Chebyshev oracles to the truncation level of the
    expansion (|a_m| <= 1e-12 per step): absolute 1e-12 in J_T -- which at J_T = 1.7e-5 is 2e-9 RELATIVE, the floor
    SURVEY.md section 0 (fact 5) predicts for two exact methods -- and 1e-12 in the pulses."""
    import mp_reference as M

    exact = M.tls_krotov_exact(5)
    p = W.to_oracle(W.c1_tls())
    ex = np.array(exact["J_T"])
    for h, tol_j, tol_p in ((O.optimize_krotov(p, 5, "expm"), 5e-14, 5e-14), (O.optimize_krotov(p, 5, "cheby"), 1e-12, 1e-12),
                            (C.optimize_krotov_c(p, 5), 1e-12, 1e-12)):
        assert np.abs(np.array(h["J_T"]) - ex).max() < tol_j
        assert np.abs(np.asarray(h["pulses"])[0] - np.array(exact["pulses"])).max() < tol_p
        assert np.abs(np.array([g[0] for g in h["g_a_int"]]) - np.array(exact["g_a_int"])).max() < 1e-12
    assert ex[-1] < 1e-3  # test_tls_optimization.jl:66 holds for the exact optimisation too
    # the reference's ExpProp run (what its own test uses) is reproduced to 1e-10 relative at every iteration
    h = O.optimize_krotov(p, 5, "expm")
    assert (np.abs(np.array(h["J_T"]) - ex) / ex).max() < 1e-9


def test_blocked_oracle_equals_per_trajectory_oracle():
    """optimize_krotov_blocked (trajectories of a generator as columns, BLAS-3) is the per-trajectory oracle up to
    BLAS summation order; two generators, a missing control term, J_T_sm (needs the global tau sum)."""
    w = W.dummy_dense(d=12, n_traj=5, n_controls=2, n_grid=21, seed=5, functional="sm")
    w.H0 = [w.H0[0], w.H0[0] * 1.1]
    w.Hc = [w.Hc[0], [w.Hc[0][0] * 0.9, None]]
    w.gen_of_traj = np.array([0, 0, 1, 1, 1])
    a = O.optimize_krotov(W.to_oracle(w), 3)
    b = O.optimize_krotov_blocked(W.to_oracle(w), 3)
    assert np.abs(np.array(a["J_T"]) - np.array(b["J_T"])).max() < 1e-14
    assert np.abs(a["pulses"] - b["pulses"]).max() < 1e-14
    assert np.abs(np.array(a["g_a_int"]) - np.array(b["g_a_int"])).max() < 1e-15
    assert np.abs(a["states"] - b["states"]).max() < 1e-14


def test_c_oracle_reports_divergence_instead_of_crashing():
    w = W.c4_ensemble(n_samples=1, n_grid=201)
    w.functional, w.lambda_a = "ss", 1e-4
    with pytest.raises(RuntimeError):
        C.optimize_krotov_c(W.to_oracle(w), 4, n_threads=2)


def test_second_order_oracle_against_50_digit_exact_propagator_optimisation():
    """Second-order update (`sigma`; documented at src/optimize.jl:104-105, TODOs at :187, :350, :369; restated from
    Reich, Ndong, Koch, J. Chem. Phys. 136, 104103 (2012) and the `krotov` Python package): the oracle's general formula
    <chi + sigma_n/2 (Psi_new(t_n) - Psi_old(t_n))| mu |Psi_new(t_n)>  against the same update written independently in
    50-digit arithmetic with the closed-form propagator -- for a constant sigma and for one that varies over the grid."""
    import mp_reference as M

    p = W.to_oracle(W.c1_tls())
    for sigma in (-2.0, lambda t: -1.0 - 0.2 * t):
        sv = O.sigma_on_intervals(sigma, p.tlist)
        assert sv.shape == (p.N_T,) and (callable(sigma) or np.ptp(sv) == 0.0)
        exact = M.tls_krotov_exact(3, sigma=list(sv))
        for method, tol in (("expm", 5e-14), ("cheby", 1e-12)):
            h = O.optimize_krotov(p, 3, method, sigma=sigma)
            assert np.abs(np.array(h["J_T"]) - np.array(exact["J_T"])).max() < tol
            assert np.abs(h["pulses"][0] - np.array(exact["pulses"])).max() < tol
            assert np.abs(np.array([g[0] for g in h["g_a_int"]]) - np.array(exact["g_a_int"])).max() < tol
    first = M.tls_krotov_exact(3)
    assert abs(first["J_T"][-1] - exact["J_T"][-1]) > 1e-2  # (the second-order term is not a no-op here)


def test_second_order_boundary_fold_identity():
    """What the device path relies on (krotov.jl_b200/second_order.py): for Hermitian generators and a sigma that is
    constant over the grid, the general second-order update equals the FIRST-order update started from the boundary
    condition chi(T) - sigma/2 Psi_old(T).  Checked inside the oracle (no product code), to the propagator's accuracy."""
    def folded(p, iters, sigma):
        wrk = O.OracleWrk(p)
        e0, e1 = wrk.pulses0, wrk.pulses1
        for k in range(p.N):
            O.krotov_initial_fw_prop(e0, p.psi0[k], k, wrk)
        O.update_result(wrk)
        J = [wrk.J_T]
        for _ in range(iters):
            def chi(Psi):
                c = O.chi_states(p.functional, wrk.tau_vals, p.weights(), p.target)
                return [np.array(ck) - 0.5 * sigma * np.array(ps) for ck, ps in zip(c, Psi)]
            O.krotov_iteration(wrk, e0, e1, chi=chi)
            O.update_result(wrk)
            J.append(wrk.J_T)
            e0, e1 = e1, e0
        return np.array(J), np.array(e0)

    for w, sigma in ((W.c1_tls(), -2.0), (W.c1_tls(), 0.7),
                     (W.dummy_dense(d=10, n_traj=3, n_controls=2, n_grid=51, functional="ss"), -2.0),
                     (W.dummy_dense(d=12, n_traj=4, n_controls=1, n_grid=31, functional="sm", seed=3), -0.4)):
        a = O.optimize_krotov(W.to_oracle(w), 3, sigma=sigma)
        J, e = folded(W.to_oracle(w), 3, sigma)
        assert np.abs(J - np.array(a["J_T"])).max() < 1e-12 and np.abs(e - a["pulses"]).max() < 1e-12
    # and it is an identity of UNITARY dynamics only: with a non-Hermitian drift the fold is wrong at first order in the loss
    w = W.dummy_dense(d=10, n_traj=3, n_controls=2, n_grid=51, functional="ss")
    w.H0 = [w.H0[0] - 0.05j * np.diag(np.arange(10.0))]
    a = O.optimize_krotov(W.to_oracle(w), 2, sigma=-2.0)
    J, e = folded(W.to_oracle(w), 2, -2.0)
    assert np.abs(e - a["pulses"]).max() > 1e-6


def test_numerical_estimate_of_A_vanishes_for_the_linear_functional():
    """A = [sum_k 2 Re<chi_k|dPsi_k> + dJ_T] / sum_k |dPsi_k|^2 is the curvature of J_T along the step: exactly zero for
    J_T_re (linear in the states), negative for the concave J_T_ss."""
    class Rec:
        def __init__(self):
            self.A = []

        def __call__(self, t):
            return -0.1

        def refresh(self, *, forward_states, forward_states0, chi_states, J_T, J_T_prev):
            self.A.append(O.numerical_estimate_A(forward_states, forward_states0, chi_states, J_T - J_T_prev))

    for functional, check in (("re", lambda A: abs(A) < 1e-12), ("ss", lambda A: A < -1e-3)):
        w = W.dummy_dense(d=10, n_traj=3, n_controls=2, n_grid=51, functional=functional)
        w.lambda_a = 0.05
        s = Rec()
        O.optimize_krotov(W.to_oracle(w), 3, sigma=s)
        assert len(s.A) == 3 and all(check(A) for A in s.A), s.A


@pytest.mark.parametrize("name,iters,sigma", [("c1_tls_exact50", 5, None), ("c1_tls_sigma_exact50", 3, -2.0)])
def test_oracles_reproduce_the_committed_exact_vectors(name, iters, sigma):
    """tests/golden/*_exact50.json are NOT oracle output: they come from the 50-digit exact-propagator optimisation
    (tests/golden/make_golden_exact.py).  The committed copy must equal a fresh run, and the oracles must meet it."""
    import mp_reference as M

    g = gold(name)
    fresh = M.tls_krotov_exact(iters, **({} if sigma is None else {"sigma": sigma}))
    assert fresh["J_T"] == g["J_T"] and fresh["pulses"] == g["pulses"]
    p = W.to_oracle(W.c1_tls())
    for method, tol in (("expm", 5e-14), ("cheby", 1e-12)):
        h = O.optimize_krotov(p, iters, method, sigma=sigma)
        assert np.abs(np.array(h["J_T"]) - np.array(g["J_T"])).max() < tol
        assert np.abs(h["pulses"][0] - np.array(g["pulses"])).max() < tol


@pytest.mark.parametrize("name", ["c2_transmon_x_g101", "two_generators_d5", "non_hermitian_d4", "nonlinear_two_generators_d5"])
def test_oracles_against_exact_propagator_vectors_of_general_problems(name):
    """tests/golden/*_exact40.json: the Krotov loop in 40-digit arithmetic with `mpmath.expm` per interval
    (tests/mp_reference.py `krotov_exact_general`, written independently of the oracles; made by make_golden_exact.py).
    Pins, beyond the two-level system: the sum over several trajectories and the J_T_sm / J_T_ss / J_T_re boundary
    conditions for N > 1, two generators, a control one generator does not depend on, complex control operators, two
    controls updated in one time step, non-linear amplitudes (a quadratic and a shaped one; mu = a'(eps_guess) H_l)
    on two generators -- and that the backward sweep of a NON-Hermitian generator runs with its adjoint.  ExpProp oracle: rounding; Chebyshev oracles: the truncation level of the expansion."""
    import mp_reference as M

    make, iters = M.exact_cases()[name]
    g = gold(name + "_exact40")
    p = W.to_oracle(make())
    runs = [(O.optimize_krotov(p, iters, "expm"), 5e-14), (O.optimize_krotov(p, iters, "cheby"), 1e-12),
            (C.optimize_krotov_c(p, iters), 1e-12)]
    for h, tol in runs:
        assert np.abs(np.array(h["J_T"]) - np.array(g["J_T"])).max() < tol
        assert np.abs(np.asarray(h["pulses"]) - np.array(g["pulses"])).max() < tol
        assert np.abs(np.array(h["g_a_int"]) - np.array(g["g_a_int"])).max() < tol
    if name == "non_hermitian_d4":  # the committed vector is what a fresh 40-digit run gives (the cheapest case: ~4 s)
        fresh = M.krotov_exact_general(p, iters)
        assert fresh["J_T"] == g["J_T"] and fresh["pulses"] == g["pulses"]


@pytest.mark.parametrize("name", ["non_hermitian_d4", "two_generators_d5"])
def test_general_second_order_oracle_against_exact_propagator_vectors(name):
    """The oracle's GENERAL second-order update (sigma varying over the grid, a non-Hermitian generator, several
    trajectories and controls; previous trajectory stored) against the 40-digit exact-propagator loop with the same
    term written independently (tests/golden/*_sigma_exact40.json).  This is the formula the device path's boundary-
    condition fold is tested against in its domain of validity (Hermitian generators, constant sigma)."""
    import mp_reference as M

    make, iters = M.exact_cases()[name]
    g = gold(name + "_sigma_exact40")
    p = W.to_oracle(make())
    sv = O.sigma_on_intervals(lambda t: -0.6 - 0.3 * t, p.tlist)
    assert np.abs(sv - np.array(g["sigma"])).max() < 1e-15
    for method, tol in (("expm", 5e-14), ("cheby", 1e-12)):
        h = O.optimize_krotov(W.to_oracle(make()), iters, method, sigma=lambda t: -0.6 - 0.3 * t)
        assert np.abs(np.array(h["J_T"]) - np.array(g["J_T"])).max() < tol
        assert np.abs(h["pulses"] - np.array(g["pulses"])).max() < tol
        assert np.abs(np.array(h["g_a_int"]) - np.array(g["g_a_int"])).max() < tol
    first = gold(name + "_exact40")
    assert abs(first["J_T"][-1] - g["J_T"][-1]) > 1e-3


def _reference_run_vectors():
    folder = os.path.join(GOLD, "julia")
    return sorted(f[:-5] for f in os.listdir(folder) if f.endswith(".json"))


def test_exported_problem_files_round_trip():
    """tools/export_problem.py writes a workload as plain binaries for `julia/reference_vectors.jl` (the unmodified
    reference on a machine with Julia); the committed exports must be exactly what the workloads produce today, and
    reading them back must give the oracle the same problem bit for bit."""
    import sys
    import tempfile

    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    import export_problem as X

    for name in ("c1_tls", "c2_transmon_x"):
        meta, p = X.load(os.path.join(GOLD, "export", name))
        make, iters = X.CASES[name]
        q = W.to_oracle(make())
        assert meta["iters"] == iters and (p.N, p.d, p.L, p.N_T) == (q.N, q.d, q.L, q.N_T)
        for a, b in ((p.tlist, q.tlist), (p.pulses, q.pulses), (p.S, q.S), (p.psi0, q.psi0), (p.target, q.target),
                     (np.array(p.H0), np.array(q.H0)), (p.lam, q.lam)):
            assert np.array_equal(np.asarray(a), np.asarray(b))
        with tempfile.TemporaryDirectory() as tmp:  # the committed files are current
            X.export(name, tmp)
            for f in os.listdir(os.path.join(tmp, name)):
                with open(os.path.join(tmp, name, f), "rb") as fa, open(os.path.join(GOLD, "export", name, f), "rb") as fb:
                    assert fa.read() == fb.read(), (name, f)
        a = O.optimize_krotov(p, 2)
        b = O.optimize_krotov(q, 2)
        assert a["J_T"] == b["J_T"] and np.array_equal(a["pulses"], b["pulses"])


@pytest.mark.parametrize("name", _reference_run_vectors() or [None])
def test_oracles_against_reference_run_vectors(name):
    """PINS THE ORACLE once a Julia run exists: tests/golden/julia/<name>.json is the output of the unmodified Krotov.jl
    on tests/golden/export/<name>/ (julia/reference_vectors.jl).  Tolerances of BASELINE.json, with the absolute floor
    two independent Chebyshev implementations have on J_T (5e-13: coefficient rounding, DESIGN.md "parity floor")."""
    if name is None:
        pytest.skip("no vectors from a Julia run of the reference are committed yet (parity unpinned, DESIGN.md section 2)")
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    import export_problem as X

    with open(os.path.join(GOLD, "julia", name + ".json")) as fh:
        g = json.load(fh)
    meta, p = X.load(os.path.join(GOLD, "export", name))
    for h in (O.optimize_krotov(p, meta["iters"]), C.optimize_krotov_c(p, meta["iters"])):
        ref = np.array(g["J_T"])
        assert np.all(np.abs(np.array(h["J_T"]) - ref) <= 1e-10 * np.abs(ref) + 5e-13)
        assert np.abs(np.asarray(h["pulses"]) - np.array(g["pulses"])).max() <= 1e-9


def test_oracles_against_exact_propagator_vector_of_the_two_transmon_problem():
    """C3's generator (two 5-level transmons, d = 25, two controls of which one starts at zero, 4 logical-basis
    trajectories, J_T_sm) on a 40-step grid with C3's time step: tests/golden/c3_two_transmon_g41_exact30.json from the
    30-digit exact-propagator loop (mpmath.expm on 25 x 25 matrices; 4.5 minutes to make, tests/golden/make_golden_exact.py
    --c3)."""
    g = gold("c3_two_transmon_g41_exact30")
    make = lambda: W.to_oracle(W.c3_two_transmon(n_grid=41, T=8.0))  # noqa: E731
    for h, tol in ((O.optimize_krotov(make(), 2, "expm"), 5e-14), (O.optimize_krotov(make(), 2, "cheby"), 1e-12),
                   (C.optimize_krotov_c(make(), 2), 1e-12)):
        assert np.abs(np.array(h["J_T"]) - np.array(g["J_T"])).max() < tol
        assert np.abs(np.asarray(h["pulses"]) - np.array(g["pulses"])).max() < tol
        assert np.abs(np.array(h["g_a_int"]) - np.array(g["g_a_int"])).max() < tol
