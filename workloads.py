"""Synthetic workloads for the BASELINE.json configs C1-C5.

Plain NumPy: no dependency on the product package or on ``oracle/``.  A
``Workload`` is the *materialised* control problem -- dense generator terms per
distinct generator, trajectory -> generator map, initial / target states, guess
controls and update shapes as callables -- from which both the product API
(``krotov_jl_b200.ControlProblem``) and the oracle's ``ProblemArrays`` are built.
Everything is deterministic (fixed ``numpy.random.default_rng`` seeds).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

TWO_PI = 2.0 * math.pi


@dataclass
class Workload:
    name: str
    tlist: np.ndarray
    H0: List[np.ndarray]  # per generator, (d, d) complex
    Hc: List[List[Optional[np.ndarray]]]  # per generator, per control
    gen_of_traj: np.ndarray  # (N,) int
    psi0: np.ndarray  # (N, d)
    target: np.ndarray  # (N, d)
    controls: List[Callable[[float], float]]  # guess controls ε_l(t)
    update_shape: Callable[[float], float]
    lambda_a: float
    functional: str  # "sm" | "ss" | "re"
    specrange: Optional[tuple] = None
    meta: dict = field(default_factory=dict)
    # non-linear control amplitudes: control l enters as shape_l(t) * sum_p amp_poly[l][p] eps^p (None = eps itself)
    amp_poly: Optional[list] = None  # per control: coefficients (ascending) or None
    amp_shape: Optional[list] = None  # per control: callable t -> factor, or None

    @property
    def N(self):
        return self.psi0.shape[0]

    @property
    def d(self):
        return self.psi0.shape[1]

    @property
    def L(self):
        return len(self.controls)

    @property
    def N_T(self):
        return len(self.tlist) - 1


# ---- shapes (same definitions as QuantumControl.Shapes; restated independently) ---------
def _blackman(t, t0, T, a=0.16):
    if t < t0 or t > T:
        return 0.0
    x = (t - t0) / (T - t0)
    return 0.5 * (1.0 - a - math.cos(TWO_PI * x) + a * math.cos(2 * TWO_PI * x))


def flattop(t, T, t_rise, t0=0.0):
    if t <= t0 or t >= T:
        return 0.0
    if t <= t0 + t_rise:
        return _blackman(t, t0, t0 + 2 * t_rise)
    if t >= T - t_rise:
        return _blackman(t, T - 2 * t_rise, T)
    return 1.0


# ---- C1: two-level system of test/test_tls_optimization.jl:12-63 ------------------------
def c1_tls(n_grid=501):
    sz = np.array([[1, 0], [0, -1]], complex)
    sx = np.array([[0, 1], [1, 0]], complex)
    tlist = np.linspace(0.0, 5.0, n_grid)
    return Workload(
        name="C1-tls",
        tlist=tlist,
        H0=[-0.5 * sz],
        Hc=[[sx]],
        gen_of_traj=np.zeros(1, int),
        psi0=np.array([[1, 0]], complex),
        target=np.array([[0, 1]], complex),
        controls=[lambda t: 0.2 * flattop(t, T=5.0, t_rise=0.3)],
        update_shape=lambda t: 1.0,
        lambda_a=1.0,
        functional="sm",
    )


# ---- C2: single transmon X gate, 3 levels ------------------------------------------------
def c2_transmon_x(n_grid=501, levels=3):
    alpha = -TWO_PI * 0.3
    n = np.arange(levels)
    H0 = np.diag(0.5 * alpha * n * (n - 1)).astype(complex)
    b = np.diag(np.sqrt(np.arange(1, levels)), 1).astype(complex)
    H1 = 0.5 * (b + b.conj().T)
    T = 50.0
    tlist = np.linspace(0.0, T, n_grid)
    psi0 = np.zeros((2, levels), complex)
    tgt = np.zeros((2, levels), complex)
    psi0[0, 0] = 1
    tgt[0, 1] = 1
    psi0[1, 1] = 1
    tgt[1, 0] = 1
    amp = TWO_PI * 0.05
    return Workload(
        name="C2-transmon-x",
        tlist=tlist,
        H0=[H0],
        Hc=[[H1]],
        gen_of_traj=np.zeros(2, int),
        psi0=psi0,
        target=tgt,
        controls=[lambda t: amp * flattop(t, T=T, t_rise=5.0)],
        update_shape=lambda t: flattop(t, T=T, t_rise=5.0),
        lambda_a=10.0,
        functional="sm",
    )


# ---- C3 / C4: two coupled transmons ------------------------------------------------------
def _two_transmon_terms(w1, w2, wd, a1, a2, J, lam, levels=5):
    n = levels
    b = np.diag(np.sqrt(np.arange(1, n)), 1).astype(complex)
    I = np.eye(n, dtype=complex)
    b1 = np.kron(b, I)
    b2 = np.kron(I, b)
    num = np.diag(np.arange(n)).astype(complex)
    anh = np.diag(np.arange(n) * (np.arange(n) - 1) / 2.0).astype(complex)
    H0 = ((w1 - wd) * np.kron(num, I) + a1 * np.kron(anh, I)
          + (w2 - wd) * np.kron(I, num) + a2 * np.kron(I, anh)
          + J * (b1.conj().T @ b2 + b1 @ b2.conj().T))
    Hre = 0.5 * (b1 + b1.conj().T + lam * (b2 + b2.conj().T))
    Him = 0.5j * (b1.conj().T - b1 + lam * (b2.conj().T - b2))
    return H0, Hre, Him


_C3 = dict(w1=TWO_PI * 4.380, w2=TWO_PI * 4.614, wd=TWO_PI * 4.498,
           a1=-TWO_PI * 0.210, a2=-TWO_PI * 0.215, J=-TWO_PI * 0.003, lam=1.03)

_SQRT_ISWAP = np.array([[1, 0, 0, 0],
                        [0, 1 / math.sqrt(2), 1j / math.sqrt(2), 0],
                        [0, 1j / math.sqrt(2), 1 / math.sqrt(2), 0],
                        [0, 0, 0, 1]], complex)
_CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], complex)


def _logical_states(levels, gate):
    d = levels * levels
    idx = [0 * levels + 0, 0 * levels + 1, 1 * levels + 0, 1 * levels + 1]
    psi0 = np.zeros((4, d), complex)
    tgt = np.zeros((4, d), complex)
    for a in range(4):
        psi0[a, idx[a]] = 1.0
        for bb in range(4):
            tgt[a, idx[bb]] = gate[bb, a]
    return psi0, tgt


def c3_two_transmon(n_grid=2001, levels=5, T=400.0):
    H0, Hre, Him = _two_transmon_terms(levels=levels, **_C3)
    psi0, tgt = _logical_states(levels, _SQRT_ISWAP)
    amp = TWO_PI * 0.035
    return Workload(
        name="C3-two-transmon-sqrt-iswap",
        tlist=np.linspace(0.0, T, n_grid),
        H0=[H0],
        Hc=[[Hre, Him]],
        gen_of_traj=np.zeros(4, int),
        psi0=psi0,
        target=tgt,
        controls=[lambda t: amp * flattop(t, T=T, t_rise=20.0), lambda t: 0.0],
        update_shape=lambda t: flattop(t, T=T, t_rise=20.0),
        lambda_a=1.0,
        functional="sm",
    )


def c4_ensemble(n_samples=256, n_grid=2001, levels=5, T=400.0, sigma=0.01, seed=20240607):
    """Robust CNOT over ``n_samples`` perturbed copies of C3's Hamiltonian.  Draws are
    sample-major in the order (w1, w2, a1, a2, J); sample 0 is unperturbed; the full
    256-sample stream is always drawn so a cut-down ensemble is a prefix of C4."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((max(n_samples, 256), 5))
    g[0, :] = 0.0
    H0s, Hcs = [], []
    for s in range(n_samples):
        par = dict(_C3)
        for j, key in enumerate(("w1", "w2", "a1", "a2", "J")):
            par[key] = _C3[key] * (1.0 + sigma * g[s, j])
        H0, Hre, Him = _two_transmon_terms(levels=levels, **par)
        H0s.append(H0)
        Hcs.append([Hre, Him])
    p0, tg = _logical_states(levels, _CNOT)
    psi0 = np.tile(p0, (n_samples, 1))
    tgt = np.tile(tg, (n_samples, 1))
    gen = np.repeat(np.arange(n_samples), 4)
    amp = TWO_PI * 0.035
    return Workload(
        name=f"C4-robust-cnot-{n_samples}x4",
        tlist=np.linspace(0.0, T, n_grid),
        H0=H0s,
        Hc=Hcs,
        gen_of_traj=gen,
        psi0=psi0,
        target=tgt,
        controls=[lambda t: amp * flattop(t, T=T, t_rise=20.0), lambda t: 0.0],
        update_shape=lambda t: flattop(t, T=T, t_rise=20.0),
        lambda_a=1.0,
        functional="sm",
        meta=dict(n_samples=n_samples),
    )


# ---- C5: dense GUE-like generator --------------------------------------------------------
def _gue(rng, d):
    A = (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / math.sqrt(2.0 * d)
    return 0.5 * (A + A.conj().T)


def c5_dense(d=4096, n_traj=64, n_grid=10001, seed=4096, amp=0.5, dt=1.0):
    """Dense random Hermitian drift + control.  Off-diagonal variance 1/(2d) gives a
    semicircle spectrum on about [-sqrt2, sqrt2]; the explicit spectral range
    (E_min, E_max) = -/+ sqrt2 (1 + 5 amp) (1.05) covers H0 + eps H1 for |eps| <= 5 amp
    so that oracle and GPU use identical Chebyshev polynomials without an O(d^3)
    diagonalisation."""
    rng = np.random.default_rng(seed)
    H0 = _gue(rng, d)
    H1 = _gue(rng, d)
    psi0 = np.zeros((n_traj, d), complex)
    for k in range(n_traj):
        psi0[k, k] = 1.0
    tg = rng.standard_normal((n_traj, d)) + 1j * rng.standard_normal((n_traj, d))
    tg /= np.linalg.norm(tg, axis=1, keepdims=True)
    T = dt * (n_grid - 1)
    R = math.sqrt(2.0) * (1.0 + 5.0 * amp) * 1.05
    return Workload(
        name=f"C5-dense-d{d}-n{n_traj}",
        tlist=np.linspace(0.0, T, n_grid),
        H0=[H0],
        Hc=[[H1]],
        gen_of_traj=np.zeros(n_traj, int),
        psi0=psi0,
        target=tg,
        controls=[lambda t: amp * flattop(t, T=T, t_rise=0.1 * T)],
        update_shape=lambda t: flattop(t, T=T, t_rise=0.1 * T),
        lambda_a=10.0,
        functional="ss",
        specrange=(-R, R),
    )


def dummy_dense(d=10, n_traj=2, n_controls=2, n_grid=51, seed=7, functional="ss", hermitian=True):
    """Small random dense problem in the spirit of QuantumControlTestUtils'
    ``dummy_control_problem`` (``test/test_iterations.jl:16-24``), own RNG."""
    rng = np.random.default_rng(seed)

    def rnd():
        A = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        A /= math.sqrt(d)
        return 0.5 * (A + A.conj().T) if hermitian else A

    H0 = rnd()
    Hc = [rnd() for _ in range(n_controls)]
    psi0 = rng.standard_normal((n_traj, d)) + 1j * rng.standard_normal((n_traj, d))
    psi0 /= np.linalg.norm(psi0, axis=1, keepdims=True)
    tg = rng.standard_normal((n_traj, d)) + 1j * rng.standard_normal((n_traj, d))
    tg /= np.linalg.norm(tg, axis=1, keepdims=True)
    T = 5.0
    phases = rng.uniform(0, TWO_PI, n_controls)
    ctrls = [(lambda t, ph=ph: 0.3 * flattop(t, T=T, t_rise=0.5) * math.cos(1.3 * t + ph)) for ph in phases]
    return Workload(
        name=f"dummy-d{d}",
        tlist=np.linspace(0, T, n_grid),
        H0=[H0],
        Hc=[Hc],
        gen_of_traj=np.zeros(n_traj, int),
        psi0=psi0,
        target=tg,
        controls=ctrls,
        update_shape=lambda t: flattop(t, T=T, t_rise=0.5),
        lambda_a=2.0,
        functional=functional,
    )


# ---- adapters ---------------------------------------------------------------------------
def midpoint_samples(f, tlist):
    """Guess control sampled like ``discretize_on_midpoints`` (ends ON the grid ends)."""
    nt = len(tlist)
    v = np.empty(nt - 1)
    v[0] = f(tlist[0])
    v[-1] = f(tlist[-1])
    for i in range(1, nt - 2):
        v[i] = f(tlist[i] + 0.5 * (tlist[i + 1] - tlist[i]))
    return v


def to_oracle(w: Workload):
    """Workload -> oracle.krotov_oracle.ProblemArrays (tests / bench cpu_baseline only)."""
    from oracle.krotov_oracle import ProblemArrays

    pulses = np.array([midpoint_samples(c, w.tlist) for c in w.controls])
    S = np.array([midpoint_samples(w.update_shape, w.tlist) for _ in w.controls])
    amp_poly, amp_shape = None, None
    if w.amp_poly is not None and any(c is not None for c in w.amp_poly):
        amp_poly = [None if c is None else [float(x) for x in c] for c in w.amp_poly]
    if w.amp_shape is not None and any(f is not None for f in w.amp_shape):
        amp_shape = np.array([np.ones(len(w.tlist) - 1) if f is None else midpoint_samples(f, w.tlist) for f in w.amp_shape])
    return ProblemArrays(
        amp_poly=amp_poly, amp_shape=amp_shape,
        tlist=np.asarray(w.tlist, float),
        H0=w.H0,
        Hc=w.Hc,
        gen_of_traj=np.asarray(w.gen_of_traj, int),
        psi0=w.psi0,
        target=w.target,
        pulses=pulses,
        S=S,
        lam=np.full(w.L, float(w.lambda_a)),
        functional=w.functional,
        specrange=w.specrange,
    )


# ---- spin chain: sparse generator on a large Hilbert space ---------------------------------------------------
def spin_chain(n_spins=6, n_traj=8, n_grid=41, T=4.0, seed=12, functional="ss"):
    """XXZ chain with open ends in a longitudinal field, driven by global sigma_x and sigma_y fields:
    d = 2^n_spins, about n_spins + 1 non-zeros per row (sparse; complex entries through sigma_y)."""
    sx = np.array([[0, 1], [1, 0]], complex)
    sy = np.array([[0, -1j], [1j, 0]], complex)
    sz = np.array([[1, 0], [0, -1]], complex)
    I2 = np.eye(2, dtype=complex)

    def op(single, site):
        out = np.array([[1.0 + 0j]])
        for s in range(n_spins):
            out = np.kron(out, single if s == site else I2)
        return out

    d = 2 ** n_spins
    H0 = np.zeros((d, d), complex)
    for i in range(n_spins - 1):
        H0 += 0.25 * (op(sx, i) @ op(sx, i + 1) + op(sy, i) @ op(sy, i + 1)) + 0.15 * op(sz, i) @ op(sz, i + 1)
    for i in range(n_spins):
        H0 += 0.1 * (1 + 0.3 * i) * op(sz, i)
    Hx = sum(op(sx, i) for i in range(n_spins)) * 0.5
    Hy = sum(op(sy, i) for i in range(n_spins)) * 0.5
    rng = np.random.default_rng(seed)
    psi0 = np.zeros((n_traj, d), complex)
    for k in range(n_traj):
        psi0[k, k % d] = 1.0
    tg = rng.standard_normal((n_traj, d)) + 1j * rng.standard_normal((n_traj, d))
    tg /= np.linalg.norm(tg, axis=1, keepdims=True)
    return Workload(
        name=f"spin-chain-{n_spins}",
        tlist=np.linspace(0, T, n_grid),
        H0=[H0],
        Hc=[[Hx, Hy]],
        gen_of_traj=np.zeros(n_traj, int),
        psi0=psi0,
        target=tg,
        controls=[lambda t: 0.4 * flattop(t, T=T, t_rise=0.4), lambda t: 0.1 * flattop(t, T=T, t_rise=0.4) * math.sin(2 * t)],
        update_shape=lambda t: flattop(t, T=T, t_rise=0.4),
        lambda_a=1.0,
        functional=functional,
    )
