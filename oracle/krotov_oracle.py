"""CPU oracle for the Krotov iteration  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``krotov.jl_b200``) never imports, links or executes anything under ``oracle/``.

PARITY UNPINNED.  The reference (JuliaQuantumControl/Krotov.jl) is pure Julia and
Julia is not installed in this image, so this restatement could not be checked
against outputs of the reference itself; the reference's own tests hold no
numeric golden vector for this path (only the inequalities of
``test/test_tls_optimization.jl:66-67``, which ``tests/`` asserts).  The
arithmetic the reference delegates to QuantumPropagators.jl / QuantumControl.jl
(un-vendored, compat ``QuantumControl >= 0.11.1``, ``Project.toml:19``; no
Manifest) is restated from its published algorithm.
What does pin the LOOP restated here independently of this file: ``tests/mp_reference.py`` (the same loop in 40-50
digit arithmetic with exact interval propagators, on the reference's two-level test problem and on small general
problems incl. several trajectories / generators / controls and a non-Hermitian generator); see DESIGN.md section 2.

What is restated, with the reference lines each function follows:

=====================================  =========================================
``krotov_initial_fw_prop``             ``src/optimize.jl:247-265``
``krotov_iteration``                   ``src/optimize.jl:279-371``
``transform_control_ranges``           ``src/optimize.jl:238-244``
``update_result``                      ``src/optimize.jl:374-397``
``optimize_krotov``                    ``src/optimize.jl:161-235`` (loop only)
``ChebyPropagator`` / ``ExpPropagator``  QuantumPropagators ``init_prop`` /
                                       ``reinit_prop!`` / ``prop_step!`` as called
                                       at ``src/optimize.jl:251,257,306,309,324,361``
``discretize*``                        ``src/workspace.jl:102,119,123``,
                                       ``src/optimize.jl:404``
``chi_*`` / ``J_T_*`` / ``taus``       ``src/workspace.jl:171-173``,
                                       ``src/optimize.jl:299-301,381-386``
=====================================  =========================================

Everything is Float64 / ComplexF64, serial, and keeps the reference's loop order
(overlaps: ``l`` outer, ``k`` inner; one lazy operator term at a time inside the
propagator).  Plain NumPy; meant for cases that finish in seconds (C1-C3 and
cut-down C4/C5).  ``krotov_oracle.c`` is the same algorithm in C for the large
CPU-baseline runs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
from scipy.linalg import expm
from scipy.special import jv

__all__ = [
    "ProblemArrays",
    "blackman",
    "flattop",
    "discretize",
    "discretize_on_midpoints",
    "cheby_coeffs",
    "specrange_diag",
    "ChebyPropagator",
    "ExpPropagator",
    "transform_control_ranges",
    "krotov_initial_fw_prop",
    "sigma_on_intervals",
    "numerical_estimate_A",
    "krotov_iteration",
    "optimize_krotov",
    "optimize_krotov_blocked",
    "taus",
    "J_T_value",
    "chi_states",
]


# --------------------------------------------------------------------------------------
# shapes and discretisation
# --------------------------------------------------------------------------------------
def blackman(t, t0, T, a=0.16):
    """Blackman window on [t0, T], zero outside (QuantumControl.Shapes.blackman)."""
    if t < t0 or t > T:
        return 0.0
    x = (t - t0) / (T - t0)
    return 0.5 * (1.0 - a - math.cos(2.0 * math.pi * x) + a * math.cos(4.0 * math.pi * x))


def flattop(t, T, t_rise, t0=0.0, t_fall=None, func="blackman"):
    """Flat-top shape: 0 outside [t0, T], Blackman (or sin^2) switch-on / -off
    (QuantumControl.Shapes.flattop, used at ``test/test_tls_optimization.jl:12``)."""
    if t_fall is None:
        t_fall = t_rise
    if t <= t0 or t >= T:
        return 0.0
    f = 1.0
    if func == "blackman":
        if t <= t0 + t_rise:
            f = blackman(t, t0, t0 + 2.0 * t_rise)
        elif t >= T - t_fall:
            f = blackman(t, T - 2.0 * t_fall, T)
    elif func == "sinsq":
        if t <= t0 + t_rise:
            f = math.sin(math.pi * (t - t0) / (2.0 * t_rise)) ** 2
        elif t >= T - t_fall:
            f = math.sin(math.pi * (t - T) / (2.0 * t_fall)) ** 2
    else:
        raise ValueError(f"unknown func {func!r}")
    return f


def discretize_on_midpoints(control, tlist):
    """N_T values: first at tlist[0], last at tlist[-1], interior at shifted midpoints.

    Follows QuantumPropagators.Controls.discretize_on_midpoints (call sites
    ``src/workspace.jl:102,119,123``): a callable is sampled at ``tlist[0]``, at the
    midpoints of the interior intervals, and at ``tlist[-1]`` (the first and last
    sample sit ON the ends of the time grid, not on the first/last midpoint).
    For a vector of length N_T: a copy (``test/test_pulse_optimization.jl:42``).
    For length N_T+1: ends kept, interior = mean of neighbours."""
    tlist = np.asarray(tlist, dtype=float)
    nt = len(tlist)
    if callable(control):
        vals = np.empty(nt - 1)
        vals[0] = control(tlist[0])
        vals[-1] = control(tlist[-1])
        for i in range(1, nt - 2):
            dt = tlist[i + 1] - tlist[i]
            vals[i] = control(tlist[i] + 0.5 * dt)
        return vals
    arr = np.asarray(control, dtype=float)
    if len(arr) == nt - 1:
        return arr.copy()
    if len(arr) == nt:
        vals = np.empty(nt - 1)
        vals[0] = arr[0]
        vals[-1] = arr[-1]
        for i in range(1, nt - 2):
            vals[i] = 2.0 * arr[i] - vals[i - 1]
        return vals
    raise ValueError("control array length must be len(tlist) or len(tlist)-1")


def discretize(control, tlist):
    """Values ON tlist (N_T+1).  For a midpoint pulse of length N_T: ends kept,
    interior i = mean of pulse values i-1 and i (``src/optimize.jl:404``)."""
    tlist = np.asarray(tlist, dtype=float)
    nt = len(tlist)
    if callable(control):
        return np.array([control(t) for t in tlist], dtype=float)
    arr = np.asarray(control, dtype=float)
    if len(arr) == nt:
        return arr.copy()
    if len(arr) == nt - 1:
        vals = np.empty(nt)
        vals[0] = arr[0]
        vals[-1] = arr[-1]
        for i in range(1, nt - 1):
            vals[i] = 0.5 * (arr[i - 1] + arr[i])
        return vals
    raise ValueError("control array length must be len(tlist) or len(tlist)-1")


# --------------------------------------------------------------------------------------
# materialised problem
# --------------------------------------------------------------------------------------
@dataclass
class ProblemArrays:
    """A control problem reduced to arrays (what oracle, C oracle and GPU all read).

    ``H0[g]`` / ``Hc[g][l]`` are dense (d, d) complex arrays per *generator* g;
    ``gen_of_traj[k]`` maps trajectory k to its generator (ensemble members that
    share a Hamiltonian share g).  ``Hc[g][l] is None`` means the generator does
    not depend on control l (``src/optimize.jl:344``)."""

    tlist: np.ndarray
    H0: List[np.ndarray]
    Hc: List[List[Optional[np.ndarray]]]
    gen_of_traj: np.ndarray
    psi0: np.ndarray  # (N, d) complex
    target: np.ndarray  # (N, d) complex
    pulses: np.ndarray  # (L, N_T) guess pulses on the midpoints
    S: np.ndarray  # (L, N_T) update shapes on the midpoints
    lam: np.ndarray  # (L,)
    weight: Optional[np.ndarray] = None  # (N,), default 1
    functional: str = "sm"  # "sm" | "ss" | "re"
    cheby_limit: float = 1e-12
    specrange_buffer: float = 0.01
    specrange: Optional[tuple] = None  # explicit (E_min, E_max) for every generator
    meta: dict = field(default_factory=dict)
    # Non-linear control amplitudes (src/optimize.jl:268-272, 337-346): the generator is
    #   H0 + sum_l a_l(eps_l(t), t) H_l,   a_l(eps, n) = amp_shape[l][n] * sum_p amp_poly[l][p] eps^p .
    # None = linear controls (a_l = eps_l).  One operator per control.
    amp_poly: Optional[list] = None  # per control: polynomial coefficients, ascending powers (None entry = linear)
    amp_shape: Optional[np.ndarray] = None  # (L, N_T) per-interval factor (a `ShapedAmplitude`), default 1

    def amplitude(self, l, n, eps):
        """a_l(eps, n): the coefficient of H_l on interval n for the control value eps."""
        v = eps
        if self.amp_poly is not None and self.amp_poly[l] is not None:
            v = 0.0
            for c in reversed(self.amp_poly[l]):  # Horner
                v = v * eps + c
        if self.amp_shape is not None:
            v = self.amp_shape[l][n] * v
        return v

    def amplitude_deriv(self, l, n, eps):
        """d a_l / d eps at (eps, n): the factor of H_l in mu_l = dH/d eps_l (`get_control_derivs` + `evaluate`)."""
        v = 1.0
        if self.amp_poly is not None and self.amp_poly[l] is not None:
            c = self.amp_poly[l]
            v = 0.0
            for q in range(len(c) - 1, 0, -1):
                v = v * eps + q * c[q]
        if self.amp_shape is not None:
            v = self.amp_shape[l][n] * v
        return v

    def amplitude_envelope(self, l, eps):
        """Coefficient of H_l used for the spectral envelope at a corner `eps` of the control range: the largest
        |shape| over the grid times the polynomial (UNPINNED convention: upstream evaluates the generator at the range
        corners; with a time-dependent shape the safe choice is its maximum)."""
        v = eps
        if self.amp_poly is not None and self.amp_poly[l] is not None:
            v = 0.0
            for c in reversed(self.amp_poly[l]):
                v = v * eps + c
        if self.amp_shape is not None:
            v = float(np.max(np.abs(self.amp_shape[l]))) * v
        return v

    @property
    def nonlinear(self):
        return self.amp_poly is not None or self.amp_shape is not None

    @property
    def N(self):
        return self.psi0.shape[0]

    @property
    def d(self):
        return self.psi0.shape[1]

    @property
    def L(self):
        return self.pulses.shape[0]

    @property
    def N_T(self):
        return len(self.tlist) - 1

    def weights(self):
        return np.ones(self.N) if self.weight is None else np.asarray(self.weight, float)


# --------------------------------------------------------------------------------------
# Chebyshev propagator
# --------------------------------------------------------------------------------------
def cheby_coeffs(Delta, dt, limit=1e-12):
    """a_0 = J_0(α), a_n = 2 J_n(α), α = |Δ dt / 2|; stop once |a_n| <= limit and n > α."""
    alpha = abs(0.5 * Delta * dt)
    coeffs = [float(jv(0, alpha))]
    eps = abs(coeffs[0])
    i = 1
    while eps > limit or i <= alpha:
        a = 2.0 * float(jv(i, alpha))
        coeffs.append(a)
        eps = abs(a)
        i += 1
    return np.array(coeffs)


def specrange_diag(G):
    """(E_min, E_max) by exact diagonalisation (``specrange_method=:diag``).  Julia's ``eigvals`` of a
    dense matrix dispatches to the Hermitian solver when ``ishermitian(G)``; same here."""
    if np.array_equal(G, G.conj().T):
        ev = np.linalg.eigvalsh(G)
        return float(ev[0]), float(ev[-1])
    ev = np.linalg.eigvals(G)
    return float(ev.real.min()), float(ev.real.max())


def transform_control_ranges(c, eps_min, eps_max, check):
    """``src/optimize.jl:238-244``: widen by 2x when checking, by 5x when (re)setting."""
    if check:
        return (min(eps_min, 2 * eps_min), max(eps_max, 2 * eps_max))
    return (min(eps_min, 5 * eps_min), max(eps_max, 5 * eps_max))


def _identity_ranges(c, eps_min, eps_max, check):
    return (eps_min, eps_max)


class _PWCPropagator:
    """Shared state of a piecewise-constant propagator: generator terms, time grid,
    the *aliased* pulse arrays (``parameters``), direction, current state / index."""

    def __init__(self, H0, Hc, tlist, parameters, backward=False, problem=None):
        self.H0 = H0
        self.Hc = Hc  # list over l, entries may be None
        self.tlist = np.asarray(tlist, float)
        self.parameters = parameters  # list over l of 1-D arrays (aliased, not copied)
        self.backward = backward
        self.state = None
        self.n = 0  # index of the time-grid point the state sits on (0-based)
        # non-linear amplitudes: coefficient of H_l = a_l(eps_l[n], n)  (None: linear controls)
        self.problem = problem if (problem is not None and problem.nonlinear) else None

    def _ops_coeffs(self, n_interval):
        """Lazy Operator(ops, coeffs) = H0 + sum_l eps_l[n] H_l for interval n."""
        ops = [self.H0]
        coeffs = [1.0]
        for l, Hl in enumerate(self.Hc):
            if Hl is not None:
                ops.append(Hl)
                e = self.parameters[l][n_interval]
                coeffs.append(e if self.problem is None else self.problem.amplitude(l, n_interval, e))
        return ops, coeffs

    def _apply(self, ops, coeffs, v):
        """mul!: one term at a time, in order (drift first)."""
        out = np.zeros_like(v)
        for op, c in zip(ops, coeffs):
            out += c * (op @ v)
        return out

    def _evaluate_at(self, vals):
        G = self.H0.copy()
        for l, Hl in enumerate(self.Hc):
            if Hl is not None:
                G = G + (vals[l] if self.problem is None else self.problem.amplitude_envelope(l, vals[l])) * Hl
        return G

    def _reset_time(self):
        self.n = len(self.tlist) - 1 if self.backward else 0

    def _interval_and_dt(self):
        if self.backward:
            ni = self.n - 1
            dt = -(self.tlist[self.n] - self.tlist[self.n - 1])
        else:
            ni = self.n
            dt = self.tlist[self.n + 1] - self.tlist[self.n]
        return ni, dt

    def _advance(self):
        self.n += -1 if self.backward else 1


class ChebyPropagator(_PWCPropagator):
    """QuantumPropagators ``Cheby`` piecewise propagator.  The backward propagator is
    built on the ADJOINT generator (``src/workspace.jl:69,150-160``) with
    ``backward=True`` (negative time step)."""

    def __init__(self, H0, Hc, tlist, parameters, backward=False, limit=1e-12,
                 specrange_buffer=0.01, specrange=None, problem=None):
        super().__init__(H0, Hc, tlist, parameters, backward, problem)
        self.limit = limit
        self.specrange_buffer = specrange_buffer
        self.explicit_specrange = specrange
        # init_prop: un-widened control ranges taken from the current pulse arrays
        self.control_ranges = [(float(np.min(p)), float(np.max(p))) for p in parameters]
        self._set_spectral_envelope()
        self.n_range_updates = 0

    def _set_spectral_envelope(self):
        if self.explicit_specrange is not None:
            E_min, E_max = self.explicit_specrange
        else:
            lo = [r[0] for r in self.control_ranges]
            hi = [r[1] for r in self.control_ranges]
            E_min, E_max = specrange_diag(self._evaluate_at(hi))
            _E_min, _E_max = specrange_diag(self._evaluate_at(lo))
            E_min = min(E_min, _E_min)
            E_max = max(E_max, _E_max)
        Delta = E_max - E_min
        delta = self.specrange_buffer * Delta
        self.E_min = E_min - delta / 2
        self.Delta = Delta + delta
        self.dt = self.tlist[1] - self.tlist[0]
        if self.backward:
            self.dt = -self.dt
        self.coeffs = cheby_coeffs(self.Delta, self.dt, self.limit)

    def reinit_prop(self, state, transform=_identity_ranges):
        self.state = np.array(state, dtype=complex)
        self._reset_time()
        need = False
        for l, ampl in enumerate(self.parameters):
            e_min, e_max = float(np.min(ampl)), float(np.max(ampl))
            c_min, c_max = transform(l, e_min, e_max, True)
            o_min, o_max = self.control_ranges[l]
            if c_min < o_min or c_max > o_max:
                need = True
        if need:
            self.control_ranges = [
                transform(l, float(np.min(a)), float(np.max(a)), False)
                for l, a in enumerate(self.parameters)
            ]
            self._set_spectral_envelope()
            self.n_range_updates += 1

    def prop_step(self):
        ni, dt = self._interval_and_dt()
        if abs(dt - self.dt) > 1e-12 * max(1.0, abs(self.dt)):
            # non-uniform grid: coefficients follow the step actually taken
            self.dt = dt
            self.coeffs = cheby_coeffs(self.Delta, dt, self.limit)
        ops, cf = self._ops_coeffs(ni)
        a = self.coeffs
        Delta, E_min = self.Delta, self.E_min
        beta = Delta / 2 + E_min
        c = -2j / Delta
        if dt < 0:
            c = -c
        v0 = self.state.copy()
        psi = a[0] * v0
        v1 = c * (self._apply(ops, cf, v0) - beta * v0)
        if len(a) > 1:
            psi = psi + a[1] * v1
        c2 = 2 * c
        for i in range(2, len(a)):
            v2 = c2 * (self._apply(ops, cf, v1) - beta * v1) + v0
            psi = psi + a[i] * v2
            v0, v1 = v1, v2
        self.state = np.exp(-1j * beta * dt) * psi
        self._advance()
        return self.state


class ExpPropagator(_PWCPropagator):
    """QuantumPropagators ``ExpProp``: U = exp(-i H_n dt) by dense matrix exponential
    (``test/test_tls_optimization.jl:58``)."""

    def reinit_prop(self, state, transform=None):
        self.state = np.array(state, dtype=complex)
        self._reset_time()

    def prop_step(self):
        ni, dt = self._interval_and_dt()
        ops, cf = self._ops_coeffs(ni)
        H = sum(c * op for op, c in zip(ops, cf))
        self.state = expm(-1j * H * dt) @ self.state
        self._advance()
        return self.state


# --------------------------------------------------------------------------------------
# functionals
# --------------------------------------------------------------------------------------
def taus(states, target):
    """τ_k = <Ψ_k^tgt | Ψ_k(T)>  (``src/optimize.jl:381``)."""
    return np.array([np.vdot(target[k], states[k]) for k in range(len(states))])


def J_T_value(kind, tau, w):
    N = len(tau)
    if kind == "sm":
        F = 0j
        for k in range(N):
            F += w[k] * tau[k]
        F /= N
        return 1.0 - abs(F) ** 2
    if kind == "ss":
        F = 0.0
        for k in range(N):
            F += w[k] * abs(tau[k]) ** 2
        return 1.0 - F / N
    if kind == "re":
        F = 0j
        for k in range(N):
            F += w[k] * tau[k]
        return 1.0 - (F / N).real
    raise ValueError(kind)


def chi_states(kind, tau, w, target):
    """χ_k = -∂J_T/∂<Ψ_k| for the built-in functionals (``make_chi`` analytic forms)."""
    N = len(tau)
    out = np.empty_like(target)
    if kind == "sm":
        s = 0j
        for k in range(N):
            s += w[k] * tau[k]
        for k in range(N):
            out[k] = (w[k] / N**2) * s * target[k]
    elif kind == "ss":
        for k in range(N):
            out[k] = (w[k] / N) * tau[k] * target[k]
    elif kind == "re":
        for k in range(N):
            out[k] = (w[k] / (2 * N)) * target[k]
    else:
        raise ValueError(kind)
    return out


# --------------------------------------------------------------------------------------
# workspace + hot path
# --------------------------------------------------------------------------------------
class OracleWrk:
    """The slice of ``KrotovWrk`` (``src/workspace.jl:30-62``) the hot path touches."""

    def __init__(self, p: ProblemArrays, prop_method="cheby", store_fw=True):
        self.p = p
        N, d, L, N_T = p.N, p.d, p.L, p.N_T
        self.tlist = np.asarray(p.tlist, float)
        self.pulses0 = [np.array(p.pulses[l], float) for l in range(L)]
        self.pulses1 = [a.copy() for a in self.pulses0]  # src/workspace.jl:125
        self.g_a_int = np.zeros(L)
        self.update_shapes = [np.asarray(p.S[l], float) for l in range(L)]
        self.lambda_vals = np.asarray(p.lam, float)
        self.fw_storage = [np.zeros((d, N_T + 1), complex) for _ in range(N)] if store_fw else [None] * N
        self.bw_storage = [np.zeros((d, N_T + 1), complex) for _ in range(N)]
        self.control_derivs = [[p.Hc[p.gen_of_traj[k]][l] for l in range(L)] for k in range(N)]
        self.fw_propagators = []
        self.bw_propagators = []
        for k in range(N):
            g = p.gen_of_traj[k]
            H0, Hc = p.H0[g], p.Hc[g]
            H0a = H0.conj().T
            Hca = [None if h is None else h.conj().T for h in Hc]
            if prop_method == "cheby":
                kw = dict(limit=p.cheby_limit, specrange_buffer=p.specrange_buffer, specrange=p.specrange, problem=p)
                self.fw_propagators.append(ChebyPropagator(H0, Hc, self.tlist, self.pulses0, False, **kw))
                self.bw_propagators.append(ChebyPropagator(H0a, Hca, self.tlist, self.pulses0, True, **kw))
            elif prop_method == "expm":
                self.fw_propagators.append(ExpPropagator(H0, Hc, self.tlist, self.pulses0, False, p))
                self.bw_propagators.append(ExpPropagator(H0a, Hca, self.tlist, self.pulses0, True, p))
            else:
                raise ValueError(prop_method)
        self.tau_vals = np.zeros(N, complex)
        self.J_T = 0.0
        self.J_T_prev = 0.0
        # second order only (`sigma`; TODO at src/optimize.jl:187 "if sigma, fw_storage0 = fw_storage"): the forward
        # trajectory of the PREVIOUS iteration, column n = Psi^(i)(t_n) -- a separate pair of arrays because the
        # iteration writes `fw_storage` shifted by one slot (sic, :367)
        self.fw_storage0 = None
        self.fw_storage1 = None

    def enable_second_order(self):
        N, d, N_T = self.p.N, self.p.d, self.p.N_T
        self.fw_storage0 = [np.zeros((d, N_T + 1), complex) for _ in range(N)]
        self.fw_storage1 = [np.zeros((d, N_T + 1), complex) for _ in range(N)]


def krotov_initial_fw_prop(eps0, phi_k, k, wrk: OracleWrk):
    """``src/optimize.jl:247-265``."""
    for prop in wrk.fw_propagators:
        prop.parameters = eps0
    wrk.fw_propagators[k].reinit_prop(phi_k, transform_control_ranges)
    Phi0 = wrk.fw_storage[k]
    if Phi0 is not None:
        Phi0[:, 0] = phi_k
    prev = None if wrk.fw_storage0 is None else wrk.fw_storage0[k]
    if prev is not None:
        prev[:, 0] = phi_k
    N_T = len(wrk.tlist) - 1
    for n in range(N_T):
        psi = wrk.fw_propagators[k].prop_step()
        if Phi0 is not None:
            Phi0[:, n + 1] = psi
        if prev is not None:
            prev[:, n + 1] = psi


def krotov_iteration(wrk: OracleWrk, eps_i, eps_ip1, chi=None, sigma_vals=None, reduce_du=None):
    """``src/optimize.jl:279-371``, same loop order, same storage slots.

    ``sigma_vals`` (one value per time interval) switches on the second-order update the reference documents
    (``src/optimize.jl:104-105``) but leaves as TODOs (``:187, :350, :369``).  It is restated from the published
    algorithm -- Reich, Ndong, Koch, J. Chem. Phys. 136, 104103 (2012), Eq. (33), as implemented by the `krotov`
    Python package's ``optimize_pulses`` (``second_order`` branch): the overlap of interval n becomes

        <chi_k(t_n)| mu |Psi_k(t_n)>  +  (sigma_n / 2) <Psi_k^(i+1)(t_n) - Psi_k^(i)(t_n)| mu |Psi_k^(i+1)(t_n)> ,

    with Psi^(i) the forward trajectory of the previous iteration (``wrk.fw_storage0``).

    ``reduce_du`` (tests of the multi-rank host logic): when this workspace holds only a shard of the trajectories, the
    callable sums the per-step overlap vector over all shards -- the one cross-rank dependency of a time step
    (``src/optimize.jl:340-349`` sums over ALL k)."""
    p = wrk.p
    tlist = wrk.tlist
    N_T = len(tlist) - 1
    N = p.N
    L = p.L
    X = wrk.bw_storage
    Phi = wrk.fw_storage
    w = p.weights()

    # backward propagation (:297-317)
    Psi = [prop.state for prop in wrk.fw_propagators]
    if chi is None:
        chi_k = chi_states(p.functional, wrk.tau_vals, w, p.target)
    else:
        chi_k = chi(Psi)
    chi_k = [np.array(c, complex) for c in chi_k]
    for k in range(N):
        wrk.bw_propagators[k].parameters = eps_i
        wrk.bw_propagators[k].reinit_prop(chi_k[k], transform_control_ranges)
        X[k][:, N_T] = chi_k[k]
        for n in range(N_T - 1, -1, -1):
            c = wrk.bw_propagators[k].prop_step()
            X[k][:, n] = c

    # pulse update and forward propagation (:321-370)
    for k in range(N):
        wrk.fw_propagators[k].parameters = eps_ip1
        wrk.fw_propagators[k].reinit_prop(p.psi0[k], transform_control_ranges)

    wrk.g_a_int[:] = 0.0
    for n in range(N_T):
        dt = tlist[n + 1] - tlist[n]
        for k in range(N):
            chi_k[k] = X[k][:, n].copy()
        du = np.zeros(L)
        for l in range(L):
            # _eval_mu (:268-276): mu_l = dH/d eps_l evaluated at eps^(i)_n, the GUESS value (:337); for a linear control
            # it is the static operator H_l, for a non-linear amplitude a_l'(eps^(i)_l[n], n) H_l
            fac = p.amplitude_deriv(l, n, eps_i[l][n]) if p.nonlinear else None
            for k in range(N):
                psi_k = wrk.fw_propagators[k].state
                mu = wrk.control_derivs[k][l]
                if mu is not None:
                    ov = np.vdot(chi_k[k], mu @ psi_k)
                    if sigma_vals is not None:  # second-order contribution (TODO at src/optimize.jl:350)
                        ov = ov + 0.5 * sigma_vals[n] * np.vdot(psi_k - wrk.fw_storage0[k][:, n], mu @ psi_k)
                    if fac is None:
                        du[l] += ov.imag
                    else:
                        du[l] += (fac * ov).imag
        if reduce_du is not None:
            du = np.asarray(reduce_du(du), float)
        for l in range(L):
            alpha = wrk.update_shapes[l][n] / wrk.lambda_vals[l]
            d_eps = alpha * du[l]
            eps_ip1[l][n] = eps_i[l][n] + d_eps
            wrk.g_a_int[l] += alpha * abs(du[l]) ** 2 * dt
        for k in range(N):
            if sigma_vals is not None and n == 0:
                wrk.fw_storage1[k][:, 0] = wrk.fw_propagators[k].state
            psi_k = wrk.fw_propagators[k].prop_step()
            if Phi[k] is not None:
                Phi[k][:, n] = psi_k  # sic: slot n (src/optimize.jl:367)
            if sigma_vals is not None:
                wrk.fw_storage1[k][:, n + 1] = psi_k
    if sigma_vals is not None:  # this iteration's trajectory is the next one's Psi^(i)
        wrk.fw_storage0, wrk.fw_storage1 = wrk.fw_storage1, wrk.fw_storage0


def update_result(wrk: OracleWrk):
    """``src/optimize.jl:374-386`` (τ and J_T only)."""
    p = wrk.p
    wrk.J_T_prev = wrk.J_T
    states = [prop.state for prop in wrk.fw_propagators]
    wrk.tau_vals = taus(states, p.target)
    wrk.J_T = J_T_value(p.functional, wrk.tau_vals, p.weights())
    return states


def sigma_on_intervals(sigma, tlist):
    """sigma(t) sampled like the pulses (``discretize_on_midpoints``): one value per time interval.  A number is a
    time-independent sigma."""
    if sigma is None:
        return None
    if callable(sigma):
        return np.asarray(discretize_on_midpoints(lambda t: float(sigma(t)), tlist), float)
    return np.full(len(tlist) - 1, float(sigma))


def numerical_estimate_A(psi_new, psi_old, chi, delta_J_T):
    """The estimate of the second-order constant A of Reich et al. (2012), Eq. (36), from one iteration, as the `krotov`
    Python package's ``second_order.numerical_estimate_A`` forms it:
        A = [ sum_k 2 Re<chi_k(T)|dPsi_k(T)> + dJ_T ] / sum_k |dPsi_k(T)|^2 ,   dPsi = Psi^(i+1)(T) - Psi^(i)(T)."""
    dpsi = [np.asarray(a) - np.asarray(b) for a, b in zip(psi_new, psi_old)]
    den = float(sum(np.vdot(x, x).real for x in dpsi))
    if den <= 1e-30:
        return 0.0
    return (sum(2.0 * np.vdot(c, x).real for c, x in zip(chi, dpsi)) + delta_J_T) / den


def optimize_krotov(p: ProblemArrays, iter_stop=5, prop_method="cheby",
                    callback: Optional[Callable] = None, store_fw=False, sigma=None):
    """The loop of ``src/optimize.jl:161-235`` with the bookkeeping the parity tests
    compare: per-iteration J_T, ∫g_a dt, τ, and the final pulses.

    ``sigma``: None (first order), a number, a callable ``sigma(t)``, optionally with a method
    ``refresh(forward_states=, forward_states0=, chi_states=, J_T=, J_T_prev=)`` called at the end of every iteration
    (the "update sigma" TODO at ``src/optimize.jl:369``) with the final-time states of this and the previous
    iteration and the chi(T) the iteration started from."""
    wrk = OracleWrk(p, prop_method=prop_method, store_fw=store_fw)
    if sigma is not None:
        wrk.enable_second_order()
    eps_i, eps_ip1 = wrk.pulses0, wrk.pulses1
    for k in range(p.N):
        krotov_initial_fw_prop(eps_i, p.psi0[k], k, wrk)
    states = update_result(wrk)
    hist = dict(J_T=[wrk.J_T], g_a_int=[], tau=[wrk.tau_vals.copy()], m_fw=[], m_bw=[], sigma=[])
    if callback is not None:
        callback(wrk, 0, eps_ip1, eps_i)
    for i in range(1, iter_stop + 1):
        sv = sigma_on_intervals(sigma, wrk.tlist)
        if sigma is not None:
            hist["sigma"].append(sv.copy())
            psi_old = [np.array(s) for s in states]
            chi_T = [np.array(c, complex) for c in chi_states(p.functional, wrk.tau_vals, p.weights(), p.target)]
        krotov_iteration(wrk, eps_i, eps_ip1, sigma_vals=sv)
        states = update_result(wrk)
        if sigma is not None and hasattr(sigma, "refresh"):
            sigma.refresh(forward_states=[np.array(s) for s in states], forward_states0=psi_old, chi_states=chi_T,
                          J_T=wrk.J_T, J_T_prev=wrk.J_T_prev)
        hist["J_T"].append(wrk.J_T)
        hist["g_a_int"].append(wrk.g_a_int.copy())
        hist["tau"].append(wrk.tau_vals.copy())
        if prop_method == "cheby":
            hist["m_fw"].append([len(q.coeffs) for q in wrk.fw_propagators])
            hist["m_bw"].append([len(q.coeffs) for q in wrk.bw_propagators])
        if callback is not None:
            callback(wrk, i, eps_ip1, eps_i)
        eps_i, eps_ip1 = eps_ip1, eps_i
    hist["pulses"] = np.array(eps_i)
    hist["optimized_controls"] = np.array([discretize(e, p.tlist) for e in eps_i])
    hist["states"] = np.array(states)
    hist["wrk"] = wrk
    return hist


def optimize_krotov_blocked(p: ProblemArrays, iter_stop=5):
    """The same algorithm as :func:`optimize_krotov` with the trajectories of one generator held as the COLUMNS of one
    (d, B) state block, so that a propagator term is one ``op @ block`` (BLAS-3) instead of B matrix-vector products.
    For dense generators with many trajectories (BASELINE config 5 at d = 4096, where the per-trajectory loop would
    stream the generator from memory once per trajectory and Chebyshev term).  Every column goes through exactly the
    arithmetic of its own :class:`ChebyPropagator` -- same polynomial, one lazy operator term at a time -- and the
    overlap sum keeps the reference's order (``l`` outer, ``k`` inner, ``src/optimize.jl:340-349``); only BLAS's
    summation order inside a matrix product differs from the per-trajectory oracle (checked in tests/test_oracle.py)."""
    N, L, N_T = p.N, p.L, p.N_T
    tlist = np.asarray(p.tlist, float)
    w = p.weights()
    groups = [np.nonzero(np.asarray(p.gen_of_traj) == g)[0] for g in range(len(p.H0))]
    pulses0 = [np.array(p.pulses[l], float) for l in range(L)]
    pulses1 = [a.copy() for a in pulses0]
    kw = dict(limit=p.cheby_limit, specrange_buffer=p.specrange_buffer, specrange=p.specrange, problem=p)
    fw = [ChebyPropagator(p.H0[g], p.Hc[g], tlist, pulses0, False, **kw) for g in range(len(groups))]
    bw = [ChebyPropagator(p.H0[g].conj().T, [None if h is None else h.conj().T for h in p.Hc[g]], tlist, pulses0,
                          True, **kw) for g in range(len(groups))]
    X = [np.zeros((N_T + 1, p.d, len(ks)), complex) for ks in groups]  # bw_storage, slot n = time-grid point n

    def block(rows, ks):  # (N, d) rows -> (d, B) columns of one group
        return np.ascontiguousarray(np.asarray(rows)[ks].T)

    def final_states():
        out = np.empty((N, p.d), complex)
        for g, ks in enumerate(groups):
            out[ks] = fw[g].state.T
        return out

    def sweep_forward(eps):  # krotov_initial_fw_prop! for every k (src/optimize.jl:247-265)
        for g, ks in enumerate(groups):
            fw[g].parameters = eps
            fw[g].reinit_prop(block(p.psi0, ks), transform_control_ranges)
            for _ in range(N_T):
                fw[g].prop_step()

    eps_i, eps_ip1 = pulses0, pulses1
    sweep_forward(eps_i)
    states = final_states()
    tau = taus(states, p.target)
    hist = dict(J_T=[J_T_value(p.functional, tau, w)], g_a_int=[], tau=[tau.copy()])
    for _ in range(1, iter_stop + 1):
        # ---- src/optimize.jl:297-317
        chi = chi_states(p.functional, tau, w, p.target)
        for g, ks in enumerate(groups):
            bw[g].parameters = eps_i
            bw[g].reinit_prop(block(chi, ks), transform_control_ranges)
            X[g][N_T] = bw[g].state
            for n in range(N_T - 1, -1, -1):
                X[g][n] = bw[g].prop_step()
        # ---- :321-370
        for g, ks in enumerate(groups):
            fw[g].parameters = eps_ip1
            fw[g].reinit_prop(block(p.psi0, ks), transform_control_ranges)
        g_a_int = np.zeros(L)
        for n in range(N_T):
            dt = tlist[n + 1] - tlist[n]
            du = np.zeros(L)
            for l in range(L):
                contrib = np.zeros(N)
                for g, ks in enumerate(groups):
                    mu = p.Hc[g][l]
                    if mu is not None:
                        fac = p.amplitude_deriv(l, n, eps_i[l][n]) if p.nonlinear else 1.0
                        contrib[ks] = (fac * np.einsum("ik,ik->k", X[g][n].conj(), mu @ fw[g].state)).imag
                for k in range(N):  # reference order: k ascending
                    du[l] += contrib[k]
            for l in range(L):
                alpha = p.S[l][n] / p.lam[l]
                eps_ip1[l][n] = eps_i[l][n] + alpha * du[l]
                g_a_int[l] += alpha * abs(du[l]) ** 2 * dt
            for g in range(len(groups)):
                fw[g].prop_step()
        states = final_states()
        tau = taus(states, p.target)
        hist["J_T"].append(J_T_value(p.functional, tau, w))
        hist["g_a_int"].append(g_a_int.copy())
        hist["tau"].append(tau.copy())
        eps_i, eps_ip1 = eps_ip1, eps_i
    hist["pulses"] = np.array(eps_i)
    hist["states"] = states
    hist["m"] = (len(fw[0].coeffs), len(bw[0].coeffs))
    return hist
