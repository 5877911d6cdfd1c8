"""Final-time functionals and their chi constructors (``QuantumControl.Functionals``; call sites
``src/workspace.jl:171-173``, ``src/optimize.jl:299-301,381-386``).

Sign convention: ``chi_k = -dJ_T/d<Psi_k|`` (``src/optimize.jl:100``).  The three analytic chi's
carry a ``krotov_builtin`` tag; when the workspace sees it, chi(T) is formed on the device."""
from __future__ import annotations

import numpy as np

__all__ = ["taus", "J_T_sm", "J_T_ss", "J_T_re", "chi_sm", "chi_ss", "chi_re", "chi_coefficients", "make_chi"]


def taus(states, trajectories, ignore_missing_target_state=False):
    """tau_k = <target_k | Psi_k>; zero where a trajectory has no target (if allowed)."""
    out = np.zeros(len(trajectories), np.complex128)
    for k, (psi, traj) in enumerate(zip(states, trajectories)):
        if traj.target_state is None:
            if not ignore_missing_target_state:
                raise ValueError(f"trajectory {k} has no target_state")
            continue
        out[k] = np.vdot(traj.target_state, psi)
    return out


def _tau_w(states, trajectories, tau):
    if tau is None:
        tau = taus(states, trajectories)
    w = np.array([t.weight for t in trajectories], np.float64)
    return np.asarray(tau, np.complex128), w


def J_T_sm(states, trajectories, tau=None):
    """Square-modulus functional ``1 - |1/N sum_k w_k tau_k|^2``."""
    tau, w = _tau_w(states, trajectories, tau)
    f = np.sum(w * tau) / len(tau)
    return 1.0 - abs(f) ** 2


def J_T_ss(states, trajectories, tau=None):
    """State-to-state functional ``1 - 1/N sum_k w_k |tau_k|^2``."""
    tau, w = _tau_w(states, trajectories, tau)
    return 1.0 - float(np.sum(w * np.abs(tau) ** 2)) / len(tau)


def J_T_re(states, trajectories, tau=None):
    """Real-part functional ``1 - Re[1/N sum_k w_k tau_k]``."""
    tau, w = _tau_w(states, trajectories, tau)
    return 1.0 - float(np.real(np.sum(w * tau))) / len(tau)


def chi_sm(states, trajectories, tau=None):
    tau, w = _tau_w(states, trajectories, tau)
    n = len(tau)
    s = np.sum(w * tau)
    return [(w[k] / n**2) * s * trajectories[k].target_state for k in range(n)]


def chi_ss(states, trajectories, tau=None):
    tau, w = _tau_w(states, trajectories, tau)
    n = len(tau)
    return [(w[k] / n) * tau[k] * trajectories[k].target_state for k in range(n)]


def chi_re(states, trajectories, tau=None):
    tau, w = _tau_w(states, trajectories, tau)
    n = len(tau)
    return [(w[k] / (2 * n)) * trajectories[k].target_state for k in range(n)]


def chi_coefficients(kind, tau, w):
    """c_k with chi_k = c_k |target_k> for the three analytic functionals, all trajectories at once (what the device
    forms in ``chi_coef_kernel``; used on the host where chi(T) is needed as an array, e.g. second_order.py)."""
    tau, w = np.asarray(tau, np.complex128), np.asarray(w, np.float64)
    n = len(tau)
    if kind == "sm":
        return (w / n**2) * np.sum(w * tau)
    if kind == "ss":
        return (w / n) * tau
    if kind == "re":
        return (w / (2 * n)).astype(np.complex128)
    raise ValueError(kind)


chi_sm.krotov_builtin = "sm"
chi_ss.krotov_builtin = "ss"
chi_re.krotov_builtin = "re"
_ANALYTIC = {J_T_sm: chi_sm, J_T_ss: chi_ss, J_T_re: chi_re}


def make_chi(J_T, trajectories, **kwargs):
    """chi for a known functional.  The reference falls back to automatic differentiation for
    arbitrary ``J_T``; that fallback does not exist here -- pass ``chi=...`` explicitly."""
    try:
        return _ANALYTIC[J_T]
    except (KeyError, TypeError):
        raise ValueError(
            "make_chi: no analytic chi is known for this J_T and automatic differentiation is not "
            "available; pass the `chi` keyword argument to `optimize`.") from None
