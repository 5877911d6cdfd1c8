"""Second-order Krotov (the ``sigma`` keyword of ``optimize``, ``src/optimize.jl:104-105``).

The reference documents ``sigma`` and leaves the implementation as TODOs (``src/optimize.jl:187, 350, 369``;
``src/workspace.jl:127-130``).  What is built here follows the published algorithm -- Reich, Ndong, Koch, J. Chem.
Phys. 136, 104103 (2012), Eq. (33), and the interface of the `krotov` Python package (``krotov.second_order``): the
overlap that drives the update of time interval n becomes

    <chi_k(t_n)| mu |Psi_k(t_n)>  +  (sigma/2) <Psi_k^(i+1)(t_n) - Psi_k^(i)(t_n)| mu |Psi_k^(i+1)(t_n)> .

How the device path computes it.  For Hermitian control operators ``Im <Psi|mu|Psi> = 0``, so the update is the
first-order update with ``chi_eff(t_n) = chi(t_n) - (sigma/2) Psi^(i)(t_n)``.  For Hermitian generators the backward
propagator under the guess pulses is the inverse of the forward propagator that produced ``Psi^(i)``, so for a
sigma that is constant over the time grid (the standard choice ``sigma = -max(eps_A, 2 A + eps_A)``, re-estimated
once per iteration)

    chi_eff(t_n) = U^dagger(T -> t_n) [ chi(T) - (sigma/2) Psi^(i)(T) ]                  for every n:

the whole second-order contribution is a change of the boundary condition of the backward sweep.  The kernels, the
storage and the multi-GPU exchange are those of the first-order iteration; no second forward storage is needed
(what the reference's TODOs reserve ``fw_storage2`` for) and the iteration costs what it cost before.  The identity
holds to the accuracy of the propagator (1e-12 per step here); ``tests/`` compare with an oracle that evaluates the
general formula from a stored previous trajectory.

Not served (``ArgumentError``): a sigma that varies over the time grid, non-Hermitian generators or control
operators (dissipative dynamics) -- both need the previous trajectory inside the sweep.
"""
from __future__ import annotations

import numpy as np

from .controls import discretize_on_midpoints
from .errors import ArgumentError

__all__ = ["Sigma", "NumericalSigma", "numerical_estimate_A", "sigma_value"]


class Sigma:
    """Base class of a second-order function: callable ``sigma(t)``, optionally ``refresh(**info)`` once per iteration
    (the "update sigma" TODO at ``src/optimize.jl:369``).  ``info`` holds ``forward_states`` / ``forward_states0`` (the
    final-time states of this and of the previous iteration, row k = trajectory k), ``chi_states`` (chi(T) the iteration
    started from, before the fold),
    ``J_T``, ``J_T_prev``, ``optimized_pulses``, ``guess_pulses``, ``trajectories``, ``result``."""

    def __call__(self, t):
        raise NotImplementedError

    def refresh(self, **info):
        pass


def numerical_estimate_A(forward_states, forward_states0, chi_states, delta_J_T):
    """Estimate of the constant A of the second-order construction from one iteration (Reich et al. 2012, Eq. (36)):
    ``A = [sum_k 2 Re<chi_k(T)|dPsi_k(T)> + dJ_T] / sum_k |dPsi_k(T)|^2`` with ``dPsi = Psi^(i+1)(T) - Psi^(i)(T)``."""
    dpsi = np.asarray(forward_states, np.complex128) - np.asarray(forward_states0, np.complex128)  # (N, d)
    den = float(np.vdot(dpsi, dpsi).real)
    if den <= 1e-30:
        return 0.0
    num = 2.0 * float(np.vdot(np.asarray(chi_states, np.complex128), dpsi).real) + delta_J_T
    return num / den


class NumericalSigma(Sigma):
    """``sigma(t) = -max(eps_A, 2 A + eps_A)`` with A re-estimated after every iteration by ``numerical_estimate_A``."""

    def __init__(self, A, eps_A=0.0):
        self.A = float(A)
        self.eps_A = float(eps_A)

    def __call__(self, t):
        return -max(self.eps_A, 2.0 * self.A + self.eps_A)

    def refresh(self, *, forward_states, forward_states0, chi_states, J_T, J_T_prev, **_):
        self.A = numerical_estimate_A(forward_states, forward_states0, chi_states, J_T - J_T_prev)


def sigma_value(sigma, tlist, full=True):
    """The one value a sigma takes over the time grid (sampled like the pulses, on the interval midpoints).
    ``full=False`` samples the first, a middle and the last interval only: the per-iteration re-check after
    ``sigma.refresh`` (the full grid was checked when the workspace was built)."""
    if not callable(sigma):
        vals = np.array([float(sigma)])
    elif full:
        vals = np.asarray(discretize_on_midpoints(lambda t: float(sigma(t)), tlist), np.float64)
    else:
        n = len(tlist) - 1
        vals = np.array([float(sigma(tlist[0])), float(sigma(0.5 * (tlist[n // 2] + tlist[n // 2 + 1]))) if n > 2
                         else float(sigma(tlist[0])), float(sigma(tlist[-1]))])
    if not np.all(np.isfinite(vals)):
        raise ArgumentError("sigma(t) is not finite on the time grid")
    if np.ptp(vals) != 0.0:
        raise ArgumentError(
            "sigma(t) varies over the time grid: the device path folds a time-independent sigma into the boundary "
            "condition of the backward sweep (see krotov_jl_b200.second_order); re-estimate it per iteration in "
            "`sigma.refresh` instead")
    return float(vals[0])
