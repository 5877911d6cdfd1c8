"""Generators: ``hamiltonian(H0, (H1, eps1), ...)`` as in QuantumControl (used by the reference's tests,
``test/test_tls_optimization.jl:16-28``)."""
from __future__ import annotations

import numpy as np

__all__ = ["Generator", "hamiltonian", "ShapedAmplitude", "PolynomialAmplitude"]


class PolynomialAmplitude:
    """A control amplitude that is a NON-LINEAR function of its control: the term enters the generator as
    ``a(eps(t), t) * op`` with ``a(eps, t) = shape(t) * sum_p coeffs[p] * eps**p`` (``coeffs`` ascending, degree
    1..4; ``shape`` a callable or a vector on the intervals of the time grid, default 1).

    The reference handles such terms through ``get_control_derivs`` / ``evaluate`` (``src/optimize.jl:268-272``):
    the derivative ``mu = a'(eps, t) op`` is evaluated at the GUESS pulse of every interval (``:337``).  Here the
    amplitude is handed to the device as its polynomial (``krotov_set_amplitudes``)."""

    def __init__(self, control, coeffs, shape=None):
        coeffs = [float(c) for c in coeffs]
        if not 2 <= len(coeffs) <= 5:
            raise ValueError("PolynomialAmplitude: 2 to 5 coefficients (degree 1 to 4)")
        self.control, self.coeffs, self.shape = control, coeffs, shape

    def __call__(self, eps, t=None):
        v = 0.0
        for c in reversed(self.coeffs):
            v = v * eps + c
        if self.shape is not None and t is not None and callable(self.shape):
            v = self.shape(t) * v
        return v


class ShapedAmplitude(PolynomialAmplitude):
    """``a(t) = shape(t) * eps(t)``: QuantumPropagators' ``ShapedAmplitude`` -- linear in the control, but the
    derivative ``mu = shape(t) op`` depends on time."""

    def __init__(self, control, shape):
        super().__init__(control, [0.0, 1.0], shape)


def _is_sparse(op):
    return hasattr(op, "tocsr") and hasattr(op, "nnz")


def _as_matrix(op):
    if _is_sparse(op):  # scipy.sparse operators stay sparse: large spin chains are never densified
        m = op.tocsr().astype(np.complex128)
        if m.shape[0] != m.shape[1]:
            raise ValueError("operators must be square matrices")
        return m
    m = np.asarray(op, dtype=np.complex128)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("operators must be square matrices")
    return m


class Generator:
    """``G(t) = sum(drift ops) + sum_i amplitudes[i](t) * control_ops[i]`` with linear controls."""

    def __init__(self, drift_ops, control_ops, amplitudes):
        self.drift_ops = [_as_matrix(o) for o in drift_ops]
        self.control_ops = [_as_matrix(o) for o in control_ops]
        self.amplitudes = list(amplitudes)
        if len(self.control_ops) != len(self.amplitudes):
            raise ValueError("one amplitude per control operator")

    @property
    def ops(self):
        return self.drift_ops + self.control_ops

    @property
    def dim(self):
        return self.ops[0].shape[0]

    def drift(self):
        out = self.drift_ops[0]
        for o in self.drift_ops[1:]:
            out = out + o
        return out.copy()

    def adjoint(self):
        adj = lambda o: o.conj().T.tocsr() if _is_sparse(o) else o.conj().T  # noqa: E731
        return Generator([adj(o) for o in self.drift_ops], [adj(o) for o in self.control_ops], self.amplitudes)

    def __repr__(self):
        return f"Generator(dim={self.dim}, drift terms={len(self.drift_ops)}, controls={len(self.amplitudes)})"


def hamiltonian(*terms):
    """``hamiltonian(H0, (H1, eps1), (H2, eps2))``: bare operators are drift terms, ``(op, control)``
    pairs are control terms.  Without any control term the plain (summed) matrix is returned, as in
    ``test/test_empty_optimization.jl:16-29``."""
    drift, cops, amps = [], [], []
    for term in terms:
        if isinstance(term, (tuple, list)) and len(term) == 2 and not np.isscalar(term[0]) and (
                callable(term[1]) or isinstance(term[1], PolynomialAmplitude) or np.ndim(term[1]) == 1) and (
                _is_sparse(term[0]) or np.ndim(term[0]) == 2):
            cops.append(term[0])
            amps.append(term[1])
        else:
            drift.append(term)
    if not cops:
        out = _as_matrix(drift[0])
        for o in drift[1:]:
            out = out + _as_matrix(o)
        return out
    if not drift:
        first = _as_matrix(cops[0])
        drift = [first * 0.0 if _is_sparse(first) else np.zeros(first.shape, np.complex128)]
    return Generator(drift, cops, amps)
