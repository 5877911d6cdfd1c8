"""``Trajectory`` and ``ControlProblem`` with the field names of QuantumControl.jl (the objects
``KrotovWrk`` reads at ``src/workspace.jl:65-76``)."""
from __future__ import annotations

import numpy as np

from .generators import Generator

__all__ = ["Trajectory", "ControlProblem"]


class Trajectory:
    """One initial state evolving under one generator; optional ``target_state`` and ``weight``.
    Extra keyword arguments (e.g. ``prop_method``) become attributes, like in the reference."""

    def __init__(self, initial_state, generator, *, target_state=None, weight=1.0, **kwargs):
        self.initial_state = np.asarray(initial_state, dtype=np.complex128)
        self.generator = generator
        self.target_state = None if target_state is None else np.asarray(target_state, dtype=np.complex128)
        self.weight = float(weight)
        self.kwargs = dict(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def adjoint(self):
        if isinstance(self.generator, Generator):
            g = self.generator.adjoint()
        elif hasattr(self.generator, "tocsr"):  # control-free scipy.sparse generator: stays sparse
            g = self.generator.conj().T.tocsr()
        else:
            g = np.asarray(self.generator).conj().T
        return Trajectory(self.initial_state, g, target_state=self.target_state, weight=self.weight, **self.kwargs)


class ControlProblem:
    """``ControlProblem(trajectories, tlist; kwargs...)``."""

    def __init__(self, trajectories, tlist, **kwargs):
        self.trajectories = list(trajectories)
        self.tlist = np.asarray(tlist, dtype=np.float64)
        self.kwargs = dict(kwargs)
