"""``KrotovResult``: same 16 fields, same order and meaning as ``src/result.jl:34-51``."""
from __future__ import annotations

import datetime as _dt
from dataclasses import dataclass, field
from typing import Any, List

import numpy as np

from .controls import discretize, get_controls

__all__ = ["KrotovResult"]


@dataclass
class KrotovResult:
    tlist: np.ndarray
    iter_start: int  # the starting iteration number
    iter_stop: int  # the maximum iteration number
    iter: int  # the current iteration number
    secs: float  # seconds that the last iteration took
    tau_vals: np.ndarray  # complex overlaps with the target states
    J_T: float  # current value of the final-time functional
    J_T_prev: float  # previous value
    guess_controls: List[np.ndarray]
    optimized_controls: List[np.ndarray]
    states: List[Any]  # forward-propagated states after each iteration
    start_local_time: _dt.datetime
    end_local_time: _dt.datetime
    records: List[tuple] = field(default_factory=list)
    converged: bool = False
    message: str = "in progress"

    @classmethod
    def from_problem(cls, problem):
        """``KrotovResult(problem)`` (``src/result.jl:53-89``)."""
        tlist = np.asarray(problem.tlist, np.float64)
        controls = get_controls(problem.trajectories)
        iter_start = int(problem.kwargs.get("iter_start", 0))
        iter_stop = int(problem.kwargs.get("iter_stop", 5000))
        guess = [discretize(c, tlist) for c in controls]
        now = _dt.datetime.now()
        return cls(
            tlist=tlist, iter_start=iter_start, iter_stop=iter_stop, iter=iter_start, secs=0.0,
            tau_vals=np.zeros(len(problem.trajectories), np.complex128), J_T=0.0, J_T_prev=0.0,
            guess_controls=guess, optimized_controls=[g.copy() for g in guess],
            states=[np.empty_like(t.initial_state) for t in problem.trajectories],
            start_local_time=now, end_local_time=now)

    def __repr__(self):
        return f"KrotovResult<{self.message}>"

    def __str__(self):
        n_it = max(self.iter - self.iter_start, 0)
        return ("Krotov Optimization Result\n"
                "--------------------------\n"
                f"- Started at {self.start_local_time.isoformat(timespec='milliseconds')}\n"
                f"- Number of trajectories: {len(self.states)}\n"
                f"- Number of iterations: {n_it}\n"
                f"- Value of functional: {self.J_T:.5e}\n"
                f"- Reason for termination: {self.message}\n"
                f"- Ended at {self.end_local_time.isoformat(timespec='milliseconds')} "
                f"({self.end_local_time - self.start_local_time})\n")


def convert_result(result):
    """``convert(KrotovResult, result)`` for results of other optimisers (``src/workspace.jl:110-113``):
    any object with the common fields is accepted."""
    if isinstance(result, KrotovResult):
        return result
    need = ["tlist", "iter_start", "iter_stop", "iter", "J_T", "J_T_prev", "guess_controls", "optimized_controls",
            "states", "records"]
    missing = [n for n in need if not hasattr(result, n)]
    if missing:
        raise TypeError(f"cannot convert {type(result).__name__} to KrotovResult: missing {missing}")
    now = _dt.datetime.now()
    return KrotovResult(
        tlist=np.asarray(result.tlist, np.float64), iter_start=int(result.iter_start), iter_stop=int(result.iter_stop),
        iter=int(result.iter), secs=float(getattr(result, "secs", 0.0)),
        tau_vals=np.asarray(getattr(result, "tau_vals", np.zeros(len(result.states))), np.complex128),
        J_T=float(result.J_T), J_T_prev=float(result.J_T_prev),
        guess_controls=[np.array(c, np.float64) for c in result.guess_controls],
        optimized_controls=[np.array(c, np.float64) for c in result.optimized_controls],
        states=list(result.states), start_local_time=getattr(result, "start_local_time", now),
        end_local_time=getattr(result, "end_local_time", now), records=list(result.records),
        converged=bool(getattr(result, "converged", False)), message=str(getattr(result, "message", "in progress")))
