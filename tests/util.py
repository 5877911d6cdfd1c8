"""Shared helpers for the tests: build the product-API problem and the oracle arrays from one Workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import krotov_jl_b200 as K  # noqa: E402
import workloads as W  # noqa: E402

_JT = {"sm": K.J_T_sm, "ss": K.J_T_ss, "re": K.J_T_re}


def to_problem(w, **kwargs):
    """Workload -> ControlProblem through the reference-style API (hamiltonian / Trajectory / ControlProblem).
    Ensemble members that share a generator share the Generator OBJECT and all trajectories share the
    control OBJECTS, as a user of the reference would write it."""
    gens = []
    amps = list(w.controls)
    for l, c in enumerate(w.controls):  # non-linear / shaped amplitudes: ONE amplitude object per control
        poly = None if w.amp_poly is None else w.amp_poly[l]
        shape = None if w.amp_shape is None else w.amp_shape[l]
        if poly is not None:
            amps[l] = K.PolynomialAmplitude(c, poly, shape)
        elif shape is not None:
            amps[l] = K.ShapedAmplitude(c, shape)
    for g in range(len(w.H0)):
        terms = [w.H0[g]]
        for l, c in enumerate(w.controls):
            if w.Hc[g][l] is not None:
                terms.append((w.Hc[g][l], amps[l]))
        gens.append(K.hamiltonian(*terms))
    trajs = [K.Trajectory(w.psi0[k], gens[int(w.gen_of_traj[k])], target_state=w.target[k]) for k in range(w.N)]
    kw = dict(prop_method=K.Cheby, J_T=_JT[w.functional], lambda_a=w.lambda_a, update_shape=w.update_shape,
              print_iters=False)
    if w.specrange is not None:
        kw.update(prop_E_min=w.specrange[0], prop_E_max=w.specrange[1])
    kw.update(kwargs)
    return K.ControlProblem(trajs, w.tlist, **kw)


def run_product(w, iters, **kwargs):
    """Optimise through the public API, recording per-iteration J_T, g_a_int, tau."""
    hist = dict(J_T=[], g_a_int=[], tau=[])
    extra_cb = kwargs.pop("callback", None)

    def record(wrk, it, eps_new, eps_old):
        if extra_cb is not None:
            extra_cb(wrk, it, eps_new, eps_old)
        hist["J_T"].append(wrk.result.J_T)
        hist["tau"].append(np.array(wrk.result.tau_vals))
        if it > 0:
            hist["g_a_int"].append(np.array(wrk.g_a_int))
        hist["pulses"] = np.array([np.array(e) for e in eps_new])
        hist["info"] = wrk.engine.info()
        hist["m_fw"] = [len(c[0]) for c in wrk.fw_settings.coeffs]

    problem = to_problem(w, iter_stop=iters, callback=record, **kwargs)
    res = K.optimize(problem, method=K.Krotov)
    hist["result"] = res
    return hist


def rel_abs(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300), np.abs(a - b)
