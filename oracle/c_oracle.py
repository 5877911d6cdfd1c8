"""ctypes front end of ``libkrotov_oracle.so`` (the C restatement in ``krotov_oracle.c``).

TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU-baseline legs import this."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libkrotov_oracle.so")
    src = os.path.join(_HERE, "krotov_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libkrotov_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_krotov_optimize.restype = ctypes.c_int
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _csr_terms(p, tol=0.0):
    """Per (generator, term) CSR blocks: rowptr [nterm][d+1], term_off, col, val."""
    d, L = p.d, p.L
    rowptr, term_off, cols, vals = [], [0], [], []
    for g in range(len(p.H0)):
        for t in range(1 + L):
            M = p.H0[g] if t == 0 else p.Hc[g][t - 1]
            rp = np.zeros(d + 1, np.int32)
            if M is not None:
                M = np.asarray(M)
                for i in range(d):
                    nz = np.nonzero(np.abs(M[i]) > tol)[0]
                    rp[i + 1] = rp[i] + len(nz)
                    cols.append(nz.astype(np.int32))
                    vals.append(M[i, nz].astype(complex))
            rowptr.append(rp)
            term_off.append(term_off[-1] + int(rp[d]))
    col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    val = np.concatenate(vals) if vals else np.zeros(0, complex)
    return (np.ascontiguousarray(np.concatenate(rowptr), np.int32), np.array(term_off, np.int32),
            np.ascontiguousarray(col, np.int32), np.ascontiguousarray(val, complex))


_KIND = {"sm": 0, "ss": 1, "re": 2}


def optimize_krotov_c(p, iters, n_threads=0):
    """Run the C oracle on a ``ProblemArrays``; returns a dict like the NumPy oracle's."""
    L_ = lib()
    d, N, L, N_T = p.d, p.N, p.L, p.N_T
    rowptr, term_off, col, val = _csr_terms(p)
    tlist = np.ascontiguousarray(p.tlist, float)
    gen = np.ascontiguousarray(p.gen_of_traj, np.int32)
    psi0 = np.ascontiguousarray(p.psi0, complex)
    tgt = np.ascontiguousarray(p.target, complex)
    w = np.ascontiguousarray(p.weights(), float)
    pulses = np.ascontiguousarray(p.pulses, float).copy()
    S = np.ascontiguousarray(p.S, float)
    lam = np.ascontiguousarray(p.lam, float)
    JT = np.zeros(iters + 1)
    ga = np.zeros((max(iters, 1), L))
    tau = np.zeros(N, complex)
    states = np.zeros((N, d), complex)
    m = np.zeros(2, np.int32)
    secs = ctypes.c_double(0.0)
    has_range = p.specrange is not None
    Emin, Emax = p.specrange if has_range else (0.0, 0.0)

    def P(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    keep = []
    if getattr(p, "nonlinear", False):
        poly = None
        deg = 1
        if p.amp_poly is not None:
            deg = max(1, max(len(c) - 1 for c in p.amp_poly if c is not None)) if any(c is not None for c in p.amp_poly) else 1
            poly = np.zeros((L, deg + 1))
            for l in range(L):
                if p.amp_poly[l] is None:
                    poly[l, 1] = 1.0
                else:
                    poly[l, : len(p.amp_poly[l])] = p.amp_poly[l]
        shape = None if p.amp_shape is None else np.ascontiguousarray(p.amp_shape, float)
        keep = [poly, shape]
        L_.oracle_set_amplitudes(ctypes.c_int(deg), None if poly is None else P(poly), None if shape is None else P(shape))

    rc = L_.oracle_krotov_optimize(
        ctypes.c_int(d), ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(N_T), ctypes.c_int(len(p.H0)),
        P(tlist), P(gen), P(rowptr), P(term_off), P(col), P(val), P(psi0), P(tgt), P(w), P(pulses), P(S), P(lam),
        ctypes.c_int(_KIND[p.functional]), ctypes.c_double(p.cheby_limit), ctypes.c_double(p.specrange_buffer),
        ctypes.c_int(int(has_range)), ctypes.c_double(Emin), ctypes.c_double(Emax), ctypes.c_int(iters),
        ctypes.c_int(n_threads), P(JT), P(ga), P(tau), P(states), P(m), ctypes.byref(secs))
    if rc != 0:
        raise RuntimeError(f"oracle_krotov_optimize failed: {rc}")
    return dict(J_T=list(JT), g_a_int=[ga[i].copy() for i in range(iters)], tau=tau, pulses=pulses,
                states=states, m=(int(m[0]), int(m[1])), secs=secs.value,
                threads=n_threads if n_threads > 0 else L_.oracle_num_threads())
