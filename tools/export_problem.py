"""Write a workload as plain files that ANY implementation can read -- the oracle, the CUDA path, and the reference
Krotov.jl itself on a machine that has Julia (`julia/reference_vectors.jl` reads exactly these files and writes the
J_T history and the optimised pulses of the UNMODIFIED reference; committed under tests/golden/julia/ they turn
"parity unpinned" into a pinned comparison).

    python tools/export_problem.py c1_tls [c2_transmon_x c3_two_transmon c4_8 ...]      # -> tests/golden/export/<name>/

Files: `problem.json` (sizes, functional, lambda_a, iterations, layout notes) and little-endian binaries
  tlist.f64        [N_T+1]
  gen_of_traj.i32  [N]                 0-based
  H0.c128          [n_gen][d][d]       row-major (C order); complex = (re, im) pairs of Float64
  Hc.c128          [n_gen][L][d][d]    row-major; a missing control term is all zeros and listed in problem.json["missing"]
  psi0.c128        [N][d]
  target.c128      [N][d]
  pulses.f64       [L][N_T]            guess pulses ON THE MIDPOINTS (what `discretize_on_midpoints` gives the reference)
  shape.f64        [L][N_T]            update shapes on the midpoints
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import workloads as W  # noqa: E402

CASES = {
    "c1_tls": (lambda: W.c1_tls(), 5),
    "c2_transmon_x": (lambda: W.c2_transmon_x(), 4),
    "c3_two_transmon": (lambda: W.c3_two_transmon(), 2),
    "c4_8": (lambda: W.c4_ensemble(n_samples=8, n_grid=201), 2),
}


def export(name, out_root):
    make, iters = CASES[name]
    w = make()
    p = W.to_oracle(w)
    out = os.path.join(out_root, name)
    os.makedirs(out, exist_ok=True)
    n_gen, d, L, N, N_T = len(p.H0), p.d, p.L, p.N, p.N_T
    Hc = np.zeros((n_gen, L, d, d), np.complex128)
    missing = []
    for g in range(n_gen):
        for l in range(L):
            if p.Hc[g][l] is None:
                missing.append([g, l])
            else:
                Hc[g, l] = p.Hc[g][l]
    np.asarray(p.tlist, "<f8").tofile(os.path.join(out, "tlist.f64"))
    np.asarray(p.gen_of_traj, "<i4").tofile(os.path.join(out, "gen_of_traj.i32"))
    np.asarray(p.H0, "<c16").tofile(os.path.join(out, "H0.c128"))
    Hc.astype("<c16").tofile(os.path.join(out, "Hc.c128"))
    np.asarray(p.psi0, "<c16").tofile(os.path.join(out, "psi0.c128"))
    np.asarray(p.target, "<c16").tofile(os.path.join(out, "target.c128"))
    np.asarray(p.pulses, "<f8").tofile(os.path.join(out, "pulses.f64"))
    np.asarray(p.S, "<f8").tofile(os.path.join(out, "shape.f64"))
    meta = dict(name=name, workload=w.name, d=d, N=N, L=L, N_T=N_T, n_gen=n_gen, functional=p.functional,
                lambda_a=[float(x) for x in p.lam], iters=iters, missing=missing,
                prop_method="Cheby", cheby_coeffs_limit=p.cheby_limit, specrange_buffer=p.specrange_buffer,
                specrange=None if p.specrange is None else list(p.specrange),
                layout="row-major (C order) arrays of little-endian Float64 / (re, im) Float64 pairs; see tools/export_problem.py")
    with open(os.path.join(out, "problem.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    return out


def load(folder):
    """The exported files back as an oracle ProblemArrays (tests: the round trip must be exact)."""
    from oracle.krotov_oracle import ProblemArrays

    with open(os.path.join(folder, "problem.json")) as fh:
        m = json.load(fh)
    d, N, L, N_T, n_gen = m["d"], m["N"], m["L"], m["N_T"], m["n_gen"]
    rd = lambda f, t, shape: np.fromfile(os.path.join(folder, f), t).reshape(shape)  # noqa: E731
    Hc = rd("Hc.c128", "<c16", (n_gen, L, d, d))
    missing = {tuple(x) for x in m["missing"]}
    return m, ProblemArrays(
        tlist=rd("tlist.f64", "<f8", (N_T + 1,)), H0=list(rd("H0.c128", "<c16", (n_gen, d, d))),
        Hc=[[None if (g, l) in missing else Hc[g, l] for l in range(L)] for g in range(n_gen)],
        gen_of_traj=rd("gen_of_traj.i32", "<i4", (N,)).astype(int), psi0=rd("psi0.c128", "<c16", (N, d)),
        target=rd("target.c128", "<c16", (N, d)), pulses=rd("pulses.f64", "<f8", (L, N_T)), S=rd("shape.f64", "<f8", (L, N_T)),
        lam=np.array(m["lambda_a"]), functional=m["functional"], cheby_limit=m["cheby_coeffs_limit"],
        specrange_buffer=m["specrange_buffer"], specrange=None if m["specrange"] is None else tuple(m["specrange"]))


if __name__ == "__main__":
    names = sys.argv[1:] or ["c1_tls", "c2_transmon_x"]
    for n in names:
        print(export(n, os.path.join(ROOT, "tests", "golden", "export")))
