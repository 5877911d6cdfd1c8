"""Generate the golden vectors in this directory with the NumPy oracle (oracle/krotov_oracle.py).

    python tests/golden/make_golden.py

The reference itself (Julia) cannot run in this image and its tests hold no numeric vector for this
path, so these are outputs of the *restatement* (PARITY UNPINNED); they pin the oracle
against regressions and give the GPU tests a fixed target that does not need the slow oracle at
BASELINE sizes (C3: d=25, N_T=2000)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import workloads as W  # noqa: E402
from oracle import krotov_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "c1_tls_cheby": (lambda: W.c1_tls(), 5, "cheby"),
    "c1_tls_expm": (lambda: W.c1_tls(), 5, "expm"),
    "c2_transmon_x": (lambda: W.c2_transmon_x(), 4, "cheby"),
    "dummy_d10": (lambda: W.dummy_dense(d=10, n_traj=2, n_controls=2), 3, "cheby"),
    "c3_two_transmon": (lambda: W.c3_two_transmon(), 2, "cheby"),
    "c4_8samples_g201": (lambda: W.c4_ensemble(n_samples=8, n_grid=201), 2, "cheby"),
}


def main(only=None):
    for name, (make, iters, method) in CASES.items():
        if only and name not in only:
            continue
        w = make()
        h = O.optimize_krotov(W.to_oracle(w), iters, method)
        out = {
            "workload": w.name, "iters": iters, "prop_method": method,
            "J_T": [float(x) for x in h["J_T"]],
            "g_a_int": [[float(v) for v in g] for g in h["g_a_int"]],
            "pulses": [[float(v) for v in row] for row in h["pulses"]],
            "tau_re": [float(v) for v in h["tau"][-1].real], "tau_im": [float(v) for v in h["tau"][-1].imag],
            "m_fw": h["m_fw"][-1][0] if h["m_fw"] else None,
        }
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(out, fh)
        print(name, out["J_T"])


if __name__ == "__main__":
    main(sys.argv[1:])
