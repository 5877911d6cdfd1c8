"""Golden vectors that do NOT come from the oracle: the two-level system of test/test_tls_optimization.jl:12-63 optimised
in 50-digit arithmetic with the closed-form propagator of every interval (tests/mp_reference.py), first order and
second order (sigma = -2), 30 significant digits stored next to the Float64 values; and the general exact-propagator
loop (mpmath.expm, 40 digits) on the small problems of mp_reference.exact_cases(): several trajectories / generators /
controls, a missing control term, complex operators, the three functionals, a non-Hermitian generator.

    python tests/golden/make_golden_exact.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import mpmath as mp  # noqa: E402

import mp_reference as M  # noqa: E402

CASES = {"c1_tls_exact50": dict(iters=5), "c1_tls_sigma_exact50": dict(iters=3, sigma=-2.0)}

for name, kw in CASES.items():
    h = M.tls_krotov_exact(**kw)
    out = {"source": "tests/mp_reference.py tls_krotov_exact(%s), mpmath dps=50" % ", ".join(f"{k}={v}" for k, v in kw.items()),
           "J_T": [float(x) for x in h["J_T"]], "J_T_30_digits": [mp.nstr(x, 30) for x in h["J_T_mp"]],
           "g_a_int": h["g_a_int"], "pulses": h["pulses"], **({"sigma": kw["sigma"]} if "sigma" in kw else {})}
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(out, fh)
    print(name, out["J_T_30_digits"])

import workloads as W  # noqa: E402

for name, (make, iters) in M.exact_cases().items():
    if "--only-new" in sys.argv and os.path.exists(os.path.join(HERE, name + "_exact40.json")):
        continue
    h = M.krotov_exact_general(W.to_oracle(make()), iters)
    out = {"source": f"tests/mp_reference.py krotov_exact_general(exact_cases()[{name!r}]), mpmath dps=40, mpmath.expm per interval",
           "iters": iters, "J_T": h["J_T"], "g_a_int": h["g_a_int"], "pulses": h["pulses"],
           "tau_re": [t.real for t in h["tau"]], "tau_im": [t.imag for t in h["tau"]]}
    with open(os.path.join(HERE, name + "_exact40.json"), "w") as fh:
        json.dump(out, fh)
    print(name, out["J_T"])

# second order with a sigma that varies over the grid, on a non-Hermitian generator and on two generators: the GENERAL
# update (previous trajectory stored), which the oracle implements and the device path refuses
for name in ("non_hermitian_d4", "two_generators_d5"):
    make, iters = M.exact_cases()[name]
    p = W.to_oracle(make())
    mids = [float(p.tlist[0])] + [float(p.tlist[i] + 0.5 * (p.tlist[i + 1] - p.tlist[i])) for i in range(1, len(p.tlist) - 2)] + [float(p.tlist[-1])]
    sigma = [-0.6 - 0.3 * t for t in mids]
    h = M.krotov_exact_general(p, iters, sigma=sigma)
    out = {"source": f"krotov_exact_general(exact_cases()[{name!r}], sigma(t) = -0.6 - 0.3 t on the midpoints), mpmath dps=40",
           "iters": iters, "sigma": sigma, "J_T": h["J_T"], "g_a_int": h["g_a_int"], "pulses": h["pulses"]}
    with open(os.path.join(HERE, name + "_sigma_exact40.json"), "w") as fh:
        json.dump(out, fh)
    print(name, "sigma(t)", out["J_T"])

if "--c3" in sys.argv:  # 4.5 minutes: mpmath.expm on 25 x 25 matrices, 30 digits
    w = W.c3_two_transmon(n_grid=41, T=8.0)
    h = M.krotov_exact_general(W.to_oracle(w), 2, dps=30)
    out = {"source": "tests/mp_reference.py krotov_exact_general(W.c3_two_transmon(n_grid=41, T=8.0)), mpmath dps=30, mpmath.expm per interval",
           "iters": 2, "J_T": h["J_T"], "g_a_int": h["g_a_int"], "pulses": h["pulses"],
           "tau_re": [t.real for t in h["tau"]], "tau_im": [t.imag for t in h["tau"]]}
    with open(os.path.join(HERE, "c3_two_transmon_g41_exact30.json"), "w") as fh:
        json.dump(out, fh)
    print("c3_two_transmon_g41", out["J_T"])
