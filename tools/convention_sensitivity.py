"""How much each RECALLED upstream convention (tags [R] / [?]) could move the numbers.

The reference's arithmetic lives in un-vendored QuantumPropagators.jl; the oracle restates its conventions from memory.
For every such convention this script re-runs the NumPy oracle with the plausible alternative and reports the largest
deviation of J_T (absolute and relative) and of the optimised pulses from the shipped convention, on C1 (two-level
system, 5 iterations), C2 (single transmon, 4 iterations) and a short C3 (two transmons, 200 steps, 2 iterations);
for C1 also against the 50-digit exact-propagator optimisation of tests/mp_reference.py, which depends on NONE of the
Chebyshev conventions.  A convention whose alternative stays below 1e-10 relative in J_T and 1e-9 in the pulses cannot
break parity with Krotov.jl whichever way upstream does it; one that exceeds it is a real parity risk and is listed as
such in DESIGN.md.

    python tools/convention_sensitivity.py > profiles/r2_convention_sensitivity.txt      (CPU only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workloads as W  # noqa: E402
from oracle import krotov_oracle as O  # noqa: E402
from scipy.special import jv  # noqa: E402


def coeffs_variant(guard):
    def f(Delta, dt, limit=1e-12):
        alpha = abs(0.5 * Delta * dt)
        out = [float(jv(0, alpha))]
        i = 1
        while abs(out[-1]) > limit or (guard == "le" and i <= alpha) or (guard == "lt" and i < alpha):
            out.append(2.0 * float(jv(i, alpha)))
            i += 1
        return np.array(out)

    return f


def envelope_variant(kind):
    def f(self):
        if self.explicit_specrange is not None:
            E_min, E_max = self.explicit_specrange
        else:
            lo = [r[0] for r in self.control_ranges]
            hi = [r[1] for r in self.control_ranges]
            E_min, E_max = O.specrange_diag(self._evaluate_at(hi))
            a, b = O.specrange_diag(self._evaluate_at(lo))
            E_min, E_max = min(E_min, a), max(E_max, b)
        Delta = E_max - E_min
        delta = self.specrange_buffer * Delta
        if kind == "half":  # shipped: buffer split over both ends
            self.E_min, self.Delta = E_min - delta / 2, Delta + delta
        elif kind == "full":  # buffer added at each end
            self.E_min, self.Delta = E_min - delta, Delta + 2 * delta
        elif kind == "none":
            self.E_min, self.Delta = E_min, Delta
        self.dt = self.tlist[1] - self.tlist[0]
        if self.backward:
            self.dt = -self.dt
        self.coeffs = O.cheby_coeffs(self.Delta, self.dt, self.limit)

    return f


def run(workload, iters, patch=None):
    saved = {}
    if patch:
        for (obj, name), val in patch.items():
            saved[(obj, name)] = getattr(obj, name)
            setattr(obj, name, val)
    try:
        p = W.to_oracle(workload())
        if patch and "pulses" in patch.get("_post", {}):
            pass
        return O.optimize_krotov(p, iters)
    finally:
        for (obj, name), val in saved.items():
            setattr(obj, name, val)


def all_midpoints(f, tlist):
    t = np.asarray(tlist, float)
    return np.array([f(0.5 * (t[i] + t[i + 1])) for i in range(len(t) - 1)])


def main():
    import mp_reference as M

    cases = [("C1 TLS (5 it.)", W.c1_tls, 5), ("C2 transmon (4 it.)", W.c2_transmon_x, 4),
             ("C3 short (200 steps, 2 it.)", lambda: W.c3_two_transmon(n_grid=201, T=40.0), 2)]
    variants = [
        ("A.1 truncation guard: stop at first |a_n| <= limit with n <= alpha [shipped] -> n < alpha",
         {(O, "cheby_coeffs"): coeffs_variant("lt")}),
        ("A.1 truncation guard -> none (first |a_n| <= limit)", {(O, "cheby_coeffs"): coeffs_variant("none")}),
        ("A.1 specrange_buffer: (E_min - d/2, Delta + d) [shipped] -> (E_min - d, Delta + 2d)",
         {(O.ChebyPropagator, "_set_spectral_envelope"): envelope_variant("full")}),
        ("A.1 specrange_buffer -> no buffer", {(O.ChebyPropagator, "_set_spectral_envelope"): envelope_variant("none")}),
        ("A.1 first reinit_prop! always re-derives the envelope [shipped] -> never widens (factors 1/1)",
         {(O, "transform_control_ranges"): lambda c, a, b, check: (a, b)}),
        ("optimize.jl:238-244 factors 2/5 [reference, not recalled] -> 2/10 (sanity row)",
         {(O, "transform_control_ranges"): lambda c, a, b, check: ((min(a, 2 * a), max(b, 2 * b)) if check else
                                                                    (min(a, 10 * a), max(b, 10 * b)))}),
    ]
    print(__doc__.split("\n\n")[0])
    print()
    base = {name: run(make, it) for name, make, it in cases}
    exact = M.tls_krotov_exact(5)
    b = base[cases[0][0]]
    dj = np.abs(np.array(b["J_T"]) - np.array(exact["J_T"]))
    print("shipped conventions vs 50-digit exact-propagator optimisation (C1): max |dJ_T| = %.2e (rel %.2e), pulses %.2e"
          % (dj.max(), (dj / np.array(exact["J_T"])).max(), np.abs(b["pulses"][0] - np.array(exact["pulses"])).max()))
    print()
    print("%-100s %-28s %10s %10s %10s %s" % ("convention -> alternative", "case", "|dJ_T|", "rel dJ_T", "|d eps|", "m"))
    for label, patch in variants:
        for name, make, it in cases:
            h = run(make, it, patch)
            ref = base[name]
            dj = np.abs(np.array(h["J_T"]) - np.array(ref["J_T"]))
            rel = (dj / np.abs(np.array(ref["J_T"]))).max()
            dp = np.abs(h["pulses"] - ref["pulses"]).max()
            m = "%d -> %d" % (ref["m_fw"][-1][0], h["m_fw"][-1][0])
            flag = "  PARITY RISK" if (rel > 1e-10 and dj.max() > 2e-15) or dp > 1e-9 else ""
            print("%-100s %-28s %10.2e %10.2e %10.2e %s%s" % (label[:100], name, dj.max(), rel, dp, m, flag))
    # midpoint rule (A.4): first / last sample ON the grid ends [shipped, V] vs every sample on its midpoint
    for name, make, it in cases:
        w = make()
        p = W.to_oracle(w)
        p.pulses = np.array([all_midpoints(c, w.tlist) for c in w.controls])
        p.S = np.array([all_midpoints(w.update_shape, w.tlist) for _ in w.controls])
        h = O.optimize_krotov(p, it)
        ref = base[name]
        dj = np.abs(np.array(h["J_T"]) - np.array(ref["J_T"]))
        print("%-100s %-28s %10.2e %10.2e %10.2e %s" % (
            "A.4 discretize_on_midpoints: ends ON the grid [shipped, verified] -> all samples on midpoints", name,
            dj.max(), (dj / np.abs(np.array(ref["J_T"]))).max(), np.abs(h["pulses"] - ref["pulses"]).max(), ""))


if __name__ == "__main__":
    main()
