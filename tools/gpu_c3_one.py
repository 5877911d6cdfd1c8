"""Ad-hoc: two iterations of C3 (single CTA of the warp kernel), for ncu source-level captures."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
K.optimize(to_problem(W.c3_two_transmon(), iter_stop=2), method=K.Krotov)
