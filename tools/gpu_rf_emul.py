"""Ad-hoc: backward-sweep time of the replicated forward sweep with several ranks EMULATED on one GPU (in-kernel counters),
against the one-rank run of the same ensemble.   python tools/gpu_rf_emul.py [samples] [ranks]"""
import os, sys
os.environ["KROTOV_PROF"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = W.c4_ensemble(n_samples=ns)
ghz = 1.965


def run(label, **kw):
    out = {}

    def cb(wrk, it, a, b):
        if it >= 1:
            out["info"] = wrk.engine.info()
            engines = getattr(wrk.engine, "engines", [wrk.engine])
            out["prof"] = [[e.profile(c) for c in range(e.info()["grid_blocks"])] for e in engines]

    run_product(w, 3, callback=cb, **kw)
    f = lambda v: v / ghz / 1e3 / w.N_T
    i = out["info"]
    print(f"{label}: exchange={i['exchange']} grid={i['grid_blocks']}x{i['block_threads']} ms={i['ms_last']:.2f} "
          f"rank_wait_ms={i.get('ms_rank_wait', 0):.3f}")
    for r, prof in enumerate(out["prof"]):
        for key in ("backward", "forward"):
            v = np.array([f(a[key]) for a in prof])
            print(f"   rank {r} {key:9s} min={v.min():.3f} mean={v.mean():.3f} max={v.max():.3f} us/step")


run("one rank")
run(f"{ranks} emulated ranks, replicate", emulate_ranks=ranks, multi_gpu="replicate")
run(f"{ranks} emulated ranks, shard", emulate_ranks=ranks, multi_gpu="shard")
