"""Developer probe: the Matern B/F launch at n = 10^6, m = 15 with the three series lengths (matern_short = 2 / 1 / 0:
5, 9, 14 terms) plus the exponential launch, CUDA-event timed; writes gpurun_out/series_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402
from bench import make_workload  # noqa: E402

n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, int(sys.argv[2]) if len(sys.argv) > 2 else 15
wl = make_workload(n, m)
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m)
res = {"n": n, "m": m}
th = [5.0, 1.0, 6.0, 0.75]
for level in (2, 1, 0):
    eng.set_option("matern_short", level)
    out = eng.bf_loglik(2, [th], [wl["beta"]], slots=[0])
    eng.time_bf_loglik(2, [th], [wl["beta"]], iters=3)
    ms = min(eng.time_bf_loglik(2, [th], [wl["beta"]], iters=20) for _ in range(3))
    res[f"matern_level{level}"] = dict(ms=ms, tiny_used=eng.get_option("matern_tiny_used"), short_used=eng.get_option("matern_short_used"),
                                       logdet=float(out[0, 0]), q=float(out[0, 1]))
eng.set_option("matern_short", 2)
eng.time_bf_loglik(0, [th[:3]], [wl["beta"]], iters=3)
res["exponential_ms"] = min(eng.time_bf_loglik(0, [th[:3]], [wl["beta"]], iters=20) for _ in range(3))
print(json.dumps(res), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"series_probe_n{n}_m{m}.json"), "w") as f:
    json.dump(res, f)
