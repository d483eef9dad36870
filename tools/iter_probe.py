"""Developer probe: wall time per sampler iteration on ONE GPU for the whole data set and for a 125k-location shard (what one
rank of an 8-GPU fit runs; its sums are partial, only the timing is meaningful).  NNGP_PDL=1 switches the programmatic
dependent launch on.  python tools/iter_probe.py [iterations]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine, _lib  # noqa: E402
from bench import make_workload, sampler_settings  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n, m = 1000000, 15
wl = make_workload(n, m)
R = _lib.load_rshim()
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
for cov, name in ((2, "matern"), (0, "exponential")):
    starting, tuning, pri = sampler_settings(cov)
    for lo, hi in ((0, n), (500000, 625000), (0, 125000)):
        eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m, loc_range=(lo, hi))
        def chain(k):
            R.rshim_set_seed(1)
            t = time.perf_counter()
            eng.rnngp(cov, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"], pri["nuUnif"], starting, tuning, k, R.rnorm, R.runif, mode=1)
            return time.perf_counter() - t
        chain(20)
        t1 = min(chain(K) for _ in range(2)); t0 = min(chain(20) for _ in range(2))
        print(f"{name:12s} shard [{lo:7d},{hi:7d}): {1e6 * (t1 - t0) / (K - 20):7.1f} us per iteration (marginal), "
              f"{1e6 * t0 / 20:7.1f} us per iteration in a 20-iteration call   pdl={'on' if os.environ.get('NNGP_PDL') else 'off'}", flush=True)
