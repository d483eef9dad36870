"""Developer probe: the ordered code-book search (mkNNIndxCB) at a given size, for an ncu launch list."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eng = Engine(0)
rng = np.random.default_rng(11)
c = rng.random((n, 2)) * np.sqrt(n / 5000.0)
c = c[np.argsort(c[:, 0], kind="stable")]
for r in range(2):
    t = time.perf_counter(); eng.mk_nn_indx(c, m, "cb"); print(f"n={n} m={m} cb: {time.perf_counter()-t:.3f} s", flush=True)
