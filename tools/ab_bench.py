"""Kernel A/B: time the fused B/F launch (n = 10^6, m = 15) with alternative builds of the library
(tools/build_variant.sh -> spnngp_b200/variants/libnngp_<name>.so), one subprocess per build.
    python tools/ab_bench.py base b6 rep ...          # parent
Writes gpurun_out/ab_bench.json."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n, m):
    from spnngp_b200 import Engine
    rng = np.random.default_rng(1)
    c = rng.random((n, 2)) * np.sqrt(n / 5000.0)
    c = c[np.argsort(c[:, 0], kind="stable")]
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    beta = np.array([1.0, 5.0])
    y = X @ beta + rng.standard_normal(n)
    eng = Engine(0)
    idx, _, lu = eng.mk_nn_indx(c, m, "cb")
    eng.set_data(c, idx, lu, X, y, m)
    res = {}
    for name, cov, theta, short in (("matern", 2, [5, 1, 6, 0.75], 1), ("matern_full", 2, [5, 1, 6, 0.75], 0),
                                    ("exponential", 0, [5, 1, 6], 1), ("gaussian", 3, [5, 1, 6], 1)):
        eng.set_option("matern_short", short)
        out = eng.bf_loglik(cov, [theta], [beta], slots=[0])
        for nb in (1, 2):
            eng.time_bf_loglik(cov, [theta] * nb, [beta] * nb, iters=3)
            ms = min(eng.time_bf_loglik(cov, [theta] * nb, [beta] * nb, iters=20) for _ in range(2))
            res[f"{name}_b{nb}_ms"] = ms
        res[f"{name}_logdet"] = float(out[0, 0]); res[f"{name}_q"] = float(out[0, 1])
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    allres = {}
    for name in sys.argv[1:]:
        lib = os.path.join(ROOT, "spnngp_b200", "variants", f"libnngp_{name}.so") if name != "default" else ""
        env = dict(os.environ)
        if lib:
            env["NNGP_B200_LIB"] = lib
        p = subprocess.run([sys.executable, __file__, "--child", "1000000", "15"], env=env, capture_output=True, text=True,
                           timeout=300)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        allres[name] = json.loads(line[0][7:]) if line else {"error": (p.stderr or p.stdout)[-400:]}
        print(name, allres[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(allres, open(os.path.join(ROOT, "gpurun_out", "ab_bench.json"), "w"), indent=1)
