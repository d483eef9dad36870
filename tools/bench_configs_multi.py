"""BASELINE configs[1] and [3] on N GPUs of one box (torchrun, one rank per GPU):

  configs[1]  spConjNNGP k-fold on BCEF (real data when data/bcef.npz exists): the 5 x 5 theta.alpha grid SHARDED over
              the ranks (zero communication until one sum at the end; dist.conj_kfold_sharded)
  configs[3]  predict.NNGP, 10^6 sites x 1,000 samples: the SITES sharded over the ranks (no collective; every rank
              holds coords / X / y, its slice of the sites, of z and of the result)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs_multi.py

Wall clock between barriers, max over ranks; one JSON line on rank 0 (also appended to gpurun_out/r02_configs_multi.jsonl).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spnngp_b200 import Engine  # noqa: E402
from spnngp_b200.dist import conj_kfold_sharded, make_allreduce, split_even  # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
which = set(sys.argv[1:]) or {"kfold", "predict"}
out = {"n_gpus": world}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def by_x(c):
    return c[np.argsort(c[:, 0], kind="stable")]


if "kfold" in which:
    rng = np.random.default_rng(11)
    m, k_fold = 15, 5
    bcef = os.path.join(ROOT, "data", "bcef.npz")
    real = os.path.exists(bcef)
    if real:
        d = np.load(bcef)
        c = np.column_stack([d["x"], d["y"]]); o = np.argsort(c[:, 0], kind="stable"); c = c[o]
        y = d["FCH"][o]; X = np.column_stack([np.ones(len(y)), d["PTC"][o]])
        n = len(y)
    else:
        n = 188717
        c = by_x(rng.random((n, 2)) * np.array([21.5, 17.0]))
        X = np.column_stack([np.ones(n), rng.standard_normal(n)])
        y = X @ np.array([1.0, 5.0]) + 2.0 * np.sin(c[:, 0]) * np.cos(c[:, 1]) + rng.standard_normal(n)
    ta = np.array([[ph, al] for ph in (0.3, 0.6, 1.5, 3, 6) for al in (0.05, 0.1, 0.25, 0.5, 1)]).T
    ho = np.stack([np.sort(rng.choice(n, n // k_fold, replace=False)) for _ in range(k_fold)]).astype(np.int32)
    ig = [2.0, float(np.var(y))]
    ar = make_allreduce(device=dev)
    warm = np.sort(rng.choice(20000, 4000, replace=False))[None, :].astype(np.int32)
    conj_kfold_sharded(eng, y[:20000], X[:20000], c[:20000], m, 0, ta, ig, warm, "cb", rank=rank, world=world, allreduce=ar)
    barrier()
    t0 = time.perf_counter()
    scores = conj_kfold_sharded(eng, y, X, c, m, 0, ta, ig, ho, "cb", rank=rank, world=world, allreduce=ar)
    t = max_over_ranks(time.perf_counter() - t0)
    best = int(np.argmin(scores[:, 1]))
    out["conj_kfold"] = dict(n=n, m=m, g=25, k_fold=k_fold, kfold_s=t, fits_per_s=25 * k_fold / t, best_set=best,
                             crps_best=float(scores[best, 1]),
                             sets_per_rank=[split_even(25, r, world)[1] - split_even(25, r, world)[0] for r in range(world)],
                             data="BCEF" if real else "synthetic BCEF-shaped")

if "predict" in which:
    n, m, q, S, distinct = 1000000, 15, 1000000, 1000, 400
    rng = np.random.default_rng(12)  # same stream on every rank: identical fit data
    W = np.sqrt(n / 5000.0)
    c = by_x(rng.random((n, 2)) * W)
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    beta = np.array([1.0, 5.0])
    y = X @ beta + rng.standard_normal(n)
    eng.set_data(c, None, None, X, y, m)  # prediction-only use of the context: sites, X, y
    bs = beta[:, None] + 0.05 * rng.standard_normal((2, S))
    base = np.array([[5.0], [1.0], [6.0], [0.75]]) * (1 + 0.02 * rng.standard_normal((4, distinct)))
    th = base[:, np.sort(rng.integers(0, distinct, S))]
    lo, hi = split_even(q, rank, world)
    qs = hi - lo
    rs = np.random.default_rng(100 + rank)  # this rank's slice of the sites and of the normals
    c0 = rs.random((qs, 2)) * W
    X0 = np.column_stack([np.ones(qs), rs.standard_normal(qs)])
    res = {}
    for cov, name in ((0, "exponential"), (2, "matern")):
        z = rs.standard_normal(qs * S)
        barrier()
        t0 = time.perf_counter()
        nn0 = eng.nn_query(c, c0, m)
        t1 = time.perf_counter()
        y0 = eng.rnngp_predict(cov, X0, c0, nn0, bs, th[: (4 if cov == 2 else 3)], z)
        t2 = time.perf_counter()
        tq, tp = max_over_ranks(t1 - t0), max_over_ranks(t2 - t1)
        res[name] = dict(nn_query_s=tq, predict_s=tp, draws_per_s=q * S / tp, mean_of_draws=float(np.mean(y0)))
        del z, y0
    out["predict"] = dict(n=n, q=q, S=S, sites_per_rank=qs, distinct_theta=int(len(np.unique(th[2]))), **res)

if rank == 0:
    line = json.dumps(out)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_configs_multi.jsonl"), "a") as f:
        f.write(line + "\n")
eng.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
