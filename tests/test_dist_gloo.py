"""CPU, world_size = 2 over gloo: the host-side logic of the multi-GPU path -- shard ranges partition the
locations, per-shard partial sums reduce to the single-process result through the allreduce hook."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    from spnngp_b200.dist import shard_range, split_even, nn_offset
    for n, m in ((1, 1), (5, 15), (100, 15), (5000, 15), (1000003, 32)):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, m, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            if n > 100 * world:
                w = [nn_offset(b, m) + b - nn_offset(a, m) - a for a, b in r]
                assert max(w) - min(w) <= 2 * (m + 1)
    assert [split_even(10, k, 3) for k in range(3)] == [(0, 3), (3, 6), (6, 10)]


def _worker(rank, world, port_file, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle as po
    from spnngp_b200.dist import make_allreduce, shard_range
    from tests.util import make_problem
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    port = po.Port(); port.set_threads(1)
    prob = make_problem(3000, 10, seed=5, port=port, domain=0.8)
    n, m = 3000, 10
    B, F, ld, _ = port.update_bf(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], n, m, [5.0, 1.0, 6.0], 0)
    r = prob["X"] @ prob["beta"] - prob["y"]
    lo, hi = shard_range(n, m, rank, world)
    # per-location terms of sum log F and Q (src/util.cpp:261-280), restricted to this rank's shard
    off, cnt = prob["nnIndxLU"][:n], prob["nnIndxLU"][n:]
    part = np.zeros(2)
    for i in range(lo, hi):
        a = B[off[i]:off[i] + cnt[i]] @ r[prob["nnIndx"][off[i]:off[i] + cnt[i]]]
        part[0] += np.log(F[i]); part[1] += (r[i] - a) ** 2 / F[i]
    ar = make_allreduce(device=None)
    assert ar is not None
    ar(part)
    full_q = port.q(B, F, r, r, n, prob["nnIndx"], prob["nnIndxLU"])
    q.put((rank, abs(part[0] - ld) / abs(ld), abs(part[1] - full_q) / abs(full_q)))
    dist.destroy_process_group()


def test_sharded_partials_allreduce_gloo(tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    f = str(tmp_path / "rdzv")
    procs = [ctx.Process(target=_worker, args=(k, 2, f, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, e1, e2 in res:
        assert e1 < 1e-13 and e2 < 1e-13


class _OracleKfoldEngine:
    """Stands in for spnngp_b200.Engine in the CPU test of the sharding logic: same conj_kfold signature, computed by
    the oracle's restatement of the R loop on the selected parameter sets."""

    def __init__(self, port):
        self.port = port

    def conj_kfold(self, y, X, coords, m, cov, theta_alpha, sigmaSqIG, ho_idx, search_type="cb", set_range=None):
        from oracle import pyoracle as po
        ta = np.asarray(theta_alpha); g = ta.shape[1]
        lo, hi = set_range if set_range is not None else (0, g)
        out = np.zeros((g, 2))
        if hi > lo:
            out[lo:hi] = po.conj_kfold(self.port, self.port, y, X, coords, m, cov, ta[:, lo:hi], sigmaSqIG, ho_idx, search_type)
        return out


def _kfold_worker(rank, world, port_file, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle as po
    from spnngp_b200.dist import conj_kfold_sharded, make_allreduce
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    port = po.Port(); port.set_threads(1)
    d = np.load(os.path.join(ROOT, "tests", "golden", "kfold_n600_m10.npz"))
    ta = d["exp_theta_alpha"]
    got = conj_kfold_sharded(_OracleKfoldEngine(port), d["y"], d["X"], d["coords"], int(d["m"]), 0, ta, d["sigmaSqIG"],
                             d["ho_idx"], "cb", rank=rank, world=world, allreduce=make_allreduce(device=None))
    q.put((rank, float(np.max(np.abs(got - d["exp_scores"]) / np.abs(d["exp_scores"])))))
    dist.destroy_process_group()


def test_kfold_parameter_sets_shard_over_ranks_gloo(tmp_path):
    """world_size 2: each rank scores half of the theta.alpha grid; the summed matrices equal the golden scores."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    f = str(tmp_path / "rdzv_kfold")
    procs = [ctx.Process(target=_kfold_worker, args=(k, 2, f, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(e < 1e-12 for _, e in res), res


def test_rebalance_equalises_finishing_times():
    """dist.rebalance / balanced_shard_range: with a per-location cost that is higher at the head of the ordering (a
    model of the Matern slow branches) the measured-time rebalancing converges to shards that finish together, sizes
    stay a partition of [0, n), and equal costs leave the even split alone."""
    from spnngp_b200.dist import balanced_shard_range, rebalance, shard_range
    n, m, world = 1000000, 15, 8

    def cost(lo, hi):  # fixed launch cost + uniform rate + an expensive head of 4,000 locations
        head = max(0, min(hi, 4000) - lo)
        return 0.03 + 0.65e-6 * (hi - lo) + 4.5e-6 * head

    state = {}
    for rank in range(world):  # every "rank" runs the same deterministic procedure
        cur = {}
        all_bounds = {}

        def set_data(lo, hi, cur=cur):
            cur["lo"], cur["hi"] = lo, hi

        def allgather(x, rank=rank, cur=cur):
            # reconstruct every rank's current shard from the same procedure: sizes are identical on all ranks by
            # construction, so recompute them from this rank's view
            sizes = cur["sizes"]
            out = []
            lo = 0
            for r in range(world):
                out.append(cost(lo, lo + int(sizes[r])))
                lo += int(sizes[r])
            return out

        # drive the loop by hand to expose the sizes to the fake allgather
        sizes = np.array([hi - lo for lo, hi in (shard_range(n, m, r, world) for r in range(world))], dtype=np.int64)
        for _ in range(3):
            cur["sizes"] = sizes
            sizes = rebalance(sizes, allgather(0.0), n)
        state[rank] = sizes
        assert int(sizes.sum()) == n and np.all(sizes > 0)
    assert all(np.array_equal(state[0], state[r]) for r in range(world))
    sizes = state[0]
    lo = 0; t = []
    for r in range(world):
        t.append(cost(lo, lo + int(sizes[r]))); lo += int(sizes[r])
    even = [cost(r * n // world, (r + 1) * n // world) for r in range(world)]
    assert max(t) < max(even) - 0.008          # the slowest shard got faster
    assert (max(t) - min(t)) < 0.25 * (max(even) - min(even))
    same = rebalance([125000] * 8, [0.11] * 8, n)
    assert np.array_equal(same, [125000] * 8)

    # the orchestrating helper ends with the engine holding the final shard
    calls = []
    lo_hi = balanced_shard_range(None, n, m, 3, world, lambda a, b: calls.append((a, b)),
                                 lambda: 0.1, lambda x: [0.1] * world, rounds=2)
    assert calls[-1] == lo_hi and len(calls) == 3
