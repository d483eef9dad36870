"""GPU, >= 2 devices: two ranks (one process per GPU), locations sharded, likelihood pieces summed by the
engine's NCCL communicator -- equal to the single-GPU result; sampler chains identical on both ranks."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, rdzv, q, reduce="mailbox"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as po
    from spnngp_b200 import Engine, _lib
    from spnngp_b200.dist import shard_range
    from tests.util import make_problem
    torch.cuda.set_device(rank)
    os.environ["NNGP_GATHER_MIN"] = "1000"  # exercise the NVLink exchange of the replicated arrays at this small n
    os.environ["NNGP_REDUCE"] = reduce      # "nccl": sums over ranks by ncclAllReduce instead of the shared-memory mailbox
    dist.init_process_group("nccl", init_method=f"file://{rdzv}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    port = po.Port(); port.set_threads(2)
    prob = make_problem(6000, 15, seed=5, port=port, domain=1.1)
    n, m = 6000, 15
    eng = Engine(rank)
    eng.comm_init_from_torch()
    lo, hi = shard_range(n, m, rank, world)
    eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m, loc_range=(lo, hi))
    out = eng.bf_loglik(0, [[5.0, 1.0, 6.0], [4.0, 1.1, 7.0]], [prob["beta"], prob["beta"]], slots=[0, 1])
    S = eng.q_forms(0)
    qr = eng.q_resid(1, prob["beta"])
    R = _lib.load_rshim(); R.rshim_set_seed(3)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.03, "tau.sq": 0.03, "phi": 0.03}
    bs, ts, acc = eng.rnngp(0, [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], None, starting, tuning, 25, R.rnorm, R.runif)
    # k-fold scores with the parameter sets split over the ranks: a second engine WITHOUT the location-sharding
    # communicator (every rank fits the whole fold for its own sets), one torch allreduce at the end
    from spnngp_b200.dist import conj_kfold_sharded, make_allreduce
    d = np.load(os.path.join(ROOT, "tests", "golden", "kfold_n600_m10.npz"))
    eng2 = Engine(rank)
    scores = conj_kfold_sharded(eng2, d["y"], d["X"], d["coords"], int(d["m"]), 2, d["matern_theta_alpha"], d["sigmaSqIG"],
                                d["ho_idx"], "cb", rank=rank, world=world,
                                allreduce=make_allreduce(device=torch.device("cuda", rank)))
    eng2.close()
    q.put((rank, out, S, qr, ts, acc, scores))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("reduce", ["mailbox", "nccl"])
def test_two_rank_nccl_matches_single_gpu(tmp_path, port, reduce):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from spnngp_b200 import Engine, _lib
    from tests.util import make_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(k, 2, str(tmp_path / "rdzv"), q, reduce)) for k in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = make_problem(6000, 15, seed=5, port=port, domain=1.1)
    eng = Engine(0)
    eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], 15)
    out = eng.bf_loglik(0, [[5.0, 1.0, 6.0], [4.0, 1.1, 7.0]], [prob["beta"], prob["beta"]], slots=[0, 1])
    S = eng.q_forms(0); qr = eng.q_resid(1, prob["beta"])
    R = _lib.load_rshim(); R.rshim_set_seed(3)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.03, "tau.sq": 0.03, "phi": 0.03}
    bs, ts, acc = eng.rnngp(0, [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], None, starting, tuning, 25, R.rnorm, R.runif)
    for r in res:
        assert np.allclose(r[1], out, rtol=1e-12) and np.allclose(r[2], S, rtol=1e-11) and abs(r[3] - qr) < 1e-11 * abs(qr)
        assert r[5] == acc and np.allclose(r[4], ts, rtol=1e-10)
    assert np.array_equal(res[0][4], res[1][4])  # both ranks hold the same chain
    d = np.load(os.path.join(ROOT, "tests", "golden", "kfold_n600_m10.npz"))
    for r in res:  # parameter sets sharded over the ranks: every rank ends up with the reference's score matrix
        assert np.max(np.abs(r[6] - d["matern_scores"]) / np.abs(d["matern_scores"])) < 1e-10
    eng.close()


def _worker_shared_gpu(rank, world, rdzv, q, mode):
    """`world` ranks on ONE device (gloo rendezvous, mailbox-only communicator): the sharded fit of a 1-GPU lease."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle as po
    from spnngp_b200 import Engine, _lib
    from spnngp_b200.dist import shard_range
    from tests.util import make_problem
    dist.init_process_group("gloo", init_method=f"file://{rdzv}", rank=rank, world_size=world)
    port = po.Port(); port.set_threads(2)
    n, m = 6000, 15
    prob = make_problem(n, m, seed=5, port=port, domain=1.1)
    eng = Engine(0)
    eng.comm_init_from_torch(nccl=False)
    assert eng.get_option("mailbox_ranks") == world
    lo, hi = shard_range(n, m, rank, world)
    eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m, loc_range=(lo, hi))
    out = eng.bf_loglik(2, [[5.0, 1.0, 6.0, 0.75], [4.0, 1.1, 7.0, 0.8]], [prob["beta"], prob["beta"]], slots=[0, 1])
    cur = eng.bf_loglik_cur(2, [4.5, 1.2, 6.5, 0.7], prob["beta"] * 1.01, 1, 0)
    S = eng.q_forms(0)
    qr = eng.q_resid(0, prob["beta"])
    tot = eng.comm_allreduce(np.array([rank + 1.0, 2.0]))
    R = _lib.load_rshim(); R.rshim_set_seed(3)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.03, "tau.sq": 0.03, "phi": 0.03, "nu": 0.01}
    bs, ts, acc = eng.rnngp(2, [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], [0.5, 1.5], starting, tuning, 25, R.rnorm, R.runif,
                            mode=mode)
    q.put((rank, out, cur, S, qr, tot, bs, ts, acc))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, 1), (3, 0)])
def test_ranks_sharing_one_gpu_mailbox_matches_unsharded(tmp_path, port, world, mode):
    """Location shards on `world` ranks that share GPU 0, partial sums exchanged through the shared-memory mailbox
    (nngp_comm_init_local): likelihood pieces, cross forms and the Matern sampler chain equal the unsharded run, and
    every rank holds bit-identical chains (fixed rank-order sums).  Runs on a single-GPU lease."""
    import torch.multiprocessing as mp
    from spnngp_b200 import Engine, _lib
    from tests.util import make_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_shared_gpu, args=(k, world, str(tmp_path / "rdzv"), q, mode)) for k in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, m = 6000, 15
    prob = make_problem(n, m, seed=5, port=port, domain=1.1)
    eng = Engine(0)
    eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
    out = eng.bf_loglik(2, [[5.0, 1.0, 6.0, 0.75], [4.0, 1.1, 7.0, 0.8]], [prob["beta"], prob["beta"]], slots=[0, 1])
    cur = eng.bf_loglik_cur(2, [4.5, 1.2, 6.5, 0.7], prob["beta"] * 1.01, 1, 0)
    S = eng.q_forms(0); qr = eng.q_resid(0, prob["beta"])
    R = _lib.load_rshim(); R.rshim_set_seed(3)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.03, "tau.sq": 0.03, "phi": 0.03, "nu": 0.01}
    bs, ts, acc = eng.rnngp(2, [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], [0.5, 1.5], starting, tuning, 25, R.rnorm, R.runif,
                            mode=mode)
    assert acc > 0
    for r in res:
        assert np.allclose(r[1], out, rtol=1e-12) and np.allclose(r[2], cur, rtol=1e-12)
        assert np.allclose(r[3], S, rtol=1e-11) and abs(r[4] - qr) < 1e-11 * abs(qr)
        assert np.array_equal(r[5], [world * (world + 1) / 2, 2.0 * world])
        assert r[8] == acc and np.allclose(r[7], ts, rtol=1e-10) and np.allclose(r[6], bs, rtol=1e-8)
    for r in res[1:]:
        assert np.array_equal(r[7], res[0][7]) and np.array_equal(r[6], res[0][6])
