"""GPU parity: ordered neighbour search through the C ABI vs the CPU oracle -- bit-exact indices and
nnDist bits, on tie-free and on gridded (heavily tied) sites (SURVEY §8a tie notes)."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import coords_flat, n_indx
from spnngp_b200._lib import c_double_p, c_int_p

pytestmark = pytest.mark.gpu


def gpu_mk_nn(engine, coords, m, search):
    cf = coords_flat(coords); n = cf.size // 2
    nI = max(n_indx(n, m), 1)
    nnIndx = np.zeros(nI, dtype=np.int32); nnDist = np.zeros(nI); lu = np.zeros(2 * n, dtype=np.int32)
    ties = C.c_int(0)
    st = engine.lib.nngp_mk_nn_indx(engine.h, n, m, cf.ctypes.data_as(c_double_p), search,
                                    nnIndx.ctypes.data_as(c_int_p), nnDist.ctypes.data_as(c_double_p),
                                    lu.ctypes.data_as(c_int_p), C.byref(ties))
    assert st == 0, engine.lib.nngp_last_error(engine.h)
    return nnIndx, nnDist, lu, ties.value


def random_sites(n, seed, sort=True):
    rng = np.random.default_rng(seed)
    c = rng.random((n, 2))
    return c[np.argsort(c[:, 0], kind="stable")] if sort else c


def grid_sites(k):
    g = np.arange(k) / k
    xx, yy = np.meshgrid(g, g, indexing="ij")
    return np.column_stack([xx.ravel(), yy.ravel()])  # ordered by x: ties everywhere


@pytest.mark.parametrize("n,m", [(5000, 15), (3000, 32), (20, 5), (700, 1), (2000, 9)])
@pytest.mark.parametrize("search", [0, 1])
def test_knn_random_bit_exact(engine, port, n, m, search):
    coords = random_sites(n, n + m)
    ref = (port.mk_nn_indx if search == 0 else port.mk_nn_indx_cb)(coords, m)
    got = gpu_mk_nn(engine, coords, m, search)
    assert np.array_equal(got[0], ref[0])
    assert np.array_equal(got[1].view(np.int64), ref[1].view(np.int64))  # nnDist bits
    assert np.array_equal(got[2], ref[2])


def test_knn_unsorted_order(engine, port):
    """Arbitrary NNGP ordering (not x-sorted): predecessors are an arbitrary subset."""
    coords = random_sites(3000, 77, sort=False)
    for search, fn in ((0, port.mk_nn_indx), (1, port.mk_nn_indx_cb)):
        ref = fn(coords, 10)
        got = gpu_mk_nn(engine, coords, 10, search)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1].view(np.int64), ref[1].view(np.int64))


@pytest.mark.parametrize("m", [8, 15, 32])
@pytest.mark.parametrize("search", [0, 1])
def test_knn_grid_ties_bit_exact(engine, port, m, search):
    """40 x 40 grid: distance ties at almost every location, and ties in u = x + y for the code-book
    sort.  The reference's Shell-sort tie order is replayed, so indices still match exactly."""
    coords = grid_sites(40)
    ref = (port.mk_nn_indx if search == 0 else port.mk_nn_indx_cb)(coords, m)
    got = gpu_mk_nn(engine, coords, m, search)
    assert np.array_equal(got[0], ref[0])
    assert np.array_equal(got[1].view(np.int64), ref[1].view(np.int64))
    if search == 1:
        assert got[3] > 0  # u-ties detected -> exact host sort path taken


def test_brute_equals_cb_on_tie_free_sites(engine):
    """Documented invariant of the reference (man/spNNGP.Rd:110-119)."""
    coords = random_sites(20000, 5)
    a = gpu_mk_nn(engine, coords, 15, 0)
    b = gpu_mk_nn(engine, coords, 15, 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.int64), b[1].view(np.int64))


def test_neighbours_are_sorted_predecessors(engine):
    """Size-independent properties: ids < i, ascending distances, distances recomputable."""
    n, m = 50000, 15
    coords = random_sites(n, 9)
    idx, dist, lu, _ = gpu_mk_nn(engine, coords, m, 1)
    off, cnt = lu[:n], lu[n:]
    assert np.all(cnt == np.minimum(np.arange(n), m))
    rows = np.repeat(np.arange(n), cnt)
    assert np.all(idx < rows)
    d = np.sqrt(((coords[rows] - coords[idx]) ** 2).sum(1))
    assert np.allclose(d, dist, rtol=1e-15)
    full = np.arange(m, n)
    D = dist[off[m]:].reshape(-1, m)
    assert np.all(np.diff(D, axis=1) >= 0) and len(full) == D.shape[0]


@pytest.mark.parametrize("n,q,m", [(4000, 1500, 15), (300, 50, 32), (50, 7, 3)])
def test_nn_query_matches_exhaustive(engine, port, n, q, m):
    coords = random_sites(n, 3, sort=False)
    coords0 = np.random.default_rng(4).random((q, 2))
    ref = port.nn_query(coords, coords0, m)
    cf, c0 = coords_flat(coords), coords_flat(coords0)
    out = np.zeros(q * m, dtype=np.int32)
    st = engine.lib.nngp_nn_query(engine.h, n, cf.ctypes.data_as(c_double_p), q, c0.ctypes.data_as(c_double_p), m,
                                  out.ctypes.data_as(c_int_p))
    assert st == 0
    assert np.array_equal(out, ref)


def _nn_query(engine, coords, coords0, m):
    cf, c0 = coords_flat(coords), coords_flat(coords0)
    n, q = len(coords), len(coords0)
    out = np.zeros(q * m, dtype=np.int32)
    st = engine.lib.nngp_nn_query(engine.h, n, cf.ctypes.data_as(c_double_p), q, c0.ctypes.data_as(c_double_p), m,
                                  out.ctypes.data_as(c_int_p))
    assert st == 0
    return out


@pytest.mark.parametrize("case", ["tied_grid", "outside_box", "clustered", "collinear", "coincident", "few_sites", "big"])
def test_nn_query_grid_edge_cases(engine, port, case):
    """The grid search (knn_grid_query_kernel) against the exhaustive oracle where its geometry is stressed: exact
    distance ties (lower site index wins), queries far outside the sites' bounding box, empty cells between clusters,
    degenerate boxes (all sites on a line / in one point), fewer sites than neighbours asked for."""
    rng = np.random.default_rng(11)
    m = 10
    if case == "tied_grid":
        g = np.arange(30) / 30.0
        xx, yy = np.meshgrid(g, g, indexing="ij")
        coords = np.column_stack([xx.ravel(), yy.ravel()])
        coords0 = np.vstack([coords[::7] + 0.0, coords[5::11] + np.array([0.5 / 30, 0.5 / 30]), rng.random((200, 2))])
    elif case == "outside_box":
        coords = rng.random((3000, 2))
        coords0 = np.vstack([rng.random((100, 2)) * 5.0 - 2.0, np.array([[-50.0, 0.5], [0.5, 80.0], [1e3, -1e3]])])
    elif case == "clustered":
        coords = np.vstack([rng.normal(0.0, 0.01, (1500, 2)), rng.normal(5.0, 0.02, (1500, 2)), [[2.5, 2.5]]])
        coords0 = np.vstack([rng.normal(0.0, 0.05, (200, 2)), rng.random((200, 2)) * 5.0])
    elif case == "collinear":
        coords = np.column_stack([rng.random(800), np.full(800, 0.25)])
        coords0 = rng.random((150, 2))
    elif case == "coincident":
        coords = np.tile(np.array([[0.3, 0.7]]), (40, 1))
        coords0 = rng.random((20, 2))
    elif case == "few_sites":
        coords = rng.random((6, 2)); coords0 = rng.random((30, 2)); m = 6
    else:
        coords = rng.random((200000, 2)) * 6.0
        coords0 = rng.random((20000, 2)) * 6.0
        m = 15
    assert np.array_equal(_nn_query(engine, coords, coords0, m), port.nn_query(coords, coords0, m))


@pytest.mark.parametrize("n,m", [(5000, 15), (600, 10), (40, 32), (300, 1), (200000, 15)])
def test_u_indx_bit_exact(engine, port, n, m):
    """nngp_mk_u_indx (reference `mkUIndx`, src/nn.cpp:25-56) vs the oracle, all three `u.search.type` values."""
    rng = np.random.default_rng(n + m)
    c = rng.random((n, 2)); c = c[np.argsort(c[:, 0], kind="stable")]
    idx, _, lu = engine.mk_nn_indx(c, m, "cb")
    for st in ((0, 1, 2) if n <= 5000 else (2,)):
        u, ulu, ui = engine.mk_u_indx(n, m, idx, search_type=st)
        uo, uluo, uio = port.mk_u_indx(n, m, idx, lu, search_type=st)
        nI = len(idx) if n > 1 else 0
        assert np.array_equal(u[:nI], uo[:nI]) and np.array_equal(ulu, uluo) and np.array_equal(ui[:nI], uio[:nI])
    # definition: location u[k] has i as its ui[k]-th neighbour
    u, ulu, ui = engine.mk_u_indx(n, m, idx)
    for i in rng.integers(0, n, 50):
        for k in range(ulu[i], ulu[i] + ulu[n + i]):
            assert idx[lu[u[k]] + ui[k]] == i
