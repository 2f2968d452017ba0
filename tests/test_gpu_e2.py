"""GPU parity tests of the HBM/L2-streamed engine (E2: cooperative kernel, frame-minor rows in global memory)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(built):
    import iib_project_ldpc_codes_b200 as L
    if L._lib.load().ldpc_b200_device_count() < 1:
        pytest.fail("no CUDA device")
    return L


def _noisy(rng, frames, n, eps, values=True):
    base = rng.integers(0, 2, (frames, n)) if values else np.zeros((frames, n), dtype=np.int64)
    return np.where(rng.random((frames, n)) < eps, 2, base).astype(np.int8)


@pytest.fixture(scope="module", params=["plain", "tma"])
def g1000_e2(L, request):
    """E2 forced on a small graph, with the lookups read by plain loads (the default) and streamed per warp through the
    TMA (cp.async.bulk + mbarrier, LDPC_B200_E2_TMA=1): every test using this fixture runs both ways."""
    import os
    vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
    g = L.Graph(vl, cl)
    g.set_engine(2)
    old = os.environ.get("LDPC_B200_E2_TMA")
    os.environ["LDPC_B200_E2_TMA"] = "1" if request.param == "tma" else "2"
    yield g
    if old is None:
        os.environ.pop("LDPC_B200_E2_TMA", None)
    else:
        os.environ["LDPC_B200_E2_TMA"] = old


@pytest.mark.parametrize("frames,its,exp", [(1500, 200, -1), (100, 200, -1), (1024, 3, -1), (2100, 60, 0), (700, 60, 25),
                                            (300, 0, -1), (300, 1, -1), (40, 2, -1)])
def test_e2_mc_counters_forced_on_small_graph(L, oracle, g1000_e2, frames, its, exp):
    g = g1000_e2
    eps, seed = [0.40, 0.43], 777
    got, st = g.mc_run(eps, seed, frames, its, expurgation=exp, return_stats=True)
    assert st["engine"] == 2
    want = oracle.best().mc_run(g.variable_lookup, g.check_lookup, 1000, 500, 3, 6, eps, seed, 0, frames, its, exp, nthreads=8)
    assert np.array_equal(got, want)


def test_e2_equals_e1(L, g1000_e2):
    vl, cl = g1000_e2.variable_lookup, g1000_e2.check_lookup
    g1 = L.Graph(vl, cl)
    eps = [0.39, 0.42]
    a, sa = g1.mc_run(eps, 5, 6000, 200, first_frame=256, return_stats=True)
    b, sb = g1000_e2.mc_run(eps, 5, 6000, 200, first_frame=256, return_stats=True)
    assert (sa["engine"], sb["engine"]) == (1, 2)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("its", [0, 1, 2, 9, 200])
def test_e2_decode_batch_values(L, oracle, g1000_e2, its):
    rng = np.random.default_rng(its + 100)
    M = _noisy(rng, 150, 1000, 0.42)
    got = g1000_e2.decode_batch(M, its)
    want = oracle.best().decode_batch(g1000_e2.variable_lookup, g1000_e2.check_lookup, 1000, 500, 3, 6, M, its)
    for a, b, name in zip(got, want, ("Mvc", "errors", "ret")):
        assert np.array_equal(a, b), name


def test_e2_hand_built_multi_edge_graph(L, oracle):
    c2v = np.array([0, 0, 1, 2, 1, 2, 3, 3], dtype=np.int32).reshape(2, 4)
    v2c = np.array([0, 0, 0, 1, 0, 1, 1, 1], dtype=np.int32).reshape(4, 2)
    g = L.Graph(v2c, c2v)
    g.set_engine(2)
    M = np.array([[a, b, c, d] for a in range(3) for b in range(3) for c in range(3) for d in range(3)], np.int8)
    got = g.decode_batch(M, 4)
    want = oracle.best().decode_batch(v2c, c2v, 4, 2, 2, 4, M, 4)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_mid_size_graph_takes_e2_automatically(L, oracle):
    """n = 30000 does not fit in shared memory (graph alone is 360 KB): automatic fallback to E2, values on."""
    n = 30000
    vl, cl = L.generate_graph(n, 3, 6, seed=3, mode=1)
    g = L.Graph(vl, cl)
    got, st = g.mc_run([0.41], 9, 96, 200, return_stats=True)
    assert st["engine"] == 2
    want = oracle.best().mc_run(vl, cl, n, n // 2, 3, 6, [0.41], 9, 0, 96, 200, -1, nthreads=8)
    assert np.array_equal(got, want)
    rng = np.random.default_rng(2)
    M = _noisy(rng, 20, n, 0.42)
    a = g.decode_batch(M, 100)
    b = oracle.best().decode_batch(vl, cl, n, n // 2, 3, 6, M, 100)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_c2_size_value_plane_goes_to_e2(L, oracle):
    """n = 10^4 with non-zero codewords: the value plane does not fit beside the graph in shared memory."""
    vl, cl = L.generate_graph(10000, 3, 6, seed=2, mode=1)
    g = L.Graph(vl, cl)
    rng = np.random.default_rng(8)
    M = _noisy(rng, 24, 10000, 0.42)
    a = g.decode_batch(M, 200)
    b = oracle.best().decode_batch(vl, cl, 10000, 5000, 3, 6, M, 200)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_c3_n_million_against_reference(L, oracle):
    """BASELINE config 3: (3,6)-regular n = 10^6, cap 200, sparse generator; the reference C decodes the same
    frames (~0.7 s per frame on a host core)."""
    n = 1_000_000
    vl, cl = L.generate_graph(n, 3, 6, seed=1, mode=1)
    g = L.Graph(vl, cl)
    got, st = g.mc_run([0.41], 2026, 12, 200, return_stats=True)
    assert st["engine"] == 2
    want = oracle.best().mc_run(vl, cl, n, n // 2, 3, 6, [0.41], 2026, 0, 12, 200, -1, nthreads=12)
    assert np.array_equal(got, want)
    pats = L.draw_patterns(n, 0.42, 4, 0, 0, 3)
    a = g.decode_batch(pats, 200)
    b = oracle.best().decode_batch(vl, cl, n, n // 2, 3, 6, pats, 200)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # full-batch size: invariants (1024 frames in lock step)
    c = g.mc_run([0.40], 1, 1024, 200)[0]
    hist = c[:201].astype(np.int64)
    assert (np.diff(hist) <= 0).all() and c[203] == 1024 and c[201] == 0 and hist[-1] == 0
    assert abs(hist[0] / (1024 * n) - 0.40) < 1e-4
