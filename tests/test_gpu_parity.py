"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs.

Bar: bit-exact (integer work).  The checker is oracle.best(): the UNMODIFIED reference C
(oracle/_ref, built from /root/reference in the build container and shipped with the snapshot)
when present, else the plain-C restatement that is pinned against it (tests/test_oracle.py).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KAT_V2C = np.array([0, 1, 0, 1, 0, 2, 1, 2, 0, 2, 1, 2], dtype=np.int32)   # message_passing.c:181-182
KAT_C2V = np.array([0, 1, 2, 4, 0, 1, 3, 5, 2, 3, 4, 5], dtype=np.int32)


@pytest.fixture(scope="module")
def L(built):
    import iib_project_ldpc_codes_b200 as L
    if L._lib.load().ldpc_b200_device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return L


@pytest.fixture(scope="module")
def g1000(L):
    vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
    return L.Graph(vl, cl)


def _noisy(rng, frames, n, eps, values=True):
    base = rng.integers(0, 2, (frames, n)) if values else np.zeros((frames, n), dtype=np.int64)
    return np.where(rng.random((frames, n)) < eps, 2, base).astype(np.int8)


def _cmp_decode(L, O, g, M, its):
    got = g.decode_batch(M, its)
    want = O.best().decode_batch(g.variable_lookup, g.check_lookup, g.n, g.k, g.dv, g.dc, M, its)
    for a, b, name in zip(got, want, ("Mvc", "errors", "ret")):
        assert np.array_equal(a, b), f"{name} mismatch (its={its}, n={g.n})"


# ---- RNG: device generator == host restatement ---------------------------------------------------
@pytest.mark.parametrize("eps", [0.0, 1.0, 0.5, 0.25, 0.40, 0.43, 1e-9, 1 - 1e-9, 2.0 ** -32, 0.3834])
def test_draw_patterns_match_host_restatement(L, oracle, eps):
    n, frames, seed, point, first = 257, 70, 0xDEADBEEF12345678, 3, 64
    got = L.draw_patterns(n, eps, seed, point, first, frames)
    want = oracle.draw_patterns(n, eps, seed, point, first, frames)
    assert np.array_equal(got, want)
    if eps in (0.0, 1.0):
        assert (got == (2 if eps == 1.0 else 0)).all()


def test_draw_patterns_rate(L):
    p = L.draw_patterns(1000, 0.40, 1, 0, 0, 2048)
    rate = (p == 2).mean()
    assert abs(rate - 0.40) < 4 * np.sqrt(0.24 / p.size)


# ---- single-frame drop-in: the known answers of SURVEY.md 8c (from the compiled reference) --------
@pytest.mark.parametrize("mvc,its,want_mvc,want_err,want_ret", [
    ([1, 1, 2, 0, 1, 0], 1, [1, 1, 1, 0, 1, 0], [0], 0),
    ([1, 1, 2, 0, 1, 0], 2, [1, 1, 1, 0, 1, 0], [0, 0], 0),
    ([1, 1, 2, 0, 1, 0], 5, [1, 1, 1, 0, 1, 0], [0] * 5, 0),
    ([2, 2, 2, 2, 1, 0], 6, [2, 2, 2, 2, 1, 0], [4] * 6, 6),
    ([2, 2, 0, 0, 1, 0], 1, [2, 2, 0, 0, 1, 0], [2], 1),
    ([0, 0, 0, 0, 0, 0], 3, [0, 0, 0, 0, 0, 0], [0, 0, 0], 0),
    ([2, 0, 1, 0, 0, 0], 3, [0, 0, 1, 0, 0, 0], [0, 0, 0], 0),
])
def test_message_passing_kat(L, mvc, its, want_mvc, want_err, want_ret):
    M = np.array(mvc, dtype=np.int32)
    err = np.zeros(its, dtype=np.int32)
    ret = L.message_passing(M, its, KAT_V2C.copy(), KAT_C2V.copy(), err, 6, 3, 2, 4)
    assert ret == want_ret and M.tolist() == want_mvc and err.tolist() == want_err


def test_message_passing_prefilled_errors_and_zero_iterations(L, oracle):
    M = np.array([2, 2, 2, 2, 1, 0], dtype=np.int32)
    err = np.full(6, 100, dtype=np.int32)
    ret = L.message_passing(M, 6, KAT_V2C.copy(), KAT_C2V.copy(), err, 6, 3, 2, 4)
    assert ret == 6 and err.tolist() == [104] * 6            # += semantics, message_passing.c:73
    M = np.array([2, 2, 2, 2, 1, 0], dtype=np.int32)
    assert L.message_passing(M, 0, KAT_V2C.copy(), KAT_C2V.copy(), np.zeros(1, np.int32), 6, 3, 2, 4) == 0
    # non-uniform prefill changes the reference's stall test; replayed exactly
    rng = np.random.default_rng(5)
    vl, cl = L.generate_graph(96, 3, 6, seed=4)
    for trial in range(20):
        M0 = _noisy(rng, 1, 96, 0.45)[0].astype(np.int32)
        pre = rng.integers(0, 3, 12).astype(np.int32)
        a, ea = M0.copy(), pre.copy()
        ra = L.message_passing(a, 12, vl.ravel().copy(), cl.ravel().copy(), ea, 96, 48, 3, 6)
        b, eb, rb = oracle.single_frame("ref" if oracle.have_ref() else "orc", M0, 12, vl, cl, 96, 48, 3, 6, errors=pre)
        assert ra == rb and np.array_equal(a, b) and np.array_equal(ea, eb)


def test_message_passing_last_writer_wins(L):
    v2c = KAT_V2C.copy()
    M = np.array([2, 0, 1, 0, 0, 0], dtype=np.int32)
    v2c[0:2] = [1, 0]
    L.message_passing(M, 3, v2c, KAT_C2V.copy(), np.zeros(3, np.int32), 6, 3, 2, 4)
    assert M.tolist() == [1, 0, 1, 0, 0, 0]


@pytest.mark.parametrize("mvc,want,err,ret", [
    ([2, 0, 0, 0], [2, 0, 0, 0], [1, 1, 1, 1], 4),
    ([0, 2, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], 0),
    ([1, 2, 0, 0], [1, 0, 0, 0], [0, 0, 0, 0], 0),
    ([2, 2, 0, 0], [2, 0, 0, 0], [1, 1, 1, 1], 4),
    ([0, 0, 0, 2], [0, 0, 0, 2], [1, 1, 1, 1], 4),
])
def test_message_passing_multi_edge_kat(L, mvc, want, err, ret):
    """Exclusion is by socket position, not variable id: a variable listed twice in a check blocks itself."""
    c2v = np.array([0, 0, 1, 2, 1, 2, 3, 3], dtype=np.int32)
    v2c = np.array([0, 0, 0, 1, 0, 1, 1, 1], dtype=np.int32)
    M = np.array(mvc, dtype=np.int32)
    e = np.zeros(4, dtype=np.int32)
    r = L.message_passing(M, 4, v2c, c2v, e, 4, 2, 2, 4)
    assert (M.tolist(), e.tolist(), r) == (want, err, ret)


def test_bad_inputs_fail_loudly(L):
    with pytest.raises(L.LdpcB200Error):
        L.message_passing(np.array([3, 0, 0, 0, 0, 0], np.int32), 2, KAT_V2C.copy(), KAT_C2V.copy(),
                          np.zeros(2, np.int32), 6, 3, 2, 4)
    bad = KAT_V2C.copy(); bad[0] = 2   # variable 0 lists check 2 which has no socket for it
    with pytest.raises(L.LdpcB200Error):
        L.Graph(bad.reshape(6, 2), KAT_C2V.reshape(3, 4))


# ---- batched decode vs the reference C, frame by frame -------------------------------------------
@pytest.mark.parametrize("its", [0, 1, 2, 3, 17, 200])
def test_decode_batch_n1000(L, oracle, g1000, its):
    rng = np.random.default_rng(its)
    _cmp_decode(L, oracle, g1000, _noisy(rng, 77, 1000, 0.42), its)


def test_decode_batch_invalid_codewords_and_extremes(L, oracle, g1000):
    rng = np.random.default_rng(11)
    M = np.concatenate([_noisy(rng, 20, 1000, 0.30), _noisy(rng, 20, 1000, 0.55),
                        np.zeros((1, 1000), np.int8), np.full((1, 1000), 2, np.int8), np.ones((1, 1000), np.int8)])
    _cmp_decode(L, oracle, g1000, M, 60)


@pytest.mark.parametrize("n,dv,dc,eps", [(6, 2, 4, 0.3), (30, 3, 6, 0.4), (996, 3, 6, 0.41), (2000, 4, 8, 0.36),
                                         (120, 2, 4, 0.3), (60, 3, 4, 0.5), (100, 5, 10, 0.3), (4000, 3, 6, 0.42)])
def test_decode_batch_shapes(L, oracle, n, dv, dc, eps):
    rng = np.random.default_rng(n + dv)
    vl, cl = L.generate_graph(n, dv, dc, seed=n, mode=1)
    g = L.Graph(vl, cl)
    frames = 45 if n <= 2000 else 33
    _cmp_decode(L, oracle, g, _noisy(rng, frames, n, eps), 40)
    _cmp_decode(L, oracle, g, _noisy(rng, frames, n, eps, values=False), 40)


def test_decode_batch_c2_size_erasure_only(L, oracle):
    """BASELINE config 2 code size (n = 10^4): the value plane does not fit next to the graph in shared
    memory, the all-zero-codeword path (what the reference drivers send, parallel_simulator.py:222) does."""
    rng = np.random.default_rng(3)
    vl, cl = L.generate_graph(10000, 3, 6, seed=2, mode=1)
    g = L.Graph(vl, cl)
    _cmp_decode(L, oracle, g, _noisy(rng, 40, 10000, 0.42, values=False), 200)


def test_decode_batch_hand_built_multi_edge_graph(L, oracle):
    c2v = np.array([0, 0, 1, 2, 1, 2, 3, 3], dtype=np.int32).reshape(2, 4)
    v2c = np.array([0, 0, 0, 1, 0, 1, 1, 1], dtype=np.int32).reshape(4, 2)
    g = L.Graph(v2c, c2v)
    M = np.array([[a, b, c, d] for a in range(3) for b in range(3) for c in range(3) for d in range(3)], np.int8)
    _cmp_decode(L, oracle, g, M, 4)


# ---- fused Monte-Carlo vs the reference frame loop on the same patterns ----------------------------
@pytest.mark.parametrize("words,team_warps", [(4, 0), (4, 1), (2, 0), (2, 1), (1, 0), (1, 1), (1, 4)])
def test_mc_counters_n1000_all_geometries(L, oracle, g1000, words, team_warps):
    g = g1000
    g.set_tuning(words, 0, team_warps)
    try:
        eps, its, frames, seed = [0.40, 0.43], 200, 1500, 12345   # 1500: not a multiple of 32/64/128
        got, st = g.mc_run(eps, seed, frames, its, return_stats=True)
        want = oracle.best().mc_run(g.variable_lookup, g.check_lookup, g.n, g.k, 3, 6, eps, seed, 0, frames, its,
                                    -1, nthreads=8)
        assert st["words"] == words
        assert np.array_equal(got, want)
    finally:
        g.set_tuning(0, 0, 0)


@pytest.mark.parametrize("expurgation", [0, 5, 40, 10 ** 6])
def test_mc_expurgation(L, oracle, g1000, expurgation):
    eps, its, frames, seed = [0.38, 0.41], 100, 1000, 99
    got, st = g1000.mc_run(eps, seed, frames, its, expurgation=expurgation, return_stats=True)
    want = oracle.best().mc_run(g1000.variable_lookup, g1000.check_lookup, 1000, 500, 3, 6, eps, seed, 0, frames,
                                its, expurgation, nthreads=8)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("its", [0, 1, 2, 3, 7])
def test_mc_small_iteration_caps(L, oracle, g1000, its):
    eps, frames, seed = [0.42], 300, 5
    got = g1000.mc_run(eps, seed, frames, its)
    want = oracle.best().mc_run(g1000.variable_lookup, g1000.check_lookup, 1000, 500, 3, 6, eps, seed, 0, frames,
                                its, -1, nthreads=4)
    assert np.array_equal(got, want)


def test_mc_eps_extremes_and_offsets(L, oracle, g1000):
    its, seed = 30, 77
    got = g1000.mc_run([0.0, 1.0, 0.2], seed, 260, its, first_frame=128 * 5)
    want = oracle.best().mc_run(g1000.variable_lookup, g1000.check_lookup, 1000, 500, 3, 6, [0.0, 1.0, 0.2], seed,
                                128 * 5, 260, its, -1, nthreads=4)
    assert np.array_equal(got, want)
    with pytest.raises(L.LdpcB200Error):
        g1000.mc_run([0.4], seed, 10, its, first_frame=64)


def test_mc_additive_over_frame_ranges(L, g1000):
    """Counters are additive over disjoint frame ranges (the multi-GPU sharding relies on it)."""
    eps, its, seed = [0.40, 0.42], 200, 2024
    whole = g1000.mc_run(eps, seed, 4096 + 100, its)
    a = g1000.mc_run(eps, seed, 1280, its, first_frame=0)
    b = g1000.mc_run(eps, seed, 4096 + 100 - 1280, its, first_frame=1280)
    assert np.array_equal(whole, a + b)


@pytest.mark.parametrize("n,dv,dc,eps,frames", [(2000, 4, 8, [0.34, 0.37], 700), (500, 3, 6, [0.40], 900),
                                                (8000, 3, 6, [0.42], 200), (64, 2, 4, [0.3], 500),
                                                (10000, 3, 6, [0.40, 0.43], 96)])
def test_mc_other_codes(L, oracle, n, dv, dc, eps, frames):
    vl, cl = L.generate_graph(n, dv, dc, seed=7, mode=1)
    g = L.Graph(vl, cl)
    got = g.mc_run(eps, 31337, frames, 200)
    want = oracle.best().mc_run(vl, cl, n, g.k, dv, dc, eps, 31337, 0, frames, 200, -1, nthreads=8)
    assert np.array_equal(got, want)


def test_mc_statistical_anchor_density_evolution(L, g1000):
    """Mean erased fraction per iteration follows the reference's modified density evolution
    (tools/density_evolution.py:18-28) at n=1000, eps=0.40: 0.4, 0.313756, 0.267938, 0.236467."""
    r = g1000.results(g1000.mc_run([0.40], 1, 200000, 200), 200)[0]
    avg = r.average_errors
    for got, want in zip(avg[:4], [0.4, 0.313756, 0.267938, 0.236467]):
        assert abs(got - want) < 2e-3
    assert 0.05 < r.block_error_rate < 0.15      # FER ~ 0.085-0.087 (code dependent)


# ---- committed golden vectors (outputs of the unmodified reference C, tests/golden/make_golden.py) ----
def test_golden_single_frame_through_dropin_symbol(L):
    from golden_util import mp_cases
    for c in mp_cases():
        M = c["inp"].astype(np.int32)
        err = c["prefill"].astype(np.int32).copy()
        ret = L.message_passing(M, c["its"], c["v2c"].copy(), c["c2v"].copy(), err if c["its"] else np.zeros(1, np.int32),
                                c["n"], c["k"], c["dv"], c["dc"])
        assert ret == c["ret"] and np.array_equal(M, c["out"])
        if c["its"]:
            assert np.array_equal(err, c["errors"])


def test_golden_monte_carlo_counters(L):
    from golden_util import mc_cases
    for c in mc_cases():
        g = L.Graph(c["v2c"].reshape(c["n"], c["dv"]), c["c2v"].reshape(-1, c["dc"]))
        for words in (0, 1, 2, 4):
            g.set_tuning(words, 0, 0)
            got = g.mc_run(c["eps"], c["seed"], c["frames"], c["its"], expurgation=c["expurgation"],
                           first_frame=c["first"])
            assert np.array_equal(got, c["counters"]), (c["n"], c["its"], c["expurgation"], words)


def test_sharded_equals_whole(L, g1000):
    """The multi-GPU split (dist.shard_frames + one sum) reproduces the single-range counters bit for bit."""
    from iib_project_ldpc_codes_b200.dist import shard_frames
    eps, its, seed, total = [0.40, 0.43], 200, 4242, 5000
    whole = g1000.mc_run(eps, seed, total, its)
    for world in (2, 4, 8):
        acc = np.zeros_like(whole)
        for r in range(world):
            lo, cnt = shard_frames(total, r, world)
            if cnt:
                acc += g1000.mc_run(eps, seed, cnt, its, first_frame=lo)
        assert np.array_equal(acc, whole)


def test_point_block_partitioning_equals_whole(L, g1000):
    """(point, block) items (dist.shard_items) launched one erasure point at a time through
    ldpc_b200_mc_launch_points: the channel generator keys every point by its index in the FULL list, so the summed
    blocks equal the unsplit run bit for bit -- also through the staged pipeline and with expurgation."""
    import torch
    from iib_project_ldpc_codes_b200.dist import shard_items
    eps, its, seed = [0.36, 0.40, 0.43], 200, 77
    for total, exp in ((5000, -1), (20000, -1), (20000, 0)):
        whole = g1000.mc_run(eps, seed, total, its, expurgation=exp, first_frame=256)
        for world in (2, 8):
            d = torch.zeros((len(eps), its + 6), dtype=torch.int64, device="cuda")
            for r in range(world):
                for p, lo, cnt in shard_items(len(eps), total, r, world, first_frame=256):
                    g1000.mc_launch_point(eps[p], p, seed, cnt, its, d[p].data_ptr(), expurgation=exp, first_frame=lo)
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy().view(np.uint64), whole), (total, exp, world)


def test_full_size_properties_c2(L):
    """BASELINE config 2 at a larger frame count: size-independent invariants of the counter block."""
    vl, cl = L.generate_graph(10000, 3, 6, seed=1, mode=1)
    g = L.Graph(vl, cl)
    its, frames = 200, 200000
    c = g.mc_run([0.38, 0.47], 12345, frames, its)
    for row, eps in zip(c, (0.38, 0.47)):
        hist = row[: its + 1].astype(np.int64)
        assert (np.diff(hist) <= 0).all()                       # erasure counts never grow
        assert row[its + 3] == frames == row[its + 4]
        assert row[its + 2] == hist[-1]                         # bit errors = final residual sum
        assert abs(hist[0] / (frames * 10000) - eps) < 1e-4     # channel rate
        assert frames <= row[its + 5] <= frames * its           # executed iterations
    assert c[0][its + 1] == 0                                   # eps=0.38 << threshold: every frame decodes
    assert c[1][its + 1] == frames                              # eps=0.47 >> threshold 0.4294: none does
