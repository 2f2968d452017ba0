"""CPU tests: the oracle restatement (oracle/mp_oracle.c) is pinned against
  (a) the committed golden vectors produced by the UNMODIFIED reference C (tests/golden/make_golden.py),
  (b) the known answers of SURVEY.md 8c,
  (c) live differential runs against oracle/_ref when it is present (the build container).
"""
import numpy as np
import pytest

from golden_util import mc_cases, mp_cases, rng_words


def test_restatement_matches_reference_golden_single_frame(oracle):
    cases = list(mp_cases())
    assert len(cases) >= 100
    for c in cases:
        M, err, ret = oracle.single_frame("orc", c["inp"], c["its"], c["v2c"], c["c2v"], c["n"], c["k"], c["dv"],
                                          c["dc"], errors=c["prefill"])
        assert ret == c["ret"]
        assert np.array_equal(M, c["out"])
        assert np.array_equal(err, c["errors"])


def test_restatement_matches_reference_golden_monte_carlo(oracle):
    for c in mc_cases():
        got = oracle.restatement().mc_run(c["v2c"], c["c2v"], c["n"], c["k"], c["dv"], c["dc"], c["eps"], c["seed"],
                                          c["first"], c["frames"], c["its"], c["expurgation"], nthreads=3)
        assert np.array_equal(got, c["counters"])


def test_rng_restatement_golden_words(oracle):
    for seed, point, q, eps, words in rng_words():
        got = np.array([oracle.bec_word(seed, point, q, v, eps) for v in range(len(words))], dtype=np.uint32)
        assert np.array_equal(got, words)


def test_rng_threshold_semantics(oracle):
    assert oracle.bec_word(1, 0, 0, 5, 0.0) == 0
    assert oracle.bec_word(1, 0, 0, 5, 1.0) == 0xFFFFFFFF
    assert oracle.bec_word(1, 0, 0, 5, 1.5) == 0xFFFFFFFF
    assert oracle.bec_word(1, 0, 0, 5, -0.1) == 0
    p = oracle.draw_patterns(500, 0.4, 3, 0, 0, 640)
    assert set(np.unique(p)) == {0, 2}
    assert abs((p == 2).mean() - 0.4) < 5 * np.sqrt(0.24 / p.size)
    # frame f depends only on (seed, point, f): windows agree
    q = oracle.draw_patterns(500, 0.4, 3, 0, 96, 64)
    assert np.array_equal(p[96:160], q)


KAT_V2C = [0, 1, 0, 1, 0, 2, 1, 2, 0, 2, 1, 2]
KAT_C2V = [0, 1, 2, 4, 0, 1, 3, 5, 2, 3, 4, 5]


@pytest.mark.parametrize("mvc,its,want_mvc,want_err,want_ret", [
    ([1, 1, 2, 0, 1, 0], 5, [1, 1, 1, 0, 1, 0], [0] * 5, 0),
    ([2, 2, 2, 2, 1, 0], 6, [2, 2, 2, 2, 1, 0], [4] * 6, 6),
    ([2, 2, 0, 0, 1, 0], 1, [2, 2, 0, 0, 1, 0], [2], 1),
    ([0] * 6, 3, [0] * 6, [0, 0, 0], 0),
    ([2, 0, 1, 0, 0, 0], 3, [0, 0, 1, 0, 0, 0], [0, 0, 0], 0),
])
def test_survey_known_answers(oracle, mvc, its, want_mvc, want_err, want_ret):
    M, err, ret = oracle.single_frame("orc", mvc, its, KAT_V2C, KAT_C2V, 6, 3, 2, 4)
    assert (M.tolist(), err.tolist(), ret) == (want_mvc, want_err, want_ret)


def test_executed_iteration_rule(oracle):
    assert oracle.executed_iterations(3, [9, 5, 2, 0, 0, 0], 6) == 4          # success at it=3
    assert oracle.executed_iterations(6, [9, 5, 5, 5, 5, 5], 6) == 3          # first skipped index
    assert oracle.executed_iterations(6, [9, 9, 9, 9, 9, 9], 6) == 2          # iterations 0 and 1 always run
    assert oracle.executed_iterations(6, [9, 8, 7, 6, 5, 4], 6) == 6          # capped
    assert oracle.executed_iterations(0, [], 0) == 0


def test_live_differential_against_compiled_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (no /root/reference); golden vectors above pin the restatement")
    rng = np.random.default_rng(1)
    for (n, dv, dc) in [(96, 3, 6), (200, 4, 8), (1000, 3, 6)]:
        vl, cl, H = oracle.ref_generate_random_code(n, dv, dc, srand_seed=2)
        k = n - n * dv // dc
        M = np.where(rng.random((60, n)) < 0.42, 2, rng.integers(0, 2, (60, n))).astype(np.int8)
        for its in (0, 1, 2, 5, 200):
            a = oracle.reference().decode_batch(vl, cl, n, k, dv, dc, M, its)
            b = oracle.restatement().decode_batch(vl, cl, n, k, dv, dc, M, its)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        for exp in (-1, 0, 9):
            a = oracle.reference().mc_run(vl, cl, n, k, dv, dc, [0.38, 0.44], 5, 128, 200, 60, exp, nthreads=4)
            b = oracle.restatement().mc_run(vl, cl, n, k, dv, dc, [0.38, 0.44], 5, 128, 200, 60, exp, nthreads=2)
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n,dv,dc,seed", [(24, 3, 6, 1), (1000, 3, 6, 1), (2000, 4, 8, 7), (60, 3, 4, 3), (10000, 3, 6, 1)])
def test_oracle_graph_twin_equals_product_generator(oracle, n, dv, dc, seed):
    """oracle/graph_twin.h lets the CPU reference arm of bench.py build the benchmark graph without loading the product
    library; it must be the same graph."""
    import iib_project_ldpc_codes_b200 as L
    vo, co = oracle.generate_graph(n, dv, dc, seed)
    vp, cp = L.generate_graph(n, dv, dc, seed=seed, mode=1)
    assert np.array_equal(vo, vp) and np.array_equal(co, cp)


def test_as_shipped_loop_restatement_runs_the_reference_object(oracle):
    """oracle/as_shipped.py (the reference's Python frame loop around the unmodified .so) decodes like the C harness:
    below threshold at n=240 most frames succeed, above none does."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    from oracle import as_shipped as A
    vl, cl = oracle.generate_graph(240, 3, 6, 2)
    frames, sec, blk = A.run_fixed_loop(240, 3, 6, 0.30, 50, cl, vl, seconds=30.0, max_frames=60)
    assert frames == 60 and blk <= 6
    frames, sec, blk = A.run_fixed_loop(240, 3, 6, 0.55, 50, cl, vl, seconds=30.0, max_frames=20)
    assert frames == 20 and blk == 20
