"""GPU parity tests of the staged Monte-Carlo path (survivor re-packing between stages, e1_staged.cuh):
same counters, bit for bit, as the unstaged kernel and as the reference frame loop."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(built):
    import iib_project_ldpc_codes_b200 as L
    if L._lib.load().ldpc_b200_device_count() < 1:
        pytest.fail("no CUDA device")
    return L


def _run(g, env_stages, *args, **kw):
    old = os.environ.get("LDPC_B200_STAGES")
    if env_stages is None:
        os.environ.pop("LDPC_B200_STAGES", None)
    else:
        os.environ["LDPC_B200_STAGES"] = str(env_stages)
    try:
        return g.mc_run(*args, return_stats=True, **kw)
    finally:
        if old is None:
            os.environ.pop("LDPC_B200_STAGES", None)
        else:
            os.environ["LDPC_B200_STAGES"] = old


@pytest.mark.parametrize("words", [4, 2, 1])
def test_staged_equals_unstaged_and_reference_n1000(L, oracle, words):
    vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
    eps, seed, frames, its = [0.40, 0.43], 321, 40000 + 77, 200
    want = oracle.best().mc_run(vl, cl, 1000, 500, 3, 6, eps, seed, 256, frames, its, -1, nthreads=16)
    for stages in (1, 2, 4, 6):
        g = L.Graph(vl, cl)      # fresh handle: the stage plan is cached per handle
        g.set_tuning(words, 0, 0)
        got, st = _run(g, stages, eps, seed, frames, its, first_frame=256)
        assert st["words"] == words and st["stages"] <= stages
        if stages > 1:
            assert st["stages"] > 1
        assert np.array_equal(got, want), f"stages={stages} words={words}"


@pytest.mark.parametrize("its", [8, 9, 33])
def test_staged_small_caps(L, oracle, its):
    vl, cl = L.generate_graph(504, 3, 6, seed=2, mode=1)
    g = L.Graph(vl, cl)
    eps, seed, frames = [0.35, 0.42, 0.5], 11, 33000
    got, st = _run(g, 4, eps, seed, frames, its)
    want = oracle.best().mc_run(vl, cl, 504, 252, 3, 6, eps, seed, 0, frames, its, -1, nthreads=16)
    assert np.array_equal(got, want)


def test_staged_c2_size_and_inflation_drops(L):
    """n = 10^4: staged == unstaged bit for bit on 2*10^5 frames per point, with fewer physical iterations."""
    vl, cl = L.generate_graph(10000, 3, 6, seed=1, mode=1)
    eps, seed, frames, its = [0.40, 0.42, 0.43], 5, 200000, 200
    g1 = L.Graph(vl, cl)
    a, sa = _run(g1, 1, eps, seed, frames, its)
    g2 = L.Graph(vl, cl)
    b, sb = _run(g2, None, eps, seed, frames, its)
    assert sa["stages"] == 1 and sb["stages"] > 1
    assert np.array_equal(a, b)
    assert sb["phys_frame_iterations"] < 0.8 * sa["phys_frame_iterations"]


def test_staged_other_degrees(L, oracle):
    vl, cl = L.generate_graph(2000, 4, 8, seed=3, mode=1)
    g = L.Graph(vl, cl)
    eps, seed, frames, its = [0.34, 0.385], 8, 34000, 120
    got, st = _run(g, 3, eps, seed, frames, its)
    want = oracle.best().mc_run(vl, cl, 2000, 1000, 4, 8, eps, seed, 0, frames, its, -1, nthreads=16)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("qcap", [1, 9, 64])
def test_draw_queue_overflow_finishes_in_thread(L, qcap):
    """team_draw continues undecided words from a shared-memory queue; when the queue is full the word is finished
    in-thread (bec_finish).  A tiny queue (test hook LDPC_B200_DRAW_QCAP) must not change a single counter, in the
    unstaged, staged and masked (expurgation) paths."""
    vl, cl = L.generate_graph(504, 3, 6, seed=3, mode=1)
    eps, seed, its = [0.37, 0.44], 99, 60
    for frames, exp in ((40000 + 5, -1), (3000, -1), (3000, 2)):
        g = L.Graph(vl, cl)
        want = g.mc_run(eps, seed, frames, its, expurgation=exp)
        os.environ["LDPC_B200_DRAW_QCAP"] = str(qcap)
        try:
            got = L.Graph(vl, cl).mc_run(eps, seed, frames, its, expurgation=exp)
        finally:
            os.environ.pop("LDPC_B200_DRAW_QCAP", None)
        assert np.array_equal(got, want), f"qcap={qcap} frames={frames} expurgation={exp}"


def test_pilot_that_chooses_one_stage_leaves_no_trace(L, oracle):
    """First call on a handle runs the stage pilot; when the plan comes out as ONE stage the unstaged kernel runs
    right after it and must start from a clean work counter (regression: it skipped the groups the pilot had pulled)."""
    vl, cl = L.generate_graph(504, 3, 6, seed=2, mode=1)
    eps, seed, frames, its = [0.35, 0.42, 0.5], 11, 33000, 8
    want = oracle.best().mc_run(vl, cl, 504, 252, 3, 6, eps, seed, 0, frames, its, -1, nthreads=16)
    os.environ["LDPC_B200_REPACK_PCT"] = "100000"      # re-packing prohibitively expensive -> the plan has one stage
    try:
        g = L.Graph(vl, cl)
        got, st = g.mc_run(eps, seed, frames, its, return_stats=True)
        again = g.mc_run(eps, seed, frames, its)
    finally:
        os.environ.pop("LDPC_B200_REPACK_PCT", None)
    assert st["stages"] == 1
    assert np.array_equal(got, want) and np.array_equal(again, want)


@pytest.mark.parametrize("frames", [8192, 9001, 20000])
def test_small_runs_use_a_small_pilot_and_stay_exact(L, oracle, frames):
    """From 8192 frames per point the staged pipeline is on, with a pilot of an eighth of the run (1024..8192 frames):
    rough cut points from a tiny lifetime histogram must not change a single counter; a later, much larger run on the
    same handle re-draws the pilot."""
    vl, cl = L.generate_graph(1000, 3, 6, seed=4, mode=1)
    eps, seed, its = [0.38, 0.41, 0.44], 21, 200
    g = L.Graph(vl, cl)
    got, st = g.mc_run(eps, seed, frames, its, return_stats=True)
    want = oracle.best().mc_run(vl, cl, 1000, 500, 3, 6, eps, seed, 0, frames, its, -1, nthreads=16)
    assert st["stages"] > 1
    assert np.array_equal(got, want)
    big = 70000
    got2 = g.mc_run(eps, seed + 1, big, its)
    want2 = oracle.best().mc_run(vl, cl, 1000, 500, 3, 6, eps, seed + 1, 0, big, its, -1, nthreads=16)
    assert np.array_equal(got2, want2)
