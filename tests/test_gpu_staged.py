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


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("words,expurgation", [(4, 0), (4, 5), (2, 0), (2, 40), (1, 3)])
def test_expurgated_two_pass_equals_reference(L, oracle, words, expurgation):
    """parallel_simulator_expurgated.py:238-243 at run sizes that take the staged pipeline: pass 1 lists the frames that end
    with erasures (ids travel in the spill records), pass 2 re-draws and counts exactly those.  Forced on for every
    point (eps = 0.43 fails often) and with forced cuts so that listed frames come out of every stage."""
    vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
    eps, seed, frames, its = [0.36, 0.40, 0.43], 99, 20000 + 33, 200
    want = oracle.best().mc_run(vl, cl, 1000, 500, 3, 6, eps, seed, 128, frames, its, expurgation, nthreads=16)
    for env in ({"LDPC_B200_EXP2PASS": 1}, {"LDPC_B200_EXP2PASS": 1, "LDPC_B200_FORCE_CUTS": "6,11,19,40"},
                {"LDPC_B200_EXP2PASS": 2}):
        g = L.Graph(vl, cl)
        g.set_tuning(words, 0, 0)
        got, st = _with_env(env, lambda: g.mc_run(eps, seed, frames, its, expurgation=expurgation, first_frame=128,
                                                  return_stats=True))
        assert np.array_equal(got, want), f"{env} words={words}"
        if env["LDPC_B200_EXP2PASS"] == 1:
            assert st["stages"] >= (5 if "LDPC_B200_FORCE_CUTS" in env else 1) and st["second_pass_groups"] == (0 if expurgation == 0 else st["second_pass_groups"])
        g.close()


def test_expurgated_two_pass_other_code_and_chunks(L, oracle):
    """(4,8) n=2000 (BASELINE config 4's code), several chunks (small spill budget), automatic choice of the strategy."""
    vl, cl = L.generate_graph(2000, 4, 8, seed=4, mode=1)
    eps, seed, frames, its = [0.31, 0.35], 5, 30000, 200
    want = oracle.best().mc_run(vl, cl, 2000, 1000, 4, 8, eps, seed, 0, frames, its, 0, nthreads=16)
    g = L.Graph(vl, cl)
    got, st = _with_env({"LDPC_B200_SPILL_MB": 2}, lambda: g.mc_run(eps, seed, frames, its, expurgation=0, return_stats=True))
    assert np.array_equal(got, want)
    assert st["kernel_launches"] > 10      # listing stages + one re-count launch per chunk
    # above the threshold nearly every frame fails: the strategy switches to the masked second pass, same counters
    hi = oracle.best().mc_run(vl, cl, 2000, 1000, 4, 8, [0.40], seed, 0, 9000, its, 0, nthreads=16)
    got_hi, st_hi = g.mc_run([0.40], seed, 9000, its, expurgation=0, return_stats=True)
    assert np.array_equal(got_hi, hi) and st_hi["stages"] == 1


def _env(**kv):
    class _E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kv}
            for k, v in kv.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = str(v)

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return _E()


@pytest.mark.parametrize("words,producers", [(4, 4), (2, 2), (2, 8), (1, 1)])
def test_overlapped_stage0_equals_reference_n1000(L, oracle, words, producers):
    """Producer / consumer stage 0 (e1_overlap.cuh), forced on a small graph: same counters as the reference loop,
    for several producer counts, ragged frame counts, eps incl. 0 and 1."""
    vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
    eps, seed, frames, its = [0.40, 0.0, 0.43, 1.0], 99, 24000 + 45, 200
    want = oracle.best().mc_run(vl, cl, 1000, 500, 3, 6, eps, seed, 128, frames, its, -1, nthreads=16)
    with _env(LDPC_B200_OVERLAP=1, LDPC_B200_PRODUCERS=producers, LDPC_B200_FORCE_CUTS="6,20"):
        g = L.Graph(vl, cl)
        g.set_tuning(words, 0, 0)
        got, st = g.mc_run(eps, seed, frames, its, first_frame=128, return_stats=True)
    assert st["overlap_producers"] == producers and st["words"] == words, st
    assert np.array_equal(got, want)
    with _env(LDPC_B200_OVERLAP=2):
        g = L.Graph(vl, cl)
        g.set_tuning(words, 0, 0)
        got2, st2 = g.mc_run(eps, seed, frames, its, first_frame=128, return_stats=True)
    assert st2["overlap_producers"] == 0
    assert np.array_equal(got2, want)


def test_overlapped_stage0_small_queue_and_expurgated(L, oracle):
    """a draw queue that overflows (words finish in-thread) and the expurgated two-pass run through the overlapped kernel"""
    vl, cl = L.generate_graph(2000, 4, 8, seed=3, mode=1)
    eps, seed, frames, its = [0.34, 0.37], 8, 20000, 120
    for exp in (-1, 0, 3):
        want = oracle.best().mc_run(vl, cl, 2000, 1000, 4, 8, eps, seed, 0, frames, its, exp, nthreads=16)
        with _env(LDPC_B200_OVERLAP=1, LDPC_B200_DRAW_QCAP=64, LDPC_B200_EXP2PASS=1, LDPC_B200_FORCE_CUTS="5,17"):
            g = L.Graph(vl, cl)
            got, st = g.mc_run(eps, seed, frames, its, expurgation=exp, return_stats=True)
        assert st["overlap_producers"] > 0, st
        assert np.array_equal(got, want), f"expurgation={exp}"


def test_overlapped_stage0_is_default_at_c2_size(L):
    """n = 10^4 (one team owns the SM): the overlapped kernel is the default and equals the plain staged kernel"""
    vl, cl = L.generate_graph(10000, 3, 6, seed=1, mode=1)
    eps, seed, frames, its = [0.40, 0.43], 5, 100000, 200
    g1 = L.Graph(vl, cl)
    a, sa = g1.mc_run(eps, seed, frames, its, return_stats=True)
    with _env(LDPC_B200_OVERLAP=2):
        g2 = L.Graph(vl, cl)
        b, sb = g2.mc_run(eps, seed, frames, its, return_stats=True)
    assert sa["overlap_producers"] > 0 and sb["overlap_producers"] == 0
    assert np.array_equal(a, b)


def test_spill_budget_follows_free_memory_and_fails_cleanly(L, oracle):
    """The spill buffers take at most a fifth of the free device memory (LDPC_B200_FAKE_FREE_MB pretends there is
    little): small chunks, same counters.  A budget that cannot be allocated is a clean error, and the handle still
    works afterwards."""
    vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
    eps, seed, frames, its = [0.40, 0.42], 77, 30000, 200
    want = oracle.best().mc_run(vl, cl, 1000, 500, 3, 6, eps, seed, 0, frames, its, -1, nthreads=16)
    with _env(LDPC_B200_FAKE_FREE_MB=5):      # 1 MB per buffer: 4096 frames per point and chunk
        g = L.Graph(vl, cl)
        got, st = g.mc_run(eps, seed, frames, its, return_stats=True)
    assert st["stages"] > 1 and np.array_equal(got, want)
    with _env(LDPC_B200_SPILL_MB=100000000):  # no cap: 40 points x 2^25 frames x 128 B = 172 GB per buffer
        g = L.Graph(vl, cl)
        with pytest.raises(L.LdpcB200Error) as ei:
            g.mc_run([0.30 + 0.001 * i for i in range(40)], seed, 1 << 34, its)
        assert "spill" in str(ei.value)
    got2 = g.mc_run(eps, seed, frames, its)
    assert np.array_equal(got2, want)


def test_background_placement_and_disk_cache(L, oracle, tmp_path):
    """The first launch runs while the placement search is still in the background (natural layout), later ones on the
    placed layout: identical counters.  The finished placement is cached on disk, a new handle reads it back
    (verified), and a corrupted file is ignored."""
    vl, cl = L.generate_graph(3000, 3, 6, seed=5, mode=1)
    eps, seed, frames, its = [0.40, 0.42], 3, 20000, 200
    want = oracle.best().mc_run(vl, cl, 3000, 1500, 3, 6, eps, seed, 0, frames, its, -1, nthreads=16)
    with _env(LDPC_B200_CACHE_DIR=str(tmp_path), LDPC_B200_CACHE=None, LDPC_B200_PLACE_SYNC=None):
        g = L.Graph(vl, cl)
        first, st = g.mc_run(eps, seed, frames, its, return_stats=True)      # search (0.5+ s) still running: natural layout
        assert np.array_equal(first, want)
        assert g.wait_placement(st["words"]) == 1                            # searched in this process
        info = g.placement_info(st["words"])
        assert info["gather_wavefronts_placed"] < 0.6 * info["gather_wavefronts_natural"]
        assert np.array_equal(g.mc_run(eps, seed, frames, its), want)        # placed layout
        files = [f for f in os.listdir(tmp_path) if f.startswith("place_")]
        assert len(files) == 1
        g2 = L.Graph(vl, cl)
        assert np.array_equal(g2.mc_run(eps, seed, frames, its), want)
        assert g2.wait_placement(st["words"]) == 2                           # read from the disk cache
        assert g2.placement_info(st["words"]) == info
        # a damaged cache file must not be trusted: flip lookups in place
        path = os.path.join(tmp_path, files[0])
        raw = bytearray(open(path, "rb").read())
        for k in range(200, 240):
            raw[-k] ^= 0x5A
        open(path, "wb").write(bytes(raw))
        g3 = L.Graph(vl, cl)
        assert np.array_equal(g3.mc_run(eps, seed, frames, its), want)
        assert g3.wait_placement(st["words"]) == 1                           # searched again
        assert np.array_equal(g3.mc_run(eps, seed, frames, its), want)
    with _env(LDPC_B200_PLACE_SYNC=1, LDPC_B200_CACHE=0):
        g4 = L.Graph(vl, cl)
        got, st4 = g4.mc_run(eps, seed, frames, its, return_stats=True)
        assert np.array_equal(got, want) and g4.wait_placement(st4["words"]) == 1
