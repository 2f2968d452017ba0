"""GPU tests of the host-side mirror of the reference drivers (iib_project_ldpc_codes_b200/simulator.py)."""
import csv
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def S(built, tmp_path):
    import iib_project_ldpc_codes_b200 as L
    if L._lib.load().ldpc_b200_device_count() < 1:
        pytest.fail("no CUDA device")
    from iib_project_ldpc_codes_b200 import simulator as S
    S.base_directory = str(tmp_path) + os.sep
    return S


def test_message_pass_decode_matches_reference_wrapper_semantics(S, oracle):
    """(int32 sequence, errors with the initial count prepended) as parallel_simulator_expurgated.py:131-167."""
    import iib_project_ldpc_codes_b200 as L
    n, dv, dc = 512, 3, 6
    vl, cl = L.generate_graph(n, dv, dc, seed=5)
    code = S.regular_LDPC_code(None, n, n // 2, dv, dc)
    np.random.seed(3)
    ch = S.BEC(0.41)
    for _ in range(10):
        y = ch.new_transmit(np.zeros(n))
        dec, errors = code.message_pass_decode(y, 50, cl, vl)
        M, err, ret = oracle.single_frame("ref" if oracle.have_ref() else "orc", y.astype(np.int32), 50, vl, cl, n, n // 2, dv, dc)
        assert dec.dtype == np.int32 and np.array_equal(dec, M)
        assert len(errors) == 51 and errors[0] == np.count_nonzero(y == 2) and np.array_equal(errors[1:], err)


def test_message_pass_decode_from_dense_parity_check(S):
    import iib_project_ldpc_codes_b200 as L
    n, dv, dc = 60, 3, 6
    vl, cl = L.generate_graph(n, dv, dc, seed=8)
    H = np.zeros((n // 2, n), dtype=np.int8)
    for c in range(n // 2):
        H[c, cl[c]] = 1
    code = S.regular_LDPC_code(H, n, n // 2, dv, dc)
    y = np.zeros(n); y[[1, 7, 30]] = 2
    dec, errors = code.message_pass_decode(y, 20)
    assert errors[0] == 3 and (dec == 0).all() and errors[-1] == 0


def test_run_simulation_fixed_ldpc_counters_csv_and_cache(S, oracle):
    p = {"BEC": 0.42, "num_tests": 3000, "iterations": 60, "n": 504, "dv": 3, "dc": 6, "optimal": False,
         "message_passing": True, "filenumber": 7}
    r = S.run_simulation_fixed_ldpc(p, stop_rule_batch=1024, run_seed=11)
    pdir = os.path.join(S.base_directory, "parity_checks")
    cl = np.load(os.path.join(pdir, "check_code_no_7_n_504_dv_3_dc_6.npy")).reshape(252, 6)
    vl = np.load(os.path.join(pdir, "variable_code_no_7_n_504_dv_3_dc_6.npy")).reshape(504, 3)
    frames = r["num_tests"]
    want = oracle.best().mc_run(vl, cl, 504, 252, 3, 6, [0.42], 11, 0, frames, 60, -1, nthreads=8)[0]
    assert np.array_equal(r["counters"], want)
    # exact ordered truncation (parallel_simulator_expurgated.py:367): the run ends with the frame of the 200th failure
    assert r["block_errors"] == 200 and frames < 3000
    before = oracle.best().mc_run(vl, cl, 504, 252, 3, 6, [0.42], 11, 0, frames - 1, 60, -1, nthreads=8)[0]
    assert int(before[61]) == 199
    assert r["density_evolution"]["max_abs_deviation"] < 0.02   # mean curve follows modified DE (tools/plotting.py:86)
    # batch-granular variant overshoots by less than one batch
    rb = S.run_simulation_fixed_ldpc(p, stop_rule_batch=1024, run_seed=11, write_csv=False, exact_stop=False)
    assert rb["block_errors"] >= 200 and rb["num_tests"] % 1024 == 0 and rb["num_tests"] - frames < 1024
    # CSV: iterations+1 rows, then the two labelled rows (parallel_simulator.py:26-32); filename schema
    path = os.path.join(S.base_directory, "report_data", "simulation_data", r["filename"])
    rows = list(csv.reader(open(path)))
    assert len(rows) == 63
    assert rows[61][0] == "Message passing block-wise error" and rows[62][0] == "Message passing bit-wise error"
    assert abs(float(rows[0][0]) - 0.42) < 0.01
    assert float(rows[61][1]) == r["block_error"]
    for key in ("regular_code_code_number=7", "_BEC=0.42", "_n=504", "_k=252", "_dv=3", "_dc=6", "_it=60", f"_num={frames}"):
        assert key in r["filename"]
    # second run reuses the cached code files
    r2 = S.run_simulation_fixed_ldpc(p, stop_rule_batch=1024, run_seed=11, write_csv=False)
    assert np.array_equal(r2["counters"], r["counters"])


def _ensemble_reference(oracle, L, n, dv, dc, eps, seed, graph_seed, first_graph, graphs, fpg, its, exp, gen_mode=1):
    F = (fpg + 127) // 128 * 128
    want = np.zeros((len(eps), its + 6), dtype=np.uint64)
    for g in range(first_graph, first_graph + graphs):
        vl, cl = L.generate_graph(n, dv, dc, seed=graph_seed * 1000003 + g, mode=gen_mode)
        want += oracle.best().mc_run(vl, cl, n, n - n * dv // dc, dv, dc, eps, seed, g * F, fpg, its, exp, nthreads=4)
    return want


@pytest.mark.parametrize("n,dv,dc,fpg,graphs,exp", [(240, 3, 6, 128, 9, 0), (240, 3, 6, 128, 9, -1), (504, 3, 6, 100, 5, 3),
                                                     (1000, 3, 6, 300, 4, -1), (2000, 4, 8, 128, 6, 0), (96, 3, 6, 32, 40, -1)])
def test_mc_run_ensemble_matches_reference_per_graph(S, oracle, n, dv, dc, fpg, graphs, exp):
    """Fresh graph per frame group (parallel_simulator_expurgated.py:200-243): the counters equal the reference frame
    loop run graph by graph on the same graphs and erasure patterns."""
    import iib_project_ldpc_codes_b200 as L
    eps, seed, gseed, its = [0.36, 0.42], 77, 5, 60
    got, st = L.mc_run_ensemble(n, dv, dc, eps, seed, graphs, its, frames_per_graph=fpg, graph_seed=gseed, first_graph=3,
                                expurgation=exp, return_stats=True)
    want = _ensemble_reference(oracle, L, n, dv, dc, eps, seed, gseed, 3, graphs, fpg, its, exp)
    assert np.array_equal(got, want)
    # additive over graph ranges
    a = L.mc_run_ensemble(n, dv, dc, eps, seed, 2, its, frames_per_graph=fpg, graph_seed=gseed, first_graph=3, expurgation=exp)
    b = L.mc_run_ensemble(n, dv, dc, eps, seed, graphs - 2, its, frames_per_graph=fpg, graph_seed=gseed, first_graph=5,
                          expurgation=exp)
    assert np.array_equal(a + b, got)


@pytest.mark.parametrize("n,dv,dc,fpg,graphs,exp", [(240, 3, 6, 128, 12, -1), (1000, 3, 6, 128, 6, 0), (2000, 4, 8, 128, 8, 0),
                                                     (2000, 4, 8, 256, 3, -1), (96, 3, 6, 32, 50, 2), (60, 3, 4, 64, 20, -1),
                                                     (600, 5, 10, 128, 5, -1)])
def test_mc_run_ensemble_in_kernel_generator(S, oracle, n, dv, dc, fpg, graphs, exp):
    """gen_mode 2: the kernel builds every graph itself in shared memory (Philox keys, bitonic sort, repair);
    the host twin rebuilds the same graphs and the reference decodes the same frames on them."""
    import iib_project_ldpc_codes_b200 as L
    eps, seed, gseed, its, first = [0.35, 0.43], 99, 0xC0FFEE1234, 60, 7
    got, st = L.mc_run_ensemble(n, dv, dc, eps, seed, graphs, its, frames_per_graph=fpg, graph_seed=gseed, gen_mode=2,
                                first_graph=first, expurgation=exp, return_stats=True)
    F = (fpg + 127) // 128 * 128
    want = np.zeros_like(got)
    for g in range(first, first + graphs):
        vl, cl = L.generate_graph_indexed(n, dv, dc, gseed, g)
        srt = np.sort(cl, axis=1)
        assert (srt[:, 1:] != srt[:, :-1]).all() and (np.bincount(cl.ravel(), minlength=n) == dv).all()
        want += oracle.best().mc_run(vl, cl, n, n - n * dv // dc, dv, dc, eps, seed, g * F, fpg, its, exp, nthreads=4)
    assert np.array_equal(got, want)


def _indexed_reference(oracle, L, n, dv, dc, eps, seed, gseed, total_frames, fpg, its, exp):
    """Reference frame loop over the graphs of the in-kernel generator: graph G decodes frames [G*F, G*F + fpg),
    the last graph only as many as are left of `total_frames`."""
    F = (fpg + 127) // 128 * 128
    want = np.zeros(its + 6, dtype=np.uint64)
    g, left = 0, total_frames
    while left > 0:
        cnt = min(fpg, left)
        vl, cl = L.generate_graph_indexed(n, dv, dc, gseed, g)
        want += oracle.best().mc_run(vl, cl, n, n - n * dv // dc, dv, dc, [eps], seed, g * F, cnt, its, exp, nthreads=4)[0]
        g += 1
        left -= cnt
    return want, g


def test_run_simulation_ensemble_expurgated(S, oracle):
    import iib_project_ldpc_codes_b200 as L
    p = {"BEC": 0.40, "expurgation": 0, "num_tests": 1000, "iterations": 40, "n": 240, "dv": 3, "dc": 6,
         "optimal": False, "message_passing": True, "seed": 3}
    r = S.run_simulation(p, write_csv=False, stop_rule_graphs=4, gen_mode=1, max_block_errors=10 ** 9)
    assert r["num_tests"] == 1000 and r["graphs"] == 8      # 7 whole graphs + 104 frames of the eighth
    want = _ensemble_reference(oracle, L, 240, 3, 6, [0.40], 3, 3, 0, 7, 128, 40, 0)[0]
    vl, cl = L.generate_graph(240, 3, 6, seed=3 * 1000003 + 7, mode=1)
    want = want + oracle.best().mc_run(vl, cl, 240, 120, 3, 6, [0.40], 3, 7 * 128, 104, 40, 0, nthreads=4)[0]
    assert np.array_equal(r["counters"], want)
    assert "regular_code_expurgated=0_BEC=0.4_n=240" in r["filename"]


def test_run_simulation_default_generator_is_the_in_kernel_one(S, oracle):
    """run_simulation(dict) exactly as parallel_simulator_expurgated.py:418-428 builds the dict, default gen_mode (2):
    the counters equal the reference frame loop on the host twins of the graphs the kernel generated."""
    import iib_project_ldpc_codes_b200 as L
    p = {"BEC": 0.36, "expurgation": 0, "num_tests": 2000, "iterations": 60, "n": 240, "dv": 3, "dc": 6,
         "optimal": False, "message_passing": True, "seed": 3}
    r = S.run_simulation(p, write_csv=False)
    assert r["num_tests"] == 2000
    want, graphs = _indexed_reference(oracle, L, 240, 3, 6, 0.36, 3, 3, 2000, 128, 60, 0)
    assert r["graphs"] == graphs and np.array_equal(r["counters"], want)
    assert r["bp_threshold"] == pytest.approx(0.429440, abs=2e-6) and "scaling_law" in r and "density_evolution" in r


def test_cli_mode0_ensemble_with_expurgation(S, oracle):
    """CLI mode 0 (parallel_simulator_expurgated.py:418-428): eps num_tests iterations n dv dc 0 seed expurgation."""
    import iib_project_ldpc_codes_b200 as L
    r = S.main(["prog", "0.36", "2000", "60", "240", "3", "6", "0", "3", "0"])
    want, graphs = _indexed_reference(oracle, L, 240, 3, 6, 0.36, 3, 3, 2000, 128, 60, 0)
    assert r["num_tests"] == 2000 and r["graphs"] == graphs and np.array_equal(r["counters"], want)
    rows = list(csv.reader(open(os.path.join(S.base_directory, "report_data", "simulation_data", r["filename"]))))
    assert len(rows) == 63 and rows[61][0] == "Message passing block-wise error"
    # a different expurgation threshold changes what is counted, not the denominators
    r5 = S.main(["prog", "0.36", "2000", "60", "240", "3", "6", "0", "3", "5"])
    want5, _ = _indexed_reference(oracle, L, 240, 3, 6, 0.36, 3, 3, 2000, 128, 60, 5)
    assert np.array_equal(r5["counters"], want5) and r5["num_tests"] == 2000


@pytest.mark.parametrize("fpg", [128, 300])
def test_ensemble_stop_rule_is_exact(S, oracle, fpg):
    """:200 `while block_error<200`: the run ends with the frame of the 200th counted failure, also inside a graph."""
    import iib_project_ldpc_codes_b200 as L
    p = {"BEC": 0.44, "expurgation": 2, "num_tests": 100000, "iterations": 40, "n": 240, "dv": 3, "dc": 6,
         "optimal": False, "message_passing": True, "seed": 9}
    r = S.run_simulation(p, write_csv=False, frames_per_graph=fpg, stop_rule_graphs=3)
    assert r["block_errors"] == 200 and r["num_tests"] < 100000
    want, graphs = _indexed_reference(oracle, L, 240, 3, 6, 0.44, 9, 9, r["num_tests"], fpg, 40, 2)
    assert np.array_equal(r["counters"], want) and r["graphs"] == graphs
    shorter, _ = _indexed_reference(oracle, L, 240, 3, 6, 0.44, 9, 9, r["num_tests"] - 1, fpg, 40, 2)
    assert int(shorter[41]) == 199      # the last frame of the run is the 200th counted block error


def test_cli_modes(S):
    r = S.main(["prog", "0.35", "256", "30", "120", "3", "6", "3", "1"])
    assert r["num_tests"] == 256
    with pytest.raises(NotImplementedError):
        S.main(["prog", "0.35", "256", "30", "120", "3", "6", "4", "1"])
    with pytest.raises(ValueError):
        S.main(["prog", "0.35", "256", "30", "120", "3", "6", "9", "1"])
