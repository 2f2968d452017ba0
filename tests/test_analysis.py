"""CPU tests of the host-side analysis and file-format mirror against golden values produced
by the reference's own code (tests/golden/make_golden_analysis.py -> tests/golden/analysis.json)."""
import csv
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "analysis.json")))


@pytest.fixture()
def S(tmp_path):
    from iib_project_ldpc_codes_b200 import simulator as S
    S.base_directory = str(tmp_path) + os.sep
    return S


def test_bp_thresholds_equal_reference_bisection(S):
    """test_de_threshold.py:17-28: (3,6) -> 0.429440, (4,8) -> 0.383447, and three more degree pairs."""
    for key, want in GOLD["thresholds"].items():
        dv, dc = (int(v) for v in key.split(","))
        assert S.calc_threshold(dv, dc) == pytest.approx(want, abs=2e-9)
    assert GOLD["thresholds"]["3,6"] == pytest.approx(0.429440, abs=1e-6)
    assert GOLD["thresholds"]["4,8"] == pytest.approx(0.383447, abs=1e-6)


def test_scaling_law_equals_reference_formula(S):
    """finite_length_scaling_calculation.py:9-21,38-44: alpha and P_B = Q(sqrt(n)(eps* - eps)/alpha) at the C5 sizes."""
    for key, want in GOLD["alpha"].items():
        dv, dc = (int(v) for v in key.split(","))
        assert float(S.calculate_alpha(S.calc_threshold(dv, dc), dv, dc)) == pytest.approx(want, rel=1e-6)
    for row in GOLD["scaling_law"]:
        pb, th, alpha = S.scaling_law_block_error(row["n"], row["dv"], row["dc"], row["eps"])
        assert pb == pytest.approx(row["pb"], rel=1e-5, abs=1e-300)


def test_modified_density_evolution_equals_reference(S):
    for row in GOLD["modified_de"]:
        got = S.modified_density_evolution(row["eps"], len(row["curve"]) - 1, row["dv"], row["dc"])
        assert np.allclose(got, row["curve"], rtol=0, atol=1e-15)


def _get_parameters(filename):
    """tools/combine_data.py:15-27, restated: split on '_', key=value tokens, the token holding '.csv' loses its suffix."""
    out = {}
    for tok in filename.split("_"):
        if "=" in tok and ".csv" not in tok:
            out[tok.split("=")[0]] = tok.split("=")[1]
        elif ".csv" in tok:
            out[tok.split("=")[0]] = tok.split("=")[1].split(".")[0]
    return out


def test_csv_and_filename_are_what_the_reference_combiner_parses(S):
    """f1: a CSV written by the mirror is consumable by tools/combine_data.py unchanged -- its filename parser (:15-27)
    finds BEC, n, k, dv, dc, it, num (golden: the reference's own get_parameters on filenames the mirror produced), and
    its row loop (:76-84: rows of length 1, re-weighted by num*n, rounded) recovers the integer counters exactly."""
    for row in GOLD["filenames"]:
        assert _get_parameters(row["filename"]) == row["parsed"]        # the restated parser == the reference's
    its, n, frames = 60, 504, 3000
    rng = np.random.default_rng(1)
    sums = np.sort(rng.integers(0, n * frames, its + 1))[::-1].astype(np.uint64)
    counters = np.concatenate([sums, np.array([17, int(sums[-1]), frames, frames, 12345], dtype=np.uint64)])
    r = S._finish("regular_code_code_number=7", 0.42, n, n // 2, 3, 6, its, counters, True)
    p = _get_parameters(r["filename"])
    assert (p["BEC"], p["n"], p["k"], p["dv"], p["dc"], p["it"], p["num"], p["number"]) == \
           ("0.42", "504", "252", "3", "6", "60", "3000", "7")
    path = os.path.join(S.base_directory, "report_data", "simulation_data", r["filename"])
    recovered = []
    with open(path) as f:
        for row in csv.reader(f, delimiter=","):
            if len(row) != 1:
                break                                                   # combine_data.py:80-81
            recovered.append(round(float(row[0]) * int(p["num"]) * int(p["n"])))
    assert recovered == [int(v) for v in sums]                          # :82 re-weighting is exact
    rows = list(csv.reader(open(path)))
    assert rows[its + 1] == ["Message passing block-wise error", str(17 / frames)]
    assert rows[its + 2][0] == "Message passing bit-wise error"


def test_fixed_code_cache_writes_the_three_files_the_reference_loads(S, monkeypatch):
    """f3: parallel_simulator_expurgated.py:303-348 looks for code_no_*, check_code_no_*, variable_code_no_* and
    regenerates unless it finds all three; the mirror writes all three (dense bool H [(n-k), n], int32 lookups)."""
    from iib_project_ldpc_codes_b200 import api

    class FakeGraph:          # no GPU in this test: the cache logic runs before any device call
        def __init__(self, *a, **k):
            raise RuntimeError("stop here")
    monkeypatch.setattr(api, "Graph", FakeGraph)
    p = {"BEC": 0.4, "num_tests": 10, "iterations": 5, "n": 120, "dv": 3, "dc": 6, "optimal": False,
         "message_passing": True, "filenumber": 3}
    with pytest.raises(RuntimeError, match="stop here"):
        S.run_simulation_fixed_ldpc(p)
    pdir = os.path.join(S.base_directory, "parity_checks")
    names = sorted(os.listdir(pdir))
    assert names == ["check_code_no_3_n_120_dv_3_dc_6.npy", "code_no_3_n_120_dv_3_dc_6.npy", "variable_code_no_3_n_120_dv_3_dc_6.npy"]
    H = np.load(os.path.join(pdir, names[1]))
    cl = np.load(os.path.join(pdir, names[0]))
    vl = np.load(os.path.join(pdir, names[2]))
    assert H.dtype == np.bool_ and H.shape == (60, 120) and cl.dtype == np.int32 and vl.dtype == np.int32
    assert (H.sum(axis=1) == 6).all() and (H.sum(axis=0) == 3).all()
    assert all(H[c, v] for c, row in enumerate(cl.reshape(60, 6)) for v in row)
    assert all(H[c, i] for i, row in enumerate(vl.reshape(120, 3)) for c in row)
