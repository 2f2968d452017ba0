#!/usr/bin/env python
"""Generate tests/golden/analysis.json from the reference's own scalar analysis code and CSV filename parser.

Run where /root/reference exists:  python tests/golden/make_golden_analysis.py
The reference's scripts execute plots / file walks at import, so only the needed function definitions are exec'd
from the source files where they lie (nothing is copied into the repository):

  calc_threshold, below_threshold           /root/reference/test_de_threshold.py:7-28
  calculate_threshold_y, calculate_alpha    /root/reference/finite_length_scaling_calculation.py:9-21
  the FER formula Q(sqrt(n)(eps*-eps)/alpha) /root/reference/finite_length_scaling_calculation.py:38-44 (norm.cdf(-z/alpha))
  modified_density_evolution                /root/reference/tools/density_evolution.py:18-28
  get_parameters (CSV filename parser)      /root/reference/tools/combine_data.py:15-27
"""
import functools
import json
import os
import sys

import numpy as np
import scipy.stats as stats

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def exec_lines(path, first, last, ns):
    src = open(path).read().split("\n")[first - 1:last]
    exec("\n".join(src), ns)


def main():
    ns = {"np": np, "functools": functools, "stats": stats, "print": lambda *a, **k: None}
    exec_lines(os.path.join(REF, "test_de_threshold.py"), 1, 28, ns)
    exec_lines(os.path.join(REF, "finite_length_scaling_calculation.py"), 9, 21, ns)
    exec_lines(os.path.join(REF, "tools", "density_evolution.py"), 18, 28, ns)
    exec_lines(os.path.join(REF, "tools", "combine_data.py"), 15, 27, ns)
    out = {"thresholds": {}, "alpha": {}, "scaling_law": [], "modified_de": [], "filenames": []}
    for dv, dc in ((3, 6), (4, 8), (3, 4), (5, 10), (3, 5)):
        th = ns["calc_threshold"](dv, dc)
        out["thresholds"][f"{dv},{dc}"] = th
        out["alpha"][f"{dv},{dc}"] = float(ns["calculate_alpha"](th, dv, dc))
    for dv, dc in ((3, 6), (4, 8)):
        th = out["thresholds"][f"{dv},{dc}"]
        alpha = out["alpha"][f"{dv},{dc}"]
        for n in (500, 1000, 2000, 4000, 8000):
            for eps in (0.30, 0.34, 0.36, 0.38, 0.40, 0.42, 0.44):
                z = np.sqrt(n) * (th - eps)
                out["scaling_law"].append({"dv": dv, "dc": dc, "n": n, "eps": eps, "pb": float(stats.norm.cdf(-z / alpha))})
    for eps, dv, dc in ((0.40, 3, 6), (0.42, 3, 6), (0.36, 4, 8)):
        out["modified_de"].append({"eps": eps, "dv": dv, "dc": dc,
                                   "curve": [float(v) for v in ns["modified_density_evolution"](eps, 12, dv, dc)]})
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from iib_project_ldpc_codes_b200 import simulator as S
    for prefix, eps, n, k, dv, dc, its, num in (("regular_code_code_number=7", 0.42, 504, 252, 3, 6, 60, 3000),
                                               ("regular_code_expurgated=0", 0.36, 2000, 1000, 4, 8, 200, 123456789),
                                               ("regular_code_expurgated=-1", 0.4, 1000, 500, 3, 6, 200, 10000)):
        fn = S._filename(prefix, eps, n, k, dv, dc, its, num)
        out["filenames"].append({"filename": fn, "parsed": ns["get_parameters"](fn)})
    json.dump(out, open(os.path.join(HERE, "analysis.json"), "w"), indent=1)
    print("wrote analysis.json:", {k: len(v) for k, v in out.items()}, file=sys.stderr)


if __name__ == "__main__":
    main()
