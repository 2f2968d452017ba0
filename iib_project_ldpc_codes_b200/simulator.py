"""Host-side mirror of the reference drivers for the message-passing path, over the GPU library.

Same names, argument meaning and outputs as the reference (file:line relative to the reference repo):

  BEC                         channels.py:4-26                           (host numpy; the fused path draws in-kernel)
  regular_LDPC_code           parallel_simulator_expurgated.py:44-53
    .message_pass_decode      parallel_simulator_expurgated.py:131-167   -> (int32 sequence, errors[max_its+1])
  run_simulation              parallel_simulator_expurgated.py:169-285   (ensemble mode, expurgation filter)
  run_simulation_fixed_ldpc   parallel_simulator_expurgated.py:287-414   (fixed code, .npy cache, no filter)
  write_message_passing_file  parallel_simulator.py:26-32                (CSV rows + filename schema)
  main(argv)                  parallel_simulator_expurgated.py:418-459   eps num_tests iterations n dv dc mode <seed|filenumber> [expurgation]

Out of scope here: the ML decoder (`optimal_decode`, modes 1/2/4/5 raise NotImplementedError).

Differences that follow from batching (documented, not hidden):
  * the 200-block-error stop rule (:200, :367) is evaluated between batches of frames; in fixed-code mode the
    batch that crosses the limit is then truncated EXACTLY at the frame of the 200th failure by bisection on the
    deterministic counters (`exact_stop=True`), in ensemble mode the overshoot is at most `stop_rule_graphs` graphs;
  * ensemble mode draws one fresh graph per FRAME GROUP (`frames_per_graph`, default 128) instead of one
    per frame: the expectation of every statistic over the ensemble is unchanged (frames are exchangeable),
    frames sharing a graph are correlated through the code only;
  * erasures come from the counter-based in-kernel generator keyed by (seed, eps point, frame), so runs ARE
    reproducible (the reference ignores its seed, random_code_generator.c:22-25, and uses unseeded numpy).
"""
from __future__ import annotations

import csv
import os
import sys
import time
from datetime import datetime

import numpy as np

from . import api

base_directory = os.environ.get("LDPC_B200_BASE_DIRECTORY", os.getcwd() + os.sep)


def _sim_dir():
    d = os.path.join(base_directory, "report_data", "simulation_data")
    os.makedirs(d, exist_ok=True)
    return d


def write_message_passing_file(filename, errors, message_passing_block_error, message_passing_bit_error):
    """parallel_simulator.py:26-32: iterations+1 rows of average erased fraction, then two labelled rows."""
    with open(os.path.join(_sim_dir(), filename), "w", newline="") as csvfile:
        writer = csv.writer(csvfile)
        for error_at_iteration in errors:
            writer.writerow([error_at_iteration])
        writer.writerow(["Message passing block-wise error", message_passing_block_error])
        writer.writerow(["Message passing bit-wise error", message_passing_bit_error])


class BEC:
    """channels.py:4-26."""

    def __init__(self, erasure_prob):
        self.erasure_prob = erasure_prob

    def transmit(self, input_binary):
        input_binary = np.where(input_binary == 0, -1, input_binary)
        random_vector = np.random.rand(len(input_binary))
        return np.where(random_vector < self.erasure_prob, 0, input_binary)

    def new_transmit(self, input_binary):
        random_vector = np.random.rand(len(input_binary))
        return np.where(random_vector < self.erasure_prob, 2, input_binary)


class regular_LDPC_code:
    """parallel_simulator_expurgated.py:44-53; the dense parity_check is carried but never needed by the
    message-passing path when the lookups are passed (and may be None here)."""

    def __init__(self, parity_check, n, k, dv, dc):
        self.parity_check = parity_check
        self.n, self.k, self.dv, self.dc = n, k, dv, dc
        self.rate = self.k / self.n

    def _lookups_from_H(self):
        H = np.asarray(self.parity_check)
        check_lookup = [list(np.nonzero(row == 1)[0]) for row in H]
        variable_lookup = [list(np.nonzero(col == 1)[0]) for col in H.T]
        return check_lookup, variable_lookup

    def message_pass_decode(self, binary_sequence, max_its, check_lookup=None, variable_lookup=None):
        if check_lookup is None or variable_lookup is None:
            cl, vl = self._lookups_from_H()
            check_lookup = cl if check_lookup is None else check_lookup
            variable_lookup = vl if variable_lookup is None else variable_lookup
        check_lookup = np.array(check_lookup, dtype="int32").flatten()
        variable_lookup = np.array(variable_lookup, dtype="int32").flatten()
        errors = np.zeros(max(max_its, 1), dtype="int32")
        binary_sequence = np.array(binary_sequence, dtype="int32")
        initial_error_count = int(np.count_nonzero(binary_sequence == 2))
        api.message_passing(binary_sequence, max_its, variable_lookup, check_lookup, errors, self.n, self.k, self.dv, self.dc)
        errors = np.insert(errors[:max_its], 0, initial_error_count)
        return binary_sequence, errors

    def optimal_decode(self, binary_sequence):
        raise NotImplementedError("ML decoding (ml_decoder.c + galois) is outside the accelerated hot path")


def _filename(prefix, eps, n, k, dv, dc, iterations, num):
    return (prefix + "_BEC=" + str(eps) + "_n=" + str(n) + "_k=" + str(k) + "_dv=" + str(dv) + "_dc=" + str(dc) +
            "_it=" + str(iterations) + "_num=" + str(num) + "_time=" + datetime.now().strftime("%d-%m-%Y-%H-%M-%S") + ".csv")


def _finish(prefix, eps, n, k, dv, dc, iterations, counters, write_csv):
    its = iterations
    frames = int(counters[its + 3])
    average_errors = counters[: its + 1].astype(np.float64) / (n * max(frames, 1))
    block = int(counters[its + 1]) / max(frames, 1)
    bit = int(counters[its + 2]) / (max(frames, 1) * n)
    filename = _filename(prefix, eps, n, k, dv, dc, iterations, frames)
    if write_csv:
        write_message_passing_file(filename, average_errors, block, bit)
    return {"filename": filename, "average_errors": average_errors, "block_error": block, "bit_error": bit,
            "num_tests": frames, "block_errors": int(counters[its + 1]), "bit_errors": int(counters[its + 2]),
            "counters": counters}


def modified_density_evolution(erasure_prob, iterations, dv, dc):
    """Expected erased fraction per flooding iteration of the reference's one-state-per-variable decoder
    (restates tools/density_evolution.py:18-28): the per-variable residual uses all dv checks."""
    results = [erasure_prob]
    last = erasure_prob
    for _ in range(iterations):
        q = 1.0 - (1.0 - last) ** (dc - 1)
        results.append(erasure_prob * q ** dv)
        last = erasure_prob * q ** (dv - 1)
    return np.array(results)


def density_evolution_overlay(average_errors, eps, dv, dc, first=8):
    """Built-in validation the reference did by eye (tools/plotting.py:86): simulated mean curve vs. modified
    density evolution over the first iterations (finite-length effects grow later)."""
    k = min(first, len(average_errors))
    de = modified_density_evolution(eps, k - 1, dv, dc)
    dev = np.abs(np.asarray(average_errors[:k], dtype=np.float64) - de)
    return {"iterations_compared": int(k), "max_abs_deviation": float(dev.max()), "de_curve": de.tolist()}


def _de_below_threshold(erasure_prob, dv, dc, iterations=100000, tolerance=1e-6):
    """test_de_threshold.py:7-15: does the density-evolution recursion of the (dv,dc) ensemble die out?"""
    prob = 1.0
    for _ in range(iterations):
        last = prob
        prob = erasure_prob * (1.0 - (1.0 - last) ** (dc - 1)) ** (dv - 1)
        if prob < tolerance * 1e-3 or last - prob < 1e-16 * max(prob, 1e-300):   # decided: vanished / reached its fixed point
            break
    return prob < tolerance


def calc_threshold(dv, dc):
    """BP threshold of the (dv,dc)-regular ensemble on the BEC by bisection (restates test_de_threshold.py:17-28;
    (3,6) -> 0.429440, (4,8) -> 0.383447)."""
    lo, hi = 0.0, 1.0
    while hi - lo > 1e-9:
        mid = 0.5 * (lo + hi)
        if _de_below_threshold(mid, dv, dc):
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def calculate_threshold_y(threshold_erasure, dv, dc):
    """finite_length_scaling_calculation.py:9-16: check-to-variable erasure probability at the critical point."""
    previous, value = 0.0, 1.0
    while abs(value - previous) > 1e-6:
        previous = value
        value = 1.0 - (1.0 - threshold_erasure * value ** (dv - 1)) ** (dc - 1)
    return value


def calculate_alpha(threshold, dv, dc):
    """finite_length_scaling_calculation.py:18-21: the scaling parameter alpha of the waterfall."""
    y = calculate_threshold_y(threshold, dv, dc)
    x = threshold * y ** (dv - 1)
    return threshold * np.sqrt(((dv - 1) / dv) * (1.0 / x - 1.0 / y))


_threshold_cache = {}


def scaling_law_block_error(n, dv, dc, erasure_prob):
    """P_B ~ Q(sqrt(n) (eps* - eps) / alpha)  (finite_length_scaling_calculation.py:23-29,38-44, without the
    n^(-2/3) shift, as the reference evaluates it)."""
    from math import erfc, sqrt
    if (dv, dc) not in _threshold_cache:
        th = calc_threshold(dv, dc)
        _threshold_cache[(dv, dc)] = (th, float(calculate_alpha(th, dv, dc)))
    th, alpha = _threshold_cache[(dv, dc)]
    z = sqrt(n) * (th - erasure_prob)
    return 0.5 * erfc(z / alpha / sqrt(2.0)), th, alpha


def validation_overlays(res, eps, n, dv, dc):
    """The checks the reference did by eye, attached to every run's result: mean erasure curve against modified
    density evolution (tools/plotting.py:86), and the measured block error rate against the finite-length scaling
    law around the BP threshold (finite_length_scaling_calculation.py:38-44, test_de_threshold.py:17-28)."""
    pb, th, alpha = scaling_law_block_error(n, dv, dc, eps)
    return {"density_evolution": density_evolution_overlay(res["average_errors"], eps, dv, dc),
            "bp_threshold": th,
            "scaling_law": {"alpha": alpha, "predicted_block_error": pb, "measured_block_error": res["block_error"],
                            "z": float(np.sqrt(n) * (th - eps))}}


def run_simulation_fixed_ldpc(parameter_set, write_csv=True, device=0, run_seed=None, stop_rule_batch=None,
                              wall_clock_limit=42000, exact_stop=True, max_block_errors=200,
                              dense_code_limit_bytes=1 << 27):
    """Fixed-code Monte-Carlo (parallel_simulator_expurgated.py:287-414).

    Code cache (:303-348): `check_code_no_*`, `variable_code_no_*` (int32 lookups) and `code_no_*` (the dense bool
    parity-check matrix [(n-k), n]) in base_directory/parity_checks, the three files the reference looks for -- a
    later reference run finds all of them and decodes the same code.  The dense file is only needed by the
    reference (the GPU path never reads it); it is written when it is at most `dense_code_limit_bytes`."""
    eps = parameter_set["BEC"]
    num_tests, iterations = parameter_set["num_tests"], parameter_set["iterations"]
    n, dv, dc = parameter_set["n"], parameter_set["dv"], parameter_set["dc"]
    if parameter_set.get("optimal"):
        raise NotImplementedError("ML decoding is outside the accelerated hot path")
    filenumber = parameter_set["filenumber"]
    k = int(n * (dc - dv) / dc)
    start = time.time()
    pdir = os.path.join(base_directory, "parity_checks")
    os.makedirs(pdir, exist_ok=True)
    tag = str(filenumber) + "_n_" + str(n) + "_dv_" + str(dv) + "_dc_" + str(dc) + ".npy"
    check_path = os.path.join(pdir, "check_code_no_" + tag)
    var_path = os.path.join(pdir, "variable_code_no_" + tag)
    code_path = os.path.join(pdir, "code_no_" + tag)
    if os.path.exists(check_path) and os.path.exists(var_path):          # :303-319 (the GPU path needs the lookups only)
        check_lookup = np.load(check_path)
        variable_lookup = np.load(var_path)
    else:                                                                 # :322-348
        variable_lookup, check_lookup = api.generate_graph(n, dv, dc, seed=int(filenumber), mode=1)
        np.save(check_path, check_lookup.ravel().astype(np.int32))
        np.save(var_path, variable_lookup.ravel().astype(np.int32))
    if not os.path.exists(code_path) and (n - k) * n <= dense_code_limit_bytes:   # :342-344, for the reference's loader
        H = np.zeros((n - k, n), dtype=bool)
        H[np.repeat(np.arange(n - k), dc), np.asarray(check_lookup).reshape(-1)] = True
        np.save(code_path, H)
    g = api.Graph(np.asarray(variable_lookup).reshape(n, dv), np.asarray(check_lookup).reshape(n - k, dc), device=device)
    seed = int(filenumber) if run_seed is None else run_seed
    total = np.zeros(iterations + 6, dtype=np.uint64)
    done = 0
    batch = stop_rule_batch or max(4096, min(num_tests, 1 << 22))
    batch = (batch + 127) // 128 * 128
    # :367  while block_error<200 and i<num_tests and elapsed < 42000
    while int(total[iterations + 1]) < max_block_errors and done < num_tests and time.time() - start < wall_clock_limit:
        cnt = min(batch, num_tests - done)
        part = g.mc_run([eps], seed, cnt, iterations, expurgation=-1, first_frame=done)[0]
        if exact_stop and int(total[iterations + 1]) + int(part[iterations + 1]) >= max_block_errors:
            # The reference stops right after the frame that brings block_error to 200 (:367).  Frames are
            # keyed by their index, so the shortest prefix of this batch reaching the limit is found by
            # bisection on the (deterministic, additive) counters: exact ordered truncation.
            need = max_block_errors - int(total[iterations + 1])
            lo, hi = 1, cnt          # smallest L in [1, cnt] with block_errors(done, done+L) >= need
            while lo < hi:
                mid = (lo + hi) // 2
                if int(g.mc_run([eps], seed, mid, iterations, expurgation=-1, first_frame=done)[0][iterations + 1]) >= need:
                    hi = mid
                else:
                    lo = mid + 1
            cnt = lo
            part = g.mc_run([eps], seed, cnt, iterations, expurgation=-1, first_frame=done)[0]
        total += part
        done += cnt
    g.close()
    res = _finish("regular_code_code_number=" + str(filenumber), eps, n, k, dv, dc, iterations, total, write_csv)
    res.update(validation_overlays(res, eps, n, dv, dc))
    return res


def run_simulation(parameter_set, write_csv=True, device=0, frames_per_graph=128, stop_rule_graphs=4096,
                   wall_clock_limit=43000, gen_mode=2, exact_stop=True, max_block_errors=200):
    """Ensemble Monte-Carlo (parallel_simulator_expurgated.py:169-285): a fresh random graph per frame group.

    gen_mode 2 (default): graph G is built by the kernel itself in shared memory (host twin:
    generate_graph_indexed(n, dv, dc, seed, G)); gen_mode 0/1: generate_graph(n, dv, dc, seed*1000003 + G, gen_mode) on
    a pool of host threads.  Graph G decodes the global frames [G*F, G*F + frames_per_graph), F = frames_per_graph
    rounded up to 128 (ldpc_b200_mc_run_ensemble).

    Stop rule (:200 `while block_error<200 and i<num_tests`): checked every `stop_rule_graphs` graphs; the batch
    that crosses the limit is then truncated by bisection on the deterministic counters -- to the graph of the
    200th counted failure and, inside that graph, to the frame of the 200th failure (exact ordered truncation;
    frames are keyed by their global index so a shorter run of the same graph decodes the same frames)."""
    eps = parameter_set["BEC"]
    num_tests, iterations = parameter_set["num_tests"], parameter_set["iterations"]
    n, dv, dc = parameter_set["n"], parameter_set["dv"], parameter_set["dc"]
    if parameter_set.get("optimal"):
        raise NotImplementedError("ML decoding is outside the accelerated hot path")
    if gen_mode == 2 and n * dv > 65535:     # the in-kernel generator sorts 16-bit socket ids
        gen_mode = 1
    seed = int(parameter_set["seed"])
    expurgation = int(parameter_set.get("expurgation", -1))
    k = int(n * (dc - dv) / dc)
    start = time.time()
    its = iterations
    fpg = int(frames_per_graph)
    stride = (fpg + 127) // 128 * 128

    def ens(first_graph, graphs, frames=fpg):
        return api.mc_run_ensemble(n, dv, dc, [eps], seed, graphs, iterations, frames_per_graph=frames,
                                   graph_seed=seed, gen_mode=gen_mode, first_graph=first_graph,
                                   expurgation=expurgation, device=device, frame_stride=stride)[0]

    total = np.zeros(iterations + 6, dtype=np.uint64)
    done, gi = 0, 0
    while int(total[its + 1]) < max_block_errors and done < num_tests and time.time() - start < wall_clock_limit:   # :200
        # whole graphs, except the last one of the run: num_tests frames in total
        graphs = min(stop_rule_graphs, (num_tests - done + fpg - 1) // fpg)
        last_frames = min(fpg, num_tests - done - (graphs - 1) * fpg)
        if last_frames == fpg:
            part = ens(gi, graphs)
        else:   # the run ends inside the last graph of this batch
            part = ens(gi + graphs - 1, 1, last_frames)
            if graphs > 1:
                part = part + ens(gi, graphs - 1)
        frames_here = (graphs - 1) * fpg + last_frames
        if exact_stop and int(total[its + 1]) + int(part[its + 1]) >= max_block_errors:
            need = max_block_errors - int(total[its + 1])
            lo, hi = 1, graphs           # smallest G with block_errors(first G graphs of this batch) >= need
            while lo < hi:
                mid = (lo + hi) // 2
                if int(ens(gi, mid)[its + 1]) >= need:
                    hi = mid
                else:
                    lo = mid + 1
            before = ens(gi, lo - 1) if lo > 1 else np.zeros_like(total)
            need -= int(before[its + 1])
            top = last_frames if lo == graphs else fpg
            a, b = 1, top                # smallest L with block_errors(first L frames of graph gi+lo-1) >= need
            while a < b:
                mid = (a + b) // 2
                if int(ens(gi + lo - 1, 1, mid)[its + 1]) >= need:
                    b = mid
                else:
                    a = mid + 1
            part = before + ens(gi + lo - 1, 1, a)
            frames_here = (lo - 1) * fpg + a
            graphs = lo
        total += part
        done += frames_here
        gi += graphs
    res = _finish("regular_code_expurgated=" + str(expurgation), eps, n, k, dv, dc, iterations, total, write_csv)
    res["graphs"] = gi
    res.update(validation_overlays(res, eps, n, dv, dc))
    return res


def main(argv=None):
    """parallel_simulator_expurgated.py:418-459 positional CLI."""
    argv = sys.argv if argv is None else argv
    erasure_prob = float(argv[1]); num_tests = int(argv[2]); iterations = int(argv[3])
    n = int(argv[4]); dv = int(argv[5]); dc = int(argv[6]); mode = int(argv[7])
    expurgation = int(argv[9]) if len(argv) > 9 else -1
    if mode == 0:
        d = {"BEC": erasure_prob, "expurgation": expurgation, "num_tests": num_tests, "iterations": iterations, "n": n,
             "dv": dv, "dc": dc, "optimal": False, "message_passing": True, "seed": int(argv[8])}
        return run_simulation(d)
    if mode == 3:
        d = {"BEC": erasure_prob, "num_tests": num_tests, "iterations": iterations, "n": n, "dv": dv, "dc": dc,
             "optimal": False, "message_passing": True, "filenumber": int(argv[8])}
        return run_simulation_fixed_ldpc(d)
    if mode in (1, 2, 4, 5):
        raise NotImplementedError("modes 1,2,4,5 use the ML decoder, which is outside the accelerated hot path")
    raise ValueError("Mode value must be in the range 0-5")


if __name__ == "__main__":
    r = main()
    print(r["filename"], "block", r["block_error"], "bit", r["bit_error"], "num", r["num_tests"])
