"""oracle/as_shipped.py -- TEST INFRASTRUCTURE (CPU baseline), not product code.

The reference's fixed-code frame loop AS IT SHIPS, restated around the unmodified compiled decoder
(oracle/_ref/message_passing.so): the Python-side cost per frame is part of what a user of the reference pays.

  BEC.new_transmit                          /root/reference/channels.py:24-26
  regular_LDPC_code.message_pass_decode     /root/reference/parallel_simulator_expurgated.py:131-167
      (np.array(...).flatten() of both lookups, zeroed errors, ct.CDLL re-opened, np.insert -- every frame)
  run_simulation_fixed_ldpc loop body       /root/reference/parallel_simulator_expurgated.py:367-392

SURVEY.md 8(d) CPU-baseline row (ii).  The reference's own driver cannot be imported unmodified (matplotlib / galois
missing, hard-coded paths, runs at import: SURVEY.md 7.2) and /root/reference does not exist on the GPU box, so the
loop is restated here, statement by statement, with numpy's unseeded global generator exactly as the reference.
Single process, single core (the reference's parallelism is separate OS processes).
"""
from __future__ import annotations

import ctypes as ct
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
base_directory = os.path.join(HERE, "_ref") + os.sep


class BEC:                                                    # channels.py:4-26
    def __init__(self, erasure_prob):
        self.erasure_prob = erasure_prob

    def new_transmit(self, input_binary):                     # :24-26
        random_vector = np.random.rand(len(input_binary))
        return np.where(random_vector < self.erasure_prob, 2, input_binary)


def message_pass_decode(n, k, dv, dc, binary_sequence, max_its, check_lookup, variable_lookup, lib_path=None):
    """parallel_simulator_expurgated.py:144-167, lookups given."""
    check_lookup = np.array(check_lookup, dtype='int32').flatten()
    variable_lookup = np.array(variable_lookup, dtype='int32').flatten()
    errors = np.zeros(max_its, dtype='int32')
    initial_error_count = len(np.nonzero(binary_sequence == 2)[0])
    binary_sequence = np.array(binary_sequence, dtype='int32')
    binary_sequence_p = binary_sequence.ctypes.data_as(ct.POINTER(ct.c_int))
    check_lookup_p = check_lookup.ctypes.data_as(ct.POINTER(ct.c_int))
    variable_lookup_p = variable_lookup.ctypes.data_as(ct.POINTER(ct.c_int))
    errors_p = errors.ctypes.data_as(ct.POINTER(ct.c_int))
    c_message_pass = ct.CDLL(lib_path or base_directory + 'message_passing.so')          # :162, re-opened every frame
    c_message_pass.message_passing(binary_sequence_p, ct.c_int(max_its), variable_lookup_p, check_lookup_p, errors_p,
                                   ct.c_int(n), ct.c_int(k), ct.c_int(dv), ct.c_int(dc))
    errors = np.insert(errors, 0, initial_error_count)
    return binary_sequence, errors


def run_fixed_loop(n, dv, dc, erasure_prob, iterations, check_lookup, variable_lookup, seconds=5.0, max_frames=10 ** 9,
                   lib_path=None):
    """The loop of :367-392 for `seconds` of wall clock (no 200-block-error stop: a throughput sample).
    lib_path: the shared object the loop opens instead of the reference's message_passing.so (the
    one-line swap to libldpc_b200.so) -- the same loop then measures the drop-in symbol.
    Returns (frames, elapsed seconds, block errors)."""
    k = int(n * (dc - dv) / dc)
    codeword = np.zeros(n)
    message_passing_error_counts = np.zeros(iterations + 1)
    block_errors = bit_errors = 0
    sim_BEC = BEC(erasure_prob)
    i = 0
    t0 = time.perf_counter()
    while i < max_frames and time.perf_counter() - t0 < seconds:
        channel_output = sim_BEC.new_transmit(codeword)
        decoded, errors = message_pass_decode(n, k, dv, dc, channel_output, iterations, check_lookup, variable_lookup, lib_path)
        message_passing_error_counts += errors
        if errors[-1] != 0:
            block_errors += 1
        bit_errors += errors[-1]
        i += 1
    return i, time.perf_counter() - t0, block_errors
