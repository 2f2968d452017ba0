"""Thin Python layer over the C ABI (include/ldpc_b200.h): graph handles, fused Monte-Carlo, batched decode.

numpy owns host buffers; PyTorch is used only by callers that want device-resident counters / streams
(bench.py, dist.py).  Nothing here computes on the CPU: every call lands in libldpc_b200.so.
"""
from __future__ import annotations

import ctypes as ct
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import P_F32, P_F64, P_I8, P_I32, P_U64, McStats, check, ptr


def counters_len(iterations: int) -> int:
    return iterations + 6


@dataclass
class McResult:
    """Counter block of one erasure point, named the way the reference drivers name them
    (parallel_simulator_expurgated.py:194-282)."""
    iterations: int
    n: int
    raw: np.ndarray  # uint64 [iterations + 6]

    @property
    def error_counts(self):  # message_passing_error_counts, length iterations+1
        return self.raw[: self.iterations + 1]

    @property
    def block_errors(self):
        return int(self.raw[self.iterations + 1])

    @property
    def bit_errors(self):
        return int(self.raw[self.iterations + 2])

    @property
    def frames(self):  # num_tests
        return int(self.raw[self.iterations + 3])

    @property
    def counted_frames(self):
        return int(self.raw[self.iterations + 4])

    @property
    def executed_iterations(self):
        return int(self.raw[self.iterations + 5])

    # the three outputs the reference writes to CSV (parallel_simulator_expurgated.py:260,278-282)
    @property
    def average_errors(self):
        return self.error_counts.astype(np.float64) / (self.n * max(self.frames, 1))

    @property
    def block_error_rate(self):
        return self.block_errors / max(self.frames, 1)

    @property
    def bit_error_rate(self):
        return self.bit_errors / (max(self.frames, 1) * self.n)


def generate_graph(n: int, dv: int, dc: int, seed: int = 1, mode: int = 1):
    """Seeded simple (dv,dc)-regular graph -> (variable_lookup [n][dv], check_lookup [m][dc]) int32
    with the contract of random_code_generator.c:34-36,53-64.  mode 0 = rejection, 1 = repair."""
    lib = _lib.load()
    if (n * dv) % dc:
        raise ValueError("n*dv must be divisible by dc")
    vl = np.zeros(n * dv, dtype=np.int32)
    cl = np.zeros(n * dv, dtype=np.int32)
    ok = lib.ldpc_b200_generate_graph(n, dv, dc, seed, mode, ptr(vl, P_I32), ptr(cl, P_I32))
    if ok != 1:
        raise RuntimeError(f"graph generation failed for n={n} dv={dv} dc={dc} mode={mode}")
    return vl.reshape(n, dv), cl.reshape(n * dv // dc, dc)


def generate_graph_indexed(n: int, dv: int, dc: int, graph_seed: int, index: int):
    """Host twin of the in-kernel ensemble generator (gen_mode 2): graph `index` for `graph_seed`."""
    lib = _lib.load()
    vl = np.zeros(n * dv, dtype=np.int32)
    cl = np.zeros(n * dv, dtype=np.int32)
    if lib.ldpc_b200_generate_graph_indexed(n, dv, dc, graph_seed, index, ptr(vl, P_I32), ptr(cl, P_I32)) != 1:
        raise RuntimeError(f"indexed graph generation failed for n={n} dv={dv} dc={dc}")
    return vl.reshape(n, dv), cl.reshape(n * dv // dc, dc)


class Graph:
    """Device-resident Tanner graph (variable_to_check_list [n][dv], check_to_variable_list [m][dc])."""

    def __init__(self, variable_lookup, check_lookup, n=None, dv=None, dc=None, device: int = 0):
        lib = _lib.load()
        vl = np.ascontiguousarray(variable_lookup, dtype=np.int32)
        cl = np.ascontiguousarray(check_lookup, dtype=np.int32)
        if n is None:
            n, dv = vl.shape
            dc = cl.shape[1]
        self.n, self.dv, self.dc = int(n), int(dv), int(dc)
        self.m = self.n * self.dv // self.dc
        self.k = self.n - self.m
        self.device = device
        self.variable_lookup = vl.reshape(self.n, self.dv)
        self.check_lookup = cl.reshape(self.m, self.dc)
        h = ct.c_void_p()
        check(lib.ldpc_b200_graph_create(ptr(vl, P_I32), ptr(cl, P_I32), self.n, self.k, self.dv, self.dc, device,
                                         ct.byref(h)))
        self._h = h
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ldpc_b200_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tuning(self, words=0, teams=0, team_warps=0):
        check(self._lib.ldpc_b200_set_tuning(self._h, words, teams, team_warps))

    def set_engine(self, engine):
        """0 = automatic, 1 = shared-memory engine (E1), 2 = HBM/L2-streamed engine (E2)."""
        check(self._lib.ldpc_b200_set_engine(self._h, engine))

    def set_place_effort(self, effort):
        check(self._lib.ldpc_b200_set_place_effort(self._h, effort))

    def wait_placement(self, words=0):
        """Blocks until the background placement searches of this handle are finished and uploaded (launches made
        before that decode on the natural layout: same results, slower kernels).  Returns the source of the
        placement for `words` words per node: 0 none, 1 searched in this process, 2 read from the disk cache."""
        src = ct.c_int(0)
        check(self._lib.ldpc_b200_placement_wait(self._h, int(words), ct.byref(src)))
        return src.value

    def placement_info(self, words):
        a, b, c = ct.c_double(), ct.c_double(), ct.c_double()
        check(self._lib.ldpc_b200_placement_info(self._h, words, ct.byref(a), ct.byref(b), ct.byref(c)))
        return {"gather_wavefronts_natural": a.value, "gather_wavefronts_placed": b.value,
                "gather_wavefronts_ideal": c.value}

    # ---- fused Monte-Carlo ----------------------------------------------------------------------
    def mc_run(self, eps, seed, num_frames, iterations, expurgation=-1, first_frame=0, return_stats=False):
        """Host-buffer call: returns uint64 [npoints][iterations+6] (and the launch statistics)."""
        eps = np.ascontiguousarray(np.atleast_1d(eps), dtype=np.float64)
        out = np.zeros((len(eps), counters_len(iterations)), dtype=np.uint64)
        st = McStats()
        check(self._lib.ldpc_b200_mc_run(self._h, ptr(eps, P_F64), len(eps), seed, first_frame, num_frames,
                                         iterations, expurgation, ptr(out, P_U64), ct.byref(st)))
        return (out, st.as_dict()) if return_stats else out

    def mc_launch(self, eps, seed, num_frames, iterations, d_counters_ptr, stream_ptr=0, expurgation=-1,
                  first_frame=0):
        """Asynchronous launch ADDING into a device buffer [npoints][iterations+6] of uint64/int64."""
        eps = np.ascontiguousarray(np.atleast_1d(eps), dtype=np.float64)
        check(self._lib.ldpc_b200_mc_launch(self._h, ptr(eps, P_F64), len(eps), seed, first_frame, num_frames,
                                            iterations, expurgation, ct.c_void_p(d_counters_ptr),
                                            ct.c_void_p(stream_ptr)))

    def mc_launch_point(self, eps_p, point, seed, num_frames, iterations, d_counters_ptr, stream_ptr=0, expurgation=-1,
                        first_frame=0):
        """One erasure point of a longer list, keyed as point index `point` (ldpc_b200_mc_launch_points): ADDS into the
        device row [iterations+6] at d_counters_ptr."""
        eps = np.ascontiguousarray([eps_p], dtype=np.float64)
        check(self._lib.ldpc_b200_mc_launch_points(self._h, ptr(eps, P_F64), 1, point, seed, first_frame, num_frames,
                                                   iterations, expurgation, ct.c_void_p(d_counters_ptr),
                                                   ct.c_void_p(stream_ptr)))

    def mc_stats(self):
        st = McStats()
        check(self._lib.ldpc_b200_mc_stats_read(self._h, ct.byref(st)))
        return st.as_dict()

    def kernel_times_ms(self, max_n=256):
        buf = np.zeros(max_n, dtype=np.float32)
        k = check(self._lib.ldpc_b200_kernel_times(self._h, ptr(buf, P_F32), max_n))
        return buf[:k].copy()

    def results(self, counters, iterations):
        return [McResult(iterations, self.n, np.asarray(row)) for row in np.asarray(counters)]

    # ---- batched decode ---------------------------------------------------------------------------
    def decode_batch(self, Mvc, iterations):
        """Mvc [frames][n] in {0,1,2} -> (decoded int8 [frames][n], errors int32 [frames][iterations]
        as message_passing.c leaves a zeroed errors[], ret int32 [frames])."""
        M = np.array(Mvc, dtype=np.int8, order="C").reshape(-1, self.n)
        frames = M.shape[0]
        errors = np.zeros((frames, max(iterations, 0)), dtype=np.int32)
        ret = np.zeros(frames, dtype=np.int32)
        check(self._lib.ldpc_b200_decode_batch(self._h, ptr(M, P_I8), frames, iterations, ptr(errors, P_I32),
                                               ptr(ret, P_I32)))
        return M, errors, ret


def mc_run_ensemble(n, dv, dc, eps, seed, num_graphs, iterations, frames_per_graph=128, graph_seed=1, gen_mode=1,
                    first_graph=0, expurgation=-1, device=0, return_stats=False, frame_stride=0):
    """Ensemble Monte-Carlo: graph G = generate_graph(n, dv, dc, graph_seed*1000003 + G, gen_mode) decodes the
    global frames [G*F, G*F + frames_per_graph), F = frame_stride (0: frames_per_graph rounded up to 128)."""
    lib = _lib.load()
    eps = np.ascontiguousarray(np.atleast_1d(eps), dtype=np.float64)
    out = np.zeros((len(eps), counters_len(iterations)), dtype=np.uint64)
    st = McStats()
    check(lib.ldpc_b200_mc_run_ensemble(device, n, dv, dc, graph_seed, gen_mode, first_graph, num_graphs,
                                        frames_per_graph, frame_stride, ptr(eps, P_F64), len(eps), seed, iterations, expurgation,
                                        ptr(out, P_U64), ct.byref(st)))
    return (out, st.as_dict()) if return_stats else out


def draw_patterns(n, eps, seed, point, first_frame, frames, device=0):
    """Erasure patterns of the in-kernel generator (computed on the GPU): int8 [frames][n] in {0,2}."""
    lib = _lib.load()
    out = np.zeros((frames, n), dtype=np.int8)
    check(lib.ldpc_b200_draw_patterns(device, n, float(eps), seed, point, first_frame, frames, ptr(out, P_I8)))
    return out


def message_passing(Mvc, iterations, variable_to_check_list, check_to_variable_list, errors, n, k, dv, dc):
    """The drop-in symbol, called the way parallel_simulator.py:152-164 calls the reference:
    int32 arrays mutated in place, returns the C return value."""
    lib = _lib.load()
    for a in (Mvc, variable_to_check_list, check_to_variable_list, errors):
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return check(lib.message_passing(ptr(Mvc, P_I32), iterations, ptr(variable_to_check_list, P_I32),
                                     ptr(check_to_variable_list, P_I32), ptr(errors, P_I32), n, k, dv, dc))


def smem_bandwidth_gbs(device=0):
    """Measured whole-GPU shared-memory read bandwidth (GB/s), the E1 roofline denominator."""
    lib = _lib.load()
    v = ct.c_double(0.0)
    check(lib.ldpc_b200_smem_probe(device, ct.byref(v)))
    return v.value
