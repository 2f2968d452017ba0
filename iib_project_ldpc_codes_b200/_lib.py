"""ctypes binding of libldpc_b200.so (include/ldpc_b200.h).

The library is built in-tree by ``iib_project_ldpc_codes_b200.build`` (nvcc, sm_100a).  There is no CPU
fallback: if the shared object is missing, loading raises; if no GPU is present, every compute entry
point returns LDPC_B200_ECUDA and the wrappers raise ``LdpcB200Error``.
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(PKG_DIR, "libldpc_b200.so")

P_I32 = ct.POINTER(ct.c_int)
P_I8 = ct.POINTER(ct.c_int8)
P_U64 = ct.POINTER(ct.c_uint64)
P_F64 = ct.POINTER(ct.c_double)
P_F32 = ct.POINTER(ct.c_float)


class LdpcB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libldpc_b200 error {code}: {msg}")
        self.code = code


class McStats(ct.Structure):
    _fields_ = [("phys_frame_iterations", ct.c_uint64), ("second_pass_groups", ct.c_uint64),
                ("engine", ct.c_int32), ("words", ct.c_int32), ("teams", ct.c_int32), ("team_warps", ct.c_int32),
                ("grid", ct.c_int32), ("smem_bytes", ct.c_int32), ("stages", ct.c_int32), ("kernel_launches", ct.c_int32),
                ("spilled_records", ct.c_uint64), ("recounted_frames", ct.c_uint64), ("record_bytes", ct.c_int32),
                ("tmem_lists", ct.c_int32),
                ("overlap_producers", ct.c_int32), ("reserved_", ct.c_int32), ("overlap_cycles", ct.c_uint64 * 4)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "overlap_cycles" else int(getattr(self, k))) for k, _ in self._fields_}


# every symbol include/ldpc_b200.h declares (tests check the shared object exports all of them)
EXPORTS = [
    "message_passing", "generate_random_code", "ldpc_b200_last_error", "ldpc_b200_device_count",
    "ldpc_b200_generate_graph", "ldpc_b200_graph_create", "ldpc_b200_graph_destroy", "ldpc_b200_counters_len",
    "ldpc_b200_mc_run", "ldpc_b200_mc_launch", "ldpc_b200_mc_launch_points", "ldpc_b200_mc_stats_read", "ldpc_b200_kernel_times",
    "ldpc_b200_decode_batch", "ldpc_b200_draw_patterns", "ldpc_b200_set_tuning", "ldpc_b200_smem_probe",
    "ldpc_b200_set_place_effort", "ldpc_b200_placement_info", "ldpc_b200_placement_wait", "ldpc_b200_set_engine", "ldpc_b200_mc_run_ensemble", "ldpc_b200_generate_graph_indexed",
]

_lib = None


def load() -> ct.CDLL:
    """Load libldpc_b200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = ct.CDLL(LIB_PATH)
    lib.ldpc_b200_last_error.restype = ct.c_char_p
    lib.ldpc_b200_device_count.restype = ct.c_int
    lib.ldpc_b200_counters_len.restype = ct.c_int
    lib.ldpc_b200_counters_len.argtypes = [ct.c_int]
    lib.message_passing.restype = ct.c_int
    lib.message_passing.argtypes = [P_I32, ct.c_int, P_I32, P_I32, P_I32, ct.c_int, ct.c_int, ct.c_int, ct.c_int]
    lib.generate_random_code.restype = ct.c_int
    lib.generate_random_code.argtypes = [ct.c_int, ct.c_int, ct.c_int, P_I32, P_I32, P_I32, ct.POINTER(ct.c_bool),
                                         ct.c_int, ct.c_bool, ct.c_int]
    lib.ldpc_b200_generate_graph.restype = ct.c_int
    lib.ldpc_b200_generate_graph.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_uint64, ct.c_int, P_I32, P_I32]
    lib.ldpc_b200_generate_graph_indexed.restype = ct.c_int
    lib.ldpc_b200_generate_graph_indexed.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_uint64, ct.c_uint64, P_I32, P_I32]
    lib.ldpc_b200_graph_create.restype = ct.c_int
    lib.ldpc_b200_graph_create.argtypes = [P_I32, P_I32, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int,
                                           ct.POINTER(ct.c_void_p)]
    lib.ldpc_b200_graph_destroy.restype = None
    lib.ldpc_b200_graph_destroy.argtypes = [ct.c_void_p]
    lib.ldpc_b200_mc_run.restype = ct.c_int
    lib.ldpc_b200_mc_run.argtypes = [ct.c_void_p, P_F64, ct.c_int, ct.c_uint64, ct.c_uint64, ct.c_uint64, ct.c_int,
                                     ct.c_int, P_U64, ct.POINTER(McStats)]
    lib.ldpc_b200_mc_run_ensemble.restype = ct.c_int
    lib.ldpc_b200_mc_run_ensemble.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_uint64, ct.c_int, ct.c_uint64,
                                              ct.c_uint64, ct.c_uint32, ct.c_uint32, P_F64, ct.c_int, ct.c_uint64, ct.c_int,
                                              ct.c_int, P_U64, ct.POINTER(McStats)]
    lib.ldpc_b200_mc_launch.restype = ct.c_int
    lib.ldpc_b200_mc_launch.argtypes = [ct.c_void_p, P_F64, ct.c_int, ct.c_uint64, ct.c_uint64, ct.c_uint64, ct.c_int,
                                        ct.c_int, ct.c_void_p, ct.c_void_p]
    lib.ldpc_b200_mc_launch_points.restype = ct.c_int
    lib.ldpc_b200_mc_launch_points.argtypes = [ct.c_void_p, P_F64, ct.c_int, ct.c_uint32, ct.c_uint64, ct.c_uint64, ct.c_uint64,
                                               ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]
    lib.ldpc_b200_mc_stats_read.restype = ct.c_int
    lib.ldpc_b200_mc_stats_read.argtypes = [ct.c_void_p, ct.POINTER(McStats)]
    lib.ldpc_b200_kernel_times.restype = ct.c_int
    lib.ldpc_b200_kernel_times.argtypes = [ct.c_void_p, P_F32, ct.c_int]
    lib.ldpc_b200_decode_batch.restype = ct.c_int
    lib.ldpc_b200_decode_batch.argtypes = [ct.c_void_p, P_I8, ct.c_uint64, ct.c_int, P_I32, P_I32]
    lib.ldpc_b200_draw_patterns.restype = ct.c_int
    lib.ldpc_b200_draw_patterns.argtypes = [ct.c_int, ct.c_int, ct.c_double, ct.c_uint64, ct.c_uint32, ct.c_uint64,
                                            ct.c_uint64, P_I8]
    lib.ldpc_b200_smem_probe.restype = ct.c_int
    lib.ldpc_b200_smem_probe.argtypes = [ct.c_int, P_F64]
    lib.ldpc_b200_set_place_effort.restype = ct.c_int
    lib.ldpc_b200_set_place_effort.argtypes = [ct.c_void_p, ct.c_int]
    lib.ldpc_b200_placement_wait.restype = ct.c_int
    lib.ldpc_b200_placement_wait.argtypes = [ct.c_void_p, ct.c_int, ct.POINTER(ct.c_int)]
    lib.ldpc_b200_placement_info.restype = ct.c_int
    lib.ldpc_b200_placement_info.argtypes = [ct.c_void_p, ct.c_int, P_F64, P_F64, P_F64]
    lib.ldpc_b200_set_engine.restype = ct.c_int
    lib.ldpc_b200_set_engine.argtypes = [ct.c_void_p, ct.c_int]
    lib.ldpc_b200_set_tuning.restype = ct.c_int
    lib.ldpc_b200_set_tuning.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int]
    _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        raise LdpcB200Error(rc, load().ldpc_b200_last_error().decode())
    return rc


def ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)
