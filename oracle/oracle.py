"""oracle/oracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

ctypes front-end to
  * liboracle.so          the plain-C restatement (oracle/mp_oracle.c) of /root/reference/message_passing.c:7-82
  * _ref/*.so             the UNMODIFIED reference C compiled where it lies (oracle/Makefile), when present

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_P_I32 = ct.POINTER(ct.c_int)
_P_I8 = ct.POINTER(ct.c_int8)
_P_U64 = ct.POINTER(ct.c_uint64)
_P_F64 = ct.POINTER(ct.c_double)


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and, if /root/reference is present, oracle/_ref (building the checker is not using it)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def ncounters(iterations: int) -> int:
    return iterations + 6


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, typ):
    return a.ctypes.data_as(typ)


class _Lib:
    """Common wrapper over {orc,ref}_{mc_run,decode_batch} which share signatures."""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        self._mc = getattr(lib, prefix + "_mc_run")
        self._mc.restype = ct.c_int
        self._mc.argtypes = [_P_I32, _P_I32, ct.c_int, ct.c_int, ct.c_int, ct.c_int, _P_F64, ct.c_int,
                             ct.c_uint64, ct.c_uint64, ct.c_uint64, ct.c_int, ct.c_int, ct.c_int, _P_U64, _P_F64]
        self._db = getattr(lib, prefix + "_decode_batch")
        self._db.restype = ct.c_int
        self._db.argtypes = [_P_I32, _P_I32, ct.c_int, ct.c_int, ct.c_int, ct.c_int, _P_I8, ct.c_long, ct.c_int,
                             _P_I32, _P_I32]

    def mc_run(self, v2c, c2v, n, k, dv, dc, eps, seed, first_frame, num_frames, iterations, expurgation=-1,
               nthreads=1, return_seconds=False):
        v2c, c2v = _i32(v2c).ravel(), _i32(c2v).ravel()
        eps = np.ascontiguousarray(np.atleast_1d(eps), dtype=np.float64)
        counters = np.zeros((len(eps), ncounters(iterations)), dtype=np.uint64)
        secs = ct.c_double(0.0)
        rc = self._mc(_p(v2c, _P_I32), _p(c2v, _P_I32), n, k, dv, dc, _p(eps, _P_F64), len(eps), seed, first_frame,
                      num_frames, iterations, expurgation, nthreads, _p(counters, _P_U64), ct.byref(secs))
        assert rc == 0
        return (counters, secs.value) if return_seconds else counters

    def decode_batch(self, v2c, c2v, n, k, dv, dc, Mvc, iterations):
        """Mvc [frames][n] values 0/1/2 -> (Mvc_out int8, errors int32 [frames][iterations], ret int32 [frames])."""
        v2c, c2v = _i32(v2c).ravel(), _i32(c2v).ravel()
        M = np.array(Mvc, dtype=np.int8, order="C").reshape(-1, n) if n else np.zeros((len(Mvc), 0), np.int8)
        frames = M.shape[0]
        errors = np.zeros((frames, iterations), dtype=np.int32)
        ret = np.zeros(frames, dtype=np.int32)
        rc = self._db(_p(v2c, _P_I32), _p(c2v, _P_I32), n, k, dv, dc, _p(M, _P_I8), frames, iterations,
                      _p(errors, _P_I32), _p(ret, _P_I32))
        assert rc == 0
        return M, errors, ret


_cache = {}


def restatement() -> _Lib:
    """The plain-C restatement (always available after build())."""
    if "orc" not in _cache:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = ct.CDLL(path)
        lib.orc_message_passing.restype = ct.c_int
        lib.orc_message_passing.argtypes = [_P_I32, ct.c_int, _P_I32, _P_I32, _P_I32, ct.c_int, ct.c_int, ct.c_int,
                                            ct.c_int]
        lib.orc_draw_patterns.restype = None
        lib.orc_draw_patterns.argtypes = [_P_I8, ct.c_int, ct.c_double, ct.c_uint64, ct.c_uint32, ct.c_uint64,
                                          ct.c_long]
        lib.orc_bec_word_export.restype = ct.c_uint32
        lib.orc_bec_word_export.argtypes = [ct.c_uint64, ct.c_uint32, ct.c_uint64, ct.c_uint32, ct.c_double]
        lib.orc_executed.restype = ct.c_int
        lib.orc_executed.argtypes = [ct.c_int, _P_I32, ct.c_int]
        _cache["orc"] = _Lib(lib, "orc")
    return _cache["orc"]


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_harness.so"))


def reference() -> _Lib:
    """Frame loops around the unmodified reference object (oracle/_ref); raises if it was never built."""
    if "ref" not in _cache:
        path = os.path.join(HERE, "_ref", "libref_harness.so")
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref not built (needs /root/reference; run `make -C oracle ref`)")
        _cache["ref"] = _Lib(ct.CDLL(path), "ref")
    return _cache["ref"]


def reference_o3() -> _Lib:
    """The same unmodified reference source compiled -O3 -march=x86-64-v3 (SURVEY 8d's second C-only row)."""
    if "ref_o3" not in _cache:
        path = os.path.join(HERE, "_ref", "libref_harness_o3.so")
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libref_harness_o3.so not built (run `make -C oracle ref`)")
        _cache["ref_o3"] = _Lib(ct.CDLL(path), "ref")
    return _cache["ref_o3"]


def have_ref_o3() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_harness_o3.so"))


def generate_graph(n, dv, dc, seed=1):
    """Oracle-side twin of the product's seeded generator (mode 1): (variable_lookup [n][dv], check_lookup [m][dc])."""
    lib = restatement().lib
    lib.orc_generate_graph.restype = ct.c_int
    lib.orc_generate_graph.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_uint64, _P_I32, _P_I32]
    vl = np.zeros(n * dv, dtype=np.int32)
    cl = np.zeros(n * dv, dtype=np.int32)
    if lib.orc_generate_graph(n, dv, dc, seed, _p(vl, _P_I32), _p(cl, _P_I32)) != 1:
        raise RuntimeError("oracle graph twin failed")
    return vl.reshape(n, dv), cl.reshape(n * dv // dc, dc)


def best() -> _Lib:
    """The strongest checker available: compiled reference if present, else the restatement."""
    return reference() if have_ref() else restatement()


def single_frame(fn_lib: str, Mvc, iterations, v2c, c2v, n, k, dv, dc, errors=None):
    """Call the single-frame symbol exactly the way parallel_simulator.py:145-164 does.

    fn_lib: 'orc' (restatement) or 'ref' (oracle/_ref/message_passing.so).  Returns (Mvc, errors, ret)."""
    M = np.array(Mvc, dtype=np.int32)
    v2c, c2v = _i32(v2c).ravel(), _i32(c2v).ravel()
    err = np.zeros(iterations, dtype=np.int32) if errors is None else np.array(errors, dtype=np.int32)
    if fn_lib == "orc":
        fn = restatement().lib.orc_message_passing
    else:
        lib = ct.CDLL(os.path.join(HERE, "_ref", "message_passing.so"))
        fn = lib.message_passing
        fn.restype = ct.c_int
        fn.argtypes = [_P_I32, ct.c_int, _P_I32, _P_I32, _P_I32, ct.c_int, ct.c_int, ct.c_int, ct.c_int]
    ret = fn(_p(M, _P_I32), iterations, _p(v2c, _P_I32), _p(c2v, _P_I32), _p(err, _P_I32), n, k, dv, dc)
    return M, err, int(ret)


def draw_patterns(n, eps, seed, point, first_frame, frames):
    """Host restatement of the device erasure-pattern generator: int8 [frames][n] in {0,2}."""
    out = np.zeros((frames, n), dtype=np.int8)
    restatement().lib.orc_draw_patterns(_p(out, _P_I8), n, float(eps), seed, point, first_frame, frames)
    return out


def bec_word(seed, point, q, var, eps):
    return int(restatement().lib.orc_bec_word_export(seed, point, q, var, float(eps)))


def executed_iterations(ret, errors, iterations):
    e = _i32(errors)
    return int(restatement().lib.orc_executed(int(ret), _p(e, _P_I32), int(iterations)))


def ref_generate_random_code(n, dv, dc, srand_seed=None):
    """Graph from the UNMODIFIED reference generator (random_code_generator.c:21), called the way
    parallel_simulator_expurgated.py:203-223 does.  first_run=False so libc rand() keeps its state;
    srand_seed (optional) reseeds libc first for reproducibility.  Returns (variable_lookup[n][dv], check_lookup[m][dc])."""
    lib = ct.CDLL(os.path.join(HERE, "_ref", "random_code_generator.so"))
    libc = ct.CDLL(None)
    if srand_seed is not None:
        libc.srand(ct.c_uint(srand_seed))
    k = n * (dc - dv) // dc
    seq = np.arange(n * dv, dtype=np.int32)
    cl = np.zeros(n * dv, dtype=np.int32)
    vl = np.zeros(n * dv, dtype=np.int32)
    H = np.zeros(n * (n - k), dtype=np.bool_)
    ok = 0
    while ok == 0:
        ok = lib.generate_random_code(ct.c_int(n), ct.c_int(dv), ct.c_int(dc), _p(vl, _P_I32), _p(cl, _P_I32),
                                      _p(seq, _P_I32), H.ctypes.data_as(ct.POINTER(ct.c_bool)), ct.c_int(0),
                                      ct.c_bool(False), ct.c_int(0))
    return vl.reshape(n, dv), cl.reshape(n - k, dc), H.reshape(n - k, n)
