"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/ldpc_b200.h
declares, fails loudly without a GPU, and the host-side graph generators honour the reference contract."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(message_passing|generate_random_code|ldpc_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from iib_project_ldpc_codes_b200 import _lib
    lib = ct.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert "message_passing" in syms and "ldpc_b200_mc_run" in syms and len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ldpc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_no_gpu_fails_loudly_not_silently(built):
    import iib_project_ldpc_codes_b200 as L
    if L._lib.load().ldpc_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    vl, cl = L.generate_graph(24, 3, 6, seed=1)
    with pytest.raises(L.LdpcB200Error):
        L.Graph(vl, cl)
    with pytest.raises(L.LdpcB200Error):
        L.message_passing(np.zeros(24, np.int32), 3, vl.ravel().copy(), cl.ravel().copy(), np.zeros(3, np.int32),
                          24, 12, 3, 6)


def _check_contract(vl, cl, n, dv, dc):
    m = n * dv // dc
    assert vl.shape == (n, dv) and cl.shape == (m, dc)
    assert ((cl >= 0) & (cl < n)).all() and ((vl >= 0) & (vl < m)).all()
    # no repeated variable inside a check (random_code_generator.c:39-47)
    s = np.sort(cl, axis=1)
    assert (s[:, 1:] != s[:, :-1]).all()
    # every variable has degree dv; variable_lookup lists its checks in ascending order (:53-64)
    assert (np.bincount(cl.ravel(), minlength=n) == dv).all()
    assert (np.diff(vl, axis=1) > 0).all()
    want = [[] for _ in range(n)]
    for c in range(m):
        for v in cl[c]:
            want[v].append(c)
    assert np.array_equal(vl, np.array([sorted(w) for w in want], dtype=np.int32))


@pytest.mark.parametrize("n,dv,dc,mode", [(24, 3, 6, 0), (1000, 3, 6, 0), (1000, 3, 6, 1), (2000, 4, 8, 1), (60, 3, 4, 1),
                                          (100000, 3, 6, 1)])
def test_seeded_generator_contract(built, n, dv, dc, mode):
    import iib_project_ldpc_codes_b200 as L
    vl, cl = L.generate_graph(n, dv, dc, seed=3, mode=mode)
    _check_contract(vl, cl, n, dv, dc)
    vl2, cl2 = L.generate_graph(n, dv, dc, seed=3, mode=mode)
    assert np.array_equal(vl, vl2) and np.array_equal(cl, cl2)
    vl3, cl3 = L.generate_graph(n, dv, dc, seed=4, mode=mode)
    assert not np.array_equal(cl, cl3)


def _call_generate_random_code(lib, n, dv, dc, dense=True):
    k = n * (dc - dv) // dc
    seq = np.arange(n * dv, dtype=np.int32)
    cl = np.zeros(n * dv, dtype=np.int32)
    vl = np.zeros(n * dv, dtype=np.int32)
    H = np.zeros(n * (n - k), dtype=np.bool_) if dense else None
    P = ct.POINTER(ct.c_int)
    ok = 0
    while ok == 0:   # parallel_simulator_expurgated.py:222-223
        ok = lib.generate_random_code(ct.c_int(n), ct.c_int(dv), ct.c_int(dc), vl.ctypes.data_as(P), cl.ctypes.data_as(P),
                                      seq.ctypes.data_as(P),
                                      H.ctypes.data_as(ct.POINTER(ct.c_bool)) if dense else None,
                                      ct.c_int(0), ct.c_bool(False), ct.c_int(0))
    return vl.reshape(n, dv), cl.reshape(n - k, dc), (H.reshape(n - k, n) if dense else None), seq


def test_generate_random_code_dropin_contract(built):
    from iib_project_ldpc_codes_b200 import _lib
    lib = ct.CDLL(_lib.LIB_PATH)
    vl, cl, H, seq = _call_generate_random_code(lib, 120, 3, 6)
    _check_contract(vl, cl, 120, 3, 6)
    assert sorted(seq.tolist()) == list(range(360))               # a permutation of the sockets, kept in place
    assert np.array_equal(cl.ravel(), seq // 3)                   # random_code_generator.c:34-36
    want_H = np.zeros_like(H)
    for c in range(60):
        want_H[c, cl[c]] = True
    assert np.array_equal(H, want_H)
    vl2, cl2, _, _ = _call_generate_random_code(lib, 120, 3, 6, dense=False)   # NULL parity_check allowed
    _check_contract(vl2, cl2, 120, 3, 6)


def test_generate_random_code_identical_to_reference_given_same_libc_rand_state(built, oracle):
    """Same libc rand() consumption as random_code_generator.c:13-19,26-47 -> identical graphs."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    from iib_project_ldpc_codes_b200 import _lib
    mine = ct.CDLL(_lib.LIB_PATH)
    libc = ct.CDLL(None)
    for (n, dv, dc, s) in [(24, 3, 6, 1), (1000, 3, 6, 7), (96, 3, 6, 3), (60, 3, 4, 9)]:
        vr, cr, Hr = oracle.ref_generate_random_code(n, dv, dc, srand_seed=s)
        libc.srand(ct.c_uint(s))
        vm, cm, Hm, _ = _call_generate_random_code(mine, n, dv, dc)
        assert np.array_equal(vr, vm) and np.array_equal(cr, cm) and np.array_equal(Hr, Hm)


def test_modified_density_evolution_matches_reference_values():
    """tools/density_evolution.py:18-28 at eps=0.40, (3,6): 0.4, 0.313756, 0.267938, 0.236467, 0.211673."""
    from iib_project_ldpc_codes_b200.simulator import modified_density_evolution
    got = modified_density_evolution(0.40, 4, 3, 6)
    assert np.allclose(got, [0.4, 0.313756, 0.267938, 0.236467, 0.211673], atol=1e-6)


@pytest.mark.parametrize("n,dv,dc", [(24, 3, 6), (1000, 3, 6), (2000, 4, 8), (60, 3, 4)])
def test_indexed_generator_contract(built, n, dv, dc):
    """Host twin of the in-kernel ensemble generator: simple regular graphs, reproducible, index-dependent."""
    import iib_project_ldpc_codes_b200 as L
    vl, cl = L.generate_graph_indexed(n, dv, dc, 12345, 7)
    _check_contract(vl, cl, n, dv, dc)
    vl2, cl2 = L.generate_graph_indexed(n, dv, dc, 12345, 7)
    assert np.array_equal(cl, cl2)
    _, cl3 = L.generate_graph_indexed(n, dv, dc, 12345, 8)
    assert not np.array_equal(cl, cl3)
