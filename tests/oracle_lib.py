"""ctypes wrapper of oracle/libjoin_oracle.so — the CPU restatement of the reference join.

Lives under tests/ on purpose: only tests, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may touch oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from enclone_b200._abi import EDGE_DTYPE, EjoinEdge, EjoinPack, EjoinParams, EjoinResult, default_params

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libjoin_oracle.so")


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_buckets", "pairs_candidate", "pairs_evaluated", "pairs_same_cdr3_len",
                                          "pairs_tier2", "pairs_accepted")] + [("ms_total", C.c_double), ("threads", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")])
        L = C.CDLL(_SO)
        L.oracle_p1.restype = C.c_double
        L.oracle_p1.argtypes = [C.c_uint32] * 3
        L.oracle_join_one.restype = C.c_int
        L.oracle_join_one.argtypes = [C.POINTER(EjoinPack), C.POINTER(EjoinParams), C.c_uint32, C.c_uint32,
                                      C.POINTER(EjoinEdge), C.c_void_p]
        L.oracle_fwr1_cdr12_counts.restype = C.c_int
        L.oracle_fwr1_cdr12_counts.argtypes = [C.POINTER(EjoinPack), C.c_uint32, C.c_uint32, C.POINTER(C.c_int * 4)]
        L.oracle_join_exacts.restype = C.c_int
        L.oracle_join_exacts.argtypes = [C.POINTER(EjoinPack), C.POINTER(EjoinParams), C.c_int, C.c_uint32, C.c_uint32,
                                         C.POINTER(EjoinResult), C.POINTER(OracleStats)]
        L.oracle_join_exacts_strided.restype = C.c_int
        L.oracle_join_exacts_strided.argtypes = [C.POINTER(EjoinPack), C.POINTER(EjoinParams), C.c_int, C.c_uint32, C.c_uint32,
                                                 C.c_uint32, C.c_uint32, C.POINTER(EjoinResult), C.POINTER(OracleStats)]
        L.oracle_accept_graph.restype = C.c_int
        L.oracle_accept_graph.argtypes = [C.POINTER(EjoinPack), C.POINTER(EjoinParams), C.c_uint32, C.c_uint32,
                                          C.POINTER(C.POINTER(EjoinEdge)), C.POINTER(C.c_uint64)]
        L.oracle_result_free.argtypes = [C.POINTER(EjoinResult)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_prepare.argtypes = [C.c_int]
        L.oracle_set_parallel_bucket.argtypes = [C.c_uint32, C.c_int]
        L.oracle_set_row_limit.argtypes = [C.c_uint32]
        L.oracle_ref_positions_differing.argtypes = [C.POINTER(EjoinPack), C.c_uint32, C.c_uint32]
        L.oracle_ref_positions_differing.restype = C.c_int
        from enclone_b200.group import EgroupAsymOpts, EgroupAsymResult, EgroupOpts, EgroupPack
        L.oracle_group_asym.restype = C.c_int
        L.oracle_group_asym.argtypes = [C.POINTER(EgroupPack), C.POINTER(C.c_uint8), C.POINTER(EgroupAsymOpts), C.POINTER(EgroupAsymResult), C.POINTER(C.c_uint64)]
        L.oracle_group_asym_free.argtypes = [C.POINTER(EgroupAsymResult)]
        L.oracle_group.restype = C.c_int
        L.oracle_group.argtypes = [C.POINTER(EgroupPack), C.POINTER(EgroupOpts), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def group(pack, opts_struct):
    """oracle/grouper_oracle.c: (group_of, chain-pair evaluations)"""
    out = np.zeros(max(pack.n_clono, 1), np.uint32)
    ev = C.c_uint64(0)
    rc = lib().oracle_group(C.byref(pack.struct), C.byref(opts_struct), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(ev))
    assert rc == 0
    return out[:pack.n_clono], ev.value


def p1(k, d, n):
    return lib().oracle_p1(k, d, n)


def join_one(pack_struct, params, k1, k2):
    e = EjoinEdge()
    ok = lib().oracle_join_one(C.byref(pack_struct), C.byref(params), k1, k2, C.byref(e), None)
    return bool(ok), e


def region_counts(pack_struct, k1, k2):
    out = (C.c_int * 4)()
    ok = lib().oracle_fwr1_cdr12_counts(C.byref(pack_struct), k1, k2, C.byref(out))
    return bool(ok), list(out)


def _edges_to_numpy(ptr, n):
    if n == 0:
        return np.zeros(0, EDGE_DTYPE)
    buf = C.string_at(ptr, n * C.sizeof(EjoinEdge))
    return np.frombuffer(buf, dtype=EDGE_DTYPE).copy()


def join_exacts(pack_struct, params=None, threads=1, bucket_lo=0, bucket_hi=0xFFFFFFFF, stride=1, phase=0):
    """Returns (edges ndarray in raw_joins order, class_of ndarray, stats dict)."""
    params = params or default_params()
    L = lib()
    L.oracle_prepare(600)
    res, st = EjoinResult(), OracleStats()
    rc = L.oracle_join_exacts_strided(C.byref(pack_struct), C.byref(params), threads, bucket_lo, bucket_hi, stride, phase,
                                      C.byref(res), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"oracle_join_exacts: {rc}")
    edges = _edges_to_numpy(res.edges, res.n_edges)
    cls = np.ctypeslib.as_array(res.class_of, shape=(res.n_info,)).copy() if res.n_info else np.zeros(0, np.uint32)
    stats = {n: getattr(st, n) for n, _ in st._fields_}
    L.oracle_result_free(C.byref(res))
    return edges, cls, stats


def set_parallel_bucket(min_entries, threads):
    """Buckets with at least `min_entries` entries are scored by `threads` workers row by row and
    replayed in scan order (join_oracle.c, join_core_parallel); 0 switches the mode off."""
    lib().oracle_set_parallel_bucket(min_entries, threads)


def ref_positions_differing(pack_struct, k1, k2):
    return lib().oracle_ref_positions_differing(C.byref(pack_struct), k1, k2)


def accept_graph(pack_struct, params, i, j):
    ep, n = C.POINTER(EjoinEdge)(), C.c_uint64()
    lib().oracle_accept_graph(C.byref(pack_struct), C.byref(params), i, j, C.byref(ep), C.byref(n))
    out = _edges_to_numpy(ep, n.value)
    lib().oracle_free(ep)
    return out


def group_asym(pack, in_center, opts_struct):
    """oracle/grouper_oracle.c, Case 2: ([[(clonotype, distance), ...], ...], distance evaluations)"""
    from enclone_b200.group import EgroupAsymResult, result_groups
    cen = np.ascontiguousarray(in_center, np.uint8)
    res, ev = EgroupAsymResult(), C.c_uint64(0)
    rc = lib().oracle_group_asym(C.byref(pack.struct), cen.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(opts_struct), C.byref(res), C.byref(ev))
    assert rc == 0
    out = result_groups(res)
    lib().oracle_group_asym_free(C.byref(res))
    return out, ev.value


def refine(pack, opts_struct):
    """oracle/refine_oracle.c: (RefineResult, stats dict) of the post-join refinement"""
    from enclone_b200.refine import ErefineOpts, ErefinePack, ErefineResult, ErefineStats, RefineResult
    L = lib()
    L.oracle_refine.restype = C.c_int
    L.oracle_refine.argtypes = [C.POINTER(ErefinePack), C.POINTER(ErefineOpts), C.POINTER(ErefineResult), C.POINTER(ErefineStats)]
    res, st = RefineResult(pack.n_exact, pack.n_chain), ErefineStats()
    rc = L.oracle_refine(C.byref(pack.struct), C.byref(opts_struct), C.byref(res.struct), C.byref(st))
    assert rc == 0, rc
    return res.trim(), st.as_dict()


def refine_opts(**kw):
    from enclone_b200.refine import ErefineOpts
    o = ErefineOpts()
    lib().oracle_refine_opts_default.argtypes = [C.POINTER(ErefineOpts)]
    lib().oracle_refine_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, int(v))
    return o
