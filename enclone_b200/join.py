"""Host-side mirror of the reference interface for the join path.

`join_exacts(pack, params)` plays the role of enclone::join::join_exacts (+ finish_join):
it returns the EquivRel over `info` and fills raw_joins / per-join records.
The flat `JoinPack` stands for (exact_clonotypes, info, refdata, dref); `params` for the join
knobs of EncloneControl.  All scoring runs in libenclone_join.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._abi import (EDGE_DTYPE, RAW_EDGE_DTYPE, EjoinEdge, EjoinParams, EjoinRawEdge, EjoinResult, EjoinStats,
                   default_params)


class EquivRel:
    """The part of equiv::EquivRel the callers use (enclone_tail/src/grouper.rs:406-445):
    class_id, orbit_reps, orbit — built from the partition labels the library returns."""

    def __init__(self, class_of):
        self._cls = class_of                     # the library's labels (min member of each class); nothing is copied or
        self._orbits = None                      # converted until a caller asks for orbits

    def class_id(self, a):
        return int(self._cls[a])

    def _build(self):
        if self._orbits is None:
            self._cls = np.asarray(self._cls, dtype=np.int64)
            order = np.argsort(self._cls, kind="stable")
            reps, starts = np.unique(self._cls[order], return_index=True)
            self._reps = reps
            self._orbits = np.split(order, starts[1:])

    def orbit_reps(self):
        self._build()
        return [int(x) for x in self._reps]

    def orbit(self, rep):
        self._build()
        i = int(np.searchsorted(self._reps, rep))
        return [int(x) for x in self._orbits[i]]

    def norbits(self):
        self._build()
        return len(self._reps)

    def labels(self):
        return self._cls


class JoinResult:
    def __init__(self, edges, class_of, stats):
        self.edges = edges                      # structured array, raw_joins order
        self.eq = EquivRel(class_of)
        self.class_of = class_of
        self.stats = stats

    @property
    def raw_joins(self):
        """Vec<(i32,i32)> of the reference: the emitted (k1,k2) in order."""
        return [(int(a), int(b)) for a, b in zip(self.edges["k1"], self.edges["k2"])]


def _edges_np(ptr, n, dtype, ctype, copy=True):
    if n == 0:
        return np.zeros(0, dtype)
    addr = C.cast(ptr, C.c_void_p).value
    view = np.frombuffer((ctype * n).from_address(addr), dtype=dtype)      # zero-copy view of library memory
    return view.copy() if copy else view


class Joiner:
    """One GPU's join engine (wraps an ejoin_handle_t)."""

    def __init__(self, params=None, device=0):
        self.L = _lib.load()
        self.params = params or default_params()
        self.h = C.c_void_p()
        rc = self.L.ejoin_create(C.byref(self.h), C.byref(self.params), device)
        if rc != 0:
            msg = self.L.ejoin_last_error(self.h).decode() if self.h else "no usable CUDA device"
            if self.h:
                self.L.ejoin_destroy(self.h)
                self.h = None
            raise _lib.EjoinError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.ejoin_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _lib.EjoinError(rc, self.L.ejoin_last_error(self.h).decode())

    def _result(self, res, copy=True):
        """copy=False returns views into the handle's pinned result buffer: valid until the next
        run/finish on this Joiner (one live result per handle, include/enclone_join.h)."""
        edges = _edges_np(res.edges, res.n_edges, EDGE_DTYPE, EjoinEdge, copy)
        if res.n_info and res.class_of:
            cls = np.ctypeslib.as_array(res.class_of, shape=(res.n_info,))
            cls = cls.copy() if copy else cls
        else:
            cls = np.zeros(0, np.uint32)
        stats = res.stats.as_dict()
        self.L.ejoin_result_free(C.byref(res))
        return JoinResult(edges, cls, stats)

    def run(self, pack_struct, copy=True):
        res = EjoinResult()
        self._check(self.L.ejoin_run(self.h, C.byref(pack_struct), C.byref(res)))
        return self._result(res, copy)

    def upload(self, pack_struct):
        self._check(self.L.ejoin_upload(self.h, C.byref(pack_struct)))

    def upload_shard(self, pack_struct, rank, world):
        self._check(self.L.ejoin_upload_shard(self.h, C.byref(pack_struct), rank, world))

    def run_resident(self):
        st = EjoinStats()
        self._check(self.L.ejoin_run_resident(self.h, C.byref(st)))
        return st.as_dict()

    def run_shard(self, pack_struct, rank, world):
        ep, n, st = C.POINTER(EjoinRawEdge)(), C.c_uint64(), EjoinStats()
        self._check(self.L.ejoin_run_shard(self.h, C.byref(pack_struct), rank, world, C.byref(ep), C.byref(n), C.byref(st)))
        raw = _edges_np(ep, n.value, RAW_EDGE_DTYPE, EjoinRawEdge)
        self.L.ejoin_raw_free(ep)
        return raw, st.as_dict()

    def run_shard_final(self, pack_struct, rank, world, copy=True):
        """One call per rank.  Returns (mode, JoinResult or None, raw edges or None, stats): mode 0 =
        contiguous bucket ranges (the result holds this rank's emitted joins), mode 1 = stride mode
        (raw accepted edges for the gather + finish)."""
        mode, res, ep, n, st = C.c_int(), EjoinResult(), C.POINTER(EjoinRawEdge)(), C.c_uint64(), EjoinStats()
        self._check(self.L.ejoin_run_shard_final(self.h, C.byref(pack_struct), rank, world, C.byref(mode), C.byref(res), C.byref(ep),
                                                 C.byref(n), C.byref(st)))
        if mode.value == 0:
            return 0, self._result(res, copy), None, st.as_dict()
        raw = _edges_np(ep, n.value, RAW_EDGE_DTYPE, EjoinRawEdge)
        self.L.ejoin_raw_free(ep)
        return 1, None, raw, st.as_dict()

    def set_stage_timing(self, on=True):
        self._check(self.L.ejoin_set_stage_timing(self.h, 1 if on else 0))

    def stage_times(self):
        """[(stage name, device ms)] of the last pipeline run on this handle (ejoin_stage_time: CUDA events)."""
        out, i = [], 0
        name, ms = C.c_char_p(), C.c_double()
        while self.L.ejoin_stage_time(self.h, i, C.byref(name), C.byref(ms)) == 0:
            out.append((name.value.decode(), ms.value))
            i += 1
        return out

    def comm_init(self, comm_id, rank, world):
        """Join the library's own communicator (ejoin_comm_init): `comm_id` is the 128-byte id rank 0 got from
        comm_id() and handed to every rank."""
        self._check(self.L.ejoin_comm_init(self.h, comm_id, rank, world))
        self.comm = (rank, world)

    def run_sharded(self, pack_struct, copy=True):
        """One call per rank (ejoin_run_sharded): rank 0 gets the whole join (raw_joins order, partition over all
        entries), the other ranks an empty result with their own stats."""
        res = EjoinResult()
        self._check(self.L.ejoin_run_sharded(self.h, C.byref(pack_struct), C.byref(res)))
        return self._result(res, copy)

    def partition(self, pack_struct, k1, k2):
        """finish_join on this handle's GPU (ejoin_partition_gpu): min-member labels over all info entries."""
        k1 = np.ascontiguousarray(k1, dtype=np.uint32)
        k2 = np.ascontiguousarray(k2, dtype=np.uint32)
        cls = np.zeros(max(int(pack_struct.n_info), 1), np.uint32)
        ncls = C.c_uint32()
        p = C.POINTER(C.c_uint32)
        self._check(self.L.ejoin_partition_gpu(self.h, C.byref(pack_struct), k1.ctypes.data_as(p), k2.ctypes.data_as(p), len(k1),
                                               cls.ctypes.data_as(p), C.byref(ncls)))
        return cls[:int(pack_struct.n_info)], int(ncls.value)

    def finish(self, pack_struct, raw_edges):
        raw = np.ascontiguousarray(raw_edges, dtype=RAW_EDGE_DTYPE)
        res = EjoinResult()
        ptr = raw.ctypes.data_as(C.POINTER(EjoinRawEdge))
        self._check(self.L.ejoin_finish(self.h, C.byref(pack_struct), ptr, len(raw), C.byref(res)))
        return self._result(res)


def join_exacts(pack, params=None, device=0):
    """join_exacts + finish_join on one GPU.  `pack` is a JoinPack / SynthPack (has .struct)."""
    j = Joiner(params, device)
    try:
        return j.run(pack.struct)
    finally:
        j.close()


class CompactPack:
    """The compact form of a pack (ejoin_pack_compact): V..J bases 2 bits per base in the layout of CloneInfo.tigsp,
    one 64-byte record per entry, pinned arenas read in place.  `.struct` is the C ejoin_pack_t; the source pack
    must stay alive (unchanged arrays are shared)."""

    def __init__(self, src, threads=None):
        import os
        self.L = _lib.load()
        self._src = src                                # keeps the shared arrays alive
        st = src.struct if hasattr(src, "struct") else src
        pp = C.POINTER(_lib.EjoinPack)()
        rc = self.L.ejoin_pack_compact(C.byref(st), threads or (os.cpu_count() or 1), C.byref(pp))
        if rc != 0:
            raise _lib.EjoinError(rc, "ejoin_pack_compact")
        self._pp = pp
        self.struct = pp.contents

    @property
    def n_info(self):
        return int(self.struct.n_info)

    def close(self):
        if getattr(self, "_pp", None):
            self.L.ejoin_pack_compact_free(self._pp)
            self._pp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def format_join_log(pack_struct, params, edge, **names):
    """JoinInfo.log of one emitted join (ejoin_format_join_log): `edge` is an EjoinEdge or one row of JoinResult.edges;
    names: index, id1, id2, vsegs1, vsegs2, dataset1, dataset2, cdr3_aa1, cdr3_aa2, bcs1, bcs2."""
    from ._abi import EjoinLogNames
    L = _lib.load()
    if not isinstance(edge, EjoinEdge):
        e = EjoinEdge.from_buffer_copy(np.asarray(edge).tobytes())
    else:
        e = edge
    nm = EjoinLogNames()
    nm.index = int(names.get("index", 0))
    for k in ("id1", "id2", "vsegs1", "vsegs2", "dataset1", "dataset2", "cdr3_aa1", "cdr3_aa2", "bcs1", "bcs2"):
        setattr(nm, k, str(names.get(k, "")).encode())
    n = C.c_size_t()
    rc = L.ejoin_format_join_log(C.byref(pack_struct), C.byref(params), C.byref(e), C.byref(nm), None, 0, C.byref(n))
    if rc != 0:
        raise _lib.EjoinError(rc, "ejoin_format_join_log")
    buf = C.create_string_buffer(n.value + 1)
    L.ejoin_format_join_log(C.byref(pack_struct), C.byref(params), C.byref(e), C.byref(nm), buf, n.value + 1, C.byref(n))
    return buf.value.decode("utf-8")


def comm_id():
    """A fresh communicator id (rank 0 calls this; ejoin_comm_id)."""
    buf = C.create_string_buffer(128)
    rc = _lib.load().ejoin_comm_id(buf)
    if rc != 0:
        raise _lib.EjoinError(rc, "ejoin_comm_id: libnccl.so.2 not usable")
    return buf.raw


def plan_shards(pack_struct, world):
    L = _lib.load()
    nb = C.c_uint32()
    L.ejoin_plan_shards(C.byref(pack_struct), world, None, None, 0, C.byref(nb))
    owner = np.zeros(nb.value + 1, np.uint32)
    bstart = np.zeros(nb.value + 2, np.uint32)
    rc = L.ejoin_plan_shards(C.byref(pack_struct), world, owner.ctypes.data_as(C.POINTER(C.c_uint32)),
                             bstart.ctypes.data_as(C.POINTER(C.c_uint32)), nb.value, C.byref(nb))
    if rc != 0:
        raise _lib.EjoinError(rc, "ejoin_plan_shards")
    return owner[:nb.value], bstart[:nb.value + 1]


def p1(k, d, n):
    return _lib.load().ejoin_p1(k, d, n)
