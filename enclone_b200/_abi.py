"""ctypes mirror of include/enclone_join.h (the C ABI of libenclone_join.so).

Only layouts live here; loading the CUDA library is in _lib.py so that CPU-only
tools (the synthetic generator, the tests' oracle wrapper) can share the structs.
"""
import ctypes as C

import numpy as np

EJOIN_ABI_VERSION = 2
EJOIN_NONE = 0xFFFFFFFF
EJOIN_NONE16 = 0xFFFF
EJOIN_F_TIGS_IN_PLACE = 1
EJOIN_F_TIGS_2BIT = 2
EJOIN_F_ENTRY_RECS = 4
EJOIN_F_CDR3_2BIT = 8
EJOIN_F_TIGS_PLANES = 16

u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))


class EjoinEntryRec(C.Structure):
    _fields_ = [("tig_off", C.c_uint32 * 2), ("cdr3_start", C.c_uint16 * 2)] + [
        (n, C.c_uint16 * 2) for n in ("v_ref_id", "j_ref_id", "dref_id", "vs_seq", "js_seq", "vname_id", "vuni_seq", "utr_seq")] + [
        (n, C.c_uint16) for n in ("c_ref_id_light", "hcomp", "fr1_start", "cdr1_start", "fr2_start", "cdr2_start", "fr3_start")] + [
        ("has_del", C.c_uint8 * 2), ("pad", C.c_uint8 * 4)]


assert C.sizeof(EjoinEntryRec) == 64


class EjoinPack(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("n_info", C.c_uint32), ("n_exact", C.c_uint32),
        ("n_refseq", C.c_uint32), ("is_bcr", C.c_uint32), ("flags", C.c_uint32),
        ("exact_id", u32p),
        ("len", u16p * 2), ("tig_off", u64p * 2), ("tig_arena", u8p), ("tig_arena_bytes", C.c_uint64),
        ("has_del", u8p * 2), ("cdr3_len", u16p * 2), ("cdr3_off", u64p * 2), ("cdr3_arena", u8p),
        ("cdr3_arena_bytes", C.c_uint64), ("cdr3_start", u16p * 2),
        ("v_ref_id", u32p * 2), ("j_ref_id", u32p * 2), ("dref_id", u32p * 2),
        ("vs_seq", u32p * 2), ("js_seq", u32p * 2), ("vname_id", u32p * 2), ("vuni_seq", u32p * 2),
        ("utr_seq", u32p * 2), ("c_ref_id_light", u32p), ("hcomp", u16p),
        ("fr1_start", u16p), ("cdr1_start", u16p), ("fr2_start", u16p), ("cdr2_start", u16p), ("fr3_start", u16p),
        ("nchains", u32p), ("ncells", u32p), ("donor_lo", u32p), ("donor_hi", u32p),
        ("refseq_off", u64p), ("refseq_len", u32p), ("refseq_arena", u8p), ("refseq_arena_bytes", C.c_uint64),
        ("tig2_arena", u64p), ("tig2_arena_words", C.c_uint64), ("entry_rec", C.POINTER(EjoinEntryRec)),
        ("cdr3_2bit", u64p), ("cdr3_2bit_words", C.c_uint64), ("cdr3_2bit_off", u32p),
    ]


class EjoinParams(C.Structure):
    _fields_ = [
        ("max_score", C.c_double), ("mult_pow", C.c_double), ("join_cdr3_ident", C.c_double),
        ("fwr1_cdr12_delta", C.c_double),
        ("cdr3_normal_len", C.c_int32), ("auto_share", C.c_int32), ("comp_filt", C.c_int32),
        ("comp_filt_bound", C.c_int32), ("max_cdr3_diffs", C.c_int32), ("cdr3_mult", C.c_int32),
        ("max_degradation", C.c_int32), ("max_diffs", C.c_int32), ("max_diffs_tcr", C.c_int32),
        ("ref_v_trim", C.c_int32), ("ref_j_trim", C.c_int32), ("easy", C.c_int32), ("old_light", C.c_int32),
        ("old_mult", C.c_int32), ("mix_donors", C.c_int32), ("degradation_mode", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


def default_params(**over):
    """Defaults of SURVEY.md §8(a) table P (help1.rs:300-363, cellranger.html.src:130-247)."""
    p = EjoinParams(
        max_score=100000.0, mult_pow=80.0, join_cdr3_ident=85.0, fwr1_cdr12_delta=20.0,
        cdr3_normal_len=42, auto_share=15, comp_filt=8, comp_filt_bound=80, max_cdr3_diffs=1000,
        cdr3_mult=5, max_degradation=2, max_diffs=1000000, max_diffs_tcr=5, ref_v_trim=15, ref_j_trim=15,
        easy=0, old_light=0, old_mult=0, mix_donors=0)
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class EjoinEdge(C.Structure):
    _fields_ = [
        ("k1", C.c_uint32), ("k2", C.c_uint32), ("shares", C.c_int32 * 2), ("indeps", C.c_int32 * 2),
        ("shares_details", (C.c_uint16 * 4) * 2), ("nrefs", C.c_uint16), ("cd", C.c_uint16),
        ("cd_chain", C.c_uint16 * 2), ("diffs", C.c_uint32), ("k", C.c_uint16), ("d", C.c_uint16),
        ("n", C.c_uint32), ("err", C.c_uint8), ("accept_reason", C.c_uint8), ("reserved", C.c_uint16),
        ("p1", C.c_double), ("mult", C.c_double), ("score", C.c_double),
    ]


EDGE_DTYPE = np.dtype([
    ("k1", "<u4"), ("k2", "<u4"), ("shares", "<i4", (2,)), ("indeps", "<i4", (2,)),
    ("shares_details", "<u2", (2, 4)), ("nrefs", "<u2"), ("cd", "<u2"), ("cd_chain", "<u2", (2,)),
    ("diffs", "<u4"), ("k", "<u2"), ("d", "<u2"), ("n", "<u4"), ("err", "u1"), ("accept_reason", "u1"),
    ("reserved", "<u2"), ("p1", "<f8"), ("mult", "<f8"), ("score", "<f8")])
assert EDGE_DTYPE.itemsize == C.sizeof(EjoinEdge), (EDGE_DTYPE.itemsize, C.sizeof(EjoinEdge))


class EjoinLogNames(C.Structure):
    _fields_ = [("index", C.c_int32)] + [(n, C.c_char_p) for n in ("id1", "id2", "vsegs1", "vsegs2", "dataset1", "dataset2", "cdr3_aa1", "cdr3_aa2",
                                                                   "bcs1", "bcs2")]


class EjoinStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_buckets", "n_groups", "pairs_candidate", "pairs_compared", "pairs_tier2", "pairs_accepted",
        "pairs_generic", "edges_after_prune", "gpu_launches", "h2d_bytes", "d2h_bytes",
        "alg_bytes_tier1", "alg_bytes_tier2", "alg_bytes_pack")] + [
        (n, C.c_double) for n in (
            "ms_h2d", "ms_device", "ms_d2h", "ms_host_post", "ms_total",
            "ms_kernel_tier1", "ms_kernel_tier2", "ms_kernel_pack", "ms_kernel_other")] + [
        ("alg_bytes_pack_t2", C.c_uint64), ("ms_kernel_pack_t2", C.c_double), ("mma_pairs", C.c_uint64), ("mma_macs", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EjoinResult(C.Structure):
    _fields_ = [("n_edges", C.c_uint64), ("edges", C.POINTER(EjoinEdge)), ("class_of", u32p),
                ("n_info", C.c_uint32), ("n_classes", C.c_uint32), ("owner", C.c_void_p), ("stats", EjoinStats)]


class EjoinRawEdge(C.Structure):
    _fields_ = [("k1", C.c_uint32), ("k2", C.c_uint32), ("cd", C.c_uint16), ("d", C.c_uint16), ("flags", C.c_uint32)]


RAW_EDGE_DTYPE = np.dtype([("k1", "<u4"), ("k2", "<u4"), ("cd", "<u2"), ("d", "<u2"), ("flags", "<u4")])
assert RAW_EDGE_DTYPE.itemsize == C.sizeof(EjoinRawEdge)

# (field, numpy dtype, per-chain?) for building an EjoinPack from numpy arrays
_PACK_ARRAYS = [
    ("exact_id", np.uint32, False), ("len", np.uint16, True), ("tig_off", np.uint64, True),
    ("has_del", np.uint8, True), ("cdr3_len", np.uint16, True), ("cdr3_off", np.uint64, True),
    ("cdr3_start", np.uint16, True), ("v_ref_id", np.uint32, True), ("j_ref_id", np.uint32, True),
    ("dref_id", np.uint32, True), ("vs_seq", np.uint32, True), ("js_seq", np.uint32, True),
    ("vname_id", np.uint32, True), ("vuni_seq", np.uint32, True), ("utr_seq", np.uint32, True),
    ("c_ref_id_light", np.uint32, False), ("hcomp", np.uint16, False),
    ("fr1_start", np.uint16, False), ("cdr1_start", np.uint16, False), ("fr2_start", np.uint16, False),
    ("cdr2_start", np.uint16, False), ("fr3_start", np.uint16, False),
    ("nchains", np.uint32, False), ("ncells", np.uint32, False), ("donor_lo", np.uint32, False),
    ("donor_hi", np.uint32, False), ("refseq_off", np.uint64, False), ("refseq_len", np.uint32, False),
]
_PACK_ARENAS = ["tig_arena", "cdr3_arena", "refseq_arena"]


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


_CT = {np.uint8: C.c_uint8, np.uint16: C.c_uint16, np.uint32: C.c_uint32, np.uint64: C.c_uint64}


class JoinPack:
    """A JoinPack held as numpy arrays plus the ctypes struct pointing at them.

    `arrays` maps field name -> ndarray (per-chain fields: a list of two arrays).
    """

    def __init__(self, arrays, is_bcr=True, owner=None):
        self.arrays = {}
        self.owner = owner
        st = EjoinPack()
        st.abi_version = EJOIN_ABI_VERSION
        st.is_bcr = 1 if is_bcr else 0
        for name, dt, per_chain in _PACK_ARRAYS:
            if per_chain:
                pair = [np.ascontiguousarray(arrays[name][c], dtype=dt) for c in range(2)]
                self.arrays[name] = pair
                getattr(st, name)[0] = _ptr(pair[0], _CT[dt])
                getattr(st, name)[1] = _ptr(pair[1], _CT[dt])
            else:
                a = np.ascontiguousarray(arrays[name], dtype=dt)
                self.arrays[name] = a
                setattr(st, name, _ptr(a, _CT[dt]))
        for name in _PACK_ARENAS:
            a = np.ascontiguousarray(arrays[name], dtype=np.uint8)
            if a.size == 0:
                a = np.zeros(16, np.uint8)
                nbytes = 0
            else:
                nbytes = a.size
            self.arrays[name] = a
            setattr(st, name, _ptr(a, C.c_uint8))
            setattr(st, name + "_bytes", nbytes)
        st.n_info = len(self.arrays["exact_id"])
        st.n_exact = len(self.arrays["nchains"])
        st.n_refseq = len(self.arrays["refseq_len"])
        self.struct = st

    @property
    def n_info(self):
        return int(self.struct.n_info)

    @classmethod
    def from_struct(cls, st, owner=None):
        """Copy a C-owned ejoin_pack_t (e.g. from the generator) into numpy arrays."""
        n, nx, nr = st.n_info, st.n_exact, st.n_refseq

        def arr(p, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(p, shape=(count,)).astype(dt, copy=True)

        counts = {"nchains": nx, "ncells": nx, "donor_lo": nx, "donor_hi": nx, "refseq_off": nr, "refseq_len": nr}
        arrays = {}
        for name, dt, per_chain in _PACK_ARRAYS:
            cnt = counts.get(name, n)
            if per_chain:
                arrays[name] = [arr(getattr(st, name)[c], cnt, dt) for c in range(2)]
            else:
                arrays[name] = arr(getattr(st, name), cnt, dt)
        for name in _PACK_ARENAS:
            arrays[name] = arr(getattr(st, name), int(getattr(st, name + "_bytes")), np.uint8)
        return cls(arrays, is_bcr=bool(st.is_bcr), owner=owner)

    def subset(self, idx):
        """A new JoinPack with only the info entries `idx` (kept in the given order)."""
        idx = np.asarray(idx, dtype=np.int64)
        a = self.arrays
        out = {}
        per_entry = {"exact_id", "c_ref_id_light", "hcomp", "fr1_start", "cdr1_start", "fr2_start", "cdr2_start", "fr3_start"}
        for name, dt, per_chain in _PACK_ARRAYS:
            if per_chain and name not in ("tig_off", "cdr3_off"):
                out[name] = [a[name][c][idx] for c in range(2)]
            elif name in per_entry:
                out[name] = a[name][idx]
            elif not per_chain:
                out[name] = a[name]
        for arena, off, ln in (("tig_arena", "tig_off", "len"), ("cdr3_arena", "cdr3_off", "cdr3_len")):
            chunks, offs, pos = [], [[], []], 0
            for k in idx:
                for c in range(2):
                    o, l = int(a[off][c][k]), int(a[ln][c][k])
                    chunks.append(a[arena][o:o + l])
                    offs[c].append(pos)
                    pos += l
            out[arena] = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
            out[off] = [np.asarray(offs[c], np.uint64) for c in range(2)]
        out["refseq_arena"] = a["refseq_arena"]
        return JoinPack(out, is_bcr=bool(self.struct.is_bcr))

    def tig(self, c, k):
        o, l = int(self.arrays["tig_off"][c][k]), int(self.arrays["len"][c][k])
        return self.arrays["tig_arena"][o:o + l].tobytes()

    def bucket_starts(self):
        l0, l1 = self.arrays["len"][0].astype(np.int64), self.arrays["len"][1].astype(np.int64)
        key = l0 * 65536 + l1
        if len(key) == 0:
            return np.zeros(1, np.int64)
        brk = np.flatnonzero(np.diff(key) != 0) + 1
        return np.concatenate([[0], brk, [len(key)]]).astype(np.int64)
