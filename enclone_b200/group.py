"""Symmetric clonotype grouping (include/enclone_group.h): ctypes view of the ABI, the flat pack, a synthetic workload.

Host-side mirror of enclone_tail::grouper::grouper(), Case 1 (/root/reference/enclone_tail/src/grouper.rs:55-766):
`group_clonotypes(pack, **opts)` takes the options of ctl.clono_group_opt under their own names (vj_refname, heavy_pc, ...)
and returns group_of[i] = smallest clonotype index of i's group.  The distance passes run on the GPU inside
libenclone_join.so; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib

EGROUP_ABI_VERSION = 1
HF_ORDER = "ACDEFGHIKLMNPQRSTVWY"


class EgroupPack(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_clono", C.c_uint32), ("n_exact", C.c_uint32), ("n_chain", C.c_uint32),
                ("clono_exact_off", C.POINTER(C.c_uint32)), ("clono_exact", C.POINTER(C.c_uint32)), ("exact_chain_off", C.POINTER(C.c_uint32)),
                ("left", C.POINTER(C.c_uint8)), ("seq_off", C.POINTER(C.c_uint64)), ("seq_len", C.POINTER(C.c_uint32)), ("seq_del_len", C.POINTER(C.c_uint32)),
                ("cdr3_start", C.POINTER(C.c_uint32)), ("cdr3_aa_off", C.POINTER(C.c_uint64)), ("cdr3_aa_len", C.POINTER(C.c_uint32)),
                ("seq_arena", C.POINTER(C.c_uint8)), ("seq_arena_bytes", C.c_uint64), ("aa_arena", C.POINTER(C.c_uint8)), ("aa_arena_bytes", C.c_uint64),
                ("col_off", C.POINTER(C.c_uint32)), ("col_vname", C.POINTER(C.c_uint32)), ("col_jname", C.POINTER(C.c_uint32)), ("col_dname", C.POINTER(C.c_uint32)),
                ("col_left", C.POINTER(C.c_uint8))]


class EgroupOpts(C.Structure):
    _fields_ = [("vj_refname", C.c_int32), ("vdj_refname", C.c_int32), ("v_heavy_refname", C.c_int32), ("vj_heavy_refname", C.c_int32), ("vdj_heavy_refname", C.c_int32),
                ("vj_len", C.c_int32), ("cdr3_len", C.c_int32), ("cdr3_heavy_len", C.c_int32), ("cdr3_light_len", C.c_int32), ("reserved", C.c_int32 * 3),
                ("heavy_pc", C.c_double), ("light_pc", C.c_double), ("cdr3_heavy_pc_hf", C.c_double), ("hf_matrix", C.POINTER(C.c_double)),
                ("aa_heavy_pc", C.c_double), ("aa_light_pc", C.c_double), ("cdr3_heavy_pc", C.c_double), ("cdr3_light_pc", C.c_double),
                ("cdr3_aa_heavy_pc", C.c_double), ("cdr3_aa_light_pc", C.c_double)]


class EgroupStats(C.Structure):
    _fields_ = [("pairs_candidate", C.c_uint64), ("pairs_evaluated", C.c_uint64), ("chain_compares", C.c_uint64), ("gpu_launches", C.c_uint64),
                ("n_groups", C.c_uint64), ("ms_device", C.c_double), ("ms_total", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class EgroupAsymOpts(C.Structure):
    _fields_ = [("bound_kind", C.c_int32), ("top", C.c_uint32), ("max_dist", C.c_double), ("min_group", C.c_uint32), ("reserved", C.c_uint32)]


class EgroupAsymResult(C.Structure):
    _fields_ = [("n_groups", C.c_uint64), ("group_off", C.POINTER(C.c_uint64)), ("member", C.POINTER(C.c_uint32)), ("dist", C.POINTER(C.c_double)),
                ("owner", C.c_void_p)]


BOUND_TOP, BOUND_MAX, BOUND_NONE = 0, 1, 2

KEY_FLAGS = ("vj_refname", "vdj_refname", "v_heavy_refname", "vj_heavy_refname", "vdj_heavy_refname", "vj_len", "cdr3_len", "cdr3_heavy_len", "cdr3_light_len")
DIST_OPTS = ("heavy_pc", "light_pc", "cdr3_heavy_pc_hf", "aa_heavy_pc", "aa_light_pc", "cdr3_heavy_pc", "cdr3_light_pc", "cdr3_aa_heavy_pc", "cdr3_aa_light_pc")

_FIELDS = (("clono_exact_off", np.uint32), ("clono_exact", np.uint32), ("exact_chain_off", np.uint32), ("left", np.uint8), ("seq_off", np.uint64),
           ("seq_len", np.uint32), ("seq_del_len", np.uint32), ("cdr3_start", np.uint32), ("cdr3_aa_off", np.uint64), ("cdr3_aa_len", np.uint32),
           ("seq_arena", np.uint8), ("aa_arena", np.uint8), ("col_off", np.uint32), ("col_vname", np.uint32), ("col_jname", np.uint32), ("col_dname", np.uint32),
           ("col_left", np.uint8))


class GroupPack:
    """The flat view of (exacts, exact_clonotypes, rsi, opt_d_val, refdata.name) the grouping reads; owns its numpy arrays."""

    def __init__(self, arrays):
        self.a = {k: np.ascontiguousarray(arrays[k], dt) for k, dt in _FIELDS}
        a = self.a
        st = EgroupPack()
        st.abi_version = EGROUP_ABI_VERSION
        st.n_clono = len(a["clono_exact_off"]) - 1
        st.n_exact = len(a["exact_chain_off"]) - 1
        st.n_chain = len(a["left"])
        for k, dt in _FIELDS:
            ct = {np.uint8: C.c_uint8, np.uint32: C.c_uint32, np.uint64: C.c_uint64}[dt]
            setattr(st, k, a[k].ctypes.data_as(C.POINTER(ct)))
        st.seq_arena_bytes = a["seq_arena"].size
        st.aa_arena_bytes = a["aa_arena"].size
        self.struct = st

    @property
    def n_clono(self):
        return self.struct.n_clono


def make_opts(hf_matrix=None, **kw):
    """egroup_opts_t from the option names of ctl.clono_group_opt; (opts, keepalive)."""
    o = EgroupOpts()
    lib = _bind()
    lib.egroup_opts_default(C.byref(o))
    keep = None
    for k, v in kw.items():
        if k in KEY_FLAGS:
            setattr(o, k, int(bool(v)))
        elif k in DIST_OPTS:
            setattr(o, k, -1.0 if v is None else float(v))
        else:
            raise TypeError(f"unknown grouping option {k}")
    if hf_matrix is not None:
        keep = np.ascontiguousarray(hf_matrix, np.float64).reshape(20, 20)
        o.hf_matrix = keep.ctypes.data_as(C.POINTER(C.c_double))
    return o, keep


_bound = False


def _bind():
    global _bound
    L = _lib.load()
    if not _bound:
        L.egroup_opts_default.argtypes = [C.POINTER(EgroupOpts)]
        L.egroup_run.argtypes = [C.POINTER(EgroupPack), C.POINTER(EgroupOpts), C.c_int, C.POINTER(C.c_uint32), C.POINTER(EgroupStats)]
        L.egroup_run.restype = C.c_int
        L.egroup_last_error.restype = C.c_char_p
        L.egroup_run_asym.argtypes = [C.POINTER(EgroupPack), C.POINTER(C.c_uint8), C.POINTER(EgroupAsymOpts), C.c_int, C.POINTER(EgroupAsymResult), C.POINTER(EgroupStats)]
        L.egroup_run_asym.restype = C.c_int
        L.egroup_asym_free.argtypes = [C.POINTER(EgroupAsymResult)]
        _bound = True
    return L


def group_clonotypes(pack, device=0, hf_matrix=None, **opts):
    """(group_of, stats) — the partition of grouper.rs:760-766 as min-member labels."""
    L = _bind()
    o, keep = make_opts(hf_matrix=hf_matrix, **opts)
    out = np.zeros(max(pack.n_clono, 1), np.uint32)
    st = EgroupStats()
    rc = L.egroup_run(C.byref(pack.struct), C.byref(o), device, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(st))
    if rc != 0:
        raise _lib.EjoinError(rc, (L.egroup_last_error() or b"").decode())
    del keep
    return out[:pack.n_clono], st.as_dict()


def asym_opts(top=None, max_dist=None, min_group=1):
    """AG_DIST_BOUND=top=N or max=D (one of them, as in the reference; neither = no bound), MIN_GROUP"""
    o = EgroupAsymOpts()
    if top is not None and max_dist is not None:
        raise TypeError("AG_DIST_BOUND is top=N or max=D, not both")
    o.bound_kind = BOUND_TOP if top is not None else (BOUND_MAX if max_dist is not None else BOUND_NONE)
    o.top = int(top or 0)
    o.max_dist = float(max_dist if max_dist is not None else 0.0)
    o.min_group = int(min_group)
    return o


def result_groups(res):
    """[[(clonotype, distance), ...], ...] — the center first, with distance 0"""
    out = []
    for g in range(res.n_groups):
        a, b = res.group_off[g], res.group_off[g + 1]
        out.append([(int(res.member[k]), float(res.dist[k])) for k in range(a, b)])
    return out


def group_asymmetric(pack, in_center, top=None, max_dist=None, min_group=1, device=0):
    """Case 2 of grouper() (AGROUP): (groups, stats); in_center[i] != 0 marks the center clonotypes."""
    L = _bind()
    o = asym_opts(top, max_dist, min_group)
    cen = np.ascontiguousarray(in_center, np.uint8)
    assert cen.size == pack.n_clono
    res, st = EgroupAsymResult(), EgroupStats()
    rc = L.egroup_run_asym(C.byref(pack.struct), cen.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(o), device, C.byref(res), C.byref(st))
    if rc != 0:
        raise _lib.EjoinError(rc, (L.egroup_last_error() or b"").decode())
    try:
        return result_groups(res), st.as_dict()
    finally:
        L.egroup_asym_free(C.byref(res))


# ---- synthetic workload ----------------------------------------------------------------------------------------------
_CODON = {}
_TAB = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"
for _i in range(64):
    _CODON["ACGT"[_i >> 4] + "ACGT"[(_i >> 2) & 3] + "ACGT"[_i & 3]] = _TAB[_i]


def translate(s):
    return "".join(_CODON.get(s[i:i + 3], "X") for i in range(0, len(s) - 2, 3))


def synth_group_pack(n_clono, seed=0, n_families=None, heavy_len=360, light_len=330, sub_rate=0.04, indel_rate=0.15, n_genes=12):
    """Clonotypes drawn from `n_families` lineage roots: a member's chains are its root's with substitutions (rate up to
    `sub_rate`) and, with probability `indel_rate`, a codon inserted or deleted (in the CDR3 or elsewhere), so that the pairwise
    distances straddle the usual thresholds.  1-3 exact subclonotypes per clonotype, 1-3 chains each (some without a heavy
    chain), gene names per column from `n_genes` names."""
    rng = np.random.default_rng(seed)
    n_families = n_families or max(1, n_clono // 6)
    B = np.frombuffer(b"ACGT", np.uint8)

    def rand_seq(n):
        # no stop codons inside, so CDR3 amino acids stay letters (the HF penalty matrix indexes letters only)
        s = list(B[rng.integers(0, 4, n)].tobytes().decode())
        for i in range(0, n - 2, 3):
            if _CODON["".join(s[i:i + 3])] == "*":
                s[i] = "C"
        return "".join(s)

    roots = []
    for _ in range(n_families):
        hl = heavy_len + 3 * int(rng.integers(-3, 4)); ll = light_len + 3 * int(rng.integers(-2, 3))
        roots.append(dict(h=rand_seq(hl), l=rand_seq(ll), hc=(3 * int(rng.integers(88, 96)), int(rng.integers(10, 22))), lc=(3 * int(rng.integers(85, 92)), int(rng.integers(8, 14))),
                          hv=int(rng.integers(0, n_genes)), hd=int(rng.integers(0, n_genes + 1)), hj=int(rng.integers(0, 6)), lv=int(rng.integers(0, n_genes)), lj=int(rng.integers(0, 5))))

    def mutate(s, cdr3):
        s = list(s)
        st, aal = cdr3
        rate = float(rng.random()) * sub_rate
        for i in np.nonzero(rng.random(len(s)) < rate)[0]:
            s[i] = "ACGT"[int(rng.integers(0, 4))]
        if rng.random() < indel_rate:
            pos = 3 * int(rng.integers(2, len(s) // 3 - 2))
            if rng.random() < 0.5:
                s[pos:pos] = list("GCT")
                if pos <= st: st += 3
                elif pos < st + 3 * aal: aal += 1
            else:
                del s[pos:pos + 3]
                if pos + 3 <= st: st -= 3
                elif pos < st + 3 * aal and aal > 4: aal -= 1
        s = "".join(s)
        for i in range(0, len(s) - 2, 3):                      # keep the frame free of stops
            if _CODON[s[i:i + 3]] == "*":
                s = s[:i] + "C" + s[i + 1:]
        st = min(st, len(s) - 3 * aal - (len(s) - 3 * aal) % 3)
        return s, (st, aal)

    clono_exact_off, clono_exact, exact_chain_off = [0], [], [0]
    left, seqs, seq_del_len, cdr3_start, aas = [], [], [], [], []
    col_off, col_v, col_j, col_d, col_left = [0], [], [], [], []
    null_id = n_genes + 16
    for _ in range(n_clono):
        r = roots[int(rng.integers(0, n_families))]
        shape = rng.random()
        cols = [("h", True), ("l", False)] if shape < 0.8 else ([("l", False)] if shape < 0.9 else [("h", True), ("l", False), ("l", False)])
        for c, is_left in cols:
            col_v.append(r["hv"] if is_left else 20 + r["lv"]); col_j.append(40 + (r["hj"] if is_left else 8 + r["lj"]))
            col_d.append(60 + r["hd"] if is_left and r["hd"] < n_genes else null_id); col_left.append(int(is_left))
        col_off.append(len(col_v))
        for _x in range(int(rng.choice([1, 1, 1, 2, 3]))):
            clono_exact.append(len(exact_chain_off) - 1)
            for c, is_left in cols:
                s, (st, aal) = mutate(r[c], r["hc" if is_left else "lc"])
                left.append(int(is_left)); seqs.append(s); cdr3_start.append(st); aas.append(translate(s[st:st + 3 * aal]))
                seq_del_len.append(len(s) + (3 if rng.random() < 0.05 else 0))
            exact_chain_off.append(len(left))
        clono_exact_off.append(len(clono_exact))
    seq_off = np.zeros(len(seqs), np.uint64); aa_off = np.zeros(len(seqs), np.uint64)
    sa, aa = bytearray(), bytearray()
    for i, (s, a) in enumerate(zip(seqs, aas)):
        seq_off[i] = len(sa); sa += s.encode(); aa_off[i] = len(aa); aa += a.encode()
    return GroupPack(dict(clono_exact_off=clono_exact_off, clono_exact=clono_exact, exact_chain_off=exact_chain_off, left=left, seq_off=seq_off,
                          seq_len=[len(s) for s in seqs], seq_del_len=seq_del_len, cdr3_start=cdr3_start, cdr3_aa_off=aa_off, cdr3_aa_len=[len(a) for a in aas],
                          seq_arena=np.frombuffer(bytes(sa), np.uint8), aa_arena=np.frombuffer(bytes(aa), np.uint8), col_off=col_off, col_vname=col_v,
                          col_jname=col_j, col_dname=col_d, col_left=col_left))


def hf_matrix_example(seed=7):
    """A symmetric 20 x 20 substitution-penalty matrix in [0, 1] with a zero diagonal (stand-in for the user's file)."""
    rng = np.random.default_rng(seed)
    m = np.round(rng.random((20, 20)), 3)
    m = (m + m.T) / 2
    np.fill_diagonal(m, 0.0)
    return m
