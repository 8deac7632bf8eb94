"""Post-join refinement (include/enclone_refine.h): ctypes view of the ABI, the flat pack, a synthetic workload.

Host-side mirror of what enclone does with `eq` and `raw_joins` after join_exacts: clonotype chain
grouping (define_mat, /root/reference/pages/heuristics.html.src:19-41), pure subclonotypes, the doublet / signature /
weak-chain filters (/root/reference/pages/default_filters.html.src:292-353) and the break-up of clonotypes no longer held
together by shared chains.  `refine_clonotypes(pack, doublet=True, ...)` takes the filter switches under the names of
ctl.clono_filt_opt_def (NDOUBLET = doublet=False, NSIG = signature=False, NWEAK_CHAINS = weak_chains=False).  Everything
runs on the GPU inside libenclone_join.so; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib

EREFINE_ABI_VERSION = 1
NONE = 0xFFFFFFFF
KEPT, DOUBLET, SIGNATURE, WEAK_CHAIN = 0, 1, 2, 3


class ErefinePack(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_info", C.c_uint32), ("n_exact", C.c_uint32), ("n_chain", C.c_uint32),
                ("exact_id", C.POINTER(C.c_uint32)), ("exact_cols", C.POINTER(C.c_uint8) * 2), ("class_of", C.POINTER(C.c_uint32)),
                ("n_joins", C.c_uint64), ("join_k1", C.POINTER(C.c_uint32)), ("join_k2", C.POINTER(C.c_uint32)),
                ("chain_off", C.POINTER(C.c_uint32)), ("ncells", C.POINTER(C.c_uint32)), ("origin_rank", C.POINTER(C.c_uint32)),
                ("seq_id", C.POINTER(C.c_uint32)), ("cdr3_id", C.POINTER(C.c_uint32)), ("order_rank", C.POINTER(C.c_uint32)),
                ("left", C.POINTER(C.c_uint8)), ("seq_len", C.POINTER(C.c_uint32)), ("seq_off", C.POINTER(C.c_uint64)),
                ("seq_arena", C.POINTER(C.c_uint8)), ("seq_arena_bytes", C.c_uint64)]


class ErefineOpts(C.Structure):
    _fields_ = [("doublet", C.c_int32), ("signature", C.c_int32), ("weak_chains", C.c_int32), ("doublet_mult", C.c_int32),
                ("signature_mult", C.c_int32), ("weak_max_cells", C.c_int32), ("weak_mult", C.c_int32), ("third_chain_diffs", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class ErefineStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_clonotypes_in", "n_clonotypes_out", "n_pure", "n_doublet", "n_signature", "n_weak",
                                          "triples_tested", "gpu_launches", "h2d_bytes")] + [("ms_device", C.c_double), ("ms_total", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ErefineResult(C.Structure):
    _fields_ = [("fate", C.POINTER(C.c_uint8)), ("clonotype_of", C.POINTER(C.c_uint32)), ("column_of", C.POINTER(C.c_uint32)),
                ("n_columns", C.POINTER(C.c_uint32))]


_FIELDS = (("exact_id", np.uint32), ("exact_cols0", np.uint8), ("exact_cols1", np.uint8), ("class_of", np.uint32), ("join_k1", np.uint32),
           ("join_k2", np.uint32), ("chain_off", np.uint32), ("ncells", np.uint32), ("origin_rank", np.uint32), ("seq_id", np.uint32),
           ("cdr3_id", np.uint32), ("order_rank", np.uint32), ("left", np.uint8), ("seq_len", np.uint32), ("seq_off", np.uint64),
           ("seq_arena", np.uint8))
_CT = {np.uint8: C.c_uint8, np.uint32: C.c_uint32, np.uint64: C.c_uint64}


class _Pinned:
    """numpy arrays in page-locked host memory (ejoin_host_alloc): uploads from them run at the copy engine's rate instead of
    going through the driver's staging buffer.  Falls back to ordinary arrays where there is no device to pin for."""

    def __init__(self):
        self.ptrs = []

    def copy(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.nbytes == 0:
            return arr
        lib = _lib.load()
        p = lib.ejoin_host_alloc(arr.nbytes)
        if not p:
            return arr
        self.ptrs.append(p)
        out = np.frombuffer((C.c_uint8 * arr.nbytes).from_address(p), dtype=arr.dtype).reshape(arr.shape)
        out[...] = arr
        return out

    def __del__(self):
        try:
            lib = _lib.load()
            for p in self.ptrs:
                lib.ejoin_host_free(p)
        except Exception:
            pass
        self.ptrs = []


class RefinePack:
    """The flat view of (exact_clonotypes, info, raw_joins, eq) the refinement reads; owns its numpy arrays
    (pinned=True: in page-locked memory, except the V..J arena, of which the library reads a few percent)."""

    def __init__(self, arrays, pinned=False):
        self.a = {k: np.ascontiguousarray(arrays[k], dt) for k, dt in _FIELDS}
        self._pin = None
        if pinned:
            self._pin = _Pinned()
            self.a = {k: (v if k == "seq_arena" else self._pin.copy(v)) for k, v in self.a.items()}
        a = self.a
        st = ErefinePack()
        st.abi_version = EREFINE_ABI_VERSION
        st.n_info, st.n_exact, st.n_chain = len(a["exact_id"]), len(a["chain_off"]) - 1, len(a["seq_id"])
        st.n_joins = len(a["join_k1"])
        for k, dt in _FIELDS:
            if k.startswith("exact_cols"):
                st.exact_cols[int(k[-1])] = a[k].ctypes.data_as(C.POINTER(C.c_uint8))
            else:
                setattr(st, k, a[k].ctypes.data_as(C.POINTER(_CT[dt])))
        st.seq_arena_bytes = a["seq_arena"].size
        self.struct = st

    n_info = property(lambda s: s.struct.n_info)
    n_exact = property(lambda s: s.struct.n_exact)
    n_chain = property(lambda s: s.struct.n_chain)


class RefineResult:
    def __init__(self, n_exact, n_chain, pinned=False):
        self._pin = _Pinned() if pinned else None
        mk = (lambda n, dt: self._pin.copy(np.empty(n, dt))) if pinned else (lambda n, dt: np.zeros(n, dt))
        self.fate = mk(max(n_exact, 1), np.uint8)
        self.clonotype_of = mk(max(n_exact, 1), np.uint32)
        self.column_of = mk(max(n_chain, 1), np.uint32)
        self.n_columns = mk(max(n_exact, 1), np.uint32)
        self.n_exact, self.n_chain = n_exact, n_chain
        st = ErefineResult()
        st.fate = self.fate.ctypes.data_as(C.POINTER(C.c_uint8))
        for k in ("clonotype_of", "column_of", "n_columns"):
            setattr(st, k, getattr(self, k).ctypes.data_as(C.POINTER(C.c_uint32)))
        self.struct = st

    def trim(self):
        if not hasattr(self, "_full"):
            self._full = (self.fate, self.clonotype_of, self.column_of, self.n_columns)      # kept for a later call that reuses this object
        self.fate, self.clonotype_of, self.n_columns = self._full[0][:self.n_exact], self._full[1][:self.n_exact], self._full[3][:self.n_exact]
        self.column_of = self._full[2][:self.n_chain]
        return self

    def same_as(self, other):
        return all(np.array_equal(getattr(self, k), getattr(other, k)) for k in ("fate", "clonotype_of", "column_of", "n_columns"))


def make_opts(**kw):
    """erefine_opts_t: enclone's defaults, then doublet= / signature= / weak_chains= (booleans) and the rule constants."""
    o = ErefineOpts()
    _bind().erefine_opts_default(C.byref(o))
    for k, v in kw.items():
        if k in ("doublet", "signature", "weak_chains"):
            setattr(o, k, int(bool(v)))
        elif k in ("doublet_mult", "signature_mult", "weak_max_cells", "weak_mult", "third_chain_diffs"):
            setattr(o, k, int(v))
        else:
            raise TypeError(f"unknown refinement option {k}")
    return o


_bound = False


def _bind():
    global _bound
    lib = _lib.load()
    if not _bound:
        lib.erefine_opts_default.argtypes = [C.POINTER(ErefineOpts)]
        lib.erefine_run.restype = C.c_int
        lib.erefine_run.argtypes = [C.POINTER(ErefinePack), C.POINTER(ErefineOpts), C.c_int, C.POINTER(ErefineResult), C.POINTER(ErefineStats)]
        lib.erefine_last_error.restype = C.c_char_p
        _bound = True
    return lib


def refine_clonotypes(pack, device=0, result=None, **opts):
    """(RefineResult, stats): fate / final clonotype / table column of every exact subclonotype and chain.
    `result`: a RefineResult to fill (e.g. a pinned one kept between calls) instead of a fresh one."""
    lib = _bind()
    o = make_opts(**opts)
    res, st = result or RefineResult(pack.n_exact, pack.n_chain), ErefineStats()
    res.fate, res.clonotype_of, res.column_of, res.n_columns = res._full if hasattr(res, "_full") else (res.fate, res.clonotype_of, res.column_of, res.n_columns)
    rc = lib.erefine_run(C.byref(pack.struct), C.byref(o), device, C.byref(res.struct), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"erefine_run: {rc}: {lib.erefine_last_error().decode()}")
    return res.trim(), st.as_dict()


# ---------------------------------------------------------------------------------------------
# Building a pack from Python objects (tests, fixtures) and a synthetic workload (tests, bench)
# ---------------------------------------------------------------------------------------------
def _components(n, a, b):
    """min-member labels of the graph on n nodes with edges (a[i], b[i])"""
    parent = np.arange(n, dtype=np.int64)

    def find(x):
        r = x
        while parent[r] != r:
            r = parent[r]
        while parent[x] != r:
            parent[x], x = r, parent[x]
        return r
    for u, v in zip(a, b):
        ru, rv = find(int(u)), find(int(v))
        if ru != rv:
            parent[max(ru, rv)] = min(ru, rv)
    return np.array([find(i) for i in range(n)], np.uint32)


def pack_from_exacts(exacts, joins, links=()):
    """exacts: list of dicts {ncells, origin (tuple of ints), chains: [{seq (str), cdr3_dna (str), cdr3_aa (str), chain_type (str)}],
    entries: [(m_heavy, m_light or None), ...]} — the info entries of the exact subclonotype as pairs of share indices;
    joins: [((x1, e1), (x2, e2)), ...] raw joins between entry e1 of exact x1 and entry e2 of exact x2.
    links: further pairs the caller's EquivRel holds together without a raw join (the onesie merger's).
    Interns the strings the way a caller of the C ABI would; class_of = what finish_join returns for these joins."""
    seqs, cdr3s, okeys, origins = {}, {}, [], sorted({tuple(e.get("origin", (0,))) for e in exacts})
    chain_off, ncells, origin_rank = [0], [], []
    seq_id, cdr3_id, left, seq_len, seq_off, arena = [], [], [], [], [], bytearray()
    exact_id, c0, c1, entry_index = [], [], [], {}
    for x, e in enumerate(exacts):
        ncells.append(e["ncells"])
        origin_rank.append(origins.index(tuple(e.get("origin", (0,)))))
        for ch in e["chains"]:
            seq_id.append(seqs.setdefault(ch["seq"], len(seqs)))
            cdr3_id.append(cdr3s.setdefault(ch["cdr3_dna"], len(cdr3s)))
            t = ch["chain_type"]
            t = "TRX" + t[3:] if t.startswith("TRB") else "TRY" + t[3:] if t.startswith("TRA") else t
            okeys.append((f"{t}:{ch['cdr3_aa']}", len(ch["seq"])))
            left.append(1 if ch["chain_type"][:3] in ("IGH", "TRB") else 0)
            seq_len.append(len(ch["seq"])); seq_off.append(len(arena)); arena += ch["seq"].encode()
        chain_off.append(len(seq_id))
        for i, (mh, ml) in enumerate(e["entries"]):
            entry_index[(x, i)] = len(exact_id)
            exact_id.append(x); c0.append(mh); c1.append(0xFF if ml is None else ml)
    ranks = {k: i for i, k in enumerate(sorted(set(okeys)))}
    k1 = [min(entry_index[a], entry_index[b]) for a, b in joins]
    k2 = [max(entry_index[a], entry_index[b]) for a, b in joins]
    n = len(exact_id)
    # finish_join: the joins, plus links between the info entries of one exact subclonotype
    first = {}
    la, lb = list(k1) + [entry_index[a] for a, _ in links], list(k2) + [entry_index[b] for _, b in links]
    for k, x in enumerate(exact_id):
        if x in first:
            la.append(first[x]); lb.append(k)
        else:
            first[x] = k
    return RefinePack(dict(exact_id=exact_id, exact_cols0=c0, exact_cols1=c1, class_of=_components(n, la, lb), join_k1=k1, join_k2=k2,
                           chain_off=chain_off, ncells=ncells, origin_rank=origin_rank, seq_id=seq_id, cdr3_id=cdr3_id,
                           order_rank=[ranks[k] for k in okeys], left=left, seq_len=seq_len, seq_off=seq_off,
                           seq_arena=np.frombuffer(bytes(arena) + b"\0" * 16, np.uint8)))


_AA = "ACDEFGHIKLMNPQRSTVWY"


def synth_exacts(n_families, seed=1, frac_doublet=0.04, frac_three=0.06, frac_onesie=0.08, frac_weak=0.03, frac_sig=0.02, big=0.15):
    """Synthetic repertoire for the refinement: clonal families of two-chain exact subclonotypes joined by a random tree of raw
    joins (plus a few redundant ones), with onesies that repeat a member's chain, three-chain members whose third chains are
    near-identical or unrelated, doublets carrying a chain of another family, rare chains (weak), and small three-chain
    contaminants glued onto large families (signature).  Returns (exacts, joins, links) for pack_from_exacts."""
    rng = np.random.default_rng(seed)
    B = np.frombuffer(b"ACGT", np.uint8)

    def rseq(n):
        return B[rng.integers(0, 4, n)]

    def mutate(s, k):
        s = s.copy()
        for p in rng.integers(0, len(s), k):
            s[p] = B[(np.searchsorted(B, s[p]) + 1 + rng.integers(0, 3)) % 4]
        return s

    def chain(s, c3, kind):
        cdr3 = s[c3:c3 + 36].tobytes().decode()
        aa = "".join(_AA[(ord(cdr3[i]) * 7 + ord(cdr3[i + 1]) * 3 + ord(cdr3[i + 2])) % 20] for i in range(0, 36, 3))
        return dict(seq=s.tobytes().decode(), cdr3_dna=cdr3, cdr3_aa="C" + aa, chain_type=kind)

    fams, exacts, joins, links = [], [], [], []
    for f in range(n_families):
        size = int(rng.integers(8, 40)) if rng.random() < big else int(rng.integers(1, 4))
        h0, l0 = rseq(int(rng.integers(340, 380))), rseq(int(rng.integers(318, 336)))
        kind_l = "IGK" if rng.random() < 0.6 else "IGL"
        members = []
        for i in range(size):
            h = mutate(h0, int(rng.integers(0, 6))) if i else h0
            l = mutate(l0, int(rng.integers(0, 3))) if i else l0
            ex = dict(ncells=int(rng.integers(1, 4)) if size < 8 else int(rng.integers(1, 12)), origin=(int(rng.integers(0, 2)),),
                      chains=[chain(h, 290, "IGH"), chain(l, 270, kind_l)], entries=[(0, 1)])
            x = len(exacts)
            exacts.append(ex)
            if members:
                y = members[int(rng.integers(0, len(members)))]
                joins.append(((y, 0), (x, 0)))
                if len(members) > 2 and rng.random() < 0.2:
                    z = members[int(rng.integers(0, len(members)))]
                    if z != y:
                        joins.append(((z, 0), (x, 0)))
            members.append(x)
        fams.append(dict(members=members, h0=h0, l0=l0, kind_l=kind_l))
    nf = len(fams)
    for f, F in enumerate(fams):
        m = F["members"]
        # onesies: one chain identical to a member's (or to nobody's)
        if rng.random() < frac_onesie * len(m):
            y = m[int(rng.integers(0, len(m)))]
            which = int(rng.integers(0, 2))
            ch = dict(exacts[y]["chains"][which]) if rng.random() < 0.8 else chain(mutate(F["l0"], 4), 270, F["kind_l"])
            x = len(exacts)
            exacts.append(dict(ncells=int(rng.integers(1, 3)), origin=(0,), chains=[ch], entries=[(0, None)]))
            if rng.random() < 0.9:           # a onesie has no raw join: the onesie merger put it into the member's class
                links.append(((y, 0), (x, 0)))
        # three-chain members: heavy + two lights, the second light near the family's or unrelated
        for _ in range(int(rng.random() < frac_three * len(m)) * 2):
            y = m[int(rng.integers(0, len(m)))]
            l2 = mutate(F["l0"], int(rng.integers(1, 25))) if rng.random() < 0.7 else rseq(len(F["l0"]))
            h = mutate(F["h0"], 2)
            ex = dict(ncells=int(rng.integers(1, 4)), origin=(0,), chains=[chain(h, 290, "IGH"), chain(mutate(F["l0"], 1), 270, F["kind_l"]), chain(l2, 270, F["kind_l"])],
                      entries=[(0, 1), (0, 2)])
            x = len(exacts)
            exacts.append(ex)
            joins.append(((y, 0), (x, 0)))
            F.setdefault("threes", []).append(x)
        th = F.get("threes", [])
        if len(th) == 2:
            joins.append(((th[0], 0), (th[1], 0)))
        # doublets: a member's two chains plus one or two chains of a member of another family
        if nf > 1 and rng.random() < frac_doublet * len(m):
            g = int(rng.integers(0, nf - 1)); g += g >= f
            y, z = m[int(rng.integers(0, len(m)))], fams[g]["members"][int(rng.integers(0, len(fams[g]["members"])))]
            cy, cz = exacts[y]["chains"], exacts[z]["chains"]
            if rng.random() < 0.5:
                ex = dict(ncells=1, origin=(0,), chains=[dict(cy[0]), dict(cy[1]), dict(cz[1])], entries=[(0, 1), (0, 2)])
            else:
                ex = dict(ncells=int(rng.integers(1, 3)), origin=(0,), chains=[dict(cy[0]), dict(cz[0]), dict(cy[1]), dict(cz[1])],
                          entries=[(0, 2), (0, 3), (1, 2), (1, 3)])
            x = len(exacts)
            exacts.append(ex)
            joins.append(((y, 0), (x, 0)))
        # weak chain: a rare extra light chain carried by one or two small members
        if len(m) >= 8 and rng.random() < frac_weak * 10:
            lw = chain(rseq(len(F["l0"])), 270, F["kind_l"])
            for _ in range(int(rng.integers(1, 3))):
                y = m[int(rng.integers(0, len(m)))]
                ex = dict(ncells=1, origin=(0,), chains=[dict(exacts[y]["chains"][0]), dict(exacts[y]["chains"][1]), dict(lw)], entries=[(0, 1), (0, 2)])
                x = len(exacts)
                exacts.append(ex)
                joins.append(((y, 0), (x, 0)))
        # signature: a one-cell three-chain exact gluing two large two-chain families through raw joins on both of its entries
        if nf > 1 and len(m) >= 8 and rng.random() < frac_sig * 10:
            g = int(rng.integers(0, nf - 1)); g += g >= f
            if len(fams[g]["members"]) >= 8:
                y, z = m[0], fams[g]["members"][0]
                cy, cz = exacts[y]["chains"], exacts[z]["chains"]
                ex = dict(ncells=1, origin=(0,), chains=[dict(cy[0]), dict(cy[1]), dict(cz[1])], entries=[(0, 1), (0, 2)])
                x = len(exacts)
                exacts.append(ex)
                joins.append(((y, 0), (x, 0)))
                joins.append(((z, 0), (x, 1)))
    return exacts, joins, links


def synth_refine_pack(n_families, seed=1, **kw):
    return pack_from_exacts(*synth_exacts(n_families, seed, **kw))


def tile_pack(pack, times, pinned=False):
    """`times` independent copies of a pack side by side (ids shifted): the bench's way to a large input."""
    a, out = pack.a, {}
    n, nx, nc = pack.n_info, pack.n_exact, pack.n_chain
    nseq, ncdr, nbytes = int(a["seq_id"].max()) + 1, int(a["cdr3_id"].max()) + 1, a["seq_arena"].size
    t = np.arange(times, dtype=np.uint64)

    def rep(v, step, dt):
        return (v[None, :].astype(np.uint64) + (t * np.uint64(step))[:, None]).reshape(-1).astype(dt)
    out["exact_id"] = rep(a["exact_id"], nx, np.uint32)
    out["exact_cols0"], out["exact_cols1"] = np.tile(a["exact_cols0"], times), np.tile(a["exact_cols1"], times)
    out["class_of"] = rep(a["class_of"], n, np.uint32)
    out["join_k1"], out["join_k2"] = rep(a["join_k1"], n, np.uint32), rep(a["join_k2"], n, np.uint32)
    out["chain_off"] = np.concatenate([rep(a["chain_off"][:-1], nc, np.uint32), np.array([nc * times], np.uint32)])
    out["ncells"], out["origin_rank"] = np.tile(a["ncells"], times), np.tile(a["origin_rank"], times)
    out["seq_id"], out["cdr3_id"] = rep(a["seq_id"], nseq, np.uint32), rep(a["cdr3_id"], ncdr, np.uint32)
    out["order_rank"], out["left"], out["seq_len"] = np.tile(a["order_rank"], times), np.tile(a["left"], times), np.tile(a["seq_len"], times)
    out["seq_off"] = rep(a["seq_off"], nbytes, np.uint64)
    out["seq_arena"] = np.tile(a["seq_arena"], times)
    return RefinePack(out, pinned=pinned)
