"""Synthetic repertoire generator (tools/synth.cpp) -> JoinPack.  Test/bench infrastructure.

The five BASELINE.json configurations are named here; seeds are
20261019 + config id.
"""
import ctypes as C
import os

import numpy as np

from ._abi import EjoinPack, JoinPack

_HERE = os.path.dirname(os.path.abspath(__file__))


class SynthCfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_cells", C.c_uint32), ("n_donors", C.c_uint32), ("is_bcr", C.c_uint32),
        ("mode", C.c_uint32), ("frac_naive", C.c_double), ("family_alpha", C.c_double),
        ("frac_public", C.c_double), ("frac_indel", C.c_double), ("frac_3chain", C.c_double),
        ("frac_4chain", C.c_double), ("frac_1chain", C.c_double), ("max_family", C.c_uint32),
        ("reserved", C.c_uint32), ("alloc", C.c_void_p), ("dealloc", C.c_void_p)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libejoin_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(path)
        _lib.esynth_cfg_default.argtypes = [C.POINTER(SynthCfg)]
        _lib.esynth_generate.argtypes = [C.POINTER(SynthCfg), C.POINTER(C.POINTER(EjoinPack)),
                                         C.POINTER(C.POINTER(C.c_uint32))]
        _lib.esynth_generate.restype = C.c_int
        _lib.esynth_free.argtypes = [C.POINTER(EjoinPack)]
        _lib.esynth_free_truth.argtypes = [C.POINTER(C.c_uint32)]
    return _lib


class _Owner:
    """Keeps the generator-owned pack alive; JoinPack views point into it."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        if self.ptr:
            self.lib.esynth_free(self.ptr)
            self.ptr = None


class SynthPack:
    """Zero-copy handle on a generated pack: `.struct` is the C ejoin_pack_t."""

    def __init__(self, lib, ptr, truth):
        self._owner = _Owner(lib, ptr)
        self.struct = ptr.contents
        self.truth = truth

    @property
    def n_info(self):
        return int(self.struct.n_info)

    def to_joinpack(self):
        return JoinPack.from_struct(self.struct)


def pinned_allocator():
    """(alloc, dealloc) function addresses of the CUDA library's pinned allocator: arenas generated with
    them are page-locked, padded and flagged EJOIN_F_TIGS_IN_PLACE (the library reads them in place)."""
    from . import _lib
    L = _lib.load()
    return dict(alloc=C.cast(L.ejoin_host_alloc, C.c_void_p).value, dealloc=C.cast(L.ejoin_host_free, C.c_void_p).value)


def generate(n_cells=10000, n_donors=1, is_bcr=True, mode=0, seed=20261019, copy=True, pinned=False, **kw):
    """Generate a repertoire.  copy=True returns a numpy-backed JoinPack (+ truth family ids);
    copy=False returns a SynthPack viewing the generator's own buffers (for big inputs);
    pinned=True (with copy=False) puts the arenas in pinned host memory."""
    lib = _load()
    if pinned:
        kw = dict(kw, **pinned_allocator())
    cfg = SynthCfg()
    lib.esynth_cfg_default(C.byref(cfg))
    cfg.seed, cfg.n_cells, cfg.n_donors, cfg.is_bcr, cfg.mode = seed, n_cells, n_donors, int(is_bcr), mode
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    pp = C.POINTER(EjoinPack)()
    tp = C.POINTER(C.c_uint32)()
    rc = lib.esynth_generate(C.byref(cfg), C.byref(pp), C.byref(tp))
    if rc != 0:
        raise RuntimeError(f"esynth_generate failed: {rc}")
    n = pp.contents.n_info
    truth = np.ctypeslib.as_array(tp, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
    lib.esynth_free_truth(tp)
    sp = SynthPack(lib, pp, truth)
    if copy:
        return sp.to_joinpack(), truth
    return sp, truth


# BASELINE.json configs.  `scale` shrinks a config for tests.
CONFIGS = {
    "C1": dict(n_cells=1654, n_donors=1, is_bcr=True, mode=0, seed=20261020),
    "C2": dict(n_cells=100_000, n_donors=1, is_bcr=True, mode=0, seed=20261021),
    "C3": dict(n_cells=1_600_000, n_donors=8, is_bcr=True, mode=0, seed=20261022),
    "C4": dict(n_cells=200_000, n_donors=1, is_bcr=True, mode=1, seed=20261023),
    "C5": dict(n_cells=1_000_000, n_donors=4, is_bcr=False, mode=0, seed=20261024),
}


def generate_config(name, scale=1.0, copy=True, pinned=False):
    cfg = dict(CONFIGS[name])
    cfg["n_cells"] = max(16, int(cfg["n_cells"] * scale))
    return generate(copy=copy, pinned=pinned, **cfg)
