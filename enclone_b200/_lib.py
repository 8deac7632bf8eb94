"""Loads libenclone_join.so (the CUDA product library).  No fallback: if the library is
missing or no CUDA device is usable, calls raise."""
import ctypes as C
import os

from ._abi import EjoinEdge, EjoinLogNames, EjoinPack, EjoinParams, EjoinRawEdge, EjoinResult, EjoinStats

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libenclone_join.so")

# every symbol include/enclone_join.h declares
EXPORTS = [
    "ejoin_params_default", "ejoin_create", "ejoin_destroy", "ejoin_last_error", "ejoin_run", "ejoin_result_free",
    "ejoin_plan_shards", "ejoin_run_shard", "ejoin_run_shard_final", "ejoin_partition_gpu", "ejoin_raw_free", "ejoin_finish", "ejoin_upload", "ejoin_upload_shard", "ejoin_run_resident",
    "ejoin_host_alloc", "ejoin_host_free", "ejoin_host_register", "ejoin_host_unregister", "ejoin_p1", "ejoin_partition",
    "ejoin_pack_compact", "ejoin_pack_compact_free", "ejoin_comm_id", "ejoin_comm_init", "ejoin_run_sharded", "ejoin_format_join_log", "ejoin_stage_time", "ejoin_set_stage_timing",
]

_lib = None


class EjoinError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libenclone_join error {code}: {msg}")
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the join has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    H = C.c_void_p
    L.ejoin_params_default.argtypes = [C.POINTER(EjoinParams)]
    L.ejoin_create.argtypes = [C.POINTER(H), C.POINTER(EjoinParams), C.c_int]
    L.ejoin_create.restype = C.c_int
    L.ejoin_destroy.argtypes = [H]
    L.ejoin_last_error.argtypes = [H]
    L.ejoin_last_error.restype = C.c_char_p
    L.ejoin_run.argtypes = [H, C.POINTER(EjoinPack), C.POINTER(EjoinResult)]
    L.ejoin_run.restype = C.c_int
    L.ejoin_result_free.argtypes = [C.POINTER(EjoinResult)]
    L.ejoin_plan_shards.argtypes = [C.POINTER(EjoinPack), C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32,
                                    C.POINTER(C.c_uint32)]
    L.ejoin_plan_shards.restype = C.c_int
    L.ejoin_run_shard.argtypes = [H, C.POINTER(EjoinPack), C.c_int, C.c_int, C.POINTER(C.POINTER(EjoinRawEdge)),
                                  C.POINTER(C.c_uint64), C.POINTER(EjoinStats)]
    L.ejoin_run_shard.restype = C.c_int
    L.ejoin_run_shard_final.argtypes = [H, C.POINTER(EjoinPack), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(EjoinResult),
                                        C.POINTER(C.POINTER(EjoinRawEdge)), C.POINTER(C.c_uint64), C.POINTER(EjoinStats)]
    L.ejoin_run_shard_final.restype = C.c_int
    L.ejoin_partition_gpu.argtypes = [H, C.POINTER(EjoinPack), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint64,
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.ejoin_partition_gpu.restype = C.c_int
    L.ejoin_raw_free.argtypes = [C.POINTER(EjoinRawEdge)]
    L.ejoin_finish.argtypes = [H, C.POINTER(EjoinPack), C.POINTER(EjoinRawEdge), C.c_uint64, C.POINTER(EjoinResult)]
    L.ejoin_finish.restype = C.c_int
    L.ejoin_upload.argtypes = [H, C.POINTER(EjoinPack)]
    L.ejoin_upload.restype = C.c_int
    L.ejoin_upload_shard.argtypes = [H, C.POINTER(EjoinPack), C.c_int, C.c_int]
    L.ejoin_upload_shard.restype = C.c_int
    L.ejoin_run_resident.argtypes = [H, C.POINTER(EjoinStats)]
    L.ejoin_run_resident.restype = C.c_int
    L.ejoin_host_alloc.argtypes = [C.c_size_t]
    L.ejoin_host_alloc.restype = C.c_void_p
    L.ejoin_host_free.argtypes = [C.c_void_p]
    L.ejoin_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.ejoin_host_register.restype = C.c_int
    L.ejoin_host_unregister.argtypes = [C.c_void_p]
    L.ejoin_host_unregister.restype = C.c_int
    L.ejoin_partition.argtypes = [C.POINTER(EjoinPack), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint64,
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.ejoin_partition.restype = C.c_int
    L.ejoin_pack_compact.argtypes = [C.POINTER(EjoinPack), C.c_int, C.POINTER(C.POINTER(EjoinPack))]
    L.ejoin_pack_compact.restype = C.c_int
    L.ejoin_pack_compact_free.argtypes = [C.POINTER(EjoinPack)]
    L.ejoin_comm_id.argtypes = [C.c_char_p]
    L.ejoin_comm_id.restype = C.c_int
    L.ejoin_comm_init.argtypes = [H, C.c_char_p, C.c_int, C.c_int]
    L.ejoin_comm_init.restype = C.c_int
    L.ejoin_run_sharded.argtypes = [H, C.POINTER(EjoinPack), C.POINTER(EjoinResult)]
    L.ejoin_run_sharded.restype = C.c_int
    L.ejoin_format_join_log.argtypes = [C.POINTER(EjoinPack), C.POINTER(EjoinParams), C.POINTER(EjoinEdge), C.POINTER(EjoinLogNames), C.c_char_p,
                                        C.c_size_t, C.POINTER(C.c_size_t)]
    L.ejoin_format_join_log.restype = C.c_int
    L.ejoin_set_stage_timing.argtypes = [H, C.c_int]
    L.ejoin_set_stage_timing.restype = C.c_int
    L.ejoin_stage_time.argtypes = [H, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double)]
    L.ejoin_stage_time.restype = C.c_int
    L.ejoin_p1.argtypes = [C.c_uint32] * 3
    L.ejoin_p1.restype = C.c_double
    _lib = L
    return L
