"""enclone_b200 — B200-native clonotype join (enclone's join_exacts / join_core / join_one).

The product is csrc/ -> libenclone_join.so (CUDA, sm_100a) behind include/enclone_join.h.
This package is the host-side mirror of the reference interface for that path.
"""
from ._abi import EJOIN_NONE, EJOIN_NONE16, JoinPack, default_params  # noqa: F401
