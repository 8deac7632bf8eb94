"""The join's log block (JoinInfo.log) rebuilt from one edge record: byte-for-byte equal to the reference's golden output
(enclone_exec/testx/inputs/outputs/enclone_test256_output:29-70, ANSI stripped; committed as tests/golden/kat256_join_log.txt
by tests/golden/make_kat256.py).  Pins share_pos_v, the per-region differences and identities, the printed precision of
score / p1 / mult, and the difference patterns of both chains of the real flaky8a/flaky8b contigs."""
import os

import pytest

import oracle_lib as O
from enclone_b200._abi import default_params
from enclone_b200.join import format_join_log
from kat import kat_pack

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = dict(index=0, id1="flaky8b.clono1", id2="flaky8a.clono0", vsegs1="107=ENST00000632630, 219=ENST00000632910",
             vsegs2="107=ENST00000632630, 219=ENST00000632910", dataset1="flaky8b = flaky8b", dataset2="flaky8a = flaky8a",
             cdr3_aa1="IGK:CMQSTHWPTWTF;IGH:CAFDFW", cdr3_aa2="IGK:CMQTTHWPTWTL;IGH:CAFDFW", bcs1="AATCCAGGTACAGACG-1", bcs2="ACGAGCCGTCGGGTCT-1")


def _golden():
    return open(os.path.join(HERE, "golden", "kat256_join_log.txt"), encoding="utf-8").read()


def test_join_log_from_the_oracle_edge_matches_the_golden_block():
    pack, _ = kat_pack()
    params = default_params(mix_donors=1)
    ok, edge = O.join_one(pack.struct, params, 0, 1)
    assert ok
    got = format_join_log(pack.struct, params, edge, **NAMES)
    want = _golden()
    assert got.split("\n") == want.split("\n")
    assert got == want


@pytest.mark.gpu
def test_join_log_from_the_gpu_edge_matches_the_golden_block():
    from enclone_b200.join import CompactPack, join_exacts
    pack, _ = kat_pack()
    params = default_params(mix_donors=1)
    for src in (pack, CompactPack(pack)):                     # ASCII and compact forms of the input
        g = join_exacts(src, params)
        assert len(g.edges) == 1
        assert format_join_log(src.struct, params, g.edges[0], **NAMES) == _golden()
