"""Shared comparison of a GPU join result against the CPU oracle."""
import numpy as np

INT_FIELDS = ["k1", "k2", "shares", "indeps", "shares_details", "nrefs", "cd", "cd_chain", "diffs", "k", "d", "n", "err"]
# p1 tolerance: both sides evaluate 1 - sum in double-double, whose cancellation noise is
# ~1e-31 absolute; otherwise they agree to the last bit or two.
P1_ABS, P1_REL = 1e-29, 4.5e-16


def assert_same_join(gpu, oracle_edges, oracle_cls, max_score=100000.0, what=""):
    ge, oe = gpu.edges, oracle_edges
    assert len(ge) == len(oe), f"{what}: {len(ge)} GPU edges vs {len(oe)} oracle edges"
    for f in INT_FIELDS:
        if not np.array_equal(ge[f], oe[f]):
            bad = np.flatnonzero((ge[f] != oe[f]).reshape(len(ge), -1).any(axis=1))[:5]
            raise AssertionError(f"{what}: field {f} differs at edges {bad}: gpu={ge[f][bad]} oracle={oe[f][bad]}")
    assert np.array_equal(gpu.class_of, oracle_cls), f"{what}: partition differs"
    if len(ge):
        # accept_reason: the auto-share and hcomp bits are integer decisions and must agree; the
        # score bit (p1*mult <= max_score) may differ only inside the p1 noise band and only
        # where another bit accepts the join anyway (then it is informational).
        gr, orr = ge["accept_reason"], oe["accept_reason"]
        assert np.array_equal(gr & 6, orr & 6), f"{what}: auto-share / hcomp accept bits differ"
        dif = (gr & 1) != (orr & 1)
        band0 = (P1_ABS + P1_REL * np.abs(oe["p1"])) * oe["mult"]
        assert not (dif & ~((np.abs(oe["score"] - max_score) <= band0) & ((gr & 6) != 0))).any(), f"{what}: score bit differs outside the band"
        # beyond the p1 table depth the library reports 0 (true value < 1e-60, below the
        # reference's own double-double noise floor)
        dp = np.abs(ge["p1"] - oe["p1"])
        assert (dp <= P1_ABS + P1_REL * np.abs(oe["p1"])).all(), f"{what}: p1 differs by {dp.max()}"
        assert np.array_equal(ge["mult"], oe["mult"]), f"{what}: mult differs"
        ds = np.abs(ge["score"] - oe["score"])
        assert (ds <= (P1_ABS + P1_REL * np.abs(oe["p1"])) * oe["mult"] * 1.0000001 + 1e-300).all(), f"{what}: score differs"
        # no decision may sit inside the tolerance band of the score threshold
        band = (P1_ABS + P1_REL * np.abs(oe["p1"])) * oe["mult"]
        only_score = ge["accept_reason"] == 1
        assert not (only_score & (np.abs(oe["score"] - max_score) <= band)).any(), f"{what}: decision inside the tolerance band"
