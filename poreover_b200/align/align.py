"""poreover/align/align.pyx replacement."""
from .. import batch

MATCH_DEFAULT, MISMATCH_DEFAULT, GAP_DEFAULT, BAND_DEFAULT = 2, -1, -1, 500  # align.pyx:10-13


def global_pair_banded(seq1, seq2, band_width=BAND_DEFAULT, match=MATCH_DEFAULT, mismatch=MISMATCH_DEFAULT,
                       gap_cost=GAP_DEFAULT):
    """Banded Needleman-Wunsch with constant gap penalty (align.pyx:100-178).

    Returns (align1, align2): two equal-length lists of single characters, '-' for gaps."""
    r = batch.align_banded_batch([seq1], [seq2], band_width, match, mismatch, gap_cost)[0]
    return list(r[0]), list(r[1])


def global_pair(seq1, seq2, match=MATCH_DEFAULT, mismatch=MISMATCH_DEFAULT, gap_cost=GAP_DEFAULT):
    """Full Needleman-Wunsch with constant gap penalty (align.pyx:29-98), the `--alignment full` path.

    Returns (align1, align2, dpMatrix) like the reference: two lists of characters and the int DP matrix."""
    r = batch.align_global_batch([seq1], [seq2], match, mismatch, gap_cost, return_dp=True)[0]
    return list(r[0]), list(r[1]), r[3]
