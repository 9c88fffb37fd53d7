"""GPU parity (bit-exact): banded NW + traceback, alignment columns, envelope."""
import numpy as np
import pytest

from poreover_b200 import batch
from poreover_b200 import align as galign
from poreover_b200.decoding import envelope as genv

pytestmark = pytest.mark.gpu


def s(x):
    return str(x)


def test_golden_alignments(golden, oracle):
    n = int(golden["aln_n"])
    for i in range(n):
        a, b, band = [s(x) for x in golden["aln%d_in" % i]]
        al = galign.global_pair_banded(a, b, int(band))
        want = golden["aln%d_out" % i]
        assert "".join(al[0]) == s(want[0]) and "".join(al[1]) == s(want[1]), i
        cols = genv.get_alignment_columns(np.array([al[0], al[1]]))
        assert [("mid".index(c[0]), c[1], c[2]) for c in cols] == [tuple(r) for r in golden["aln%d_cols" % i].tolist()]
        U, V = [int(x) for x in golden["aln%d_UV" % i]]
        for pad in (5, 150):
            env = genv.build_envelope(np.zeros((U, 5)), np.zeros((V, 5)), cols, golden["aln%d_s2s1" % i],
                                      golden["aln%d_s2s2" % i], padding=pad)
            assert np.array_equal(env, golden["aln%d_env_pad%d" % (i, pad)]), (i, pad)
    # an empty sequence: the unmodified reference (oracle/_ref/align*.so, checked in the build container) returns the
    # all-gap alignment -- its row loop, and the division in it, never run (align.pyx:120-171)
    assert galign.global_pair_banded("", "ACGT") == (["-"] * 4, list("ACGT"))
    assert galign.global_pair_banded("ACGT", "") == (list("ACGT"), ["-"] * 4)
    assert galign.global_pair_banded("", "") == ([], [])


def _mutate(rng, a, rate):
    out = []
    for ch in a:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append("ACGT"[rng.integers(0, 4)])
        elif r < rate:
            out.append(ch)
            out.append("ACGT"[rng.integers(0, 4)])
        else:
            out.append(ch)
    return "".join(out) or "A"


def test_fuzz_batch_vs_oracle(oracle):
    rng = np.random.default_rng(11)
    s1, s2 = [], []
    for k in range(120):
        n = int(rng.integers(1, 400))
        a = "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
        b = _mutate(rng, a, rng.uniform(0, 0.6))
        if k % 7 == 0:
            b = b[: max(1, len(b) // 3)]
        s1.append(a)
        s2.append(b)
    for band, sc in ((500, (2, -1, -1)), (10, (2, -1, -1)), (3, (2, -1, -1)), (1, (2, -1, -1)), (25, (3, -2, -2))):
        got = batch.align_banded_batch(s1, s2, band, *sc)
        for a, b, g in zip(s1, s2, got):
            w = oracle.global_pair_banded(a, b, band, *sc)
            assert g[0] == "".join(w[0]) and g[1] == "".join(w[1]), (band, a, b)
            assert g[2] == sum(x == y for x, y in zip(w[0], w[1]))


def test_long_pair_and_envelope(golden, oracle):
    """A ~2000-base pair with the default band, then the envelope from the real mappings (pair fixture)."""
    for i in range(int(golden["pair_n"])):
        b1, b2 = s(golden["pair%d_basecall1" % i]), s(golden["pair%d_basecall2" % i])
        al = galign.global_pair_banded(b1, b2)
        want = golden["pair%d_align" % i]
        assert "".join(al[0]) == s(want[0]) and "".join(al[1]) == s(want[1])
        lp1, lp2 = golden["pair%d_lp1" % i], golden["pair%d_lp2_rc" % i]
        _, maps, _, _ = batch.viterbi_batch([lp1, lp2], "bonito")
        cols = genv.get_alignment_columns(np.array([al[0], al[1]]))
        env = genv.build_envelope(lp1, lp2, cols, maps[0], maps[1], padding=5)
        assert np.array_equal(env, golden["pair%d_env" % i])
    rng = np.random.default_rng(3)
    a = "".join("ACGT"[i] for i in rng.integers(0, 4, size=2100))
    b = _mutate(rng, a, 0.12)
    g = batch.align_banded_batch([a, b], [b, a])
    for (x, y), r in zip(((a, b), (b, a)), g):
        w = oracle.global_pair_banded(x, y)
        assert r[0] == "".join(w[0]) and r[1] == "".join(w[1])
