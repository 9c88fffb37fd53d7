"""GPU parity at BASELINE.json sizes: T ~ 5000 pairs against the unmodified reference core (oracle/_ref, else
the oracle port), plus size-independent properties on larger batches (determinism, batch-position
invariance, the reference's own 2D(y,y) == 1D(y) identity on a diagonal envelope)."""
import numpy as np
import pytest

from poreover_b200 import _lib, batch, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pairs(first, n, T):
    l1, l2 = [], []
    for k in range(first, first + n):
        p1, p2, _ = synth.make_pair(k, T)
        l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
    return l1, l2


def test_pairs_T5000_vs_reference(oracle):
    backend = "ref" if oracle.have_ref() else "port"
    l1, l2 = _pairs(100, 3, 5000)
    res = batch.pair_decode_batch(l1, l2, "bonito", beam_width=25, rc2=True)
    for a, b, r in zip(l1, l2, res):
        want = oracle.pair_decode(a, oracle.reverse_complement(b, "bonito"), "bonito", 25, backend=backend, with_score=True)
        assert r["basecall1"] == want["basecall1"] and r["basecall2"] == want["basecall2"]
        assert r["identity"] == want["identity"]
        assert r["consensus"] == want["consensus"]
        assert abs(r["score"] - want["score"]) < TOL
        assert not (r["status"] & (_lib.ST_POOL_OVERFLOW | _lib.ST_SHORT_BEAM_SKIP | _lib.ST_UNSET_BAND))


def test_single_reads_T5000_beam(oracle):
    backend = "ref" if oracle.have_ref() else "port"
    arrays = [synth.bonito_log_prob(synth.make_read(700 + i, 5000)[0]) for i in range(2)]
    for W in (25, 100):
        seqs, sc, st = batch.beam_search_batch(arrays, W, "ctc_merge_repeats")
        for a, g, gs in zip(arrays, seqs, sc):
            w, ws = oracle.beam_search(a, W, "ctc_merge_repeats", backend, True)
            assert g == w and abs(gs - ws) < TOL
        assert not (st & _lib.ST_POOL_OVERFLOW).any()


def test_batch_properties_200_pairs():
    """Determinism and independence of batch position / neighbours at a size the oracle could not finish."""
    l1, l2 = _pairs(300, 40, 3000)
    big1, big2 = l1 * 5, l2 * 5
    r1 = batch.pair_decode_batch(big1, big2, "bonito", 25, rc2=True)
    perm = np.random.default_rng(1).permutation(len(big1))
    r2 = batch.pair_decode_batch([big1[i] for i in perm], [big2[i] for i in perm], "bonito", 25, rc2=True)
    for j, i in enumerate(perm):
        assert r1[i]["consensus"] == r2[j]["consensus"] and r1[i]["score"] == r2[j]["score"]
    for i in range(40):
        for rep in range(1, 5):
            assert r1[i]["consensus"] == r1[i + 40 * rep]["consensus"] and r1[i]["score"] == r1[i + 40 * rep]["score"]
    assert all(not (r["status"] & _lib.ST_POOL_OVERFLOW) for r in r1)
    # consensus is at least as close to either 1D basecall's length as they are to each other (sanity)
    assert all(abs(len(r["consensus"]) - r["length1"]) < 0.2 * r["length1"] for r in r1)


def test_diagonal_envelope_equals_1d():
    """tests/test_beam.py::beam_2d_same::test_diagonal_envelope of the reference, at T = 5000."""
    y = synth.bonito_log_prob(synth.make_read(900, 5000)[0])
    T = len(y)
    env = np.array([(i, i + 1) for i in range(T)])
    s1, _, _ = batch.beam_search_batch([y], 25, "ctc")
    s2, _, _ = batch.beam_search_2d_batch([y], [y], [env], 25, "ctc", "row")
    assert s1[0] == s2[0]


def test_long_pair_wide_band(oracle):
    """Config-4 style stress: a long pair whose envelope is widened (padding 150) so bands exceed 300 steps."""
    p1, p2, _ = synth.make_pair(5000, 12000)
    lp1 = synth.bonito_log_prob(p1)
    lp2 = np.ascontiguousarray(oracle.reverse_complement(synth.bonito_log_prob(p2), "bonito"))
    want = oracle.pair_decode(lp1, lp2, "bonito", 5, padding=150, with_score=True)
    env = want["envelope"]
    assert (env[:, 1] - env[:, 0]).max() > 300
    seqs, sc, st = batch.beam_search_2d_batch([lp1], [lp2], [env], 5, "ctc_merge_repeats", "row_col")
    assert seqs[0] == want["consensus"] and abs(sc[0] - want["score"]) < TOL
    assert not (st[0] & _lib.ST_POOL_OVERFLOW)


def test_real_pair_from_reference_data(tmp_path):
    """The reference's own real-data pair (data/reads/read1.npy + read2.npy, PoreOverNet logits, 62,000 and
    75,600 timesteps, bands up to ~1450 wide) through the fused pipeline; expectations recorded from the real
    reference by tests/golden/make_golden_real.py."""
    import os
    from poreover_b200.decoding import decode as gdecode
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_pair.npz"))
    arrays = []
    for name in ("read1", "read2"):
        f = tmp_path / (name + ".npy")
        np.save(f, g[name])
        arrays.append(gdecode.model_from_trace(str(f), "poreover").device_array())
    assert arrays[0].shape == (62000, 5) and arrays[1].shape == (75600, 5)
    for W in (5, 25):
        r = batch.pair_decode_batch([arrays[0]], [arrays[1]], "poreover", beam_width=W, rc2=True)[0]
        assert r["basecall1"] == str(g["basecall1"])
        assert (r["length1"], r["length2"]) == (int(g["length1"]), int(g["length2"]))
        assert r["identity"] == float(g["identity"])
        assert r["consensus"] == str(g["consensus_w%d" % W]), W
        assert not (r["status"] & (_lib.ST_POOL_OVERFLOW | _lib.ST_SHORT_BEAM_SKIP | _lib.ST_UNSET_BAND))
