"""Banded Viterbi acceptor (cpp_viterbi_acceptor, Forward.h:14-121) on the GPU against the oracle restatement and
the unmodified reference function (oracle/_ref), bit-exact paths."""
import numpy as np
import pytest

from oracle import oracle as O
from poreover_b200 import batch, synth
from poreover_b200.decoding import decoding_cpp

pytestmark = pytest.mark.gpu


def _case(seed, T):
    lp = synth.bonito_log_prob(synth.make_read(seed, T)[0])
    return lp, O.beam_search(lp, 25, "ctc")


@pytest.mark.parametrize("seed,T,band", [(1, 300, 1000), (2, 600, 40), (3, 1500, 100), (4, 1000, 12), (5, 2500, 1000),
                                         (6, 97, 7), (7, 5000, 1000)])
def test_acceptor_matches_oracle_and_reference(seed, T, band):
    lp, lab = _case(seed, T)
    want = O.viterbi_acceptor(lp, lab, band, "port")
    if O.have_ref():
        assert np.array_equal(want, O.viterbi_acceptor(lp, lab, band, "ref"))
    got = decoding_cpp.cpp_viterbi_acceptor(lp, lab, band_size=band)
    assert got.dtype.kind == "i" and np.array_equal(got, want)
    # every base is placed once, in order
    assert "".join("ACGT"[i] for i in got[got != 4]) == lab


def test_acceptor_batch_views_and_dtypes():
    """Ragged batch, float64 input, and the reverse-complement VIEW of a read against the materialised array."""
    reads, labels, wants = [], [], []
    for seed, T in ((11, 400), (12, 1234), (13, 64)):
        lp, lab = _case(seed, T)
        reads.append(lp); labels.append(lab); wants.append(O.viterbi_acceptor(lp, lab, 50, "port"))
    paths, st = batch.viterbi_acceptor_batch(reads, labels, 50)
    assert all(np.array_equal(p, w) for p, w in zip(paths, wants)) and not st.any()
    paths64, _ = batch.viterbi_acceptor_batch([r.astype(np.float64) for r in reads], labels, 50)
    assert all(np.array_equal(p, w) for p, w in zip(paths64, wants))
    lp = reads[1]
    rc = np.ascontiguousarray(O.reverse_complement(lp, "bonito"))
    lab_rc = O.beam_search(rc, 25, "ctc")
    want_rc = O.viterbi_acceptor(rc, lab_rc, 50, "port")
    got_rc, _ = batch.viterbi_acceptor_batch([lp], [lab_rc], 50, rc=np.ones(1, np.uint8))
    assert np.array_equal(got_rc[0], want_rc)


def test_acceptor_unplaceable_label_is_flagged():
    """A label that cannot fit inside the band: the reference's traceback never terminates; here it is a status."""
    lp, lab = _case(21, 400)
    short = lab[:40]  # T/L = 10 with a 1-wide band: row L is stored for t <= T-9 only, (L, T-1) never exists
    with pytest.raises(RuntimeError):
        O.viterbi_acceptor(lp, short, 1, "port")
    _, st = batch.viterbi_acceptor_batch([lp], [short], 1)
    assert st[0] & batch._lib.ST_UNSET_BAND
    with pytest.raises(RuntimeError):
        decoding_cpp.cpp_viterbi_acceptor(lp, short, band_size=1)


@pytest.mark.parametrize("T,lab,band", [(1, "A", 1000), (5, "ACGTA", 1000), (5, "ACGTA", 1), (6, "ACGTAC", 2),
                                        (8, "AC", 3), (30, "ACGTTGCA", 2), (30, "A", 1000), (7, "AAAAAAA", 1000),
                                        (64, "ACGT" * 16, 5), (33, "T" * 5, 40)])
def test_acceptor_edge_shapes(T, lab, band):
    """Tiny reads, one label, a base on every timestep, bands wider than the read and 1-wide bands: oracle, the
    unmodified reference function and the kernel agree entry by entry."""
    rng = np.random.default_rng(T * 131 + len(lab))
    lp = np.log(rng.dirichlet(np.ones(5) * 0.5, size=T).astype(np.float32))
    want = O.viterbi_acceptor(lp, lab, band, "port")
    if O.have_ref():
        assert np.array_equal(want, O.viterbi_acceptor(lp, lab, band, "ref"))
    got, st = batch.viterbi_acceptor_batch([lp], [lab], band)
    assert not st[0] and np.array_equal(got[0], want)


def test_acceptor_rejects_more_bases_than_timesteps():
    lp = np.log(np.full((4, 5), 0.2, dtype=np.float32))
    with pytest.raises(Exception):
        batch.viterbi_acceptor_batch([lp], ["ACGTACGT"], 10)
