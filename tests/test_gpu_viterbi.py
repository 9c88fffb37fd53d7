"""GPU parity: CUDA best-path decode + sequence mapping vs the oracle and the golden fixtures (bit-exact)."""
import numpy as np
import pytest

from poreover_b200 import _lib, batch, synth
from poreover_b200.decoding import transducer

pytestmark = pytest.mark.gpu


def _check(O, arrays, kind, rc=None, layout=_lib.BLANK_LAST, oracle_arrays=None):
    seqs, maps, paths, st = batch.viterbi_batch(arrays, kind, rc=rc, layout=layout, return_path=True)
    oracle_arrays = oracle_arrays or arrays
    for i, a in enumerate(oracle_arrays):
        want_seq, want_path = O.viterbi(a, kind)
        assert seqs[i] == want_seq
        assert np.array_equal(paths[i], want_path)
        want_map = O.sequence_mapping(want_path, kind)
        got = maps[i][1:] if st[i] & _lib.ST_MAPPING_WRAP else maps[i]
        assert np.array_equal(got, want_map)


def test_golden(golden, oracle):
    csv = golden["csv_log_prob"]
    seqs, maps, paths, st = batch.viterbi_batch([csv], "poreover", return_path=True)
    assert seqs[0] == str(golden["csv_viterbi_seq"])
    assert np.array_equal(paths[0], golden["csv_viterbi_path"])
    assert np.array_equal(maps[0], golden["csv_s2s"])
    for k in (0, 1):
        lp = golden["syn%d_log_prob" % k]
        seqs, maps, paths, st = batch.viterbi_batch([lp, lp], "bonito", rc=[0, 1], return_path=True)
        assert seqs[0] == str(golden["syn%d_viterbi_seq" % k])
        assert seqs[1] == str(golden["syn%d_rc_viterbi_seq" % k])
        assert np.array_equal(paths[0], golden["syn%d_viterbi_path" % k])
        assert np.array_equal(maps[0], golden["syn%d_s2s" % k])
        # file layout (blank first) handled on the device == loader permutation on the host
        prob = golden["syn%d_prob" % k]
        with np.errstate(divide="ignore"):
            raw = np.log(prob)
        s2, _, _, _ = batch.viterbi_batch([raw], "bonito", layout=_lib.BLANK_FIRST)
        assert s2[0] == str(golden["syn%d_viterbi_seq" % k])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ragged_batch(oracle, dtype):
    rng = np.random.default_rng(0)
    lens = [1, 2, 3, 4, 5, 7, 31, 32, 33, 127, 128, 129, 500, 1023, 4097]
    arrays = []
    for i, T in enumerate(lens):
        p, _ = synth.make_read(100 + i, T)
        arrays.append(synth.bonito_log_prob(p).astype(dtype))
    rc = rng.integers(0, 2, size=len(arrays)).astype(np.uint8)
    ora = [oracle.reverse_complement(a, "bonito") if r else a for a, r in zip(arrays, rc)]
    for kind in ("bonito", "poreover"):
        _check(oracle, arrays, kind, rc=rc, oracle_arrays=ora)


def test_ties_and_wrap(oracle):
    # exact ties: first index wins (numpy argmax); a base tied with blank resolves to the base
    y = np.log(np.array([[0.2, 0.2, 0.2, 0.2, 0.2], [0.1, 0.4, 0.05, 0.05, 0.4], [0.3, 0.3, 0.1, 0.0, 0.3],
                         [0.0, 0.0, 0.0, 0.0, 1.0], [0.25, 0.25, 0.25, 0.25, 0.0]], dtype=np.float32) + 0.0)
    _check(oracle, [y], "bonito")
    _check(oracle, [y], "poreover")
    # first and last frame carry the same base: the reference mapping drops the first base (A.8-Q4)
    z = np.full((6, 5), -5.0, dtype=np.float32)
    for t, k in enumerate([2, 4, 1, 1, 4, 2]):
        z[t, k] = -0.1
    seqs, maps, paths, st = batch.viterbi_batch([z], "bonito", return_path=True)
    assert seqs[0] == "GCG" and (st[0] & _lib.ST_MAPPING_WRAP)
    assert np.array_equal(maps[0][1:], oracle.sequence_mapping(paths[0], "bonito"))


def test_small_alphabet_and_transducer_api(golden, oracle):
    y = np.log(golden["toy0_y"])
    m = transducer.poreover(y, alphabet="AB")
    assert m.viterbi_decode().translate(str.maketrans("AC", "AB")) == str(golden["toy0_viterbi"])
    lp = golden["syn0_log_prob"]
    m = transducer.bonito(lp)
    seq, path = m.viterbi_decode(return_path=True)
    assert seq == str(golden["syn0_viterbi_seq"]) and np.array_equal(path, golden["syn0_viterbi_path"])
    assert np.array_equal(m.sequence_mapping(), golden["syn0_s2s"])
    m.reverse_complement()
    assert m.viterbi_decode() == str(golden["syn0_rc_viterbi_seq"])


def test_large_roundtrip_property():
    """Full bench size property: decoding a planted noiseless matrix returns the planted (collapsed) sequence."""
    rng = np.random.default_rng(5)
    T, n = 5000, 64
    arrays, want = [], []
    for _ in range(n):
        sym = rng.integers(0, 5, size=T)
        y = np.full((T, 5), -4.0, dtype=np.float32)
        y[np.arange(T), sym] = -0.05
        arrays.append(y)
        keep = (sym != 4) & (np.concatenate(([True], sym[1:] != sym[:-1])))
        want.append("".join("ACGT"[s] for s in sym[keep]))
    seqs, maps, _, _ = batch.viterbi_batch(arrays, "bonito")
    assert seqs == want
    assert all(len(m) == len(s) for m, s in zip(maps, seqs))
