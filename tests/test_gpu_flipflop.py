"""GPU parity (bit-exact): flip-flop Viterbi, FP64 8-state DP."""
import numpy as np
import pytest

from poreover_b200 import batch, synth
from poreover_b200.decoding import transducer

pytestmark = pytest.mark.gpu


def test_golden(golden, oracle):
    for k in (0, 1):
        tr = golden["ff%d_trace" % k]
        seqs, maps, paths = batch.flipflop_viterbi_batch([tr], return_path=True)
        assert seqs[0] == str(golden["ff%d_seq" % k])
        assert np.array_equal(paths[0], golden["ff%d_path" % k])
        assert np.array_equal(maps[0], golden["ff%d_s2s" % k])
        lp = synth.flipflop_log_prob(tr)
        m = transducer.flipflop(lp)
        seq, path = m.viterbi_decode(return_path=True)
        assert seq == str(golden["ff%d_seq" % k]) and np.array_equal(path, golden["ff%d_path" % k])


def test_batch_vs_oracle(oracle):
    traces = [synth.make_flipflop_trace(40 + i, T) for i, T in enumerate((1, 2, 9, 33, 257, 1000, 5000))]
    rc = np.array([0, 1, 0, 1, 1, 0, 1], dtype=np.uint8)
    seqs, maps, paths = batch.flipflop_viterbi_batch(traces, rc=rc, return_path=True)
    for tr, r, sq, mp, pa in zip(traces, rc, seqs, maps, paths):
        lp = synth.flipflop_log_prob(tr)
        if r:
            lp = oracle.reverse_complement(lp, "flipflop")
        w_seq, w_path = oracle.viterbi(lp, "flipflop")
        assert sq == w_seq and np.array_equal(pa, w_path)
        assert np.array_equal(mp, oracle.sequence_mapping(w_path, "flipflop"))
    # float64 input path (csv traces, decode.py:83-88) gives the same answer as the uint8 table path
    lps = [synth.flipflop_log_prob(t) for t in traces]
    seqs2, _, _ = batch.flipflop_viterbi_batch(lps, rc=rc)
    assert seqs2 == seqs
