"""GPU parity: cpp_forward (label forward log-probability) and global_pair (full NW)."""
import numpy as np
import pytest

from poreover_b200 import batch, synth
from poreover_b200 import align as galign
from poreover_b200.decoding import decoding_cpp

pytestmark = pytest.mark.gpu


def test_forward_golden(golden, oracle):
    y = golden["csv_log_prob"]
    lab = str(golden["csv_beam1d_w25"])
    assert abs(decoding_cpp.cpp_forward(y, lab) - float(golden["csv_forward_w25"])) < 1e-5
    for k in (0, 1):
        lp = golden["syn%d_log_prob" % k]
        seq = str(golden["syn%d_viterbi_seq" % k])
        assert abs(decoding_cpp.cpp_forward(lp, seq, "ACGT", "ctc_merge_repeats") - float(golden["syn%d_forward_bonito" % k])) < 1e-5
        assert abs(decoding_cpp.cpp_forward(lp, seq, "ACGT", "ctc") - float(golden["syn%d_forward_ctc" % k])) < 1e-5


def test_forward_batch_vs_oracle(oracle):
    rng = np.random.default_rng(4)
    arrays, labels = [], []
    for i, T in enumerate((3, 17, 300, 1500)):
        a = synth.bonito_log_prob(synth.make_read(50 + i, T)[0])
        arrays.append(a)
        L = max(1, int(0.3 * T))  # 450 > 256: exercises the tiled path
        labels.append("".join("ACGT"[j] for j in rng.integers(0, 4, size=L)))
    labels[2] = oracle.viterbi(arrays[2], "bonito")[0]
    for model in ("ctc", "ctc_merge_repeats"):
        got = batch.forward_batch(arrays, labels, model)
        for a, l, g in zip(arrays, labels, got):
            w = oracle.forward(a, l, model)
            assert (np.isinf(w) and np.isinf(g)) or abs(g - w) < 1e-5, (model, len(a), g, w)


def test_global_pair(golden, oracle):
    al = galign.global_pair("ACGTTGCAAC", "ACTTGGCAC")
    assert "".join(al[0]) == str(golden["alnfull_out"][0]) and "".join(al[1]) == str(golden["alnfull_out"][1])
    assert np.array_equal(al[2], golden["alnfull_dp"])
    rng = np.random.default_rng(9)
    s1, s2 = [], []
    for k in range(40):
        n = int(rng.integers(1, 300))
        a = "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
        b = "".join(ch for ch in a if rng.random() > 0.1) or "A"
        s1.append(a); s2.append(b)
    got = batch.align_global_batch(s1, s2, return_dp=True)
    for a, b, g in zip(s1, s2, got):
        w = oracle.global_pair(a, b)
        assert g[0] == "".join(w[0]) and g[1] == "".join(w[1])
        assert np.array_equal(g[3], w[2])
