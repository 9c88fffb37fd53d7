"""The legacy prefix search's CPU restatement (oracle/prefix_oracle.py) against the golden vectors the unmodified
reference produced (tests/golden/prefix_golden.npz, tests/golden/make_prefix_golden.py)."""
import os

import numpy as np

from oracle import prefix_oracle as PO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prefix_golden.npz"))


def test_oracle_1d_matches_the_reference_goldens():
    for k in range(int(G["n1d"])):
        y = G["y1d_%d" % k]
        if len(y) > 130:
            continue  # the pure-Python loops take seconds per case beyond this; the GPU test covers T = 400
        for fl, tag in (("numpy", "np"), ("cy", "cy")):
            lab, p = PO.prefix_search(y, y.shape[1] - 1, fl)
            assert lab == G["lab1d_%s_%d" % (tag, k)].tolist(), (k, fl)
            assert abs(p - float(G["p1d_%s_%d" % (tag, k)])) < 1e-9, (k, fl)


def test_oracle_gamma_and_2d_match_the_reference_goldens():
    for k in range(int(G["n2d"])):
        y1, y2 = G["y2d_a_%d" % k], G["y2d_b_%d" % k]
        if len(y1) > 40:
            continue
        for fl, tag in (("numpy", "np"), ("cy", "cy")):
            g = PO.pair_gamma(y1, y2, fl)
            assert np.allclose(g, G["gamma_%s_%d" % (tag, k)], rtol=0, atol=1e-9), (k, fl)
            lab, p = PO.pair_prefix_search(y1, y2, y1.shape[1] - 1, fl)
            assert lab == G["lab2d_%s_%d" % (tag, k)].tolist(), (k, fl)
            assert abs(p - float(G["p2d_%s_%d" % (tag, k)])) < 1e-9, (k, fl)


def test_oracle_on_the_reference_tests_own_toy_tables():
    """tests/test_prefix.py:66-85 of the reference: brute-force top labels of three toy tables."""
    def brute(y):
        T, S = y.shape
        best = {}
        import itertools
        for path in itertools.product(range(S), repeat=T):
            lab = tuple(c for c in path if c != S - 1)
            best[lab] = best.get(lab, 0.0) + float(np.prod([y[t, c] for t, c in enumerate(path)]))
        lab = max(best.items(), key=lambda kv: kv[1])
        return list(lab[0]), np.log(lab[1])
    for y in (np.array([[0.1, 0.6, 0.3], [0.4, 0.2, 0.4], [0.4, 0.3, 0.3], [0.2, 0.8, 0]]),
              np.array([[0.7, 0.2, 0.1], [0.2, 0.3, 0.5], [0.7, 0.2, 0.1], [0.05, 0.05, 0.9]]),
              np.array([[0.7, 0.2, 0.1], [0.2, 0.3, 0.5]])):
        with np.errstate(divide="ignore"):
            lab, p = PO.prefix_search(np.log(y), 2, "numpy")
        want, wp = brute(y)
        assert lab == want and abs(p - wp) < 1e-9


def test_prefix_host_logic_window_cuts_and_alphabet_checks(monkeypatch):
    """decode --algorithm prefix cuts a read into windows exactly as decode.py:181-188 does and joins the labels; the
    module rejects alphabets it cannot map onto the kernels.  The GPU call is replaced by the oracle here (host logic
    only: no CUDA in the CPU suite)."""
    from collections import OrderedDict

    import pytest

    from poreover_b200 import batch
    from poreover_b200.decoding import decode, prefix_search, transducer

    calls = []

    def fake(arrays, flavour=1, device=None):
        calls.append([len(a) for a in arrays])
        labs = [np.array(PO.prefix_search(a, a.shape[1] - 1, "cy" if flavour == 1 else "numpy")[0], dtype=np.uint8)
                for a in arrays]
        return labs, np.zeros(len(arrays)), np.zeros(len(arrays), np.int32)

    monkeypatch.setattr(batch, "prefix_search_batch", fake)
    rng = np.random.default_rng(4)

    def table(T):
        x = rng.random((T, 5)) ** 5
        x /= x.sum(axis=1, keepdims=True)
        return np.log(x)

    reads = [table(T) for T in (7, 20, 21, 40)]
    models = [transducer.poreover(r) for r in reads]
    seqs = decode.decode_models(models, "prefix", window=10)
    assert calls == [[7, 10, 10, 10, 10, 1, 10, 10, 10, 10]]  # T = 20: a window and then the empty-free tail (i + window < T)
    for r, s in zip(reads, seqs):
        want, i = "", 0
        while i + 10 < len(r):
            want += "".join("ACGT"[c] for c in PO.prefix_search(r[i:i + 10], 4, "cy")[0])
            i += 10
        want += "".join("ACGT"[c] for c in PO.prefix_search(r[i:], 4, "cy")[0])
        assert s == want
    assert prefix_search.prefix_search_windows(reads[2], 10) == seqs[2]
    y = table(6)
    with pytest.raises(NotImplementedError):
        prefix_search.prefix_search_log(y, alphabet=OrderedDict([("C", 1), ("A", 0), ("G", 2), ("T", 3)]))
    with pytest.raises(ValueError):
        prefix_search.prefix_search_log(y, alphabet=OrderedDict([("A", 0), ("B", 1)]))  # 5 columns, 2 letters
    with pytest.raises(NotImplementedError):
        prefix_search.prefix_search_log(y, return_forward=True)
    assert prefix_search.remove_gaps("A-C--G") == "ACG"
    assert prefix_search.greedy_search(np.log(np.array([[.7, .1, .1, .05, .05], [.1, .1, .1, .1, .6], [.1, .6, .1, .1, .1]]))) == "AC"
