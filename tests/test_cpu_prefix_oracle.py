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
