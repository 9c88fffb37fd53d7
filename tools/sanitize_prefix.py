"""Small cases of the legacy prefix-search kernels for compute-sanitizer (memcheck / racecheck / synccheck), checked
against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import prefix_oracle as PO
from poreover_b200 import _lib, batch

rng = np.random.default_rng(1)


def table(T, S=5, peaked=4):
    x = rng.random((T, S)) ** peaked
    x /= x.sum(axis=1, keepdims=True)
    return np.log(x)


bad = 0
ys = [table(T) for T in (1, 5, 17, 33)] + [table(9, 3)] * 0
for fl, name in ((_lib.PREFIX_NUMPY, "numpy"), (_lib.PREFIX_CY, "cy")):
    labs, sc, st = batch.prefix_search_batch(ys, fl)
    for y, l, s in zip(ys, labs, sc):
        want, p = PO.prefix_search(y, 4, name)
        bad += (l.tolist() != want) or abs(s - p) > 1e-6
    p1 = [table(6), table(11), table(3, 3)][:2]
    p2 = [table(7), table(9)]
    labs, sc, st = batch.pair_prefix_search_batch(p1, p2, fl)
    gam = batch.pair_gamma_batch(p1, p2, fl)
    for a, b, l, s, g in zip(p1, p2, labs, sc, gam):
        want, p = PO.pair_prefix_search(a, b, 4, name)
        bad += (l.tolist() != want) or abs(s - p) > 1e-6 or not np.allclose(g, PO.pair_gamma(a, b, name), atol=1e-8, rtol=0)
    v = batch.forward_vec(ys[2], 1, 1, PO.forward_vec_log(-1, 0, ys[2]), fl)
    bad += not np.allclose(v, PO.forward_vec_log(1, 1, ys[2], PO.forward_vec_log(-1, 0, ys[2]), name), atol=1e-10, rtol=0)
print("done, mismatches:", int(bad))
