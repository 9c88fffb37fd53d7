"""Generates tests/golden/prefix_golden.npz by running the UNMODIFIED reference package (oracle/_ref/refpkg.zip + its
compiled Cython helpers) on seeded inputs: prefix_search_log, prefix_search_log_cy, pair_gamma_log,
decoding_cy.pair_gamma_log, pair_prefix_search_log and pair_prefix_search_log_cy.  Run in the build container (needs
oracle/_ref, i.e. /root/reference at build time): python tests/golden/make_prefix_golden.py"""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_python_driver  # noqa: E402

ref_python_driver.load_reference_package()
from poreover.decoding import decoding_cy, prefix_search  # noqa: E402


def table(rng, T, S, peaked):
    """a random probability table (rows sum to 1, blank last), log domain; `peaked` sharpens the rows"""
    x = rng.random((T, S)) ** peaked
    x[:, -1] *= 1.5
    x /= x.sum(axis=1, keepdims=True)
    return np.log(x)


def main():
    rng = np.random.default_rng(20250)
    out = {}
    alph = {2: OrderedDict([("A", 0), ("B", 1)]), 4: prefix_search.DNA_alphabet}
    k = 0
    for T, S, peaked in [(4, 3, 1), (9, 3, 3), (12, 5, 2), (40, 5, 4), (120, 5, 6), (400, 5, 8)]:
        for rep in range(3):
            y = table(rng, T, S, peaked)
            a = alph[S - 1]
            l0, p0 = prefix_search.prefix_search_log(y, alphabet=a)
            l1, p1 = prefix_search.prefix_search_log_cy(y, alphabet=a)
            out["y1d_%d" % k] = y
            out["lab1d_np_%d" % k] = np.array([a[c] for c in l0], dtype=np.uint8)
            out["p1d_np_%d" % k] = p0
            out["lab1d_cy_%d" % k] = np.array([a[c] for c in l1], dtype=np.uint8)
            out["p1d_cy_%d" % k] = p1
            k += 1
    out["n1d"] = k
    k = 0
    for U, V, S, peaked in [(4, 4, 3, 1), (5, 7, 3, 2), (10, 9, 5, 3), (30, 34, 5, 5), (60, 55, 5, 8)]:
        for rep in range(2):
            # two noisy views of the same sharpened table so that the reads agree on something
            base = table(rng, max(U, V), S, peaked)
            y1 = np.log(np.exp(base[:U]) * 0.8 + 0.2 * np.exp(table(rng, U, S, 1)))
            y2 = np.log(np.exp(base[:V]) * 0.8 + 0.2 * np.exp(table(rng, V, S, 1)))
            a = alph[S - 1]
            out["y2d_a_%d" % k] = y1
            out["y2d_b_%d" % k] = y2
            out["gamma_np_%d" % k] = prefix_search.pair_gamma_log(y1, y2)
            out["gamma_cy_%d" % k] = np.asarray(decoding_cy.pair_gamma_log(y1, y2))
            l0, p0 = prefix_search.pair_prefix_search_log(y1, y2, alphabet=a)
            l1, p1 = prefix_search.pair_prefix_search_log_cy(y1, y2, alphabet=a)
            out["lab2d_np_%d" % k] = np.array([a[c] for c in l0], dtype=np.uint8)
            out["p2d_np_%d" % k] = p0
            out["lab2d_cy_%d" % k] = np.array([a[c] for c in l1], dtype=np.uint8)
            out["p2d_cy_%d" % k] = p1
            k += 1
    out["n2d"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "prefix_golden.npz"), **out)
    print("1D cases", out["n1d"], "2D cases", out["n2d"], "label lengths 1D",
          [len(out["lab1d_np_%d" % i]) for i in range(out["n1d"])], "2D", [len(out["lab2d_np_%d" % i]) for i in range(out["n2d"])])


if __name__ == "__main__":
    main()
