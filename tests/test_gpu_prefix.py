"""Legacy prefix search on the GPU (csrc/prefix.cu) against (a) the golden vectors the unmodified reference produced
(tests/golden/prefix_golden.npz), (b) the CPU restatement oracle/prefix_oracle.py on seeded inputs, edge cases
included, and (c) the command line (`decode --algorithm prefix --window N`, decode.py:179-188).
Bar: labels identical; scores within 1e-6 (north star: 1e-4); gamma within 1e-8."""
import os
import subprocess
import sys
from collections import OrderedDict

import numpy as np
import pytest

from oracle import prefix_oracle as PO
from poreover_b200 import _lib, batch
from poreover_b200.decoding import decoding_cy, prefix_search

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prefix_golden.npz"))
FLAVOURS = ((_lib.PREFIX_NUMPY, "np", "numpy"), (_lib.PREFIX_CY, "cy", "cy"))


def _table(rng, T, S, peaked):
    x = rng.random((T, S)) ** peaked
    x[:, -1] *= 1.5
    x /= x.sum(axis=1, keepdims=True)
    return np.log(x)


def test_1d_search_equals_the_reference_goldens():
    n = int(G["n1d"])
    for S in (3, 5):  # one batch per alphabet size
        ks = [k for k in range(n) if G["y1d_%d" % k].shape[1] == S]
        for fl, tag, _ in FLAVOURS:
            labels, score, st = batch.prefix_search_batch([G["y1d_%d" % k] for k in ks], fl)
            for j, k in enumerate(ks):
                assert labels[j].tolist() == G["lab1d_%s_%d" % (tag, k)].tolist(), (k, tag)
                assert abs(score[j] - float(G["p1d_%s_%d" % (tag, k)])) < 1e-6, (k, tag)
                assert st[j] == 0


def test_gamma_equals_the_reference_goldens():
    for k in range(int(G["n2d"])):
        y1, y2 = G["y2d_a_%d" % k], G["y2d_b_%d" % k]
        for fl, tag, _ in FLAVOURS:
            g = batch.pair_gamma_batch([y1], [y2], fl)[0]
            want = G["gamma_%s_%d" % (tag, k)]
            assert g.shape == want.shape
            assert np.allclose(g, want, rtol=0, atol=1e-8), (k, tag, float(np.abs(g - want).max()))


def test_2d_search_equals_the_reference_goldens():
    n = int(G["n2d"])
    for S in (3, 5):
        ks = [k for k in range(n) if G["y2d_a_%d" % k].shape[1] == S]
        for fl, tag, _ in FLAVOURS:
            labels, score, st = batch.pair_prefix_search_batch([G["y2d_a_%d" % k] for k in ks],
                                                               [G["y2d_b_%d" % k] for k in ks], fl)
            for j, k in enumerate(ks):
                assert labels[j].tolist() == G["lab2d_%s_%d" % (tag, k)].tolist(), (k, tag)
                assert abs(score[j] - float(G["p2d_%s_%d" % (tag, k)])) < 1e-6, (k, tag)
                assert st[j] == 0


def test_searches_equal_the_oracle_on_seeded_tables_and_edge_cases():
    rng = np.random.default_rng(77)
    ys = [_table(rng, int(rng.integers(1, 60)), 5, int(rng.integers(1, 7))) for _ in range(40)]
    # edge cases: a single row, zero probabilities (log 0 = -inf), a deterministic table, one- and three-letter alphabets
    with np.errstate(divide="ignore"):
        hard = np.log(np.array([[0, 0, 0, 0, 1.0], [1.0, 0, 0, 0, 0], [0, 1.0, 0, 0, 0], [0, 0, 0, 0, 1.0]]))
        zeros = np.log(np.array([[0.5, 0, 0.25, 0, 0.25], [0, 0.5, 0, 0.25, 0.25], [0.25, 0.25, 0.25, 0.25, 0]]))
    ys += [ys[0][:1], hard, zeros]
    for fl, _, name in FLAVOURS:
        labels, score, _ = batch.prefix_search_batch(ys, fl)
        for j, y in enumerate(ys):
            with np.errstate(divide="ignore", invalid="ignore"):
                lab, p = PO.prefix_search(y, 4, name)
            assert labels[j].tolist() == lab, (j, name)
            assert score[j] == p or abs(score[j] - p) < 1e-6, (j, name)
    for S in (2, 4):
        small = [_table(rng, int(rng.integers(2, 30)), S, 3) for _ in range(10)]
        for fl, _, name in FLAVOURS:
            labels, score, _ = batch.prefix_search_batch(small, fl)
            for j, y in enumerate(small):
                lab, p = PO.prefix_search(y, S - 1, name)
                assert labels[j].tolist() == lab and abs(score[j] - p) < 1e-6, (S, j, name)
    # pairs of unequal lengths, both flavours, one batch
    p1 = [_table(rng, int(rng.integers(2, 26)), 5, 4) for _ in range(12)]
    p2 = [_table(rng, int(rng.integers(2, 26)), 5, 4) for _ in p1]
    for fl, _, name in FLAVOURS:
        labels, score, _ = batch.pair_prefix_search_batch(p1, p2, fl)
        gam = batch.pair_gamma_batch(p1, p2, fl)
        for j in range(len(p1)):
            lab, p = PO.pair_prefix_search(p1[j], p2[j], 4, name)
            assert labels[j].tolist() == lab and abs(score[j] - p) < 1e-6, (j, name)
            assert np.allclose(gam[j], PO.pair_gamma(p1[j], p2[j], name), rtol=0, atol=1e-8)
    with np.errstate(divide="ignore"):
        det = np.log(np.array([[0, 0, 1.0], [1.0, 0, 0], [0, 1.0, 0]]))  # tests/test_prefix.py:130 of the reference
    toy = OrderedDict([("A", 0), ("B", 1)])
    assert prefix_search.pair_prefix_search_log(det, det, alphabet=toy) == ("AB", 0.0)
    assert prefix_search.pair_prefix_search_log_cy(det, det, alphabet=toy) == ("AB", 0.0)


def test_forward_helpers_equal_the_oracle():
    rng = np.random.default_rng(5)
    y = _table(rng, 17, 5, 2)
    label = [2, 2, 0, 3, 1]
    alpha = prefix_search.forward(label, y)
    alpha_cy = prefix_search.forward(label, y, fw_fn=decoding_cy.forward_vec_log)
    prev = PO.forward_vec_log(-1, 0, y)
    prev_cy = PO.forward_vec_log(-1, 0, y, flavour="cy")
    assert np.allclose(alpha[0], prev, rtol=0, atol=1e-10)
    for i, s in enumerate(label):
        prev = PO.forward_vec_log(s, i + 1, y, previous=prev)
        prev_cy = PO.forward_vec_log(s, i + 1, y, previous=prev_cy, flavour="cy")
        assert np.allclose(alpha[i + 1], prev, rtol=0, atol=1e-10)
        assert np.allclose(alpha_cy[i + 1], prev_cy, rtol=0, atol=1e-10)
    assert alpha_cy[3, 0] == -9999.0 and alpha[3, 0] == -np.inf  # the two flavours' log 0 (decoding_cy.pyx:18)


def test_decode_command_line_prefix_windows(tmp_path):
    """`decode --basecaller poreover --algorithm prefix --window 50`: every window searched on its own, labels joined
    (decode.py:179-188); the loader takes the logarithm of probability tables (decode.py:41-51)."""
    rng = np.random.default_rng(9)
    d = tmp_path / "reads"
    d.mkdir()
    want = {}
    for k, T in enumerate((30, 100, 149, 150, 151)):
        p = np.exp(_table(rng, T, 5, 5)).astype(np.float32)
        p /= p.sum(axis=1, keepdims=True)
        np.save(d / ("w%d.npy" % k), p)
        lp = np.log(p).astype(np.float64)
        seq, i = "", 0
        while i + 50 < T:
            seq += "".join("ACGT"[c] for c in PO.prefix_search(lp[i:i + 50], 4, "cy")[0])
            i += 50
        seq += "".join("ACGT"[c] for c in PO.prefix_search(lp[i:], 4, "cy")[0])
        want["w%d" % k] = seq
    out = tmp_path / "out_prefix"
    cmd = [sys.executable, "-m", "poreover_b200", "decode", str(d), "--basecaller", "poreover", "--out", str(out),
           "--algorithm", "prefix", "--window", "50"]
    subprocess.run(cmd, check=True, env=dict(os.environ, PYTHONPATH=ROOT), cwd=ROOT, capture_output=True)
    recs = {}
    for block in open(str(out) + ".fasta").read().split(">")[1:]:
        name, seq = block.split("\n", 1)
        recs[os.path.splitext(os.path.basename(name.strip()))[0]] = seq.replace("\n", "")
    assert recs == want
