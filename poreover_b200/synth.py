"""Synthetic Bonito-shaped CTC probability matrices with a planted sequence.

Files written by ``save_pair`` are exactly what ``data/bonito022.patch:7-10`` of the reference makes
bonito emit: float32 probabilities, shape (T, 5), blank in column 0, so the reference loader
(decode.py:41-51, :76-80) and ours accept them unchanged.
"""
import os

import numpy as np

_COMP = np.array([3, 2, 1, 0])


def _lay(rng, seq, T, head_blank, tail_blank, blank_dom_head=0, blank_dom_tail=0):
    """Probabilities (T,5) float32, blank FIRST, for planted base indices ``seq`` (0..3)."""
    L = len(seq)
    lo, hi = max(1, head_blank), T - max(1, tail_blank)
    L = min(L, max(hi - lo, 0))
    pos = np.sort(rng.choice(np.arange(lo, hi), size=L, replace=False)) if L else np.zeros(0, dtype=np.int64)
    sym = np.zeros(T, dtype=np.int64)  # 0 = blank, 1..4 = bases (file layout)
    sym[pos] = np.asarray(seq[:L]) + 1
    c = rng.uniform(0.5, 0.99, size=T)
    if blank_dom_head:
        c[:blank_dom_head] = 0.96
    if blank_dom_tail:
        c[T - blank_dom_tail:] = 0.96
    noise = rng.standard_gamma(0.3, size=(T, 5)) + 1e-12
    noise /= noise.sum(axis=1, keepdims=True)
    p = noise * (1.0 - c)[:, None]
    p[np.arange(T), sym] += c
    p = np.maximum(p, 1e-7)
    p /= p.sum(axis=1, keepdims=True)
    return p.astype(np.float32)


def make_read(seed, T=5000, density=0.4):
    """One read: (probs float32 (T,5) blank-first, planted base indices)."""
    rng = np.random.default_rng(seed)
    L = int(round(density * T))
    seq = rng.integers(0, 4, size=L)
    return _lay(rng, seq, T, 1, 1, blank_dom_head=3), seq


def make_pair(k, T=5000, density=0.4):
    """Pair k: seeds 1000+k (read 1) and 2000+k (read 2).  Read 2 carries the reverse complement of the
    planted sequence with independent timing and noise and T2 ~ 1.03-1.05 T1.  Rule 1: first 3 frames of
    read 1 and last 3 frames of read 2 are blank-dominant; rule 2: first/last frame of each read blank."""
    r1 = np.random.default_rng(1000 + k)
    r2 = np.random.default_rng(2000 + k)
    L = int(round(density * T))
    seq = r1.integers(0, 4, size=L)
    p1 = _lay(r1, seq, T, 3, 1, blank_dom_head=3)
    T2 = int(round(T * r2.uniform(1.03, 1.05)))
    rc = _COMP[seq[::-1]]
    p2 = _lay(r2, rc, T2, 1, 3, blank_dom_tail=3)
    return p1, p2, seq


def save_pair(dirname, k, T=5000, blank_last=False):
    """Write the two reads of pair k as .npy probability files: bonito order (blank first, data/bonito022.patch:7-10)
    or, with blank_last, the PoreOverNet order that `--basecaller poreover` expects (decode.py:108-110)."""
    p1, p2, _ = make_pair(k, T)
    if blank_last:
        p1, p2 = np.ascontiguousarray(p1[:, [1, 2, 3, 4, 0]]), np.ascontiguousarray(p2[:, [1, 2, 3, 4, 0]])
    f1, f2 = "pair%05d_1.npy" % k, "pair%05d_2.npy" % k
    np.save(os.path.join(dirname, f1), p1)
    np.save(os.path.join(dirname, f2), p2)
    return f1, f2


def bonito_log_prob(p):
    """What decode.model_from_trace(path, 'bonito') yields for probabilities p (decode.py:41-51, :76-80):
    float32 log, blank moved last.  (transducer.__init__ then widens to float64, exactly.)"""
    with np.errstate(divide="ignore"):
        return np.log(p)[:, [1, 2, 3, 4, 0]]


def make_flipflop_trace(seed, T=5000):
    """Synthetic Guppy/Flappie style T x 8 uint8 trace (decode.py:53-65, :89-104)."""
    rng = np.random.default_rng(seed)
    state = np.zeros(T, dtype=np.int64)
    s = int(rng.integers(0, 4))
    for t in range(T):
        if rng.random() < 0.4:
            b = int(rng.integers(0, 4))
            s = (b + 4) if (b == (s & 3) and s < 4) else b
        state[t] = s
    w = rng.standard_gamma(0.3, size=(T, 8))
    w /= w.sum(axis=1, keepdims=True)
    c = rng.uniform(0.5, 0.99, size=T)
    p = w * (1 - c)[:, None]
    p[np.arange(T), state] += c
    return np.clip(np.rint(p * 255), 0, 255).astype(np.uint8)


def flipflop_log_prob(trace):
    """decode.py:92-93"""
    eps = 0.0000001
    return np.log((trace + eps) / (255 + eps))
