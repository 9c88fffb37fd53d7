"""One-off parity sweep at the bench's size: N synthetic pairs (T ~ 5000, beam width 25) through the fused GPU pipeline
against the unmodified reference core (oracle/_ref) on all host cores.  Prints the number of identical consensus
strings and the largest score difference.   usage: parity_sweep.py [n_pairs] [first_seed]"""
import multiprocessing as mp
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as O
from poreover_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
FIRST = int(sys.argv[2]) if len(sys.argv) > 2 else 5000


def ref(k):
    p1, p2, _ = synth.make_pair(k, 5000)
    lp1, lp2 = synth.bonito_log_prob(p1), O.reverse_complement(synth.bonito_log_prob(p2), "bonito")
    r = O.pair_decode(lp1, lp2, "bonito", 25, backend="ref" if O.have_ref() else "port", with_score=True)
    return r.get("consensus", ""), r.get("score", 0.0), r["basecall1"], r["basecall2"], r.get("identity", -1.0)


if __name__ == "__main__":
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        want = pool.map(ref, range(FIRST, FIRST + N), chunksize=1)
    from poreover_b200 import batch
    l1, l2 = [], []
    for k in range(FIRST, FIRST + N):
        p1, p2, _ = synth.make_pair(k, 5000)
        l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
    got = batch.pair_decode_batch(l1, l2, "bonito", beam_width=25, rc2=True)
    same = sum(g.get("consensus", "") == w[0] for g, w in zip(got, want))
    b_same = sum(g["basecall1"] == w[2] and g["basecall2"] == w[3] for g, w in zip(got, want))
    i_same = sum(g.get("identity", -1.0) == w[4] for g, w in zip(got, want))
    dsc = max(abs(g.get("score", 0.0) - w[1]) for g, w in zip(got, want))
    flags = int(np.bitwise_or.reduce([g["status"] for g in got]))
    print("pairs %d  consensus identical %d  1D basecalls identical %d  identity identical %d  max |score diff| %.3g  status bits %d"
          % (N, same, b_same, i_same, dsc, flags))
