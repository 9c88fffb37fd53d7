import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
l1, l2, envs = [], [], []
for k, T in ((10, 200), (11, 350), (12, 600), (13, 601), (14, 900), (15, 1200)):
    p1, p2, _ = synth.make_pair(k, T)
    lp1 = synth.bonito_log_prob(p1)
    lp2 = np.ascontiguousarray(O.reverse_complement(synth.bonito_log_prob(p2), "bonito"))
    r = O.pair_decode(lp1, lp2, "bonito", 25, method="row")
    l1.append(lp1); l2.append(lp2); envs.append(r["envelope"])
want = {}
for method in ("row", "row_col"):
    for model in ("ctc_merge_repeats", "ctc"):
        for W in (5, 25):
            want[(method, model, W)] = [O.beam_search_2d(a, b, e, W, model, method, with_score=True) for a, b, e in zip(l1, l2, envs)]
bad = 0
for rep in range(reps):
    for key, w in want.items():
        method, model, W = key
        seqs, sc, st = batch.beam_search_2d_batch(l1, l2, envs, W, model, method)
        for i, (g, gs) in enumerate(zip(seqs, sc)):
            if g != w[i][0] or abs(gs - w[i][1]) > 1e-4:
                bad += 1
                d = next((j for j in range(min(len(g), len(w[i][0]))) if g[j] != w[i][0][j]), -1)
                print("MISMATCH rep", rep, key, "item", i, "T", len(l1[i]), "len", len(g), len(w[i][0]), "first diff at", d, "score", gs, w[i][1], "status", st[i])
print("done, mismatches:", bad)
