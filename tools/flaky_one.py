import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth
reps = int(sys.argv[1]); nb = int(sys.argv[2])
p1, p2, _ = synth.make_pair(13, 601)
lp1 = synth.bonito_log_prob(p1)
lp2 = np.ascontiguousarray(O.reverse_complement(synth.bonito_log_prob(p2), "bonito"))
r = O.pair_decode(lp1, lp2, "bonito", 25, method="row")
env = r["envelope"]
w = O.beam_search_2d(lp1, lp2, env, 25, "ctc_merge_repeats", "row", with_score=True)
print("env lo monotone:", bool(np.all(np.diff(env[:,0])>=0)), "hi monotone:", bool(np.all(np.diff(env[:,1])>=0)), "max width", int((env[:,1]-env[:,0]).max()))
vals = {}
for rep in range(reps):
    seqs, sc, st = batch.beam_search_2d_batch([lp1]*nb, [lp2]*nb, [env]*nb, 25, "ctc_merge_repeats", "row")
    for g, gs in zip(seqs, sc):
        key = (g == w[0], round(gs, 7))
        vals[key] = vals.get(key, 0) + 1
print("oracle", w[1], vals)
