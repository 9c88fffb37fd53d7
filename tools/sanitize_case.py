"""Small search cases for compute-sanitizer (memcheck / racecheck): one pair, both traversals, both trees."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 150
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bad = 0
for k in range(reps):
    p1, p2, _ = synth.make_pair(20 + k, T)
    lp1 = synth.bonito_log_prob(p1)
    lp2 = np.ascontiguousarray(O.reverse_complement(synth.bonito_log_prob(p2), "bonito"))
    r = O.pair_decode(lp1, lp2, "bonito", 25)
    env = r["envelope"]
    for method in ("row", "row_col"):
        for model in ("ctc_merge_repeats", "ctc"):
            for W in (5, 25):
                seqs, sc, st = batch.beam_search_2d_batch([lp1], [lp2], [env], W, model, method)
                w, ws = O.beam_search_2d(lp1, lp2, env, W, model, method, with_score=True)
                ok = seqs[0] == w and abs(sc[0] - ws) < 1e-4
                if not ok:
                    bad += 1
                    print("MISMATCH", k, method, model, W, sc[0], ws, len(seqs[0]), len(w))
    s1, sc1, _ = batch.beam_search_batch([lp1], 25, "ctc_merge_repeats")
    if s1[0] != O.beam_search(lp1, 25, "ctc_merge_repeats"):
        bad += 1
        print("MISMATCH 1d", k)
print("done, mismatches:", bad)
