import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth
arrays = [synth.bonito_log_prob(synth.make_read(300 + i, T)[0]) for i, T in enumerate((1, 2, 5, 40, 333, 900, 1500))]
for model in ("ctc_merge_repeats", "ctc"):
    for W in (5, 25, 100):
        seqs, sc, st = batch.beam_search_batch(arrays, W, model)
        for a, g, gs in zip(arrays, seqs, sc):
            w, ws = O.beam_search(a, W, model, with_score=True)
            if g != w or abs(gs - ws) > 1e-4:
                print("MISMATCH", model, W, len(a), g == w, gs, ws)
print("done")
