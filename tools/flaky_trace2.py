import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["POB_DEBUG_TRACE"] = "1"
STEP = int(sys.argv[1])
os.environ["POB_DEBUG_STEP"] = str(STEP)
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth, _lib
p1, p2, _ = synth.make_pair(13, 601)
lp1 = synth.bonito_log_prob(p1)
lp2 = np.ascontiguousarray(O.reverse_complement(synth.bonito_log_prob(p2), "bonito"))
r = O.pair_decode(lp1, lp2, "bonito", 25, method="row")
env = r["envelope"]
w = O.beam_search_2d(lp1, lp2, env, 25, "ctc_merge_repeats", "row", with_score=True)
L = _lib.lib()
L.pob_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
good = bad = None
for rep in range(80):
    seqs, sc, st = batch.beam_search_2d_batch([lp1], [lp2], [env], 25, "ctc_merge_repeats", "row")
    tr = np.zeros(4000 + 4 * 160 * 10)
    L.pob_debug_trace(_lib.get_ctx().h, tr.ctypes.data_as(C.c_void_p), len(tr))
    d = tr[4000:].reshape(4, 160, 10)
    if abs(sc[0] - w[1]) < 1e-5: good = d
    else: bad = d
    if good is not None and bad is not None: break
print("have good", good is not None, "bad", bad is not None)
np.set_printoptions(linewidth=200, suppress=True, precision=4)
for k in range(4):
    g = {int(r[0]): r for r in good[k] if r[0] > 0}
    b = {int(r[0]): r for r in bad[k] if r[0] > 0}
    print("step", STEP - 2 + k, "env", env[STEP-2+k].tolist(), "n good", len(g), "n bad", len(b), "only good", sorted(set(g) - set(b))[:10], "only bad", sorted(set(b) - set(g))[:10])
    for o in sorted(set(g) & set(b)):
        if abs(g[o][4] - b[o][4]) > 1e-6 or (np.isinf(g[o][4]) != np.isinf(b[o][4])):
            print("   order", o, "pslot/porder_now/parent_order", g[o][1:4], b[o][1:4], "good score,p0,p.hi0,p.state,p.eidx,slot", g[o][4:10], "bad", b[o][4:10])
