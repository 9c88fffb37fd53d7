import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["POB_DEBUG_TRACE"] = "1"
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth, _lib
p1, p2, _ = synth.make_pair(13, 601)
lp1 = synth.bonito_log_prob(p1)
lp2 = np.ascontiguousarray(O.reverse_complement(synth.bonito_log_prob(p2), "bonito"))
r = O.pair_decode(lp1, lp2, "bonito", 25, method="row")
env = r["envelope"]
# reference per-step top scores
ref = O.ref()
ref.ref_set_trace(1)
w = O.beam_search_2d(lp1, lp2, env, 25, "ctc_merge_repeats", "row", backend="ref", with_score=True)
n = ref.ref_trace_len()
top = np.zeros(n); dep = np.zeros(n, dtype=np.int32)
ref.ref_trace_get(top.ctypes.data_as(C.c_void_p), dep.ctypes.data_as(C.c_void_p))
ref.ref_set_trace(0)
L = _lib.lib()
L.pob_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
good = bad = None
for rep in range(60):
    seqs, sc, st = batch.beam_search_2d_batch([lp1], [lp2], [env], 25, "ctc_merge_repeats", "row")
    tr = np.zeros(2 * 700)
    L.pob_debug_trace(_lib.get_ctx().h, tr.ctypes.data_as(C.c_void_p), len(tr))
    tr = tr.reshape(-1, 2)
    if abs(sc[0] - w[1]) < 1e-5:
        good = tr
    else:
        bad = tr
    if good is not None and bad is not None:
        break
print("ref steps", n, "have good", good is not None, "have bad", bad is not None)
if good is not None:
    d = np.abs(good[:n, 0] - top)
    print("good vs ref: max |dtop|", d.max(), "at", int(d.argmax()))
if bad is not None:
    d = np.abs(bad[:n, 0] - top)
    i = int(np.argmax(d > 1e-6)) if (d > 1e-6).any() else -1
    print("bad vs ref: first step with |dtop|>1e-6:", i, "values", bad[max(i-1,0):i+3, 0], top[max(i-1,0):i+3])
    if good is not None:
        ds = np.abs(bad[:n, 1] - good[:n, 1])
        j = int(np.argmax(ds > 1e-6)) if (ds > 1e-6).any() else -1
        print("bad vs good: first step where the beam-score sum differs:", j, bad[max(j-1,0):j+2, 1], good[max(j-1,0):j+2, 1])
        print("envelope rows around:", env[max(j-2,0):j+3].tolist() if j >= 0 else None)
