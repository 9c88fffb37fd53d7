"""Minimal driver for ncu: one (or a few) pob_pair_decode calls over N synthetic pairs, device-resident inputs."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poreover_b200 import _lib, batch, synth
from poreover_b200._lib import ReadsT, check, lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 444
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
W = int(sys.argv[4]) if len(sys.argv) > 4 else 25
uniq = min(n, int(os.environ.get('POB_PROF_UNIQUE', '64')))
l1, l2 = [], []
for k in range(uniq):
    p1, p2, _ = synth.make_pair(k, T)
    l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
l1 = (l1 * (n // uniq + 1))[:n]; l2 = (l2 * (n // uniq + 1))[:n]
ctx = _lib.get_ctx(0); L = lib()
b1 = batch.ReadBatch(l1); b2 = batch.ReadBatch(l2, rc=np.ones(n, np.uint8))
def dev(b):
    return ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens),
                  ctx.to_device(b.rc) if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)
d1, d2 = dev(b1), dev(b2)
r1, r2 = b1.total_rows, b2.total_rows
o = [ctx.malloc(x) for x in (r1 + 64, 4 * n + 64, r2 + 64, 4 * n + 64, r1 + r2 + 64, 4 * n + 64, 8 * n + 64, 16 * n + 64, 4 * n + 64)]
ctx.profile(True)
for c in range(calls):
    ctx.profile_reset()
    check(L.pob_pair_decode(ctx.h, _lib.DEVICE, C.byref(d1), C.byref(d2), 1, W, 5, 500, 1, *o), "pair_decode")
    ctx.sync()
    p = ctx.profile_get()
    print("call", c, {k: round(v["ms"], 3) for k, v in p.items()}, ctx.counters())
try:
    ex = C.c_ulonglong(0)
    L.pob_debug_exact_prunes(C.byref(ex), 0)
    print("exact prune passes:", ex.value, "of", ctx.counters())
except AttributeError:
    pass
st = ctx.from_device(o[8], (n,), np.int32)
print("status bits:", np.bitwise_or.reduce(st), "pairs/s (beam only):", n / (p["beam_pair"]["ms"] / 1e3))
