"""Wall-clock stages of pob_pair_decode(POB_HOST) on a small chunk (POB_DEBUG_TIMING=1): where a host-buffer call spends
its time next to the same call on device-resident inputs.   usage: host_call_timing.py [n_pairs] [calls]"""
import ctypes as C, os, sys, time
os.environ["POB_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poreover_b200 import _lib, batch, synth
from poreover_b200._lib import ReadsT, check, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 313
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
l1, l2 = [], []
for k in range(n):
    p1, p2, _ = synth.make_pair(k, 5000)
    l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
ctx = _lib.get_ctx(0); L = lib()
b1 = batch.ReadBatch(l1); b2 = batch.ReadBatch(l2, rc=np.ones(n, np.uint8))
def pinned(a):
    p = _lib.vp(); check(L.pob_malloc_host(a.nbytes + 64, C.byref(p))); C.memmove(p.value, a.ctypes.data, a.nbytes); return p.value
def host(b):
    return ReadsT(pinned(b.data), b.row_off.ctypes.data, b.lens.ctypes.data, b.rc.ctypes.data if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)
def dev(b):
    return ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens), ctx.to_device(b.rc) if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)
r1, r2 = b1.total_rows, b2.total_rows
ho = [np.zeros(x, np.uint8) for x in (r1 + 64, 4 * n + 64, r2 + 64, 4 * n + 64, r1 + r2 + 64, 4 * n + 64, 8 * n + 64, 16 * n + 64, 4 * n + 64)]
do = [ctx.malloc(x) for x in (r1 + 64, 4 * n + 64, r2 + 64, 4 * n + 64, r1 + r2 + 64, 4 * n + 64, 8 * n + 64, 16 * n + 64, 4 * n + 64)]
h1, h2, d1, d2 = host(b1), host(b2), dev(b1), dev(b2)
for where, a, o in ((_lib.DEVICE, (d1, d2), do), (_lib.HOST, (h1, h2), [x.ctypes.data for x in ho])):
    for c in range(calls):
        sys.stderr.write("---- %s call %d\n" % ("HOST" if where == _lib.HOST else "DEVICE", c))
        t0 = time.perf_counter()
        check(L.pob_pair_decode(ctx.h, where, C.byref(a[0]), C.byref(a[1]), 1, 25, 5, 500, 1, *o), "pair_decode")
        ctx.sync()
        sys.stderr.write("total %.3f ms\n" % (1e3 * (time.perf_counter() - t0)))
