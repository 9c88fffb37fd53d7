"""ncu driver: the HBM-bound Viterbi kernel on 10k reads x T=5000 (1.02 GB of float32 log-probabilities)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poreover_b200 import _lib, batch, synth
from poreover_b200._lib import ReadsT, check, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
T = 5000
uniq = [synth.bonito_log_prob(synth.make_read(i, T)[0]) for i in range(50)]
b = batch.ReadBatch((uniq * (n // 50 + 1))[:n], rc=(np.arange(n) % 2).astype(np.uint8))
ctx = _lib.get_ctx(0); L = lib()
d = ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens), ctx.to_device(b.rc), b.n, 5, b.dtype, b.layout)
rows = b.total_rows
o = [ctx.malloc(rows + 64), ctx.malloc(4 * rows + 64), ctx.malloc(4 * n + 64), ctx.malloc(4 * n + 64)]
ctx.profile(True)
for i in range(6):
    check(L.pob_viterbi(ctx.h, _lib.DEVICE, C.byref(d), 1, o[0], o[1], None, o[2], o[3]), "viterbi")
ctx.sync()
p = ctx.profile_get()["viterbi_ctc"]
lens = ctx.from_device(o[2], (n,), np.int32)
alg = float(b.lens.sum()) * 20 + float(lens.sum()) * 5 + n * 8
print("ms/launch", p["ms"] / p["launches"], "GB/s", alg / (p["ms"] / p["launches"] / 1e3) / 1e9)
