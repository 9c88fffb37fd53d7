import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poreover_b200 import _lib, batch, synth
from poreover_b200._lib import ReadsT, check, lib
n, T = 10000, 5000
uniq = [synth.bonito_log_prob(synth.make_read(i, T)[0]) for i in range(50)]
ctx = _lib.get_ctx(0); L = lib()
for name, rc in (("forward", np.zeros(n, np.uint8)), ("half rc", (np.arange(n) % 2).astype(np.uint8)), ("all rc", np.ones(n, np.uint8))):
    b = batch.ReadBatch((uniq * (n // 50 + 1))[:n], rc=rc)
    d = ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens), ctx.to_device(b.rc), b.n, 5, b.dtype, b.layout)
    rows = b.total_rows
    o = [ctx.malloc(rows + 64), ctx.malloc(4 * rows + 64), ctx.malloc(4 * n + 64), ctx.malloc(4 * n + 64)]
    for i in range(3):
        check(L.pob_viterbi(ctx.h, _lib.DEVICE, C.byref(d), 1, o[0], o[1], None, o[2], o[3]), "viterbi")
    ctx.sync(); ctx.profile(True); ctx.profile_reset()
    for i in range(10):
        check(L.pob_viterbi(ctx.h, _lib.DEVICE, C.byref(d), 1, o[0], o[1], None, o[2], o[3]), "viterbi")
    p = ctx.profile_get()["viterbi_ctc"]; ctx.profile(False)
    lens = ctx.from_device(o[2], (n,), np.int32)
    alg = float(b.lens.sum()) * 20 + float(lens.sum()) * 5 + n * 8
    ms = p["ms"] / p["launches"]
    print("%-8s ms/launch %.4f  GB/s %.0f  frac of 6554: %.3f" % (name, ms, alg / (ms / 1e3) / 1e9, alg / (ms / 1e3) / 1e9 / 6554.2))
    for x in o: ctx.free(x)
