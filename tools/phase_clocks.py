"""Cycle attribution of the beam engine's phases: builds an instrumented copy of the library
(-DPOB_PHASE_CLOCKS -> build/libporeover_b200_clk.so), runs N synthetic pairs and prints the share of
thread-0 cycles between consecutive marks.   usage: phase_clocks.py [build|run] [n_pairs] [T] [W]"""
import ctypes as C, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "libporeover_b200_clk.so")
NAMES = {0: "item setup+seed", 1: "sweep: setup + clean-max scan", 2: "sweep: phase A (private)", 3: "sweep: publish + barrier",
         4: "sweep: phase B loop", 5: "sweep: bookkeeping + keys", 6: "prune (rank)", 7: "expand X1 classify",
         8: "expand X2 retire/inspect", 9: "expand X3 create/revive", 10: "expand X4 trace ids", 11: "single-t update (skip steps)", 12: "sweep: call + views", 13: "sweep: deferred_finalize",
         14: "sweep: per-thread setup", 15: "sweep: clean-max scan (thread 0)"}
if sys.argv[1] == "build":
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(ROOT, "poreover_b200", "csrc", "*.cu")))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                           "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-DPOB_PHASE_CLOCKS"] + os.environ.get("POB_CLK_FLAGS", "").split() + ["-o", OUT] + srcs)
    sys.exit(0)
os.environ["POB_DEBUG_LIB"] = OUT
sys.path.insert(0, ROOT)
import numpy as np
from poreover_b200 import _lib, batch, synth
from poreover_b200._lib import ReadsT, check, lib
n = int(sys.argv[2]) if len(sys.argv) > 2 else 444
T = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
W = int(sys.argv[4]) if len(sys.argv) > 4 else 25
uniq = min(n, 64)
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
clk = (C.c_ulonglong * 32)()
raw = C.CDLL(OUT)
ctx.profile(True)
for c in range(2):
    ctx.profile_reset()
    raw.pob_debug_phase_clocks(clk, 1)
    check(L.pob_pair_decode(ctx.h, _lib.DEVICE, C.byref(d1), C.byref(d2), 1, W, 5, 500, 1, *o), "pair_decode")
    ctx.sync()
    p = ctx.profile_get()
raw.pob_debug_phase_clocks(clk, 0)
cnt = ctx.counters()
tot = float(sum(clk))
print("beam ms", p["beam_pair"]["ms"], cnt, "steps/pair", cnt["steps"] / n)
for i in range(16):
    print("%5.1f%%  %8.0f cyc/step  %s" % (100 * clk[i] / tot, clk[i] / cnt["steps"], NAMES[i]))
