"""Throughput of the other BASELINE.json configurations through the batch API (host buffers in, host results out;
kernel times from the library's CUDA-event profile).   usage: bench_configs.py [cfg2|cfg4|cfg5|all] [scale]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poreover_b200 import _lib, batch, synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ctx = _lib.get_ctx(0)
out = {}


def timed(fn, reps=2):
    fn()
    ctx.profile(True); ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    dt = (time.perf_counter() - t0) / reps
    prof = {k: v["ms"] / reps for k, v in ctx.profile_get().items() if v["ms"] > 0}
    ctx.profile(False)
    return r, dt, prof


if which in ("cfg2", "all"):
    n = int(2000 * scale)
    uniq = [synth.bonito_log_prob(synth.make_read(i, 5000)[0]) for i in range(min(n, 500))]
    reads = (uniq * (n // len(uniq) + 1))[:n]
    for W in (25, 100):
        (seqs, sc, st), dt, prof = timed(lambda: batch.beam_search_batch(reads, W, "ctc_merge_repeats"), reps=1)
        out["cfg2_beam_w%d" % W] = {"reads": n, "T": 5000, "e2e_reads_per_s": n / dt, "kernel_ms": prof,
                                    "kernel_reads_per_s": n / (prof.get("beam_single", 1e9) / 1e3),
                                    "overflow": int(sum(1 for s in st if s & 4))}
    big = (uniq * (10000 // len(uniq) + 1))[:10000]
    _, dt, prof = timed(lambda: batch.viterbi_batch(big, "bonito"))
    out["cfg2_viterbi"] = {"reads": 10000, "e2e_reads_per_s": 10000 / dt, "kernel_ms": prof}

if which in ("cfg5", "all"):
    n = int(10000 * scale)
    uniq = [synth.make_flipflop_trace(i, 5000) for i in range(200)]
    traces = (uniq * (n // 200 + 1))[:n]
    (res), dt, prof = timed(lambda: batch.flipflop_viterbi_batch(traces))
    out["cfg5_flipflop"] = {"reads": n, "T": 5000, "e2e_reads_per_s": n / dt, "kernel_ms": prof,
                            "kernel_reads_per_s": n / (prof.get("viterbi_flipflop", 1e9) / 1e3)}

if which in ("cfg4", "all"):
    n = int(32 * scale)
    l1, l2 = [], []
    rng = np.random.default_rng(7)
    for k in range(n):
        T = int(rng.integers(50000, 100001))
        p1, p2, _ = synth.make_pair(5000 + k, T)
        l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
    res, dt, prof = timed(lambda: batch.pair_decode_batch(l1, l2, "bonito", beam_width=25, padding=150, rc2=True), reps=1)
    out["cfg4_long_pairs"] = {"pairs": n, "T": "50k-100k", "padding": 150, "e2e_pairs_per_s": n / dt, "kernel_ms": prof,
                              "consensus_bases": int(sum(len(r.get("consensus", "")) for r in res)),
                              "status_or": int(np.bitwise_or.reduce([r["status"] for r in res]))}
print(json.dumps(out, indent=1))
