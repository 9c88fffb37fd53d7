"""Throughput of the legacy prefix search (decode --algorithm prefix): N windows of T rows on the GPU (device time of
the kernel + the host-buffer call) next to the unmodified reference's prefix_search_log_cy on a sample of the same
windows on one host core.   usage: bench_prefix.py [n_windows] [T] [ref_sample] [n_pairs] [U]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from poreover_b200 import _lib, batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sample = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n2 = int(sys.argv[4]) if len(sys.argv) > 4 else 256
U = int(sys.argv[5]) if len(sys.argv) > 5 else 60
rng = np.random.default_rng(3)


def table(T, peaked=8):
    x = rng.random((T, 5)) ** peaked
    x[:, -1] *= 1.5
    x /= x.sum(axis=1, keepdims=True)
    return np.log(x)


uniq = [table(T) for _ in range(64)]
wins = [uniq[i % 64] for i in range(n)]
ctx = _lib.get_ctx(0)
batch.prefix_search_batch(wins[:64], _lib.PREFIX_CY)  # warm-up
ctx.profile(True); ctx.profile_reset()
t0 = time.perf_counter()
labels, score, st = batch.prefix_search_batch(wins, _lib.PREFIX_CY)
dt = time.perf_counter() - t0
k1 = ctx.profile_get()["prefix_search"]
base = [uniq[i % 64][:U] for i in range(n2)]
p1 = [np.log(0.8 * np.exp(b) + 0.2 * np.exp(table(U, 1))) for b in base]
p2 = [np.log(0.8 * np.exp(b) + 0.2 * np.exp(table(U, 1))) for b in base]
batch.pair_prefix_search_batch(p1[:8], p2[:8], _lib.PREFIX_CY)
ctx.profile_reset()
t0 = time.perf_counter()
l2, s2, _ = batch.pair_prefix_search_batch(p1, p2, _lib.PREFIX_CY)
dt2 = time.perf_counter() - t0
k2 = ctx.profile_get()["pair_prefix_search"]
out = {"windows": n, "T": T, "bases_per_window": float(np.mean([len(l) for l in labels])),
       "gpu_call_s": dt, "gpu_kernel_ms": k1["ms"], "windows_per_s_call": n / dt,
       "pairs": n2, "U": U, "pair_call_s": dt2, "pair_kernel_ms": k2["ms"], "pairs_per_s_call": n2 / dt2,
       "bases_per_pair": float(np.mean([len(l) for l in l2]))}
try:
    import ref_python_driver
    ref_python_driver.load_reference_package()
    from poreover.decoding import prefix_search as RP
    t0 = time.perf_counter()
    same = 0
    for i in range(sample):
        lab, p = RP.prefix_search_log_cy(wins[i])
        same += ["ACGT".index(c) for c in lab] == labels[i].tolist() and abs(p - score[i]) < 1e-6
    rt = time.perf_counter() - t0
    t0 = time.perf_counter()
    same2 = 0
    for i in range(min(sample, n2)):
        lab, p = RP.pair_prefix_search_log_cy(p1[i], p2[i])
        same2 += ["ACGT".index(c) for c in lab] == l2[i].tolist() and abs(p - s2[i]) < 1e-6
    rt2 = time.perf_counter() - t0
    out.update({"reference_windows_per_s_one_core": sample / rt, "reference_sample": sample, "sample_identical": int(same),
                "reference_pairs_per_s_one_core": min(sample, n2) / rt2, "pair_sample_identical": int(same2)})
except Exception as e:  # noqa: BLE001
    out["reference"] = "unavailable: %r" % (e,)
print(json.dumps(out))
