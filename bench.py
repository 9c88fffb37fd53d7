#!/usr/bin/env python
"""Benchmark of the pair-decode hot path (BASELINE.json metric: pair-decoded pairs/s and consensus
Mbases/s, next to the reference CPU decoder on the host cores).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (rank 0 only)

A "step" is one pass of the hot path (viterbi x2 -> mapping -> banded NW -> envelope -> row_col beam
search) over one batch of synthetic Bonito-shaped pairs (T ~ 5000, beam width 25, --reverse_complement).
The workload is BASELINE.json configs[2]: one 10k-pair batch (T ~ 5000, beam width 25); it fits one GPU, so every
GPU decodes its own `--pairs-per-gpu` = 10000 pairs per step (weak scaling; pairs shard by pair, no data-path
collective).  `--unique-pairs` distinct synthetic pairs are generated per rank and tiled to the batch size.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from poreover_b200 import synth  # noqa: E402


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def make_pairs(first, count, T, unique=None):
    unique = min(count, unique or count)
    l1, l2 = [], []
    for k in range(first, first + unique):
        p1, p2, _ = synth.make_pair(k, T)
        l1.append(synth.bonito_log_prob(p1))  # exactly what the reference loader hands to the decoders
        l2.append(synth.bonito_log_prob(p2))
    reps = -(-count // unique)
    return (l1 * reps)[:count], (l2 * reps)[:count]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- reference arm
_REF_DATA = None


def _ref_worker(i):
    from oracle import oracle as O
    lp1, lp2 = _REF_DATA[0][i], _REF_DATA[1][i]
    backend = "ref" if O.have_ref() else "port"
    r = O.pair_decode(lp1, O.reverse_complement(lp2, "bonito"), "bonito", _REF_DATA[2], padding=5, method="row_col",
                      backend=backend)
    return len(r.get("consensus", "")), r.get("skipped", 0)


def cpu_reference(l1, l2, beam_width, n_pairs, steps, warmup, cores=None):
    """The reference's CPU implementation of the path on the host cores: its own C++ search core and Cython
    aligner compiled unmodified into oracle/_ref (else the oracle port), one pair per worker process like the
    reference's multiprocessing.Pool (pair_decode.py:292-297)."""
    global _REF_DATA
    import multiprocessing as mp
    from oracle import oracle as O
    O.port()
    cores = cores or os.cpu_count() or 1
    try:
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available // (768 << 20))))
    except ImportError:
        pass
    _REF_DATA = (l1, l2, beam_width)
    ctx = mp.get_context("fork")
    times, bases = [], 0
    with ctx.Pool(processes=cores) as pool:
        for s in range(warmup + steps):
            idx = [(s * n_pairs + j) % len(l1) for j in range(n_pairs)]
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, idx, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                bases += sum(r[0] for r in res)
    total = sum(times)
    return {"pairs_per_s": n_pairs * len(times) / total, "bases_per_s": bases / total, "cores": cores,
            "kind": "reference" if O.have_ref() else "port", "ms_per_step": 1e3 * total / len(times),
            "sample": "%d pairs per step x %d steps (same synthetic pairs as the GPU arm, T~%d, beam %d)"
                      % (n_pairs, len(times), len(l1[0]), beam_width)}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_pairs = args.ref_pairs or 4 * cores  # a subsample of at least 4 x cores pairs
    uniq = min(max(n_pairs, 8), 256)
    l1, l2 = make_pairs(0, uniq, args.T)
    r = cpu_reference(l1, l2, args.beam_width, n_pairs, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "pair_decode_pairs_per_s", "value": r["pairs_per_s"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "consensus_mbases_per_s": r["bases_per_s"] / 1e6,
        "config": workload_config(args, args.pairs_per_gpu),  # the arm's workload; this run's bounded sample is in cpu_baseline
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["pairs_per_s"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, pairs_per_unit):
    return {"workload": "pair-decode of synthetic bonito pairs (2 reads, T~%d x5 CTC, --reverse_complement, "
                        "beam_width %d, padding 5, banded NW 500, row_col)" % (args.T, args.beam_width),
            "pairs_per_gpu_per_step": pairs_per_unit, "unique_pairs_per_gpu": min(pairs_per_unit, args.unique_pairs),
            "T": args.T, "beam_width": args.beam_width,
            "cache": "inputs per step (%.0f MB/GPU) exceed the 126 MB L2" % (pairs_per_unit * 2.04 * args.T * 20 / 1e6),
            "sharding": "by pair, no collective"}


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    rank, world, local = dist_env()
    from poreover_b200 import _lib, batch
    from poreover_b200._lib import ReadsT, check, lib, ptr

    use_dist = world > 1
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.get_ctx(local)
    L = lib()
    P = args.pairs_per_gpu
    l1, l2 = make_pairs(rank * args.unique_pairs, P, args.T, args.unique_pairs)
    b1 = batch.ReadBatch(l1)
    b2 = batch.ReadBatch(l2, rc=np.ones(P, dtype=np.uint8))
    n = P
    rows1, rows2 = b1.total_rows, b2.total_rows
    kind, method = _lib.KIND["bonito"], _lib.METHOD["row_col"]

    def barrier():
        ctx.sync()
        if use_dist:
            dist.barrier()

    # ---------------- device-resident leg: inputs already in HBM when the timed region starts
    def dev_reads(b):
        d = ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens),
                   ctx.to_device(b.rc) if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)
        return d

    d1, d2 = dev_reads(b1), dev_reads(b2)
    o_seq1, o_seq2 = ctx.malloc(rows1 + 64), ctx.malloc(rows2 + 64)
    o_cons = ctx.malloc(rows1 + rows2 + 64)
    o_l1, o_l2, o_lc, o_st = (ctx.malloc(4 * n + 64) for _ in range(4))
    o_sc, o_stats = ctx.malloc(8 * n + 64), ctx.malloc(16 * n + 64)

    def step_device():
        check(L.pob_pair_decode(ctx.h, _lib.DEVICE, C.byref(d1), C.byref(d2), kind, args.beam_width, 5, 500, method,
                                o_seq1, o_l1, o_seq2, o_l2, o_cons, o_lc, o_sc, o_stats, o_st), "pob_pair_decode")

    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.profile(True)
    ctx.profile_reset()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    prof = ctx.profile_get()
    counters = ctx.counters()
    ctx.profile(False)
    if use_dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    else:
        ms_max = ms
    cons_len = ctx.from_device(o_lc, (n,), np.int32)
    status = ctx.from_device(o_st, (n,), np.int32)
    bases = int(cons_len.sum())
    if use_dist:
        t = torch.tensor([bases], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        bases_all = float(t.item())
    else:
        bases_all = float(bases)
    value = world * P * args.steps / (ms_max / 1e3)
    mbases = bases_all * args.steps / (ms_max / 1e3) / 1e6

    # ---------------- end-to-end leg: the call a user makes, HOST buffers (pinned), copies inside the timed region
    def pinned_like(a):
        p = _lib.vp()
        check(L.pob_malloc_host(max(a.nbytes, 1) + 64, C.byref(p)), "pob_malloc_host")
        C.memmove(p.value, a.ctypes.data, a.nbytes)
        return p.value

    def host_reads(b):
        return ReadsT(pinned_like(b.data), b.row_off.ctypes.data, b.lens.ctypes.data,
                      b.rc.ctypes.data if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)

    h1, h2 = host_reads(b1), host_reads(b2)
    hs1, hs2 = np.zeros(rows1 + 64, np.uint8), np.zeros(rows2 + 64, np.uint8)
    hc = np.zeros(rows1 + rows2 + 64, np.uint8)
    hl1, hl2, hlc, hst = (np.zeros(n, np.int32) for _ in range(4))
    hsc, hstats = np.zeros(n, np.float64), np.zeros((n, 4), np.int32)

    def step_host():
        check(L.pob_pair_decode(ctx.h, _lib.HOST, C.byref(h1), C.byref(h2), kind, args.beam_width, 5, 500, method,
                                ptr(hs1), ptr(hl1), ptr(hs2), ptr(hl2), ptr(hc), ptr(hlc), ptr(hsc), ptr(hstats),
                                ptr(hst)), "pob_pair_decode(host)")

    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    if use_dist:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    barrier()
    assert np.array_equal(hlc, cons_len), "host and device legs disagree"
    h2d = b1.data.nbytes + b2.data.nbytes + 2 * (b1.row_off.nbytes + b1.lens.nbytes) + n
    d2h = rows1 + rows2 + (rows1 + rows2) + n * (4 * 4 + 8)
    e2e = {"value": world * P * args.steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / args.steps}

    # ---------------- roofline of the HBM-bound kernel (north star: Viterbi >= 60% of HBM peak): 10k reads T=5000
    roof = None
    beam = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # first reads of the pairs (T rows each), tiled to --viterbi-reads: the same shape as tools/prof_viterbi.py, the
        # workload of the committed ncu capture that `traffic` comes from
        reps = max(1, int(np.ceil(args.viterbi_reads / float(P))))
        arrays = (l1 * reps)[:args.viterbi_reads]
        vb = batch.ReadBatch(arrays)
        dv = dev_reads(vb)
        rows = vb.total_rows
        v_seq, v_s2s = ctx.malloc(rows + 64), ctx.malloc(4 * rows + 64)
        v_len, v_st = ctx.malloc(4 * vb.n + 64), ctx.malloc(4 * vb.n + 64)

        def vit():
            check(L.pob_viterbi(ctx.h, _lib.DEVICE, C.byref(dv), kind, v_seq, v_s2s, None, v_len, v_st), "pob_viterbi")

        for _ in range(3):
            vit()
        ctx.sync()
        ctx.profile(True)
        ctx.profile_reset()
        for _ in range(10):
            vit()
        vp = ctx.profile_get()["viterbi_ctc"]
        ctx.profile(False)
        lens = ctx.from_device(v_len, (vb.n,), np.int32)
        alg_bytes = float(vb.lens.sum()) * 20 + float(lens.sum()) * 5 + vb.n * 8  # 20T in, L bases + 4L mapping out
        t_ms = vp["ms"] / vp["launches"]
        ach = alg_bytes / (t_ms / 1e3) / 1e9
        # DRAM traffic of one launch of this kernel on this workload, from the committed ncu --set full capture
        traffic, traffic_src = None, None
        try:
            ns = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r01_f.json")))["viterbi5_f32_kernel"]
            if vb.n == 10000 and args.T == 5000:
                traffic = ns["dram_traffic_bytes_per_launch"]
                traffic_src = "profiles/ncu_summary_r01_f.json (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"
        except (OSError, ValueError, KeyError):
            pass
        roof = {"kernel": "viterbi_ctc (%d reads, T=%d, %.2f GB in)" % (vb.n, args.T, float(vb.lens.sum()) * 20 / 1e9),
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src, "ms_per_launch": t_ms}
        bk = prof.get("beam_pair", {"ms": 0.0, "launches": 1})
        tot_ms = sum(v["ms"] for v in prof.values())
        beam = {"kernel": "beam_pair", "bound": "fp64-add + fp32-sfu issue / dependent-chain latency (not hbm, not tensor)",
                "cell_updates_per_launch": counters["cell_updates"], "ms_per_launch": bk["ms"] / max(1, bk["launches"]),
                "cell_updates_per_s": counters["cell_updates"] / max(1e-9, bk["ms"] / max(1, bk["launches"]) / 1e3),
                "share_of_step": bk["ms"] / tot_ms if tot_ms else None}
        # the same kernel in the HBM frame, for comparison: minimal bytes a pair needs (20 (U+V) in, 8 U of
        # envelope, the consensus out) against the measured peak, and its DRAM traffic from the committed ncu capture
        # (444 pairs per launch there, scaled per pair): the traffic is the engine's per-node windows streaming
        # through L2, not re-reads of the input
        beam_ms = bk["ms"] / max(1, bk["launches"])
        alg = 20.0 * (rows1 + rows2) + 8.0 * rows1 + bases
        beam["hbm_frame"] = {"algorithmic_bytes_per_launch": alg, "achieved": alg / max(1e-9, beam_ms / 1e3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": alg / max(1e-9, beam_ms / 1e3) / 1e9 / peak}
        try:
            nb = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r01_f.json")))["beam_kernel"]
            per_pair = (nb["dram_read"] + nb["dram_write"]) * 1e9 / nb["grid"]
            beam["hbm_frame"]["traffic"] = per_pair * P
            beam["hbm_frame"]["traffic_source"] = ("profiles/ncu_summary_r01_f.json: (dram__bytes_read.sum + "
                                                   "dram__bytes_write.sum) / 444 pairs x pairs per launch")
        except (OSError, ValueError, KeyError):
            beam["hbm_frame"]["traffic"] = None
        # ... and in the instruction-issue frame, the resource the kernel does use: warp instructions per pair from the
        # same capture x pairs per launch / the live kernel time, against 4 issue slots per SM and clock (148 SMs)
        try:
            nb = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r01_f.json")))["beam_kernel"]
            sm_mhz = float((clocks or {}).get("sm_mhz") or 0.0) or 1965.0
            issued = float(nb["warp_instructions"]) / float(nb["grid"]) * P / max(1e-9, beam_ms / 1e3)
            peak_issue = 148 * 4 * sm_mhz * 1e6
            beam["issue_frame"] = {"achieved": issued / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instructions/s",
                                   "frac": issued / peak_issue,
                                   "source": "profiles/ncu_summary_r01_f.json smsp__inst_executed.sum / 444 pairs x pairs "
                                             "per launch / live kernel time; peak = 148 SMs x 4 schedulers x SM clock"}
        except Exception:  # the frame is an annotation: never let it take the bench line down
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        npairs = min(P, max(4 * cores, 4))  # at least 4 x cores pairs (10-30 s of CPU work)
        r = cpu_reference(l1[:min(P, 256)], l2[:min(P, 256)], args.beam_width, npairs, 1, 0)
        cpu = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"], "consensus_mbases_per_s": r["bases_per_s"] / 1e6}

    if rank == 0:
        line = {
            "metric": "pair_decode_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, P), "consensus_mbases_per_s": mbases,
            "e2e": e2e, "gpu_launches": int(counters["launches"]), "clocks": clocks,
            "roofline": roof, "roofline_beam": beam, "cpu_baseline": cpu,
            "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "pairs_skipped": int(((status & (16 | 32 | 8 | 64)) != 0).sum()),
            "pairs_pool_overflow": int(((status & 4) != 0).sum()),
        }
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=10000)
    ap.add_argument("--unique-pairs", type=int, default=2500, help="distinct synthetic pairs per GPU (tiled to the batch)")
    ap.add_argument("--T", type=int, default=5000)
    ap.add_argument("--beam-width", type=int, default=25)
    ap.add_argument("--viterbi-reads", type=int, default=10000)
    ap.add_argument("--ref-pairs", type=int, default=0, help="pairs per step of the reference arm (default: 4 x host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
