#!/usr/bin/env python
"""Benchmark of the decoding hot path (BASELINE.json metric: pair-decoded pairs/s and consensus Mbases/s at
1/2/4/8 B200, next to the reference CPU decoder on the host cores).

    python bench.py --gpus N --steps K --warmup W                  # our CUDA path, BASELINE configs[2]
    python bench.py --impl reference --gpus N --steps K ...        # the reference's CPU path (rank 0 only)
    python bench.py --config cfg2_viterbi|cfg2_beam25|cfg2_beam100|cfg4_long|cfg5_flipflop   # the other configs

Default workload (configs[2]): ONE job of `--pairs` = 10,000 synthetic Bonito-shaped pairs (T ~ 5000, beam width 25,
--reverse_complement) decoded by all N GPUs together -- STRONG scaling, the way BASELINE.json states it ("sharded by
pair across 8xB200").  A "step" is one pass of the whole hot path (viterbi x2 -> mapping -> banded NW -> envelope ->
row_col beam search) over that job: every rank pulls chunks of pairs from a host work queue (poreover_b200.multigpu.
WorkQueue: fetch-and-add on torch.distributed's TCPStore; one process per GPU, two GPU calls in flight per process) and
decodes them with pob_pair_decode; no data-path collective exists or is faked.
  value : device-resident leg (the job's inputs sit in the HBM of every GPU before the timed region), CUDA events.
  e2e   : the same job through pob_pair_decode(POB_HOST) on pinned host buffers, H2D + D2H inside the timed region, and
          the per-pair records gathered on rank 0 (gather_object over gloo) inside it too; wall clock, max over ranks.
For N > 1 the line also carries "weak": the replica measurement of round 1 (every GPU decodes its own 10k pairs).
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


def measured_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        peaks = {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    return hbm, src


def ncu_summary():
    """Key metrics of the committed `ncu --set full` captures of this round (profiles/ncu_summary_r02.json)."""
    for name in ("ncu_summary_r02.json", "ncu_summary_r01_f.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), "profiles/" + name
        except (OSError, ValueError):
            continue
    return {}, None


# --------------------------------------------------------------------------------------- reference arm
_REF_DATA = None


def _ref_worker(i):
    from oracle import oracle as O
    kind, data, W = _REF_DATA
    backend = "ref" if O.have_ref() else "port"
    if kind == "pair":
        lp1, lp2 = data[0][i], data[1][i]
        r = O.pair_decode(lp1, O.reverse_complement(lp2, "bonito"), "bonito", W, padding=data[2], method="row_col",
                          backend=backend)
        return len(r.get("consensus", "")), r.get("skipped", 0)
    if kind == "beam":
        return len(O.beam_search(data[i], W, "ctc_merge_repeats", backend=backend)), 0
    if kind == "viterbi":
        return len(O.viterbi(data[i], "bonito")[0]), 0
    if kind == "flipflop":
        return len(O.viterbi(synth.flipflop_log_prob(data[i]), "flipflop")[0]), 0
    raise ValueError(kind)


def cpu_reference(kind, data, beam_width, n_items, steps, warmup, n_data, cores=None):
    """The reference's CPU implementation of the path on the host cores, one item per worker process like the
    reference's multiprocessing.Pool (pair_decode.py:292-297, decode.py:158-162).  Searches and the banded aligner
    are the reference's own C++ / Cython compiled unmodified into oracle/_ref (else the oracle port); the numpy /
    pure-Python glue between them (argmax Viterbi, sequence mapping, envelope) runs from the oracle port."""
    global _REF_DATA
    import multiprocessing as mp
    from oracle import oracle as O
    O.port()
    loaded = []
    if O.have_ref():
        # load the reference's compiled core in the PARENT, before the fork: the workers inherit the mappings and the
        # run's record of loaded native libraries shows oracle/_ref/*.so directly
        O.ref()
        loaded.append("oracle/_ref/libporeover_ref.so")
        if kind == "pair":
            O.ref_align_module()
            loaded.append("oracle/_ref/align*.so")
    cores = cores or os.cpu_count() or 1
    try:
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available // (768 << 20))))
    except ImportError:
        pass
    _REF_DATA = (kind, data, beam_width)
    ctx = mp.get_context("fork")
    times, bases = [], 0
    with ctx.Pool(processes=cores) as pool:
        for s in range(warmup + steps):
            idx = [(s * n_items + j) % n_data for j in range(n_items)]
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, idx, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                bases += sum(r[0] for r in res)
    total = sum(times)
    ref_core = O.have_ref() and kind in ("pair", "beam")
    return {"items_per_s": n_items * len(times) / total, "bases_per_s": bases / total, "cores": cores,
            "kind": "reference" if ref_core else "port",
            "kind_detail": ("reference C++ search core + Cython banded aligner compiled unmodified (oracle/_ref), numpy / "
                            "Python glue between them from the oracle port" if ref_core else
                            "plain-C oracle port of the reference's numpy / Python code (oracle/poreover_oracle.c)"),
            "native_loaded_in_parent": loaded, "ms_per_step": 1e3 * total / len(times),
            "sample": "%d items per step x %d steps (same synthetic inputs as the GPU arm)" % (n_items, len(times))}


def cpu_reference_python(n_pairs, T, beam_width, cores=None):
    """BASELINE.md section 3, literally: the UNMODIFIED reference package's own driver, pair_decode.pair_decode(Namespace(
    ..., threads=cores)) with its multiprocessing.Pool, on .npy files of the same synthetic pairs -- run by
    oracle/ref_python_driver.py from the build outputs under oracle/_ref/ (the package byte-compiled into refpkg.zip, its
    three Cython extensions compiled from the sources where they lay).  One pass over n_pairs pairs; None when those
    outputs were not built."""
    import shutil
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not (os.path.exists(os.path.join(ref, "refpkg.zip")) and os.path.isdir(os.path.join(ref, "refext"))):
        return None
    cores = cores or os.cpu_count() or 1
    d = tempfile.mkdtemp(prefix="pob_refpy_")
    try:
        names = [synth.save_pair(d, k, T) for k in range(n_pairs)]
        with open(os.path.join(d, "pairs.txt"), "w") as f:
            for a, b in names:
                f.write("%s %s\n" % (a, b))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_python_driver.py"), os.path.join(d, "pairs.txt"),
                            d, os.path.join(d, "ref"), str(cores), str(beam_width)], capture_output=True, text=True,
                           timeout=1800)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"error": (r.stderr or r.stdout)[-300:]}
        res = json.loads(line[-1])
        return {"value": n_pairs / res["seconds"], "unit": "pairs/s", "cores": cores, "kind": "reference",
                "kind_detail": "the unmodified reference package end to end: poreover.decoding.pair_decode.pair_decode("
                               "Namespace(..., threads=%d)) on .npy files, its own multiprocessing.Pool, loaders, numpy "
                               "Viterbi, Cython aligner and C++ search (oracle/_ref/refpkg.zip + refext/)" % cores,
                "sample": "%d pairs, one pass (pairs 0..%d of the GPU arm's synthetic set, T~%d, beam %d)"
                          % (n_pairs, n_pairs - 1, T, beam_width),
                "consensus_mbases_per_s": res["consensus_bases"] / res["seconds"] / 1e6, "seconds": res["seconds"]}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# --------------------------------------------------------------------------------------- workloads
CONFIGS = {
    "cfg3_pairs": "BASELINE configs[2]: pair-decode of 10k synthetic bonito pairs (T~5000, beam_width 25, --reverse_complement, "
                  "padding 5, banded NW 500, row_col), one job sharded by pair over the GPUs",
    "cfg2_viterbi": "BASELINE configs[1]: best-path (Viterbi) decode + base->timestep mapping of 10k synthetic T=5000 reads",
    "cfg2_beam25": "BASELINE configs[1]: single-read CTC prefix beam search, beam_width 25, over 10k synthetic T=5000 reads",
    "cfg2_beam100": "BASELINE configs[1]: single-read CTC prefix beam search, beam_width 100, over 10k synthetic T=5000 reads",
    "cfg4_long": "BASELINE configs[3]: pair-decode of 1k synthetic pairs at T=50k-100k, padding 150 (wide envelopes), beam_width 25",
    "cfg5_flipflop": "BASELINE configs[4]: flip-flop trace Viterbi (8 states, 40 transitions) over 10k synthetic T=5000 uint8 traces",
}


def workload_config(args, extra=None):
    c = {"workload": CONFIGS[args.config], "name": args.config, "T": args.T, "beam_width": args.beam_width,
         "sharding": "by pair / read over a host work queue (TCPStore fetch-and-add), no collective"}
    if extra:
        c.update(extra)
    return c


def view(rt, lo, hi):
    """ReadsT of the slice [lo, hi) of a packed batch (host or device pointers): same data, offset index arrays."""
    from poreover_b200._lib import ReadsT
    return ReadsT(rt.data, rt.row_off + 8 * lo, (rt.row_len + 4 * lo) if rt.row_len else None,
                  (rt.rc + lo) if rt.rc else None, hi - lo, rt.n_states, rt.dtype, rt.layout)


def beam_roofline(prof, counters, n_pairs_per_launch, launches, clocks, alg_bytes):
    """Frames of the dominant kernel (the pair beam search): forward cell updates/s and its share of the step from
    this run; instruction-issue, FP64-pipe and shared-memory fractions, warp instructions per cell update and the
    DRAM : algorithmic ratio from the committed ncu capture of the same kernel, scaled by this run's kernel time."""
    hbm, hbm_src = measured_peaks()
    bk = prof.get("beam_pair", {"ms": 0.0, "launches": 1})
    tot_ms = sum(v["ms"] for v in prof.values())
    beam_ms = bk["ms"] / max(1, launches)  # per step
    ns, ns_src = ncu_summary()
    nb = ns.get("beam_kernel", {})
    sm_mhz = float((clocks or {}).get("sm_mhz") or 0.0) or 1965.0
    peak_issue = 148 * 4 * sm_mhz * 1e6
    roof = {"kernel": "beam_kernel (pair search, %.1f %% of the step's kernel time)" % (100.0 * bk["ms"] / tot_ms if tot_ms else 0),
            "bound": "issue", "unit": "G warp-instructions/s", "peak": peak_issue / 1e9, "achieved": None, "frac": None,
            "traffic": None, "peak_source": "148 SMs x 4 schedulers x SM clock under load (%.0f MHz); the kernel is neither "
                                            "HBM- nor tensor-bound: FP64 forward cells + shared-memory bookkeeping" % sm_mhz,
            "cell_updates_per_step": counters["cell_updates"], "ms_per_step": beam_ms,
            "cell_updates_per_s": counters["cell_updates"] / max(1e-9, beam_ms / 1e3),
            "share_of_step": bk["ms"] / tot_ms if tot_ms else None}
    if nb and nb.get("grid"):
        # the capture is one wave of T=5000 pairs: scale it by forward cell updates (the unit of work both runs count),
        # so that other read lengths and envelope widths use the same per-update figures
        cap_upd = float(nb.get("cell_updates") or 0.0)
        scale = (counters["cell_updates"] / cap_upd) if cap_upd else n_pairs_per_launch / float(nb["grid"])
        inst = float(nb["warp_instructions"]) * scale
        dram = (nb["dram_read"] + nb["dram_write"]) * 1e9 * scale
        issued = inst / max(1e-9, beam_ms / 1e3)
        roof.update({
            "achieved": issued / 1e9, "frac": issued / peak_issue,
            "traffic": dram,
            "warp_instructions_per_cell_update": inst / max(1.0, counters["cell_updates"]),
            "dram_to_algorithmic": dram / max(1.0, alg_bytes),
            "pipe_fp64_frac": nb.get("pipe_fp64_pct", 0) / 100.0, "pipe_fma_frac": nb.get("pipe_fma_pct", 0) / 100.0,
            "pipe_alu_frac": nb.get("pipe_alu_pct", 0) / 100.0, "pipe_xu_frac": nb.get("pipe_xu_pct", 0) / 100.0,
            "smem_wavefront_frac": nb.get("smem_wavefronts_pct", 0) / 100.0,
            "source": "%s: smsp__inst_executed.sum, dram__bytes_{read,write}.sum and pipe utilisations of one ncu --set full "
                      "capture (%d pairs, %.3g cell updates), scaled per %s to this step / this run's kernel time"
                      % (ns_src, int(nb["grid"]), cap_upd, "cell update" if cap_upd else "pair")})
    roof["hbm_frame"] = {"algorithmic_bytes_per_step": alg_bytes, "achieved": alg_bytes / max(1e-9, beam_ms / 1e3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": alg_bytes / max(1e-9, beam_ms / 1e3) / 1e9 / hbm,
                         "peak_source": hbm_src}
    return roof


def viterbi_roofline(ctx, L, reads, n_reads, T):
    """HBM roofline of the Viterbi kernel on n_reads reads (north star: >= 60 % of the HBM peak)."""
    from poreover_b200 import _lib, batch
    from poreover_b200._lib import ReadsT, check
    hbm, hbm_src = measured_peaks()
    reps = max(1, int(np.ceil(n_reads / float(len(reads)))))
    vb = batch.ReadBatch((reads * reps)[:n_reads])
    dv = ReadsT(ctx.to_device(vb.data), ctx.to_device(vb.row_off), ctx.to_device(vb.lens), None, vb.n, vb.n_states,
                vb.dtype, vb.layout)
    rows = vb.total_rows
    v_seq, v_s2s = ctx.malloc(rows + 64), ctx.malloc(4 * rows + 64)
    v_len, v_st = ctx.malloc(4 * vb.n + 64), ctx.malloc(4 * vb.n + 64)

    def vit():
        check(L.pob_viterbi(ctx.h, _lib.DEVICE, C.byref(dv), _lib.KIND["bonito"], v_seq, v_s2s, None, v_len, v_st), "pob_viterbi")

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
    ns, ns_src = ncu_summary()
    traffic = None
    try:
        if vb.n == 10000 and T == 5000:
            traffic = ns["viterbi5_f32_kernel"]["dram_traffic_bytes_per_launch"]
    except KeyError:
        pass
    for p in (v_seq, v_s2s, v_len, v_st, dv.data, dv.row_off, dv.row_len):
        ctx.free(p)
    return {"kernel": "viterbi5_f32_kernel (%d reads, T=%d, %.2f GB in)" % (vb.n, T, float(vb.lens.sum()) * 20 / 1e9),
            "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
            "traffic_source": (ns_src + " (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)") if traffic else None,
            "algorithmic_bytes_per_launch": alg_bytes, "peak_source": hbm_src, "ms_per_launch": t_ms,
            "reads_per_s": vb.n / (t_ms / 1e3)}, lens


# --------------------------------------------------------------------------------------- our arm, configs[2]
def run_pairs(args):
    rank, world, local = dist_env()
    from poreover_b200 import _lib, batch, multigpu
    from poreover_b200._lib import ReadsT, check, lib, ptr

    use_dist = world > 1
    store = host_group = None
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")  # host-side gather of the records; the data path has no collective
        store = dist.distributed_c10d._get_default_store()
    L = lib()
    lanes = [_lib.get_ctx(local)] + [_lib.Context(local) for _ in range(max(1, args.lanes) - 1)]
    ctx = lanes[0]
    G = args.pairs
    pad = 150 if args.config == "cfg4_long" else 5
    if args.config == "cfg4_long":
        rng = np.random.default_rng(7)
        l1, l2 = [], []
        uniq = min(G, args.unique_pairs)
        for k in range(uniq):
            T = int(rng.integers(50000, 100001))
            p1, p2, _ = synth.make_pair(5000 + k, T)
            l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
        reps = -(-G // uniq)
        l1, l2 = (l1 * reps)[:G], (l2 * reps)[:G]
    else:
        l1, l2 = make_pairs(0, G, args.T, args.unique_pairs)  # the same job on every rank
    b1 = batch.ReadBatch(l1)
    b2 = batch.ReadBatch(l2, rc=np.ones(G, dtype=np.uint8))
    rows1, rows2 = b1.total_rows, b2.total_rows
    kind, method = _lib.KIND["bonito"], _lib.METHOD["row_col"]
    # chunks: `--chunks-per-rank` per rank, so that two GPU calls in flight per rank leave every rank several pulls
    chunk = max(1, -(-G // (world * args.chunks_per_rank)))

    def barrier():
        for c in lanes:
            c.sync()
        if use_dist:
            dist.barrier()

    def allmax(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    step_no = [0]
    store_lock = threading.Lock()  # one store round trip at a time per process (queue pulls and record hand-over)
    last_bounds = [None]

    def drain(decode_chunk):
        """One step of the job: this rank's share of the queue.  Returns the chunks it decoded."""
        step_no[0] += 1
        q = multigpu.WorkQueue(G, chunk, store, key="pob_bench_q%d" % step_no[0],
                               pullers=(world * len(lanes)) if (use_dist and args.taper) else 0, min_chunk=args.min_chunk)
        last_bounds[0] = q.bounds
        return multigpu.drain_queue(q, decode_chunk, len(lanes), lock=store_lock)

    # ---------------- device-resident leg: the job's inputs already in HBM when the timed region starts
    def dev_reads(b):
        return ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens),
                      ctx.to_device(b.rc) if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)

    d1, d2 = dev_reads(b1), dev_reads(b2)
    o_seq1, o_seq2 = ctx.malloc(rows1 + 64), ctx.malloc(rows2 + 64)
    o_cons = ctx.malloc(rows1 + rows2 + 64)
    o_l1, o_l2, o_lc, o_st = (ctx.malloc(4 * G + 64) for _ in range(4))
    o_sc, o_stats = ctx.malloc(8 * G + 64), ctx.malloc(16 * G + 64)

    def decode_device(lo, hi, lane):
        v1, v2 = view(d1, lo, hi), view(d2, lo, hi)
        check(L.pob_pair_decode(lanes[lane].h, _lib.DEVICE, C.byref(v1), C.byref(v2), kind, args.beam_width, pad, 500,
                                method, o_seq1, o_l1 + 4 * lo, o_seq2, o_l2 + 4 * lo, o_cons, o_lc + 4 * lo,
                                o_sc + 8 * lo, o_stats + 16 * lo, o_st + 4 * lo), "pob_pair_decode")
        return None

    for _ in range(args.warmup):
        drain(decode_device)
        barrier()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.timer_start()
    mine = []
    for _ in range(args.steps):
        mine = drain(decode_device)
        if use_dist:
            for c in lanes:
                c.sync()
            dist.barrier(group=host_group)  # a step is the whole job: nobody starts the next one early
    for c in lanes:
        c.sync()
    ms = ctx.timer_stop()  # stop event recorded after every lane's stream has drained
    barrier()
    clocks = sampler.stop()
    counters = {"cell_updates": 0, "steps": 0, "launches": 0}
    ms_max = allmax(ms)
    pairs_mine = sum(hi - lo for lo, hi, _ in mine)
    # per-kernel times and forward cell updates of this rank's share of the last step: one more UN-TIMED pass over the
    # same chunks, one call at a time on one stream (with two calls in flight the kernels of one wait for the SMs the
    # other holds, and their CUDA-event times overlap)
    cell_updates, launches_pass = 0, 0
    heavy = args.config == "cfg4_long"  # long pairs: profile one chunk and scale (a pass takes minutes)
    ctx.profile(True)
    ctx.profile_reset()
    done = 0
    for lo, hi, _ in (mine[:1] if heavy else mine):
        decode_device(lo, hi, 0)
        ctx.sync()
        cell_updates += ctx.counters()["cell_updates"]
        done += hi - lo
    prof = ctx.profile_get()
    ctx.profile(False)
    scale_up = pairs_mine / float(done) if done else 1.0
    for v in prof.values():
        launches_pass += v["launches"]
        v["ms"] *= scale_up
    counters["cell_updates"] = int(cell_updates * scale_up)
    launches_step = int(launches_pass * scale_up)
    cons_len = ctx.from_device(o_lc, (G,), np.int32)
    status = ctx.from_device(o_st, (G,), np.int32)
    value = G * args.steps / (ms_max / 1e3)
    # every pair decoded exactly once, and to the same consensus length as a single-GPU pass over the whole job
    check_rec = None
    one_gpu = None
    lens_mine = np.full(G, -1, np.int64)
    for lo, hi, _ in mine:
        lens_mine[lo:hi] = cons_len[lo:hi]
    if use_dist:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(lens_mine, gathered, dst=0, group=host_group)
    else:
        gathered = [lens_mine]
    if rank == 0:
        cover = np.stack(gathered)
        once = bool(((cover >= 0).sum(axis=0) == 1).all())
        job = cover.max(axis=0)
        if args.config == "cfg4_long" and not use_dist:
            one_gpu = job.astype(np.int32)  # long pairs: a pass takes minutes; the job above WAS a single-GPU pass
        else:
            whole = view(d1, 0, G), view(d2, 0, G)
            check(L.pob_pair_decode(ctx.h, _lib.DEVICE, C.byref(whole[0]), C.byref(whole[1]), kind, args.beam_width, pad,
                                    500, method, o_seq1, o_l1, o_seq2, o_l2, o_cons, o_lc, o_sc, o_stats, o_st), "pob_pair_decode")
            one_gpu = ctx.from_device(o_lc, (G,), np.int32)
        check_rec = {"every_pair_decoded_once": once, "consensus_lengths_equal_single_gpu_pass": bool(np.array_equal(job, one_gpu)),
                     "pairs": int(G), "consensus_bases": int(one_gpu.sum())}
        bases_job = int(one_gpu.sum())
    else:
        bases_job = 0
    mbases = allsum(float(bases_job)) * args.steps / (ms_max / 1e3) / 1e6
    barrier()

    # ---------------- end-to-end leg: the call a user makes, HOST buffers (pinned), copies inside the timed region
    def pinned_like(a):
        p = _lib.vp()
        check(L.pob_malloc_host(max(a.nbytes, 1) + 64, C.byref(p)), "pob_malloc_host")
        C.memmove(p.value, a.ctypes.data, a.nbytes)
        return p.value

    def pinned_zeros(shape, dtype):
        """numpy view of freshly pinned host memory (cudaMallocHost): copies to and from it are asynchronous"""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _lib.vp()
        check(L.pob_malloc_host(max(nbytes, 1) + 64, C.byref(p)), "pob_malloc_host")
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        a[...] = 0
        return a

    def host_reads(b):
        # the whole descriptor in pinned memory: probabilities and the (small) index arrays
        return ReadsT(pinned_like(b.data), pinned_like(b.row_off), pinned_like(b.lens),
                      pinned_like(b.rc) if b.rc is not None else None, b.n, b.n_states, b.dtype, b.layout)

    h1, h2 = host_reads(b1), host_reads(b2)
    pin_out = not args.pageable_outputs
    zeros = pinned_zeros if pin_out else np.zeros
    hs1, hs2 = zeros(rows1 + 64, np.uint8), zeros(rows2 + 64, np.uint8)
    hc = zeros(rows1 + rows2 + 64, np.uint8)
    hl1, hl2, hlc, hst = (zeros(G, np.int32) for _ in range(4))
    hsc, hstats = zeros(G, np.float64), zeros((G, 4), np.int32)
    A = lambda a: a.ctypes.data  # noqa: E731
    coff = b1.row_off + b2.row_off
    bytes_io = [0, 0]

    def decode_host(lo, hi, lane):
        v1, v2 = view(h1, lo, hi), view(h2, lo, hi)
        check(L.pob_pair_decode(lanes[lane].h, _lib.HOST, C.byref(v1), C.byref(v2), kind, args.beam_width, pad, 500, method,
                                A(hs1), A(hl1) + 4 * lo, A(hs2), A(hl2) + 4 * lo, A(hc), A(hlc) + 4 * lo,
                                A(hsc) + 8 * lo, A(hstats) + 16 * lo, A(hst) + 4 * lo), "pob_pair_decode(host)")
        # the chunk's records: consensus strings + lengths + status, as the command line would hand them to its writer.
        # Other ranks hand them to rank 0 through the host store as soon as the chunk is done (bytes under a per-chunk
        # key), so the hand-over overlaps the decoding of the next chunks; rank 0 collects them after its own share.
        c0, c1 = int(coff[lo]), int(coff[hi])
        rec = (hlc[lo:hi].copy(), hst[lo:hi].copy(), hc[c0:c1].tobytes())
        if use_dist and rank != 0:
            blob = rec[0].tobytes() + rec[1].tobytes() + rec[2]
            with store_lock:
                store.set("pob_rec_%d_%d" % (step_no[0], lo), blob)
            return None
        return rec

    e2e_parts = {"drain_s": 0.0, "gather_s": 0.0}

    def e2e_step():
        ta = time.perf_counter()
        got = drain(decode_host)
        tb = time.perf_counter()
        e2e_parts["drain_s"] += tb - ta
        if use_dist and rank == 0:
            # results gathered on the host: every chunk rank 0 did not decode itself arrives through the store
            have = {lo for lo, _, _ in got}
            bounds = last_bounds[0]
            for k in range(len(bounds) - 1):
                lo, hi = bounds[k], bounds[k + 1]
                if lo in have:
                    continue
                key = "pob_rec_%d_%d" % (step_no[0], lo)
                blob = store.get(key)  # blocks until the owning rank has set it
                store.delete_key(key)
                m = 4 * (hi - lo)
                got.append((lo, hi, (np.frombuffer(blob[:m], np.int32), np.frombuffer(blob[m:2 * m], np.int32), blob[2 * m:])))
            e2e_parts["gather_s"] += time.perf_counter() - tb
        return [got]

    for _ in range(0 if heavy else max(1, min(args.warmup, 2))):
        e2e_step()
        if use_dist:
            dist.barrier(group=host_group)
    barrier()
    e2e_parts["drain_s"] = e2e_parts["gather_s"] = 0.0
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = e2e_step()
        if use_dist:
            dist.barrier(group=host_group)
    e2e_s = allmax(time.perf_counter() - t0)
    barrier()
    e2e_ok = None
    if rank == 0:
        lens = np.full(G, -1, np.int64)
        for part in last:
            for lo, hi, (cl, _st, _blob) in part:
                lens[lo:hi] = cl
        e2e_ok = bool(np.array_equal(lens, one_gpu))
    row_bytes = 20  # 5 float32 per timestep
    h2d = int((rows1 + rows2) * row_bytes + 2 * (b1.row_off.nbytes + b1.lens.nbytes) + G)
    d2h = int(rows1 + rows2 + (rows1 + rows2) + G * (4 * 4 + 8))
    e2e = {"value": G * args.steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * e2e_s / args.steps, "records_gathered_on_rank0_inside": bool(use_dist),
           "host_buffers": "inputs pinned (cudaMallocHost); outputs " + ("pinned" if pin_out else "pageable"),
           "rank0_ms_per_step": {"decode_own_chunks": 1e3 * e2e_parts["drain_s"] / args.steps,
                                 "gather_records": 1e3 * e2e_parts["gather_s"] / args.steps},
           "consensus_lengths_equal_single_gpu_pass": e2e_ok}

    # ---------------- N > 1: the replica (weak-scaling) measurement of round 1 next to the strong one
    weak = None
    if use_dist and not args.no_weak:
        whole = view(d1, 0, G), view(d2, 0, G)

        def replica():
            check(L.pob_pair_decode(ctx.h, _lib.DEVICE, C.byref(whole[0]), C.byref(whole[1]), kind, args.beam_width, pad,
                                    500, method, o_seq1, o_l1, o_seq2, o_l2, o_cons, o_lc, o_sc, o_stats, o_st), "pob_pair_decode")

        replica()
        barrier()
        ctx.timer_start()
        wsteps = min(args.steps, 3)
        for _ in range(wsteps):
            replica()
        wms = allmax(ctx.timer_stop())
        barrier()
        weak = {"value": world * G * wsteps / (wms / 1e3), "unit": "pairs/s", "scaling": "weak", "steps": wsteps,
                "pairs_per_gpu_per_step": G, "note": "every GPU decodes its own copy of the 10k-pair job (replicas, as in round 1)"}

    # ---------------- rooflines: the dominant kernel (beam search) and the HBM-bound Viterbi kernel
    roof = roof_v = None
    if rank == 0:
        alg = 20.0 * (rows1 + rows2) + 8.0 * rows1 + float(bases_job)  # 20 (U+V) in, 8 U envelope, consensus out
        share = pairs_mine / float(G) if G else 1.0
        roof = beam_roofline(prof, counters, pairs_mine, 1, clocks, alg * share)
        if args.config == "cfg3_pairs":
            roof_v, _ = viterbi_roofline(ctx, L, l1[:min(G, 2500)], args.viterbi_reads, args.T)

    cpu = cpu_py = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == "cfg4_long":
        cpu = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
               "sample": "not timed in this run: one T=50k-100k pair with padding 150 at beam width 25 takes the reference "
                         "core several minutes on one core; tests/test_gpu_configs.py times and checks two such pairs at "
                         "beam width 5 (about 100 s of CPU for both)"}
    elif rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        npairs = min(G, max(4 * cores, 4)) if args.config == "cfg3_pairs" else min(G, max(cores // 4, 2))
        nd = min(G, 256)
        r = cpu_reference("pair", (l1[:nd], l2[:nd], pad), args.beam_width, npairs, 1, 0, nd)
        cpu = {"value": r["items_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
               "kind_detail": r["kind_detail"], "sample": r["sample"], "consensus_mbases_per_s": r["bases_per_s"] / 1e6}
        if args.config == "cfg3_pairs":
            cpu_py = cpu_reference_python(min(G, 4 * cores), args.T, args.beam_width, cores)

    if rank == 0:
        line = {
            "metric": "pair_decode_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {
                "pairs_per_step_whole_job": G, "unique_pairs": min(G, args.unique_pairs), "chunk_pairs": chunk,
                "chunks_taper_to": (args.min_chunk if (use_dist and args.taper) else None),
                "gpu_calls_in_flight_per_rank": len(lanes), "padding": pad,
                "cache": "inputs per step (%.0f MB) exceed the 126 MB L2" % ((rows1 + rows2) * 20 / 1e6)}),
            "consensus_mbases_per_s": mbases, "e2e": e2e, "gpu_launches": int(launches_step * args.steps), "clocks": clocks,
            "roofline": roof, "roofline_viterbi": roof_v, "cpu_baseline": cpu, "cpu_baseline_python": cpu_py, "weak": weak,
            "check": check_rec,
            "kernels_ms_per_step_rank0": {k: v["ms"] for k, v in prof.items()},
            "kernels_note": "per-kernel CUDA-event times of rank 0's share of one step, re-run one call at a time after the timed region",
            "pairs_skipped": int(((status & (16 | 32 | 8 | 64)) != 0).sum()),
            "pairs_pool_overflow": int(((status & 4) != 0).sum()),
        }
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------- our arm, single-read configs
def run_reads(args):
    """configs[1] (Viterbi, 1D beam search at widths 25 / 100) and configs[4] (flip-flop Viterbi): 10k single reads,
    sharded over the ranks in equal static slices (reads are independent; weak scaling: every rank its own `--reads`)."""
    rank, world, local = dist_env()
    from poreover_b200 import _lib, batch
    from poreover_b200._lib import ReadsT, check, lib, ptr

    use_dist = world > 1
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = lib()
    ctx = _lib.get_ctx(local)
    n = args.reads
    uniq = min(n, args.unique_reads)
    flip = args.config == "cfg5_flipflop"
    if flip:
        data = [synth.make_flipflop_trace(rank * uniq + i, args.T) for i in range(uniq)]
    else:
        data = [synth.bonito_log_prob(synth.make_read(rank * uniq + i, args.T)[0]) for i in range(uniq)]
    reads = (data * (-(-n // uniq)))[:n]
    W = {"cfg2_beam25": 25, "cfg2_beam100": 100}.get(args.config, args.beam_width)
    b = batch.ReadBatch(reads, dtype=np.uint8 if flip else None)
    rows = b.total_rows
    dR = ReadsT(ctx.to_device(b.data), ctx.to_device(b.row_off), ctx.to_device(b.lens), None, b.n, b.n_states, b.dtype, b.layout)
    o_seq, o_s2s = ctx.malloc(rows + 64), ctx.malloc(4 * rows + 64)
    o_len, o_st, o_sc = ctx.malloc(4 * n + 64), ctx.malloc(4 * n + 64), ctx.malloc(8 * n + 64)
    o_path = ctx.malloc(rows + 64)
    lut = np.log((np.arange(256) + 1e-7) / (255 + 1e-7))  # decode.py:92-93, computed by the host
    d_lut = ctx.to_device(lut)
    hR = b.struct()
    h_seq, h_s2s = np.zeros(rows + 64, np.uint8), np.zeros(rows + 64, np.int32)
    h_len, h_st, h_sc = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
    model = _lib.MODEL["ctc_merge_repeats"]

    def call(where, R, seq, s2s, ln, st, sc, path, lutp):
        if args.config == "cfg2_viterbi":
            check(L.pob_viterbi(ctx.h, where, C.byref(R), _lib.KIND["bonito"], seq, s2s, None, ln, st), "pob_viterbi")
        elif flip:
            check(L.pob_viterbi_flipflop(ctx.h, where, C.byref(R), lutp, seq, s2s, path, ln), "pob_viterbi_flipflop")
        else:
            check(L.pob_beam_search(ctx.h, where, C.byref(R), W, model, seq, ln, sc, st), "pob_beam_search")

    def step_device():
        call(_lib.DEVICE, dR, o_seq, o_s2s, o_len, o_st, o_sc, o_path, d_lut)

    def step_host():
        call(_lib.HOST, hR, ptr(h_seq), ptr(h_s2s), ptr(h_len), ptr(h_st), ptr(h_sc), None, ptr(lut))

    def barrier():
        ctx.sync()
        if use_dist:
            dist.barrier()

    def allmax(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.profile(True); ctx.profile_reset()
    sampler = ClockSampler(local); sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = allmax(ctx.timer_stop())
    barrier()
    clocks = sampler.stop()
    prof = ctx.profile_get()
    counters = ctx.counters()
    ctx.profile(False)
    lens = ctx.from_device(o_len, (n,), np.int32)
    value = world * n * args.steps / (ms / 1e3)
    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    ctx.sync()
    e2e_s = allmax(time.perf_counter() - t0)
    assert np.array_equal(h_len, lens), "host and device legs disagree"
    esz = 1 if flip else 4
    e2e = {"value": world * n * args.steps / e2e_s, "unit": "reads/s", "ms_per_step": 1e3 * e2e_s / args.steps,
           "h2d_bytes_per_step": int(rows * b.n_states * esz + b.row_off.nbytes + b.lens.nbytes),
           "d2h_bytes_per_step": int(rows * (5 if args.config != "cfg2_beam25" and args.config != "cfg2_beam100" else 1) + n * 16)}
    hbm, hbm_src = measured_peaks()
    kname = {"cfg2_viterbi": "viterbi_ctc", "cfg5_flipflop": "viterbi_flipflop"}.get(args.config, "beam_single")
    k_ms = prof.get(kname, {"ms": 0.0})["ms"] / args.steps
    in_bytes = float(b.lens.sum()) * b.n_states * esz
    alg = in_bytes + float(lens.sum()) * (5 if kname != "beam_single" else 1) + n * 8
    roof = {"kernel": kname, "bound": "hbm", "achieved": alg / max(1e-9, k_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": alg / max(1e-9, k_ms / 1e3) / 1e9 / hbm, "traffic": None, "algorithmic_bytes_per_launch": alg,
            "peak_source": hbm_src, "ms_per_launch": k_ms}
    if kname == "beam_single":
        roof["note"] = ("the single-read search is bound by its dependent per-timestep chain (one CTA per read), not by HBM; "
                        "cell updates/s: %.3g" % (counters["cell_updates"] / max(1e-9, k_ms / 1e3)))
    if kname == "viterbi_flipflop":
        roof["note"] = "bound by the FP64 dependent chain over T (8 states per read), not by HBM"
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        kind = {"cfg2_viterbi": "viterbi", "cfg5_flipflop": "flipflop"}.get(args.config, "beam")
        per = {"viterbi": 64 * cores, "flipflop": 8 * cores, "beam": (2 if W <= 25 else 1) * cores}[kind]
        nd = min(uniq, max(per, 8))
        r = cpu_reference(kind, data[:nd], W, min(per, n), 1, 0, nd)
        cpu = {"value": r["items_per_s"], "unit": "reads/s", "cores": r["cores"], "kind": r["kind"],
               "kind_detail": r["kind_detail"], "sample": r["sample"]}
    if rank == 0:
        line = {"metric": args.config + "_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64" if kname != "viterbi_ctc" else "f32", "data": "synthetic",
                "config": workload_config(args, {"reads_per_gpu_per_step": n, "unique_reads": uniq, "beam_width": W,
                                                 "cache": "inputs per step (%.0f MB) exceed the 126 MB L2" % (in_bytes / 1e6)}),
                "e2e": e2e, "gpu_launches": int(sum(v["launches"] for v in prof.values())), "clocks": clocks, "roofline": roof,
                "cpu_baseline": cpu, "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
                "consensus_mbases_per_s": float(lens.sum()) * world * args.steps / (ms / 1e3) / 1e6}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu_py = None
    if args.config in ("cfg3_pairs", "cfg4_long"):
        n_items = args.ref_pairs or (4 * cores if args.config == "cfg3_pairs" else max(cores // 4, 2))
        uniq = min(max(n_items, 8), 256)
        if args.config == "cfg4_long":
            rng = np.random.default_rng(7)
            l1, l2 = [], []
            for k in range(min(uniq, n_items)):
                T = int(rng.integers(50000, 100001))
                p1, p2, _ = synth.make_pair(5000 + k, T)
                l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
            pad = 150
        else:
            l1, l2 = make_pairs(0, uniq, args.T)
            pad = 5
        r = cpu_reference("pair", (l1, l2, pad), args.beam_width, n_items, args.steps, args.warmup, len(l1))
        metric, unit = "pair_decode_pairs_per_s", "pairs/s"
        if args.config == "cfg3_pairs":
            cpu_py = cpu_reference_python(n_items, args.T, args.beam_width, cores)
    else:
        W = {"cfg2_beam25": 25, "cfg2_beam100": 100}.get(args.config, args.beam_width)
        kind = {"cfg2_viterbi": "viterbi", "cfg5_flipflop": "flipflop"}.get(args.config, "beam")
        n_items = args.ref_pairs or {"viterbi": 64 * cores, "flipflop": 8 * cores, "beam": (2 if W <= 25 else 1) * cores}[kind]
        nd = min(n_items, 512)
        if kind == "flipflop":
            data = [synth.make_flipflop_trace(i, args.T) for i in range(nd)]
        else:
            data = [synth.bonito_log_prob(synth.make_read(i, args.T)[0]) for i in range(nd)]
        r = cpu_reference(kind, data, W, n_items, args.steps, args.warmup, nd)
        metric, unit = args.config + "_reads_per_s", "reads/s"
    line = {
        "impl": "reference", "metric": metric, "value": r["items_per_s"], "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if args.config == "cfg3_pairs" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "consensus_mbases_per_s": r["bases_per_s"] / 1e6,
        "config": workload_config(args),  # the arm's workload; this run's bounded sample is in cpu_baseline
        "cpu_baseline": {"value": r["items_per_s"], "unit": unit, "cores": r["cores"], "kind": r["kind"],
                         "kind_detail": r["kind_detail"], "native_loaded_in_parent": r["native_loaded_in_parent"],
                         "sample": r["sample"]},
        "cpu_baseline_python": cpu_py,
        "e2e": {"value": r["items_per_s"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3_pairs", choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs of the whole job (default 10000; cfg4_long: 1000)")
    ap.add_argument("--pairs-per-gpu", type=int, default=0, help="(round 1 name) same as --pairs")
    ap.add_argument("--unique-pairs", type=int, default=2500, help="distinct synthetic pairs (tiled to the job)")
    ap.add_argument("--reads", type=int, default=10000, help="single-read configs: reads per GPU per step")
    ap.add_argument("--unique-reads", type=int, default=1000)
    ap.add_argument("--chunks-per-rank", type=int, default=4, help="the job is cut into world x this many chunks")
    ap.add_argument("--lanes", type=int, default=2, help="GPU calls in flight per rank (contexts / streams)")
    ap.add_argument("--T", type=int, default=5000)
    ap.add_argument("--beam-width", type=int, default=25)
    ap.add_argument("--viterbi-reads", type=int, default=10000)
    ap.add_argument("--ref-pairs", type=int, default=0, help="items per step of the reference arm (default: by config and cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--taper", action="store_true", help="N > 1: chunks that shrink towards the end of the job (guided self-scheduling); "
                    "measured slower (9.2k vs 10.9k pairs/s at N=2, 313-pair chunks): chunks below the ~444 resident pairs under-fill a GPU")
    ap.add_argument("--min-chunk", type=int, default=96, help="smallest chunk of the tapered queue")
    ap.add_argument("--pageable-outputs", action="store_true", help="e2e leg: results into pageable instead of pinned host arrays")
    args = ap.parse_args()
    if args.pairs_per_gpu and not args.pairs:
        args.pairs = args.pairs_per_gpu
    if not args.pairs:
        args.pairs = 1000 if args.config == "cfg4_long" else 10000
    if args.config == "cfg4_long":
        args.unique_pairs = min(args.unique_pairs, 64)
        args.lanes = 1  # the node pools of wide-band pairs take tens of GB per context: one GPU call at a time
    if args.warmup < 3 and args.impl == "ours" and args.config != "cfg4_long":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.config in ("cfg3_pairs", "cfg4_long"):
        run_pairs(args)
    else:
        run_reads(args)


if __name__ == "__main__":
    main()
