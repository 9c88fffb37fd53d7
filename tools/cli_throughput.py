"""End-to-end throughput of the `pair-decode` command line path: .npy files on disk -> FASTA / log files.

    python tools/cli_throughput.py [--pairs 8192] [--unique 512] [--T 5000] [--beam_width 25] [--out result.json]

Writes `unique` synthetic pairs (SURVEY.md 8(d) recipe) to a scratch directory, lists them `pairs` times over in a
pairs file, and times (wall clock, one process, one GPU) what `python -m poreover_b200 pair-decode` does with it:
poreover_b200.multigpu.decode_pairs_all_gpus + pair_decode.write_results.  Next to the whole-run figure it reports
the two pipeline stages on their own -- the host stage (file reads + numpy log into packed batches) and the GPU stage
(pob_pair_decode + result formatting) -- so that one can see which of them the run waits for.
"""
import argparse
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

# POB_TREE=<dir> times another checkout of the package (e.g. an older commit unpacked under build/)
sys.path.insert(0, os.environ.get("POB_TREE") or os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def _gen(job):
    from poreover_b200 import synth
    d, k, T = job
    return synth.save_pair(d, k, T)


def decode_mode(a, d, names):
    """`python -m poreover_b200 decode DIR --basecaller bonito --algorithm ...` on a directory of 2 x pairs reads
    (symbolic links to the unique files): multigpu.decode_files_all_gpus + the FASTA output of decode.decode."""
    from pathlib import Path

    from poreover_b200 import multigpu
    from poreover_b200.__main__ import build_parser
    from poreover_b200.decoding import decode as dec
    uniq = [n for pair in names for n in pair]
    rd = os.path.join(d, "reads")
    os.mkdir(rd)
    files = []
    for i in range(2 * a.pairs):
        f = os.path.join(rd, "read%06d.npy" % i)
        os.symlink(os.path.join(d, uniq[i % len(uniq)]), f)
        files.append(f)
    args = build_parser().parse_args(["decode", rd, "--basecaller", "bonito", "--algorithm", a.decode, "--beam_width",
                                      str(a.beam_width), "--out", os.path.join(d, "dec")])

    def whole_run(fs):
        seqs = multigpu.decode_files_all_gpus(args, fs)
        with open(args.out + '.fasta', 'w') as out_fasta:
            for p, sq in zip(fs, seqs):
                print(dec.fasta_format(Path(p).stem, sq), file=out_fasta)
        return seqs

    whole_run(files[:min(len(files), 4096)])
    t0 = time.perf_counter()
    seqs = whole_run(files)
    t_run = time.perf_counter() - t0
    out = {"metric": "decode_cli_reads_per_s", "algorithm": a.decode, "value": len(files) / t_run, "unit": "reads/s",
           "reads": len(files), "T": a.T, "beam_width": a.beam_width if a.decode == "beam" else None,
           "mbases_per_s": sum(len(s) for s in seqs) / t_run / 1e6, "wall_s": t_run, "host_cores": os.cpu_count(),
           "tree": os.environ.get("POB_TREE", "this checkout"),
           "what": "files on disk (page cache) -> .fasta, one process, one GPU, wall clock"}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(out) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8192)
    ap.add_argument("--unique", type=int, default=512)
    ap.add_argument("--T", type=int, default=5000)
    ap.add_argument("--beam_width", type=int, default=25)
    ap.add_argument("--out", default=None)
    ap.add_argument("--whole-only", action="store_true", help="only the whole-run figure (works on older trees)")
    ap.add_argument("--decode", choices=["viterbi", "beam"], default=None,
                    help="time the single-read `decode` command line on 2 x --pairs reads instead of pair-decode")
    a = ap.parse_args()
    from poreover_b200 import multigpu
    from poreover_b200.__main__ import build_parser
    from poreover_b200.decoding import pair_decode as pd

    d = tempfile.mkdtemp(prefix="pob_cli_")
    t0 = time.perf_counter()
    with ProcessPoolExecutor() as ex:
        names = list(ex.map(_gen, [(d, k, a.T) for k in range(a.unique)], chunksize=8))
    t_gen = time.perf_counter() - t0
    if a.decode:
        return decode_mode(a, d, names)
    pair_list = [list(names[i % a.unique]) for i in range(a.pairs)]
    with open(os.path.join(d, "pairs.txt"), "w") as f:
        for p in pair_list:
            f.write("%s %s\n" % (p[0], p[1]))
    args = build_parser().parse_args(["pair-decode", os.path.join(d, "pairs.txt"), "--dir", d, "--basecaller", "bonito",
                                      "--reverse_complement", "--beam_width", str(a.beam_width), "--out",
                                      os.path.join(d, "run")])

    def whole_run(pairs):
        res = multigpu.decode_pairs_all_gpus(args, pairs)
        with open(args.out + '.1d.fasta', 'w') as f1, open(args.out + '.2d.fasta', 'w') as f2, \
                open(args.out + '.log', 'w') as lf:
            pd.write_results(args, res, f1, f2, lf)
        return res

    whole_run(pair_list[:min(len(pair_list), 2048)])  # warm-up: context, arena growth, page cache
    t0 = time.perf_counter()
    res = whole_run(pair_list)
    t_run = time.perf_counter() - t0
    done = sum(1 for r in res if r is not None and len(r) == 3)
    bases = sum(len(r[1].split("\n", 1)[1].replace("\n", "")) for r in res if r is not None and len(r) == 3)

    if a.whole_only:
        out = {"metric": "pair_decode_cli_pairs_per_s", "value": a.pairs / t_run, "unit": "pairs/s", "pairs": a.pairs,
               "decoded": done, "wall_s": t_run, "tree": os.environ.get("POB_TREE", "this checkout")}
        print(json.dumps(out))
        if a.out:
            with open(a.out, "w") as f:
                f.write(json.dumps(out) + "\n")
        return
    from poreover_b200 import ingest
    # the stages on their own, same chunking as the run
    chunk = max(8, min(4096, -(-len(pair_list) // 4)))
    subs = [pair_list[i:i + chunk] for i in range(0, len(pair_list), chunk)]
    t0 = time.perf_counter()
    payloads = [pd.load_pairs(args, s) for s in subs[:2]]
    t_load = (time.perf_counter() - t0) / sum(len(s) for s in subs[:2])
    t0 = time.perf_counter()
    for p in payloads:
        pd.decode_loaded(args, p)
    t_gpu = (time.perf_counter() - t0) / sum(len(s) for s in subs[:2])
    # the reference-style loader (one transducer object per file, decode.py:67-112) on a sample, single thread
    from poreover_b200.decoding import decode
    sample = [os.path.join(d, n) for p in pair_list[:256] for n in p]
    t0 = time.perf_counter()
    for p in sample:
        m = decode.model_from_trace(p, "bonito")
        m.log_prob, m.transition  # what transducer.__init__ builds eagerly in the reference (transducer.py:16, :22)
    t_ref_loader = (time.perf_counter() - t0) / 256

    out = {"metric": "pair_decode_cli_pairs_per_s", "value": a.pairs / t_run, "unit": "pairs/s",
           "pairs": a.pairs, "unique_pairs": a.unique, "T": a.T, "beam_width": a.beam_width, "decoded": done,
           "consensus_mbases_per_s": bases / t_run / 1e6, "wall_s": t_run, "chunk": chunk,
           "loader_threads": ingest.n_threads(), "host_cores": os.cpu_count(),
           "pinned_batches": os.environ.get("POREOVER_B200_PINNED", "0") == "1",
           "gpu_lanes": int(os.environ.get("POREOVER_B200_GPU_LANES", "2")),
           "host_stage_pairs_per_s": 1.0 / t_load, "gpu_stage_pairs_per_s": 1.0 / t_gpu,
           "reference_style_loader_pairs_per_s_1thread": 1.0 / t_ref_loader, "synth_s": t_gen,
           "what": "files on disk (page cache) -> .1d.fasta/.2d.fasta/.log, one process, one GPU, wall clock"}
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
