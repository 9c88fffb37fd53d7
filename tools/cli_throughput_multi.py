"""End-to-end throughput of the `pair-decode` command line path on SEVERAL GPUs of one box: .npy files on disk ->
FASTA / log files through poreover_b200.multigpu.decode_pairs_all_gpus (one process per GPU under torchrun, chunks
pulled from the host work queue, records gathered on rank 0, which writes the files).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/cli_throughput_multi.py \
        [--pairs 40960] [--unique 512] [--out result.json]
"""
import argparse
import datetime
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=40960)
    ap.add_argument("--unique", type=int, default=512)
    ap.add_argument("--T", type=int, default=5000)
    ap.add_argument("--beam_width", type=int, default=25)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch.distributed as dist
    from poreover_b200 import ingest, multigpu, synth
    from poreover_b200.__main__ import build_parser
    from poreover_b200.decoding import pair_decode as pd
    dist.init_process_group("gloo", timeout=datetime.timedelta(minutes=30))
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [None]
    if rank == 0:
        box[0] = tempfile.mkdtemp(prefix="pob_cli_")
    dist.broadcast_object_list(box, src=0)
    d = box[0]
    # every rank writes its share of the unique files
    names = [None] * a.unique
    for k in range(rank, a.unique, world):
        names[k] = synth.save_pair(d, k, a.T)
    gathered = [None] * world
    dist.all_gather_object(gathered, names)
    names = [next(g[k] for g in gathered if g[k] is not None) for k in range(a.unique)]
    pair_list = [list(names[i % a.unique]) for i in range(a.pairs)]
    args = build_parser().parse_args(["pair-decode", os.path.join(d, "pairs.txt"), "--dir", d, "--basecaller", "bonito",
                                      "--reverse_complement", "--beam_width", str(a.beam_width), "--out", os.path.join(d, "run")])

    def whole_run(pairs):
        res = multigpu.decode_pairs_all_gpus(args, pairs)
        if res is not None:
            with open(args.out + '.1d.fasta', 'w') as f1, open(args.out + '.2d.fasta', 'w') as f2, open(args.out + '.log', 'w') as lf:
                pd.write_results(args, res, f1, f2, lf)
        return res

    whole_run(pair_list[:min(len(pair_list), 1024 * world)])  # warm-up: contexts, arenas, page cache
    dist.barrier()
    t0 = time.perf_counter()
    res = whole_run(pair_list)
    dist.barrier()
    t_run = time.perf_counter() - t0
    # the host stage of one rank on its own (files -> packed batches), for the record
    sub = pair_list[:2048]
    t1 = time.perf_counter()
    pd.load_pairs(args, sub)
    t_load = (time.perf_counter() - t1) / len(sub)
    loads = [None] * world
    dist.all_gather_object(loads, 1.0 / t_load)
    if rank == 0:
        done = sum(1 for r in res if r is not None and len(r) == 3)
        out = {"metric": "pair_decode_cli_pairs_per_s", "value": a.pairs / t_run, "unit": "pairs/s", "n_gpus": world,
               "pairs": a.pairs, "unique_pairs": a.unique, "T": a.T, "beam_width": a.beam_width, "decoded": done,
               "wall_s": t_run, "host_cores": os.cpu_count(), "loader_threads_per_rank": ingest.n_threads(),
               "host_stage_pairs_per_s_per_rank": loads, "host_stage_pairs_per_s_all_ranks": sum(loads),
               "what": "files on disk (page cache) -> .1d.fasta/.2d.fasta/.log, one process per GPU under torchrun, wall clock"}
        line = json.dumps(out)
        print(line)
        if a.out:
            with open(a.out, "w") as f:
                f.write(line + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
