"""Where one chunk of the `pair-decode` command line spends its time: loader, pob_pair_decode through the batch API
(pageable vs pinned packed batches), record formatting.   usage: cli_stage_times.py [pairs] [unique]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from concurrent.futures import ProcessPoolExecutor
import numpy as np

def _gen(job):
    from poreover_b200 import synth
    d, k, T = job
    return synth.save_pair(d, k, T)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    uniq = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    from poreover_b200 import batch, ingest
    from poreover_b200.__main__ import build_parser
    from poreover_b200.decoding import pair_decode as pd
    d = tempfile.mkdtemp(prefix="pob_cli_")
    with ProcessPoolExecutor() as ex:
        names = list(ex.map(_gen, [(d, k, 5000) for k in range(uniq)], chunksize=8))
    pairs = [list(names[i % uniq]) for i in range(n)]
    args = build_parser().parse_args(["pair-decode", "x", "--dir", d, "--basecaller", "bonito", "--reverse_complement",
                                      "--beam_width", "25", "--out", os.path.join(d, "run")])
    for pinned in (0, 1):
        os.environ["POREOVER_B200_PINNED"] = str(pinned)
        for rep in range(3):
            t0 = time.perf_counter(); payload = pd.load_pairs(args, pairs); t1 = time.perf_counter()
            raw = pd.decode_loaded(args, payload, fmt=False); t2 = time.perf_counter()
            res = pd.format_decoded(args, raw); t3 = time.perf_counter()
            print("pinned=%d rep %d: load %.3f s (%.0f/s)  pob_pair_decode+records %.3f s (%.0f/s)  format %.3f s (%.0f/s)"
                  % (pinned, rep, t1 - t0, n / (t1 - t0), t2 - t1, n / (t2 - t1), t3 - t2, n / (t3 - t2)))
