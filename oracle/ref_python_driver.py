"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference package's own pair-decode driver (pair_decode.pair_decode
with its multiprocessing pool: BASELINE.md section 3) on a pairs file, from the build outputs under oracle/_ref/
(refpkg.zip: the package byte-compiled; refext/ and align*.so: its Cython extensions).  Executed as a child process by
bench.py's cpu_baseline_python leg and by tests; prints one JSON line with the wall-clock seconds of the call.

usage: ref_python_driver.py <pairs.txt> <dir> <out prefix> <threads> <beam_width> [padding]"""
import argparse
import glob
import importlib.machinery
import importlib.util
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def load_reference_package():
    sys.path.insert(0, os.path.join(REF, "refpkg.zip"))
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    import numpy as np
    if not hasattr(np, "product"):
        np.product = np.prod
    # the compiled extensions cannot live inside the zip: load them under their package names first
    exts = {"poreover.decoding.decoding_cpp": glob.glob(os.path.join(REF, "refext", "decoding_cpp*.so")),
            "poreover.decoding.decoding_cy": glob.glob(os.path.join(REF, "refext", "decoding_cy*.so")),
            "poreover.align.align": glob.glob(os.path.join(REF, "align*.so"))}
    for name, paths in exts.items():
        loader = importlib.machinery.ExtensionFileLoader(name, paths[0])
        spec = importlib.util.spec_from_loader(name, loader, origin=paths[0])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        loader.exec_module(mod)
    import poreover  # noqa: F401  (from the zip; finds the extensions in sys.modules)
    from poreover.decoding import pair_decode
    return pair_decode


if __name__ == "__main__":
    pairs, d, out, threads, W = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    padding = int(sys.argv[6]) if len(sys.argv) > 6 else 5
    pd = load_reference_package()
    # every pair-decode flag of the reference's parser with its default (__main__.py:65-91)
    ns = argparse.Namespace(**{"in": [pairs], "dir": d, "out": out, "basecaller": "bonito", "reverse_complement": True,
                               "threads": threads, "beam_width": W, "padding": padding, "alignment": "banded",
                               "single": "viterbi", "logging": "info", "algorithm": "beam", "method": "envelope",
                               "beam_search_method": "row_col", "skip_matches": False, "skip_threshold": 10,
                               "diagonal_envelope": False, "diagonal_width": 10, "window": 200, "debug": False,
                               "debug_envelope": False, "matches": 10, "indels": 100})
    t0 = time.perf_counter()
    pd.pair_decode(ns)
    dt = time.perf_counter() - t0
    bases = 0
    try:
        with open(out + ".2d.fasta") as f:
            bases = sum(len(l.strip()) for l in f if l.strip() and not l.startswith(">"))
    except OSError:
        pass
    print(json.dumps({"seconds": dt, "consensus_bases": bases}), flush=True)
