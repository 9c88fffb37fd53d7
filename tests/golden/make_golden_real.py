"""Golden fixture from the reference's own real-data pair (data/reads/read1.npy, read2.npy: PoreOverNet logits,
62,000 / 75,600 timesteps).  Runs the REAL reference's pair_decode_helper (same scratch build as
make_golden.py) with --basecaller poreover --reverse_complement at beam widths 5 and 25 and records inputs
and outputs in tests/golden/real_pair.npz.  Run in the build container only."""
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    scratch = os.environ.get("POREOVER_REF_SCRATCH", os.path.join(tempfile.gettempdir(), "ref_scratch"))
    os.makedirs(scratch, exist_ok=True)
    mg.build_reference(scratch)
    import warnings
    warnings.simplefilter("ignore")
    from poreover.decoding import pair_decode, decode
    d = os.path.join(mg.REF, "data", "reads")
    G = {"read1": np.load(os.path.join(d, "read1.npy")), "read2": np.load(os.path.join(d, "read2.npy"))}
    for W in (5, 25):
        ns = Namespace(**{"in": ["read1.npy", "read2.npy"], "dir": d, "basecaller": "poreover", "reverse_complement": True,
                          "out": "out", "threads": 1, "method": "envelope", "single": "viterbi", "logging": "info",
                          "debug": False, "algorithm": "beam", "alignment": "banded", "beam_width": W,
                          "debug_envelope": False, "diagonal_envelope": False, "diagonal_width": 50, "padding": 5,
                          "skip_matches": False, "skip_threshold": 10, "beam_search_method": "row_col", "window": 200})
        r = pair_decode.pair_decode_helper(ns)
        assert len(r) == 3
        G["consensus_w%d" % W] = r[1].split("\n", 1)[1].replace("\n", "")
        G["identity"] = r[2]["sequence_identity"]
        G["length1"], G["length2"] = r[2]["length1"], r[2]["length2"]
        print("W", W, "consensus", len(G["consensus_w%d" % W]), "identity", G["identity"], G["length1"], G["length2"])
    m1 = decode.model_from_trace(os.path.join(d, "read1.npy"), "poreover")
    G["basecall1"] = m1.viterbi_decode()
    out = os.path.join(HERE, "real_pair.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
