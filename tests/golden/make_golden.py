"""Generate the committed golden fixtures by running the REAL reference (Python + Cython + C++).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It copies the reference to a scratch dir, builds its three Cython extensions with the reference's own
setup.py (five absent imports are stubbed), imports it, and records inputs and
outputs of every function on the hot path as small .npz files next to this script.  The fixtures are
what pins the oracle (oracle/poreover_oracle.c) and, through it, the CUDA path.
"""
import os
import shutil
import subprocess
import sys
import tempfile
from argparse import Namespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"

STUBS = {
    "h5py.py": "",
    "mappy.py": "",
    "Bio/__init__.py": "SeqIO = None\n",
    "progressbar.py": (
        "class _S:\n    def wrap_stderr(self): pass\nstreams = _S()\n"
        "class ProgressBar:\n    def __init__(self, max_value=None): pass\n    def update(self, *a, **k): pass\n"
        "def progressbar(it): return it\n"
    ),
    "tensorflow.py": (
        "import sys\nclass _T:\n    def __getattr__(self, k): return self\n    def __call__(self, *a, **k): return self\n"
        "sys.modules['tensorflow'] = _T()\n"
    ),
}


def build_reference(scratch):
    ref = os.path.join(scratch, "ref")
    if not os.path.exists(os.path.join(ref, "poreover", "align")) or not any(
        f.endswith(".so") for f in os.listdir(os.path.join(ref, "poreover", "align"))
    ):
        shutil.rmtree(ref, ignore_errors=True)
        shutil.copytree(REF, ref)
        subprocess.check_call(["chmod", "-R", "u+w", ref])
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=ref,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    stubs = os.path.join(scratch, "stubs")
    for name, body in STUBS.items():
        p = os.path.join(stubs, name)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(body)
    sys.path.insert(0, ref)
    sys.path.insert(0, stubs)
    np.product = np.prod  # tests/testing.py:73 uses the alias NumPy 2 removed
    return ref


def mutate(rng, s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append("ACGT"[rng.integers(0, 4)])
            continue
        if r < rate:
            out.append(ch)
            out.append("ACGT"[rng.integers(0, 4)])
            continue
        out.append(ch)
    return "".join(out)


def main():
    scratch = os.environ.get("POREOVER_REF_SCRATCH", os.path.join(tempfile.gettempdir(), "ref_scratch"))
    os.makedirs(scratch, exist_ok=True)
    ref = build_reference(scratch)
    import warnings

    warnings.simplefilter("ignore")
    from poreover.decoding import decoding_cpp, transducer, envelope, pair_decode, decode  # the reference
    import poreover.align as align
    from poreover_b200 import synth

    rng = np.random.default_rng(7)
    G = {}

    # ---------------- real-data fixture of the reference's own tests: tests/poreover.csv
    m = decode.model_from_trace(os.path.join(REF, "tests", "poreover.csv"))
    csv_lp = m.log_prob.copy()
    G["csv_log_prob"] = csv_lp
    seq, path = m.viterbi_decode(return_path=True)
    G["csv_viterbi_seq"] = seq
    G["csv_viterbi_path"] = np.asarray(path, dtype=np.int64)
    G["csv_s2s"] = np.asarray(pair_decode.get_sequence_mapping(path, "poreover")[0], dtype=np.int64)
    T = len(csv_lp)
    for W in (10, 25):
        G["csv_beam1d_w%d" % W] = decoding_cpp.cpp_beam_search(csv_lp, beam_width_=W)
    G["csv_beam2d_same_w10"] = decoding_cpp.cpp_beam_search_2d(csv_lp, csv_lp, beam_width_=10)
    env10 = np.array([(max(0, i - 10), min(i + 10, T)) for i in range(T)])
    G["csv_env10"] = env10.astype(np.int32)
    G["csv_beam2d_env10_w10"] = decoding_cpp.cpp_beam_search_2d(csv_lp, csv_lp, env10.tolist(), beam_width_=10, method_="row")
    G["csv_beam2d_env10_w10_rowcol"] = decoding_cpp.cpp_beam_search_2d(csv_lp, csv_lp, env10.tolist(), beam_width_=10, method_="row_col")
    envfull = np.tile([0, T - 1], (T, 1))
    G["csv_beam2d_full_w25"] = decoding_cpp.cpp_beam_search_2d(csv_lp, csv_lp)
    G["csv_beam2d_fullenv_w25"] = decoding_cpp.cpp_beam_search_2d(csv_lp, csv_lp, envfull.tolist())
    envdiag = np.array([(i, i + 1) for i in range(T)])
    G["csv_beam2d_diag_w25"] = decoding_cpp.cpp_beam_search_2d(csv_lp, csv_lp, envdiag.tolist())
    G["csv_forward_w25"] = decoding_cpp.cpp_forward(csv_lp, G["csv_beam1d_w25"])

    # ---------------- known-answer toys of the reference tests (brute-force truth from tests/testing.py)
    sys.path.insert(0, os.path.join(ref, "tests"))
    from testing import poreover_profile, joint_profile  # noqa

    toys = [
        np.array([[0.8, 0.1, 0.1], [0.1, 0.3, 0.6], [0.7, 0.2, 0.1], [0.1, 0.1, 0.8]]),  # test_beam.py:13
        np.array([[0.4, 0.5, 0.1], [0.4, 0.2, 0.4], [0.3, 0.5, 0.2]]),  # test_beam.py:19
        np.array([[0.7, 0.2, 0.1], [0.2, 0.3, 0.5], [0.7, 0.2, 0.1], [0.05, 0.05, 0.9]]),  # test_beam.py:44
    ]
    for i, y in enumerate(toys):
        G["toy%d_y" % i] = y
        prof = poreover_profile(y, ("A", "B", ""))
        G["toy%d_top_label" % i] = prof.top_label()[0]
        G["toy%d_beam1d" % i] = decoding_cpp.cpp_beam_search(np.log(y), alphabet_="AB")
        G["toy%d_viterbi" % i] = prof.viterbi_decode()
        labs = ["A", "B", "AB", "BA", "AA", "BB", "ABA", "BAB", "AAB"]
        G["toy%d_forward_labels" % i] = np.array(labs)
        G["toy%d_forward" % i] = np.array([decoding_cpp.cpp_forward(np.log(y), l, alphabet_="AB") for l in labs])
        G["toy%d_label_prob" % i] = np.array([prof.label_prob(l) for l in labs])
    jp = joint_profile(poreover_profile(toys[0], ("A", "B", "")), poreover_profile(toys[2], ("A", "B", "")))
    G["toy_joint_top"] = jp.top_label()[0]
    G["toy_joint_beam2d"] = decoding_cpp.cpp_beam_search_2d(np.log(toys[0]), np.log(toys[2]), alphabet_="AB")

    # ---------------- synthetic bonito reads: viterbi / mapping / 1D beam / forward
    tmp = os.path.join(scratch, "npy")
    os.makedirs(tmp, exist_ok=True)
    for k, T in ((0, 300), (1, 700)):
        p, _ = synth.make_read(500 + k, T)
        f = os.path.join(tmp, "read%d.npy" % k)
        np.save(f, p)
        for bc, mt in (("bonito", "ctc_merge_repeats"), ):
            m = decode.model_from_trace(f, bc)
            G["syn%d_prob" % k] = p
            G["syn%d_log_prob" % k] = m.log_prob.astype(np.float32)
            assert np.array_equal(G["syn%d_log_prob" % k].astype(np.float64), m.log_prob)
            seq, path = m.viterbi_decode(return_path=True)
            G["syn%d_viterbi_seq" % k] = seq
            G["syn%d_viterbi_path" % k] = np.asarray(path, dtype=np.int64)
            G["syn%d_s2s" % k] = np.asarray(pair_decode.get_sequence_mapping(path, "bonito")[0], dtype=np.int64)
            for W in (5, 25):
                G["syn%d_beam1d_bonito_w%d" % (k, W)] = decoding_cpp.cpp_beam_search(m.log_prob, W, "ACGT", mt)
                G["syn%d_beam1d_ctc_w%d" % (k, W)] = decoding_cpp.cpp_beam_search(m.log_prob, W, "ACGT", "ctc")
            G["syn%d_forward_bonito" % k] = decoding_cpp.cpp_forward(m.log_prob, seq, "ACGT", mt)
            G["syn%d_forward_ctc" % k] = decoding_cpp.cpp_forward(m.log_prob, seq, "ACGT", "ctc")
            m.reverse_complement()
            G["syn%d_rc_viterbi_seq" % k] = m.viterbi_decode()
        # poreover-kind viterbi on the same matrix (argmax, repeats kept)
        mp = transducer.poreover(G["syn%d_log_prob" % k])
        seq, path = mp.viterbi_decode(return_path=True)
        G["syn%d_viterbi_seq_poreover" % k] = seq
        G["syn%d_s2s_poreover" % k] = np.asarray(pair_decode.get_sequence_mapping(path, "poreover")[0], dtype=np.int64)

    # ---------------- flip-flop viterbi on synthetic uint8 traces (decode.py:92-93 transform)
    for k, T in ((0, 200), (1, 900)):
        tr = synth.make_flipflop_trace(900 + k, T)
        eps = 0.0000001
        lp = np.log((tr + eps) / (255 + eps))
        mf = transducer.flipflop(lp)
        seq, path = mf.viterbi_decode(return_path=True)
        G["ff%d_trace" % k] = tr
        G["ff%d_seq" % k] = seq
        G["ff%d_path" % k] = np.asarray(path, dtype=np.int64)
        G["ff%d_s2s" % k] = np.asarray(pair_decode.get_sequence_mapping(path, "flipflop")[0], dtype=np.int64)

    # ---------------- banded alignment + columns + envelope
    cases = []
    for n, rate, band in ((12, 0.2, 500), (40, 0.15, 3), (40, 0.15, 1), (150, 0.1, 10), (700, 0.12, 500),
                          (1300, 0.1, 500), (90, 0.5, 500), (1, 0.0, 500), (5, 0.0, 2)):
        a = "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
        b = mutate(rng, a, rate) or "A"
        cases.append((a, b, band))
    cases.append(("ACGTACGT", "ACGACGT", 500))
    cases.append(("A" * 30, "A" * 22, 500))
    cases.append(("ACGT" * 20, "TTTT", 500))
    G["aln_n"] = len(cases)
    for i, (a, b, band) in enumerate(cases):
        al = align.global_pair_banded(a, b, band)
        G["aln%d_in" % i] = np.array([a, b, str(band)])
        G["aln%d_out" % i] = np.array(["".join(al[0]), "".join(al[1])])
        arr = np.array([list(s) for s in al[:2]])
        cols = envelope.get_alignment_columns(arr)
        G["aln%d_cols" % i] = np.array([("mid".index(c[0]), c[1], c[2]) for c in cols], dtype=np.int32).reshape(-1, 3)
        # a plausible sequence->signal mapping: strictly increasing frame indices
        la = max(max(c[1] for c in cols) + 1, len(a))
        lb = max(max(c[2] for c in cols) + 1, len(b))
        s1 = np.sort(rng.choice(np.arange(1, 3 * la + 2), size=len(a), replace=False))
        s2 = np.sort(rng.choice(np.arange(1, 3 * lb + 2), size=len(b), replace=False))
        U, V = 3 * la + 5, 3 * lb + 4
        for pad in (5, 150):
            env = envelope.build_envelope(np.zeros((U, 5)), np.zeros((V, 5)), cols, list(s1), list(s2), padding=pad)
            G["aln%d_env_pad%d" % (i, pad)] = np.asarray(env, dtype=np.int64)
        G["aln%d_s2s1" % i] = s1.astype(np.int64)
        G["aln%d_s2s2" % i] = s2.astype(np.int64)
        G["aln%d_UV" % i] = np.array([U, V])
    al = align.global_pair("ACGTTGCAAC", "ACTTGGCAC")
    G["alnfull_out"] = np.array(["".join(al[0]), "".join(al[1])])
    G["alnfull_dp"] = np.asarray(al[2], dtype=np.int32)

    # ---------------- full pair path through pair_decode_helper (pair_decode.py:305-531)
    def ns(f1, f2, bc, W, method="row_col"):
        return Namespace(**{"in": [f1, f2], "dir": tmp, "basecaller": bc, "reverse_complement": True, "out": "out",
                            "threads": 1, "method": "envelope", "single": "viterbi", "logging": "info", "debug": False,
                            "algorithm": "beam", "alignment": "banded", "beam_width": W, "debug_envelope": False,
                            "diagonal_envelope": False, "diagonal_width": 50, "padding": 5, "skip_matches": False,
                            "skip_threshold": 10, "beam_search_method": method, "window": 200})

    pair_cases = [(0, 400, 5), (1, 400, 25), (2, 1000, 25), (3, 1500, 25)]
    G["pair_n"] = len(pair_cases)
    for i, (k, T, W) in enumerate(pair_cases):
        f1, f2 = synth.save_pair(tmp, k, T)
        r = pair_decode.pair_decode_helper(ns(f1, f2, "bonito", W))
        assert len(r) == 3, r
        m1 = decode.model_from_trace(os.path.join(tmp, f1), "bonito")
        m2 = decode.model_from_trace(os.path.join(tmp, f2), "bonito")
        m2.reverse_complement()
        b1, p1 = m1.viterbi_decode(True)
        b2, p2 = m2.viterbi_decode(True)
        al = align.global_pair_banded(b1, b2)
        arr = np.array([list(s) for s in al[:2]])
        cols = envelope.get_alignment_columns(arr)
        env = envelope.build_envelope(m1.log_prob, m2.log_prob, cols, pair_decode.get_sequence_mapping(p1, "bonito")[0],
                                      pair_decode.get_sequence_mapping(p2, "bonito")[0], padding=5)
        cons = decoding_cpp.cpp_beam_search_2d(m1.log_prob, m2.log_prob, env.tolist(), beam_width_=W,
                                               method_="row_col", model_="ctc_merge_repeats")
        fasta2d = r[1].split("\n", 1)[1].replace("\n", "")
        assert fasta2d == cons
        G["pair%d_args" % i] = np.array([k, T, W])
        G["pair%d_lp1" % i] = m1.log_prob.astype(np.float32)
        G["pair%d_lp2_rc" % i] = np.ascontiguousarray(m2.log_prob).astype(np.float32)
        G["pair%d_basecall1" % i] = b1
        G["pair%d_basecall2" % i] = b2
        G["pair%d_align" % i] = np.array(["".join(al[0]), "".join(al[1])])
        G["pair%d_env" % i] = np.asarray(env, dtype=np.int64)
        G["pair%d_consensus" % i] = cons
        G["pair%d_identity" % i] = r[2]["sequence_identity"]
        G["pair%d_fasta1d" % i] = r[0]
        G["pair%d_fasta2d" % i] = r[1]
        # the same pair through the other schedules / tree
        G["pair%d_consensus_row" % i] = decoding_cpp.cpp_beam_search_2d(
            m1.log_prob, m2.log_prob, env.tolist(), beam_width_=W, method_="row", model_="ctc_merge_repeats")
        G["pair%d_consensus_ctc" % i] = decoding_cpp.cpp_beam_search_2d(
            m1.log_prob, m2.log_prob, env.tolist(), beam_width_=W, method_="row_col", model_="ctc")

    out = os.path.join(HERE, "golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB,", len(G), "entries")


if __name__ == "__main__":
    main()
