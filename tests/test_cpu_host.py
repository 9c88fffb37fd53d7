"""CPU: host-side logic, the C-ABI library's exported surface, and the multi-process work queue."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from poreover_b200 import _lib, batch, multigpu, synth
from poreover_b200.decoding import decode as gdecode
from poreover_b200.decoding import envelope as genv
from poreover_b200.decoding import pair_decode as gpd

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "poreover_b200.h")).read()
    declared = set(re.findall(r"\b(pob_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pob_ctx", "pob_reads_t"}
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libporeover_b200.so lacks %s" % name
        assert name in _lib.SIGNATURES, "python binding lacks %s" % name
    l = _lib.lib()
    assert l.pob_abi_version() == 1
    assert l.pob_strerror(-1) == b"invalid argument"
    assert l.pob_kernel_name(5) == b"beam_pair"


def test_no_cpu_fallback():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    lp = np.log(np.full((10, 5), 0.2, dtype=np.float32))
    with pytest.raises(_lib.PoreoverB200Error):
        batch.viterbi_batch([lp], "bonito")
    with pytest.raises(_lib.PoreoverB200Error):
        batch.pair_decode_batch([lp], [lp])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poreover_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), "%s mentions the oracle" % f


def test_sequence_mapping_host(oracle):
    rng = np.random.default_rng(0)
    for kind, hi in (("poreover", 5), ("bonito", 5), ("flipflop", 8)):
        for T in (1, 2, 7, 100, 1000):
            for _ in range(5):
                path = rng.integers(0, hi, size=T)
                if rng.random() < 0.3:
                    path[-1] = path[0]
                got, sig = gpd.get_sequence_mapping(path, kind)
                assert np.array_equal(got, oracle.sequence_mapping(path, kind)), (kind, T)


def test_alignment_columns_and_fasta(golden, oracle):
    for i in range(int(golden["aln_n"])):
        rows = [list(str(x)) for x in golden["aln%d_out" % i]]
        cols = genv.get_alignment_columns(np.array(rows))
        assert [("mid".index(c[0]), c[1], c[2]) for c in cols] == [tuple(r) for r in golden["aln%d_cols" % i].tolist()]
    assert gdecode.fasta_format("r", "A" * 130) == ">r\n" + "A" * 60 + "\n" + "A" * 60 + "\n" + "A" * 10 + "\n"
    assert gdecode.fasta_format("r", "A" * 60) == ">r\n" + "A" * 60 + "\n"  # reference quirk: no empty last line
    assert str(golden["pair0_fasta2d"]).startswith(">consensus;pair00000_1;pair00000_2\n")


def test_loaders(tmp_path, golden):
    p = golden["syn0_prob"]
    f = tmp_path / "r.npy"
    np.save(f, p)
    m = gdecode.model_from_trace(str(f), "bonito")
    assert m.kind == "bonito" and np.array_equal(m.log_prob, golden["syn0_log_prob"].astype(np.float64))
    assert m.device_array().dtype == np.float32
    logits = np.random.default_rng(1).normal(size=(3, 40, 5)).astype(np.float32)
    g = tmp_path / "l.npy"
    np.save(g, logits)
    m = gdecode.model_from_trace(str(g), "poreover")
    assert m.log_prob.shape == (120, 5)
    assert np.allclose(np.exp(m.log_prob).sum(axis=1), 1, atol=1e-5)
    m.reverse_complement()
    assert m.log_prob.shape == (120, 5) and m.device_array().flags["C_CONTIGUOUS"]


def test_read_batch_packing():
    arrays = [np.full((t, 5), float(t), dtype=np.float32) for t in (1, 5, 8, 3)]
    b = batch.ReadBatch(arrays, rc=[0, 1, 0, 1])
    assert np.all(b.row_off % 4 == 0) and list(b.lens) == [1, 5, 8, 3]
    for a, o in zip(arrays, b.row_off[:-1]):
        assert np.array_equal(b.data[o:o + len(a)], a)
    s = b.struct()
    assert s.n == 4 and s.n_states == 5 and s.dtype == _lib.F32


def test_cli_parser_matches_reference_flags():
    from poreover_b200.__main__ import build_parser
    a = build_parser().parse_args(["pair-decode", "pairs.txt", "--basecaller", "bonito", "--reverse_complement"])
    # defaults of poreover/__main__.py:69-91
    assert (a.beam_width, a.padding, a.alignment, a.single, a.beam_search_method, a.method) == \
        (5, 5, "banded", "viterbi", "row_col", "envelope")
    d = build_parser().parse_args(["decode", "x.npy", "--basecaller", "bonito"])
    assert (d.algorithm, d.beam_width, d.window, d.out) == ("viterbi", 25, 400, "out")


def test_work_queue_single_process():
    items = list(range(50))
    out = multigpu.run_sharded(items, [i % 7 for i in items], lambda c: [x + 1 for x in c], chunk=8)
    assert out == [x + 1 for x in items]


def test_work_queue_two_ranks_gloo(tmp_path):
    res = tmp_path / "res.json"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "_queue_worker.py"), str(res)]
    subprocess.run(cmd, check=True, env=env, timeout=300, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = json.load(open(res))
    assert [tuple(x[:2]) for x in r["out"]] == [(i, i * i) for i in range(103)]
    assert sum(r["counts"]) == 103 and all(c > 0 for c in r["counts"])
    assert {x[2] for x in r["out"]} == {0, 1}  # both ranks pulled work from the shared queue


def test_strong_scaling_step_and_failure_two_ranks_gloo(tmp_path):
    """bench.py's strong-scaling step on the host side (two lanes per rank on one queue, records to rank 0 through the
    store), run_sharded when a rank fails (nobody hangs) and a second queue in the same process group."""
    res = tmp_path / "strong.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "_strong_worker.py"), str(res)]
    subprocess.run(cmd, check=True, timeout=300, capture_output=True)
    rep = json.load(open(res))
    assert len(rep["steps"]) == 3
    for s in rep["steps"]:
        assert s["covered"] == 1000 and 0 < s["own_chunks"] <= s["chunks"]  # every pair exactly once
    assert sum(s["own_chunks"] for s in rep["steps"]) < sum(s["chunks"] for s in rep["steps"])  # rank 1 pulled work too
    assert rep["errors"][1].startswith("ValueError: boom")
    assert rep["errors"][0].startswith("RuntimeError: decoding failed on another rank") and "boom" in rep["errors"][0]
    assert rep["second_run"] == [2 * x for x in range(23)]


def test_get_anchors_matches_reference_loop():
    """The vectorised get_anchors against a literal transcription of the reference loop's rules on random alignments."""
    from poreover_b200.decoding.pair_decode import get_anchors
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(1, 400))
        a1 = rng.choice(list("ACGT-"), size=n, p=[0.23, 0.23, 0.23, 0.23, 0.08])
        a2 = np.where(rng.random(n) < 0.8, a1, rng.choice(list("ACGT-"), size=n))
        both = (a1 == '-') & (a2 == '-')
        a2[both] = 'A'
        ranges, types = [], []
        start, count, prev = 0, 1, 'START'
        for i in range(n):
            st = 'mat' if a1[i] == a2[i] else ('ins' if a1[i] == '-' else ('del' if a2[i] == '-' else 'mis'))
            if prev == st and st != 'mis':
                count += 1
            else:
                if (prev in ('ins', 'del') and count >= 3) or (prev == 'mat' and count >= 4):
                    ranges.append((start, i)); types.append(prev)
                prev, count, start = st, 1, i
        got = get_anchors(np.array([a1, a2]), matches=4, indels=3)
        assert got == (ranges, types)
