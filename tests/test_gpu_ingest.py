"""GPU: the command-line ingest path (poreover_b200/ingest.py).  Batches that keep bonito's file column order
(POB_BLANK_FIRST) must decode exactly like the permuted arrays of the reference loader, through every kernel of the
pair path, and the chunked two-stage pipeline must return what one big call returns."""
import os
import subprocess
import sys
from argparse import Namespace

import numpy as np
import pytest

from poreover_b200 import _lib, batch, ingest, synth
from poreover_b200.decoding import decode as gdecode
from poreover_b200.decoding import pair_decode as gpd

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _args(d, **over):
    base = {"in": [], "dir": str(d), "basecaller": "bonito", "reverse_complement": True, "out": "out", "threads": 1,
            "method": "envelope", "single": "viterbi", "logging": "info", "debug": False, "algorithm": "beam",
            "alignment": "banded", "beam_width": 25, "debug_envelope": False, "diagonal_envelope": False,
            "diagonal_width": 50, "padding": 5, "skip_matches": False, "skip_threshold": 10,
            "beam_search_method": "row_col", "window": 200}
    base.update(over)
    return Namespace(**base)


def _files(tmp_path, n, T0, blank_last=False):
    pairs = [list(synth.save_pair(str(tmp_path), 300 + k, T0 + 41 * (k % 5), blank_last=blank_last)) for k in range(n)]
    f1 = [os.path.join(str(tmp_path), p[0]) for p in pairs]
    f2 = [os.path.join(str(tmp_path), p[1]) for p in pairs]
    return pairs, f1, f2


@pytest.mark.parametrize("basecaller,W,method", [("bonito", 25, "row_col"), ("bonito", 5, "row"), ("poreover", 10, "row_col")])
def test_file_order_batches_decode_like_loader_arrays(tmp_path, basecaller, W, method):
    pairs, f1, f2 = _files(tmp_path, 7, 500, blank_last=basecaller == "poreover")
    a1 = [gdecode.model_from_trace(p, basecaller).device_array() for p in f1]
    a2 = [gdecode.model_from_trace(p, basecaller).device_array() for p in f2]
    want = batch.pair_decode_batch(a1, a2, kind=basecaller, beam_width=W, method=method, rc2=True)
    b1 = ingest.load_reads(f1, basecaller)
    b2 = ingest.load_reads(f2, basecaller, rc=1)
    assert b1.layout == (_lib.BLANK_FIRST if basecaller == "bonito" else _lib.BLANK_LAST)
    got = batch.pair_decode_batch(b1, b2, kind=basecaller, beam_width=W, method=method)
    assert got == want  # basecalls, consensus, identity, status and the FP64 ranking score, bit for bit
    assert all(not r["skipped"] and len(r["consensus"]) > 100 for r in got)
    # the single-read kernels on the same batches
    for b, arrs in ((b1, a1), (b2, [a[::-1, [3, 2, 1, 0, 4]] for a in a2])):
        s_got, m_got, p_got, _ = batch.viterbi_batch(b, basecaller, return_path=True)
        s_want, m_want, p_want, _ = batch.viterbi_batch(arrs, basecaller, return_path=True)
        assert s_got == s_want and all(np.array_equal(x, y) for x, y in zip(m_got, m_want))
        assert all(np.array_equal(x, y) for x, y in zip(p_got, p_want))
    model = gdecode.MODEL_TYPE[basecaller]
    s_got, sc_got, _ = batch.beam_search_batch(b1, 5, model)
    s_want, sc_want, _ = batch.beam_search_batch(a1, 5, model)
    assert s_got == s_want and np.array_equal(sc_got, sc_want)


def test_chunked_pipeline_equals_one_call(tmp_path):
    pairs, f1, f2 = _files(tmp_path, 9, 400)
    args = _args(tmp_path)
    one = gpd.decode_pairs(args, pairs, chunk=64)
    many = gpd.decode_pairs(args, pairs, chunk=2)  # five chunks, each loaded while the previous one is decoded
    assert one == many and all(len(r) == 3 and r[2]["skipped"] == 0 for r in one)
    staged = gpd.decode_pairs(_args(tmp_path, alignment="full"), pairs, chunk=4)
    assert [r[1] for r in staged] == [r[1] for r in one]  # same consensus records through the staged flags' loader


def test_pair_decode_command_line(tmp_path):
    """`python -m poreover_b200 pair-decode pairs.txt` (pair_decode.py:230-303): the three output files, records in
    input order, sequences equal to the in-process batch call on the reference loader's arrays."""
    pairs, f1, f2 = _files(tmp_path, 6, 450)
    with open(tmp_path / "pairs.txt", "w") as f:
        for p in pairs:
            f.write("%s %s\n" % (p[0], p[1]))
    out = tmp_path / "run"
    cmd = [sys.executable, "-m", "poreover_b200", "pair-decode", str(tmp_path / "pairs.txt"), "--dir", str(tmp_path),
           "--basecaller", "bonito", "--reverse_complement", "--beam_width", "25", "--out", str(out)]
    subprocess.run(cmd, check=True, env=dict(os.environ, PYTHONPATH=ROOT), cwd=ROOT, capture_output=True)
    a1 = [gdecode.model_from_trace(p, "bonito").device_array() for p in f1]
    a2 = [gdecode.model_from_trace(p, "bonito").device_array() for p in f2]
    want = batch.pair_decode_batch(a1, a2, kind="bonito", beam_width=25, rc2=True)

    def records(path):
        return [(b.split("\n", 1)[0], b.split("\n", 1)[1].replace("\n", "")) for b in open(path).read().split(">")[1:]]

    two_d = records(str(out) + ".2d.fasta")
    assert [s for _, s in two_d] == [r["consensus"] for r in want]
    assert [n for n, _ in two_d] == ["consensus;%s;%s" % (p[0][:-4], p[1][:-4]) for p in pairs]
    one_d = records(str(out) + ".1d.fasta")
    assert [s for _, s in one_d] == [x for r in want for x in (r["basecall1"], r["basecall2"])]
    log = [l.split("\t") for l in open(str(out) + ".log") if not l.startswith("#")]
    assert [(l[0], l[1], int(l[2]), int(l[3])) for l in log] == \
        [(p[0], p[1], r["length1"], r["length2"]) for p, r in zip(pairs, want)]
