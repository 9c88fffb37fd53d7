"""Edge cases through the C ABI: empty batches, zero-length and one-timestep reads inside ragged batches, pairs the
reference skips (length mismatch, low identity) next to pairs it decodes, and the `decode` command line."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from poreover_b200 import _lib, batch, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(seed, T):
    return synth.bonito_log_prob(synth.make_read(seed, T)[0])


def test_empty_batches():
    assert batch.viterbi_batch([], "bonito")[0] == []
    assert batch.beam_search_batch([], 25, "ctc")[0] == []
    assert batch.flipflop_viterbi_batch([])[0] == []
    assert batch.pair_decode_batch([], [], "bonito") == []


def test_zero_length_and_single_timestep_reads():
    empty = np.zeros((0, 5), dtype=np.float32)
    one = np.log(np.array([[0.1, 0.6, 0.1, 0.1, 0.1]], dtype=np.float32))
    blank = np.log(np.array([[0.1, 0.1, 0.1, 0.1, 0.6]], dtype=np.float32))
    reads = [_read(1, 200), empty, one, blank, _read(2, 77)]
    seqs, maps, _, st = batch.viterbi_batch(reads, "bonito")
    assert seqs[1] == "" and st[1] & _lib.ST_EMPTY
    assert seqs[2] == "C" and list(maps[2]) == [0] and seqs[3] == ""
    for i in (0, 4):
        assert seqs[i] == O.viterbi(reads[i], "bonito")[0] and not st[i]
    bs, sc, bst = batch.beam_search_batch(reads, 25, "ctc_merge_repeats")
    assert bs[1] == "" and bst[1] & _lib.ST_EMPTY
    for i in (0, 2, 3, 4):
        w, ws = O.beam_search(reads[i], 25, "ctc_merge_repeats", with_score=True)
        assert bs[i] == w and abs(sc[i] - ws) < 1e-4, i
    tr = [synth.make_flipflop_trace(3, 50), np.zeros((0, 8), dtype=np.uint8), synth.make_flipflop_trace(4, 1)]
    ff = batch.flipflop_viterbi_batch(tr)[0]
    assert ff[1] == ""
    for i in (0, 2):
        assert ff[i] == O.viterbi(synth.flipflop_log_prob(tr[i]), "flipflop")[0]


def test_skipped_pairs_inside_a_batch():
    """pair_decode.py:372-375 (|len1-len2| > 1000) and :395-398 (identity < 0.5) next to ordinary pairs."""
    p1, p2, _ = synth.make_pair(40, 1200)
    q1, q2, _ = synth.make_pair(41, 900)
    long1 = synth.make_read(42, 6000)[0]          # ~2400 bases against ~480: length skip
    short2 = synth.make_read(43, 1200)[0]
    unrel1, unrel2 = synth.make_read(44, 1000)[0], synth.make_read(45, 1000)[0]  # unrelated sequences: identity skip
    a1 = [synth.bonito_log_prob(x) for x in (p1, long1, unrel1, q1)]
    a2 = [synth.bonito_log_prob(x) for x in (p2, short2, unrel2, q2)]
    res = batch.pair_decode_batch(a1, a2, "bonito", beam_width=25, rc2=True)
    for k, r in enumerate(res):
        want = O.pair_decode(a1[k], O.reverse_complement(a2[k], "bonito"), "bonito", 25, with_score=True)
        assert r["basecall1"] == want["basecall1"] and r["basecall2"] == want["basecall2"]
        assert bool(r["skipped"]) == bool(want["skipped"]), k
        if want["skipped"]:
            assert "consensus" not in r
        else:
            assert r["consensus"] == want["consensus"] and abs(r["score"] - want["score"]) < 1e-4
    assert res[1]["status"] & _lib.ST_SKIPPED_LENGTH
    assert res[2]["status"] & _lib.ST_SKIPPED_IDENTITY and res[2]["identity"] < 0.5
    assert not res[0]["skipped"] and not res[3]["skipped"]


def test_decode_command_line(tmp_path):
    """`python -m poreover_b200 decode DIR --basecaller bonito [--algorithm beam]` (decode.py:114-192): one FASTA record
    per file, sequences equal to the oracle's."""
    d = tmp_path / "reads"
    d.mkdir()
    want = {}
    for k in range(5):
        p = synth.make_read(60 + k, 300 + 50 * k)[0]
        np.save(d / ("r%d.npy" % k), p)
        lp = synth.bonito_log_prob(p)
        want["r%d" % k] = (O.viterbi(lp, "bonito")[0], O.beam_search(lp, 5, "ctc_merge_repeats"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    for algo, idx in (("viterbi", 0), ("beam", 1)):
        out = tmp_path / ("out_" + algo)
        cmd = [sys.executable, "-m", "poreover_b200", "decode", str(d), "--basecaller", "bonito", "--out", str(out),
               "--algorithm", algo, "--beam_width", "5"]
        subprocess.run(cmd, check=True, env=env, cwd=ROOT, capture_output=True)
        text = open(str(out) + ".fasta").read()
        recs = {}
        for block in text.split(">")[1:]:
            name, seq = block.split("\n", 1)
            recs[os.path.splitext(os.path.basename(name.strip()))[0]] = seq.replace("\n", "")
        assert set(recs) == set(want)
        for name, seq in recs.items():
            assert seq == want[name][idx], (algo, name)
