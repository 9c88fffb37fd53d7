"""GPU parity of the prefix beam searches against the oracle / golden fixtures.

Bar (BASELINE.json north_star): decoded strings identical; ranking score of the returned node within
1e-4 absolute of the FP64 reference."""
import numpy as np
import pytest

from poreover_b200 import _lib, batch, synth
from poreover_b200.decoding import decoding_cpp

pytestmark = pytest.mark.gpu
TOL = 1e-4


def s(x):
    return str(x)


def _pair(oracle, k, T):
    p1, p2, _ = synth.make_pair(k, T)
    lp1 = synth.bonito_log_prob(p1)
    lp2 = synth.bonito_log_prob(p2)
    return lp1, lp2, np.ascontiguousarray(oracle.reverse_complement(lp2, "bonito"))


def test_beam_1d_golden(golden):
    for k in (0, 1):
        lp = golden["syn%d_log_prob" % k]
        for W in (5, 25):
            assert decoding_cpp.cpp_beam_search(lp, W, "ACGT", "ctc_merge_repeats") == s(golden["syn%d_beam1d_bonito_w%d" % (k, W)])
            assert decoding_cpp.cpp_beam_search(lp, W, "ACGT", "ctc") == s(golden["syn%d_beam1d_ctc_w%d" % (k, W)])
    y = golden["csv_log_prob"]
    for W in (10, 25):
        assert decoding_cpp.cpp_beam_search(y, beam_width_=W) == s(golden["csv_beam1d_w%d" % W])


def test_beam_1d_batch_scores(oracle):
    arrays = [synth.bonito_log_prob(synth.make_read(300 + i, T)[0]) for i, T in enumerate((1, 2, 5, 40, 333, 900, 1500))]
    for model in ("ctc_merge_repeats", "ctc"):
        for W in (5, 25, 100):
            seqs, sc, st = batch.beam_search_batch(arrays, W, model)
            for a, g, gs in zip(arrays, seqs, sc):
                w, ws = oracle.beam_search(a, W, model, with_score=True)
                assert g == w, (model, W, len(a))
                assert abs(gs - ws) < TOL
            assert not (st & _lib.ST_POOL_OVERFLOW).any()


def test_beam_2d_csv_golden(golden):
    y = golden["csv_log_prob"]
    T = len(y)
    f = decoding_cpp.cpp_beam_search_2d
    assert f(y, y, beam_width_=10) == s(golden["csv_beam2d_same_w10"])
    env10 = golden["csv_env10"]
    assert f(y, y, env10.tolist(), beam_width_=10, method_="row") == s(golden["csv_beam2d_env10_w10"])
    assert f(y, y, env10.tolist(), beam_width_=10, method_="row_col") == s(golden["csv_beam2d_env10_w10_rowcol"])
    assert f(y, y) == s(golden["csv_beam2d_full_w25"])
    envfull = np.tile([0, T - 1], (T, 1))
    assert f(y, y, envfull.tolist()) == s(golden["csv_beam2d_fullenv_w25"])
    envdiag = np.array([(i, i + 1) for i in range(T)])
    assert f(y, y, envdiag.tolist()) == s(golden["csv_beam2d_diag_w25"])


def test_pair_golden(golden, oracle):
    for i in range(int(golden["pair_n"])):
        k, T, W = [int(x) for x in golden["pair%d_args" % i]]
        lp1, lp2 = golden["pair%d_lp1" % i], golden["pair%d_lp2_rc" % i]
        env = golden["pair%d_env" % i]
        f = decoding_cpp.cpp_beam_search_2d
        assert f(lp1, lp2, env.tolist(), W, "ACGT", "ctc_merge_repeats", "row_col") == s(golden["pair%d_consensus" % i])
        assert f(lp1, lp2, env.tolist(), W, "ACGT", "ctc_merge_repeats", "row") == s(golden["pair%d_consensus_row" % i])
        assert f(lp1, lp2, env.tolist(), W, "ACGT", "ctc", "row_col") == s(golden["pair%d_consensus_ctc" % i])


@pytest.mark.parametrize("method", ["row_col", "row"])
def test_pair_batch_vs_oracle(oracle, method):
    l1, l2, envs, want = [], [], [], {}
    for k, T in ((10, 200), (11, 350), (12, 600), (13, 601), (14, 900), (15, 1200)):
        lp1, _, lp2rc = _pair(oracle, k, T)
        r = oracle.pair_decode(lp1, lp2rc, "bonito", 25, method=method)
        l1.append(lp1); l2.append(lp2rc); envs.append(r["envelope"])
    for model in ("ctc_merge_repeats", "ctc"):
        for W in (5, 25):
            seqs, sc, st = batch.beam_search_2d_batch(l1, l2, envs, W, model, method)
            for a, b, e, g, gs in zip(l1, l2, envs, seqs, sc):
                w, ws = oracle.beam_search_2d(a, b, e, W, model, method, with_score=True)
                assert g == w, (model, W, len(a))
                assert abs(gs - ws) < TOL, (gs, ws)
            assert not (st & _lib.ST_POOL_OVERFLOW).any()


def test_fused_pair_decode(golden, oracle):
    """pob_pair_decode: viterbi x2 -> mapping -> banded NW -> envelope -> row_col search, all on the device."""
    l1, l2 = [], []
    for i in range(int(golden["pair_n"])):
        k, T, W = [int(x) for x in golden["pair%d_args" % i]]
        p1, p2, _ = synth.make_pair(k, T)
        l1.append(synth.bonito_log_prob(p1)); l2.append(synth.bonito_log_prob(p2))
    res = batch.pair_decode_batch(l1, l2, "bonito", beam_width=25, rc2=True)
    for i, r in enumerate(res):
        W = int(golden["pair%d_args" % i][2])
        assert r["basecall1"] == s(golden["pair%d_basecall1" % i])
        assert r["basecall2"] == s(golden["pair%d_basecall2" % i])
        assert r["identity"] == float(golden["pair%d_identity" % i])
        if W == 25:
            assert r["consensus"] == s(golden["pair%d_consensus" % i])
        want = oracle.pair_decode(l1[i], oracle.reverse_complement(l2[i], "bonito"), "bonito", 25, with_score=True)
        assert r["consensus"] == want["consensus"]
        assert abs(r["score"] - want["score"]) < TOL
    # unrelated reads: identity < 0.5 -> skipped exactly like the reference
    a = synth.bonito_log_prob(synth.make_read(1, 500)[0])
    b = synth.bonito_log_prob(synth.make_read(2, 520)[0])
    r = batch.pair_decode_batch([a], [b], "bonito", 25, rc2=True)[0]
    want = oracle.pair_decode(a, oracle.reverse_complement(b, "bonito"), "bonito", 25)
    assert r["skipped"] == want["skipped"] == 1 and (r["status"] & _lib.ST_SKIPPED_IDENTITY)
    assert r["identity"] == want["identity"]
