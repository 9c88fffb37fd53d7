"""CPU: pin the oracle (oracle/poreover_oracle.c) to the golden vectors recorded from the real
reference, and to the reference's own compiled C++ (oracle/_ref) where that exists."""
import numpy as np
import pytest

from poreover_b200 import synth

AB = str.maketrans("AC", "AB")  # toys use alphabet "AB"; the oracle always spells ACGT


def s(x):
    return str(x)


def test_viterbi_csv(golden, oracle):
    seq, path = oracle.viterbi(golden["csv_log_prob"], "poreover")
    assert seq == s(golden["csv_viterbi_seq"])
    assert np.array_equal(path, golden["csv_viterbi_path"])
    assert np.array_equal(oracle.sequence_mapping(path, "poreover"), golden["csv_s2s"])


@pytest.mark.parametrize("k", [0, 1])
def test_viterbi_bonito(golden, oracle, k):
    lp = golden["syn%d_log_prob" % k]
    assert np.array_equal(synth.bonito_log_prob(golden["syn%d_prob" % k]), lp)  # loader transform
    seq, path = oracle.viterbi(lp, "bonito")
    assert seq == s(golden["syn%d_viterbi_seq" % k])
    assert np.array_equal(path, golden["syn%d_viterbi_path" % k])
    assert np.array_equal(oracle.sequence_mapping(path, "bonito"), golden["syn%d_s2s" % k])
    seq, path = oracle.viterbi(lp, "poreover")
    assert seq == s(golden["syn%d_viterbi_seq_poreover" % k])
    assert np.array_equal(oracle.sequence_mapping(path, "poreover"), golden["syn%d_s2s_poreover" % k])
    rc = oracle.reverse_complement(lp, "bonito")
    assert oracle.viterbi(rc, "bonito")[0] == s(golden["syn%d_rc_viterbi_seq" % k])


@pytest.mark.parametrize("k", [0, 1])
def test_viterbi_flipflop(golden, oracle, k):
    lp = synth.flipflop_log_prob(golden["ff%d_trace" % k])
    seq, path = oracle.viterbi(lp, "flipflop")
    assert seq == s(golden["ff%d_seq" % k])
    assert np.array_equal(path, golden["ff%d_path" % k])
    assert np.array_equal(oracle.sequence_mapping(path, "flipflop"), golden["ff%d_s2s" % k])


def test_toys(golden, oracle):
    for i in range(3):
        y = np.log(golden["toy%d_y" % i])
        got = oracle.beam_search(y, 25, "ctc").translate(AB)
        assert got == s(golden["toy%d_beam1d" % i]) == s(golden["toy%d_top_label" % i])
        assert oracle.viterbi(y, "poreover")[0].translate(AB) == s(golden["toy%d_viterbi" % i])
        for lab, want, p in zip(golden["toy%d_forward_labels" % i], golden["toy%d_forward" % i],
                                golden["toy%d_label_prob" % i]):
            f = oracle.forward(y, s(lab).translate(str.maketrans("AB", "AC")), "ctc")
            assert f == want or abs(f - want) < 1e-12
            if p > 0:
                assert abs(np.exp(f) - p) < 1e-12  # brute-force path enumeration (tests/testing.py)
    got = oracle.beam_search_2d(np.log(golden["toy0_y"]), np.log(golden["toy2_y"]), None, 25, "ctc", "row")
    assert got.translate(AB) == s(golden["toy_joint_beam2d"]) == s(golden["toy_joint_top"])


def test_csv_beams(golden, oracle):
    y = golden["csv_log_prob"]
    T = len(y)
    for W in (10, 25):
        assert oracle.beam_search(y, W, "ctc") == s(golden["csv_beam1d_w%d" % W])
    assert oracle.beam_search_2d(y, y, None, 10, "ctc", "row") == s(golden["csv_beam2d_same_w10"])
    env10 = golden["csv_env10"]
    assert oracle.beam_search_2d(y, y, env10, 10, "ctc", "row") == s(golden["csv_beam2d_env10_w10"])
    assert oracle.beam_search_2d(y, y, env10, 10, "ctc", "row_col") == s(golden["csv_beam2d_env10_w10_rowcol"])
    assert oracle.beam_search_2d(y, y, None, 25, "ctc", "row") == s(golden["csv_beam2d_full_w25"])
    envfull = np.tile([0, T - 1], (T, 1))
    assert oracle.beam_search_2d(y, y, envfull, 25, "ctc", "row") == s(golden["csv_beam2d_fullenv_w25"])
    envdiag = np.array([(i, i + 1) for i in range(T)])
    assert oracle.beam_search_2d(y, y, envdiag, 25, "ctc", "row") == s(golden["csv_beam2d_diag_w25"])
    # reference test_beam.py: 2D(y,y) == 1D(y) in these configurations
    assert s(golden["csv_beam2d_diag_w25"]) == s(golden["csv_beam1d_w25"])
    f = oracle.forward(y, s(golden["csv_beam1d_w25"]), "ctc")
    assert abs(f - float(golden["csv_forward_w25"])) < 1e-9


@pytest.mark.parametrize("k", [0, 1])
def test_syn_beam1d(golden, oracle, k):
    lp = golden["syn%d_log_prob" % k]
    for W in (5, 25):
        assert oracle.beam_search(lp, W, "ctc_merge_repeats") == s(golden["syn%d_beam1d_bonito_w%d" % (k, W)])
        assert oracle.beam_search(lp, W, "ctc") == s(golden["syn%d_beam1d_ctc_w%d" % (k, W)])
    seq = s(golden["syn%d_viterbi_seq" % k])
    assert abs(oracle.forward(lp, seq, "ctc_merge_repeats") - float(golden["syn%d_forward_bonito" % k])) < 1e-9
    assert abs(oracle.forward(lp, seq, "ctc") - float(golden["syn%d_forward_ctc" % k])) < 1e-9


def test_alignment_and_envelope(golden, oracle):
    for i in range(int(golden["aln_n"])):
        a, b, band = [s(x) for x in golden["aln%d_in" % i]]
        al = oracle.global_pair_banded(a, b, int(band))
        want = golden["aln%d_out" % i]
        assert "".join(al[0]) == s(want[0]) and "".join(al[1]) == s(want[1]), i
        cols = oracle.alignment_columns(al)
        wc = golden["aln%d_cols" % i]
        assert [("mid".index(c[0]), c[1], c[2]) for c in cols] == [tuple(r) for r in wc.tolist()]
        U, V = golden["aln%d_UV" % i]
        for pad in (5, 150):
            env = oracle.build_envelope(int(U), int(V), cols, golden["aln%d_s2s1" % i], golden["aln%d_s2s2" % i], pad)
            assert np.array_equal(env, golden["aln%d_env_pad%d" % (i, pad)]), (i, pad)
    al = oracle.global_pair("ACGTTGCAAC", "ACTTGGCAC")
    assert "".join(al[0]) == s(golden["alnfull_out"][0]) and "".join(al[1]) == s(golden["alnfull_out"][1])
    assert np.array_equal(al[2], golden["alnfull_dp"])
    with pytest.raises(ZeroDivisionError):
        oracle.global_pair_banded("", "ACGT")


def test_pair_path(golden, oracle):
    for i in range(int(golden["pair_n"])):
        k, T, W = [int(x) for x in golden["pair%d_args" % i]]
        p1, p2, _ = synth.make_pair(k, T)
        lp1 = synth.bonito_log_prob(p1)
        lp2 = oracle.reverse_complement(synth.bonito_log_prob(p2), "bonito")
        assert np.array_equal(lp1, golden["pair%d_lp1" % i])  # the generator is deterministic
        assert np.array_equal(lp2, golden["pair%d_lp2_rc" % i])
        r = oracle.pair_decode(lp1, lp2, "bonito", W, padding=5, method="row_col")
        assert r["basecall1"] == s(golden["pair%d_basecall1" % i])
        assert r["basecall2"] == s(golden["pair%d_basecall2" % i])
        assert np.array_equal(r["envelope"], golden["pair%d_env" % i])
        assert r["identity"] == float(golden["pair%d_identity" % i])
        assert r["consensus"] == s(golden["pair%d_consensus" % i])
        env = golden["pair%d_env" % i]
        assert oracle.beam_search_2d(lp1, lp2, env, W, "ctc_merge_repeats", "row") == s(golden["pair%d_consensus_row" % i])
        assert oracle.beam_search_2d(lp1, lp2, env, W, "ctc", "row_col") == s(golden["pair%d_consensus_ctc" % i])


def test_port_vs_compiled_reference(oracle):
    """The restatement against the UNMODIFIED reference C++ (oracle/_ref), strings and ranking scores."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    for k in range(4):
        p1, p2, _ = synth.make_pair(40 + k, 350)
        lp1 = synth.bonito_log_prob(p1).astype(np.float64)
        lp2 = oracle.reverse_complement(synth.bonito_log_prob(p2), "bonito").astype(np.float64)
        for W in (5, 25):
            for kind, model in (("bonito", "ctc_merge_repeats"), ("poreover", "ctc")):
                a = oracle.pair_decode(lp1, lp2, kind, W, backend="port", with_score=True)
                b = oracle.pair_decode(lp1, lp2, kind, W, backend="ref", with_score=True)
                assert a["consensus"] == b["consensus"]
                assert abs(a["score"] - b["score"]) < 1e-9
                env = a["envelope"]
                for backend in ("ref", "stock"):
                    assert oracle.beam_search_2d(lp1, lp2, env, W, model, "row", backend="port") == \
                        oracle.beam_search_2d(lp1, lp2, env, W, model, "row", backend=backend)
                x = oracle.beam_search(lp1, W, model, "port", True)
                y = oracle.beam_search(lp1, W, model, "ref", True)
                assert x[0] == y[0] and abs(x[1] - y[1]) < 1e-9
        al = oracle.ref_align_module().global_pair_banded(a["basecall1"], a["basecall2"], 7)
        mine = oracle.global_pair_banded(a["basecall1"], a["basecall2"], 7)
        assert al[0] == mine[0] and al[1] == mine[1]
    # small no-envelope "row" search (O(U*V)), both trees
    p1, p2, _ = synth.make_pair(77, 60)
    lp1 = synth.bonito_log_prob(p1).astype(np.float64)
    lp2 = oracle.reverse_complement(synth.bonito_log_prob(p2), "bonito").astype(np.float64)
    for model in ("ctc", "ctc_merge_repeats"):
        assert oracle.beam_search_2d(lp1, lp2, None, 10, model, "row", backend="port") == \
            oracle.beam_search_2d(lp1, lp2, None, 10, model, "row", backend="ref")


def test_oracle_acceptor_matches_reference_function(oracle):
    """orc_viterbi_acceptor against the unmodified viterbi_acceptor_poreover (Forward.h:14-121) in oracle/_ref."""
    O = oracle
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    for seed, T, band in [(1, 300, 1000), (2, 600, 40), (3, 1500, 100), (4, 1000, 5), (5, 97, 7)]:
        lp = synth.bonito_log_prob(synth.make_read(seed, T)[0])
        lab = O.beam_search(lp, 25, "ctc")
        a = O.viterbi_acceptor(lp, lab, band, "port")
        b = O.viterbi_acceptor(lp, lab, band, "ref")
        assert np.array_equal(a, b), (seed, T, band)
        assert "".join("ACGT"[i] for i in a[a != 4]) == lab
