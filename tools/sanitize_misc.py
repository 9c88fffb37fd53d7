"""Small cases of the non-search kernels for compute-sanitizer: viterbi (cp.async ring), flip-flop (4 reads per warp),
acceptor, banded / full NW, envelope, and one fused pair-decode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from poreover_b200 import batch, synth

bad = 0
reads = [synth.bonito_log_prob(synth.make_read(s, T)[0]) for s, T in ((1, 130), (2, 517), (3, 64), (4, 3), (5, 1000))]
seqs, maps, _, st = batch.viterbi_batch(reads, "bonito", rc=[0, 1, 0, 1, 1])
for lp, rc, s in zip(reads, [0, 1, 0, 1, 1], seqs):
    want = O.viterbi(O.reverse_complement(lp, "bonito") if rc else lp, "bonito")[0]
    bad += s != want
tr = [synth.make_flipflop_trace(s, T) for s, T in ((1, 100), (2, 257), (3, 9), (4, 640), (5, 33), (6, 1))]
ff = batch.flipflop_viterbi_batch(tr)
for t, s in zip(tr, ff[0]):
    bad += s != O.viterbi(synth.flipflop_log_prob(t), "flipflop")[0]
lab = [O.beam_search(r, 25, "ctc") for r in reads[:3]]
paths, _ = batch.viterbi_acceptor_batch(reads[:3], lab, 20)
for lp, l, p in zip(reads[:3], lab, paths):
    bad += not np.array_equal(p, O.viterbi_acceptor(lp, l, 20))
p1, p2, _ = synth.make_pair(7, 300)
r = batch.pair_decode_batch([synth.bonito_log_prob(p1)], [synth.bonito_log_prob(p2)], "bonito", 25, rc2=True)[0]
w = O.pair_decode(synth.bonito_log_prob(p1), O.reverse_complement(synth.bonito_log_prob(p2), "bonito"), "bonito", 25)
bad += r["consensus"] != w["consensus"]
g = batch.align_global_batch([r["basecall1"]], [r["basecall2"]])[0]
bad += (g[0], g[1]) != tuple("".join(x) for x in O.global_pair(r["basecall1"], r["basecall2"])[:2])
# banded NW, row-owner fill kernel at several slot counts (band 1/3/10 -> 64 slots, 40 -> 128, 500 -> 1024), uneven
# lengths (several rows of a thread in one pair) and a band wider than the kernel's 1024 threads (generic kernel)
rng = np.random.default_rng(11)
for n1, n2, band in ((90, 100, 1), (300, 120, 3), (57, 211, 10), (400, 380, 40), (700, 650, 500), (150, 140, 600)):
    a = "".join(rng.choice(list("ACGT"), size=n1))
    b = "".join(rng.choice(list("ACGT"), size=n2))
    got = batch.align_banded_batch([a, b], [b, a], band_width=band)
    for (x, y), gt in zip(((a, b), (b, a)), got):
        want = O.global_pair_banded(x, y, band)
        bad += (gt[0], gt[1]) != ("".join(want[0]), "".join(want[1]))
print("done, mismatches:", bad)
