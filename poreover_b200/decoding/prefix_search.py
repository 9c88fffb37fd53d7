"""Legacy prefix search: the reference's decoding/prefix_search.py with the searches and the gamma matrix on the GPU.

Same names, arguments and return values as the reference module (decoding/prefix_search.py); the work is done by
pob_prefix_search, pob_pair_gamma and pob_pair_prefix_search (csrc/prefix.cu), which reproduce the reference's two
arithmetic flavours: the plain functions use np.logaddexp / scipy logsumexp with LOG_0 = -inf, the *_cy functions the
Cython helpers' log(exp(a) + exp(b)) and -9999 (decoding_cy.pyx:18, :127-156, :177-220).

What stays on the host: spelling indices with the alphabet, greedy_search / remove_gaps (string utilities,
prefix_search.py:16-30) and forward_vec_no_gap_log (a shift and an add of two vectors, :67-79).  Not provided:
return_forward=True (an (alphabet, T, T) table that only the unreachable `pair-decode --algorithm prefix` branch reads,
pair_decode.py:175-186 behind the always-false assert of :224) and the envelope variant of the 2D search
(PairPrefixSearch.cpp: double free, Gamma.h:100).
"""
from collections import OrderedDict

import numpy as np

from .. import _lib, batch
from . import decoding_cy  # noqa: F401  (the reference module exposes it: prefix_search.decoding_cy, tests/test_prefix.py:194)

# Default alphabet (prefix_search.py:14)
DNA_alphabet = OrderedDict([('A', 0), ('C', 1), ('G', 2), ('T', 3)])

LOG_0 = -float('Inf')
LOG_1 = 0.


def remove_gaps(a):
    """prefix_search.py:16-23: drops '-' and nothing else (repeats are kept)."""
    return ''.join(i for i in a if i != '-')


def greedy_search(logits, alphabet=['A', 'C', 'G', 'T', '-']):
    """prefix_search.py:25-29: highest-probability label at each step, gaps removed."""
    return remove_gaps(np.take(alphabet, np.argmax(logits, axis=1)))


def _letters(alphabet, y):
    """index -> letter for an alphabet dict whose indices are 0..len-1 with the blank in the last column of y"""
    y = np.asarray(y)
    if y.ndim != 2:
        raise ValueError("expected a (time, alphabet + blank) table")
    if len(alphabet) < 1 or len(alphabet) > 4:
        raise NotImplementedError("the GPU searches handle alphabets of one to four letters; got %r" % (alphabet,))
    if sorted(alphabet.values()) != list(range(len(alphabet))) or y.shape[1] != len(alphabet) + 1:
        raise ValueError("alphabet indices must be 0..%d and the table must have one more column (the blank)"
                         % (len(alphabet) - 1))
    inv = [None] * len(alphabet)
    for c, i in alphabet.items():
        inv[i] = c
    if list(alphabet.values()) != list(range(len(alphabet))):
        # the reference evaluates the candidates in dict order and indexes the forward vectors by letter index
        # (prefix_search.py:305): only alphabets listed in index order are consistent there
        raise NotImplementedError("alphabet must list its letters in index order")
    return inv


def _spell(indices, inv):
    return ''.join(inv[int(i)] for i in indices)


def _search_1d(y, alphabet, flavour, return_forward):
    if return_forward:
        raise NotImplementedError("return_forward=True serves only the unreachable pair-decode prefix branch "
                                  "(pair_decode.py:175-186, :224) and is not built")
    inv = _letters(alphabet, y)
    labels, score, _ = batch.prefix_search_batch([y], flavour)
    return (_spell(labels[0], inv), float(score[0]))


def prefix_search_log(y, alphabet=DNA_alphabet, return_forward=False):
    """prefix_search.py:116-174 (numpy arithmetic): (top label, its log label probability)."""
    return _search_1d(y, alphabet, _lib.PREFIX_NUMPY, return_forward)


def prefix_search_log_cy(y_, alphabet=DNA_alphabet, return_forward=False):
    """prefix_search.py:176-238 (Cython helpers' arithmetic)."""
    return _search_1d(y_, alphabet, _lib.PREFIX_CY, return_forward)


def prefix_search_windows(log_prob, window, alphabet=DNA_alphabet):
    """decode --algorithm prefix (decode.py:179-188): the table cut into windows of `window` rows, every window
    searched on its own (one batch), the labels concatenated."""
    inv = _letters(alphabet, log_prob)
    T = len(log_prob)
    cuts = []
    i = 0
    while i + window < T:  # decode.py:185-187
        cuts.append((i, i + window))
        i += window
    cuts.append((i, T))
    labels, _, _ = batch.prefix_search_batch([log_prob[a:b] for a, b in cuts], _lib.PREFIX_CY)
    return ''.join(_spell(l, inv) for l in labels)


def pair_gamma_log(y1, y2):
    """prefix_search.py:35-65: dense (U+1) x (V+1) matrix, gamma[0, 0] = log P(both reads spell the same label)."""
    return batch.pair_gamma_batch([y1], [y2], _lib.PREFIX_NUMPY)[0]


def forward_vec_log(s, i, y, previous=None):
    """prefix_search.py:81-97: one column of the 1D forward algorithm on letter s at label position i."""
    return batch.forward_vec(y, s, i, previous, _lib.PREFIX_NUMPY)


def forward(l, y, fw_fn=forward_vec_log):
    """prefix_search.py:99-114: the full forward matrix, (len(l) + 1) x len(y), column by column."""
    prev = fw_fn(-1, 0, y)
    alpha = np.zeros((len(l) + 1, len(y)))
    alpha[0] = prev
    for i, s in enumerate(l):
        prev = fw_fn(s, i + 1, y, prev)
        alpha[i + 1] = prev
    return alpha


def forward_vec_no_gap_log(l, y, fw0):
    """prefix_search.py:67-79: forward variable of paths that do not end on a gap (vector glue, host)."""
    return np.insert(fw0[:-1], 0, LOG_1 if len(l) == 1 else LOG_0) + y[:, l[-1]]


def pair_prefix_search_log(y1, y2, alphabet=DNA_alphabet):
    """prefix_search.py:247-310 (numpy arithmetic): (top label, log label probability given agreement)."""
    inv = _letters(alphabet, y1)
    labels, score, _ = batch.pair_prefix_search_batch([y1], [y2], _lib.PREFIX_NUMPY)
    return (_spell(labels[0], inv), float(score[0]))


def pair_prefix_search_log_cy(y1_, y2_, alphabet=DNA_alphabet):
    """prefix_search.py:312-385 (Cython helpers' arithmetic for gamma and the forward vectors)."""
    inv = _letters(alphabet, y1_)
    labels, score, _ = batch.pair_prefix_search_batch([y1_], [y2_], _lib.PREFIX_CY)
    return (_spell(labels[0], inv), float(score[0]))
