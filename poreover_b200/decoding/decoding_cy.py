"""Drop-in for the one function of the reference's legacy Cython module poreover/decoding/decoding_cy.pyx that its test
suite reaches outside the legacy prefix search: viterbi_acceptor (decoding_cy.pyx:60-123), the unbanded twin of
decoding_cpp.cpp_viterbi_acceptor.  It runs on the same GPU kernel (poreover_b200/csrc/acceptor.cu) with a band that
covers the whole matrix.  forward_vec_log and pair_gamma_log (decoding_cy.pyx:127-156, :177-220), the helpers of the
legacy prefix search, run on csrc/prefix.cu with the Cython arithmetic (log(exp(a) + exp(b)), -9999 for log 0)."""
import numpy as np

from .. import _lib, batch
from . import decoding_cpp


def forward_vec_log(s, i, y, previous=None):
    """decoding_cy.pyx:127-156"""
    return batch.forward_vec(y, s, i, previous, _lib.PREFIX_CY)


def pair_gamma_log(y1, y2):
    """decoding_cy.pyx:177-220"""
    return batch.pair_gamma_batch([y1], [y2], _lib.PREFIX_CY)[0]


def viterbi_acceptor(y, label_, alphabet='ACGT', band_size=0):
    y = np.asarray(y)
    band = int(band_size) if band_size else max(y.shape[0], len(label_)) + 1  # 0 = no band (decoding_cy.pyx:60)
    return decoding_cpp.cpp_viterbi_acceptor(y, label_, band, alphabet)
