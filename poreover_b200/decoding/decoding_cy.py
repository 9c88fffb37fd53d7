"""Drop-in for the one function of the reference's legacy Cython module poreover/decoding/decoding_cy.pyx that its test
suite reaches outside the legacy prefix search: viterbi_acceptor (decoding_cy.pyx:60-123), the unbanded twin of
decoding_cpp.cpp_viterbi_acceptor.  It runs on the same GPU kernel (poreover_b200/csrc/acceptor.cu) with a band that
covers the whole matrix.  The rest of decoding_cy serves `--algorithm prefix` only and is out of scope."""
import numpy as np

from . import decoding_cpp


def viterbi_acceptor(y, label_, alphabet='ACGT', band_size=0):
    y = np.asarray(y)
    band = int(band_size) if band_size else max(y.shape[0], len(label_)) + 1  # 0 = no band (decoding_cy.pyx:60)
    return decoding_cpp.cpp_viterbi_acceptor(y, label_, band, alphabet)
