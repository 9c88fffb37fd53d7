"""Drop-in for the reference's Cython module poreover/decoding/decoding_cpp.pyx: same function names,
arguments, defaults and return values; the searches run on the GPU (poreover_b200/csrc/beam.cu)."""
import numpy as np

from .. import batch


def _prep(y_):
    y = np.asarray(y_)
    if y.dtype != np.float32:
        y = np.asarray(y, dtype=np.float64)  # decoding_cpp.pyx:96: coerced to float64
    return np.ascontiguousarray(y)


def _check_alphabet(alphabet_, y):
    """The reference takes any alphabet (decoding_cpp.pyx:88, :107: the gap state is column len(alphabet_)); the GPU
    searches handle one to four letters (two to five states)."""
    if y.ndim != 2 or y.shape[1] != len(alphabet_) + 1:
        raise ValueError("alphabet %r needs %d states (letters + blank), the matrix has %s"
                         % (alphabet_, len(alphabet_) + 1, y.shape[1:] if y.ndim > 1 else y.shape))
    if not 1 <= len(alphabet_) <= 4:
        raise NotImplementedError("the GPU searches handle alphabets of one to four letters; got %r" % (alphabet_,))


def _spell(seq, alphabet_):
    """the kernels spell base index k as "ACGT"[k]; translate to the caller's alphabet"""
    return seq if alphabet_ == "ACGT" else seq.translate(str.maketrans("ACGT"[:len(alphabet_)], alphabet_))


def _indices(label_, alphabet_):
    return label_ if alphabet_ == "ACGT" else label_.translate(str.maketrans(alphabet_, "ACGT"[:len(alphabet_)]))


def cpp_beam_search(y_, beam_width_=25, alphabet_="ACGT", model_="ctc"):
    """decoding_cpp.pyx:88-103 -> beam_search (BeamSearch.h:400) -> beam_search_ (:18-58)."""
    y = _prep(y_)
    _check_alphabet(alphabet_, y)
    seqs, _, _ = batch.beam_search_batch([y], beam_width_, model_)
    return _spell(seqs[0], alphabet_)


def cpp_beam_search_2d(y1_, y2_, envelope_ranges_=None, beam_width_=25, alphabet_="ACGT", model_="ctc", method_="row"):
    """decoding_cpp.pyx:107-139 -> beam_search (BeamSearch.h:411 / :440)."""
    y1, y2 = _prep(y1_), _prep(y2_)
    _check_alphabet(alphabet_, y1)
    if y1.dtype != y2.dtype:
        y1, y2 = y1.astype(np.float64), y2.astype(np.float64)
    env = None if envelope_ranges_ is None else [np.asarray(envelope_ranges_, dtype=np.intc)]
    seqs, _, _ = batch.beam_search_2d_batch([y1], [y2], env, beam_width_, model_, method_)
    return _spell(seqs[0], alphabet_)


def cpp_forward(y_, label_, alphabet_="ACGT", model_="ctc"):
    """decoding_cpp.pyx:49-65 -> forward (PrefixTree.h:710-759): log-probability of `label_` given y_."""
    y = _prep(y_)
    _check_alphabet(alphabet_, y)
    return float(batch.forward_batch([y], [_indices(label_, alphabet_)], model_)[0])


def cpp_viterbi_acceptor(y_, label_, band_size=1000, alphabet_="ACGT"):
    """decoding_cpp.pyx:69-84 -> viterbi_acceptor_poreover (Forward.h:14-121): int array, one entry per timestep,
    len(alphabet_) for blank or the index of the base emitted there."""
    y = _prep(y_)
    _check_alphabet(alphabet_, y)
    paths, st = batch.viterbi_acceptor_batch([y], [_indices(label_, alphabet_)], band_size)
    if st[0] & batch._lib.ST_UNSET_BAND:
        raise RuntimeError("viterbi_acceptor: the label cannot be placed inside the band "
                           "(the reference's traceback does not terminate here, Forward.h:107-116)")
    return paths[0]
