"""Drop-in for poreover/decoding/transducer.py: the model object and its best-path decode, with the
decode running on the GPU (poreover_b200/csrc/viterbi.cu).  Same class names, constructor arguments,
attributes and return values as the reference (transducer.py:11-106)."""
import numpy as np

from .. import _lib, batch


def remove_repeated(s):
    """transducer.py:4-9"""
    out = ''
    for i in range(len(s)):
        if (i == 0) or (s[i - 1] != s[i]):
            out += s[i]
    return out


class transducer:
    """Table of log-probabilities, T x num_states (transducer.py:11-25)."""

    def __init__(self, log_prob, kind, alphabet):
        log_prob = np.asarray(log_prob)
        # The reference widens to float64 (transducer.py:16).  float32 input widens exactly, so the
        # float32 original is kept for the device (half the bytes, identical results); the float64 copy and the
        # all-ones transition table (transducer.py:22) are built on first access -- the batch drivers never read
        # them, and at a few thousand reads per second their allocation was most of the loader's time.
        self._f32 = np.ascontiguousarray(log_prob) if log_prob.dtype == np.float32 else None
        self._lp64 = None if self._f32 is not None else log_prob.astype(np.float64)
        self._transition = None
        self.t_max = len(log_prob)
        self.alphabet = alphabet
        self.num_states = len(alphabet)
        self.kind = kind
        assert (self.num_states == (log_prob.shape[1] if log_prob.ndim > 1 else len(log_prob[0])))

    @property
    def log_prob(self):
        if self._lp64 is None:
            self._lp64 = self._f32.astype(np.float64)
        return self._lp64

    @log_prob.setter
    def log_prob(self, value):
        self._lp64 = value
        self._f32 = None  # an assigned table replaces the loader's float32 original

    @property
    def transition(self):
        if self._transition is None:
            self._transition = np.ones((self.t_max, self.num_states))
        return self._transition

    @transition.setter
    def transition(self, value):
        self._transition = value

    def __getitem__(self, i):
        return self.log_prob.__getitem__(i)

    def device_array(self):
        """The array handed to the C ABI: float32 when that is exact, else float64."""
        if self._f32 is not None:
            return self._f32
        return np.ascontiguousarray(self.log_prob)

    def _reverse(self, perm):
        if self._f32 is not None:
            self._f32 = np.ascontiguousarray(self._f32[::-1, perm])
            self._lp64 = None
        else:
            self.log_prob = self.log_prob[::-1, perm]

    def argmax_decode(self, return_path=False):
        """transducer.py:27-33 (repeats kept, blanks dropped: the 'poreover' rule)"""
        seqs, _, paths, _ = batch.viterbi_batch([self.device_array()], "poreover", return_path=True)
        if return_path:
            return seqs[0], paths[0]
        return seqs[0]

    def viterbi_decode(self, return_path=False):
        raise NotImplementedError

    def sequence_mapping(self):
        """sequence_to_signal of get_sequence_mapping (pair_decode.py:114-142), from the same kernel launch."""
        seqs, maps, paths, st = batch.viterbi_batch([self.device_array()], self.kind, return_path=True)
        m = maps[0]
        if st[0] & _lib.ST_MAPPING_WRAP:
            m = m[1:]  # path[0] == path[-1]: the reference's negative-index wrap drops the first base
        return m

    def __repr__(self):
        return 'transducer(kind=%s, alphabet=%s, t_max=%s)' % (self.kind, self.alphabet, self.t_max)


class poreover(transducer):
    def __init__(self, log_prob, alphabet="ACGT"):
        super().__init__(log_prob, 'poreover', np.array(list(alphabet) + ['']))

    def reverse_complement(self):
        # (A,C,G,T,-)/(0,1,2,3,4) => (T,G,C,A,-)/(3,2,1,0,4)   transducer.py:68-70
        self._reverse([3, 2, 1, 0, 4])

    def viterbi_decode(self, return_path=False):
        return self.argmax_decode(return_path)


class bonito(transducer):
    def __init__(self, log_prob, alphabet="ACGT"):
        super().__init__(log_prob, 'bonito', np.array(list(alphabet) + ['']))

    def reverse_complement(self):
        self._reverse([3, 2, 1, 0, 4])  # transducer.py:79-81

    def viterbi_decode(self, return_path=False):
        """transducer.py:83-89: argmax, collapse runs, drop blanks; returns the UNcollapsed argmax path."""
        seqs, _, paths, _ = batch.viterbi_batch([self.device_array()], "bonito", return_path=True)
        if return_path is True:
            return seqs[0], paths[0]
        return seqs[0]


class flipflop(transducer):
    def __init__(self, log_prob):
        super().__init__(log_prob, 'flipflop', np.array(['A', 'C', 'G', 'T', 'a', 'c', 'g', 't']))
        self.transition = np.array([
            [1, 1, 1, 1, 1, 0, 0, 0],
            [1, 1, 1, 1, 0, 1, 0, 0],
            [1, 1, 1, 1, 0, 0, 1, 0],
            [1, 1, 1, 1, 0, 0, 0, 1],
            [1, 1, 1, 1, 1, 0, 0, 0],
            [1, 1, 1, 1, 0, 1, 0, 0],
            [1, 1, 1, 1, 0, 0, 1, 0],
            [1, 1, 1, 1, 0, 0, 0, 1]
        ])

    def reverse_complement(self):
        self._reverse([3, 2, 1, 0, 7, 6, 5, 4])  # transducer.py:104-106

    def viterbi_decode(self, return_path=False):
        """transducer.py:35-59 on the GPU (FP64 8-state max-sum DP)."""
        seqs, _, paths = batch.flipflop_viterbi_batch([np.ascontiguousarray(self.log_prob)], return_path=True)
        if return_path:
            return seqs[0], paths[0]
        return seqs[0]
