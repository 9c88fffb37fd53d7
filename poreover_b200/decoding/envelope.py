"""Drop-in for poreover/decoding/envelope.py: alignment columns and the per-timestep envelope.

build_envelope runs on the GPU (envelope_kernel in poreover_b200/csrc/align.cu), fused with the column
extraction; get_alignment_columns is kept as the cheap host-side accessor the reference exposes."""
import numpy as np

from .. import batch


def get_alignment_columns(alignment):
    """envelope.py:26-44: list of (label, x_index, y_index) per alignment column."""
    x_index = -1
    y_index = -1
    alignment_col = []
    rows = np.asarray(alignment)
    for (x, y) in rows.T:
        if x != '-':
            x_index += 1
        if y != '-':
            y_index += 1
        label = 'i' if x == '-' else ('d' if y == '-' else 'm')
        alignment_col.append((label, x_index, y_index))
    return alignment_col


def _rows_from_columns(alignment_col):
    """Rebuild gap patterns from (label, x, y) columns: only gap/non-gap matters to the envelope."""
    r1 = "".join('-' if c[0] == 'i' else 'N' for c in alignment_col)
    r2 = "".join('-' if c[0] == 'd' else 'N' for c in alignment_col)
    return r1, r2


def build_envelope(y1, y2, alignment_col, sequence_to_signal1, sequence_to_signal2, padding=150):
    """envelope.py:46-87.  Returns an int ndarray of shape (len(y1), 2)."""
    U, V = len(y1), len(y2)
    rows = _rows_from_columns(alignment_col)
    return batch.build_envelope_batch([rows], [sequence_to_signal1], [sequence_to_signal2], [U], [V], padding)[0]
