"""Drop-in for poreover/decoding/decode.py: format loaders, FASTA formatting and the single-read driver.

Loaders keep the reference's numpy preprocessing bit for bit (decode.py:29-112): np.log / log-softmax, the
bonito column permutation and the flip-flop uint8 transform happen on the host exactly as before, so the
arrays that cross the C ABI are the arrays the reference's decoders saw.  The decoding itself is batched:
all reads of a run go to the GPU in one call instead of one process-pool task per file (decode.py:158-162).
"""
import glob
import logging
import os
import sys
from pathlib import Path

import numpy as np

from . import transducer
from .. import batch


def fasta_format(name, seq, width=60):
    """decode.py:20-27: header line, the sequence in lines of `width`, no empty last line unless seq is empty."""
    return '>' + name + '\n' + '\n'.join([seq[i:i + width] for i in range(0, len(seq), width)]) + '\n'


def _logsumexp(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0)
    return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(a - m), axis=axis))


def logit_to_log_likelihood(logits):
    """decode.py:34-39 (scipy.special.logsumexp over axis 2, so logits must be 3-D)."""
    try:
        from scipy.special import logsumexp
    except ImportError:  # pragma: no cover
        logsumexp = _logsumexp
    return (logits.T - logsumexp(logits, axis=2).T).T


def load_logits(file_path, flatten=False):
    """decode.py:41-51"""
    read_reshape = np.load(file_path)
    if np.isclose(np.sum(read_reshape[0]), 1):
        with np.errstate(divide="ignore"):
            read_reshape = np.log(read_reshape)
    else:
        read_reshape = logit_to_log_likelihood(read_reshape)
    if flatten and len(read_reshape.shape) > 2:
        return np.concatenate(read_reshape)
    return read_reshape


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:
        raise ImportError("reading .hdf5 / .fast5 traces needs h5py, which is not installed here") from e


def trace_from_flappie(p):
    """decode.py:53-59"""
    hdf = _h5py().File(p, 'r')
    read_id = list(hdf)[0]
    trace = np.array(hdf[read_id]['trace'])
    hdf.close()
    return trace


def trace_from_guppy(p):
    """decode.py:61-65"""
    hdf = _h5py().File(p, 'r')
    trace = np.array(hdf['/Analyses/Basecall_1D_000/BaseCalled_template/Trace'])
    hdf.close()
    return trace


def model_from_trace(f, basecaller=""):
    """decode.py:67-112: infer the model type from the file and build the transducer."""
    file_name, file_extension = os.path.splitext(f)
    if file_extension == '.npy' and basecaller == 'poreover':
        trace = load_logits(f, flatten=True)
        model = transducer.poreover(trace)
    elif file_extension == '.npy' and basecaller == 'bonito':
        trace = load_logits(f, flatten=True)
        trace = trace[::, [1, 2, 3, 4, 0]]
        model = transducer.bonito(trace)
    elif file_extension == '.csv':
        with np.errstate(divide="ignore"):
            trace = np.log(np.loadtxt(f, delimiter=',', skiprows=1))
        if trace.shape[1] == 5:
            model = transducer.poreover(trace)
        elif trace.shape[1] == 8:
            model = transducer.flipflop(trace)
    elif file_extension == '.hdf5' or basecaller == 'flappie':
        trace = trace_from_flappie(f)
        eps = 0.0000001
        trace = np.log((trace + eps) / (255 + eps))
        model = transducer.flipflop(trace)
    elif file_extension == '.fast5' or basecaller == 'guppy':
        trace = trace_from_guppy(f)
        eps = 0.0000001
        trace = np.log((trace + eps) / (255 + eps))
        model = transducer.flipflop(trace)
    else:
        if basecaller == "":
            print("Problem loading the trace probabilities, please specify where they came from with --basecaller [poreover/guppy/flappie]")
        else:
            print("Problem loading the trace probabilities")
        sys.exit(1)
    return model


MODEL_TYPE = {'poreover': 'ctc', 'bonito': 'ctc_merge_repeats', 'guppy': 'ctc_flipflop', 'flappie': 'ctc_flipflop',
              'flipflop': 'ctc_flipflop'}  # decode.py:172


def decode_models(models, algorithm="viterbi", beam_width=25, device=None, window=400):
    """Decode a list of transducers in as few GPU calls as possible (one per model kind).  `window`: rows per window of
    --algorithm prefix (decode.py:180-188)."""
    out = [None] * len(models)
    by_kind = {}
    for i, m in enumerate(models):
        by_kind.setdefault(m.kind, []).append(i)
    for kind, idx in by_kind.items():
        arrays = [models[i].device_array() for i in idx]
        if algorithm == 'viterbi':
            if kind == 'flipflop':
                seqs = batch.flipflop_viterbi_batch(arrays, device=device, return_maps=False)[0]
            else:
                seqs = batch.viterbi_batch(arrays, kind, device=device, return_maps=False)[0]
        elif algorithm == 'beam':
            if kind == 'flipflop':
                raise NotImplementedError("flip-flop beam search is out of scope (the reference's own test fails)")
            seqs = batch.beam_search_batch(arrays, beam_width, MODEL_TYPE[kind], device=device)[0]
        else:
            # legacy prefix search, window by window (decode.py:179-188); every window of every read in one GPU call
            assert kind == "poreover"  # decode.py:180
            from . import prefix_search
            cuts, owner = [], []
            for j, a in enumerate(arrays):
                a64 = np.asarray(a, dtype=np.float64)
                k = 0
                while k + window < len(a64):
                    cuts.append(a64[k:k + window])
                    owner.append(j)
                    k += window
                cuts.append(a64[k:])
                owner.append(j)
            labels = batch.prefix_search_batch(cuts, batch._lib.PREFIX_CY, device=device)[0]
            seqs = [""] * len(arrays)
            for j, lab in zip(owner, labels):
                seqs[j] += "".join("ACGT"[int(c)] for c in lab)
        for i, s in zip(idx, seqs):
            out[i] = s
    return out


def decode(args):
    """decode.py:114-167.  Same flags and output files; --threads is accepted and ignored (the batch is the
    unit of parallelism on the GPU)."""
    logger = logging.getLogger("poreover_b200")
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter('%(message)s'))
        logger.addHandler(handler)
    logger.setLevel(logging.INFO)
    logger.info('PoreOver decode (B200 backend)')
    in_path = getattr(args, 'in')
    in_files = in_path
    if len(in_path) == 1 and os.path.isdir(in_path[0]):
        file_ext = {'guppy': '.fast5', 'flappie': '.hdf5', 'bonito': '.npy', 'poreover': '.npy'}[args.basecaller]
        in_files = sorted(glob.glob("{}/*{}".format(in_path[0], file_ext)))
    if len(in_files) > 1:
        logger.info("found {} reads to decode".format(len(in_files)))
        logger.info("writing sequences to {0}.fasta".format(args.out))
    # the reference farms files out to a process pool (decode.py:158-162); here they go through the GPU in chunks,
    # and under `torchrun` every rank decodes the chunks it pulls from a host work queue on its own GPU
    from .. import multigpu
    seqs = multigpu.decode_files_all_gpus(args, in_files)
    if seqs is None:
        return  # non-zero ranks of a multi-process launch
    with open(args.out + '.fasta', 'w') as out_fasta:
        for p, sq in zip(in_files, seqs):
            print(fasta_format(Path(p).stem, sq), file=out_fasta)


def decode_helper(in_path, args):
    """decode.py:169-192 for a single file."""
    model = model_from_trace(in_path, args.basecaller)
    sequence = decode_models([model], args.algorithm, args.beam_width, window=getattr(args, "window", 400))[0]
    return fasta_format(Path(in_path).stem, sequence)
