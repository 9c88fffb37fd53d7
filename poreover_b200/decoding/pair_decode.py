"""Drop-in for poreover/decoding/pair_decode.py: the 1D^2 pair driver.

Kept: function names and signatures of the hot-path helpers (get_sequence_mapping, fasta_format,
pair_decode_helper, pair_decode), the argparse Namespace fields, the output files and their formats
(SURVEY.md A.9).  Changed: pairs are not farmed out to a process pool one at a time
(pair_decode.py:292-297); all pairs of a run go through pob_pair_decode in batches, and with several GPUs
the batches are pulled from a host work queue by one process per GPU (poreover_b200/multigpu.py).
Out of scope here, as in SURVEY.md section 2: --method split/align, --skip_matches, --single beam,
--algorithm prefix, --alignment full, --diagonal_envelope (they raise NotImplementedError).
"""
import logging
import os
import sys
from pathlib import Path

import numpy as np

from . import decode
from .. import batch


def fasta_format(name, seq, width=60):
    """pair_decode.py:43-51"""
    return decode.fasta_format(name, seq, width)


def get_sequence_mapping(path, kind):
    """pair_decode.py:114-142: (sequence_to_signal, signal_to_sequence) from a best path.

    Pure index bookkeeping over a path the caller already holds on the host; the fused device pipeline
    computes sequence_to_signal inside the Viterbi kernel instead (poreover_b200/csrc/viterbi.cu)."""
    path = np.asarray(path)
    n = len(path)
    idx = np.arange(n)
    if kind == 'poreover':
        keep = path < 4
    elif kind == 'flipflop':
        keep = np.ones(n, dtype=bool)
        if n:
            keep[1:] = path[1:] != path[:-1]
        s2s = idx[keep]
        sig2seq = (np.cumsum(keep) - 1).tolist() if n else []
        return s2s.tolist(), sig2seq
    elif kind == 'bonito':
        prev = np.roll(path, 1)  # path[i-1] with python's wrap-around at i == 0 (pair_decode.py:136)
        keep = (path != 4) & (path != prev)
    else:
        return [], []
    s2s = idx[keep]
    return s2s.tolist(), list(range(len(s2s)))


_UNSUPPORTED = (
    ("method", "envelope", "--method split/align are deprecated in the reference and not on the GPU path"),
    ("single", "viterbi", "--single beam (re-squiggle) is not on the GPU path yet"),
    ("algorithm", "beam", "--algorithm prefix is the legacy search and not on the GPU path"),
    ("alignment", "banded", "--alignment full is not on the GPU path yet"),
)


def _check_args(args):
    for name, ok, why in _UNSUPPORTED:
        if getattr(args, name, ok) != ok:
            raise NotImplementedError(why)
    if getattr(args, "skip_matches", False) or getattr(args, "diagonal_envelope", False):
        raise NotImplementedError("--skip_matches / --diagonal_envelope are not on the GPU path yet")
    if getattr(args, "beam_search_method", "row_col") not in ("row", "row_col"):
        raise NotImplementedError("--beam_search_method grid is marked 'still testing' in the reference; not built")


def _paths(args, in_path):
    path1, path2 = Path(in_path[0]), Path(in_path[1])
    if path1.suffix == ".fast5":  # pair_decode.py:316-319
        path1 = path1.with_suffix(".npy")
    if path2.suffix == ".fast5":
        path2 = path2.with_suffix(".npy")
    return path1, path2


def decode_pairs(args, pair_list, device=None, chunk=1024):
    """Decode [(name1, name2), ...] -> list of pair_decode_helper-style results, in input order."""
    _check_args(args)
    results = [None] * len(pair_list)
    for c0 in range(0, len(pair_list), chunk):
        sub = pair_list[c0:c0 + chunk]
        m1, m2, meta = [], [], []
        for in_path in sub:
            path1, path2 = _paths(args, in_path)
            a = decode.model_from_trace(os.path.join(args.dir, path1), args.basecaller)
            b = decode.model_from_trace(os.path.join(args.dir, path2), args.basecaller)
            assert a.kind == b.kind
            m1.append(a.device_array())
            m2.append(b.device_array())
            meta.append((in_path, path1, path2, a.kind))
        kind = meta[0][3]
        if kind == 'flipflop':
            raise NotImplementedError("flip-flop pair decoding is out of scope (README.md:97 of the reference)")
        res = batch.pair_decode_batch(m1, m2, kind=kind, beam_width=args.beam_width, padding=args.padding,
                                      method=args.beam_search_method, rc2=bool(args.reverse_complement),
                                      device=device)
        for k, (r, (in_path, path1, path2, _)) in enumerate(zip(res, meta)):
            if r["status"] & (batch._lib.ST_MAPPING_WRAP | batch._lib.ST_EMPTY):
                results[c0 + k] = None  # the reference's assertion fires and the pool drops the pair silently
                continue
            summary = {'read1': in_path[0], 'read2': in_path[1], 'length1': r["length1"], 'length2': r["length2"]}
            if r["status"] & batch._lib.ST_SKIPPED_LENGTH:
                summary['skipped'] = 1
                results[c0 + k] = [summary]
                continue
            summary['sequence_identity'] = r["identity"]
            if r["skipped"]:
                summary['skipped'] = 1
                results[c0 + k] = [summary]
                continue
            summary['skipped'] = 0
            results[c0 + k] = (
                fasta_format(in_path[0], r["basecall1"]) + fasta_format(in_path[1], r["basecall2"]),
                fasta_format('consensus;{};{}'.format(path1.stem, path2.stem), r["consensus"]),
                summary)
    return results


def pair_decode_helper(args):
    """pair_decode.py:305-531 for one pair: returns (1D fasta, 2D fasta, summary) or [summary] when skipped."""
    in_path = getattr(args, 'in')
    if len(in_path) != 2:
        logging.getLogger("poreover_b200").error("ERROR: Exactly two reads are required")
    r = decode_pairs(args, [in_path])[0]
    if r is None:
        raise AssertionError("len(sequence_to_signal) != len(basecall) (pair_decode.py:379)")
    return r


def write_results(args, results, out_1d_f, out_2d_f, log_f):
    """The parent-side callback of the reference (pair_decode.py:272-283)."""
    keys = ["read1", "read2", "length1", "length2", "sequence_identity", "skipped"]
    for x in results:
        if x is None:
            continue
        if len(x) == 3:
            print(x[0], file=out_1d_f)
            print(x[1], file=out_2d_f)
            print('\t'.join(map(str, [x[2].get(k, "") for k in keys])), file=log_f)
        elif len(x) == 1:
            print('\t'.join(map(str, [x[0].get(k, "") for k in keys])), file=log_f)


def pair_decode(args):
    """pair_decode.py:230-303."""
    logger = logging.getLogger("poreover_b200")
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter('%(message)s'))
        logger.addHandler(handler)
    logger.setLevel(logging.DEBUG if getattr(args, "logging", "info") == "debug" else logging.INFO)
    logger.info('PoreOver pair-decode (B200 backend)')
    in_path = getattr(args, 'in')
    if len(in_path) == 1:
        with open(in_path[0], 'r') as read_pairs:
            pair_list = [line.split() for line in read_pairs if line.strip()]
        logger.info("found {} read pairs in {}".format(len(pair_list), in_path[0]))
        logger.info("writing sequences to {0}.1d.fasta and {0}.2d.fasta".format(args.out))
        logger.info("pair alignment statistics saved to {}.log".format(args.out))
        from .. import multigpu
        results = multigpu.decode_pairs_all_gpus(args, pair_list)
        if results is None:
            return  # non-zero ranks of a multi-process launch
        with open(args.out + '.1d.fasta', 'w') as f1, open(args.out + '.2d.fasta', 'w') as f2, \
                open(args.out + '.log', 'w', 1) as lf:
            print('# PoreOver pair-decode', file=lf)
            print('# ' + str(vars(args)), file=lf)
            print('# ' + '\t'.join(["read1", "read2", "length1", "length2", "sequence_identity", "skipped"]), file=lf)
            write_results(args, results, f1, f2, lf)
    else:
        seqs_1d, seq_2d, summary = pair_decode_helper(args)
        print(summary, file=sys.stderr)
        with open(args.out + '.fasta', 'w') as out_fasta:
            print(seq_2d, file=out_fasta)
