"""Drop-in for poreover/decoding/pair_decode.py: the 1D^2 pair driver.

Kept: function names and signatures of the hot-path helpers (get_sequence_mapping, fasta_format,
pair_decode_helper, pair_decode), the argparse Namespace fields, the output files and their formats
(SURVEY.md A.9).  Changed: pairs are not farmed out to a process pool one at a time
(pair_decode.py:292-297); all pairs of a run go through pob_pair_decode in batches, and with several GPUs
the batches are pulled from a host work queue by one process per GPU (poreover_b200/multigpu.py).
--alignment full, --diagonal_envelope, --skip_matches and --single beam run the same kernels stage by
stage (viterbi -> aligner -> envelope -> one batched search over whole pairs or over the boxes between anchors),
with the anchor / box bookkeeping of pair_decode.py:412-452 on the host.
Out of scope here, as in SURVEY.md section 2: --method split/align, --algorithm prefix (they raise
NotImplementedError).
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
    ("algorithm", "beam", "pair-decode --algorithm prefix cannot run in the reference either (pair_decode.py:224 asserts "
                          "kind == 'poreover' on a value that is 'ctc'); the legacy searches themselves are in "
                          "poreover_b200.decoding.prefix_search and `decode --algorithm prefix`"),
)


def _check_args(args):
    for name, ok, why in _UNSUPPORTED:
        if getattr(args, name, ok) != ok:
            raise NotImplementedError(why)
    if getattr(args, "beam_search_method", "row_col") not in ("row", "row_col"):
        raise NotImplementedError("--beam_search_method grid is marked 'still testing' in the reference; not built")


def _staged(args):
    """True when a flag asks for something the fused pob_pair_decode call does not do."""
    return (getattr(args, "alignment", "banded") == "full" or getattr(args, "skip_matches", False)
            or getattr(args, "diagonal_envelope", False) or getattr(args, "single", "viterbi") == "beam")


def get_anchors(alignment, matches, indels):
    """pair_decode.py:53-89: column ranges of long runs of matches / insertions / deletions.

    alignment: 2 x C array of single characters.  Run-length encoding of the column states; a mismatch column
    always starts a new run (and never forms an anchor), and -- as in the reference -- a run that reaches the
    last column is never closed, so it yields no anchor."""
    a1, a2 = np.asarray(alignment[0]), np.asarray(alignment[1])
    n = len(a1)
    if n == 0:
        return [], []
    state = np.where(a1 == a2, 0, np.where(a1 == '-', 1, np.where(a2 == '-', 2, 3)))  # mat ins del mis
    new_run = np.ones(n, dtype=bool)
    new_run[1:] = (state[1:] != state[:-1]) | (state[1:] == 3)
    starts = np.flatnonzero(new_run)
    ends = np.append(starts[1:], n)
    names = ('mat', 'ins', 'del')
    ranges, types = [], []
    for s0, e0 in zip(starts[:-1], ends[:-1]):  # the final run is still open when the loop ends
        st = state[s0]
        if st == 3:
            continue
        if e0 - s0 >= (matches if st == 0 else indels):
            ranges.append((int(s0), int(e0)))
            types.append(names[st])
    return ranges, types


def _alignment_to_sequence(alignment):
    """pair_decode.py:402-410: running count of non-gap characters per row (1-based at the first base)."""
    return np.cumsum(np.asarray(alignment) != '-', axis=1)


def _boxes_and_anchors(alignment, s2s1, s2s2, U, V, skip_threshold):
    """pair_decode.py:412-452: anchors (start signal index, sequence) and the boxes of signal between them."""
    a2s = _alignment_to_sequence(alignment)
    anchor_ranges, anchor_type = get_anchors(alignment, matches=skip_threshold, indels=100)
    assert len(anchor_ranges) > 0, \
        'No matches/indels of sufficient length found in alignment. Try decreasing --matches or --indels'
    anchors, boxes = [], []
    for i, (cs, ce) in enumerate(anchor_ranges):
        row = 1 if anchor_type[i] == 'ins' else 0
        anchors.append((int(s2s1[a2s[0, cs]]), ''.join(alignment[row, cs:ce])))
        if i > 0:
            pe = anchor_ranges[i - 1][1]
            boxes.append((int(s2s1[a2s[0, pe]]), int(s2s1[a2s[0, cs]]), int(s2s2[a2s[1, pe]]), int(s2s2[a2s[1, cs]])))
        else:
            boxes.append((0, int(s2s1[a2s[0, cs]]), 0, int(s2s2[a2s[1, cs]])))
    le = anchor_ranges[-1][1]
    boxes.append((int(s2s1[a2s[0, le]]), U, int(s2s2[a2s[1, le]]), V))
    return anchors, boxes


def _decode_pairs_staged(args, meta, m1, m2, kind, device):
    """--alignment full / --diagonal_envelope / --skip_matches: pair_decode.py:357-531 stage by stage."""
    n = len(meta)
    rc2 = np.full(n, 1 if args.reverse_complement else 0, dtype=np.uint8)
    model = decode.MODEL_TYPE[kind]
    method = args.beam_search_method
    out = [None] * n
    U = [len(a) for a in m1]
    V = [len(a) for a in m2]
    if getattr(args, "diagonal_envelope", False):
        # pair_decode.py:497-498; no 1D decoding, the helper returns (consensus fasta, summary) only (:525-527)
        w = args.diagonal_width
        envs = []
        for u_, v_ in zip(U, V):
            mid = (np.arange(u_) / u_ * v_).astype(int)
            envs.append(np.stack([np.maximum(mid - w, 0), np.minimum(mid + w, v_)], axis=1))
        seqs, _, _ = batch.beam_search_2d_batch(m1, m2, envs, args.beam_width, model, method, rc2=rc2, device=device)
        for k, (in_path, path1, path2, _) in enumerate(meta):
            out[k] = (fasta_format('consensus;{};{}'.format(args.method, path1.stem, path2.stem), seqs[k]),
                      {'read1': in_path[0], 'read2': in_path[1]})
        return out
    if getattr(args, "single", "viterbi") == "beam":
        # pair_decode.py:363-370: 1D beam search with cpp_beam_search's DEFAULTS (width 25, model "ctc", whatever
        # the basecaller), then the banded acceptor turns each basecall into a path; mapping on the host
        seq1 = batch.beam_search_batch(m1, 25, "ctc", device=device)[0]
        seq2 = batch.beam_search_batch(m2, 25, "ctc", rc=rc2, device=device)[0]
        # an empty basecall makes the reference's acceptor read label_int[0] (Forward.h:51): that pool task dies and
        # only its pair is lost (pair_decode.py:295 has no error callback); the acceptor gets a placeholder label and
        # the pair is dropped below
        dead = [len(a) == 0 or len(b) == 0 for a, b in zip(seq1, seq2)]
        if any(dead):
            _warn("%d pair(s) dropped: empty --single beam basecall" % sum(dead))
        pth1, sa1 = batch.viterbi_acceptor_batch(m1, [s or "A" for s in seq1], 1000, device=device)
        pth2, sa2 = batch.viterbi_acceptor_batch(m2, [s or "A" for s in seq2], 1000, rc=rc2, device=device)
        map1 = [np.asarray(get_sequence_mapping(p, kind)[0], dtype=np.int64) for p in pth1]
        map2 = [np.asarray(get_sequence_mapping(p, kind)[0], dtype=np.int64) for p in pth2]
        wrap = batch._lib.ST_MAPPING_WRAP
        st1 = np.array([wrap if (len(m) != len(s) or (a & batch._lib.ST_UNSET_BAND) or d) else 0
                        for m, s, a, d in zip(map1, seq1, sa1, dead)])
        st2 = np.array([wrap if (len(m) != len(s) or (a & batch._lib.ST_UNSET_BAND)) else 0
                        for m, s, a in zip(map2, seq2, sa2)])
    else:
        seq1, map1, _, st1 = batch.viterbi_batch(m1, kind, device=device)
        seq2, map2, _, st2 = batch.viterbi_batch(m2, kind, rc=rc2, device=device)
    live = []
    for k, (in_path, path1, path2, _) in enumerate(meta):
        summary = {'read1': in_path[0], 'read2': in_path[1], 'length1': len(seq1[k]), 'length2': len(seq2[k])}
        if abs(len(seq1[k]) - len(seq2[k])) > 1000:  # checked before the mapping assertions (pair_decode.py:372-375)
            summary['skipped'] = 1
            out[k] = [summary]
            continue
        if (st1[k] | st2[k]) & batch._lib.ST_MAPPING_WRAP:
            continue  # the reference's assertion (pair_decode.py:379 / :382) fires and the pool drops the pair
        if len(seq1[k]) == 0 and len(seq2[k]) == 0:
            continue  # 0 / 0 identity: the task dies in the reference
        live.append(k)
    if not live:
        return out
    if getattr(args, "alignment", "banded") == "full":
        alns = batch.align_global_batch([seq1[k] for k in live], [seq2[k] for k in live], device=device)
    else:
        alns = batch.align_banded_batch([seq1[k] for k in live], [seq2[k] for k in live], device=device)
    todo, arrs = [], {}
    for k, al in zip(live, alns):
        in_path = meta[k][0]
        arr = np.array([list(al[0]), list(al[1])])
        ident = np.sum(arr[0] == arr[1]) / len(arr[0])
        summary = {'read1': in_path[0], 'read2': in_path[1], 'length1': len(seq1[k]), 'length2': len(seq2[k]),
                   'sequence_identity': ident}
        if ident < 0.5:
            summary['skipped'] = 1
            out[k] = [summary]
            continue
        summary['skipped'] = 0
        out[k] = summary
        arrs[k] = arr
        todo.append(k)
    if not todo:
        return out
    envs = batch.build_envelope_batch([(''.join(arrs[k][0]), ''.join(arrs[k][1])) for k in todo],
                                      [map1[k] for k in todo], [map2[k] for k in todo], [U[k] for k in todo],
                                      [V[k] for k in todo], padding=args.padding, device=device)
    # one batched search over whole pairs, or over the boxes between anchors (pair_decode.py:512-522)
    it1, it2, itenv, owner = [], [], [], []
    pieces = {k: [] for k in todo}
    for k, env in zip(todo, envs):
        if not getattr(args, "skip_matches", False):
            it1.append(m1[k]); it2.append(m2[k]); itenv.append(env); owner.append((k, 0))
            continue
        # a pair without anchors (assertion, pair_decode.py:431) or with an empty box (IndexError, :516) kills its own
        # pool task in the reference and nothing else: the pair is dropped, the rest of the chunk goes on
        try:
            anchors, boxes = _boxes_and_anchors(arrs[k], map1[k], map2[k], U[k], V[k], args.skip_threshold)
            mine = []
            y2 = m2[k]
            for b in boxes:
                e = env[b[0]:b[1]].copy()
                if len(e) == 0:
                    raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # pair_decode.py:516
                v0, v1 = int(e[0, 0]), int(e[-1, 1])
                # read 2 is a reverse-complement VIEW here: logical rows [v0, v1) are physical rows [V-v1, V-v0)
                y2_ = y2[V[k] - v1:V[k] - v0] if args.reverse_complement else y2[v0:v1]
                mine.append((m1[k][b[0]:b[1]], y2_, e - v0, (k, b[0])))
        except (AssertionError, IndexError) as e:
            _warn("pair %s %s dropped: %s" % (meta[k][0][0], meta[k][0][1], e))
            out[k] = None
            pieces.pop(k)
            continue
        pieces[k].extend(anchors)
        for a_, b_, e_, o_ in mine:
            it1.append(a_); it2.append(b_); itenv.append(e_); owner.append(o_)
    if not it1:
        return out
    seqs, _, st_search = batch.beam_search_2d_batch(it1, it2, itenv, args.beam_width, model, method,
                                                    rc2=np.full(len(it1), 1 if args.reverse_complement else 0, dtype=np.uint8),
                                                    device=device)
    for (k, _), s_ in zip(owner, st_search):
        _warn_status(meta[k][0], int(s_))
    for (k, start), sq in zip(owner, seqs):
        pieces[k].append((start, sq))
    for k in todo:
        if k not in pieces:
            continue
        in_path, path1, path2, _ = meta[k]
        joined = ''.join(x[1] for x in sorted(pieces[k]))  # by first signal index, then sequence (:522)
        out[k] = (fasta_format(in_path[0], seq1[k]) + fasta_format(in_path[1], seq2[k]),
                  fasta_format('consensus;{};{}'.format(path1.stem, path2.stem), joined), out[k])
    return out


def _warn(msg):
    logging.getLogger("poreover_b200").warning("WARNING: " + msg)


def _warn_status(in_path, status):
    """The searches flag what they could not do the reference's way; none of it may pass silently."""
    L = batch._lib
    if status & L.ST_POOL_OVERFLOW:
        _warn("pair %s %s: the search's node pool overflowed (very wide envelope): its consensus may deviate from the "
              "reference's" % (in_path[0], in_path[1]))
    if status & (L.ST_SHORT_BEAM_SKIP | L.ST_UNSET_BAND):
        _warn("pair %s %s: the envelope drives the reference's search into undefined behaviour (BeamSearch.h:309, "
              ":317); a defined rule was used instead" % (in_path[0], in_path[1]))


def _paths(args, in_path):
    path1, path2 = Path(in_path[0]), Path(in_path[1])
    if path1.suffix == ".fast5":  # pair_decode.py:316-319
        path1 = path1.with_suffix(".npy")
    if path2.suffix == ".fast5":
        path2 = path2.with_suffix(".npy")
    return path1, path2


def load_pairs(args, sub):
    """Host stage of a chunk of pairs: the files of both reads -> packed log-probability batches (or, for the staged
    flags, per-read arrays).  No GPU work; runs on the loader threads while the previous chunk is being decoded.
    A pair whose files cannot be loaded (missing, corrupt, wrong shape) is dropped with a warning, like the pool task
    that dies alone in the reference (pair_decode.py:295); the payload carries the chunk positions that survived."""
    try:
        return _load_pairs(args, sub) + (list(range(len(sub))), len(sub))
    except NotImplementedError:
        raise
    except Exception as e:  # noqa: BLE001 -- find the pair(s) that cannot be loaded
        if len(sub) == 1:
            _warn("pair %s dropped: %s: %s" % (" ".join(sub[0][:2]), type(e).__name__, e))
            return [], [], [], None, [], 1
        keep = []
        for k, in_path in enumerate(sub):
            try:
                _load_pairs(args, [in_path])
                keep.append(k)
            except NotImplementedError:
                raise
            except Exception as e1:  # noqa: BLE001
                _warn("pair %s dropped: %s: %s" % (" ".join(in_path[:2]), type(e1).__name__, e1))
        if not keep:
            return [], [], [], None, [], len(sub)
        return _load_pairs(args, [sub[k] for k in keep]) + (keep, len(sub))


def _load_pairs(args, sub):
    from .. import ingest
    meta, files1, files2 = [], [], []
    for in_path in sub:
        path1, path2 = _paths(args, in_path)
        files1.append(os.path.join(args.dir, path1))
        files2.append(os.path.join(args.dir, path2))
        meta.append((in_path, path1, path2))
    if _staged(args):
        models = ingest.load_models(files1 + files2, args.basecaller)
        a, b = models[:len(sub)], models[len(sub):]
        for x, y in zip(a, b):
            assert x.kind == y.kind
        kind = a[0].kind if a else None
        meta = [m + (kind,) for m in meta]
        return meta, [x.device_array() for x in a], [y.device_array() for y in b], kind
    try:
        alloc = ingest.packed_alloc()
        b1 = ingest.load_reads(files1, args.basecaller, alloc=alloc)
        b2 = ingest.load_reads(files2, args.basecaller, rc=1 if args.reverse_complement else 0, alloc=alloc)
    except NotImplementedError:
        raise NotImplementedError("flip-flop pair decoding is out of scope (README.md:97 of the reference)")
    for x, y in zip(b1.kinds, b2.kinds):
        assert x == y
    if b1.dtype != b2.dtype:
        # one side came from float64 files: widen the other (the reference widens everything, transducer.py:16)
        wide = lambda b: b if b.np_dtype == np.float64 else batch.ReadBatch(  # noqa: E731
            [b.data[o:o + l].astype(np.float64) for o, l in zip(b.row_off[:-1], b.lens)], rc=b.rc, layout=b.layout)
        k1, k2 = b1.kinds, b2.kinds
        b1, b2 = wide(b1), wide(b2)
        b1.kinds, b2.kinds = k1, k2
    kind = b1.kinds[0] if b1.kinds else None
    return [m + (kind,) for m in meta], b1, b2, kind


def decode_loaded(args, payload, device=None, fmt=True):
    """GPU stage of a chunk: load_pairs' payload -> pair_decode_helper-style results, in the chunk's order.
    With fmt=False the records are returned raw, for format_decoded on another thread."""
    meta, m1, m2, kind = payload[:4]
    keep, total = (payload[4], payload[5]) if len(payload) > 4 else (list(range(len(meta))), len(meta))
    if len(meta) == 0:
        raw = ("done", [None] * total)
        return raw[1] if fmt else raw
    if kind == 'flipflop':
        raise NotImplementedError("flip-flop pair decoding is out of scope (README.md:97 of the reference)")
    if _staged(args):
        done = _scatter(_decode_pairs_staged(args, meta, m1, m2, kind, device), keep, total)
        return done if fmt else ("done", done)
    res = batch.pair_decode_batch(m1, m2, kind=kind, beam_width=args.beam_width, padding=args.padding,
                                  method=args.beam_search_method, device=device)  # rc and layout travel in the batches
    raw = ("raw", meta, res, keep, total)
    return format_decoded(args, raw) if fmt else raw


def _scatter(results, keep, total):
    """results of the pairs that loaded -> the chunk's order, None where a pair was dropped"""
    if len(keep) == total:
        return results
    out = [None] * total
    for k, r in zip(keep, results):
        out[k] = r
    return out


def format_decoded(args, raw):
    """Host stage after the GPU: pob_pair_decode's records -> the tuples pair_decode_helper returns
    (pair_decode.py:383-398, :525-531)."""
    if not raw:
        return []
    if raw[0] == "done":
        return raw[1]
    _, meta, res = raw[:3]
    keep, total = (raw[3], raw[4]) if len(raw) > 3 else (list(range(len(meta))), len(meta))
    results = [None] * len(meta)
    for k, (r, (in_path, path1, path2, _)) in enumerate(zip(res, meta)):
        summary = {'read1': in_path[0], 'read2': in_path[1], 'length1': r["length1"], 'length2': r["length2"]}
        if r["status"] & batch._lib.ST_SKIPPED_LENGTH:  # checked before the mapping assertions (pair_decode.py:372-375)
            summary['skipped'] = 1
            results[k] = [summary]
            continue
        if r["status"] & (batch._lib.ST_MAPPING_WRAP | batch._lib.ST_EMPTY):
            continue  # the reference's assertion fires and the pool drops the pair silently
        _warn_status(in_path, r["status"])
        summary['sequence_identity'] = r["identity"]
        if r["skipped"]:
            summary['skipped'] = 1
            results[k] = [summary]
            continue
        summary['skipped'] = 0
        results[k] = (
            fasta_format(in_path[0], r["basecall1"]) + fasta_format(in_path[1], r["basecall2"]),
            fasta_format('consensus;{};{}'.format(path1.stem, path2.stem), r["consensus"]),
            summary)
    return _scatter(results, keep, total)


def decode_pairs(args, pair_list, device=None, chunk=4096):
    """Decode [(name1, name2), ...] -> list of pair_decode_helper-style results, in input order.

    Chunks of pairs go through a three-stage pipeline: files are loaded a chunk ahead (ingest.py), two GPU calls are
    in flight on two contexts (the second fills the SMs the first one's last wave leaves idle), and the records of
    the oldest chunk are formatted meanwhile."""
    from .. import ingest, multigpu
    _check_args(args)
    results = [None] * len(pair_list)

    def size_of(in_path):
        try:
            return os.path.getsize(os.path.join(args.dir, _paths(args, in_path)[0]))
        except (OSError, IndexError):
            return 0

    cost = [size_of(p) for p in pair_list]  # chunks of long reads are cut by bytes, and get one GPU call at a time
    q = multigpu.WorkQueue(len(pair_list), chunk, ramp=1, weights=cost, weight_budget=multigpu.CHUNK_BYTES)

    def gpu_stage(payload):
        with batch._lib.borrow_ctx(device) as ctx:  # two of these run at a time, each on its own stream and arena
            return decode_loaded(args, payload, ctx, fmt=False)

    for c, raw in ingest.Lookahead(q.next, lambda c: load_pairs(args, pair_list[c[0]:c[1]]), gpu_stage,
                                   workers=multigpu._lanes_for(cost)):
        res = format_decoded(args, raw)
        results[c[0]:c[0] + len(res)] = res
    return results


def pair_decode_helper(args):
    """pair_decode.py:305-531 for one pair: returns (1D fasta, 2D fasta, summary) or [summary] when skipped."""
    in_path = getattr(args, 'in')
    if len(in_path) != 2:
        logging.getLogger("poreover_b200").error("ERROR: Exactly two reads are required")
    r = decode_pairs(args, [in_path])[0]
    if r is None:
        raise AssertionError("len(sequence_to_signal) != len(basecall) (pair_decode.py:379)")
    return r  # 3-tuple, [summary] when skipped, or (consensus fasta, summary) with --diagonal_envelope (:525-527)


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
        elif len(x) == 2:
            # --diagonal_envelope: the reference's callback drops this shape silently (pair_decode.py:272-283 only
            # handles 3 and 1); the consensus is written here because that is what the flag is for
            print(x[0], file=out_2d_f)
            print('\t'.join(map(str, [x[1].get(k, "") for k in keys])), file=log_f)
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
        r = pair_decode_helper(args)
        seq_2d, summary = (r[1], r[2]) if len(r) == 3 else ((r[0], r[1]) if len(r) == 2 else ("", r[0]))
        print(summary, file=sys.stderr)
        with open(args.out + '.fasta', 'w') as out_fasta:
            print(seq_2d, file=out_fasta)
