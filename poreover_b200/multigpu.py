"""Multi-GPU sharding of pair-decode: one process per GPU, a host work queue, results gathered on the host.

Read pairs are independent (the reference treats them so: pair_decode.py:292-297), so there is NO data-path
collective.  Ranks pull chunks of pairs from a shared counter (torch.distributed's TCPStore `add`, i.e. an
atomic fetch-and-add on the host), longest pairs first so the tail is short, and rank 0 collects the
per-pair records with gather_object over the host (gloo) group.  Single process: plain loop.
"""
import os


def dist_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class WorkQueue:
    """Chunked dynamic queue over range(n_items) shared by all ranks of a torch.distributed job.  The first `ramp`
    chunks are a quarter of the size: a pipelined consumer starts its first GPU call sooner.  With `weights` (one
    per item, in queue order) a chunk also ends once it holds `weight_budget`: batches of long reads stay small.
    With `pullers` (how many consumers pull concurrently: ranks x calls in flight) the chunks taper towards the end of
    the queue (guided self-scheduling: at most 1 / (2 pullers) of what is left, never below `min_chunk`), so that the
    consumers finish within a small chunk of each other instead of a full one."""

    def __init__(self, n_items, chunk, store=None, key="poreover_b200_queue", ramp=0, weights=None, weight_budget=None,
                 pullers=0, min_chunk=64):
        self.n, self.chunk, self.store, self.key = n_items, max(1, chunk), store, key
        self._local = 0
        self.bounds = [0]  # the same on every rank: chunk k = [bounds[k], bounds[k+1])
        while self.bounds[-1] < n_items:
            lo = self.bounds[-1]
            small = len(self.bounds) - 1 < ramp
            step = max(1, self.chunk // 4) if small else self.chunk
            if pullers > 0:
                step = max(min(min_chunk, self.chunk), min(step, -(-(n_items - lo) // (2 * pullers))))
            hi = min(n_items, lo + step)
            if pullers > 0 and weights is None and n_items - hi < min_chunk // 2:
                hi = n_items  # no crumb at the end
            if weights is not None and weight_budget:
                budget, acc, k = (weight_budget / 4 if small else weight_budget), 0, lo
                while k < hi and (k == lo or acc + weights[k] <= budget):
                    acc += weights[k]
                    k += 1
                hi = k
            self.bounds.append(hi)

    def next(self):
        if self.store is None:
            k = self._local
            self._local += 1
        else:
            k = self.store.add(self.key, 1) - 1  # atomic on the store's host
        if k + 1 >= len(self.bounds):
            return None
        return self.bounds[k], self.bounds[k + 1]


def drain_queue(queue, work, lanes=2, lock=None):
    """Empty this rank's share of `queue` with `lanes` host threads (GPU calls in flight): each thread pulls the next
    chunk (lo, hi) and runs work(lo, hi, lane).  Returns [(lo, hi, result)] in the order this rank pulled them.  The
    queue is shared by all ranks of the job (fetch-and-add on the store), so a rank that finishes early simply pulls
    more; an exception in any lane stops the others after their current chunk and is re-raised here."""
    import threading
    lock, out, errs = lock or threading.Lock(), [], []  # pass `lock` to share the store connection with other users

    def lane_loop(lane):
        try:
            while not errs:
                with lock:  # one store round trip at a time per process
                    c = queue.next()
                if c is None:
                    return
                r = work(c[0], c[1], lane)
                with lock:
                    out.append((c[0], c[1], r))
        except BaseException as e:  # noqa: BLE001 -- handed to the caller
            errs.append(e)

    if lanes <= 1:
        lane_loop(0)
    else:
        ths = [threading.Thread(target=lane_loop, args=(k,), daemon=True) for k in range(lanes)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    if errs:
        raise errs[0]
    return out


_queue_serial = [0]


def _next_queue_key():
    """A store key no earlier queue of this process group has used (every rank calls run_sharded the same number of
    times, so the serial agrees across ranks)."""
    _queue_serial[0] += 1
    return "poreover_b200_queue_%d" % _queue_serial[0]


def run_sharded(items, cost, process_chunk, chunk=256, group=None, store=None, load_chunk=None, finish_chunk=None,
                cost_budget=None, workers=None):
    """Process `items` across all ranks.  cost[i] orders the queue (descending).  process_chunk(list) ->
    list of results.  With load_chunk, a chunk goes through a pipeline -- payload = load_chunk(list) on a host
    thread, process_chunk(payload) on a second thread (the GPU call), then finish_chunk(its result) -> list of
    results on the calling thread -- so that loading chunk k+2, decoding chunk k+1 and formatting chunk k overlap
    (the queue itself is only touched from the calling thread).  cost_budget caps the summed cost of a chunk;
    workers = GPU calls in flight (default: ingest.Lookahead's).  Returns the full result list in input order on
    rank 0, None elsewhere."""
    from .ingest import Lookahead
    rank, world, _ = dist_info()
    order = sorted(range(len(items)), key=lambda i: -cost[i])
    q = WorkQueue(len(order), chunk, store if world > 1 else None, key=_next_queue_key(), ramp=world if load_chunk else 0,
                  weights=[cost[i] for i in order] if cost_budget else None, weight_budget=cost_budget)
    mine = {}
    if load_chunk is None:
        stream = ((c, process_chunk([items[i] for i in order[c[0]:c[1]]])) for c in iter(q.next, None))
    else:
        stream = Lookahead(q.next, lambda c: load_chunk([items[i] for i in order[c[0]:c[1]]]), process_chunk,
                           workers=workers)
    failure = None
    try:
        for c, res in stream:
            if finish_chunk is not None:
                res = finish_chunk(res)
            for i, r in zip(order[c[0]:c[1]], res):
                mine[i] = r
    except BaseException as e:  # noqa: BLE001
        if world == 1:
            raise
        # the other ranks are (or will be) waiting in the gather below: join it with the error instead of leaving them
        # blocked until the process group times out, then re-raise here; rank 0 raises for everyone
        failure = e
        mine = {"__error__": "rank %d: %s: %s" % (rank, type(e).__name__, e)}
    if world == 1:
        return [mine.get(i) for i in range(len(items))]
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if failure is not None:
        raise failure
    if rank != 0:
        return None
    errors = [part["__error__"] for part in gathered if "__error__" in part]
    if errors:
        raise RuntimeError("decoding failed on another rank: " + "; ".join(errors))
    out = [None] * len(items)
    for part in gathered:
        for i, r in part.items():
            out[i] = r
    return out


# A chunk holds at most this many bytes of (first-read) input files: 4096 pairs of T~5000 are 0.4 GB, so short reads
# are limited by the pair count and long reads (2 MB per file at T = 100k) by this.
CHUNK_BYTES = 1 << 30
# Reads longer than this (file bytes; ~T = 50k) get wide envelopes and node pools of tens of GB per context: one GPU
# call in flight instead of two, as before the pipelined command line.
LONG_READ_BYTES = 1 << 20


def _lanes_for(cost):
    return 1 if cost and max(cost) > LONG_READ_BYTES else None


def _host_group():
    """(world, local rank, store, group) of the host-side (gloo) process group; single process: (1, local, None, None)."""
    rank, world, local = dist_info()
    if world == 1:
        return world, local, None, None
    import datetime
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", timeout=datetime.timedelta(hours=4))
    return world, local, dist.distributed_c10d._get_default_store(), dist.group.WORLD


def decode_files_all_gpus(args, in_files, chunk=4096):
    """`decode` CLI path (decode.py:114-167): files are independent, every rank loads and decodes the chunks it pulls
    on its own GPU; rank 0 gets the sequences in input order (None elsewhere)."""
    from .decoding import decode as dec
    world, local, store, group = _host_group()

    def size_of(p):
        try:
            return os.path.getsize(p)
        except OSError:
            return 0

    def load(paths):
        from . import ingest
        npy = args.basecaller in ('poreover', 'bonito') and all(os.path.splitext(p)[1] == '.npy' for p in paths)
        if npy and args.algorithm in ('viterbi', 'beam'):
            return ingest.load_reads(paths, args.basecaller)  # one packed batch of a single model kind
        return ingest.load_models(paths, args.basecaller)

    def work(payload):
        from . import _lib, batch
        with _lib.borrow_ctx(local) as ctx:  # two of these run at a time, each on its own stream and arena
            if isinstance(payload, batch.ReadBatch):
                if args.algorithm == 'viterbi':
                    return batch.viterbi_batch(payload, args.basecaller, device=ctx, return_maps=False)[0]
                return batch.beam_search_batch(payload, args.beam_width, dec.MODEL_TYPE[args.basecaller], device=ctx)[0]
            return dec.decode_models(payload, args.algorithm, args.beam_width, device=ctx,
                                     window=getattr(args, "window", 400))

    chunk = max(8, min(chunk, -(-len(in_files) // (4 * world))))
    cost = [size_of(p) for p in in_files]
    return run_sharded(in_files, cost, work, chunk, group, store, load_chunk=load, cost_budget=CHUNK_BYTES,
                       workers=_lanes_for(cost))


def decode_pairs_all_gpus(args, pair_list, chunk=4096):
    """CLI path: every rank decodes the chunks it pulls on its own GPU (LOCAL_RANK)."""
    from .decoding import pair_decode as pd
    rank, world, local = dist_info()
    store = group = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo", timeout=datetime.timedelta(hours=4))
        group = dist.group.WORLD
        store = dist.distributed_c10d._get_default_store()

    def cost_of(p):
        try:
            path1, _ = pd._paths(args, p)
            return os.path.getsize(os.path.join(args.dir, path1))
        except OSError:
            return 0

    cost = [cost_of(p) for p in pair_list]
    # small runs: at least ~4 chunks per rank so that every GPU gets work; large runs: 4096-pair batches (a B200 holds
    # 444 pairs in flight, so a batch should be several waves; the next batch is loaded while this one is decoded)
    pd._check_args(args)
    chunk = max(8, min(chunk, -(-len(pair_list) // (4 * world))))
    from . import _lib

    def gpu_stage(payload):
        with _lib.borrow_ctx(local) as ctx:  # two of these run at a time, each on its own stream and arena
            return pd.decode_loaded(args, payload, device=ctx, fmt=False)

    return run_sharded(pair_list, cost, gpu_stage, chunk, group, store, load_chunk=lambda sub: pd.load_pairs(args, sub),
                       finish_chunk=lambda raw: pd.format_decoded(args, raw), cost_budget=CHUNK_BYTES,
                       workers=_lanes_for(cost))
