"""Helper for test_cpu_host.py (gloo, world size 2): the bench's strong-scaling step on the host side -- a job's chunks
pulled by two lanes per rank from one queue (a fresh store key per step), records handed to rank 0 through the store --
and run_sharded's behaviour when one rank fails (every rank must come back, nobody may hang in the gather)."""
import datetime
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch.distributed as dist  # noqa: E402

from poreover_b200 import multigpu  # noqa: E402

dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=120))
rank = dist.get_rank()
store = dist.distributed_c10d._get_default_store()
lock = threading.Lock()
G, chunk = 1000, 61
report = {"steps": []}
dist.barrier()  # both ranks at the starting line
for step in range(3):
    q = multigpu.WorkQueue(G, chunk, store, key="strong_q%d" % step)

    def work(lo, hi, lane):
        time.sleep(0.001 * (1 + 2 * rank))  # uneven ranks
        if rank != 0:
            with lock:
                store.set("rec_%d_%d" % (step, lo), bytes([lane]) * (hi - lo))
            return None
        return bytes([lane]) * (hi - lo)

    got = multigpu.drain_queue(q, work, lanes=2, lock=lock)
    if rank == 0:
        have = {lo: r for lo, hi, r in got}
        covered = 0
        for k in range(len(q.bounds) - 1):
            lo, hi = q.bounds[k], q.bounds[k + 1]
            blob = have[lo] if lo in have else store.get("rec_%d_%d" % (step, lo))
            assert len(blob) == hi - lo
            covered += len(blob)
        report["steps"].append({"covered": covered, "own_chunks": len(got), "chunks": len(q.bounds) - 1})
    dist.barrier()

# a failure on rank 1 only: both ranks must leave run_sharded with an exception (rank 0 through the gather)
def bad_work(items):
    if rank == 1:
        raise ValueError("boom")
    time.sleep(0.2)  # rank 0 is slow: rank 1 is sure to pull a chunk (and fail on it)
    return [x for x in items]

try:
    multigpu.run_sharded(list(range(40)), [1] * 40, bad_work, chunk=5, group=dist.group.WORLD, store=store)
    report_err = None
except Exception as e:  # noqa: BLE001
    report_err = "%s: %s" % (type(e).__name__, e)
errs = [None, None]
dist.all_gather_object(errs, report_err)
# a second queue in the same process group starts from its own counter (unique keys)
out = multigpu.run_sharded(list(range(23)), [1] * 23, lambda items: [x * 2 for x in items], chunk=4, group=dist.group.WORLD, store=store)
if rank == 0:
    report["errors"] = errs
    report["second_run"] = out
    with open(sys.argv[1], "w") as f:
        json.dump(report, f)
dist.destroy_process_group()
