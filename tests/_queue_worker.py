"""Helper for test_cpu_host.py: run the host work queue under torch.distributed (gloo), world size 2."""
import datetime
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch.distributed as dist  # noqa: E402

from poreover_b200 import multigpu  # noqa: E402

dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=120))
rank = dist.get_rank()
store = dist.distributed_c10d._get_default_store()
items = list(range(103))
cost = [(i * 37) % 11 for i in items]
seen = []


def work(chunk):
    seen.extend(chunk)
    time.sleep(0.002 * len(chunk) * (1 + rank))  # uneven ranks: the queue must balance dynamically
    return [(x, x * x, rank) for x in chunk]


def load(chunk):
    # host stage of the two-stage pipeline: runs one chunk ahead on the prefetch thread
    time.sleep(0.001 * len(chunk))
    return list(chunk)


out = multigpu.run_sharded(items, cost, work, chunk=7, group=dist.group.WORLD, store=store, load_chunk=load)
counts = [None, None]
dist.all_gather_object(counts, len(seen))
if rank == 0:
    with open(sys.argv[1], "w") as f:
        json.dump({"out": out, "counts": counts}, f)
else:
    assert out is None
dist.destroy_process_group()
