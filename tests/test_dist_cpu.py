"""world_size-2 gloo run of the N>1 host path: stream sharding is a partition and the counter all-reduce sums
the per-rank statistics (the only collective of the design)."""
import os
import socket
import sys

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from ofdm_uhd_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.streams_of_rank(64, world, rank)
    lo, hi = sharding.split_frames(1001, world, rank)
    c = torch.tensor([hi - lo, len(mine), rank + 1, 0, 0, 0, 0, 0], dtype=torch.int64)
    sharding.reduce_counters(c)
    if rank == 0:
        torch.save(c, out)
    dist.destroy_process_group()


def test_two_rank_counters(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "c.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    c = torch.load(out)
    assert c.tolist()[:3] == [1001, 64, 3]
