"""Host-side multi-GPU logic on CPU: contiguous frame shards, batch cutting and the result gather
(world_size 2, gloo)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from instance_stixels_b200 import sharding
from instance_stixels_b200._lib import SECTION_DTYPE


def test_shards_partition_the_stream():
    for n in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(n, world, r)
                assert 0 <= a <= b <= n
                seen.extend(range(a, b))
            assert seen == list(range(n))
            sizes = [sharding.shard_range(n, world, r)[1] - sharding.shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_batches_cover_shard():
    got = [i for a, b in sharding.batches(10, 75, 16) for i in range(a, b)]
    assert got == list(range(10, 75))
    assert all(b - a <= 16 for a, b in sharding.batches(10, 75, 16))


def test_compact_sections_roundtrip():
    rng = np.random.default_rng(0)
    sec = np.zeros((5, 200), dtype=SECTION_DTYPE)
    sec["type"] = -1
    lens = [0, 3, 199, 1, 7]
    for c, n in enumerate(lens):
        sec["type"][c, :n] = rng.integers(0, 3, n)
        sec["vB"][c, :n] = rng.integers(0, 100, n)
    flat = sharding.compact_sections(sec)
    assert len(flat) == sum(lens)
    for c, n in enumerate(lens):
        mine = flat[flat["column"] == c]["section"]
        assert np.array_equal(np.ascontiguousarray(mine).view(np.uint8), np.ascontiguousarray(sec[c, :n]).view(np.uint8))


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(n_frames, world, rank)
    # the per-frame "result" is a variable-length array tagged with its frame id
    local = [np.full(f % 5 + 1, f, dtype=np.int32) for f in range(a, b)]
    out = sharding.gather_frames(local, n_frames, dist, dst=0)
    if rank == 0:
        q.put([x.tolist() for x in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_frames = 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == [[f] * (f % 5 + 1) for f in range(n_frames)]
