"""CPU, world_size 2 over gloo: the batch sharding + single all-gather of the multi-GPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from afldm_b200 import parallel


def test_shard_bounds_cover_the_batch():
    for n in (1, 2, 7, 16, 17, 128):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [parallel.shard_bounds(16, r, 8) for r in range(8)][3] == (6, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(0)
        latents = torch.randn(total, 4, 8, 8, generator=g)          # every rank draws the same seed-0 batch
        mine = parallel.shard_batch(latents)
        lo, hi = parallel.shard_bounds(total, rank, world_size)
        assert mine.shape[0] == hi - lo and torch.equal(mine, latents[lo:hi])
        result = mine * 2.0 + 1.0                                   # stands in for denoise + decode
        full = parallel.gather_frames(result, total)
        q.put((rank, torch.equal(full, latents * 2.0 + 1.0), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shard_and_gather_world2_even():
    for rank, ok, shape in _run(16):
        assert ok and shape == (16, 4, 8, 8)


def test_shard_and_gather_world2_ragged():
    for rank, ok, shape in _run(5):
        assert ok and shape == (5, 4, 8, 8)
