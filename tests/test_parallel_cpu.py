"""N>1 host logic on CPU with the gloo backend, world_size 2 (no GPU): sharding, the single weight broadcast,
and the rank-count independence of which image gets which seed / which slot."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stable_diffusion_burn_b200 import parallel


def test_shard_images_partition():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            got = sorted(i for r in range(world) for i in parallel.shard_images(n, r, world))
            assert got == list(range(n))
    assert parallel.shard_images(64, 3, 8) == list(range(3, 64, 8))
    with pytest.raises(ValueError):
        parallel.shard_images(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weights: rank 0 holds the arena, the others start from zeros
        arena = torch.arange(1000, dtype=torch.float32) * 0.5 if rank == 0 else torch.zeros(1000)
        parallel.broadcast_arena(arena, 0)
        ok_w = bool(torch.equal(arena, torch.arange(1000, dtype=torch.float32) * 0.5))
        # "sampling": image i is a constant plane derived from its seed only (stands in for the device call)
        idx = parallel.shard_images(n_images, rank, world)
        imgs = np.stack([np.full((4, 4, 3), parallel.image_seed(1234, i) % 251, np.uint8) for i in idx]) if idx else np.zeros((0, 4, 4, 3), np.uint8)
        full = parallel.gather_images(imgs, n_images, rank, world)
        if rank == 0:
            q.put((ok_w, full.tolist()))
        else:
            q.put((ok_w, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 8])
def test_world2_gloo_broadcast_and_gather(n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res)
    full = [r[1] for r in res if r[1] is not None][0]
    want = np.stack([np.full((4, 4, 3), (1234 + i) % 251, np.uint8) for i in range(n_images)])
    assert np.array_equal(np.array(full, np.uint8), want)  # same images as a 1-rank run would produce
