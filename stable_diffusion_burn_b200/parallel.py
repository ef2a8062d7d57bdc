"""Multi-GPU plumbing of the sampling path: one process per GPU, images are independent units.

  * image i of a job runs on rank  i mod world  (shard_images)
  * weights: rank 0 owns the fp32 master arena; ONE broadcast ships it at init (broadcast_arena); there is no
    collective on the sampling path itself
  * results: every rank keeps its own images; gather_images concatenates them in image order on rank 0.

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_images(n_images: int, rank: int, world: int) -> list[int]:
    """Indices of the images this rank samples (round-robin, identical on every rank)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_images, world))


def image_seed(base_seed: int, image_index: int) -> int:
    """Per-image RNG stream (seed = 1234 + image index) — independent of the rank count."""
    return base_seed + image_index


def broadcast_arena(arena, src: int = 0):
    """In-place broadcast of the flat fp32 weight arena (a torch tensor view of sdb_weight_arena)."""
    import torch.distributed as dist
    dist.broadcast(arena, src)
    return arena


def broadcast_weights(ctx, rank: int, world: int):
    """The ONE collective of the path, issued by the library (sdb_broadcast_weights): rank 0 creates an ncclUniqueId, the 128
    bytes travel over the process group that already exists (any backend), every rank then joins the library's own communicator,
    which broadcasts the fp32 master arena + the per-norm eps table from rank 0 and is destroyed."""
    import torch
    import torch.distributed as dist
    uid = None
    if rank == 0:
        try:
            uid = ctx.nccl_unique_id()
        except Exception as e:  # libnccl.so.2 not loadable by the library: every rank takes the fallback below
            import sys
            print(f"[sdb200] library NCCL unavailable ({e}); broadcasting the weight arena through torch.distributed",
                  file=sys.stderr, flush=True)
    box = [uid]
    dist.broadcast_object_list(box, src=0, device=torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None)
    if box[0] is not None:
        ctx.broadcast_weights(box[0], rank, world)
        return
    # fallback (the round-1 path): the same ONE broadcast of the fp32 master arena, issued through the existing process group.
    # Carries tensors only: per-norm eps values of a dump-dir are not shipped (every rank must load the directory itself).
    ptr, nbytes = ctx.weight_arena()

    class _Arena:
        __cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 3}
    arena = torch.as_tensor(_Arena(), device=torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(arena, 0)
    torch.cuda.synchronize()


def gather_images(local_images: np.ndarray, n_images: int, rank: int, world: int):
    """local_images [k,H,W,3] uint8 for shard_images(...) -> [n_images,H,W,3] on rank 0 (None elsewhere).
    NCCL moves device memory only: under that backend the staging buffers live on the rank's current CUDA device."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    k_max = (n_images + world - 1) // world
    h, w, c = local_images.shape[1:]
    pad = torch.zeros((k_max, h, w, c), dtype=torch.uint8, device=dev)
    pad[: local_images.shape[0]] = torch.from_numpy(local_images).to(dev)
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    out = np.zeros((n_images, h, w, c), np.uint8)
    for r in range(world):
        idx = shard_images(n_images, r, world)
        out[idx] = bufs[r][: len(idx)].cpu().numpy()
    return out
