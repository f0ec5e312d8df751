"""Multi-GPU plumbing for the base-sharded MSM.

One process per GPU; each rank reduces its base shard to ONE affine point, then a single
all-gather of world_size x 104 B (200 B for G2) and a local fold.  NCCL has no elliptic-curve
reduction operator, so "all-reduce of partial sums" is all-gather + adds; gathering per-rank partial
points instead of bucket arrays is strictly less traffic (104 B vs ~100 MB per rank).

torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests); the additions are libbzk's
host group law (`bzk_g1_add` / `bzk_g2_add`), which needs no GPU context."""
import ctypes as ct

import numpy as np

from . import _lib

G1_BYTES, G2_BYTES = 104, 200


def _add(fn_name, nbytes, a, b):
    lib = _lib.load()
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros(nbytes, dtype=np.uint8)
    st = getattr(lib, fn_name)(ct.c_void_p(a.ctypes.data), ct.c_void_p(b.ctypes.data), ct.c_void_p(out.ctypes.data))
    if st != 0:
        raise _lib.BzkError(st, "group addition")
    return out


def g1_add(a, b):
    return _add("bzk_g1_add", G1_BYTES, a, b)


def g2_add(a, b):
    return _add("bzk_g2_add", G2_BYTES, a, b)


def fold(points, kind="g1"):
    """sum of an [n, 104|200] array of wire images."""
    add = g1_add if kind == "g1" else g2_add
    acc = np.asarray(points[0], dtype=np.uint8)
    for p in points[1:]:
        acc = add(acc, p)
    return acc


def allgather_fold(partial, kind="g1", group=None, device="cpu"):
    """every rank contributes its partial sum image; every rank returns the total."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(partial, dtype=np.uint8)
    nbytes = G1_BYTES if kind == "g1" else G2_BYTES
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint8)).to(device)
    gathered = torch.empty(world * nbytes, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return fold(gathered.cpu().numpy().reshape(world, nbytes), kind)


def shard_range(n, rank, world):
    """contiguous base range [lo, hi) owned by `rank` (contiguous index ranges)."""
    return n * rank // world, n * (rank + 1) // world
