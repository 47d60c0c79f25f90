"""Multi-GPU plumbing for launchers that run one process per GPU (torchrun):
shot partitioning and the single reduce of the stacked images.  The data path
has no other collective; in-process multi-GPU runs use rtm_stack_reduce (C ABI) instead."""
from __future__ import annotations

import numpy as np


def partition_shots(nrec: int, world: int, rank: int) -> range:
    """Contiguous block of shot indices of `rank` (same rule as csrc/driver.cpp)."""
    base, extra = divmod(nrec, world)
    first = rank * base + min(rank, extra)
    return range(first, first + base + (1 if rank < extra else 0))


class _DevBuf:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def stack_tensor(engine):
    """The engine's device-resident stack (up then down) as a torch tensor view."""
    import torch
    ptr, nfl, _ = engine.stack_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    return torch.as_tensor(_DevBuf(ptr, nfl), device=dev)


def reduce_stack(t, nshots: int, dst: int = 0, group=None):
    """Sum the per-rank stacks into rank `dst` (one reduce) and the shot counts alongside.
    `t` is a tensor holding [up_sum, down_sum] (device tensor for NCCL, CPU tensor for gloo).
    Returns the total number of shots on `dst`, else None."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return nshots
    dist.reduce(t, dst=dst, group=group)
    cnt = torch.tensor([nshots], dtype=torch.int64, device=t.device)
    dist.reduce(cnt, dst=dst, group=group)
    return int(cnt.item()) if dist.get_rank(group) == dst else None


def finalize(t, nrec: int, iNorm: int, mod_NX: int, mod_NZ: int):
    """Stack -> image on the host (kernel.cu:1042-1059)."""
    from . import stack_finalize
    a = t.detach().cpu().numpy().astype(np.float32)
    n = mod_NX * mod_NZ
    img, ill = stack_finalize(a[:n], a[n:2 * n], nrec, iNorm)
    return img.reshape(mod_NX, mod_NZ), ill.reshape(mod_NX, mod_NZ)
