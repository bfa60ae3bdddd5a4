"""Multi-GPU sharding of the bake path: one process per GPU, the scene BVH
replicated, sample points sharded by contiguous global index ranges, and one exchange step —
the per-sample AO floats — over torch.distributed (NCCL on GPUs; gloo in the CPU tests).

Ray RNG streams are functions of the *global* sample index and stratum only, so the gathered
AO array is bit-identical to a single-GPU run whatever the rank count.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r of R owns the global sample range [floor(r*n/R), floor((r+1)*n/R))."""
    return (rank * total) // world, ((rank + 1) * total) // world


def owned_mask(total: int, part: int, num_parts: int, block_samples: int = 16384) -> np.ndarray:
    """Samples traced by `part` under the interleaved partition (aobake_compute_ao_interleaved):
    super-blocks of block_samples samples, dealt round-robin."""
    return (np.arange(total, dtype=np.int64) // block_samples) % num_parts == part


def gather_shards_(full, world: int, group=None):
    """In-place exchange: `full` is a 1-D tensor of all samples' AO in which this rank has
    filled its own shard_range; afterwards every rank holds every shard.  Ragged shards are
    sent as one broadcast per owner (no padding, no staging copies)."""
    import torch.distributed as dist
    n = full.numel()
    for r in range(world):
        b, e = shard_range(n, r, world)
        if e > b:
            dist.broadcast(full[b:e], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return full


class _DevArray:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, device: int):
    import torch
    if n == 0:
        return torch.empty(0, dtype=torch.float32, device=f"cuda:{device}")
    return torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{device}")


class DistributedBaker:
    """bake::computeAO + mapAOToVertices across the ranks of an initialised process group."""

    def __init__(self, baker, rank: int, world: int, device: int, group=None):
        self.bk, self.rank, self.world, self.device, self.group = baker, rank, world, device, group

    def compute_ao(self, rays_per_sample: int, scene_offset: float, scene_maxdistance: float, gather: bool = True,
                   download: bool = False, interleave: bool = True, block_samples: int = 16384) -> Optional[np.ndarray]:
        """interleave=True (default): rank r traces the 64k-sample super-blocks with index % R == r
        (even load on scenes whose regions differ in traversal cost) and the resident ao[] arrays
        are summed in place with one all-reduce — every other rank contributes exact zeros, so the
        result is bit-identical to a single-GPU bake.  interleave=False: contiguous ranges and one
        broadcast per owner."""
        import torch
        import torch.distributed as dist
        total = self.bk.num_samples
        if interleave and self.world > 1:
            self.bk.compute_ao_interleaved(self.rank, self.world, rays_per_sample, scene_offset, scene_maxdistance,
                                           block_samples)
        else:
            b, e = shard_range(total, self.rank, self.world)
            self.bk.compute_ao(rays_per_sample, scene_offset, scene_maxdistance, download=False, begin=b, end=e)
        if gather and self.world > 1:
            ptr, n = self.bk.ao_device_ptr()
            full = device_tensor(ptr, n, self.device)      # the context's resident ao[] — exchanged in place
            self.bk.synchronize()
            if interleave:
                dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            else:
                gather_shards_(full, self.world, self.group)
            torch.cuda.synchronize(self.device)
        if download:
            import torch
            ptr, n = self.bk.ao_device_ptr()
            return device_tensor(ptr, n, self.device).cpu().numpy()
        return None
