"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, the in-place ragged
exchange, and that sharded AO reproduces the single-rank result bit for bit.  The per-shard
compute stand-in here is the oracle (this is a test; the product computes shards on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optix_prime_baking_b200 import scenes
from optix_prime_baking_b200.multi_gpu import gather_shards_, owned_mask, shard_range


def test_shard_ranges_partition():
    for total in (0, 1, 7, 1000, 3007584, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.oracle_binding import Oracle
    scene, blockers = scenes.config1_sphere(16, 16)
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers)
    total, per = orc.distribute_samples(1, 501)          # odd total: ragged shards
    sb = orc.sample_instances(per, 1)
    b, e = shard_range(total, rank, world)
    ao, _ = orc.compute_ao(sb, 16, off, maxd, begin=b, end=e)
    full = torch.full((total,), -1.0, dtype=torch.float32)
    full[b:e] = torch.from_numpy(ao)
    gather_shards_(full, world)
    np.save(os.path.join(out_dir, f"ao_{rank}.npy"), full.numpy())
    # interleaved partition: owned super-blocks carry AO, everything else exact zeros, one all-reduce
    mask = owned_mask(total, rank, world, 64)
    ao_full, _ = orc.compute_ao(sb, 16, off, maxd)        # stand-in for the per-part GPU launch
    inter = torch.from_numpy(np.where(mask, ao_full, np.float32(0.0)).astype(np.float32))
    dist.all_reduce(inter, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"ao_inter_{rank}.npy"), inter.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tests.oracle_binding import Oracle
    scene, blockers = scenes.config1_sphere(16, 16)
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers)
    total, per = orc.distribute_samples(1, 501)
    sb = orc.sample_instances(per, 1)
    want, _ = orc.compute_ao(sb, 16, off, maxd)
    for r in range(world):
        got = np.load(tmp_path / f"ao_{r}.npy")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        got = np.load(tmp_path / f"ao_inter_{r}.npy")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_owned_masks_partition():
    for total, parts, block in [(10007, 3, 64), (1, 2, 32), (200000, 8, 65536), (131072, 2, 65536)]:
        cover = sum(owned_mask(total, p, parts, block).astype(np.int32) for p in range(parts))
        assert np.all(cover == 1)
