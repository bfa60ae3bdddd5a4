"""torchrun worker for the multi-GPU parity test: every rank bakes its shard on its own GPU,
the AO array is exchanged over NCCL, and rank 0 checks it against a single-GPU bake."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optix_prime_baking_b200 import api, scenes  # noqa: E402
from optix_prime_baking_b200.multi_gpu import DistributedBaker  # noqa: E402


def _fresh_id(rank):
    """A new NCCL unique id from rank 0, shipped over torch.distributed (one id per communicator)."""
    t = torch.zeros(128, dtype=torch.uint8, device='cuda')
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(api.Baker.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return t.cpu().numpy().tobytes()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name, (scene, blockers) in {"sphere": scenes.config1_sphere(48, 48),
                                    "instanced": scenes.config4_instanced(2, 20, 20, with_ground=True)}.items():
        off, maxd = scenes.default_distances(scene)
        with api.Baker(device=local) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(2, 10007)
            sb_host = bk.sample_instances(per, 2, download=True)
            ao_c = DistributedBaker(bk, rank, world, local).compute_ao(64, off, maxd, gather=True, download=True, interleave=False)
            ao = DistributedBaker(bk, rank, world, local).compute_ao(64, off, maxd, gather=True, download=True, interleave=True,
                                                                      block_samples=2048)
            assert np.array_equal(ao.view(np.uint32), ao_c.view(np.uint32)), 'interleaved != contiguous sharding'
            # the same exchange done natively by libaobake.so (dlopen'd NCCL); the 128-byte id travels over torch here
            idt = torch.zeros(128, dtype=torch.uint8, device='cuda')
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(api.Baker.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, src=0)
            bk.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
            ao_native = bk.compute_ao_distributed(64, off, maxd)
            # host-buffer forms: 1/N of the scene per rank + all-gather, only the owned sample super-blocks uploaded
            with api.Baker(device=local) as b2:
                b2.comm_init(rank, world, _fresh_id(rank))
                b2.set_scene(scene, blockers, distributed=True)
                b2.set_samples(sb_host, per, distributed=True)
                ao_host = b2.compute_ao_distributed(64, off, maxd)
                v_host = b2.map_ao_to_vertices(api.FILTER_AREA_BASED, distributed=True)
                b2.comm_destroy()
            assert np.array_equal(ao_host.view(np.uint32), ao_native.view(np.uint32)), 'set_scene/set_samples_distributed path differs'

            v_dist = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1, distributed=True)   # instances split over the ranks
            v_repl = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)
            for a, b in zip(v_dist, v_repl):
                assert np.abs(a - b).max() < 1e-4, 'distributed vertex map differs from the replicated one'
            va_dist = bk.map_ao_to_vertices(api.FILTER_AREA_BASED, distributed=True)
            for a, b in zip(v_host, va_dist):
                assert np.abs(a - b).max() < 1e-6
            bk.comm_destroy()
            assert np.array_equal(ao_native.view(np.uint32), ao.view(np.uint32)), 'native NCCL exchange differs'
            v_area = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)
            for a, b in zip(va_dist, v_area):
                assert np.abs(a - b).max() < 1e-6
        if rank == 0:
            with api.Baker(device=local) as ref:
                ref.set_scene(scene, blockers)
                ref.sample_instances(per, 2, download=False)
                want = ref.compute_ao(64, off, maxd)
                want_v = ref.map_ao_to_vertices(api.FILTER_AREA_BASED)
            assert np.array_equal(ao.view(np.uint32), want.view(np.uint32)), f"{name}: sharded AO differs from single-GPU AO"
            for a, b in zip(v_area, want_v):
                assert np.abs(a - b).max() < 1e-5
            print(f"mgpu ok: {name} world={world} samples={total}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
