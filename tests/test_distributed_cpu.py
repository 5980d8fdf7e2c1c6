"""world_size-2 gloo test of the multi-GPU plumbing (bench.py's N > 1 path): shard parameters per
rank, all-reduce of the accumulation buffers, resolve on rank 0. The per-rank render is done by the
test-only host emulation of the kernel source; on the GPU box the same plumbing drives the CUDA path
(tests/test_gpu_parity.py checks the shard semantics of the kernels themselves)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    pkg = importlib.import_module("sycl-ray-tracer_b200")
    scenes = importlib.import_module("sycl-ray-tracer_b200.scenes")
    import _hostemu
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = scenes.cornell_scene(2)
    emu = _hostemu.Scene(data)
    cam = pkg.Camera((48, 32), data.camera_position, data.camera_direction, data.camera_focal_length)
    if mode == "tile":
        shard = {"rank": rank, "world": world, "tile_size": 16, "seed_salt": 0}
    else:
        shard = {"rank": rank, "world": world, "tile_size": 0, "seed_salt": (rank * 0x9E3779B9) & 0xFFFFFFFF}
    if mode == "progressive":   # config 5: spp slices rendered in batches chained with RT_RENDER_RESUME, one reduction at the end
        f, rays_total = None, 0
        for n in (1, 2, 1):
            f = emu.render(cam, 1, 6, n, shard=shard, resume=f)
            rays_total += f["ray_count"]
        f["ray_count"] = rays_total
    else:
        f = emu.render(cam, 1, 6, 2, shard=shard)
    acc = torch.from_numpy(f["accum"].copy())
    rays = torch.tensor([f["ray_count"]], dtype=torch.int64)
    dist.all_reduce(acc)
    dist.all_reduce(rays)
    if rank == 0:
        np.save(out, acc.numpy())
        np.save(out + ".rays.npy", rays.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tile", "spp", "progressive"])
def test_two_ranks_gloo(tmp_path, mode, pkg, scenes, hostemu):
    out = str(tmp_path / f"acc_{mode}.npy")
    port = 29500 + (os.getpid() % 2000) + {"tile": 0, "spp": 1, "progressive": 2}[mode]
    mp.spawn(_worker, args=(2, port, mode, out), nprocs=2, join=True)
    acc, rays = np.load(out), int(np.load(out + ".rays.npy")[0])
    data = scenes.cornell_scene(2)
    emu = hostemu.Scene(data)
    cam = pkg.Camera((48, 32), data.camera_position, data.camera_direction, data.camera_focal_length)
    if mode == "tile":   # bit-identical to the unsharded frame
        full = emu.render(cam, 1, 6, 2)
        assert np.array_equal(acc.view(np.uint32), full["accum"].view(np.uint32)) and rays == full["ray_count"]
    elif mode == "progressive":   # 1 + 2 + 1 samples per rank = one 4-sample frame per rank, reduced once
        parts = [emu.render(cam, 1, 6, 4, shard={"rank": r, "world": 2, "tile_size": 0, "seed_salt": (r * 0x9E3779B9) & 0xFFFFFFFF})
                 for r in range(2)]
        assert np.array_equal(acc.view(np.uint32), (parts[0]["accum"] + parts[1]["accum"]).view(np.uint32))
        assert (acc[..., 3] == 8).all() and rays == parts[0]["ray_count"] + parts[1]["ray_count"]
    else:                # sum of independently salted streams; sample count adds up
        parts = [emu.render(cam, 1, 6, 2, shard={"rank": r, "world": 2, "tile_size": 0, "seed_salt": (r * 0x9E3779B9) & 0xFFFFFFFF})
                 for r in range(2)]
        assert np.array_equal(acc.view(np.uint32), (parts[0]["accum"] + parts[1]["accum"]).view(np.uint32))
        assert (acc[..., 3] == 4).all() and rays == parts[0]["ray_count"] + parts[1]["ray_count"]
