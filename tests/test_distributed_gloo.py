"""world_size-2 gloo test of the N>1 host logic (CPU): each rank produces the SUM buffer of ITS Philox
sample shard (with the CPU oracle standing in for the device kernels, tests only), the buffers are reduced
to rank 0 with the same reduce_sum_to_root() the GPU path uses, and rank 0 must hold the single-rank image."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, total_spp, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpppathtracer_b200 import scenes, distributed as ptd
    from oracle.ref_host import PortOracle
    sc = scenes.demo_scene(48, 27)
    orc = PortOracle()
    orc.load_scene(sc)
    orc.set_sky(scenes.synthetic_sky(64, 64))
    first, count = ptd.shard_samples(total_spp, world, rank)
    img, _ = orc.render(sc.camera, count, 8, seed=9, first_sample=first)
    buf = np.concatenate([img, np.full(img.shape[:2] + (1,), float(count), np.float32)], axis=2)
    t = torch.from_numpy(buf.copy())
    ptd.reduce_sum_to_root(t, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_spp_sharded_reduce_equals_single_rank(tmp_path):
    world, total_spp = 2, 6
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total_spp, str(tmp_path)), nprocs=world, join=True)
    from cpppathtracer_b200 import scenes, distributed as ptd
    from oracle.ref_host import PortOracle
    sc = scenes.demo_scene(48, 27)
    orc = PortOracle()
    orc.load_scene(sc)
    orc.set_sky(scenes.synthetic_sky(64, 64))
    full, _ = orc.render(sc.camera, total_spp, 8, seed=9)
    red = np.load(tmp_path / "reduced.npy")
    assert (red[..., 3] == total_spp).all()
    # same samples, different fp32 summation order -> equal to rounding, not bitwise
    assert np.allclose(red[..., :3], full, rtol=1e-6, atol=1e-6)
    assert np.allclose(ptd.mean_from_sum(red), full / total_spp, rtol=1e-6, atol=1e-6)
