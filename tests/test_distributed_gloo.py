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


class _HostRenderer:
    """Stand-in with the Renderer methods render_sharded() uses; the SUM buffer is a CPU tensor filled by the oracle."""

    def __init__(self, sc, sky):
        from oracle.ref_host import PortOracle
        self.orc = PortOracle(); self.orc.load_scene(sc); self.orc.set_sky(sky)
        self.buf = None

    def ensure_frame(self, w, h):
        if self.buf is None or self.buf.shape[:2] != (h, w):
            self.buf = torch.zeros((h, w, 4), dtype=torch.float32)

    def reset_accum(self):
        if self.buf is not None:
            self.buf.zero_()

    def render(self, cam, spp, max_depth, seed, first_sample=0, sync=True):
        img, _ = self.orc.render(cam, spp, max_depth, seed=seed, first_sample=first_sample)
        self.buf[..., :3] += torch.from_numpy(img)
        self.buf[..., 3] += float(spp)

    def accum_tensor(self):
        if self.buf is None:
            raise RuntimeError("no frame rendered yet")        # what pt_accum_device_ptr answers before the first frame
        return self.buf

    def synchronize(self):
        pass


def _worker_sharded(rank, world, port, total_spp, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpppathtracer_b200 import scenes, distributed as ptd
    sc = scenes.demo_scene(48, 27)
    r = _HostRenderer(sc, scenes.synthetic_sky(64, 64))
    t = ptd.render_sharded(r, sc.camera, total_spp, 8, seed=9, rank=rank, world=world)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_empty_shard_still_joins_the_reduce(tmp_path):
    """world 2, one sample in total: rank 1 renders nothing but must own a zeroed SUM buffer and enter the collective
    (it used to raise 'no frame rendered yet' before the reduce and leave rank 0 waiting in it)."""
    world, total_spp = 2, 1
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_sharded, args=(world, port, total_spp, str(tmp_path)), nprocs=world, join=True)
    from cpppathtracer_b200 import scenes
    from oracle.ref_host import PortOracle
    sc = scenes.demo_scene(48, 27)
    orc = PortOracle(); orc.load_scene(sc); orc.set_sky(scenes.synthetic_sky(64, 64))
    full, _ = orc.render(sc.camera, 1, 8, seed=9)
    red = np.load(tmp_path / "reduced.npy")
    assert (red[..., 3] == 1).all() and np.array_equal(red[..., :3], full)
