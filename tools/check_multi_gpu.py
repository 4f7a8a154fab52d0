#!/usr/bin/env python
"""Multi-GPU determinism + config C4 (run under torchrun, one rank per GPU):
  1. spp-sharded render (shard_samples) + one NCCL reduce == the same samples rendered by rank 0 alone (fp32 tolerance);
  2. C4: 3840x2160, 1024 spp progressive, sample ranges split across the ranks, one reduce at the end -> Mpaths/s."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes, distributed as ptd

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sky = scenes.synthetic_sky(1280, 1280)
out = {"world": world}

# 1. determinism
sc = scenes.sphere_field(1000, 960, 540)
r = api.Renderer(sc, sky)
r.set_stream(torch.cuda.current_stream().cuda_stream)
total_spp = 8 * world + 3          # not divisible by the world size on purpose
t = ptd.render_sharded(r, sc.camera, total_spp, 8, seed=11, rank=rank, world=world)
if rank == 0:
    sharded = t.cpu().numpy().reshape(540, 960, 4).copy()
    r.reset_accum(); r.render(sc.camera, total_spp, 8, seed=11)
    alone = ptd.accum_tensor(r).cpu().numpy().reshape(540, 960, 4)
    out["determinism"] = {"total_spp": total_spp, "samples_ok": bool((sharded[..., 3] == total_spp).all()),
                          "max_abs_diff_of_sums": float(np.abs(sharded[..., :3] - alone[..., :3]).max()),
                          "max_rel_diff": float((np.abs(sharded[..., :3] - alone[..., :3]) / np.maximum(alone[..., :3], 1e-3)).max()),
                          "bitwise_equal_fraction": float((sharded[..., :3] == alone[..., :3]).mean())}
r.close()

# 2. C4
sc = scenes.sphere_field(1000, 3840, 2160)
r = api.Renderer(sc, sky)
r.set_stream(torch.cuda.current_stream().cuda_stream)
spp = 1024
ptd.render_sharded(r, sc.camera, 16 * world, 8, seed=5, rank=rank, world=world)        # warm-up (allocations, NCCL)
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ptd.render_sharded(r, sc.camera, spp, 8, seed=5, rank=rank, world=world, sync=False)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    paths = 3840 * 2160 * spp
    out["c4"] = {"ms": float(ms.item()), "gpaths_per_s": paths / float(ms.item()) / 1e6, "paths": paths,
                 "what": "3840x2160 x 1024 spp, depth 8, sample ranges sharded over the ranks, one NCCL reduce of the 133 MB SUM buffer"}
    print(json.dumps(out))
r.close()
if world > 1: dist.destroy_process_group()
