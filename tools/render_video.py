#!/usr/bin/env python
"""Config C5 from the command line: `python tools/render_video.py --frames 120 --spp 64 [--out DIR]`
(under torchrun: frames round-robin over ranks, no collective)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import scenes, video

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=120); ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spheres", type=int, default=1000); ap.add_argument("--out", default=None)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
sc = scenes.sphere_field(a.spheres, a.width, a.height)
st = video.render_video(sc, scenes.synthetic_sky(1280, 1280), a.frames, a.spp, rank=rank, world=world, out_dir=a.out)
paths = a.width * a.height * a.spp * len(st)
print(json.dumps({"rank": rank, "frames": len(st), "mpaths_s_device": paths / sum(s["render_ms"] for s in st) / 1e3,
                  "mpaths_s_wall": paths / sum(s["wall_ms"] for s in st) / 1e3,
                  "build_ms_mean": sum(s["build_ms"] for s in st) / len(st)}))
