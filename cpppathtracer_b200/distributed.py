"""Multi-GPU plumbing for the render path: one process per GPU, torch.distributed for the exchange.

The reference has no multi-GPU code at all.  The path shards two ways:
  * samples-per-pixel: rank g renders the Philox sample range shard_samples(spp, G, g) of EVERY pixel into
    its own fp32 SUM buffer; one reduce(sum) of that buffer to rank 0 per frame is the only exchange;
  * frames of a camera path: frame f belongs to rank frame_owner(f, G); no exchange at all.
Because the stream is keyed by (pixel, sample, bounce), any rank reproduces any sample range exactly, so
the G-GPU image equals the 1-GPU image up to the fp32 summation order of the reduce.
"""
from __future__ import annotations

import numpy as np


def shard_samples(total_spp: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous sample range (first, count) of `rank`; the first `total_spp % world` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world) or total_spp < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_spp, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def frame_owner(frame: int, world: int) -> int:
    """Round-robin frame partition of the video path (config C5)."""
    return frame % world


def frames_of_rank(n_frames: int, world: int, rank: int) -> list[int]:
    return [f for f in range(n_frames) if frame_owner(f, world) == rank]


class _CudaView:
    """Zero-copy view of a device buffer owned by libpt_b200.so, through the CUDA array interface."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def accum_tensor(renderer):
    """torch.Tensor aliasing the renderer's device-resident SUM buffer (float4 per pixel)."""
    import torch
    ptr, n = renderer.accum_device_ptr()
    return torch.as_tensor(_CudaView(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))


def _scene_stream(renderer):
    """The scene's stream as a torch stream, or None when it IS torch's current stream (nothing to order)."""
    import torch
    ptr = int(getattr(renderer, "stream_ptr", 0) or 0)
    cur = torch.cuda.current_stream()
    if ptr == int(cur.cuda_stream):
        return None
    return torch.cuda.ExternalStream(ptr) if ptr else torch.cuda.default_stream()


def reduce_sum_to_root(t, dst: int = 0, group=None, renderer=None):
    """The per-frame exchange: one reduce(sum) of the SUM buffers to rank `dst` (NCCL on GPUs, gloo in tests).

    The collective is enqueued on torch's current stream.  When the renderer works on another stream
    (Renderer.set_stream) the two are ordered with events on the device, both ways: the reduce waits for the
    render that fills the buffer, and the scene's next kernel (pt_reset_accum, the next frame) waits for the reduce."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        other = _scene_stream(renderer) if (renderer is not None and t.is_cuda) else None
        if other is not None:
            import torch
            torch.cuda.current_stream().wait_event(other.record_event())
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
        if other is not None:
            other.wait_event(torch.cuda.current_stream().record_event())
    return t


def mean_from_sum(sum_rgba: np.ndarray) -> np.ndarray:
    """rgb mean from a (.., 4) SUM buffer whose 4th channel counts the accumulated samples."""
    n = np.maximum(sum_rgba[..., 3:4], 1.0)
    return sum_rgba[..., :3] / n


def render_sharded(renderer, cam, total_spp: int, max_depth: int = 8, seed: int = 0, rank: int = 0, world: int = 1,
                   reduce: bool = True, sync: bool = True):
    """Render this rank's sample shard of the frame into its SUM buffer and (optionally) reduce to rank 0.
    Returns the aliasing tensor (valid on every rank; holds the full sum on rank 0 after the reduce).

    A rank whose shard is empty (world > total_spp) still owns a zeroed SUM buffer of the frame's size and still
    enters the collective -- skipping it would leave the other ranks waiting in the reduce."""
    first, count = shard_samples(total_spp, world, rank)
    renderer.ensure_frame(cam.width, cam.height)
    renderer.reset_accum()
    if count > 0:
        renderer.render(cam, count, max_depth, seed, first_sample=first, sync=False)
    # (a host-side stand-in for the renderer -- the gloo test of this logic -- hands out its own tensor)
    t = renderer.accum_tensor() if hasattr(renderer, "accum_tensor") else accum_tensor(renderer)
    if reduce:
        reduce_sum_to_root(t, renderer=renderer)
    if sync:
        renderer.synchronize()
        if reduce:
            import torch
            if t.is_cuda:
                torch.cuda.current_stream().synchronize()
    return t
