"""Frame sharding across the GPUs of one box: one process per GPU (torch.distributed), scene replicated, the frame split by
image rows or by sample-index range, and ONE reduction of the film accumulators per frame (NCCL over NVLink on GPUs; the
same code runs on gloo/CPU tensors in tests).  The reference has no distributed path; the only shared
output of the hot path is the per-pixel Sensor accumulator (sensor.rs:76-77), so a sum of accumulators is the whole exchange.

  mode "tile": rank r renders rows y with y % world == r and ALL samples -> every pixel is summed in the reference's
               sample order on one GPU, other ranks contribute exact zeros: the reduced film is bitwise equal to 1 GPU.
  mode "spp" : rank r renders samples [r*spp/world, (r+1)*spp/world) of every pixel -> best balance; per-pixel float sums are
               re-associated across ranks (tolerance-level equality).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class Shard:
    row_offset: int = 0
    row_stride: int = 0
    spp_begin: int = 0
    spp_end: int = 0

    def as_kwargs(self):
        return dict(row_offset=self.row_offset, row_stride=self.row_stride, spp_begin=self.spp_begin, spp_end=self.spp_end)


def shard_plan(rank: int, world: int, mode: str, spp: int, spp_begin: int = 0, spp_end: int | None = None) -> Shard:
    """The slice of the frame rank `rank` of `world` renders.  [spp_begin, spp_end) restricts the whole job to a sample window."""
    spp_end = spp if spp_end is None else spp_end
    if not (0 <= rank < world) or not (0 <= spp_begin <= spp_end <= spp):
        raise ValueError("bad shard request")
    if mode == "tile":
        return Shard(row_offset=rank, row_stride=world, spp_begin=spp_begin, spp_end=spp_end)
    if mode == "spp":
        n = spp_end - spp_begin
        b = spp_begin + (n * rank) // world
        e = spp_begin + (n * (rank + 1)) // world
        return Shard(row_offset=0, row_stride=0, spp_begin=b, spp_end=e)
    raise ValueError(f"unknown shard mode {mode!r}")


def reduce_film(acc, dst: int = 0, all_ranks: bool = False):
    """Sum the film accumulators over ranks (in place on `dst`, or everywhere with all_ranks).  One collective per frame."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return acc
    if all_ranks:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
    return acc


def render_sharded(image, sampler, mode: str = "tile", device_acc=None, stream=None, spp_window=None, max_slots: int = 0):
    """Render this rank's shard of `image` (a RendererImage) into a device accumulator (torch CUDA tensor [H, W, 3] f32, zeroed
    by the caller), then reduce to rank 0.  Returns the (reduced on rank 0) accumulator tensor."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    from . import capi
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    r = image.renderer
    spp = r.args.spp
    w0, w1 = spp_window if spp_window else (0, spp)
    shard = shard_plan(rank, world, mode, spp, w0, w1)
    ctx = r.args.scene.ctx
    if device_acc is None:
        device_acc = torch.zeros((image.height, image.width, 3), dtype=torch.float32, device=f"cuda:{torch.cuda.current_device()}")
    p = r.params(sampler, max_slots=max_slots, **shard.as_kwargs())
    if p.spp_begin == 0 and p.spp_end == 0:
        p.spp_end = spp  # an explicit full range (0,0 would also mean "all")
    if shard.spp_begin == shard.spp_end:
        pass  # nothing to render on this rank (more ranks than samples): contributes zeros
    else:
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        ctx.check(ctx.lib.tcpt_render_device(ctx.handle, C.byref(p), C.c_void_p(device_acc.data_ptr()), C.c_void_p(s)))
        image.stats = ctx.stats()
    reduce_film(device_acc, dst=0)
    return device_acc
