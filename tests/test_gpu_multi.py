"""Multi-GPU film through the C ABI (tcpt_comm_init / tcpt_render_sharded / tcpt_group_*): N ranks must reproduce the one-GPU film --
bitwise in tile mode (every pixel summed on one rank in the reference's sample order, sensor.rs:76-77), within rounding in spp mode.
The two-rank cases need two visible GPUs and are skipped otherwise (`gpurun --gpus 2`); the one-GPU cases drive the same entry
points (shard rule, device accumulators on a caller's stream) rank by rank into one accumulator."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode,world", [("tile", 2), ("tile", 5), ("spp", 2), ("spp", 8)])
def test_shards_rendered_rank_by_rank_on_one_gpu(bundle_factory, mode, world):
    """What every rank of a `world`-GPU job would render (tcpt_shard_params), through tcpt_render_device on a caller-owned stream into ONE
    device accumulator: the sum the NCCL reduce would produce.  Tile mode: bitwise the full frame; spp mode: the same samples, re-associated."""
    import torch
    from toy_cpu_pathtracing_b200 import capi
    w, h, spp = 200, 150, 32
    b = bundle_factory(10, w, h)
    ctx = b.scene.ctx
    img = b.image("mis", spp)
    full = img.render("sobol").accumulators.copy()
    acc = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.Stream(device=0)
    job = img.renderer.params("sobol")
    paths = 0
    for r in range(world):
        mine = capi.RenderParams()
        assert ctx.lib.tcpt_shard_params(C.byref(job), capi.SHARD_MODES[mode], r, world, C.byref(mine)) == 0
        part = torch.zeros_like(acc)
        ctx.check(ctx.lib.tcpt_render_device(ctx.handle, C.byref(mine), C.c_void_p(part.data_ptr()), C.c_void_p(stream.cuda_stream)))
        paths += ctx.stats()["paths"]
        stream.synchronize()
        acc += part            # rank order = NCCL's tree order does not matter in tile mode: all but one addend are exact zeros
    assert paths == w * h * spp
    got = acc.cpu().numpy()
    if mode == "tile":
        assert np.array_equal(got.view(np.uint32), full.view(np.uint32))
    else:
        assert np.abs(got - full).max() <= 1e-5 * max(1.0, np.abs(full).max())


def test_render_sharded_without_a_communicator_is_render(bundle_factory):
    w, h, spp = 200, 150, 16
    b = bundle_factory(3, w, h)
    a = b.image("mis", spp).render("sobol")
    c = b.image("mis", spp).render_sharded("sobol", mode="tile")
    d = b.image("mis", spp).render_sharded("sobol", mode="spp")
    for other in (c, d):
        assert np.array_equal(a.accumulators.view(np.uint32), other.accumulators.view(np.uint32))
        assert np.array_equal(a.pixels.view(np.uint32), other.pixels.view(np.uint32))
    assert c.stats["reduce_ms"] == 0.0
    # a block of sample indices: tone-mapped over the block's own sample count
    e = b.image("mis", spp).render_sharded("sobol", mode="spp", spp_window=(4, 12))
    f = b.image("mis", spp).render("sobol", spp_begin=4, spp_end=12)
    assert np.array_equal(e.accumulators.view(np.uint32), f.accumulators.view(np.uint32))
    assert e.stats["paths"] == w * h * 8


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_reproduce_the_one_gpu_film(tmp_path):
    """Two processes, one GPU each (torchrun): tcpt_comm_init + tcpt_render_sharded, rank 0 compares with its own one-GPU render."""
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "mgpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_OK tile" in r.stdout and "MGPU_OK spp" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_group_api_one_process_two_gpus(tables):
    """tcpt_group_*: one host process, one context per GPU, one host thread per GPU inside tcpt_group_render, one complete frame out."""
    import toy_cpu_pathtracing_b200 as tp
    from toy_cpu_pathtracing_b200 import capi, scenes
    lib = capi.load_library()
    w, h, spp = 200, 150, 32
    # the one-GPU film
    scene = tp.Scene(device=0)
    cam = tp.Camera(45.0, w, h)
    scenes.load_scene(19, scene, cam)
    scene.build(cam)
    ref = tp.RendererImage(w, h, tp.SrgbRendererMis(tp.RendererArgs((w, h), spp, scene, cam))).render("sobol")
    # the group
    devs = (C.c_int32 * 2)(0, 1)
    g = C.c_void_p()
    assert lib.tcpt_group_create(devs, 2, C.byref(g)) == 0, lib.tcpt_group_last_error(g)
    try:
        assert lib.tcpt_group_size(g) == 2
        std, tab = tables
        assert lib.tcpt_group_set_tables(g, std, len(std), capi.as_ptr(tab, C.c_float), tab.size) == 0
        ctx0 = capi.Context.__new__(capi.Context)          # a borrowed handle: the group owns the context
        ctx0.lib, ctx0.handle, ctx0.has_gpu, ctx0.comm_rank, ctx0.comm_size = lib, C.c_void_p(lib.tcpt_group_context(g, 0)), True, 0, 2
        gs = tp.Scene(context=ctx0)
        scenes.load_scene(19, gs, tp.Camera(45.0, w, h))
        lib.tcpt_scene_clear(ctx0.handle)
        gs.desc.replay(gs)
        pos = np.asarray(cam.position, dtype=np.float32)
        assert lib.tcpt_group_build(g, capi.as_ptr(pos, C.c_float)) == 0, lib.tcpt_group_last_error(g)
        p = tp.SrgbRendererMis(tp.RendererArgs((w, h), spp, gs, cam)).params("sobol")
        acc = np.zeros((h, w, 3), np.float32); srgb = np.zeros((h, w, 3), np.float32)
        for mode in ("tile", "spp"):
            assert lib.tcpt_group_render(g, C.byref(p), capi.SHARD_MODES[mode], capi.as_ptr(acc, C.c_float), capi.as_ptr(srgb, C.c_float)) == 0, lib.tcpt_group_last_error(g)
            if mode == "tile":
                assert np.array_equal(acc.view(np.uint32), ref.accumulators.view(np.uint32))
                assert np.array_equal(srgb.view(np.uint32), ref.pixels.view(np.uint32))
            else:
                assert np.abs(acc - ref.accumulators).max() <= 1e-5 * max(1.0, np.abs(ref.accumulators).max())
            s0, s1 = capi.Stats(), capi.Stats()
            lib.tcpt_get_stats(lib.tcpt_group_context(g, 0), C.byref(s0)); lib.tcpt_get_stats(lib.tcpt_group_context(g, 1), C.byref(s1))
            assert s0.paths + s1.paths == w * h * spp and s0.paths > 0 and s1.paths > 0
        ctx0.handle = C.c_void_p()     # do not destroy the borrowed context
    finally:
        lib.tcpt_group_destroy(g)
