"""Worker of tests/test_gpu_multi.py::test_two_ranks_reproduce_the_one_gpu_film, launched by torchrun with one rank per GPU.
Every rank: scene 19 at 200x150, its own libtcpt context on its GPU, tcpt_comm_init (the NCCL unique id travels through a gloo
process group), then tcpt_render_sharded in tile and in spp mode.  Rank 0 also renders the frame alone and compares."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    import toy_cpu_pathtracing_b200 as tp
    from toy_cpu_pathtracing_b200 import scenes
    from toy_cpu_pathtracing_b200.multi_gpu import init_comm

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # plumbing only: carries the NCCL unique id; the film reduce is libtcpt's own ncclReduce
    w, h, spp = 200, 150, 64
    scene = tp.Scene(device=local)
    cam = tp.Camera(45.0, w, h)
    scenes.load_scene(19, scene, cam)
    scene.build(cam)
    init_comm(scene.ctx, rank, world)
    mk = lambda: tp.RendererImage(w, h, tp.SrgbRendererMis(tp.RendererArgs((w, h), spp, scene, cam)))  # noqa: E731
    results = {}
    for mode in ("tile", "spp"):
        img = mk().render_sharded("sobol", mode=mode)
        results[mode] = (img.accumulators.copy(), img.pixels.copy(), dict(img.stats))
        paths = torch.tensor([img.stats["paths"]], dtype=torch.int64)
        dist.all_reduce(paths)
        assert int(paths) == w * h * spp, (mode, int(paths))
        if world > 1:
            assert img.stats["reduce_ms"] > 0.0
    # a block of sample indices only (what a progressive host does)
    part = mk().render_sharded("sobol", mode="spp", spp_window=(16, 48))
    dist.barrier()
    scene.ctx.comm_destroy()
    if rank == 0:
        alone = mk().render("sobol")
        acc, pix, _ = results["tile"]
        assert np.array_equal(acc.view(np.uint32), alone.accumulators.view(np.uint32)), "tile mode is not bitwise the one-GPU film"
        assert np.array_equal(pix.view(np.uint32), alone.pixels.view(np.uint32))
        print("MGPU_OK tile", flush=True)
        acc, pix, _ = results["spp"]
        err = np.abs(acc - alone.accumulators).max() / max(1.0, np.abs(alone.accumulators).max())
        assert err <= 1e-5, err
        assert np.abs(pix - alone.pixels).max() <= 1e-5
        ref = mk().render("sobol", spp_begin=16, spp_end=48)
        assert np.abs(part.accumulators - ref.accumulators).max() <= 1e-5 * max(1.0, np.abs(ref.accumulators).max())
        print(f"MGPU_OK spp max_rel_err {err:.2e}", flush=True)
    else:
        assert not results["tile"][0].any()     # host buffers are only written on rank 0
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
