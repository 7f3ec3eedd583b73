"""The N > 1 path on CPU: two gloo ranks shard a frame (rows or sample ranges), render their shard (here with the ORACLE as
the per-rank renderer, which tests may use), reduce the film with the product's reduce_film(), and must reproduce the
single-process frame: bitwise for row shards, within rounding for sample-range shards."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_plan_partitions_everything():
    from toy_cpu_pathtracing_b200.multi_gpu import shard_plan
    for world in (1, 2, 3, 8):
        rows = set()
        for r in range(world):
            s = shard_plan(r, world, "tile", 512)
            rows |= set(range(s.row_offset, 150, s.row_stride))
            assert (s.spp_begin, s.spp_end) == (0, 512)
        assert rows == set(range(150))
        got = []
        for r in range(world):
            s = shard_plan(r, world, "spp", 512, 8, 8 + 5 * 3)
            got += list(range(s.spp_begin, s.spp_end))
        assert got == list(range(8, 23))
    with pytest.raises(ValueError):
        shard_plan(2, 2, "tile", 16)
    with pytest.raises(ValueError):
        shard_plan(0, 2, "cube", 16)


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import toy_cpu_pathtracing_b200 as tp
    from toy_cpu_pathtracing_b200 import capi, scenes
    from toy_cpu_pathtracing_b200.multi_gpu import reduce_film, shard_plan
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, spp = 32, 24, 8
    scene = tp.Scene(require_gpu=False)
    cam = tp.Camera(45.0, w, h)
    scenes.load_scene(10, scene, cam)
    std, tab = capi.load_tables()
    osc = oracle.scene_from_description(scene.desc, cam.position, std, tab)
    sh = shard_plan(rank, world, mode, spp)
    acc = np.zeros((h, w, 3), dtype=np.float32)
    if mode == "tile":
        full, _, _ = osc.render(osc.params(w, h, spp, "mis", "sobol", cam, threads=2))
        acc[sh.row_offset::sh.row_stride] = full[sh.row_offset::sh.row_stride]     # this rank's rows, all samples
    else:
        p = osc.params(w, h, spp, "mis", "sobol", cam, threads=2)
        xy = np.array([[x, y] for y in range(h) for x in range(w)], dtype=np.uint32)
        for s in range(sh.spp_begin, sh.spp_end):                                   # this rank's sample indices, all pixels
            acc += osc.path_samples(p, xy, np.full(len(xy), s, np.uint32)).reshape(h, w, 3)
    t = torch.from_numpy(acc)
    reduce_film(t, dst=0)
    if rank == 0:
        np.save(Path(out_dir) / f"{mode}.npy", t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tile", "spp"])
def test_two_rank_film_reduce_matches_single_process(built, tables, tmp_path, mode):
    import torch.multiprocessing as mp
    import toy_cpu_pathtracing_b200 as tp
    from toy_cpu_pathtracing_b200 import scenes
    from oracle import oracle
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "tile" else 1)
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / f"{mode}.npy")
    w, h, spp = 32, 24, 8
    scene = tp.Scene(require_gpu=False)
    cam = tp.Camera(45.0, w, h)
    scenes.load_scene(10, scene, cam)
    osc = oracle.scene_from_description(scene.desc, cam.position, tables[0], tables[1])
    ref, _, _ = osc.render(osc.params(w, h, spp, "mis", "sobol", cam, threads=2))
    if mode == "tile":
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    else:
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
