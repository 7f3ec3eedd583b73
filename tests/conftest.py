import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """Build libtcpt.so / liboracle.so if they are not there yet (they travel to the GPU box prebuilt)."""
    from toy_cpu_pathtracing_b200 import capi
    from oracle import oracle
    if not capi.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    oracle.build_library()
    return True


@pytest.fixture(scope="session")
def tables(built):
    from toy_cpu_pathtracing_b200 import capi
    return capi.load_tables()


class SceneBundle:
    """A config scene built once for both backends: .scene (libtcpt), .oracle (CPU restatement), .camera"""

    def __init__(self, scene_id, width, height, tables, require_gpu, **kw):
        import toy_cpu_pathtracing_b200 as tp
        from toy_cpu_pathtracing_b200 import scenes
        from oracle import oracle
        self.tp = tp
        self.width, self.height = width, height
        self.scene = tp.Scene(device=0, require_gpu=require_gpu)
        self.camera = tp.Camera(45.0, width, height)
        if callable(scene_id):            # a test's own scene: loader(scene, camera, **kw)
            scene_id(self.scene, self.camera, **kw)
        elif scene_id == "soup":
            scenes.load_soup(self.scene, self.camera, **kw)
        else:
            scenes.load_scene(scene_id, self.scene, self.camera, **kw)
        self.scene.build(self.camera)
        self.oracle = oracle.scene_from_description(self.scene.desc, self.camera.position, tables[0], tables[1])

    def image(self, integrator, spp, seed=0, max_depth=16):
        tp = self.tp
        r = tp.RENDERERS[integrator](tp.RendererArgs((self.width, self.height), spp, self.scene, self.camera, seed=seed), max_depth=max_depth)
        return tp.RendererImage(self.width, self.height, r)

    def oparams(self, integrator, sampler, spp, seed=0, max_depth=16, **kw):
        return self.oracle.params(self.width, self.height, spp, integrator, sampler, self.camera, seed=seed, max_depth=max_depth, **kw)


@pytest.fixture(scope="session")
def bundle_factory(tables):
    cache = {}

    def make(scene_id, width, height, require_gpu=True, **kw):
        key = (scene_id, width, height, require_gpu, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = SceneBundle(scene_id, width, height, tables, require_gpu, **kw)
        return cache[key]

    return make
