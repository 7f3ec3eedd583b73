#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_*.npz: small frames and hit records produced by the ORACLE (the Rust reference cannot run
here, so these are regression pins of the restatement, not reference outputs; the Sobol vectors in sobol_kat.json are the
only fixtures derived independently of the oracle).  Usage: python tools/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import toy_cpu_pathtracing_b200 as tp  # noqa: E402
from toy_cpu_pathtracing_b200 import capi, scenes  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = [(3, {}, "mis", "sobol"), (10, {}, "nee", "sobol"), (17, {}, "mis", "sobol"), (17, {"coat": False}, "pt", "random"), (19, {}, "mis", "sobol"),
         (7, {}, "mis", "sobol"), (8, {}, "mis", "sobol"), (1, {}, "mis", "sobol"), (2, {}, "nee", "sobol"), ("lights", {"directional": True}, "mis", "sobol")]
W, H, SPP = 24, 18, 8


def main():
    std, tab = capi.load_tables()
    out = ROOT / "tests" / "golden"
    for sid, kw, integ, smp in CASES:
        scene = tp.Scene(require_gpu=False)
        cam = tp.Camera(45.0, W, H)
        scenes.load_scene(sid, scene, cam, **kw)
        osc = oracle.scene_from_description(scene.desc, cam.position, std, tab)
        acc, srgb, st = osc.render(osc.params(W, H, SPP, integ, smp, cam, threads=1))
        closest, shadow = osc.record_rays(osc.params(W, H, 1, integ, smp, cam, window=(8, 6, 16, 12)))
        rays = np.concatenate([closest, np.full((len(closest), 1), np.finfo(np.float32).max, np.float32)], 1)[:400]
        hits, _, _ = osc.trace(rays)
        name = f"oracle_scene{sid}{''.join('_' + ('nocoat' if k == 'coat' else k) for k in kw)}_{integ}_{smp}.npz"
        np.savez_compressed(out / name, acc=acc, rays=rays, hits=hits, counts=np.array([st["closest_rays"], st["shadow_rays"]], dtype=np.int64))
        print(name, acc.mean(), st["closest_rays"], st["shadow_rays"])


if __name__ == "__main__":
    main()
