"""Developer probe (not a test): GPU vs oracle on sampler streams, recorded rays and per-path samples.  Run under gpurun."""
import sys, time, traceback
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import capi, scenes
from oracle import oracle

std, tab = capi.load_tables()


def setup(scene_id, w, h, **kw):
    sc = tp.Scene(device=0)
    cam = tp.Camera(45.0, w, h)
    scenes.load_scene(scene_id, sc, cam, **kw)
    t = time.time(); sc.build(cam); tb = time.time() - t
    osc = oracle.scene_from_description(sc.desc, cam.position, std, tab)
    return sc, cam, osc, tb


def main():
    which = [int(a) for a in sys.argv[1:]] or [3, 10, 17, 19]
    for sid in which:
        w, h, spp = 64, 48, 16
        try:
            sc, cam, osc, tb = setup(sid, w, h)
            print(f"=== scene {sid}: build {tb:.2f}s, depth {sc.ctx.stats()['max_bvh_depth']}")
            # recorded rays
            p = osc.params(w, h, 2, "mis", "sobol", cam, window=(8, 8, 40, 40))
            closest, shadow = osc.record_rays(p)
            rays = np.concatenate([closest, np.full((len(closest), 1), np.finfo(np.float32).max, np.float32)], 1)
            o_hit, nb, nt = osc.trace(rays)
            g_hit = sc.trace(rays)
            bad = np.any(o_hit != g_hit, axis=1)
            print(f"closest rays {len(rays)}: mismatches {bad.sum()}  (oracle box {nb} tri {nt})")
            if bad.any():
                i = np.nonzero(bad)[0][:5]
                print(" oracle", o_hit[i], "\n gpu", g_hit[i], "\n rays", rays[i])
            o_any, _, _ = osc.trace(shadow, any_hit=True)
            g_any = sc.trace(shadow, any_hit=True)
            print(f"shadow rays {len(shadow)}: mismatches {(o_any[:, 0] != g_any[:, 0]).sum()} occluded {o_any[:,0].sum()}")
            for integ in ("pt", "nee", "mis"):
                for smp in ("sobol", "random"):
                    rnd = tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=0))
                    img = tp.RendererImage(w, h, rnd)
                    rng = np.random.default_rng(1)
                    n = 4000
                    xy = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.uint32)
                    si = rng.integers(0, spp, n).astype(np.uint32)
                    g = img.path_samples(smp, xy, si)
                    o = osc.path_samples(osc.params(w, h, spp, integ, smp, cam), xy, si)
                    err = np.abs(g - o).max(1); mag = np.abs(o).max(1) + 1e-6
                    rel = err / mag
                    t = time.time(); img.render(smp); tg = time.time() - t
                    acc, _, st = osc.render(osc.params(w, h, spp, integ, smp, cam))
                    ga, oa = img.accumulators / spp, acc / spp
                    mre = np.abs(ga - oa).mean() / np.abs(oa).mean()
                    s = img.stats
                    print(f"{integ:3s} {smp:6s}: paths rel>1e-4: {(rel > 1e-4).sum()}/{n} exact {(err == 0).sum()} nan g/o {np.isnan(g).any(1).sum()}/{np.isnan(o).any(1).sum()} | image MRE {mre:.2e} "
                          f"| gpu {s['render_ms']:.1f} ms {s['closest_rays'] + s['shadow_rays']} rays (oracle {st['closest_rays'] + st['shadow_rays']}) wall {tg:.2f}s oracle {st['seconds']:.2f}s")
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
