"""Developer probe: where does non-kernel time of a bench step go?  Per step: wall, render_ms (events around the pass), sum of stage times."""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import scenes
W, H, S = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 8
timing = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = tp.Scene(device=0); cam = tp.Camera(45.0, W, H)
scenes.load_scene(19, sc, cam); sc.build(cam)
ctx = sc.ctx
r = tp.RENDERERS["mis"](tp.RendererArgs((W, H), 4096, sc, cam, seed=0))
acc = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
stream = torch.cuda.Stream(device=0); torch.cuda.set_stream(stream)
ctx.set_option("stage_timing", timing)
for k in range(10):
    p = r.params("sobol", spp_begin=k * S, spp_end=(k + 1) * S)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.check(ctx.lib.tcpt_render_device(ctx.handle, C.byref(p), C.c_void_p(acc.data_ptr()), C.c_void_p(stream.cuda_stream)))
    t1 = time.perf_counter()
    st = ctx.stats()
    ssum = st["trace_closest_ms"] + st["trace_shadow_ms"] + st["shade_ms"] + st["generate_ms"] + st["film_ms"]
    print(f"step {k}: wall {1e3 * (t1 - t0):7.2f} ms render_ms {st['render_ms']:7.2f} stage sum {ssum:7.2f} prefix {st['sobol_prefix_ms']:.2f}", flush=True)
