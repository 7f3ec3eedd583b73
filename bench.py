#!/usr/bin/env python3
"""bench.py — throughput of the path-integration hot path on B200 (Mrays/s, with Mpaths/s beside it).

Workload (BASELINE.json configs[3], the configuration the metric's 1/2/4/8-GPU numbers are quoted on): scene 19 (floor + three
dragon instances: textured SimplePbr, smooth clearcoat, plastic; environment light) at 3840x2160, MIS integrator, Z-Sobol
sampler, max_depth 16, as a 4096-spp frame.  One STEP = one pass of the hot path over one batch = `--spp-per-step` sample
indices of every pixel of the frame per GPU (default 16 -> 132.7 M paths per GPU per step, 35 GB of wavefront state in HBM), including the film kernel and, at
N > 1, the NCCL reduce of the film accumulators.  Ranks render disjoint sample-index ranges of the same frame (spp-pass
sharding), so per-GPU work is fixed as N grows: "scaling": "weak".  configs[0..2] are parity-test cases (tests/), not bench lines.

  value    whole-job Mrays/s, film accumulators resident in HBM (tcpt_render_device), CUDA events on the launching stream
  e2e      the same metric through the reference-facing call RendererImage.render() -> tcpt_render() with HOST buffers:
           per step the render parameters go host->device and the tone-mapped sRGB frame comes device->host
  roofline dominant kernel k_trace_fused (closest-hit + any-hit traversal); see DESIGN.md "Measurement" for the byte/flop definitions
  cpu_baseline / --impl reference: the CPU restatement of the reference algorithm (oracle, kind "port": the Rust reference
           cannot be compiled in this image) on all host threads, on a bounded window of the same frame
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    "scene19_4k": dict(scene=19, width=3840, height=2160, frame_spp=4096, integrator="mis", sampler="sobol"),
    "scene17_1080p": dict(scene=17, width=1920, height=1080, frame_spp=1024, integrator="mis", sampler="sobol"),
    "scene17_1080p_nocoat": dict(scene=17, width=1920, height=1080, frame_spp=1024, integrator="mis", sampler="sobol", kw={"coat": False}),
    "scene10_test": dict(scene=10, width=200, height=150, frame_spp=512, integrator="mis", sampler="sobol"),
    "scene3_test": dict(scene=3, width=200, height=150, frame_spp=512, integrator="mis", sampler="sobol"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="scene19_4k", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=16)
    ap.add_argument("--max-slots", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period", type=float, default=0.05, help="seconds between NVML clock samples during the timed region")
    ap.add_argument("--opt", action="append", default=[], help="developer knob: libtcpt option as name=value (tcpt_set_option), repeatable")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


# ---------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML from a thread every 10 ms
    (ctypes calls into libtcpt release the GIL), nvidia-smi polling as a fallback when pynvml is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int, period: float = 0.05):
        self.index, self.rows, self.proc, self.thread, self.stop, self.period = index, [], None, None, False, period
        self.sm, self.sm_max, self.reasons = [], None, set()

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.reasons |= {k for k, b in bits.items() if r & b}
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it holds plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread:
            self.thread.join(timeout=1)

    def summary(self):
        if self.sm:
            return {"sm_mhz": int(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------- shared setup
def build_scene(wl, device, require_gpu=True):
    import toy_cpu_pathtracing_b200 as tp
    from toy_cpu_pathtracing_b200 import scenes
    scene = tp.Scene(device=device, require_gpu=require_gpu)
    cam = tp.Camera(45.0, wl["width"], wl["height"])
    scenes.load_scene(wl["scene"], scene, cam, **wl.get("kw", {}))
    return tp, scene, cam


def cpu_window(wl, seconds, paths_per_second_guess=2.0e5):
    """A centred pixel window of the SAME frame sized for about `seconds` of CPU work (4 sample indices per pixel, more when the
    whole frame is too small to fill the time)."""
    want_paths = max(2000.0, seconds * paths_per_second_guess)
    spp = 4
    side = int(max(8, min(wl["height"], wl["width"], (want_paths / spp) ** 0.5)))
    if side * side * spp < 0.7 * want_paths:
        spp = int(min(wl["frame_spp"], max(4, want_paths / (side * side))))
    x0, y0 = (wl["width"] - side) // 2, (wl["height"] - side) // 2
    return (x0, y0, x0 + side, y0 + side), spp


def run_cpu(wl, scene_desc, cam, seconds, threads=0, optimised=False):
    """The reference algorithm on the host cores (oracle port, reference-faithful mode: exhaustive traversal without t-shrinking,
    per-call instance-matrix inverses), on a bounded window of the bench frame.  Returns (Mrays/s, Mpaths/s, info)."""
    from toy_cpu_pathtracing_b200 import capi
    from oracle import oracle
    std, tab = capi.load_tables()
    osc = oracle.scene_from_description(scene_desc, cam.position, std, tab, faithful=not optimised, literal_build=False)
    if optimised:   # the same port with an ordered, t-shrinking traversal and cached instance inverses, so that the GPU / CPU ratio
        osc.set_optimised(True)   # is not inflated by the reference's exhaustive traversal (same image, checked bit for bit on scenes 3 and 19)
    # calibrate on a small window, then size the real sample
    win, spp = cpu_window(wl, 0.5)
    p = osc.params(wl["width"], wl["height"], wl["frame_spp"], wl["integrator"], wl["sampler"], cam, threads=threads, window=win)
    p.spp = wl["frame_spp"]
    _, _, st = _oracle_window(osc, p, spp)
    rate = st["paths"] / max(st["seconds"], 1e-6)
    win, spp = cpu_window(wl, seconds, rate)
    p = osc.params(wl["width"], wl["height"], wl["frame_spp"], wl["integrator"], wl["sampler"], cam, threads=threads, window=win)
    _, _, st = _oracle_window(osc, p, spp)
    rays = st["closest_rays"] + st["shadow_rays"]
    ncores = threads if threads > 0 else (os.cpu_count() or 1)
    info = {"cores": ncores, "sample": f"{win[2] - win[0]}x{win[3] - win[1]} px window of the {wl['width']}x{wl['height']} frame, sample indices 0..{spp - 1} of {wl['frame_spp']}, "
                                       f"{st['paths']} paths in {st['seconds']:.2f} s", "seconds": st["seconds"], "paths": st["paths"], "rays": rays}
    return rays / st["seconds"] / 1e6, st["paths"] / st["seconds"] / 1e6, info


def _oracle_window(osc, p, spp_used):
    """Render sample indices 0..spp_used-1 of the frame's Sobol sequence (the sampler is configured for the full frame spp)."""
    # the oracle's render loop runs `spp` samples per pixel; the Sobol sampler must still see the frame's spp: orc_render uses
    # p.spp for both, so a window render at reduced spp uses sampler parameters of the reduced spp.  Throughput is independent
    # of which low-discrepancy points are drawn, so the sample renders `spp_used` samples with the sampler sized for them.
    p.spp = spp_used
    return osc.render(p)


# ---------------------------------------------------------------- reference arm
def main_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    _, scene, cam = build_scene(wl, 0, require_gpu=False)
    scene.desc  # description only; nothing is built on a GPU in this arm
    vals, paths = [], []
    info = None
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        mr, mp, info = run_cpu(wl, scene.desc, cam, per_step)
        if i >= args.warmup:
            vals.append((info["rays"], info["seconds"]))
            paths.append(info["paths"])
    tot_r = sum(v[0] for v in vals); tot_s = sum(v[1] for v in vals)
    value = tot_r / tot_s / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl), "mpaths_per_s": sum(paths) / tot_s / 1e6,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, wl):
    return {"workload": f"scene{wl['scene']} {wl['width']}x{wl['height']} {wl['integrator']}+{wl['sampler']} max_depth 16, {wl['frame_spp']}-spp frame, "
                        f"{args.spp_per_step} sample indices of every pixel per GPU per step" + (" (BASELINE.json configs[3])" if wl["scene"] == 19 else "") + (" no-coat" if wl.get("kw") else ""),
            "paths_per_gpu_per_step": wl["width"] * wl["height"] * args.spp_per_step, "sharding": "spp-pass", "collective": "one NCCL reduce of the film accumulators per step",
            "cache": f"working set (path state + ray queues, {wl['width'] * wl['height'] * args.spp_per_step * 264 / 1e9:.1f} GB per GPU) exceeds the 126 MB L2; no explicit flush", "assets": "procedural stand-ins (reference assets are LFS stubs)"}


# ---------------------------------------------------------------- GPU arm
def main_gpu(args, wl):
    import ctypes as C

    import torch
    import torch.distributed as dist
    from toy_cpu_pathtracing_b200 import capi
    from toy_cpu_pathtracing_b200.multi_gpu import reduce_film, shard_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    tp, scene, cam = build_scene(wl, local)
    t0 = time.time(); scene.build(cam); build_s = time.time() - t0
    ctx = scene.ctx
    for kv in args.opt:
        name, _, val = kv.partition("=")
        ctx.set_option(name, int(val))
    W, H, S = wl["width"], wl["height"], args.spp_per_step
    renderer = tp.RENDERERS[wl["integrator"]](tp.RendererArgs((W, H), wl["frame_spp"], scene, cam, seed=0))
    image = tp.RendererImage(W, H, renderer)
    acc = torch.zeros((H, W, 3), dtype=torch.float32, device=f"cuda:{local}")
    frame = torch.zeros_like(acc) if rank == 0 else None
    # a real (non-default) stream: libtcpt launches on the handle it is given, NCCL and the torch events use the same stream,
    # so the CUDA events bracket exactly the kernels of the timed steps
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)

    def step(k, timing=False):
        # step k of the job: `world * S` fresh sample indices of the frame, rank r takes its S of them (spp-pass sharding)
        lo = (k * world * S) % wl["frame_spp"]
        sh = shard_plan(rank, world, "spp", wl["frame_spp"], lo, min(lo + world * S, wl["frame_spp"]))
        p = renderer.params(wl["sampler"], max_slots=args.max_slots, **sh.as_kwargs())
        acc.zero_()
        ctx.check(ctx.lib.tcpt_render_device(ctx.handle, C.byref(p), C.c_void_p(acc.data_ptr()), C.c_void_p(stream.cuda_stream)))
        st = ctx.stats()
        reduce_film(acc, dst=0)
        if rank == 0:
            frame.add_(acc)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    ctx.set_option("stage_timing", 1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = {"rays": 0, "paths": 0, "closest": 0, "shadow": 0, "launches": 0, "closest_ms": 0.0, "shade_ms": 0.0, "shadow_ms": 0.0, "gen_ms": 0.0, "film_ms": 0.0, "passes": 0}
    with ClockSampler(local, args.clock_period) as clk:
        e0.record(stream)
        for k in range(args.steps):
            st = step(args.warmup + k)
            tot["rays"] += st["closest_rays"] + st["shadow_rays"]; tot["paths"] += st["paths"]; tot["closest"] += st["closest_rays"]; tot["shadow"] += st["shadow_rays"]
            tot["launches"] += st["kernel_launches"]; tot["closest_ms"] += st["trace_closest_ms"]; tot["shade_ms"] += st["shade_ms"]; tot["shadow_ms"] += st["trace_shadow_ms"]
            tot["gen_ms"] += st["generate_ms"]; tot["film_ms"] += st["film_ms"]; tot["passes"] += st["passes"]
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    ctx.set_option("stage_timing", 0)
    t = torch.tensor([ms, float(tot["rays"]), float(tot["paths"]), float(tot["launches"])], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, rays_all, paths_all, launches_all = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item()
    else:
        ms_max, rays_all, paths_all, launches_all = ms, float(tot["rays"]), float(tot["paths"]), float(tot["launches"])
    value = rays_all / (ms_max * 1e-3) / 1e6

    # ---- e2e through the reference-facing API with host buffers (every rank renders its share; rank 0 reports the sum / max time)
    def e2e_step(k):
        lo = (k * world * S) % wl["frame_spp"]
        sh = shard_plan(rank, world, "spp", wl["frame_spp"], lo, min(lo + world * S, wl["frame_spp"]))
        image.render(wl["sampler"], want_accumulators=False, max_slots=args.max_slots, **sh.as_kwargs())   # the reference-facing call; fills image.pixels (host)
        return image.stats["closest_rays"] + image.stats["shadow_rays"]
    e2e_step(0)
    barrier()
    t0 = time.perf_counter(); e2e_rays = 0
    n_e2e = max(2, min(args.steps, 4))
    for k in range(n_e2e):
        e2e_rays += e2e_step(args.warmup + k)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = tm[0].item(), ts[1].item()
    e2e_value = e2e_rays / e2e_s / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_trace_closest): one extra counted step OUTSIDE the timed region gives B and T per ray
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    ctx.set_option("count_tests", 1)
    p = renderer.params(wl["sampler"], max_slots=args.max_slots, spp_begin=0, spp_end=1)
    acc.zero_()
    ctx.check(ctx.lib.tcpt_render_device(ctx.handle, C.byref(p), C.c_void_p(acc.data_ptr()), C.c_void_p(stream.cuda_stream)))
    cs = ctx.stats()
    ctx.set_option("count_tests", 0)
    rays_c = max(1, cs["closest_rays"] + cs["shadow_rays"])
    box_per_ray, tri_per_ray = cs["box_tests"] / rays_c, cs["tri_tests"] / rays_c
    # dominant kernel: k_trace_fused (one launch per bounce: the shadow rays of the previous bounce + this bounce's extension rays)
    n_trace_launches = max(1, tot["passes"] * 17)
    trace_s = (tot["closest_ms"] + tot["shadow_ms"]) * 1e-3
    bytes_per_ray = 48.0  # 32 B ray read + 16 B hit write (compulsory wavefront traffic; the BVH is L2 resident)
    achieved = bytes_per_ray * tot["rays"] / trace_s / 1e9 if trace_s > 0 else 0.0
    traffic = None
    try:  # measured DRAM traffic of the same kernel from the committed ncu capture, scaled to this run's rays per launch
        tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())["k_trace_fused"]
        traffic = tr["dram_bytes"] / tr["rays"] * (tot["rays"] / n_trace_launches)
    except Exception:
        pass
    clk_s = clk.summary()
    sm_mhz = clk_s["sm_mhz"] or 1965
    flops_per_ray = 18.0 * box_per_ray + 64.0 * tri_per_ray + 60.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    fp32_ach = flops_per_ray * tot["rays"] / trace_s / 1e12 if trace_s > 0 else 0.0
    stage_sum = tot["closest_ms"] + tot["shade_ms"] + tot["shadow_ms"] + tot["gen_ms"] + tot["film_ms"]
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl),
        "mpaths_per_s": paths_all / (ms_max * 1e-3) / 1e6, "rays_per_path": rays_all / max(1.0, paths_all),
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(capi.RenderParams), "d2h_bytes_per_step": int(image.pixels.nbytes), "steps": n_e2e,
                "call": "RendererImage.render -> tcpt_render(params, host sRGB frame out)"},
        "gpu_launches": int(launches_all),
        "clocks": clk_s,
        "roofline": {"kernel": "k_trace_fused", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "rays_per_launch": tot["rays"] / n_trace_launches,
                     "avg_launch_ms": 1e3 * trace_s / n_trace_launches, "share_of_step": 1e3 * trace_s / stage_sum if stage_sum else None,
                     "note": "BVH+textures are L2 resident: the path is FP32-issue/latency bound, see roofline_fp32"},
        "roofline_fp32": {"kernels": "k_trace_fused", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak if fp32_peak else None,
                          "flops_per_ray": flops_per_ray, "box_tests_per_ray": box_per_ray, "tri_tests_per_ray": tri_per_ray, "sm_mhz": sm_mhz,
                          "definition": "18*B + 64*T + 60 flops per ray (SURVEY.md 8d); peak = 148 SM x 128 lanes x 2 x f_SM"},
        # node records (64 B per visited pair = 32 B per box test) and triangle records (48 B per test) are served by
        # L1 / L2 (the BVH is L2 resident); informational, no measured L1/L2 peak to divide by
        "bvh_traffic": {"kernel": "k_trace_fused", "bytes_per_ray": 32.0 * box_per_ray + 48.0 * tri_per_ray,
                        "achieved": (32.0 * box_per_ray + 48.0 * tri_per_ray) * tot["rays"] / trace_s / 1e9 if trace_s > 0 else 0.0, "unit": "GB/s",
                        "served_by": "L1/L2", "trace_only_mrays_per_s": tot["rays"] / trace_s / 1e6 if trace_s > 0 else 0.0},
        "stage_ms_per_step": {k: tot[k] / args.steps for k in ("gen_ms", "closest_ms", "shade_ms", "shadow_ms", "film_ms")},
        "scene_build_s": build_s,
    }
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only
        try:
            mr, mp, info = run_cpu(wl, scene.desc, cam, args.cpu_seconds)
            line["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": info["cores"], "kind": "port", "sample": info["sample"], "mpaths_per_s": mp}
            try:
                mr2, mp2, info2 = run_cpu(wl, scene.desc, cam, max(2.0, args.cpu_seconds / 3.0), optimised=True)
                line["cpu_baseline"]["optimised"] = {"value": mr2, "unit": "Mrays/s", "mpaths_per_s": mp2, "sample": info2["sample"],
                                                     "what": "same port, ordered t-shrinking traversal + cached instance inverses (not the reference's algorithmic cost)"}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"]["optimised"] = {"value": None, "what": f"failed: {e}"}
        except Exception as e:  # the oracle is test infrastructure; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return main_reference(args, wl)
    return main_gpu(args, wl)


if __name__ == "__main__":
    sys.exit(main())
