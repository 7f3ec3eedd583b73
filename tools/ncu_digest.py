#!/usr/bin/env python3
"""Digest an ncu launch list (tools/prof.sh <tag> traffic) into profiles/ncu_traffic.json: measured DRAM bytes, device time and FP32
instruction counts of every kernel of ONE bench step, per kernel family, next to the ray / vertex / path counts of the same step
(from the twin run without ncu) and the hash of the kernel sources it was captured from.  bench.py scales these per-ray / per-vertex
figures to its own run and refuses them when the sources have changed.
usage: ncu_digest.py launches.csv plain.json tag [workload]"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def family(name):
    for key in ("k_trace_fused", "k_trace_closest", "k_trace_shadow", "k_shade_all", "k_shade", "k_generate", "k_sobol_prefix", "k_sobol_pass", "k_film", "k_finalize", "k_aov"):
        if key in name:
            return "k_shade" if key == "k_shade_all" else key
    return None


def soup(csv_path, plain_path, tag, workload):
    """Launches of `bench.py --workload soup_* --steps 1 --warmup 0`: k_trace_rays #0 = the primaries that produce the hit points,
    #1 = the timed incoherent closest-hit batch, #2 = coherent, #3 = any hit (then the counting launches)."""
    rows = [r for r in csv.reader(open(csv_path, newline="")) if len(r) >= 15 and r[0].isdigit() and "k_trace_rays" in r[4]]
    launches = collections.OrderedDict()
    for r in rows:
        launches.setdefault(int(r[0]), {})[r[12]] = float(r[14].replace(",", ""))
    L = [launches[k] for k in sorted(launches)]
    plain = json.loads([ln for ln in open(plain_path) if ln.startswith("{")][-1])
    names = {1: "incoherent_closest", 2: "coherent_closest", 3: "incoherent_anyhit"}
    out = {"captured": tag, "kernel_source_sha": bench.kernel_source_sha()}
    for i, name in names.items():
        rays = plain[name]["rays"]
        db = L[i].get("dram__bytes_read.sum", 0.0) + L[i].get("dram__bytes_write.sum", 0.0)
        out[name] = {"dram_bytes_per_ray": db / rays, "ncu_ms": L[i].get("gpu__time_duration.sum", 0.0) / 1e6, "rays": rays,
                     "algorithmic_bytes_per_ray": plain[name]["bytes_per_ray"]}
    out["incoherent_dram_bytes_per_ray"] = out["incoherent_closest"]["dram_bytes_per_ray"]
    dst = ROOT / "profiles" / "ncu_traffic.json"
    doc = json.loads(dst.read_text())
    doc.setdefault("soups", {})[workload] = out
    dst.write_text(json.dumps(doc, indent=1))
    print(json.dumps(out, indent=1))


def main():
    if sys.argv[1] == "--soup":
        return soup(*sys.argv[2:6])
    csv_path, plain_path, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    workload = sys.argv[4] if len(sys.argv) > 4 else "scene19_4k"
    rows = [r for r in csv.reader(open(csv_path, newline="")) if len(r) >= 15 and r[0].isdigit()]
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r[0]), {"name": r[4]})
        try:
            d[r[12]] = float(r[14].replace(",", ""))
        except ValueError:
            pass
    # ONE step: from the first k_sobol_pass / k_generate up to and including the first k_film
    ids = sorted(launches)
    first = next(i for i in ids if "k_generate" in launches[i]["name"] or "k_sobol_pass" in launches[i]["name"])
    last = next(i for i in ids if i > first and "k_film" in launches[i]["name"])
    step = [launches[i] for i in ids if first <= i <= last and family(launches[i]["name"])]
    kern = collections.OrderedDict()
    per_launch = []
    for l in step:
        f = family(l["name"])
        k = kern.setdefault(f, {"launches": 0, "dram_bytes": 0.0, "time_ms": 0.0, "fadd": 0.0, "fmul": 0.0, "ffma": 0.0, "inst_thread": 0.0, "inst_warp": 0.0})
        k["launches"] += 1
        db = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        k["dram_bytes"] += db
        k["time_ms"] += l.get("gpu__time_duration.sum", 0.0) / 1e6
        k["fadd"] += l.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 0.0)
        k["fmul"] += l.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 0.0)
        k["ffma"] += l.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 0.0)
        k["inst_thread"] += l.get("smsp__thread_inst_executed.sum", 0.0)
        k["inst_warp"] += l.get("smsp__inst_executed.sum", 0.0)
        per_launch.append({"kernel": f, "ms": l.get("gpu__time_duration.sum", 0.0) / 1e6, "dram_mb": db / 1e6,
                           "lanes": (l.get("smsp__thread_inst_executed.sum", 0.0) / l["smsp__inst_executed.sum"]) if l.get("smsp__inst_executed.sum") else None})
    plain = json.loads([ln for ln in open(plain_path) if ln.startswith("{")][-1])
    c = plain["counts_rank0"]
    steps = plain["steps"]
    out = {"_comment": "ncu per-launch metrics (dram__bytes_read/write.sum, gpu__time_duration.sum, smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on.sum, "
                       "smsp__{thread_,}inst_executed.sum) summed per kernel family over the launches of ONE step (first k_generate .. first k_film) of "
                       "`python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras`; counts from the same command without ncu. Cold-cache, serialised replays: "
                       "use the per-ray figures and the kernel SHARES, not the absolute times.",
           "captured": tag, "workload": workload, "kernel_source_sha": bench.kernel_source_sha(),
           "paths": c["paths"] // steps, "closest_rays": c["closest_rays"] // steps, "shadow_rays": c["shadow_rays"] // steps,
           "total_dram_bytes": sum(k["dram_bytes"] for k in kern.values()), "total_time_ms": sum(k["time_ms"] for k in kern.values()),
           "kernels": kern}
    rays = out["closest_rays"] + out["shadow_rays"]
    t = out["total_time_ms"]
    out["summary"] = {f: {"share_of_step": k["time_ms"] / t, "dram_bytes_per_ray" if "trace" in f else "dram_bytes_per_vertex" if f == "k_shade" else "dram_bytes_per_path":
                          k["dram_bytes"] / (rays if "trace" in f else out["closest_rays"] if f == "k_shade" else out["paths"]),
                          "lanes_per_warp_instruction": (k["inst_thread"] / k["inst_warp"]) if k["inst_warp"] else None,
                          "flops_fadd_fmul_2ffma": k["fadd"] + k["fmul"] + 2 * k["ffma"]} for f, k in kern.items()}
    dst = ROOT / "profiles" / "ncu_traffic.json"
    old = {}
    try:
        old = json.loads(dst.read_text())
    except Exception:
        pass
    if "soups" in old:
        out["soups"] = old["soups"]
    dst.write_text(json.dumps(out, indent=1))
    (ROOT / "profiles" / f"{tag}_step_launches.json").write_text(json.dumps(per_launch, indent=0))
    print(json.dumps(out["summary"], indent=1))


if __name__ == "__main__":
    main()
