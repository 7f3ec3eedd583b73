"""ORACLE (test infrastructure, NOT product code): ctypes wrapper over oracle/liboracle.so, the CPU restatement of the
reference's path-integration hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  PARITY UNPINNED against the reference's own outputs below the Sobol known-answer vectors and the BVH
topology checks (the Rust reference cannot be built here and holds no fixtures below image level); pinned to restatement-independent
ground truth instead by tests/test_{intersection,analytic}_ground_truth.py and tests/test_rgb2spec.py."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
LIB = ORACLE_DIR / "liboracle.so"
f32 = np.float32


class OrcRenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("seed", C.c_uint32), ("max_depth", C.c_uint32),
                ("integrator", C.c_int32), ("sampler", C.c_int32), ("exposure", C.c_float), ("fov_deg", C.c_float),
                ("cam_pos", C.c_float * 3), ("cam_dir", C.c_float * 3), ("cam_up", C.c_float * 3), ("threads", C.c_int32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32), ("stride", C.c_uint32)]


class OrcStats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("paths", C.c_uint64), ("box_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("seconds", C.c_double)]


def build_library(force: bool = False) -> Path:
    if force or not LIB.exists():
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_library()
        _lib = C.CDLL(str(LIB))
        _lib.orc_scene_new.restype = C.c_void_p
        _lib.orc_build.restype = C.c_double
        _lib.orc_trace_mt.restype = C.c_double
        for name in ("orc_scene_free", "orc_set_tables", "orc_add_mesh", "orc_set_tangent_source", "orc_add_single_triangle", "orc_add_texture", "orc_add_material", "orc_rgb_to_coeffs", "orc_add_primitive",
                     "orc_add_env_light", "orc_add_delta_light", "orc_set_modes", "orc_set_optimised", "orc_build", "orc_render", "orc_path_samples", "orc_record_rays", "orc_trace", "orc_sobol_probe",
                     "orc_sampler_stream", "orc_get_bvh", "orc_get_mesh_tangents"):
            getattr(_lib, name).argtypes = None
    return _lib


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _vp(h):
    return C.c_void_p(h)


class OracleScene:
    """Backend of toy_cpu_pathtracing_b200.scene.SceneDescription.replay for the CPU oracle."""

    def __init__(self, std_tables: bytes, rgb2spec: np.ndarray, faithful: bool = False, literal_build: bool = False):
        self.l = lib()
        self.h = self.l.orc_scene_new()
        self._keep = [std_tables, np.ascontiguousarray(rgb2spec, dtype=f32)]
        rc = self.l.orc_set_tables(_vp(self.h), std_tables, C.c_size_t(len(std_tables)), _p(self._keep[1]), C.c_size_t(self._keep[1].size))
        if rc != 0:
            raise RuntimeError(f"orc_set_tables failed: {rc}")
        self.l.orc_set_modes(_vp(self.h), int(faithful), int(literal_build))

    def close(self):
        if self.h:
            self.l.orc_scene_free(_vp(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- replay protocol
    def add_mesh(self, pos, nrm, uv, idx):
        return self.l.orc_add_mesh(_vp(self.h), _p(pos), _p(nrm), _p(uv) if uv is not None else None, len(pos), _p(idx, C.c_uint32), len(idx))

    def set_tangent_source(self, geometry, tri):
        return self.l.orc_set_tangent_source(_vp(self.h), geometry, _p(tri, C.c_uint32), len(tri))

    def add_single_triangle(self, pos, nrm, uv):
        return self.l.orc_add_single_triangle(_vp(self.h), _p(pos), _p(nrm), _p(uv))

    def add_texture(self, arr):
        hgt, wid = arr.shape[:2]
        return self.l.orc_add_texture(_vp(self.h), _p(arr, C.c_uint8), C.c_uint32(wid), C.c_uint32(hgt), C.c_uint32(1 if arr.ndim == 2 else arr.shape[2]))

    def add_material(self, desc):
        return self.l.orc_add_material(_vp(self.h), C.byref(desc))  # orc_material_desc has the layout of tcpt_material_desc

    def add_primitive(self, geometry, material, l2w):
        return self.l.orc_add_primitive(_vp(self.h), geometry, material, _p(l2w))

    def add_env_light(self, intensity, rgb, l2w):
        hgt, wid = rgb.shape[:2]
        return self.l.orc_add_env_light(_vp(self.h), C.c_float(intensity), _p(rgb), C.c_uint32(wid), C.c_uint32(hgt), _p(l2w))

    def add_delta_light(self, kind, intensity, spectrum, angle_inner, angle_outer, l2w):
        return self.l.orc_add_delta_light(_vp(self.h), kind, C.c_float(intensity), C.byref(spectrum), C.c_float(angle_inner), C.c_float(angle_outer), _p(l2w))

    def build(self, cam_pos) -> float:
        a = np.asarray(cam_pos, dtype=f32)
        return self.l.orc_build(_vp(self.h), _p(a))

    def set_optimised(self, on: bool = True):
        """bench.py only: ordered, t-shrinking traversal + cached inverses (the "optimised CPU" figure of SURVEY 8d); call after build()."""
        self.l.orc_set_optimised(_vp(self.h), int(on))

    # --- queries
    @staticmethod
    def params(width, height, spp, integrator, sampler, camera, seed=0, max_depth=16, exposure=1.0, threads=0, window=(0, 0, 0, 0), stride=1) -> OrcRenderParams:
        p = OrcRenderParams()
        p.width, p.height, p.spp, p.seed, p.max_depth = width, height, spp, seed, max_depth
        p.integrator = {"pt": 0, "nee": 1, "mis": 2, "albedo": 3, "normal": 4}[integrator]
        p.sampler = {"random": 0, "sobol": 1}[sampler]
        p.exposure, p.fov_deg, p.threads = exposure, camera.fov, threads
        for k in range(3):
            p.cam_pos[k], p.cam_dir[k], p.cam_up[k] = float(camera.position[k]), float(camera.direction[k]), float(camera.up[k])
        p.x0, p.y0, p.x1, p.y1 = window
        p.stride = stride
        return p

    def render(self, p: OrcRenderParams):
        acc = np.zeros((p.height, p.width, 3), dtype=f32)
        srgb = np.zeros((p.height, p.width, 3), dtype=f32)
        st = OrcStats()
        self.l.orc_render(_vp(self.h), C.byref(p), _p(acc), _p(srgb), C.byref(st))
        return acc, srgb, {k: getattr(st, k) for k, _ in st._fields_}

    def path_samples(self, p: OrcRenderParams, pixels_xy, sample_indices) -> np.ndarray:
        xy = np.ascontiguousarray(pixels_xy, dtype=np.uint32)
        si = np.ascontiguousarray(sample_indices, dtype=np.uint32)
        out = np.zeros((len(si), 3), dtype=f32)
        self.l.orc_path_samples(_vp(self.h), C.byref(p), _p(xy, C.c_uint32), _p(si, C.c_uint32), len(si), _p(out))
        return out

    def record_rays(self, p: OrcRenderParams, max_rays=1 << 20):
        closest = np.zeros((max_rays, 6), dtype=f32)
        shadow = np.zeros((max_rays, 7), dtype=f32)
        nc, ns = C.c_int(0), C.c_int(0)
        self.l.orc_record_rays(_vp(self.h), C.byref(p), _p(closest), max_rays, C.byref(nc), _p(shadow), max_rays, C.byref(ns))
        return closest[: nc.value].copy(), shadow[: ns.value].copy()

    def trace(self, rays: np.ndarray, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=f32)
        out = np.zeros((len(rays), 6), dtype=np.int32)
        nb, nt = C.c_uint64(0), C.c_uint64(0)
        self.l.orc_trace(_vp(self.h), _p(rays), len(rays), int(any_hit), _p(out, C.c_int32), C.byref(nb), C.byref(nt))
        return out, nb.value, nt.value

    def trace_mt(self, rays: np.ndarray, any_hit=False, threads=0):
        """orc_trace over all host threads; returns (hits, seconds, box tests, triangle tests)."""
        rays = np.ascontiguousarray(rays, dtype=f32)
        out = np.zeros((len(rays), 6), dtype=np.int32)
        nb, nt = C.c_uint64(0), C.c_uint64(0)
        sec = self.l.orc_trace_mt(_vp(self.h), _p(rays), len(rays), int(any_hit), _p(out, C.c_int32), int(threads), C.byref(nb), C.byref(nt))
        return out, float(sec), nb.value, nt.value

    def sobol_probe(self, spp, w, h, seed, px, py, sample_index):
        vals = np.zeros(4, dtype=f32)
        idx = np.zeros(3, dtype=np.uint64)
        morton = C.c_uint32(0)
        self.l.orc_sobol_probe(_vp(self.h), C.c_uint32(spp), C.c_uint32(w), C.c_uint32(h), C.c_uint32(seed), C.c_uint32(px), C.c_uint32(py), C.c_uint32(sample_index),
                               _p(vals), _p(idx, C.c_uint64), C.byref(morton))
        return vals, idx, morton.value

    def sampler_stream(self, sampler, spp, w, h, seed, px, py, sample_index, kinds):
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        out = np.zeros(int(sum(1 if x == 1 else 2 for x in kinds)), dtype=f32)
        self.l.orc_sampler_stream(_vp(self.h), {"random": 0, "sobol": 1}[sampler], C.c_uint32(spp), C.c_uint32(w), C.c_uint32(h), C.c_uint32(seed), C.c_uint32(px),
                                  C.c_uint32(py), C.c_uint32(sample_index), _p(k, C.c_int32), len(k), _p(out))
        return out

    def get_bvh(self, which: int) -> np.ndarray:
        n = self.l.orc_get_bvh(_vp(self.h), which, None, 0)
        out = np.zeros((n, 8), dtype=np.uint32)
        self.l.orc_get_bvh(_vp(self.h), which, _p(out, C.c_uint32), n)
        return out

    def rgb_to_coeffs(self, rgb, gamma_encoded=True):
        a = np.asarray(rgb, dtype=f32)
        cs = np.zeros(3, dtype=f32)
        ix = np.zeros(4, dtype=np.int32)
        rc = self.l.orc_rgb_to_coeffs(_vp(self.h), _p(a), int(gamma_encoded), _p(cs), _p(ix, C.c_int32))
        if rc != 0:
            raise ValueError("component > 1 (the reference panics)")
        return cs, ix

    def mesh_tangents(self, geometry: int) -> np.ndarray:
        n = self.l.orc_get_mesh_tangents(_vp(self.h), geometry, None, 0)
        out = np.zeros((n, 3), dtype=f32)
        if n:
            self.l.orc_get_mesh_tangents(_vp(self.h), geometry, _p(out), n)
        return out


def cdf_search(cdf: np.ndarray, u: np.ndarray) -> np.ndarray:
    cdf = np.ascontiguousarray(cdf, dtype=f32); u = np.ascontiguousarray(u, dtype=f32)
    out = np.zeros(len(u), dtype=np.uint32)
    lib().orc_cdf_search(_p(cdf), len(cdf), _p(u), len(u), _p(out, C.c_uint32))
    return out


def build_bvh_boxes(boxes: np.ndarray, literal: bool) -> np.ndarray:
    l = lib()
    b = np.ascontiguousarray(boxes, dtype=f32)
    n = l.orc_build_bvh_boxes(_p(b), len(b), int(literal), None, 0)
    out = np.zeros((n, 8), dtype=np.uint32)
    l.orc_build_bvh_boxes(_p(b), len(b), int(literal), _p(out, C.c_uint32), n)
    return out


def scene_from_description(desc, cam_pos, std_tables, rgb2spec, faithful=False, literal_build=False) -> OracleScene:
    s = OracleScene(std_tables, rgb2spec, faithful=faithful, literal_build=literal_build)
    desc.replay(s)
    s.build(cam_pos)
    return s
