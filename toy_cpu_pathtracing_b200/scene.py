"""Host-side mirror of the reference's `scene` crate construction API (scene/src/scene.rs:36-76, scene/src/lib.rs re-exports).

A `Scene` records meshes, textures, materials and primitives with the reference's own vocabulary
(`load_obj`/`add_mesh`, `create_primitive(GeometryPrimitive | EnvironmentLightPrimitive)`, `LambertMaterial::new`, ...) as a
neutral description.  `Scene.build(camera)` hands it to libtcpt (C++ BVH builder with the reference's exact SAH topology,
flattening, upload).  The same description can be replayed into the CPU oracle by tests (oracle/oracle.py), which is how
parity is checked; the product never touches the oracle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import capi
from .assets import MeshData

f32 = np.float32


# ------------------------------------------------------------------ spectra / parameters (spectrum crate, scene/src/material/parameter.rs)
@dataclass
class ColorSrgb:            # color::ColorSrgb<NoneToneMap>::new(r, g, b): gamma-encoded sRGB
    r: float
    g: float
    b: float
    gamma_encoded = True


@dataclass
class ColorSrgbLinear:      # color::ColorSrgbLinear::new
    r: float
    g: float
    b: float
    gamma_encoded = False


@dataclass
class ConstantSpectrum:     # spectrum::ConstantSpectrum::new(c)
    c: float


@dataclass
class RgbAlbedoSpectrum:    # spectrum::RgbAlbedoSpectrum::<C>::new(color)
    color: object


@dataclass
class PresetSpectrum:       # one of the DenselySampledSpectrum presets of spectrum/src/presets.rs:336-462 (metal eta/k, glass eta)
    name: str


class _Presets:
    @staticmethod
    def cie_illum_d6500():
        return "D65"

    def __getattr__(self, name):     # presets::au_eta(), presets::glass_sf11_eta(), ...
        if name in capi.PRESETS:
            return lambda: PresetSpectrum(name)
        raise AttributeError(name)


presets = _Presets()


@dataclass
class RgbTexture:           # scene::RgbTexture::load_srgb -> RGB8, gamma encoded
    data: np.ndarray        # (H, W, 3) uint8

    @staticmethod
    def load_srgb(array_or_path):
        return RgbTexture(_load_image(array_or_path, 3))


@dataclass
class FloatTexture:         # scene::FloatTexture::load(path, gamma_corrected) -> gray8
    data: np.ndarray        # (H, W) uint8
    gamma_corrected: bool = False

    @staticmethod
    def load(array_or_path, gamma_corrected=False):
        return FloatTexture(_load_image(array_or_path, 1), gamma_corrected)


@dataclass
class NormalTexture:        # scene::NormalTexture::load(path, flip_y)
    data: np.ndarray
    flip_y: bool = False

    @staticmethod
    def load(array_or_path, flip_y=False):
        return NormalTexture(_load_image(array_or_path, 3), flip_y)


def decode_image(path) -> np.ndarray:
    """Decoder only (OpenCV): the file's own channels and sample depth, channel order L | LA | RGB | RGBA, dtype uint8 | uint16 | float32."""
    import os
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    a = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if a is None:
        raise FileNotFoundError(path)
    if a.ndim == 3 and a.shape[2] >= 3:
        a = a[..., [2, 1, 0] + ([3] if a.shape[2] == 4 else [])]        # BGR(A) -> RGB(A)
    if a.dtype not in (np.uint8, np.uint16, np.float32):
        a = a.astype(np.float32)
    return np.ascontiguousarray(a)


IMG_KINDS = {"rgb8": (0, np.uint8, 3), "luma8": (1, np.uint8, 1), "rgb32f": (2, np.float32, 3)}


def convert_image(a: np.ndarray, kind: str) -> np.ndarray:
    """DynamicImage::to_rgb8 / to_luma8 / to_rgb32f of the `image` crate on a decoded array (libtcpt: csrc/host_image.h)."""
    a = np.ascontiguousarray(a)
    hgt, wid = a.shape[:2]
    ch = 1 if a.ndim == 2 else a.shape[2]
    code, dtype, out_ch = IMG_KINDS[kind]
    st = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}[a.dtype]
    out = np.zeros((hgt, wid, out_ch) if out_ch == 3 else (hgt, wid), dtype=dtype)
    rc = capi.load_library().tcpt_image_convert(a.ctypes.data_as(C.c_void_p), wid, hgt, ch, st, code, out.ctypes.data_as(C.c_void_p))
    if rc != capi.TCPT_OK:
        raise ValueError("tcpt_image_convert: unsupported image layout")
    return out


def _load_image(src, channels):
    """texture/loader.rs:43-87: load_rgb_image (channels 3) keeps Rgb8 and converts everything else with to_rgb8; load_grayscale_image
    (channels 1) keeps Luma8, drops the alpha of LumaA8 and converts everything else with to_luma8.  Arrays are taken as already loaded."""
    if isinstance(src, np.ndarray):
        a = np.ascontiguousarray(src, dtype=np.uint8)
        return a[..., 0] if channels == 1 and a.ndim == 3 else a
    a = decode_image(src)
    return convert_image(a, "rgb8" if channels == 3 else "luma8")


class SpectrumType:
    Albedo = "albedo"


@dataclass
class SpectrumParameter:
    spectrum: object = None
    tex: Optional[RgbTexture] = None

    @staticmethod
    def constant(spectrum):
        return SpectrumParameter(spectrum=spectrum)

    Constant = constant

    @staticmethod
    def texture(tex: RgbTexture, spectrum_type=SpectrumType.Albedo):
        return SpectrumParameter(tex=tex)


@dataclass
class FloatParameter:
    value: float = 0.0
    tex: Optional[FloatTexture] = None

    @staticmethod
    def constant(v):
        return FloatParameter(value=float(v))

    @staticmethod
    def texture(tex: FloatTexture):
        return FloatParameter(tex=tex)


@dataclass
class NormalParameter:
    tex: Optional[NormalTexture] = None

    @staticmethod
    def none():
        return NormalParameter()

    @staticmethod
    def texture(tex: NormalTexture):
        return NormalParameter(tex=tex)


# ------------------------------------------------------------------ materials (scene/src/material/impls/*.rs constructors)
@dataclass
class LambertMaterial:
    albedo: SpectrumParameter
    normal: NormalParameter

    @classmethod
    def new(cls, albedo, normal):
        return cls(albedo, normal)


@dataclass
class EmissiveMaterial:
    radiance: SpectrumParameter
    intensity: FloatParameter

    @classmethod
    def new(cls, radiance, intensity):
        return cls(radiance, intensity)


@dataclass
class PlasticMaterial:
    eta: float
    color: SpectrumParameter
    normal: NormalParameter
    thin_surface: bool
    roughness: FloatParameter

    @classmethod
    def new(cls, eta, color, normal, thin_surface, roughness):
        return cls(eta, color, normal, thin_surface, roughness)


@dataclass
class SimplePbrMaterial:
    base_color: SpectrumParameter
    metallic: FloatParameter
    roughness: FloatParameter
    normal: NormalParameter
    ior: FloatParameter

    @classmethod
    def new(cls, base_color, metallic, roughness, normal, ior):
        return cls(base_color, metallic, roughness, normal, ior)


@dataclass
class SimpleClearcoatPbrMaterial:
    base_color: SpectrumParameter
    metallic: FloatParameter
    roughness: FloatParameter
    normal: NormalParameter
    ior: FloatParameter
    clearcoat_ior: FloatParameter
    clearcoat_roughness: FloatParameter
    clearcoat_tint_color: SpectrumParameter
    clearcoat_thickness: FloatParameter

    @classmethod
    def new(cls, *a):
        return cls(*a)


class MetalType:            # metal_material.rs:15-27; value = (eta preset, k preset) as MetalMaterial::get_eta / get_k choose them (:84-108)
    Gold = ("au_eta", "au_k")
    Silver = ("ag_eta", "ag_k")
    Copper = ("cu_eta", "cu_k")
    Aluminum = ("al_eta", "al_k")
    Brass = ("cu_zn_eta", "cu_zn_k")


@dataclass
class MetalMaterial:        # metal_material.rs:39-75
    metal_type: tuple
    normal: NormalParameter
    roughness: FloatParameter

    @classmethod
    def new(cls, metal_type, normal, roughness):
        return cls(metal_type, normal, roughness)

    new_with_roughness = new


class GlassType:            # glass_material.rs:13-29; value = eta preset (GlassMaterial::get_eta, :49-60)
    Bk7 = "glass_bk7_eta"
    Baf10 = "glass_baf10_eta"
    Fk51a = "glass_fk51a_eta"
    Lasf9 = "glass_lasf9_eta"
    Sf5 = "glass_sf5_eta"
    Sf10 = "glass_sf10_eta"
    Sf11 = "glass_sf11_eta"


@dataclass
class GlassMaterial:        # glass_material.rs:31-47
    glass_type: str
    normal: NormalParameter
    thin_surface: bool
    roughness: FloatParameter

    @classmethod
    def new(cls, glass_type, normal, thin_surface, roughness):
        return cls(glass_type, normal, thin_surface, roughness)


# ------------------------------------------------------------------ transforms (math/src/transform.rs:84-161): T.translate(v) = translation * T, etc.
class Transform:
    def __init__(self, m=None):
        self.m = np.eye(4, dtype=f32) if m is None else np.asarray(m, dtype=f32)

    @staticmethod
    def identity():
        return Transform()

    @staticmethod
    def from_translate(v):      # Transform::from_translate (math/src/transform.rs:100-104)
        return Transform().translate(v)

    @staticmethod
    def from_scale(v):
        return Transform().scale(v)

    @staticmethod
    def from_rotate_y(degrees):  # Transform::from_rotate(glam::Quat::from_rotation_y(deg.to_radians()))
        return Transform().rotate_y(degrees)

    def translate(self, v):
        t = np.eye(4, dtype=f32)
        t[:3, 3] = np.asarray(v, dtype=f32)
        return Transform((t @ self.m).astype(f32))

    def scale(self, v):
        s = np.diag(np.array([v[0], v[1], v[2], 1.0], dtype=f32))
        return Transform((s @ self.m).astype(f32))

    def rotate_y(self, degrees):
        a = np.deg2rad(f32(degrees))
        c, s = f32(np.cos(a)), f32(np.sin(a))
        r = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=f32)
        return Transform((r @ self.m).astype(f32))

    def column_major(self) -> np.ndarray:
        return np.ascontiguousarray(self.m.T.reshape(-1), dtype=f32)


@dataclass
class GeometryPrimitive:            # CreatePrimitiveDesc::GeometryPrimitive
    geometry_index: int
    surface_material: object
    transform: Transform = field(default_factory=Transform)


@dataclass
class EnvironmentLightPrimitive:    # CreatePrimitiveDesc::EnvironmentLightPrimitive (texture given as an (H,W,3) f32 array or an EXR path)
    intensity: float
    texture: object
    transform: Transform = field(default_factory=Transform)


@dataclass
class SingleTrianglePrimitive:      # CreatePrimitiveDesc::SingleTrianglePrimitive { positions, normals, uvs, surface_material, transform }
    positions: object
    normals: object
    uvs: object
    surface_material: object
    transform: Transform = field(default_factory=Transform)


@dataclass
class PointLightPrimitive:          # CreatePrimitiveDesc::PointLightPrimitive { intensity, spectrum, transform }
    intensity: float
    spectrum: object
    transform: Transform = field(default_factory=Transform)


@dataclass
class SpotLightPrimitive:           # CreatePrimitiveDesc::SpotLightPrimitive { angle_inner, angle_outer, intensity, spectrum, transform }
    angle_inner: float
    angle_outer: float
    intensity: float
    spectrum: object
    transform: Transform = field(default_factory=Transform)


@dataclass
class DirectionalLightPrimitive:    # CreatePrimitiveDesc::DirectionalLightPrimitive { intensity, spectrum, transform }
    intensity: float
    spectrum: object
    transform: Transform = field(default_factory=Transform)


class CreatePrimitiveDesc:
    GeometryPrimitive = GeometryPrimitive
    EnvironmentLightPrimitive = EnvironmentLightPrimitive
    SingleTrianglePrimitive = SingleTrianglePrimitive
    PointLightPrimitive = PointLightPrimitive
    SpotLightPrimitive = SpotLightPrimitive
    DirectionalLightPrimitive = DirectionalLightPrimitive


def load_obj(path) -> MeshData:
    """TriangleMesh::load_obj (geometry/impls/triangle_mesh.rs:141-243) through libtcpt's loader (csrc/host_obj.h: tobj 4.0.3's
    `single_index + triangulate` rules and the reference's concatenation of the models, quirks included).  Host-side only: needs no GPU."""
    lib = capi.load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    if lib.tcpt_obj_load(str(path).encode(), C.byref(h), err, len(err)) != capi.TCPT_OK:
        raise ValueError(f"load_obj({path}): {err.value.decode()}")
    try:
        counts = (C.c_uint32 * 5)()
        lib.tcpt_obj_counts(h, counts)
        nv, nn, nt, ntri, nmodels = (int(c) for c in counts)
        pos, nrm, uvs = np.zeros((nv, 3), f32), np.zeros((nn, 3), f32), np.zeros((nt, 2), f32)
        idx, ttri = np.zeros((ntri, 3), np.uint32), np.zeros(ntri, np.uint32)
        lib.tcpt_obj_copy(h, capi.as_ptr(pos, C.c_float), capi.as_ptr(nrm, C.c_float), capi.as_ptr(uvs, C.c_float), capi.as_ptr(idx, C.c_uint32), capi.as_ptr(ttri, C.c_uint32))
    finally:
        lib.tcpt_obj_free(h)
    if nn != nv:
        raise ValueError("OBJ files must carry vn normals on every face vertex (the reference panics without them, triangle_mesh.rs:57-61)")
    if nt not in (0, nv):
        raise ValueError("texcoords on only some of the vertices (the reference would misindex them)")
    identity = np.array_equal(ttri, np.arange(ntri, dtype=np.uint32))
    return MeshData(pos, nrm, uvs if nt else None, idx, tangent_tri=None if identity else ttri)


# ------------------------------------------------------------------ scene description
class SceneDescription:
    """Neutral record of everything added to a Scene; replayable into any backend exposing the add_*/build protocol."""

    def __init__(self):
        self.meshes: list[MeshData] = []
        self.textures: list[np.ndarray] = []
        self._tex_ids: dict[int, tuple] = {}   # id(array) -> (index, array kept alive so the id stays unique)
        self.materials: list[capi.MaterialDesc] = []
        self.primitives: list[tuple] = []   # ("geom", geometry, material, l2w16) | ("env", intensity, rgb, l2w16) | ("delta", kind, intensity, spectrum, inner, outer, l2w16)

    # --- textures are de-duplicated by object identity so one image shared by several parameters is uploaded once
    def _texture(self, arr: np.ndarray) -> int:
        key = id(arr)
        if key not in self._tex_ids:
            self._tex_ids[key] = (len(self.textures), arr)
            self.textures.append(np.ascontiguousarray(arr, dtype=np.uint8))
        return self._tex_ids[key][0]

    def _spectrum(self, p: SpectrumParameter) -> capi.SpectrumParam:
        out = capi.SpectrumParam(capi.SPEC_CONSTANT, (C.c_float * 3)(0, 0, 0), -1)
        if p is None:
            return out
        if p.tex is not None:
            out.kind, out.texture = capi.SPEC_TEXTURE_SRGB, self._texture(p.tex.data)
            return out
        s = p.spectrum
        if isinstance(s, ConstantSpectrum):
            out.kind = capi.SPEC_CONSTANT
            out.value[0] = s.c
        elif isinstance(s, RgbAlbedoSpectrum):
            out.kind = capi.SPEC_RGB_ALBEDO_SRGB if s.color.gamma_encoded else capi.SPEC_RGB_ALBEDO_LINEAR
            out.value[0], out.value[1], out.value[2] = s.color.r, s.color.g, s.color.b
        elif isinstance(s, PresetSpectrum):
            out.kind, out.texture = capi.SPEC_PRESET, capi.PRESETS.index(s.name)
        elif s == "D65":
            out.kind = capi.SPEC_D65
        else:
            raise TypeError(f"unsupported spectrum {s!r}")
        return out

    def _float(self, p: Optional[FloatParameter]) -> capi.FloatParam:
        if p is None:
            return capi.FloatParam(0, 0.0, -1, 0)
        if p.tex is not None:
            return capi.FloatParam(1, 0.0, self._texture(p.tex.data), int(p.tex.gamma_corrected))
        return capi.FloatParam(0, p.value, -1, 0)

    def _normal(self, p: Optional[NormalParameter]) -> capi.NormalParam:
        if p is None or p.tex is None:
            return capi.NormalParam(-1, 0)
        return capi.NormalParam(self._texture(p.tex.data), int(p.tex.flip_y))

    def material_desc(self, m) -> capi.MaterialDesc:
        d = capi.MaterialDesc()
        d.normal = capi.NormalParam(-1, 0)
        d.color = self._spectrum(None)
        d.coat_tint = self._spectrum(None)
        for name in ("intensity", "roughness", "metallic", "ior", "coat_ior", "coat_roughness", "coat_thickness"):
            setattr(d, name, self._float(None))
        d.eta = 1.5
        if isinstance(m, LambertMaterial):
            d.type, d.color, d.normal = capi.MAT_LAMBERT, self._spectrum(m.albedo), self._normal(m.normal)
        elif isinstance(m, EmissiveMaterial):
            d.type, d.color, d.intensity = capi.MAT_EMISSIVE, self._spectrum(m.radiance), self._float(m.intensity)
        elif isinstance(m, PlasticMaterial):
            d.type, d.eta, d.color, d.normal = capi.MAT_PLASTIC, m.eta, self._spectrum(m.color), self._normal(m.normal)
            d.thin_surface, d.roughness = int(m.thin_surface), self._float(m.roughness)
        elif isinstance(m, (SimplePbrMaterial, SimpleClearcoatPbrMaterial)):
            d.type = capi.MAT_SIMPLE_PBR if isinstance(m, SimplePbrMaterial) else capi.MAT_CLEARCOAT_PBR
            d.color, d.metallic, d.roughness = self._spectrum(m.base_color), self._float(m.metallic), self._float(m.roughness)
            d.normal, d.ior = self._normal(m.normal), self._float(m.ior)
            if isinstance(m, SimpleClearcoatPbrMaterial):
                d.coat_ior, d.coat_roughness = self._float(m.clearcoat_ior), self._float(m.clearcoat_roughness)
                d.coat_tint, d.coat_thickness = self._spectrum(m.clearcoat_tint_color), self._float(m.clearcoat_thickness)
        elif isinstance(m, MetalMaterial):
            d.type, d.normal, d.roughness = capi.MAT_METAL, self._normal(m.normal), self._float(m.roughness)
            d.color = self._spectrum(SpectrumParameter.constant(PresetSpectrum(m.metal_type[0])))
            d.coat_tint = self._spectrum(SpectrumParameter.constant(PresetSpectrum(m.metal_type[1])))
        elif isinstance(m, GlassMaterial):
            d.type, d.normal, d.roughness = capi.MAT_GLASS, self._normal(m.normal), self._float(m.roughness)
            d.thin_surface = int(m.thin_surface)
            d.color = self._spectrum(SpectrumParameter.constant(PresetSpectrum(m.glass_type)))
        else:
            raise TypeError(f"unsupported material {m!r}")
        return d

    def replay(self, backend):
        """backend: add_mesh(pos, nrm, uv|None, idx) / add_texture(arr) / add_material(desc) / add_primitive(g, m, l2w) / add_env_light(i, rgb, l2w) /
        add_delta_light(kind, intensity, spectrum_param, angle_inner, angle_outer, l2w)"""
        for mesh in self.meshes:
            if mesh.single:
                backend.add_single_triangle(mesh.positions, mesh.normals, mesh.uvs)
            else:
                g = backend.add_mesh(mesh.positions, mesh.normals, mesh.uvs, mesh.indices)
                if mesh.tangent_tri is not None:
                    backend.set_tangent_source(g, mesh.tangent_tri)
        for t in self.textures:
            backend.add_texture(t)
        for m in self.materials:
            backend.add_material(m)
        for p in self.primitives:
            if p[0] == "geom":
                backend.add_primitive(p[1], p[2], p[3])
            elif p[0] == "delta":
                backend.add_delta_light(*p[1:])
            else:
                backend.add_env_light(p[1], p[2], p[3])


class Scene:
    """scene::Scene (scene/src/scene.rs:36-76) over libtcpt."""

    def __init__(self, device: int = 0, context: Optional[capi.Context] = None, require_gpu: bool = True, describe_only: bool = False):
        # describe_only: record the scene (self.desc) without creating a libtcpt context -- bench.py's reference arm replays the
        # description into the CPU oracle and must not load the product library
        self.ctx = None if describe_only else (context or capi.Context(device, require_gpu=require_gpu))
        self.desc = SceneDescription()
        self.built = False

    # Scene::load_obj (scene.rs:54-57): returns the geometry index
    def load_obj(self, path_or_mesh) -> int:
        mesh = path_or_mesh if isinstance(path_or_mesh, MeshData) else load_obj(path_or_mesh)
        self.desc.meshes.append(mesh)
        return len(self.desc.meshes) - 1

    add_mesh = load_obj

    # Scene::create_primitive (scene.rs:59-62)
    def create_primitive(self, desc) -> int:
        d = self.desc
        if isinstance(desc, GeometryPrimitive):
            d.materials.append(d.material_desc(desc.surface_material))
            d.primitives.append(("geom", desc.geometry_index, len(d.materials) - 1, desc.transform.column_major()))
        elif isinstance(desc, EnvironmentLightPrimitive):
            tex = desc.texture
            if not isinstance(tex, np.ndarray):
                tex = convert_image(decode_image(tex), "rgb32f")       # image::open(path).to_rgb32f() (environment_light.rs:36-37)
            d.primitives.append(("env", float(desc.intensity), np.ascontiguousarray(tex, dtype=f32), desc.transform.column_major()))
        elif isinstance(desc, SingleTrianglePrimitive):
            tri = MeshData(np.asarray(desc.positions, dtype=f32).reshape(3, 3), np.asarray(desc.normals, dtype=f32).reshape(3, 3),
                           np.asarray(desc.uvs, dtype=f32).reshape(3, 2), np.array([[0, 1, 2]], dtype=np.uint32), single=True)
            d.meshes.append(tri)
            d.materials.append(d.material_desc(desc.surface_material))
            d.primitives.append(("geom", len(d.meshes) - 1, len(d.materials) - 1, desc.transform.column_major()))
        elif isinstance(desc, (PointLightPrimitive, SpotLightPrimitive, DirectionalLightPrimitive)):
            kind = capi.LIGHT_POINT if isinstance(desc, PointLightPrimitive) else capi.LIGHT_SPOT if isinstance(desc, SpotLightPrimitive) else capi.LIGHT_DIRECTIONAL
            spec = d._spectrum(SpectrumParameter.constant(desc.spectrum))
            d.primitives.append(("delta", kind, float(desc.intensity), spec, float(getattr(desc, "angle_inner", 0.0)), float(getattr(desc, "angle_outer", 0.0)),
                                 desc.transform.column_major()))
        else:
            raise TypeError(desc)
        return len(d.primitives) - 1

    # Scene::build(&camera) (scene.rs:64-76): bakes the camera position, builds BLAS/TLAS, flattens, uploads
    def build(self, camera) -> None:
        lib, h, ctx = self.ctx.lib, self.ctx.handle, self.ctx
        ctx.check(lib.tcpt_scene_clear(h))
        self.desc.replay(self)
        pos = np.asarray(camera.position, dtype=f32)
        ctx.check(lib.tcpt_scene_build(h, capi.as_ptr(pos, C.c_float)), allow_no_gpu=True)
        self.built = True

    # BASELINE.json configs[4]: a triangle soup as a traversal-only scene whose BVH is built on the device (tcpt_scene_build_soup, csrc/lbvh.cuh)
    def build_soup(self, triangles) -> None:
        """triangles: (n, 3, 3) f32 vertices in world space (= Render space: one primitive, identity transform, primitive index 0)."""
        tri = np.ascontiguousarray(triangles, dtype=f32).reshape(-1, 9)
        self.ctx.check(self.ctx.lib.tcpt_scene_build_soup(self.ctx.handle, capi.as_ptr(tri, C.c_float), len(tri)))
        self.built = True

    def soup_build_info(self) -> dict:
        ms, rec, lev = C.c_double(0), C.c_uint64(0), C.c_uint32(0)
        self.ctx.check(self.ctx.lib.tcpt_soup_build_info(self.ctx.handle, C.byref(ms), C.byref(rec), C.byref(lev)))
        return {"build_ms": ms.value, "records": rec.value, "levels": lev.value}

    # --- backend protocol used by SceneDescription.replay
    def add_mesh(self, pos, nrm, uv, idx):
        lib, h = self.ctx.lib, self.ctx.handle
        uvp = capi.as_ptr(uv, C.c_float) if uv is not None else None
        return self.ctx.check(lib.tcpt_scene_add_mesh(h, capi.as_ptr(pos, C.c_float), capi.as_ptr(nrm, C.c_float), uvp, len(pos), capi.as_ptr(idx, C.c_uint32), len(idx)))

    def set_tangent_source(self, geometry, tri):
        return self.ctx.check(self.ctx.lib.tcpt_scene_set_tangent_source(self.ctx.handle, geometry, capi.as_ptr(tri, C.c_uint32), len(tri)))

    def add_single_triangle(self, pos, nrm, uv):
        return self.ctx.check(self.ctx.lib.tcpt_scene_add_single_triangle(self.ctx.handle, capi.as_ptr(pos, C.c_float), capi.as_ptr(nrm, C.c_float), capi.as_ptr(uv, C.c_float)))

    def add_texture(self, arr):
        hgt, wid = arr.shape[:2]
        ch = 1 if arr.ndim == 2 else arr.shape[2]
        return self.ctx.check(self.ctx.lib.tcpt_scene_add_texture(self.ctx.handle, capi.as_ptr(arr, C.c_uint8), wid, hgt, ch))

    def add_material(self, desc):
        return self.ctx.check(self.ctx.lib.tcpt_scene_add_material(self.ctx.handle, C.byref(desc)))

    def add_primitive(self, geometry, material, l2w):
        return self.ctx.check(self.ctx.lib.tcpt_scene_add_primitive(self.ctx.handle, geometry, material, capi.as_ptr(l2w, C.c_float)))

    def add_env_light(self, intensity, rgb, l2w):
        hgt, wid = rgb.shape[:2]
        return self.ctx.check(self.ctx.lib.tcpt_scene_add_env_light(self.ctx.handle, intensity, capi.as_ptr(rgb, C.c_float), wid, hgt, capi.as_ptr(l2w, C.c_float)))

    def add_delta_light(self, kind, intensity, spectrum, angle_inner, angle_outer, l2w):
        return self.ctx.check(self.ctx.lib.tcpt_scene_add_delta_light(self.ctx.handle, kind, intensity, C.byref(spectrum), angle_inner, angle_outer, capi.as_ptr(l2w, C.c_float)))

    # --- introspection used by the bit-exactness tests
    def get_bvh(self, which: int) -> np.ndarray:
        n = self.ctx.check(self.ctx.lib.tcpt_get_bvh(self.ctx.handle, which, None, 0))
        out = np.zeros((n, 8), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.tcpt_get_bvh(self.ctx.handle, which, capi.as_ptr(out, C.c_uint32), n))
        return out

    def get_wide_bvh(self, which: int):
        """(records [n, 8 rows, 4 children] u32, first_record, slot_base) of the device's 4-wide layout (include/tcpt_flat.h)."""
        first, sbase = C.c_uint32(0), C.c_uint32(0)
        n = self.ctx.check(self.ctx.lib.tcpt_get_wide_bvh(self.ctx.handle, which, None, 0, C.byref(first), C.byref(sbase)))
        out = np.zeros((n, 8, 4), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.tcpt_get_wide_bvh(self.ctx.handle, which, capi.as_ptr(out, C.c_uint32), n, C.byref(first), C.byref(sbase)))
        return out, first.value, sbase.value

    def sampler_stream(self, sampler, spp, width, height, seed, px, py, sample_index, kinds) -> np.ndarray:
        """The values a sampler hands out for one (pixel, sample): kinds = 1 (get_1d) / 2 (get_2d) per call (parity probe)."""
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        out = np.zeros(int(sum(1 if x == 1 else 2 for x in kinds)), dtype=f32)
        self.ctx.check(self.ctx.lib.tcpt_sampler_stream(self.ctx.handle, capi.SAMPLERS[sampler], spp, width, height, seed, px, py, sample_index,
                                                        capi.as_ptr(k, C.c_int32), len(k), capi.as_ptr(out, C.c_float)))
        return out

    def mesh_tangents(self, geometry: int) -> np.ndarray:
        n = self.ctx.check(self.ctx.lib.tcpt_get_mesh_tangents(self.ctx.handle, geometry, None, 0))
        out = np.zeros((n, 3), dtype=f32)
        if n:
            self.ctx.lib.tcpt_get_mesh_tangents(self.ctx.handle, geometry, capi.as_ptr(out, C.c_float), n)
        return out

    def rgb_to_coeffs(self, rgb, gamma_encoded=True):
        a = np.asarray(rgb, dtype=f32)
        cs = np.zeros(3, dtype=f32)
        ix = np.zeros(4, dtype=np.int32)
        self.ctx.check(self.ctx.lib.tcpt_rgb_to_coeffs(self.ctx.handle, capi.as_ptr(a, C.c_float), int(gamma_encoded), capi.as_ptr(cs, C.c_float), capi.as_ptr(ix, C.c_int32)))
        return cs, ix

    def trace(self, rays: np.ndarray, any_hit=False) -> np.ndarray:
        """rays: (n,7) f32 {o, d, tmax} -> (n,6) int32 {prim, tri, t bits, b0 bits, b1 bits, b2 bits}"""
        rays = np.ascontiguousarray(rays, dtype=f32)
        out = np.zeros((len(rays), 6), dtype=np.int32)
        self.ctx.check(self.ctx.lib.tcpt_trace(self.ctx.handle, capi.as_ptr(rays, C.c_float), len(rays), int(any_hit), capi.as_ptr(out, C.c_int32)))
        return out
