"""All twenty scenes of the reference's `match args.scene` (renderer/src/main.rs:70-92; BASELINE.json names 3, 10, 17 and 19), restated from
the reference's scene builders with stand-in assets.

Materials, transforms, lights and cameras are the reference's (renderer/src/scene/scene_N.rs); every mesh,
texture and HDRI it loads is a git-LFS stub in the checkout, so geometry and images come from the
deterministic generators in assets.py (Cornell walls = axis-aligned quads spanning x,z in [-2.5,2.5], y in [0,5]).
"""
from __future__ import annotations

import functools

import numpy as np

from . import assets
from .scene import (ColorSrgb, ColorSrgbLinear, ConstantSpectrum, CreatePrimitiveDesc, EmissiveMaterial, FloatParameter, FloatTexture,
                    GlassMaterial, GlassType, LambertMaterial, MetalMaterial, MetalType, NormalParameter, NormalTexture, PlasticMaterial, RgbAlbedoSpectrum, RgbTexture, SimpleClearcoatPbrMaterial,
                    SimplePbrMaterial, SpectrumParameter, SpectrumType, Transform, presets)

GP = CreatePrimitiveDesc.GeometryPrimitive


# the reference's own asset files (renderer/assets/, git-LFS payloads): used instead of the stand-ins when TCPT_ASSET_DIR points at a
# checkout that holds them; an LFS pointer stub or a missing file falls back to the procedural stand-in
_REAL_ASSETS = {
    "bunny": "bunny.obj", "dragon": "dragon.min.obj", "box": "box.obj", "light": "light.obj", "hidari": "hidari.obj", "migi": "migi.obj",
    "yuka": "yuka.obj", "oku": "oku.obj", "tenjou": "tenjou.obj",
    "bunny_basecolor": "bunny-material-0/BaseColor.png", "bunny_normal": "bunny-material-0/Normal.png",
    "bunny1_basecolor": "bunny-material-1/BaseColor.png", "bunny1_normal": "bunny-material-1/Normal.png",
    "dragon_basecolor": "dragon-material/BaseColor.png", "dragon_normal": "dragon-material/Normal.png", "dragon_metallic": "dragon-material/Metallic.png",
    "dragon_roughness": "dragon-material/Roughness.png", "dragon_coat_thickness": "dragon-material/ClearcoatThickness.png",
    "sky": "sky/scythian_tombs_2_1k.exr",
}
_GRAY = {"dragon_metallic", "dragon_roughness", "dragon_coat_thickness"}


def real_asset_path(name: str):
    """Path of the reference's own file for `name` if TCPT_ASSET_DIR holds its payload, else None."""
    import os
    from pathlib import Path
    root = os.environ.get("TCPT_ASSET_DIR")
    if not root or name not in _REAL_ASSETS:
        return None
    p = Path(root) / _REAL_ASSETS[name]
    if not p.is_file() or p.stat().st_size < 1024 and p.read_bytes().startswith(b"version https://git-lfs"):
        return None
    return p


def _load_real(name: str, path):
    from .scene import _load_image, load_obj
    if str(path).endswith(".obj"):
        return load_obj(path)
    if str(path).endswith(".exr"):
        from .scene import convert_image, decode_image
        return convert_image(decode_image(path), "rgb32f")
    return _load_image(path, 1 if name in _GRAY else 3)


@functools.lru_cache(maxsize=None)
def _asset(name: str):
    real = real_asset_path(name)
    if real is not None:
        return _load_real(name, real)
    if name == "bunny":
        return assets.blob(center=(-0.8, 1.0, 0.4), radius=1.0, nu=50, nv=50)
    if name == "dragon":
        return assets.knot(scale=0.12, nu=250, nv=40)
    if name == "box":
        return assets.box((0.5, 0.0, -1.7), (1.9, 2.8, -0.3), rot_y_deg=-18.0)
    if name == "light":
        return assets.box((-0.75, 4.90, -0.75), (0.75, 4.99, 0.75))
    h = 2.5
    if name == "hidari":   # left wall, faces +x
        return assets.quad((-h, 0, h), (-h, 0, -h), (-h, 5, -h), (-h, 5, h), (1, 0, 0))
    if name == "migi":     # right wall, faces -x
        return assets.quad((h, 0, -h), (h, 0, h), (h, 5, h), (h, 5, -h), (-1, 0, 0))
    if name == "yuka":     # floor, faces +y
        return assets.quad((-h, 0, h), (h, 0, h), (h, 0, -h), (-h, 0, -h), (0, 1, 0))
    if name == "oku":      # back wall, faces +z
        return assets.quad((-h, 0, -h), (h, 0, -h), (h, 5, -h), (-h, 5, -h), (0, 0, 1))
    if name == "tenjou":   # ceiling, faces -y
        return assets.quad((-h, 5, -h), (h, 5, -h), (h, 5, h), (-h, 5, h), (0, -1, 0))
    if name == "bunny_basecolor":
        return assets.base_color_texture(1024, seed=0)
    if name == "bunny_normal":
        return assets.normal_texture(1024, seed=10)
    if name == "bunny1_basecolor":
        return assets.base_color_texture(1024, seed=1)
    if name == "bunny1_normal":
        return assets.normal_texture(1024, seed=11)
    if name == "dragon_coat_thickness":
        return assets.gray_texture(1024, seed=22, lo=0.1, hi=0.95)
    if name == "dragon_basecolor":
        return assets.base_color_texture(1024, seed=3)
    if name == "dragon_normal":
        return assets.normal_texture(1024, seed=13)
    if name == "dragon_metallic":
        return assets.gray_texture(1024, seed=20, threshold=0.5)
    if name == "dragon_roughness":
        return assets.gray_texture(1024, seed=21, lo=0.15, hi=0.85)
    if name == "sky":
        return assets.sky_hdri(1024, 512)
    raise KeyError(name)


def _grey(v=0.8):
    return SpectrumParameter.constant(RgbAlbedoSpectrum(ColorSrgb(v, v, v)))


def _lambert(r, g, b):
    return LambertMaterial.new(SpectrumParameter.constant(RgbAlbedoSpectrum(ColorSrgb(r, g, b))), NormalParameter.none())


def _cornell_rest(scene, with_box=True, with_lamp=True):
    """box, hidari, migi, yuka, oku, tenjou, light: identical in scenes 3, 10 and 17 (scene_3.rs:33-108)."""
    if with_box:
        scene.create_primitive(GP(scene.load_obj(_asset("box")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("hidari")), _lambert(0.9, 0.0, 0.0), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("migi")), _lambert(0.0, 0.9, 0.0), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("yuka")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("oku")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("tenjou")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    if with_lamp:
        scene.create_primitive(GP(scene.load_obj(_asset("light")),
                              EmissiveMaterial.new(SpectrumParameter.constant(presets.cie_illum_d6500()), FloatParameter.constant(10.0)), Transform.identity()))


def _cornell_camera(camera):
    d = np.array([0.0, -0.9, -3.2], dtype=np.float32)
    camera.set_look_to((0.0, 3.15221, 6.0), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def load_scene_3(scene, camera):
    """Cornell box with a textured, normal-mapped Lambert bunny (scene_3.rs:13-115)."""
    tex = RgbTexture.load_srgb(_asset("bunny_basecolor"))
    nrm = NormalTexture.load(_asset("bunny_normal"), False)
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")),
                              LambertMaterial.new(SpectrumParameter.texture(tex, SpectrumType.Albedo), NormalParameter.texture(nrm)), Transform.identity()))
    _cornell_rest(scene)
    _cornell_camera(camera)


def load_scene_10(scene, camera):
    """Same box; the bunny is thin-film plastic, eta 1.8, roughness 0 (scene_10.rs:13-112)."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")),
                              PlasticMaterial.new(1.8, SpectrumParameter.Constant(ConstantSpectrum(1.0)), NormalParameter.none(), True, FloatParameter.constant(0.0)),
                              Transform.identity()))
    _cornell_rest(scene)
    d = np.array([0.0, -1.0, -3.0], dtype=np.float32)
    camera.set_look_to((0.0, 3.5, 6.0), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def _camera_10(camera):
    d = np.array([0.0, -1.0, -3.0], dtype=np.float32)
    camera.set_look_to((0.0, 3.5, 6.0), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def load_scene_0(scene, camera):
    """Plain Cornell box: grey Lambert bunny and box, coloured walls, ceiling lamp (scene_0.rs:12-107)."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_4(scene, camera):
    """Scene 3 with the second bunny material (bunny-material-1 base colour + normal map) and the scene-10 camera (scene_4.rs:12-115)."""
    tex = RgbTexture.load_srgb(_asset("bunny1_basecolor"))
    nrm = NormalTexture.load(_asset("bunny1_normal"), False)
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")),
                              LambertMaterial.new(SpectrumParameter.texture(tex, SpectrumType.Albedo), NormalParameter.texture(nrm)), Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_5(scene, camera):
    """Grey Lambert bunny with the normal map only, close-up camera (scene_5.rs:12-113)."""
    nrm = NormalTexture.load(_asset("bunny_normal"), False)
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), LambertMaterial.new(_grey(0.8), NormalParameter.texture(nrm)), Transform.identity()))
    _cornell_rest(scene)
    d = np.array([0.0, -0.5, -2.0], dtype=np.float32)
    camera.set_look_to((0.3, 1.6, 2.8), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def load_scene_1(scene, camera):
    """Two bunny instances (one rotated 30 degrees about Y and translated), the floor, one inline SingleTriangle rotated 60 degrees and
    two point lights (scene_1.rs:12-87)."""
    bunny = scene.load_obj(_asset("bunny"))
    scene.create_primitive(GP(bunny, _lambert(0.5, 0.5, 0.8), Transform.identity()))
    scene.create_primitive(GP(bunny, _lambert(0.5, 0.8, 0.5), Transform.from_rotate_y(30.0).translate((-1.0, 1.0, 3.0))))
    scene.create_primitive(GP(scene.load_obj(_asset("yuka")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(CreatePrimitiveDesc.SingleTrianglePrimitive(
        [(-2.0, 0.0, 0.0), (2.0, 0.0, 0.0), (-2.0, 4.0, 0.0)], [(0.0, 0.0, 1.0)] * 3, [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0)],
        _lambert(0.8, 0.5, 0.5), Transform.from_rotate_y(60.0)))
    scene.create_primitive(CreatePrimitiveDesc.PointLightPrimitive(10.0, presets.cie_illum_d6500(), Transform.from_translate((0.0, 3.0, 0.0))))
    scene.create_primitive(CreatePrimitiveDesc.PointLightPrimitive(10.0, presets.cie_illum_d6500(), Transform.from_translate((3.0, 5.0, 0.0))))
    d = np.array([0.0, -1.0, -3.0], dtype=np.float32)
    camera.set_look_to((0.0, 3.5, 7.0), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def load_scene_2(scene, camera):
    """Cornell box without the ceiling lamp, lit by one point light at (0, 3, 0) (scene_2.rs:12-102): delta-light NEE, no MIS weight."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("box")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("hidari")), _lambert(0.9, 0.0, 0.0), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("migi")), _lambert(0.0, 0.9, 0.0), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("yuka")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("oku")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(GP(scene.load_obj(_asset("tenjou")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    scene.create_primitive(CreatePrimitiveDesc.PointLightPrimitive(10.0, presets.cie_illum_d6500(), Transform.from_translate((0.0, 3.0, 0.0))))
    _camera_10(camera)


def load_scene_lights(scene, camera, directional=False):
    """Not a reference scene: the Cornell box under a point and a spot light next to the lamp, so that the power-weighted light choice
    mixes delta and area lights (the reference builds Spot and Directional lights in no scene of its own;
    primitive/impls/{spot,directional}_light.rs).  directional=True adds a directional light: its shadow rays start ON the surface
    without any offset (common.rs:70-72), so whether a surface shadows itself is decided by rounding noise -- a knife edge of the
    reference that parity tests have to bound separately."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    _cornell_rest(scene)
    scene.create_primitive(CreatePrimitiveDesc.PointLightPrimitive(6.0, presets.cie_illum_d6500(), Transform.from_translate((-1.5, 2.0, 1.0))))
    scene.create_primitive(CreatePrimitiveDesc.SpotLightPrimitive(0.3, 0.9, 3.0, RgbAlbedoSpectrum(ColorSrgb(1.0, 0.7, 0.4)),
                                                                  Transform.identity().rotate_y(200.0).translate((1.5, 3.5, 2.0))))
    if directional:
        scene.create_primitive(CreatePrimitiveDesc.DirectionalLightPrimitive(0.8, ConstantSpectrum(1.0), Transform.identity().rotate_y(35.0)))
    _camera_10(camera)


def load_scene_tri_lamp(scene, camera, textured=True):
    """Not a reference scene: the Cornell box lit by an EmissiveSingleTriangle instead of the lamp mesh (the reference builds that
    primitive in no scene of its own; primitive/impls/emissive_single_triangle.rs, chosen by repository.rs:84-105 when a
    SingleTrianglePrimitive carries an emissive material).  A textured intensity shows the one place where it differs from a
    one-triangle EmissiveTriangleMesh: the light sample's uv is the pair of random numbers (:190-252), a hit's uv is interpolated."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    _cornell_rest(scene, with_lamp=False)
    intensity = FloatParameter.texture(FloatTexture.load(assets.gray_texture(256, seed=33, lo=0.3, hi=1.0), False)) if textured else FloatParameter.constant(0.7)
    scene.create_primitive(CreatePrimitiveDesc.SingleTrianglePrimitive(
        [(-1.2, 0.0, -1.0), (1.2, 0.0, -1.0), (0.0, 0.0, 1.3)], [(0.0, -1.0, 0.0)] * 3, [(0.0, 0.0), (1.0, 0.0), (0.5, 1.0)],
        EmissiveMaterial.new(SpectrumParameter.constant(presets.cie_illum_d6500()), intensity), Transform.from_rotate_y(20.0).translate((0.2, 4.6, 0.1))))
    _camera_10(camera)


def load_scene_6(scene, camera):
    """Cornell box with a mirror-smooth gold bunny (scene_6.rs:13-111): MetalMaterial / ConductorBsdf, specular reflection."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), MetalMaterial.new(MetalType.Gold, NormalParameter.none(), FloatParameter.constant(0.0)), Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_7(scene, camera):
    """Four instances of the bunny in gold with roughness 0.05 / 0.25 / 0.5 / 0.75 (scene_7.rs:13-131): rough conductor, instance transforms."""
    bunny = scene.load_obj(_asset("bunny"))
    for pos, rough in zip(((-1.3, 0.0, -0.5), (-0.5, 0.0, -0.5), (0.3, 0.0, -0.5), (1.1, 0.0, -0.5)), (0.05, 0.25, 0.5, 0.75)):
        scene.create_primitive(GP(bunny, MetalMaterial.new(MetalType.Gold, NormalParameter.none(), FloatParameter.constant(rough)),
                                  Transform.identity().scale((0.6, 0.6, 0.6)).translate(pos)))
    _cornell_rest(scene)
    d = np.array([0.0, -0.7, -2.5], dtype=np.float32)
    camera.set_look_to((0.0, 2.5, 5.0), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def load_scene_8(scene, camera):
    """SF11 glass bunny, smooth, not thin (scene_8.rs:13-112): dispersive eta(lambda) -> the path keeps only its hero wavelength."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), GlassMaterial.new(GlassType.Sf11, NormalParameter.none(), False, FloatParameter.constant(0.0)), Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_9(scene, camera):
    """Same box; the bunny is solid plastic, eta 1.8, roughness 0, thin film off (scene_9.rs:13-113)."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")),
                              PlasticMaterial.new(1.8, SpectrumParameter.Constant(ConstantSpectrum(1.0)), NormalParameter.none(), False, FloatParameter.constant(0.0)),
                              Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


_FOUR_POSITIONS = ((-1.3, 0.0, -0.5), (-0.5, 0.0, -0.5), (0.3, 0.0, -0.5), (1.1, 0.0, -0.5))


def load_scene_11(scene, camera):
    """SF11 glass bunny with roughness 0.2 (scene_11.rs:12-117): rough dielectric reflection and transmission."""
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), GlassMaterial.new(GlassType.Sf11, NormalParameter.none(), False, FloatParameter.constant(0.2)), Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_12(scene, camera):
    """Four BK7 glass bunnies, roughness 0.05 / 0.25 / 0.5 / 0.75 (scene_12.rs:12-130)."""
    bunny = scene.load_obj(_asset("bunny"))
    for pos, rough in zip(_FOUR_POSITIONS, (0.05, 0.25, 0.5, 0.75)):
        scene.create_primitive(GP(bunny, GlassMaterial.new(GlassType.Bk7, NormalParameter.none(), False, FloatParameter.constant(rough)),
                                  Transform.from_scale((0.6, 0.6, 0.6)).translate(pos)))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_13(scene, camera):
    """Solid light-blue plastic bunny, eta 1.5, smooth (scene_13.rs:12-120)."""
    color = SpectrumParameter.constant(RgbAlbedoSpectrum(ColorSrgbLinear(0.4, 0.9, 1.0)))
    scene.create_primitive(GP(scene.load_obj(_asset("bunny")), PlasticMaterial.new(1.5, color, NormalParameter.none(), False, FloatParameter.constant(0.0)), Transform.identity()))
    _cornell_rest(scene)
    _camera_10(camera)


def load_scene_14(scene, camera):
    """Four coloured plastic bunnies, roughness 0.05 / 0.1 / 0.3 / 0.5 (scene_14.rs:12-140)."""
    bunny = scene.load_obj(_asset("bunny"))
    colors = ((1.0, 0.5, 0.5), (0.5, 1.0, 0.5), (0.5, 0.5, 1.0), (1.0, 0.8, 0.4))
    for pos, rough, c in zip(_FOUR_POSITIONS, (0.05, 0.1, 0.3, 0.5), colors):
        mat = PlasticMaterial.new(1.5, SpectrumParameter.constant(RgbAlbedoSpectrum(ColorSrgb(*c))), NormalParameter.none(), False, FloatParameter.constant(rough))
        scene.create_primitive(GP(bunny, mat, Transform.from_scale((0.6, 0.6, 0.6)).translate(pos)))
    _cornell_rest(scene)
    _camera_10(camera)


def _dragon_transform():
    return Transform.identity().rotate_y(120.0).scale((2.5, 2.5, 2.5)).translate((0.0, 0.0, 0.5))


def load_scene_15(scene, camera):
    """Textured SimplePbr dragon (base colour, metallic, roughness, normal map) in the Cornell box (scene_15.rs:12-155)."""
    pbr = SimplePbrMaterial.new(SpectrumParameter.texture(RgbTexture.load_srgb(_asset("dragon_basecolor")), SpectrumType.Albedo),
                                FloatParameter.texture(FloatTexture.load(_asset("dragon_metallic"), False)),
                                FloatParameter.texture(FloatTexture.load(_asset("dragon_roughness"), False)),
                                NormalParameter.texture(NormalTexture.load(_asset("dragon_normal"), False)), FloatParameter.constant(1.5))
    scene.create_primitive(GP(scene.load_obj(_asset("dragon")), pbr, _dragon_transform()))
    _cornell_rest(scene)
    _cornell_camera(camera)


def load_scene_16(scene, camera):
    """Scene 17 with a smooth coat (clearcoat roughness 0.01) (scene_16.rs:12-155)."""
    scene.create_primitive(GP(scene.load_obj(_asset("dragon")), _clearcoat(0.01), _dragon_transform()))
    _cornell_rest(scene)
    _cornell_camera(camera)


def load_scene_18(scene, camera):
    """Smooth clearcoat whose thickness comes from a texture (scene_18.rs:12-160)."""
    scene.create_primitive(GP(scene.load_obj(_asset("dragon")),
                              _clearcoat(0.01, FloatParameter.texture(FloatTexture.load(_asset("dragon_coat_thickness"), False))), _dragon_transform()))
    _cornell_rest(scene)
    _cornell_camera(camera)


def _clearcoat(coat_roughness, thickness=None):
    tint = SpectrumParameter.constant(RgbAlbedoSpectrum(ColorSrgb(0.7, 0.8, 1.0)))
    return SimpleClearcoatPbrMaterial.new(_grey(0.8), FloatParameter.constant(1.0), FloatParameter.constant(0.7), NormalParameter.none(), FloatParameter.constant(1.5),
                                          FloatParameter.constant(1.5), FloatParameter.constant(coat_roughness), tint, thickness or FloatParameter.constant(0.8))


def load_scene_17(scene, camera, coat=True):
    """Clearcoat-PBR dragon in the Cornell box (scene_17.rs:13-155).  coat=False swaps in the uncoated SimplePbr substrate
    (BASELINE.json config 3 'no-coat'; the reference has no such toggle, only pictures/scene17.no-coat.mis.png)."""
    mat = _clearcoat(0.75) if coat else SimplePbrMaterial.new(_grey(0.8), FloatParameter.constant(1.0), FloatParameter.constant(0.7), NormalParameter.none(), FloatParameter.constant(1.5))
    xf = Transform.identity().rotate_y(120.0).scale((2.5, 2.5, 2.5)).translate((0.0, 0.0, 0.5))
    scene.create_primitive(GP(scene.load_obj(_asset("dragon")), mat, xf))
    _cornell_rest(scene)
    _cornell_camera(camera)


def load_scene_19(scene, camera):
    """Floor + three instances of one dragon mesh (textured SimplePbr, smooth clearcoat, blue plastic) under an environment light (scene_19.rs:17-153)."""
    scene.create_primitive(GP(scene.load_obj(_asset("yuka")), _lambert(0.8, 0.8, 0.8), Transform.identity()))
    dragon = scene.load_obj(_asset("dragon"))
    pbr = SimplePbrMaterial.new(SpectrumParameter.texture(RgbTexture.load_srgb(_asset("dragon_basecolor")), SpectrumType.Albedo),
                                FloatParameter.texture(FloatTexture.load(_asset("dragon_metallic"), False)),
                                FloatParameter.texture(FloatTexture.load(_asset("dragon_roughness"), False)),
                                NormalParameter.texture(NormalTexture.load(_asset("dragon_normal"), False)), FloatParameter.constant(1.5))
    scene.create_primitive(GP(dragon, pbr, Transform.identity()))
    scene.create_primitive(GP(dragon, _clearcoat(0.01), Transform.identity().translate((0.5, 0.0, 0.5))))
    plastic = PlasticMaterial.new(1.5, SpectrumParameter.constant(RgbAlbedoSpectrum(ColorSrgbLinear(0.4, 0.9, 1.0))), NormalParameter.none(), False, FloatParameter.constant(0.0))
    scene.create_primitive(GP(dragon, plastic, Transform.identity().translate((-0.5, 0.0, -0.5))))
    scene.create_primitive(CreatePrimitiveDesc.EnvironmentLightPrimitive(1.0, _asset("sky"), Transform.identity()))
    d = np.array([1.5, -0.4, -2.5], dtype=np.float32)
    camera.set_look_to((-1.5, 0.8, 2.5), d / np.float32(np.sqrt(np.float32((d * d).sum()))), (0.0, 1.0, 0.0))


def load_soup(scene, camera, n_triangles: int, seed: int = 42):
    """BASELINE.json config 5: synthetic triangle soup lit by a small emissive quad; camera at (0,0,3) looking at the origin."""
    scene.create_primitive(GP(scene.load_obj(assets.triangle_soup(n_triangles, seed)), _lambert(0.7, 0.7, 0.7), Transform.identity()))
    light = assets.quad((-0.5, 1.6, -0.5), (0.5, 1.6, -0.5), (0.5, 1.6, 0.5), (-0.5, 1.6, 0.5), (0, -1, 0))
    scene.create_primitive(GP(scene.load_obj(light), EmissiveMaterial.new(SpectrumParameter.constant(presets.cie_illum_d6500()), FloatParameter.constant(10.0)), Transform.identity()))
    camera.set_look_to((0.0, 0.0, 3.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0))


SCENES = {0: load_scene_0, 1: load_scene_1, 2: load_scene_2, 4: load_scene_4, 5: load_scene_5, 11: load_scene_11, 12: load_scene_12, 13: load_scene_13,
          14: load_scene_14, 15: load_scene_15, 16: load_scene_16, 18: load_scene_18, "lights": load_scene_lights, "tri_lamp": load_scene_tri_lamp, 3: load_scene_3, 6: load_scene_6, 7: load_scene_7, 8: load_scene_8, 9: load_scene_9, 10: load_scene_10, 17: load_scene_17, 19: load_scene_19}


def load_scene(scene_id, scene, camera, **kw):
    """The `match args.scene` of renderer/src/main.rs:70-92 for the scenes in scope."""
    if scene_id not in SCENES:
        raise ValueError(f"scene {scene_id} is outside the B200 hot-path scope (available: {sorted(SCENES)})")
    SCENES[scene_id](scene, camera, **kw)
