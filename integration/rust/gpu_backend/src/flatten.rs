//! Walks what was handed to `Scene::{load_obj, create_primitive}` into the arrays libtcpt takes (the Rust twin of
//! toy_cpu_pathtracing_b200/csrc/host_scene.cpp + scene.py's SceneDescription::replay, which are the TESTED versions).
//!
//! libtcpt rebuilds the BLAS / TLAS itself with the reference's exact SAH topology (bit-identical, tests/test_bvh_topology.py), so
//! nothing of `scene::Bvh` has to be exported; what it needs is the construction data.  `GpuScene` therefore RECORDS the calls next to
//! forwarding them to the real `Scene`, instead of reaching into `Scene`'s private repositories.
//!
//! The `scene` / `spectrum` crates keep their parameter fields private (`LambertMaterial.albedo` lambert_material.rs:15-20,
//! `TypedRgbTexture.data` rgb_texture.rs:18-21, `FloatTexture.gamma_corrected` float_texture.rs:16-19, ...).  A maintainer adds the one
//! read-only trait below to those crates (`impl GpuDescribe for LambertMaterial { ... }` etc., each a few lines returning its own
//! fields); integration/rust/README.md lists every impl.  Nothing else of the reference changes.
use std::collections::HashMap;
use std::sync::Arc;

use crate::ffi::*;

/// What a material / spectrum / texture tells the flattener about itself.  To be implemented inside `scene` and `spectrum`
/// (see README.md) and reached through `SurfaceMaterial::as_any` / `SpectrumTrait::as_any` downcasts or a `fn gpu(&self)` hook.
pub enum SpectrumDesc {
    Constant(f32),                // ConstantSpectrum::new(c)
    RgbAlbedoSrgb([f32; 3]),      // RgbAlbedoSpectrum::<ColorSrgb<NoneToneMap>>::new(color)   (gamma-encoded value)
    RgbAlbedoLinear([f32; 3]),    // RgbAlbedoSpectrum::<ColorSrgbLinear<NoneToneMap>>::new(color)
    D65,                          // presets::cie_illum_d6500()
    Preset(i32),                  // presets::au_eta() .. glass_sf11_eta(): TCPT_PRESET_* id
}
pub enum SpectrumParamDesc<'a> {
    Constant(SpectrumDesc),
    TextureSrgb { key: usize, rgb8: &'a [u8], width: u32, height: u32 }, // key = Arc::as_ptr of the texture (de-duplication)
}
pub enum FloatParamDesc<'a> {
    Constant(f32),
    Texture { key: usize, gray8: &'a [u8], width: u32, height: u32, gamma_corrected: bool },
}
pub enum NormalParamDesc<'a> {
    None,
    Texture { key: usize, rgb8: &'a [u8], width: u32, height: u32, flip_y: bool },
}
pub enum MaterialDesc<'a> {
    Lambert { albedo: SpectrumParamDesc<'a>, normal: NormalParamDesc<'a> },
    Emissive { radiance: SpectrumParamDesc<'a>, intensity: FloatParamDesc<'a> },
    Plastic { eta: f32, color: SpectrumParamDesc<'a>, normal: NormalParamDesc<'a>, thin_surface: bool, roughness: FloatParamDesc<'a> },
    SimplePbr { base_color: SpectrumParamDesc<'a>, metallic: FloatParamDesc<'a>, roughness: FloatParamDesc<'a>, normal: NormalParamDesc<'a>, ior: FloatParamDesc<'a> },
    ClearcoatPbr {
        base_color: SpectrumParamDesc<'a>, metallic: FloatParamDesc<'a>, roughness: FloatParamDesc<'a>, normal: NormalParamDesc<'a>, ior: FloatParamDesc<'a>,
        coat_ior: FloatParamDesc<'a>, coat_roughness: FloatParamDesc<'a>, coat_thickness: FloatParamDesc<'a>, coat_tint: SpectrumParamDesc<'a>,
    },
    Metal { eta_preset: i32, k_preset: i32, normal: NormalParamDesc<'a>, roughness: FloatParamDesc<'a> },
    Glass { eta_preset: i32, normal: NormalParamDesc<'a>, thin_surface: bool, roughness: FloatParamDesc<'a> },
}
pub trait GpuDescribe {
    fn gpu_material(&self) -> MaterialDesc<'_>;
}

/// One recorded `create_primitive` call (CreatePrimitiveDesc, primitive/create_desc.rs:11-76) in POD form.
pub enum Recorded {
    Geometry { geometry: usize, material: scene::Material, local_to_world: [f32; 16] },
    SingleTriangle { positions: [f32; 9], normals: [f32; 9], uvs: [f32; 6], material: scene::Material, local_to_world: [f32; 16] },
    Delta { kind: i32, intensity: f32, spectrum: SpectrumDesc, angle_inner: f32, angle_outer: f32, local_to_world: [f32; 16] },
    Environment { intensity: f32, rgb: Vec<f32>, width: u32, height: u32, local_to_world: [f32; 16] },
}

fn check(ctx: *mut TcptCtx, rc: i32) -> i32 {
    if rc < 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(tcpt_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("libtcpt: {msg} (code {rc})"); // the reference panics on scene errors too (scene.rs:82,95,109)
    }
    rc
}

pub struct Uploader {
    pub ctx: *mut TcptCtx,
    textures: HashMap<usize, i32>, // Arc pointer -> libtcpt texture index: a texture shared by several materials is uploaded once
}
impl Uploader {
    pub fn new(ctx: *mut TcptCtx) -> Self {
        Self { ctx, textures: HashMap::new() }
    }
    fn texture(&mut self, key: usize, data: &[u8], width: u32, height: u32, channels: u32) -> i32 {
        if let Some(&i) = self.textures.get(&key) {
            return i;
        }
        let i = check(self.ctx, unsafe { tcpt_scene_add_texture(self.ctx, data.as_ptr(), width, height, channels) });
        self.textures.insert(key, i);
        i
    }
    fn spectrum(d: &SpectrumDesc) -> TcptSpectrumParam {
        match *d {
            SpectrumDesc::Constant(c) => TcptSpectrumParam { kind: SPEC_CONSTANT, value: [c, 0.0, 0.0], texture: -1 },
            SpectrumDesc::RgbAlbedoSrgb(v) => TcptSpectrumParam { kind: SPEC_RGB_ALBEDO_SRGB, value: v, texture: -1 },
            SpectrumDesc::RgbAlbedoLinear(v) => TcptSpectrumParam { kind: SPEC_RGB_ALBEDO_LINEAR, value: v, texture: -1 },
            SpectrumDesc::D65 => TcptSpectrumParam { kind: SPEC_D65, value: [0.0; 3], texture: -1 },
            SpectrumDesc::Preset(id) => TcptSpectrumParam { kind: SPEC_PRESET, value: [0.0; 3], texture: id },
        }
    }
    fn spectrum_param(&mut self, p: &SpectrumParamDesc) -> TcptSpectrumParam {
        match p {
            SpectrumParamDesc::Constant(s) => Self::spectrum(s),
            SpectrumParamDesc::TextureSrgb { key, rgb8, width, height } => {
                TcptSpectrumParam { kind: SPEC_TEXTURE_SRGB, value: [0.0; 3], texture: self.texture(*key, rgb8, *width, *height, 3) }
            }
        }
    }
    fn float_param(&mut self, p: &FloatParamDesc) -> TcptFloatParam {
        match p {
            FloatParamDesc::Constant(v) => TcptFloatParam { kind: 0, value: *v, texture: -1, gamma_corrected: 0 },
            FloatParamDesc::Texture { key, gray8, width, height, gamma_corrected } => {
                TcptFloatParam { kind: 1, value: 0.0, texture: self.texture(*key, gray8, *width, *height, 1), gamma_corrected: *gamma_corrected as i32 }
            }
        }
    }
    fn normal_param(&mut self, p: &NormalParamDesc) -> TcptNormalParam {
        match p {
            NormalParamDesc::None => TcptNormalParam::default(),
            NormalParamDesc::Texture { key, rgb8, width, height, flip_y } => {
                TcptNormalParam { texture: self.texture(*key, rgb8, *width, *height, 3), flip_y: *flip_y as i32 }
            }
        }
    }
    /// tcpt_material_desc of one material (field meaning: include/tcpt.h, "material description").
    pub fn material(&mut self, m: &MaterialDesc) -> i32 {
        let mut d = TcptMaterialDesc::default();
        match m {
            MaterialDesc::Lambert { albedo, normal } => {
                d.ty = MAT_LAMBERT; d.color = self.spectrum_param(albedo); d.normal = self.normal_param(normal);
            }
            MaterialDesc::Emissive { radiance, intensity } => {
                d.ty = MAT_EMISSIVE; d.color = self.spectrum_param(radiance); d.intensity = self.float_param(intensity);
            }
            MaterialDesc::Plastic { eta, color, normal, thin_surface, roughness } => {
                d.ty = MAT_PLASTIC; d.eta = *eta; d.color = self.spectrum_param(color); d.normal = self.normal_param(normal);
                d.thin_surface = *thin_surface as i32; d.roughness = self.float_param(roughness);
            }
            MaterialDesc::SimplePbr { base_color, metallic, roughness, normal, ior } => {
                d.ty = MAT_SIMPLE_PBR; d.color = self.spectrum_param(base_color); d.metallic = self.float_param(metallic);
                d.roughness = self.float_param(roughness); d.normal = self.normal_param(normal); d.ior = self.float_param(ior);
            }
            MaterialDesc::ClearcoatPbr { base_color, metallic, roughness, normal, ior, coat_ior, coat_roughness, coat_thickness, coat_tint } => {
                d.ty = MAT_CLEARCOAT_PBR; d.color = self.spectrum_param(base_color); d.metallic = self.float_param(metallic);
                d.roughness = self.float_param(roughness); d.normal = self.normal_param(normal); d.ior = self.float_param(ior);
                d.coat_ior = self.float_param(coat_ior); d.coat_roughness = self.float_param(coat_roughness);
                d.coat_thickness = self.float_param(coat_thickness); d.coat_tint = self.spectrum_param(coat_tint);
            }
            MaterialDesc::Metal { eta_preset, k_preset, normal, roughness } => {
                d.ty = MAT_METAL; d.color = Self::spectrum(&SpectrumDesc::Preset(*eta_preset)); d.coat_tint = Self::spectrum(&SpectrumDesc::Preset(*k_preset));
                d.normal = self.normal_param(normal); d.roughness = self.float_param(roughness);
            }
            MaterialDesc::Glass { eta_preset, normal, thin_surface, roughness } => {
                d.ty = MAT_GLASS; d.color = Self::spectrum(&SpectrumDesc::Preset(*eta_preset)); d.normal = self.normal_param(normal);
                d.thin_surface = *thin_surface as i32; d.roughness = self.float_param(roughness);
            }
        }
        check(self.ctx, unsafe { tcpt_scene_add_material(self.ctx, &d) })
    }

    /// TriangleMesh after tobj (geometry/impls/triangle_mesh.rs:130-139, fields are pub): positions / normals / uvs / indices as they are.
    /// Vertex normals are normalised and per-triangle tangents generated inside libtcpt exactly like `load_obj` does (:168-225).
    pub fn mesh<Id: scene::SceneId>(&mut self, mesh: &scene::geometry::impls::TriangleMesh<Id>) -> i32 {
        let pos: Vec<f32> = mesh.positions.iter().flat_map(|p| [p.x(), p.y(), p.z()]).collect();
        let nrm: Vec<f32> = mesh.normals.iter().flat_map(|n| [n.x(), n.y(), n.z()]).collect();
        let uvs: Vec<f32> = mesh.uvs.iter().flat_map(|t| [t.x, t.y]).collect();
        let uv_ptr = if uvs.is_empty() { std::ptr::null() } else { uvs.as_ptr() };
        check(self.ctx, unsafe {
            tcpt_scene_add_mesh(self.ctx, pos.as_ptr(), nrm.as_ptr(), uv_ptr, mesh.positions.len() as i32, mesh.indices.as_ptr(), (mesh.indices.len() / 3) as i32)
        })
    }

    /// Replays the recorded primitives in creation order (TLAS item order and light order depend on it: primitive/bvh.rs:74-80,
    /// light_sampler.rs:168-187), then `tcpt_scene_build` = Scene::build(&camera) (scene.rs:64-76).
    pub fn primitives(&mut self, recorded: &[Recorded], describe: &dyn Fn(&scene::Material) -> MaterialDesc<'_>, geometry_map: &[i32], cam_pos: [f32; 3]) {
        let mut material_cache: HashMap<usize, i32> = HashMap::new();
        for r in recorded {
            match r {
                Recorded::Geometry { geometry, material, local_to_world } => {
                    let key = Arc::as_ptr(material) as *const () as usize;
                    let mi = *material_cache.entry(key).or_insert_with(|| self.material(&describe(material)));
                    check(self.ctx, unsafe { tcpt_scene_add_primitive(self.ctx, geometry_map[*geometry], mi, local_to_world.as_ptr()) });
                }
                Recorded::SingleTriangle { positions, normals, uvs, material, local_to_world } => {
                    let gi = check(self.ctx, unsafe { tcpt_scene_add_single_triangle(self.ctx, positions.as_ptr(), normals.as_ptr(), uvs.as_ptr()) });
                    let mi = self.material(&describe(material)); // emissive -> EmissiveSingleTriangle inside libtcpt (repository.rs:84-105)
                    check(self.ctx, unsafe { tcpt_scene_add_primitive(self.ctx, gi, mi, local_to_world.as_ptr()) });
                }
                Recorded::Delta { kind, intensity, spectrum, angle_inner, angle_outer, local_to_world } => {
                    let s = Self::spectrum(spectrum);
                    check(self.ctx, unsafe { tcpt_scene_add_delta_light(self.ctx, *kind, *intensity, &s, *angle_inner, *angle_outer, local_to_world.as_ptr()) });
                }
                Recorded::Environment { intensity, rgb, width, height, local_to_world } => {
                    check(self.ctx, unsafe { tcpt_scene_add_env_light(self.ctx, *intensity, rgb.as_ptr(), *width, *height, local_to_world.as_ptr()) });
                }
            }
        }
        check(self.ctx, unsafe { tcpt_scene_build(self.ctx, cam_pos.as_ptr()) });
    }
}
