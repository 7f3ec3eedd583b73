// ORACLE (test infrastructure).  Camera, film sensor and the pt / nee / mis integrators, restated from
//   /root/reference/renderer/src/{camera,filter,sensor,tone_map}.rs,
//   renderer/src/renderer/{base_renderer,common,pt_renderer,nee_renderer,mis_renderer}.rs.
#pragma once
#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

#include "oscene.h"

namespace orc {

enum Integrator : int { INTEG_PT = 0, INTEG_NEE = 1, INTEG_MIS = 2, INTEG_ALBEDO = 3, INTEG_NORMAL = 4 };
enum SamplerKind : int { SAMPLER_RANDOM = 0, SAMPLER_SOBOL = 1 };

struct Camera {
    Vec3 position = Vec3(0, 0, 0), direction = Vec3(0, 0, -1), up = Vec3(0, 1, 0);
    float fov = 45.0f;
    uint32_t width = 800, height = 600;
    float filter_width = 1.0f;
    // Camera::set_look_to (camera.rs:39-48)
    void set_look_to(Vec3 pos, Vec3 dir, Vec3 up_) { position = pos; direction = normalize(dir); up = normalize(up_); }
    // Camera::generate_ray (camera.rs:51-65)
    Ray generate_ray(float x, float y) const {
        float aspect = (float)width / (float)height;
        float fov_rad = fov * (PI_F / 180.0f);  // f32::to_radians
        float scale = std::tan(fov_rad / 2.0f);
        float dir_x = (2.0f * x / (float)width - 1.0f) * aspect * scale;
        float dir_y = (1.0f - 2.0f * y / (float)height) * scale;
        Vec3 rd = normalize(Vec3(dir_x, dir_y, -1.0f));
        // glam Mat3::look_to_rh(dir, up): f = normalize(dir), s = normalize(f x up), u = s x f,
        // cols (s.x,u.x,-f.x),(s.y,u.y,-f.y),(s.z,u.z,-f.z); the camera multiplies by its transpose => cols s, u, -f
        Vec3 f = direction;  // already normalised by set_look_to; glam only asserts it
        Vec3 s = normalize(cross(f, up));
        Vec3 u = cross(s, f);
        Vec3 nf = -f;
        Vec3 d = (s * rd.x + u * rd.y) + nf * rd.z;
        return Ray{Vec3(0, 0, 0), normalize(d)};
    }
    // Camera::sample_ray + BoxFilter::sample (camera.rs:68-81, filter.rs:24-30)
    Ray sample_ray(uint32_t px, uint32_t py, Vec2 uv) const {
        float fx = uv.x * filter_width - filter_width * 0.5f;
        float fy = uv.y * filter_width - filter_width * 0.5f;
        float x = (float)px + fx + 0.5f;
        float y = (float)py + fy + 0.5f;
        return generate_ray(x, y);
    }
};

// Sensor (sensor.rs:41-88), Reinhard (tone_map.rs:20-28)
struct Sensor {
    Vec3 acc = Vec3(0, 0, 0);
    float exposure = 1.0f;
    static Vec3 sample_to_rgb(const Tables& T, const SampledWavelengths& wl, const SampledSpectrum& s, float exposure) {
        int count = wl.is_secondary_terminated() ? 1 : NS;
        Vec3 xyz(0, 0, 0);
        for (int k = 0; k < count; ++k) {
            float l = wl.lambda[k];
            uint32_t i = f2u_sat(std::floor(l - LAMBDA_MIN));
            if (i == (uint32_t)N_DENSE) i = 0;
            float pdfk = wl.pdf[k];
            float a = s.v[k] / pdfk;  // `s.value(index) / pdf.value(index)`: plain f32 division (sensor.rs:62-63)
            float nc = a / (float)NS;
            float lam_cmf = LAMBDA_MIN + (float)i;
            xyz.x += nc * dense_value(T.cie_x, lam_cmf);
            xyz.y += nc * dense_value(T.cie_y, lam_cmf);
            xyz.z += nc * dense_value(T.cie_z, lam_cmf);
        }
        Vec3 rgb = mul(T.xyz_to_rgb, xyz);
        return rgb * exposure;
    }
    void add_sample(const Tables& T, const SampledWavelengths& wl, const SampledSpectrum& s) { acc = acc + sample_to_rgb(T, wl, s, exposure); }
    static Vec3 finalize_no_tone_map(Vec3 acc, uint32_t spp) {  // Sensor<GamutSrgb, NoneToneMap, GammaSrgb>::to_rgb (albedo_renderer.rs:38-39)
        Vec3 c = vmax(acc / (float)spp, Vec3(0, 0, 0));
        return Vec3(srgb_oetf(c.x), srgb_oetf(c.y), srgb_oetf(c.z));
    }
    static Vec3 finalize(Vec3 acc, uint32_t spp) {
        Vec3 avg = acc / (float)spp;
        Vec3 c = vmax(avg, Vec3(0, 0, 0));
        Vec3 tm(c.x / (1.0f + c.x), c.y / (1.0f + c.y), c.z / (1.0f + c.z));
        return Vec3(srgb_oetf(tm.x), srgb_oetf(tm.y), srgb_oetf(tm.z));
    }
};

struct RenderParams {
    uint32_t width, height, spp, seed, max_depth;
    int integrator, sampler;
    float exposure;
};

struct ProbeRecord {  // optional per-path trace used by parity tests
    std::vector<Ray>* closest_rays = nullptr;
    std::vector<Ray>* shadow_rays = nullptr;
    std::vector<float>* shadow_tmax = nullptr;
};

inline float balance_heuristic(float a, float b) { return (a == 0.0f && b == 0.0f) ? 0.0f : a / (a + b); }  // common.rs:15-20

struct PathTracer {
    const Scene& scene;
    const Camera& cam;
    RenderParams rp;
    ProbeRecord probe;

    PathTracer(const Scene& s, const Camera& c, const RenderParams& p) : scene(s), cam(c), rp(p) {}

    bool evaluate_emissive(const MaterialContext& mc, const Intersection& is, const SampledWavelengths& wl, SampledSpectrum* out) const {
        const Material& m = scene.materials[is.si.material];
        if (m.type != MAT_EMISSIVE) return false;
        *out = scene.emissive_radiance(mc, m, is.si.uv, wl);  // base_renderer.rs:54-73 (uniform EDF: frame independent)
        return true;
    }

    // next ray construction (base_renderer.rs:111-121)
    static Ray spawn_ray(const SurfaceInteraction& si, Vec3 wi_render) {
        float sign = dot(si.normal, wi_render) < 0.0f ? -1.0f : 1.0f;
        Vec3 origin = si.position + (sign * si.normal) * 1e-5f;
        return Ray{origin, wi_render}.move_forward(1e-5f);
    }

    // NEE for all strategies: returns contribution and MIS weight (nee_renderer.rs:18-102, mis_renderer.rs:21-123, common.rs)
    void nee(const MaterialContext& mc, SamplerBase& smp, const SampledWavelengths& wl, const Mat4& r2t, const Intersection& hit,
             bool with_mis, SampledSpectrum* contribution, float* mis_weight, RayStats* st) const {
        *contribution = SampledSpectrum::zero();
        *mis_weight = 1.0f;
        const Material& mat = scene.materials[hit.si.material];
        LightSampler ls = scene.light_sampler(mc, wl);
        float u = smp.get_1d();
        float p_light;
        int li = scene.sample_light(ls, u, &p_light);
        if (li < 0) return;
        float s = smp.get_1d();
        Vec2 uv = smp.get_2d();
        int prim = scene.light_list[li];
        const SurfaceInteraction& sp = hit.si;
        TangentShadingPoint tsp;
        auto make_tsp = [&]() { tsp.normal = transform_normal(r2t, sp.normal); tsp.uv = sp.uv; };
        if (scene.primitives[prim].kind == PRIM_ENV_LIGHT) {
            Scene::InfiniteSample is = scene.sample_infinite_light(prim, wl, uv);
            Ray shadow = Ray{sp.position, is.wi}.move_forward(1e-4f);
            if (probe.shadow_rays) { probe.shadow_rays->push_back(shadow); probe.shadow_tmax->push_back(std::numeric_limits<float>::max()); }
            bool visible = !scene.intersect_p(shadow, std::numeric_limits<float>::max(), st);
            if (!visible) return;
            Vec3 wo = transform_vector3(r2t, hit.wo), wi = transform_vector3(r2t, is.wi);
            make_tsp();
            SampledSpectrum f = material_evaluate(mc, mat, wl, wo, wi, tsp);
            if (with_mis) *mis_weight = balance_heuristic(is.pdf_dir, material_pdf(mc, mat, wl, wo, wi, tsp));
            *contribution = f * is.radiance / (is.pdf_dir * p_light);
            return;
        }
        const Primitive& LP = scene.primitives[prim];
        if (LP.kind == PRIM_POINT_LIGHT || LP.kind == PRIM_SPOT_LIGHT) {
            // PointLight / SpotLight::calculate_intensity (point_light.rs:75-88, spot_light.rs:98-122) + evaluate_delta_point_light (common.rs:23-55)
            Vec3 lpos = transform_point3(LP.local_to_render, Vec3(0, 0, 0));
            SampledSpectrum inten = LP.light_spectrum.sample(scene.T, wl) * LP.light_intensity;
            if (LP.kind == PRIM_SPOT_LIGHT) {
                Vec3 wi_l = normalize(lpos - sp.position);
                // quirk: the cosine of the angle to the axis is compared with the cone ANGLES themselves, smoothstep(outer, inner, cos)
                float cos_theta = transform_vector3(LP.render_to_local, wi_l).z;
                float t = clampf((cos_theta - LP.angle_outer) / (LP.angle_inner - LP.angle_outer), 0.0f, 1.0f);
                inten = inten * (t * t * (3.0f - 2.0f * t));
            }
            Vec3 dv = lpos - sp.position;
            Ray shadow = Ray{sp.position, normalize(dv)}.move_forward(1e-4f);
            float t = length(dv) - 2.0f * 1e-4f;
            if (probe.shadow_rays) { probe.shadow_rays->push_back(shadow); probe.shadow_tmax->push_back(t); }
            if (scene.intersect_p(shadow, t, st)) return;
            Vec3 wo = transform_vector3(r2t, hit.wo), wi = transform_vector3(r2t, normalize(dv));
            make_tsp();
            SampledSpectrum f = material_evaluate(mc, mat, wl, wo, wi, tsp);
            *contribution = f * inten / (length_squared(dv) * p_light);  // delta lights: no MIS (weight stays 1)
            return;
        }
        if (LP.kind == PRIM_DIRECTIONAL_LIGHT) {
            // DirectionalLight::calculate_intensity (directional_light.rs:92-107) + evaluate_delta_directional_light (common.rs:58-79):
            // the shadow ray starts ON the surface (no forward offset) and is unbounded
            Vec3 dir = normalize(transform_vector3(LP.local_to_render, Vec3(0, 0, 1)));
            SampledSpectrum inten = LP.light_spectrum.sample(scene.T, wl) * LP.light_intensity;
            Ray shadow = Ray{sp.position, dir};
            if (probe.shadow_rays) { probe.shadow_rays->push_back(shadow); probe.shadow_tmax->push_back(std::numeric_limits<float>::max()); }
            if (scene.intersect_p(shadow, std::numeric_limits<float>::max(), st)) return;
            Vec3 wo = transform_vector3(r2t, hit.wo), wi = transform_vector3(r2t, normalize(dir));
            make_tsp();
            SampledSpectrum f = material_evaluate(mc, mat, wl, wo, wi, tsp);
            *contribution = f * inten / p_light;
            return;
        }
        Scene::AreaSample as = scene.sample_area_light(mc, prim, sp.position, wl, s, uv);
        Vec3 dv = as.position - sp.position;
        Ray shadow = Ray{sp.position, normalize(dv)}.move_forward(1e-4f);
        float t = length(dv) - 2.0f * 1e-4f;
        if (probe.shadow_rays) { probe.shadow_rays->push_back(shadow); probe.shadow_tmax->push_back(t); }
        bool visible = !scene.intersect_p(shadow, t, st);
        if (!visible) return;
        Vec3 wo = transform_vector3(r2t, hit.wo);
        Vec3 wi = transform_vector3(r2t, normalize(dv));
        make_tsp();
        SampledSpectrum f = material_evaluate(mc, mat, wl, wo, wi, tsp);
        float distance2 = length_squared(dv);
        Vec3 light_normal = transform_normal(r2t, as.light_normal);
        float cos_light = std::fabs(dot(light_normal, -wi));
        float g = cos_light / distance2;
        if (with_mis) *mis_weight = balance_heuristic(as.pdf_dir, material_pdf(mc, mat, wl, wo, wi, tsp));
        *contribution = f * as.radiance * g / (as.pdf * p_light);
    }

    // BaseSrgbRenderer::render, one (pixel, sample_index) path (base_renderer.rs:160-276); returns the RGB the sensor would add
    // AlbedoRenderer / NormalRenderer (renderer/src/renderer/{albedo,normal}_renderer.rs): one camera ray, no offset along the ray
    Vec3 aov_path(SamplerBase& smp, uint32_t px, uint32_t py, uint32_t sample_index, RayStats* st) const {
        const Tables& T = scene.T;
        smp.start_pixel_sample(px, py, sample_index);
        const float FMAX = std::numeric_limits<float>::max();
        if (rp.integrator == INTEG_NORMAL) {  // normal_renderer.rs:33-69: the pixel sample is the FIRST thing drawn
            Vec2 uvp = smp.get_2d_pixel();
            Ray ray = cam.sample_ray(px, py, uvp);
            Intersection hit;
            if (!scene.intersect(ray, FMAX, &hit, st)) return Vec3(0, 0, 0);
            Vec3 n = hit.si.shading_normal;
            if (scene.materials[hit.si.material].type != MAT_EMISSIVE) n = transform_normal(shading_transform(hit.si), hit.si.shading_normal);
            return Vec3(n.x * 0.5f + 0.5f, n.y * 0.5f + 0.5f, n.z * 0.5f + 0.5f) * 1.0f;
        }
        float u = smp.get_1d();  // albedo_renderer.rs:41-66
        SampledWavelengths wl = SampledWavelengths::new_uniform(u);
        Vec2 uvp = smp.get_2d_pixel();
        Ray ray = cam.sample_ray(px, py, uvp);
        Intersection hit;
        if (!scene.intersect(ray, FMAX, &hit, st)) return Vec3(0, 0, 0);
        const Material& m = scene.materials[hit.si.material];
        if (m.type == MAT_EMISSIVE) return Vec3(0, 0, 0);
        MaterialContext mc{&T, &scene.textures, smp.aux_base(), 0};
        SampledSpectrum a = material_albedo(mc, m, wl, hit.si.uv) * 1.0f;
        SampledSpectrum out;
        for (int i = 0; i < NS; ++i) out.v[i] = a.v[i] * dense_value(T.d65, wl.lambda[i]);  // multiply_spectrum (sampled_spectrum.rs:270-281)
        return Sensor::sample_to_rgb(T, wl, out, 1.0f);
    }

    Vec3 trace_path(SamplerBase& smp, uint32_t px, uint32_t py, uint32_t sample_index, RayStats* st) const {
        if (rp.integrator >= INTEG_ALBEDO) return aov_path(smp, px, py, sample_index, st);
        const Tables& T = scene.T;
        smp.start_pixel_sample(px, py, sample_index);
        MaterialContext mc{&T, &scene.textures, smp.aux_base(), 0};
        SampledSpectrum throughput = SampledSpectrum::one();
        SampledSpectrum contribution = SampledSpectrum::zero();
        float u = smp.get_1d();
        SampledWavelengths wl = SampledWavelengths::new_uniform(u);
        Vec2 uvp = smp.get_2d_pixel();
        Ray ray = cam.sample_ray(px, py, uvp).move_forward(1e-5f);
        throughput *= 1.0f;  // filter weight
        const float FMAX = std::numeric_limits<float>::max();
        Intersection hit;
        if (probe.closest_rays) probe.closest_rays->push_back(ray);
        if (!scene.intersect(ray, FMAX, &hit, st)) {
            contribution += throughput * scene.evaluate_infinite_light_radiance(ray.d, wl);
            return Sensor::sample_to_rgb(T, wl, contribution, rp.exposure);
        }
        SampledSpectrum le;
        if (evaluate_emissive(mc, hit, wl, &le)) contribution += throughput * le;

        for (uint32_t depth = 1; depth <= rp.max_depth; ++depth) {
            const Material& mat = scene.materials[hit.si.material];
            if (mat.type == MAT_EMISSIVE) break;  // no BSDF (base_renderer.rs:199-202)
            mc.depth = depth;
            Mat4 r2t = shading_transform(hit.si);
            Vec3 wo = transform_vector3(r2t, hit.wo);
            TangentShadingPoint tsp;
            tsp.normal = transform_normal(r2t, hit.si.normal);
            tsp.uv = hit.si.uv;
            float uc = smp.get_1d();
            Vec2 uv = smp.get_2d();
            MaterialSample ms = material_sample(mc, mat, uc, uv, wl, wo, tsp);

            if (ms.is_non_specular() && rp.integrator != INTEG_PT) {
                SampledSpectrum c;
                float w;
                nee(mc, smp, wl, r2t, hit, rp.integrator == INTEG_MIS, &c, &w, st);
                if (rp.integrator == INTEG_MIS) contribution += throughput * c * w;  // mis_renderer.rs:148
                else contribution += throughput * c;                                 // nee_renderer.rs:126
            }

            // process_bsdf_sampling (base_renderer.rs:95-139)
            bool have_next = false;
            Intersection next;
            Ray next_ray{};
            Vec3 wi_render(0, 0, 0);
            if (ms.is_sampled) {
                wi_render = transform_vector3(inverse(r2t), ms.wi);
                next_ray = spawn_ray(hit.si, wi_render);
                if (probe.closest_rays) probe.closest_rays->push_back(next_ray);
                have_next = scene.intersect(next_ray, FMAX, &next, st);
            }
            if (!have_next) {
                // calculate_bsdf_infinite_light_contribution
                if (ms.is_sampled && rp.integrator != INTEG_NEE) {
                    Ray bg = spawn_ray(hit.si, wi_render);
                    SampledSpectrum radiance = scene.evaluate_infinite_light_radiance(bg.d, wl);
                    if (rp.integrator == INTEG_PT) {
                        contribution += throughput * ms.f * radiance / ms.pdf;  // pt_renderer.rs:80-81
                    } else {
                        LightSampler ls = scene.light_sampler(mc, wl);
                        float light_pdf = scene.pdf_infinite_light_sample(ls, bg.d);
                        float w = balance_heuristic(ms.pdf, light_pdf);
                        float tf = 1.0f / ms.pdf;
                        contribution += throughput * ms.f * radiance * tf * w;  // mis_renderer.rs:228-229
                    }
                }
                break;
            }
            float tf = 1.0f / ms.pdf;
            SampledSpectrum next_emissive = SampledSpectrum::zero();
            SampledSpectrum le2;
            if (evaluate_emissive(mc, next, wl, &le2)) next_emissive = ms.f * le2 * tf;
            SampledSpectrum modifier = ms.f * tf;
            // calculate_bsdf_contribution
            if (rp.integrator == INTEG_PT) {
                contribution += throughput * next_emissive;
            } else if (rp.integrator == INTEG_NEE) {
                if (ms.is_specular()) contribution += throughput * next_emissive;
            } else {
                if (ms.is_specular()) contribution += throughput * next_emissive;
                else if (ms.is_sampled) {
                    LightSampler ls = scene.light_sampler(mc, wl);
                    float pdf_light = scene.pdf_light_sample(ls, hit.si.position, next);
                    float w = balance_heuristic(ms.pdf, pdf_light);
                    contribution += throughput * next_emissive * w;
                }
            }
            throughput *= modifier;
            hit = next;
            // apply_russian_roulette (base_renderer.rs:76-92)
            float p_rr = throughput.max_value();
            if (!(p_rr >= 1.0f)) {
                float ur = smp.get_1d();
                if (ur < p_rr) throughput /= p_rr;
                else break;
            }
        }
        return Sensor::sample_to_rgb(T, wl, contribution, rp.exposure);
    }

    // RendererImage::render (renderer.rs:120-134): dynamic per-pixel scheduling over `threads` host threads
    void render(float* out_acc, float* out_srgb, int threads, RayStats* total, uint32_t x0 = 0, uint32_t y0 = 0, uint32_t x1 = 0, uint32_t y1 = 0, uint32_t stride = 1) const {
        if (x1 == 0) x1 = rp.width;
        if (y1 == 0) y1 = rp.height;
        uint32_t rw = (x1 - x0 + stride - 1) / stride, rh = (y1 - y0 + stride - 1) / stride;   // pixels x0, x0 + stride, ... of the window
        std::atomic<uint32_t> next{0};
        std::vector<RayStats> stats(threads);
        auto work = [&](int tid) {
            std::unique_ptr<SamplerBase> smp;
            if (rp.sampler == SAMPLER_SOBOL) smp.reset(new ZSobolSampler(scene.T.sobol, rp.spp, rp.width, rp.height, rp.seed));
            else smp.reset(new RandomSampler(rp.seed));
            for (;;) {
                uint32_t i = next.fetch_add(1);
                if (i >= rw * rh) break;
                uint32_t px = x0 + (i % rw) * stride, py = y0 + (i / rw) * stride;
                Vec3 acc(0, 0, 0);
                for (uint32_t s = 0; s < rp.spp; ++s) acc = acc + trace_path(*smp, px, py, s, &stats[tid]);
                size_t o = ((size_t)py * rp.width + px) * 3;
                if (out_acc) { out_acc[o] = acc.x; out_acc[o + 1] = acc.y; out_acc[o + 2] = acc.z; }
                if (out_srgb) { Vec3 c = rp.integrator == INTEG_NORMAL ? acc / (float)rp.spp : rp.integrator == INTEG_ALBEDO ? Sensor::finalize_no_tone_map(acc, rp.spp) : Sensor::finalize(acc, rp.spp); out_srgb[o] = c.x; out_srgb[o + 1] = c.y; out_srgb[o + 2] = c.z; }
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
        if (total) for (auto& s : stats) { total->closest += s.closest; total->shadow += s.shadow; total->tc.box_tests += s.tc.box_tests; total->tc.tri_tests += s.tc.tri_tests; }
    }
};

}  // namespace orc
