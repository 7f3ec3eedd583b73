// ORACLE (test infrastructure, NOT product code).  C entry points over the CPU restatement in oracle/*.h so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it through ctypes.
// The product library (toy_cpu_pathtracing_b200/csrc) never links or calls anything in this directory.
// PARITY UNPINNED below the Sobol known-answer vectors: the Rust reference cannot be built here.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "orender.h"

using namespace orc;

extern "C" {

typedef struct { int32_t kind; float value[3]; int32_t texture; } orc_spectrum_param;  // kind: 0 const,1 rgb albedo (sRGB gamma),2 rgb albedo (linear),3 D65,4 texture
typedef struct { int32_t kind; float value; int32_t texture; int32_t gamma_corrected; } orc_float_param;
typedef struct { int32_t texture; int32_t flip_y; } orc_normal_param;
typedef struct {
    int32_t type;
    orc_spectrum_param color;
    orc_float_param intensity;
    orc_normal_param normal;
    float eta;
    int32_t thin_surface;
    orc_float_param roughness, metallic, ior, coat_ior, coat_roughness, coat_thickness;
    orc_spectrum_param coat_tint;
} orc_material_desc;

typedef struct {
    uint32_t width, height, spp, seed, max_depth;
    int32_t integrator, sampler;
    float exposure, fov_deg;
    float cam_pos[3], cam_dir[3], cam_up[3];
    int32_t threads;
    uint32_t x0, y0, x1, y1;  // pixel window (0,0,0,0 = full frame)
    uint32_t stride;          // > 1: only every stride-th pixel of the window in x and in y is rendered (a sparse sample of the WHOLE frame for bench.py)
} orc_render_params;

typedef struct {
    uint64_t closest_rays, shadow_rays, paths, box_tests, tri_tests;
    double seconds;
} orc_stats;

struct OrcScene {
    Scene scene;
    std::vector<float> rgb2spec;
    bool tables_set = false;
};

void* orc_scene_new() { return new OrcScene(); }
void orc_scene_free(void* h) { delete (OrcScene*)h; }

// std_tables = contents of data/std_tables.bin; rgb2spec = contents of data/srgb_table.bin (64 z nodes + table)
int orc_set_tables(void* h, const void* std_tables, size_t std_len, const float* rgb2spec, size_t rgb2spec_len) {
    OrcScene* s = (OrcScene*)h;
    const size_t base_len = 8 + 104 * 4 + 4 * N_DENSE * 4;
    const bool v1 = std_len == base_len && !std::memcmp(std_tables, "TCPTSTD1", 8);
    const bool v2 = std_len >= base_len + 4 && !std::memcmp(std_tables, "TCPTSTD2", 8);
    if (!v1 && !v2) return -1;
    if (rgb2spec_len != 64 + (size_t)3 * 64 * 64 * 64 * 3) return -2;
    Tables& T = s->scene.T;
    const uint8_t* p = (const uint8_t*)std_tables + 8;
    std::memcpy(T.sobol, p, 104 * 4); p += 104 * 4;
    std::memcpy(T.cie_x, p, N_DENSE * 4); p += N_DENSE * 4;
    std::memcpy(T.cie_y, p, N_DENSE * 4); p += N_DENSE * 4;
    std::memcpy(T.cie_z, p, N_DENSE * 4); p += N_DENSE * 4;
    std::memcpy(T.d65, p, N_DENSE * 4); p += N_DENSE * 4;
    T.presets.clear();
    if (v2) {
        uint32_t n; std::memcpy(&n, p, 4); p += 4;
        if (std_len != base_len + 4 + (size_t)n * N_DENSE * 4) return -1;
        T.presets.resize((size_t)n * N_DENSE);
        std::memcpy(T.presets.data(), p, T.presets.size() * 4);
    }
    s->rgb2spec.assign(rgb2spec, rgb2spec + rgb2spec_len);
    std::memcpy(T.z_nodes, s->rgb2spec.data(), 64 * 4);
    T.rgb2spec = s->rgb2spec.data() + 64;
    srgb_matrices(&T.rgb_to_xyz, &T.xyz_to_rgb);
    s->tables_set = true;
    return 0;
}

int orc_add_mesh(void* h, const float* pos, const float* nrm, const float* uv, int nverts, const uint32_t* idx, int ntris) {
    OrcScene* s = (OrcScene*)h;
    Mesh m;
    m.positions.resize(nverts); m.normals.resize(nverts);
    for (int i = 0; i < nverts; ++i) { m.positions[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]); m.normals[i] = Vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]); }
    if (uv) { m.uvs.resize(nverts); for (int i = 0; i < nverts; ++i) { m.uvs[i].x = uv[2 * i]; m.uvs[i].y = uv[2 * i + 1]; } }
    m.indices.assign(idx, idx + (size_t)ntris * 3);
    m.finalize();
    s->scene.meshes.push_back(std::move(m));
    return (int)s->scene.meshes.size() - 1;
}

// multi-model OBJ files: the reference's tangent loop (triangle_mesh.rs:181-226) runs over every triangle gathered so far after each model
// and pushes again, so tangents[j] of a later triangle j is the tangent of an earlier one: tri[j] names it (oracle/obj_oracle.py)
int orc_set_tangent_source(void* h, int geometry, const uint32_t* tri, int n) {
    OrcScene* s = (OrcScene*)h;
    Mesh& m = s->scene.meshes[geometry];
    if (m.tangents.empty()) return 0;
    std::vector<Vec3> moved(n);
    for (int t = 0; t < n; ++t) moved[t] = m.tangents[tri[t]];
    m.tangents.swap(moved);
    return 0;
}

// geometry of CreatePrimitiveDesc::SingleTrianglePrimitive (single_triangle.rs:24-41)
int orc_add_single_triangle(void* h, const float pos[9], const float nrm[9], const float uv[6]) {
    OrcScene* s = (OrcScene*)h;
    Mesh m;
    m.single = true;
    for (int i = 0; i < 3; ++i) {
        m.positions.push_back(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
        m.normals.push_back(Vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        Vec2 t; t.x = uv[2 * i]; t.y = uv[2 * i + 1];
        m.uvs.push_back(t);
    }
    m.indices = {0, 1, 2};
    m.finalize();
    s->scene.meshes.push_back(std::move(m));
    return (int)s->scene.meshes.size() - 1;
}

int orc_add_texture(void* h, const uint8_t* data, uint32_t w, uint32_t hgt, uint32_t channels) {
    OrcScene* s = (OrcScene*)h;
    Texture t; t.w = w; t.h = hgt; t.channels = channels;
    t.data.assign(data, data + (size_t)w * hgt * channels);
    s->scene.textures.push_back(std::move(t));
    return (int)s->scene.textures.size() - 1;
}

static SpectrumParam conv_spec(const Tables& T, const orc_spectrum_param& p) {
    SpectrumParam r;
    switch (p.kind) {
        case 0: r.spectrum = make_constant_spectrum(p.value[0]); break;
        case 1: r.spectrum = make_rgb_albedo(T, Vec3(p.value[0], p.value[1], p.value[2]), true); break;
        case 2: r.spectrum = make_rgb_albedo(T, Vec3(p.value[0], p.value[1], p.value[2]), false); break;
        case 3: r.spectrum.kind = SPEC_D65; break;
        case 4: r.is_texture = true; r.texture = p.texture; break;
        case 5: r.spectrum.kind = SPEC_PRESET; r.spectrum.table = p.texture; break;
    }
    return r;
}
static FloatParam conv_float(const orc_float_param& p) { FloatParam r; r.is_texture = p.kind == 1; r.value = p.value; r.texture = p.texture; r.gamma_corrected = p.gamma_corrected != 0; return r; }

int orc_add_material(void* h, const orc_material_desc* d) {
    OrcScene* s = (OrcScene*)h;
    if (!s->tables_set) return -1;
    const Tables& T = s->scene.T;
    Material m;
    m.type = d->type;
    m.color = conv_spec(T, d->color);
    m.intensity = conv_float(d->intensity);
    m.normal.texture = d->normal.texture; m.normal.flip_y = d->normal.flip_y != 0;
    m.eta = d->eta; m.thin_surface = d->thin_surface != 0;
    m.roughness = conv_float(d->roughness); m.metallic = conv_float(d->metallic); m.ior = conv_float(d->ior);
    m.coat_ior = conv_float(d->coat_ior); m.coat_roughness = conv_float(d->coat_roughness); m.coat_thickness = conv_float(d->coat_thickness);
    m.coat_tint = conv_spec(T, d->coat_tint);
    s->scene.materials.push_back(m);
    return (int)s->scene.materials.size() - 1;
}

// reads back the sigmoid-polynomial coefficients a constant RGB spectrum parameter resolved to (table-index parity checks)
int orc_rgb_to_coeffs(void* h, const float rgb[3], int gamma_encoded, float coeffs[3], int32_t index[4]) {
    OrcScene* s = (OrcScene*)h;
    Rgb2SpecIndex ix;
    if (!rgb_to_coeffs(s->scene.T, Vec3(rgb[0], rgb[1], rgb[2]), gamma_encoded != 0, coeffs, &ix)) return -1;
    index[0] = ix.m; index[1] = ix.zi; index[2] = ix.yi; index[3] = ix.xi;
    return 0;
}

static Mat4 mat_from(const float m[16]) { Mat4 r; std::memcpy(r.c, m, 64); return r; }  // column major

int orc_add_primitive(void* h, int geometry, int material, const float local_to_world[16]) {
    OrcScene* s = (OrcScene*)h;
    Primitive p;
    p.geometry = geometry; p.material = material;
    p.local_to_world = mat_from(local_to_world);
    p.kind = s->scene.materials[material].type == MAT_EMISSIVE ? PRIM_EMISSIVE_MESH : PRIM_MESH;
    if (p.kind == PRIM_EMISSIVE_MESH) s->scene.init_emissive(p);
    s->scene.primitives.push_back(std::move(p));
    return (int)s->scene.primitives.size() - 1;
}

int orc_add_env_light(void* h, float intensity, const float* rgb, uint32_t w, uint32_t hgt, const float local_to_world[16]) {
    OrcScene* s = (OrcScene*)h;
    EnvLight e; e.intensity = intensity; e.w = w; e.h = hgt;
    e.data.assign(rgb, rgb + (size_t)w * hgt * 3);
    s->scene.init_env(e);
    s->scene.envs.push_back(std::move(e));
    Primitive p; p.kind = PRIM_ENV_LIGHT; p.env = (int)s->scene.envs.size() - 1; p.local_to_world = mat_from(local_to_world);
    s->scene.primitives.push_back(std::move(p));
    return (int)s->scene.primitives.size() - 1;
}

// CreatePrimitiveDesc::{PointLight,SpotLight,DirectionalLight}Primitive (primitive/repository.rs:108-134); kind = PRIM_POINT_LIGHT ...
int orc_add_delta_light(void* h, int kind, float intensity, const orc_spectrum_param* spectrum, float angle_inner, float angle_outer, const float local_to_world[16]) {
    OrcScene* s = (OrcScene*)h;
    if (kind < PRIM_POINT_LIGHT || kind > PRIM_DIRECTIONAL_LIGHT || !s->tables_set) return -1;
    Primitive p; p.kind = kind; p.light_intensity = intensity; p.angle_inner = angle_inner; p.angle_outer = angle_outer;
    p.light_spectrum = conv_spec(s->scene.T, *spectrum).spectrum;
    p.local_to_world = mat_from(local_to_world);
    s->scene.primitives.push_back(std::move(p));
    return (int)s->scene.primitives.size() - 1;
}

// faithful = reference cost model (exhaustive traversal is always used; this adds per-call inverses/attribute work);
// literal_build = O(N^2) SAH sweep exactly as the reference, else the prefix/suffix sweep (bit-identical result)
// bench.py's "optimised CPU" figure: ordered, t-shrinking traversal in the TLAS and every BLAS, cached instance inverses (see Bvh::shrink).
// Call after orc_build.
void orc_set_optimised(void* h, int on) {
    OrcScene* s = (OrcScene*)h;
    s->scene.tlas.shrink = on != 0;
    for (auto& m : s->scene.meshes) m.bvh.shrink = on != 0;
    if (on) s->scene.faithful = false;
}
void orc_set_modes(void* h, int faithful, int literal_build) { OrcScene* s = (OrcScene*)h; s->scene.faithful = faithful != 0; s->scene.literal_build = literal_build != 0; }

double orc_build(void* h, const float cam_pos[3]) {
    OrcScene* s = (OrcScene*)h;
    auto t0 = std::chrono::steady_clock::now();
    s->scene.build(Vec3(cam_pos[0], cam_pos[1], cam_pos[2]));
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static Camera make_camera(const orc_render_params* p) {
    Camera c; c.fov = p->fov_deg; c.width = p->width; c.height = p->height;
    c.set_look_to(Vec3(p->cam_pos[0], p->cam_pos[1], p->cam_pos[2]), Vec3(p->cam_dir[0], p->cam_dir[1], p->cam_dir[2]), Vec3(p->cam_up[0], p->cam_up[1], p->cam_up[2]));
    return c;
}
static RenderParams make_rp(const orc_render_params* p) { return RenderParams{p->width, p->height, p->spp, p->seed, p->max_depth, p->integrator, p->sampler, p->exposure}; }

int orc_render(void* h, const orc_render_params* p, float* out_acc, float* out_srgb, orc_stats* stats) {
    OrcScene* s = (OrcScene*)h;
    Camera cam = make_camera(p);
    PathTracer pt(s->scene, cam, make_rp(p));
    RayStats rs;
    int threads = p->threads > 0 ? p->threads : (int)std::thread::hardware_concurrency();
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t stride = p->stride > 1 ? p->stride : 1;
    pt.render(out_acc, out_srgb, threads, &rs, p->x0, p->y0, p->x1, p->y1, stride);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        uint32_t rw = ((p->x1 ? p->x1 : p->width) - p->x0 + stride - 1) / stride, rh = ((p->y1 ? p->y1 : p->height) - p->y0 + stride - 1) / stride;
        stats->closest_rays = rs.closest; stats->shadow_rays = rs.shadow; stats->paths = (uint64_t)rw * rh * p->spp;
        stats->box_tests = rs.tc.box_tests; stats->tri_tests = rs.tc.tri_tests; stats->seconds = sec;
    }
    return 0;
}

// per-sample probe: RGB contribution of individual (pixel, sample) paths, in the order given
int orc_path_samples(void* h, const orc_render_params* p, const uint32_t* pixels_xy, const uint32_t* sample_indices, int n, float* out_rgb) {
    OrcScene* s = (OrcScene*)h;
    Camera cam = make_camera(p);
    // every (pixel, sample) path is a pure function of its coordinates: the list is cut into blocks handed to worker threads
    // (p->threads, 0 = all cores), each with its own sampler and tracer state; the values do not depend on the thread count
    int threads = p->threads > 0 ? p->threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (n < 2048) threads = 1;
    std::atomic<int> next{0};
    auto work = [&]() {
        PathTracer pt(s->scene, cam, make_rp(p));
        std::unique_ptr<SamplerBase> smp;
        if (p->sampler == SAMPLER_SOBOL) smp.reset(new ZSobolSampler(s->scene.T.sobol, p->spp, p->width, p->height, p->seed));
        else smp.reset(new RandomSampler(p->seed));
        for (;;) {
            const int b = next.fetch_add(1024);
            if (b >= n) break;
            const int e = b + 1024 < n ? b + 1024 : n;
            for (int i = b; i < e; ++i) {
                Vec3 c = pt.trace_path(*smp, pixels_xy[2 * i], pixels_xy[2 * i + 1], sample_indices[i], nullptr);
                out_rgb[3 * i] = c.x; out_rgb[3 * i + 1] = c.y; out_rgb[3 * i + 2] = c.z;
            }
        }
    };
    if (threads == 1) work();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    return 0;
}

// record the rays a render issues (closest-hit and shadow) for traversal parity tests; returns counts through n_closest/n_shadow
int orc_record_rays(void* h, const orc_render_params* p, float* closest_od, int max_closest, int* n_closest, float* shadow_odt, int max_shadow, int* n_shadow) {
    OrcScene* s = (OrcScene*)h;
    Camera cam = make_camera(p);
    PathTracer pt(s->scene, cam, make_rp(p));
    std::vector<Ray> cr, sr; std::vector<float> st;
    pt.probe.closest_rays = &cr; pt.probe.shadow_rays = &sr; pt.probe.shadow_tmax = &st;
    std::unique_ptr<SamplerBase> smp;
    if (p->sampler == SAMPLER_SOBOL) smp.reset(new ZSobolSampler(s->scene.T.sobol, p->spp, p->width, p->height, p->seed));
    else smp.reset(new RandomSampler(p->seed));
    uint32_t x1 = p->x1 ? p->x1 : p->width, y1 = p->y1 ? p->y1 : p->height;
    for (uint32_t y = p->y0; y < y1; ++y) for (uint32_t x = p->x0; x < x1; ++x) for (uint32_t k = 0; k < p->spp; ++k) pt.trace_path(*smp, x, y, k, nullptr);
    *n_closest = (int)std::min<size_t>(cr.size(), max_closest);
    *n_shadow = (int)std::min<size_t>(sr.size(), max_shadow);
    for (int i = 0; i < *n_closest; ++i) { float* o = closest_od + 6 * i; o[0] = cr[i].o.x; o[1] = cr[i].o.y; o[2] = cr[i].o.z; o[3] = cr[i].d.x; o[4] = cr[i].d.y; o[5] = cr[i].d.z; }
    for (int i = 0; i < *n_shadow; ++i) { float* o = shadow_odt + 7 * i; o[0] = sr[i].o.x; o[1] = sr[i].o.y; o[2] = sr[i].o.z; o[3] = sr[i].d.x; o[4] = sr[i].d.y; o[5] = sr[i].d.z; o[6] = st[i]; }
    return 0;
}

// closest-hit / any-hit of explicit rays.  rays = n x {o[3], d[3], tmax}.  out_hit = n x {prim, tri, t bits, b0 bits, b1 bits, b2 bits} (prim = -1 on miss)
static void trace_range(OrcScene* s, const float* rays, int b, int e, int any_hit, int32_t* out_hit, RayStats& st) {
    for (int i = b; i < e; ++i) {
        const float* r = rays + 7 * (size_t)i;
        Ray ray{Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5])};
        int32_t* o = out_hit + 6 * (size_t)i;
        if (any_hit) {
            o[0] = s->scene.intersect_p(ray, r[6], &st) ? 1 : 0;
            o[1] = o[2] = o[3] = o[4] = o[5] = 0;
        } else {
            Intersection is;
            if (s->scene.intersect(ray, r[6], &is, &st)) {
                o[0] = is.primitive; o[1] = (int32_t)is.tri;
                std::memcpy(&o[2], &is.t_hit, 4); std::memcpy(&o[3], &is.bary[0], 4); std::memcpy(&o[4], &is.bary[1], 4); std::memcpy(&o[5], &is.bary[2], 4);
            } else { o[0] = -1; o[1] = o[2] = o[3] = o[4] = o[5] = 0; }
        }
    }
}
int orc_trace(void* h, const float* rays, int n, int any_hit, int32_t* out_hit, uint64_t* box_tests, uint64_t* tri_tests) {
    OrcScene* s = (OrcScene*)h;
    RayStats st;
    trace_range(s, rays, 0, n, any_hit, out_hit, st);
    if (box_tests) *box_tests = st.tc.box_tests;
    if (tri_tests) *tri_tests = st.tc.tri_tests;
    return 0;
}
// the same over `threads` host threads (0 = all cores), blocks of 4096 rays handed out dynamically; returns the wall-clock seconds of
// the tracing alone (bench.py's CPU figure for the triangle-soup workloads)
double orc_trace_mt(void* h, const float* rays, int n, int any_hit, int32_t* out_hit, int threads, uint64_t* box_tests, uint64_t* tri_tests) {
    OrcScene* s = (OrcScene*)h;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    std::vector<RayStats> st(threads);
    std::atomic<int> next{0};
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
        for (;;) {
            const int b = next.fetch_add(4096);
            if (b >= n) break;
            trace_range(s, rays, b, b + 4096 < n ? b + 4096 : n, any_hit, out_hit, st[t]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& t : pool) t.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t nb = 0, nt = 0;
    for (auto& x : st) { nb += x.tc.box_tests; nt += x.tc.tri_tests; }
    if (box_tests) *box_tests = nb;
    if (tri_tests) *tri_tests = nt;
    return sec;
}

// the CDF search of EnvironmentLight::sample_infinite_light alone (environment_light.rs:218-223)
int orc_cdf_search(const float* cdf, int n, const float* u, int m, uint32_t* out) {
    for (int i = 0; i < m; ++i) out[i] = (uint32_t)Scene::sample_from_cdf(cdf, (size_t)n, u[i]);
    return 0;
}

// Sobol known-answer probe: start_pixel_sample(p, i); get_1d; get_2d; get_1d  -> 4 floats + the 3 sample indices
int orc_sobol_probe(void* h, uint32_t spp, uint32_t w, uint32_t hgt, uint32_t seed, uint32_t px, uint32_t py, uint32_t sample_index, float out_vals[4], uint64_t out_index[3], uint32_t* morton) {
    OrcScene* s = (OrcScene*)h;
    ZSobolSampler z(s->scene.T.sobol, spp, w, hgt, seed);
    z.start_pixel_sample(px, py, sample_index);
    *morton = z.morton_index;
    out_index[0] = z.get_sample_index(); out_vals[0] = z.get_1d();
    out_index[1] = z.get_sample_index(); Vec2 v = z.get_2d(); out_vals[1] = v.x; out_vals[2] = v.y;
    out_index[2] = z.get_sample_index(); out_vals[3] = z.get_1d();
    return 0;
}

// generic stream probe: n_dims draws following the pattern in `kinds` (1 = get_1d, 2 = get_2d)
int orc_sampler_stream(void* h, int sampler, uint32_t spp, uint32_t w, uint32_t hgt, uint32_t seed, uint32_t px, uint32_t py, uint32_t sample_index, const int32_t* kinds, int n, float* out) {
    OrcScene* s = (OrcScene*)h;
    std::unique_ptr<SamplerBase> smp;
    if (sampler == SAMPLER_SOBOL) smp.reset(new ZSobolSampler(s->scene.T.sobol, spp, w, hgt, seed));
    else smp.reset(new RandomSampler(seed));
    smp->start_pixel_sample(px, py, sample_index);
    int k = 0;
    for (int i = 0; i < n; ++i) {
        if (kinds[i] == 1) out[k++] = smp->get_1d();
        else { Vec2 v = smp->get_2d(); out[k++] = v.x; out[k++] = v.y; }
    }
    return k;
}

// BVH topology dump.  which = -1: TLAS, else geometry index.  Returns node count; fills up to max_nodes records of 8 x u32:
// {kind, value, min.x, min.y, min.z, max.x, max.y, max.z} (floats as bit patterns; items carry only kind/value)
int orc_get_bvh(void* h, int which, uint32_t* out, int max_nodes) {
    OrcScene* s = (OrcScene*)h;
    const Bvh& b = which < 0 ? s->scene.tlas : s->scene.meshes[which].bvh;
    int n = (int)b.nodes.size();
    for (int i = 0; i < n && i < max_nodes; ++i) {
        const FlatNode& nd = b.nodes[i];
        uint32_t* o = out + 8 * i;
        o[0] = nd.kind; o[1] = nd.value;
        float f[6] = {nd.bounds.mn.x, nd.bounds.mn.y, nd.bounds.mn.z, nd.bounds.mx.x, nd.bounds.mx.y, nd.bounds.mx.z};
        if (nd.kind == NODE_ITEM) std::memset(f, 0, sizeof f);
        std::memcpy(o + 2, f, 24);
    }
    return n;
}

// standalone build of a BVH over explicit boxes (topology tests of the builder alone)
int orc_build_bvh_boxes(const float* boxes, int n, int literal, uint32_t* out, int max_nodes) {
    std::vector<Bounds> ib(n);
    for (int i = 0; i < n; ++i) ib[i] = Bounds{Vec3(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]), Vec3(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5])};
    Bvh b;
    b.build(ib, literal != 0);
    int cnt = (int)b.nodes.size();
    for (int i = 0; i < cnt && i < max_nodes; ++i) {
        const FlatNode& nd = b.nodes[i];
        uint32_t* o = out + 8 * i;
        o[0] = nd.kind; o[1] = nd.value;
        float f[6] = {nd.bounds.mn.x, nd.bounds.mn.y, nd.bounds.mn.z, nd.bounds.mx.x, nd.bounds.mx.y, nd.bounds.mx.z};
        if (nd.kind == NODE_ITEM) std::memset(f, 0, sizeof f);
        std::memcpy(o + 2, f, 24);
    }
    return cnt;
}

// mesh tangents after finalize (host-side parity of load-time tangent generation)
int orc_get_mesh_tangents(void* h, int geometry, float* out, int max_tris) {
    OrcScene* s = (OrcScene*)h;
    const Mesh& m = s->scene.meshes[geometry];
    int n = (int)m.tangents.size();
    for (int i = 0; i < n && i < max_tris; ++i) { out[3 * i] = m.tangents[i].x; out[3 * i + 1] = m.tangents[i].y; out[3 * i + 2] = m.tangents[i].z; }
    return n;
}

}  // extern "C"
