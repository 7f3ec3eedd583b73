// Host scene container and flattener (product code; C++ stand-in for the Rust host side the north star describes).
// Mirrors the construction half of scene::Scene (/root/reference/scene/src/scene.rs:54-76): meshes, textures, materials,
// primitives, environment light; build() bakes world_to_render, builds BLAS/TLAS with the reference's topology and emits
// the POD layout of include/tcpt_flat.h.
#pragma once
#include <string>
#include <vector>

#include "../../include/tcpt_flat.h"
#include "host_bvh.h"
#include "host_math.h"

namespace tcpt {

struct HostTables {
    bool set = false;
    uint32_t sobol[104];
    std::vector<float> cie_x, cie_y, cie_z, d65;  // 470 each
    std::vector<float> rgb2spec;                  // 64 z nodes + 3*64^3*3
    std::vector<float> presets;                   // n x 470 dense metal / glass tables
};

struct HostMesh {
    std::vector<V3> positions, normals, tangents;
    std::vector<V2> uvs;
    std::vector<uint32_t> indices;
    Box bounds;
    BuiltBvh bvh;
    bool built = false;
    bool single = false;  // SingleTriangle primitive: see tcpt_flat_geometry.single
};

struct HostTexture { std::vector<uint8_t> data; uint32_t w, h, channels; };

struct HostPrimitive {
    int kind = 0, geometry = -1, material = -1, env = -1;
    M4 local_to_world;
    std::vector<float> area_list, area_table;
    float area_sum = 0.0f;
    float light_intensity = 0.0f, angle_inner = 0.0f, angle_outer = 0.0f;  // delta lights (kind 3..5)
    tcpt_flat_spectrum light_spectrum{};
};

struct HostEnv {
    float intensity;
    uint32_t w, h;
    std::vector<float> data, marginal, conditional;
    std::vector<uint32_t> marginal_guide, conditional_guide;  // see tcpt_flat_env
    uint32_t guide_h = 0, guide_w = 0;
    float total_weight;
    tcpt_flat_spectrum integrated;
};

// Owns the vectors a tcpt_flat_scene points into.
struct FlatStorage {
    std::vector<tcpt_bvh_node> nodes;
    std::vector<int32_t> tlas_items;
    std::vector<float> tri_verts, positions, normals, uvs, tangents, area_list, area_table, env_floats;
    std::vector<uint32_t> indices, env_guides;
    std::vector<tcpt_flat_geometry> geometries;
    std::vector<tcpt_flat_primitive> primitives;
    std::vector<tcpt_flat_material> materials;
    std::vector<tcpt_flat_texture> textures;
    std::vector<uint8_t> texture_bytes;
    std::vector<int32_t> light_list;
    std::vector<tcpt_flat_env> envs;
    tcpt_flat_scene view{};
};

// guide table of a CDF search (tcpt_flat_env): out[j] = #{i : cdf[i] <= j/G}, j = 0..G (G a power of two: j/G is exact in f32)
void build_cdf_guide(const float* cdf, uint32_t n, uint32_t G, uint32_t* out);

class HostScene {
   public:
    HostTables tables;
    std::vector<HostMesh> meshes;
    std::vector<HostTexture> textures;
    std::vector<tcpt_flat_material> materials;
    std::vector<HostPrimitive> primitives;
    std::vector<HostEnv> envs;
    BuiltBvh tlas;
    std::vector<int> tlas_prims;
    std::vector<int> geom_flat;       // mesh index -> index into the flat geometry table of the last build (-1: not referenced)
    bool use_binned_builder = false;  // soups only (outside topology-parity scope)
    std::string error;

    void clear();
    int add_mesh(const float* pos, const float* nrm, const float* uv, int nverts, const uint32_t* idx, int ntris);
    int set_tangent_source(int mesh, const uint32_t* tri, int ntris);  // multi-model OBJ files: see host_obj.h
    int add_single_triangle(const float pos[9], const float nrm[9], const float uv[6]);
    int add_texture(const uint8_t* data, uint32_t w, uint32_t h, uint32_t channels);
    int add_material(const tcpt_material_desc& d);
    int add_primitive(int geometry, int material, const float l2w[16]);
    int add_env_light(float intensity, const float* rgb, uint32_t w, uint32_t h, const float l2w[16]);
    int add_delta_light(int kind, float intensity, const tcpt_spectrum_param& spectrum, float angle_inner, float angle_outer, const float l2w[16]);
    // Scene::build; fills `out`
    int build(const float cam_pos[3], FlatStorage& out);

    bool rgb_to_coeffs(const float rgb[3], bool gamma_encoded, float coeffs[3], int32_t index[4]) const;
    int dump_bvh(int which, uint32_t* out, int max_nodes) const;
    static int dump_built(const BuiltBvh& b, uint32_t* out, int max_nodes);

   private:
    tcpt_flat_spectrum resolve_spectrum(const tcpt_spectrum_param& p) const;
    tcpt_flat_spectrum illuminant_from_rgb(const float rgb[3]) const;
};

}  // namespace tcpt
