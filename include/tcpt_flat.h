/* tcpt_flat.h — the flattened, GPU-friendly scene layout (POD, #[repr(C)]-mirrorable).
 *
 * This is the lower of the two boundaries: a host (the reference's Rust `scene` crate through a `flatten.rs`, or the C++
 * builder in csrc/host_scene.cpp which stands in for it here) produces one tcpt_flat_scene and hands it to
 * tcpt_upload_flat_scene(); everything device-side reads only these arrays.
 *
 * BVH nodes (TLAS and all BLAS, concatenated) are 32-byte records fetched as two 16-byte loads:
 *     lo = {min.x, min.y, min.z, bits(a)}     hi = {max.x, max.y, max.z, bits(b)}
 *     inner: b == 0, a = index of the SECOND child (first child = this + 1)         [scene/src/bvh.rs:259-266 second_offset]
 *     leaf : b = item_count (> 0), a = first item slot                               [scene/src/bvh.rs:271-286]
 * Node indices are relative to the BVH's node_base and enumerate inner/leaf nodes in the reference's pre-order (item records
 * removed), so "later leaf in DFS order" == larger node index; item slots enumerate leaf items in the same order.
 * BLAS items are stored pre-gathered: tri_verts[3*slot+k] = {p_k.xyz, bits(k==0 ? triangle_index : k==1 ? degenerate_flag : 0)}.
 */
#ifndef TCPT_FLAT_H
#define TCPT_FLAT_H
#include <stdint.h>

#include "tcpt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TCPT_MAX_LIGHTS 16
#define TCPT_TRAVERSAL_STACK 96

typedef struct { float lo[4]; float hi[4]; } tcpt_bvh_node;

typedef struct {
    uint32_t node_base, node_count;   /* into bvh_nodes */
    uint32_t slot_base, tri_count;    /* into tri_verts (x3) */
    uint32_t vertex_base;             /* into positions/normals/uvs */
    uint32_t index_base;              /* into indices (x3 per triangle) */
    uint32_t tangent_base;            /* into tangents (per triangle), valid when has_uv */
    uint32_t has_uv;
} tcpt_flat_geometry;

typedef struct {
    float l2r[12];     /* local_to_render, 3 rows x 4 columns stored column major 4x(x,y,z): c0.xyz c1.xyz c2.xyz c3.xyz */
    float r2l[12];     /* inverse (glam Mat4::inverse), same storage */
    int32_t geometry;  /* -1 for the environment light */
    int32_t material;
    int32_t kind;      /* 0 mesh, 1 emissive mesh, 2 environment light */
    int32_t light_index; /* position in light_list or -1 */
    uint32_t identity; /* l2r is exactly the identity */
    uint32_t area_base; /* emissive: into area_list / area_table (tri_count entries) */
    float area_sum;
    int32_t env;
} tcpt_flat_primitive;

typedef struct { int32_t kind; float c[3]; float scale; int32_t texture; } tcpt_flat_spectrum; /* kind: 0 const(c[0]) 1 sigmoid(c) 2 sigmoid*scale*D65 3 D65 4 texture */
typedef struct { int32_t is_texture; float value; int32_t texture; int32_t gamma_corrected; } tcpt_flat_float;
typedef struct {
    int32_t type;
    tcpt_flat_spectrum color, coat_tint;
    tcpt_flat_float intensity, roughness, metallic, ior, coat_ior, coat_roughness, coat_thickness;
    int32_t normal_texture, normal_flip_y;
    float eta;
    int32_t thin_surface;
} tcpt_flat_material;

typedef struct { uint64_t offset; uint32_t width, height, channels, pad; } tcpt_flat_texture; /* offset into texture_bytes */

typedef struct {
    float intensity;
    uint32_t width, height;
    uint64_t data_offset;        /* into env_floats: h*w*3 */
    uint64_t marginal_offset;    /* into env_floats: h */
    uint64_t conditional_offset; /* into env_floats: h*w */
    float total_weight;
    tcpt_flat_spectrum integrated;
    int32_t primitive;
} tcpt_flat_env;

typedef struct {
    const tcpt_bvh_node* bvh_nodes; uint64_t n_bvh_nodes;
    uint32_t tlas_node_count;               /* TLAS occupies bvh_nodes[0 .. tlas_node_count) */
    const int32_t* tlas_items; uint32_t n_tlas_items; /* primitive index per TLAS item slot */
    const float* tri_verts; uint64_t n_tri_slots;     /* 12 floats per slot */
    const float* positions; const float* normals; const float* uvs; uint64_t n_vertices; /* 3,3,2 floats per vertex */
    const uint32_t* indices; uint64_t n_triangles;
    const float* tangents;                  /* 3 floats per triangle (meshes with UVs) */
    const tcpt_flat_geometry* geometries; uint32_t n_geometries;
    const tcpt_flat_primitive* primitives; uint32_t n_primitives;
    const tcpt_flat_material* materials; uint32_t n_materials;
    const tcpt_flat_texture* textures; uint32_t n_textures;
    const uint8_t* texture_bytes; uint64_t n_texture_bytes;
    const float* area_list; const float* area_table; uint64_t n_area;
    const int32_t* light_list; uint32_t n_lights;   /* primitive indices, LightSamplerFactory order (light_sampler.rs:168-187) */
    const tcpt_flat_env* envs; uint32_t n_envs;
    const float* env_floats; uint64_t n_env_floats;
    uint32_t max_bvh_depth;                 /* TLAS depth + deepest BLAS depth, must be < TCPT_TRAVERSAL_STACK */
} tcpt_flat_scene;

/* copies everything to the device owned by ctx; replaces any previously uploaded scene */
int tcpt_upload_flat_scene(tcpt_ctx* ctx, const tcpt_flat_scene* scene);

#ifdef __cplusplus
}
#endif
#endif
