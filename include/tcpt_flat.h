/* tcpt_flat.h — the flattened, GPU-friendly scene layout (POD, #[repr(C)]-mirrorable).
 *
 * This is the lower of the two boundaries: a host (the reference's Rust `scene` crate through a `flatten.rs`, or the C++
 * builder in csrc/host_scene.cpp which stands in for it here) produces one tcpt_flat_scene and hands it to
 * tcpt_upload_flat_scene(); everything device-side reads only these arrays.
 *
 * BVH nodes (TLAS and all BLAS, concatenated) are 128-byte 4-WIDE records fetched as 16-byte rows.  A record holds up to four
 * children of a 4-wide COLLAPSE of the reference's binary tree (scene/src/bvh.rs:92-295): starting from a binary node's two
 * children, the largest-area inner child is replaced by its own two children until four are held (DFS order is kept):
 *     row 0 = lo.x of children 0..3   row 1 = hi.x   row 2 = lo.y   row 3 = hi.y   row 4 = lo.z   row 5 = hi.z
 *     row 6 = entry of children 0..3  row 7 = item count of children 0..3 (0 for inner / absent children; informational)
 *     entry, inner child : index of the child's own record in bvh_nodes (ABSOLUTE, < 2^31)
 *     entry, leaf child  : 0x80000000 | (item_count - 1) << 27 | first item slot (ABSOLUTE into tri_verts / tlas_items, < 2^27 - 1),
 *                          item_count <= 16                                                     [scene/src/bvh.rs:271-286]
 *     entry, absent child: 0xffffffff, box = (+inf, -inf)
 * so that the near / far planes of all four children along one axis are two adjacent 16-byte rows (picked by the sign of the ray
 * direction) and an entry is what the traversal stack holds.  Every reference LEAF is kept with its own box bits and its item
 * order; a reference leaf of more than 16 items becomes a record of up to four item ranges that all carry the leaf's box (nested
 * for more than 64).  Reference inner boxes that end up interior to a record are not stored: an inner box is the exact min / max
 * merge of the boxes below it and the slab test is monotone in the box, so a ray that passes a leaf box passes every ancestor
 * box -- the leaves reached, hence the candidates, are the reference's (csrc/dtraverse.cuh).  Record 0 of the array is
 * the TLAS root; tcpt_flat_geometry.node_base is the BLAS root record.  A BVH whose root is a leaf is one record with one child.
 * Records are emitted in pre-order; item slots enumerate leaf items in the reference's leaf order, so "later leaf in DFS order" ==
 * larger first-slot (tcpt_get_bvh() returns the reference's own flattened order for comparison).
 * BLAS items are stored pre-gathered: tri_verts[3*slot+k] = {p_k.xyz, bits(w_k)} with w_0 = triangle index,
 * w_1 = degenerate flag (|e1 x e2|^2 == 0, math/src/ray.rs:50-57), w_2 = first slot of the leaf holding this slot.
 * TLAS items: tlas_items[2*slot] = primitive index, tlas_items[2*slot+1] = first slot of the leaf holding this slot.
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

typedef struct { float q[32]; } tcpt_bvh_node; /* 8 rows x 4 children, see above */

typedef struct {
    uint32_t node_base, node_count;   /* into bvh_nodes: the BLAS root record, number of records */
    uint32_t slot_base, tri_count;    /* into tri_verts (x3): first item slot */
    uint32_t vertex_base;             /* into positions/normals/uvs */
    uint32_t index_base;              /* into indices (x3 per triangle) */
    uint32_t tangent_base;            /* into tangents (per triangle), valid when has_uv */
    uint32_t has_uv;
    uint32_t single;                  /* SingleTriangle primitive (primitive/impls/single_triangle.rs): no box tests, tangent used as stored */
} tcpt_flat_geometry;

typedef struct { int32_t kind; float c[3]; float scale; int32_t texture; } tcpt_flat_spectrum; /* kind: 0 const(c[0]) 1 sigmoid(c) 2 sigmoid*scale*D65 3 D65 4 texture 5 dense preset table (texture = preset id) */

typedef struct {
    float l2r[12];     /* local_to_render, 3 rows x 4 columns stored column major 4x(x,y,z): c0.xyz c1.xyz c2.xyz c3.xyz */
    float r2l[12];     /* inverse (glam Mat4::inverse), same storage */
    int32_t geometry;  /* -1 for the environment light */
    int32_t material;
    int32_t kind;      /* 0 mesh, 1 emissive mesh, 2 environment light, 3 point light, 4 spot light, 5 directional light */
    int32_t light_index; /* position in light_list or -1 */
    uint32_t identity; /* l2r is exactly the identity */
    uint32_t area_base; /* emissive: into area_list / area_table (tri_count entries) */
    float area_sum;
    int32_t env;
    /* delta lights (kind 3..5): intensity, spectrum, spot cone angles, DirectionalLight::preprocess area = pi r^2 of the scene's bounding sphere */
    float light_intensity, angle_inner, angle_outer, dir_area;
    tcpt_flat_spectrum light_spectrum;
} tcpt_flat_primitive;

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
    /* guide tables for the CDF searches (an acceleration structure, not reference data): for a CDF of n entries and G = the
     * power of two >= n, guide[j] = #{i : cdf[i] <= j/G}, j = 0..G, so the search for u only looks at cdf[guide[j] .. guide[j+1])
     * with j = floor(u*G).  marginal: guide_h + 1 entries; conditional: height rows of guide_w + 1 entries. */
    uint32_t guide_h, guide_w;
    uint64_t marginal_guide_offset, conditional_guide_offset; /* into env_guides */
} tcpt_flat_env;

typedef struct {
    const tcpt_bvh_node* bvh_nodes; uint64_t n_bvh_nodes;
    uint32_t tlas_node_count;               /* TLAS occupies bvh_nodes[0 .. tlas_node_count) */
    const int32_t* tlas_items; uint32_t n_tlas_items; /* 2 x int32 per TLAS item slot: {primitive, leaf_first_slot} */
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
    const uint32_t* env_guides; uint64_t n_env_guides;
    uint32_t max_bvh_depth;                 /* worst-case height of the traversal stack (siblings waiting along the deepest TLAS path + 1 + the same in the deepest BLAS), must be < TCPT_TRAVERSAL_STACK */
} tcpt_flat_scene;

/* copies everything to the device owned by ctx; replaces any previously uploaded scene */
int tcpt_upload_flat_scene(tcpt_ctx* ctx, const tcpt_flat_scene* scene);

#ifdef __cplusplus
}
#endif
#endif
