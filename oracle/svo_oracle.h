/*
 * oracle/svo_oracle.h -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's SVO construction path
 * (AdamYuan/SparseVoxelOctree: Voxelizer + OctreeBuilder).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library, and there only as the checker / the CPU baseline.  The
 * product (sparsevoxeloctree_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or fixtures for this path
 *  and cannot run in this image (no Vulkan ICD).  What pins this oracle:
 *  - PINNED to the reference's own code for everything its shaders define: the checked-in
 *    SPIR-V binaries (shader/include/spirv/<name>.u32) are EXECUTED by oracle/spirv_interp.py and
 *    their outputs committed (tests/golden/spirv_*.npz): orc_build equals the four compute
 *    shaders word for word; the geometry and fragment stages equal voxelizer.geom / .frag
 *    (tests/test_spirv_golden.py), plus hand-derived KATs (tests/test_oracle_kat.py).
 *  - UNPINNED for the stages the reference does not define: the fixed-function rasterizer
 *    (coverage, snapping, depth interpolation) and the texture unit (texture() filtering / LOD /
 *    sRGB decode, and the linear blits that build the mip chains) live in the Vulkan driver.  The
 *    arithmetic used here for them is the "pinned arithmetic" stated in DESIGN.md section 3.
 */
#ifndef SVO_ORACLE_H
#define SVO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One draw per material: mirrors Scene::DrawCmd (src/Scene.hpp:28-34, Scene.cpp:157-170). */
typedef struct {
	uint32_t first_index, index_count;
	uint32_t texture_id;   /* 0xffffffff = untextured (voxelizer.frag:35) */
	uint32_t albedo_rgba8; /* packUnorm4x8(vec4(albedo,0)), R in bits 0-7 (Scene.cpp:164) */
} orc_draw;

/* An unpacked voxel fragment (voxelizer.frag:17-25 voxel position + colour & 0xffffff). */
typedef struct {
	uint32_t x, y, z, rgb;
} orc_frag;

enum { ORC_CENTER = 0, ORC_CONSERVATIVE_EXACT = 1, ORC_CONSERVATIVE_DILATE = 2 };

/* voxelizer.frag:40-42 packing and octree_tag_node.comp:38-39 unpacking (levels <= 12). */
void orc_pack_fragment(uint32_t x, uint32_t y, uint32_t z, uint32_t colour, uint32_t out[2]);
void orc_unpack_fragment(const uint32_t in[2], uint32_t *x, uint32_t *y, uint32_t *z, uint32_t *rgb);

/* OctreeBuilder.cpp:42-45: the reference's octree buffer size guess, in 32-bit words. */
uint32_t orc_octree_entry_num(uint32_t fragment_count, uint32_t level);

/*
 * Voxelize (voxelizer.vert + voxelizer.geom + rasterizer state at Voxelizer.cpp:112-129 +
 * voxelizer.frag).  Positions are read with a byte stride (20 for the reference's Vertex,
 * 12 for tight float3).  shard_lo/hi (may be NULL) is a half-open voxel box; fragments
 * outside it are dropped (used by the octant-sharding tests).  Fragments are written in
 * draw order, triangle order, then row-major pixel order when nthreads == 1; with
 * nthreads > 1 the order is arbitrary (atomic append, like voxelizer.frag:37).
 * Returns the fragment count (also when it exceeds cap; only cap entries are written).
 */
int64_t orc_voxelize(const void *positions, uint32_t pos_stride_bytes, const uint32_t *indices,
                     const orc_draw *draws, uint32_t n_draws, uint32_t level, int mode,
                     const uint32_t *shard_lo, const uint32_t *shard_hi, orc_frag *out, int64_t cap,
                     int nthreads);

/*
 * Textured materials (voxelizer.frag:27-36).  A texture is its base level as stbi_load(..., 4) returns it
 * (Scene.cpp:247): RGBA8, sRGB-encoded colour, row 0 first, tight rows.  orc_texset_create builds the mip
 * chains the way Scene::load_textures does (linear blits, Scene.cpp:262-296).  texcoords: 2 floats per
 * vertex with a byte stride (the reference's Vertex: positions + 12, stride 20).
 */
typedef struct {
	const uint8_t *rgba8;
	uint32_t width, height;
} orc_texture;
typedef struct orc_texset orc_texset;
orc_texset *orc_texset_create(const orc_texture *tex, uint32_t n);
void orc_texset_destroy(orc_texset *s);
/* level data of one texture (debug / tests); returns the level count or -1 */
int orc_texset_level(const orc_texset *s, uint32_t tex, uint32_t level, uint32_t *w, uint32_t *h, const uint8_t **data);
int64_t orc_voxelize_textured(const void *positions, uint32_t pos_stride_bytes, const void *texcoords, uint32_t uv_stride_bytes,
                              const uint32_t *indices, const orc_draw *draws, uint32_t n_draws, const orc_texset *texset,
                              uint32_t level, int mode, const uint32_t *shard_lo, const uint32_t *shard_hi, orc_frag *out,
                              int64_t cap, int nthreads);
/* colour of pixel (px,py) of one textured triangle: 0xff000000 | rgb, or 0 when discarded; lod_out = {hi, lo, delta*256} */
uint32_t orc_debug_sample(const orc_texset *s, uint32_t tex, const float *p0, const float *p1, const float *p2, const float *uv0,
                          const float *uv1, const float *uv2, uint32_t level, int32_t px, int32_t py, uint32_t lod_out[3]);

/* the interpolated gTexcoord and the texture() value (before the alpha test / packing) of one pixel of a textured triangle */
void orc_debug_texture_fetch(const orc_texset *s, uint32_t tex, const float *p0, const float *p1, const float *p2, const float *uv0,
                             const float *uv1, const float *uv2, uint32_t level, int32_t px, int32_t py, double uv_out[2],
                             float rgba_out[4]);
/* voxelizer.frag:28-30,35,42 on a given sample value: 0xff000000 | (packUnorm4x8(x) & 0xffffff), or 0 when discarded */
uint32_t orc_debug_shade(const float rgba[4]);

/* Debug views for the SPIR-V cross-checks: geometry-stage outputs {axis, gAABB[4], gDepthRange[2]} + snapped window
 * coordinates of one triangle; and its covered pixels with the pinned fp64 depth (before voxelizer.frag). */
void orc_debug_tri_setup(const float *p0, const float *p1, const float *p2, uint32_t level, uint32_t out_axis_aabb_zr[7],
                         int32_t out_xy_snapped[6]);
int64_t orc_debug_raster_pixels(const float *p0, const float *p1, const float *p2, uint32_t level, int mode, int32_t *out_px,
                                int32_t *out_py, double *out_z, int64_t cap);

void orc_debug_dilate(const float *p0, const float *p1, const float *p2, uint32_t level, float out[9]);

/*
 * The OctreeBuilder level loop (OctreeBuilder.cpp:142-210 driving octree_init_node /
 * octree_tag_node / octree_alloc_node / octree_modify_arg).  Fragments are tagged in
 * input order and nodes allocated in window order when nthreads == 1.
 * Returns the octree range in bytes ((counter+1)*32, OctreeBuilder.cpp:212-214), or -1
 * if the build would write past cap_words (the reference would write out of bounds).
 */
int64_t orc_build(const orc_frag *frags, int64_t n_frags, uint32_t level, uint32_t *words,
                  uint64_t cap_words, int nthreads);

/*
 * Canonicalise a node buffer by Morton depth-first traversal from the root block
 * (node word layout by use in octree.glsl:87-110, octree_tag_node.comp:26,48,59).
 * Emits every non-empty node as (depth, morton at that depth, word with the child
 * pointer masked out of internal nodes).  Returns the node count (also when > cap),
 * or a negative code when the buffer violates the layout:
 *  -1 pointer not a multiple of 8 / out of range / zero, -2 leaf above the last level,
 *  -3 internal node at the last level, -4 non-empty word without bit 31.
 */
int64_t orc_canonicalise(const uint32_t *words, uint64_t n_words, uint32_t level, uint8_t *out_depth,
                         uint64_t *out_morton, uint32_t *out_word, int64_t cap);

/*
 * Octree_RayMarchLeaf (octree.glsl:179-340): the reference's primary-ray traversal of the node buffer, restated
 * for verification of built trees (the octree occupies [1,2]^3).  Returns hit; outputs as the shader's.
 */
int orc_raymarch_leaf(const uint32_t *octree, const float o[3], const float d[3], float o_pos[3], float o_colour[3],
                      float o_normal[3], uint32_t *o_iter);

/* Morton code of a voxel: child slot = x | y<<1 | z<<2 per level, MSB first (tag_node.comp:24-25). */
uint64_t orc_morton(uint32_t x, uint32_t y, uint32_t z, uint32_t level);

#ifdef __cplusplus
}
#endif
#endif
