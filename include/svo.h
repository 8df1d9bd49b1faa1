/*
 * include/svo.h -- the drop-in boundary of svo-b200: a C ABI (plain pointers and sizes, no
 * C++ / torch types) over the hand-written sm_100a CUDA implementation of the reference's SVO
 * construction path (AdamYuan/SparseVoxelOctree Voxelizer + OctreeBuilder).
 *
 * The reference has no FFI for this path; its boundary is two C++ classes with one caller
 * (src/LoaderThread.cpp:53-54,64,74) and one consumer (src/Octree.cpp:24-27).  Each entry point
 * below cites the reference interface it replaces (paths relative to the reference tree).  The C++
 * mirror classes with the reference's method names live in sparsevoxeloctree_b200/host/, the
 * Python (ctypes) mirror in sparsevoxeloctree_b200/api.py, and the binding a reference maintainer
 * would add is shown in INTEGRATION.md.
 *
 * Conventions: caller owns handles, the library owns device memory; no exceptions cross the ABI;
 * every call returns SVO_OK (0) or a negative svo_status, with a message in svo_last_error()
 * (thread-local).  A handle is used by one thread at a time.  All device work is enqueued on the
 * `stream` argument (a cudaStream_t passed as void*; NULL = the legacy default stream).
 * There is no CPU fallback anywhere behind this ABI: without a CUDA device calls fail.
 */
#ifndef SVO_B200_H
#define SVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SVO_API
#else
#define SVO_API __attribute__((visibility("default")))
#endif

typedef enum svo_status {
	SVO_OK = 0,
	SVO_ERR_INVALID_ARGUMENT = -1,
	SVO_ERR_CUDA = -2,           /* a CUDA runtime call failed; svo_last_error() has the cudaError string */
	SVO_ERR_UNSUPPORTED = -3,    /* e.g. IPC in a build without it */
	SVO_ERR_CAPACITY = -4,       /* > 2^32-1 fragments, or an octree of >= 2^30 words (30-bit node pointers) */
	SVO_ERR_NOT_READY = -5       /* result queried before the producing call */
} svo_status;

/* Rasterization rule.  The stock reference is always conservative: Mode A when
 * VK_EXT_conservative_rasterization exists (src/Voxelizer.cpp:92-99,120-127).  SVO_CENTER is
 * BASELINE.json's "non-conservative raster". */
typedef enum svo_raster_mode {
	SVO_CENTER = 0,            /* centre sample + top-left rule on the undilated triangle */
	SVO_CONSERVATIVE_EXACT = 1, /* every pixel square touching the triangle (exact 2-D SAT) = reference Mode A */
	SVO_CONSERVATIVE_DILATE = 2 /* reference Mode B (no VK_EXT_conservative_rasterization): the triangle dilated by
	                               shader/voxelizer_conservative.geom:46-87, then centre sampled, depth clipped */
} svo_raster_mode;

/* One draw per material: Scene::DrawCmd (src/Scene.hpp:28-34; filled at src/Scene.cpp:157-170,
 * consumed by Scene::CmdDraw src/Scene.cpp:450-463). */
typedef struct svo_draw {
	uint32_t first_index, index_count;
	uint32_t texture_id;   /* 0xffffffff = untextured (shader/voxelizer.frag:35) */
	uint32_t albedo_rgba8; /* glm::packUnorm4x8(vec4(albedo, 0)), R in bits 0-7 (src/Scene.cpp:164) */
} svo_draw;

/* The mesh hand-off that Scene keeps private (src/Scene.hpp:26,35-40): the vertex buffer
 * (struct Vertex {vec3 pos; vec2 uv}, stride 20, src/Scene.cpp:16-19; a tight float3 array with
 * stride 12 is accepted too), the u32 index buffer and the draw list.  Positions must already be
 * normalised to [-1,1]^3 (src/Scene.cpp:90-99). */
/* One texture as Scene::load_textures reads it (src/Scene.cpp:245-262): the stbi_load(..., 4) image, RGBA8 with
 * sRGB-encoded colour (VK_FORMAT_R8G8B8A8_SRGB), row 0 first, tight rows; always a HOST pointer.  The library
 * builds the mip chain on the device like CmdGenerateMipmap2D (linear blits, src/Scene.cpp:290-295). */
typedef struct svo_texture {
	const uint8_t *rgba8;
	uint32_t width, height; /* 1..16384 */
} svo_texture;

typedef struct svo_mesh {
	const void *positions;          /* first vertex position; HOST pointer unless on_device != 0 */
	uint32_t position_stride_bytes; /* >= 12, multiple of 4 */
	uint32_t on_device;             /* 1: positions/texcoords/indices are device pointers (borrowed, must outlive the scene) */
	const uint32_t *indices;
	uint64_t n_vertices, n_indices;
	const svo_draw *draws; /* always a host pointer */
	uint32_t n_draws;
	/* textured materials (shader/voxelizer.frag:27-36); all zero / NULL for untextured scenes */
	uint32_t n_textures;
	const svo_texture *textures;    /* host array; svo_draw.texture_id indexes it */
	const void *texcoords;          /* first vertex's uv (2 floats); for the reference's Vertex: positions + 12, same stride */
	uint32_t texcoord_stride_bytes; /* >= 8, multiple of 4 */
} svo_mesh;

/* A power-of-two sub-cube of the grid (octant sharding): the cube of side
 * 2^(level - shard_level) voxels whose origin is cube_index * side.  Fragments are emitted in
 * cube-local coordinates; the builder then builds the (level - shard_level)-deep subtree. */
typedef struct svo_shard {
	uint32_t shard_level;   /* 0 = whole grid */
	uint32_t cube_index[3]; /* each < 2^shard_level */
} svo_shard;

typedef struct svo_scene svo_scene;
typedef struct svo_voxelizer svo_voxelizer;
typedef struct svo_builder svo_builder;

/* ---- library ------------------------------------------------------------------------------- */
SVO_API const char *svo_last_error(void);
SVO_API const char *svo_version(void);
SVO_API int svo_device_count(void); /* < 0: CUDA unavailable */
SVO_API uint64_t svo_launch_count(void); /* kernels launched by this library since it was loaded */

/* ---- Scene (input side) -------------------------------------------------------------------
 * Replaces the GPU-buffer half of Scene::Create / load_buffers_and_draw_cmd
 * (src/Scene.cpp:145-223,387-417): uploads (or borrows) vertex/index buffers and keeps the draw
 * list.  Host pointers may be released when the call returns if they are pageable; pinned host
 * memory must stay valid until the stream has run the copy.
 * Positions must be normalised to [-1,1]^3 (what src/Scene.cpp:90-99 guarantees): a triangle with a non-finite
 * vertex or a coordinate beyond +-2 is skipped, not clipped.  Every index must be < n_vertices
 * (SVO_ERR_INVALID_ARGUMENT otherwise; checked on the device, so borrowed device buffers are covered too).
 * svo_*_destroy() release device memory stream-ordered on the stream of the handle's last call. */
SVO_API int svo_scene_create(const svo_mesh *mesh, int device, void *stream, svo_scene **out);
SVO_API void svo_scene_destroy(svo_scene *scene);
SVO_API uint64_t svo_scene_triangle_count(const svo_scene *scene);
/* One mip level of a texture as the scene holds it on the device (the images Scene::load_textures leaves after
 * CmdGenerateMipmap2D, src/Scene.cpp:290-295): *d_texels = DEVICE pointer to width*height RGBA8 texels.
 * Returns the texture's level count (ImageBase::QueryMipLevel) or a negative svo_status. */
SVO_API int svo_scene_texture_level(const svo_scene *scene, uint32_t texture, uint32_t level, uint32_t *width, uint32_t *height,
                                    const uint32_t **d_texels);

/* ---- Voxelizer ----------------------------------------------------------------------------
 * svo_voxelizer_create = Voxelizer::Create (src/Voxelizer.hpp:43-45, src/Voxelizer.cpp:5-26):
 * synchronous; runs the count pass (src/Voxelizer.cpp:134-165) and allocates the exact-size
 * fragment list, so svo_voxelizer_fragment_count() is final when it returns.
 * shard may be NULL (whole grid). */
SVO_API int svo_voxelizer_create(svo_scene *scene, uint32_t level, int mode, const svo_shard *shard, void *stream,
                                 svo_voxelizer **out);
/* A "pre-voxelized" source: a voxelizer whose fragment list is supplied by the caller (n 64-bit fragments,
 * morton << 24 | rgb, any order; host pointer unless on_device).  Lets OctreeBuilder run on fragment lists
 * produced elsewhere -- e.g. the reference's own voxelizer output (src/Voxelizer.hpp:52) re-packed -- and is
 * what the SPIR-V golden tests use.  svo_voxelizer_voxelize() then just restores the list. */
SVO_API int svo_voxelizer_create_from_fragments(int device, uint32_t level, const uint64_t *fragments, uint64_t n, int on_device,
                                                void *stream, svo_voxelizer **out);
/* A window of the whole grid (half-open voxel box): only fragments inside it are produced, in GLOBAL coordinates at
 * the full level.  Used by the multi-GPU slab split (each rank builds the part of the tree under its octants). */
SVO_API int svo_voxelizer_create_windowed(svo_scene *scene, uint32_t level, int mode, const uint32_t window_lo[3],
                                          const uint32_t window_hi[3], void *stream, svo_voxelizer **out);
SVO_API void svo_voxelizer_destroy(svo_voxelizer *vox);
/* Voxelizer::CmdVoxelize (src/Voxelizer.hpp:49, src/Voxelizer.cpp:167-179): enqueues the fragment
 * emission on the stream (the reference records it into a command buffer).  On the brick path only the small
 * triangles are emitted here (see svo_voxelizer_fragments). */
SVO_API int svo_voxelizer_voxelize(svo_voxelizer *vox, void *stream);
SVO_API uint32_t svo_voxelizer_level(const svo_voxelizer *vox);            /* Voxelizer::GetLevel */
SVO_API uint32_t svo_voxelizer_resolution(const svo_voxelizer *vox);       /* Voxelizer::GetVoxelResolution */
SVO_API uint64_t svo_voxelizer_fragment_count(const svo_voxelizer *vox);   /* Voxelizer::GetVoxelFragmentCount */
/* Voxelizer::GetVoxelFragmentList (src/Voxelizer.hpp:52): DEVICE pointer to fragment_count 64-bit
 * fragments, (morton(x,y,z) << 24) | rgb, morton slot order x | y<<1 | z<<2 per level
 * (shader/octree_tag_node.comp:24-25), in triangle order.
 * The list is valid between svo_voxelizer_voxelize() and the next svo_builder_build()/prepare() on this voxelizer:
 * the build CONSUMES it (sorts it in place and reuses the storage), unlike the reference's CmdBuild, which only reads
 * it.  After a build, svo_voxelizer_export_reference_fragments() and a second build return SVO_ERR_NOT_READY until
 * svo_voxelizer_voxelize() has run again.
 * Brick path (svo_debug_set_build_path): svo_voxelizer_voxelize() emits only the small triangles' fragments -- the
 * builder bins the large triangles and never needs theirs.  This call (and svo_voxelizer_export_reference_fragments)
 * then completes the list first: it enqueues the large triangles' emission on the voxelizer's last stream and waits
 * for it, so the returned list is whole and may be read on any stream. */
SVO_API const uint64_t *svo_voxelizer_fragments(const svo_voxelizer *vox);
/* The reference's own fragment packing (shader/voxelizer.frag:40-42, uvec2 per fragment, levels
 * <= 12): converts the fragment list into d_out (DEVICE, fragment_count * 8 bytes) on the stream. */
SVO_API int svo_voxelizer_export_reference_fragments(const svo_voxelizer *vox, uint32_t *d_out, void *stream);

/* ---- OctreeBuilder ------------------------------------------------------------------------
 * svo_builder_create = OctreeBuilder::Create (src/OctreeBuilder.hpp:36-37, src/OctreeBuilder.cpp:8-22). */
SVO_API int svo_builder_create(svo_voxelizer *vox, void *stream, svo_builder **out);
SVO_API void svo_builder_destroy(svo_builder *b);
/* OctreeBuilder::CmdBuild (src/OctreeBuilder.hpp:41, src/OctreeBuilder.cpp:142-210): enqueues sort,
 * de-duplication and the level build.  The octree buffer is sized exactly, which costs one
 * internal stream synchronisation mid-build (the reference guesses the size, OctreeBuilder.cpp:42-45). */
SVO_API int svo_builder_build(svo_builder *b, void *stream);
/* The same build in two phases, for callers that place the node words themselves (multi-GPU stitch over NVLink,
 * externally allocated / Vulkan-exported memory):
 *   svo_builder_prepare  : sort + reduce + levels + the size read-back (one stream sync); afterwards
 *                          svo_builder_octree_range_bytes / leaf_count / level_counts are valid;
 *   svo_builder_emit_to  : writes the node words into d_dst (DEVICE memory of this or, through P2P / CUDA IPC, of
 *                          another GPU).  skip_root = 0: the whole tree, child pointers = word index + bias.
 *                          skip_root = 1: blocks 1.. are written from d_dst[0] on with pointers already valid for a
 *                          buffer in which d_dst sits at word offset pointer_bias_words; the 8 root words are kept
 *                          aside (svo_builder_root_words) so that the caller can merge several subtrees' roots.
 *                          skip_root = 2: the same one level deeper -- the root block AND the depth-1 blocks (1 + N_1
 *                          blocks, N_1 = level_counts[1] <= 8) are kept aside (svo_builder_top_words; the depth-1 blocks
 *                          follow the root's non-empty slots in order), blocks 1 + N_1 .. are written from d_dst[0] on.
 *                          Parts of the grid cut at depth-2 cell borders can then be built separately (several per GPU,
 *                          emitted while the next one is being built) and their top blocks summed.
 * This is the fused "emit + transfer": the kernel that produces the words stores them across NVLink. */
SVO_API int svo_builder_prepare(svo_builder *b, void *stream);
SVO_API int svo_builder_emit_to(svo_builder *b, uint32_t *d_dst, uint32_t pointer_bias_words, int skip_root, void *stream);
/* Compact gather (multi-GPU, brick path).  A tree built from wall-sized triangles is mostly "flat" bricks -- 16 identical
 * leaf blocks that their 16-byte record describes completely -- and pointer blocks that follow from the records and two
 * ranks per brick.  Instead of storing the finished node words across NVLink (svo_builder_emit_to: 11 bytes per leaf
 * cross the link), a prepared build can send
 *   - the node words of the windows above depth L-2 and the leaf blocks of the rasterized (not flat) bricks, to their
 *     final places in d_dst, and
 *   - 32 bytes per brick (record, rank of its first depth L-1 / depth L-2 node) to d_tables
 *     (svo_builder_compact_bytes(b) bytes, 16-byte aligned; this or a peer GPU's memory),
 * and the GPU that owns the destination buffer writes everything else itself at local HBM speed:
 * svo_expand_compact(device, d_tables, plan, d_dst) with the same d_dst (as seen from that GPU) and the four plan words
 * svo_builder_emit_compact_to returned (host values: ship them with the sizes).  Together the two calls write exactly what
 * svo_builder_emit_to(b, d_dst, pointer_bias_words, skip_root) writes; the blocks kept aside (svo_builder_root_words /
 * svo_builder_top_words) are the same.  svo_builder_compact_bytes is 0 for a build that took the fragment-sort path
 * (no compact form: use svo_builder_emit_to).  The caller orders svo_expand_compact after the arrival of the tables
 * (stream order on one device; a collective or an IPC event between processes).
 * Replaces nothing in the reference (single device); the gather is north_star's "subtrees gathered to GPU 0".
 * svo_builder_push_tables does the table part alone (and returns the plan words), svo_builder_emit_compact_to with
 * d_tables = NULL the rest: a host that sends the tables first lets the owner of the buffer expand them while the
 * other stores are still crossing the link. */
SVO_API uint64_t svo_builder_compact_bytes(const svo_builder *b);
SVO_API int svo_builder_push_tables(svo_builder *b, uint32_t pointer_bias_words, int skip_root, void *d_tables, uint64_t plan[4], void *stream);
SVO_API int svo_builder_emit_compact_to(svo_builder *b, uint32_t *d_dst, uint32_t pointer_bias_words, int skip_root, void *d_tables,
                                        uint64_t plan[4], void *stream);
SVO_API int svo_expand_compact(int device, const void *d_tables, const uint64_t plan[4], uint32_t *d_dst, void *stream);
SVO_API int svo_builder_root_words(svo_builder *b, uint32_t out[8], void *stream);
SVO_API int svo_builder_top_words(svo_builder *b, uint32_t out[72], uint32_t *n_blocks, void *stream);
SVO_API uint32_t svo_builder_level(const svo_builder *b); /* OctreeBuilder::GetLevel */
/* OctreeBuilder::GetOctreeRange (src/OctreeBuilder.hpp:42, src/OctreeBuilder.cpp:212-214): bytes. */
SVO_API uint64_t svo_builder_octree_range_bytes(const svo_builder *b);
/* OctreeBuilder::GetOctree (src/OctreeBuilder.hpp:43): DEVICE pointer to the node words in the
 * reference layout (shader/octree.glsl:87-110): 0 empty; 0x80000000|child_block_word_index
 * internal; 0xC0000000|min(n,63)<<24|BGR leaf; root block at word 0; levels in contiguous windows
 * top-down, Morton order inside a level. */
SVO_API const uint32_t *svo_builder_octree(const svo_builder *b);
SVO_API uint64_t svo_builder_leaf_count(const svo_builder *b);
/* node counts per depth: out[d] = non-empty nodes at depth d, d = 0..level (out[0] = 1 if any) */
SVO_API int svo_builder_level_counts(const svo_builder *b, uint64_t *out, uint32_t n_out);
/* Multi-GPU stitch support: copy the whole node buffer to d_dst + dst_word_offset while
 * adding `base_words` to every internal node's child pointer (leaves untouched), so that the subtree can
 * live at word offset base_words of a larger buffer whose root block points at it -- see
 * sparsevoxeloctree_b200/sharded.py.  d_dst may be peer (P2P / CUDA-IPC mapped) memory: the rebase is
 * fused with the NVLink transfer. */
SVO_API int svo_builder_rebase_copy(const svo_builder *b, uint32_t *d_dst, uint64_t dst_word_offset,
                                    uint32_t base_words, void *stream);

/* ---- timing (LoaderThread.cpp:57-97 writes 4 GPU timestamps around CmdVoxelize / CmdBuild) -- */
enum { SVO_PHASE_RASTER = 0,      /* fragment emission (CmdVoxelize) */
       SVO_PHASE_SORT_HIST = 1,   /* digit histograms of all passes + bin scan */
       SVO_PHASE_SORT_PASSES = 2, /* the onesweep passes (sort_passes launches of one kernel) */
       SVO_PHASE_REDUCE = 3,      /* de-duplicate + colour reduce */
       SVO_PHASE_LEVELS = 4,      /* bottom-up parent compaction, one launch per level */
       SVO_PHASE_EMIT = 5,        /* size read-back + node word emission */
       SVO_PHASE_COUNT = 6 };
/* Milliseconds of the last voxelize (RASTER) / build (others), from cudaEvents on the call's stream.
 * Synchronises on the recorded events.  sort_passes receives the number of radix passes. */
SVO_API int svo_voxelizer_last_ms(svo_voxelizer *vox, float *raster_ms);
SVO_API int svo_builder_last_ms(svo_builder *b, float *phase_ms /*[SVO_PHASE_COUNT]*/, uint32_t *sort_passes);

/* ---- single kernels, exposed so the parity tests can exercise them in isolation ------------- */
/* Stable LSD radix sort of n 64-bit keys on bits [begin_bit, end_bit); d_keys is sorted in place
 * (d_tmp: n keys of scratch). */
SVO_API int svo_sort_u64(uint64_t *d_keys, uint64_t *d_tmp, uint64_t n, uint32_t begin_bit, uint32_t end_bit,
                         int device, void *stream);
/* Test switch: the sort keeps 32-bit look-back words below 2^30 keys and 64-bit ones above; on != 0 forces the
 * 64-bit variant for every size so that the tests reach it without sorting a billion keys. */
SVO_API void svo_debug_force_wide_sort_state(int on);
/* Profiling switch: on != 0 makes every sort record a cudaEvent after each of its kernels (histogram, every radix
 * pass, bucket sort); svo_builder_sort_step_ms then returns the milliseconds between consecutive events of the
 * builder's last sort (out[0..n), n = return value <= cap; < 0: svo_status).  Off by default: no events, no cost. */
SVO_API void svo_debug_profile_passes(int on);
SVO_API int svo_builder_sort_step_ms(svo_builder *b, float *out, uint32_t cap);
/* Build path switch, read when a voxelizer is created.  -1 (default): automatic -- the brick path (large triangles
 * binned to 8^3-voxel bricks and rasterized straight into the three deepest tree levels; only the small triangles'
 * fragments are emitted and sorted) when level - shard_level >= 4 and the large triangles hold at least a fifth of
 * the fragments; 0: never (every fragment is emitted, sorted and reduced); 1: whenever the level allows it.  Both
 * paths produce the same node buffer bit for bit.  svo_builder_build_path: the path the builder's last build took
 * (0 / 1); with 1 the phases SORT_HIST / SORT_PASSES / REDUCE of svo_builder_last_ms hold: small triangles' fragments
 * (sort + reduce) / pair generation + pair sort / the brick kernel. */
SVO_API void svo_debug_set_build_path(int mode);
SVO_API int svo_builder_build_path(const svo_builder *b);
/* Brick path only, after a build (or prepare + emit_to): counts = { (brick, triangle) pairs incl. small records, bricks,
 * leaves of small triangles, bricks that needed pixels (the others are flat: one triangle, one depth voxel) },
 * ms = { k_brick_flat + k_brick_raster, k_brick_ranks, (nothing: the keys are written by k_brick_ranks), k_brick_emit } of the last build (cudaEvents). */
SVO_API int svo_builder_brick_stats(svo_builder *b, uint64_t counts[4], float ms[4]);

/* ---- the consumer side, for verification ----------------------------------------------------
 * Octree_RayMarchLeaf (shader/octree.glsl:179-340, the primary-ray traversal octree_tracer.frag:36 runs on the
 * node buffer), one ray per thread: lets a built tree be checked / rendered without Vulkan.  The octree
 * occupies [1,2]^3 (octree.glsl:53).  All pointers are DEVICE pointers; origins / dirs hold 3 floats per ray
 * (dirs need not be normalised; camera rays: shader/camera.glsl:12-15). */
typedef struct svo_ray_hit {
	float pos[3];    /* o_pos: hit position, pushed just outside the entry face */
	float colour[3]; /* o_color: unpackUnorm4x8(leaf).xyz */
	float normal[3]; /* o_normal: axis-aligned entry-face normal */
	uint32_t hit;    /* the function's return value */
	uint32_t iter;   /* o_iter: traversal iterations (the tracer's "iteration" view) */
} svo_ray_hit;
SVO_API int svo_octree_raymarch_leaf(int device, const uint32_t *d_octree, uint64_t n_rays, const float *d_origins,
                                     const float *d_dirs, svo_ray_hit *d_hits, void *stream);

/* ---- plain device memory helpers for callers without a CUDA binding (tests, ctypes) --------- */
SVO_API int svo_device_malloc(int device, uint64_t bytes, void **out);
SVO_API int svo_device_free(int device, void *ptr);
SVO_API int svo_memcpy_h2d(int device, void *d_dst, const void *h_src, uint64_t bytes, void *stream);
SVO_API int svo_memcpy_d2h(int device, void *h_dst, const void *d_src, uint64_t bytes, void *stream);
SVO_API int svo_memcpy_d2d(int device, void *d_dst, const void *d_src, uint64_t bytes, void *stream);
SVO_API int svo_stream_synchronize(int device, void *stream);

/* ---- CUDA IPC: lets the other ranks (one process per GPU) store their subtrees straight into rank 0's
 * stitched node buffer over NVLink (svo_builder_rebase_copy with a mapped peer pointer).  The buffer
 * must come from svo_device_malloc (cudaMalloc), not from a stream-ordered pool. */
#define SVO_IPC_HANDLE_BYTES 64
SVO_API int svo_ipc_export(int device, void *d_ptr, unsigned char handle[SVO_IPC_HANDLE_BYTES]);
SVO_API int svo_ipc_open(int device, const unsigned char handle[SVO_IPC_HANDLE_BYTES], void **out);
SVO_API int svo_ipc_close(int device, void *d_ptr);

/* ---- hand-off to Vulkan: the node buffer as an external-memory file descriptor -------------------------------
 * What Octree::Update (src/Octree.cpp:22-35) takes from OctreeBuilder::GetOctree() (src/OctreeBuilder.hpp:43) is a
 * myvk::Buffer; a CUDA-built tree reaches the unmodified OctreeTracer / PathTracer through VK_KHR_external_memory_fd.
 * svo_builder_export_fd puts the node words of a prepared (svo_builder_prepare) or built builder into memory
 * allocated with cuMemCreate(CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) -- a prepared builder emits straight into it,
 * no copy -- and returns a file descriptor (cuMemExportToShareableHandle) plus the allocation size (the range rounded
 * up to the allocation granularity: the size vkAllocateMemory must be given).  The Vulkan side imports it with
 * VkImportMemoryFdInfoKHR{handleType = OPAQUE_FD} (INTEGRATION.md has the myvk::BufferBase subclass); the import takes
 * the descriptor over, otherwise the caller closes it.  The memory stays alive until svo_builder_destroy and every
 * importer has released it.  *d_ptr (may be NULL) receives the CUDA address of the exported buffer. */
SVO_API int svo_builder_export_fd(svo_builder *b, int *fd, uint64_t *alloc_size, const uint32_t **d_ptr, void *stream);
/* The importing side restated in CUDA, for verification without a Vulkan driver: maps `size` bytes of an exported
 * descriptor (cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer -- the calls a CUDA consumer of a Vulkan
 * allocation makes; where the driver refuses a CUDA-exported descriptor there, cuMemImportFromShareableHandle).
 * Returns 1 / 2 for the path that mapped it, or a negative svo_status.  The import consumes the descriptor. */
SVO_API int svo_external_memory_import_fd(int device, int fd, uint64_t size, void **import_handle, void **d_ptr);
SVO_API int svo_external_memory_release(int device, void *import_handle);

/* ---- the whole loader sequence on several GPUs of one process ------------------------------------------------
 * svo_build_sharded replaces, for a C/C++ host, the reference's single-device sequence Scene::Create ->
 * Voxelizer::Create -> OctreeBuilder::Create -> CmdVoxelize + CmdBuild (src/LoaderThread.cpp:51-89) by the
 * octant-sharded build of SURVEY.md section 8e: the mesh (host memory) is uploaded to every device, device k
 * voxelizes and builds the subtrees of its octants (x / xy / xyz split for 2 / 4 / 8 devices), and the node words
 * are stored over NVLink peer access into ONE buffer on devices[0], stitched under a shared root -- the buffer
 * Octree::Update would take.  Synchronous.  n_devices = 1, 2, 4 or 8 (a device may be listed more than once);
 * level 14 is built as 8 cube-local level-13 octants.  The stitched tree must stay below 2^30 words. */
typedef struct svo_sharded svo_sharded;
SVO_API int svo_build_sharded(const svo_mesh *mesh, uint32_t level, int mode, const int *devices, uint32_t n_devices,
                              svo_sharded **out);
SVO_API int svo_sharded_rebuild(svo_sharded *sh); /* voxelize + build + stitch again (same scene, same buffers) */
SVO_API const uint32_t *svo_sharded_octree(const svo_sharded *sh);       /* DEVICE pointer on devices[0] */
SVO_API uint64_t svo_sharded_octree_range_bytes(const svo_sharded *sh);
SVO_API uint64_t svo_sharded_leaf_count(const svo_sharded *sh);
SVO_API uint64_t svo_sharded_fragment_count(const svo_sharded *sh);
SVO_API float svo_sharded_last_ms(const svo_sharded *sh);                /* host wall clock of the last build + stitch */
SVO_API void svo_sharded_destroy(svo_sharded *sh);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H */
