// svo_b200.cu -- the C ABI of include/svo.h over the sm_100a kernels (single translation unit).
//
// Host-side orchestration only: it owns device buffers, enqueues the kernels of raster.cuh / sort.cuh /
// build.cuh on the caller's stream and reads back the few scalars that size the next allocation.
// There is no CPU implementation of any phase here: without a CUDA device every entry point fails.
#include "../../include/svo.h"

#include <stdarg.h>

#ifndef SVO_EMU
#include <cuda.h> // driver API types; the entry points are looked up at run time (no link dependency on libcuda)
#include <unistd.h>
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "build.cuh"
#include "brick.cuh"
#include "raster.cuh"
#include "raymarch.cuh"
#include "sort.cuh"

namespace svo {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}
const char *get_error() { return g_err; }

int dev_alloc(void **p, uint64_t bytes, cudaStream_t s) {
	SVO_CUDA_TRY(cudaMallocAsync(p, bytes, s));
	return 0;
}
void dev_free(void *p, cudaStream_t s) {
	if (p) cudaFreeAsync(p, s);
}

#ifndef SVO_EMU
int configure_device_pool(int device) {
	static bool done[64] = {};
	if (device < 0 || device >= 64 || done[device]) return 0;
	cudaMemPool_t pool;
	SVO_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
	uint64_t thr = UINT64_MAX; // keep freed blocks cached: build/destroy cycles do not hit the driver allocator
	SVO_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
	done[device] = true;
	return 0;
}
int sm_count(int device) {
	int n = 0;
	if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 148;
	return n;
}
#else
int configure_device_pool(int) { return 0; }
int sm_count(int) { return 4; }
#endif

struct DeviceGuard {
	int prev = -1;
	bool ok = true;
	explicit DeviceGuard(int dev) {
		if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
		if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
	}
	~DeviceGuard() {
		if (prev >= 0) cudaSetDevice(prev);
	}
};

struct Timer {
	cudaEvent_t a = nullptr, b = nullptr;
	bool recorded = false;
	int init() {
		if (!a) {
			SVO_CUDA_TRY(cudaEventCreate(&a));
			SVO_CUDA_TRY(cudaEventCreate(&b));
		}
		return 0;
	}
	void destroy() {
		if (a) cudaEventDestroy(a);
		if (b) cudaEventDestroy(b);
		a = b = nullptr;
	}
};

} // namespace svo

using namespace svo;

struct svo_scene {
	int device = 0;
	cudaStream_t last_stream = nullptr; // stream of the last call that enqueued work: destroy frees on it
	SceneView view{};
	std::vector<DrawRec> draws;
	DevBuf<unsigned char> pos;
	DevBuf<uint32_t> idx;
	DevBuf<DrawRec> d_draws;
	uint64_t n_vertices = 0;
	// textured materials
	bool textured = false; // some draw samples a texture: the voxelizer runs its TEX kernel variants
	DevBuf<unsigned char> uv;
	DevBuf<uint32_t> texels;
	DevBuf<TexDesc> tex_desc;
	DevBuf<float> tex_decode, tex_enc;
	DevBuf<double> tex_lod;
	std::vector<TexDesc> h_desc;
};

struct svo_voxelizer {
	svo_scene *scene = nullptr;
	int device = 0;
	cudaStream_t last_stream = nullptr;
	uint32_t level = 0;       // full-grid level
	uint32_t key_level = 0;   // level of the emitted (shard-local) keys
	RasterParams rp{};
	uint64_t n_frag = 0, n_frag_small = 0, n_frag_large = 0;
	uint32_t n_large = 0, n_rows = 0;
	DevBuf<uint64_t> tri_off; // n_tri + 1, small-class fragment offsets
	DevBuf<LargeTri> large;
	DevBuf<UvMap> large_uv; // textured scenes: texture-coordinate map of every large triangle
	DevBuf<uint32_t> row_off, row_xy, row_li;
	DevBuf<uint64_t> frags;
	DevBuf<uint64_t> ext_frags; // svo_voxelizer_create_from_fragments: the caller's fragment list (voxelize re-copies it)
	// brick path (brick.cuh): the large triangles are binned to 8^3-voxel bricks instead of being emitted as fragments
	bool brick = false;
	bool large_emitted = false;   // the large triangles' fragments are in frags (brick path: only on request)
	uint64_t n_tile_rows = 0, n_pairs_large = 0;
	DevBuf<uint64_t> tr_base;     // [n_large + 1] first tile row of every large triangle
	DevBuf<uint64_t> pair_off;    // [n_tile_rows + 1] first pair of every tile row
	bool voxelized = false;
	Timer t_raster;
};

struct svo_builder {
	svo_voxelizer *vox = nullptr;
	int device = 0;
	cudaStream_t last_stream = nullptr;
	uint32_t level = 0;
	DevBuf<uint64_t> tmp;      // sort ping-pong partner of the fragment list
	DevBuf<uint32_t> leaf;     // leaf words
	DevBuf<uint32_t> first;    // pooled per-level first-child arrays
	DevBuf<unsigned char> slot;     // pooled per-depth child-slot arrays
	DevBuf<uint64_t> counts;   // device: node count per depth 0..level
	DevBuf<uint64_t> lb_state;
	DevBuf<uint32_t> tickets;
	DevBuf<uint32_t> octree;
	SortScratch sort_scratch;
	ScanScratch scan_scratch;
	DevBuf<uint64_t> rf_cnt01, rf_cnt2, rf_pre01, rf_pre2; // per reduce tile: run counts and their exclusive prefixes
	// brick path (brick.cuh)
	DevBuf<uint64_t> pairs_a, pairs_b, brick_u64, brick_scalars;
	DevBuf<uint32_t> pair_flags, brick_first, small_leaf, brick_u32, brick_temp;
	uint64_t n_pairs = 0, n_small_leaves = 0, n_bricks = 0; // of the last build
	cudaEvent_t ev_brick[5] = {};                            // around k_brick_flat + k_brick_raster, the scans, k_brick_emit
	uint64_t n_slow = 0;                                     // bricks that needed pixels (the others are flat)
	int path = 0;                             // 0: every fragment sorted; 1: bricks
	BrickArgs brick_args{};                   // the arrays of the last brick build
	uint64_t h_counts[MAX_LEVEL + 1] = {};
	cudaStream_t aux = nullptr;   // second stream: work that does not depend on what the caller's stream is busy with
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	uint64_t *h_pinned = nullptr; // 8 words of page-locked host memory: read-backs the host does not block the stream for
	uint64_t range_bytes = 0;
	uint32_t sort_passes = 0;
	bool built = false;     // the node words are in b->octree
	bool prepared = false;  // sort / reduce / levels done, sizes known: ready for an emit
	bool emitted = false;   // svo_builder_emit_to has run after the last prepare (phase times are complete)
	EmitParams ep{};
	DevBuf<uint32_t> root_scratch; // svo_builder_emit_to(skip_root): the top blocks go here
	uint32_t top_blocks = 0;       // how many (1, or 1 + depth-1 nodes)
	// svo_builder_export_fd: the node words in exportable (cuMemCreate) memory
	unsigned long long export_va = 0, export_handle = 0, export_size = 0;
	cudaEvent_t ev[SVO_PHASE_COUNT + 1] = {};
};

// at most min(F, 8^d) nodes at depth d
static uint64_t node_cap(uint64_t F, uint32_t d) {
	const uint64_t cap8 = d >= 11 ? UINT64_MAX : (1ull << (3 * d));
	return F < cap8 ? F : cap8;
}

static int fail(int code, const char *msg) {
	set_error("%s", msg);
	return code;
}

// Scene::load_textures (src/Scene.cpp:225-300): upload the base levels, build the mip chains on the device with
// linear blits, and the lookup tables of texture.cuh.  Synchronous (host staging buffers live on this stack).
// Small page-locked host slots (8 words each) for read-backs the host waits for through an event instead of blocking
// the stream.  One allocation per process (cudaHostAlloc / cudaFreeHost per builder cost more than a small build and
// synchronise the device); slots are handed out and taken back under a mutex.
static std::mutex g_pin_mutex;
static uint64_t *g_pin_base = nullptr;
static std::vector<uint32_t> g_pin_free;
constexpr uint32_t PIN_SLOTS = 4096, PIN_WORDS = 8;
static uint64_t *pinned_slot_acquire() {
	std::lock_guard<std::mutex> lock(g_pin_mutex);
	if (!g_pin_base) {
#ifdef SVO_EMU
		g_pin_base = static_cast<uint64_t *>(calloc((size_t)PIN_SLOTS * PIN_WORDS, sizeof(uint64_t)));
#else
		if (cudaHostAlloc(reinterpret_cast<void **>(&g_pin_base), (size_t)PIN_SLOTS * PIN_WORDS * sizeof(uint64_t), cudaHostAllocPortable) != cudaSuccess) {
			(void)cudaGetLastError();
			g_pin_base = nullptr;
		}
#endif
		if (!g_pin_base) return nullptr;
		for (uint32_t i = PIN_SLOTS; i-- > 0;) g_pin_free.push_back(i);
	}
	if (g_pin_free.empty()) return new (std::nothrow) uint64_t[PIN_WORDS](); // (more handles than slots: pageable memory, still correct)
	const uint32_t i = g_pin_free.back();
	g_pin_free.pop_back();
	return g_pin_base + (size_t)i * PIN_WORDS;
}
static void pinned_slot_release(uint64_t *p) {
	if (!p) return;
	std::lock_guard<std::mutex> lock(g_pin_mutex);
	if (!g_pin_base || p < g_pin_base || p >= g_pin_base + (size_t)PIN_SLOTS * PIN_WORDS) {
		delete[] p;
		return;
	}
	g_pin_free.push_back((uint32_t)((p - g_pin_base) / PIN_WORDS));
}

static double srgb_to_linear(double x) { return x <= 0.04045 ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4); }
static int upload_textures(svo_scene *sc, const svo_mesh *mesh, cudaStream_t s) {
	const uint32_t n = mesh->n_textures;
	std::vector<TexDesc> desc(n);
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const svo_texture &t = mesh->textures[i];
		if (!t.rgba8 || t.width == 0 || t.height == 0 || t.width > TEX_MAX_SIZE || t.height > TEX_MAX_SIZE)
			return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: texture must be 1..16384 texels wide and high");
		TexDesc &d = desc[i];
		d = TexDesc{};
		d.w = t.width, d.h = t.height;
		uint32_t levels = 1; // ImageBase::QueryMipLevel (dep/MyVK/include/myvk/ImageBase.hpp:10-17,49)
		while ((t.width | t.height) >> levels) ++levels;
		d.levels = levels;
		for (uint32_t l = 0; l < levels; ++l) {
			d.off[l] = (uint32_t)total;
			total += (uint64_t)std::max(t.width >> l, 1u) * std::max(t.height >> l, 1u);
		}
		if (total >= (1ull << 32)) return fail(SVO_ERR_CAPACITY, "svo_scene_create: more than 2^32 texels");
	}
	std::vector<float> decode(256), enc(256);
	std::vector<double> lod(128);
	for (int c = 0; c < 256; ++c) {
		decode[c] = (float)srgb_to_linear((double)c / 255.0);
		enc[c] = c ? (float)srgb_to_linear(((double)c - 0.5) / 255.0) : 0.0f;
	}
	for (int k = 0; k < 128; ++k) lod[k] = std::exp2((double)k / 128.0);
	SVO_TRY(sc->texels.alloc(total, s));
	SVO_TRY(sc->tex_desc.alloc(n, s));
	SVO_TRY(sc->tex_decode.alloc(256, s));
	SVO_TRY(sc->tex_enc.alloc(256, s));
	SVO_TRY(sc->tex_lod.alloc(128, s));
	SVO_CUDA_TRY(cudaMemcpyAsync(sc->tex_desc.p, desc.data(), n * sizeof(TexDesc), cudaMemcpyHostToDevice, s));
	SVO_CUDA_TRY(cudaMemcpyAsync(sc->tex_decode.p, decode.data(), 256 * sizeof(float), cudaMemcpyHostToDevice, s));
	SVO_CUDA_TRY(cudaMemcpyAsync(sc->tex_enc.p, enc.data(), 256 * sizeof(float), cudaMemcpyHostToDevice, s));
	SVO_CUDA_TRY(cudaMemcpyAsync(sc->tex_lod.p, lod.data(), 128 * sizeof(double), cudaMemcpyHostToDevice, s));
	TexView tv{sc->texels.p, sc->tex_desc.p, sc->tex_decode.p, sc->tex_enc.p, sc->tex_lod.p, n};
	for (uint32_t i = 0; i < n; ++i) {
		const TexDesc &d = desc[i];
		SVO_CUDA_TRY(cudaMemcpyAsync(sc->texels.p + d.off[0], mesh->textures[i].rgba8, (uint64_t)d.w * d.h * 4, cudaMemcpyHostToDevice, s));
		SVO_LAUNCH_INDEP(std::min<uint32_t>(div_up((uint64_t)d.w * d.h, 256), 1024u), 256, s, k_tex_has_alpha, (const uint32_t *)sc->texels.p,
		                 sc->tex_desc.p, i);
		for (uint32_t l = 1; l < d.levels; ++l) {
			const uint32_t ws = std::max(d.w >> (l - 1), 1u), hs = std::max(d.h >> (l - 1), 1u);
			const uint32_t wd = std::max(d.w >> l, 1u), hd = std::max(d.h >> l, 1u);
			SVO_LAUNCH_INDEP(div_up((uint64_t)wd * hd, 256), 256, s, k_mip_downsample, tv, sc->texels.p, d.off[l - 1], ws, hs, d.off[l], wd, hd);
		}
	}
	SVO_CUDA_TRY(cudaGetLastError());
	// texture coordinates
	if (mesh->on_device) {
		sc->view.uv = (const unsigned char *)mesh->texcoords;
	} else {
		const uint64_t bytes = mesh->n_vertices ? (mesh->n_vertices - 1) * (uint64_t)mesh->texcoord_stride_bytes + 8 : 0;
		SVO_TRY(sc->uv.alloc(bytes, s));
		if (bytes) SVO_CUDA_TRY(cudaMemcpyAsync(sc->uv.p, mesh->texcoords, bytes, cudaMemcpyHostToDevice, s));
		sc->view.uv = sc->uv.p;
	}
	sc->view.uv_stride = mesh->texcoord_stride_bytes;
	sc->view.tex = tv;
	sc->h_desc = desc;
	SVO_CUDA_TRY(cudaStreamSynchronize(s));
	return SVO_OK;
}

// ---- exportable memory (driver API virtual memory management, entry points resolved through the runtime) ----------
#ifndef SVO_EMU
namespace {
template <class F> F driver_fn(const char *name) {
	void *p = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
	return reinterpret_cast<F>(p);
}
struct Vmm {
	CUresult (*granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
	CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
	CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
	CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
	CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
	CUresult (*export_handle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
	CUresult (*import_handle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
	CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
	CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
	CUresult (*address_free)(CUdeviceptr, size_t) = nullptr;
	bool ok = false;
	Vmm() {
		granularity = driver_fn<decltype(granularity)>("cuMemGetAllocationGranularity");
		create = driver_fn<decltype(create)>("cuMemCreate");
		reserve = driver_fn<decltype(reserve)>("cuMemAddressReserve");
		map = driver_fn<decltype(map)>("cuMemMap");
		set_access = driver_fn<decltype(set_access)>("cuMemSetAccess");
		export_handle = driver_fn<decltype(export_handle)>("cuMemExportToShareableHandle");
		import_handle = driver_fn<decltype(import_handle)>("cuMemImportFromShareableHandle");
		unmap = driver_fn<decltype(unmap)>("cuMemUnmap");
		release = driver_fn<decltype(release)>("cuMemRelease");
		address_free = driver_fn<decltype(address_free)>("cuMemAddressFree");
		ok = granularity && create && reserve && map && set_access && export_handle && import_handle && unmap && release && address_free;
	}
};
const Vmm &vmm() {
	static Vmm v;
	return v;
}
CUmemAllocationProp export_prop(int device) {
	CUmemAllocationProp prop{};
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	prop.location.id = device;
	prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
	return prop;
}
} // namespace
static void release_export(svo_builder *b) {
	if (!b->export_va) return;
	const Vmm &v = vmm();
	cudaStreamSynchronize(b->last_stream);
	v.unmap((CUdeviceptr)b->export_va, b->export_size);
	v.release((CUmemGenericAllocationHandle)b->export_handle);
	v.address_free((CUdeviceptr)b->export_va, b->export_size);
	b->export_va = b->export_handle = b->export_size = 0;
}
#else
static void release_export(svo_builder *) {}
#endif

#ifndef SVO_EMU
namespace {
struct ImportedMemory {
	int path = 0; // 1: cudaImportExternalMemory, 2: cuMemImportFromShareableHandle
	cudaExternalMemory_t ext = nullptr;
	CUmemGenericAllocationHandle handle = 0;
	CUdeviceptr va = 0;
	size_t size = 0;
};
} // namespace
#endif
extern "C" {

const char *svo_last_error(void) { return get_error(); }
const char *svo_version(void) {
#ifdef SVO_EMU
	return "svo-b200 0.1 (CPU kernel-logic emulation build: tests only)";
#else
	return "svo-b200 0.1 (sm_100a)";
#endif
}
uint64_t svo_launch_count(void) { return svo::g_launches; }
int svo_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
	return n;
}

// ------------------------------------------------------------------------------------------------------
int svo_scene_create(const svo_mesh *mesh, int device, void *stream, svo_scene **out) {
	if (!mesh || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: null argument");
	*out = nullptr;
	if (mesh->position_stride_bytes < 12 || (mesh->position_stride_bytes & 3))
		return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: position stride must be >= 12 and a multiple of 4");
	if (mesh->n_indices % 3) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: index count is not a multiple of 3");
	if ((mesh->n_indices && (!mesh->indices || !mesh->positions)) || (mesh->n_draws && !mesh->draws))
		return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: null buffer");
	DeviceGuard guard(device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "svo_scene_create: cudaSetDevice failed (no CUDA device?)");
	SVO_TRY(configure_device_pool(device));
	cudaStream_t s = (cudaStream_t)stream;
	svo_scene *sc = new (std::nothrow) svo_scene();
	if (!sc) return fail(SVO_ERR_CUDA, "out of host memory");
	sc->device = device;
	sc->last_stream = s;
	sc->n_vertices = mesh->n_vertices;
	uint64_t tri_base = 0;
	for (uint32_t d = 0; d < mesh->n_draws; ++d) {
		const svo_draw &dr = mesh->draws[d];
		if (dr.texture_id != 0xffffffffu) {
			if (dr.texture_id >= mesh->n_textures || !mesh->textures || !mesh->texcoords || mesh->texcoord_stride_bytes < 8 ||
			    (mesh->texcoord_stride_bytes & 3)) {
				delete sc;
				return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: a textured draw needs its texture and texture coordinates");
			}
			if (dr.index_count) sc->textured = true;
		}
		if ((uint64_t)dr.first_index + dr.index_count > mesh->n_indices || dr.index_count % 3) {
			delete sc;
			return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: draw range outside the index buffer");
		}
		if (dr.index_count == 0) continue;
		DrawRec r;
		r.first_index = dr.first_index;
		r.tri_base = (uint32_t)tri_base;
		r.tri_count = dr.index_count / 3;
		r.rgb = dr.albedo_rgba8 & 0xffffffu;
		r.tex = dr.texture_id;
		sc->draws.push_back(r);
		tri_base += r.tri_count;
	}
	if (tri_base >= (1ull << 32)) {
		delete sc;
		return fail(SVO_ERR_CAPACITY, "more than 2^32-1 triangles");
	}
	int rc = 0;
	do {
		if (mesh->on_device) {
			sc->view.pos = (const unsigned char *)mesh->positions;
			sc->view.idx = mesh->indices;
		} else {
			const uint64_t pbytes = mesh->n_vertices * (uint64_t)mesh->position_stride_bytes;
			if ((rc = sc->pos.alloc(pbytes, s))) break;
			if ((rc = sc->idx.alloc(mesh->n_indices, s))) break;
			if (pbytes && cudaMemcpyAsync(sc->pos.p, mesh->positions, pbytes, cudaMemcpyHostToDevice, s) != cudaSuccess) rc = SVO_ERR_CUDA;
			if (mesh->n_indices &&
			    cudaMemcpyAsync(sc->idx.p, mesh->indices, mesh->n_indices * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
				rc = SVO_ERR_CUDA;
			if (rc) {
				set_error("svo_scene_create: host to device copy failed");
				break;
			}
			sc->view.pos = sc->pos.p;
			sc->view.idx = sc->idx.p;
		}
		if (sc->textured && (rc = upload_textures(sc, mesh, s))) break;
		if ((rc = sc->d_draws.alloc(sc->draws.size(), s))) break;
		// every index must address a vertex (a malformed mesh would make the raster kernels read out of bounds)
		DevBuf<uint32_t> max_index;
		uint32_t h_max = 0;
		if (mesh->n_indices) {
			if ((rc = max_index.alloc(1, s))) break;
			if (cudaMemsetAsync(max_index.p, 0, sizeof(uint32_t), s) != cudaSuccess) rc = SVO_ERR_CUDA;
			SVO_LAUNCH(std::min<uint32_t>(div_up(mesh->n_indices, 256 * 8), 1184u), 256, 0, s, k_max_index, sc->view.idx, mesh->n_indices,
			           max_index.p);
			if (cudaMemcpyAsync(&h_max, max_index.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = SVO_ERR_CUDA;
		}
		if (!sc->draws.empty() && cudaMemcpyAsync(sc->d_draws.p, sc->draws.data(), sc->draws.size() * sizeof(DrawRec),
		                                          cudaMemcpyHostToDevice, s) != cudaSuccess) {
			rc = fail(SVO_ERR_CUDA, "svo_scene_create: draw list copy failed");
			break;
		}
		// the draw list is staged from a host vector owned by the scene: make the copy complete before returning
		if (cudaStreamSynchronize(s) != cudaSuccess) rc = fail(SVO_ERR_CUDA, "svo_scene_create: stream sync failed");
		max_index.release(s);
		if (!rc && mesh->n_indices && h_max >= mesh->n_vertices)
			rc = fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_create: an index is >= n_vertices");
	} while (0);
	if (rc) {
		svo_scene_destroy(sc);
		return rc;
	}
	sc->view.stride = mesh->position_stride_bytes;
	sc->view.draws = sc->d_draws.p;
	sc->view.n_draws = (uint32_t)sc->draws.size();
	sc->view.n_tri = tri_base;
	*out = sc;
	return SVO_OK;
}

void svo_scene_destroy(svo_scene *sc) {
	if (!sc) return;
	DeviceGuard guard(sc->device);
	const cudaStream_t s = sc->last_stream; // stream-ordered frees: after the work that may still read the buffers
	sc->pos.release(s);
	sc->idx.release(s);
	sc->d_draws.release(s);
	sc->uv.release(s), sc->texels.release(s), sc->tex_desc.release(s), sc->tex_decode.release(s), sc->tex_enc.release(s);
	sc->tex_lod.release(s);
	delete sc;
}
uint64_t svo_scene_triangle_count(const svo_scene *sc) { return sc ? sc->view.n_tri : 0; }
int svo_scene_texture_level(const svo_scene *sc, uint32_t texture, uint32_t level, uint32_t *width, uint32_t *height,
                            const uint32_t **d_texels) {
	if (!sc || !width || !height || !d_texels) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_texture_level: null argument");
	if (texture >= sc->h_desc.size() || level >= sc->h_desc[texture].levels)
		return fail(SVO_ERR_INVALID_ARGUMENT, "svo_scene_texture_level: no such texture level");
	const TexDesc &d = sc->h_desc[texture];
	*width = std::max(d.w >> level, 1u), *height = std::max(d.h >> level, 1u);
	*d_texels = sc->texels.p + d.off[level];
	return (int)d.levels;
}

// ---- brick path switch -------------------------------------------------------------------------------
// -1 (default): bricks when the large triangles hold a good part of the fragments; 0: never (every fragment is emitted
// and sorted); 1: whenever the level allows it.  Read when a voxelizer is created.
static int g_build_path = -2; // -2: not set yet (the environment variable SVO_BUILD_PATH = 0 / 1 is looked at once)
static bool brick_path_wanted(const svo_voxelizer *v) {
	if (g_build_path == -2) {
		const char *e = getenv("SVO_BUILD_PATH");
		g_build_path = (e && (e[0] == '0' || e[0] == '1') && !e[1]) ? e[0] - '0' : -1;
	}
	if (v->key_level < 4 || v->n_frag_large == 0 || v->n_large == 0 || v->n_large > PAIR_LI_MASK || g_build_path == 0) return false;
	if (g_build_path == 1) return true;
	// enough large-triangle fragments to pay for the path's extra launches (a Sponza-sized build of 5e6 fragments is
	// launch bound: 0.38 ms sorted, 0.57 ms binned), and at least a fifth of all fragments
	return v->n_frag_large >= (8ull << 20) && v->n_frag_large * 4 >= v->n_frag_small;
}
static dim3 brick_pair_grid(const svo_voxelizer *v) {
	// one warp per large triangle; few triangles with hundreds of tile rows each: deal a triangle's rows out over up to 256 warps
	// (a wall has 512 rows of 512 tiles)
	const uint32_t wgrid = div_up((uint64_t)v->n_large * 32, RASTER_BLOCK);
	return dim3(wgrid, v->n_large < 65536u ? std::min(256u, std::max(1u, (1u << v->level) / 16u)) : 1u);
}

// ------------------------------------------------------------------------------------------------------
static int voxelizer_create_impl(svo_scene *scene, uint32_t level, int mode, const svo_shard *shard, const uint32_t *win_lo,
                                 const uint32_t *win_hi, void *stream, svo_voxelizer **out);

int svo_voxelizer_create(svo_scene *scene, uint32_t level, int mode, const svo_shard *shard, void *stream, svo_voxelizer **out) {
	return voxelizer_create_impl(scene, level, mode, shard, nullptr, nullptr, stream, out);
}

int svo_voxelizer_create_windowed(svo_scene *scene, uint32_t level, int mode, const uint32_t window_lo[3], const uint32_t window_hi[3],
                                  void *stream, svo_voxelizer **out) {
	if (!window_lo || !window_hi) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_voxelizer_create_windowed: null window");
	return voxelizer_create_impl(scene, level, mode, nullptr, window_lo, window_hi, stream, out);
}

static int voxelizer_create_impl(svo_scene *scene, uint32_t level, int mode, const svo_shard *shard, const uint32_t *win_lo,
                                 const uint32_t *win_hi, void *stream, svo_voxelizer **out) {
	if (!scene || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_voxelizer_create: null argument");
	*out = nullptr;
	if (mode != SVO_CENTER && mode != SVO_CONSERVATIVE_EXACT && mode != SVO_CONSERVATIVE_DILATE)
		return fail(SVO_ERR_INVALID_ARGUMENT, "unknown raster mode");
	uint32_t sl = shard ? shard->shard_level : 0;
	if (level < 1 || level > 14 || sl >= level) return fail(SVO_ERR_INVALID_ARGUMENT, "level must be 1..14 and shard_level < level");
	const uint32_t key_level = level - sl;
	if (3 * key_level + 24 > 64) return fail(SVO_ERR_CAPACITY, "3*(level - shard_level) + 24 colour bits must fit 64-bit fragments: shard levels >= 14");
	if (shard)
		for (int k = 0; k < 3; ++k)
			if (shard->cube_index[k] >= (1u << sl)) return fail(SVO_ERR_INVALID_ARGUMENT, "shard cube index out of range");
	DeviceGuard guard(scene->device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "cudaSetDevice failed");
	cudaStream_t s = (cudaStream_t)stream;
	svo_voxelizer *v = new (std::nothrow) svo_voxelizer();
	if (!v) return fail(SVO_ERR_CUDA, "out of host memory");
	v->scene = scene;
	v->device = scene->device;
	v->last_stream = s;
	scene->last_stream = s;
	v->level = level;
	v->key_level = key_level;
	v->rp.res = 1u << level;
	v->rp.mode = mode == SVO_CENTER ? MODE_CENTER : (mode == SVO_CONSERVATIVE_EXACT ? MODE_CONSERVATIVE : MODE_DILATE);
	const uint32_t side = 1u << key_level;
	for (int k = 0; k < 3; ++k) {
		const uint32_t ci = shard ? shard->cube_index[k] : 0;
		v->rp.sb.lo[k] = ci * side;
		v->rp.sb.hi[k] = ci * side + side;
		v->rp.origin[k] = ci * side;
		if (win_lo) { // a window of the whole grid: fragments keep their global coordinates
			if (win_lo[k] >= win_hi[k] || win_hi[k] > (1u << level)) {
				delete v;
				return fail(SVO_ERR_INVALID_ARGUMENT, "svo_voxelizer_create_windowed: bad window");
			}
			v->rp.sb.lo[k] = win_lo[k], v->rp.sb.hi[k] = win_hi[k];
		}
	}

	// ---- the count pass (Voxelizer::count_and_create_fragment_list, src/Voxelizer.cpp:134-165) ----
	const uint64_t T = scene->view.n_tri;
	DevBuf<uint32_t> cnt_small, rows, row_len, row_x0;
	DevBuf<uint64_t> large_index, row_prefix, dense_index, frag_prefix;
	ScanScratch ss;
	int rc = 0;
	do {
		if ((rc = v->t_raster.init())) break;
		if ((rc = v->tri_off.alloc(T + 1, s))) break;
		if (T == 0) {
			if (cudaMemsetAsync(v->tri_off.p, 0, sizeof(uint64_t), s) != cudaSuccess) rc = fail(SVO_ERR_CUDA, "memset failed");
			break;
		}
		if ((rc = cnt_small.alloc(T, s)) || (rc = rows.alloc(T, s)) || (rc = large_index.alloc(T + 1, s)) || (rc = row_prefix.alloc(T + 1, s)))
			break;
		const uint32_t tgrid = div_up(T, RASTER_BLOCK);
		if (scene->textured)
			SVO_LAUNCH_INDEP(tgrid, RASTER_BLOCK, s, k_classify_count<true>, scene->view, v->rp, cnt_small.p, rows.p);
		else
			SVO_LAUNCH_INDEP(tgrid, RASTER_BLOCK, s, k_classify_count<false>, scene->view, v->rp, cnt_small.p, rows.p);
		if ((rc = exclusive_scan((const uint32_t *)cnt_small.p, v->tri_off.p, T, ss, s))) break;
		// large-triangle index and first row of every large triangle: two scans of rows[] (entries != 0, and the values)
		if ((rc = exclusive_scan<uint32_t, true>((const uint32_t *)rows.p, large_index.p, T, ss, s))) break;
		if ((rc = exclusive_scan((const uint32_t *)rows.p, row_prefix.p, T, ss, s))) break;
		uint64_t h_small = 0, h_nlarge = 0, rows_sparse = 0;
		if (cudaMemcpyAsync(&h_small, v->tri_off.p + T, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
		    cudaMemcpyAsync(&h_nlarge, large_index.p + T, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
		    cudaMemcpyAsync(&rows_sparse, row_prefix.p + T, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
		    cudaStreamSynchronize(s) != cudaSuccess) {
			rc = fail(SVO_ERR_CUDA, "count pass failed");
			set_error("count pass failed: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}
		v->n_frag_small = h_small;
		v->n_large = (uint32_t)h_nlarge;
		if (v->n_large) {
			if (rows_sparse >= (1ull << 32)) {
				rc = fail(SVO_ERR_CAPACITY, "too many rows");
				break;
			}
			if ((rc = v->large.alloc(v->n_large, s)) || (rc = row_len.alloc(rows_sparse, s)) || (rc = row_x0.alloc(rows_sparse, s)) ||
			    (rc = dense_index.alloc(rows_sparse + 1, s)) || (rc = frag_prefix.alloc(rows_sparse + 1, s)))
				break;
			SVO_LAUNCH_INDEP(tgrid, RASTER_BLOCK, s, k_large_collect, T, (const uint32_t *)rows.p, (const uint64_t *)large_index.p,
			                 (const uint64_t *)row_prefix.p, v->large.p);
			const uint32_t wgrid = div_up((uint64_t)v->n_large * 32, RASTER_BLOCK);
			// few large triangles with thousands of rows each: spread a triangle's rows over up to 16 warps
			const dim3 wgrid2(wgrid, v->n_large < 65536u ? std::min(16u, div_up((uint64_t)(1u << level), LARGE_ROW_CHUNK)) : 1u);
			if (scene->textured) {
				if ((rc = v->large_uv.alloc(v->n_large, s))) break;
				SVO_LAUNCH_INDEP(wgrid2, RASTER_BLOCK, s, k_large_rows<true>, scene->view, v->rp, v->n_large, v->large.p, v->large_uv.p,
				                 row_len.p, row_x0.p);
			} else
				SVO_LAUNCH_INDEP(wgrid2, RASTER_BLOCK, s, k_large_rows<false>, scene->view, v->rp, v->n_large, v->large.p, (UvMap *)nullptr,
				                 row_len.p, row_x0.p);
			if ((rc = exclusive_scan<uint32_t, true>((const uint32_t *)row_len.p, dense_index.p, rows_sparse, ss, s))) break;
			if ((rc = exclusive_scan((const uint32_t *)row_len.p, frag_prefix.p, rows_sparse, ss, s))) break;
			uint64_t h_rows = 0, h_frag = 0;
			if (cudaMemcpyAsync(&h_rows, dense_index.p + rows_sparse, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
			    cudaMemcpyAsync(&h_frag, frag_prefix.p + rows_sparse, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
			    cudaStreamSynchronize(s) != cudaSuccess) {
				rc = fail(SVO_ERR_CUDA, "row pass failed");
				break;
			}
			v->n_rows = (uint32_t)h_rows;
			v->n_frag_large = h_frag;
			if (h_frag >= 0xffffffffull) {
				rc = fail(SVO_ERR_CAPACITY, "more than 2^32-2 fragments (the reference's counter is 32-bit too)");
				break;
			}
			if ((rc = v->row_off.alloc((uint64_t)v->n_rows + 1, s)) || (rc = v->row_xy.alloc(v->n_rows, s)) ||
			    (rc = v->row_li.alloc(v->n_rows, s)))
				break;
			DenseRows dr{v->row_off.p, v->row_xy.p, v->row_li.p};
			SVO_LAUNCH_INDEP(wgrid2, RASTER_BLOCK, s, k_rows_compact, v->n_large, (const LargeTri *)v->large.p, (const uint32_t *)row_len.p,
			                 (const uint32_t *)row_x0.p, (const uint64_t *)dense_index.p, (const uint64_t *)frag_prefix.p, rows_sparse, dr);
			// brick path: count the (brick, triangle) pairs of the large triangles (part of the count pass: exact sizes)
			v->brick = brick_path_wanted(v);
			if (v->brick) {
				DevBuf<uint32_t> n_tr, pair_cnt;
				do {
					if ((rc = n_tr.alloc(v->n_large, s)) || (rc = v->tr_base.alloc((uint64_t)v->n_large + 1, s))) break;
					SVO_LAUNCH_INDEP(div_up(v->n_large, RASTER_BLOCK), RASTER_BLOCK, s, k_brick_tile_rows, v->rp, v->n_large,
					                 (const LargeTri *)v->large.p, n_tr.p);
					if ((rc = exclusive_scan((const uint32_t *)n_tr.p, v->tr_base.p, v->n_large, ss, s))) break;
					uint64_t h_tr = 0;
					if (cudaMemcpyAsync(&h_tr, v->tr_base.p + v->n_large, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
					    cudaStreamSynchronize(s) != cudaSuccess) {
						rc = fail(SVO_ERR_CUDA, "tile row pass failed");
						break;
					}
					v->n_tile_rows = h_tr;
					if ((rc = pair_cnt.alloc(h_tr, s)) || (rc = v->pair_off.alloc(h_tr + 1, s))) break;
					SVO_LAUNCH(brick_pair_grid(v), RASTER_BLOCK, 0, s, k_brick_pairs<false>, v->rp, v->n_large, (const LargeTri *)v->large.p,
					           (const uint64_t *)v->tr_base.p, pair_cnt.p, (const uint64_t *)nullptr, (uint64_t *)nullptr);
					if ((rc = exclusive_scan((const uint32_t *)pair_cnt.p, v->pair_off.p, h_tr, ss, s))) break;
					uint64_t h_pairs = 0;
					if (cudaMemcpyAsync(&h_pairs, v->pair_off.p + h_tr, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
					    cudaStreamSynchronize(s) != cudaSuccess) {
						rc = fail(SVO_ERR_CUDA, "pair count pass failed");
						break;
					}
					v->n_pairs_large = h_pairs;
					if (h_pairs >= (1ull << 31)) v->brick = false; // (brick_first holds 32-bit pair indices)
				} while (0);
				n_tr.release(s), pair_cnt.release(s);
				if (rc) break;
			}
		}
		v->n_frag = v->n_frag_small + v->n_frag_large;
		if (v->n_frag >= 0xffffffffull) {
			rc = fail(SVO_ERR_CAPACITY, "more than 2^32-2 fragments (the reference's counter is 32-bit too)");
			break;
		}
	} while (0);
	if (!rc) rc = v->frags.alloc(v->n_frag, s);
	if (!rc && cudaGetLastError() != cudaSuccess) rc = fail(SVO_ERR_CUDA, "voxelizer count pass: kernel launch failed");
	cnt_small.release(s), rows.release(s), row_len.release(s), row_x0.release(s);
	large_index.release(s), row_prefix.release(s), dense_index.release(s), frag_prefix.release(s);
	ss.state.release(s), ss.ticket.release(s);
	if (rc) {
		svo_voxelizer_destroy(v);
		return rc;
	}
	*out = v;
	return SVO_OK;
}

int svo_voxelizer_create_from_fragments(int device, uint32_t level, const uint64_t *fragments, uint64_t n, int on_device, void *stream,
                                        svo_voxelizer **out) {
	if (!out || (n && !fragments)) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_voxelizer_create_from_fragments: null argument");
	*out = nullptr;
	if (level < 1 || 3 * level + 24 > 64) return fail(SVO_ERR_INVALID_ARGUMENT, "level must be 1..13");
	if (n >= 0xffffffffull) return fail(SVO_ERR_CAPACITY, "more than 2^32-2 fragments");
	DeviceGuard guard(device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "cudaSetDevice failed (no CUDA device?)");
	SVO_TRY(configure_device_pool(device));
	cudaStream_t s = (cudaStream_t)stream;
	svo_voxelizer *v = new (std::nothrow) svo_voxelizer();
	if (!v) return fail(SVO_ERR_CUDA, "out of host memory");
	v->device = device;
	v->last_stream = s;
	v->level = v->key_level = level;
	v->n_frag = n;
	int rc = v->t_raster.init();
	if (!rc) rc = v->frags.alloc(n, s);
	if (!rc) rc = v->ext_frags.alloc(n, s);
	if (!rc && n &&
	    cudaMemcpyAsync(v->ext_frags.p, fragments, n * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s) != cudaSuccess)
		rc = fail(SVO_ERR_CUDA, "fragment copy failed");
	if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = fail(SVO_ERR_CUDA, "stream sync failed");
	if (rc) {
		svo_voxelizer_destroy(v);
		return rc;
	}
	*out = v;
	return SVO_OK;
}

void svo_voxelizer_destroy(svo_voxelizer *v) {
	if (!v) return;
	DeviceGuard guard(v->device);
	const cudaStream_t s = v->last_stream;
	v->large_uv.release(s);
	v->tri_off.release(s), v->large.release(s), v->row_off.release(s), v->row_xy.release(s), v->row_li.release(s), v->frags.release(s);
	v->ext_frags.release(s);
	v->tr_base.release(s), v->pair_off.release(s);
	v->t_raster.destroy();
	delete v;
}

static int emit_large_fragments(svo_voxelizer *v, cudaStream_t s) {
	const SceneView &sv = v->scene->view;
	const bool tex = v->scene->textured;
	if (v->n_frag_large) {
		DenseRows dr{v->row_off.p, v->row_xy.p, v->row_li.p};
		if (tex)
		{
			SVO_LAUNCH(div_up(v->n_frag_large, EMIT_TILE), EMIT_BLOCK, 0, s, k_emit_large<true>, v->rp, sv.tex, (const LargeTri *)v->large.p,
			           (const UvMap *)v->large_uv.p, dr, v->n_rows, v->n_frag_large, v->frags.p + v->n_frag_small);
			// rows of large alpha-tested triangles (none in most scenes: the kernel then exits at once)
			SVO_LAUNCH(div_up(v->n_rows, RASTER_BLOCK), RASTER_BLOCK, 0, s, k_emit_alpha_rows<true>, v->rp, sv.tex,
			           (const LargeTri *)v->large.p, (const UvMap *)v->large_uv.p, dr, v->n_rows, v->frags.p + v->n_frag_small);
		}
		else
			SVO_LAUNCH(div_up(v->n_frag_large, EMIT_TILE), EMIT_BLOCK, 0, s, k_emit_large<false>, v->rp, sv.tex, (const LargeTri *)v->large.p,
			           (const UvMap *)nullptr, dr, v->n_rows, v->n_frag_large, v->frags.p + v->n_frag_small);
	}
	SVO_CUDA_TRY(cudaGetLastError());
	v->large_emitted = true;
	return SVO_OK;
}

int svo_voxelizer_voxelize(svo_voxelizer *v, void *stream) {
	if (!v) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_voxelizer_voxelize: null handle");
	DeviceGuard guard(v->device);
	cudaStream_t s = (cudaStream_t)stream;
	v->last_stream = s;
	if (v->scene) v->scene->last_stream = s;
	if (!v->scene) { // fragment list supplied by the caller: restore it (the builder sorts the list in place)
		SVO_CUDA_TRY(cudaEventRecord(v->t_raster.a, s));
		if (v->n_frag) SVO_CUDA_TRY(cudaMemcpyAsync(v->frags.p, v->ext_frags.p, v->n_frag * 8, cudaMemcpyDeviceToDevice, s));
		SVO_CUDA_TRY(cudaEventRecord(v->t_raster.b, s));
		v->t_raster.recorded = v->voxelized = true;
		return SVO_OK;
	}
	const SceneView &sv = v->scene->view;
	SVO_CUDA_TRY(cudaEventRecord(v->t_raster.a, s));
	const bool tex = v->scene->textured;
	if (v->n_frag_small) {
		if (tex)
			SVO_LAUNCH(div_up(sv.n_tri, RASTER_BLOCK), RASTER_BLOCK, 0, s, k_emit_small<true>, sv, v->rp, (const uint64_t *)v->tri_off.p,
			           v->frags.p);
		else
			SVO_LAUNCH(div_up(sv.n_tri, RASTER_BLOCK), RASTER_BLOCK, 0, s, k_emit_small<false>, sv, v->rp, (const uint64_t *)v->tri_off.p,
			           v->frags.p);
	}
	// brick path: the large triangles are never emitted as fragments (the builder bins them: brick.cuh) unless somebody
	// asks for the fragment list (svo_voxelizer_fragments / svo_voxelizer_export_reference_fragments)
	v->large_emitted = false;
	if (v->n_frag_large && !v->brick) SVO_TRY(emit_large_fragments(v, s));
	SVO_CUDA_TRY(cudaEventRecord(v->t_raster.b, s));
	SVO_CUDA_TRY(cudaGetLastError());
	v->t_raster.recorded = true;
	v->voxelized = true;
	return SVO_OK;
}

uint32_t svo_voxelizer_level(const svo_voxelizer *v) { return v ? v->level : 0; }
uint32_t svo_voxelizer_resolution(const svo_voxelizer *v) { return v ? 1u << v->level : 0; }
uint64_t svo_voxelizer_fragment_count(const svo_voxelizer *v) { return v ? v->n_frag : 0; }
// brick path: the list is completed on request (the large triangles' fragments are not part of the build any more)
static int complete_fragment_list(const svo_voxelizer *cv) {
	svo_voxelizer *v = const_cast<svo_voxelizer *>(cv);
	if (!v->brick || !v->voxelized || v->large_emitted || !v->n_frag_large) return SVO_OK;
	DeviceGuard guard(v->device);
	SVO_TRY(emit_large_fragments(v, v->last_stream));
	SVO_CUDA_TRY(cudaStreamSynchronize(v->last_stream)); // the caller may read the list on any stream
	return SVO_OK;
}
const uint64_t *svo_voxelizer_fragments(const svo_voxelizer *v) {
	if (!v || complete_fragment_list(v) != SVO_OK) return nullptr;
	return v->frags.p;
}

int svo_voxelizer_export_reference_fragments(const svo_voxelizer *v, uint32_t *d_out, void *stream) {
	if (!v || !d_out) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (v->key_level > 12) return fail(SVO_ERR_UNSUPPORTED, "the reference fragment packing holds 12 bits per axis (voxelizer.frag:40-42)");
	if (!v->voxelized) return fail(SVO_ERR_NOT_READY, "voxelize first");
	SVO_TRY(complete_fragment_list(v));
	DeviceGuard guard(v->device);
	cudaStream_t s = (cudaStream_t)stream;
	if (v->n_frag)
		SVO_LAUNCH_INDEP(div_up(v->n_frag, 256), 256, s, k_export_reference_fragments, (const uint64_t *)v->frags.p, v->n_frag,
		                 reinterpret_cast<uint2 *>(d_out));
	SVO_CUDA_TRY(cudaGetLastError());
	return SVO_OK;
}

int svo_voxelizer_last_ms(svo_voxelizer *v, float *raster_ms) {
	if (!v || !raster_ms) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!v->t_raster.recorded) return fail(SVO_ERR_NOT_READY, "voxelize first");
	DeviceGuard guard(v->device);
	SVO_CUDA_TRY(cudaEventSynchronize(v->t_raster.b));
	SVO_CUDA_TRY(cudaEventElapsedTime(raster_ms, v->t_raster.a, v->t_raster.b));
	return SVO_OK;
}

// ------------------------------------------------------------------------------------------------------
int svo_builder_create(svo_voxelizer *vox, void *stream, svo_builder **out) {
	if (!vox || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_create: null argument");
	*out = nullptr;
	DeviceGuard guard(vox->device);
	cudaStream_t s = (cudaStream_t)stream;
	svo_builder *b = new (std::nothrow) svo_builder();
	if (!b) return fail(SVO_ERR_CUDA, "out of host memory");
	b->vox = vox;
	b->device = vox->device;
	b->last_stream = s;
	b->level = vox->key_level;
	int rc = 0;
	do {
		for (int i = 0; i <= SVO_PHASE_COUNT; ++i)
			if (cudaEventCreate(&b->ev[i]) != cudaSuccess) rc = fail(SVO_ERR_CUDA, "cudaEventCreate failed");
		for (int i = 0; i < 5; ++i)
			if (cudaEventCreate(&b->ev_brick[i]) != cudaSuccess) rc = fail(SVO_ERR_CUDA, "cudaEventCreate failed");
		b->h_pinned = pinned_slot_acquire();
		if (!b->h_pinned) rc = fail(SVO_ERR_CUDA, "no page-locked host memory");
		if (rc) break;
		const uint64_t F = vox->n_frag;
		if ((rc = b->tmp.alloc(F, s)) || (rc = b->leaf.alloc(F, s)) || (rc = b->counts.alloc(MAX_LEVEL + 2, s))) break;
		// pooled arrays: first[d] has one entry per depth-(d-1) node (<= min(F, 8^(d-1))), slot[d] one per depth-d node
		uint64_t pool_first = 0, pool_slot = 0;
		for (uint32_t d = 1; d <= b->level; ++d) {
			pool_first += node_cap(F, d - 1);
			pool_slot += (node_cap(F, d) + 7) & ~7ull; // every depth's slot array starts 8-byte aligned (k_parent_compact stores 8 slots at once)
		}
		if ((rc = b->first.alloc(pool_first, s)) || (rc = b->slot.alloc(pool_slot, s))) break;
		const uint64_t tiles = (F + CMP_TILE - 1) / CMP_TILE + 1;
		if ((rc = b->lb_state.alloc(tiles * (b->level + 4), s)) || (rc = b->tickets.alloc(b->level + 2, s))) break;
	} while (0);
	if (rc) {
		svo_builder_destroy(b);
		return rc;
	}
	*out = b;
	return SVO_OK;
}

void svo_builder_destroy(svo_builder *b) {
	if (!b) return;
	DeviceGuard guard(b->device);
	const cudaStream_t s = b->last_stream;
	b->tmp.release(s), b->leaf.release(s), b->first.release(s), b->slot.release(s), b->counts.release(s), b->lb_state.release(s);
	b->tickets.release(s), b->octree.release(s), b->root_scratch.release(s);
	b->sort_scratch.release(s);
	b->scan_scratch.state.release(s), b->scan_scratch.ticket.release(s);
	b->rf_cnt01.release(s), b->rf_cnt2.release(s), b->rf_pre01.release(s), b->rf_pre2.release(s);
	b->pairs_a.release(s), b->pairs_b.release(s), b->brick_u64.release(s), b->brick_scalars.release(s);
	b->pair_flags.release(s), b->brick_first.release(s), b->small_leaf.release(s), b->brick_u32.release(s), b->brick_temp.release(s);
	for (int i = 0; i <= SVO_PHASE_COUNT; ++i)
		if (b->ev[i]) cudaEventDestroy(b->ev[i]);
	for (int i = 0; i < 5; ++i)
		if (b->ev_brick[i]) cudaEventDestroy(b->ev_brick[i]);
	if (b->aux) cudaStreamSynchronize(b->aux), cudaStreamDestroy(b->aux);
	if (b->ev_fork) cudaEventDestroy(b->ev_fork);
	if (b->ev_join) cudaEventDestroy(b->ev_join);
	pinned_slot_release(b->h_pinned);
	release_export(b);
	delete b;
}

// count the runs of every tile, scan over the tiles, then the big kernel: no tile waits for a neighbour
static int reduce_sorted(svo_builder *b, const uint64_t *sorted, uint64_t F, uint32_t K, const FusedOut &fo, cudaStream_t s) {
	const uint32_t rf_tiles = div_up(F, RF_TILE);
	SVO_TRY(b->rf_cnt01.reserve(rf_tiles, s));
	SVO_TRY(b->rf_cnt2.reserve(rf_tiles, s));
	SVO_TRY(b->rf_pre01.reserve((uint64_t)rf_tiles + 1, s));
	SVO_TRY(b->rf_pre2.reserve((uint64_t)rf_tiles + 1, s));
	const uint64_t *p01 = b->rf_pre01.p, *p2 = b->rf_pre2.p;
#define SVO_REDUCE_CASE(KK)                                                                                                        \
	{                                                                                                                              \
		SVO_LAUNCH(rf_tiles, RF_BLOCK, 0, s, k_reduce_count<KK>, sorted, F, b->rf_cnt01.p, b->rf_cnt2.p);                            \
		SVO_TRY(exclusive_scan((const uint64_t *)b->rf_cnt01.p, b->rf_pre01.p, rf_tiles, b->scan_scratch, s));                      \
		if (KK >= 3) SVO_TRY(exclusive_scan((const uint64_t *)b->rf_cnt2.p, b->rf_pre2.p, rf_tiles, b->scan_scratch, s));           \
		SVO_LAUNCH(rf_tiles, RF_BLOCK, 0, s, k_reduce_fused<KK>, sorted, F, fo, rf_tiles, p01, p2);                                  \
	}
	if (K == 1) SVO_REDUCE_CASE(1) else if (K == 2) SVO_REDUCE_CASE(2) else SVO_REDUCE_CASE(3)
#undef SVO_REDUCE_CASE
	return 0;
}

// the builder's second stream and its fork / join events, created on first use
static int ensure_aux_stream(svo_builder *b) {
	if (b->aux) return SVO_OK;
	if (cudaStreamCreateWithFlags(&b->aux, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming) != cudaSuccess)
		return fail(SVO_ERR_CUDA, "cannot create the builder's auxiliary stream");
	return SVO_OK;
}

// The brick path of svo_builder_prepare (brick.cuh): small triangles' fragments sorted and reduced on their own, large
// triangles binned; on return *keys_top holds the depth L-3 keys (counts[L-3] of them: the non-empty bricks) and *free_buf is free.
//   ev[0..1] small fragments: sort + reduce + small records (+ the read-back of their number)
//   ev[1..2] pairs: generation, sort by brick, brick heads        ev[2..3] k_brick_flat, k_brick_raster, k_brick_ranks
static int prepare_bricks(svo_builder *b, cudaStream_t s, const uint64_t *first_off, const uint64_t *slot_off, uint64_t **free_buf,
                          uint64_t **keys_top) {
	svo_voxelizer *v = b->vox;
	const uint32_t L = b->level;
	const int n_sm = sm_count(b->device);
	const uint64_t ns = v->n_frag_small, npl = v->n_pairs_large;
	uint64_t *A = v->frags.p, *B = b->tmp.p; // A: sorted small fragments, dead after their reduce; B: their unique Morton codes
	SVO_TRY(b->brick_scalars.reserve(4, s));
	SVO_CUDA_TRY(cudaMemsetAsync(b->brick_scalars.p, 0, 4 * sizeof(uint64_t), s));
	uint64_t *d_nsl = b->brick_scalars.p, *d_nsb = b->brick_scalars.p + 1;
	uint64_t *h_small = b->h_pinned; // leaves of small triangles, bricks that hold some
	h_small[0] = h_small[1] = 0;
	b->sort_passes = 0;
	// The pair list: the large triangles' pairs first (in triangle order), the small records behind them; the pair sort is
	// stable, so a brick's small record ends up as its last pair (k_brick_raster handles it first).  The large pairs do
	// not depend on the small chain: they are generated while the host waits for the small chain's two counts.
	SVO_TRY(b->pairs_a.reserve(npl + ns, s)); // (at most one small record per small fragment)
	if (npl) { // on the builder's second stream, next to the small chain
		SVO_TRY(ensure_aux_stream(b));
		SVO_CUDA_TRY(cudaEventRecord(b->ev_fork, s));
		SVO_CUDA_TRY(cudaStreamWaitEvent(b->aux, b->ev_fork, 0));
		SVO_LAUNCH(brick_pair_grid(v), RASTER_BLOCK, 0, b->aux, k_brick_pairs<true>, v->rp, v->n_large, (const LargeTri *)v->large.p,
		           (const uint64_t *)v->tr_base.p, (uint32_t *)nullptr, (const uint64_t *)v->pair_off.p, b->pairs_a.p);
		SVO_CUDA_TRY(cudaEventRecord(b->ev_join, b->aux));
	}
	if (ns) {
		uint64_t *sorted = A;
		SVO_TRY(radix_sort_u64(v->frags.p, b->tmp.p, ns, 24, 24 + 3 * L, b->sort_scratch, b->device, n_sm, s, &sorted, &b->sort_passes, nullptr));
		A = sorted, B = sorted == v->frags.p ? b->tmp.p : v->frags.p;
		SVO_TRY(b->small_leaf.reserve(ns, s));
		FusedOut fo{};
		fo.leaf = b->small_leaf.p;
		fo.keys_top = B;
		fo.count[0] = d_nsl;
		SVO_TRY(reduce_sorted(b, A, ns, 1, fo, s));
		const uint32_t g = std::min<uint32_t>(div_up(ns, 256), (uint32_t)n_sm * 8u);
		SVO_LAUNCH(g, 256, 0, s, k_brick_small_records, (const uint64_t *)B, (const uint64_t *)d_nsl, b->pairs_a.p + npl,
		           reinterpret_cast<unsigned long long *>(d_nsb));
		SVO_CUDA_TRY(cudaMemcpyAsync(h_small, b->brick_scalars.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
	}
	SVO_CUDA_TRY(cudaEventRecord(b->ev[1], s));
	if (npl) SVO_CUDA_TRY(cudaStreamWaitEvent(s, b->ev_join, 0));
	if (ns) SVO_CUDA_TRY(cudaEventSynchronize(b->ev[1])); // (the two counts have arrived)
	b->n_small_leaves = h_small[0];
	const uint64_t nsb = h_small[1], n_pairs = npl + nsb;
	if (n_pairs >= (1ull << 32)) return fail(SVO_ERR_CAPACITY, "brick path: more than 2^32-1 (brick, triangle) pairs");
	b->n_pairs = n_pairs;
	SVO_TRY(b->pairs_b.reserve(n_pairs, s));
	SVO_TRY(b->pair_flags.reserve(n_pairs, s));
	SVO_TRY(b->brick_first.reserve(n_pairs + 1, s));
	uint64_t *pairs = b->pairs_a.p;
	uint32_t pair_passes = 0;
	SVO_TRY(radix_sort_u64(b->pairs_a.p, b->pairs_b.p, n_pairs, PAIR_SORT_BEGIN, 33 + 3 * (L - BRICK_LOG), b->sort_scratch, b->device, n_sm, s,
	                       &pairs, &pair_passes, nullptr));
	uint64_t *brick_code = pairs == b->pairs_a.p ? b->pairs_b.p : b->pairs_a.p; // (the other pair buffer is free after the sort)
	uint64_t *d_nbricks = b->brick_scalars.p + 3;
	{
		const uint32_t tiles = div_up(n_pairs, SCAN_TILE);
		SVO_TRY(b->scan_scratch.state.reserve((uint64_t)tiles + 1, s));
		SVO_TRY(b->scan_scratch.ticket.reserve(1, s));
		SVO_CUDA_TRY(cudaMemsetAsync(b->scan_scratch.state.p, 0, ((uint64_t)tiles + 1) * sizeof(uint64_t), s));
		SVO_CUDA_TRY(cudaMemsetAsync(b->scan_scratch.ticket.p, 0, sizeof(uint32_t), s));
		SVO_LAUNCH(tiles, SCAN_BLOCK, 0, s, k_brick_heads, (const uint64_t *)pairs, n_pairs, b->brick_first.p, brick_code, d_nbricks,
		           b->scan_scratch.state.p, b->scan_scratch.ticket.p);
	}
	SVO_CUDA_TRY(cudaEventRecord(b->ev[2], s));

	// per-brick arrays are sized for the upper bound "one brick per pair"; entries past the real number of bricks stay 0
	const uint64_t nbd = n_pairs;
	SVO_TRY(b->brick_u32.reserve(nbd * 4, s));
	SVO_TRY(b->brick_u64.reserve((nbd + 1) * 3, s));
	SVO_TRY(b->brick_temp.reserve(nbd * BRICK_CELLS, s)); // 2 KB per brick (fails with SVO_ERR_CUDA when the device cannot hold it: SVO_BUILD_PATH=0 sorts every fragment instead)
	SVO_CUDA_TRY(cudaMemsetAsync(b->brick_u32.p, 0, nbd * 4 * sizeof(uint32_t), s)); // the records (their counts: w)
	BrickArgs a{};
	a.pairs = pairs, a.brick_first = b->brick_first.p, a.brick_code = brick_code, a.n_bricks = d_nbricks;
	a.large = v->large.p, a.luv = v->large_uv.p, a.tv = v->scene->view.tex, a.rp = v->rp;
	a.small_keys = B, a.small_leaf = b->small_leaf.p, a.n_small = d_nsl;
	a.rec = reinterpret_cast<uint4 *>(b->brick_u32.p); // (16-byte aligned: the pool hands out 256-byte aligned blocks)
	for (int j = 0; j < 3; ++j) a.rank[j] = b->brick_u64.p + (nbd + 1) * j;
	a.n_bound = nbd;
	a.temp = b->brick_temp.p;
	a.keys_top = A;
	for (uint32_t j = 0; j < 3; ++j) a.count[j] = b->counts.p + (L - j);
	a.count3 = b->counts.p + (L - 3), a.first_l2 = b->first.p + first_off[L - 2], a.slot_l2 = b->slot.p + slot_off[L - 2];
	a.slow_list = b->pair_flags.p; // (n_pairs words: at most one entry per brick)
	a.n_slow = reinterpret_cast<unsigned long long *>(b->brick_scalars.p + 2);
	const uint32_t rgrid = div_up(nbd, (uint64_t)BRICK_WARPS * BRICK_BPW);
	SVO_CUDA_TRY(cudaEventRecord(b->ev_brick[0], s));
	SVO_LAUNCH(div_up(nbd, 256), 256, 0, s, k_brick_flat, a);
	if (v->scene->textured)
		SVO_LAUNCH(rgrid, BRICK_BLOCK, 0, s, k_brick_raster<true>, a);
	else
		SVO_LAUNCH(rgrid, BRICK_BLOCK, 0, s, k_brick_raster<false>, a);
	SVO_CUDA_TRY(cudaEventRecord(b->ev_brick[1], s));
	{ // ranks: three scans in one launch, which also writes the depth L-3 level (keys, first children, slots) and the four level counts
		const uint32_t tiles = div_up(nbd + 1, SCAN_TILE); // (one element more than records: it receives the totals)
		SVO_TRY(b->scan_scratch.state.reserve((uint64_t)(tiles + 1) * 3, s));
		SVO_TRY(b->scan_scratch.ticket.reserve(3, s));
		SVO_CUDA_TRY(cudaMemsetAsync(b->scan_scratch.state.p, 0, (uint64_t)(tiles + 1) * 3 * sizeof(uint64_t), s));
		SVO_CUDA_TRY(cudaMemsetAsync(b->scan_scratch.ticket.p, 0, 3 * sizeof(uint32_t), s));
		SVO_LAUNCH(tiles, SCAN_BLOCK, 0, s, k_brick_ranks, a, b->brick_u64.p + (nbd + 1), b->brick_u64.p + (nbd + 1) * 2, b->scan_scratch.state.p,
		           b->scan_scratch.ticket.p, (uint64_t)(tiles + 1));
	}
	SVO_CUDA_TRY(cudaEventRecord(b->ev_brick[2], s));
	b->brick_args = a; // k_brick_emit (after the sizes are known) works on the same arrays
	SVO_CUDA_TRY(cudaEventRecord(b->ev[3], s));
	SVO_CUDA_TRY(cudaGetLastError());
	*keys_top = A, *free_buf = B;
	return 0;
}

int svo_builder_prepare(svo_builder *b, void *stream) {
	if (!b) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_prepare: null handle");
	svo_voxelizer *v = b->vox;
	if (!v->voxelized) return fail(SVO_ERR_NOT_READY, "svo_builder_build: voxelize first");
	DeviceGuard guard(b->device);
	cudaStream_t s = (cudaStream_t)stream;
	b->last_stream = v->last_stream = s;
	const uint64_t F = v->n_frag;
	const uint32_t L = b->level;
	const int n_sm = sm_count(b->device);
	b->built = b->prepared = b->emitted = false;
	// The build consumes the fragment list: it is sorted in place and the two fragment-sized buffers then serve as key
	// ping-pong buffers for the upper levels.  A second build (or a fragment export) needs a new CmdVoxelize first.
	v->voxelized = false;

	SVO_CUDA_TRY(cudaEventRecord(b->ev[0], s));
	const uint32_t K = L < 3 ? L : 3;
	const uint64_t tiles_f = (F + CMP_TILE - 1) / CMP_TILE + 1;
	SVO_CUDA_TRY(cudaMemsetAsync(b->lb_state.p, 0, tiles_f * (L + 4) * sizeof(uint64_t), s));
	SVO_CUDA_TRY(cudaMemsetAsync(b->tickets.p, 0, (L + 2) * sizeof(uint32_t), s));
	SVO_CUDA_TRY(cudaMemsetAsync(b->counts.p, 0, (MAX_LEVEL + 2) * sizeof(uint64_t), s));
	uint64_t first_off[MAX_LEVEL + 2] = {}, slot_off[MAX_LEVEL + 2] = {};
	{
		uint64_t fo = 0, so = 0;
		for (uint32_t d = L; d >= 1; --d) {
			first_off[d] = fo, slot_off[d] = so;
			fo += node_cap(F, d - 1), so += (node_cap(F, d) + 7) & ~7ull;
		}
	}
	uint64_t *sorted = v->frags.p, *other = b->tmp.p;
	b->path = v->brick ? 1 : 0;
	if (v->brick) {
		// ---- large triangles binned to bricks, small ones sorted on their own: the three deepest levels (brick.cuh) ----
		SVO_TRY(prepare_bricks(b, s, first_off, slot_off, &sorted, &other));
	} else {
		// ---- sort by Morton code (stable) ----
		SVO_TRY(radix_sort_u64(v->frags.p, b->tmp.p, F, 24, 24 + 3 * L, b->sort_scratch, b->device, n_sm, s, &sorted, &b->sort_passes, b->ev[1]));
		other = sorted == v->frags.p ? b->tmp.p : v->frags.p;
		SVO_CUDA_TRY(cudaEventRecord(b->ev[2], s));
		// ---- de-duplicate + colour reduce + the two deepest parent levels, fused (build.cuh) ----
		if (F) {
			FusedOut fo{};
			fo.leaf = b->leaf.p;
			fo.slot0 = b->slot.p + slot_off[L];
			fo.first1 = b->first.p + first_off[L];
			if (L >= 2) fo.slot1 = b->slot.p + slot_off[L - 1], fo.first2 = b->first.p + first_off[L - 1];
			fo.keys_top = other;
			for (uint32_t j = 0; j < K; ++j) fo.count[j] = b->counts.p + (L - j);
			SVO_TRY(reduce_sorted(b, sorted, F, K, fo, s));
		}
		SVO_CUDA_TRY(cudaEventRecord(b->ev[3], s));
	}

	// ---- remaining levels (L-K+1)..1: unique parents, first child, child mask (small from here on) ----
	// key buffers ping-pong between the two fragment-sized buffers (the sorted fragments are dead after the reduce)
	const uint32_t pgrid = (uint32_t)n_sm * 4u;
	uint64_t *kin = other, *kout = sorted;
	uint32_t d = b->path == 1 ? L - 3 : L - K + 1; // (the brick path leaves the depth L-3 keys: a brick is a depth L-3 node)
	for (; d > TAIL_DEPTH && F; --d) {
		const uint64_t in_cap8 = d >= 11 ? UINT64_MAX : (1ull << (3 * d));
		const uint64_t in_cap = F < in_cap8 ? F : in_cap8;
		const uint64_t tiles = (in_cap + CMP_TILE - 1) / CMP_TILE + 1;
		uint32_t g = (uint32_t)(tiles < pgrid ? tiles : pgrid);
		SVO_LAUNCH(g, CMP_BLOCK, 0, s, k_parent_compact, (const uint64_t *)kin, (const uint64_t *)(b->counts.p + d), kout,
		           b->first.p + first_off[d], b->slot.p + slot_off[d], b->lb_state.p + tiles_f * (3 + L - d), b->tickets.p + (L - d + 1),
		           b->counts.p + d - 1);
		uint64_t *t = kin;
		kin = kout;
		kout = t;
	}
	if (d >= 1 && F) { // the levels near the root (<= 8^TAIL_DEPTH keys): one single-block launch for all of them
		TailArgs ta{};
		ta.keys[0] = kin, ta.keys[1] = kout, ta.counts = b->counts.p, ta.d0 = d;
		for (uint32_t e = 1; e <= d; ++e) ta.first[e] = b->first.p + first_off[e], ta.slot[e] = b->slot.p + slot_off[e];
		SVO_LAUNCH(1, TAIL_BLOCK, 0, s, k_parent_tail, ta);
	}
	SVO_CUDA_TRY(cudaEventRecord(b->ev[4], s));

	// ---- exact sizing: the one host round trip of the build ----
	SVO_CUDA_TRY(cudaMemcpyAsync(b->h_counts, b->counts.p, (L + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
	b->n_bricks = 0;
	if (b->path == 1) {
		SVO_CUDA_TRY(cudaMemcpyAsync(&b->n_bricks, b->brick_scalars.p + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
		SVO_CUDA_TRY(cudaMemcpyAsync(&b->n_slow, b->brick_scalars.p + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
	}
	SVO_CUDA_TRY(cudaStreamSynchronize(s));
	EmitParams &ep = b->ep;
	ep = EmitParams{};
	ep.level = L;
	uint64_t blocks = 1;
	ep.block_base[1] = 0;
	for (uint32_t d = 2; d <= L + 1; ++d) {
		ep.block_base[d] = blocks;
		if (d <= L) blocks += b->h_counts[d - 1];
	}
	ep.total_blocks = blocks;
	if (blocks * 8 >= (1ull << 30)) return fail(SVO_ERR_CAPACITY, "octree needs >= 2^30 words: 30-bit child pointers (octree.glsl:110) cannot address it");
	for (uint32_t d = 0; d <= L; ++d) ep.count[d] = b->h_counts[d];
	for (uint32_t d = 1; d <= L; ++d) {
		ep.first[d] = b->first.p + first_off[d];
		ep.slot[d] = b->slot.p + slot_off[d];
	}
	ep.leaf = b->leaf.p;
	if (b->path == 1) ep.total_blocks = ep.block_base[L - 1]; // the two deepest windows are written brick by brick (k_brick_emit)
	b->range_bytes = blocks * 8 * sizeof(uint32_t); // (counter + 1) * 8 * 4, src/OctreeBuilder.cpp:212-214
	b->prepared = true;
	return SVO_OK;
}

// Emit the node words of a prepared build.  skip_root = 0: blocks 0.. go to d_dst[0..), child pointers are
// block_index*8 + pointer_bias_words.  skip_root = 1 (multi-GPU stitch): blocks 1.. go to d_dst[0..) and the root
// block to the builder's root scratch; child pointers are (block_index-1)*8 + pointer_bias_words, i.e. they are
// already valid in a buffer where d_dst sits at word offset pointer_bias_words.  d_dst may be peer memory.
static int emit_into(svo_builder *b, uint32_t *d_dst, uint32_t bias, int skip_root, cudaStream_t s, uint32_t brick_parts = BRICK_EMIT_ALL,
                     uint64_t *plan = nullptr) {
	EmitParams ep = b->ep;
	// skip_root: 0 = whole tree; 1 = the root block is kept aside; 2 = the root block and the depth-1 blocks are
	ep.block_shift = skip_root == 0 ? 0u : (skip_root == 1 ? 1u : 1u + (uint32_t)b->h_counts[1]);
	ep.ptr_bias = bias;
	b->top_blocks = ep.block_shift;
	if (skip_root) {
		SVO_TRY(b->root_scratch.reserve(9 * 8, s));
		ep.root_dst = b->root_scratch.p;
	}
	if (b->h_counts[b->level] == 0) { // empty scene: a zeroed root block
		SVO_CUDA_TRY(cudaMemsetAsync(skip_root ? b->root_scratch.p : d_dst, 0, 8 * sizeof(uint32_t), s));
	} else {
		auto k_direct = k_emit_octree<false>;
		auto k_staged = k_emit_octree<true>;
		// brick path: the upper windows (small) are written on the builder's second stream, next to k_brick_emit
		cudaStream_t su = s;
		if (b->path == 1 && b->n_bricks) {
			SVO_TRY(ensure_aux_stream(b));
			su = b->aux;
			SVO_CUDA_TRY(cudaEventRecord(b->ev_fork, s));
			SVO_CUDA_TRY(cudaStreamWaitEvent(su, b->ev_fork, 0));
		}
		if (skip_root) // stitched into another (usually a peer GPU's) buffer
			SVO_LAUNCH(div_up(ep.total_blocks, EMITO_BLOCK), EMITO_BLOCK, 0, su, k_staged, ep, d_dst);
		else
			SVO_LAUNCH(div_up(ep.total_blocks, EMITO_BLOCK), EMITO_BLOCK, 0, su, k_direct, ep, d_dst);
		if (su != s) SVO_CUDA_TRY(cudaEventRecord(b->ev_join, su));
		if (b->path == 1) {
			BrickEmit be{ep.block_base[b->level - 1], ep.block_base[b->level], ep.block_shift, ep.ptr_bias, b->n_bricks, brick_parts};
			if (plan) plan[0] = be.n_bricks, plan[1] = be.block_l1, plan[2] = be.block_l, plan[3] = (uint64_t)be.block_shift | ((uint64_t)be.ptr_bias << 32);
			SVO_CUDA_TRY(cudaEventRecord(b->ev_brick[3], s));
			if (b->n_bricks)
				SVO_LAUNCH_INDEP(div_up(b->n_bricks * BRICK_EMIT_LANES, BRICK_BLOCK), BRICK_BLOCK, s, k_brick_emit, b->brick_args, be, d_dst);
			SVO_CUDA_TRY(cudaEventRecord(b->ev_brick[4], s));
		}
		if (su != s) SVO_CUDA_TRY(cudaStreamWaitEvent(s, b->ev_join, 0));
	}
	SVO_CUDA_TRY(cudaGetLastError());
	return SVO_OK;
}

int svo_builder_emit_to(svo_builder *b, uint32_t *d_dst, uint32_t pointer_bias_words, int skip_root, void *stream) {
	if (!b || !d_dst) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_emit_to: null argument");
	if (!b->prepared) return fail(SVO_ERR_NOT_READY, "svo_builder_emit_to: prepare first");
	if (skip_root < 0 || skip_root > 2 || (skip_root == 2 && b->level < 3))
		return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_emit_to: skip_root is 0, 1 or 2 (2 needs level >= 3)");
	if ((uint64_t)pointer_bias_words + b->range_bytes / 4 >= (1ull << 30))
		return fail(SVO_ERR_CAPACITY, "biased child pointers would exceed 30 bits (octree.glsl:110)");
	DeviceGuard guard(b->device);
	b->last_stream = (cudaStream_t)stream;
	SVO_TRY(emit_into(b, d_dst, pointer_bias_words, skip_root, (cudaStream_t)stream));
	SVO_CUDA_TRY(cudaEventRecord(b->ev[5], (cudaStream_t)stream)); // svo_builder_last_ms covers prepare + emit_to as well
	b->emitted = true;
	return SVO_OK;
}

// ---- compact gather (multi-GPU, brick path): see svo.h ----
uint64_t svo_builder_compact_bytes(const svo_builder *b) {
	return (b && b->prepared && b->path == 1 && b->h_counts[b->level]) ? b->n_bricks * 32ull : 0;
}

static int compact_args_ok(svo_builder *b, const char *who, uint32_t pointer_bias_words, int skip_root) {
	int rc = SVO_OK;
	const char *why = nullptr;
	if (!b->prepared) rc = SVO_ERR_NOT_READY, why = "prepare first";
	else if (!svo_builder_compact_bytes(b)) rc = SVO_ERR_UNSUPPORTED, why = "the build did not take the brick path (use svo_builder_emit_to)";
	else if (skip_root < 0 || skip_root > 2 || (skip_root == 2 && b->level < 3)) rc = SVO_ERR_INVALID_ARGUMENT, why = "skip_root is 0, 1 or 2 (2 needs level >= 3)";
	else if ((uint64_t)pointer_bias_words + b->range_bytes / 4 >= (1ull << 30))
		rc = SVO_ERR_CAPACITY, why = "biased child pointers would exceed 30 bits (octree.glsl:110)";
	if (rc != SVO_OK) set_error("%s: %s", who, why);
	return rc;
}

// records and ranks -> d_tables (three copies: 16 + 8 + 8 bytes per brick), and the plan words of svo_expand_compact
static int push_tables(svo_builder *b, uint32_t bias, int skip_root, void *d_tables, uint64_t plan[4], cudaStream_t s) {
	const uint32_t L = b->level;
	const uint32_t shift = skip_root == 0 ? 0u : (skip_root == 1 ? 1u : 1u + (uint32_t)b->h_counts[1]);
	plan[0] = b->n_bricks, plan[1] = b->ep.block_base[L - 1], plan[2] = b->ep.block_base[L], plan[3] = (uint64_t)shift | ((uint64_t)bias << 32);
	const uint64_t n = b->n_bricks;
	char *t = static_cast<char *>(d_tables);
	SVO_CUDA_TRY(cudaMemcpyAsync(t, b->brick_args.rec, n * 16, cudaMemcpyDefault, s));
	SVO_CUDA_TRY(cudaMemcpyAsync(t + n * 16, b->brick_args.rank[1], n * 8, cudaMemcpyDefault, s));
	SVO_CUDA_TRY(cudaMemcpyAsync(t + n * 24, b->brick_args.rank[2], n * 8, cudaMemcpyDefault, s));
	return SVO_OK;
}

int svo_builder_push_tables(svo_builder *b, uint32_t pointer_bias_words, int skip_root, void *d_tables, uint64_t plan[4], void *stream) {
	if (!b || !d_tables || !plan) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_push_tables: null argument");
	SVO_TRY(compact_args_ok(b, "svo_builder_push_tables", pointer_bias_words, skip_root));
	DeviceGuard guard(b->device);
	b->last_stream = (cudaStream_t)stream;
	return push_tables(b, pointer_bias_words, skip_root, d_tables, plan, (cudaStream_t)stream);
}

int svo_builder_emit_compact_to(svo_builder *b, uint32_t *d_dst, uint32_t pointer_bias_words, int skip_root, void *d_tables, uint64_t plan[4],
                                void *stream) {
	if (!b || !d_dst || (d_tables && !plan)) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_emit_compact_to: null argument");
	SVO_TRY(compact_args_ok(b, "svo_builder_emit_compact_to", pointer_bias_words, skip_root));
	DeviceGuard guard(b->device);
	cudaStream_t s = (cudaStream_t)stream;
	b->last_stream = s;
	// records and ranks to the tables (first: the owner of the buffer can start on them while the rest is still crossing);
	// the upper windows and the leaf blocks of the rasterized bricks go to their final places
	if (d_tables) SVO_TRY(push_tables(b, pointer_bias_words, skip_root, d_tables, plan, s));
	SVO_TRY(emit_into(b, d_dst, pointer_bias_words, skip_root, s, BRICK_EMIT_COPY, nullptr));
	SVO_CUDA_TRY(cudaEventRecord(b->ev[5], s));
	b->emitted = true;
	return SVO_OK;
}

int svo_expand_compact(int device, const void *d_tables, const uint64_t plan[4], uint32_t *d_dst, void *stream) {
	if (!d_tables || !plan || !d_dst) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_expand_compact: null argument");
	const uint64_t n = plan[0];
	if (!n) return SVO_OK;
	if (reinterpret_cast<uintptr_t>(d_tables) % 16) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_expand_compact: the tables must be 16-byte aligned");
	DeviceGuard guard(device);
	const char *t = static_cast<const char *>(d_tables);
	BrickArgs a{};
	a.rec = reinterpret_cast<uint4 *>(const_cast<char *>(t));
	a.rank[1] = reinterpret_cast<const uint64_t *>(t + n * 16);
	a.rank[2] = reinterpret_cast<const uint64_t *>(t + n * 24);
	BrickEmit be{plan[1], plan[2], (uint32_t)plan[3], (uint32_t)(plan[3] >> 32), n, BRICK_EMIT_FLAT | BRICK_EMIT_PTRS};
	cudaStream_t s = (cudaStream_t)stream;
	SVO_LAUNCH_INDEP(div_up(n * BRICK_EMIT_LANES, BRICK_BLOCK), BRICK_BLOCK, s, k_brick_emit, a, be, d_dst);
	SVO_CUDA_TRY(cudaGetLastError());
	return SVO_OK;
}

int svo_builder_top_words(svo_builder *b, uint32_t out[72], uint32_t *n_blocks, void *stream) {
	if (!b || !out || !n_blocks) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!b->root_scratch.p || !b->top_blocks) return fail(SVO_ERR_NOT_READY, "svo_builder_top_words: emit with skip_root first");
	DeviceGuard guard(b->device);
	*n_blocks = b->top_blocks;
	SVO_CUDA_TRY(cudaMemcpyAsync(out, b->root_scratch.p, (size_t)b->top_blocks * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	SVO_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	return SVO_OK;
}

int svo_builder_root_words(svo_builder *b, uint32_t out[8], void *stream) {
	if (!b || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!b->root_scratch.p) return fail(SVO_ERR_NOT_READY, "svo_builder_root_words: emit with skip_root first");
	DeviceGuard guard(b->device);
	SVO_CUDA_TRY(cudaMemcpyAsync(out, b->root_scratch.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	SVO_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	return SVO_OK;
}

int svo_builder_build(svo_builder *b, void *stream) {
	SVO_TRY(svo_builder_prepare(b, stream));
	DeviceGuard guard(b->device);
	cudaStream_t s = (cudaStream_t)stream;
	SVO_TRY(b->octree.reserve(b->range_bytes / 4, s));
	SVO_TRY(emit_into(b, b->octree.p, 0, 0, s));
	SVO_CUDA_TRY(cudaEventRecord(b->ev[5], s));
	b->built = true;
	return SVO_OK;
}

uint32_t svo_builder_level(const svo_builder *b) { return b ? b->level : 0; }
uint64_t svo_builder_octree_range_bytes(const svo_builder *b) { return (b && (b->built || b->prepared)) ? b->range_bytes : 0; }
const uint32_t *svo_builder_octree(const svo_builder *b) { return (b && b->built) ? b->octree.p : nullptr; }
uint64_t svo_builder_leaf_count(const svo_builder *b) { return (b && (b->built || b->prepared)) ? b->h_counts[b->level] : 0; }
int svo_builder_level_counts(const svo_builder *b, uint64_t *out, uint32_t n_out) {
	if (!b || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!b->built && !b->prepared) return fail(SVO_ERR_NOT_READY, "build first");
	for (uint32_t d = 0; d < n_out; ++d) out[d] = d <= b->level ? b->h_counts[d] : 0;
	return SVO_OK;
}

int svo_builder_rebase_copy(const svo_builder *b, uint32_t *d_dst, uint64_t dst_word_offset, uint32_t base_words, void *stream) {
	if (!b || !d_dst) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!b->built) return fail(SVO_ERR_NOT_READY, "build first");
	if ((dst_word_offset & 3) || (base_words & 7)) return fail(SVO_ERR_INVALID_ARGUMENT, "offsets must keep 8-word block alignment");
	DeviceGuard guard(b->device);
	cudaStream_t s = (cudaStream_t)stream;
	const uint64_t n_vec = b->range_bytes / 16;
	SVO_LAUNCH_INDEP(div_up(n_vec, 256), 256, s, k_rebase_copy, reinterpret_cast<const uint4 *>(b->octree.p),
	                 reinterpret_cast<uint4 *>(d_dst + dst_word_offset), n_vec, base_words);
	SVO_CUDA_TRY(cudaGetLastError());
	return SVO_OK;
}

int svo_builder_last_ms(svo_builder *b, float *phase_ms, uint32_t *sort_passes) {
	if (!b || !phase_ms) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!b->built && !b->emitted) return fail(SVO_ERR_NOT_READY, "build (or prepare + emit_to) first");
	DeviceGuard guard(b->device);
	SVO_CUDA_TRY(cudaEventSynchronize(b->ev[5]));
	phase_ms[SVO_PHASE_RASTER] = 0.f;
	if (b->vox->t_raster.recorded) SVO_CUDA_TRY(cudaEventElapsedTime(&phase_ms[SVO_PHASE_RASTER], b->vox->t_raster.a, b->vox->t_raster.b));
	for (int i = 0; i < 5; ++i) SVO_CUDA_TRY(cudaEventElapsedTime(&phase_ms[SVO_PHASE_SORT_HIST + i], b->ev[i], b->ev[i + 1]));
	if (sort_passes) *sort_passes = b->sort_passes;
	return SVO_OK;
}

// ------------------------------------------------------------------------------------------------------
int svo_builder_export_fd(svo_builder *b, int *fd, uint64_t *alloc_size, const uint32_t **d_ptr, void *stream) {
	if (!b || !fd || !alloc_size) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_builder_export_fd: null argument");
	if (!b->prepared && !b->built) return fail(SVO_ERR_NOT_READY, "svo_builder_export_fd: prepare or build first");
#ifdef SVO_EMU
	(void)d_ptr, (void)stream;
	return fail(SVO_ERR_UNSUPPORTED, "no external memory in the emulation build");
#else
	DeviceGuard guard(b->device);
	cudaStream_t s = (cudaStream_t)stream;
	b->last_stream = s;
	const Vmm &v = vmm();
	if (!v.ok) return fail(SVO_ERR_UNSUPPORTED, "the CUDA driver lacks the virtual memory management entry points");
	release_export(b);
	const CUmemAllocationProp prop = export_prop(b->device);
	size_t gran = 0;
	if (v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
		return fail(SVO_ERR_CUDA, "cuMemGetAllocationGranularity failed");
	const size_t size = (b->range_bytes + gran - 1) / gran * gran;
	CUmemGenericAllocationHandle h = 0;
	CUdeviceptr va = 0;
	if (v.create(&h, size, &prop, 0) != CUDA_SUCCESS) return fail(SVO_ERR_CUDA, "cuMemCreate (exportable allocation) failed");
	if (v.reserve(&va, size, 0, 0, 0) != CUDA_SUCCESS) {
		v.release(h);
		return fail(SVO_ERR_CUDA, "cuMemAddressReserve failed");
	}
	CUmemAccessDesc acc{};
	acc.location = prop.location;
	acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	if (v.map(va, size, 0, h, 0) != CUDA_SUCCESS || v.set_access(va, size, &acc, 1) != CUDA_SUCCESS) {
		v.release(h);
		v.address_free(va, size);
		return fail(SVO_ERR_CUDA, "cuMemMap / cuMemSetAccess failed");
	}
	b->export_va = va, b->export_handle = h, b->export_size = size;
	uint32_t *dst = reinterpret_cast<uint32_t *>(va);
	if (b->built) // already emitted into the builder's own buffer: one device copy
		SVO_CUDA_TRY(cudaMemcpyAsync(dst, b->octree.p, b->range_bytes, cudaMemcpyDeviceToDevice, s));
	else { // prepared: the emit kernel writes the exported memory directly
		SVO_TRY(emit_into(b, dst, 0, 0, s));
		SVO_CUDA_TRY(cudaEventRecord(b->ev[5], s));
		b->emitted = true;
	}
	if (size > b->range_bytes) SVO_CUDA_TRY(cudaMemsetAsync(reinterpret_cast<unsigned char *>(dst) + b->range_bytes, 0, size - b->range_bytes, s));
	SVO_CUDA_TRY(cudaStreamSynchronize(s)); // the importer may read as soon as it holds the descriptor
	int out_fd = -1;
	if (v.export_handle(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || out_fd < 0)
		return fail(SVO_ERR_CUDA, "cuMemExportToShareableHandle failed");
	*fd = out_fd, *alloc_size = size;
	if (d_ptr) *d_ptr = dst;
	return SVO_OK;
#endif
}

int svo_external_memory_import_fd(int device, int fd, uint64_t size, void **import_handle, void **d_ptr) {
	if (!import_handle || !d_ptr || fd < 0 || size == 0) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_external_memory_import_fd: bad argument");
#ifdef SVO_EMU
	(void)device;
	return fail(SVO_ERR_UNSUPPORTED, "no external memory in the emulation build");
#else
	DeviceGuard guard(device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "cudaSetDevice failed");
	ImportedMemory *im = new (std::nothrow) ImportedMemory();
	if (!im) return fail(SVO_ERR_CUDA, "out of host memory");
	im->size = size;
	// what a CUDA consumer of a Vulkan allocation does (and the mirror image of vkImportMemoryFdKHR on this descriptor)
	cudaExternalMemoryHandleDesc hd{};
	hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
	hd.handle.fd = dup(fd); // a successful import owns its descriptor; keep ours for the second path
	hd.size = size;
	void *ptr = nullptr;
	if (cudaImportExternalMemory(&im->ext, &hd) == cudaSuccess) {
		cudaExternalMemoryBufferDesc bd{};
		bd.offset = 0, bd.size = size;
		if (cudaExternalMemoryGetMappedBuffer(&ptr, im->ext, &bd) == cudaSuccess) {
			im->path = 1;
			close(fd);
			*import_handle = im, *d_ptr = ptr;
			return 1;
		}
		cudaDestroyExternalMemory(im->ext);
		im->ext = nullptr;
	} else
		close(hd.handle.fd);
	(void)cudaGetLastError();
	const Vmm &v = vmm();
	CUmemAccessDesc acc{};
	acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE, acc.location.id = device;
	acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	if (v.ok && v.import_handle(&im->handle, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS &&
	    v.reserve(&im->va, size, 0, 0, 0) == CUDA_SUCCESS && v.map(im->va, size, 0, im->handle, 0) == CUDA_SUCCESS &&
	    v.set_access(im->va, size, &acc, 1) == CUDA_SUCCESS) {
		im->path = 2;
		close(fd);
		*import_handle = im, *d_ptr = reinterpret_cast<void *>(im->va);
		return 2;
	}
	delete im;
	return fail(SVO_ERR_CUDA, "neither cudaImportExternalMemory nor cuMemImportFromShareableHandle accepted the descriptor");
#endif
}
int svo_external_memory_release(int device, void *import_handle) {
	if (!import_handle) return SVO_OK;
#ifdef SVO_EMU
	(void)device;
	return SVO_OK;
#else
	DeviceGuard guard(device);
	ImportedMemory *im = static_cast<ImportedMemory *>(import_handle);
	if (im->path == 1)
		cudaDestroyExternalMemory(im->ext);
	else if (im->path == 2) {
		const Vmm &v = vmm();
		v.unmap(im->va, im->size);
		v.release(im->handle);
		v.address_free(im->va, im->size);
	}
	delete im;
	return SVO_OK;
#endif
}

// ------------------------------------------------------------------------------------------------------
int svo_sort_u64(uint64_t *d_keys, uint64_t *d_tmp, uint64_t n, uint32_t begin_bit, uint32_t end_bit, int device, void *stream) {
	if ((n && (!d_keys || !d_tmp)) || begin_bit > end_bit || end_bit > 64) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_sort_u64: bad argument");
	DeviceGuard guard(device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "cudaSetDevice failed");
	SVO_TRY(configure_device_pool(device));
	cudaStream_t s = (cudaStream_t)stream;
	SortScratch sc;
	uint64_t *res = d_keys;
	uint32_t np = 0;
	int rc = radix_sort_u64(d_keys, d_tmp, n, begin_bit, end_bit, sc, device, sm_count(device), s, &res, &np, nullptr);
	if (!rc && res != d_keys && cudaMemcpyAsync(d_keys, res, n * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = fail(SVO_ERR_CUDA, "copy back failed");
	sc.release(s);
	return rc;
}

static_assert(sizeof(svo_ray_hit) == sizeof(RayHit), "svo_ray_hit layout");
int svo_octree_raymarch_leaf(int device, const uint32_t *d_octree, uint64_t n_rays, const float *d_origins, const float *d_dirs,
                             svo_ray_hit *d_hits, void *stream) {
	if (n_rays && (!d_octree || !d_origins || !d_dirs || !d_hits)) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_octree_raymarch_leaf: null argument");
	if (n_rays >= (1ull << 38)) return fail(SVO_ERR_CAPACITY, "svo_octree_raymarch_leaf: too many rays");
	DeviceGuard guard(device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "cudaSetDevice failed");
	if (n_rays == 0) return SVO_OK;
	SVO_LAUNCH_INDEP(div_up(n_rays, 128), 128, (cudaStream_t)stream, k_raymarch_leaf, d_octree, n_rays, d_origins, d_dirs,
	                 reinterpret_cast<RayHit *>(d_hits));
	SVO_CUDA_TRY(cudaGetLastError());
	return SVO_OK;
}

int svo_device_malloc(int device, uint64_t bytes, void **out) {
	if (!out) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	DeviceGuard guard(device);
	if (!guard.ok) return fail(SVO_ERR_CUDA, "cudaSetDevice failed");
	SVO_CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
	return SVO_OK;
}
int svo_device_free(int device, void *ptr) {
	DeviceGuard guard(device);
	SVO_CUDA_TRY(cudaFree(ptr));
	return SVO_OK;
}
int svo_memcpy_h2d(int device, void *d_dst, const void *h_src, uint64_t bytes, void *stream) {
	DeviceGuard guard(device);
	SVO_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
	return SVO_OK;
}
int svo_memcpy_d2h(int device, void *h_dst, const void *d_src, uint64_t bytes, void *stream) {
	DeviceGuard guard(device);
	SVO_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	SVO_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	return SVO_OK;
}
int svo_memcpy_d2d(int device, void *d_dst, const void *d_src, uint64_t bytes, void *stream) {
	DeviceGuard guard(device);
	SVO_CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
	return SVO_OK;
}
#ifndef SVO_EMU
int svo_ipc_export(int device, void *d_ptr, unsigned char handle[SVO_IPC_HANDLE_BYTES]) {
	static_assert(sizeof(cudaIpcMemHandle_t) == SVO_IPC_HANDLE_BYTES, "handle size");
	if (!d_ptr || !handle) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	DeviceGuard guard(device);
	cudaIpcMemHandle_t h;
	SVO_CUDA_TRY(cudaIpcGetMemHandle(&h, d_ptr));
	memcpy(handle, &h, sizeof(h));
	return SVO_OK;
}
int svo_ipc_open(int device, const unsigned char handle[SVO_IPC_HANDLE_BYTES], void **out) {
	if (!out || !handle) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	DeviceGuard guard(device);
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	SVO_CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
	return SVO_OK;
}
int svo_ipc_close(int device, void *d_ptr) {
	DeviceGuard guard(device);
	SVO_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
	return SVO_OK;
}
#else
int svo_ipc_export(int, void *, unsigned char *) { return fail(SVO_ERR_UNSUPPORTED, "no IPC in the emulation build"); }
int svo_ipc_open(int, const unsigned char *, void **) { return fail(SVO_ERR_UNSUPPORTED, "no IPC in the emulation build"); }
int svo_ipc_close(int, void *) { return fail(SVO_ERR_UNSUPPORTED, "no IPC in the emulation build"); }
#endif
int svo_stream_synchronize(int device, void *stream) {
	DeviceGuard guard(device);
	SVO_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	return SVO_OK;
}

void svo_debug_force_wide_sort_state(int on) { svo::g_force_wide_sort_state = on != 0; }
void svo_debug_profile_passes(int on) { svo::g_profile_passes = on != 0; }
void svo_debug_set_build_path(int mode) { g_build_path = mode < 0 ? -1 : (mode > 0 ? 1 : 0); }
int svo_builder_build_path(const svo_builder *b) { return b ? b->path : 0; }
int svo_builder_brick_stats(svo_builder *b, uint64_t counts[4], float ms[4]) {
	if (!b || !counts || !ms) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	if (!(b->built || b->emitted) || b->path != 1) return fail(SVO_ERR_NOT_READY, "svo_builder_brick_stats: no brick build");
	DeviceGuard guard(b->device);
	counts[0] = b->n_pairs, counts[1] = b->n_bricks, counts[2] = b->n_small_leaves, counts[3] = b->n_slow;
	SVO_CUDA_TRY(cudaEventSynchronize(b->ev_brick[4]));
	SVO_CUDA_TRY(cudaEventElapsedTime(&ms[0], b->ev_brick[0], b->ev_brick[1]));
	SVO_CUDA_TRY(cudaEventElapsedTime(&ms[1], b->ev_brick[1], b->ev_brick[2]));
	SVO_CUDA_TRY(cudaEventElapsedTime(&ms[2], b->ev_brick[2], b->ev[3]));
	SVO_CUDA_TRY(cudaEventElapsedTime(&ms[3], b->ev_brick[3], b->ev_brick[4]));
	return SVO_OK;
}
#if SVO_OS_CLOCKS
// experiment builds only (not declared in svo.h): per-phase cycle sums of the onesweep tiles; reset != 0 clears them
SVO_API int svo_debug_onesweep_clocks(unsigned long long out[12], int reset) {
	if (cudaMemcpyFromSymbol(out, svo::g_os_clocks, sizeof(unsigned long long) * 8) != cudaSuccess) return -2;
	if (cudaMemcpyFromSymbol(out + 8, svo::g_os_walk, sizeof(unsigned long long) * 4) != cudaSuccess) return -2;
	if (reset) {
		unsigned long long z[8] = {};
		if (cudaMemcpyToSymbol(svo::g_os_clocks, z, sizeof(z)) != cudaSuccess) return -2;
		if (cudaMemcpyToSymbol(svo::g_os_walk, z, sizeof(unsigned long long) * 4) != cudaSuccess) return -2;
	}
	return 0;
}
#endif
int svo_builder_sort_step_ms(svo_builder *b, float *out, uint32_t cap) {
	if (!b || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	DeviceGuard guard(b->device);
	const SortScratch &sc = b->sort_scratch;
	int n = 0;
	for (int i = 0; i + 1 < sc.n_ev && (uint32_t)n < cap; ++i, ++n) {
		SVO_CUDA_TRY(cudaEventSynchronize(sc.ev[i + 1]));
		SVO_CUDA_TRY(cudaEventElapsedTime(&out[n], sc.ev[i], sc.ev[i + 1]));
	}
	return n;
}

#ifdef SVO_EMU
// test hook of the emulation build only (see scan.cuh)
SVO_API void svo_emu_set_lookback_aggregate_only(int on) { svo::g_emu_lookback_aggregate_only = on; }
#endif

} // extern "C"

#include "sharded.inl"
