// sharded.inl -- svo_build_sharded: the octant-sharded build over several GPUs of ONE process (included by svo_b200.cu).
//
// The reference builds on a single VkDevice from one loader thread (src/LoaderThread.cpp:51-89); a C++ host that wants
// the multi-GPU build calls this instead of Voxelizer::Create / OctreeBuilder::Create / CmdVoxelize / CmdBuild.  Same
// plan as sparsevoxeloctree_b200/sharded.py (one process per GPU over torch.distributed, what bench.py drives), minus
// the collectives: with all devices in one address space the sizes are exchanged in host memory.
//   levels <= 13 ("slab" mode): device k builds the voxel window made of its octants in global coordinates
//       (svo_voxelizer_create_windowed + svo_builder_prepare), then emits its node words, child pointers already final,
//       straight into the stitched buffer on devices[0] over NVLink peer access (svo_builder_emit_to, skip_root).  A slab
//       built on the brick path crosses in compact form (svo_builder_emit_compact_to: upper windows, the rasterized
//       bricks' leaf blocks and 32 bytes per brick) and devices[0] generates the rest itself (svo_expand_compact).
//   level 14 ("octant" mode; 42 Morton bits + 24 colour bits do not fit a 64-bit fragment): one cube-local level-13
//       build per octant, round-robin over the devices; svo_builder_rebase_copy adds the subtree's base to every child
//       pointer while storing into the stitched buffer.
// One host thread per device runs that device's work, so the devices' host-side waits (the count pass and the size
// read-back of every build) overlap.

namespace {

struct ShardPart {
	int device = 0;
	uint32_t octant = 0; // octant mode: which child of the root
	svo_scene *scene = nullptr;
	svo_voxelizer *vox = nullptr;
	svo_builder *builder = nullptr;
	uint64_t body_words = 0; // node words this part contributes behind the root block
	uint64_t base_words = 0; // where they go in the stitched buffer
	uint64_t compact_bytes = 0, stage_off = 0, plan[4] = {}; // slab mode, brick path, not on devices[0]: the compact gather
	int rc = SVO_OK;
	char err[256] = "";
};

void part_fail(ShardPart &p, int rc) {
	p.rc = rc;
	snprintf(p.err, sizeof(p.err), "%s", svo_last_error()); // the failing call ran on this thread
}

} // namespace

struct svo_sharded {
	std::vector<int> devices;
	std::vector<cudaStream_t> streams; // one per device
	std::vector<std::vector<ShardPart>> parts; // per device
	uint32_t level = 0;
	bool slab = false;
	uint32_t *octree = nullptr; // on devices[0] (cudaMalloc: peers write into it)
	uint64_t capacity_words = 0, total_words = 0;
	char *stage = nullptr; // on devices[0]: the per-brick tables of the other devices' slabs (compact gather)
	uint64_t stage_capacity = 0;
	uint64_t n_frag = 0, n_leaf = 0;
	float last_ms = 0.f;
};

template <class F> static int for_each_device(svo_sharded *sh, F fn) {
#ifdef SVO_EMU
	for (size_t k = 0; k < sh->devices.size(); ++k) fn(k); // (the kernel emulator runs one launch at a time)
#else
	std::vector<std::thread> th;
	for (size_t k = 0; k < sh->devices.size(); ++k) th.emplace_back([&, k]() { fn(k); });
	for (auto &t : th) t.join();
#endif
	for (auto &dev_parts : sh->parts)
		for (ShardPart &p : dev_parts)
			if (p.rc != SVO_OK) {
				set_error("svo_build_sharded (device %d): %s", p.device, p.err);
				return p.rc;
			}
	return SVO_OK;
}

static int sharded_run(svo_sharded *sh) {
	const auto t0 = std::chrono::steady_clock::now();
	// phase 1: every device voxelizes and builds (slab mode: up to the size read-back only)
	SVO_TRY(for_each_device(sh, [&](size_t k) {
		for (ShardPart &p : sh->parts[k]) {
			void *s = sh->streams[k];
			int rc = svo_voxelizer_voxelize(p.vox, s);
			if (rc == SVO_OK) rc = sh->slab ? svo_builder_prepare(p.builder, s) : svo_builder_build(p.builder, s);
			if (rc != SVO_OK) {
				part_fail(p, rc);
				return;
			}
			const uint64_t words = svo_builder_octree_range_bytes(p.builder) / 4;
			const bool empty = svo_builder_leaf_count(p.builder) == 0;
			p.body_words = empty ? 0 : (sh->slab ? words - 8 : words); // slab: the part's own root block is merged, not copied
		}
	}));
	// plan: root block, then the parts' bodies in device order (slab) / octant order
	std::vector<ShardPart *> order;
	for (auto &dev_parts : sh->parts)
		for (ShardPart &p : dev_parts) order.push_back(&p);
	if (!sh->slab) std::sort(order.begin(), order.end(), [](const ShardPart *a, const ShardPart *b) { return a->octant < b->octant; });
	uint64_t run = 8;
	sh->n_frag = sh->n_leaf = 0;
	for (ShardPart *p : order) {
		p->base_words = run;
		run += p->body_words;
		sh->n_frag += svo_voxelizer_fragment_count(p->vox);
		sh->n_leaf += svo_builder_leaf_count(p->builder);
	}
	if (run >= (1ull << 30)) return fail(SVO_ERR_CAPACITY, "stitched octree needs >= 2^30 words: 30-bit child pointers cannot address it");
	sh->total_words = run;
	uint64_t stage_run = 0;
	for (size_t k = 0; k < sh->parts.size(); ++k)
		for (ShardPart &p : sh->parts[k]) {
			p.compact_bytes = (sh->slab && k != 0 && p.body_words) ? (svo_builder_compact_bytes(p.builder) + 255) / 256 * 256 : 0;
			p.stage_off = stage_run;
			stage_run += p.compact_bytes;
		}
	{
		DeviceGuard guard(sh->devices[0]);
		if (stage_run > sh->stage_capacity) {
			if (sh->stage) cudaFree(sh->stage);
			sh->stage = nullptr;
			sh->stage_capacity = stage_run + stage_run / 4;
			SVO_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&sh->stage), sh->stage_capacity));
		}
		if (run > sh->capacity_words) {
			if (sh->octree) cudaFree(sh->octree);
			sh->octree = nullptr;
			sh->capacity_words = run + run / 4;
			SVO_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&sh->octree), sh->capacity_words * sizeof(uint32_t)));
		}
	}
	// phase 2: every device stores its node words into the stitched buffer (peer stores over NVLink for devices != devices[0])
	uint32_t *const dst = sh->octree;
	SVO_TRY(for_each_device(sh, [&](size_t k) {
		for (ShardPart &p : sh->parts[k]) {
			if (!p.body_words) continue;
			void *s = sh->streams[k];
			const int rc = !sh->slab          ? svo_builder_rebase_copy(p.builder, dst, p.base_words, (uint32_t)p.base_words, s)
			               : p.compact_bytes ? svo_builder_emit_compact_to(p.builder, dst + p.base_words, (uint32_t)p.base_words, 1,
			                                                               sh->stage + p.stage_off, p.plan, s)
			                                 : svo_builder_emit_to(p.builder, dst + p.base_words, (uint32_t)p.base_words, 1, s);
			if (rc != SVO_OK) {
				part_fail(p, rc);
				return;
			}
		}
		if (svo_stream_synchronize(sh->devices[k], sh->streams[k]) != SVO_OK && !sh->parts[k].empty()) part_fail(sh->parts[k][0], SVO_ERR_CUDA);
	}));
	// compact gather: the tables have arrived (every stream was synchronised above); devices[0] writes the flat bricks'
	// leaf blocks and the pointer blocks of the two deepest windows of the other devices' slabs
	for (ShardPart *p : order)
		if (p->compact_bytes) SVO_TRY(svo_expand_compact(sh->devices[0], sh->stage + p->stage_off, p->plan, dst + p->base_words, sh->streams[0]));
	SVO_TRY(svo_stream_synchronize(sh->devices[0], sh->streams[0]));
	// the root block: slab mode merges the parts' root blocks (disjoint octants: a sum); octant mode points at the subtrees
	uint32_t root[8] = {};
	for (ShardPart *p : order) {
		if (!p->body_words) continue;
		if (sh->slab) {
			uint32_t r[8];
			SVO_TRY(svo_builder_root_words(p->builder, r, nullptr)); // (the emitting stream was synchronised above)
			for (int i = 0; i < 8; ++i) root[i] += r[i];
		} else
			root[p->octant] = 0x80000000u | (uint32_t)p->base_words;
	}
	{
		DeviceGuard guard(sh->devices[0]);
		SVO_CUDA_TRY(cudaMemcpy(dst, root, sizeof(root), cudaMemcpyHostToDevice));
	}
	sh->last_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return SVO_OK;
}

extern "C" {

void svo_sharded_destroy(svo_sharded *sh) {
	if (!sh) return;
	for (size_t k = 0; k < sh->parts.size(); ++k)
		for (ShardPart &p : sh->parts[k]) {
			svo_builder_destroy(p.builder);
			svo_voxelizer_destroy(p.vox);
			svo_scene_destroy(p.scene);
		}
	for (size_t k = 0; k < sh->streams.size(); ++k) {
		DeviceGuard guard(sh->devices[k]);
		if (sh->streams[k]) {
			cudaStreamSynchronize(sh->streams[k]);
			cudaStreamDestroy(sh->streams[k]);
		}
	}
	if (sh->octree || sh->stage) {
		DeviceGuard guard(sh->devices[0]);
		if (sh->octree) cudaFree(sh->octree);
		if (sh->stage) cudaFree(sh->stage);
	}
	delete sh;
}

int svo_build_sharded(const svo_mesh *mesh, uint32_t level, int mode, const int *devices, uint32_t n_devices, svo_sharded **out) {
	if (!mesh || !devices || !out) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_build_sharded: null argument");
	*out = nullptr;
	if (n_devices != 1 && n_devices != 2 && n_devices != 4 && n_devices != 8)
		return fail(SVO_ERR_INVALID_ARGUMENT, "svo_build_sharded: 1, 2, 4 or 8 devices (the x / xy / xyz split of the root's octants)");
	if (level < 2 || level > 14) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_build_sharded: level must be 2..14");
	if (mesh->on_device) return fail(SVO_ERR_INVALID_ARGUMENT, "svo_build_sharded: the mesh must be in host memory (it is uploaded to every device)");
	svo_sharded *sh = new (std::nothrow) svo_sharded();
	if (!sh) return fail(SVO_ERR_CUDA, "out of host memory");
	sh->level = level;
	sh->slab = 3 * level + 24 <= 64;
	sh->devices.assign(devices, devices + n_devices);
	sh->streams.assign(n_devices, nullptr);
	sh->parts.resize(n_devices);
	int rc = SVO_OK;
	for (uint32_t k = 0; k < n_devices && rc == SVO_OK; ++k) {
		DeviceGuard guard(devices[k]);
		if (!guard.ok || cudaStreamCreateWithFlags(&sh->streams[k], cudaStreamNonBlocking) != cudaSuccess) {
			rc = fail(SVO_ERR_CUDA, "svo_build_sharded: cannot use one of the devices");
			break;
		}
		if (devices[k] != devices[0]) { // the stitched buffer lives on devices[0]: peers store into it
			int can = 0;
			cudaDeviceCanAccessPeer(&can, devices[k], devices[0]);
			if (!can) rc = fail(SVO_ERR_UNSUPPORTED, "svo_build_sharded: no peer access to devices[0]");
			else {
				const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(SVO_ERR_CUDA, "cudaDeviceEnablePeerAccess failed");
				(void)cudaGetLastError();
			}
		}
	}
	// the parts: slab mode one window per device, octant mode the 8 octants round-robin
	const uint32_t res = 1u << level, half = res >> 1;
	const uint32_t axes = n_devices == 1 ? 0 : (n_devices == 2 ? 1 : (n_devices == 4 ? 2 : 3));
	for (uint32_t k = 0; k < n_devices && rc == SVO_OK; ++k) {
		if (sh->slab) {
			ShardPart p;
			p.device = devices[k];
			sh->parts[k].push_back(p);
		} else
			for (uint32_t o = k; o < 8; o += n_devices) {
				ShardPart p;
				p.device = devices[k], p.octant = o;
				sh->parts[k].push_back(p);
			}
	}
	if (rc == SVO_OK)
		rc = for_each_device(sh, [&](size_t k) {
			void *s = sh->streams[k];
			svo_scene *scene = nullptr; // one upload per device, shared by the device's parts
			for (ShardPart &p : sh->parts[k]) {
				int r = SVO_OK;
				if (!scene) r = svo_scene_create(mesh, p.device, s, &scene);
				if (r == SVO_OK && &p == &sh->parts[k][0]) p.scene = scene; // the first part owns it
				if (r == SVO_OK) {
					if (sh->slab) {
						uint32_t lo[3] = {0, 0, 0}, hi[3] = {res, res, res};
						for (uint32_t a = 0; a < axes; ++a) {
							const uint32_t bit = ((uint32_t)k >> a) & 1u;
							lo[a] = bit * half, hi[a] = bit * half + half;
						}
						r = svo_voxelizer_create_windowed(scene, level, mode, lo, hi, s, &p.vox);
					} else {
						svo_shard shd{};
						shd.shard_level = 1;
						shd.cube_index[0] = p.octant & 1u, shd.cube_index[1] = (p.octant >> 1) & 1u, shd.cube_index[2] = (p.octant >> 2) & 1u;
						r = svo_voxelizer_create(scene, level, mode, &shd, s, &p.vox);
					}
				}
				if (r == SVO_OK) r = svo_builder_create(p.vox, s, &p.builder);
				if (r != SVO_OK) {
					part_fail(p, r);
					return;
				}
			}
		});
	if (rc == SVO_OK) rc = sharded_run(sh);
	if (rc != SVO_OK) {
		char keep[512];
		snprintf(keep, sizeof(keep), "%s", svo_last_error());
		svo_sharded_destroy(sh);
		set_error("%s", keep);
		return rc;
	}
	*out = sh;
	return SVO_OK;
}

int svo_sharded_rebuild(svo_sharded *sh) {
	if (!sh) return fail(SVO_ERR_INVALID_ARGUMENT, "null argument");
	return sharded_run(sh);
}
const uint32_t *svo_sharded_octree(const svo_sharded *sh) { return sh ? sh->octree : nullptr; }
uint64_t svo_sharded_octree_range_bytes(const svo_sharded *sh) { return sh ? sh->total_words * 4 : 0; }
uint64_t svo_sharded_leaf_count(const svo_sharded *sh) { return sh ? sh->n_leaf : 0; }
uint64_t svo_sharded_fragment_count(const svo_sharded *sh) { return sh ? sh->n_frag : 0; }
float svo_sharded_last_ms(const svo_sharded *sh) { return sh ? sh->last_ms : 0.f; }

} // extern "C"
