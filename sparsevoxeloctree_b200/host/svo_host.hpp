// svo_host.hpp -- C++ host-side mirror of the reference's SVO construction classes over the C ABI (include/svo.h).
//
// The reference's boundary for this path is two C++ classes with one caller and one consumer
// (src/LoaderThread.cpp:51-89, src/Octree.cpp:22-35).  These classes keep the reference's names, argument
// meaning and ownership model (everything is a std::shared_ptr; factories return nullptr on failure and log
// nothing else -- src/Scene.cpp:398-401, dep/MyVK/src/Buffer.cpp:45-47), with the Vulkan handles replaced by
// a CUDA stream:
//
//   reference                                                  here
//   ---------------------------------------------------------  -----------------------------------------------
//   Scene::Create(queue, filename, notif*)                     Scene::Create(mesh, device, stream)   (mesh hand-off)
//   Voxelizer::Create(scene, command_pool, octree_level)       Voxelizer::Create(scene, octree_level, stream[, mode, shard])
//   Voxelizer::CmdVoxelize(command_buffer)                     Voxelizer::CmdVoxelize(stream)
//   OctreeBuilder::Create(voxelizer, command_pool)             OctreeBuilder::Create(voxelizer, stream)
//   OctreeBuilder::CmdBuild(command_buffer)                    OctreeBuilder::CmdBuild(stream)
//   OctreeBuilder::GetOctreeRange(command_pool) -> bytes       OctreeBuilder::GetOctreeRange() -> bytes
//   OctreeBuilder::GetOctree() -> myvk::Buffer                 OctreeBuilder::GetOctree() -> device pointer
//   Octree::Update(command_pool, builder)                      Octree::Update(builder)
//
// Header only; link against libsvo_b200.so.  No CPU fallback exists behind these calls.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/svo.h"

namespace svo_host {

typedef void *Stream; // cudaStream_t

struct Vertex { // src/Scene.cpp:16-19
	float m_position[3];
	float m_texcoord[2];
};

// What Scene keeps private in the reference (src/Scene.hpp:26,35-40): vertex + index buffers and the draw list.
struct TextureData { // one stbi_load(..., 4) image (src/Scene.cpp:245-262): RGBA8, sRGB colour
	uint32_t width{}, height{};
	std::vector<uint8_t> rgba8;
};
struct MeshData {
	std::vector<Vertex> vertices;
	std::vector<uint32_t> indices;
	std::vector<svo_draw> draws;
	std::vector<TextureData> textures; // indexed by svo_draw::texture_id
};

// Scene::load_meshes (src/Scene.cpp:36-143) without tinyobjloader: Wavefront OBJ + MTL -> MeshData.
// One mesh per material, vertices (position, (u, 1 - v)), positions normalised to [-1,1]^3 in fp32 (Scene.cpp:90-99),
// empty meshes dropped, the rest sorted by size (Scene.cpp:123-132), one draw per mesh with
// packUnorm4x8(vec4(Kd, 0)) and the texture id of map_Kd (ids in material order, Scene.cpp:101-121).
// Polygons are triangulated as a fan.  Image decoding is the caller's: texture_filenames receives the paths
// (Scene.cpp:139-142) and MeshData::textures is to be filled from them (the reference uses stb_image, Scene.cpp:247);
// draws whose texture is not supplied must have texture_id reset to 0xffffffff.
inline bool LoadObj(const std::string &filename, MeshData *out, std::vector<std::string> *texture_filenames) {
	struct Material {
		std::string name, map_kd;
		float kd[3] = {0.f, 0.f, 0.f};
		bool has_kd = false;
		std::vector<Vertex> vertices;
	};
	const size_t slash = filename.find_last_of("/\\");
	const std::string base_dir = slash == std::string::npos ? std::string() : filename.substr(0, slash + 1);
	std::ifstream obj(filename);
	if (!obj) {
		fprintf(stderr, "Failed to load %s\n", filename.c_str());
		return false;
	}
	std::vector<float> v, vt;
	std::vector<Material> mats;
	std::map<std::string, size_t> mat_index;
	auto load_mtl = [&](const std::string &path) {
		std::ifstream mtl(path);
		std::string line;
		Material *cur = nullptr;
		while (std::getline(mtl, line)) {
			line = line.substr(0, line.find('#'));
			std::istringstream ls(line);
			std::string tok;
			if (!(ls >> tok)) continue;
			if (tok == "newmtl") {
				std::string name;
				ls >> name;
				if (mat_index.count(name)) {
					cur = nullptr;
					continue;
				}
				mat_index[name] = mats.size();
				mats.emplace_back();
				cur = &mats.back();
				cur->name = name;
			} else if (cur && tok == "Kd") {
				ls >> cur->kd[0] >> cur->kd[1] >> cur->kd[2];
				cur->has_kd = true;
			} else if (cur && tok == "map_Kd") {
				std::string w;
				while (ls >> w) cur->map_kd = w; // options precede the file name
				if (!cur->has_kd) cur->kd[0] = cur->kd[1] = cur->kd[2] = 0.6f;
			}
		}
	};
	long cur_mat = -1;
	float pmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, pmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	std::string line;
	while (std::getline(obj, line)) {
		line = line.substr(0, line.find('#'));
		std::istringstream ls(line);
		std::string tok;
		if (!(ls >> tok)) continue;
		if (tok == "v") {
			float x = 0, y = 0, z = 0;
			ls >> x >> y >> z;
			v.insert(v.end(), {x, y, z});
		} else if (tok == "vt") {
			float a = 0, b = 0;
			ls >> a >> b;
			vt.insert(vt.end(), {a, b});
		} else if (tok == "mtllib") {
			std::string name;
			ls >> name;
			load_mtl(base_dir + name);
		} else if (tok == "usemtl") {
			std::string name;
			ls >> name;
			auto it = mat_index.find(name);
			cur_mat = it == mat_index.end() ? -1 : (long)it->second;
		} else if (tok == "f") {
			if (cur_mat < 0) {
				fprintf(stderr, "face without a material\n");
				return false;
			}
			std::vector<Vertex> corners;
			std::string c;
			while (ls >> c) {
				long vi = strtol(c.c_str(), nullptr, 10), ti = 0;
				const size_t s1 = c.find('/');
				if (s1 != std::string::npos && s1 + 1 < c.size() && c[s1 + 1] != '/') ti = strtol(c.c_str() + s1 + 1, nullptr, 10);
				vi = vi > 0 ? vi - 1 : (long)(v.size() / 3) + vi;
				Vertex vert{};
				for (int k = 0; k < 3; ++k) vert.m_position[k] = v[3 * vi + k];
				if (ti != 0) {
					ti = ti > 0 ? ti - 1 : (long)(vt.size() / 2) + ti;
					vert.m_texcoord[0] = vt[2 * ti];
					vert.m_texcoord[1] = 1.0f - vt[2 * ti + 1];
				}
				corners.push_back(vert);
			}
			for (size_t k = 1; k + 1 < corners.size(); ++k)
				for (const Vertex &vert : {corners[0], corners[k], corners[k + 1]}) {
					mats[cur_mat].vertices.push_back(vert);
					for (int a = 0; a < 3; ++a) {
						pmin[a] = std::min(pmin[a], vert.m_position[a]);
						pmax[a] = std::max(pmax[a], vert.m_position[a]);
					}
				}
		}
	}
	if (mats.empty()) {
		fprintf(stderr, "No material found\n");
		return false;
	}
	// normalize all the vertices to [-1, 1] (Scene.cpp:90-99)
	const float extent = std::max(pmax[0] - pmin[0], std::max(pmax[1] - pmin[1], pmax[2] - pmin[2])) * 0.5f;
	const float inv_extent = 1.0f / extent;
	float center[3];
	for (int a = 0; a < 3; ++a) center[a] = (pmax[a] + pmin[a]) * 0.5f;
	std::map<std::string, uint32_t> tex_ids;
	texture_filenames->clear();
	std::vector<uint32_t> tex_of(mats.size(), 0xffffffffu);
	for (size_t i = 0; i < mats.size(); ++i) {
		if (mats[i].vertices.empty() || mats[i].map_kd.empty()) continue;
		auto it = tex_ids.find(mats[i].map_kd);
		if (it == tex_ids.end()) {
			it = tex_ids.emplace(mats[i].map_kd, (uint32_t)tex_ids.size()).first;
			std::string rel = mats[i].map_kd;
			std::replace(rel.begin(), rel.end(), '\\', '/');
			texture_filenames->push_back(base_dir + rel);
		}
		tex_of[i] = it->second;
	}
	std::vector<size_t> order;
	for (size_t i = 0; i < mats.size(); ++i)
		if (!mats[i].vertices.empty()) order.push_back(i);
	if (order.empty()) {
		fprintf(stderr, "Empty mesh\n");
		return false;
	}
	std::stable_sort(order.begin(), order.end(), [&](size_t l, size_t r) { return mats[l].vertices.size() > mats[r].vertices.size(); });
	out->vertices.clear(), out->indices.clear(), out->draws.clear();
	for (size_t i : order) {
		auto pack = [](float c) { return (uint32_t)std::nearbyint(std::min(std::max(c, 0.f), 1.f) * 255.f); };
		svo_draw d{};
		d.first_index = (uint32_t)out->vertices.size();
		d.index_count = (uint32_t)mats[i].vertices.size();
		d.texture_id = tex_of[i];
		d.albedo_rgba8 = pack(mats[i].kd[0]) | (pack(mats[i].kd[1]) << 8) | (pack(mats[i].kd[2]) << 16);
		out->draws.push_back(d);
		for (Vertex vert : mats[i].vertices) {
			for (int a = 0; a < 3; ++a) vert.m_position[a] = (vert.m_position[a] - center[a]) * inv_extent;
			out->indices.push_back((uint32_t)out->vertices.size());
			out->vertices.push_back(vert);
		}
	}
	return true;
}

class Scene {
	svo_scene *m_handle{};

public:
	~Scene() { svo_scene_destroy(m_handle); }
	static std::shared_ptr<Scene> Create(const MeshData &mesh, int device = 0, Stream stream = nullptr) {
		svo_mesh m{};
		m.positions = mesh.vertices.data();
		m.position_stride_bytes = sizeof(Vertex);
		m.indices = mesh.indices.data();
		m.n_vertices = mesh.vertices.size();
		m.n_indices = mesh.indices.size();
		m.draws = mesh.draws.data();
		m.n_draws = (uint32_t)mesh.draws.size();
		std::vector<svo_texture> tex(mesh.textures.size());
		for (size_t i = 0; i < tex.size(); ++i) tex[i] = svo_texture{mesh.textures[i].rgba8.data(), mesh.textures[i].width, mesh.textures[i].height};
		if (!tex.empty() && !mesh.vertices.empty()) {
			m.n_textures = (uint32_t)tex.size();
			m.textures = tex.data();
			m.texcoords = mesh.vertices[0].m_texcoord;
			m.texcoord_stride_bytes = sizeof(Vertex);
		}
		auto ret = std::make_shared<Scene>();
		if (svo_scene_create(&m, device, stream, &ret->m_handle) != SVO_OK) {
			fprintf(stderr, "Scene::Create: %s\n", svo_last_error());
			return nullptr;
		}
		return ret;
	}
	svo_scene *GetHandle() const { return m_handle; }
	uint64_t GetTriangleCount() const { return svo_scene_triangle_count(m_handle); }
};

class Voxelizer { // src/Voxelizer.hpp
	std::shared_ptr<Scene> m_scene_ptr;
	svo_voxelizer *m_handle{};

public:
	~Voxelizer() { svo_voxelizer_destroy(m_handle); }
	static std::shared_ptr<Voxelizer> Create(const std::shared_ptr<Scene> &scene, uint32_t octree_level, Stream stream = nullptr,
	                                         int mode = SVO_CONSERVATIVE_EXACT, const svo_shard *shard = nullptr) {
		auto ret = std::make_shared<Voxelizer>();
		ret->m_scene_ptr = scene;
		if (svo_voxelizer_create(scene->GetHandle(), octree_level, mode, shard, stream, &ret->m_handle) != SVO_OK) {
			fprintf(stderr, "Voxelizer::Create: %s\n", svo_last_error());
			return nullptr;
		}
		// the reference logs the same line (src/Voxelizer.cpp:163)
		fprintf(stderr, "Voxel fragment list created with %llu voxels (%f MB)\n", (unsigned long long)ret->GetVoxelFragmentCount(),
		        ret->GetVoxelFragmentCount() * 8 / 1000000.0);
		return ret;
	}
	const std::shared_ptr<Scene> &GetScenePtr() const { return m_scene_ptr; }
	uint32_t GetLevel() const { return svo_voxelizer_level(m_handle); }
	void CmdVoxelize(Stream stream = nullptr) const { svo_voxelizer_voxelize(m_handle, stream); }
	uint32_t GetVoxelResolution() const { return svo_voxelizer_resolution(m_handle); }
	uint64_t GetVoxelFragmentCount() const { return svo_voxelizer_fragment_count(m_handle); }
	const uint64_t *GetVoxelFragmentList() const { return svo_voxelizer_fragments(m_handle); } // device pointer
	svo_voxelizer *GetHandle() const { return m_handle; }
};

class OctreeBuilder { // src/OctreeBuilder.hpp
	std::shared_ptr<Voxelizer> m_voxelizer_ptr;
	svo_builder *m_handle{};

public:
	~OctreeBuilder() { svo_builder_destroy(m_handle); }
	static std::shared_ptr<OctreeBuilder> Create(const std::shared_ptr<Voxelizer> &voxelizer, Stream stream = nullptr) {
		auto ret = std::make_shared<OctreeBuilder>();
		ret->m_voxelizer_ptr = voxelizer;
		if (svo_builder_create(voxelizer->GetHandle(), stream, &ret->m_handle) != SVO_OK) {
			fprintf(stderr, "OctreeBuilder::Create: %s\n", svo_last_error());
			return nullptr;
		}
		return ret;
	}
	const std::shared_ptr<Voxelizer> &GetVoxelizerPtr() const { return m_voxelizer_ptr; }
	uint32_t GetLevel() const { return svo_builder_level(m_handle); }
	int CmdBuild(Stream stream = nullptr) const { return svo_builder_build(m_handle, stream); }
	uint64_t GetOctreeRange() const { return svo_builder_octree_range_bytes(m_handle); } // bytes, like VkDeviceSize
	const uint32_t *GetOctree() const { return svo_builder_octree(m_handle); }           // device pointer
	// The same build in two phases, for hosts that place the node words themselves (several GPUs, externally allocated
	// memory; include/svo.h): Prepare, then EmitTo -- or, for a slab that took the brick path (CompactBytes() > 0), the
	// compact gather: PushTables + EmitCompactTo on the sending GPU, ExpandCompact on the GPU that owns the buffer.
	int Prepare(Stream stream = nullptr) const { return svo_builder_prepare(m_handle, stream); }
	int EmitTo(uint32_t *d_dst, uint32_t pointer_bias_words = 0, int skip_root = 0, Stream stream = nullptr) const {
		return svo_builder_emit_to(m_handle, d_dst, pointer_bias_words, skip_root, stream);
	}
	uint64_t CompactBytes() const { return svo_builder_compact_bytes(m_handle); }
	int PushTables(uint32_t pointer_bias_words, int skip_root, void *d_tables, uint64_t plan[4], Stream stream = nullptr) const {
		return svo_builder_push_tables(m_handle, pointer_bias_words, skip_root, d_tables, plan, stream);
	}
	int EmitCompactTo(uint32_t *d_dst, uint32_t pointer_bias_words, int skip_root, void *d_tables, uint64_t plan[4], Stream stream = nullptr) const {
		return svo_builder_emit_compact_to(m_handle, d_dst, pointer_bias_words, skip_root, d_tables, plan, stream);
	}
	static int ExpandCompact(int device, const void *d_tables, const uint64_t plan[4], uint32_t *d_dst, Stream stream = nullptr) {
		return svo_expand_compact(device, d_tables, plan, d_dst, stream);
	}
	// Queue-family ownership transfer (src/OctreeBuilder.cpp:215-220): nothing to record for a CUDA-produced buffer;
	// the Vulkan side acquires external memory from VK_QUEUE_FAMILY_EXTERNAL.
	void CmdTransferOctreeOwnership(Stream, uint32_t, uint32_t) const {}
	svo_builder *GetHandle() const { return m_handle; }
};

class Octree { // src/Octree.hpp
	std::shared_ptr<OctreeBuilder> m_builder; // owns the device allocation (the reference copies the myvk::Buffer pointer)
	const uint32_t *m_buffer{};
	uint64_t m_range{};
	uint32_t m_level{};

public:
	static std::shared_ptr<Octree> Create() { return std::make_shared<Octree>(); }
	void Update(const std::shared_ptr<OctreeBuilder> &builder) { // src/Octree.cpp:22-35
		m_builder = builder;
		m_range = builder->GetOctreeRange();
		m_buffer = builder->GetOctree();
		m_level = builder->GetLevel();
	}
	bool Empty() const { return m_buffer == nullptr; }
	const uint32_t *GetBuffer() const { return m_buffer; }
	uint32_t GetLevel() const { return m_level; }
	uint64_t GetRange() const { return m_range; }
	// What OctreeTracer does with the buffer, for checks without Vulkan: Octree_RayMarchLeaf (shader/octree.glsl:179-340)
	// over n rays; origins / dirs (3 floats per ray) and hits are DEVICE pointers.
	int RayMarchLeaf(uint64_t n_rays, const float *d_origins, const float *d_dirs, svo_ray_hit *d_hits, int device = 0,
	                 Stream stream = nullptr) const {
		return svo_octree_raymarch_leaf(device, m_buffer, n_rays, d_origins, d_dirs, d_hits, stream);
	}
};

} // namespace svo_host
