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
#include <cstdint>
#include <cstdio>
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
};

} // namespace svo_host
