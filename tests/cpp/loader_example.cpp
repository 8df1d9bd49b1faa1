// The reference loader sequence (src/LoaderThread.cpp:51-116) written against the C++ mirror classes:
// Scene -> Voxelizer -> OctreeBuilder -> CmdVoxelize + CmdBuild -> Octree::Update, on a small procedural mesh.
// Prints "fragments range level root[0..7]".  With --mesh <in.bin> <level> <mode> <out.bin> the mesh comes from a file
// (u64 nv, ni, nd; nv*3 floats; ni u32 indices; nd svo_draw) and the whole node buffer goes to out.bin (u64 fragments,
// u64 range, words): tests/test_cpp_host.py feeds both hosts the same bytes and compares the buffers word for word.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../sparsevoxeloctree_b200/host/svo_host.hpp"

using namespace svo_host;

int main(int argc, char **argv) {
	if (argc > 2 && std::string(argv[1]) == "--obj") { // dump what LoadObj makes of a file (no GPU needed)
		MeshData mesh;
		std::vector<std::string> tex;
		if (!LoadObj(argv[2], &mesh, &tex)) return 1;
		printf("%zu %zu %zu\n", mesh.vertices.size(), mesh.draws.size(), tex.size());
		for (const svo_draw &d : mesh.draws) printf("d %u %u %u %u\n", d.first_index, d.index_count, d.texture_id, d.albedo_rgba8);
		for (const Vertex &v : mesh.vertices) {
			uint32_t b[5];
			memcpy(b, &v, sizeof(b));
			printf("v %08x %08x %08x %08x %08x\n", b[0], b[1], b[2], b[3], b[4]);
		}
		for (const std::string &t : tex) printf("t %s\n", t.c_str());
		return 0;
	}
	uint32_t level = argc > 1 ? (uint32_t)atoi(argv[1]) : 7;
	const int n = argc > 2 ? atoi(argv[2]) : 33;
	MeshData mesh;
	const bool from_file = argc > 5 && std::string(argv[1]) == "--mesh";
	int mode = SVO_CONSERVATIVE_EXACT;
	if (from_file) {
		FILE *f = fopen(argv[2], "rb");
		if (!f) return 1;
		uint64_t hdr[3];
		if (fread(hdr, 8, 3, f) != 3) return 1;
		std::vector<float> pos(hdr[0] * 3);
		mesh.indices.resize(hdr[1]);
		mesh.draws.resize(hdr[2]);
		if (fread(pos.data(), 4, pos.size(), f) != pos.size() || fread(mesh.indices.data(), 4, hdr[1], f) != hdr[1] ||
		    fread(mesh.draws.data(), sizeof(svo_draw), hdr[2], f) != hdr[2])
			return 1;
		fclose(f);
		for (uint64_t i = 0; i < hdr[0]; ++i) mesh.vertices.push_back({{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, {0.f, 0.f}});
		level = (uint32_t)atoi(argv[3]);
		mode = atoi(argv[4]);
	}
	for (int i = 0; i < n && !from_file; ++i)
		for (int j = 0; j < n; ++j) {
			float x = -1.f + 2.f * i / (n - 1), z = -1.f + 2.f * j / (n - 1);
			float y = 0.4f * std::sin(3.f * x) * std::cos(2.f * z);
			mesh.vertices.push_back({{x, y, z}, {0.f, 0.f}});
		}
	for (int i = 0; i + 1 < n && !from_file; ++i)
		for (int j = 0; j + 1 < n; ++j) {
			uint32_t a = i * n + j, b = (i + 1) * n + j, c = (i + 1) * n + j + 1, d = i * n + j + 1;
			for (uint32_t v : {a, b, c, a, c, d}) mesh.indices.push_back(v);
		}
	if (!from_file) mesh.draws.push_back({0, (uint32_t)mesh.indices.size(), 0xffffffffu, 0x00C83C32u});

	if (svo_device_count() < 1) {
		fprintf(stderr, "no CUDA device\n");
		return 2;
	}
	if (from_file && getenv("SVO_DEVICES")) { // the same mesh through svo_build_sharded on the listed devices ("0,1,2,3")
		std::vector<int> devices;
		for (const char *p = getenv("SVO_DEVICES"); *p;) {
			devices.push_back(atoi(p));
			while (*p && *p != ',') ++p;
			if (*p == ',') ++p;
		}
		svo_mesh m{};
		m.positions = mesh.vertices.data(), m.position_stride_bytes = sizeof(Vertex);
		m.indices = mesh.indices.data(), m.n_vertices = mesh.vertices.size(), m.n_indices = mesh.indices.size();
		m.draws = mesh.draws.data(), m.n_draws = (uint32_t)mesh.draws.size();
		svo_sharded *sh = nullptr;
		if (svo_build_sharded(&m, level, mode, devices.data(), (uint32_t)devices.size(), &sh) != SVO_OK) {
			fprintf(stderr, "svo_build_sharded: %s\n", svo_last_error());
			return 1;
		}
		const uint64_t range = svo_sharded_octree_range_bytes(sh);
		std::vector<uint32_t> words(range / 4);
		svo_memcpy_d2h(devices[0], words.data(), svo_sharded_octree(sh), range, nullptr);
		printf("%llu %llu %u sharded x%zu %.3f ms\n", (unsigned long long)svo_sharded_fragment_count(sh), (unsigned long long)range, level,
		       devices.size(), svo_sharded_last_ms(sh));
		FILE *f = fopen(argv[5], "wb");
		if (!f) return 1;
		const uint64_t hdr[2] = {svo_sharded_fragment_count(sh), range};
		fwrite(hdr, 8, 2, f);
		fwrite(words.data(), 4, words.size(), f);
		fclose(f);
		svo_sharded_destroy(sh);
		return 0;
	}
	auto scene = Scene::Create(mesh);
	if (!scene) return 1;
	auto voxelizer = Voxelizer::Create(scene, level, nullptr, mode);
	if (!voxelizer) return 1;
	auto builder = OctreeBuilder::Create(voxelizer);
	if (!builder) return 1;
	voxelizer->CmdVoxelize();
	if (builder->CmdBuild() != SVO_OK) {
		fprintf(stderr, "CmdBuild: %s\n", svo_last_error());
		return 1;
	}
	auto octree = Octree::Create();
	octree->Update(builder);
	uint32_t root[8];
	svo_memcpy_d2h(0, root, octree->GetBuffer(), sizeof(root), nullptr);
	printf("%llu %llu %u", (unsigned long long)voxelizer->GetVoxelFragmentCount(), (unsigned long long)octree->GetRange(), octree->GetLevel());
	for (uint32_t w : root) printf(" %08x", w);
	printf("\n");
	fprintf(stderr, "Octree range: %llu (%f MB)\n", (unsigned long long)octree->GetRange(), octree->GetRange() / 1000000.0f);
	if (from_file) {
		std::vector<uint32_t> words(octree->GetRange() / 4);
		svo_memcpy_d2h(0, words.data(), octree->GetBuffer(), octree->GetRange(), nullptr);
		FILE *f = fopen(argv[5], "wb");
		if (!f) return 1;
		const uint64_t hdr[2] = {voxelizer->GetVoxelFragmentCount(), octree->GetRange()};
		fwrite(hdr, 8, 2, f);
		fwrite(words.data(), 4, words.size(), f);
		fclose(f);
	}
	return 0;
}
