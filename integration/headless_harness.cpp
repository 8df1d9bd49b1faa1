// integration/headless_harness.cpp -- the reference's OWN Voxelizer + OctreeBuilder run without a window (SURVEY.md
// section 8, row f3): the loader sequence of src/LoaderThread.cpp:43-116 and the device set-up of
// src/Application.cpp:161-293, minus GLFW, surface, swapchain and UI.  It is compiled by oracle/ref_harness.mk against
// the reference sources where they lie (Scene / Voxelizer / OctreeBuilder / Counter + MyVK + volk + VMA + tinyobj +
// meshoptimizer + stb_image + spdlog, nothing copied) into oracle/_ref/svo_ref_headless.  It needs a Vulkan loader and
// an ICD at RUN time (lavapipe on the host cores, or the NVIDIA ICD): neither exists in this image, so today it
// builds and exits with "no Vulkan" -- the day an ICD is present it produces the driver-level baseline and diff.
//
//   svo_ref_headless <scene.obj> <level> <out_prefix>
// writes <out_prefix>.frags (u32 x 2 per fragment, the reference's packing, voxelizer.frag:40-42) and
// <out_prefix>.octree (the node words, OctreeBuilder::GetOctreeRange bytes) and prints the four timestamps the
// reference prints (LoaderThread.cpp:92-97) plus the wall clock of Voxelizer::Create (the count pass).
// tests/parity.py-style comparison: oracle.unpack the fragments, oracle.canonicalise the octree.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include <spdlog/spdlog.h>

#include "OctreeBuilder.hpp"
#include "Scene.hpp"
#include "Voxelizer.hpp"
#include "myvk/Buffer.hpp"
#include "myvk/CommandBuffer.hpp"
#include "myvk/Fence.hpp"
#include "myvk/Instance.hpp"
#include "myvk/QueryPool.hpp"
#include "myvk/QueueSelector.hpp"

static bool dump(const myvk::Ptr<myvk::CommandPool> &pool, const myvk::Ptr<myvk::BufferBase> &src, VkDeviceSize bytes,
                 const std::string &path) {
	const auto &device = pool->GetDevicePtr();
	auto staging = myvk::Buffer::Create(device, bytes, VMA_ALLOCATION_CREATE_MAPPED_BIT | VMA_ALLOCATION_CREATE_HOST_ACCESS_RANDOM_BIT,
	                                    VK_BUFFER_USAGE_TRANSFER_DST_BIT);
	if (!staging) return false;
	auto fence = myvk::Fence::Create(device);
	auto cb = myvk::CommandBuffer::Create(pool);
	cb->Begin(VK_COMMAND_BUFFER_USAGE_ONE_TIME_SUBMIT_BIT);
	cb->CmdCopy(src, staging, {{0, 0, bytes}});
	cb->End();
	cb->Submit(fence);
	fence->Wait();
	std::ofstream f(path, std::ios::binary);
	f.write(static_cast<const char *>(staging->GetMappedData()), (std::streamsize)bytes);
	return bool(f);
}

int main(int argc, char **argv) {
	if (argc < 4) {
		fprintf(stderr, "usage: %s <scene.obj> <level> <out_prefix>\n", argv[0]);
		return 2;
	}
	const uint32_t level = (uint32_t)atoi(argv[2]);
	const std::string prefix = argv[3];
	if (volkInitialize() != VK_SUCCESS) {
		fprintf(stderr, "no Vulkan: volkInitialize failed (no loader library)\n");
		return 3;
	}
	auto instance = myvk::Instance::Create({}); // no window-system extensions
	if (!instance) {
		fprintf(stderr, "no Vulkan: vkCreateInstance failed\n");
		return 3;
	}
	auto physical_devices = myvk::PhysicalDevice::Fetch(instance);
	if (physical_devices.empty()) {
		fprintf(stderr, "no Vulkan: no physical device (no ICD)\n");
		return 3;
	}
	const auto &physical_device = physical_devices[0]; // like Application.cpp:191
	spdlog::info("Physical Device: {}", physical_device->GetProperties().vk10.deviceName);
	std::vector<const char *> extensions;
	// the reference picks Mode A (hardware conservative raster) when the extension exists, else Mode B (Voxelizer.cpp:92-99)
	if (physical_device->GetExtensionSupport(VK_EXT_CONSERVATIVE_RASTERIZATION_EXTENSION_NAME)) {
		extensions.push_back(VK_EXT_CONSERVATIVE_RASTERIZATION_EXTENSION_NAME);
		spdlog::info("EXT_conservative_rasterization supported: raster mode A");
	} else
		spdlog::warn("EXT_conservative_rasterization not supported: raster mode B (voxelizer_conservative.geom)");
	myvk::Ptr<myvk::Queue> queue;
	auto features = physical_device->GetDefaultFeatures();
	features.vk12.descriptorBindingPartiallyBound = VK_TRUE; // Application.cpp:270-272
	auto device = myvk::Device::Create(physical_device, myvk::GenericQueueSelector{&queue}, features, extensions);
	if (!device || !queue) {
		fprintf(stderr, "no Vulkan: device creation failed\n");
		return 3;
	}
	auto pool = myvk::CommandPool::Create(queue);

	std::atomic<const char *> notification{""};
	auto scene = Scene::Create(queue, argv[1], &notification);
	if (!scene) {
		fprintf(stderr, "Scene::Create failed for %s\n", argv[1]);
		return 1;
	}
	const auto t0 = std::chrono::steady_clock::now();
	auto voxelizer = Voxelizer::Create(scene, pool, level); // includes the count pass (Voxelizer.cpp:134-165)
	const auto t1 = std::chrono::steady_clock::now();
	auto builder = OctreeBuilder::Create(voxelizer, pool);

	auto fence = myvk::Fence::Create(device);
	auto query_pool = myvk::QueryPool::Create(device, VK_QUERY_TYPE_TIMESTAMP, 4);
	auto cb = myvk::CommandBuffer::Create(pool);
	cb->Begin(VK_COMMAND_BUFFER_USAGE_ONE_TIME_SUBMIT_BIT);
	cb->CmdResetQueryPool(query_pool);
	cb->CmdWriteTimestamp(VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, query_pool, 0);
	voxelizer->CmdVoxelize(cb);
	cb->CmdWriteTimestamp(VK_PIPELINE_STAGE_BOTTOM_OF_PIPE_BIT, query_pool, 1);
	const std::vector<VkBufferMemoryBarrier> frag_barrier = {
	    voxelizer->GetVoxelFragmentList()->GetMemoryBarrier(VK_ACCESS_SHADER_WRITE_BIT, VK_ACCESS_SHADER_READ_BIT)};
	cb->CmdPipelineBarrier(VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT, VK_PIPELINE_STAGE_COMPUTE_SHADER_BIT, {}, frag_barrier, {});
	cb->CmdWriteTimestamp(VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, query_pool, 2);
	builder->CmdBuild(cb);
	cb->CmdWriteTimestamp(VK_PIPELINE_STAGE_BOTTOM_OF_PIPE_BIT, query_pool, 3);
	cb->End();
	cb->Submit(fence);
	fence->Wait();
	uint64_t ts[4];
	query_pool->GetResults64(ts, VK_QUERY_RESULT_WAIT_BIT);
	const double period_ms = physical_device->GetProperties().vk10.limits.timestampPeriod * 1e-6; // (the reference assumes 1 ns per tick)
	const VkDeviceSize range = builder->GetOctreeRange(pool);
	printf("{\"level\": %u, \"fragments\": %u, \"octree_range_bytes\": %llu, \"count_pass_ms\": %.3f, \"voxelize_ms\": %.3f, "
	       "\"build_ms\": %.3f, \"svo_build_ms\": %.3f}\n",
	       level, voxelizer->GetVoxelFragmentCount(), (unsigned long long)range,
	       std::chrono::duration<double, std::milli>(t1 - t0).count(), double(ts[1] - ts[0]) * period_ms, double(ts[3] - ts[2]) * period_ms,
	       double(ts[3] - ts[0]) * period_ms);
	if (!dump(pool, voxelizer->GetVoxelFragmentList(), (VkDeviceSize)voxelizer->GetVoxelFragmentCount() * 8, prefix + ".frags") ||
	    !dump(pool, builder->GetOctree(), range, prefix + ".octree"))
		return 1;
	return 0;
}
