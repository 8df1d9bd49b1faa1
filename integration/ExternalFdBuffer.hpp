// integration/ExternalFdBuffer.hpp -- the Vulkan half of the octree hand-off.
//
// A myvk::BufferBase (dep/MyVK/include/myvk/BufferBase.hpp:12-22) whose memory is the file descriptor returned by
// svo_builder_export_fd (include/svo.h): the node words written by the CUDA builder become the storage buffer that
// Octree::Update binds (src/Octree.cpp:22-35), so OctreeTracer and PathTracer read them unchanged.  Needs the device
// extensions VK_KHR_external_memory and VK_KHR_external_memory_fd (add them next to the swapchain extension in
// Application::create_device, src/Application.cpp:270-272).  Drop this file into src/ of the reference; it only uses
// what the reference already vendors (volk, myvk).  The three one-line changes that go with it are in INTEGRATION.md.
#ifndef SVO_EXTERNAL_FD_BUFFER_HPP
#define SVO_EXTERNAL_FD_BUFFER_HPP

#include <myvk/BufferBase.hpp>

class ExternalFdBuffer final : public myvk::BufferBase {
private:
	myvk::Ptr<myvk::Device> m_device_ptr;
	VkDeviceMemory m_memory{VK_NULL_HANDLE};

public:
	// fd, alloc_size: the results of svo_builder_export_fd.  On success Vulkan owns the descriptor.
	static myvk::Ptr<ExternalFdBuffer> Create(const myvk::Ptr<myvk::Device> &device, int fd, VkDeviceSize alloc_size,
	                                          VkBufferUsageFlags usage = VK_BUFFER_USAGE_STORAGE_BUFFER_BIT) {
		auto ret = std::make_shared<ExternalFdBuffer>();
		ret->m_device_ptr = device;
		ret->m_size = alloc_size;

		VkExternalMemoryBufferCreateInfo external_info = {VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO};
		external_info.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
		VkBufferCreateInfo buffer_info = {VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO};
		buffer_info.pNext = &external_info;
		buffer_info.size = alloc_size;
		buffer_info.usage = usage;
		buffer_info.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
		if (vkCreateBuffer(device->GetHandle(), &buffer_info, nullptr, &ret->m_buffer) != VK_SUCCESS)
			return nullptr;

		VkMemoryRequirements requirements;
		vkGetBufferMemoryRequirements(device->GetHandle(), ret->m_buffer, &requirements);
		const VkPhysicalDeviceMemoryProperties &props = device->GetPhysicalDevicePtr()->GetMemoryProperties();
		uint32_t type = UINT32_MAX;
		for (uint32_t i = 0; i < props.memoryTypeCount; ++i)
			if ((requirements.memoryTypeBits >> i & 1u) &&
			    (props.memoryTypes[i].propertyFlags & VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT)) {
				type = i;
				break;
			}
		if (type == UINT32_MAX || requirements.size > alloc_size)
			return nullptr; // (the destructor releases the buffer)

		VkImportMemoryFdInfoKHR import_info = {VK_STRUCTURE_TYPE_IMPORT_MEMORY_FD_INFO_KHR};
		import_info.handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
		import_info.fd = fd;
		VkMemoryAllocateInfo allocate_info = {VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO};
		allocate_info.pNext = &import_info;
		allocate_info.allocationSize = alloc_size;
		allocate_info.memoryTypeIndex = type;
		if (vkAllocateMemory(device->GetHandle(), &allocate_info, nullptr, &ret->m_memory) != VK_SUCCESS)
			return nullptr;
		if (vkBindBufferMemory(device->GetHandle(), ret->m_buffer, ret->m_memory, 0) != VK_SUCCESS)
			return nullptr;
		return ret;
	}

	const myvk::Ptr<myvk::Device> &GetDevicePtr() const override { return m_device_ptr; }

	~ExternalFdBuffer() override {
		if (m_buffer) vkDestroyBuffer(m_device_ptr->GetHandle(), m_buffer, nullptr);
		if (m_memory) vkFreeMemory(m_device_ptr->GetHandle(), m_memory, nullptr);
	}
};

#endif
