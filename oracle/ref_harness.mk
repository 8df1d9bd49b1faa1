# oracle/ref_harness.mk -- builds the reference's own Voxelizer + OctreeBuilder behind a window-less main
# (integration/headless_harness.cpp) from the sources where they lie under $(REF); nothing is copied, outputs go to
# oracle/_ref/ only.  Not the reference's build system: plain g++ on the files the path needs.
#   make -f oracle/ref_harness.mk            (from the repository root; REF=/root/reference by default)
# The binary needs a Vulkan loader + ICD at run time (see the header of the harness); without one it prints "no Vulkan".
REF ?= /root/reference
OUT := oracle/_ref
CXX ?= g++
CXXFLAGS := -std=c++20 -O2 -w -DVK_NO_PROTOTYPES -pthread
INC := -I$(REF)/src -I$(REF)/dep -I$(REF)/dep/meshoptimizer/src -I$(REF)/dep/glm -I$(REF)/dep/spdlog/include \
       -I$(REF)/dep/MyVK/include -I$(REF)/dep/MyVK/dep/volk -I$(REF)/dep/MyVK/dep/vulkan -I$(REF)/dep/MyVK/dep/vma \
       -I$(REF)/shader/include
MYVK := ImageBase Image BufferBase Buffer CommandBuffer CommandPool Device Instance PhysicalDevice Queue QueueSelector Fence Semaphore \
        ImageView RenderPass PipelineBase PipelineLayout DescriptorSetLayout ShaderModule GraphicsPipeline ComputePipeline Framebuffer \
        DescriptorPool DescriptorSet Sampler ObjectTracker QueryPool FramebufferBase ImagelessFramebuffer
SRC := $(addprefix $(REF)/src/,Scene.cpp Voxelizer.cpp OctreeBuilder.cpp Counter.cpp) \
       $(addprefix $(REF)/dep/MyVK/src/,$(addsuffix .cpp,$(MYVK))) \
       $(REF)/dep/MyVK/dep/vma/vk_mem_alloc.cpp $(REF)/dep/stb_image.cpp $(REF)/dep/tiny_obj_loader.cpp \
       $(addprefix $(REF)/dep/meshoptimizer/src/,indexgenerator.cpp vcacheoptimizer.cpp overdrawoptimizer.cpp vfetchoptimizer.cpp)
OBJ := $(patsubst $(REF)/%.cpp,$(OUT)/obj/%.o,$(SRC))

$(OUT)/svo_ref_headless: integration/headless_harness.cpp $(OBJ) $(OUT)/obj/volk.o
	$(CXX) $(CXXFLAGS) $(INC) -o $@ $^ -ldl

$(OUT)/obj/%.o: $(REF)/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $(INC) -c -o $@ $<

$(OUT)/obj/volk.o: $(REF)/dep/MyVK/dep/volk/volk.c
	@mkdir -p $(dir $@)
	gcc -O2 -w -DVK_NO_PROTOTYPES -I$(REF)/dep/MyVK/dep/vulkan -c -o $@ $<

clean:
	rm -rf $(OUT)/obj $(OUT)/svo_ref_headless
