// common.cuh -- platform layer shared by all translation units of libsvo_b200.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a (CUDA runtime, real kernels).
// SVO_EMU build: g++ with tests/cpu_emu/cuda_emu.h -- a kernel-LOGIC emulator (one OS thread per CUDA
// thread, blocks run one after another) used only by the CPU-only unit tests to exercise the kernels'
// indexing / scan / ranking logic where no GPU exists.  It is never built into, loaded by, or reachable
// from the product package: sparsevoxeloctree_b200 loads libsvo_b200.so only and fails loudly without it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef SVO_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#define SVO_HD __host__ __device__
#define SVO_DEV __device__ __forceinline__

#ifdef SVO_EMU
// cooperative kernel (uses __syncthreads / warp collectives): one OS thread per CUDA thread
#define SVO_LAUNCH(grid, block, smem, stream, kernel, ...) \
	(++svo::g_launches, svo_emu::launch((grid), (block), (smem), true, [=]() { kernel(__VA_ARGS__); }))
// independent-thread kernel (no barrier, no warp collective): threads are run one after another
#define SVO_LAUNCH_INDEP(grid, block, stream, kernel, ...) \
	(++svo::g_launches, svo_emu::launch((grid), (block), 0, false, [=]() { kernel(__VA_ARGS__); }))
#define SVO_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(svo_emu::tctx.dyn_smem)
#define SVO_EMU_WARP_ORDER() __syncwarp()
#else
#define SVO_EMU_WARP_ORDER() ((void)0)
#define SVO_LAUNCH(grid, block, smem, stream, kernel, ...) \
	(++svo::g_launches, kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
#define SVO_LAUNCH_INDEP(grid, block, stream, kernel, ...) \
	(++svo::g_launches, kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__))
#define SVO_DYN_SMEM(type, name)                               \
	extern __shared__ __align__(16) unsigned char svo_dyn_smem_raw[]; \
	type *name = reinterpret_cast<type *>(svo_dyn_smem_raw)
#endif

namespace svo {

// kernels launched by this library since load (host-side count; the bench reports it as gpu_launches)
inline uint64_t g_launches = 0;

constexpr uint32_t FULL_MASK = 0xffffffffu;
constexpr int WARP = 32;

// ---- error plumbing (no exceptions across the C ABI) ----------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

#define SVO_CUDA_TRY(expr)                                                                         \
	do {                                                                                           \
		cudaError_t svo_err__ = (expr);                                                            \
		if (svo_err__ != cudaSuccess) {                                                            \
			svo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(svo_err__), __FILE__, __LINE__); \
			return -2; /* SVO_ERR_CUDA */                                                          \
		}                                                                                          \
	} while (0)

#define SVO_TRY(expr)              \
	do {                           \
		int svo_rc__ = (expr);     \
		if (svo_rc__ != 0) return svo_rc__; \
	} while (0)

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// ---- stream-ordered device memory -----------------------------------------------------------------
// cudaMallocAsync from the device's default pool with the release threshold raised, so that steady-state
// create/build/destroy cycles recycle memory without touching the driver allocator.
int dev_alloc(void **p, uint64_t bytes, cudaStream_t s);
void dev_free(void *p, cudaStream_t s);
int configure_device_pool(int device);
int sm_count(int device);

template <class T> struct DevBuf {
	T *p = nullptr;
	uint64_t n = 0;
	int alloc(uint64_t count, cudaStream_t s) {
		release(s);
		n = count;
		return dev_alloc(reinterpret_cast<void **>(&p), (count ? count : 1) * sizeof(T), s);
	}
	// grow-only reuse
	int reserve(uint64_t count, cudaStream_t s) {
		if (p && n >= count) return 0;
		return alloc(count, s);
	}
	void release(cudaStream_t s) {
		if (p) dev_free(p, s);
		p = nullptr;
		n = 0;
	}
};

} // namespace svo
