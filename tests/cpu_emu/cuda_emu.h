// tests/cpu_emu/cuda_emu.h -- a minimal CUDA kernel-LOGIC emulator for the CPU-only unit tests.
//
// TEST INFRASTRUCTURE ONLY.  Compiled (with g++ -DSVO_EMU) together with the product's .cu sources into
// tests/_build/libsvo_emu.so, which only tests/test_emu_*.py load, by explicit path.  It lets the tests that
// run without a GPU execute the kernels' indexing, scan, ranking and look-back logic on tiny inputs:
//   * every CUDA thread of a block is an OS thread; blocks run one after another, in blockIdx order;
//   * __syncthreads / warp collectives are std::barrier rendezvous (threads that return drop out);
//   * warp collectives must be called by the whole (non-exited) warp with a full mask;
//   * __shared__ variables are function statics (valid because only one block is live at a time).
// It does not model memory ordering, occupancy, or concurrency between blocks -- those are covered by the
// GPU tests.  The product package never loads this library and has no CPU fallback.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct float3 { float x, y, z; };
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return {x, y}; }

// ---- runtime API subset ----------------------------------------------------------------------------
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef void *cudaStream_t;
struct svo_emu_event { std::chrono::steady_clock::time_point t; };
typedef svo_emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) {
	*p = aligned_alloc(256, (n + 255) / 256 * 256);
	memset(*p, 0xCD, n); // poison: reads of uninitialised device memory show up
	return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
constexpr unsigned cudaStreamNonBlocking = 1;
constexpr cudaError_t cudaErrorPeerAccessAlreadyEnabled = 704;
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; } // (launches run to completion in call order)
inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new svo_emu_event(); return cudaSuccess; }
constexpr unsigned cudaEventDisableTiming = 2;
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
	*ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
	return cudaSuccess;
}

namespace svo_emu {

struct WarpState {
	std::unique_ptr<std::barrier<>> bar;
	uint64_t slot[32];
	unsigned live_mask = 0;
};
struct BlockState {
	std::unique_ptr<std::barrier<>> bar;
	std::vector<WarpState> warps;
};
struct ThreadCtx {
	uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
	dim3 blockDim, gridDim;
	void *dyn_smem = nullptr;
	BlockState *blk = nullptr;
	WarpState *warp = nullptr;
	unsigned lane = 0;
	bool cooperative = false;
};
inline thread_local ThreadCtx tctx;

inline void fail(const char *msg) {
	fprintf(stderr, "cuda_emu: %s\n", msg);
	abort();
}

template <class F> void launch(dim3 grid, dim3 block, size_t smem, bool cooperative, F fn) {
	if (block.y != 1 || block.z != 1 || grid.z != 1) fail("only 1-D blocks and 1-D / 2-D grids are emulated");
	std::vector<unsigned char> dyn(smem + 16);
	for (unsigned by = 0; by < grid.y; ++by)
	for (unsigned b = 0; b < grid.x; ++b) {
		memset(dyn.data(), 0xCD, dyn.size());
		if (!cooperative) {
			for (unsigned t = 0; t < block.x; ++t) {
				tctx = ThreadCtx();
				tctx.threadIdx = {t, 0, 0}, tctx.blockIdx = {b, by, 0};
				tctx.blockDim = block, tctx.gridDim = grid;
				tctx.dyn_smem = dyn.data();
				tctx.lane = t % 32;
				fn();
			}
			continue;
		}
		BlockState bs;
		bs.bar = std::make_unique<std::barrier<>>(block.x);
		unsigned nw = (block.x + 31) / 32;
		bs.warps.resize(nw);
		for (unsigned w = 0; w < nw; ++w) {
			unsigned cnt = std::min(32u, block.x - w * 32);
			bs.warps[w].bar = std::make_unique<std::barrier<>>(cnt);
			bs.warps[w].live_mask = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
		}
		std::vector<std::thread> th;
		th.reserve(block.x);
		for (unsigned t = 0; t < block.x; ++t) {
			th.emplace_back([&, t]() {
				tctx = ThreadCtx();
				tctx.threadIdx = {t, 0, 0}, tctx.blockIdx = {b, by, 0};
				tctx.blockDim = block, tctx.gridDim = grid;
				tctx.dyn_smem = dyn.data();
				tctx.blk = &bs, tctx.warp = &bs.warps[t / 32], tctx.lane = t % 32;
				tctx.cooperative = true;
				fn();
				// a thread that returns no longer takes part in barriers (CUDA semantics)
				tctx.warp->bar->arrive_and_drop();
				tctx.blk->bar->arrive_and_drop();
			});
		}
		for (auto &x : th) x.join();
	}
}

inline void warp_sync() {
	if (!tctx.cooperative) fail("warp collective in a kernel launched with SVO_LAUNCH_INDEP");
	tctx.warp->bar->arrive_and_wait();
}
// all-lanes exchange: returns a pointer to the 32 published values (valid until the next collective)
inline const uint64_t *warp_publish(unsigned mask, uint64_t v) {
	if (mask != 0xffffffffu) fail("emulated warp collectives need a full mask");
	tctx.warp->slot[tctx.lane] = v;
	warp_sync();
	return tctx.warp->slot;
}
} // namespace svo_emu

#define threadIdx (svo_emu::tctx.threadIdx)
#define blockIdx (svo_emu::tctx.blockIdx)
#define blockDim (svo_emu::tctx.blockDim)
#define gridDim (svo_emu::tctx.gridDim)

inline void __syncthreads() {
	if (!svo_emu::tctx.cooperative) svo_emu::fail("__syncthreads in a kernel launched with SVO_LAUNCH_INDEP");
	svo_emu::tctx.blk->bar->arrive_and_wait();
}
inline void __syncwarp(unsigned = 0xffffffffu) { svo_emu::warp_sync(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned __activemask() { return svo_emu::tctx.cooperative ? svo_emu::tctx.warp->live_mask : 1u << svo_emu::tctx.lane; }

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
	static_assert(sizeof(T) <= 8, "");
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *s = svo_emu::warp_publish(mask, raw);
	unsigned lane = svo_emu::tctx.lane;
	unsigned base = lane / width * width;
	uint64_t r = s[base + ((unsigned)src % width)];
	svo_emu::warp_sync();
	T out;
	memcpy(&out, &r, sizeof(T));
	return out;
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *s = svo_emu::warp_publish(mask, raw);
	unsigned lane = svo_emu::tctx.lane;
	unsigned base = lane / width * width;
	uint64_t r = (lane - base >= delta) ? s[lane - delta] : raw;
	svo_emu::warp_sync();
	T out;
	memcpy(&out, &r, sizeof(T));
	return out;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *s = svo_emu::warp_publish(mask, raw);
	unsigned lane = svo_emu::tctx.lane;
	unsigned base = lane / width * width;
	uint64_t r = (lane - base + delta < (unsigned)width) ? s[lane + delta] : raw;
	svo_emu::warp_sync();
	T out;
	memcpy(&out, &r, sizeof(T));
	return out;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *s = svo_emu::warp_publish(mask, raw);
	(void)width;
	uint64_t r = s[svo_emu::tctx.lane ^ (unsigned)lanemask];
	svo_emu::warp_sync();
	T out;
	memcpy(&out, &r, sizeof(T));
	return out;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
	const uint64_t *s = svo_emu::warp_publish(mask, pred ? 1 : 0);
	unsigned r = 0, live = svo_emu::tctx.warp->live_mask;
	for (int i = 0; i < 32; ++i)
		if (((live >> i) & 1u) && s[i]) r |= 1u << i;
	svo_emu::warp_sync();
	return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == svo_emu::tctx.warp->live_mask; }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *s = svo_emu::warp_publish(mask, raw);
	unsigned r = 0, live = svo_emu::tctx.warp->live_mask;
	for (int i = 0; i < 32; ++i)
		if (((live >> i) & 1u) && s[i] == raw) r |= 1u << i;
	svo_emu::warp_sync();
	return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
	const uint64_t *s = svo_emu::warp_publish(mask, v);
	unsigned r = 0, live = svo_emu::tctx.warp->live_mask;
	for (int i = 0; i < 32; ++i)
		if ((live >> i) & 1u) r += (unsigned)s[i];
	svo_emu::warp_sync();
	return r;
}

inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
	return (unsigned)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}

inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
	const uint64_t *s = svo_emu::warp_publish(mask, v);
	unsigned r = 0, live = svo_emu::tctx.warp->live_mask;
	for (int i = 0; i < 32; ++i)
		if ((live >> i) & 1u) r |= (unsigned)s[i];
	svo_emu::warp_sync();
	return r;
}

// ---- atomics (device-wide; blocks are sequential but threads of a block are concurrent) --------------
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMax(unsigned *p, unsigned v) {
	unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
	return old;
}
inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) {
	__atomic_compare_exchange_n(p, &cmp, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
	return cmp;
}

// ---- intrinsics --------------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v) {
	unsigned r = 0;
	for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
	return r;
}
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
