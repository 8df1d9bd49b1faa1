// scan.cuh -- warp / block scans and the single-pass "decoupled look-back" chained scan used by the
// count->scan->emit voxelizer, the de-duplication and the level compaction kernels.
#pragma once
#include "common.cuh"

namespace svo {

// ---- warp level -------------------------------------------------------------------------------------
template <class T> SVO_DEV T warp_inclusive_sum(T v, int lane) {
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		T o = __shfl_up_sync(FULL_MASK, v, d);
		if (lane >= d) v += o;
	}
	return v;
}
template <class T> SVO_DEV T warp_sum(T v) {
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
	return v;
}

// ---- block level ------------------------------------------------------------------------------------
// Exclusive sum over the block (all threads call; blockDim.x = BLOCK, a multiple of 32).
// Returns the thread's exclusive prefix; block_total receives the sum on every thread.
template <int BLOCK, class T> SVO_DEV T block_exclusive_sum(T v, T &block_total, T *s_warp /* BLOCK/32 + 1 entries */) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	constexpr int NW = BLOCK / 32;
	T inc = warp_inclusive_sum(v, lane);
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		T w = lane < NW ? s_warp[lane] : T(0);
		T winc = warp_inclusive_sum(w, lane);
		if (lane < NW) s_warp[lane] = winc - w; // exclusive prefix of each warp
		if (lane == NW - 1) s_warp[NW] = winc;  // total
	}
	__syncthreads();
	T excl = s_warp[warp] + inc - v;
	block_total = s_warp[NW];
	__syncthreads(); // s_warp may be reused by the caller right away
	return excl;
}

// ---- chained (single pass) scan ---------------------------------------------------------------------
// One 64-bit word per tile: status in the top 2 bits, value in the low 62, so that a single store
// publishes both (no fence needed between flag and payload).
constexpr uint64_t LB_INVALID = 0, LB_AGGREGATE = 1, LB_PREFIX = 2;
SVO_DEV uint64_t lb_pack(uint64_t status, uint64_t v) { return (status << 62) | (v & 0x3fffffffffffffffull); }
SVO_DEV uint64_t lb_status(uint64_t s) { return s >> 62; }
SVO_DEV uint64_t lb_value(uint64_t s) { return s & 0x3fffffffffffffffull; }
SVO_DEV uint64_t lb_load(const uint64_t *p) { return *reinterpret_cast<const volatile uint64_t *>(p); }
SVO_DEV void lb_store(uint64_t *p, uint64_t v) { *reinterpret_cast<volatile uint64_t *>(p) = v; }

#ifdef SVO_EMU
// test hook: when set, tiles publish aggregates only (never inclusive prefixes), which forces the
// look-back to walk the whole chain in windows of 32 -- the path a sequential emulation never takes.
inline int g_emu_lookback_aggregate_only = 0;
#endif

// Called by one warp of the block (all 32 lanes).  Publishes this tile's aggregate, walks back over the
// predecessors' states, publishes the inclusive prefix and returns the exclusive prefix (valid on every lane).
// state[] must be zero (LB_INVALID) when the kernel starts, and tiles must be numbered in start order (dynamic
// ticket), which guarantees forward progress.
// A step of the walk looks at 32 * LB_WIDE predecessors: every lane has LB_WIDE independent loads in flight, so the
// chain of a scan whose tiles all start together advances 32 * LB_WIDE tiles per memory round trip (with one state
// per lane the three rank scans of 2.8e6 brick records -- 1350 tiles -- took 76 us, 42 round trips of the chain).
#ifndef SVO_LB_WIDE
#define SVO_LB_WIDE 4
#endif
constexpr int LB_WIDE = SVO_LB_WIDE;
SVO_DEV uint64_t lookback_exclusive(uint64_t *state, uint32_t tile, uint64_t aggregate, int lane) {
	if (tile == 0) {
		if (lane == 0) lb_store(&state[0], lb_pack(LB_PREFIX, aggregate));
		return 0;
	}
	if (lane == 0) lb_store(&state[tile], lb_pack(LB_AGGREGATE, aggregate));
	uint64_t exclusive = 0;
	int64_t base = (int64_t)tile - 1;
	for (bool done = false; !done; base -= 32 * LB_WIDE) {
		uint64_t s[LB_WIDE];
#pragma unroll
		for (int j = 0; j < LB_WIDE; ++j) { // nearest predecessors first: sub-window j holds tiles base - 32 j - lane
			const int64_t idx = base - 32 * j - lane;
			s[j] = idx >= 0 ? lb_load(&state[idx]) : lb_pack(LB_PREFIX, 0);
		}
#pragma unroll
		for (int j = 0; j < LB_WIDE; ++j) {
			if (done) break; // (warp-uniform)
			const int64_t idx = base - 32 * j - lane;
			while (__any_sync(FULL_MASK, lb_status(s[j]) == LB_INVALID)) {
#if defined(__CUDA_ARCH__)
				__nanosleep(40); // the predecessors are still working: leave the issue slots to them
#endif
				if (lb_status(s[j]) == LB_INVALID) s[j] = lb_load(&state[idx]);
			}
			const unsigned pmask = __ballot_sync(FULL_MASK, lb_status(s[j]) == LB_PREFIX);
			const int first = pmask ? (__ffs((int)pmask) - 1) : 32;
			const uint64_t v = lane <= first ? lb_value(s[j]) : 0;
			exclusive += warp_sum(v);
			done = pmask != 0;
		}
	}
#ifdef SVO_EMU
	if (g_emu_lookback_aggregate_only) return exclusive;
#endif
	if (lane == 0) lb_store(&state[tile], lb_pack(LB_PREFIX, exclusive + aggregate));
	return exclusive;
}

// Dynamic tile ticket (tile ids in start order).  Returns the tile id on every thread.
SVO_DEV uint32_t take_ticket(uint32_t *counter, uint32_t *s_ticket) {
	if (threadIdx.x == 0) *s_ticket = atomicAdd(counter, 1u);
	__syncthreads();
	uint32_t t = *s_ticket;
	__syncthreads();
	return t;
}

// ---- device-wide exclusive scan of uint32 -> uint64 offsets ------------------------------------------
constexpr int SCAN_BLOCK = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

// out[i] = sum_{j<i} f(in[j]) (64-bit), out[n] = total.  state: tiles+1 words zeroed; ticket zeroed.
// NONZERO = false: f(v) = v; NONZERO = true: f(v) = (v != 0), i.e. the scan numbers the non-zero entries.
// (Two counters are never packed into one scanned word: the look-back keeps 62 value bits per tile, so a packed
// count << 40 | sum silently wrapped at 2^22 counted entries.)
template <class InT, bool NONZERO>
__global__ void __launch_bounds__(SCAN_BLOCK)
    k_exclusive_scan(const InT *__restrict__ in, uint64_t *__restrict__ out, uint64_t n, uint64_t *state, uint32_t *ticket) {
	__shared__ uint64_t s_warp[SCAN_BLOCK / 32 + 1];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_prefix;
	const uint32_t tile = take_ticket(ticket, &s_ticket);
	const uint64_t base = (uint64_t)tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	InT v[SCAN_ITEMS];
	uint64_t sum = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		v[i] = base + i < n ? in[base + i] : InT(0);
		if (NONZERO) v[i] = v[i] != InT(0) ? InT(1) : InT(0);
		sum += v[i];
	}
	uint64_t total;
	uint64_t excl = block_exclusive_sum<SCAN_BLOCK, uint64_t>(sum, total, s_warp);
	if (threadIdx.x < 32) {
		uint64_t p = lookback_exclusive(state, tile, total, threadIdx.x);
		if (threadIdx.x == 0) s_prefix = p;
	}
	__syncthreads();
	uint64_t run = s_prefix + excl;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		if (base + i < n) out[base + i] = run;
		run += v[i];
	}
	// the thread that owns the last element also writes the total
	if (n > 0 && base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = run;
	if (n == 0 && tile == 0 && threadIdx.x == 0) out[0] = 0;
}

struct ScanScratch {
	DevBuf<uint64_t> state;
	DevBuf<uint32_t> ticket;
};

// host wrapper; temp storage is grown on demand and reused
template <class InT, bool NONZERO = false>
inline int exclusive_scan(const InT *in, uint64_t *out, uint64_t n, ScanScratch &sc, cudaStream_t s) {
	const uint32_t tiles = n ? div_up(n, SCAN_TILE) : 1;
	SVO_TRY(sc.state.reserve(tiles + 1, s));
	SVO_TRY(sc.ticket.reserve(1, s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.state.p, 0, (tiles + 1) * sizeof(uint64_t), s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.ticket.p, 0, sizeof(uint32_t), s));
	auto k = k_exclusive_scan<InT, NONZERO>;
	SVO_LAUNCH(tiles, SCAN_BLOCK, 0, s, k, in, out, n, sc.state.p, sc.ticket.p);
	SVO_CUDA_TRY(cudaGetLastError());
	return 0;
}

} // namespace svo
