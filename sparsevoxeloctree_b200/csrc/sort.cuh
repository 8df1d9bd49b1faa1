// sort.cuh -- hand-written "onesweep" least-significant-digit radix sort of 64-bit fragment keys.
//
// Replaces the reference's per-level pointer chase over ALL fragments (octree_tag_node.comp:18-31, run L
// times by OctreeBuilder.cpp:167-209): once the fragments are sorted by Morton code the tree is built
// bottom-up from contiguous runs.  Only the 3*level Morton bits are sorted (bits [24, 24+3L) of a
// fragment); the sort is stable, so fragments of one voxel keep their emission order (the colour
// running average of octree_tag_node.comp:48-57 is order dependent).
//
//   k_radix_histogram : one read of the keys -> digit histograms of every pass (shared-memory bins)
//   k_radix_scan_bins : exclusive scan of each pass's bins
//   k_onesweep_pass   : per pass, ONE read + ONE write of the keys: per-tile ranking with
//                       __match_any_sync, per-digit decoupled look-back across tiles (tiles numbered
//                       by a ticket so predecessors are always running), shared-memory reorder,
//                       coalesced run-wise scatter.
// Traffic: 8*F*(2P+1) bytes for P passes -- HBM bound by design.
#pragma once
#include "scan.cuh"

namespace svo {

constexpr int RADIX_BITS = 8, RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;
constexpr int HIST_BLOCK = 256, HIST_ITEMS = 16;

struct SortPasses {
	uint32_t n_pass;
	uint32_t shift[MAX_PASSES];
	uint32_t mask[MAX_PASSES];
};
inline SortPasses make_passes(uint32_t begin_bit, uint32_t end_bit) {
	SortPasses sp{};
	uint32_t b = begin_bit;
	while (b < end_bit && sp.n_pass < MAX_PASSES) {
		uint32_t w = end_bit - b < (uint32_t)RADIX_BITS ? end_bit - b : (uint32_t)RADIX_BITS;
		sp.shift[sp.n_pass] = b;
		sp.mask[sp.n_pass] = (1u << w) - 1u;
		++sp.n_pass;
		b += w;
	}
	return sp;
}

// ---- histogram of every pass in one read ---------------------------------------------------------------
__global__ void __launch_bounds__(HIST_BLOCK)
    k_radix_histogram(const uint64_t *__restrict__ keys, uint64_t n, SortPasses sp, uint32_t *__restrict__ g_hist /*[pass][RADIX]*/) {
	__shared__ uint32_t s_hist[MAX_PASSES * RADIX];
	for (uint32_t i = threadIdx.x; i < sp.n_pass * RADIX; i += HIST_BLOCK) s_hist[i] = 0;
	__syncthreads();
	const uint64_t per_block = (uint64_t)HIST_BLOCK * HIST_ITEMS;
	for (uint64_t base = (uint64_t)blockIdx.x * per_block; base < n; base += (uint64_t)gridDim.x * per_block) {
#pragma unroll 4
		for (int i = 0; i < HIST_ITEMS; ++i) {
			const uint64_t idx = base + (uint64_t)i * HIST_BLOCK + threadIdx.x;
			const bool ok = idx < n;
			const uint64_t k = ok ? keys[idx] : 0;
			for (uint32_t p = 0; p < sp.n_pass; ++p) {
				const uint32_t d = ok ? ((uint32_t)(k >> sp.shift[p]) & sp.mask[p]) : 0xffffffffu;
				// spatially coherent fragments share their upper digits: one add per warp when uniform
				const uint32_t d0 = __shfl_sync(FULL_MASK, d, 0);
				if (__all_sync(FULL_MASK, d == d0)) {
					if ((threadIdx.x & 31) == 0 && ok) atomicAdd(&s_hist[p * RADIX + d], 32u);
				} else if (ok)
					atomicAdd(&s_hist[p * RADIX + d], 1u);
			}
		}
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < sp.n_pass * RADIX; i += HIST_BLOCK) {
		const uint32_t c = s_hist[i];
		if (c) atomicAdd(&g_hist[i], c);
	}
}

// exclusive scan of each pass's RADIX bins, in place (grid = n_pass blocks of RADIX threads)
__global__ void __launch_bounds__(RADIX) k_radix_scan_bins(uint32_t *g_hist) {
	__shared__ uint32_t s_warp[RADIX / 32 + 1];
	uint32_t *h = g_hist + blockIdx.x * RADIX;
	const uint32_t v = h[threadIdx.x];
	uint32_t total;
	const uint32_t e = block_exclusive_sum<RADIX, uint32_t>(v, total, s_warp);
	h[threadIdx.x] = e;
}

// ---- one onesweep pass ---------------------------------------------------------------------------------
// Look-back state: one word per (tile, digit).  The 2 status bits rotate with the pass number so the state
// array is zeroed once per sort, not once per pass: in pass k a word is "not ready" while it still holds the
// previous pass's final code.
//   pass k: STALE = 2k, AGGREGATE = 2k+1, PREFIX = 2k+2  (mod 4)
template <class StateT> struct LbCodec {
	static constexpr int VBITS = sizeof(StateT) * 8 - 2;
	static constexpr StateT VMASK = (StateT(1) << VBITS) - 1;
	static SVO_DEV StateT pack(uint32_t code, uint64_t v) { return (StateT(code & 3u) << VBITS) | (StateT(v) & VMASK); }
	static SVO_DEV uint32_t code(StateT s) { return (uint32_t)(s >> VBITS); }
	static SVO_DEV uint64_t value(StateT s) { return (uint64_t)(s & VMASK); }
};

template <int BLOCK, int ITEMS, class StateT>
__global__ void __launch_bounds__(BLOCK)
    k_onesweep_pass(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out, uint64_t n, uint32_t shift,
                    uint32_t mask, uint32_t pass, const uint32_t *__restrict__ g_bins /* exclusive, this pass */,
                    StateT *state /*[tiles][RADIX]*/, uint32_t *ticket) {
	static_assert(BLOCK % 32 == 0 && BLOCK >= RADIX, "one thread per digit is assumed");
	constexpr int NW = BLOCK / 32;
	constexpr int TILE = BLOCK * ITEMS;
	constexpr int NB = RADIX + 1; // bin RADIX collects the padding of the last tile
	using LB = LbCodec<StateT>;
	SVO_DYN_SMEM(uint64_t, s_keys);                                     // TILE keys
	uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_keys + TILE);     // NW * NB
	uint32_t *s_tile_off = s_hist + NW * NB;                            // NB: first slot of each digit inside the tile
	uint64_t *s_gofs = reinterpret_cast<uint64_t *>(s_tile_off + NB + (NB & 1)); // RADIX: global offset minus tile offset
	__shared__ uint32_t s_scan[BLOCK / 32 + 1];
	__shared__ uint32_t s_ticket;

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t tile = take_ticket(ticket, &s_ticket);
	const uint64_t tile_base = (uint64_t)tile * TILE;
	const uint32_t tile_count = (uint32_t)(n - tile_base < (uint64_t)TILE ? n - tile_base : (uint64_t)TILE);

	for (int i = threadIdx.x; i < NW * NB; i += BLOCK) s_hist[i] = 0;

	// warp-striped load: warp w owns [w*32*ITEMS, (w+1)*32*ITEMS), item i of lane l is element i*32 + l
	uint64_t key[ITEMS];
	uint32_t rank[ITEMS];
	const uint32_t wbase = warp * 32 * ITEMS;
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t e = wbase + i * 32 + lane;
		key[i] = e < tile_count ? keys_in[tile_base + e] : ~0ull;
	}
	__syncthreads();

	// rank inside the warp, digit by digit group (stable: items in increasing i, lanes in increasing l)
	uint32_t *wh = s_hist + warp * NB;
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t e = wbase + i * 32 + lane;
		const uint32_t d = e < tile_count ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)RADIX;
		const unsigned peers = __match_any_sync(FULL_MASK, d);
		const int leader = __ffs((int)peers) - 1;
		uint32_t base = 0;
		if (lane == leader) {
			base = wh[d];
			wh[d] = base + (uint32_t)__popc(peers);
		}
		base = __shfl_sync(FULL_MASK, base, leader);
		rank[i] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
		__syncwarp();
	}
	__syncthreads();

	// per digit: exclusive prefix over the warps, tile total, look-back over the preceding tiles
	uint32_t bin_total = 0;
	if (threadIdx.x < NB) {
		const uint32_t d = threadIdx.x;
		uint32_t run = 0;
#pragma unroll
		for (int w = 0; w < NW; ++w) {
			const uint32_t c = s_hist[w * NB + d];
			s_hist[w * NB + d] = run;
			run += c;
		}
		bin_total = run;
	}
	// BLOCK >= RADIX; the padding bin (d == RADIX) is handled by thread RADIX when BLOCK > RADIX, else folded below
	uint32_t scan_in = threadIdx.x < RADIX ? bin_total : 0u;
	uint32_t tile_total;
	const uint32_t tile_off = block_exclusive_sum<BLOCK, uint32_t>(scan_in, tile_total, s_scan);
	if (threadIdx.x < RADIX) s_tile_off[threadIdx.x] = tile_off;
	if (threadIdx.x == 0) s_tile_off[RADIX] = tile_total; // padding sorts after every real key
	if (BLOCK == RADIX && threadIdx.x == 0) {
		// padding bin: prefix over warps (only the last tile has padding)
		uint32_t run = 0;
		for (int w = 0; w < NW; ++w) {
			const uint32_t c = s_hist[w * NB + RADIX];
			s_hist[w * NB + RADIX] = run;
			run += c;
		}
	}

	if (threadIdx.x < RADIX) {
		const uint32_t d = threadIdx.x;
		const uint32_t STALE = (2u * pass) & 3u, AGG = (2u * pass + 1u) & 3u, PRE = (2u * pass + 2u) & 3u;
		(void)STALE;
		StateT *my = state + (uint64_t)tile * RADIX + d;
		uint64_t excl = 0;
		if (tile == 0) {
			*reinterpret_cast<volatile StateT *>(my) = LB::pack(PRE, bin_total);
		} else {
			*reinterpret_cast<volatile StateT *>(my) = LB::pack(AGG, bin_total);
			int64_t t = (int64_t)tile - 1;
			for (;;) {
				const StateT s = *reinterpret_cast<const volatile StateT *>(state + (uint64_t)t * RADIX + d);
				const uint32_t c = LB::code(s);
				if (c == PRE) {
					excl += LB::value(s);
					break;
				}
				if (c == AGG) {
					excl += LB::value(s);
					--t; // t >= 0 always: tile 0 publishes PRE
				}
				// otherwise: not published yet in this pass, poll again
			}
#ifdef SVO_EMU
			if (!g_emu_lookback_aggregate_only)
#endif
				*reinterpret_cast<volatile StateT *>(my) = LB::pack(PRE, excl + bin_total);
		}
		s_gofs[d] = (uint64_t)g_bins[d] + excl - (uint64_t)tile_off;
	}
	__syncthreads();

	// reorder through shared memory
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t e = wbase + i * 32 + lane;
		const uint32_t d = e < tile_count ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)RADIX;
		const uint32_t pos = s_tile_off[d] + s_hist[warp * NB + d] + rank[i];
		s_keys[pos] = key[i];
	}
	__syncthreads();

	// scatter: slot idx of the tile goes to g_bins[d] + (keys of digit d in earlier tiles) + (idx - tile_off[d])
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t idx = i * BLOCK + threadIdx.x;
		if (idx < tile_count) {
			const uint64_t k = s_keys[idx];
			const uint32_t d = (uint32_t)(k >> shift) & mask;
			keys_out[s_gofs[d] + idx] = k;
		}
	}
}

template <int BLOCK, int ITEMS> constexpr size_t onesweep_smem_bytes() {
	return (size_t)BLOCK * ITEMS * 8 + (size_t)(BLOCK / 32) * (RADIX + 1) * 4 + (size_t)(RADIX + 2) * 4 + (size_t)RADIX * 8;
}

struct SortScratch {
	DevBuf<uint32_t> hist;   // MAX_PASSES * RADIX
	DevBuf<uint32_t> ticket; // MAX_PASSES
	DevBuf<unsigned char> state;
	bool attr_set = false;
};

constexpr int OS_BLOCK = 256, OS_ITEMS = 16;

// Sorts n keys on bits [begin_bit, end_bit).  Ping-pongs between a and b; *result receives the buffer that
// holds the sorted keys.  Stable.
inline int radix_sort_u64(uint64_t *a, uint64_t *b, uint64_t n, uint32_t begin_bit, uint32_t end_bit, SortScratch &sc,
                          int n_sm, cudaStream_t s, uint64_t **result, uint32_t *n_pass_out, cudaEvent_t ev_after_hist) {
	const SortPasses sp = make_passes(begin_bit, end_bit);
	if (n_pass_out) *n_pass_out = sp.n_pass;
	*result = a;
	if (n <= 1 || sp.n_pass == 0) {
		if (ev_after_hist) SVO_CUDA_TRY(cudaEventRecord(ev_after_hist, s));
		return 0;
	}
	constexpr int TILE = OS_BLOCK * OS_ITEMS;
	const uint32_t tiles = div_up(n, TILE);
	const bool wide = n >= (1ull << 30);
	const size_t state_bytes = (size_t)tiles * RADIX * (wide ? 8 : 4);
	SVO_TRY(sc.hist.reserve(MAX_PASSES * RADIX, s));
	SVO_TRY(sc.ticket.reserve(MAX_PASSES, s));
	SVO_TRY(sc.state.reserve(state_bytes, s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.hist.p, 0, MAX_PASSES * RADIX * sizeof(uint32_t), s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.ticket.p, 0, MAX_PASSES * sizeof(uint32_t), s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.state.p, 0, state_bytes, s));

	uint32_t hgrid = div_up(n, (uint64_t)HIST_BLOCK * HIST_ITEMS);
	const uint32_t hmax = (uint32_t)(n_sm > 0 ? n_sm : 148) * 8u;
	if (hgrid > hmax) hgrid = hmax;
	SVO_LAUNCH(hgrid, HIST_BLOCK, 0, s, k_radix_histogram, (const uint64_t *)a, n, sp, sc.hist.p);
	SVO_LAUNCH(sp.n_pass, RADIX, 0, s, k_radix_scan_bins, sc.hist.p);
	SVO_CUDA_TRY(cudaGetLastError());
	if (ev_after_hist) SVO_CUDA_TRY(cudaEventRecord(ev_after_hist, s));

	constexpr size_t smem = onesweep_smem_bytes<OS_BLOCK, OS_ITEMS>();
	auto k32 = k_onesweep_pass<OS_BLOCK, OS_ITEMS, uint32_t>;
	auto k64 = k_onesweep_pass<OS_BLOCK, OS_ITEMS, uint64_t>;
#ifndef SVO_EMU
	if (!sc.attr_set) {
		SVO_CUDA_TRY(cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		SVO_CUDA_TRY(cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		sc.attr_set = true;
	}
#endif
	uint64_t *src = a, *dst = b;
	for (uint32_t p = 0; p < sp.n_pass; ++p) {
		if (wide)
			SVO_LAUNCH(tiles, OS_BLOCK, smem, s, k64, (const uint64_t *)src, dst, n, sp.shift[p], sp.mask[p], p,
			           (const uint32_t *)(sc.hist.p + p * RADIX), reinterpret_cast<uint64_t *>(sc.state.p), sc.ticket.p + p);
		else
			SVO_LAUNCH(tiles, OS_BLOCK, smem, s, k32, (const uint64_t *)src, dst, n, sp.shift[p], sp.mask[p], p,
			           (const uint32_t *)(sc.hist.p + p * RADIX), reinterpret_cast<uint32_t *>(sc.state.p), sc.ticket.p + p);
		uint64_t *t = src;
		src = dst;
		dst = t;
	}
	SVO_CUDA_TRY(cudaGetLastError());
	*result = src;
	return 0;
}

} // namespace svo
