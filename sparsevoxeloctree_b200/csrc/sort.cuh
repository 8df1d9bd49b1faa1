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
//   k_onesweep_pass   : per pass, ONE read + ONE write of the keys: per-tile ranking by one warp vote per
//                       digit bit + one shared-memory atomic per digit group, per-digit decoupled look-back
//                       across tiles (tiles numbered by a ticket, i.e. in start order, so predecessors are always
//                       running; 4 predecessor states in flight per step), shared-memory reorder, run-wise
//                       coalesced scatter.  (profiles/r02_onesweep_experiments.txt: what else was tried.)
// Digits are up to 9 bits wide (512 bins): 36 Morton bits at level 12 take 4 passes.
// Traffic: 8*F*(2P+1) bytes for P passes -- HBM bound by design.
#pragma once
#include "scan.cuh"

namespace svo {

#ifndef SVO_MAX_RADIX_BITS
#define SVO_MAX_RADIX_BITS 9
#endif
constexpr int MAX_RADIX_BITS = SVO_MAX_RADIX_BITS, MAX_RADIX = 512;
constexpr int MAX_PASSES = 8;
#ifndef SVO_HIST_BLOCK
#define SVO_HIST_BLOCK 1024
#endif
#ifndef SVO_HIST_ITEMS
#define SVO_HIST_ITEMS 8
#endif
#ifndef SVO_HIST_GRID
#define SVO_HIST_GRID 2 // persistent blocks per SM
#endif
constexpr int HIST_BLOCK = SVO_HIST_BLOCK, HIST_ITEMS = SVO_HIST_ITEMS;

struct SortPasses {
	uint32_t n_pass;
	uint32_t max_bits; // widest digit
	uint32_t shift[MAX_PASSES];
	uint32_t bits[MAX_PASSES];
	uint32_t mask[MAX_PASSES];
};
// Fewest passes with digits <= MAX_RADIX_BITS, widths balanced (36 bits -> 9,9,9,9 ; 30 -> 8,8,7,7).
inline SortPasses make_passes(uint32_t begin_bit, uint32_t end_bit) {
	SortPasses sp{};
	const uint32_t total = end_bit - begin_bit;
	if (total == 0) return sp;
	uint32_t np = (total + MAX_RADIX_BITS - 1) / MAX_RADIX_BITS;
	if (np > MAX_PASSES) np = MAX_PASSES; // 64 bits / 9 = 8 passes at most
	uint32_t b = begin_bit;
	for (uint32_t p = 0; p < np; ++p) {
		const uint32_t left = end_bit - b, passes_left = np - p;
		const uint32_t w = (left + passes_left - 1) / passes_left;
		sp.shift[p] = b;
		sp.bits[p] = w;
		sp.mask[p] = (1u << w) - 1u;
		if (w > sp.max_bits) sp.max_bits = w;
		b += w;
	}
	sp.n_pass = np;
	return sp;
}

// ---- histogram of every pass in one read ---------------------------------------------------------------
// Spatially coherent fragments share their upper digits: one OR-reduction of (key ^ lane 0's key) tells, for
// every pass at once, whether the whole warp falls into one bin -- then a single lane adds 32.
// The pass loop is unrolled (NPASS is a template parameter) and runs on the two 32-bit halves of
// key >> shift[0], so a digit costs one funnel shift and one AND instead of a variable 64-bit shift.
// NARROW: every relative shift is below 32 (always the case for up to 4 passes of 9 bits): no second form needed
template <bool NARROW> SVO_DEV uint32_t hist_digit(uint32_t lo, uint32_t hi, uint32_t r /*relative shift, uniform*/, uint32_t mask) {
	return (NARROW || r < 32u ? __funnelshift_r(lo, hi, r) : hi >> (r - 32u)) & mask;
}
// Shared-memory add through a 32-bit shared-window address computed once per thread: taking &s_hist[i] as a generic
// pointer makes the compiler rebuild the window base (S2UR SR_CgaCtaId, ...) in front of every atomic.
SVO_DEV uint32_t shared_address(const void *p) {
#if defined(__CUDA_ARCH__)
	uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
	asm volatile("" : "+r"(a)); // opaque: otherwise the base is rematerialised (4 instructions) at every use
	return a;
#else
	(void)p;
	return 0u;
#endif
}
SVO_DEV void shared_add(uint32_t addr, uint32_t *generic, uint32_t v) {
#if defined(__CUDA_ARCH__)
	(void)generic;
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
#else
	(void)addr;
	atomicAdd(generic, v);
#endif
}
template <int NPASS, bool NARROW>
__global__ void __launch_bounds__(HIST_BLOCK)
    k_radix_histogram(const uint64_t *__restrict__ keys, uint64_t n, SortPasses sp, uint32_t *__restrict__ g_hist /*[pass][MAX_RADIX]*/) {
	__shared__ uint32_t s_hist[NPASS * MAX_RADIX];
	for (uint32_t i = threadIdx.x; i < NPASS * MAX_RADIX; i += HIST_BLOCK) s_hist[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t s0 = sp.shift[0];
	const uint32_t s_base = shared_address(s_hist);
	const uint64_t per_block = (uint64_t)HIST_BLOCK * HIST_ITEMS;
	for (uint64_t base = (uint64_t)blockIdx.x * per_block; base < n; base += (uint64_t)gridDim.x * per_block) {
		const bool full = base + per_block <= n;
#pragma unroll 8
		for (int i = 0; i < HIST_ITEMS; ++i) {
			const uint64_t idx = base + (uint64_t)i * HIST_BLOCK + threadIdx.x;
			const bool ok = full || idx < n;
			const uint64_t m = (ok ? keys[idx] : 0ull) >> s0;
			const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
			const uint32_t pad = ok ? 0u : ~0u;
			const uint32_t dlo = __reduce_or_sync(FULL_MASK, (lo ^ __shfl_sync(FULL_MASK, lo, 0)) | pad);
			const uint32_t dhi = __reduce_or_sync(FULL_MASK, (hi ^ __shfl_sync(FULL_MASK, hi, 0)) | pad);
#pragma unroll
			for (int p = 0; p < NPASS; ++p) {
				const uint32_t r = sp.shift[p] - s0;
				const uint32_t d = hist_digit<NARROW>(lo, hi, r, sp.mask[p]);
				const bool uniform = hist_digit<NARROW>(dlo, dhi, r, sp.mask[p]) == 0u;
				// (grouping by __match_any_sync measured slower: MATCH costs more than the adds)
				if (uniform) {
					if (lane == 0) shared_add(s_base + ((p * MAX_RADIX + d) << 2), &s_hist[p * MAX_RADIX + d], 32u);
				} else if (ok)
					shared_add(s_base + ((p * MAX_RADIX + d) << 2), &s_hist[p * MAX_RADIX + d], 1u);
			}
		}
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < NPASS * MAX_RADIX; i += HIST_BLOCK) {
		const uint32_t c = s_hist[i];
		if (c) atomicAdd(&g_hist[i], c);
	}
}

// exclusive scan of each pass's MAX_RADIX bins, in place (grid = n_pass blocks of MAX_RADIX threads)
__global__ void __launch_bounds__(MAX_RADIX) k_radix_scan_bins(uint32_t *g_hist) {
	__shared__ uint32_t s_warp[MAX_RADIX / 32 + 1];
	uint32_t *h = g_hist + blockIdx.x * MAX_RADIX;
	const uint32_t v = h[threadIdx.x];
	uint32_t total;
	const uint32_t e = block_exclusive_sum<MAX_RADIX, uint32_t>(v, total, s_warp);
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

#ifndef SVO_OS_TICKET
#define SVO_OS_TICKET 1 // 1: tiles are numbered by a ticket (start order); 0: by blockIdx.x (relies on in-order dispatch)
#endif
#ifndef SVO_OS_LB_WINDOW
#define SVO_OS_LB_WINDOW 4 // predecessor states in flight per look-back step (2: 0.931, 4: 0.891, 8: 0.908 ms per pass)
#endif

template <int BLOCK, int ITEMS, int RBITS> struct OnesweepCfg {
	static constexpr int RADIX = 1 << RBITS;
	static constexpr int NB = RADIX + 1; // bin RADIX collects the padding of the last tile
	static constexpr int NW = BLOCK / 32;
	static constexpr int TILE = BLOCK * ITEMS;
	static constexpr int DPT = RADIX > BLOCK ? RADIX / BLOCK : 1; // consecutive digits per digit-thread
	static constexpr int DTHREADS = RADIX / DPT;                  // threads that own digits
	static constexpr int NBP = NB + (NB & 1);
	static constexpr size_t SMEM = (size_t)TILE * 8 + (size_t)NW * NB * 4 + (size_t)NBP * 4 + (size_t)RADIX * 4 + 64;
	static_assert(BLOCK % 32 == 0 && DTHREADS % 32 == 0 && DTHREADS <= BLOCK, "digit threads must be whole warps");
};

// SVO_OS_CLOCKS (experiments only): thread 0 of every tile adds the cycles it spent in each phase to g_os_clocks[]
#ifndef SVO_OS_CLOCKS
#define SVO_OS_CLOCKS 0
#endif
#if SVO_OS_CLOCKS && defined(__CUDACC__)
__device__ unsigned long long g_os_clocks[8];
__device__ unsigned long long g_os_walk[4]; // thread 0: look-back steps, states consumed, empty polls, walks
#define SVO_CLK(i)                                                                     \
	do {                                                                               \
		if (threadIdx.x == 0) {                                                        \
			const long long now = clock64();                                           \
			atomicAdd(&g_os_clocks[i], (unsigned long long)(now - clk_prev));          \
			clk_prev = now;                                                            \
		}                                                                              \
	} while (0)
#define SVO_CLK_INIT long long clk_prev = clock64()
#else
#define SVO_CLK(i) ((void)0)
#define SVO_CLK_INIT ((void)0)
#endif
#ifndef SVO_OS_EXPERIMENT
#define SVO_OS_EXPERIMENT 0 // timing experiments only (bit 0: linear writes, bit 1: no look-back); results are wrong when set
#endif
#ifndef SVO_OS_SPLIT_SMEM
#define SVO_OS_SPLIT_SMEM 1 // 0.873 -> 0.864 ms per pass
#endif
#ifndef SVO_OS_BRANCHY_ATOMIC
#define SVO_OS_BRANCHY_ATOMIC 0
#endif

SVO_DEV void lb_backoff() {
#if defined(__CUDA_ARCH__)
	__nanosleep(64);
#endif
}

// peers &= (lanes whose digit agrees with mine in one bit): test, vote, conditional complement, and -- 4 instructions
SVO_DEV unsigned split_by_bit(unsigned peers, uint32_t d, uint32_t bitmask) {
#if defined(__CUDA_ARCH__)
	asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 t, p, 0xffffffff;\n\t"
	    "@!p not.b32 t, t;\n\tand.b32 %0, %0, t;\n\t}"
	    : "+r"(peers)
	    : "r"(d), "r"(bitmask));
	return peers;
#else
	const bool p = (d & bitmask) != 0u;
	const unsigned b = __ballot_sync(FULL_MASK, p);
	return peers & (p ? b : ~b);
#endif
}
// the lanes of the warp whose NBITS-bit digit equals mine: one vote per digit bit, straight-line code (rounds over the
// distinct values of the row -- cheaper in instructions for coherent digits -- measured slower: their data-dependent
// branches keep the compiler from interleaving the items of a thread)
template <int NBITS> SVO_DEV unsigned warp_peers(uint32_t d) {
	unsigned peers = FULL_MASK;
#pragma unroll
	for (int bb = 0; bb < NBITS; ++bb) peers = split_by_bit(peers, d, 1u << bb);
	return peers;
}

// Shared-memory atomic add executed by the group leaders only, as ONE predicated instruction: written as a branch
// the compiler wraps it in a divergence region (BSSY / BRA / BSYNC) per item.
SVO_DEV uint32_t leader_atomic_add(uint32_t *addr, uint32_t v, bool leader) {
#if defined(__CUDA_ARCH__) && !SVO_OS_BRANCHY_ATOMIC
	uint32_t old = 0;
	const uint32_t a = (uint32_t)__cvta_generic_to_shared(addr);
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
	             : "+r"(old)
	             : "r"(a), "r"(v), "r"((uint32_t)leader)
	             : "memory");
	return old;
#else
	return leader ? atomicAdd(addr, v) : 0u;
#endif
}

struct PassArgs {
	const uint64_t *in;
	uint64_t *out;
	uint64_t n;
	uint32_t shift, mask;
	const uint32_t *bins; // exclusive digit offsets of this pass
	void *state;          // [tiles][RADIX] look-back words
	uint32_t *ticket;     // zeroed; tiles are handed out in start order
	uint32_t pass;        // pass number (the look-back status codes rotate with it)
};

// One tile.  FULL = the tile holds TILE keys (every tile but possibly the last): no bounds checks, no padding bin.
template <int BLOCK, int ITEMS, int RBITS, class StateT, bool FULL>
SVO_DEV void onesweep_tile(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out, uint32_t tile, uint32_t tile_count,
                           uint32_t shift, uint32_t mask, uint32_t pass, const uint32_t *__restrict__ g_bins, StateT *state,
                           uint64_t *s_keys, uint32_t *s_wsum) {
	using C = OnesweepCfg<BLOCK, ITEMS, RBITS>;
	constexpr int RADIX = C::RADIX, NB = C::NB, NW = C::NW, TILE = C::TILE, DPT = C::DPT, DTHREADS = C::DTHREADS;
	using LB = LbCodec<StateT>;
	uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_keys + TILE);  // NW * NB: per-warp digit counters, then warp prefixes
	uint32_t *s_tile_off = s_hist + NW * NB;                         // NB: first slot of each digit inside the tile
	uint32_t *s_gofs = s_tile_off + C::NBP;                          // RADIX: global offset minus tile offset (mod 2^32)
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t wbase = warp * 32 * ITEMS;

	// warp-striped load: warp w owns [w*32*ITEMS, (w+1)*32*ITEMS), item i of lane l is element i*32 + l
	uint64_t key[ITEMS];
	{
		const uint64_t *src = keys_in + (uint64_t)tile * TILE + wbase + lane;
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) key[i] = (FULL || wbase + i * 32 + lane < tile_count) ? src[i * 32] : ~0ull;
	}
	for (int i = threadIdx.x; i < NW * NB; i += BLOCK) s_hist[i] = 0; // overlaps the loads in flight
	const uint32_t AGG = (2u * pass + 1u) & 3u, PRE = (2u * pass + 2u) & 3u;
	uint32_t bin_total[DPT];
	uint32_t my_sum = 0, inc = 0;
	SVO_CLK_INIT;
	__syncthreads();
	SVO_CLK(0);

	// rank inside the warp.  Stable: items in increasing i, lanes in increasing l.  The leader (lowest lane) of
	// each digit group bumps the warp's counter with ONE shared-memory atomic; a warp's atomics execute in
	// program order, so no warp barrier is needed between items and the stages pipeline across items.
	uint32_t *wh = s_hist + warp * NB;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t rank[ITEMS];
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t d = (FULL || wbase + i * 32 + lane < tile_count) ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)RADIX;
		const unsigned peers = warp_peers<RBITS + (FULL ? 0 : 1)>(d); // (bit RBITS: the padding bin of a partial tile)
		const uint32_t below = (uint32_t)__popc(peers & lt_mask);
		const uint32_t base = leader_atomic_add(&wh[d], (uint32_t)__popc(peers), below == 0u);
		SVO_EMU_WARP_ORDER(); // hardware issues a warp's atomics in program order; the emulator's lanes are free-running
		rank[i] = __shfl_sync(FULL_MASK, base, __ffs((int)peers) - 1) + below;
	}
	SVO_CLK(1);
	__syncthreads();
	SVO_CLK(2);

	// per digit: exclusive prefix over the warps and the tile total; publish the aggregate right away
	if (threadIdx.x < DTHREADS) {
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			const uint32_t d = threadIdx.x * DPT + j;
			uint32_t run = 0;
#pragma unroll
			for (int w = 0; w < NW; ++w) {
				const uint32_t c = s_hist[w * NB + d];
				s_hist[w * NB + d] = run;
				run += c;
			}
			bin_total[j] = run;
			my_sum += run;
			*reinterpret_cast<volatile StateT *>(state + (uint64_t)tile * RADIX + d) = LB::pack(tile == 0 ? PRE : AGG, run);
		}
	}
	if (!FULL && threadIdx.x == BLOCK - 1) { // padding bin of the last tile: prefix over the warps
		uint32_t run = 0;
		for (int w = 0; w < NW; ++w) {
			const uint32_t c = s_hist[w * NB + RADIX];
			s_hist[w * NB + RADIX] = run;
			run += c;
		}
	}
	// exclusive scan of the tile totals over the digits (digit threads are whole warps)
	if (threadIdx.x < DTHREADS) {
		inc = warp_inclusive_sum(my_sum, lane);
		if (lane == 31) s_wsum[warp] = inc;
	}
	__syncthreads();
	if (threadIdx.x < DTHREADS) {
		uint32_t wpre = 0;
#pragma unroll
		for (int w = 0; w < DTHREADS / 32; ++w) wpre += w < warp ? s_wsum[w] : 0u;
		uint32_t run = wpre + inc - my_sum; // first slot of this thread's first digit
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			s_tile_off[threadIdx.x * DPT + j] = run;
			run += bin_total[j];
		}
	}
	if (!FULL && threadIdx.x == 0) s_tile_off[RADIX] = tile_count; // padding sorts after every real key
	__syncthreads();
	SVO_CLK(3);

	// reorder through shared memory (needs tile-local offsets only: runs while predecessors publish)
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t d = (FULL || wbase + i * 32 + lane < tile_count) ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)RADIX;
#if SVO_OS_SPLIT_SMEM
		{ // low and high halves in separate arrays: 32-bit stores to random slots conflict less than 64-bit ones
			const uint32_t pos = s_tile_off[d] + s_hist[warp * NB + d] + rank[i];
			reinterpret_cast<uint32_t *>(s_keys)[pos] = (uint32_t)key[i];
			reinterpret_cast<uint32_t *>(s_keys)[TILE + pos] = (uint32_t)(key[i] >> 32);
		}
#else
		s_keys[s_tile_off[d] + s_hist[warp * NB + d] + rank[i]] = key[i];
#endif
	}
	SVO_CLK(4);

	// Decoupled look-back, one thread per digit, SVO_OS_LB_WINDOW predecessor states in flight per step.  (Measured: the
	// walks are short -- ~20 predecessors, 5 steps -- and never find an unpublished state; what they cost is the
	// >1000 cycles every dependent global load takes on an SM whose memory queue is full of key traffic.  Starting the
	// walk before the reorder, wider windows, one polling warp per tile and dedicated scanner blocks were all slower.)
	constexpr int W = SVO_OS_LB_WINDOW;
	if (threadIdx.x < DTHREADS) {
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			const uint32_t d = threadIdx.x * DPT + j;
			uint32_t excl = 0; // offsets are taken mod 2^32 (n < 2^32)
			if (tile != 0 && !(SVO_OS_EXPERIMENT & 2)) {
				const volatile StateT *p = state + (size_t)(tile - 1) * RADIX + d; // nearest predecessor not yet summed
				uint32_t left = tile;                                            // predecessors from p backwards (>= 1)
				for (;;) {
					StateT st[W];
#pragma unroll
					for (int q = 0; q < W; ++q)
						if ((uint32_t)q < left) st[q] = p[-(ptrdiff_t)q * RADIX];
					uint32_t used = 0;
					bool done = false;
#pragma unroll
					for (int q = 0; q < W; ++q) {
						if (done || (uint32_t)q >= left || used != (uint32_t)q) break;
						const uint32_t c = LB::code(st[q]);
						if (c != PRE && c != AGG) break; // not published yet: poll from here again
						excl += (uint32_t)LB::value(st[q]);
						++used;
						done = c == PRE;
					}
#if SVO_OS_CLOCKS && defined(__CUDACC__)
					if (threadIdx.x == 0) {
						atomicAdd(&g_os_walk[0], 1ull), atomicAdd(&g_os_walk[1], (unsigned long long)used);
						if (used == 0u) atomicAdd(&g_os_walk[2], 1ull);
						if (done) atomicAdd(&g_os_walk[3], 1ull);
					}
#endif
					if (done) break;
					if (used == 0u) lb_backoff();
					p -= (ptrdiff_t)used * RADIX, left -= used; // tile 0 always holds a prefix: left never reaches 0
				}
#ifdef SVO_EMU
				if (!g_emu_lookback_aggregate_only)
#endif
					*reinterpret_cast<volatile StateT *>(state + (size_t)tile * RADIX + d) = LB::pack(PRE, (uint64_t)excl + bin_total[j]);
			}
			s_gofs[d] = g_bins[d] + excl - s_tile_off[d];
		}
	}
	SVO_CLK(5);
	__syncthreads();
	SVO_CLK(6);

	// scatter: slot idx of the tile goes to bins[d] + (keys of digit d in earlier tiles) + (idx - tile_off[d])
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t idx = i * BLOCK + threadIdx.x;
		if (FULL || idx < tile_count) {
#if SVO_OS_SPLIT_SMEM
			const uint64_t k = (uint64_t)reinterpret_cast<const uint32_t *>(s_keys)[idx] |
			                   ((uint64_t)reinterpret_cast<const uint32_t *>(s_keys)[TILE + idx] << 32);
#else
			const uint64_t k = s_keys[idx];
#endif
			const uint32_t d = (uint32_t)(k >> shift) & mask;
#if (SVO_OS_EXPERIMENT & 1)
			keys_out[(uint64_t)tile * TILE + idx] = k; (void)d;
#else
			keys_out[(uint32_t)(s_gofs[d] + idx)] = k;
#endif
		}
	}
	SVO_CLK(7);
}

// Tiles are numbered in start order by a ticket, so every predecessor of a running tile has been started (and
// running blocks are never preempted) -- the forward progress the look-back needs, without relying on the order in
// which the hardware dispatches the blocks of a grid (+2.9 % per pass against numbering by blockIdx.x).
template <int BLOCK, int ITEMS, int RBITS, int MINB, class StateT>
__global__ void __launch_bounds__(BLOCK, MINB) k_onesweep_pass(PassArgs pa) {
	using C = OnesweepCfg<BLOCK, ITEMS, RBITS>;
	SVO_DYN_SMEM(uint64_t, s_keys);
	__shared__ uint32_t s_wsum[C::RADIX / 32 + 1];
	const uint64_t n = pa.n;
#if SVO_OS_TICKET
	__shared__ uint32_t s_tile;
	if (threadIdx.x == 0) s_tile = atomicAdd(pa.ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
#else
	const uint32_t tile = blockIdx.x;
#endif
	const uint64_t tile_base = (uint64_t)tile * C::TILE;
	const uint32_t tile_count = (uint32_t)(n - tile_base < (uint64_t)C::TILE ? n - tile_base : (uint64_t)C::TILE);
	StateT *state = reinterpret_cast<StateT *>(pa.state);
	if (tile_count == (uint32_t)C::TILE)
		onesweep_tile<BLOCK, ITEMS, RBITS, StateT, true>(pa.in, pa.out, tile, tile_count, pa.shift, pa.mask, pa.pass, pa.bins, state, s_keys, s_wsum);
	else
		onesweep_tile<BLOCK, ITEMS, RBITS, StateT, false>(pa.in, pa.out, tile, tile_count, pa.shift, pa.mask, pa.pass, pa.bins, state, s_keys, s_wsum);
}

inline bool g_force_wide_sort_state = false; // svo_debug_force_wide_sort_state (tests)
inline bool g_profile_passes = false;        // svo_debug_profile_passes: an event after every sort kernel

constexpr int SORT_MAX_EVENTS = 2 * MAX_PASSES + 8;
struct SortScratch {
	DevBuf<uint32_t> hist;    // MAX_PASSES * MAX_RADIX digit bins (+ tickets behind them)
	DevBuf<unsigned char> state;
	cudaEvent_t ev[SORT_MAX_EVENTS] = {};
	int n_ev = 0; // events recorded by the last sort (profiling only)
	void release(cudaStream_t s) {
		hist.release(s), state.release(s);
		for (auto &e : ev)
			if (e) cudaEventDestroy(e), e = nullptr;
	}
	int mark(cudaStream_t s) { // profiling: one more event on the stream
		if (!g_profile_passes || n_ev >= SORT_MAX_EVENTS) return 0;
		if (!ev[n_ev]) SVO_CUDA_TRY(cudaEventCreate(&ev[n_ev]));
		SVO_CUDA_TRY(cudaEventRecord(ev[n_ev++], s));
		return 0;
	}
};
constexpr uint32_t SORT_HIST_WORDS = MAX_PASSES * MAX_RADIX, SORT_TICKETS = 16; // tickets live behind the bins

// tuning point (tests/bench can override at compile time)
#ifndef SVO_OS_BLOCK
#define SVO_OS_BLOCK 256
#endif
#ifndef SVO_OS_ITEMS
#define SVO_OS_ITEMS 22
#endif
#ifndef SVO_OS_MINB
#define SVO_OS_MINB 3
#endif
constexpr int OS_BLOCK = SVO_OS_BLOCK, OS_ITEMS = SVO_OS_ITEMS, OS_MINB = SVO_OS_MINB, OS_TILE = OS_BLOCK * OS_ITEMS;

template <int RBITS, class StateT> inline int launch_onesweep_pass(const PassArgs &pa, int device, cudaStream_t s) {
	using C = OnesweepCfg<OS_BLOCK, OS_ITEMS, RBITS>;
	auto k = k_onesweep_pass<OS_BLOCK, OS_ITEMS, RBITS, OS_MINB, StateT>;
#ifndef SVO_EMU
	static bool attr_set[64] = {}; // per device: a single-process multi-GPU host launches on every device
	const int di = device >= 0 && device < 64 ? device : 0;
	if (!attr_set[di]) {
		SVO_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
		attr_set[di] = true;
	}
#else
	(void)device;
#endif
	SVO_LAUNCH(div_up(pa.n, OS_TILE), OS_BLOCK, C::SMEM, s, k, pa);
	return 0;
}
inline int launch_onesweep(bool nine, bool wide, const PassArgs &pa, int device, cudaStream_t s) {
	if (nine) return wide ? launch_onesweep_pass<9, uint64_t>(pa, device, s) : launch_onesweep_pass<9, uint32_t>(pa, device, s);
	return wide ? launch_onesweep_pass<8, uint64_t>(pa, device, s) : launch_onesweep_pass<8, uint32_t>(pa, device, s);
}

inline int launch_radix_histogram(const uint64_t *keys, uint64_t n, const SortPasses &sp, uint32_t *hist, int n_sm, cudaStream_t s) {
	uint32_t hgrid = div_up(n, (uint64_t)HIST_BLOCK * HIST_ITEMS);
	const uint32_t hmax = (uint32_t)(n_sm > 0 ? n_sm : 148) * (uint32_t)SVO_HIST_GRID;
	if (hgrid > hmax) hgrid = hmax;
	const bool narrow = sp.shift[sp.n_pass - 1] - sp.shift[0] < 32u;
	switch (sp.n_pass) {
#define SVO_HIST_CASE(NP)                                                                                                   \
	case NP: {                                                                                                              \
		if (narrow) {                                                                                                       \
			SVO_LAUNCH(hgrid, HIST_BLOCK, 0, s, (k_radix_histogram<NP, true>), keys, n, sp, hist);   \
		} else {                                                                                                            \
			SVO_LAUNCH(hgrid, HIST_BLOCK, 0, s, (k_radix_histogram<NP, false>), keys, n, sp, hist);  \
		}                                                                                                                   \
		break;                                                                                                              \
	}
		SVO_HIST_CASE(1) SVO_HIST_CASE(2) SVO_HIST_CASE(3) SVO_HIST_CASE(4) SVO_HIST_CASE(5) SVO_HIST_CASE(6) SVO_HIST_CASE(7)
		SVO_HIST_CASE(8)
#undef SVO_HIST_CASE
	}
	return 0;
}

// Sorts n keys on bits [begin_bit, end_bit).  Ping-pongs between a and b; *result receives the buffer that
// holds the sorted keys.  Stable.
inline int radix_sort_u64(uint64_t *a, uint64_t *b, uint64_t n, uint32_t begin_bit, uint32_t end_bit, SortScratch &sc, int device,
                          int n_sm, cudaStream_t s, uint64_t **result, uint32_t *n_pass_out, cudaEvent_t ev_after_hist) {
	const SortPasses sp = make_passes(begin_bit, end_bit);
	if (n_pass_out) *n_pass_out = sp.n_pass;
	*result = a;
	if (n <= 1 || sp.n_pass == 0) {
		if (ev_after_hist) SVO_CUDA_TRY(cudaEventRecord(ev_after_hist, s));
		return 0;
	}
	if (n >= (1ull << 32)) {
		set_error("radix_sort_u64: more than 2^32-1 keys");
		return -4;
	}
	const uint32_t tiles = div_up(n, OS_TILE);
	const bool wide = n >= (1ull << 30) || g_force_wide_sort_state;
	const bool nine = sp.max_bits > 8;
	const uint32_t radix = nine ? 512u : 256u;
	const size_t state_bytes = (size_t)tiles * radix * (wide ? 8 : 4);
	SVO_TRY(sc.hist.reserve(SORT_HIST_WORDS + SORT_TICKETS, s));
	SVO_TRY(sc.state.reserve(state_bytes, s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.hist.p, 0, (SORT_HIST_WORDS + SORT_TICKETS) * sizeof(uint32_t), s));
	SVO_CUDA_TRY(cudaMemsetAsync(sc.state.p, 0, state_bytes, s));
	SVO_TRY(sc.mark(s));
	SVO_TRY(launch_radix_histogram(a, n, sp, sc.hist.p, n_sm, s));
	SVO_LAUNCH(sp.n_pass, MAX_RADIX, 0, s, k_radix_scan_bins, sc.hist.p);
	SVO_CUDA_TRY(cudaGetLastError());
	if (ev_after_hist) SVO_CUDA_TRY(cudaEventRecord(ev_after_hist, s));
	SVO_TRY(sc.mark(s));

	uint64_t *src = a, *dst = b;
	for (uint32_t p = 0; p < sp.n_pass; ++p) {
		PassArgs pa{};
		pa.in = src, pa.out = dst, pa.n = n;
		pa.shift = sp.shift[p], pa.mask = sp.mask[p];
		pa.bins = sc.hist.p + p * MAX_RADIX;
		pa.state = sc.state.p;
		pa.ticket = sc.hist.p + SORT_HIST_WORDS + p;
		pa.pass = p;
		SVO_TRY(launch_onesweep(nine, wide, pa, device, s));
		SVO_TRY(sc.mark(s));
		uint64_t *t = src;
		src = dst;
		dst = t;
	}
	SVO_CUDA_TRY(cudaGetLastError());
	*result = src;
	return 0;
}

} // namespace svo
