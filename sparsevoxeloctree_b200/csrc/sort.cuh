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
//   k_onesweep_pass   : per pass, ONE read + ONE write of the keys.  Persistent thread blocks draw tiles from a
//                       ticket (tile ids in start order: the forward progress the look-back needs, whatever order
//                       the hardware dispatches blocks in); the next tile is pulled into shared memory by one bulk
//                       asynchronous copy (cp.async.bulk + mbarrier) while the current one is ranked; ranking by
//                       warp votes (one round per DISTINCT digit value in a row of 32 keys -- spatially coherent
//                       fragments share their upper digits -- falling back to one vote per digit bit) + one
//                       shared-memory atomic per digit group; decoupled look-back across tiles done by ONE warp
//                       per tile with 16-byte loads of the predecessors' state rows; shared-memory reorder,
//                       run-wise coalesced scatter.
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

#ifndef SVO_OS_TMA
#define SVO_OS_TMA 1 // 1: the next tile is prefetched into shared memory by cp.async.bulk; 0: plain global loads
#endif
#ifndef SVO_OS_LB_WARP
#define SVO_OS_LB_WARP 1 // 1: one warp per tile walks the predecessors with 16-byte loads; 0: one thread per digit
#endif
#ifndef SVO_OS_LB_WINDOW
#define SVO_OS_LB_WINDOW 2 // predecessor rows in flight per look-back step (warp look-back)
#endif
#ifndef SVO_OS_RANK_LOOP
#define SVO_OS_RANK_LOOP 4 // rounds of the distinct-value ranking before the per-bit votes take over (0: votes only)
#endif

template <int BLOCK, int ITEMS, int RBITS, class StateT> struct OnesweepCfg {
	static constexpr int RADIX = 1 << RBITS;
	static constexpr int NB = RADIX + 1; // bin RADIX collects the padding of the last tile
	static constexpr int NW = BLOCK / 32;
	static constexpr int TILE = BLOCK * ITEMS;
	static constexpr int DPT = RADIX > BLOCK ? RADIX / BLOCK : 1; // consecutive digits per digit-thread
	static constexpr int DTHREADS = RADIX / DPT;                  // threads that own digits
	static constexpr int NBP = NB + (NB & 1);
	static constexpr int DPL = RADIX / 32; // digits per lane of the look-back warp
	// dynamic shared memory, in this order (every piece a multiple of 16 bytes)
	static constexpr size_t OFF_STAGE = 0;                                           // TILE keys: bulk-copy landing buffer
	static constexpr size_t OFF_KEYS = SVO_OS_TMA ? (size_t)TILE * 8 : 0;            // TILE keys: reorder buffer (two 32-bit halves)
	static constexpr size_t OFF_HIST = OFF_KEYS + (size_t)TILE * 8;                  // NW * NB counters (+ pad)
	static constexpr size_t HIST_BYTES = (((size_t)NW * NB * 4) + 15) & ~size_t(15);
	static constexpr size_t OFF_TOFF = OFF_HIST + HIST_BYTES;                        // NBP
	static constexpr size_t OFF_GOFS = OFF_TOFF + (((size_t)NBP * 4 + 15) & ~size_t(15)); // RADIX
	static constexpr size_t OFF_TOTAL = OFF_GOFS + (size_t)RADIX * 4;                // RADIX
	static constexpr size_t SMEM = OFF_TOTAL + (size_t)RADIX * 4;
	static_assert(BLOCK % 32 == 0 && DTHREADS % 32 == 0 && DTHREADS <= BLOCK, "digit threads must be whole warps");
	static_assert(DPL * sizeof(StateT) % 16 == 0, "a lane's states are read with 16-byte loads");
};

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
	__nanosleep(40);
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

// The lanes of the warp whose digit equals mine.  Fragments arrive in raster order, so the 32 keys of a row mostly
// share their upper digits: up to LOOP rounds of "take the first unmatched lane's value, vote on equality" (6
// instructions per distinct value; the loop condition is warp-uniform) before falling back to one vote per digit bit
// (4 instructions per bit, whatever the values).
template <int NBITS, int LOOP> SVO_DEV unsigned warp_peers(uint32_t d) {
	unsigned peers = 0, rem = FULL_MASK;
#pragma unroll
	for (int it = 0; it < LOOP; ++it) {
		if (rem == 0u) break;
		const uint32_t dv = __shfl_sync(FULL_MASK, d, __ffs((int)rem) - 1);
		const unsigned m = __ballot_sync(FULL_MASK, d == dv);
		if (d == dv) peers = m;
		rem &= ~m;
	}
	if (rem != 0u) {
		peers = FULL_MASK;
#pragma unroll
		for (int bb = 0; bb < NBITS; ++bb) peers = split_by_bit(peers, d, 1u << bb);
	}
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

// ---- bulk asynchronous copy (TMA unit, 1-D) + mbarrier --------------------------------------------------------
// global -> shared copy of `bytes` (a multiple of 16; both addresses 16-byte aligned) issued by ONE thread; completion
// is signalled on an mbarrier as a transaction count.  The CPU emulator copies synchronously.
SVO_DEV void mbar_init(uint64_t *mbar, uint32_t count) {
#if defined(__CUDA_ARCH__)
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
	*mbar = 0;
	(void)count;
#endif
}
SVO_DEV void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *mbar) {
#if defined(__CUDA_ARCH__)
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
	const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
	// the generic-proxy reads of the landing buffer (previous tile) are ordered before the async-proxy writes
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
	             "l"((uint64_t)__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(bar)
	             : "memory");
#else
	memcpy(smem_dst, gmem_src, bytes);
	(void)mbar;
#endif
}
SVO_DEV void mbar_wait(uint64_t *mbar, uint32_t parity) {
#if defined(__CUDA_ARCH__)
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
	             "r"(parity)
	             : "memory");
#else
	(void)mbar, (void)parity;
#endif
}

// 16-byte loads / stores of look-back states that must not be served from a stale cache line
template <class StateT> SVO_DEV void load_states16(const StateT *p, StateT *out /* 16 / sizeof(StateT) values */) {
#if defined(__CUDA_ARCH__)
	if (sizeof(StateT) == 4) {
		uint32_t a, b, c, d;
		asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
		out[0] = (StateT)a, out[1] = (StateT)b, out[2] = (StateT)c, out[3] = (StateT)d;
	} else {
		uint64_t a, b;
		asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		out[0] = (StateT)a, out[1] = (StateT)b;
	}
#else
	for (unsigned i = 0; i < 16 / sizeof(StateT); ++i) out[i] = *reinterpret_cast<const volatile StateT *>(p + i);
#endif
}
template <class StateT> SVO_DEV void store_states16(StateT *p, const StateT *v) {
#if defined(__CUDA_ARCH__)
	if (sizeof(StateT) == 4)
		asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)v[0]), "r"((uint32_t)v[1]), "r"((uint32_t)v[2]),
		             "r"((uint32_t)v[3])
		             : "memory");
	else
		asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((uint64_t)v[0]), "l"((uint64_t)v[1]) : "memory");
#else
	for (unsigned i = 0; i < 16 / sizeof(StateT); ++i) *reinterpret_cast<volatile StateT *>(p + i) = v[i];
#endif
}

struct PassArgs {
	const uint64_t *in;
	uint64_t *out;
	uint64_t n;
	uint32_t tiles;
	uint32_t shift, mask;
	const uint32_t *bins; // exclusive digit offsets of this pass
	void *state;          // [tiles][RADIX] look-back words
	uint32_t *ticket;     // zeroed; tiles are handed out in start order
	// Device-side mode switch of the hybrid sort (bucket.cuh): *mode is 0 (hybrid) or 1 (classic: a bucket was too
	// large); nullptr counts as 0.  The pass runs when bit *mode of run_mask is set, as pass number pass_in_mode[*mode]
	// among the passes that run in that mode (the look-back status codes rotate with it).
	const uint32_t *mode;
	uint32_t run_mask;
	uint32_t pass_in_mode[2];
	uint32_t use_bulk; // the input is 16-byte aligned: whole tiles arrive by cp.async.bulk
};

// One tile.  FULL = the tile holds TILE keys (every tile but possibly the last): no bounds checks, no padding bin.
// STAGED = its keys already sit in the landing buffer s_stage (bulk copy), else they are loaded from global memory.
// after_load() is called by every thread once the whole block has its keys in registers (the landing buffer is free).
template <int BLOCK, int ITEMS, int RBITS, class StateT, bool FULL, class AfterLoad>
SVO_DEV void onesweep_tile(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out, uint32_t shift, uint32_t mask,
                           const uint32_t *__restrict__ g_bins, StateT *state, uint32_t pass, uint32_t tile, uint32_t tile_count, bool staged,
                           unsigned char *smem, uint32_t *s_wsum, AfterLoad after_load) {
	using C = OnesweepCfg<BLOCK, ITEMS, RBITS, StateT>;
	constexpr int RADIX = C::RADIX, NB = C::NB, NW = C::NW, TILE = C::TILE, DPT = C::DPT, DTHREADS = C::DTHREADS, DPL = C::DPL;
	using LB = LbCodec<StateT>;
	const uint64_t *s_stage = reinterpret_cast<const uint64_t *>(smem + C::OFF_STAGE);
	uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem + C::OFF_KEYS);
	uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + C::OFF_HIST);     // NW * NB: per-warp digit counters, then warp prefixes
	uint32_t *s_tile_off = reinterpret_cast<uint32_t *>(smem + C::OFF_TOFF); // NB: first slot of each digit inside the tile
	uint32_t *s_gofs = reinterpret_cast<uint32_t *>(smem + C::OFF_GOFS);     // RADIX: global offset minus tile offset (mod 2^32)
	uint32_t *s_total = reinterpret_cast<uint32_t *>(smem + C::OFF_TOTAL);   // RADIX: keys of each digit in this tile
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t wbase = warp * 32 * ITEMS;

	// warp-striped: warp w owns [w*32*ITEMS, (w+1)*32*ITEMS), item i of lane l is element i*32 + l
	uint64_t key[ITEMS];
	if (SVO_OS_TMA && staged) {
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) key[i] = s_stage[wbase + i * 32 + lane];
	} else {
		const uint64_t *src = keys_in + (uint64_t)tile * TILE + wbase + lane;
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) key[i] = (FULL || wbase + i * 32 + lane < tile_count) ? src[i * 32] : ~0ull;
	}
	for (int i = threadIdx.x; i < NW * NB; i += BLOCK) s_hist[i] = 0; // overlaps the loads in flight
	const uint32_t AGG = (2u * pass + 1u) & 3u, PRE = (2u * pass + 2u) & 3u;
	uint32_t bin_total[DPT];
	uint32_t my_sum = 0, inc = 0;
	__syncthreads();
	after_load();

	// rank inside the warp.  Stable: items in increasing i, lanes in increasing l.  The leader (lowest lane) of
	// each digit group bumps the warp's counter with ONE shared-memory atomic; a warp's atomics execute in
	// program order, so no warp barrier is needed between items and the stages pipeline across items.
	uint32_t *wh = s_hist + warp * NB;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t rank[ITEMS];
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t d = (FULL || wbase + i * 32 + lane < tile_count) ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)RADIX;
		const unsigned peers = warp_peers<RBITS + (FULL ? 0 : 1), SVO_OS_RANK_LOOP>(d); // (bit RBITS: the padding bin of a partial tile)
		const uint32_t below = (uint32_t)__popc(peers & lt_mask);
		const uint32_t base = leader_atomic_add(&wh[d], (uint32_t)__popc(peers), below == 0u);
		SVO_EMU_WARP_ORDER(); // hardware issues a warp's atomics in program order; the emulator's lanes are free-running
		rank[i] = __shfl_sync(FULL_MASK, base, __ffs((int)peers) - 1) + below;
	}
	__syncthreads();

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
			if (SVO_OS_LB_WARP) s_total[d] = run;
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

#if SVO_OS_LB_WARP
	// Decoupled look-back by ONE warp: lane l owns digits [l*DPL, (l+1)*DPL) and reads them from a predecessor's state
	// row with 16-byte loads, SVO_OS_LB_WINDOW rows in flight.  A tile publishes a whole row at a time (aggregates after
	// its ranking, prefixes after its own look-back), so rows are consumed whole: the walk stops at the first row that
	// holds prefixes.  The other warps wait at the barrier below instead of polling.
	if (warp == 0) {
		constexpr int VPS = 16 / (int)sizeof(StateT); // states per 16-byte load
		constexpr int W = SVO_OS_LB_WINDOW;
		uint32_t excl[DPL];
#pragma unroll
		for (int j = 0; j < DPL; ++j) excl[j] = 0; // offsets are taken mod 2^32 (n < 2^32)
		if (tile != 0 && !(SVO_OS_EXPERIMENT & 2)) {
			uint32_t p = tile - 1;                                      // nearest predecessor not yet summed
			uint32_t open = DPL == 32 ? ~0u : ((1u << DPL) - 1u);       // this lane's digits that still lack a prefix
			bool walking = true;
			while (walking) {
				StateT st[W][DPL];
#pragma unroll
				for (int w = 0; w < W; ++w)
					if ((uint32_t)w <= p) {
						const StateT *row = state + (size_t)(p - w) * RADIX + lane * DPL;
#pragma unroll
						for (int j = 0; j < DPL; j += VPS) load_states16(row + j, &st[w][j]);
					}
				uint32_t consumed = 0;
#pragma unroll
				for (int w = 0; w < W; ++w) {
					if (!walking || (uint32_t)w > p) break; // warp-uniform
					bool ready = true;
#pragma unroll
					for (int j = 0; j < DPL; ++j) {
						const uint32_t c = LB::code(st[w][j]);
						ready = ready && (!((open >> j) & 1u) || c == AGG || c == PRE);
					}
					if (!__all_sync(FULL_MASK, ready)) break; // this row is still being published: poll it again
#pragma unroll
					for (int j = 0; j < DPL; ++j)
						if ((open >> j) & 1u) {
							excl[j] += (uint32_t)LB::value(st[w][j]);
							if (LB::code(st[w][j]) == PRE) open &= ~(1u << j);
						}
					++consumed;
					walking = __any_sync(FULL_MASK, open != 0u) != 0;
				}
				if (consumed == 0) lb_backoff();
				p -= consumed; // never passes tile 0: its row holds prefixes, which close every digit
			}
			StateT mine[DPL];
#pragma unroll
			for (int j = 0; j < DPL; ++j) mine[j] = LB::pack(PRE, (uint64_t)excl[j] + s_total[lane * DPL + j]);
#ifdef SVO_EMU
			if (!g_emu_lookback_aggregate_only)
#endif
			{
				StateT *row = state + (size_t)tile * RADIX + lane * DPL;
#pragma unroll
				for (int j = 0; j < DPL; j += VPS) store_states16(row + j, &mine[j]);
			}
		}
#pragma unroll
		for (int j = 0; j < DPL; ++j) {
			const uint32_t d = lane * DPL + j;
			s_gofs[d] = g_bins[d] + excl[j] - s_tile_off[d];
		}
	}
#else
	// decoupled look-back, one thread per digit, 4 predecessor states in flight
	if (threadIdx.x < DTHREADS) {
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			const uint32_t d = threadIdx.x * DPT + j;
			uint32_t excl = 0; // offsets are taken mod 2^32 (n < 2^32)
			if (tile != 0 && !(SVO_OS_EXPERIMENT & 2)) {
				const volatile StateT *p = state + (size_t)(tile - 1) * RADIX + d; // nearest predecessor not yet summed
				uint32_t left = tile;                                            // predecessors from p backwards
				for (;;) {
					if (left >= 4u) {
						const StateT s0 = p[0], s1 = p[-RADIX], s2 = p[-2 * RADIX], s3 = p[-3 * RADIX];
						const uint32_t c0 = LB::code(s0), c1 = LB::code(s1), c2 = LB::code(s2), c3 = LB::code(s3);
						if (c0 != PRE && c0 != AGG) { lb_backoff(); continue; }
						excl += (uint32_t)LB::value(s0);
						if (c0 == PRE) break;
						if (c1 != PRE && c1 != AGG) { p -= RADIX, left -= 1u; continue; }
						excl += (uint32_t)LB::value(s1);
						if (c1 == PRE) break;
						if (c2 != PRE && c2 != AGG) { p -= 2 * RADIX, left -= 2u; continue; }
						excl += (uint32_t)LB::value(s2);
						if (c2 == PRE) break;
						if (c3 != PRE && c3 != AGG) { p -= 3 * RADIX, left -= 3u; continue; }
						excl += (uint32_t)LB::value(s3);
						if (c3 == PRE) break;
						p -= 4 * RADIX, left -= 4u;
					} else { // the first few tiles: one state at a time (tile 0 always holds a prefix)
						const StateT s0 = p[0];
						const uint32_t c0 = LB::code(s0);
						if (c0 != PRE && c0 != AGG) { lb_backoff(); continue; }
						excl += (uint32_t)LB::value(s0);
						if (c0 == PRE) break;
						p -= RADIX, left -= 1u;
					}
				}
#ifdef SVO_EMU
				if (!g_emu_lookback_aggregate_only)
#endif
					*reinterpret_cast<volatile StateT *>(state + (size_t)tile * RADIX + d) = LB::pack(PRE, (uint64_t)excl + bin_total[j]);
			}
			s_gofs[d] = g_bins[d] + excl - s_tile_off[d];
		}
	}
#endif
	__syncthreads();

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
}

// Persistent blocks: each draws tiles from the ticket until none is left.  With SVO_OS_TMA the ticket for the NEXT tile
// is drawn, and its bulk copy issued, as soon as the current tile's keys have left the landing buffer -- the copy
// then has the whole ranking / look-back / scatter of the current tile to complete.  (Forward progress: the smallest
// unfinished tile is always either being processed, or held by a block whose current tile is smaller and finished.)
template <int BLOCK, int ITEMS, int RBITS, int MINB, class StateT>
__global__ void __launch_bounds__(BLOCK, MINB) k_onesweep_pass(PassArgs pa) {
	using C = OnesweepCfg<BLOCK, ITEMS, RBITS, StateT>;
	SVO_DYN_SMEM(unsigned char, smem);
	__shared__ uint32_t s_wsum[C::RADIX / 32 + 1];
	__shared__ uint32_t s_next;
	__shared__ uint64_t s_mbar;
	const uint32_t mode = pa.mode ? *pa.mode : 0u;
	if (!((pa.run_mask >> mode) & 1u)) return;
	const uint32_t pass = mode ? pa.pass_in_mode[1] : pa.pass_in_mode[0];
	// (kernel parameters are copied into scalars: handing the struct to the tile function by reference put it on the stack)
	const uint64_t *const keys_in = pa.in;
	uint64_t *const keys_out = pa.out;
	const uint32_t shift = pa.shift, mask = pa.mask;
	const uint32_t *const g_bins = pa.bins;
	StateT *const state = reinterpret_cast<StateT *>(pa.state);
	uint32_t *const ticket = pa.ticket;
	const uint32_t tiles = pa.tiles;
	const uint32_t last_count = (uint32_t)(pa.n - (uint64_t)(tiles - 1) * C::TILE); // keys of the last tile (1..TILE)
	const bool bulk = SVO_OS_TMA && pa.use_bulk;
	uint32_t *const p_next = &s_next;
	uint64_t *const p_mbar = &s_mbar;
	auto tile_is_bulk = [=](uint32_t t) { return bulk && t < tiles && (t + 1 < tiles || last_count == (uint32_t)C::TILE); };
	auto take_next = [=]() { // thread 0: draw the next tile and start its copy into the (free) landing buffer
		if (threadIdx.x == 0) {
			const uint32_t t = atomicAdd(ticket, 1u);
			*p_next = t;
			if (tile_is_bulk(t)) bulk_load(smem + C::OFF_STAGE, keys_in + (uint64_t)t * C::TILE, (uint32_t)C::TILE * 8u, p_mbar);
		}
	};
	if (SVO_OS_TMA && threadIdx.x == 0) mbar_init(&s_mbar, 1);
	take_next();
	__syncthreads();
	uint32_t tile = s_next;
	uint32_t phase = 0; // parity of the mbarrier: one completed copy per bulk-loaded tile
	while (tile < tiles) {
		const bool staged = tile_is_bulk(tile);
		const uint32_t count = tile + 1 < tiles ? (uint32_t)C::TILE : last_count;
		if (staged) {
			mbar_wait(&s_mbar, phase);
			phase ^= 1u;
		}
		if (count == (uint32_t)C::TILE)
			onesweep_tile<BLOCK, ITEMS, RBITS, StateT, true>(keys_in, keys_out, shift, mask, g_bins, state, pass, tile, count, staged, smem, s_wsum,
			                                                 take_next);
		else
			onesweep_tile<BLOCK, ITEMS, RBITS, StateT, false>(keys_in, keys_out, shift, mask, g_bins, state, pass, tile, count, false, smem, s_wsum,
			                                                  take_next);
		__syncthreads(); // the scatter has read the reorder buffer and s_gofs; s_next is visible
		tile = s_next;
	}
}

inline bool g_force_wide_sort_state = false; // svo_debug_force_wide_sort_state (tests)
inline bool g_profile_passes = false;        // svo_debug_profile_passes: an event after every pass

constexpr int SORT_MAX_EVENTS = 2 * MAX_PASSES + 4;
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
#define SVO_OS_MINB 2
#endif
constexpr int OS_BLOCK = SVO_OS_BLOCK, OS_ITEMS = SVO_OS_ITEMS, OS_MINB = SVO_OS_MINB, OS_TILE = OS_BLOCK * OS_ITEMS;

// Persistent grid: as many blocks as fit on the device at once (per device: set the dynamic shared-memory limit and ask
// the occupancy calculator once per kernel and device).
template <int RBITS, class StateT> inline int launch_onesweep_pass(const PassArgs &pa, int device, int n_sm, cudaStream_t s) {
	using C = OnesweepCfg<OS_BLOCK, OS_ITEMS, RBITS, StateT>;
	auto k = k_onesweep_pass<OS_BLOCK, OS_ITEMS, RBITS, OS_MINB, StateT>;
	uint32_t grid;
#ifndef SVO_EMU
	static int per_sm[64] = {};
	const int di = device >= 0 && device < 64 ? device : 0;
	if (!per_sm[di]) {
		SVO_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
		int nb = 0;
		SVO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, OS_BLOCK, C::SMEM));
		if (nb < 1) {
			set_error("k_onesweep_pass does not fit on an SM (%zu bytes of shared memory)", (size_t)C::SMEM);
			return -2;
		}
		per_sm[di] = nb;
	}
	grid = (uint32_t)per_sm[di] * (uint32_t)(n_sm > 0 ? n_sm : 148);
#else
	(void)device, (void)n_sm;
	grid = 3;
#endif
	if (grid > pa.tiles) grid = pa.tiles;
	SVO_LAUNCH(grid, OS_BLOCK, C::SMEM, s, k, pa);
	return 0;
}
inline int launch_onesweep(bool nine, bool wide, const PassArgs &pa, int device, int n_sm, cudaStream_t s) {
	if (nine) return wide ? launch_onesweep_pass<9, uint64_t>(pa, device, n_sm, s) : launch_onesweep_pass<9, uint32_t>(pa, device, n_sm, s);
	return wide ? launch_onesweep_pass<8, uint64_t>(pa, device, n_sm, s) : launch_onesweep_pass<8, uint32_t>(pa, device, n_sm, s);
}

inline int launch_radix_histogram(const uint64_t *keys, uint64_t n, const SortPasses &sp, uint32_t *hist, int n_sm, cudaStream_t s) {
	uint32_t hgrid = div_up(n, (uint64_t)HIST_BLOCK * HIST_ITEMS);
	const uint32_t hmax = (uint32_t)(n_sm > 0 ? n_sm : 148) * (uint32_t)SVO_HIST_GRID;
	if (hgrid > hmax) hgrid = hmax;
	const bool narrow = sp.shift[sp.n_pass - 1] - sp.shift[0] < 32u;
	switch (sp.n_pass) {
#define SVO_HIST_CASE(NP)                                                                              \
	case NP: {                                                                                         \
		if (narrow) {                                                                                  \
			SVO_LAUNCH(hgrid, HIST_BLOCK, 0, s, (k_radix_histogram<NP, true>), keys, n, sp, hist);     \
		} else {                                                                                       \
			SVO_LAUNCH(hgrid, HIST_BLOCK, 0, s, (k_radix_histogram<NP, false>), keys, n, sp, hist);    \
		}                                                                                              \
		break;                                                                                         \
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
	sc.n_ev = 0;
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
		pa.in = src, pa.out = dst, pa.n = n, pa.tiles = tiles;
		pa.shift = sp.shift[p], pa.mask = sp.mask[p];
		pa.bins = sc.hist.p + p * MAX_RADIX;
		pa.state = sc.state.p;
		pa.ticket = sc.hist.p + SORT_HIST_WORDS + p;
		pa.mode = nullptr, pa.run_mask = 1u, pa.pass_in_mode[0] = pa.pass_in_mode[1] = p;
		pa.use_bulk = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 ? 1u : 0u;
		SVO_TRY(launch_onesweep(nine, wide, pa, device, n_sm, s));
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
