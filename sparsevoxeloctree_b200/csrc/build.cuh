// build.cuh -- the OctreeBuilder kernels: de-duplicate + colour-reduce the sorted fragments, compact the
// levels of unique Morton keys into their parents (bottom-up), then emit the node words.
//
// Replaces the reference level loop (src/OctreeBuilder.cpp:142-210: octree_init_node / octree_tag_node /
// octree_alloc_node / octree_modify_arg, 4L-2 dependent dispatches, F*L fragment re-reads and F*L(L+1)/2
// dependent pointer loads) with streaming passes over sorted keys:
//   k_reduce_count   : runs per tile at the three deepest granularities; a scan over the tiles then gives every
//                      tile its position in the outputs, so that no tile waits for another one in
//   k_reduce_fused   : one pass over the sorted fragments produces the three deepest levels at once:
//                      runs of equal Morton code -> leaves (colours folded with the reference's integer
//                      running average in run = emission order, octree_tag_node.comp:48-57), and the runs of
//                      equal key>>3 / key>>6 -> the depth L-1 / L-2 nodes.
//   k_parent_compact : keys of depth d -> unique parents (key >> 3) of depth d-1 (the small upper levels).
//                      Sizes stay on the device; persistent blocks draw tiles from a ticket.
//   k_emit_octree    : one thread per 8-word child block: 32 B written once, zeros included, so there is no
//                      separate init pass (octree_init_node) and no allocation atomics (octree_alloc_node).
// Per depth d the passes leave: slot[d][i] = child slot (key & 7) of node i, and first[d][j] = index of the
// first depth-d child of depth-(d-1) node j; children of a node are contiguous (sorted order).
// Layout produced = what the reference produces when its per-level allocation happens to run in Morton
// order: root block at word 0, level windows top-down (octree_modify_arg.comp:9-13), child pointer =
// word index of the child block (octree_alloc_node.comp:21).
#pragma once
#include "scan.cuh"
#include "svo_math.cuh"

namespace svo {

constexpr int CMP_BLOCK = 256, CMP_ITEMS = 8, CMP_TILE = CMP_BLOCK * CMP_ITEMS;
static_assert(CMP_ITEMS == 8, "k_parent_compact packs a thread's 8 slot bytes into one 64-bit store");
constexpr int MAX_LEVEL = 16;

// ---- de-duplicate + colour reduce + the deepest K-1 parent levels, fused ---------------------------------
// K = min(3, level) granularities at once:
//   j = 0: runs of equal Morton code (key >> 24) -> leaves (depth L)
//   j = 1: runs of equal key >> 27               -> depth L-1 nodes
//   j = 2: runs of equal key >> 30               -> depth L-2 nodes
// One tile per thread block.  Elements are mapped to threads striped (element = row*BLOCK + tid), so
// shared-memory traffic is conflict free and the in-warp ranks come from ballots.
#ifndef SVO_RF_BLOCK
#define SVO_RF_BLOCK 256
#endif
#ifndef SVO_RF_ITEMS
#define SVO_RF_ITEMS 8
#endif
#ifndef SVO_RF_MINB
#define SVO_RF_MINB 8
#endif
constexpr int RF_BLOCK = SVO_RF_BLOCK, RF_ITEMS = SVO_RF_ITEMS, RF_TILE = RF_BLOCK * RF_ITEMS, RF_NW = RF_BLOCK / 32;
static_assert(RF_TILE <= 65536 && RF_ITEMS * RF_NW <= 512 && (RF_ITEMS * RF_NW) % 32 == 0, "the (row, warp) count matrix is scanned by one warp");
constexpr int RF_CPL = RF_ITEMS * RF_NW / 32; // counts per lane in that scan

struct FusedOut {
	uint32_t *leaf;           // [count0] leaf words
	unsigned char *slot0;     // [count0] child slot of each leaf (null: not wanted)
	uint32_t *first1;         // [count1] first leaf of each depth-(L-1) node
	unsigned char *slot1;     // [count1] child slot of each depth-(L-1) node
	uint32_t *first2;         // [count2] first depth-(L-1) child of each depth-(L-2) node
	uint64_t *keys_top;       // keys (shifted) of granularity K-1, for the remaining levels
	uint64_t *count[3];       // device scalars receiving the number of runs per granularity
};

// A tile's position in the outputs (the number of runs in front of it, per granularity) comes from a count pass:
// k_reduce_count counts the runs of every tile, a scan over the tiles leaves the prefixes in pre01 / pre2.  The keys
// are read twice, but no tile ever waits for another one: a chained scan inside k_reduce_fused (decoupled look-back
// over the tiles' run counts) measured 1.31 ms on C4 against 0.25 + 0.87 ms this way -- its tiles spent a third of
// their time waiting for their neighbours' aggregates.

// runs per tile and granularity: cnt01[tile] = runs of equal key>>24 | runs of equal key>>27 << 32, cnt2[tile] = key>>30
template <int K>
__global__ void __launch_bounds__(RF_BLOCK)
    k_reduce_count(const uint64_t *__restrict__ frags, uint64_t n, uint64_t *__restrict__ cnt01, uint64_t *__restrict__ cnt2) {
	__shared__ uint32_t s_c[3];
	const int lane = threadIdx.x & 31;
	const uint64_t tile_base = (uint64_t)blockIdx.x * RF_TILE;
	if (threadIdx.x < 3) s_c[threadIdx.x] = 0;
	uint64_t key[RF_ITEMS], prev0[RF_ITEMS]; // prev0: the element before lane 0's (only lane 0 loads it)
#pragma unroll
	for (int i = 0; i < RF_ITEMS; ++i) {
		const uint64_t idx = tile_base + i * RF_BLOCK + threadIdx.x;
		key[i] = idx < n ? frags[idx] : frags[n - 1]; // out-of-range slots repeat the last fragment: no run starts there
		prev0[i] = 0;
		if (lane == 0) prev0[i] = idx == 0 ? ~key[i] : (idx - 1 < n ? frags[idx - 1] : frags[n - 1]);
	}
	__syncthreads();
	uint32_t c[3] = {0, 0, 0};
#pragma unroll
	for (int i = 0; i < RF_ITEMS; ++i) {
		const uint64_t up = __shfl_up_sync(FULL_MASK, key[i], 1);
		const uint64_t x = key[i] ^ (lane == 0 ? prev0[i] : up);
#pragma unroll
		for (int j = 0; j < K; ++j) c[j] += (uint32_t)__popc(__ballot_sync(FULL_MASK, (x >> (24 + 3 * j)) != 0));
	}
	if (lane == 0) {
#pragma unroll
		for (int j = 0; j < K; ++j) atomicAdd(&s_c[j], c[j]);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		cnt01[blockIdx.x] = (uint64_t)s_c[0] | ((uint64_t)s_c[1] << 32);
		cnt2[blockIdx.x] = s_c[2];
	}
}

template <int K>
__global__ void __launch_bounds__(RF_BLOCK, SVO_RF_MINB)
    k_reduce_fused(const uint64_t *__restrict__ frags, uint64_t n, FusedOut out, uint32_t tiles, const uint64_t *__restrict__ pre01,
                   const uint64_t *__restrict__ pre2) {
	__shared__ uint64_t s_keys[RF_TILE + 2]; // [0] = the element before the tile, [TILE+1] = the one after
	__shared__ uint32_t s_cnt[3][RF_ITEMS * RF_NW];
	__shared__ uint32_t s_total[3];
	__shared__ uint16_t s_start[RF_TILE]; // (tiles with many fragments per voxel) first element of every leaf run
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint32_t tile = blockIdx.x;
	const uint64_t tile_base = (uint64_t)tile * RF_TILE;
	const uint64_t last = frags[n - 1];

	// Out-of-range slots repeat the last fragment (no run starts there); the element "before" fragment 0
	// differs from it in every bit, so fragment 0 starts a run at every granularity.
#pragma unroll
	for (int i = 0; i < RF_ITEMS; ++i) {
		const uint32_t e = i * RF_BLOCK + threadIdx.x;
		s_keys[e + 1] = tile_base + e < n ? frags[tile_base + e] : last;
	}
	if (threadIdx.x == 0) s_keys[0] = tile_base > 0 ? frags[tile_base - 1] : ~frags[0];
	if (threadIdx.x == 32) s_keys[RF_TILE + 1] = tile_base + RF_TILE < n ? frags[tile_base + RF_TILE] : ~last;
	__syncthreads();

	// run-start flags + in-warp ranks
	uint32_t packed[RF_ITEMS]; // per row: flags (bits 0..2) and the lane's exclusive rank per granularity (5 bits each)
#pragma unroll
	for (int i = 0; i < RF_ITEMS; ++i) {
		const uint32_t e = i * RF_BLOCK + threadIdx.x;
		const uint64_t x = s_keys[e + 1] ^ s_keys[e];
		uint32_t pk = 0;
#pragma unroll
		for (int j = 0; j < K; ++j) {
			const bool f = (x >> (24 + 3 * j)) != 0;
			const unsigned b = __ballot_sync(FULL_MASK, f);
			pk |= (f ? 1u : 0u) << j;
			pk |= (uint32_t)__popc(b & lt_mask) << (3 + 5 * j);
			if (lane == 0) s_cnt[j][i * RF_NW + warp] = (uint32_t)__popc(b);
		}
		packed[i] = pk;
	}
	__syncthreads();

	// warps 0..K-1: exclusive scan of the (row, warp) counts of one granularity
	if (warp < K) {
		const int j = warp;
		uint32_t c[RF_CPL], sum = 0;
#pragma unroll
		for (int q = 0; q < RF_CPL; ++q) sum += (c[q] = s_cnt[j][RF_CPL * lane + q]);
		const uint32_t inc = warp_inclusive_sum(sum, lane);
		uint32_t run = inc - sum;
#pragma unroll
		for (int q = 0; q < RF_CPL; ++q) {
			s_cnt[j][RF_CPL * lane + q] = run;
			run += c[q];
		}
		if (lane == 31) s_total[j] = inc;
	}
	__syncthreads();
	// runs in front of this tile (k_reduce_count + scan)
	uint64_t agg[K], pre[K];
#pragma unroll
	for (int j = 0; j < K; ++j) agg[j] = s_total[j];
	{
		const uint64_t a = pre01[tile];
		pre[0] = a & 0xffffffffull;
		if (K >= 2) pre[K >= 2 ? 1 : 0] = a >> 32;
		if (K >= 3) pre[K >= 3 ? 2 : 0] = pre2[tile];
	}
	if (tile == tiles - 1 && threadIdx.x == 0) {
#pragma unroll
		for (int j = 0; j < K; ++j) *out.count[j] = pre[j] + agg[j];
	}

	// run owners write their node.  The colour fold of a voxel is sequential (the reference's running average is order
	// dependent), one thread per voxel.  When most elements start a run (about one fragment per voxel) the owner folds
	// in place; when voxels hold many fragments only a few lanes own runs, so the runs are first listed densely and
	// then dealt out one per thread -- every lane folds (the San-Miguel-like workload: 3.7 fragments per voxel).
	const uint64_t p0 = pre[0], p1 = K >= 2 ? pre[K >= 2 ? 1 : 0] : 0, p2 = K >= 3 ? pre[K >= 3 ? 2 : 0] : 0;
	const bool listed = s_total[0] * 4u <= (uint32_t)RF_TILE * 3u; // block-uniform
	auto fold_leaf = [&](uint32_t e, uint64_t key) {
		uint32_t acc = leaf_first((uint32_t)(key & 0xffffffu));
		if (((s_keys[e + 2] ^ key) >> 24) == 0) { // the voxel has more fragments: fold them in emission order
			for (uint64_t q = (uint64_t)e + 1; tile_base + q < n; ++q) {
				const uint64_t kk = q < RF_TILE ? s_keys[q + 1] : frags[tile_base + q];
				if ((kk >> 24) != (key >> 24)) break;
				acc = leaf_accumulate(acc, (uint32_t)(kk & 0xffffffu));
			}
		}
		return acc;
	};
#pragma unroll
	for (int i = 0; i < RF_ITEMS; ++i) {
		const uint32_t pk = packed[i];
		if (!(pk & 1u)) continue;
		const uint32_t e = i * RF_BLOCK + threadIdx.x;
		const uint64_t key = s_keys[e + 1];
		const uint32_t l0 = s_cnt[0][i * RF_NW + warp] + ((pk >> 3) & 31u); // index of the run inside the tile
		const uint64_t u0 = p0 + l0;
		if (listed)
			s_start[l0] = (uint16_t)e;
		else {
			out.leaf[u0] = fold_leaf(e, key);
			if (out.slot0) out.slot0[u0] = (unsigned char)((key >> 24) & 7u);
			if (K == 1) out.keys_top[u0] = key >> 24;
		}
		if (K >= 2 && (pk & 2u)) {
			const uint64_t u1 = p1 + s_cnt[1][i * RF_NW + warp] + ((pk >> 8) & 31u);
			out.first1[u1] = (uint32_t)u0;
			out.slot1[u1] = (unsigned char)((key >> 27) & 7u);
			if (K == 2) out.keys_top[u1] = key >> 27;
			if (K >= 3 && (pk & 4u)) {
				const uint64_t u2 = p2 + s_cnt[2][i * RF_NW + warp] + ((pk >> 13) & 31u);
				out.first2[u2] = (uint32_t)u1;
				out.keys_top[u2] = key >> 30;
			}
		}
	}
	if (listed) {
		__syncthreads();
		for (uint32_t r = threadIdx.x; r < s_total[0]; r += RF_BLOCK) {
			const uint32_t e = s_start[r];
			const uint64_t key = s_keys[e + 1];
			out.leaf[p0 + r] = fold_leaf(e, key);
			if (out.slot0) out.slot0[p0 + r] = (unsigned char)((key >> 24) & 7u);
			if (K == 1) out.keys_top[p0 + r] = key >> 24;
		}
	}
}

// ---- one level up (upper levels: small) ----------------------------------------------------------------------
// keys_in: n = *n_ptr sorted unique keys of depth d.  Writes slot_in[i] = key & 7 for every node, and for every
// run of equal key >> 3: keys_out[u] = parent key, first_out[u] = index of the run's first node; *n_out = runs.
__global__ void __launch_bounds__(CMP_BLOCK)
    k_parent_compact(const uint64_t *__restrict__ keys_in, const uint64_t *__restrict__ n_ptr, uint64_t *__restrict__ keys_out,
                     uint32_t *__restrict__ first_out, unsigned char *__restrict__ slot_in, uint64_t *state, uint32_t *ticket,
                     uint64_t *n_out) {
	__shared__ uint64_t s_warp[CMP_BLOCK / 32 + 1];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_prefix;
	const uint64_t n = *n_ptr;
	const uint64_t n_tiles = (n + CMP_TILE - 1) / CMP_TILE;
	if (n == 0) {
		if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = 0;
		return;
	}
	for (;;) {
		const uint32_t tile = take_ticket(ticket, &s_ticket);
		if (tile >= n_tiles) break;
		const uint64_t base = (uint64_t)tile * CMP_TILE + (uint64_t)threadIdx.x * CMP_ITEMS;
		uint64_t k[CMP_ITEMS];
		bool head[CMP_ITEMS];
		uint64_t prev = (base > 0 && base < n) ? keys_in[base - 1] : 0;
		uint32_t cnt = 0;
		const bool whole = base + CMP_ITEMS <= n;
		if (whole) { // 16-byte loads (base is a multiple of 8 keys)
#pragma unroll
			for (int i = 0; i < CMP_ITEMS; i += 2) {
				const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(keys_in + base + i);
				k[i] = v.x, k[i + 1] = v.y;
			}
		}
#pragma unroll
		for (int i = 0; i < CMP_ITEMS; ++i) {
			const uint64_t idx = base + i;
			if (!whole) k[i] = idx < n ? keys_in[idx] : 0;
			head[i] = idx < n && (idx == 0 || (k[i] >> 3) != (prev >> 3));
			prev = k[i];
			cnt += head[i] ? 1u : 0u;
		}
		if (base + CMP_ITEMS <= n) { // the thread's 8 slot bytes in one aligned store (slot arrays are 8-byte aligned)
			uint64_t packed = 0;
#pragma unroll
			for (int i = 0; i < CMP_ITEMS; ++i) packed |= (k[i] & 7ull) << (8 * i);
			*reinterpret_cast<uint64_t *>(slot_in + base) = packed;
		} else {
#pragma unroll
			for (int i = 0; i < CMP_ITEMS; ++i)
				if (base + i < n) slot_in[base + i] = (unsigned char)(k[i] & 7u);
		}
		uint64_t total;
		const uint64_t excl = block_exclusive_sum<CMP_BLOCK, uint64_t>((uint64_t)cnt, total, s_warp);
		if (threadIdx.x < 32) {
			const uint64_t p = lookback_exclusive(state, tile, total, threadIdx.x);
			if (threadIdx.x == 0) {
				s_prefix = p;
				if (tile == n_tiles - 1) *n_out = p + total;
			}
		}
		__syncthreads();
		uint64_t u = s_prefix + excl;
#pragma unroll
		for (int i = 0; i < CMP_ITEMS; ++i) {
			if (!head[i]) continue;
			keys_out[u] = k[i] >> 3;
			first_out[u] = (uint32_t)(base + i);
			++u;
		}
		__syncthreads();
	}
}

// The levels near the root in ONE launch: a level of depth d has at most 8^d nodes, so from depth TAIL_DEPTH upwards a
// single block walks level after level (what k_parent_compact does per level, without ticket, look-back and the launch
// in between: 6.5 us per level on a B200).  Thread t owns TAIL_ITEMS consecutive keys, held in registers; keys ping-pong
// between the two buffers through global memory (a block sees its own stores after __syncthreads).
constexpr uint32_t TAIL_DEPTH = 4; // the tail starts with the keys of this depth (<= 4096)
constexpr int TAIL_BLOCK = 1024, TAIL_ITEMS = 4;
static_assert((1u << (3 * TAIL_DEPTH)) <= (uint32_t)TAIL_BLOCK * TAIL_ITEMS, "one block holds a whole level");
struct TailArgs {
	uint64_t *keys[2];           // keys[0]: in (depth d0 keys), keys[1]: the other buffer
	uint64_t *counts;            // counts[d], d = 0..level: device scalars (counts[d0] is valid on entry)
	uint32_t *first[TAIL_DEPTH + 1];     // [d]: first_out of the step that consumes the depth-d keys
	unsigned char *slot[TAIL_DEPTH + 1]; // [d]: slot_in of that step
	uint32_t d0;
};
__global__ void __launch_bounds__(TAIL_BLOCK) k_parent_tail(TailArgs a) {
	__shared__ uint32_t s_warp[TAIL_BLOCK / 32 + 1];
	uint32_t cur = 0;
#pragma unroll 1
	for (uint32_t d = a.d0; d >= 1; --d, cur ^= 1u) {
		const uint32_t n = (uint32_t)a.counts[d];
		const uint64_t *kin = a.keys[cur];
		uint64_t *kout = a.keys[cur ^ 1u];
		uint32_t *first = d == 1 ? a.first[1] : (d == 2 ? a.first[2] : (d == 3 ? a.first[3] : a.first[4]));
		unsigned char *slot = d == 1 ? a.slot[1] : (d == 2 ? a.slot[2] : (d == 3 ? a.slot[3] : a.slot[4]));
		static_assert(TAIL_DEPTH == 4, "the selects above");
		const uint32_t lo = threadIdx.x * TAIL_ITEMS;
		uint64_t k[TAIL_ITEMS], prev = 0;
		if (lo > 0 && lo < n) prev = kin[lo - 1];
#pragma unroll
		for (int i = 0; i < TAIL_ITEMS; ++i) k[i] = lo + i < n ? kin[lo + i] : 0;
		uint32_t heads = 0, cnt = 0;
#pragma unroll
		for (int i = 0; i < TAIL_ITEMS; ++i) {
			if (lo + i < n) {
				slot[lo + i] = (unsigned char)(k[i] & 7u);
				if (lo + i == 0 || (k[i] >> 3) != (prev >> 3)) heads |= 1u << i, ++cnt;
			}
			prev = k[i];
		}
		uint32_t total;
		uint32_t u = block_exclusive_sum<TAIL_BLOCK, uint32_t>(cnt, total, s_warp);
#pragma unroll
		for (int i = 0; i < TAIL_ITEMS; ++i)
			if ((heads >> i) & 1u) kout[u] = k[i] >> 3, first[u] = lo + i, ++u;
		if (threadIdx.x == 0) a.counts[d - 1] = total;
		__syncthreads(); // kout and counts[d - 1] are read by the whole block in the next round
	}
}

// ---- node words --------------------------------------------------------------------------------------------
struct EmitParams {
	uint32_t level;
	uint64_t total_blocks;
	uint64_t block_base[MAX_LEVEL + 2];       // block_base[d], d = 1..level+1: first 8-word block of the window of depth-d nodes
	uint64_t count[MAX_LEVEL + 1];            // nodes per depth
	const uint32_t *first[MAX_LEVEL + 1];     // [d]: per depth-(d-1) node, index of its first child among the depth-d nodes
	const unsigned char *slot[MAX_LEVEL + 1]; // [d]: per depth-d node, its child slot
	const uint32_t *leaf;                     // leaf words of the depth-`level` nodes
	// placement: block g is written at words[(g - block_shift) * 8]; a child pointer to block c is
	// (c - block_shift) * 8 + ptr_bias.  block_shift > 0 sends the first block_shift blocks (the root block, or the root
	// block and the depth-1 blocks) to root_dst instead, so that a subtree can be emitted straight into a larger (possibly
	// peer-GPU) buffer at word offset ptr_bias and its top blocks merged with those of other subtrees.
	uint32_t block_shift, ptr_bias;
	uint32_t *root_dst;
};

// One thread builds one 8-word block.  STAGED = false: the thread stores its 32 bytes directly (fastest into local
// HBM).  STAGED = true: the warp's 32 blocks (1 KB, contiguous) are transposed through shared memory so that every
// store instruction writes 512 contiguous bytes -- full sectors, which is what counts when `words` is a peer GPU's
// memory and the stores cross NVLink (450 -> 670 GB/s measured).
#ifndef SVO_EMITO_BLOCK
#define SVO_EMITO_BLOCK 128
#endif
constexpr int EMITO_BLOCK = SVO_EMITO_BLOCK;
template <bool STAGED>
__global__ void __launch_bounds__(EMITO_BLOCK) k_emit_octree(EmitParams ep, uint32_t *__restrict__ words) {
	__shared__ uint4 s_stage[STAGED ? EMITO_BLOCK * 2 : 1];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t g = (uint64_t)blockIdx.x * EMITO_BLOCK + threadIdx.x;
	uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	if (g < ep.total_blocks) {
		uint32_t d = ep.level; // the window of depth-d nodes' blocks that holds g: search from the deepest (largest) one
#pragma unroll 1
		while (d > 1u && g < ep.block_base[d]) --d;
		const uint64_t j = g - ep.block_base[d];
		const uint32_t *first = ep.first[d];
		const unsigned char *slot = ep.slot[d];
		const uint32_t c0 = first[j];
		const bool leaf_level = d == ep.level;
		const uint32_t c1 = j + 1 < ep.count[d - 1] ? first[j + 1] : (uint32_t)ep.count[d];
		const uint32_t nc = c1 - c0; // 1..8 children, contiguous, slots ascending
		uint32_t m = 0;
#pragma unroll
		for (int q = 0; q < 8; ++q)
			if ((uint32_t)q < nc) m |= 1u << slot[c0 + q];
		const uint64_t child_base = ep.block_base[d + 1];
		uint32_t c = c0;
#pragma unroll
		for (int sl = 0; sl < 8; ++sl) {
			if ((m >> sl) & 1u) {
				w[sl] = leaf_level ? ep.leaf[c] : (0x80000000u | ((uint32_t)((child_base + c - ep.block_shift) << 3) + ep.ptr_bias));
				++c;
			}
		}
	}
	if (g < ep.block_shift) { // a top block of a subtree that is stitched elsewhere
		uint4 *r = reinterpret_cast<uint4 *>(ep.root_dst + g * 8);
		r[0] = make_uint4(w[0], w[1], w[2], w[3]);
		r[1] = make_uint4(w[4], w[5], w[6], w[7]);
	}
	if (!STAGED) {
		if (g < ep.total_blocks && g >= ep.block_shift) {
			uint4 *o = reinterpret_cast<uint4 *>(words + (g - ep.block_shift) * 8);
			o[0] = make_uint4(w[0], w[1], w[2], w[3]);
			o[1] = make_uint4(w[4], w[5], w[6], w[7]);
		}
		return;
	}
	uint4 *ws = s_stage + warp * 64;
	ws[2 * lane] = make_uint4(w[0], w[1], w[2], w[3]);
	ws[2 * lane + 1] = make_uint4(w[4], w[5], w[6], w[7]);
	__syncwarp();
	// the warp's blocks g0 .. g0+31 occupy 64 consecutive uint4; lane l stores vectors l and 32 + l
	const uint64_t g0 = (uint64_t)blockIdx.x * EMITO_BLOCK + warp * 32;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint64_t blk = g0 + (uint64_t)((h * 32 + lane) >> 1); // block this vector belongs to
		if (blk < ep.total_blocks && blk >= ep.block_shift)
			reinterpret_cast<uint4 *>(words)[(blk - ep.block_shift) * 2 + ((h * 32 + lane) & 1)] = ws[h * 32 + lane];
	}
}

// Multi-GPU stitch: copy a built subtree to dst (possibly peer memory mapped through CUDA IPC), adding
// base_words to every internal child pointer (leaves untouched).  One uint4 per thread.
__global__ void __launch_bounds__(256)
    k_rebase_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint64_t n_vec, uint32_t base_words) {
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n_vec) return;
	uint4 v = src[i];
	v.x = (v.x & 0xC0000000u) == 0x80000000u ? v.x + base_words : v.x;
	v.y = (v.y & 0xC0000000u) == 0x80000000u ? v.y + base_words : v.y;
	v.z = (v.z & 0xC0000000u) == 0x80000000u ? v.z + base_words : v.z;
	v.w = (v.w & 0xC0000000u) == 0x80000000u ? v.w + base_words : v.w;
	dst[i] = v;
}

} // namespace svo
