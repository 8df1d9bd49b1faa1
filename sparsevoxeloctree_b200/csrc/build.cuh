// build.cuh -- the OctreeBuilder kernels: de-duplicate + colour-reduce the sorted fragments, compact one
// level of unique Morton keys into its parents (bottom-up), then emit the node words.
//
// Replaces the reference level loop (src/OctreeBuilder.cpp:142-210: octree_init_node / octree_tag_node /
// octree_alloc_node / octree_modify_arg, 4L-2 dependent dispatches, F*L fragment re-reads and F*L(L+1)/2
// dependent pointer loads) with streaming passes over sorted keys:
//   k_dedup_reduce   : runs of equal Morton code -> one leaf; colours folded with the reference's integer
//                      running average in run (= emission) order (octree_tag_node.comp:48-57).
//   k_parent_compact : keys of depth d -> unique parents (key >> 3) of depth d-1, with each parent's first
//                      child index and 8-bit child mask.  Sizes stay on the device; persistent blocks draw
//                      tiles from a ticket and chain their counts by decoupled look-back.
//   k_emit_octree    : one thread per 8-word child block: 32 B written once, zeros included, so there is no
//                      separate init pass (octree_init_node) and no allocation atomics (octree_alloc_node).
// Layout produced = what the reference produces when its per-level allocation happens to run in Morton
// order: root block at word 0, level windows top-down (octree_modify_arg.comp:9-13), child pointer =
// word index of the child block (octree_alloc_node.comp:21).
#pragma once
#include "scan.cuh"
#include "svo_math.cuh"

namespace svo {

constexpr int CMP_BLOCK = 256, CMP_ITEMS = 8, CMP_TILE = CMP_BLOCK * CMP_ITEMS;
constexpr int MAX_LEVEL = 16;

// ---- de-duplicate + colour reduce ----------------------------------------------------------------------
__global__ void __launch_bounds__(CMP_BLOCK)
    k_dedup_reduce(const uint64_t *__restrict__ frags, uint64_t n, uint64_t *__restrict__ out_keys,
                   uint32_t *__restrict__ out_leaf, uint64_t *state, uint32_t *ticket, uint64_t *count_out) {
	__shared__ uint64_t s_warp[CMP_BLOCK / 32 + 1];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_prefix;
	const uint64_t n_tiles = (n + CMP_TILE - 1) / CMP_TILE;
	if (n == 0) {
		if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = 0;
		return;
	}
	for (;;) {
		const uint32_t tile = take_ticket(ticket, &s_ticket);
		if (tile >= n_tiles) break;
		const uint64_t base = (uint64_t)tile * CMP_TILE + (uint64_t)threadIdx.x * CMP_ITEMS;
		uint64_t k[CMP_ITEMS];
		bool head[CMP_ITEMS];
		uint64_t prev = (base > 0 && base < n) ? frags[base - 1] : 0;
		uint32_t cnt = 0;
#pragma unroll
		for (int i = 0; i < CMP_ITEMS; ++i) {
			const uint64_t idx = base + i;
			k[i] = idx < n ? frags[idx] : 0;
			const bool first = idx == 0;
			head[i] = idx < n && (first || (k[i] >> 24) != (prev >> 24));
			prev = k[i];
			cnt += head[i] ? 1u : 0u;
		}
		uint64_t total;
		const uint64_t excl = block_exclusive_sum<CMP_BLOCK, uint64_t>((uint64_t)cnt, total, s_warp);
		if (threadIdx.x < 32) {
			const uint64_t p = lookback_exclusive(state, tile, total, threadIdx.x);
			if (threadIdx.x == 0) {
				s_prefix = p;
				if (tile == n_tiles - 1) *count_out = p + total;
			}
		}
		__syncthreads();
		uint64_t u = s_prefix + excl;
#pragma unroll
		for (int i = 0; i < CMP_ITEMS; ++i) {
			if (!head[i]) continue;
			const uint64_t key = k[i] >> 24;
			uint32_t acc = leaf_first((uint32_t)(k[i] & 0xffffffu));
			for (uint64_t j = base + i + 1; j < n; ++j) {
				const uint64_t kk = frags[j];
				if ((kk >> 24) != key) break;
				acc = leaf_accumulate(acc, (uint32_t)(kk & 0xffffffu));
			}
			out_keys[u] = key;
			out_leaf[u] = acc;
			++u;
		}
		__syncthreads(); // s_prefix is rewritten by the next tile
	}
}

// ---- one level up ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CMP_BLOCK)
    k_parent_compact(const uint64_t *__restrict__ keys_in, const uint64_t *__restrict__ n_ptr, uint64_t *__restrict__ keys_out,
                     uint32_t *__restrict__ first_out, unsigned char *__restrict__ mask_out, uint64_t *state, uint32_t *ticket,
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
#pragma unroll
		for (int i = 0; i < CMP_ITEMS; ++i) {
			const uint64_t idx = base + i;
			k[i] = idx < n ? keys_in[idx] : 0;
			head[i] = idx < n && (idx == 0 || (k[i] >> 3) != (prev >> 3));
			prev = k[i];
			cnt += head[i] ? 1u : 0u;
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
			const uint64_t parent = k[i] >> 3;
			uint32_t m = 1u << (uint32_t)(k[i] & 7u);
			for (uint64_t j = base + i + 1; j < n; ++j) { // at most 7 more children
				const uint64_t kk = keys_in[j];
				if ((kk >> 3) != parent) break;
				m |= 1u << (uint32_t)(kk & 7u);
			}
			keys_out[u] = parent;
			first_out[u] = (uint32_t)(base + i);
			mask_out[u] = (unsigned char)m;
			++u;
		}
		__syncthreads();
	}
}

// ---- node words --------------------------------------------------------------------------------------------
struct EmitParams {
	uint32_t level;
	uint64_t total_blocks;
	uint64_t block_base[MAX_LEVEL + 2];      // block_base[d], d = 1..level+1: first 8-word block of the window of depth-d nodes
	const uint32_t *first[MAX_LEVEL + 1];    // [d]: per depth-(d-1) node, index of its first child among the depth-d nodes
	const unsigned char *mask[MAX_LEVEL + 1];
	const uint32_t *leaf;                    // leaf words of the depth-`level` nodes
};

__global__ void __launch_bounds__(256) k_emit_octree(EmitParams ep, uint32_t *__restrict__ words) {
	const uint64_t g = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (g >= ep.total_blocks) return;
	uint32_t d = 1;
#pragma unroll 1
	while (d < ep.level && g >= ep.block_base[d + 1]) ++d;
	const uint64_t j = g - ep.block_base[d];
	uint32_t c = ep.first[d][j];
	const uint32_t m = ep.mask[d][j];
	const bool leaf_level = d == ep.level;
	const uint64_t child_base = ep.block_base[d + 1];
	uint32_t w[8];
#pragma unroll
	for (int s = 0; s < 8; ++s) {
		if ((m >> s) & 1u) {
			w[s] = leaf_level ? ep.leaf[c] : (0x80000000u | (uint32_t)((child_base + c) << 3));
			++c;
		} else
			w[s] = 0u;
	}
	uint4 *o = reinterpret_cast<uint4 *>(words + g * 8);
	o[0] = make_uint4(w[0], w[1], w[2], w[3]);
	o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Multi-GPU stitch: copy blocks [1, total) of a built subtree to dst (possibly peer memory), adding
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
