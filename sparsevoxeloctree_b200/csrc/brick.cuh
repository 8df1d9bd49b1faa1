// brick.cuh -- the binned path for large triangles: no fragment list, no fragment sort.
//
// Wall- and floor-sized triangles make most of the fragments of an interior scene (config C4: 97 %), and every one of
// those fragments used to cross HBM nine times on its way into the tree (emit, histogram, 4 x sort read + write,
// reduce).  Here the large triangles are BINNED instead: the grid is cut into bricks of 8^3 voxels, every large
// triangle lists the bricks its plane can touch ((brick, triangle) pairs: k_brick_pairs), the pairs -- 1/64 of the
// fragments -- are sorted by brick Morton code with the same onesweep sort, and one warp per brick then rasterizes the
// brick's triangles, in triangle order, straight into a dense 512-cell grid in shared memory (k_brick_build): coverage
// by the same exact integer edge functions, depth by the same fp64 plane, colours folded by the same running average as
// on the fragment path, so the leaves are bit-identical to sorting the emitted fragments.  The cell index is the
// Morton code inside the brick, so reading the grid in order IS the sorted, de-duplicated leaf list; the three deepest
// levels (leaves, depth L-1, depth L-2) come out of the same pass in the layout k_reduce_fused produces, and the
// upper levels / node emission run unchanged.
//
// Small triangles (candidate rectangle <= 256 pixels) keep the fragment path -- emit, sort, reduce -- on their own
// (short) list; their leaves enter the bricks as a "small record" that sorts in front of the brick's triangles, which
// is exactly the fragment order of the list (small class first, then large triangles by id): the fold continues from
// the small fragments' leaf word.
//
// Replaces, for these triangles: voxelizer.frag:18-44 (fragment append) + octree_tag_node.comp:18-60 (L descents per
// fragment).  north_star: "a binned tile pass for large triangles".
#pragma once
#include "build.cuh"
#include "raster.cuh"

namespace svo {

constexpr int BRICK_LOG = 3;                // 8^3 voxels: three tree levels inside a brick
constexpr int BRICK_CELLS = 512;
constexpr int BRICK_WARPS = 8, BRICK_BLOCK = BRICK_WARPS * 32; // one warp per brick

// A pair word: brick Morton code << 33 | is_large << 32 | payload.  payload = index of the large triangle, or (small
// record, is_large = 0: sorts first inside its brick) the index of the brick's first leaf in the small-leaf list.
SVO_HD inline uint64_t pair_large(uint64_t brick, uint32_t li) { return (brick << 33) | (1ull << 32) | (uint64_t)li; }
SVO_HD inline uint64_t pair_small(uint64_t brick, uint32_t first) { return (brick << 33) | (uint64_t)first; }
constexpr uint32_t PAIR_SORT_BEGIN = 32; // sorted bits: [32, 33 + 3 * (level - 3))

SVO_HD inline void screen_axes(uint32_t axis, uint32_t &wx, uint32_t &wy) { // world axes of screen x / y (voxelizer.frag:24)
	wx = axis == 0u ? 1u : (axis == 1u ? 2u : 0u);
	wy = axis == 0u ? 2u : (axis == 1u ? 0u : 1u);
}

// rows of 8x8-pixel tiles a large triangle's pixel rectangle spans (tiles are aligned to the shard-local grid)
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_brick_tile_rows(RasterParams rp, uint32_t n_large, const LargeTri *__restrict__ large, uint32_t *__restrict__ n_tr) {
	const uint32_t li = blockIdx.x * RASTER_BLOCK + threadIdx.x;
	if (li >= n_large) return;
	const TriSetup &ts = large[li].ts;
	uint32_t wx, wy;
	screen_axes(ts.axis, wx, wy);
	const int32_t oy = (int32_t)pick3(rp.origin, wy);
	n_tr[li] = (uint32_t)(((ts.py1 - oy) >> BRICK_LOG) - ((ts.py0 - oy) >> BRICK_LOG) + 1);
}

// One warp per large triangle (a triangle's tile rows are dealt out over gridDim.y warps).  Per tile row: the exact
// spans of its 8 pixel rows give the tiles it touches; per tile the depth plane at the corners of the touched
// rectangle bounds the depth bricks (depth is monotone in x and in y, so the corners hold the extremes).  The list is a
// superset -- a brick whose pixels all miss the triangle simply receives nothing.
//   EMIT = false: pair_cnt[tr_base[li] + row] = pairs of the tile row
//   EMIT = true : the pairs, written at pair_off[tr_base[li] + row] in (tile, depth brick) order.  The list is in
//                 triangle order, so the stable sort by brick leaves every brick's triangles in triangle order.
template <bool EMIT>
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_brick_pairs(RasterParams rp, uint32_t n_large, const LargeTri *__restrict__ large, const uint64_t *__restrict__ tr_base,
                  uint32_t *__restrict__ pair_cnt, const uint64_t *__restrict__ pair_off, uint64_t *__restrict__ pairs) {
	const uint32_t li = (blockIdx.x * RASTER_BLOCK + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (li >= n_large) return; // whole warp leaves together
	const TriSetup &ts = large[li].ts;
	const uint32_t axis = ts.axis;
	uint32_t wx, wy;
	screen_axes(axis, wx, wy);
	const int32_t ox = (int32_t)pick3(rp.origin, wx), oy = (int32_t)pick3(rp.origin, wy);
	const uint32_t oz = pick3(rp.origin, axis);
	const int32_t t0 = (ts.py0 - oy) >> BRICK_LOG;
	const int32_t ntr = ((ts.py1 - oy) >> BRICK_LOG) - t0 + 1;
	const uint64_t base = tr_base[li];
	for (int32_t tr = (int32_t)blockIdx.y; tr < ntr; tr += (int32_t)gridDim.y) {
		const int32_t ty = t0 + tr;
		int32_t xlo = 0x7fffffff, xhi = -0x7fffffff;
		if (lane < 8) {
			const int32_t py = oy + ty * 8 + lane;
			if (py >= ts.py0 && py <= ts.py1) {
				int32_t a, b;
				row_span(ts, py, a, b);
				row_span_depth_window(ts, rp.res, py, a, b);
				if (a <= b) xlo = a, xhi = b;
			}
		}
#pragma unroll
		for (int d = 4; d > 0; d >>= 1) {
			xlo = tmin(xlo, __shfl_xor_sync(FULL_MASK, xlo, d));
			xhi = tmax(xhi, __shfl_xor_sync(FULL_MASK, xhi, d));
		}
		xlo = __shfl_sync(FULL_MASK, xlo, 0), xhi = __shfl_sync(FULL_MASK, xhi, 0);
		uint32_t total = 0;
		if (xlo <= xhi) { // warp-uniform
			const int32_t tx0 = (xlo - ox) >> BRICK_LOG, tx1 = (xhi - ox) >> BRICK_LOG;
			const int32_t ylo = tmax(ts.py0, oy + ty * 8), yhi = tmin(ts.py1, oy + ty * 8 + 7);
			const double r_lo = depth_row_term(ts, ylo), r_hi = depth_row_term(ts, yhi);
			uint64_t o = EMIT ? pair_off[base + tr] : 0;
			for (int32_t tb = tx0; tb <= tx1; tb += 32) {
				const int32_t tx = tb + lane;
				uint32_t cnt = 0, zb0 = 0;
				if (tx <= tx1) {
					const int32_t ax = tmax(xlo, ox + tx * 8), bx = tmin(xhi, ox + tx * 8 + 7);
					const uint32_t u0 = pixel_depth_row(ts, rp.res, ax, r_lo), u1 = pixel_depth_row(ts, rp.res, bx, r_lo);
					const uint32_t u2 = pixel_depth_row(ts, rp.res, ax, r_hi), u3 = pixel_depth_row(ts, rp.res, bx, r_hi);
					uint32_t umin = tmin(tmin(u0, u1), tmin(u2, u3)), umax = tmax(tmax(u0, u1), tmax(u2, u3));
					bool any = true;
					if (ts.cull_depth) { // the shard's depth window (fragments outside are dropped)
						umin = tmax(umin, ts.zs_lo), umax = tmin(umax, ts.zs_hi - 1u);
						any = umin <= umax;
					}
					if (any) {
						zb0 = (umin - oz) >> BRICK_LOG;
						cnt = ((umax - oz) >> BRICK_LOG) - zb0 + 1u;
					}
				}
				const uint32_t inc = warp_inclusive_sum(cnt, lane);
				if (EMIT) {
					uint64_t w = o + total + inc - cnt;
					for (uint32_t q = 0; q < cnt; ++q) {
						uint32_t bx, by, bz;
						unswizzle(axis, (uint32_t)tx, (uint32_t)ty, zb0 + q, bx, by, bz);
						pairs[w++] = pair_large(morton3(bx, by, bz), li);
					}
				}
				total += __shfl_sync(FULL_MASK, inc, 31);
			}
		}
		if (!EMIT && lane == 0) pair_cnt[base + tr] = total;
	}
}

// Small records: one per brick that holds leaves of small triangles.  keys = the sorted unique Morton codes of those
// leaves (n = *n_ptr of them).  The records are appended in any order (one per brick: the pair sort orders them).
__global__ void __launch_bounds__(256)
    k_brick_small_records(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ n_ptr, uint64_t *__restrict__ records,
                          unsigned long long *__restrict__ n_records) {
	const uint64_t n = *n_ptr;
	const int lane = threadIdx.x & 31;
	for (uint64_t i0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) & ~31ull; i0 < n; i0 += (uint64_t)gridDim.x * 256) {
		const uint64_t i = i0 + lane;
		bool head = false;
		uint64_t brick = 0;
		if (i < n) {
			brick = keys[i] >> (3 * BRICK_LOG);
			head = i == 0 || (keys[i - 1] >> (3 * BRICK_LOG)) != brick;
		}
		const unsigned b = __ballot_sync(FULL_MASK, head);
		uint64_t pos = 0;
		if (lane == 0 && b) pos = atomicAdd(n_records, (unsigned long long)__popc(b));
		pos = __shfl_sync(FULL_MASK, pos, 0);
		if (head) records[pos + __popc(b & ((1u << lane) - 1u))] = pair_small(brick, (uint32_t)i);
	}
}

// first pair of every brick in the sorted pair list: flags, (scan by exclusive_scan<.., true>), scatter
__global__ void __launch_bounds__(256) k_brick_head_flags(const uint64_t *__restrict__ pairs, uint64_t n, uint32_t *__restrict__ flags) {
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	flags[i] = (i == 0 || (pairs[i] >> 33) != (pairs[i - 1] >> 33)) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
    k_brick_head_scatter(const uint64_t *__restrict__ pairs, const uint32_t *__restrict__ flags, const uint64_t *__restrict__ idx, uint64_t n,
                         uint32_t *__restrict__ brick_first, uint64_t *__restrict__ brick_code) {
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	if (flags[i]) brick_first[idx[i]] = (uint32_t)i, brick_code[idx[i]] = pairs[i] >> 33;
	if (i == n - 1) brick_first[idx[n]] = (uint32_t)n; // idx[n] = number of bricks
}

struct BrickArgs {
	const uint64_t *pairs;       // sorted by brick (stable)
	const uint32_t *brick_first; // [n_bricks + 1] first pair of every brick
	const uint64_t *brick_code;  // [n_bricks] Morton code of every brick
	const uint64_t *n_bricks;    // device scalar
	const LargeTri *large;
	const UvMap *luv;
	TexView tv;
	RasterParams rp;
	const uint64_t *small_keys; // leaves of the small triangles: sorted unique Morton codes, leaf words, how many
	const uint32_t *small_leaf;
	const uint64_t *n_small;
	FusedOut out;          // the three deepest levels, as k_reduce_fused<3> leaves them
	uint64_t *state;       // 3 look-back chains (leaves, depth L-1, depth L-2 nodes) of state_stride words each, zeroed
	uint64_t state_stride;
	uint32_t *ticket;      // zeroed
};

SVO_DEV uint32_t spread3(uint32_t v) { return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4); } // 3 bits -> every third bit
// bit i of the result = byte i of b is not zero
SVO_DEV uint32_t nonzero_bytes4(uint32_t b) {
	uint32_t t = b | (b >> 4);
	t |= t >> 2;
	t |= t >> 1;
	return ((t & 0x01010101u) * 0x01020408u) >> 24;
}

// One warp rasterizes BRICK_BPW consecutive bricks, each into its own grid; a block = BRICK_TILE consecutive bricks = one
// element ("tile", numbered by a ticket) of the three output scans.  The scans' look-back runs once per tile, between
// the rasterization and the write-out (measured with one brick per warp: 39 % of the warp time was spent at that
// barrier; several bricks per warp amortise it).
#ifndef SVO_BRICK_BPW
#define SVO_BRICK_BPW 4
#endif
constexpr int BRICK_BPW = SVO_BRICK_BPW, BRICK_TILE = BRICK_WARPS * BRICK_BPW;
constexpr size_t BRICK_SMEM = (size_t)BRICK_TILE * BRICK_CELLS * 4;
constexpr int LT_WORDS = (int)(sizeof(LargeTri) / 8);
static_assert(sizeof(LargeTri) % 8 == 0 && LT_WORDS <= 32 && BRICK_BPW < 32, "a LargeTri is staged by one warp, 8 bytes per lane");

template <bool TEX> __global__ void __launch_bounds__(BRICK_BLOCK) k_brick_build(BrickArgs a) {
	SVO_DYN_SMEM(uint32_t, s_grid);                        // [BRICK_TILE][512] leaf words, 0 = empty
	__shared__ uint32_t s_bal[BRICK_TILE][BRICK_CELLS / 32]; // occupancy of every grid, 32 cells per word
	__shared__ uint32_t s_cnt[3][BRICK_WARPS];
	__shared__ uint64_t s_base[3][BRICK_WARPS];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_tri[BRICK_WARPS][BRICK_BPW][LT_WORDS]; // the triangle being rasterized, per warp and brick slot
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint64_t nb = *a.n_bricks;
	const uint32_t n_tiles = (uint32_t)((nb + BRICK_TILE - 1) / BRICK_TILE);
	const uint32_t tile = take_ticket(a.ticket, &s_ticket);
	if (tile >= n_tiles) return; // (the grid is sized from the number of pairs, an upper bound)
	const uint32_t sx3 = spread3((uint32_t)lane & 7u), sy3a = spread3((uint32_t)lane >> 3), sy3b = spread3(((uint32_t)lane >> 3) + 4u);

	// The metadata of the warp's bricks is fetched up front, lane q for brick q -- first pair index, Morton code, first
	// pair -- and the first triangle of every brick is staged in shared memory by 22 lanes at once: three dependent
	// round trips per BRICK_BPW bricks instead of four per brick (the kernel is bound by these latencies, not by HBM).
	const uint64_t brick0 = (uint64_t)tile * BRICK_TILE + (uint32_t)(warp * BRICK_BPW);
	uint32_t pf = 0;
	uint64_t code = 0, pr_first = 0;
	if (lane <= BRICK_BPW && brick0 + lane <= nb) pf = a.brick_first[brick0 + lane];
	if (lane < BRICK_BPW && brick0 + lane < nb) {
		code = a.brick_code[brick0 + lane];
		pr_first = a.pairs[pf];
	}
#pragma unroll
	for (int q = 0; q < BRICK_BPW; ++q) {
		const uint64_t pr = __shfl_sync(FULL_MASK, pr_first, q);
		if (brick0 + q < nb && ((pr >> 32) & 1ull) && lane < LT_WORDS)
			s_tri[warp][q][lane] = reinterpret_cast<const uint64_t *>(a.large + (uint32_t)pr)[lane];
	}

	uint32_t c0q[BRICK_BPW]; // leaves per brick
	uint64_t n1q[BRICK_BPW]; // occupancy of the brick's 64 depth L-1 nodes (bit = cell index >> 3)
	uint64_t idq[BRICK_BPW]; // Morton code of the brick
	uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
	for (int q = 0; q < BRICK_BPW; ++q) {
		const uint64_t brick = brick0 + q;
		uint32_t *g = s_grid + (size_t)(warp * BRICK_BPW + q) * BRICK_CELLS;
		{
			uint4 *g4 = reinterpret_cast<uint4 *>(g);
#pragma unroll
			for (int k = 0; k < BRICK_CELLS / 128; ++k) g4[k * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
		}
		__syncwarp();
		c0q[q] = 0, n1q[q] = 0, idq[q] = 0;
		const uint32_t p0 = __shfl_sync(FULL_MASK, pf, q), p1 = __shfl_sync(FULL_MASK, pf, q + 1);
		const uint64_t brick_id = __shfl_sync(FULL_MASK, code, q);
		const uint64_t pr0 = __shfl_sync(FULL_MASK, pr_first, q);
		if (brick >= nb) continue; // warp-uniform
		idq[q] = brick_id;
		const uint32_t bx = compact1by2_10((uint32_t)brick_id), by = compact1by2_10((uint32_t)(brick_id >> 1)),
		               bz = compact1by2_10((uint32_t)(brick_id >> 2));
		for (uint32_t p = p0; p < p1; ++p) {
			const uint64_t pr = p == p0 ? pr0 : a.pairs[p]; // warp-uniform
			if (!((pr >> 32) & 1ull)) {
				// the brick's leaves of small triangles (they come first in fragment order): plain stores, cells are unique
				const uint64_t ns = *a.n_small;
				for (uint64_t i = (uint64_t)(uint32_t)pr + lane;; i += 32) {
					bool ok = i < ns;
					uint64_t key = 0;
					if (ok) key = a.small_keys[i], ok = (key >> (3 * BRICK_LOG)) == brick_id;
					if (ok) g[(uint32_t)key & (BRICK_CELLS - 1)] = a.small_leaf[i];
					if (!__all_sync(FULL_MASK, ok)) break;
				}
			} else {
				// one large triangle: the brick's 8x8 pixels, two per lane; a triangle puts at most one fragment into a
				// voxel, so the read-modify-write of a cell needs no atomics, and triangles follow each other in order
				const uint32_t li = (uint32_t)pr;
				if (p != p0) { // (the first pair's triangle is staged already)
					if (lane < LT_WORDS) s_tri[warp][q][lane] = reinterpret_cast<const uint64_t *>(a.large + li)[lane];
					__syncwarp();
				}
				const LargeTri &lt = *reinterpret_cast<const LargeTri *>(s_tri[warp][q]);
				const TriSetup &ts = lt.ts;
				const uint32_t axis = ts.axis;
				uint32_t wx, wy;
				screen_axes(axis, wx, wy);
				const uint32_t tcx = wx == 0u ? bx : (wx == 1u ? by : bz), tcy = wy == 0u ? bx : (wy == 1u ? by : bz);
				const uint32_t tcz = axis == 0u ? bx : (axis == 1u ? by : bz);
				const int32_t px = (int32_t)(pick3(a.rp.origin, wx) + tcx * 8u) + (lane & 7);
				const int32_t py0 = (int32_t)(pick3(a.rp.origin, wy) + tcy * 8u) + (lane >> 3);
				const uint32_t z_base = pick3(a.rp.origin, axis) + tcz * 8u;
				const uint32_t tex = TEX ? lt.textured : 0u;
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const int32_t py = py0 + 4 * h;
					bool ok = px >= ts.px0 && px <= ts.px1 && py >= ts.py0 && py <= ts.py1 && pixel_covered(ts, px, py);
					uint32_t uz = 0, rgb = lt.rgb;
					ok = ok && pixel_fragment(ts, a.rp.res, px, py, uz);
					const uint32_t dz = uz - z_base;
					ok = ok && dz < 8u;
					if (TEX && ok && (tex & 1u)) {
						const bool pass = sample_colour(a.tv, a.luv[li], px, py, rgb);
						if (tex & 2u) ok = pass; // alpha-tested texture: voxelizer.frag:29-30 discards
					}
					if (ok) {
						const uint32_t cell = (sx3 << wx) | ((h ? sy3b : sy3a) << wy) | (spread3(dz) << axis);
						const uint32_t w = g[cell];
						g[cell] = w ? leaf_accumulate(w, rgb) : leaf_first(rgb);
					}
				}
			}
			__syncwarp();
		}
		// occupancy, 32 cells at a time (cell index = Morton code inside the brick = output order)
		uint32_t c0 = 0;
		uint64_t n1 = 0;
#pragma unroll
		for (int k = 0; k < BRICK_CELLS / 32; ++k) {
			const uint32_t b = __ballot_sync(FULL_MASK, g[k * 32 + lane] != 0u);
			if (lane == 0) s_bal[warp * BRICK_BPW + q][k] = b;
			if (b) { // warp-uniform
				c0 += (uint32_t)__popc(b);
				n1 |= (uint64_t)nonzero_bytes4(b) << (4 * k);
			}
		}
		c0q[q] = c0, n1q[q] = n1;
		t0 += c0, t1 += (uint32_t)__popcll(n1);
		t2 += (uint32_t)__popc(nonzero_bytes4((uint32_t)n1) | (nonzero_bytes4((uint32_t)(n1 >> 32)) << 4));
	}
	if (lane == 0) s_cnt[0][warp] = t0, s_cnt[1][warp] = t1, s_cnt[2][warp] = t2;
	__syncthreads();
	if (warp < 3) { // one look-back chain per output array; lane w also turns the warps' counts into their offsets
		const uint32_t mine = lane < BRICK_WARPS ? s_cnt[warp][lane] : 0u;
		const uint32_t inc = warp_inclusive_sum(mine, lane);
		const uint64_t total = __shfl_sync(FULL_MASK, inc, 31);
		const uint64_t excl = lookback_exclusive(a.state + (uint64_t)warp * a.state_stride, tile, total, lane);
		if (lane < BRICK_WARPS) s_base[warp][lane] = excl + inc - mine;
		if (lane == 0 && tile == n_tiles - 1) *a.out.count[warp] = excl + total;
	}
	__syncthreads();
	if (t0 == 0u) return; // warp-uniform (no barrier follows)
	uint64_t base0 = s_base[0][warp], base1 = s_base[1][warp], base2 = s_base[2][warp];

#pragma unroll
	for (int q = 0; q < BRICK_BPW; ++q) {
		const uint64_t n1 = n1q[q];
		if (n1 == 0ull) continue; // warp-uniform
		const uint32_t *g = s_grid + (size_t)(warp * BRICK_BPW + q) * BRICK_CELLS;
		const uint32_t n2 = nonzero_bytes4((uint32_t)n1) | (nonzero_bytes4((uint32_t)(n1 >> 32)) << 4); // depth L-2 nodes
		uint32_t run0 = 0;
		for (int k = 0; k < BRICK_CELLS / 32; ++k) {
			const uint32_t nm = (uint32_t)(n1 >> (4 * k)) & 0xfu; // the depth L-1 nodes of cells 32k .. 32k+31
			if (!nm) continue;                                    // warp-uniform
			const uint32_t b = s_bal[warp * BRICK_BPW + q][k];
			const uint32_t run1 = (uint32_t)__popcll(n1 & ((1ull << (4 * k)) - 1ull));
			if ((b >> lane) & 1u) {
				const uint64_t u = base0 + run0 + (uint32_t)__popc(b & lt_mask);
				a.out.leaf[u] = g[k * 32 + lane];
				a.out.slot0[u] = (unsigned char)(lane & 7);
			}
			if (lane < 4 && ((nm >> lane) & 1u)) {
				const uint64_t u1 = base1 + run1 + (uint32_t)__popc(nm & lt_mask);
				a.out.first1[u1] = (uint32_t)(base0 + run0 + (uint32_t)__popc(b & ((1u << (8 * lane)) - 1u)));
				a.out.slot1[u1] = (unsigned char)((k * 4 + lane) & 7);
			}
			// the first occupied word of a depth L-2 node (64 cells = words 2m, 2m+1) also writes that node
			if (lane == 0 && (!(k & 1) || !((n1 >> (4 * (k - 1))) & 0xfull))) {
				const uint32_t m = (uint32_t)k >> 1;
				const uint64_t u2 = base2 + (uint32_t)__popc(n2 & ((1u << m) - 1u));
				a.out.first2[u2] = (uint32_t)(base1 + run1);
				a.out.keys_top[u2] = (idq[q] << 3) | (uint64_t)m;
			}
			run0 += (uint32_t)__popc(b);
		}
		base0 += c0q[q], base1 += (uint32_t)__popcll(n1), base2 += (uint32_t)__popc(n2);
	}
}

} // namespace svo
