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
constexpr int BRICK_WARPS = 8, BRICK_BLOCK = BRICK_WARPS * 32; // a warp works on one brick at a time

// A pair word: brick Morton code << 33 | is_large << 32 | payload.  payload = index of the large triangle, or (small
// record, is_large = 0: first inside its brick) the index of the brick's first leaf in the small-leaf list.
SVO_HD inline uint64_t pair_large(uint64_t brick, uint32_t li) { return (brick << 33) | (1ull << 32) | (uint64_t)li; }
// Payload of a large pair: triangle index in bits 0..25; bit 31 = FLAT: the triangle covers all 64 pixels of the brick's
// footprint with one and the same depth voxel (bits 28..30: that voxel's position inside the brick; bits 26..27: the
// triangle's dominant axis) and takes its colour from the draw -- the whole content of such a pair is known without
// looking at a single pixel.  Walls and floors parallel to the grid are made of such pairs.
constexpr uint32_t PAIR_LI_MASK = 0x03ffffffu, PAIR_FLAT = 0x80000000u;
// brick record, word w: counts in bits 0..20; bit 21 = the brick is one flat pair (axis: bits 22..23, depth parity: bit 24,
// colour: word z bits 8..31) -- its leaf blocks are not in temp
constexpr uint32_t REC_FLAT = 1u << 21;
SVO_HD inline uint32_t pair_flat_bits(uint32_t axis, uint32_t dz) { return PAIR_FLAT | (dz << 28) | (axis << 26); }
SVO_HD inline uint64_t pair_small(uint64_t brick, uint32_t first) { return (brick << 33) | (uint64_t)first; }
constexpr uint32_t PAIR_SORT_BEGIN = 33; // sorted bits: [33, 33 + 3 * (level - 3)) -- the brick code only: small records are put in front
                                         // of the large pairs before the (stable) sort, so they stay first inside their brick

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
	const bool flat_ok = EMIT && large[li].textured == 0u && li <= PAIR_LI_MASK; // (textured triangles: the colour varies per pixel)
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
				uint32_t cnt = 0, zb0 = 0, flat = 0;
				if (tx <= tx1) {
					const int32_t ax = tmax(xlo, ox + tx * 8), bx = tmin(xhi, ox + tx * 8 + 7);
					const uint32_t u0 = pixel_depth_row(ts, rp.res, ax, r_lo), u1 = pixel_depth_row(ts, rp.res, bx, r_lo);
					const uint32_t u2 = pixel_depth_row(ts, rp.res, ax, r_hi), u3 = pixel_depth_row(ts, rp.res, bx, r_hi);
					if (EMIT && flat_ok && bx - ax == 7 && yhi - ylo == 7 && u0 == u1 && u0 == u2 && u0 == u3) {
						// FLAT: the four corner pixels of the whole 8 x 8 tile are covered and give fragments with the same depth
						// voxel.  The edge functions are linear and the depth voxel is monotone in x and in y, so every pixel of
						// the tile is then covered, passes the depth clip / window like the corners, and lands in that voxel.
						bool all = true;
#pragma unroll
						for (int i = 0; i < 3; ++i) { // pixel_covered at the corners: one evaluation, stepped 7 columns / 7 rows
							const int64_t e = ts.ea[i] * (int64_t)ax + ts.eb[i] * (int64_t)ylo + ts.ec[i];
							const int64_t ex = e + 7 * ts.ea[i], ey = e + 7 * ts.eb[i];
							all = all && (e | ex | ey | (ex + 7 * ts.eb[i])) >= 0;
						}
						if (all && (ts.clip_z || ts.cull_depth)) { // fragments can be dropped: ask pixel_fragment itself
							uint32_t c;
							all = pixel_fragment(ts, rp.res, ax, ylo, c) && pixel_fragment(ts, rp.res, bx, ylo, c) &&
							      pixel_fragment(ts, rp.res, ax, yhi, c) && pixel_fragment(ts, rp.res, bx, yhi, c);
						}
						if (all) flat = pair_flat_bits(axis, (u0 - oz) & 7u); // (u0 = the corners' depth voxel: pixel_depth_row is pixel_fragment's arithmetic)
					}
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
						pairs[w++] = pair_large(morton3(bx, by, bz), li | flat); // (a flat tile has one depth brick: cnt = 1)
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
	__shared__ uint32_t s_warp[8];
	__shared__ unsigned long long s_base;
	const uint64_t n = *n_ptr;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint64_t i0 = (uint64_t)blockIdx.x * 256; i0 < n; i0 += (uint64_t)gridDim.x * 256) { // block-uniform
		const uint64_t i = i0 + threadIdx.x;
		bool head = false;
		uint64_t brick = 0;
		if (i < n) {
			brick = keys[i] >> (3 * BRICK_LOG);
			head = i == 0 || (keys[i - 1] >> (3 * BRICK_LOG)) != brick;
		}
		const unsigned b = __ballot_sync(FULL_MASK, head);
		if (lane == 0) s_warp[warp] = (uint32_t)__popc(b);
		__syncthreads();
		if (threadIdx.x == 0) { // one global atomic per 256 keys
			uint32_t t = 0;
			for (int w = 0; w < 8; ++w) {
				const uint32_t c = s_warp[w];
				s_warp[w] = t;
				t += c;
			}
			s_base = t ? atomicAdd(n_records, (unsigned long long)t) : 0ull;
		}
		__syncthreads();
		if (head) records[s_base + s_warp[warp] + (uint32_t)__popc(b & ((1u << lane) - 1u))] = pair_small(brick, (uint32_t)i);
		__syncthreads();
	}
}

// The brick table from the sorted pair list, in one pass (chained scan over the tiles, numbered by a ticket): entry u =
// first pair of brick u and its code -- Morton code in bits 0..29, its coordinates (10 bits each) from bit 32 --,
// brick_first[n_bricks] = n, *n_bricks.  state: tiles + 1 words, zeroed; ticket: zeroed.
__global__ void __launch_bounds__(SCAN_BLOCK)
    k_brick_heads(const uint64_t *__restrict__ pairs, uint64_t n, uint32_t *__restrict__ brick_first, uint64_t *__restrict__ brick_code,
                  uint64_t *__restrict__ n_bricks, uint64_t *state, uint32_t *ticket) {
	__shared__ uint64_t s_warp[SCAN_BLOCK / 32 + 1];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_prefix;
	const uint32_t tile = take_ticket(ticket, &s_ticket);
	const uint64_t base = (uint64_t)tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint64_t code[SCAN_ITEMS];
	uint64_t prev = (base > 0 && base <= n) ? pairs[base - 1] >> 33 : ~0ull;
	uint32_t heads = 0, cnt = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		code[i] = base + i < n ? pairs[base + i] >> 33 : prev;
		if (base + i < n && code[i] != prev) heads |= 1u << i, ++cnt;
		prev = code[i];
	}
	uint64_t total;
	const uint64_t excl = block_exclusive_sum<SCAN_BLOCK, uint64_t>((uint64_t)cnt, total, s_warp);
	if (threadIdx.x < 32) {
		const uint64_t p = lookback_exclusive(state, tile, total, threadIdx.x);
		if (threadIdx.x == 0) s_prefix = p;
	}
	__syncthreads();
	uint64_t u = s_prefix + excl;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		if ((heads >> i) & 1u) {
			const uint64_t m = code[i];
			brick_first[u] = (uint32_t)(base + i);
			brick_code[u] = m | ((uint64_t)compact1by2_10((uint32_t)m) << 32) | ((uint64_t)compact1by2_10((uint32_t)(m >> 1)) << 42) |
			                ((uint64_t)compact1by2_10((uint32_t)(m >> 2)) << 52);
			++u;
		}
	}
	if (n > 0 && base <= n - 1 && n - 1 < base + SCAN_ITEMS) brick_first[u] = (uint32_t)n, *n_bricks = u; // the thread that owns the last pair
}

struct BrickArgs {
	const uint64_t *pairs;       // sorted by brick (stable)
	const uint32_t *brick_first; // [n_bricks + 1] first pair of every brick
	const uint64_t *brick_code;  // [n_bricks] Morton code of every brick (bits 0..29) and its x, y, z (bits 32.., 42.., 52..)
	const uint64_t *n_bricks;    // device scalar
	const LargeTri *large;
	const UvMap *luv;
	TexView tv;
	RasterParams rp;
	const uint64_t *small_keys; // leaves of the small triangles: sorted unique Morton codes, leaf words, how many
	const uint32_t *small_leaf;
	const uint64_t *n_small;
	// per brick (arrays sized for the upper bound "number of pairs"; entries past n_bricks stay zero)
	uint32_t *temp;         // [512] the 8-word leaf blocks of the brick's depth L-1 nodes, dense, in Morton order
	uint4 *rec;             // x, y: occupancy of the depth L-1 nodes 2l (bit l of x) and 2l+1 (bit l of y); z bits 0..7: of the 8 depth
	                        // L-2 nodes; w: number of leaves | depth L-1 nodes << 10 | depth L-2 nodes << 17 (zeroed: bricks past
	                        // n_bricks count 0); flat bricks: see REC_FLAT
	const uint64_t *rank[3]; // exclusive scans of cnt[] (rank[j][n] = total)
	uint64_t n_bound;       // entries of the per-brick arrays
	uint32_t *slow_list;     // the bricks k_brick_raster has to rasterize (all but the flat ones), in any order
	unsigned long long *n_slow; // how many (zeroed)
	uint64_t *keys_top;      // per non-empty brick = depth L-3 node: Morton code (what k_parent_compact builds the upper levels from)
	uint32_t *first_l2;      // per depth L-3 node: index of its first depth L-2 child
	unsigned char *slot_l2;  // per depth L-2 node: its child slot
	uint64_t *count[3];      // device scalars: leaves (zeroed: accumulated), depth L-1, depth L-2 nodes
	uint64_t *count3;        // depth L-3 nodes
};

// 3 bits -> every third bit (0, 1, 8, 9, 64, 65, 72, 73): one byte permute picks the entry of a table held in two registers
SVO_DEV uint32_t spread3(uint32_t v) {
#if defined(__CUDA_ARCH__)
	return __byte_perm(0x09080100u, 0x49484140u, v);
#else
	return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
#endif
}
// bit i of the result = byte i of b is not zero
SVO_DEV uint32_t nonzero_bytes4(uint32_t b) {
	uint32_t t = b | (b >> 4);
	t |= t >> 2;
	t |= t >> 1;
	return ((t & 0x01010101u) * 0x01020408u) >> 24;
}

// One warp rasterizes BRICK_BPW consecutive bricks, one after the other, into its 512-cell grid; nothing is shared
// between warps and no warp waits for another one: what a brick needs from its neighbours (the ranks of its nodes in the
// level arrays) is left to k_brick_ranks (three scans over the per-brick counts) and k_brick_emit.
#ifndef SVO_BRICK_BPW
#define SVO_BRICK_BPW 8
#endif
constexpr int BRICK_BPW = SVO_BRICK_BPW;
constexpr int LT_WORDS = (int)(sizeof(LargeTri) / 8);
static_assert(sizeof(LargeTri) % 8 == 0 && LT_WORDS <= 32 && BRICK_BPW < 32, "a LargeTri is staged by one warp, 8 bytes per lane");

// coverage of pixel (X + dx, Y + dy), 0 <= dx, dy < 8: the three edge functions (pixel_covered) evaluated at (X, Y) and
// stepped -- ea = 256 A, eb = 256 B with 32-bit A, B (tri_setup), so the step 256 (A dx + B dy) is one 32-bit product sum
SVO_DEV void brick_edges(const TriSetup &ts, int32_t X, int32_t Y, int32_t dx, int32_t dy, int64_t (&e)[3]) {
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		const int64_t e0 = ts.ea[i] * (int64_t)X + ts.eb[i] * (int64_t)Y + ts.ec[i];
		const int32_t step = (int32_t)(ts.ea[i] >> 8) * dx + (int32_t)(ts.eb[i] >> 8) * dy;
		e[i] = e0 + ((int64_t)step << 8);
	}
}

// Bricks whose only pair is FLAT, one thread per brick: 64 voxels in one plane of the brick, one fragment each, all of
// the draw's colour.  The 16 depth L-1 nodes (4 x 4 in the plane) hold four leaves each, in the same slots: 16
// identical blocks, described completely by the 16-byte record written here (occupancy from the axis and the plane's
// position, colour, slot parity -- REC_FLAT); k_brick_emit generates the blocks.  No pixel is looked at, the triangle
// record is not read beyond its colour; all other bricks are put on k_brick_raster's list.
SVO_DEV bool brick_is_solo_flat(uint32_t p0, uint32_t p1, uint64_t first_pair) {
	return p1 - p0 == 1u && ((first_pair >> 32) & 1ull) && ((uint32_t)first_pair & PAIR_FLAT);
}
__global__ void __launch_bounds__(256) k_brick_flat(BrickArgs a) {
	const uint64_t brick = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	const int lane = threadIdx.x & 31;
	const bool valid = brick < *a.n_bricks;
	uint32_t p0 = 0, p1 = 0;
	uint64_t pr = 0;
	if (valid) p0 = a.brick_first[brick], p1 = a.brick_first[brick + 1], pr = a.pairs[p0];
	const bool flat = valid && brick_is_solo_flat(p0, p1, pr);
	{ // every other brick goes on the raster kernel's list (one atomic per warp)
		const unsigned slow = __ballot_sync(FULL_MASK, valid && !flat);
		unsigned long long pos = 0;
		if (lane == 0 && slow) pos = atomicAdd(a.n_slow, (unsigned long long)__popc(slow));
		pos = __shfl_sync(FULL_MASK, pos, 0);
		if (valid && !flat) a.slow_list[pos + (uint32_t)__popc(slow & ((1u << lane) - 1u))] = (uint32_t)brick;
	}
	if (!flat) return;
	const uint32_t pl = (uint32_t)pr, rgb = a.large[pl & PAIR_LI_MASK].rgb & 0xffffffu;
	const uint32_t axis = (pl >> 26) & 3u, dz = (pl >> 28) & 7u;
	uint32_t wx, wy;
	screen_axes(axis, wx, wy);
	// the plane's 4 x 4 nodes: Morton index inside the brick with two bits per axis; node 2l -> bit l of x, 2l + 1 -> bit l of y
	const uint32_t nd = dz >> 1, jd = ((nd & 1u) | ((nd & 2u) << 2)) << axis;
	uint32_t xb = 0, yb = 0, n2b = 0;
#pragma unroll
	for (uint32_t k = 0; k < 16u; ++k) {
		const uint32_t nx = k & 3u, ny = k >> 2;
		const uint32_t j = (((nx & 1u) | ((nx & 2u) << 2)) << wx) | (((ny & 1u) | ((ny & 2u) << 2)) << wy) | jd;
		if (j & 1u) yb |= 1u << (j >> 1); else xb |= 1u << (j >> 1);
	}
#pragma unroll
	for (uint32_t k = 0; k < 4u; ++k) n2b |= 1u << (((k & 1u) << wx) | ((k >> 1) << wy) | ((dz >> 2) << axis));
	a.rec[brick] = make_uint4(xb, yb, n2b | (rgb << 8), 64u | (16u << 10) | (4u << 17) | REC_FLAT | (axis << 22) | ((dz & 1u) << 24));
}

#ifndef SVO_BRICK_MINB
#define SVO_BRICK_MINB 4 // resident blocks per SM asked of the compiler (64 registers per thread)
#endif
template <bool TEX> __global__ void __launch_bounds__(BRICK_BLOCK, SVO_BRICK_MINB) k_brick_raster(BrickArgs a) {
	__align__(16) __shared__ uint32_t s_grid[BRICK_WARPS][BRICK_CELLS]; // leaf words (zeroed per brick; bits: which cells have a first writer)
	__shared__ uint32_t s_bits[BRICK_WARPS][BRICK_CELLS / 32];
	__shared__ uint64_t s_tri[BRICK_WARPS][BRICK_BPW][LT_WORDS]; // the first triangle of every brick of the warp
	__align__(16) __shared__ uint64_t s_meta[BRICK_WARPS][BRICK_BPW][6];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t nb = *a.n_slow; // bricks on the list (k_brick_flat)
	const uint64_t brick0 = ((uint64_t)blockIdx.x * BRICK_WARPS + warp) * BRICK_BPW; // the warp's first list entry
	if (brick0 >= nb) return; // whole warp (no block barrier in this kernel)
	uint32_t *g = s_grid[warp], *bits = s_bits[warp];
	const int32_t dx = lane & 7, dy = lane >> 3;
	const uint32_t sx3 = spread3((uint32_t)dx), sy3a = spread3((uint32_t)dy), sy3b = spread3((uint32_t)dy + 4u);

	// The metadata of the warp's bricks is fetched up front, lane q for list entry q -- brick, pair range, Morton code, first pair,
	// -- and the first triangle of every brick is staged in shared memory by 22 lanes at once: three
	// dependent round trips per BRICK_BPW bricks instead of four per brick (latency, not HBM, is what this kernel waits for).
	if (lane < BRICK_BPW && brick0 + lane < nb) {
		const uint32_t brick = a.slow_list[brick0 + lane];
		const uint32_t p0 = a.brick_first[brick], p1 = a.brick_first[brick + 1];
		uint64_t *m = s_meta[warp][lane];
		m[0] = (uint64_t)p0 | ((uint64_t)p1 << 32);
		m[1] = a.brick_code[brick];
		m[2] = a.pairs[p0];
		m[3] = brick;
		m[4] = a.pairs[p1 - 1]; // the brick's small record, if it has one, is its last pair (the sort is stable, small records are appended)
	}
	__syncwarp();
#pragma unroll
	for (int q = 0; q < BRICK_BPW; ++q) {
		if (brick0 + q >= nb) break;
		const uint64_t pr = s_meta[warp][q][2];
		if (((pr >> 32) & 1ull) && lane < LT_WORDS)
			s_tri[warp][q][lane] = reinterpret_cast<const uint64_t *>(a.large + ((uint32_t)pr & PAIR_LI_MASK))[lane];
	}

#pragma unroll 1
	for (int q = 0; q < BRICK_BPW; ++q) {
		if (brick0 + q >= nb) break; // warp-uniform
		const uint64_t brick = s_meta[warp][q][3];
		if (lane < BRICK_CELLS / 32) bits[lane] = 0u;
		{ // (empty cells must read as zero when the finished blocks are copied out below)
			uint4 *g4 = reinterpret_cast<uint4 *>(g);
#pragma unroll
			for (int k = 0; k < BRICK_CELLS / 128; ++k) g4[k * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
		}
		__syncwarp();
		const ulonglong2 m01 = *reinterpret_cast<const ulonglong2 *>(&s_meta[warp][q][0]);
		const ulonglong2 m23 = *reinterpret_cast<const ulonglong2 *>(&s_meta[warp][q][2]);
		const uint32_t p0 = (uint32_t)m01.x, p1 = (uint32_t)(m01.x >> 32);
		const uint64_t brick_id = m01.y & 0x3fffffffull, pr0 = m23.x;
		const uint32_t bxyz = (uint32_t)(m01.y >> 32);
		// first voxel of the brick, full-grid coordinates
		const uint32_t vb0 = a.rp.origin[0] + 8u * (bxyz & 1023u), vb1 = a.rp.origin[1] + 8u * ((bxyz >> 10) & 1023u),
		               vb2 = a.rp.origin[2] + 8u * (bxyz >> 20);
		// The brick's leaves of small triangles come first in fragment order, its large triangles follow by id: the small
		// record (the last pair) is handled before the others.
		const uint64_t pr_last = s_meta[warp][q][4];
		const uint32_t has_small = ((pr_last >> 32) & 1ull) ? 0u : 1u;
		if (has_small) { // plain stores, cells are unique
			const uint64_t ns = *a.n_small;
			for (uint64_t i = (uint64_t)(uint32_t)pr_last + lane;; i += 32) {
				bool ok = i < ns;
				uint64_t key = 0;
				if (ok) key = a.small_keys[i], ok = (key >> (3 * BRICK_LOG)) == brick_id;
				if (ok) {
					const uint32_t cell = (uint32_t)key & (BRICK_CELLS - 1);
					g[cell] = a.small_leaf[i];
					atomicOr(&bits[cell >> 5], 1u << (cell & 31u));
				}
				if (!__all_sync(FULL_MASK, ok)) break;
			}
			__syncwarp();
		}
		for (uint32_t p = p0; p < p1 - has_small; ++p) {
			const uint64_t pr = p == p0 ? pr0 : a.pairs[p]; // warp-uniform
			{
				// one large triangle: the brick's 8x8 pixels, two per lane (rows dy and dy + 4); a triangle puts at most one
				// fragment into a voxel, so the fold of a cell needs no atomics, and triangles follow each other in order
				const uint32_t li = (uint32_t)pr & PAIR_LI_MASK;
				if (p != p0) { // (the first pair's triangle is staged already)
					if (lane < LT_WORDS) s_tri[warp][q][lane] = reinterpret_cast<const uint64_t *>(a.large + li)[lane];
					__syncwarp();
				}
				const LargeTri &lt = *reinterpret_cast<const LargeTri *>(s_tri[warp][q]);
				const TriSetup &ts = lt.ts;
				const uint32_t axis = ts.axis;
				uint32_t wx, wy;
				screen_axes(axis, wx, wy);
				const int32_t X = (int32_t)(wx == 0u ? vb0 : (wx == 1u ? vb1 : vb2)), Y = (int32_t)(wy == 0u ? vb0 : (wy == 1u ? vb1 : vb2));
				const uint32_t z_base = axis == 0u ? vb0 : (axis == 1u ? vb1 : vb2);
				const int32_t px = X + dx;
				const uint32_t tex = TEX ? lt.textured : 0u;
				int64_t e[3];
				brick_edges(ts, X, Y, dx, dy, e);
				const bool in_x = px >= ts.px0 && px <= ts.px1;
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const int32_t py = Y + dy + 4 * h;
					if (h) {
#pragma unroll
						for (int i = 0; i < 3; ++i) e[i] += ts.eb[i] * 4; // four rows down
					}
					bool ok = in_x && py >= ts.py0 && py <= ts.py1 && (e[0] | e[1] | e[2]) >= 0;
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
						const uint32_t bit = 1u << (cell & 31u);
						const uint32_t old = atomicOr(&bits[cell >> 5], bit);
						g[cell] = (old & bit) ? leaf_accumulate(g[cell], rgb) : leaf_first(rgb);
					}
				}
			}
			__syncwarp();
		}
		// The brick's depth L-1 nodes: lane l owns cells 16 l .. 16 l + 15 = the nodes 2l and 2l + 1.  Every occupied node
		// leaves its finished 8-word block (leaf words, zeros for empty children) in temp, at its rank inside the brick.
		const uint32_t hw = (bits[lane >> 1] >> (16 * (lane & 1))) & 0xffffu;
		const bool o0 = (hw & 0xffu) != 0u, o1 = (hw >> 8) != 0u;
		const uint32_t b0 = __ballot_sync(FULL_MASK, o0), b1 = __ballot_sync(FULL_MASK, o1);
		const uint32_t c0 = warp_sum((uint32_t)__popc(hw));
		const uint32_t n2 = __ballot_sync(FULL_MASK, lane < 8 && (((b0 | b1) >> (4 * lane)) & 0xfu) != 0u); // 8 nodes = 4 lanes
		if (lane == 0) // w: the brick's counts (leaves 10 bits, depth L-1 nodes 7 bits, depth L-2 nodes 4 bits) for the rank scans
			a.rec[brick] = make_uint4(b0, b1, n2, c0 | ((uint32_t)(__popc(b0) + __popc(b1)) << 10) | ((uint32_t)__popc(n2) << 17));
		if (hw) {
			const uint32_t lt = (1u << lane) - 1u;
			uint32_t rank = (uint32_t)(__popc(b0 & lt) + __popc(b1 & lt));
			uint4 *dst = reinterpret_cast<uint4 *>(a.temp + brick * BRICK_CELLS);
			const uint4 *src = reinterpret_cast<const uint4 *>(g + 16 * lane);
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const uint32_t m = (hw >> (8 * h)) & 0xffu;
				if (m) {
					dst[2 * rank] = src[2 * h], dst[2 * rank + 1] = src[2 * h + 1];
					++rank;
				}
			}
		}
		__syncwarp(); // the grid and the bits are reused by the next brick
	}
}

// rank of the depth L-1 node `node` (0..63) among the brick's occupied ones; node 2l is bit l of x, node 2l+1 bit l of y
SVO_DEV uint32_t brick_node_rank(uint32_t x, uint32_t y, uint32_t node) {
	const uint32_t l = node >> 1, below = (1u << l) - 1u;
	return (uint32_t)(__popc(x & below) + __popc(y & below)) + ((node & 1u) ? ((x >> l) & 1u) : 0u);
}

// The ranks of every brick's nodes in the two levels above the leaves, the node counts of the four deepest levels, and
// the depth L-3 level itself (a brick IS a depth L-3 node): three exclusive scans over the 16-byte records -- depth L-1
// nodes (bits 10..16 of word w), depth L-2 nodes (17..20), non-empty bricks -- in one pass.  Records are read striped
// (consecutive lanes consecutive records: 512 bytes per warp load) and the three counts travel packed in one 64-bit
// word through the warp scans (a tile of 2048 bricks holds <= 2^17, 2^14 and 2^11 of them); warps 0, 1, 2 walk the
// three look-back chains at the same time; leaves are only counted.  With its ranks in hand a thread writes what
// k_parent_compact would have derived from the depth L-2 keys: the brick's Morton code as a depth L-3 key, the index of
// its first depth L-2 child and those children's slots.  rank1 / rank2: n + 1 entries each (entry n = total).
constexpr uint32_t RANK_SH2 = 18, RANK_SH3 = 33;
constexpr uint64_t RANK_M1 = (1ull << RANK_SH2) - 1, RANK_M2 = (1ull << (RANK_SH3 - RANK_SH2)) - 1;
static_assert(SCAN_TILE * 64ull <= RANK_M1 && SCAN_TILE * 8ull <= RANK_M2, "packed tile sums");
#ifndef SVO_RANKS_MINB
#define SVO_RANKS_MINB 3
#endif
__global__ void __launch_bounds__(SCAN_BLOCK, SVO_RANKS_MINB)
    k_brick_ranks(BrickArgs a, uint64_t *__restrict__ rank1, uint64_t *__restrict__ rank2, uint64_t *state, uint32_t *ticket, uint64_t state_stride) {
	constexpr int NW = SCAN_BLOCK / 32, NC = SCAN_ITEMS * NW; // (row, warp) cells of a tile
	static_assert(NC % 32 == 0 && NC <= 128, "the cells are scanned by one warp");
	__shared__ uint64_t s_cell[NC];
	__shared__ uint32_t s_ticket;
	__shared__ unsigned long long s_leaves;
	__shared__ uint64_t s_prefix[3];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t n = a.n_bound;
	const uint32_t tile = take_ticket(ticket, &s_ticket);
	if (threadIdx.x == 0) s_leaves = 0;
	const uint64_t base = (uint64_t)tile * SCAN_TILE + threadIdx.x;
	uint64_t inc[SCAN_ITEMS]; // inclusive warp scan of the packed counts, row by row
	uint32_t own[SCAN_ITEMS]; // the record's counts (bits 10..20) | occupancy of its 8 depth L-2 nodes << 21
	uint32_t leaves = 0;
	auto packed = [](uint32_t o) {
		return (uint64_t)((o >> 10) & 0x7fu) | ((uint64_t)((o >> 17) & 0xfu) << RANK_SH2) | ((uint64_t)((o >> 21) != 0u) << RANK_SH3);
	};
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		const uint64_t e = base + (uint64_t)i * SCAN_BLOCK;
		uint4 r = make_uint4(0u, 0u, 0u, 0u);
		if (e < n) r = a.rec[e];
		own[i] = (r.w & 0x1ffc00u) | ((r.z & 0xffu) << 21);
		leaves += r.w & 0x3ffu;
		inc[i] = warp_inclusive_sum(packed(own[i]), lane);
		if (lane == 31) s_cell[i * NW + warp] = inc[i];
	}
	leaves = warp_sum(leaves);
	__syncthreads();
	if (lane == 0 && leaves) atomicAdd(&s_leaves, (unsigned long long)leaves);
	if (warp == 0) { // exclusive scan of the cells, in element order (row-major); the tile's packed total is left in s_prefix[0]
		constexpr int CPL = NC / 32;
		uint64_t c[CPL], t = 0;
#pragma unroll
		for (int k = 0; k < CPL; ++k) c[k] = s_cell[lane * CPL + k], t += c[k];
		const uint64_t tinc = warp_inclusive_sum(t, lane);
		uint64_t run = tinc - t;
#pragma unroll
		for (int k = 0; k < CPL; ++k) s_cell[lane * CPL + k] = run, run += c[k];
		if (lane == 31) s_prefix[0] = tinc;
	}
	__syncthreads();
	const uint64_t total = s_prefix[0];
	if (threadIdx.x == 0 && s_leaves) atomicAdd(reinterpret_cast<unsigned long long *>(a.count[0]), s_leaves);
	__syncthreads();
	// (the bricks' codes are fetched now: their latency passes while the look-back warps walk their chains)
	uint32_t code[SCAN_ITEMS];
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		const uint64_t e = base + (uint64_t)i * SCAN_BLOCK;
		code[i] = (own[i] >> 21) ? (uint32_t)a.brick_code[e] & 0x3fffffffu : 0u;
	}
	if (warp < 3) { // warp y walks the look-back chain of scan y
		const uint64_t mine = warp == 0 ? (total & RANK_M1) : (warp == 1 ? ((total >> RANK_SH2) & RANK_M2) : (total >> RANK_SH3));
		const uint64_t p = lookback_exclusive(state + (uint64_t)warp * state_stride, tile, mine, lane);
		if (lane == 0) s_prefix[warp] = p;
	}
	__syncthreads();
	const uint64_t p1 = s_prefix[0], p2 = s_prefix[1], p3 = s_prefix[2];
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		const uint64_t e = base + (uint64_t)i * SCAN_BLOCK;
		if (e > n) continue;
		const uint64_t pk = s_cell[i * NW + warp] + inc[i] - packed(own[i]);
		const uint64_t r1 = p1 + (pk & RANK_M1), r2 = p2 + ((pk >> RANK_SH2) & RANK_M2), r3 = p3 + (pk >> RANK_SH3);
		rank1[e] = r1, rank2[e] = r2; // (e == n: the element behind the last record receives the totals)
		if (e == n) *a.count[1] = r1, *a.count[2] = r2, *a.count3 = r3;
		uint32_t n2 = own[i] >> 21;
		if (n2) { // a non-empty brick (records past the last brick are zero)
			a.keys_top[r3] = code[i];
			a.first_l2[r3] = (uint32_t)r2;
			unsigned char *sl = a.slot_l2 + r2;
			for (; n2; n2 &= n2 - 1u) *sl++ = (unsigned char)(__ffs((int)n2) - 1);
		}
	}
}

// Node words of the two deepest windows, straight from the bricks (the bulk of the node buffer): the 8-word blocks of
// a brick's depth L-1 nodes (leaf words: a copy of the brick's part of temp) and of its depth L-2 nodes (pointers to
// the former) are consecutive in their windows, at the brick's ranks.  Placement as in k_emit_octree: block g ->
// words[(g - block_shift) * 8], a pointer to block c is (c - block_shift) * 8 + ptr_bias.  BRICK_EMIT_LANES lanes per brick.
// Which part of the work one launch does (BrickEmit::parts): a multi-GPU gather sends only what cannot be regenerated
// -- the leaf blocks of the rasterized bricks (BRICK_EMIT_COPY, stored over NVLink) and the 32 bytes of record + ranks per
// brick -- and the GPU that owns the stitched buffer generates the flat bricks' blocks and all pointer blocks itself, at
// local HBM speed, from the records alone (BRICK_EMIT_FLAT | BRICK_EMIT_PTRS: a.rec and a.rank[1..2] are all that is read).
constexpr uint32_t BRICK_EMIT_COPY = 1u, BRICK_EMIT_FLAT = 2u, BRICK_EMIT_PTRS = 4u, BRICK_EMIT_ALL = 7u;
struct BrickEmit {
	uint64_t block_l1, block_l; // first block of the window that holds the children of the depth L-2 / of the depth L-1 nodes
	uint32_t block_shift, ptr_bias;
	uint64_t n_bricks;
	uint32_t parts;
};
#ifndef SVO_BRICK_EMIT_LANES
#define SVO_BRICK_EMIT_LANES 4 // lanes per brick
#endif
constexpr uint32_t BRICK_EMIT_LANES = SVO_BRICK_EMIT_LANES;
static_assert(BRICK_EMIT_LANES == 1 || BRICK_EMIT_LANES == 2 || BRICK_EMIT_LANES == 4 || BRICK_EMIT_LANES == 8 || BRICK_EMIT_LANES == 16, "a power of two: lanes share the brick's 16..64 leaf blocks and 8 pointer blocks");
// one 8-word block with a single 256-bit store (sm_100: STG.256); p is 32-byte aligned
SVO_DEV void store_block(uint32_t *p, const uint4 &lo, const uint4 &hi) {
#if defined(__CUDA_ARCH__) && !defined(SVO_EMIT_NO_V8)
#ifndef SVO_EMIT_ST
#define SVO_EMIT_ST "st.global.v8.b32"
#endif
	asm volatile(SVO_EMIT_ST " [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x),
	             "r"(hi.y), "r"(hi.z), "r"(hi.w)
	             : "memory");
#else
	reinterpret_cast<uint4 *>(p)[0] = lo, reinterpret_cast<uint4 *>(p)[1] = hi;
#endif
}
__global__ void __launch_bounds__(BRICK_BLOCK) k_brick_emit(BrickArgs a, BrickEmit be, uint32_t *__restrict__ words) {
	const uint64_t tid = (uint64_t)blockIdx.x * BRICK_BLOCK + threadIdx.x;
	const uint64_t brick = tid / BRICK_EMIT_LANES;
	const uint32_t sub = threadIdx.x % BRICK_EMIT_LANES;
	if (brick >= be.n_bricks) return;
	// (all three loads are issued before anything depends on one of them: the kernel waits for latency, not for bandwidth)
	const uint4 rec = a.rec[brick];
	const uint64_t r1 = a.rank[1][brick], r2 = a.rank[2][brick];
	const uint32_t c1 = (uint32_t)(__popc(rec.x) + __popc(rec.y));
	if (c1 == 0u) return;
	if (!(be.parts & ((rec.w & REC_FLAT) ? (BRICK_EMIT_FLAT | BRICK_EMIT_PTRS) : (BRICK_EMIT_COPY | BRICK_EMIT_PTRS)))) return;
	const uint64_t g1 = be.block_l + r1 - be.block_shift; // where the brick's first leaf block goes
	uint32_t *dst = words + g1 * 8;
	if (rec.w & REC_FLAT) { // 16 identical blocks: slot s holds the leaf iff its bit `axis` equals the plane's depth parity
		const uint32_t axis = (rec.w >> 22) & 3u, par = (rec.w >> 24) & 1u, leaf = leaf_first(rec.z >> 8);
		uint4 lo, hi;
		lo.x = ((0u >> axis) & 1u) == par ? leaf : 0u, lo.y = ((1u >> axis) & 1u) == par ? leaf : 0u;
		lo.z = ((2u >> axis) & 1u) == par ? leaf : 0u, lo.w = ((3u >> axis) & 1u) == par ? leaf : 0u;
		hi.x = ((4u >> axis) & 1u) == par ? leaf : 0u, hi.y = ((5u >> axis) & 1u) == par ? leaf : 0u;
		hi.z = ((6u >> axis) & 1u) == par ? leaf : 0u, hi.w = ((7u >> axis) & 1u) == par ? leaf : 0u;
		if (be.parts & BRICK_EMIT_FLAT) {
#pragma unroll
			for (uint32_t i = 0; i < 16u; i += BRICK_EMIT_LANES) store_block(dst + 8u * (i + sub), lo, hi); // consecutive lanes consecutive blocks
		}
	} else if (be.parts & BRICK_EMIT_COPY) {
		const uint4 *src = reinterpret_cast<const uint4 *>(a.temp + brick * BRICK_CELLS);
		for (uint32_t i = sub; i < c1; i += BRICK_EMIT_LANES) store_block(dst + 8u * i, src[2u * i], src[2u * i + 1u]);
	}
	const uint32_t n2 = rec.z & 0xffu;
	if (!(be.parts & BRICK_EMIT_PTRS)) return;
#pragma unroll
	for (uint32_t nd = sub; nd < 8u; nd += BRICK_EMIT_LANES) { // a depth L-2 node: one block of pointers to its children's blocks
		if (!((n2 >> nd) & 1u)) continue;
		uint32_t c = (uint32_t)g1 + brick_node_rank(rec.x, rec.y, 8u * nd);
		uint32_t w[8];
#pragma unroll
		for (int sl = 0; sl < 8; ++sl) {
			const uint32_t node = 8u * nd + (uint32_t)sl;
			const bool occ = (((node & 1u) ? rec.y : rec.x) >> (node >> 1)) & 1u;
			w[sl] = occ ? (0x80000000u | ((c++ << 3) + be.ptr_bias)) : 0u;
		}
		const uint64_t g = be.block_l1 + r2 + (uint32_t)__popc(n2 & ((1u << nd) - 1u)) - be.block_shift;
		store_block(words + g * 8, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
	}
}
} // namespace svo
