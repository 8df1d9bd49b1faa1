// raster.cuh -- the Voxelizer kernels: triangle-parallel fixed-point rasterization into 64-bit Morton
// fragments with count -> scan -> emit (no per-fragment global atomic; deterministic order).
//
// Replaces voxelizer.vert/.geom/.frag + the fixed-function rasterizer + the count pass
// (src/Voxelizer.cpp:134-179).  Two work classes, decided per triangle from its candidate rectangle:
//   small (area <= SMALL_AREA pixels): one thread walks the
//          rectangle with exact integer edge functions, stepped per pixel; the search for the next covered
//          pixel is a loop of its own so that the warp's lanes meet again before the per-fragment work;
//   large: warps compute exact per-row spans (integer division, no per-pixel tests); emission is
//          output-parallel over the span pixels (load-balanced row search, 8 consecutive fragments per thread
//          with the depth plane and the (x, y) Morton code kept in registers), so huge wall/floor triangles
//          are spread over the whole GPU and stored with fully coalesced 16-byte writes.
// Morton codes come from a 1024-entry spread table in shared memory (two lookups per coordinate above level 10).
// Textured draws (texture.cuh) sample per fragment; the TEX template parameter keeps the untextured kernels free
// of that code.  Large triangles of alpha-tested textures: rows counted by sampling, emitted by k_emit_alpha_rows.
// Fragment order: all small-triangle fragments in triangle order, then all large-triangle fragments in
// triangle / row / x order.  A voxel receives at most one fragment per triangle, so together with the
// stable sort this fixes the colour-averaging order per voxel: (class, triangle id).
#pragma once
#include "scan.cuh"
#include "svo_math.cuh"
#include "texture.cuh"

namespace svo {

constexpr int64_t SMALL_AREA = 256; // candidate-rectangle pixels handled by a single thread
#ifndef SVO_RASTER_BLOCK
#define SVO_RASTER_BLOCK 128
#endif
constexpr int RASTER_BLOCK = SVO_RASTER_BLOCK;
constexpr int LARGE_ROW_CHUNK = 128; // rows of a large triangle handled by one warp before it strides on
#ifndef SVO_EMIT_BLOCK
#define SVO_EMIT_BLOCK 256
#endif
#ifndef SVO_EMIT_ITEMS
#define SVO_EMIT_ITEMS 8
#endif
constexpr int EMIT_BLOCK = SVO_EMIT_BLOCK, EMIT_ITEMS = SVO_EMIT_ITEMS, EMIT_TILE = EMIT_BLOCK * EMIT_ITEMS;

struct DrawRec {
	uint32_t first_index, tri_base, tri_count, rgb;
	uint32_t tex; // 0xffffffff = untextured (voxelizer.frag:35)
};
struct SceneView {
	const unsigned char *pos;
	uint32_t stride;
	const uint32_t *idx;
	const DrawRec *draws;
	uint32_t n_draws;
	uint64_t n_tri;
	const unsigned char *uv; // texture coordinates (2 floats per vertex), only read for textured draws
	uint32_t uv_stride;
	TexView tex;
};
struct RasterParams {
	uint32_t res;       // 1 << level (full grid)
	int mode;
	ShardBox sb;        // voxel window, full-grid coordinates
	uint32_t origin[3]; // sb.lo: fragments are emitted in shard-local coordinates
};

// largest vertex index of the mesh (svo_scene_create validates it against the vertex count)
__global__ void __launch_bounds__(256) k_max_index(const uint32_t *__restrict__ idx, uint64_t n, uint32_t *__restrict__ out) {
	uint32_t m = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) m = idx[i] > m ? idx[i] : m;
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const uint32_t o = __shfl_xor_sync(FULL_MASK, m, d);
		m = o > m ? o : m;
	}
	if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

SVO_DEV uint32_t find_draw(const SceneView &sv, uint64_t t) {
	uint32_t lo = 0, hi = sv.n_draws; // last draw with tri_base <= t
	while (hi - lo > 1) {
		uint32_t mid = (lo + hi) >> 1;
		if (sv.draws[mid].tri_base <= t) lo = mid; else hi = mid;
	}
	return lo;
}

// How a triangle's fragments get their colour: the draw's albedo, or a texture sample through um.
// alpha = the texture has texels that can fail the alpha test (voxelizer.frag:29-30), so fragments may be discarded.
struct TriShade {
	uint32_t rgb;
	bool textured, alpha;
	UvMap um;
};

// TEX = the scene has textured draws (the untextured instantiation carries no texture code at all)
template <bool TEX>
SVO_DEV bool load_and_setup(const SceneView &sv, const RasterParams &rp, uint64_t t, TriSetup &ts, TriShade &sh) {
	const uint32_t d = find_draw(sv, t);
	const DrawRec dr = sv.draws[d];
	const uint32_t *ix = sv.idx + dr.first_index + 3u * (uint32_t)(t - dr.tri_base);
	float p[3][3];
	uint32_t vi[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		vi[i] = __ldg(ix + i);
		const float *v = reinterpret_cast<const float *>(sv.pos + (size_t)vi[i] * sv.stride);
		p[i][0] = __ldg(v), p[i][1] = __ldg(v + 1), p[i][2] = __ldg(v + 2);
	}
	sh.rgb = dr.rgb;
	sh.textured = sh.alpha = false;
	const bool ok = tri_setup(p[0], p[1], p[2], rp.res, rp.mode, rp.sb, ts);
	if (TEX && ok && dr.tex != 0xffffffffu) {
		float uv[3][2];
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			const float *v = reinterpret_cast<const float *>(sv.uv + (size_t)vi[i] * sv.uv_stride);
			uv[i][0] = __ldg(v), uv[i][1] = __ldg(v + 1);
		}
		uvmap_setup(sv.tex, dr.tex, p, uv, ts.axis, rp.res, sh.um);
		sh.textured = true;
		sh.alpha = sv.tex.desc[dr.tex].has_alpha != 0u;
	}
	return ok;
}

// A fragment = morton(voxel) << 24 | rgb, with the Morton spread taken from a 1024-entry table in shared memory
// (part1by2_10 of every 10-bit value): 6 lookups instead of ~90 ALU instructions.
SVO_DEV void fill_spread_table(uint32_t *s_lut, int n_threads) {
	for (uint32_t i = threadIdx.x; i < 1024u; i += (uint32_t)n_threads) s_lut[i] = part1by2_10(i);
}
SVO_DEV uint64_t make_fragment_lut(const TriSetup &ts, const RasterParams &rp, const uint32_t *s_lut, int32_t px, int32_t py, uint32_t uz,
                                   uint32_t rgb) {
	uint32_t vx, vy, vz;
	unswizzle(ts.axis, (uint32_t)px, (uint32_t)py, uz, vx, vy, vz);
	vx -= rp.origin[0], vy -= rp.origin[1], vz -= rp.origin[2];
	const uint32_t m_lo = s_lut[vx & 1023u] | (s_lut[vy & 1023u] << 1) | (s_lut[vz & 1023u] << 2);
	uint32_t m_hi = 0;
	if (rp.res > 1024u) m_hi = s_lut[(vx >> 10) & 1023u] | (s_lut[(vy >> 10) & 1023u] << 1) | (s_lut[(vz >> 10) & 1023u] << 2);
	return ((uint64_t)((m_lo >> 8) | (m_hi << 22)) << 32) | (uint64_t)((m_lo << 24) | (rgb & 0xffffffu));
}

// Row-major walk over the covered pixels of a small triangle's candidate rectangle.  The three edge functions are
// stepped (one 64-bit add per edge and pixel) instead of re-evaluated, and the search for the next covered pixel is
// a tight loop of its own: the lanes of a warp -- one triangle each, rectangles of different sizes -- meet again
// before the expensive per-fragment work, so that work runs with most lanes active.
struct SmallWalk {
	int64_t e[3], erow[3];
	int32_t px, py;
	SVO_DEV void begin(const TriSetup &ts) {
		px = ts.px0, py = ts.py0;
#pragma unroll
		for (int i = 0; i < 3; ++i) e[i] = erow[i] = ts.ea[i] * (int64_t)px + ts.eb[i] * (int64_t)py + ts.ec[i];
	}
	SVO_DEV void step(const TriSetup &ts) {
		if (px < ts.px1) {
			++px;
#pragma unroll
			for (int i = 0; i < 3; ++i) e[i] += ts.ea[i];
		} else {
			px = ts.px0, ++py;
#pragma unroll
			for (int i = 0; i < 3; ++i) e[i] = erow[i] += ts.eb[i];
		}
	}
	// moves to the next covered pixel at or after the current one; false when the rectangle is exhausted
	SVO_DEV bool next(const TriSetup &ts) {
		while (py <= ts.py1) {
			if ((e[0] | e[1] | e[2]) >= 0) return true; // pixel_covered: all three edge functions >= 0
			step(ts);
		}
		return false;
	}
};

// ---- pass 1: classify + count small ------------------------------------------------------------------
// cnt_small[t]  = fragments of a small triangle (0 for large / culled)
// rows[t]       = candidate rows of a large triangle (>= 1), 0 for small / culled ones
// Triangles of an alpha-tested texture produce a fragment only where the sample passes the alpha test
// (voxelizer.frag:29-30 discards before the counter): small ones sample while they count; large ones get their rows
// counted by sampling (k_large_rows) and emitted by k_emit_alpha_rows instead of the span arithmetic.
template <bool TEX>
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_classify_count(SceneView sv, RasterParams rp, uint32_t *__restrict__ cnt_small, uint32_t *__restrict__ rows) {
	const uint64_t t = (uint64_t)blockIdx.x * RASTER_BLOCK + threadIdx.x;
	if (t >= sv.n_tri) return;
	TriSetup ts;
	TriShade sh;
	uint32_t cnt = 0, pk = 0;
	if (load_and_setup<TEX>(sv, rp, t, ts, sh)) {
		if (ts.full_area <= SMALL_AREA) {
			SmallWalk wk;
			wk.begin(ts);
			while (wk.next(ts)) {
				uint32_t uz, c;
				if ((!ts.cull_depth && !ts.clip_z) || pixel_fragment(ts, rp.res, wk.px, wk.py, uz))
					if (!(TEX && sh.alpha) || sample_colour(sv.tex, sh.um, wk.px, wk.py, c)) ++cnt;
				wk.step(ts);
			}
		} else
			pk = (uint32_t)(ts.py1 - ts.py0 + 1);
	}
	cnt_small[t] = cnt;
	rows[t] = pk;
}

// ---- pass 1b: large triangles ---------------------------------------------------------------------------
struct LargeTri {
	TriSetup ts;
	uint32_t rgb;
	uint32_t tri;      // triangle id
	uint32_t row_base; // first row in the (sparse) row arrays
	uint32_t textured; // bit 0: colour comes from luv[li] (textured scenes only); bit 1: alpha-tested texture
};

// gather the large triangles into a dense list (order = triangle order)
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_large_collect(uint64_t n_tri, const uint32_t *__restrict__ rows, const uint64_t *__restrict__ large_index /* scan of rows != 0 */,
                    const uint64_t *__restrict__ row_prefix /* scan of rows */, LargeTri *__restrict__ large) {
	const uint64_t t = (uint64_t)blockIdx.x * RASTER_BLOCK + threadIdx.x;
	if (t >= n_tri) return;
	if (rows[t]) {
		LargeTri &lt = large[large_index[t]];
		lt.tri = (uint32_t)t;
		lt.row_base = (uint32_t)row_prefix[t]; // the host checks that all rows together stay below 2^32
	}
}

// one warp per large triangle: exact row spans.  row_len[r] = fragments of the row (0: empty) ; row_x0[r] = first pixel
template <bool TEX>
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_large_rows(SceneView sv, RasterParams rp, uint32_t n_large, LargeTri *__restrict__ large, UvMap *__restrict__ luv,
                 uint32_t *__restrict__ row_len, uint32_t *__restrict__ row_x0) {
	// a triangle's rows are shared out in chunks of LARGE_ROW_CHUNK among gridDim.y warps (a wall has 4096 rows)
	const uint32_t li = (blockIdx.x * RASTER_BLOCK + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (li >= n_large) return; // whole warp leaves together
	LargeTri &lt = large[li];
	TriSetup ts;
	TriShade sh;
	load_and_setup<TEX>(sv, rp, lt.tri, ts, sh); // true by construction (classified large)
	if (lane == 0 && blockIdx.y == 0) {
		lt.ts = ts;
		lt.rgb = sh.rgb;
		lt.textured = (sh.textured ? 1u : 0u) | (sh.alpha ? 2u : 0u);
		if (TEX && sh.textured) luv[li] = sh.um;
	}
	const int32_t h = ts.py1 - ts.py0 + 1;
	for (int32_t c = (int32_t)blockIdx.y * LARGE_ROW_CHUNK; c < h; c += (int32_t)gridDim.y * LARGE_ROW_CHUNK)
	for (int32_t r = c + lane; r < h && r < c + LARGE_ROW_CHUNK; r += 32) {
		int32_t x_lo, x_hi;
		row_span(ts, ts.py0 + r, x_lo, x_hi);
		row_span_depth_window(ts, rp.res, ts.py0 + r, x_lo, x_hi);
		uint32_t len = x_hi >= x_lo ? (uint32_t)(x_hi - x_lo + 1) : 0u;
		if (TEX && sh.alpha && len) { // alpha-tested: the row holds as many fragments as samples survive
			len = 0;
			uint32_t c;
			for (int32_t px = x_lo; px <= x_hi; ++px) len += sample_colour(sv.tex, sh.um, px, ts.py0 + r, c) ? 1u : 0u;
		}
		row_len[lt.row_base + r] = len;
		row_x0[lt.row_base + r] = (uint32_t)x_lo;
	}
}

// dense non-empty rows: off (first large-fragment ordinal), x0 | y << 16, owning large triangle
struct DenseRows {
	uint32_t *off; // n_rows + 1
	uint32_t *xy;
	uint32_t *li;
};
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_rows_compact(uint32_t n_large, const LargeTri *__restrict__ large, const uint32_t *__restrict__ row_len,
                   const uint32_t *__restrict__ row_x0, const uint64_t *__restrict__ dense_index /* scan of row_len != 0 */,
                   const uint64_t *__restrict__ frag_prefix /* scan of row_len */, uint64_t n_rows_sparse, DenseRows out) {
	// one warp per large triangle again, so that a row knows its triangle without a search
	const uint32_t li = (blockIdx.x * RASTER_BLOCK + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (li >= n_large) return;
	const LargeTri &lt = large[li];
	const int32_t h = lt.ts.py1 - lt.ts.py0 + 1;
	for (int32_t c = (int32_t)blockIdx.y * LARGE_ROW_CHUNK; c < h; c += (int32_t)gridDim.y * LARGE_ROW_CHUNK)
	for (int32_t r = c + lane; r < h && r < c + LARGE_ROW_CHUNK; r += 32) {
		const uint64_t s = (uint64_t)lt.row_base + r;
		if (row_len[s]) {
			const uint32_t k = (uint32_t)dense_index[s];
			out.off[k] = (uint32_t)frag_prefix[s];
			out.xy[k] = row_x0[s] | ((uint32_t)(lt.ts.py0 + r) << 16);
			out.li[k] = li;
		}
	}
	if (li == 0 && lane == 0 && blockIdx.y == 0) {
		out.off[dense_index[n_rows_sparse]] = (uint32_t)frag_prefix[n_rows_sparse];
	}
}

// ---- pass 2: emit ---------------------------------------------------------------------------------------
template <bool TEX>
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_emit_small(SceneView sv, RasterParams rp, const uint64_t *__restrict__ tri_off, uint64_t *__restrict__ frags) {
	__shared__ uint32_t s_lut[1024];
	fill_spread_table(s_lut, RASTER_BLOCK);
	__syncthreads();
	const uint64_t t = (uint64_t)blockIdx.x * RASTER_BLOCK + threadIdx.x;
	if (t >= sv.n_tri) return;
	const uint64_t o0 = tri_off[t], o1 = tri_off[t + 1];
	if (o0 == o1) return; // large, culled or empty
	TriSetup ts;
	TriShade sh;
	if (!load_and_setup<TEX>(sv, rp, t, ts, sh)) return;
	uint64_t o = o0;
	SmallWalk wk;
	wk.begin(ts);
	while (wk.next(ts)) {
		uint32_t uz, rgb = sh.rgb;
		if (pixel_fragment(ts, rp.res, wk.px, wk.py, uz))
			if (!(TEX && sh.textured) || sample_colour(sv.tex, sh.um, wk.px, wk.py, rgb))
				frags[o++] = make_fragment_lut(ts, rp, s_lut, wk.px, wk.py, uz, rgb);
		wk.step(ts);
	}
}

// output-parallel expansion of the dense row spans: fragment ordinal j -> (row, x).  A block owns EMIT_TILE
// consecutive fragments, a thread EMIT_ITEMS consecutive ones: one search per thread, then x advances by
// incrementing the Morton-spread coordinate, the depth plane's row term is hoisted, and the fragments leave in
// 16-byte stores.
template <bool TEX>
__global__ void __launch_bounds__(EMIT_BLOCK)
    k_emit_large(RasterParams rp, TexView tv, const LargeTri *__restrict__ large, const UvMap *__restrict__ luv, DenseRows rows,
                 uint32_t n_rows, uint64_t n_frag_large, uint64_t *__restrict__ frags /* already offset to the large region */) {
	__shared__ uint32_t s_off[EMIT_TILE + 2];
	__shared__ uint32_t s_lut[1024]; // part1by2_10: Morton spread of 10 coordinate bits
	__shared__ uint32_t s_first, s_last;
	const int lane = threadIdx.x & 31;
	const uint64_t c0 = (uint64_t)blockIdx.x * EMIT_TILE;
	const uint64_t c1 = c0 + EMIT_TILE < n_frag_large ? c0 + EMIT_TILE : n_frag_large;
	if (threadIdx.x < 64) {
		// warp 0: last row with off <= c0 (rows are non-empty, so it contains fragment c0); warp 1: last row with
		// off <= c1 - 1 (contains the tile's last fragment).  32-ary searches, one ballot per level.
		const uint64_t target = threadIdx.x < 32 ? c0 : c1 - 1;
		uint32_t lo = 0, hi = n_rows; // invariant: off[lo] <= target < off[hi]
		while (hi - lo > 1) {
			const uint32_t step = (hi - lo + 31u) / 32u;
			const uint32_t idx = lo + (uint32_t)lane * step;
			const bool ok = idx < hi && rows.off[idx] <= target; // monotone in the lane; lane 0 always holds
			const unsigned b = __ballot_sync(FULL_MASK, ok);
			const uint32_t k = 31u - (uint32_t)__clz((int)b);
			lo += k * step;
			hi = lo + step < hi ? lo + step : hi;
		}
		if (lane == 0) (threadIdx.x < 32 ? s_first : s_last) = lo;
	}
	for (uint32_t i = threadIdx.x; i < 1024u; i += EMIT_BLOCK) s_lut[i] = part1by2_10(i);
	__syncthreads();
	const uint32_t r_first = s_first, n_staged = s_last - s_first + 1u; // rows that intersect [c0, c1): at most EMIT_TILE
	for (uint32_t i = threadIdx.x; i <= n_staged; i += EMIT_BLOCK) s_off[i] = rows.off[r_first + i]; // + the end of the last one
	__syncthreads();

	const uint64_t j0 = c0 + (uint64_t)threadIdx.x * EMIT_ITEMS;
	if (j0 >= c1) return;
	uint32_t lo = 0, hi = n_staged; // last staged row with off <= j0
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (s_off[mid] <= j0) lo = mid; else hi = mid;
	}
	// Per row the thread keeps the depth plane and the Morton code of (x, y) in registers; per fragment it evaluates
	// the depth (the same operations as pixel_depth_row), looks up the spread of the depth voxel, and steps x
	// directly in spread form.  Keys are assembled as two 32-bit halves: the low 10 bits of the three coordinates
	// make the low 30 Morton bits, bits 10.. (levels above 10 only) the rest.
	const bool deep = rp.res > 1024u;
	uint64_t out[EMIT_ITEMS];
	uint32_t row_end = 0; // forces the row set-up on the first fragment
	double row_term = 0.0, dzdx = 0.0, dx = 0.0;
	uint32_t zr_lo = 0, zr_hi = 0, oz = 0, rgb = 0;
	uint32_t xs = 0, xmask = 0, xone = 0, xl = 0, row_lo = 0, row_hi = 0, hi_xrow = 0, shx = 0, shz = 0;
	int32_t px = 0, py_tex = 0;
	const UvMap *um = nullptr; // non-null: the row belongs to a textured triangle
	--lo;
#pragma unroll
	for (int k = 0; k < EMIT_ITEMS; ++k) {
		const uint64_t j = j0 + k;
		if (j < c1) {
			if (j >= row_end) { // next row (the first one included)
				++lo;
				const uint32_t r = r_first + lo;
				const uint32_t xy = rows.xy[r];
				const LargeTri *lt = &large[rows.li[r]];
				row_end = s_off[lo + 1];
				px = (int32_t)(xy & 0xffffu) + (int32_t)(j - s_off[lo]);
				const int32_t py = (int32_t)(xy >> 16);
				row_term = depth_row_term(lt->ts, py);
				dzdx = lt->ts.dzdx, zr_lo = lt->ts.zr_lo, zr_hi = lt->ts.zr_hi;
				dx = (double)((px * 256 + 128) - lt->ts.X0); // exact; advancing it by 256.0 per pixel stays exact
				// voxel = axis 0: (uz, px, py); 1: (py, uz, px); 2: (px, py, uz)   (voxelizer.frag:24)
				const uint32_t axis = lt->ts.axis;
				const uint32_t wx = axis == 0u ? 1u : (axis == 1u ? 2u : 0u); // world axis of screen x
				const uint32_t wy = axis == 0u ? 2u : (axis == 1u ? 0u : 1u);
				shx = wx, shz = axis;
				oz = pick3(rp.origin, axis);
				xl = (uint32_t)px - pick3(rp.origin, wx);
				const uint32_t yl = (uint32_t)py - pick3(rp.origin, wy);
				xmask = 0x09249249u << wx, xone = 1u << wx;
				xs = s_lut[xl & 1023u] << wx;
				row_lo = s_lut[yl & 1023u] << wy;
				row_hi = s_lut[(yl >> 10) & 1023u] << wy;
				hi_xrow = (s_lut[(xl >> 10) & 1023u] << wx) | row_hi;
				rgb = lt->rgb & 0xffffffu;
				if (TEX) {
					um = (lt->textured & 1u) ? &luv[rows.li[r]] : nullptr;
					if (lt->textured & 2u) um = nullptr; // alpha-tested rows are rewritten by k_emit_alpha_rows
					py_tex = py;
				}
			}
			if (TEX && um) sample_colour(tv, *um, px, py_tex, rgb); // opaque texture (alpha-tested ones never come here)
			const uint32_t zl = depth_voxel_range(rp.res, dfma(dzdx, dx, row_term), zr_lo, zr_hi) - oz;
			const uint32_t m_lo = xs | (s_lut[zl & 1023u] << shz) | row_lo;
			uint32_t m_hi = 0;
			if (deep) m_hi = hi_xrow | (s_lut[(zl >> 10) & 1023u] << shz);
			out[k] = ((uint64_t)((m_lo >> 8) | (m_hi << 22)) << 32) | (uint64_t)((m_lo << 24) | rgb);
			xs = ((xs | ~xmask) + xone) & xmask; // x + 1 in spread form: the gaps are filled so that the carry ripples
			++xl, ++px;
			dx += 256.0;
			if (deep && xs == 0u) hi_xrow = (s_lut[(xl >> 10) & 1023u] << shx) | row_hi; // x crossed a multiple of 1024
		} else
			out[k] = 0;
	}
	uint64_t *dst = frags + j0;
	if (j0 + EMIT_ITEMS <= c1 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
		for (int k = 0; k < EMIT_ITEMS; k += 2) *reinterpret_cast<ulonglong2 *>(dst + k) = make_ulonglong2(out[k], out[k + 1]);
	} else {
#pragma unroll
		for (int k = 0; k < EMIT_ITEMS; ++k)
			if (j0 + k < c1) dst[k] = out[k];
	}
}

// Rows of large alpha-tested triangles: one thread per dense row walks the row's span, samples, and writes the
// surviving fragments at the row's offset (the count pass counted them the same way).  Runs after k_emit_large,
// whose span arithmetic does not apply to these rows (whatever it put at their ordinals is overwritten here).
template <bool TEX>
__global__ void __launch_bounds__(RASTER_BLOCK)
    k_emit_alpha_rows(RasterParams rp, TexView tv, const LargeTri *__restrict__ large, const UvMap *__restrict__ luv, DenseRows rows,
                      uint32_t n_rows, uint64_t *__restrict__ frags /* already offset to the large region */) {
	__shared__ uint32_t s_lut[1024];
	fill_spread_table(s_lut, RASTER_BLOCK);
	__syncthreads();
	const uint32_t r = blockIdx.x * RASTER_BLOCK + threadIdx.x;
	if (!TEX || r >= n_rows) return;
	const uint32_t li = rows.li[r];
	const LargeTri &lt = large[li];
	if (!(lt.textured & 2u)) return;
	const int32_t py = (int32_t)(rows.xy[r] >> 16);
	int32_t x_lo, x_hi;
	row_span(lt.ts, py, x_lo, x_hi);
	row_span_depth_window(lt.ts, rp.res, py, x_lo, x_hi);
	uint64_t o = rows.off[r];
	for (int32_t px = x_lo; px <= x_hi; ++px) {
		uint32_t rgb, uz = 0;
		if (!sample_colour(tv, luv[li], px, py, rgb)) continue;
		pixel_fragment(lt.ts, rp.res, px, py, uz); // inside the depth window by construction of the span
		frags[o++] = make_fragment_lut(lt.ts, rp, s_lut, px, py, uz, rgb);
	}
}

// voxelizer.frag:40-42 packing of our fragments (levels <= 12), for consumers of the reference format
__global__ void __launch_bounds__(256)
    k_export_reference_fragments(const uint64_t *__restrict__ frags, uint64_t n, uint2 *__restrict__ out) {
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const uint64_t f = frags[i];
	const uint64_t m = f >> 24;
	const uint32_t x = compact1by2(m), y = compact1by2(m >> 1), z = compact1by2(m >> 2);
	out[i] = make_uint2(x | (y << 12) | ((z & 0xffu) << 24), ((z >> 8) << 28) | (uint32_t)(f & 0xffffffu));
}

} // namespace svo
