// svo_math.cuh -- per-triangle / per-fragment arithmetic of the voxelizer and the leaf colour rule,
// written once as __host__ __device__ so that the same code is compiled into the sm_100a kernels and
// into the CPU kernel-logic emulator used by the CPU-only tests (tests/cpu_emu; never a product path).
//
// This is the "pinned arithmetic" of DESIGN.md section 3: one IEEE operation per written operator
// (explicit _rn intrinsics on the device, -ffp-contract=off on the host), fp32 where the reference
// shaders are fp32, exact integers for coverage, fp64 for the depth plane.
#pragma once
#include "common.cuh"

namespace svo {

// ---- no-contraction fp helpers -------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
SVO_HD inline float fmul(float a, float b) { return __fmul_rn(a, b); }
SVO_HD inline float fadd(float a, float b) { return __fadd_rn(a, b); }
SVO_HD inline float fsub(float a, float b) { return __fsub_rn(a, b); }
SVO_HD inline double dmul(double a, double b) { return __dmul_rn(a, b); }
SVO_HD inline double dsub(double a, double b) { return __dsub_rn(a, b); }
SVO_HD inline double dadd(double a, double b) { return __dadd_rn(a, b); }
SVO_HD inline double ddiv(double a, double b) { return __ddiv_rn(a, b); }
SVO_HD inline double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
SVO_HD inline float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SVO_HD inline float fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
SVO_HD inline float fmul(float a, float b) { return a * b; }
SVO_HD inline float fadd(float a, float b) { return a + b; }
SVO_HD inline float fsub(float a, float b) { return a - b; }
SVO_HD inline double dmul(double a, double b) { return a * b; }
SVO_HD inline double dsub(double a, double b) { return a - b; }
SVO_HD inline double dadd(double a, double b) { return a + b; }
SVO_HD inline double ddiv(double a, double b) { return a / b; }
SVO_HD inline double dfma(double a, double b, double c) { return fma(a, b, c); }
SVO_HD inline float ffma(float a, float b, float c) { return fmaf(a, b, c); }
SVO_HD inline float fdiv(float a, float b) { return a / b; }
#endif

// GLSL uint(float): truncation; negative / NaN pinned to 0, too large saturates.
SVO_HD inline uint32_t f2u_sat(float f) {
	if (!(f > 0.0f)) return 0u;
	if (f >= 4294967296.0f) return 0xffffffffu;
	return (uint32_t)f;
}
SVO_HD inline float glsl_min(float x, float y) { return y < x ? y : x; }
SVO_HD inline float glsl_max(float x, float y) { return x < y ? y : x; }

// a[i] for i in 0..2 by selects: indexing a kernel-parameter array with a run-time index makes the compiler test
// every possible constant-bank slot in turn (dozens of instructions)
SVO_HD inline uint32_t pick3(const uint32_t (&a)[3], uint32_t i) { return i == 0u ? a[0] : (i == 1u ? a[1] : a[2]); }
template <class T> SVO_HD inline T tmin(T a, T b) { return b < a ? b : a; }
template <class T> SVO_HD inline T tmax(T a, T b) { return a < b ? b : a; }
SVO_HD inline int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }

// floor(n / d) for d > 0, exact
SVO_HD inline int64_t floor_div(int64_t n, int64_t d) {
	int64_t q = n / d;
	if ((n % d != 0) && (n < 0)) --q;
	return q;
}
SVO_HD inline int64_t ceil_div(int64_t n, int64_t d) { return -floor_div(-n, d); }

// ---- Morton --------------------------------------------------------------------------------------
// spread the low 10 bits of v to every third bit
SVO_HD inline uint32_t part1by2_10(uint32_t v) {
	v &= 0x3ffu;
	v = (v | (v << 16)) & 0x030000ffu;
	v = (v | (v << 8)) & 0x0300f00fu;
	v = (v | (v << 4)) & 0x030c30c3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}
// spread up to 20 bits (levels <= 20) : low 30 Morton bits from the low 10 coordinate bits, rest above
SVO_HD inline uint64_t part1by2(uint32_t v) {
	return (uint64_t)part1by2_10(v) | ((uint64_t)part1by2_10(v >> 10) << 30);
}
// child slot = x | y<<1 | z<<2 per level, most significant level first (octree_tag_node.comp:24-25)
SVO_HD inline uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
	return part1by2(x) | (part1by2(y) << 1) | (part1by2(z) << 2);
}
// x -> x + 1 directly in spread form (bits at positions 0, 3, 6, ...): fill the gaps with ones so the carry ripples
SVO_HD inline uint64_t part1by2_increment(uint64_t m) {
	const uint64_t mask = 0x1249249249249249ull;
	return ((m | ~mask) + 1ull) & mask;
}
SVO_HD inline uint32_t compact1by2_10(uint32_t v) {
	v &= 0x09249249u;
	v = (v | (v >> 2)) & 0x030c30c3u;
	v = (v | (v >> 4)) & 0x0300f00fu;
	v = (v | (v >> 8)) & 0x030000ffu;
	v = (v | (v >> 16)) & 0x3ffu;
	return v;
}
SVO_HD inline uint32_t compact1by2(uint64_t m) {
	return compact1by2_10((uint32_t)(m & 0x3fffffffu)) | (compact1by2_10((uint32_t)((m >> 30) & 0x3fffffffu)) << 10);
}

// ---- triangle setup ------------------------------------------------------------------------------
// voxelizer.vert:8-11 + voxelizer.geom:15-42 + the viewport transform configured at Voxelizer.cpp:115-116.
struct TriSetup {
	// coverage: pixel (px,py) is covered iff ea[i]*px + eb[i]*py + ec[i] >= 0 for i = 0..2, inside the
	// pixel rectangle [px0,px1] x [py0,py1] (already intersected with gAABB, the viewport and the shard).
	int64_t ea[3], eb[3], ec[3];
	int32_t px0, px1, py0, py1; // empty when px0 > px1 or py0 > py1
	// depth plane through the snapped vertices, fp64
	double dzdx, dzdy, z0;
	int32_t X0, Y0;
	uint32_t zr_lo, zr_hi; // gDepthRange (voxelizer.geom:41-42)
	uint32_t axis;         // gAxis
	// size of the candidate rectangle BEFORE the shard window is applied: the small/large work
	// classification uses it so that a triangle is classified identically in every shard
	int64_t full_area;
	// shard window along the depth axis (fragments outside are dropped); cull_depth = window is a strict subset
	uint32_t zs_lo, zs_hi;
	bool cull_depth;
	// Mode B only: the dilated vertices' extrapolated depth can leave [0,1]; such fragments are depth-clipped
	bool clip_z;
};

enum { MODE_CENTER = 0, MODE_CONSERVATIVE = 1, MODE_DILATE = 2 };

// voxelizer_conservative.geom:46-87 -- the software-conservative dilation the reference falls back to without
// VK_EXT_conservative_rasterization (Voxelizer.cpp:92-99).  One fp32 rounding per operator, with fused
// multiply-adds exactly where the reference's compiled SPIR-V has them (GetBarycentric's numerators and
// denominator); bit-identical to the executed binary (tests/test_spirv_golden.py).  q = projected vertices
// (ndc x, ndc y, depth); replaced by the three emitted vertices.
SVO_HD inline void dilate_mode_b(float (&qx)[3], float (&qy)[3], float (&qz)[3], float normal_axis, uint32_t res) {
	const int LA[3] = {2, 0, 1}, LB[3] = {1, 2, 0}; // line0 = cross(ndc2, ndc1), line1 = cross(ndc0, ndc2), line2 = cross(ndc1, ndc0)
	const float inv = fdiv(1.0f, (float)res);
	float lx[3], ly[3], lz[3];
	for (int i = 0; i < 3; ++i) {
		const float ax = qx[LA[i]], ay = qy[LA[i]], bx = qx[LB[i]], by = qy[LB[i]];
		lx[i] = fsub(ay, by);
		ly[i] = fsub(bx, ax);
		lz[i] = fsub(fmul(ax, by), fmul(ay, bx));
		const float d = fadd(fmul(inv, fabsf(lx[i])), fmul(inv, fabsf(ly[i])));
		lz[i] = normal_axis < 0.0f ? fadd(lz[i], d) : fsub(lz[i], d);
	}
	const int IA[3] = {2, 0, 1}, IB[3] = {1, 2, 0}; // intersect0 = cross(line2, line1), 1 = cross(line0, line2), 2 = cross(line1, line0)
	const float ax = qx[0], ay = qy[0], bx = qx[1], by = qy[1], cx = qx[2], cy = qy[2];
	const float den = ffma(fsub(by, cy), fsub(ax, cx), fmul(fsub(cx, bx), fsub(ay, cy)));
	float ox[3], oy[3], oz[3];
	for (int i = 0; i < 3; ++i) {
		const int u = IA[i], v = IB[i];
		const float ix = fsub(fmul(ly[u], lz[v]), fmul(lz[u], ly[v]));
		const float iy = fsub(fmul(lz[u], lx[v]), fmul(lx[u], lz[v]));
		const float iz = fsub(fmul(lx[u], ly[v]), fmul(ly[u], lx[v]));
		const float px = fdiv(ix, iz), py = fdiv(iy, iz);
		const float l0 = fdiv(ffma(fsub(by, cy), fsub(px, cx), fmul(fsub(cx, bx), fsub(py, cy))), den);
		const float l1 = fdiv(ffma(fsub(cy, ay), fsub(px, cx), fmul(fsub(ax, cx), fsub(py, cy))), den);
		const float l2 = fsub(fsub(1.0f, l0), l1);
		ox[i] = px, oy[i] = py;
		oz[i] = fadd(fadd(fmul(l0, qz[0]), fmul(l1, qz[1])), fmul(l2, qz[2]));
	}
	for (int i = 0; i < 3; ++i) qx[i] = ox[i], qy[i] = oy[i], qz[i] = oz[i];
}

// Shard window in voxel coordinates (half-open), used to clip the pixel rectangle early and to cull
// fragments by depth; whole grid: lo = 0, hi = res.
struct ShardBox {
	uint32_t lo[3], hi[3];
};

// returns false when the triangle cannot produce a fragment
SVO_HD inline bool tri_setup(const float *p0, const float *p1, const float *p2, uint32_t res, int mode, const ShardBox &sb,
                             TriSetup &t) {
	const float *p[3] = {p0, p1, p2};
	bool valid = true;
	for (int i = 0; i < 3; ++i)
		for (int k = 0; k < 3; ++k)
			if (!(fabsf(p[i][k]) <= 2.0f)) valid = false; // guard band, rejects NaN / Inf
	float e1[3], e2[3];
	for (int k = 0; k < 3; ++k) {
		e1[k] = fsub(p1[k], p0[k]);
		e2[k] = fsub(p2[k], p0[k]);
	}
	// voxelizer.geom:28-32
	const float nsx = fsub(fmul(e1[1], e2[2]), fmul(e1[2], e2[1]));
	const float nsy = fsub(fmul(e1[2], e2[0]), fmul(e1[0], e2[2]));
	const float nsz = fsub(fmul(e1[0], e2[1]), fmul(e1[1], e2[0]));
	const float nx = fabsf(nsx), ny = fabsf(nsy), nz = fabsf(nsz);
	uint32_t axis = (nx > ny && nx > nz) ? 0u : ((ny > nz) ? 1u : 2u);
	t.axis = axis;
	if (!valid) return false;

	// Project (voxelizer.geom:15-19): axis 0 -> v.yzx, 1 -> v.zxy, 2 -> v.xyz ; z = (z+1)*0.5
	const int sx = axis == 0u ? 1 : (axis == 1u ? 2 : 0);
	const int sy = axis == 0u ? 2 : (axis == 1u ? 0 : 1);
	const int sz = axis == 0u ? 0 : (axis == 1u ? 1 : 2);
	const float fres = (float)res;
	float qx[3], qy[3], qz[3];
	for (int i = 0; i < 3; ++i) { // (selects, not run-time indexing: the vertices stay in registers)
		qx[i] = sx == 0 ? p[i][0] : (sx == 1 ? p[i][1] : p[i][2]);
		qy[i] = sy == 0 ? p[i][0] : (sy == 1 ? p[i][1] : p[i][2]);
		qz[i] = fmul(fadd(sz == 0 ? p[i][0] : (sz == 1 ? p[i][1] : p[i][2]), 1.0f), 0.5f);
	}
	// gAABB / gDepthRange (voxelizer.geom:39-42)
	uint32_t ax0 = f2u_sat(fmul(fmul(fadd(glsl_min(qx[0], glsl_min(qx[1], qx[2])), 1.0f), 0.5f), fres));
	uint32_t ay0 = f2u_sat(fmul(fmul(fadd(glsl_min(qy[0], glsl_min(qy[1], qy[2])), 1.0f), 0.5f), fres));
	uint32_t ax1 = f2u_sat(fmul(fmul(fadd(glsl_max(qx[0], glsl_max(qx[1], qx[2])), 1.0f), 0.5f), fres));
	uint32_t ay1 = f2u_sat(fmul(fmul(fadd(glsl_max(qy[0], glsl_max(qy[1], qy[2])), 1.0f), 0.5f), fres));
	t.zr_lo = f2u_sat(fmul(glsl_min(qz[0], glsl_min(qz[1], qz[2])), fres));
	t.zr_hi = f2u_sat(fmul(glsl_max(qz[0], glsl_max(qz[1], qz[2])), fres));

	// Mode B: rasterize the dilated triangle instead (gAABB / gDepthRange stay those of the original one)
	const bool centre_rule = mode != MODE_CONSERVATIVE;
	t.clip_z = mode == MODE_DILATE;
	if (mode == MODE_DILATE) {
		dilate_mode_b(qx, qy, qz, axis == 0u ? nsx : (axis == 1u ? nsy : nsz), res);
		for (int i = 0; i < 3; ++i) // NaN / Inf / far outside: degenerate input, dropped
			if (!(fabsf(qx[i]) <= 4.0f) || !(fabsf(qy[i]) <= 4.0f) || !(fabsf(qz[i]) <= 1e6f)) return false;
	}

	// window coordinates, snapped to 1/256 pixel (round half even)
	int32_t X[3], Y[3];
	float zf[3];
	for (int i = 0; i < 3; ++i) {
		float xf = fmul(fmul(fadd(qx[i], 1.0f), 0.5f), fres);
		float yf = fmul(fmul(fadd(qy[i], 1.0f), 0.5f), fres);
		X[i] = (int32_t)rintf(fmul(xf, 256.0f));
		Y[i] = (int32_t)rintf(fmul(yf, 256.0f));
		zf[i] = qz[i];
	}
	int64_t a2 = (int64_t)(X[1] - X[0]) * (int64_t)(Y[2] - Y[0]) - (int64_t)(X[2] - X[0]) * (int64_t)(Y[1] - Y[0]);
	if (a2 < 0) { // CULL_NONE (Voxelizer.cpp:117-118): normalise the winding
		int32_t ti;
		float tf;
		ti = X[1], X[1] = X[2], X[2] = ti;
		ti = Y[1], Y[1] = Y[2], Y[2] = ti;
		tf = zf[1], zf[1] = zf[2], zf[2] = tf;
		a2 = -a2;
	}
	if (a2 == 0 && centre_rule) return false;

	const int32_t xmin = tmin(X[0], tmin(X[1], X[2])), xmax = tmax(X[0], tmax(X[1], X[2]));
	const int32_t ymin = tmin(Y[0], tmin(Y[1], Y[2])), ymax = tmax(Y[0], tmax(Y[1], Y[2]));

	// edges: E_i(P) = (Xb-Xa)*(Py-Ya) - (Yb-Ya)*(Px-Xa) = A*Px + B*Py + C ; interior E_i > 0 when a2 > 0
	if (a2 > 0) {
		const int EA[3] = {1, 2, 0}, EB[3] = {2, 0, 1};
		for (int i = 0; i < 3; ++i) {
			int64_t A = -(int64_t)(Y[EB[i]] - Y[EA[i]]), B = (int64_t)(X[EB[i]] - X[EA[i]]);
			int64_t Cc = -(A * (int64_t)X[EA[i]] + B * (int64_t)Y[EA[i]]);
			int64_t slack;
			if (centre_rule) {
				bool top_left = (A > 0) || (A == 0 && B > 0);
				slack = top_left ? 0 : -1; // E > 0  <=>  E - 1 >= 0
			} else
				slack = 128 * (iabs64(A) + iabs64(B)); // max of E over the pixel square >= 0
			// centre of pixel (px,py) = (256px+128, 256py+128)
			t.ea[i] = 256 * A;
			t.eb[i] = 256 * B;
			t.ec[i] = Cc + 128 * (A + B) + slack;
		}
	} else {
		// zero snapped area (conservative only): the longest edge as a segment, |E| <= support
		// (selects instead of run-time indices, so that X / Y stay in registers)
		const int SA[3] = {0, 1, 2}, SB[3] = {1, 2, 0};
		int64_t best_d = -1, A = 0, B = 0, Cc = 0;
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			const int64_t dx = X[SB[i]] - X[SA[i]], dy = Y[SB[i]] - Y[SA[i]];
			const int64_t d = dx * dx + dy * dy;
			if (d > best_d) {
				best_d = d;
				A = -dy, B = dx;
				Cc = -(A * (int64_t)X[SA[i]] + B * (int64_t)Y[SA[i]]);
			}
		}
		int64_t sup = 128 * (iabs64(A) + iabs64(B));
		t.ea[0] = 256 * A, t.eb[0] = 256 * B, t.ec[0] = Cc + 128 * (A + B) + sup;
		t.ea[1] = -256 * A, t.eb[1] = -256 * B, t.ec[1] = -(Cc + 128 * (A + B)) + sup;
		t.ea[2] = 0, t.eb[2] = 0, t.ec[2] = 0;
	}

	// candidate pixel rectangle
	int64_t px0, px1, py0, py1;
	if (centre_rule) { // centres inside the closed snapped bounding box
		px0 = ceil_div((int64_t)xmin - 128, 256), px1 = floor_div((int64_t)xmax - 128, 256);
		py0 = ceil_div((int64_t)ymin - 128, 256), py1 = floor_div((int64_t)ymax - 128, 256);
	} else { // closed squares [256p, 256p+256] touching the closed bounding box
		px0 = ceil_div((int64_t)xmin, 256) - 1, px1 = floor_div((int64_t)xmax, 256);
		py0 = ceil_div((int64_t)ymin, 256) - 1, py1 = floor_div((int64_t)ymax, 256);
	}
	// viewport / scissor (Voxelizer.cpp:115-116), gAABB discard (voxelizer.frag:21-22)
	px0 = tmax(px0, (int64_t)0), py0 = tmax(py0, (int64_t)0);
	px1 = tmin(px1, (int64_t)res - 1), py1 = tmin(py1, (int64_t)res - 1);
	px0 = tmax(px0, (int64_t)ax0), py0 = tmax(py0, (int64_t)ay0);
	px1 = tmin(px1, (int64_t)ax1), py1 = tmin(py1, (int64_t)ay1);
	t.full_area = (px0 > px1 || py0 > py1) ? 0 : (px1 - px0 + 1) * (py1 - py0 + 1);
	// shard window on the two screen axes: voxel = axis 0: (uz, px, py); 1: (py, uz, px); 2: (px, py, uz)
	const int wx = axis == 0u ? 1 : (axis == 1u ? 2 : 0); // world axis of screen x
	const int wy = axis == 0u ? 2 : (axis == 1u ? 0 : 1);
	px0 = tmax(px0, (int64_t)pick3(sb.lo, wx)), px1 = tmin(px1, (int64_t)pick3(sb.hi, wx) - 1);
	py0 = tmax(py0, (int64_t)pick3(sb.lo, wy)), py1 = tmin(py1, (int64_t)pick3(sb.hi, wy) - 1);
	t.px0 = (int32_t)px0, t.px1 = (int32_t)px1, t.py0 = (int32_t)py0, t.py1 = (int32_t)py1;
	if (px0 > px1 || py0 > py1) return false;
	// depth range vs shard along the depth axis (fragments are clamped into [zr_lo, zr_hi])
	{
		uint32_t zlo = tmin(t.zr_lo, res - 1u), zhi = tmin(t.zr_hi, res - 1u);
		const uint32_t wlo = pick3(sb.lo, (uint32_t)sz), whi = pick3(sb.hi, (uint32_t)sz);
		if (zhi < wlo || zlo >= whi) return false;
		t.zs_lo = wlo, t.zs_hi = whi;
		t.cull_depth = zlo < wlo || zhi >= whi;
	}

	// depth plane (fp64)
	t.X0 = X[0], t.Y0 = Y[0];
	t.z0 = (double)zf[0];
	if (a2 == 0) {
		t.dzdx = 0.0, t.dzdy = 0.0; // provoking-vertex depth for degenerate primitives
	} else {
		double dz1 = dsub((double)zf[1], (double)zf[0]), dz2 = dsub((double)zf[2], (double)zf[0]);
		double dx1 = (double)(X[1] - X[0]), dy1 = (double)(Y[1] - Y[0]);
		double dx2 = (double)(X[2] - X[0]), dy2 = (double)(Y[2] - Y[0]);
		double da2 = (double)a2;
		t.dzdx = ddiv(dsub(dmul(dz1, dy2), dmul(dz2, dy1)), da2);
		t.dzdy = ddiv(dsub(dmul(dz2, dx1), dmul(dz1, dx2)), da2);
	}
	return true;
}

SVO_HD inline bool pixel_covered(const TriSetup &t, int32_t px, int32_t py) {
	bool in = true;
#pragma unroll
	for (int i = 0; i < 3; ++i) in = in && (t.ea[i] * (int64_t)px + t.eb[i] * (int64_t)py + t.ec[i] >= 0);
	return in;
}

// Row span [x_lo, x_hi] of covered pixels of row py inside [px0, px1]; empty when x_lo > x_hi.
// Exact: solves ea*px + (eb*py + ec) >= 0 for px with integer floor/ceil division.
SVO_HD inline void row_span(const TriSetup &t, int32_t py, int32_t &x_lo, int32_t &x_hi) {
	int64_t lo = t.px0, hi = t.px1;
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		const int64_t r = t.eb[i] * (int64_t)py + t.ec[i];
		const int64_t a = t.ea[i];
		if (a > 0) {
			lo = tmax(lo, ceil_div(-r, a));
		} else if (a < 0) {
			hi = tmin(hi, floor_div(r, -a));
		} else if (r < 0) {
			hi = lo - 1;
		}
	}
	if (hi < lo) hi = lo - 1;
	x_lo = (int32_t)lo, x_hi = (int32_t)hi;
}

// depth voxel of pixel (px,py): voxelizer.frag:18-20,23 (+ the pinned final clamp to res-1).
// Split into the row term and the pixel term so that span kernels can hoist the row term: the operations
// (and therefore the rounding) are identical whichever way it is called.
SVO_HD inline double depth_row_term(const TriSetup &t, int32_t py) {
	const int32_t cy = py * 256 + 128;
	return dfma(t.dzdy, (double)(cy - t.Y0), t.z0);
}
SVO_HD inline double plane_z_row(const TriSetup &t, int32_t px, double row_term) {
	const int32_t cx = px * 256 + 128;
	return dfma(t.dzdx, (double)(cx - t.X0), row_term);
}
SVO_HD inline uint32_t depth_voxel_range(uint32_t res, double z, uint32_t zr_lo, uint32_t zr_hi) {
	double zs = dmul(z, (double)res);
#if defined(__CUDA_ARCH__)
	// the conversion saturates (negative and NaN -> 0, huge -> 2^32-1), the last min() below does the rest
	uint32_t uz = __double2uint_rz(zs);
#else
	uint32_t uz = !(zs > 0.0) ? 0u : (zs >= (double)res ? res - 1u : (uint32_t)zs);
#endif
	uz = tmax(uz, zr_lo);
	uz = tmin(uz, zr_hi);
	uz = tmin(uz, res - 1u);
	return uz;
}
SVO_HD inline uint32_t depth_voxel(const TriSetup &t, uint32_t res, double z) { return depth_voxel_range(res, z, t.zr_lo, t.zr_hi); }
SVO_HD inline uint32_t pixel_depth_row(const TriSetup &t, uint32_t res, int32_t px, double row_term) {
	return depth_voxel(t, res, plane_z_row(t, px, row_term));
}
SVO_HD inline uint32_t pixel_depth(const TriSetup &t, uint32_t res, int32_t px, int32_t py) {
	return pixel_depth_row(t, res, px, depth_row_term(t, py));
}
// Does the covered pixel (px,py) produce a fragment, and at which depth voxel?  Drops fragments that are depth
// clipped (Mode B only: depthClampEnable = 0, dep/MyVK/src/GraphicsPipeline.cpp:45-50) or fall outside the
// shard's depth window.
SVO_HD inline bool pixel_fragment(const TriSetup &t, uint32_t res, int32_t px, int32_t py, uint32_t &uz) {
	const double z = plane_z_row(t, px, depth_row_term(t, py));
	if (t.clip_z && !(z >= 0.0 && z <= 1.0)) return false;
	uz = depth_voxel(t, res, z);
	return !t.cull_depth || (uz >= t.zs_lo && uz < t.zs_hi);
}

// Narrow a row span to the pixels that survive the depth clip / the shard's depth window.  Depth is a monotone
// function of px along a row (one fma with a fixed slope, then monotone clamps), so the survivors form one
// interval whose ends are found by bisection with the exact per-pixel depth.
SVO_HD inline void row_span_depth_window(const TriSetup &t, uint32_t res, int32_t py, int32_t &x_lo, int32_t &x_hi) {
	if ((!t.cull_depth && !t.clip_z) || x_lo > x_hi) return;
	const bool rising = !(t.dzdx < 0.0);
	const double row_term = depth_row_term(t, py);
	// "too low" holds on a prefix (rising) or suffix (falling) of the row; "too high" the other way round
	auto below = [&](int32_t px) {
		const double z = plane_z_row(t, px, row_term);
		return (t.clip_z && z < 0.0) || (t.cull_depth && depth_voxel(t, res, z) < t.zs_lo);
	};
	auto above = [&](int32_t px) {
		const double z = plane_z_row(t, px, row_term);
		return (t.clip_z && z > 1.0) || (t.cull_depth && depth_voxel(t, res, z) >= t.zs_hi);
	};
	int32_t lo = x_lo, hi = x_hi;
	if (rising) {
		int32_t a = lo, b = hi + 1; // first px that is not below
		while (a < b) { int32_t m = a + (b - a) / 2; if (below(m)) a = m + 1; else b = m; }
		lo = a;
		a = lo, b = hi + 1; // last px that is not above
		while (a < b) { int32_t m = a + (b - a) / 2; if (!above(m)) a = m + 1; else b = m; }
		hi = a - 1;
	} else {
		int32_t a = lo, b = hi + 1;
		while (a < b) { int32_t m = a + (b - a) / 2; if (above(m)) a = m + 1; else b = m; }
		lo = a;
		a = lo, b = hi + 1;
		while (a < b) { int32_t m = a + (b - a) / 2; if (!below(m)) a = m + 1; else b = m; }
		hi = a - 1;
	}
	x_lo = lo, x_hi = hi;
}

// un-swizzle (voxelizer.frag:24): axis 0 -> u.zxy, 1 -> u.yzx, 2 -> u.xyz
SVO_HD inline void unswizzle(uint32_t axis, uint32_t ux, uint32_t uy, uint32_t uz, uint32_t &vx, uint32_t &vy, uint32_t &vz) {
	if (axis == 0u)
		vx = uz, vy = ux, vz = uy;
	else if (axis == 1u)
		vx = uy, vy = uz, vz = ux;
	else
		vx = ux, vy = uy, vz = uz;
}

// ---- leaf colour: octree_tag_node.comp:10-16,48-57 ------------------------------------------------
// first writer: 0xC1000000 | rgb ; later writers: running integer average, count saturating at 63
SVO_HD inline uint32_t leaf_first(uint32_t rgb) { return 0xC1000000u | (rgb & 0xffffffu); }
// floor(x / d) for 0 <= x <= 255*63+255, 1 <= d <= 64.  On the device: (x + 0.5) * approx(1/d), truncated.  x and d are
// integers, so (x + 0.5)/d is at least 1/128 away from every integer while the reciprocal approximation and the two
// roundings are off by less than 2^-20 relative (< 0.02 absolute): the truncation is exact, at a fifth of the cost
// of an integer division.  (The leaf fold is the hot loop of scenes with many fragments per voxel.)
SVO_HD inline uint32_t div_small(uint32_t x, uint32_t d, float rcp_d) {
#if defined(__CUDA_ARCH__)
	(void)d;
	return (uint32_t)__fmul_rn(__fadd_rn((float)x, 0.5f), rcp_d);
#else
	(void)rcp_d;
	return x / d;
#endif
}
SVO_HD inline uint32_t leaf_accumulate(uint32_t prev, uint32_t rgb) {
	uint32_t w = (prev >> 24) & 0x3fu;
#if defined(__CUDA_ARCH__)
	const float rcp = __frcp_rn((float)(w + 1u));
#else
	const float rcp = 0.0f;
#endif
	uint32_t r = div_small((prev & 0xffu) * w + (rgb & 0xffu), w + 1u, rcp);
	uint32_t g = div_small(((prev >> 8) & 0xffu) * w + ((rgb >> 8) & 0xffu), w + 1u, rcp);
	uint32_t b = div_small(((prev >> 16) & 0xffu) * w + ((rgb >> 16) & 0xffu), w + 1u, rcp);
	uint32_t nw = tmin(w + 1u, 0x3fu);
	return (nw << 24) | (r & 0xffu) | ((g & 0xffu) << 8) | ((b & 0xffu) << 16) | 0xC0000000u;
}

} // namespace svo
