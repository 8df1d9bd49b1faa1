/*
 * oracle/svo_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see svo_oracle.h).
 *
 * PARITY STATUS: see svo_oracle.h -- pinned against the reference's own SPIR-V binaries executed
 * by oracle/spirv_interp.py (tests/golden/spirv_*.npz) for everything the shaders define; the
 * fixed-function rasterizer stage is unpinned (driver-defined) and uses DESIGN.md section 3.
 * Every function cites the reference file:line (relative to /root/reference) that it restates.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: the pinned arithmetic below is "one IEEE operation per
 * written operator", fp32 where the shaders are fp32.
 */
#include "svo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Fragment packing: voxelizer.frag:40-42 / octree_tag_node.comp:38-39
 * ---------------------------------------------------------------------------------------- */
void orc_pack_fragment(uint32_t x, uint32_t y, uint32_t z, uint32_t colour, uint32_t out[2]) {
	out[0] = x | (y << 12u) | ((z & 0xffu) << 24u);
	out[1] = ((z >> 8u) << 28u) | (colour & 0x00ffffffu);
}
void orc_unpack_fragment(const uint32_t in[2], uint32_t *x, uint32_t *y, uint32_t *z, uint32_t *rgb) {
	*x = in[0] & 0xfffu;
	*y = (in[0] >> 12u) & 0xfffu;
	*z = (in[0] >> 24u) | ((in[1] >> 28u) << 8u);
	*rgb = in[1] & 0xffffffu;
}

/* OctreeBuilder.cpp:42-45 with Config.hpp:20-21 (u32 arithmetic, integer level/3). */
uint32_t orc_octree_entry_num(uint32_t fragment_count, uint32_t level) {
	const uint32_t kOctreeNodeNumMin = 1000000u, kOctreeNodeNumMax = 500000000u;
	uint32_t ratio = level / 3u;
	uint32_t n = fragment_count * ratio;
	if (n < kOctreeNodeNumMin)
		n = kOctreeNodeNumMin;
	if (n > kOctreeNodeNumMax)
		n = kOctreeNodeNumMax;
	return n;
}

uint64_t orc_morton(uint32_t x, uint32_t y, uint32_t z, uint32_t level) {
	uint64_t m = 0;
	for (uint32_t b = 0; b < level; ++b) {
		m |= (uint64_t)((x >> b) & 1u) << (3u * b);
		m |= (uint64_t)((y >> b) & 1u) << (3u * b + 1u);
		m |= (uint64_t)((z >> b) & 1u) << (3u * b + 2u);
	}
	return m;
}

/* ------------------------------------------------------------------------------------------
 * Voxelizer
 * ---------------------------------------------------------------------------------------- */

/* GLSL uint(float): truncation toward zero; negative / NaN are undefined in GLSL and are
 * pinned to 0 here, too-large saturates. */
static uint32_t f2u(float f) {
	if (!(f > 0.0f))
		return 0u;
	if (f >= 4294967296.0f)
		return 0xffffffffu;
	return (uint32_t)f;
}
static float fmin3(float a, float b, float c) {
	float m = c < b ? c : b; /* GLSL min(x,y) = y < x ? y : x */
	return m < a ? m : a;
}
static float fmax3(float a, float b, float c) {
	float m = b < c ? c : b;
	return a < m ? m : a;
}

typedef struct {
	uint32_t axis;
	int32_t X[3], Y[3]; /* snapped window coordinates, 1/256 pixel */
	float zf[3];        /* depth in [0,1] */
	uint32_t aabb[4];   /* gAABB  (voxelizer.geom:39-40) */
	uint32_t zr[2];     /* gDepthRange (voxelizer.geom:41-42) */
	int64_t area2;      /* > 0 after orientation normalisation; 0 = degenerate */
	int valid;
	int dilated;        /* Mode B: vertices come from voxelizer_conservative.geom; depth clip applies */
} orc_tri;

/* voxelizer_conservative.geom:46-87 -- the software-conservative dilation used when
 * VK_EXT_conservative_rasterization is absent (Voxelizer.cpp:92-99).  q[i] = projected vertex
 * (ndc x, ndc y, depth); normal_axis = normal[axis].  One fp32 rounding per operator, with the three
 * fused multiply-adds exactly where the reference's compiled SPIR-V has them (GetBarycentric's
 * numerators and denominator; checked against the executed binary, tests/test_spirv_golden.py).
 * The dilated vertices (emit order) replace q. */
static void dilate_mode_b(float q[3][3], float normal_axis, uint32_t res) {
	float line[3][3];
	const int LA[3] = {2, 0, 1}, LB[3] = {1, 2, 0}; /* line0 = cross(ndc2, ndc1), line1 = cross(ndc0, ndc2), line2 = cross(ndc1, ndc0) */
	const float inv = 1.0f / (float)res;
	for (int i = 0; i < 3; ++i) {
		const float ax = q[LA[i]][0], ay = q[LA[i]][1], bx = q[LB[i]][0], by = q[LB[i]][1];
		line[i][0] = ay * 1.0f - 1.0f * by;
		line[i][1] = 1.0f * bx - ax * 1.0f;
		line[i][2] = ax * by - ay * bx;
		const float d = inv * fabsf(line[i][0]) + inv * fabsf(line[i][1]);
		if (normal_axis < 0.0f) line[i][2] = line[i][2] + d; else line[i][2] = line[i][2] - d;
	}
	/* intersect0 = cross(line2, line1), intersect1 = cross(line0, line2), intersect2 = cross(line1, line0) */
	const int IA[3] = {2, 0, 1}, IB[3] = {1, 2, 0};
	float out[3][3];
	const float ax = q[0][0], ay = q[0][1], bx = q[1][0], by = q[1][1], cx = q[2][0], cy = q[2][1];
	const float den = fmaf(by - cy, ax - cx, (cx - bx) * (ay - cy));
	for (int i = 0; i < 3; ++i) {
		const float *u = line[IA[i]], *v = line[IB[i]];
		const float ix = u[1] * v[2] - u[2] * v[1], iy = u[2] * v[0] - u[0] * v[2], iz = u[0] * v[1] - u[1] * v[0];
		const float px = ix / iz, py = iy / iz;
		const float l0 = fmaf(by - cy, px - cx, (cx - bx) * (py - cy)) / den;
		const float l1 = fmaf(cy - ay, px - cx, (ax - cx) * (py - cy)) / den;
		const float l2 = (1.0f - l0) - l1;
		out[i][0] = px, out[i][1] = py;
		out[i][2] = (l0 * q[0][2] + l1 * q[1][2]) + l2 * q[2][2];
	}
	memcpy(q, out, sizeof(out));
}

/* voxelizer.vert:8-11 (pass-through) + voxelizer.geom:15-42 + viewport transform
 * (Voxelizer.cpp:115-116: viewport (0,0,res,res), depth 0..1, no y flip). */
static void orc_tri_setup(const float *p0, const float *p1, const float *p2, uint32_t res, int mode, orc_tri *t) {
	const float *p[3] = {p0, p1, p2};
	float e1[3], e2[3], n[3], w[3];
	for (int k = 0; k < 3; ++k) {
		e1[k] = p1[k] - p0[k];
		e2[k] = p2[k] - p0[k];
	}
	/* voxelizer.geom:28 cross(pos1 - pos0, pos2 - pos0) */
	n[0] = e1[1] * e2[2] - e1[2] * e2[1];
	n[1] = e1[2] * e2[0] - e1[0] * e2[2];
	n[2] = e1[0] * e2[1] - e1[1] * e2[0];
	for (int k = 0; k < 3; ++k)
		w[k] = fabsf(n[k]);
	/* voxelizer.geom:30-32 */
	t->axis = (w[0] > w[1] && w[0] > w[2]) ? 0u : ((w[1] > w[2]) ? 1u : 2u);

	float q[3][3];
	t->valid = 1;
	for (int i = 0; i < 3; ++i) {
		/* Project(): axis 0 -> v.yzx, axis 1 -> v.zxy, axis 2 -> v.xyz; z = (z+1)*0.5  (voxelizer.geom:15-19) */
		if (t->axis == 0u) {
			q[i][0] = p[i][1], q[i][1] = p[i][2], q[i][2] = p[i][0];
		} else if (t->axis == 1u) {
			q[i][0] = p[i][2], q[i][1] = p[i][0], q[i][2] = p[i][1];
		} else {
			q[i][0] = p[i][0], q[i][1] = p[i][1], q[i][2] = p[i][2];
		}
		q[i][2] = (q[i][2] + 1.0f) * 0.5f;
		t->zf[i] = q[i][2];
		for (int k = 0; k < 3; ++k)
			if (!(fabsf(p[i][k]) <= 2.0f)) /* guard band; also rejects NaN/Inf */
				t->valid = 0;
	}
	const float fres = (float)res;
	/* voxelizer.geom:39-42 */
	t->aabb[0] = f2u((fmin3(q[0][0], q[1][0], q[2][0]) + 1.0f) * 0.5f * fres);
	t->aabb[1] = f2u((fmin3(q[0][1], q[1][1], q[2][1]) + 1.0f) * 0.5f * fres);
	t->aabb[2] = f2u((fmax3(q[0][0], q[1][0], q[2][0]) + 1.0f) * 0.5f * fres);
	t->aabb[3] = f2u((fmax3(q[0][1], q[1][1], q[2][1]) + 1.0f) * 0.5f * fres);
	t->zr[0] = f2u(fmin3(q[0][2], q[1][2], q[2][2]) * fres);
	t->zr[1] = f2u(fmax3(q[0][2], q[1][2], q[2][2]) * fres);
	t->dilated = 0;
	if (!t->valid)
		return;
	if (mode == ORC_CONSERVATIVE_DILATE) { /* gAABB / gDepthRange above stay those of the ORIGINAL triangle (conservative.geom:69-74) */
		dilate_mode_b(q, n[t->axis], res);
		t->dilated = 1;
		for (int i = 0; i < 3; ++i) {
			t->zf[i] = q[i][2];
			if (!(fabsf(q[i][0]) <= 4.0f) || !(fabsf(q[i][1]) <= 4.0f) || !(fabsf(q[i][2]) <= 1e6f)) /* NaN / Inf / far outside: */
				t->valid = 0;                                                                    /* degenerate input, dropped */
		}
		if (!t->valid)
			return;
	}
	/* viewport transform x_f = (x+1)*res/2 (identical to the AABB expression) and snapping
	 * to 8 sub-pixel bits, round-half-even (pinned). */
	for (int i = 0; i < 3; ++i) {
		float xf = (q[i][0] + 1.0f) * 0.5f * fres;
		float yf = (q[i][1] + 1.0f) * 0.5f * fres;
		t->X[i] = (int32_t)rintf(xf * 256.0f);
		t->Y[i] = (int32_t)rintf(yf * 256.0f);
	}
	int64_t a2 = (int64_t)(t->X[1] - t->X[0]) * (int64_t)(t->Y[2] - t->Y[0]) -
	             (int64_t)(t->X[2] - t->X[0]) * (int64_t)(t->Y[1] - t->Y[0]);
	if (a2 < 0) { /* CULL_NONE: both windings rasterize (Voxelizer.cpp:117-118); normalise to a2 > 0 */
		int32_t ti;
		float tf;
		ti = t->X[1], t->X[1] = t->X[2], t->X[2] = ti;
		ti = t->Y[1], t->Y[1] = t->Y[2], t->Y[2] = ti;
		tf = t->zf[1], t->zf[1] = t->zf[2], t->zf[2] = tf;
		a2 = -a2;
	}
	t->area2 = a2;
}

static int64_t edge_fn(const orc_tri *t, int a, int b, int64_t px, int64_t py) {
	return (int64_t)(t->X[b] - t->X[a]) * (py - t->Y[a]) - (int64_t)(t->Y[b] - t->Y[a]) * (px - t->X[a]);
}
static int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }

/* Coverage of pixel (px,py) -- the rasterizer the reference configures at Voxelizer.cpp:112-129.
 * mode ORC_CENTER: centre sample + top-left rule (Vulkan spec, basic polygon rasterization).
 * mode ORC_CONSERVATIVE_EXACT: VK_CONSERVATIVE_RASTERIZATION_MODE_OVERESTIMATE with
 *   extraPrimitiveOverestimationSize 0 (Voxelizer.cpp:120-127): every pixel square that
 *   touches the (snapped) triangle, closed-set separating-axis test. */
static int covered(const orc_tri *t, int mode, int32_t px, int32_t py) {
	static const int EA[3] = {1, 2, 0}, EB[3] = {2, 0, 1};
	const int64_t cx = (int64_t)px * 256 + 128, cy = (int64_t)py * 256 + 128;
	if (mode != ORC_CONSERVATIVE_EXACT) { /* centre sample + top-left rule: plain (Mode C) or on the dilated triangle (Mode B) */
		if (t->area2 == 0)
			return 0;
		for (int i = 0; i < 3; ++i) {
			int64_t A = -(int64_t)(t->Y[EB[i]] - t->Y[EA[i]]), B = (int64_t)(t->X[EB[i]] - t->X[EA[i]]);
			int64_t e = edge_fn(t, EA[i], EB[i], cx, cy);
			int top_left = (A > 0) || (A == 0 && B > 0);
			if (e < 0 || (e == 0 && !top_left))
				return 0;
		}
		return 1;
	}
	/* conservative: box axes first */
	int32_t xmin = t->X[0], xmax = t->X[0], ymin = t->Y[0], ymax = t->Y[0];
	for (int i = 1; i < 3; ++i) {
		if (t->X[i] < xmin) xmin = t->X[i];
		if (t->X[i] > xmax) xmax = t->X[i];
		if (t->Y[i] < ymin) ymin = t->Y[i];
		if (t->Y[i] > ymax) ymax = t->Y[i];
	}
	if ((int64_t)px * 256 > xmax || (int64_t)px * 256 + 256 < xmin)
		return 0;
	if ((int64_t)py * 256 > ymax || (int64_t)py * 256 + 256 < ymin)
		return 0;
	if (t->area2 > 0) {
		for (int i = 0; i < 3; ++i) {
			int64_t A = -(int64_t)(t->Y[EB[i]] - t->Y[EA[i]]), B = (int64_t)(t->X[EB[i]] - t->X[EA[i]]);
			int64_t e = edge_fn(t, EA[i], EB[i], cx, cy);
			if (e + 128 * (iabs64(A) + iabs64(B)) < 0)
				return 0;
		}
		return 1;
	}
	/* zero snapped area: a segment or a point (degenerateTrianglesRasterized behaviour, pinned):
	 * the longest edge is the segment; squares touching it are covered. */
	static const int SA[3] = {0, 1, 2}, SB[3] = {1, 2, 0};
	int best = 0;
	int64_t best_d = -1;
	for (int i = 0; i < 3; ++i) {
		int64_t dx = t->X[SB[i]] - t->X[SA[i]], dy = t->Y[SB[i]] - t->Y[SA[i]];
		int64_t d = dx * dx + dy * dy;
		if (d > best_d)
			best_d = d, best = i;
	}
	int64_t A = -(int64_t)(t->Y[SB[best]] - t->Y[SA[best]]), B = (int64_t)(t->X[SB[best]] - t->X[SA[best]]);
	int64_t e = edge_fn(t, SA[best], SB[best], cx, cy);
	return iabs64(e) <= 128 * (iabs64(A) + iabs64(B));
}

/* Depth at the pixel centre: the triangle's plane through the snapped vertices evaluated in
 * fp64 (pinned); extrapolated when the centre is outside (conservative mode). */
typedef struct {
	double dzdx, dzdy, z0;
	int32_t X0, Y0;
} orc_plane;
static void plane_setup(const orc_tri *t, orc_plane *pl) {
	pl->X0 = t->X[0], pl->Y0 = t->Y[0];
	pl->z0 = (double)t->zf[0];
	if (t->area2 == 0) { /* provoking-vertex depth for degenerate primitives */
		pl->dzdx = pl->dzdy = 0.0;
		return;
	}
	double dz1 = (double)t->zf[1] - (double)t->zf[0], dz2 = (double)t->zf[2] - (double)t->zf[0];
	double dx1 = (double)(t->X[1] - t->X[0]), dy1 = (double)(t->Y[1] - t->Y[0]);
	double dx2 = (double)(t->X[2] - t->X[0]), dy2 = (double)(t->Y[2] - t->Y[0]);
	double a2 = (double)t->area2;
	pl->dzdx = (dz1 * dy2 - dz2 * dy1) / a2;
	pl->dzdy = (dz2 * dx1 - dz1 * dx2) / a2;
}

/* voxelizer.frag:17-25 GetVoxePos for pixel (px,py); returns 0 when the fragment is discarded. */
static int frag_voxel(const orc_tri *t, const orc_plane *pl, uint32_t res, int32_t px, int32_t py, uint32_t v[3]) {
	int32_t cx = px * 256 + 128, cy = py * 256 + 128;
	double z = fma(pl->dzdx, (double)(cx - pl->X0), fma(pl->dzdy, (double)(cy - pl->Y0), pl->z0));
	if (t->dilated && !(z >= 0.0 && z <= 1.0))
		return 0; /* depth clip (depthClampEnable = 0, dep/MyVK/src/GraphicsPipeline.cpp:45-50): the dilated vertices' extrapolated depth can leave [0,1] */
	double zs = z * (double)res; /* v.z *= float(kVoxelResolution) */
	uint32_t uz = !(zs > 0.0) ? 0u : (zs >= (double)res ? res - 1u : (uint32_t)zs); /* clamp(uvec3(v),0,res-1) */
	uint32_t ux = (uint32_t)px, uy = (uint32_t)py;
	/* voxelizer.frag:21-22 */
	if (ux < t->aabb[0] || ux > t->aabb[2] || uy < t->aabb[1] || uy > t->aabb[3])
		return 0;
	/* voxelizer.frag:23 clamp(u.z, lo, hi) = min(max(u.z, lo), hi) */
	if (uz < t->zr[0]) uz = t->zr[0];
	if (uz > t->zr[1]) uz = t->zr[1];
	/* pinned deviation: a depth of exactly 1.0 yields zr = res; the reference's traversal treats
	 * coordinate res like res-1 for L < 12 and corrupts the packing at L = 12. */
	if (uz > res - 1u) uz = res - 1u;
	/* voxelizer.frag:24 */
	if (t->axis == 0u)
		v[0] = uz, v[1] = ux, v[2] = uy;
	else if (t->axis == 1u)
		v[0] = uy, v[1] = uz, v[2] = ux;
	else
		v[0] = ux, v[1] = uy, v[2] = uz;
	return 1;
}

static int32_t floor_div256(int32_t v) { return v >> 8; } /* arithmetic shift = floor for negatives */

/* ------------------------------------------------------------------------------------------
 * Textures: voxelizer.frag:27-32 (texture(uTextures[id], gTexcoord), alpha < 0.5 discard,
 * packUnorm4x8), images VK_FORMAT_R8G8B8A8_SRGB with a full mip chain generated by linear blits
 * (Scene.cpp:262-296, dep/MyVK/src/CommandBuffer.cpp:338-393), sampler LINEAR / LINEAR mipmap /
 * REPEAT, no anisotropy, no LOD clamp (Scene.cpp:409-411, dep/MyVK/src/Sampler.cpp:13-38).
 *
 * UNPINNED like the rasterizer: filtering precision, the LOD approximation and sRGB conversion are
 * the Vulkan driver's.  The arithmetic below is the "ideal" formulas of the Vulkan spec (texel
 * filtering, scale factor / LOD operation, blit image) with pinned rounding:
 *  - sRGB decode through a 256-entry fp32 table; encode = nearest code in the sRGB domain, by
 *    comparing with the decoded mid-points; alpha is linear (a / 255).
 *  - texture coordinates: the affine map through the three ORIGINAL projected vertices (fp32 window
 *    coordinates taken to fp64), evaluated at the pixel centre in fp64 -- for Mode B too, where the
 *    reference extrapolates the same affine map to the dilated vertices (conservative.geom:21-27,76-86).
 *  - LOD: constant per triangle (the map is affine): rho^2 = max(|d(uv*size)/dx|^2, |d(uv*size)/dy|^2),
 *    lambda = log2(rho) floored to 1/256 with an exponent split and a 128-entry threshold table (no
 *    transcendental at sample time); lambda <= 0 = magnification (level 0).
 *  - bilinear weights from fp64 coordinates rounded to fp32, lerps in fp32, one rounding per operator.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	uint32_t w, h;
	uint8_t *px; /* RGBA8, sRGB-encoded colour, row-major */
} orc_level;
typedef struct {
	uint32_t levels;
	orc_level *lv;
} orc_tex;
struct orc_texset {
	uint32_t n;
	orc_tex *t;
	float decode[256];   /* sRGB code -> linear */
	float enc_thr[256];  /* linear value of the mid-point between codes c-1 and c */
	double lod_thr[128]; /* 2^(k/128) */
};

static double srgb_to_linear(double x) { return x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4); }

static uint32_t mip_level_count(uint32_t w, uint32_t h) { /* ImageBase::QueryMipLevel (dep/MyVK/include/myvk/ImageBase.hpp:10-17,49) */
	uint32_t x = w | h, ret = 1;
	if (x & 0x80000000u)
		return 32u;
	while (x >> ret)
		++ret;
	return ret;
}

static uint8_t encode_srgb(const orc_texset *s, float x) {
	uint32_t c = 0;
	for (uint32_t k = 1; k < 256; ++k)
		if (s->enc_thr[k] <= x)
			c = k;
	return (uint8_t)c;
}
static float unorm_alpha(uint8_t a) { return (float)a / 255.0f; }
static uint8_t pack_unorm8(float x) { /* packUnorm4x8: round(clamp(c, 0, 1) * 255), ties to even; NaN -> 0 */
	if (!(x > 0.0f))
		return 0;
	if (x > 1.0f)
		x = 1.0f;
	return (uint8_t)rintf(x * 255.0f);
}
static float lerpf(float a, float b, float t) { return a + t * (b - a); }

/* four texels of one level at integer columns i0,i1 / rows j0,j1, filtered: out = linear RGBA */
static void filter4(const orc_texset *s, const orc_level *L, int64_t i0, int64_t i1, int64_t j0, int64_t j1, float a, float b,
                    float out[4]) {
	const uint8_t *t00 = L->px + 4 * ((size_t)j0 * L->w + (size_t)i0), *t10 = L->px + 4 * ((size_t)j0 * L->w + (size_t)i1);
	const uint8_t *t01 = L->px + 4 * ((size_t)j1 * L->w + (size_t)i0), *t11 = L->px + 4 * ((size_t)j1 * L->w + (size_t)i1);
	for (int c = 0; c < 4; ++c) {
		float v00, v10, v01, v11;
		if (c < 3)
			v00 = s->decode[t00[c]], v10 = s->decode[t10[c]], v01 = s->decode[t01[c]], v11 = s->decode[t11[c]];
		else
			v00 = unorm_alpha(t00[c]), v10 = unorm_alpha(t10[c]), v01 = unorm_alpha(t01[c]), v11 = unorm_alpha(t11[c]);
		out[c] = lerpf(lerpf(v00, v10, a), lerpf(v01, v11, a), b);
	}
}

/* vkCmdBlitImage with VK_FILTER_LINEAR from level l-1 to level l (CommandBuffer.cpp:353-381): the source
 * coordinate of a destination texel centre is scaled by the size ratio, then linearly filtered with
 * clamp-to-edge; sRGB images filter in linear space and re-encode. */
static void downsample(const orc_texset *s, const orc_level *src, orc_level *dst) {
	const double sx = (double)src->w / (double)dst->w, sy = (double)src->h / (double)dst->h;
	for (uint32_t j = 0; j < dst->h; ++j)
		for (uint32_t i = 0; i < dst->w; ++i) {
			const double U = ((double)i + 0.5) * sx - 0.5, V = ((double)j + 0.5) * sy - 0.5;
			const double fu = floor(U), fv = floor(V);
			const float a = (float)(U - fu), b = (float)(V - fv);
			int64_t i0 = (int64_t)fu, i1 = i0 + 1, j0 = (int64_t)fv, j1 = j0 + 1;
			if (i0 < 0) i0 = 0;
			if (j0 < 0) j0 = 0;
			if (i1 > (int64_t)src->w - 1) i1 = (int64_t)src->w - 1;
			if (j1 > (int64_t)src->h - 1) j1 = (int64_t)src->h - 1;
			if (i0 > (int64_t)src->w - 1) i0 = (int64_t)src->w - 1;
			if (j0 > (int64_t)src->h - 1) j0 = (int64_t)src->h - 1;
			float v[4];
			filter4(s, src, i0, i1, j0, j1, a, b, v);
			uint8_t *o = dst->px + 4 * ((size_t)j * dst->w + i);
			for (int c = 0; c < 3; ++c)
				o[c] = encode_srgb(s, v[c]);
			o[3] = pack_unorm8(v[3]);
		}
}

orc_texset *orc_texset_create(const orc_texture *tex, uint32_t n) {
	orc_texset *s = (orc_texset *)calloc(1, sizeof(orc_texset));
	if (!s)
		return NULL;
	for (int c = 0; c < 256; ++c) {
		s->decode[c] = (float)srgb_to_linear((double)c / 255.0);
		s->enc_thr[c] = c ? (float)srgb_to_linear(((double)c - 0.5) / 255.0) : 0.0f;
	}
	for (int k = 0; k < 128; ++k)
		s->lod_thr[k] = exp2((double)k / 128.0);
	s->n = n;
	s->t = (orc_tex *)calloc(n ? n : 1, sizeof(orc_tex));
	for (uint32_t i = 0; i < n; ++i) {
		orc_tex *t = &s->t[i];
		if (tex[i].width == 0 || tex[i].height == 0 || !tex[i].rgba8) {
			orc_texset_destroy(s);
			return NULL;
		}
		t->levels = mip_level_count(tex[i].width, tex[i].height);
		t->lv = (orc_level *)calloc(t->levels, sizeof(orc_level));
		uint32_t w = tex[i].width, h = tex[i].height;
		for (uint32_t l = 0; l < t->levels; ++l) {
			t->lv[l].w = w, t->lv[l].h = h;
			t->lv[l].px = (uint8_t *)malloc((size_t)w * h * 4);
			if (l == 0)
				memcpy(t->lv[0].px, tex[i].rgba8, (size_t)w * h * 4);
			else
				downsample(s, &t->lv[l - 1], &t->lv[l]);
			w = w / 2 ? w / 2 : 1; /* std::max(mip_width / 2, 1) (CommandBuffer.cpp:362-363) */
			h = h / 2 ? h / 2 : 1;
		}
	}
	return s;
}
void orc_texset_destroy(orc_texset *s) {
	if (!s)
		return;
	for (uint32_t i = 0; i < s->n && s->t; ++i) {
		for (uint32_t l = 0; l < s->t[i].levels && s->t[i].lv; ++l)
			free(s->t[i].lv[l].px);
		free(s->t[i].lv);
	}
	free(s->t);
	free(s);
}
int orc_texset_level(const orc_texset *s, uint32_t tex, uint32_t level, uint32_t *w, uint32_t *h, const uint8_t **data) {
	if (!s || tex >= s->n || level >= s->t[tex].levels)
		return -1;
	*w = s->t[tex].lv[level].w, *h = s->t[tex].lv[level].h, *data = s->t[tex].lv[level].px;
	return (int)s->t[tex].levels;
}

/* the affine texture-coordinate map of one triangle and its (constant) level of detail */
typedef struct {
	double x0, y0, u0, v0, dudx, dudy, dvdx, dvdy;
	uint32_t hi, lo; /* mip levels blended */
	float delta;     /* weight of lo */
} orc_uvmap;

static void uvmap_setup(const orc_texset *s, const orc_tex *tex, const float *const p[3], const float *const uv[3], uint32_t axis,
                        uint32_t res, orc_uvmap *m) {
	const float fres = (float)res;
	double x[3], y[3];
	for (int i = 0; i < 3; ++i) { /* Project() (voxelizer.geom:15-19) + viewport, like orc_tri_setup */
		const float qx = axis == 0u ? p[i][1] : (axis == 1u ? p[i][2] : p[i][0]);
		const float qy = axis == 0u ? p[i][2] : (axis == 1u ? p[i][0] : p[i][1]);
		x[i] = (double)((qx + 1.0f) * 0.5f * fres);
		y[i] = (double)((qy + 1.0f) * 0.5f * fres);
	}
	m->x0 = x[0], m->y0 = y[0], m->u0 = (double)uv[0][0], m->v0 = (double)uv[0][1];
	const double dx1 = x[1] - x[0], dy1 = y[1] - y[0], dx2 = x[2] - x[0], dy2 = y[2] - y[0];
	const double det = dx1 * dy2 - dx2 * dy1;
	m->dudx = m->dudy = m->dvdx = m->dvdy = 0.0;
	if (det != 0.0 && isfinite(det)) {
		const double du1 = (double)uv[1][0] - m->u0, du2 = (double)uv[2][0] - m->u0;
		const double dv1 = (double)uv[1][1] - m->v0, dv2 = (double)uv[2][1] - m->v0;
		m->dudx = (du1 * dy2 - du2 * dy1) / det;
		m->dudy = (du2 * dx1 - du1 * dx2) / det;
		m->dvdx = (dv1 * dy2 - dv2 * dy1) / det;
		m->dvdy = (dv2 * dx1 - dv1 * dx2) / det;
	}
	const double W = (double)tex->lv[0].w, H = (double)tex->lv[0].h;
	const double ax = m->dudx * W, ay = m->dvdx * H, bx = m->dudy * W, by = m->dvdy * H;
	const double r2x = ax * ax + ay * ay, r2y = bx * bx + by * by;
	const double r2 = r2y > r2x ? r2y : r2x;
	const uint32_t q = tex->levels - 1u;
	m->hi = m->lo = 0, m->delta = 0.0f;
	if (!(r2 > 1.0))
		return; /* lambda <= 0 (or NaN): magnification */
	if (!isfinite(r2)) {
		m->hi = m->lo = q;
		return;
	}
	int E;
	const double f2 = frexp(r2, &E) * 2.0; /* r2 = f2 * 2^(E-1), f2 in [1,2) */
	int k = 0;
	for (int i = 1; i < 128; ++i)
		if (s->lod_thr[i] <= f2)
			k = i;
	const int64_t lam256 = 128 * (int64_t)(E - 1) + k; /* floor(256 * log2(rho)) = floor(128 * log2(r2)) */
	const int64_t hi = lam256 >> 8;
	if (hi >= (int64_t)q) {
		m->hi = m->lo = q;
		return;
	}
	m->hi = (uint32_t)hi, m->lo = (uint32_t)hi + 1u;
	m->delta = (float)(lam256 & 255) / 256.0f;
	if (m->delta == 0.0f)
		m->lo = m->hi;
}

static void sample_level(const orc_texset *s, const orc_level *L, double u, double v, float out[4]) {
	double U = u * (double)L->w - 0.5, V = v * (double)L->h - 0.5;
	if (!(fabs(U) < 4503599627370496.0)) U = 0.0; /* non-finite / absurd coordinates: pinned to texel 0 */
	if (!(fabs(V) < 4503599627370496.0)) V = 0.0;
	const double fu = floor(U), fv = floor(V);
	const float a = (float)(U - fu), b = (float)(V - fv);
	int64_t i0 = (int64_t)fu % (int64_t)L->w, j0 = (int64_t)fv % (int64_t)L->h; /* VK_SAMPLER_ADDRESS_MODE_REPEAT */
	if (i0 < 0) i0 += L->w;
	if (j0 < 0) j0 += L->h;
	const int64_t i1 = i0 + 1 == (int64_t)L->w ? 0 : i0 + 1, j1 = j0 + 1 == (int64_t)L->h ? 0 : j0 + 1;
	filter4(s, L, i0, i1, j0, j1, a, b, out);
}

/* voxelizer.frag:28-30,35,42: x = texture(...); if (x.a < 0.5) discard; packUnorm4x8(x) & 0xffffff */
static int shade(const float c[4], uint32_t *rgb) {
	if (c[3] < 0.5f)
		return 0;
	*rgb = (uint32_t)pack_unorm8(c[0]) | ((uint32_t)pack_unorm8(c[1]) << 8) | ((uint32_t)pack_unorm8(c[2]) << 16);
	return 1;
}

/* voxelizer.frag:27-36: returns 0 when the fragment is discarded (alpha < 0.5), else 1 and the packed colour */
/* texture(uTextures[id], gTexcoord) at the centre of pixel (px,py): the interpolated coordinates and the filtered value */
static void sample_value(const orc_texset *s, const orc_tex *tex, const orc_uvmap *m, int32_t px, int32_t py, double uv_out[2],
                         float c[4]) {
	const double cx = (double)px + 0.5, cy = (double)py + 0.5;
	const double u = fma(m->dudx, cx - m->x0, fma(m->dudy, cy - m->y0, m->u0));
	const double v = fma(m->dvdx, cx - m->x0, fma(m->dvdy, cy - m->y0, m->v0));
	if (uv_out)
		uv_out[0] = u, uv_out[1] = v;
	sample_level(s, &tex->lv[m->hi], u, v, c);
	if (m->lo != m->hi) {
		float d[4];
		sample_level(s, &tex->lv[m->lo], u, v, d);
		for (int k = 0; k < 4; ++k)
			c[k] = lerpf(c[k], d[k], m->delta);
	}
}
static int sample_colour(const orc_texset *s, const orc_tex *tex, const orc_uvmap *m, int32_t px, int32_t py, uint32_t *rgb) {
	float c[4];
	sample_value(s, tex, m, px, py, NULL, c);
	return shade(c, rgb);
}
/* debug view for the SPIR-V cross-check: what the fixed-function stages hand to voxelizer.frag for one pixel of a textured
 * triangle -- the interpolated gTexcoord and the value texture() returns (pinned arithmetic) */
void orc_debug_texture_fetch(const orc_texset *s, uint32_t tex, const float *p0, const float *p1, const float *p2, const float *uv0,
                             const float *uv1, const float *uv2, uint32_t level, int32_t px, int32_t py, double uv_out[2],
                             float rgba_out[4]) {
	orc_tri t;
	orc_tri_setup(p0, p1, p2, 1u << level, ORC_CENTER, &t);
	const float *const p[3] = {p0, p1, p2}, *const uv[3] = {uv0, uv1, uv2};
	orc_uvmap m;
	uvmap_setup(s, &s->t[tex], p, uv, t.axis, 1u << level, &m);
	sample_value(s, &s->t[tex], &m, px, py, uv_out, rgba_out);
}
uint32_t orc_debug_shade(const float rgba[4]) {
	uint32_t rgb = 0;
	return shade(rgba, &rgb) ? (0xff000000u | rgb) : 0u;
}

uint32_t orc_debug_sample(const orc_texset *s, uint32_t tex, const float *p0, const float *p1, const float *p2, const float *uv0,
                          const float *uv1, const float *uv2, uint32_t level, int32_t px, int32_t py, uint32_t lod_out[3]) {
	orc_tri t;
	orc_tri_setup(p0, p1, p2, 1u << level, ORC_CENTER, &t);
	const float *const p[3] = {p0, p1, p2}, *const uv[3] = {uv0, uv1, uv2};
	orc_uvmap m;
	uvmap_setup(s, &s->t[tex], p, uv, t.axis, 1u << level, &m);
	if (lod_out)
		lod_out[0] = m.hi, lod_out[1] = m.lo, lod_out[2] = (uint32_t)(m.delta * 256.0f);
	uint32_t rgb = 0;
	return sample_colour(s, &s->t[tex], &m, px, py, &rgb) ? (0xff000000u | rgb) : 0u;
}

int64_t orc_voxelize_textured(const void *positions, uint32_t pos_stride_bytes, const void *texcoords, uint32_t uv_stride_bytes,
                              const uint32_t *indices, const orc_draw *draws, uint32_t n_draws, const orc_texset *texset,
                              uint32_t level, int mode, const uint32_t *shard_lo, const uint32_t *shard_hi, orc_frag *out,
                              int64_t cap, int nthreads) {
	if (level < 1 || level > 16 || (mode != ORC_CENTER && mode != ORC_CONSERVATIVE_EXACT && mode != ORC_CONSERVATIVE_DILATE))
		return -1;
	for (uint32_t d = 0; d < n_draws; ++d)
		if (draws[d].texture_id != 0xffffffffu && (!texset || !texcoords || draws[d].texture_id >= texset->n))
			return -2; /* a textured draw needs texture coordinates and its texture */
	const uint32_t res = 1u << level;
	const unsigned char *pbase = (const unsigned char *)positions, *tbase = (const unsigned char *)texcoords;
	int64_t counter = 0; /* uCounter (voxelizer.frag:5,37) */
	if (nthreads < 1)
		nthreads = 1;

	for (uint32_t d = 0; d < n_draws; ++d) { /* Scene::CmdDraw: one draw per material (Scene.cpp:450-463) */
		const uint32_t ntri = draws[d].index_count / 3u;
		const uint32_t albedo = draws[d].albedo_rgba8 & 0xffffffu;
		const orc_tex *tex = draws[d].texture_id != 0xffffffffu ? &texset->t[draws[d].texture_id] : NULL;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) if (nthreads > 1)
		for (uint32_t k = 0; k < ntri; ++k) {
			const uint32_t *ix = indices + draws[d].first_index + 3u * k;
			const float *const p[3] = {(const float *)(pbase + (size_t)ix[0] * pos_stride_bytes),
			                           (const float *)(pbase + (size_t)ix[1] * pos_stride_bytes),
			                           (const float *)(pbase + (size_t)ix[2] * pos_stride_bytes)};
			orc_tri t;
			orc_tri_setup(p[0], p[1], p[2], res, mode, &t);
			if (!t.valid)
				continue;
			if (t.area2 == 0 && mode != ORC_CONSERVATIVE_EXACT)
				continue;
			orc_plane pl;
			plane_setup(&t, &pl);
			orc_uvmap um;
			if (tex) {
				const float *const uv[3] = {(const float *)(tbase + (size_t)ix[0] * uv_stride_bytes),
				                            (const float *)(tbase + (size_t)ix[1] * uv_stride_bytes),
				                            (const float *)(tbase + (size_t)ix[2] * uv_stride_bytes)};
				uvmap_setup(texset, tex, p, uv, t.axis, res, &um);
			}
			int32_t xmin = t.X[0], xmax = t.X[0], ymin = t.Y[0], ymax = t.Y[0];
			for (int i = 1; i < 3; ++i) {
				if (t.X[i] < xmin) xmin = t.X[i];
				if (t.X[i] > xmax) xmax = t.X[i];
				if (t.Y[i] < ymin) ymin = t.Y[i];
				if (t.Y[i] > ymax) ymax = t.Y[i];
			}
			/* generous candidate range; every pixel is put through the full rule. Viewport/scissor
			 * (Voxelizer.cpp:115-116) limits pixels to [0,res). */
			int32_t px0 = floor_div256(xmin) - 1, px1 = floor_div256(xmax) + 1;
			int32_t py0 = floor_div256(ymin) - 1, py1 = floor_div256(ymax) + 1;
			if (px0 < 0) px0 = 0;
			if (py0 < 0) py0 = 0;
			if (px1 > (int32_t)res - 1) px1 = (int32_t)res - 1;
			if (py1 > (int32_t)res - 1) py1 = (int32_t)res - 1;
			for (int32_t py = py0; py <= py1; ++py)
				for (int32_t px = px0; px <= px1; ++px) {
					if (!covered(&t, mode, px, py))
						continue;
					uint32_t v[3];
					if (!frag_voxel(&t, &pl, res, px, py, v))
						continue;
					if (shard_lo && shard_hi) {
						if (v[0] < shard_lo[0] || v[0] >= shard_hi[0] || v[1] < shard_lo[1] ||
						    v[1] >= shard_hi[1] || v[2] < shard_lo[2] || v[2] >= shard_hi[2])
							continue;
					}
					uint32_t colour = albedo; /* voxelizer.frag:35 */
					if (tex && !sample_colour(texset, tex, &um, px, py, &colour))
						continue; /* voxelizer.frag:29-30 discard */
					int64_t cur; /* uint cur = atomicAdd(uCounter, 1u)  (voxelizer.frag:37) */
					if (nthreads > 1)
						cur = __atomic_fetch_add(&counter, 1, __ATOMIC_RELAXED);
					else
						cur = counter++;
					if (out && cur < cap) { /* uCountOnly == 0 (voxelizer.frag:39-43) */
						out[cur].x = v[0], out[cur].y = v[1], out[cur].z = v[2];
						out[cur].rgb = colour;
					}
				}
		}
	}
	return counter;
}

int64_t orc_voxelize(const void *positions, uint32_t pos_stride_bytes, const uint32_t *indices,
                     const orc_draw *draws, uint32_t n_draws, uint32_t level, int mode,
                     const uint32_t *shard_lo, const uint32_t *shard_hi, orc_frag *out, int64_t cap,
                     int nthreads) {
	return orc_voxelize_textured(positions, pos_stride_bytes, NULL, 0, indices, draws, n_draws, NULL, level, mode, shard_lo,
	                             shard_hi, out, cap, nthreads);
}

/* Debug views used by the SPIR-V cross-checks (tests/golden/make_spirv_golden.py): the geometry-stage
 * outputs of one triangle, and its covered pixels with the pinned depth, before voxelizer.frag. */
void orc_debug_tri_setup(const float *p0, const float *p1, const float *p2, uint32_t level, uint32_t out_axis_aabb_zr[7],
                         int32_t out_xy_snapped[6]) {
	orc_tri t;
	orc_tri_setup(p0, p1, p2, 1u << level, ORC_CONSERVATIVE_EXACT, &t);
	out_axis_aabb_zr[0] = t.axis;
	for (int i = 0; i < 4; ++i) out_axis_aabb_zr[1 + i] = t.aabb[i];
	out_axis_aabb_zr[5] = t.zr[0], out_axis_aabb_zr[6] = t.zr[1];
	for (int i = 0; i < 3; ++i) out_xy_snapped[2 * i] = t.valid ? t.X[i] : 0, out_xy_snapped[2 * i + 1] = t.valid ? t.Y[i] : 0;
}
int64_t orc_debug_raster_pixels(const float *p0, const float *p1, const float *p2, uint32_t level, int mode, int32_t *out_px,
                                int32_t *out_py, double *out_z, int64_t cap) {
	const uint32_t res = 1u << level;
	orc_tri t;
	orc_tri_setup(p0, p1, p2, res, mode, &t);
	if (!t.valid || (t.area2 == 0 && mode != ORC_CONSERVATIVE_EXACT)) return 0;
	orc_plane pl;
	plane_setup(&t, &pl);
	int32_t xmin = t.X[0], xmax = t.X[0], ymin = t.Y[0], ymax = t.Y[0];
	for (int i = 1; i < 3; ++i) {
		if (t.X[i] < xmin) xmin = t.X[i];
		if (t.X[i] > xmax) xmax = t.X[i];
		if (t.Y[i] < ymin) ymin = t.Y[i];
		if (t.Y[i] > ymax) ymax = t.Y[i];
	}
	int32_t px0 = floor_div256(xmin) - 1, px1 = floor_div256(xmax) + 1, py0 = floor_div256(ymin) - 1, py1 = floor_div256(ymax) + 1;
	if (px0 < 0) px0 = 0;
	if (py0 < 0) py0 = 0;
	if (px1 > (int32_t)res - 1) px1 = (int32_t)res - 1;
	if (py1 > (int32_t)res - 1) py1 = (int32_t)res - 1;
	int64_t n = 0;
	for (int32_t py = py0; py <= py1; ++py)
		for (int32_t px = px0; px <= px1; ++px) {
			if (!covered(&t, mode, px, py)) continue;
			if (n < cap) {
				out_px[n] = px, out_py[n] = py;
				out_z[n] = fma(pl.dzdx, (double)(px * 256 + 128 - pl.X0), fma(pl.dzdy, (double)(py * 256 + 128 - pl.Y0), pl.z0));
			}
			++n;
		}
	return n;
}

/* ------------------------------------------------------------------------------------------
 * OctreeBuilder
 * ---------------------------------------------------------------------------------------- */

/* octree_tag_node.comp:10-16 */
static void leaf_to_uvec4(uint32_t val, uint32_t v[4]) {
	v[0] = val & 0xffu, v[1] = (val >> 8u) & 0xffu, v[2] = (val >> 16u) & 0xffu, v[3] = (val >> 24u) & 0x3fu;
}
static uint32_t uvec4_to_leaf(const uint32_t v[4]) {
	uint32_t w = v[3] < 0x3fu ? v[3] : 0x3fu;
	return (w << 24u) | (v[0] & 0xffu) | ((v[1] & 0xffu) << 8u) | ((v[2] & 0xffu) << 16u) | 0xC0000000u;
}

/* octree_tag_node.comp:18-31 TraverseOctree */
static uint32_t traverse_octree(const uint32_t *octree, uint32_t res, const uint32_t voxel_pos[3], int *is_leaf) {
	uint32_t level_dim = res;
	uint32_t pos[3] = {voxel_pos[0], voxel_pos[1], voxel_pos[2]};
	uint32_t idx = 0u, cur = 0u;
	do {
		level_dim >>= 1;
		uint32_t cx = pos[0] >= level_dim, cy = pos[1] >= level_dim, cz = pos[2] >= level_dim;
		idx = cur | cx | (cy << 1u) | (cz << 2u);
		cur = __atomic_load_n(&octree[idx], __ATOMIC_RELAXED) & 0x3fffffffu;
		pos[0] -= cx * level_dim, pos[1] -= cy * level_dim, pos[2] -= cz * level_dim;
	} while (cur != 0u && level_dim > 1u);
	*is_leaf = level_dim == 1u;
	return idx;
}

/* octree_tag_node.comp:33-60 main(), one invocation */
static void tag_node(uint32_t *octree, uint32_t res, const orc_frag *f) {
	uint32_t pos[3] = {f->x, f->y, f->z};
	int is_leaf;
	uint32_t idx = traverse_octree(octree, res, pos, &is_leaf);
	if (is_leaf) {
		uint32_t prev_val = 0u, cur_val, new_val = 0xC1000000u | (f->rgb & 0xffffffu);
		uint32_t rgba[4];
		leaf_to_uvec4(new_val, rgba);
		for (;;) { /* atomicCompSwap loop (octree_tag_node.comp:50) */
			uint32_t expected = prev_val;
			if (__atomic_compare_exchange_n(&octree[idx], &expected, new_val, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
				break;
			cur_val = expected;
			prev_val = cur_val;
			uint32_t prev_rgba[4], cur_rgba[4];
			leaf_to_uvec4(prev_val, prev_rgba);
			for (int c = 0; c < 3; ++c)
				prev_rgba[c] *= prev_rgba[3];
			for (int c = 0; c < 4; ++c)
				cur_rgba[c] = prev_rgba[c] + rgba[c];
			for (int c = 0; c < 3; ++c)
				cur_rgba[c] /= cur_rgba[3];
			new_val = uvec4_to_leaf(cur_rgba);
		}
	} else
		__atomic_store_n(&octree[idx], 0x80000000u, __ATOMIC_RELAXED); /* plain store, benign race */
}

int64_t orc_build(const orc_frag *frags, int64_t n_frags, uint32_t level, uint32_t *words,
                  uint64_t cap_words, int nthreads) {
	if (level < 1 || level > 16)
		return -1;
	const uint32_t res = 1u << level;
	if (nthreads < 1)
		nthreads = 1;
	/* OctreeBuilder.cpp:27-31 build_info = {allocBegin 0, allocNum 8}; Counter reset 0 (:14-15) */
	uint64_t alloc_begin = 0, alloc_num = 8;
	uint32_t counter = 0;
	for (uint32_t i = 1; i <= level; ++i) { /* OctreeBuilder.cpp:167 */
		if (alloc_begin + alloc_num > cap_words)
			return -1;
		/* octree_init_node.comp:7-11 */
		memset(words + alloc_begin, 0, (size_t)alloc_num * sizeof(uint32_t));
		/* octree_tag_node.comp, ceil(F/64) groups (OctreeBuilder.cpp:163,177-178) */
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
		for (int64_t f = 0; f < n_frags; ++f)
			tag_node(words, res, &frags[f]);
		if (i != level) {
			/* octree_alloc_node.comp:9-23 (one atomicAdd per flagged word; the subgroup
			 * aggregation changes only which id a word gets, which is arbitrary anyway) */
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
			for (uint64_t k = alloc_begin; k < alloc_begin + alloc_num; ++k) {
				if (words[k] & 0x80000000u) {
					uint32_t cur = nthreads > 1 ? __atomic_fetch_add(&counter, 1u, __ATOMIC_RELAXED) : counter++;
					words[k] = ((cur + 1u) << 3u) | 0x80000000u;
				}
			}
			/* octree_modify_arg.comp:9-13 */
			alloc_begin += alloc_num;
			alloc_num = ((uint64_t)counter << 3u) - alloc_begin + 8u;
		}
	}
	/* OctreeBuilder.cpp:212-214 */
	return ((int64_t)counter + 1) * 8 * (int64_t)sizeof(uint32_t);
}

/* ------------------------------------------------------------------------------------------
 * Canonical form (test helper; layout per octree.glsl:87-110)
 * ---------------------------------------------------------------------------------------- */
int64_t orc_canonicalise(const uint32_t *words, uint64_t n_words, uint32_t level, uint8_t *out_depth,
                         uint64_t *out_morton, uint32_t *out_word, int64_t cap) {
	if (level < 1 || level > 16 || n_words < 8)
		return -1;
	/* explicit DFS stack: (block word index, next slot, morton prefix) per depth */
	uint32_t blk[17];
	uint32_t slot[17];
	uint64_t pre[17];
	int64_t n = 0;
	int d = 1;
	blk[1] = 0, slot[1] = 0, pre[1] = 0;
	while (d >= 1) {
		if (slot[d] == 8u) {
			--d;
			continue;
		}
		uint32_t s = slot[d]++;
		uint32_t w = words[blk[d] + s];
		if (w == 0u)
			continue;
		if (!(w & 0x80000000u))
			return -4;
		uint64_t m = (pre[d] << 3u) | s;
		if (w & 0x40000000u) { /* leaf */
			if ((uint32_t)d != level)
				return -2;
			if (n < cap) {
				if (out_depth) out_depth[n] = (uint8_t)d;
				if (out_morton) out_morton[n] = m;
				if (out_word) out_word[n] = w;
			}
			++n;
		} else {
			if ((uint32_t)d == level)
				return -3;
			uint32_t ptr = w & 0x3fffffffu;
			if (ptr == 0u || (ptr & 7u) || (uint64_t)ptr + 8u > n_words)
				return -1;
			if (n < cap) {
				if (out_depth) out_depth[n] = (uint8_t)d;
				if (out_morton) out_morton[n] = m;
				if (out_word) out_word[n] = 0x80000000u;
			}
			++n;
			++d;
			blk[d] = ptr, slot[d] = 0, pre[d] = m;
		}
	}
	return n;
}

/* Debug view for the SPIR-V cross-check of Mode B: the three vertices voxelizer_conservative.geom emits
 * (ndc x, ndc y, depth), in emit order, as fp32. */
void orc_debug_dilate(const float *p0, const float *p1, const float *p2, uint32_t level, float out[9]) {
	const float *p[3] = {p0, p1, p2};
	float e1[3], e2[3], n[3], w[3], q[3][3];
	for (int k = 0; k < 3; ++k) e1[k] = p1[k] - p0[k], e2[k] = p2[k] - p0[k];
	n[0] = e1[1] * e2[2] - e1[2] * e2[1];
	n[1] = e1[2] * e2[0] - e1[0] * e2[2];
	n[2] = e1[0] * e2[1] - e1[1] * e2[0];
	for (int k = 0; k < 3; ++k) w[k] = fabsf(n[k]);
	const uint32_t axis = (w[0] > w[1] && w[0] > w[2]) ? 0u : ((w[1] > w[2]) ? 1u : 2u);
	for (int i = 0; i < 3; ++i) {
		if (axis == 0u) q[i][0] = p[i][1], q[i][1] = p[i][2], q[i][2] = p[i][0];
		else if (axis == 1u) q[i][0] = p[i][2], q[i][1] = p[i][0], q[i][2] = p[i][1];
		else q[i][0] = p[i][0], q[i][1] = p[i][1], q[i][2] = p[i][2];
		q[i][2] = (q[i][2] + 1.0f) * 0.5f;
	}
	dilate_mode_b(q, n[axis], 1u << level);
	memcpy(out, q, sizeof(q));
}

/* ------------------------------------------------------------------------------------------
 * The consumer side, for verification only: Octree_RayMarchLeaf (octree.glsl:179-340, the variant
 * octree_tracer.frag:36 calls), the stack-based parametric octree traversal of Laine & Karras,
 * "Efficient Sparse Voxel Octrees" (2010), which the reference adopts.  The octree occupies [1,2]^3
 * and the ray is mirrored so that every direction component is negative; positions are handled as
 * fp32 bit patterns, one mantissa bit per level (hence the 23-entry stack).  fp32, one rounding per
 * operator, with fused multiply-adds exactly where the reference's compiled octree_tracer.frag has
 * them (every "k * t_coef - t_bias" and "pos * t_coef - t_bias"), so that it is bit-identical to
 * the executed binary (tests/golden/spirv_tracer_*.npz).
 * ---------------------------------------------------------------------------------------- */
static uint32_t f2bits(float f) {
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}
static float bits2f(uint32_t u) {
	float f;
	memcpy(&f, &u, 4);
	return f;
}
static float min3f(float a, float b, float c) { /* GLSL min(min(a,b),c), min(x,y) = y < x ? y : x */
	float m = b < a ? b : a;
	return c < m ? c : m;
}
static float max3f(float a, float b, float c) {
	float m = a < b ? b : a;
	return m < c ? c : m;
}

int orc_raymarch_leaf(const uint32_t *octree, const float o[3], const float d_in[3], float o_pos[3], float o_colour[3],
                      float o_normal[3], uint32_t *o_iter) {
	enum { STACK = 23 };            /* octree.glsl:43 */
	const float EPS = 3.552713678800501e-15f; /* octree.glsl:44 */
	uint32_t stack[STACK] = {0};
	float d[3], t_coef[3], t_bias[3], pos[3] = {1.0f, 1.0f, 1.0f};
	uint32_t oct_mask = 0u, iter = 0u;
	for (int k = 0; k < 3; ++k) { /* octree.glsl:182-199 */
		d[k] = fabsf(d_in[k]) > EPS ? d_in[k] : (d_in[k] >= 0 ? EPS : -EPS);
		t_coef[k] = 1.0f / -fabsf(d[k]);
		t_bias[k] = t_coef[k] * o[k];
		if (d[k] > 0.0f) {
			oct_mask ^= 1u << k;
			t_bias[k] = fmaf(3.0f, t_coef[k], -t_bias[k]);
		}
	}
	/* octree.glsl:201-205: the active span of t */
	float t_min = max3f(fmaf(2.0f, t_coef[0], -t_bias[0]), fmaf(2.0f, t_coef[1], -t_bias[1]), fmaf(2.0f, t_coef[2], -t_bias[2]));
	const float t_max = min3f(t_coef[0] - t_bias[0], t_coef[1] - t_bias[1], t_coef[2] - t_bias[2]);
	t_min = t_min < 0.0f ? 0.0f : t_min;
	float h = t_max;
	uint32_t parent = 0u, cur = 0u, idx = 0u;
	for (int k = 0; k < 3; ++k) /* octree.glsl:207-216: first child */
		if (fmaf(1.5f, t_coef[k], -t_bias[k]) > t_min)
			idx ^= 1u << k, pos[k] = 1.5f;
	uint32_t scale = STACK - 1;
	float scale_exp2 = 0.5f;

	while (scale < STACK) { /* octree.glsl:221-302 */
		++iter;
		if (cur == 0u)
			cur = octree[parent + (idx ^ oct_mask)];
		float t_corner[3];
		for (int k = 0; k < 3; ++k)
			t_corner[k] = fmaf(pos[k], t_coef[k], -t_bias[k]);
		const float tc_max = min3f(t_corner[0], t_corner[1], t_corner[2]);
		if ((cur & 0x80000000u) != 0u && t_min <= t_max) {
			const float half = scale_exp2 * 0.5f;
			if ((cur & 0x40000000u) != 0u)
				break; /* leaf */
			if (tc_max < h) /* push */
				stack[scale] = parent;
			h = tc_max;
			parent = cur & 0x3fffffffu;
			idx = 0u;
			--scale;
			scale_exp2 = half;
			for (int k = 0; k < 3; ++k)
				if (half * t_coef[k] + t_corner[k] > t_min)
					idx ^= 1u << k, pos[k] += scale_exp2;
			cur = 0u;
			continue;
		}
		/* advance */
		uint32_t step_mask = 0u;
		for (int k = 0; k < 3; ++k)
			if (t_corner[k] <= tc_max)
				step_mask ^= 1u << k, pos[k] -= scale_exp2;
		t_min = tc_max;
		idx ^= step_mask;
		if ((idx & step_mask) != 0u) { /* pop: the highest differing mantissa bit names the level to return to */
			uint32_t differing = 0u;
			for (int k = 0; k < 3; ++k)
				if (step_mask & (1u << k))
					differing |= f2bits(pos[k]) ^ f2bits(pos[k] + scale_exp2);
			int msb = -1; /* findMSB(0) = -1 -> scale = 0xffffffff >= STACK */
			for (int b = 31; b >= 0; --b)
				if (differing & (1u << b)) {
					msb = b;
					break;
				}
			scale = (uint32_t)msb;
			if (scale >= STACK)
				break;
			scale_exp2 = bits2f((scale - STACK + 127u) << 23u);
			parent = stack[scale];
			uint32_t sh[3];
			for (int k = 0; k < 3; ++k) {
				sh[k] = f2bits(pos[k]) >> scale;
				pos[k] = bits2f(sh[k] << scale);
			}
			idx = (sh[0] & 1u) | ((sh[1] & 1u) << 1u) | ((sh[2] & 1u) << 2u);
			h = 0.0f;
			cur = 0u;
		}
	}

	/* octree.glsl:304-339: normal of the entry face, un-mirror, outputs */
	float t_corner[3], norm[3] = {0.0f, 0.0f, 0.0f};
	for (int k = 0; k < 3; ++k)
		t_corner[k] = fmaf(t_coef[k], pos[k] + scale_exp2, -t_bias[k]);
	const int axis = (t_corner[0] > t_corner[1] && t_corner[0] > t_corner[2]) ? 0 : (t_corner[1] > t_corner[2] ? 1 : 2);
	norm[axis] = -1.0f;
	for (int k = 0; k < 3; ++k) {
		if ((oct_mask & (1u << k)) == 0u)
			norm[k] = -norm[k];
		else
			pos[k] = 3.0f - scale_exp2 - pos[k];
	}
	for (int k = 0; k < 3; ++k) {
		float p = o[k] + t_min * d[k];
		const float lo = pos[k], hi = pos[k] + scale_exp2;
		p = p < lo ? lo : p; /* clamp(x, lo, hi) = min(max(x, lo), hi) */
		p = hi < p ? hi : p;
		if (norm[k] != 0.0f)
			p = norm[k] > 0.0f ? pos[k] + scale_exp2 + EPS * 2.0f : pos[k] - EPS;
		o_pos[k] = p;
		o_normal[k] = norm[k] == 0.0f ? 0.0f : norm[k]; /* no negative zero */
		o_colour[k] = (float)((cur >> (8 * k)) & 0xffu) / 255.0f; /* unpackUnorm4x8(cur).xyz */
	}
	*o_iter = iter;
	return scale < STACK && t_min <= t_max;
}

