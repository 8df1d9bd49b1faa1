// raymarch.cuh -- the consumer side of the node buffer, for verification of built trees without Vulkan
// (SURVEY.md section 8 row f4): primary-ray traversal with the semantics of the reference's
// Octree_RayMarchLeaf (shader/octree.glsl:179-340, called by octree_tracer.frag:36), i.e. the parametric
// stack traversal of Laine & Karras, "Efficient Sparse Voxel Octrees" (2010).
//
// One thread per ray.  The octree occupies [1,2]^3; the ray is mirrored so that all direction components are
// negative; a cube position is an fp32 bit pattern with one mantissa bit per level, which is why popping to the
// common ancestor is a find-MSB on the differing bits and why the parent stack has 23 entries.
// fp32 with one rounding per operator and fused multiply-adds exactly where the reference's compiled
// octree_tracer.frag has them, so hits, iteration counts, normals and positions are bit-identical to the
// executed reference binary (tests/golden/spirv_tracer_*.npz).
#pragma once
#include "svo_math.cuh"

namespace svo {

struct RayHit {
	float pos[3], colour[3], normal[3];
	uint32_t hit, iter;
};

SVO_DEV float fmin3(float a, float b, float c) { // GLSL min(min(a,b),c) with min(x,y) = y < x ? y : x
	const float m = b < a ? b : a;
	return c < m ? c : m;
}
SVO_DEV float fmax3(float a, float b, float c) {
	const float m = a < b ? b : a;
	return m < c ? c : m;
}

SVO_DEV void raymarch_leaf(const uint32_t *__restrict__ octree, const float (&org)[3], const float (&dir)[3], RayHit &out) {
	constexpr uint32_t LEVELS = 23;                // mantissa bits = deepest addressable level
	constexpr float TINY = 3.552713678800501e-15f; // 2^-48: keeps 1/d finite for axis-parallel rays
	uint32_t parents[LEVELS];
	float d[3], coef[3], bias[3], cube[3];
	uint32_t mirror = 0u;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		d[k] = fabsf(dir[k]) > TINY ? dir[k] : (dir[k] >= 0.0f ? TINY : -TINY);
		coef[k] = fdiv(-1.0f, fabsf(d[k])); // t_k(x) = x * coef_k - bias_k
		bias[k] = fmul(coef[k], org[k]);
		if (d[k] > 0.0f) {
			mirror |= 1u << k;
			bias[k] = ffma(3.0f, coef[k], -bias[k]);
		}
		cube[k] = 1.0f;
	}
	float t_lo = fmax3(ffma(2.0f, coef[0], -bias[0]), ffma(2.0f, coef[1], -bias[1]), ffma(2.0f, coef[2], -bias[2]));
	const float t_hi = fmin3(fsub(coef[0], bias[0]), fsub(coef[1], bias[1]), fsub(coef[2], bias[2]));
	t_lo = t_lo < 0.0f ? 0.0f : t_lo;
	float h = t_hi;
	uint32_t block = 0u, word = 0u, slot = 0u, iter = 0u;
#pragma unroll
	for (int k = 0; k < 3; ++k)
		if (ffma(1.5f, coef[k], -bias[k]) > t_lo) slot |= 1u << k, cube[k] = 1.5f;
	uint32_t level = LEVELS - 1u; // mantissa bit of the current cube size
	float size = 0.5f;

	while (level < LEVELS) {
		++iter;
		if (word == 0u) word = __ldg(octree + block + (slot ^ mirror));
		float tc[3];
#pragma unroll
		for (int k = 0; k < 3; ++k) tc[k] = ffma(cube[k], coef[k], -bias[k]);
		const float t_exit = fmin3(tc[0], tc[1], tc[2]);
		if ((word & 0x80000000u) && t_lo <= t_hi) { // occupied and still inside the active span: descend, or stop at a leaf
			if (word & 0x40000000u) break;
			const float half = fmul(size, 0.5f);
			if (t_exit < h) parents[level] = block;
			h = t_exit;
			block = word & 0x3fffffffu;
			slot = 0u;
			--level;
			size = half;
#pragma unroll
			for (int k = 0; k < 3; ++k)
				if (fadd(fmul(half, coef[k]), tc[k]) > t_lo) slot |= 1u << k, cube[k] = fadd(cube[k], half);
			word = 0u;
			continue;
		}
		// step to the neighbour across the exit face(s)
		uint32_t moved = 0u;
#pragma unroll
		for (int k = 0; k < 3; ++k)
			if (tc[k] <= t_exit) moved |= 1u << k, cube[k] = fsub(cube[k], size);
		t_lo = t_exit;
		slot ^= moved;
		if (slot & moved) { // left the parent: pop to the lowest common ancestor
			uint32_t differing = 0u;
#pragma unroll
			for (int k = 0; k < 3; ++k)
				if (moved & (1u << k)) differing |= __float_as_uint(cube[k]) ^ __float_as_uint(fadd(cube[k], size));
			level = differing ? 31u - (uint32_t)__clz((int)differing) : 0xffffffffu;
			if (level >= LEVELS) break; // left the root cube
			size = __uint_as_float((level - LEVELS + 127u) << 23);
			block = parents[level];
			uint32_t bits[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				bits[k] = __float_as_uint(cube[k]) >> level;
				cube[k] = __uint_as_float(bits[k] << level);
			}
			slot = (bits[0] & 1u) | ((bits[1] & 1u) << 1) | ((bits[2] & 1u) << 2);
			h = 0.0f;
			word = 0u;
		}
	}

	// entry face = the axis with the largest entry parameter; undo the mirroring
	float te[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) te[k] = ffma(coef[k], fadd(cube[k], size), -bias[k]);
	const int axis = (te[0] > te[1] && te[0] > te[2]) ? 0 : (te[1] > te[2] ? 1 : 2);
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		float n = k == axis ? -1.0f : 0.0f;
		if (mirror & (1u << k))
			cube[k] = fsub(fsub(3.0f, size), cube[k]);
		else
			n = -n;
		float p = fadd(org[k], fmul(t_lo, d[k]));
		const float lo = cube[k], hi = fadd(cube[k], size);
		p = p < lo ? lo : p;
		p = hi < p ? hi : p;
		if (n != 0.0f) p = n > 0.0f ? fadd(hi, fmul(TINY, 2.0f)) : fsub(lo, TINY);
		out.pos[k] = p;
		out.normal[k] = n == 0.0f ? 0.0f : n;
		out.colour[k] = fdiv((float)((word >> (8 * k)) & 0xffu), 255.0f); // unpackUnorm4x8(leaf).xyz
	}
	out.iter = iter;
	out.hit = (level < LEVELS && t_lo <= t_hi) ? 1u : 0u;
}

__global__ void __launch_bounds__(128) k_raymarch_leaf(const uint32_t *__restrict__ octree, uint64_t n_rays, const float *__restrict__ origins,
                                                       const float *__restrict__ dirs, RayHit *__restrict__ hits) {
	const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
	if (i >= n_rays) return;
	const float o[3] = {origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]};
	const float d[3] = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
	RayHit h;
	raymarch_leaf(octree, o, d, h);
	hits[i] = h;
}

} // namespace svo
