// bucket.cuh -- the builder's main path: sort RUNS of fragments, not fragments, then finish bucket by bucket on chip.
//
// Replaces the reference's tag / alloc level loop (shader/octree_tag_node.comp:18-60, src/OctreeBuilder.cpp:167-209),
// like sort.cuh + build.cuh do, but with far fewer bytes moved.  Fragments leave the voxelizer in raster order, so
// consecutive fragments mostly fall into the same small cube of voxels (a "bucket": the voxels that share the upper
// Morton bits; 16^3 voxels up to level 12).  Instead of pushing every 8-byte fragment through 4 radix passes
// (96 bytes of traffic per fragment, hist + passes + reduce), the fragment list is cut into runs of equal bucket:
//
//   k_runs           one read of the fragments: a run = consecutive fragments (inside an aligned group of 32) of one
//                    bucket -> 8-byte record (bucket, start, length), written in stream order (chained scan)
//   radix sort       of the RECORDS by bucket (sort.cuh, stable: the runs of a bucket keep their stream order);
//                    ~15x fewer keys than fragments on wall-dominated scenes, and only the bucket bits
//   k_run_scan       fragment offset of every sorted run (= where it would sit in a bucket-sorted fragment list) and
//                    the list of bucket starts;  k_bucket_plan cuts that order into segments of whole buckets
//   k_bucket_sort_reduce  one thread block per segment: gathers the segment's fragments run by run (runs are
//                    contiguous pieces of the fragment list: coalesced), sorts them by the remaining Morton bits in
//                    shared memory (stable LSD passes, warp-vote ranking as in the onesweep kernel), and reduces the
//                    sorted keys right there: leaves with the reference's colour running average, and the two parent
//                    levels above them -- the outputs k_reduce_fused (build.cuh) produced from a fully sorted list.
//                    The segments' positions in the outputs come from a decoupled look-back over three counters.
// Fragment traffic: one read for the runs, one gathered read for the segments -- 16 bytes per fragment.
//
// The path needs coherent input and bounded buckets.  Both are checked on the device: more runs than fragments / 4, or
// a bucket with more fragments than a block can hold, sets *mode = 1 and the classic path (full onesweep sort of the
// fragments + k_reduce_count / k_reduce_fused) runs instead -- every kernel of either path is launched and looks at
// *mode first, so no host round trip is needed.
#pragma once
#include "build.cuh"
#include "sort.cuh"

namespace svo {

inline bool g_use_buckets = true; // svo_debug_use_bucket_path(0): classic path only (tests, comparisons)

// ---- run records --------------------------------------------------------------------------------------------------
// record = bucket << 37 | start << 5 | (length - 1): bucket <= 24 bits, start < 2^32, length 1..32
constexpr uint32_t REC_BUCKET_SHIFT = 37, REC_START_SHIFT = 5;
constexpr uint32_t MAX_BUCKET_BITS = 24; // bucket ids: the bits of the Morton code above the low LBITS
SVO_HD inline uint32_t rec_bucket(uint64_t r) { return (uint32_t)(r >> REC_BUCKET_SHIFT); }
SVO_HD inline uint32_t rec_start(uint64_t r) { return (uint32_t)(r >> REC_START_SHIFT); }
SVO_HD inline uint32_t rec_len(uint64_t r) { return ((uint32_t)r & 31u) + 1u; }

// Morton bits sorted inside a bucket: 12 (16^3 voxels), more only where the bucket id would not fit 24 bits
inline uint32_t bucket_low_bits(uint32_t level) {
	const uint32_t total = 3 * level;
	uint32_t lb = total > 12 + MAX_BUCKET_BITS ? total - MAX_BUCKET_BITS : 12;
	return lb < total ? lb : total;
}

constexpr int RUN_BLOCK = 256, RUN_ITEMS = 16, RUN_TILE = RUN_BLOCK * RUN_ITEMS, RUN_NW = RUN_BLOCK / 32;
static_assert(RUN_ITEMS * RUN_NW == 128, "the (row, warp) run counts are scanned by one warp, 4 per lane");

// device-side scalars of one build
struct BucketCtl {
	uint32_t mode;      // 0: this path; 1: classic path (set by k_runs / k_bucket_plan)
	uint32_t n_runs;    // records written (<= capacity)
	uint32_t n_heads;   // buckets that hold fragments
	uint32_t ticket[5]; // k_runs, k_run_scan, k_bucket_sort_reduce, spare
};

// fragments -> run records in stream order.  Tiles of RUN_TILE fragments numbered by a ticket; the number of runs in
// front of a tile comes from a chained scan (one word per tile).
__global__ void __launch_bounds__(RUN_BLOCK)
    k_runs(const uint64_t *__restrict__ frags, uint64_t n, uint32_t bucket_shift, uint64_t *__restrict__ rec, uint32_t rec_cap,
           uint64_t *state, BucketCtl *ctl) {
	__shared__ uint32_t s_cnt[RUN_ITEMS * RUN_NW];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_prefix;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint32_t tile = take_ticket(&ctl->ticket[0], &s_ticket);
	const uint32_t n_tiles = (uint32_t)((n + RUN_TILE - 1) / RUN_TILE);
	if (tile >= n_tiles) return;
	const uint64_t base = (uint64_t)tile * RUN_TILE;
	uint32_t bkt[RUN_ITEMS], heads[RUN_ITEMS], nvalid[RUN_ITEMS];
#pragma unroll
	for (int i = 0; i < RUN_ITEMS; ++i) {
		const uint64_t idx = base + (uint64_t)i * RUN_BLOCK + threadIdx.x;
		const uint64_t row0 = idx - lane; // first fragment of this row of 32
		nvalid[i] = row0 >= n ? 0u : (n - row0 >= 32u ? 32u : (uint32_t)(n - row0));
		const bool valid = (uint32_t)lane < nvalid[i];
		bkt[i] = valid ? (uint32_t)(frags[idx] >> bucket_shift) : 0u;
	}
#pragma unroll
	for (int i = 0; i < RUN_ITEMS; ++i) {
		const uint32_t prev = __shfl_up_sync(FULL_MASK, bkt[i], 1);
		const bool valid = (uint32_t)lane < nvalid[i];
		heads[i] = __ballot_sync(FULL_MASK, valid && (lane == 0 || bkt[i] != prev));
		if (lane == 0) s_cnt[i * RUN_NW + warp] = (uint32_t)__popc(heads[i]);
	}
	__syncthreads();
	// warp 0: exclusive scan of the 128 (row, warp) counts (stream order = row-major), then the tile's place in the list
	if (warp == 0) {
		uint32_t c[4], sum = 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) sum += (c[q] = s_cnt[4 * lane + q]);
		const uint32_t inc = warp_inclusive_sum(sum, lane);
		uint32_t run = inc - sum;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			s_cnt[4 * lane + q] = run;
			run += c[q];
		}
		const uint32_t total = __shfl_sync(FULL_MASK, inc, 31);
		const uint64_t p = lookback_exclusive(state, tile, (uint64_t)total, lane);
		if (lane == 0) {
			s_prefix = p;
			if (tile == n_tiles - 1) { // the last tile knows the number of runs
				const uint64_t all = p + total;
				ctl->n_runs = all > rec_cap ? rec_cap : (uint32_t)all;
				if (all > rec_cap) ctl->mode = 1u; // incoherent input: sorting the fragments themselves is cheaper
			}
		}
	}
	__syncthreads();
	const uint64_t tile_first = s_prefix;
#pragma unroll
	for (int i = 0; i < RUN_ITEMS; ++i) {
		const uint32_t m = heads[i];
		if (!((m >> lane) & 1u)) continue;
		const uint32_t rest = lane == 31 ? 0u : (m >> (lane + 1));
		const uint32_t end = rest ? (uint32_t)lane + (uint32_t)__ffs((int)rest) : nvalid[i]; // next head, or the end of the row
		const uint64_t slot = tile_first + s_cnt[i * RUN_NW + warp] + (uint32_t)__popc(m & lt_mask);
		const uint64_t idx = base + (uint64_t)i * RUN_BLOCK + threadIdx.x;
		if (slot < rec_cap)
			rec[slot] = ((uint64_t)bkt[i] << REC_BUCKET_SHIFT) | (idx << REC_START_SHIFT) | (uint64_t)(end - (uint32_t)lane - 1u);
	}
}

// ---- sorted runs -> fragment offsets, bucket starts ------------------------------------------------------------------
// roff[r] = fragments of the runs in front of sorted run r (roff[n_runs] = all fragments); for every run that starts a
// bucket: head_pos[h] = roff[r], head_run[h] = r (h numbers the non-empty buckets; entry n_heads closes the list).
// One chained scan of (starts-a-bucket << 32 | length): both sums stay far below the look-back word's 62 value bits
// (lengths sum to < 2^32 fragments, heads to < 2^30).
constexpr int RS_BLOCK = 256, RS_ITEMS = 8, RS_TILE = RS_BLOCK * RS_ITEMS;
__global__ void __launch_bounds__(RS_BLOCK)
    k_run_scan(const uint64_t *__restrict__ rec, uint32_t *__restrict__ roff, uint32_t *__restrict__ head_pos,
               uint32_t *__restrict__ head_run, uint64_t *state, BucketCtl *ctl) {
	__shared__ uint64_t s_warp[RS_BLOCK / 32 + 1];
	__shared__ uint32_t s_ticket;
	__shared__ uint64_t s_prefix;
	if (ctl->mode != 0u) return;
	const uint32_t n = ctl->n_runs;
	const uint32_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
	if (n == 0) {
		if (blockIdx.x == 0 && threadIdx.x == 0) {
			ctl->n_heads = 0;
			roff[0] = 0, head_pos[0] = 0, head_run[0] = 0;
		}
		return;
	}
	for (;;) {
		const uint32_t tile = take_ticket(&ctl->ticket[1], &s_ticket);
		if (tile >= n_tiles) break;
		const uint32_t base = tile * RS_TILE + threadIdx.x * RS_ITEMS;
		uint64_t v[RS_ITEMS];
		uint64_t sum = 0;
		uint32_t prev = (base > 0 && base < n) ? rec_bucket(rec[base - 1]) : 0xffffffffu; // (no bucket id reaches 2^32 - 1)
#pragma unroll
		for (int i = 0; i < RS_ITEMS; ++i) {
			v[i] = 0;
			if (base + i < n) {
				const uint64_t r = rec[base + i];
				const uint32_t b = rec_bucket(r);
				v[i] = ((uint64_t)(b != prev ? 1u : 0u) << 32) | rec_len(r);
				prev = b;
			}
			sum += v[i];
		}
		uint64_t total;
		const uint64_t excl = block_exclusive_sum<RS_BLOCK, uint64_t>(sum, total, s_warp);
		if (threadIdx.x < 32) {
			const uint64_t p = lookback_exclusive(state, tile, total, threadIdx.x);
			if (threadIdx.x == 0) s_prefix = p;
		}
		__syncthreads();
		uint64_t run = s_prefix + excl;
#pragma unroll
		for (int i = 0; i < RS_ITEMS; ++i) {
			if (base + i < n) {
				roff[base + i] = (uint32_t)run;
				if (v[i] >> 32) {
					const uint32_t h = (uint32_t)(run >> 32);
					head_pos[h] = (uint32_t)run;
					head_run[h] = base + i;
				}
			}
			run += v[i];
		}
		if (base <= n - 1 && n - 1 < base + RS_ITEMS) { // the thread that owns the last run closes both lists
			const uint32_t h = (uint32_t)(run >> 32);
			roff[n] = (uint32_t)run;
			head_pos[h] = (uint32_t)run, head_run[h] = n;
			ctl->n_heads = h;
		}
		__syncthreads();
	}
}

// ---- segments ----------------------------------------------------------------------------------------------------------
// The bucket-sorted order is cut at multiples of BS_T fragments, each cut moved up to the next bucket start: cell i owns
// the buckets that START in [i*BS_T, (i+1)*BS_T).  cell_pos[i] / cell_run[i] = first fragment offset / sorted run of cell i.
// A cell's last bucket may reach far beyond the cell: when that makes the cell larger than a block can hold, the block
// takes that bucket as a segment of its own (last_pos / last_run = where it starts); a bucket above BS_CAP fragments
// cannot be sorted on chip at all -> classic path.
#ifndef SVO_BS_BLOCK
#define SVO_BS_BLOCK 512
#endif
#ifndef SVO_BS_ITEMS
#define SVO_BS_ITEMS 16
#endif
constexpr int BS_BLOCK = SVO_BS_BLOCK, BS_ITEMS = SVO_BS_ITEMS, BS_CAP = BS_BLOCK * BS_ITEMS, BS_T = BS_CAP / 2, BS_NW = BS_BLOCK / 32;
struct CellPlan {
	uint32_t *cell_pos, *cell_run; // n_cells + 1 entries
	uint32_t *last_pos, *last_run; // n_cells entries, valid where the cell exceeds BS_CAP
	uint32_t n_cells;              // n_frag / BS_T + 1
};
__global__ void __launch_bounds__(256)
    k_bucket_plan(const uint32_t *__restrict__ head_pos, const uint32_t *__restrict__ head_run, CellPlan cp, BucketCtl *ctl) {
	if (ctl->mode != 0u) return;
	const uint32_t n_heads = ctl->n_heads;
	for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < n_heads; k += gridDim.x * 256) {
		const uint32_t s = head_pos[k], e = head_pos[k + 1];
		if (e - s > (uint32_t)BS_CAP) {
			ctl->mode = 1u;
			continue;
		}
		const uint32_t c0 = s / BS_T, c1 = e / BS_T;
		for (uint32_t i = c0 + 1; i <= c1; ++i) cp.cell_pos[i] = e, cp.cell_run[i] = head_run[k + 1]; // at most 2 cells
		if (c1 > c0) cp.last_pos[c0] = s, cp.last_run[c0] = head_run[k];
		if (k == 0) cp.cell_pos[0] = 0, cp.cell_run[0] = 0;
		if (k == n_heads - 1) cp.cell_pos[cp.n_cells] = e, cp.cell_run[cp.n_cells] = head_run[k + 1];
	}
	if (n_heads == 0 && blockIdx.x == 0 && threadIdx.x == 0) // no fragment at all
		for (uint32_t i = 0; i <= cp.n_cells; ++i) cp.cell_pos[i] = 0, cp.cell_run[i] = 0;
}

// ---- the segment kernel ----------------------------------------------------------------------------------------------
constexpr int BS_RBITS = 9, BS_RADIX = 1 << BS_RBITS, BS_NB = BS_RADIX + 1;
static_assert(BS_RADIX == BS_BLOCK, "one digit per thread in the digit scans");
struct BsSmem {
	static constexpr size_t OFF_KEYS = 0;                                      // BS_CAP keys as two 32-bit halves
	static constexpr size_t OFF_HIST = (size_t)BS_CAP * 8;                     // BS_NW * BS_NB counters; the reduce reuses it (run starts)
	static constexpr size_t HIST_BYTES = (((size_t)BS_NW * BS_NB * 4) + 15) & ~size_t(15);
	static constexpr size_t OFF_TOFF = OFF_HIST + HIST_BYTES;                  // BS_NB
	static constexpr size_t OFF_CNT = OFF_TOFF + (((size_t)BS_NB * 4 + 15) & ~size_t(15)); // 3 * BS_ITEMS * BS_NW
	static constexpr size_t BYTES = OFF_CNT + (size_t)3 * BS_ITEMS * BS_NW * 4;
	static_assert(HIST_BYTES >= (size_t)BS_CAP * 2, "the run-start list (16-bit) reuses the counter area");
};
static_assert(BS_ITEMS * BS_NW == 256, "the (row, warp) count matrix of the reduce is scanned by one warp, 8 per lane");

// One stable LSD pass over the block's keys on `nbits` (<= 9) bits from `shift`: ranks by warp votes, prefix over the warps,
// scatter through shared memory.  key[] holds the thread's elements (warp w owns [w*ipl*32, (w+1)*ipl*32), item i of lane
// l is element i*32 + l of that range), n of them are real.  reload = read the permuted keys back into key[].
SVO_DEV void bs_sort_pass(uint64_t (&key)[BS_ITEMS], uint32_t n, uint32_t ipl, uint32_t shift, uint32_t nbits, bool reload,
                          uint32_t *s_lo, uint32_t *s_hi, uint32_t *s_hist, uint32_t *s_tile_off, uint32_t *s_wsum) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t wbase = (uint32_t)warp * ipl * 32u;
	const uint32_t mask = (1u << nbits) - 1u, lt_mask = (1u << lane) - 1u;
	for (int i = threadIdx.x; i < BS_NW * BS_NB; i += BS_BLOCK) s_hist[i] = 0;
	__syncthreads();
	uint32_t *wh = s_hist + warp * BS_NB;
	uint32_t rank[BS_ITEMS];
#pragma unroll
	for (int i = 0; i < BS_ITEMS; ++i) {
		rank[i] = 0;
		if ((uint32_t)i >= ipl) continue; // block-uniform
		const uint32_t row = wbase + (uint32_t)i * 32u; // first element of this row
		const uint32_t vmask = row >= n ? 0u : (n - row >= 32u ? FULL_MASK : ((1u << (n - row)) - 1u));
		const bool valid = (vmask >> lane) & 1u;
		const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & mask) : 0u;
		const unsigned peers = warp_peers<BS_RBITS>(d) & vmask; // lanes without an element stay out of every group
		const uint32_t below = (uint32_t)__popc(peers & lt_mask);
		const uint32_t base = leader_atomic_add(&wh[d], (uint32_t)__popc(peers), valid && below == 0u);
		SVO_EMU_WARP_ORDER();
		rank[i] = __shfl_sync(FULL_MASK, base, peers ? __ffs((int)peers) - 1 : 0) + below;
	}
	__syncthreads();
	// one digit per thread: exclusive prefix over the warps, then over the digits
	{
		const uint32_t d = threadIdx.x;
		uint32_t run = 0;
#pragma unroll
		for (int w = 0; w < BS_NW; ++w) {
			const uint32_t c = s_hist[w * BS_NB + d];
			s_hist[w * BS_NB + d] = run;
			run += c;
		}
		const uint32_t inc = warp_inclusive_sum(run, lane);
		if (lane == 31) s_wsum[warp] = inc;
		__syncthreads();
		uint32_t wpre = 0;
#pragma unroll
		for (int w = 0; w < BS_NW; ++w) wpre += w < warp ? s_wsum[w] : 0u;
		s_tile_off[d] = wpre + inc - run;
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < BS_ITEMS; ++i) {
		if ((uint32_t)i >= ipl) continue;
		if (wbase + (uint32_t)i * 32u + lane >= n) continue;
		const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
		const uint32_t pos = s_tile_off[d] + s_hist[warp * BS_NB + d] + rank[i];
		s_lo[pos] = (uint32_t)key[i];
		s_hi[pos] = (uint32_t)(key[i] >> 32);
	}
	__syncthreads();
	if (reload) {
#pragma unroll
		for (int i = 0; i < BS_ITEMS; ++i) {
			const uint32_t e = wbase + (uint32_t)i * 32u + lane;
			if ((uint32_t)i < ipl && e < n) key[i] = (uint64_t)s_lo[e] | ((uint64_t)s_hi[e] << 32);
		}
	}
}

struct BucketSortArgs {
	const uint64_t *frags;   // the fragment list (stream order)
	const uint64_t *rec;     // sorted run records
	const uint32_t *roff;    // fragment offset of every sorted run
	CellPlan cp;
	uint32_t low_bits;       // Morton bits below the bucket id
	uint32_t *bucket_tab;    // gridDim.x * BS_CAP words: per block, the bucket id of every bucket of the segment in hand
	uint64_t *state;         // 3 * (2 * n_cells) look-back words (leaf / parent / grandparent counters), zeroed
	FusedOut out;
	BucketCtl *ctl;
};

// All of a segment: gather, sort, reduce.  K = 3 granularities (levels >= 3 take this path).
__global__ void __launch_bounds__(BS_BLOCK, 2) k_bucket_sort_reduce(BucketSortArgs a) {
	SVO_DYN_SMEM(unsigned char, smem);
	uint32_t *s_lo = reinterpret_cast<uint32_t *>(smem + BsSmem::OFF_KEYS), *s_hi = s_lo + BS_CAP;
	uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + BsSmem::OFF_HIST);
	uint16_t *s_start = reinterpret_cast<uint16_t *>(smem + BsSmem::OFF_HIST); // reduce: first element of every leaf run
	uint32_t *s_tile_off = reinterpret_cast<uint32_t *>(smem + BsSmem::OFF_TOFF);
	uint32_t(*s_cnt)[BS_ITEMS * BS_NW] = reinterpret_cast<uint32_t(*)[BS_ITEMS * BS_NW]>(smem + BsSmem::OFF_CNT);
	__shared__ uint32_t s_wsum[BS_NW + 1];
	__shared__ uint32_t s_total[3];
	__shared__ uint64_t s_pre[3];
	__shared__ uint32_t s_ticket, s_nbuckets;
	if (a.ctl->mode != 0u) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	// persistent blocks: cells are drawn from a ticket (start order: the look-back's predecessors are always running)
	for (;;) {
	const uint32_t cell = take_ticket(&a.ctl->ticket[2], &s_ticket);
	if (cell >= a.cp.n_cells) return;
	const uint32_t n_vt = 2u * a.cp.n_cells; // look-back tiles: two per cell
	uint64_t *st0 = a.state, *st1 = a.state + n_vt, *st2 = a.state + 2 * (size_t)n_vt;
	const uint32_t pos0 = a.cp.cell_pos[cell], pos1 = a.cp.cell_pos[cell + 1];
	const uint32_t run0 = a.cp.cell_run[cell], run1 = a.cp.cell_run[cell + 1];
	const bool split = pos1 - pos0 > (uint32_t)BS_CAP; // the cell's last bucket is taken separately
	const uint32_t mid_pos = split ? a.cp.last_pos[cell] : pos1, mid_run = split ? a.cp.last_run[cell] : run1;
	uint32_t *my_tab = a.bucket_tab + (size_t)blockIdx.x * BS_CAP;
	const uint32_t bshift = 24u + a.low_bits;             // a key's bucket id = key >> bshift
	const uint64_t low_mask = (1ull << bshift) - 1ull;

	for (uint32_t part = 0; part < 2u; ++part) {
		const uint32_t vt = 2u * cell + part;
		const uint32_t p_lo = part == 0 ? pos0 : mid_pos, p_hi = part == 0 ? mid_pos : pos1;
		const uint32_t r_lo = part == 0 ? run0 : mid_run, r_hi = part == 0 ? mid_run : run1;
		const uint32_t n = p_hi - p_lo;
		const bool last_vt = vt == n_vt - 1u;
		if (n == 0u) { // nothing here: pass the counters on (three warps, one look-back each)
			if (warp < 3) {
				uint64_t *st = warp == 0 ? st0 : (warp == 1 ? st1 : st2);
				const uint64_t p = lookback_exclusive(st, vt, 0ull, lane);
				if (last_vt && lane == 0) *a.out.count[warp] = p;
			}
			continue;
		}

		// ---- gather: warps take the segment's runs in turn; a run is a contiguous piece of the fragment list
		for (uint32_t r = r_lo + warp; r < r_hi; r += BS_NW) {
			const uint64_t rc = a.rec[r];
			const uint32_t dst = a.roff[r] - p_lo, len = rec_len(rc);
			if ((uint32_t)lane < len) {
				const uint64_t k = a.frags[(uint64_t)rec_start(rc) + lane];
				s_lo[dst + lane] = (uint32_t)k, s_hi[dst + lane] = (uint32_t)(k >> 32);
			}
		}
		__syncthreads();

		// ---- to registers; number the buckets of the segment (its keys are sorted by bucket already)
		const uint32_t ipl = (n + BS_BLOCK - 1) / BS_BLOCK; // items per lane, 1..BS_ITEMS
		const uint32_t wbase = (uint32_t)warp * ipl * 32u;
		uint64_t key[BS_ITEMS];
		uint32_t bflag = 0; // bit i: element i of this thread starts a bucket
		uint32_t wcount = 0;
#pragma unroll
		for (int i = 0; i < BS_ITEMS; ++i) {
			key[i] = 0;
			const uint32_t e = wbase + (uint32_t)i * 32u + lane;
			bool head = false;
			if ((uint32_t)i < ipl && e < n) {
				key[i] = (uint64_t)s_lo[e] | ((uint64_t)s_hi[e] << 32);
				head = e > 0u && (s_hi[e - 1] >> (bshift - 32u)) != (uint32_t)(key[i] >> bshift); // (bshift >= 36)
			}
			const unsigned b = __ballot_sync(FULL_MASK, head);
			bflag |= (head ? 1u : 0u) << i;
			wcount += (uint32_t)__popc(b);
		}
		if (lane == 0) s_wsum[warp] = wcount;
		__syncthreads();
		uint32_t brank = 0; // bucket number of the element in hand, running over the warp's rows
#pragma unroll
		for (int w = 0; w < BS_NW; ++w) brank += w < warp ? s_wsum[w] : 0u;
		if (threadIdx.x == BS_BLOCK - 1) s_nbuckets = brank + wcount + 1u;
#pragma unroll
		for (int i = 0; i < BS_ITEMS; ++i) {
			const bool head = (bflag >> i) & 1u;
			const unsigned b = __ballot_sync(FULL_MASK, head);
			const uint32_t mine = brank + (uint32_t)__popc(b & (lt_mask | (1u << lane))); // heads up to and including this lane
			brank += (uint32_t)__popc(b);
			const uint32_t e = wbase + (uint32_t)i * 32u + lane;
			if ((uint32_t)i < ipl && e < n) {
				if (head || e == 0u) my_tab[mine] = (uint32_t)(key[i] >> bshift);
				key[i] = (key[i] & low_mask) | ((uint64_t)mine << bshift); // bucket id -> bucket number (few bits)
			}
		}
		__syncthreads();

		// ---- sort by (bucket number, low Morton bits): stable LSD passes of <= 9 bits
		const uint32_t nb = s_nbuckets;
		const uint32_t kbits = nb > 1u ? 32u - (uint32_t)__clz((int)(nb - 1u)) : 0u;
		const uint32_t total_bits = a.low_bits + kbits;
		const uint32_t n_pass = (total_bits + BS_RBITS - 1) / BS_RBITS;
		uint32_t sh = 24u;
		for (uint32_t p = 0; p < n_pass; ++p) {
			const uint32_t left = 24u + total_bits - sh, wbits = (left + (n_pass - p) - 1) / (n_pass - p);
			bs_sort_pass(key, n, ipl, sh, wbits, p + 1 < n_pass, s_lo, s_hi, s_hist, s_tile_off, s_wsum);
			sh += wbits;
		}
		// (the sorted keys are in shared memory; elements are block-striped from here on: element = row * BS_BLOCK + tid)

		// ---- reduce: run starts at the three granularities + in-warp ranks (as k_reduce_fused)
		const uint32_t rows = ipl; // rows of BS_BLOCK elements
		uint32_t packed[BS_ITEMS];
#pragma unroll
		for (int i = 0; i < BS_ITEMS; ++i) {
			packed[i] = 0;
			if ((uint32_t)i >= rows) continue; // block-uniform
			const uint32_t e = (uint32_t)i * BS_BLOCK + threadIdx.x;
			uint64_t x = 0;
			if (e < n) {
				const uint64_t k = (uint64_t)s_lo[e] | ((uint64_t)s_hi[e] << 32);
				x = e == 0u ? ~0ull : (k ^ ((uint64_t)s_lo[e - 1] | ((uint64_t)s_hi[e - 1] << 32)));
			}
			uint32_t pk = 0;
#pragma unroll
			for (int j = 0; j < 3; ++j) {
				const bool f = (x >> (24 + 3 * j)) != 0;
				const unsigned b = __ballot_sync(FULL_MASK, f);
				pk |= (f ? 1u : 0u) << j;
				pk |= (uint32_t)__popc(b & lt_mask) << (3 + 5 * j);
				if (lane == 0) s_cnt[j][i * BS_NW + warp] = (uint32_t)__popc(b);
			}
			packed[i] = pk;
		}
		__syncthreads();
		// warps 0..2: exclusive scan of the (row, warp) counts of one granularity, then that counter's look-back
		if (warp < 3) {
			const int j = warp;
			constexpr int CPL = BS_ITEMS * BS_NW / 32;
			uint32_t c[CPL], sum = 0;
#pragma unroll
			for (int q = 0; q < CPL; ++q) {
				const uint32_t idx = CPL * lane + q;
				c[q] = idx / BS_NW < rows ? s_cnt[j][idx] : 0u;
				sum += c[q];
			}
			const uint32_t inc = warp_inclusive_sum(sum, lane);
			uint32_t run = inc - sum;
#pragma unroll
			for (int q = 0; q < CPL; ++q) {
				s_cnt[j][CPL * lane + q] = run;
				run += c[q];
			}
			const uint32_t total = __shfl_sync(FULL_MASK, inc, 31);
			uint64_t *st = j == 0 ? st0 : (j == 1 ? st1 : st2);
			const uint64_t p = lookback_exclusive(st, vt, (uint64_t)total, lane);
			if (lane == 0) {
				s_total[j] = total, s_pre[j] = p;
				if (last_vt) *a.out.count[j] = p + total;
			}
		}
		__syncthreads();
		const uint64_t p0 = s_pre[0], p1 = s_pre[1], p2 = s_pre[2];
		const uint32_t n_leaf = s_total[0];
		const bool listed = n_leaf * 4u <= n * 3u; // block-uniform: many fragments per voxel -> deal the runs out, one per thread
		auto key_at = [&](uint32_t e) { return (uint64_t)s_lo[e] | ((uint64_t)s_hi[e] << 32); };
		auto fold_leaf = [&](uint32_t e, uint64_t k) { // the reference's running average over the voxel's fragments, in list order
			uint32_t acc = leaf_first((uint32_t)(k & 0xffffffu));
			for (uint32_t q = e + 1; q < n; ++q) {
				const uint64_t kk = key_at(q);
				if ((kk >> 24) != (k >> 24)) break;
				acc = leaf_accumulate(acc, (uint32_t)(kk & 0xffffffu));
			}
			return acc;
		};
#pragma unroll
		for (int i = 0; i < BS_ITEMS; ++i) {
			const uint32_t pk = packed[i];
			if (!(pk & 1u)) continue;
			const uint32_t e = (uint32_t)i * BS_BLOCK + threadIdx.x;
			const uint64_t k = key_at(e);
			const uint32_t l0 = s_cnt[0][i * BS_NW + warp] + ((pk >> 3) & 31u); // index of the run inside the segment
			const uint64_t u0 = p0 + l0;
			if (!listed) {
				a.out.leaf[u0] = fold_leaf(e, k);
				a.out.slot0[u0] = (unsigned char)((k >> 24) & 7u);
			}
			if (pk & 2u) {
				const uint64_t u1 = p1 + s_cnt[1][i * BS_NW + warp] + ((pk >> 8) & 31u);
				a.out.first1[u1] = (uint32_t)u0;
				a.out.slot1[u1] = (unsigned char)((k >> 27) & 7u);
				if (pk & 4u) {
					const uint64_t u2 = p2 + s_cnt[2][i * BS_NW + warp] + ((pk >> 13) & 31u);
					a.out.first2[u2] = (uint32_t)u1;
					// the real key again: bucket number -> bucket id
					const uint64_t real = (k & low_mask) | ((uint64_t)my_tab[(uint32_t)(k >> bshift)] << bshift);
					a.out.keys_top[u2] = real >> 30;
				}
			}
		}
		if (listed) {
			__syncthreads(); // everybody is done with the counters' area before the run-start list overwrites... (s_cnt is separate; s_hist is free)
#pragma unroll
			for (int i = 0; i < BS_ITEMS; ++i) {
				const uint32_t pk = packed[i];
				if (pk & 1u) s_start[s_cnt[0][i * BS_NW + warp] + ((pk >> 3) & 31u)] = (uint16_t)((uint32_t)i * BS_BLOCK + threadIdx.x);
			}
			__syncthreads();
			for (uint32_t r = threadIdx.x; r < n_leaf; r += BS_BLOCK) {
				const uint32_t e = s_start[r];
				const uint64_t k = key_at(e);
				a.out.leaf[p0 + r] = fold_leaf(e, k);
				a.out.slot0[p0 + r] = (unsigned char)((k >> 24) & 7u);
			}
		}
		__syncthreads(); // the next part reuses the shared memory
	}
	}
}

} // namespace svo
