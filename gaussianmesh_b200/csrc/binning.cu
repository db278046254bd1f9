#include <cstdlib>
// Tile binning without a global sort.
//
// The reference builds 64-bit (tile | depth) keys for every (Gaussian, tile) instance, radix-sorts
// all R of them over 41-45 bits and then searches the sorted keys for tile boundaries
// (dgr/cuda_rasterizer/rasterizer_impl.cu:70-138, 464-495).  Here neither the tile nor the coarse depth
// order ever enters a sort:
//   * preprocess counted instances per (tile, depth bucket) -- the bucket is a monotone function of depth
//     (state.h), so bucket order is depth order;
//   * `tile_scan` turns the counts into segment offsets (tile segments 4-aligned, buckets packed inside);
//   * `emit` drops each instance into its (tile, bucket) range;
//   * `sort_pack` sorts every bucket by (depth, gaussian id): buckets of up to 128 instances in REGISTERS by
//     one warp (bitonic network over shuffles, no shared memory, no barriers), larger ones by the whole
//     block in shared memory; then gathers the splat data and writes the packed, tile-contiguous records
//     that the blend kernels stage with bulk copies.
//
// Ordering contract (what the blend result depends on): within a tile, ascending depth bits, ties
// broken by ascending Gaussian id -- identical to a stable radix sort of keys emitted in id order
// (SURVEY.md 7 "sort tie order").
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

// ------------------------------------------------------------------------------------------------
// tile_scan: one thread per tile.  total = sum of the tile's bucket counts; a chained scan over blocks
// (ticketed, so a block only ever waits for blocks that have already started) of the 4-aligned totals
// gives tile_start; inside a tile the bucket counts become bucket STARTS (bucket_cursor), which emit
// advances to bucket ENDS.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = kScanTiles;
constexpr unsigned long long kFlagAggregate = 1ull << 32, kFlagInclusive = 2ull << 32;

__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int num_tiles, int bucket_log2, GeometryState g, uint32_t capacity)
{
	pdl_sync();
	__shared__ uint32_t warp_sums[kScanThreads / 32];
	__shared__ uint32_t s_block, s_base;
	const int tid = threadIdx.x;
	unsigned long long* const ticket = g.scan_state + kMaxScanBlocks;
	if (tid == 0)
		s_block = (uint32_t)atomicAdd(ticket, 1ull);
	__syncthreads();
	const uint32_t blk = s_block;
	const int t = (int)blk * kScanTiles + tid;
	const int B = 1 << bucket_log2;

	uint32_t cnt[kMaxBuckets];
	uint32_t total = 0;
	if (t < num_tiles) {
		const uint32_t* c = g.bucket_cursor + ((size_t)t << bucket_log2);
#pragma unroll
		for (int b = 0; b < kMaxBuckets; b++) {
			cnt[b] = (b < B) ? c[b] : 0u;
			total += cnt[b];
		}
	}
	const uint32_t aligned = (total + (kSegAlign - 1)) & ~(uint32_t)(kSegAlign - 1);

	// block-wide scan of `aligned`
	uint32_t incl = aligned;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if ((tid & 31) >= o) incl += v;
	}
	if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
	__syncthreads();
	uint32_t warp_off = 0, block_total = 0;
#pragma unroll
	for (int w = 0; w < kScanThreads / 32; w++) {
		if (w < (tid >> 5)) warp_off += warp_sums[w];
		block_total += warp_sums[w];
	}

	// chained scan across blocks
	if (tid == 0) {
		volatile unsigned long long* state = g.scan_state;
		uint32_t prefix = 0;
		if (blk > 0) {
			state[blk] = kFlagAggregate | block_total;
			__threadfence();
			int j = (int)blk - 1;
			while (true) {
				const unsigned long long v = state[j];
				if (v & kFlagInclusive) { prefix += (uint32_t)v; break; }
				if (v & kFlagAggregate) { prefix += (uint32_t)v; j--; }
			}
		}
		state[blk] = kFlagInclusive | (unsigned long long)(prefix + block_total);
		__threadfence();
		s_base = prefix;
		if (blk == gridDim.x - 1) {
			const uint32_t run = prefix + block_total;
			g.header->num_rendered = run;
			g.header->capacity = capacity;
			g.header->overflow = (run > capacity) ? 1u : 0u;
			g.header->num_tiles = (uint32_t)num_tiles;
			g.header->bucket_log2 = (uint32_t)bucket_log2;
		}
	}
	__syncthreads();

	if (t < num_tiles) {
		const uint32_t start = s_base + warp_off + incl - aligned;
		g.tile_start[t] = start;
		g.tile_count[t] = total;
		uint32_t* c = g.bucket_cursor + ((size_t)t << bucket_log2);
		uint32_t at = start;
#pragma unroll
		for (int b = 0; b < kMaxBuckets; b++) {
			if (b < B) {
				c[b] = at;
				at += cnt[b];
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// emit: one thread per Gaussian, same rectangle walk, the same culling predicate and the same bucket as
// the count in preprocess (so counts and emitted instances agree exactly).  Slot order inside a bucket
// is arbitrary; sort_pack makes it deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
emit_kernel(int P, const int* __restrict__ radii, GeometryState g, BinningState b, uint32_t capacity,
            int W, int H, int tiles_x, int tiles_y, int bucket_log2)
{
	pdl_sync();
	const int idx = blockIdx.x * kEmitThreads + threadIdx.x;
	if (idx >= P)
		return;
	const int radius = radii[idx];
	if (!(radius > 0))
		return;
	// kept-tile mask recorded by preprocess (zero for culled splats and for rectangles of more than 64 tiles,
	// which large_tiles_kernel places)
	unsigned long long kept = g.tile_mask[idx];
	if (kept == 0ull)
		return;
	const float2 xy = g.means2D[idx];
	int x0, y0, x1, y1;
	tile_rect(xy, radius, tiles_x, tiles_y, x0, y0, x1, y1);
	const float depth = g.depths[idx];
	const uint64_t key = ((uint64_t)__float_as_uint(depth) << 32) | (uint32_t)idx;
	const uint32_t bucket = g.depth_lut[depth_fine_bin(depth)];
	uint32_t* const cursor = g.bucket_cursor + bucket;

	{
		// replay the mask, four instances at a time: the four slot atomics are in flight
		// together, so a Gaussian costs ceil(n/4) memory round trips instead of n
		const uint32_t w = (uint32_t)(x1 - x0);
		const uint32_t inv_w = (65536u + w - 1u) / w;       // (bit * inv_w) >> 16 == bit / w for bit < 64, w <= 64
		while (kept) {
			uint32_t tiles[4], pos[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				tiles[u] = 0xffffffffu;
				if (kept) {
					const uint32_t bit = (uint32_t)__ffsll((long long)kept) - 1u;
					kept &= kept - 1ull;
					const uint32_t row = (bit * inv_w) >> 16;
					tiles[u] = (uint32_t)(y0 + (int)row) * (uint32_t)tiles_x + (uint32_t)x0 + (bit - row * w);
				}
			}
#pragma unroll
			for (int u = 0; u < 4; u++)
				if (tiles[u] != 0xffffffffu)
					pos[u] = atomicAdd(&cursor[(size_t)tiles[u] << bucket_log2], 1u);
#pragma unroll
			for (int u = 0; u < 4; u++)
				if (tiles[u] != 0xffffffffu && pos[u] < capacity)
					b.keys[pos[u]] = key;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// large_tiles: Gaussians whose tile rectangle exceeds 64 tiles (close-up splats).  One warp per Gaussian,
// lanes stride over the rectangle.  The SAME kernel binary runs the count pass (phase 0, before tile_scan)
// and the placement pass (phase 1, with emit), so the culling predicate gives identical answers in both.
// ------------------------------------------------------------------------------------------------
constexpr int kLargeThreads = 256;

__global__ void __launch_bounds__(kLargeThreads)
large_tiles_kernel(int phase, const int* __restrict__ radii, GeometryState g, uint64_t* __restrict__ keys,
                   uint32_t capacity, int W, int H, int tiles_x, int tiles_y, int bucket_log2)
{
	pdl_sync();
	const int lane = threadIdx.x & 31;
	const uint32_t warps = gridDim.x * (kLargeThreads / 32);
	const uint32_t n = g.header->num_large;
	for (uint32_t e = blockIdx.x * (kLargeThreads / 32) + (threadIdx.x >> 5); e < n; e += warps) {
		const uint32_t idx = g.large_list[e];
		const float4 co = g.conic_opacity[idx];
		const float thr = cull_threshold(co.w);
		if (thr < 0.0f)
			continue;
		const float2 xy = g.means2D[idx];
		const float depth = g.depths[idx];
		const uint64_t key = ((uint64_t)__float_as_uint(depth) << 32) | idx;
		uint32_t* const cursor = g.bucket_cursor + g.depth_lut[depth_fine_bin(depth)];
		int x0, y0, x1, y1;
		tile_rect(xy, radii[idx], tiles_x, tiles_y, x0, y0, x1, y1);
		const int w = x1 - x0, area = w * (y1 - y0);
		for (int t = lane; t < area; t += 32) {
			const int ty = y0 + t / w, tx = x0 + t % w;
			const float px0 = (float)(tx * kTile), py0 = (float)(ty * kTile);
			const float px1 = fminf(px0 + (kTile - 1), (float)(W - 1));
			const float py1 = fminf(py0 + (kTile - 1), (float)(H - 1));
			if (rect_cannot_contribute(xy.x, xy.y, co.x, co.y, co.z, thr, px0, py0, px1, py1))
				continue;
			const uint32_t tile = (uint32_t)(ty * tiles_x + tx);
			const uint32_t pos = atomicAdd(&cursor[(size_t)tile << bucket_log2], 1u);
			if (phase == 1 && pos < capacity)
				keys[pos] = key;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// sort_pack: one block (8 warps) per tile.
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortSmem = 4096;       // keys; 32 KB -- block-wide path for oversized buckets
constexpr int kWarpSortMax = 256;     // largest bucket one warp sorts in registers (8 keys per lane)
constexpr int kMidKeys = 2048;        // largest bucket of the small-block class of big_bucket_sort_pack_kernel (16 KB of keys)

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int lane_mask)
{
	const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, lane_mask);
	const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), lane_mask);
	return ((uint64_t)hi << 32) | lo;
}

// Bitonic sort of 32*E keys held E per lane; element index e = r * 32 + lane (register r).
// Ascending in e on return.  Padding keys must be UINT64_MAX.
template <int E>
__device__ __forceinline__ void warp_bitonic(uint64_t (&key)[E], int lane)
{
#pragma unroll
	for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			if (j >= 32) {
				// partner lives in another register of the same lane; bit k of e depends only on r here
#pragma unroll
				for (int r = 0; r < E; r++) {
					const int pr = r ^ (j >> 5);
					if (pr > r) {
						const bool up = (((r << 5) & k) == 0);
						const uint64_t a = key[r], c = key[pr];
						const bool swap = ((a > c) == up);          // distinct keys: a < c is !(a > c)
						key[r] = swap ? c : a;
						key[pr] = swap ? a : c;
					}
				}
			} else {
#pragma unroll
				for (int r = 0; r < E; r++) {
					const int e = (r << 5) | lane;
					const bool up = ((e & k) == 0);
					const bool lower = ((lane & j) == 0);
					const uint64_t mine = key[r];
					const uint64_t other = shfl_xor_u64(mine, j);
					// keys are distinct (the Gaussian id is part of them; equal padding keys may go either way), so
					// "keep the smaller" is "keep mine iff mine < other": one 64-bit compare and one select
					const bool take_min = (lower == up);
					key[r] = ((mine < other) == take_min) ? mine : other;
				}
			}
		}
	}
}

__device__ __forceinline__ void pack_one(const GeometryState& g, const BinningState& b, uint32_t dst, uint32_t id)
{
	const float2 xy = g.means2D[id];
	const float4 co = g.conic_opacity[id];
	const float4 rgb = g.rgb_clamp[id];
	// streaming stores: the records are read once by the blend kernels, the per-Gaussian arrays stay in L2
	__stcs(&b.rec_conic[dst], co);
	__stcs(&b.rec_xyrg[dst], make_float4(xy.x, xy.y, rgb.x, rgb.y));
	__stcs(&b.rec_bid[dst], make_float2(rgb.z, __uint_as_float(id)));
}

// Sort the n <= 32*E keys at `src` (global or shared memory) in registers and write their packed records to list
// positions s0 .. s0+n-1.
template <int E>
__device__ __forceinline__ void warp_sort_pack(const GeometryState& g, const BinningState& b, const uint64_t* src,
                                               uint32_t s0, uint32_t n, int lane)
{
	uint64_t key[E];
#pragma unroll
	for (int r = 0; r < E; r++) {
		const uint32_t e = (uint32_t)(r << 5) | lane;
		key[r] = (e < n) ? src[e] : ~0ull;
	}
	warp_bitonic<E>(key, lane);
#pragma unroll
	for (int r = 0; r < E; r++) {
		const uint32_t e = (uint32_t)(r << 5) | lane;
		if (e < n)
			pack_one(g, b, s0 + e, (uint32_t)key[r]);
	}
}

template <typename Ptr>
__device__ __forceinline__ void compare_exchange(Ptr a, uint32_t i, uint32_t j, uint32_t n)
{
	if (j < n) {
		const uint64_t x = a[i], y = a[j];
		if (x > y) { a[i] = y; a[j] = x; }
	}
}

// Block-wide bitonic network in the "flip" formulation (every compare-exchange orders ascending), which
// needs no padding: a partner index >= n behaves as +infinity and the exchange is simply skipped.
// k and s are powers of two, so every index is built from shifts and masks (lk = log2 k, ls = log2 s).
template <typename Ptr>
__device__ __forceinline__ void block_bitonic(Ptr a, uint32_t n)
{
	uint32_t lm = 0;
	while ((1u << lm) < n) lm++;
	const uint32_t half = (1u << lm) >> 1;
	for (uint32_t lk = 1; lk <= lm; lk++) {
		const uint32_t hk_mask = (1u << (lk - 1)) - 1u;
		for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
			const uint32_t off = t & hk_mask;
			const uint32_t blk = (t >> (lk - 1)) << lk;
			compare_exchange(a, blk + off, blk + ((1u << lk) - 1u - off), n);
		}
		__syncthreads();
		for (int ls = (int)lk - 2; ls >= 0; ls--) {
			const uint32_t s_mask = (1u << ls) - 1u;
			for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
				const uint32_t i = ((t >> ls) << (ls + 1)) + (t & s_mask);
				compare_exchange(a, i, i + (1u << ls), n);
			}
			__syncthreads();
		}
	}
}

// One WARP per (tile, bucket): buckets of up to kWarpSortMax instances are sorted in registers and packed with
// no shared memory and no block barrier, so every warp is an independent latency chain.
__global__ void __launch_bounds__(kSortThreads)
bucket_sort_pack_kernel(GeometryState g, BinningState b, uint32_t capacity, int bucket_log2, uint32_t num_buckets_total)
{
	pdl_sync();
	const int lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * kSortWarps + (threadIdx.x >> 5);
	if (gw >= num_buckets_total)
		return;
	const uint32_t tile = gw >> bucket_log2;
	const uint32_t bk = gw & ((1u << bucket_log2) - 1u);
	// bucket bounds: emit left bucket_cursor at the END of each bucket
	const uint32_t s0 = (bk == 0) ? g.tile_start[tile] : g.bucket_cursor[gw - 1];
	const uint32_t s1 = min(g.bucket_cursor[gw], capacity);
	if (s0 >= s1)
		return;
	const uint32_t n = s1 - s0;
	if (n <= 32) warp_sort_pack<1>(g, b, b.keys + s0, s0, n, lane);
	else if (n <= 64) warp_sort_pack<2>(g, b, b.keys + s0, s0, n, lane);
	else if (n <= 128) warp_sort_pack<4>(g, b, b.keys + s0, s0, n, lane);
	else if (n <= kWarpSortMax) warp_sort_pack<8>(g, b, b.keys + s0, s0, n, lane);
	else if (lane == 0) {
		// left to big_bucket_sort_pack_kernel: two size classes, two lists in one array (the second grows from the end)
		if (n <= (uint32_t)kMidKeys) g.big_list[atomicAdd(&g.header->num_big, 1u)] = gw;
		else g.big_list[kMaxBucketEntries - 1 - atomicAdd(&g.header->num_huge, 1u)] = gw;
	}
}

// The buckets the warp kernel skipped (more than kWarpSortMax instances), one BLOCK per bucket, taken from the
// list the warp kernel filled.  Gaussians bound to a surface put most of a tile's instances into one or two of the
// global depth buckets, so this path has to be as cheap per key as the warp path: the block splits the bucket into
// up to 128 SUB-BUCKETS by linear interpolation of the depth bits between the bucket's own minimum and maximum
// (monotone in depth, equal depths share a sub-bucket -- so concatenating sorted sub-buckets is the sorted bucket),
// scatters the keys through shared memory, and its eight warps sort the sub-buckets in registers exactly like
// bucket_sort_pack_kernel.  A sub-bucket that still exceeds kWarpSortMax (many equal depths) sends the whole
// bucket through the block-wide bitonic network; a bucket larger than shared memory is sorted in place in global
// memory (rare, slow, still exact).
constexpr int kSubMax = 128;          // sub-buckets per big bucket
constexpr int kSubTarget = 48;        // aimed-at keys per sub-bucket
constexpr size_t kBigSmemBytes = (size_t)kSortSmem * sizeof(uint64_t);

// kHuge = false: the buckets of up to kMidKeys keys, 128-thread blocks with 16 KB of shared memory (twelve per SM: the
// work is a chain of latencies per bucket, so buckets in flight are what counts); kHuge = true: the rest, 256 threads, 32 KB.
template <int kT, int kRegKeys, bool kHuge>
__global__ void __launch_bounds__(kT)
big_bucket_sort_pack_kernel(GeometryState g, BinningState b, uint32_t capacity, int bucket_log2, uint32_t smem_keys)
{
	pdl_sync();
	extern __shared__ __align__(16) unsigned char big_smem[];
	uint64_t* const s_part = reinterpret_cast<uint64_t*>(big_smem);      // [kSortSmem] keys partitioned by sub-bucket
	__shared__ uint32_t s_cnt[kSubMax], s_off[kSubMax + 1], s_lo[(kT / 32)], s_hi[(kT / 32)], s_fallback;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t num_big = kHuge ? g.header->num_huge : g.header->num_big;
	__shared__ uint32_t s_ticket;
	for (;;) {
		// the blocks draw buckets from a shared ticket (sizes vary by an order of magnitude: a static stride left the
		// grid waiting for its unluckiest block)
		if (threadIdx.x == 0)
			s_ticket = atomicAdd(kHuge ? &g.header->huge_ticket : &g.header->big_ticket, 1u);
		__syncthreads();
		const uint32_t e = s_ticket;
		if (e >= num_big)
			break;
		const uint32_t gw = g.big_list[kHuge ? kMaxBucketEntries - 1 - e : e];
		const uint32_t bk = gw & ((1u << bucket_log2) - 1u);
		const uint32_t s0 = (bk == 0) ? g.tile_start[gw >> bucket_log2] : g.bucket_cursor[gw - 1];
		const uint32_t n = min(g.bucket_cursor[gw], capacity) - s0;
		uint64_t* keys = b.keys + s0;
		if (n > smem_keys) {
			block_bitonic(keys, n);
			for (uint32_t i = threadIdx.x; i < n; i += kT)
				pack_one(g, b, s0 + i, (uint32_t)keys[i]);
			continue;
		}
		// The keys were just written by emit and sit in L2.  A bucket of up to kRegKeys keys per thread (1,024: nearly every
		// oversized bucket of a surface scene) is read ONCE into registers, all loads in flight together; the three passes
		// (depth range, sub-bucket histogram, scatter) then run out of registers instead of paying an L2 round trip each.
		const bool in_regs = n <= (uint32_t)(kT * kRegKeys);
		uint64_t kr[kRegKeys];
		if (in_regs) {
#pragma unroll
			for (int u = 0; u < kRegKeys; u++) {
				const uint32_t i = threadIdx.x + (uint32_t)u * kT;
				kr[u] = i < n ? keys[i] : ~0ull;
			}
		}
		uint32_t lo = 0xffffffffu, hi = 0u;
		if (in_regs) {
#pragma unroll
			for (int u = 0; u < kRegKeys; u++) {
				if (threadIdx.x + (uint32_t)u * kT < n) {
					const uint32_t d = (uint32_t)(kr[u] >> 32);
					lo = min(lo, d);
					hi = max(hi, d);
				}
			}
		} else {
			for (uint32_t i = threadIdx.x; i < n; i += kT) {
				const uint32_t d = (uint32_t)(keys[i] >> 32);
				lo = min(lo, d);
				hi = max(hi, d);
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
			hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
		}
		if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
		if (threadIdx.x < kSubMax) s_cnt[threadIdx.x] = 0u;
		if (threadIdx.x == 0) s_fallback = 0u;
		__syncthreads();
#pragma unroll
		for (int w = 0; w < (kT / 32); w++) {
			lo = min(lo, s_lo[w]);
			hi = max(hi, s_hi[w]);
		}
		uint32_t S = 1;
		while (S < (uint32_t)kSubMax && S * kSubTarget < n) S <<= 1;
		// float arithmetic is monotone (conversion, multiplication by a positive constant, truncation)
		const float scale = (float)S / ((float)(hi - lo) + 1.0f);
		auto sub_of = [&](uint64_t k) {
			return min(S - 1u, (uint32_t)((float)((uint32_t)(k >> 32) - lo) * scale));
		};
		if (in_regs) {
#pragma unroll
			for (int u = 0; u < kRegKeys; u++)
				if (threadIdx.x + (uint32_t)u * kT < n)
					atomicAdd(&s_cnt[sub_of(kr[u])], 1u);
		} else {
			for (uint32_t i = threadIdx.x; i < n; i += kT)
				atomicAdd(&s_cnt[sub_of(keys[i])], 1u);
		}
		__syncthreads();
		if (warp == 0) {
			// exclusive scan of up to 128 counts, four per lane
			uint32_t c[4], sum = 0;
#pragma unroll
			for (int u = 0; u < 4; u++) {
				c[u] = (uint32_t)(4 * lane + u) < S ? s_cnt[4 * lane + u] : 0u;
				sum += c[u];
			}
			uint32_t incl = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o) incl += v;
			}
			uint32_t at = incl - sum;
#pragma unroll
			for (int u = 0; u < 4; u++) {
				if ((uint32_t)(4 * lane + u) < S) {
					s_off[4 * lane + u] = at;
					s_cnt[4 * lane + u] = at;              // becomes the scatter cursor
				}
				at += c[u];
			}
			if (lane == 31) s_off[S] = at;
		}
		__syncthreads();
		if (in_regs) {
#pragma unroll
			for (int u = 0; u < kRegKeys; u++)
				if (threadIdx.x + (uint32_t)u * kT < n)
					s_part[atomicAdd(&s_cnt[sub_of(kr[u])], 1u)] = kr[u];
		} else {
			for (uint32_t i = threadIdx.x; i < n; i += kT) {
				const uint64_t k = keys[i];
				s_part[atomicAdd(&s_cnt[sub_of(k)], 1u)] = k;
			}
		}
		__syncthreads();
		for (uint32_t sb = warp; sb < S; sb += (kT / 32)) {
			const uint32_t o0 = s_off[sb], m = s_off[sb + 1] - o0;
			if (m == 0) continue;
			if (m <= 32) warp_sort_pack<1>(g, b, s_part + o0, s0 + o0, m, lane);
			else if (m <= 64) warp_sort_pack<2>(g, b, s_part + o0, s0 + o0, m, lane);
			else if (m <= 128) warp_sort_pack<4>(g, b, s_part + o0, s0 + o0, m, lane);
			else if (m <= kWarpSortMax) warp_sort_pack<8>(g, b, s_part + o0, s0 + o0, m, lane);
			else if (lane == 0) s_fallback = 1u;
		}
		__syncthreads();
		if (s_fallback) {
			// s_part is a permutation of the bucket; the total order is unique, so re-packing is idempotent
			block_bitonic(s_part, n);
			for (uint32_t i = threadIdx.x; i < n; i += kT)
				pack_one(g, b, s0 + i, (uint32_t)s_part[i]);
		}
		__syncthreads();   // the shared arrays are reused by the next bucket
	}
}

} // namespace

int launch_tile_scan(int num_tiles, const GeometryState& g, uint32_t capacity, const ViewParams& vp, cudaStream_t stream)
{
	if (num_tiles <= 0)
		return GM_OK;
	launch_k(tile_scan_kernel, dim3((num_tiles + kScanTiles - 1) / kScanTiles), dim3(kScanThreads), 0, stream, num_tiles, vp.bucket_log2, g, capacity);
	return GM_OK;
}

int launch_large_tiles(int phase, const int* radii, const GeometryState& g, const BinningState* b, uint32_t capacity,
                       const ViewParams& vp, cudaStream_t stream)
{
	launch_k(large_tiles_kernel, dim3(148), dim3(kLargeThreads), 0, stream, phase, radii, g, b ? b->keys : nullptr, capacity, vp.W, vp.H,
	                                                     vp.tiles_x, vp.tiles_y, vp.bucket_log2);
	return GM_OK;
}

int launch_emit(int P, const int* radii, const GeometryState& g, const BinningState& b, uint32_t capacity,
                const ViewParams& vp, cudaStream_t stream)
{
	if (P <= 0)
		return GM_OK;
	launch_k(emit_kernel, dim3((P + kEmitThreads - 1) / kEmitThreads), dim3(kEmitThreads), 0, stream, 
		P, radii, g, b, capacity, vp.W, vp.H, vp.tiles_x, vp.tiles_y, vp.bucket_log2);
	launch_large_tiles(1, radii, g, &b, capacity, vp, stream);
	return GM_OK;
}

int launch_sort_pack(int num_tiles, const GeometryState& g, const BinningState& b, uint32_t capacity,
                     const ViewParams& vp, cudaStream_t stream)
{
	if (num_tiles <= 0)
		return GM_OK;
	const uint32_t total = (uint32_t)num_tiles << vp.bucket_log2;
	launch_k(bucket_sort_pack_kernel, dim3((total + kSortWarps - 1) / kSortWarps), dim3(kSortThreads), 0, stream, 
		g, b, capacity, vp.bucket_log2, total);
	// 32 KB of dynamic shared memory, 48 registers: six blocks per SM
	// 16 KB of dynamic shared memory and 64 registers at 128 threads: twelve blocks per SM
	launch_k(big_bucket_sort_pack_kernel<128, 8, false>, dim3(148 * 12), dim3(128), (size_t)kMidKeys * sizeof(uint64_t), stream, g, b, capacity,
	         vp.bucket_log2, (uint32_t)kMidKeys);
	// usually empty: the buckets beyond kMidKeys keys (32 KB of shared memory; beyond that, sorted in place in global memory)
	launch_k(big_bucket_sort_pack_kernel<256, 4, true>, dim3(148 * 2), dim3(256), kBigSmemBytes, stream, g, b, capacity, vp.bucket_log2,
	         (uint32_t)kSortSmem);
	return GM_OK;
}

} // namespace gm
