// Tile binning without a global sort.
//
// The reference builds 64-bit (tile | depth) keys for every (Gaussian, tile) instance, radix-sorts
// all R of them over 41-45 bits and then searches the sorted keys for tile boundaries
// (dgr/cuda_rasterizer/rasterizer_impl.cu:70-138, 464-495).  Here the tile component never enters
// a sort: preprocess counted instances per tile, `tile_scan` turns the counts into segment
// offsets, `emit` drops each instance into its tile's segment, and `sort_pack` orders every
// segment by (depth, gaussian id) inside shared memory and writes the packed, tile-contiguous
// splat records that the blend kernels stage with bulk copies.
//
// Ordering contract (what the blend result depends on): within a tile, ascending depth bits, ties
// broken by ascending Gaussian id -- identical to a stable radix sort of keys emitted in id order
// (SURVEY.md 7 "sort tie order").
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

// ------------------------------------------------------------------------------------------------
// tile_scan: exclusive scan of the aligned per-tile counts (one block).
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int num_tiles, GeometryState g, uint32_t capacity)
{
	__shared__ uint32_t warp_sums[kScanThreads / 32];
	const int tid = threadIdx.x;
	const int per_thread = (num_tiles + kScanThreads - 1) / kScanThreads;
	const int begin = min(num_tiles, tid * per_thread);
	const int end = min(num_tiles, begin + per_thread);

	uint32_t local = 0;
	for (int t = begin; t < end; t++)
		local += (g.tile_count[t] + (kSegAlign - 1)) & ~(uint32_t)(kSegAlign - 1);

	// block-wide exclusive scan of `local`
	uint32_t incl = local;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if ((tid & 31) >= o) incl += v;
	}
	if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
	__syncthreads();
	if (tid < 32) {
		uint32_t w = warp_sums[tid];
		uint32_t wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
			if (tid >= o) wi += v;
		}
		warp_sums[tid] = wi - w;   // exclusive
	}
	__syncthreads();
	uint32_t run = warp_sums[tid >> 5] + incl - local;

	for (int t = begin; t < end; t++) {
		g.tile_start[t] = run;
		g.tile_fill[t] = 0;
		run += (g.tile_count[t] + (kSegAlign - 1)) & ~(uint32_t)(kSegAlign - 1);
	}
	if (tid == kScanThreads - 1) {
		g.header->num_rendered = run;
		g.header->capacity = capacity;
		g.header->overflow = (run > capacity) ? 1u : 0u;
		g.header->num_tiles = (uint32_t)num_tiles;
	}
}

// ------------------------------------------------------------------------------------------------
// emit: one thread per Gaussian, same rectangle walk and the same culling predicate as the count
// in preprocess (so counts and emitted instances agree exactly).  Slot order inside a segment is
// arbitrary; sort_pack makes it deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
emit_kernel(int P, const int* __restrict__ radii, GeometryState g, BinningState b, uint32_t capacity,
            int W, int H, int tiles_x, int tiles_y)
{
	const int idx = blockIdx.x * kEmitThreads + threadIdx.x;
	if (idx >= P)
		return;
	const int radius = radii[idx];
	if (!(radius > 0))
		return;
	const float4 co = g.conic_opacity[idx];
	const float thr = cull_threshold(co.w);
	if (thr < 0.0f)
		return;
	const float2 xy = g.means2D[idx];
	const uint64_t key = ((uint64_t)__float_as_uint(g.depths[idx]) << 32) | (uint32_t)idx;
	int x0, y0, x1, y1;
	tile_rect(xy, radius, tiles_x, tiles_y, x0, y0, x1, y1);
	for (int ty = y0; ty < y1; ty++) {
		const float py0 = (float)(ty * kTile);
		const float py1 = fminf(py0 + (kTile - 1), (float)(H - 1));
		for (int tx = x0; tx < x1; tx++) {
			const float px0 = (float)(tx * kTile);
			const float px1 = fminf(px0 + (kTile - 1), (float)(W - 1));
			if (rect_cannot_contribute(xy.x, xy.y, co.x, co.y, co.z, thr, px0, py0, px1, py1))
				continue;
			const int tile = ty * tiles_x + tx;
			const uint32_t slot = atomicAdd(&g.tile_fill[tile], 1u);
			const uint32_t pos = g.tile_start[tile] + slot;
			if (pos < capacity)
				b.keys[pos] = key;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// sort_pack: one block per tile.  Bitonic sorting network in the "flip" formulation (every
// compare-exchange orders ascending), which needs no padding: a partner index >= n behaves as
// +infinity and the exchange is simply skipped.  Segments up to kSortSmem keys are sorted in
// shared memory; longer ones in place in global memory (rare, slow, still exact).
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortSmem = 4096;   // keys; 32 KB

template <typename Ptr>
__device__ __forceinline__ void compare_exchange(Ptr a, uint32_t i, uint32_t j, uint32_t n)
{
	if (j < n) {
		const uint64_t x = a[i], y = a[j];
		if (x > y) { a[i] = y; a[j] = x; }
	}
}

// k and s are powers of two, so every index is built from shifts and masks (lk = log2 k, ls = log2 s).
template <typename Ptr>
__device__ __forceinline__ void bitonic_sort(Ptr a, uint32_t n)
{
	uint32_t lm = 0;
	while ((1u << lm) < n) lm++;
	const uint32_t half = (1u << lm) >> 1;
	for (uint32_t lk = 1; lk <= lm; lk++) {
		// flip step: i in the lower half of each k-block pairs with its mirror image
		const uint32_t hk_mask = (1u << (lk - 1)) - 1u;
		for (uint32_t t = threadIdx.x; t < half; t += kSortThreads) {
			const uint32_t off = t & hk_mask;
			const uint32_t blk = (t >> (lk - 1)) << lk;
			compare_exchange(a, blk + off, blk + ((1u << lk) - 1u - off), n);
		}
		__syncthreads();
		for (int ls = (int)lk - 2; ls >= 0; ls--) {
			const uint32_t s_mask = (1u << ls) - 1u;
			for (uint32_t t = threadIdx.x; t < half; t += kSortThreads) {
				const uint32_t i = ((t >> ls) << (ls + 1)) + (t & s_mask);
				compare_exchange(a, i, i + (1u << ls), n);
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(kSortThreads)
sort_pack_kernel(GeometryState g, BinningState b, uint32_t capacity)
{
	__shared__ uint64_t s_keys[kSortSmem];
	const int tile = blockIdx.x;
	const uint32_t start = g.tile_start[tile];
	if (start >= capacity)
		return;
	const uint32_t n = min(g.tile_count[tile], capacity - start);
	if (n == 0)
		return;

	uint64_t* keys = b.keys + start;
	const uint64_t* sorted;
	if (n <= kSortSmem) {
		for (uint32_t i = threadIdx.x; i < n; i += kSortThreads)
			s_keys[i] = keys[i];
		__syncthreads();
		bitonic_sort(s_keys, n);
		sorted = s_keys;
	} else {
		__syncthreads();
		bitonic_sort(keys, n);
		sorted = keys;
	}

	// pack: gather the per-Gaussian splat data in blend order
	for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) {
		const uint32_t id = (uint32_t)sorted[i];
		const float2 xy = g.means2D[id];
		const float4 co = g.conic_opacity[id];
		const float4 rgb = g.rgb_clamp[id];
		b.rec_conic[start + i] = co;
		b.rec_xyrg[start + i] = make_float4(xy.x, xy.y, rgb.x, rgb.y);
		b.rec_bid[start + i] = make_float2(rgb.z, __uint_as_float(id));
	}
}

} // namespace

int launch_tile_scan(int num_tiles, const GeometryState& g, uint32_t capacity, cudaStream_t stream)
{
	tile_scan_kernel<<<1, kScanThreads, 0, stream>>>(num_tiles, g, capacity);
	return GM_OK;
}

int launch_emit(int P, const int* radii, const GeometryState& g, const BinningState& b, uint32_t capacity,
                const ViewParams& vp, cudaStream_t stream)
{
	if (P <= 0)
		return GM_OK;
	emit_kernel<<<(P + kEmitThreads - 1) / kEmitThreads, kEmitThreads, 0, stream>>>(
		P, radii, g, b, capacity, vp.W, vp.H, vp.tiles_x, vp.tiles_y);
	return GM_OK;
}

int launch_sort_pack(int num_tiles, const GeometryState& g, const BinningState& b, uint32_t capacity,
                     cudaStream_t stream)
{
	sort_pack_kernel<<<num_tiles, kSortThreads, 0, stream>>>(g, b, capacity);
	return GM_OK;
}

} // namespace gm
