// Layout of the three caller-owned scratch chunks (geometry / image / binning).
//
// These replace the reference's GeometryState / ImageState / BinningState
// (dgr/cuda_rasterizer/rasterizer_impl.h:21-73, rasterizer_impl.cu:155-194).  The chunks are
// opaque to callers, so only the *protocol* is kept (caller sizes them with required<T>(), the
// library carves 128-byte aligned sub-arrays, forward -> backward must see them unchanged); the
// contents are laid out for the B200 pipeline described in DESIGN.md:
//
//   geometry : frame header, per-Gaussian SoA state, per-tile counters (count / start / fill)
//   image    : per-pixel final transmittance and last-contributor index
//   binning  : per-instance (depth,id) keys bucketed by tile, then three packed, tile-contiguous
//              record arrays that the blend kernels stage with cp.async.bulk (TMA)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime_api.h>
#include "../../include/gm_rasterizer.h"

namespace gm {

constexpr int kTile = GM_TILE;                 // 16x16 pixel tiles (config.h:16-17)
constexpr int kTilePixels = kTile * kTile;
constexpr int kSegAlign = 4;                   // tile segments start on a multiple of 4 instances
                                               // => every record array offset is 32-byte aligned
constexpr size_t kChunkAlign = 128;

// Depth buckets.  Every tile's segment of the instance list is pre-partitioned into B = 2^bucket_log2
// consecutive depth buckets by a MONOTONE map depth -> bucket (fine bin of the float's bit pattern, then a
// lookup table built from the frame's depth histogram), so that sorting each small bucket by (depth, id)
// sorts the whole tile.  B is the largest power of two <= kMaxBuckets with tiles * B <= kMaxBucketEntries.
constexpr int kMaxBuckets = 16;
constexpr size_t kMaxBucketEntries = (size_t)1 << 20;
constexpr int kScanTiles = 256;                // tiles per block of the tile scan
constexpr int kMaxScanBlocks = GM_MAX_TILES / kScanTiles;
constexpr int kDepthBins = 4096;               // fine bins: (bits(z) - bits(0.2)) >> 16, 128 per octave
constexpr uint32_t kDepthBinBase = 0x3E4CCCCDu; // bit pattern of 0.2f, the near-plane cull threshold
constexpr int kDepthBinShift = 16;

// Written by the tile-scan kernel, read by later stages and by gm_forward_status().
struct FrameHeader {
	uint32_t num_rendered;   // padded instance total of this view (sum of aligned tile counts)
	uint32_t num_visible;    // Gaussians with radii > 0
	uint32_t overflow;       // 1 if num_rendered exceeded the binning capacity handed to gm_forward
	uint32_t capacity;       // binning capacity in instances used for this frame
	uint32_t num_tiles;
	uint32_t bucket_log2;
	uint32_t num_large;      // Gaussians whose tile rectangle exceeds 64 tiles (walked by large_tiles_kernel)
	uint32_t num_big;        // (tile, bucket) ranges too long for the register sort (big_bucket_sort_pack_kernel)
	uint32_t big_ticket;     // next entry of the big list to be taken by a block of big_bucket_sort_pack_kernel
	uint32_t num_huge;       // ranges beyond the small blocks' shared memory: listed from the END of big_list downwards
	uint32_t huge_ticket;
	uint32_t cull_masks;     // 1: the forward blend left its per-(chunk, warp block) survivor masks in the key array (cull_mask_fits)
	uint32_t pad[20];
};
static_assert(sizeof(FrameHeader) == 128, "FrameHeader must be one 128-byte line");

struct GeometryState {
	FrameHeader* header;
	float* depths;            // [P]   view-space z                                (forward.cu:249)
	int* internal_radii;      // [P]   used when the caller passes radii == NULL   (rasterizer_impl.cu:364)
	float2* means2D;          // [P]   pixel-space centre                          (forward.cu:251)
	float* cov3D;             // [6P]  world covariance, scale/rot path only       (forward.cu:146-151)
	float4* conic_opacity;    // [P]   (a, b, c, opacity)                          (forward.cu:253)
	float4* rgb_clamp;        // [P]   (r, g, b, clamp bits as uint)               (forward.cu:63-70)
	unsigned long long* tile_mask;   // [P] bit (ty-y0)*(x1-x0)+(tx-x0): tile kept by the exact culling (rects of <= 64 tiles)
	uint32_t* large_list;     // [P] ids of the Gaussians with rectangles of more than 64 tiles
	uint32_t* tile_count;     // [GM_MAX_TILES] instances per tile after exact tile culling
	uint32_t* tile_start;     // [GM_MAX_TILES] first instance of the tile (multiple of kSegAlign)
	uint32_t* bucket_cursor;  // [kMaxBucketEntries] per (tile, bucket): count -> start -> end (see binning.cu)
	uint32_t* big_list;       // [kMaxBucketEntries] flat (tile, bucket) ids of the oversized buckets
	unsigned long long* scan_state;  // [kMaxScanBlocks + 1] chained-scan descriptors (flag << 32 | value) + ticket
	uint32_t* depth_hist;     // [kDepthBins] visible-depth histogram of this frame
	uint8_t* depth_lut;       // [kDepthBins] fine depth bin -> bucket (monotone)

	static GeometryState fromChunk(char*& chunk, size_t P);
};

struct ImageState {
	float* accum_alpha;       // [N] final transmittance T                         (forward.cu:369)
	uint32_t* n_contrib;      // [N] 1-based list position of the last contributor (forward.cu:370)

	static ImageState fromChunk(char*& chunk, size_t N);
};

struct BinningState {
	uint64_t* keys;           // [R] (depth bits << 32) | gaussian id, bucketed by tile, unsorted
	float4* rec_conic;        // [R] (a, b, c, opacity)            sorted front-to-back within a tile
	float4* rec_xyrg;         // [R] (x, y, r, g)
	float2* rec_bid;          // [R] (b, gaussian id as bits)

	static BinningState fromChunk(char*& chunk, size_t R);
};

template <typename T>
inline void obtain(char*& chunk, T*& ptr, size_t count, size_t alignment = kChunkAlign)
{
	uintptr_t at = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(uintptr_t)(alignment - 1);
	ptr = reinterpret_cast<T*>(at);
	chunk = reinterpret_cast<char*>(ptr + count);
}

template <typename T>
inline size_t required(size_t n)
{
	char* end = nullptr;
	T::fromChunk(end, n);
	return reinterpret_cast<size_t>(end) + kChunkAlign;
}

inline GeometryState GeometryState::fromChunk(char*& chunk, size_t P)
{
	GeometryState g;
	// header | scan_state | depth_hist | bucket_cursor are contiguous: one memset clears the per-frame counters
	// (frame_clear_bytes); the per-Gaussian arrays follow
	obtain(chunk, g.header, 1);
	obtain(chunk, g.scan_state, (size_t)kMaxScanBlocks + 1);
	obtain(chunk, g.depth_hist, (size_t)kDepthBins);
	obtain(chunk, g.bucket_cursor, kMaxBucketEntries);
	obtain(chunk, g.depths, P);
	obtain(chunk, g.internal_radii, P);
	obtain(chunk, g.means2D, P);
	obtain(chunk, g.cov3D, 6 * P);
	obtain(chunk, g.conic_opacity, P);
	obtain(chunk, g.rgb_clamp, P);
	obtain(chunk, g.tile_mask, P);
	obtain(chunk, g.large_list, P);
	obtain(chunk, g.tile_count, (size_t)GM_MAX_TILES);
	obtain(chunk, g.tile_start, (size_t)GM_MAX_TILES);
	obtain(chunk, g.big_list, kMaxBucketEntries);
	obtain(chunk, g.depth_lut, (size_t)kDepthBins);
	return g;
}

// bytes from the frame header to the end of the bucket cursors in use (num_entries = tiles << bucket_log2)
inline size_t frame_clear_bytes(const GeometryState& g, size_t num_entries)
{
	return (size_t)(reinterpret_cast<const char*>(g.bucket_cursor + num_entries) - reinterpret_cast<const char*>(g.header));
}

inline ImageState ImageState::fromChunk(char*& chunk, size_t N)
{
	ImageState s;
	obtain(chunk, s.accum_alpha, N);
	obtain(chunk, s.n_contrib, N);
	return s;
}

inline BinningState BinningState::fromChunk(char*& chunk, size_t R)
{
	// Round up so that a full-width bulk copy of the last (padded) batch never leaves the chunk.
	size_t Rp = (R + 7) & ~(size_t)7;
	BinningState b;
	obtain(chunk, b.keys, Rp);
	obtain(chunk, b.rec_conic, Rp);
	obtain(chunk, b.rec_xyrg, Rp);
	obtain(chunk, b.rec_bid, Rp);
	return b;
}

// The forward blend leaves, per 32-record chunk of a tile's list and per 8x4 warp block of the tile, the 32-bit mask of the
// records that survived the exact culling; the backward blend reads the mask instead of evaluating the test again.  The
// masks live in the KEY array, which is dead once sort_pack has packed the records: chunk c of tile t is entry
// ((start_t >> 5) + t + c) -- consecutive tiles cannot collide because start_{t+1} >= start_t + n_t -- eight words each.
// They fit when the frame is not extremely sparse; the forward decides per frame and says so in the header.
__host__ __device__ inline bool cull_mask_fits(uint32_t num_rendered, uint32_t num_tiles, uint32_t capacity)
{
	return ((size_t)(num_rendered >> 5) + num_tiles + 8) * 8 <= (size_t)capacity * 2 && num_rendered <= capacity;
}

// Per-view constants.  The four small arrays stay DEVICE pointers (the reference API hands
// them over as device tensors, rasterizer.h:24-132); kernels load them once per thread block.
struct ViewParams {
	const float* view;    // [16] column-major W2C as stored by scene/cameras.py:48 (auxiliary.h:57-65)
	const float* proj;    // [16] full projection (scene/cameras.py:49-50)
	const float* campos;  // [3]
	const float* bg;      // [3]
	float tan_fovx, tan_fovy;
	float focal_x, focal_y;
	float scale_modifier;
	int W, H;
	int tiles_x, tiles_y;
	int D, M;
	int bucket_log2;      // log2 of the number of depth buckets per tile
};

inline int bucket_log2_for(int num_tiles)
{
	int lg = 0;
	while ((1 << (lg + 1)) <= kMaxBuckets && (size_t)num_tiles << (lg + 1) <= kMaxBucketEntries)
		lg++;
	return lg;
}

} // namespace gm
