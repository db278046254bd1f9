// Forward per-tile front-to-back alpha blend.
//
// Replaces renderCUDA<3> of dgr/cuda_rasterizer/forward.cu:261-374.  Same per-pixel arithmetic
// (forward.cu:327-362) and the same outputs (out_color [3,H,W], final T, last contributor), but:
//   * the tile's depth-sorted splat records are CONTIGUOUS in memory (binning.cu) and are staged
//     into shared memory with 1-D bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP), double
//     buffered, instead of 256 indexed gathers per batch;
//   * colour is staged with the record instead of being fetched from global memory per
//     contributing pixel (forward.cu:355);
//   * each warp owns an 8x4 pixel block and its 32 lanes first test 32 splats in parallel against
//     that block (rect_cannot_contribute, exact), so only splats that can reach the block are
//     evaluated per pixel.  Skipped splats are exactly those for which every pixel of the block
//     would hit one of the reference's `continue`s, so no output bit changes.
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 256;
constexpr int kBatch = 256;

// Per-warp queue of the splats of one 32-splat chunk that can reach the warp's pixel block, in list order.
struct WarpQueue {
	float4 conic[32];
	float4 xyrg[32];
	float2 bpos[32];    // (blue, 1-based list position as bits)
};

struct __align__(128) FwdSmem {
	float4 conic[2][kBatch];
	float4 xyrg[2][kBatch];
	float2 bid[2][kBatch];
	WarpQueue queue[kThreads / 32];
	uint64_t full[2];
};

__global__ void __launch_bounds__(kThreads)
blend_forward_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                     int W, int H, int tiles_x, const float* __restrict__ bg_color,
                     float* __restrict__ out_color)
{
	__shared__ FwdSmem s;

	const int tile = blockIdx.x;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	// 8x4 pixel block per warp
	const int bx0 = tile_x * kTile + (warp & 1) * 8;
	const int by0 = tile_y * kTile + (warp >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;
	const float wx0 = (float)bx0, wy0 = (float)by0;
	const float wx1 = (float)min(bx0 + 7, W - 1), wy1 = (float)min(by0 + 3, H - 1);

	const uint32_t start = g.tile_start[tile];
	uint32_t n = 0;
	if (start < capacity)
		n = min(g.tile_count[tile], capacity - start);
	const int num_batches = (int)((n + kBatch - 1) / kBatch);

	if (tid == 0) {
		mbar_init(&s.full[0], 1);
		mbar_init(&s.full[1], 1);
		fence_mbar_init();
	}
	__syncthreads();

	auto issue = [&](int batch) {
		const uint32_t off = start + (uint32_t)batch * kBatch;
		const uint32_t cnt = min((uint32_t)kBatch, n - (uint32_t)batch * kBatch);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		const int buf = batch & 1;
		mbar_arrive_expect_tx(&s.full[buf], cnt4 * 40u);
		bulk_g2s(s.conic[buf], b.rec_conic + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.xyrg[buf], b.rec_xyrg + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.bid[buf], b.rec_bid + off, cnt4 * 8u, &s.full[buf]);
	};
	if (tid == 0 && num_batches > 0)
		issue(0);

	// forward.cu:294-298
	bool done = !inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;

	for (int batch = 0; batch < num_batches; batch++) {
		// forward.cu:303-306 (block-wide early exit); this barrier also frees the buffer that the
		// next bulk copy overwrites, because every thread has finished batch-1 by now.
		if (__syncthreads_and(done)) {
			// the bulk copy of this batch may still be in flight; it must land before the block
			// (and with it the shared memory it targets) is retired
			if (tid == 0)
				mbar_wait(&s.full[batch & 1], (uint32_t)(batch >> 1) & 1u);
			break;
		}
		if (tid == 0 && batch + 1 < num_batches)
			issue(batch + 1);
		const int buf = batch & 1;
		mbar_wait(&s.full[buf], (uint32_t)(batch >> 1) & 1u);

		if (__all_sync(0xffffffffu, done))
			continue;

		const int cnt = (int)min((uint32_t)kBatch, n - (uint32_t)batch * kBatch);
		WarpQueue& q = s.queue[warp];
		for (int base = 0; base < cnt; base += 32) {
			// cull 32 splats in parallel against the warp's 8x4 pixel block and compact the survivors
			const int j = base + lane;
			bool keep = false;
			float4 co, xr;
			if (j < cnt) {
				co = s.conic[buf][j];
				xr = s.xyrg[buf][j];
				keep = !rect_cannot_contribute(xr.x, xr.y, co.x, co.y, co.z, cull_threshold(co.w),
				                               wx0, wy0, wx1, wy1);
			}
			const uint32_t mask = __ballot_sync(0xffffffffu, keep);
			if (mask == 0)
				continue;
			if (keep) {
				const int pos = __popc(mask & ((1u << lane) - 1u));
				q.conic[pos] = co;
				q.xyrg[pos] = xr;
				q.bpos[pos] = make_float2(s.bid[buf][j].x, __uint_as_float((uint32_t)(batch * kBatch + j + 1)));
			}
			__syncwarp();
			const int n_keep = __popc(mask);
			if (!done) {
#pragma unroll 2
				for (int i = 0; i < n_keep; i++) {
					const float4 c4 = q.conic[i];
					const float4 x4 = q.xyrg[i];
					// forward.cu:331-335
					const float dx = x4.x - pixf_x, dy = x4.y - pixf_y;
					const float power = -0.5f * (c4.x * dx * dx + c4.z * dy * dy) - c4.y * dx * dy;
					if (power > 0.0f)
						continue;
					// forward.cu:343-345
					const float alpha = min(0.99f, c4.w * expf(power));
					if (alpha < 1.0f / 255.0f)
						continue;
					const float test_T = T * (1 - alpha);
					if (test_T < 0.0001f) {
						done = true;                   // forward.cu:347-351
						break;
					}
					const float2 bp = q.bpos[i];
					// forward.cu:354-355
					C0 += x4.z * alpha * T;
					C1 += x4.w * alpha * T;
					C2 += bp.x * alpha * T;
					T = test_T;
					last_contributor = __float_as_uint(bp.y);
				}
			}
			__syncwarp();   // the queue is rewritten by the next chunk
			if (__all_sync(0xffffffffu, done))
				break;
		}
	}

	// forward.cu:366-373
	if (inside) {
		const uint32_t pix_id = (uint32_t)W * py + px;
		img.accum_alpha[pix_id] = T;
		img.n_contrib[pix_id] = last_contributor;
		const size_t HW = (size_t)H * W;
		out_color[0 * HW + pix_id] = C0 + T * bg_color[0];
		out_color[1 * HW + pix_id] = C1 + T * bg_color[1];
		out_color[2 * HW + pix_id] = C2 + T * bg_color[2];
	}
}

} // namespace

int launch_blend_forward(const GeometryState& g, const BinningState& b, const ImageState& img, uint32_t capacity,
                         const ViewParams& vp, float* out_color, cudaStream_t stream)
{
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	if (num_tiles <= 0)
		return GM_OK;
	blend_forward_kernel<<<num_tiles, kThreads, 0, stream>>>(g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, out_color);
	return GM_OK;
}

} // namespace gm
