// Backward of the per-tile alpha blend.
//
// Replaces renderCUDA<3> of dgr/cuda_rasterizer/backward.cu:399-557.  Per-pixel arithmetic follows
// backward.cu:476-554 line by line; what changes is how the nine per-(pixel, splat) partial
// gradients reach global memory.  The reference issues nine global float atomicAdds per
// contributing pair (backward.cu:523,545-554).  Here
//   * the tile's sorted splat records are bulk-copied (TMA) back to front, starting at the last
//     batch any pixel of the tile actually used (the reference walks the whole list);
//   * a warp (8x4 pixel block) culls 32 splats in parallel against its block, exactly as the
//     forward does;
//   * partials are summed across the warp with shuffles, then across the 8 warps of the tile in
//     shared memory, and leave the SM as ONE atomic per (tile, splat, component).
// Summation order differs from the reference's (which is itself non-deterministic), hence the
// 1e-3 relative tolerance of the gradient parity tests.
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 256;
constexpr int kBatch = 256;
constexpr int kComp = 9;   // dcolor r,g,b | dmean2D x,y | dconic a,b,c | dopacity

struct __align__(128) BwdSmem {
	float4 conic[2][kBatch];
	float4 xyrg[2][kBatch];
	float2 bid[2][kBatch];
	float acc[kComp][kBatch + 1];   // +1: component c of splat j lands in bank (c + j) % 32
	uint32_t touched[kBatch];
	uint64_t full[2];
	uint32_t warp_max[kThreads / 32];
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__global__ void __launch_bounds__(kThreads)
blend_backward_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                      int W, int H, int tiles_x, const float* __restrict__ bg_color,
                      const float* __restrict__ dL_dpixels,
                      float* __restrict__ dL_dmean2D,   // [P,3]
                      float* __restrict__ dL_dconic2D,  // [P,4]
                      float* __restrict__ dL_dopacity,  // [P]
                      float* __restrict__ dL_dcolors)   // [P,3]
{
	__shared__ BwdSmem s;

	const int tile = blockIdx.x;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const int bx0 = tile_x * kTile + (warp & 1) * 8;
	const int by0 = tile_y * kTile + (warp >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const uint32_t pix_id = (uint32_t)W * py + px;
	const float pixf_x = (float)px, pixf_y = (float)py;
	const float wx0 = (float)bx0, wy0 = (float)by0;
	const float wx1 = (float)min(bx0 + 7, W - 1), wy1 = (float)min(by0 + 3, H - 1);

	const uint32_t start = g.tile_start[tile];
	uint32_t n = 0;
	if (start < capacity)
		n = min(g.tile_count[tile], capacity - start);

	// backward.cu:430-448
	const float T_final = inside ? img.accum_alpha[pix_id] : 0.0f;
	float T = T_final;
	const uint32_t last_contributor = inside ? min(img.n_contrib[pix_id], n) : 0u;

	// how far back does this warp / this tile have to go?
	uint32_t warp_last = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
	if (lane == 0)
		s.warp_max[warp] = warp_last;
	for (int c = 0; c < kComp; c++)
		s.acc[c][tid] = 0.0f;
	s.touched[tid] = 0u;
	if (tid == 0) {
		mbar_init(&s.full[0], 1);
		mbar_init(&s.full[1], 1);
		fence_mbar_init();
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int w = 0; w < kThreads / 32; w++)
		tile_last = max(tile_last, s.warp_max[w]);
	if (tile_last == 0)
		return;

	float accum_rec0 = 0.0f, accum_rec1 = 0.0f, accum_rec2 = 0.0f;
	float dL_dpixel0 = 0.0f, dL_dpixel1 = 0.0f, dL_dpixel2 = 0.0f;
	if (inside) {
		const size_t HW = (size_t)H * W;
		dL_dpixel0 = dL_dpixels[0 * HW + pix_id];
		dL_dpixel1 = dL_dpixels[1 * HW + pix_id];
		dL_dpixel2 = dL_dpixels[2 * HW + pix_id];
	}
	float last_alpha = 0.0f;
	float last_color0 = 0.0f, last_color1 = 0.0f, last_color2 = 0.0f;

	// backward.cu:455-461, 531-533
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];

	auto issue = [&](int batch, int buf) {
		const uint32_t off = start + (uint32_t)batch * kBatch;
		const uint32_t cnt = min((uint32_t)kBatch, n - (uint32_t)batch * kBatch);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		mbar_arrive_expect_tx(&s.full[buf], cnt4 * 40u);
		bulk_g2s(s.conic[buf], b.rec_conic + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.xyrg[buf], b.rec_xyrg + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.bid[buf], b.rec_bid + off, cnt4 * 8u, &s.full[buf]);
	};

	const int batch_hi = (int)((tile_last - 1) / kBatch);
	if (tid == 0)
		issue(batch_hi, 0);

	for (int it = 0, batch = batch_hi; batch >= 0; it++, batch--) {
		const int buf = it & 1;
		if (tid == 0 && batch > 0)
			issue(batch - 1, buf ^ 1);
		mbar_wait(&s.full[buf], (uint32_t)(it >> 1) & 1u);

		const int batch_base = batch * kBatch;
		// positions >= warp_last are behind every pixel of this warp (backward.cu:487-489)
		const int cnt = min(min(kBatch, (int)n - batch_base), (int)warp_last - batch_base);
		for (int base = (cnt > 0) ? ((cnt - 1) & ~31) : -1; base >= 0; base -= 32) {
			const int j = base + lane;
			bool keep = false;
			if (j < cnt) {
				const float4 co = s.conic[buf][j];
				const float4 xr = s.xyrg[buf][j];
				keep = !rect_cannot_contribute(xr.x, xr.y, co.x, co.y, co.z, cull_threshold(co.w),
				                               wx0, wy0, wx1, wy1);
			}
			uint32_t mask = __ballot_sync(0xffffffffu, keep);
			while (mask) {
				const int k = 31 - __clz(mask);
				mask &= ~(1u << k);
				const int jj = base + k;
				const float4 co = s.conic[buf][jj];
				const float4 xr = s.xyrg[buf][jj];

				// backward.cu:487-501
				bool active = (uint32_t)(batch_base + jj) < last_contributor;
				const float dx = xr.x - pixf_x, dy = xr.y - pixf_y;
				const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
				active = active && !(power > 0.0f);
				const float G = expf(power);
				const float alpha = min(0.99f, co.w * G);
				active = active && !(alpha < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, active))
					continue;

				float g_c0 = 0.0f, g_c1 = 0.0f, g_c2 = 0.0f;
				float g_mx = 0.0f, g_my = 0.0f, g_ca = 0.0f, g_cb = 0.0f, g_cc = 0.0f, g_op = 0.0f;
				if (active) {
					const float cb = s.bid[buf][jj].x;
					// backward.cu:503-524
					T = T / (1.f - alpha);
					const float dchannel_dcolor = alpha * T;
					float dL_dalpha = 0.0f;
					accum_rec0 = last_alpha * last_color0 + (1.f - last_alpha) * accum_rec0;
					last_color0 = xr.z;
					dL_dalpha += (xr.z - accum_rec0) * dL_dpixel0;
					g_c0 = dchannel_dcolor * dL_dpixel0;
					accum_rec1 = last_alpha * last_color1 + (1.f - last_alpha) * accum_rec1;
					last_color1 = xr.w;
					dL_dalpha += (xr.w - accum_rec1) * dL_dpixel1;
					g_c1 = dchannel_dcolor * dL_dpixel1;
					accum_rec2 = last_alpha * last_color2 + (1.f - last_alpha) * accum_rec2;
					last_color2 = cb;
					dL_dalpha += (cb - accum_rec2) * dL_dpixel2;
					g_c2 = dchannel_dcolor * dL_dpixel2;
					// backward.cu:525-534
					dL_dalpha *= T;
					last_alpha = alpha;
					float bg_dot_dpixel = 0.0f;
					bg_dot_dpixel += bg0 * dL_dpixel0;
					bg_dot_dpixel += bg1 * dL_dpixel1;
					bg_dot_dpixel += bg2 * dL_dpixel2;
					dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
					// backward.cu:537-554
					const float dL_dG = co.w * dL_dalpha;
					const float gdx = G * dx;
					const float gdy = G * dy;
					const float dG_ddelx = -gdx * co.x - gdy * co.y;
					const float dG_ddely = -gdy * co.z - gdx * co.y;
					g_mx = dL_dG * dG_ddelx * ddelx_dx;
					g_my = dL_dG * dG_ddely * ddely_dy;
					g_ca = -0.5f * gdx * dx * dL_dG;
					g_cb = -0.5f * gdx * dy * dL_dG;
					g_cc = -0.5f * gdy * dy * dL_dG;
					g_op = G * dL_dalpha;
				}
				g_c0 = warp_sum(g_c0); g_c1 = warp_sum(g_c1); g_c2 = warp_sum(g_c2);
				g_mx = warp_sum(g_mx); g_my = warp_sum(g_my);
				g_ca = warp_sum(g_ca); g_cb = warp_sum(g_cb); g_cc = warp_sum(g_cc);
				g_op = warp_sum(g_op);
				if (lane < kComp) {
					float v = g_c0;
					v = (lane == 1) ? g_c1 : v; v = (lane == 2) ? g_c2 : v;
					v = (lane == 3) ? g_mx : v; v = (lane == 4) ? g_my : v;
					v = (lane == 5) ? g_ca : v; v = (lane == 6) ? g_cb : v;
					v = (lane == 7) ? g_cc : v; v = (lane == 8) ? g_op : v;
					atomicAdd(&s.acc[lane][jj], v);
					if (lane == 0)
						s.touched[jj] = 1u;
				}
			}
		}
		__syncthreads();   // all warps have added their partials of this batch

		// flush: thread t owns splat t of the batch
		if (s.touched[tid]) {
			const uint32_t id = __float_as_uint(s.bid[buf][tid].y);
			atomicAdd(&dL_dcolors[3 * (size_t)id + 0], s.acc[0][tid]);
			atomicAdd(&dL_dcolors[3 * (size_t)id + 1], s.acc[1][tid]);
			atomicAdd(&dL_dcolors[3 * (size_t)id + 2], s.acc[2][tid]);
			atomicAdd(&dL_dmean2D[3 * (size_t)id + 0], s.acc[3][tid]);
			atomicAdd(&dL_dmean2D[3 * (size_t)id + 1], s.acc[4][tid]);
			atomicAdd(&dL_dconic2D[4 * (size_t)id + 0], s.acc[5][tid]);
			atomicAdd(&dL_dconic2D[4 * (size_t)id + 1], s.acc[6][tid]);
			atomicAdd(&dL_dconic2D[4 * (size_t)id + 3], s.acc[7][tid]);
			atomicAdd(&dL_dopacity[id], s.acc[8][tid]);
#pragma unroll
			for (int c = 0; c < kComp; c++)
				s.acc[c][tid] = 0.0f;
			s.touched[tid] = 0u;
		}
		__syncthreads();   // accumulators clean, buffer `buf` free for the copy issued next round
	}
}

} // namespace

int launch_blend_backward(const GeometryState& g, const BinningState& b, const ImageState& img, uint32_t capacity,
                          const ViewParams& vp, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                          float* dL_dopacity, float* dL_dcolor, cudaStream_t stream)
{
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	if (num_tiles <= 0)
		return GM_OK;
	blend_backward_kernel<<<num_tiles, kThreads, 0, stream>>>(
		g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
	return GM_OK;
}

} // namespace gm
