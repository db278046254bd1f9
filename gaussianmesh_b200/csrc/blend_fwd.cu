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
// The default is blend_forward_pairs_kernel<4, true, 1>: two splats per iteration in packed fp32x2, two 128-thread blocks per
// tile, survivors paired across chunks; it also leaves the survivor masks of its culling for the backward blend
// (state.h: cull_mask_fits).  blend_forward_kernel (one splat per iteration), the <8, false> instantiation (one block per
// tile) and blend_forward_ring_kernel (no block barrier) are kept for A/B measurements.
#include <cstdlib>
#include "common.cuh"
#include "packed.cuh"
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
	pdl_sync();
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

// ---- two splats per iteration, packed fp32x2 arithmetic -------------------------------------------
// Same decomposition as above (one CTA per tile, a warp per 8x4 pixel block, exact warp-level culling into a
// per-warp queue), but the queue holds the surviving splats in PAIRS and the per-pixel arithmetic that does
// not depend on the running transmittance -- d = xy - pixf, power, exp, opacity * G, feature * alpha
// (forward.cu:331-343,355) -- is evaluated for both splats of a pair with one FADD2 / FMUL2 / FFMA2 per pair of
// scalar operations (packed.cuh; sm_100a).  Every packed operation rounds exactly like the scalar one the
// reference compiles to, so the alpha / T decisions stay bit-identical.  Only the T recurrence
// (forward.cu:346-356) is walked splat by splat, and the reference's `continue` / `break` become predicates
// that zero the splat's alpha: feature * 0 * T added to C leaves C untouched and T * (1 - 0) == T, so no output
// bit moves.  `done` is carried as the pixel's alpha threshold (1/255 while live, +inf afterwards).
struct WarpQueueP {
	// [field][slot][4]: slot k holds splats 2k (A) and 2k+1 (B) of the compacted chunk
	//   0: xA xB yA yB   1: aA aB -bA -bB   2: cA cB oA oB   3: rA rB gA gB   4: bA bB posA posB (1-based, as bits)
	float v[5][17][4];   // 33 entries: up to 32 survivors of a chunk behind one carried over
};

template <int kWarpsF>
struct __align__(128) FwdSmemP {
	float4 conic[2][kBatch];
	float4 xyrg[2][kBatch];
	float2 bid[2][kBatch];
	WarpQueueP queue[kWarpsF];
	uint64_t full[2];
};

// Work folded into the kernel for the training step (gm_forward_ex): the L1 loss of the finished pixels and the
// clearing of the gradient accumulators the backward adds into.  All-null = plain forward.
struct FwdEpilogue {
	const float* target_f;
	const uint8_t* target_u8;
	float* loss;
	float* dL_dimg;
	float4* zero_ptr;
	size_t zero_vec4;
	float inv_numel;
};

// kWarpsF = 8: one block per 16x16 tile; kWarpsF = 4: two blocks per tile, each owning a 16x8 half (a barrier among four
// warps instead of eight, and the block-wide early exit of forward.cu:303-306 per half tile).  kCarry: the survivors of the
// culling are paired across the 32-splat chunks (an odd one waits for the next chunk instead of being paired with a pad).
// kMinB: minimum resident blocks the register allocation is held to (the half-tile kernel measured fastest unconstrained:
// 80 registers, six blocks per SM).
template <int kWarpsF, bool kCarry, int kMinB = 32 / kWarpsF>
__global__ void __launch_bounds__(kWarpsF * 32, kMinB)
blend_forward_pairs_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                           int W, int H, int tiles_x, const float* __restrict__ bg_color,
                           float* __restrict__ out_color, FwdEpilogue epi)
{
	pdl_sync();
	__shared__ FwdSmemP<kWarpsF> s;
	__shared__ float s_loss[kWarpsF];
	if (epi.zero_ptr != nullptr) {
		// this block's slice of the buffer to clear: plain stores, issued before the blend loop and retired behind it
		const size_t per = (epi.zero_vec4 + gridDim.x - 1) / gridDim.x;
		const size_t z0 = (size_t)blockIdx.x * per, z1 = min(z0 + per, epi.zero_vec4);
		for (size_t i = z0 + threadIdx.x; i < z1; i += kWarpsF * 32)
			epi.zero_ptr[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}

	constexpr int kPerTile = (kThreads / 32) / kWarpsF;   // blocks per tile
	const int tile = blockIdx.x / kPerTile;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tw = (blockIdx.x % kPerTile) * kWarpsF + warp;   // the warp's 8x4 pixel block inside the tile

	const int bx0 = tile_x * kTile + (tw & 1) * 8;
	const int by0 = tile_y * kTile + (tw >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const f2 npx = pk1(-(float)px), npy = pk1(-(float)py);
	const float wx0 = (float)bx0, wy0 = (float)by0;
	const float wx1 = (float)min(bx0 + 7, W - 1), wy1 = (float)min(by0 + 3, H - 1);

	const uint32_t start = g.tile_start[tile];
	uint32_t n = 0;
	if (start < capacity)
		n = min(g.tile_count[tile], capacity - start);
	const int num_batches = (int)((n + kBatch - 1) / kBatch);
	// survivor masks of the culling for the backward pass (state.h: cull_mask_fits), in the dead key array
	const bool leave_masks = cull_mask_fits(g.header->num_rendered, g.header->num_tiles, capacity);
	uint32_t* const cull_masks = reinterpret_cast<uint32_t*>(b.keys) + ((size_t)(start >> 5) + (size_t)tile) * 8 + tw;
	if (blockIdx.x == 0 && threadIdx.x == 0)
		g.header->cull_masks = leave_masks ? 1u : 0u;

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
	const float kInf = __int_as_float(0x7f800000);
	float thr = inside ? 1.0f / 255.0f : kInf;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
	const f2 neg_half = pk1(-0.5f);
	int carry = 0;   // kCarry: queue entry 0 holds a survivor of an earlier chunk that has not been blended yet

	for (int batch = 0; batch < num_batches; batch++) {
		// forward.cu:303-306 (block-wide early exit); this barrier also frees the buffer that the
		// next bulk copy overwrites, because every thread has finished batch-1 by now.
		if (__syncthreads_and(thr == kInf)) {
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

		if (__all_sync(0xffffffffu, thr == kInf))
			continue;

		const int cnt = (int)min((uint32_t)kBatch, n - (uint32_t)batch * kBatch);
		WarpQueueP& q = s.queue[warp];
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
			if (leave_masks && lane == 0)
				cull_masks[(size_t)(batch * (kBatch / 32) + (base >> 5)) * 8] = mask;
			const int n_keep = __popc(mask);
			int n_pairs;
			int total = 0;
			if (kCarry) {
				const bool last_chunk = batch == num_batches - 1 && base + 32 >= cnt;
				total = carry + n_keep;
				if (total == 0 || (n_keep == 0 && !last_chunk))
					continue;
				if (keep) {
					const int pos = carry + __popc(mask & ((1u << lane) - 1u));
					const int slot = pos >> 1, h = pos & 1;
					q.v[0][slot][h] = xr.x;
					q.v[0][slot][2 + h] = xr.y;
					q.v[1][slot][h] = co.x;
					q.v[1][slot][2 + h] = -co.y;
					q.v[2][slot][h] = co.z;
					q.v[2][slot][2 + h] = co.w;
					q.v[3][slot][h] = xr.z;
					q.v[3][slot][2 + h] = xr.w;
					q.v[4][slot][h] = s.bid[buf][j].x;
					q.v[4][slot][2 + h] = __uint_as_float((uint32_t)(batch * kBatch + j + 1));
				}
				if (last_chunk && (total & 1)) {
					// the list's very last entry, if odd, is paired with a splat of opacity 0 (alpha == 0: skipped)
					if (lane < 5) {
						q.v[lane][total >> 1][1] = 0.0f;
						q.v[lane][total >> 1][3] = 0.0f;
					}
					total++;
				}
				n_pairs = total >> 1;
			} else {
				if (mask == 0)
					continue;
				// the lane behind the last survivor pads an odd queue with a splat of opacity 0 (alpha == 0: skipped)
				const int pos = keep ? __popc(mask & ((1u << lane) - 1u)) : n_keep;
				const bool pad = !keep && (n_keep & 1) && lane == (__ffs(~mask) - 1);
				if (keep || pad) {
					const int slot = pos >> 1, h = pos & 1;
					q.v[0][slot][h] = keep ? xr.x : 0.0f;
					q.v[0][slot][2 + h] = keep ? xr.y : 0.0f;
					q.v[1][slot][h] = keep ? co.x : 0.0f;
					q.v[1][slot][2 + h] = keep ? -co.y : 0.0f;
					q.v[2][slot][h] = keep ? co.z : 0.0f;
					q.v[2][slot][2 + h] = keep ? co.w : 0.0f;
					q.v[3][slot][h] = keep ? xr.z : 0.0f;
					q.v[3][slot][2 + h] = keep ? xr.w : 0.0f;
					q.v[4][slot][h] = keep ? s.bid[buf][j].x : 0.0f;
					q.v[4][slot][2 + h] = __uint_as_float(keep ? (uint32_t)(batch * kBatch + j + 1) : 0u);
				}
				n_pairs = (n_keep + 1) >> 1;
			}
			__syncwarp();
#pragma unroll 2
			for (int k = 0; k < n_pairs; k++) {
				const ulonglong2 XY = *reinterpret_cast<const ulonglong2*>(q.v[0][k]);
				const ulonglong2 AB = *reinterpret_cast<const ulonglong2*>(q.v[1][k]);
				const ulonglong2 CO = *reinterpret_cast<const ulonglong2*>(q.v[2][k]);
				// forward.cu:331-335: d = xy - pixf; power = -0.5 (a dx dx + c dy dy) - b dx dy
				const f2 dx = add2(XY.x, npx), dy = add2(XY.y, npy);
				f2 t = mul2(dy, CO.x);
				const f2 u = mul2(dx, AB.x);
				t = mul2(dy, t);
				const f2 sq = fma2(dx, u, t);
				const f2 v = mul2(dx, AB.y);
				const f2 w = mul2(dy, v);
				const f2 power = fma2(sq, neg_half, w);
				// forward.cu:343: alpha = min(0.99, opacity * exp(power))
				const f2 al = mul2(CO.y, exp2x(power));
				const float aA = fminf(lo(al), 0.99f), aB = fminf(hi(al), 0.99f);
				const float4 BP = *reinterpret_cast<const float4*>(q.v[4][k]);

				// splat A -- forward.cu:336,344: the two `continue`s zero alpha; :346-351: test_T = T (1 - alpha), a
				// skipped splat gives T back (T >= 1e-4 while the pixel is live), a terminating one freezes the pixel
				const bool sA = (lo(power) > 0.0f) | (aA < thr);
				const float eA = sA ? 0.0f : aA;
				const float ttA = __fmul_rn(T, __fadd_rn(1.0f, -eA));
				const bool tA = ttA < 0.0001f;
				thr = tA ? kInf : thr;
				const float wA = tA ? 0.0f : eA;
				const float TA = T;
				T = tA ? T : ttA;
				last_contributor = (sA | tA) ? last_contributor : __float_as_uint(BP.z);
				// splat B
				const bool sB = (hi(power) > 0.0f) | (aB < thr);
				const float eB = sB ? 0.0f : aB;
				const float ttB = __fmul_rn(T, __fadd_rn(1.0f, -eB));
				const bool tB = ttB < 0.0001f;
				thr = tB ? kInf : thr;
				const float wB = tB ? 0.0f : eB;
				const float TB = T;
				T = tB ? T : ttB;
				last_contributor = (sB | tB) ? last_contributor : __float_as_uint(BP.w);

				// forward.cu:354-355: C += (feature * alpha) * T
				const ulonglong2 RG = *reinterpret_cast<const ulonglong2*>(q.v[3][k]);
				const f2 w2 = pk(wA, wB);
				const f2 cr = mul2(w2, RG.x), cg = mul2(w2, RG.y), cb = mul2(w2, pk(BP.x, BP.y));
				C0 = __fmaf_rn(TB, hi(cr), __fmaf_rn(TA, lo(cr), C0));
				C1 = __fmaf_rn(TB, hi(cg), __fmaf_rn(TA, lo(cg), C1));
				C2 = __fmaf_rn(TB, hi(cb), __fmaf_rn(TA, lo(cb), C2));
			}
			__syncwarp();   // the queue is rewritten by the next chunk
			if (__all_sync(0xffffffffu, thr == kInf))
				break;
			if (kCarry) {
				carry = total & 1;
				if (carry && n_pairs > 0) {
					// the odd survivor (half A of slot n_pairs) becomes entry 0 of the next chunk's list
					if (lane < 10) {
						const int f = lane >> 1, e = (lane & 1) * 2;
						q.v[f][0][e] = q.v[f][n_pairs][e];
					}
					__syncwarp();
				}
			}
		}
	}

	// forward.cu:366-373
	float part = 0.0f;
	if (inside) {
		const uint32_t pix_id = (uint32_t)W * py + px;
		img.accum_alpha[pix_id] = T;
		img.n_contrib[pix_id] = last_contributor;
		const size_t HW = (size_t)H * W;
		const float o0 = C0 + T * bg_color[0], o1 = C1 + T * bg_color[1], o2 = C2 + T * bg_color[2];
		out_color[0 * HW + pix_id] = o0;
		out_color[1 * HW + pix_id] = o1;
		out_color[2 * HW + pix_id] = o2;
		if (epi.loss != nullptr) {
			// utils/loss_utils.py:17-18 on the pixel just finished: |image - target| and its gradient
			const float oc[3] = {o0, o1, o2};
#pragma unroll
			for (int ch = 0; ch < 3; ch++) {
				const size_t at = ch * HW + pix_id;
				const float t = epi.target_u8 != nullptr ? (float)epi.target_u8[at] / 255.0f : epi.target_f[at];
				const float d = oc[ch] - t;
				part += fabsf(d);
				if (epi.dL_dimg != nullptr)
					epi.dL_dimg[at] = (d > 0.0f ? epi.inv_numel : (d < 0.0f ? -epi.inv_numel : 0.0f));
			}
		}
	}
	if (epi.loss != nullptr) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			part += __shfl_xor_sync(0xffffffffu, part, o);
		if (lane == 0)
			s_loss[warp] = part;
		__syncthreads();
		if (tid == 0) {
			float sum = 0.0f;
#pragma unroll
			for (int w = 0; w < kWarpsF; w++)
				sum += s_loss[w];
			atomicAdd(epi.loss, sum * epi.inv_numel);
		}
	}
}


// ---- the same packed-pair kernel without a block-wide barrier in the batch loop -------------------------
// The staged records live in a ring of kStagesF stages.  A warp waits only for the stage it needs (mbarrier), and
// when it is through with a stage it counts itself off; the warp that arrives last refills the stage with the batch
// kStagesF further on.  A warp whose 32 pixels have all terminated (or whose block is culled) therefore never holds
// the others back, and nobody waits at a barrier for the slowest warp of every batch (16 % of the stall samples of the
// barrier version).  The reference's block-wide early exit (forward.cu:303-306) becomes: once all eight warps are
// done, the refilling warp stops issuing copies and publishes the first batch that will never arrive (`stop_at`);
// every warp drains the copies already in flight and leaves.
constexpr int kBatchF = 128;
constexpr int kStagesF = 3;

struct __align__(128) FwdSmemR {
	float4 conic[kStagesF][kBatchF];
	float4 xyrg[kStagesF][kBatchF];
	float2 bid[kStagesF][kBatchF];
	WarpQueueP queue[kThreads / 32];
	uint64_t full[kStagesF];
	uint64_t empty[kStagesF];          // one arrival per thread that is through with the stage
	uint32_t released[kStagesF];
	uint32_t done_warps;
	int stop_at;
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t"
		"}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}

__global__ void __launch_bounds__(kThreads)
blend_forward_ring_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                          int W, int H, int tiles_x, const float* __restrict__ bg_color,
                          float* __restrict__ out_color)
{
	pdl_sync();
	__shared__ FwdSmemR s;

	const int tile = blockIdx.x;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	// 8x4 pixel block per warp
	const int bx0 = tile_x * kTile + (warp & 1) * 8;
	const int by0 = tile_y * kTile + (warp >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const f2 npx = pk1(-(float)px), npy = pk1(-(float)py);
	const float wx0 = (float)bx0, wy0 = (float)by0;
	const float wx1 = (float)min(bx0 + 7, W - 1), wy1 = (float)min(by0 + 3, H - 1);

	const uint32_t start = g.tile_start[tile];
	uint32_t n = 0;
	if (start < capacity)
		n = min(g.tile_count[tile], capacity - start);
	const int num_batches = (int)((n + kBatchF - 1) / kBatchF);

	auto issue = [&](int batch, int st) {
		const uint32_t off = start + (uint32_t)batch * kBatchF;
		const uint32_t cnt = min((uint32_t)kBatchF, n - (uint32_t)batch * kBatchF);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		mbar_arrive_expect_tx(&s.full[st], cnt4 * 40u);
		bulk_g2s(s.conic[st], b.rec_conic + off, cnt4 * 16u, &s.full[st]);
		bulk_g2s(s.xyrg[st], b.rec_xyrg + off, cnt4 * 16u, &s.full[st]);
		bulk_g2s(s.bid[st], b.rec_bid + off, cnt4 * 8u, &s.full[st]);
	};
	if (tid == 0) {
#pragma unroll
		for (int st = 0; st < kStagesF; st++) {
			mbar_init(&s.full[st], 1);
			mbar_init(&s.empty[st], kThreads);
			s.released[st] = 0;
		}
		s.done_warps = 0;
		s.stop_at = num_batches;
		fence_mbar_init();
#pragma unroll
		for (int st = 0; st < kStagesF; st++)
			if (st < num_batches)
				issue(st, st);
	}
	__syncthreads();

	// forward.cu:294-298
	const float kInf = __int_as_float(0x7f800000);
	float thr = inside ? 1.0f / 255.0f : kInf;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
	const f2 neg_half = pk1(-0.5f);
	WarpQueueP& q = s.queue[warp];
	bool warp_done = __all_sync(0xffffffffu, thr == kInf);
	if (warp_done && lane == 0)
		atomicAdd(&s.done_warps, 1u);

	int st = 0;
	uint32_t parity = 0;
	for (int batch = 0; batch < num_batches; batch++) {
		// wait for this batch -- or learn that it will never be issued (every warp of the tile is done)
		bool stopped = false;
		while (!mbar_try_wait(&s.full[st], parity)) {
			if (*reinterpret_cast<volatile int*>(&s.stop_at) <= batch) {
				stopped = true;
				break;
			}
		}
		if (stopped)
			break;

		if (!warp_done) {
			const int cnt = (int)min((uint32_t)kBatchF, n - (uint32_t)batch * kBatchF);
			for (int base = 0; base < cnt; base += 32) {
				// cull 32 splats in parallel against the warp's 8x4 pixel block and compact the survivors
				const int j = base + lane;
				bool keep = false;
				float4 co, xr;
				if (j < cnt) {
					co = s.conic[st][j];
					xr = s.xyrg[st][j];
					keep = !rect_cannot_contribute(xr.x, xr.y, co.x, co.y, co.z, cull_threshold(co.w),
					                               wx0, wy0, wx1, wy1);
				}
				const uint32_t mask = __ballot_sync(0xffffffffu, keep);
				if (mask == 0)
					continue;
				const int n_keep = __popc(mask);
				{
					// the lane behind the last survivor pads an odd queue with a splat of opacity 0 (alpha == 0: skipped)
					const int pos = keep ? __popc(mask & ((1u << lane) - 1u)) : n_keep;
					const bool pad = !keep && (n_keep & 1) && lane == (__ffs(~mask) - 1);
					if (keep || pad) {
						const int slot = pos >> 1, h = pos & 1;
						q.v[0][slot][h] = keep ? xr.x : 0.0f;
						q.v[0][slot][2 + h] = keep ? xr.y : 0.0f;
						q.v[1][slot][h] = keep ? co.x : 0.0f;
						q.v[1][slot][2 + h] = keep ? -co.y : 0.0f;
						q.v[2][slot][h] = keep ? co.z : 0.0f;
						q.v[2][slot][2 + h] = keep ? co.w : 0.0f;
						q.v[3][slot][h] = keep ? xr.z : 0.0f;
						q.v[3][slot][2 + h] = keep ? xr.w : 0.0f;
						q.v[4][slot][h] = keep ? s.bid[st][j].x : 0.0f;
						q.v[4][slot][2 + h] = __uint_as_float(keep ? (uint32_t)(batch * kBatchF + j + 1) : 0u);
					}
				}
				__syncwarp();
				const int n_pairs = (n_keep + 1) >> 1;
#pragma unroll 2
				for (int k = 0; k < n_pairs; k++) {
					const ulonglong2 XY = *reinterpret_cast<const ulonglong2*>(q.v[0][k]);
					const ulonglong2 AB = *reinterpret_cast<const ulonglong2*>(q.v[1][k]);
					const ulonglong2 CO = *reinterpret_cast<const ulonglong2*>(q.v[2][k]);
					// forward.cu:331-335: d = xy - pixf; power = -0.5 (a dx dx + c dy dy) - b dx dy
					const f2 dx = add2(XY.x, npx), dy = add2(XY.y, npy);
					f2 t = mul2(dy, CO.x);
					const f2 u = mul2(dx, AB.x);
					t = mul2(dy, t);
					const f2 sq = fma2(dx, u, t);
					const f2 v = mul2(dx, AB.y);
					const f2 w = mul2(dy, v);
					const f2 power = fma2(sq, neg_half, w);
					// forward.cu:343: alpha = min(0.99, opacity * exp(power))
					const f2 al = mul2(CO.y, exp2x(power));
					const float aA = fminf(lo(al), 0.99f), aB = fminf(hi(al), 0.99f);
					const float4 BP = *reinterpret_cast<const float4*>(q.v[4][k]);

					// splat A -- forward.cu:336,344: the two `continue`s zero alpha; :346-351: test_T = T (1 - alpha), a
					// skipped splat gives T back (T >= 1e-4 while the pixel is live), a terminating one freezes the pixel
					const bool sA = (lo(power) > 0.0f) | (aA < thr);
					const float eA = sA ? 0.0f : aA;
					const float ttA = __fmul_rn(T, __fadd_rn(1.0f, -eA));
					const bool tA = ttA < 0.0001f;
					thr = tA ? kInf : thr;
					const float wA = tA ? 0.0f : eA;
					const float TA = T;
					T = tA ? T : ttA;
					last_contributor = (sA | tA) ? last_contributor : __float_as_uint(BP.z);
					// splat B
					const bool sB = (hi(power) > 0.0f) | (aB < thr);
					const float eB = sB ? 0.0f : aB;
					const float ttB = __fmul_rn(T, __fadd_rn(1.0f, -eB));
					const bool tB = ttB < 0.0001f;
					thr = tB ? kInf : thr;
					const float wB = tB ? 0.0f : eB;
					const float TB = T;
					T = tB ? T : ttB;
					last_contributor = (sB | tB) ? last_contributor : __float_as_uint(BP.w);

					// forward.cu:354-355: C += (feature * alpha) * T
					const ulonglong2 RG = *reinterpret_cast<const ulonglong2*>(q.v[3][k]);
					const f2 w2 = pk(wA, wB);
					const f2 cr = mul2(w2, RG.x), cg = mul2(w2, RG.y), cb = mul2(w2, pk(BP.x, BP.y));
					C0 = __fmaf_rn(TB, hi(cr), __fmaf_rn(TA, lo(cr), C0));
					C1 = __fmaf_rn(TB, hi(cg), __fmaf_rn(TA, lo(cg), C1));
					C2 = __fmaf_rn(TB, hi(cb), __fmaf_rn(TA, lo(cb), C2));
				}
				__syncwarp();   // the queue is rewritten by the next chunk
				if (__all_sync(0xffffffffu, thr == kInf)) {
					warp_done = true;
					if (lane == 0)
						atomicAdd(&s.done_warps, 1u);
					break;
				}
			}
		}

		// release the stage; the warp that arrives last refills it, unless every warp of the tile has finished
		__syncwarp();
		mbar_arrive(&s.empty[st]);                                       // release: this thread's reads of the stage are done
		if (lane == 0) {
			const uint32_t old = atomicAdd(&s.released[st], 1u);
			if ((old & (kThreads / 32 - 1)) == kThreads / 32 - 1 && batch + kStagesF < num_batches) {
				mbar_wait(&s.empty[st], parity);                          // acquire: every thread's arrival (returns at once)
				if (*reinterpret_cast<volatile uint32_t*>(&s.done_warps) == kThreads / 32)
					atomicMin(&s.stop_at, batch + kStagesF);
				else {
					fence_proxy_async();
					issue(batch + kStagesF, st);
				}
			}
		}
		if (++st == kStagesF) {
			st = 0;
			parity ^= 1u;
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
                         const ViewParams& vp, float* out_color, const gm_forward_epilogue* epilogue, bool* epilogue_done,
                         cudaStream_t stream)
{
	if (epilogue_done != nullptr)
		*epilogue_done = false;
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	if (num_tiles <= 0)
		return GM_OK;
	// The default is the packed-pair kernel with one block barrier per 256-record batch (two blocks per tile).  GM_BLEND_FWD=ring selects the
	// variant over a three-stage ring without block barriers (bit-identical output, measured 5 % slower: waiting warps spin
	// on the mbarrier and take issue slots from the working ones -- DESIGN.md 8), GM_BLEND_SCALAR=1 the one-pixel-per-thread
	// kernel; both are kept for A/B measurements.
	static const bool scalar = std::getenv("GM_BLEND_SCALAR") != nullptr && std::getenv("GM_BLEND_SCALAR")[0] == '1';
	static const bool ring = std::getenv("GM_BLEND_FWD") != nullptr && std::getenv("GM_BLEND_FWD")[0] == 'r';
	if (!scalar && ring) {
		launch_k(blend_forward_ring_kernel, dim3(num_tiles), dim3(kThreads), 0, stream, g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, out_color);
		return GM_OK;
	}
	if (scalar) {
		launch_k(blend_forward_kernel, dim3(num_tiles), dim3(kThreads), 0, stream, g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, out_color);
		return GM_OK;
	}
	FwdEpilogue e = {};
	if (epilogue != nullptr) {
		if (epilogue->loss != nullptr && epilogue->target != nullptr) {
			e.target_f = epilogue->target_is_u8 ? nullptr : static_cast<const float*>(epilogue->target);
			e.target_u8 = epilogue->target_is_u8 ? static_cast<const uint8_t*>(epilogue->target) : nullptr;
			e.loss = epilogue->loss;
			e.dL_dimg = epilogue->dL_dimg;
			e.inv_numel = 1.0f / (3.0f * (float)vp.W * (float)vp.H);
		}
		if (epilogue->zero_ptr != nullptr && epilogue->zero_floats > 0) {
			e.zero_ptr = reinterpret_cast<float4*>(epilogue->zero_ptr);
			e.zero_vec4 = epilogue->zero_floats / 4;
		}
		if (epilogue_done != nullptr)
			*epilogue_done = true;
	}
	// two 128-thread blocks per tile, survivors paired across chunks; GM_BLEND_FWD=tile selects one 256-thread block per tile
	// with per-chunk padding (the default until late in round 2; A/B measurements)
	static const bool whole_tile = std::getenv("GM_BLEND_FWD") != nullptr && std::getenv("GM_BLEND_FWD")[0] == 't';
	if (whole_tile)
		launch_k(blend_forward_pairs_kernel<8, false>, dim3(num_tiles), dim3(kThreads), 0, stream, g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, out_color, e);
	else
		launch_k(blend_forward_pairs_kernel<4, true, 1>, dim3(num_tiles * 2), dim3(kThreads / 2), 0, stream, g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, out_color, e);
	return GM_OK;
}

} // namespace gm
