// Backward of the per-tile alpha blend.
//
// Replaces renderCUDA<3> of dgr/cuda_rasterizer/backward.cu:399-557.  The per-pixel recurrence
// (backward.cu:476-534: T <- T/(1-alpha), accum_rec, dL/dalpha with the background term) is kept; what
// changes is everything around it.  The reference issues nine global float atomicAdds per contributing
// (pixel, splat) pair (backward.cu:523,545-554).  Here
//   * the tile's sorted splat records are bulk-copied (TMA) back to front, starting at the last batch any
//     pixel of the tile actually used (the reference walks the whole list);
//   * a warp owns an 8x4 pixel block and culls 32 splats in parallel against it, exactly as the forward;
//   * per pair each lane produces nine MOMENTS (3 colour sums and sum q, q dx, q dy, q dx^2, q dx dy,
//     q dy^2 with q = G dL/dalpha); the splat-constant factors of backward.cu:537-554 (opacity, conic
//     entries, -1/2, W/2, H/2) are applied once per (warp, splat) when the sums leave the warp;
//   * the nine moments are summed over the 32 pixels with a transposing butterfly (12 shuffles instead of
//     45) and parked in a warp-private shared-memory row; after 32 splats every lane owns one splat and
//     sends its gradients to global memory with nine atomics -- one set per (warp, splat) that actually
//     received a contribution, instead of one per (pixel, splat).
// Summation order differs from the reference's (which is itself non-deterministic), hence the 1e-3
// relative tolerance of the gradient parity tests.
//
// Four kernels share this decomposition and differ in how the nine sums leave the warp: blend_backward_kernel (one splat
// per iteration) and blend_backward_pairs_kernel (two, packed fp32x2) use the butterfly described above,
// blend_backward_mma_kernel the tensor cores, and blend_backward_cols_kernel -- THE DEFAULT, at the end of the file --
// parks two numbers per (pixel, splat) and lets lanes own splats when the sums are formed.  It also reads the survivor
// masks the forward blend left behind instead of culling again (state.h: cull_mask_fits).
#include <cstdlib>
#include "common.cuh"
#include "packed.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kBatch = 256;
constexpr int kComp = 9;   // colour r,g,b | q | q dx | q dy | q dx^2 | q dx dy | q dy^2

// Per-warp queue of the splats of one 32-splat chunk that can reach the warp's pixel block, back to front,
// and the parked sums of their nine moments.
struct WarpQueue {
	float4 conic[32];
	float4 xyrg[32];
	float2 bid[32];               // (blue, gaussian id as bits)
	uint32_t pos[32];             // 0-based position in the tile's list
	float park[32 * kComp];       // [queue slot * 9 + component]: stride 9 is bank-conflict free
};

struct __align__(128) BwdSmem {
	float4 conic[2][kBatch];
	float4 xyrg[2][kBatch];
	float2 bid[2][kBatch];
	WarpQueue queue[kWarps];
	uint64_t full[2];
	uint32_t warp_max[kWarps];
};

// Sum nine per-lane values over the warp.  On return lane 4q (q = 0..7) holds the total of v[q] and lane 2
// holds the total of v[8] (so do lanes 4q+1 resp. 4q+2, 4q+3 -- only one of each is used).
__device__ __forceinline__ float butterfly9(const float (&v)[kComp], int lane)
{
	// xor 16: 8 values -> 4 per lane
	const bool h4 = lane & 16;
	float w[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const float keep = h4 ? v[i + 4] : v[i];
		const float send = h4 ? v[i] : v[i + 4];
		w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
	float s = v[8] + __shfl_xor_sync(0xffffffffu, v[8], 16);
	// xor 8: 4 -> 2
	const bool h3 = lane & 8;
	float u[2];
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const float keep = h3 ? w[i + 2] : w[i];
		const float send = h3 ? w[i] : w[i + 2];
		u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
	s += __shfl_xor_sync(0xffffffffu, s, 8);
	// xor 4: 2 -> 1
	const bool h2 = lane & 4;
	float t;
	{
		const float keep = h2 ? u[1] : u[0];
		const float send = h2 ? u[0] : u[1];
		t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
	}
	s += __shfl_xor_sync(0xffffffffu, s, 4);
	// xor 2: lanes with bit 1 clear keep t (value index lane >> 2), the others keep s (value 8)
	const bool h1 = lane & 2;
	{
		const float keep = h1 ? s : t;
		const float send = h1 ? t : s;
		t = keep + __shfl_xor_sync(0xffffffffu, send, 2);
	}
	t += __shfl_xor_sync(0xffffffffu, t, 1);
	return t;
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
	pdl_sync();
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
	if (tid == 0) {
		mbar_init(&s.full[0], 1);
		mbar_init(&s.full[1], 1);
		fence_mbar_init();
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int w = 0; w < kWarps; w++)
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

	// backward.cu:455-461: d(pixel offset)/d(NDC mean); applied once per (warp, splat) below
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	// backward.cu:531-534: dL/dalpha += (-T_final / (1 - alpha)) * (bg . dL/dpixel)
	const float bg_term = -T_final * (bg_color[0] * dL_dpixel0 + bg_color[1] * dL_dpixel1 + bg_color[2] * dL_dpixel2);

	auto issue = [&](int batch, int buf) {
		const uint32_t off = start + (uint32_t)batch * kBatch;
		const uint32_t cnt = min((uint32_t)kBatch, n - (uint32_t)batch * kBatch);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		mbar_arrive_expect_tx(&s.full[buf], cnt4 * 40u);
		bulk_g2s(s.conic[buf], b.rec_conic + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.xyrg[buf], b.rec_xyrg + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.bid[buf], b.rec_bid + off, cnt4 * 8u, &s.full[buf]);
	};

	WarpQueue& q = s.queue[warp];
	const int batch_hi = (int)((tile_last - 1) / kBatch);
	if (tid == 0)
		issue(batch_hi, 0);

	for (int it = 0, batch = batch_hi; batch >= 0; it++, batch--) {
		const int buf = it & 1;
		// every warp has finished batch+1 (buffer buf^1) before it is overwritten
		if (it > 0)
			__syncthreads();
		if (tid == 0 && batch > 0)
			issue(batch - 1, buf ^ 1);
		mbar_wait(&s.full[buf], (uint32_t)(it >> 1) & 1u);

		const int batch_base = batch * kBatch;
		// positions >= warp_last are behind every pixel of this warp (backward.cu:487-489)
		const int cnt = min(min(kBatch, (int)n - batch_base), (int)warp_last - batch_base);
		for (int base = (cnt > 0) ? ((cnt - 1) & ~31) : -1; base >= 0; base -= 32) {
			// cull 32 splats in parallel against the warp's 8x4 pixel block; compact the survivors back to front
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
				const int slot = __popc(mask >> lane) - 1;          // highest list position first
				q.conic[slot] = co;
				q.xyrg[slot] = xr;
				q.bid[slot] = s.bid[buf][j];
				q.pos[slot] = (uint32_t)(batch_base + j);
			}
			__syncwarp();
			const int n_keep = __popc(mask);
			uint32_t touched = 0;
			for (int i = 0; i < n_keep; i++) {
				const float4 c4 = q.conic[i];
				const float4 x4 = q.xyrg[i];
				// backward.cu:487-501
				const float dx = x4.x - pixf_x, dy = x4.y - pixf_y;
				const float power = -0.5f * (c4.x * dx * dx + c4.z * dy * dy) - c4.y * dx * dy;
				const float G = expf(power);
				const float alpha = min(0.99f, c4.w * G);
				const bool active = (q.pos[i] < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, active))
					continue;
				touched |= 1u << i;

				float v[kComp];
#pragma unroll
				for (int c = 0; c < kComp; c++)
					v[c] = 0.0f;
				if (active) {
					const float cb = q.bid[i].x;
					// backward.cu:503-534
					const float rcp = __fdividef(1.f, 1.f - alpha);   // MUFU.RCP: 1 ulp, far inside the 1e-3 gradient tolerance
					T = T * rcp;
					const float dchannel_dcolor = alpha * T;
					const float keep_prev = 1.f - last_alpha;
					accum_rec0 = last_alpha * last_color0 + keep_prev * accum_rec0;
					accum_rec1 = last_alpha * last_color1 + keep_prev * accum_rec1;
					accum_rec2 = last_alpha * last_color2 + keep_prev * accum_rec2;
					last_color0 = x4.z; last_color1 = x4.w; last_color2 = cb;
					last_alpha = alpha;
					float dL_dalpha = (x4.z - accum_rec0) * dL_dpixel0 + (x4.w - accum_rec1) * dL_dpixel1 +
					                  (cb - accum_rec2) * dL_dpixel2;
					dL_dalpha = dL_dalpha * T + bg_term * rcp;
					v[0] = dchannel_dcolor * dL_dpixel0;
					v[1] = dchannel_dcolor * dL_dpixel1;
					v[2] = dchannel_dcolor * dL_dpixel2;
					// moments of q = G dL/dalpha (backward.cu:537-554 with the splat constants factored out)
					const float qq = G * dL_dalpha;
					const float qdx = qq * dx, qdy = qq * dy;
					v[3] = qq;
					v[4] = qdx;
					v[5] = qdy;
					v[6] = qdx * dx;
					v[7] = qdx * dy;
					v[8] = qdy * dy;
				}
				const float total = butterfly9(v, lane);
				if ((lane & 3) == 0)
					q.park[i * kComp + (lane >> 2)] = total;
				else if (lane == 2)
					q.park[i * kComp + 8] = total;
			}
			__syncwarp();
			if ((touched >> lane) & 1u) {
				// this lane sends queue slot `lane` to global memory
				const float* m = q.park + lane * kComp;
				const float4 c4 = q.conic[lane];
				const uint32_t id = __float_as_uint(q.bid[lane].y);
				const float a = c4.x, bb = c4.y, c = c4.z, o = c4.w;
				const float Sq = m[3], Sx = m[4], Sy = m[5], Sxx = m[6], Sxy = m[7], Syy = m[8];
				atomicAdd(&dL_dcolors[3 * (size_t)id + 0], m[0]);
				atomicAdd(&dL_dcolors[3 * (size_t)id + 1], m[1]);
				atomicAdd(&dL_dcolors[3 * (size_t)id + 2], m[2]);
				// dL/dG = o dL/dalpha;  dG/ddelx = -G (a dx + b dy);  dG/ddely = -G (c dy + b dx)
				atomicAdd(&dL_dmean2D[3 * (size_t)id + 0], -o * ddelx_dx * (a * Sx + bb * Sy));
				atomicAdd(&dL_dmean2D[3 * (size_t)id + 1], -o * ddely_dy * (c * Sy + bb * Sx));
				const float h = -0.5f * o;
				atomicAdd(&dL_dconic2D[4 * (size_t)id + 0], h * Sxx);
				atomicAdd(&dL_dconic2D[4 * (size_t)id + 1], h * Sxy);
				atomicAdd(&dL_dconic2D[4 * (size_t)id + 3], h * Syy);
				atomicAdd(&dL_dopacity[id], Sq);
			}
			__syncwarp();   // queue and park rows are rewritten by the next chunk
		}
	}
}

// ---- two splats per iteration, packed fp32x2 arithmetic -------------------------------------------
// Same decomposition as above, but the queue holds the surviving splats in PAIRS (A = further back, B = the next
// one towards the camera) and everything that does not sit on the T / accum_rec recurrences is evaluated for
// both splats with one FADD2 / FMUL2 / FFMA2 per pair of scalar operations (packed.cuh): d, power, G, alpha
// (bit-identical to the forward's, so both passes take the same alpha / T decisions), the dL/dalpha dot product
// and the nine moments.  An inactive splat is carried with alpha = 0 and G = 0: T / (1 - 0) == T, the pending
// (last_alpha, last_color) term is folded into accum_rec one splat early (accum_rec <- la lc + (1 - la) accum_rec
// followed by la <- 0 is the same recurrence), and all nine moments are zero.  The 2 x 9 moments are summed over
// the 32 pixels by one transposing butterfly: the first exchange (lane ^ 16) also separates the two splats,
// lanes 0-15 finish splat A and lanes 16-31 splat B.
struct WarpQueueP {
	// [field][slot][4]: slot k holds splats 2k (A) and 2k+1 (B) of the compacted chunk, back to front
	//   0: xA xB yA yB   1: aA aB -bA -bB   2: cA cB oA oB   3: rA rB gA gB   4: bA bB posA posB (0-based, as bits)
	float v[5][16][4];
	uint32_t id[32];
	float park[32 * kComp];
};

struct __align__(128) BwdSmemP {
	float4 conic[2][kBatch];
	float4 xyrg[2][kBatch];
	float2 bid[2][kBatch];
	WarpQueueP queue[kWarps];
	uint64_t full[2];
	uint32_t warp_max[kWarps];
};

// Sum nine packed pairs over the warp.  On return lane l holds, for splat (l >> 4) of the pair, the total of
// component (l & 1) ? 8 : ((l >> 1) & 7).
__device__ __forceinline__ float butterfly18(const f2 (&v)[kComp], int lane)
{
	// xor 16: separates the splats -- lanes 0-15 keep the lo halves, lanes 16-31 the hi halves
	// (PRMT with a per-lane selector: one instruction per select, where `?:` on the two halves turns into
	// pairs of predicated moves)
	const uint32_t sel_keep = (lane & 16) ? 0x7654u : 0x3210u, sel_send = sel_keep ^ 0x4444u;
	float w[kComp];
#pragma unroll
	for (int i = 0; i < kComp; i++) {
		const uint32_t a = __float_as_uint(lo(v[i])), b = __float_as_uint(hi(v[i]));
		const float keep = __uint_as_float(__byte_perm(a, b, sel_keep));
		const float send = __uint_as_float(__byte_perm(a, b, sel_send));
		w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
	// xor 8: 8 -> 4
	const bool h3 = lane & 8;
	float u[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const float keep = h3 ? w[i + 4] : w[i];
		const float send = h3 ? w[i] : w[i + 4];
		u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
	float e = w[8] + __shfl_xor_sync(0xffffffffu, w[8], 8);
	// xor 4: 4 -> 2
	const bool h2 = lane & 4;
	float y[2];
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const float keep = h2 ? u[i + 2] : u[i];
		const float send = h2 ? u[i] : u[i + 2];
		y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
	}
	e += __shfl_xor_sync(0xffffffffu, e, 4);
	// xor 2: 2 -> 1
	const bool h1 = lane & 2;
	float z;
	{
		const float keep = h1 ? y[1] : y[0];
		const float send = h1 ? y[0] : y[1];
		z = keep + __shfl_xor_sync(0xffffffffu, send, 2);
	}
	e += __shfl_xor_sync(0xffffffffu, e, 2);
	// xor 1: even lanes finish z, odd lanes finish component 8
	const bool h0 = lane & 1;
	const float keep = h0 ? e : z;
	const float send = h0 ? z : e;
	return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__global__ void __launch_bounds__(kThreads, 4)
blend_backward_pairs_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                            int W, int H, int tiles_x, const float* __restrict__ bg_color,
                            const float* __restrict__ dL_dpixels,
                            float* __restrict__ dL_dmean2D,   // [P,3]
                            float* __restrict__ dL_dconic2D,  // [P,4]
                            float* __restrict__ dL_dopacity,  // [P]
                            float* __restrict__ dL_dcolors)   // [P,3]
{
	pdl_sync();
	__shared__ BwdSmemP s;

	const int tile = blockIdx.x;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const int bx0 = tile_x * kTile + (warp & 1) * 8;
	const int by0 = tile_y * kTile + (warp >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const uint32_t pix_id = (uint32_t)W * py + px;
	const f2 npx = pk1(-(float)px), npy = pk1(-(float)py);
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
	if (tid == 0) {
		mbar_init(&s.full[0], 1);
		mbar_init(&s.full[1], 1);
		fence_mbar_init();
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int w = 0; w < kWarps; w++)
		tile_last = max(tile_last, s.warp_max[w]);
	if (tile_last == 0)
		return;

	float dL_dpixel0 = 0.0f, dL_dpixel1 = 0.0f, dL_dpixel2 = 0.0f;
	if (inside) {
		const size_t HW = (size_t)H * W;
		dL_dpixel0 = dL_dpixels[0 * HW + pix_id];
		dL_dpixel1 = dL_dpixels[1 * HW + pix_id];
		dL_dpixel2 = dL_dpixels[2 * HW + pix_id];
	}
	// accum_rec and the pending (last_alpha * last_color, 1 - last_alpha) term of backward.cu:509-515
	float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
	float pend0 = 0.0f, pend1 = 0.0f, pend2 = 0.0f, keep_prev = 1.0f;

	// backward.cu:455-461: d(pixel offset)/d(NDC mean); applied once per (warp, splat) below
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	// backward.cu:531-534: dL/dalpha += (-T_final / (1 - alpha)) * (bg . dL/dpixel)
	const f2 bg_term = pk1(-T_final * (bg_color[0] * dL_dpixel0 + bg_color[1] * dL_dpixel1 + bg_color[2] * dL_dpixel2));
	const f2 dLp0 = pk1(dL_dpixel0), dLp1 = pk1(dL_dpixel1), dLp2 = pk1(dL_dpixel2);
	const f2 neg_half = pk1(-0.5f), neg_one = pk1(-1.0f), one = pk1(1.0f);

	auto issue = [&](int batch, int buf) {
		const uint32_t off = start + (uint32_t)batch * kBatch;
		const uint32_t cnt = min((uint32_t)kBatch, n - (uint32_t)batch * kBatch);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		mbar_arrive_expect_tx(&s.full[buf], cnt4 * 40u);
		bulk_g2s(s.conic[buf], b.rec_conic + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.xyrg[buf], b.rec_xyrg + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.bid[buf], b.rec_bid + off, cnt4 * 8u, &s.full[buf]);
	};

	WarpQueueP& q = s.queue[warp];
	// where this lane parks its butterfly18 result of pair 0 (entry lane >> 4, component as described there)
	float* const park_lane = q.park + (lane >> 4) * kComp + ((lane & 1) ? 8 : ((lane >> 1) & 7));
	const bool park_writer = (lane & 1) == 0 || (lane & 15) == 1;
	const int batch_hi = (int)((tile_last - 1) / kBatch);
	if (tid == 0)
		issue(batch_hi, 0);

	for (int it = 0, batch = batch_hi; batch >= 0; it++, batch--) {
		const int buf = it & 1;
		// every warp has finished batch+1 (buffer buf^1) before it is overwritten
		if (it > 0)
			__syncthreads();
		if (tid == 0 && batch > 0)
			issue(batch - 1, buf ^ 1);
		mbar_wait(&s.full[buf], (uint32_t)(it >> 1) & 1u);

		const int batch_base = batch * kBatch;
		// positions >= warp_last are behind every pixel of this warp (backward.cu:487-489)
		const int cnt = min(min(kBatch, (int)n - batch_base), (int)warp_last - batch_base);
		for (int base = (cnt > 0) ? ((cnt - 1) & ~31) : -1; base >= 0; base -= 32) {
			// cull 32 splats in parallel against the warp's 8x4 pixel block; compact the survivors back to front
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
			const int n_keep = __popc(mask);
			{
				// one non-surviving lane pads an odd queue with a splat that no pixel accepts (position 2^32 - 1)
				const bool pad = !keep && (n_keep & 1) && lane == (__ffs(~mask) - 1);
				if (keep || pad) {
					const int at = keep ? __popc(mask >> lane) - 1 : n_keep;   // highest list position first
					const int slot = at >> 1, h = at & 1;
					const float2 bi = keep ? s.bid[buf][j] : make_float2(0.0f, 0.0f);
					q.v[0][slot][h] = keep ? xr.x : 0.0f;
					q.v[0][slot][2 + h] = keep ? xr.y : 0.0f;
					q.v[1][slot][h] = keep ? co.x : 0.0f;
					q.v[1][slot][2 + h] = keep ? -co.y : 0.0f;
					q.v[2][slot][h] = keep ? co.z : 0.0f;
					q.v[2][slot][2 + h] = keep ? co.w : 0.0f;
					q.v[3][slot][h] = keep ? xr.z : 0.0f;
					q.v[3][slot][2 + h] = keep ? xr.w : 0.0f;
					q.v[4][slot][h] = bi.x;
					q.v[4][slot][2 + h] = __uint_as_float(keep ? (uint32_t)(batch_base + j) : 0xffffffffu);
					q.id[at] = __float_as_uint(bi.y);
				}
			}
			__syncwarp();
			const int n_pairs = (n_keep + 1) >> 1;
			uint32_t touched = 0, pair_bits = 3u;
			float* park_at = park_lane;
			for (int k = 0; k < n_pairs; k++, pair_bits <<= 2, park_at += 2 * kComp) {
				const ulonglong2 XY = *reinterpret_cast<const ulonglong2*>(q.v[0][k]);
				const ulonglong2 AB = *reinterpret_cast<const ulonglong2*>(q.v[1][k]);
				const ulonglong2 CO = *reinterpret_cast<const ulonglong2*>(q.v[2][k]);
				const float4 BP = *reinterpret_cast<const float4*>(q.v[4][k]);
				// backward.cu:487-501, the forward's instruction sequence
				const f2 dx = add2(XY.x, npx), dy = add2(XY.y, npy);
				f2 t = mul2(dy, CO.x);
				const f2 u = mul2(dx, AB.x);
				t = mul2(dy, t);
				const f2 sq = fma2(dx, u, t);
				const f2 vv = mul2(dx, AB.y);
				const f2 ww = mul2(dy, vv);
				const f2 power = fma2(sq, neg_half, ww);
				const f2 G = exp2x(power);
				const f2 al = mul2(CO.y, G);
				const float aA = fminf(lo(al), 0.99f), aB = fminf(hi(al), 0.99f);
				const bool skipA = (__float_as_uint(BP.z) >= last_contributor) | (lo(power) > 0.0f) | (aA < 1.0f / 255.0f);
				const bool skipB = (__float_as_uint(BP.w) >= last_contributor) | (hi(power) > 0.0f) | (aB < 1.0f / 255.0f);
				if (__all_sync(0xffffffffu, skipA & skipB))
					continue;
				touched |= pair_bits;

				const f2 e2 = pk(skipA ? 0.0f : aA, skipB ? 0.0f : aB);
				const f2 Ge = pk(skipA ? 0.0f : lo(G), skipB ? 0.0f : hi(G));
				// backward.cu:503-507: T <- T / (1 - alpha).  MUFU.RCP: 1 ulp, far inside the 1e-3 gradient tolerance
				const f2 om = fma2(e2, neg_one, one);
				const float rcpA = rcp_approx_ftz(lo(om)), rcpB = rcp_approx_ftz(hi(om));
				const float TA = T * rcpA, TB = TA * rcpB;
				T = TB;
				const f2 T2 = pk(TA, TB), rcp2 = pk(rcpA, rcpB);
				const f2 dchannel_dcolor = mul2(e2, T2);
				// backward.cu:509-521: accum_rec, walked A then B
				const ulonglong2 RG = *reinterpret_cast<const ulonglong2*>(q.v[3][k]);
				const f2 col0 = RG.x, col1 = RG.y, col2 = pk(BP.x, BP.y);
				const f2 ac0 = mul2(e2, col0), ac1 = mul2(e2, col1), ac2 = mul2(e2, col2);
				const float accA0 = fmaf(keep_prev, acc0, pend0), accA1 = fmaf(keep_prev, acc1, pend1),
				            accA2 = fmaf(keep_prev, acc2, pend2);
				acc0 = fmaf(lo(om), accA0, lo(ac0));
				acc1 = fmaf(lo(om), accA1, lo(ac1));
				acc2 = fmaf(lo(om), accA2, lo(ac2));
				pend0 = hi(ac0); pend1 = hi(ac1); pend2 = hi(ac2);
				keep_prev = hi(om);
				const f2 d0 = fma2(pk(accA0, acc0), neg_one, col0);
				const f2 d1 = fma2(pk(accA1, acc1), neg_one, col1);
				const f2 d2 = fma2(pk(accA2, acc2), neg_one, col2);
				f2 dL_dalpha = fma2(d2, dLp2, fma2(d1, dLp1, mul2(d0, dLp0)));
				// backward.cu:526-534
				dL_dalpha = fma2(dL_dalpha, T2, mul2(bg_term, rcp2));

				f2 v[kComp];
				v[0] = mul2(dchannel_dcolor, dLp0);
				v[1] = mul2(dchannel_dcolor, dLp1);
				v[2] = mul2(dchannel_dcolor, dLp2);
				// moments of q = G dL/dalpha (backward.cu:537-554 with the splat constants factored out)
				const f2 qq = mul2(Ge, dL_dalpha);
				const f2 qdx = mul2(qq, dx), qdy = mul2(qq, dy);
				v[3] = qq;
				v[4] = qdx;
				v[5] = qdy;
				v[6] = mul2(qdx, dx);
				v[7] = mul2(qdx, dy);
				v[8] = mul2(qdy, dy);
				const float total = butterfly18(v, lane);
				if (park_writer)
					*park_at = total;
			}
			__syncwarp();
			if ((touched >> lane) & 1u) {
				// this lane sends queue entry `lane` to global memory (unless no pixel accepted it)
				const float* m = q.park + lane * kComp;
				const float Sq = m[3], Sx = m[4], Sy = m[5], Sxx = m[6], Sxy = m[7], Syy = m[8];
				const float m0 = m[0], m1 = m[1], m2 = m[2];
				const uint32_t any = (__float_as_uint(m0) | __float_as_uint(m1) | __float_as_uint(m2) | __float_as_uint(Sq) |
				                      __float_as_uint(Sx) | __float_as_uint(Sy) | __float_as_uint(Sxx) | __float_as_uint(Sxy) |
				                      __float_as_uint(Syy)) << 1;
				if (any != 0) {
					const int slot = lane >> 1, h = lane & 1;
					const float a = q.v[1][slot][h], bb = -q.v[1][slot][2 + h], c = q.v[2][slot][h], o = q.v[2][slot][2 + h];
					const uint32_t id = q.id[lane];
					atomicAdd(&dL_dcolors[3 * (size_t)id + 0], m0);
					atomicAdd(&dL_dcolors[3 * (size_t)id + 1], m1);
					atomicAdd(&dL_dcolors[3 * (size_t)id + 2], m2);
					// dL/dG = o dL/dalpha;  dG/ddelx = -G (a dx + b dy);  dG/ddely = -G (c dy + b dx)
					atomicAdd(&dL_dmean2D[3 * (size_t)id + 0], -o * ddelx_dx * (a * Sx + bb * Sy));
					atomicAdd(&dL_dmean2D[3 * (size_t)id + 1], -o * ddely_dy * (c * Sy + bb * Sx));
					const float hh = -0.5f * o;
					atomicAdd(&dL_dconic2D[4 * (size_t)id + 0], hh * Sxx);
					atomicAdd(&dL_dconic2D[4 * (size_t)id + 1], hh * Sxy);
					atomicAdd(&dL_dconic2D[4 * (size_t)id + 3], hh * Syy);
					atomicAdd(&dL_dopacity[id], Sq);
				}
			}
			__syncwarp();   // queue and park rows are rewritten by the next chunk
		}
	}
}


// ---- tensor-core reduction over the warp's 32 pixels ----------------------------------------------
// The per-(pixel, splat) work above ends in nine sums over the 32 pixels of the warp.  With q = G dL/dalpha and
// w = alpha T those sums are two small matrix products over the pixel index p:
//     M[s, k] = sum_p q[s, p] * Phi[p, k]      Phi[p, :] = (1, u, v, u^2, u v, v^2),  u, v = pixel offset from the block centre
//     C[s, c] = sum_p w[s, p] * dL/dpixel[p, c]
// (backward.cu:523,537-554 sum q dx, q dx^2, ... with dx = x_s - px: the splat-centred moments follow from the
// pixel-centred ones by a shift applied once per (warp, splat), see `flush`).  Each lane parks (q, w) of its pixel in a
// warp-private shared-memory slab, 8 splats deep, and both products run on the tensor cores as ONE 16 x 32 A operand
// (rows 0-7: q of the eight splats, rows 8-15: their w) against the two B operands (mma.sync.m16n8k8, TF32 inputs, FP32
// accumulation): q and w are split into a TF32 head and an exact fp32 tail (two MMAs, 21+ mantissa bits), Phi is exact
// in TF32 (multiples of 1/4 below 16), dL/dpixel is split head / tail across two columns of its B operand.  This
// replaces the 18-value transposing butterfly (37 % of the pairs kernel's instructions) with two 64-bit shared-memory
// stores per pair of splats and ~20 instructions per splat at flush time.
//
// The record stages are released per warp (a counter per stage; the warp that arrives last refills the stage), so there
// is no block-wide barrier in the loop and a warp with few surviving splats runs ahead of its neighbours.
//
// All shared-memory traffic of the inner loops goes through 32-bit shared-window addresses kept in registers (inline
// ld.shared / st.shared): left to itself ptxas re-derives them from %tid and %cgaid in every iteration (S2R, a
// ~100-cycle instruction) when registers are short.
constexpr int kBatchM = 64;       // records per stage
constexpr int kStagesM = 4;
constexpr int kFlushPairs = 4;    // slab depth in pairs (8 splats = the 8 + 8 rows of one MMA)
constexpr int kSlabPitch = 80;    // floats per slab row: 32 pixels x (q, w) + 16 pad (conflict-free LDS.128 fragments)

struct WarpQueueM {
	// [field][slot][4]: slot k holds splats 2k (A) and 2k+1 (B) of the compacted chunk, back to front
	//   0: xA xB yA yB   1: aA aB -bA -bB   2: cA cB oA oB   3: rA rB gA gB   4: bA bB posA posB (0-based, as bits)
	float v[5][16][4];
	uint32_t id[32];
	float slab[2 * kFlushPairs][kSlabPitch];   // [splat row][pixel][q, w]; reused as the 8 x 12 transpose buffer of a flush
	uint32_t dfr[32][8];                       // this warp's B fragments of dL/dpixel, per lane
};

struct __align__(128) BwdSmemM {
	float4 conic[kStagesM][kBatchM];
	float4 xyrg[kStagesM][kBatchM];
	float2 bid[kStagesM][kBatchM];
	WarpQueueM queue[kWarps];
	uint32_t phi[32][8];                       // B fragments of the pixel monomials, per lane (same for every warp)
	uint64_t full[kStagesM];
	uint64_t empty[kStagesM];                  // one arrival per thread that is through with the stage
	uint32_t released[kStagesM];
	uint32_t warp_max[kWarps];
};

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
	return v;
}

__device__ __forceinline__ uint4 lds128u(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}

__device__ __forceinline__ uint2 lds64u(uint32_t addr)
{
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
	return v;
}

__device__ __forceinline__ ulonglong2 lds128p(uint32_t addr)   // two packed fp32 pairs
{
	ulonglong2 v;
	asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr) : "memory");
	return v;
}

__device__ __forceinline__ float lds32(uint32_t addr)
{
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
	return v;
}

__device__ __forceinline__ void sts64(uint32_t addr, float a, float b)
{
	asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void sts32(uint32_t addr, float a)
{
	asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}

// fire-and-forget global float add (SASS REDG: no return value, nothing for a later instruction to wait on)
__device__ __forceinline__ void red_add(float* addr, float v)
{
	asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// vector forms (sm_90+): one request for two / four adjacent floats; the address must be 8- / 16-byte aligned
__device__ __forceinline__ void red_add2(float* addr, float a, float b)
{
	asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// x = head + tail: head = x with the 13 low mantissa bits cleared (a TF32 value), tail exact in fp32 (the tensor core
// reads its top 19 bits): together 21+ mantissa bits of x enter the product
__device__ __forceinline__ void tf32_split(float x, uint32_t& head, uint32_t& tail)
{
	head = __float_as_uint(x) & 0xffffe000u;
	tail = __float_as_uint(x - __uint_as_float(head));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
	             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kThreads, 4)
blend_backward_mma_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                          int W, int H, int tiles_x, const float* __restrict__ bg_color,
                          const float* __restrict__ dL_dpixels,
                          float* __restrict__ dL_dmean2D,   // [P,3]
                          float* __restrict__ dL_dconic2D,  // [P,4]
                          float* __restrict__ dL_dopacity,  // [P]
                          float* __restrict__ dL_dcolors)   // [P,3]
{
	pdl_sync();
	extern __shared__ __align__(128) unsigned char smem_raw[];
	BwdSmemM& s = *reinterpret_cast<BwdSmemM*>(smem_raw);

	const int tile = blockIdx.x;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const int bx0 = tile_x * kTile + (warp & 1) * 8;
	const int by0 = tile_y * kTile + (warp >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const uint32_t pix_id = (uint32_t)W * py + px;
	const f2 npx = pk1(-(float)px), npy = pk1(-(float)py);
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

	uint32_t warp_last = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
	if (lane == 0)
		s.warp_max[warp] = warp_last;
	if (tid == 0) {
#pragma unroll
		for (int st = 0; st < kStagesM; st++) {
			mbar_init(&s.full[st], 1);
			mbar_init(&s.empty[st], kThreads);
			s.released[st] = 0;
		}
		fence_mbar_init();
	}
	// B fragments of mma.m16n8k8 (row.col): this thread holds B[k = tg][n = gid] and B[k = tg + 4][n = gid]; k-step i maps
	// k = tg, tg + 4 to the pixels 8 i + 2 tg, 8 i + 2 tg + 1 of the warp's block (lane index = pixel index), the order in
	// which the A fragments are read from the slab.  Column n of Phi is the n-th monomial of the pixel's offset from the
	// block centre (columns 6, 7 are zero); it is the same for every warp.
	const int gid = lane >> 2, tg = lane & 3;
	if (warp == 0) {
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 2; j++) {
				const int p = 8 * i + 2 * tg + j;
				const float u = (float)(p & 7) - 3.5f, v = (float)(p >> 3) - 1.5f;
				const float mono = gid == 0 ? 1.0f : gid == 1 ? u : gid == 2 ? v : gid == 3 ? u * u : gid == 4 ? u * v
				                   : gid == 5 ? v * v : 0.0f;
				s.phi[lane][2 * i + j] = __float_as_uint(mono);
			}
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int w = 0; w < kWarps; w++)
		tile_last = max(tile_last, s.warp_max[w]);
	if (tile_last == 0)
		return;

	float dL_dpixel0 = 0.0f, dL_dpixel1 = 0.0f, dL_dpixel2 = 0.0f;
	if (inside) {
		const size_t HW = (size_t)H * W;
		dL_dpixel0 = dL_dpixels[0 * HW + pix_id];
		dL_dpixel1 = dL_dpixels[1 * HW + pix_id];
		dL_dpixel2 = dL_dpixels[2 * HW + pix_id];
	}

	WarpQueueM& q = s.queue[warp];
	// the warp's dL/dpixel as B fragments: columns 2 c, 2 c + 1 = TF32 head and fp32 tail of channel c, columns 6, 7 zero.
	// The 3 x 32 values are exchanged through the (not yet used) slab.
	q.slab[0][lane] = dL_dpixel0;
	q.slab[0][32 + lane] = dL_dpixel1;
	q.slab[1][lane] = dL_dpixel2;
	__syncwarp();
	{
		const int c = gid >> 1;
		const float* src = c == 0 ? &q.slab[0][0] : c == 1 ? &q.slab[0][32] : &q.slab[1][0];
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 2; j++) {
				uint32_t head, tail;
				tf32_split(src[8 * i + 2 * tg + j], head, tail);
				q.dfr[lane][2 * i + j] = gid >= 6 ? 0u : (gid & 1) ? tail : head;
			}
	}
	__syncwarp();

	// accum_rec and the pending (last_alpha * last_color, 1 - last_alpha) term of backward.cu:509-515
	float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
	float pend0 = 0.0f, pend1 = 0.0f, pend2 = 0.0f, keep_prev = 1.0f;

	// backward.cu:455-461: d(pixel offset)/d(NDC mean); applied once per (warp, splat) in flush
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	// backward.cu:531-534: dL/dalpha += (-T_final / (1 - alpha)) * (bg . dL/dpixel)
	const f2 bg_term = pk1(-T_final * (bg_color[0] * dL_dpixel0 + bg_color[1] * dL_dpixel1 + bg_color[2] * dL_dpixel2));
	const f2 dLp0 = pk1(dL_dpixel0), dLp1 = pk1(dL_dpixel1), dLp2 = pk1(dL_dpixel2);
	const f2 neg_half = pk1(-0.5f), neg_one = pk1(-1.0f), one = pk1(1.0f);

	auto issue = [&](int batch, int st) {
		const uint32_t off = start + (uint32_t)batch * kBatchM;
		const uint32_t cnt = min((uint32_t)kBatchM, n - (uint32_t)batch * kBatchM);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		mbar_arrive_expect_tx(&s.full[st], cnt4 * 40u);
		bulk_g2s(s.conic[st], b.rec_conic + off, cnt4 * 16u, &s.full[st]);
		bulk_g2s(s.xyrg[st], b.rec_xyrg + off, cnt4 * 16u, &s.full[st]);
		bulk_g2s(s.bid[st], b.rec_bid + off, cnt4 * 8u, &s.full[st]);
	};

	// shared-window addresses, made opaque so that they stay in registers
	uint32_t q_base = smem_u32(&q);
	uint32_t frag_base = q_base + (uint32_t)offsetof(WarpQueueM, slab) + (uint32_t)(gid * kSlabPitch + 4 * tg) * 4u;
	uint32_t slab_lane = q_base + (uint32_t)offsetof(WarpQueueM, slab) + (uint32_t)lane * 8u;
	asm volatile("" : "+r"(q_base), "+r"(frag_base), "+r"(slab_lane));
	const uint32_t phi_lane = smem_u32(&s.phi[lane][0]);
	const float cx = (float)bx0 + 3.5f, cy = (float)by0 + 1.5f;
	constexpr uint32_t kField = 16 * 16;   // bytes per queue field

	// Reduce the slab (up to 4 pairs = 8 splats, queue slots first .. first + 3) on the tensor cores and send every touched
	// splat's gradients to global memory: lane l < 8 owns splat l of the group (pair l >> 1, half l & 1).
	auto flush = [&](int first, uint32_t touched) {
		__syncwarp();
		float pm[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dm[4] = {0.0f, 0.0f, 0.0f, 0.0f};
		const uint32_t dfr_lane = q_base + (uint32_t)offsetof(WarpQueueM, dfr) + (uint32_t)lane * 32u;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			// A fragment of k-step i: row gid = q of splat gid, row gid + 8 = its w; pixels 8 i + 2 tg and + 1
			const float4 a = lds128(frag_base + 64u * i);
			const uint2 bp = lds64u(phi_lane + 8u * i), bd = lds64u(dfr_lane + 8u * i);
			uint32_t hi[4], lo[4];
			tf32_split(a.x, hi[0], lo[0]); tf32_split(a.y, hi[1], lo[1]);
			tf32_split(a.z, hi[2], lo[2]); tf32_split(a.w, hi[3], lo[3]);
			mma_tf32(pm, hi, bp.x, bp.y);
			mma_tf32(dm, hi, bd.x, bd.y);
			mma_tf32(pm, lo, bp.x, bp.y);
			mma_tf32(dm, lo, bd.x, bd.y);
		}
		__syncwarp();   // every fragment has been read: the slab is reused to transpose the results
		// C fragments: (row gid, columns 2 tg, 2 tg + 1) of q x Phi; (row gid + 8, same columns) of w x dL/dpixel
		const uint32_t park = q_base + (uint32_t)offsetof(WarpQueueM, slab);   // [8 splats][12]: Sq Su Sv Suu Suv Svv | colour 0 1 2
		if (tg < 3) {
			sts64(park + (uint32_t)(gid * 12 + 2 * tg) * 4u, pm[0], pm[1]);
			sts32(park + (uint32_t)(gid * 12 + 6 + tg) * 4u, dm[2] + dm[3]);
		}
		__syncwarp();
		if (lane < 8 && ((touched >> (lane >> 1)) & 1u)) {
			const float4 m0 = lds128(park + (uint32_t)lane * 48u);
			const float4 m1 = lds128(park + (uint32_t)lane * 48u + 16u);
			const float c2 = lds32(park + (uint32_t)lane * 48u + 32u);
			const float Sq = m0.x, Su = m0.y, Sv = m0.z, Suu = m0.w, Suv = m1.x, Svv = m1.y, c0 = m1.z, c1 = m1.w;
			const uint32_t any = (__float_as_uint(c0) | __float_as_uint(c1) | __float_as_uint(c2) | __float_as_uint(Sq) |
			                      __float_as_uint(Su) | __float_as_uint(Sv) | __float_as_uint(Suu) | __float_as_uint(Suv) |
			                      __float_as_uint(Svv)) << 1;
			if (any != 0) {
				const int slot = first + (lane >> 1), h = lane & 1;
				const float a = q.v[1][slot][h], bb = -q.v[1][slot][2 + h], c = q.v[2][slot][h], o = q.v[2][slot][2 + h];
				const uint32_t id = q.id[2 * slot + h];
				// dx = x_s - px = X - u with X = x_s - (block centre): shift the pixel-centred moments to the splat
				const float X = q.v[0][slot][h] - cx, Y = q.v[0][slot][2 + h] - cy;
				const float Sx = fmaf(X, Sq, -Su), Sy = fmaf(Y, Sq, -Sv);
				const float Sxx = fmaf(X, fmaf(X, Sq, -2.0f * Su), Suu);
				const float Sxy = fmaf(X, fmaf(Y, Sq, -Sv), fmaf(-Y, Su, Suv));
				const float Syy = fmaf(Y, fmaf(Y, Sq, -2.0f * Sv), Svv);
				red_add(&dL_dcolors[3 * (size_t)id + 0], c0);
				red_add(&dL_dcolors[3 * (size_t)id + 1], c1);
				red_add(&dL_dcolors[3 * (size_t)id + 2], c2);
				// dL/dG = o dL/dalpha;  dG/ddelx = -G (a dx + b dy);  dG/ddely = -G (c dy + b dx)
				red_add(&dL_dmean2D[3 * (size_t)id + 0], -o * ddelx_dx * (a * Sx + bb * Sy));
				red_add(&dL_dmean2D[3 * (size_t)id + 1], -o * ddely_dy * (c * Sy + bb * Sx));
				const float hh = -0.5f * o;
				red_add(&dL_dconic2D[4 * (size_t)id + 0], hh * Sxx);
				red_add(&dL_dconic2D[4 * (size_t)id + 1], hh * Sxy);
				red_add(&dL_dconic2D[4 * (size_t)id + 3], hh * Syy);
				red_add(&dL_dopacity[id], Sq);
			}
		}
		__syncwarp();   // the slab is rewritten from here on
	};

	const int batch_hi = (int)((tile_last - 1) / kBatchM);
	if (tid == 0) {
#pragma unroll
		for (int st = 0; st < kStagesM; st++)
			if (batch_hi - st >= 0)
				issue(batch_hi - st, st);
	}

	int st = 0;
	uint32_t parity = 0;
	for (int batch = batch_hi; batch >= 0; batch--) {
		mbar_wait(&s.full[st], parity);

		const int batch_base = batch * kBatchM;
		// positions >= warp_last are behind every pixel of this warp (backward.cu:487-489)
		const int cnt = min(min(kBatchM, (int)n - batch_base), (int)warp_last - batch_base);
		for (int base = (cnt > 0) ? ((cnt - 1) & ~31) : -1; base >= 0; base -= 32) {
			// cull 32 splats in parallel against the warp's 8x4 pixel block; compact the survivors back to front
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
				// one non-surviving lane pads an odd queue with a splat that no pixel accepts (position 2^32 - 1)
				const bool pad = !keep && (n_keep & 1) && lane == (__ffs(~mask) - 1);
				if (keep || pad) {
					const int at = keep ? __popc(mask >> lane) - 1 : n_keep;   // highest list position first
					const int slot = at >> 1, h = at & 1;
					const float2 bi = keep ? s.bid[st][j] : make_float2(0.0f, 0.0f);
					q.v[0][slot][h] = keep ? xr.x : 0.0f;
					q.v[0][slot][2 + h] = keep ? xr.y : 0.0f;
					q.v[1][slot][h] = keep ? co.x : 0.0f;
					q.v[1][slot][2 + h] = keep ? -co.y : 0.0f;
					q.v[2][slot][h] = keep ? co.z : 0.0f;
					q.v[2][slot][2 + h] = keep ? co.w : 0.0f;
					q.v[3][slot][h] = keep ? xr.z : 0.0f;
					q.v[3][slot][2 + h] = keep ? xr.w : 0.0f;
					q.v[4][slot][h] = bi.x;
					q.v[4][slot][2 + h] = __uint_as_float(keep ? (uint32_t)(batch_base + j) : 0xffffffffu);
					q.id[at] = __float_as_uint(bi.y);
				}
			}
			__syncwarp();
			const int n_pairs = (n_keep + 1) >> 1;
			uint32_t touched = 0, bit = 1u;
			uint32_t slot_addr = q_base, st_addr = slab_lane;
			for (int k = 0; k < n_pairs; k++, slot_addr += 16u) {
				const ulonglong2 XY = lds128p(slot_addr);
				const ulonglong2 AB = lds128p(slot_addr + kField);
				const ulonglong2 CO = lds128p(slot_addr + 2 * kField);
				const float4 BP = lds128(slot_addr + 4 * kField);
				// backward.cu:487-501, the forward's instruction sequence
				const f2 dx = add2(XY.x, npx), dy = add2(XY.y, npy);
				f2 t = mul2(dy, CO.x);
				const f2 u = mul2(dx, AB.x);
				t = mul2(dy, t);
				const f2 sq = fma2(dx, u, t);
				const f2 vv = mul2(dx, AB.y);
				const f2 ww = mul2(dy, vv);
				const f2 power = fma2(sq, neg_half, ww);
				const f2 G = exp2x(power);
				const f2 al = mul2(CO.y, G);
				const float aA = fminf(lo(al), 0.99f), aB = fminf(hi(al), 0.99f);
				const bool skipA = (__float_as_uint(BP.z) >= last_contributor) | (lo(power) > 0.0f) | (aA < 1.0f / 255.0f);
				const bool skipB = (__float_as_uint(BP.w) >= last_contributor) | (hi(power) > 0.0f) | (aB < 1.0f / 255.0f);
				if (!__all_sync(0xffffffffu, skipA & skipB)) {
					touched |= bit;
					const f2 e2 = pk(skipA ? 0.0f : aA, skipB ? 0.0f : aB);
					const f2 Ge = pk(skipA ? 0.0f : lo(G), skipB ? 0.0f : hi(G));
					// backward.cu:503-507: T <- T / (1 - alpha).  MUFU.RCP: 1 ulp, far inside the 1e-3 gradient tolerance
					const f2 om = fma2(e2, neg_one, one);
					const float rcpA = rcp_approx_ftz(lo(om)), rcpB = rcp_approx_ftz(hi(om));
					const float TA = T * rcpA, TB = TA * rcpB;
					T = TB;
					const f2 T2 = pk(TA, TB), rcp2 = pk(rcpA, rcpB);
					// backward.cu:509-521: accum_rec, walked A then B
					const ulonglong2 RG = lds128p(slot_addr + 3 * kField);
					const f2 col0 = RG.x, col1 = RG.y, col2 = pk(BP.x, BP.y);
					const f2 ac0 = mul2(e2, col0), ac1 = mul2(e2, col1), ac2 = mul2(e2, col2);
					const float accA0 = fmaf(keep_prev, acc0, pend0), accA1 = fmaf(keep_prev, acc1, pend1),
					            accA2 = fmaf(keep_prev, acc2, pend2);
					acc0 = fmaf(lo(om), accA0, lo(ac0));
					acc1 = fmaf(lo(om), accA1, lo(ac1));
					acc2 = fmaf(lo(om), accA2, lo(ac2));
					pend0 = hi(ac0); pend1 = hi(ac1); pend2 = hi(ac2);
					keep_prev = hi(om);
					const f2 d0 = fma2(pk(accA0, acc0), neg_one, col0);
					const f2 d1 = fma2(pk(accA1, acc1), neg_one, col1);
					const f2 d2 = fma2(pk(accA2, acc2), neg_one, col2);
					f2 dL_dalpha = fma2(d2, dLp2, fma2(d1, dLp1, mul2(d0, dLp0)));
					// backward.cu:526-534
					dL_dalpha = fma2(dL_dalpha, T2, mul2(bg_term, rcp2));
					// q = G dL/dalpha (backward.cu:537-554) and w = alpha T (backward.cu:523) of this pixel: rows 2 r (splat A)
					// and 2 r + 1 (splat B) of the slab
					const f2 qq = mul2(Ge, dL_dalpha), wt = mul2(e2, T2);
					sts64(st_addr, lo(qq), lo(wt));
					sts64(st_addr + kSlabPitch * 4u, hi(qq), hi(wt));
				}
				bit <<= 1;
				st_addr += 2u * kSlabPitch * 4u;
				if (bit == (1u << kFlushPairs) || k == n_pairs - 1) {
					if (touched != 0)
						flush(k & ~(kFlushPairs - 1), touched);
					touched = 0;
					bit = 1u;
					st_addr = slab_lane;
				}
			}
			__syncwarp();   // the queue is rewritten by the next chunk
		}

		// release the stage; the warp that arrives last refills it with the batch kStagesM further on
		__syncwarp();
		mbar_arrive(&s.empty[st]);                                       // release: this thread's reads of the stage are done
		if (lane == 0) {
			const uint32_t old = atomicAdd(&s.released[st], 1u);
			if ((old & (kWarps - 1)) == kWarps - 1 && batch - kStagesM >= 0) {
				mbar_wait(&s.empty[st], parity);                          // acquire: every thread's arrival (returns at once)
				fence_proxy_async();
				issue(batch - kStagesM, st);
			}
		}
		if (++st == kStagesM) {
			st = 0;
			parity ^= 1u;
		}
	}
}


// ---- column sums: lanes own splats when the sums leave the warp -------------------------------------
// The nine sums over the warp's 32 pixels are not formed per splat with shuffles.  Each lane parks (q, w) = (G dL/dalpha,
// alpha T) of its pixel in a warp-private slab, one row per evaluated splat, ACROSS the 32-splat culling chunks; when 16
// rows are full the roles turn: lane l takes splat l & 15 and walks 16 of its 32 pixels (half l >> 4 of the block), forming
//     sum_p q[s, p] * (1, u, v, u^2, u v, v^2)[p]      and      sum_p w[s, p] * dL/dpixel[p, :]
// with u, v the pixel's offset from the block centre (compile-time constants of the unrolled loop; a row of eight pixels is
// first reduced to three partial sums).  One exchange joins the two halves, and lanes 0-15 shift the pixel-centred moments
// to their splat (dx = X - u), apply the splat-constant factors of backward.cu:537-554 and send the gradients (vector REDs).
// Against the 18-value butterfly of the pairs kernel (about 37 instructions per splat) this costs about 14 per splat, and
// nothing in it sits on the alpha / T recurrence.  The survivors of the culling are evaluated in QUADS (two pairs per
// iteration, no branch inside: the alpha test of the second pair overlaps the recurrences of the first) formed across
// chunk boundaries: up to three survivors are carried to the next chunk instead of being padded.
constexpr int kSlabRowsC = 16;
constexpr int kSlabPitchC = 66;   // floats per slab row: 32 pixels x (q, w) + 2 (rows 8 bytes apart modulo 128: conflict-free column reads)

struct WarpQueueC {
	// [field][slot][4]: slot k holds entries 2k (A, further back) and 2k+1 (B) of the compacted survivors, back to front
	//   0: xA xB yA yB   1: aA aB -bA -bB   2: cA cB oA oB   3: rA rB gA gB   4: bA bB posA posB (0-based, as bits)
	//   5: idA idB - -   (35 entries: up to 32 survivors of a chunk behind up to three carried over)
	float v[6][18][4];
	float slab[kSlabRowsC * kSlabPitchC];
	float table[kSlabRowsC / 2][17];   // per parked pair: fields 0, 1, 2, 5 of its queue slot (what the flush needs); odd pitch
	float4 dlp[33];                    // dL/dpixel of the warp's 32 pixels, pixel p at p + (p >> 4): the two halves on different banks
};

template <int kBatchC, int kWarpsC>
struct __align__(128) BwdSmemC {
	float4 conic[2][kBatchC];
	float4 xyrg[2][kBatchC];
	float2 bid[2][kBatchC];
	uint32_t cmask[2][(kBatchC / 32) * 8];   // the forward's survivor masks of the staged chunks, [chunk][warp block of the tile]
	WarpQueueC queue[kWarpsC];
	uint64_t full[2];
	uint32_t warp_max[kWarpsC];
};

// kWarpsC = 8: one block per 16x16 tile; kWarpsC = 4: two blocks per tile, each owning a 16x8 half (more, smaller blocks:
// finer register / occupancy steps, a barrier among four warps instead of eight, list trimming per half tile)
template <int kBatchC, int kMinBlocks, int kWarpsC>
__global__ void __launch_bounds__(kWarpsC * 32, kMinBlocks)
blend_backward_cols_kernel(GeometryState g, BinningState b, ImageState img, uint32_t capacity,
                           int W, int H, int tiles_x, const float* __restrict__ bg_color,
                           const float* __restrict__ dL_dpixels,
                           float* __restrict__ dL_dmean2D,   // [P,3]
                           float* __restrict__ dL_dconic2D,  // [P,4]
                           float* __restrict__ dL_dopacity,  // [P]
                           float* __restrict__ dL_dcolors)   // [P,3]
{
	pdl_sync();
	extern __shared__ __align__(128) unsigned char smem_raw[];
	BwdSmemC<kBatchC, kWarpsC>& s = *reinterpret_cast<BwdSmemC<kBatchC, kWarpsC>*>(smem_raw);

	constexpr int kPerTile = kWarps / kWarpsC;          // blocks per tile
	const int tile = blockIdx.x / kPerTile;
	const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tw = (blockIdx.x % kPerTile) * kWarpsC + warp;   // the warp's 8x4 block inside the tile

	const int bx0 = tile_x * kTile + (tw & 1) * 8;
	const int by0 = tile_y * kTile + (tw >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const uint32_t pix_id = (uint32_t)W * py + px;
	const f2 npx = pk1(-(float)px), npy = pk1(-(float)py);
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

	uint32_t warp_last = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
	if (lane == 0)
		s.warp_max[warp] = warp_last;
	if (tid == 0) {
		mbar_init(&s.full[0], 1);
		mbar_init(&s.full[1], 1);
		fence_mbar_init();
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int w = 0; w < kWarpsC; w++)
		tile_last = max(tile_last, s.warp_max[w]);
	if (tile_last == 0)
		return;

	float dL_dpixel0 = 0.0f, dL_dpixel1 = 0.0f, dL_dpixel2 = 0.0f;
	if (inside) {
		const size_t HW = (size_t)H * W;
		dL_dpixel0 = dL_dpixels[0 * HW + pix_id];
		dL_dpixel1 = dL_dpixels[1 * HW + pix_id];
		dL_dpixel2 = dL_dpixels[2 * HW + pix_id];
	}
	WarpQueueC& q = s.queue[warp];
	q.dlp[lane + (lane >> 4)] = make_float4(dL_dpixel0, dL_dpixel1, dL_dpixel2, 0.0f);

	// accum_rec and the pending (last_alpha * last_color, 1 - last_alpha) term of backward.cu:509-515
	float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
	float pend0 = 0.0f, pend1 = 0.0f, pend2 = 0.0f, keep_prev = 1.0f;

	// backward.cu:455-461: d(pixel offset)/d(NDC mean); applied once per (warp, splat) in flush
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	// backward.cu:531-534: dL/dalpha += (-T_final / (1 - alpha)) * (bg . dL/dpixel)
	const f2 bg_term = pk1(-T_final * (bg_color[0] * dL_dpixel0 + bg_color[1] * dL_dpixel1 + bg_color[2] * dL_dpixel2));
	const f2 dLp0 = pk1(dL_dpixel0), dLp1 = pk1(dL_dpixel1), dLp2 = pk1(dL_dpixel2);
	const f2 neg_half = pk1(-0.5f), neg_one = pk1(-1.0f), one = pk1(1.0f);

	// the forward left the survivor masks of its culling in the key array (state.h: cull_mask_fits): same test, same answer
	const bool have_masks = g.header->cull_masks != 0u;
	const uint32_t* const cull_masks = reinterpret_cast<const uint32_t*>(b.keys) + ((size_t)(start >> 5) + (size_t)tile) * 8;
	constexpr uint32_t kMaskBytes = (kBatchC / 32) * 8 * 4;
	auto issue = [&](int batch, int buf) {
		const uint32_t off = start + (uint32_t)batch * kBatchC;
		const uint32_t cnt = min((uint32_t)kBatchC, n - (uint32_t)batch * kBatchC);
		const uint32_t cnt4 = (cnt + 3u) & ~3u;
		mbar_arrive_expect_tx(&s.full[buf], cnt4 * 40u + (have_masks ? kMaskBytes : 0u));
		if (have_masks)
			bulk_g2s(s.cmask[buf], cull_masks + (size_t)batch * (kBatchC / 32) * 8, kMaskBytes, &s.full[buf]);
		bulk_g2s(s.conic[buf], b.rec_conic + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.xyrg[buf], b.rec_xyrg + off, cnt4 * 16u, &s.full[buf]);
		bulk_g2s(s.bid[buf], b.rec_bid + off, cnt4 * 8u, &s.full[buf]);
	};

	// vector REDs need 8- / 16-byte aligned gradient arrays (true for every allocation of the host side; checked, not assumed)
	const bool vec_red = (((uintptr_t)dL_dcolors | (uintptr_t)dL_dmean2D) & 7u) == 0 && ((uintptr_t)dL_dconic2D & 15u) == 0;
	// Shared-window addresses of everything the inner loops touch, made opaque so that they stay in registers: left to
	// itself ptxas re-derives them from %tid and %cgaid in every iteration (S2R, a ~100-cycle instruction).
	uint32_t q_base = smem_u32(&q);
	// this lane's word of a queue slot that is copied to the table when the slot's pair is parked (lanes 0-15: fields 0, 1, 2, 5)
	// (lanes 0-15 serve the first pair of a quad, lanes 16-31 the second: next slot, next table row)
	uint32_t copy_src = q_base + (uint32_t)((((lane >> 2) & 3) == 3 ? 5 : ((lane >> 2) & 3)) * 18 * 16 + (lane & 3) * 4 + (lane >> 4) * 16);
	uint32_t slab_lane = q_base + (uint32_t)offsetof(WarpQueueC, slab) + (uint32_t)lane * 8u;
	uint32_t table_lane = q_base + (uint32_t)offsetof(WarpQueueC, table) + (uint32_t)(lane & 15) * 4u + (uint32_t)(lane >> 4) * 68u;
	asm volatile("" : "+r"(q_base), "+r"(copy_src), "+r"(slab_lane), "+r"(table_lane));
	constexpr uint32_t kField = 18 * 16;   // bytes per queue field
	const float cx = (float)bx0 + 3.5f, cy = (float)by0 + 1.5f;

	// Sum the parked rows (rows 0 .. rows-1, `rows` even and warp-uniform) over the pixels and send each splat's gradients.
	auto flush = [&](int rows) {
		__syncwarp();
		const int half = lane >> 4;
		// row (lane & 15) of the slab, pixels 16 half .. 16 half + 15
		const uint32_t row_addr = slab_lane + (uint32_t)(lane & 15) * (kSlabPitchC * 4u - 8u);
		const uint32_t dl_addr = q_base + (uint32_t)offsetof(WarpQueueC, dlp) + (uint32_t)half * 272u;
		float Sq = 0.0f, Su = 0.0f, Sv = 0.0f, Suu = 0.0f, Suv = 0.0f, Svv = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
#pragma unroll
		for (int r = 0; r < 2; r++) {
			float R0 = 0.0f, R1 = 0.0f, R2 = 0.0f;
#pragma unroll
			for (int u = 0; u < 8; u++) {
				const uint2 qwb = lds64u(row_addr + 8u * (8 * r + u));
				const float2 qw = make_float2(__uint_as_float(qwb.x), __uint_as_float(qwb.y));
				const float4 d = lds128(dl_addr + 16u * (8 * r + u));
				const float uu = (float)u - 3.5f;
				R0 += qw.x;
				R1 = fmaf(qw.x, uu, R1);
				R2 = fmaf(qw.x, uu * uu, R2);
				c0 = fmaf(qw.y, d.x, c0);
				c1 = fmaf(qw.y, d.y, c1);
				c2 = fmaf(qw.y, d.z, c2);
			}
			const float vv = (float)(2 * half + r) - 1.5f;
			Sq += R0;
			Su += R1;
			Suu += R2;
			Sv = fmaf(vv, R0, Sv);
			Suv = fmaf(vv, R1, Suv);
			Svv = fmaf(vv * vv, R0, Svv);
		}
		Sq += __shfl_xor_sync(0xffffffffu, Sq, 16);
		Su += __shfl_xor_sync(0xffffffffu, Su, 16);
		Sv += __shfl_xor_sync(0xffffffffu, Sv, 16);
		Suu += __shfl_xor_sync(0xffffffffu, Suu, 16);
		Suv += __shfl_xor_sync(0xffffffffu, Suv, 16);
		Svv += __shfl_xor_sync(0xffffffffu, Svv, 16);
		c0 += __shfl_xor_sync(0xffffffffu, c0, 16);
		c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
		c2 += __shfl_xor_sync(0xffffffffu, c2, 16);
		if (lane < rows) {
			const uint32_t any = (__float_as_uint(c0) | __float_as_uint(c1) | __float_as_uint(c2) | __float_as_uint(Sq) |
			                      __float_as_uint(Su) | __float_as_uint(Sv) | __float_as_uint(Suu) | __float_as_uint(Suv) |
			                      __float_as_uint(Svv)) << 1;
			if (any != 0) {
				const uint32_t t = table_lane + (uint32_t)(lane >> 1) * 60u;   // table[lane >> 1] + (lane & 1)
				const float a = lds32(t + 16u), bb = -lds32(t + 24u), c = lds32(t + 32u), o = lds32(t + 40u);
				const uint32_t id = __float_as_uint(lds32(t + 48u));
				// dx = x_s - px = X - u with X = x_s - (block centre): shift the pixel-centred moments to the splat
				const float X = lds32(t) - cx, Y = lds32(t + 8u) - cy;
				const float Sx = fmaf(X, Sq, -Su), Sy = fmaf(Y, Sq, -Sv);
				const float Sxx = fmaf(X, fmaf(X, Sq, -2.0f * Su), Suu);
				const float Sxy = fmaf(X, fmaf(Y, Sq, -Sv), fmaf(-Y, Su, Suv));
				const float Syy = fmaf(Y, fmaf(Y, Sq, -2.0f * Sv), Svv);
				// dL/dG = o dL/dalpha;  dG/ddelx = -G (a dx + b dy);  dG/ddely = -G (c dy + b dx)
				const float gx = -o * ddelx_dx * (a * Sx + bb * Sy), gy = -o * ddely_dy * (c * Sy + bb * Sx);
				const float hh = -0.5f * o;
				if (vec_red) {
					// rows of three floats start 8-byte aligned for even ids only: the aligned pair goes out as one request,
					// the third value alone.  dL/dmean2D.z is not written by this pass: adding 0 to it is free.
					const bool odd = id & 1u;
					float* const pc = dL_dcolors + 3 * (size_t)id;
					red_add2(pc + (odd ? 1 : 0), odd ? c1 : c0, odd ? c2 : c1);
					red_add(pc + (odd ? 0 : 2), odd ? c0 : c2);
					float* const pm = dL_dmean2D + 3 * (size_t)id;
					red_add2(pm + (odd ? 1 : 0), odd ? gy : gx, odd ? 0.0f : gy);
					if (odd)
						red_add(pm, gx);
					red_add4(dL_dconic2D + 4 * (size_t)id, hh * Sxx, hh * Sxy, 0.0f, hh * Syy);
				} else {
					red_add(&dL_dcolors[3 * (size_t)id + 0], c0);
					red_add(&dL_dcolors[3 * (size_t)id + 1], c1);
					red_add(&dL_dcolors[3 * (size_t)id + 2], c2);
					red_add(&dL_dmean2D[3 * (size_t)id + 0], gx);
					red_add(&dL_dmean2D[3 * (size_t)id + 1], gy);
					red_add(&dL_dconic2D[4 * (size_t)id + 0], hh * Sxx);
					red_add(&dL_dconic2D[4 * (size_t)id + 1], hh * Sxy);
					red_add(&dL_dconic2D[4 * (size_t)id + 3], hh * Syy);
				}
				red_add(&dL_dopacity[id], Sq);
			}
		}
		__syncwarp();   // slab and table are rewritten from here on
	};

	const int batch_hi = (int)((tile_last - 1) / kBatchC);
	if (tid == 0)
		issue(batch_hi, 0);

	// One pair of splats up to the alpha test (backward.cu:487-501, the forward's instruction sequence); independent of the
	// T recurrence, so the two pairs of a quad overlap.
	struct Front {
		f2 e2, Ge;        // alpha and G of the two splats, 0 where the pixel skips the splat
	};
	auto front = [&](uint32_t slot_addr) -> Front {
		const ulonglong2 XY = lds128p(slot_addr);
		const ulonglong2 AB = lds128p(slot_addr + kField);
		const ulonglong2 CO = lds128p(slot_addr + 2 * kField);
		const uint2 PP = lds64u(slot_addr + 4 * kField + 8u);
		const f2 dx = add2(XY.x, npx), dy = add2(XY.y, npy);
		f2 t = mul2(dy, CO.x);
		const f2 u = mul2(dx, AB.x);
		t = mul2(dy, t);
		const f2 sq = fma2(dx, u, t);
		const f2 vv = mul2(dx, AB.y);
		const f2 ww = mul2(dy, vv);
		const f2 power = fma2(sq, neg_half, ww);
		const f2 G = exp2x(power);
		const f2 al = mul2(CO.y, G);
		const float aA = fminf(lo(al), 0.99f), aB = fminf(hi(al), 0.99f);
		const bool skipA = (PP.x >= last_contributor) | (lo(power) > 0.0f) | (aA < 1.0f / 255.0f);
		const bool skipB = (PP.y >= last_contributor) | (hi(power) > 0.0f) | (aB < 1.0f / 255.0f);
		Front f;
		f.e2 = pk(skipA ? 0.0f : aA, skipB ? 0.0f : aB);
		f.Ge = pk(skipA ? 0.0f : lo(G), skipB ? 0.0f : hi(G));
		return f;
	};
	// The recurrences of backward.cu:503-534 for the pair, A then B; parks (q, w) of both splats in slab rows `at`, `at + 1`.
	// A splat the pixel skips is carried with alpha = 0 and G = 0: T / (1 - 0) == T, the pending (last_alpha, last_color) term
	// is folded into accum_rec one splat early, q = w = 0.
	auto back = [&](uint32_t slot_addr, const Front& f, uint32_t dst) {
		const f2 e2 = f.e2;
		// backward.cu:503-507: T <- T / (1 - alpha).  MUFU.RCP: 1 ulp, far inside the 1e-3 gradient tolerance
		const f2 om = fma2(e2, neg_one, one);
		const float rcpA = rcp_approx_ftz(lo(om)), rcpB = rcp_approx_ftz(hi(om));
		const float TA = T * rcpA, TB = TA * rcpB;
		T = TB;
		const f2 T2 = pk(TA, TB), rcp2 = pk(rcpA, rcpB);
		// backward.cu:509-521: accum_rec, walked A then B
		const ulonglong2 RG = lds128p(slot_addr + 3 * kField);
		const f2 col2 = lds128p(slot_addr + 4 * kField).x;
		const f2 col0 = RG.x, col1 = RG.y;
		const f2 ac0 = mul2(e2, col0), ac1 = mul2(e2, col1), ac2 = mul2(e2, col2);
		const float accA0 = fmaf(keep_prev, acc0, pend0), accA1 = fmaf(keep_prev, acc1, pend1),
		            accA2 = fmaf(keep_prev, acc2, pend2);
		acc0 = fmaf(lo(om), accA0, lo(ac0));
		acc1 = fmaf(lo(om), accA1, lo(ac1));
		acc2 = fmaf(lo(om), accA2, lo(ac2));
		pend0 = hi(ac0); pend1 = hi(ac1); pend2 = hi(ac2);
		keep_prev = hi(om);
		const f2 d0 = fma2(pk(accA0, acc0), neg_one, col0);
		const f2 d1 = fma2(pk(accA1, acc1), neg_one, col1);
		const f2 d2 = fma2(pk(accA2, acc2), neg_one, col2);
		f2 dL_dalpha = fma2(d2, dLp2, fma2(d1, dLp1, mul2(d0, dLp0)));
		// backward.cu:526-534
		dL_dalpha = fma2(dL_dalpha, T2, mul2(bg_term, rcp2));
		// q = G dL/dalpha (backward.cu:537-554) and w = alpha T (backward.cu:523) of this pixel
		const f2 qq = mul2(f.Ge, dL_dalpha), wt = mul2(e2, T2);
		sts64(dst, lo(qq), lo(wt));
		sts64(dst + kSlabPitchC * 4u, hi(qq), hi(wt));
	};

	int carry = 0;      // queue entries 0 .. carry-1 hold survivors of earlier chunks that have not been evaluated yet (< 4)
	int row = 0;        // parked slab rows
	for (int it = 0, batch = batch_hi; batch >= 0; it++, batch--) {
		const int buf = it & 1;
		// every warp has finished batch+1 (buffer buf^1) before it is overwritten
		if (it > 0)
			__syncthreads();
		if (tid == 0 && batch > 0)
			issue(batch - 1, buf ^ 1);
		mbar_wait(&s.full[buf], (uint32_t)(it >> 1) & 1u);

		const int batch_base = batch * kBatchC;
		// positions >= warp_last are behind every pixel of this warp (backward.cu:487-489)
		const int cnt = min(min(kBatchC, (int)n - batch_base), (int)warp_last - batch_base);
		for (int base = (cnt > 0) ? ((cnt - 1) & ~31) : -1; base >= 0; base -= 32) {
			// cull 32 splats in parallel against the warp's 8x4 pixel block; append the survivors back to front
			const int j = base + lane;
			bool keep = false;
			float4 co, xr;
			uint32_t mask;
			if (have_masks) {
				mask = s.cmask[buf][(base >> 5) * 8 + tw];
				if (cnt - base < 32)
					mask &= (1u << (cnt - base)) - 1u;       // records behind this warp's last contributor
				keep = (mask >> lane) & 1u;
				if (keep) {
					co = s.conic[buf][j];
					xr = s.xyrg[buf][j];
				}
			} else {
				if (j < cnt) {
					co = s.conic[buf][j];
					xr = s.xyrg[buf][j];
					keep = !rect_cannot_contribute(xr.x, xr.y, co.x, co.y, co.z, cull_threshold(co.w),
					                               wx0, wy0, wx1, wy1);
				}
				mask = __ballot_sync(0xffffffffu, keep);
			}
			const bool last_chunk = (batch | base) == 0;
			const int n_keep = __popc(mask);
			int total = carry + n_keep;
			if (total == 0 || (n_keep == 0 && !last_chunk))
				continue;
			if (keep) {
				const int at = carry + __popc(mask >> lane) - 1;   // highest list position first
				const int slot = at >> 1, h = at & 1;
				const float2 bi = s.bid[buf][j];
				q.v[0][slot][h] = xr.x;
				q.v[0][slot][2 + h] = xr.y;
				q.v[1][slot][h] = co.x;
				q.v[1][slot][2 + h] = -co.y;
				q.v[2][slot][h] = co.z;
				q.v[2][slot][2 + h] = co.w;
				q.v[3][slot][h] = xr.z;
				q.v[3][slot][2 + h] = xr.w;
				q.v[4][slot][h] = bi.x;
				q.v[4][slot][2 + h] = __uint_as_float((uint32_t)(batch_base + j));
				q.v[5][slot][h] = bi.y;
			}
			if (last_chunk && (total & 3)) {
				// the list's very end is filled up to a whole quad with splats that no pixel accepts (position 2^32 - 1)
				const int pads = (-total) & 3;
				if (lane < 6 * pads) {
					const int e = total + lane / 6, f = lane % 6;
					q.v[f][e >> 1][e & 1] = 0.0f;
					q.v[f][e >> 1][2 + (e & 1)] = f == 4 ? __uint_as_float(0xffffffffu) : 0.0f;
				}
				total += pads;
			}
			__syncwarp();
			const int n_quads = total >> 2;
			uint32_t slot_addr = q_base;
			for (int k = 0; k < n_quads; k++, slot_addr += 32u) {
				// this lane's word of the two slots for the flush table
				const float copy_word = lds32(copy_src + (uint32_t)k * 32u);
				const Front f0 = front(slot_addr), f1 = front(slot_addr + 16u);
				const uint32_t dst = slab_lane + (uint32_t)row * (kSlabPitchC * 4u);
				back(slot_addr, f0, dst);
				back(slot_addr + 16u, f1, dst + 2u * kSlabPitchC * 4u);
				sts32(table_lane + (uint32_t)row * 34u, copy_word);
				row += 4;
				if (row == kSlabRowsC) {
					flush(kSlabRowsC);
					row = 0;
				}
			}
			__syncwarp();   // every lane is through with the queue
			carry = total & 3;
			if (carry != 0 && n_quads > 0) {
				// the survivors behind the last whole quad (slots 2 n_quads, 2 n_quads + 1) move to the front of the next chunk's list
				if (lane < 24) {
					const int f = lane >> 2, w = lane & 3;
					q.v[f][0][w] = q.v[f][2 * n_quads][w];
					q.v[f][1][w] = q.v[f][2 * n_quads + 1][w];
				}
				__syncwarp();
			}
		}
	}
	if (row > 0)
		flush(row);
}

} // namespace

int launch_blend_backward(const GeometryState& g, const BinningState& b, const ImageState& img, uint32_t capacity,
                          const ViewParams& vp, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                          float* dL_dopacity, float* dL_dcolor, cudaStream_t stream)
{
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	if (num_tiles <= 0)
		return GM_OK;
	// The default is the column-sum kernel (two blocks per tile).  GM_BLEND_BWD=pairs selects the packed-pair kernel with
	// the shuffle butterfly (the default of round 1 and most of round 2), GM_BLEND_BWD=mma the variant that reduces over the
	// pixels on the tensor cores, GM_BLEND_SCALAR=1 the one-splat-per-iteration kernel; all kept for A/B measurements
	// (DESIGN.md 8).
	static const bool scalar = std::getenv("GM_BLEND_SCALAR") != nullptr && std::getenv("GM_BLEND_SCALAR")[0] == '1';
	static const char variant = std::getenv("GM_BLEND_BWD") != nullptr ? std::getenv("GM_BLEND_BWD")[0] : 'c';
	if (scalar)
		launch_k(blend_backward_kernel, dim3(num_tiles), dim3(kThreads), 0, stream, 
			g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
	else if (variant == 'p')
		launch_k(blend_backward_pairs_kernel, dim3(num_tiles), dim3(kThreads), 0, stream, 
			g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
	else if (variant != 'm') {
		constexpr int kWarpsC = 4;
		auto kern = blend_backward_cols_kernel<128, 5, kWarpsC>;
		constexpr size_t smem = sizeof(BwdSmemC<128, kWarpsC>);
		// the opt-in shared-memory size is a per-device function attribute
		static bool opted_in_c[64] = {};
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev < 0 || dev >= 64 || !opted_in_c[dev]) {
			const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (attr != cudaSuccess) {
				set_last_error("blend_backward shared memory", attr);
				return GM_ERR_CUDA;
			}
			if (dev >= 0 && dev < 64)
				opted_in_c[dev] = true;
		}
		launch_k(kern, dim3(num_tiles * (kWarps / kWarpsC)), dim3(kWarpsC * 32), smem, stream,
			g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
	}
	else {
		// the opt-in shared-memory size is a per-device function attribute
		static bool opted_in[64] = {};
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev < 0 || dev >= 64 || !opted_in[dev]) {
			const cudaError_t attr = cudaFuncSetAttribute(blend_backward_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                              (int)sizeof(BwdSmemM));
			if (attr != cudaSuccess) {
				set_last_error("blend_backward shared memory", attr);
				return GM_ERR_CUDA;
			}
			if (dev >= 0 && dev < 64)
				opted_in[dev] = true;
		}
		launch_k(blend_backward_mma_kernel, dim3(num_tiles), dim3(kThreads), sizeof(BwdSmemM), stream, 
			g, b, img, capacity, vp.W, vp.H, vp.tiles_x, vp.bg, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
	}
	return GM_OK;
}

} // namespace gm
