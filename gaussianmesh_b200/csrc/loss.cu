// Training-loss kernels around the rasterizer op: the photometric loss of
// train_mesh_gaussian.py:91-94, (1 - lambda) L1 + lambda (1 - SSIM), with its gradient, and the mesh-restrict
// regulariser.
//
//   l1 / ssim            utils/loss_utils.py:17-18,36-82: 11x11 Gaussian window (sigma 1.5), zero padding, per channel,
//                        C1 = 0.01^2, C2 = 0.03^2, mean over all elements.  The reference runs five grouped conv2d
//                        launches plus ~15 elementwise kernels (and their autograd mirrors); here one kernel produces
//                        the L1 sum, the SSIM sum and the three partial-derivative maps of the SSIM map, a second
//                        convolves those maps into dL/dimg and adds the L1 sign term.  Separable passes through shared
//                        memory, one 16x16 output tile per block; the last block to finish folds the two sums into
//                        (loss, L1, SSIM), so the whole loss is two launches and no host round trip.
//   mesh_restrict_loss   utils/loss_utils.py:84-107: sum(max(0, max_k scale_k - weight * sqrt(|(v2-v1) x (v3-v1)|)))
#include "common.cuh"
#include "packed.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kT = 32;            // output tile edge
constexpr int kR = 5;             // window radius (11 taps)
constexpr int kIn = kT + 2 * kR;  // 42
constexpr int kSeg = 4;           // outputs per thread along the filter direction (register blocking)
constexpr int kPT = 256;          // threads per block
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// gaussian(11, 1.5) of utils/loss_utils.py:23-25, normalised
__constant__ float c_win[11] = {0.00102838f, 0.00759876f, 0.03600077f, 0.10936069f, 0.21300554f, 0.26601172f,
                                0.21300554f, 0.10936069f, 0.03600077f, 0.00759876f, 0.00102838f};

// The same window as (tap t, tap t - 1) pairs, t = 0 .. 11 (taps -1 and 11 are zero): two ADJACENT outputs j, j + 1 read
// input j + t through taps t and t - 1, so one packed FFMA2 (sm_100a) advances both -- 12 instead of 22 FMA
// instructions per pair of outputs, each half rounding exactly like the scalar chain (a zero tap adds an exact zero).
__constant__ float2 c_win2[12] = {{0.00102838f, 0.0f},        {0.00759876f, 0.00102838f}, {0.03600077f, 0.00759876f},
                                  {0.10936069f, 0.03600077f}, {0.21300554f, 0.10936069f}, {0.26601172f, 0.21300554f},
                                  {0.21300554f, 0.26601172f}, {0.10936069f, 0.21300554f}, {0.03600077f, 0.10936069f},
                                  {0.00759876f, 0.03600077f}, {0.00102838f, 0.00759876f}, {0.0f, 0.00102838f}};

__device__ __forceinline__ f2 win2(int t)
{
	return pk(c_win2[t].x, c_win2[t].y);
}

// head of the caller's scratch chunk; the three derivative maps follow at kMapsOffset
struct PhotoSums {
	float l1, ssim;
	unsigned int done;
	unsigned int pad;
};
constexpr size_t kMapsOffset = 128;

__device__ __forceinline__ float block_sum(float v, float* warp_part)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	const int tid = threadIdx.x;
	__syncthreads();                                           // warp_part may still be read from a previous call
	if ((tid & 31) == 0) warp_part[tid >> 5] = v;
	__syncthreads();
	float s = 0.0f;
	if (tid == 0)
		for (int w = 0; w < kPT / 32; w++) s += warp_part[w];
	return s;
}

// Both kernels filter a 32x32 output tile (42x42 with the halo) separably through shared memory.  Every thread
// produces kSeg = 4 adjacent outputs along the filter direction from 14 staged inputs, so a tap costs one FMA and
// 0.3 shared-memory loads instead of one FMA and one load.
//   horizontal pass: work item = (row of the 42, group of 4 columns): 42 x 8 items over 256 threads
//   vertical pass:   work item = (group of 4 rows, column): warp w owns rows 4w..4w+3, lane = column
// Stage the 42x42 input tile of M planes (zero outside the image = conv2d's padding): warp w takes rows w, w+8, ...,
// three rows at a time with every global load issued before the first shared-memory store.
template <int M>
__device__ __forceinline__ void load_tile(float (*dst)[kIn][kIn + 1], const float* const (&src)[M], int H, int W,
                                          int x0, int y0, int tid)
{
	const int lane = tid & 31, warp = tid >> 5;
	const int gx_a = x0 - kR + lane, gx_b = gx_a + 32;
	const bool in_a = gx_a >= 0 && gx_a < W, in_b = lane + 32 < kIn && gx_b >= 0 && gx_b < W;
#pragma unroll
	for (int r0 = 0; r0 < kIn; r0 += 3 * (kPT / 32)) {
		float va[3][M], vb[3][M];
#pragma unroll
		for (int u = 0; u < 3; u++) {
			const int ly = r0 + u * (kPT / 32) + warp;
			const int gy = y0 - kR + ly;
			const bool row = ly < kIn && gy >= 0 && gy < H;
			const size_t base = (size_t)gy * W;
#pragma unroll
			for (int m = 0; m < M; m++) {
				va[u][m] = (row && in_a) ? __ldg(src[m] + base + gx_a) : 0.0f;
				vb[u][m] = (row && in_b) ? __ldg(src[m] + base + gx_b) : 0.0f;
			}
		}
#pragma unroll
		for (int u = 0; u < 3; u++) {
			const int ly = r0 + u * (kPT / 32) + warp;
			if (ly < kIn) {
#pragma unroll
				for (int m = 0; m < M; m++) {
					dst[m][ly][lane] = va[u][m];
					if (lane + 32 < kIn) dst[m][ly][lane + 32] = vb[u][m];
				}
			}
		}
	}
}

template <int Q>
__device__ __forceinline__ void vertical4(const float (*s_h)[kIn][kT + 1], int row0, int col, float (&out)[Q][kSeg])
{
#pragma unroll
	for (int q = 0; q < Q; q++) {
		float v[kSeg + 10];
#pragma unroll
		for (int k = 0; k < kSeg + 10; k++)
			v[k] = s_h[q][row0 + k][col];
#pragma unroll
		for (int j = 0; j < kSeg; j += 2) {
			f2 acc = pk1(0.0f);
#pragma unroll
			for (int t = 0; t < 12; t++)
				acc = fma2(win2(t), pk1(v[j + t]), acc);
			out[q][j] = lo(acc);
			out[q][j + 1] = hi(acc);
		}
	}
}

__global__ void __launch_bounds__(kPT)
photometric_forward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                           PhotoSums* __restrict__ sums, float* __restrict__ dmaps /* [3][C][H][W] or null */,
                           size_t plane_all, float lambda, float inv_numel, float* __restrict__ out /* [3] */)
{
	pdl_sync();
	__shared__ float s_in[2][kIn][kIn + 1];
	__shared__ float s_h[5][kIn][kT + 1];
	__shared__ float warp_part[kPT / 32];
	float (*const s_x)[kIn + 1] = s_in[0];
	float (*const s_y)[kIn + 1] = s_in[1];
	const int ch = blockIdx.z;
	const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
	const size_t plane = (size_t)H * W;
	const int tid = threadIdx.x;
	{
		const float* const src[2] = {img1 + ch * plane, img2 + ch * plane};
		load_tile<2>(s_in, src, H, W, x0, y0, tid);
	}
	__syncthreads();
	// horizontal pass, five quantities: x, y, x^2, y^2, xy
	for (int i = tid; i < kIn * (kT / kSeg); i += kPT) {
		const int ly = i / (kT / kSeg), c0 = (i % (kT / kSeg)) * kSeg;
		float u[kSeg + 10], v[kSeg + 10];
#pragma unroll
		for (int k = 0; k < kSeg + 10; k++) {
			u[k] = s_x[ly][c0 + k];
			v[k] = s_y[ly][c0 + k];
		}
		f2 acc[5][kSeg / 2];
#pragma unroll
		for (int j = 0; j < kSeg / 2; j++)
#pragma unroll
			for (int q = 0; q < 5; q++) acc[q][j] = pk1(0.0f);
#pragma unroll
		for (int k = 0; k < kSeg + 10; k++) {
			const float xx = u[k] * u[k], yy = v[k] * v[k], xy = u[k] * v[k];
#pragma unroll
			for (int j = 0; j < kSeg / 2; j++) {
				const int t = k - 2 * j;          // outputs 2j and 2j + 1 read input k through taps t and t - 1
				if (t >= 0 && t < 12) {
					const f2 w = win2(t);
					acc[0][j] = fma2(w, pk1(u[k]), acc[0][j]); acc[1][j] = fma2(w, pk1(v[k]), acc[1][j]);
					acc[2][j] = fma2(w, pk1(xx), acc[2][j]); acc[3][j] = fma2(w, pk1(yy), acc[3][j]);
					acc[4][j] = fma2(w, pk1(xy), acc[4][j]);
				}
			}
		}
#pragma unroll
		for (int q = 0; q < 5; q++)
#pragma unroll
			for (int j = 0; j < kSeg / 2; j++) {
				s_h[q][ly][c0 + 2 * j] = lo(acc[q][j]);
				s_h[q][ly][c0 + 2 * j + 1] = hi(acc[q][j]);
			}
	}
	__syncthreads();
	const int lx = tid & 31, ly0 = (tid >> 5) * kSeg;
	const int gx = x0 + lx;
	float mom[5][kSeg];
	vertical4<5>(s_h, ly0, lx, mom);
	float val_sum = 0.0f, l1_sum = 0.0f;
#pragma unroll
	for (int j = 0; j < kSeg; j++) {
		const int gy = y0 + ly0 + j;
		if (gx < W && gy < H) {
			const float mu1 = mom[0][j], mu2 = mom[1][j];
			const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
			const float s1 = mom[2][j] - mu1_sq, s2 = mom[3][j] - mu2_sq, s12 = mom[4][j] - mu12;
			const float A1 = 2.0f * mu12 + kC1, A2 = 2.0f * s12 + kC2;
			const float B1 = mu1_sq + mu2_sq + kC1, B2 = s1 + s2 + kC2;
			// MUFU.RCP reciprocals (1 ulp): far inside the 1e-5 / 1e-3 tolerances of the loss and its gradient
			const float rB1 = __fdividef(1.0f, B1), rB2 = __fdividef(1.0f, B2);
			const float inv = rB1 * rB2;
			const float val = A1 * A2 * inv;                       // utils/loss_utils.py:77
			val_sum += val;
			l1_sum += fabsf(s_x[ly0 + j + kR][lx + kR] - s_y[ly0 + j + kR][lx + kR]);   // utils/loss_utils.py:18
			if (dmaps != nullptr) {
				const size_t at = ch * plane + (size_t)gy * W + gx;
				// partial derivatives of the map w.r.t. the three window averages img1 enters: E[x], E[x^2], E[xy]
				dmaps[at] = 2.0f * mu2 * (A2 - A1) * inv - val * 2.0f * mu1 * (rB1 - rB2);
				dmaps[plane_all + at] = -val * rB2;
				dmaps[2 * plane_all + at] = 2.0f * A1 * inv;
			}
		}
	}
	const float s_ssim = block_sum(val_sum, warp_part);
	const float s_l1 = block_sum(l1_sum, warp_part);
	if (tid == 0) {
		atomicAdd(&sums->ssim, s_ssim);
		atomicAdd(&sums->l1, s_l1);
		__threadfence();
		const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
		if (atomicAdd(&sums->done, 1u) == total - 1) {
			// last block: every other block's sums are visible (fence before its ticket)
			const float L1 = atomicAdd(&sums->l1, 0.0f) * inv_numel;
			const float S = atomicAdd(&sums->ssim, 0.0f) * inv_numel;
			out[0] = (1.0f - lambda) * L1 + lambda * (1.0f - S);   // train_mesh_gaussian.py:94 without mrloss
			out[1] = L1;
			out[2] = S;
		}
	}
}

// dL/dimg = (1 - lambda) sign(x - y) / numel  -  lambda / numel * (conv(D_mu) + 2 x conv(D_xx) + y conv(D_xy))
__global__ void __launch_bounds__(kPT)
photometric_backward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                            const float* __restrict__ dmaps, size_t plane_all, float l1_scale, float ssim_scale,
                            float* __restrict__ dL_dimg1)
{
	pdl_sync();
	__shared__ float s_d[3][kIn][kIn + 1];
	__shared__ float s_h[3][kIn][kT + 1];
	const int ch = blockIdx.z;
	const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
	const size_t plane = (size_t)H * W;
	const int tid = threadIdx.x;
	{
		const float* const src[3] = {dmaps + ch * plane, dmaps + plane_all + ch * plane, dmaps + 2 * plane_all + ch * plane};
		load_tile<3>(s_d, src, H, W, x0, y0, tid);
	}
	__syncthreads();
	for (int i = tid; i < kIn * (kT / kSeg); i += kPT) {
		const int ly = i / (kT / kSeg), c0 = (i % (kT / kSeg)) * kSeg;
#pragma unroll
		for (int m = 0; m < 3; m++) {
			float v[kSeg + 10];
#pragma unroll
			for (int k = 0; k < kSeg + 10; k++)
				v[k] = s_d[m][ly][c0 + k];
#pragma unroll
			for (int j = 0; j < kSeg; j += 2) {
				f2 acc = pk1(0.0f);
#pragma unroll
				for (int t = 0; t < 12; t++)
					acc = fma2(win2(t), pk1(v[j + t]), acc);
				s_h[m][ly][c0 + j] = lo(acc);
				s_h[m][ly][c0 + j + 1] = hi(acc);
			}
		}
	}
	__syncthreads();
	const int lx = tid & 31, ly0 = (tid >> 5) * kSeg;
	const int gx = x0 + lx;
	float acc[3][kSeg];
	vertical4<3>(s_h, ly0, lx, acc);
#pragma unroll
	for (int j = 0; j < kSeg; j++) {
		const int gy = y0 + ly0 + j;
		if (gx < W && gy < H) {
			const size_t at = ch * plane + (size_t)gy * W + gx;
			const float x = img1[at], y = img2[at];
			const float d = x - y;
			const float sgn = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
			dL_dimg1[at] = l1_scale * sgn + ssim_scale * (acc[0][j] + 2.0f * x * acc[1][j] + y * acc[2][j]);
		}
	}
}

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
mesh_restrict_kernel(int P, const float* __restrict__ scale, const float* __restrict__ v1, const float* __restrict__ v2,
                     const float* __restrict__ v3, float weight, float* __restrict__ loss, float* __restrict__ dL_dscale,
                     int accumulate)
{
	pdl_sync();
	__shared__ float warp_part[kThreads / 32];
	float part = 0.0f;
	for (int i = blockIdx.x * kThreads + threadIdx.x; i < P; i += gridDim.x * kThreads) {
		const float s0 = scale[3 * i], s1 = scale[3 * i + 1], s2 = scale[3 * i + 2];
		const float abx = v2[3 * i] - v1[3 * i], aby = v2[3 * i + 1] - v1[3 * i + 1], abz = v2[3 * i + 2] - v1[3 * i + 2];
		const float acx = v3[3 * i] - v1[3 * i], acy = v3[3 * i + 1] - v1[3 * i + 1], acz = v3[3 * i + 2] - v1[3 * i + 2];
		const float cx = aby * acz - abz * acy, cy = abz * acx - abx * acz, cz = abx * acy - aby * acx;
		const float r = sqrtf(sqrtf(cx * cx + cy * cy + cz * cz));        // "circumradius" of loss_utils.py:87-94
		const float mx = fmaxf(s0, fmaxf(s1, s2));
		const float l = mx - weight * r;
		const bool on = l > 0.0f;
		part += on ? l : 0.0f;
		if (dL_dscale != nullptr) {
			// gradient of max goes to the first maximal component
			const int k = (s0 >= s1 && s0 >= s2) ? 0 : (s1 >= s2 ? 1 : 2);
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float g = (on && k == c) ? 1.0f : 0.0f;
				if (!accumulate)
					dL_dscale[3 * i + c] = g;
				else if (g != 0.0f)
					dL_dscale[3 * i + c] += g;
			}
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		part += __shfl_xor_sync(0xffffffffu, part, o);
	if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
	__syncthreads();
	if (threadIdx.x == 0) {
		float s = 0.0f;
		for (int w = 0; w < kThreads / 32; w++) s += warp_part[w];
		atomicAdd(loss, s);
	}
}

} // namespace

size_t photometric_scratch_bytes(int C, int H, int W)
{
	if (C <= 0 || H <= 0 || W <= 0) return kMapsOffset;
	return kMapsOffset + 3 * (size_t)C * H * W * sizeof(float);
}

int launch_photometric(int C, int H, int W, const float* img, const float* gt, float lambda, char* scratch, float* out,
                       float* dL_dimg, cudaStream_t stream)
{
	PhotoSums* sums = reinterpret_cast<PhotoSums*>(scratch);
	float* dmaps = reinterpret_cast<float*>(scratch + kMapsOffset);
	cudaMemsetAsync(sums, 0, sizeof(PhotoSums), stream);
	if (C <= 0 || H <= 0 || W <= 0) {
		cudaMemsetAsync(out, 0, 3 * sizeof(float), stream);
		return GM_OK;
	}
	const size_t numel = (size_t)C * H * W;
	const float inv_numel = 1.0f / (float)numel;
	const dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, C), block(kPT);
	launch_k(photometric_forward_kernel, dim3(grid), dim3(block), 0, stream, H, W, img, gt, sums, dL_dimg ? dmaps : nullptr, numel, lambda,
	                                                       inv_numel, out);
	if (dL_dimg != nullptr)
		launch_k(photometric_backward_kernel, dim3(grid), dim3(block), 0, stream, H, W, img, gt, dmaps, numel, (1.0f - lambda) * inv_numel,
		                                                        -lambda * inv_numel, dL_dimg);
	return GM_OK;
}

int launch_mesh_restrict(int P, const float* scale, const float* v1, const float* v2, const float* v3, float weight,
                         float* loss, float* dL_dscale, int accumulate, cudaStream_t stream)
{
	cudaMemsetAsync(loss, 0, sizeof(float), stream);
	if (P <= 0) return GM_OK;
	const int blocks = min(148 * 8, (P + kThreads - 1) / kThreads);
	launch_k(mesh_restrict_kernel, dim3(blocks), dim3(kThreads), 0, stream, P, scale, v1, v2, v3, weight, loss, dL_dscale, accumulate);
	return GM_OK;
}

} // namespace gm
