// Training-loss kernels around the rasterizer op: SSIM with its gradient and the mesh-restrict
// regulariser of train_mesh_gaussian.py:92-94.
//
//   ssim                 utils/loss_utils.py:36-82: 11x11 Gaussian window (sigma 1.5), zero padding, per channel,
//                        C1 = 0.01^2, C2 = 0.03^2, mean over all elements.  The reference runs five grouped conv2d
//                        launches plus ~15 elementwise kernels (and their autograd mirrors); here one kernel produces
//                        the SSIM sum and the three partial-derivative maps, a second convolves those maps into
//                        dL/dimg1.  Separable passes through shared memory, one 16x16 output tile per block.
//   mesh_restrict_loss   utils/loss_utils.py:84-107: sum(max(0, max_k scale_k - weight * sqrt(|(v2-v1) x (v3-v1)|)))
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kT = 16;            // output tile edge
constexpr int kR = 5;             // window radius (11 taps)
constexpr int kIn = kT + 2 * kR;  // 26
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// gaussian(11, 1.5) of utils/loss_utils.py:23-25, normalised
__constant__ float c_win[11] = {0.00102838f, 0.00759876f, 0.03600077f, 0.10936069f, 0.21300554f, 0.26601172f,
                                0.21300554f, 0.10936069f, 0.03600077f, 0.00759876f, 0.00102838f};

__device__ __forceinline__ float block_sum(float v, float* warp_part)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	const int tid = threadIdx.y * kT + threadIdx.x;
	if ((tid & 31) == 0) warp_part[tid >> 5] = v;
	__syncthreads();
	float s = 0.0f;
	if (tid == 0)
		for (int w = 0; w < kT * kT / 32; w++) s += warp_part[w];
	return s;
}

__global__ void __launch_bounds__(kT * kT)
ssim_forward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                    float* __restrict__ ssim_sum, float* __restrict__ dmaps /* [3][C][H][W] or null */, size_t plane_all)
{
	__shared__ float s_x[kIn][kIn + 1], s_y[kIn][kIn + 1];
	__shared__ float s_h[5][kIn][kT + 1];
	__shared__ float warp_part[kT * kT / 32];
	const int ch = blockIdx.z;
	const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
	const size_t plane = (size_t)H * W;
	const float* a = img1 + ch * plane;
	const float* b = img2 + ch * plane;
	const int tid = threadIdx.y * kT + threadIdx.x;

	for (int i = tid; i < kIn * kIn; i += kT * kT) {
		const int ly = i / kIn, lx = i % kIn;
		const int gy = y0 + ly - kR, gx = x0 + lx - kR;
		const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
		s_x[ly][lx] = in ? a[(size_t)gy * W + gx] : 0.0f;       // zero padding (conv2d padding = 5)
		s_y[ly][lx] = in ? b[(size_t)gy * W + gx] : 0.0f;
	}
	__syncthreads();
	// horizontal pass: 26 rows x 16 columns, five quantities
	for (int i = tid; i < kIn * kT; i += kT * kT) {
		const int ly = i / kT, lx = i % kT;
		float m1 = 0, m2 = 0, xx = 0, yy = 0, xy = 0;
#pragma unroll
		for (int k = 0; k < 11; k++) {
			const float w = c_win[k], u = s_x[ly][lx + k], v = s_y[ly][lx + k];
			m1 += w * u; m2 += w * v; xx += w * u * u; yy += w * v * v; xy += w * u * v;
		}
		s_h[0][ly][lx] = m1; s_h[1][ly][lx] = m2; s_h[2][ly][lx] = xx; s_h[3][ly][lx] = yy; s_h[4][ly][lx] = xy;
	}
	__syncthreads();
	const int lx = threadIdx.x, ly = threadIdx.y;
	const int gx = x0 + lx, gy = y0 + ly;
	float val = 0.0f;
	if (gx < W && gy < H) {
		float mu1 = 0, mu2 = 0, exx = 0, eyy = 0, exy = 0;
#pragma unroll
		for (int k = 0; k < 11; k++) {
			const float w = c_win[k];
			mu1 += w * s_h[0][ly + k][lx]; mu2 += w * s_h[1][ly + k][lx];
			exx += w * s_h[2][ly + k][lx]; eyy += w * s_h[3][ly + k][lx]; exy += w * s_h[4][ly + k][lx];
		}
		const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
		const float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
		const float A1 = 2.0f * mu12 + kC1, A2 = 2.0f * s12 + kC2;
		const float B1 = mu1_sq + mu2_sq + kC1, B2 = s1 + s2 + kC2;
		const float inv = 1.0f / (B1 * B2);
		val = A1 * A2 * inv;                                   // utils/loss_utils.py:77
		if (dmaps != nullptr) {
			const size_t at = ch * plane + (size_t)gy * W + gx;
			// partial derivatives of the map w.r.t. the three window averages img1 enters: E[x], E[x^2], E[xy]
			dmaps[at] = (2.0f * mu2 * A2 - 2.0f * mu2 * A1) * inv - val * (2.0f * mu1 / B1 - 2.0f * mu1 / B2);
			dmaps[plane_all + at] = -val / B2;
			dmaps[2 * plane_all + at] = 2.0f * A1 * inv;
		}
	}
	const float s = block_sum(val, warp_part);
	if (tid == 0)
		atomicAdd(ssim_sum, s);
}

// dL/dimg1 += scale * (conv(D_mu) + 2 x conv(D_xx) + y conv(D_xy))
__global__ void __launch_bounds__(kT * kT)
ssim_backward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                     const float* __restrict__ dmaps, size_t plane_all, const float* __restrict__ scale_dev, float scale_host,
                     float* __restrict__ dL_dimg1, int accumulate)
{
	__shared__ float s_d[3][kIn][kIn + 1];
	__shared__ float s_h[3][kIn][kT + 1];
	const int ch = blockIdx.z;
	const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
	const size_t plane = (size_t)H * W;
	const int tid = threadIdx.y * kT + threadIdx.x;
	for (int i = tid; i < kIn * kIn; i += kT * kT) {
		const int ly = i / kIn, lx = i % kIn;
		const int gy = y0 + ly - kR, gx = x0 + lx - kR;
		const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
		const size_t at = ch * plane + (size_t)gy * W + gx;
#pragma unroll
		for (int m = 0; m < 3; m++)
			s_d[m][ly][lx] = in ? dmaps[m * plane_all + at] : 0.0f;
	}
	__syncthreads();
	for (int i = tid; i < kIn * kT; i += kT * kT) {
		const int ly = i / kT, lx = i % kT;
		float acc[3] = {0, 0, 0};
#pragma unroll
		for (int k = 0; k < 11; k++)
#pragma unroll
			for (int m = 0; m < 3; m++) acc[m] += c_win[k] * s_d[m][ly][lx + k];
#pragma unroll
		for (int m = 0; m < 3; m++) s_h[m][ly][lx] = acc[m];
	}
	__syncthreads();
	const int lx = threadIdx.x, ly = threadIdx.y;
	const int gx = x0 + lx, gy = y0 + ly;
	if (gx < W && gy < H) {
		float acc[3] = {0, 0, 0};
#pragma unroll
		for (int k = 0; k < 11; k++)
#pragma unroll
			for (int m = 0; m < 3; m++) acc[m] += c_win[k] * s_h[m][ly + k][lx];
		const size_t at = ch * plane + (size_t)gy * W + gx;
		const float scale = scale_host * (scale_dev ? scale_dev[0] : 1.0f);
		const float g = scale * (acc[0] + 2.0f * img1[at] * acc[1] + img2[at] * acc[2]);
		dL_dimg1[at] = accumulate ? dL_dimg1[at] + g : g;
	}
}

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
mesh_restrict_kernel(int P, const float* __restrict__ scale, const float* __restrict__ v1, const float* __restrict__ v2,
                     const float* __restrict__ v3, float weight, float* __restrict__ loss, float* __restrict__ dL_dscale)
{
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
			dL_dscale[3 * i] = (on && k == 0) ? 1.0f : 0.0f;
			dL_dscale[3 * i + 1] = (on && k == 1) ? 1.0f : 0.0f;
			dL_dscale[3 * i + 2] = (on && k == 2) ? 1.0f : 0.0f;
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

int launch_ssim_forward(int C, int H, int W, const float* img1, const float* img2, float* ssim_sum, float* dmaps,
                        cudaStream_t stream)
{
	cudaMemsetAsync(ssim_sum, 0, sizeof(float), stream);
	if (C <= 0 || H <= 0 || W <= 0) return GM_OK;
	const dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, C), block(kT, kT);
	ssim_forward_kernel<<<grid, block, 0, stream>>>(H, W, img1, img2, ssim_sum, dmaps, (size_t)C * H * W);
	return GM_OK;
}

int launch_ssim_backward(int C, int H, int W, const float* img1, const float* img2, const float* dmaps,
                         const float* scale_dev, float scale_host, float* dL_dimg1, int accumulate, cudaStream_t stream)
{
	if (C <= 0 || H <= 0 || W <= 0) return GM_OK;
	const dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, C), block(kT, kT);
	ssim_backward_kernel<<<grid, block, 0, stream>>>(H, W, img1, img2, dmaps, (size_t)C * H * W, scale_dev, scale_host,
	                                                 dL_dimg1, accumulate);
	return GM_OK;
}

int launch_mesh_restrict(int P, const float* scale, const float* v1, const float* v2, const float* v3, float weight,
                         float* loss, float* dL_dscale, cudaStream_t stream)
{
	cudaMemsetAsync(loss, 0, sizeof(float), stream);
	if (P <= 0) return GM_OK;
	const int blocks = min(148 * 8, (P + kThreads - 1) / kThreads);
	mesh_restrict_kernel<<<blocks, kThreads, 0, stream>>>(P, scale, v1, v2, v3, weight, loss, dL_dscale);
	return GM_OK;
}

} // namespace gm
