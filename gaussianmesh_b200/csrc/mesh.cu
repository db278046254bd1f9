// Mesh-bound Gaussian kernels: the per-face bind (training), the per-face local-frame deform
// (editing), the rotated-direction SH colour of the edit renderer, and the L1 loss of config 4.
//
// Each kernel fuses what the reference runs as a chain of Jittor elementwise/gather/bmm ops:
//   bind      scene/mesh_based_gaussian_model.py:34-43,122-152
//   deform    edittool/__init__.py:103-131 (+ strip_symmetric, edittool/general_utils.py:26-37)
//   colour    edittool/__init__.py:442-448 + edittool/sh_utils.py:34-89
//   l1        utils/loss_utils.py:17-18
// All are one thread per Gaussian (or per element), streaming, HBM-bound.
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float3 load3(const float* p, int i)
{
	return make_float3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
}

__device__ __forceinline__ void softmax3(const float3 l, float& b0, float& b1, float& b2)
{
	const float m = fmaxf(l.x, fmaxf(l.y, l.z));
	const float e0 = expf(l.x - m), e1 = expf(l.y - m), e2 = expf(l.z - m);
	const float s = e0 + e1 + e2;
	b0 = e0 / s; b1 = e1 / s; b2 = e2 / s;
}

__global__ void __launch_bounds__(kThreads)
mesh_bind_forward_kernel(int P, const float* __restrict__ bc_logits, const float* __restrict__ distance,
                         const float* __restrict__ v1, const float* __restrict__ v2, const float* __restrict__ v3,
                         const float* __restrict__ normal, const float* __restrict__ r, float alpha_distance,
                         const float* __restrict__ log_scale, const float* __restrict__ rot_raw,
                         const float* __restrict__ opacity_logit,
                         float* __restrict__ xyz, float* __restrict__ scale, float* __restrict__ rot,
                         float* __restrict__ opacity)
{
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	if (xyz != nullptr) {
		float b0, b1, b2;
		softmax3(load3(bc_logits, i), b0, b1, b2);
		const float3 a = load3(v1, i), b = load3(v2, i), c = load3(v3, i), nrm = load3(normal, i);
		// mesh_based_gaussian_model.py:143,149-150
		const float k = alpha_distance * r[i] * (sigmoidf(distance[i]) - 0.5f);
		xyz[3 * i + 0] = (b0 * a.x + b1 * b.x + b2 * c.x) + k * nrm.x;
		xyz[3 * i + 1] = (b0 * a.y + b1 * b.y + b2 * c.y) + k * nrm.y;
		xyz[3 * i + 2] = (b0 * a.z + b1 * b.z + b2 * c.z) + k * nrm.z;
	}
	if (scale != nullptr) {
		const float3 ls = load3(log_scale, i);
		scale[3 * i + 0] = expf(ls.x); scale[3 * i + 1] = expf(ls.y); scale[3 * i + 2] = expf(ls.z);
	}
	if (rot != nullptr) {
		const float4 q = reinterpret_cast<const float4*>(rot_raw)[i];
		// jt.normalize: x / sqrt(max(sum(x^2), eps)), eps = 1e-30
		const float nrm = sqrtf(fmaxf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w, 1e-30f));
		reinterpret_cast<float4*>(rot)[i] = make_float4(q.x / nrm, q.y / nrm, q.z / nrm, q.w / nrm);
	}
	if (opacity != nullptr)
		opacity[i] = sigmoidf(opacity_logit[i]);
}

__global__ void __launch_bounds__(kThreads)
mesh_bind_backward_kernel(int P, const float* __restrict__ bc_logits, const float* __restrict__ distance,
                          const float* __restrict__ v1, const float* __restrict__ v2, const float* __restrict__ v3,
                          const float* __restrict__ normal, const float* __restrict__ r, float alpha_distance,
                          const float* __restrict__ log_scale, const float* __restrict__ rot_raw,
                          const float* __restrict__ opacity_logit,
                          const float* __restrict__ dL_dxyz, const float* __restrict__ dL_dscale,
                          const float* __restrict__ dL_drot, const float* __restrict__ dL_dopacity,
                          float* __restrict__ dL_dbc, float* __restrict__ dL_ddist,
                          float* __restrict__ dL_dlog_scale, float* __restrict__ dL_drot_raw,
                          float* __restrict__ dL_dopacity_logit)
{
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	if (dL_dxyz != nullptr && dL_dbc != nullptr) {
		float b0, b1, b2;
		softmax3(load3(bc_logits, i), b0, b1, b2);
		const float3 gx = load3(dL_dxyz, i);
		const float3 a = load3(v1, i), b = load3(v2, i), c = load3(v3, i), nrm = load3(normal, i);
		const float g0 = gx.x * a.x + gx.y * a.y + gx.z * a.z;
		const float g1 = gx.x * b.x + gx.y * b.y + gx.z * b.z;
		const float g2 = gx.x * c.x + gx.y * c.y + gx.z * c.z;
		const float dot = b0 * g0 + b1 * g1 + b2 * g2;
		dL_dbc[3 * i + 0] = b0 * (g0 - dot);
		dL_dbc[3 * i + 1] = b1 * (g1 - dot);
		dL_dbc[3 * i + 2] = b2 * (g2 - dot);
		if (dL_ddist != nullptr) {
			const float sg = sigmoidf(distance[i]);
			dL_ddist[i] = (gx.x * nrm.x + gx.y * nrm.y + gx.z * nrm.z) * alpha_distance * r[i] * sg * (1.0f - sg);
		}
	}
	if (dL_dscale != nullptr && dL_dlog_scale != nullptr) {
		const float3 ls = load3(log_scale, i), gs = load3(dL_dscale, i);
		dL_dlog_scale[3 * i + 0] = gs.x * expf(ls.x);
		dL_dlog_scale[3 * i + 1] = gs.y * expf(ls.y);
		dL_dlog_scale[3 * i + 2] = gs.z * expf(ls.z);
	}
	if (dL_drot != nullptr && dL_drot_raw != nullptr) {
		const float4 q = reinterpret_cast<const float4*>(rot_raw)[i];
		const float4 gq = reinterpret_cast<const float4*>(dL_drot)[i];
		const float ss = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
		float4 out;
		if (ss > 1e-30f) {
			const float inv = 1.0f / sqrtf(ss);
			const float4 y = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
			const float yg = y.x * gq.x + y.y * gq.y + y.z * gq.z + y.w * gq.w;
			out = make_float4((gq.x - y.x * yg) * inv, (gq.y - y.y * yg) * inv, (gq.z - y.z * yg) * inv,
			                  (gq.w - y.w * yg) * inv);
		} else {
			const float inv = 1.0f / sqrtf(1e-30f);   // clamped branch: the norm is a constant
			out = make_float4(gq.x * inv, gq.y * inv, gq.z * inv, gq.w * inv);
		}
		reinterpret_cast<float4*>(dL_drot_raw)[i] = out;
	}
	if (dL_dopacity != nullptr && dL_dopacity_logit != nullptr) {
		const float o = sigmoidf(opacity_logit[i]);
		dL_dopacity_logit[i] = dL_dopacity[i] * o * (1.0f - o);
	}
}

// edittool/__init__.py:103-131
__global__ void __launch_bounds__(kThreads)
deform_kernel(int P, const float* __restrict__ V, const float* __restrict__ Vd, const float* __restrict__ VR,
              const float* __restrict__ VS, const int* __restrict__ tri, const float* __restrict__ w,
              const float* __restrict__ pos_in, const float* __restrict__ cov_in, int cov_full,
              float* __restrict__ pos_out, float* __restrict__ cov6_out, float* __restrict__ rot_out)
{
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	float dp[3] = {0, 0, 0};
	float Rs[9], Ss[9];
#pragma unroll
	for (int e = 0; e < 9; e++) { Rs[e] = 0.0f; Ss[e] = 0.0f; }
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const int v = tri[3 * i + k];
		const float wk = w[3 * i + k];
#pragma unroll
		for (int a = 0; a < 3; a++)
			dp[a] += wk * (Vd[3 * v + a] - V[3 * v + a]);   // :116-117
#pragma unroll
		for (int e = 0; e < 9; e++) {
			Rs[e] += wk * VR[9 * (size_t)v + e];            // :120-121
			Ss[e] += wk * VS[9 * (size_t)v + e];            // :124-125
		}
	}
	// R_g = Rs^T (:122), A = R_g S_g (:127)
	float Rg[9], A[9];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			Rg[3 * a + b] = Rs[3 * b + a];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			A[3 * a + b] = Rg[3 * a + 0] * Ss[0 + b] + Rg[3 * a + 1] * Ss[3 + b] + Rg[3 * a + 2] * Ss[6 + b];
	float C[9];
	if (cov_full) {
#pragma unroll
		for (int e = 0; e < 9; e++) C[e] = cov_in[9 * (size_t)i + e];
	} else {
		const float* c6 = cov_in + 6 * (size_t)i;
		C[0] = c6[0]; C[1] = c6[1]; C[2] = c6[2];
		C[3] = c6[1]; C[4] = c6[3]; C[5] = c6[4];
		C[6] = c6[2]; C[7] = c6[4]; C[8] = c6[5];
	}
	// Sigma' = (A C) A^T (:129)
	float AC[9];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			AC[3 * a + b] = A[3 * a + 0] * C[0 + b] + A[3 * a + 1] * C[3 + b] + A[3 * a + 2] * C[6 + b];
	float Sg[9];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			Sg[3 * a + b] = AC[3 * a + 0] * A[3 * b + 0] + AC[3 * a + 1] * A[3 * b + 1] + AC[3 * a + 2] * A[3 * b + 2];

	// strip_symmetric: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
	float* c6o = cov6_out + 6 * (size_t)i;
	c6o[0] = Sg[0]; c6o[1] = Sg[1]; c6o[2] = Sg[2]; c6o[3] = Sg[4]; c6o[4] = Sg[5]; c6o[5] = Sg[8];
#pragma unroll
	for (int a = 0; a < 3; a++)
		pos_out[3 * i + a] = pos_in[3 * i + a] + dp[a];     // :131
	if (rot_out != nullptr) {
#pragma unroll
		for (int e = 0; e < 9; e++) rot_out[9 * (size_t)i + e] = Rg[e];
	}
}

// edittool/__init__.py:442-448
__global__ void __launch_bounds__(kThreads)
sh_rotated_kernel(int P, int D, int M, const float* __restrict__ pos, const float* __restrict__ campos,
                  const float* __restrict__ rot, const float* __restrict__ shs, float* __restrict__ rgb)
{
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	const float3 p = load3(pos, i);
	float dx = p.x - campos[0], dy = p.y - campos[1], dz = p.z - campos[2];
	const float len = sqrtf(dx * dx + dy * dy + dz * dz);
	dx /= len; dy /= len; dz /= len;
	float x = dx, y = dy, z = dz;
	if (rot != nullptr) {
		const float* R = rot + 9 * (size_t)i;
		// R_g^T dir
		x = R[0] * dx + R[3] * dy + R[6] * dz;
		y = R[1] * dx + R[4] * dy + R[7] * dz;
		z = R[2] * dx + R[5] * dy + R[8] * dz;
	}
	const ShDir d = sh_dir(x, y, z);
	const float* sh = shs + (size_t)i * M * 3;
#pragma unroll
	for (int ch = 0; ch < 3; ch++) {
		const float v = sh_channel(D, d, [sh, ch](int k) { return sh[3 * k + ch]; });
		rgb[3 * i + ch] = fmaxf(v + 0.5f, 0.0f);
	}
}

// utils/loss_utils.py:17-18
__global__ void __launch_bounds__(kThreads)
l1_kernel(size_t numel, const float* __restrict__ img, const float* __restrict__ target, float inv_numel,
          float* __restrict__ loss, float* __restrict__ dL_dimg)
{
	__shared__ float warp_part[kThreads / 32];
	float part = 0.0f;
	for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < numel; i += (size_t)gridDim.x * kThreads) {
		const float d = img[i] - target[i];
		part += fabsf(d);
		if (dL_dimg != nullptr)
			dL_dimg[i] = (d > 0.0f ? inv_numel : (d < 0.0f ? -inv_numel : 0.0f));
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		part += __shfl_xor_sync(0xffffffffu, part, o);
	if ((threadIdx.x & 31) == 0)
		warp_part[threadIdx.x >> 5] = part;
	__syncthreads();
	if (threadIdx.x == 0) {
		float s = 0.0f;
		for (int w = 0; w < kThreads / 32; w++) s += warp_part[w];
		atomicAdd(loss, s * inv_numel);
	}
}

} // namespace

int launch_mesh_bind_forward(int P, const float* bc_logits, const float* distance, const float* v1, const float* v2,
                             const float* v3, const float* normal, const float* r, float alpha_distance,
                             const float* log_scale, const float* rot_raw, const float* opacity_logit, float* xyz,
                             float* scale, float* rot, float* opacity, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	mesh_bind_forward_kernel<<<(P + kThreads - 1) / kThreads, kThreads, 0, stream>>>(
		P, bc_logits, distance, v1, v2, v3, normal, r, alpha_distance, log_scale, rot_raw, opacity_logit, xyz, scale,
		rot, opacity);
	return GM_OK;
}

int launch_mesh_bind_backward(int P, const float* bc_logits, const float* distance, const float* v1, const float* v2,
                              const float* v3, const float* normal, const float* r, float alpha_distance,
                              const float* log_scale, const float* rot_raw, const float* opacity_logit,
                              const float* dL_dxyz, const float* dL_dscale, const float* dL_drot,
                              const float* dL_dopacity, float* dL_dbc, float* dL_ddist, float* dL_dlog_scale,
                              float* dL_drot_raw, float* dL_dopacity_logit, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	mesh_bind_backward_kernel<<<(P + kThreads - 1) / kThreads, kThreads, 0, stream>>>(
		P, bc_logits, distance, v1, v2, v3, normal, r, alpha_distance, log_scale, rot_raw, opacity_logit, dL_dxyz,
		dL_dscale, dL_drot, dL_dopacity, dL_dbc, dL_ddist, dL_dlog_scale, dL_drot_raw, dL_dopacity_logit);
	return GM_OK;
}

int launch_deform(int P, const float* V, const float* Vd, const float* VR, const float* VS, const int* tri,
                  const float* w, const float* pos_in, const float* cov_in, int cov_full, float* pos_out,
                  float* cov6_out, float* rot_out, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	deform_kernel<<<(P + kThreads - 1) / kThreads, kThreads, 0, stream>>>(P, V, Vd, VR, VS, tri, w, pos_in, cov_in,
	                                                                        cov_full, pos_out, cov6_out, rot_out);
	return GM_OK;
}

int launch_sh_rotated(int P, int D, int M, const float* pos, const float* campos, const float* rot, const float* shs,
                      float* rgb, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	sh_rotated_kernel<<<(P + kThreads - 1) / kThreads, kThreads, 0, stream>>>(P, D, M, pos, campos, rot, shs, rgb);
	return GM_OK;
}

int launch_l1(size_t numel, const float* img, const float* target, float* loss, float* dL_dimg, cudaStream_t stream)
{
	cudaMemsetAsync(loss, 0, sizeof(float), stream);
	if (numel == 0) return GM_OK;
	const int blocks = (int)min((size_t)148 * 8, (numel + kThreads - 1) / kThreads);
	l1_kernel<<<blocks, kThreads, 0, stream>>>(numel, img, target, 1.0f / (float)numel, loss, dL_dimg);
	return GM_OK;
}

} // namespace gm
