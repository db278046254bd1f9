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
	pdl_sync();
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
	pdl_sync();
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
	pdl_sync();
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
	pdl_sync();
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

// utils/loss_utils.py:17-18.  TT = float, or uint8_t for a target kept as the 8-bit image the reference's loader reads
// (utils/general_utils.py PILtoTorch: uint8 / 255.0, the same fp32 division here) -- a quarter of the upload.
template <typename TT>
__device__ __forceinline__ float target_value(const TT* t, size_t i);
template <>
__device__ __forceinline__ float target_value<float>(const float* t, size_t i) { return t[i]; }
template <>
__device__ __forceinline__ float target_value<uint8_t>(const uint8_t* t, size_t i) { return (float)t[i] / 255.0f; }

template <typename TT>
__global__ void __launch_bounds__(kThreads)
l1_kernel(size_t numel, const float* __restrict__ img, const TT* __restrict__ target, float inv_numel,
          float* __restrict__ loss, float* __restrict__ dL_dimg)
{
	pdl_sync();
	__shared__ float warp_part[kThreads / 32];
	float part = 0.0f;
	for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < numel; i += (size_t)gridDim.x * kThreads) {
		const float d = img[i] - target_value<TT>(target, i);
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

__global__ void __launch_bounds__(kThreads)
u8_to_float_kernel(size_t numel, const uint8_t* __restrict__ src, float* __restrict__ dst)
{
	pdl_sync();
	// four pixels per thread: one 32-bit load, one 128-bit store
	const size_t i4 = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
	if (i4 + 4 <= numel && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
		const uchar4 v = *reinterpret_cast<const uchar4*>(src + i4);
		*reinterpret_cast<float4*>(dst + i4) = make_float4((float)v.x / 255.0f, (float)v.y / 255.0f, (float)v.z / 255.0f,
		                                                   (float)v.w / 255.0f);
	} else {
		for (size_t i = i4; i < numel && i < i4 + 4; i++)
			dst[i] = (float)src[i] / 255.0f;
	}
}

// ---- the reference's "Python" pipeline variants (gaussian_renderer/__init__.py:78-94) ------------------------
// compute_cov3D_python: pc.get_covariance(scaling_modifier) = strip_symmetric(L L^T), L = R(q / |q|) diag(mod * s)
// (utils/general_utils.py:64-109).  Unlike the CUDA branch (forward.cu:127) the quaternion IS normalised here.
__device__ __forceinline__ void quat_to_rot(float r, float x, float y, float z, float (&R)[9])
{
	// utils/general_utils.py:75-94, row-major R[3 * row + col]
	R[0] = 1.0f - 2.0f * (y * y + z * z); R[1] = 2.0f * (x * y - r * z); R[2] = 2.0f * (x * z + r * y);
	R[3] = 2.0f * (x * y + r * z); R[4] = 1.0f - 2.0f * (x * x + z * z); R[5] = 2.0f * (y * z - r * x);
	R[6] = 2.0f * (x * z - r * y); R[7] = 2.0f * (y * z + r * x); R[8] = 1.0f - 2.0f * (x * x + y * y);
}

__global__ void __launch_bounds__(kThreads)
cov3d_python_forward_kernel(int P, const float* __restrict__ scales, float mod, const float* __restrict__ rotations,
                            float* __restrict__ cov6)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	const float3 s = load3(scales, i);
	const float4 q = make_float4(rotations[4 * i], rotations[4 * i + 1], rotations[4 * i + 2], rotations[4 * i + 3]);
	const float norm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
	float R[9];
	quat_to_rot(q.x / norm, q.y / norm, q.z / norm, q.w / norm, R);
	const float sc[3] = {mod * s.x, mod * s.y, mod * s.z};
	float L[9];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			L[3 * a + b] = R[3 * a + b] * sc[b];
	float* o = cov6 + 6 * (size_t)i;
	// actual_covariance = L @ L^T, then (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
	const int ra[6] = {0, 0, 0, 1, 1, 2}, rb[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
	for (int e = 0; e < 6; e++)
		o[e] = L[3 * ra[e] + 0] * L[3 * rb[e] + 0] + L[3 * ra[e] + 1] * L[3 * rb[e] + 1] + L[3 * ra[e] + 2] * L[3 * rb[e] + 2];
}

__global__ void __launch_bounds__(kThreads)
cov3d_python_backward_kernel(int P, const float* __restrict__ scales, float mod, const float* __restrict__ rotations,
                             const float* __restrict__ dL_dcov6, float* __restrict__ dL_dscale, float* __restrict__ dL_drot)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	const float3 s = load3(scales, i);
	const float4 qr = make_float4(rotations[4 * i], rotations[4 * i + 1], rotations[4 * i + 2], rotations[4 * i + 3]);
	const float norm = sqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w);
	const float r = qr.x / norm, x = qr.y / norm, y = qr.z / norm, z = qr.w / norm;
	float R[9];
	quat_to_rot(r, x, y, z, R);
	const float sc[3] = {mod * s.x, mod * s.y, mod * s.z};
	const float* g = dL_dcov6 + 6 * (size_t)i;
	// the six packed entries are read from the upper triangle only: dL/dSigma = G (upper), dL/dL = (G + G^T) L
	const float Gs[9] = {2.0f * g[0], g[1], g[2], g[1], 2.0f * g[3], g[4], g[2], g[4], 2.0f * g[5]};
	float dR[9], ds[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++) {
			const float dLab = Gs[3 * a + 0] * R[0 + b] * sc[b] + Gs[3 * a + 1] * R[3 + b] * sc[b] + Gs[3 * a + 2] * R[6 + b] * sc[b];
			dR[3 * a + b] = dLab * sc[b];
			ds[b] += dLab * R[3 * a + b];
		}
	dL_dscale[3 * i + 0] = mod * ds[0];
	dL_dscale[3 * i + 1] = mod * ds[1];
	dL_dscale[3 * i + 2] = mod * ds[2];
	// dR/dq of utils/general_utils.py:86-94
	const float dr = 2.0f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
	const float dx = 2.0f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.0f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.0f * x * dR[8]);
	const float dy = 2.0f * (-2.0f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.0f * y * dR[8]);
	const float dz = 2.0f * (-2.0f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.0f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
	// q = raw / |raw|
	const float dot = r * dr + x * dx + y * dy + z * dz;
	dL_drot[4 * i + 0] = (dr - r * dot) / norm;
	dL_drot[4 * i + 1] = (dx - x * dot) / norm;
	dL_drot[4 * i + 2] = (dy - y * dot) / norm;
	dL_drot[4 * i + 3] = (dz - z * dot) / norm;
}

// convert_SHs_python (gaussian_renderer/__init__.py:87-92) and its edit-time twin (edittool/__init__.py:442-448) are
// sh_rotated_kernel above; this is their backward: dL/dshs and, through the normalised direction, dL/dpos.
// The chain is the one of eval_sh (utils/sh_utils.py:57-112) written per channel.
__global__ void __launch_bounds__(kThreads)
sh_rotated_backward_kernel(int P, int D, int M, const float* __restrict__ pos, const float* __restrict__ campos,
                           const float* __restrict__ rot, const float* __restrict__ shs, const float* __restrict__ dL_drgb,
                           float* __restrict__ dL_dshs, float* __restrict__ dL_dpos)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	const float3 p = load3(pos, i);
	const float ox = p.x - campos[0], oy = p.y - campos[1], oz = p.z - campos[2];
	const float len = sqrtf(ox * ox + oy * oy + oz * oz);
	const float dx = ox / len, dy = oy / len, dz = oz / len;
	float x = dx, y = dy, z = dz;
	const float* R = rot != nullptr ? rot + 9 * (size_t)i : nullptr;
	if (R != nullptr) {
		x = R[0] * dx + R[3] * dy + R[6] * dz;
		y = R[1] * dx + R[4] * dy + R[7] * dz;
		z = R[2] * dx + R[5] * dy + R[8] * dz;
	}
	const ShDir d = sh_dir(x, y, z);
	const float* sh = shs + (size_t)i * M * 3;
	float* dsh = dL_dshs != nullptr ? dL_dshs + (size_t)i * M * 3 : nullptr;
	const float xx = d.xx, yy = d.yy, zz = d.zz, xy = d.xy, yz = d.yz, xz = d.xz;
	// basis values b_k(dir) and their gradients
	float bk[16], bx[16], by[16], bz[16];
#pragma unroll
	for (int k = 0; k < 16; k++) { bk[k] = 0.0f; bx[k] = 0.0f; by[k] = 0.0f; bz[k] = 0.0f; }
	bk[0] = GM_SH_C0;
	if (D > 0) {
		bk[1] = -GM_SH_C1 * y; by[1] = -GM_SH_C1;
		bk[2] = GM_SH_C1 * z;  bz[2] = GM_SH_C1;
		bk[3] = -GM_SH_C1 * x; bx[3] = -GM_SH_C1;
	}
	if (D > 1) {
		bk[4] = GM_SH_C2_0 * xy; bx[4] = GM_SH_C2_0 * y; by[4] = GM_SH_C2_0 * x;
		bk[5] = GM_SH_C2_1 * yz; by[5] = GM_SH_C2_1 * z; bz[5] = GM_SH_C2_1 * y;
		bk[6] = GM_SH_C2_2 * (2.0f * zz - xx - yy); bx[6] = GM_SH_C2_2 * -2.0f * x; by[6] = GM_SH_C2_2 * -2.0f * y; bz[6] = GM_SH_C2_2 * 4.0f * z;
		bk[7] = GM_SH_C2_3 * xz; bx[7] = GM_SH_C2_3 * z; bz[7] = GM_SH_C2_3 * x;
		bk[8] = GM_SH_C2_4 * (xx - yy); bx[8] = GM_SH_C2_4 * 2.0f * x; by[8] = GM_SH_C2_4 * -2.0f * y;
	}
	if (D > 2) {
		bk[9] = GM_SH_C3_0 * y * (3.0f * xx - yy); bx[9] = GM_SH_C3_0 * 6.0f * xy; by[9] = GM_SH_C3_0 * 3.0f * (xx - yy);
		bk[10] = GM_SH_C3_1 * xy * z; bx[10] = GM_SH_C3_1 * yz; by[10] = GM_SH_C3_1 * xz; bz[10] = GM_SH_C3_1 * xy;
		bk[11] = GM_SH_C3_2 * y * (4.0f * zz - xx - yy); bx[11] = GM_SH_C3_2 * -2.0f * xy;
		by[11] = GM_SH_C3_2 * (4.0f * zz - xx - 3.0f * yy); bz[11] = GM_SH_C3_2 * 8.0f * yz;
		bk[12] = GM_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy); bx[12] = GM_SH_C3_3 * -6.0f * xz;
		by[12] = GM_SH_C3_3 * -6.0f * yz; bz[12] = GM_SH_C3_3 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
		bk[13] = GM_SH_C3_4 * x * (4.0f * zz - xx - yy); bx[13] = GM_SH_C3_4 * (4.0f * zz - 3.0f * xx - yy);
		by[13] = GM_SH_C3_4 * -2.0f * xy; bz[13] = GM_SH_C3_4 * 8.0f * xz;
		bk[14] = GM_SH_C3_5 * z * (xx - yy); bx[14] = GM_SH_C3_5 * 2.0f * xz; by[14] = GM_SH_C3_5 * -2.0f * yz; bz[14] = GM_SH_C3_5 * (xx - yy);
		bk[15] = GM_SH_C3_6 * x * (xx - 3.0f * yy); bx[15] = GM_SH_C3_6 * 3.0f * (xx - yy); by[15] = GM_SH_C3_6 * -6.0f * xy;
	}
	const int n = (D + 1) * (D + 1);
	float gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll
	for (int ch = 0; ch < 3; ch++) {
		// forward value of the channel decides the clamp (max(v + 0.5, 0) has zero slope below)
		const float v = sh_channel(D, d, [sh, ch](int k) { return sh[3 * k + ch]; });
		const float gch = (v + 0.5f < 0.0f) ? 0.0f : dL_drgb[3 * i + ch];
		for (int k = 0; k < M; k++) {
			const float c = (k < n) ? sh[3 * k + ch] : 0.0f;
			if (dsh != nullptr)
				dsh[3 * k + ch] = (k < n) ? gch * bk[k & 15] : 0.0f;
			if (k < n) {
				gx += gch * c * bx[k & 15];
				gy += gch * c * by[k & 15];
				gz += gch * c * bz[k & 15];
			}
		}
	}
	if (dL_dpos == nullptr)
		return;
	// dir' = R^T dir: d/d dir = R (d/d dir')
	float hx = gx, hy = gy, hz = gz;
	if (R != nullptr) {
		hx = R[0] * gx + R[1] * gy + R[2] * gz;
		hy = R[3] * gx + R[4] * gy + R[5] * gz;
		hz = R[6] * gx + R[7] * gy + R[8] * gz;
	}
	// dir = o / |o|
	const float dot = dx * hx + dy * hy + dz * hz;
	dL_dpos[3 * i + 0] = (hx - dx * dot) / len;
	dL_dpos[3 * i + 1] = (hy - dy * dot) / len;
	dL_dpos[3 * i + 2] = (hz - dz * dot) / len;
}

// SingleObjectDeform.load_mesh (edittool/__init__.py:65-101, the face-id branch): the Gaussian's face -> its three
// vertex ids, and the area-ratio barycentric weights of the projected point (edittool/general_utils.py:73-88),
// evaluated in float64 as numpy does.
__global__ void __launch_bounds__(kThreads)
load_mesh_kernel(int P, int num_faces, const double* __restrict__ vertex, const int* __restrict__ faces,
                 const long long* __restrict__ face_id, const float* __restrict__ proj_pos,
                 int* __restrict__ gaussian_triangles, double* __restrict__ weights)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P)
		return;
	long long f = face_id[i];
	f = f < 0 ? 0 : (f >= num_faces ? num_faces - 1 : f);
	const int t0 = faces[3 * f], t1 = faces[3 * f + 1], t2 = faces[3 * f + 2];
	gaussian_triangles[3 * i + 0] = t0; gaussian_triangles[3 * i + 1] = t1; gaussian_triangles[3 * i + 2] = t2;
	const double gx = proj_pos[3 * i], gy = proj_pos[3 * i + 1], gz = proj_pos[3 * i + 2];
	double e[3][3];
	const int tv[3] = {t0, t1, t2};
#pragma unroll
	for (int k = 0; k < 3; k++) {
		e[k][0] = gx - vertex[3 * (size_t)tv[k]];
		e[k][1] = gy - vertex[3 * (size_t)tv[k] + 1];
		e[k][2] = gz - vertex[3 * (size_t)tv[k] + 2];
	}
	auto cross_norm = [](const double (&a)[3], const double (&b)[3]) {
		const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
		return sqrt(cx * cx + cy * cy + cz * cz);
	};
	const double s1 = cross_norm(e[1], e[2]), s2 = cross_norm(e[0], e[2]), s3 = cross_norm(e[0], e[1]);
	const double sum = s1 + s2 + s3;
	weights[3 * i + 0] = s1 / sum; weights[3 * i + 1] = s2 / sum; weights[3 * i + 2] = s3 / sum;
}

} // namespace

int launch_mesh_bind_forward(int P, const float* bc_logits, const float* distance, const float* v1, const float* v2,
                             const float* v3, const float* normal, const float* r, float alpha_distance,
                             const float* log_scale, const float* rot_raw, const float* opacity_logit, float* xyz,
                             float* scale, float* rot, float* opacity, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(mesh_bind_forward_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, 
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
	launch_k(mesh_bind_backward_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, 
		P, bc_logits, distance, v1, v2, v3, normal, r, alpha_distance, log_scale, rot_raw, opacity_logit, dL_dxyz,
		dL_dscale, dL_drot, dL_dopacity, dL_dbc, dL_ddist, dL_dlog_scale, dL_drot_raw, dL_dopacity_logit);
	return GM_OK;
}

int launch_deform(int P, const float* V, const float* Vd, const float* VR, const float* VS, const int* tri,
                  const float* w, const float* pos_in, const float* cov_in, int cov_full, float* pos_out,
                  float* cov6_out, float* rot_out, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(deform_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, V, Vd, VR, VS, tri, w, pos_in, cov_in,
	                                                                        cov_full, pos_out, cov6_out, rot_out);
	return GM_OK;
}

int launch_sh_rotated(int P, int D, int M, const float* pos, const float* campos, const float* rot, const float* shs,
                      float* rgb, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(sh_rotated_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, D, M, pos, campos, rot, shs, rgb);
	return GM_OK;
}

int launch_cov3d_python(int P, const float* scales, float mod, const float* rotations, float* cov6, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(cov3d_python_forward_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, scales, mod, rotations, cov6);
	return GM_OK;
}

int launch_cov3d_python_backward(int P, const float* scales, float mod, const float* rotations, const float* dL_dcov6,
                                 float* dL_dscale, float* dL_drot, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(cov3d_python_backward_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, scales, mod, rotations, dL_dcov6,
	                                                                                  dL_dscale, dL_drot);
	return GM_OK;
}

int launch_sh_rotated_backward(int P, int D, int M, const float* pos, const float* campos, const float* rot, const float* shs,
                               const float* dL_drgb, float* dL_dshs, float* dL_dpos, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(sh_rotated_backward_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, D, M, pos, campos, rot, shs, dL_drgb,
	                                                                                dL_dshs, dL_dpos);
	return GM_OK;
}

int launch_load_mesh(int P, int num_faces, const double* vertex, const int* faces, const long long* face_id,
                     const float* proj_pos, int* gaussian_triangles, double* weights, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(load_mesh_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, num_faces, vertex, faces, face_id, proj_pos,
	                                                                      gaussian_triangles, weights);
	return GM_OK;
}

int launch_l1(size_t numel, const float* img, const void* target, int target_is_u8, float* loss, float* dL_dimg, cudaStream_t stream)
{
	cudaMemsetAsync(loss, 0, sizeof(float), stream);
	if (numel == 0) return GM_OK;
	const int blocks = (int)min((size_t)148 * 8, (numel + kThreads - 1) / kThreads);
	if (target_is_u8)
		launch_k(l1_kernel<uint8_t>, dim3(blocks), dim3(kThreads), 0, stream, numel, img, static_cast<const uint8_t*>(target),
		         1.0f / (float)numel, loss, dL_dimg);
	else
		launch_k(l1_kernel<float>, dim3(blocks), dim3(kThreads), 0, stream, numel, img, static_cast<const float*>(target),
		         1.0f / (float)numel, loss, dL_dimg);
	return GM_OK;
}

int launch_u8_to_float(size_t numel, const uint8_t* src, float* dst, cudaStream_t stream)
{
	if (numel == 0) return GM_OK;
	const size_t threads = (numel + 3) / 4;
	launch_k(u8_to_float_kernel, dim3((unsigned int)((threads + kThreads - 1) / kThreads)), dim3(kThreads), 0, stream, numel, src, dst);
	return GM_OK;
}

} // namespace gm
