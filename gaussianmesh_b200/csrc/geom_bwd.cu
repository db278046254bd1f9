// Per-Gaussian backward: one fused kernel for what the reference runs as two
// (computeCov2DCUDA, dgr/cuda_rasterizer/backward.cu:144-274, then preprocessCUDA<3> backward,
// backward.cu:346-396 with the SH and covariance helpers at :20-139 and :278-341).
//
// Fusing removes one pass over means/cov3D/radii and the global round trip of dL_dcov3D and the
// partially accumulated dL_dmean3D.  Formulas keep the reference's association order.
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv)   // auxiliary.h:107-117
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	float3 r;
	r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return r;
}

// kOverwrite: the five per-Gaussian output rows need not be pre-zeroed -- this kernel writes every element of
// them (zeros for Gaussians that were not rendered and for SH coefficients above the active degree).
// kCoop (needs kVecSH, M == 16, and kOverwrite or degree 3): the SH rows are not read / written by their owner
// thread with twelve 128-bit accesses at a 192-byte lane stride -- one request then touches 32 cache lines and the
// L1 tag stage, not HBM, sets the pace (ncu: l1tex 50 %) -- but staged per WARP through shared memory: the 32 rows of
// a warp are one contiguous 6 KB block, copied with fully coalesced cp.async while the covariance part runs, read and
// overwritten in place by their owners (row pitch 13 float4: conflict-free LDS.128 / STS.128), and stored back the
// same way.
constexpr int kRowF4 = 12;         // float4 per SH row (M == 16)
constexpr int kRowPitchF4 = 13;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

template <bool kVecSH, bool kOverwrite, bool kCoop>
__global__ void __launch_bounds__(kThreads, 5)
geometry_backward_kernel(int P,
                         const float* __restrict__ means3D,
                         const int* __restrict__ radii,
                         const float* __restrict__ shs,
                         const float* __restrict__ scales,
                         const float* __restrict__ rotations,
                         const float* __restrict__ cov3Ds,       // precomputed or geom.cov3D
                         ViewParams vp,
                         GeometryState g,
                         const float* __restrict__ dL_dmean2D,   // [P,3]
                         const float* __restrict__ dL_dconics,   // [P,4]
                         const float* __restrict__ dL_dcolor,    // [P,3]
                         float* __restrict__ dL_dmean3D,         // [P,3]
                         float* __restrict__ dL_dcov3D,          // [P,6]
                         float* __restrict__ dL_dsh,             // [P,M,3]
                         float* __restrict__ dL_dscale,          // [P,3]
                         float* __restrict__ dL_drot)            // [P,4]
{
	pdl_sync();
	__shared__ float s_view[16];
	__shared__ float s_proj[16];
	__shared__ float s_cam[3];
	if (threadIdx.x < 16) {
		s_view[threadIdx.x] = vp.view[threadIdx.x];
		s_proj[threadIdx.x] = vp.proj[threadIdx.x];
	}
	if (threadIdx.x < 3)
		s_cam[threadIdx.x] = vp.campos[threadIdx.x];
	__syncthreads();

	const int idx = blockIdx.x * kThreads + threadIdx.x;
	const bool valid = idx < P;
	const bool visible = valid && radii[idx] > 0;
	const int lane = threadIdx.x & 31;
	extern __shared__ float4 s_rows_all[];
	float4* const s_rows = s_rows_all + (threadIdx.x >> 5) * (32 * kRowPitchF4);
	const int row0 = idx - lane;                                   // first Gaussian of this warp
	const int warp_floats4 = max(0, min(32, P - row0)) * kRowF4;
	if (kCoop) {
		const float4* src = reinterpret_cast<const float4*>(shs) + (size_t)row0 * kRowF4;
#pragma unroll
		for (int i = 0; i < kRowF4; i++) {
			const int f = lane + 32 * i;
			if (f < warp_floats4) {
				const int r = f / kRowF4, c = f - r * kRowF4;
				cp_async16(&s_rows[r * kRowPitchF4 + c], src + f);
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	}
	if (!visible) {
		if (kOverwrite && valid) {
#pragma unroll
			for (int i = 0; i < 3; i++) dL_dmean3D[3 * idx + i] = 0.0f;
#pragma unroll
			for (int i = 0; i < 6; i++) dL_dcov3D[6 * idx + i] = 0.0f;
			if (shs != nullptr && !kCoop) {
				float* dsh = dL_dsh + (size_t)idx * vp.M * 3;
				if (kVecSH) {
					for (int j4 = 0; j4 < (vp.M * 3) / 4; j4++)
						reinterpret_cast<float4*>(dsh)[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
				} else {
					for (int i = 0; i < vp.M * 3; i++) dsh[i] = 0.0f;
				}
			}
			if (scales != nullptr) {
#pragma unroll
				for (int i = 0; i < 3; i++) dL_dscale[3 * idx + i] = 0.0f;
				reinterpret_cast<float4*>(dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
			}
		}
		if (!kCoop)
			return;
	}

	float3 mean = make_float3(0.f, 0.f, 0.f), dmean = make_float3(0.f, 0.f, 0.f);
	float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	if (visible) {
	mean = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);

	// ------------------------------------------------------------------ backward.cu:157-273
	float cov6[6];
#pragma unroll
	for (int i = 0; i < 6; i++)
		cov6[i] = cov3Ds[6 * idx + i];
	const float3 dL_dconic = make_float3(dL_dconics[4 * idx], dL_dconics[4 * idx + 1], dL_dconics[4 * idx + 3]);

	const float h_x = vp.focal_x, h_y = vp.focal_y;
	const Cov2DFrame f = cov2d_frame(mean, h_x, h_y, vp.tan_fovx, vp.tan_fovy, s_view);
	const float3 t = f.t;
	const float limx = 1.3f * vp.tan_fovx;
	const float limy = 1.3f * vp.tan_fovy;
	const float x_grad_mul = f.txtz < -limx || f.txtz > limx ? 0 : 1;
	const float y_grad_mul = f.tytz < -limy || f.tytz > limy ? 0 : 1;

	const Mat3 W = mat3_cols(
		s_view[0], s_view[4], s_view[8],
		s_view[1], s_view[5], s_view[9],
		s_view[2], s_view[6], s_view[10]);
	const Mat3 Vrk = vrk_from_cov6(cov6);
	const Mat3& T = f.T;
	Mat3 cov2D = mul(mul(transpose(T), transpose(Vrk)), T);

	const float a = cov2D.c[0][0] += 0.3f;
	const float b = cov2D.c[0][1];
	const float c = cov2D.c[1][1] += 0.3f;

	const float denom = a * c - b * b;
	float dL_da = 0, dL_db = 0, dL_dc = 0;
	const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

	if (denom2inv != 0) {
		dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
		dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
		dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

		dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
		dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
		dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
		dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
		dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
		dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
	} else {
#pragma unroll
		for (int i = 0; i < 6; i++)
			dcov[i] = 0;
	}
#pragma unroll
	for (int i = 0; i < 6; i++)
		dL_dcov3D[6 * idx + i] = dcov[i];

	const float dL_dT00 = 2 * (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_da +
		(T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_db;
	const float dL_dT01 = 2 * (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_da +
		(T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_db;
	const float dL_dT02 = 2 * (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_da +
		(T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_db;
	const float dL_dT10 = 2 * (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_dc +
		(T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_db;
	const float dL_dT11 = 2 * (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_dc +
		(T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_db;
	const float dL_dT12 = 2 * (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_dc +
		(T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_db;

	const float dL_dJ00 = W.c[0][0] * dL_dT00 + W.c[0][1] * dL_dT01 + W.c[0][2] * dL_dT02;
	const float dL_dJ02 = W.c[2][0] * dL_dT00 + W.c[2][1] * dL_dT01 + W.c[2][2] * dL_dT02;
	const float dL_dJ11 = W.c[1][0] * dL_dT10 + W.c[1][1] * dL_dT11 + W.c[1][2] * dL_dT12;
	const float dL_dJ12 = W.c[2][0] * dL_dT10 + W.c[2][1] * dL_dT11 + W.c[2][2] * dL_dT12;

	const float tz = 1.f / t.z;
	const float tz2 = tz * tz;
	const float tz3 = tz2 * tz;

	const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
	const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
	const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;

	// backward.cu:269-273: first part of dL/dmean3D (assignment in the reference)
	dmean = transform_vec_4x3_transpose(make_float3(dL_dtx, dL_dty, dL_dtz), s_view);

	// ------------------------------------------------------------------ backward.cu:370-387
	{
		const float* proj = s_proj;
		const float3 m = mean;
		const float4 m_hom = transform_point_4x4(m, proj);
		const float m_w = 1.0f / (m_hom.w + 0.0000001f);
		const float d2x = dL_dmean2D[3 * idx + 0], d2y = dL_dmean2D[3 * idx + 1];
		const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
		const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
		float3 d;
		d.x = (proj[0] * m_w - proj[3] * mul1) * d2x + (proj[1] * m_w - proj[3] * mul2) * d2y;
		d.y = (proj[4] * m_w - proj[7] * mul1) * d2x + (proj[5] * m_w - proj[7] * mul2) * d2y;
		d.z = (proj[8] * m_w - proj[11] * mul1) * d2x + (proj[9] * m_w - proj[11] * mul2) * d2y;
		dmean.x += d.x; dmean.y += d.y; dmean.z += d.z;
	}

	// ------------------------------------------------------------------ backward.cu:20-139
	// dL/dsh[k] = basis_k(dir) * dL/dRGB (clamp-masked, :31-34), dL/ddir = sum_k grad basis_k * (sh_k . dL/dRGB),
	// then through the normalisation of dir (dnormvdv, :128-138).  The row of SH coefficients is read and the
	// row of gradients written with 128-bit accesses.
	}   // visible

	if (shs != nullptr) {
		if (kCoop) {
			asm volatile("cp.async.wait_all;" ::: "memory");
			__syncwarp();
		}
		float4* const my_row = s_rows + lane * kRowPitchF4;
		if (visible) {
		const int deg = vp.D;
		const float3 dir_orig = make_float3(mean.x - s_cam[0], mean.y - s_cam[1], mean.z - s_cam[2]);
		const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
		const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
		const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;

		const int n_floats = 3 * (deg + 1) * (deg + 1);
		float sh[48];
		if (kCoop) {
#pragma unroll
			for (int j = 0; j < kRowF4; j++) {
				const float4 t = my_row[j];
				sh[4 * j + 0] = t.x; sh[4 * j + 1] = t.y; sh[4 * j + 2] = t.z; sh[4 * j + 3] = t.w;
			}
		} else {
			load_sh<kVecSH>(shs + (size_t)idx * vp.M * 3, n_floats, sh);
		}
		const uint32_t clamp_bits = __float_as_uint(g.rgb_clamp[idx].w);
		float dRGB[3];
#pragma unroll
		for (int ch = 0; ch < 3; ch++)
			dRGB[ch] = dL_dcolor[3 * idx + ch] * (((clamp_bits >> ch) & 1u) ? 0 : 1);

		// basis values and their gradients w.r.t. the unit direction (zero above `deg`)
		float B[16], Gx[16], Gy[16], Gz[16];
#pragma unroll
		for (int k = 0; k < 16; k++) { B[k] = 0.f; Gx[k] = 0.f; Gy[k] = 0.f; Gz[k] = 0.f; }
		B[0] = GM_SH_C0;
		if (deg > 0) {
			B[1] = -GM_SH_C1 * y; Gy[1] = -GM_SH_C1;
			B[2] = GM_SH_C1 * z;  Gz[2] = GM_SH_C1;
			B[3] = -GM_SH_C1 * x; Gx[3] = -GM_SH_C1;
		}
		if (deg > 1) {
			B[4] = GM_SH_C2_0 * xy; Gx[4] = GM_SH_C2_0 * y; Gy[4] = GM_SH_C2_0 * x;
			B[5] = GM_SH_C2_1 * yz; Gy[5] = GM_SH_C2_1 * z; Gz[5] = GM_SH_C2_1 * y;
			B[6] = GM_SH_C2_2 * (2.f * zz - xx - yy);
			Gx[6] = GM_SH_C2_2 * 2.f * -x; Gy[6] = GM_SH_C2_2 * 2.f * -y; Gz[6] = GM_SH_C2_2 * 2.f * 2.f * z;
			B[7] = GM_SH_C2_3 * xz; Gx[7] = GM_SH_C2_3 * z; Gz[7] = GM_SH_C2_3 * x;
			B[8] = GM_SH_C2_4 * (xx - yy); Gx[8] = GM_SH_C2_4 * 2.f * x; Gy[8] = GM_SH_C2_4 * 2.f * -y;
		}
		if (deg > 2) {
			B[9] = GM_SH_C3_0 * y * (3.f * xx - yy);
			Gx[9] = GM_SH_C3_0 * 3.f * 2.f * xy; Gy[9] = GM_SH_C3_0 * 3.f * (xx - yy);
			B[10] = GM_SH_C3_1 * xy * z;
			Gx[10] = GM_SH_C3_1 * yz; Gy[10] = GM_SH_C3_1 * xz; Gz[10] = GM_SH_C3_1 * xy;
			B[11] = GM_SH_C3_2 * y * (4.f * zz - xx - yy);
			Gx[11] = GM_SH_C3_2 * -2.f * xy; Gy[11] = GM_SH_C3_2 * (-3.f * yy + 4.f * zz - xx); Gz[11] = GM_SH_C3_2 * 4.f * 2.f * yz;
			B[12] = GM_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
			Gx[12] = GM_SH_C3_3 * -3.f * 2.f * xz; Gy[12] = GM_SH_C3_3 * -3.f * 2.f * yz; Gz[12] = GM_SH_C3_3 * 3.f * (2.f * zz - xx - yy);
			B[13] = GM_SH_C3_4 * x * (4.f * zz - xx - yy);
			Gx[13] = GM_SH_C3_4 * (-3.f * xx + 4.f * zz - yy); Gy[13] = GM_SH_C3_4 * -2.f * xy; Gz[13] = GM_SH_C3_4 * 4.f * 2.f * xz;
			B[14] = GM_SH_C3_5 * z * (xx - yy);
			Gx[14] = GM_SH_C3_5 * 2.f * xz; Gy[14] = GM_SH_C3_5 * -2.f * yz; Gz[14] = GM_SH_C3_5 * (xx - yy);
			B[15] = GM_SH_C3_6 * x * (xx - 3.f * yy);
			Gx[15] = GM_SH_C3_6 * 3.f * (xx - yy); Gy[15] = GM_SH_C3_6 * -3.f * 2.f * xy;
		}

		float3 dL_ddir = make_float3(0.f, 0.f, 0.f);
		float out[48];
#pragma unroll
		for (int k = 0; k < 16; k++) {
			const float w = sh[3 * k] * dRGB[0] + sh[3 * k + 1] * dRGB[1] + sh[3 * k + 2] * dRGB[2];
			dL_ddir.x += Gx[k] * w; dL_ddir.y += Gy[k] * w; dL_ddir.z += Gz[k] * w;
			out[3 * k] = B[k] * dRGB[0]; out[3 * k + 1] = B[k] * dRGB[1]; out[3 * k + 2] = B[k] * dRGB[2];
		}
		// only coefficients up to `deg` are written (the caller's zeros stay above it, as in the reference)
		if (kCoop) {
			// `out` is zero above the active degree (B[k] == 0 there)
#pragma unroll
			for (int j = 0; j < kRowF4; j++)
				my_row[j] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
		} else {
		float* dsh = dL_dsh + (size_t)idx * vp.M * 3;
		if (kOverwrite) {
			// `out` is zero above the active degree (B[k] == 0 there); rows longer than 16 coefficients get zeros
			if (kVecSH) {
				float4* dsh4 = reinterpret_cast<float4*>(dsh);
#pragma unroll
				for (int j4 = 0; j4 < 12; j4++)
					if (4 * j4 < vp.M * 3)
						dsh4[j4] = make_float4(out[4 * j4], out[4 * j4 + 1], out[4 * j4 + 2], out[4 * j4 + 3]);
				for (int j4 = 12; j4 < (vp.M * 3) / 4; j4++)
					dsh4[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
			} else {
#pragma unroll
				for (int i = 0; i < 48; i++)
					if (i < vp.M * 3) dsh[i] = out[i];
				for (int i = 48; i < vp.M * 3; i++) dsh[i] = 0.0f;
			}
		} else if (kVecSH) {
			float4* dsh4 = reinterpret_cast<float4*>(dsh);
#pragma unroll
			for (int j4 = 0; j4 < 12; j4++) {
				if (4 * j4 + 4 <= n_floats)
					dsh4[j4] = make_float4(out[4 * j4], out[4 * j4 + 1], out[4 * j4 + 2], out[4 * j4 + 3]);
				else {
#pragma unroll
					for (int e = 0; e < 4; e++)
						if (4 * j4 + e < n_floats) dsh[4 * j4 + e] = out[4 * j4 + e];
				}
			}
		} else {
#pragma unroll
			for (int i = 0; i < 48; i++)
				if (i < n_floats) dsh[i] = out[i];
		}
		}
		const float3 d = dnormvdv3(dir_orig, dL_ddir);
		dmean.x += d.x; dmean.y += d.y; dmean.z += d.z;
		} else if (kCoop && kOverwrite) {
#pragma unroll
			for (int j = 0; j < kRowF4; j++)
				my_row[j] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		if (kCoop) {
			// rows of Gaussians that were not rendered: zeros when overwriting, untouched otherwise
			const uint32_t row_mask = kOverwrite ? 0xffffffffu : __ballot_sync(0xffffffffu, visible);
			__syncwarp();
			float4* dst = reinterpret_cast<float4*>(dL_dsh) + (size_t)row0 * kRowF4;
#pragma unroll
			for (int i = 0; i < kRowF4; i++) {
				const int f = lane + 32 * i;
				if (f < warp_floats4) {
					const int r = f / kRowF4, c = f - r * kRowF4;
					if ((row_mask >> r) & 1u)
						dst[f] = s_rows[r * kRowPitchF4 + c];
				}
			}
		}
	}
	if (!visible)
		return;

	dL_dmean3D[3 * idx + 0] = dmean.x;
	dL_dmean3D[3 * idx + 1] = dmean.y;
	dL_dmean3D[3 * idx + 2] = dmean.z;

	// ------------------------------------------------------------------ backward.cu:278-341
	if (scales != nullptr) {
		const float4 q = reinterpret_cast<const float4*>(rotations)[idx];
		const float r = q.x, x = q.y, y = q.z, z = q.w;
		const Mat3 R = mat3_cols(
			1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
			2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
			2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
		Mat3 S = mat3_cols(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
		const float3 sv = make_float3(vp.scale_modifier * scales[3 * idx], vp.scale_modifier * scales[3 * idx + 1],
		                              vp.scale_modifier * scales[3 * idx + 2]);
		S.c[0][0] = sv.x; S.c[1][1] = sv.y; S.c[2][2] = sv.z;
		const Mat3 M = mul(S, R);

		const Mat3 dL_dSigma = mat3_cols(
			dcov[0], 0.5f * dcov[1], 0.5f * dcov[2],
			0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
			0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
		Mat3 M2;
#pragma unroll
		for (int i = 0; i < 3; i++)
#pragma unroll
			for (int j = 0; j < 3; j++)
				M2.c[i][j] = M.c[i][j] * 2.0f;
		const Mat3 dL_dM = mul(M2, dL_dSigma);
		const Mat3 Rt = transpose(R);
		Mat3 dL_dMt = transpose(dL_dM);

		dL_dscale[3 * idx + 0] = Rt.c[0][0] * dL_dMt.c[0][0] + Rt.c[0][1] * dL_dMt.c[0][1] + Rt.c[0][2] * dL_dMt.c[0][2];
		dL_dscale[3 * idx + 1] = Rt.c[1][0] * dL_dMt.c[1][0] + Rt.c[1][1] * dL_dMt.c[1][1] + Rt.c[1][2] * dL_dMt.c[1][2];
		dL_dscale[3 * idx + 2] = Rt.c[2][0] * dL_dMt.c[2][0] + Rt.c[2][1] * dL_dMt.c[2][1] + Rt.c[2][2] * dL_dMt.c[2][2];

#pragma unroll
		for (int j = 0; j < 3; j++) {
			dL_dMt.c[0][j] *= sv.x;
			dL_dMt.c[1][j] *= sv.y;
			dL_dMt.c[2][j] *= sv.z;
		}

		float4 dq;
		dq.x = 2 * z * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * y * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) + 2 * x * (dL_dMt.c[1][2] - dL_dMt.c[2][1]);
		dq.y = 2 * y * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * z * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) + 2 * r * (dL_dMt.c[1][2] - dL_dMt.c[2][1]) - 4 * x * (dL_dMt.c[2][2] + dL_dMt.c[1][1]);
		dq.z = 2 * x * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * r * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) + 2 * z * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * y * (dL_dMt.c[2][2] + dL_dMt.c[0][0]);
		dq.w = 2 * r * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * x * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) + 2 * y * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * z * (dL_dMt.c[1][1] + dL_dMt.c[0][0]);
		// backward.cu:339-340: gradient w.r.t. the quaternion as given (no normalisation Jacobian)
		reinterpret_cast<float4*>(dL_drot)[idx] = dq;
	}
}

} // namespace

int launch_geometry_backward(int P, const float* means3D, const int* radii, const float* shs, const float* scales,
                             const float* rotations, const float* cov3Ds, const ViewParams& vp,
                             const GeometryState& g, const float* dL_dmean2D, const float* dL_dconic,
                             const float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                             float* dL_dscale, float* dL_drot, bool overwrite, cudaStream_t stream)
{
	if (P <= 0)
		return GM_OK;
	const dim3 grid((P + kThreads - 1) / kThreads);
	const bool vec = sh_rows_vectorizable(shs, vp.M) && sh_rows_vectorizable(dL_dsh, vp.M) &&
	                 vp.M * 3 >= 3 * (vp.D + 1) * (vp.D + 1);
	const bool coop = vec && shs != nullptr && vp.M == 16 && (overwrite || vp.D == 3);
	constexpr size_t kCoopSmem = (size_t)(kThreads / 32) * 32 * kRowPitchF4 * sizeof(float4);
#define GM_LAUNCH_GEOM(V, O, C) do { \
		if (C) cudaFuncSetAttribute(geometry_backward_kernel<V, O, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem); \
		launch_k(geometry_backward_kernel<V, O, C>, dim3(grid), dim3(kThreads), (C) ? kCoopSmem : 0, stream,  \
			P, means3D, radii, shs, scales, rotations, cov3Ds, vp, g, dL_dmean2D, dL_dconic, dL_dcolor, \
			dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot); } while (0)
	if (coop && overwrite) GM_LAUNCH_GEOM(true, true, true);
	else if (coop) GM_LAUNCH_GEOM(true, false, true);
	else if (vec && overwrite) GM_LAUNCH_GEOM(true, true, false);
	else if (vec) GM_LAUNCH_GEOM(true, false, false);
	else if (overwrite) GM_LAUNCH_GEOM(false, true, false);
	else GM_LAUNCH_GEOM(false, false, false);
#undef GM_LAUNCH_GEOM
	return GM_OK;
}

} // namespace gm
