// Device-side math shared by the sm_100a kernels.
//
// Parity note (SURVEY.md 7 "hard parts"): the discrete decisions of the pipeline -- near cull,
// det == 0, ceil(3 sqrt(lambda)), the tile rectangle, alpha < 1/255, T < 1e-4 -- flip on 1-ulp
// differences.  Every formula below therefore keeps the reference's association order so that
// nvcc's default FMA contraction produces the same roundings; the file:line each one follows is
// cited at the function.  No fast-math anywhere.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "state.h"

namespace gm {

// ---- error handling ---------------------------------------------------------------------------
void set_last_error(const char* what, cudaError_t err);
int check_stage(const char* what, bool debug, cudaStream_t stream);

// ---- constants (auxiliary.h:20-38) -------------------------------------------------------------
#define GM_SH_C0 0.28209479177387814f
#define GM_SH_C1 0.4886025119029199f
#define GM_SH_C2_0 1.0925484305920792f
#define GM_SH_C2_1 -1.0925484305920792f
#define GM_SH_C2_2 0.31539156525252005f
#define GM_SH_C2_3 -1.0925484305920792f
#define GM_SH_C2_4 0.5462742152960396f
#define GM_SH_C3_0 -0.5900435899266435f
#define GM_SH_C3_1 2.890611442640554f
#define GM_SH_C3_2 -0.4570457994644658f
#define GM_SH_C3_3 0.3731763325901154f
#define GM_SH_C3_4 -0.4570457994644658f
#define GM_SH_C3_5 1.445305721320277f
#define GM_SH_C3_6 -0.5900435899266435f

// ---- 3x3 matrix with glm's conventions ---------------------------------------------------------
// c[i][j] is COLUMN i, ROW j; the product below evaluates each entry in the
// order glm 0.9.9 does (detail/type_mat3x3.inl operator*): A[0][r]*B[c][0] + A[1][r]*B[c][1] +
// A[2][r]*B[c][2].
struct Mat3 {
	float c[3][3];
};

__device__ __forceinline__ Mat3 mat3_cols(float a0, float a1, float a2, float b0, float b1, float b2,
                                          float c0, float c1, float c2)
{
	Mat3 m;
	m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
	m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
	m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
	return m;
}

__device__ __forceinline__ Mat3 mul(const Mat3& A, const Mat3& B)
{
	Mat3 R;
#pragma unroll
	for (int col = 0; col < 3; col++)
#pragma unroll
		for (int row = 0; row < 3; row++)
			R.c[col][row] = A.c[0][row] * B.c[col][0] + A.c[1][row] * B.c[col][1] + A.c[2][row] * B.c[col][2];
	return R;
}

__device__ __forceinline__ Mat3 transpose(const Mat3& A)
{
	Mat3 R;
#pragma unroll
	for (int col = 0; col < 3; col++)
#pragma unroll
		for (int row = 0; row < 3; row++)
			R.c[col][row] = A.c[row][col];
	return R;
}

// ---- point transforms (auxiliary.h:57-96) ------------------------------------------------------
__device__ __forceinline__ float3 transform_point_4x3(const float3& p, const float* m)
{
	return make_float3(
		m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
		m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
		m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}

__device__ __forceinline__ float4 transform_point_4x4(const float3& p, const float* m)
{
	return make_float4(
		m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
		m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
		m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
		m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}

__device__ __forceinline__ float3 transform_vec_4x3_transpose(const float3& p, const float* m)
{
	return make_float3(
		m[0] * p.x + m[1] * p.y + m[2] * p.z,
		m[4] * p.x + m[5] * p.y + m[6] * p.z,
		m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

// auxiliary.h:40-43 -- the literals are DOUBLE in the reference, so this is evaluated in fp64 and
// rounded once.
__device__ __forceinline__ float ndc_to_pix(float v, int S)
{
	return ((v + 1.0) * S - 1.0) * 0.5;
}

// auxiliary.h:45-55
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, int tiles_x, int tiles_y,
                                          int& x0, int& y0, int& x1, int& y1)
{
	x0 = min(tiles_x, max(0, (int)((p.x - max_radius) / kTile)));
	y0 = min(tiles_y, max(0, (int)((p.y - max_radius) / kTile)));
	x1 = min(tiles_x, max(0, (int)((p.x + max_radius + kTile - 1) / kTile)));
	y1 = min(tiles_y, max(0, (int)((p.y + max_radius + kTile - 1) / kTile)));
}

// forward.cu:118-152.  The quaternion is used as given (NOT normalised, forward.cu:127).
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot, float* cov3D)
{
	Mat3 S = mat3_cols(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
	S.c[0][0] = mod * scale.x;
	S.c[1][1] = mod * scale.y;
	S.c[2][2] = mod * scale.z;

	float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
	Mat3 R = mat3_cols(
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));

	Mat3 M = mul(S, R);
	Mat3 Sigma = mul(transpose(M), M);

	cov3D[0] = Sigma.c[0][0];
	cov3D[1] = Sigma.c[0][1];
	cov3D[2] = Sigma.c[0][2];
	cov3D[3] = Sigma.c[1][1];
	cov3D[4] = Sigma.c[1][2];
	cov3D[5] = Sigma.c[2][2];
}

// The pieces of forward.cu:74-113 that both the forward and the backward (backward.cu:157-190)
// need: clamped view-space point, T = W * J.
struct Cov2DFrame {
	float3 t;            // view-space mean after the 1.3*tanfov clamp
	float txtz, tytz;    // unclamped ratios (backward needs them for the gradient masks)
	Mat3 T;
};

__device__ __forceinline__ Cov2DFrame cov2d_frame(const float3& mean, float focal_x, float focal_y,
                                                  float tan_fovx, float tan_fovy, const float* view)
{
	Cov2DFrame f;
	float3 t = transform_point_4x3(mean, view);
	const float limx = 1.3f * tan_fovx;
	const float limy = 1.3f * tan_fovy;
	f.txtz = t.x / t.z;
	f.tytz = t.y / t.z;
	t.x = min(limx, max(-limx, f.txtz)) * t.z;
	t.y = min(limy, max(-limy, f.tytz)) * t.z;
	f.t = t;

	Mat3 J = mat3_cols(
		focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
		0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
		0.0f, 0.0f, 0.0f);
	Mat3 W = mat3_cols(
		view[0], view[4], view[8],
		view[1], view[5], view[9],
		view[2], view[6], view[10]);
	f.T = mul(W, J);
	return f;
}

__device__ __forceinline__ Mat3 vrk_from_cov6(const float* cov3D)
{
	return mat3_cols(
		cov3D[0], cov3D[1], cov3D[2],
		cov3D[1], cov3D[3], cov3D[4],
		cov3D[2], cov3D[4], cov3D[5]);
}

// forward.cu:74-113: returns (cov[0][0] + 0.3, cov[0][1], cov[1][1] + 0.3)
__device__ __forceinline__ float3 cov2d(const float3& mean, float focal_x, float focal_y, float tan_fovx,
                                        float tan_fovy, const float* cov3D, const float* view)
{
	Cov2DFrame f = cov2d_frame(mean, focal_x, focal_y, tan_fovx, tan_fovy, view);
	Mat3 Vrk = vrk_from_cov6(cov3D);
	Mat3 cov = mul(mul(transpose(f.T), transpose(Vrk)), f.T);
	cov.c[0][0] += 0.3f;
	cov.c[1][1] += 0.3f;
	return make_float3(cov.c[0][0], cov.c[0][1], cov.c[1][1]);
}

// Degree 0..3 real SH basis dotted with one colour channel; forward.cu:20-62 with the glm::vec3
// arithmetic written out per channel (sh points at coefficient 0 of that channel, stride 3).
// (x, y, z) is the unit view direction.
struct ShDir {
	float x, y, z, xx, yy, zz, xy, yz, xz;
};

__device__ __forceinline__ ShDir sh_dir(float x, float y, float z)
{
	ShDir d;
	d.x = x; d.y = y; d.z = z;
	d.xx = x * x; d.yy = y * y; d.zz = z * z;
	d.xy = x * y; d.yz = y * z; d.xz = x * z;
	return d;
}

template <typename ShFetch>
__device__ __forceinline__ float sh_channel(int deg, const ShDir& d, ShFetch sh)
{
	float result = GM_SH_C0 * sh(0);
	if (deg > 0) {
		result = result - GM_SH_C1 * d.y * sh(1) + GM_SH_C1 * d.z * sh(2) - GM_SH_C1 * d.x * sh(3);
		if (deg > 1) {
			result = result +
				GM_SH_C2_0 * d.xy * sh(4) +
				GM_SH_C2_1 * d.yz * sh(5) +
				GM_SH_C2_2 * (2.0f * d.zz - d.xx - d.yy) * sh(6) +
				GM_SH_C2_3 * d.xz * sh(7) +
				GM_SH_C2_4 * (d.xx - d.yy) * sh(8);
			if (deg > 2) {
				result = result +
					GM_SH_C3_0 * d.y * (3.0f * d.xx - d.yy) * sh(9) +
					GM_SH_C3_1 * d.xy * d.z * sh(10) +
					GM_SH_C3_2 * d.y * (4.0f * d.zz - d.xx - d.yy) * sh(11) +
					GM_SH_C3_3 * d.z * (2.0f * d.zz - 3.0f * d.xx - 3.0f * d.yy) * sh(12) +
					GM_SH_C3_4 * d.x * (4.0f * d.zz - d.xx - d.yy) * sh(13) +
					GM_SH_C3_5 * d.z * (d.xx - d.yy) * sh(14) +
					GM_SH_C3_6 * d.x * (d.xx - 3.0f * d.yy) * sh(15);
			}
		}
	}
	return result;
}

// First n_floats (= 3 (deg+1)^2) SH floats of one Gaussian into registers.  kVec: 128-bit loads; needs the
// Gaussian's row 16-byte aligned and M*3 a multiple of 4, so a partially needed float4 still lies inside the row.
template <bool kVec>
__device__ __forceinline__ void load_sh(const float* __restrict__ sh, int n_floats, float (&c)[48])
{
	if (kVec) {
		const float4* v = reinterpret_cast<const float4*>(sh);
#pragma unroll
		for (int j = 0; j < 12; j++) {
			float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
			if (4 * j < n_floats)
				t = __ldg(v + j);
			c[4 * j + 0] = t.x; c[4 * j + 1] = t.y; c[4 * j + 2] = t.z; c[4 * j + 3] = t.w;
		}
	} else {
#pragma unroll
		for (int i = 0; i < 48; i++)
			c[i] = (i < n_floats) ? __ldg(sh + i) : 0.0f;
	}
}

__host__ __device__ __forceinline__ bool sh_rows_vectorizable(const void* shs, int M)
{
	return shs != nullptr && (reinterpret_cast<uintptr_t>(shs) & 15u) == 0 && M > 0 && ((M * 3) & 3) == 0;
}

// ---- exact (output-preserving) rectangle culling -------------------------------------------------
// A splat contributes to a pixel only if power <= 0 and opacity*exp(power) >= 1/255
// (forward.cu:336-345, backward.cu:494-501).  For an axis-aligned pixel rectangle
// [px0,px1]x[py0,py1] (inclusive pixel centres) this returns true when NO pixel of the rectangle
// can pass those tests, so the (splat, rectangle) pair may be dropped without changing any output
// bit.  q(d) = a dx^2 + 2 b dx dy + c dy^2 is minimised over the continuous rectangle (a lower
// bound for every pixel), an fp32 error bound proportional to the largest term magnitude inside the
// rectangle is subtracted, and the comparison keeps a 1e-3 margin in the exponent.
//
// thr = log(255 * opacity); a splat with opacity < 1/255 (thr < 0) can never pass the alpha test.
__device__ __forceinline__ bool rect_cannot_contribute(float mx, float my, float a, float b, float c, float thr,
                                                       float px0, float py0, float px1, float py1)
{
	// Written without branches (every candidate is evaluated, the answer is selected): sort_pack runs this test for the
	// eight warp blocks of a tile back to back, and straight-line code lets the eight evaluations overlap.  The values
	// are those of the obvious early-out formulation; a division by a non-positive conic entry only feeds a discarded lane.
	const bool dead = thr < 0.0f;
	// offsets of the rectangle edges relative to the centre (d = mean - pixel, sign is irrelevant
	// for the quadratic form as long as both axes use the same convention)
	const float dx0 = px0 - mx, dx1 = px1 - mx;
	const float dy0 = py0 - my, dy1 = py1 - my;
	const bool in_x = (dx0 <= 0.0f) && (dx1 >= 0.0f);
	const bool in_y = (dy0 <= 0.0f) && (dy1 >= 0.0f);
	const bool covers = in_x && in_y;                       // the centre lies inside the rectangle
	const bool degenerate = !(a > 0.0f) || !(c > 0.0f);     // degenerate conic: leave it to the per-pixel test

	// The edge minimiser only has to be near the true one: q is stationary there, so the few-ulp error of the
	// approximate division moves q by a second-order amount, far below the margins kept at the end.
	float qmin = 3.0e38f;
	{
		const float dx = (dx0 > 0.0f) ? dx0 : dx1;                  // facing vertical edge
		const float dy = fminf(dy1, fmaxf(dy0, __fdividef(-b * dx, c)));
		const float q = a * dx * dx + 2.0f * b * dx * dy + c * dy * dy;
		qmin = in_x ? qmin : fminf(qmin, q);
	}
	{
		const float dy = (dy0 > 0.0f) ? dy0 : dy1;                  // facing horizontal edge
		const float dx = fminf(dx1, fmaxf(dx0, __fdividef(-b * dy, a)));
		const float q = a * dx * dx + 2.0f * b * dx * dy + c * dy * dy;
		qmin = in_y ? qmin : fminf(qmin, q);
	}
	const float ex = fmaxf(fabsf(dx0), fabsf(dx1));
	const float ey = fmaxf(fabsf(dy0), fabsf(dy1));
	const float mag = a * ex * ex + 2.0f * fabsf(b) * ex * ey + c * ey * ey;
	// 32 ulp of the largest term covers the rounding of q here and of `power` in the blend kernels
	const bool beyond = qmin - 4.0e-6f * mag > 2.0f * thr + 2.0e-3f;
	return dead || (!covers && !degenerate && beyond);
}

__device__ __forceinline__ float cull_threshold(float opacity)
{
	// log(255*o) rounded up a little; negative (=> never contributes) only if o*1 < 1/255 exactly as
	// the blend kernels would evaluate it (alpha = o * exp(power) <= o for power <= 0).
	// __logf (MUFU.LG2) is within ~1e-6 here (255*o in [1, 255]); 1e-4 on top keeps the threshold an upper bound.
	return (opacity < 1.0f / 255.0f) ? -1.0f : fmaxf(0.0f, __logf(255.0f * opacity)) + 1.0e-4f;
}

// ---- depth buckets (state.h) --------------------------------------------------------------------
// Monotone in z for z > 0.2 (IEEE bit patterns of positive floats order like the values).
__device__ __forceinline__ uint32_t depth_fine_bin(float z)
{
	const uint32_t bits = __float_as_uint(z);
	const uint32_t d = (bits > kDepthBinBase) ? (bits - kDepthBinBase) >> kDepthBinShift : 0u;
	return min(d, (uint32_t)(kDepthBins - 1));
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// Every kernel of the per-frame chain is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts
// with pdl_sync(): griddepcontrol.wait blocks until the previous kernel of the stream has completed and its writes are
// visible (so nothing here relaxes the stream order of memory), griddepcontrol.launch_dependents then lets the NEXT
// kernel's blocks become resident as soon as every block of this one has got past that point -- they fill the SM slots
// the last wave leaves idle and sit at their own wait, so the launch latency and the block-scheduling ramp of each
// kernel overlap the tail of its predecessor instead of following it.  GM_PDL=0 turns the attribute off (A/B).
__device__ __forceinline__ void pdl_sync()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl_enabled() ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- small PTX wrappers: mbarrier + 1-D bulk copy (TMA, SASS UBLKCP) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// orders this thread's earlier generic-proxy accesses (and those it has synchronised with) before later async-proxy
// operations (bulk copies) on the same shared memory
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t"
		"}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global -> shared::cta bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

} // namespace gm
