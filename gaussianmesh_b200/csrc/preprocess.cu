// Per-Gaussian forward stage: near cull, projection, 3D->2D covariance, conic, radius, tile
// rectangle, SH->RGB, and -- new in this design -- the per-tile instance COUNT with exact
// rectangle culling, so that no prefix sum over Gaussians and no global key sort is needed
// afterwards (DESIGN.md "binning").
//
// Replaces preprocessCUDA<3> (dgr/cuda_rasterizer/forward.cu:155-256), checkFrustum
// (rasterizer_impl.cu:54-66) and the tiles_touched/InclusiveSum pair (forward.cu:255,
// rasterizer_impl.cu:407).
#include <cstdio>
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 256;

// auxiliary.h:138-163 -- only the near plane is tested (the +-1.3 NDC test is commented out in the
// reference).  prefiltered == true turns a culled point into a device trap, as in the reference.
__device__ __forceinline__ bool near_cull_passes(const float3& p_orig, const float* view, bool prefiltered,
                                                 float3& p_view)
{
	p_view = transform_point_4x3(p_orig, view);
	if (p_view.z <= 0.2f) {
		if (prefiltered) {
			printf("Point is filtered although prefiltered is set. This shouldn't happen!");
			__trap();
		}
		return false;
	}
	return true;
}

__global__ void __launch_bounds__(kThreads)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view, uint8_t* __restrict__ present)
{
	pdl_sync();
	int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= P)
		return;
	float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
	float3 pv;
	present[idx] = near_cull_passes(p, view, false, pv) ? 1 : 0;
}

// ---- depth buckets: histogram of visible depths, then the monotone fine-bin -> bucket table ------------
constexpr int kHistThreads = 512;

__global__ void __launch_bounds__(kHistThreads)
depth_histogram_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                       uint32_t* __restrict__ hist)
{
	pdl_sync();
	__shared__ uint32_t s_hist[kDepthBins];
	for (int i = threadIdx.x; i < kDepthBins; i += kHistThreads)
		s_hist[i] = 0;
	const float v2 = view[2], v6 = view[6], v10 = view[10], v14 = view[14];
	__syncthreads();
	for (int idx = blockIdx.x * kHistThreads + threadIdx.x; idx < P; idx += gridDim.x * kHistThreads) {
		const float z = v2 * means3D[3 * idx] + v6 * means3D[3 * idx + 1] + v10 * means3D[3 * idx + 2] + v14;
		if (z > 0.2f)
			atomicAdd(&s_hist[depth_fine_bin(z)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < kDepthBins; i += kHistThreads) {
		const uint32_t c = s_hist[i];
		if (c != 0)
			atomicAdd(&hist[i], c);
	}
}

// One block.  lut[i] = floor(B * (number of depths in bins < i) / total): equal-count buckets, monotone in i.
constexpr int kLutThreads = 1024;
static_assert(kDepthBins % kLutThreads == 0, "bins per thread");

__global__ void __launch_bounds__(kLutThreads)
bucket_lut_kernel(const uint32_t* __restrict__ hist, uint8_t* __restrict__ lut, int bucket_log2)
{
	pdl_sync();
	constexpr int kPer = kDepthBins / kLutThreads;
	__shared__ uint32_t warp_sums[kLutThreads / 32];
	const int tid = threadIdx.x;
	uint32_t c[kPer], local = 0;
#pragma unroll
	for (int i = 0; i < kPer; i++) { c[i] = hist[tid * kPer + i]; local += c[i]; }
	uint32_t incl = local;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if ((tid & 31) >= o) incl += v;
	}
	if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
	__syncthreads();
	if (tid < 32) {
		const uint32_t w = warp_sums[tid];
		uint32_t wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
			if (tid >= o) wi += v;
		}
		warp_sums[tid] = wi;   // inclusive
	}
	__syncthreads();
	const uint32_t total = warp_sums[kLutThreads / 32 - 1];
	uint32_t below = ((tid >> 5) ? warp_sums[(tid >> 5) - 1] : 0u) + incl - local;
	const uint32_t B = 1u << bucket_log2;
#pragma unroll
	for (int i = 0; i < kPer; i++) {
		const uint32_t b = total ? (uint32_t)(((uint64_t)below << bucket_log2) / total) : 0u;
		lut[tid * kPer + i] = (uint8_t)min(b, B - 1u);
		below += c[i];
	}
}

// 4 resident blocks per SM (64 registers, a 32-byte spill): measured faster than 3 blocks at 79 registers
// kCoop (SH path, M == 16, degree 3, 16-byte aligned rows): the SH rows are not fetched by their owner thread with
// twelve 128-bit loads at a 192-byte lane stride (one request = 32 cache lines: the L1 tag stage becomes the limit)
// but staged per WARP: its 32 rows are one contiguous 6 KB block, copied into shared memory with coalesced cp.async
// issued before anything else, and read by the owners after the geometry part.  Rows are kept at their natural
// 12-float4 pitch (48 KB per block, so four blocks still fit an SM) with the column index XOR-swizzled by
// (row >> 1) & 3, which makes the owners' LDS.128 conflict-free.
constexpr int kRowF4 = 12;

__device__ __forceinline__ int row_slot(int r, int c) { return r * kRowF4 + (c ^ ((r >> 1) & 3)); }

template <bool kVecSH, bool kCoop>
__global__ void __launch_bounds__(kThreads, 4)
preprocess_kernel(int P,
                  const float* __restrict__ means3D,
                  const float* __restrict__ scales,
                  const float* __restrict__ rotations,
                  const float* __restrict__ opacities,
                  const float* __restrict__ shs,
                  const float* __restrict__ cov3D_precomp,
                  const float* __restrict__ colors_precomp,
                  ViewParams vp,
                  int* __restrict__ radii,
                  GeometryState g,
                  bool prefiltered)
{
	pdl_sync();
	__shared__ float s_view[16];
	__shared__ float s_proj[16];
	__shared__ float s_cam[3];
	__shared__ __align__(16) uint8_t s_lut[kDepthBins];
	static_assert(kDepthBins == kThreads * 16, "one 16-byte load per thread");
	reinterpret_cast<uint4*>(s_lut)[threadIdx.x] = reinterpret_cast<const uint4*>(g.depth_lut)[threadIdx.x];
	if (threadIdx.x < 16) {
		s_view[threadIdx.x] = vp.view[threadIdx.x];
		s_proj[threadIdx.x] = vp.proj[threadIdx.x];
	}
	if (threadIdx.x < 3)
		s_cam[threadIdx.x] = vp.campos[threadIdx.x];
	__syncthreads();

	const int idx = blockIdx.x * kThreads + threadIdx.x;
	bool visible = false;
	extern __shared__ float4 s_rows_all[];
	float4* const s_rows = s_rows_all + (threadIdx.x >> 5) * (32 * kRowF4);
	if (kCoop) {
		const int lane = threadIdx.x & 31;
		const int row0 = idx - lane;
		const int n4 = max(0, min(32, P - row0)) * kRowF4;
		const float4* src = reinterpret_cast<const float4*>(shs) + (size_t)row0 * kRowF4;
#pragma unroll
		for (int i = 0; i < kRowF4; i++) {
			const int f = lane + 32 * i;
			if (f < n4) {
				const int r = f / kRowF4;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
				             ::"r"(smem_u32(&s_rows[row_slot(r, f - r * kRowF4)])), "l"(src + f) : "memory");
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	}

	if (idx < P) {
		// forward.cu:186-187: a Gaussian that exits early keeps radius 0 (and touches no tile).
		int my_radius_i = 0;

		do {
			const float3 p_orig = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
			float3 p_view;
			if (!near_cull_passes(p_orig, s_view, prefiltered, p_view))
				break;

			// forward.cu:195-198
			const float4 p_hom = transform_point_4x4(p_orig, s_proj);
			const float p_w = 1.0f / (p_hom.w + 0.0000001f);
			const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);

			// forward.cu:200-211
			float cov6[6];
			if (cov3D_precomp != nullptr) {
#pragma unroll
				for (int i = 0; i < 6; i++)
					cov6[i] = cov3D_precomp[6 * idx + i];
			} else {
				const float3 sc = make_float3(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]);
				const float4 q = reinterpret_cast<const float4*>(rotations)[idx];
				cov3d_from_scale_rot(sc, vp.scale_modifier, q, cov6);
#pragma unroll
				for (int i = 0; i < 6; i++)
					g.cov3D[6 * idx + i] = cov6[i];
			}

			// forward.cu:213-223
			const float3 cov = cov2d(p_orig, vp.focal_x, vp.focal_y, vp.tan_fovx, vp.tan_fovy, cov6, s_view);
			const float det = (cov.x * cov.z - cov.y * cov.y);
			if (det == 0.0f)
				break;
			const float det_inv = 1.f / det;
			const float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);

			// forward.cu:225-237
			const float mid = 0.5f * (cov.x + cov.z);
			const float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
			const float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
			const float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
			const float2 point_image = make_float2(ndc_to_pix(p_proj.x, vp.W), ndc_to_pix(p_proj.y, vp.H));
			int x0, y0, x1, y1;
			tile_rect(point_image, (int)my_radius, vp.tiles_x, vp.tiles_y, x0, y0, x1, y1);
			if ((x1 - x0) * (y1 - y0) == 0)
				break;

			// forward.cu:239-247 (SH -> RGB with clamp mask) or precomputed colours
			float4 rgbc = make_float4(0.f, 0.f, 0.f, 0.f);
			if (kCoop) {
				// evaluated after the do-block, once the warp's rows have landed
			} else if (colors_precomp == nullptr) {
				// forward.cu:25-27: dir = (pos - campos) / length
				float dx = p_orig.x - s_cam[0], dy = p_orig.y - s_cam[1], dz = p_orig.z - s_cam[2];
				const float len = sqrtf(dx * dx + dy * dy + dz * dz);
				dx = dx / len; dy = dy / len; dz = dz / len;
				const ShDir d = sh_dir(dx, dy, dz);
				float sh[48];
				load_sh<kVecSH>(shs + (size_t)idx * vp.M * 3, 3 * (vp.D + 1) * (vp.D + 1), sh);
				float r = sh_channel(vp.D, d, [&sh](int k) { return sh[3 * k + 0]; });
				float gch = sh_channel(vp.D, d, [&sh](int k) { return sh[3 * k + 1]; });
				float b = sh_channel(vp.D, d, [&sh](int k) { return sh[3 * k + 2]; });
				r += 0.5f; gch += 0.5f; b += 0.5f;
				const uint32_t bits = (r < 0 ? 1u : 0u) | (gch < 0 ? 2u : 0u) | (b < 0 ? 4u : 0u);
				rgbc = make_float4(max(r, 0.0f), max(gch, 0.0f), max(b, 0.0f), __uint_as_float(bits));
			} else {
				rgbc = make_float4(colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2],
				                   __uint_as_float(0u));
			}

			// forward.cu:249-255
			const float opacity = opacities[idx];
			g.depths[idx] = p_view.z;
			g.means2D[idx] = point_image;
			g.conic_opacity[idx] = make_float4(conic.x, conic.y, conic.z, opacity);
			if (!kCoop)
				g.rgb_clamp[idx] = rgbc;
			my_radius_i = (int)my_radius;
			visible = true;

			// Per-(tile, depth bucket) instance count.  Tiles whose 16x16 pixel block cannot receive any
			// alpha >= 1/255 from this Gaussian are skipped (rect_cannot_contribute is output-preserving).
			const float thr = cull_threshold(opacity);
			const uint32_t bucket = s_lut[depth_fine_bin(p_view.z)];
			unsigned long long kept = 0ull;
			if (thr >= 0.0f) {
				// The exact test below can only pass inside the ellipse q(d) <= 2 thr + slack, whose axis-aligned
				// extent is sqrt((2 thr + slack) * cov_xx) by sqrt((2 thr + slack) * cov_yy).  Walking only the
				// tiles that box overlaps (padded by a pixel and 1%) skips most corner tiles of the 3-sigma square
				// without changing which tiles are kept.
				const float lvl = 2.0f * thr + 0.05f;
				const float ex = 1.01f * sqrtf(lvl * cov.x) + 1.0f, ey = 1.01f * sqrtf(lvl * cov.z) + 1.0f;
				// (a degenerate conic is never culled by the exact test, so it keeps the whole square)
				const bool tight = conic.x > 0.0f && conic.z > 0.0f && ex < 1.0e6f && ey < 1.0e6f;
				const int tx_lo = tight ? max(x0, (int)floorf((point_image.x - ex) * (1.0f / kTile))) : x0;
				const int tx_hi = tight ? min(x1, (int)floorf((point_image.x + ex) * (1.0f / kTile)) + 1) : x1;
				const int ty_lo = tight ? max(y0, (int)floorf((point_image.y - ey) * (1.0f / kTile))) : y0;
				const int ty_hi = tight ? min(y1, (int)floorf((point_image.y + ey) * (1.0f / kTile)) + 1) : y1;
				const int w = x1 - x0;
				if (w * (y1 - y0) > 64) {
					// too many tiles for a 64-bit mask (and for one thread): large_tiles_kernel walks this rectangle
					// with a whole warp, once to count and once to place, in ONE kernel binary so both passes agree
					g.large_list[atomicAdd(&g.header->num_large, 1u)] = (uint32_t)idx;
				} else if (tight && conic.x * conic.z - conic.y * conic.y > 0.0f) {
					// Row by row in closed form.  A tile [px0,px1] x [py0,py1] can receive alpha >= 1/255 only if it meets
					// the ellipse q(d) = a dx^2 + 2 b dx dy + c dy^2 <= Lq; within one tile row the ellipse cut by the row's
					// band is convex, so the tiles it meets are exactly those whose pixel span overlaps the band's
					// x-interval [xmin, xmax] -- one interval per row instead of one quadratic minimisation per tile (the
					// per-tile loop ran at the pace of the warp's largest splat).  Lq carries the margins of
					// rect_cannot_contribute (2e-3 in the exponent, 32 ulp of the largest term over the walked box); the
					// blend kernels cull exactly per 8x4 block again, so a spare tile costs time, never a bit.
					const float a = conic.x, b = conic.y, c = conic.z;
					const float det_c = a * c - b * b;
					const float bx = fmaxf(fabsf((float)(tx_lo * kTile) - point_image.x), fabsf((float)(tx_hi * kTile) - point_image.x));
					const float by = fmaxf(fabsf((float)(ty_lo * kTile) - point_image.y), fabsf((float)(ty_hi * kTile) - point_image.y));
					const float Lq = 2.0f * thr + 2.0e-3f + 4.0e-6f * (a * bx * bx + 2.0f * fabsf(b) * bx * by + c * by * by);
					const float inv_a = 1.0f / a;
					const float dx_ext = sqrtf(Lq * c / det_c);          // largest |dx| on the ellipse, reached at dy = -/+ (b/c) dx_ext
					const float dy_at = b / c * dx_ext;
					const float aL = a * Lq;
					for (int ty = ty_lo; ty < ty_hi; ty++) {
						const float y0b = (float)(ty * kTile) - point_image.y;
						const float y1b = fminf((float)(ty * kTile + (kTile - 1)), (float)(vp.H - 1)) - point_image.y;
						const float dyM = fminf(y1b, fmaxf(y0b, -dy_at));     // band point nearest to the dx-maximiser
						const float dym = fminf(y1b, fmaxf(y0b, dy_at));      // ... and to the dx-minimiser
						const float DM = aL - det_c * dyM * dyM, Dm = aL - det_c * dym * dym;
						if (DM < 0.0f && Dm < 0.0f)
							continue;                                       // the band lies outside the ellipse's dy range
						const float xmax = (-b * dyM + sqrtf(fmaxf(DM, 0.0f))) * inv_a;
						const float xmin = (-b * dym - sqrtf(fmaxf(Dm, 0.0f))) * inv_a;
						const int ta = max(tx_lo, (int)ceilf((point_image.x + xmin - (float)(kTile - 1)) * (1.0f / kTile) - 1.0e-3f));
						const int tb = min(tx_hi - 1, (int)floorf((point_image.x + xmax) * (1.0f / kTile) + 1.0e-3f));
						uint32_t* const row = g.bucket_cursor + (((size_t)ty * vp.tiles_x) << vp.bucket_log2) + bucket;
						for (int tx = ta; tx <= tb; tx++) {
							atomicAdd(&row[(size_t)tx << vp.bucket_log2], 1u);
							kept |= 1ull << ((ty - y0) * w + (tx - x0));
						}
					}
				} else {
					for (int ty = ty_lo; ty < ty_hi; ty++) {
						const float py0 = (float)(ty * kTile);
						const float py1 = fminf(py0 + (kTile - 1), (float)(vp.H - 1));
						uint32_t* const row = g.bucket_cursor + (((size_t)ty * vp.tiles_x) << vp.bucket_log2) + bucket;
						for (int tx = tx_lo; tx < tx_hi; tx++) {
							const float px0 = (float)(tx * kTile);
							const float px1 = fminf(px0 + (kTile - 1), (float)(vp.W - 1));
							if (!rect_cannot_contribute(point_image.x, point_image.y, conic.x, conic.y, conic.z, thr,
							                            px0, py0, px1, py1)) {
								atomicAdd(&row[(size_t)tx << vp.bucket_log2], 1u);
								kept |= 1ull << ((ty - y0) * w + (tx - x0));
							}
						}
					}
				}
			}
			// emit replays this mask instead of re-evaluating the culling (rects of more than 64 tiles re-evaluate)
			g.tile_mask[idx] = kept;
		} while (false);

		radii[idx] = my_radius_i;
	}

	if (kCoop) {
		asm volatile("cp.async.wait_all;" ::: "memory");
		__syncwarp();
		if (visible) {
			// forward.cu:239-247 (SH -> RGB with clamp mask); forward.cu:25-27: dir = (pos - campos) / length
			const int lane = threadIdx.x & 31;
			float dx = means3D[3 * idx] - s_cam[0], dy = means3D[3 * idx + 1] - s_cam[1], dz = means3D[3 * idx + 2] - s_cam[2];
			const float len = sqrtf(dx * dx + dy * dy + dz * dz);
			dx = dx / len; dy = dy / len; dz = dz / len;
			const ShDir d = sh_dir(dx, dy, dz);
			float sh[48];
#pragma unroll
			for (int j = 0; j < kRowF4; j++) {
				const float4 t = s_rows[row_slot(lane, j)];
				sh[4 * j + 0] = t.x; sh[4 * j + 1] = t.y; sh[4 * j + 2] = t.z; sh[4 * j + 3] = t.w;
			}
			float r = sh_channel(vp.D, d, [&sh](int k) { return sh[3 * k + 0]; });
			float gch = sh_channel(vp.D, d, [&sh](int k) { return sh[3 * k + 1]; });
			float b = sh_channel(vp.D, d, [&sh](int k) { return sh[3 * k + 2]; });
			r += 0.5f; gch += 0.5f; b += 0.5f;
			const uint32_t bits = (r < 0 ? 1u : 0u) | (gch < 0 ? 2u : 0u) | (b < 0 ? 4u : 0u);
			g.rgb_clamp[idx] = make_float4(max(r, 0.0f), max(gch, 0.0f), max(b, 0.0f), __uint_as_float(bits));
		}
	}

	const uint32_t vis_mask = __ballot_sync(0xffffffffu, visible);
	if ((threadIdx.x & 31) == 0 && vis_mask != 0)
		atomicAdd(&g.header->num_visible, (uint32_t)__popc(vis_mask));
}

} // namespace

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream)
{
	if (P <= 0)
		return GM_OK;
	launch_k(mark_visible_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, means3D, viewmatrix, present);
	return GM_OK;
}

int launch_depth_buckets(int P, const float* means3D, const ViewParams& vp, const GeometryState& g, cudaStream_t stream)
{
	// one clear for every per-frame counter (frame header, scan descriptors, depth histogram, bucket cursors): a memset
	// between two kernels would break the programmatic launch chain, so it comes first
	cudaMemsetAsync(g.header, 0, frame_clear_bytes(g, (size_t)(vp.tiles_x * vp.tiles_y) << vp.bucket_log2), stream);
	if (P > 0) {
		const int hist_blocks = min((P + kHistThreads - 1) / kHistThreads, 148 * 4);
		launch_k(depth_histogram_kernel, dim3(hist_blocks), dim3(kHistThreads), 0, stream, P, means3D, vp.view, g.depth_hist);
	}
	launch_k(bucket_lut_kernel, dim3(1), dim3(kLutThreads), 0, stream, g.depth_hist, g.depth_lut, vp.bucket_log2);
	return GM_OK;
}

int launch_preprocess(int P, const float* means3D, const float* scales, const float* rotations,
                      const float* opacities, const float* shs, const float* cov3D_precomp,
                      const float* colors_precomp, const ViewParams& vp, int* radii,
                      const GeometryState& g, bool prefiltered, cudaStream_t stream)
{
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	(void)num_tiles;      // the per-frame counters were cleared by launch_depth_buckets, the first stage of the frame
	if (P <= 0)
		return GM_OK;
	const dim3 grid((P + kThreads - 1) / kThreads);
	const bool vec = colors_precomp == nullptr && sh_rows_vectorizable(shs, vp.M) && vp.M * 3 >= 3 * (vp.D + 1) * (vp.D + 1);
	constexpr size_t kCoopSmem = (size_t)(kThreads / 32) * 32 * kRowF4 * sizeof(float4);      // 48 KB
	if (vec && vp.M == 16 && vp.D == 3) {
		cudaFuncSetAttribute(preprocess_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem);
		launch_k(preprocess_kernel<true, true>, dim3(grid), dim3(kThreads), kCoopSmem, stream, 
			P, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, vp, radii, g, prefiltered);
	} else if (vec)
		launch_k(preprocess_kernel<true, false>, dim3(grid), dim3(kThreads), 0, stream, 
			P, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, vp, radii, g, prefiltered);
	else
		launch_k(preprocess_kernel<false, false>, dim3(grid), dim3(kThreads), 0, stream, 
			P, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, vp, radii, g, prefiltered);
	launch_large_tiles(0, radii, g, nullptr, 0, vp, stream);      // count pass for rectangles of more than 64 tiles
	return GM_OK;
}

} // namespace gm
