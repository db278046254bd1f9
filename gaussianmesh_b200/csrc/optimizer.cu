// Optimizer-side kernels of one training iteration: the Adam update of every parameter tensor in
// ONE launch, and the densification statistics gathered after each backward.
//
//   adam            scene/mesh_based_gaussian_model.py:243-258 builds `jt.nn.Adam(l, lr=0.0, eps=1e-15)` over seven
//                   parameter groups with their own learning rates; train_mesh_gaussian.py:136-147 steps it once per
//                   iteration.  Jittor is not vendored in /root/reference (and pins no version); the update restated
//                   here is the one its optim.Adam.step publishes:
//                       m <- b0 m + (1 - b0) g;   v <- b1 v + (1 - b1) g g
//                       p <- p - m * (lr sqrt(1 - b1^n) / (1 - b0^n)) / (sqrt(v) + eps)
//                   (bias correction folded into the step size, eps added to the UNcorrected sqrt(v)).
//                   The reference launches ~10 elementwise kernels per group; this is one pass, 16 B read + 12 B
//                   written per element, 128-bit accesses.  A tensor may carry two learning rates with a period
//                   (features kept as one [P,16,3] tensor: the first 3 floats of every 48 are f_dc, the rest f_rest
//                   with lr / 20), so the per-frame concat of get_features (:167-170) never has to exist.
//   adam (view-parallel)  SURVEY.md 8f-4 "view-parallel DP": every rank renders its own view and holds the whole
//                   model; the gradient exchange, the update and the parameter broadcast are ONE kernel over NVLink
//                   peer memory.  Rank r owns the r-th contiguous shard of the flat parameter vector: it LOADS that
//                   shard of every rank's gradient buffer (P2P reads, summed and scaled by 1/world), applies Adam
//                   with its shard of the moments (the only copy -- optimizer state is sharded) and STORES the new
//                   parameters into every rank's parameter buffer (P2P writes).  Per rank and step (world-1)/world
//                   of the gradient bytes cross NVLink in each direction, against 2x that plus a full-size Adam pass
//                   for all-reduce followed by a replicated update.
//   densify_stats   train_mesh_gaussian.py:117-121 + scene/mesh_based_gaussian_model.py:587-589:
//                   max_radii2D = max(max_radii2D, radii), bc_gradient_accum += |dL/dmean2D.xy|, denom += 1 for the
//                   visible Gaussians -- five indexed Jittor kernels, one here.
#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;                                   // float4 per thread
constexpr size_t kChunk = (size_t)kThreads * kVecPerThread * 4;    // elements per block

using AdamTensor = gm_adam_tensor;
constexpr int kAdamMaxTensors = 8;                                 // per launch

struct AdamTable {
	AdamTensor t[kAdamMaxTensors];
	unsigned int first_block[kAdamMaxTensors + 1];
	int count;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b0, float b1, float step, float eps)
{
	m = b0 * m + (1.0f - b0) * g;
	v = b1 * v + (1.0f - b1) * g * g;
	p = p - m * step / (sqrtf(v) + eps);
}

__global__ void __launch_bounds__(kThreads)
adam_kernel(const __grid_constant__ AdamTable tab, float b0, float b1, float bias, float eps,
            const uint32_t* __restrict__ skip_flag)
{
	pdl_sync();
	// gated update: a frame whose binning chunk overflowed rendered only part of its tiles, so its gradients are
	// partial -- the update is dropped on the device, without a host round trip (FrameHeader::overflow)
	if (skip_flag != nullptr && *skip_flag != 0u)
		return;
	int k = 0;
#pragma unroll
	for (int i = 1; i < kAdamMaxTensors; i++)
		if (i < tab.count && blockIdx.x >= tab.first_block[i])
			k = i;
	const AdamTensor& t = tab.t[k];
	const size_t base = (size_t)(blockIdx.x - tab.first_block[k]) * kChunk;
	const float step_tail = t.lr * bias, step_head = t.lr_head * bias;
	const bool vec = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
	                   reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15u) == 0;
#pragma unroll
	for (int j = 0; j < kVecPerThread; j++) {
		const size_t i = base + ((size_t)j * kThreads + threadIdx.x) * 4;
		if (i >= t.numel)
			break;
		float step[4];
		unsigned int r = t.period > 0 ? (unsigned int)(i % t.period) : 0u;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			step[c] = (t.period > 0 && r < t.split) ? step_head : step_tail;
			r = (r + 1 == t.period) ? 0u : r + 1;
		}
		if (vec && i + 4 <= t.numel) {
			float4 p = *reinterpret_cast<const float4*>(t.param + i);
			const float4 g = *reinterpret_cast<const float4*>(t.grad + i);
			float4 m = *reinterpret_cast<const float4*>(t.exp_avg + i);
			float4 v = *reinterpret_cast<const float4*>(t.exp_avg_sq + i);
			adam_one(p.x, g.x, m.x, v.x, b0, b1, step[0], eps);
			adam_one(p.y, g.y, m.y, v.y, b0, b1, step[1], eps);
			adam_one(p.z, g.z, m.z, v.z, b0, b1, step[2], eps);
			adam_one(p.w, g.w, m.w, v.w, b0, b1, step[3], eps);
			*reinterpret_cast<float4*>(t.param + i) = p;
			*reinterpret_cast<float4*>(t.exp_avg + i) = m;
			*reinterpret_cast<float4*>(t.exp_avg_sq + i) = v;
		} else {
			for (int c = 0; c < 4 && i + c < t.numel; c++)
				adam_one(t.param[i + c], t.grad[i + c], t.exp_avg[i + c], t.exp_avg_sq[i + c], b0, b1, step[c], eps);
		}
	}
}

struct ShardedAdamArgs {
	const float* grads[GM_MAX_PEERS];
	float* params[GM_MAX_PEERS];
	gm_adam_segment seg[kAdamMaxTensors];
	int world, num_segments;
	size_t lo, hi;                 // this rank's shard of the flat vector, multiples of 4
	float* exp_avg;                // [hi - lo]
	float* exp_avg_sq;
};

__global__ void __launch_bounds__(kThreads)
adam_sharded_p2p_kernel(const __grid_constant__ ShardedAdamArgs a, float b0, float b1, float bias, float eps,
                        float inv_world)
{
	pdl_sync();
	const size_t i = a.lo + ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
	if (i >= a.hi)
		return;
	// which tensor does this float4 belong to?  Segments start at multiples of 32 floats, so it never straddles two.
	int k = -1;
#pragma unroll
	for (int s = 0; s < kAdamMaxTensors; s++)
		if (s < a.num_segments && i >= a.seg[s].offset && i < a.seg[s].offset + a.seg[s].numel)
			k = s;
	if (k < 0)
		return;                                                    // alignment padding between tensors
	const gm_adam_segment& sg = a.seg[k];
	const size_t rel = i - sg.offset;
	float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
	for (int r = 0; r < GM_MAX_PEERS; r++) {
		if (r < a.world) {
			const float4 t = *reinterpret_cast<const float4*>(a.grads[r] + i);    // peer memory for r != this rank
			g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
		}
	}
	float gv[4] = {g.x * inv_world, g.y * inv_world, g.z * inv_world, g.w * inv_world};
	float4 p4 = *reinterpret_cast<const float4*>(a.params[0] + i);                 // params[0] is this rank's own buffer
	float pv[4] = {p4.x, p4.y, p4.z, p4.w};
	float4 m4 = *reinterpret_cast<const float4*>(a.exp_avg + (i - a.lo));
	float4 v4 = *reinterpret_cast<const float4*>(a.exp_avg_sq + (i - a.lo));
	float mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
	unsigned int r = sg.period > 0 ? (unsigned int)(rel % sg.period) : 0u;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		const float lr = (sg.period > 0 && r < sg.split) ? sg.lr_head : sg.lr;
		r = (r + 1 == sg.period) ? 0u : r + 1;
		if (rel + c < sg.numel)
			adam_one(pv[c], gv[c], mv[c], vv[c], b0, b1, lr * bias, eps);
	}
	*reinterpret_cast<float4*>(a.exp_avg + (i - a.lo)) = make_float4(mv[0], mv[1], mv[2], mv[3]);
	*reinterpret_cast<float4*>(a.exp_avg_sq + (i - a.lo)) = make_float4(vv[0], vv[1], vv[2], vv[3]);
	const float4 out = make_float4(pv[0], pv[1], pv[2], pv[3]);
#pragma unroll
	for (int q = 0; q < GM_MAX_PEERS; q++)
		if (q < a.world)
			*reinterpret_cast<float4*>(a.params[q] + i) = out;
}

// ---- the same exchange through the NVSwitch multicast mapping (NVLS) -----------------------------------------------
// With a multicast address over every rank's gradient vector, ONE multimem.ld_reduce returns the sum of the N copies --
// the reduction happens inside the switch, so each rank pulls 1/N of the vector once instead of N-1 peer copies of it --
// and ONE multimem.st writes the updated parameters into all N parameter vectors.  Per GPU and step the NVLink bytes
// drop from 2 (N-1)/N * bytes(vector) to about 2/N * bytes(vector) in each direction.  Four 128-bit reductions are in
// flight per thread before the first is consumed.
struct ShardedAdamMcArgs {
	const float* grads_mc;         // multicast address of the flat gradient vectors
	float* params_mc;              // multicast address of the flat parameter vectors
	const float* params_local;     // this rank's own parameter vector (unicast)
	gm_adam_segment seg[kAdamMaxTensors];
	int num_segments;
	size_t lo, hi;
	float* exp_avg;
	float* exp_avg_sq;
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr)
{
	float4 v;
	asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
	             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
	return v;
}

__device__ __forceinline__ void multimem_st(float* mc_addr, float4 v)
{
	asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
	             ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kMcUnroll = 4;

__global__ void __launch_bounds__(kThreads)
adam_sharded_mc_kernel(const __grid_constant__ ShardedAdamMcArgs a, float b0, float b1, float bias, float eps,
                       float inv_world)
{
	pdl_sync();
	const size_t base = a.lo + (size_t)blockIdx.x * (kThreads * kMcUnroll * 4) + (size_t)threadIdx.x * 4;
	float4 g[kMcUnroll];
	bool live[kMcUnroll];
	int seg_of[kMcUnroll];
#pragma unroll
	for (int u = 0; u < kMcUnroll; u++) {
		const size_t i = base + (size_t)u * kThreads * 4;
		int k = -1;
#pragma unroll
		for (int s = 0; s < kAdamMaxTensors; s++)
			if (s < a.num_segments && i >= a.seg[s].offset && i < a.seg[s].offset + a.seg[s].numel)
				k = s;
		seg_of[u] = k;
		live[u] = i < a.hi && k >= 0;                              // k < 0: alignment padding between tensors
		if (live[u])
			g[u] = multimem_ld_reduce_add(a.grads_mc + i);          // sum over the ranks, reduced in the switch
	}
#pragma unroll
	for (int u = 0; u < kMcUnroll; u++) {
		if (!live[u])
			continue;
		const size_t i = base + (size_t)u * kThreads * 4;
		const gm_adam_segment& sg = a.seg[seg_of[u]];
		const size_t rel = i - sg.offset;
		float gv[4] = {g[u].x * inv_world, g[u].y * inv_world, g[u].z * inv_world, g[u].w * inv_world};
		const float4 p4 = *reinterpret_cast<const float4*>(a.params_local + i);
		float pv[4] = {p4.x, p4.y, p4.z, p4.w};
		const float4 m4 = *reinterpret_cast<const float4*>(a.exp_avg + (i - a.lo));
		const float4 v4 = *reinterpret_cast<const float4*>(a.exp_avg_sq + (i - a.lo));
		float mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
		unsigned int r = sg.period > 0 ? (unsigned int)(rel % sg.period) : 0u;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const float lr = (sg.period > 0 && r < sg.split) ? sg.lr_head : sg.lr;
			r = (r + 1 == sg.period) ? 0u : r + 1;
			if (rel + c < sg.numel)
				adam_one(pv[c], gv[c], mv[c], vv[c], b0, b1, lr * bias, eps);
		}
		*reinterpret_cast<float4*>(a.exp_avg + (i - a.lo)) = make_float4(mv[0], mv[1], mv[2], mv[3]);
		*reinterpret_cast<float4*>(a.exp_avg_sq + (i - a.lo)) = make_float4(vv[0], vv[1], vv[2], vv[3]);
		multimem_st(a.params_mc + i, make_float4(pv[0], pv[1], pv[2], pv[3]));   // into every rank's parameter vector
	}
}

__global__ void __launch_bounds__(kThreads)
densify_stats_kernel(int P, const int* __restrict__ radii, const float* __restrict__ dL_dmean2D,
                     float* __restrict__ max_radii2D, float* __restrict__ grad_accum, float* __restrict__ denom,
                     const uint32_t* __restrict__ skip_flag)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= P || (skip_flag != nullptr && *skip_flag != 0u))
		return;
	const int r = radii[i];
	if (r <= 0)                                                    // visibility_filter = radii > 0
		return;
	max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
	const float gx = dL_dmean2D[3 * i], gy = dL_dmean2D[3 * i + 1];
	grad_accum[i] += sqrtf(gx * gx + gy * gy);
	denom[i] += 1.0f;
}

} // namespace

int launch_adam(int n, const gm_adam_tensor* tensors, int step, float beta1, float beta2, float eps,
                const uint32_t* skip_flag, cudaStream_t stream)
{
	// Python-float arithmetic of the reference optimiser: double, rounded once
	const double bias = sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
	for (int i = 0; i < n;) {
		AdamTable tab;
		tab.count = 0;
		unsigned int blocks = 0;
		for (; i < n && tab.count < kAdamMaxTensors; i++) {
			if (tensors[i].numel == 0)
				continue;
			tab.t[tab.count] = tensors[i];
			tab.first_block[tab.count] = blocks;
			blocks += (unsigned int)((tensors[i].numel + kChunk - 1) / kChunk);
			tab.count++;
		}
		tab.first_block[tab.count] = blocks;
		if (blocks > 0)
			launch_k(adam_kernel, dim3(blocks), dim3(kThreads), 0, stream, tab, beta1, beta2, (float)bias, eps, skip_flag);
	}
	return GM_OK;
}

int launch_adam_sharded_p2p(int world, int rank, const float* const* grads, float* const* params, int num_segments,
                            const gm_adam_segment* segments, size_t total, float* exp_avg, float* exp_avg_sq, int step,
                            float beta1, float beta2, float eps, cudaStream_t stream)
{
	ShardedAdamArgs a;
	a.world = world;
	a.num_segments = num_segments;
	// own buffer first, then the peers in ring order: every rank starts its P2P traffic on a different link
	for (int q = 0; q < world; q++) {
		a.grads[q] = grads[(rank + q) % world];
		a.params[q] = params[(rank + q) % world];
	}
	for (int s = 0; s < num_segments; s++)
		a.seg[s] = segments[s];
	size_t lo, hi;
	gm_adam_shard_range(total, world, rank, &lo, &hi);
	a.lo = lo;
	a.hi = hi;
	a.exp_avg = exp_avg;
	a.exp_avg_sq = exp_avg_sq;
	if (hi <= lo)
		return GM_OK;
	const double bias = sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
	const size_t vecs = (hi - lo + 3) / 4;
	launch_k(adam_sharded_p2p_kernel, dim3((unsigned int)((vecs + kThreads - 1) / kThreads)), dim3(kThreads), 0, stream, 
		a, beta1, beta2, (float)bias, eps, 1.0f / (float)world);
	return GM_OK;
}

int launch_adam_sharded_mc(int world, int rank, const float* grads_mc, float* params_mc, const float* params_local,
                           int num_segments, const gm_adam_segment* segments, size_t total, float* exp_avg, float* exp_avg_sq,
                           int step, float beta1, float beta2, float eps, cudaStream_t stream)
{
	ShardedAdamMcArgs a;
	a.grads_mc = grads_mc;
	a.params_mc = params_mc;
	a.params_local = params_local;
	a.num_segments = num_segments;
	for (int s = 0; s < num_segments; s++)
		a.seg[s] = segments[s];
	size_t lo, hi;
	gm_adam_shard_range(total, world, rank, &lo, &hi);
	a.lo = lo;
	a.hi = hi;
	a.exp_avg = exp_avg;
	a.exp_avg_sq = exp_avg_sq;
	if (hi <= lo)
		return GM_OK;
	const double bias = sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
	const size_t per_block = (size_t)kThreads * kMcUnroll * 4;
	launch_k(adam_sharded_mc_kernel, dim3((unsigned int)((hi - lo + per_block - 1) / per_block)), dim3(kThreads), 0, stream,
	         a, beta1, beta2, (float)bias, eps, 1.0f / (float)world);
	return GM_OK;
}

int launch_densify_stats(int P, const int* radii, const float* dL_dmean2D, float* max_radii2D, float* grad_accum,
                         float* denom, const uint32_t* skip_flag, cudaStream_t stream)
{
	if (P <= 0) return GM_OK;
	launch_k(densify_stats_kernel, dim3((P + kThreads - 1) / kThreads), dim3(kThreads), 0, stream, P, radii, dL_dmean2D, max_radii2D,
	                                                                               grad_accum, denom, skip_flag);
	return GM_OK;
}

} // namespace gm
