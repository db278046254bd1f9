// Host-side entry points of libCudaRasterizer.so:
//   * the flat C ABI declared in include/gm_rasterizer.h, and
//   * the C++ symbols of namespace CudaRasterizer that the reference's Jittor glue links against
//     (declared in csrc/cuda_rasterizer/rasterizer.h, mirroring dgr/cuda_rasterizer/rasterizer.h).
// No device memory is allocated here; all scratch lives in the caller's three chunks.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace gm {

static thread_local std::string g_last_error;

void set_last_error(const char* what, cudaError_t err)
{
	g_last_error = std::string(what) + ": " + cudaGetErrorString(err);
}

bool pdl_enabled()
{
	static const bool on = !(std::getenv("GM_PDL") != nullptr && std::getenv("GM_PDL")[0] == '0');
	return on;
}

// The reference's CHECK_CUDA (auxiliary.h:165-172): synchronise + check only in debug mode.  Launch
// configuration errors are cheap to see, so they are reported in both modes.
int check_stage(const char* what, bool debug, cudaStream_t stream)
{
	cudaError_t err = cudaGetLastError();
	if (err == cudaSuccess && debug)
		err = cudaStreamSynchronize(stream);
	if (err == cudaSuccess && debug)
		err = cudaGetLastError();
	if (err != cudaSuccess) {
		set_last_error(what, err);
		if (debug)
			fprintf(stderr, "\n[CUDA ERROR] in stage %s: %s\n", what, cudaGetErrorString(err));
		return GM_ERR_CUDA;
	}
	return GM_OK;
}

// ---- optional per-stage timing (gm_profile_begin / gm_profile_end) -----------------------------
// When enabled, every stage launch is bracketed by two events recorded on the launch stream; the
// elapsed times are summed per stage in gm_profile_end.  Used by bench.py for the roofline of the
// dominant kernel; off by default (no events, no overhead).
enum Stage { kStDepthBuckets, kStPreprocess, kStTileScan, kStEmit, kStSortPack, kStBlendFwd, kStBlendBwd, kStGeomBwd, kStL1,
             kStMeshBindFwd, kStMeshBindBwd, kStDeform, kStShRotated, kStMarkVisible, kStAcapRest, kStAcapGetRS,
             kStPhotometric, kStMeshRestrict, kStAdam, kStDensifyStats, kStCov3dPython, kStLoadMesh,
             kNumStages };
static const char* const kStageNames[kNumStages] = {
	"depth_buckets", "preprocess", "tile_scan", "emit", "sort_pack", "blend_forward", "blend_backward", "geometry_backward",
	"l1_loss", "mesh_bind_forward", "mesh_bind_backward", "deform", "sh_to_rgb_rotated", "mark_visible", "acap_rest",
	"acap_get_rs", "photometric_loss", "mesh_restrict_loss", "adam", "densify_stats", "cov3d_python", "load_mesh"};

struct StageRecord { int stage; cudaEvent_t start, stop; };
static std::mutex g_profile_mutex;
static bool g_profile_on = false;
static std::vector<StageRecord> g_profile_records;
static uint64_t g_launch_counts[kNumStages];

struct StageScope {
	cudaStream_t stream;
	cudaEvent_t stop = nullptr;
	StageScope(int stage, cudaStream_t s) : stream(s)
	{
		if (!g_profile_on)
			return;
		std::lock_guard<std::mutex> lock(g_profile_mutex);
		g_launch_counts[stage]++;
		StageRecord r;
		r.stage = stage;
		if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess)
			return;
		cudaEventRecord(r.start, stream);
		stop = r.stop;
		g_profile_records.push_back(r);
	}
	~StageScope()
	{
		if (stop != nullptr)
			cudaEventRecord(stop, stream);
	}
};

static bool make_view(ViewParams& vp, int D, int M, const float* background, int width, int height,
                      float scale_modifier, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      float tan_fovx, float tan_fovy)
{
	vp.view = viewmatrix;
	vp.proj = projmatrix;
	vp.campos = cam_pos;
	vp.bg = background;
	vp.tan_fovx = tan_fovx;
	vp.tan_fovy = tan_fovy;
	// rasterizer_impl.cu:359-360
	vp.focal_y = height / (2.0f * tan_fovy);
	vp.focal_x = width / (2.0f * tan_fovx);
	vp.scale_modifier = scale_modifier;
	vp.W = width;
	vp.H = height;
	vp.tiles_x = (width + kTile - 1) / kTile;
	vp.tiles_y = (height + kTile - 1) / kTile;
	vp.D = D;
	vp.M = M;
	vp.bucket_log2 = bucket_log2_for(vp.tiles_x * vp.tiles_y);
	return width > 0 && height > 0;
}

// Largest instance count whose BinningState fits in `bytes`.
static uint32_t binning_capacity_instances(size_t bytes)
{
	// 48 bytes per instance (8 key + 16 + 16 + 8 record) plus alignment slack: start from the upper bound
	// and step down to the largest multiple of 8 whose carved chunk fits.
	size_t r = (bytes / 48) & ~(size_t)7;
	if (r > 0xfffffff0u)
		r = 0xfffffff0u;
	while (r > 0 && required<BinningState>(r) > bytes)
		r -= 8;
	return (uint32_t)r;
}

// The kernels read a quaternion row with one 128-bit access: [P,4] float rows are 16 bytes, so only the base can be off.
static bool misaligned16(const void* p)
{
	return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) != 0;
}

static int reject_misaligned(const char* what)
{
	g_last_error = std::string(what) + " must be 16-byte aligned (quaternion rows are read as float4)";
	return GM_ERR_BAD_ARGUMENT;
}

static int forward_stage0(char* geom_buffer, int P, const ViewParams& vp, const float* means3D, const float* shs,
                          const float* colors_precomp, const float* opacities, const float* scales,
                          const float* rotations, const float* cov3D_precomp, bool prefiltered, int* radii,
                          uint32_t capacity, bool debug, cudaStream_t stream, GeometryState& geom)
{
	if (P < 0 || means3D == nullptr || opacities == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	if (colors_precomp == nullptr && shs == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
		return GM_ERR_BAD_ARGUMENT;
	if (cov3D_precomp == nullptr && misaligned16(rotations))
		return reject_misaligned("rotations");
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	if (num_tiles > GM_MAX_TILES)
		return GM_ERR_TOO_MANY_TILES;

	geom = GeometryState::fromChunk(geom_buffer, (size_t)P);
	if (radii == nullptr)
		radii = geom.internal_radii;

	{ StageScope scope_(kStDepthBuckets, stream); launch_depth_buckets(P, means3D, vp, geom, stream); }
	if (int rc = check_stage("depth_buckets", debug, stream)) return rc;
	{ StageScope scope_(kStPreprocess, stream); launch_preprocess(P, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, vp, radii, geom,
	                  prefiltered, stream); }
	if (int rc = check_stage("preprocess", debug, stream)) return rc;
	{ StageScope scope_(kStTileScan, stream); launch_tile_scan(num_tiles, geom, capacity, vp, stream); }
	if (int rc = check_stage("tile_scan", debug, stream)) return rc;
	return GM_OK;
}

static int forward_stage1(const GeometryState& geom, char* binning_buffer, char* image_buffer, int P,
                          const ViewParams& vp, uint32_t capacity, const int* radii, float* out_color, bool debug,
                          cudaStream_t stream, const gm_forward_epilogue* epilogue = nullptr)
{
	if (out_color == nullptr || image_buffer == nullptr || (capacity > 0 && binning_buffer == nullptr))
		return GM_ERR_BAD_ARGUMENT;
	const int num_tiles = vp.tiles_x * vp.tiles_y;
	BinningState binning = BinningState::fromChunk(binning_buffer, (size_t)capacity);
	ImageState img = ImageState::fromChunk(image_buffer, (size_t)vp.W * vp.H);
	if (radii == nullptr)
		radii = geom.internal_radii;

	{ StageScope scope_(kStEmit, stream); launch_emit(P, radii, geom, binning, capacity, vp, stream); }
	if (int rc = check_stage("emit", debug, stream)) return rc;
	{ StageScope scope_(kStSortPack, stream); launch_sort_pack(num_tiles, geom, binning, capacity, vp, stream); }
	if (int rc = check_stage("sort_pack", debug, stream)) return rc;
	bool folded = false;
	{ StageScope scope_(kStBlendFwd, stream); launch_blend_forward(geom, binning, img, capacity, vp, out_color, epilogue, &folded, stream); }
	if (int rc = check_stage("blend_forward", debug, stream)) return rc;
	if (epilogue != nullptr && !folded) {
		// an A/B kernel variant that has no epilogue: the same work as separate launches
		if (epilogue->loss != nullptr && epilogue->target != nullptr) {
			StageScope scope_(kStL1, stream);
			launch_l1((size_t)3 * vp.W * vp.H, out_color, epilogue->target, epilogue->target_is_u8, epilogue->loss, epilogue->dL_dimg, stream);
		}
		if (epilogue->zero_ptr != nullptr && epilogue->zero_floats > 0)
			cudaMemsetAsync(epilogue->zero_ptr, 0, epilogue->zero_floats * sizeof(float), stream);
		if (int rc = check_stage("forward epilogue", debug, stream)) return rc;
	}
	return GM_OK;
}

} // namespace gm

using namespace gm;

extern "C" {

const char* gm_version(void) { return "gaussianmesh-b200 0.1.0 sm_100a"; }
const char* gm_last_error(void) { return g_last_error.c_str(); }

int gm_profile_num_stages(void) { return kNumStages; }
const char* gm_profile_stage_name(int stage) { return (stage >= 0 && stage < kNumStages) ? kStageNames[stage] : ""; }

void gm_profile_begin(void)
{
	std::lock_guard<std::mutex> lock(g_profile_mutex);
	for (auto& r : g_profile_records) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
	g_profile_records.clear();
	memset(g_launch_counts, 0, sizeof(g_launch_counts));
	g_profile_on = true;
}

int gm_profile_end(float* ms_per_stage, uint64_t* launches_per_stage)
{
	std::lock_guard<std::mutex> lock(g_profile_mutex);
	g_profile_on = false;
	int rc = GM_OK;
	for (int i = 0; i < kNumStages; i++) {
		if (ms_per_stage) ms_per_stage[i] = 0.0f;
		if (launches_per_stage) launches_per_stage[i] = g_launch_counts[i];
	}
	for (auto& r : g_profile_records) {
		float ms = 0.0f;
		cudaError_t err = cudaEventSynchronize(r.stop);
		if (err == cudaSuccess)
			err = cudaEventElapsedTime(&ms, r.start, r.stop);
		if (err != cudaSuccess) { set_last_error("profile_end", err); rc = GM_ERR_CUDA; }
		else if (ms_per_stage) ms_per_stage[r.stage] += ms;
		cudaEventDestroy(r.start);
		cudaEventDestroy(r.stop);
	}
	g_profile_records.clear();
	return rc;
}

size_t gm_required_geom(size_t P) { return required<GeometryState>(P); }
size_t gm_required_image(size_t N) { return required<ImageState>(N); }
size_t gm_required_binning(size_t R) { return required<BinningState>(R); }
size_t gm_binning_capacity(size_t bytes) { return binning_capacity_instances(bytes); }

int gm_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, gm_stream_t stream)
{
	(void)projmatrix;   // the reference computes p_proj but only tests view-space z (auxiliary.h:153)
	if (P < 0 || (P > 0 && (means3D == nullptr || viewmatrix == nullptr || present == nullptr)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStMarkVisible, (cudaStream_t)stream); launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream); }
	return check_stage("mark_visible", false, (cudaStream_t)stream);
}

int gm_forward_0(char* geom_buffer, int P, int D, int M, const float* background, int width, int height,
                 const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                 const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                 float tan_fovy, int prefiltered, int* radii, int debug, gm_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	ViewParams vp;
	if (!make_view(vp, D, M, background, width, height, scale_modifier, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy))
		return GM_ERR_BAD_ARGUMENT;
	if (geom_buffer == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	GeometryState geom;
	if (int rc = forward_stage0(geom_buffer, P, vp, means3D, shs, colors_precomp, opacities, scales, rotations,
	                            cov3D_precomp, prefiltered != 0, radii, 0xfffffff0u, debug != 0, stream, geom))
		return rc;
	// rasterizer_impl.cu:409-411: the instance total goes back to the host (blocking)
	uint32_t num_rendered = 0;
	cudaError_t err = cudaMemcpyAsync(&num_rendered, &geom.header->num_rendered, sizeof(uint32_t),
	                                  cudaMemcpyDeviceToHost, stream);
	if (err == cudaSuccess)
		err = cudaStreamSynchronize(stream);
	if (err != cudaSuccess) {
		set_last_error("forward_0 readback", err);
		return GM_ERR_CUDA;
	}
	return (int)num_rendered;
}

int gm_forward_1(char* geom_buffer, char* binning_buffer, char* image_buffer, int P, int D, int M,
                 int num_rendered, const float* background, int width, int height, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                 float scale_modifier, const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                 float tan_fovy, int prefiltered, float* out_color, int* radii, int debug, gm_stream_t stream_)
{
	(void)means3D; (void)shs; (void)colors_precomp; (void)opacities; (void)scales; (void)rotations;
	(void)cov3D_precomp; (void)prefiltered;
	cudaStream_t stream = (cudaStream_t)stream_;
	ViewParams vp;
	if (!make_view(vp, D, M, background, width, height, scale_modifier, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy))
		return GM_ERR_BAD_ARGUMENT;
	if (geom_buffer == nullptr || num_rendered < 0 || P < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (vp.tiles_x * vp.tiles_y > GM_MAX_TILES)
		return GM_ERR_TOO_MANY_TILES;
	GeometryState geom = GeometryState::fromChunk(geom_buffer, (size_t)P);
	return forward_stage1(geom, binning_buffer, image_buffer, P, vp, (uint32_t)num_rendered, radii, out_color,
	                      debug != 0, stream);
}

int gm_forward_ex(char* geom_buffer, char* binning_buffer, size_t binning_capacity, char* image_buffer, int P,
                  int D, int M, const float* background, int width, int height, const float* means3D,
                  const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                  float scale_modifier, const float* rotations, const float* cov3D_precomp,
                  const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                  float tan_fovy, int prefiltered, float* out_color, int* radii, int debug,
                  uint32_t* frame_info_host, const gm_forward_epilogue* epilogue, gm_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	ViewParams vp;
	if (!make_view(vp, D, M, background, width, height, scale_modifier, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy))
		return GM_ERR_BAD_ARGUMENT;
	if (geom_buffer == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	if (epilogue != nullptr) {
		if ((reinterpret_cast<uintptr_t>(epilogue->zero_ptr) & 15u) != 0 || (epilogue->zero_floats & 3u) != 0)
			return GM_ERR_BAD_ARGUMENT;
		// the loss is accumulated per tile: cleared here, ahead of the whole launch chain
		if (epilogue->loss != nullptr && cudaMemsetAsync(epilogue->loss, 0, sizeof(float), stream) != cudaSuccess)
			return GM_ERR_CUDA;
	}
	const uint32_t capacity = binning_capacity_instances(binning_capacity);
	GeometryState geom;
	if (int rc = forward_stage0(geom_buffer, P, vp, means3D, shs, colors_precomp, opacities, scales, rotations,
	                            cov3D_precomp, prefiltered != 0, radii, capacity, debug != 0, stream, geom))
		return rc;
	if (frame_info_host != nullptr) {
		// FrameHeader starts with {num_rendered, num_visible, overflow, capacity}
		cudaError_t err = cudaMemcpyAsync(frame_info_host, geom.header, 4 * sizeof(uint32_t),
		                                  cudaMemcpyDeviceToHost, stream);
		if (err != cudaSuccess) {
			set_last_error("forward count readback", err);
			return GM_ERR_CUDA;
		}
	}
	return forward_stage1(geom, binning_buffer, image_buffer, P, vp, capacity, radii, out_color, debug != 0, stream, epilogue);
}

int gm_forward(char* geom_buffer, char* binning_buffer, size_t binning_capacity, char* image_buffer, int P,
               int D, int M, const float* background, int width, int height, const float* means3D,
               const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
               float scale_modifier, const float* rotations, const float* cov3D_precomp,
               const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
               float tan_fovy, int prefiltered, float* out_color, int* radii, int debug,
               uint32_t* frame_info_host, gm_stream_t stream_)
{
	return gm_forward_ex(geom_buffer, binning_buffer, binning_capacity, image_buffer, P, D, M, background, width, height,
	                     means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
	                     projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, radii, debug, frame_info_host,
	                     nullptr, stream_);
}

int gm_forward_status(const char* geom_buffer, int* num_rendered, int* num_visible, gm_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (geom_buffer == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	char* p = const_cast<char*>(geom_buffer);
	GeometryState geom = GeometryState::fromChunk(p, 0);
	FrameHeader h;
	cudaError_t err = cudaMemcpyAsync(&h, geom.header, sizeof(FrameHeader), cudaMemcpyDeviceToHost, stream);
	if (err == cudaSuccess)
		err = cudaStreamSynchronize(stream);
	if (err != cudaSuccess) {
		set_last_error("forward_status", err);
		return GM_ERR_CUDA;
	}
	if (num_rendered) *num_rendered = (int)h.num_rendered;
	if (num_visible) *num_visible = (int)h.num_visible;
	return h.overflow ? GM_ERR_BINNING_OVERFLOW : GM_OK;
}

void gm_geom_view(char* geom_buffer, size_t P, void** out6)
{
	GeometryState g = GeometryState::fromChunk(geom_buffer, P);
	out6[0] = g.depths; out6[1] = g.means2D; out6[2] = g.cov3D; out6[3] = g.conic_opacity;
	out6[4] = g.rgb_clamp; out6[5] = g.tile_count;
}

int gm_backward_ex(int P, int D, int M, int R, const float* background, int width, int height,
                const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                float scale_modifier, const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                float* dL_drot, int debug, int flags, gm_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	ViewParams vp;
	if (!make_view(vp, D, M, background, width, height, scale_modifier, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy))
		return GM_ERR_BAD_ARGUMENT;
	if (P < 0 || R < 0 || geom_buffer == nullptr || image_buffer == nullptr || dL_dpix == nullptr ||
	    dL_dmean2D == nullptr || dL_dconic == nullptr || dL_dopacity == nullptr || dL_dcolor == nullptr ||
	    dL_dmean3D == nullptr || dL_dcov3D == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	// rasterizer_impl.cu:576-596: SH / scale-rot gradients only exist on those paths
	const bool sh_path = (colors_precomp == nullptr);
	const bool sr_path = (cov3D_precomp == nullptr);
	if ((sh_path && (shs == nullptr || dL_dsh == nullptr)) ||
	    (sr_path && (scales == nullptr || rotations == nullptr || dL_dscale == nullptr || dL_drot == nullptr)))
		return GM_ERR_BAD_ARGUMENT;
	if (vp.tiles_x * vp.tiles_y > GM_MAX_TILES)
		return GM_ERR_TOO_MANY_TILES;
	if (sr_path && (misaligned16(rotations) || misaligned16(dL_drot)))
		return reject_misaligned("rotations / dL_drot");
	if (P == 0)
		return GM_OK;

	GeometryState geom = GeometryState::fromChunk(geom_buffer, (size_t)P);
	BinningState binning = BinningState::fromChunk(binning_buffer, (size_t)R);
	ImageState img = ImageState::fromChunk(image_buffer, (size_t)width * height);
	if (radii == nullptr)
		radii = geom.internal_radii;

	{ StageScope scope_(kStBlendBwd, stream); launch_blend_backward(geom, binning, img, (uint32_t)R, vp, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity,
	                      dL_dcolor, stream); }
	if (int rc = check_stage("blend_backward", debug != 0, stream)) return rc;

	const float* cov3D_ptr = sr_path ? geom.cov3D : cov3D_precomp;
	{ StageScope scope_(kStGeomBwd, stream); launch_geometry_backward(P, means3D, radii, sh_path ? shs : nullptr, sr_path ? scales : nullptr,
	                         sr_path ? rotations : nullptr, cov3D_ptr, vp, geom, dL_dmean2D, dL_dconic, dL_dcolor,
	                         dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, (flags & GM_BACKWARD_OVERWRITE) != 0, stream); }
	return check_stage("geometry_backward", debug != 0, stream);
}

int gm_backward(int P, int D, int M, int R, const float* background, int width, int height,
                const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                float scale_modifier, const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                float* dL_drot, int debug, gm_stream_t stream_)
{
	return gm_backward_ex(P, D, M, R, background, width, height, means3D, shs, colors_precomp, scales, scale_modifier,
	                      rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii,
	                      geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity,
	                      dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug, 0, stream_);
}

int gm_mesh_bind_forward(int P, const float* bc_logits, const float* distance, const float* vertex1,
                         const float* vertex2, const float* vertex3, const float* normal, const float* r,
                         float alpha_distance, const float* log_scale, const float* rot_raw,
                         const float* opacity_logit, float* xyz, float* scale, float* rot, float* opacity,
                         gm_stream_t stream)
{
	if (P < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (xyz != nullptr && (!bc_logits || !distance || !vertex1 || !vertex2 || !vertex3 || !normal || !r))
		return GM_ERR_BAD_ARGUMENT;
	if ((scale != nullptr && !log_scale) || (rot != nullptr && !rot_raw) || (opacity != nullptr && !opacity_logit))
		return GM_ERR_BAD_ARGUMENT;
	if (rot != nullptr && (misaligned16(rot) || misaligned16(rot_raw)))
		return reject_misaligned("rot / rot_raw");
	{ StageScope scope_(kStMeshBindFwd, (cudaStream_t)stream); launch_mesh_bind_forward(P, bc_logits, distance, vertex1, vertex2, vertex3, normal, r, alpha_distance, log_scale,
	                         rot_raw, opacity_logit, xyz, scale, rot, opacity, (cudaStream_t)stream); }
	return check_stage("mesh_bind_forward", false, (cudaStream_t)stream);
}

int gm_mesh_bind_backward(int P, const float* bc_logits, const float* distance, const float* vertex1,
                          const float* vertex2, const float* vertex3, const float* normal, const float* r,
                          float alpha_distance, const float* log_scale, const float* rot_raw,
                          const float* opacity_logit, const float* dL_dxyz, const float* dL_dscale,
                          const float* dL_drot, const float* dL_dopacity, float* dL_dbc_logits,
                          float* dL_ddistance, float* dL_dlog_scale, float* dL_drot_raw,
                          float* dL_dopacity_logit, gm_stream_t stream)
{
	if (P < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (dL_dxyz != nullptr && dL_dbc_logits != nullptr &&
	    (!bc_logits || !distance || !vertex1 || !vertex2 || !vertex3 || !normal || !r))
		return GM_ERR_BAD_ARGUMENT;
	if ((dL_dscale && dL_dlog_scale && !log_scale) || (dL_drot && dL_drot_raw && !rot_raw) ||
	    (dL_dopacity && dL_dopacity_logit && !opacity_logit))
		return GM_ERR_BAD_ARGUMENT;
	if (dL_drot && dL_drot_raw && (misaligned16(dL_drot) || misaligned16(dL_drot_raw) || misaligned16(rot_raw)))
		return reject_misaligned("rot_raw / dL_drot / dL_drot_raw");
	{ StageScope scope_(kStMeshBindBwd, (cudaStream_t)stream); launch_mesh_bind_backward(P, bc_logits, distance, vertex1, vertex2, vertex3, normal, r, alpha_distance, log_scale,
	                          rot_raw, opacity_logit, dL_dxyz, dL_dscale, dL_drot, dL_dopacity, dL_dbc_logits,
	                          dL_ddistance, dL_dlog_scale, dL_drot_raw, dL_dopacity_logit, (cudaStream_t)stream); }
	return check_stage("mesh_bind_backward", false, (cudaStream_t)stream);
}

int gm_deform_gaussians(int P, int num_vertices, const float* vertex_rest, const float* vertex_deformed,
                        const float* vertex_R, const float* vertex_S, const int* gaussian_triangles,
                        const float* weights, const float* pos_in, const float* cov_in, int cov_in_is_full,
                        float* pos_out, float* cov6_out, float* rot_out, gm_stream_t stream)
{
	if (P < 0 || num_vertices < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (P > 0 && (!vertex_rest || !vertex_deformed || !vertex_R || !vertex_S || !gaussian_triangles || !weights ||
	              !pos_in || !cov_in || !pos_out || !cov6_out))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStDeform, (cudaStream_t)stream); launch_deform(P, vertex_rest, vertex_deformed, vertex_R, vertex_S, gaussian_triangles, weights, pos_in, cov_in,
	              cov_in_is_full, pos_out, cov6_out, rot_out, (cudaStream_t)stream); }
	return check_stage("deform", false, (cudaStream_t)stream);
}

int gm_sh_to_rgb_rotated(int P, int D, int M, const float* pos, const float* campos, const float* rot,
                         const float* shs, float* rgb, gm_stream_t stream)
{
	if (P < 0 || D < 0 || D > 3 || M < (D + 1) * (D + 1))
		return GM_ERR_BAD_ARGUMENT;
	if (P > 0 && (!pos || !campos || !shs || !rgb))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStShRotated, (cudaStream_t)stream); launch_sh_rotated(P, D, M, pos, campos, rot, shs, rgb, (cudaStream_t)stream); }
	return check_stage("sh_to_rgb_rotated", false, (cudaStream_t)stream);
}

int gm_sh_to_rgb_rotated_backward(int P, int D, int M, const float* pos, const float* campos, const float* rot,
                                  const float* shs, const float* dL_drgb, float* dL_dshs, float* dL_dpos, gm_stream_t stream)
{
	if (P < 0 || D < 0 || D > 3 || M < (D + 1) * (D + 1))
		return GM_ERR_BAD_ARGUMENT;
	if (P > 0 && (!pos || !campos || !shs || !dL_drgb || (!dL_dshs && !dL_dpos)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStShRotated, (cudaStream_t)stream);
	  launch_sh_rotated_backward(P, D, M, pos, campos, rot, shs, dL_drgb, dL_dshs, dL_dpos, (cudaStream_t)stream); }
	return check_stage("sh_to_rgb_rotated_backward", false, (cudaStream_t)stream);
}

int gm_cov3d_from_scale_rot(int P, const float* scales, float scale_modifier, const float* rotations, float* cov6,
                            gm_stream_t stream)
{
	if (P < 0 || (P > 0 && (!scales || !rotations || !cov6)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStCov3dPython, (cudaStream_t)stream);
	  launch_cov3d_python(P, scales, scale_modifier, rotations, cov6, (cudaStream_t)stream); }
	return check_stage("cov3d_from_scale_rot", false, (cudaStream_t)stream);
}

int gm_cov3d_from_scale_rot_backward(int P, const float* scales, float scale_modifier, const float* rotations,
                                     const float* dL_dcov6, float* dL_dscale, float* dL_drot, gm_stream_t stream)
{
	if (P < 0 || (P > 0 && (!scales || !rotations || !dL_dcov6 || !dL_dscale || !dL_drot)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStCov3dPython, (cudaStream_t)stream);
	  launch_cov3d_python_backward(P, scales, scale_modifier, rotations, dL_dcov6, dL_dscale, dL_drot, (cudaStream_t)stream); }
	return check_stage("cov3d_from_scale_rot_backward", false, (cudaStream_t)stream);
}

int gm_load_mesh(int P, int num_vertices, int num_faces, const double* vertex, const int32_t* faces, const int64_t* face_id,
                 const float* proj_pos, int32_t* gaussian_triangles, double* weights, gm_stream_t stream)
{
	if (P < 0 || num_vertices < 0 || num_faces < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (P > 0 && (num_faces == 0 || !vertex || !faces || !face_id || !proj_pos || !gaussian_triangles || !weights))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStLoadMesh, (cudaStream_t)stream);
	  launch_load_mesh(P, num_faces, vertex, faces, reinterpret_cast<const long long*>(face_id), proj_pos, gaussian_triangles,
	                   weights, (cudaStream_t)stream); }
	return check_stage("load_mesh", false, (cudaStream_t)stream);
}

int gm_acap_build_rings(int num_vertices, int num_faces, const int32_t* faces_host, int32_t* ring_offsets_host,
                        int32_t* ring_neighbours_host, int32_t* face_offsets_host, int32_t* face_list_host)
{
	if (num_vertices < 0 || num_faces < 0 || (num_faces > 0 && !faces_host) || !ring_offsets_host || !face_offsets_host ||
	    (num_faces > 0 && (!ring_neighbours_host || !face_list_host)))
		return GM_ERR_BAD_ARGUMENT;
	return acap_build_rings_host(num_vertices, num_faces, faces_host, ring_offsets_host, ring_neighbours_host,
	                             face_offsets_host, face_list_host);
}

int gm_acap_rest(int num_vertices, const double* vertex_rest, const int32_t* faces, const int32_t* ring_offsets,
                 const int32_t* ring_neighbours, const int32_t* face_offsets, const int32_t* face_list,
                 double* sqrt_w, double* rest_normals, double* ata_inv, gm_stream_t stream)
{
	if (num_vertices < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (num_vertices > 0 && (!vertex_rest || !faces || !ring_offsets || !ring_neighbours || !face_offsets || !face_list ||
	                         !sqrt_w || !rest_normals || !ata_inv))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStAcapRest, (cudaStream_t)stream);
	  launch_acap_rest(num_vertices, vertex_rest, faces, ring_offsets, ring_neighbours, face_offsets, face_list, sqrt_w,
	                   rest_normals, ata_inv, (cudaStream_t)stream); }
	return check_stage("acap_rest", false, (cudaStream_t)stream);
}

int gm_acap_get_rs(int num_vertices, const double* vertex_rest, const double* vertex_deformed, const int32_t* faces,
                   const int32_t* ring_offsets, const int32_t* ring_neighbours, const int32_t* face_offsets,
                   const int32_t* face_list, const double* sqrt_w, const double* rest_normals, const double* ata_inv,
                   double* normals_scratch, float* R_out, float* S_out, gm_stream_t stream)
{
	if (num_vertices < 0)
		return GM_ERR_BAD_ARGUMENT;
	if (num_vertices > 0 && (!vertex_rest || !vertex_deformed || !faces || !ring_offsets || !ring_neighbours ||
	                         !face_offsets || !face_list || !sqrt_w || !rest_normals || !ata_inv || !normals_scratch ||
	                         !R_out || !S_out))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStAcapGetRS, (cudaStream_t)stream);
	  launch_acap_get_rs(num_vertices, vertex_rest, vertex_deformed, faces, ring_offsets, ring_neighbours, face_offsets,
	                     face_list, sqrt_w, rest_normals, ata_inv, normals_scratch, R_out, S_out, (cudaStream_t)stream); }
	return check_stage("acap_get_rs", false, (cudaStream_t)stream);
}

int gm_l1_loss(size_t numel, const float* img, const float* target, float* loss, float* dL_dimg, gm_stream_t stream)
{
	if (loss == nullptr || (numel > 0 && (!img || !target)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStL1, (cudaStream_t)stream); launch_l1(numel, img, target, 0, loss, dL_dimg, (cudaStream_t)stream); }
	return check_stage("l1_loss", false, (cudaStream_t)stream);
}

int gm_l1_loss_u8(size_t numel, const float* img, const uint8_t* target, float* loss, float* dL_dimg, gm_stream_t stream)
{
	if (loss == nullptr || (numel > 0 && (!img || !target)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStL1, (cudaStream_t)stream); launch_l1(numel, img, target, 1, loss, dL_dimg, (cudaStream_t)stream); }
	return check_stage("l1_loss", false, (cudaStream_t)stream);
}

int gm_image_u8_to_float(size_t numel, const uint8_t* src, float* dst, gm_stream_t stream)
{
	if (numel > 0 && (!src || !dst))
		return GM_ERR_BAD_ARGUMENT;
	launch_u8_to_float(numel, src, dst, (cudaStream_t)stream);
	return check_stage("image_u8_to_float", false, (cudaStream_t)stream);
}

size_t gm_photometric_scratch_bytes(int C, int H, int W) { return photometric_scratch_bytes(C, H, W); }

int gm_photometric_loss(int C, int H, int W, const float* img, const float* gt, float lambda_dssim, char* scratch,
                        float* out, float* dL_dimg, gm_stream_t stream)
{
	if (C < 0 || H < 0 || W < 0 || out == nullptr || scratch == nullptr)
		return GM_ERR_BAD_ARGUMENT;
	if ((size_t)C * H * W > 0 && (!img || !gt))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStPhotometric, (cudaStream_t)stream);
	  launch_photometric(C, H, W, img, gt, lambda_dssim, scratch, out, dL_dimg, (cudaStream_t)stream); }
	return check_stage("photometric_loss", false, (cudaStream_t)stream);
}

int gm_mesh_restrict_loss(int P, const float* scale, const float* vertex1, const float* vertex2, const float* vertex3,
                          float weight, float* loss, float* dL_dscale, int accumulate, gm_stream_t stream)
{
	if (P < 0 || loss == nullptr || (P > 0 && (!scale || !vertex1 || !vertex2 || !vertex3)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStMeshRestrict, (cudaStream_t)stream);
	  launch_mesh_restrict(P, scale, vertex1, vertex2, vertex3, weight, loss, dL_dscale, accumulate, (cudaStream_t)stream); }
	return check_stage("mesh_restrict_loss", false, (cudaStream_t)stream);
}

int gm_adam_step_gated(int num_tensors, const gm_adam_tensor* tensors_host, int step, float beta1, float beta2, float eps,
                       const uint32_t* skip_flag, gm_stream_t stream)
{
	if (num_tensors < 0 || step < 1 || (num_tensors > 0 && tensors_host == nullptr))
		return GM_ERR_BAD_ARGUMENT;
	for (int i = 0; i < num_tensors; i++) {
		const gm_adam_tensor& t = tensors_host[i];
		if (t.numel > 0 && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq))
			return GM_ERR_BAD_ARGUMENT;
		if (t.period > 0 && t.split > t.period)
			return GM_ERR_BAD_ARGUMENT;
	}
	{ StageScope scope_(kStAdam, (cudaStream_t)stream);
	  launch_adam(num_tensors, tensors_host, step, beta1, beta2, eps, skip_flag, (cudaStream_t)stream); }
	return check_stage("adam", false, (cudaStream_t)stream);
}

int gm_adam_step(int num_tensors, const gm_adam_tensor* tensors_host, int step, float beta1, float beta2, float eps,
                 gm_stream_t stream)
{
	return gm_adam_step_gated(num_tensors, tensors_host, step, beta1, beta2, eps, nullptr, stream);
}

const uint32_t* gm_frame_overflow_flag(const char* geom_buffer)
{
	if (geom_buffer == nullptr)
		return nullptr;
	char* p = const_cast<char*>(geom_buffer);
	GeometryState geom = GeometryState::fromChunk(p, 0);
	return &geom.header->overflow;
}

void gm_adam_shard_range(size_t total, int world, int rank, size_t* lo, size_t* hi)
{
	// contiguous shards of whole 32-float blocks (a block never straddles two tensors)
	const size_t blocks = (total + 31) / 32;
	const size_t base = world > 0 ? blocks / (size_t)world : 0, extra = world > 0 ? blocks % (size_t)world : 0;
	const size_t r = (size_t)rank;
	const size_t b0 = r * base + (r < extra ? r : extra);
	const size_t b1 = b0 + base + (r < extra ? 1 : 0);
	*lo = b0 * 32 < total ? b0 * 32 : total;
	*hi = b1 * 32 < total ? b1 * 32 : total;
}

int gm_adam_step_sharded_p2p(int world, int rank, const float* const* grads_host, float* const* params_host,
                             int num_segments, const gm_adam_segment* segments_host, size_t total, float* exp_avg,
                             float* exp_avg_sq, int step, float beta1, float beta2, float eps, gm_stream_t stream)
{
	if (world < 1 || world > GM_MAX_PEERS || rank < 0 || rank >= world || num_segments < 0 || num_segments > 8 || step < 1)
		return GM_ERR_BAD_ARGUMENT;
	if (!grads_host || !params_host || (num_segments > 0 && !segments_host) || (total & 3) != 0)
		return GM_ERR_BAD_ARGUMENT;
	for (int q = 0; q < world; q++)
		if (!grads_host[q] || !params_host[q])
			return GM_ERR_BAD_ARGUMENT;
	for (int s = 0; s < num_segments; s++) {
		const gm_adam_segment& sg = segments_host[s];
		if ((sg.offset & 31) != 0 || sg.offset + sg.numel > total || (sg.period > 0 && sg.split > sg.period))
			return GM_ERR_BAD_ARGUMENT;
	}
	size_t lo, hi;
	gm_adam_shard_range(total, world, rank, &lo, &hi);
	if (hi > lo && (!exp_avg || !exp_avg_sq))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStAdam, (cudaStream_t)stream);
	  launch_adam_sharded_p2p(world, rank, grads_host, params_host, num_segments, segments_host, total, exp_avg,
	                          exp_avg_sq, step, beta1, beta2, eps, (cudaStream_t)stream); }
	return check_stage("adam_sharded_p2p", false, (cudaStream_t)stream);
}

int gm_adam_step_sharded_mc(int world, int rank, const float* grads_multicast, float* params_multicast,
                            const float* params_local, int num_segments, const gm_adam_segment* segments_host, size_t total,
                            float* exp_avg, float* exp_avg_sq, int step, float beta1, float beta2, float eps, gm_stream_t stream)
{
	if (world < 1 || world > GM_MAX_PEERS || rank < 0 || rank >= world || num_segments < 0 || num_segments > 8 || step < 1)
		return GM_ERR_BAD_ARGUMENT;
	if (!grads_multicast || !params_multicast || !params_local || (num_segments > 0 && !segments_host) || (total & 3) != 0)
		return GM_ERR_BAD_ARGUMENT;
	for (int s = 0; s < num_segments; s++) {
		const gm_adam_segment& sg = segments_host[s];
		if ((sg.offset & 31) != 0 || sg.offset + sg.numel > total || (sg.period > 0 && sg.split > sg.period))
			return GM_ERR_BAD_ARGUMENT;
	}
	size_t lo, hi;
	gm_adam_shard_range(total, world, rank, &lo, &hi);
	if (hi > lo && (!exp_avg || !exp_avg_sq))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStAdam, (cudaStream_t)stream);
	  launch_adam_sharded_mc(world, rank, grads_multicast, params_multicast, params_local, num_segments, segments_host, total,
	                         exp_avg, exp_avg_sq, step, beta1, beta2, eps, (cudaStream_t)stream); }
	return check_stage("adam_sharded_mc", false, (cudaStream_t)stream);
}

int gm_densify_stats_gated(int P, const int32_t* radii, const float* dL_dmean2D, float* max_radii2D, float* grad_accum,
                           float* denom, const uint32_t* skip_flag, gm_stream_t stream)
{
	if (P < 0 || (P > 0 && (!radii || !dL_dmean2D || !max_radii2D || !grad_accum || !denom)))
		return GM_ERR_BAD_ARGUMENT;
	{ StageScope scope_(kStDensifyStats, (cudaStream_t)stream);
	  launch_densify_stats(P, radii, dL_dmean2D, max_radii2D, grad_accum, denom, skip_flag, (cudaStream_t)stream); }
	return check_stage("densify_stats", false, (cudaStream_t)stream);
}

int gm_densify_stats(int P, const int32_t* radii, const float* dL_dmean2D, float* max_radii2D, float* grad_accum,
                     float* denom, gm_stream_t stream)
{
	return gm_densify_stats_gated(P, radii, dL_dmean2D, max_radii2D, grad_accum, denom, nullptr, stream);
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// C++ ABI of the reference (legacy default stream, exceptions instead of error codes).
// ------------------------------------------------------------------------------------------------
namespace CudaRasterizer {

static void raise_on_error(int rc, const char* where)
{
	if (rc >= 0)
		return;
	std::string msg = std::string(where) + " failed (" + std::to_string(rc) + ")";
	if (rc == GM_ERR_CUDA)
		msg += std::string(": ") + gm_last_error();
	throw std::runtime_error(msg);
}

GeometryState GeometryState::fromChunk(char*& chunk, size_t P)
{
	GeometryState s;
	s.begin = chunk;
	gm::GeometryState::fromChunk(chunk, P);
	s.end = chunk;
	return s;
}

ImageState ImageState::fromChunk(char*& chunk, size_t N)
{
	ImageState s;
	s.begin = chunk;
	gm::ImageState::fromChunk(chunk, N);
	s.end = chunk;
	return s;
}

BinningState BinningState::fromChunk(char*& chunk, size_t R)
{
	BinningState s;
	s.begin = chunk;
	gm::BinningState::fromChunk(chunk, R);
	s.end = chunk;
	return s;
}

void Rasterizer::markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
	raise_on_error(gm_mark_visible(P, means3D, viewmatrix, projmatrix, reinterpret_cast<uint8_t*>(present), nullptr),
	               "markVisible");
}

int Rasterizer::forward_0(char* geometryBuffer, const int P, int D, int M, const float* background,
	const int width, int height, const float* means3D, const float* shs, const float* colors_precomp,
	const float* opacities, const float* scales, const float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	const float tan_fovx, float tan_fovy, const bool prefiltered, int* radii, bool debug)
{
	int rc = gm_forward_0(geometryBuffer, P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
	                      scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
	                      tan_fovy, prefiltered, radii, debug, nullptr);
	raise_on_error(rc, "forward_0");
	return rc;
}

void Rasterizer::forward_1(char* geometryBuffer, char* binningBuffer, char* imageBuffer, const int P, int D, int M,
	int num_rendered, const float* background, const int width, int height, const float* means3D,
	const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
	const float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
	const float* projmatrix, const float* cam_pos, const float tan_fovx, float tan_fovy, const bool prefiltered,
	float* out_color, int* radii, bool debug)
{
	raise_on_error(gm_forward_1(geometryBuffer, binningBuffer, imageBuffer, P, D, M, num_rendered, background, width,
	                            height, means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
	                            cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered,
	                            out_color, radii, debug, nullptr),
	               "forward_1");
}

int Rasterizer::forward(std::function<char* (size_t)> geometryBuffer, std::function<char* (size_t)> binningBuffer,
	std::function<char* (size_t)> imageBuffer, const int P, int D, int M, const float* background,
	const int width, int height, const float* means3D, const float* shs, const float* colors_precomp,
	const float* opacities, const float* scales, const float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	const float tan_fovx, float tan_fovy, const bool prefiltered, float* out_color, int* radii, bool debug)
{
	char* geom = geometryBuffer(required<GeometryState>(P));
	char* img = imageBuffer(required<ImageState>((size_t)width * height));
	int R = forward_0(geom, P, D, M, background, width, height, means3D, shs, colors_precomp, opacities, scales,
	                  scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
	                  prefiltered, radii, debug);
	char* binning = binningBuffer(required<BinningState>(R));
	forward_1(geom, binning, img, P, D, M, R, background, width, height, means3D, shs, colors_precomp, opacities,
	          scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
	          prefiltered, out_color, radii, debug);
	return R;
}

void Rasterizer::backward(const int P, int D, int M, int R, const float* background, const int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
	const float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
	const float* projmatrix, const float* campos, const float tan_fovx, float tan_fovy, const int* radii,
	char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix, float* dL_dmean2D,
	float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
	float* dL_dscale, float* dL_drot, bool debug)
{
	raise_on_error(gm_backward(P, D, M, R, background, width, height, means3D, shs, colors_precomp, scales,
	                           scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx,
	                           tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dmean2D,
	                           dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot,
	                           debug, nullptr),
	               "backward");
}

} // namespace CudaRasterizer
