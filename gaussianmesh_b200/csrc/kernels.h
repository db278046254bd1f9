// Host-side launch prototypes of the sm_100a kernels (one .cu per pipeline stage).
#pragma once
#include <cuda_runtime_api.h>
#include "state.h"

namespace gm {

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream);

int launch_depth_buckets(int P, const float* means3D, const ViewParams& vp, const GeometryState& g, cudaStream_t stream);

int launch_preprocess(int P, const float* means3D, const float* scales, const float* rotations,
                      const float* opacities, const float* shs, const float* cov3D_precomp,
                      const float* colors_precomp, const ViewParams& vp, int* radii,
                      const GeometryState& g, bool prefiltered, cudaStream_t stream);

int launch_tile_scan(int num_tiles, const GeometryState& g, uint32_t capacity, const ViewParams& vp, cudaStream_t stream);

// phase 0: count the instances of the large-rectangle Gaussians per (tile, bucket); phase 1: place them.
int launch_large_tiles(int phase, const int* radii, const GeometryState& g, const BinningState* b, uint32_t capacity,
                       const ViewParams& vp, cudaStream_t stream);

int launch_emit(int P, const int* radii, const GeometryState& g, const BinningState& b, uint32_t capacity,
                const ViewParams& vp, cudaStream_t stream);

int launch_sort_pack(int num_tiles, const GeometryState& g, const BinningState& b, uint32_t capacity,
                     const ViewParams& vp, cudaStream_t stream);

// `epilogue` may be NULL; *epilogue_done tells the caller whether the selected kernel folded it in (the A/B variants do not)
int launch_blend_forward(const GeometryState& g, const BinningState& b, const ImageState& img, uint32_t capacity,
                         const ViewParams& vp, float* out_color, const gm_forward_epilogue* epilogue, bool* epilogue_done,
                         cudaStream_t stream);

int launch_blend_backward(const GeometryState& g, const BinningState& b, const ImageState& img, uint32_t capacity,
                          const ViewParams& vp, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                          float* dL_dopacity, float* dL_dcolor, cudaStream_t stream);

int launch_geometry_backward(int P, const float* means3D, const int* radii, const float* shs, const float* scales,
                             const float* rotations, const float* cov3Ds, const ViewParams& vp,
                             const GeometryState& g, const float* dL_dmean2D, const float* dL_dconic,
                             const float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                             float* dL_dscale, float* dL_drot, bool overwrite, cudaStream_t stream);

int launch_mesh_bind_forward(int P, const float* bc_logits, const float* distance, const float* v1, const float* v2,
                             const float* v3, const float* normal, const float* r, float alpha_distance,
                             const float* log_scale, const float* rot_raw, const float* opacity_logit, float* xyz,
                             float* scale, float* rot, float* opacity, cudaStream_t stream);

int launch_mesh_bind_backward(int P, const float* bc_logits, const float* distance, const float* v1, const float* v2,
                              const float* v3, const float* normal, const float* r, float alpha_distance,
                              const float* log_scale, const float* rot_raw, const float* opacity_logit,
                              const float* dL_dxyz, const float* dL_dscale, const float* dL_drot,
                              const float* dL_dopacity, float* dL_dbc, float* dL_ddist, float* dL_dlog_scale,
                              float* dL_drot_raw, float* dL_dopacity_logit, cudaStream_t stream);

int launch_deform(int P, const float* V, const float* Vd, const float* VR, const float* VS, const int* tri,
                  const float* w, const float* pos_in, const float* cov_in, int cov_full, float* pos_out,
                  float* cov6_out, float* rot_out, cudaStream_t stream);

int launch_sh_rotated(int P, int D, int M, const float* pos, const float* campos, const float* rot, const float* shs,
                      float* rgb, cudaStream_t stream);

int launch_cov3d_python(int P, const float* scales, float mod, const float* rotations, float* cov6, cudaStream_t stream);
int launch_cov3d_python_backward(int P, const float* scales, float mod, const float* rotations, const float* dL_dcov6,
                                 float* dL_dscale, float* dL_drot, cudaStream_t stream);
int launch_sh_rotated_backward(int P, int D, int M, const float* pos, const float* campos, const float* rot, const float* shs,
                               const float* dL_drgb, float* dL_dshs, float* dL_dpos, cudaStream_t stream);
int launch_load_mesh(int P, int num_faces, const double* vertex, const int* faces, const long long* face_id,
                     const float* proj_pos, int* gaussian_triangles, double* weights, cudaStream_t stream);

int launch_acap_rest(int Vn, const double* V, const int* F, const int* ring_off, const int* ring, const int* face_off,
                     const int* face_list, double* sqrt_w, double* normals, double* ata_inv, cudaStream_t stream);

int launch_acap_get_rs(int Vn, const double* V0, const double* V1, const int* F, const int* ring_off, const int* ring,
                       const int* face_off, const int* face_list, const double* sqrt_w, const double* n0,
                       const double* ata_inv, double* n1_scratch, float* R_out, float* S_out, cudaStream_t stream);

int acap_build_rings_host(int Vn, int Fn, const int* F, int* ring_off, int* ring, int* face_off, int* face_list);

int launch_l1(size_t numel, const float* img, const void* target, int target_is_u8, float* loss, float* dL_dimg,
              cudaStream_t stream);
int launch_u8_to_float(size_t numel, const uint8_t* src, float* dst, cudaStream_t stream);

size_t photometric_scratch_bytes(int C, int H, int W);
int launch_photometric(int C, int H, int W, const float* img, const float* gt, float lambda, char* scratch, float* out,
                       float* dL_dimg, cudaStream_t stream);
int launch_mesh_restrict(int P, const float* scale, const float* v1, const float* v2, const float* v3, float weight,
                         float* loss, float* dL_dscale, int accumulate, cudaStream_t stream);
int launch_adam(int n, const gm_adam_tensor* tensors, int step, float beta1, float beta2, float eps,
                const uint32_t* skip_flag, cudaStream_t stream);
int launch_adam_sharded_p2p(int world, int rank, const float* const* grads, float* const* params, int num_segments,
                            const gm_adam_segment* segments, size_t total, float* exp_avg, float* exp_avg_sq, int step,
                            float beta1, float beta2, float eps, cudaStream_t stream);
int launch_adam_sharded_mc(int world, int rank, const float* grads_mc, float* params_mc, const float* params_local,
                           int num_segments, const gm_adam_segment* segments, size_t total, float* exp_avg, float* exp_avg_sq,
                           int step, float beta1, float beta2, float eps, cudaStream_t stream);
int launch_densify_stats(int P, const int* radii, const float* dL_dmean2D, float* max_radii2D, float* grad_accum,
                         float* denom, const uint32_t* skip_flag, cudaStream_t stream);

} // namespace gm
