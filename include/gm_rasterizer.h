/*
 * gm_rasterizer.h -- flat C ABI of libCudaRasterizer.so (B200 / sm_100a build).
 *
 * This is the drop-in boundary for the one hot path this repository replaces:
 *   mesh-face bind/deform -> per-Gaussian preprocess -> tile binning/sort -> per-tile alpha blend
 *   (forward and backward).
 * Every entry takes plain device pointers, sizes and a CUDA stream; none allocates or frees
 * device memory (the caller owns every byte, exactly as with the reference library).
 * All pointers are DEVICE pointers unless the parameter name ends in _host.
 *
 * Citations are relative to the reference checkout
 * (gaussian_renderer/diff_gaussian_rasterizater/ abbreviated as dgr/).
 *
 * Return value of every int-returning entry: GM_OK (0) or a negative GM_ERR_* code.  With
 * debug != 0 each stage is followed by a stream synchronise + error check (the reference's
 * CHECK_CUDA, dgr/cuda_rasterizer/auxiliary.h:165-172); with debug == 0 only launch-time
 * errors are reported, as in the reference.
 */
#ifndef GM_RASTERIZER_H_INCLUDED
#define GM_RASTERIZER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_OK 0
#define GM_ERR_CUDA (-1)            /* a CUDA runtime call or kernel failed (see gm_last_error)      */
#define GM_ERR_BAD_ARGUMENT (-2)    /* inconsistent pointers / sizes (e.g. neither SH nor colours)    */
#define GM_ERR_TOO_MANY_TILES (-3)  /* ceil(W/16)*ceil(H/16) exceeds GM_MAX_TILES                     */
#define GM_ERR_BINNING_OVERFLOW (-4)/* binning chunk smaller than the instance count of this view     */

#define GM_TILE 16                  /* BLOCK_X == BLOCK_Y, dgr/cuda_rasterizer/config.h:16-17         */
#define GM_MAX_TILES (1 << 18)      /* per-tile counters live in the geometry chunk       */

typedef void* gm_stream_t;          /* a cudaStream_t; NULL = the legacy default stream               */

/* ---- library info ------------------------------------------------------------------------ */
const char* gm_version(void);       /* "gaussianmesh-b200 <semver> sm_100a"                           */
const char* gm_last_error(void);    /* text of the last GM_ERR_CUDA on this host thread               */

/* ---- optional per-stage timing.  Between gm_profile_begin() and gm_profile_end() every kernel stage
 *      launched through this library is bracketed by two CUDA events on its launch stream;
 *      gm_profile_end() waits for them and returns, per stage, the summed elapsed milliseconds and
 *      the number of launches (arrays of gm_profile_num_stages() entries; either may be NULL).
 *      Off by default: no events are recorded and nothing is added to the stream. ----------------- */
int gm_profile_num_stages(void);
const char* gm_profile_stage_name(int stage);
void gm_profile_begin(void);
int gm_profile_end(float* ms_per_stage, uint64_t* launches_per_stage);

/* ---- opaque chunk sizing; replaces CudaRasterizer::required<T>()
 *      (dgr/cuda_rasterizer/rasterizer_impl.h:67-73, used at dgr/rasterize_points.py:74-76,186) --- */
size_t gm_required_geom(size_t P);          /* GeometryState for P Gaussians                        */
size_t gm_required_image(size_t N);         /* ImageState for N = W*H pixels                        */
size_t gm_required_binning(size_t R);       /* BinningState for R (Gaussian, tile) instances        */
/*  Inverse of gm_required_binning: the instance capacity gm_forward derives from a binning chunk of
 *  `bytes` bytes.  This is the R to hand to gm_backward after a gm_forward over that chunk. */
size_t gm_binning_capacity(size_t bytes);

/* ---- markVisible; replaces Rasterizer::markVisible (dgr/cuda_rasterizer/rasterizer.h:24-29,
 *      rasterizer_impl.cu:54-66,141-153).  present[P] is a byte mask (C++ bool). ---------------- */
int gm_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, gm_stream_t stream);

/* ---- two-phase forward; replaces Rasterizer::forward_0 / forward_1
 *      (dgr/cuda_rasterizer/rasterizer.h:31-76, rasterizer_impl.cu:338-513), the pair the
 *      reference's glue calls (dgr/rasterize_points.py:164-266).
 *
 *  gm_forward_0: preprocess + per-tile instance count + tile offsets.  Returns the number of
 *  (Gaussian, tile) instances the caller must size the binning chunk for (>= 0), or GM_ERR_*.
 *  Like the reference it ends with a blocking 4-byte device->host read.
 *  Null-pointer variants as in the reference (rasterizer_impl.cu:364-367, rasterize_points.py:
 *  162-163): colors_precomp == NULL -> SH path (shs [P,M,3], degree D); cov3D_precomp == NULL ->
 *  scales [P,3] / rotations [P,4] path (quaternion used UN-normalised, forward.cu:127);
 *  radii == NULL -> radii kept only inside the geometry chunk. */
int gm_forward_0(char* geom_buffer, int P, int D, int M, const float* background, int width,
                 int height, const float* means3D, const float* shs, const float* colors_precomp,
                 const float* opacities, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                 int prefiltered, int* radii, int debug, gm_stream_t stream);

/*  gm_forward_1: instance emit + per-tile depth sort + record packing + front-to-back blend.
 *  num_rendered is the value gm_forward_0 returned; binning_buffer must hold
 *  gm_required_binning(num_rendered) bytes, image_buffer gm_required_image(W*H).
 *  out_color is planar [3,H,W] (forward.cu:368-373). */
int gm_forward_1(char* geom_buffer, char* binning_buffer, char* image_buffer, int P, int D, int M,
                 int num_rendered, const float* background, int width, int height,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* opacities, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                 int prefiltered, float* out_color, int* radii, int debug, gm_stream_t stream);

/* ---- single-call forward; replaces Rasterizer::forward (rasterizer.h:79-100,
 *      rasterizer_impl.cu:198-336) WITHOUT its host synchronisation: the caller hands in a
 *      binning chunk of binning_capacity bytes sized from a high-water mark; the frame's
 *      counters are written asynchronously to frame_info_host[4] (pinned host memory, may be
 *      NULL): [0] instances this view needs, [1] visible Gaussians, [2] overflow flag,
 *      [3] capacity in instances.  If the view needs more than the chunk holds, the frame is
 *      rendered from the instances that fit (tiles past the capacity come out as background),
 *      [0] still receives the full requirement and [2] is 1; gm_forward_status() on the geometry
 *      chunk then returns GM_ERR_BINNING_OVERFLOW. */
int gm_forward(char* geom_buffer, char* binning_buffer, size_t binning_capacity,
               char* image_buffer, int P, int D, int M, const float* background, int width,
               int height, const float* means3D, const float* shs, const float* colors_precomp,
               const float* opacities, const float* scales, float scale_modifier,
               const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
               const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
               int prefiltered, float* out_color, int* radii, int debug,
               uint32_t* frame_info_host, gm_stream_t stream);

/*  gm_forward with work folded into the epilogue of the blend kernel (the training step's next two launches):
 *    target / loss / dL_dimg   L1 loss against `target` ([3,H,W]; float32, or uint8 when target_is_u8 -- value / 255.0f):
 *                              *loss = mean |image - target| (zeroed by this call, accumulated per tile), dL_dimg = its
 *                              gradient sign(image - target) / (3 H W) -- what gm_l1_loss computes (utils/loss_utils.py:17-18);
 *    zero_ptr / zero_floats    a float buffer (16-byte aligned, a multiple of 4 floats) that is cleared: the accumulated
 *                              gradient buffers gm_backward adds into (rasterize_points.py:302-305 allocates them zeroed).
 *  Every field may be NULL / 0.  The blend kernel is issue-bound with the memory system idle, so both ride for free. */
typedef struct gm_forward_epilogue {
	const void* target;
	int target_is_u8;
	float* loss;
	float* dL_dimg;
	float* zero_ptr;
	size_t zero_floats;
} gm_forward_epilogue;
int gm_forward_ex(char* geom_buffer, char* binning_buffer, size_t binning_capacity,
                  char* image_buffer, int P, int D, int M, const float* background, int width,
                  int height, const float* means3D, const float* shs, const float* colors_precomp,
                  const float* opacities, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                  int prefiltered, float* out_color, int* radii, int debug,
                  uint32_t* frame_info_host, const gm_forward_epilogue* epilogue, gm_stream_t stream);

/*  Blocking: synchronises `stream`, reads the frame header of a geometry chunk and reports
 *  GM_OK / GM_ERR_BINNING_OVERFLOW; *num_rendered / *num_visible may be NULL. */
int gm_forward_status(const char* geom_buffer, int* num_rendered, int* num_visible,
                      gm_stream_t stream);

/*  Diagnostic: device pointers into a geometry chunk carved for P Gaussians, for tests that compare
 *  per-Gaussian state with the reference's GeometryState (rasterizer_impl.h:21-39).  out[0] depths
 *  f32[P], out[1] means2D f32[P,2], out[2] cov3D f32[P,6], out[3] conic_opacity f32[P,4],
 *  out[4] rgb+clamp-bits f32[P,4], out[5] per-tile instance counts u32[tiles]. */
void gm_geom_view(char* geom_buffer, size_t P, void** out6);

/* ---- backward; replaces Rasterizer::backward (rasterizer.h:102-132,
 *      rasterizer_impl.cu:515-608).  The three chunks must be those of the matching forward,
 *      unmodified.  All nine gradient buffers MUST be zero-initialised by the caller
 *      (dgr/rasterize_points.py:302-310): dL_dmean2D [P,3] (NDC units, .z unused),
 *      dL_dconic [P,4] (.x=a, .y=b, .w=c, .z unused; backward.cu:549-551), dL_dopacity [P],
 *      dL_dcolor [P,3], dL_dmean3D [P,3], dL_dcov3D [P,6], dL_dsh [P,M,3], dL_dscale [P,3],
 *      dL_drot [P,4] (gradient w.r.t. the UN-normalised quaternion, backward.cu:340). */
int gm_backward(int P, int D, int M, int R, const float* background, int width, int height,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* scales, float scale_modifier, const float* rotations,
                const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                char* geom_buffer, char* binning_buffer, char* image_buffer,
                const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                float* dL_dscale, float* dL_drot, int debug, gm_stream_t stream);

/*  gm_backward with flags.  GM_BACKWARD_OVERWRITE: dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale and dL_drot need
 *  NOT be pre-zeroed -- the call writes every element of them (zeros for Gaussians that were not rendered and
 *  for SH coefficients above the active degree), which saves the caller a 256 B/Gaussian memset.  The four
 *  accumulated buffers (dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor) must still be zero. */
#define GM_BACKWARD_OVERWRITE 1
int gm_backward_ex(int P, int D, int M, int R, const float* background, int width, int height,
                   const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                   float scale_modifier, const float* rotations, const float* cov3D_precomp,
                   const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                   float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                   const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                   float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                   float* dL_dscale, float* dL_drot, int debug, int flags, gm_stream_t stream);

/* ---- mesh-bound Gaussian parametrisation (training side); replaces the Jittor op chains of
 *      MeshBasedGaussianModel.get_xyz / get_scaling / get_rotation / get_opacity
 *      (scene/mesh_based_gaussian_model.py:34-43,122-152):
 *        xyz     = softmax(bc) . (v1,v2,v3) + alpha_distance * r * (sigmoid(distance) - 0.5) * normal
 *        scale   = exp(log_scale);  rot = q / max(|q|, 1e-12);  opacity = sigmoid(opacity_logit)
 *      Any of the activation input/output pairs may be NULL to skip that activation. --------- */
int gm_mesh_bind_forward(int P, const float* bc_logits /*[P,3]*/, const float* distance /*[P,1]*/,
                         const float* vertex1, const float* vertex2, const float* vertex3 /*[P,3]*/,
                         const float* normal /*[P,3]*/, const float* r /*[P,1]*/,
                         float alpha_distance,
                         const float* log_scale /*[P,3]*/, const float* rot_raw /*[P,4]*/,
                         const float* opacity_logit /*[P,1]*/,
                         float* xyz /*[P,3]*/, float* scale /*[P,3]*/, float* rot /*[P,4]*/,
                         float* opacity /*[P,1]*/, gm_stream_t stream);

/*  Vector-Jacobian product of the above: given dL/dxyz, dL/dscale, dL/drot, dL/dopacity (any may
 *  be NULL) writes (not accumulates) dL/dbc_logits, dL/ddistance, dL/dlog_scale, dL/drot_raw,
 *  dL/dopacity_logit. */
int gm_mesh_bind_backward(int P, const float* bc_logits, const float* distance,
                          const float* vertex1, const float* vertex2, const float* vertex3,
                          const float* normal, const float* r, float alpha_distance,
                          const float* log_scale, const float* rot_raw, const float* opacity_logit,
                          const float* dL_dxyz, const float* dL_dscale, const float* dL_drot,
                          const float* dL_dopacity,
                          float* dL_dbc_logits, float* dL_ddistance, float* dL_dlog_scale,
                          float* dL_drot_raw, float* dL_dopacity_logit, gm_stream_t stream);

/* ---- edit-time per-face local-frame transform; replaces SingleObjectDeform.deform_gaussian
 *      (edittool/__init__.py:103-131) and the per-frame strip_symmetric
 *      (edittool/general_utils.py:26-37):
 *        dpos_g = sum_k w_k (V' - V)[tri_k];  R_g = (sum_k w_k R[tri_k])^T;  S_g = sum_k w_k S[tri_k]
 *        A = R_g S_g;  Sigma' = A Sigma A^T;  pos' = pos + dpos_g
 *      cov_in is either packed [P,6] (cov_in_is_full == 0) or full [P,3,3] (== 1).
 *      Outputs: pos_out [P,3], cov6_out [P,6] (packed xx,xy,xz,yy,yz,zz), rot_out [P,3,3] (= R_g). */
int gm_deform_gaussians(int P, int num_vertices, const float* vertex_rest /*[Vn,3]*/,
                        const float* vertex_deformed /*[Vn,3]*/, const float* vertex_R /*[Vn,3,3]*/,
                        const float* vertex_S /*[Vn,3,3]*/, const int* gaussian_triangles /*[P,3]*/,
                        const float* weights /*[P,3]*/, const float* pos_in /*[P,3]*/,
                        const float* cov_in, int cov_in_is_full,
                        float* pos_out, float* cov6_out, float* rot_out, gm_stream_t stream);

/* ---- edit-time per-frame colour; replaces the Jittor chain of
 *      ObjectVisualTool.render_gaussian (edittool/__init__.py:442-448) + eval_sh
 *      (edittool/sh_utils.py:34-89):  dir = normalize(pos - campos); dir' = R_g^T dir;
 *      rgb = max(eval_sh(D, shs, dir') + 0.5, 0).   shs is [P,M,3]; rot may be NULL (identity). */
int gm_sh_to_rgb_rotated(int P, int D, int M, const float* pos /*[P,3]*/, const float* campos /*[3]*/,
                         const float* rot /*[P,3,3] or NULL*/, const float* shs,
                         float* rgb /*[P,3]*/, gm_stream_t stream);

/* ---- backward of gm_sh_to_rgb_rotated: dL/dshs [P,M,3] (zero above degree D) and dL/dpos [P,3] through the
 *      normalised (and, with rot, rotated) view direction; either output may be NULL.  With rot == NULL this is the
 *      autograd of the reference's convert_SHs_python branch (gaussian_renderer/__init__.py:87-92,
 *      utils/sh_utils.py:57-112), clamp included. */
int gm_sh_to_rgb_rotated_backward(int P, int D, int M, const float* pos, const float* campos, const float* rot,
                                  const float* shs, const float* dL_drgb /*[P,3]*/, float* dL_dshs, float* dL_dpos,
                                  gm_stream_t stream);

/* ---- the compute_cov3D_python branch (gaussian_renderer/__init__.py:78-79): pc.get_covariance(scaling_modifier) =
 *      strip_symmetric(L L^T), L = R(q / |q|) diag(modifier * s) (utils/general_utils.py:64-109,
 *      scene/mesh_based_gaussian_model.py:24-29,176-177).  The quaternion is normalised here, unlike the CUDA
 *      branch (forward.cu:127).  cov6 is [P,6] packed xx,xy,xz,yy,yz,zz.  The backward takes dL/dcov6 as the
 *      rasterizer returns it (six independent inputs) and writes dL/dscale [P,3], dL/drot [P,4]. */
int gm_cov3d_from_scale_rot(int P, const float* scales, float scale_modifier, const float* rotations, float* cov6,
                            gm_stream_t stream);
int gm_cov3d_from_scale_rot_backward(int P, const float* scales, float scale_modifier, const float* rotations,
                                     const float* dL_dcov6, float* dL_dscale, float* dL_drot, gm_stream_t stream);

/* ---- SingleObjectDeform.load_mesh, face-id branch (edittool/__init__.py:87-101): gaussian_triangles[i] =
 *      faces[face_id[i]] and the area-ratio barycentric weights of the projected point
 *      (get_barycentric_coordinate, edittool/general_utils.py:73-88), float64 like numpy.
 *      vertex [Vn,3] float64, faces [Fn,3] int32, face_id [P] int64, proj_pos [P,3] float32 (get_proj_xyz);
 *      outputs gaussian_triangles [P,3] int32, weights [P,3] float64. */
int gm_load_mesh(int P, int num_vertices, int num_faces, const double* vertex, const int32_t* faces,
                 const int64_t* face_id, const float* proj_pos, int32_t* gaussian_triangles, double* weights,
                 gm_stream_t stream);

/* ---- ACAP rotation / shear per vertex; replaces pyACAP.pyACAP(mesh).GetRS(ref_V, def_V, _R = 1, ncpu)
 *      (edittool/__init__.py:102,109; ACAP/pyACAPv1.zip: mainpy.cpp:60-64, src/FeatureVector.cpp:81-173,428-590,
 *      src/Align.cpp:31-100).  float64 arithmetic like the reference.
 *
 *  gm_acap_build_rings (HOST arrays): cyclically ordered one-ring neighbours of every vertex and the vertex ->
 *  incident-face lists of a triangle mesh.  ring_neighbours_host needs 3*num_faces + num_vertices entries,
 *  face_list_host 3*num_faces; the two offset arrays num_vertices + 1.
 *  gm_acap_rest (device, once per rest mesh): per ring entry the factor applied to the edge vector (fourth root of
 *  the cotangent weight), the unit vertex normals and per vertex (sum p p^T)^-1 (9 doubles, row-major).
 *  gm_acap_get_rs (device, per deformed mesh): R_out[Vn,3,3] (= the TRANSPOSE of the polar rotation, as pyACAP
 *  returns it) and S_out[Vn,3,3], float32 row-major, ready for gm_deform_gaussians. */
int gm_acap_build_rings(int num_vertices, int num_faces, const int32_t* faces_host /*[Fn,3]*/,
                        int32_t* ring_offsets_host, int32_t* ring_neighbours_host,
                        int32_t* face_offsets_host, int32_t* face_list_host);
int gm_acap_rest(int num_vertices, const double* vertex_rest /*[Vn,3]*/, const int32_t* faces /*[Fn,3]*/,
                 const int32_t* ring_offsets, const int32_t* ring_neighbours, const int32_t* face_offsets,
                 const int32_t* face_list, double* sqrt_w /*[ring entries]*/, double* rest_normals /*[Vn,3]*/,
                 double* ata_inv /*[Vn,9]*/, gm_stream_t stream);
int gm_acap_get_rs(int num_vertices, const double* vertex_rest, const double* vertex_deformed /*[Vn,3]*/,
                   const int32_t* faces, const int32_t* ring_offsets, const int32_t* ring_neighbours,
                   const int32_t* face_offsets, const int32_t* face_list, const double* sqrt_w,
                   const double* rest_normals, const double* ata_inv, double* normals_scratch /*[Vn,3]*/,
                   float* R_out /*[Vn,3,3]*/, float* S_out /*[Vn,3,3]*/, gm_stream_t stream);

/* ---- L1 loss of config 4 (utils/loss_utils.py:17-18): writes mean|img-target| to *loss and
 *      dL/dimg = sign(img-target)/numel to dL_dimg (may be NULL).  numel = 3*W*H. ------------- */
int gm_l1_loss(size_t numel, const float* img, const float* target, float* loss /*[1]*/,
               float* dL_dimg, gm_stream_t stream);

/*  Ground-truth images stay 8-bit in the reference's loaders until utils/general_utils.py:22-27 (PILtoTorch) divides by
 *  255.0; keeping the target as uint8 quarters the per-step upload.  gm_l1_loss_u8 is gm_l1_loss with target[i] =
 *  (float)u8[i] / 255.0f evaluated inside the kernel; gm_image_u8_to_float writes that float image (for the losses that
 *  take a float target). */
int gm_l1_loss_u8(size_t numel, const float* img, const uint8_t* target, float* loss, float* dL_dimg, gm_stream_t stream);
int gm_image_u8_to_float(size_t numel, const uint8_t* src, float* dst, gm_stream_t stream);

/* ---- the rest of one training iteration around the op (train_mesh_gaussian.py:85-147) -------
 *
 *  gm_photometric_loss: out[0] = (1 - lambda) L1 + lambda (1 - SSIM), out[1] = L1 = mean|img - gt|,
 *  out[2] = SSIM(img, gt) (utils/loss_utils.py:17-18,36-82: 11x11 Gaussian window, sigma 1.5, zero padding, per
 *  channel, mean over all C*H*W elements; train_mesh_gaussian.py:91,94).  If dL_dimg != NULL it receives
 *  d out[0] / d img ([C,H,W]).  `scratch` is a caller-owned chunk of gm_photometric_scratch_bytes(C,H,W) bytes.
 *  Two kernel launches, no host synchronisation. */
size_t gm_photometric_scratch_bytes(int C, int H, int W);
int gm_photometric_loss(int C, int H, int W, const float* img, const float* gt, float lambda_dssim, char* scratch,
                        float* out /*[3]*/, float* dL_dimg, gm_stream_t stream);

/*  gm_mesh_restrict_loss (utils/loss_utils.py:84-107, train_mesh_gaussian.py:93):
 *  *loss = sum_i max(0, max_k scale[i,k] - weight * sqrt(|(v2-v1) x (v3-v1)|)).  If dL_dscale != NULL it receives the
 *  gradient w.r.t. scale ([P,3]; 1 at the first maximal component of an active Gaussian) -- written when
 *  accumulate == 0, added when accumulate != 0 (on top of the rasterizer's dL/dscale). */
int gm_mesh_restrict_loss(int P, const float* scale /*[P,3]*/, const float* vertex1, const float* vertex2,
                          const float* vertex3 /*[P,3] each*/, float weight, float* loss /*[1]*/, float* dL_dscale,
                          int accumulate, gm_stream_t stream);

/*  gm_adam_step: one Adam update of every listed tensor in one pass (the reference's
 *  jt.nn.Adam(l, lr=0.0, eps=1e-15), scene/mesh_based_gaussian_model.py:248-258, stepped at
 *  train_mesh_gaussian.py:136-147):  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g g;
 *  p <- p - m * lr * sqrt(1 - b2^step) / (1 - b1^step) / (sqrt(v) + eps).   `tensors_host` is a HOST array; `step`
 *  is 1-based.  If period > 0, element i uses lr_head when (i mod period) < split and lr otherwise (one
 *  [P,16,3] feature tensor carrying the f_dc and f_rest learning rates). */
typedef struct gm_adam_tensor {
	float* param;
	const float* grad;
	float* exp_avg;
	float* exp_avg_sq;
	size_t numel;
	float lr;
	float lr_head;
	uint32_t period;
	uint32_t split;
} gm_adam_tensor;
int gm_adam_step(int num_tensors, const gm_adam_tensor* tensors_host, int step, float beta1, float beta2, float eps,
                 gm_stream_t stream);

/*  Device-side gate for the sync-free training loop: gm_forward never reads the instance count back
 *  (rasterizer_impl.cu:411 does), so a frame that outgrew its binning chunk is only known on the device when the
 *  optimizer runs.  gm_frame_overflow_flag returns the address of that frame's overflow word inside the geometry chunk;
 *  the _gated variants read it (NULL = ungated) and drop the update / the statistics when it is set -- a partially
 *  rendered frame never reaches the parameters.  The host learns of the overflow from gm_forward's frame_info_host. */
const uint32_t* gm_frame_overflow_flag(const char* geom_buffer);
int gm_adam_step_gated(int num_tensors, const gm_adam_tensor* tensors_host, int step, float beta1, float beta2, float eps,
                       const uint32_t* skip_flag, gm_stream_t stream);

/*  View-parallel training (no counterpart in the reference, which is single-GPU): every rank holds
 *  the whole model in ONE flat float vector of `total` floats (tensors at 32-float aligned offsets, described by
 *  `segments_host`), renders its own view, and leaves its gradient in a flat buffer of the same layout.
 *  gm_adam_step_sharded_p2p is the gradient exchange + update + parameter broadcast in one kernel over peer memory:
 *  rank r loads shard r (gm_adam_shard_range) of every rank's gradient buffer, averages, applies the Adam update of
 *  gm_adam_step with ITS shard of the moments (exp_avg / exp_avg_sq hold hi - lo floats) and stores the new parameters
 *  into every rank's parameter buffer.  grads_host / params_host are HOST arrays of `world` DEVICE pointers
 *  (index = rank; entries of other ranks are peer mappings of their buffers, e.g. CUDA IPC / symmetric memory).
 *  The caller orders the kernel between two cross-rank barriers (all gradients written before; all parameter
 *  stores landed after).  world <= GM_MAX_PEERS, num_segments <= 8. */
#define GM_MAX_PEERS 8
typedef struct gm_adam_segment {
	size_t offset;          /* first float of the tensor in the flat vector (multiple of 32) */
	size_t numel;
	float lr;
	float lr_head;
	uint32_t period;
	uint32_t split;
} gm_adam_segment;
void gm_adam_shard_range(size_t total, int world, int rank, size_t* lo, size_t* hi);
int gm_adam_step_sharded_p2p(int world, int rank, const float* const* grads_host, float* const* params_host,
                             int num_segments, const gm_adam_segment* segments_host, size_t total,
                             float* exp_avg, float* exp_avg_sq, int step, float beta1, float beta2, float eps,
                             gm_stream_t stream);

/*  The same exchange through a MULTICAST mapping of the two vectors (NVSwitch / NVLS; torch symmetric memory exposes
 *  it as handle.multicast_ptr): multimem.ld_reduce.add sums the N gradient copies inside the switch, multimem.st writes
 *  the new parameters into all N parameter vectors.  grads_multicast / params_multicast are the multicast addresses,
 *  params_local this rank's own (unicast) parameter vector; everything else as gm_adam_step_sharded_p2p.  Per GPU the
 *  NVLink traffic drops from 2 (N-1)/N to about 2/N of the vector in each direction. */
int gm_adam_step_sharded_mc(int world, int rank, const float* grads_multicast, float* params_multicast,
                            const float* params_local, int num_segments, const gm_adam_segment* segments_host, size_t total,
                            float* exp_avg, float* exp_avg_sq, int step, float beta1, float beta2, float eps,
                            gm_stream_t stream);

/*  gm_densify_stats (train_mesh_gaussian.py:117-121, scene/mesh_based_gaussian_model.py:587-589): for every Gaussian
 *  with radii > 0:  max_radii2D = max(max_radii2D, radii);  grad_accum += |dL_dmean2D.xy|;  denom += 1. */
int gm_densify_stats(int P, const int32_t* radii, const float* dL_dmean2D /*[P,3]*/, float* max_radii2D /*[P]*/,
                     float* grad_accum /*[P]*/, float* denom /*[P]*/, gm_stream_t stream);
int gm_densify_stats_gated(int P, const int32_t* radii, const float* dL_dmean2D, float* max_radii2D, float* grad_accum,
                           float* denom, const uint32_t* skip_flag, gm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GM_RASTERIZER_H_INCLUDED */
