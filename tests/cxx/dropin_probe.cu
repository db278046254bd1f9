// TEST INFRASTRUCTURE: a stand-alone host program that uses libCudaRasterizer.so exactly the way the reference's
// Jittor glue does (gaussian_renderer/diff_gaussian_rasterizater/rasterize_points.py:63-274, 276-398): it includes
// the SHIPPED cuda_rasterizer/rasterizer_impl.h, sizes the three byte buffers with the header template
// CudaRasterizer::required<T>(), and calls CudaRasterizer::Rasterizer::{forward_0, forward_1, backward, markVisible}
// through their C++ (Itanium-mangled) symbols on the legacy default stream.  No gm_* entry is used.
//
//   dropin_probe <in.bin> <out.bin>
// in.bin : int32 P, D, M, W, H; float tanfovx, tanfovy, scale_modifier; bg[3]; view[16]; proj[16]; campos[3];
//          means3D[P*3]; shs[P*M*3]; opacities[P]; scales[P*3]; rotations[P*4]; dL_dpix[3*H*W]
// out.bin: int32 num_rendered; color[3*H*W]; radii[P] (int32); visible[P] (uint8);
//          dL_dmeans3D[P*3]; dL_dsh[P*M*3]; dL_dopacity[P]; dL_dscales[P*3]; dL_drot[P*4]; dL_dmeans2D[P*3]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "rasterizer_impl.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

template <typename T>
static T* upload(const std::vector<T>& h)
{
	T* d = nullptr;
	if (cudaMalloc(&d, h.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
	cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
	return d;
}

template <typename T>
static T* zeros(size_t n)
{
	T* d = nullptr;
	if (cudaMalloc(&d, n * sizeof(T) + 16) != cudaSuccess) return nullptr;
	cudaMemset(d, 0, n * sizeof(T));
	return d;
}

template <typename T>
static bool read_vec(FILE* f, std::vector<T>& v, size_t n)
{
	v.resize(n);
	return fread(v.data(), sizeof(T), n, f) == n;
}

template <typename T>
static void write_dev(FILE* f, const T* d, size_t n)
{
	std::vector<T> h(n);
	cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost);
	fwrite(h.data(), sizeof(T), n, f);
}

int main(int argc, char** argv)
{
	if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
	FILE* f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 1; }
	int hdr[5];
	float fl[3];
	if (fread(hdr, sizeof(int), 5, f) != 5 || fread(fl, sizeof(float), 3, f) != 3) return 1;
	const int P = hdr[0], D = hdr[1], M = hdr[2], W = hdr[3], H = hdr[4];
	std::vector<float> bg, view, proj, campos, means3D, shs, opac, scales, rots, dLdpix;
	if (!read_vec(f, bg, 3) || !read_vec(f, view, 16) || !read_vec(f, proj, 16) || !read_vec(f, campos, 3) ||
	    !read_vec(f, means3D, (size_t)P * 3) || !read_vec(f, shs, (size_t)P * M * 3) || !read_vec(f, opac, P) ||
	    !read_vec(f, scales, (size_t)P * 3) || !read_vec(f, rots, (size_t)P * 4) || !read_vec(f, dLdpix, (size_t)3 * H * W))
		return 1;
	fclose(f);

	float *d_bg = upload(bg), *d_view = upload(view), *d_proj = upload(proj), *d_cam = upload(campos);
	float *d_means = upload(means3D), *d_shs = upload(shs), *d_op = upload(opac), *d_sc = upload(scales), *d_rot = upload(rots);
	float* d_dLdpix = upload(dLdpix);

	// compute_buffer_size (rasterize_points.py:63-86)
	const size_t geom_size = CudaRasterizer::required<CudaRasterizer::GeometryState>(P);
	const size_t img_size = CudaRasterizer::required<CudaRasterizer::ImageState>((size_t)W * H);
	char* geom = zeros<char>(geom_size);
	char* image = zeros<char>(img_size);
	int* radii = zeros<int>(P);
	float* color = zeros<float>((size_t)3 * H * W);

	// RasterizeGaussiansCUDA, first snippet (:123-189)
	const int num_rendered = CudaRasterizer::Rasterizer::forward_0(geom, P, D, M, d_bg, W, H, d_means, d_shs, nullptr, d_op,
	                                                                d_sc, fl[2], d_rot, nullptr, d_view, d_proj, d_cam, fl[0],
	                                                                fl[1], false, radii, false);
	const size_t binning_size = CudaRasterizer::required<CudaRasterizer::BinningState>(num_rendered);
	char* binning = zeros<char>(binning_size);
	// second snippet (:196-268)
	CudaRasterizer::Rasterizer::forward_1(geom, binning, image, P, D, M, num_rendered, d_bg, W, H, d_means, d_shs, nullptr,
	                                      d_op, d_sc, fl[2], d_rot, nullptr, d_view, d_proj, d_cam, fl[0], fl[1], false, color,
	                                      radii, false);
	CK(cudaDeviceSynchronize());

	// RasterizeGaussiansBackwardCUDA (:276-398): nine zero-filled gradient tensors
	float *g_m2d = zeros<float>((size_t)P * 3), *g_conic = zeros<float>((size_t)P * 4), *g_op = zeros<float>(P),
	      *g_col = zeros<float>((size_t)P * 3), *g_m3d = zeros<float>((size_t)P * 3), *g_cov = zeros<float>((size_t)P * 6),
	      *g_sh = zeros<float>((size_t)P * M * 3), *g_sc = zeros<float>((size_t)P * 3), *g_rot = zeros<float>((size_t)P * 4);
	CudaRasterizer::Rasterizer::backward(P, D, M, num_rendered, d_bg, W, H, d_means, d_shs, nullptr, d_sc, fl[2], d_rot, nullptr,
	                                     d_view, d_proj, d_cam, fl[0], fl[1], radii, geom, binning, image, d_dLdpix, g_m2d,
	                                     g_conic, g_op, g_col, g_m3d, g_cov, g_sh, g_sc, g_rot, false);
	CK(cudaDeviceSynchronize());

	bool* present = zeros<bool>(P);
	CudaRasterizer::Rasterizer::markVisible(P, d_means, d_view, d_proj, present);
	CK(cudaDeviceSynchronize());

	FILE* o = fopen(argv[2], "wb");
	if (!o) { perror(argv[2]); return 1; }
	fwrite(&num_rendered, sizeof(int), 1, o);
	write_dev(o, color, (size_t)3 * H * W);
	write_dev(o, radii, P);
	write_dev(o, reinterpret_cast<unsigned char*>(present), P);
	write_dev(o, g_m3d, (size_t)P * 3);
	write_dev(o, g_sh, (size_t)P * M * 3);
	write_dev(o, g_op, P);
	write_dev(o, g_sc, (size_t)P * 3);
	write_dev(o, g_rot, (size_t)P * 4);
	write_dev(o, g_m2d, (size_t)P * 3);
	fclose(o);
	printf("dropin_probe: P=%d %dx%d num_rendered=%d geom=%zu image=%zu binning=%zu bytes\n", P, W, H, num_rendered, geom_size,
	       img_size, binning_size);
	return 0;
}
