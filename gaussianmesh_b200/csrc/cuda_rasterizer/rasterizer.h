// Drop-in replacement header for the reference's cuda_rasterizer/rasterizer.h.
//
// The reference's Jittor glue (dgr/rasterize_points.py:8-10, 12-37) compiles inline C++ with
// -I<package>/cuda_rasterizer and links -lCudaRasterizer, calling the static methods declared
// here.  libCudaRasterizer.so (sm_100a build) exports exactly these C++ symbols as thin wrappers
// over the flat C ABI of include/gm_rasterizer.h, so the glue works unchanged.  Argument order,
// types and meaning follow dgr/cuda_rasterizer/rasterizer.h:20-133.
#ifndef CUDA_RASTERIZER_H_INCLUDED
#define CUDA_RASTERIZER_H_INCLUDED

#include <cstddef>
#include <functional>
#include <vector>

namespace CudaRasterizer
{
	class Rasterizer
	{
	public:
		static void markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present);

		// preprocess + per-tile instance count; returns the instance total (blocking 4-byte read)
		static int forward_0(char* geometryBuffer, const int P, int D, int M, const float* background,
			const int width, int height, const float* means3D, const float* shs,
			const float* colors_precomp, const float* opacities, const float* scales,
			const float scale_modifier, const float* rotations, const float* cov3D_precomp,
			const float* viewmatrix, const float* projmatrix, const float* cam_pos,
			const float tan_fovx, float tan_fovy, const bool prefiltered, int* radii,
			bool debug = false);

		// emit + per-tile sort + blend
		static void forward_1(char* geometryBuffer, char* binningBuffer, char* imageBuffer,
			const int P, int D, int M, int num_rendered, const float* background,
			const int width, int height, const float* means3D, const float* shs,
			const float* colors_precomp, const float* opacities, const float* scales,
			const float scale_modifier, const float* rotations, const float* cov3D_precomp,
			const float* viewmatrix, const float* projmatrix, const float* cam_pos,
			const float tan_fovx, float tan_fovy, const bool prefiltered, float* out_color,
			int* radii, bool debug);

		static int forward(std::function<char* (size_t)> geometryBuffer,
			std::function<char* (size_t)> binningBuffer, std::function<char* (size_t)> imageBuffer,
			const int P, int D, int M, const float* background, const int width, int height,
			const float* means3D, const float* shs, const float* colors_precomp,
			const float* opacities, const float* scales, const float scale_modifier,
			const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
			const float* projmatrix, const float* cam_pos, const float tan_fovx, float tan_fovy,
			const bool prefiltered, float* out_color, int* radii = nullptr, bool debug = false);

		static void backward(const int P, int D, int M, int R, const float* background,
			const int width, int height, const float* means3D, const float* shs,
			const float* colors_precomp, const float* scales, const float scale_modifier,
			const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
			const float* projmatrix, const float* campos, const float tan_fovx, float tan_fovy,
			const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
			const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
			float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
			float* dL_drot, bool debug);
	};
};

#endif
