// Drop-in replacement header for the reference's cuda_rasterizer/rasterizer_impl.h.
//
// The reference's glue instantiates CudaRasterizer::required<T>() from THIS header and calls the
// exported T::fromChunk (dgr/rasterize_points.py:74-76,186; dgr/cuda_rasterizer/
// rasterizer_impl.h:67-73).  The three state types are opaque to every caller, so only their
// names, the fromChunk signature and the sizing protocol are kept; the real layouts live in
// gaussianmesh_b200/csrc/state.h.
#pragma once

#include <cstddef>
#include <cstdint>
#include "rasterizer.h"

namespace CudaRasterizer
{
	// Extent of one carved chunk: [begin, end).  fromChunk advances `chunk` to `end`.
	struct GeometryState
	{
		char* begin;
		char* end;
		static GeometryState fromChunk(char*& chunk, size_t P);
	};

	struct ImageState
	{
		char* begin;
		char* end;
		static ImageState fromChunk(char*& chunk, size_t N);
	};

	struct BinningState
	{
		char* begin;
		char* end;
		static BinningState fromChunk(char*& chunk, size_t R);
	};

	template<typename T>
	size_t required(size_t n)
	{
		char* size = nullptr;
		T::fromChunk(size, n);
		return ((size_t)size) + 128;
	}
};
