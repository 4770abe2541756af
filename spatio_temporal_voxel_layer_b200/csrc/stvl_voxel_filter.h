// stvl_voxel_filter.h — device VOXEL pre-filter (see stvl_voxel_filter.cu)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace stvl {
size_t voxel_filter_scratch_bytes(unsigned long long n);
// out[0 .. *n_out) receives the centroids (compact, one float4 per emitted leaf); n_out is a DEVICE counter
cudaError_t launch_voxel_filter(const uint8_t * data, unsigned long long n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz,
                                float leaf, double zlo, double zhi, int min_points, const float * transform,
                                void * scratch, float4 * out, unsigned long long * n_out, int sm_count, cudaStream_t s);
}  // namespace stvl
