// stvl_voxel_filter.h — device VOXEL pre-filter (see stvl_voxel_filter.cu)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace stvl {
size_t voxel_filter_temp_bytes(unsigned long long n);
cudaError_t launch_voxel_filter(const uint8_t * data, unsigned long long n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz,
                                float leaf, double zlo, double zhi, int min_points, const double origin[3], const float * transform,
                                unsigned long long * keys_a, unsigned long long * keys_b, uint32_t * vals_a, uint32_t * vals_b,
                                void * temp, size_t temp_bytes, float4 * out, cudaStream_t s);
}  // namespace stvl
