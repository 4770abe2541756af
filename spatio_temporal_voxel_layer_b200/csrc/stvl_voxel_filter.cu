// stvl_voxel_filter.cu — the VOXEL pre-filter of MeasurementBuffer::BufferROSCloud on the device
// (reference measurement_buffer.cpp:153-164: pcl::VoxelGrid<PCLPointCloud2>, leaf = (float)voxel_size, filter field "z"
// in [min_obstacle_height, max_obstacle_height], setDownsampleAllData(false), min points per voxel).
//
// PCL bins the accepted points by floor(p * inverse_leaf), sorts (bin, point index) pairs and emits the float centroid of
// every bin that holds enough points.  Here: one kernel computes a 63-bit bin key per point, a STABLE radix sort
// (cub::DeviceRadixSort — library code, like PCL's std::sort upstream) orders the point indices by bin with ascending
// point index inside a bin, and one kernel walks each bin sequentially, adding the floats in that order — the order the
// CPU oracle defines (PCL's own order inside a bin is unspecified: std::sort is unstable).  The output cloud has one
// float4 slot per input point: bin heads carry the centroid, every other slot carries the sensor origin, which Mark
// skips (distance^2 < 0.0001, reference spatio_temporal_voxel_grid.cpp:289).
#include <cub/device/device_radix_sort.cuh>
#include <math_constants.h>

#include "stvl_device.cuh"
#include "stvl_voxel_filter.h"

namespace stvl {

namespace {
constexpr unsigned long long kNoBin = ~0ull;
struct VfTransform { int on; float m[12]; };

__device__ __forceinline__ void load_xyz(const uint8_t * data, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz,
                                         unsigned long long i, float & x, float & y, float & z, const VfTransform & tf)
{
  const uint8_t * b = data + i * step;
  uint32_t w[3];
  const uint32_t off[3] = {ox, oy, oz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const uint8_t * q = b + off[a];
    if ((reinterpret_cast<uintptr_t>(q) & 3u) == 0) w[a] = *reinterpret_cast<const uint32_t *>(q);
    else w[a] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
  }
  x = __uint_as_float(w[0]); y = __uint_as_float(w[1]); z = __uint_as_float(w[2]);
  if (tf.on) transform_point(tf.m, x, y, z);
}

__global__ void __launch_bounds__(256) vf_keys_kernel(const uint8_t * __restrict__ data, unsigned long long n, uint32_t step,
                                                      uint32_t ox, uint32_t oy, uint32_t oz, float inv_leaf, double zlo, double zhi,
                                                      unsigned long long * __restrict__ keys, uint32_t * __restrict__ vals, const VfTransform tf)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x, y, z;
  load_xyz(data, step, ox, oy, oz, i, x, y, z, tf);
  unsigned long long key = kNoBin;
  // PCL: skip non-finite points and points whose filter field is outside the limits (value > max || value < min)
  if (isfinite(x) && isfinite(y) && isfinite(z) && !((double)z < zlo || (double)z > zhi)) {
    const float bx = floorf(__fmul_rn(x, inv_leaf)), by = floorf(__fmul_rn(y, inv_leaf)), bz = floorf(__fmul_rn(z, inv_leaf));
    const float lim = 1048575.0f;
    if (fabsf(bx) < lim && fabsf(by) < lim && fabsf(bz) < lim)
      key = ((unsigned long long)(uint32_t)((int)bx + (1 << 20)) << 42) | ((unsigned long long)(uint32_t)((int)by + (1 << 20)) << 21) |
            (unsigned long long)(uint32_t)((int)bz + (1 << 20));
  }
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) vf_centroid_kernel(const uint8_t * __restrict__ data, unsigned long long n, uint32_t step,
                                                          uint32_t ox, uint32_t oy, uint32_t oz,
                                                          const unsigned long long * __restrict__ keys, const uint32_t * __restrict__ vals,
                                                          int min_points, float4 fill, float4 * __restrict__ out, const VfTransform tf)
{
  const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long key = keys[j];
  float4 r = fill;
  if (key != kNoBin && (j == 0 || keys[j - 1] != key)) {
    // head of a bin: add the member points in ascending point index (the sort is stable), in float like PCL
    float sx = 0.f, sy = 0.f, sz = 0.f;
    unsigned long long c = 0;
    for (unsigned long long m = j; m < n && keys[m] == key; ++m, ++c) {
      float x, y, z;
      load_xyz(data, step, ox, oy, oz, vals[m], x, y, z, tf);
      sx = __fadd_rn(sx, x); sy = __fadd_rn(sy, y); sz = __fadd_rn(sz, z);
    }
    if (c >= (unsigned long long)(min_points > 0 ? min_points : 0)) {
      const float cf = (float)c;
      r = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), 0.f);
    }
  }
  out[j] = r;
}
}  // namespace

size_t voxel_filter_temp_bytes(unsigned long long n)
{
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 63);
  return tmp;
}

cudaError_t launch_voxel_filter(const uint8_t * data, unsigned long long n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz,
                                float leaf, double zlo, double zhi, int min_points, const double origin[3], const float * transform,
                                unsigned long long * keys_a, unsigned long long * keys_b, uint32_t * vals_a, uint32_t * vals_b,
                                void * temp, size_t temp_bytes, float4 * out, cudaStream_t s)
{
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const float inv_leaf = 1.0f / leaf;   // PCL: inverse_leaf_size_ = 1 / leaf_size_ (float)
  VfTransform tf;
  tf.on = transform != nullptr;
  for (int i = 0; i < 12; ++i) tf.m[i] = transform ? transform[i] : 0.f;
  vf_keys_kernel<<<blocks, 256, 0, s>>>(data, n, step, ox, oy, oz, inv_leaf, zlo, zhi, keys_a, vals_a, tf);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, vals_a, vals_b, (int)n, 0, 63, s);
  if (e != cudaSuccess) return e;
  const float4 fill = make_float4((float)origin[0], (float)origin[1], (float)origin[2], 0.f);
  vf_centroid_kernel<<<blocks, 256, 0, s>>>(data, n, step, ox, oy, oz, keys_b, vals_b, min_points, fill, out, tf);
  return cudaGetLastError();
}

}  // namespace stvl
