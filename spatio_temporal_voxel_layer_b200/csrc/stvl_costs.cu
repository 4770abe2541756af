// stvl_costs.cu — SpatioTemporalVoxelLayer::updateCosts (reference spatio_temporal_voxel_layer.cpp:666-696) on
// device-resident byte grids: the footprint clearing (nav2 Costmap2D::setConvexPolygonCost) on the layer's costmap and
// CostmapLayer::updateWithOverwrite / updateWithMax of the bounds window into the master grid.
//
// nav2_costmap_2d is not vendored by the reference (package.xml dependency, version unpinned): the routines are
// restated from its published sources; the tests compare with a literal CPU restatement (sort + column scan).
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

#include "stvl_device.cuh"
#include "stvl_kernels.h"

namespace stvl {

namespace {
// nav2 Costmap2D::worldToMap: wx >= origin, (unsigned)((wx - origin) / resolution) < size
__device__ __forceinline__ bool costs_world_to_map(double wx, double wy, const CostmapDev & p, unsigned & mx, unsigned & my)
{
  if (wx < p.origin_x || wy < p.origin_y) return false;
  mx = __double2uint_rz(__ddiv_rn(__dsub_rn(wx, p.origin_x), p.resolution));
  my = __double2uint_rz(__ddiv_rn(__dsub_rn(wy, p.origin_y), p.resolution));
  return mx < p.size_x && my < p.size_y;
}
}  // namespace

// One CTA.  setConvexPolygonCost = map the vertices (any vertex off the map: nothing happens), walk the outline with
// nav2's Bresenham (raytraceLine / bresenham2D: edge i runs from vertex i to vertex i+1, the last one closes the
// polygon), then fill every column between its lowest and highest outline cell (convexFillCells sorts the outline
// cells by x and scans them; per-column min / max of y is the same set because a closed outline puts at least two
// cells into every column it spans).
__global__ void __launch_bounds__(256)
polygon_cost_kernel(uint8_t * __restrict__ costmap, const CostmapDev p, const __grid_constant__ PolygonParams poly, const uint8_t cost,
                    int32_t * __restrict__ colmin, int32_t * __restrict__ colmax)
{
  __shared__ unsigned s_mx[kMaxPolygonVertices], s_my[kMaxPolygonVertices];
  __shared__ int s_fail;
  __shared__ unsigned s_x0, s_x1;
  const unsigned tid = threadIdx.x, n = poly.n;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  if (tid < n) {
    unsigned mx = 0, my = 0;
    if (!costs_world_to_map(poly.xy[2 * tid], poly.xy[2 * tid + 1], p, mx, my)) atomicExch(&s_fail, 1);
    s_mx[tid] = mx; s_my[tid] = my;
  }
  __syncthreads();
  if (s_fail || n < 3) return;   // nav2: returns false before touching the map / convexFillCells needs a triangle
  if (tid == 0) {
    unsigned x0 = s_mx[0], x1 = s_mx[0];
    for (unsigned i = 1; i < n; ++i) { x0 = min(x0, s_mx[i]); x1 = max(x1, s_mx[i]); }
    s_x0 = x0; s_x1 = x1;
  }
  __syncthreads();
  const unsigned x0 = s_x0, x1 = s_x1;
  for (unsigned x = x0 + tid; x <= x1; x += blockDim.x) { colmin[x] = INT_MAX; colmax[x] = -1; }
  __syncthreads();
  if (tid < n) {
    // raytraceLine(at, x0, y0, x1, y1) with max_length = UINT_MAX, min_length = 0, then bresenham2D
    const unsigned ax = s_mx[tid], ay = s_my[tid], bx = s_mx[(tid + 1) % n], by = s_my[(tid + 1) % n];
    const int dx = (int)bx - (int)ax, dy = (int)by - (int)ay;
    const unsigned abs_dx = (unsigned)abs(dx), abs_dy = (unsigned)abs(dy);
    const int offset_dx = dx > 0 ? 1 : -1, offset_dy = (dy > 0 ? 1 : -1) * (int)p.size_x;
    unsigned offset = ay * p.size_x + ax;
    const bool x_dominant = abs_dx >= abs_dy;
    const unsigned abs_da = x_dominant ? abs_dx : abs_dy, abs_db = x_dominant ? abs_dy : abs_dx;
    const int offset_a = x_dominant ? offset_dx : offset_dy, offset_b = x_dominant ? offset_dy : offset_dx;
    int error_b = (int)(abs_da / 2);
    for (unsigned i = 0; i <= abs_da; ++i) {
      const unsigned cx = offset % p.size_x, cy = offset / p.size_x;
      atomicMin(&colmin[cx], (int)cy);
      atomicMax(&colmax[cx], (int)cy);
      if (i == abs_da) break;
      offset += offset_a;
      error_b += abs_db;
      if ((unsigned)error_b >= abs_da) { offset += offset_b; error_b -= abs_da; }
    }
  }
  __syncthreads();
  const unsigned warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
  for (unsigned x = x0 + warp; x <= x1; x += n_warps) {
    const int lo = colmin[x], hi = colmax[x];
    for (int y = lo + (int)lane; y <= hi; y += 32) costmap[(size_t)y * p.size_x + x] = cost;
  }
}

// CostmapLayer::updateWithOverwrite (method 0) / updateWithMax (method 1) over the window; 255 = NO_INFORMATION.
// Four cells per 32-bit word with the byte-wise SIMD intrinsics:
//   take = (layer != 255) && (method 0 || master == 255 || master < layer);   master = take ? layer : master
__device__ __forceinline__ uint32_t combine4(uint32_t c, uint32_t m, int method)
{
  const uint32_t known = ~__vcmpeq4(c, 0xFFFFFFFFu);
  const uint32_t take = method == 0 ? known : (known & (__vcmpeq4(m, 0xFFFFFFFFu) | __vcmpgtu4(c, m)));
  return (m & ~take) | (c & take);
}
// One thread per 16-byte aligned chunk of a window row (a row of the window starts and ends anywhere: the first and the
// last chunk of a row are done byte by byte).
__global__ void __launch_bounds__(256)
update_costs_kernel(const uint8_t * __restrict__ layer, uint8_t * __restrict__ master, const uint32_t size_x, const int min_i, const int min_j,
                    const int max_i, const int max_j, const int method, const unsigned chunks_per_row)
{
  const unsigned long long total = (unsigned long long)chunks_per_row * (unsigned)(max_j - min_j);
  const bool aligned = ((reinterpret_cast<uintptr_t>(layer) | reinterpret_cast<uintptr_t>(master)) & 15u) == 0 && (size_x & 15u) == 0;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    const size_t row = (size_t)(min_j + (int)(t / chunks_per_row)) * size_x;
    const int c0 = (min_i & ~15) + (int)(t % chunks_per_row) * 16;   // chunk [c0, c0 + 16) of the row
    const int lo = max(c0, min_i), hi = min(c0 + 16, max_i);
    if (aligned && lo == c0 && hi == c0 + 16) {
      const uint4 c = *reinterpret_cast<const uint4 *>(layer + row + c0);
      if ((c.x & c.y & c.z & c.w) == 0xFFFFFFFFu) continue;   // nothing known in this chunk
      uint4 m = *reinterpret_cast<const uint4 *>(master + row + c0);
      m.x = combine4(c.x, m.x, method); m.y = combine4(c.y, m.y, method); m.z = combine4(c.z, m.z, method); m.w = combine4(c.w, m.w, method);
      *reinterpret_cast<uint4 *>(master + row + c0) = m;
    } else {
      for (int i = lo; i < hi; ++i) {
        const uint8_t c = layer[row + i];
        if (c == 255) continue;
        if (method == 0) { master[row + i] = c; continue; }
        const uint8_t old_cost = master[row + i];
        if (old_cost == 255 || old_cost < c) master[row + i] = c;
      }
    }
  }
}

cudaError_t launch_polygon_cost(uint8_t * costmap_dev, const CostmapDev & p, const PolygonParams & poly, uint8_t cost, int32_t * colmin,
                                int32_t * colmax, cudaStream_t s)
{
  polygon_cost_kernel<<<1, 256, 0, s>>>(costmap_dev, p, poly, cost, colmin, colmax);
  return cudaGetLastError();
}
cudaError_t launch_update_costs(const uint8_t * layer_dev, uint8_t * master_dev, uint32_t size_x, int min_i, int min_j, int max_i, int max_j,
                                int method, int sm_count, cudaStream_t s)
{
  if (max_i <= min_i || max_j <= min_j) return cudaSuccess;
  const unsigned chunks_per_row = (unsigned)((((max_i + 15) & ~15) - (min_i & ~15)) / 16);
  const unsigned long long total = (unsigned long long)chunks_per_row * (unsigned long long)(max_j - min_j);
  unsigned long long blocks = (total + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (unsigned long long)sm_count * 16) blocks = (unsigned long long)sm_count * 16;
  update_costs_kernel<<<(unsigned)blocks, 256, 0, s>>>(layer_dev, master_dev, size_x, min_i, min_j, max_i, max_j, method, chunks_per_row);
  return cudaGetLastError();
}

}  // namespace stvl
