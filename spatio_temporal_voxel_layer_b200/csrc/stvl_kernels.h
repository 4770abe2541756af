// stvl_kernels.h — launch interface between the C-ABI host code (stvl_api.cu) and the kernels.
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>
#include "stvl_device.cuh"

namespace stvl {

// optional per-sweep lists (NULL = not wanted)
struct SweepOutputs {
  float * points;            // survivors' voxel centres, xyz float32 (GetOccupancyPointCloud)
  unsigned long long points_cap;
  int32_t * cleared;         // (ix,iy) per cleared voxel (cleared_cells)
  unsigned long long cleared_cap;
  int32_t * ambiguous;       // (ix,iy,iz) of libm-dependent razor-edge decisions
  unsigned long long ambiguous_cap;
};

struct CostmapDev {
  double origin_x, origin_y, resolution;
  double inv_resolution;     // 1 / resolution (rounded): shortcut of worldToMap's division, see map_cell
  uint32_t size_x, size_y;
  int32_t mark_threshold;
  uint8_t default_value, lethal;
  // bitmap output (sharded cycle): this rank's packed cell bitmap has size_y rows of bit_wpr words and covers the 32-cell
  // words [bit_w_lo, bit_w_lo + bit_wpr) of a costmap row
  int32_t bit_w_lo, bit_wpr;
};

// sharded cycle: the per-rank cell bitmaps the root expands into costmap bytes.  Rank r's bitmap is packed: size_y rows of
// (w_hi - w_lo) words covering the 32-cell words [w_lo, w_hi) of a costmap row — the x interval its shard can mark (bit
// mx & 31 of word mx >> 5 = cell (mx, my) is lethal).
constexpr int kMaxShardRanks = 16;
struct ExpandParams {
  const uint32_t * bits[kMaxShardRanks];
  int32_t w_lo[kMaxShardRanks], w_hi[kMaxShardRanks];   // word interval per row that rank r owns (and sends)
  int32_t n_ranks;
  int32_t pad;
};

// updateCosts on device byte grids (stvl_costs.cu)
constexpr int kMaxPolygonVertices = 64;
struct PolygonParams {
  double xy[2 * kMaxPolygonVertices];
  uint32_t n;
  uint32_t pad;
};
cudaError_t launch_polygon_cost(uint8_t * costmap_dev, const CostmapDev & p, const PolygonParams & poly, uint8_t cost, int32_t * colmin,
                                int32_t * colmax, cudaStream_t s);
cudaError_t launch_update_costs(const uint8_t * layer_dev, uint8_t * master_dev, uint32_t size_x, int min_i, int min_j, int max_i, int max_j,
                                int method, int sm_count, cudaStream_t s);

cudaError_t launch_mark(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl = false,
                        const CostmapDev * fused_costmap = nullptr, uint8_t * costmap_dev = nullptr, bool costmap_bits = false,
                        uint32_t * cursor_host = nullptr);   // cursor_host: where GridView::mark_cursor stands (advanced by the launch)
uint32_t mark_blocks_for(unsigned long long n_points);
uint32_t mark_grid_for(const MarkBatch & batch, int sm_count);
uint32_t mark_costmap_tiles_capacity(const MarkBatch & batch, int sm_count);
cudaError_t launch_load_voxels(const GridView & g, const int32_t * x, const int32_t * y, const int32_t * z, const double * t,
                               unsigned long long n, cudaStream_t s);
cudaError_t launch_sweep(const GridView & g, const DevFrustum * frustums_dev, const FrustumParams * fparam, int n_frustums, double now,
                         const SweepOutputs & out, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s, bool pdl, bool share_sms,
                         uint8_t * costmap_reset = nullptr, size_t costmap_bytes = 0, int costmap_value = 0);
cudaError_t launch_costmap(const GridView & g, const CostmapDev & p, uint8_t * costmap_dev, uint32_t tiles_upper_bound, int sm_count, cudaStream_t s,
                           bool fill_default, bool pdl = false, bool bits = false);
cudaError_t launch_expand_bitmaps(const ExpandParams & e, const CostmapDev & p, uint8_t * costmap_dev, int sm_count, cudaStream_t s);
cudaError_t launch_rebuild_free_ring(const GridView & g, int sm_count, cudaStream_t s);
cudaError_t launch_columns(const GridView & g, int32_t * ix, int32_t * iy, uint32_t * cnt, unsigned long long cap,
                           uint32_t tiles_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_columns_to_dense(const GridView & g, int ix0, int iy0, uint32_t nx, uint32_t ny, uint32_t * dense,
                                    uint32_t tiles_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_costmap_from_dense(const GridView & g, const uint32_t * dense, int ix0, int iy0, uint32_t nx, uint32_t ny,
                                      const CostmapDev & p, uint8_t * costmap_dev, int sm_count, cudaStream_t s);
cudaError_t launch_columns_to_window(const GridView & g, int ix0, int iy0, uint32_t nx, uint32_t ny, void * win, int count_bytes,
                                     uint32_t tiles_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_costmap_from_window(const GridView & g, const void * win, int count_bytes, int ix0, int iy0, uint32_t nx, uint32_t ny,
                                       const CostmapDev & p, uint8_t * costmap_dev, int sm_count, cudaStream_t s);
cudaError_t launch_fill_where_zero(uint8_t * p, size_t n, uint8_t v, int sm_count, cudaStream_t s);
cudaError_t launch_reset_area(const GridView & g, double sx, double sy, double ex, double ey, int invert,
                              uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_count_active(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_dump_voxels(const GridView & g, int32_t * x, int32_t * y, int32_t * z, double * t, unsigned long long cap,
                               uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_rebuild_hash(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s);
cudaError_t launch_remap_tile_counts(const GridView & g, const uint64_t * old_keys, const uint32_t * old_counts, uint32_t n_old, int sm_count,
                                     cudaStream_t s);
cudaError_t launch_reassign_tiles(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s);


// launch with (optionally) the programmatic-stream-serialization attribute, see pdl_wait()
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_ex(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool pdl, Args &&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args &&>(args))...);
}

}  // namespace stvl
