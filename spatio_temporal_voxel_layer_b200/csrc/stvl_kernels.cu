// stvl_kernels.cu — sm_100a kernels of the per-cycle voxel update.
//
//   mark_kernel<kVec4, kCostmap>   Mark / operator()            reference spatio_temporal_voxel_grid.cpp:254-309
//                                  (kCostmap: also carries the costmap pass of the fused cycle)
//   sweep_triage_kernel            ClearFrustums sweep, pass 1  reference .cpp:96-224: one thread per leaf, temporal / frustum
//                                                               culling at leaf level, column counts of untouched leaves
//   sweep_leaf_kernel              ClearFrustums sweep, pass 2  per-voxel decay + frustum IsInside for the queued leaves
//   costmap_kernel                 UpdateROSCostmap grid loop   reference spatio_temporal_voxel_layer.cpp:699-718
//   reset_area_kernel              ResetGridArea                reference .cpp:397-418
//   + small utility kernels (column compaction, dumps, hash rebuild, dense scatter for the multi-GPU merge)
//
// The three kernels of a cycle are chained by programmatic dependent launch (pdl_wait / pdl_launch_dependents below).
//
// Everything is integer / FP64 scalar work bound by HBM traffic: no tensor cores.  All
// arithmetic whose result feeds a comparison or a stored value uses the _rn intrinsics
// (never contracted into FMA) in the reference's operation order, see SURVEY.md §7.
#include <math_constants.h>
#include <algorithm>
#include <utility>

#include "stvl_common.cuh"

namespace stvl {

#ifdef STVL_TRACE
__device__ unsigned long long g_trace[4][1024][8];
extern "C" int stvl_debug_trace(void * out, size_t bytes)
{
  if (bytes != sizeof(g_trace)) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)) != cudaSuccess) return -3;
  return 0;
}
#endif

// block reduction of the marked-cell count / bounds (warp shuffle, shared memory, then ONE set of global atomics per CTA:
// a thousand warps updating the same five words would serialise in L2); the last CTA of the grid publishes.
// s_cm: five ints of shared memory, initialised {0, INT_MAX, INT_MAX, INT_MIN, INT_MIN} before the CTA's previous barrier.
__device__ __forceinline__ void costmap_reduce_and_publish(Counters * c, int * s_cm, unsigned n, int mnx, int mny, int mxx, int mxy)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n += __shfl_xor_sync(kFull, n, d);
    mnx = min(mnx, __shfl_xor_sync(kFull, mnx, d)); mny = min(mny, __shfl_xor_sync(kFull, mny, d));
    mxx = max(mxx, __shfl_xor_sync(kFull, mxx, d)); mxy = max(mxy, __shfl_xor_sync(kFull, mxy, d));
  }
  if ((threadIdx.x & 31) == 0 && n) {
    atomicAdd(&s_cm[0], (int)n);
    atomicMin(&s_cm[1], mnx); atomicMin(&s_cm[2], mny); atomicMax(&s_cm[3], mxx); atomicMax(&s_cm[4], mxy);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_cm[0]) {
      atomicAdd(&c->cm_acc.n_marked, (unsigned long long)s_cm[0]);
      atomicMin(&c->cm_acc.mk_min_x, s_cm[1]); atomicMin(&c->cm_acc.mk_min_y, s_cm[2]);
      atomicMax(&c->cm_acc.mk_max_x, s_cm[3]); atomicMax(&c->cm_acc.mk_max_y, s_cm[4]);
    }
    __threadfence();
    if (atomicAdd(&c->cm_ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      costmap_publish_result(c);
      c->cm_ticket = 0;
      __threadfence();
    }
  }
}

// =====================================================================================
// Costmap kernel (helpers: see above the Mark kernel, which can carry the costmap pass in the fused cycle)
// =====================================================================================
constexpr int kCostmapTilesPerWarp = 8;
template <bool kBits>
__global__ void __launch_bounds__(256) costmap_kernel(const GridView g, const CostmapDev p, uint8_t * __restrict__ costmap)
{
  const unsigned lane = threadIdx.x & 31;
  __shared__ int s_cm[5];
  if (threadIdx.x == 0) { s_cm[0] = 0; s_cm[1] = s_cm[2] = INT_MAX; s_cm[3] = s_cm[4] = INT_MIN; }
  TRACE(3, 0);
  // (fused cycle with a small Mark grid: launched programmatically between the leaf pass and Mark — the column counts are
  // ready when the sweep is, and Mark's cloud streaming overlaps this kernel as well; no-ops otherwise)
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t n_tiles = min(g.ctr->tile_n, g.tile_capacity);
  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  TRACE(3, 1);
  unsigned n = 0;
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
  // each warp owns batches of 8 consecutive tiles; the 8 x 256 B count rows and the 8 tile keys are requested together
  // (one DRAM round trip per batch instead of one per tile)
  for (uint32_t t0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kCostmapTilesPerWarp; t0 < n_tiles;
       t0 += warps_total * kCostmapTilesPerWarp) {
    costmap_tile_batch<kCostmapTilesPerWarp, kBits>(g, p, costmap, t0, n_tiles, lane, n, mnx, mny, mxx, mxy);
  }
  __syncthreads();
  TRACE(3, 2);
  costmap_reduce_and_publish(g.ctr, s_cm, n, mnx, mny, mxx, mxy);
  TRACE(3, 3);
}

__global__ void columns_prologue_kernel(const GridView g) { g.ctr->n_columns_out = 0; }

// occupied columns of the last sweep as (ix, iy, count) — GetFlattenedCostmap, reference .cpp:312-317
__global__ void __launch_bounds__(256) columns_kernel(const GridView g, int32_t * __restrict__ oix, int32_t * __restrict__ oiy,
                                                      uint32_t * __restrict__ ocnt, unsigned long long cap)
{
  const uint32_t n_tiles = min(g.ctr->tile_n, g.tile_capacity);
  const unsigned long long total = (unsigned long long)n_tiles * 64;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t cnt = g.tile_count[i];
    if (!cnt) continue;
    const unsigned long long o = atomicAdd(&g.ctr->n_columns_out, 1ull);
    if (o < cap) {
      int bx, by;
      unpack_tile_key(g.tile_key[i >> 6], bx, by);
      const int col = (int)(i & 63);
      if (oix) oix[o] = bx * 8 + (col >> 3);
      if (oiy) oiy[o] = by * 8 + (col & 7);
      if (ocnt) ocnt[o] = cnt;
    }
  }
}

// multi-GPU merge: scatter-add this shard's column counts into a dense window (then reduced over NVLink)
__global__ void __launch_bounds__(256) columns_to_dense_kernel(const GridView g, int ix0, int iy0, uint32_t nx, uint32_t ny, uint32_t * __restrict__ dense)
{
  const uint32_t n_tiles = min(g.ctr->tile_n, g.tile_capacity);
  const unsigned long long total = (unsigned long long)n_tiles * 64;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t cnt = g.tile_count[i];
    if (!cnt) continue;
    int bx, by;
    unpack_tile_key(g.tile_key[i >> 6], bx, by);
    const int col = (int)(i & 63);
    const long long dx = (long long)(bx * 8 + (col >> 3)) - ix0, dy = (long long)(by * 8 + (col & 7)) - iy0;
    if (dx >= 0 && dy >= 0 && dx < (long long)nx && dy < (long long)ny) atomicAdd(&dense[(size_t)dy * nx + dx], cnt);
  }
}

__global__ void __launch_bounds__(256) costmap_from_dense_kernel(const GridView g, const uint32_t * __restrict__ dense, int ix0, int iy0,
                                                                 uint32_t nx, uint32_t ny, const CostmapDev p, uint8_t * __restrict__ costmap)
{
  const unsigned long long total = (unsigned long long)nx * ny;
  __shared__ int s_cm[5];
  if (threadIdx.x == 0) { s_cm[0] = 0; s_cm[1] = s_cm[2] = INT_MAX; s_cm[3] = s_cm[4] = INT_MIN; }
  unsigned n = 0;
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t cnt = dense[i];
    if (cnt == 0 || (int)cnt < p.mark_threshold) continue;
    const int ix = ix0 + (int)(i % nx), iy = iy0 + (int)(i / nx);
    unsigned mx, my;
    if (world_to_map(index_to_world(ix, g.vs, g.half_vs), index_to_world(iy, g.vs, g.half_vs), p, mx, my)) {
      costmap[(size_t)my * p.size_x + mx] = p.lethal;
      ++n; mnx = min(mnx, ix); mny = min(mny, iy); mxx = max(mxx, ix); mxy = max(mxy, iy);
    }
  }
  __syncthreads();
  costmap_reduce_and_publish(g.ctr, s_cm, n, mnx, mny, mxx, mxy);
}

// typed variants for the in-library sharded merge (u8 windows move 4x fewer bytes over NVLink than u32 ones; a column
// lives in exactly one tile of one rank's shard, so the scatter is a plain saturating store)
template <typename T>
__global__ void __launch_bounds__(256) columns_to_window_kernel(const GridView g, int ix0, int iy0, uint32_t nx, uint32_t ny, T * __restrict__ win)
{
  const uint32_t n_tiles = min(g.ctr->tile_n, g.tile_capacity);
  const unsigned long long total = (unsigned long long)n_tiles * 64;
  const uint32_t cap = sizeof(T) == 1 ? 255u : 0xFFFFFFFFu;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t cnt = g.tile_count[i];
    if (!cnt) continue;
    int bx, by;
    unpack_tile_key(g.tile_key[i >> 6], bx, by);
    const int col = (int)(i & 63);
    const long long dx = (long long)(bx * 8 + (col >> 3)) - ix0, dy = (long long)(by * 8 + (col & 7)) - iy0;
    if (dx >= 0 && dy >= 0 && dx < (long long)nx && dy < (long long)ny) win[(size_t)dy * nx + dx] = (T)(cnt < cap ? cnt : cap);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) costmap_from_window_kernel(const GridView g, const T * __restrict__ win, int ix0, int iy0,
                                                                  uint32_t nx, uint32_t ny, const CostmapDev p, uint8_t * __restrict__ costmap)
{
  const unsigned long long total = (unsigned long long)nx * ny;
  __shared__ int s_cm[5];
  if (threadIdx.x == 0) { s_cm[0] = 0; s_cm[1] = s_cm[2] = INT_MAX; s_cm[3] = s_cm[4] = INT_MIN; }
  unsigned n = 0;
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t cnt = win[i];
    if (cnt == 0 || (int)cnt < p.mark_threshold) continue;
    const int ix = ix0 + (int)(i % nx), iy = iy0 + (int)(i / nx);
    unsigned mx, my;
    if (world_to_map(index_to_world(ix, g.vs, g.half_vs), index_to_world(iy, g.vs, g.half_vs), p, mx, my)) {
      costmap[(size_t)my * p.size_x + mx] = p.lethal;
      ++n; mnx = min(mnx, ix); mny = min(mny, iy); mxx = max(mxx, ix); mxy = max(mxy, iy);
    }
  }
  __syncthreads();
  costmap_reduce_and_publish(g.ctr, s_cm, n, mnx, mny, mxx, mxy);
}

// =====================================================================================
// ResetGridArea — reference .cpp:397-418 (x,y only: whole voxel columns are cleared)
// =====================================================================================
__global__ void __launch_bounds__(256) reset_area_kernel(const GridView g, double sx, double sy, double ex, double ey, int invert)
{
  const unsigned lane = threadIdx.x & 31;
  const uint32_t n_slots = slots_in_use(g);
  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  for (uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < n_slots; slot += warps_total) {
    const uint64_t key = g.leaf_key[slot];
    if (key == kEmptyKey) continue;
    int bx, by, bz;
    unpack_leaf_key(key, bx, by, bz);
    uint16_t * m = reinterpret_cast<uint16_t *>(g.leaf_mask + (size_t)slot * 16) + lane;
    uint32_t m16 = *m;
    if (!m16) continue;
    const uint32_t before = m16;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int col = 2 * (int)lane + k;
      const double wx = index_to_world(bx * 8 + (col >> 3), g.vs, g.half_vs), wy = index_to_world(by * 8 + (col & 7), g.vs, g.half_vs);
      const bool in_x = wx > sx && wx < ex, in_y = wy > sy && wy < ey;
      const bool in_range = in_x && in_y;
      if (in_range == (invert != 0)) m16 &= k ? 0x00FFu : 0xFF00u;
    }
    if (m16 != before) *m = (uint16_t)m16;
  }
}

// =====================================================================================
// dumps, counts, maintenance
// =====================================================================================
__global__ void __launch_bounds__(256) count_active_kernel(const GridView g)
{
  const uint32_t n_slots = slots_in_use(g);
  unsigned long long n = 0;
  unsigned live = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned long long)n_slots * 16; i += (unsigned long long)gridDim.x * blockDim.x) {
    if (g.leaf_key[i >> 4] == kEmptyKey) continue;
    n += __popc(g.leaf_mask[i]);
    if ((i & 15) == 0) ++live;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { n += __shfl_xor_sync(kFull, n, d); live += __shfl_xor_sync(kFull, live, d); }
  if ((threadIdx.x & 31) == 0) { if (n) atomicAdd(&g.ctr->n_active, n); if (live) atomicAdd(&g.ctr->live_leaves, live); }
}

__global__ void __launch_bounds__(256) dump_voxels_kernel(const GridView g, int32_t * __restrict__ ox, int32_t * __restrict__ oy,
                                                          int32_t * __restrict__ oz, double * __restrict__ ot, unsigned long long cap)
{
  const unsigned lane = threadIdx.x & 31;
  const uint32_t n_slots = slots_in_use(g);
  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  for (uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < n_slots; slot += warps_total) {
    const uint64_t key = g.leaf_key[slot];
    if (key == kEmptyKey) continue;
    int bx, by, bz;
    unpack_leaf_key(key, bx, by, bz);
    uint32_t m = reinterpret_cast<const uint16_t *>(g.leaf_mask + (size_t)slot * 16)[lane];
    const int cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFull, incl, d); if ((int)lane >= d) incl += v; }
    const int total = __shfl_sync(kFull, incl, 31);
    if (!total) continue;
    unsigned long long b0 = 0;
    if (lane == 0) b0 = atomicAdd(&g.ctr->n_voxels_out, (unsigned long long)total);
    b0 = __shfl_sync(kFull, b0, 0);
    unsigned long long o = b0 + (incl - cnt);
    while (m) {
      const int b = __ffs(m) - 1; m &= m - 1;
      const unsigned li = lane * 16 + b;
      if (o < cap) {
        if (ox) ox[o] = bx * 8 + (int)(li >> 6);
        if (oy) oy[o] = by * 8 + (int)((li >> 3) & 7);
        if (oz) oz[o] = bz * 8 + (int)(li & 7);
        if (ot) ot[o] = g.leaf_time[(size_t)slot * kLeafVoxels + li];
      }
      ++o;
    }
  }
}

// re-insert every live leaf into a freshly cleared hash (drops tombstones / overflow markers)
__global__ void __launch_bounds__(256) rebuild_hash_kernel(const GridView g)
{
  const uint32_t n_slots = slots_in_use(g);
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += gridDim.x * blockDim.x) {
    const uint64_t key = g.leaf_key[slot];
    if (key == kEmptyKey) continue;
    uint32_t h = (uint32_t)mix64(key) & g.hash_mask;
    for (;;) {
      const uint64_t old = atomicCAS(&g.hash[h].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) { g.hash[h].val = slot; g.leaf_hpos[slot] = h; break; }
      h = (h + 1) & g.hash_mask;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) g.ctr->n_tombs = 0;
}

// after the tile pool was rebuilt from scratch: give every live leaf its tile again
__global__ void __launch_bounds__(256) reassign_tiles_kernel(const GridView g)
{
  const uint32_t n_slots = slots_in_use(g);
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += gridDim.x * blockDim.x) {
    const uint64_t key = g.leaf_key[slot];
    if (key == kEmptyKey) continue;
    int bx, by, bz;
    unpack_leaf_key(key, bx, by, bz);
    g.leaf_tile[slot] = tile_find_or_insert(g, bx, by);
  }
}

// after a tile rebuild renumbered the tiles: carry the last sweep's per-column counts (the reference's _cost_map) over from a
// copy of the old tile keys / count rows, so that a costmap query between the rebuild and the next sweep still sees them
__global__ void __launch_bounds__(256) remap_tile_counts_kernel(const GridView g, const uint64_t * __restrict__ old_keys,
                                                                const uint32_t * __restrict__ old_counts, uint32_t n_old)
{
  const unsigned lane = threadIdx.x & 31;
  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  for (uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_old; t += warps_total) {
    const uint2 c = reinterpret_cast<const uint2 *>(old_counts + (size_t)t * 64)[lane];
    if (!__any_sync(kFull, (c.x | c.y) != 0)) continue;
    uint32_t nt = 0;
    if (lane == 0) { int bx, by; unpack_tile_key(old_keys[t], bx, by); nt = tile_find_or_insert(g, bx, by); }
    nt = __shfl_sync(kFull, nt, 0);
    if (nt != kOverflowSlot) reinterpret_cast<uint2 *>(g.tile_count + (size_t)nt * 64)[lane] = c;
  }
}

// =====================================================================================
// host-callable launchers
// =====================================================================================
static inline int grid_for(unsigned long long work_items, int per_block, int max_blocks)
{
  unsigned long long b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > (unsigned long long)max_blocks) b = max_blocks;
  return (int)b;
}

// bits: costmap_dev is this rank's packed cell bitmap (size_y rows of p.bit_wpr words) instead of bytes
cudaError_t launch_costmap(const GridView & g, const CostmapDev & p, uint8_t * costmap_dev, uint32_t tiles_upper_bound, int sm_count, cudaStream_t s,
                           bool fill_default, bool pdl, bool bits)
{
  if (fill_default) {
    const size_t bytes = bits ? (size_t)p.bit_wpr * 4u * p.size_y : (size_t)p.size_x * p.size_y;
    const cudaError_t e = cudaMemsetAsync(costmap_dev, bits ? 0 : p.default_value, bytes, s);   // Costmap2D::resetMaps, :705
    if (e != cudaSuccess) return e;
  }
  // 64 tiles per CTA (8 per warp): the per-CTA statistics reduction is amortised and the grid stays a fraction of a wave
  const unsigned grid = (unsigned)grid_for(tiles_upper_bound, 64, sm_count * 4);
  if (bits) return launch_ex(costmap_kernel<true>, grid, 256u, 0, s, pdl && !fill_default, g, p, costmap_dev);
  return launch_ex(costmap_kernel<false>, grid, 256u, 0, s, pdl && !fill_default, g, p, costmap_dev);
}

// Root rank of the sharded cycle: expand the ranks' cell bitmaps into costmap bytes (lethal where any rank marked the cell,
// the default value elsewhere — this IS Costmap2D::resetMaps + the grid loop of UpdateROSCostmap for the merged map).
// One thread per 16 cells (one 128-bit store).
__global__ void __launch_bounds__(256) expand_bitmaps_kernel(const __grid_constant__ ExpandParams e, const CostmapDev p, uint8_t * __restrict__ costmap)
{
  const uint32_t gpr = (p.size_x + 15u) >> 4;             // 16-cell groups per row
  const unsigned long long total = (unsigned long long)gpr * p.size_y;
  const uint32_t dflt = p.default_value, lethal = p.lethal;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / gpr), gx = (uint32_t)(i % gpr);
    const int w = (int)(gx >> 1);
    uint32_t bits = 0;
    for (int r = 0; r < e.n_ranks; ++r)
      if (w >= e.w_lo[r] && w < e.w_hi[r]) bits |= e.bits[r][(size_t)y * (uint32_t)(e.w_hi[r] - e.w_lo[r]) + (uint32_t)(w - e.w_lo[r])];
    bits = (gx & 1u) ? (bits >> 16) : (bits & 0xFFFFu);
    uint32_t out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) v |= (((bits >> (4 * q + b)) & 1u) ? lethal : dflt) << (8 * b);
      out[q] = v;
    }
    const size_t cell = (size_t)y * p.size_x + (size_t)gx * 16;
    const uint32_t n_valid = min(16u, p.size_x - gx * 16u);
    if (n_valid == 16 && ((reinterpret_cast<uintptr_t>(costmap) + cell) & 15u) == 0) {
      *reinterpret_cast<uint4 *>(costmap + cell) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
      for (uint32_t k = 0; k < n_valid; ++k) costmap[cell + k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
    }
  }
}
cudaError_t launch_expand_bitmaps(const ExpandParams & e, const CostmapDev & p, uint8_t * costmap_dev, int sm_count, cudaStream_t s)
{
  const unsigned long long groups = (unsigned long long)((p.size_x + 15u) >> 4) * p.size_y;
  if (groups == 0) return cudaSuccess;
  expand_bitmaps_kernel<<<grid_for(groups, 256, sm_count * 8), 256, 0, s>>>(e, p, costmap_dev);
  return cudaGetLastError();
}

// after the leaf pool was resized: rebuild the free ring from the slots below the bump pointer that hold no leaf
__global__ void __launch_bounds__(256) rebuild_free_ring_kernel(const GridView g)
{
  const uint32_t n_slots = slots_in_use(g);
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += gridDim.x * blockDim.x) {
    if (g.leaf_key[slot] != kEmptyKey) continue;
    const uint32_t i = atomicAdd(&g.ctr->free_push, 1u);
    g.free_slots[i & g.free_ring_mask] = slot;
  }
}
cudaError_t launch_rebuild_free_ring(const GridView & g, int sm_count, cudaStream_t s)
{
  // (the caller has reset free_pop / free_push to 0 and clamped bump to the old capacity on the stream before this)
  rebuild_free_ring_kernel<<<grid_for(g.leaf_capacity, 256, sm_count * 8), 256, 0, s>>>(g);
  return cudaGetLastError();
}
cudaError_t launch_columns(const GridView & g, int32_t * ix, int32_t * iy, uint32_t * cnt, unsigned long long cap,
                           uint32_t tiles_upper_bound, int sm_count, cudaStream_t s)
{
  columns_prologue_kernel<<<1, 1, 0, s>>>(g);
  columns_kernel<<<grid_for((unsigned long long)tiles_upper_bound * 64, 256, sm_count * 8), 256, 0, s>>>(g, ix, iy, cnt, cap);
  return cudaGetLastError();
}
cudaError_t launch_columns_to_dense(const GridView & g, int ix0, int iy0, uint32_t nx, uint32_t ny, uint32_t * dense,
                                    uint32_t tiles_upper_bound, int sm_count, cudaStream_t s)
{
  columns_to_dense_kernel<<<grid_for((unsigned long long)tiles_upper_bound * 64, 256, sm_count * 8), 256, 0, s>>>(g, ix0, iy0, nx, ny, dense);
  return cudaGetLastError();
}
cudaError_t launch_costmap_from_dense(const GridView & g, const uint32_t * dense, int ix0, int iy0, uint32_t nx, uint32_t ny,
                                      const CostmapDev & p, uint8_t * costmap_dev, int sm_count, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(costmap_dev, p.default_value, (size_t)p.size_x * p.size_y, s);
  if (e != cudaSuccess) return e;
  costmap_from_dense_kernel<<<grid_for((unsigned long long)nx * ny, 256, sm_count * 8), 256, 0, s>>>(g, dense, ix0, iy0, nx, ny, p, costmap_dev);
  return cudaGetLastError();
}
cudaError_t launch_columns_to_window(const GridView & g, int ix0, int iy0, uint32_t nx, uint32_t ny, void * win, int count_bytes,
                                     uint32_t tiles_upper_bound, int sm_count, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(win, 0, (size_t)nx * ny * count_bytes, s);
  if (e != cudaSuccess) return e;
  const int blocks = grid_for((unsigned long long)tiles_upper_bound * 64, 256, sm_count * 8);
  if (count_bytes == 1) columns_to_window_kernel<uint8_t><<<blocks, 256, 0, s>>>(g, ix0, iy0, nx, ny, static_cast<uint8_t *>(win));
  else columns_to_window_kernel<uint32_t><<<blocks, 256, 0, s>>>(g, ix0, iy0, nx, ny, static_cast<uint32_t *>(win));
  return cudaGetLastError();
}
cudaError_t launch_costmap_from_window(const GridView & g, const void * win, int count_bytes, int ix0, int iy0, uint32_t nx, uint32_t ny,
                                       const CostmapDev & p, uint8_t * costmap_dev, int sm_count, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(costmap_dev, p.default_value, (size_t)p.size_x * p.size_y, s);
  if (e != cudaSuccess) return e;
  const int blocks = grid_for((unsigned long long)nx * ny, 256 * 8, sm_count * 8);
  if (count_bytes == 1) costmap_from_window_kernel<uint8_t><<<blocks, 256, 0, s>>>(g, static_cast<const uint8_t *>(win), ix0, iy0, nx, ny, p, costmap_dev);
  else costmap_from_window_kernel<uint32_t><<<blocks, 256, 0, s>>>(g, static_cast<const uint32_t *>(win), ix0, iy0, nx, ny, p, costmap_dev);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256) fill_where_zero_kernel(uint8_t * __restrict__ p, size_t n, uint8_t v)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) if (p[i] == 0) p[i] = v;
}
cudaError_t launch_fill_where_zero(uint8_t * p, size_t n, uint8_t v, int sm_count, cudaStream_t s)
{
  fill_where_zero_kernel<<<grid_for(n, 256 * 16, sm_count * 8), 256, 0, s>>>(p, n, v);
  return cudaGetLastError();
}
cudaError_t launch_reset_area(const GridView & g, double sx, double sy, double ex, double ey, int invert,
                              uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s)
{
  reset_area_kernel<<<grid_for(leaf_slots_upper_bound, 8, sm_count * 8), 256, 0, s>>>(g, sx, sy, ex, ey, invert);
  return cudaGetLastError();
}
cudaError_t launch_count_active(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s)
{
  count_active_kernel<<<grid_for((unsigned long long)leaf_slots_upper_bound * 16, 256, sm_count * 8), 256, 0, s>>>(g);
  return cudaGetLastError();
}
cudaError_t launch_dump_voxels(const GridView & g, int32_t * x, int32_t * y, int32_t * z, double * t, unsigned long long cap,
                               uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s)
{
  dump_voxels_kernel<<<grid_for(leaf_slots_upper_bound, 8, sm_count * 8), 256, 0, s>>>(g, x, y, z, t, cap);
  return cudaGetLastError();
}
cudaError_t launch_rebuild_hash(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s)
{
  rebuild_hash_kernel<<<grid_for(leaf_slots_upper_bound, 256, sm_count * 8), 256, 0, s>>>(g);
  return cudaGetLastError();
}
cudaError_t launch_remap_tile_counts(const GridView & g, const uint64_t * old_keys, const uint32_t * old_counts, uint32_t n_old, int sm_count,
                                     cudaStream_t s)
{
  if (n_old == 0) return cudaSuccess;
  remap_tile_counts_kernel<<<grid_for(n_old, 8, sm_count * 8), 256, 0, s>>>(g, old_keys, old_counts, n_old);
  return cudaGetLastError();
}
cudaError_t launch_reassign_tiles(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s)
{
  reassign_tiles_kernel<<<grid_for(leaf_slots_upper_bound, 256, sm_count * 8), 256, 0, s>>>(g);
  return cudaGetLastError();
}

}  // namespace stvl
