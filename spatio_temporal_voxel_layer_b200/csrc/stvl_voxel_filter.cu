// stvl_voxel_filter.cu — the VOXEL pre-filter of MeasurementBuffer::BufferROSCloud on the device
// (reference measurement_buffer.cpp:153-164: pcl::VoxelGrid<PCLPointCloud2>, leaf = (float)voxel_size, filter field "z"
// in [min_obstacle_height, max_obstacle_height], setDownsampleAllData(false), min points per voxel).
//
// PCL bins the accepted points by floor(p * inverse_leaf), sorts (bin, point index) pairs and emits the float centroid of
// every bin that holds enough points.  Here, hand-written (no library sort):
//   vf_bin_kernel       one thread per point: transform, limits, 63-bit bin key, find-or-insert in a hash of bins, per-bin
//                       count (one atomic per warp and bin: lanes of a warp that hit the same bin are aggregated)
//   vf_offsets_kernel   one thread per hash slot: every occupied bin gets a segment of the index array (warp-aggregated
//                       cursor) and goes onto the list of bins
//   vf_scatter_kernel   one thread per point: its index goes into its bin's segment (again one atomic per warp and bin; the
//                       lanes of a group write in lane order, so every segment is a few ascending runs)
//   vf_centroid_kernel  one WARP per bin: the segment is sorted (shuffle bitonic network up to 32 points, shared-memory
//                       bitonic up to 512, repeated minimum extraction beyond) and the member points are added in
//                       ASCENDING POINT INDEX, in float like PCL — the order the CPU oracle defines (PCL's own order inside
//                       a bin is unspecified: std::sort is unstable) — then divided by the count
// The output is COMPACT: out[0 .. *n_out) holds one centroid per emitted bin (order unspecified; Mark does not care) and
// Mark reads the count from device memory (DevReading::n_points_dev), so it only ever sees centroids.
#include <math_constants.h>
#include <algorithm>

#include "stvl_common.cuh"
#include "stvl_voxel_filter.h"

namespace stvl {

namespace {
constexpr unsigned long long kNoBin = ~0ull;
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr int kVfWarps = 4;                 // warps per CTA of the centroid kernel
constexpr int kVfSmemMax = 512;             // points per bin the shared-memory sort takes
struct VfTransform { int on; float m[12]; };
struct VfCounters { uint32_t cursor, n_bins; };

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

__global__ void __launch_bounds__(256) vf_bin_kernel(const uint8_t * __restrict__ data, unsigned long long n, uint32_t step,
                                                     uint32_t ox, uint32_t oy, uint32_t oz, float inv_leaf, double zlo, double zhi,
                                                     unsigned long long * __restrict__ keys, uint32_t hash_mask, uint32_t * __restrict__ cnt,
                                                     uint32_t * __restrict__ pt_slot, const VfTransform tf)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  uint32_t slot = kNoSlot;
  if (i < n) {
    float x, y, z;
    load_xyz(data, step, ox, oy, oz, i, x, y, z, tf);
    // PCL: skip non-finite points and points whose filter field is outside the limits (value > max || value < min)
    if (isfinite(x) && isfinite(y) && isfinite(z) && !((double)z < zlo || (double)z > zhi)) {
      const float bx = floorf(__fmul_rn(x, inv_leaf)), by = floorf(__fmul_rn(y, inv_leaf)), bz = floorf(__fmul_rn(z, inv_leaf));
      const float lim = 1048575.0f;
      if (fabsf(bx) < lim && fabsf(by) < lim && fabsf(bz) < lim) {
        const unsigned long long key = ((unsigned long long)(uint32_t)((int)bx + (1 << 20)) << 42) |
                                       ((unsigned long long)(uint32_t)((int)by + (1 << 20)) << 21) | (unsigned long long)(uint32_t)((int)bz + (1 << 20));
        uint32_t h = (uint32_t)mix64(key) & hash_mask;
        for (;;) {   // the table holds at least twice as many slots as there are points: a free slot always exists
          unsigned long long k = ld_volatile_u64(&keys[h]);
          if (k == kNoBin) k = atomicCAS(&keys[h], kNoBin, key);
          if (k == kNoBin || k == key) { slot = h; break; }
          h = (h + 1) & hash_mask;
        }
      }
    }
    pt_slot[i] = slot;
  }
  // per-bin counts, one atomic per warp and bin
  const unsigned peers = __match_any_sync(kFull, slot);
  if (slot != kNoSlot && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&cnt[slot], (uint32_t)__popc(peers));
}

__global__ void __launch_bounds__(256) vf_offsets_kernel(const uint32_t * __restrict__ cnt, uint32_t n_slots, uint32_t * __restrict__ off,
                                                         uint32_t * __restrict__ bins, VfCounters * __restrict__ ctr)
{
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t c = h < n_slots ? cnt[h] : 0u;
  // warp-aggregated: exclusive scan of the counts for the segment cursor, ballot rank for the bin list
  uint32_t incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(kFull, incl, d); if ((int)lane >= d) incl += v; }
  const uint32_t total = __shfl_sync(kFull, incl, 31);
  const unsigned occupied = __ballot_sync(kFull, c != 0);
  uint32_t base = 0, bbase = 0;
  if (lane == 0 && total) { base = atomicAdd(&ctr->cursor, total); bbase = atomicAdd(&ctr->n_bins, (uint32_t)__popc(occupied)); }
  base = __shfl_sync(kFull, base, 0);
  bbase = __shfl_sync(kFull, bbase, 0);
  if (c) {
    off[h] = base + incl - c;
    bins[bbase + __popc(occupied & ((1u << lane) - 1u))] = h;
  }
}

__global__ void __launch_bounds__(256) vf_scatter_kernel(const uint32_t * __restrict__ pt_slot, unsigned long long n, const uint32_t * __restrict__ off,
                                                         uint32_t * __restrict__ fill, uint32_t * __restrict__ seg)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t slot = i < n ? pt_slot[i] : kNoSlot;
  const unsigned peers = __match_any_sync(kFull, slot);
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (slot != kNoSlot && (int)lane == leader) base = atomicAdd(&fill[slot], (uint32_t)__popc(peers));
  base = __shfl_sync(kFull, base, leader);
  if (slot != kNoSlot) seg[off[slot] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;
}

__global__ void __launch_bounds__(kVfWarps * 32) vf_centroid_kernel(const uint8_t * __restrict__ data, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz,
                                                                    const uint32_t * __restrict__ cnt, const uint32_t * __restrict__ off,
                                                                    const uint32_t * __restrict__ bins, const VfCounters * __restrict__ ctr,
                                                                    const uint32_t * __restrict__ seg, int min_points, float4 * __restrict__ out,
                                                                    unsigned long long * __restrict__ n_out, const VfTransform tf)
{
  __shared__ uint32_t s_idx[kVfWarps][kVfSmemMax];
  __shared__ float s_xyz[kVfWarps][3][kVfSmemMax];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_bins = ctr->n_bins;
  const uint32_t need = (uint32_t)(min_points > 0 ? min_points : 0);
  for (uint32_t b = blockIdx.x * kVfWarps + warp; b < n_bins; b += gridDim.x * kVfWarps) {
    const uint32_t slot = bins[b];
    const uint32_t c = cnt[slot], o = off[slot];
    if (c < need) continue;   // PCL: a leaf needs at least min_points_per_voxel points
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (c <= 32) {
      // one index per lane, bitonic network over shuffles (padding sorts to the end)
      uint32_t v = lane < c ? seg[o + lane] : 0xFFFFFFFFu;
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const uint32_t w = __shfl_xor_sync(kFull, v, j);
          const bool up = ((lane & k) == 0);            // ascending block
          const bool lower = ((lane & j) == 0);         // this lane keeps the smaller of the pair in an ascending block
          v = (lower == up) ? min(v, w) : max(v, w);
        }
      }
      float x = 0.f, y = 0.f, z = 0.f;
      if (lane < c) load_xyz(data, step, ox, oy, oz, v, x, y, z, tf);
      // the additions in ascending point index (every lane computes the same chain)
      for (uint32_t j = 0; j < c; ++j) {
        sx = __fadd_rn(sx, __shfl_sync(kFull, x, (int)j));
        sy = __fadd_rn(sy, __shfl_sync(kFull, y, (int)j));
        sz = __fadd_rn(sz, __shfl_sync(kFull, z, (int)j));
      }
    } else if (c <= (uint32_t)kVfSmemMax) {
      uint32_t p2 = 64;
      while (p2 < c) p2 <<= 1;
      for (uint32_t j = lane; j < p2; j += 32) s_idx[warp][j] = j < c ? seg[o + j] : 0xFFFFFFFFu;
      __syncwarp();
      for (uint32_t k = 2; k <= p2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t t = lane; t < p2; t += 32) {
            const uint32_t u = t ^ j;
            if (u > t) {
              const uint32_t a = s_idx[warp][t], bb = s_idx[warp][u];
              const bool up = ((t & k) == 0);
              if ((a > bb) == up) { s_idx[warp][t] = bb; s_idx[warp][u] = a; }
            }
          }
          __syncwarp();
        }
      }
      for (uint32_t j = lane; j < c; j += 32) {
        float x, y, z;
        load_xyz(data, step, ox, oy, oz, s_idx[warp][j], x, y, z, tf);
        s_xyz[warp][0][j] = x; s_xyz[warp][1][j] = y; s_xyz[warp][2][j] = z;
      }
      __syncwarp();
      if (lane < 3) {   // lanes 0, 1, 2 add one coordinate each, in ascending point index
        float s = 0.f;
        for (uint32_t j = 0; j < c; ++j) s = __fadd_rn(s, s_xyz[warp][lane][j]);
        sx = s;
      }
      sy = __shfl_sync(kFull, sx, 1);
      sz = __shfl_sync(kFull, sx, 2);
      sx = __shfl_sync(kFull, sx, 0);
      __syncwarp();
    } else {
      // very dense bin: repeated extraction of the smallest index above the last one taken
      long long last = -1;
      for (uint32_t j = 0; j < c; ++j) {
        uint32_t m = 0xFFFFFFFFu;
        for (uint32_t t = lane; t < c; t += 32) { const uint32_t v = seg[o + t]; if ((long long)v > last && v < m) m = v; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(kFull, m, d));
        float x, y, z;
        load_xyz(data, step, ox, oy, oz, m, x, y, z, tf);
        sx = __fadd_rn(sx, x); sy = __fadd_rn(sy, y); sz = __fadd_rn(sz, z);
        last = (long long)m;
      }
    }
    if (lane == 0) {
      const float cf = (float)c;
      const unsigned long long oi = atomicAdd(n_out, 1ull);
      out[oi] = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), 0.f);
    }
  }
}

uint32_t vf_hash_slots(unsigned long long n)
{
  uint64_t p = 64;
  while (p < 2 * n) p <<= 1;
  return (uint32_t)p;
}
}  // namespace

// scratch layout: [keys u64 x H | cnt u32 x H | fill u32 x H | counters | off u32 x H | bins u32 x n | pt_slot u32 x n | seg u32 x n]
size_t voxel_filter_scratch_bytes(unsigned long long n)
{
  const size_t H = vf_hash_slots(n);
  return H * 8 + H * 4 * 3 + 256 + (size_t)n * 4 * 3 + 256;
}

cudaError_t launch_voxel_filter(const uint8_t * data, unsigned long long n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz,
                                float leaf, double zlo, double zhi, int min_points, const float * transform,
                                void * scratch, float4 * out, unsigned long long * n_out, int sm_count, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(n_out, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess || n == 0) return e;
  const uint32_t H = vf_hash_slots(n);
  uint8_t * base = static_cast<uint8_t *>(scratch);
  unsigned long long * keys = reinterpret_cast<unsigned long long *>(base);
  uint32_t * cnt = reinterpret_cast<uint32_t *>(base + (size_t)H * 8);
  uint32_t * fill = cnt + H;
  VfCounters * ctr = reinterpret_cast<VfCounters *>(fill + H);
  uint32_t * off = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(ctr) + 256);
  uint32_t * bins = off + H;
  uint32_t * pt_slot = bins + n;
  uint32_t * seg = pt_slot + n;
  e = cudaMemsetAsync(keys, 0xFF, (size_t)H * 8, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(cnt, 0, (size_t)H * 8 + 256, s);   // counts, fill cursors, counters
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const float inv_leaf = 1.0f / leaf;   // PCL: inverse_leaf_size_ = 1 / leaf_size_ (float)
  VfTransform tf;
  tf.on = transform != nullptr;
  for (int i = 0; i < 12; ++i) tf.m[i] = transform ? transform[i] : 0.f;
  vf_bin_kernel<<<blocks, 256, 0, s>>>(data, n, step, ox, oy, oz, inv_leaf, zlo, zhi, keys, H - 1, cnt, pt_slot, tf);
  vf_offsets_kernel<<<(H + 255) / 256, 256, 0, s>>>(cnt, H, off, bins, ctr);
  vf_scatter_kernel<<<blocks, 256, 0, s>>>(pt_slot, n, off, fill, seg);
  const unsigned cblocks = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((n + kVfWarps - 1) / kVfWarps, (unsigned long long)sm_count * 8));
  vf_centroid_kernel<<<cblocks, kVfWarps * 32, 0, s>>>(data, step, ox, oy, oz, cnt, off, bins, ctr, seg, min_points, out, n_out, tf);
  return cudaGetLastError();
}

}  // namespace stvl
