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
#include <cuda_pipeline.h>
#include <math_constants.h>
#include <algorithm>
#include <utility>

#include "stvl_device.cuh"
#include "stvl_kernels.h"

namespace stvl {

constexpr unsigned kFull = 0xFFFFFFFFu;

// fire-and-forget global reductions (no return value, no scoreboard wait)
__device__ __forceinline__ void red_or_u32(uint32_t * p, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
// "later reading overwrites" as a SIGNED 64-bit max on the double's bit pattern: every clock sample of a batched launch is
// > 0, and a stored time may have been pushed below zero by a sweep (value - accel, reference .cpp:190), whose bit pattern
// is a negative integer — an unsigned max would keep the stale negative time
__device__ __forceinline__ void red_max_s64(unsigned long long * p, unsigned long long v) { asm volatile("red.global.max.s64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running, once every CTA of the predecessor has executed
// launch_dependents (or exited).  It must not touch anything the predecessor writes before pdl_wait(), which returns
// when the predecessor grid has completed and its writes are visible.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// =====================================================================================
// hash helpers
// =====================================================================================
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t * p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t * p) { return *reinterpret_cast<const volatile uint64_t *>(p); }

// find-or-insert the column tile of leaf column (bx,by); tiles are never freed between rebuilds
__device__ uint32_t tile_find_or_insert(const GridView & g, int bx, int by)
{
  const uint64_t key = pack_tile_key(bx, by);
  uint32_t h = (uint32_t)mix64(key) & g.tile_hash_mask;
  for (uint32_t probe = 0; probe <= g.tile_hash_mask; ++probe) {
    uint64_t k = ld_volatile_u64(&g.tile_hash_keys[h]);
    if (k == kEmptyKey) {
      const uint64_t old = atomicCAS((unsigned long long *)&g.tile_hash_keys[h], (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) {
        const uint32_t t = atomicAdd(&g.ctr->tile_n, 1u);
        if (t >= g.tile_capacity) {
          atomicExch(&g.ctr->tile_overflow, 1u);
          __threadfence();
          *reinterpret_cast<volatile uint32_t *>(&g.tile_hash_vals[h]) = kOverflowSlot;
          return kOverflowSlot;
        }
        g.tile_key[t] = key;
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(&g.tile_hash_vals[h]) = t;
        return t;
      }
      k = old;
    }
    if (k == key) {
      uint32_t v;
      while ((v = ld_volatile_u32(&g.tile_hash_vals[h])) == kInvalidSlot) { __nanosleep(20); }
      return v;
    }
    h = (h + 1) & g.tile_hash_mask;
  }
  atomicExch(&g.ctr->tile_overflow, 1u);
  return kOverflowSlot;
}

// hand out a leaf slot: first the free stack (as it stood when the kernel started), then fresh slots
__device__ __forceinline__ uint32_t leaf_alloc(const GridView & g, uint32_t free_n0, uint32_t high_water0)
{
  const uint32_t i = atomicAdd(&g.ctr->alloc_cursor, 1u);
  uint32_t slot;
  if (i < free_n0) {
    slot = g.free_slots[free_n0 - 1 - i];
  } else {
    const unsigned long long s = (unsigned long long)high_water0 + (i - free_n0);
    slot = s >= g.leaf_capacity ? kOverflowSlot : (uint32_t)s;
  }
  return slot;
}

// find-or-insert a leaf; returns its pool slot (or kOverflowSlot).  Probes use ordinary (L1-cacheable) 128-bit loads:
// a stale line can only show EMPTY where a key has since been published (the CAS then returns the truth) or an
// unpublished slot (then the volatile spin below reads L2); keys never change while a Mark kernel runs.
__device__ uint32_t leaf_find_or_insert(const GridView & g, uint64_t key, int bx, int by, uint32_t free_n0, uint32_t high_water0,
                                        double tmin_init)
{
  uint32_t h = (uint32_t)mix64(key) & g.hash_mask;
  for (uint32_t probe = 0; probe <= g.hash_mask; ++probe) {
    const uint4 e = *reinterpret_cast<const uint4 *>(&g.hash[h]);
    uint64_t k = (uint64_t)e.x | ((uint64_t)e.y << 32);
    uint32_t v = e.z;
    if (k == kEmptyKey) {
      const uint64_t old = atomicCAS(&g.hash[h].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) {
        const uint32_t slot = leaf_alloc(g, free_n0, high_water0);
        if (slot == kOverflowSlot) {
          atomicExch(&g.ctr->leaf_overflow, 1u);
        } else {
          const uint32_t tile = tile_find_or_insert(g, bx, by);
          g.leaf_key[slot] = key;
          g.leaf_hpos[slot] = h;
          g.leaf_tile[slot] = tile;
          g.leaf_tmin[slot] = tmin_init;
        }
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(&g.hash[h].val) = slot;
        return slot;
      }
      k = old;
      v = kInvalidSlot;
    }
    if (k == key) {
      while (v == kInvalidSlot) { v = ld_volatile_u32(&g.hash[h].val); if (v == kInvalidSlot) __nanosleep(20); }
      return v;
    }
    h = (h + 1) & g.hash_mask;   // other key or tombstone: keep probing
  }
  atomicExch(&g.ctr->leaf_overflow, 1u);
  return kOverflowSlot;
}

// fold the leaves handed out by the previous Mark / load kernel into the allocator state.  Runs single-threaded
// between kernels (stream order), so allocating kernels need no per-block epilogue.
__device__ __forceinline__ void fold_alloc(const GridView & g)
{
  Counters * c = g.ctr;
  const uint32_t cursor = ld_volatile_u32(&c->alloc_cursor);
  if (cursor == 0) return;
  const uint32_t free_n = ld_volatile_u32(&c->free_n), hw = ld_volatile_u32(&c->high_water);
  const uint32_t from_free = cursor < free_n ? cursor : free_n;
  unsigned long long nhw = (unsigned long long)hw + (cursor - from_free);
  if (nhw > g.leaf_capacity) nhw = g.leaf_capacity;
  c->free_n = free_n - from_free;
  c->high_water = (uint32_t)nhw;
  c->alloc_cursor = 0;
  if (g.live) {   // mirror for the host's grid sizing
    volatile LiveCounters * l = g.live;
    // (posted stores only: a LOAD from host-mapped memory here would put a PCIe round trip at the tail of the kernel)
    l->high_water = (uint32_t)nhw; l->free_n = free_n - from_free; l->tile_n = ld_volatile_u32(&c->tile_n);
  }
}
__global__ void fold_alloc_kernel(const GridView g) { fold_alloc(g); }

// Opt-in phase tracing for profiling builds only (python -m spatio_temporal_voxel_layer_b200.build --trace builds a
// SEPARATE libstvl_b200_trace.so; the product library compiles none of this).  Thread 0 of every CTA stamps
// %globaltimer at a few phase boundaries; profiles/trace_phases.py turns the stamps into a per-kernel timeline.
#ifdef STVL_TRACE
__device__ unsigned long long g_trace[4][1024][8];
__device__ __forceinline__ void trace_stamp(int kernel, int phase)
{
  if (threadIdx.x == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[kernel][blockIdx.x][phase] = t;
    if (phase == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_trace[kernel][blockIdx.x][7] = sm; }
  }
}
extern "C" int stvl_debug_trace(void * out, size_t bytes)
{
  if (bytes != sizeof(g_trace)) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)) != cudaSuccess) return -3;
  return 0;
}
#define TRACE(k, ph) trace_stamp(k, ph)
#define TRACE_VAL(k, ph, v) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_trace[k][blockIdx.x][ph] = (unsigned long long)(v); } while (0)
#else
#define TRACE(k, ph) ((void)0)
#define TRACE_VAL(k, ph, v) ((void)0)
#endif

// =====================================================================================
// Costmap — reference spatio_temporal_voxel_layer.cpp:699-718 (grid loop) on the tile counts
// =====================================================================================
// nav2 Costmap2D::worldToMap (external): wx >= origin, (unsigned)((wx-origin)/resolution) < size
__device__ __forceinline__ bool world_to_map(double wx, double wy, const CostmapDev & p, unsigned & mx, unsigned & my)
{
  if (wx < p.origin_x || wy < p.origin_y) return false;
  mx = __double2uint_rz(__ddiv_rn(__dsub_rn(wx, p.origin_x), p.resolution));
  my = __double2uint_rz(__ddiv_rn(__dsub_rn(wy, p.origin_y), p.resolution));
  return mx < p.size_x && my < p.size_y;
}

// last CTA of the costmap pass: publish the result and re-arm the accumulators (no separate prologue launch)
__device__ __forceinline__ void costmap_publish_result(Counters * c)
{
  volatile CostmapCtr * a = &c->cm_acc;
  c->cm.n_marked = a->n_marked;
  c->cm.mk_min_x = a->mk_min_x; c->cm.mk_min_y = a->mk_min_y; c->cm.mk_max_x = a->mk_max_x; c->cm.mk_max_y = a->mk_max_y;
  a->n_marked = 0;
  a->mk_min_x = INT_MAX; a->mk_min_y = INT_MAX; a->mk_max_x = INT_MIN; a->mk_max_y = INT_MIN;
}

// The grid loop of UpdateROSCostmap (:708-717) for `n_batch` consecutive column tiles starting at t0, by one warp: the
// count rows (256 B each, two columns per lane) and tile keys of the whole batch are requested together.
template <int kBatch>
__device__ __forceinline__ void costmap_tile_batch(const GridView & g, const CostmapDev & p, uint8_t * __restrict__ costmap, uint32_t t0,
                                                   uint32_t n_tiles, unsigned lane, unsigned & n, int & mnx, int & mny, int & mxx, int & mxy)
{
  uint2 cc[kBatch];
  unsigned long long keys[kBatch];
#pragma unroll
  for (int u = 0; u < kBatch; ++u) {
    const uint32_t t = t0 + u;
    cc[u] = t < n_tiles ? reinterpret_cast<const uint2 *>(g.tile_count + (size_t)t * 64)[lane] : make_uint2(0, 0);   // columns 2*lane, 2*lane+1
    keys[u] = t < n_tiles ? g.tile_key[t] : 0ull;
  }
#pragma unroll
  for (int u = 0; u < kBatch; ++u) {
    if ((cc[u].x | cc[u].y) == 0) continue;
    int bx, by;
    unpack_tile_key(keys[u], bx, by);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t cnt = k ? cc[u].y : cc[u].x;
      const int col = 2 * (int)lane + k;
      if (cnt == 0 || (int)cnt < p.mark_threshold) continue;   // static_cast<int>(count) >= _mark_threshold, :712
      const int ix = bx * 8 + (col >> 3), iy = by * 8 + (col & 7);
      unsigned mx, my;
      if (world_to_map(index_to_world(ix, g.vs, g.half_vs), index_to_world(iy, g.vs, g.half_vs), p, mx, my)) {
        costmap[(size_t)my * p.size_x + mx] = p.lethal;   // :715
        ++n; mnx = min(mnx, ix); mny = min(mny, iy); mxx = max(mxx, ix); mxy = max(mxy, iy);   // touch(), :716
      }
    }
  }
}

// block reduction of the marked-cell count / bounds into the device accumulators; with own_ticket the last CTA of the
// grid also publishes (the fused Mark + costmap kernel publishes on Mark's ticket instead)
__device__ __forceinline__ void costmap_reduce_and_publish(Counters * c, unsigned n, int mnx, int mny, int mxx, int mxy, bool own_ticket = true)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n += __shfl_xor_sync(kFull, n, d);
    mnx = min(mnx, __shfl_xor_sync(kFull, mnx, d)); mny = min(mny, __shfl_xor_sync(kFull, mny, d));
    mxx = max(mxx, __shfl_xor_sync(kFull, mxx, d)); mxy = max(mxy, __shfl_xor_sync(kFull, mxy, d));
  }
  __shared__ unsigned int s_n;
  __shared__ int s_b[4];
  if (threadIdx.x == 0) { s_n = 0; s_b[0] = s_b[1] = INT_MAX; s_b[2] = s_b[3] = INT_MIN; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && n) {
    atomicAdd(&s_n, n);
    atomicMin(&s_b[0], mnx); atomicMin(&s_b[1], mny); atomicMax(&s_b[2], mxx); atomicMax(&s_b[3], mxy);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_n) {
      atomicAdd(&c->cm_acc.n_marked, (unsigned long long)s_n);
      atomicMin(&c->cm_acc.mk_min_x, s_b[0]); atomicMin(&c->cm_acc.mk_min_y, s_b[1]);
      atomicMax(&c->cm_acc.mk_max_x, s_b[2]); atomicMax(&c->cm_acc.mk_max_y, s_b[3]);
    }
    if (!own_ticket) return;
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
// Mark — reference .cpp:269-309
// =====================================================================================
struct PointXYZ { float x, y, z; bool ok; };

template <bool kVec4>
__device__ __forceinline__ PointXYZ load_point(const DevReading & r, unsigned long long i)
{
  PointXYZ p;
  if (i >= r.n_points) { p.x = p.y = p.z = 0.f; p.ok = false; return p; }
  if (kVec4 || r.vec4) {
    // coalesced 128-bit streaming load: the cloud is read exactly once
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(r.data) + i);
    p.x = v.x; p.y = v.y; p.z = v.z;
  } else {
    const uint8_t * b = r.data + i * r.point_step;
    if (((reinterpret_cast<uintptr_t>(b) | r.off_x | r.off_y | r.off_z) & 3u) == 0) {
      p.x = __ldcs(reinterpret_cast<const float *>(b + r.off_x));
      p.y = __ldcs(reinterpret_cast<const float *>(b + r.off_y));
      p.z = __ldcs(reinterpret_cast<const float *>(b + r.off_z));
    } else {
      uint32_t w[3];
      const uint32_t off[3] = {r.off_x, r.off_y, r.off_z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const uint8_t * q = b + off[a];
        w[a] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
      }
      p.x = __uint_as_float(w[0]); p.y = __uint_as_float(w[1]); p.z = __uint_as_float(w[2]);
    }
  }
  p.ok = true;
  return p;
}

constexpr int kMarkThreads = 256;
constexpr int kMarkPointsPerThread = 4;
constexpr int kMarkPointsPerBlock = kMarkThreads * kMarkPointsPerThread;
constexpr int kMarkTable = 128;       // block-local (reading, leaf) table (shared memory)
constexpr int kMarkProbes = 16;
constexpr size_t kMarkStageBytes = 2 * (size_t)kMarkPointsPerBlock * 16;   // two cp.async stages of one chunk each

// write one voxel straight to the global store (fallback when the block-local table is full)
__device__ __forceinline__ void mark_voxel_direct(const GridView & g, uint64_t key, unsigned l, double now, unsigned long long now_bits,
                                                  bool use_max, uint32_t free_n0, uint32_t high_water0, double tmin_init)
{
  int bx, by, bz;
  unpack_leaf_key(key, bx, by, bz);
  const uint32_t slot = leaf_find_or_insert(g, key, bx, by, free_n0, high_water0, tmin_init);
  if (slot == kOverflowSlot) return;
  atomicOr(&g.leaf_mask[(size_t)slot * 16 + (l >> 5)], 1u << (l & 31));
  double * tp = &g.leaf_time[(size_t)slot * kLeafVoxels + l];
  if (use_max) atomicMax(reinterpret_cast<long long *>(tp), (long long)now_bits);
  else {
    *tp = now;   // MarkGridPoint: setValueOn overwrite, .cpp:421-430
    if (g.leaf_tmin[slot] > now) g.leaf_tmin[slot] = -CUDART_INF;   // clocks are not monotone here
  }
}

// Mark — persistent CTAs that pull 1024-point chunks of the batched clouds from a device-side cursor:
//   1. packed clouds are streamed global -> shared memory with cp.async, the NEXT chunk in flight while the current one
//      is processed (the id of the chunk after that is requested from the cursor at the same time, so the atomic's
//      latency is hidden too).  Dynamic chunks keep every resident CTA busy whatever the mix of valid / invalid pixels,
//      and — under programmatic dependent launch — the CTAs that become resident early, while the sweep still occupies
//      most of the SMs, take most of the work; CTAs that only become resident late find the cursor exhausted.
//   2. range / z-window tests and key quantisation in the reference's exact arithmetic (.cpp:285-303), written
//      branch-free so the four points of a thread advance together
//   3. block-level dedupe in shared memory: a depth image puts ~100 points into every voxel and ~1000 into every
//      8^3 leaf, so the CTA ORs its points into a table of {(reading, leaf), 512-bit mask} that lives across chunks
//      (a 1024-point chunk of a depth image touches 4-10 leaves)
//   4. ONE flush at the end: one global hash probe per distinct leaf, one 32-bit OR per touched mask word and one
//      64-bit max (or store) per distinct voxel.  Nothing before the flush touches the grid, so all of 1-3 may overlap
//      the sweep that precedes Mark on the stream (programmatic dependent launch).
__device__ __forceinline__ int mark_reading_of(const MarkBatch & batch, uint32_t chunk)
{
  int ri = 0;
#pragma unroll 1
  for (int k = 1; k < (int)batch.n_readings; ++k) if (chunk >= batch.first_block[k]) ri = k;
  return ri;
}

// key of the block-local table: reading index (the mark time differs per reading) + leaf coordinate, x relative to the
// batch's reference leaf so that everything fits 63 bits; leaves farther than 2^16 leaves from it bypass the table
constexpr int kMarkRelBias = 1 << 16;
__device__ __forceinline__ uint64_t pack_table_key(int ri, int rx, int by, int bz)
{
  return ((uint64_t)(uint32_t)ri << 59) | ((uint64_t)(uint32_t)(rx + kMarkRelBias) << 42) |
         ((uint64_t)(uint32_t)(by + kLeafBias) << 21) | (uint64_t)(uint32_t)(bz + kLeafBias);
}

// kVec4: every reading of the batch is a packed x,y,z,pad float4 cloud (the RGBD / depth-image layout).
// kCostmap: the fused update cycle — this launch also runs the costmap pass over the column counts of the sweep that
// precedes it (they are independent of Mark: the costmap reflects the pre-Mark survivors, reference .cpp:149-224 then
// spatio_temporal_voxel_layer.cpp:699-718); the costmap bytes were reset by the sweep's triage kernel.
template <bool kVec4, bool kCostmap>
__global__ void __launch_bounds__(kMarkThreads, kVec4 ? 5 : 4)
mark_kernel(const GridView g, const __grid_constant__ MarkBatch batch, const CostmapDev cm, uint8_t * __restrict__ costmap)
{
  __shared__ unsigned long long s_key[kMarkTable];
  __shared__ uint32_t s_mask[kMarkTable][16];
  __shared__ uint32_t s_slot[kMarkTable];
  __shared__ double s_now[kMarkTable];
  __shared__ unsigned int s_drop[2];
  __shared__ uint32_t s_ids[2];        // chunk ids handed to this CTA by the cursor (for iterations i+1, i+2)
  __shared__ unsigned int s_fill;      // claimed table entries
  TRACE(0, 0);
  // The reading descriptors stay in the kernel-parameter bank and are only ever read with warp-uniform addresses
  // (every thread of the CTA reads the same field of the same reading): per-thread addresses into the constant bank
  // serialise, and copying the whole batch to shared memory that way cost every CTA ~2 us before its first load.
  const unsigned lane = threadIdx.x & 31;
  const bool use_max = batch.use_atomic_max != 0;
  const uint32_t total = batch.total_blocks;
  // Everything up to the flush only reads the clouds and the kernel parameters, so under programmatic dependent
  // launch it overlaps the tail of the sweep; the grid state (allocator counters, hash, leaves) is touched only after
  // sync_with_predecessor().  (The chunk cursor is only ever touched by Mark kernels, and consecutive launches alternate
  // between two cursors, see Counters::mark_cursor.)
  uint32_t * const cursor = &g.ctr->mark_cursor[batch.cursor_sel & 1u];
  if (threadIdx.x == 0) {   // the first two chunks of this CTA: requested before anything else is set up
    const uint32_t c0 = atomicAdd(cursor, 2u);
    s_ids[0] = c0; s_ids[1] = c0 + 1;
  }
  uint32_t free_n0 = 0, high_water0 = 0;
  bool grid_state_ok = false;
  auto sync_with_predecessor = [&]() {
    if (!grid_state_ok) {
      pdl_wait();
      free_n0 = ld_volatile_u32(&g.ctr->free_n);
      high_water0 = ld_volatile_u32(&g.ctr->high_water);
      grid_state_ok = true;
    }
  };
  unsigned int drop_range = 0, drop_nonfinite = 0;
  unsigned cm_n = 0;               // fused costmap pass: marked cells and their index bounds
  int cm_mnx = INT_MAX, cm_mny = INT_MAX, cm_mxx = INT_MIN, cm_mxy = INT_MIN;
  uint64_t last_key = kEmptyKey;   // this thread's most recent table key and its entry
  int last_e = -1;

  // Cloud staging.  Packed float4 clouds (kVec4) are streamed global -> shared memory with cp.async (LDGSTS), two
  // 16 KB stages per CTA: the next chunk is in flight while the current one is processed and the copies hold no
  // registers, which buys a fifth resident CTA per SM.  Other PointCloud2 layouts prefetch through registers.
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float4 * s_stage = reinterpret_cast<float4 *>(s_dyn);   // [2][kMarkPointsPerBlock] when kVec4
  PointXYZ nxt[kVec4 ? 1 : kMarkPointsPerThread];
  auto prefetch = [&](int rix, uint32_t chunk_id, int stage) {
    const DevReading & rd = batch.r[rix];
    const unsigned long long base = (unsigned long long)(chunk_id - batch.first_block[rix]) * kMarkPointsPerBlock;
    if constexpr (kVec4) {
      const unsigned long long n_pts = rd.n_points;
      const float4 * src = reinterpret_cast<const float4 *>(rd.data);
#pragma unroll
      for (int k = 0; k < kMarkPointsPerThread; ++k) {
        const unsigned long long i = base + (unsigned long long)k * kMarkThreads + threadIdx.x;
        if (i < n_pts)
          __pipeline_memcpy_async(&s_stage[stage * kMarkPointsPerBlock + k * kMarkThreads + threadIdx.x], src + i, 16);
      }
      __pipeline_commit();
    } else {
#pragma unroll
      for (int k = 0; k < kMarkPointsPerThread; ++k) nxt[k] = load_point<false>(rd, base + (unsigned long long)k * kMarkThreads + threadIdx.x);
    }
  };
  if (threadIdx.x < 2) s_drop[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_fill = 0;
  for (int i = threadIdx.x; i < kMarkTable; i += kMarkThreads) s_key[i] = kEmptyKey;
#pragma unroll
  for (int i = threadIdx.x; i < kMarkTable * 16; i += kMarkThreads) (&s_mask[0][0])[i] = 0;
  __syncthreads();
  uint32_t chunk = s_ids[0];
  int ri = 0;
  if (chunk < total) {
    ri = mark_reading_of(batch, chunk);
    prefetch(ri, chunk, 0);
  }
  TRACE(0, 1);

  // ---- flush: resolve the table's leaves in the global hash, push masks and times, empty the table ---------------------
  auto flush = [&](const bool final) {
    if (final) {
      TRACE(0, 3);   // all chunks processed
      pdl_launch_dependents();   // the next cycle's triage kernel may become resident and stage its parameters
    }
    sync_with_predecessor();
    if constexpr (kCostmap) {
      // the costmap pass rides on the warps that would otherwise idle while the first warps resolve the leaves
      if (final && threadIdx.x >= kMarkTable) {
        const uint32_t n_tiles = min(ld_volatile_u32(&g.ctr->tile_n), g.tile_capacity);
        constexpr uint32_t kCmWarps = (kMarkThreads - kMarkTable) / 32, kCmBatch = 4;
        const uint32_t cw = blockIdx.x * kCmWarps + ((threadIdx.x - kMarkTable) >> 5);
        for (uint32_t t0 = cw * kCmBatch; t0 < n_tiles; t0 += gridDim.x * kCmWarps * kCmBatch)
          costmap_tile_batch<(int)kCmBatch>(g, cm, costmap, t0, n_tiles, lane, cm_n, cm_mnx, cm_mny, cm_mxx, cm_mxy);
      }
    }
    if (threadIdx.x < kMarkTable) {   // one thread per distinct (reading, leaf) resolves (or creates) the leaf in the global hash
      const unsigned long long tkey = s_key[threadIdx.x];
      uint32_t slot = kOverflowSlot;
      if (tkey != kEmptyKey) {
        const int ri = (int)(tkey >> 59);
        const int bx = (int)((tkey >> 42) & 0x1FFFF) - kMarkRelBias + batch.ref_bx;
        const int by = (int)((tkey >> 21) & 0x1FFFFF) - kLeafBias, bz = (int)(tkey & 0x1FFFFF) - kLeafBias;
        slot = leaf_find_or_insert(g, pack_leaf_key(bx, by, bz), bx, by, free_n0, high_water0, batch.min_now);
        s_now[threadIdx.x] = batch.r[ri].now;
        s_key[threadIdx.x] = kEmptyKey;
      }
      s_slot[threadIdx.x] = slot;
    }
    last_key = kEmptyKey; last_e = -1;
    __syncthreads();
    if (threadIdx.x == 0) s_fill = 0;   // (after the barrier: every thread has taken its flush decision from it)
    if (final) TRACE(0, 4);   // leaves resolved in the global hash
    // per touched mask word one 32-bit OR, per distinct voxel one 64-bit max (batched launch, monotone clocks) or a
    // plain store (per-reading launches: exact overwrite semantics of MarkGridPoint, .cpp:421-430)
#pragma unroll
    for (int i = threadIdx.x; i < kMarkTable * 16; i += kMarkThreads) {
      uint32_t m = (&s_mask[0][0])[i];
      if (!m) continue;
      (&s_mask[0][0])[i] = 0;
      const uint32_t slot = s_slot[i >> 4];
      if (slot == kOverflowSlot) continue;
      const double now = s_now[i >> 4];
      const unsigned long long now_bits = (unsigned long long)__double_as_longlong(now);
      const unsigned w = i & 15;
      red_or_u32(&g.leaf_mask[(size_t)slot * 16 + w], m);
      double * tp = g.leaf_time + (size_t)slot * kLeafVoxels + w * 32;
      if (!use_max && g.leaf_tmin[slot] > now) g.leaf_tmin[slot] = -CUDART_INF;   // clocks are not monotone here
      while (m) {
        const int b = __ffs(m) - 1; m &= m - 1;
        if (use_max) red_max_s64(reinterpret_cast<unsigned long long *>(tp + b), now_bits);
        else tp[b] = now;
      }
    }
    __syncthreads();
  };

  uint32_t n_chunks_done = 0;   // (only read by the trace build)
  for (uint32_t it = 0; chunk < total; ++it) {
    ++n_chunks_done;
    const int stage = (int)(it & 1u);
    PointXYZ pts[kMarkPointsPerThread];
    if constexpr (!kVec4) {
#pragma unroll
      for (int k = 0; k < kMarkPointsPerThread; ++k) pts[k] = nxt[k];
    }
    // the next chunk of this CTA (handed out one iteration ago): start its loads; ask the cursor for the one after it
    const uint32_t chunk_next = s_ids[stage ^ 1];
    const bool last = chunk_next >= total;
    int ri_next = ri;
    uint32_t after_next = 0;
    if (!last) {
      ri_next = mark_reading_of(batch, chunk_next);
      prefetch(ri_next, chunk_next, stage ^ 1);
      if (threadIdx.x == 0) after_next = atomicAdd(cursor, 1u);
    }
    const DevReading & par = batch.r[ri];   // warp-uniform parameter-bank reads
    if constexpr (kVec4) {
      // every thread only reads what it copied itself: waiting for its own cp.async group is enough
      if (last) __pipeline_wait_prior(0); else __pipeline_wait_prior(1);
      // points of this chunk that exist (only the reading's last chunk is short): one 64-bit subtraction per chunk, then
      // 32-bit compares
      const unsigned long long base = (unsigned long long)(chunk - batch.first_block[ri]) * kMarkPointsPerBlock;
      const unsigned long long left = par.n_points - base;
      const unsigned in_chunk = left < (unsigned long long)kMarkPointsPerBlock ? (unsigned)left : (unsigned)kMarkPointsPerBlock;
#pragma unroll
      for (int k = 0; k < kMarkPointsPerThread; ++k) {
        const float4 v = s_stage[stage * kMarkPointsPerBlock + k * kMarkThreads + threadIdx.x];
        pts[k].x = v.x; pts[k].y = v.y; pts[k].z = v.z;
        pts[k].ok = (unsigned)(k * kMarkThreads) + threadIdx.x < in_chunk;
      }
    }
    if (it == 0) TRACE(0, 2);   // first chunk's data has arrived

    // ---- cheap rejection first (branch-free): depth images are full of invalid pixels and floor / ceiling returns
    const double ox = par.ox, oy = par.oy, oz = par.oz, now_r = par.now;
    const float range2 = par.mark_range_2, range2_reject = par.range2_reject;
    const float oxf = par.oxf, oyf = par.oyf, ozf = par.ozf;
    const bool zwin = par.filter == 2;
    const float zlo = zwin ? par.zmin_f : -CUDART_INF_F, zhi = zwin ? par.zmax_f : CUDART_INF_F;
    if (par.has_tf) {   // cloud still in the sensor frame: tf2::doTransform first (measurement_buffer.cpp:141-146)
#pragma unroll
      for (int k = 0; k < kMarkPointsPerThread; ++k) transform_point(par.tf, pts[k].x, pts[k].y, pts[k].z);
    }
    bool pre[kMarkPointsPerThread];
#pragma unroll
    for (int k = 0; k < kMarkPointsPerThread; ++k) {
      const float fx = pts[k].x, fy = pts[k].y, fz = pts[k].z;
      const bool finite = isfinite(fx) && isfinite(fy) && isfinite(fz);
      drop_nonfinite += (pts[k].ok && !finite) ? 1u : 0u;
      // PassThrough z window (reference measurement_buffer.cpp:165-175) on exactly equivalent float thresholds, and a
      // single-precision range screen that only ever rejects points the exact test below would reject too
      const float ax = fx - oxf, ay = fy - oyf, az = fz - ozf;
      const float d2f = ax * ax + ay * ay + az * az;
      pre[k] = pts[k].ok && finite && !(fz < zlo || fz > zhi) && !(d2f > range2_reject);
    }
    // ---- exact arithmetic + block-local dedupe for the survivors ------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kMarkPointsPerThread; ++k) {
      if (!pre[k]) continue;
      const float fx = pts[k].x, fy = pts[k].y, fz = pts[k].z;
      const double xd = (double)fx, yd = (double)fy, zd = (double)fz;
      // .cpp:285-291  (double arithmetic, narrowed to float, compared against the float range^2 and 0.0001)
      const double dx = __dsub_rn(xd, ox), dy = __dsub_rn(yd, oy), dz = __dsub_rn(zd, oz);
      const float distance_2 = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      // (the remaining rejections are predicates, not branches: nearly every point that gets here passes them, and a
      // warp runs the arithmetic below as long as one lane needs it anyway)
      const bool in_range = !(distance_2 > range2 || (double)distance_2 < 0.0001);
      // .cpp:293-303: shift negatives by one voxel (z tests y's sign — sic), multiply by 1/vs, truncate toward zero
      const double X = fx < 0 ? __dsub_rn(xd, g.vs) : xd;
      const double Y = fy < 0 ? __dsub_rn(yd, g.vs) : yd;
      const double Z = fy < 0 ? __dsub_rn(zd, g.vs) : zd;
      const double gx = __dmul_rn(X, g.inv_vs), gy = __dmul_rn(Y, g.inv_vs), gz = __dmul_rn(Z, g.inv_vs);
      const double lim = 8388608.0;   // 2^23 voxels = 2^20 leaves
      const bool in_lim = fabs(gx) < lim && fabs(gy) < lim && fabs(gz) < lim;
      drop_range += (in_range && !in_lim) ? 1u : 0u;
      const int ix = __double2int_rz(gx), iy = __double2int_rz(gy), iz = __double2int_rz(gz);   // (saturated garbage when !in_lim: masked)
      const int bx = ix >> 3, by = iy >> 3, bz = iz >> 3;
      const bool owns = g.shard_count <= 1 || shard_owns(bx, g.shard_index, g.shard_count, g.shard_width);
      if (!(in_range && in_lim && owns)) continue;
      const unsigned l = (unsigned)(((ix & 7) * 8 + (iy & 7)) * 8 + (iz & 7));
      // find-or-claim the (reading, leaf) in the block-local table (consecutive points of a thread mostly share it)
      const int rx = bx - batch.ref_bx;
      const bool in_reach = rx >= -kMarkRelBias && rx < kMarkRelBias;
      const uint64_t key = pack_table_key(ri, rx, by, bz);
      int e = -1;
      if (in_reach) {
        if (key == last_key) e = last_e;
        else {
          uint32_t h = ((uint32_t)bx * 73856093u ^ (uint32_t)by * 19349663u ^ (uint32_t)bz * 83492791u ^ (uint32_t)ri * 2654435761u) & (kMarkTable - 1);
#pragma unroll 1
          for (int p = 0; p < kMarkProbes; ++p) {
            unsigned long long cur = s_key[h];
            if (cur == kEmptyKey) {
              cur = atomicCAS(&s_key[h], (unsigned long long)kEmptyKey, (unsigned long long)key);
              if (cur == kEmptyKey) atomicAdd(&s_fill, 1u);
            }
            if (cur == key || cur == kEmptyKey) { e = (int)h; break; }
            h = (h + 1) & (kMarkTable - 1);
          }
          last_key = key; last_e = e;
        }
      }
      if (e >= 0) {
        const uint32_t bit = 1u << (l & 31);
        if (!(s_mask[e][l >> 5] & bit)) atomicOr(&s_mask[e][l >> 5], bit);
      } else {   // table full (scattered cloud) or leaf out of the table's reach: straight to the global store
        sync_with_predecessor();
        mark_voxel_direct(g, pack_leaf_key(bx, by, bz), l, now_r, (unsigned long long)__double_as_longlong(now_r), use_max,
                          free_n0, high_water0, batch.min_now);
      }
    }
    if (!last && threadIdx.x == 0) s_ids[stage] = after_next;   // consumed two iterations from now
    __syncthreads();   // (the only barrier of the chunk loop: hands the next chunk id to the other warps)
    // A filling table is flushed early: large clouds give a CTA dozens of chunks (a chunk of an organised cloud touches
    // 4-10 leaves), and points that find the table full take the slow per-point path.  Under programmatic dependent
    // launch this blocks until the sweep is done — the 3-4 chunks per CTA of a depth-camera cycle never get here.
    if (!last && s_fill >= (unsigned)(kMarkTable * 5 / 8)) flush(false);
    chunk = chunk_next;
    ri = ri_next;
  }
  flush(true);

  TRACE(0, 5);   // flushed
#ifdef STVL_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 1024) g_trace[0][blockIdx.x][7] |= (unsigned long long)n_chunks_done << 16;
#endif
  // dropped-point statistics: warp reduce, shared-memory add, one global atomic per CTA
  if (__ballot_sync(kFull, (drop_range | drop_nonfinite) != 0)) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      drop_range += __shfl_xor_sync(kFull, drop_range, d);
      drop_nonfinite += __shfl_xor_sync(kFull, drop_nonfinite, d);
    }
    if (lane == 0) {
      if (drop_range) atomicAdd(&s_drop[0], drop_range);
      if (drop_nonfinite) atomicAdd(&s_drop[1], drop_nonfinite);
    }
  }
  __syncthreads();
  if constexpr (kCostmap) costmap_reduce_and_publish(g.ctr, cm_n, cm_mnx, cm_mny, cm_mxx, cm_mxy, false);
  if (threadIdx.x == 0) {
    if (s_drop[0]) atomicAdd(&g.ctr->dropped_range, (unsigned long long)s_drop[0]);   // rare: points beyond +-2^23 voxels
    // One atomic per CTA: the ticket also carries this CTA's non-finite point count (every CTA of a depth image has some;
    // separate counters on the same cache line tripled the same-address atomic traffic when all CTAs finish together).
    // The last CTA commits the leaves handed out by this launch to the allocator state (and publishes the costmap result).
    __threadfence();
    const unsigned long long mine = 1ull + ((unsigned long long)s_drop[1] << 16);
    const unsigned long long before = atomicAdd(&g.ctr->mark_ticket64, mine);
    if ((before & 0xFFFFull) == (unsigned long long)(gridDim.x - 1)) {
      __threadfence();
      fold_alloc(g);
      if constexpr (kCostmap) costmap_publish_result(g.ctr);
      g.ctr->dropped_nonfinite += (before + mine) >> 16;
      g.ctr->mark_ticket64 = 0;
      *cursor = 0;
      __threadfence();
    }
    TRACE(0, 6);
  }
}

// bulk load of (index, time) pairs — MarkGridPoint semantics (overwrite), for tests / restore
__global__ void __launch_bounds__(256)
load_voxels_kernel(const GridView g, const int32_t * __restrict__ vx, const int32_t * __restrict__ vy,
                   const int32_t * __restrict__ vz, const double * __restrict__ vt, unsigned long long n)
{
  const uint32_t free_n0 = g.ctr->free_n;
  const uint32_t high_water0 = g.ctr->high_water;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int ix = vx[i], iy = vy[i], iz = vz[i];
    const int bx = ix >> 3, by = iy >> 3, bz = iz >> 3;
    if (!(leaf_coord_ok(bx) && leaf_coord_ok(by) && leaf_coord_ok(bz))) {
      atomicAdd(&g.ctr->dropped_range, 1ull);
    } else if (shard_owns(bx, g.shard_index, g.shard_count, g.shard_width)) {
      const uint32_t slot = leaf_find_or_insert(g, pack_leaf_key(bx, by, bz), bx, by, free_n0, high_water0, CUDART_INF);
      if (slot != kOverflowSlot) {
        // keep the leaf's lower bound of mark times exact (CAS loop: loaded times are arbitrary doubles, NaN = unknown)
        {
          const double tv = vt[i];
          unsigned long long * tp = reinterpret_cast<unsigned long long *>(&g.leaf_tmin[slot]);
          unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tp);
          for (;;) {
            const double c = __longlong_as_double((long long)cur);
            const double want = (tv == tv) ? fmin(c, tv) : -CUDART_INF;
            if (__double_as_longlong(want) == (long long)cur) break;
            const unsigned long long old = atomicCAS(tp, cur, (unsigned long long)__double_as_longlong(want));
            if (old == cur) break;
            cur = old;
          }
        }
        const unsigned li = (unsigned)(((ix & 7) * 8 + (iy & 7)) * 8 + (iz & 7));
        atomicOr(&g.leaf_mask[(size_t)slot * 16 + (li >> 5)], 1u << (li & 31));
        g.leaf_time[(size_t)slot * kLeafVoxels + li] = vt[i];
      }
    }
  }
}

// =====================================================================================
// Clear sweep — reference .cpp:149-224
// =====================================================================================
constexpr int kSweepThreads = 256;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kDenseLeafVoxels = 33;    // leaves with more than one 32-voxel batch take the CTA-per-leaf path of pass 2

// classify a whole leaf against one frustum: 0 = no voxel centre can be inside, 1 = test per voxel, 2 = all inside
__device__ __forceinline__ int classify_leaf(const DevFrustum & f, int bx, int by, int bz, double vs, double half_vs)
{
  if (f.type < 0) return 0;
  if (!f.cull) return 1;
  const int x0 = bx * 8, y0 = by * 8, z0 = bz * 8;
  if (x0 + 7 < f.lo[0] || x0 > f.hi[0] || y0 + 7 < f.lo[1] || y0 > f.hi[1] || z0 + 7 < f.lo[2] || z0 > f.hi[2]) return 0;
  // bounding sphere of the 512 voxel centres (conservative radius)
  const double cx = ((double)x0 + 3.5) * vs + half_vs, cy = ((double)y0 + 3.5) * vs + half_vs, cz = ((double)z0 + 3.5) * vs + half_vs;
  const double rad = 6.0621778264910705 * vs * 1.000001 + 1e-6;   // 3.5*sqrt(3)
  if (f.type == 0) {
    bool all_in = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double s = f.p[6 * k] * (cx - f.p[6 * k + 3]) + f.p[6 * k + 1] * (cy - f.p[6 * k + 4]) + f.p[6 * k + 2] * (cz - f.p[6 * k + 5]);
      if (s > rad) return 0;
      if (s > -rad) all_in = false;
    }
    return all_in ? 2 : 1;
  }
  // lidar hourglass: radial range and vertical cone, in the sensor frame (approximate transform + margins)
  double tx, ty, tz;
  quat_rotate_exact(f.p[3], f.p[4], f.p[5], f.p[6], cx - f.p[0], cy - f.p[1], cz - f.p[2], tx, ty, tz);
  const double rc = sqrt(tx * tx + ty * ty);
  if (rc - rad > f.aux[1] || rc + rad < f.aux[0]) return 0;
  const double zmin = fabs(tz) - rad;
  if (zmin > 0 && zmin - f.aux[3] > f.aux[2] * (rc + rad) + 1e-6) return 0;
  return 1;
}

__device__ __forceinline__ void free_leaf(const GridView & g, uint32_t slot)
{
  const uint32_t h = g.leaf_hpos[slot];
  g.hash[h].key = kTombKey;
  g.hash[h].val = kInvalidSlot;
  g.leaf_key[slot] = kEmptyKey;
  const uint32_t i = atomicAdd(&g.ctr->free_n, 1u);
  g.free_slots[i] = slot;
  atomicAdd(&g.ctr->n_tombs, 1u);
}

// block-level reduction of the per-thread sweep statistics into the device counters
struct SweepAcc {
  unsigned int swept = 0, surv = 0, clear = 0, rewr = 0, amb = 0;
  int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
};
__device__ __forceinline__ void sweep_publish(const GridView & g, SweepAcc a)
{
  __shared__ unsigned int s_red[5];
  __shared__ int s_bb[4];
  if (threadIdx.x == 0) { s_red[0] = s_red[1] = s_red[2] = s_red[3] = s_red[4] = 0; s_bb[0] = s_bb[1] = INT_MAX; s_bb[2] = s_bb[3] = INT_MIN; }
  __syncthreads();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a.swept += __shfl_xor_sync(kFull, a.swept, d); a.surv += __shfl_xor_sync(kFull, a.surv, d);
    a.clear += __shfl_xor_sync(kFull, a.clear, d); a.rewr += __shfl_xor_sync(kFull, a.rewr, d);
    a.amb += __shfl_xor_sync(kFull, a.amb, d);
    a.min_x = min(a.min_x, __shfl_xor_sync(kFull, a.min_x, d)); a.min_y = min(a.min_y, __shfl_xor_sync(kFull, a.min_y, d));
    a.max_x = max(a.max_x, __shfl_xor_sync(kFull, a.max_x, d)); a.max_y = max(a.max_y, __shfl_xor_sync(kFull, a.max_y, d));
  }
  if ((threadIdx.x & 31) == 0 && (a.swept | a.surv | a.clear | a.rewr | a.amb)) {
    atomicAdd(&s_red[0], a.swept); atomicAdd(&s_red[1], a.surv); atomicAdd(&s_red[2], a.clear);
    atomicAdd(&s_red[3], a.rewr); atomicAdd(&s_red[4], a.amb);
    atomicMin(&s_bb[0], a.min_x); atomicMin(&s_bb[1], a.min_y); atomicMax(&s_bb[2], a.max_x); atomicMax(&s_bb[3], a.max_y);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    SweepCtr * c = &g.ctr->sw[g.sw_cur];
    if (s_red[0]) atomicAdd(&c->n_swept, (unsigned long long)s_red[0]);
    if (s_red[1]) atomicAdd(&c->n_survived, (unsigned long long)s_red[1]);
    if (s_red[2]) {
      atomicAdd(&c->n_cleared, (unsigned long long)s_red[2]);
      atomicMin(&c->cb_min_x, s_bb[0]); atomicMin(&c->cb_min_y, s_bb[1]);
      atomicMax(&c->cb_max_x, s_bb[2]); atomicMax(&c->cb_max_y, s_bb[3]);
    }
    if (s_red[3]) atomicAdd(&c->n_rewritten, (unsigned long long)s_red[3]);
    if (s_red[4]) atomicAdd(&c->n_ambiguous, (unsigned long long)s_red[4]);
  }
}

__device__ __forceinline__ void stage_frustums(DevFrustum * fr, const DevFrustum * frustums_gmem, int n_frustums)
{
  // the frustum parameter blocks (plane sets / hourglass parameters) live in shared memory for the whole kernel
  const int words = n_frustums * (int)(sizeof(DevFrustum) / 8);
  const double * src = reinterpret_cast<const double *>(frustums_gmem);
  double * dst = reinterpret_cast<double *>(fr);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// can a leaf whose oldest mark time is tmin contain a voxel that expires by decay alone?  The decay duration is
// non-increasing in a voxel's age for the linear model; the exponential / persistent ones never go negative for
// voxel_decay >= 0 (reference .cpp:320-331).  tmin = -inf ("unknown") always answers yes.
__device__ __forceinline__ bool leaf_may_expire(const GridView & g, double now, double tmin)
{
  if (g.decay_model == 0) return __dsub_rn(g.voxel_decay, __dsub_rn(now, tmin)) < 0.;
  return !(g.voxel_decay >= 0.0);
}

// Sweep, pass 1 — ONE THREAD PER LEAF, coalesced streaming over the dense per-leaf arrays (key, tmin, tile, 64 B value
// mask).  A leaf outside every frustum's bounding box whose oldest voxel is too young to expire cannot change: its
// survivors are counted straight from the mask and its 4 KB time block is never touched.  Every other leaf goes to
// the global work queue of pass 2.
__global__ void __launch_bounds__(kSweepThreads)
sweep_triage_kernel(const GridView g, const __grid_constant__ FrustumParams fparam, const DevFrustum * __restrict__ frustums_gmem,
                    const int n_frustums, const double now, const int force_full, uint8_t * __restrict__ cm_reset,
                    const unsigned long long cm_bytes, const int cm_value, const unsigned dense_threshold)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevFrustum * fr = reinterpret_cast<DevFrustum *>(smem_raw);
  TRACE(1, 0);
  pdl_launch_dependents();   // the leaf pass may become resident now; it waits for this grid before reading the queue
  // frustum blocks: from the kernel parameters (small sets) or from the device buffer an earlier stream operation
  // filled; neither is written by the predecessor kernel, so they are staged before waiting for it
  stage_frustums(fr, frustums_gmem ? frustums_gmem : fparam.f, n_frustums);
  pdl_wait();                // launched programmatically behind the previous cycle's Mark kernel: the grid is ours from here
  const uint32_t n_slots = g.ctr->high_water;
  TRACE(1, 1);
  // re-arm the state of the NEXT sweep while this one runs: zero the other tile buffer and the other counter set
  {
    const uint32_t n_words = (g.ctr->tile_n < g.tile_capacity ? g.ctr->tile_n : g.tile_capacity) * 64u;
    uint4 * t4 = reinterpret_cast<uint4 *>(g.tile_count_next);
    for (uint32_t i = blockIdx.x * kSweepThreads + threadIdx.x; i < n_words / 4; i += gridDim.x * kSweepThreads) t4[i] = make_uint4(0, 0, 0, 0);
    // fused update cycle: Costmap2D::resetMaps (spatio_temporal_voxel_layer.cpp:705) for the costmap pass that rides on
    // this cycle's Mark kernel — saves the memset launch between Mark and the costmap
    if (cm_reset) {
      const unsigned head = (unsigned)((16u - (unsigned)(reinterpret_cast<uintptr_t>(cm_reset) & 15u)) & 15u);
      const unsigned long long h = head < cm_bytes ? head : cm_bytes;
      const unsigned long long n16 = (cm_bytes - h) / 16;
      const unsigned v = (unsigned)cm_value * 0x01010101u;
      uint4 * c4 = reinterpret_cast<uint4 *>(cm_reset + h);
      for (unsigned long long i = (unsigned long long)blockIdx.x * kSweepThreads + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * kSweepThreads)
        c4[i] = make_uint4(v, v, v, v);
      if (blockIdx.x == 0) {
        for (unsigned long long i = threadIdx.x; i < h; i += kSweepThreads) cm_reset[i] = (uint8_t)cm_value;
        for (unsigned long long i = h + n16 * 16 + threadIdx.x; i < cm_bytes; i += kSweepThreads) cm_reset[i] = (uint8_t)cm_value;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      SweepCtr * c = &g.ctr->sw[g.sw_cur ^ 1];
      c->n_swept = c->n_survived = c->n_cleared = c->n_rewritten = c->n_ambiguous = 0;
      c->cb_min_x = c->cb_min_y = INT_MAX; c->cb_max_x = c->cb_max_y = INT_MIN;
      c->n_points_out = c->n_cleared_out = c->n_ambiguous_out = 0;
    }
  }
  SweepAcc acc;
  TRACE(1, 2);
  for (uint32_t slot = blockIdx.x * kSweepThreads + threadIdx.x; slot < n_slots; slot += gridDim.x * kSweepThreads) {
    // all loads of the leaf are independent of each other: one DRAM round trip per leaf
    const uint64_t key = g.leaf_key[slot];
    const double tmin = g.leaf_tmin[slot];
    const uint32_t tile = g.leaf_tile[slot];
    const uint4 * mp = reinterpret_cast<const uint4 *>(g.leaf_mask + (size_t)slot * 16);
    uint4 mw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mw[q] = mp[q];
    if (key == kEmptyKey) continue;   // free slot
    uint32_t any = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) any |= mw[q].x | mw[q].y | mw[q].z | mw[q].w;
    if (any == 0) { free_leaf(g, slot); continue; }   // nothing left in this leaf: pruneGrid, .cpp:223
    bool full_path = force_full || leaf_may_expire(g, now, tmin);
    if (!full_path) {
      int bx, by, bz;
      unpack_leaf_key(key, bx, by, bz);
      const int x0 = bx * 8, y0 = by * 8, z0 = bz * 8;
      for (int f = 0; f < n_frustums && !full_path; ++f) {
        const DevFrustum & F = fr[f];
        if (F.type < 0) continue;
        if (F.cull && (x0 + 7 < F.lo[0] || x0 > F.hi[0] || y0 + 7 < F.lo[1] || y0 > F.hi[1] || z0 + 7 < F.lo[2] || z0 > F.hi[2])) continue;
        // inside the frustum's bounding box: bounding sphere of the leaf against the plane set / hourglass
        full_path = classify_leaf(F, bx, by, bz, g.vs, g.half_vs) != 0;
      }
    }
    if (full_path) {
      // leaves with more than one 32-voxel batch are processed by a whole CTA (the per-voxel pass of a leaf is a long
      // dependent chain of exact double arithmetic: ~2 us per batch, the critical path of the whole sweep when one
      // warp walks 4-16 batches).  They are queued from the end of the queue array.
      unsigned pc = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) pc += __popc(mw[q].x) + __popc(mw[q].y) + __popc(mw[q].z) + __popc(mw[q].w);
      // warp-aggregated queue pushes: one same-address atomic per warp and queue instead of one per leaf (several
      // hundred returning atomics on one counter serialise for microseconds at the end of the kernel)
      const bool is_dense = pc >= dense_threshold;
      const unsigned peers = __match_any_sync(__activemask(), is_dense ? 1 : 0);
      const unsigned lane_id = threadIdx.x & 31;
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if ((int)lane_id == leader) base = atomicAdd(is_dense ? &g.ctr->dense_n : &g.ctr->queue_n, (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      const uint32_t qpos = base + (uint32_t)__popc(peers & ((1u << lane_id) - 1u));
      g.sweep_queue[is_dense ? g.leaf_capacity - 1u - qpos : qpos] = slot;
      continue;
    }
    // PopulateCostmapAndPointcloud .cpp:242-250 for every (surviving) voxel: per-column popcounts, two adjacent u32
    // column counters per 64-bit reduction (no carry: counts stay < 2^32)
    unsigned cnt = 0;
    unsigned long long * tc = reinterpret_cast<unsigned long long *>(g.tile_count + (size_t)tile * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t wv[4] = {mw[q].x, mw[q].y, mw[q].z, mw[q].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t x = wv[j];
        if (!x) continue;
        cnt += __popc(x);
        x = x - ((x >> 1) & 0x55555555u);
        x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
        x = (x + (x >> 4)) & 0x0F0F0F0Fu;   // four per-column counts, one per byte
        if (tile != kOverflowSlot) {
          const unsigned long long lo = (unsigned long long)(x & 0xFFu) | ((unsigned long long)((x >> 8) & 0xFFu) << 32);
          const unsigned long long hi = (unsigned long long)((x >> 16) & 0xFFu) | ((unsigned long long)(x >> 24) << 32);
          if (lo) atomicAdd(tc + (q * 4 + j) * 2, lo);
          if (hi) atomicAdd(tc + (q * 4 + j) * 2 + 1, hi);
        }
      }
    }
    acc.swept += cnt; acc.surv += cnt;
  }
  __syncthreads();
  TRACE(1, 3);
  sweep_publish(g, acc);
  TRACE(1, 4);
}

// every CTA of pass 2 read queue_n when it started; the last one to finish empties the queue for the next sweep
__device__ __forceinline__ void leaf_pass_done(const GridView & g)
{
  __threadfence();
  if (atomicAdd(&g.ctr->leaf_ticket, 1u) == gridDim.x - 1) {
    g.ctr->queue_n = 0;
    g.ctr->dense_n = 0;
    g.ctr->leaf_ticket = 0;
    __threadfence();
  }
}

// Sweep, pass 2 — ONE WARP PER QUEUED LEAF (CTAs >= n_dense_ctas) or ONE CTA PER DENSE LEAF, each of its eight warps
// taking every eighth 32-voxel batch of the leaf's compacted voxel list (CTAs < n_dense_ctas): leaf-level frustum classification, exact per-voxel frustum tests,
// decay, rewrite of accelerated mark times, clearing (reference .cpp:158-221).
__global__ void __launch_bounds__(kSweepThreads, 4)
sweep_leaf_kernel(const GridView g, const __grid_constant__ FrustumParams fparam, const DevFrustum * __restrict__ frustums_gmem,
                  const int n_frustums, const double now, const SweepOutputs out, const uint32_t n_dense_ctas)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevFrustum * fr = reinterpret_cast<DevFrustum *>(smem_raw);
  __shared__ uint16_t s_list[kSweepWarps][kLeafVoxels];
  __shared__ uint32_t s_mask[2][kSweepWarps][16];   // [iteration parity]: warp 0 of a dense leaf's CTA reads its siblings' copies
  __shared__ uint32_t s_part[kSweepWarps][kMaxFrustums / 32];
  __shared__ uint32_t s_full[kSweepWarps][kMaxFrustums / 32];
  __shared__ double s_val[kSweepWarps][128];
  __shared__ double s_pmin[2][kSweepWarps];   // dense leaves: [iteration parity][warp] smallest surviving time of the warp's batches
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const bool dense = blockIdx.x < n_dense_ctas;
  const uint16_t * mask16 = reinterpret_cast<const uint16_t *>(g.leaf_mask);
  TRACE(2, 0);
  // Programmatic dependent launch: this grid becomes resident while the triage pass is still running and lets Mark do
  // the same.  The frustum blocks were uploaded before the triage pass started, so they are staged before the wait.
  pdl_launch_dependents();
  stage_frustums(fr, frustums_gmem ? frustums_gmem : fparam.f, n_frustums);
  pdl_wait();
  TRACE(2, 1);
  // work items: dense CTA b takes dense leaf number b (stored from the end of the queue array), the others queue
  // entry (blockIdx - n_dense_ctas) * 8 + warp.  When there are more dense leaves than the dense CTAs can take in two
  // rounds (large maps whose leaves all need the per-voxel pass) throughput matters, not the latency of one leaf: the
  // dense CTAs then also work one leaf per warp (coop == false).
  // Shorten the dependent-load chain of a warp's first leaf: its queue entry and leaf header are requested
  // speculatively (assuming the cooperative mode for dense CTAs), together with the queue length; a stale queue entry
  // is still a valid slot index, so wasted loads are harmless.
  const uint32_t qi_spec = dense ? blockIdx.x : (blockIdx.x - n_dense_ctas) * kSweepWarps + warp;
  uint32_t slot0 = qi_spec < g.leaf_capacity ? g.sweep_queue[dense ? g.leaf_capacity - 1u - qi_spec : qi_spec] : 0u;
  if (slot0 >= g.leaf_capacity) slot0 = 0;
  const uint64_t key0 = g.leaf_key[slot0];
  const uint32_t m16_0 = mask16[(size_t)slot0 * 32 + lane];
  const uint32_t tile0 = g.leaf_tile[slot0];
  const double tmin0 = g.leaf_tmin[slot0];
  const uint32_t n_queued = dense ? g.ctr->dense_n : g.ctr->queue_n;
  const bool coop = dense && n_queued <= 2u * n_dense_ctas;
  const uint32_t qi0 = coop ? blockIdx.x : (dense ? blockIdx.x : blockIdx.x - n_dense_ctas) * kSweepWarps + warp;
  const uint32_t q_stride = coop ? n_dense_ctas : (dense ? n_dense_ctas : gridDim.x - n_dense_ctas) * kSweepWarps;
  TRACE(2, 2);
  if ((coop ? blockIdx.x : (dense ? blockIdx.x : blockIdx.x - n_dense_ctas) * kSweepWarps) >= n_queued) {   // nothing for this CTA (uniform per block)
    if (threadIdx.x == 0) leaf_pass_done(g);
    return;
  }

  const int n_chunks = (n_frustums + 31) >> 5;
  const bool skip_model_ok = !out.points;
  unsigned int c_swept = 0, c_surv = 0, c_clear = 0, c_rewr = 0, c_amb = 0;
  int bb_min_x = INT_MAX, bb_min_y = INT_MAX, bb_max_x = INT_MIN, bb_max_y = INT_MIN;

  uint32_t iter = 0;
  for (uint32_t qi = qi0; qi < n_queued; qi += q_stride, ++iter) {
    const bool first = qi == qi_spec && qi == qi0;   // the speculative loads above are this leaf's
    const uint32_t slot = first ? slot0 : g.sweep_queue[dense ? g.leaf_capacity - 1u - qi : qi];
    const uint64_t key = first ? key0 : g.leaf_key[slot];
    const uint32_t m16 = first ? m16_0 : mask16[(size_t)slot * 32 + lane];   // each lane owns two voxel columns = local indices [lane*16, lane*16+16)
    uint32_t (* const my_mask)[16] = s_mask[iter & 1];
    const uint32_t tile = first ? tile0 : g.leaf_tile[slot];
    const double tmin = first ? tmin0 : g.leaf_tmin[slot];

    int bx, by, bz;
    unpack_leaf_key(key, bx, by, bz);
    // leaf-level culling: one lane per frustum
    unsigned any_cand = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const int f = c * 32 + (int)lane;
      const int cls = f < n_frustums ? classify_leaf(fr[f], bx, by, bz, g.vs, g.half_vs) : 0;
      const unsigned pm = __ballot_sync(kFull, cls == 1), fm = __ballot_sync(kFull, cls == 2);
      if (lane == 0) { s_part[warp][c] = pm; s_full[warp][c] = fm; }
      any_cand |= pm | fm;
    }
    // temporal culling: outside every frustum and too young to expire -> the 4 KB time block is never touched
    if (!any_cand && skip_model_ok) {
      if (!leaf_may_expire(g, now, tmin)) {
        if (coop && warp != 0) continue;   // (all warps of a dense leaf's CTA get here: one of them counts)
        const unsigned cnt = __popc(m16);
        c_swept += cnt; c_surv += cnt;
        if (tile != kOverflowSlot) {
          const unsigned long long c0 = __popc(m16 & 0xFFu), c1 = __popc(m16 >> 8);
          if (c0 | c1) atomicAdd(reinterpret_cast<unsigned long long *>(g.tile_count + (size_t)tile * 64 + 2 * lane), c0 | (c1 << 32));
        }
        continue;
      }
    }
    // redistribute the active voxels evenly over the lanes: per-lane popcount, warp scan, index list in smem
    const int cnt = __popc(m16);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFull, incl, d); if ((int)lane >= d) incl += v; }
    const int total = __shfl_sync(kFull, incl, 31);
    {
      int pos = incl - cnt;
      uint32_t m = m16;
      while (m) { const int b = __ffs(m) - 1; m &= m - 1; s_list[warp][pos++] = (uint16_t)(lane * 16 + b); }
    }
    // the 16 mask words of the leaf (lane j < 16 holds word j), assembled from the 16-bit halves by shuffles
    uint32_t w_orig;
    {
      const uint32_t lo = __shfl_sync(kFull, m16, (2 * lane) & 31), hi = __shfl_sync(kFull, m16, (2 * lane + 1) & 31);
      w_orig = lo | (hi << 16);
    }
    if (lane < 16) my_mask[warp][lane] = w_orig;

    __syncwarp();

    const double * tbase = g.leaf_time + (size_t)slot * kLeafVoxels;
    double vmin = CUDART_INF;   // smallest mark time among this leaf's survivors
    // 32-voxel batches of the compacted list: all of them (one warp per leaf) or every eighth (dense leaf, CTA per leaf)
    const int bstep = coop ? kSweepWarps : 1;
    int nb = 0;
    for (int i0 = (coop ? (int)warp : 0) * 32; i0 < total; i0 += 32 * bstep, ++nb) {
      // the mark times of this warp's next four batches are requested together (one DRAM round trip per 128 voxels)
      if ((nb & 3) == 0) {
        double pv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = i0 + u * 32 * bstep + (int)lane;
          pv[u] = j < total ? __ldcs(tbase + s_list[warp][j]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s_val[warp][u * 32 + lane] = pv[u];
        __syncwarp();
      }
      const int i = i0 + (int)lane;
      const bool have = i < total;
      bool cleared = false, survived = false;
      int ix = 0, iy = 0, iz = 0;
      unsigned li = 0;
      if (have) {
        li = s_list[warp][i];
        const double v = s_val[warp][(nb & 3) * 32 + lane];
        ix = bx * 8 + (int)(li >> 6); iy = by * 8 + (int)((li >> 3) & 7); iz = bz * 8 + (int)(li & 7);
        ++c_swept;
        // .cpp:167-169
        const double t = __dsub_rn(now, v);
        double base_dur;
        bool amb = false;
        if (g.decay_model == 0) base_dur = __dsub_rn(g.voxel_decay, t);
        else if (g.decay_model == 1) base_dur = __dmul_rn(g.voxel_decay, exp(-t));
        else base_dur = g.voxel_decay;
        // .cpp:171-198: first containing frustum, in reading order, decides
        int hit = -1;
        if (any_cand) {
          const double wx = index_to_world(ix, g.vs, g.half_vs), wy = index_to_world(iy, g.vs, g.half_vs), wz = index_to_world(iz, g.vs, g.half_vs);
          for (int c = 0; c < n_chunks && hit < 0; ++c) {
            const unsigned fm = s_full[warp][c];
            unsigned m = s_part[warp][c] | fm;
            while (m) {
              const int b = __ffs(m) - 1; m &= m - 1;
              const int f = c * 32 + b;
              bool in;
              if ((fm >> b) & 1u) in = true;
              else if (fr[f].type == 0) in = depth_is_inside(fr[f].p, wx, wy, wz);
              else in = lidar_is_inside(fr[f].p, wx, wy, wz, amb);
              if (in) { hit = f; break; }
            }
          }
        }
        if (hit >= 0) {
          // .cpp:179-197; GetFrustumAcceleration .cpp:334-341 = ((1./6.)*factor) * ((t*t)*t)
          const double accel = __dmul_rn(fr[hit].c16f, __dmul_rn(__dmul_rn(t, t), t));
          const double tud = __dsub_rn(base_dur, accel);
          if (g.decay_model == 1 && fabs(tud) <= 1e-12 * fmax(fabs(base_dur), fabs(accel))) amb = true;
          if (tud < 0.) cleared = true;
          else {
            const double nv = __dsub_rn(v, accel);
            if (__double_as_longlong(nv) != __double_as_longlong(v)) { g.leaf_time[(size_t)slot * kLeafVoxels + li] = nv; ++c_rewr; }
            vmin = fmin(vmin, nv);
          }
        } else if (base_dur < 0.) {
          cleared = true;   // .cpp:203-211
        } else {
          vmin = fmin(vmin, v);
        }
        survived = !cleared;
        if (cleared) {
          atomicAnd(&my_mask[warp][li >> 5], ~(1u << (li & 31)));
          ++c_clear;
          bb_min_x = min(bb_min_x, ix); bb_max_x = max(bb_max_x, ix);
          bb_min_y = min(bb_min_y, iy); bb_max_y = max(bb_max_y, iy);
        } else {
          ++c_surv;
        }
        if (amb) {
          ++c_amb;
          if (out.ambiguous) {
            const unsigned long long a = atomicAdd(&g.ctr->sw[g.sw_cur].n_ambiguous_out, 1ull);
            if (a < out.ambiguous_cap) { out.ambiguous[3 * a] = ix; out.ambiguous[3 * a + 1] = iy; out.ambiguous[3 * a + 2] = iz; }
          }
        }
      }
      // optional lists: warp-aggregated appends
      if (out.points) {   // PopulateCostmapAndPointcloud .cpp:234-240: Point32 = (float) voxel centre
        const unsigned sm = __ballot_sync(kFull, survived);
        if (sm) {
          unsigned long long b0 = 0;
          if (lane == (unsigned)(__ffs(sm) - 1)) b0 = atomicAdd(&g.ctr->sw[g.sw_cur].n_points_out, (unsigned long long)__popc(sm));
          b0 = __shfl_sync(kFull, b0, __ffs(sm) - 1);
          if (survived) {
            const unsigned long long o = b0 + __popc(sm & ((1u << lane) - 1));
            if (o < out.points_cap) {
              out.points[3 * o] = (float)index_to_world(ix, g.vs, g.half_vs);
              out.points[3 * o + 1] = (float)index_to_world(iy, g.vs, g.half_vs);
              out.points[3 * o + 2] = (float)index_to_world(iz, g.vs, g.half_vs);
            }
          }
        }
      }
      if (out.cleared) {   // cleared_cells.insert, .cpp:213-215
        const unsigned cm = __ballot_sync(kFull, cleared);
        if (cm) {
          unsigned long long b0 = 0;
          if (lane == (unsigned)(__ffs(cm) - 1)) b0 = atomicAdd(&g.ctr->sw[g.sw_cur].n_cleared_out, (unsigned long long)__popc(cm));
          b0 = __shfl_sync(kFull, b0, __ffs(cm) - 1);
          if (cleared) {
            const unsigned long long o = b0 + __popc(cm & ((1u << lane) - 1));
            if (o < out.cleared_cap) { out.cleared[2 * o] = ix; out.cleared[2 * o + 1] = iy; }
          }
        }
      }
    }
    __syncwarp();
    // write back changed mask words; count survivors per (x,y) column into the leaf's column tile
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vmin = fmin(vmin, __shfl_xor_sync(kFull, vmin, d));
    if (coop) {
      // The warps of a dense leaf's CTA each cleared bits in their own copy of the mask: they meet at a barrier (they
      // all run the same iterations) and warp 0 combines the copies and takes the leaf-level decisions.  Everything
      // shared is double-buffered by iteration parity: a warp can only write the same buffer again after the NEXT
      // barrier, which warp 0 reaches after it has read this one.
      if (lane == 0) s_pmin[iter & 1][warp] = vmin;
      __syncthreads();
      if (warp != 0) continue;
#pragma unroll
      for (int k = 1; k < kSweepWarps; ++k) vmin = fmin(vmin, s_pmin[iter & 1][k]);
      if (lane < 16) {
        uint32_t a = my_mask[0][lane];
#pragma unroll
        for (int k = 1; k < kSweepWarps; ++k) a &= my_mask[k][lane];
        my_mask[0][lane] = a;
      }
      __syncwarp();
    }
    uint32_t w = 0;
    if (lane < 16) {
      w = my_mask[warp][lane];
      if (w != w_orig) g.leaf_mask[(size_t)slot * 16 + lane] = w;
    }
    const bool leaf_alive = __ballot_sync(kFull, w != 0) != 0;
    if (leaf_alive && lane == 0 && !(vmin == tmin)) g.leaf_tmin[slot] = vmin;   // exact lower bound again
    if (!leaf_alive) {
      if (lane == 0) free_leaf(g, slot);
    } else if (tile != kOverflowSlot) {
      const uint32_t two = reinterpret_cast<const uint16_t *>(my_mask[warp])[lane];
      const unsigned long long c0 = __popc(two & 0xFFu), c1 = __popc(two >> 8);
      // PopulateCostmapAndPointcloud .cpp:242-250: _cost_map[(x,y)] += 1 per surviving voxel.  The two adjacent
      // u32 column counters of this lane are bumped by ONE 64-bit reduction (no carry: counts stay < 2^32)
      if (c0 | c1) atomicAdd(reinterpret_cast<unsigned long long *>(g.tile_count + (size_t)tile * 64 + 2 * lane), c0 | (c1 << 32));
    }
    __syncwarp();
  }   // queued leaves

  SweepAcc acc;
  acc.swept = c_swept; acc.surv = c_surv; acc.clear = c_clear; acc.rewr = c_rewr; acc.amb = c_amb;
  acc.min_x = bb_min_x; acc.min_y = bb_min_y; acc.max_x = bb_max_x; acc.max_y = bb_max_y;
  TRACE_VAL(2, 5, iter);       // leaves this warp went through
  TRACE_VAL(2, 6, c_swept);    // voxels lane 0 of warp 0 tested
  __syncthreads();
  TRACE(2, 3);
  sweep_publish(g, acc);
  if (threadIdx.x == 0) leaf_pass_done(g);
  TRACE(2, 4);
}

// =====================================================================================
// Costmap kernel (helpers: see above the Mark kernel, which can carry the costmap pass in the fused cycle)
// =====================================================================================
constexpr int kCostmapTilesPerWarp = 8;
__global__ void __launch_bounds__(256) costmap_kernel(const GridView g, const CostmapDev p, uint8_t * __restrict__ costmap)
{
  const unsigned lane = threadIdx.x & 31;
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
    costmap_tile_batch<kCostmapTilesPerWarp>(g, p, costmap, t0, n_tiles, lane, n, mnx, mny, mxx, mxy);
  }
  __syncthreads();
  TRACE(3, 2);
  costmap_reduce_and_publish(g.ctr, n, mnx, mny, mxx, mxy);
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
  costmap_reduce_and_publish(g.ctr, n, mnx, mny, mxx, mxy);
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
  costmap_reduce_and_publish(g.ctr, n, mnx, mny, mxx, mxy);
}

// =====================================================================================
// ResetGridArea — reference .cpp:397-418 (x,y only: whole voxel columns are cleared)
// =====================================================================================
__global__ void __launch_bounds__(256) reset_area_kernel(const GridView g, double sx, double sy, double ex, double ey, int invert)
{
  const unsigned lane = threadIdx.x & 31;
  const uint32_t n_slots = g.ctr->high_water;
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
  const uint32_t n_slots = g.ctr->high_water;
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
  const uint32_t n_slots = g.ctr->high_water;
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
  const uint32_t n_slots = g.ctr->high_water;
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
  const uint32_t n_slots = g.ctr->high_water;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += gridDim.x * blockDim.x) {
    const uint64_t key = g.leaf_key[slot];
    if (key == kEmptyKey) continue;
    int bx, by, bz;
    unpack_leaf_key(key, bx, by, bz);
    g.leaf_tile[slot] = tile_find_or_insert(g, bx, by);
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

cudaError_t launch_fold_alloc(const GridView & g, cudaStream_t s)
{
  fold_alloc_kernel<<<1, 1, 0, s>>>(g);
  return cudaGetLastError();
}
// launch with (optionally) the programmatic-stream-serialization attribute, see pdl_wait()
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool pdl, Args &&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

template <bool kVec4, bool kCostmap>
static cudaError_t launch_mark_variant(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                                       const CostmapDev & cm, uint8_t * costmap)
{
  static int occ = 0;
  const size_t smem = kVec4 ? kMarkStageBytes : 0;
  if (occ == 0) {
    int n = 0;
    if (kVec4) cudaFuncSetAttribute(mark_kernel<kVec4, kCostmap>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMarkStageBytes);
    // All kernels of the cycle ask for the same (maximum) shared-memory carve-out: they overlap on the same SMs under
    // programmatic dependent launch, and an SM only changes its L1 / shared split when it is empty — with the default
    // split of the sweep kernels only two of Mark's 43 KB CTAs fit next to them, and they would keep the SM from ever
    // draining.
    cudaFuncSetAttribute(mark_kernel<kVec4, kCostmap>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mark_kernel<kVec4, kCostmap>, kMarkThreads, smem);
    occ = (e == cudaSuccess && n >= 1) ? n : 4;
  }
  // persistent CTAs, as many as are resident at once (never more than there are chunk pairs): each pulls chunks from
  // the device-side cursor with the next chunk's loads in flight
  const uint32_t resident = (uint32_t)sm_count * (uint32_t)occ;
  const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(resident, 65535u), (batch.total_blocks + 1) / 2));
  return launch_ex(mark_kernel<kVec4, kCostmap>, grid, kMarkThreads, smem, s, pdl, g, batch, cm, costmap);
}
// pdl: the caller guarantees that the kernel's inputs (clouds) were complete before the PREDECESSOR of this launch in the
// stream started, and that the predecessor is one of this library's sweep kernels.  cm / costmap != NULL: fused costmap pass.
cudaError_t launch_mark(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                        const CostmapDev * cm, uint8_t * costmap)
{
  if (batch.total_blocks == 0) return cudaSuccess;
  bool all_vec4 = true;
  for (uint32_t i = 0; i < batch.n_readings; ++i) all_vec4 = all_vec4 && batch.r[i].vec4 != 0;
  const CostmapDev none{};
  if (cm && costmap) {
    return all_vec4 ? launch_mark_variant<true, true>(g, batch, sm_count, s, pdl, *cm, costmap)
                    : launch_mark_variant<false, true>(g, batch, sm_count, s, pdl, *cm, costmap);
  }
  return all_vec4 ? launch_mark_variant<true, false>(g, batch, sm_count, s, pdl, none, nullptr)
                  : launch_mark_variant<false, false>(g, batch, sm_count, s, pdl, none, nullptr);
}
uint32_t mark_grid_for(const MarkBatch & batch, int sm_count)
{
  // (launch_mark_variant's grid, with the nominal 5 resident CTAs per SM)
  return std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>((uint32_t)sm_count * 5u, 65535u), (batch.total_blocks + 1) / 2));
}
uint32_t mark_blocks_for(unsigned long long n_points)
{
  return (uint32_t)((n_points + kMarkPointsPerBlock - 1) / kMarkPointsPerBlock);
}
cudaError_t launch_load_voxels(const GridView & g, const int32_t * x, const int32_t * y, const int32_t * z, const double * t,
                               unsigned long long n, cudaStream_t s)
{
  if (n == 0) return cudaSuccess;
  load_voxels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, x, y, z, t, n);
  return cudaGetLastError();
}
// frustums_dev == NULL: the (<= kParamFrustums) blocks travel in *fparam
cudaError_t launch_sweep(const GridView & g, const DevFrustum * frustums_dev, const FrustumParams * fparam, int n_frustums, double now,
                         const SweepOutputs & out, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s, uint8_t * costmap_reset,
                         size_t costmap_bytes, int costmap_value)
{
  static const FrustumParams no_params{};
  const FrustumParams & fp = (frustums_dev == nullptr && fparam) ? *fparam : no_params;
  cudaError_t e;
  static bool carveout_set = false;
  if (!carveout_set) {   // see launch_mark_variant
    cudaFuncSetAttribute(sweep_triage_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(sweep_leaf_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  const size_t dyn = (size_t)n_frustums * sizeof(DevFrustum);
  if (dyn > 32 * 1024) {
    e = cudaFuncSetAttribute(sweep_triage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sweep_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  // pass 1: one thread per leaf slot.  pub_voxels needs every survivor's coordinates: everything takes pass 2 then.
  const int force_full = out.points ? 1 : 0;
  // Both passes are launched programmatically: everything they do before pdl_wait() only touches kernel parameters /
  // the frustum buffer, so the attribute is safe whatever precedes them on the stream.
  // Grid sizes: every kernel of the cycle is grid-stride, and under programmatic dependent launch a kernel only starts
  // once ALL CTAs of its predecessor are resident — so the grids are kept to a fraction of a wave (the leaf pass and
  // Mark get their turn early) unless the pool is large enough to need the extra memory parallelism.
  const bool big = leaf_slots_upper_bound > (1u << 18);
  const int triage_cap = sm_count * (big ? 8 : 2);
  e = launch_ex(sweep_triage_kernel, (unsigned)grid_for(leaf_slots_upper_bound, kSweepThreads, triage_cap), kSweepThreads, dyn, s, true,
                g, fp, frustums_dev, n_frustums, now, force_full, costmap_reset, (unsigned long long)costmap_bytes, costmap_value,
                big ? 0xFFFFFFFFu : (unsigned)kDenseLeafVoxels);
  if (e != cudaSuccess) return e;
  // pass 2: one warp per queued leaf; the queue length is only known on the device, CTAs without work exit at once.
  // Programmatic dependent launch: its CTAs become resident (and stage the frustum blocks) while pass 1 drains.
  // The first CTAs take the leaves with more than one 32-voxel batch (a CTA per leaf), the others one leaf per warp.
  // Large pools are throughput bound: no CTA-per-leaf mode there, every queued leaf gets a warp.
  const unsigned n_dense_ctas = big ? 0u : (unsigned)grid_for(leaf_slots_upper_bound, 1, sm_count * 2);
  const unsigned n_sparse_ctas = (unsigned)grid_for(leaf_slots_upper_bound, kSweepWarps, sm_count * (big ? 4 : 2));
  return launch_ex(sweep_leaf_kernel, n_dense_ctas + n_sparse_ctas, kSweepThreads, dyn, s, true,
                   g, fp, frustums_dev, n_frustums, now, out, (uint32_t)n_dense_ctas);
}
cudaError_t launch_costmap(const GridView & g, const CostmapDev & p, uint8_t * costmap_dev, uint32_t tiles_upper_bound, int sm_count, cudaStream_t s,
                           bool fill_default, bool pdl)
{
  if (fill_default) {
    const cudaError_t e = cudaMemsetAsync(costmap_dev, p.default_value, (size_t)p.size_x * p.size_y, s);   // Costmap2D::resetMaps, :705
    if (e != cudaSuccess) return e;
  }
  // 64 tiles per CTA (8 per warp): the per-CTA statistics reduction is amortised and the grid stays a fraction of a wave
  return launch_ex(costmap_kernel, (unsigned)grid_for(tiles_upper_bound, 64, sm_count * 4), 256u, 0, s, pdl && !fill_default, g, p, costmap_dev);
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
cudaError_t launch_reassign_tiles(const GridView & g, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s)
{
  reassign_tiles_kernel<<<grid_for(leaf_slots_upper_bound, 256, sm_count * 8), 256, 0, s>>>(g);
  return cudaGetLastError();
}

}  // namespace stvl
