// stvl_common.cuh — device helpers shared by the kernels of the per-cycle voxel update: PTX wrappers (reductions,
// programmatic dependent launch, mbarrier / bulk copy), the leaf hash and the fold-free leaf allocator, the costmap tile pass.
#pragma once
#include <math_constants.h>
#include <climits>
#include <cstdint>

#include "stvl_device.cuh"
#include "stvl_kernels.h"

namespace stvl {

constexpr unsigned kFull = 0xFFFFFFFFu;

// fire-and-forget global reductions (no return value, no scoreboard wait)
__device__ __forceinline__ void red_or_u32(uint32_t * p, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_and_u32(uint32_t * p, uint32_t v) { asm volatile("red.global.and.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add_u32(uint32_t * p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add_u64(unsigned long long * p, unsigned long long v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
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

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t * p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t * p) { return *reinterpret_cast<const volatile uint64_t *>(p); }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long * p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ void prefetch_l2(const void * p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ---- mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP / SYNCS) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16 B aligned); completion is signalled on `bar` as
// transaction bytes.  The clouds are read exactly once: evict-first keeps them from displacing the grid's lines in L2.
__device__ __forceinline__ void bulk_load(void * dst_smem, const void * src_gmem, uint32_t bytes, uint64_t * bar, unsigned long long policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// named barrier over a subset of the CTA's warps (id 1..15; `threads` a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

// =====================================================================================
// column tiles, leaf hash, leaf allocator
// =====================================================================================
// find-or-insert the column tile of leaf column (bx,by); tiles are never freed between rebuilds
__device__ inline uint32_t tile_find_or_insert(const GridView & g, int bx, int by)
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

// hand out a leaf slot: the free ring first, then fresh slots from the bump pointer.  No sweep (the only producer of free
// slots) runs concurrently with an allocating kernel, so free_push is constant here and a pop that overshoots it is undone.
__device__ __forceinline__ uint32_t leaf_alloc(const GridView & g)
{
  Counters * c = g.ctr;
  const uint32_t push = ld_volatile_u32(&c->free_push);
  if ((int32_t)(push - ld_volatile_u32(&c->free_pop)) > 0) {
    const uint32_t i = atomicAdd(&c->free_pop, 1u);
    if ((int32_t)(push - i) > 0) return g.free_slots[i & g.free_ring_mask];
    atomicSub(&c->free_pop, 1u);
  }
  const uint32_t s = atomicAdd(&c->bump, 1u);
  return s < g.leaf_capacity ? s : kOverflowSlot;
}

// find-or-insert a leaf; returns its pool slot (or kOverflowSlot).  Probes use ordinary (L1-cacheable) 128-bit loads:
// a stale line can only show EMPTY where a key has since been published (the CAS then returns the truth) or an
// unpublished slot (then the volatile spin below reads L2); keys never change while a Mark kernel runs.
__device__ inline uint32_t leaf_find_or_insert(const GridView & g, uint64_t key, int bx, int by, double tmin_init)
{
  uint32_t h = (uint32_t)mix64(key) & g.hash_mask;
  for (uint32_t probe = 0; probe <= g.hash_mask; ++probe) {
    const uint4 e = *reinterpret_cast<const uint4 *>(&g.hash[h]);
    uint64_t k = (uint64_t)e.x | ((uint64_t)e.y << 32);
    uint32_t v = e.z;
    if (k == kEmptyKey) {
      const uint64_t old = atomicCAS(&g.hash[h].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) {
        const uint32_t slot = leaf_alloc(g);
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

// give an emptied leaf back (sweep kernels only): tombstone in the hash, slot onto the free ring
__device__ __forceinline__ void free_leaf(const GridView & g, uint32_t slot)
{
  const uint32_t h = g.leaf_hpos[slot];
  g.hash[h].key = kTombKey;
  g.hash[h].val = kInvalidSlot;
  g.leaf_key[slot] = kEmptyKey;
  const uint32_t i = atomicAdd(&g.ctr->free_push, 1u);
  g.free_slots[i & g.free_ring_mask] = slot;
  atomicAdd(&g.ctr->n_tombs, 1u);
}

// number of pool slots a scan has to visit
__device__ __forceinline__ uint32_t slots_in_use(const GridView & g)
{
  const uint32_t b = ld_volatile_u32(&g.ctr->bump);
  return b < g.leaf_capacity ? b : g.leaf_capacity;
}

// Opt-in phase tracing for profiling builds only (python -m spatio_temporal_voxel_layer_b200.build --trace builds a
// SEPARATE libstvl_b200_trace.so; the product library compiles none of this).  Thread 0 of every CTA stamps
// %globaltimer at a few phase boundaries; profiles/trace_phases.py turns the stamps into a per-kernel timeline.
#ifdef STVL_TRACE
extern __device__ unsigned long long g_trace[4][1024][8];
__device__ __forceinline__ void trace_stamp(int kernel, int phase)
{
  if (threadIdx.x == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[kernel][blockIdx.x][phase] = t;
    if (phase == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_trace[kernel][blockIdx.x][7] = sm; }
  }
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
// One coordinate of nav2 Costmap2D::worldToMap (external): (unsigned)((w - origin) / resolution), w >= origin.  The division
// is the reference's; a product with the cached 1/resolution decides it wherever that is safe: the product differs from the
// correctly rounded quotient by less than 5e-16 relative, i.e. by less than 1.1e-6 for cell indices below 2^31, so unless it
// lies within 1e-5 of an integer both truncate to the same cell (with the origin a multiple of the resolution the product
// sits half a cell away from the integers for every voxel centre).  Larger quotients and near-integers take the division.
__device__ __forceinline__ unsigned map_cell(double d, double resolution, double inv_resolution)
{
  const double q = __dmul_rn(d, inv_resolution);
  const double r = rint(q);
  if (q < 2147483648.0 && fabs(__dsub_rn(q, r)) > 1e-5) return __double2uint_rz(q);
  return __double2uint_rz(__ddiv_rn(d, resolution));
}
// nav2 Costmap2D::worldToMap (external): wx >= origin, (unsigned)((wx-origin)/resolution) < size
__device__ __forceinline__ bool world_to_map(double wx, double wy, const CostmapDev & p, unsigned & mx, unsigned & my)
{
  if (wx < p.origin_x || wy < p.origin_y) return false;
  mx = map_cell(__dsub_rn(wx, p.origin_x), p.resolution, p.inv_resolution);
  my = map_cell(__dsub_rn(wy, p.origin_y), p.resolution, p.inv_resolution);
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

// The grid loop of UpdateROSCostmap (:708-717) for `kBatch` consecutive column tiles starting at t0, by one warp: the
// count rows (256 B each, two columns per lane) and tile keys of the whole batch are requested together.
// kBits: the output is this rank's packed bitmap of costmap cells (CostmapDev::bit_w_lo / bit_wpr) instead of bytes — the
// form the sharded cycle ships to the root rank.
template <int kBatch, bool kBits>
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
        if constexpr (kBits) {
          const int w = (int)(mx >> 5) - p.bit_w_lo;
          if ((unsigned)w < (unsigned)p.bit_wpr)
            red_or_u32(reinterpret_cast<uint32_t *>(costmap) + (size_t)my * (unsigned)p.bit_wpr + (unsigned)w, 1u << (mx & 31));
        } else {
          costmap[(size_t)my * p.size_x + mx] = p.lethal;   // :715
        }
        ++n; mnx = min(mnx, ix); mny = min(mny, iy); mxx = max(mxx, ix); mxy = max(mxy, iy);   // touch(), :716
      }
    }
  }
}

// reduction of one warp's marked-cell count / bounds into the device accumulators
__device__ __forceinline__ void costmap_warp_accumulate(Counters * c, unsigned n, int mnx, int mny, int mxx, int mxy)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n += __shfl_xor_sync(kFull, n, d);
    mnx = min(mnx, __shfl_xor_sync(kFull, mnx, d)); mny = min(mny, __shfl_xor_sync(kFull, mny, d));
    mxx = max(mxx, __shfl_xor_sync(kFull, mxx, d)); mxy = max(mxy, __shfl_xor_sync(kFull, mxy, d));
  }
  if ((threadIdx.x & 31) == 0 && n) {
    atomicAdd(&c->cm_acc.n_marked, (unsigned long long)n);
    atomicMin(&c->cm_acc.mk_min_x, mnx); atomicMin(&c->cm_acc.mk_min_y, mny);
    atomicMax(&c->cm_acc.mk_max_x, mxx); atomicMax(&c->cm_acc.mk_max_y, mxy);
  }
}

}  // namespace stvl
