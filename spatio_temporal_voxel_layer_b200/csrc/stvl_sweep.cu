// stvl_sweep.cu — ClearFrustums / TemporalClearAndGenerateCostmap (reference spatio_temporal_voxel_grid.cpp:96-224) as ONE
// kernel.
//
// Every CTA walks passes over the leaf slots (dealt to the CTAs in pairs or fours of consecutive slots, so that a run of leaves
// that all need work — a frustum's interior, a part of the map that expires together — is spread over the whole grid); a
// pass has ONE CTA barrier:
//   A  one THREAD per leaf: header (key, tmin, tile, 64 B value mask) in one round trip.  A leaf outside every frustum whose
//      oldest voxel is too young to expire cannot change: its survivors are counted straight from the mask (two adjacent u32
//      column counters per 64-bit reduction) and its 4 KB time block is never touched.  A leaf no frustum can touch but in
//      which something may have expired puts its occupied voxel COLUMNS (8 mark times = one 64 B run) on a column queue;
//      a leaf a frustum may touch goes on the leaf queue.  Both queues live in shared memory; the occupied 64 B time runs
//      are prefetched into L2.
//   B' one THREAD per queued column: 64 B of mark times, the plain decay test of .cpp:203-211 per voxel, the new mask byte,
//      the column's survivor count; the leaf's new lower time bound is gathered in shared memory.
//   B  one WARP per (queued leaf, occupied 32-bit mask word), taken dynamically.  Lane l owns bit l of the word, so its mark
//      times are one coalesced 256 B run and the survivors come back as one ballot — the new mask word.  Per voxel: the
//      reference's exact arithmetic (decay model, first containing frustum in reading order, acceleration, strict <),
//      rewrite of accelerated times, per-column survivor counts, optional lists; the leaf's frustum classification (one lane
//      per frustum) is redone per word, its new lower time bound is gathered in shared memory like the columns'.
//   C  (after the next pass's barrier) the thread that queued a leaf writes its new lower time bound or frees it.
// Three sets of queues rotate, so a warp that runs out of work starts phase A of the next pass while the others finish.
// No global work queue, no second launch, no grid-wide synchronisation: a leaf's whole sweep stays inside one CTA.
#include <algorithm>
#include <cstdlib>
#include <utility>

#include "stvl_common.cuh"

namespace stvl {

constexpr int kSweepThreads = 512;
constexpr int kSweepWarps = kSweepThreads / 32;

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

struct SweepAcc {
  unsigned int swept = 0, surv = 0, clear = 0, rewr = 0, amb = 0;
  int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
};

// can a leaf whose oldest mark time is tmin contain a voxel that expires by decay alone?  The decay duration is
// non-increasing in a voxel's age for the linear model; the exponential / persistent ones never go negative for
// voxel_decay >= 0 (reference .cpp:320-331).  tmin = -inf ("unknown") always answers yes.
__device__ __forceinline__ bool leaf_may_expire(const GridView & g, double now, double tmin)
{
  if (g.decay_model == 0) return __dsub_rn(g.voxel_decay, __dsub_rn(now, tmin)) < 0.;
  return !(g.voxel_decay >= 0.0);
}

// leaf queue entry: pool slot (bits 0-31) | thread that owns the leaf in this pass (32-40) | which of its 16 mask words are
// occupied (41-56) | kTouched when a frustum may contain one of its voxels.  The warps take (leaf, word) tasks from it.
constexpr unsigned long long kTouched = 1ull << 57;
// column queue entry: (thread that owns the leaf in this pass) << 14 | column << 8 | the column's mask byte
constexpr int kColQueue = 4096;

// total order of doubles as unsigned integers (for the shared-memory minimum of a leaf's surviving mark times)
__device__ __forceinline__ unsigned long long ordered_bits(double x)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_value(unsigned long long e)
{
  return __longlong_as_double((long long)((e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e));
}
constexpr unsigned long long kNoSurvivor = ~0ull;

// dynamic shared memory behind the frustum blocks
struct SweepShared {
  uint32_t cq[3][kColQueue];                  // column queues
  unsigned long long lmin[2][kSweepThreads];  // per leaf whose columns are queued (indexed by the owning thread): minimum of the
  unsigned long long lkey[2][kSweepThreads];  //   surviving mark times (ordered_bits), key, slot, tile
  uint32_t lslot[2][kSweepThreads];
  uint32_t ltile[2][kSweepThreads];
  uint32_t lcls[2][kSweepThreads];            // frustum classes of the leaf, two bits per frustum (sets of up to 16 frustums)
};

__global__ void __launch_bounds__(kSweepThreads, 2)
sweep_kernel(const __grid_constant__ GridView g, const __grid_constant__ FrustumParams fparam, const DevFrustum * __restrict__ frustums_gmem,
             const int n_frustums, const double now, const __grid_constant__ SweepOutputs out, uint8_t * __restrict__ cm_reset,
             const unsigned long long cm_bytes, const int cm_value)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevFrustum * fr = reinterpret_cast<DevFrustum *>(smem_raw);   // frustum parameter blocks, staged for the whole kernel
  SweepShared & sh = *reinterpret_cast<SweepShared *>(smem_raw + (((size_t)n_frustums * sizeof(DevFrustum) + 15) & ~(size_t)15));
  // three work queues in rotation: tile pass t fills queue t % 3 in phase A and drains it in phase B; a warp that runs out
  // of phase-B work goes on to phase A of pass t + 1 (queue (t + 1) % 3) without waiting for the others, and queue
  // (t + 2) % 3 — drained in pass t - 1 — is re-armed meanwhile.  ONE CTA barrier per pass.
  __shared__ unsigned long long s_q[3][kSweepThreads];
  __shared__ unsigned int s_qn[3], s_qnext[3], s_cn[3], s_climit[3];
  __shared__ uint32_t s_part[kSweepWarps][kMaxFrustums / 32];
  __shared__ uint32_t s_full[kSweepWarps][kMaxFrustums / 32];
  __shared__ unsigned int s_red[5];
  __shared__ int s_bb[4];
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  TRACE(1, 0);
  // Programmatic dependent launch: the Mark kernel behind this one may become resident now and stream its clouds; it waits
  // for this grid before it touches the store.  The frustum blocks come from the kernel parameters (small sets) or from a
  // device buffer an earlier stream operation filled; neither is written by the predecessor kernel.
  pdl_launch_dependents();
  {
    const int words = n_frustums * (int)(sizeof(DevFrustum) / 8);
    const double * src = reinterpret_cast<const double *>(frustums_gmem ? frustums_gmem : fparam.f);
    double * dst = reinterpret_cast<double *>(fr);
    for (int i = threadIdx.x; i < words; i += kSweepThreads) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    s_qn[0] = s_qn[1] = s_qn[2] = 0; s_qnext[0] = s_qnext[1] = s_qnext[2] = 0; s_cn[0] = s_cn[1] = s_cn[2] = 0; s_climit[0] = s_climit[1] = s_climit[2] = kColQueue;
    s_red[0] = s_red[1] = s_red[2] = s_red[3] = s_red[4] = 0;
    s_bb[0] = s_bb[1] = INT_MAX; s_bb[2] = s_bb[3] = INT_MIN;
  }
  __syncthreads();
  pdl_wait();                // launched programmatically behind the previous cycle's Mark kernel: the grid is ours from here
  const uint32_t n_slots = slots_in_use(g);
  TRACE(1, 1);
  // re-arm the state of the NEXT sweep while this one runs: zero the other tile buffer and the other counter set
  {
    const uint32_t n_tiles = min(ld_volatile_u32(&g.ctr->tile_n), g.tile_capacity);
    const uint32_t n_words = n_tiles * 64u;
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
      if (g.live) {   // mirror of the allocator counters for the host's grid sizing (posted stores only)
        volatile LiveCounters * l = g.live;
        l->high_water = n_slots; l->free_n = ld_volatile_u32(&g.ctr->free_push) - ld_volatile_u32(&g.ctr->free_pop); l->tile_n = n_tiles;
      }
    }
  }
  const bool force_full = out.points != nullptr;   // pub_voxels needs every survivor's coordinates
  const int n_chunks = (n_frustums + 31) >> 5;
  const bool cache_classes = n_frustums <= 16;   // phase A classifies a leaf against every frustum once, the warp pass reuses it
  SweepAcc acc;
  TRACE(1, 2);

  // Passes: thread t of CTA b takes, in pass p, the leaf (p * 512 + t) of the CTA's share.  Slots are dealt to the CTAs in
  // pairs of consecutive ones (fine enough that the few hundred leaves a frustum touches — consecutive slots — end up in as
  // many CTAs; the 8-byte header fields still fill half a 32 B sector), large stores in fours (full sectors: there the
  // header traffic matters and every CTA has many leaves of every kind anyway).
  const uint32_t kDeal = n_slots > 262144u ? 4u : 2u;
  const uint32_t n_deals = (n_slots + kDeal - 1) / kDeal;
  const uint32_t n_passes = (n_deals + gridDim.x * (kSweepThreads / kDeal) - 1) / (gridDim.x * (kSweepThreads / kDeal));
  auto slot_of = [&](uint32_t pass) {
    const uint32_t q = pass * kSweepThreads + threadIdx.x;
    return ((q / kDeal) * gridDim.x + blockIdx.x) * kDeal + (q % kDeal);
  };
  uint32_t qi = 0;   // queues of this pass (pass number modulo 3)
  // bit p & 1: this thread queued the columns of a leaf in pass p and has not yet written the leaf's new bound (that happens
  // behind the NEXT pass's barrier, i.e. after this thread's phase A of that pass — hence two bits, and the leaf's slot is
  // kept in the parity-indexed shared arrays)
  uint32_t leaf_pending = 0;
  auto finish_leaf = [&](uint32_t pb) {   // phase C
    if ((leaf_pending >> pb) & 1u) {
      const uint32_t cslot = sh.lslot[pb][threadIdx.x];
      const unsigned long long m = sh.lmin[pb][threadIdx.x];
      if (m == kNoSurvivor) free_leaf(g, cslot);            // every voxel of the leaf was cleared: pruneGrid, .cpp:223
      else g.leaf_tmin[cslot] = ordered_value(m);           // exact lower bound again
      leaf_pending &= ~(1u << pb);
    }
  };
  for (uint32_t pass = 0; pass < n_passes; ++pass, qi = qi == 2 ? 0 : qi + 1) {
    const uint32_t pb = pass & 1u;
    // ---- phase A: one thread per leaf -------------------------------------------------------------------------------------
    {
      const uint32_t slot = slot_of(pass);
      unsigned long long entry = 0;
      bool queue_it = false;
      if (slot < n_slots) {
        const uint64_t key = g.leaf_key[slot];
        const double tmin = g.leaf_tmin[slot];
        const uint32_t tile = g.leaf_tile[slot];
        const uint4 * mp = reinterpret_cast<const uint4 *>(g.leaf_mask + (size_t)slot * 16);
        uint4 mw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) mw[q] = mp[q];
        if (pass == 0) TRACE(2, 0);
        if (key != kEmptyKey) {
          uint32_t any = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) any |= mw[q].x | mw[q].y | mw[q].z | mw[q].w;
          if (any == 0) {
            free_leaf(g, slot);   // nothing left in this leaf: pruneGrid, .cpp:223
          } else {
            // can a frustum contain a voxel centre of this leaf?  (bounding box, then the leaf's bounding sphere against the
            // plane set / hourglass)
            bool touches = force_full;
            uint32_t classes = 0;   // (small frustum sets: every frustum's class, kept for the warp pass)
            {
              int bx, by, bz;
              unpack_leaf_key(key, bx, by, bz);
              const int x0 = bx * 8, y0 = by * 8, z0 = bz * 8;
              for (int f = 0; f < n_frustums && (cache_classes || !touches); ++f) {
                const DevFrustum & F = fr[f];
                if (F.type < 0) continue;
                if (F.cull && (x0 + 7 < F.lo[0] || x0 > F.hi[0] || y0 + 7 < F.lo[1] || y0 > F.hi[1] || z0 + 7 < F.lo[2] || z0 > F.hi[2])) continue;
                const int cls = classify_leaf(F, bx, by, bz, g.vs, g.half_vs);
                touches = touches || cls != 0;
                if (cache_classes) classes |= (uint32_t)cls << (2 * f);
              }
            }
            if (touches || leaf_may_expire(g, now, tmin)) {
              const uint32_t wv[16] = {mw[0].x, mw[0].y, mw[0].z, mw[0].w, mw[1].x, mw[1].y, mw[1].z, mw[1].w,
                                       mw[2].x, mw[2].y, mw[2].z, mw[2].w, mw[3].x, mw[3].y, mw[3].z, mw[3].w};
              bool as_columns = false;
              if (!touches) {
                // decay test only: one queue entry per occupied voxel column, if the queue has room for all of them
                unsigned n_cols = 0;
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                  uint32_t t = wv[w] | (wv[w] >> 4);
                  t |= t >> 2; t |= t >> 1;
                  n_cols += __popc(t & 0x01010101u);
                }
                unsigned c0 = atomicAdd(&s_cn[qi], n_cols);
                if (c0 + n_cols <= (unsigned)kColQueue) {
                  as_columns = true;
                  sh.lmin[pb][threadIdx.x] = kNoSurvivor;
                  sh.lkey[pb][threadIdx.x] = key;
                  sh.lslot[pb][threadIdx.x] = slot;
                  sh.ltile[pb][threadIdx.x] = tile;
                  leaf_pending |= 1u << pb;
#pragma unroll
                  for (int w = 0; w < 16; ++w) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                      const uint32_t byte = (wv[w] >> (8 * c)) & 0xFFu;
                      if (byte) sh.cq[qi][c0++] = (threadIdx.x << 14) | ((uint32_t)(w * 4 + c) << 8) | byte;
                    }
                  }
                } else {
                  atomicMin(&s_climit[qi], c0);   // no room (nor for anybody after this reservation): the leaf takes the warp-per-leaf pass
                }
              }
              if (!as_columns) {
                unsigned nz = 0;
#pragma unroll
                for (int w = 0; w < 16; ++w) nz |= wv[w] ? (1u << w) : 0u;
                queue_it = true;
                entry = (unsigned long long)slot | ((unsigned long long)threadIdx.x << 32) | ((unsigned long long)nz << 41) | (touches ? kTouched : 0ull);
                sh.lmin[pb][threadIdx.x] = kNoSurvivor;
                sh.lkey[pb][threadIdx.x] = key;
                sh.lslot[pb][threadIdx.x] = slot;
                sh.ltile[pb][threadIdx.x] = tile;
                sh.lcls[pb][threadIdx.x] = classes;
                leaf_pending |= 1u << pb;
              }
              // the leaf's mark times will be needed in phase B: pull the occupied 64 B column runs into L2 now
              const double * tb = g.leaf_time + (size_t)slot * kLeafVoxels;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t wv[4] = {mw[q].x, mw[q].y, mw[q].z, mw[q].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t x = wv[j];
                  if (!x) continue;
                  const double * wb = tb + (q * 4 + j) * 32;
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    if ((x >> (8 * c)) & 0xFFu) prefetch_l2(wb + 8 * c);
                }
              }
            } else {
              // nothing in this leaf can change: PopulateCostmapAndPointcloud .cpp:242-250 for every voxel straight from the
              // mask — per-column popcounts, two adjacent u32 column counters per 64-bit reduction (no carry: counts < 2^32)
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
                    if (lo) red_add_u64(tc + (q * 4 + j) * 2, lo);
                    if (hi) red_add_u64(tc + (q * 4 + j) * 2 + 1, hi);
                  }
                }
              }
              acc.swept += cnt; acc.surv += cnt;
            }
          }
        }
      }
      // warp-aggregated push
      const unsigned qm = __ballot_sync(kFull, queue_it);
      if (qm) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&s_qn[qi], (unsigned)__popc(qm));
        base = __shfl_sync(kFull, base, 0);
        if (queue_it) s_q[qi][base + __popc(qm & ((1u << lane) - 1u))] = entry;
      }
    }
    if (pass == 0) TRACE(2, 1);
    __syncthreads();
    if (pass == 0) TRACE(2, 2);
    if (threadIdx.x == 0) {   // re-arm the queues of pass t + 2 (those of pass t - 1: everybody is done with them)
      const uint32_t qz = qi == 0 ? 2 : qi - 1;
      s_qn[qz] = 0; s_qnext[qz] = 0; s_cn[qz] = 0; s_climit[qz] = kColQueue;
    }
    const uint32_t n_items = s_qn[qi];
    const uint32_t n_cols_queued = min(s_cn[qi], s_climit[qi]);
    // ---- phase C of the previous pass: its columns have all been worked on ---------------------------------------------------
    if (pass > 0) finish_leaf(pb ^ 1u);
    // ---- phase B': one thread per queued voxel column (leaves no frustum can touch): the decay test of .cpp:203-211 ----------
    for (uint32_t i = threadIdx.x; i < n_cols_queued; i += kSweepThreads) {
      const uint32_t it = sh.cq[qi][i];
      const uint32_t owner = it >> 14, col = (it >> 8) & 63u, byte = it & 0xFFu;
      const uint32_t cslot = sh.lslot[pb][owner];
      const double2 * tp = reinterpret_cast<const double2 *>(g.leaf_time + (size_t)cslot * kLeafVoxels + col * 8);
      const double2 t01 = __ldcs(tp), t23 = __ldcs(tp + 1), t45 = __ldcs(tp + 2), t67 = __ldcs(tp + 3);
      const double tv[8] = {t01.x, t01.y, t23.x, t23.y, t45.x, t45.y, t67.x, t67.y};
      uint32_t nb = byte;
      double vmin = CUDART_INF;
      if (g.decay_model == 0) {   // linear: duration = decay - (now - v), .cpp:167-169 / :320-324
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const bool expired = __dsub_rn(g.voxel_decay, __dsub_rn(now, tv[k])) < 0.;
          if ((byte >> k) & 1u) {
            if (expired) nb &= ~(1u << k);
            else vmin = fmin(vmin, tv[k]);
          }
        }
      } else {
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
          if (!((byte >> k) & 1u)) continue;
          const double v = tv[k];
          const double t = __dsub_rn(now, v);
          const double base_dur = g.decay_model == 1 ? __dmul_rn(g.voxel_decay, exp(-t)) : g.voxel_decay;
          if (base_dur < 0.) nb &= ~(1u << k);
          else vmin = fmin(vmin, v);
        }
      }
      const unsigned n_in = __popc(byte), n_left = __popc(nb);
      acc.swept += n_in; acc.surv += n_left; acc.clear += n_in - n_left;
      if (nb != byte) {
        reinterpret_cast<uint8_t *>(g.leaf_mask)[(size_t)cslot * 64 + col] = (uint8_t)nb;   // this thread alone owns the column's mask byte
        int bx, by, bz;
        unpack_leaf_key(sh.lkey[pb][owner], bx, by, bz);
        const int ix = bx * 8 + (int)(col >> 3), iy = by * 8 + (int)(col & 7u);
        acc.min_x = min(acc.min_x, ix); acc.max_x = max(acc.max_x, ix);
        acc.min_y = min(acc.min_y, iy); acc.max_y = max(acc.max_y, iy);
        if (out.cleared) {   // cleared_cells.insert, .cpp:213-215: one entry per cleared voxel
          const unsigned n_cl = n_in - n_left;
          const unsigned long long o = atomicAdd(&g.ctr->sw[g.sw_cur].n_cleared_out, (unsigned long long)n_cl);
          for (unsigned c = 0; c < n_cl; ++c)
            if (o + c < out.cleared_cap) { out.cleared[2 * (o + c)] = ix; out.cleared[2 * (o + c) + 1] = iy; }
        }
      }
      if (nb) {
        // PopulateCostmapAndPointcloud .cpp:242-250: _cost_map[(x,y)] += 1 per surviving voxel of the column
        const uint32_t ctile = sh.ltile[pb][owner];
        if (ctile != kOverflowSlot) red_add_u32(g.tile_count + (size_t)ctile * 64 + col, n_left);
        atomicMin(&sh.lmin[pb][owner], ordered_bits(vmin));
      }
    }
    if (pass == 0) { TRACE(1, 6); TRACE_VAL(1, 5, (unsigned long long)n_items | ((unsigned long long)n_cols_queued << 32)); }
    // ---- phase B: one warp per (queued leaf, mask word), taken dynamically.  Lane l owns bit l of the word: its 32 mark times
    // are one coalesced 256 B run, the survivors come back as one ballot = the new mask word -----------------------------------
    for (;;) {
      uint32_t task = 0;
      if (lane == 0) task = atomicAdd(&s_qnext[qi], 1u);
      task = __shfl_sync(kFull, task, 0);
      if (task >= n_items * 16u) break;
      const unsigned long long ent = s_q[qi][task >> 4];
      const int w = (int)(task & 15u);
      if (!((ent >> (41 + w)) & 1ull)) continue;   // an empty word of the leaf
      const uint32_t bslot = (uint32_t)ent;
      const uint32_t owner = (uint32_t)(ent >> 32) & 0x1FFu;
      const bool touched = (ent & kTouched) != 0;
      const uint32_t m = g.leaf_mask[(size_t)bslot * 16 + w];
      const bool have = ((m >> lane) & 1u) != 0;
      const double v = have ? __ldcs(g.leaf_time + (size_t)bslot * kLeafVoxels + w * 32 + (int)lane) : 0.0;
      const uint64_t key = sh.lkey[pb][owner];
      const uint32_t tile = sh.ltile[pb][owner];
      int bx, by, bz;
      unpack_leaf_key(key, bx, by, bz);
      // leaf-level culling: one lane per frustum
      unsigned any_cand = 0;
      if (touched) {
        const uint32_t classes = sh.lcls[pb][owner];
        for (int c = 0; c < n_chunks; ++c) {
          const int f = c * 32 + (int)lane;
          const int cls = f >= n_frustums ? 0 : cache_classes ? (int)((classes >> (2 * (f & 15))) & 3u) : classify_leaf(fr[f], bx, by, bz, g.vs, g.half_vs);
          const unsigned pm = __ballot_sync(kFull, cls == 1), fm = __ballot_sync(kFull, cls == 2);
          if (lane == 0) { s_part[warp][c] = pm; s_full[warp][c] = fm; }
          any_cand |= pm | fm;
        }
        __syncwarp();
      }
      double vmin = CUDART_INF;   // smallest mark time among the word's survivors
      const unsigned li = (unsigned)w * 32u + lane;
      const int ix = bx * 8 + (int)(li >> 6), iy = by * 8 + (int)((li >> 3) & 7), iz = bz * 8 + (int)(li & 7);
      bool cleared = false;
      if (have) {
        ++acc.swept;
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
            unsigned cm = s_part[warp][c] | fm;
            while (cm) {
              const int b = __ffs(cm) - 1; cm &= cm - 1;
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
            if (__double_as_longlong(nv) != __double_as_longlong(v)) { g.leaf_time[(size_t)bslot * kLeafVoxels + li] = nv; ++acc.rewr; }
            vmin = nv;
          }
        } else if (base_dur < 0.) {
          cleared = true;   // .cpp:203-211
        } else {
          vmin = v;
        }
        if (cleared) {
          ++acc.clear;
          acc.min_x = min(acc.min_x, ix); acc.max_x = max(acc.max_x, ix);
          acc.min_y = min(acc.min_y, iy); acc.max_y = max(acc.max_y, iy);
        } else {
          ++acc.surv;
        }
        if (amb) {
          ++acc.amb;
          if (out.ambiguous) {
            const unsigned long long a = atomicAdd(&g.ctr->sw[g.sw_cur].n_ambiguous_out, 1ull);
            if (a < out.ambiguous_cap) { out.ambiguous[3 * a] = ix; out.ambiguous[3 * a + 1] = iy; out.ambiguous[3 * a + 2] = iz; }
          }
        }
      }
      const unsigned cmk = __ballot_sync(kFull, cleared);
      const uint32_t sm = m & ~cmk;                                    // the word's survivors = its new mask
      if (cmk && lane == 0) g.leaf_mask[(size_t)bslot * 16 + w] = sm;   // this warp is the only writer of the word
      // PopulateCostmapAndPointcloud .cpp:242-250: _cost_map[(x,y)] += 1 per surviving voxel — the word holds four voxel
      // columns, one per byte; lanes 0 and 1 each push a pair of adjacent u32 column counters as one 64-bit reduction
      if (sm && tile != kOverflowSlot && lane < 2) {
        unsigned long long * tc = reinterpret_cast<unsigned long long *>(g.tile_count + (size_t)tile * 64);
        uint32_t y = sm - ((sm >> 1) & 0x55555555u);
        y = (y & 0x33333333u) + ((y >> 2) & 0x33333333u);
        y = (y + (y >> 4)) & 0x0F0F0F0Fu;
        y >>= 16 * lane;
        const unsigned long long pair = (unsigned long long)(y & 0xFFu) | ((unsigned long long)((y >> 8) & 0xFFu) << 32);
        if (pair) red_add_u64(tc + w * 2 + (int)lane, pair);
      }
      // optional lists: warp-aggregated appends
      if (out.points && sm) {   // PopulateCostmapAndPointcloud .cpp:234-240: Point32 = (float) voxel centre
        unsigned long long b0 = 0;
        if (lane == 0) b0 = atomicAdd(&g.ctr->sw[g.sw_cur].n_points_out, (unsigned long long)__popc(sm));
        b0 = __shfl_sync(kFull, b0, 0);
        if ((sm >> lane) & 1u) {
          const unsigned long long o = b0 + __popc(sm & ((1u << lane) - 1u));
          if (o < out.points_cap) {
            out.points[3 * o] = (float)index_to_world(ix, g.vs, g.half_vs);
            out.points[3 * o + 1] = (float)index_to_world(iy, g.vs, g.half_vs);
            out.points[3 * o + 2] = (float)index_to_world(iz, g.vs, g.half_vs);
          }
        }
      }
      if (out.cleared && cmk) {   // cleared_cells.insert, .cpp:213-215
        unsigned long long b0 = 0;
        if (lane == 0) b0 = atomicAdd(&g.ctr->sw[g.sw_cur].n_cleared_out, (unsigned long long)__popc(cmk));
        b0 = __shfl_sync(kFull, b0, 0);
        if (cleared) {
          const unsigned long long o = b0 + __popc(cmk & ((1u << lane) - 1u));
          if (o < out.cleared_cap) { out.cleared[2 * o] = ix; out.cleared[2 * o + 1] = iy; }
        }
      }
      // the word's share of the leaf's new lower time bound (phase C turns the minimum into the bound, or frees the leaf)
      if (sm) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vmin = fmin(vmin, __shfl_xor_sync(kFull, vmin, d));
        if (lane == 0) atomicMin(&sh.lmin[pb][owner], ordered_bits(vmin));
      }
    }
  }
  if (n_passes > 0) {   // phase C of the last pass
    __syncthreads();
    finish_leaf((n_passes - 1) & 1u);
  }
  TRACE(1, 3);

  // ---- statistics: warp reduce, shared-memory add, a handful of global reductions per CTA -------------------------------------
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    acc.swept += __shfl_xor_sync(kFull, acc.swept, d); acc.surv += __shfl_xor_sync(kFull, acc.surv, d);
    acc.clear += __shfl_xor_sync(kFull, acc.clear, d); acc.rewr += __shfl_xor_sync(kFull, acc.rewr, d);
    acc.amb += __shfl_xor_sync(kFull, acc.amb, d);
    acc.min_x = min(acc.min_x, __shfl_xor_sync(kFull, acc.min_x, d)); acc.min_y = min(acc.min_y, __shfl_xor_sync(kFull, acc.min_y, d));
    acc.max_x = max(acc.max_x, __shfl_xor_sync(kFull, acc.max_x, d)); acc.max_y = max(acc.max_y, __shfl_xor_sync(kFull, acc.max_y, d));
  }
  if (lane == 0 && (acc.swept | acc.surv | acc.clear | acc.rewr | acc.amb)) {
    atomicAdd(&s_red[0], acc.swept); atomicAdd(&s_red[1], acc.surv); atomicAdd(&s_red[2], acc.clear);
    atomicAdd(&s_red[3], acc.rewr); atomicAdd(&s_red[4], acc.amb);
    atomicMin(&s_bb[0], acc.min_x); atomicMin(&s_bb[1], acc.min_y); atomicMax(&s_bb[2], acc.max_x); atomicMax(&s_bb[3], acc.max_y);
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
  TRACE(1, 4);
}

// =====================================================================================
// host-callable launcher
// =====================================================================================
namespace {
int env_int(const char * name, int dflt, int lo, int hi)
{
  const char * v = std::getenv(name);
  if (!v || !*v) return dflt;
  const int x = std::atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
}  // namespace

// frustums_dev == NULL: the (<= kParamFrustums) blocks travel in *fparam
cudaError_t launch_sweep(const GridView & g, const DevFrustum * frustums_dev, const FrustumParams * fparam, int n_frustums, double now,
                         const SweepOutputs & out, uint32_t leaf_slots_upper_bound, int sm_count, cudaStream_t s, bool pdl, bool share_sms,
                         uint8_t * costmap_reset, size_t costmap_bytes, int costmap_value)
{
  static const FrustumParams no_params{};
  const FrustumParams & fp = (frustums_dev == nullptr && fparam) ? *fparam : no_params;
  static bool carveout_set = false;
  if (!carveout_set) {   // see launch_mark_variant
    cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  const size_t dyn = (((size_t)n_frustums * sizeof(DevFrustum) + 15) & ~(size_t)15) + sizeof(SweepShared);
  static size_t dyn_max = 0;
  if (dyn > dyn_max) {
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    dyn_max = dyn;
  }
  // Persistent grid, two CTAs per SM (one while a Mark kernel is to share the SMs, see STVL_SWEEP_CTAS_PER_SM_FUSED); small
  // stores get one CTA per 128 leaves, so that a CTA's warps have leaves to share.
  static const int ctas_alone = env_int("STVL_SWEEP_CTAS_PER_SM", 2, 1, 2), ctas_shared = env_int("STVL_SWEEP_CTAS_PER_SM_FUSED", 2, 1, 2);
  const unsigned max_ctas = (unsigned)(sm_count * (share_sms ? ctas_shared : ctas_alone));
  const unsigned long long want = ((unsigned long long)leaf_slots_upper_bound + 127) / 128;
  const unsigned grid = (unsigned)std::min<unsigned long long>(max_ctas, std::max<unsigned long long>(want, 1));
  return launch_ex(sweep_kernel, grid, (unsigned)kSweepThreads, dyn, s, pdl, g, fp, frustums_dev, n_frustums, now, out, costmap_reset, (unsigned long long)costmap_bytes,
                         costmap_value);
}

}  // namespace stvl
