// stvl_mark.cu — Mark / operator() (reference spatio_temporal_voxel_grid.cpp:254-309) as a warp-specialised sm_100a kernel.
//
// One persistent CTA of 32 warps per SM.  The batch's points (all readings laid end to end) are cut into equal slices, four
// per CTA, dealt round-robin: every CTA gets the same number of points from four distant places of the batch (floor rows
// that the range screen rejects are cheap, wall rows are not), with no work cursor and no atomics on the hand-out.  A slice
// is streamed as units of up to 1984 points (31 KB, never across two readings).  Inside a CTA:
//   producer warp   one elected lane streams the units global -> shared memory with the bulk-copy engine (cp.async.bulk,
//                   completion on an mbarrier as transaction bytes; SASS UBLKCP / SYNCS), six units in flight per SM; the
//                   first copies are issued before the CTA has finished setting up its tables
//   31 consumer     wait on the stage's "full" mbarrier, pull two points each into registers, release the stage ("empty"
//   warps           mbarrier, one arrival per warp) and run the reference's per-point arithmetic: range / z-window tests and
//                   voxel quantisation on a single-precision fast path that is PROVEN to agree with the reference's double
//                   arithmetic (see quantise_axis_fast), the exact double sequence for the few points it cannot decide;
//                   survivors are ORed into a CTA-local table of {(reading, leaf) -> 512-bit mask} in shared memory
// Nothing before the flush touches the grid, so all of the streaming may overlap the sweep that precedes Mark on the stream
// (programmatic dependent launch).  The flush resolves every table entry in the global leaf hash once and pushes one
// 32-bit OR per touched mask word and one 64-bit max (or store) per distinct voxel.  In the fused cycle all 32 warps then
// run the costmap pass over the sweep's column counts.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>

#include "stvl_common.cuh"

namespace stvl {

constexpr int kMarkConsumerWarps = 31;
constexpr int kMarkConsumers = kMarkConsumerWarps * 32;
constexpr int kMarkThreads = kMarkConsumers + 32;
constexpr int kMarkPPT = 2;                      // points per consumer thread per unit
constexpr int kMarkChunk = kMarkConsumers * kMarkPPT;   // 1984 points = 31 KB of packed float4 points: one stage of the ring
constexpr int kMarkSlices = 4;                   // slices per CTA
constexpr int kMarkFilter = 8192;                // direct-mapped filter of voxel codes, entries (a power of two)
constexpr int kMarkFilterShift = 19;             // 32 - log2(kMarkFilter)
constexpr int kMarkQueue = 8192;                 // voxels to write, entries
constexpr int kMarkQueueFlush = 2048;            // the producer asks for an early flush above this fill; at most kMarkStagesMax
                                                 // units (5952 points) are still consumed before every warp has seen the request
constexpr int kMarkStagesMax = 3;
// voxel code: reading (4 bits) | x (10) | y (10) | z (8), relative to the voxel of the reading's origin
constexpr int kCodeBiasXY = 512, kCodeBiasZ = 128;
static_assert(kMarkQueueFlush + kMarkStagesMax * kMarkChunk <= kMarkQueue, "queue sized for the consumers' lag behind a flush request");
static_assert((1 << (32 - kMarkFilterShift)) == kMarkFilter, "filter index = upper bits of the multiplicative hash");
// per-stage word the producer posts: what the stage holds | sensor-frame cloud | reading | threads with two points | points.
// Thread t < half holds points t and half + t of the unit (half = a multiple of 32: warps beyond it sit the unit out).
constexpr uint32_t kMetaEnd = 1u << 31, kMetaFlush = 1u << 30, kMetaTf = 1u << 29;
static_assert(kMarkThreads == 1024, "one CTA of 32 warps");
static_assert(kMarkPPT == 2, "the consumer loop is written for one pair of points per thread");

struct PointXYZ { float x, y, z; bool ok; };

// generic PointCloud2 layouts (anything that is not packed x,y,z,pad float4): straight from global memory
__device__ __forceinline__ PointXYZ load_point_generic(const DevReading & r, unsigned long long i, bool exists)
{
  PointXYZ p;
  if (!exists) { p.x = p.y = p.z = 0.f; p.ok = false; return p; }
  if (r.vec4) {
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

// write one voxel straight to the global store (a voxel beyond the reach of the 32-bit code, or a full queue)
__device__ __noinline__ void mark_voxel_direct(const GridView & g, int bx, int by, int bz, unsigned l, double now, bool use_max, double tmin_init)
{
  const uint32_t slot = leaf_find_or_insert(g, pack_leaf_key(bx, by, bz), bx, by, tmin_init);
  if (slot == kOverflowSlot) return;
  atomicOr(&g.leaf_mask[(size_t)slot * 16 + (l >> 5)], 1u << (l & 31));
  double * tp = &g.leaf_time[(size_t)slot * kLeafVoxels + l];
  if (use_max) red_max_s64(reinterpret_cast<unsigned long long *>(tp), (unsigned long long)__double_as_longlong(now));
  else {
    *tp = now;   // MarkGridPoint: setValueOn overwrite, .cpp:421-430
    if (g.leaf_tmin[slot] > now) g.leaf_tmin[slot] = -CUDART_INF;   // clocks are not monotone here
  }
}

// The reference's per-point arithmetic in double (.cpp:285-303): range test on the float-narrowed squared distance, the
// -vs shift of negative coordinates (z keyed on y's sign — sic), multiplication by the cached 1/vs, truncation toward zero.
// Returns (ix, iy, iz, status) by value (registers): status 1 = marked, 0 = skipped, 2 = passed the range test but lies
// beyond +-2^23 voxels, 3 = a coordinate is not finite (UB upstream: dropped and counted here), 4 = outside the z window.
__device__ __noinline__ int4 quantise_exact(float fx, float fy, float fz, double ox, double oy, double oz, float range2, double vs, double inv_vs,
                                            float zlo, float zhi)
{
  if (!(fabsf(fx) <= 3.402823466e38f && fabsf(fy) <= 3.402823466e38f && fabsf(fz) <= 3.402823466e38f)) return make_int4(0, 0, 0, 3);
  if (fz < zlo || fz > zhi) return make_int4(0, 0, 0, 4);   // PassThrough z window, measurement_buffer.cpp:165-175
  const double xd = (double)fx, yd = (double)fy, zd = (double)fz;
  const double dx = __dsub_rn(xd, ox), dy = __dsub_rn(yd, oy), dz = __dsub_rn(zd, oz);
  const float distance_2 = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  if (distance_2 > range2 || (double)distance_2 < 0.0001) return make_int4(0, 0, 0, 0);   // .cpp:289
  const double X = fx < 0 ? __dsub_rn(xd, vs) : xd;
  const double Y = fy < 0 ? __dsub_rn(yd, vs) : yd;
  const double Z = fy < 0 ? __dsub_rn(zd, vs) : zd;
  const double gx = __dmul_rn(X, inv_vs), gy = __dmul_rn(Y, inv_vs), gz = __dmul_rn(Z, inv_vs);
  const double lim = 8388608.0;   // 2^23 voxels = 2^20 leaves
  if (!(fabs(gx) < lim && fabs(gy) < lim && fabs(gz) < lim)) return make_int4(0, 0, 0, 2);
  return make_int4(__double2int_rz(gx), __double2int_rz(gy), __double2int_rz(gz), 1);
}

// One axis of the single-precision fast path.  S = shifted coordinate in float (x, or x - vs rounded to float; vs is a float
// widened to double, so the reference's S differs from ours by at most one float rounding); a = |S|.  The reference's
// |S * (1/vs)| in double differs from a * (float)(1/vs) by a relative error below 3 * 2^-24 + 2^-52 (shift rounding, rounding
// of 1/vs to float, and the float product we do NOT form), so it lies in the real interval [a * c_lo, a * c_hi] with
//   c_lo = fl((float)(1/vs) * (1 - 5 * 2^-24)),   c_hi = fl((float)(1/vs) * (1 + 3 * 2^-23))
// (one more rounding each: 1 - 5 * 2^-24 leaves two ulps of slack below 1 - 4 * 2^-24, likewise above).  A fused
// multiply-add rounded toward zero gives floor(a * c) EXACTLY: a * c + 2^23 is formed without rounding and truncated to a
// float in [2^23, 2^24), whose ulp is 1.  If both ends of the interval have the same floor, no integer lies in it or on its
// ends, so the reference truncates to that same integer; otherwise `diff` gets a non-zero bit and the point takes the
// exact path.  Sums of 2^24 and more (indices beyond 2^23) always differ: the ends are more than two ulps apart there.
// (Infinite products compare equal — such points never get here: their squared distance is not finite either.)
__device__ __forceinline__ int quantise_axis_fast(float S, float c_lo, float c_hi, uint32_t & diff)
{
  const float a = fabsf(S);
  const uint32_t t_lo = __float_as_uint(__fmaf_rz(a, c_lo, 8388608.0f));
  const uint32_t t_hi = __float_as_uint(__fmaf_rz(a, c_hi, 8388608.0f));
  diff |= t_lo ^ t_hi;
  const int fl = (int)(t_lo & 0x7FFFFFu);   // floor(|S| / vs)
  return S < 0.0f ? -fl : fl;              // the reference truncates toward zero
}
// x and y, whose shift hangs on their own sign (.cpp:293-298): for f < 0 the reference computes trunc((f - vs) * (1/vs)) =
// -floor((|f| + vs) / vs) = -(floor(|f| / vs) + 1) = ~floor(|f| / vs), so the shift never has to be formed: floor(|f| / vs)
// from the interval test above (a = |f| exactly, one rounding fewer), complemented for negative f.  What the reference adds
// to u = |f| / vs on the way — rounding (f - vs), 1/vs and the product in double — is below (u + 1) * 4e-16, far inside the
// interval's slack of 1.8e-7 * u around every integer >= 1; around 0 the slack is u itself, so negatives with |f| < 1e-12
// (u < 4e-16 needs vs > 2500 m) and -0.0 (whose sign bit is set although the reference does not shift it) are left to the
// exact path: as signed integers exactly those bit patterns are below that of -1e-12f.
__device__ __forceinline__ int quantise_xy_fast(float f, float c_lo, float c_hi, uint32_t & diff, bool & tiny)
{
  const float a = fabsf(f);
  const uint32_t t_lo = __float_as_uint(__fmaf_rz(a, c_lo, 8388608.0f));
  const uint32_t t_hi = __float_as_uint(__fmaf_rz(a, c_hi, 8388608.0f));
  diff |= t_lo ^ t_hi;
  tiny = tiny | (__float_as_int(f) < (int)0xAB8CBCCC);      // -1e-12f
  return (int)(t_lo & 0x7FFFFFu) ^ (__float_as_int(f) >> 31);
}

// kVec4: every reading of the batch is a packed x,y,z,pad float4 cloud (the RGBD / depth-image layout), streamed through
// the bulk-copy pipeline.  Other layouts are read from global memory by the consumer threads themselves (same chunk
// hand-out by the producer warp, no staging).
// kCostmap: the fused update cycle — this launch also runs the costmap pass over the column counts of the sweep that
// precedes it (they are independent of Mark: the costmap reflects the pre-Mark survivors, reference .cpp:149-224 then
// spatio_temporal_voxel_layer.cpp:699-718); the costmap bytes were reset by the sweep kernel.  kCmBits: the pass writes a
// cell bitmap instead of bytes (sharded cycle).
template <bool kVec4, bool kCostmap, bool kCmBits>
__global__ void __launch_bounds__(kMarkThreads, 1)
mark_kernel(const __grid_constant__ GridView g, const __grid_constant__ MarkBatch batch, const __grid_constant__ CostmapDev cm,
            uint8_t * __restrict__ costmap, const int n_stages)
{
  extern __shared__ __align__(128) unsigned char s_dyn[];
  // dynamic shared memory: [stages (kVec4 only)] [filter] [queue]
  const size_t stage_bytes = kVec4 ? (size_t)n_stages * kMarkChunk * 16 : 0;
  uint32_t * const s_filter = reinterpret_cast<uint32_t *>(s_dyn + stage_bytes);
  uint32_t * const s_queue = s_filter + kMarkFilter;
  __shared__ double s_rnow[kMaxBatchReadings];
  // per reading: the single-precision screen of the range / z tests {range2_reject, range2_accept, near_accept, ox}
  // {oy, oz, zlo, zhi} and the voxel code's frame {ref_ix - bias, ref_iy - bias, ref_iz - bias, reading << 28}
  __shared__ __align__(16) float s_par[kMaxBatchReadings][12];
  __shared__ __align__(8) uint64_t s_full[kMarkStagesMax], s_empty[kMarkStagesMax];
  __shared__ unsigned long long s_base[kMarkStagesMax];   // per stage: first point of the unit in its reading (generic layouts)
  __shared__ uint32_t s_meta[kMarkStagesMax];
  __shared__ unsigned int s_qn;            // queued voxels (may run past kMarkQueue: the excess went straight to the store)
  __shared__ unsigned int s_flushes;       // early flushes done (the producer requests at most one at a time)
  __shared__ unsigned int s_drop[2];
  __shared__ int s_cm[5];                  // fused costmap pass: marked cells and their bounds, gathered per CTA

  TRACE(0, 0);
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const bool use_max = batch.use_atomic_max != 0;

  // =====================================================================================================================
  // producer warp
  // =====================================================================================================================
  if (warp == kMarkConsumerWarps) {
    // slices of `slice_len` points of the batch's point index space; this CTA takes slices blockIdx.x, blockIdx.x + gridDim.x, ...
    const unsigned long long total = batch.first_point[batch.n_readings];
    const unsigned long long slice_len = batch.slice_len;
    unsigned long long slice0 = (unsigned long long)blockIdx.x * slice_len;   // first point of the current slice
    unsigned long long pos = slice0;                                          // next point to hand out
    int stage = 0;            // ring position of the next hand-out ...
    uint32_t phase = 0;       // ... and how often the ring has wrapped (parity)
    unsigned flush_requests = 0;
    unsigned long long policy = 0;
    bool done = false;
    auto advance = [&]() { if (++stage == n_stages) { stage = 0; phase ^= 1u; } };
    // hand out the next unit (or the end marker); false once the end marker has been posted
    auto produce = [&]() -> bool {
      for (;;) {
        unsigned long long slice_end = slice0 + slice_len < total ? slice0 + slice_len : total;
        if (pos >= slice_end) {   // next slice of this CTA
          slice0 += (unsigned long long)gridDim.x * slice_len;
          pos = slice0;
          if (slice0 < total) continue;
          mbar_wait(&s_empty[stage], phase ^ 1u);
          s_meta[stage] = kMetaEnd;
          mbar_arrive(&s_full[stage]);
          return false;
        }
        int ri = 0;
#pragma unroll 1
        for (int k = 1; k < (int)batch.n_readings; ++k) if (pos >= batch.first_point[k]) ri = k;
        const DevReading & rd = batch.r[ri];
        const unsigned long long base = pos - batch.first_point[ri];
        // what is left of the slice inside this reading, in equal units of at most kMarkChunk points (multiples of 64, so
        // that a unit splits into two halves of whole warps)
        const unsigned long long piece_end = slice_end < batch.first_point[ri + 1] ? slice_end : batch.first_point[ri + 1];
        const unsigned long long piece = piece_end - pos;
        const unsigned long long n_units = (piece + kMarkChunk - 1) / kMarkChunk;
        unsigned long long want = ((piece + n_units - 1) / n_units + 63) & ~63ull;
        if (want > piece) want = piece;
        pos += want;
        // points of the reading: the host's count, or (output of the VOXEL pre-filter) a device counter below that bound
        unsigned long long n_pts = rd.n_points;
        if (rd.n_points_dev) { const unsigned long long nd = ld_volatile_u64(rd.n_points_dev); if (nd < n_pts) n_pts = nd; }
        if (base >= n_pts) continue;   // beyond the actual count: nothing to hand out
        const uint32_t cnt = (uint32_t)(n_pts - base < want ? n_pts - base : want);
        const uint32_t half = (((cnt + 1) >> 1) + 31u) & ~31u;
        mbar_wait(&s_empty[stage], phase ^ 1u);
        // a filling queue is flushed early (large or scattered clouds): the marker travels in-band, so every consumer
        // sees it between the same two units
        if (*reinterpret_cast<volatile unsigned int *>(&s_qn) >= (unsigned)kMarkQueueFlush &&
            *reinterpret_cast<volatile unsigned int *>(&s_flushes) == flush_requests) {
          ++flush_requests;
          s_meta[stage] = kMetaFlush;
          mbar_arrive(&s_full[stage]);
          advance();
          mbar_wait(&s_empty[stage], phase ^ 1u);
        }
        s_meta[stage] = (rd.has_tf ? kMetaTf : 0u) | ((uint32_t)ri << 24) | ((half >> 5) << 16) | cnt;
        s_base[stage] = base;
        if constexpr (kVec4) {
          mbar_arrive_expect_tx(&s_full[stage], cnt * 16u);
          bulk_load(s_dyn + (size_t)stage * kMarkChunk * 16, rd.data + base * 16ull, cnt * 16u, &s_full[stage], policy);
        } else {
          mbar_arrive(&s_full[stage]);
        }
        advance();
        return true;
      }
    };
    if (lane == 0) {
      for (int s = 0; s < n_stages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kMarkConsumerWarps); }
      mbar_fence_init();
      s_qn = 0; s_flushes = 0; s_drop[0] = s_drop[1] = 0;
      s_cm[0] = 0; s_cm[1] = s_cm[2] = INT_MAX; s_cm[3] = s_cm[4] = INT_MIN;
      policy = l2_evict_first_policy();
      // the first copies go out while the consumer warps are still clearing the filter (s_qn is 0: no flush marker yet)
      for (int s = 0; s < n_stages && !done; ++s) done = !produce();
    }
    __syncwarp();
    named_bar_sync(3, kMarkThreads);
    if (lane == 0) { while (!done) done = !produce(); }
    __syncwarp();
  } else {
    // ===================================================================================================================
    // consumer warps
    // ===================================================================================================================
    if (threadIdx.x < kMaxBatchReadings) {
      const DevReading & r = batch.r[threadIdx.x < batch.n_readings ? threadIdx.x : 0];
      s_rnow[threadIdx.x] = r.now;
      const bool zwin = r.filter == 2;
      float * sp = s_par[threadIdx.x];
      sp[0] = r.range2_reject; sp[1] = r.range2_accept; sp[2] = r.near_accept; sp[3] = r.oxf;
      sp[4] = r.oyf; sp[5] = r.ozf; sp[6] = zwin ? r.zmin_f : -CUDART_INF_F; sp[7] = zwin ? r.zmax_f : CUDART_INF_F;
      sp[8] = __int_as_float(r.ref_ix - kCodeBiasXY); sp[9] = __int_as_float(r.ref_iy - kCodeBiasXY);
      sp[10] = __int_as_float(r.ref_iz - kCodeBiasZ); sp[11] = __uint_as_float(threadIdx.x << 28);
    }
    // an entry starts out holding a code that hashes to ANOTHER entry, so no lookup can ever match it
    for (int i = threadIdx.x; i < kMarkFilter; i += kMarkConsumers) s_filter[i] = i == 0 ? 1u : 0u;
    named_bar_sync(3, kMarkThreads);   // barriers initialised (producer), filter cleared
    TRACE(0, 1);

    bool grid_ok = false;       // this thread has waited for the predecessor grid (the sweep)
    auto sync_with_predecessor = [&]() { if (!grid_ok) { pdl_wait(); grid_ok = true; } };
    unsigned int drop_range = 0, drop_nonfinite = 0;
    const float vs_f = batch.vs_f, c_lo = batch.c_lo, c_hi = batch.c_hi;   // (float)vs and the two ends of 1/vs, see launch_mark
    const uint32_t my_point = smem_u32(s_dyn) + threadIdx.x * 16u;       // this thread's first point of stage 0
    const uint32_t filter_addr = smem_u32(s_filter), qn_addr = smem_u32(&s_qn);
    const bool sharded = g.shard_count > 1;

    // ---- flush: write the queued voxels to the store, empty the queue (and, between units, the filter) -------------------
    auto flush = [&](const bool final) {
      named_bar_sync(1, kMarkConsumers);   // every consumer warp is done with the units before the flush point
      if (final) TRACE(0, 3);
      sync_with_predecessor();
      // launch_dependents only AFTER our own wait: the next cycle's sweep (and the Mark behind it) cannot become resident
      // before the grid preceding this one has completed, which bounds the programmatic-launch chain to two kernels in flight
      if (final) pdl_launch_dependents();
      const int n_queued = (int)min(*reinterpret_cast<volatile unsigned int *>(&s_qn), (unsigned)kMarkQueue);
      // per queued voxel: one probe of the global leaf hash, one 32-bit OR into the leaf's mask, one 64-bit max (batched
      // launch, monotone clocks) or a plain store (per-reading launches: exact overwrite semantics of MarkGridPoint,
      // .cpp:421-430) of the mark time
      for (int i = threadIdx.x; i < n_queued; i += kMarkConsumers) {
        const uint32_t code = s_queue[i];
        const int ri = (int)(code >> 28);
        const int x = (int)((code >> 18) & 1023u) + __float_as_int(s_par[ri][8]);
        const int y = (int)((code >> 8) & 1023u) + __float_as_int(s_par[ri][9]);
        const int z = (int)(code & 255u) + __float_as_int(s_par[ri][10]);
        const int bx = x >> 3, by = y >> 3, bz = z >> 3;
        const uint32_t slot = leaf_find_or_insert(g, pack_leaf_key(bx, by, bz), bx, by, batch.min_now);
        if (slot == kOverflowSlot) continue;
        const unsigned l = (unsigned)(((x & 7) * 8 + (y & 7)) * 8 + (z & 7));
        const double now = s_rnow[ri];
        red_or_u32(&g.leaf_mask[(size_t)slot * 16 + (l >> 5)], 1u << (l & 31));
        double * tp = g.leaf_time + (size_t)slot * kLeafVoxels + l;
        if (use_max) red_max_s64(reinterpret_cast<unsigned long long *>(tp), (unsigned long long)__double_as_longlong(now));
        else {
          *tp = now;
          if (g.leaf_tmin[slot] > now) g.leaf_tmin[slot] = -CUDART_INF;   // clocks are not monotone here
        }
      }
      if (final) TRACE(0, 4);
      if (!final) {
        named_bar_sync(1, kMarkConsumers);   // the queue has been read
        for (int i = threadIdx.x; i < kMarkFilter; i += kMarkConsumers) s_filter[i] = i == 0 ? 1u : 0u;
        if (threadIdx.x == 0) { s_qn = 0; atomicAdd(&s_flushes, 1u); }
        named_bar_sync(1, kMarkConsumers);
      }
    };

    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (;; first = false) {
      mbar_wait(&s_full[stage], phase);
      const uint32_t meta = s_meta[stage];
      if (meta & kMetaEnd) break;
      if (meta & kMetaFlush) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);
        flush(false);
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        continue;
      }
      const int ri = (int)((meta >> 24) & 15u);
      const unsigned in_chunk = meta & 0xFFFFu;
      const unsigned half = ((meta >> 16) & 0xFFu) << 5;
      if (warp * 32u >= half) {   // this warp has no points in the unit
        __syncwarp();   // every lane has seen the stage's phase: after the arrival below the barrier may complete again
        if (lane == 0) mbar_arrive(&s_empty[stage]);
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        continue;
      }
      const float4 pa = *reinterpret_cast<const float4 *>(&s_par[ri][0]), pb = *reinterpret_cast<const float4 *>(&s_par[ri][4]);
      const float4 pc = *reinterpret_cast<const float4 *>(&s_par[ri][8]);
      PointXYZ pts[2];
      if constexpr (kVec4) {
        const uint32_t addr = my_point + (uint32_t)stage * (uint32_t)(kMarkChunk * 16);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr + (uint32_t)k * (half * 16u)));
          pts[k].x = v.x; pts[k].y = v.y; pts[k].z = v.z;
          pts[k].ok = (unsigned)k * half + threadIdx.x < in_chunk;
        }
      } else {
        const DevReading & par = batch.r[ri];
        const unsigned long long base = s_base[stage];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const unsigned idx = (unsigned)k * half + threadIdx.x;
          pts[k] = load_point_generic(par, base + idx, idx < in_chunk);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[stage]);   // the stage may be refilled: this warp holds its points in registers
      if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      if (first) TRACE(0, 2);
      if (meta & kMetaTf) {   // cloud still in the sensor frame: tf2::doTransform first (measurement_buffer.cpp:141-146)
        const DevReading & par = batch.r[ri];
#pragma unroll
        for (int k = 0; k < 2; ++k) transform_point(par.tf, pts[k].x, pts[k].y, pts[k].z);
      }
      const float range2_reject = pa.x, range2_accept = pa.y, near_accept = pa.z, oxf = pa.w, oyf = pb.x, ozf = pb.y, zlo = pb.z, zhi = pb.w;
      // ---- single-precision screen of the range and z tests, both points of the pair interleaved -----------------------------
      float d2f[2];
      bool cand[2], nonfin[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float fx = pts[k].x, fy = pts[k].y, fz = pts[k].z;
        // a float evaluation of the squared distance whose error is far below the margins of the thresholds: above
        // range2_reject the exact test certainly rejects, inside [near_accept, range2_accept] it certainly accepts.  A NaN or
        // infinite coordinate makes d2f NaN or +inf.
        const float ax = __fsub_rn(fx, oxf), ay = __fsub_rn(fy, oyf), az = __fsub_rn(fz, ozf);
        d2f[k] = fmaf(az, az, fmaf(ay, ay, __fmul_rn(ax, ax)));
        // PassThrough z window (reference measurement_buffer.cpp:165-175) on exactly equivalent float thresholds
        const bool zok = !((fz < zlo) | (fz > zhi));
        nonfin[k] = pts[k].ok & !(d2f[k] < CUDART_INF_F);
        cand[k] = pts[k].ok & zok & !(d2f[k] > range2_reject);
      }
      // whole warps of rejected points (floor, ceiling, beyond the obstacle range) stop here
      if (!__any_sync(kFull, cand[0] | cand[1] | nonfin[0] | nonfin[1])) continue;
      // ---- single-precision quantisation ---------------------------------------------------------------------------------------
      int ix[2], iy[2], iz[2];
      bool take[2], slow[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float fx = pts[k].x, fy = pts[k].y, fz = pts[k].z;
        const bool inwin = (d2f[k] <= range2_accept) & (d2f[k] >= near_accept);
        // .cpp:293-303 in float: negatives are shifted by one voxel (z tests y's sign — sic), scaled by 1/vs, truncated toward zero
        const float Z = fy < 0 ? __fsub_rn(fz, vs_f) : fz;
        uint32_t diff = 0;
        bool tiny = false;
        ix[k] = quantise_xy_fast(fx, c_lo, c_hi, diff, tiny);
        iy[k] = quantise_xy_fast(fy, c_lo, c_hi, diff, tiny);
        iz[k] = quantise_axis_fast(Z, c_lo, c_hi, diff);
        const bool decided = inwin & (diff == 0) & !tiny;
        take[k] = cand[k] & decided;
        // everything the fast path cannot decide, and everything that is not finite (counted), takes the exact sequence
        slow[k] = nonfin[k] | (cand[k] & !decided);
      }
      // ---- the exact double sequence for the points the fast path could not decide (rare; whole warps skip it) ------------------
      if (__any_sync(kFull, slow[0] | slow[1])) {
        const DevReading & par = batch.r[ri];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (slow[k]) {
            const int4 q = quantise_exact(pts[k].x, pts[k].y, pts[k].z, par.ox, par.oy, par.oz, par.mark_range_2, g.vs, g.inv_vs, zlo, zhi);
            ix[k] = q.x; iy[k] = q.y; iz[k] = q.z;
            take[k] = q.w == 1;
            drop_range += q.w == 2 ? 1u : 0u;
            drop_nonfinite += q.w == 3 ? 1u : 0u;
          }
        }
      }
      if (sharded) {   // spatial shard: foreign leaves are dropped
#pragma unroll
        for (int k = 0; k < 2; ++k) take[k] = take[k] && shard_owns(ix[k] >> 3, g.shard_index, g.shard_count, g.shard_width);
      }
      // ---- CTA-local dedupe: 32-bit voxel code, direct-mapped filter, queue of voxels to write --------------------------------
      const int fx0 = __float_as_int(pc.x), fy0 = __float_as_int(pc.y), fz0 = __float_as_int(pc.z);
      const uint32_t ri28 = __float_as_uint(pc.w);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        bool push = false;
        uint32_t code = 0;
        if (take[k]) {
          const uint32_t rx = (uint32_t)(ix[k] - fx0), ry = (uint32_t)(iy[k] - fy0), rz = (uint32_t)(iz[k] - fz0);
          if ((((rx | ry) >> 10) | (rz >> 8)) == 0u) {
            code = ri28 | (rx << 18) | (ry << 8) | rz;
            const uint32_t fa = filter_addr + (((code * 0x9E3779B1u) >> kMarkFilterShift) << 2);
            uint32_t cur;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(fa));
            if (cur != code) {   // not seen (or forgotten): remember it and queue it.  The exchange lets exactly one of the
              uint32_t old;      // lanes (and warps) that find the same voxel missing at the same time do the queueing
              asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(fa), "r"(code) : "memory");
              push = old != code;
            }
          } else {   // farther from the sensor than the code reaches: straight to the global store
            sync_with_predecessor();
            mark_voxel_direct(g, ix[k] >> 3, iy[k] >> 3, iz[k] >> 3, (unsigned)(((ix[k] & 7) * 8 + (iy[k] & 7)) * 8 + (iz[k] & 7)), s_rnow[ri], use_max,
                              batch.min_now);
          }
        }
        const unsigned pm = __ballot_sync(kFull, push);
        if (pm) {   // warp-aggregated append
          unsigned q0 = 0;
          if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(q0) : "r"(qn_addr), "r"((unsigned)__popc(pm)) : "memory");
          q0 = __shfl_sync(kFull, q0, 0) + __popc(pm & ((1u << lane) - 1u));
          if (push) {
            if (q0 < (unsigned)kMarkQueue) s_queue[q0] = code;
            else {   // (cannot happen while the flush threshold holds; kept for safety)
              sync_with_predecessor();
              mark_voxel_direct(g, ix[k] >> 3, iy[k] >> 3, iz[k] >> 3, (unsigned)(((ix[k] & 7) * 8 + (iy[k] & 7)) * 8 + (iz[k] & 7)), s_rnow[ri],
                                use_max, batch.min_now);
            }
          }
        }
      }
    }
    flush(true);
    TRACE(0, 5);

    // dropped-point statistics: warp reduce, shared-memory add, one fire-and-forget global reduction per CTA
    if (__any_sync(kFull, (drop_range | drop_nonfinite) != 0)) {
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
    named_bar_sync(1, kMarkConsumers);
    if (threadIdx.x == 0) {
      if (s_drop[0]) red_add_u64(&g.ctr->dropped_range, (unsigned long long)s_drop[0]);   // rare: points beyond +-2^23 voxels
      if (s_drop[1]) red_add_u64(&g.ctr->dropped_nonfinite, (unsigned long long)s_drop[1]);
      TRACE(0, 6);
    }
  }

  // =====================================================================================================================
  // costmap pass (fused cycle), all 32 warps: it needs the sweep's column counts, nothing of Mark
  // =====================================================================================================================
  if constexpr (kCostmap) {
    TRACE(3, 0);
    pdl_wait();
    const uint32_t n_tiles = min(ld_volatile_u32(&g.ctr->tile_n), g.tile_capacity);
    TRACE(3, 1);
    constexpr uint32_t kCmWarps = kMarkThreads / 32;
    const uint32_t cw = blockIdx.x * kCmWarps + warp, n_cw = gridDim.x * kCmWarps;
    unsigned cm_n = 0;
    int cm_mnx = INT_MAX, cm_mny = INT_MAX, cm_mxx = INT_MIN, cm_mxy = INT_MIN;
    // tiles per warp and round trip: as few as keep every warp of the grid busy (the pass is a chain of dependent round
    // trips, not a stream)
    if (n_tiles <= n_cw) {
      if (cw < n_tiles) costmap_tile_batch<1, kCmBits>(g, cm, costmap, cw, n_tiles, lane, cm_n, cm_mnx, cm_mny, cm_mxx, cm_mxy);
    } else if (n_tiles <= 2 * n_cw) {
      if (cw * 2 < n_tiles) costmap_tile_batch<2, kCmBits>(g, cm, costmap, cw * 2, n_tiles, lane, cm_n, cm_mnx, cm_mny, cm_mxx, cm_mxy);
    } else {
      for (uint32_t t0 = cw * 4; t0 < n_tiles; t0 += n_cw * 4)
        costmap_tile_batch<4, kCmBits>(g, cm, costmap, t0, n_tiles, lane, cm_n, cm_mnx, cm_mny, cm_mxx, cm_mxy);
    }
    TRACE(3, 2);
    // marked cells and bounds: warp shuffle, then shared memory, then ONE set of global atomics per CTA (thousands of warps
    // updating the same five words would serialise in L2)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      cm_n += __shfl_xor_sync(kFull, cm_n, d);
      cm_mnx = min(cm_mnx, __shfl_xor_sync(kFull, cm_mnx, d)); cm_mny = min(cm_mny, __shfl_xor_sync(kFull, cm_mny, d));
      cm_mxx = max(cm_mxx, __shfl_xor_sync(kFull, cm_mxx, d)); cm_mxy = max(cm_mxy, __shfl_xor_sync(kFull, cm_mxy, d));
    }
    if (lane == 0 && cm_n) {
      atomicAdd(&s_cm[0], (int)cm_n);
      atomicMin(&s_cm[1], cm_mnx); atomicMin(&s_cm[2], cm_mny); atomicMax(&s_cm[3], cm_mxx); atomicMax(&s_cm[4], cm_mxy);
    }
    named_bar_sync(2, kMarkThreads);
    TRACE(3, 3);
    if (threadIdx.x == 0) {
      if (s_cm[0]) {
        atomicAdd(&g.ctr->cm_acc.n_marked, (unsigned long long)s_cm[0]);
        atomicMin(&g.ctr->cm_acc.mk_min_x, s_cm[1]); atomicMin(&g.ctr->cm_acc.mk_min_y, s_cm[2]);
        atomicMax(&g.ctr->cm_acc.mk_max_x, s_cm[3]); atomicMax(&g.ctr->cm_acc.mk_max_y, s_cm[4]);
      }
      __threadfence();
      if (atomicAdd(&g.ctr->cm_ticket, 1u) == gridDim.x - 1) {   // last CTA publishes the result and re-arms the accumulators
        __threadfence();
        costmap_publish_result(g.ctr);
        g.ctr->cm_ticket = 0;
        __threadfence();
      }
      TRACE(3, 4);
    }
  }
}

// bulk load of (index, time) pairs — MarkGridPoint semantics (overwrite), for tests / restore
__global__ void __launch_bounds__(256)
load_voxels_kernel(const GridView g, const int32_t * __restrict__ vx, const int32_t * __restrict__ vy,
                   const int32_t * __restrict__ vz, const double * __restrict__ vt, unsigned long long n)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int ix = vx[i], iy = vy[i], iz = vz[i];
    const int bx = ix >> 3, by = iy >> 3, bz = iz >> 3;
    if (!(leaf_coord_ok(bx) && leaf_coord_ok(by) && leaf_coord_ok(bz))) {
      atomicAdd(&g.ctr->dropped_range, 1ull);
    } else if (shard_owns(bx, g.shard_index, g.shard_count, g.shard_width)) {
      const uint32_t slot = leaf_find_or_insert(g, pack_leaf_key(bx, by, bz), bx, by, CUDART_INF);
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
// host-callable launchers
// =====================================================================================
namespace {
int env_int(const char * name, int dflt, int lo, int hi)
{
  const char * v = std::getenv(name);
  if (!v || !*v) return dflt;
  const int x = std::atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
// bulk-copy stages per CTA (STVL_MARK_STAGES is a tuning knob for profiling; the default is what the measurements in
// DESIGN.md were taken with)
int mark_stages(bool vec4) { static const int v = env_int("STVL_MARK_STAGES", 3, 2, kMarkStagesMax); return vec4 ? v : 2; }
size_t mark_smem_bytes(bool vec4) { return (vec4 ? (size_t)mark_stages(true) * kMarkChunk * 16 : 0) + (size_t)(kMarkFilter + kMarkQueue) * 4; }
}  // namespace

// one CTA per SM; small batches get one CTA per unit
uint32_t mark_grid_for(const MarkBatch & batch, int sm_count)
{
  return std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)sm_count, batch.total_blocks));
}
// the batch's point index space (readings laid end to end) and the slice length: kMarkSlices slices per CTA, a multiple of 64
static void mark_set_slices(MarkBatch & b, uint32_t grid)
{
  unsigned long long total = 0;
  for (uint32_t k = 0; k < b.n_readings; ++k) { b.first_point[k] = total; total += b.r[k].n_points; }
  for (uint32_t k = b.n_readings; k <= (uint32_t)kMaxBatchReadings; ++k) b.first_point[k] = total;
  const unsigned long long n_slices = (unsigned long long)grid * kMarkSlices;
  unsigned long long len = ((total + n_slices - 1) / n_slices + 63ull) & ~63ull;
  b.slice_len = std::max<unsigned long long>(len, 1024ull);
}
// column tiles the fused variant's costmap pass can take without becoming the critical path of the launch
uint32_t mark_costmap_tiles_capacity(const MarkBatch & batch, int sm_count)
{
  return mark_grid_for(batch, sm_count) * (uint32_t)(kMarkThreads / 32) * 64u;
}

template <bool kVec4, bool kCostmap, bool kCmBits>
static cudaError_t launch_mark_variant(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                                       const CostmapDev & cm, uint8_t * costmap)
{
  static bool configured = false;
  const size_t smem = mark_smem_bytes(kVec4);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mark_kernel<kVec4, kCostmap, kCmBits>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // All kernels of the cycle ask for the same (maximum) shared-memory carve-out: they overlap on the same SMs under
    // programmatic dependent launch, and an SM only changes its L1 / shared split when it is empty.
    cudaFuncSetAttribute(mark_kernel<kVec4, kCostmap, kCmBits>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  // constants of the single-precision quantisation (quantise_axis_fast), rounded here exactly as the proof there assumes
  MarkBatch b = batch;
  b.vs_f = (float)g.vs;                                              // exact: vs is a float widened to double
  b.c_lo = (float)g.inv_vs * 0.999999701976776123046875f;            // (float)(1/vs) * (1 - 5 * 2^-24), one rounding
  b.c_hi = (float)g.inv_vs * 1.00000035762786865234375f;             // (float)(1/vs) * (1 + 3 * 2^-23), one rounding
  mark_set_slices(b, mark_grid_for(batch, sm_count));
  for (uint32_t k = 0; k < b.n_readings; ++k) {   // frame of the 32-bit voxel codes: the voxel the reading's origin falls into
    const double o[3] = {b.r[k].ox, b.r[k].oy, b.r[k].oz};
    int32_t ref[3];
    for (int a = 0; a < 3; ++a) {
      const double q = std::floor(o[a] / g.vs);
      ref[a] = (std::isfinite(q) && std::fabs(q) < 1e9) ? (int32_t)q : 0;
    }
    b.r[k].ref_ix = ref[0]; b.r[k].ref_iy = ref[1]; b.r[k].ref_iz = ref[2];
  }
  return launch_ex(mark_kernel<kVec4, kCostmap, kCmBits>, mark_grid_for(batch, sm_count), (unsigned)kMarkThreads, smem, s, pdl, g, b, cm, costmap,
                   mark_stages(kVec4));
}

// pdl: the caller guarantees that the kernel's inputs (clouds) were complete before the PREDECESSOR of this launch in the
// stream started, and that the predecessor is this library's sweep kernel.  cm / costmap != NULL: fused costmap pass
// (cm_bits: into a cell bitmap).
cudaError_t launch_mark(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                        const CostmapDev * cm, uint8_t * costmap, bool cm_bits)
{
  if (batch.total_blocks == 0) return cudaSuccess;
  bool all_vec4 = true;
  for (uint32_t i = 0; i < batch.n_readings; ++i) all_vec4 = all_vec4 && batch.r[i].vec4 != 0;
  const CostmapDev none{};
  if (cm && costmap) {
    if (cm_bits)
      return all_vec4 ? launch_mark_variant<true, true, true>(g, batch, sm_count, s, pdl, *cm, costmap)
                      : launch_mark_variant<false, true, true>(g, batch, sm_count, s, pdl, *cm, costmap);
    return all_vec4 ? launch_mark_variant<true, true, false>(g, batch, sm_count, s, pdl, *cm, costmap)
                    : launch_mark_variant<false, true, false>(g, batch, sm_count, s, pdl, *cm, costmap);
  }
  return all_vec4 ? launch_mark_variant<true, false, false>(g, batch, sm_count, s, pdl, none, nullptr)
                  : launch_mark_variant<false, false, false>(g, batch, sm_count, s, pdl, none, nullptr);
}
uint32_t mark_blocks_for(unsigned long long n_points)
{
  return (uint32_t)((n_points + kMarkChunk - 1) / kMarkChunk);
}
cudaError_t launch_load_voxels(const GridView & g, const int32_t * x, const int32_t * y, const int32_t * z, const double * t,
                               unsigned long long n, cudaStream_t s)
{
  if (n == 0) return cudaSuccess;
  load_voxels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, x, y, z, t, n);
  return cudaGetLastError();
}

}  // namespace stvl
