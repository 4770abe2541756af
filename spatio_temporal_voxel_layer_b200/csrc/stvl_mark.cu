// stvl_mark.cu — Mark / operator() (reference spatio_temporal_voxel_grid.cpp:254-309) as a bulk-copy pipeline kernel for sm_100a.
//
// One persistent CTA of 32 warps per SM, every warp a self-contained pipeline (no producer warp, no CTA-wide barrier while
// the clouds stream).  The batch's points (all readings laid end to end) are cut into UNITS of 128 points (2 KB of packed
// float4 points, never across two readings).  A warp owns two 2 KB buffers and one mbarrier each:
//   * lane 0 streams a unit global -> shared memory with the bulk-copy engine (cp.async.bulk, completion on the buffer's
//     mbarrier as transaction bytes; SASS UBLKCP / SYNCS), evict-first in L2;
//   * the warp waits for its OLDER buffer, pulls four points per lane into registers, re-issues that buffer right away
//     (64 units = 128 KB in flight per SM) and runs the reference's per-point arithmetic on the four points: range /
//     z-window tests and voxel quantisation on a single-precision fast path that is PROVEN to agree with the reference's
//     double arithmetic (see quantise_axis_fast), the exact double sequence for the few points it cannot decide;
//   * survivors go through a CTA-wide direct-mapped filter of 32-bit voxel codes (dedupe) into a queue of voxels to write.
// Hand-out of the units.  A launch that has the GPU to itself deals them round-robin to the CTAs — every CTA gets the same
// number of points from all over the batch (floor rows that the range screen rejects are cheap, wall rows are not) — and,
// inside a CTA, round-robin to its warps: no atomics.  A launch of the fused cycle starts its CTAs one by one as the sweep
// leaves the SMs; there a CTA takes GROUPS of 32 units from a global cursor (never reset: every launch is told where it
// stands) and hands a group's units to whichever warp asks next.
// The first copies are in flight before the CTA has set up its tables.  Nothing before the flush touches the grid, so all
// of the streaming may overlap the sweep that precedes Mark on the stream (programmatic dependent launch).  The flush
// resolves every queued voxel in the global leaf hash once and pushes one 32-bit OR into the leaf's mask and one 64-bit max
// (or store) of the mark time.  A queue that fills up (huge or scattered clouds) is flushed early: the warp that notices
// raises a flag, every warp joins at its next unit.  In the fused cycle the first eight warps of every CTA run the costmap
// pass over the sweep's column counts BEFORE they join the unit stream.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>

#include "stvl_common.cuh"

namespace stvl {

constexpr int kMarkWarps = 32;                   // warps of a Mark CTA that has its SM to itself ...
constexpr int kMarkWarpsShared = 16;             // ... and of one that shares it with a CTA of the sweep kernel (fused cycle): 512 threads
                                                 // x 64 registers each, so that one CTA of either kernel fits an SM at the same time
constexpr int kMarkPPT = 4;                      // points per lane per unit
constexpr int kMarkChunk = 32 * kMarkPPT;        // a unit: 128 points = 2 KB of packed float4 points, one warp-private buffer
constexpr int kMarkCostmapWarps = 8;             // warps of a CTA that run the fused cycle's costmap pass before they join the unit stream
constexpr int kMarkGroup = 32;                  // dynamic hand-out: units a CTA claims at a time
constexpr int kMarkBufs = 2;                     // buffers (units in flight) per warp
constexpr int kMarkFilter = 8192;                // direct-mapped filter of voxel codes, entries (a power of two)
constexpr int kMarkFilterShift = 19;             // 32 - log2(kMarkFilter)
constexpr int kMarkQueue = 8192;                 // voxels to write, entries
constexpr int kMarkQueueFlush = 4096;            // an early flush is requested at this fill; every warp still finishes the unit
                                                 // it is working on (at most 32 x 128 more voxels) before it joins
// voxel code: reading (4 bits) | x (10) | y (10) | z (8), relative to the voxel of the reading's origin
constexpr int kCodeBiasXY = 512, kCodeBiasZ = 128;
static_assert(kMarkQueueFlush + kMarkWarps * kMarkChunk <= kMarkQueue, "queue sized for the warps' lag behind a flush request");
static_assert((1 << (32 - kMarkFilterShift)) == kMarkFilter, "filter index = upper bits of the multiplicative hash");
// what a warp knows about the unit in one of its buffers: points (8 bits, 0 = empty unit) | reading << 8 | sensor-frame cloud
constexpr uint32_t kMetaEnd = 0xFFFFFFFFu, kMetaTf = 1u << 16;
constexpr uint32_t kNoUnit = 0xFFFFFFFFu;
constexpr uint32_t kNoCode = 0xFFFFFFFFu;   // "nothing to queue" (the one voxel whose code this is goes straight to the store)
static_assert(kMarkPPT == 4 && kMarkChunk == 128, "the unit arithmetic is written for four points per lane");

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

// a queued voxel code back to (leaf, voxel-in-leaf): the inverse of the packing in the kernel below
struct CodeVoxel { int bx, by, bz; unsigned l; int ri; };
__device__ __forceinline__ CodeVoxel decode_code(uint32_t code, const float (* par)[12])
{
  CodeVoxel v;
  v.ri = (int)(code >> 28);
  const int x = (int)((code >> 18) & 1023u) + __float_as_int(par[v.ri][8]);
  const int y = (int)((code >> 8) & 1023u) + __float_as_int(par[v.ri][9]);
  const int z = (int)(code & 255u) + __float_as_int(par[v.ri][10]);
  v.bx = x >> 3; v.by = y >> 3; v.bz = z >> 3;
  v.l = (unsigned)(((x & 7) * 8 + (y & 7)) * 8 + (z & 7));
  return v;
}
// a voxel code that found the queue full: straight to the global store
__device__ __noinline__ void mark_code_direct(const GridView & g, uint32_t code, const float (* par)[12], const double * rnow, bool use_max, double tmin_init)
{
  const CodeVoxel v = decode_code(code, par);
  mark_voxel_direct(g, v.bx, v.by, v.bz, v.l, rnow[v.ri], use_max, tmin_init);
}

// kVec4: every reading of the batch is a packed x,y,z,pad float4 cloud (the RGBD / depth-image layout), streamed through
// the bulk-copy pipeline.  Other layouts are read from global memory by the lanes themselves (same unit hand-out, no
// staging).
// kCostmap: the fused update cycle — this launch also runs the costmap pass over the column counts of the sweep that
// precedes it (they are independent of Mark: the costmap reflects the pre-Mark survivors, reference .cpp:149-224 then
// spatio_temporal_voxel_layer.cpp:699-718); the costmap bytes were reset by the sweep kernel.  kCmBits: the pass writes a
// cell bitmap instead of bytes (sharded cycle).
//
// Dynamic shared memory, one base register for everything the unit loop touches (byte offsets):
constexpr uint32_t kSmFilter = 0;                                   // kMarkFilter x u32
constexpr uint32_t kSmQueue = kSmFilter + kMarkFilter * 4;          // kMarkQueue x u32
constexpr uint32_t kSmPar = kSmQueue + kMarkQueue * 4;              // per reading 12 x f32: the single-precision screen {range2_reject, range2_accept,
                                                                    //   near_accept, ox} {oy, oz, zlo, zhi} and the voxel code's frame {ref_ix - bias,
                                                                    //   ref_iy - bias, ref_iz - bias, reading << 28}
constexpr uint32_t kSmNow = kSmPar + kMaxBatchReadings * 48;        // per reading f64: its clock sample
constexpr uint32_t kSmQn = kSmNow + kMaxBatchReadings * 8;          // u32 queued voxels (may run past kMarkQueue: the excess went straight to the store)
constexpr uint32_t kSmFlushReq = kSmQn + 4;                         // u32 a warp has seen the queue above kMarkQueueFlush
constexpr uint32_t kSmWarp = kSmQn + 64;                            // per warp: kMarkBufs buffers of kMarkChunk points, then their mbarriers
constexpr uint32_t kSmWarpBar = kMarkBufs * kMarkChunk * 16;        //   (offset of the barriers inside a warp's region)
constexpr uint32_t kSmWarpStride = kSmWarpBar + 16;                 //   4112 B
static_assert(kSmWarp % 16 == 0 && kSmWarpStride % 16 == 0 && kMarkBufs == 2, "128-bit shared loads of the points");

template <bool kVec4, bool kCostmap, bool kCmBits, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, kWarps == kMarkWarps ? 1 : 2)
mark_kernel(const __grid_constant__ GridView g, const __grid_constant__ MarkBatch batch, const __grid_constant__ CostmapDev cm,
            uint8_t * __restrict__ costmap)
{
  constexpr int kThreads = kWarps * 32;
  extern __shared__ __align__(128) unsigned char s_dyn[];
  uint32_t * const s_filter = reinterpret_cast<uint32_t *>(s_dyn + kSmFilter);
  uint32_t * const s_queue = reinterpret_cast<uint32_t *>(s_dyn + kSmQueue);
  float (* const s_par)[12] = reinterpret_cast<float (*)[12]>(s_dyn + kSmPar);
  double * const s_rnow = reinterpret_cast<double *>(s_dyn + kSmNow);
  volatile unsigned int * const s_qn = reinterpret_cast<volatile unsigned int *>(s_dyn + kSmQn);
  volatile unsigned int * const s_flush_req = reinterpret_cast<volatile unsigned int *>(s_dyn + kSmFlushReq);
  __shared__ unsigned int s_done;          // warps that have run out of units
  // dynamic hand-out: {end << 32 | next} of the group of units the CTA is working through, the group claimed ahead of need
  // (index + 1, 0 = not there yet), and "no more groups"
  __shared__ unsigned long long s_units;
  __shared__ unsigned int s_claimed, s_exhausted;
  __shared__ unsigned int s_drop[2];
  __shared__ int s_cm[5];                  // fused costmap pass: marked cells and their bounds, gathered per CTA

  TRACE(0, 0);
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const bool use_max = batch.use_atomic_max != 0;
  // the one shared-memory address the unit loop computes with (opaque to the compiler, so it is not formed again and again)
  uint32_t sbase;
  asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(sbase) : "l"(s_dyn));
  const uint32_t wbar = sbase + kSmWarp + warp * kSmWarpStride + kSmWarpBar;   // this warp's two mbarriers
  const uint32_t laddr = wbar - kSmWarpBar + lane * 16u;                        // this lane's first point in buffer 0

  // ---- unit hand-out: units are dealt round-robin to the CTAs, a CTA's units round-robin to its warps ---------------------
  // reading k holds units [first_block[k], first_block[k + 1]) of kMarkChunk points; lane k keeps first_block[k]
  const uint32_t n_tickets = batch.units_per_cta + (blockIdx.x < batch.units_extra ? 1u : 0u);   // units of this CTA (static hand-out)
  const bool dynamic = batch.dynamic_units != 0;
  const uint32_t first_unit = lane < batch.n_readings ? batch.first_block[lane] : 0xFFFFFFFFu;
  // one more group for the CTA (dynamic hand-out; one lane): its index, beyond n_groups when there is none left.  Every CTA
  // claims until it is told so ONCE, so a launch advances the cursor by exactly n_groups - gridDim.x + gridDim.x = n_groups.
  auto claim_group = [&]() -> uint32_t { return gridDim.x + (atomicAdd(g.mark_cursor, 1u) - batch.cursor_base); };
  // the warp's next unit.  Static: the CTA's unit number `ticket`.  Dynamic: the next unit of the CTA's current group; the lane
  // that takes a group's first unit claims the group after it, the lane that finds the group used up installs that one.
  auto next_unit = [&](const uint32_t ticket) -> uint32_t {
    if (!dynamic) return ticket < n_tickets ? blockIdx.x + ticket * gridDim.x : kNoUnit;
    uint32_t u = kNoUnit;
    if (lane == 0) {
      for (;;) {
        const unsigned long long old = atomicAdd(&s_units, 1ull);
        const uint32_t cur = (uint32_t)old, end = (uint32_t)(old >> 32);
        if (cur < end) {
          if ((cur & (uint32_t)(kMarkGroup - 1)) == 0u) *reinterpret_cast<volatile unsigned int *>(&s_claimed) = claim_group() + 1u;
          u = cur;
          break;
        }
        if (*reinterpret_cast<volatile unsigned int *>(&s_exhausted)) break;
        if (cur == end) {   // first to find the group used up: install the one claimed ahead
          unsigned int c;
          while ((c = *reinterpret_cast<volatile unsigned int *>(&s_claimed)) == 0u) { }
          *reinterpret_cast<volatile unsigned int *>(&s_claimed) = 0u;
          __threadfence_block();   // ... before the group below becomes visible: its first taker writes s_claimed again
          const uint32_t grp = c - 1u;
          if (grp >= batch.n_groups) { *reinterpret_cast<volatile unsigned int *>(&s_exhausted) = 1u; break; }
          const uint32_t lo = grp * (uint32_t)kMarkGroup, hi = min(lo + (uint32_t)kMarkGroup, batch.total_blocks);
          atomicExch(&s_units, ((unsigned long long)hi << 32) | lo);
        } else {
          __nanosleep(32);   // another warp is installing the next group
        }
      }
    }
    return __shfl_sync(kFull, u, 0);
  };
  // what buffer b will hold for unit u; lane 0 starts the copy.  Returns the unit's meta word (kMetaEnd: no such unit).
  auto issue = [&](const uint32_t u, const uint32_t b, unsigned long long & base_out) -> uint32_t {
    if (u == kNoUnit) return kMetaEnd;
    const int ri = __popc(__ballot_sync(kFull, u >= first_unit)) - 1;   // lane 0 holds 0: ri >= 0
    const MarkStreamSource & rd = batch.src[ri];   // (the few fields the hand-out needs, packed next to first_block: the batch
                                                   // description is cold in the constant cache when the kernel starts)
    const unsigned long long base = (unsigned long long)(u - __shfl_sync(kFull, first_unit, ri)) * (unsigned)kMarkChunk;
    // points of the reading: the host's count, or (output of the VOXEL pre-filter) a device counter below that bound
    unsigned long long n_pts = rd.n_points;
    if (rd.n_points_dev) { const unsigned long long nd = ld_volatile_u64(rd.n_points_dev); if (nd < n_pts) n_pts = nd; }
    base_out = base;
    if (base >= n_pts) {   // beyond the actual count: an empty unit (its barrier phase completes all the same)
      if constexpr (kVec4) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(wbar + b * 8u) : "memory"); }
      return (uint32_t)ri << 8;
    }
    const uint32_t cnt = n_pts - base < (unsigned long long)kMarkChunk ? (uint32_t)(n_pts - base) : (uint32_t)kMarkChunk;
    if constexpr (kVec4) {
      if (lane == 0) {
        const uint32_t bar = wbar + b * 8u;
        const unsigned long long policy = l2_evict_first_policy();   // the clouds are read once: keep them from displacing the grid's lines
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(cnt * 16u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     :: "r"(laddr + b * (uint32_t)(kMarkChunk * 16)), "l"(rd.data + base * 16ull), "r"(cnt * 16u), "r"(bar), "l"(policy) : "memory");
      }
    }
    return cnt | ((uint32_t)ri << 8) | (rd.has_tf ? kMetaTf : 0u);
  };

  // the first units are on their way before the CTA sets up its tables
  unsigned long long base_cur = 0, base_nxt = 0;   // first point of the unit in its reading (generic layouts read from there)
  if constexpr (kVec4) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(wbar) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(wbar + 8u) : "memory");
      mbar_fence_init();
    }
    __syncwarp();
  }
  // (dynamic hand-out: unit `warp` of the CTA's first group, group blockIdx.x; the claim of its second group is under way
  // while the tables are set up, and every warp takes its second unit from it after the barrier below)
  // The warps that run the costmap pass first (fused variant) take no unit before it; the rest of the first group is handed out
  // like every later group.
  constexpr uint32_t kLateWarps = kCostmap ? (uint32_t)kMarkCostmapWarps : 0u;
  static_assert(kWarps - (int)kLateWarps <= kMarkGroup, "a CTA's first group gives every early warp one unit");
  uint32_t meta_cur, meta_nxt = kMetaEnd;
  uint32_t claimed_first = 0;
  if (dynamic) {
    if (threadIdx.x == 0) claimed_first = claim_group();
    const uint32_t u = blockIdx.x * (uint32_t)kMarkGroup + (warp - kLateWarps);
    meta_cur = issue((warp >= kLateWarps && u < batch.total_blocks) ? u : kNoUnit, 0, base_cur);
  } else {
    meta_cur = issue(next_unit(warp), 0, base_cur);
    meta_nxt = issue(next_unit(warp + kWarps), 1, base_nxt);
  }

  if (warp == 0) {
    if (lane < kMaxBatchReadings) {
      const DevReading & r = batch.r[lane < batch.n_readings ? lane : 0];
      s_rnow[lane] = r.now;
      const bool zwin = r.filter == 2;
      float * sp = s_par[lane];
      sp[0] = r.range2_reject; sp[1] = r.range2_accept; sp[2] = r.near_accept; sp[3] = r.oxf;
      sp[4] = r.oyf; sp[5] = r.ozf; sp[6] = zwin ? r.zmin_f : -CUDART_INF_F; sp[7] = zwin ? r.zmax_f : CUDART_INF_F;
      sp[8] = __int_as_float(r.ref_ix - kCodeBiasXY); sp[9] = __int_as_float(r.ref_iy - kCodeBiasXY);
      sp[10] = __int_as_float(r.ref_iz - kCodeBiasZ); sp[11] = __uint_as_float(lane << 28);
    }
    if (lane == 0) {
      *s_qn = 0; *s_flush_req = 0; s_done = 0; s_drop[0] = s_drop[1] = 0;
      s_claimed = 0; s_exhausted = 0; s_units = 0;
      if (dynamic) {   // what the early warps left of the first group; the group claimed above comes next
        const uint32_t hi = min((blockIdx.x + 1u) * (uint32_t)kMarkGroup, batch.total_blocks);
        const uint32_t lo = min(blockIdx.x * (uint32_t)kMarkGroup + ((uint32_t)kWarps - kLateWarps), hi);
        s_units = ((unsigned long long)hi << 32) | lo;
        s_claimed = claimed_first + 1u;
      }
      s_cm[0] = 0; s_cm[1] = s_cm[2] = INT_MAX; s_cm[3] = s_cm[4] = INT_MIN;
    }
  }
  // an entry starts out holding a code that hashes to ANOTHER entry, so no lookup can ever match it
  for (int i = threadIdx.x; i < kMarkFilter / 4; i += kThreads) reinterpret_cast<uint4 *>(s_filter)[i] = make_uint4(i == 0 ? 1u : 0u, 0u, 0u, 0u);
  named_bar_sync(1, kThreads);
  TRACE(0, 1);

  bool grid_ok = false;       // this thread has waited for the predecessor grid (the sweep)
  auto sync_with_predecessor = [&]() { if (!grid_ok) { pdl_wait(); grid_ok = true; } };
  unsigned int drop_range = 0, drop_nonfinite = 0;
  const float vs_f = batch.vs_f, c_lo = batch.c_lo, c_hi = batch.c_hi;   // (float)vs and the two ends of 1/vs, see launch_mark
  const bool sharded = g.shard_count > 1;

  // ---- flush: write the queued voxels to the store.  Every warp of the CTA takes part; the flush that finds all warps out
  // of units is the last one (returns true) -----------------------------------------------------------------------------------
  auto flush = [&]() -> bool {
    named_bar_sync(1, kThreads);   // every warp is between two units (or out of units)
    const bool final = *reinterpret_cast<volatile unsigned int *>(&s_done) == (unsigned)kWarps;
    if (final) TRACE(0, 3);
    sync_with_predecessor();
    // launch_dependents only AFTER our own wait: the next cycle's sweep (and the Mark behind it) cannot become resident
    // before the grid preceding this one has completed, which bounds the programmatic-launch chain to two kernels in flight
    if (final) pdl_launch_dependents();
    const int n_queued = (int)min(*s_qn, (unsigned)kMarkQueue);
    // per queued voxel: one probe of the global leaf hash, one 32-bit OR into the leaf's mask, one 64-bit max (batched
    // launch, monotone clocks) or a plain store (per-reading launches: exact overwrite semantics of MarkGridPoint,
    // .cpp:421-430) of the mark time
    for (int i = threadIdx.x; i < n_queued; i += kThreads) {
      const CodeVoxel v = decode_code(s_queue[i], s_par);
      const uint32_t slot = leaf_find_or_insert(g, pack_leaf_key(v.bx, v.by, v.bz), v.bx, v.by, batch.min_now);
      if (slot == kOverflowSlot) continue;
      const double now = s_rnow[v.ri];
      red_or_u32(&g.leaf_mask[(size_t)slot * 16 + (v.l >> 5)], 1u << (v.l & 31));
      double * tp = g.leaf_time + (size_t)slot * kLeafVoxels + v.l;
      if (use_max) red_max_s64(reinterpret_cast<unsigned long long *>(tp), (unsigned long long)__double_as_longlong(now));
      else {
        *tp = now;
        if (g.leaf_tmin[slot] > now) g.leaf_tmin[slot] = -CUDART_INF;   // clocks are not monotone here
      }
    }
    if (final) { TRACE(0, 4); return true; }
    // the filter keeps what it knows (a voxel written once needs no second write); the queue starts over
    named_bar_sync(1, kThreads);   // the queue has been read
    if (threadIdx.x == 0) { *s_qn = 0; *s_flush_req = 0; }
    named_bar_sync(1, kThreads);
    return false;
  };

  // =====================================================================================================================
  // costmap pass (fused cycle): it needs the sweep's column counts, nothing of Mark.  The first kMarkCostmapWarps warps of the
  // CTA run it FIRST — while their own first units are in flight and the other warps stream — and then join the unit stream
  // (units are handed out dynamically in this variant, so late joiners simply take fewer), which keeps the pass off the tail
  // of the cycle: after the sweep, only the flush is left.
  // =====================================================================================================================
  if constexpr (kCostmap) {
    if (warp < (unsigned)kMarkCostmapWarps) {
      TRACE(3, 0);
      sync_with_predecessor();
      const uint32_t n_tiles = min(ld_volatile_u32(&g.ctr->tile_n), g.tile_capacity);
      TRACE(3, 1);
      const uint32_t cw = blockIdx.x * (uint32_t)kMarkCostmapWarps + warp, n_cw = gridDim.x * (uint32_t)kMarkCostmapWarps;
      unsigned cm_n = 0;
      int cm_mnx = INT_MAX, cm_mny = INT_MAX, cm_mxx = INT_MIN, cm_mxy = INT_MIN;
      // tiles per warp and round trip: as few as keep every warp of the pass busy (it is a chain of dependent round trips,
      // not a stream)
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
      named_bar_sync(2, kMarkCostmapWarps * 32);
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

  if (dynamic) {
    // (a warp without a first unit — a short first group, or more warps than a group has units — still takes part)
    if (meta_cur == kMetaEnd) { meta_cur = issue(next_unit(0), 0, base_cur); if (meta_cur != kMetaEnd) meta_nxt = issue(next_unit(0), 1, base_nxt); }
    else meta_nxt = issue(next_unit(0), 1, base_nxt);
  }
  // ---- the warp's stream of units ---------------------------------------------------------------------------------------------
  // the warp's it-th unit sits in buffer it & 1; every unit (an empty one too) completes one phase of its
  // buffer's barrier
  bool first = true;
  for (uint32_t it = 0;; ++it) {
    const uint32_t meta = meta_cur;
    if (meta == kMetaEnd) break;
    const uint32_t b = it & 1u;
    const unsigned in_unit = meta & 0xFFu;
    const uint32_t ri = (meta >> 8) & 15u;
    float px[kMarkPPT], py[kMarkPPT], pz[kMarkPPT];
    if constexpr (kVec4) {
      const uint32_t bar = wbar + b * 8u, parity = (it >> 1) & 1u;
      uint32_t ok;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
      } while (!ok);
      const uint32_t addr = laddr + b * (uint32_t)(kMarkChunk * 16);
#pragma unroll
      for (int k = 0; k < kMarkPPT; ++k) {   // (an empty or short unit leaves stale points in the buffer: masked below)
        float w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(px[k]), "=f"(py[k]), "=f"(pz[k]), "=f"(w) : "r"(addr + (uint32_t)k * 512u));
      }
    } else {
      const DevReading & par = batch.r[ri];
#pragma unroll
      for (int k = 0; k < kMarkPPT; ++k) {
        const unsigned idx = (unsigned)k * 32u + lane;
        const PointXYZ p = load_point_generic(par, base_cur + idx, idx < in_unit);
        px[k] = p.x; py[k] = p.y; pz[k] = p.z;
      }
    }
    // the buffer may be refilled: the warp holds its points in registers
    __syncwarp();
    meta_cur = meta_nxt; base_cur = base_nxt;
    meta_nxt = issue(next_unit(warp + (it + 2u) * (uint32_t)kWarps), b, base_nxt);
    if (first) { TRACE(0, 2); first = false; }
    if (in_unit) {
      const uint32_t par_addr = sbase + kSmPar + ri * 48u;
      float4 pa, pb;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pa.x), "=f"(pa.y), "=f"(pa.z), "=f"(pa.w) : "r"(par_addr));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(pb.x), "=f"(pb.y), "=f"(pb.z), "=f"(pb.w) : "r"(par_addr));
      if (meta & kMetaTf) {   // cloud still in the sensor frame: tf2::doTransform first (measurement_buffer.cpp:141-146)
        const DevReading & par = batch.r[ri];
#pragma unroll
        for (int k = 0; k < kMarkPPT; ++k) transform_point(par.tf, px[k], py[k], pz[k]);
      }
      const float range2_reject = pa.x, range2_accept = pa.y, near_accept = pa.z, oxf = pa.w, oyf = pb.x, ozf = pb.y, zlo = pb.z, zhi = pb.w;
      // ---- single-precision screen of the range and z tests, the four points interleaved ------------------------------------
      // per point two bits: FAST (bit k) = certainly passes the reference's range and z tests; OPEN (bit 4 + k) = neither that
      // nor certainly rejected (inside the margin of a threshold; not finite — counted whatever its z —; float distance
      // overflowed): decided below
      uint32_t fl = 0;
#pragma unroll
      for (int k = 0; k < kMarkPPT; ++k) {
        const float fx = px[k], fy = py[k], fz = pz[k];
        // a float evaluation of the squared distance whose error is far below the margins of the thresholds: above
        // range2_reject the exact test certainly rejects, inside [near_accept, range2_accept] it certainly accepts.  A NaN or
        // infinite coordinate makes d2f NaN or +inf: every comparison below is false then.
        const float ax = __fsub_rn(fx, oxf), ay = __fsub_rn(fy, oyf), az = __fsub_rn(fz, ozf);
        const float d2f = fmaf(az, az, fmaf(ay, ay, __fmul_rn(ax, ax)));
        const bool there = (unsigned)k * 32u + lane < in_unit;
        // PassThrough z window (reference measurement_buffer.cpp:165-175) on exactly equivalent float thresholds
        const bool fast = there & (fz >= zlo) & (fz <= zhi) & (d2f >= near_accept) & (d2f <= range2_accept);
        const bool rejected = (d2f < CUDART_INF_F) & ((fz < zlo) | (fz > zhi) | (d2f > range2_reject));
        const bool open = there & !fast & !rejected;
        fl |= (fast ? 1u << k : 0u) | (open ? 16u << k : 0u);
      }
      // whole warps of rejected points (floor, ceiling, beyond the obstacle range) stop here
      if (__any_sync(kFull, fl != 0)) {
        float4 pc;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+32];" : "=f"(pc.x), "=f"(pc.y), "=f"(pc.z), "=f"(pc.w) : "r"(par_addr));
        const int fx0 = __float_as_int(pc.x), fy0 = __float_as_int(pc.y), fz0 = __float_as_int(pc.z);
        const uint32_t ri28 = __float_as_uint(pc.w);
        uint32_t code[kMarkPPT];   // the point's voxel code if it is to be queued, else kNoCode
#pragma unroll
        for (int k = 0; k < kMarkPPT; ++k) {
          code[k] = kNoCode;
          // slots of 32 consecutive points without a candidate (the image rows are coherent) skip the quantisation
          if (!__any_sync(kFull, (fl & (0x11u << k)) != 0)) continue;
          const float fx = px[k], fy = py[k], fz = pz[k];
          bool fast = (fl & (1u << k)) != 0, open = (fl & (16u << k)) != 0;
          // points that are not finite (invalid depth pixels; UB upstream) are dropped and counted right here; everything else
          // the screen left open goes to the exact sequence
          if (__any_sync(kFull, open)) {
            if (open && !(fabsf(fx) <= 3.402823466e38f && fabsf(fy) <= 3.402823466e38f && fabsf(fz) <= 3.402823466e38f)) {
              drop_nonfinite += 1u;
              open = false;
            }
            if (!__any_sync(kFull, fast | open)) continue;
          }
          // ---- single-precision quantisation -------------------------------------------------------------------------------
          // .cpp:293-303 in float: negatives are shifted by one voxel (z tests y's sign — sic), scaled by 1/vs, truncated toward zero
          const float Z = fy < 0 ? __fsub_rn(fz, vs_f) : fz;
          uint32_t diff = 0;
          bool tiny = false;
          int ix = quantise_xy_fast(fx, c_lo, c_hi, diff, tiny);
          int iy = quantise_xy_fast(fy, c_lo, c_hi, diff, tiny);
          int iz = quantise_axis_fast(Z, c_lo, c_hi, diff);
          const bool decided = (diff == 0) & !tiny;
          bool take = fast & decided;
          // ---- the exact double sequence for the points the fast path could not decide (rare; whole warps skip it) ----------
          const bool slow = open | (fast & !decided);
          if (__any_sync(kFull, slow)) {
            if (slow) {
              const DevReading & par = batch.r[ri];
              const int4 q = quantise_exact(fx, fy, fz, par.ox, par.oy, par.oz, par.mark_range_2, g.vs, g.inv_vs, zlo, zhi);
              ix = q.x; iy = q.y; iz = q.z;
              take = q.w == 1;
              drop_range += q.w == 2 ? 1u : 0u;
              drop_nonfinite += q.w == 3 ? 1u : 0u;
            }
          }
          if (sharded) take = take && shard_owns(ix >> 3, g.shard_index, g.shard_count, g.shard_width);   // spatial shard: foreign leaves are dropped
          // ---- CTA-local dedupe: 32-bit voxel code, direct-mapped filter -----------------------------------------------------
          if (take) {
            const uint32_t rx = (uint32_t)(ix - fx0), ry = (uint32_t)(iy - fy0), rz = (uint32_t)(iz - fz0);
            const uint32_t c = ri28 | (rx << 18) | (ry << 8) | rz;
            if (((((rx | ry) >> 10) | (rz >> 8)) == 0u) & (c != kNoCode)) {
              const uint32_t fa = sbase + (((c * 0x9E3779B1u) >> kMarkFilterShift) << 2);
              uint32_t cur;
              asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(cur) : "r"(fa), "n"(kSmFilter));
              if (cur != c) {   // not seen (or forgotten): remember it and queue it.  The exchange lets exactly one of the
                uint32_t old;   // lanes (and warps) that find the same voxel missing at the same time do the queueing
                asm volatile("atom.shared.exch.b32 %0, [%1+%3], %2;" : "=r"(old) : "r"(fa), "r"(c), "n"(kSmFilter) : "memory");
                if (old != c) code[k] = c;
              }
            } else {   // farther from the sensor than the code reaches: straight to the global store
              sync_with_predecessor();
              mark_voxel_direct(g, ix >> 3, iy >> 3, iz >> 3, (unsigned)(((ix & 7) * 8 + (iy & 7)) * 8 + (iz & 7)), s_rnow[ri], use_max, batch.min_now);
            }
          }
        }
        // ---- warp-aggregated append of the unit's new voxels to the queue ----------------------------------------------------
        unsigned pm[kMarkPPT], n_push = 0;
#pragma unroll
        for (int k = 0; k < kMarkPPT; ++k) { pm[k] = __ballot_sync(kFull, code[k] != kNoCode); n_push += __popc(pm[k]); }
        if (n_push) {
          unsigned q0 = 0;
          if (lane == 0) {
            asm volatile("atom.shared.add.u32 %0, [%1+%3], %2;" : "=r"(q0) : "r"(sbase), "r"(n_push), "n"(kSmQn) : "memory");
            if (q0 < (unsigned)kMarkQueueFlush && q0 + n_push >= (unsigned)kMarkQueueFlush) *s_flush_req = 1u;
          }
          q0 = __shfl_sync(kFull, q0, 0);
          const unsigned below = (1u << lane) - 1u;
#pragma unroll
          for (int k = 0; k < kMarkPPT; ++k) {
            if (code[k] != kNoCode) {
              const unsigned q = q0 + __popc(pm[k] & below);
              if (q < (unsigned)kMarkQueue) asm volatile("st.shared.u32 [%0+%2], %1;" :: "r"(sbase + q * 4u), "r"(code[k]), "n"(kSmQueue) : "memory");
              else {   // (cannot happen while the flush threshold holds; kept for safety)
                sync_with_predecessor();
                mark_code_direct(g, code[k], s_par, s_rnow, use_max, batch.min_now);
              }
            }
            q0 += __popc(pm[k]);
          }
        }
      }
    }
    // a filling queue is flushed early (large or scattered clouds): every warp joins between two of its units
    if (*s_flush_req) flush();
  }
  if (lane == 0) atomicAdd(&s_done, 1u);
  while (!flush()) { }
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
  named_bar_sync(1, kThreads);
  if (threadIdx.x == 0) {
    if (s_drop[0]) red_add_u64(&g.ctr->dropped_range, (unsigned long long)s_drop[0]);   // rare: points beyond +-2^23 voxels
    if (s_drop[1]) red_add_u64(&g.ctr->dropped_nonfinite, (unsigned long long)s_drop[1]);
    TRACE(0, 6);
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
size_t mark_smem_bytes(bool vec4, int warps) { return (size_t)kSmWarp + (vec4 ? (size_t)warps * kSmWarpStride : 0); }
// warps of a Mark CTA of the fused cycle (STVL_MARK_WARPS_FUSED: tuning knob, 32 or 16.  With 16 — and one sweep CTA per SM,
// STVL_SWEEP_CTAS_PER_SM_FUSED=1 — a CTA of either kernel fits an SM at the same time and Mark's streaming really overlaps the
// sweep; measured on config 2 the two kernels then slow each other down by what the overlap saves, so 32 is the default)
int mark_warps_shared()
{
  static const int w = [] { const char * v = std::getenv("STVL_MARK_WARPS_FUSED"); return (v && std::atoi(v) == kMarkWarpsShared) ? kMarkWarpsShared : kMarkWarps; }();
  return w;
}
}  // namespace

// one CTA per SM; small batches get one CTA per unit
uint32_t mark_grid_for(const MarkBatch & batch, int sm_count) { return std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)sm_count, batch.total_blocks)); }
// ... and with the dynamic hand-out (fused cycle) one CTA per group of units
static uint32_t mark_groups(const MarkBatch & batch) { return (batch.total_blocks + (uint32_t)kMarkGroup - 1) / (uint32_t)kMarkGroup; }
static uint32_t mark_grid_dynamic(const MarkBatch & batch, int sm_count) { return std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)sm_count, mark_groups(batch))); }
// column tiles the fused variant's costmap pass can take without becoming the critical path of the launch
uint32_t mark_costmap_tiles_capacity(const MarkBatch & batch, int sm_count)
{
  return mark_grid_dynamic(batch, sm_count) * (uint32_t)kMarkCostmapWarps * 64u;
}

template <bool kVec4, bool kCostmap, bool kCmBits, int kWarps>
static cudaError_t launch_mark_variant(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                                       const CostmapDev & cm, uint8_t * costmap, uint32_t * cursor)
{
  static bool configured = false;
  const size_t smem = mark_smem_bytes(kVec4, kWarps);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mark_kernel<kVec4, kCostmap, kCmBits, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // All kernels of the cycle ask for the same (maximum) shared-memory carve-out: they overlap on the same SMs under
    // programmatic dependent launch, and an SM only changes its L1 / shared split when it is empty.
    cudaFuncSetAttribute(mark_kernel<kVec4, kCostmap, kCmBits, kWarps>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  // constants of the single-precision quantisation (quantise_axis_fast), rounded here exactly as the proof there assumes
  MarkBatch b = batch;
  b.vs_f = (float)g.vs;                                              // exact: vs is a float widened to double
  b.c_lo = (float)g.inv_vs * 0.999999701976776123046875f;            // (float)(1/vs) * (1 - 5 * 2^-24), one rounding
  b.c_hi = (float)g.inv_vs * 1.00000035762786865234375f;             // (float)(1/vs) * (1 + 3 * 2^-23), one rounding
  // Unit hand-out.  A launch of the fused cycle starts its CTAs one by one as the sweep leaves the SMs (and its costmap warps
  // join the stream late), so there the units are handed out dynamically, in groups, from the handle's cursor; a launch that has
  // the GPU to itself deals them round-robin: CTA c takes units c, c + grid, c + 2 grid, ...  (STVL_MARK_DYNAMIC=0/1 forces either for launches that do not carry the costmap pass.)
  static const int force_dynamic = std::getenv("STVL_MARK_DYNAMIC") ? std::atoi(std::getenv("STVL_MARK_DYNAMIC")) : -1;
  const bool dynamic = cursor != nullptr && g.mark_cursor != nullptr && (kCostmap || (force_dynamic < 0 ? pdl : force_dynamic != 0));
  const uint32_t grid = dynamic ? mark_grid_dynamic(batch, sm_count) : mark_grid_for(batch, sm_count);
  b.units_per_cta = b.total_blocks / grid;
  b.units_extra = b.total_blocks % grid;
  b.dynamic_units = dynamic ? 1u : 0u;
  b.n_groups = mark_groups(batch);
  b.cursor_base = cursor ? *cursor : 0u;
  for (uint32_t k = 0; k < b.n_readings; ++k) {
    b.src[k].data = b.r[k].data; b.src[k].n_points = b.r[k].n_points; b.src[k].n_points_dev = b.r[k].n_points_dev; b.src[k].has_tf = b.r[k].has_tf;
  }
  for (uint32_t k = 0; k < b.n_readings; ++k) {   // frame of the 32-bit voxel codes: the voxel the reading's origin falls into
    const double o[3] = {b.r[k].ox, b.r[k].oy, b.r[k].oz};
    int32_t ref[3];
    for (int a = 0; a < 3; ++a) {
      const double q = std::floor(o[a] / g.vs);
      ref[a] = (std::isfinite(q) && std::fabs(q) < 1e9) ? (int32_t)q : 0;
    }
    b.r[k].ref_ix = ref[0]; b.r[k].ref_iy = ref[1]; b.r[k].ref_iz = ref[2];
  }
  const cudaError_t e = launch_ex(mark_kernel<kVec4, kCostmap, kCmBits, kWarps>, grid, (unsigned)(kWarps * 32), smem, s, pdl, g, b, cm, costmap);
  // every CTA claims groups until it is told ONCE that there is none left: the launch moves the cursor by this much
  if (e == cudaSuccess && dynamic) *cursor += (b.n_groups > grid ? b.n_groups - grid : 0u) + grid;
  return e;
}

// pdl: the caller guarantees that the kernel's inputs (clouds) were complete before the PREDECESSOR of this launch in the
// stream started, and that the predecessor is this library's sweep kernel.  cm / costmap != NULL: fused costmap pass
// (cm_bits: into a cell bitmap).
template <bool kCostmap, bool kCmBits>
static cudaError_t launch_mark_layout(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                                      const CostmapDev & cm, uint8_t * costmap, bool all_vec4, bool shared_sm, uint32_t * cursor)
{
  if (shared_sm && mark_warps_shared() == kMarkWarpsShared)
    return all_vec4 ? launch_mark_variant<true, kCostmap, kCmBits, kMarkWarpsShared>(g, batch, sm_count, s, pdl, cm, costmap, cursor)
                    : launch_mark_variant<false, kCostmap, kCmBits, kMarkWarpsShared>(g, batch, sm_count, s, pdl, cm, costmap, cursor);
  return all_vec4 ? launch_mark_variant<true, kCostmap, kCmBits, kMarkWarps>(g, batch, sm_count, s, pdl, cm, costmap, cursor)
                  : launch_mark_variant<false, kCostmap, kCmBits, kMarkWarps>(g, batch, sm_count, s, pdl, cm, costmap, cursor);
}
cudaError_t launch_mark(const GridView & g, const MarkBatch & batch, int sm_count, cudaStream_t s, bool pdl,
                        const CostmapDev * cm, uint8_t * costmap, bool cm_bits, uint32_t * cursor)
{
  if (batch.total_blocks == 0) return cudaSuccess;
  bool all_vec4 = true;
  for (uint32_t i = 0; i < batch.n_readings; ++i) all_vec4 = all_vec4 && batch.r[i].vec4 != 0;
  const CostmapDev none{};
  // a launch of the fused cycle (chained behind the sweep, or carrying the costmap pass) shares the SMs with the sweep kernel
  const bool shared_sm = pdl || (cm && costmap);
  if (cm && costmap) {
    if (cm_bits) return launch_mark_layout<true, true>(g, batch, sm_count, s, pdl, *cm, costmap, all_vec4, shared_sm, cursor);
    return launch_mark_layout<true, false>(g, batch, sm_count, s, pdl, *cm, costmap, all_vec4, shared_sm, cursor);
  }
  return launch_mark_layout<false, false>(g, batch, sm_count, s, pdl, none, nullptr, all_vec4, shared_sm, cursor);
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
