// stvl_device.cuh — device-side data layout of the B200 sparse voxel grid.
//
// The reference keeps mark times in an openvdb::DoubleGrid (5-4-3 tree with 8^3
// leaves; reference spatio_temporal_voxel_grid.hpp:182, .cpp:80).  Here the store
// is a GPU hash of 8x8x8 LEAVES:
//
//   leaf hash   : open addressing, 64-bit packed leaf coordinate -> pool slot
//   leaf pool   : per slot  key (8 B) | value mask 512 bit (64 B) | 512 mark times f64 (4 KB)
//                 local index li = (lx*8 + ly)*8 + lz  (z fastest: a voxel column is one 64 B run)
//   column tiles: one per (bx,by) leaf column: 64 x u32 per-(x,y) voxel counts, refilled by
//                 every clear sweep (the reference's _cost_map, .cpp:227-251)
//
// Only mask words and the 32 B sectors of touched times ever move through HBM, a
// cleared voxel is one bit flip (no compaction pass), and frustum culling happens
// once per leaf instead of once per voxel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace stvl {

constexpr int kLeafDim = 8;
constexpr int kLeafVoxels = 512;
constexpr int kLeafBias = 1 << 20;              // leaf coords in [-2^20, 2^20)
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint64_t kTombKey = ~0ull - 1;
constexpr uint32_t kInvalidSlot = 0xFFFFFFFFu;  // hash value not yet published
constexpr uint32_t kOverflowSlot = 0xFFFFFFFEu; // pool exhausted when this leaf was requested
constexpr int kMaxFrustums = 512;
constexpr int kMaxBatchReadings = 16;

// -------- device-resident counters ---------------------------------------------------
// Results of one clear sweep.  Two sets: sweep k accumulates into set k&1 while its first kernel re-arms the other
// set for sweep k+1, so no separate "zero the counters" launch is needed.
struct SweepCtr {
  unsigned long long n_swept, n_survived, n_cleared, n_rewritten, n_ambiguous;
  int32_t cb_min_x, cb_min_y, cb_max_x, cb_max_y;
  unsigned long long n_points_out, n_cleared_out, n_ambiguous_out;
};
struct CostmapCtr {
  unsigned long long n_marked;
  int32_t mk_min_x, mk_min_y, mk_max_x, mk_max_y;
};
struct Counters {
  // leaf pool allocator, fold-free: a slot comes from the free ring (entries [free_pop, free_push) of free_slots, indices
  // taken modulo the ring size) and, when that is empty, from the bump pointer.  Mark / load kernels only pop and bump,
  // sweeps only push, and the two never run concurrently on a handle.
  uint32_t free_pop, free_push;   // monotone (wrapping) counters
  uint32_t bump;           // slots [0, min(bump, capacity)) have been handed out at least once
  uint32_t n_tombs;        // tombstones in the leaf hash
  uint32_t leaf_overflow;  // a Mark ran out of leaves (host grows the pool and replays)
  // tile pool
  uint32_t tile_n;
  uint32_t tile_overflow;
  uint32_t cm_ticket;      // last-CTA-done ticket of the costmap pass (its last CTA publishes cm_acc -> cm)
  uint32_t pad1[4];
  // dropped input points
  unsigned long long dropped_range, dropped_nonfinite;
  SweepCtr sw[2];
  CostmapCtr cm_acc;       // accumulators of the running costmap pass
  CostmapCtr cm;           // published result of the last costmap pass
  unsigned long long n_columns_out;
  // dumps
  unsigned long long n_voxels_out, n_active;
  uint32_t live_leaves;
  uint32_t pad0;
};

// A few allocator counters mirrored into host-mapped pinned memory by the kernels that change them (plain posted
// stores over PCIe, no stream operation): the host sizes its grids from them between synchronisations.  Estimates
// only — every kernel takes its real extents from Counters on the device and strides over them.
struct LiveCounters {
  uint32_t high_water, free_n, tile_n, seq;   // high_water = min(bump, capacity), free_n = free_push - free_pop
};

struct __align__(16) HashEntry {
  unsigned long long key;   // packed leaf coordinate, kEmptyKey or kTombKey
  uint32_t val;             // pool slot; kInvalidSlot until published
  uint32_t pad;
};

// -------- grid view handed to every kernel ----------------------------------------
struct GridView {
  // leaf hash: 16 B entries {key, slot} so that one 128-bit load resolves a probe
  HashEntry * hash;
  uint32_t hash_mask;      // capacity - 1 (power of two)
  // leaf pool
  uint64_t * leaf_key;     // kEmptyKey = slot is free
  uint32_t * leaf_hpos;    // position of the leaf in the hash (for tombstoning)
  uint32_t * leaf_tile;    // column tile of the leaf
  uint32_t * leaf_mask;    // 16 x u32 per leaf
  double * leaf_time;      // 512 per leaf
  double * leaf_tmin;      // lower bound of the mark times stored in the leaf (-inf = unknown): lets a sweep skip the
                           // 4 KB time block of leaves in which nothing can expire
  uint32_t * free_slots;   // free ring, free_ring_mask + 1 entries (a power of two >= leaf_capacity)
  uint32_t free_ring_mask;
  uint32_t leaf_capacity;
  // column tiles
  uint64_t * tile_hash_keys;
  uint32_t * tile_hash_vals;
  uint32_t tile_hash_mask;
  uint64_t * tile_key;
  uint32_t * tile_count;   // 64 per tile: the buffer the LAST sweep filled (what costmap / column queries read)
  uint32_t * tile_count_next;   // the other buffer: zeroed during a sweep for the one after it
  int sw_cur;              // which SweepCtr set / tile buffer the current (last) sweep uses
  uint32_t tile_capacity;
  Counters * ctr;
  uint32_t * mark_cursor;   // Mark's group hand-out: a monotone (wrapping) counter, never reset (every launch is told where it stands)
  LiveCounters * live;     // host-mapped mirror (may be NULL)
  // constants
  double vs, inv_vs, half_vs;
  double voxel_decay;
  int decay_model;
  // spatial shard (multi-GPU): keep leaf iff floor_mod(floor_div(bx, shard_width), shard_count) == shard_index
  int shard_index, shard_count, shard_width;
};

// -------- frustum as the sweep kernel sees it --------------------------------------
struct DevFrustum {
  int32_t type;          // 0 depth, 1 lidar, -1 never inside
  int32_t cull;          // 1: aabb / bounding data below are valid
  double c16f;           // (1./6.) * accel_factor, reference .cpp:338
  double p[36];          // stvl_frustum_t::p
  int32_t lo[3], hi[3];  // conservative voxel-index AABB (inclusive)
  double aux[4];         // lidar: min_d, max_d, tan(vFOV/2), |pad|
};

// up to kParamFrustums prepared frustum blocks travel as a kernel parameter (no upload, no stream operation between
// the kernels of consecutive cycles); larger sets go through a device buffer
constexpr int kParamFrustums = 8;
struct FrustumParams {
  DevFrustum f[kParamFrustums];
};

// -------- one marking reading on the device -----------------------------------------
struct DevReading {
  const uint8_t * data;
  unsigned long long n_points;                  // number of points (an upper bound when n_points_dev is set)
  const unsigned long long * n_points_dev;      // non-NULL: the actual count lives on the device (output of the VOXEL pre-filter)
  uint32_t point_step, off_x, off_y, off_z;
  double ox, oy, oz;
  double now;
  double zmin, zmax;
  float mark_range_2;
  int32_t filter;         // STVL_FILTER_PASSTHROUGH applies the z window in-kernel
  float zmin_f, zmax_f;   // float thresholds EXACTLY equivalent to the double window for float z: zmin rounded up, zmax rounded down
  float range2_reject;    // single-precision screen: a float distance^2 above this is certainly out of range ...
  float range2_accept;    // ... and one inside [near_accept, range2_accept] certainly passes the reference's exact test;
  float near_accept;      //     anything else takes the exact double path
  float oxf, oyf, ozf;    // (float) origin, for that screen
  float tf[12];           // sensor -> global float affine (row-major 3x4), applied first when has_tf
  uint32_t has_tf;
  uint32_t first_block;   // first unit of this reading in the batch's unit index space (units of 128 points)
  uint32_t vec4;          // 1: point_step==16, offsets 0/4/8, 16 B aligned -> float4 loads
  int32_t ref_ix, ref_iy, ref_iz;   // voxel of the reading's origin: frame of Mark's 32-bit voxel codes (set at launch)
};
// what Mark's unit hand-out needs of a reading (copied from DevReading at launch)
struct MarkStreamSource {
  const uint8_t * data;
  unsigned long long n_points;
  const unsigned long long * n_points_dev;
  uint32_t has_tf, pad;
};
struct MarkBatch {
  uint32_t first_block[kMaxBatchReadings];   // copy of r[k].first_block, one cache line: lane k of every warp keeps entry k
  MarkStreamSource src[kMaxBatchReadings];
  DevReading r[kMaxBatchReadings];
  uint32_t n_readings;
  uint32_t total_blocks;
  uint32_t use_atomic_max;  // 1: batched mode (monotone clocks), 0: plain stores
  int32_t ref_bx;           // reference leaf x of the batch (first reading's origin): block-local table keys store bx - ref_bx
  float vs_f;               // (float)vs
  float c_lo, c_hi;         // (float)(1/vs) scaled down / up by a few ulps: the two ends of the fast quantisation (set at launch)
  // unit hand-out (set at launch).  Static: units round-robin over the CTAs and, inside a CTA, over its warps.  Dynamic (fused
  // cycle, whose CTAs start at different times as the sweep leaves the SMs): groups of kMarkGroup units; CTA c starts with
  // group c, further groups come from GridView::mark_cursor, whose value when this launch starts is cursor_base
  uint32_t units_per_cta;   // total_blocks / grid and ...
  uint32_t units_extra;     // ... total_blocks % grid: the first units_extra CTAs take one unit more
  uint32_t dynamic_units;
  uint32_t cursor_base;
  uint32_t n_groups;        // ceil(total_blocks / kMarkGroup)
  uint32_t pad2;
  double min_now;           // smallest clock sample in the batch: initial leaf_tmin of leaves created by it
};

// -------- helpers -------------------------------------------------------------------
__host__ __device__ inline uint64_t pack_leaf_key(int bx, int by, int bz)
{
  return ((uint64_t)(uint32_t)(bx + kLeafBias) << 42) | ((uint64_t)(uint32_t)(by + kLeafBias) << 21) |
         (uint64_t)(uint32_t)(bz + kLeafBias);
}
__host__ __device__ inline void unpack_leaf_key(uint64_t k, int & bx, int & by, int & bz)
{
  bx = (int)((k >> 42) & 0x1FFFFF) - kLeafBias;
  by = (int)((k >> 21) & 0x1FFFFF) - kLeafBias;
  bz = (int)(k & 0x1FFFFF) - kLeafBias;
}
__host__ __device__ inline uint64_t pack_tile_key(int bx, int by)
{
  return ((uint64_t)(uint32_t)(bx + kLeafBias) << 21) | (uint64_t)(uint32_t)(by + kLeafBias);
}
__host__ __device__ inline void unpack_tile_key(uint64_t k, int & bx, int & by)
{
  bx = (int)((k >> 21) & 0x1FFFFF) - kLeafBias;
  by = (int)(k & 0x1FFFFF) - kLeafBias;
}
__host__ __device__ inline bool leaf_coord_ok(int b) { return b >= -kLeafBias && b < kLeafBias; }
__host__ __device__ inline uint64_t mix64(uint64_t x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}
__host__ __device__ inline int floor_div(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__host__ __device__ inline bool shard_owns(int bx, int shard_index, int shard_count, int shard_width)
{
  if (shard_count <= 1) return true;
  int s = floor_div(bx, shard_width) % shard_count;
  if (s < 0) s += shard_count;
  return s == shard_index;
}

#ifdef __CUDACC__
// tf2::doTransform of one point in float (reference measurement_buffer.cpp:141-146; tf2_sensor_msgs applies an
// Eigen::Transform<float,3,Affine>): out_i = ((m_i0*x + m_i1*y) + m_i2*z) + m_i3, no FMA
__device__ __forceinline__ void transform_point(const float * __restrict__ m, float & x, float & y, float & z)
{
  const float ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
  const float oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], x), __fmul_rn(m[5], y)), __fmul_rn(m[6], z)), m[7]);
  const float oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], x), __fmul_rn(m[9], y)), __fmul_rn(m[10], z)), m[11]);
  x = ox; y = oy; z = oz;
}

// IndexToWorld, reference .cpp:446-460: OpenVDB indexToWorld (coord*vs) THEN + vs/2 — two roundings, never an FMA.
__device__ __forceinline__ double index_to_world(int i, double vs, double half_vs)
{
  return __dadd_rn(__dmul_rn((double)i, vs), half_vs);
}

// DepthCameraFrustum::IsInside, reference depth_camera_frustum.cpp:286-304 (+ Dot :322-339).
// The reference normalises p_delta per plane (1 sqrt + 3 div).  sign(n . normalize(d)) can only differ from
// sign(n . d) inside a band of a few ulp around zero, so the un-normalised dot decides everywhere outside a
// 1e-12 relative guard band; inside the band the reference's exact operation sequence is replayed
// (correctly-rounded sqrt/div, no FMA), so the result is bit-identical to the CPU path everywhere.
__device__ __forceinline__ bool depth_is_inside(const double * __restrict__ pl, double x, double y, double z)
{
  // pass 1, branch-free: the six planes are independent chains (a lone warp on a dense leaf is latency bound, so
  // instruction-level parallelism matters more than early exit)
  bool outside = false;
  unsigned in_band = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double dx = __dsub_rn(x, pl[6 * k + 3]), dy = __dsub_rn(y, pl[6 * k + 4]), dz = __dsub_rn(z, pl[6 * k + 5]);
    const double tx = pl[6 * k + 0] * dx, ty = pl[6 * k + 1] * dy, tz = pl[6 * k + 2] * dz;
    const double s = tx + ty + tz;
    const double band = 1e-12 * (fabs(tx) + fabs(ty) + fabs(tz));
    outside = outside || (s > band);
    in_band |= (s >= -band && s <= band) ? (1u << k) : 0u;
  }
  if (outside) return false;   // some plane certainly has n . normalize(d) > 0
  // pass 2 (rare): planes whose dot is within the guard band replay the reference's exact operation sequence:
  // Eigen normalize() = divide by sqrt(squaredNorm) when squaredNorm > 0, squaredNorm = (x*x + y*y) + z*z;
  // Dot = (nx*dx + ny*dy) + nz*dz
  while (in_band) {
    const int k = __ffs(in_band) - 1;
    in_band &= in_band - 1;
    double dx = __dsub_rn(x, pl[6 * k + 3]), dy = __dsub_rn(y, pl[6 * k + 4]), dz = __dsub_rn(z, pl[6 * k + 5]);
    const double z2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (z2 > 0.0) {
      const double n = __dsqrt_rn(z2);
      dx = __ddiv_rn(dx, n); dy = __ddiv_rn(dy, n); dz = __ddiv_rn(dz, n);
    }
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(pl[6 * k + 0], dx), __dmul_rn(pl[6 * k + 1], dy)), __dmul_rn(pl[6 * k + 2], dz));
    if (dot > 0.) return false;
  }
  return true;
}

// Eigen QuaternionBase::_transformVector (q * v): uv = qv x v; uv += uv; v + w*uv + qv x uv — no FMA.
__device__ __forceinline__ void quat_rotate_exact(double qx, double qy, double qz, double qw,
                                                  double vx, double vy, double vz,
                                                  double & rx, double & ry, double & rz)
{
  double ux = __dsub_rn(__dmul_rn(qy, vz), __dmul_rn(qz, vy));
  double uy = __dsub_rn(__dmul_rn(qz, vx), __dmul_rn(qx, vz));
  double uz = __dsub_rn(__dmul_rn(qx, vy), __dmul_rn(qy, vx));
  ux = __dadd_rn(ux, ux); uy = __dadd_rn(uy, uy); uz = __dadd_rn(uz, uz);
  const double cx = __dsub_rn(__dmul_rn(qy, uz), __dmul_rn(qz, uy));
  const double cy = __dsub_rn(__dmul_rn(qz, ux), __dmul_rn(qx, uz));
  const double cz = __dsub_rn(__dmul_rn(qx, uy), __dmul_rn(qy, ux));
  rx = __dadd_rn(__dadd_rn(vx, __dmul_rn(qw, ux)), cx);
  ry = __dadd_rn(__dadd_rn(vy, __dmul_rn(qw, uy)), cy);
  rz = __dadd_rn(__dadd_rn(vz, __dmul_rn(qw, uz)), cz);
}

// ThreeDimensionalLidarFrustum::IsInside, reference three_dimensional_lidar_frustum.cpp:77-116.
// `ambiguous` is raised when the hFOV test hinged on atan() within 1e-12 rad (CUDA vs glibc libm may differ there).
__device__ __forceinline__ bool lidar_is_inside(const double * __restrict__ p, double x, double y, double z, bool & ambiguous)
{
  double tx, ty, tz;
  quat_rotate_exact(p[3], p[4], p[5], p[6], __dsub_rn(x, p[0]), __dsub_rn(y, p[1]), __dsub_rn(z, p[2]), tx, ty, tz);
  const double r2 = __dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty));
  if (r2 > p[10] || r2 < p[9]) return false;
  const double v_padded = __dadd_rn(fabs(tz), p[7]);
  if (__ddiv_rn(__dmul_rn(v_padded, v_padded), r2) > p[8]) return false;
  if (p[12] == 0.0) {
    const double half_pi = 1.5707963267948966;   // M_PI / 2
    if (tx > 0) {
      const double a = fabs(atan(__ddiv_rn(ty, tx)));
      if (fabs(a - p[11]) <= 1e-12) ambiguous = true;
      if (a > p[11]) return false;
    } else {
      const double a = __dadd_rn(fabs(atan(__ddiv_rn(tx, ty))), half_pi);
      if (fabs(a - p[11]) <= 1e-12) ambiguous = true;
      if (a > p[11]) return false;
    }
  }
  return true;
}
#endif  // __CUDACC__

}  // namespace stvl
