// stvl_api.cu — C-ABI of the B200 voxel grid (see include/stvl_b200.h): device memory, streams,
// allocator bookkeeping and the call sequencing around the kernels in stvl_kernels.cu.
//
// There is deliberately no CPU path in this file: every data-path entry point launches CUDA kernels and
// fails with STVL_ERR_NO_DEVICE / STVL_ERR_CUDA when that is impossible.
#include <algorithm>
#include <climits>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the symbols are resolved with dlopen at run time

#include "stvl_b200.h"
#include "stvl_device.cuh"
#include "stvl_geometry.h"
#include "stvl_kernels.h"
#include "stvl_voxel_filter.h"

using namespace stvl;

// the ctypes / cgo-style bindings rely on these layouts (tests/test_cabi_host.py)
static_assert(sizeof(stvl_config_t) == 56 && sizeof(stvl_reading_t) == 152 && sizeof(stvl_sweep_stats_t) == 88 &&
              sizeof(stvl_costmap_params_t) == 40 && sizeof(stvl_costmap_result_t) == 48 && sizeof(stvl_frustum_t) == 304 &&
              sizeof(stvl_pool_stats_t) == 88, "C-ABI struct layout changed");

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char * fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail(STVL_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

uint32_t pow2_ceil(uint64_t v)
{
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

template <typename T>
struct DevBuf {
  T * p = nullptr;
  size_t cap = 0;   // elements
};

}  // namespace

#ifdef STVL_TRACE
// trace build only: wall-clock split of the host side of the fused call (printed by stvl_destroy)
namespace { double g_host_us[4] = {0, 0, 0, 0}; unsigned long long g_host_calls = 0;
inline double host_now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); } }
#define HOST_T(var) const double var = host_now_us()
#define HOST_ADD(i, a, b) (g_host_us[i] += (b) - (a))
#else
#define HOST_T(var) do {} while (0)
#define HOST_ADD(i, a, b) do {} while (0)
#endif


struct stvl_grid {
  stvl_config_t cfg{};
  int device = 0;
  int sm_count = 148;
  int sm_spare = 0;     // SMs the persistent kernels of the running cycle leave free (sharded cycle: NCCL's copy kernels run beside them)
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copy_done = nullptr, ev_mark_done = nullptr;
  GridView v{};
  Counters * h_ctr = nullptr;       // pinned mirror
  LiveCounters * h_live = nullptr;  // host-mapped, written by the kernels (grid-sizing estimates between fetches)
  Counters last{};                  // counters at the last fetch
  // staging
  DevBuf<uint8_t> stage;
  DevBuf<DevFrustum> frustums;
  static constexpr int kFrustumRing = 4;
  DevFrustum * h_frustums[kFrustumRing] = {nullptr, nullptr, nullptr, nullptr};   // pinned, kMaxFrustums each
  cudaEvent_t ev_frustums[kFrustumRing] = {nullptr, nullptr, nullptr, nullptr};   // "copy out of slot i finished"
  int frustum_slot = 0;
  DevBuf<float> points;
  DevBuf<int32_t> cleared, ambiguous;
  DevBuf<uint8_t> costmap;
  // VOXEL pre-filter scratch (keys / point indices double-buffered for the radix sort, CUB temp, filtered clouds)
  DevBuf<uint8_t> vf_scratch;
  DevBuf<float4> vf_out;
  DevBuf<unsigned long long> vf_nout;   // per filtered reading: number of centroids (device counters read by Mark)
  size_t costmap_clean_bytes = 0;     // the internal costmap buffer is already filled with costmap_clean_value ...
  int costmap_clean_value = -1;       // ... over this many bytes (re-armed after every copy-out)
  DevBuf<int32_t> col_ix, col_iy;
  DevBuf<uint32_t> col_cnt;
  DevBuf<int32_t> dump_x, dump_y, dump_z;
  DevBuf<int32_t> poly_cols;        // per-column y range of a footprint polygon's outline (stvl_update_costs_device)
  DevBuf<double> dump_t;
  // host bookkeeping
  uint32_t out_flags = STVL_OUT_CLEARED | STVL_OUT_AMBIGUOUS;
  uint64_t active_upper = 0;        // upper bound of active voxels
  uint64_t added_since_sweep = 0;   // points marked / voxels loaded after the last enqueued sweep
  bool counters_fresh = true;       // `last` reflects the device state
  bool marks_unverified = false;    // a Mark ran since the allocator flags were last checked
  uint64_t pending_leaf_bound = 0;  // upper bound of leaves the unverified Marks can have allocated
  uint64_t pending_tile_bound = 0;  // same for column tiles
  std::vector<MarkBatch> replay;    // batches of the unverified Mark (device data still intact)
  double time_bound = 0.0;          // every stored mark time is <= this ...
  bool time_bound_inf = false;      // ... unless a sweep may have raised times
  bool swept = false;
  uint32_t * tile_buf[2] = {nullptr, nullptr};   // double-buffered column tiles (see GridView::tile_count)
  int sweep_parity = 0;
  uint64_t launches = 0;
  uint32_t mark_cursor_host = 0;   // value of *v.mark_cursor once every Mark launched so far has run
  bool record_mark_event = false;   // set by stvl_mark around its launches (ev_mark_done guards the host-cloud staging buffer)
  size_t device_bytes = 0;
  // in-library sharded merge (stvl_comm_init / stvl_costmap_sharded_device)
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_n = 1;
  cudaStream_t merge_stream = nullptr;
  cudaEvent_t ev_scatter[2] = {nullptr, nullptr}, ev_reduce[2] = {nullptr, nullptr};
  DevBuf<uint8_t> window[2];
  // bitmap merge of the sharded fused cycle: this rank's packed cell bitmap (double-buffered) and, on the root, the
  // staging area the other ranks' bitmaps are received into
  DevBuf<uint32_t> bits[2];
  DevBuf<uint32_t> bits_stage[2];
  int bit_w_lo[kMaxShardRanks] = {0}, bit_w_hi[kMaxShardRanks] = {0};   // word interval of every rank for the current geometry
  stvl_costmap_params_t bits_params{};   // geometry the intervals were computed for
  bool bits_geom_valid = false;
  int bits_pending = -1;                 // buffer whose exchange is in flight and not expanded yet
  int bits_expanding = -1;               // buffer whose expansion was put on the side stream by the current call
  cudaEvent_t ev_bits_done[2] = {nullptr, nullptr}, ev_expand[2] = {nullptr, nullptr}, ev_main_pos = nullptr;
  uint64_t bits_k = 0;
  uint64_t win_k = 0;
  int win_pending = -1;                 // window whose reduce is in flight and not thresholded yet
  int win_geom[4] = {0, 0, 0, 0};       // ix0, iy0, nx, ny of the pending window
  int win_count_bytes = 4;
  // pipelined host-buffer cycle (stvl_update_cycle_async): double-buffered staging of clouds / costmap bytes / results
  struct AsyncSlot {
    DevBuf<uint8_t> stage, costmap;
    CostmapCtr * d_res = nullptr;        // device snapshot of the cycle's costmap result
    CostmapCtr * h_res = nullptr;        // pinned
    cudaEvent_t ev_h2d = nullptr, ev_cycle = nullptr, ev_d2h = nullptr;
    bool busy = false;
    uint64_t ticket = 0;
  } aslot[2];
  cudaStream_t d2h_stream = nullptr;
  uint64_t next_ticket = 1;
  // optional per-stage timing (stvl_enable_timing)
  uint32_t timing_mask = 0;
  struct TimedInterval { cudaEvent_t a, b; int stage; };
  std::vector<TimedInterval> timing_pending;
  std::vector<cudaEvent_t> timing_pool;
  double vs = 0, inv_vs = 0, half_vs = 0;
};

namespace {

template <typename T>
int ensure(stvl_grid * g, DevBuf<T> & b, size_t n)
{
  if (n <= b.cap) return STVL_OK;
  size_t ncap = std::max<size_t>(n, b.cap * 2);
  ncap = std::max<size_t>(ncap, 1024);
  if (b.p) { CU(cudaStreamSynchronize(g->stream)); CU(cudaStreamSynchronize(g->copy_stream)); CU(cudaFree(b.p)); g->device_bytes -= b.cap * sizeof(T); b.p = nullptr; b.cap = 0; }
  CU(cudaMalloc(&b.p, ncap * sizeof(T)));
  b.cap = ncap;
  g->device_bytes += ncap * sizeof(T);
  return STVL_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define ENTER(g)                                                              \
  if (!(g)) return fail(STVL_ERR_INVALID, "null grid handle");               \
  DeviceGuard guard__((g)->device);                                          \
  if (!guard__.ok) return fail(STVL_ERR_CUDA, "cudaSetDevice(%d) failed", (g)->device)

// RAII bracket of one stage's launches with CUDA events (no-op unless the stage is selected)
struct StageTimer {
  stvl_grid * g;
  int stage;
  cudaEvent_t a = nullptr, b = nullptr;
  StageTimer(stvl_grid * g_, int stage_) : g(g_), stage(stage_)
  {
    if (!(g->timing_mask & (uint32_t)stage)) return;
    auto get = [&]() {
      cudaEvent_t e = nullptr;
      if (!g->timing_pool.empty()) { e = g->timing_pool.back(); g->timing_pool.pop_back(); }
      else if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
      return e;
    };
    a = get(); b = get();
    if (a && b) cudaEventRecord(a, g->stream);
  }
  ~StageTimer()
  {
    if (!a || !b) return;
    cudaEventRecord(b, g->stream);
    g->timing_pending.push_back({a, b, stage});
  }
};

int alloc_pool(stvl_grid * g, uint32_t leaf_capacity, GridView * nv)
{
  GridView v = g->v;
  v.leaf_capacity = leaf_capacity;
  const uint32_t hcap = pow2_ceil((uint64_t)leaf_capacity * 4);
  v.hash_mask = hcap - 1;
  CU(cudaMalloc(&v.hash, (size_t)hcap * sizeof(HashEntry)));
  CU(cudaMalloc(&v.leaf_key, (size_t)leaf_capacity * 8));
  CU(cudaMalloc(&v.leaf_hpos, (size_t)leaf_capacity * 4));
  CU(cudaMalloc(&v.leaf_tile, (size_t)leaf_capacity * 4));
  CU(cudaMalloc(&v.leaf_mask, (size_t)leaf_capacity * 64));
  CU(cudaMalloc(&v.leaf_time, (size_t)leaf_capacity * kLeafVoxels * 8));
  CU(cudaMalloc(&v.leaf_tmin, (size_t)leaf_capacity * 8));
  const uint32_t ring = pow2_ceil(leaf_capacity);
  v.free_ring_mask = ring - 1;
  CU(cudaMalloc(&v.free_slots, (size_t)ring * 4));
  *nv = v;
  return STVL_OK;
}
size_t pool_bytes(const GridView & v)
{
  return ((size_t)v.hash_mask + 1) * sizeof(HashEntry) + (size_t)v.leaf_capacity * (8 + 4 + 4 + 64 + kLeafVoxels * 8 + 8) + ((size_t)v.free_ring_mask + 1) * 4;
}
void free_pool(GridView & v)
{
  cudaFree(v.hash); cudaFree(v.leaf_key); cudaFree(v.leaf_hpos);
  cudaFree(v.leaf_tile); cudaFree(v.leaf_mask); cudaFree(v.leaf_time); cudaFree(v.leaf_tmin); cudaFree(v.free_slots);
  v.hash = nullptr; v.leaf_key = nullptr; v.leaf_hpos = nullptr;
  v.leaf_tile = nullptr; v.leaf_mask = nullptr; v.leaf_time = nullptr; v.leaf_tmin = nullptr; v.free_slots = nullptr;
}
int alloc_tiles(stvl_grid * g, uint32_t tile_capacity, GridView * nv)
{
  GridView v = *nv;
  (void)g;
  v.tile_capacity = tile_capacity;
  const uint32_t hcap = pow2_ceil((uint64_t)tile_capacity * 4);
  v.tile_hash_mask = hcap - 1;
  CU(cudaMalloc(&v.tile_hash_keys, (size_t)hcap * 8));
  CU(cudaMalloc(&v.tile_hash_vals, (size_t)hcap * 4));
  CU(cudaMalloc(&v.tile_key, (size_t)tile_capacity * 8));
  CU(cudaMalloc(&v.tile_count, (size_t)tile_capacity * 256 * 2));   // two buffers back to back
  v.tile_count_next = v.tile_count + (size_t)tile_capacity * 64;
  *nv = v;
  return STVL_OK;
}
size_t tile_bytes(const GridView & v) { return ((size_t)v.tile_hash_mask + 1) * 12 + (size_t)v.tile_capacity * (8 + 512); }
void free_tiles(GridView & v)
{
  cudaFree(v.tile_hash_keys); cudaFree(v.tile_hash_vals); cudaFree(v.tile_key);
  cudaFree(v.tile_count < v.tile_count_next ? v.tile_count : v.tile_count_next);
  v.tile_hash_keys = nullptr; v.tile_hash_vals = nullptr; v.tile_key = nullptr; v.tile_count = nullptr; v.tile_count_next = nullptr;
}

// sentinel state of the counters (bounding boxes start empty)
int init_counters(stvl_grid * g)
{
  Counters c;
  std::memset(&c, 0, sizeof(c));
  for (int k = 0; k < 2; ++k) { c.sw[k].cb_min_x = c.sw[k].cb_min_y = INT_MAX; c.sw[k].cb_max_x = c.sw[k].cb_max_y = INT_MIN; }
  c.cm_acc.mk_min_x = c.cm_acc.mk_min_y = INT_MAX; c.cm_acc.mk_max_x = c.cm_acc.mk_max_y = INT_MIN;
  c.cm = c.cm_acc;
  *g->h_ctr = c;
  CU(cudaMemcpyAsync(g->v.ctr, g->h_ctr, sizeof(Counters), cudaMemcpyHostToDevice, g->stream));
  CU(cudaStreamSynchronize(g->stream));
  g->last = c;
  return STVL_OK;
}
// point the view at the tile buffer / counter set of sweep parity p
void select_sweep_buffers(stvl_grid * g, int p)
{
  uint32_t * base = g->v.tile_count < g->v.tile_count_next ? g->v.tile_count : g->v.tile_count_next;
  uint32_t * other = base + (size_t)g->v.tile_capacity * 64;
  g->v.tile_count = p ? other : base;
  g->v.tile_count_next = p ? base : other;
  g->v.sw_cur = p;
  g->sweep_parity = p;
}

// wipe the whole store (ResetGrid / creation)
int clear_pool(stvl_grid * g)
{
  GridView & v = g->v;
  cudaStream_t s = g->stream;
  CU(cudaMemsetAsync(v.hash, 0xFF, ((size_t)v.hash_mask + 1) * sizeof(HashEntry), s));
  CU(cudaMemsetAsync(v.leaf_key, 0xFF, (size_t)v.leaf_capacity * 8, s));
  CU(cudaMemsetAsync(v.leaf_mask, 0, (size_t)v.leaf_capacity * 64, s));
  CU(cudaMemsetAsync(v.leaf_time, 0, (size_t)v.leaf_capacity * kLeafVoxels * 8, s));
  CU(cudaMemsetAsync(v.tile_hash_keys, 0xFF, ((size_t)v.tile_hash_mask + 1) * 8, s));
  CU(cudaMemsetAsync(v.tile_hash_vals, 0xFF, ((size_t)v.tile_hash_mask + 1) * 4, s));
  select_sweep_buffers(g, 0);
  CU(cudaMemsetAsync(v.tile_count, 0, (size_t)v.tile_capacity * 512, s));
  { int rc = init_counters(g); if (rc != STVL_OK) return rc; }
  g->counters_fresh = true;
  g->marks_unverified = false;
  g->pending_leaf_bound = g->pending_tile_bound = 0;
  g->replay.clear();
  g->active_upper = 0;
  g->added_since_sweep = 0;
  g->time_bound = 0.0;
  g->time_bound_inf = false;
  g->swept = false;
  return STVL_OK;
}

// leaves in use / slots handed out, from a counter snapshot
inline uint32_t ctr_high_water(const stvl_grid * g, const Counters & c) { return std::min<uint32_t>(c.bump, g->v.leaf_capacity); }
inline uint32_t ctr_free_n(const Counters & c) { return c.free_push - c.free_pop; }

// copy the device counters to the host mirror and wait for the stream
int fetch_counters(stvl_grid * g)
{
  CU(cudaMemcpyAsync(g->h_ctr, g->v.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, g->stream));
  CU(cudaStreamSynchronize(g->stream));
  g->last = *g->h_ctr;
  g->counters_fresh = true;
  g->pending_leaf_bound = g->pending_tile_bound = 0;
  // tighten the active-voxel bound: survivors of the last sweep + everything added after it
  if (g->swept) g->active_upper = g->last.sw[g->sweep_parity].n_survived + g->added_since_sweep;
  return STVL_OK;
}

int rebuild_leaf_hash(stvl_grid * g)
{
  GridView & v = g->v;
  CU(cudaMemsetAsync(v.hash, 0xFF, ((size_t)v.hash_mask + 1) * sizeof(HashEntry), g->stream));
  CU(launch_rebuild_hash(v, v.leaf_capacity, g->sm_count, g->stream));
  g->launches += 1;
  return STVL_OK;
}
// Rebuild the tile table from the live leaves (drops the tiles of emptied leaf columns, renumbers the rest).
// old_keys / old_counts (n_old tiles): the tile keys and the count rows of the LAST sweep as they were before the rebuild;
// they are carried over to the renumbered tiles, so the reference's _cost_map survives until the next sweep refills it.
int rebuild_tiles(stvl_grid * g, const uint64_t * old_keys = nullptr, const uint32_t * old_counts = nullptr, uint32_t n_old = 0)
{
  GridView & v = g->v;
  CU(cudaMemsetAsync(v.tile_hash_keys, 0xFF, ((size_t)v.tile_hash_mask + 1) * 8, g->stream));
  CU(cudaMemsetAsync(v.tile_hash_vals, 0xFF, ((size_t)v.tile_hash_mask + 1) * 4, g->stream));
  { const int p = g->sweep_parity; select_sweep_buffers(g, 0); CU(cudaMemsetAsync(v.tile_count, 0, (size_t)v.tile_capacity * 512, g->stream)); select_sweep_buffers(g, p); }
  CU(cudaMemsetAsync(&v.ctr->tile_n, 0, 4, g->stream));
  CU(cudaMemsetAsync(&v.ctr->tile_overflow, 0, 4, g->stream));
  CU(launch_reassign_tiles(v, v.leaf_capacity, g->sm_count, g->stream));
  g->launches += 1;
  if (old_keys && old_counts && n_old) {
    CU(launch_remap_tile_counts(v, old_keys, old_counts, n_old, g->sm_count, g->stream));
    g->launches += 1;
  }
  return STVL_OK;
}

// double the leaf pool (keeps every leaf in its slot), then rebuild the hash at the new size
int grow_leaves(stvl_grid * g)
{
  GridView old = g->v;
  if (old.leaf_capacity >= (1u << 28)) return fail(STVL_ERR_CAPACITY, "leaf pool cannot grow beyond 2^28 leaves");
  const uint32_t ncap = old.leaf_capacity * 2;
  GridView nv{};
  int rc = alloc_pool(g, ncap, &nv);
  if (rc != STVL_OK) return fail(STVL_ERR_CAPACITY, "growing the leaf pool to %u leaves failed: %s", ncap, g_err);
  cudaStream_t s = g->stream;
  const size_t oc = old.leaf_capacity, nc = ncap;
  CU(cudaMemcpyAsync(nv.leaf_key, old.leaf_key, oc * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemsetAsync(nv.leaf_key + oc, 0xFF, (nc - oc) * 8, s));
  CU(cudaMemcpyAsync(nv.leaf_tile, old.leaf_tile, oc * 4, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(nv.leaf_mask, old.leaf_mask, oc * 64, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemsetAsync(nv.leaf_mask + oc * 16, 0, (nc - oc) * 64, s));
  CU(cudaMemcpyAsync(nv.leaf_time, old.leaf_time, oc * kLeafVoxels * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemsetAsync(nv.leaf_time + oc * kLeafVoxels, 0, (nc - oc) * kLeafVoxels * 8, s));
  CU(cudaMemcpyAsync(nv.leaf_tmin, old.leaf_tmin, oc * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemsetAsync(&old.ctr->leaf_overflow, 0, 4, s));
  // the bump pointer ran past the old capacity while the pool was exhausted: clamp it, and rebuild the free ring (its size
  // changed) from the slots below it that hold no leaf
  {
    int rc2 = fetch_counters(g);
    if (rc2 != STVL_OK) return rc2;
    const uint32_t bump = std::min<uint32_t>(g->last.bump, old.leaf_capacity);
    g->h_ctr->bump = bump; g->h_ctr->free_pop = 0; g->h_ctr->free_push = 0;
    CU(cudaMemcpyAsync(&old.ctr->bump, &g->h_ctr->bump, 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(&old.ctr->free_pop, &g->h_ctr->free_pop, 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(&old.ctr->free_push, &g->h_ctr->free_push, 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    g->counters_fresh = false;
  }
  g->device_bytes += pool_bytes(nv);
  g->v = nv;
  CU(launch_rebuild_free_ring(g->v, g->sm_count, s));
  g->launches += 1;
  rc = rebuild_leaf_hash(g);
  if (rc != STVL_OK) return rc;
  CU(cudaStreamSynchronize(s));
  g->device_bytes -= pool_bytes(old);
  free_pool(old);
  return STVL_OK;
}
int grow_tiles(stvl_grid * g)
{
  const GridView old = g->v;
  const uint32_t n_old = std::min<uint32_t>(g->last.tile_n, old.tile_capacity);
  // The rebuild inserts a tile for every live leaf column AND for every old tile that still carries counts of the last
  // sweep; a doubled pool may still be too small for that (scattered clouds: about one tile per leaf), in which case count
  // rows would be dropped.  So: double until the rebuild fits, always carrying over from the ORIGINAL pool, which stays
  // intact until then.
  for (uint64_t cap = (uint64_t)old.tile_capacity * 2;; cap *= 2) {
    if (cap > (1ull << 28)) { g->v = old; return fail(STVL_ERR_CAPACITY, "column tile pool cannot grow beyond 2^28 tiles"); }
    GridView nv = old;
    int rc = alloc_tiles(g, (uint32_t)cap, &nv);
    if (rc != STVL_OK) { free_tiles(nv); g->v = old; return fail(STVL_ERR_CAPACITY, "growing the column tile pool failed: %s", g_err); }
    g->v = nv;
    select_sweep_buffers(g, g->sweep_parity);
    rc = rebuild_tiles(g, old.tile_key, old.tile_count, n_old);
    if (rc != STVL_OK) return rc;
    uint32_t overflow = 0;
    CU(cudaMemcpyAsync(&overflow, &g->v.ctr->tile_overflow, 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    if (!overflow) break;
    GridView failed = g->v;
    free_tiles(failed);
  }
  g->counters_fresh = false;
  g->device_bytes += tile_bytes(g->v);
  g->device_bytes -= tile_bytes(old);
  { GridView o = old; free_tiles(o); }
  return STVL_OK;
}

uint32_t tiles_upper(const stvl_grid * g);

// The fused update cycle (stvl_update_cycle_device): the sweep kernel resets the costmap bytes, the costmap pass rides on
// the Mark kernel, and the kernels are chained by programmatic dependent launch.
struct CycleFusion {
  CostmapDev cm{};
  uint8_t * costmap = nullptr;
  bool bits = false;           // the costmap pass writes this rank's packed cell bitmap (sharded cycle)
  bool pdl = false;            // the Mark launch directly follows this library's sweep kernel on the stream
  bool costmap_done = false;   // out: a launch carried the costmap pass
};

int launch_batches(stvl_grid * g, const std::vector<MarkBatch> & batches, CycleFusion * fz = nullptr)
{
  StageTimer timer__(g, STVL_TIME_MARK);
  bool pdl = fz && fz->pdl && !(g->timing_mask & STVL_TIME_MARK);
  if (fz && fz->costmap && !fz->costmap_done && !batches.empty() &&
      (uint64_t)tiles_upper(g) > (uint64_t)mark_costmap_tiles_capacity(batches.back(), g->sm_count - g->sm_spare)) {
    // Mark's grid is too small to carry the costmap pass (small clouds, large map): the costmap kernel goes FIRST — it
    // only needs the sweep — chained programmatically behind it, and Mark behind it, so that Mark's cloud streaming
    // overlaps it too (Mark's wait for this kernel implies the sweep: this kernel waited for it).
    CU(launch_costmap(g->v, fz->cm, fz->costmap, tiles_upper(g), g->sm_count, g->stream, false, pdl, fz->bits));
    fz->costmap_done = true;
    g->launches += 1;
  }
  for (size_t i = 0; i < batches.size(); ++i) {
    // the costmap pass rides on dedicated warps of every Mark CTA
    const bool carry = fz && fz->costmap && !fz->costmap_done && i + 1 == batches.size();
    const MarkBatch & b = batches[i];
    HOST_T(tm0);
    CU(launch_mark(g->v, b, g->sm_count - g->sm_spare, g->stream, pdl, carry ? &fz->cm : nullptr, carry ? fz->costmap : nullptr, carry && fz->bits,
                   &g->mark_cursor_host));
    HOST_T(tm1);
    HOST_ADD(3, tm0, tm1);
    if (carry) fz->costmap_done = true;
    pdl = false;                       // only the first launch follows the sweep
    g->launches += 1;
  }
  // (stvl_mark only: the staging buffer of host clouds is free again once these launches have run)
  if (g->record_mark_event) CU(cudaEventRecord(g->ev_mark_done, g->stream));
  return STVL_OK;
}

// Make sure the last Mark did not run out of leaves / tiles; if it did, grow and replay it (Mark is idempotent:
// same bits, same times).  Needs the counters, so it synchronises unless they are already fresh.
int verify_marks(stvl_grid * g)
{
  if (!g->marks_unverified) return STVL_OK;
  // No synchronisation needed when the pools provably could not overflow: leaves in use at the last fetch plus
  // everything the Marks since then can have allocated (each reading touches at most the leaves inside its
  // obstacle_range ball) still fits.  `last` is older than those Marks, frees since then only help.
  if (!g->counters_fresh) {
    const uint64_t leaves_known = (uint64_t)ctr_high_water(g, g->last) - ctr_free_n(g->last);
    if (!g->last.leaf_overflow && !g->last.tile_overflow &&
        leaves_known + g->pending_leaf_bound <= g->v.leaf_capacity &&
        (uint64_t)g->last.tile_n + g->pending_tile_bound <= g->v.tile_capacity) {
      g->marks_unverified = false;
      g->replay.clear();
      return STVL_OK;   // pending_*_bound stay: they are only reset by a fetch
    }
  }
  for (int attempt = 0; attempt < 32; ++attempt) {
    if (!g->counters_fresh) { int rc = fetch_counters(g); if (rc != STVL_OK) return rc; }
    if (!g->last.leaf_overflow && !g->last.tile_overflow) {
      g->marks_unverified = false;
      g->replay.clear();
      return STVL_OK;
    }
    if (g->last.leaf_overflow) { int rc = grow_leaves(g); if (rc != STVL_OK) return rc; }
    if (g->last.tile_overflow) { int rc = grow_tiles(g); if (rc != STVL_OK) return rc; }
    int rc = launch_batches(g, g->replay);
    if (rc != STVL_OK) return rc;
    g->counters_fresh = false;
  }
  return fail(STVL_ERR_CAPACITY, "leaf pool still exhausted after repeated growth");
}

// after any fetch: housekeeping decisions that only need the counters
int after_fetch(stvl_grid * g)
{
  int rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  return STVL_OK;
}

// upper bounds of the pool slots / tiles in use, for grid sizing (exact counters at the last fetch + what the Marks
// since then can have added)
// Used for grid sizing only (every kernel strides over the real extents it reads on the device): between fetches the
// proven bound only grows, so it is tightened with the kernels' own mirror of the counters plus slack for what the
// cycles still in flight can add.
uint32_t slots_upper(const stvl_grid * g)
{
  const uint64_t hw = ctr_high_water(g, g->last);
  uint64_t n = std::min<uint64_t>(g->v.leaf_capacity, hw + g->pending_leaf_bound + 1);
  if (g->h_live && !g->counters_fresh) {
    const uint64_t live = *reinterpret_cast<const volatile uint32_t *>(&g->h_live->high_water);
    if (live) n = std::min<uint64_t>(n, std::max<uint64_t>(live, hw) * 5 / 4 + 8192);
  }
  return (uint32_t)n;
}
uint32_t tiles_upper(const stvl_grid * g)
{
  uint64_t n = std::min<uint64_t>(g->v.tile_capacity, (uint64_t)g->last.tile_n + g->pending_tile_bound + 1);
  if (g->h_live && !g->counters_fresh) {
    const uint64_t live = *reinterpret_cast<const volatile uint32_t *>(&g->h_live->tile_n);
    if (live) n = std::min<uint64_t>(n, std::max<uint64_t>(live, g->last.tile_n) * 5 / 4 + 2048);
  }
  return (uint32_t)n;
}

int build_batches(stvl_grid * g, const stvl_reading_t * readings, uint32_t n, const std::vector<const uint8_t *> & dev_ptrs,
                  const std::vector<const unsigned long long *> & n_dev, std::vector<MarkBatch> * out, uint64_t * total_points, double * max_now, bool * monotone)
{
  MarkBatch cur{};
  double prev_now = -std::numeric_limits<double>::infinity();
  *monotone = true;
  *total_points = 0;
  *max_now = -std::numeric_limits<double>::infinity();
  for (uint32_t i = 0; i < n; ++i) {
    const stvl_reading_t & r = readings[i];
    if (!r.marking || r.n_points == 0) continue;   // `if (obs._marking)`, reference .cpp:273
    if (!(r.now >= prev_now)) *monotone = false;
    if (!(r.now > 0.0) || !std::isfinite(r.now)) *monotone = false;
    prev_now = r.now;
    *max_now = std::max(*max_now, r.now);
    if (cur.n_readings == kMaxBatchReadings) { out->push_back(cur); cur = MarkBatch{}; }
    cur.min_now = cur.n_readings ? std::min(cur.min_now, r.now) : r.now;
    if (cur.n_readings == 0) {   // reference leaf of the batch's block-local table keys
      const double lx = std::floor(r.origin[0] / (8.0 * g->vs));
      cur.ref_bx = (std::isfinite(lx) && std::fabs(lx) < 1e9) ? (int32_t)lx : 0;
    }
    DevReading & d = cur.r[cur.n_readings++];
    d.data = dev_ptrs[i];
    d.n_points = r.n_points;
    d.n_points_dev = n_dev[i];
    d.point_step = r.point_step; d.off_x = r.off_x; d.off_y = r.off_y; d.off_z = r.off_z;
    d.ox = r.origin[0]; d.oy = r.origin[1]; d.oz = r.origin[2];
    d.now = r.now;
    d.zmin = r.min_obstacle_height; d.zmax = r.max_obstacle_height;
    // (double)z < zmin  <=>  z < roundup_f32(zmin)   and   (double)z > zmax  <=>  z > rounddown_f32(zmax)   for float z
    d.zmin_f = (float)d.zmin; if ((double)d.zmin_f < d.zmin) d.zmin_f = std::nextafterf(d.zmin_f, INFINITY);
    d.zmax_f = (float)d.zmax; if ((double)d.zmax_f > d.zmax) d.zmax_f = std::nextafterf(d.zmax_f, -INFINITY);
    if (std::isnan(d.zmin)) d.zmin_f = -INFINITY;   // comparisons with NaN limits are false upstream: nothing is rejected
    if (std::isnan(d.zmax)) d.zmax_f = INFINITY;
    d.mark_range_2 = (float)(r.obstacle_range * r.obstacle_range);   // `float mark_range_2 = range * range`, .cpp:274
    // single-precision screen: a point whose float distance^2 exceeds range2_reject is certainly out of range, one inside
    // [near_accept, range2_accept] certainly passes; everything else takes the exact double test
    d.oxf = (float)d.ox; d.oyf = (float)d.oy; d.ozf = (float)d.oz;
    {
      const float m = std::fmax(std::fmax(std::fabs(d.oxf), std::fabs(d.oyf)), std::fmax(std::fabs(d.ozf), 1.0f));
      // The float evaluation d2f of the squared distance differs from the reference's float-narrowed double sum by at most
      // 2.1e-7 * d * (m + d) (the origin rounded to float, the three float differences) + 4.5e-7 * d^2 (squares, sums, the
      // reference's own narrowing); the margin is more than four times that at d^2 = range^2 and at d^2 = 1e-4
      const float margin = d.mark_range_2 * 2e-6f + 1e-6f * m * (std::sqrt(std::fmax(d.mark_range_2, 0.f)) + 1.0f) + 1e-7f;
      d.range2_reject = d.mark_range_2 + margin;
      if (!(d.range2_reject == d.range2_reject)) d.range2_reject = INFINITY;
      // inside [near_accept, range2_accept] the exact test (.cpp:289: !(d2 > range^2 || d2 < 0.0001)) certainly passes;
      // NaN / negative thresholds simply send every point to the exact path
      d.range2_accept = d.mark_range_2 - margin;
      d.near_accept = 0.0001f + margin;
      if (!(d.range2_accept == d.range2_accept) || !(d.near_accept == d.near_accept)) { d.range2_accept = -INFINITY; d.near_accept = INFINITY; }
    }
    d.filter = r.filter;
    d.has_tf = r.has_transform ? 1u : 0u;
    for (int k = 0; k < 12; ++k) d.tf[k] = r.transform[k];
    d.first_block = cur.total_blocks;
    cur.first_block[cur.n_readings - 1] = cur.total_blocks;
    d.vec4 = (r.point_step == 16 && r.off_x == 0 && r.off_y == 4 && r.off_z == 8 && ((uintptr_t)d.data & 15) == 0) ? 1u : 0u;
    cur.total_blocks += mark_blocks_for(r.n_points);
    *total_points += r.n_points;
    // leaves / tiles this reading can touch: points farther than obstacle_range from the origin are skipped
    // (reference .cpp:289), so at most the leaves intersecting that ball (+1 for the negative-coordinate shift)
    const double span = std::isfinite(r.obstacle_range) ? 2.0 * std::fabs(r.obstacle_range) / (8.0 * g->vs) + 3.0 : 1e18;
    const double lb = std::min((double)r.n_points, span * span * span), tb = std::min((double)r.n_points, span * span);
    g->pending_leaf_bound += lb < 1e18 ? (uint64_t)lb : (uint64_t)r.n_points;
    g->pending_tile_bound += tb < 1e18 ? (uint64_t)tb : (uint64_t)r.n_points;
  }
  if (cur.n_readings) out->push_back(cur);
  return STVL_OK;
}

int validate_readings(const stvl_reading_t * readings, uint32_t n)
{
  if (n && !readings) return fail(STVL_ERR_INVALID, "readings is NULL");
  for (uint32_t i = 0; i < n; ++i) {
    const stvl_reading_t & r = readings[i];
    if (!r.marking || r.n_points == 0) continue;
    if (!r.data) return fail(STVL_ERR_INVALID, "reading %u: data is NULL", i);
    if (r.point_step < 12 || r.off_x + 4 > r.point_step || r.off_y + 4 > r.point_step || r.off_z + 4 > r.point_step)
      return fail(STVL_ERR_INVALID, "reading %u: x/y/z offsets (%u,%u,%u) do not fit point_step %u", i, r.off_x, r.off_y, r.off_z, r.point_step);
    if (r.filter != STVL_FILTER_NONE && r.filter != STVL_FILTER_PASSTHROUGH && r.filter != STVL_FILTER_VOXEL) return fail(STVL_ERR_INVALID, "reading %u: unknown filter %d", i, r.filter);
    if (r.filter == STVL_FILTER_VOXEL && r.n_points > 0x7FFFFFFFull) return fail(STVL_ERR_INVALID, "reading %u: too many points for the VOXEL filter", i);
    if (r.n_points > (1ull << 40)) return fail(STVL_ERR_INVALID, "reading %u: n_points too large", i);
  }
  return STVL_OK;
}

// VOXEL pre-filter (reference measurement_buffer.cpp:153-164) for the readings that ask for it: each such reading is
// replaced by its filtered cloud — compact centroids, their count in a device counter (see stvl_voxel_filter.cu) — before
// Mark sees it.  n_dev[i] receives the counter's address (NULL for unfiltered readings).
int apply_voxel_filters(stvl_grid * g, std::vector<stvl_reading_t> & rs, std::vector<const uint8_t *> & dev_ptrs,
                        std::vector<const unsigned long long *> & n_dev)
{
  size_t total = 0, max_n = 0, n_filtered = 0;
  for (const stvl_reading_t & r : rs)
    if (r.marking && r.n_points && r.filter == STVL_FILTER_VOXEL) { total += (size_t)r.n_points; max_n = std::max<size_t>(max_n, (size_t)r.n_points); ++n_filtered; }
  if (total == 0) return STVL_OK;
  int rc;
  rc = ensure(g, g->vf_scratch, voxel_filter_scratch_bytes(max_n)); if (rc != STVL_OK) return rc;
  rc = ensure(g, g->vf_out, total); if (rc != STVL_OK) return rc;
  rc = ensure(g, g->vf_nout, n_filtered); if (rc != STVL_OK) return rc;
  size_t off = 0, k = 0;
  for (size_t i = 0; i < rs.size(); ++i) {
    stvl_reading_t & r = rs[i];
    if (!(r.marking && r.n_points && r.filter == STVL_FILTER_VOXEL)) continue;
    CU(launch_voxel_filter(dev_ptrs[i], r.n_points, r.point_step, r.off_x, r.off_y, r.off_z, g->cfg.voxel_size,
                           r.min_obstacle_height, r.max_obstacle_height, r.voxel_min_points, r.has_transform ? r.transform : nullptr,
                           g->vf_scratch.p, g->vf_out.p + off, g->vf_nout.p + k, g->sm_count, g->stream));
    g->launches += 4;
    dev_ptrs[i] = reinterpret_cast<const uint8_t *>(g->vf_out.p + off);
    n_dev[i] = g->vf_nout.p + k;
    r.point_step = 16; r.off_x = 0; r.off_y = 4; r.off_z = 8;
    r.filter = STVL_FILTER_NONE;   // the z window was applied by the filter
    r.has_transform = 0;           // ... and so was the sensor -> global transform
    off += (size_t)r.n_points;
    ++k;
  }
  return STVL_OK;
}

int run_mark(stvl_grid * g, const stvl_reading_t * readings_in, uint32_t n, const std::vector<const uint8_t *> & dev_ptrs_in,
             CycleFusion * fz = nullptr)
{
  std::vector<stvl_reading_t> rs(readings_in, readings_in + n);
  std::vector<const uint8_t *> dev_ptrs(dev_ptrs_in);
  std::vector<const unsigned long long *> n_dev(n, nullptr);
  {
    const uint64_t launches_before = g->launches;
    const int rc = apply_voxel_filters(g, rs, dev_ptrs, n_dev);
    if (rc != STVL_OK) return rc;
    if (fz && g->launches != launches_before) fz->pdl = false;   // filter kernels sit between the sweep and Mark
  }
  const stvl_reading_t * readings = rs.data();
  std::vector<MarkBatch> batches;
  uint64_t total_points = 0;
  double max_now = 0;
  bool monotone = true;
  int rc = build_batches(g, readings, n, dev_ptrs, n_dev, &batches, &total_points, &max_now, &monotone);
  if (rc != STVL_OK) return rc;
  if (batches.empty()) return STVL_OK;
  // batched launches resolve "later reading overwrites" by atomicMax, which is only the same thing when clocks are
  // monotone and no stored time exceeds them; otherwise one launch per reading with plain stores (exact overwrite).
  double first_now = batches[0].r[0].now;
  const bool use_max = monotone && !g->time_bound_inf && first_now >= g->time_bound;
  if (use_max) {
    for (MarkBatch & b : batches) b.use_atomic_max = 1;
  } else {
    std::vector<MarkBatch> seq;
    for (const MarkBatch & b : batches)
      for (uint32_t k = 0; k < b.n_readings; ++k) {
        MarkBatch one{};
        one.n_readings = 1;
        one.r[0] = b.r[k];
        one.r[0].first_block = 0;
        one.first_block[0] = 0;
        {
          const double lx = std::floor(b.r[k].ox / (8.0 * g->vs));
          one.ref_bx = (std::isfinite(lx) && std::fabs(lx) < 1e9) ? (int32_t)lx : 0;
        }
        one.total_blocks = mark_blocks_for(b.r[k].n_points);
        one.use_atomic_max = 0;
        one.min_now = b.r[k].now;
        seq.push_back(one);
      }
    batches.swap(seq);
  }
  rc = launch_batches(g, batches, fz);
  if (rc != STVL_OK) return rc;
  g->replay = batches;
  g->marks_unverified = true;
  g->counters_fresh = false;
  g->active_upper += total_points;
  g->added_since_sweep += total_points;
  if (!g->time_bound_inf) g->time_bound = std::max(g->time_bound, max_now);
  return STVL_OK;
}

void world_bbox(const stvl_grid * g, const int32_t idx[4], double out[4])
{
  for (int k = 0; k < 4; ++k) out[k] = (double)idx[k] * g->vs + g->half_vs;
}

CostmapDev to_dev(const stvl_costmap_params_t & p)
{
  CostmapDev d;
  d.origin_x = p.origin_x; d.origin_y = p.origin_y; d.resolution = p.resolution;
  d.inv_resolution = 1.0 / p.resolution;
  d.size_x = p.size_x; d.size_y = p.size_y; d.mark_threshold = p.mark_threshold;
  d.default_value = p.default_value; d.lethal = p.lethal_value;
  return d;
}

void fill_costmap_result(const stvl_grid * g, stvl_costmap_result_t * res)
{
  if (!res) return;
  std::memset(res, 0, sizeof(*res));
  res->n_marked_columns = g->last.cm.n_marked;
  res->has_bounds = g->last.cm.n_marked > 0;
  if (res->has_bounds) {
    const int32_t idx[4] = {g->last.cm.mk_min_x, g->last.cm.mk_min_y, g->last.cm.mk_max_x, g->last.cm.mk_max_y};
    world_bbox(g, idx, res->touched);
  }
}

}  // namespace

// ---- NCCL, loaded at run time -------------------------------------------------------------------------------------------
struct NcclApi {
  void * lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char * (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi * nccl_api()
{
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // prefer the copy that is already in the process (e.g. the one torch bundles), else the system library
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) {
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
      api.Reduce = reinterpret_cast<decltype(api.Reduce)>(dlsym(api.lib, "ncclReduce"));
      api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(api.lib, "ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(api.lib, "ncclRecv"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.lib, "ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.lib, "ncclGroupEnd"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Reduce && api.Send && api.Recv && api.GroupStart &&
               api.GroupEnd && api.GetErrorString;
    }
  }
  return api.ok ? &api : nullptr;
}
#define NC(expr)                                                                                        \
  do {                                                                                                  \
    ncclResult_t r__ = (expr);                                                                          \
    if (r__ != ncclSuccess) return fail(STVL_ERR_CUDA, "%s: %s", #expr, nccl_api()->GetErrorString(r__)); \
  } while (0)

// voxel-column window that covers every column whose centre can map into the costmap (same rule as sharding.dense_window)
void window_geometry(const stvl_grid * g, const stvl_costmap_params_t & p, int geom[4])
{
  geom[0] = (int)std::floor(p.origin_x / g->vs) - 2;
  geom[1] = (int)std::floor(p.origin_y / g->vs) - 2;
  geom[2] = (int)std::ceil(p.size_x * p.resolution / g->vs) + 5;
  geom[3] = (int)std::ceil(p.size_y * p.resolution / g->vs) + 5;
}

// root: threshold the reduced window `b` into costmap bytes once its reduce has finished
int threshold_window(stvl_grid * g, int b, const stvl_costmap_params_t & p, uint8_t * costmap_dev, int root)
{
  CU(cudaStreamWaitEvent(g->stream, g->ev_reduce[b], 0));
  if (g->win_count_bytes == 0) {
    // bytes mode: the reduce left the merged {0, lethal} bytes in the root's window; copy them out, and give unknown-space
    // layers (default_value != 0) their default where nothing was marked
    if (g->comm_rank == root && costmap_dev) {
      const size_t bytes = (size_t)p.size_x * p.size_y;
      CU(cudaMemcpyAsync(costmap_dev, g->window[b].p, bytes, cudaMemcpyDeviceToDevice, g->stream));
      if (p.default_value != 0) {
        CU(launch_fill_where_zero(costmap_dev, bytes, p.default_value, g->sm_count, g->stream));
        g->launches += 1;
      }
    }
    return STVL_OK;
  }
  if (g->comm_rank == root && costmap_dev) {
    StageTimer timer__(g, STVL_TIME_COSTMAP);
    CU(launch_costmap_from_window(g->v, g->window[b].p, g->win_count_bytes, g->win_geom[0], g->win_geom[1], (uint32_t)g->win_geom[2],
                                  (uint32_t)g->win_geom[3], to_dev(p), costmap_dev, g->sm_count, g->stream));
    g->launches += 1;
    g->counters_fresh = false;
  }
  return STVL_OK;
}

// ---- bitmap merge of the sharded fused cycle ------------------------------------------------------------------------------
// 32-cell word interval [w_lo, w_hi) of a costmap row that rank r's x-slabs can mark (hull over its slabs), for every rank.
// Pure host arithmetic (exposed as stvl_shard_word_intervals).
void compute_word_intervals(double vs, int shard_count, int shard_width, int n_ranks, int own_rank, const stvl_costmap_params_t & p,
                            int * w_lo, int * w_hi)
{
  const int n = n_ranks, W = std::max(shard_width, 1);
  const double half_vs = vs / 2.0;
  const int wpr = (int)((p.size_x + 31u) >> 5);
  // leaf-slab range that can touch the costmap's x extent
  const double x_lo = p.origin_x - 2.0 * vs, x_hi = p.origin_x + (double)p.size_x * p.resolution + 2.0 * vs;
  const long long sl_lo = (long long)std::floor(x_lo / (8.0 * vs * W)) - 1, sl_hi = (long long)std::floor(x_hi / (8.0 * vs * W)) + 1;
  for (int r = 0; r < n; ++r) { w_lo[r] = wpr; w_hi[r] = 0; }
  for (long long sl = sl_lo; sl <= sl_hi; ++sl) {
    int r = (int)(sl % n); if (r < 0) r += n;
    if (shard_count <= 1) r = own_rank;
    // voxel columns of the slab: ix in [sl*W*8, (sl+1)*W*8); centres at ix*vs + vs/2
    const double cx0 = (double)(sl * W * 8) * vs + half_vs, cx1 = (double)((sl + 1) * W * 8 - 1) * vs + half_vs;
    long long m0 = (long long)std::floor((cx0 - p.origin_x) / p.resolution) - 1, m1 = (long long)std::floor((cx1 - p.origin_x) / p.resolution) + 1;
    if (m1 < 0 || m0 >= (long long)p.size_x) continue;
    m0 = std::max<long long>(m0, 0); m1 = std::min<long long>(m1, (long long)p.size_x - 1);
    w_lo[r] = std::min<int>(w_lo[r], (int)(m0 >> 5));
    w_hi[r] = std::max<int>(w_hi[r], (int)(m1 >> 5) + 1);
  }
  for (int r = 0; r < n; ++r) if (w_hi[r] <= w_lo[r]) { w_lo[r] = 0; w_hi[r] = 0; }   // nothing of rank r maps here
}
void shard_word_intervals(stvl_grid * g, const stvl_costmap_params_t & p)
{
  compute_word_intervals(g->vs, g->cfg.shard_count, g->cfg.shard_width, g->comm_n, g->comm_rank, p, g->bit_w_lo, g->bit_w_hi);
}

// every non-root rank sends its bitmap `b`, the root receives them back to back into its staging area (one NCCL group)
int exchange_bits(stvl_grid * g, int b, uint32_t size_y, int root, cudaStream_t xs)
{
  NcclApi * n = nccl_api();
  if (g->comm_n == 1) return STVL_OK;
  if (g->comm_rank != root) {
    const size_t words = (size_t)(g->bit_w_hi[g->comm_rank] - g->bit_w_lo[g->comm_rank]) * size_y;
    if (words) NC(n->Send(g->bits[b].p, words, ncclUint32, root, g->comm, xs));
    return STVL_OK;
  }
  NC(n->GroupStart());
  size_t off = 0;
  for (int r = 0; r < g->comm_n; ++r) {
    if (r == root) continue;
    const size_t words = (size_t)(g->bit_w_hi[r] - g->bit_w_lo[r]) * size_y;
    if (words) NC(n->Recv(g->bits_stage[b].p + off, words, ncclUint32, r, g->comm, xs));
    off += words;
  }
  NC(n->GroupEnd());
  return STVL_OK;
}

int launch_expand(stvl_grid * g, int b, const stvl_costmap_params_t & p, uint8_t * costmap_dev, int root, cudaStream_t xs)
{
  ExpandParams e{};
  e.n_ranks = g->comm_n;
  size_t off = 0;
  for (int r = 0; r < g->comm_n; ++r) {
    e.w_lo[r] = g->bit_w_lo[r]; e.w_hi[r] = g->bit_w_hi[r];
    if (r == root) e.bits[r] = g->bits[b].p;
    else { e.bits[r] = g->bits_stage[b].p + off; off += (size_t)(g->bit_w_hi[r] - g->bit_w_lo[r]) * p.size_y; }
  }
  CU(launch_expand_bitmaps(e, to_dev(p), costmap_dev, g->sm_count, xs));
  g->launches += 1;
  return STVL_OK;
}

// root: expand the pending bitmaps into costmap bytes on the MAIN stream (blocking mode, flush)
int expand_pending_bits(stvl_grid * g, const stvl_costmap_params_t & p, uint8_t * costmap_dev, int root)
{
  if (g->bits_pending < 0) return STVL_OK;
  const int b = g->bits_pending;
  g->bits_pending = -1;
  CU(cudaStreamWaitEvent(g->stream, g->ev_bits_done[b], 0));
  if (g->comm_rank == root && costmap_dev) return launch_expand(g, b, p, costmap_dev, root, g->stream);
  return STVL_OK;
}

// root: expand the pending bitmaps on the SIDE stream, not before everything enqueued on the main stream so far is done;
// the caller makes the main stream wait for ev_expand at the end of its own enqueue
int expand_pending_bits_async(stvl_grid * g, const stvl_costmap_params_t & p, uint8_t * costmap_dev, int root)
{
  const int b = g->bits_pending;
  g->bits_pending = -1;
  g->bits_expanding = -1;
  if (b < 0 || g->comm_rank != root || !costmap_dev) return STVL_OK;
  CU(cudaEventRecord(g->ev_main_pos, g->stream));
  CU(cudaStreamWaitEvent(g->merge_stream, g->ev_main_pos, 0));
  // (the exchange of buffer b was enqueued on the side stream: stream order covers it)
  int rc = launch_expand(g, b, p, costmap_dev, root, g->merge_stream);
  if (rc != STVL_OK) return rc;
  CU(cudaEventRecord(g->ev_expand[b], g->merge_stream));
  g->bits_expanding = b;
  return STVL_OK;
}

// =====================================================================================
extern "C" {

int stvl_abi_version(void) { return STVL_ABI_VERSION; }
const char * stvl_last_error(void) { return g_err; }

int stvl_device_count(int * count)
{
  if (!count) return fail(STVL_ERR_INVALID, "count is NULL");
  *count = 0;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return STVL_OK; }
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++*count;
  }
  return STVL_OK;
}

int stvl_create(const stvl_config_t * cfg, stvl_grid_t ** out)
{
  if (!cfg || !out) return fail(STVL_ERR_INVALID, "cfg/out is NULL");
  *out = nullptr;
  if (!(cfg->voxel_size > 0.f) || !std::isfinite(cfg->voxel_size)) return fail(STVL_ERR_INVALID, "voxel_size must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(STVL_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
  }
  int dev = cfg->device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return fail(STVL_ERR_CUDA, "cudaGetDevice failed"); }
  if (dev >= ndev) return fail(STVL_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev);
  int major = 0, minor = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10) return fail(STVL_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", dev, major, minor);

  stvl_grid * g = new (std::nothrow) stvl_grid();
  if (!g) return fail(STVL_ERR_INVALID, "out of host memory");
  g->cfg = *cfg;
  g->device = dev;
  g->sm_count = sms > 0 ? sms : 148;
  DeviceGuard guard(dev);
  if (!guard.ok) { delete g; return fail(STVL_ERR_CUDA, "cudaSetDevice(%d) failed", dev); }
  // `_voxel_size(voxel_size)`: float widened to double (reference .cpp:51,54); OpenVDB caches 1/scale
  g->vs = (double)cfg->voxel_size;
  g->inv_vs = 1.0 / g->vs;
  g->half_vs = g->vs / 2.0;

  auto bail = [&](int rc) { stvl_destroy(g); return rc; };
#define CUB(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return bail(fail(STVL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__))); } while (0)
  CUB(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
  CUB(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
  CUB(cudaEventCreateWithFlags(&g->ev_copy_done, cudaEventDisableTiming));
  CUB(cudaEventCreateWithFlags(&g->ev_mark_done, cudaEventDisableTiming));
  CUB(cudaMallocHost(&g->h_ctr, sizeof(Counters)));
  if (cudaHostAlloc(&g->h_live, sizeof(LiveCounters), cudaHostAllocMapped) == cudaSuccess) {
    std::memset(g->h_live, 0, sizeof(LiveCounters));
    LiveCounters * dev_live = nullptr;
    if (cudaHostGetDevicePointer(&dev_live, g->h_live, 0) == cudaSuccess) g->v.live = dev_live;
  } else {
    g->h_live = nullptr;
    cudaGetLastError();
  }
  for (int i = 0; i < stvl_grid::kFrustumRing; ++i) {
    CUB(cudaMallocHost(&g->h_frustums[i], sizeof(DevFrustum) * kMaxFrustums));
    CUB(cudaEventCreateWithFlags(&g->ev_frustums[i], cudaEventDisableTiming));
  }
  CUB(cudaMalloc(&g->v.ctr, sizeof(Counters)));
  g->device_bytes += sizeof(Counters);
  CUB(cudaMalloc(&g->v.mark_cursor, sizeof(uint32_t)));
  CUB(cudaMemset(g->v.mark_cursor, 0, sizeof(uint32_t)));

  const uint32_t leaf_cap = cfg->leaf_capacity ? std::max<uint32_t>(cfg->leaf_capacity, 64u) : (1u << 18);
  GridView nv = g->v;
  nv.vs = g->vs; nv.inv_vs = g->inv_vs; nv.half_vs = g->half_vs;
  nv.voxel_decay = cfg->voxel_decay; nv.decay_model = cfg->decay_model;
  nv.shard_index = cfg->shard_index; nv.shard_count = cfg->shard_count; nv.shard_width = std::max(cfg->shard_width, 1);
  g->v = nv;
  int rc = alloc_pool(g, leaf_cap, &nv);
  if (rc != STVL_OK) { free_pool(nv); return bail(rc); }
  g->v = nv;
  rc = alloc_tiles(g, std::max<uint32_t>(leaf_cap / 2, 64u), &nv);
  if (rc != STVL_OK) { g->v = nv; return bail(rc); }
  g->v = nv;
  g->device_bytes += pool_bytes(g->v) + tile_bytes(g->v);
  rc = clear_pool(g);
  if (rc != STVL_OK) return bail(rc);
  if (cfg->max_points) { rc = ensure(g, g->stage, (size_t)cfg->max_points * 16); if (rc != STVL_OK) return bail(rc); }
  CUB(cudaStreamSynchronize(g->stream));
#undef CUB
  *out = g;
  return STVL_OK;
}

int stvl_destroy(stvl_grid_t * g)
{
#ifdef STVL_TRACE
  if (g_host_calls) std::fprintf(stderr, "[stvl trace] host side of %llu fused cycles: clear %.2f us (of which the two sweep launches %.2f), mark %.2f us (of which the Mark launch %.2f)\n",
                                 g_host_calls, g_host_us[0] / g_host_calls, g_host_us[2] / g_host_calls, g_host_us[1] / g_host_calls, g_host_us[3] / g_host_calls);
#endif
  if (!g) return STVL_OK;
  DeviceGuard guard(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->copy_stream) cudaStreamSynchronize(g->copy_stream);
  if (g->comm) stvl_comm_destroy(g);
  cudaFree(g->window[0].p); cudaFree(g->window[1].p);
  cudaFree(g->bits[0].p); cudaFree(g->bits[1].p); cudaFree(g->bits_stage[0].p); cudaFree(g->bits_stage[1].p);
  free_pool(g->v);
  free_tiles(g->v);
  cudaFree(g->v.ctr);
  cudaFree(g->v.mark_cursor);
  cudaFree(g->stage.p); cudaFree(g->frustums.p); cudaFree(g->points.p); cudaFree(g->cleared.p); cudaFree(g->ambiguous.p);
  cudaFree(g->costmap.p); cudaFree(g->col_ix.p); cudaFree(g->col_iy.p); cudaFree(g->col_cnt.p);
  cudaFree(g->dump_x.p); cudaFree(g->dump_y.p); cudaFree(g->dump_z.p); cudaFree(g->dump_t.p);
  cudaFree(g->vf_scratch.p); cudaFree(g->vf_out.p); cudaFree(g->vf_nout.p);
  if (g->h_ctr) cudaFreeHost(g->h_ctr);
  if (g->h_live) cudaFreeHost(g->h_live);
  for (int i = 0; i < stvl_grid::kFrustumRing; ++i) {
    if (g->h_frustums[i]) cudaFreeHost(g->h_frustums[i]);
    if (g->ev_frustums[i]) cudaEventDestroy(g->ev_frustums[i]);
  }
  for (auto & iv : g->timing_pending) { cudaEventDestroy(iv.a); cudaEventDestroy(iv.b); }
  for (auto e : g->timing_pool) cudaEventDestroy(e);
  for (auto & a : g->aslot) {
    cudaFree(a.stage.p); cudaFree(a.costmap.p); cudaFree(a.d_res);
    if (a.h_res) cudaFreeHost(a.h_res);
    if (a.ev_h2d) cudaEventDestroy(a.ev_h2d);
    if (a.ev_cycle) cudaEventDestroy(a.ev_cycle);
    if (a.ev_d2h) cudaEventDestroy(a.ev_d2h);
  }
  if (g->d2h_stream) cudaStreamDestroy(g->d2h_stream);
  if (g->ev_copy_done) cudaEventDestroy(g->ev_copy_done);
  if (g->ev_mark_done) cudaEventDestroy(g->ev_mark_done);
  if (g->stream) cudaStreamDestroy(g->stream);
  if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
  cudaGetLastError();
  delete g;
  return STVL_OK;
}

int stvl_build_depth_frustum(double vfov, double hfov, double min_dist, double max_dist, const double origin[3],
                             const double quat_xyzw[4], double accel_factor, stvl_frustum_t * out)
{
  if (!origin || !quat_xyzw || !out) return fail(STVL_ERR_INVALID, "NULL argument");
  return build_depth_frustum(vfov, hfov, min_dist, max_dist, origin, quat_xyzw, accel_factor, out);
}
int stvl_build_lidar_frustum(double vfov, double vfov_padding, double hfov, double min_dist, double max_dist,
                             const double origin[3], const double quat_xyzw[4], double accel_factor, stvl_frustum_t * out)
{
  if (!origin || !quat_xyzw || !out) return fail(STVL_ERR_INVALID, "NULL argument");
  return build_lidar_frustum(vfov, vfov_padding, hfov, min_dist, max_dist, origin, quat_xyzw, accel_factor, out);
}

int stvl_build_transform(const double translation[3], const double quat_xyzw[4], float m[12])
{
  if (!translation || !quat_xyzw || !m) return fail(STVL_ERR_INVALID, "NULL argument");
  // Eigen::Quaternionf::toRotationMatrix, in float, products before sums (host code: built with -ffp-contract=off)
  const float qx = (float)quat_xyzw[0], qy = (float)quat_xyzw[1], qz = (float)quat_xyzw[2], qw = (float)quat_xyzw[3];
  const float dx = 2.0f * qx, dy = 2.0f * qy, dz = 2.0f * qz;
  const float wx = dx * qw, wy = dy * qw, wz = dz * qw;
  const float xx = dx * qx, xy = dy * qx, xz = dz * qx;
  const float yy = dy * qy, yz = dz * qy, zz = dz * qz;
  const float rot[9] = {1.0f - (yy + zz), xy - wz, xz + wy,
                        xy + wz, 1.0f - (xx + zz), yz - wx,
                        xz - wy, yz + wx, 1.0f - (xx + yy)};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) m[4 * r + c] = rot[3 * r + c];
    m[4 * r + 3] = (float)translation[r];
  }
  return STVL_OK;
}

int stvl_set_outputs(stvl_grid_t * g, uint32_t flags)
{
  if (!g) return fail(STVL_ERR_INVALID, "null grid handle");
  g->out_flags = flags;
  return STVL_OK;
}

namespace {
int clear_frustums_impl(stvl_grid * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums, CycleFusion * fz)
{
  if (n_frustums && !frustums) return fail(STVL_ERR_INVALID, "frustums is NULL");
  if (n_frustums > (uint32_t)kMaxFrustums) return fail(STVL_ERR_INVALID, "at most %d clearing frustums per sweep", kMaxFrustums);
  // the allocator flags of the previous Mark must be known before its voxels are swept
  int rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  // hash / tile maintenance, decided from the last known counters
  if (g->last.n_tombs > (g->v.hash_mask + 1) / 8) { rc = rebuild_leaf_hash(g); if (rc != STVL_OK) return rc; g->last.n_tombs = 0; }
  if (g->last.tile_n > g->v.tile_capacity / 4 * 3) {
    rc = rebuild_tiles(g); if (rc != STVL_OK) return rc;
    rc = fetch_counters(g); if (rc != STVL_OK) return rc;
    if (g->last.tile_n > g->v.tile_capacity / 2) { rc = grow_tiles(g); if (rc != STVL_OK) return rc; }
  }
  // frustum blocks -> device (the previous sweep that read d_frustums is ordered before on the stream)
  bool may_raise = false;
  FrustumParams fparam;
  const DevFrustum * frustums_dev = nullptr;
  if (n_frustums && n_frustums <= (uint32_t)kParamFrustums) {
    // small sets (the usual handful of sensors) travel as kernel parameters: no upload, no stream operation
    for (uint32_t i = 0; i < n_frustums; ++i) {
      prepare_frustum(frustums[i], g->vs, &fparam.f[i]);
      if (!(frustums[i].accel_factor >= 0.0)) may_raise = true;
    }
  } else if (n_frustums) {
    rc = ensure(g, g->frustums, n_frustums);
    if (rc != STVL_OK) return rc;
    // ring of pinned staging slots: a slot is reused only after the copy that read it has finished
    const int slot = g->frustum_slot;
    g->frustum_slot = (slot + 1) % stvl_grid::kFrustumRing;
    CU(cudaEventSynchronize(g->ev_frustums[slot]));
    DevFrustum * hf = g->h_frustums[slot];
    for (uint32_t i = 0; i < n_frustums; ++i) {
      prepare_frustum(frustums[i], g->vs, &hf[i]);
      if (!(frustums[i].accel_factor >= 0.0)) may_raise = true;
    }
    CU(cudaMemcpyAsync(g->frustums.p, hf, sizeof(DevFrustum) * n_frustums, cudaMemcpyHostToDevice, g->stream));
    CU(cudaEventRecord(g->ev_frustums[slot], g->stream));
    frustums_dev = g->frustums.p;
  }
  // stored times only ever go down in a sweep if now >= every stored time and all factors are >= 0
  if (!(now >= g->time_bound) || may_raise) g->time_bound_inf = true;

  SweepOutputs out{};
  // list buffers are sized by the active-voxel bound; in a long run without synchronising calls that bound only
  // grows (by the points marked), so tighten it with one counter fetch before paying for a reallocation
  const bool wants_lists = g->cfg.pub_voxels || (g->out_flags & STVL_OUT_CLEARED);
  if (wants_lists && !g->counters_fresh &&
      ((g->cfg.pub_voxels && g->active_upper > g->points.cap / 3) || ((g->out_flags & STVL_OUT_CLEARED) && g->active_upper > g->cleared.cap / 2))) {
    rc = fetch_counters(g);
    if (rc != STVL_OK) return rc;
  }
  const uint64_t bound = g->active_upper;
  if (g->cfg.pub_voxels) {
    rc = ensure(g, g->points, (size_t)bound * 3 + 3); if (rc != STVL_OK) return rc;
    out.points = g->points.p; out.points_cap = bound;
  }
  if (g->out_flags & STVL_OUT_CLEARED) {
    rc = ensure(g, g->cleared, (size_t)bound * 2 + 2); if (rc != STVL_OK) return rc;
    out.cleared = g->cleared.p; out.cleared_cap = bound;
  }
  if (g->out_flags & STVL_OUT_AMBIGUOUS) {
    const size_t cap = 65536;
    rc = ensure(g, g->ambiguous, cap * 3); if (rc != STVL_OK) return rc;
    out.ambiguous = g->ambiguous.p; out.ambiguous_cap = cap;
  }
  select_sweep_buffers(g, g->sweep_parity ^ 1);   // this sweep's tile buffer / counter set (re-armed by the previous sweep)
  StageTimer timer__(g, STVL_TIME_SWEEP);
  if (fz && fz->costmap) {
    HOST_T(ts0);
    // fused cycle: the sweep also resets the costmap bytes (or this rank's cell bitmap) and shares the SMs with the Mark
    // kernel that streams its clouds behind it
    const size_t reset_bytes = fz->bits ? (size_t)fz->cm.bit_wpr * 4u * fz->cm.size_y : (size_t)fz->cm.size_x * fz->cm.size_y;
    CU(launch_sweep(g->v, frustums_dev, &fparam, (int)n_frustums, now, out, slots_upper(g), g->sm_count - g->sm_spare, g->stream, true, true, fz->costmap,
                    reset_bytes, fz->bits ? 0 : (int)fz->cm.default_value));
    HOST_T(ts1);
    HOST_ADD(2, ts0, ts1);
    fz->pdl = !(g->timing_mask & STVL_TIME_SWEEP);   // nothing may be enqueued between the sweep and Mark
  } else {
    // (the programmatic attribute is safe whatever precedes the sweep on the stream: before its wait it only reads kernel
    // parameters / the frustum buffer)
    CU(launch_sweep(g->v, frustums_dev, &fparam, (int)n_frustums, now, out, slots_upper(g), g->sm_count, g->stream, true, false));
  }
  g->launches += 1;
  g->counters_fresh = false;
  g->swept = true;
  g->added_since_sweep = 0;
  return STVL_OK;
}
}  // namespace

int stvl_clear_frustums(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums)
{
  ENTER(g);
  return clear_frustums_impl(g, now, frustums, n_frustums, nullptr);
}

int stvl_mark(stvl_grid_t * g, const stvl_reading_t * readings, uint32_t n_readings)
{
  ENTER(g);
  int rc = validate_readings(readings, n_readings);
  if (rc != STVL_OK) return rc;
  // a second Mark before the first was verified would overwrite its staged input: verify first
  rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  size_t total = 0;
  std::vector<size_t> offs(n_readings, 0);
  for (uint32_t i = 0; i < n_readings; ++i) {
    const stvl_reading_t & r = readings[i];
    if (!r.marking || r.n_points == 0) continue;
    offs[i] = total;
    total += ((size_t)r.n_points * r.point_step + 255) & ~(size_t)255;
  }
  if (total == 0) return STVL_OK;
  rc = ensure(g, g->stage, total);
  if (rc != STVL_OK) return rc;
  // H2D on the copy stream (overlaps a sweep still running on the main stream); the staging buffer may only be
  // overwritten once the previous Mark kernel has consumed it
  CU(cudaStreamWaitEvent(g->copy_stream, g->ev_mark_done, 0));   // recorded right behind the previous Mark's launches: the copy overlaps this cycle's sweep
  std::vector<const uint8_t *> dev_ptrs(n_readings, nullptr);
  for (uint32_t i = 0; i < n_readings; ++i) {
    const stvl_reading_t & r = readings[i];
    if (!r.marking || r.n_points == 0) continue;
    dev_ptrs[i] = g->stage.p + offs[i];
    CU(cudaMemcpyAsync(g->stage.p + offs[i], r.data, (size_t)r.n_points * r.point_step, cudaMemcpyHostToDevice, g->copy_stream));
  }
  CU(cudaEventRecord(g->ev_copy_done, g->copy_stream));
  CU(cudaStreamWaitEvent(g->stream, g->ev_copy_done, 0));
  g->record_mark_event = true;
  rc = run_mark(g, readings, n_readings, dev_ptrs);
  g->record_mark_event = false;
  if (rc != STVL_OK) return rc;
  // the caller owns the clouds again as soon as they are on the device
  CU(cudaEventSynchronize(g->ev_copy_done));
  return STVL_OK;
}

int stvl_mark_device(stvl_grid_t * g, const stvl_reading_t * readings, uint32_t n_readings)
{
  ENTER(g);
  int rc = validate_readings(readings, n_readings);
  if (rc != STVL_OK) return rc;
  rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  std::vector<const uint8_t *> dev_ptrs(n_readings, nullptr);
  for (uint32_t i = 0; i < n_readings; ++i) dev_ptrs[i] = static_cast<const uint8_t *>(readings[i].data);
  return run_mark(g, readings, n_readings, dev_ptrs);
}

int stvl_costmap_device(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap_dev, stvl_costmap_result_t * res)
{
  ENTER(g);
  if (!p || !costmap_dev) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  {
    StageTimer timer__(g, STVL_TIME_COSTMAP);
    CU(launch_costmap(g->v, to_dev(*p), costmap_dev, tiles_upper(g), g->sm_count, g->stream, true));
  }
  g->launches += 1;
  g->counters_fresh = false;
  if (!res) return STVL_OK;   // fully asynchronous: the caller synchronises / reads results later
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  fill_costmap_result(g, res);
  return after_fetch(g);
}

int stvl_costmap(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap, stvl_costmap_result_t * res)
{
  ENTER(g);
  if (!p) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  const size_t bytes = (size_t)p->size_x * p->size_y;
  const uint8_t * before = g->costmap.p;
  int rc = ensure(g, g->costmap, std::max<size_t>(bytes, 1));
  if (rc != STVL_OK) return rc;
  if (g->costmap.p != before) g->costmap_clean_value = -1;   // fresh allocation: contents unknown
  {
    StageTimer timer__(g, STVL_TIME_COSTMAP);
    const bool prefilled = g->costmap_clean_value == (int)p->default_value && g->costmap_clean_bytes == bytes;
    CU(launch_costmap(g->v, to_dev(*p), g->costmap.p, tiles_upper(g), g->sm_count, g->stream, !prefilled));
  }
  g->launches += 1;
  if (costmap && bytes) CU(cudaMemcpyAsync(costmap, g->costmap.p, bytes, cudaMemcpyDeviceToHost, g->stream));
  rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  // resetMaps for the NEXT cycle, off the critical path (runs while the caller works on this cycle's bytes)
  if (bytes) {
    CU(cudaMemsetAsync(g->costmap.p, p->default_value, bytes, g->stream));
    g->costmap_clean_value = (int)p->default_value;
    g->costmap_clean_bytes = bytes;
  }
  fill_costmap_result(g, res);
  return after_fetch(g);
}

int stvl_get_sweep_stats(stvl_grid_t * g, stvl_sweep_stats_t * out)
{
  ENTER(g);
  if (!out) return fail(STVL_ERR_INVALID, "out is NULL");
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  std::memset(out, 0, sizeof(*out));
  const SweepCtr & c = g->last.sw[g->sweep_parity];
  out->n_swept = c.n_swept; out->n_survived = c.n_survived; out->n_cleared = c.n_cleared;
  out->n_rewritten = c.n_rewritten; out->n_ambiguous = c.n_ambiguous;
  if (c.n_cleared) {
    out->cleared_bbox_index[0] = c.cb_min_x; out->cleared_bbox_index[1] = c.cb_min_y;
    out->cleared_bbox_index[2] = c.cb_max_x; out->cleared_bbox_index[3] = c.cb_max_y;
    world_bbox(g, out->cleared_bbox_index, out->cleared_bbox_world);
  }
  return after_fetch(g);
}

int stvl_get_columns(stvl_grid_t * g, int32_t * ix, int32_t * iy, uint32_t * count, uint64_t cap, uint64_t * n)
{
  ENTER(g);
  if (!n) return fail(STVL_ERR_INVALID, "n is NULL");
  int rc;
  if (cap) {
    rc = ensure(g, g->col_ix, cap); if (rc != STVL_OK) return rc;
    rc = ensure(g, g->col_iy, cap); if (rc != STVL_OK) return rc;
    rc = ensure(g, g->col_cnt, cap); if (rc != STVL_OK) return rc;
  }
  CU(launch_columns(g->v, cap ? g->col_ix.p : nullptr, cap ? g->col_iy.p : nullptr, cap ? g->col_cnt.p : nullptr, cap,
                    tiles_upper(g), g->sm_count, g->stream));
  g->launches += 2;
  rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  *n = g->last.n_columns_out;
  const uint64_t m = std::min<uint64_t>(*n, cap);
  if (m) {
    if (ix) CU(cudaMemcpyAsync(ix, g->col_ix.p, m * 4, cudaMemcpyDeviceToHost, g->stream));
    if (iy) CU(cudaMemcpyAsync(iy, g->col_iy.p, m * 4, cudaMemcpyDeviceToHost, g->stream));
    if (count) CU(cudaMemcpyAsync(count, g->col_cnt.p, m * 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
  }
  rc = after_fetch(g);
  if (rc != STVL_OK) return rc;
  return *n > cap && cap ? fail(STVL_ERR_RANGE, "column buffer too small: %llu > %llu", (unsigned long long)*n, (unsigned long long)cap) : STVL_OK;
}

int stvl_get_cleared(stvl_grid_t * g, int32_t * ix, int32_t * iy, uint64_t cap, uint64_t * n)
{
  ENTER(g);
  if (!n) return fail(STVL_ERR_INVALID, "n is NULL");
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  *n = (g->out_flags & STVL_OUT_CLEARED) ? g->last.sw[g->sweep_parity].n_cleared_out : 0;
  const uint64_t m = std::min<uint64_t>(*n, cap);
  if (m && (ix || iy)) {
    std::vector<int32_t> tmp(m * 2);
    CU(cudaMemcpy(tmp.data(), g->cleared.p, m * 8, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < m; ++i) { if (ix) ix[i] = tmp[2 * i]; if (iy) iy[i] = tmp[2 * i + 1]; }
  }
  rc = after_fetch(g);
  if (rc != STVL_OK) return rc;
  return *n > cap && cap ? fail(STVL_ERR_RANGE, "cleared buffer too small") : STVL_OK;
}

int stvl_get_points(stvl_grid_t * g, float * xyz, uint64_t cap, uint64_t * n)
{
  ENTER(g);
  if (!n) return fail(STVL_ERR_INVALID, "n is NULL");
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  *n = g->cfg.pub_voxels ? g->last.sw[g->sweep_parity].n_points_out : 0;
  const uint64_t m = std::min<uint64_t>(*n, cap);
  if (m && xyz) CU(cudaMemcpy(xyz, g->points.p, m * 12, cudaMemcpyDeviceToHost));
  rc = after_fetch(g);
  if (rc != STVL_OK) return rc;
  return *n > cap && cap ? fail(STVL_ERR_RANGE, "point buffer too small") : STVL_OK;
}

int stvl_get_ambiguous(stvl_grid_t * g, int32_t * ix, int32_t * iy, int32_t * iz, uint64_t cap, uint64_t * n)
{
  ENTER(g);
  if (!n) return fail(STVL_ERR_INVALID, "n is NULL");
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  *n = (g->out_flags & STVL_OUT_AMBIGUOUS) ? std::min<uint64_t>(g->last.sw[g->sweep_parity].n_ambiguous_out, 65536) : 0;
  const uint64_t m = std::min<uint64_t>(*n, cap);
  if (m) {
    std::vector<int32_t> tmp(m * 3);
    CU(cudaMemcpy(tmp.data(), g->ambiguous.p, m * 12, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < m; ++i) { if (ix) ix[i] = tmp[3 * i]; if (iy) iy[i] = tmp[3 * i + 1]; if (iz) iz[i] = tmp[3 * i + 2]; }
  }
  return after_fetch(g);
}

int stvl_reset(stvl_grid_t * g)
{
  ENTER(g);
  int rc = clear_pool(g);
  if (rc != STVL_OK) return rc;
  CU(cudaStreamSynchronize(g->stream));
  return STVL_OK;
}

int stvl_reset_area(stvl_grid_t * g, double start_x, double start_y, double end_x, double end_y, int invert_area)
{
  ENTER(g);
  int rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  CU(launch_reset_area(g->v, start_x, start_y, end_x, end_y, invert_area, slots_upper(g), g->sm_count, g->stream));
  g->launches += 1;
  return STVL_OK;
}

int stvl_active_count(stvl_grid_t * g, uint64_t * n)
{
  ENTER(g);
  if (!n) return fail(STVL_ERR_INVALID, "n is NULL");
  int rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  CU(cudaMemsetAsync(&g->v.ctr->n_active, 0, 8, g->stream));
  CU(cudaMemsetAsync(&g->v.ctr->live_leaves, 0, 4, g->stream));
  CU(launch_count_active(g->v, slots_upper(g), g->sm_count, g->stream));
  g->launches += 1;
  rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  *n = g->last.n_active;
  if (!g->swept) g->active_upper = g->last.n_active;
  return STVL_OK;
}

int stvl_get_voxels(stvl_grid_t * g, int32_t * ix, int32_t * iy, int32_t * iz, double * t, uint64_t cap, uint64_t * n)
{
  ENTER(g);
  if (!n) return fail(STVL_ERR_INVALID, "n is NULL");
  int rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  if (cap) {
    rc = ensure(g, g->dump_x, cap); if (rc != STVL_OK) return rc;
    rc = ensure(g, g->dump_y, cap); if (rc != STVL_OK) return rc;
    rc = ensure(g, g->dump_z, cap); if (rc != STVL_OK) return rc;
    rc = ensure(g, g->dump_t, cap); if (rc != STVL_OK) return rc;
  }
  CU(cudaMemsetAsync(&g->v.ctr->n_voxels_out, 0, 8, g->stream));
  CU(launch_dump_voxels(g->v, cap ? g->dump_x.p : nullptr, cap ? g->dump_y.p : nullptr, cap ? g->dump_z.p : nullptr,
                        cap ? g->dump_t.p : nullptr, cap, slots_upper(g), g->sm_count, g->stream));
  g->launches += 1;
  rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  *n = g->last.n_voxels_out;
  const uint64_t m = std::min<uint64_t>(*n, cap);
  if (m) {
    if (ix) CU(cudaMemcpyAsync(ix, g->dump_x.p, m * 4, cudaMemcpyDeviceToHost, g->stream));
    if (iy) CU(cudaMemcpyAsync(iy, g->dump_y.p, m * 4, cudaMemcpyDeviceToHost, g->stream));
    if (iz) CU(cudaMemcpyAsync(iz, g->dump_z.p, m * 4, cudaMemcpyDeviceToHost, g->stream));
    if (t) CU(cudaMemcpyAsync(t, g->dump_t.p, m * 8, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
  }
  return *n > cap && cap ? fail(STVL_ERR_RANGE, "voxel buffer too small") : STVL_OK;
}

int stvl_load_voxels(stvl_grid_t * g, const int32_t * ix, const int32_t * iy, const int32_t * iz, const double * t, uint64_t n)
{
  ENTER(g);
  if (n == 0) return STVL_OK;
  if (!ix || !iy || !iz || !t) return fail(STVL_ERR_INVALID, "NULL argument");
  int rc = verify_marks(g);
  if (rc != STVL_OK) return rc;
  rc = ensure(g, g->dump_x, n); if (rc != STVL_OK) return rc;
  rc = ensure(g, g->dump_y, n); if (rc != STVL_OK) return rc;
  rc = ensure(g, g->dump_z, n); if (rc != STVL_OK) return rc;
  rc = ensure(g, g->dump_t, n); if (rc != STVL_OK) return rc;
  CU(cudaMemcpyAsync(g->dump_x.p, ix, n * 4, cudaMemcpyHostToDevice, g->stream));
  CU(cudaMemcpyAsync(g->dump_y.p, iy, n * 4, cudaMemcpyHostToDevice, g->stream));
  CU(cudaMemcpyAsync(g->dump_z.p, iz, n * 4, cudaMemcpyHostToDevice, g->stream));
  CU(cudaMemcpyAsync(g->dump_t.p, t, n * 8, cudaMemcpyHostToDevice, g->stream));
  double tmax = -std::numeric_limits<double>::infinity();
  for (uint64_t i = 0; i < n; ++i) tmax = std::max(tmax, t[i]);
  for (int attempt = 0; attempt < 32; ++attempt) {
    CU(launch_load_voxels(g->v, g->dump_x.p, g->dump_y.p, g->dump_z.p, g->dump_t.p, n, g->stream));
    g->launches += 1;
    rc = fetch_counters(g);
    if (rc != STVL_OK) return rc;
    if (!g->last.leaf_overflow && !g->last.tile_overflow) break;
    if (g->last.leaf_overflow) { rc = grow_leaves(g); if (rc != STVL_OK) return rc; }
    if (g->last.tile_overflow) { rc = grow_tiles(g); if (rc != STVL_OK) return rc; }
    if (attempt == 31) return fail(STVL_ERR_CAPACITY, "leaf pool still exhausted after repeated growth");
  }
  g->active_upper += n;
  g->added_since_sweep += n;
  if (!(tmax <= g->time_bound)) { if (std::isfinite(tmax)) g->time_bound = std::max(g->time_bound, tmax); else g->time_bound_inf = true; }
  return STVL_OK;
}

double stvl_index_to_world(const stvl_grid_t * g, int32_t index)
{
  if (!g) return std::numeric_limits<double>::quiet_NaN();
  return (double)index * g->vs + g->half_vs;
}
double stvl_voxel_size(const stvl_grid_t * g) { return g ? g->vs : std::numeric_limits<double>::quiet_NaN(); }

int stvl_columns_to_dense(stvl_grid_t * g, int32_t ix0, int32_t iy0, uint32_t nx, uint32_t ny, uint32_t * dense_dev)
{
  ENTER(g);
  if (!dense_dev) return fail(STVL_ERR_INVALID, "dense_dev is NULL");
  CU(launch_columns_to_dense(g->v, ix0, iy0, nx, ny, dense_dev, tiles_upper(g), g->sm_count, g->stream));
  g->launches += 1;
  return STVL_OK;   // asynchronous on the handle's stream (stvl_get_stream): order the collective after it
}

int stvl_costmap_from_dense(stvl_grid_t * g, const uint32_t * dense_dev, int32_t ix0, int32_t iy0, uint32_t nx, uint32_t ny,
                            const stvl_costmap_params_t * p, uint8_t * costmap_dev, stvl_costmap_result_t * res)
{
  ENTER(g);
  if (!dense_dev || !p || !costmap_dev) return fail(STVL_ERR_INVALID, "NULL argument");
  CU(launch_costmap_from_dense(g->v, dense_dev, ix0, iy0, nx, ny, to_dev(*p), costmap_dev, g->sm_count, g->stream));
  g->launches += 1;
  g->counters_fresh = false;
  if (!res) return STVL_OK;   // fully asynchronous
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  fill_costmap_result(g, res);
  return after_fetch(g);
}

int stvl_shard_word_intervals(float voxel_size, int shard_count, int shard_width, int n_ranks, const stvl_costmap_params_t * p,
                              int32_t * w_lo, int32_t * w_hi)
{
  if (!p || !w_lo || !w_hi) return fail(STVL_ERR_INVALID, "NULL argument");
  if (n_ranks < 1 || n_ranks > kMaxShardRanks) return fail(STVL_ERR_INVALID, "n_ranks must be 1..%d", kMaxShardRanks);
  if (!(voxel_size > 0.f) || !(p->resolution > 0)) return fail(STVL_ERR_INVALID, "voxel_size and resolution must be positive");
  int lo[kMaxShardRanks], hi[kMaxShardRanks];
  compute_word_intervals((double)voxel_size, shard_count, shard_width, n_ranks, 0, *p, lo, hi);
  for (int r = 0; r < n_ranks; ++r) { w_lo[r] = lo[r]; w_hi[r] = hi[r]; }
  return STVL_OK;
}

int stvl_nccl_unique_id(uint8_t id[128])
{
  if (!id) return fail(STVL_ERR_INVALID, "id is NULL");
  NcclApi * n = nccl_api();
  if (!n) return fail(STVL_ERR_INVALID, "libnccl.so.2 could not be loaded");
  ncclUniqueId u;
  NC(n->GetUniqueId(&u));
  static_assert(sizeof(u) == 128, "ncclUniqueId size");
  std::memcpy(id, &u, 128);
  return STVL_OK;
}

int stvl_comm_init(stvl_grid_t * g, const uint8_t id[128], int rank, int n_ranks)
{
  ENTER(g);
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(STVL_ERR_INVALID, "bad communicator arguments");
  NcclApi * n = nccl_api();
  if (!n) return fail(STVL_ERR_INVALID, "libnccl.so.2 could not be loaded");
  if (g->comm) return fail(STVL_ERR_INVALID, "communicator already initialised");
  ncclUniqueId u;
  std::memcpy(&u, id, 128);
  NC(n->CommInitRank(&g->comm, n_ranks, u, rank));
  g->comm_rank = rank; g->comm_n = n_ranks;
  CU(cudaStreamCreateWithFlags(&g->merge_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    CU(cudaEventCreateWithFlags(&g->ev_scatter[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g->ev_reduce[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g->ev_bits_done[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g->ev_expand[k], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&g->ev_main_pos, cudaEventDisableTiming));
  g->bits_geom_valid = false;
  g->bits_pending = -1;
  return STVL_OK;
}

int stvl_comm_destroy(stvl_grid_t * g)
{
  ENTER(g);
  if (!g->comm) return STVL_OK;
  CU(cudaStreamSynchronize(g->stream));
  CU(cudaStreamSynchronize(g->merge_stream));
  nccl_api()->CommDestroy(g->comm);
  g->comm = nullptr;
  for (int k = 0; k < 2; ++k) {
    cudaEventDestroy(g->ev_scatter[k]); cudaEventDestroy(g->ev_reduce[k]); cudaEventDestroy(g->ev_bits_done[k]); cudaEventDestroy(g->ev_expand[k]);
    g->ev_scatter[k] = g->ev_reduce[k] = g->ev_bits_done[k] = g->ev_expand[k] = nullptr;
  }
  cudaEventDestroy(g->ev_main_pos); g->ev_main_pos = nullptr;
  g->bits_pending = -1;
  cudaStreamDestroy(g->merge_stream);
  g->merge_stream = nullptr;
  g->win_pending = -1;
  return STVL_OK;
}

int stvl_costmap_sharded_device(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root,
                                int count_bytes, int pipelined)
{
  ENTER(g);
  if (!p) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!g->comm) return fail(STVL_ERR_INVALID, "stvl_comm_init has not been called");
  if (count_bytes != 0 && count_bytes != 1 && count_bytes != 4) return fail(STVL_ERR_INVALID, "count_bytes must be 0 (bytes mode), 1 or 4");
  if (root < 0 || root >= g->comm_n) return fail(STVL_ERR_INVALID, "bad root");
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  int geom[4];
  window_geometry(g, *p, geom);
  if (count_bytes == 0) { geom[0] = geom[1] = 0; geom[2] = (int)p->size_x; geom[3] = (int)p->size_y; }   // the window IS a costmap
  const size_t cells = (size_t)geom[2] * geom[3];
  if (g->win_pending >= 0 && (count_bytes != g->win_count_bytes || std::memcmp(geom, g->win_geom, sizeof(geom)) != 0)) {
    // the costmap moved or changed size while a window was in flight: finish that one first
    int rc = stvl_costmap_sharded_flush(g, p, nullptr, root);
    if (rc != STVL_OK) return rc;
  }
  const int b = (int)(g->win_k & 1);
  const size_t cell_bytes = count_bytes == 0 ? 1 : (size_t)count_bytes;
  if (g->window[b].cap < cells * cell_bytes) {
    CU(cudaStreamSynchronize(g->merge_stream));
    int rc = ensure(g, g->window[b], cells * cell_bytes);
    if (rc != STVL_OK) return rc;
  }
  std::memcpy(g->win_geom, geom, sizeof(geom));
  g->win_count_bytes = count_bytes;
  // the reduce that used this window two calls ago must be done before it is overwritten
  CU(cudaStreamWaitEvent(g->stream, g->ev_reduce[b], 0));
  if (count_bytes == 0) {
    // bytes mode (shards that own whole voxel columns, i.e. x / y slabs): threshold locally into {0, lethal} bytes —
    // exactly stvl_costmap with default 0 — and let the reduce (max) assemble the costmap in the root's buffer
    CostmapDev d = to_dev(*p);
    d.default_value = 0;
    StageTimer timer__(g, STVL_TIME_COSTMAP);
    CU(launch_costmap(g->v, d, g->window[b].p, tiles_upper(g), g->sm_count, g->stream, true));
    g->counters_fresh = false;
  } else {
    CU(launch_columns_to_window(g->v, geom[0], geom[1], (uint32_t)geom[2], (uint32_t)geom[3], g->window[b].p, count_bytes,
                                tiles_upper(g), g->sm_count, g->stream));
  }
  g->launches += 1;
  CU(cudaEventRecord(g->ev_scatter[b], g->stream));
  NcclApi * n = nccl_api();
  const ncclDataType_t dt = count_bytes == 4 ? ncclUint32 : ncclUint8;
  const ncclRedOp_t op = count_bytes == 0 ? ncclMax : ncclSum;
  void * recv = g->window[b].p;   // in place
  g->win_k += 1;
  if (!pipelined) {
    if (g->win_pending >= 0) { int rc = threshold_window(g, g->win_pending, *p, nullptr, root); if (rc != STVL_OK) return rc; g->win_pending = -1; }
    NC(n->Reduce(g->window[b].p, recv, cells, dt, op, root, g->comm, g->stream));
    CU(cudaEventRecord(g->ev_reduce[b], g->stream));
    return threshold_window(g, b, *p, costmap_dev, root);
  }
  // pipelined: the reduce runs on the side stream behind the scatter; the library's stream goes on with the next cycle
  CU(cudaStreamWaitEvent(g->merge_stream, g->ev_scatter[b], 0));
  NC(n->Reduce(g->window[b].p, recv, cells, dt, op, root, g->comm, g->merge_stream));
  CU(cudaEventRecord(g->ev_reduce[b], g->merge_stream));
  if (g->win_pending >= 0) {
    int rc = threshold_window(g, g->win_pending, *p, costmap_dev, root);   // the PREVIOUS call's bytes
    if (rc != STVL_OK) return rc;
  }
  g->win_pending = b;
  return STVL_OK;
}

int stvl_costmap_sharded_flush(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root)
{
  ENTER(g);
  if (!p) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!g->comm) return fail(STVL_ERR_INVALID, "stvl_comm_init has not been called");
  if (g->bits_pending >= 0) {   // the sharded fused cycle's bitmap exchange: expand the last cycle's cells now
    int rc = expand_pending_bits(g, *p, costmap_dev, root);
    if (rc != STVL_OK) return rc;
  }
  if (g->win_pending < 0) return STVL_OK;
  const int b = g->win_pending;
  g->win_pending = -1;
  return threshold_window(g, b, *p, costmap_dev, root);
}

int stvl_get_pool_stats(stvl_grid_t * g, stvl_pool_stats_t * out)
{
  ENTER(g);
  if (!out) return fail(STVL_ERR_INVALID, "out is NULL");
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  std::memset(out, 0, sizeof(*out));
  const Counters & c = g->last;
  out->leaf_capacity = g->v.leaf_capacity;
  out->leaves_free = ctr_free_n(c);
  out->leaves_in_use = ctr_high_water(g, c) - ctr_free_n(c);
  out->hash_capacity = (uint64_t)g->v.hash_mask + 1;
  out->hash_tombstones = c.n_tombs;
  out->tiles_in_use = c.tile_n;
  out->tile_capacity = g->v.tile_capacity;
  out->points_dropped_out_of_range = c.dropped_range;
  out->points_dropped_nonfinite = c.dropped_nonfinite;
  out->kernel_launches = g->launches;
  out->device_bytes = g->device_bytes;
  return after_fetch(g);
}

namespace {
// Fused form of clear -> mark -> costmap, three launches: the triage kernel also resets the costmap bytes, the leaf
// pass and Mark are chained by programmatic dependent launch (Mark's cloud streaming overlaps the sweep's tail; it
// waits for the sweep before it touches the grid) and the costmap pass rides on the Mark kernel.  The clouds must be
// complete on the device before this call is enqueued (stream order on stvl_get_stream() is enough).
int fused_cycle(stvl_grid * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums, const stvl_reading_t * readings,
                uint32_t n_readings, const CostmapDev & cm, uint8_t * costmap_dev, bool bits = false)
{
  HOST_T(t0);
  int rc = validate_readings(readings, n_readings);
  if (rc != STVL_OK) return rc;
  CycleFusion fz;
  fz.cm = cm;
  fz.costmap = costmap_dev;
  fz.bits = bits;
  rc = clear_frustums_impl(g, now, frustums, n_frustums, &fz);
  if (rc != STVL_OK) return rc;
  HOST_T(t1);
  std::vector<const uint8_t *> dev_ptrs(n_readings, nullptr);
  for (uint32_t i = 0; i < n_readings; ++i) dev_ptrs[i] = static_cast<const uint8_t *>(readings[i].data);
  rc = run_mark(g, readings, n_readings, dev_ptrs, &fz);
  if (rc != STVL_OK) return rc;
  HOST_T(t2);
  HOST_ADD(0, t0, t1); HOST_ADD(1, t1, t2);
#ifdef STVL_TRACE
  ++g_host_calls;
#endif
  if (!fz.costmap_done) {   // nothing to mark in this cycle: the plain costmap kernel (the bytes are reset already)
    StageTimer timer__(g, STVL_TIME_COSTMAP);
    CU(launch_costmap(g->v, fz.cm, costmap_dev, tiles_upper(g), g->sm_count, g->stream, false, false, bits));
    g->launches += 1;
  }
  g->counters_fresh = false;
  return STVL_OK;
}
}  // namespace

int stvl_update_cycle_device(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                             const stvl_reading_t * readings, uint32_t n_readings,
                             const stvl_costmap_params_t * p, uint8_t * costmap_dev)
{
  ENTER(g);
  if (!p || !costmap_dev) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  return fused_cycle(g, now, frustums, n_frustums, readings, n_readings, to_dev(*p), costmap_dev);
}

namespace {
int sharded_cycle(stvl_grid * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                  const stvl_reading_t * readings, uint32_t n_readings,
                  const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root, int pipelined);
}  // namespace

int stvl_update_cycle_sharded_device(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                                     const stvl_reading_t * readings, uint32_t n_readings,
                                     const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root, int pipelined)
{
  ENTER(g);
  return sharded_cycle(g, now, frustums, n_frustums, readings, n_readings, p, costmap_dev, root, pipelined);
}

namespace {
int sharded_cycle(stvl_grid * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                  const stvl_reading_t * readings, uint32_t n_readings,
                  const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root, int pipelined)
{
  if (!p) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!g->comm) return fail(STVL_ERR_INVALID, "stvl_comm_init has not been called");
  if (root < 0 || root >= g->comm_n) return fail(STVL_ERR_INVALID, "bad root");
  if (g->comm_n > kMaxShardRanks) return fail(STVL_ERR_INVALID, "at most %d ranks", kMaxShardRanks);
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  if (g->comm_rank == root && !costmap_dev) return fail(STVL_ERR_INVALID, "costmap_dev is NULL on the root rank");
  // x-slab shards own whole voxel columns, so the merged costmap is the union of the ranks' lethal cells: every rank's
  // costmap pass writes a packed BITMAP of the cells of its own x interval (1 bit per cell instead of a byte, and only the
  // rank's own interval instead of an N-times window), the bitmaps travel to the root with one grouped ncclSend / ncclRecv
  // on a side stream, and the root expands them into bytes (which is also its Costmap2D::resetMaps).
  if (g->bits_pending >= 0 && (!g->bits_geom_valid || std::memcmp(&g->bits_params, p, sizeof(*p)) != 0)) {
    int rc = expand_pending_bits(g, g->bits_params, costmap_dev, root);   // the costmap moved or changed size: finish the old one first
    if (rc != STVL_OK) return rc;
  }
  if (!g->bits_geom_valid || std::memcmp(&g->bits_params, p, sizeof(*p)) != 0) {
    CU(cudaStreamSynchronize(g->merge_stream));
    shard_word_intervals(g, *p);
    g->bits_params = *p;
    g->bits_geom_valid = true;
  }
  const int me = g->comm_rank;
  const size_t my_words = (size_t)(g->bit_w_hi[me] - g->bit_w_lo[me]) * p->size_y;
  size_t stage_words = 0;
  for (int r = 0; r < g->comm_n; ++r) if (r != root) stage_words += (size_t)(g->bit_w_hi[r] - g->bit_w_lo[r]) * p->size_y;
  const int b = (int)(g->bits_k & 1);
  if (g->bits[b].cap < std::max<size_t>(my_words, 1) || (me == root && g->bits_stage[b].cap < std::max<size_t>(stage_words, 1))) {
    CU(cudaStreamSynchronize(g->merge_stream));
    int rc = ensure(g, g->bits[b], std::max<size_t>(my_words, 1)); if (rc != STVL_OK) return rc;
    if (me == root) { rc = ensure(g, g->bits_stage[b], std::max<size_t>(stage_words, 1)); if (rc != STVL_OK) return rc; }
  }
  if (pipelined && g->bits_pending >= 0) {
    // cycle k-1's cells: expanded on the side stream while this cycle's kernels run on the main one.  The expansion must
    // not start before everything the caller enqueued so far (reads of cycle k-2's bytes) is done ...
    int rc = expand_pending_bits_async(g, *p, costmap_dev, root);
    if (rc != STVL_OK) return rc;
  }
  CU(cudaStreamWaitEvent(g->stream, g->ev_bits_done[b], 0));   // the exchange that used this bitmap two cycles ago
  CostmapDev d = to_dev(*p);
  d.bit_w_lo = g->bit_w_lo[me];
  d.bit_wpr = g->bit_w_hi[me] - g->bit_w_lo[me];
  // NCCL's send / receive kernels of the previous cycle's exchange run on the side stream while this cycle's kernels run: a
  // persistent grid of one (Mark) or two (sweep) register-file-filling CTAs per SM would have to wait for every SM such a copy
  // CTA sits on, so the sharded cycle leaves a few SMs to them: 4 on a sending rank, 2 + 2 per peer (at most 16) on the root,
  // which receives from every peer (STVL_SHARD_SPARE_SMS overrides; measured at N = 2: 60.1 us per cycle with none, 57.6 with 4,
  // 53.7 with 4 and NCCL_MAX_P2P_NCHANNELS=1)
  static const int spare_env = [] { const char * v = std::getenv("STVL_SHARD_SPARE_SMS"); const int x = v ? std::atoi(v) : -1; return x > 32 ? 32 : x; }();
  const int spare_dflt = g->comm_rank == root ? std::min(2 + 2 * (g->comm_n - 1), 16) : 4;
  g->sm_spare = g->comm_n > 1 ? std::min(spare_env >= 0 ? spare_env : spare_dflt, g->sm_count / 2) : 0;
  int rc = fused_cycle(g, now, frustums, n_frustums, readings, n_readings, d, reinterpret_cast<uint8_t *>(g->bits[b].p), true);
  g->sm_spare = 0;
  if (rc != STVL_OK) return rc;
  g->bits_k += 1;
  cudaStream_t xs = pipelined ? g->merge_stream : g->stream;
  if (pipelined) {
    CU(cudaEventRecord(g->ev_scatter[b], g->stream));
    CU(cudaStreamWaitEvent(g->merge_stream, g->ev_scatter[b], 0));
  }
  rc = exchange_bits(g, b, p->size_y, root, xs);
  if (rc != STVL_OK) return rc;
  CU(cudaEventRecord(g->ev_bits_done[b], xs));
  if (pipelined) {
    // ... and this call returns with the main stream ordered behind the expansion of cycle k-1: what the caller enqueues
    // next sees cycle k-1's bytes
    if (g->bits_expanding >= 0) { CU(cudaStreamWaitEvent(g->stream, g->ev_expand[g->bits_expanding], 0)); g->bits_expanding = -1; }
    g->bits_pending = b;
    return STVL_OK;
  }
  g->bits_pending = b;
  return expand_pending_bits(g, *p, costmap_dev, root);
}
}  // namespace

int stvl_update_cycle_wait(stvl_grid_t * g, uint64_t ticket, stvl_costmap_result_t * res)
{
  ENTER(g);
  for (auto & a : g->aslot) {
    if (!a.busy || a.ticket != ticket) continue;
    CU(cudaEventSynchronize(a.ev_d2h));
    a.busy = false;
    if (res) {
      std::memset(res, 0, sizeof(*res));
      res->n_marked_columns = a.h_res->n_marked;
      res->has_bounds = a.h_res->n_marked > 0;
      if (res->has_bounds) {
        const int32_t idx[4] = {a.h_res->mk_min_x, a.h_res->mk_min_y, a.h_res->mk_max_x, a.h_res->mk_max_y};
        world_bbox(g, idx, res->touched);
      }
    }
    return STVL_OK;
  }
  return fail(STVL_ERR_INVALID, "ticket %llu is not in flight", (unsigned long long)ticket);
}

int stvl_update_cycle_async(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                            const stvl_reading_t * readings, uint32_t n_readings,
                            const stvl_costmap_params_t * p, uint8_t * costmap_host, uint64_t * ticket)
{
  ENTER(g);
  if (!p || !ticket) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  int rc = validate_readings(readings, n_readings);
  if (rc != STVL_OK) return rc;
  const uint64_t t = g->next_ticket;
  stvl_grid::AsyncSlot & a = g->aslot[t & 1];
  if (a.busy) {   // two cycles in flight already: the older one has to finish (its result is dropped unless it was waited for)
    CU(cudaEventSynchronize(a.ev_d2h));
    a.busy = false;
  }
  if (!a.ev_h2d) {
    CU(cudaEventCreateWithFlags(&a.ev_h2d, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&a.ev_cycle, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&a.ev_d2h, cudaEventDisableTiming));
    CU(cudaMalloc(&a.d_res, sizeof(CostmapCtr)));
    CU(cudaMallocHost(&a.h_res, sizeof(CostmapCtr)));
    if (!g->d2h_stream) CU(cudaStreamCreateWithFlags(&g->d2h_stream, cudaStreamNonBlocking));
  }
  // staging layout of this cycle's clouds
  size_t total = 0;
  std::vector<size_t> offs(n_readings, 0);
  for (uint32_t i = 0; i < n_readings; ++i) {
    const stvl_reading_t & r = readings[i];
    if (!r.marking || r.n_points == 0) continue;
    offs[i] = total;
    total += ((size_t)r.n_points * r.point_step + 255) & ~(size_t)255;
  }
  const size_t bytes = (size_t)p->size_x * p->size_y;
  rc = ensure(g, a.stage, std::max<size_t>(total, 1)); if (rc != STVL_OK) return rc;
  rc = ensure(g, a.costmap, std::max<size_t>(bytes, 1)); if (rc != STVL_OK) return rc;
  // H2D on the copy stream: the staging buffer is free once the Mark kernel of the cycle that used it (two cycles ago) has run
  CU(cudaStreamWaitEvent(g->copy_stream, a.ev_cycle, 0));
  std::vector<stvl_reading_t> rs(readings, readings + n_readings);
  for (uint32_t i = 0; i < n_readings; ++i) {
    const stvl_reading_t & r = readings[i];
    if (!r.marking || r.n_points == 0) continue;
    CU(cudaMemcpyAsync(a.stage.p + offs[i], r.data, (size_t)r.n_points * r.point_step, cudaMemcpyHostToDevice, g->copy_stream));
    rs[i].data = a.stage.p + offs[i];
  }
  CU(cudaEventRecord(a.ev_h2d, g->copy_stream));
  // the fused cycle on the main stream; the device costmap of this slot is free once its previous read-back is done
  CU(cudaStreamWaitEvent(g->stream, a.ev_h2d, 0));
  CU(cudaStreamWaitEvent(g->stream, a.ev_d2h, 0));
  const bool sharded = g->comm != nullptr;     // a shard of a multi-GPU map: this rank's cells are merged into rank 0's costmap
  if (sharded) rc = sharded_cycle(g, now, frustums, n_frustums, rs.data(), n_readings, p, a.costmap.p, 0, 0);
  else rc = fused_cycle(g, now, frustums, n_frustums, rs.data(), n_readings, to_dev(*p), a.costmap.p);
  if (rc != STVL_OK) return rc;
  CU(cudaMemcpyAsync(a.d_res, &g->v.ctr->cm, sizeof(CostmapCtr), cudaMemcpyDeviceToDevice, g->stream));
  CU(cudaEventRecord(a.ev_cycle, g->stream));
  // D2H on its own stream: overlaps the next cycle's kernels and H2D copy
  CU(cudaStreamWaitEvent(g->d2h_stream, a.ev_cycle, 0));
  if (costmap_host && bytes && (!sharded || g->comm_rank == 0))
    CU(cudaMemcpyAsync(costmap_host, a.costmap.p, bytes, cudaMemcpyDeviceToHost, g->d2h_stream));
  CU(cudaMemcpyAsync(a.h_res, a.d_res, sizeof(CostmapCtr), cudaMemcpyDeviceToHost, g->d2h_stream));
  CU(cudaEventRecord(a.ev_d2h, g->d2h_stream));
  a.busy = true;
  a.ticket = t;
  g->next_ticket = t + 1;
  *ticket = t;
  return STVL_OK;
}

int stvl_update_costs_device(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * layer_costmap_dev, uint8_t * master_dev,
                             const double * footprint_xy, uint32_t n_footprint, int combination_method,
                             int32_t min_i, int32_t min_j, int32_t max_i, int32_t max_j)
{
  ENTER(g);
  if (!p || !layer_costmap_dev) return fail(STVL_ERR_INVALID, "NULL argument");
  if (!(p->resolution > 0)) return fail(STVL_ERR_INVALID, "resolution must be positive");
  if (n_footprint && !footprint_xy) return fail(STVL_ERR_INVALID, "footprint_xy is NULL");
  if (n_footprint > (uint32_t)kMaxPolygonVertices) return fail(STVL_ERR_INVALID, "at most %d footprint vertices", kMaxPolygonVertices);
  if (master_dev && (min_i < 0 || min_j < 0 || max_i > (int64_t)p->size_x || max_j > (int64_t)p->size_y))
    return fail(STVL_ERR_INVALID, "window [%d,%d) x [%d,%d) outside the %u x %u map", min_i, max_i, min_j, max_j, p->size_x, p->size_y);
  if (n_footprint >= 3 && p->size_x && p->size_y) {
    int rc = ensure(g, g->poly_cols, (size_t)p->size_x * 2);
    if (rc != STVL_OK) return rc;
    PolygonParams poly{};
    poly.n = n_footprint;
    for (uint32_t i = 0; i < 2 * n_footprint; ++i) poly.xy[i] = footprint_xy[i];
    CU(launch_polygon_cost(layer_costmap_dev, to_dev(*p), poly, 0 /* FREE_SPACE */, g->poly_cols.p, g->poly_cols.p + p->size_x, g->stream));
    g->launches += 1;
  }
  if (master_dev && (combination_method == 0 || combination_method == 1)) {
    CU(launch_update_costs(layer_costmap_dev, master_dev, p->size_x, min_i, min_j, max_i, max_j, combination_method, g->sm_count, g->stream));
    g->launches += 1;
  }
  return STVL_OK;
}

int stvl_get_costmap_result(stvl_grid_t * g, stvl_costmap_result_t * res)
{
  ENTER(g);
  if (!res) return fail(STVL_ERR_INVALID, "NULL argument");
  int rc = fetch_counters(g);
  if (rc != STVL_OK) return rc;
  fill_costmap_result(g, res);
  return after_fetch(g);
}

int stvl_enable_timing(stvl_grid_t * g, uint32_t stage_mask)
{
  if (!g) return fail(STVL_ERR_INVALID, "null grid handle");
  g->timing_mask = stage_mask;
  return STVL_OK;
}

int stvl_get_timing(stvl_grid_t * g, stvl_timing_t * out)
{
  ENTER(g);
  if (!out) return fail(STVL_ERR_INVALID, "out is NULL");
  std::memset(out, 0, sizeof(*out));
  CU(cudaStreamSynchronize(g->stream));
  for (const auto & iv : g->timing_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, iv.a, iv.b) == cudaSuccess) {
      if (iv.stage == STVL_TIME_SWEEP) { out->sweep_ms += ms; ++out->n_sweep; }
      else if (iv.stage == STVL_TIME_MARK) { out->mark_ms += ms; ++out->n_mark; }
      else { out->costmap_ms += ms; ++out->n_costmap; }
    }
    g->timing_pool.push_back(iv.a);
    g->timing_pool.push_back(iv.b);
  }
  g->timing_pending.clear();
  return STVL_OK;
}

int stvl_get_stream(stvl_grid_t * g, void ** stream)
{
  if (!g || !stream) return fail(STVL_ERR_INVALID, "NULL argument");
  *stream = (void *)g->stream;
  return STVL_OK;
}

int stvl_synchronize(stvl_grid_t * g)
{
  ENTER(g);
  CU(cudaStreamSynchronize(g->copy_stream));
  CU(cudaStreamSynchronize(g->stream));
  return STVL_OK;
}

}  // extern "C"
