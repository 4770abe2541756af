/*
 * stvl_b200.h — C-ABI of the B200-native spatio-temporal voxel grid.
 *
 * This is the drop-in boundary.  The reference has no FFI for
 * this path: the caller is the C++ class volume_grid::SpatioTemporalVoxelGrid
 * (reference: spatio_temporal_voxel_layer/include/spatio_temporal_voxel_layer/
 * spatio_temporal_voxel_grid.hpp:122-189).  Every entry point below replaces the
 * OpenVDB-backed body of one of that class's methods; the C++ façade in
 * spatio_temporal_voxel_layer_b200/host/ keeps the reference signatures and calls
 * these.  Paths in the comments are relative to the reference package root
 * (`stvl/` = spatio_temporal_voxel_layer/).
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures.
 *   - every call returns an int status (STVL_OK == 0).  Nothing throws or aborts
 *     across the boundary; stvl_last_error() gives a message (reference error
 *     convention: bool + std::cout, stvl/src/spatio_temporal_voxel_grid.cpp:188,385-392).
 *   - `now` (seconds, double) is always an ARGUMENT.  The reference samples
 *     _clock->now() inside the hot functions (.cpp:155,275); the façade passes it.
 *   - one call at a time per handle (the façade holds _grid_lock like the
 *     reference, .cpp:101,258,382,401); calls may come from different host threads.
 *   - the caller owns every host pointer it passes; the library never keeps one
 *     past the call.  Output buffers are caller-allocated.
 *   - there is NO CPU fallback: if no CUDA device is usable stvl_create fails.
 */
#ifndef STVL_B200_H_
#define STVL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STVL_ABI_VERSION 2

/* status codes */
enum {
  STVL_OK = 0,
  STVL_ERR_INVALID = 1,   /* bad argument */
  STVL_ERR_CUDA = 2,      /* CUDA runtime error (message in stvl_last_error) */
  STVL_ERR_NO_DEVICE = 3, /* no usable sm_100 device: the library has no CPU path */
  STVL_ERR_CAPACITY = 4,  /* leaf pool exhausted and growth failed */
  STVL_ERR_RANGE = 5      /* output buffer too small */
};

/* decay models — GlobalDecayModel, stvl/include/.../spatio_temporal_voxel_grid.hpp:81-86.
 * Any value other than 0/1 takes the PERSISTENT branch (.cpp:320-331). */
enum { STVL_DECAY_LINEAR = 0, STVL_DECAY_EXPONENTIAL = 1, STVL_DECAY_PERSISTENT = 2 };

/* frustum model types — ModelType, stvl/include/.../measurement_reading.h:49-53 */
enum { STVL_MODEL_DEPTH_CAMERA = 0, STVL_MODEL_3D_LIDAR = 1, STVL_MODEL_INVALID = -1 };

/* cloud pre-filters — buffer::Filters, stvl/include/.../measurement_buffer.hpp (NONE/VOXEL/PASSTHROUGH),
 * applied upstream in MeasurementBuffer::BufferROSCloud (stvl/src/measurement_buffer.cpp:153-175). */
enum { STVL_FILTER_NONE = 0, STVL_FILTER_VOXEL = 1, STVL_FILTER_PASSTHROUGH = 2 };

/*
 * One clearing frustum, already transformed into the global frame on the host.
 * Built by stvl_build_depth_frustum / stvl_build_lidar_frustum (or by the façade's
 * geometry::DepthCameraFrustum / ThreeDimensionalLidarFrustum, which call them).
 *
 * type == STVL_MODEL_DEPTH_CAMERA: p[] = 6 planes in the reference's stored order
 *   top, left, bottom, right, far, near (stvl/src/frustum_models/depth_camera_frustum.cpp:109-149),
 *   each {nx, ny, nz, px, py, pz} = VectorWithPt3D{x,y,z,initial_point}
 *   (stvl/include/.../frustum_models/frustum.hpp:63-93) after TransformFrames.
 * type == STVL_MODEL_3D_LIDAR: p[0..2] = position, p[3..6] = conjugate orientation
 *   (x,y,z,w), p[7] = vFOV padding, p[8] = tan^2(vFOV/2), p[9] = min_d^2,
 *   p[10] = max_d^2, p[11] = hFOV/2, p[12] = full-circle flag (0.0/1.0)
 *   (stvl/src/frustum_models/three_dimensional_lidar_frustum.cpp:44-74).
 * type == STVL_MODEL_INVALID: never contains a point (vFOV==0 && hFOV==0,
 *   depth_camera_frustum.cpp:71-74,289-291).
 */
typedef struct stvl_frustum {
  int32_t type;
  int32_t reserved;
  double accel_factor; /* MeasurementReading::_decay_acceleration → frustum_model::accel_factor (hpp:105-119) */
  double p[36];
} stvl_frustum_t;

/* Grid construction parameters — SpatioTemporalVoxelGrid ctor (hpp:129-133, .cpp:49-62). */
typedef struct stvl_config {
  float voxel_size;        /* ctor takes `const float &`; stored as (double)(float) (.cpp:51,54) */
  int32_t decay_model;     /* STVL_DECAY_* */
  double background_value; /* OpenVDB background; carried for the façade, not used on device */
  double voxel_decay;
  int32_t pub_voxels;      /* build the occupancy point list each sweep (.cpp:234-240) */
  int32_t device;          /* CUDA device ordinal; -1 = current device */
  uint32_t leaf_capacity;  /* initial 8x8x8 leaf pool size (0 = default); grows on demand */
  uint32_t max_points;     /* initial point staging capacity per Mark call (0 = default); grows */
  /* spatial shard for the multi-GPU config: this handle only keeps
   * voxels whose leaf x-index bx satisfies floor_mod(floor_div(bx, shard_width), shard_count) == shard_index.
   * shard_count <= 1 disables sharding. */
  int32_t shard_index;
  int32_t shard_count;
  int32_t shard_width;     /* slab width in leaves (>=1) */
  int32_t reserved;
} stvl_config_t;

/* One marking reading — observation::MeasurementReading (measurement_reading.h:60-125)
 * flattened to what operator()(…) reads (.cpp:269-309), plus the optional
 * BufferROSCloud pre-filter (measurement_buffer.cpp:153-175). */
typedef struct stvl_reading {
  const void * data;       /* PointCloud2::data; HOST pointer (or device pointer for stvl_mark_device) */
  uint64_t n_points;       /* width*height */
  uint32_t point_step;     /* bytes per point */
  uint32_t off_x, off_y, off_z; /* byte offsets of the FLOAT32 fields named x,y,z */
  double origin[3];        /* _origin */
  double obstacle_range;   /* _obstacle_range_in_m */
  double now;              /* per-reading clock sample (.cpp:275) */
  int32_t marking;         /* _marking truthiness (.cpp:273) */
  int32_t filter;          /* STVL_FILTER_*; NONE = cloud was filtered upstream already; PASSTHROUGH / VOXEL run on the device */
  double min_obstacle_height, max_obstacle_height; /* z window of the filter */
  int32_t voxel_min_points; /* VOXEL filter only */
  int32_t has_transform;    /* != 0: the cloud is still in the sensor frame; apply `transform` first (then the filter, then Mark) */
  /* tf2::doTransform(cloud, cld_global, tf_stamped) of BufferROSCloud (stvl/src/measurement_buffer.cpp:141-146) on the
   * device: row-major 3x4 float affine [R | t] (tf2_sensor_msgs builds an Eigen::Transform<float,3,Affine> from the
   * TransformStamped and applies it to every x,y,z in FLOAT): out_i = ((m[i][0]*x + m[i][1]*y) + m[i][2]*z) + m[i][3]. */
  float transform[12];
} stvl_reading_t;

/* Result of one clear sweep. */
typedef struct stvl_sweep_stats {
  uint64_t n_swept;     /* active voxels visited */
  uint64_t n_survived;  /* voxels still active after the sweep */
  uint64_t n_cleared;   /* voxels expired by decay or frustum acceleration */
  uint64_t n_rewritten; /* survivors whose mark time was lowered (value - accel) */
  uint64_t n_ambiguous; /* decisions that depended on libm exp/atan within 1e-12 relative (listed, see stvl_get_ambiguous) */
  int32_t cleared_bbox_index[4]; /* min_ix, min_iy, max_ix, max_iy over cleared voxels (valid if n_cleared>0) */
  double cleared_bbox_world[4];  /* same as voxel-centre world coords: min_x, min_y, max_x, max_y */
} stvl_sweep_stats_t;

/* 2-D costmap target — the subset of nav2_costmap_2d::Costmap2D that
 * SpatioTemporalVoxelLayer::UpdateROSCostmap reads (stvl/src/spatio_temporal_voxel_layer.cpp:699-725). */
typedef struct stvl_costmap_params {
  double origin_x, origin_y, resolution;
  uint32_t size_x, size_y;
  int32_t mark_threshold;      /* _mark_threshold (:712) */
  uint8_t default_value;       /* default_value_ used by resetMaps (:705, :137-141) */
  uint8_t lethal_value;        /* nav2_costmap_2d::LETHAL_OBSTACLE = 254 */
  uint8_t pad[2];
} stvl_costmap_params_t;

typedef struct stvl_costmap_result {
  uint64_t n_marked_columns; /* columns that passed threshold AND worldToMap */
  int32_t has_bounds;        /* 0 if nothing was marked */
  int32_t reserved;
  double touched[4];         /* min_x, min_y, max_x, max_y of touch() over marked columns (world) */
} stvl_costmap_result_t;

typedef struct stvl_grid stvl_grid_t; /* opaque */

/* ---- lifecycle --------------------------------------------------------- */
int stvl_abi_version(void);
const char * stvl_last_error(void); /* thread-local message of the last failing call */
/* number of CUDA devices usable by this library (compute capability 10.x); host-only call */
int stvl_device_count(int * count);

/* SpatioTemporalVoxelGrid::SpatioTemporalVoxelGrid + InitializeGrid (.cpp:49-93) */
int stvl_create(const stvl_config_t * cfg, stvl_grid_t ** out);
/* ~SpatioTemporalVoxelGrid (.cpp:65-71) */
int stvl_destroy(stvl_grid_t * g);

/* ---- host-side frustum geometry (no GPU needed) ------------------------- */
/* DepthCameraFrustum ctor + ComputePlaneNormals + SetPosition/SetOrientation + TransformModel
 * (stvl/src/frustum_models/depth_camera_frustum.cpp:45-174).  quat = (x,y,z,w). */
int stvl_build_depth_frustum(double vfov, double hfov, double min_dist, double max_dist,
                             const double origin[3], const double quat_xyzw[4],
                             double accel_factor, stvl_frustum_t * out);
/* ThreeDimensionalLidarFrustum ctor + SetPosition/SetOrientation + TransformModel
 * (stvl/src/frustum_models/three_dimensional_lidar_frustum.cpp:44-74). */
int stvl_build_lidar_frustum(double vfov, double vfov_padding, double hfov,
                             double min_dist, double max_dist,
                             const double origin[3], const double quat_xyzw[4],
                             double accel_factor, stvl_frustum_t * out);

/* The float affine tf2::doTransform applies to a cloud for a TransformStamped (translation, rotation quaternion x,y,z,w)
 * — stvl/src/measurement_buffer.cpp:141-146; tf2_sensor_msgs builds Translation3f * Quaternionf (Eigen, float).
 * Fills the row-major 3x4 `transform` of stvl_reading_t. */
int stvl_build_transform(const double translation[3], const double quat_xyzw[4], float transform[12]);

/* ---- the per-cycle hot path --------------------------------------------- */
/* ClearFrustums + TemporalClearAndGenerateCostmap (.cpp:96-224): one sweep over the
 * active voxels with ONE `now`; frustums in reading order, first containing one wins.
 * Asynchronous: results are fetched with stvl_get_sweep_stats / stvl_costmap / … */
int stvl_clear_frustums(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums);
/* Mark + operator() (.cpp:254-309).  Host clouds are copied to the device inside the call. */
int stvl_mark(stvl_grid_t * g, const stvl_reading_t * readings, uint32_t n_readings);
/* Same, but readings[i].data are DEVICE pointers already resident in HBM (16-byte aligned).
 * LIFETIME: the call only enqueues work, and a Mark that runs out of leaf / tile pool is replayed after the pool has
 * grown — detected at the next stvl_* call on the handle.  The device clouds handed to stvl_mark_device,
 * stvl_update_cycle_device and stvl_update_cycle_sharded_device must therefore stay valid and UNMODIFIED until the next
 * stvl_* call on the handle has returned (host clouds passed to stvl_mark are copied inside the call and are the
 * caller's again when it returns). */
int stvl_mark_device(stvl_grid_t * g, const stvl_reading_t * readings, uint32_t n_readings);
/* UpdateROSCostmap's grid loop (stvl/src/spatio_temporal_voxel_layer.cpp:699-718) fused with
 * GetFlattenedCostmap: resetMaps to default_value, then lethal_value for every voxel column with
 * (int)count >= mark_threshold that worldToMap accepts.  Uses the column counts of the LAST sweep
 * (pre-Mark state).  `costmap` is a HOST buffer of size_x*size_y bytes (may be NULL
 * to skip the copy-out).  Synchronises. */
int stvl_costmap(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap, stvl_costmap_result_t * res);
/* Same, but leaves the bytes in a DEVICE buffer the caller owns.  With res == NULL the call only enqueues work
 * (no synchronisation); results are then available after stvl_synchronize / any synchronising call. */
int stvl_costmap_device(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap_dev, stvl_costmap_result_t * res);

/* One whole update cycle with device-resident inputs, in the reference's order (stvl/src/spatio_temporal_voxel_layer.cpp:
 * 772-795): the results are those of stvl_clear_frustums, stvl_mark_device, stvl_costmap_device(res = NULL), but the
 * three stages are fused into two launches (the sweep kernel also resets the costmap bytes, the costmap pass rides on
 * dedicated warps of the Mark kernel, the kernels overlap by programmatic dependent launch).  Enqueue only, no synchronisation;
 * the clouds must be complete on the device when the call is made (stream order on stvl_get_stream() suffices).
 * stvl_get_costmap_result returns the marked-cell count / bounds afterwards. */
int stvl_update_cycle_device(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                             const stvl_reading_t * readings, uint32_t n_readings,
                             const stvl_costmap_params_t * p, uint8_t * costmap_dev);
/* The same cycle on a spatial shard (stvl_config_t.shard_*, stvl_comm_init).  x-slab shards own whole voxel columns, so
 * the merged costmap is the union of the ranks' lethal cells: the fused cycle's costmap pass writes a packed BITMAP of the
 * cells of this rank's own x interval (1 bit per cell), the bitmaps travel to `root` with one grouped ncclSend / ncclRecv,
 * and the root expands them into the bytes of `costmap_dev` (lethal_value where any rank marked the cell, default_value
 * elsewhere).  Blocking (pipelined = 0: the bytes are this cycle's when the call's work has run), or pipelined one cycle
 * behind: exchange and expansion run on a side stream while the next cycle's kernels run, the bytes of cycle k are in
 * `costmap_dev` (in stream order on stvl_get_stream()) after call k+1, and stvl_costmap_sharded_flush drains the last
 * one.  costmap_dev is only used on the root.  Needs shard_count == the communicator's size. */
int stvl_update_cycle_sharded_device(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                                     const stvl_reading_t * readings, uint32_t n_readings,
                                     const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root, int pipelined);
/* The geometry of that exchange (host arithmetic only, no GPU): for every rank r of an n_ranks-way x-slab sharding
 * (rank r keeps the leaves with floor(bx / shard_width) mod shard_count == r), the interval [w_lo[r], w_hi[r]) of 32-cell
 * words of a costmap row that its voxel columns can map into.  Rank r's packed bitmap has size_y rows of
 * (w_hi[r] - w_lo[r]) words; bit (mx & 31) of word (mx >> 5) - w_lo[r] of row my is cell (mx, my). */
int stvl_shard_word_intervals(float voxel_size, int shard_count, int shard_width, int n_ranks, const stvl_costmap_params_t * p,
                              int32_t * w_lo, int32_t * w_hi);
/* SpatioTemporalVoxelLayer::updateCosts (stvl/src/spatio_temporal_voxel_layer.cpp:666-696) on DEVICE byte grids of
 * p->size_x * p->size_y cells: if n_footprint >= 3, the footprint polygon (world x,y pairs, convex) is set to
 * FREE_SPACE (0) in the layer's costmap — nav2 Costmap2D::setConvexPolygonCost(_transformed_footprint, FREE_SPACE), :682-684
 * (nothing happens if a vertex is off the map; at most 64 vertices) — then, if master_dev != NULL, the window
 * [min_i, max_i) x [min_j, max_j) of the layer's costmap is combined into the master grid: combination_method 0 =
 * CostmapLayer::updateWithOverwrite, 1 = updateWithMax (:686-695), anything else leaves the master untouched.
 * Enqueue only (ordered on stvl_get_stream()). */
int stvl_update_costs_device(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * layer_costmap_dev, uint8_t * master_dev,
                             const double * footprint_xy, uint32_t n_footprint, int combination_method,
                             int32_t min_i, int32_t min_j, int32_t max_i, int32_t max_j);

/* Pipelined form of the whole cycle for HOST buffers (what a costmap plugin that keeps its clouds and its costmap in host
 * memory calls): readings[i].data are HOST clouds (pinned memory makes the copies asynchronous), `costmap_host` receives
 * the size_x*size_y bytes.  The call enqueues H2D copy -> fused cycle -> D2H copy on three streams with double-buffered
 * device staging and returns a ticket at once; cycle k+1's H2D copy overlaps cycle k's kernels and read-back.
 * The clouds of a cycle and its costmap_host belong to the library until stvl_update_cycle_wait(ticket) has returned
 * (which also delivers the marked-cell count / bounds of that cycle).  At most two cycles may be in flight: submitting a
 * third waits for the oldest.  Cycles complete in submission order.
 * On a handle with a communicator (a shard of a multi-GPU map, stvl_comm_init) the cycle is the sharded one: every rank
 * uploads its own clouds, the ranks' cells are merged into rank 0 (blocking merge on the cycle's stream) and only rank 0
 * reads bytes back into its costmap_host (the merged map); the result counts are each rank's own. */
int stvl_update_cycle_async(stvl_grid_t * g, double now, const stvl_frustum_t * frustums, uint32_t n_frustums,
                            const stvl_reading_t * readings, uint32_t n_readings,
                            const stvl_costmap_params_t * p, uint8_t * costmap_host, uint64_t * ticket);
int stvl_update_cycle_wait(stvl_grid_t * g, uint64_t ticket, stvl_costmap_result_t * res);

/* The result (marked columns, touched bounds — the touch() calls of spatio_temporal_voxel_layer.cpp:716) of the last
 * costmap pass enqueued by stvl_costmap_device(res = NULL) or stvl_update_cycle_device.  Synchronises. */
int stvl_get_costmap_result(stvl_grid_t * g, stvl_costmap_result_t * res);

/* ---- results of the last sweep ------------------------------------------ */
/* Which optional per-sweep lists the sweep kernel builds (default: STVL_OUT_CLEARED | STVL_OUT_AMBIGUOUS;
 * the occupancy point list follows cfg.pub_voxels).  Counts and the cleared bounding box are always produced. */
enum { STVL_OUT_CLEARED = 1, STVL_OUT_AMBIGUOUS = 2 };
int stvl_set_outputs(stvl_grid_t * g, uint32_t flags);
int stvl_get_sweep_stats(stvl_grid_t * g, stvl_sweep_stats_t * out); /* synchronises */
/* GetFlattenedCostmap (.cpp:312-317): occupied voxel columns of the last sweep as index pairs and
 * counts; world (x,y) of a column = voxel centre, see stvl_index_to_world.  Any of the arrays may be NULL. */
int stvl_get_columns(stvl_grid_t * g, int32_t * ix, int32_t * iy, uint32_t * count, uint64_t cap, uint64_t * n);
/* cleared_cells of the last sweep (.cpp:213-215), one entry per cleared VOXEL (duplicates per column
 * possible; the reference's unordered_set dedupes on insert). */
int stvl_get_cleared(stvl_grid_t * g, int32_t * ix, int32_t * iy, uint64_t cap, uint64_t * n);
/* GetOccupancyPointCloud (.cpp:344-376): survivors' voxel centres as FLOAT32 xyz triples; only if pub_voxels. */
int stvl_get_points(stvl_grid_t * g, float * xyz, uint64_t cap, uint64_t * n);
/* voxels whose decision hinged on a libm exp/atan result within the listing band (north_star: "must be listed") */
int stvl_get_ambiguous(stvl_grid_t * g, int32_t * ix, int32_t * iy, int32_t * iz, uint64_t cap, uint64_t * n);

/* ---- grid state ---------------------------------------------------------- */
/* ResetGrid (.cpp:379-394) */
int stvl_reset(stvl_grid_t * g);
/* ResetGridArea (.cpp:397-418): clear voxels whose centre is (strictly) inside [start,end] — or outside, if invert_area */
int stvl_reset_area(stvl_grid_t * g, double start_x, double start_y, double end_x, double end_y, int invert_area);
/* number of active voxels (IsGridEmpty == (count == 0), .cpp:472-477); synchronises */
int stvl_active_count(stvl_grid_t * g, uint64_t * n);
/* dump every active voxel (index triple + mark time) — parity tests and SaveGrid-style export */
int stvl_get_voxels(stvl_grid_t * g, int32_t * ix, int32_t * iy, int32_t * iz, double * t, uint64_t cap, uint64_t * n);
/* bulk-load voxels (index triple + mark time), overwriting like MarkGridPoint (.cpp:421-430); HOST arrays */
int stvl_load_voxels(stvl_grid_t * g, const int32_t * ix, const int32_t * iy, const int32_t * iz, const double * t, uint64_t n);

/* IndexToWorld (.cpp:446-460): coord*vs then + vs/2 (two roundings). Host-only helper. */
double stvl_index_to_world(const stvl_grid_t * g, int32_t index);
/* stored voxel size (double)(float)voxel_size */
double stvl_voxel_size(const stvl_grid_t * g);

/* ---- multi-GPU merge of the column counts ---------------- */
/* Scatter-add the last sweep's column counts into a dense DEVICE grid of nx*ny uint32 covering voxel
 * columns [ix0, ix0+nx) x [iy0, iy0+ny) (row-major, y slow).  The caller zeroes / reduces the grid
 * (ncclReduce over NVLink) between ranks.  Asynchronous on the handle's stream (stvl_get_stream): enqueue the
 * collective in stream order after it. */
int stvl_columns_to_dense(stvl_grid_t * g, int32_t ix0, int32_t iy0, uint32_t nx, uint32_t ny, uint32_t * dense_dev);
/* Threshold a (reduced) dense column-count grid into costmap bytes, same rule as stvl_costmap.  res == NULL: enqueue only. */
int stvl_costmap_from_dense(stvl_grid_t * g, const uint32_t * dense_dev, int32_t ix0, int32_t iy0, uint32_t nx, uint32_t ny,
                            const stvl_costmap_params_t * p, uint8_t * costmap_dev, stvl_costmap_result_t * res);

/* In-library variant of the same merge (what the C++ host side uses): NCCL is loaded at run time (dlopen of
 * libnccl.so.2 — the copy already in the process when one is, e.g. torch's) and the reduce runs on a side stream.
 *   stvl_nccl_unique_id : rank 0 creates the id; the caller ships its 128 bytes to the other ranks (any channel)
 *   stvl_comm_init      : collective; one communicator per handle
 *   stvl_costmap_sharded_device : scatter this rank's column counts into a window over the costmap (count_bytes 1 or 4
 *       per cell; 1 saturates at 255 and is exact for disjoint shards / thresholds <= 255), ncclReduce(sum) to `root`,
 *       root thresholds into costmap_dev (same rule as stvl_costmap).  count_bytes == 0 ("bytes mode", for shards that
 *       own whole voxel columns such as x-slabs): every rank thresholds its own columns into {0, lethal} bytes and
 *       ncclReduce(max) assembles them in the root's window, which is then copied to costmap_dev; no 4 B/cell count window
 *       is moved and the root does not re-scan one.  pipelined == 0: everything of THIS cycle, in stream
 *       order.  pipelined == 1: the reduce of this call overlaps the next cycle's sweep and Mark (the column tiles are
 *       double-buffered), and root writes the bytes of the PREVIOUS call; stvl_costmap_sharded_flush writes the last ones.
 * All enqueue-only (no host synchronisation). */
int stvl_nccl_unique_id(uint8_t id[128]);
int stvl_comm_init(stvl_grid_t * g, const uint8_t id[128], int rank, int n_ranks);
int stvl_comm_destroy(stvl_grid_t * g);
int stvl_costmap_sharded_device(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root,
                                int count_bytes, int pipelined);
int stvl_costmap_sharded_flush(stvl_grid_t * g, const stvl_costmap_params_t * p, uint8_t * costmap_dev, int root);

/* ---- diagnostics ---------------------------------------------------------- */
typedef struct stvl_pool_stats {
  uint64_t leaf_capacity, leaves_in_use, leaves_free, hash_capacity, hash_tombstones;
  uint64_t tiles_in_use, tile_capacity, points_dropped_out_of_range, points_dropped_nonfinite;
  uint64_t kernel_launches; /* kernels launched by this handle since creation */
  uint64_t device_bytes;    /* device memory held */
} stvl_pool_stats_t;
int stvl_get_pool_stats(stvl_grid_t * g, stvl_pool_stats_t * out);
/* Per-stage device timing: when enabled, the library brackets the selected stages' kernels with CUDA events on its
 * stream.  stvl_get_timing synchronises, adds up the finished intervals since the last call and resets. */
enum { STVL_TIME_SWEEP = 1, STVL_TIME_MARK = 2, STVL_TIME_COSTMAP = 4 };
typedef struct stvl_timing {
  double sweep_ms, mark_ms, costmap_ms;       /* summed kernel time per stage */
  uint64_t n_sweep, n_mark, n_costmap;        /* intervals summed */
} stvl_timing_t;
int stvl_enable_timing(stvl_grid_t * g, uint32_t stage_mask);
int stvl_get_timing(stvl_grid_t * g, stvl_timing_t * out);
/* raw CUDA stream (cudaStream_t) the handle launches on — for CUDA-event timing by the caller */
int stvl_get_stream(stvl_grid_t * g, void ** stream);
/* block until all queued work of the handle is done */
int stvl_synchronize(stvl_grid_t * g);

#ifdef __cplusplus
}
#endif
#endif /* STVL_B200_H_ */
