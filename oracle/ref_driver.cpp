// ref_driver.cpp — C entry points around the REFERENCE'S OWN translation units (TEST INFRASTRUCTURE, never shipped).
//
// oracle/Makefile compiles, unmodified and from where they lie under /root/reference,
//     spatio_temporal_voxel_layer/src/spatio_temporal_voxel_grid.cpp
//     spatio_temporal_voxel_layer/src/frustum_models/depth_camera_frustum.cpp
//     spatio_temporal_voxel_layer/src/frustum_models/three_dimensional_lidar_frustum.cpp
// against the stand-in headers in oracle/ref_shim/ (OpenVDB, Eigen, PCL, ROS 2 and Boost are not in this image and cannot
// be fetched) and links them with this file into oracle/_ref/libstvl_ref.so.  The reference's own logic — reading order,
// first-frustum-wins, strict comparisons, the quantisation quirks, the decay algebra, the plane construction — therefore
// runs as written; what the stand-ins restate is the third-party arithmetic only (see the headers).  tests/test_ref_pin.py
// checks the restated oracle (oracle/stvl_oracle.cpp) against this library bit for bit, and tests/golden/ref_*.npz holds
// outputs generated with it (the library itself cannot travel to the GPU box: /root/reference does not exist there).
#include <cstdint>
#include <cstring>
#include <memory>
#include <unordered_set>
#include <vector>

// the plane set of a DepthCameraFrustum is private; the driver dumps it to pin ComputePlaneNormals / TransformModel
#define private public
#define protected public
#include "spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp"
#undef private
#undef protected

namespace
{
struct RefGrid
{
  double now = 0.0;
  rclcpp::Clock::SharedPtr clock;
  std::unique_ptr<volume_grid::SpatioTemporalVoxelGrid> grid;
  std::unordered_set<volume_grid::occupany_cell> cleared;
};

observation::MeasurementReading make_reading(int model_type, const double * p)
{
  // p: vfov, vfov_padding, hfov, min_z, max_z, ox, oy, oz, qx, qy, qz, qw, accel
  observation::MeasurementReading r;
  r._vertical_fov_in_rad = p[0]; r._vertical_fov_padding_in_m = p[1]; r._horizontal_fov_in_rad = p[2];
  r._min_z_in_m = p[3]; r._max_z_in_m = p[4];
  r._origin.x = p[5]; r._origin.y = p[6]; r._origin.z = p[7];
  r._orientation.x = p[8]; r._orientation.y = p[9]; r._orientation.z = p[10]; r._orientation.w = p[11];
  r._decay_acceleration = p[12];
  r._marking = 0.; r._clearing = 1.;
  r._obstacle_range_in_m = 0.;
  r._model_type = (ModelType)model_type;
  return r;
}
}  // namespace

extern "C" {

void * ref_create(float voxel_size, double background, int decay_model, double voxel_decay, int pub_voxels)
{
  RefGrid * g = new RefGrid;
  g->clock = std::make_shared<rclcpp::Clock>([g]() { return g->now; });
  g->grid.reset(new volume_grid::SpatioTemporalVoxelGrid(g->clock, voxel_size, background, decay_model, voxel_decay, pub_voxels != 0));
  return g;
}
void ref_destroy(void * h) { delete static_cast<RefGrid *>(h); }

// Mark(): one reading; `data` is a PointCloud2 payload of n points with the given step and x/y/z offsets
void ref_mark(void * h, const uint8_t * data, uint64_t n, uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
              const double origin[3], double obstacle_range, double marking, double now)
{
  RefGrid * g = static_cast<RefGrid *>(h);
  g->now = now;
  sensor_msgs::msg::PointCloud2 cloud;
  cloud.height = 1; cloud.width = (uint32_t)n; cloud.point_step = point_step; cloud.row_step = point_step * (uint32_t)n;
  const char * names[3] = {"x", "y", "z"};
  const uint32_t offs[3] = {off_x, off_y, off_z};
  for (int i = 0; i < 3; ++i) { sensor_msgs::msg::PointField f; f.name = names[i]; f.offset = offs[i]; f.datatype = sensor_msgs::msg::PointField::FLOAT32; f.count = 1; cloud.fields.push_back(f); }
  cloud.data.assign(data, data + (size_t)n * point_step);
  observation::MeasurementReading r(cloud, obstacle_range);
  r._origin.x = origin[0]; r._origin.y = origin[1]; r._origin.z = origin[2];
  r._marking = marking; r._clearing = 0.;
  std::vector<observation::MeasurementReading> v;
  v.push_back(r);
  g->grid->Mark(v);
}

// ClearFrustums(): n clearing readings (model_type[i], params + 13 * i)
void ref_clear_frustums(void * h, double now, uint32_t n, const int32_t * model_type, const double * params)
{
  RefGrid * g = static_cast<RefGrid *>(h);
  g->now = now;
  std::vector<observation::MeasurementReading> v;
  v.reserve(n);
  for (uint32_t i = 0; i < n; ++i) v.push_back(make_reading(model_type[i], params + 13 * i));
  g->cleared.clear();
  g->grid->ClearFrustums(v, g->cleared);
}

uint64_t ref_active_count(void * h)
{
  RefGrid * g = static_cast<RefGrid *>(h);
  uint64_t n = 0;
  for (auto it = g->grid->_grid->cbeginValueOn(); it.test(); ++it) ++n;
  return n;
}
void ref_get_voxels(void * h, int32_t * ix, int32_t * iy, int32_t * iz, double * t, uint64_t cap)
{
  RefGrid * g = static_cast<RefGrid *>(h);
  uint64_t n = 0;
  for (auto it = g->grid->_grid->cbeginValueOn(); it.test() && n < cap; ++it, ++n) {
    const openvdb::Coord c = it.getCoord();
    ix[n] = c.x(); iy[n] = c.y(); iz[n] = c.z(); t[n] = it.getValue();
  }
}
uint64_t ref_costmap_size(void * h) { return static_cast<RefGrid *>(h)->grid->GetFlattenedCostmap()->size(); }
void ref_get_costmap(void * h, double * x, double * y, uint32_t * c, uint64_t cap)
{
  uint64_t n = 0;
  for (const auto & kv : *static_cast<RefGrid *>(h)->grid->GetFlattenedCostmap()) {
    if (n >= cap) break;
    x[n] = kv.first.x; y[n] = kv.first.y; c[n] = kv.second; ++n;
  }
}
uint64_t ref_get_cleared(void * h, double * x, double * y, uint64_t cap)
{
  RefGrid * g = static_cast<RefGrid *>(h);
  uint64_t n = 0;
  for (const auto & c : g->cleared) { if (x && n < cap) { x[n] = c.x; y[n] = c.y; } ++n; }
  return n;
}
uint64_t ref_get_points(void * h, float * xyz, uint64_t cap)
{
  RefGrid * g = static_cast<RefGrid *>(h);
  std::unique_ptr<sensor_msgs::msg::PointCloud2> pc2(new sensor_msgs::msg::PointCloud2);
  g->grid->GetOccupancyPointCloud(pc2);
  const uint64_t n = (uint64_t)pc2->width * pc2->height;
  if (xyz) {
    sensor_msgs::PointCloud2ConstIterator<float> ix(*pc2, "x"), iy(*pc2, "y"), iz(*pc2, "z");
    for (uint64_t i = 0; i < n && i < cap; ++i, ++ix, ++iy, ++iz) { xyz[3 * i] = *ix; xyz[3 * i + 1] = *iy; xyz[3 * i + 2] = *iz; }
  }
  return n;
}
int ref_reset(void * h) { return static_cast<RefGrid *>(h)->grid->ResetGrid() ? 1 : 0; }
void ref_reset_area(void * h, double sx, double sy, double ex, double ey, int invert)
{
  static_cast<RefGrid *>(h)->grid->ResetGridArea(volume_grid::occupany_cell(sx, sy), volume_grid::occupany_cell(ex, ey), invert != 0);
}
// the grid's own conversions and decay algebra (protected members)
void ref_index_to_world(void * h, int32_t ix, int32_t iy, int32_t iz, double out[3])
{
  const openvdb::Vec3d w = static_cast<RefGrid *>(h)->grid->IndexToWorld(openvdb::Coord(ix, iy, iz));
  out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
}
double ref_decay_duration(void * h, double t) { return static_cast<RefGrid *>(h)->grid->GetTemporalClearingDuration(t); }
double ref_frustum_acceleration(void * h, double t, double factor) { return static_cast<RefGrid *>(h)->grid->GetFrustumAcceleration(t, factor); }

// frustum models on their own: build one (ctor + SetPosition + SetOrientation + TransformModel, as ClearFrustums does) and
// query IsInside for n points; for a depth camera also dump the six transformed planes {normal xyz, point xyz}
static geometry::Frustum * build_frustum(int model_type, const double * p)
{
  geometry::Frustum * f = nullptr;
  if (model_type == DEPTH_CAMERA) f = new geometry::DepthCameraFrustum(p[0], p[2], p[3], p[4]);
  else f = new geometry::ThreeDimensionalLidarFrustum(p[0], p[1], p[2], p[3], p[4]);
  geometry_msgs::msg::Point o; o.x = p[5]; o.y = p[6]; o.z = p[7];
  geometry_msgs::msg::Quaternion q; q.x = p[8]; q.y = p[9]; q.z = p[10]; q.w = p[11];
  f->SetPosition(o);
  f->SetOrientation(q);
  f->TransformModel();
  return f;
}
void ref_frustum_is_inside(int model_type, const double * params, const double * xyz, uint64_t n, uint8_t * out)
{
  geometry::Frustum * f = build_frustum(model_type, params);
  for (uint64_t i = 0; i < n; ++i) out[i] = f->IsInside(openvdb::Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])) ? 1 : 0;
  delete f;
}
int ref_depth_frustum_planes(const double * params, double out36[36])
{
  geometry::DepthCameraFrustum * f = static_cast<geometry::DepthCameraFrustum *>(build_frustum(DEPTH_CAMERA, params));
  const int valid = f->_valid_frustum ? 1 : 0;
  if (valid) {
    for (int k = 0; k < 6; ++k) {
      const geometry::VectorWithPt3D & pl = f->_plane_normals[k];
      out36[6 * k] = pl.x; out36[6 * k + 1] = pl.y; out36[6 * k + 2] = pl.z;
      out36[6 * k + 3] = pl.initial_point[0]; out36[6 * k + 4] = pl.initial_point[1]; out36[6 * k + 5] = pl.initial_point[2];
    }
  }
  delete f;
  return valid;
}

}  // extern "C"
