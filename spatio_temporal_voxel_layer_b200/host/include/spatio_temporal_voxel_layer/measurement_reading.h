// observation::MeasurementReading — the input record of Mark / ClearFrustums.  Field names, types and meaning follow
// the reference (spatio_temporal_voxel_layer/include/spatio_temporal_voxel_layer/measurement_reading.h:60-125) so
// that callers written against it compile unchanged; note _marking / _clearing are doubles tested for truthiness
// upstream (measurement_reading.h:123, spatio_temporal_voxel_grid.cpp:273).
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__MEASUREMENT_READING_H_
#define SPATIO_TEMPORAL_VOXEL_LAYER__MEASUREMENT_READING_H_

#include <memory>

#include "geometry_msgs/msg/point.hpp"
#include "geometry_msgs/msg/quaternion.hpp"
#include "sensor_msgs/msg/point_cloud2.hpp"

enum ModelType
{
  DEPTH_CAMERA = 0,
  THREE_DIMENSIONAL_LIDAR = 1
};

namespace observation
{
struct MeasurementReading
{
  MeasurementReading()
  : _cloud(std::make_shared<sensor_msgs::msg::PointCloud2>()) {}

  MeasurementReading(
    geometry_msgs::msg::Point & origin, sensor_msgs::msg::PointCloud2 cloud, double obstacle_range, double min_z,
    double max_z, double vFOV, double vFOVPadding, double hFOV, double decay_acceleration, bool marking, bool clearing,
    ModelType model_type)
  : _origin(origin), _cloud(std::make_shared<sensor_msgs::msg::PointCloud2>(std::move(cloud))),
    _obstacle_range_in_m(obstacle_range), _min_z_in_m(min_z), _max_z_in_m(max_z), _vertical_fov_in_rad(vFOV),
    _vertical_fov_padding_in_m(vFOVPadding), _horizontal_fov_in_rad(hFOV), _marking(marking), _clearing(clearing),
    _decay_acceleration(decay_acceleration), _model_type(model_type) {}

  MeasurementReading(sensor_msgs::msg::PointCloud2 cloud, double obstacle_range)
  : _cloud(std::make_shared<sensor_msgs::msg::PointCloud2>(std::move(cloud))), _obstacle_range_in_m(obstacle_range) {}

  // The reference deep-copies the cloud on every copy (measurement_reading.h:100-116), i.e. 34 MB per cycle for seven
  // depth cameras before the hot path even starts.  The cloud is immutable once buffered, so copies share it here.
  MeasurementReading(const MeasurementReading &) = default;
  MeasurementReading & operator=(const MeasurementReading &) = default;

  geometry_msgs::msg::Point _origin;
  geometry_msgs::msg::Quaternion _orientation;
  std::shared_ptr<sensor_msgs::msg::PointCloud2> _cloud;
  double _obstacle_range_in_m = 0, _min_z_in_m = 0, _max_z_in_m = 0;
  double _vertical_fov_in_rad = 0, _vertical_fov_padding_in_m = 0, _horizontal_fov_in_rad = 0;
  double _marking = 0, _clearing = 0, _decay_acceleration = 0;
  ModelType _model_type = DEPTH_CAMERA;

  // ---- not in the reference struct: the BufferROSCloud pre-filter (measurement_buffer.cpp:153-175) deferred to the
  // device.  0 = NONE (cloud already filtered, the reference's situation), 1 = VOXEL, 2 = PASSTHROUGH.
  int _device_filter = 0;
  double _min_obstacle_height = 0.0, _max_obstacle_height = 0.0;
  int _voxel_min_points = 0;
  // the cloud is still in the sensor frame: tf2::doTransform (measurement_buffer.cpp:141-146) deferred to the device as
  // well — row-major 3x4 float affine (stvl_build_transform)
  bool _has_device_transform = false;
  float _device_transform[12] = {0};
};
}  // namespace observation
#endif  // SPATIO_TEMPORAL_VOXEL_LAYER__MEASUREMENT_READING_H_
