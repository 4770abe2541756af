// buffer::MeasurementBuffer — per-source queue of readings between the sensor callbacks and the update cycle.
// The reference's buffer (measurement_buffer.hpp:84-170, measurement_buffer.cpp) also does the TF lookups and the PCL
// VoxelGrid / PassThrough filtering on the subscriber thread; TF / PCL / ROS message plumbing is outside the hot path
// and outside this build (SURVEY.md §8 scope), so this buffer takes readings that are already in the global frame and
// keeps the part of the contract the update cycle relies on: GetReadings, observation persistence, ClearAfterReading,
// Lock / Unlock, enable flag.  The z-window / voxel filter parameters travel with the buffer so the Mark kernel can
// apply them on the device.
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__MEASUREMENT_BUFFER_HPP_
#define SPATIO_TEMPORAL_VOXEL_LAYER__MEASUREMENT_BUFFER_HPP_

#include <list>
#include <mutex>
#include <string>
#include <vector>

#include "rclcpp/rclcpp.hpp"
#include "spatio_temporal_voxel_layer/measurement_reading.h"
#include "stvl_b200.h"

namespace buffer
{
enum class Filters
{
  NONE = 0,
  VOXEL = 1,
  PASSTHROUGH = 2
};

class MeasurementBuffer
{
public:
  MeasurementBuffer(
    const std::string & source_name, rclcpp::Clock::SharedPtr clock, const double & observation_keep_time,
    const double & expected_update_rate, const double & min_obstacle_height, const double & max_obstacle_height,
    const double & obstacle_range, const double & min_z, const double & max_z, const double & vFOV,
    const double & vFOVPadding, const double & hFOV, const double & decay_acceleration, const bool & marking,
    const bool & clearing, const Filters & filter, const int & voxel_min_points, const bool & clear_buffer_after_reading,
    const ModelType & model_type)
  : _source_name(source_name), _clock(clock), _observation_keep_time(observation_keep_time),
    _expected_update_rate(expected_update_rate), _min_obstacle_height(min_obstacle_height),
    _max_obstacle_height(max_obstacle_height), _obstacle_range(obstacle_range), _min_z(min_z), _max_z(max_z),
    _vertical_fov(vFOV), _vertical_fov_padding(vFOVPadding), _horizontal_fov(hFOV),
    _decay_acceleration(decay_acceleration), _marking(marking), _clearing(clearing), _filter(filter),
    _voxel_min_points(voxel_min_points), _clear_buffer_after_reading(clear_buffer_after_reading), _enabled(true),
    _model_type(model_type), _last_updated(clock->now().seconds()) {}

  // BufferROSCloud (measurement_buffer.cpp:90-190) minus TF and PCL: the sensor pose and the cloud arrive in the
  // global frame; the per-source parameters are stamped on the reading exactly as upstream (:123-133)
  void BufferCloud(
    const geometry_msgs::msg::Point & origin, const geometry_msgs::msg::Quaternion & orientation,
    std::shared_ptr<sensor_msgs::msg::PointCloud2> cloud)
  {
    BufferCloudImpl(origin, orientation, cloud, false);
  }

  // The same with the cloud still in the SENSOR frame: (translation, rotation) is the sensor -> global TransformStamped
  // the reference looks up at :101-116/:141-145.  The sensor origin (0,0,0 in its own frame) maps to the translation,
  // its orientation to the rotation; the cloud's doTransform (:146) runs on the device inside Mark.
  void BufferSensorFrameCloud(
    const geometry_msgs::msg::Point & translation, const geometry_msgs::msg::Quaternion & rotation,
    std::shared_ptr<sensor_msgs::msg::PointCloud2> cloud)
  {
    BufferCloudImpl(translation, rotation, cloud, true);
  }

private:
  void BufferCloudImpl(
    const geometry_msgs::msg::Point & origin, const geometry_msgs::msg::Quaternion & orientation,
    std::shared_ptr<sensor_msgs::msg::PointCloud2> cloud, bool sensor_frame)
  {
    observation::MeasurementReading r;
    if (sensor_frame) {
      const double t[3] = {origin.x, origin.y, origin.z};
      const double q[4] = {orientation.x, orientation.y, orientation.z, orientation.w};
      r._has_device_transform = stvl_build_transform(t, q, r._device_transform) == 0;
    }
    r._origin = origin;
    r._orientation = orientation;
    r._obstacle_range_in_m = _obstacle_range;
    r._min_z_in_m = _min_z;
    r._max_z_in_m = _max_z;
    r._vertical_fov_in_rad = _vertical_fov;
    r._vertical_fov_padding_in_m = _vertical_fov_padding;
    r._horizontal_fov_in_rad = _horizontal_fov;
    r._decay_acceleration = _decay_acceleration;
    r._clearing = _clearing;
    r._marking = _marking;
    r._model_type = _model_type;
    // the z-window filter the reference runs here with PCL is applied by the Mark kernel on the device
    r._device_filter = static_cast<int>(_filter);
    r._min_obstacle_height = _min_obstacle_height;
    r._max_obstacle_height = _max_obstacle_height;
    r._voxel_min_points = _voxel_min_points;
    if (!(_clearing && !_marking) && cloud) {   // clearing-only sources do not buffer points (:135-138)
      r._cloud = cloud;
    }
    _observation_list.push_front(r);
    _stamps.push_front(_clock->now().seconds());
    _last_updated = _clock->now().seconds();
    RemoveStaleObservations();
  }

public:
  // GetReadings (measurement_buffer.cpp:193-204)
  void GetReadings(std::vector<observation::MeasurementReading> & observations)
  {
    RemoveStaleObservations();
    for (auto it = _observation_list.begin(); it != _observation_list.end(); ++it) {
      observations.push_back(*it);
    }
  }
  // RemoveStaleObservations (measurement_buffer.cpp:207-228): keep_time 0 keeps only the newest reading
  void RemoveStaleObservations(void)
  {
    if (_observation_list.empty()) {return;}
    if (_observation_keep_time == 0.) {
      _observation_list.erase(++_observation_list.begin(), _observation_list.end());
      _stamps.erase(++_stamps.begin(), _stamps.end());
      return;
    }
    const double now = _clock->now().seconds();
    auto it = _observation_list.begin();
    auto st = _stamps.begin();
    for (; it != _observation_list.end(); ++it, ++st) {
      if (now - *st > _observation_keep_time) {
        _observation_list.erase(it, _observation_list.end());
        _stamps.erase(st, _stamps.end());
        return;
      }
    }
  }
  void ResetAllMeasurements(void) {_observation_list.clear(); _stamps.clear();}
  bool ClearAfterReading(void) {return _clear_buffer_after_reading;}
  bool UpdatedAtExpectedRate(void) const
  {
    if (_expected_update_rate == 0.) {return true;}
    return (_clock->now().seconds() - _last_updated) <= _expected_update_rate;
  }
  bool IsEnabled(void) const {return _enabled;}
  void SetEnabled(const bool & enabled) {_enabled = enabled;}
  void ResetLastUpdatedTime(void) {_last_updated = _clock->now().seconds();}
  void Lock(void) {_lock.lock();}
  void Unlock(void) {_lock.unlock();}
  std::string GetSourceName(void) const {return _source_name;}
  Filters GetFilter(void) const {return _filter;}
  double GetMinObstacleHeight(void) const {return _min_obstacle_height;}
  double GetMaxObstacleHeight(void) const {return _max_obstacle_height;}
  int GetVoxelMinPoints(void) const {return _voxel_min_points;}
  void SetMinObstacleHeight(const double & v) {_min_obstacle_height = v;}
  void SetMaxObstacleHeight(const double & v) {_max_obstacle_height = v;}
  void SetMinZ(const double & v) {_min_z = v;}
  void SetMaxZ(const double & v) {_max_z = v;}
  void SetVerticalFovAngle(const double & v) {_vertical_fov = v;}
  void SetVerticalFovPadding(const double & v) {_vertical_fov_padding = v;}
  void SetHorizontalFovAngle(const double & v) {_horizontal_fov = v;}

private:
  std::string _source_name;
  rclcpp::Clock::SharedPtr _clock;
  double _observation_keep_time, _expected_update_rate;
  double _min_obstacle_height, _max_obstacle_height, _obstacle_range, _min_z, _max_z;
  double _vertical_fov, _vertical_fov_padding, _horizontal_fov, _decay_acceleration;
  bool _marking, _clearing;
  Filters _filter;
  int _voxel_min_points;
  bool _clear_buffer_after_reading, _enabled;
  ModelType _model_type;
  double _last_updated;
  std::list<observation::MeasurementReading> _observation_list;
  std::list<double> _stamps;
  std::recursive_mutex _lock;
};
}  // namespace buffer
#endif
