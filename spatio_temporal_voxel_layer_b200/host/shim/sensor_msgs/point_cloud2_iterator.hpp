// The two helpers of sensor_msgs/point_cloud2_iterator.hpp the grid uses to emit its occupancy cloud
// (reference spatio_temporal_voxel_grid.cpp:352-375): PointCloud2Modifier::setPointCloud2Fields for xyz FLOAT32.
#ifndef STVL_SHIM_SENSOR_MSGS_POINT_CLOUD2_ITERATOR_HPP_
#define STVL_SHIM_SENSOR_MSGS_POINT_CLOUD2_ITERATOR_HPP_
#include "sensor_msgs/msg/point_cloud2.hpp"
namespace sensor_msgs {
class PointCloud2Modifier
{
public:
  explicit PointCloud2Modifier(msg::PointCloud2 & c) : c_(c) {}
  // xyz, FLOAT32, 16-byte point step like the real setPointCloud2FieldsByString(1, "xyz") (x,y,z + 4 B padding)
  void setXYZ(uint32_t n)
  {
    c_.fields.clear();
    const char * names[3] = {"x", "y", "z"};
    for (uint32_t i = 0; i < 3; ++i) { msg::PointField f; f.name = names[i]; f.offset = 4 * i; f.datatype = msg::PointField::FLOAT32; f.count = 1; c_.fields.push_back(f); }
    c_.point_step = 16;
    c_.width = n; c_.height = 1;
    c_.row_step = c_.point_step * n;
    c_.data.assign((size_t)c_.row_step, 0);
  }
private:
  msg::PointCloud2 & c_;
};
}  // namespace sensor_msgs
#endif
