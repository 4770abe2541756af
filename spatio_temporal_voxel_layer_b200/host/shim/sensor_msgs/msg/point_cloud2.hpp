// Minimal sensor_msgs::msg::PointCloud2 / PointField (layout-compatible field names) for builds without ROS 2.
#ifndef STVL_SHIM_SENSOR_MSGS_POINT_CLOUD2_HPP_
#define STVL_SHIM_SENSOR_MSGS_POINT_CLOUD2_HPP_
#include <cstdint>
#include <string>
#include <vector>
namespace sensor_msgs { namespace msg {
struct PointField
{
  enum : uint8_t { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct Header { std::string frame_id; double stamp = 0; };
struct PointCloud2
{
  Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
} }
#endif
