// Stand-in for sensor_msgs/msg/PointCloud2 and PointField (field names, offsets, point_step, raw little-endian data).
#ifndef STVL_REF_SHIM_SENSOR_MSGS_POINT_CLOUD2_HPP_
#define STVL_REF_SHIM_SENSOR_MSGS_POINT_CLOUD2_HPP_
#include <cstdint>
#include <string>
#include <vector>
namespace sensor_msgs { namespace msg {
struct PointField
{
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2
{
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
} }
#endif
