#ifndef STVL_SHIM_GEOMETRY_MSGS_POINT32_HPP_
#define STVL_SHIM_GEOMETRY_MSGS_POINT32_HPP_
namespace geometry_msgs { namespace msg {
struct Point32 { float x = 0, y = 0, z = 0; };
} }
#endif
