#ifndef STVL_REF_SHIM_GEOMETRY_MSGS_POINT32_HPP_
#define STVL_REF_SHIM_GEOMETRY_MSGS_POINT32_HPP_
namespace geometry_msgs { namespace msg { struct Point32 { float x = 0.f, y = 0.f, z = 0.f; }; } }
#endif
