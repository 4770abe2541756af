#ifndef STVL_REF_SHIM_GEOMETRY_MSGS_QUATERNION_HPP_
#define STVL_REF_SHIM_GEOMETRY_MSGS_QUATERNION_HPP_
namespace geometry_msgs { namespace msg { struct Quaternion { double x = 0., y = 0., z = 0., w = 1.; }; } }
#endif
