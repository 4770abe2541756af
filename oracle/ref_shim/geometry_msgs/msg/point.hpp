#ifndef STVL_REF_SHIM_GEOMETRY_MSGS_POINT_HPP_
#define STVL_REF_SHIM_GEOMETRY_MSGS_POINT_HPP_
namespace geometry_msgs { namespace msg { struct Point { double x = 0., y = 0., z = 0.; }; } }
#endif
