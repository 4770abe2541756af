// stvl_geometry.h — host-side frustum geometry (see stvl_geometry.cpp)
#pragma once
#include "stvl_b200.h"
#include "stvl_device.cuh"

namespace stvl {
int build_depth_frustum(double vfov, double hfov, double min_d, double max_d, const double origin[3],
                        const double quat_xyzw[4], double accel, stvl_frustum_t * out);
int build_lidar_frustum(double vfov, double vpad, double hfov, double min_d, double max_d, const double origin[3],
                        const double quat_xyzw[4], double accel, stvl_frustum_t * out);
void prepare_frustum(const stvl_frustum_t & in, double vs, DevFrustum * out);
}  // namespace stvl
