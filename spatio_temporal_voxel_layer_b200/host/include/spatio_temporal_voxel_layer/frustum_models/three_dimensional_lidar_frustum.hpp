// geometry::ThreeDimensionalLidarFrustum — analytic hourglass of a spinning lidar (reference
// three_dimensional_lidar_frustum.hpp/.cpp).
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__FRUSTUM_MODELS__THREE_DIMENSIONAL_LIDAR_FRUSTUM_HPP_
#define SPATIO_TEMPORAL_VOXEL_LAYER__FRUSTUM_MODELS__THREE_DIMENSIONAL_LIDAR_FRUSTUM_HPP_
#include "spatio_temporal_voxel_layer/frustum_models/frustum.hpp"

namespace geometry
{
class ThreeDimensionalLidarFrustum : public Frustum
{
public:
  ThreeDimensionalLidarFrustum(const double & vFOV, const double & vFOVPadding, const double & hFOV, const double & min_dist, const double & max_dist)
  : _vFOV(vFOV), _vFOVPadding(vFOVPadding), _hFOV(hFOV), _min_d(min_dist), _max_d(max_dist) {}
  // ctor precompute + conjugate orientation (reference three_dimensional_lidar_frustum.cpp:44-74)
  void TransformModel(void) override
  {
    const double o[3] = {_position.x, _position.y, _position.z};
    const double q[4] = {_orientation.x, _orientation.y, _orientation.z, _orientation.w};
    stvl_build_lidar_frustum(_vFOV, _vFOVPadding, _hFOV, _min_d, _max_d, o, q, 0.0, &_block);
  }
private:
  double _vFOV, _vFOVPadding, _hFOV, _min_d, _max_d;
};
}  // namespace geometry
#endif
