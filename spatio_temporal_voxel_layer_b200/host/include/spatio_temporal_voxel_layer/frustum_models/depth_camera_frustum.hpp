// geometry::DepthCameraFrustum — six-plane depth camera frustum (reference depth_camera_frustum.hpp/.cpp).
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__FRUSTUM_MODELS__DEPTH_CAMERA_FRUSTUM_HPP_
#define SPATIO_TEMPORAL_VOXEL_LAYER__FRUSTUM_MODELS__DEPTH_CAMERA_FRUSTUM_HPP_
#include "spatio_temporal_voxel_layer/frustum_models/frustum.hpp"

namespace geometry
{
class DepthCameraFrustum : public Frustum
{
public:
  DepthCameraFrustum(const double & vFOV, const double & hFOV, const double & min_dist, const double & max_dist)
  : _vFOV(vFOV), _hFOV(hFOV), _min_d(min_dist), _max_d(max_dist) {}
  // ComputePlaneNormals + TransformFrames of the six planes (reference depth_camera_frustum.cpp:67-174)
  void TransformModel(void) override
  {
    const double o[3] = {_position.x, _position.y, _position.z};
    const double q[4] = {_orientation.x, _orientation.y, _orientation.z, _orientation.w};
    stvl_build_depth_frustum(_vFOV, _hFOV, _min_d, _max_d, o, q, 0.0, &_block);
  }
private:
  double _vFOV, _hFOV, _min_d, _max_d;
};
}  // namespace geometry
#endif
