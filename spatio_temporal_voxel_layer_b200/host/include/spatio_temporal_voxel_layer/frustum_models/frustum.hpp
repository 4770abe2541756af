// geometry::Frustum — abstract sensor frustum, host side.  The reference's interface (frustum.hpp:96-116) is kept:
// SetPosition / SetOrientation / TransformModel build the model in the global frame.  IsInside — the reference's
// inner-loop predicate — is evaluated ON THE GPU by the clear sweep from the parameter block Block() returns.
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__FRUSTUM_MODELS__FRUSTUM_HPP_
#define SPATIO_TEMPORAL_VOXEL_LAYER__FRUSTUM_MODELS__FRUSTUM_HPP_

#include "geometry_msgs/msg/point.hpp"
#include "geometry_msgs/msg/quaternion.hpp"
#include "stvl_b200.h"

namespace geometry
{
class Frustum
{
public:
  Frustum() {}
  virtual ~Frustum(void) {}
  // set pose of the sensor in global space
  virtual void SetPosition(const geometry_msgs::msg::Point & origin) { _position = origin; }
  virtual void SetOrientation(const geometry_msgs::msg::Quaternion & quat) { _orientation = quat; }
  // transform model to the current coordinates: fills the device parameter block
  virtual void TransformModel(void) = 0;
  // the block the sweep kernel stages in shared memory (valid after TransformModel)
  const stvl_frustum_t & Block(double accel_factor) { _block.accel_factor = accel_factor; return _block; }

protected:
  geometry_msgs::msg::Point _position;
  geometry_msgs::msg::Quaternion _orientation;
  stvl_frustum_t _block{};
};
}  // namespace geometry
#endif
