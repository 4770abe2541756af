// Stand-in for the parts of rclcpp the reference's grid touches: Clock::now().seconds() (reference
// spatio_temporal_voxel_grid.cpp:155,275) with an injectable time source, plus the types measurement_buffer.hpp declares
// members of.  TEST INFRASTRUCTURE (oracle/_ref).
#ifndef STVL_REF_SHIM_RCLCPP_HPP_
#define STVL_REF_SHIM_RCLCPP_HPP_
#include <functional>
#include <memory>
namespace rclcpp
{
class Time
{
public:
  explicit Time(double s = 0.0) : s_(s) {}
  double seconds() const { return s_; }
private:
  double s_;
};
class Duration { public: explicit Duration(double s = 0.0) : s_(s) {} double seconds() const { return s_; } private: double s_; };
class Logger {};
class Clock
{
public:
  typedef std::shared_ptr<Clock> SharedPtr;
  explicit Clock(std::function<double()> source) : source_(source) {}
  Time now() const { return Time(source_()); }
private:
  std::function<double()> source_;
};
class Node { public: typedef std::shared_ptr<Node> SharedPtr; };
}  // namespace rclcpp
#endif
