// Minimal stand-in for the parts of rclcpp the hot path touches (Clock::now().seconds(), reference
// spatio_temporal_voxel_grid.cpp:155,275).  Only used when the real ROS 2 headers are not on the include path;
// in a ROS workspace the real <rclcpp/rclcpp.hpp> is found first and this directory is simply not added.
#ifndef STVL_SHIM_RCLCPP_HPP_
#define STVL_SHIM_RCLCPP_HPP_
#include <chrono>
#include <memory>
#include <functional>

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

class Clock
{
public:
  typedef std::shared_ptr<Clock> SharedPtr;
  Clock() {}
  // tests inject a deterministic time source; the default is the system clock
  explicit Clock(std::function<double()> source) : source_(source) {}
  Time now() const
  {
    if (source_) return Time(source_());
    using namespace std::chrono;
    return Time(duration_cast<duration<double>>(system_clock::now().time_since_epoch()).count());
  }
private:
  std::function<double()> source_;
};
}  // namespace rclcpp
#endif
