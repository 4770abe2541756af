#include "rclcpp/rclcpp.hpp"
