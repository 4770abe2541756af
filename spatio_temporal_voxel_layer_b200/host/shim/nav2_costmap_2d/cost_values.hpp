#ifndef STVL_SHIM_NAV2_COST_VALUES_HPP_
#define STVL_SHIM_NAV2_COST_VALUES_HPP_
namespace nav2_costmap_2d
{
static constexpr unsigned char NO_INFORMATION = 255;
static constexpr unsigned char LETHAL_OBSTACLE = 254;
static constexpr unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;
static constexpr unsigned char FREE_SPACE = 0;
}
#endif
