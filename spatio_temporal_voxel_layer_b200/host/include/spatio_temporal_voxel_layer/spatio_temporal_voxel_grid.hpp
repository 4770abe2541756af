// volume_grid::SpatioTemporalVoxelGrid — drop-in for the reference class of the same name
// (spatio_temporal_voxel_layer/include/spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp:122-189).
// Same public signatures; the OpenVDB DoubleGrid behind them is replaced by the GPU leaf hash of libstvl_b200.so.
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__SPATIO_TEMPORAL_VOXEL_GRID_HPP_
#define SPATIO_TEMPORAL_VOXEL_LAYER__SPATIO_TEMPORAL_VOXEL_GRID_HPP_

#include <math.h>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "rclcpp/rclcpp.hpp"
#include "sensor_msgs/msg/point_cloud2.hpp"
#include "sensor_msgs/point_cloud2_iterator.hpp"
#include "geometry_msgs/msg/point.hpp"
#include "geometry_msgs/msg/point32.hpp"

#include "spatio_temporal_voxel_layer/measurement_reading.h"
#include "spatio_temporal_voxel_layer/frustum_models/depth_camera_frustum.hpp"
#include "spatio_temporal_voxel_layer/frustum_models/three_dimensional_lidar_frustum.hpp"
#include "stvl_b200.h"

namespace volume_grid
{
enum GlobalDecayModel
{
  LINEAR = 0,
  EXPONENTIAL = 1,
  PERSISTENT = 2
};

// Structure for an occupied cell for map (voxel-centre world x,y — reference hpp:89-102)
struct occupany_cell
{
  occupany_cell(const double & _x, const double & _y) : x(_x), y(_y) {}
  bool operator==(const occupany_cell & other) const { return x == other.x && y == other.y; }
  double x, y;
};
}  // namespace volume_grid

namespace std
{
template<>
struct hash<volume_grid::occupany_cell>
{
  std::size_t operator()(const volume_grid::occupany_cell & k) const
  {
    return (std::hash<double>()(k.x) ^ (std::hash<double>()(k.y) << 1)) >> 1;
  }
};
}  // namespace std

namespace volume_grid
{
class SpatioTemporalVoxelGrid
{
public:
  SpatioTemporalVoxelGrid(
    rclcpp::Clock::SharedPtr clock, const float & voxel_size, const double & background_value, const int & decay_model,
    const double & voxel_decay, const bool & pub_voxels);
  ~SpatioTemporalVoxelGrid(void);

  // Core marking and clearing functions
  void Mark(const std::vector<observation::MeasurementReading> & marking_observations);
  void operator()(const observation::MeasurementReading & obs) const;
  void ClearFrustums(
    const std::vector<observation::MeasurementReading> & clearing_observations,
    std::unordered_set<occupany_cell> & cleared_cells);

  // Get the pointcloud of the underlying occupancy grid
  void GetOccupancyPointCloud(std::unique_ptr<sensor_msgs::msg::PointCloud2> & pc2);
  std::unordered_map<occupany_cell, uint> * GetFlattenedCostmap();

  // Clear the grid
  bool ResetGrid(void);
  void ResetGridArea(const occupany_cell & start, const occupany_cell & end, bool invert_area = false);

  // Save the grid with size information (simple binary dump of (ix,iy,iz,time); .vdb needs OpenVDB)
  bool SaveGrid(const std::string & file_name, double & map_size_bytes);
  // Extension (the reference has no loader; its saved .vdb files are read back by vdb2pc only): replaces the grid's
  // contents with the voxels of a file written by SaveGrid ("<file_name>.stvl"; the voxel size must match).
  bool LoadGrid(const std::string & file_name);

  // ---- additions for the fused GPU path (not in the reference) ----------------------------------------------------
  // UpdateROSCostmap's grid loop on the device: bytes of the layer's costmap + touched bounds, no host map in between
  bool FlattenToCostmap(
    double origin_x, double origin_y, double resolution, unsigned int size_x, unsigned int size_y, int mark_threshold,
    unsigned char default_value, unsigned char * costmap, stvl_costmap_result_t * result);
  // bounding box of the cells cleared by the last ClearFrustums (all UpdateROSCostmap needs from cleared_cells)
  bool GetClearedBounds(double bounds[4]);
  // when false, ClearFrustums does not materialise cleared_cells (use GetClearedBounds); default true = reference behaviour
  void SetExactClearedCells(bool on);
  stvl_grid_t * Handle() { return _grid; }

protected:
  bool MarkReadings(const observation::MeasurementReading * obs, size_t n) const;

  rclcpp::Clock::SharedPtr _clock;
  mutable stvl_grid_t * _grid;
  int _decay_model;
  double _background_value, _voxel_size, _voxel_decay;
  bool _pub_voxels;
  bool _exact_cleared_cells;
  std::unordered_map<occupany_cell, uint> * _cost_map;
  bool _cost_map_valid;
  mutable std::mutex _grid_lock;
};
}  // namespace volume_grid
#endif  // SPATIO_TEMPORAL_VOXEL_LAYER__SPATIO_TEMPORAL_VOXEL_GRID_HPP_
