// spatio_temporal_voxel_layer::SpatioTemporalVoxelLayer — the nav2 costmap plugin, i.e. the CALLER of the hot path
// (reference spatio_temporal_voxel_layer.hpp:90-197, spatio_temporal_voxel_layer.cpp).  The update cycle
// (updateBounds -> ClearFrustums -> Mark -> UpdateROSCostmap, updateCosts) and its parameters are kept; ROS node
// plumbing (parameter server, subscriptions, TF message filters, services, dynamic parameters) is outside the hot
// path and outside this build, so parameters arrive in a plain struct and readings are pushed into the buffers by the
// embedding application (or by the ROS callbacks of a full ROS build).
#ifndef SPATIO_TEMPORAL_VOXEL_LAYER__SPATIO_TEMPORAL_VOXEL_LAYER_HPP_
#define SPATIO_TEMPORAL_VOXEL_LAYER__SPATIO_TEMPORAL_VOXEL_LAYER_HPP_

#include <memory>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "nav2_costmap_2d/costmap_layer.hpp"
#include "nav2_costmap_2d/cost_values.hpp"
#include "spatio_temporal_voxel_layer/measurement_buffer.hpp"
#include "spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp"

namespace spatio_temporal_voxel_layer
{
// per-source parameters, names and defaults as declared in onInitialize (reference spatio_temporal_voxel_layer.cpp:178-202)
struct SourceParameters
{
  std::string name;
  double observation_persistence = 0.0;
  double expected_update_rate = 0.0;
  double min_obstacle_height = 0.0, max_obstacle_height = 3.0;
  bool marking = true, clearing = false;
  double obstacle_range = 2.5;
  double min_z = 0.0, max_z = 10.0;
  double vertical_fov_angle = 0.7, vertical_fov_padding = 0.0, horizontal_fov_angle = 1.04;
  double decay_acceleration = 0.0;
  std::string filter = "passthrough";
  int voxel_min_points = 0;
  bool clear_after_reading = false;
  bool enabled = true;
  int model_type = 0;
};
// layer parameters, names and defaults as declared in onInitialize (reference spatio_temporal_voxel_layer.cpp:87-127)
struct LayerParameters
{
  bool enabled = true;
  bool publish_voxel_map = false;
  double voxel_size = 0.05;
  int combination_method = 1;
  int mark_threshold = 0;
  bool update_footprint_enabled = true;
  bool track_unknown_space = false;   // costmap-level parameter: NO_INFORMATION vs FREE_SPACE default (:137-141)
  int decay_model = 0;
  double voxel_decay = -1.0;
  bool mapping_mode = false;
  double map_save_duration = 60.0;
  std::vector<SourceParameters> observation_sources;
};

class SpatioTemporalVoxelLayer : public nav2_costmap_2d::CostmapLayer
{
public:
  SpatioTemporalVoxelLayer(void);
  virtual ~SpatioTemporalVoxelLayer(void);

  // stands in for the ROS parameter server + lifecycle node the reference reads in onInitialize
  void configure(
    const LayerParameters & params, rclcpp::Clock::SharedPtr clock, nav2_costmap_2d::LayeredCostmap * parent,
    const std::string & name = "stvl_layer");

  // Core Functions
  virtual void onInitialize(void);
  virtual void updateBounds(
    double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y);
  virtual void updateCosts(nav2_costmap_2d::Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j);
  virtual void matchSize(void);
  virtual void reset(void);
  virtual void clearArea(int start_x, int start_y, int end_x, int end_y, bool invert_area = false) override;
  virtual bool isClearable() {return true;}

  // Functions for sensor feeds
  bool GetMarkingObservations(std::vector<observation::MeasurementReading> & marking_observations) const;
  bool GetClearingObservations(std::vector<observation::MeasurementReading> & clearing_observations) const;
  void ObservationsResetAfterReading() const;
  std::shared_ptr<buffer::MeasurementBuffer> GetBuffer(const std::string & source_name) const;
  bool AddStaticObservations(const observation::MeasurementReading & obs);
  bool RemoveStaticObservations(void);
  void setFootprint(const std::vector<geometry_msgs::msg::Point> & footprint) {_footprint = footprint;}
  void setMarkThreshold(int t) {_mark_threshold = t;}

  // Functions to interact with maps
  void UpdateROSCostmap(
    double * min_x, double * min_y, double * max_x, double * max_y,
    std::unordered_set<volume_grid::occupany_cell> & cleared_cells);
  bool updateFootprint(
    double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y);
  void ResetGrid(void);
  bool SaveGrid(const std::string & file_name, double & map_size_bytes);
  // the last published occupancy cloud (the reference publishes it on "voxel_grid", :798-805)
  const sensor_msgs::msg::PointCloud2 * LastVoxelCloud() const {return _last_voxel_cloud.get();}
  volume_grid::SpatioTemporalVoxelGrid * Grid() {return _voxel_grid.get();}
  // true (default): UpdateROSCostmap runs fused on the GPU (stvl_costmap) and only the cleared cells' bounding box is
  // fetched; false: the reference's literal host loops over GetFlattenedCostmap() and cleared_cells
  void setFusedCostmap(bool on);

private:
  LayerParameters _params;
  rclcpp::Clock::SharedPtr _clock;
  std::vector<std::shared_ptr<buffer::MeasurementBuffer>> _observation_buffers, _marking_buffers, _clearing_buffers;
  bool _publish_voxels, _mapping_mode, was_reset_;
  std::string _global_frame;
  double _voxel_size, _voxel_decay;
  int _combination_method, _mark_threshold;
  volume_grid::GlobalDecayModel _decay_model;
  bool _update_footprint_enabled, _enabled, _fused_costmap;
  double _last_map_save_time, _map_save_duration;
  std::vector<geometry_msgs::msg::Point> _footprint, _transformed_footprint;
  std::vector<observation::MeasurementReading> _static_observations;
  std::unique_ptr<volume_grid::SpatioTemporalVoxelGrid> _voxel_grid;
  std::unique_ptr<sensor_msgs::msg::PointCloud2> _last_voxel_cloud;
  mutable std::recursive_mutex _voxel_grid_lock;
};
}  // namespace spatio_temporal_voxel_layer
#endif
