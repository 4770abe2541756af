// spatio_temporal_voxel_layer::SpatioTemporalVoxelLayer — update-cycle host code around the GPU grid.
// Follows reference spatio_temporal_voxel_layer/src/spatio_temporal_voxel_layer.cpp for the functions on the cycle:
// onInitialize :69-360 (parameters -> grid + buffers), Get*Observations :466-526, updateFootprint :529-548,
// reset :590-606, updateCosts :666-696, UpdateROSCostmap :699-725, updateBounds :728-809, clearArea :950-963.
#include <cmath>
#include <ctime>
#include <iostream>

#include "spatio_temporal_voxel_layer/spatio_temporal_voxel_layer.hpp"

namespace spatio_temporal_voxel_layer
{

/*****************************************************************************/
SpatioTemporalVoxelLayer::SpatioTemporalVoxelLayer(void)
: _publish_voxels(false), _mapping_mode(false), was_reset_(false), _voxel_size(0.05), _voxel_decay(-1.0),
  _combination_method(1), _mark_threshold(0), _decay_model(volume_grid::LINEAR), _update_footprint_enabled(true),
  _enabled(true), _fused_costmap(true), _last_map_save_time(0.0), _map_save_duration(60.0)
/*****************************************************************************/
{
}

/*****************************************************************************/
SpatioTemporalVoxelLayer::~SpatioTemporalVoxelLayer(void)
/*****************************************************************************/
{
  _voxel_grid.reset();
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::configure(
  const LayerParameters & params, rclcpp::Clock::SharedPtr clock, nav2_costmap_2d::LayeredCostmap * parent,
  const std::string & name)
/*****************************************************************************/
{
  _params = params;
  _clock = clock;
  layered_costmap_ = parent;
  name_ = name;
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::onInitialize(void)
/*****************************************************************************/
{
  // initializes the layer from its parameters
  _global_frame = layered_costmap_ ? layered_costmap_->getGlobalFrameID() : std::string("map");
  _enabled = _params.enabled;
  enabled_ = _enabled;
  _publish_voxels = _params.publish_voxel_map;
  _voxel_size = _params.voxel_size;
  _combination_method = _params.combination_method;
  _mark_threshold = _params.mark_threshold;
  _update_footprint_enabled = _params.update_footprint_enabled;
  _decay_model = static_cast<volume_grid::GlobalDecayModel>(_params.decay_model);
  _voxel_decay = _params.voxel_decay;
  _mapping_mode = _params.mapping_mode;
  _map_save_duration = _params.map_save_duration;
  _last_map_save_time = _clock->now().seconds();

  if (_params.track_unknown_space) {
    default_value_ = nav2_costmap_2d::NO_INFORMATION;
  } else {
    default_value_ = nav2_costmap_2d::FREE_SPACE;
  }

  // the background value handed to the grid is the layer's default cost, as upstream (:157-159)
  _voxel_grid = std::make_unique<volume_grid::SpatioTemporalVoxelGrid>(
    _clock, static_cast<float>(_voxel_size), static_cast<double>(default_value_), static_cast<int>(_decay_model),
    _voxel_decay, _publish_voxels);
  _voxel_grid->SetExactClearedCells(!_fused_costmap);

  matchSize();

  // one buffer per observation source; marking / clearing lists as upstream (:204-266, :344-352)
  _observation_buffers.clear(); _marking_buffers.clear(); _clearing_buffers.clear();
  for (const SourceParameters & s : _params.observation_sources) {
    buffer::Filters filter = buffer::Filters::NONE;
    if (s.filter == "passthrough") {
      filter = buffer::Filters::PASSTHROUGH;
    } else if (s.filter == "voxel") {
      filter = buffer::Filters::VOXEL;
    }
    ModelType model_type = s.model_type == 1 ? THREE_DIMENSIONAL_LIDAR : DEPTH_CAMERA;
    auto b = std::make_shared<buffer::MeasurementBuffer>(
      s.name, _clock, s.observation_persistence, s.expected_update_rate, s.min_obstacle_height, s.max_obstacle_height,
      s.obstacle_range, s.min_z, s.max_z, s.vertical_fov_angle, s.vertical_fov_padding, s.horizontal_fov_angle,
      s.decay_acceleration, s.marking, s.clearing, filter, s.voxel_min_points, s.clear_after_reading, model_type);
    b->SetEnabled(s.enabled);
    _observation_buffers.push_back(b);
    if (s.marking) {_marking_buffers.push_back(b);}
    if (s.clearing) {_clearing_buffers.push_back(b);}
  }
  current_ = true;
  was_reset_ = false;
}

/*****************************************************************************/
std::shared_ptr<buffer::MeasurementBuffer> SpatioTemporalVoxelLayer::GetBuffer(const std::string & source_name) const
/*****************************************************************************/
{
  for (const auto & b : _observation_buffers) {
    if (b->GetSourceName() == source_name) {return b;}
  }
  return nullptr;
}

/*****************************************************************************/
bool SpatioTemporalVoxelLayer::GetMarkingObservations(
  std::vector<observation::MeasurementReading> & marking_observations) const
/*****************************************************************************/
{
  // get marking observations and static marked areas
  bool current = true;
  for (unsigned int i = 0; i != _marking_buffers.size(); ++i) {
    _marking_buffers[i]->Lock();
    if (_marking_buffers[i]->IsEnabled()) {
      _marking_buffers[i]->GetReadings(marking_observations);
    }
    current = current && _marking_buffers[i]->UpdatedAtExpectedRate();
    _marking_buffers[i]->Unlock();
  }
  marking_observations.insert(marking_observations.end(), _static_observations.begin(), _static_observations.end());
  return current;
}

/*****************************************************************************/
bool SpatioTemporalVoxelLayer::GetClearingObservations(
  std::vector<observation::MeasurementReading> & clearing_observations) const
/*****************************************************************************/
{
  // get clearing observations
  bool current = true;
  for (unsigned int i = 0; i != _clearing_buffers.size(); ++i) {
    _clearing_buffers[i]->Lock();
    if (_clearing_buffers[i]->IsEnabled()) {
      _clearing_buffers[i]->GetReadings(clearing_observations);
    }
    current = current && _clearing_buffers[i]->UpdatedAtExpectedRate();
    _clearing_buffers[i]->Unlock();
  }
  return current;
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::ObservationsResetAfterReading() const
/*****************************************************************************/
{
  for (unsigned int i = 0; i != _clearing_buffers.size(); ++i) {
    _clearing_buffers[i]->Lock();
    if (_clearing_buffers[i]->ClearAfterReading()) {
      _clearing_buffers[i]->ResetAllMeasurements();
    }
    _clearing_buffers[i]->Unlock();
  }
  for (unsigned int i = 0; i != _marking_buffers.size(); ++i) {
    _marking_buffers[i]->Lock();
    if (_marking_buffers[i]->ClearAfterReading()) {
      _marking_buffers[i]->ResetAllMeasurements();
    }
    _marking_buffers[i]->Unlock();
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelLayer::updateFootprint(
  double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y)
/*****************************************************************************/
{
  // updates layer costmap to include footprint for clearing in voxel grid
  if (!_update_footprint_enabled) {
    return false;
  }
  // nav2_costmap_2d::transformFootprint
  const double cos_th = std::cos(robot_yaw), sin_th = std::sin(robot_yaw);
  _transformed_footprint.clear();
  for (const auto & p : _footprint) {
    geometry_msgs::msg::Point q;
    q.x = robot_x + (p.x * cos_th - p.y * sin_th);
    q.y = robot_y + (p.x * sin_th + p.y * cos_th);
    _transformed_footprint.push_back(q);
  }
  for (unsigned int i = 0; i < _transformed_footprint.size(); i++) {
    touch(_transformed_footprint[i].x, _transformed_footprint[i].y, min_x, min_y, max_x, max_y);
  }
  return true;
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::reset(void)
/*****************************************************************************/
{
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);
  // reset layer
  Costmap2D::resetMaps();
  this->ResetGrid();
  current_ = false;
  was_reset_ = true;
  for (auto & b : _observation_buffers) {
    b->ResetLastUpdatedTime();
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelLayer::AddStaticObservations(const observation::MeasurementReading & obs)
/*****************************************************************************/
{
  // observations to always be added to the map each update cycle not marked
  try {
    _static_observations.push_back(obs);
    return true;
  } catch (...) {
    return false;
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelLayer::RemoveStaticObservations(void)
/*****************************************************************************/
{
  _static_observations.clear();
  return true;
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::ResetGrid(void)
/*****************************************************************************/
{
  if (!_voxel_grid->ResetGrid()) {
    std::cout << "Did not clear level set in " << getName() << "!" << std::endl;
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelLayer::SaveGrid(const std::string & file_name, double & map_size_bytes)
/*****************************************************************************/
{
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);
  return _voxel_grid->SaveGrid(file_name, map_size_bytes);
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::matchSize(void)
/*****************************************************************************/
{
  // match the master costmap size, volume_grid maintains full w/ expiration.
  CostmapLayer::matchSize();
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::setFusedCostmap(bool on)
/*****************************************************************************/
{
  _fused_costmap = on;
  if (_voxel_grid) {
    _voxel_grid->SetExactClearedCells(!on);
  }
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::updateCosts(
  nav2_costmap_2d::Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j)
/*****************************************************************************/
{
  // update costs in master_grid with costmap_
  if (!_enabled) {
    return;
  }
  // if not current due to reset, set current now after clearing
  if (!current_ && was_reset_) {
    was_reset_ = false;
    current_ = true;
  }
  if (_update_footprint_enabled) {
    setConvexPolygonCost(_transformed_footprint, nav2_costmap_2d::FREE_SPACE);
  }
  switch (_combination_method) {
    case 0:
      updateWithOverwrite(master_grid, min_i, min_j, max_i, max_j);
      break;
    case 1:
      updateWithMax(master_grid, min_i, min_j, max_i, max_j);
      break;
    default:
      break;
  }
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::UpdateROSCostmap(
  double * min_x, double * min_y, double * max_x, double * max_y,
  std::unordered_set<volume_grid::occupany_cell> & cleared_cells)
/*****************************************************************************/
{
  if (_fused_costmap) {
    // resetMaps + threshold + worldToMap + LETHAL write + touch, all on the device; the bytes land in costmap_
    stvl_costmap_result_t res;
    if (!_voxel_grid->FlattenToCostmap(
        origin_x_, origin_y_, resolution_, size_x_, size_y_, _mark_threshold, default_value_, costmap_.data(), &res))
    {
      std::cout << "Failed to flatten the voxel grid: " << stvl_last_error() << std::endl;
      return;
    }
    if (res.has_bounds) {
      touch(res.touched[0], res.touched[1], min_x, min_y, max_x, max_y);
      touch(res.touched[2], res.touched[3], min_x, min_y, max_x, max_y);
    }
    // cleared cells only ever feed touch(): their bounding box is equivalent
    double cb[4];
    if (_voxel_grid->GetClearedBounds(cb)) {
      touch(cb[0], cb[1], min_x, min_y, max_x, max_y);
      touch(cb[2], cb[3], min_x, min_y, max_x, max_y);
    }
    return;
  }

  // the reference's literal host loops (:705-724)
  Costmap2D::resetMaps();
  std::unordered_map<volume_grid::occupany_cell, uint>::iterator it;
  for (it = _voxel_grid->GetFlattenedCostmap()->begin(); it != _voxel_grid->GetFlattenedCostmap()->end(); ++it) {
    uint map_x, map_y;
    if (static_cast<int>(it->second) >= _mark_threshold && worldToMap(it->first.x, it->first.y, map_x, map_y)) {
      costmap_[getIndex(map_x, map_y)] = nav2_costmap_2d::LETHAL_OBSTACLE;
      touch(it->first.x, it->first.y, min_x, min_y, max_x, max_y);
    }
  }
  std::unordered_set<volume_grid::occupany_cell>::iterator cell;
  for (cell = cleared_cells.begin(); cell != cleared_cells.end(); ++cell) {
    touch(cell->x, cell->y, min_x, min_y, max_x, max_y);
  }
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::updateBounds(
  double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y)
/*****************************************************************************/
{
  // grabs new max bounds for the costmap
  if (!_enabled) {
    return;
  }
  // costmap mutex first, then the grid lock (clearArea dead-lock note, reference :736-742)
  std::unique_lock<Costmap2D::mutex_t> cm_lock(*getMutex());
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);

  if (layered_costmap_ && layered_costmap_->isRolling()) {
    updateOrigin(robot_x - getSizeInMetersX() / 2, robot_y - getSizeInMetersY() / 2);
  }
  useExtraBounds(min_x, min_y, max_x, max_y);

  bool current = true;
  std::vector<observation::MeasurementReading> marking_observations, clearing_observations;
  current = GetMarkingObservations(marking_observations) && current;
  current = GetClearingObservations(clearing_observations) && current;
  ObservationsResetAfterReading();
  current_ = current;

  std::unordered_set<volume_grid::occupany_cell> cleared_cells;

  // navigation mode: clear observations, mapping mode: save maps and publish
  bool should_save = false;
  if (_mapping_mode) {
    should_save = _clock->now().seconds() - _last_map_save_time > _map_save_duration;
  }
  if (!_mapping_mode) {
    _voxel_grid->ClearFrustums(clearing_observations, cleared_cells);
  } else if (should_save) {
    _last_map_save_time = _clock->now().seconds();
    time_t rawtime;
    struct tm timeinfo;
    char time_buffer[100];
    time(&rawtime);
    localtime_r(&rawtime, &timeinfo);
    strftime(time_buffer, 100, "%F-%r", &timeinfo);
    double bytes = 0.;
    _voxel_grid->SaveGrid(time_buffer, bytes);
  }

  // mark observations
  _voxel_grid->Mark(marking_observations);

  // update the ROS Layered Costmap
  UpdateROSCostmap(min_x, min_y, max_x, max_y, cleared_cells);

  // publish point cloud in navigation mode
  if (_publish_voxels && !_mapping_mode) {
    std::unique_ptr<sensor_msgs::msg::PointCloud2> pc2 = std::make_unique<sensor_msgs::msg::PointCloud2>();
    _voxel_grid->GetOccupancyPointCloud(pc2);
    pc2->header.frame_id = _global_frame;
    _last_voxel_cloud = std::move(pc2);
  }

  // update footprint
  updateFootprint(robot_x, robot_y, robot_yaw, min_x, min_y, max_x, max_y);
}

/*****************************************************************************/
void SpatioTemporalVoxelLayer::clearArea(int start_x, int start_y, int end_x, int end_y, bool invert_area)
/*****************************************************************************/
{
  // convert map coords to world coords
  volume_grid::occupany_cell start_world(0, 0);
  volume_grid::occupany_cell end_world(0, 0);
  mapToWorld(start_x, start_y, start_world.x, start_world.y);
  mapToWorld(end_x, end_y, end_world.x, end_world.y);

  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);
  _voxel_grid->ResetGridArea(start_world, end_world, invert_area);
  CostmapLayer::clearArea(start_x, start_y, end_x, end_y, invert_area);
}

}  // namespace spatio_temporal_voxel_layer
