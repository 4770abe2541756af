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

SpatioTemporalVoxelLayer::SpatioTemporalVoxelLayer(void)
: _publish_voxels(false), _mapping_mode(false), was_reset_(false), _voxel_size(0.05), _voxel_decay(-1.0),
  _combination_method(1), _mark_threshold(0), _decay_model(volume_grid::LINEAR), _update_footprint_enabled(true),
  _enabled(true), _fused_costmap(true), _last_map_save_time(0.0), _map_save_duration(60.0)
{
}

SpatioTemporalVoxelLayer::~SpatioTemporalVoxelLayer(void)
{
  _voxel_grid.reset();
}

void SpatioTemporalVoxelLayer::configure(
  const LayerParameters & params, rclcpp::Clock::SharedPtr clock, nav2_costmap_2d::LayeredCostmap * parent,
  const std::string & name)
{
  _params = params;
  _clock = clock;
  layered_costmap_ = parent;
  name_ = name;
}

void SpatioTemporalVoxelLayer::onInitialize(void)
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

std::shared_ptr<buffer::MeasurementBuffer> SpatioTemporalVoxelLayer::GetBuffer(const std::string & source_name) const
{
  for (const auto & candidate : _observation_buffers) {
    if (candidate->GetSourceName() == source_name) {return candidate;}
  }
  return nullptr;
}

namespace
{
// RAII for MeasurementBuffer::Lock / Unlock (upstream brackets every buffer access by hand)
struct ScopedBufferLock
{
  explicit ScopedBufferLock(buffer::MeasurementBuffer & b) : _b(b) {_b.Lock();}
  ~ScopedBufferLock() {_b.Unlock();}
  buffer::MeasurementBuffer & _b;
};

// readings of every enabled buffer of `group`, appended to `out`; false if any buffer is behind its expected rate
// (the body the reference repeats for the marking and the clearing buffers, :466-513)
bool collectReadings(
  const std::vector<std::shared_ptr<buffer::MeasurementBuffer>> & group, std::vector<observation::MeasurementReading> & out)
{
  bool all_current = true;
  for (const auto & source : group) {
    ScopedBufferLock guard(*source);
    if (source->IsEnabled()) {source->GetReadings(out);}
    if (!source->UpdatedAtExpectedRate()) {all_current = false;}
  }
  return all_current;
}

// "clear_after_reading" sources forget their readings once a cycle has consumed them (:516-526)
void dropConsumedReadings(const std::vector<std::shared_ptr<buffer::MeasurementBuffer>> & group)
{
  for (const auto & source : group) {
    ScopedBufferLock guard(*source);
    if (source->ClearAfterReading()) {source->ResetAllMeasurements();}
  }
}
}  // namespace

bool SpatioTemporalVoxelLayer::GetMarkingObservations(std::vector<observation::MeasurementReading> & marking_observations) const
{
  const bool all_current = collectReadings(_marking_buffers, marking_observations);
  // statically added observations are marked every cycle on top of the sensor readings
  for (const auto & fixed : _static_observations) {marking_observations.push_back(fixed);}
  return all_current;
}

bool SpatioTemporalVoxelLayer::GetClearingObservations(std::vector<observation::MeasurementReading> & clearing_observations) const
{
  return collectReadings(_clearing_buffers, clearing_observations);
}

void SpatioTemporalVoxelLayer::ObservationsResetAfterReading() const
{
  dropConsumedReadings(_clearing_buffers);
  dropConsumedReadings(_marking_buffers);
}

bool SpatioTemporalVoxelLayer::updateFootprint(
  double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y)
{
  if (!_update_footprint_enabled) {return false;}
  // nav2_costmap_2d::transformFootprint: rotate the footprint by the robot's yaw, move it to the robot; every vertex
  // extends the update bounds (:529-548)
  const double c = std::cos(robot_yaw), s = std::sin(robot_yaw);
  _transformed_footprint.resize(_footprint.size());
  for (size_t k = 0; k < _footprint.size(); ++k) {
    geometry_msgs::msg::Point & v = _transformed_footprint[k];
    v.x = robot_x + (_footprint[k].x * c - _footprint[k].y * s);
    v.y = robot_y + (_footprint[k].x * s + _footprint[k].y * c);
    v.z = 0.0;
    touch(v.x, v.y, min_x, min_y, max_x, max_y);
  }
  return true;
}

void SpatioTemporalVoxelLayer::reset(void)
{
  // layer reset (:590-606): default costs, empty grid, "not current" until the next updateCosts, fresh rate timers
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);
  Costmap2D::resetMaps();
  ResetGrid();
  current_ = false;
  was_reset_ = true;
  for (const auto & source : _observation_buffers) {source->ResetLastUpdatedTime();}
}

bool SpatioTemporalVoxelLayer::AddStaticObservations(const observation::MeasurementReading & obs)
{
  try {
    _static_observations.push_back(obs);
  } catch (...) {
    return false;
  }
  return true;
}

bool SpatioTemporalVoxelLayer::RemoveStaticObservations(void)
{
  _static_observations.clear();
  return true;
}

void SpatioTemporalVoxelLayer::ResetGrid(void)
{
  if (_voxel_grid->ResetGrid()) {return;}
  std::cout << "Did not clear level set in " << getName() << "!" << std::endl;
}

bool SpatioTemporalVoxelLayer::SaveGrid(const std::string & file_name, double & map_size_bytes)
{
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);
  return _voxel_grid->SaveGrid(file_name, map_size_bytes);
}

void SpatioTemporalVoxelLayer::matchSize(void)
{
  // only the 2-D layer follows the master's size; the voxel grid is unbounded and expires by itself
  CostmapLayer::matchSize();
}

void SpatioTemporalVoxelLayer::setFusedCostmap(bool on)
{
  _fused_costmap = on;
  if (_voxel_grid) {_voxel_grid->SetExactClearedCells(!on);}
}

void SpatioTemporalVoxelLayer::updateCosts(
  nav2_costmap_2d::Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j)
{
  if (!_enabled) {return;}
  // a reset leaves the layer "not current" until its first merge into the master (:676-679)
  if (was_reset_ && !current_) {
    current_ = true;
    was_reset_ = false;
  }
  // the robot cannot be an obstacle to itself (:682-684)
  if (_update_footprint_enabled) {setConvexPolygonCost(_transformed_footprint, nav2_costmap_2d::FREE_SPACE);}
  if (_combination_method == 0) {
    updateWithOverwrite(master_grid, min_i, min_j, max_i, max_j);
  } else if (_combination_method == 1) {
    updateWithMax(master_grid, min_i, min_j, max_i, max_j);
  }
}

void SpatioTemporalVoxelLayer::UpdateROSCostmap(
  double * min_x, double * min_y, double * max_x, double * max_y,
  std::unordered_set<volume_grid::occupany_cell> & cleared_cells)
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

  // host path, semantically the reference's loops (:705-724): default costs everywhere, LETHAL where a voxel column holds
  // at least mark_threshold voxels and lies on the map, bounds extended by every such column and every cleared cell
  Costmap2D::resetMaps();
  for (const auto & column : *_voxel_grid->GetFlattenedCostmap()) {
    if (static_cast<int>(column.second) < _mark_threshold) {continue;}
    unsigned int mx = 0, my = 0;
    if (!worldToMap(column.first.x, column.first.y, mx, my)) {continue;}
    costmap_[getIndex(mx, my)] = nav2_costmap_2d::LETHAL_OBSTACLE;
    touch(column.first.x, column.first.y, min_x, min_y, max_x, max_y);
  }
  for (const auto & freed : cleared_cells) {touch(freed.x, freed.y, min_x, min_y, max_x, max_y);}
}

namespace
{
// "%F-%r" of the wall clock: the file name upstream gives its periodic map dumps (:778-786)
std::string wallClockStamp()
{
  const time_t t = time(nullptr);
  struct tm parts;
  localtime_r(&t, &parts);
  char text[100];
  strftime(text, sizeof(text), "%F-%r", &parts);
  return text;
}
}  // namespace

void SpatioTemporalVoxelLayer::updateBounds(
  double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y)
{
  if (!_enabled) {return;}
  // lock order as upstream (:736-742): the costmap's mutex, then the grid's (clearArea takes them the same way round)
  std::unique_lock<Costmap2D::mutex_t> cm_lock(*getMutex());
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);

  if (layered_costmap_ && layered_costmap_->isRolling()) {
    updateOrigin(robot_x - getSizeInMetersX() / 2, robot_y - getSizeInMetersY() / 2);
  }
  useExtraBounds(min_x, min_y, max_x, max_y);

  // this cycle's readings; the layer is "current" only if every source delivers at its expected rate
  std::vector<observation::MeasurementReading> marking_observations, clearing_observations;
  const bool marking_current = GetMarkingObservations(marking_observations);
  const bool clearing_current = GetClearingObservations(clearing_observations);
  ObservationsResetAfterReading();
  current_ = marking_current && clearing_current;

  std::unordered_set<volume_grid::occupany_cell> cleared_cells;
  if (_mapping_mode) {
    // mapping mode never clears; it dumps the grid every map_save_duration seconds instead (:764-786)
    const double t = _clock->now().seconds();
    if (t - _last_map_save_time > _map_save_duration) {
      _last_map_save_time = t;
      double bytes = 0.;
      _voxel_grid->SaveGrid(wallClockStamp(), bytes);
    }
  } else {
    _voxel_grid->ClearFrustums(clearing_observations, cleared_cells);   // decay + frustum acceleration (:773)
  }
  _voxel_grid->Mark(marking_observations);                              // (:792)
  UpdateROSCostmap(min_x, min_y, max_x, max_y, cleared_cells);          // (:795)

  if (_publish_voxels && !_mapping_mode) {
    // navigation mode publishes the occupancy cloud on "voxel_grid" (:798-805); kept for the caller here
    auto cloud = std::make_unique<sensor_msgs::msg::PointCloud2>();
    _voxel_grid->GetOccupancyPointCloud(cloud);
    cloud->header.frame_id = _global_frame;
    _last_voxel_cloud = std::move(cloud);
  }
  updateFootprint(robot_x, robot_y, robot_yaw, min_x, min_y, max_x, max_y);   // (:808)
}

void SpatioTemporalVoxelLayer::clearArea(int start_x, int start_y, int end_x, int end_y, bool invert_area)
{
  // the rectangle in world coordinates for the voxel grid, in cells for the 2-D layer (:950-963)
  volume_grid::occupany_cell from(0, 0), to(0, 0);
  mapToWorld(start_x, start_y, from.x, from.y);
  mapToWorld(end_x, end_y, to.x, to.y);
  std::lock_guard<std::recursive_mutex> lock(_voxel_grid_lock);
  _voxel_grid->ResetGridArea(from, to, invert_area);
  CostmapLayer::clearArea(start_x, start_y, end_x, end_y, invert_area);
}

}  // namespace spatio_temporal_voxel_layer
