// volume_grid::SpatioTemporalVoxelGrid on the GPU — the C++ façade that keeps the reference's interface
// (reference spatio_temporal_voxel_layer/src/spatio_temporal_voxel_grid.cpp) and forwards every method to the C-ABI of
// libstvl_b200.so (include/stvl_b200.h).  No voxel is touched on the host: there is no CPU fallback.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <stdexcept>

#include "spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp"

namespace volume_grid
{

namespace
{
// byte offset of the FLOAT32 field `name` — what sensor_msgs::PointCloud2ConstIterator<float>(cloud, name) resolves
bool field_offset(const sensor_msgs::msg::PointCloud2 & cloud, const char * name, uint32_t & off)
{
  for (const auto & f : cloud.fields) {
    if (f.name == name) {
      off = f.offset;
      return f.datatype == sensor_msgs::msg::PointField::FLOAT32;
    }
  }
  return false;
}
}  // namespace

/*****************************************************************************/
SpatioTemporalVoxelGrid::SpatioTemporalVoxelGrid(
  rclcpp::Clock::SharedPtr clock, const float & voxel_size, const double & background_value, const int & decay_model,
  const double & voxel_decay, const bool & pub_voxels)
: _clock(clock), _grid(nullptr), _decay_model(decay_model), _background_value(background_value),
  _voxel_size(voxel_size), _voxel_decay(voxel_decay), _pub_voxels(pub_voxels), _exact_cleared_cells(true),
  _cost_map(new std::unordered_map<occupany_cell, uint>), _cost_map_valid(true)
/*****************************************************************************/
{
  // InitializeGrid (reference .cpp:73-93): the OpenVDB grid + transform become a GPU leaf hash with the same scale
  stvl_config_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.voxel_size = voxel_size;
  cfg.decay_model = decay_model;
  cfg.background_value = background_value;
  cfg.voxel_decay = voxel_decay;
  cfg.pub_voxels = pub_voxels ? 1 : 0;
  cfg.device = -1;
  cfg.shard_count = 1;
  cfg.shard_width = 1;
  if (stvl_create(&cfg, &_grid) != STVL_OK) {
    // the reference's constructor cannot fail; a missing GPU is unrecoverable for this layer
    throw std::runtime_error(std::string("SpatioTemporalVoxelGrid: ") + stvl_last_error());
  }
}

/*****************************************************************************/
SpatioTemporalVoxelGrid::~SpatioTemporalVoxelGrid(void)
/*****************************************************************************/
{
  if (_cost_map) {
    delete _cost_map;
  }
  stvl_destroy(_grid);
}

/*****************************************************************************/
void SpatioTemporalVoxelGrid::ClearFrustums(
  const std::vector<observation::MeasurementReading> & clearing_readings,
  std::unordered_set<occupany_cell> & cleared_cells)
/*****************************************************************************/
{
  std::unique_lock<std::mutex> lock(_grid_lock);

  // (an empty grid sweeps nothing; the reference's IsGridEmpty early-out .cpp:104-108 has the same observable result)
  _cost_map->clear();
  _cost_map_valid = false;

  // frustum objects in READING ORDER; unknown model types are skipped (reference .cpp:120-144)
  std::vector<stvl_frustum_t> blocks;
  blocks.reserve(clearing_readings.size());
  for (auto it = clearing_readings.begin(); it != clearing_readings.end(); ++it) {
    std::unique_ptr<geometry::Frustum> frustum;
    if (it->_model_type == DEPTH_CAMERA) {
      frustum.reset(new geometry::DepthCameraFrustum(
        it->_vertical_fov_in_rad, it->_horizontal_fov_in_rad, it->_min_z_in_m, it->_max_z_in_m));
    } else if (it->_model_type == THREE_DIMENSIONAL_LIDAR) {
      frustum.reset(new geometry::ThreeDimensionalLidarFrustum(
        it->_vertical_fov_in_rad, it->_vertical_fov_padding_in_m, it->_horizontal_fov_in_rad, it->_min_z_in_m,
        it->_max_z_in_m));
    } else {
      continue;
    }
    frustum->SetPosition(it->_origin);
    frustum->SetOrientation(it->_orientation);
    frustum->TransformModel();
    blocks.push_back(frustum->Block(it->_decay_acceleration));
  }

  // TemporalClearAndGenerateCostmap (reference .cpp:149-224): one clock sample for the whole sweep
  const double cur_time = _clock->now().seconds();
  if (stvl_clear_frustums(_grid, cur_time, blocks.empty() ? nullptr : blocks.data(), (uint32_t)blocks.size()) != STVL_OK) {
    std::cout << "Failed to clear frustums: " << stvl_last_error() << std::endl;
    return;
  }
  if (_exact_cleared_cells) {
    uint64_t n = 0;
    if (stvl_get_cleared(_grid, nullptr, nullptr, 0, &n) == STVL_OK && n > 0) {
      std::vector<int32_t> ix(n), iy(n);
      uint64_t m = 0;
      if (stvl_get_cleared(_grid, ix.data(), iy.data(), n, &m) == STVL_OK) {
        for (uint64_t i = 0; i < m && i < n; ++i) {
          cleared_cells.insert(occupany_cell(stvl_index_to_world(_grid, ix[i]), stvl_index_to_world(_grid, iy[i])));
        }
      }
    }
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelGrid::MarkReadings(const observation::MeasurementReading * obs, size_t n) const
/*****************************************************************************/
{
  std::vector<stvl_reading_t> readings;
  readings.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const observation::MeasurementReading & o = obs[i];
    if (!o._marking) {   // reference .cpp:273
      continue;
    }
    const sensor_msgs::msg::PointCloud2 & cloud = *(o._cloud);
    stvl_reading_t r;
    std::memset(&r, 0, sizeof(r));
    if (!field_offset(cloud, "x", r.off_x) || !field_offset(cloud, "y", r.off_y) || !field_offset(cloud, "z", r.off_z)) {
      if ((size_t)cloud.width * cloud.height != 0) {
        std::cout << "Failed to mark: cloud has no FLOAT32 x/y/z fields." << std::endl;
      }
      continue;
    }
    r.data = cloud.data.data();
    r.n_points = (uint64_t)cloud.width * cloud.height;
    r.point_step = cloud.point_step;
    r.origin[0] = o._origin.x; r.origin[1] = o._origin.y; r.origin[2] = o._origin.z;
    r.obstacle_range = o._obstacle_range_in_m;
    r.now = _clock->now().seconds();   // one clock sample per reading, reference .cpp:275
    r.marking = 1;
    // NONE: the cloud was filtered upstream as in the reference (measurement_buffer.cpp:153-175); otherwise the buffer
    // deferred its z-window filter to the Mark kernel
    r.filter = o._device_filter;
    r.min_obstacle_height = o._min_obstacle_height;
    r.max_obstacle_height = o._max_obstacle_height;
    r.voxel_min_points = o._voxel_min_points;
    r.has_transform = o._has_device_transform ? 1 : 0;
    std::memcpy(r.transform, o._device_transform, sizeof(r.transform));
    readings.push_back(r);
  }
  if (readings.empty()) {
    return true;
  }
  if (stvl_mark(_grid, readings.data(), (uint32_t)readings.size()) != STVL_OK) {
    std::cout << "Failed to mark point." << " (" << stvl_last_error() << ")" << std::endl;
    return false;
  }
  return true;
}

/*****************************************************************************/
void SpatioTemporalVoxelGrid::Mark(const std::vector<observation::MeasurementReading> & marking_readings)
/*****************************************************************************/
{
  std::unique_lock<std::mutex> lock(_grid_lock);
  // mark the grid: all readings of the cycle go to the device in one batched launch (later readings overwrite)
  if (marking_readings.size() > 0) {
    MarkReadings(marking_readings.data(), marking_readings.size());
  }
}

/*****************************************************************************/
void SpatioTemporalVoxelGrid::operator()(const observation::MeasurementReading & obs) const
/*****************************************************************************/
{
  MarkReadings(&obs, 1);
}

/*****************************************************************************/
std::unordered_map<occupany_cell, uint> * SpatioTemporalVoxelGrid::GetFlattenedCostmap()
/*****************************************************************************/
{
  // the reference fills _cost_map inside the sweep (.cpp:227-251); here the per-column counts live on the device and
  // the host map is only materialised when somebody asks for it
  if (!_cost_map_valid) {
    _cost_map->clear();
    uint64_t n = 0;
    if (stvl_get_columns(_grid, nullptr, nullptr, nullptr, 0, &n) == STVL_OK && n > 0) {
      std::vector<int32_t> ix(n), iy(n);
      std::vector<uint32_t> cnt(n);
      uint64_t m = 0;
      if (stvl_get_columns(_grid, ix.data(), iy.data(), cnt.data(), n, &m) == STVL_OK) {
        _cost_map->reserve((size_t)m);
        for (uint64_t i = 0; i < m && i < n; ++i) {
          _cost_map->insert(std::make_pair(
            occupany_cell(stvl_index_to_world(_grid, ix[i]), stvl_index_to_world(_grid, iy[i])), (uint)cnt[i]));
        }
      }
    }
    _cost_map_valid = true;
  }
  return _cost_map;
}

/*****************************************************************************/
void SpatioTemporalVoxelGrid::GetOccupancyPointCloud(std::unique_ptr<sensor_msgs::msg::PointCloud2> & pc2)
/*****************************************************************************/
{
  // x,y,z FLOAT32, height 1, dense (reference .cpp:344-376)
  uint64_t n = 0;
  stvl_get_points(_grid, nullptr, 0, &n);
  std::vector<float> xyz((size_t)n * 3);
  uint64_t m = 0;
  if (n > 0 && stvl_get_points(_grid, xyz.data(), n, &m) != STVL_OK) {
    m = 0;
  }
  if (m > n) {m = n;}
  pc2->width = (uint32_t)m;
  pc2->height = 1;
  pc2->is_dense = true;
  pc2->fields.clear();
  const char * names[3] = {"x", "y", "z"};
  for (uint32_t i = 0; i < 3; ++i) {
    sensor_msgs::msg::PointField f;
    f.name = names[i];
    f.offset = 4 * i;
    f.datatype = sensor_msgs::msg::PointField::FLOAT32;
    f.count = 1;
    pc2->fields.push_back(f);
  }
  pc2->point_step = 16;   // what PointCloud2Modifier::setPointCloud2FieldsByString(1, "xyz") produces
  pc2->row_step = pc2->point_step * pc2->width;
  pc2->data.assign((size_t)pc2->row_step, 0);
  for (uint64_t i = 0; i < m; ++i) {
    std::memcpy(&pc2->data[i * 16], &xyz[3 * i], 12);
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelGrid::ResetGrid(void)
/*****************************************************************************/
{
  std::unique_lock<std::mutex> lock(_grid_lock);
  if (stvl_reset(_grid) == STVL_OK) {
    uint64_t n = 1;
    if (stvl_active_count(_grid, &n) == STVL_OK && n == 0) {
      return true;
    }
  }
  std::cout << "Failed to reset costmap, please try again." << std::endl;
  return false;
}

/*****************************************************************************/
void SpatioTemporalVoxelGrid::ResetGridArea(const occupany_cell & start, const occupany_cell & end, bool invert_area)
/*****************************************************************************/
{
  std::unique_lock<std::mutex> lock(_grid_lock);
  if (stvl_reset_area(_grid, start.x, start.y, end.x, end.y, invert_area ? 1 : 0) != STVL_OK) {
    std::cout << "Failed to clear area: " << stvl_last_error() << std::endl;
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelGrid::SaveGrid(const std::string & file_name, double & map_size_bytes)
/*****************************************************************************/
{
  // The reference writes an OpenVDB .vdb (reference .cpp:480-496); OpenVDB is exactly what this build does without, so
  // the grid is dumped as a documented flat binary: "STVLGRD1", voxel_size f64, n u64, then n x (ix,iy,iz i32, t f64).
  try {
    uint64_t n = 0;
    if (stvl_active_count(_grid, &n) != STVL_OK) {
      map_size_bytes = 0.;
      return false;
    }
    std::vector<int32_t> ix(n), iy(n), iz(n);
    std::vector<double> t(n);
    uint64_t m = 0;
    if (n && stvl_get_voxels(_grid, ix.data(), iy.data(), iz.data(), t.data(), n, &m) != STVL_OK) {
      map_size_bytes = 0.;
      return false;
    }
    std::ofstream f(file_name + ".stvl", std::ios::binary);
    if (!f) {
      map_size_bytes = 0.;
      return false;
    }
    const char magic[8] = {'S', 'T', 'V', 'L', 'G', 'R', 'D', '1'};
    const double vs = stvl_voxel_size(_grid);
    f.write(magic, 8);
    f.write(reinterpret_cast<const char *>(&vs), 8);
    f.write(reinterpret_cast<const char *>(&m), 8);
    for (uint64_t i = 0; i < m; ++i) {
      f.write(reinterpret_cast<const char *>(&ix[i]), 4);
      f.write(reinterpret_cast<const char *>(&iy[i]), 4);
      f.write(reinterpret_cast<const char *>(&iz[i]), 4);
      f.write(reinterpret_cast<const char *>(&t[i]), 8);
    }
    stvl_pool_stats_t ps;
    map_size_bytes = (stvl_get_pool_stats(_grid, &ps) == STVL_OK) ? (double)ps.device_bytes : (double)(m * 20);
    return f.good();
  } catch (...) {
    map_size_bytes = 0.;
    return false;
  }
  return false;  // best offense is a good defense
}

/*****************************************************************************/
bool SpatioTemporalVoxelGrid::LoadGrid(const std::string & file_name)
/*****************************************************************************/
{
  // reader of SaveGrid's flat binary: "STVLGRD1", voxel_size f64, n u64, then n x (ix,iy,iz i32, t f64)
  std::unique_lock<std::mutex> lock(_grid_lock);
  try {
    std::ifstream f(file_name + ".stvl", std::ios::binary);
    char magic[8];
    double vs = 0.;
    uint64_t n = 0;
    if (!f || !f.read(magic, 8) || std::memcmp(magic, "STVLGRD1", 8) != 0) {return false;}
    f.read(reinterpret_cast<char *>(&vs), 8);
    f.read(reinterpret_cast<char *>(&n), 8);
    if (!f || vs != stvl_voxel_size(_grid)) {
      std::cout << "Saved grid has another voxel size." << std::endl;
      return false;
    }
    std::vector<int32_t> ix(n), iy(n), iz(n);
    std::vector<double> t(n);
    for (uint64_t i = 0; i < n; ++i) {
      f.read(reinterpret_cast<char *>(&ix[i]), 4);
      f.read(reinterpret_cast<char *>(&iy[i]), 4);
      f.read(reinterpret_cast<char *>(&iz[i]), 4);
      f.read(reinterpret_cast<char *>(&t[i]), 8);
    }
    if (!f) {return false;}
    if (stvl_reset(_grid) != STVL_OK) {return false;}
    return n == 0 || stvl_load_voxels(_grid, ix.data(), iy.data(), iz.data(), t.data(), n) == STVL_OK;
  } catch (...) {
    return false;
  }
}

/*****************************************************************************/
bool SpatioTemporalVoxelGrid::FlattenToCostmap(
  double origin_x, double origin_y, double resolution, unsigned int size_x, unsigned int size_y, int mark_threshold,
  unsigned char default_value, unsigned char * costmap, stvl_costmap_result_t * result)
/*****************************************************************************/
{
  stvl_costmap_params_t p;
  std::memset(&p, 0, sizeof(p));
  p.origin_x = origin_x; p.origin_y = origin_y; p.resolution = resolution;
  p.size_x = size_x; p.size_y = size_y;
  p.mark_threshold = mark_threshold;
  p.default_value = default_value;
  p.lethal_value = 254;   // nav2_costmap_2d::LETHAL_OBSTACLE
  return stvl_costmap(_grid, &p, costmap, result) == STVL_OK;
}

/*****************************************************************************/
bool SpatioTemporalVoxelGrid::GetClearedBounds(double bounds[4])
/*****************************************************************************/
{
  stvl_sweep_stats_t st;
  if (stvl_get_sweep_stats(_grid, &st) != STVL_OK || st.n_cleared == 0) {
    return false;
  }
  for (int k = 0; k < 4; ++k) {bounds[k] = st.cleared_bbox_world[k];}
  return true;
}

/*****************************************************************************/
void SpatioTemporalVoxelGrid::SetExactClearedCells(bool on)
/*****************************************************************************/
{
  _exact_cleared_cells = on;
  stvl_set_outputs(_grid, (on ? (uint32_t)STVL_OUT_CLEARED : 0u) | (uint32_t)STVL_OUT_AMBIGUOUS);
}

}  // namespace volume_grid
