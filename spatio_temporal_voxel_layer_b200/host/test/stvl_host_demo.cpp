// stvl_host_demo — drives the C++ plugin (SpatioTemporalVoxelLayer -> SpatioTemporalVoxelGrid -> C-ABI -> CUDA) through
// a scenario file written by tests/test_host_facade.py and dumps what a nav2 LayeredCostmap would observe per cycle.
//   stvl_host_demo <scenario.bin> <out.bin> [--literal]
// exit 0 = ran, 3 = no usable GPU (the layer has no CPU fallback), 2 = bad input.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <vector>

#include "spatio_temporal_voxel_layer/spatio_temporal_voxel_layer.hpp"

namespace
{
template<typename T> bool rd(std::ifstream & f, T & v) { return (bool)f.read(reinterpret_cast<char *>(&v), sizeof(T)); }
template<typename T> void wr(std::ofstream & f, const T & v) { f.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
}

int main(int argc, char ** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: %s scenario.bin out.bin [--literal]\n", argv[0]); return 2; }
  const bool literal = argc > 3 && std::string(argv[3]) == "--literal";
  std::ifstream in(argv[1], std::ios::binary);
  char magic[8];
  if (!in.read(magic, 8) || std::memcmp(magic, "STVLSCN1", 8) != 0) { std::fprintf(stderr, "bad scenario file\n"); return 2; }

  using namespace spatio_temporal_voxel_layer;
  LayerParameters lp;
  float vs; int32_t decay_model, mark_threshold, track_unknown, publish; double voxel_decay;
  rd(in, vs); rd(in, decay_model); rd(in, voxel_decay); rd(in, mark_threshold); rd(in, track_unknown); rd(in, publish);
  lp.voxel_size = vs; lp.decay_model = decay_model; lp.voxel_decay = voxel_decay; lp.mark_threshold = mark_threshold;
  lp.track_unknown_space = track_unknown != 0; lp.publish_voxel_map = publish != 0; lp.update_footprint_enabled = true;
  double ox, oy, res; uint32_t sx, sy;
  rd(in, ox); rd(in, oy); rd(in, res); rd(in, sx); rd(in, sy);
  uint32_t n_sources;
  rd(in, n_sources);
  for (uint32_t s = 0; s < n_sources; ++s) {
    SourceParameters sp;
    sp.name = "source" + std::to_string(s);
    double p[9]; int32_t q[4];
    for (double & v : p) rd(in, v);
    for (int32_t & v : q) rd(in, v);
    sp.obstacle_range = p[0]; sp.min_z = p[1]; sp.max_z = p[2]; sp.vertical_fov_angle = p[3]; sp.vertical_fov_padding = p[4];
    sp.horizontal_fov_angle = p[5]; sp.decay_acceleration = p[6]; sp.min_obstacle_height = p[7]; sp.max_obstacle_height = p[8];
    sp.marking = q[0] != 0; sp.clearing = q[1] != 0; sp.model_type = q[2];
    sp.filter = q[3] == 2 ? "passthrough" : (q[3] == 1 ? "voxel" : "none");
    lp.observation_sources.push_back(sp);
  }

  double now = 0.0;
  auto clock = std::make_shared<rclcpp::Clock>([&now]() { return now; });
  nav2_costmap_2d::LayeredCostmap parent("map", false);
  parent.getCostmap()->resizeMap(sx, sy, res, ox, oy);
  SpatioTemporalVoxelLayer layer;
  layer.configure(lp, clock, &parent);
  layer.setFusedCostmap(!literal);
  {
    // a 0.7 m x 0.5 m robot footprint; the robot pose of a cycle is the first source's origin, yaw = 0.3 * cycle
    std::vector<geometry_msgs::msg::Point> fp(4);
    const double hx = 0.35, hy = 0.25;
    fp[0].x = hx; fp[0].y = hy; fp[1].x = -hx; fp[1].y = hy; fp[2].x = -hx; fp[2].y = -hy; fp[3].x = hx; fp[3].y = -hy;
    layer.setFootprint(fp);
  }
  try {
    layer.onInitialize();
  } catch (const std::exception & e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }

  std::ofstream out(argv[2], std::ios::binary);
  out.write("STVLOUT1", 8);
  uint32_t n_cycles;
  rd(in, n_cycles);
  wr(out, n_cycles);
  for (uint32_t c = 0; c < n_cycles; ++c) {
    rd(in, now);
    double robot_x = 0.0, robot_y = 0.0;
    for (uint32_t s = 0; s < n_sources; ++s) {
      geometry_msgs::msg::Point o; geometry_msgs::msg::Quaternion q;
      rd(in, o.x); rd(in, o.y); rd(in, o.z); rd(in, q.x); rd(in, q.y); rd(in, q.z); rd(in, q.w);
      if (s == 0) { robot_x = o.x; robot_y = o.y; }
      uint32_t n_points;
      rd(in, n_points);
      auto cloud = std::make_shared<sensor_msgs::msg::PointCloud2>();
      cloud->width = n_points; cloud->height = 1; cloud->point_step = 16; cloud->row_step = 16 * n_points;
      const char * names[3] = {"x", "y", "z"};
      for (uint32_t i = 0; i < 3; ++i) {
        sensor_msgs::msg::PointField f; f.name = names[i]; f.offset = 4 * i; f.datatype = sensor_msgs::msg::PointField::FLOAT32; f.count = 1;
        cloud->fields.push_back(f);
      }
      cloud->data.resize((size_t)n_points * 16);
      in.read(reinterpret_cast<char *>(cloud->data.data()), (std::streamsize)cloud->data.size());
      layer.GetBuffer("source" + std::to_string(s))->BufferCloud(o, q, cloud);
    }
    double b[4] = {1e308, 1e308, -1e308, -1e308};
    layer.updateBounds(robot_x, robot_y, 0.3 * c, &b[0], &b[1], &b[2], &b[3]);
    // master grid merge: updateCosts = footprint clearing in the layer's costmap + updateWithMax over the whole map
    layer.updateCosts(*parent.getCostmap(), 0, 0, (int)sx, (int)sy);
    // what the cycle produced: costmap bytes (after the footprint clearing), bounds, flattened costmap, occupancy cloud size
    const uint32_t cells = sx * sy;
    wr(out, cells);
    out.write(reinterpret_cast<const char *>(layer.getCharMap()), cells);
    for (double v : b) wr(out, v);
    auto * cm = layer.Grid()->GetFlattenedCostmap();
    const uint64_t ncol = cm->size();
    wr(out, ncol);
    for (const auto & kv : *cm) { wr(out, kv.first.x); wr(out, kv.first.y); const uint32_t cnt = kv.second; wr(out, cnt); }
    const uint64_t npts = layer.LastVoxelCloud() ? layer.LastVoxelCloud()->width : 0;
    wr(out, npts);
  }
  // final state of the master costmap and the grid itself
  out.write(reinterpret_cast<const char *>(parent.getCostmap()->getCharMap()), (std::streamsize)sx * sy);
  double bytes = 0.;
  const bool saved = layer.SaveGrid(std::string(argv[2]) + ".grid", bytes);
  // round trip of the saved grid: load it back (replaces the store) and save it again
  double bytes2 = 0.;
  const bool reloaded = saved && layer.Grid()->LoadGrid(std::string(argv[2]) + ".grid") && layer.SaveGrid(std::string(argv[2]) + ".grid2", bytes2);
  const uint32_t ok = saved ? (reloaded ? 1u : 2u) : 0u;
  wr(out, ok);
  std::printf("stvl_host_demo: %u cycles, %s costmap path, grid saved=%u (%.0f device bytes)\n", n_cycles, literal ? "literal host" : "fused GPU", ok, bytes);
  return 0;
}
