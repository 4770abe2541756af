// The subset of nav2_costmap_2d::{Costmap2D, Layer, CostmapLayer} that SpatioTemporalVoxelLayer uses
// (reference spatio_temporal_voxel_layer.cpp:666-725, 728-808, 950-963), restated from nav2's public behaviour:
// worldToMap / mapToWorld / getIndex / resetMaps / touch / updateWithMax / updateWithOverwrite / updateOrigin.
#ifndef STVL_SHIM_NAV2_COSTMAP_LAYER_HPP_
#define STVL_SHIM_NAV2_COSTMAP_LAYER_HPP_
#include <algorithm>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>
#include <cstdlib>
#include "geometry_msgs/msg/point.hpp"
#include "nav2_costmap_2d/cost_values.hpp"

namespace nav2_costmap_2d
{
class Costmap2D
{
public:
  typedef std::recursive_mutex mutex_t;
  Costmap2D() {}
  Costmap2D(unsigned int sx, unsigned int sy, double res, double ox, double oy, unsigned char def = 0)
  { resizeMap(sx, sy, res, ox, oy); default_value_ = def; resetMaps(); }
  virtual ~Costmap2D() {}
  void resizeMap(unsigned int sx, unsigned int sy, double res, double ox, double oy)
  {
    size_x_ = sx; size_y_ = sy; resolution_ = res; origin_x_ = ox; origin_y_ = oy;
    costmap_.assign((size_t)sx * sy, default_value_);
  }
  void resetMaps() { if (!costmap_.empty()) std::memset(costmap_.data(), default_value_, costmap_.size()); }
  bool worldToMap(double wx, double wy, unsigned int & mx, unsigned int & my) const
  {
    if (wx < origin_x_ || wy < origin_y_) return false;
    mx = static_cast<unsigned int>((wx - origin_x_) / resolution_);
    my = static_cast<unsigned int>((wy - origin_y_) / resolution_);
    return mx < size_x_ && my < size_y_;
  }
  void mapToWorld(unsigned int mx, unsigned int my, double & wx, double & wy) const
  {
    wx = origin_x_ + (mx + 0.5) * resolution_;
    wy = origin_y_ + (my + 0.5) * resolution_;
  }
  unsigned int getIndex(unsigned int mx, unsigned int my) const { return my * size_x_ + mx; }
  unsigned char * getCharMap() { return costmap_.data(); }
  const unsigned char * getCharMap() const { return costmap_.data(); }
  unsigned char getCost(unsigned int mx, unsigned int my) const { return costmap_[getIndex(mx, my)]; }
  void setCost(unsigned int mx, unsigned int my, unsigned char c) { costmap_[getIndex(mx, my)] = c; }
  unsigned int getSizeInCellsX() const { return size_x_; }
  unsigned int getSizeInCellsY() const { return size_y_; }
  double getSizeInMetersX() const { return (size_x_ - 1 + 0.5) * resolution_; }
  double getSizeInMetersY() const { return (size_y_ - 1 + 0.5) * resolution_; }
  double getOriginX() const { return origin_x_; }
  double getOriginY() const { return origin_y_; }
  double getResolution() const { return resolution_; }
  void setDefaultValue(unsigned char c) { default_value_ = c; }
  unsigned char getDefaultValue() const { return default_value_; }
  // rolling window: nav2 shifts the overlapping region; STVL refills the whole layer every cycle anyway
  virtual void updateOrigin(double new_origin_x, double new_origin_y)
  {
    const int cell_ox = static_cast<int>((new_origin_x - origin_x_) / resolution_);
    const int cell_oy = static_cast<int>((new_origin_y - origin_y_) / resolution_);
    origin_x_ = origin_x_ + cell_ox * resolution_;
    origin_y_ = origin_y_ + cell_oy * resolution_;
    resetMaps();
  }
  mutex_t * getMutex() { return &access_; }

  // setConvexPolygonCost / convexFillCells / polygonOutlineCells / raytraceLine / bresenham2D, as nav2 has them: map the
  // vertices (a vertex off the map: return false, nothing set), collect the Bresenham outline of every edge and of the
  // closing edge, sort the cells by x, fill every column between its lowest and highest outline cell.
  struct MapLocation { unsigned int x, y; };
  bool setConvexPolygonCost(const std::vector<geometry_msgs::msg::Point> & polygon, unsigned char cost_value)
  {
    std::vector<MapLocation> map_polygon;
    for (const auto & pt : polygon) {
      MapLocation loc;
      if (!worldToMap(pt.x, pt.y, loc.x, loc.y)) return false;
      map_polygon.push_back(loc);
    }
    std::vector<MapLocation> cells;
    convexFillCells(map_polygon, cells);
    for (const auto & c : cells) costmap_[getIndex(c.x, c.y)] = cost_value;
    return true;
  }
  void polygonOutlineCells(const std::vector<MapLocation> & polygon, std::vector<MapLocation> & cells) const
  {
    for (size_t i = 0; i + 1 < polygon.size(); ++i) raytraceLine(cells, polygon[i].x, polygon[i].y, polygon[i + 1].x, polygon[i + 1].y);
    if (!polygon.empty()) raytraceLine(cells, polygon.back().x, polygon.back().y, polygon[0].x, polygon[0].y);
  }
  void convexFillCells(const std::vector<MapLocation> & polygon, std::vector<MapLocation> & cells) const
  {
    if (polygon.size() < 3) return;
    polygonOutlineCells(polygon, cells);
    size_t i = 0;
    while (i + 1 < cells.size()) {   // "quick bubble sort to sort points by x"
      if (cells[i].x > cells[i + 1].x) { std::swap(cells[i], cells[i + 1]); if (i > 0) --i; } else { ++i; }
    }
    i = 0;
    MapLocation min_pt, max_pt;
    const unsigned int min_x = cells[0].x, max_x = cells[cells.size() - 1].x;
    for (unsigned int x = min_x; x <= max_x; ++x) {
      if (i + 1 >= cells.size()) break;
      if (cells[i].y < cells[i + 1].y) { min_pt = cells[i]; max_pt = cells[i + 1]; } else { min_pt = cells[i + 1]; max_pt = cells[i]; }
      i += 2;
      while (i < cells.size() && cells[i].x == x) {
        if (cells[i].y < min_pt.y) min_pt = cells[i]; else if (cells[i].y > max_pt.y) max_pt = cells[i];
        ++i;
      }
      for (unsigned int y = min_pt.y; y <= max_pt.y; ++y) cells.push_back(MapLocation{x, y});
    }
  }
  void raytraceLine(std::vector<MapLocation> & cells, unsigned int x0, unsigned int y0, unsigned int x1, unsigned int y1) const
  {
    const int dx = (int)x1 - (int)x0, dy = (int)y1 - (int)y0;
    const unsigned int abs_dx = (unsigned int)std::abs(dx), abs_dy = (unsigned int)std::abs(dy);
    const int offset_dx = dx > 0 ? 1 : -1, offset_dy = (dy > 0 ? 1 : -1) * (int)size_x_;
    const unsigned int offset = y0 * size_x_ + x0;
    if (abs_dx >= abs_dy) { bresenham2D(cells, abs_dx, abs_dy, (int)(abs_dx / 2), offset_dx, offset_dy, offset); return; }
    bresenham2D(cells, abs_dy, abs_dx, (int)(abs_dy / 2), offset_dy, offset_dx, offset);
  }
  void bresenham2D(std::vector<MapLocation> & cells, unsigned int abs_da, unsigned int abs_db, int error_b, int offset_a, int offset_b,
                   unsigned int offset) const
  {
    for (unsigned int i = 0; i < abs_da; ++i) {
      cells.push_back(MapLocation{offset % size_x_, offset / size_x_});
      offset += offset_a;
      error_b += abs_db;
      if ((unsigned int)error_b >= abs_da) { offset += offset_b; error_b -= abs_da; }
    }
    cells.push_back(MapLocation{offset % size_x_, offset / size_x_});
  }

protected:
  unsigned int size_x_ = 0, size_y_ = 0;
  double resolution_ = 0.05, origin_x_ = 0, origin_y_ = 0;
  std::vector<unsigned char> costmap_;
  unsigned char default_value_ = 0;
  mutex_t access_;
};

class LayeredCostmap
{
public:
  LayeredCostmap(const std::string & frame, bool rolling) : frame_(frame), rolling_(rolling) {}
  bool isRolling() const { return rolling_; }
  const std::string & getGlobalFrameID() const { return frame_; }
  Costmap2D * getCostmap() { return &master_; }
private:
  std::string frame_;
  bool rolling_;
  Costmap2D master_;
};

class Layer
{
public:
  virtual ~Layer() {}
  virtual void onInitialize() {}
  virtual void updateBounds(double robot_x, double robot_y, double robot_yaw, double * min_x, double * min_y, double * max_x, double * max_y) = 0;
  virtual void updateCosts(Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j) = 0;
  virtual void reset() = 0;
  virtual bool isClearable() = 0;
  virtual void matchSize() {}
  bool isCurrent() const { return current_; }
  const std::string & getName() const { return name_; }
protected:
  LayeredCostmap * layered_costmap_ = nullptr;
  std::string name_ = "stvl_layer";
  bool current_ = false;
  bool enabled_ = true;
};

class CostmapLayer : public Layer, public Costmap2D
{
public:
  virtual void matchSize()
  {
    if (!layered_costmap_) return;
    Costmap2D * m = layered_costmap_->getCostmap();
    resizeMap(m->getSizeInCellsX(), m->getSizeInCellsY(), m->getResolution(), m->getOriginX(), m->getOriginY());
  }
  virtual void clearArea(int start_x, int start_y, int end_x, int end_y, bool invert)
  {
    current_ = false;
    unsigned char * grid = getCharMap();
    for (int x = 0; x < static_cast<int>(getSizeInCellsX()); x++) {
      const bool xrange = x > start_x && x < end_x;
      for (int y = 0; y < static_cast<int>(getSizeInCellsY()); y++) {
        if ((xrange && y > start_y && y < end_y) == invert) continue;
        const int index = getIndex(x, y);
        if (grid[index] != NO_INFORMATION) grid[index] = NO_INFORMATION;
      }
    }
  }
  void addExtraBounds(double mx0, double my0, double mx1, double my1)
  {
    extra_min_x_ = std::min(mx0, extra_min_x_); extra_max_x_ = std::max(mx1, extra_max_x_);
    extra_min_y_ = std::min(my0, extra_min_y_); extra_max_y_ = std::max(my1, extra_max_y_);
    has_extra_bounds_ = true;
  }
protected:
  void touch(double x, double y, double * min_x, double * min_y, double * max_x, double * max_y)
  {
    *min_x = std::min(x, *min_x); *min_y = std::min(y, *min_y);
    *max_x = std::max(x, *max_x); *max_y = std::max(y, *max_y);
  }
  void useExtraBounds(double * min_x, double * min_y, double * max_x, double * max_y)
  {
    if (!has_extra_bounds_) return;
    *min_x = std::min(extra_min_x_, *min_x); *min_y = std::min(extra_min_y_, *min_y);
    *max_x = std::max(extra_max_x_, *max_x); *max_y = std::max(extra_max_y_, *max_y);
    extra_min_x_ = extra_min_y_ = 1e6; extra_max_x_ = extra_max_y_ = -1e6;
    has_extra_bounds_ = false;
  }
  void updateWithMax(Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j)
  {
    if (!enabled_) return;
    unsigned char * master = master_grid.getCharMap();
    const unsigned int span = master_grid.getSizeInCellsX();
    for (int j = min_j; j < max_j; j++) {
      unsigned int it = span * j + min_i;
      for (int i = min_i; i < max_i; i++, it++) {
        if (costmap_[it] == NO_INFORMATION) continue;
        const unsigned char old_cost = master[it];
        if (old_cost == NO_INFORMATION || old_cost < costmap_[it]) master[it] = costmap_[it];
      }
    }
  }
  void updateWithOverwrite(Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j)
  {
    if (!enabled_) return;
    unsigned char * master = master_grid.getCharMap();
    const unsigned int span = master_grid.getSizeInCellsX();
    for (int j = min_j; j < max_j; j++) {
      unsigned int it = span * j + min_i;
      for (int i = min_i; i < max_i; i++, it++) {
        if (costmap_[it] != NO_INFORMATION) master[it] = costmap_[it];
      }
    }
  }
  bool has_extra_bounds_ = false;
  double extra_min_x_ = 1e6, extra_max_x_ = -1e6, extra_min_y_ = 1e6, extra_max_y_ = -1e6;
};
}  // namespace nav2_costmap_2d
#endif
