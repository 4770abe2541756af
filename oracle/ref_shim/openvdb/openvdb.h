// Stand-in for the part of OpenVDB 10 the reference's SpatioTemporalVoxelGrid touches (reference
// spatio_temporal_voxel_grid.cpp:73-93, 149-224, 421-496).  TEST INFRASTRUCTURE: it exists only so that the reference's own
// translation units can be compiled unmodified under oracle/_ref/ (OpenVDB is not in this image and cannot be fetched).
// What matters numerically is the transform: createLinearTransform(scale matrix) simplifies to a UniformScaleMap whose
// applyMap is ijk * scale and whose applyInverseMap multiplies by the cached 1.0 / scale (OpenVDB math/Maps.h, ScaleMap:
// mScaleValuesInverse = 1 / mScaleValues); everything else is container semantics (active voxels with a double value,
// background value, iteration over the active set — here in key order, the reference's results do not depend on it).
#ifndef STVL_REF_SHIM_OPENVDB_H_
#define STVL_REF_SHIM_OPENVDB_H_
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

namespace openvdb
{
typedef double Real;

struct Vec3d
{
  double v[3];
  Vec3d() : v{0., 0., 0.} {}
  Vec3d(double x, double y, double z) : v{x, y, z} {}
  double & operator[](int i) { return v[i]; }
  const double & operator[](int i) const { return v[i]; }
  double & x() { return v[0]; }
  double & y() { return v[1]; }
  double & z() { return v[2]; }
  const double & x() const { return v[0]; }
  const double & y() const { return v[1]; }
  const double & z() const { return v[2]; }
};

struct Coord
{
  int32_t c[3];
  Coord() : c{0, 0, 0} {}
  // the reference constructs Coord(double, double, double): the doubles convert to Int32 by truncation toward zero
  Coord(int32_t x, int32_t y, int32_t z) : c{x, y, z} {}
  int32_t x() const { return c[0]; }
  int32_t y() const { return c[1]; }
  int32_t z() const { return c[2]; }
  bool operator<(const Coord & o) const { return std::tie(c[0], c[1], c[2]) < std::tie(o.c[0], o.c[1], o.c[2]); }
};

namespace math
{
enum Axis { X_AXIS = 0, Y_AXIS = 1, Z_AXIS = 2 };

struct Mat4d
{
  double m[4][4];
  static Mat4d identity() { Mat4d r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1. : 0.; return r; }
  // the reference pre-scales an identity by (vs, vs, vs), pre-translates by 0 and pre-rotates by 0 about Z: the matrix is
  // diag(vs, vs, vs, 1) exactly (the rotation by angle 0 multiplies by cos 0 = 1 and adds sin 0 = 0 terms)
  void preScale(const Vec3d & s) { for (int j = 0; j < 4; ++j) { m[0][j] *= s[0]; m[1][j] *= s[1]; m[2][j] *= s[2]; } }
  void preTranslate(const Vec3d & t) { for (int j = 0; j < 4; ++j) m[3][j] += t[0] * m[0][j] + t[1] * m[1][j] + t[2] * m[2][j]; }
  void preRotate(Axis, double angle) { (void)angle; /* angle == 0 at the only call site: identity */ }
};

class Transform
{
public:
  typedef std::shared_ptr<Transform> Ptr;
  static Ptr createLinearTransform(const Mat4d & m)
  {
    Ptr t(new Transform);
    t->scale_ = m.m[0][0];              // uniform scale (the call site passes diag(vs, vs, vs, 1))
    t->inv_scale_ = 1.0 / t->scale_;    // ScaleMap caches the inverse
    return t;
  }
  Vec3d indexToWorld(const Coord & c) const { return Vec3d((double)c.c[0] * scale_, (double)c.c[1] * scale_, (double)c.c[2] * scale_); }
  Vec3d worldToIndex(const Vec3d & p) const { return Vec3d(p[0] * inv_scale_, p[1] * inv_scale_, p[2] * inv_scale_); }
private:
  double scale_ = 1., inv_scale_ = 1.;
};

template <typename T>
struct Ray { typedef Vec3d Vec3T; };
}  // namespace math
typedef math::Mat4d Mat4d;

enum GridClass { GRID_UNKNOWN = 0, GRID_LEVEL_SET = 1 };
struct FloatMetadata { explicit FloatMetadata(float v) : value(v) {} float value; };
inline void initialize() {}

class DoubleGrid
{
public:
  typedef std::shared_ptr<DoubleGrid> Ptr;
  struct Voxel { double value; bool on; };
  typedef std::map<Coord, Voxel> Store;

  static Ptr create(double background) { Ptr g(new DoubleGrid); g->background_ = background; return g; }
  void setTransform(math::Transform::Ptr t) { xform_ = t; }
  void setName(const std::string &) {}
  void insertMeta(const std::string &, const FloatMetadata &) {}
  void setGridClass(GridClass) {}
  Vec3d indexToWorld(const Coord & c) const { return xform_->indexToWorld(c); }
  Vec3d worldToIndex(const Vec3d & p) const { return xform_->worldToIndex(p); }

  class Accessor
  {
  public:
    explicit Accessor(DoubleGrid * g) : g_(g) {}
    void setValueOn(const Coord & c, double v) { Voxel & x = g_->store_[c]; x.value = v; x.on = true; }
    void setValueOff(const Coord & c, double v) { Voxel & x = g_->store_[c]; x.value = v; x.on = false; }
    double getValue(const Coord & c) const { auto it = g_->store_.find(c); return it == g_->store_.end() ? g_->background_ : it->second.value; }
    bool isValueOn(const Coord & c) const { auto it = g_->store_.find(c); return it != g_->store_.end() && it->second.on; }
  private:
    DoubleGrid * g_;
  };
  Accessor getAccessor() { return Accessor(this); }

  // iteration over the active voxels; voxels switched off or rewritten during the walk stay in the container (std::map
  // iterators are stable), exactly like OpenVDB's value iterator over an unchanged tree topology
  class ValueOnCIter
  {
  public:
    ValueOnCIter(Store::const_iterator it, Store::const_iterator end) : it_(it), end_(end) { skip(); }
    bool test() const { return it_ != end_; }
    ValueOnCIter & operator++() { ++it_; skip(); return *this; }
    Coord getCoord() const { return it_->first; }
    double getValue() const { return it_->second.value; }
  private:
    void skip() { while (it_ != end_ && !it_->second.on) ++it_; }
    Store::const_iterator it_, end_;
  };
  ValueOnCIter cbeginValueOn() const { return ValueOnCIter(store_.begin(), store_.end()); }

  // prune(): inactive voxels holding the background value are dropped (per-voxel semantics; OpenVDB additionally collapses
  // constant regions into tiles, which does not change any voxel's value or state)
  void pruneGrid() { for (auto it = store_.begin(); it != store_.end();) { if (!it->second.on) it = store_.erase(it); else ++it; } }
  void clear() { store_.clear(); }
  bool empty() const { for (const auto & kv : store_) if (kv.second.on) return false; return true; }
  uint64_t memUsage() const { return (uint64_t)store_.size() * (sizeof(Coord) + sizeof(Voxel)); }
  const Store & store() const { return store_; }
private:
  Store store_;
  double background_ = 0.;
  math::Transform::Ptr xform_;
};
typedef std::vector<DoubleGrid::Ptr> GridPtrVec;

namespace io
{
// SaveGrid (reference .cpp:480-496) writes a .vdb; the stand-in accepts the call and writes nothing
class File
{
public:
  explicit File(const std::string &) {}
  void write(const GridPtrVec &) {}
  void close() {}
};
}  // namespace io
}  // namespace openvdb
#endif
