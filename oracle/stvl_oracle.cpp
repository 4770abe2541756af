// stvl_oracle.cpp — CPU restatement of the reference's per-cycle voxel update.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (spatio_temporal_voxel_layer_b200/,
// include/) links, imports or executes this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may.
//
// PINNING.  The reference ships no tests, golden vectors or fixtures and its third-party
// dependencies (OpenVDB 10.0.1, Eigen3, PCL, ROS 2 / nav2) are absent here with no network, so the real binary cannot be
// built.  What IS done (round 2): the reference's OWN translation units of the path — spatio_temporal_voxel_grid.cpp,
// depth_camera_frustum.cpp, three_dimensional_lidar_frustum.cpp — are compiled UNMODIFIED from /root/reference against
// stand-in headers for those dependencies (oracle/ref_shim/, oracle/ref_driver.cpp -> oracle/_ref/libstvl_ref.so), and
// tests/test_ref_pin.py checks this restatement against that build bit for bit: frustum plane sets, IsInside decisions
// (points on the planes included), the quantisation quirks of operator(), the decay algebra, multi-cycle scenes in every
// decay model, cleared cells, published points, ResetGridArea.  tests/golden/*.npz are certified the same way
// (tests/golden/make_ref_golden.py, ref_certified.json).  So the reference's own logic is pinned; STILL UNPINNED is the
// third-party arithmetic listed below, which both this file and the stand-in headers restate from published behaviour
// (Eigen's evaluation order, OpenVDB's scale map, nav2's worldToMap, PCL's filters).  The hand-derived known-answer tests
// of round 1 remain in tests/test_oracle_kat.py.
//
// Every function cites the reference lines it follows.  Paths are relative to
// /root/reference/spatio_temporal_voxel_layer/ :
//   GRID   = src/spatio_temporal_voxel_grid.cpp
//   GRIDH  = include/spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp
//   DEPTH  = src/frustum_models/depth_camera_frustum.cpp
//   LIDAR  = src/frustum_models/three_dimensional_lidar_frustum.cpp
//   FRUST  = include/spatio_temporal_voxel_layer/frustum_models/frustum.hpp
//   LAYER  = src/spatio_temporal_voxel_layer.cpp
//   BUFFER = src/measurement_buffer.cpp
//
// Third-party arithmetic that is NOT in /root/reference and is restated from its
// published behaviour:
//   OpenVDB v10.0.1 (pinned by openvdb_vendor/CMakeLists.txt:9): DoubleGrid as a
//     key->double container; UniformScaleMap: indexToWorld = ijk * vs,
//     worldToIndex = xyz * (1/vs)  (GRID:446-469).
//   Eigen3 (unpinned): AngleAxisd::toRotationMatrix, Affine3d products, cross,
//     normalize (divide by sqrt(squaredNorm) when > 0), Quaterniond::toRotationMatrix,
//     inverse, conjugate, quaternion*vector.  squaredNorm of a Vector3d is taken as
//     (x*x + y*y) + z*z (the SSE2 Packet2d reduction of an x86-64 build);
//     Affine3d::rotation() (an SVD polar factor) is taken as linear().
//   nav2_costmap_2d (unpinned): Costmap2D::worldToMap / getIndex / touch / resetMaps.
//   PCL (unpinned): PassThrough and VoxelGrid on PCLPointCloud2.
//   glibc libm: exp, tan, atan.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: no FMA contraction, as in a
// baseline x86-64 build of the reference).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <unordered_map>
#include <unordered_set>
#include <algorithm>
#include <functional>

namespace {

// ---------------------------------------------------------------------------------
// Minimal fixed-size linear algebra in Eigen's evaluation order
// ---------------------------------------------------------------------------------
struct V3 { double x, y, z; };
struct M3 { double m[3][3]; };

inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 sub(const V3 & a, const V3 & b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 add(const V3 & a, const V3 & b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 scale(const V3 & a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
// Eigen cross: (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
inline V3 cross(const V3 & a, const V3 & b)
{
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen Vector3d::squaredNorm on x86-64/SSE2: predux([x2,y2]) then + z2
inline double squared_norm(const V3 & a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
// Eigen normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)
inline void normalize(V3 & a)
{
  const double z = squared_norm(a);
  if (z > 0.0) {
    const double n = std::sqrt(z);
    a.x /= n; a.y /= n; a.z /= n;
  }
}
inline V3 mul(const M3 & R, const V3 & v)
{
  return v3((R.m[0][0] * v.x + R.m[0][1] * v.y) + R.m[0][2] * v.z,
            (R.m[1][0] * v.x + R.m[1][1] * v.y) + R.m[1][2] * v.z,
            (R.m[2][0] * v.x + R.m[2][1] * v.y) + R.m[2][2] * v.z);
}
inline M3 mul(const M3 & A, const M3 & B)
{
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i][j] = (A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j]) + A.m[i][2] * B.m[2][j];
  return C;
}
// Eigen AngleAxis<double>::toRotationMatrix()
inline M3 angle_axis(double angle, const V3 & axis)
{
  M3 res;
  const V3 sin_axis = scale(axis, std::sin(angle));
  const double c = std::cos(angle);
  const V3 cos1_axis = scale(axis, 1.0 - c);
  double tmp;
  tmp = cos1_axis.x * axis.y;
  res.m[0][1] = tmp - sin_axis.z;
  res.m[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  res.m[0][2] = tmp + sin_axis.y;
  res.m[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  res.m[1][2] = tmp - sin_axis.x;
  res.m[2][1] = tmp + sin_axis.x;
  res.m[0][0] = cos1_axis.x * axis.x + c;
  res.m[1][1] = cos1_axis.y * axis.y + c;
  res.m[2][2] = cos1_axis.z * axis.z + c;
  return res;
}
struct Quat { double x, y, z, w; };
// Eigen QuaternionBase::toRotationMatrix()
inline M3 quat_to_matrix(const Quat & q)
{
  M3 res;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  res.m[0][0] = 1.0 - (tyy + tzz);
  res.m[0][1] = txy - twz;
  res.m[0][2] = txz + twy;
  res.m[1][0] = txy + twz;
  res.m[1][1] = 1.0 - (txx + tzz);
  res.m[1][2] = tyz - twx;
  res.m[2][0] = txz - twy;
  res.m[2][1] = tyz + twx;
  res.m[2][2] = 1.0 - (txx + tyy);
  return res;
}
// Eigen QuaternionBase::_transformVector: uv = vec x v; uv += uv; v + w*uv + vec x uv
inline V3 quat_rotate(const Quat & q, const V3 & v)
{
  const V3 qv = v3(q.x, q.y, q.z);
  V3 uv = cross(qv, v);
  uv = add(uv, uv);
  const V3 c2 = cross(qv, uv);
  return v3((v.x + q.w * uv.x) + c2.x, (v.y + q.w * uv.y) + c2.y, (v.z + q.w * uv.z) + c2.z);
}
// Eigen quaternion inverse(): conjugate / squaredNorm when squaredNorm > 0;
// 4-coefficient squaredNorm on SSE2: [x2+z2, y2+w2] then the horizontal add.
inline Quat quat_inverse(const Quat & q)
{
  const double n2 = (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w);
  if (n2 > 0.0) return Quat{-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
  return Quat{0.0, 0.0, 0.0, 0.0};
}

// ---------------------------------------------------------------------------------
// Frustum blocks (layout identical to stvl_frustum_t in include/stvl_b200.h; the
// oracle keeps its own definition so that it does not depend on product headers)
// ---------------------------------------------------------------------------------
struct FrustumBlock {
  int32_t type;       // 0 depth camera, 1 3-D lidar, -1 invalid   (measurement_reading.h:49-53)
  int32_t reserved;
  double accel_factor;
  double p[36];
};
static_assert(sizeof(FrustumBlock) == 304, "layout must match stvl_frustum_t");

// DEPTH:67-157 ComputePlaneNormals, in the sensor frame (+Z forward).
// out: 6 x {nx,ny,nz,px,py,pz} in the stored order top, left, bottom, right, far, near.
bool depth_compute_plane_normals(double vFOV, double hFOV, double min_d, double max_d, double * out)
{
  if (vFOV == 0 && hFOV == 0) {  // DEPTH:71-74
    return false;
  }
  const V3 X = v3(1, 0, 0), Y = v3(0, 1, 0);
  V3 deflected[4];
  // DEPTH:79-95 — rx * ry * Z is the third column of Rx*Ry
  M3 rx = angle_axis(vFOV / 2., X);
  M3 ry = angle_axis(hFOV / 2., Y);
  auto third_col = [](const M3 & A, const M3 & B) {
      const M3 C = mul(A, B);
      return v3(C.m[0][2], C.m[1][2], C.m[2][2]);
    };
  deflected[0] = third_col(rx, ry);
  rx = angle_axis(-vFOV / 2., X);
  deflected[1] = third_col(rx, ry);
  ry = angle_axis(-hFOV / 2., Y);
  deflected[2] = third_col(rx, ry);
  rx = angle_axis(vFOV / 2., X);
  deflected[3] = third_col(rx, ry);

  // DEPTH:98-104
  V3 pt[8];
  for (int i = 0; i < 4; ++i) {
    pt[2 * i] = scale(deflected[i], min_d);
    pt[2 * i + 1] = scale(deflected[i], max_d);
  }
  auto store = [&](int k, const V3 & n, const V3 & p) {
      out[6 * k + 0] = n.x; out[6 * k + 1] = n.y; out[6 * k + 2] = n.z;
      out[6 * k + 3] = p.x; out[6 * k + 4] = p.y; out[6 * k + 5] = p.z;
    };
  // DEPTH:109-115 top
  const V3 v_01 = sub(pt[1], pt[0]);
  const V3 v_13 = sub(pt[3], pt[1]);
  V3 T_n = cross(v_13, v_01); normalize(T_n); store(0, T_n, pt[0]);
  // DEPTH:117-123 left
  const V3 v_23 = sub(pt[3], pt[2]);
  const V3 v_35 = sub(pt[5], pt[3]);
  V3 T_l = cross(v_35, v_23); normalize(T_l); store(1, T_l, pt[2]);
  // DEPTH:125-131 bottom
  const V3 v_45 = sub(pt[5], pt[4]);
  const V3 v_57 = sub(pt[7], pt[5]);
  V3 T_b = cross(v_57, v_45); normalize(T_b); store(2, T_b, pt[4]);
  // DEPTH:133-139 right
  const V3 v_67 = sub(pt[7], pt[6]);
  const V3 v_71 = sub(pt[1], pt[7]);
  V3 T_r = cross(v_71, v_67); normalize(T_r); store(3, T_r, pt[6]);
  // DEPTH:142-144 far
  V3 T_1 = cross(v_57, v_71); normalize(T_1); store(4, T_1, pt[7]);
  // DEPTH:147-149 near = VectorWithPt3D(T_1, pt_[2]) * -1  (FRUST:76-79: a*x, point kept)
  store(5, v3(-1 * T_1.x, -1 * T_1.y, -1 * T_1.z), pt[2]);
  return true;
}

// DEPTH:160-174 TransformModel + FRUST:83-89 TransformFrames.
// T = Identity; T.pretranslate(q^-1 * p); T.prerotate(q)  =>  linear = R(q), translation = R(q)*(q^-1*p)
void depth_transform_model(const V3 & position, const Quat & q, double * planes)
{
  const V3 t0 = quat_rotate(quat_inverse(q), position);
  const M3 R = quat_to_matrix(q);
  const V3 translation = mul(R, t0);
  for (int k = 0; k < 6; ++k) {
    double * pl = planes + 6 * k;
    V3 vec_t = mul(R, v3(pl[0], pl[1], pl[2]));   // rotation() taken as linear()
    normalize(vec_t);
    const V3 p = add(mul(R, v3(pl[3], pl[4], pl[5])), translation);
    pl[0] = vec_t.x; pl[1] = vec_t.y; pl[2] = vec_t.z;
    pl[3] = p.x; pl[4] = p.y; pl[5] = p.z;
  }
}

// DEPTH:286-304 IsInside + DEPTH:322-339 Dot
bool depth_is_inside(const FrustumBlock & f, const V3 & pt)
{
  if (f.type != 0) return false;  // _valid_frustum == false, DEPTH:289-291
  for (int k = 0; k < 6; ++k) {
    const double * pl = f.p + 6 * k;
    V3 p_delta = v3(pt.x - pl[3], pt.y - pl[4], pt.z - pl[5]);
    normalize(p_delta);
    const double dot = pl[0] * p_delta.x + pl[1] * p_delta.y + pl[2] * p_delta.z;
    if (dot > 0.) return false;
  }
  return true;
}

// LIDAR:44-60 ctor precompute + LIDAR:70-74 TransformModel
void lidar_build(double vFOV, double vFOVPadding, double hFOV, double min_d, double max_d,
                 const V3 & position, const Quat & q, double * p)
{
  const double hFOVhalf = hFOV / 2.0;
  const double tan_vFOVhalf = std::tan(vFOV / 2.0);
  p[0] = position.x; p[1] = position.y; p[2] = position.z;
  p[3] = -q.x; p[4] = -q.y; p[5] = -q.z; p[6] = q.w;    // conjugate
  p[7] = vFOVPadding;
  p[8] = tan_vFOVhalf * tan_vFOVhalf;
  p[9] = min_d * min_d;
  p[10] = max_d * max_d;
  p[11] = hFOVhalf;
  p[12] = (hFOV > 6.27) ? 1.0 : 0.0;
  for (int i = 13; i < 36; ++i) p[i] = 0.0;
}

// LIDAR:77-116 IsInside
bool lidar_is_inside(const FrustumBlock & f, const V3 & pt)
{
  const double * p = f.p;
  const Quat qc{p[3], p[4], p[5], p[6]};
  const V3 tp = quat_rotate(qc, sub(pt, v3(p[0], p[1], p[2])));
  const double radial_distance_squared = (tp.x * tp.x) + (tp.y * tp.y);
  if (radial_distance_squared > p[10] || radial_distance_squared < p[9]) return false;
  const double v_padded = std::fabs(tp.z) + p[7];
  if ((v_padded * v_padded / radial_distance_squared) > p[8]) return false;
  if (p[12] == 0.0) {
    const double half_pi = M_PI / 2;
    if (tp.x > 0) {
      if (std::fabs(std::atan(tp.y / tp.x)) > p[11]) return false;
    } else if (std::fabs(std::atan(tp.x / tp.y)) + half_pi > p[11]) {
      return false;
    }
  }
  return true;
}

bool frustum_is_inside(const FrustumBlock & f, const V3 & pt)
{
  if (f.type == 0) return depth_is_inside(f, pt);
  if (f.type == 1) return lidar_is_inside(f, pt);
  return false;
}

// ---------------------------------------------------------------------------------
// Voxel container standing in for openvdb::DoubleGrid: open-addressing map from the
// packed (ix,iy,iz) key to the mark time.  Container semantics only (GRID:421-443).
// ---------------------------------------------------------------------------------
struct Coord { int32_t x, y, z; };

class VoxelMap {
public:
  VoxelMap() { rehash(1u << 16); }
  size_t size() const { return n_; }
  void clear() { std::fill(used_.begin(), used_.end(), 0); n_ = 0; }
  void set(const Coord & c, double v)   // setValueOn: insert or overwrite
  {
    if ((n_ + 1) * 2 > cap_) rehash(cap_ * 2);
    size_t i = slot(c);
    while (used_[i]) {
      if (same(keys_[i], c)) { vals_[i] = v; return; }
      i = (i + 1) & (cap_ - 1);
    }
    used_[i] = 1; keys_[i] = c; vals_[i] = v; ++n_;
  }
  void erase(const Coord & c)   // setValueOff(background): backward-shift deletion
  {
    size_t i = slot(c);
    while (used_[i]) {
      if (same(keys_[i], c)) break;
      i = (i + 1) & (cap_ - 1);
    }
    if (!used_[i]) return;
    size_t j = i;
    for (;; ) {
      j = (j + 1) & (cap_ - 1);
      if (!used_[j]) break;
      const size_t k = slot(keys_[j]);
      // can the entry at j move into the hole at i ?
      const bool between = (i <= j) ? (i < k && k <= j) : (i < k || k <= j);
      if (!between) {
        keys_[i] = keys_[j]; vals_[i] = vals_[j];
        i = j;
      }
    }
    used_[i] = 0; --n_;
  }
  // iteration over raw slots (order is unspecified, as with any container change;
  // every reference observable downstream is order-free: sets, counts, maps)
  size_t capacity() const { return cap_; }
  bool used(size_t i) const { return used_[i] != 0; }
  const Coord & key(size_t i) const { return keys_[i]; }
  double & val(size_t i) { return vals_[i]; }

private:
  static bool same(const Coord & a, const Coord & b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
  size_t slot(const Coord & c) const
  {
    uint64_t h = (uint64_t)(uint32_t)c.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)c.y + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= ((uint64_t)(uint32_t)c.z + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    return (size_t)h & (cap_ - 1);
  }
  void rehash(size_t cap)
  {
    std::vector<Coord> ok; std::vector<double> ov; std::vector<uint8_t> ou;
    ok.swap(keys_); ov.swap(vals_); ou.swap(used_);
    cap_ = cap; n_ = 0;
    keys_.assign(cap_, Coord{0, 0, 0}); vals_.assign(cap_, 0.0); used_.assign(cap_, 0);
    for (size_t i = 0; i < ou.size(); ++i) if (ou[i]) set(ok[i], ov[i]);
  }
  std::vector<Coord> keys_;
  std::vector<double> vals_;
  std::vector<uint8_t> used_;
  size_t cap_ = 0, n_ = 0;
};

// GRIDH:89-102 occupany_cell and GRIDH:196-203 its hash
struct occupany_cell {
  occupany_cell(const double & _x, const double & _y) : x(_x), y(_y) {}
  bool operator==(const occupany_cell & other) const { return x == other.x && y == other.y; }
  double x, y;
};
struct occupany_cell_hash {
  std::size_t operator()(const occupany_cell & k) const
  {
    return (std::hash<double>()(k.x) ^ (std::hash<double>()(k.y) << 1)) >> 1;
  }
};
struct Point32 { float x, y, z; };

struct OracleGrid {
  int decay_model;
  double background_value, voxel_size, voxel_decay;
  bool pub_voxels;
  double inv_voxel_size;  // OpenVDB ScaleMap caches 1/scale
  VoxelMap grid;
  std::vector<Point32> grid_points;                                    // GRIDH:186
  std::unordered_map<occupany_cell, unsigned, occupany_cell_hash> cost_map;   // GRIDH:187
  std::unordered_set<occupany_cell, occupany_cell_hash> cleared_cells;        // LAYER:764 (caller-owned upstream)

  // GRID:446-460 IndexToWorld: OpenVDB indexToWorld (coord * vs) then + vs/2
  V3 IndexToWorld(const Coord & c) const
  {
    V3 w = v3((double)c.x * voxel_size, (double)c.y * voxel_size, (double)c.z * voxel_size);
    const double center_offset = voxel_size / 2.0;
    w.x += center_offset; w.y += center_offset; w.z += center_offset;
    return w;
  }
  // GRID:463-469 WorldToIndex: OpenVDB worldToIndex = xyz * (1/vs)
  V3 WorldToIndex(const V3 & v) const
  {
    return v3(v.x * inv_voxel_size, v.y * inv_voxel_size, v.z * inv_voxel_size);
  }
  // GRID:320-331
  double GetTemporalClearingDuration(const double & time_delta) const
  {
    if (decay_model == 0) {
      return voxel_decay - time_delta;
    } else if (decay_model == 1) {
      return voxel_decay * std::exp(-time_delta);
    }
    return voxel_decay;
  }
  // GRID:334-341
  static double GetFrustumAcceleration(const double & time_delta, const double & acceleration_factor)
  {
    const double acceleration = 1. / 6. * acceleration_factor * (time_delta * time_delta * time_delta);
    return acceleration;
  }
  // GRID:227-251
  void PopulateCostmapAndPointcloud(const Coord & pt)
  {
    const V3 pose_world = IndexToWorld(pt);
    if (pub_voxels) {
      Point32 point;
      point.x = (float)pose_world.x; point.y = (float)pose_world.y; point.z = (float)pose_world.z;
      grid_points.push_back(point);
    }
    auto cell = cost_map.find(occupany_cell(pose_world.x, pose_world.y));
    if (cell != cost_map.end()) {
      cell->second += 1;
    } else {
      cost_map.insert(std::make_pair(occupany_cell(pose_world.x, pose_world.y), 1u));
    }
  }
  // GRID:149-224 TemporalClearAndGenerateCostmap (cur_time injected, GRID:155)
  void TemporalClearAndGenerateCostmap(const FrustumBlock * frustums, int n_frustums, double cur_time)
  {
    std::vector<Coord> to_clear;   // erasures are applied after the walk; decisions are per-voxel independent
    for (size_t s = 0; s < grid.capacity(); ++s) {
      if (!grid.used(s)) continue;
      const Coord pt_index = grid.key(s);
      const V3 pose_world = IndexToWorld(pt_index);
      bool frustum_cycle = false;
      bool cleared_point = false;
      const double value = grid.val(s);
      const double time_since_marking = cur_time - value;
      const double base_duration_to_decay = GetTemporalClearingDuration(time_since_marking);
      for (int f = 0; f < n_frustums; ++f) {
        if (frustum_is_inside(frustums[f], pose_world)) {
          frustum_cycle = true;
          const double frustum_acceleration = GetFrustumAcceleration(time_since_marking, frustums[f].accel_factor);
          const double time_until_decay = base_duration_to_decay - frustum_acceleration;
          if (time_until_decay < 0.) {
            cleared_point = true;
            to_clear.push_back(pt_index);
            break;
          } else {
            const double updated_mark = value - frustum_acceleration;
            grid.val(s) = updated_mark;   // MarkGridPoint: setValueOn overwrite, GRID:421-430
            break;
          }
        }
      }
      if (!frustum_cycle) {
        if (base_duration_to_decay < 0.) {
          cleared_point = true;
          to_clear.push_back(pt_index);
        }
      }
      if (cleared_point) {
        cleared_cells.insert(occupany_cell(pose_world.x, pose_world.y));
      } else {
        PopulateCostmapAndPointcloud(pt_index);
      }
    }
    for (const Coord & c : to_clear) grid.erase(c);   // ClearGridPoint, GRID:433-443; pruneGrid GRID:223
  }
  // GRID:96-146 ClearFrustums.  Frustum objects are built by the caller (blocks).
  void ClearFrustums(const FrustumBlock * frustums, int n_frustums, double cur_time)
  {
    cleared_cells.clear();   // upstream: a fresh set per updateBounds, LAYER:764
    if (grid.size() == 0) {  // IsGridEmpty, GRID:104-108
      grid_points.clear();
      cost_map.clear();
      return;
    }
    grid_points.clear();
    cost_map.clear();
    TemporalClearAndGenerateCostmap(frustums, n_frustums, cur_time);
  }
  // GRID:269-309 operator()(MeasurementReading) with cur_time injected (GRID:275)
  void MarkReading(const uint8_t * data, uint64_t n_points, uint32_t point_step,
                   uint32_t off_x, uint32_t off_y, uint32_t off_z,
                   const double origin[3], double obstacle_range, double marking, double cur_time,
                   uint64_t * n_nonfinite)
  {
    if (!marking) return;
    float mark_range_2 = obstacle_range * obstacle_range;
    for (uint64_t i = 0; i < n_points; ++i) {
      float fx, fy, fz;
      std::memcpy(&fx, data + i * point_step + off_x, 4);
      std::memcpy(&fy, data + i * point_step + off_y, 4);
      std::memcpy(&fz, data + i * point_step + off_z, 4);
      // upstream converts NaN/Inf to int (UB); we drop non-finite points and count them
      if (!std::isfinite(fx) || !std::isfinite(fy) || !std::isfinite(fz)) {
        if (n_nonfinite) ++*n_nonfinite;
        continue;
      }
      float distance_2 =
        (fx - origin[0]) * (fx - origin[0]) +
        (fy - origin[1]) * (fy - origin[1]) +
        (fz - origin[2]) * (fz - origin[2]);
      if (distance_2 > mark_range_2 || distance_2 < 0.0001) {
        continue;
      }
      double x = fx < 0 ? fx - voxel_size : fx;
      double y = fy < 0 ? fy - voxel_size : fy;
      double z = fy < 0 ? fz - voxel_size : fz;   // sic: z tests y's sign, GRID:295
      const V3 mark_grid = WorldToIndex(v3(x, y, z));
      // openvdb::Coord(double,double,double) -> Int32 truncation toward zero, GRID:300-303
      grid.set(Coord{(int32_t)mark_grid.x, (int32_t)mark_grid.y, (int32_t)mark_grid.z}, cur_time);
    }
  }
  // GRID:397-418 ResetGridArea
  void ResetGridArea(double sx, double sy, double ex, double ey, bool invert_area)
  {
    std::vector<Coord> to_clear;
    for (size_t s = 0; s < grid.capacity(); ++s) {
      if (!grid.used(s)) continue;
      const Coord pt_index = grid.key(s);
      const V3 pose_world = IndexToWorld(pt_index);
      const bool in_x_range = pose_world.x > sx && pose_world.x < ex;
      const bool in_y_range = pose_world.y > sy && pose_world.y < ey;
      const bool in_range = in_x_range && in_y_range;
      if (in_range == invert_area) to_clear.push_back(pt_index);
    }
    for (const Coord & c : to_clear) grid.erase(c);
  }
};

// nav2_costmap_2d::Costmap2D::worldToMap (external, restated)
inline bool worldToMap(double wx, double wy, double origin_x, double origin_y, double resolution,
                       unsigned size_x, unsigned size_y, unsigned & mx, unsigned & my)
{
  if (wx < origin_x || wy < origin_y) return false;
  mx = static_cast<unsigned int>((wx - origin_x) / resolution);
  my = static_cast<unsigned int>((wy - origin_y) / resolution);
  if (mx < size_x && my < size_y) return true;
  return false;
}
// nav2_costmap_2d::Layer::touch
inline void touch(double x, double y, double * b)
{
  b[0] = std::min(x, b[0]); b[1] = std::min(y, b[1]);
  b[2] = std::max(x, b[2]); b[3] = std::max(y, b[3]);
}

}  // namespace

// =====================================================================================
// C interface (ctypes / bench)
// =====================================================================================
extern "C" {

// GRID:49-93 ctor + InitializeGrid.  voxel_size arrives as float and is widened (GRID:51,54).
void * oracle_create(float voxel_size, double background_value, int decay_model, double voxel_decay, int pub_voxels)
{
  OracleGrid * g = new OracleGrid();
  g->decay_model = decay_model;
  g->background_value = background_value;
  g->voxel_size = voxel_size;
  g->voxel_decay = voxel_decay;
  g->pub_voxels = pub_voxels != 0;
  g->inv_voxel_size = 1.0 / g->voxel_size;
  return g;
}
void oracle_destroy(void * h) { delete static_cast<OracleGrid *>(h); }
double oracle_voxel_size(void * h) { return static_cast<OracleGrid *>(h)->voxel_size; }

// DEPTH ctor + SetPosition/SetOrientation + TransformModel; returns block.type
int oracle_build_depth_frustum(double vfov, double hfov, double min_d, double max_d,
                               const double origin[3], const double quat_xyzw[4], double accel, void * out)
{
  FrustumBlock * b = static_cast<FrustumBlock *>(out);
  std::memset(b, 0, sizeof(*b));
  b->accel_factor = accel;
  if (!depth_compute_plane_normals(vfov, hfov, min_d, max_d, b->p)) {
    b->type = -1;
    return b->type;
  }
  depth_transform_model(v3(origin[0], origin[1], origin[2]),
                        Quat{quat_xyzw[0], quat_xyzw[1], quat_xyzw[2], quat_xyzw[3]}, b->p);
  b->type = 0;
  return b->type;
}
int oracle_build_lidar_frustum(double vfov, double vpad, double hfov, double min_d, double max_d,
                               const double origin[3], const double quat_xyzw[4], double accel, void * out)
{
  FrustumBlock * b = static_cast<FrustumBlock *>(out);
  std::memset(b, 0, sizeof(*b));
  b->accel_factor = accel;
  lidar_build(vfov, vpad, hfov, min_d, max_d, v3(origin[0], origin[1], origin[2]),
              Quat{quat_xyzw[0], quat_xyzw[1], quat_xyzw[2], quat_xyzw[3]}, b->p);
  b->type = 1;
  return b->type;
}
// IsInside for n points (xyz doubles) against one block -> out[i] in {0,1}
void oracle_is_inside(const void * block, const double * xyz, uint64_t n, uint8_t * out)
{
  const FrustumBlock & b = *static_cast<const FrustumBlock *>(block);
  for (uint64_t i = 0; i < n; ++i) out[i] = frustum_is_inside(b, v3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])) ? 1 : 0;
}

void oracle_clear_frustums(void * h, double now, const void * blocks, int n_frustums)
{
  static_cast<OracleGrid *>(h)->ClearFrustums(static_cast<const FrustumBlock *>(blocks), n_frustums, now);
}
// GRID:254-266 Mark = operator() per reading; call once per reading, in order.
void oracle_mark(void * h, const void * data, uint64_t n_points, uint32_t point_step,
                 uint32_t off_x, uint32_t off_y, uint32_t off_z, const double origin[3],
                 double obstacle_range, double marking, double now, uint64_t * n_nonfinite)
{
  static_cast<OracleGrid *>(h)->MarkReading(static_cast<const uint8_t *>(data), n_points, point_step,
    off_x, off_y, off_z, origin, obstacle_range, marking, now, n_nonfinite);
}
uint64_t oracle_active_count(void * h) { return static_cast<OracleGrid *>(h)->grid.size(); }
uint64_t oracle_get_voxels(void * h, int32_t * ix, int32_t * iy, int32_t * iz, double * t, uint64_t cap)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  uint64_t n = 0;
  for (size_t s = 0; s < g->grid.capacity(); ++s) {
    if (!g->grid.used(s)) continue;
    if (n < cap) {
      const Coord & c = g->grid.key(s);
      if (ix) {ix[n] = c.x;}
      if (iy) {iy[n] = c.y;}
      if (iz) {iz[n] = c.z;}
      if (t) {t[n] = g->grid.val(s);}
    }
    ++n;
  }
  return n;
}
void oracle_load_voxels(void * h, const int32_t * ix, const int32_t * iy, const int32_t * iz, const double * t, uint64_t n)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  for (uint64_t i = 0; i < n; ++i) g->grid.set(Coord{ix[i], iy[i], iz[i]}, t[i]);
}
// GRID:312-317 GetFlattenedCostmap
uint64_t oracle_costmap_size(void * h) { return static_cast<OracleGrid *>(h)->cost_map.size(); }
uint64_t oracle_get_costmap(void * h, double * x, double * y, uint32_t * count, uint64_t cap)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  uint64_t n = 0;
  for (const auto & kv : g->cost_map) {
    if (n < cap) { x[n] = kv.first.x; y[n] = kv.first.y; count[n] = kv.second; }
    ++n;
  }
  return n;
}
uint64_t oracle_get_cleared(void * h, double * x, double * y, uint64_t cap)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  uint64_t n = 0;
  for (const auto & c : g->cleared_cells) {
    if (n < cap) { x[n] = c.x; y[n] = c.y; }
    ++n;
  }
  return n;
}
// GRID:344-376 GetOccupancyPointCloud payload
uint64_t oracle_get_points(void * h, float * xyz, uint64_t cap)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  const uint64_t n = g->grid_points.size();
  for (uint64_t i = 0; i < n && i < cap; ++i) {
    xyz[3 * i] = g->grid_points[i].x; xyz[3 * i + 1] = g->grid_points[i].y; xyz[3 * i + 2] = g->grid_points[i].z;
  }
  return n;
}
// GRID:379-394 ResetGrid
int oracle_reset(void * h)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  g->grid.clear();
  return g->grid.size() == 0;
}
void oracle_reset_area(void * h, double sx, double sy, double ex, double ey, int invert)
{
  static_cast<OracleGrid *>(h)->ResetGridArea(sx, sy, ex, ey, invert != 0);
}
// LAYER:699-725 UpdateROSCostmap.  bounds[4] = {min_x,min_y,max_x,max_y} in/out.
// Returns the number of columns written.
uint64_t oracle_update_ros_costmap(void * h, double origin_x, double origin_y, double resolution,
                                   uint32_t size_x, uint32_t size_y, int mark_threshold,
                                   uint8_t default_value, uint8_t lethal, uint8_t * costmap, double * bounds)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  std::memset(costmap, default_value, (size_t)size_x * size_y);   // Costmap2D::resetMaps, LAYER:705
  uint64_t n = 0;
  for (auto it = g->cost_map.begin(); it != g->cost_map.end(); ++it) {
    unsigned map_x, map_y;
    if (static_cast<int>(it->second) >= mark_threshold &&
      worldToMap(it->first.x, it->first.y, origin_x, origin_y, resolution, size_x, size_y, map_x, map_y))
    {
      costmap[(size_t)map_y * size_x + map_x] = lethal;   // getIndex
      touch(it->first.x, it->first.y, bounds);
      ++n;
    }
  }
  for (const auto & cell : g->cleared_cells) touch(cell.x, cell.y, bounds);   // LAYER:720-724
  return n;
}

// quantisation probe for KATs: world float -> index, exactly as GRID:293-303
void oracle_quantise(void * h, float fx, float fy, float fz, int32_t * out)
{
  OracleGrid * g = static_cast<OracleGrid *>(h);
  double x = fx < 0 ? fx - g->voxel_size : fx;
  double y = fy < 0 ? fy - g->voxel_size : fy;
  double z = fy < 0 ? fz - g->voxel_size : fz;
  const V3 m = g->WorldToIndex(v3(x, y, z));
  out[0] = (int32_t)m.x; out[1] = (int32_t)m.y; out[2] = (int32_t)m.z;
}
void oracle_index_to_world(void * h, int32_t ix, int32_t iy, int32_t iz, double * out)
{
  const V3 w = static_cast<OracleGrid *>(h)->IndexToWorld(Coord{ix, iy, iz});
  out[0] = w.x; out[1] = w.y; out[2] = w.z;
}
double oracle_decay_duration(void * h, double t) { return static_cast<OracleGrid *>(h)->GetTemporalClearingDuration(t); }
double oracle_frustum_acceleration(double t, double factor) { return OracleGrid::GetFrustumAcceleration(t, factor); }

// ---------------------------------------------------------------------------------
// BUFFER:153-175 cloud pre-filters (PCL, restated).  in: n points at point_step with
// float x,y,z offsets.  out_xyz: caller buffer of n*3 floats.  Returns kept count.
// ---------------------------------------------------------------------------------
// BUFFER:141-146  tf2::doTransform(cloud, *cld_global, tf_stamped).  The algorithm lives in tf2_sensor_msgs (ROS 2
// geometry2, not vendored in the reference; package.xml depends on tf2_sensor_msgs without a pin): it builds
//   Eigen::Transform<float,3,Isometry> t = Translation3f(tx,ty,tz) * Quaternionf(w,x,y,z)
// and writes t * Vector3f(x,y,z) for every point.  Restated: Eigen's Quaternion::toRotationMatrix in float, then
// out_i = ((m_i0*x + m_i1*y) + m_i2*z) + t_i (Eigen's 3x3 coefficient product sums left to right; no FMA contraction —
// this file is built with -ffp-contract=off).  Parity unpinned like the rest of the oracle.
void oracle_transform_from_quat(const double * translation, const double * quat_xyzw, float * m)
{
  const float x = (float)quat_xyzw[0], y = (float)quat_xyzw[1], z = (float)quat_xyzw[2], w = (float)quat_xyzw[3];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  m[0] = 1.0f - (tyy + tzz); m[1] = txy - twz;          m[2] = txz + twy;           m[3] = (float)translation[0];
  m[4] = txy + twz;          m[5] = 1.0f - (txx + tzz); m[6] = tyz - twx;           m[7] = (float)translation[1];
  m[8] = txz - twy;          m[9] = tyz + twx;          m[10] = 1.0f - (txx + tyy); m[11] = (float)translation[2];
}
// in: n points (float x,y,z at the given offsets); out_xyz: n*3 floats, every point transformed (NaN stays NaN)
void oracle_transform_cloud(const void * data, uint64_t n, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                            uint32_t off_z, const float * m, float * out_xyz)
{
  const uint8_t * d = static_cast<const uint8_t *>(data);
  for (uint64_t i = 0; i < n; ++i) {
    float x, y, z;
    std::memcpy(&x, d + i * point_step + off_x, 4);
    std::memcpy(&y, d + i * point_step + off_y, 4);
    std::memcpy(&z, d + i * point_step + off_z, 4);
    for (int r = 0; r < 3; ++r) {
      const float a = m[4 * r] * x, b = m[4 * r + 1] * y, c = m[4 * r + 2] * z;
      const float s1 = a + b;
      const float s2 = s1 + c;
      out_xyz[3 * i + r] = s2 + m[4 * r + 3];
    }
  }
}
// pcl::PassThrough<PCLPointCloud2> on "z", limits [lo,hi], keep_organized=false:
// drops points with non-finite x/y/z, and points whose z is outside the limits.
uint64_t oracle_filter_passthrough(const void * data, uint64_t n, uint32_t point_step,
                                   uint32_t off_x, uint32_t off_y, uint32_t off_z,
                                   double lo, double hi, float * out_xyz)
{
  const uint8_t * d = static_cast<const uint8_t *>(data);
  uint64_t k = 0;
  for (uint64_t i = 0; i < n; ++i) {
    float x, y, z;
    std::memcpy(&x, d + i * point_step + off_x, 4);
    std::memcpy(&y, d + i * point_step + off_y, 4);
    std::memcpy(&z, d + i * point_step + off_z, 4);
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    if (z < lo || z > hi) continue;   // PCL: outside if (value > max || value < min)
    out_xyz[3 * k] = x; out_xyz[3 * k + 1] = y; out_xyz[3 * k + 2] = z;
    ++k;
  }
  return k;
}
// pcl::VoxelGrid<PCLPointCloud2>, leaf = (float)voxel_size, filter field "z" in [lo,hi],
// downsample_all_data=false, min_points_per_voxel.  Points are binned by
// floor(p * inverse_leaf) relative to the min bin of the accepted points; the centroid of
// every bin holding >= min_points points is emitted (float accumulation in ascending point
// order — PCL sorts (bin,index) pairs with std::sort, whose order among equal bins is
// unspecified: ascending index is OUR definition).  Output order: ascending bin.
uint64_t oracle_filter_voxel(const void * data, uint64_t n, uint32_t point_step,
                             uint32_t off_x, uint32_t off_y, uint32_t off_z,
                             float leaf, double lo, double hi, int min_points, float * out_xyz)
{
  const uint8_t * d = static_cast<const uint8_t *>(data);
  struct P { float x, y, z; };
  std::vector<P> pts; pts.reserve(n);
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint64_t i = 0; i < n; ++i) {
    P p;
    std::memcpy(&p.x, d + i * point_step + off_x, 4);
    std::memcpy(&p.y, d + i * point_step + off_y, 4);
    std::memcpy(&p.z, d + i * point_step + off_z, 4);
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    if (p.z < lo || p.z > hi) continue;   // getMinMax3D with the filter field limits
    pts.push_back(p);
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  if (pts.empty()) return 0;
  const float inv = 1.0f / leaf;
  int mnb[3], mxb[3];
  for (int a = 0; a < 3; ++a) {
    mnb[a] = static_cast<int>(std::floor(mn[a] * inv));
    mxb[a] = static_cast<int>(std::floor(mx[a] * inv));
  }
  const int64_t dx = (int64_t)mxb[0] - mnb[0] + 1, dy = (int64_t)mxb[1] - mnb[1] + 1;
  struct IP { int64_t bin; uint32_t idx; };
  std::vector<IP> ip(pts.size());
  for (size_t i = 0; i < pts.size(); ++i) {
    const int ijk0 = static_cast<int>(std::floor(pts[i].x * inv) - static_cast<float>(mnb[0]));
    const int ijk1 = static_cast<int>(std::floor(pts[i].y * inv) - static_cast<float>(mnb[1]));
    const int ijk2 = static_cast<int>(std::floor(pts[i].z * inv) - static_cast<float>(mnb[2]));
    ip[i].bin = (int64_t)ijk0 + (int64_t)ijk1 * dx + (int64_t)ijk2 * dx * dy;
    ip[i].idx = (uint32_t)i;
  }
  std::sort(ip.begin(), ip.end(), [](const IP & a, const IP & b) {
      return a.bin != b.bin ? a.bin < b.bin : a.idx < b.idx; });
  uint64_t k = 0;
  size_t i = 0;
  while (i < ip.size()) {
    size_t j = i;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    while (j < ip.size() && ip[j].bin == ip[i].bin) {
      sx += pts[ip[j].idx].x; sy += pts[ip[j].idx].y; sz += pts[ip[j].idx].z;
      ++j;
    }
    const size_t cnt = j - i;
    if (cnt >= (size_t)std::max(min_points, 0)) {
      const float c = static_cast<float>(cnt);
      out_xyz[3 * k] = sx / c; out_xyz[3 * k + 1] = sy / c; out_xyz[3 * k + 2] = sz / c;
      ++k;
    }
    i = j;
  }
  return k;
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// LAYER:666-696 updateCosts: footprint clearing + combination into the master grid.  The byte-grid routines live in
// nav2_costmap_2d (external, not vendored, version unpinned: package.xml depends on nav2_costmap_2d); restated from the
// published sources of Costmap2D::setConvexPolygonCost / polygonOutlineCells / convexFillCells / raytraceLine /
// bresenham2D and CostmapLayer::updateWithMax / updateWithOverwrite.  Parity unpinned like the rest of this file.
// ---------------------------------------------------------------------------------
namespace {
struct MapLocation { unsigned x, y; };
inline int sign_of(int v) { return v > 0 ? 1 : -1; }
// Costmap2D::bresenham2D with the PolygonOutlineCells action (collect every visited cell)
void bresenham2d(std::vector<MapLocation> & cells, unsigned size_x, unsigned abs_da, unsigned abs_db, int error_b, int offset_a,
                 int offset_b, unsigned offset, unsigned max_length)
{
  const unsigned end = std::min(max_length, abs_da);
  for (unsigned i = 0; i < end; ++i) {
    cells.push_back(MapLocation{offset % size_x, offset / size_x});
    offset += offset_a;
    error_b += abs_db;
    if ((unsigned)error_b >= abs_da) { offset += offset_b; error_b -= abs_da; }
  }
  cells.push_back(MapLocation{offset % size_x, offset / size_x});
}
// Costmap2D::raytraceLine(at, x0, y0, x1, y1) with max_length = UINT_MAX, min_length = 0
void raytrace_line(std::vector<MapLocation> & cells, unsigned size_x, unsigned x0, unsigned y0, unsigned x1, unsigned y1)
{
  const int dx = (int)x1 - (int)x0, dy = (int)y1 - (int)y0;
  const unsigned abs_dx = (unsigned)std::abs(dx), abs_dy = (unsigned)std::abs(dy);
  const int offset_dx = sign_of(dx), offset_dy = sign_of(dy) * (int)size_x;
  const unsigned offset = y0 * size_x + x0;
  // dist / scale only matter for a finite max_length: scale = 1 here
  if (abs_dx >= abs_dy) {
    bresenham2d(cells, size_x, abs_dx, abs_dy, (int)(abs_dx / 2), offset_dx, offset_dy, offset, abs_dx);
    return;
  }
  bresenham2d(cells, size_x, abs_dy, abs_dx, (int)(abs_dy / 2), offset_dy, offset_dx, offset, abs_dy);
}
}  // namespace

extern "C" {
// Costmap2D::setConvexPolygonCost: returns 0 (and changes nothing) if a vertex is off the map.  polygon_xy = n x (x, y).
int oracle_set_convex_polygon_cost(uint8_t * costmap, double origin_x, double origin_y, double resolution, uint32_t size_x,
                                   uint32_t size_y, const double * polygon_xy, uint32_t n, uint8_t cost)
{
  std::vector<MapLocation> map_polygon;
  for (uint32_t i = 0; i < n; ++i) {
    MapLocation loc;
    if (!worldToMap(polygon_xy[2 * i], polygon_xy[2 * i + 1], origin_x, origin_y, resolution, size_x, size_y, loc.x, loc.y)) return 0;
    map_polygon.push_back(loc);
  }
  std::vector<MapLocation> cells;
  // convexFillCells
  if (map_polygon.size() >= 3) {
    // polygonOutlineCells: every edge, then the closing edge
    for (size_t i = 0; i + 1 < map_polygon.size(); ++i)
      raytrace_line(cells, size_x, map_polygon[i].x, map_polygon[i].y, map_polygon[i + 1].x, map_polygon[i + 1].y);
    raytrace_line(cells, size_x, map_polygon.back().x, map_polygon.back().y, map_polygon[0].x, map_polygon[0].y);
    // "quick bubble sort to sort points by x"
    MapLocation swap;
    size_t i = 0;
    while (i + 1 < cells.size()) {
      if (cells[i].x > cells[i + 1].x) {
        swap = cells[i]; cells[i] = cells[i + 1]; cells[i + 1] = swap;
        if (i > 0) --i;
      } else {
        ++i;
      }
    }
    i = 0;
    MapLocation min_pt, max_pt;
    const unsigned min_x = cells[0].x, max_x = cells[cells.size() - 1].x;
    // walk through each column and mark cells inside the polygon
    for (unsigned x = min_x; x <= max_x; ++x) {
      if (i + 1 >= cells.size()) break;
      if (cells[i].y < cells[i + 1].y) { min_pt = cells[i]; max_pt = cells[i + 1]; }
      else { min_pt = cells[i + 1]; max_pt = cells[i]; }
      i += 2;
      while (i < cells.size() && cells[i].x == x) {
        if (cells[i].y < min_pt.y) min_pt = cells[i];
        else if (cells[i].y > max_pt.y) max_pt = cells[i];
        ++i;
      }
      for (unsigned y = min_pt.y; y <= max_pt.y; ++y) cells.push_back(MapLocation{x, y});
    }
  }
  for (const MapLocation & c : cells) costmap[(size_t)c.y * size_x + c.x] = cost;
  return 1;
}

// CostmapLayer::updateWithOverwrite (method 0) / updateWithMax (method 1) over [min_i,max_i) x [min_j,max_j); 255 = NO_INFORMATION
void oracle_update_costs(const uint8_t * layer, uint8_t * master, uint32_t size_x, int min_i, int min_j, int max_i, int max_j, int method)
{
  for (int j = min_j; j < max_j; ++j) {
    size_t it = (size_t)size_x * j + min_i;
    for (int i = min_i; i < max_i; ++i, ++it) {
      if (method == 0) {
        if (layer[it] != 255) master[it] = layer[it];
      } else if (method == 1) {
        if (layer[it] == 255) continue;
        const uint8_t old_cost = master[it];
        if (old_cost == 255 || old_cost < layer[it]) master[it] = layer[it];
      }
    }
  }
}
}  // extern "C"
