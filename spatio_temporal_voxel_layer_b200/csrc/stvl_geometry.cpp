// stvl_geometry.cpp — host-side frustum geometry of the product library.
//
// The reference builds its frustum objects on the host once per clearing reading per cycle
// (spatio_temporal_voxel_grid.cpp:120-144); that stays host code here.  These functions produce the
// stvl_frustum_t parameter blocks the sweep kernel stages in shared memory, and the conservative culling
// data (voxel-index AABB) that lets the kernel skip whole 8^3 leaves.
//
// Compiled with -ffp-contract=off: the plane blocks must not depend on FMA availability.
#include <cmath>
#include <cstring>
#include <limits>

#include "stvl_b200.h"
#include "stvl_device.cuh"
#include "stvl_geometry.h"

namespace stvl {
namespace {

struct Vec { double v[3]; };
struct Mat { double a[9]; };   // row-major

Vec vsub(const Vec & p, const Vec & q) { return Vec{{p.v[0] - q.v[0], p.v[1] - q.v[1], p.v[2] - q.v[2]}}; }
Vec vadd(const Vec & p, const Vec & q) { return Vec{{p.v[0] + q.v[0], p.v[1] + q.v[1], p.v[2] + q.v[2]}}; }
Vec vscale(const Vec & p, double s) { return Vec{{p.v[0] * s, p.v[1] * s, p.v[2] * s}}; }
Vec vcross(const Vec & p, const Vec & q)
{
  return Vec{{p.v[1] * q.v[2] - p.v[2] * q.v[1], p.v[2] * q.v[0] - p.v[0] * q.v[2], p.v[0] * q.v[1] - p.v[1] * q.v[0]}};
}
// Eigen normalize(): divide by sqrt(squaredNorm) when it is > 0; squaredNorm = (x^2 + y^2) + z^2
void vnormalize(Vec & p)
{
  const double n2 = (p.v[0] * p.v[0] + p.v[1] * p.v[1]) + p.v[2] * p.v[2];
  if (n2 > 0.0) {
    const double n = std::sqrt(n2);
    p.v[0] /= n; p.v[1] /= n; p.v[2] /= n;
  }
}
Vec mat_vec(const Mat & m, const Vec & p)
{
  Vec r;
  for (int i = 0; i < 3; ++i) r.v[i] = (m.a[3 * i] * p.v[0] + m.a[3 * i + 1] * p.v[1]) + m.a[3 * i + 2] * p.v[2];
  return r;
}
Mat mat_mat(const Mat & l, const Mat & r)
{
  Mat o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      o.a[3 * i + j] = (l.a[3 * i] * r.a[j] + l.a[3 * i + 1] * r.a[3 + j]) + l.a[3 * i + 2] * r.a[6 + j];
  return o;
}
// rotation by `angle` about a unit axis (Eigen AngleAxisd::toRotationMatrix evaluation order)
Mat rotation_about(double angle, const Vec & axis)
{
  const double s = std::sin(angle), c = std::cos(angle);
  const Vec sa = vscale(axis, s), ca = vscale(axis, 1.0 - c);
  Mat m;
  double t = ca.v[0] * axis.v[1];
  m.a[1] = t - sa.v[2]; m.a[3] = t + sa.v[2];
  t = ca.v[0] * axis.v[2];
  m.a[2] = t + sa.v[1]; m.a[6] = t - sa.v[1];
  t = ca.v[1] * axis.v[2];
  m.a[5] = t - sa.v[0]; m.a[7] = t + sa.v[0];
  m.a[0] = ca.v[0] * axis.v[0] + c; m.a[4] = ca.v[1] * axis.v[1] + c; m.a[8] = ca.v[2] * axis.v[2] + c;
  return m;
}
// unit quaternion (x,y,z,w) -> rotation matrix (Eigen toRotationMatrix evaluation order)
Mat rotation_of(const double q[4])
{
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  Mat m;
  m.a[0] = 1.0 - (tyy + tzz); m.a[1] = txy - twz; m.a[2] = txz + twy;
  m.a[3] = txy + twz; m.a[4] = 1.0 - (txx + tzz); m.a[5] = tyz - twx;
  m.a[6] = txz - twy; m.a[7] = tyz + twx; m.a[8] = 1.0 - (txx + tyy);
  return m;
}
// q * v (Eigen _transformVector)
Vec rotate_by(const double q[4], const Vec & p)
{
  const Vec qv{{q[0], q[1], q[2]}};
  Vec uv = vcross(qv, p);
  uv = vadd(uv, uv);
  const Vec c2 = vcross(qv, uv);
  return Vec{{(p.v[0] + q[3] * uv.v[0]) + c2.v[0], (p.v[1] + q[3] * uv.v[1]) + c2.v[1], (p.v[2] + q[3] * uv.v[2]) + c2.v[2]}};
}

void put_plane(double * out, int k, const Vec & n, const Vec & p)
{
  for (int i = 0; i < 3; ++i) { out[6 * k + i] = n.v[i]; out[6 * k + 3 + i] = p.v[i]; }
}

// intersection of three planes n_i . (x - p_i) = 0 by Cramer's rule; false if (nearly) singular
bool three_planes(const double * a, const double * b, const double * c, double out[3])
{
  const double d[3] = {a[0] * a[3] + a[1] * a[4] + a[2] * a[5], b[0] * b[3] + b[1] * b[4] + b[2] * b[5],
                       c[0] * c[3] + c[1] * c[4] + c[2] * c[5]};
  const double det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
  if (!(std::fabs(det) > 1e-9)) return false;
  out[0] = (d[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (d[1] * c[2] - b[2] * d[2]) + a[2] * (d[1] * c[1] - b[1] * d[2])) / det;
  out[1] = (a[0] * (d[1] * c[2] - b[2] * d[2]) - d[0] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * d[2] - d[1] * c[0])) / det;
  out[2] = (a[0] * (b[1] * d[2] - d[1] * c[1]) - a[1] * (b[0] * d[2] - d[1] * c[0]) + d[0] * (b[0] * c[1] - b[1] * c[0])) / det;
  return std::isfinite(out[0]) && std::isfinite(out[1]) && std::isfinite(out[2]);
}

}  // namespace

// DepthCameraFrustum: ctor + ComputePlaneNormals (depth_camera_frustum.cpp:45-157), then SetPosition /
// SetOrientation / TransformModel (:160-174, :306-319) with VectorWithPt3D::TransformFrames (frustum.hpp:83-89).
int build_depth_frustum(double vfov, double hfov, double min_d, double max_d, const double origin[3],
                        const double q[4], double accel, stvl_frustum_t * out)
{
  std::memset(out, 0, sizeof(*out));
  out->accel_factor = accel;
  if (vfov == 0 && hfov == 0) {   // "ability to construct with bogus values": never contains anything
    out->type = STVL_MODEL_INVALID;
    return STVL_OK;
  }
  const Vec ex{{1, 0, 0}}, ey{{0, 1, 0}};
  // four corner rays, CCW: R_x(+-vFOV/2) * R_y(+-hFOV/2) * Z  = third column of the product
  const double sv[4] = {+1, -1, -1, +1}, sh[4] = {+1, +1, -1, -1};
  Vec pts[8];
  for (int i = 0; i < 4; ++i) {
    const Mat m = mat_mat(rotation_about(sv[i] * vfov / 2., ex), rotation_about(sh[i] * hfov / 2., ey));
    const Vec ray{{m.a[2], m.a[5], m.a[8]}};
    pts[2 * i] = vscale(ray, min_d);
    pts[2 * i + 1] = vscale(ray, max_d);
  }
  double * pl = out->p;
  const Vec v01 = vsub(pts[1], pts[0]), v13 = vsub(pts[3], pts[1]);
  Vec n = vcross(v13, v01); vnormalize(n); put_plane(pl, 0, n, pts[0]);          // top
  const Vec v23 = vsub(pts[3], pts[2]), v35 = vsub(pts[5], pts[3]);
  n = vcross(v35, v23); vnormalize(n); put_plane(pl, 1, n, pts[2]);              // left
  const Vec v45 = vsub(pts[5], pts[4]), v57 = vsub(pts[7], pts[5]);
  n = vcross(v57, v45); vnormalize(n); put_plane(pl, 2, n, pts[4]);              // bottom
  const Vec v67 = vsub(pts[7], pts[6]), v71 = vsub(pts[1], pts[7]);
  n = vcross(v71, v67); vnormalize(n); put_plane(pl, 3, n, pts[6]);              // right
  Vec f = vcross(v57, v71); vnormalize(f); put_plane(pl, 4, f, pts[7]);          // far
  put_plane(pl, 5, Vec{{-1 * f.v[0], -1 * f.v[1], -1 * f.v[2]}}, pts[2]);        // near

  // T = Identity; pretranslate(q^-1 * origin); prerotate(q)  =>  linear R(q), translation R(q) * (q^-1 * origin)
  const double n2 = (q[0] * q[0] + q[2] * q[2]) + (q[1] * q[1] + q[3] * q[3]);
  double qi[4] = {0, 0, 0, 0};
  if (n2 > 0.0) { qi[0] = -q[0] / n2; qi[1] = -q[1] / n2; qi[2] = -q[2] / n2; qi[3] = q[3] / n2; }
  const Mat R = rotation_of(q);
  const Vec tr = mat_vec(R, rotate_by(qi, Vec{{origin[0], origin[1], origin[2]}}));
  for (int k = 0; k < 6; ++k) {
    Vec nn = mat_vec(R, Vec{{pl[6 * k], pl[6 * k + 1], pl[6 * k + 2]}});
    vnormalize(nn);
    const Vec pp = vadd(mat_vec(R, Vec{{pl[6 * k + 3], pl[6 * k + 4], pl[6 * k + 5]}}), tr);
    put_plane(pl, k, nn, pp);
  }
  out->type = STVL_MODEL_DEPTH_CAMERA;
  // a proper, bounded hexahedron: the sweep may cull leaves against the vertices' bounding box
  out->reserved = (vfov > 1e-6 && vfov < 3.0 && hfov > 1e-6 && hfov < 3.0 && min_d >= 0 && max_d > min_d && std::isfinite(max_d)) ? 1 : 0;
  return STVL_OK;
}

// ThreeDimensionalLidarFrustum ctor precompute + TransformModel (three_dimensional_lidar_frustum.cpp:44-74)
int build_lidar_frustum(double vfov, double vpad, double hfov, double min_d, double max_d, const double origin[3],
                        const double q[4], double accel, stvl_frustum_t * out)
{
  std::memset(out, 0, sizeof(*out));
  out->accel_factor = accel;
  double * p = out->p;
  const double hfov_half = hfov / 2.0;
  const double tan_half = std::tan(vfov / 2.0);
  p[0] = origin[0]; p[1] = origin[1]; p[2] = origin[2];
  p[3] = -q[0]; p[4] = -q[1]; p[5] = -q[2]; p[6] = q[3];
  p[7] = vpad;
  p[8] = tan_half * tan_half;
  p[9] = min_d * min_d;
  p[10] = max_d * max_d;
  p[11] = hfov_half;
  p[12] = hfov > 6.27 ? 1.0 : 0.0;
  out->type = STVL_MODEL_3D_LIDAR;
  out->reserved = (vfov > 0 && vfov < 3.0 && min_d >= 0 && max_d > min_d && std::isfinite(max_d) && std::isfinite(vpad)) ? 1 : 0;
  return STVL_OK;
}

// block -> what the sweep kernel consumes: (1./6.)*factor, conservative voxel-index AABB, lidar scalars
void prepare_frustum(const stvl_frustum_t & in, double vs, DevFrustum * out)
{
  std::memset(out, 0, sizeof(*out));
  out->type = in.type;
  out->c16f = 1. / 6. * in.accel_factor;   // reference .cpp:338: 1. / 6. * acceleration_factor * (...)
  std::memcpy(out->p, in.p, sizeof(in.p));
  out->cull = 0;
  if (in.type != STVL_MODEL_DEPTH_CAMERA && in.type != STVL_MODEL_3D_LIDAR) { out->type = -1; return; }
  double lo[3], hi[3];
  bool ok = in.reserved == 1;
  if (ok && in.type == STVL_MODEL_DEPTH_CAMERA) {
    // vertices = side(k) ^ side(k+1) ^ {near, far}
    for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<double>::infinity(); hi[a] = -lo[a]; }
    for (int k = 0; k < 4 && ok; ++k) {
      for (int cap = 4; cap <= 5 && ok; ++cap) {
        double x[3];
        ok = three_planes(in.p + 6 * k, in.p + 6 * ((k + 1) & 3), in.p + 6 * cap, x);
        for (int a = 0; a < 3 && ok; ++a) { lo[a] = std::fmin(lo[a], x[a]); hi[a] = std::fmax(hi[a], x[a]); }
      }
    }
  } else if (ok) {
    const double tan_half = std::sqrt(in.p[8]), min_d = std::sqrt(in.p[9]), max_d = std::sqrt(in.p[10]);
    const double zmax = tan_half * max_d + std::fabs(in.p[7]);
    const double rad = std::sqrt(max_d * max_d + zmax * zmax) * (1.0 + 1e-9) + 1e-6;
    ok = std::isfinite(rad);
    for (int a = 0; a < 3; ++a) { lo[a] = in.p[a] - rad; hi[a] = in.p[a] + rad; }
    out->aux[0] = min_d; out->aux[1] = max_d; out->aux[2] = tan_half; out->aux[3] = std::fabs(in.p[7]);
  }
  if (!ok) return;
  for (int a = 0; a < 3; ++a) {
    // voxel centre = i*vs + vs/2  =>  i in [(lo - vs/2)/vs, (hi - vs/2)/vs], widened by two voxels
    const double l = std::floor((lo[a] - 1e-6) / vs - 0.5) - 2.0, h = std::ceil((hi[a] + 1e-6) / vs - 0.5) + 2.0;
    if (!(std::fabs(l) < 2e9 && std::fabs(h) < 2e9)) return;
    out->lo[a] = (int32_t)l; out->hi[a] = (int32_t)h;
  }
  out->cull = 1;
}

}  // namespace stvl
