// stvl2pc — converts a grid saved by SpatioTemporalVoxelGrid::SaveGrid ("<name>.stvl") into a point cloud file, the
// counterpart of the reference's vdb2pc (reference spatio_temporal_voxel_layer/src/vdb2pc.cpp:60-93 reads the .vdb back and
// emits grid->indexToWorld(coord) — the voxel's lower corner, coord * voxel_size, without the half-voxel offset — per active
// voxel).  Pure host code: no GPU, no library.
//
//   stvl2pc <file.stvl> <out.pcd> [--centres] [--times]
//     --centres   voxel centres (coord * vs + vs / 2, IndexToWorld of the grid) instead of vdb2pc's lower corners
//     --times     add the mark time as a fourth field (FLOAT64 "t")
// .stvl layout: "STVLGRD1", voxel_size f64, n u64, then n x (ix, iy, iz i32, t f64), little endian.
// Output: ASCII PCD v0.7, fields x y z (FLOAT32) [t].
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

int main(int argc, char ** argv)
{
  if (argc < 3) {
    std::fprintf(stderr, "usage: stvl2pc <file.stvl> <out.pcd> [--centres] [--times]\n");
    return 2;
  }
  bool centres = false, times = false;
  for (int i = 3; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--centres")) centres = true;
    else if (!std::strcmp(argv[i], "--times")) times = true;
    else { std::fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
  }
  std::ifstream f(argv[1], std::ios::binary);
  char magic[8];
  double vs = 0.;
  uint64_t n = 0;
  if (!f || !f.read(magic, 8) || std::memcmp(magic, "STVLGRD1", 8) != 0) {
    std::fprintf(stderr, "No valid grid inside of provided file.\n");
    return 1;
  }
  f.read(reinterpret_cast<char *>(&vs), 8);
  f.read(reinterpret_cast<char *>(&n), 8);
  struct Rec { int32_t x, y, z; double t; };
  std::vector<Rec> recs(n);
  for (uint64_t i = 0; i < n; ++i) {
    f.read(reinterpret_cast<char *>(&recs[i].x), 4);
    f.read(reinterpret_cast<char *>(&recs[i].y), 4);
    f.read(reinterpret_cast<char *>(&recs[i].z), 4);
    f.read(reinterpret_cast<char *>(&recs[i].t), 8);
  }
  if (!f) {
    std::fprintf(stderr, "truncated file: expected %llu voxels\n", (unsigned long long)n);
    return 1;
  }
  std::FILE * o = std::fopen(argv[2], "w");
  if (!o) { std::fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
  std::fprintf(o, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n");
  std::fprintf(o, times ? "FIELDS x y z t\nSIZE 4 4 4 8\nTYPE F F F F\nCOUNT 1 1 1 1\n" : "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n");
  std::fprintf(o, "WIDTH %llu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %llu\nDATA ascii\n", (unsigned long long)n, (unsigned long long)n);
  const double off = centres ? vs / 2.0 : 0.0;
  for (const Rec & r : recs) {
    // pcl::PointXYZ holds floats: the doubles are narrowed like vdb2pc's pcl::PointXYZ(pt[0], pt[1], pt[2])
    const float x = (float)((double)r.x * vs + off), y = (float)((double)r.y * vs + off), z = (float)((double)r.z * vs + off);
    if (times) std::fprintf(o, "%.9g %.9g %.9g %.17g\n", x, y, z, r.t);
    else std::fprintf(o, "%.9g %.9g %.9g\n", x, y, z);
  }
  std::fclose(o);
  std::printf("stvl2pc: %llu voxels, voxel size %.17g\n", (unsigned long long)n, vs);
  return 0;
}
