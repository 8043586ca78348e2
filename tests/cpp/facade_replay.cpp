// Pillar replay (the intent of the reference's test/test_ESDF_Map.cpp:42-103) written against
// the fiesta::ESDFMap facade exactly as one would write it against the reference class.  Prints checkpoint values that
// tests/test_gpu_facade.py compares with tests/golden/reference_golden.json.
#include <cstdio>
#include "ESDFMap.h"

static void checkpoint(fiesta::ESDFMap &m, const char *tag) {
  long finite = 0; double sum = 0;
  for (int x = 0; x < 64; ++x) for (int y = 0; y < 64; ++y) for (int z = 0; z < 64; ++z) {
    // GetDistance(Vector3i) reads unknown/unreached as +10000 (ESDFMap.cpp:477-479)
    double d = m.GetDistance(Eigen::Vector3d(-6.4 + (x + 0.5) * 0.2, -6.4 + (y + 0.5) * 0.2, (z + 0.5) * 0.2));
    if (d >= 0 && d < 10000) { ++finite; sum += d; }
  }
  std::printf("%s finite %ld sum %.6f\n", tag, finite, sum);
}

int main() {
  fiesta::ESDFMap *map = new fiesta::ESDFMap(Eigen::Vector3d(-6.4, -6.4, 0.0), 0.2, Eigen::Vector3d(12.8, 12.8, 12.8));
  map->SetParameters(0.97, 0.03, 0.30, 0.90, 0.80);
  std::printf("grid_total_size_ %d\n", map->grid_total_size_);
  for (int x = 0; x < 64; ++x) for (int y = 0; y < 64; ++y) for (int z = 0; z < 64; ++z) map->SetOccupancy(Eigen::Vector3i(x, y, z), 0);
  if (map->CheckUpdate()) { map->SetOriginalRange(); map->UpdateOccupancy(true); map->UpdateESDF(); }
  const int s[5] = {12, 22, 32, 42, 52};
  for (int a = 0; a < 5; ++a) for (int b = 0; b < 5; ++b) {
    for (int z = 0; z < 25; ++z) {
      int idx = map->SetOccupancy(Eigen::Vector3d(-6.4 + (s[a] + 0.5) * 0.2, -6.4 + (s[b] + 0.5) * 0.2, (z + 0.5) * 0.2), 1);
      if (idx != s[a] * 64 * 64 + s[b] * 64 + z) { std::printf("bad index %d\n", idx); return 1; }
    }
    if (map->CheckUpdate()) { map->UpdateOccupancy(true); map->UpdateESDF(); }
  }
  std::printf("GetDistance(30,30,10) %.9f\n", map->GetDistance(Eigen::Vector3i(30, 30, 10)));
  Eigen::Vector3d g(0, 0, 0);
  double d = map->GetDistWithGradTrilinear(Eigen::Vector3d(0.33, -1.27, 2.51), g);
  std::printf("trilinear %.9f %.9f %.9f %.9f\n", d, g(0), g(1), g(2));
  std::printf("out_of_map %d %.1f %.1f\n", map->SetOccupancy(Eigen::Vector3d(99, 0, 0), 1), map->GetDistance(Eigen::Vector3d(99, 0, 0)),
              map->GetDistWithGradTrilinear(Eigen::Vector3d(99, 0, 0), g));
  sensor_msgs::PointCloud pc;
  map->GetPointCloud(pc, 0, 63);
  std::printf("occupied_points %zu\n", pc.points.size());
  visualization_msgs::Marker mk;
  map->GetSliceMarker(mk, 8, 100, Eigen::Vector4d(0, 1, 0, 1), 2.0);
  std::printf("slice_points %zu\n", mk.points.size());
  delete map;
  return 0;
}
