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
  {   // planner-side additions of the facade: CUDA-graph query plan and pinned host mirror return the same bits as the per-call queries
    fiesta_query_plan *plan = map->MakeQueryPlan(4);
    double *qp = fiesta_query_plan_positions(plan);
    const double pts[4][3] = {{0.33, -1.27, 2.51}, {-3.05, 2.2, 0.71}, {5.9, 5.9, 11.3}, {99, 0, 0}};
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) qp[3 * i + k] = pts[i][k];
    fiesta_query_plan_run(plan);
    fiesta_host_mirror *mir = map->MakeHostMirror();
    fiesta_host_mirror_refresh(mir, nullptr);
    int same = 1;
    for (int i = 0; i < 4; ++i) {
      Eigen::Vector3d gi(0, 0, 0);
      const double di = map->GetDistWithGradTrilinear(Eigen::Vector3d(pts[i][0], pts[i][1], pts[i][2]), gi);
      double gm[3];
      const double dm = fiesta_host_mirror_get_dist_grad_trilinear(mir, pts[i], gm);
      const double *gp = fiesta_query_plan_gradients(plan) + 3 * i;
      same = same && di == fiesta_query_plan_distances(plan)[i] && di == dm;
      for (int k = 0; k < 3; ++k) same = same && gi(k) == gp[k] && gi(k) == gm[k];
      same = same && map->GetDistance(Eigen::Vector3d(pts[i][0], pts[i][1], pts[i][2])) == fiesta_host_mirror_get_distance_pos(mir, pts[i]);
    }
    std::printf("plan_and_mirror_equal %d\n", same);
    fiesta_host_mirror_destroy(mir);
    fiesta_query_plan_destroy(plan);
  }
  sensor_msgs::PointCloud pc;
  map->GetPointCloud(pc, 0, 63);
  std::printf("occupied_points %zu\n", pc.points.size());
  visualization_msgs::Marker mk;
  map->GetSliceMarker(mk, 8, 100, Eigen::Vector4d(0, 1, 0, 1), 2.0);
  std::printf("slice_points %zu\n", mk.points.size());
  delete map;
  return 0;
}
