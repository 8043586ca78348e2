// fiesta_b200 -- header-compatible C++ facade for the reference class `fiesta::ESDFMap`.
//
// Put this directory BEFORE the reference's include/ on the include path and `#include "ESDFMap.h"` in Fiesta.h
// resolves here: same namespace, class name, constructor and public methods as /root/reference/include/ESDFMap.h:111-164
// (dense-array + PROBABILISTIC build, the shipped configuration: parameters.h:9-14), same public field
// `grid_total_size_`, same sentinel returns.  Every method forwards 1:1 to the C ABI in include/fiesta_b200.h, which runs
// on the B200.  Ownership matches the reference (Fiesta.h:96,137): `new ESDFMap(...)` / `delete`.
//
// Differences a maintainer should know:
//  * RaycastFrame() is an ADDITION: one call replaces the whole Fiesta::RaycastMultithread loop (Fiesta.h:281-303) and is
//    the fast path.  The per-call SetOccupancy() path still works unchanged (events are staged and applied on the device
//    at the next UpdateOccupancy), so Fiesta.h compiles and runs without edits.
//  * SetOccupancy() is not thread-safe; the reference's threaded ray casting (ray_cast_num_thread > 0) is itself racy
//    (Fiesta.h:294-300, ESDFMap.cpp:430-433).  Use RaycastFrame() instead of host threads.
//  * A failed construction (no sm_100 GPU, out of memory) throws std::runtime_error -- there is no CPU fallback.
#ifndef ESDF_MAP_H
#define ESDF_MAP_H

#include <Eigen/Eigen>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>
#include <visualization_msgs/Marker.h>
#include <sensor_msgs/PointCloud.h>
#if defined(__has_include)
#if __has_include("parameters.h")
#include "parameters.h"   // the reference's own header (macros, dirs_, Parameters) when built inside the FIESTA tree
#endif
#endif
#include "../fiesta_b200.h"

namespace fiesta {

class ESDFMap {
  fiesta_map *h_ = nullptr;
  double resolution_;
  Eigen::Vector3d origin_;
  Eigen::Vector3i grid_size_, min_vec_, max_vec_;
  double min_occupancy_log_ = 0;

  static void check(int rc, const char *what) {
    if (rc != FIESTA_OK) throw std::runtime_error(std::string(what) + ": " + fiesta_last_error());
  }
  void Pos2VoxHost(const Eigen::Vector3d &pos, Eigen::Vector3i &vox) const {   // ESDFMap.cpp:74-77
    for (int i = 0; i < 3; ++i) vox(i) = (int)std::floor((pos(i) - origin_(i)) / resolution_);
  }

 public:
  int grid_total_size_;   // ESDFMap.h:115, read by Fiesta.h:107-108

  // ESDFMap(origin, resolution, map_size) -- ESDFMap.h:116, ESDFMap.cpp:171-213
  ESDFMap(Eigen::Vector3d origin, double resolution, Eigen::Vector3d map_size, int device = 0)
      : resolution_(resolution), origin_(origin) {
    fiesta_config cfg = {};
    for (int i = 0; i < 3; ++i) { cfg.origin[i] = origin(i); cfg.map_size[i] = map_size(i); }
    cfg.resolution = resolution;
    cfg.device = device;
    check(fiesta_create(&cfg, &h_), "fiesta_create");
    grid_total_size_ = fiesta_grid_total_size(h_);
    int gs[3];
    fiesta_grid_size(h_, gs);
    grid_size_ = Eigen::Vector3i(gs[0], gs[1], gs[2]);
    min_vec_ = Eigen::Vector3i(0, 0, 0);
    max_vec_ = Eigen::Vector3i(gs[0] - 1, gs[1] - 1, gs[2] - 1);
  }
  ~ESDFMap() { fiesta_destroy(h_); }
  ESDFMap(const ESDFMap &) = delete;
  ESDFMap &operator=(const ESDFMap &) = delete;
  fiesta_map *handle() { return h_; }

  // ESDFMap.h:124
  void SetParameters(double p_hit, double p_miss, double p_min, double p_max, double p_occ) {
    check(fiesta_set_parameters(h_, p_hit, p_miss, p_min, p_max, p_occ), "SetParameters");
    min_occupancy_log_ = std::log(p_occ / (1 - p_occ));
  }

  // ESDFMap.h:128-130 -- the per-frame driver calls (Fiesta.h:507-514)
  bool CheckUpdate() { return fiesta_check_update(h_) != 0; }
  bool UpdateOccupancy(bool global_map) {
    int r = fiesta_update_occupancy(h_, global_map ? 1 : 0);
    if (r < 0) check(-r, "UpdateOccupancy");
    return r > 0;
  }
  void UpdateESDF() { check(fiesta_update_esdf(h_), "UpdateESDF"); }

  // ESDFMap.h:133-136
  int SetOccupancy(Eigen::Vector3d pos, int occ) { double p[3] = {pos(0), pos(1), pos(2)}; return fiesta_set_occupancy_pos(h_, p, occ); }
  int SetOccupancy(Eigen::Vector3i vox, int occ) { int v[3] = {vox(0), vox(1), vox(2)}; return fiesta_set_occupancy_vox(h_, v, occ); }
  int GetOccupancy(Eigen::Vector3d pos) { double p[3] = {pos(0), pos(1), pos(2)}; return fiesta_get_occupancy_pos(h_, p); }
  int GetOccupancy(Eigen::Vector3i vox) { int v[3] = {vox(0), vox(1), vox(2)}; return fiesta_get_occupancy_vox(h_, v); }

  // ESDFMap.h:139-141
  double GetDistance(Eigen::Vector3d pos) { double p[3] = {pos(0), pos(1), pos(2)}; return fiesta_get_distance_pos(h_, p); }
  double GetDistance(Eigen::Vector3i vox) { int v[3] = {vox(0), vox(1), vox(2)}; return fiesta_get_distance_vox(h_, v); }
  double GetDistWithGradTrilinear(Eigen::Vector3d pos, Eigen::Vector3d &grad) {
    double p[3] = {pos(0), pos(1), pos(2)}, g[3] = {0, 0, 0};
    double d = fiesta_get_dist_grad_trilinear(h_, p, g);
    grad(0) = g[0]; grad(1) = g[1]; grad(2) = g[2];
    return d;
  }

  // ESDFMap.h:148-149
  void SetUpdateRange(Eigen::Vector3d min_pos, Eigen::Vector3d max_pos, bool new_vec = true) {
    double a[3] = {min_pos(0), min_pos(1), min_pos(2)}, b[3] = {max_pos(0), max_pos(1), max_pos(2)};
    check(fiesta_set_update_range(h_, a, b, new_vec ? 1 : 0), "SetUpdateRange");
    // host mirror of the box for the visualisation loops below (ESDFMap.cpp:794-809)
    for (int i = 0; i < 3; ++i) {
      double lo = std::max(min_pos(i), origin_(i)), hi = std::min(max_pos(i), origin_(i) + grid_size_(i) * resolution_);
      min_vec_(i) = (int)std::floor((lo - origin_(i)) / resolution_);
      max_vec_(i) = std::min((int)std::floor((hi - resolution_ / 2 - origin_(i)) / resolution_), grid_size_(i) - 1);
    }
  }
  void SetOriginalRange() {
    check(fiesta_set_original_range(h_), "SetOriginalRange");
    min_vec_ = Eigen::Vector3i(0, 0, 0);
    max_vec_ = Eigen::Vector3i(grid_size_(0) - 1, grid_size_(1) - 1, grid_size_(2) - 1);
  }

  // ---- additions (fast paths; not in the reference class) ----
  // One call for Fiesta::RaycastMultithread (Fiesta.h:281-303): `cloud` holds n points as packed float xyz in the sensor
  // frame, `transform` is Fiesta's transform_ (row-major 4x4).  Serial-mode semantics, computed on the GPU.
  void RaycastFrame(const float *cloud_xyz, long n, const double transform_row_major[16], double min_ray_length, double max_ray_length) {
    fiesta_raycast_params p = {min_ray_length, max_ray_length};
    check(fiesta_raycast_frame(h_, cloud_xyz, n, transform_row_major, &p), "RaycastFrame");
  }
  void GetDistWithGradTrilinearBatch(const double *pos_xyz, long n, double *dist, double *grad_xyz) {
    check(fiesta_get_dist_grad_trilinear_batch(h_, pos_xyz, n, dist, grad_xyz), "GetDistWithGradTrilinearBatch");
  }

  // ---- visualisation (ESDFMap.h:144-145; off the hot path: one device->host dump per call) ----
  void GetPointCloud(sensor_msgs::PointCloud &m, int vis_lower_bound, int vis_upper_bound) {
    m.header.frame_id = "world";
    m.points.clear();
    std::vector<double> occ((size_t)grid_total_size_);
    check(fiesta_export_occupancy(h_, occ.data()), "export_occupancy");
    const int gyz = grid_size_(1) * grid_size_(2);
    for (int x = min_vec_(0); x <= max_vec_(0); ++x)
      for (int y = min_vec_(1); y <= max_vec_(1); ++y)
        for (int z = std::max(min_vec_(2), vis_lower_bound); z <= std::min(max_vec_(2), vis_upper_bound); ++z)
          if (occ[(size_t)x * gyz + (size_t)y * grid_size_(2) + z] > min_occupancy_log_) {
            geometry_msgs::Point32 p;
            p.x = (float)((x + 0.5) * resolution_ + origin_(0));
            p.y = (float)((y + 0.5) * resolution_ + origin_(1));
            p.z = (float)((z + 0.5) * resolution_ + origin_(2));
            m.points.push_back(p);
          }
  }
  void GetSliceMarker(visualization_msgs::Marker &m, int slice, int id, Eigen::Vector4d /*color*/, double max_dist) {
    m.header.frame_id = "world";
    m.id = id;
    m.type = visualization_msgs::Marker::POINTS;
    m.action = visualization_msgs::Marker::MODIFY;
    m.scale.x = m.scale.y = m.scale.z = resolution_;
    m.pose.orientation.w = 1; m.pose.orientation.x = m.pose.orientation.y = m.pose.orientation.z = 0;
    m.points.clear();
    m.colors.clear();
    if (slice < 0 || slice >= grid_size_(2)) return;
    std::vector<double> dist((size_t)grid_total_size_);
    check(fiesta_export_distance(h_, dist.data()), "export_distance");
    const int gyz = grid_size_(1) * grid_size_(2);
    for (int x = min_vec_(0); x <= max_vec_(0); ++x)
      for (int y = min_vec_(1); y <= max_vec_(1); ++y) {
        const double d = dist[(size_t)x * gyz + (size_t)y * grid_size_(2) + slice];
        if (d < 0 || d >= FIESTA_INFINITY) continue;
        geometry_msgs::Point p;
        p.x = (x + 0.5) * resolution_ + origin_(0); p.y = (y + 0.5) * resolution_ + origin_(1); p.z = (slice + 0.5) * resolution_ + origin_(2);
        m.points.push_back(p);
        m.colors.push_back(Rainbow(d <= max_dist ? d / max_dist : 1));
      }
  }

 private:
  // HSV rainbow with s = v = 1 (same mapping as the reference's RainbowColorMap, ESDFMap.cpp:584-637)
  static std_msgs::ColorRGBA Rainbow(double h) {
    std_msgs::ColorRGBA c;
    c.a = 1;
    h = (h - std::floor(h)) * 6;
    const int i = (int)std::floor(h);
    double f = h - i;
    if (!(i & 1)) f = 1 - f;
    const float n = (float)(1 - f);
    const float lut[7][3] = {{1, n, 0}, {n, 1, 0}, {0, 1, n}, {0, n, 1}, {n, 0, 1}, {1, 0, n}, {1, n, 0}};
    const int k = (i >= 0 && i <= 6) ? i : 0;
    c.r = lut[k][0]; c.g = lut[k][1]; c.b = lut[k][2];
    return c;
  }
};

}  // namespace fiesta
#endif  // ESDF_MAP_H
