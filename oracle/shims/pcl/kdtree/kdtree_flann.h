// oracle shim: parameters.h hard-defines DEBUG, so ESDFMap::CheckWithGroundTruth (never called)
// must still compile; a brute-force nearest neighbour stands in for PCL's k-d tree.
#ifndef FIESTA_ORACLE_PCL_SHIM
#define FIESTA_ORACLE_PCL_SHIM
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZ { float x, y, z; PointXYZ() : x(0), y(0), z(0) {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {} };
template <typename P> struct PointCloud { typedef std::shared_ptr<PointCloud<P>> Ptr; unsigned width = 0, height = 0; std::vector<P> points; };
template <typename P> struct KdTreeFLANN {
  typename PointCloud<P>::Ptr c;
  void setInputCloud(const typename PointCloud<P>::Ptr &p) { c = p; }
  int nearestKSearch(const P &q, int, std::vector<int> &idx, std::vector<float> &d2) {
    float best = 3.4e38f; int bi = -1;
    for (std::size_t i = 0; i < c->points.size(); ++i) {
      float dx = c->points[i].x - q.x, dy = c->points[i].y - q.y, dz = c->points[i].z - q.z;
      float s = dx * dx + dy * dy + dz * dz; if (s < best) { best = s; bi = (int)i; }
    }
    idx[0] = bi; d2[0] = best; return bi >= 0;
  }
};
}
#endif
