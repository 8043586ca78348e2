// oracle shim: plain structs with the fields ESDFMap::GetSliceMarker writes.
#ifndef FIESTA_ORACLE_MARKER_SHIM
#define FIESTA_ORACLE_MARKER_SHIM
#include <string>
#include <vector>
namespace std_msgs { struct Header { std::string frame_id; }; struct ColorRGBA { float r, g, b, a; }; }
namespace geometry_msgs {
struct Point { double x, y, z; };
struct Point32 { float x, y, z; };
struct Quaternion { double x, y, z, w; };
struct Vector3 { double x, y, z; };
struct Pose { Point position; Quaternion orientation; };
}
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, POINTS = 8, TEXT_VIEW_FACING = 9 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2 };
  std_msgs::Header header; int id; int type; int action;
  geometry_msgs::Pose pose; geometry_msgs::Vector3 scale; std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points; std::vector<std_msgs::ColorRGBA> colors; std::string text;
};
}
#endif
