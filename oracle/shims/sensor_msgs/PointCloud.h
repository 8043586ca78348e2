#ifndef FIESTA_ORACLE_POINTCLOUD_SHIM
#define FIESTA_ORACLE_POINTCLOUD_SHIM
#include <visualization_msgs/Marker.h>
namespace sensor_msgs { struct PointCloud { std_msgs::Header header; std::vector<geometry_msgs::Point32> points; }; }
#endif
