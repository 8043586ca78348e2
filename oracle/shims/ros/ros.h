// oracle shim: parameters.h only names ros::NodeHandle in a declaration.
#ifndef FIESTA_ORACLE_ROS_SHIM
#define FIESTA_ORACLE_ROS_SHIM
namespace ros { class NodeHandle; }
#endif
