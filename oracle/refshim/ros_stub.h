// ros_stub.h -- TEST INFRASTRUCTURE.  Empty stand-ins for the ROS types that estimator/estimator.{h,cpp}, utility/visualization.h and
// utility/CameraPoseVisualization.h name (publishers, messages, tf, cv_bridge): enough for those files to compile unmodified; nothing
// is published anywhere.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "ros/assert.h"
namespace ros {
struct Time { double t; Time() : t(0) {} explicit Time(double s) : t(s) {} double toSec() const { return t; } static Time now() { return Time(); } };
struct Duration { explicit Duration(double = 0) {} void sleep() const {} };
struct Publisher { template <typename M> void publish(const M &) const {} int getNumSubscribers() const { return 0; } };
struct NodeHandle { template <typename M> Publisher advertise(const std::string &, int) { return Publisher(); } };
inline bool ok() { return true; }
}  // namespace ros
namespace std_msgs {
struct Header { unsigned seq; ros::Time stamp; std::string frame_id; Header() : seq(0) {} };
struct Float32 { float data; }; struct Bool { bool data; };
struct ColorRGBA { float r, g, b, a; };
}  // namespace std_msgs
namespace geometry_msgs {
struct Point { double x, y, z; }; struct Point32 { float x, y, z; }; struct Vector3 { double x, y, z; };
struct Quaternion { double x, y, z, w; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PointStamped { std_msgs::Header header; Point point; };
struct PoseWithCovariance { Pose pose; };
struct Twist { Vector3 linear, angular; }; struct TwistWithCovariance { Twist twist; };
}  // namespace geometry_msgs
namespace sensor_msgs {
struct ChannelFloat32 { std::string name; std::vector<float> values; };
struct PointCloud { std_msgs::Header header; std::vector<geometry_msgs::Point32> points; std::vector<ChannelFloat32> channels; };
typedef std::shared_ptr<const PointCloud> PointCloudConstPtr;
struct Imu { std_msgs::Header header; geometry_msgs::Vector3 angular_velocity, linear_acceleration; };
struct Image { std_msgs::Header header; unsigned height, width, step; std::string encoding; unsigned char is_bigendian; std::vector<unsigned char> data; };
typedef std::shared_ptr<const Image> ImageConstPtr; typedef std::shared_ptr<Image> ImagePtr;
namespace image_encodings { const std::string MONO8 = "mono8", BGR8 = "bgr8"; }
}  // namespace sensor_msgs
namespace nav_msgs {
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; };
}  // namespace nav_msgs
namespace visualization_msgs {
struct Marker {
    enum { LINE_STRIP = 4, LINE_LIST = 5, ADD = 0, SPHERE_LIST = 7, POINTS = 8, DELETE = 2 };
    std_msgs::Header header; std::string ns; int id, type, action; geometry_msgs::Pose pose; geometry_msgs::Vector3 scale; std_msgs::ColorRGBA color;
    ros::Duration lifetime; std::vector<geometry_msgs::Point> points; std::vector<std_msgs::ColorRGBA> colors;
    Marker() : id(0), type(0), action(0) {}
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs
