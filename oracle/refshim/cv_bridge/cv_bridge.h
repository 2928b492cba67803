#pragma once
#include "../ros_stub.h"
#include "../opencv2/opencv.hpp"
namespace cv_bridge {
struct CvImage { std_msgs::Header header; std::string encoding; cv::Mat image; sensor_msgs::ImagePtr toImageMsg() const { return sensor_msgs::ImagePtr(); } CvImage() {} CvImage(const std_msgs::Header &h, const std::string &e, const cv::Mat &m) : header(h), encoding(e), image(m) {} };
typedef std::shared_ptr<CvImage> CvImagePtr; typedef std::shared_ptr<const CvImage> CvImageConstPtr;
}  // namespace cv_bridge
