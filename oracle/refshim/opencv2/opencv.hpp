// Stand-in for <opencv2/opencv.hpp> (TEST INFRASTRUCTURE): only what estimator/feature_manager.{h,cpp} names, so that the file compiles.
// FeatureManager::solvePoseByPnP / initFramePoseByPnP (cv::solvePnP) are NOT exercised through oracle/_ref: solvePnP below reports failure.
#pragma once
#include <vector>
#include <set>
#include <map>
#include <list>
namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point3f { float x, y, z; Point3f() : x(0), y(0), z(0) {} Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };
class Mat { public: Mat() {} };
template <typename T> class Mat_ : public Mat {
  public:
    Mat_(int, int) {}
    struct Comma { Comma &operator,(T) { return *this; } operator Mat() const { return Mat(); } };
    Comma operator<<(T) { return Comma(); }
};
template <typename A> inline void eigen2cv(const A &, Mat &) {}
template <typename A> inline void cv2eigen(const Mat &, A &) {}
inline void Rodrigues(const Mat &, Mat &) {}
template <typename A, typename B> inline bool solvePnP(const A &, const B &, const Mat &, const Mat &, Mat &, Mat &, bool) { return false; }
}  // namespace cv
