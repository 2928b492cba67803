// Stand-in for the OpenCV C++ headers (TEST INFRASTRUCTURE): only the NAMES that estimator/feature_manager.{h,cpp} and
// camera_models/src/camera_models/{Camera,PinholeCamera}.cc mention, so that those files compile unmodified.  Nothing here computes:
// the functions that would need real OpenCV (cv::solvePnP, cv::findHomography, cv::FileStorage I/O, cv::Mat pixels ...) report failure
// or do nothing, and the code paths that call them are NOT exercised through oracle/_ref (only Eigen-only functions are: triangulation,
// depth shift, liftProjective / spaceToPlane / distortion).
#pragma once
#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>
#include <cmath>
#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 5
#define CV_8UC1 0
typedef unsigned char uchar;
namespace cv {
typedef unsigned char uchar;
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T x_, T y_) : x(x_), y(y_) {} };
typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point_<int> Point2i; typedef Point2i Point;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {} };
typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} bool operator==(const Size &o) const { return width == o.width && height == o.height; } };
class Mat {
  public:
    int rows, cols;
    Mat() : rows(0), cols(0) {}
    Mat(int r, int c, int) : rows(r), cols(c) {}
    Mat(Size s, int) : rows(s.height), cols(s.width) {}
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    static Mat zeros(Size s, int t) { return Mat(s, t); }
    static Mat eye(int r, int c, int t) { return Mat(r, c, t); }
    template <typename T> T &at(int, int) { static T dummy; return dummy; }
    template <typename T> const T &at(int, int) const { static T dummy; return dummy; }
    template <typename T> T &at(int) { static T dummy; return dummy; }
    template <typename T> const T &at(int) const { static T dummy; return dummy; }
    bool empty() const { return rows == 0 || cols == 0; }
    Mat clone() const { return *this; }
    Mat t() const { return *this; }
    Mat inv() const { return *this; }
    Mat operator*(const Mat &) const { return *this; }
};
template <typename T> class Mat_ : public Mat {
  public:
    Mat_(int r, int c) : Mat(r, c, 0) {}
    struct Comma { Comma &operator,(T) { return *this; } operator Mat() const { return Mat(); } };
    Comma operator<<(T) { return Comma(); }
};
typedef const Mat &InputArray;
struct OutputArrayStub {                                           // cv::OutputArray as Camera.cc uses it
    Mat *m;
    OutputArrayStub() : m(nullptr) {}
    OutputArrayStub(Mat &x) : m(&x) {}
    OutputArrayStub(const Mat &) : m(nullptr) {}                    // cv::noArray()
    bool needed() const { return m != nullptr; }
    void create(int r, int c, int t) const { if (m) *m = Mat(r, c, t); }
    Mat getMat() const { return m ? *m : Mat(); }
};
typedef OutputArrayStub OutputArray;
inline Mat noArray() { return Mat(); }
template <typename T> inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline double norm(const Point_<T> &a) { return std::sqrt((double)a.x * a.x + (double)a.y * a.y); }
struct FileNode {
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    bool isNone() const { return true; }
    operator int() const { return 0; } operator double() const { return 0.0; } operator float() const { return 0.f; } operator std::string() const { return std::string(); }
    template <typename T> void operator>>(T &) const {}
};
class FileStorage {
  public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    template <typename T> FileStorage &operator<<(const T &) { return *this; }
};
enum { DECOMP_LU = 0, DECOMP_NORMAL = 16, INTER_LINEAR = 1, SOLVEPNP_ITERATIVE = 0 };
template <typename A> inline void eigen2cv(const A &, Mat &) {}
template <typename A> inline void cv2eigen(const Mat &, A &) {}
inline void Rodrigues(const Mat &, Mat &) {}
template <typename A, typename B> inline bool solvePnP(const A &, const B &, const Mat &, const Mat &, Mat &, Mat &, bool = false, int = 0) { return false; }
template <typename A, typename B> inline Mat findHomography(const A &, const B &) { return Mat(); }
inline bool solve(const Mat &, const Mat &, Mat &, int = 0) { return false; }
inline void convertMaps(const Mat &, const Mat &, Mat &, Mat &, int, bool = false) {}
template <typename A, typename B> inline void projectPoints(const A &, const Mat &, const Mat &, const Mat &, const Mat &, B &) {}
}  // namespace cv
