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
#include <cstring>
#include <memory>
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
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
// A small image container: enough of cv::Mat for FeatureTracker (8-bit single-channel frames and masks, shared buffers, clone, at<>).
class Mat {
  public:
    int rows, cols, type_;
    std::shared_ptr<std::vector<unsigned char>> buf;
    unsigned char *data;
    Mat() : rows(0), cols(0), type_(0), data(nullptr) {}
    Mat(int r, int c, int t) : rows(r), cols(c), type_(t), buf(new std::vector<unsigned char>((size_t)r * c * 8, 0)), data(buf->data()) {}
    Mat(int r, int c, int t, const Scalar &s) : rows(r), cols(c), type_(t), buf(new std::vector<unsigned char>((size_t)r * c * 8, (unsigned char)s.v[0])), data(buf->data()) {}
    Mat(int r, int c, int t, void *external) : rows(r), cols(c), type_(t), data((unsigned char *)external) {}      // wraps caller memory (no copy)
    Mat(Size s, int t) : Mat(s.height, s.width, t) {}
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    static Mat zeros(Size s, int t) { return Mat(s, t); }
    static Mat eye(int r, int c, int t) { return Mat(r, c, t); }
    template <typename T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + ((size_t)y * cols + x) * sizeof(T)); }
    template <typename T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + ((size_t)y * cols + x) * sizeof(T)); }
    template <typename T> T &at(int i) { return *reinterpret_cast<T *>(data + (size_t)i * sizeof(T)); }
    template <typename T> const T &at(int i) const { return *reinterpret_cast<const T *>(data + (size_t)i * sizeof(T)); }
    template <typename T> T &at(const Point_<int> &p) { return at<T>(p.y, p.x); }
    template <typename T> T &at(const Point_<float> &p) { return at<T>((int)lrintf(p.y), (int)lrintf(p.x)); }       // Point2f -> Point: saturate_cast rounds
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    Mat clone() const { Mat m(rows, cols, type_); if (data) memcpy(m.data, data, (size_t)rows * cols * (type_ == 0 ? 1 : 8)); return m; }
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
struct TermCriteria { enum { COUNT = 1, MAX_ITER = 1, EPS = 2 }; int type, maxCount; double epsilon; TermCriteria(int t = COUNT + EPS, int c = 30, double e = 0.01) : type(t), maxCount(c), epsilon(e) {} };
enum { OPTFLOW_USE_INITIAL_FLOW = 4, FM_RANSAC = 8 };
// The three OpenCV routines FeatureTracker::trackImage() computes with are NOT restated here: they are forwarded to callbacks that the test
// installs and that run the real cv2 (ref_glue.cpp: ref_set_cv_callbacks) -- the reference's tracker code then runs on its own third-party library.
typedef void (*LkCallback)(const unsigned char *prev, const unsigned char *next, int rows, int cols, int n, const float *prev_pts, float *next_pts, unsigned char *status,
                           float *err, int win, int max_level, int crit_count, double crit_eps, int flags);
typedef int (*GfttCallback)(const unsigned char *img, int rows, int cols, const unsigned char *mask, int max_corners, double quality, double min_dist, float *out, int cap);
typedef void (*CircleCallback)(unsigned char *img, int rows, int cols, int cx, int cy, int radius, int color, int thickness);
inline LkCallback g_lk_cb = nullptr;
inline GfttCallback g_gftt_cb = nullptr;
inline CircleCallback g_circle_cb = nullptr;
inline void calcOpticalFlowPyrLK(const Mat &prev, const Mat &next, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts, std::vector<uchar> &status,
                                 std::vector<float> &err, Size winSize = Size(21, 21), int maxLevel = 3, TermCriteria criteria = TermCriteria(), int flags = 0) {
    const int n = (int)prevPts.size();
    if (!(flags & OPTFLOW_USE_INITIAL_FLOW)) nextPts = prevPts; else nextPts.resize(n);
    status.assign(n, 0); err.assign(n, 0.f);
    if (n && g_lk_cb) g_lk_cb(prev.data, next.data, prev.rows, prev.cols, n, &prevPts[0].x, &nextPts[0].x, status.data(), err.data(), winSize.width, maxLevel, criteria.maxCount, criteria.epsilon, flags);
}
inline void goodFeaturesToTrack(const Mat &img, std::vector<Point2f> &corners, int maxCorners, double quality, double minDist, const Mat &mask = Mat()) {
    std::vector<float> out(2 * (size_t)(maxCorners > 0 ? maxCorners : 4096));
    const int n = g_gftt_cb ? g_gftt_cb(img.data, img.rows, img.cols, mask.empty() ? nullptr : mask.data, maxCorners, quality, minDist, out.data(), (int)out.size() / 2) : 0;
    corners.resize(n);
    for (int i = 0; i < n; i++) corners[i] = Point2f(out[2 * i], out[2 * i + 1]);
}
inline void circle(Mat &img, Point_<float> c, int radius, const Scalar &color, int thickness = 1, int = 8, int = 0) {
    if (g_circle_cb && img.type() == 0) g_circle_cb(img.data, img.rows, img.cols, (int)lrintf(c.x), (int)lrintf(c.y), radius, (int)color.v[0], thickness);
}
inline void cvtColor(const Mat &, Mat &, int) {}
inline void hconcat(const Mat &, const Mat &, Mat &) {}
inline void arrowedLine(Mat &, Point_<float>, Point_<float>, const Scalar &, int = 1, int = 8, int = 0, double = 0.1) {}
template <typename A, typename B> inline Mat findFundamentalMat(const A &, const B &, int, double, double, std::vector<uchar> &status) { status.clear(); return Mat(); }
}  // namespace cv
#define CV_GRAY2RGB 8
inline int cvRound(double v) { return (int)lrint(v); }
