// TEST SHIM — not OpenCV.  The smallest subset of the cv:: value types that the my_slam adapter layer
// (monocular-visual-odometry_b200/my_slam_adapter/) touches, with OpenCV's field names, memory layout and call
// syntax, so that the adapters can be compiled and exercised in an image that has no OpenCV C++ headers.
// With a real OpenCV on the include path this directory is simply not used.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };

struct KeyPoint {           // features2d: same field order and size (28 bytes) as cv::KeyPoint
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(pt_), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct DMatch {             // 16 bytes, as cv::DMatch
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0;
  DMatch() {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char *data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data((unsigned char *)ext), type_(type) { step = (size_t)c * elemSize(); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    buf_ = std::shared_ptr<unsigned char>((unsigned char *)std::calloc((size_t)r * step + 64, 1), std::free);
    data = buf_.get();
  }
  Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; ++r) std::memcpy(m.data + r * m.step, data + r * step, m.step); return m; }
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  int depth() const { return type_ & 7; }
  size_t elemSize() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()] * channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  template <class T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
  template <class T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
  template <class T> T &at(int r, int c) { return ptr<T>(r)[c]; }
  template <class T> const T &at(int r, int c) const { return ptr<T>(r)[c]; }
 protected:
  int type_ = 0;
  std::shared_ptr<unsigned char> buf_;
};

template <class T> struct MatDepth_;
template <> struct MatDepth_<unsigned char> { enum { value = CV_8U }; };
template <> struct MatDepth_<int> { enum { value = CV_32S }; };
template <> struct MatDepth_<float> { enum { value = CV_32F }; };
template <> struct MatDepth_<double> { enum { value = CV_64F }; };

template <class T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(MatDepth_<T>::value, 1)) {}
  Mat_(const Mat &m) : Mat(m) {}
  T &operator()(int r, int c) { return this->template at<T>(r, c); }
  const T &operator()(int r, int c) const { return this->template at<T>(r, c); }
};
typedef Mat_<unsigned char> Mat1b;

}  // namespace cv
