// SHIM, not OpenCV -- TEST INFRASTRUCTURE.  Declares just enough of the cv:: names that the reference's
// postprocess.cpp / postprocess.cu / preprocess.cu / common.hpp mention, so that those files COMPILE
// from where they lie under /root/reference (oracle/Makefile, target `ref`).  Only the pieces the
// checked functions execute are functional: cv::Mat as a plain (rows, cols, type, data) view and
// cv::invertAffineTransform (restated, pinned against cv2 4.13 in tests/test_oracle_cpu.py).
// Drawing / resizing entry points are declared and abort if ever called.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <numeric>
#include <memory>
#include <functional>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <math.h>
#include <sstream>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_PI 3.1415926535897932384626433832795

namespace cv {
[[noreturn]] inline void shim_abort(const char* what) {
    std::fprintf(stderr, "opencv shim: %s is not implemented (test infrastructure)\n", what);
    std::abort();
}
template <typename T>
struct Point_ {
    T x{}, y{};
    Point_() = default;
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
using Point = Point_<int>;
using Point2f = Point_<float>;
template <typename T>
struct Size_ {
    T width{}, height{};
    Size_() = default;
    Size_(T w, T h) : width(w), height(h) {}
};
using Size = Size_<int>;
template <typename T>
struct Rect_ {
    T x{}, y{}, width{}, height{};
    Rect_() = default;
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Point_<T> tl() const { return {x, y}; }
    Point_<T> br() const { return {x + width, y + height}; }
    T area() const { return width * height; }
};
using Rect = Rect_<int>;
struct Scalar {
    double val[4]{};
    Scalar() = default;
    Scalar(double a, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};
template <typename T, int N>
struct Vec {
    T val[N]{};
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
using Vec3b = Vec<unsigned char, 3>;
using Vec3f = Vec<float, 3>;

class Mat {
   public:
    int rows{0}, cols{0}, type_{0};
    unsigned char* data{nullptr};
    size_t step{0};
    std::vector<unsigned char> own;
    Mat() = default;
    Mat(int r, int c, int type, void* d) : rows(r), cols(c), type_(type), data(static_cast<unsigned char*>(d)) {
        step = (size_t)c * elemSize();
    }
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
        step = (size_t)c * elemSize();
        own.assign(step * r, 0);
        data = own.data();
    }
    Mat(int r, int c, int type, const Scalar&) : Mat(r, c, type) {}
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), type_(o.type_), data(o.data), step(o.step), own(o.own) {
        if (!own.empty()) data = own.data();
    }
    Mat& operator=(const Mat& o) {
        rows = o.rows; cols = o.cols; type_ = o.type_; step = o.step; own = o.own;
        data = own.empty() ? o.data : own.data();
        return *this;
    }
    size_t elemSize() const { return type_ == CV_8UC3 ? 3 : (type_ == CV_32F ? 4 : 1); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
    template <typename T>
    T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T>
    const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T>
    T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T>
    const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    // scale_mask (yolov8/src/postprocess.cpp:207-226) crops then resizes: the shim only RECORDS the crop rectangle and the
    // target size (ref_v8_scale_mask_rect reads them back); pixel work stays out of the shim
    Mat operator()(const Rect& r) const {
        shim_last_crop()[0] = r.x; shim_last_crop()[1] = r.y; shim_last_crop()[2] = r.width; shim_last_crop()[3] = r.height;
        return Mat(r.height > 0 ? r.height : 0, r.width > 0 ? r.width : 0, type_, data);
    }
    static int* shim_last_crop() { static thread_local int v[6] = {0, 0, 0, 0, 0, 0}; return v; }
    Mat clone() const { shim_abort("Mat::clone"); }
    bool empty() const { return data == nullptr; }
    Size size() const { return Size(cols, rows); }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    int type() const { return type_; }
    void copyTo(Mat&) const { shim_abort("Mat::copyTo"); }
    void copyTo(Mat&&) const { shim_abort("Mat::copyTo"); }
    void convertTo(Mat&, int, double = 1, double = 0) const { shim_abort("Mat::convertTo"); }
    Mat& setTo(const Scalar&) { shim_abort("Mat::setTo"); }
};

enum { FONT_HERSHEY_PLAIN = 1, FONT_HERSHEY_SIMPLEX = 0, LINE_AA = 16, INTER_LINEAR = 1, BORDER_CONSTANT = 0 };

inline void rectangle(Mat&, Rect, const Scalar&, int = 1, int = 8, int = 0) { shim_abort("rectangle"); }
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { shim_abort("rectangle"); }
inline void putText(Mat&, const std::string&, Point, int, double, const Scalar&, int = 1, int = 8, bool = false) {
    shim_abort("putText");
}
inline Size getTextSize(const std::string&, int, double, int, int*) { shim_abort("getTextSize"); }
inline void resize(const Mat&, Mat& dst, Size s, double = 0, double = 0, int = 1) {
    Mat::shim_last_crop()[4] = s.width;
    Mat::shim_last_crop()[5] = s.height;
    dst = Mat(0, 0, 0, nullptr);
    dst.cols = s.width;
    dst.rows = s.height;
}
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { shim_abort("line"); }
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) { shim_abort("circle"); }
inline void polylines(Mat&, const std::vector<std::vector<Point>>&, bool, const Scalar&, int = 1, int = 8, int = 0) {
    shim_abort("polylines");
}
inline void polylines(Mat&, const std::vector<Point>&, bool, const Scalar&, int = 1, int = 8, int = 0) {
    shim_abort("polylines");
}
inline Mat imread(const std::string&, int = 1) { shim_abort("imread"); }
inline bool imwrite(const std::string&, const Mat&) { shim_abort("imwrite"); }
inline void exp(const Mat&, Mat&) { shim_abort("exp"); }

// OpenCV >= 4 CV_32F branch (softfloat determinant and A = M * float(D) products, double translation)
inline void invertAffineTransform(const Mat& M_, Mat& iM_) {
    const float* M = M_.ptr<float>(0);
    const size_t st = M_.step / sizeof(float), ist = iM_.step / sizeof(float);
    float* iM = iM_.ptr<float>(0);
    volatile float p0 = M[0] * M[st + 1], p1 = M[1] * M[st];
    volatile float det = p0 - p1;
    double D = (double)det;
    D = D != 0 ? 1. / D : 0;
    const float Df = (float)D;
    volatile float fA11 = M[st + 1] * Df, fA22 = M[0] * Df, fA12 = (-M[1]) * Df, fA21 = (-M[st]) * Df;
    const double A11 = fA11, A22 = fA22, A12 = fA12, A21 = fA21;
    const double b1 = -A11 * M[2] - A12 * M[st + 2];
    const double b2 = -A21 * M[2] - A22 * M[st + 2];
    iM[0] = (float)A11;
    iM[1] = (float)A12;
    iM[2] = (float)b1;
    iM[ist] = (float)A21;
    iM[ist + 1] = (float)A22;
    iM[ist + 2] = (float)b2;
}
}  // namespace cv
