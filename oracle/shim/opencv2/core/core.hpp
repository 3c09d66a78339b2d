// Test-infrastructure stand-in for the handful of OpenCV declarations that the
// reference's image operators touch.  NOT OpenCV, NOT product
// code: it exists only so that /root/reference/src/{FOVUndistorter,
// PhotometricUndistorter}.cpp and the compat headers can be compiled in a
// container without OpenCV.  Pixel decode is delegated to a registry the test
// harness fills (with pixels decoded by the real cv2 wheel) or to a PGM reader.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cassert>
#include <cmath>
#ifdef MDC_SHIM_WITH_MATH_H
#include <math.h>   // pulls the float overloads of tan/sqrt into the global namespace (g++ >= 6)
#endif
#include <string>
#include <vector>
#include <memory>
#include <iostream>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_AA 16
#define CV_8UC3 16
#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_LOAD_IMAGE_GRAYSCALE 0

typedef unsigned char uchar;
typedef unsigned short ushort;

namespace cv {

inline int shim_elem_size(int type) { return type == CV_8U ? 1 : (type == CV_16U ? 2 : (type == CV_8UC3 ? 3 : (type == CV_64F ? 8 : 4))); }

struct Vec3b {
    unsigned char v[3];
    Vec3b() { v[0] = v[1] = v[2] = 0; }
    Vec3b(unsigned char a, unsigned char b, unsigned char c) { v[0] = a; v[1] = b; v[2] = c; }
};
struct Scalar { double v; Scalar(double a = 0) : v(a) {} Scalar(double a, double, double) : v(a) {} };
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point { int x, y; Point() : x(0), y(0) {} Point(int a, int b) : x(a), y(b) {} };

class Mat {
public:
    int rows, cols;
    unsigned char* data;

    Mat() : rows(0), cols(0), data(0), type_(CV_8U) {}
    Mat(int r, int c, int t) : rows(r), cols(c), data(0), type_(t) {
        own_.reset(new std::vector<unsigned char>((size_t)r * c * shim_elem_size(t)));
        data = own_->empty() ? 0 : &(*own_)[0];
    }
    Mat(int r, int c, int t, void* ext) : rows(r), cols(c), data((unsigned char*)ext), type_(t) {}

    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    template <typename T> T& at(int i) { return ((T*)data)[i]; }
    template <typename T> const T& at(int i) const { return ((const T*)data)[i]; }
    template <typename T> T& at(int r, int c) { return ((T*)data)[(size_t)r * cols + c]; }
    Mat& setTo(const Scalar& s) {
        if (type_ == CV_32F) for (size_t i = 0; i < (size_t)rows * cols; i++) at<float>((int)i) = (float)s.v;
        return *this;
    }

    // convertTo: dst = saturate(round-half-even(src * alpha + beta)); CV_32F source, CV_8U / CV_16U destination only
    void convertTo(Mat& dst, int type, double alpha = 1, double beta = 0) const;

    Mat operator*(double s) const {
        Mat m(rows, cols, type_);
        if (type_ == CV_32F)
            for (size_t i = 0; i < (size_t)rows * cols; i++) m.at<float>((int)i) = (float)(at<float>((int)i) * s);
        return m;
    }

private:
    int type_;
    std::shared_ptr<std::vector<unsigned char> > own_;
};

// implemented in oracle/shim/shim_impl.cpp
Mat imread(const std::string& path, int flags);
Mat imdecode(const Mat& buf, int flags);
bool imwrite(const std::string& path, const Mat& m);
void imshow(const std::string& name, const Mat& m);
int waitKey(int ms);
// stand-in: ignores the points and returns the next 3x3 homography (CV_64F) of the file named by $MDC_SHIM_HOMOGRAPHIES
Mat findHomography(const std::vector<Point2f>& src, const std::vector<Point2f>& dst);
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8) {}

}  // namespace cv

// registry used by the test harness: pixels decoded elsewhere (cv2) for `path`
extern "C" void mdc_shim_register_image(const char* path, int rows, int cols, int type, const void* pixels);
extern "C" void mdc_shim_clear_images();
