// TEST INFRASTRUCTURE.  Stand-in for the aruco marker library (not installed), just enough for the reference's
// main_vignetteCalib.cpp to compile unmodified: the "detector" reports exactly one marker per image; the geometry the program
// derives from it comes from cv::findHomography, whose stand-in reads the homographies from a side file (shim_impl.cpp).
#pragma once
#include <vector>
#include "opencv2/core/core.hpp"

namespace aruco {
struct Marker : public std::vector<cv::Point2f> {
    Marker() : std::vector<cv::Point2f>(4) {}
};
class MarkerDetector {
public:
    void detect(const cv::Mat&, std::vector<Marker>& out) {
        out.clear();
        out.push_back(Marker());
    }
};
}  // namespace aruco
