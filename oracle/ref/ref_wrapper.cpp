// TEST INFRASTRUCTURE — C wrapper around the *reference's own* classes.
//
// This file is compiled together with the UNMODIFIED reference translation
// units /root/reference/src/FOVUndistorter.cpp and PhotometricUndistorter.cpp
// (never copied into this repo) against the stand-in headers in oracle/shim/,
// producing oracle/_ref/libmdc_oracle_ref*.so (see oracle/Makefile).  It is the
// executable oracle of SURVEY.md §8c: the parity tests compare the CUDA path
// against it, and bench.py's cpu_baseline / --impl reference legs time it.
// Only tests/, __graft_entry__.smoke() and bench.py may load it.
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <cstring>
#include <cstdint>
#include <opencv2/core/core.hpp>
#include "Eigen/Core"
// The private tables (remapX/remapY, vignetteMapInv) are read for bit-compare;
// class layout does not depend on access specifiers, the reference TUs
// themselves are compiled without this define.
#define private public
#include "FOVUndistorter.h"
#include "PhotometricUndistorter.h"
#undef private

extern "C" {

// ---------------------------------------------------------------- FOV rectifier
void* oref_fov_create(const char* camera_txt) { return new UndistorterFOV(camera_txt); }
void oref_fov_destroy(void* h) { delete (UndistorterFOV*)h; }
int oref_fov_valid(void* h) { return ((UndistorterFOV*)h)->isValid() ? 1 : 0; }
void oref_fov_dims(void* h, int* dims4) {
    UndistorterFOV* u = (UndistorterFOV*)h;
    dims4[0] = u->getInputDims()[0]; dims4[1] = u->getInputDims()[1];
    dims4[2] = u->getOutputDims()[0]; dims4[3] = u->getOutputDims()[1];
}
const float* oref_fov_remap_x(void* h) { return ((UndistorterFOV*)h)->remapX; }
const float* oref_fov_remap_y(void* h) { return ((UndistorterFOV*)h)->remapY; }
void oref_fov_K(void* h, float* krect9, float* korg9) {
    UndistorterFOV* u = (UndistorterFOV*)h;
    Eigen::Matrix3f a = u->getK_rect(), b = u->getK_org();
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { krect9[r * 3 + c] = a(r, c); korg9[r * 3 + c] = b(r, c); }
}
float oref_fov_omega(void* h) { return ((UndistorterFOV*)h)->getOmega(); }
void oref_fov_original_calibration(void* h, float* v5) {
    Eigen::VectorXf v = ((UndistorterFOV*)h)->getOriginalCalibration();
    for (int i = 0; i < 5; i++) v5[i] = v[i];
}
void oref_fov_distort(void* h, float* x, float* y, int n) { ((UndistorterFOV*)h)->distortCoordinates(x, y, n); }
void oref_fov_undistort_f32(void* h, const float* in, float* out, int n_in, int n_out) {
    ((UndistorterFOV*)h)->undistort<float>(in, out, n_in, n_out);
}
void oref_fov_undistort_u8(void* h, const unsigned char* in, float* out, int n_in, int n_out) {
    ((UndistorterFOV*)h)->undistort<unsigned char>(in, out, n_in, n_out);
}

// ---------------------------------------------------------- photometric un-mapper
void* oref_photo_create(const char* pcalib, const char* vignette, int w, int h) {
    return new PhotometricUndistorter(std::string(pcalib), std::string(vignette), w, h);
}
void oref_photo_destroy(void* h) { delete (PhotometricUndistorter*)h; }
float* oref_photo_ginv(void* h) { return ((PhotometricUndistorter*)h)->getGInv(); }
float* oref_photo_g(void* h) { return ((PhotometricUndistorter*)h)->getG(); }
int oref_photo_valid_vignette(void* h) { return ((PhotometricUndistorter*)h)->validVignette ? 1 : 0; }
int oref_photo_valid_gamma(void* h) { return ((PhotometricUndistorter*)h)->validGamma ? 1 : 0; }
const float* oref_photo_vignette_map(void* h) { return ((PhotometricUndistorter*)h)->vignetteMap; }
const float* oref_photo_vignette_map_inv(void* h) { return ((PhotometricUndistorter*)h)->vignetteMapInv; }
void oref_photo_unmap(void* h, unsigned char* in, float* out, int n, int gamma, int vignette, int kill) {
    ((PhotometricUndistorter*)h)->unMapImage(in, out, n, gamma != 0, vignette != 0, kill != 0);
}

// ----------------------------------------------------- CPU baseline timing loops
// The as-shipped per-frame sequence of DatasetReader::getImage (photo + rect
// mode, BenchmarkDatasetReader.h:222-223): unMapImage into a temp buffer, then
// undistort<float>.  `threads` workers with private buffers (both calls are
// re-entrant); each worker handles frames t, t+threads, ...
// Decode and the per-frame ExposureImage allocation are excluded on both sides.
// Returns wall seconds; per-stage seconds of worker 0 in stage_s[0..1].
double oref_time_frames(void* hf, void* hp, const unsigned char* frames, int n_distinct,
                        int n_frames, int threads, int gamma, int vignette, int kill, double* stage_s) {
    UndistorterFOV* u = (UndistorterFOV*)hf;
    PhotometricUndistorter* p = (PhotometricUndistorter*)hp;
    const int n_in = u->getInputDims()[0] * u->getInputDims()[1];
    const int n_out = u->getOutputDims()[0] * u->getOutputDims()[1];
    if (threads < 1) threads = 1;
    std::vector<std::vector<float> > tmp(threads, std::vector<float>(n_in)), out(threads, std::vector<float>(n_out));
    std::vector<double> s0(threads, 0.0), s1(threads, 0.0);
    auto worker = [&](int t) {
        for (int f = t; f < n_frames; f += threads) {
            unsigned char* src = const_cast<unsigned char*>(frames) + (size_t)(f % n_distinct) * n_in;
            auto a = std::chrono::steady_clock::now();
            p->unMapImage(src, &tmp[t][0], n_in, gamma != 0, vignette != 0, kill != 0);
            auto b = std::chrono::steady_clock::now();
            u->undistort<float>(&tmp[t][0], &out[t][0], n_in, n_out);
            auto c = std::chrono::steady_clock::now();
            s0[t] += std::chrono::duration<double>(b - a).count();
            s1[t] += std::chrono::duration<double>(c - b).count();
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.push_back(std::thread(worker, t));
        for (size_t t = 0; t < pool.size(); t++) pool[t].join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stage_s) { stage_s[0] = s0[0]; stage_s[1] = s1[0]; }
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
