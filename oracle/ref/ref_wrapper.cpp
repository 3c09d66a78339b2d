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

// ----------------------------------------------------- all-core CPU arm: persistent worker pool
// What a maintainer would do to run the reference's per-frame path on every core: one worker per hardware thread, pinned,
// with PRIVATE temp/output buffers that live as long as the pool and are first-touched by the worker that uses them, one
// replica of the two reference objects (remap tables, vignette map) and of the input frames per NUMA node, built by a
// thread running on that node, and a start barrier so that the timed region contains frame work only (steady_clock inside
// the library: from releasing the workers to the last one finishing).  Each frame is the reference's own sequence
// (BenchmarkDatasetReader.h:222-223): unMapImage into the temp buffer, then undistort<float>.
}  // extern "C" (pool internals below are C++)

#include <sched.h>
#include <unistd.h>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <mutex>

namespace {

struct NodeReplica {
    UndistorterFOV* fov;
    PhotometricUndistorter* photo;
    std::vector<unsigned char> frames;
    NodeReplica() : fov(0), photo(0) {}
};

struct RefPool {
    std::string camera, pcalib, vignette;
    int w, h, n_in, n_out, n_distinct, flags[3];
    const unsigned char* src_frames;
    std::vector<int> cpus;              // cpu each worker is pinned to
    std::vector<int> node_of_worker;
    std::vector<NodeReplica> nodes;
    std::vector<std::thread> threads;
    std::mutex mu, build_mu;
    std::condition_variable cv_go, cv_done;
    long generation;
    int frames_per_worker, pending, ready;
    bool quit, failed;
    std::vector<double> busy_s;
    std::chrono::steady_clock::time_point t_last_done;
    RefPool() : generation(0), frames_per_worker(0), pending(0), ready(0), quit(false), failed(false) {}
};

int node_of_cpu(int cpu) {
    for (int node = 0; node < 64; ++node) {
        char path[128];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        std::ifstream f(path);
        if (!f.good()) continue;
        std::string list;
        std::getline(f, list);
        size_t pos = 0;
        while (pos < list.size()) {
            int a = 0, b = 0, used = 0;
            if (sscanf(list.c_str() + pos, "%d-%d%n", &a, &b, &used) == 2) { if (cpu >= a && cpu <= b) return node; }
            else if (sscanf(list.c_str() + pos, "%d%n", &a, &used) == 1) { if (cpu == a) return node; }
            else break;
            pos += used;
            if (pos < list.size() && list[pos] == ',') ++pos;
        }
    }
    return 0;
}

void pool_worker(RefPool* P, int t) {
    const char* pin = getenv("MDC_REF_PIN");      // "0": leave placement to the kernel (A/B of the pinning itself)
    if (!(pin && pin[0] == '0')) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(P->cpus[t], &set);
        sched_setaffinity(0, sizeof set, &set);
    }
    const int node = P->node_of_worker[t];
    {   // the first worker of a node builds that node's replica (its pages are first-touched here, on this node)
        std::lock_guard<std::mutex> lk(P->build_mu);
        NodeReplica& R = P->nodes[node];
        if (!R.fov) {
            R.fov = new UndistorterFOV(P->camera.c_str());
            R.photo = new PhotometricUndistorter(P->pcalib, P->vignette, P->w, P->h);
            R.frames.assign(P->src_frames, P->src_frames + (size_t)P->n_distinct * P->n_in);
            if (!R.fov->isValid() || !R.photo->validVignette) P->failed = true;
        }
    }
    std::vector<float> tmp(P->n_in, 0.0f), out(P->n_out, 0.0f);      // private, first-touched by this (pinned) thread
    NodeReplica& R = P->nodes[node];
    long seen = 0;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        ++P->ready;
        P->cv_done.notify_all();
    }
    for (;;) {
        int n;
        {
            std::unique_lock<std::mutex> lk(P->mu);
            P->cv_go.wait(lk, [&] { return P->quit || P->generation != seen; });
            if (P->quit) return;
            seen = P->generation;
            n = P->frames_per_worker;
        }
        const auto a = std::chrono::steady_clock::now();
        for (int f = 0; f < n; ++f) {
            unsigned char* src = &R.frames[(size_t)((t + f) % P->n_distinct) * P->n_in];
            R.photo->unMapImage(src, &tmp[0], P->n_in, P->flags[0] != 0, P->flags[1] != 0, P->flags[2] != 0);
            R.fov->undistort<float>(&tmp[0], &out[0], P->n_in, P->n_out);
        }
        const auto b = std::chrono::steady_clock::now();
        {
            std::lock_guard<std::mutex> lk(P->mu);
            P->busy_s[t] = std::chrono::duration<double>(b - a).count();
            if (--P->pending == 0) { P->t_last_done = b; P->cv_done.notify_all(); }
        }
    }
}

}  // namespace

extern "C" {

// threads <= 0: one worker per CPU of the process's affinity mask.  Returns 0 on failure.
void* oref_pool_create(const char* camera_txt, const char* pcalib, const char* vignette, int w, int h, const unsigned char* frames,
                       int n_distinct, int threads, int gamma, int vignette_flag, int kill) {
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    std::vector<int> cpus;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
    if (cpus.empty()) cpus.push_back(0);
    if (threads <= 0) threads = (int)cpus.size();
    RefPool* P = new RefPool();
    P->camera = camera_txt; P->pcalib = pcalib; P->vignette = vignette;
    P->w = w; P->h = h; P->n_distinct = n_distinct; P->src_frames = frames;
    P->flags[0] = gamma; P->flags[1] = vignette_flag; P->flags[2] = kill;
    {   // geometry from a throw-away object (the replicas are built by the workers)
        UndistorterFOV probe(camera_txt);
        if (!probe.isValid()) { delete P; return 0; }
        P->n_in = probe.getInputDims()[0] * probe.getInputDims()[1];
        P->n_out = probe.getOutputDims()[0] * probe.getOutputDims()[1];
    }
    // placement of fewer workers than CPUs: "spread" deals them round-robin over the NUMA nodes (each node's CPUs in affinity
    // order), "compact" fills the CPUs in affinity order
    const char* spread_env = getenv("MDC_REF_SPREAD");
    if (spread_env && spread_env[0] == '1') {
        std::vector<std::vector<int> > by_node;
        for (size_t i = 0; i < cpus.size(); ++i) {
            const size_t nd = (size_t)node_of_cpu(cpus[i]);
            if (by_node.size() <= nd) by_node.resize(nd + 1);
            by_node[nd].push_back(cpus[i]);
        }
        std::vector<int> order;
        for (size_t k = 0; order.size() < cpus.size(); ++k)
            for (size_t nd = 0; nd < by_node.size(); ++nd)
                if (k < by_node[nd].size()) order.push_back(by_node[nd][k]);
        cpus.swap(order);
    }
    int max_node = 0;
    for (int t = 0; t < threads; ++t) {
        P->cpus.push_back(cpus[t % cpus.size()]);
        P->node_of_worker.push_back(node_of_cpu(P->cpus.back()));
        if (P->node_of_worker.back() > max_node) max_node = P->node_of_worker.back();
    }
    P->nodes.resize(max_node + 1);
    P->busy_s.assign(threads, 0.0);
    for (int t = 0; t < threads; ++t) P->threads.push_back(std::thread(pool_worker, P, t));
    {
        std::unique_lock<std::mutex> lk(P->mu);
        P->cv_done.wait(lk, [&] { return P->ready == threads; });
    }
    P->src_frames = 0;      // every node holds its own copy now
    if (P->failed) { extern void oref_pool_destroy(void*); oref_pool_destroy(P); return 0; }
    return P;
}

int oref_pool_threads(void* h) { return (int)((RefPool*)h)->threads.size(); }
int oref_pool_numa_nodes(void* h) {
    RefPool* P = (RefPool*)h;
    int n = 0;
    for (size_t i = 0; i < P->nodes.size(); ++i) n += P->nodes[i].fov != 0;
    return n;
}

// Every worker processes frames_per_worker frames.  Returns the seconds from releasing the workers to the last one finishing
// (steady_clock, inside this call); busy_minmax[0..1] = shortest / longest per-worker busy time.
double oref_pool_run(void* h, int frames_per_worker, double* busy_minmax) {
    RefPool* P = (RefPool*)h;
    std::chrono::steady_clock::time_point t0;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        P->frames_per_worker = frames_per_worker;
        P->pending = (int)P->threads.size();
        ++P->generation;
        t0 = std::chrono::steady_clock::now();
    }
    P->cv_go.notify_all();
    std::unique_lock<std::mutex> lk(P->mu);
    P->cv_done.wait(lk, [&] { return P->pending == 0; });
    if (busy_minmax) {
        busy_minmax[0] = busy_minmax[1] = P->busy_s[0];
        for (size_t t = 1; t < P->busy_s.size(); ++t) {
            if (P->busy_s[t] < busy_minmax[0]) busy_minmax[0] = P->busy_s[t];
            if (P->busy_s[t] > busy_minmax[1]) busy_minmax[1] = P->busy_s[t];
        }
    }
    return std::chrono::duration<double>(P->t_last_done - t0).count();
}

void oref_pool_destroy(void* h) {
    RefPool* P = (RefPool*)h;
    if (!P) return;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        P->quit = true;
    }
    P->cv_go.notify_all();
    for (size_t t = 0; t < P->threads.size(); ++t) P->threads[t].join();
    for (size_t i = 0; i < P->nodes.size(); ++i) { delete P->nodes[i].fov; delete P->nodes[i].photo; }
    delete P;
}

}  // extern "C"
