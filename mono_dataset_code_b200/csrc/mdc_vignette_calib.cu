// vignetteCalib's alternating optimiser on the GPU: main_vignetteCalib.cpp:395-585.
//
// State as the reference holds it when its loop starts (:395): n float images [wI*hI] (NaN = invalidated pixel), n
// plane-to-image maps p2x/p2y [gw*gh] (NaN = plane point not visible; finite entries keep all four bilinear taps inside the
// image, :352-356), planeColor [gw*gh], vignetteFactor [wI*hI].  Every array is device-resident, image-major, contiguous.
//
//   plane step (:400-446)     one thread per plane point, sequential over the images  ->  FF/FC accumulate in the reference's
//                             order, planeColor = FC/FF is BIT-IDENTICAL; E is a sum over 10^9 doubles (order-dependent).
//   vignette step (:458-533)  one thread per (image, plane point): bilinear scatter-add of 8 float terms into TT/CT with
//                             global fp32 atomics.  The reference adds them in (image, point) order; any parallel order
//                             differs in the last bits (compared to 1e-5 relative).  Then CT/TT, the maximum, the division.
//   smoothing (:542-566)      NaN-aware 3x3 mean, ping-pong buffers, bit-identical.
//
// All float arithmetic is spelled out with round-to-nearest intrinsics in the reference's evaluation order (no FMA).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "mdc_b200.h"
#include "mdc_internal.h"

#define VC_CHECK(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            mdc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return MDC_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

namespace {

struct Taps { float w00, w10, w01, w11; int base; };      // weights of bp[0], bp[1], bp[width], bp[1+width]

// the index/weight part of getInterpolatedElement (main_vignetteCalib.cpp:52-62)
__device__ __forceinline__ Taps vc_taps(float x, float y, int width) {
    const int ix = static_cast<int>(x), iy = static_cast<int>(y);
    const float dx = __fsub_rn(x, static_cast<float>(ix)), dy = __fsub_rn(y, static_cast<float>(iy));
    const float dxdy = __fmul_rn(dx, dy);
    Taps t;
    t.w11 = dxdy;
    t.w01 = __fsub_rn(dy, dxdy);
    t.w10 = __fsub_rn(dx, dxdy);
    t.w00 = __fadd_rn(__fsub_rn(__fsub_rn(1.0f, dx), dy), dxdy);
    t.base = ix + iy * width;
    return t;
}
// ... and its blend (:65-68): dxdy*bp[1+w] + (dy-dxdy)*bp[w] + (dx-dxdy)*bp[1] + (1-dx-dy+dxdy)*bp[0], left to right
__device__ __forceinline__ float vc_blend(const float* __restrict__ mat, const Taps& t, int width) {
    const float* bp = mat + t.base;
    float r = __fmul_rn(t.w11, __ldg(bp + 1 + width));
    r = __fadd_rn(r, __fmul_rn(t.w01, __ldg(bp + width)));
    r = __fadd_rn(r, __fmul_rn(t.w10, __ldg(bp + 1)));
    return __fadd_rn(r, __fmul_rn(t.w00, __ldg(bp)));
}

// CTA-wide reduction of (E, R) partials -> two global atomics per CTA
__device__ __forceinline__ void vc_flush_stats(double e, double r, double* __restrict__ stats) {
    __shared__ double se[32], sr[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = (blockDim.x + 31) >> 5;
    if (lane == 0) { se[warp] = e; sr[warp] = r; }
    __syncthreads();
    if (warp == 0) {
        e = lane < warps ? se[lane] : 0.0;
        r = lane < warps ? sr[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            e += __shfl_xor_sync(0xffffffffu, e, o);
            r += __shfl_xor_sync(0xffffffffu, r, o);
        }
        if (lane == 0) { atomicAdd(stats, e); atomicAdd(stats + 1, r); }
    }
}

// The reference's outlier test `abs(residual) > oth2` (:423, :481) on a double residual: with only `int abs(int)` visible to
// unqualified lookup (libstdc++ before GCC 6 - the reference's era - and the oracle build) the residual is truncated to int first;
// with <cmath>'s overloads in the global namespace it is fabs.  integer_abs != 0 reproduces the former: residuals are squares
// (>= 0) or NaN, and an out-of-range / NaN conversion yields INT_MIN on x86, which is never an outlier.
__device__ __forceinline__ bool vc_outlier(double residual, double oth2, int integer_abs) {
    if (!integer_abs) return fabs(residual) > oth2;
    return residual < 2147483648.0 && static_cast<double>(__double2int_rz(residual)) > oth2;
}

constexpr int kVcBatch = 4;      // images whose maps and taps are fetched together (memory-level parallelism)

// ---- plane step: for each plane point the optimum is sum(color*fac) / sum(fac*fac) over the images that see it
__global__ void __launch_bounds__(256) vc_plane_kernel(const float* __restrict__ images, const float* __restrict__ p2x,
                                                       const float* __restrict__ p2y, int n, int gwgh, int wI, size_t npx,
                                                       const float* __restrict__ vignette, float* __restrict__ plane_color,
                                                       double oth2, int integer_abs, double* __restrict__ stats) {
    double e_sum = 0.0, r_cnt = 0.0;
    const int stride = gridDim.x * blockDim.x;
    for (int pi = blockIdx.x * blockDim.x + threadIdx.x; pi < gwgh; pi += stride) {
        const float pc = plane_color[pi];
        float ff = 0.0f, fc = 0.0f;
        // software pipeline: the maps of the NEXT batch of images are requested before the taps of the current batch are gathered,
        // so a batch costs one memory round trip instead of two dependent ones
        float mx[kVcBatch], my[kVcBatch];
        auto fetch_maps = [&](int img0, float (&x)[kVcBatch], float (&y)[kVcBatch]) {
#pragma unroll
            for (int j = 0; j < kVcBatch; ++j) {
                const int img = img0 + j;
                const bool in = img < n;
                x[j] = in ? __ldg(p2x + static_cast<size_t>(img) * gwgh + pi) : __int_as_float(0x7fc00000);
                y[j] = in ? __ldg(p2y + static_cast<size_t>(img) * gwgh + pi) : 0.0f;
            }
        };
        fetch_maps(0, mx, my);
        for (int img0 = 0; img0 < n; img0 += kVcBatch) {
            float nx[kVcBatch], ny[kVcBatch];
            fetch_maps(img0 + kVcBatch, nx, ny);
            float color[kVcBatch], fac[kVcBatch];
#pragma unroll
            for (int j = 0; j < kVcBatch; ++j) {
                color[j] = fac[j] = 0.0f;
                if (!isnan(mx[j])) {                                                       // :414
                    const Taps t = vc_taps(mx[j], my[j], wI);
                    color[j] = vc_blend(images + static_cast<size_t>(img0 + j) * npx, t, wI);
                    fac[j] = vc_blend(vignette, t, wI);
                }
            }
#pragma unroll
            for (int j = 0; j < kVcBatch; ++j) {
                if (isnan(mx[j]) || isnan(fac[j]) || isnan(color[j])) continue;            // :414, :420-421
                const float d = __fsub_rn(color[j], __fmul_rn(pc, fac[j]));
                const double residual = static_cast<double>(__fmul_rn(d, d));
                if (vc_outlier(residual, oth2, integer_abs)) { e_sum += oth2; r_cnt += 1.0; continue; }      // :424-429
                ff = __fadd_rn(ff, __fmul_rn(fac[j], fac[j]));
                fc = __fadd_rn(fc, __fmul_rn(color[j], fac[j]));
                if (isnan(pc)) continue;
                e_sum += residual;
                r_cnt += 1.0;
            }
#pragma unroll
            for (int j = 0; j < kVcBatch; ++j) { mx[j] = nx[j]; my[j] = ny[j]; }
        }
        plane_color[pi] = (ff < 1.0f) ? __int_as_float(0x7fc00000) : __fdiv_rn(fc, ff);    // :441-445
    }
    vc_flush_stats(e_sum, r_cnt, stats);
}

// ---- vignette step: bilinear scatter of the normal equations  TT += w*cP*cP,  CT += w*cI*cP
__global__ void __launch_bounds__(256) vc_vignette_kernel(const float* __restrict__ images, const float* __restrict__ p2x,
                                                          const float* __restrict__ p2y, size_t total, int gwgh, int wI, size_t npx,
                                                          const float* __restrict__ plane_color, const float* __restrict__ vignette,
                                                          float* __restrict__ tt, float* __restrict__ ct, double oth2, int integer_abs,
                                                          double* __restrict__ stats) {
    double e_sum = 0.0, r_cnt = 0.0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const float x = __ldg(p2x + idx);
        if (isnan(x)) continue;                                                            // :469
        const float y = __ldg(p2y + idx);
        const size_t img = idx / static_cast<size_t>(gwgh);
        const int pi = static_cast<int>(idx - img * gwgh);
        const float cP = __ldg(plane_color + pi);
        if (isnan(cP)) continue;                                                           // :477
        const Taps t = vc_taps(x, y, wI);
        const float cI = vc_blend(images + img * npx, t, wI);
        if (isnan(cI)) continue;                                                           // :478
        const float fac = vc_blend(vignette, t, wI);
        const float d = __fsub_rn(cI, __fmul_rn(cP, fac));
        const double residual = static_cast<double>(__fmul_rn(d, d));
        if (vc_outlier(residual, oth2, integer_abs)) { e_sum += oth2; r_cnt += 1.0; continue; }              // :481-486
        float* a = tt + t.base;
        float* b = ct + t.base;
        atomicAdd(a, __fmul_rn(__fmul_rn(t.w00, cP), cP));
        atomicAdd(a + 1, __fmul_rn(__fmul_rn(t.w10, cP), cP));
        atomicAdd(a + wI, __fmul_rn(__fmul_rn(t.w01, cP), cP));
        atomicAdd(a + 1 + wI, __fmul_rn(__fmul_rn(t.w11, cP), cP));
        atomicAdd(b, __fmul_rn(__fmul_rn(t.w00, cI), cP));
        atomicAdd(b + 1, __fmul_rn(__fmul_rn(t.w10, cI), cP));
        atomicAdd(b + wI, __fmul_rn(__fmul_rn(t.w01, cI), cP));
        atomicAdd(b + 1 + wI, __fmul_rn(__fmul_rn(t.w11, cI), cP));
        if (isnan(fac)) continue;
        e_sum += residual;
        r_cnt += 1.0;
    }
    vc_flush_stats(e_sum, r_cnt, stats);
}

// vignette = CT/TT where TT >= 1, NaN elsewhere; running maximum of the finite factors (:507-517)
__global__ void __launch_bounds__(256) vc_divide_kernel(const float* __restrict__ tt, const float* __restrict__ ct, size_t npx,
                                                        float* __restrict__ vignette, int* __restrict__ max_bits) {
    float m = 0.0f;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx; i += stride) {
        float v = __int_as_float(0x7fc00000);
        if (!(tt[i] < 1.0f)) {
            v = __fdiv_rn(ct[i], tt[i]);
            if (v > m) m = v;
        }
        vignette[i] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // m >= 0 and not NaN, so the integer order of the bit patterns is the float order
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(max_bits, __float_as_int(m));
}
// normalise to vignette max. factor 1 (:521-523)
__global__ void __launch_bounds__(256) vc_normalise_kernel(float* __restrict__ vignette, size_t npx, const int* __restrict__ max_bits) {
    const float m = __int_as_float(*max_bits);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx; i += stride) vignette[i] = __fdiv_rn(vignette[i], m);
}

// one round of the NaN-aware 3x3 mean (:547-563), src -> dst
__global__ void __launch_bounds__(256) vc_smooth_kernel(const float* __restrict__ src, float* __restrict__ dst, int wI, int hI) {
    const size_t npx = static_cast<size_t>(wI) * hI;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx; i += stride) {
        const int y = static_cast<int>(i / wI), x = static_cast<int>(i - static_cast<size_t>(y) * wI);
        float sum = 0.0f, num = 0.0f;
        auto take = [&](bool inside, ptrdiff_t off) {
            if (!inside) return;
            const float v = src[static_cast<ptrdiff_t>(i) + off];
            if (!isnan(v)) { sum = __fadd_rn(sum, v); num = __fadd_rn(num, 1.0f); }
        };
        const bool r = x < wI - 1, l = x > 0, d = y < hI - 1, u = y > 0;
        take(r && d, 1 + wI);
        take(r, 1);
        take(r && u, 1 - wI);
        take(d, wI);
        take(true, 0);
        take(u, -wI);
        take(d && l, -1 + wI);
        take(l, -1);
        take(u && l, -1 - wI);
        dst[i] = (num > 0.0f) ? __fdiv_rn(sum, num) : src[i];
    }
}

unsigned vc_blocks(size_t work) {
    size_t b = (work + 255) / 256;
    if (b > 148u * 32u) b = 148u * 32u;
    return static_cast<unsigned>(b < 1 ? 1 : b);
}

struct Problem { const float *images, *p2x, *p2y; int n, gwgh, wI, hI; };

int check_problem(const char* who, mdc_ctx* c, const Problem& p, const void* a, const void* b) {
    if (!c || !p.images || !p.p2x || !p.p2y || !a || !b || p.n < 0 || p.gwgh < 1 || p.wI < 2 || p.hI < 2) {
        mdc_set_error("%s: bad argument", who);
        return MDC_ERR_INVALID_ARG;
    }
    return MDC_OK;
}

// scratch layout (floats): [0, A) first accumulator, [A, 2A) second; then 2 doubles of stats and the max word
struct Scratch {
    float *acc0 = nullptr, *acc1 = nullptr;
    double* stats = nullptr;
    int* max_bits = nullptr;
    void* base = nullptr;
};
int alloc_scratch(Scratch* s, size_t acc_elems) {
    const size_t bytes = 2 * acc_elems * sizeof(float) + 64;
    VC_CHECK(cudaMalloc(&s->base, bytes));
    s->stats = static_cast<double*>(s->base);
    s->max_bits = reinterpret_cast<int*>(s->stats + 2);
    s->acc0 = reinterpret_cast<float*>(static_cast<char*>(s->base) + 64);
    s->acc1 = s->acc0 + acc_elems;
    return MDC_OK;
}

int plane_step(mdc_ctx* c, const Problem& p, const float* d_vignette, float* d_plane_color, double oth2, int integer_abs, double* d_stats, cudaStream_t s) {
    VC_CHECK(cudaMemsetAsync(d_stats, 0, 2 * sizeof(double), s));
    vc_plane_kernel<<<vc_blocks(static_cast<size_t>(p.gwgh)), 256, 0, s>>>(p.images, p.p2x, p.p2y, p.n, p.gwgh, p.wI, static_cast<size_t>(p.wI) * p.hI,
                                                                           d_vignette, d_plane_color, oth2, integer_abs, d_stats);
    VC_CHECK(cudaGetLastError());
    mdc_ctx_add_launches(c, 1);
    return MDC_OK;
}

int vignette_step(mdc_ctx* c, const Problem& p, const float* d_plane_color, float* d_vignette, double oth2, int integer_abs, const Scratch& sc, cudaStream_t s) {
    const size_t npx = static_cast<size_t>(p.wI) * p.hI, total = static_cast<size_t>(p.n) * p.gwgh;
    VC_CHECK(cudaMemsetAsync(sc.base, 0, 64 + 2 * npx * sizeof(float), s));
    if (total) {
        vc_vignette_kernel<<<vc_blocks(total), 256, 0, s>>>(p.images, p.p2x, p.p2y, total, p.gwgh, p.wI, npx, d_plane_color, d_vignette, sc.acc0,
                                                            sc.acc1, oth2, integer_abs, sc.stats);
        VC_CHECK(cudaGetLastError());
    }
    vc_divide_kernel<<<vc_blocks(npx), 256, 0, s>>>(sc.acc0, sc.acc1, npx, d_vignette, sc.max_bits);
    vc_normalise_kernel<<<vc_blocks(npx), 256, 0, s>>>(d_vignette, npx, sc.max_bits);
    VC_CHECK(cudaGetLastError());
    mdc_ctx_add_launches(c, 3);
    return MDC_OK;
}

int smooth(mdc_ctx* c, const float* d_vignette, int wI, int hI, int iterations, float* d_out, float* d_tmp, cudaStream_t s) {
    const size_t npx = static_cast<size_t>(wI) * hI;
    VC_CHECK(cudaMemcpyAsync(d_out, d_vignette, npx * sizeof(float), cudaMemcpyDeviceToDevice, s));
    for (int it = 0; it < iterations; ++it) {
        VC_CHECK(cudaMemcpyAsync(d_tmp, d_out, npx * sizeof(float), cudaMemcpyDeviceToDevice, s));
        vc_smooth_kernel<<<vc_blocks(npx), 256, 0, s>>>(d_tmp, d_out, wI, hI);
        VC_CHECK(cudaGetLastError());
        mdc_ctx_add_launches(c, 1);
    }
    return MDC_OK;
}

}  // namespace

extern "C" int mdc_vc_plane_step(mdc_ctx* c, const float* d_images, const float* d_p2x, const float* d_p2y, int n, int gw, int gh, int wI, int hI,
                                 const float* d_vignette, float* d_plane_color, double outlier_th2, int integer_abs, double stats_host[2]) {
    const Problem p{d_images, d_p2x, d_p2y, n, gw * gh, wI, hI};
    int rc = check_problem("mdc_vc_plane_step", c, p, d_vignette, d_plane_color);
    if (rc != MDC_OK) return rc;
    VC_CHECK(cudaSetDevice(mdc_ctx_device_ordinal(c)));
    cudaStream_t s = static_cast<cudaStream_t>(mdc_ctx_stream_handle(c));
    double* d_stats = nullptr;
    VC_CHECK(cudaMalloc(&d_stats, 2 * sizeof(double)));
    rc = plane_step(c, p, d_vignette, d_plane_color, outlier_th2, integer_abs, d_stats, s);
    double st[2] = {0, 0};
    if (rc == MDC_OK && cudaMemcpyAsync(st, d_stats, sizeof st, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = MDC_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == MDC_OK) { mdc_set_error("mdc_vc_plane_step: %s", cudaGetErrorString(cudaGetLastError())); rc = MDC_ERR_CUDA; }
    cudaFree(d_stats);
    if (stats_host) { stats_host[0] = st[0]; stats_host[1] = st[1]; }
    return rc;
}

extern "C" int mdc_vc_vignette_step(mdc_ctx* c, const float* d_images, const float* d_p2x, const float* d_p2y, int n, int gw, int gh, int wI, int hI,
                                    const float* d_plane_color, float* d_vignette, double outlier_th2, int integer_abs, double stats_host[2]) {
    const Problem p{d_images, d_p2x, d_p2y, n, gw * gh, wI, hI};
    int rc = check_problem("mdc_vc_vignette_step", c, p, d_plane_color, d_vignette);
    if (rc != MDC_OK) return rc;
    VC_CHECK(cudaSetDevice(mdc_ctx_device_ordinal(c)));
    cudaStream_t s = static_cast<cudaStream_t>(mdc_ctx_stream_handle(c));
    Scratch sc;
    if ((rc = alloc_scratch(&sc, static_cast<size_t>(wI) * hI)) != MDC_OK) return rc;
    rc = vignette_step(c, p, d_plane_color, d_vignette, outlier_th2, integer_abs, sc, s);
    double st[2] = {0, 0};
    if (rc == MDC_OK && cudaMemcpyAsync(st, sc.stats, sizeof st, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = MDC_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == MDC_OK) { mdc_set_error("mdc_vc_vignette_step: %s", cudaGetErrorString(cudaGetLastError())); rc = MDC_ERR_CUDA; }
    cudaFree(sc.base);
    if (stats_host) { stats_host[0] = st[0]; stats_host[1] = st[1]; }
    return rc;
}

extern "C" int mdc_vc_smooth(mdc_ctx* c, const float* d_vignette, int wI, int hI, int iterations, float* d_out) {
    if (!c || !d_vignette || !d_out || wI < 1 || hI < 1 || iterations < 0) { mdc_set_error("mdc_vc_smooth: bad argument"); return MDC_ERR_INVALID_ARG; }
    VC_CHECK(cudaSetDevice(mdc_ctx_device_ordinal(c)));
    cudaStream_t s = static_cast<cudaStream_t>(mdc_ctx_stream_handle(c));
    float* tmp = nullptr;
    VC_CHECK(cudaMalloc(&tmp, static_cast<size_t>(wI) * hI * sizeof(float)));
    int rc = smooth(c, d_vignette, wI, hI, iterations, d_out, tmp, s);
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == MDC_OK) { mdc_set_error("mdc_vc_smooth: %s", cudaGetErrorString(cudaGetLastError())); rc = MDC_ERR_CUDA; }
    cudaFree(tmp);
    return rc;
}

// The reference's loop (:395-585): per iteration plane step, vignette step (+ normalisation), smoothed copy for output.
// log_host, if given, receives [max_iterations][4] = {E_plane, R_plane, E_vignette, R_vignette}.
extern "C" int mdc_vignette_calib(mdc_ctx* c, const float* d_images, const float* d_p2x, const float* d_p2y, int n, int gw, int gh, int wI, int hI,
                                  int max_iterations, int outlier_th, int integer_abs, float* d_plane_color, float* d_vignette, float* d_smoothed,
                                  double* log_host) {
    const Problem p{d_images, d_p2x, d_p2y, n, gw * gh, wI, hI};
    int rc = check_problem("mdc_vignette_calib", c, p, d_plane_color, d_vignette);
    if (rc != MDC_OK) return rc;
    if (max_iterations < 0) { mdc_set_error("mdc_vignette_calib: bad argument"); return MDC_ERR_INVALID_ARG; }
    VC_CHECK(cudaSetDevice(mdc_ctx_device_ordinal(c)));
    cudaStream_t s = static_cast<cudaStream_t>(mdc_ctx_stream_handle(c));
    const size_t npx = static_cast<size_t>(wI) * hI;
    Scratch sc;
    if ((rc = alloc_scratch(&sc, npx)) != MDC_OK) return rc;
    double* d_pstats = nullptr;
    if (cudaMalloc(&d_pstats, 2 * sizeof(double)) != cudaSuccess) { cudaFree(sc.base); mdc_set_error("mdc_vignette_calib: out of device memory"); return MDC_ERR_CUDA; }
    for (int it = 0; it < max_iterations && rc == MDC_OK; ++it) {
        double oth2 = static_cast<double>(outlier_th) * outlier_th;           // :397-398 (int arithmetic in the reference)
        if (it < max_iterations / 2) oth2 = 10000.0 * 10000.0;
        double ps[2] = {0, 0}, vs[2] = {0, 0};
        rc = plane_step(c, p, d_vignette, d_plane_color, oth2, integer_abs, d_pstats, s);
        if (rc == MDC_OK && cudaMemcpyAsync(ps, d_pstats, sizeof ps, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = MDC_ERR_CUDA;
        if (rc == MDC_OK) rc = vignette_step(c, p, d_plane_color, d_vignette, oth2, integer_abs, sc, s);
        if (rc == MDC_OK && cudaMemcpyAsync(vs, sc.stats, sizeof vs, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = MDC_ERR_CUDA;
        if (rc == MDC_OK && cudaStreamSynchronize(s) != cudaSuccess) { mdc_set_error("mdc_vignette_calib: %s", cudaGetErrorString(cudaGetLastError())); rc = MDC_ERR_CUDA; }
        if (rc != MDC_OK) break;
        printf("%f residual terms => %f\n", ps[1], sqrtf(static_cast<float>(ps[0] / ps[1])));      // :448
        printf("%f residual terms => %f\n", vs[1], sqrtf(static_cast<float>(vs[0] / vs[1])));      // :519
        if (log_host) { log_host[4 * it] = ps[0]; log_host[4 * it + 1] = ps[1]; log_host[4 * it + 2] = vs[0]; log_host[4 * it + 3] = vs[1]; }
    }
    if (rc == MDC_OK && d_smoothed) {
        rc = smooth(c, d_vignette, wI, hI, 4, d_smoothed, sc.acc0, s);
        if (cudaStreamSynchronize(s) != cudaSuccess && rc == MDC_OK) { mdc_set_error("mdc_vignette_calib: %s", cudaGetErrorString(cudaGetLastError())); rc = MDC_ERR_CUDA; }
    }
    cudaFree(d_pstats);
    cudaFree(sc.base);
    return rc;
}
