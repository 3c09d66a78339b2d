// Probe: throughput of the scatter-add pattern of the vignetteCalib vignette step (8 atomics per sample: a 2x2 footprint in two
// accumulator images of 1280x1024, footprints of neighbouring threads next to each other) with
//   mode 0: fp32 RED.ADD        (what the step used: order-dependent sums)
//   mode 1: 64-bit integer RED  (fixed point: order-independent sums)
//   mode 2: 32-bit integer RED  (for scale)
// usage: global_atomic_probe [samples_in_millions=256]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

template <int kMode>
__global__ void __launch_bounds__(256) scatter(void* acc0, void* acc1, size_t total, int w, int h, int gw) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        // plane point (px, py) of image `img` lands near a smooth function of (px, py, img): neighbours in the grid are neighbours in the image
        const size_t img = idx / (static_cast<size_t>(gw) * gw);
        const int pi = static_cast<int>(idx - img * gw * gw), py = pi / gw, px = pi - py * gw;
        const int x = 100 + (px * 9) / 10 + static_cast<int>(img % 37), y = 60 + (py * 8) / 10 + static_cast<int>(img % 23);
        if (x + 1 >= w || y + 1 >= h) continue;
        const size_t base = static_cast<size_t>(y) * w + x;
        const float v = 1.0f + 0.001f * static_cast<float>(pi & 1023);
        if (kMode == 0) {
            float *a = static_cast<float*>(acc0) + base, *b = static_cast<float*>(acc1) + base;
            atomicAdd(a, v); atomicAdd(a + 1, v); atomicAdd(a + w, v); atomicAdd(a + w + 1, v);
            atomicAdd(b, v); atomicAdd(b + 1, v); atomicAdd(b + w, v); atomicAdd(b + w + 1, v);
        } else if (kMode == 1) {
            unsigned long long *a = static_cast<unsigned long long*>(acc0) + base, *b = static_cast<unsigned long long*>(acc1) + base;
            const unsigned long long q = static_cast<unsigned long long>(__double_as_longlong(static_cast<double>(v) * 1048576.0 + 6755399441055744.0) - 0x4338000000000000ll);
            atomicAdd(a, q); atomicAdd(a + 1, q); atomicAdd(a + w, q); atomicAdd(a + w + 1, q);
            atomicAdd(b, q); atomicAdd(b + 1, q); atomicAdd(b + w, q); atomicAdd(b + w + 1, q);
        } else {
            unsigned *a = static_cast<unsigned*>(acc0) + base, *b = static_cast<unsigned*>(acc1) + base;
            const unsigned q = static_cast<unsigned>(v * 1024.0f);
            atomicAdd(a, q); atomicAdd(a + 1, q); atomicAdd(a + w, q); atomicAdd(a + w + 1, q);
            atomicAdd(b, q); atomicAdd(b + 1, q); atomicAdd(b + w, q); atomicAdd(b + w + 1, q);
        }
    }
}

template <int kMode>
static void run(const char* name, size_t total) {
    const int w = 1280, h = 1024, gw = 1000;
    void *a0, *a1;
    cudaMalloc(&a0, static_cast<size_t>(w) * h * 8); cudaMalloc(&a1, static_cast<size_t>(w) * h * 8);
    cudaMemset(a0, 0, static_cast<size_t>(w) * h * 8); cudaMemset(a1, 0, static_cast<size_t>(w) * h * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    scatter<kMode><<<148 * 32, 256>>>(a0, a1, total / 8, w, h, gw);
    cudaEventRecord(e0);
    scatter<kMode><<<148 * 32, 256>>>(a0, a1, total, w, h, gw);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-22s %7.3f ms for %zu M samples: %6.1f G samples/s = %7.1f G atomics/s\n", name, ms, total >> 20, total / (ms * 1e-3) / 1e9, 8.0 * total / (ms * 1e-3) / 1e9);
    cudaFree(a0); cudaFree(a1);
}

int main(int argc, char** argv) {
    const size_t total = static_cast<size_t>(argc > 1 ? atoi(argv[1]) : 256) << 20;
    run<0>("fp32 RED", total);
    run<1>("64-bit integer RED", total);
    run<2>("32-bit integer RED", total);
    return 0;
}
