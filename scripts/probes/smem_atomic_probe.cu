// Probe: what does one histogram update cost on the shared-memory pipe?  Every thread walks a pseudo-random byte stream (one byte
// per update, like the G-step) and updates a lane-sliced 256-bin histogram in shared memory with
//   mode 0: one native ATOMS.ADD.32            row stride 128 B, slot = lane        (1 limb)
//   mode 1: three native ATOMS.ADD.32          three planes of mode 0               (3 limbs of 17 bits = fixed-point sums)
//   mode 2: three native ATOMS.ADD.32          row stride 64 B, slot = lane & 15    (half the memory, lanes l / l+16 share banks)
//   mode 3: atomicAdd(u64) = LDS.64 + CAS loop row stride 128 B, slot = lane & 15   (what the fp64 / 64-bit integer version does)
//   mode 4: atomicAdd(double)                  same layout as 3
// usage: smem_atomic_probe [threads=896] [ctas_per_sm=1] [iters=4096]; prints clocks per warp-update per SM for every mode.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

template <int kMode>
__global__ void probe(int iters, unsigned long long* sink, long long* clocks) {
    extern __shared__ __align__(16) uint8_t smem[];
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t x = 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t b = (x >> (8 * j)) & 0xffu;
            const double prod = __longlong_as_double(0x4338000000000000ll | (static_cast<long long>(x) << 7) | i);      // a "converted" sample
            const uint32_t lo = static_cast<uint32_t>(__double2loint(prod)), hi = static_cast<uint32_t>(__double2hiint(prod));
            if (kMode == 0) {
                atomicAdd(reinterpret_cast<unsigned*>(smem + b * 128 + lane * 4), lo & 0x1ffffu);
            } else if (kMode == 1) {
                uint8_t* p = smem + b * 128 + lane * 4;
                atomicAdd(reinterpret_cast<unsigned*>(p), lo & 0x1ffffu);
                atomicAdd(reinterpret_cast<unsigned*>(p + 32768), __funnelshift_r(lo, hi, 17) & 0x1ffffu);
                atomicAdd(reinterpret_cast<int*>(p + 65536), static_cast<int>(hi - 0x43380000u) >> 2);
            } else if (kMode == 2) {
                uint8_t* p = smem + b * 64 + (lane & 15) * 4;
                atomicAdd(reinterpret_cast<unsigned*>(p), lo & 0x1ffffu);
                atomicAdd(reinterpret_cast<unsigned*>(p + 16384), __funnelshift_r(lo, hi, 17) & 0x1ffffu);
                atomicAdd(reinterpret_cast<int*>(p + 32768), static_cast<int>(hi - 0x43380000u) >> 2);
            } else if (kMode == 3) {
                atomicAdd(reinterpret_cast<unsigned long long*>(smem + b * 128 + (lane & 15) * 8), static_cast<unsigned long long>(__double_as_longlong(prod)) - 0x4338000000000000ull);
            } else {
                atomicAdd(reinterpret_cast<double*>(smem + b * 128 + (lane & 15) * 8), prod);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < 128 * 1024 / 8; i += blockDim.x) s += reinterpret_cast<unsigned long long*>(smem)[i];
    if (s == 0x1234567ull) sink[0] = s;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int kMode>
static void run(const char* name, int threads, int per_sm, int iters, int sms) {
    const int smem = per_sm == 1 ? 128 * 1024 : 100 * 1024;
    cudaFuncSetAttribute(probe<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    unsigned long long* sink; long long* clocks;
    const int grid = sms * per_sm;
    cudaMalloc(&sink, 8); cudaMalloc(&clocks, grid * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<kMode><<<grid, threads, smem>>>(16, sink, clocks);
    cudaEventRecord(a);
    probe<kMode><<<grid, threads, smem>>>(iters, sink, clocks);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long c0 = 0; cudaMemcpy(&c0, clocks, 8, cudaMemcpyDeviceToHost);
    const double warp_updates_per_sm = 4.0 * iters * (threads / 32) * per_sm;
    printf("%-34s %4d thr x %d CTA/SM: %7.3f ms, %6.2f clk per warp-update per SM (CTA 0: %lld clk), %7.1f G updates/s\n", name, threads, per_sm, ms,
           c0 / (4.0 * iters * (threads / 32) * per_sm), c0, warp_updates_per_sm * 32 * sms / (ms * 1e-3) / 1e9);
    cudaFree(sink); cudaFree(clocks);
}

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 896, per_sm = argc > 2 ? atoi(argv[2]) : 1, iters = argc > 3 ? atoi(argv[3]) : 4096;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("1 x ATOMS.ADD.32, 32 slots", threads, per_sm, iters, sms);
    run<1>("3 x ATOMS.ADD.32, 32 slots", threads, per_sm, iters, sms);
    run<2>("3 x ATOMS.ADD.32, 16 slots", threads, per_sm, iters, sms);
    run<3>("atomicAdd(u64) CAS loop, 16 slots", threads, per_sm, iters, sms);
    run<4>("atomicAdd(f64) CAS loop, 16 slots", threads, per_sm, iters, sms);
    return 0;
}
