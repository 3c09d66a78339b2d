// Probe: texture gather (tld4) on a u8 PITCH-2D (linear memory) texture object.
//   1. does it work at all (the CUDA guide documents tex2Dgather for CUDA arrays created with cudaArrayTextureGather only),
//   2. in which order do the four texels come back,
//   3. up to which texture height (maxTexture2DGather vs maxTexture2DLinear),
//   4. how many lanes per clock and SM the TEX pipe sustains for it, alone and next to LUT look-ups + stores on the LSU pipe.
// usage: tex_gather_probe            prints one line per check; exit code 0 if gather works with the expected order
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 3; } } while (0)

__device__ __forceinline__ void tld4_u8(unsigned long long tex, float u, float v, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& w) {
    asm volatile("tld4.r.2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6}];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(tex), "f"(u), "f"(v));
}
__host__ __device__ inline uint8_t pixel(int x, int y) { return static_cast<uint8_t>((x * 37 + y * 101 + ((x * y) >> 3)) & 0xff); }

__global__ void fill(uint8_t* img, int W, long long H) {
    const long long n = static_cast<long long>(W) * H;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        img[i] = pixel(static_cast<int>(i % W), static_cast<int>(i / W));
}
// one gather per thread at integer positions (xi, yi); out[4*i..] = x, y, z, w components
__global__ void gather_at(unsigned long long tex, const int* xs, const int* ys, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x, y, z, w;
    tld4_u8(tex, static_cast<float>(xs[i] + 1), static_cast<float>(ys[i] + 1), x, y, z, w);
    out[4 * i] = x; out[4 * i + 1] = y; out[4 * i + 2] = z; out[4 * i + 3] = w;
}
// throughput: every warp walks `iters` frames of a 4-row strip like K1 does: lane = x (step sx source pixels), 4 gathers per iteration;
// mode bit 1: + 16 lane-replicated LUT look-ups, bit 2: + 4 streaming stores
__global__ void __launch_bounds__(256, 3) rate(unsigned long long tex, int W, int H, int iters, float sx, float sy, int mode, float* out, size_t out_stride) {
    extern __shared__ float lut[];
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) lut[i] = static_cast<float>(i >> 5);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t lut_lane = static_cast<uint32_t>(__cvta_generic_to_shared(lut)) + 4u * lane;
    const int tiles_x = W / 128;
    const float u0 = 4.0f + (warp % tiles_x) * 100.0f + lane * sx;
    float v[4];
    for (int q = 0; q < 4; ++q) v[q] = 4.0f + ((warp / tiles_x) % 64) * 8.0f + q * sy;
    float acc = 0.f;
    float* o = out + (static_cast<size_t>(warp) * 4) * 32 + lane;
    for (int it = 0; it < iters; ++it) {
        uint32_t b[4][4];
        for (int q = 0; q < 4; ++q) tld4_u8(tex, floorf(u0) + 1.0f, floorf(v[q]) + 1.0f, b[q][0], b[q][1], b[q][2], b[q][3]);
        float px[4];
        for (int q = 0; q < 4; ++q) {
            float g = 0.f;
            for (int k = 0; k < 4; ++k) {
                float t;
                if (mode & 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(lut_lane + (b[q][k] << 7)));
                else t = __uint_as_float(0x3f800000u | b[q][k]);
                g += t;
            }
            px[q] = g;
            v[q] += static_cast<float>(H);
        }
        if (mode & 2) { for (int q = 0; q < 4; ++q) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(o + q * 32), "f"(px[q]) : "memory"); o += out_stride; }
        else acc += px[0] + px[1] + px[2] + px[3];
    }
    if (acc == 12345.678f) out[0] = acc;
}

static int make_tex(cudaTextureObject_t* t, uint8_t* ptr, int W, long long H) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = ptr; rd.res.pitch2D.desc = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = H; rd.res.pitch2D.pitchInBytes = W;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaError_t e = cudaCreateTextureObject(t, &rd, &td, nullptr);
    if (e != cudaSuccess) { printf("cudaCreateTextureObject(%d x %lld): %s\n", W, H, cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s: maxTexture2DGather %d x %d, maxTexture2DLinear %d x %d (pitch %d), textureAlignment %zu, texturePitchAlignment %zu, SMs %d\n",
           prop.name, prop.maxTexture2DGather[0], prop.maxTexture2DGather[1], prop.maxTexture2DLinear[0], prop.maxTexture2DLinear[1],
           prop.maxTexture2DLinear[2], prop.textureAlignment, prop.texturePitchAlignment, prop.multiProcessorCount);
    const int W = 1280;
    const long long Hmax = 65000;
    uint8_t* img;
    CK(cudaMalloc(&img, static_cast<size_t>(W) * Hmax));
    fill<<<1024, 256>>>(img, W, Hmax);
    CK(cudaDeviceSynchronize());
    int ok_order = 0;
    const long long heights[] = {1024, 24576, 32768, 49152, 65000};
    for (long long H : heights) {
        cudaTextureObject_t tex;
        if (make_tex(&tex, img, W, H)) continue;
        const int n = 4096;
        std::vector<int> xs(n), ys(n);
        uint32_t s = 99u + static_cast<uint32_t>(H);
        for (int i = 0; i < n; ++i) {
            s = s * 1664525u + 1013904223u; xs[i] = (s >> 8) % (W - 1);
            s = s * 1664525u + 1013904223u; ys[i] = static_cast<int>((s >> 4) % static_cast<uint32_t>(H - 1));
            if (i < 64) ys[i] = static_cast<int>(H - 2 - i);      // the last rows of the stack
        }
        int *dx, *dy; uint32_t* dout;
        CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&dy, n * 4)); CK(cudaMalloc(&dout, n * 16));
        CK(cudaMemcpy(dx, xs.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, ys.data(), n * 4, cudaMemcpyHostToDevice));
        gather_at<<<(n + 127) / 128, 128>>>(tex, dx, dy, n, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("height %lld: gather kernel failed: %s\n", H, cudaGetErrorString(e)); return 4; }
        std::vector<uint32_t> out(4 * n);
        CK(cudaMemcpy(out.data(), dout, n * 16, cudaMemcpyDeviceToHost));
        // expected (GL/D3D order): x = (xi, yi+1), y = (xi+1, yi+1), z = (xi+1, yi), w = (xi, yi)
        int bad = 0, first = -1;
        for (int i = 0; i < n; ++i) {
            const uint32_t ex[4] = {pixel(xs[i], ys[i] + 1), pixel(xs[i] + 1, ys[i] + 1), pixel(xs[i] + 1, ys[i]), pixel(xs[i], ys[i])};
            for (int k = 0; k < 4; ++k) if (out[4 * i + k] != ex[k]) { ++bad; if (first < 0) first = i; }
        }
        printf("height %lld: %s (%d of %d components differ from the x=(0,1) y=(1,1) z=(1,0) w=(0,0) order)\n", H, bad ? "MISMATCH" : "OK", bad, 4 * n);
        if (bad && first >= 0) {
            const int i = first;
            printf("  first bad sample at (%d, %d): got %u %u %u %u; texels (0,0)=%u (1,0)=%u (0,1)=%u (1,1)=%u\n", xs[i], ys[i], out[4 * i], out[4 * i + 1],
                   out[4 * i + 2], out[4 * i + 3], pixel(xs[i], ys[i]), pixel(xs[i] + 1, ys[i]), pixel(xs[i], ys[i] + 1), pixel(xs[i] + 1, ys[i] + 1));
        }
        if (!bad && H == 1024) ok_order = 1;
        cudaDestroyTextureObject(tex); cudaFree(dx); cudaFree(dy); cudaFree(dout);
    }
    // ---- throughput
    {
        const int H = 1024, frames = 24;
        cudaTextureObject_t tex;
        if (make_tex(&tex, img, W, static_cast<long long>(H) * frames) == 0) {
            const int grid = prop.multiProcessorCount * 3;
            const size_t warps = static_cast<size_t>(grid) * 8, out_stride = warps * 4 * 32;
            float* out;
            CK(cudaMalloc(&out, out_stride * frames * sizeof(float)));
            CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
            int clk_khz = 0;
            cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
            const float steps[][2] = {{2.9f, 1.6f}, {1.0f, 1.0f}, {0.5f, 0.5f}};
            for (auto& st : steps)
                for (int mode = 0; mode < 4; ++mode) {
                    cudaEvent_t a, b;
                    cudaEventCreate(&a); cudaEventCreate(&b);
                    rate<<<grid, 256, 32768>>>(tex, W, H, frames, st[0], st[1], mode, out, out_stride);
                    cudaEventRecord(a);
                    const int reps = 20;
                    for (int r = 0; r < reps; ++r) rate<<<grid, 256, 32768>>>(tex, W, H, frames, st[0], st[1], mode, out, out_stride);
                    cudaEventRecord(b);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); return 5; }
                    float ms = 0;
                    cudaEventElapsedTime(&ms, a, b);
                    const double lanes = static_cast<double>(warps) * 32 * 4 * frames * reps;        // gathers (one per lane)
                    const double per_clk_sm = lanes / (ms * 1e-3) / (clk_khz * 1e3) / prop.multiProcessorCount;
                    printf("rate: step %.1f x %.1f px/lane, +lut %d, +stores %d: %.3f ms / %d launches, %.2f G gathers/s = %.2f lanes/clk/SM at %d MHz (%.0f G output px/s equivalent)\n",
                           st[0], st[1], mode & 1, (mode >> 1) & 1, ms, reps, lanes / (ms * 1e-3) / 1e9, per_clk_sm, clk_khz / 1000, lanes / (ms * 1e-3) / 1e9);
                }
        }
    }
    return ok_order ? 0 : 1;
}
