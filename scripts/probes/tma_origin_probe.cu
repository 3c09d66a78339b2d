// Probe: does a tiled u8 TMA load (cp.async.bulk.tensor.2d) accept a box origin whose x is not a multiple of 16 bytes?
// usage: tma_origin_probe <x0> <box_w> [elem_bytes=1]     x0 / box_w in ELEMENTS of the map's data type (u8, u16 or u32 view of
// the same bytes); prints OK / MISMATCH, or dies with the CUDA error.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int x0, int y0, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
    const uint32_t buf_a = static_cast<uint32_t>(__cvta_generic_to_shared(buf));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.proxy.async.shared::cta;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes));
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(buf_a), "l"(&map), "r"(x0), "r"(y0), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
    const int es_b = argc > 3 ? atoi(argv[3]) : 1;          // element size of the tensor map's view
    const int x0 = (argc > 1 ? atoi(argv[1]) : 0) * es_b, bw = (argc > 2 ? atoi(argv[2]) : 32) * es_b, bh = 8, W = 256, H = 64, y0 = 3;   // bytes
    std::vector<uint8_t> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = static_cast<uint8_t>((i * 7 + (i >> 8) * 13) & 0xff);
    uint8_t *d, *o;
    cudaMalloc(&d, W * H); cudaMalloc(&o, bw * bh);
    cudaMemcpy(d, h.data(), W * H, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)(W / es_b), (cuuint64_t)H}, strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {(cuuint32_t)(bw / es_b), (cuuint32_t)bh}, es[2] = {1, 1};
    const CUtensorMapDataType dt = es_b == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : es_b == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUresult r = enc(&map, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("elem %d B, origin byte %d, box %d B: ", es_b, x0, bw);
    if (r != CUDA_SUCCESS) { printf("x0=%d bw=%d encode failed %d\n", x0, bw, (int)r); return 2; }
    probe<<<1, 128, bw * bh>>>(map, x0 / es_b, y0, bw * bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("x0=%d bw=%d CUDA error: %s\n", x0, bw, cudaGetErrorString(e)); return 3; }
    std::vector<uint8_t> g(bw * bh);
    cudaMemcpy(g.data(), o, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) bad += g[y * bw + x] != h[(y0 + y) * W + x0 + x];
    printf("x0=%d bw=%d %s (%d mismatches)\n", x0, bw, bad ? "MISMATCH" : "OK", bad);
    return bad ? 1 : 0;
}
