// Native NCCL broadcast of the calibration tables (see include/mdc_b200_nccl.h).
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>

#include "mdc_b200.h"
#include "mdc_b200_nccl.h"

namespace {
// error text goes through stderr here: mdc_set_error lives in the other library and is not part of its ABI
#define NC_CHECK(expr, code)                                                                      \
    do {                                                                                          \
        if (!(expr)) { fprintf(stderr, "mdc_ctx_create_broadcast: %s failed (%s:%d)\n", #expr, __FILE__, __LINE__); return code; } \
    } while (0)
}  // namespace

extern "C" int mdc_ctx_create_broadcast(void* nccl_comm, int rank, int root, int device, const mdc_fov* fov, const mdc_photo* photo,
                                        mdc_ctx** out) {
    if (!nccl_comm || !out) return MDC_ERR_INVALID_ARG;
    *out = nullptr;
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    NC_CHECK(cudaSetDevice(device) == cudaSuccess, MDC_ERR_CUDA);
    cudaStream_t s;
    NC_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess, MDC_ERR_CUDA);

    // geometry + which tables exist
    int meta_h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rank == root) {
        if (fov) mdc_fov_dims(fov, &meta_h[0], &meta_h[1], &meta_h[2], &meta_h[3]);
        meta_h[4] = fov && mdc_fov_is_valid(fov);
        meta_h[5] = photo && mdc_photo_valid_gamma(photo);
        meta_h[6] = photo && mdc_photo_valid_vignette(photo);
    }
    int* meta_d = nullptr;
    NC_CHECK(cudaMalloc(&meta_d, sizeof meta_h) == cudaSuccess, MDC_ERR_CUDA);
    NC_CHECK(cudaMemcpyAsync(meta_d, meta_h, sizeof meta_h, cudaMemcpyHostToDevice, s) == cudaSuccess, MDC_ERR_CUDA);
    NC_CHECK(ncclBroadcast(meta_d, meta_d, 8, ncclInt, root, comm, s) == ncclSuccess, MDC_ERR_CUDA);
    NC_CHECK(cudaMemcpyAsync(meta_h, meta_d, sizeof meta_h, cudaMemcpyDeviceToHost, s) == cudaSuccess, MDC_ERR_CUDA);
    NC_CHECK(cudaStreamSynchronize(s) == cudaSuccess, MDC_ERR_CUDA);
    cudaFree(meta_d);
    const int in_w = meta_h[0], in_h = meta_h[1], out_w = meta_h[2], out_h = meta_h[3];
    const bool has_fov = meta_h[4] != 0, has_g = meta_h[5] != 0, has_v = meta_h[6] != 0;

    auto bcast = [&](const float* host_src, size_t count, float** dev) -> bool {
        if (cudaMalloc(dev, count * sizeof(float)) != cudaSuccess) return false;
        if (rank == root && cudaMemcpyAsync(*dev, host_src, count * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) return false;
        return ncclBroadcast(*dev, *dev, count, ncclFloat, root, comm, s) == ncclSuccess;
    };
    float *rx = nullptr, *ry = nullptr, *g = nullptr, *v = nullptr;
    const size_t n_out = static_cast<size_t>(out_w) * out_h, n_in = static_cast<size_t>(in_w) * in_h;
    if (has_fov) {
        NC_CHECK(bcast(rank == root ? mdc_fov_remap_x(fov) : nullptr, n_out, &rx), MDC_ERR_CUDA);
        NC_CHECK(bcast(rank == root ? mdc_fov_remap_y(fov) : nullptr, n_out, &ry), MDC_ERR_CUDA);
    }
    if (has_g) NC_CHECK(bcast(rank == root ? mdc_photo_ginv(const_cast<mdc_photo*>(photo)) : nullptr, 256, &g), MDC_ERR_CUDA);
    if (has_v) NC_CHECK(bcast(rank == root ? mdc_photo_vignette_map_inv(photo) : nullptr, n_in, &v), MDC_ERR_CUDA);
    NC_CHECK(cudaStreamSynchronize(s) == cudaSuccess, MDC_ERR_CUDA);
    cudaStreamDestroy(s);

    int rc = mdc_ctx_create_from_device_tables(device, in_w, in_h, out_w, out_h, rx, ry, g, v, out);
    if (rc != MDC_OK) { cudaFree(rx); cudaFree(ry); cudaFree(g); cudaFree(v); return rc; }
    return mdc_ctx_take_table_ownership(*out);   // the context now frees the four buffers
}
