// Native C++ multi-GPU init check (2 GPUs, one process, one thread per rank):
// rank 0 parses the calibration files; mdc_ctx_create_broadcast ships the tables over NCCL; rank 1 — which
// never saw the files — prepares frames with its received tables and dumps the result for comparison
// with the CPU oracle (tests/test_nccl_cpp.py).     nccl_bcast_check <dataset_dir/> <frames.bin> <n> <out.bin>
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mdc_b200.h"
#include "mdc_b200_nccl.h"

int main(int argc, char** argv) {
    if (argc < 5) { printf("usage: nccl_bcast_check <dataset_dir/> <frames.bin> <n> <out.bin>\n"); return 1; }
    const std::string dir = argv[1];
    const int n = atoi(argv[3]);
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 2) { printf("need 2 GPUs, have %d\n", ndev); return 77; }

    mdc_fov* fov = 0; mdc_photo* photo = 0;
    if (mdc_fov_create((dir + "camera.txt").c_str(), &fov) != MDC_OK) return 2;
    int iw, ih, ow, oh;
    mdc_fov_dims(fov, &iw, &ih, &ow, &oh);
    if (mdc_photo_create((dir + "pcalib.txt").c_str(), (dir + "vignette.png").c_str(), iw, ih, &photo) != MDC_OK) return 2;

    std::vector<unsigned char> frames((size_t)n * iw * ih);
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) return 3;
    fclose(f);

    ncclComm_t comms[2];
    int devs[2] = {0, 1};
    if (ncclCommInitAll(comms, 2, devs) != ncclSuccess) return 4;

    std::vector<float> out[2];
    int status[2] = {0, 0};
    auto worker = [&](int rank) {
        mdc_ctx* ctx = 0;
        // only rank 0 passes the host models
        status[rank] = mdc_ctx_create_broadcast(comms[rank], rank, 0, devs[rank], rank == 0 ? fov : 0, rank == 0 ? photo : 0, &ctx);
        if (status[rank] != MDC_OK) return;
        out[rank].resize((size_t)n * ow * oh);
        float* lv[1] = {out[rank].data()};
        status[rank] = mdc_prepare_batch_host(ctx, frames.data(), n, MDC_RECTIFY | MDC_REMOVE_GAMMA | MDC_REMOVE_VIGNETTE | MDC_NAN_OVEREXPOSED, lv, 1);
        mdc_ctx_destroy(ctx);
    };
    std::thread t0(worker, 0), t1(worker, 1);
    t0.join(); t1.join();
    ncclCommDestroy(comms[0]); ncclCommDestroy(comms[1]);
    if (status[0] != MDC_OK || status[1] != MDC_OK) { printf("status %d %d: %s\n", status[0], status[1], mdc_last_error()); return 5; }

    f = fopen(argv[4], "wb");
    fwrite(out[1].data(), sizeof(float), out[1].size(), f);      // the NON-root rank's result
    fclose(f);
    // both ranks must agree bit for bit
    for (size_t i = 0; i < out[0].size(); i++) {
        unsigned a, b;
        memcpy(&a, &out[0][i], 4); memcpy(&b, &out[1][i], 4);
        if (a != b && !(out[0][i] != out[0][i] && out[1][i] != out[1][i])) { printf("rank mismatch at %zu\n", i); return 6; }
    }
    printf("nccl_bcast_check done\n");
    return 0;
}
