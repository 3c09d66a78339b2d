/*
 * mdc_b200_nccl.h — optional native multi-GPU init for C++ hosts (libmdc_b200_nccl.so, links libnccl).
 *
 * The reference is single-process / single-device; this is the B200-side equivalent of "construct the
 * DatasetReader's two calibration objects" when frames are sharded over several GPUs: rank
 * `root` has parsed the files and built the tables on its host (mdc_fov_create / mdc_photo_create), the four
 * tables travel once over NVLink with ncclBroadcast, and every rank gets a device context that owns its copy.
 * There is no collective after this call.  Python hosts use torch.distributed instead
 * (mono_dataset_code_b200/sharding.py); both paths end in mdc_ctx_create_from_device_tables.
 */
#ifndef MDC_B200_NCCL_H
#define MDC_B200_NCCL_H
#include "mdc_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* nccl_comm: an initialised ncclComm_t for this rank (passed as void* to keep nccl.h out of this header).
 * fov / photo: the host models on `root`, ignored (may be NULL) elsewhere.  device: this rank's CUDA device. */
int mdc_ctx_create_broadcast(void* nccl_comm, int rank, int root, int device, const mdc_fov* fov, const mdc_photo* photo,
                             mdc_ctx** out);

#ifdef __cplusplus
}
#endif
#endif
