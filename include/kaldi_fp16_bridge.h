/* kaldi_fp16_bridge.h -- drop-in C ABI for the reference's memory / transfer bridge.
 *
 * Names, argument order and return conventions (0 / -1 or NULL, message via bridge_last_error)
 * are those of /root/reference/cpp/include/bridge.h; the citation after each prototype is the
 * declaration it replaces.  Callers: internal/gpu/{bridge,tensor,ops}.go through cgo.
 */
#ifndef KALDI_FP16_B200_BRIDGE_H
#define KALDI_FP16_B200_BRIDGE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *bridge_last_error(void); /* bridge.h:14 */
void bridge_clear_error(void);       /* bridge.h:15 */

/* cudaSetDevice; additionally refuses a device that is not sm_100 (no fallback) -- bridge.h:18 */
int bridge_gpu_init(int device_id);
int bridge_gpu_get_free_memory(size_t *free_bytes, size_t *total_bytes); /* bridge.h:19 */
int bridge_gpu_sync(void);                                               /* bridge.h:20 */

void *bridge_gpu_malloc(size_t bytes); /* bridge.h:23 */
void bridge_gpu_free(void *ptr);       /* bridge.h:24 */
void *bridge_host_alloc(size_t bytes); /* pinned -- bridge.h:25 */
void bridge_host_free(void *ptr);      /* bridge.h:26 */

/* blocking copies, `count` ELEMENTS -- bridge.h:29-32 */
int bridge_transfer_fp16(void *dst_device, const uint16_t *src_host, size_t count);
int bridge_read_fp16(uint16_t *dst_host, const void *src_device, size_t count);
int bridge_transfer_int32(void *dst_device, const int32_t *src_host, size_t count);
int bridge_transfer_float32(void *dst_device, const float *src_host, size_t count);
/* addition: read FP32 back (gradients / master weights in tests) */
int bridge_read_float32(float *dst_host, const void *src_device, size_t count);

/* packed minibatch: one allocation, 256-byte aligned sections -- bridge.h:33-50 */
typedef struct {
    void *d_features;    /* fp16 [total_frames * feat_dim] */
    void *d_ivectors;    /* fp16 [batch_size * ivec_dim]   */
    void *d_csr_row_ptr; /* int32 [num_states + 1]         */
    void *d_csr_col_idx; /* int32 [num_arcs]               */
    void *d_csr_labels;  /* int32 [num_arcs]               */
    void *d_csr_weights; /* float [num_arcs]               */
    void *d_buffer;
    size_t total_bytes;
    size_t features_bytes;
    size_t ivectors_bytes;
    size_t csr_rowptr_bytes;
    size_t csr_colidx_bytes;
    size_t csr_labels_bytes;
    size_t csr_weights_bytes;
} GPUBatchPtrs;

int bridge_batch_alloc(int total_frames, int feat_dim, int batch_size, int ivec_dim, int num_states,
                       int num_arcs, GPUBatchPtrs *out);                                  /* bridge.h:52 */
int bridge_batch_transfer(const GPUBatchPtrs *ptrs, const void *host_buf, size_t total_bytes); /* bridge.h:54 */
void bridge_batch_free(GPUBatchPtrs *ptrs);                                               /* bridge.h:55 */
void bridge_gpu_memset(void *ptr, int value, size_t bytes);                               /* bridge.h:56 */
int bridge_fp16_to_fp32_gpu(float *dst_device, const void *src_device, size_t count);     /* bridge.h:57 */
int bridge_fp32_to_fp16_gpu(void *dst_device, const float *src_device, size_t count);     /* bridge.h:58 */

#ifdef __cplusplus
}
#endif
#endif /* KALDI_FP16_B200_BRIDGE_H */
