/* y5obb.h — C ABI of the B200-native yolov5_obb hot path (liby5obb.so).
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host.
 * All entry points are asynchronous on `stream` (a cudaStream_t passed as void*) unless stated, never
 * call cudaDeviceSynchronize, and return 0 on success or a Y5OBB_E* code (the Python shims raise
 * RuntimeError on != 0, as the reference's AT_ASSERTM / AT_CUDA_CHECK do).
 *
 * Each declaration cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef Y5OBB_H_
#define Y5OBB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y5OBB_OK 0
#define Y5OBB_EINVAL 1     /* bad argument (null pointer, negative size, unsupported shape) */
#define Y5OBB_EWORKSPACE 2 /* workspace too small: call the matching *_workspace_bytes */
#define Y5OBB_ECUDA 3      /* a CUDA call failed; y5obb_last_cuda_error() has the cudaError_t */
#define Y5OBB_EARCH 4      /* device is not sm_100 */

/* ---- library ------------------------------------------------------------------------------- */
int y5obb_abi_version(void);            /* bumps when a signature changes */
int y5obb_last_cuda_error(void);        /* cudaError_t of the last failing CUDA call on this thread */
const char* y5obb_build_info(void);     /* "sm_100a nvcc <ver> <date>" */

/* ---- rotated NMS ---------------------------------------------------------------------------
 * Replaces  utils/nms_rotated/src/nms_rotated_ext.cpp:25-39  (nms_rotated(dets, scores, thr))
 *           utils/nms_rotated/src/nms_rotated_cuda.cu:71-134 (sort, 64x64 IoU bit-matrix, D2H, host scan)
 *           utils/nms_rotated/nms_rotated_wrapper.py:26-42   (min(w,h) < 0.001 pre-filter)
 * Everything — sort, IoU tiles, greedy scan, compaction — runs on the device; nothing is copied to
 * the host.
 *
 * flags */
#define Y5OBB_NMS_STRICT_GT 1   /* suppress if IoU >  thr (reference CUDA, nms_rotated_cuda.cu:60); else */
                                /* suppress if IoU >= thr (reference CPU, nms_rotated_cpu.cpp:55)         */
#define Y5OBB_NMS_DROP_SMALL 2  /* boxes with min(w,h) < 0.001 never enter NMS (nms_rotated_wrapper.py:32) */

/* Workspace needed for n_total boxes over n_images images with at most max_per_image boxes in any one
 * image (pass n_total if unknown). */
size_t y5obb_nms_workspace_bytes(int64_t n_total, int64_t n_images, int64_t max_per_image);

/* One NMS problem.  dets5 [n,5] = (cx, cy, w, h, theta_rad) fp32 row-major, scores [n] fp32.
 * keep_out [n] int64 receives indices into dets5 in descending-score order (ties: lower index first);
 * n_keep_out [1] int64 receives the count. */
int y5obb_nms_rotated_f32(const float* dets5, const float* scores, int64_t n, float iou_thr, int flags,
                          int64_t* keep_out, int64_t* n_keep_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* n_images independent problems in one launch sequence.  image_ids [n_total] int32 in [0, n_images)
 * (any order).  Image b's kept indices (into dets5) are written in descending-score order to
 * keep_out[seg_off_out[b] .. seg_off_out[b] + n_keep_out[b]); seg_off_out has n_images + 1 entries.
 * max_keep > 0 stops each image after that many keeps (non_max_suppression_obb's max_det,
 * utils/general.py:854-855); 0 = unlimited. */
int y5obb_nms_rotated_batched_f32(const float* dets5, const float* scores, const int32_t* image_ids,
                                  int64_t n_total, int64_t n_images, int64_t max_per_image,
                                  float iou_thr, int flags, int64_t max_keep,
                                  int64_t* keep_out, int64_t* n_keep_out, int64_t* seg_off_out,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Pairwise rotated IoU of n independent pairs (a[i], b[i]), each 5 floats.  Device restatement of
 * utils/nms_rotated/src/box_iou_rotated_utils.h:334-360 (single_box_iou_rotated<float>). */
int y5obb_rbox_iou_pairs_f32(const float* a5, const float* b5, float* iou_out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* Y5OBB_H_ */
