/* yolo_b200.h -- C ABI of the B200-native YOLOv3 inference path (libyolo_b200.so).
 *
 * The reference (ydixon/yolo_v3) is pure Python; it has no FFI for this path.  The entry points
 * below are what a ctypes binding for the path binds (yolo_v3_b200/_lib.py is that binding; the
 * stub a reference maintainer would add is in INTEGRATION.md).  Each one names the reference
 * interface it stands in for.  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - every function returns YB_OK (0) or a negative yb_status; yb_last_error() has the text;
 *     nothing throws across the ABI;
 *   - "dev" pointers are CUDA device pointers owned by the caller; the library owns packed
 *     weights, TMA descriptors and the activation arena (grown lazily per (B,H,W), freed in
 *     yb_destroy).  Nothing is allocated on the steady-state path;
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*, so this
 *     header needs no CUDA include) and is asynchronous unless stated otherwise;
 *   - one yb_ctx per (device, host thread); calls on one ctx are serialised by the caller;
 *   - tensors: images NCHW fp32 in [0,1] exactly as the reference feeds them (test.py:32);
 *     detections [B, N, 5+C] fp32, N = 3*(H/32*W/32 + H/16*W/16 + H/8*W/8), rows ordered
 *     stride-32 head, stride-16 head, stride-8 head (= torch.cat((det1,det2,det3),1), test.py:36).
 */
#ifndef YOLO_B200_H_
#define YOLO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yb_ctx yb_ctx;

typedef enum {
    YB_OK = 0,
    YB_E_ARG = -1,        /* bad argument / shape (H or W not a multiple of 32, B <= 0, ...) */
    YB_E_KEY = -2,        /* unknown state_dict key or wrong element count */
    YB_E_STATE = -3,      /* call order (forward before finalize, ...) */
    YB_E_CUDA = -4,       /* CUDA runtime / driver error (text in yb_last_error) */
    YB_E_NCCL = -5,       /* NCCL error or libnccl not loadable */
    YB_E_CAP = -6,        /* output capacity too small */
    YB_E_NOMEM = -7,
    YB_E_UNSUPPORTED = -8 /* e.g. fp16 tensor-core mode on a non-sm_100 device */
} yb_status;

/* precision_mode for yb_finalize */
#define YB_MODE_FP32 0    /* fp32 activations + fp32 CUDA-core implicit GEMM (debugging aid: slow, fixed summation order) */
#define YB_MODE_FP16 1    /* fp16 activations/weights, fp32 accumulate in TMEM on tcgen05 tensor cores (perf path) */
#define YB_MODE_FP32_TC 2 /* fp32-grade parity with the reference ON the tensor cores: every activation and weight is an
                           * fp16 pair hi + lo (22 mantissa bits), three tcgen05 kind::f16 partial products per k-step
                           * (hi*hi + hi*lo + lo*hi) accumulated in fp32 in TMEM, fp32 epilogue.  What precision='fp32' runs. */

/* ---- lifetime -------------------------------------------------------------------------------- */

/* Replaces YoloNet.__init__(img_dim, anchors, numClass) (darknet.py:167-195).  anchors: 9 (w,h)
 * pairs in pixels, may be NULL for the reference default. */
int yb_create(yb_ctx** out, int device, int num_classes, const float anchors[18]);
void yb_destroy(yb_ctx* ctx);
/* ctx may be NULL: returns the text of the last error raised by a call that had no ctx. */
const char* yb_last_error(const yb_ctx* ctx);
int yb_num_tensors(const yb_ctx* ctx);
/* i-th state_dict key in registration order and its element count (438 keys for 80 classes). */
const char* yb_tensor_key(const yb_ctx* ctx, int i, size_t* numel);

/* ---- weights --------------------------------------------------------------------------------- */

/* Replaces nn.Module.load_state_dict(...) entry by entry (darknet.py:240-243, train.py:26).  Keys
 * are the reference's own ("feature.mlist.2.conv1.conv.weight", "pre_det3.mlist.6.bias", ...);
 * values fp32, conv weights [Cout,Cin,k,k].  "*.num_batches_tracked" is accepted and ignored. */
int yb_set_tensor(yb_ctx* ctx, const char* key, const float* data, size_t n, int on_host);
int yb_get_tensor(const yb_ctx* ctx, const char* key, float* host_out, size_t n);

/* Replaces WeightManager.loadWeight (darknet.py:249-303): `host` is the float stream that follows
 * the 5-int32 header of a darknet .weights file; per BN conv: bn.bias, bn.weight, running_mean,
 * running_var, conv.weight; per plain conv: bias, weight; modules in cfg order.  backbone_only=1
 * is Darknet.loadWeight (darknet.py:102-104, darknet53.conv.74).  *consumed <- floats used. */
int yb_load_darknet_blob(yb_ctx* ctx, const float* host, size_t nfloats, int backbone_only, size_t* consumed);
/* Inverse of the above (the reference leaves saveWeight(format='darknet') unimplemented,
 * darknet.py:237-238).  Pass host_out=NULL to query the float count. */
int yb_save_darknet_blob(const yb_ctx* ctx, float* host_out, size_t capacity, int backbone_only, size_t* written);

/* Folds BN (eval, eps 1e-5) into per-channel fp32 scale/bias, packs the weights for the chosen
 * mode and uploads them.  Must be called after the weights change and before forward.  Synchronous. */
int yb_finalize(yb_ctx* ctx, int precision_mode);

/* Element type of the images handed to yb_forward / yb_forward_logits / yb_backbone / yb_detect (x_nchw is then read as
 * that type).  The reference's callers hold fp32 tensors (imgs.cuda(), test.py:32): YB_INPUT_F32, the default.  With
 * YB_INPUT_F16 (YB_MODE_FP16 only) the stem reads fp16 [B,3,H,W] and produces the SAME bits -- the fp32 path rounds every
 * pixel to fp16 (round to nearest even, what Tensor.half() does on the host) before the tensor core sees it -- for half
 * the host-to-device bytes. */
#define YB_INPUT_F32 0
#define YB_INPUT_F16 1
int yb_set_input_dtype(yb_ctx* ctx, int dtype);

/* CUDA-graph replay of the convolution launches that follow the stem (they read and write plan-owned buffers only): mode 0
 * never, 1 always, 2 (default) automatically for launch-bound shapes (B*H*W <= 2^20 pixels, e.g. one 416x416 image -- 80
 * dependent kernels of a few microseconds each).  The graph of a (B,H,W) plan is captured on its second call; capture
 * failure falls back to stream launches silently.  No reference counterpart (PyTorch eager launches every kernel). */
int yb_set_graph_mode(yb_ctx* ctx, int mode);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* Replaces YoloNet.forward(x, target=None) + torch.cat((det1,det2,det3),1) (darknet.py:198-231,
 * test.py:35-36).  x_nchw: dev [B,3,H,W] fp32.  det_cat: dev [B,N,5+C] fp32; det1/det2/det3 are
 * the row ranges [0,n1), [n1,n1+n2), [n1+n2,N) of it. */
int yb_forward(yb_ctx* ctx, const float* x_nchw, int B, int H, int W, float* det_cat, void* stream);

/* Same convolution stack, but returns the three raw head maps (what pre_detN.mlist[6] outputs,
 * darknet.py:118) as dev NCHW fp32 [B,3*(5+C),H/32,W/32], [.., H/16, W/16], [.., H/8, W/8].
 * Parity / debugging aid for the conv kernels. */
int yb_forward_logits(yb_ctx* ctx, const float* x_nchw, int B, int H, int W,
                      float* logits32, float* logits16, float* logits8, void* stream);

/* Replaces Darknet.forward (darknet.py:83-88): backbone only.  feat_nchw: dev [B,1024,H/32,W/32] fp32. */
int yb_backbone(yb_ctx* ctx, const float* x_nchw, int B, int H, int W, float* feat_nchw, void* stream);

/* Replaces the three YoloLayer.forward(x, img_dim, None) calls (yololayer.py:31-59,97-105) on raw
 * head maps given in the reference layout (dev NCHW fp32), writing the concatenated det tensor. */
int yb_decode(yb_ctx* ctx, const float* logits32, const float* logits16, const float* logits8,
              int B, int H, int W, float* det_cat, void* stream);

/* Replaces utils.postprocessing(detections, num_classes, obj_conf_thr, nms_thr, is_eval, use_nms)
 * (utils.py:226-258 incl. get_nms_detections :148-202, iou_vectorized :98-119, get_raw_detections
 * :204-224, boundingbox.bbox_cxcywh_to_x1y1x2y2 :25-29).  det_cat: dev [B,N,5+C] (not modified).
 * Outputs (dev): rows7 [B,cap,7] = x1,y1,x2,y2,obj,score,cls, class-ascending / score-descending
 * per image (candidate order when use_nms=0); counts [B] = rows produced per image (if > cap the
 * image was truncated to cap rows: caller should retry with a larger cap); src_index [B,cap] =
 * flat box index of each row (may be NULL); cand_counts [B] = boxes (box,class pairs when is_eval)
 * that passed the threshold (may be NULL; all zero <=> the reference returns []).
 * Tie-break is fixed: score descending, then candidate order ascending. */
int yb_postprocess(yb_ctx* ctx, const float* det_cat, int B, int N, float obj_conf_thr, float nms_thr,
                   int is_eval, int use_nms, float* rows7, int* counts, int* src_index, int* cand_counts,
                   int cap, void* stream);

/* Replaces the notebook's own inline post-process (yolo_detect.ipynb cell 35 with its helpers in cells 30 and 33:
 * torch_unique, iou_vectorized, reduce_row_by_column, nms), which differs from utils.postprocessing: a box is a
 * candidate when its OBJECTNESS > obj_conf_thr (>= 0), its class is the arg-max of the raw class probabilities,
 * x2 = (cx - w/2) + w, boxes of a class are ordered by objectness (descending, ties in candidate order) and the
 * pair-list reduction of cells 33 is greedy NMS in that order.  Same outputs as yb_postprocess; rows7 column 5 holds
 * the winning class probability. */
int yb_postprocess_notebook(yb_ctx* ctx, const float* det_cat, int B, int N, float obj_conf_thr, float nms_thr,
                            float* rows7, int* counts, int* src_index, int* cand_counts, int cap, void* stream);

/* forward + postprocess with no intermediate host round trip (what test.py:35-36 does per batch). */
int yb_detect(yb_ctx* ctx, const float* x_nchw, int B, int H, int W, float obj_conf_thr, float nms_thr,
              int is_eval, int use_nms, float* rows7, int* counts, int* src_index, int* cand_counts,
              int cap, void* stream);

/* Replaces boundingbox.correct_yolo_boxes(bboxes, org_w, org_h, img_w, img_h, is_letterbox)
 * (boundingbox.py:139-149 = letterbox_reverse :95-116 or rescale_bbox :119-137, then x1y1x2y2 -> xywh,
 * :10-15), the step every caller applies right after postprocessing (test.py:41, evaluate.py:187), for a
 * whole batch on the device.  boxes: dev, B x cap rows of `row_stride` floats whose first four are
 * x1,y1,x2,y2 in network-input pixels (row_stride 7 for rows7, 4 for bare boxes); counts: dev [B] valid rows
 * per image, or NULL for "all cap rows"; org_wh_host: HOST int[B][2] = original (w,h) of every image;
 * out_xywh: dev [B,cap,4] = x,y,w,h in original-image pixels.  Rows whose four coordinates sum to 0 are
 * passed through unchanged, as in the reference. */
int yb_correct_boxes(yb_ctx* ctx, const float* boxes, int row_stride, const int* counts, int B, int cap,
                     const int* org_wh_host, int img_w, int img_h, int is_letterbox, float* out_xywh, void* stream);

/* Replaces utils.load_image(path, 'letterbox', dim) after the file decode (utils.py:60-72) for a batch of images of
 * different sizes: letterbox_image (utils.py:44-57 = letterbox_transforms :34-42, cv2.resize(..., INTER_CUBIC) to the
 * box, paste on a grey-128 canvas) followed by torch.from_numpy(img).float().permute(2,0,1) / 255.  The resize
 * arithmetic is OpenCV's portable 8-bit path, bit for bit (csrc/preprocess.cu).
 * imgs_dev: HOST array of B device pointers to uint8 RGB HWC images (row pitch w*3); hw_host: HOST int[B][2] = (h, w)
 * of every image; dim_w, dim_h: the reference's `dim` = (outer_w, outer_h), from which every image's box is derived;
 * canvas_h, canvas_w: the canvas.  The reference allocates np.full(dim + (3,)), i.e. canvas_h = dim[0], canvas_w =
 * dim[1] -- the same thing for the square sizes it uses; a box that does not fit the canvas makes numpy raise and this
 * call return YB_E_ARG.  (dim = (w, h) of the image itself with canvas (h, w) is the plain float()/255 + HWC->CHW of
 * load_image(mode=None).)  offset_rule 0: box offsets as utils.letterbox_transforms (w//2 - box_w//2); 1: as the
 * dataset transform IaaLetterbox (transforms.py:144-212: same cubic resize, offsets (w - box_w)//2, canvas [dim_h,dim_w]).
 * out_nchw: dev [B,3,canvas_h,canvas_w] fp32; canvas_hwc: dev uint8 [B,canvas_h,canvas_w,3], the canvas
 * letterbox_image itself returns (either output may be NULL, not both); trans_host: HOST float[B][5] = box_w, box_h,
 * box_x, box_y, ratio (the reference's `trans`), may be NULL. */
int yb_letterbox(yb_ctx* ctx, const uint8_t* const* imgs_dev, const int* hw_host, int B, int dim_w, int dim_h,
                 int canvas_h, int canvas_w, int offset_rule, float* out_nchw, uint8_t* canvas_hwc, float* trans_host,
                 void* stream);

/* Replaces utils.load_image(path, 'resize', dim) after the file decode (utils.py:68-71): cv2.resize(img, dim) (default
 * INTER_LINEAR; OpenCV's 8-bit fixed-point path, bit for bit) then float()/255 and HWC->CHW, for a batch.  dim_w, dim_h:
 * the reference's `dim` = cv2's dsize (w, h); out_nchw: dev [B,3,dim_h,dim_w] fp32; out_hwc: dev uint8 [B,dim_h,dim_w,3]
 * (either may be NULL, not both); other arguments as yb_letterbox. */
int yb_resize(yb_ctx* ctx, const uint8_t* const* imgs_dev, const int* hw_host, int B, int dim_w, int dim_h,
              float* out_nchw, uint8_t* out_hwc, void* stream);

/* ---- multi-GPU (absent in the reference; batch sharding) ------------------------- */

/* 128-byte NCCL unique id, created on rank 0 and shipped to the other ranks by the host plumbing
 * (torch.distributed / gloo object broadcast). */
int yb_comm_unique_id(uint8_t id_out[128]);
int yb_comm_init(yb_ctx* ctx, const uint8_t id[128], int rank, int world);
/* ncclBroadcast of the finalized (packed) weight blob from `root`; every rank must have called
 * yb_finalize with the same mode (so the buffers exist) -- non-root contents are overwritten. */
int yb_bcast_weights(yb_ctx* ctx, int root, void* stream);
/* ncclAllGather of fixed-capacity detection rows + counts: all_rows [world*B_local,cap,7],
 * all_counts [world*B_local], rank-major. */
int yb_allgather_dets(yb_ctx* ctx, const float* rows7, const int* counts, int B_local, int cap,
                      float* all_rows, int* all_counts, void* stream);

/* ---- introspection for bench.py / tests -------------------------------------------------------- */

/* Number of kernels this library launched on behalf of ctx since creation (kernels replayed by a captured graph included). */
long long yb_launch_count(const yb_ctx* ctx);
/* Number of forward calls whose post-stem convolution launches were replayed from a captured CUDA graph (yb_set_graph_mode). */
long long yb_graph_replays(const yb_ctx* ctx);
/* Device time (ms) of the convolution stack / decode / post-process sections of the last
 * yb_forward/yb_detect call with profiling enabled.  yb_set_profiling level: 0 off, 1 = one CUDA event at
 * each section boundary (does not disturb the back-to-back launches inside the convolution stack),
 * 2 = additionally an event after every convolution (per-layer times, serialises the launches), 3 = (yb_detect
 * only) section events recorded without any synchronisation, so that calls still run back to back; yb_get_section_ms
 * then synchronises once and returns the average over the (at most 16) calls made since the previous query. */
int yb_set_profiling(yb_ctx* ctx, int enabled);
int yb_get_section_ms(yb_ctx* ctx, float* conv_ms, float* decode_ms, float* post_ms);
/* Per-layer timing of the last profiled call: ms[i] for the i-th convolution (75), returns count. */
int yb_get_layer_ms(yb_ctx* ctx, float* ms, int capacity);
/* Watchdog words of the tensor-core kernel (host-mapped, survive a device trap): [0]=1 if a
 * pipeline wait timed out, [1]=block, [2]=role (0 producer,1 mma,2 epilogue), [3]=barrier, [4]=parity.
 * Returns the number of words copied. */
int yb_debug_words(const yb_ctx* ctx, int* out, int n);
/* Standalone single-convolution entry used by the kernel unit tests: runs layer `layer_index`
 * (0..74, cfg order) in the finalized mode on a dev NHWC tensor of the mode's element type
 * (fp32 or fp16), in [B,H,W,Cin] -> out [B,Ho,Wo,Cout(+pad)], with optional residual (same shape as out). */
int yb_run_layer(yb_ctx* ctx, int layer_index, const void* in_nhwc, int B, int H, int W,
                 const void* residual_nhwc, void* out_nhwc, void* stream);
/* Unit-test entry of the fused first kernel (YB_MODE_FP16): stem + the first stride-2 convolution (reference
 * darknet.py:66-69) on a dev NCHW image of the current input element type -> out [B, H/2, W/2, 64] fp16 NHWC.
 * H even, W a multiple of 4 (fp32 images) / 8 (fp16); YB_E_UNSUPPORTED for widths whose last column tile would be
 * less than 80 % full (the network path then runs the two layers as separate kernels). */
int yb_run_stem_block(yb_ctx* ctx, const void* x_nchw, int B, int H, int W, void* out_nhwc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLO_B200_H_ */
