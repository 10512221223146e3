// Internal declarations shared by the translation units of libyolo_b200.so.
// Public C ABI: include/yolo_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/yolo_b200.h"

namespace yb {

constexpr float kBnEps = 1e-5f;   // nn.BatchNorm2d default (reference darknet.py:39)
constexpr float kLeaky = 0.1f;    // reference darknet.py:41

// One convolution of the network, in darknet-cfg order (reference darknet.py:72-79,107-118,153-157).
struct Layer {
    std::string key;      // state_dict prefix
    int cin = 0, cout = 0, ks = 1, stride = 1;
    bool bn = true;
    int cout_pad = 0;     // cout rounded up to 16 (heads: 255 -> 256)
    // host master copies (fp32, reference layouts)
    std::vector<float> w;                     // [cout][cin][ks][ks]
    std::vector<float> g, b, mean, var;       // bn.weight, bn.bias, running_mean, running_var  (bn)
    std::vector<float> bias;                  // conv bias (plain head conv)
    // device, filled by finalize()
    float* d_scale = nullptr;                 // [cout_pad]  gamma/sqrt(var+eps)  (1 for heads)
    float* d_bias = nullptr;                  // [cout_pad]  beta-mean*scale      (conv bias for heads)
    float* d_w32 = nullptr;                   // [ks*ks][cin][cout_pad] fp32      (YB_MODE_FP32)
    __half* d_w16 = nullptr;                  // [cout_pad][ks*ks*cin] fp16, K-major (YB_MODE_FP16);
                                              // YB_MODE_FP32_TC: [cout_pad][ks*ks][cin block][wh | wl][bke], rows pre-scaled by 2^s
};

// cudaFuncSetAttribute applies to the CURRENT device only, and a process may hold contexts on several devices
// (one yb_ctx per device and host thread): run `set` once per device, serialised, before the first launch there.
class PerDeviceOnce {
public:
    template <class F>
    cudaError_t run(F&& set) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const unsigned long long bit = 1ull << (dev & 63);
        std::lock_guard<std::mutex> lock(mu_);
        if (done_ & bit) return cudaSuccess;
        e = set();
        if (e == cudaSuccess) done_ |= bit;
        return e;
    }

private:
    std::mutex mu_;
    unsigned long long done_ = 0;
};

// Tuning / A-B overrides are read from the environment only in builds with -DYB_EXPERIMENTS (make EXPERIMENTS=1, what
// tools/layer_bench.py --sweep needs); the release library never calls getenv and has one code path per layer shape.
inline const char* tune_env(const char* name) {
#ifdef YB_EXPERIMENTS
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

// NHWC activation view: `p` already includes the channel offset of a concat slice.
struct TView {
    void* p = nullptr;
    int B = 0, H = 0, W = 0, C = 0;
    long ld = 0;          // elements between consecutive pixels
    long lo = 0;          // YB_MODE_FP32_TC: channel offset of the lo half (0 in the other modes)
};

// Arguments common to both convolution kernels.
struct ConvArgs {
    const void* in; long in_ld;
    void* out; long out_ld;
    const void* res; long res_ld;     // nullable residual, added after the activation
    const float* scale; const float* bias;
    int B, H, W, Cin;
    int Ho, Wo, Cout;                 // Cout = channels actually stored (cout_pad for fp32 head maps)
    int ks, stride, pad;
    int leaky;                        // LeakyReLU(0.1) after scale/bias
    int upsample;                     // nearest x2: every output pixel is written to a 2x2 block of a [B,2Ho,2Wo] tensor
    int out_f32;                      // output element is fp32 regardless of the activation type (head maps)
    // YB_MODE_FP32_TC ("split"): activations are fp16 pairs hi + lo sharing the pixel pitch; *_lo = channel offset of the
    // lo half relative to the hi pointer (in / out / res point at the hi half)
    int split;
    long in_lo, out_lo, res_lo;
};

// conv_simt.cu
template <typename T>
cudaError_t launch_conv_simt(const ConvArgs& a, const float* w32, int cout_pad, cudaStream_t s);
template <typename T>
cudaError_t launch_stem(const float* x_nchw, T* out_nhwc, const float* w32, const float* scale, const float* bias,
                        int B, int H, int W, cudaStream_t s);
template <typename T>
cudaError_t launch_nhwc_to_nchw_f32(const T* in, long in_ld, int C, int B, int HW, float* out, cudaStream_t s);
cudaError_t launch_f32_to_f16(const float* in, __half* out, size_t n, cudaStream_t s);
// YB_MODE_FP32_TC ("split" activations: fp16 hi | lo per pixel)
cudaError_t launch_stem_split(const float* x_nchw, __half* out_nhwc64, const float* w32, const float* scale, const float* bias,
                              int B, int H, int W, cudaStream_t s);
cudaError_t launch_split_to_nchw_f32(const __half* in, long in_ld, long lo, int C, int B, int HW, float* out, cudaStream_t s);
cudaError_t launch_f32_to_split(const float* in, __half* out, size_t M, int C, cudaStream_t s);
cudaError_t launch_split_to_f32(const __half* in, float* out, size_t M, int C, cudaStream_t s);
// nearest x2 of an fp16 tensor: in [B,H,W] pixels of C values (halves = 2: + C lo values in_lo later) -> the 2x2 blocks of out [B,2H,2W]
cudaError_t launch_upsample2x(const __half* in, long in_ld, long in_lo, __half* out, long out_ld, long out_lo, int C,
                              int B, int H, int W, int halves, cudaStream_t s);

// conv_tc.cu  (tcgen05 + TMA implicit GEMM)
struct TcPlan {
    CUtensorMap tmA, tmB, tmOut, tmRes;
    int epi_staged = 0, ring = 0, sub_bytes = 128, cs = 0, n_sub = 0, b_resident = 0;
    int swz = 128;        // 128: 64-channel k-blocks, 64: 32-channel k-blocks (Cin == 32)
    int BN = 0, n_tiles = 0, m_tiles = 0, stages = 0, tmem_cols = 0;
    int num_kblocks = 0, cin_blocks = 0, kps = 1, cta2 = 0, cout_pad = 0, tab_bytes = 0;
    int srel = 0;
    int epi_split = 0;    // 32-column sub-tiles: the two halves of the epilogue warps take alternate sub-tiles
    int split = 0;        // YB_MODE_FP32_TC: hi/lo operands, three k sections per (tap, channel block) (conv_tc.cu)
    int chunk_pairs = 0, n_chunks = 0;   // split mode: unit pairs per TMEM chunk, chunks per tile
    int grid = 0;
    size_t smem = 0;
    long M = 0;
};
// Builds TMA descriptors + tile configuration for one layer invocation. Returns "" or an error text.
std::string tc_make_plan(TcPlan& p, const ConvArgs& a, const __half* w16, int cout_pad, int K, int num_sms);
cudaError_t tc_launch(const TcPlan& p, const ConvArgs& a, int* dbg, cudaStream_t s);
bool tc_supported(const ConvArgs& a);
// stem_halo.cu  (the Cin = 3 stem from a halo patch: planar TMA load, pixel-major fp16 conversion, two taps per tcgen05.mma
// through the leading-dimension offset of a non-swizzled descriptor).  Needs W % 4 == 0 (fp32 images) / W % 8 == 0 (fp16).
struct StemHaloPlan {
    CUtensorMap tmOut;
    int tiles_x = 0, tiles_y = 0, total_tiles = 0, grid = 0;
    size_t smem = 0;
};
bool stem_halo_supported(int W, int in_f16);
std::string stem_halo_make_plan(StemHaloPlan& p, __half* out, long out_ld, int B, int H, int W, int num_sms);
// sb_host: the stem's scale[32] | bias[32] on the HOST -- they travel as kernel parameters (constant-bank operands)
cudaError_t stem_halo_launch(const StemHaloPlan& p, const void* x, int in_f16, int B, int H, int W, const __half* w16,
                             const float* sb_host, int* dbg, cudaStream_t s);

// stem_halo_split.cu  (the halo stem of YB_MODE_FP32_TC: hi/lo operand pairs, output [B,H,W,32 hi | 32 lo]; fp32 images, W % 4 == 0)
bool stem_split_supported(int W);
void stem_split_host_params(const float* w27x32, const float* scale, const float* bias, float* sc_eff, float* bi_out, int* shift);
std::string stem_split_make_plan(StemHaloPlan& p, __half* out, long out_ld, int B, int H, int W, int num_sms);
cudaError_t stem_split_launch(const StemHaloPlan& p, const float* x, int B, int H, int W, const float* w32, const float* sc_eff_host,
                              const float* bias_host, const int* shift_host, int* dbg, cudaStream_t s);

// stem_block.cu  (stem + the first stride-2 convolution in one kernel: the 32-channel stem output never leaves the SM)
struct StemBlockPlan {
    CUtensorMap tmW1, tmOut;
    int tiles_x = 0, tiles_y = 0, total_tiles = 0, grid = 0;
    size_t smem = 0;
};
bool stem_block_supported(int H, int W, int in_f16);
// w1: layer 1's fp16 weights [64][9*32] (k = tap*32 + c); out: layer 1's output [B, H/2, W/2, out_ld >= 64]
std::string stem_block_make_plan(StemBlockPlan& p, const __half* w1, __half* out, long out_ld, int B, int H, int W, int num_sms);
cudaError_t stem_block_launch(const StemBlockPlan& p, const void* x, int in_f16, int B, int H, int W, const __half* w0,
                              const float* sb0_host, const float* scale1, const float* bias1, int* dbg, cudaStream_t s);

// conv_halo.cu  (3x3 stride-1 layers with Cin = 32 / 64 from a halo tile: every input pixel staged once)
struct HaloPlan {
    CUtensorMap tmIn, tmB, tmOut, tmRes;
    int swz = 128, stride = 1, cout_pad = 0, tiles_x = 0, tiles_y = 0, n_tiles = 0, total_tiles = 0;
    int ring = 0, tab_bytes = 0, slot_bytes = 0, stages = 0, grid = 0;
    int pair = 0;         // CTA pairs: M = 256 (two spatial tiles) x N = 128 per cta_group::2 UMMA (Cout = 128, Cin = 64, stride 1)
    size_t smem = 0;
};
bool halo_supported(const ConvArgs& a);
std::string halo_make_plan(HaloPlan& p, const ConvArgs& a, const __half* w16, int cout_pad, int K, int num_sms);
cudaError_t halo_launch(const HaloPlan& p, const ConvArgs& a, int* dbg, cudaStream_t s);

// decode.cu
struct DecodeScale {
    const float* logits;  // NHWC [B,h,w,ld] (internal) or NCHW [B,3*(5+C),h,w] (API)
    long ld;              // NHWC pixel pitch in floats (ignored for NCHW)
    int h, w;
    int row_off;          // first row of this scale in det_cat
    float stride;         // pixels per cell
    float aw[3], ah[3];   // anchors / stride, fp32 (reference yololayer.py:37-38)
};
cudaError_t launch_decode(const DecodeScale sc[3], int nchw, int B, int attrs, int n_total, float* det, cudaStream_t s);
// One warp per grid cell over the NHWC head maps (pixel pitch sc[].ld, a multiple of 4 floats).  mode bit 0: write
// det_cat (standalone decode), bit 1: score the rows into rowcount / rowcand (front end of the non-eval post-process).
// `list`: scratch of at least 1 + (cells of the batch) ints for the live-cell list of the two-kernel scoring form (may be
// NULL: single-kernel form); *extra_launches <- kernels launched beyond the first.
cudaError_t launch_decode_cells(const DecodeScale sc[3], int B, int attrs, int n_total, int mode, float* det, float thr,
                                int* rowcount, float* rowcand, int* list, int* extra_launches, int num_sms, cudaStream_t s);

// postprocess.cu
struct PostBuffers {
    int B = 0, N = 0, cand_cap = 0, sort_cap = 0, C = 0;
    int* rowcount = nullptr;    // [B][N]
    int* rowoff = nullptr;      // [B][N]
    int* cand_total = nullptr;  // [B]
    float* rowcand = nullptr;   // [B][N][8]   (non-eval: candidate of the row, written only when it passes)
    float* cand = nullptr;      // [B][cand_cap][8]  x1,y1,x2,y2,obj,score,cls,boxidx(bits)
    unsigned long long* keys = nullptr;  // [B][sort_cap]
    float4* sbox = nullptr;     // [B][cand_cap] boxes in sorted order
    unsigned char* keep = nullptr;  // [B][cand_cap]
    int* seg = nullptr;         // [B][C][2] start,end
    size_t bytes = 0;
};
struct PostArgs {
    const float* det; int B, N, C;      // det may be NULL when pre_scored

    float conf_thr, nms_thr;
    int is_eval, use_nms;
    float* rows7; int* counts; int* src_index; int* cand_counts; int cap;
    int pre_scored = 0;                 // rowcount / rowcand were already filled by launch_decode_cells (non-eval only)
    int variant = 0;                    // 1: the notebook's inline post-process (yolo_detect.ipynb cell 35), non-eval only
};
cudaError_t launch_postprocess(const PostArgs& a, PostBuffers& buf, long long* launches, cudaStream_t s);
cudaError_t launch_correct_boxes(const float* boxes, int row_stride, const int* counts, int B, int cap, const float* params_dev,
                                 float* out, cudaStream_t s);

// preprocess.cu
struct LbImage {
    const unsigned char* src;   // dev, uint8 RGB HWC, row pitch sw*3
    int sh, sw;                 // source size
    int box_w, box_h, box_x, box_y;
    double scale_x, scale_y;    // 1.0 / ((double)box / src), as cv::resize computes it
    int interp;                 // 0: INTER_CUBIC (letterbox), 1: INTER_LINEAR (plain resize)
};
cudaError_t launch_letterbox(const LbImage* imgs_dev, int B, int canvas_h, int canvas_w, float* out, unsigned char* canvas,
                             cudaStream_t s);

}  // namespace yb
