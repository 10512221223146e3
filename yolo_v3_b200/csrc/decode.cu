// K2: anchor decode of the three head maps -- YoloLayer.forward with target=None
// (reference yololayer.py:31-59, 97-105), all three scales in one launch, written straight into the
// concatenated [B, N, 5+C] tensor the callers build with torch.cat (test.py:36).
//
//   channel = a*(5+C) + attr          (view at yololayer.py:42)
//   row     = (y*w + x)*3 + a         (permute at yololayer.py:104)
//   bx = (sigmoid(tx) + x) * stride   by likewise          (yololayer.py:57, :98)
//   bw = (exp(tw) * (anchor_w/stride)) * stride            (yololayer.py:59, :98)
//   conf, cls = sigmoid(t)
//
// Full-precision expf and IEEE division (no -use_fast_math); every product is a separate rounding
// as in the reference.  HBM-bound: reads 4 B and writes 4 B per element, both fully coalesced.
#include "yb_internal.h"

namespace yb {
namespace {

struct DecodeParams {
    DecodeScale sc[3];
    long cells_before[4];   // prefix sum of B*h*w per scale
    int B, attrs, n_total;
};

__device__ __forceinline__ float sigmoidf_rn(float t) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-t))); }

// One exponential and one IEEE division per element, no divergent branches: the per-attribute
// differences are folded into (use_exp, offset, mul).  (s + 0.0f) * 1.0f == s exactly, so conf/cls
// come out bit-identical to a plain sigmoid.
__device__ __forceinline__ float decode_one(const DecodeScale& s, int a, int attr, int x, int y, float t) {
    const bool wh = attr == 2 || attr == 3;
    const float e = expf(wh ? t : -t);
    const float sig = __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
    const float off = attr == 0 ? (float)x : (attr == 1 ? (float)y : 0.f);
    const float mul = attr < 2 ? s.stride : 1.0f;
    const float r_sig = __fmul_rn(__fadd_rn(sig, off), mul);
    const float anc = attr == 2 ? s.aw[a] : s.ah[a];
    const float r_wh = __fmul_rn(__fmul_rn(e, anc), s.stride);
    return wh ? r_wh : r_sig;
}

// NHWC (engine-internal) input: element (b, p, c) -> out[b][row_off*attrs + p*3*attrs + c]: a flat
// elementwise map.  A block owns kCellsPerBlock consecutive cells of one image and scale (so the cell ->
// (x, y) arithmetic is two scalar divisions per block and increments afterwards); thread t owns channel
// t (anchor/attr computed once) and walks the cells, four loads in flight at a time; consecutive threads
// touch consecutive addresses on both sides.
constexpr int kCellsPerBlock = 16;
constexpr int kDecodeUnroll = 4;

__global__ void __launch_bounds__(256) decode_nhwc_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ det,
                                                          int blocks_s0, int blocks_s1) {
    const int ch = 3 * P.attrs;
    int blk = blockIdx.x;
    const int si = blk >= blocks_s0 + blocks_s1 ? 2 : (blk >= blocks_s0 ? 1 : 0);
    blk -= si == 2 ? blocks_s0 + blocks_s1 : (si == 1 ? blocks_s0 : 0);
    const DecodeScale& s = P.sc[si];
    const int hw = s.h * s.w;
    const int per_img = (hw + kCellsPerBlock - 1) / kCellsPerBlock;
    const int b = blk / per_img, p0 = (blk - b * per_img) * kCellsPerBlock;
    const int np = min(kCellsPerBlock, hw - p0);
    const int y0 = p0 / s.w, x0 = p0 - y0 * s.w;
    const float* in = s.logits + ((long)b * hw + p0) * s.ld;
    float* out = det + ((long)b * P.n_total + s.row_off) * P.attrs + (long)p0 * ch;
    for (int c = threadIdx.x; c < ch; c += blockDim.x) {
        const int a = c / P.attrs, attr = c - a * P.attrs;
        int x = x0, y = y0;
        for (int i0 = 0; i0 < np; i0 += kDecodeUnroll) {
            float t[kDecodeUnroll];
#pragma unroll
            for (int u = 0; u < kDecodeUnroll; ++u) t[u] = i0 + u < np ? __ldg(in + (long)(i0 + u) * s.ld + c) : 0.f;
#pragma unroll
            for (int u = 0; u < kDecodeUnroll; ++u) {
                if (i0 + u < np) out[(long)(i0 + u) * ch + c] = decode_one(s, a, attr, x, y, t[u]);
                if (++x == s.w) { x = 0; ++y; }
            }
        }
    }
}

// NCHW (reference layout, API boundary) input: a block owns 32 consecutive cells of one image and
// scale; reads each channel row coalesced along the cells, transposes through shared memory, and
// writes the 32*3*attrs contiguous output floats coalesced.
__global__ void __launch_bounds__(256) decode_nchw_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ det,
                                                          int blocks_s0, int blocks_s1) {
    extern __shared__ float tile[];   // [ch][33]
    const int ch = 3 * P.attrs;
    int blk = blockIdx.x;
    const int si = blk >= blocks_s0 + blocks_s1 ? 2 : (blk >= blocks_s0 ? 1 : 0);
    blk -= si == 2 ? blocks_s0 + blocks_s1 : (si == 1 ? blocks_s0 : 0);
    const DecodeScale& s = P.sc[si];
    const int hw = s.h * s.w;
    const int per_img = (hw + 31) / 32;
    const int b = blk / per_img, p0 = (blk % per_img) * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int np = min(32, hw - p0);
    for (int c = warp; c < ch; c += 8) {
        float v = 0.f;
        if (lane < np) {
            const float t = s.logits[((long)b * ch + c) * hw + p0 + lane];
            const int p = p0 + lane;
            v = decode_one(s, c / P.attrs, c % P.attrs, p % s.w, p / s.w, t);
        }
        tile[c * 33 + lane] = v;
    }
    __syncthreads();
    float* o = det + ((long)b * P.n_total + s.row_off) * P.attrs + (long)p0 * ch;
    const int n = np * ch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = tile[(i % ch) * 33 + i / ch];
}

}  // namespace

cudaError_t launch_decode(const DecodeScale sc[3], int nchw, int B, int attrs, int n_total, float* det, cudaStream_t s) {
    DecodeParams P;
    P.B = B; P.attrs = attrs; P.n_total = n_total;
    P.cells_before[0] = 0;
    for (int i = 0; i < 3; ++i) {
        P.sc[i] = sc[i];
        P.cells_before[i + 1] = P.cells_before[i] + (long)B * sc[i].h * sc[i].w;
    }
    if (!nchw) {
        int nb[3];
        for (int i = 0; i < 3; ++i) nb[i] = B * ((sc[i].h * sc[i].w + kCellsPerBlock - 1) / kCellsPerBlock);
        decode_nhwc_kernel<<<nb[0] + nb[1] + nb[2], 256, 0, s>>>(P, det, nb[0], nb[1]);
    } else {
        int nb[3];
        for (int i = 0; i < 3; ++i) nb[i] = B * ((sc[i].h * sc[i].w + 31) / 32);
        const size_t smem = (size_t)3 * attrs * 33 * sizeof(float);
        static bool attr_set = false;
        if (!attr_set && smem > 48 * 1024) {
            cudaFuncSetAttribute(decode_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set = true;
        }
        decode_nchw_kernel<<<nb[0] + nb[1] + nb[2], 256, smem, s>>>(P, det, nb[0], nb[1]);
    }
    return cudaGetLastError();
}

}  // namespace yb
