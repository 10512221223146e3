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
// (x, y) arithmetic is two scalar divisions per block).  64 threads cover one cell with a 16-byte load of
// four consecutive channels each (the head maps are padded to a multiple of 16 channels, so the loads are
// aligned); the four 64-thread groups of a block take every fourth cell, with four such loads in flight
// per thread -- the kernel is latency-bound otherwise.  Stores are scalar (the 3*(5+C)-float output rows
// are only 4-byte aligned) but consecutive threads still write consecutive addresses.
constexpr int kCellsPerBlock = 32;

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
    const float* in = s.logits + ((long)b * hw + p0) * s.ld;
    float* out = det + ((long)b * P.n_total + s.row_off) * P.attrs + (long)p0 * ch;
    const int grp = threadIdx.x >> 6, q = threadIdx.x & 63;
    for (int c0 = q * 4; c0 < ch; c0 += 256) {                     // one pass for up to 256 channels
        int a[4], attr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = (c0 + u) / P.attrs; attr[u] = (c0 + u) - a[u] * P.attrs; }
        for (int i0 = grp; i0 < np; i0 += 16) {                    // cells i0, i0+4, i0+8, i0+12 of this group
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 4 * u;
                t[u] = i < np ? __ldg(reinterpret_cast<const float4*>(in + (long)i * s.ld + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 4 * u;
                if (i >= np) break;
                const int p = p0 + i;
                const int y = p / s.w, x = p - y * s.w;
                float* o = out + (long)i * ch + c0;
                const float v[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c0 + e < ch) o[e] = decode_one(s, a[e], attr[e], x, y, v[e]);
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
