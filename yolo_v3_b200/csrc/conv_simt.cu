// fp32 CUDA-core kernels: the parity-mode convolution (YB_MODE_FP32), the Cin=3 stem (both modes)
// and the layout converters at the API boundary.
//
// conv_simt_kernel computes conv_bn_relu (reference darknet.py:27-44) / res_layer (darknet.py:46-53) /
// the plain head conv (darknet.py:118) / UpsampleGroup's conv+nearest-x2 (darknet.py:159-162) as an
// NHWC implicit GEMM with fp32 FMA accumulation in a fixed (tap, channel) order, so results are
// fp32-grade against the reference's CPU path.  It is the correctness anchor, not the fast path:
// the tensor-core kernel in conv_tc.cu is what YB_MODE_FP16 runs.
#include <algorithm>

#include "yb_internal.h"

namespace yb {

namespace {

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

constexpr int TM = 64, TN = 64, TK = 16;

// packed fp32 FMA (two independent IEEE fmas per instruction on sm_100: same results as two fmaf calls)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a, const float* __restrict__ w, int cout_pad) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const long M = (long)a.B * a.Ho * a.Wo;
    const long m0 = (long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;

    // A-tile loader: pixel lp, channels [lk, lk+4)
    const int lp = tid >> 2, lk = (tid & 3) * 4;
    const long m = m0 + lp;
    const bool mvalid = m < M;
    int img = 0, oy = 0, ox = 0;
    if (mvalid) {
        img = (int)(m / ((long)a.Ho * a.Wo));
        int r = (int)(m % ((long)a.Ho * a.Wo));
        oy = r / a.Wo;
        ox = r % a.Wo;
    }
    // B-tile loader: k row bk, columns [bn, bn+4)
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    const bool nvalid = (n0 + bn) < cout_pad;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* in = reinterpret_cast<const T*>(a.in);
    const int taps = a.ks * a.ks;
    for (int tap = 0; tap < taps; ++tap) {
        const int ky = tap / a.ks, kx = tap % a.ks;
        const int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
        const bool valid = mvalid && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
        const T* src = in + (((long)img * a.H + iy) * a.W + ix) * a.in_ld;
        for (int c0 = 0; c0 < a.Cin; c0 += TK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = to_f32(src[c0 + lk + i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[lk + i][lp] = av[i];
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nvalid) bv = *reinterpret_cast<const float4*>(w + ((long)tap * a.Cin + c0 + bk) * cout_pad + n0 + bn);
            *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tm]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tn]);
                const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
                const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // epilogue: scale/bias (folded eval BN or conv bias), LeakyReLU, residual, store
    const T* res = reinterpret_cast<const T*>(a.res);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long mm = m0 + tm + i;
        if (mm >= M) continue;
        const int im = (int)(mm / ((long)a.Ho * a.Wo));
        const int r = (int)(mm % ((long)a.Ho * a.Wo));
        const int y = r / a.Wo, x = r % a.Wo;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn + j;
            if (n >= a.Cout) continue;
            float v = fmaf(acc[i][j], a.scale[n], a.bias[n]);
            if (a.leaky) v = v > 0.f ? v : v * kLeaky;
            if (res) v += to_f32(res[mm * a.res_ld + n]);
            if (a.upsample) {
                T* o = reinterpret_cast<T*>(a.out);
                const long W2 = 2L * a.Wo;
                const long base = ((long)im * 2 * a.Ho + 2 * y) * W2 + 2 * x;
                const T tv = from_f32<T>(v);
                o[(base) * a.out_ld + n] = tv;
                o[(base + 1) * a.out_ld + n] = tv;
                o[(base + W2) * a.out_ld + n] = tv;
                o[(base + W2 + 1) * a.out_ld + n] = tv;
            } else if (a.out_f32) {
                reinterpret_cast<float*>(a.out)[mm * a.out_ld + n] = v;
            } else {
                reinterpret_cast<T*>(a.out)[mm * a.out_ld + n] = from_f32<T>(v);
            }
        }
    }
}

// Stem: conv 3->32, 3x3, stride 1, pad 1 straight from the caller's NCHW fp32 image to NHWC.
// One thread per output pixel, 32 fp32 accumulators; the 32x32 tile a warp produces is contiguous
// in NHWC memory, so it is staged in shared memory and written with 16-byte coalesced stores.
// SPLIT (YB_MODE_FP32_TC): T = __half and every pixel is written as 64 values, hi[32] | lo[32] with hi = RN16(v),
// lo = RN16(v - hi) -- the activation format of the split-mode tensor-core convolutions (conv_tc.cu).
template <typename T, bool SPLIT = false>
__global__ void __launch_bounds__(128) stem_kernel(const float* __restrict__ x, T* __restrict__ out,
                                                   const float* __restrict__ w, const float* __restrict__ scale,
                                                   const float* __restrict__ bias, int B, int H, int W) {
    constexpr int CH = SPLIT ? 64 : 32;             // values stored per pixel
    __shared__ __align__(16) float ws[27 * 32];
    __shared__ float ss[32], sb[32];
    __shared__ __align__(16) T stage[4][32][CH + 8];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < 32) { ss[threadIdx.x] = scale[threadIdx.x]; sb[threadIdx.x] = bias[threadIdx.x]; }
    __syncthreads();
    const long total = (long)B * H * W;
    const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 acc2[16];                                  // 32 fp32 accumulators, updated two at a time (fma.rn.f32x2)
#pragma unroll
    for (int n = 0; n < 16; ++n) acc2[n] = make_float2(0.f, 0.f);
    if (pix < total) {
        const int b = (int)(pix / ((long)H * W));
        const int r = (int)(pix % ((long)H * W));
        const int y = r / W, xx = r % W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = y + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = xx + kx - 1;
                const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float v = ok ? __ldg(x + (((long)b * 3 + c) * H + iy) * W + ix) : 0.f;
                    // the 32 weights of this tap as eight 16-byte broadcast loads (scalar loads made the kernel LDS-bound:
                    // one shared-memory instruction per FMA)
                    const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 3 + c) * 32);
                    const float2 vv = make_float2(v, v);
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const float4 w4 = wr[n];
                        acc2[2 * n] = ffma2(vv, make_float2(w4.x, w4.y), acc2[2 * n]);
                        acc2[2 * n + 1] = ffma2(vv, make_float2(w4.z, w4.w), acc2[2 * n + 1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        float v = fmaf((n & 1) ? acc2[n >> 1].y : acc2[n >> 1].x, ss[n], sb[n]);
        v = v > 0.f ? v : v * kLeaky;
        const T hi = from_f32<T>(v);
        stage[warp][lane][n] = hi;
        if constexpr (SPLIT) stage[warp][lane][32 + n] = from_f32<T>(v - to_f32(hi));
    }
    __syncwarp();
    // the warp's 32 pixels x CH values are one contiguous run of 32*CH*sizeof(T) bytes
    const long warp_pix0 = (long)blockIdx.x * blockDim.x + warp * 32;
    constexpr int VEC = 16 / sizeof(T);              // elements per 16-byte store
    constexpr int CHUNKS = 32 * CH / VEC;            // per warp
    for (int ch = lane; ch < CHUNKS; ch += 32) {
        const int e = ch * VEC;
        const int p = e / CH, c = e % CH;
        if (warp_pix0 + p < total) {
            const uint4 v = *reinterpret_cast<const uint4*>(&stage[warp][p][c]);
            *reinterpret_cast<uint4*>(out + (warp_pix0 + p) * CH + c) = v;
        }
    }
}

// NHWC (pitch in_ld, first C channels) -> dense NCHW fp32, 32x32 smem transpose.
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, long in_ld, int C, int HW, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        tile[i][tx] = (p < HW && c < C) ? to_f32(in[((long)b * HW + p) * in_ld + c]) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        if (p < HW && c < C) out[((long)b * C + c) * HW + p] = tile[tx][i];
    }
}

// split activations [pixel][.. hi at 0, lo at `lo` ..] (pitch in_ld, first C channels) -> dense NCHW fp32
__global__ void split_to_nchw_kernel(const __half* __restrict__ in, long in_ld, long lo, int C, int HW, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        float v = 0.f;
        if (p < HW && c < C) {
            const __half* q = in + ((long)b * HW + p) * in_ld + c;
            v = __half2float(q[0]) + __half2float(q[lo]);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        if (p < HW && c < C) out[((long)b * C + c) * HW + p] = tile[tx][i];
    }
}

// fp32 [M][C] <-> split fp16 [M][2C] (hi | lo); test / API-boundary converters of yb_run_layer in YB_MODE_FP32_TC
__global__ void f32_to_split_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t M, int C) {
    const size_t n = M * (size_t)C;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const size_t m = i / C;
        const int c = (int)(i - m * C);
        const float v = in[i];
        const __half hi = __float2half_rn(v);
        out[m * 2 * C + c] = hi;
        out[m * 2 * C + C + c] = __float2half_rn(v - __half2float(hi));
    }
}
__global__ void split_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, size_t M, int C) {
    const size_t n = M * (size_t)C;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const size_t m = i / C;
        const int c = (int)(i - m * C);
        out[i] = __half2float(in[m * 2 * C + c]) + __half2float(in[m * 2 * C + C + c]);
    }
}

// nearest-neighbour x2 of an fp16 tensor into a channel slice of the concat buffer (UpsampleGroup, darknet.py:161-162): the
// 1x1 "up" convolution writes a plain tensor through its staged TMA-store epilogue and this copy replicates it -- one thread
// per 16-byte chunk of an input pixel.  `halves` = 2 in YB_MODE_FP32_TC (C hi values, then C lo values in_lo later), else 1.
__global__ void upsample2x_kernel(const __half* __restrict__ in, long in_ld, long in_lo, __half* __restrict__ out, long out_ld,
                                  long out_lo, int C, int H, int W, int halves, long total) {
    const int cpp = C / 8;                                 // chunks per half of a pixel
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ch = (int)(i % (halves * cpp));
        const long pix = i / (halves * cpp);
        const int half_i = ch / cpp, c = (ch - half_i * cpp) * 8;
        const int x = (int)(pix % W);
        const long by = pix / W;                           // b * H + y
        const uint4 v = *reinterpret_cast<const uint4*>(in + pix * in_ld + (half_i ? in_lo : 0) + c);
        __half* o = out + ((by * 2) * (2L * W) + 2 * x) * out_ld + (half_i ? out_lo : 0) + c;
        *reinterpret_cast<uint4*>(o) = v;
        *reinterpret_cast<uint4*>(o + out_ld) = v;
        *reinterpret_cast<uint4*>(o + 2L * W * out_ld) = v;
        *reinterpret_cast<uint4*>(o + 2L * W * out_ld + out_ld) = v;
    }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __float2half_rn(in[i]);
}

}  // namespace

template <typename T>
cudaError_t launch_conv_simt(const ConvArgs& a, const float* w32, int cout_pad, cudaStream_t s) {
    const long M = (long)a.B * a.Ho * a.Wo;
    dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((a.Cout + TN - 1) / TN));
    conv_simt_kernel<T><<<grid, 256, 0, s>>>(a, w32, cout_pad);
    return cudaGetLastError();
}
template cudaError_t launch_conv_simt<float>(const ConvArgs&, const float*, int, cudaStream_t);
template cudaError_t launch_conv_simt<__half>(const ConvArgs&, const float*, int, cudaStream_t);

template <typename T>
cudaError_t launch_stem(const float* x, T* out, const float* w32, const float* scale, const float* bias,
                        int B, int H, int W, cudaStream_t s) {
    const long total = (long)B * H * W;
    stem_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(x, out, w32, scale, bias, B, H, W);
    return cudaGetLastError();
}
template cudaError_t launch_stem<float>(const float*, float*, const float*, const float*, const float*, int, int, int, cudaStream_t);
template cudaError_t launch_stem<__half>(const float*, __half*, const float*, const float*, const float*, int, int, int, cudaStream_t);

cudaError_t launch_stem_split(const float* x, __half* out, const float* w32, const float* scale, const float* bias,
                              int B, int H, int W, cudaStream_t s) {
    const long total = (long)B * H * W;
    stem_kernel<__half, true><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(x, out, w32, scale, bias, B, H, W);
    return cudaGetLastError();
}
cudaError_t launch_split_to_nchw_f32(const __half* in, long in_ld, long lo, int C, int B, int HW, float* out, cudaStream_t s) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
    split_to_nchw_kernel<<<grid, block, 0, s>>>(in, in_ld, lo, C, HW, out);
    return cudaGetLastError();
}
cudaError_t launch_upsample2x(const __half* in, long in_ld, long in_lo, __half* out, long out_ld, long out_lo, int C,
                              int B, int H, int W, int halves, cudaStream_t s) {
    const long total = (long)B * H * W * halves * (C / 8);
    const int blocks = (int)std::min<long>((total + 255) / 256, 148 * 8);
    upsample2x_kernel<<<blocks, 256, 0, s>>>(in, in_ld, in_lo, out, out_ld, out_lo, C, H, W, halves, total);
    return cudaGetLastError();
}
cudaError_t launch_f32_to_split(const float* in, __half* out, size_t M, int C, cudaStream_t s) {
    f32_to_split_kernel<<<1184, 256, 0, s>>>(in, out, M, C);
    return cudaGetLastError();
}
cudaError_t launch_split_to_f32(const __half* in, float* out, size_t M, int C, cudaStream_t s) {
    split_to_f32_kernel<<<1184, 256, 0, s>>>(in, out, M, C);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_nhwc_to_nchw_f32(const T* in, long in_ld, int C, int B, int HW, float* out, cudaStream_t s) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
    nhwc_to_nchw_kernel<T><<<grid, block, 0, s>>>(in, in_ld, C, HW, out);
    return cudaGetLastError();
}
template cudaError_t launch_nhwc_to_nchw_f32<float>(const float*, long, int, int, int, float*, cudaStream_t);
template cudaError_t launch_nhwc_to_nchw_f32<__half>(const __half*, long, int, int, int, float*, cudaStream_t);

cudaError_t launch_f32_to_f16(const float* in, __half* out, size_t n, cudaStream_t s) {
    f32_to_f16_kernel<<<1184, 256, 0, s>>>(in, out, n);
    return cudaGetLastError();
}

}  // namespace yb
