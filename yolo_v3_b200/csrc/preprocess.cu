// N1: pre-process -- utils.load_image(path, mode, dim) after the file decode (reference utils.py:60-72):
//   'letterbox': letterbox_image (utils.py:44-57) = cv2.resize(img, (box_w, box_h), INTER_CUBIC) pasted at the centred
//                offset of a grey (128) canvas;   'resize': cv2.resize(img, dim) (INTER_LINEAR, utils.py:69);
// then torch.from_numpy(img).float().permute(2,0,1) / 255.
//
// The resize arithmetic is OpenCV's (third-party, not vendored in the reference; opencv-python 4.13.0 in this image).
// This kernel restates OpenCV's own portable code path for 8-bit images bit for bit (modules/imgproc/src/resize.cpp:
// resizeGeneric_ + HResizeCubic<uchar,int,short> + VResizeCubic<..., VResizeCubicVec_32s8u>):
//   * per axis: f = (float)((d + 0.5) * scale - 0.5) evaluated in double, s = floor(f), four taps
//     cvRound(interpolateCubic(f - s) * 2048) with A = -0.75, every fp32 operation rounded separately;
//   * horizontal pass: exact int32 sum of the four border-replicated taps;
//   * vertical pass: fp32  S0*b0 + (S1*b1 + (S2*b2 + S3*b3)),  b = tap / 2^22, no FMA, round-half-even, saturate --
//     except the last (box_w*3) % 8 elements of every row, which OpenCV's scalar tail computes as
//     (S0*b0 + S1*b1 + S2*b2 + S3*b3 + 2^21) >> 22 in int32.
// (OpenCV builds with Intel IPP route the call to ippiResizeCubic, which differs from this by at most one grey level;
// tests/golden/letterbox_golden.npz holds both, see oracle/yolo_oracle.py.)
//
// HBM-bound byte work: one thread per canvas pixel reads its 4x4 source neighbourhood (3 bytes per pixel, L1/L2
// resident: neighbouring threads share 3/4 of it) and writes the three fp32 planes with coalesced stores;
// algorithmic traffic = source bytes + 12 bytes per canvas pixel.
#include <cstdlib>

#include "yb_internal.h"

namespace yb {
namespace {

__device__ __forceinline__ void cubic_taps(int d, double scale, int& s0, int (&tap)[4]) {
    const double fd = __dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    float f = __double2float_rn(fd);
    const float fl = floorf(f);
    s0 = (int)fl;
    const float x = __fsub_rn(f, fl);
    const float A = -0.75f;
    const float x1 = __fadd_rn(x, 1.f), xm = __fsub_rn(1.f, x);
    float c[4];
    // ((A*(x+1) - 5A)*(x+1) + 8A)*(x+1) - 4A
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), -3.75f), x1), -6.0f), x1), -3.0f);
    // ((A+2)*x - (A+3))*x*x + 1
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, x), 2.25f), x), x), 1.f);
    // ((A+2)*(1-x) - (A+3))*(1-x)*(1-x) + 1
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, xm), 2.25f), xm), xm), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) tap[k] = max(-32768, min(32767, __float2int_rn(__fmul_rn(c[k], 2048.f))));   // saturate_cast<short>
}

// INTER_LINEAR taps (cv2.resize(img, dim), utils.py:69): the x axis clamps offset and fraction at the borders, the y
// axis keeps its fraction and clips only the row indices (resize.cpp set-up loops).
__device__ __forceinline__ void linear_taps(int d, double scale, int ssize, bool clamp, int& s0, int (&tap)[2]) {
    const double fd = __dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    const float f = __double2float_rn(fd);
    const float fl = floorf(f);
    s0 = (int)fl;
    float x = __fsub_rn(f, fl);
    if (clamp) {
        if (s0 < 0) { x = 0.f; s0 = 0; }
        if (s0 >= ssize - 1) { x = 0.f; s0 = ssize - 1; }
    }
    tap[0] = max(-32768, min(32767, __float2int_rn(__fmul_rn(__fsub_rn(1.f, x), 2048.f))));
    tap[1] = max(-32768, min(32767, __float2int_rn(__fmul_rn(x, 2048.f))));
}

// One canvas pixel inside the box, evaluated directly from the source image (4x4 / 2x2 neighbourhood).
__device__ __forceinline__ void direct_pixel(const LbImage& im, int dx, int dy, int (&val)[3]) {
    if (im.interp == 1) {
        // bilinear: exact int32 horizontal pass, then ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2
        // (VResizeLinearVec_32s8u and its scalar tail compute the same expression)
        int sx, sy, xa[2], yb[2];
        linear_taps(dx, im.scale_x, im.sw, true, sx, xa);
        linear_taps(dy, im.scale_y, im.sh, false, sy, yb);
        const int x0 = sx, x1 = min(sx + 1, im.sw - 1);
        const unsigned char* r0 = im.src + (size_t)max(0, min(im.sh - 1, sy)) * im.sw * 3;
        const unsigned char* r1 = im.src + (size_t)max(0, min(im.sh - 1, sy + 1)) * im.sw * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int S0 = (int)__ldg(r0 + x0 * 3 + c) * xa[0] + (int)__ldg(r0 + x1 * 3 + c) * xa[1];
            const int S1 = (int)__ldg(r1 + x0 * 3 + c) * xa[0] + (int)__ldg(r1 + x1 * 3 + c) * xa[1];
            const int t = ((yb[0] * (S0 >> 4)) >> 16) + ((yb[1] * (S1 >> 4)) >> 16);
            val[c] = max(0, min(255, (t + 2) >> 2));
        }
        return;
    }
    int sx, sy, xa[4], yb[4];
    cubic_taps(dx, im.scale_x, sx, xa);
    cubic_taps(dy, im.scale_y, sy, yb);
    int H[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ry = max(0, min(im.sh - 1, sy - 1 + k));
        const unsigned char* row = im.src + (size_t)ry * im.sw * 3;
        H[k][0] = H[k][1] = H[k][2] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rx = max(0, min(im.sw - 1, sx - 1 + j));
            const unsigned char* px = row + rx * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) H[k][c] += (int)__ldg(px + c) * xa[j];
        }
    }
    const int vec_end = (im.box_w * 3) / 8 * 8;
    const float sc = 1.f / 4194304.f;                                  // 1 / (2048 * 2048)
    float bf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) bf[k] = __fmul_rn((float)yb[k], sc);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int r;
        if (dx * 3 + c < vec_end) {
            float t = __fmul_rn((float)H[3][c], bf[3]);
            t = __fadd_rn(__fmul_rn((float)H[2][c], bf[2]), t);
            t = __fadd_rn(__fmul_rn((float)H[1][c], bf[1]), t);
            t = __fadd_rn(__fmul_rn((float)H[0][c], bf[0]), t);
            r = __float2int_rn(t);
        } else {
            r = (H[0][c] * yb[0] + H[1][c] * yb[1] + H[2][c] * yb[2] + H[3][c] * yb[3] + (1 << 21)) >> 22;
        }
        val[c] = max(0, min(255, r));
    }
}

__device__ __forceinline__ void store_pixel(float* __restrict__ out, unsigned char* __restrict__ canvas, size_t plane, int b,
                                            int canvas_w, int x, int y, const int (&val)[3]) {
    if (out) {                                                             // float() / 255, HWC -> CHW (utils.py:71)
        float* o = out + (size_t)b * 3 * plane + (size_t)y * canvas_w + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c * plane] = __fdiv_rn((float)val[c], 255.f);
    }
    if (canvas) {                                                          // the HWC canvas letterbox_image returns
        unsigned char* q = canvas + ((size_t)b * plane + (size_t)y * canvas_w + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) q[c] = (unsigned char)val[c];
    }
}

// First version: one thread per canvas pixel.  Kept for A/B timing (YB_LB_DIRECT=1).
__global__ void __launch_bounds__(256) letterbox_kernel(const LbImage* __restrict__ imgs, int canvas_h, int canvas_w,
                                                        float* __restrict__ out, unsigned char* __restrict__ canvas) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
    if (x >= canvas_w || y >= canvas_h) return;
    const LbImage im = imgs[b];
    const int dx = x - im.box_x, dy = y - im.box_y;
    int val[3] = {128, 128, 128};
    if (dx >= 0 && dx < im.box_w && dy >= 0 && dy < im.box_h) direct_pixel(im, dx, dy, val);
    store_pixel(out, canvas, (size_t)canvas_h * canvas_w, b, canvas_w, x, y, val);
}

// ---- two-pass tiled version -------------------------------------------------------------------------------
// The direct kernel recomputes the horizontal pass of every source row four times (once per vertical tap of every
// output row that uses it) and every thread re-derives taps its neighbours share; ncu shows it instruction-issue-
// bound (84 % issue-active at 11 % of DRAM throughput).  Here a block owns a 32 x 32 tile of the canvas: the taps
// of its 32 columns and 32 rows are computed once into shared memory, pass 1 computes the horizontal sums of
// exactly the source rows the tile needs, once each (kept as fp32: |sum| < 2^24, so exact), pass 2 combines them
// vertically, and the /255 comes from a 256-entry table of the IEEE quotients.  Same arithmetic, same results.
constexpr int kLbTile = 32;
constexpr int kLbMaxRows = 100;            // source rows a tile may need: up to ~3x vertical downscale (else: direct)

__global__ void __launch_bounds__(256) letterbox_tiled_kernel(const LbImage* __restrict__ imgs, int canvas_h, int canvas_w,
                                                              float* __restrict__ out, unsigned char* __restrict__ canvas) {
    __shared__ float Hs[kLbMaxRows * kLbTile * 3];
    __shared__ int xtap[kLbTile][4], xoff[kLbTile][4];       // per column: taps, byte offset of the (clamped) source pixel
    __shared__ int ytap[kLbTile][4], ysrc[kLbTile];          // per row: taps, unclamped first source row (sy)
    __shared__ float ybf[kLbTile][4];                        // per row: taps / 2^22 (cubic vector path)
    __shared__ float lut[256];
    const int b = blockIdx.z, tid = threadIdx.x;
    const LbImage im = imgs[b];
    const size_t plane = (size_t)canvas_h * canvas_w;
    const int tx0 = blockIdx.x * kLbTile, ty0 = blockIdx.y * kLbTile;
    const int lx = tid & 31, ly = tid >> 5;                  // 32 columns x 8 rows of threads
    const int x = tx0 + lx;
    const int ntap = im.interp == 1 ? 2 : 4;
    // the rows / columns of the tile that lie inside the box, in box coordinates
    const int dy_lo = max(ty0 - im.box_y, 0);
    const int dy_hi = min(min(ty0 + kLbTile, canvas_h) - im.box_y, im.box_h);                   // rows [dy_lo, dy_hi)
    const int dx = x - im.box_x;
    const bool col_in = x < canvas_w && dx >= 0 && dx < im.box_w;
    const bool tile_in = dy_lo < dy_hi && tx0 + kLbTile > im.box_x && tx0 < im.box_x + im.box_w;
    lut[tid] = __fdiv_rn((float)tid, 255.f);
    if (tile_in && ly == 0) {                                // warp 0: taps of the tile's columns
        int sx = 0, t4[4] = {0, 0, 0, 0}, o4[4] = {0, 0, 0, 0};
        if (col_in) {
            if (im.interp == 1) {
                int t2[2];
                linear_taps(dx, im.scale_x, im.sw, true, sx, t2);
                t4[0] = t2[0]; t4[1] = t2[1];
                o4[0] = sx * 3; o4[1] = min(sx + 1, im.sw - 1) * 3;
            } else {
                cubic_taps(dx, im.scale_x, sx, t4);
#pragma unroll
                for (int j = 0; j < 4; ++j) o4[j] = max(0, min(im.sw - 1, sx - 1 + j)) * 3;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { xtap[lx][j] = t4[j]; xoff[lx][j] = o4[j]; }
    }
    if (tile_in && ly == 1) {                                // warp 1: taps of the tile's rows
        const int dy = ty0 + lx - im.box_y;
        int sy = 0, t4[4] = {0, 0, 0, 0};
        if (dy >= dy_lo && dy < dy_hi) {
            if (im.interp == 1) {
                int t2[2];
                linear_taps(dy, im.scale_y, im.sh, false, sy, t2);
                t4[0] = t2[0]; t4[1] = t2[1];
            } else {
                cubic_taps(dy, im.scale_y, sy, t4);
            }
        }
        ysrc[lx] = sy;
#pragma unroll
        for (int k = 0; k < 4; ++k) { ytap[lx][k] = t4[k]; ybf[lx][k] = __fmul_rn((float)t4[k], 1.f / 4194304.f); }
    }
    __syncthreads();
    int r_lo = 0, r_hi = -1;
    if (tile_in) {
        const int s_first = ysrc[dy_lo + im.box_y - ty0], s_last = ysrc[dy_hi - 1 + im.box_y - ty0];
        const int first = im.interp == 1 ? s_first : s_first - 1, last = im.interp == 1 ? s_last + 1 : s_last + 2;
        r_lo = max(0, min(im.sh - 1, first));
        r_hi = max(0, min(im.sh - 1, last));
    }
    const int nrows = r_hi - r_lo + 1;
    const bool staged = tile_in && nrows <= kLbMaxRows;
    // ---- pass 1: horizontal sums of source rows r_lo..r_hi for this thread's column
    if (staged && col_in) {
        int xa[4], xo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { xa[j] = xtap[lx][j]; xo[j] = xoff[lx][j]; }
        for (int r = ly; r < nrows; r += 8) {
            const unsigned char* row = im.src + (size_t)(r_lo + r) * im.sw * 3;
            int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < ntap) {
                    const unsigned char* px = row + xo[j];
                    h0 += (int)__ldg(px) * xa[j];
                    h1 += (int)__ldg(px + 1) * xa[j];
                    h2 += (int)__ldg(px + 2) * xa[j];
                }
            }
            float* h = Hs + (r * kLbTile + lx) * 3;
            h[0] = (float)h0; h[1] = (float)h1; h[2] = (float)h2;
        }
    }
    __syncthreads();
    // ---- pass 2: vertical combination, four output rows per thread
    if (x >= canvas_w) return;
    const int vec_end = (im.box_w * 3) / 8 * 8;
    for (int yy = ly; yy < kLbTile; yy += 8) {
        const int y = ty0 + yy;
        if (y >= canvas_h) break;
        const int dy = y - im.box_y;
        int val[3] = {128, 128, 128};
        if (col_in && dy >= 0 && dy < im.box_h) {
            if (!staged) {
                direct_pixel(im, dx, dy, val);              // strong downscale: more source rows than the buffer holds
            } else if (im.interp == 1) {
                const int sy = ysrc[yy], b0 = ytap[yy][0], b1 = ytap[yy][1];
                const float* h0 = Hs + ((max(0, min(im.sh - 1, sy)) - r_lo) * kLbTile + lx) * 3;
                const float* h1 = Hs + ((max(0, min(im.sh - 1, sy + 1)) - r_lo) * kLbTile + lx) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int S0 = __float2int_rn(h0[c]), S1 = __float2int_rn(h1[c]);
                    val[c] = max(0, min(255, ((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16)) + 2) >> 2));
                }
            } else {
                const int sy = ysrc[yy];
                const float* h[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = Hs + ((max(0, min(im.sh - 1, sy - 1 + k)) - r_lo) * kLbTile + lx) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float H0 = h[0][c], H1 = h[1][c], H2 = h[2][c], H3 = h[3][c];
                    int r;
                    if (dx * 3 + c < vec_end) {
                        float t = __fmul_rn(H3, ybf[yy][3]);
                        t = __fadd_rn(__fmul_rn(H2, ybf[yy][2]), t);
                        t = __fadd_rn(__fmul_rn(H1, ybf[yy][1]), t);
                        t = __fadd_rn(__fmul_rn(H0, ybf[yy][0]), t);
                        r = __float2int_rn(t);
                    } else {
                        r = (__float2int_rn(H0) * ytap[yy][0] + __float2int_rn(H1) * ytap[yy][1] + __float2int_rn(H2) * ytap[yy][2] +
                             __float2int_rn(H3) * ytap[yy][3] + (1 << 21)) >> 22;
                    }
                    val[c] = max(0, min(255, r));
                }
            }
        }
        if (out) {
            float* o = out + (size_t)b * 3 * plane + (size_t)y * canvas_w + x;
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c * plane] = lut[val[c]];
        }
        if (canvas) {
            unsigned char* q = canvas + ((size_t)b * plane + (size_t)y * canvas_w + x) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) q[c] = (unsigned char)val[c];
        }
    }
}

}  // namespace

cudaError_t launch_letterbox(const LbImage* imgs_dev, int B, int canvas_h, int canvas_w, float* out, unsigned char* canvas,
                             cudaStream_t s) {
    static const bool direct = tune_env("YB_LB_DIRECT") && atoi(tune_env("YB_LB_DIRECT")) != 0;   // first version, kept for A/B timing
    if (direct) {
        const dim3 grid((canvas_w + 31) / 32, (canvas_h + 7) / 8, B);
        letterbox_kernel<<<grid, 256, 0, s>>>(imgs_dev, canvas_h, canvas_w, out, canvas);
    } else {
        const dim3 grid((canvas_w + kLbTile - 1) / kLbTile, (canvas_h + kLbTile - 1) / kLbTile, B);
        letterbox_tiled_kernel<<<grid, 256, 0, s>>>(imgs_dev, canvas_h, canvas_w, out, canvas);
    }
    return cudaGetLastError();
}

}  // namespace yb
