// K0s (split): the Cin = 3 stem of the fp32-grade tensor-core mode (YB_MODE_FP32_TC) -- stem_halo.cu's halo-patch scheme with
// hi/lo operand pairs, so that the last convolution of that mode leaves the CUDA cores.
//
// conv_tc.cu's split mode (see its comment on split2): every operand is a pair of fp16 numbers hi = RN16(v), lo = RN16(v - hi)
// carrying 22 bits of v, and the GEMM accumulates the three significant partial products lo*hi + hi*lo + hi*hi in fp32.
// Here, per unit of 12 output rows x 38 columns:
//   * one planar TMA load of the fp32 NCHW patch (14 x 44 x 3);
//   * the converters scale every pixel by 2^8 (the image lives in [0, 1]: its lo parts would be fp16 subnormals) and write TWO
//     pixel-major arrays, hi and lo, of 32-byte rows (the pixel and its two right neighbours x 3 channels, 32B swizzle);
//   * the weights are scaled per output channel by 2^s(n) so that max|w'| lies in [2048, 4096) (their lo parts stay normal
//     fp16 numbers) and split the same way into two tables per filter row; s(n) comes from the host (kernel parameter) and
//     2^-(s(n)+8) is folded into the epilogue scale -- powers of two, exact;
//   * per M-tile and filter row three tcgen05.mma (M = 128, N = 32, K = 16): xl*wh, xh*wl, xh*wh -- the corrections first --
//     into the tile's 32 TMEM columns: nine instructions per accumulator, so the tcgen05 accumulator's truncation
//     (1.7e-8 per instruction, profiles/r02a_tc_accum_probe.txt) stays three orders below the 2e-6 bar of the layer test and
//     no second accumulation level is needed;
//   * epilogue: scale / bias (kernel parameters), LeakyReLU in fp32, hi/lo split of the result, 128-byte staged rows
//     [32 hi | 32 lo], one TMA store of a {64, 38, 12, 1} box of the [B, H, W, 64] pair tensor.
//
// reference: darknet.py:37-44 (conv_bn_relu), :66 -- in fp32, which is what this mode stands in for.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "yb_internal.h"

namespace yb {
namespace {

constexpr int kTP = 40, kTC = 38, kTR = 3, kTU = 4;
constexpr int kTUR = kTU * kTR;                 // 12 output rows per unit
constexpr int kTPatchRows = kTUR + 2;
constexpr int kTPatchPix = kTPatchRows * kTP;   // 560
constexpr int kTCvtPix = 576;
constexpr int kTTileValid = kTR * kTC;          // 114
constexpr int kTRawStages = 2, kTCvtStages = 2, kTAccStages = 4, kTRing = 2;
constexpr int kTThreads = 896;                  // warp 0 TMA, 1 + 3 MMA (1: TMEM alloc), 2 store issuer, 4-11 converters, 12-27 epilogue
constexpr int kTCvtGroups = 2;
constexpr int kTCvtPer = (kTPatchPix + 127) / 128;
constexpr int kTRawPitch = 44, kTRawMask = 3;   // fp32 images only: box start rounded down to 4 pixels (16 bytes)
constexpr uint32_t kTRawSlot = 7424;            // 3 x 14 x 44 fp32 = 7392 B
constexpr uint32_t kTCvtHalf = kTCvtPix * 32;   // 18432: one pixel-major array (hi or lo)
constexpr uint32_t kTCvtSlot = 2 * kTCvtHalf;
constexpr uint32_t kTStgSlot = 57 * 1024;       // 4 x 114 rows x 128 B = 58368
constexpr uint32_t kTOffW = 2048, kTOffStg = 8192;   // weight tables: filter row j -> hi at j*2048, lo at j*2048 + 1024
constexpr uint32_t kTOffCvt = kTOffStg + kTRing * kTStgSlot;
constexpr uint32_t kTOffRaw = kTOffCvt + kTCvtStages * kTCvtSlot;
constexpr uint32_t kTSmem = kTOffRaw + kTRawStages * kTRawSlot + 1024;
static_assert(kTOffCvt % 256 == 0 && kTCvtHalf % 256 == 0 && kTOffRaw % 128 == 0 && kTStgSlot % 1024 == 0, "alignment");
static_assert(kTU * kTTileValid * 128 <= kTStgSlot && 3 * kTPatchRows * kTRawPitch * 4 <= kTRawSlot, "slot sizes");
static_assert(kTSmem <= 227 * 1024, "shared memory budget");
constexpr float kTPixelScale = 256.f;           // 2^8

struct StemSplitArgs {
    int tiles_x, tiles_y, total_tiles;
    float sc[32];                               // scale[n] * 2^-(shift[n] + 8)
    float bi[32];
    int shift[32];                              // weights of channel n are multiplied by 2^shift[n] before the hi/lo split
    const float* w32;                           // [27][32] fp32, k = (ky*3+kx)*3 + c
    int* dbg;
};

__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}

struct SplitWalk {
    int tx, ty, img, dx, dy, dimg, tiles_x, tiles_y;
    __device__ __forceinline__ SplitWalk(const StemSplitArgs& a, int first, int step) : tiles_x(a.tiles_x), tiles_y(a.tiles_y) {
        tx = first % tiles_x; int t = first / tiles_x;
        ty = t % tiles_y; img = t / tiles_y;
        dx = step % tiles_x; t = step / tiles_x;
        dy = t % tiles_y; dimg = t / tiles_y;
    }
    __device__ __forceinline__ void next() {
        tx += dx;
        int carry = 0;
        if (tx >= tiles_x) { tx -= tiles_x; carry = 1; }
        ty += dy + carry;
        carry = 0;
        if (ty >= tiles_y) { ty -= tiles_y; carry = 1; }
        img += dimg + carry;
    }
    __device__ __forceinline__ int x0() const { return tx * kTC; }
    __device__ __forceinline__ int y0() const { return ty * kTUR; }
};

// hi / lo halves of two fp32 values (conv_tc.cu's split2)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kTThreads, 1)
stem_halo_split_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmOut, const StemSplitArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    // header: rfull[4] | rempty[4] | cfull[4] | cempty[4] | tfull[4] | tempty[4] | sempty[4] | sready[4] | tmem_ptr
    const uint32_t rfull0 = base, rempty0 = base + 32, cfull0 = base + 64, cempty0 = base + 96;
    const uint32_t tfull0 = base + 128, tempty0 = base + 160, sempty0 = base + 192, sready0 = base + 224;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 256);
    const uint32_t wsm = base + kTOffW, stg0 = base + kTOffStg, cvt0 = base + kTOffCvt, raw0 = base + kTOffRaw;
    constexpr int RPLANE = kTPatchRows * kTRawPitch;
    constexpr uint32_t RAW_BYTES = 3u * RPLANE * sizeof(float);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int unit_first = blockIdx.x, unit_step = gridDim.x;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmOut); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(rfull0 + 8 * s, 1); mbar_init(rempty0 + 8 * s, 4);
            mbar_init(cfull0 + 8 * s, 4); mbar_init(cempty0 + 8 * s, 1);
            mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 16);
            mbar_init(sempty0 + 8 * s, 1); mbar_init(sready0 + 8 * s, 16);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 32 * kTU * kTAccStages);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 192) {
        // weight tables: filter row ky, 16-byte chunk h, output channel n -> chunk h of the 32-byte K-major rows (32B swizzle) of
        // the hi and of the lo table: the nine weights (ky,0..2) x 3 channels scaled by 2^shift[n], + 7 zeros
        const int idx = threadIdx.x - 128;
        const int ky = idx >> 6, h = (idx >> 5) & 1, n = idx & 31;
        uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
        const int sh = a.shift[n];
        if (h == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                split_pair(ldexpf(__ldg(a.w32 + (ky * 9 + 2 * i) * 32 + n), sh), ldexpf(__ldg(a.w32 + (ky * 9 + 2 * i + 1) * 32 + n), sh), hi[i], lo[i]);
        } else {
            split_pair(ldexpf(__ldg(a.w32 + (ky * 9 + 8) * 32 + n), sh), 0.f, hi[0], lo[0]);
        }
        const uint32_t off = kTOffW + ky * 2048 + n * 32 + ((h ^ ((n >> 2) & 1)) << 4);
        *reinterpret_cast<uint4*>(gen + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(gen + off + 1024) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    // the converted stages start as zeros: rows 560..575 and bytes 18..31 of every row are never written
    for (uint32_t i = threadIdx.x; i < kTCvtStages * kTCvtSlot / 16; i += kTThreads)
        *reinterpret_cast<uint4*>(gen + kTOffCvt + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait_prior();

    if (warp == 0) {
        // ===== TMA producer: one planar patch per unit =====
        Slot<kTRawStages> rs(0);
        SplitWalk t(a, unit_first, unit_step);
        for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, t.next(), rs.advance(1)) {
            mbar_wait(rempty0 + 8 * rs.i, rs.ph ^ 1, a.dbg, 0, (int)rs.i);
            if (elect_one()) {
                mbar_arrive_expect_tx(rfull0 + 8 * rs.i, RAW_BYTES);
                tma_load_4d(&tmX, raw0 + rs.i * kTRawSlot, rfull0 + 8 * rs.i, (t.x0() - 1) & ~kTRawMask, t.y0() - 1, 0, t.img);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers (two warps, alternate units): per M-tile and filter row xl*wh, xh*wl, xh*wh; consecutive
        // instructions go to different M-tiles =====
        const int k = warp >> 1;
        const uint32_t idesc = make_idesc(32);
        uint64_t adH[3], bdH[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            adH[j] = desc_sw32(cvt0) + (uint64_t)(2 * j * kTP);
            bdH[j] = desc_sw32(wsm + j * 2048);
        }
        constexpr uint64_t A_LO = kTCvtHalf >> 4, B_LO = 1024 >> 4;
        Slot<kTCvtStages> cs((uint32_t)k % kTCvtStages);
        Slot<kTAccStages> as((uint32_t)k);
        for (int unit = unit_first + k * unit_step; unit < a.total_tiles; unit += 2 * unit_step, as.advance(2)) {
            mbar_wait(tempty0 + 8 * as.i, as.ph ^ 1, a.dbg, 1, 100 + (int)as.i);
            mbar_wait(cfull0 + 8 * cs.i, cs.ph, a.dbg, 1, (int)cs.i);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t a_off = (uint64_t)(cs.i * (kTCvtSlot >> 4));
                const uint32_t d_tmem = tmem_base + as.i * (32 * kTU);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) {       // 0: xl*wh, 1: xh*wl, 2: xh*wh
                        const uint64_t ad = adH[j] + a_off + (p == 0 ? A_LO : 0);
                        const uint64_t bd = bdH[j] + (p == 1 ? B_LO : 0);
#pragma unroll
                        for (int t = 0; t < kTU; ++t)
                            umma_f16(d_tmem + t * 32, ad + (uint64_t)(2 * t * kTR * kTP), bd, idesc, (j | p) != 0);
                    }
                }
                umma_commit(cempty0 + 8 * cs.i);
                umma_commit(tfull0 + 8 * as.i);
            }
            __syncwarp();
            cs.ph ^= 1u;                           // two converted stages, two issuers: issuer k always uses stage k
        }
    } else if (warp == 2) {
        // ===== store issuer: one {64 (hi | lo), 38, 12, 1} box per unit =====
        if (lane == 0) {
            SplitWalk t(a, unit_first, unit_step);
            Slot<kTRing> ss(0);
            uint32_t prev = 0;
            bool first = true;
            for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, t.next(), ss.advance(1)) {
                mbar_wait(sready0 + 8 * ss.i, ss.ph, a.dbg, 4, 700 + (int)ss.i);
                tma_store_4d(&tmOut, stg0 + ss.i * kTStgSlot, 0, t.x0(), t.y0(), t.img);
                tma_store_commit();
                if (!first) {
                    tma_store_wait_read<1>();
                    mbar_arrive(sempty0 + 8 * prev);
                }
                first = false;
                prev = ss.i;
            }
            tma_store_wait_all();
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + 4 * kTCvtGroups) {
        // ===== converters (two groups of four warps, alternate units = alternate stages) =====
        const int grp = (warp - 4) >> 2;
        const int tid = threadIdx.x & 127;
        int q[kTCvtPer], room[kTCvtPer];
#pragma unroll
        for (int i = 0; i < kTCvtPer; ++i) {
            const int p = tid + 128 * i;
            q[i] = (p / kTP) * kTRawPitch + p % kTP;
            room[i] = min(2, kTP - 1 - p % kTP);
        }
        const bool last_ok = tid + 128 * (kTCvtPer - 1) < kTPatchPix;
        const uint32_t c0 = ((tid >> 2) & 1) << 4;
        uint32_t ph = 0;                               // group g always works in raw stage g and converted stage g
        int tx = (unit_first + grp * unit_step) % a.tiles_x;
        const int dtx = (kTCvtGroups * unit_step) % a.tiles_x;
        for (int unit = unit_first + grp * unit_step; unit < a.total_tiles; unit += kTCvtGroups * unit_step, ph ^= 1u) {
            const int d = (tx * kTC - 1) & kTRawMask;
            tx += dtx; if (tx >= a.tiles_x) tx -= a.tiles_x;
            mbar_wait(rfull0 + 8 * grp, ph, a.dbg, 5, grp);
            mbar_wait(cempty0 + 8 * grp, ph ^ 1, a.dbg, 5, 100 + grp);
            const float* rp = reinterpret_cast<const float*>(gen + kTOffRaw + grp * kTRawSlot) + d;
            uint8_t* dst = gen + kTOffCvt + grp * kTCvtSlot + tid * 32;
#pragma unroll
            for (int i = 0; i < kTCvtPer; ++i) {
                if (i < kTCvtPer - 1 || last_ok) {
                    float v[10];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        v[3 * k] = v[3 * k + 1] = v[3 * k + 2] = 0.f;
                        if (k <= room[i]) {
                            v[3 * k] = rp[q[i] + k] * kTPixelScale; v[3 * k + 1] = rp[RPLANE + q[i] + k] * kTPixelScale;
                            v[3 * k + 2] = rp[2 * RPLANE + q[i] + k] * kTPixelScale;
                        }
                    }
                    v[9] = 0.f;
                    uint32_t hi[5], lo[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) split_pair(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
                    *reinterpret_cast<uint4*>(dst + i * 4096 + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint32_t*>(dst + i * 4096 + (c0 ^ 16u)) = hi[4];
                    *reinterpret_cast<uint4*>(dst + kTCvtHalf + i * 4096 + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint32_t*>(dst + kTCvtHalf + i * 4096 + (c0 ^ 16u)) = lo[4];
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(rempty0 + 8 * grp); mbar_arrive(cfull0 + 8 * grp); }
        }
    } else if (warp >= 12) {
        // ===== epilogue: warp (q, t) drains lane quarter q of M-tile t; staging row = t*114 + r*38 + c, 128 bytes: 32 hi | 32 lo =====
        const int q = warp & 3, t = (warp - 12) >> 2;
        const int m = q * 32 + lane;
        const int r = m / kTP, c = m - r * kTP;
        const bool valid = r < kTR && c < kTC;
        const int mp = t * kTTileValid + r * kTC + c;
        const int xr = mp & 7;
        Slot<kTAccStages> as(0);
        Slot<kTRing> ss(0);
        for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, as.advance(1), ss.advance(1)) {
            mbar_wait(tfull0 + 8 * as.i, as.ph, a.dbg, 2, 200 + (int)as.i);
            tc_fence_after();
            uint32_t r0[16], r1[16];
            const uint32_t taddr = tmem_base + as.i * (32 * kTU) + t * 32 + ((uint32_t)(q * 32) << 16);
            tmem_ld16(taddr, r0);
            tmem_ld16(taddr + 16, r1);
            mbar_wait(sempty0 + 8 * ss.i, ss.ph ^ 1, a.dbg, 2, 500 + (int)ss.i);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as.i);
            if (valid) {
                uint8_t* srow = gen + kTOffStg + ss.i * kTStgSlot + mp * 128;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {              // channels 8cc .. 8cc+7: hi chunk cc, lo chunk 4 + cc
                    const uint32_t* rr = cc < 2 ? r0 : r1;
                    const int j0 = 8 * (cc & 1);
                    uint4 ph4, pl4;
                    uint32_t* ph = reinterpret_cast<uint32_t*>(&ph4);
                    uint32_t* pl = reinterpret_cast<uint32_t*>(&pl4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ch = 8 * cc + 2 * e;
                        float y0 = fmaf(__uint_as_float(rr[j0 + 2 * e]), a.sc[ch], a.bi[ch]);
                        float y1 = fmaf(__uint_as_float(rr[j0 + 2 * e + 1]), a.sc[ch + 1], a.bi[ch + 1]);
                        y0 = leaky(y0); y1 = leaky(y1);
                        split_pair(y0, y1, ph[e], pl[e]);
                    }
                    *reinterpret_cast<uint4*>(srow + ((cc ^ xr) << 4)) = ph4;
                    *reinterpret_cast<uint4*>(srow + (((4 + cc) ^ xr) << 4)) = pl4;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sready0 + 8 * ss.i);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32 * kTU * kTAccStages);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;

}  // namespace

bool stem_split_supported(int W) { return W % 4 == 0; }

// Per output channel: the power of two that brings max|w| into [2048, 4096); w = host copy of the stem's fp32 weights
// [27][32] (k = (ky*3+kx)*3 + c).  sc_eff = scale * 2^-(shift + 8).
void stem_split_host_params(const float* w27x32, const float* scale, const float* bias, float* sc_eff, float* bi_out, int* shift) {
    for (int n = 0; n < 32; ++n) {
        float mx = 0.f;
        for (int k = 0; k < 27; ++k) mx = std::max(mx, std::fabs(w27x32[k * 32 + n]));
        int e = 0;
        if (mx > 0.f && std::isfinite(mx)) std::frexp(mx, &e);      // mx = m * 2^e, m in [0.5, 1)
        shift[n] = mx > 0.f ? 12 - e : 0;                           // max|w * 2^shift| in [2048, 4096)
        shift[n] = std::max(-40, std::min(40, shift[n]));
        sc_eff[n] = std::ldexp(scale[n], -(shift[n] + 8));
        bi_out[n] = bias[n];
    }
}

std::string stem_split_make_plan(StemHaloPlan& p, __half* out, long out_ld, int B, int H, int W, int num_sms) {
    if (!g_enc) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeTiled not available from the driver";
        g_enc = reinterpret_cast<EncodeTiledFn>(f);
    }
    if (out_ld % 8) return "stem output pitch must be a multiple of 8 channels";
    p.tiles_x = (W + kTC - 1) / kTC;
    p.tiles_y = (H + kTUR - 1) / kTUR;
    p.total_tiles = B * p.tiles_x * p.tiles_y;
    p.grid = std::min(p.total_tiles, num_sms);
    p.smem = kTSmem;
    const cuuint32_t es4[4] = {1, 1, 1, 1};
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};       // 32 hi | 32 lo
    cuuint64_t st[3] = {(cuuint64_t)out_ld * 2, (cuuint64_t)W * out_ld * 2, (cuuint64_t)H * W * out_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kTC, (cuuint32_t)kTUR, 1};
    CUresult r = g_enc(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out, dims, st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled(split stem output) failed with CUresult " + std::to_string((int)r);
    return "";
}

cudaError_t stem_split_launch(const StemHaloPlan& p, const float* x, int B, int H, int W, const float* w32, const float* sc_eff_host,
                              const float* bias_host, const int* shift_host, int* dbg, cudaStream_t s) {
    CUtensorMap tmX;
    {
        const cuuint32_t es4[4] = {1, 1, 1, 1};
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        cuuint64_t st[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
        cuuint32_t box[4] = {(cuuint32_t)kTRawPitch, (cuuint32_t)kTPatchRows, 3, 1};
        CUresult r = g_enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    static PerDeviceOnce attr_once;
    {
        cudaError_t e = attr_once.run([] {
            return cudaFuncSetAttribute(stem_halo_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem);
        });
        if (e != cudaSuccess) return e;
    }
    StemSplitArgs a;
    a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.total_tiles = p.total_tiles;
    for (int i = 0; i < 32; ++i) { a.sc[i] = sc_eff_host[i]; a.bi[i] = bias_host[i]; a.shift[i] = shift_host[i]; }
    a.w32 = w32; a.dbg = dbg;
    static const bool pdl = !(tune_env("YB_TC_PDL") && atoi(tune_env("YB_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(kTThreads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, stem_halo_split_kernel, tmX, p.tmOut, a);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace yb
