// K0s: the Cin = 3 stem (3 -> 32, 3x3, stride 1) from a HALO PATCH, straight from the caller's NCHW image.
//
// The first tensor-core stem (round 1-2, conv_tc.cu) built one im2col row per output pixel in registers: 27 scalar loads,
// 14 conversions and four 16-byte stores per pixel -- ~1 800 warp instructions per 128-pixel tile, instruction-issue-bound
// at 0.236 ms against 0.140 ms of HBM time (profiles/r02_full_raw_layer0.csv).  Here every input pixel is touched ONCE:
//
//   * a tile is 3 output rows x 38 output columns (the geometry of conv_halo.cu); its input patch -- 3 planes x 5 x 40
//     pixels of the NCHW image, zero-filled outside the image by TMA (= the convolution's padding) -- is one 4-D TMA load
//     of the caller's tensor (fp32 or fp16; the box is widened to a 16-byte aligned first column, see RawGeom);
//   * converter warps turn the planar patch into pixel-major fp16: 16 bytes per pixel (3 channels + 5 zeros) at patch
//     index p = r*40 + c, one pass, 3 shared-memory loads + 1 store per pixel;
//   * the tensor core reads that array through NON-swizzled K-major descriptors: 8 fp16 of one pixel are one 16-byte
//     core-matrix row, eight consecutive pixels one core matrix (SBO = 128 B), and the descriptor's leading-dimension
//     byte offset -- the distance between the two K halves of a K = 16 instruction -- is free: LBO = 16 B makes the second
//     half "the next pixel", LBO = 38 * 16 B "next row, two columns back".  So one tcgen05.mma covers TWO taps of all 128
//     GEMM rows (row m = r*40 + c reads patch pixels m + tap offset), and the nine taps are five instructions
//     (M = 128, N = 32, K = 16) against a 5 KB weight table laid out the same way;
//   * epilogue as everywhere (TMEM -> scale/bias -> LeakyReLU -> fp16 -> 64B-swizzled staging -> TMA store of a
//     {32 ch, 38, 3, 1} box, clipped at the image border); four epilogue groups take tiles round-robin because the
//     per-tile hand-over chain (~800 cycles) would otherwise pace the kernel.
//
// reference: darknet.py:37-44 (conv_bn_relu: Conv2d(3, 32, 3, 1, 1, bias=False) -> BatchNorm2d -> LeakyReLU(0.1)), :66.
#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "yb_internal.h"

namespace yb {
namespace {

constexpr int kSP = 40;                        // patch pitch in pixels = tile columns + 2
constexpr int kSC = 38;                        // output columns per M-tile
constexpr int kSR = 3;                         // output rows per M-tile (GEMM row m = r*40 + c)
constexpr int kSU = 4;                         // M-tiles per UNIT: 12 output rows x 38 columns from one 14 x 40 patch
constexpr int kSUR = kSU * kSR;                // output rows per unit
constexpr int kSPatchRows = kSUR + 2;
constexpr int kSPatchPix = kSPatchRows * kSP;  // 560 pixels per plane
constexpr int kSCvtPix = 576;                  // pixel-major rows per stage: M-tile 3's last K half reads up to 360 + 82 + 127 = 569
constexpr int kSTileValid = kSR * kSC;         // 114 staged rows per M-tile
constexpr int kSRawStages = 4, kSCvtStages = 4, kSAccStages = 4, kSRing = 3;
constexpr int kSThreads = 896;                 // warp 0 TMA, 1 + 3 MMA (1: TMEM alloc), 2 store issuer, 4-11 converters, 12-27 epilogue
constexpr int kSCvtGroups = 2;
constexpr int kSCvtPer = (kSPatchPix + 127) / 128;   // pixels per converter thread and unit (5)
constexpr uint32_t kSRawSlot = 7424;           // 3 x 14 x 44 fp32 = 7392 B, padded to 128
// TMA wants the first element of a box row 16-byte aligned in global memory, and here the innermost dimension is x: the
// raw box starts at the patch's first column rounded down to 4 (fp32) / 8 (fp16) pixels and is 44 / 48 columns wide; the
// converters skip the 0..3 (0..7) extra columns.
template <typename TIn> struct RawGeom;
template <> struct RawGeom<float> { static constexpr int kPitch = 44, kMask = 3; };
template <> struct RawGeom<__half> { static constexpr int kPitch = 48, kMask = 7; };
constexpr uint32_t kSCvtSlot = kSCvtPix * 32;  // 18432: 32 bytes per pixel (one K = 16 row), 32B swizzle
constexpr uint32_t kSStgSlot = 58 * 512;       // 4 x 114 compact rows x 32 fp16 = 29184 B, padded to the 512-byte swizzle period
constexpr uint32_t kSOffW = 2048, kSOffStg = 8192;
constexpr uint32_t kSOffCvt = kSOffStg + kSRing * kSStgSlot;
constexpr uint32_t kSOffRaw = kSOffCvt + kSCvtStages * kSCvtSlot;
constexpr uint32_t kSSmem = kSOffRaw + kSRawStages * kSRawSlot + 1024;
static_assert(kSOffRaw % 128 == 0, "TMA destination alignment");
static_assert(kSU * kSTileValid * 64 <= kSStgSlot && 3 * kSPatchRows * 44 * 4 <= kSRawSlot, "slot sizes");
static_assert(kSSmem <= 227 * 1024, "shared memory budget");

// tap pairs of the five instructions: first tap's patch offset (pixels), distance to the second tap (pixels), tap indices
// (-1: no tap, zero weights)
constexpr int kSMma = 3;                       // instructions per M-tile: one per filter row

// K-major operand with 32-byte rows and the 32B swizzle (layout type 6): SBO = 8 rows = 256 bytes
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}

struct StemHaloArgs {
    int tiles_x, tiles_y, total_tiles;
    float sb[64];                               // scale[32] | bias[32] BY VALUE: the epilogue reads them as constant-bank operands
    const __half* w;                            // [32][32] fp16, k = (ky*3+kx)*3 + c, zero padded
    int* dbg;
};



// tile = (img * tiles_y + ty) * tiles_x + tx, walked with a constant step without divisions
struct StemWalk {
    int tx, ty, img, dx, dy, dimg, tiles_x, tiles_y;
    __device__ __forceinline__ StemWalk(const StemHaloArgs& a, int first, int step) : tiles_x(a.tiles_x), tiles_y(a.tiles_y) {
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
    __device__ __forceinline__ int x0() const { return tx * kSC; }
    __device__ __forceinline__ int y0() const { return ty * kSUR; }
};


__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ra)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&ra);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ra)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&ra);
}


template <typename TIn>
__global__ void __launch_bounds__(kSThreads, 1)
stem_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmOut, const StemHaloArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    // header: rfull[4] | rempty[4] | cfull[4] | cempty[4] | tfull[4] | tempty[4] | sempty[4] | sready[4] | tmem_ptr
    const uint32_t rfull0 = base, rempty0 = base + 32, cfull0 = base + 64, cempty0 = base + 96;
    const uint32_t tfull0 = base + 128, tempty0 = base + 160, sempty0 = base + 192, sready0 = base + 224;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 256);
    const uint32_t wsm = base + kSOffW, stg0 = base + kSOffStg, cvt0 = base + kSOffCvt, raw0 = base + kSOffRaw;
    constexpr int RP = RawGeom<TIn>::kPitch, RMASK = RawGeom<TIn>::kMask;
    constexpr int RPLANE = kSPatchRows * RP;
    constexpr uint32_t RAW_BYTES = 3u * RPLANE * sizeof(TIn);

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
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 32 * kSU * kSAccStages);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 320) {
        // weight table: filter row ky, output channel n -> one 32-byte K-major row (32B swizzle) = the nine weights of taps
        // (ky,0), (ky,1), (ky,2) x 3 channels, + 7 zeros; idx also enumerates the 16-byte chunk h of the row
        const int idx = threadIdx.x - 128;
        const int ky = idx >> 6, h = (idx >> 5) & 1, n = idx & 31;
        if (ky < 3) {
            const unsigned short* wr = reinterpret_cast<const unsigned short*>(a.w) + n * 32 + ky * 9;   // k = (ky*3+kx)*3 + c
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (h == 0) {
                v.x = (uint32_t)__ldg(wr) | ((uint32_t)__ldg(wr + 1) << 16);
                v.y = (uint32_t)__ldg(wr + 2) | ((uint32_t)__ldg(wr + 3) << 16);
                v.z = (uint32_t)__ldg(wr + 4) | ((uint32_t)__ldg(wr + 5) << 16);
                v.w = (uint32_t)__ldg(wr + 6) | ((uint32_t)__ldg(wr + 7) << 16);
            } else {
                v.x = (uint32_t)__ldg(wr + 8);
            }
            *reinterpret_cast<uint4*>(gen + kSOffW + ky * 1024 + n * 32 + ((h ^ ((n >> 2) & 1)) << 4)) = v;
        }
    }
    // the converted stages start as zeros: rows 560..575 are never written and only feed by-product GEMM rows
    for (uint32_t i = threadIdx.x; i < kSCvtStages * kSCvtSlot / 16; i += kSThreads)
        *reinterpret_cast<uint4*>(gen + kSOffCvt + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait_prior();             // the output buffer may still be read by the previous step's kernels

    if (warp == 0) {
        // ===== TMA producer: one planar patch per unit =====
        Slot<kSRawStages> rs(0);
        StemWalk t(a, unit_first, unit_step);
        for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, t.next(), rs.advance(1)) {
            mbar_wait(rempty0 + 8 * rs.i, rs.ph ^ 1, a.dbg, 0, (int)rs.i);
            if (elect_one()) {
                mbar_arrive_expect_tx(rfull0 + 8 * rs.i, RAW_BYTES);
                tma_load_4d(&tmX, raw0 + rs.i * kSRawSlot, rfull0 + 8 * rs.i, (t.x0() - 1) & ~RMASK, t.y0() - 1, 0, t.img);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers (two warps, alternate units): per unit 4 M-tiles x five instructions of two taps each; M-tile t
        // reads the patch from pixel t*120 on and accumulates into its own 32 TMEM columns; one commit pair per unit =====
        const int k = warp >> 1;
        const uint32_t idesc = make_idesc(32);
        // instruction ky: the 32-byte rows of pixels m + ky*40 .. (each holds the pixel and its two right neighbours = taps kx = 0, 1, 2)
        uint64_t ad[kSMma], bd[kSMma];
#pragma unroll
        for (int j = 0; j < kSMma; ++j) {
            ad[j] = make_desc_sw32(cvt0) + (uint64_t)(2 * j * kSP);
            bd[j] = make_desc_sw32(wsm + j * 1024);
        }
        Slot<kSCvtStages> cs((uint32_t)k);
        Slot<kSAccStages> as((uint32_t)k);
        for (int unit = unit_first + k * unit_step; unit < a.total_tiles; unit += 2 * unit_step, cs.advance(2), as.advance(2)) {
            mbar_wait(tempty0 + 8 * as.i, as.ph ^ 1, a.dbg, 1, 100 + (int)as.i);
            mbar_wait(cfull0 + 8 * cs.i, cs.ph, a.dbg, 1, (int)cs.i);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t a_off = (uint64_t)(cs.i * (kSCvtSlot >> 4));
                const uint32_t d_tmem = tmem_base + as.i * (32 * kSU);
                // filter-row-major order: consecutive instructions accumulate into different M-tiles (back-to-back instructions on
                // one accumulator serialise on its read-modify-write, ~150 cycles each for these short N = 32 instructions)
#pragma unroll
                for (int j = 0; j < kSMma; ++j) {
#pragma unroll
                    for (int t = 0; t < kSU; ++t)
                        umma_f16(d_tmem + t * 32, ad[j] + a_off + (uint64_t)(2 * t * kSR * kSP), bd[j], idesc, j != 0);
                }
                umma_commit(cempty0 + 8 * cs.i);
                umma_commit(tfull0 + 8 * as.i);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ===== store issuer: one {32 ch, 38, 12, 1} box per unit =====
        if (lane == 0) {
            StemWalk t(a, unit_first, unit_step);
            Slot<kSRing> ss(0);
            uint32_t prev = 0;
            bool first = true;
            for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, t.next(), ss.advance(1)) {
                mbar_wait(sready0 + 8 * ss.i, ss.ph, a.dbg, 4, 700 + (int)ss.i);
                tma_store_4d(&tmOut, stg0 + ss.i * kSStgSlot, 0, t.x0(), t.y0(), t.img);
                tma_store_commit();
                if (!first) {                                 // one store (29 KB) may stay unread; the one before goes back
                    tma_store_wait_read<1>();
                    mbar_arrive(sempty0 + 8 * prev);
                }
                first = false;
                prev = ss.i;
            }
            tma_store_wait_all();
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + 4 * kSCvtGroups) {
        // ===== converters (two groups of four warps, alternate units): planar fp32 / fp16 patch -> pixel-major fp16, 16 bytes
        // per pixel; a thread converts up to five pixels per unit =====
        const int grp = (warp - 4) >> 2;
        const int tid = threadIdx.x & 127;
        int q[kSCvtPer];
        int room[kSCvtPer];                                   // right neighbours inside the patch row (0..2)
#pragma unroll
        for (int i = 0; i < kSCvtPer; ++i) {
            const int p = tid + 128 * i;
            q[i] = (p / kSP) * RP + p % kSP;                  // raw index of patch pixel p (+ the unit's column shift)
            room[i] = min(2, kSP - 1 - p % kSP);
        }
        const bool last_ok = tid + 128 * (kSCvtPer - 1) < kSPatchPix;
        Slot<kSRawStages> rs((uint32_t)grp);
        Slot<kSCvtStages> cs((uint32_t)grp);
        int tx = (unit_first + grp * unit_step) % a.tiles_x;
        const int dtx = (kSCvtGroups * unit_step) % a.tiles_x;
        for (int unit = unit_first + grp * unit_step; unit < a.total_tiles;
             unit += kSCvtGroups * unit_step, rs.advance(kSCvtGroups), cs.advance(kSCvtGroups)) {
            const int d = (tx * kSC - 1) & RMASK;
            tx += dtx; if (tx >= a.tiles_x) tx -= a.tiles_x;
            mbar_wait(rfull0 + 8 * rs.i, rs.ph, a.dbg, 5, (int)rs.i);
            const TIn* rp = reinterpret_cast<const TIn*>(gen + kSOffRaw + rs.i * kSRawSlot) + d;
            // 32-byte row of patch pixel p: its three channels, those of its two right neighbours in the patch row (zeros past
            // the row end: they only meet by-product GEMM rows), seven zeros (never written: the stages start zeroed)
            uint32_t pk[kSCvtPer][5];
#pragma unroll
            for (int i = 0; i < kSCvtPer; ++i) {
                float v[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) v[k] = 0.f;
                if (i < kSCvtPer - 1 || last_ok) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (k <= room[i]) {
                            v[3 * k] = raw_ld(rp + q[i] + k); v[3 * k + 1] = raw_ld(rp + RPLANE + q[i] + k); v[3 * k + 2] = raw_ld(rp + 2 * RPLANE + q[i] + k);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                    pk[i][k] = *reinterpret_cast<const uint32_t*>(&h);
                }
                const __half2 h8 = __floats2half2_rn(v[8], 0.f);
                pk[i][4] = *reinterpret_cast<const uint32_t*>(&h8);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(rempty0 + 8 * rs.i);  // the raw patch is in registers
            mbar_wait(cempty0 + 8 * cs.i, cs.ph ^ 1, a.dbg, 5, 100 + (int)cs.i);
            // row p = tid + 128 i at p*32; 16-byte chunk j at ((j ^ (p >> 2)) & 1) << 4 (32B swizzle; p >> 2 has the parity of tid >> 2)
            uint8_t* dst = gen + kSOffCvt + cs.i * kSCvtSlot + tid * 32;
            const uint32_t c0 = ((tid >> 2) & 1) << 4;
#pragma unroll
            for (int i = 0; i < kSCvtPer; ++i)
                if (i < kSCvtPer - 1 || last_ok) {
                    *reinterpret_cast<uint4*>(dst + i * 4096 + c0) = make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
                    *reinterpret_cast<uint32_t*>(dst + i * 4096 + (c0 ^ 16u)) = pk[i][4];
                }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(cfull0 + 8 * cs.i);
        }
    } else if (warp >= 12) {
        // ===== epilogue: sixteen warps per unit -- warp (q, t) drains TMEM lane quarter q of M-tile t.  TMEM lane m = GEMM row m
        // = patch-pitch pixel (r, c) of the M-tile; staging row = t*114 + r*38 + c (the unit's 12 x 38 box in row-major order).
        // A thread owns one pixel and all 32 channels; scale / bias are kernel parameters, i.e. constant-bank operands (read
        // from a shared-memory table -- sixteen broadcast LDS.128 per thread, four wavefronts each -- they kept the LSU pipe
        // 80 % busy: profiles/r04a_full_raw_stem_halo.csv) =====
        const int q = warp & 3, t = (warp - 12) >> 2;
        const int m = q * 32 + lane;
        const int r = m / kSP, c = m - r * kSP;
        const bool valid = r < kSR && c < kSC;
        const int mp = t * kSTileValid + r * kSC + c;
        const int xr = (mp >> 1) & 3;
        Slot<kSAccStages> as(0);
        Slot<kSRing> ss(0);
        for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, as.advance(1), ss.advance(1)) {
            mbar_wait(tfull0 + 8 * as.i, as.ph, a.dbg, 2, 200 + (int)as.i);
            tc_fence_after();
            uint32_t r0[16], r1[16];
            const uint32_t taddr = tmem_base + as.i * (32 * kSU) + t * 32 + ((uint32_t)(q * 32) << 16);
            tmem_ld16(taddr, r0);
            tmem_ld16(taddr + 16, r1);
            mbar_wait(sempty0 + 8 * ss.i, ss.ph ^ 1, a.dbg, 2, 500 + (int)ss.i);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as.i);  // accumulator drained into registers
            if (valid) {
                uint8_t* srow = gen + kSOffStg + ss.i * kSStgSlot + mp * 64;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {              // 16-byte chunk cc = channels 8cc .. 8cc+7
                    const uint32_t* rr = cc < 2 ? r0 : r1;
                    const int j0 = 8 * (cc & 1);
                    uint4 pk;
                    __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ch = 8 * cc + 2 * e;
                        const float y0 = fmaf(__uint_as_float(rr[j0 + 2 * e]), a.sb[ch], a.sb[32 + ch]);
                        const float y1 = fmaf(__uint_as_float(rr[j0 + 2 * e + 1]), a.sb[ch + 1], a.sb[32 + ch + 1]);
                        ph2[e] = __floats2half2_rn(fmaxf(y0, y0 * kLeaky), fmaxf(y1, y1 * kLeaky));     // LeakyReLU(0.1) = max(v, 0.1 v)
                    }
                    *reinterpret_cast<uint4*>(srow + ((cc ^ xr) << 4)) = pk;
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
        tmem_dealloc(tmem_base, 32 * kSU * kSAccStages);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;

std::string enc_err(const char* what, CUresult r) { return std::string(what) + " failed with CUresult " + std::to_string((int)r); }

std::string load_encoder() {
    if (g_enc) return "";
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeTiled not available from the driver";
    g_enc = reinterpret_cast<EncodeTiledFn>(f);
    return "";
}

}  // namespace

bool stem_halo_supported(int W, int in_f16) { return W % (in_f16 ? 8 : 4) == 0; }

std::string stem_halo_make_plan(StemHaloPlan& p, __half* out, long out_ld, int B, int H, int W, int num_sms) {
    std::string e = load_encoder();
    if (!e.empty()) return e;
    if (out_ld % 8) return "stem output pitch must be a multiple of 8 channels";
    p.tiles_x = (W + kSC - 1) / kSC;
    p.tiles_y = (H + kSUR - 1) / kSUR;
    p.total_tiles = B * p.tiles_x * p.tiles_y;
    p.grid = std::min(p.total_tiles, num_sms);
    p.smem = kSSmem;
    const cuuint32_t es4[4] = {1, 1, 1, 1};
    cuuint64_t dims[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)out_ld * 2, (cuuint64_t)W * out_ld * 2, (cuuint64_t)H * W * out_ld * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)kSC, (cuuint32_t)kSUR, 1};
    CUresult r = g_enc(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out, dims, st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return enc_err("cuTensorMapEncodeTiled(stem output)", r);
    return "";
}

cudaError_t stem_halo_launch(const StemHaloPlan& p, const void* x, int in_f16, int B, int H, int W, const __half* w16,
                             const float* sb_host, int* dbg, cudaStream_t s) {
    // the image is the caller's tensor: its tensor map is encoded per call (host-side, ~1 us)
    CUtensorMap tmX;
    {
        const size_t es = in_f16 ? 2 : 4;
        const cuuint32_t es4[4] = {1, 1, 1, 1};
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        cuuint64_t st[3] = {(cuuint64_t)W * es, (cuuint64_t)H * W * es, (cuuint64_t)3 * H * W * es};
        cuuint32_t box[4] = {(cuuint32_t)(in_f16 ? RawGeom<__half>::kPitch : RawGeom<float>::kPitch), (cuuint32_t)kSPatchRows, 3, 1};
        CUresult r = g_enc(&tmX, in_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims,
                           st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    static PerDeviceOnce attr_once;
    {
        cudaError_t e = attr_once.run([] {
            cudaError_t r = cudaFuncSetAttribute(stem_halo_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSSmem);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(stem_halo_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSSmem);
            return r;
        });
        if (e != cudaSuccess) return e;
    }
    StemHaloArgs a;
    a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.total_tiles = p.total_tiles;
    for (int i = 0; i < 64; ++i) a.sb[i] = sb_host[i];
    a.w = w16; a.dbg = dbg;
    static const bool pdl = !(tune_env("YB_TC_PDL") && atoi(tune_env("YB_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(kSThreads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = in_f16 ? cudaLaunchKernelEx(&cfg, stem_halo_kernel<__half>, tmX, p.tmOut, a)
                                 : cudaLaunchKernelEx(&cfg, stem_halo_kernel<float>, tmX, p.tmOut, a);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace yb
