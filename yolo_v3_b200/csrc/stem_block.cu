// K0f: the first TWO layers in one kernel -- stem (3 -> 32, 3x3, stride 1) and the first down-sampling convolution
// (32 -> 64, 3x3, stride 2) -- with the 608 x 608 x 32 stem output kept ON CHIP.
//
// As separate kernels the pair moves 142 MB (image) + 757 MB (stem output written) + 757 MB (read back) + 378 MB (layer-1
// output) per 32-image batch: both run at their HBM bound, 0.157 + 0.175 ms.  Fused, only the image is read and the
// 304 x 304 x 64 tensor written (520 MB = 0.08 ms of HBM time); the bound moves on chip (the per-unit chains of the epilogue
// warps and the short N = 32 / 64 tensor-core instructions: 0.293 ms).
//
//   unit      = one layer-1 output tile of conv_halo.cu's stride-2 geometry: 3 rows x 38 columns (GEMM row m = r*40 + c).
//   needs     = stem outputs rows 2*y0-1 .. 2*y0+5, columns 2*x0-1 .. 2*x0+75 (7 x 77), i.e. the image patch rows
//               2*y0-2 .. 2*y0+6, columns 2*x0-2 .. 2*x0+76 (9 x 79; pitch 80), zero-filled outside the image by TMA.
//   stem      = stem_halo.cu's scheme on that patch: planar TMA load -> converter warps -> pixel-major fp16 (16 B per pixel:
//               its three channels, its right neighbour's three, two zeros) -> non-swizzled K-major descriptors, one
//               tcgen05.mma per filter row (LBO = 2 pixels), five M-tiles of 128 patch-pitch pixels (L = Y*80 + X), each
//               with its own 32 TMEM columns.
//   hand-over = the epilogue warps drain the stem accumulators (scale / bias / LeakyReLU, ZERO where the stem pixel lies
//               outside the image: that is layer 1's padding), and write fp16 rows of 64 bytes into four PARITY PLANES
//               (odd / even stem rows x odd / even stem columns, 4 x 40 pixels each, 64B swizzle) -- exactly what
//               conv_halo.cu's stride-2 kernel loads from HBM with four strided TMA boxes.
//   layer 1   = nine taps = nine unit-pitch views of the planes (descriptor start shifted by whole pixels), two K = 16
//               instructions each, M = 128 x N = 64; resident weights (36 KB); epilogue -> 128B-swizzled staging -> TMA store
//               of a {64 ch, 38, 3, 1} box.
//
// Two units are in flight: the epilogue warps form two groups of eight that take alternate units (group g works in stage g
// of everything: stem accumulator set, parity planes, layer-1 accumulator, staging slot), and the MMA issuer runs layer 1 of
// unit u as soon as its planes are written and the stem of unit u+2 as soon as the hand-over of u has drained that
// accumulator set.
//
// reference: darknet.py:37-44 (conv_bn_relu), :66-69 (Darknet.__init__: conv 3->32, then the stride-2 conv of stage 0).
#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "yb_internal.h"

namespace yb {
namespace {

constexpr int kFP = 40, kFC = 38, kFR = 3;        // layer-1 tile: patch pitch, output columns, output rows
constexpr int kFPlanePix = 4 * kFP;               // 160 pixels per parity plane
constexpr uint32_t kFPlaneBytes = kFPlanePix * 64;          // 10240
constexpr uint32_t kFPlaneSlot = 4 * kFPlaneBytes + 48 * 64;   // 44032: the last view reads plane 3 + (1*40+1) + 127 rows
constexpr int kFIP = 80;                          // image-patch pitch (stem GEMM row L = Y*80 + X)
constexpr int kFIRows = 9;                        // image-patch rows
constexpr int kFIPix = kFIRows * kFIP;            // 720
constexpr int kFSY = 7, kFSX = 77;                // stem outputs needed per unit
constexpr int kFMT = 5;                           // stem M-tiles per unit (7 * 80 = 560 GEMM rows)
constexpr int kFCvtPix = 816;                     // M-tile 4, filter row 2, K half 1, + neighbour: 512 + 160 + 2 + 127 = 801
constexpr uint32_t kFCvtSlot = kFCvtPix * 16;     // 13056
constexpr uint32_t kFRawSlot = 9088;              // 3 x 9 x 84 fp32 = 9072 B, padded to 128
constexpr int kFRawStages = 3, kFCvtStages = 2, kFPlaneStages = 2, kFRing = 2;
constexpr int kFThreads = 768;                    // warp 0 TMA, 1 store issuer, 2 MMA, 3 TMEM alloc, 4-7 converters, 8-23 epilogue
constexpr int kFCvtPer = (kFIPix + 127) / 128;    // 6 pixels per converter thread and unit
constexpr uint32_t kFStgSlot = 128 * 128;         // layer-1 staging: 114 compact rows x 64 fp16
constexpr uint32_t kFWBytes = 9 * 64 * 64;        // resident layer-1 weights: 9 taps x 64 rows x 32 fp16
constexpr uint32_t kFOffTab = 1024;               // layer-1 scale[64] | bias[64]
constexpr uint32_t kFOffW0 = 2048;                // stem weight table (3 KB)
constexpr uint32_t kFOffW1 = 5120;                // layer-1 weights (1024-aligned: 64B-swizzled TMA boxes)
constexpr uint32_t kFOffStg = kFOffW1 + kFWBytes;                     // 41984 (1024-aligned)
constexpr uint32_t kFOffPlane = kFOffStg + kFRing * kFStgSlot;        // 74752
constexpr uint32_t kFOffCvt = kFOffPlane + kFPlaneStages * kFPlaneSlot;   // 162816
constexpr uint32_t kFOffRaw = kFOffCvt + kFCvtStages * kFCvtSlot;     // 188928
constexpr uint32_t kFSmem = kFOffRaw + kFRawStages * kFRawSlot + 1024;
static_assert(kFOffW1 % 1024 == 0 && kFOffStg % 1024 == 0 && kFOffPlane % 1024 == 0 && kFPlaneSlot % 1024 == 0, "swizzle alignment");
static_assert(kFOffRaw % 128 == 0 && kFRawSlot % 128 == 0, "TMA destination alignment");
static_assert(kFSmem <= 227 * 1024, "shared memory budget");
constexpr uint32_t kFAcc1Cols = 32 * kFMT;        // 160 TMEM columns per stem accumulator set
constexpr uint32_t kFAcc2Base = 2 * kFAcc1Cols;   // layer-1 accumulators: 2 x 64 columns from column 320

__device__ __forceinline__ float2 ffma2x(float2 a, float2 b, float2 c) {
    unsigned long long ra;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ra)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&ra);
}
__device__ __forceinline__ float2 fmul2x(float2 a, float2 b) {
    unsigned long long ra;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ra)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&ra);
}

template <typename TIn> struct FRaw;
template <> struct FRaw<float> { static constexpr int kPitch = 84, kMask = 3; };
template <> struct FRaw<__half> { static constexpr int kPitch = 88, kMask = 7; };

struct StemBlockArgs {
    int tiles_x, tiles_y, total_tiles;
    int H, W;                                     // image = stem output size
    float sb0[64];                                // stem scale[32] | bias[32], by value (constant-bank operands)
    const __half* w0;                             // stem weights [32][32] fp16, k = (ky*3+kx)*3 + c
    const float* scale1; const float* bias1;      // layer 1, [64]
    int* dbg;
};

struct BlockWalk {
    int tx, ty, img, dx, dy, dimg, tiles_x, tiles_y;
    __device__ __forceinline__ BlockWalk(const StemBlockArgs& a, int first, int step) : tiles_x(a.tiles_x), tiles_y(a.tiles_y) {
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
    __device__ __forceinline__ int x0() const { return tx * kFC; }    // layer-1 output coordinates
    __device__ __forceinline__ int y0() const { return ty * kFR; }
};

template <typename TIn>
__global__ void __launch_bounds__(kFThreads, 1)
stem_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmOut, const StemBlockArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    // header (8-byte barriers): rfull[3] rempty[3] | cfull[2] cempty[2] | t1full[2] t1empty[2] | pfull[2] pempty[2] |
    //                           t2full[2] t2empty[2] | sready[2] sempty[2] | wbar | tmem_ptr
    const uint32_t rfull0 = base, rempty0 = base + 24, cfull0 = base + 48, cempty0 = base + 64;
    const uint32_t t1full0 = base + 80, t1empty0 = base + 96, pfull0 = base + 112, pempty0 = base + 128;
    const uint32_t t2full0 = base + 144, t2empty0 = base + 160, sready0 = base + 176, sempty0 = base + 192, wbar = base + 208;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 224);
    const uint32_t w0sm = base + kFOffW0, w1sm = base + kFOffW1, stg0 = base + kFOffStg, plane0 = base + kFOffPlane;
    const uint32_t cvt0 = base + kFOffCvt, raw0 = base + kFOffRaw;
    constexpr int RP = FRaw<TIn>::kPitch, RMASK = FRaw<TIn>::kMask;
    constexpr int RPLANE = kFIRows * RP;
    constexpr uint32_t RAW_BYTES = 3u * RPLANE * sizeof(TIn);
    static_assert(RAW_BYTES <= kFRawSlot, "raw slot");

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int unit_first = blockIdx.x, unit_step = gridDim.x;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmOut); }
    if (warp == 2 && lane == 0) {
        for (int s = 0; s < kFRawStages; ++s) { mbar_init(rfull0 + 8 * s, 1); mbar_init(rempty0 + 8 * s, 4); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(cfull0 + 8 * s, 4); mbar_init(cempty0 + 8 * s, 1);
            mbar_init(t1full0 + 8 * s, 1); mbar_init(t1empty0 + 8 * s, 8);      // eight epilogue warps per unit (see below)
            mbar_init(pfull0 + 8 * s, 8); mbar_init(pempty0 + 8 * s, 1);
            mbar_init(t2full0 + 8 * s, 1); mbar_init(t2empty0 + 8 * s, 8);
            mbar_init(sready0 + 8 * s, 8); mbar_init(sempty0 + 8 * s, 1);
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 3) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
    for (int i = threadIdx.x; i < 128; i += kFThreads) {
        float* tab = reinterpret_cast<float*>(gen + kFOffTab);
        tab[i] = i < 64 ? __ldg(a.scale1 + i) : __ldg(a.bias1 + i - 64);
    }
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 192) {
        // stem weight table: filter row ky, K half h, output channel n -> 8 fp16: h = 0: taps (ky,0), (ky,1) (3 channels each)
        // + 2 zeros; h = 1: tap (ky,2) + 5 zeros
        const int idx = threadIdx.x - 128;
        const int ky = idx >> 6, h = (idx >> 5) & 1, n = idx & 31;
        const unsigned short* wr = reinterpret_cast<const unsigned short*>(a.w0) + n * 32 + (ky * 3 + 2 * h) * 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        v.x = (uint32_t)__ldg(wr) | ((uint32_t)__ldg(wr + 1) << 16);
        v.y = (uint32_t)__ldg(wr + 2);
        if (h == 0) {
            v.y |= (uint32_t)__ldg(wr + 3) << 16;
            v.z = (uint32_t)__ldg(wr + 4) | ((uint32_t)__ldg(wr + 5) << 16);
        }
        *reinterpret_cast<uint4*>(gen + kFOffW0 + ky * 1024 + h * 512 + n * 16) = v;
    }
    // converted stages and parity planes start as zeros: the never-written rows only feed by-product GEMM rows, but they must
    // not hold NaN patterns next to zero weights
    for (uint32_t i = threadIdx.x; i < (kFPlaneStages * kFPlaneSlot + kFCvtStages * kFCvtSlot) / 16; i += kFThreads)
        *reinterpret_cast<uint4*>(gen + kFOffPlane + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait_prior();             // the output buffer may still be read by the previous step's kernels

    if (warp == 0) {
        // ===== TMA producer: the resident layer-1 weights once, then one planar image patch per unit =====
        if (unit_first < a.total_tiles && elect_one()) {
            mbar_arrive_expect_tx(wbar, kFWBytes);
#pragma unroll 1
            for (int t = 0; t < 9; ++t) tma_load_2d(&tmW1, w1sm + t * 4096, wbar, t * 32, 0);
        }
        __syncwarp();
        Slot<kFRawStages> rs(0);
        BlockWalk t(a, unit_first, unit_step);
        for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, t.next(), rs.advance(1)) {
            mbar_wait(rempty0 + 8 * rs.i, rs.ph ^ 1, a.dbg, 0, (int)rs.i);
            if (elect_one()) {
                mbar_arrive_expect_tx(rfull0 + 8 * rs.i, RAW_BYTES);
                tma_load_4d(&tmX, raw0 + rs.i * kFRawSlot, rfull0 + 8 * rs.i, (2 * t.x0() - 2) & ~RMASK, 2 * t.y0() - 2, 0, t.img);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ===== MMA issuer: stem(0); then per unit u: stem(u+1), layer1(u) =====
        const uint32_t idesc0 = make_idesc(32), idesc1 = make_idesc(64);
        uint64_t ad0[3], bd0[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            ad0[j] = make_desc_plain(cvt0, 2, 8) + (uint64_t)(j * kFIP);       // filter row j; K half 1 = two pixels on
            bd0[j] = make_desc_plain(w0sm + j * 1024, 32, 8);
        }
        const uint64_t bd1 = make_smem_desc<64>(w1sm);
        const uint64_t ad1 = make_smem_desc<64>(plane0);
        auto stem = [&](uint32_t s) {          // s = u & 1: converted stage and stem accumulator set of unit u
            if (elect_one()) {
                const uint64_t a_off = (uint64_t)(s * (kFCvtSlot >> 4));
                const uint32_t d_tmem = tmem_base + s * kFAcc1Cols;
                // filter-row-major order: consecutive instructions accumulate into different M-tiles
#pragma unroll
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                    for (int t = 0; t < kFMT; ++t) umma_f16(d_tmem + t * 32, ad0[j] + a_off + (uint64_t)(t * 128), bd0[j], idesc0, j != 0);
                }
                umma_commit(cempty0 + 8 * s);
                umma_commit(t1full0 + 8 * s);
            }
            __syncwarp();
        };
        const int n_units = unit_first < a.total_tiles ? (a.total_tiles - unit_first + unit_step - 1) / unit_step : 0;
        // stem(0), stem(1); then per unit u: layer1(u) as soon as its planes are written, stem(u+2) as soon as the hand-over
        // of u has drained that accumulator set (the two halves of the epilogue warps work on alternate units)
        if (n_units > 0) {
            mbar_wait(cfull0, 0, a.dbg, 1, 0);
            tc_fence_after();
            stem(0);
            mbar_wait(wbar, 0, a.dbg, 1, 600);
        }
        if (n_units > 1) {
            mbar_wait(cfull0 + 8, 0, a.dbg, 1, 1);
            tc_fence_after();
            stem(1);
        }
        for (int u = 0; u < n_units; ++u) {
            const uint32_t s = (uint32_t)u & 1u, ph = ((uint32_t)u >> 1) & 1u;
            mbar_wait(t2empty0 + 8 * s, ph ^ 1, a.dbg, 1, 200 + (int)s);
            mbar_wait(pfull0 + 8 * s, ph, a.dbg, 1, 300 + (int)s);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_tmem = tmem_base + kFAcc2Base + s * 64;
                const uint64_t a_base = ad1 + (uint64_t)(s * (kFPlaneSlot >> 4));
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int ky = t / 3, kx = t % 3;
                    // plane (ky odd?, kx odd?) at pixel offset (ky/2, kx/2)
                    const int pix = (((ky & 1) << 1) | (kx & 1)) * kFPlanePix + (ky >> 1) * kFP + (kx >> 1);
                    const uint64_t ad = a_base + (uint64_t)((pix * 64) >> 4);
                    const uint64_t bd = bd1 + (uint64_t)(t * (4096 >> 4));
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc1, (t | k) != 0);
                }
                umma_commit(pempty0 + 8 * s);
                umma_commit(t2full0 + 8 * s);
            }
            __syncwarp();
            if (u + 2 < n_units) {
                const uint32_t ph2 = ((uint32_t)(u + 2) >> 1) & 1u;
                mbar_wait(t1empty0 + 8 * s, ph2 ^ 1, a.dbg, 1, 100 + (int)s);
                mbar_wait(cfull0 + 8 * s, ph2, a.dbg, 1, (int)s);
                tc_fence_after();
                stem(s);
            }
        }
    } else if (warp == 1) {
        // ===== store issuer =====
        if (lane == 0) {
            BlockWalk t(a, unit_first, unit_step);
            Slot<kFRing> ss(0);
            uint32_t prev = 0;
            bool first = true;
            for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, t.next(), ss.advance(1)) {
                mbar_wait(sready0 + 8 * ss.i, ss.ph, a.dbg, 4, 700 + (int)ss.i);
                tma_store_4d(&tmOut, stg0 + ss.i * kFStgSlot, 0, t.x0(), t.y0(), t.img);
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
    } else if (warp >= 4 && warp < 8) {
        // ===== converters: planar fp32 / fp16 patch (9 x 80 pixels) -> pixel-major fp16, 16 bytes per pixel =====
        const int tid = threadIdx.x & 127;
        int q[kFCvtPer];
        bool edge[kFCvtPer];
#pragma unroll
        for (int i = 0; i < kFCvtPer; ++i) {
            const int p = tid + 128 * i;
            q[i] = (p / kFIP) * RP + p % kFIP;
            edge[i] = p % kFIP == kFIP - 1;
        }
        const bool last_ok = tid + 128 * (kFCvtPer - 1) < kFIPix;
        Slot<kFRawStages> rs(0);
        Slot<kFCvtStages> cs(0);
        int tx = unit_first % a.tiles_x;
        const int dtx = unit_step % a.tiles_x;
        for (int unit = unit_first; unit < a.total_tiles; unit += unit_step, rs.advance(1), cs.advance(1)) {
            const int d = (2 * tx * kFC - 2) & RMASK;
            tx += dtx; if (tx >= a.tiles_x) tx -= a.tiles_x;
            mbar_wait(rfull0 + 8 * rs.i, rs.ph, a.dbg, 5, (int)rs.i);
            const TIn* rp = reinterpret_cast<const TIn*>(gen + kFOffRaw + rs.i * kFRawSlot) + d;
            uint32_t pk[kFCvtPer][3];
#pragma unroll
            for (int i = 0; i < kFCvtPer; ++i) {
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
                if (i < kFCvtPer - 1 || last_ok) {
                    v0 = raw_ld(rp + q[i]); v1 = raw_ld(rp + RPLANE + q[i]); v2 = raw_ld(rp + 2 * RPLANE + q[i]);
                    if (!edge[i]) { n0 = raw_ld(rp + q[i] + 1); n1 = raw_ld(rp + RPLANE + q[i] + 1); n2 = raw_ld(rp + 2 * RPLANE + q[i] + 1); }
                }
                const __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, n0), h2 = __floats2half2_rn(n1, n2);
                pk[i][0] = *reinterpret_cast<const uint32_t*>(&h0); pk[i][1] = *reinterpret_cast<const uint32_t*>(&h1);
                pk[i][2] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(rempty0 + 8 * rs.i);  // the raw patch is in registers
            mbar_wait(cempty0 + 8 * cs.i, cs.ph ^ 1, a.dbg, 5, 100 + (int)cs.i);
            uint8_t* dst = gen + kFOffCvt + cs.i * kFCvtSlot + tid * 16;
#pragma unroll
            for (int i = 0; i < kFCvtPer; ++i)
                if (i < kFCvtPer - 1 || last_ok) *reinterpret_cast<uint4*>(dst + i * 2048) = make_uint4(pk[i][0], pk[i][1], pk[i][2], 0u);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(cfull0 + 8 * cs.i);
        }
    } else if (warp >= 8) {
        // ===== epilogue warps.  TWO independent groups of eight warps (q = TMEM lane quarter, half = column half) take
        // alternate units: group g always works in stage g of everything (stem accumulator set, parity planes, layer-1
        // accumulator, staging slot).  Per unit: hand-over (stem accumulators -> parity planes), then -- once the MMA issuer
        // has run layer 1 from those planes -- the layer-1 epilogue.  With one group of sixteen warps the dependent chain of a
        // unit (five tcgen05.ld round trips, proxy fences, waiting for the layer-1 MMAs) paced the kernel at ~3 800 cycles per
        // unit while the tensor pipe idled (profiles/r04e_stem_block_timeline.txt); two chains overlap each other's waits =====
        const int q = warp & 3, idx = (warp - 8) >> 2;
        const int half = idx & 1;
        const uint32_t g = (uint32_t)(idx >> 1);
        const int ml = q * 32 + lane;
        // hand-over: this thread's stem pixel of M-tile t is L = 128 t + ml = (Y, X); channels 16*half .. 16*half+15 = 16-byte
        // chunks 2*half, 2*half+1 of the pixel's 64-byte plane row (64B swizzle: chunk c at (c ^ (pix >> 1 & 3)) << 4)
        int hoff[kFMT];            // byte offset of the pixel's row inside a plane slot | swizzle phase in bits 0-1, or -1
#pragma unroll
        for (int t = 0; t < kFMT; ++t) {
            const int L = 128 * t + ml, Y = L / kFIP, X = L - Y * kFIP;
            const int pl = ((Y & 1) << 1) | (X & 1), pix = (Y >> 1) * kFP + (X >> 1);
            hoff[t] = (Y < kFSY && X < kFSX) ? (int)(pl * kFPlaneBytes + pix * 64 + ((pix >> 1) & 3)) : -1;
        }
        // layer-1 epilogue: GEMM row ml = (r, c), staging row r*38 + c, channels 32*half .. 32*half+31
        const int r1 = ml / kFP, c1 = ml - r1 * kFP;
        const bool valid1 = r1 < kFR && c1 < kFC;
        const int mp = r1 * kFC + c1;
        const int xr = mp & 7;
        const float4* sc1 = reinterpret_cast<const float4*>(gen + kFOffTab) + half * 8;          // layer-1 scale / bias of this thread's
        const float4* bi1 = reinterpret_cast<const float4*>(gen + kFOffTab + 256) + half * 8;    // 32 channels (shared-memory table)
        float2 sc0[8], bi0[8];                                                                   // stem scale / bias of its 16 channels
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            sc0[e] = make_float2(a.sb0[half * 16 + 2 * e], a.sb0[half * 16 + 2 * e + 1]);
            bi0[e] = make_float2(a.sb0[32 + half * 16 + 2 * e], a.sb0[32 + half * 16 + 2 * e + 1]);
        }
        const int n_units = unit_first < a.total_tiles ? (a.total_tiles - unit_first + unit_step - 1) / unit_step : 0;
        BlockWalk tw(a, unit_first + (int)g * unit_step, 2 * unit_step);      // position of this group's current unit
        uint8_t* pbase = gen + kFOffPlane + g * kFPlaneSlot;
        uint8_t* sbase = gen + kFOffStg + g * kFStgSlot;
        uint32_t ph = 0;
        for (int u = (int)g; u < n_units; u += 2, ph ^= 1u, tw.next()) {
            // ---- hand-over of unit u: stem output (gy, gx) = (2*y0 - 1 + Y, 2*x0 - 1 + X); outside the image -> zero (layer
            // 1's padding).  Only units on the image border can hold such pixels.
            const int gy0 = 2 * tw.y0() - 1, gx0 = 2 * tw.x0() - 1;
            const bool interior = gy0 >= 0 && gy0 + kFSY <= a.H && gx0 >= 0 && gx0 + kFSX <= a.W;
            mbar_wait(t1full0 + 8 * g, ph, a.dbg, 2, 200 + (int)g);
            tc_fence_after();
            mbar_wait(pempty0 + 8 * g, ph ^ 1, a.dbg, 2, 300 + (int)g);
#pragma unroll
            for (int t = 0; t < kFMT; ++t) {
                uint32_t v[16];
                tmem_ld16(tmem_base + g * kFAcc1Cols + (uint32_t)(t * 32 + half * 16) + ((uint32_t)(q * 32) << 16), v);
                tmem_ld_wait();
                if (hoff[t] >= 0) {
                    bool in_img = true;
                    if (!interior) {
                        const int L = 128 * t + ml, Y = L / kFIP;
                        const int gy = gy0 + Y, gx = gx0 + (L - Y * kFIP);
                        in_img = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                    }
                    uint8_t* prow = pbase + (hoff[t] & ~3);
                    const int sw = hoff[t] & 3;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 y = ffma2x(make_float2(__uint_as_float(v[8 * h + 2 * e]), __uint_as_float(v[8 * h + 2 * e + 1])),
                                                    sc0[4 * h + e], bi0[4 * h + e]);
                            const float2 z = fmul2x(y, make_float2(kLeaky, kLeaky));
                            ph2[e] = __floats2half2_rn(fmaxf(y.x, z.x), fmaxf(y.y, z.y));      // LeakyReLU(0.1) = max(v, 0.1 v)
                        }
                        if (!in_img) pk = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(prow + (((2 * half + h) ^ sw) << 4)) = pk;
                    }
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(t1empty0 + 8 * g); mbar_arrive(pfull0 + 8 * g); }
            // ---- layer-1 epilogue of unit u
            mbar_wait(t2full0 + 8 * g, ph, a.dbg, 2, 400 + (int)g);
            tc_fence_after();
            uint32_t r0[16], r1v[16];
            const uint32_t tcol = tmem_base + kFAcc2Base + g * 64 + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16);
            tmem_ld16(tcol, r0);
            tmem_ld16(tcol + 16, r1v);
            mbar_wait(sempty0 + 8 * g, ph ^ 1, a.dbg, 2, 500 + (int)g);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t2empty0 + 8 * g);    // accumulator drained into registers
            if (valid1) {
                uint8_t* srow = sbase + (uint32_t)mp * 128u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {                 // 16-byte chunk 4*half + c = channels 32*half + 8c .. + 7
                    const uint32_t* rr = c < 2 ? r0 : r1v;
                    const int j0 = 8 * (c & 1);
                    const float4 s0 = sc1[2 * c], s1 = sc1[2 * c + 1], b0 = bi1[2 * c], b1 = bi1[2 * c + 1];
                    uint4 pk;
                    __half2* ph2 = reinterpret_cast<__half2*>(&pk);
                    ph2[0] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j0 + 0]), s0.x, b0.x)), leaky(fmaf(__uint_as_float(rr[j0 + 1]), s0.y, b0.y)));
                    ph2[1] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j0 + 2]), s0.z, b0.z)), leaky(fmaf(__uint_as_float(rr[j0 + 3]), s0.w, b0.w)));
                    ph2[2] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j0 + 4]), s1.x, b1.x)), leaky(fmaf(__uint_as_float(rr[j0 + 5]), s1.y, b1.y)));
                    ph2[3] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j0 + 6]), s1.z, b1.z)), leaky(fmaf(__uint_as_float(rr[j0 + 7]), s1.w, b1.w)));
                    *reinterpret_cast<uint4*>(srow + (((4 * half + c) ^ xr) << 4)) = pk;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sready0 + 8 * g);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;

std::string enc_err(const char* what, CUresult r) { return std::string(what) + " failed with CUresult " + std::to_string((int)r); }

}  // namespace

bool stem_block_supported(int H, int W, int in_f16) {
    if (H % 2 || W % (in_f16 ? 8 : 4)) return false;
    const int Wo = W / 2, tx = (Wo + kFC - 1) / kFC;
    return Wo * 5 >= tx * kFC * 4;              // at least 80 % of the tile columns are real outputs (as conv_halo.cu)
}

std::string stem_block_make_plan(StemBlockPlan& p, const __half* w1, __half* out, long out_ld, int B, int H, int W, int num_sms) {
    if (!g_enc) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeTiled not available from the driver";
        g_enc = reinterpret_cast<EncodeTiledFn>(f);
    }
    if (out_ld % 8) return "layer-1 output pitch must be a multiple of 8 channels";
    const int Ho = H / 2, Wo = W / 2;
    p.tiles_x = (Wo + kFC - 1) / kFC;
    p.tiles_y = (Ho + kFR - 1) / kFR;
    p.total_tiles = B * p.tiles_x * p.tiles_y;
    p.grid = std::min(p.total_tiles, num_sms);
    p.smem = kFSmem;
    const cuuint32_t es4[4] = {1, 1, 1, 1};
    const cuuint32_t es2[2] = {1, 1};
    {   // layer-1 weights [64][9*32] fp16, one tap x 64 rows per box, 64B swizzle
        cuuint64_t dims[2] = {288, 64};
        cuuint64_t st[1] = {288 * 2};
        cuuint32_t box[2] = {32, 64};
        CUresult r = g_enc(&p.tmW1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w1), dims, st, box, es2,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return enc_err("cuTensorMapEncodeTiled(layer-1 weights)", r);
    }
    {   // output: NHWC as (C, W, H, N); box = 64 ch x 38 x 3 x 1, 128B swizzle (staging rows are 128 bytes)
        cuuint64_t dims[4] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
        cuuint64_t st[3] = {(cuuint64_t)out_ld * 2, (cuuint64_t)Wo * out_ld * 2, (cuuint64_t)Ho * Wo * out_ld * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)kFC, (cuuint32_t)kFR, 1};
        CUresult r = g_enc(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out, dims, st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return enc_err("cuTensorMapEncodeTiled(layer-1 output)", r);
    }
    return "";
}

cudaError_t stem_block_launch(const StemBlockPlan& p, const void* x, int in_f16, int B, int H, int W, const __half* w0,
                              const float* sb0_host, const float* scale1, const float* bias1, int* dbg, cudaStream_t s) {
    // the image is the caller's tensor: its tensor map is encoded per call (host-side, ~1 us)
    CUtensorMap tmX;
    {
        const size_t es = in_f16 ? 2 : 4;
        const cuuint32_t es4[4] = {1, 1, 1, 1};
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        cuuint64_t st[3] = {(cuuint64_t)W * es, (cuuint64_t)H * W * es, (cuuint64_t)3 * H * W * es};
        cuuint32_t box[4] = {(cuuint32_t)(in_f16 ? FRaw<__half>::kPitch : FRaw<float>::kPitch), (cuuint32_t)kFIRows, 3, 1};
        CUresult r = g_enc(&tmX, in_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims,
                           st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    static PerDeviceOnce attr_once;
    {
        cudaError_t e = attr_once.run([] {
            cudaError_t r = cudaFuncSetAttribute(stem_block_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmem);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(stem_block_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmem);
            return r;
        });
        if (e != cudaSuccess) return e;
    }
    StemBlockArgs a;
    a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.total_tiles = p.total_tiles;
    a.H = H; a.W = W;
    for (int i = 0; i < 64; ++i) a.sb0[i] = sb0_host[i];
    a.w0 = w0; a.scale1 = scale1; a.bias1 = bias1; a.dbg = dbg;
    static const bool pdl = !(tune_env("YB_TC_PDL") && atoi(tune_env("YB_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(kFThreads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = in_f16 ? cudaLaunchKernelEx(&cfg, stem_block_kernel<__half>, tmX, p.tmW1, p.tmOut, a)
                                 : cudaLaunchKernelEx(&cfg, stem_block_kernel<float>, tmX, p.tmW1, p.tmOut, a);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace yb
