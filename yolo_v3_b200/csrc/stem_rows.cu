// K1s: the Cin = 3 stem (3 -> 32, 3x3, stride 1) from a PIXEL-ROW patch, for image widths that are multiples of 38.
//
// stem_tc_kernel (conv_tc.cu) has every producer thread gather the 27 taps of its output pixel and write a 64-byte
// im2col row: 27 loads, 14 conversions and 4 shared-memory stores per pixel, which makes that kernel instruction-issue
// bound (profiles/r01e_full_raw_stem.csv).  Here the patch of a 3 x 38 output tile is stored once as [pixel][8 fp16]
// (c0, c1, c2, 0...: 16 bytes per pixel, 5 x 40 pixels) -- 3 loads, 2 conversions and 1 store per PATCH pixel -- and the
// tensor core reads the taps out of it directly.  In the un-swizzled K-major layout a K = 16 MMA reads, for row m, one
// 16-byte chunk at start + 16*m and a second one LBO bytes further, and both the start row and LBO are free
// (tools/probes/umma_shift_probe.cu): the first chunk is the pixel array seen from tap a, the second the same array seen
// from tap b.  Five MMAs (tap pairs (0,1) (2,3) (4,5) (6,7) (8,-)) of M = 128, N = 32, K = 16 make one tile.
// GEMM rows follow the patch pitch (m = r*40 + c), 114 of 128 are real outputs -- as in conv_halo.cu, whose epilogue
// (compact staging rows, 4-D TMA store box {32 ch, 38, 3, 1}) this kernel shares in spirit.
#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "yb_internal.h"

namespace yb {
namespace {

constexpr int kSP = 40, kSC = 38, kSR = 3;
constexpr int kSPatchPix = (kSR + 2) * kSP;        // 200
constexpr int kSSlotBytes = 216 * 16;              // the last view reads up to row 82 + 127 = 209
constexpr int kSStages = 12;
constexpr int kSAcc = 4;                           // TMEM accumulators of 32 columns
constexpr int kSRing = 4;                          // epilogue staging buffers: 128 rows x 64 bytes
constexpr int kSThreads = 832;                     // warps 0-7 producers, 8-23 epilogue (4 groups of 4), 24 MMA + TMEM, 25 store issuer
constexpr int kSMmaWarp = 24, kSStoreWarp = 25;
constexpr uint32_t kSStgBytes = 128 * 64;

struct StemRowsArgs {
    const float* x;
    int B, H, W;
    int tiles_x, tiles_y, total_tiles;
    const __half* w;                               // [32][32] fp16, k = (ky*3+kx)*3 + c (the layout stem_tc_kernel uses)
    const float* scale; const float* bias;
    int* dbg;
};

__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo_bytes) {
    // un-swizzled K-major: LBO = bytes between the two 16-byte chunks of a K = 16 step, SBO = 128 (8 rows of 16 bytes)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tma_store_4d_s(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(kSThreads, 1) stem_rows_kernel(const __grid_constant__ CUtensorMap tmOut, const StemRowsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    // header: full[12] | empty[12] | tfull[4] | tempty[4] | sempty[4] | sready[4] | tmem_ptr
    const uint32_t full0 = base, empty0 = base + 96, tfull0 = base + 192, tempty0 = base + 224, sempty0 = base + 256, sready0 = base + 288;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 320);
    float* tab = reinterpret_cast<float*>(gen + 512);                  // scale[32] | bias[32]
    const uint32_t wsm = base + 1024;                                  // 5 MMAs x [2 chunks][32 rows][16 B] = 5 KB
    const uint32_t stg0 = base + 8 * 1024;                             // kSRing x 8 KB
    const uint32_t stage0 = stg0 + kSRing * kSStgBytes;                // kSStages x kSSlotBytes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmOut);
    if (warp == kSMmaWarp && lane == 0) {
        for (int s = 0; s < kSStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int i = 0; i < kSAcc; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
        for (int i = 0; i < kSRing; ++i) { mbar_init(sempty0 + 8 * i, 1); mbar_init(sready0 + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kSMmaWarp) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 32 * kSAcc);
    if (threadIdx.x < 64) tab[threadIdx.x] = threadIdx.x < 32 ? __ldg(a.scale + threadIdx.x) : __ldg(a.bias + threadIdx.x - 32);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 320) {
        // weights of MMA i, chunk j = tap 2i+j (zero for the tenth): 8 fp16 per output channel n, rows 16 bytes apart
        const int e = threadIdx.x - 64, i = e / 64, j = (e >> 5) & 1, n = e & 31, t = 2 * i + j;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t < 9) {
            const __half* w = a.w + n * 32 + t * 3;
            const __half z = __float2half(0.f);
            __half2 p0 = __halves2half2(w[0], w[1]), p1 = __halves2half2(w[2], z);
            v.x = *reinterpret_cast<uint32_t*>(&p0);
            v.y = *reinterpret_cast<uint32_t*>(&p1);
        }
        *reinterpret_cast<uint4*>(gen + 1024 + i * 1024 + j * 512 + n * 16) = v;
    }
    if (threadIdx.x < 16 * kSStages) {
        // rows 200..215 of every patch slot are only ever read into by-product GEMM rows; keep them finite anyway
        const int s = threadIdx.x / 16, r = threadIdx.x % 16;
        *reinterpret_cast<uint4*>(gen + (stage0 - base) + s * kSSlotBytes + (kSPatchPix + r) * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait_prior();              // the output buffer may still be read by the previous step's kernels

    const int tile_first = blockIdx.x, tile_step = gridDim.x;
    if (warp < 8) {
        // ===== producers: every warp builds whole patches on its own (tiles i = warp, warp + 8, ... of this CTA), seven
        // pixels per lane, so eight tiles are in the works at once: a patch costs one proxy fence + one barrier arrival,
        // and with a single 256-thread group that chain (about 1600 cycles) was paid tile after tile =====
        const long HW = (long)a.H * a.W;
        auto fetch = [&](int tile, float (&v)[7][3]) {
            int gy0 = 0, gx0 = 0;
            const float* img = a.x;
            const bool live = tile < a.total_tiles;
            if (live) {
                const int xt = tile % a.tiles_x, rest = tile / a.tiles_x;
                gy0 = (rest % a.tiles_y) * kSR - 1;
                gx0 = xt * kSC - 1;
                img = a.x + (long)(rest / a.tiles_y) * 3 * HW;
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int pp = lane + 32 * k, ppy = pp / kSP;              // patch pixel (row, column): division by a constant
                const int gy = gy0 + ppy, gx = gx0 + pp - ppy * kSP;
                const bool ok = live && lane + 32 * k < kSPatchPix && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                const float* p = img + (long)gy * a.W + gx;
                v[k][0] = ok ? __ldg(p) : 0.f;
                v[k][1] = ok ? __ldg(p + HW) : 0.f;
                v[k][2] = ok ? __ldg(p + 2 * HW) : 0.f;
            }
        };
        float v[7][3];
        int i = warp;                                      // index of this CTA's i-th tile
        fetch(tile_first + i * tile_step, v);
        for (; tile_first + i * tile_step < a.total_tiles; i += 8) {
            uint32_t h0[7], h1[7];                         // pack (waits for the loads issued one iteration ago) ...
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const __half2 a0 = __floats2half2_rn(v[k][0], v[k][1]), a1 = __floats2half2_rn(v[k][2], 0.f);
                h0[k] = *reinterpret_cast<const uint32_t*>(&a0);
                h1[k] = *reinterpret_cast<const uint32_t*>(&a1);
            }
            fetch(tile_first + (i + 8) * tile_step, v);    // ... and put the next patch's loads in flight
            const int stage = i % kSStages;
            const uint32_t phase = (uint32_t)(i / kSStages) & 1u;
            mbar_wait(empty0 + 8 * stage, phase ^ 1, a.dbg, 0, stage);
            uint8_t* slot = gen + (stage0 - base) + stage * kSSlotBytes;
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (lane + 32 * k < kSPatchPix) *reinterpret_cast<uint4*>(slot + (lane + 32 * k) * 16) = make_uint4(h0[k], h1[k], 0u, 0u);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * stage);
        }
    } else if (warp == kSMmaWarp) {
        // ===== MMA issuer: five K = 16 steps, each pairing two taps through the leading-dimension offset =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(32);
            int stage = 0;
            uint32_t phase = 0, acc = 0, acc_phase = 0;
            for (int tile = tile_first; tile < a.total_tiles; tile += tile_step) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, a.dbg, 1, 100 + (int)acc);
                mbar_wait(full0 + 8 * stage, phase, a.dbg, 1, stage);
                tc_fence_after();
                const uint32_t slot = stage0 + stage * kSSlotBytes;
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const int ta = 2 * m, tb = m < 4 ? 2 * m + 1 : 2 * m;        // the tenth "tap" has zero weights
                    const int oa = (ta / 3) * kSP + ta % 3, ob = (tb / 3) * kSP + tb % 3;
                    umma_f16(tmem_base + acc * 32, desc_noswz(slot + oa * 16, (uint32_t)((m < 4 ? ob - oa : 1) * 16)),
                             desc_noswz(wsm + m * 1024, 512u), idesc, m != 0);
                }
                umma_commit(empty0 + 8 * stage);
                umma_commit(tfull0 + 8 * acc);
                if (++stage == kSStages) { stage = 0; phase ^= 1; }
                if (++acc == kSAcc) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == kSStoreWarp) {
        // ===== store issuer =====
        if (lane == 0) {
            uint32_t g = 0;
            for (int tile = tile_first; tile < a.total_tiles; tile += tile_step, ++g) {
                const int xt = tile % a.tiles_x, rest = tile / a.tiles_x;
                const uint32_t buf = g % kSRing, ph = (g / kSRing) & 1u;
                mbar_wait(sready0 + 8 * buf, ph, a.dbg, 4, 700 + (int)buf);
                tma_store_4d_s(&tmOut, stg0 + buf * kSStgBytes, 0, xt * kSC, (rest % a.tiles_y) * kSR, rest / a.tiles_y);
                tma_store_commit();
                if (g >= 2) {
                    tma_store_wait_read<2>();
                    mbar_arrive(sempty0 + 8 * ((g - 2) % kSRing));
                }
            }
            tma_store_wait_all();
        }
        __syncwarp();
    } else if (warp >= 8 && warp < 24) {
        // ===== epilogue: four groups of four warps, group k takes this CTA's tiles k, k+4, ... (its own accumulator and
        // its own staging buffer).  A tile's epilogue is a dependent chain -- accumulator ready, TMEM load, math, shared
        // stores, proxy fence, hand-over to the store issuer -- of roughly a thousand cycles; with all epilogue warps on
        // one tile that chain, not any pipe, set the pace of both stem kernels =====
        const int qd = warp & 3, grp = (warp - 8) >> 2;
        const int m = qd * 32 + lane;
        const int r = m / kSP, c = m - r * kSP;
        const bool valid = r < kSR && c < kSC;
        const int mp = r * kSC + c;
        const int xr = (mp >> 1) & 3;
        const uint32_t acc = (uint32_t)grp, buf = (uint32_t)grp;      // kSAcc == kSRing == 4 groups
        uint8_t* srow = gen + (stg0 - base) + buf * kSStgBytes + mp * 64;
        uint32_t n = 0;                                               // uses of this group's accumulator / buffer so far
        for (int tile = tile_first + grp * tile_step; tile < a.total_tiles; tile += 4 * tile_step, ++n) {
            const uint32_t ph = n & 1u;
            mbar_wait(tfull0 + 8 * acc, ph, a.dbg, 2, 200 + (int)acc);
            tc_fence_after();
            uint32_t r0[16], r1[16];
            const uint32_t taddr = tmem_base + acc * 32 + ((uint32_t)(qd * 32) << 16);
            tmem_ld16(taddr, r0);
            tmem_ld16(taddr + 16, r1);
            mbar_wait(sempty0 + 8 * buf, ph ^ 1, a.dbg, 2, 500 + (int)buf);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (valid) {
#pragma unroll
                for (int h = 0; h < 4; ++h) {                          // four 16-byte chunks = 32 channels
                    const uint32_t* rr = h < 2 ? r0 : r1;
                    const float4 s0 = reinterpret_cast<const float4*>(tab)[2 * h], s1 = reinterpret_cast<const float4*>(tab)[2 * h + 1];
                    const float4 b0 = reinterpret_cast<const float4*>(tab + 32)[2 * h], b1 = reinterpret_cast<const float4*>(tab + 32)[2 * h + 1];
                    const int j = (h & 1) * 8;
                    uint4 pk;
                    __half2* ph2 = reinterpret_cast<__half2*>(&pk);
                    ph2[0] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j + 0]), s0.x, b0.x)), leaky(fmaf(__uint_as_float(rr[j + 1]), s0.y, b0.y)));
                    ph2[1] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j + 2]), s0.z, b0.z)), leaky(fmaf(__uint_as_float(rr[j + 3]), s0.w, b0.w)));
                    ph2[2] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j + 4]), s1.x, b1.x)), leaky(fmaf(__uint_as_float(rr[j + 5]), s1.y, b1.y)));
                    ph2[3] = __floats2half2_rn(leaky(fmaf(__uint_as_float(rr[j + 6]), s1.z, b1.z)), leaky(fmaf(__uint_as_float(rr[j + 7]), s1.w, b1.w)));
                    *reinterpret_cast<uint4*>(srow + ((h ^ xr) << 4)) = pk;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sready0 + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kSMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32 * kSAcc);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc_s = nullptr;

}  // namespace

bool stem_rows_supported(int B, int H, int W) {
    // Experimental, off by default (YB_STEM_ROWS=1 turns it on; read at every call so that tests can switch it):
    // correct on every shape tested, but at 608x608 batch 32 it takes 0.344 ms against 0.270 ms for stem_tc_kernel.
    // ncu (profiles/r01l_full_raw_stem_rows.csv): the tensor pipe is 40 % active with five M128 x N32 x K16 MMAs per
    // tile -- an MMA this small still occupies the pipe for ~76 cycles, so 2.5x more MMAs per tile cost more than the
    // 9x cheaper patch construction saves -- and the single-buffered producers wait out their loads (long-scoreboard
    // stalls dominate).
    const char* e = getenv("YB_STEM_ROWS");
    return e && atoi(e) != 0 && B > 0 && H > 0 && W % kSC == 0;
}

std::string stem_rows_make_plan(StemRowsPlan& p, __half* out, long out_ld, int B, int H, int W, int num_sms) {
    if (!g_enc_s) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeTiled not available from the driver";
        g_enc_s = reinterpret_cast<EncodeTiledFn>(f);
    }
    p.tiles_x = W / kSC;
    p.tiles_y = (H + kSR - 1) / kSR;
    p.total_tiles = B * p.tiles_x * p.tiles_y;
    p.grid = std::min(p.total_tiles, num_sms);
    p.smem = 1024 + 8 * 1024 + (size_t)kSRing * kSStgBytes + (size_t)kSStages * kSSlotBytes;
    cuuint64_t dims[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)out_ld * 2, (cuuint64_t)W * out_ld * 2, (cuuint64_t)H * W * out_ld * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)kSC, (cuuint32_t)kSR, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_enc_s(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled(stem rows output) failed with CUresult " + std::to_string((int)r);
    return "";
}

cudaError_t stem_rows_launch(const StemRowsPlan& p, const float* x, int B, int H, int W, const __half* w16, const float* scale,
                             const float* bias, int* dbg, cudaStream_t s) {
    StemRowsArgs a{};
    a.x = x; a.B = B; a.H = H; a.W = W;
    a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.total_tiles = p.total_tiles;
    a.w = w16; a.scale = scale; a.bias = bias; a.dbg = dbg;
    static PerDeviceOnce attr_once;
    {
        cudaError_t e = attr_once.run([] { return cudaFuncSetAttribute(stem_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); });
        if (e != cudaSuccess) return e;
    }
    static const bool pdl = !(getenv("YB_TC_PDL") && atoi(getenv("YB_TC_PDL")) == 0);
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, stem_rows_kernel, p.tmOut, a);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace yb
