// K1: fused convolution as a tcgen05 implicit GEMM (YB_MODE_FP16), sm_100a only.
//
// Computes what conv_bn_relu / res_layer / UpsampleGroup / the plain head conv compute in the
// reference (darknet.py:27-53, :118, :153-162) for NHWC fp16 activations:
//
//     out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] ),   m = (image, oy, ox),  k = (ky, kx, cin)
//
//   * A is never materialised: the producer warp issues TMA loads -- an im2col-mode tensor map for
//     3x3 (stride 1 or 2; out-of-bounds pixels are zero-filled by the hardware = the conv padding), a
//     plain tiled map for 1x1 -- of 128 output pixels x one k-block (64 channels of one filter tap)
//     straight into 128B-swizzled shared memory; weights [Cout][ky][kx][Cin] come through a tiled map;
//   * one elected thread issues tcgen05.mma (kind::f16, M=128, N=BN<=256, K=16) with the fp32
//     accumulator in TMEM; tcgen05.commit releases smem stages and publishes finished accumulators;
//   * TMEM holds two accumulators, so the four epilogue warps drain tile i (tcgen05.ld -> fp32
//     scale/bias (folded eval BN) -> LeakyReLU -> +residual -> fp16 -> 16-byte stores, optionally to
//     the 2x2 nearest-upsampled positions of a concat slice) while the MMA warp works on tile i+1;
//   * persistent CTAs (one per SM), static round-robin over the (m-tile, n-tile) grid.
//
// Layers with Cin == 32 use 32-channel k-blocks with the 64B swizzle; everything else 64-channel
// k-blocks with the 128B swizzle.  The Cin == 3 stem runs on CUDA cores (conv_simt.cu).
#include <algorithm>

#include "yb_internal.h"

namespace yb {
namespace {

constexpr int kBM = 128;
constexpr int kMaxStages = 8;
constexpr int kThreads = 256;            // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr size_t kSmemBudget = 227 * 1024;
constexpr size_t kSmemHeader = 1024;     // barriers + tmem pointer
constexpr long long kWatchdogCycles = 4000000000LL;

struct TcArgs {
    long M;
    int Ho, Wo, HoWo;
    int ks, stride, pad;
    int cin_blocks, num_kblocks;
    int BN, n_tiles, m_tiles;
    int stages, tmem_cols;
    const float* scale; const float* bias;
    void* out; long out_ld; int out_f32;
    const __half* res; long res_ld;
    int leaky, upsample;
    int* dbg;
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrival must not hang the GPU.  On timeout the waiter records who it is in
// host-visible memory and traps; the host sees a launch failure with the diagnostic attached.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* dbg, int role, int which) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kWatchdogCycles) {
            if (dbg) {
                dbg[1] = (int)blockIdx.x; dbg[2] = role; dbg[3] = which; dbg[4] = (int)parity;
                __threadfence_system();
                dbg[0] = 1;
                __threadfence_system();
            }
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major, set to 1), SBO >> 4 in [32,46) = bytes
// between 8-row groups, descriptor version 1 in [46,48), layout type in [61,64)
// (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
template <int SWZ>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    constexpr uint64_t sbo = (8 * SWZ) >> 4;
    constexpr uint64_t layout = SWZ == 128 ? 2 : 4;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits [4,6) = 1), A/B fp16
// (0), both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24); }

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : v * kLeaky; }

// Epilogue for 16 consecutive channels of one output pixel.
__device__ __forceinline__ void epilogue16(const TcArgs& a, const uint32_t (&acc)[16], int n, bool valid, long m,
                                           long o00, long W2ld) {
    float v[16];
    const float4* sc = reinterpret_cast<const float4*>(a.scale + n);
    const float4* bi = reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 s4 = __ldg(sc + q), b4 = __ldg(bi + q);
        v[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]), s4.x, b4.x);
        v[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), s4.y, b4.y);
        v[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), s4.z, b4.z);
        v[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), s4.w, b4.w);
    }
    if (a.leaky) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = leaky(v[i]);
    }
    if (!valid) return;
    if (a.out_f32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + m * a.out_ld + n);
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        return;
    }
    if (a.res) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.res + m * a.res_ld + n);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint4 r = rp[q];
            const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h[i]);
                v[8 * q + 2 * i] += f.x;
                v[8 * q + 2 * i + 1] += f.y;
            }
        }
    }
    uint4 pk[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        __half2* h = reinterpret_cast<__half2*>(&pk[q]);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[8 * q + 2 * i], v[8 * q + 2 * i + 1]);
    }
    __half* ob = reinterpret_cast<__half*>(a.out);
    if (!a.upsample) {
        uint4* o = reinterpret_cast<uint4*>(ob + m * a.out_ld + n);
        o[0] = pk[0]; o[1] = pk[1];
    } else {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            uint4* o = reinterpret_cast<uint4*>(ob + o00 + (d >> 1) * W2ld + (d & 1) * a.out_ld + n);
            o[0] = pk[0]; o[1] = pk[1];
        }
    }
}

template <int SWZ>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    constexpr int BKE = SWZ / 2;                  // fp16 elements per k-block row
    constexpr uint32_t A_BYTES = kBM * SWZ;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;           // swizzle atoms need 1024-byte alignment
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t B_BYTES = (uint32_t)a.BN * SWZ;
    const uint32_t stage_bytes = A_BYTES + ((B_BYTES + 1023u) & ~1023u);
    // header: full[8] | empty[8] | tmem_full[2] | tmem_empty[2] | tmem_ptr
    const uint32_t full0 = base, empty0 = base + 8 * kMaxStages;
    const uint32_t tfull0 = base + 16 * kMaxStages, tempty0 = tfull0 + 16;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 16 * kMaxStages + 32);
    const uint32_t stage0 = base + (uint32_t)kSmemHeader;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = a.m_tiles * a.n_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t acc_stride = (uint32_t)a.tmem_cols >> 1;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
                const long m0 = (long)m_tile * kBM;
                const int n0 = n_tile * a.BN;
                int cw = 0, chh = 0, cn = 0;
                if (a.ks == 3) {
                    cn = (int)(m0 / a.HoWo);
                    const int r = (int)(m0 - (long)cn * a.HoWo);
                    const int p = r / a.Wo, q = r - p * a.Wo;
                    cw = q * a.stride - a.pad;
                    chh = p * a.stride - a.pad;
                }
                for (int kb = 0; kb < a.num_kblocks; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1, a.dbg, 0, stage);
                    const uint32_t sA = stage0 + stage * stage_bytes, sB = sA + A_BYTES;
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_arrive_expect_tx(fb, A_BYTES + B_BYTES);
                    if (a.ks == 1) {
                        tma_load_2d(&tmA, sA, fb, kb * BKE, (int)m0);
                    } else {
                        const int tap = kb / a.cin_blocks, cb = kb - tap * a.cin_blocks;
                        const int kh = tap / 3, kw = tap - kh * 3;
                        tma_load_im2col(&tmA, sA, fb, cb * BKE, cw, chh, cn, (uint16_t)kw, (uint16_t)kh);
                    }
                    tma_load_2d(&tmB, sB, fb, kb * BKE, n0);
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(a.BN);
            int stage = 0;
            uint32_t phase = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, a.dbg, 1, 100 + (int)acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * acc_stride;
                for (int kb = 0; kb < a.num_kblocks; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase, a.dbg, 1, stage);
                    tc_fence_after();
                    const uint32_t sA = stage0 + stage * stage_bytes, sB = sA + A_BYTES;
                    const uint64_t adesc = make_smem_desc<SWZ>(sA), bdesc = make_smem_desc<SWZ>(sB);
#pragma unroll
                    for (int k = 0; k < BKE / 16; ++k)   // 16 fp16 = 32 bytes per MMA: descriptor address += 2
                        umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(empty0 + 8 * stage);     // frees this smem stage once the MMAs have read it
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull0 + 8 * acc);           // accumulator complete
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp - 4;                               // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
            const long m = (long)m_tile * kBM + row;
            const int n0 = n_tile * a.BN;
            const bool valid = m < a.M;
            long o00 = 0, W2ld = 0;
            if (a.upsample && valid) {
                const int img = (int)(m / a.HoWo);
                const int r = (int)(m - (long)img * a.HoWo);
                const int y = r / a.Wo, x = r - y * a.Wo;
                W2ld = 2L * a.Wo * a.out_ld;
                o00 = (((long)img * 2 * a.Ho + 2 * y) * 2 * a.Wo + 2 * x) * a.out_ld;
            }
            mbar_wait(tfull0 + 8 * acc, acc_phase, a.dbg, 2, 200 + (int)acc);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < a.BN; c0 += 32) {
                uint32_t r0[16], r1[16];
                const bool two = c0 + 16 < a.BN;
                tmem_ld16(taddr + c0, r0);
                if (two) tmem_ld16(taddr + c0 + 16, r1);
                tmem_ld_wait();
                epilogue16(a, r0, n0 + c0, valid, m, o00, W2ld);
                if (two) epilogue16(a, r1, n0 + c0 + 16, valid, m, o00, W2ld);
            }
            tc_fence_before();
            mbar_arrive(tempty0 + 8 * acc);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

std::string load_driver_entry_points() {
    if (g_encode_tiled && g_encode_im2col) return "";
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeTiled not available from the driver";
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f);
    f = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &qr);
    if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeIm2col not available from the driver";
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f);
    return "";
}

std::string cu_err(const char* what, CUresult r) { return std::string(what) + " failed with CUresult " + std::to_string((int)r); }

}  // namespace

bool tc_supported(const ConvArgs& a) {
    if (a.ks != 1 && a.ks != 3) return false;
    if (a.ks == 1 && a.stride != 1) return false;
    if (a.Cin != 32 && a.Cin % 64 != 0) return false;
    if (a.in_ld % 8 || a.out_ld % 8 || (a.res && a.res_ld % 8)) return false;
    return true;
}

std::string tc_make_plan(TcPlan& p, const ConvArgs& a, const __half* w16, int cout_pad, int K, int num_sms) {
    std::string e = load_driver_entry_points();
    if (!e.empty()) return e;
    p.swz = a.Cin == 32 ? 64 : 128;
    const int bke = p.swz / 2;
    p.cin_blocks = a.Cin / bke;
    p.num_kblocks = a.ks * a.ks * p.cin_blocks;
    p.M = (long)a.B * a.Ho * a.Wo;
    p.m_tiles = (int)((p.M + kBM - 1) / kBM);
    // tile N: the divisor of cout_pad (<= 256, multiple of 16) with the fewest persistent waves x width
    int best_bn = 0;
    double best_cost = 0;
    for (int bn = std::min(cout_pad, 256); bn >= 16; bn -= 16) {
        if (cout_pad % bn) continue;
        if (bn < 64 && bn != cout_pad) break;
        const long tiles = (long)p.m_tiles * (cout_pad / bn);
        const long waves = (tiles + num_sms - 1) / num_sms;
        const double cost = (double)waves * (bn + 24);      // +24: per-tile fixed cost in units of N columns
        if (!best_bn || cost < best_cost) { best_bn = bn; best_cost = cost; }
    }
    if (!best_bn) return "no valid tile width for cout_pad=" + std::to_string(cout_pad);
    p.BN = best_bn;
    p.n_tiles = cout_pad / p.BN;
    int tc = 32;
    while (tc < 2 * p.BN) tc <<= 1;
    p.tmem_cols = tc;
    const size_t stage_bytes = (size_t)kBM * p.swz + (((size_t)p.BN * p.swz + 1023) & ~(size_t)1023);
    p.stages = (int)std::min<size_t>(kMaxStages, (kSmemBudget - kSmemHeader - 1024) / stage_bytes);
    if (p.stages < 2) return "not enough shared memory for two pipeline stages";
    p.smem = kSmemHeader + 1024 + p.stages * stage_bytes;
    p.grid = (int)std::min<long>((long)p.m_tiles * p.n_tiles, num_sms);

    const CUtensorMapSwizzle swz = p.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    // B: weights [cout_pad][K] fp16, K contiguous; box = one k-block x BN rows
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(__half)};
        cuuint32_t box[2] = {(cuuint32_t)bke, (cuuint32_t)p.BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = g_encode_tiled(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w16), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cu_err("cuTensorMapEncodeTiled(weights)", r);
    }
    if (a.ks == 1) {
        // A: [M][Cin] with pixel pitch in_ld; rows past M are zero-filled
        cuuint64_t dims[2] = {(cuuint64_t)a.Cin, (cuuint64_t)p.M};
        cuuint64_t strides[1] = {(cuuint64_t)a.in_ld * sizeof(__half)};
        cuuint32_t box[2] = {(cuuint32_t)bke, (cuuint32_t)kBM};
        cuuint32_t es[2] = {1, 1};
        CUresult r = g_encode_tiled(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.in), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cu_err("cuTensorMapEncodeTiled(activations)", r);
    } else {
        // A: NHWC as (C, W, H, N); im2col box of 128 output pixels x one k-block of channels.  The
        // bounding box of filter-window origins is [-pad, dim-1+upper] with upper = pad-(ks-1); the
        // traversal stride is the conv stride; the (kx, ky) tap arrives as the instruction's offsets.
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.in_ld * sizeof(__half), (cuuint64_t)a.W * a.in_ld * sizeof(__half),
                                 (cuuint64_t)a.H * a.W * a.in_ld * sizeof(__half)};
        int lower[2] = {-a.pad, -a.pad};
        int upper[2] = {a.pad - (a.ks - 1), a.pad - (a.ks - 1)};
        cuuint32_t es[4] = {1, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1};
        CUresult r = g_encode_im2col(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a.in), dims, strides, lower, upper,
                                     (cuuint32_t)bke, (cuuint32_t)kBM, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cu_err("cuTensorMapEncodeIm2col(activations)", r);
        // Driver workaround that CUTLASS applies for im2col descriptors of small tensors
        // (cute/atom/copy_traits_sm90_im2col.hpp, driver <= 13.1, tensor < 128 KiB): clear bit 21 of
        // the descriptor's second 64-bit word.
        int drv = 0;
        if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010) {
            const size_t bytes = ((size_t)(a.B - 1) * a.H * a.W + (size_t)(a.H - 1) * a.W + (a.W - 1)) * a.in_ld * sizeof(__half) +
                                 (size_t)a.Cin * sizeof(__half);
            if (bytes < 131072) reinterpret_cast<uint64_t*>(&p.tmA)[1] &= ~(1ull << 21);
        }
    }
    return "";
}

cudaError_t tc_launch(const TcPlan& p, const ConvArgs& a, int* dbg, cudaStream_t s) {
    TcArgs t;
    t.M = p.M;
    t.Ho = a.Ho; t.Wo = a.Wo; t.HoWo = a.Ho * a.Wo;
    t.ks = a.ks; t.stride = a.stride; t.pad = a.pad;
    t.cin_blocks = p.cin_blocks; t.num_kblocks = p.num_kblocks;
    t.BN = p.BN; t.n_tiles = p.n_tiles; t.m_tiles = p.m_tiles;
    t.stages = p.stages; t.tmem_cols = p.tmem_cols;
    t.scale = a.scale; t.bias = a.bias;
    t.out = a.out; t.out_ld = a.out_ld; t.out_f32 = a.out_f32;
    t.res = static_cast<const __half*>(a.res); t.res_ld = a.res_ld;
    t.leaky = a.leaky; t.upsample = a.upsample;
    t.dbg = dbg;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (p.swz == 128) conv_tc_kernel<128><<<p.grid, kThreads, p.smem, s>>>(p.tmA, p.tmB, t);
    else conv_tc_kernel<64><<<p.grid, kThreads, p.smem, s>>>(p.tmA, p.tmB, t);
    return cudaGetLastError();
}

}  // namespace yb
