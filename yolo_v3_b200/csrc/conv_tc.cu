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
//   * TMEM holds two accumulators, so the four epilogue warps drain tile i while the MMA warp works on
//     tile i+1: tcgen05.ld -> fp32 scale/bias (folded eval BN) -> LeakyReLU -> +residual -> fp16,
//     staged through a ring of 128B-swizzled shared-memory sub-tiles (128 rows x 128 bytes) that are
//     written back with TMA stores (cp.async.bulk.tensor ... bulk_group: full-line coalesced writes,
//     M tail clipped by the tensor map); the residual sub-tiles are prefetched into the same ring by
//     a dedicated TMA load warp, so the in-place "x + f(x)" costs no exposed global-load latency.
//     The two 1x1 "up" convs instead store each pixel directly to its 2x2 nearest-upsampled block of
//     the concat slice;
//   * persistent CTAs (one per SM), static round-robin over the (m-tile, n-tile) grid.
//
// Layers with Cin == 32 use 32-channel k-blocks with the 64B swizzle; everything else 64-channel
// k-blocks with the 128B swizzle.  The Cin == 3 stem lives in stem_halo.cu / stem_block.cu.
//
// Every role reads its warp index through a shuffle broadcast, which the compiler knows to be warp-uniform (what
// cutlass::canonical_warp_idx_sync does): the role dispatch is then a uniform branch and the single-thread issue loops and
// the epilogue run on the uniform datapath instead of being littered with R2UR copies (validated in round 2:
// profiles/r02a_uw_ab.txt, +5 % on the whole step).
//
// Build flavours: the release library reads no environment variable.  -DYB_EXPERIMENTS (make EXPERIMENTS=1) adds the
// tuning overrides used by tools/layer_bench.py sweeps (tune_env) and the in-kernel clock64 time line (YB_TC_TRACE).
#include <algorithm>
#include <cstdlib>

#include "yb_internal.h"
#include "tc_ptx.cuh"

namespace yb {
namespace {

constexpr int kBM = 128;
constexpr int kMaxStages = 16;
constexpr int kThreads = 768;            // warps 0-1 TMA producers, 2 MMA, 3 TMEM alloc + store issuer, 4 residual TMA, 8-23 epilogue
constexpr int kThreadsSplit = 640;       // split mode: warp 0 TMA producer, 1 residual TMA, 2 MMA, 3 TMEM alloc + store issuer, 4-19 epilogue
                                         // (five warpgroups: 96 registers per thread -- the epilogue keeps 32 fp32 accumulators)
constexpr int kEpiWarps = 16;
constexpr int kMaxRing = 4;              // epilogue staging ring depth
constexpr size_t kSmemBudget = 227 * 1024;
constexpr size_t kSmemHeader = 1024;     // barriers + tmem pointer

struct TcArgs {
    long M;
    int Ho, Wo, HoWo;
    int ks, stride, pad;
    int cin_blocks, num_kblocks;
    int kps, num_iters;      // k-blocks per pipeline stage, stages per tile (num_kblocks / kps)
    int BN, n_tiles, m_tiles;
    int stages, tmem_cols;
    const float* scale; const float* bias;
    const float* tab;        // device-side: smem copy of scale[cout_pad] | bias[cout_pad] (set inside the kernel)
    int cout_pad, tab_bytes; // tab_bytes = 2*cout_pad*4 rounded up to 1 KB
    void* out; long out_ld; int out_f32;
    const __half* res; long res_ld;
    int leaky, upsample;
    // staged (TMA store) epilogue
    int epi_staged;          // 1: smem ring + TMA store, 0: direct global stores
    int ring;                // staging buffers
    int sub_bytes;           // bytes per row of a sub-tile (128 or 64) == swizzle span
    int cs;                  // columns per sub-tile
    int n_sub;               // sub-tiles per tile (BN / cs)
    int has_res;
    int b_resident;          // 1: the CTA's whole weight slab [BN][K] is loaded once and stays in smem
    int srel;                // store issuer: staging-buffer stores allowed to stay unread (0, 1 or 2)
    int epi_split;           // sub-tiles of 32 columns occupy only half of the sixteen epilogue warps: the halves take alternate
                             // sub-tiles (n_sub even) or alternate tiles (n_sub == 1), two hand-over chains in flight
    // split mode (SPLIT instantiations, YB_MODE_FP32_TC): every activation is a pair of fp16 tensors hi + lo sharing one
    // pixel pitch; a_lo / out_lo / res_lo = channel offset of the lo half relative to the hi pointer; split_out = the
    // output is written as such a pair (everything but the fp32 head maps)
    int a_lo, out_lo, res_lo, split_out;
    int chunk_pairs, n_chunks;   // two-level accumulation: unit pairs (12 MMAs each) per TMEM chunk, chunks per tile
    int* dbg;
    long long* trace;        // optional [6 roles][64 tiles][4] clock64 stamps of CTA 0 (YB_TC_TRACE=1)
};

// clock64 time line of CTA 0 (YB_TC_TRACE=1): only in -DYB_EXPERIMENTS builds, the release kernels carry no trace points
#ifdef YB_EXPERIMENTS
#define YB_TRACE(role, idx, slot)                                                                   \
    do {                                                                                            \
        if (a.trace && blockIdx.x == 0 && (idx) < 64) a.trace[((role) * 64 + (idx)) * 4 + (slot)] = clock64(); \
    } while (0)
#else
#define YB_TRACE(role, idx, slot) do { } while (0)
#endif

// Epilogue for 16 consecutive channels of one output pixel.
__device__ __forceinline__ void epilogue16(const TcArgs& a, const uint32_t (&acc)[16], int n, bool valid, long m,
                                           long o00, long W2ld) {
    float v[16];
    // scale/bias come from the shared-memory table: with ~227 KB of smem carved out there is no L1 left,
    // so global loads here would each pay an L2 round trip
    const float4* sc = reinterpret_cast<const float4*>(a.tab + n);
    const float4* bi = reinterpret_cast<const float4*>(a.tab + a.cout_pad + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 s4 = sc[q], b4 = bi[q];
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

// Staged epilogue for 16 consecutive channels of one output pixel: scale/bias, LeakyReLU, residual
// (read from the swizzled staging sub-tile the load warp filled), convert, write back in place.
// `srow` points at this pixel's row of the sub-tile; chunk c (16 bytes) lives at ((c ^ xr) << 4).
__device__ __forceinline__ void epilogue16_staged(const TcArgs& a, const uint32_t (&acc)[16], int n, uint8_t* srow,
                                                  int grp, int xr) {
    float v[16];
    const float4* sc = reinterpret_cast<const float4*>(a.tab + n);            // shared-memory table (see above)
    const float4* bi = reinterpret_cast<const float4*>(a.tab + a.cout_pad + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 s4 = sc[q], b4 = bi[q];
        v[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]), s4.x, b4.x);
        v[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), s4.y, b4.y);
        v[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), s4.z, b4.z);
        v[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), s4.w, b4.w);
    }
    if (a.leaky) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = leaky(v[i]);
    }
    if (a.out_f32) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float4* p = reinterpret_cast<float4*>(srow + (((grp * 4 + h) ^ xr) << 4));
            *p = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
        return;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint4* p = reinterpret_cast<uint4*>(srow + (((grp * 2 + h) ^ xr) << 4));
        if (a.has_res) {
            const uint4 r = *p;
            const __half2* hh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hh[i]);
                v[8 * h + 2 * i] += f.x;
                v[8 * h + 2 * i + 1] += f.y;
            }
        }
        uint4 pk;
        __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) ph[i] = __floats2half2_rn(v[8 * h + 2 * i], v[8 * h + 2 * i + 1]);
        *p = pk;
    }
}

// ---- split mode (YB_MODE_FP32_TC): fp32-grade convolutions on the fp16 tensor pipe ----------------------------------
// Every activation x is stored as two fp16 tensors, hi = RN16(x) and lo = RN16(x - hi) (x - hi is exact in fp32, so
// hi + lo carries 22 bits of x), and every weight w -- pre-scaled per output channel by a power of two so that its lo part
// stays a normal fp16 number, the scale is undone in the fp32 epilogue -- as wh + wl.  The GEMM accumulates the three
// significant partial products xh*wh + xh*wl + xl*wh (exact in the fp32 accumulator's product stage; the dropped xl*wl
// term is 2^-22 relative).  Per (tap, channel block) the producer loads TWO k-block units -- unit 0 = (A lo, B wh), unit 1 =
// (A hi, B wl); the weight rows are packed [tap][channel block][wh | wl][64] to match -- and the MMA warp issues THREE groups
// on them: xl*wh and xh*wl (the two corrections first, while the accumulator is small), then xh*wh from unit 1's A tile and
// unit 0's B tile, so that every operand tile crosses L2 -> shared memory once (four tiles per three MMA groups).  tools/split_numerics.py
// (profiles/r02_split_numerics_cpu.txt): the operand representation error of this scheme through all 75 layers is 3e-6 of
// max|logit| -- below the fp32 oracle's own 5e-6.
//
// TWO-LEVEL ACCUMULATION.  The tcgen05 fp32 accumulator does not round to nearest: every MMA truncates the running sum
// (profiles/r02a_tc_accum_probe.txt, r02c_split_error_vs_chain_length.txt: the result shrinks toward zero by 1.7e-8 per
// MMA in the chain -- 1.4e-5 for the 864 MMAs of a 512->1024 3x3 layer, against 1e-6 for an fp32 GEMM on the CPU), so
// a long K loop into one TMEM accumulator cannot be fp32-grade.  The SPLIT kernels therefore accumulate in TMEM only over
// a CHUNK of one (tap, channel block) pair of units -- 12 MMAs, of which only the last four carry the main term -- and the
// epilogue warps add every finished chunk into fp32 REGISTER accumulators with round-to-nearest adds while the tensor
// core works on the next chunk in the other TMEM buffer.  A thread owns at most 32 accumulators (two 16-column groups),
// which caps the tile width at 128 columns (64 for the fp32 head maps).
__device__ __forceinline__ void split2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 f = __half22float2(hi);
    lo = __floats2half2_rn(a - f.x, b - f.y);
}

// Staged split epilogue: the staging slot holds the hi sub-tile followed (lo_delta bytes later) by the lo sub-tile, both
// 128B- (or 64B-) swizzled like the single sub-tile of the fp16 mode; the residual arrives the same way.
__device__ __forceinline__ void epilogue16_staged_split(const TcArgs& a, const float (&acc)[16], int n, uint8_t* srow,
                                                        uint32_t lo_delta, int grp, int xr) {
    // eight values (one 16-byte chunk of the hi and of the lo sub-tile) at a time: the caller holds up to 64 accumulators
    const float4* sc = reinterpret_cast<const float4*>(a.tab + n);
    const float4* bi = reinterpret_cast<const float4*>(a.tab + a.cout_pad + n);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 s4 = sc[2 * h + q], b4 = bi[2 * h + q];
            v[4 * q + 0] = fmaf(acc[8 * h + 4 * q + 0], s4.x, b4.x);
            v[4 * q + 1] = fmaf(acc[8 * h + 4 * q + 1], s4.y, b4.y);
            v[4 * q + 2] = fmaf(acc[8 * h + 4 * q + 2], s4.z, b4.z);
            v[4 * q + 3] = fmaf(acc[8 * h + 4 * q + 3], s4.w, b4.w);
        }
        if (a.leaky) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = leaky(v[i]);
        }
        uint4* p = reinterpret_cast<uint4*>(srow + (((grp * 2 + h) ^ xr) << 4));
        uint4* pl = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p) + lo_delta);
        if (a.has_res) {
            const uint4 r = *p, rl = *pl;
            const __half2* hh = reinterpret_cast<const __half2*>(&r);
            const __half2* ll = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fh = __half22float2(hh[i]), fl = __half22float2(ll[i]);
                v[2 * i] += fh.x + fl.x;                  // hi + lo is exact in fp32: one rounding, like the reference's x + f(x)
                v[2 * i + 1] += fh.y + fl.y;
            }
        }
        uint4 pk, pkl;
        __half2* ph = reinterpret_cast<__half2*>(&pk);
        __half2* pq = reinterpret_cast<__half2*>(&pkl);
#pragma unroll
        for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], ph[i], pq[i]);
        *p = pk;
        *pl = pkl;
    }
}

// (m-unit, n-tile) of the tiles a CTA visits and the staging-ring slot of its sub-tiles, carried without divisions: these
// loops run on one thread each or sit on the epilogue's per-tile dependent chain, which is what paces the layers with
// small tiles (profiles/r01s_trace_layer2_ring*.txt) -- a 32-bit division there costs ~150-200 cycles of latency.
struct TileWalk {
    int m_unit, n_tile;
    int dq, dr, n_tiles;
    __device__ __forceinline__ TileWalk(int first, int step, int nt) : n_tiles(nt) {
        m_unit = first / nt; n_tile = first - m_unit * nt;
        dq = step / nt; dr = step - dq * nt;
    }
    __device__ __forceinline__ void next() {
        m_unit += dq; n_tile += dr;
        if (n_tile >= n_tiles) { n_tile -= n_tiles; ++m_unit; }
    }
};
struct RingWalk {
    uint32_t buf = 0, ph = 0;          // slot g % ring and parity (g / ring) & 1 of the g-th sub-tile
    __device__ __forceinline__ void next(uint32_t ring) { if (++buf == ring) { buf = 0; ph ^= 1u; } }
};

// Resident weights: the CTA's slab [BN / NCTA rows][K] is loaded once (one elected thread) and stays in shared memory.
// The plan makes the number of tiles in flight a multiple of n_tiles, so every tile of a CTA has the same n-tile.  In
// pair mode each CTA keeps its half of the rows and both halves complete on the leader's barrier (the leader's MMA warp
// is the only reader), which therefore expects the bytes of both.
template <bool CTA2>
__device__ __forceinline__ void load_resident_weights(const CUtensorMap* tmB, const TcArgs& a, uint32_t bres0, uint32_t b_slot,
                                                      uint32_t b_bytes, uint32_t bres_bar, int bke, uint32_t rank) {
    constexpr int NCTA = CTA2 ? 2 : 1;
    const int n0 = (((int)blockIdx.x / NCTA) % a.n_tiles) * a.BN + (int)rank * (a.BN / NCTA);
    if (rank == 0) mbar_arrive_expect_tx(bres_bar, (uint32_t)a.num_kblocks * b_bytes * NCTA);
#pragma unroll 1
    for (int kb = 0; kb < a.num_kblocks; ++kb) {
        if constexpr (CTA2) tma_load_2d_pair(tmB, bres0 + kb * b_slot, bres_bar, kb * bke, n0);
        else tma_load_2d(tmB, bres0 + kb * b_slot, bres_bar, kb * bke, n0);
    }
}

// SUBS (split mode): sub-tiles a thread of the epilogue accumulates in registers -- 2: tiles up to 128 columns, 32
// accumulators; 4: 256-column tiles, 64 accumulators, for which the four role warps donate registers to the epilogue
// warpgroups with setmaxnreg (96 -> 56 / 104).  Such a kernel must not call out-of-line functions from the re-partitioned
// regions -- ptxas fails its register allocation (C7600) -- hence mbar_wait<kDonate>.
template <int SWZ, bool CTA2, bool SPLIT = false, int SUBS = 2>
__global__ void __launch_bounds__(SPLIT ? kThreadsSplit : kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes, const TcArgs a_in) {
    constexpr int BKE = SWZ / 2;                  // fp16 elements per k-block row
    constexpr bool kDonate = SPLIT && SUBS == 4;  // register donation (setmaxnreg): no out-of-line calls in this kernel
    TcArgs a = a_in;
    constexpr uint32_t A_BYTES = kBM * SWZ;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;           // swizzle atoms need 1024-byte alignment
    uint8_t* gen = smem_raw + (base - raw);
    // CTA pair: rank r owns rows [2*m_unit + r]*128.. of the 256-row UMMA tile and half of the B rows
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    constexpr int NCTA = CTA2 ? 2 : 1;
    const uint32_t B_BYTES = (uint32_t)(a.BN / NCTA) * SWZ;
    const uint32_t kb_bytes = A_BYTES + (a.b_resident ? 0u : ((B_BYTES + 1023u) & ~1023u));   // one k-block inside a stage
    const uint32_t stage_bytes = kb_bytes * (uint32_t)a.kps;
    // header: full[16] | empty[16] | tmem_full[2] | tmem_empty[2] | tmem_ptr | sfull[4] | sempty[4] | bres
    const uint32_t full0 = base, empty0 = base + 8 * kMaxStages;
    const uint32_t tfull0 = base + 16 * kMaxStages, tempty0 = tfull0 + 16;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 16 * kMaxStages + 32);
    const uint32_t sfull0 = base + 16 * kMaxStages + 64, sempty0 = sfull0 + 32, bres_bar = sempty0 + 32, sready0 = bres_bar + 16;
    // staging ring (epilogue), resident weight slab, then the operand pipeline stages; all 1024-byte aligned
    const uint32_t stg_half = (uint32_t)kBM * (uint32_t)a.sub_bytes;
    const bool split_out = SPLIT && a.split_out;                       // staging slots hold a hi and a lo sub-tile
    const uint32_t stg_bytes = split_out ? 2u * stg_half : stg_half;
    a.tab = reinterpret_cast<const float*>(gen + kSmemHeader);          // scale | bias table
    const uint32_t stg0 = base + (uint32_t)kSmemHeader + (uint32_t)a.tab_bytes;
    const uint32_t B_SLOT = (B_BYTES + 1023u) & ~1023u;
    const uint32_t bres0 = stg0 + (a.epi_staged ? (uint32_t)a.ring * ((stg_bytes + 1023u) & ~1023u) : 0u);
    const uint32_t stage0 = bres0 + (a.b_resident ? (uint32_t)a.num_kblocks * B_SLOT : 0u);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int total_tiles = a.m_tiles * a.n_tiles;          // m_tiles counts 256-row units in pair mode
    const int tile_first = blockIdx.x / NCTA, tile_step = gridDim.x / NCTA;

    constexpr int NT = SPLIT ? kThreadsSplit : kThreads;
    for (int i = threadIdx.x; i < a.cout_pad; i += NT) {
        float* tab = reinterpret_cast<float*>(gen + kSmemHeader);
        tab[i] = __ldg(a.scale + i);
        tab[a.cout_pad + i] = __ldg(a.bias + i);
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (a.epi_staged) prefetch_tmap(&tmOut);
        if (a.epi_staged && a.has_res) prefetch_tmap(&tmRes);
    }
    if (warp == 2 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        // arrivals per accumulator: every staged-epilogue warp that reads it (split by tiles: only one half of them does)
        const int acc_readers = !a.epi_staged ? 4 : (a.epi_split && a.n_sub == 1) ? kEpiWarps / 2 : kEpiWarps;
        const int sub_writers = a.epi_split ? kEpiWarps / 2 : kEpiWarps;
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, acc_readers * NCTA); }
        for (int i = 0; i < kMaxRing; ++i) { mbar_init(sfull0 + 8 * i, 1); mbar_init(sempty0 + 8 * i, 1); mbar_init(sready0 + 8 * i, sub_writers); }
        mbar_init(bres_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        if constexpr (CTA2) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr)), (uint32_t)a.tmem_cols);
        else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), (uint32_t)a.tmem_cols);
    }
    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // peer barriers must exist before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t acc_stride = (uint32_t)a.tmem_cols >> 1;
    if (threadIdx.x == 0) YB_TRACE(5, 0, 1);
    pdl_launch_dependents();      // everything above (barriers, TMEM, table) overlaps the previous layer's tail
    __syncwarp();
    pdl_wait_prior();             // activations written by the previous layer are complete and visible
    if (threadIdx.x == 0) YB_TRACE(5, 0, 2);
    [[maybe_unused]] auto split_epilogue = [&]() {
        // ===== epilogue of the split mode: every finished TMEM chunk is added into fp32 register accumulators (round to
        // nearest -- the second level of the accumulation, see the comment on split2), then the usual staged hand-over:
        // scale/bias, LeakyReLU, residual, hi/lo split, swizzled smem sub-tile, TMA store.  Warp w owns TMEM lanes
        // 32*(w%4).. and the 16-column group (w-4)/4 of each of the (at most SUBS) sub-tiles =====
        const int q = warp & 3, part = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int xr = a.sub_bytes == 128 ? (row & 7) : ((row >> 1) & 3);
        const bool active = part < (a.cs >> 4);
        uint32_t accb = 0, acc_phase = 0;
        RingWalk rw;
        const int n_wrap = a.n_tiles * a.BN;
        int n0 = (tile_first % a.n_tiles) * a.BN;
        const int dn0 = (tile_step % a.n_tiles) * a.BN;
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            float sacc[SUBS][16];
#pragma unroll
            for (int j = 0; j < SUBS; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) sacc[j][i] = 0.f;
            for (int ch = 0; ch < a.n_chunks; ++ch) {
                mbar_wait<kDonate>(tfull0 + 8 * accb, acc_phase, a.dbg, 2, 200 + (int)accb);
                tc_fence_after();
                const uint32_t tcol = tmem_base + accb * acc_stride + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 16);
                // SUBS == 2: both sub-tiles in one round trip to TMEM; SUBS == 4: one at a time (16 values in flight next to
                // the 64 accumulators -- the chunk's MMAs take four times as long as these four round trips)
                constexpr int STEP = SUBS == 2 ? 2 : 1;
#pragma unroll
                for (int j = 0; j < SUBS; j += STEP) {
                    uint32_t r0[16];
                    [[maybe_unused]] uint32_t r1[16];
                    const bool h0 = active && j < a.n_sub;
                    [[maybe_unused]] const bool h1 = STEP == 2 && active && j + 1 < a.n_sub;
                    if (h0) tmem_ld16(tcol + (uint32_t)(j * a.cs), r0);
                    if constexpr (STEP == 2) { if (h1) tmem_ld16(tcol + (uint32_t)((j + 1) * a.cs), r1); }
                    tmem_ld_wait();
                    if (j + STEP >= SUBS) {                    // last read of this chunk: the buffer goes back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (CTA2) mbar_arrive_leader(tempty0 + 8 * accb); else mbar_arrive(tempty0 + 8 * accb);
                        }
                    }
                    if (h0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sacc[j][i] = __fadd_rn(sacc[j][i], __uint_as_float(r0[i]));
                    }
                    if constexpr (STEP == 2) {
                        if (h1) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) sacc[j + 1][i] = __fadd_rn(sacc[j + 1][i], __uint_as_float(r1[i]));
                        }
                    }
                }
                accb ^= 1;
                if (accb == 0) acc_phase ^= 1;
            }
#pragma unroll
            for (int j = 0; j < SUBS; ++j) {
                if (j < a.n_sub) {
                    const uint32_t buf = rw.buf, ph = rw.ph;
                    if (a.has_res) mbar_wait<kDonate>(sfull0 + 8 * buf, ph, a.dbg, 2, 400 + (int)buf);
                    else mbar_wait<kDonate>(sempty0 + 8 * buf, ph ^ 1, a.dbg, 2, 500 + (int)buf);
                    uint8_t* srow = gen + (stg0 - base) + buf * stg_bytes + (uint32_t)row * (uint32_t)a.sub_bytes;
                    const int nb = n0 + j * a.cs + part * 16;
                    if (active) {
                        if constexpr (SUBS == 4) {             // (the fp32 head maps keep the narrow tiles of the SUBS == 2 kernels)
                            epilogue16_staged_split(a, sacc[j], nb, srow, stg_half, part, xr);
                        } else if (split_out) {
                            epilogue16_staged_split(a, sacc[j], nb, srow, stg_half, part, xr);
                        } else {                               // fp32 head maps: the plain staged epilogue on the summed values
                            uint32_t rr[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(sacc[j][i]);
                            epilogue16_staged(a, rr, nb, srow, part, xr);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sready0 + 8 * buf);
                    rw.next((uint32_t)a.ring);
                }
            }
            n0 += dn0;
            if (n0 >= n_wrap) n0 -= n_wrap;
        }
    };
    // SUBS == 4: 640 threads x 96 registers at launch.  Warpgroup 0 (the four role warps: single-thread loops) keeps 56, each
    // of the four epilogue warpgroups grows to 104: 64 fp32 accumulators + a 16-value TMEM load + addressing.  The two
    // setmaxnreg instructions sit at the heads of the two branches, so that each dominates the code it governs.
    if (kDonate && warp >= 4) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        split_epilogue();
    } else {
    if constexpr (kDonate) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp < (SPLIT ? 1 : 2)) {
        // ===== TMA producers (two warps take alternate pipeline stages -- one warp takes all of them in split mode; whole warp runs the loop, one
        // elected lane issues).  Both walk the same stage/coordinate sequence and act on their own parity. =====
        {
            const uint32_t pw = (uint32_t)warp;
            uint32_t itg = 0;                             // running stage counter across tiles
            int stage = 0;
            uint32_t phase = 0;
            if (pw == 0 && a.b_resident && tile_first < total_tiles && elect_one())
                load_resident_weights<CTA2>(&tmB, a, bres0, B_SLOT, B_BYTES, bres_bar, BKE, rank);
            __syncwarp();
            // The loop below runs on one thread; everything that can be is carried incrementally
            // (no divisions, no address recomputation) because its latency paces the pipeline.
            const uint32_t tx_bytes = (a.b_resident ? A_BYTES : A_BYTES + B_BYTES) * (uint32_t)a.kps * NCTA;
            const bool load_b = !a.b_resident;
            uint32_t sA = stage0, fb = full0, eb = empty0;
            int ti = 0;
            TileWalk tw(tile_first, tile_step, a.n_tiles);
            for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++ti, tw.next()) {
                const int m_unit = tw.m_unit, n_tile = tw.n_tile;
                const int m0 = (m_unit * NCTA + (int)rank) * kBM;
                const int n0 = n_tile * a.BN + (int)rank * (a.BN / NCTA);   // this CTA's half of the weight rows
                if (pw == 0) YB_TRACE(0, ti, 0);
                if (a.ks == 1) {
                    int kc = 0;
                    [[maybe_unused]] int cc = 0, sec = 0;      // SPLIT: channel of the A box inside its half, unit 0 (lo) / 1 (hi)
                    for (int it = 0; it < a.num_iters; ++it, ++itg) {
                        const bool mine = SPLIT || (itg & 1u) == pw;
                        if (mine) mbar_wait<kDonate>(eb, phase ^ 1, a.dbg, 0, stage);
                        const bool el = mine && elect_one();
                        if (leader && el) mbar_arrive_expect_tx(fb, tx_bytes);
                        uint32_t dst = sA;
                        for (int j = 0; j < a.kps; ++j) {
                            if (el) {
                                const int ac = SPLIT ? cc + (sec == 0 ? a.a_lo : 0) : kc;
                                if constexpr (CTA2) {
                                    tma_load_2d_pair(&tmA, dst, fb, ac, m0);
                                    if (load_b) tma_load_2d_pair(&tmB, dst + A_BYTES, fb, kc, n0);
                                } else {
                                    tma_load_2d(&tmA, dst, fb, ac, m0);
                                    if (load_b) tma_load_2d(&tmB, dst + A_BYTES, fb, kc, n0);
                                }
                            }
                            if constexpr (SPLIT) {
                                if (++sec == 2) { sec = 0; cc += BKE; }
                            }
                            kc += BKE;
                            dst += kb_bytes;
                        }
                        sA += stage_bytes; fb += 8; eb += 8;
                        if (++stage == a.stages) { stage = 0; phase ^= 1; sA = stage0; fb = full0; eb = empty0; }
                    }
                } else {
                    const int cn = m0 / a.HoWo;
                    const int r = m0 - cn * a.HoWo;
                    const int p = r / a.Wo, q = r - p * a.Wo;
                    const int cw = q * a.stride - a.pad, chh = p * a.stride - a.pad;
                    int kc = 0, cc = 0;                      // k coordinate of the weights, channel coordinate of A
                    [[maybe_unused]] int sec = 0;            // SPLIT: unit 0 (lo) / 1 (hi) of the current (tap, channel block)
                    uint16_t kw = 0, kh = 0;
                    const int cend = a.cin_blocks * BKE;
                    for (int it = 0; it < a.num_iters; ++it, ++itg) {
                        const bool mine = SPLIT || (itg & 1u) == pw;
                        if (mine) mbar_wait<kDonate>(eb, phase ^ 1, a.dbg, 0, stage);
                        const bool el = mine && elect_one();
                        if (leader && el) mbar_arrive_expect_tx(fb, tx_bytes);
                        uint32_t dst = sA;
                        for (int j = 0; j < a.kps; ++j) {
                            if (el) {
                                const int ac = SPLIT ? cc + (sec == 0 ? a.a_lo : 0) : cc;
                                if constexpr (CTA2) {
                                    tma_load_im2col_pair(&tmA, dst, fb, ac, cw, chh, cn, kw, kh);
                                    if (load_b) tma_load_2d_pair(&tmB, dst + A_BYTES, fb, kc, n0);
                                } else {
                                    tma_load_im2col(&tmA, dst, fb, ac, cw, chh, cn, kw, kh);
                                    if (load_b) tma_load_2d(&tmB, dst + A_BYTES, fb, kc, n0);
                                }
                            }
                            kc += BKE;
                            bool next_cb = true;
                            if constexpr (SPLIT) { if (++sec < 2) next_cb = false; else sec = 0; }
                            if (next_cb) {
                                cc += BKE;
                                if (cc == cend) { cc = 0; if (++kw == 3) { kw = 0; ++kh; } }
                            }
                            dst += kb_bytes;
                        }
                        sA += stage_bytes; fb += 8; eb += 8;
                        if (++stage == a.stages) { stage = 0; phase ^= 1; sA = stage0; fb = full0; eb = empty0; }
                    }
                }
                if (pw == 0) YB_TRACE(0, ti, 1);
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===== MMA issuer (whole warp of the leader CTA, one elected lane issues) =====
        if (leader) {
            const uint32_t idesc = make_idesc(a.BN, kBM * NCTA);
            int stage = 0;
            uint32_t phase = 0, acc = 0, acc_phase = 0;
            // descriptors are carried incrementally: the start-address field counts 16-byte units
            const uint64_t adesc0 = make_smem_desc<SWZ>(stage0);
            const uint64_t stage_inc = stage_bytes >> 4, kb_inc = kb_bytes >> 4, b_off = A_BYTES >> 4;
            const uint64_t bres_desc = make_smem_desc<SWZ>(bres0), bres_inc = B_SLOT >> 4;
            uint64_t adesc = adesc0;
            uint32_t fbar = full0, ebar = empty0;
            bool ready = false;
            if (a.b_resident && tile_first < total_tiles) mbar_wait<kDonate>(bres_bar, 0, a.dbg, 1, 600);
            int ti = 0;
            if constexpr (SPLIT) {
                // Split mode: the pipeline carries k-block UNITS in pairs -- unit 0 = (A lo, B wh), unit 1 = (A hi, B wl), in one
                // stage (kps == 2) or in two consecutive stages (kps == 1; the stage count is even) -- and every pair gets three
                // MMA groups: lo*wh, hi*wl, then hi*wh from unit 1's A tile and unit 0's B tile.  A TMEM chunk is chunk_pairs
                // pairs; the finished chunk goes to the epilogue warps, the next one starts in the other TMEM buffer.
                const int pairs = a.num_kblocks >> 1;
                const uint32_t upair = a.kps == 2 ? 1u : 2u;                    // stages per pair
                const uint32_t u1_off = a.kps == 2 ? kb_bytes : stage_bytes;      // byte distance of unit 1 from unit 0
                for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++ti) {
                    YB_TRACE(1, ti, 0);
                    mbar_wait<kDonate>(tempty0 + 8 * acc, acc_phase ^ 1, a.dbg, 1, 100 + (int)acc);
                    tc_fence_after();
                    YB_TRACE(1, ti, 1);
                    uint32_t d_tmem = tmem_base + acc * acc_stride;
                    int pc = 0;
                    uint32_t fresh = 0;                                           // 0: the chunk's first MMA overwrites the accumulator
                    for (int p = 0; p < pairs; ++p) {
                        if (pc == a.chunk_pairs) {
                            if (elect_one()) {
                                if constexpr (CTA2) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);
                            }
                            acc ^= 1;
                            if (acc == 0) acc_phase ^= 1;
                            mbar_wait<kDonate>(tempty0 + 8 * acc, acc_phase ^ 1, a.dbg, 1, 100 + (int)acc);
                            tc_fence_after();
                            d_tmem = tmem_base + acc * acc_stride;
                            pc = 0; fresh = 0;
                        }
                        ++pc;
                        mbar_wait<kDonate>(full0 + 8 * stage, phase, a.dbg, 1, stage);
                        if (upair == 2) mbar_wait<kDonate>(full0 + 8 * (stage + 1), phase, a.dbg, 1, stage + 1);
                        tc_fence_after();
                        const uint32_t s0 = stage0 + (uint32_t)stage * stage_bytes;
                        const uint64_t a0 = make_smem_desc<SWZ>(s0), a1 = make_smem_desc<SWZ>(s0 + u1_off);
                        const uint64_t b0 = a.b_resident ? bres_desc + (uint64_t)(2 * p) * bres_inc : a0 + b_off;
                        const uint64_t b1 = a.b_resident ? b0 + bres_inc : a1 + b_off;
                        if (elect_one()) {
                            if constexpr (CTA2) {
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16_pair(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, fresh | (uint32_t)k);
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16_pair(d_tmem, a1 + 2 * k, b1 + 2 * k, idesc, 1);
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16_pair(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, 1);
                                umma_commit_pair(empty0 + 8 * stage);
                                if (upair == 2) umma_commit_pair(empty0 + 8 * (stage + 1));
                            } else {
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, fresh | (uint32_t)k);
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16(d_tmem, a1 + 2 * k, b1 + 2 * k, idesc, 1);
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, 1);
                                umma_commit(empty0 + 8 * stage);
                                if (upair == 2) umma_commit(empty0 + 8 * (stage + 1));
                            }
                        }
                        fresh = 1;
                        stage += (int)upair;
                        if (stage >= a.stages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) {
                        if constexpr (CTA2) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);   // last chunk complete
                    }
                    YB_TRACE(1, ti, 2);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            } else {
            for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++ti) {
                YB_TRACE(1, ti, 0);
                mbar_wait<kDonate>(tempty0 + 8 * acc, acc_phase ^ 1, a.dbg, 1, 100 + (int)acc);
                tc_fence_after();
                YB_TRACE(1, ti, 1);
                const uint32_t d_tmem = tmem_base + acc * acc_stride;
                int kb = 0;
                for (int it = 0; it < a.num_iters; ++it) {
                    if (!ready) mbar_wait<kDonate>(fbar, phase, a.dbg, 1, stage);
                    tc_fence_after();
                    // probe the NEXT stage's barrier now: its latency overlaps the MMA issue below
                    const bool wrap = stage + 1 == a.stages;
                    ready = mbar_test_wait(wrap ? full0 : fbar + 8, wrap ? phase ^ 1 : phase);
                    const bool el = elect_one();
                    uint64_t ad = adesc;
                    for (int j = 0; j < a.kps; ++j, ++kb) {
                        const uint64_t bd = a.b_resident ? bres_desc + (uint64_t)kb * bres_inc : ad + b_off;
                        if (el) {
                            if constexpr (CTA2) {
#pragma unroll
                                for (int k = 0; k < BKE / 16; ++k) umma_f16_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                            } else {
                                umma_f16(d_tmem, ad, bd, idesc, kb != 0);
#pragma unroll
                                for (int k = 1; k < BKE / 16; ++k)   // 16 fp16 = 32 bytes per MMA: descriptor address += 2
                                    umma_f16_imm<1>(d_tmem, ad + 2 * k, bd + 2 * k, idesc);
                            }
                        }
                        ad += kb_inc;
                    }
                    if (el) {
                        if constexpr (CTA2) umma_commit_pair(ebar); else umma_commit(ebar);   // frees the stage (in both CTAs)
                    }
                    adesc += stage_inc; fbar += 8; ebar += 8;
                    if (++stage == a.stages) { stage = 0; phase ^= 1; adesc = adesc0; fbar = full0; ebar = empty0; }
                }
                if (elect_one()) {
                    if constexpr (CTA2) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);   // accumulator complete
                }
                YB_TRACE(1, ti, 2);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            }
        }
        __syncwarp();
    } else if (warp == (SPLIT ? 1 : 4)) {
        // ===== residual prefetch: TMA loads of the residual sub-tiles into the staging ring =====
        if (lane == 0 && a.epi_staged && a.has_res) {
            RingWalk rw;
            TileWalk tw(tile_first, tile_step, a.n_tiles);
            for (int tile = tile_first; tile < total_tiles; tile += tile_step, tw.next()) {
                const int m_unit = tw.m_unit, n_tile = tw.n_tile;
                const int m0 = (m_unit * NCTA + (int)rank) * kBM, n0 = n_tile * a.BN;
                for (int j = 0; j < a.n_sub; ++j, rw.next((uint32_t)a.ring)) {
                    const uint32_t buf = rw.buf, ph = rw.ph;
                    mbar_wait<kDonate>(sempty0 + 8 * buf, ph ^ 1, a.dbg, 3, 300 + (int)buf);
                    mbar_arrive_expect_tx(sfull0 + 8 * buf, stg_bytes);
                    tma_load_2d(&tmRes, stg0 + buf * stg_bytes, sfull0 + 8 * buf, n0 + j * a.cs, m0);
                    if (split_out) tma_load_2d(&tmRes, stg0 + buf * stg_bytes + stg_half, sfull0 + 8 * buf, a.res_lo + n0 + j * a.cs, m0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===== store issuer: waits until the eight epilogue warps have filled a staging sub-tile, writes it
        // back with one TMA store and recycles the buffer once the store has drained it =====
        if (lane == 0 && a.epi_staged) {
            uint32_t g = 0, prev1 = 0, prev2 = 0;          // slots of sub-tiles g-1 and g-2
            RingWalk rw;
            TileWalk tw(tile_first, tile_step, a.n_tiles);
            for (int tile = tile_first; tile < total_tiles; tile += tile_step, tw.next()) {
                const int m_unit = tw.m_unit, n_tile = tw.n_tile;
                const int m0 = (m_unit * NCTA + (int)rank) * kBM, n0 = n_tile * a.BN;
                for (int j = 0; j < a.n_sub; ++j, ++g, rw.next((uint32_t)a.ring)) {
                    const uint32_t buf = rw.buf, ph = rw.ph;
                    mbar_wait<kDonate>(sready0 + 8 * buf, ph, a.dbg, 4, 700 + (int)buf);
                    tma_store_2d(&tmOut, stg0 + buf * stg_bytes, n0 + j * a.cs, m0);
                    if (split_out) tma_store_2d(&tmOut, stg0 + buf * stg_bytes + stg_half, a.out_lo + n0 + j * a.cs, m0);
                    tma_store_commit();
                    // Recycle buffers: keep at most `srel` stores unread.  srel = ring - 2 is the latest release that
                    // still lets the epilogue warps start sub-tile g+1 while sub-tile g is being finished; with the
                    // two-deep ring that means handing the buffer back as soon as THIS store has read it (releasing
                    // buffer g-1 only after store g was issued chained the sub-tiles one after another).
                    if (a.srel == 0) {
                        tma_store_wait_read<0>();
                        mbar_arrive(sempty0 + 8 * buf);
                    } else if (a.srel == 1) {
                        if (g >= 1) { tma_store_wait_read<1>(); mbar_arrive(sempty0 + 8 * prev1); }
                    } else {
                        if (g >= 2) { tma_store_wait_read<2>(); mbar_arrive(sempty0 + 8 * prev2); }
                    }
                    prev2 = prev1; prev1 = buf;
                }
            }
            tma_store_wait_all();
        }
        __syncwarp();
    } else if (SPLIT && !kDonate && warp >= 4 && warp < 20) {
        split_epilogue();
    } else if (warp >= 8 && a.epi_staged) {
        // ===== epilogue (staged): TMEM -> registers -> swizzled smem sub-tile -> TMA store =====
        // Sixteen warps: warp w reads TMEM lanes 32*(w%4).. (its rows); the four warps of a row quarter each
        // take one 16-column group of every sub-tile, so a thread's share is 16 values and four warps per
        // scheduler hide each other's instruction latency (the epilogue is issue-bound, not memory-bound).
        const int q = warp & 3, praw = (warp - 8) >> 2;
        // split mode (32-column sub-tiles): warps with praw 0-1 form group 0, praw 2-3 group 1; the groups take the even /
        // odd sub-tiles of the CTA's sub-tile sequence G = tile index * n_sub + j -- alternate sub-tiles of every tile when
        // n_sub is even, alternate tiles when n_sub == 1 -- so two hand-over chains are in flight instead of one
        const int egrp = a.epi_split ? (praw >> 1) : 0;
        const int part = a.epi_split ? (praw & 1) : praw;
        const bool split_tiles = a.epi_split && a.n_sub == 1;
        const int jstep = (a.epi_split && !split_tiles) ? 2 : 1;
        const int j_first = jstep == 2 ? egrp : 0;
        const int j_last = a.n_sub - jstep + j_first;         // the last sub-tile of a tile this warp reads
        const int tmul = split_tiles ? 2 : 1;                 // tiles advanced per iteration
        const int row = q * 32 + lane;
        const int xr = a.sub_bytes == 128 ? (row & 7) : ((row >> 1) & 3);
        const int ngrp = a.cs >> 4;                           // 16-column groups per sub-tile: 1, 2 or 4
        const bool active = part < ngrp;
        const bool issuer = (warp == 8 && lane == 0);
        uint32_t acc = split_tiles ? (uint32_t)egrp : 0u, acc_phase = 0;
        int ti = split_tiles ? egrp : 0;
        RingWalk rw;
        if (a.epi_split && egrp) rw.next((uint32_t)a.ring);   // group 1 starts at G = 1
        // only the tile's first column is needed here: n0 = (tile % n_tiles) * BN, advanced modulo the padded width
        const int n_wrap = a.n_tiles * a.BN;
        const int tile_begin = tile_first + (split_tiles ? egrp * tile_step : 0);
        int n0 = (tile_begin % a.n_tiles) * a.BN;
        const int dn0 = ((tmul * tile_step) % a.n_tiles) * a.BN;
        for (int tile = tile_begin; tile < total_tiles; tile += tmul * tile_step, ti += tmul) {
            if (issuer) YB_TRACE(2, ti, 0);
            mbar_wait<kDonate>(tfull0 + 8 * acc, acc_phase, a.dbg, 2, 200 + (int)acc);
            tc_fence_after();
            if (issuer) YB_TRACE(2, ti, 1);
            const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
            for (int j = j_first; j < a.n_sub; j += jstep) {
                const uint32_t buf = rw.buf, ph = rw.ph;
                uint32_t r0[16];
                const uint32_t tcol = taddr + (uint32_t)(j * a.cs + part * 16);
                if (active) tmem_ld16(tcol, r0);
                // the buffer is ours once the residual landed (res layers) or its previous store drained
                if (a.has_res) mbar_wait<kDonate>(sfull0 + 8 * buf, ph, a.dbg, 2, 400 + (int)buf);
                else mbar_wait<kDonate>(sempty0 + 8 * buf, ph ^ 1, a.dbg, 2, 500 + (int)buf);
                if (issuer && j == 0) YB_TRACE(3, ti, 0);
                tmem_ld_wait();
                if (issuer && j == 0) YB_TRACE(3, ti, 1);
                if (j == j_last) {                            // accumulator fully read by this warp: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CTA2) mbar_arrive_leader(tempty0 + 8 * acc); else mbar_arrive(tempty0 + 8 * acc);
                    }
                }
                uint8_t* srow = gen + (stg0 - base) + buf * stg_bytes + (uint32_t)row * (uint32_t)a.sub_bytes;
                const int nb = n0 + j * a.cs;
                if (active) epilogue16_staged(a, r0, nb + part * 16, srow, part, xr);
                if (issuer && j == 0) YB_TRACE(3, ti, 2);
                fence_async_smem();                           // generic-proxy smem writes -> visible to the TMA engine
                if (issuer && j == 0) YB_TRACE(3, ti, 3);
                __syncwarp();
                if (lane == 0) mbar_arrive(sready0 + 8 * buf);   // no block-wide barrier: warps run ahead independently
                if (issuer && j == 0) YB_TRACE(4, ti, 0);
                rw.next((uint32_t)a.ring);
                if (a.epi_split) rw.next((uint32_t)a.ring);   // the other group's sub-tile
            }
            if (issuer) YB_TRACE(2, ti, 2);
            if (split_tiles) {
                acc_phase ^= 1;                               // this group always reads accumulator `egrp`
            } else {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            n0 += dn0;
            if (n0 >= n_wrap) n0 -= n_wrap;
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== epilogue (direct, warps 8-11 only): TMEM -> registers -> global, used by the nearest-upsample layers.
        // (Spreading the columns over all sixteen warps changes nothing -- profiles/README.md r01x: the 2x2-replicated
        // 32-byte stores, one 768-byte pixel pitch apart per lane, are what these two layers wait for.) =====
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        uint32_t acc = 0, acc_phase = 0;
        TileWalk tw((int)blockIdx.x, (int)gridDim.x, a.n_tiles);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tw.next()) {
            const int m_tile = tw.m_unit, n_tile = tw.n_tile;
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
            mbar_wait<kDonate>(tfull0 + 8 * acc, acc_phase, a.dbg, 2, 200 + (int)acc);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    }

    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // the peer may still signal this CTA's barriers / read its smem
    if (warp == 3) {
        tc_fence_after();
        if constexpr (CTA2) tmem_dealloc_pair(tmem_base, (uint32_t)a.tmem_cols);
        else tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
    if (threadIdx.x == 0) YB_TRACE(5, 0, 3);
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
    p.cout_pad = cout_pad;
    const int bke = p.swz / 2;
    p.cin_blocks = a.Cin / bke;
    p.split = a.split;
    p.num_kblocks = a.ks * a.ks * p.cin_blocks * (a.split ? 2 : 1);   // split mode: two units per (tap, channel block)
    if (a.split && K != 2 * a.ks * a.ks * a.Cin) return "split mode expects the [tap][channel block][wh|wl] weight packing";
    if (a.split && a.upsample) return "split mode: the upsample layers run as a plain convolution followed by a copy kernel";
    const bool split_out = a.split && !a.out_f32;
    p.M = (long)a.B * a.Ho * a.Wo;
    p.m_tiles = (int)((p.M + kBM - 1) / kBM);
    // tile N: the kernel is bound by bytes brought into the SM (A: 128 rows, B: BN rows per k-block), so
    // the cost of a candidate is (persistent waves) x (128 + BN); measured in profiles/r01_layer_sweep_v3.txt
    int best_bn = 0;
    double best_cost = 0;
    // split mode: a thread of the epilogue keeps its share of the tile in registers (two-level accumulation), at most
    // 32 values: 128 columns of fp16 pairs (four warps per 64-column sub-tile) or 64 columns of fp32 (two per 32-column one)
    // (the fp32 head maps keep the SUBS = 2 kernels: 64 columns)
    static const bool split_wide = !(tune_env("YB_SPLIT_WIDE") && atoi(tune_env("YB_SPLIT_WIDE")) == 0);
    const int bn_max = a.split ? (a.out_f32 ? 64 : (split_wide ? 256 : 128)) : 256;
    for (int bn = std::min(cout_pad, bn_max); bn >= 16; bn -= 16) {
        if (cout_pad % bn) continue;
        if (bn < 64 && bn != cout_pad) break;
        const long tiles = (long)p.m_tiles * (cout_pad / bn);
        const long waves = (tiles + num_sms - 1) / num_sms;
        const double cost = (double)waves * (128 + bn);
        if (!best_bn || cost < best_cost) { best_bn = bn; best_cost = cost; }
    }
    if (!best_bn && a.split) {
        // a head whose padded width has no divisor in [64 columns of fp32 ..]: e.g. 20 classes, 75 channels padded to 80 ->
        // 16-column tiles (one 64-byte fp32 sub-tile each)
        for (int bn = std::min(cout_pad, bn_max); bn >= 16 && !best_bn; bn -= 16)
            if (cout_pad % bn == 0 && (bn * (a.out_f32 ? 4 : 2)) % 64 == 0) best_bn = bn;
    }
    if (!best_bn) return "no valid tile width for cout_pad=" + std::to_string(cout_pad);
    p.BN = best_bn;
    if (const char* e = tune_env("YB_TC_BN")) {
        const int bn = atoi(e);
        if (bn >= 16 && bn <= bn_max && bn % 16 == 0 && cout_pad % bn == 0) p.BN = bn;
    }
    p.n_tiles = cout_pad / p.BN;
    // CTA pairs (cta_group::2): 256-row UMMA tiles, each CTA stages only half of the weight rows, which
    // both halves the bytes every SM must ingest per FLOP and frees shared memory for deeper pipelines.
    // BN = 128 3x3 layers whose weight slab (BN x K) is too large to stay resident in one CTA but whose half fits
    // (64 -> 128 stride 2 at 152x152: 147 KB) also run as pairs, each CTA keeping its half resident: streamed, the slab
    // is re-fetched for every tile and the layer is bound by L2 -> SM traffic (1.9 GB, 158 us at the ~6.3 KB/clk the
    // L2 delivers, against 88 us of HBM time; profiles/README.md "L2 -> SM model").
    bool pair128 = false;
    {
        const size_t slot_one = ((size_t)p.BN * p.swz + 1023) & ~(size_t)1023, slot_half = ((size_t)(p.BN / 2) * p.swz + 1023) & ~(size_t)1023;
        pair128 = p.swz == 128 && p.BN == 128 && a.ks == 3 && slot_one * p.num_kblocks > 96 * 1024 &&
                  slot_half * p.num_kblocks <= 96 * 1024 && p.m_tiles > 4 * num_sms;
        if (const char* e = tune_env("YB_TC_PAIR128")) pair128 = pair128 && atoi(e) != 0;
    }
    p.cta2 = p.swz == 128 && (p.BN == 256 || pair128 || (a.split && p.BN == 128 && !a.out_f32)) && !a.upsample && p.m_tiles >= 4 && num_sms % 2 == 0;
    if (const char* e = tune_env("YB_TC_CTA2")) p.cta2 = p.cta2 && atoi(e) != 0;
    const int ncta = p.cta2 ? 2 : 1;
    if (p.cta2) p.m_tiles = (p.m_tiles + 1) / 2;            // 256-row units from here on
    int tc = 32;
    while (tc < 2 * p.BN) tc <<= 1;
    p.tmem_cols = tc;
    // epilogue: TMA-store staging whenever a tile row splits into whole 128- (or 64-) byte sub-tiles
    const int esz = a.out_f32 ? 4 : 2;
    p.epi_staged = 0; p.ring = 0; p.sub_bytes = 128;
    if (!a.upsample) {
        if ((p.BN * esz) % 128 == 0) { p.epi_staged = 1; p.sub_bytes = 128; }
        else if ((p.BN * esz) % 64 == 0) { p.epi_staged = 1; p.sub_bytes = 64; }
    }
    p.cs = p.sub_bytes / esz;
    p.n_sub = p.epi_staged ? p.BN / p.cs : 0;
    if (p.epi_staged) p.ring = a.res && !split_out ? 4 : 2;        // split mode: a slot holds a hi and a lo sub-tile
    if (const char* e = tune_env("YB_TC_RING")) if (p.epi_staged) p.ring = std::max(2, std::min(kMaxRing, atoi(e)));
    // 32-column sub-tiles (Cout = 32 in fp16, the fp32 head maps) keep only eight of the sixteen epilogue warps busy:
    // the two halves then work on alternate sub-tiles, each with its own slot of a four-deep ring
    p.epi_split = p.epi_staged && p.cs == 32 && (p.n_sub == 1 || p.n_sub % 2 == 0) && !a.split;
    if (a.split && (!p.epi_staged || p.n_sub > (a.out_f32 ? 2 : 4))) return "split mode: tile does not fit the register accumulators";
    if (const char* e = tune_env("YB_TC_EPISPLIT")) p.epi_split = p.epi_split && atoi(e) != 0;
    if (p.epi_split) p.ring = 4;
    const size_t stg_bytes = ((size_t)kBM * p.sub_bytes * (split_out ? 2 : 1) + 1023) & ~(size_t)1023;
    const size_t ring_bytes = p.epi_staged ? p.ring * stg_bytes : 0;
    p.grid = p.cta2 ? 2 * (int)std::min<long>((long)p.m_tiles * p.n_tiles, num_sms / 2)
                    : (int)std::min<long>((long)p.m_tiles * p.n_tiles, num_sms);
    // small-K layers: keep the CTA's whole weight slab resident (halves the L2->SM traffic of 1x1 convs)
    const size_t b_slot = ((size_t)(p.BN / ncta) * p.swz + 1023) & ~(size_t)1023;
    const size_t bres_bytes = b_slot * p.num_kblocks;
    p.b_resident = (!p.cta2 || p.BN == 128) && bres_bytes <= 96 * 1024 && (p.grid / ncta) % p.n_tiles == 0 &&
                   p.m_tiles > 2 * num_sms / ncta;
    if (const char* e = tune_env("YB_TC_BRES")) p.b_resident = p.b_resident && atoi(e) != 0;
    // (Negative results of rounds 1-2, removed from the code and kept as measurements under profiles/: an L2 prefetch of
    // the A operand some tiles ahead makes the memory-bound layers SLOWER -- the TMA request path, not DRAM latency,
    // limits them; touching the weights before griddepcontrol.wait, an L2 persistence window over layer outputs and a
    // channel-blocked activation layout change nothing or lose.)
    p.srel = 1;
    if (const char* e = tune_env("YB_TC_SREL")) p.srel = std::max(0, std::min(std::min(2, p.ring - 1), atoi(e)));
    const size_t kb_bytes = (size_t)kBM * p.swz + (p.b_resident ? 0 : b_slot);
    p.tab_bytes = (int)(((size_t)2 * cout_pad * sizeof(float) + 1023) & ~(size_t)1023);
    const size_t fixed = kSmemHeader + 1024 + p.tab_bytes + ring_bytes + (p.b_resident ? bres_bytes : 0);
    // k-blocks per pipeline stage: the producer and MMA loops each run on one thread and cost a few hundred
    // cycles per stage, so a stage should carry >= ~500 tensor-core cycles (one k-block is (BKE/16)*BN/2)
    p.kps = 1;
    {
        const int kb_cycles = (bke / 16) * p.BN / 2;
        for (int c = 4; c >= 2; --c)
            if (p.num_kblocks % c == 0 && c * kb_bytes <= 48 * 1024 && kb_cycles * (c - 1) < 768) { p.kps = c; break; }
        // resident weights, no residual ring: two stages holding ALL taps of a tile fit -> one barrier round
        // trip per tile (measured on the 32->64 stride-2 layer: 0.286 -> 0.247 ms, profiles/README.md)
        if (p.b_resident && !a.res && p.num_kblocks <= 9 && fixed + 2 * p.num_kblocks * kb_bytes <= kSmemBudget)
            p.kps = p.num_kblocks;
        if (const char* e = tune_env("YB_TC_KPS")) { const int c = atoi(e); if (c >= 1 && p.num_kblocks % c == 0 && c * kb_bytes <= 96 * 1024) p.kps = c; }
        // split mode: a pair of units (one (tap, channel block)) per stage, or one unit per stage
        if (a.split) p.kps = 2 * kb_bytes <= 48 * 1024 ? 2 : 1;
    }
    // a resident slab next to a four-deep residual ring can leave room for fewer than two multi-k-block stages
    while (p.kps > 1 && (kSmemBudget - fixed) / (kb_bytes * p.kps) < 2) {
        int c = p.kps - 1;
        while (c > 1 && p.num_kblocks % c) --c;
        p.kps = c;
    }
    if (a.split && p.kps != 2) p.kps = 1;
    const size_t stage_bytes = kb_bytes * p.kps;
    p.stages = (int)std::min<size_t>(kMaxStages, (kSmemBudget - fixed) / stage_bytes);
    if (const char* e = tune_env("YB_TC_STAGES")) p.stages = std::max(2, std::min(p.stages, atoi(e)));
    if (a.split && p.kps == 1) p.stages &= ~1;            // unit pairs must not straddle the end of the stage ring
    if (p.stages < 2) return "not enough shared memory for two pipeline stages";
    p.smem = fixed + p.stages * stage_bytes;
    if (a.split) {
        // unit pairs per chunk: one (12 MMAs of K = 16; 6 with 32-channel k-blocks, where two keep the chunk long enough
        // for the register pass of the previous one to hide behind it)
        p.chunk_pairs = p.swz == 64 ? 2 : 1;
        if (const char* e = tune_env("YB_SPLIT_CHUNK")) p.chunk_pairs = std::max(1, std::min(64, atoi(e)));
        p.n_chunks = (p.num_kblocks / 2 + p.chunk_pairs - 1) / p.chunk_pairs;
    }

    const CUtensorMapSwizzle swz = p.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    // B: weights [cout_pad][K] fp16, K contiguous; box = one k-block x BN rows
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(__half)};
        cuuint32_t box[2] = {(cuuint32_t)bke, (cuuint32_t)(p.BN / ncta)};   // pair mode: each CTA loads half the rows
        cuuint32_t es[2] = {1, 1};
        CUresult r = g_encode_tiled(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w16), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cu_err("cuTensorMapEncodeTiled(weights)", r);
    }
    if (a.ks == 1) {
        // A: [M][Cin] with pixel pitch in_ld; rows past M are zero-filled
        cuuint64_t dims[2] = {(cuuint64_t)(a.split ? a.in_lo + a.Cin : a.Cin), (cuuint64_t)p.M};
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
        cuuint64_t dims[4] = {(cuuint64_t)(a.split ? a.in_lo + a.Cin : a.Cin), (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
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
    if (p.epi_staged) {
        const CUtensorMapSwizzle eswz = p.sub_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        const CUtensorMapDataType dt = a.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        cuuint64_t dims[2] = {(cuuint64_t)(split_out ? a.out_lo + a.Cout : a.Cout), (cuuint64_t)p.M};
        cuuint64_t strides[1] = {(cuuint64_t)a.out_ld * esz};
        cuuint32_t box[2] = {(cuuint32_t)p.cs, (cuuint32_t)kBM};
        cuuint32_t es[2] = {1, 1};
        CUresult r = g_encode_tiled(&p.tmOut, dt, 2, a.out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, eswz,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cu_err("cuTensorMapEncodeTiled(output)", r);
        if (a.res) {
            cuuint64_t rstrides[1] = {(cuuint64_t)a.res_ld * sizeof(__half)};
            cuuint64_t rdims[2] = {(cuuint64_t)(split_out ? a.res_lo + a.Cout : a.Cout), (cuuint64_t)p.M};
            r = g_encode_tiled(&p.tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.res), rdims, rstrides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, eswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return cu_err("cuTensorMapEncodeTiled(residual)", r);
        } else {
            p.tmRes = p.tmOut;
        }
    } else {
        p.tmOut = p.tmB;
        p.tmRes = p.tmB;
    }
    return "";
}


cudaError_t tc_launch(const TcPlan& p, const ConvArgs& a, int* dbg, cudaStream_t s) {
    TcArgs t;
    t.trace = nullptr;
#ifdef YB_EXPERIMENTS
    static long long* trace_dev = nullptr;
    static const bool trace_on = tune_env("YB_TC_TRACE") && atoi(tune_env("YB_TC_TRACE")) != 0;
    if (trace_on && !trace_dev) cudaMalloc(&trace_dev, 6 * 64 * 4 * sizeof(long long));
    if (trace_on) cudaMemsetAsync(trace_dev, 0, 6 * 64 * 4 * sizeof(long long), s);
    t.trace = trace_on ? trace_dev : nullptr;
#endif
    t.M = p.M;
    t.Ho = a.Ho; t.Wo = a.Wo; t.HoWo = a.Ho * a.Wo;
    t.ks = a.ks; t.stride = a.stride; t.pad = a.pad;
    t.cin_blocks = p.cin_blocks; t.num_kblocks = p.num_kblocks;
    t.kps = p.kps; t.num_iters = p.num_kblocks / p.kps;
    t.BN = p.BN; t.n_tiles = p.n_tiles; t.m_tiles = p.m_tiles;
    t.stages = p.stages; t.tmem_cols = p.tmem_cols;
    t.scale = a.scale; t.bias = a.bias;
    t.tab = nullptr; t.cout_pad = p.cout_pad; t.tab_bytes = p.tab_bytes;
    t.out = a.out; t.out_ld = a.out_ld; t.out_f32 = a.out_f32;
    t.res = static_cast<const __half*>(a.res); t.res_ld = a.res_ld;
    t.leaky = a.leaky; t.upsample = a.upsample;
    t.epi_staged = p.epi_staged; t.ring = p.ring; t.sub_bytes = p.sub_bytes; t.cs = p.cs; t.n_sub = p.n_sub;
    t.has_res = a.res != nullptr;
    t.b_resident = p.b_resident;
    t.srel = p.srel;
    t.epi_split = p.epi_split;
    t.a_lo = (int)a.in_lo; t.out_lo = (int)a.out_lo; t.res_lo = (int)a.res_lo;
    t.split_out = p.split && !a.out_f32;
    t.chunk_pairs = p.chunk_pairs; t.n_chunks = p.n_chunks;
    t.dbg = dbg;
    static PerDeviceOnce attr_once;
    {
        cudaError_t e = attr_once.run([] {
            cudaError_t r = cudaFuncSetAttribute(conv_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
            if (r == cudaSuccess) {     // split mode (YB_MODE_FP32_TC)
                r = cudaFuncSetAttribute(conv_tc_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
                if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_tc_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
                if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_tc_kernel<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
                if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_tc_kernel<128, false, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
                if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_tc_kernel<128, true, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
            }
            return r;
        });
        if (e != cudaSuccess) return e;
    }
    {
        static const bool pdl = !(tune_env("YB_TC_PDL") && atoi(tune_env("YB_TC_PDL")) == 0);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(p.grid);
        cfg.blockDim = dim3(p.split ? kThreadsSplit : kThreads);
        cfg.dynamicSmemBytes = p.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (p.cta2) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 2;
            attr[na].val.clusterDim.y = 1;
            attr[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        cudaError_t e;
        if (p.split && p.n_sub > 2) {
            if (p.cta2) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, true, true, 4>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
            else e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, false, true, 4>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
        } else if (p.split) {
            if (p.cta2) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, true, true>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
            else if (p.swz == 128) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, false, true>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
            else e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, false, true>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
        }
        else if (p.cta2) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, true>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
        else if (p.swz == 128) e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, false>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
        else e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, false>, p.tmA, p.tmB, p.tmOut, p.tmRes, t);
        if (e != cudaSuccess) return e;
    }
#ifdef YB_EXPERIMENTS
    if (trace_on) {                       // debugging aid: dump CTA 0's per-tile time line (cycles)
        static int dumps = 0;
        cudaStreamSynchronize(s);
        long long h[6 * 64 * 4];
        cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
        if (dumps++ % 8 == 7) {
            const long long t0 = h[0];
            fprintf(stderr, "[tc trace] BN=%d kblocks=%d kps=%d stages=%d ring=%d bres=%d n_sub=%d cta2=%d grid=%d\n", p.BN,
                    p.num_kblocks, p.kps, p.stages, p.ring, p.b_resident, p.n_sub, p.cta2, p.grid);
            for (int i = 0; i < 12; ++i)
                fprintf(stderr, "[tc trace] tile %2d  prod start %7lld end %7lld | mma start %7lld tempty-ok %7lld issued %7lld | epi start %7lld tfull-ok %7lld done %7lld\n",
                        i, h[(0 * 64 + i) * 4] - t0, h[(0 * 64 + i) * 4 + 1] - t0, h[(1 * 64 + i) * 4] - t0, h[(1 * 64 + i) * 4 + 1] - t0,
                        h[(1 * 64 + i) * 4 + 2] - t0, h[(2 * 64 + i) * 4] - t0, h[(2 * 64 + i) * 4 + 1] - t0, h[(2 * 64 + i) * 4 + 2] - t0);
            {
                int last = 0;
                for (int i = 0; i < 64; ++i) if (h[(2 * 64 + i) * 4 + 2]) last = i;
                // (the entry stamp is taken before griddepcontrol.wait, so the preceding memset may erase it)
                fprintf(stderr, "[tc trace] kernel: setup-done %7lld pdl-wait-done %7lld | first tile MMAs issued %7lld | last tile (%d) epilogue done %7lld | exit %7lld\n",
                        h[(5 * 64) * 4 + 1] - t0, h[(5 * 64) * 4 + 2] - t0, h[(1 * 64) * 4 + 2] - t0, last,
                        h[(2 * 64 + last) * 4 + 2] - t0, h[(5 * 64) * 4 + 3] - t0);
            }
            for (int i = 0; i < 6; ++i)
                fprintf(stderr, "[tc trace] tile %2d  epilogue sub-tile 0: buffer-ok %7lld tmem-loaded %7lld math+smem %7lld fence %7lld arrived %7lld\n", i,
                        h[(3 * 64 + i) * 4] - t0, h[(3 * 64 + i) * 4 + 1] - t0, h[(3 * 64 + i) * 4 + 2] - t0, h[(3 * 64 + i) * 4 + 3] - t0,
                        h[(4 * 64 + i) * 4] - t0);
        }
    }
#endif
    return cudaGetLastError();
}

}  // namespace yb
