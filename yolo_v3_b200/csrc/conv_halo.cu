// K1h: 3x3 convolution (stride 1, and stride 2 through parity planes) of the wide, shallow layers (Cin = 32 or 64 at 304x304 / 152x152) from a HALO TILE.
//
// conv_tc.cu fetches the A operand of a 3x3 layer with nine im2col TMA loads per tile, i.e. every input pixel travels
// from L2 to shared memory nine times.  For the layers with few channels that traffic, not DRAM and not the tensor
// pipe, is the bound (profiles/README.md: 6-8 TB/s through L2 at 30-47 % tensor-pipe activity).  Here each input
// pixel is staged ONCE: a tile is 3 output rows x 38 output columns, its input patch (5 x 40 pixels, zero-filled
// outside the image by TMA) is one tiled TMA load, and the nine taps are nine views of that patch -- the UMMA
// descriptor of tap (ky, kx) simply starts (ky*40 + kx) pixels into it.  That works because a K-major descriptor
// may start at any row of a TMA-written swizzled tile (base_offset 0; tools/probes/umma_shift_probe.cu).
//
//   * GEMM rows follow the PATCH pitch: row m = r*40 + c (r < 3, c < 40), so that tap (ky, kx) of row m is patch
//     pixel m + ky*40 + kx for every m: one descriptor, 128 consecutive 128-byte (Cin = 64) or 64-byte (Cin = 32)
//     rows.  Rows with c >= 38 or m >= 120 are by-products (they read the neighbour's pixels) and are dropped.
//   * weights: the CTA's whole [BN = 64][9*Cin] slab is resident in shared memory (36 / 72 KB);
//   * epilogue as in conv_tc.cu (TMEM -> scale/bias -> LeakyReLU -> +residual -> fp16 -> swizzled staging -> TMA
//     store), except that the staging row of GEMM row m is the COMPACT index r*38 + c and the residual load / output
//     store are 4-D boxes {64 ch, 38, 3, 1} of the NHWC tensor (rows past the image are clipped by the tensor map).
//
// Same warp roles, barriers, TMEM double buffering, watchdog and programmatic dependent launch as conv_tc.cu.
#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "yb_internal.h"

namespace yb {
namespace {

constexpr int kHP = 40;                       // patch pitch in pixels = tile columns + 2
constexpr int kHC = 38;                       // output columns per tile
constexpr int kHR = 3;                        // output rows per tile
constexpr int kHPatchPix = (kHR + 2) * kHP;   // 200 pixels per TMA load
constexpr int kHSlotPix = 216;                // rows reserved per stage: the last tap reads up to row 2*40+2+127 = 209
// stride 2: the input patch is staged as four parity planes (odd/even input rows x odd/even input columns, each 4 x 40
// pixels, loaded with a traversal stride of 2), so that every tap is again a unit-pitch view: tap ky in {0,2} reads the
// odd-row plane at row offset ky/2, ky = 1 the even-row plane, and likewise for kx.
constexpr int kHPlanePix = 4 * kHP;           // 160 pixels per plane
constexpr int kHSlotPix2 = 4 * kHPlanePix + 48;   // last view: plane 3 + (1*40+1) + 127 = 648 < 688
constexpr int kHValid = kHR * kHC;            // 114 output pixels per tile
constexpr int kHBN = 64;                      // output channels per tile
constexpr int kHThreads = 768;                // warp 0 producer, 2 MMA, 3 TMEM alloc + store issuer, 4 residual, 8-23 epilogue
constexpr int kHEpiWarps = 16;
constexpr int kHMaxStages = 12;
constexpr int kHMaxRing = 4;
constexpr size_t kHSmemBudget = 227 * 1024;
constexpr uint32_t kHStgBytes = 128 * 128;    // one staging sub-tile: 128 rows x 64 fp16

struct HaloArgs {
    int tiles_x, tiles_y, n_tiles, total_tiles;
    int stages, slot_bytes;
    const float* scale; const float* bias;
    int cout_pad, tab_bytes;
    int leaky, has_res, ring;
    int* dbg;
};


// Position of the tiles a CTA visits -- tile = ((img * tiles_y + ty) * tiles_x + tx) * n_tiles + n_tile -- carried
// without divisions: every role decodes its tile once per iteration, and for the epilogue warps that decode sits on
// the per-tile dependent chain that paces these layers (a 32-bit division costs ~150-200 cycles of latency there).
// The grid is a multiple of n_tiles (halo_make_plan), so a CTA's n-tile never changes.
struct TileWalk {
    int n_tile, tx, ty, img;
    int dx, dy, dimg, tiles_x, tiles_y;
    __device__ __forceinline__ TileWalk(const HaloArgs& a, int first, int step) : tiles_x(a.tiles_x), tiles_y(a.tiles_y) {
        n_tile = first % a.n_tiles;
        int mt = first / a.n_tiles;
        tx = mt % tiles_x; mt /= tiles_x;
        ty = mt % tiles_y; img = mt / tiles_y;
        int dm = step / a.n_tiles;
        dx = dm % tiles_x; dm /= tiles_x;
        dy = dm % tiles_y; dimg = dm / tiles_y;
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
    __device__ __forceinline__ int x0() const { return tx * kHC; }
    __device__ __forceinline__ int y0() const { return ty * kHR; }
};
struct RingWalk {
    uint32_t buf = 0, ph = 0;          // slot g % ring and parity (g / ring) & 1 of the g-th tile
    __device__ __forceinline__ void next(uint32_t ring) { if (++buf == ring) { buf = 0; ph ^= 1u; } }
};

__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(kTmaCacheDefault) : "memory");
}

// PAIR (Cout = 128, Cin = 64, stride 1): two CTAs of a cluster run one cta_group::2 UMMA of M = 256 (two spatial tiles,
// one patch per CTA) x N = 128 (each CTA keeps the 64 weight rows of its half resident).  A single CTA with N = 64 pays
// ~76 cycles per M128 x N64 x K16 instruction against 32 of tensor work -- every instruction re-reads its 4 KB of A and
// 2 KB of B from shared memory, 192 B/clk against the 128 B/clk a SM delivers -- and needs two passes over the patch;
// the pair reads 6 KB per 64 tensor cycles.  Tile t of the pair's round is tile 2*p + rank, so the tile walk is the
// same (first = blockIdx.x, step = gridDim.x); an odd tile count leaves rank 1 of the last pair a tile at image index
// B, which TMA zero-fills on load and clips on store.
// (warp index through a shuffle broadcast = uniform role dispatch, as in conv_tc.cu)
template <int SWZ, int STRIDE, bool PAIR>
__global__ void __launch_bounds__(kHThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes, const HaloArgs a) {
    static_assert(!PAIR || (SWZ == 128 && STRIDE == 1), "pair mode: 64-channel stride-1 layers only");
    constexpr int BKE = SWZ / 2;                                // fp16 per pixel = Cin
    constexpr uint32_t B_SLOT = kHBN * SWZ;                     // one tap of this CTA's weight rows: 64 rows x Cin
    constexpr int NCTA = PAIR ? 2 : 1;
    constexpr int NSUB = PAIR ? 2 : 1;                          // 64-channel sub-tiles per tile
    constexpr uint32_t ACC_COLS = kHBN * NSUB;                  // TMEM columns per accumulator = UMMA N
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    // header: full[12] | empty[12] | tfull[2] | tempty[2] | tmem_ptr | sfull[4] | sempty[4] | sready[4] | bres
    const uint32_t full0 = base, empty0 = base + 8 * kHMaxStages;
    const uint32_t tfull0 = base + 16 * kHMaxStages, tempty0 = tfull0 + 16;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 16 * kHMaxStages + 32);
    const uint32_t sfull0 = base + 16 * kHMaxStages + 64, sempty0 = sfull0 + 32, sready0 = sempty0 + 32, bres_bar = sready0 + 32;
    const float* tab = reinterpret_cast<const float*>(gen + 1024);         // scale[cout_pad] | bias[cout_pad]
    const uint32_t stg0 = base + 1024 + (uint32_t)a.tab_bytes;
    const uint32_t bres0 = stg0 + (uint32_t)a.ring * kHStgBytes;
    const uint32_t stage0 = bres0 + 9 * B_SLOT;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int tile_first = blockIdx.x, tile_step = gridDim.x;
    // both CTAs of a pair run the same number of rounds: the bound is tested on the leader's tile
    const int tile_end = a.total_tiles + (int)rank;

    for (int i = threadIdx.x; i < a.cout_pad; i += kHThreads) {
        float* t = reinterpret_cast<float*>(gen + 1024);
        t[i] = __ldg(a.scale + i);
        t[a.cout_pad + i] = __ldg(a.bias + i);
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmIn); prefetch_tmap(&tmB); prefetch_tmap(&tmOut);
        if (a.has_res) prefetch_tmap(&tmRes);
    }
    if (warp == 2 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, kHEpiWarps * NCTA); }
        for (int i = 0; i < kHMaxRing; ++i) { mbar_init(sfull0 + 8 * i, 1); mbar_init(sempty0 + 8 * i, 1); mbar_init(sready0 + 8 * i, kHEpiWarps); }
        mbar_init(bres_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        if constexpr (PAIR) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 2 * ACC_COLS);
        else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 2 * ACC_COLS);
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // peer barriers must exist before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    // the resident weight slab (one elected thread of warp 0)
    auto load_slab = [&]() {
        // the grid is a multiple of n_tiles: a CTA's n-tile never changes; pair mode: this CTA's half of the 128 rows
        const int n0 = PAIR ? (int)rank * kHBN : (tile_first % a.n_tiles) * kHBN;
        if (leader) mbar_arrive_expect_tx(bres_bar, 9 * B_SLOT * NCTA);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            if constexpr (PAIR) tma_load_2d_pair(&tmB, bres0 + t * B_SLOT, bres_bar, t * BKE, n0);
            else tma_load_2d(&tmB, bres0 + t * B_SLOT, bres_bar, t * BKE, n0);
        }
    };
    pdl_wait_prior();

    if (warp == 0) {
        // ===== TMA producer: the resident weight slab once, then one patch per tile =====
        if (tile_first < tile_end && elect_one()) load_slab();
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        TileWalk t(a, tile_first, tile_step);
        for (int tile = tile_first; tile < tile_end; tile += tile_step, t.next()) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1, a.dbg, 0, stage);
            if (elect_one()) {
                const uint32_t slot = stage0 + stage * a.slot_bytes, fb = full0 + 8 * stage;
                if (STRIDE == 1) {
                    if (leader) mbar_arrive_expect_tx(fb, (uint32_t)kHPatchPix * SWZ * NCTA);
                    if constexpr (PAIR) tma_load_4d_pair(&tmIn, slot, fb, 0, t.x0() - 1, t.y0() - 1, t.img);
                    else tma_load_4d(&tmIn, slot, fb, 0, t.x0() - 1, t.y0() - 1, t.img);
                } else {
                    // planes: 0 = odd rows / odd cols, 1 = odd rows / even cols, 2 = even rows / odd cols, 3 = even / even
                    mbar_arrive_expect_tx(fb, 4u * kHPlanePix * SWZ);
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl)
                        tma_load_4d(&tmIn, slot + pl * kHPlanePix * SWZ, fb, 0, 2 * t.x0() - 1 + (pl & 1), 2 * t.y0() - 1 + (pl >> 1), t.img);
                }
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 2) {
        // ===== MMA issuer (pair mode: the leader's warp issues for both CTAs): nine taps = nine shifted views of the patch =====
        if (leader) {
            const uint32_t idesc = make_idesc((int)ACC_COLS, 128 * NCTA);
            int stage = 0;
            uint32_t phase = 0, acc = 0, acc_phase = 0;
            if (tile_first < tile_end) mbar_wait(bres_bar, 0, a.dbg, 1, 600);
            const uint64_t bdesc0 = make_smem_desc<SWZ>(bres0);
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, a.dbg, 1, 100 + (int)acc);
                mbar_wait(full0 + 8 * stage, phase, a.dbg, 1, stage);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                    const uint64_t adesc0 = make_smem_desc<SWZ>(stage0 + stage * a.slot_bytes);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int ky = t / 3, kx = t % 3;
                        // stride 1: pixel (ky, kx) of the patch; stride 2: plane (ky odd?, kx odd?) at offset (ky/2, kx/2)
                        const int pix = STRIDE == 1 ? ky * kHP + kx
                                                    : (((ky & 1) << 1) | (kx & 1)) * kHPlanePix + (ky >> 1) * kHP + (kx >> 1);
                        const uint64_t ad = adesc0 + (uint64_t)((pix * SWZ) >> 4);
                        const uint64_t bd = bdesc0 + (uint64_t)(t * (B_SLOT >> 4));
#pragma unroll
                        for (int k = 0; k < BKE / 16; ++k) {
                            if constexpr (PAIR) umma_f16_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (t | k) != 0);
                            else umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (t | k) != 0);
                        }
                    }
                    if constexpr (PAIR) { umma_commit_pair(empty0 + 8 * stage); umma_commit_pair(tfull0 + 8 * acc); }
                    else { umma_commit(empty0 + 8 * stage); umma_commit(tfull0 + 8 * acc); }
                }
                __syncwarp();
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        // ===== residual prefetch into the staging ring (compact 3 x 38 pixel rows, 64 channels per slot) =====
        if (lane == 0 && a.has_res) {
            TileWalk t(a, tile_first, tile_step);
            RingWalk rw;
            for (int tile = tile_first; tile < tile_end; tile += tile_step, t.next()) {
#pragma unroll 1
                for (int j = 0; j < NSUB; ++j, rw.next((uint32_t)a.ring)) {
                    const uint32_t buf = rw.buf, ph = rw.ph;
                    mbar_wait(sempty0 + 8 * buf, ph ^ 1, a.dbg, 3, 300 + (int)buf);
                    mbar_arrive_expect_tx(sfull0 + 8 * buf, (uint32_t)kHValid * 128u);
                    tma_load_4d(&tmRes, stg0 + buf * kHStgBytes, sfull0 + 8 * buf, (t.n_tile * NSUB + j) * kHBN, t.x0(), t.y0(), t.img);
                }
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===== store issuer =====
        if (lane == 0) {
            TileWalk t(a, tile_first, tile_step);
            RingWalk rw;
            uint32_t prev = 0;
            bool first = true;
            for (int tile = tile_first; tile < tile_end; tile += tile_step, t.next()) {
#pragma unroll 1
                for (int j = 0; j < NSUB; ++j, rw.next((uint32_t)a.ring)) {
                    const uint32_t buf = rw.buf, ph = rw.ph;
                    mbar_wait(sready0 + 8 * buf, ph, a.dbg, 4, 700 + (int)buf);
                    tma_store_4d(&tmOut, stg0 + buf * kHStgBytes, (t.n_tile * NSUB + j) * kHBN, t.x0(), t.y0(), t.img);
                    tma_store_commit();
                    if (!first) {                             // the previous store has finished reading its buffer
                        tma_store_wait_read<1>();
                        mbar_arrive(sempty0 + 8 * prev);
                    }
                    first = false;
                    prev = buf;
                }
            }
            tma_store_wait_all();
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ===== epilogue: TMEM lane m = GEMM row m = patch-pitch pixel (r, c); staging row = r*38 + c =====
        const int q = warp & 3, part = (warp - 8) >> 2;        // TMEM lane quarter, 16-column group of the 64
        const int m = q * 32 + lane;
        const int r = m / kHP, c = m - r * kHP;
        const bool valid = r < kHR && c < kHC;
        const int mp = r * kHC + c;                             // compact row (only used when valid)
        const int xr = mp & 7;
        uint32_t acc = 0, acc_phase = 0;
        RingWalk rw;
        const int n0 = PAIR ? 0 : (tile_first % a.n_tiles) * kHBN;   // the grid is a multiple of n_tiles: constant per CTA
        // ... and so are this thread's sixteen channels (single-CTA mode): their scale / bias pairs live in registers for
        // the whole kernel; pair mode (two sub-tiles = 32 channels per thread) reads the shared-memory table
        float4 sc4[4], bi4[4];
        if constexpr (!PAIR) {
            const float4* sc = reinterpret_cast<const float4*>(tab + n0 + part * 16);
            const float4* bi = reinterpret_cast<const float4*>(tab + a.cout_pad + n0 + part * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sc4[i] = sc[i]; bi4[i] = bi[i]; }
        }
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            mbar_wait(tfull0 + 8 * acc, acc_phase, a.dbg, 2, 200 + (int)acc);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < NSUB; ++j, rw.next((uint32_t)a.ring)) {
                const uint32_t buf = rw.buf, ph = rw.ph;
                uint32_t r0[16];
                tmem_ld16(tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * kHBN + part * 16), r0);
                if (a.has_res) mbar_wait(sfull0 + 8 * buf, ph, a.dbg, 2, 400 + (int)buf);
                else mbar_wait(sempty0 + 8 * buf, ph ^ 1, a.dbg, 2, 500 + (int)buf);
                tmem_ld_wait();
                if (j == NSUB - 1) {                            // accumulator drained into registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR) mbar_arrive_leader(tempty0 + 8 * acc); else mbar_arrive(tempty0 + 8 * acc);
                    }
                }
                if (valid) {
                    uint8_t* srow = gen + (stg0 - base) + buf * kHStgBytes + (uint32_t)mp * 128u;
                    float v[16];
                    if constexpr (PAIR) {
                        const float4* sc = reinterpret_cast<const float4*>(tab + j * kHBN + part * 16);
                        const float4* bi = reinterpret_cast<const float4*>(tab + a.cout_pad + j * kHBN + part * 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { sc4[i] = sc[i]; bi4[i] = bi[i]; }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 s4 = sc4[i], b4 = bi4[i];
                        v[4 * i + 0] = fmaf(__uint_as_float(r0[4 * i + 0]), s4.x, b4.x);
                        v[4 * i + 1] = fmaf(__uint_as_float(r0[4 * i + 1]), s4.y, b4.y);
                        v[4 * i + 2] = fmaf(__uint_as_float(r0[4 * i + 2]), s4.z, b4.z);
                        v[4 * i + 3] = fmaf(__uint_as_float(r0[4 * i + 3]), s4.w, b4.w);
                    }
                    if (a.leaky) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = leaky(v[i]);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4* p = reinterpret_cast<uint4*>(srow + (((part * 2 + h) ^ xr) << 4));
                        if (a.has_res) {
                            const uint4 rr = *p;
                            const __half2* hh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(hh[i]);
                                v[8 * h + 2 * i] += f.x;
                                v[8 * h + 2 * i + 1] += f.y;
                            }
                        }
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int i = 0; i < 4; ++i) ph2[i] = __floats2half2_rn(v[8 * h + 2 * i], v[8 * h + 2 * i + 1]);
                        *p = pk;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(sready0 + 8 * buf);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // the peer may still signal this CTA's barriers / read its smem
    if (warp == 3) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 2 * ACC_COLS);
        else tmem_dealloc(tmem_base, 2 * ACC_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;

std::string halo_err(const char* what, CUresult r) { return std::string(what) + " failed with CUresult " + std::to_string((int)r); }

}  // namespace

bool halo_supported(const ConvArgs& a) {
    static const bool enabled = !(tune_env("YB_HALO") && atoi(tune_env("YB_HALO")) == 0);
    if (!enabled) return false;
    if (a.ks != 3 || a.pad != 1 || a.upsample || a.out_f32) return false;
    if (a.stride == 1) {
        if (a.Cin != 32 && a.Cin != 64) return false;
    } else if (a.stride == 2) {
        // four parity planes per stage: fits next to the resident weights only for 64-byte pixels
        static const bool s2 = !(tune_env("YB_HALO_S2") && atoi(tune_env("YB_HALO_S2")) == 0);
        if (!s2 || a.Cin != 32 || a.H % 2 || a.W % 2) return false;
    } else {
        return false;
    }
    if (a.Cout % kHBN != 0 || a.Cout > 128) return false;
    // any output width: the last column tile may be partial -- its patch load zero-fills and its store box is clipped by the
    // tensor map, exactly like the last row strip -- as long as at least 80 % of the tile columns are real outputs
    // (208 -> 6 tiles, 91 %; 128 -> 4, 84 %; 104 -> 3, 91 %; 64 -> 2, 84 %; multiples of 38: 100 %)
    const int tx = (a.Wo + kHC - 1) / kHC;
    if (a.Wo * 5 < tx * kHC * 4) return false;
    // (the scheme trades 11 % of the tensor work -- 114 useful rows of 128 -- for ~5x less L2->SM traffic; with these
    // channel counts the im2col kernel is traffic-bound, so every eligible shape takes it)
    return !(a.in_ld % 8 || a.out_ld % 8 || (a.res && a.res_ld % 8));
}

std::string halo_make_plan(HaloPlan& p, const ConvArgs& a, const __half* w16, int cout_pad, int K, int num_sms) {
    if (!g_enc) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return "cuTensorMapEncodeTiled not available from the driver";
        g_enc = reinterpret_cast<EncodeTiledFn>(f);
    }
    p.swz = a.Cin == 32 ? 64 : 128;
    p.cout_pad = cout_pad;
    p.stride = a.stride;
    p.tiles_x = (a.Wo + kHC - 1) / kHC;
    p.tiles_y = (a.Ho + kHR - 1) / kHR;
    // Cout = 128 with 128-byte pixels: CTA pairs, all 128 channels in one UMMA (see the kernel's header comment)
    p.pair = p.swz == 128 && a.stride == 1 && a.Cout == 2 * kHBN && num_sms % 2 == 0;
    if (const char* e = tune_env("YB_HALO_PAIR")) p.pair = p.pair && atoi(e) != 0;
    p.n_tiles = p.pair ? 1 : a.Cout / kHBN;
    p.total_tiles = a.B * p.tiles_x * p.tiles_y * p.n_tiles;
    p.ring = (a.res || p.pair) ? 4 : 2;           // slots hold 64 channels: a pair-mode tile takes two
    p.tab_bytes = (int)(((size_t)2 * cout_pad * sizeof(float) + 1023) & ~(size_t)1023);
    p.slot_bytes = (int)(((size_t)(a.stride == 1 ? kHSlotPix : kHSlotPix2) * p.swz + 1023) & ~(size_t)1023);
    const size_t fixed = 1024 + 1024 + p.tab_bytes + (size_t)p.ring * kHStgBytes + (size_t)9 * kHBN * p.swz;
    p.stages = (int)std::min<size_t>(kHMaxStages, (kHSmemBudget - fixed) / p.slot_bytes);
    if (p.stages < 2) return "not enough shared memory for two patch stages";
    p.smem = fixed + (size_t)p.stages * p.slot_bytes;
    if (p.pair) {
        p.grid = std::min(2 * ((p.total_tiles + 1) / 2), num_sms);
        p.grid -= p.grid % 2;
    } else {
        p.grid = std::min(p.total_tiles, num_sms);
        p.grid -= p.grid % p.n_tiles;
    }
    if (p.grid <= 0) return "grid too small for the tile split";
    const CUtensorMapSwizzle swz = p.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const cuuint32_t es4[4] = {1, 1, 1, 1};
    const cuuint32_t es2[2] = {1, 1};
    {   // input patch: NHWC as (C, W, H, N); box = Cin x 40 x 5 x 1, out-of-image pixels read as zero (the conv padding)
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t st[3] = {(cuuint64_t)a.in_ld * 2, (cuuint64_t)a.W * a.in_ld * 2, (cuuint64_t)a.H * a.W * a.in_ld * 2};
        cuuint32_t box[4] = {(cuuint32_t)a.Cin, (cuuint32_t)kHP, (cuuint32_t)(kHR + 2), 1};
        cuuint32_t es_in[4] = {1, 1, 1, 1};
        if (a.stride == 2) {        // one parity plane: 40 x 4 pixels traversed with stride 2 (box extent = 2x the count)
            box[1] = 2 * kHP; box[2] = 2 * 4;
            es_in[1] = 2; es_in[2] = 2;
        }
        CUresult r = g_enc(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a.in), dims, st, box, es_in,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return halo_err("cuTensorMapEncodeTiled(halo input)", r);
    }
    {   // weights [cout_pad][K], one tap x 64 rows per box
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout_pad};
        cuuint64_t st[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)a.Cin, (cuuint32_t)kHBN};
        CUresult r = g_enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w16), dims, st, box, es2,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return halo_err("cuTensorMapEncodeTiled(halo weights)", r);
    }
    {   // output / residual: box = 64 ch x 38 x 3 x 1, 128B swizzle (staging rows are 128 bytes)
        cuuint64_t dims[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)a.B};
        cuuint64_t st[3] = {(cuuint64_t)a.out_ld * 2, (cuuint64_t)a.Wo * a.out_ld * 2, (cuuint64_t)a.Ho * a.Wo * a.out_ld * 2};
        cuuint32_t box[4] = {(cuuint32_t)kHBN, (cuuint32_t)kHC, (cuuint32_t)kHR, 1};
        CUresult r = g_enc(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.out, dims, st, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return halo_err("cuTensorMapEncodeTiled(halo output)", r);
        if (a.res) {
            cuuint64_t rst[3] = {(cuuint64_t)a.res_ld * 2, (cuuint64_t)a.Wo * a.res_ld * 2, (cuuint64_t)a.Ho * a.Wo * a.res_ld * 2};
            r = g_enc(&p.tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a.res), dims, rst, box, es4,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return halo_err("cuTensorMapEncodeTiled(halo residual)", r);
        } else {
            p.tmRes = p.tmOut;
        }
    }
    return "";
}

cudaError_t halo_launch(const HaloPlan& p, const ConvArgs& a, int* dbg, cudaStream_t s) {
    HaloArgs h;
    h.tiles_x = p.tiles_x; h.tiles_y = p.tiles_y; h.n_tiles = p.n_tiles; h.total_tiles = p.total_tiles;
    h.stages = p.stages; h.slot_bytes = p.slot_bytes;
    h.scale = a.scale; h.bias = a.bias;
    h.cout_pad = p.cout_pad; h.tab_bytes = p.tab_bytes;
    h.leaky = a.leaky; h.has_res = a.res != nullptr; h.ring = p.ring;
    h.dbg = dbg;
    static PerDeviceOnce attr_once;
    {
        cudaError_t e = attr_once.run([] {
            cudaError_t r = cudaFuncSetAttribute(conv_halo_kernel<128, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmemBudget);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_halo_kernel<128, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmemBudget);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_halo_kernel<64, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmemBudget);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(conv_halo_kernel<64, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmemBudget);
            return r;
        });
        if (e != cudaSuccess) return e;
    }
    static const bool pdl = !(tune_env("YB_TC_PDL") && atoi(tune_env("YB_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(kHThreads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (p.pair) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e;
    if (p.stride == 2) e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<64, 2, false>, p.tmIn, p.tmB, p.tmOut, p.tmRes, h);
    else if (p.pair) e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<128, 1, true>, p.tmIn, p.tmB, p.tmOut, p.tmRes, h);
    else if (p.swz == 128) e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<128, 1, false>, p.tmIn, p.tmB, p.tmOut, p.tmRes, h);
    else e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<64, 1, false>, p.tmIn, p.tmB, p.tmOut, p.tmRes, h);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace yb
