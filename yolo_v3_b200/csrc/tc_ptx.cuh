// PTX wrappers shared by the tensor-core convolution kernels (conv_tc.cu, conv_halo.cu): mbarrier, TMA, tcgen05.
// Everything is `static` to the including translation unit (anonymous namespace).
#pragma once

#include "yb_internal.h"

namespace yb {
namespace {

constexpr long long kWatchdogCycles = 4000000000LL;

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch: let the next kernel of the stream be launched while this one runs, and
// wait for the previous one (completion + memory flush) before touching anything it produced.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of a converged warp.  The producer and MMA loops are executed by their WHOLE warp with
// warp-uniform values and only the asynchronous instructions are guarded by this predicate: that lets
// ptxas keep coordinates / descriptors in uniform registers instead of broadcasting them lane by lane.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

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
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {   // non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrival must not hang the GPU.  On timeout the waiter records who it is in
// host-visible memory and traps; the host sees a launch failure with the diagnostic attached.  The
// polling loop lives out of line so that the producer / MMA loops stay a few instructions long: those
// loops run on ONE thread each and their instruction latency, not bandwidth, paces the whole pipeline.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* dbg, int role, int which) {
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
// INL: the polling loop is inlined.  Kernels that re-partition the register file with setmaxnreg must not call out-of-line
// functions from the re-partitioned regions (ptxas fails the register allocation of such a kernel, C7600).
template <bool INL = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* dbg, int role, int which) {
    if (mbar_try_wait(bar, parity)) return;
    if constexpr (INL) {
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
    } else {
        mbar_wait_slow(bar, parity, dbg, role, which);
    }
}
// Same, for waits that are long by construction (the sixteen epilogue warps waiting for the next accumulator while the
// main loop runs): back off between probes instead of spinning -- the step runs at the board's power cap, and warps
// that poll cost issue slots and energy.
__device__ __noinline__ void mbar_wait_relaxed_slow(uint32_t bar, uint32_t parity, int* dbg, int role, int which, unsigned sleep_ns) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
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
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, int* dbg, int role, int which, int sleep_ns) {
    if (mbar_try_wait(bar, parity)) return;
    if (sleep_ns > 0) mbar_wait_relaxed_slow(bar, parity, dbg, role, which, (unsigned)sleep_ns);
    else mbar_wait_slow(bar, parity, dbg, role, which);
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
// L2 prefetch of a box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_im2col(const CUtensorMap* tm, int c, int w, int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.im2col [%0, {%1, %2, %3, %4}], {%5, %6};"
                 ::"l"(tm), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster share one 256-row UMMA; loads of both
// CTAs complete on the leader's (rank 0) mbarrier, whose shared::cluster address is the local one with
// the peer bit (bit 24) cleared.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
constexpr uint64_t kTmaCacheDefault = 0x1000000000000000ull;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "l"(kTmaCacheDefault) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_pair(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c, int w, int h, int n,
                                                     uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h), "l"(kTmaCacheDefault)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the barrier at this offset in CTA rank 0
    asm volatile(
        "{\n\t.reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t}"
        ::"r"(bar) : "memory");
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
// same, accumulate flag known at compile time (no setp in the issue loop)
template <int ACC>
__device__ __forceinline__ void umma_f16_imm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC) : "memory");
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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
// non-swizzled K-major operand: start address, LBO (between the two K halves) and SBO (between 8-row groups), all >> 4
__device__ __forceinline__ uint64_t make_desc_plain(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)lbo16 << 16) | ((uint64_t)sbo16 << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits [4,6) = 1), A/B fp16
// (0), both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n, int m = 128) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : v * kLeaky; }

// A round-robin slot counter: index g % N and parity (g / N) & 1 of the g-th use, advanced by a constant step < 2N
template <int N>
struct Slot {
    uint32_t i, ph;
    __device__ __forceinline__ explicit Slot(uint32_t first) : i(first % N), ph((first / N) & 1u) {}
    __device__ __forceinline__ void advance(uint32_t step) { i += step; if (i >= (uint32_t)N) { i -= N; ph ^= 1u; } }
};

// element of a planar image patch staged in shared memory (fp32 or fp16 images)
__device__ __forceinline__ float raw_ld(const float* p) { return *p; }
__device__ __forceinline__ float raw_ld(const __half* p) { return __half2float(*p); }

}  // namespace
}  // namespace yb
