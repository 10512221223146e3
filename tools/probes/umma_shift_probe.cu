// Hardware probe (not part of the product): can a tcgen05 K-major shared-memory descriptor start at an arbitrary
// row of a swizzled tile?  The early, L2-bound 3x3 layers would stop re-reading every input pixel nine times if
// one halo tile in shared memory could serve all taps through descriptors shifted by whole pixels (rows).
//
// For shift k = 0..8 rows the kernel multiplies rows [k, k+128) of a TMA-loaded A tile (136 rows x 64 fp16, 128B
// swizzle -- and a second configuration with 32 fp16 = 64-byte rows, 64B swizzle) by a 64-row B tile and the host
// compares with the exact integer result, for two ways of filling the descriptor's base_offset field
// (0, and (start_address >> 7) & 7 as the PTX ISA describes for unaligned pattern starts).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/umma_shift_probe.bin tools/probes/umma_shift_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int kRowsA = 136, kM = 128, kN = 64, kShifts = 9;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a lost arrival must not hang the GPU
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) if (clock64() - t0 > 2000000000LL) return false;
    return true;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}

// K-major descriptor; SWZ = swizzle span in bytes (128 or 64); base_offset in bits [49,52)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int swz, uint32_t base_offset) {
    const uint64_t sbo = (uint64_t)(8 * swz) >> 4, layout = swz == 128 ? 2 : 4;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) | (layout << 61);
}

// out[shift][variant][m][n] fp32; status[0] = 1 on a barrier timeout
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                               int swz, float* out, int* status) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t bar_load = base, bar_mma = base + 8;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 16);
    const uint32_t sA = base + 1024;                                    // 136 rows x swz bytes (<= 17408 B)
    const uint32_t sB = sA + 18 * 1024;                                 // 64 rows x swz bytes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kelems = swz / 2;                                          // fp16 per row: 64 or 32
    if (threadIdx.x == 0) {
        mbar_init(bar_load, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(tmem_ptr))), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_load, (uint32_t)(kRowsA + kN) * swz);
        tma_load_2d(&tmA, sA, bar_load, 0, 0);
        tma_load_2d(&tmB, sB, bar_load, 0, 0);
    }
    bool ok = mbar_wait(bar_load, 0);
    uint32_t phase = 0;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    for (int k = 0; k < kShifts && ok; ++k) {
        for (int variant = 0; variant < 2 && ok; ++variant) {
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = sA + (uint32_t)k * swz;
                const uint32_t bo = variant ? ((a0 >> 7) & 7u) : 0u;
                const uint64_t ad = make_desc(a0, swz, bo), bd = make_desc(sB, swz, 0);
                for (int kk = 0; kk < kelems / 16; ++kk) umma_f16(tmem, ad + 2 * kk, bd + 2 * kk, idesc, kk != 0);
                umma_commit(bar_mma);
            }
            ok = mbar_wait(bar_mma, phase);
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ok) {
                const int m = warp * 32 + lane;
                float* o = out + (((size_t)k * 2 + variant) * kM + m) * kN;
                for (int c0 = 0; c0 < kN; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    for (int i = 0; i < 16; ++i) o[c0 + i] = __uint_as_float(r[i]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    }
    if (!ok && threadIdx.x == 0) status[0] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}

// ---- second question: un-swizzled 16-byte rows with a free leading-dimension offset ---------------------------
// For a Cin = 3 stem the patch can be kept as [pixel][8 fp16] (16 bytes per pixel).  In the un-swizzled K-major
// layout a K = 16 MMA reads, for row m, one 16-byte chunk at start + m*16 and a second one LBO bytes further: if LBO
// may be ANY multiple of 16 bytes, the second chunk can be "the same pixel array seen from another filter tap", and
// one MMA covers two taps.  out2[case][m][n]; cases: (row shift of chunk 0, pixel distance of chunk 1).
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
constexpr int kNsRows = 384, kNsCases = 6;
__constant__ int c_ns_shift[kNsCases] = {0, 1, 3, 0, 5, 40};
__constant__ int c_ns_dist[kNsCases] = {1, 1, 38, 40, 77, 82};

__global__ void __launch_bounds__(128, 1) probe_noswz(const __half* __restrict__ A, const __half* __restrict__ B, float* out, int* status) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t bar_mma = base + 8;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 16);
    const uint32_t sA = base + 1024;                                    // [384 rows][8 fp16]
    const uint32_t sB = sA + 8 * 1024;                                  // [2 chunks][64 rows][8 fp16]: LBO = 1024, SBO = 128
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kNsRows; i += 128) reinterpret_cast<uint4*>(gen + 1024)[i] = reinterpret_cast<const uint4*>(A)[i];
    for (int i = threadIdx.x; i < 2 * kN; i += 128) reinterpret_cast<uint4*>(gen + 1024 + 8 * 1024)[i] = reinterpret_cast<const uint4*>(B)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(tmem_ptr))), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_ptr;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    bool ok = true;
    uint32_t phase = 0;
    for (int cs = 0; cs < kNsCases && ok; ++cs) {
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = make_desc_noswz(sA + (uint32_t)c_ns_shift[cs] * 16u, (uint32_t)c_ns_dist[cs] * 16u, 128u);
            const uint64_t bd = make_desc_noswz(sB, 1024u, 128u);
            umma_f16(tmem, ad, bd, idesc, 0);
            umma_commit(bar_mma);
        }
        ok = mbar_wait(bar_mma, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            const int m = warp * 32 + lane;
            float* o = out + ((size_t)cs * kM + m) * kN;
            for (int c0 = 0; c0 < kN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int i = 0; i < 16; ++i) o[c0 + i] = __uint_as_float(r[i]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (!ok && threadIdx.x == 0) status[0] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}

int run_noswz() {
    std::vector<__half> hA((size_t)kNsRows * 8), hB((size_t)2 * kN * 8);
    std::vector<int> iA(hA.size()), iB(hB.size());
    srand(77);
    for (size_t i = 0; i < hA.size(); ++i) { iA[i] = rand() % 7 - 3; hA[i] = __float2half((float)iA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { iB[i] = rand() % 5 - 2; hB[i] = __float2half((float)iB[i]); }
    __half *dA, *dB; float* dO; int* dS;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dO, sizeof(float) * kNsCases * kM * kN)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xFF, sizeof(float) * kNsCases * kM * kN)); CK(cudaMemset(dS, 0, 4));
    CK(cudaFuncSetAttribute(probe_noswz, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    probe_noswz<<<1, 128, 32 * 1024>>>(dA, dB, dO, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("no-swizzle probe: kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    int st = 0; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    std::vector<float> hO((size_t)kNsCases * kM * kN);
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    const int shift[kNsCases] = {0, 1, 3, 0, 5, 40}, dist[kNsCases] = {1, 1, 38, 40, 77, 82};
    printf("un-swizzled 16-byte rows, K = 16 = [row m+shift | row m+shift+dist], barrier timeout: %d\n", st);
    for (int cs = 0; cs < kNsCases; ++cs) {
        long bad = 0;
        for (int m = 0; m < kM; ++m)
            for (int n = 0; n < kN; ++n) {
                int ref = 0;
                for (int c = 0; c < 8; ++c) {
                    ref += iA[(size_t)(m + shift[cs]) * 8 + c] * iB[(size_t)(0 * kN + n) * 8 + c];
                    ref += iA[(size_t)(m + shift[cs] + dist[cs]) * 8 + c] * iB[(size_t)(1 * kN + n) * 8 + c];
                }
                if (hO[((size_t)cs * kM + m) * kN + n] != (float)ref) ++bad;
            }
        printf("  start row %2d, LBO = %2d rows: %s (%ld / %d wrong)\n", shift[cs], dist[cs], bad ? "MISMATCH" : "exact", bad, kM * kN);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dS);
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int run(int swz, EncodeTiledFn enc) {
    const int ke = swz / 2;
    std::vector<__half> hA((size_t)kRowsA * ke), hB((size_t)kN * ke);
    std::vector<int> iA(hA.size()), iB(hB.size());
    srand(1234 + swz);
    for (size_t i = 0; i < hA.size(); ++i) { iA[i] = rand() % 7 - 3; hA[i] = __float2half((float)iA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { iB[i] = rand() % 5 - 2; hB[i] = __float2half((float)iB[i]); }
    __half *dA, *dB; float* dO; int* dS;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dO, sizeof(float) * kShifts * 2 * kM * kN)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xFF, sizeof(float) * kShifts * 2 * kM * kN)); CK(cudaMemset(dS, 0, 4));
    CUtensorMap tmA, tmB;
    const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint32_t es[2] = {1, 1};
    {
        cuuint64_t dims[2] = {(cuuint64_t)ke, (cuuint64_t)kRowsA}; cuuint64_t st[1] = {(cuuint64_t)ke * 2}; cuuint32_t box[2] = {(cuuint32_t)ke, (cuuint32_t)kRowsA};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)ke, (cuuint64_t)kN}; cuuint64_t st[1] = {(cuuint64_t)ke * 2}; cuuint32_t box[2] = {(cuuint32_t)ke, (cuuint32_t)kN};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    probe<<<1, 128, 64 * 1024>>>(tmA, tmB, swz, dO, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swizzle %dB: kernel failed: %s\n", swz, cudaGetErrorString(e)); return 1; }
    int st = 0; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    std::vector<float> hO((size_t)kShifts * 2 * kM * kN);
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    printf("swizzle %dB (rows of %d fp16), barrier timeout: %d\n", swz, ke, st);
    for (int k = 0; k < kShifts; ++k) {
        for (int v = 0; v < 2; ++v) {
            long bad = 0;
            for (int m = 0; m < kM; ++m)
                for (int n = 0; n < kN; ++n) {
                    int ref = 0;
                    for (int c = 0; c < ke; ++c) ref += iA[(size_t)(m + k) * ke + c] * iB[(size_t)n * ke + c];
                    if (hO[(((size_t)k * 2 + v) * kM + m) * kN + n] != (float)ref) ++bad;
                }
            printf("  shift %d rows, base_offset %s: %s (%ld / %d wrong)\n", k, v ? "(addr>>7)&7" : "0          ", bad ? "MISMATCH" : "exact", bad, kM * kN);
        }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dS);
    return 0;
}

int main() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || !f) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
    int rc = run(128, enc);
    rc |= run(64, enc);
    rc |= run_noswz();
    return rc;
}
