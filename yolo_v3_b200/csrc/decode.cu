// K2: anchor decode of the three head maps -- YoloLayer.forward with target=None
// (reference yololayer.py:31-59, 97-105), all three scales in one launch, written straight into the
// concatenated [B, N, 5+C] tensor the callers build with torch.cat (test.py:36).
//
//   channel = a*(5+C) + attr          (view at yololayer.py:42)
//   row     = (y*w + x)*3 + a         (permute at yololayer.py:104)
//   bx = (sigmoid(tx) + x) * stride   by likewise          (yololayer.py:57, :98)
//   bw = (exp(tw) * (anchor_w/stride)) * stride            (yololayer.py:59, :98)
//   conf, cls = sigmoid(t)
//
// Full-precision expf and IEEE division (no -use_fast_math); every product is a separate rounding
// as in the reference.  HBM-bound: reads 4 B and writes 4 B per element, both fully coalesced.
#include <algorithm>
#include <cstdlib>

#include "yb_internal.h"

namespace yb {
namespace {

struct DecodeParams {
    DecodeScale sc[3];
    long cells_before[4];   // prefix sum of B*h*w per scale
    int B, attrs, n_total;
};

__device__ __forceinline__ float sigmoidf_rn(float t) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-t))); }

// One exponential and one IEEE division per element, no divergent branches: the per-attribute
// differences are folded into (use_exp, offset, mul).  (s + 0.0f) * 1.0f == s exactly, so conf/cls
// come out bit-identical to a plain sigmoid.
__device__ __forceinline__ float decode_one(const DecodeScale& s, int a, int attr, int x, int y, float t) {
    const bool wh = attr == 2 || attr == 3;
    const float e = expf(wh ? t : -t);
    const float sig = __frcp_rn(__fadd_rn(1.0f, e));     // 1 / (1 + e), correctly rounded = __fdiv_rn(1, 1 + e) bit for bit
    const float off = attr == 0 ? (float)x : (attr == 1 ? (float)y : 0.f);
    const float mul = attr < 2 ? s.stride : 1.0f;
    const float r_sig = __fmul_rn(__fadd_rn(sig, off), mul);
    const float anc = attr == 2 ? s.aw[a] : s.ah[a];
    const float r_wh = __fmul_rn(__fmul_rn(e, anc), s.stride);
    return wh ? r_wh : r_sig;
}

// NHWC (engine-internal) input: element (b, p, c) -> out[b][row_off*attrs + p*3*attrs + c]: a flat
// elementwise map.  A block owns kCellsPerBlock consecutive cells of one image and scale (so the cell ->
// (x, y) arithmetic is two scalar divisions per block).  64 threads cover one cell with a 16-byte load of
// four consecutive channels each (the head maps are padded to a multiple of 16 channels, so the loads are
// aligned); the four 64-thread groups of a block take every fourth cell, with four such loads in flight
// per thread -- the kernel is latency-bound otherwise.  Stores are scalar (the 3*(5+C)-float output rows
// are only 4-byte aligned) but consecutive threads still write consecutive addresses.  (Staging the block's output in
// shared memory for aligned 16-byte stores is slower, see decode_nhwc_staged_kernel below.)
constexpr int kCellsPerBlock = 32;

__global__ void __launch_bounds__(256) decode_nhwc_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ det,
                                                          int blocks_s0, int blocks_s1) {
    const int ch = 3 * P.attrs;
    int blk = blockIdx.x;
    const int si = blk >= blocks_s0 + blocks_s1 ? 2 : (blk >= blocks_s0 ? 1 : 0);
    blk -= si == 2 ? blocks_s0 + blocks_s1 : (si == 1 ? blocks_s0 : 0);
    const DecodeScale& s = P.sc[si];
    const int hw = s.h * s.w;
    const int per_img = (hw + kCellsPerBlock - 1) / kCellsPerBlock;
    const int b = blk / per_img, p0 = (blk - b * per_img) * kCellsPerBlock;
    const int np = min(kCellsPerBlock, hw - p0);
    const float* in = s.logits + ((long)b * hw + p0) * s.ld;
    float* out = det + ((long)b * P.n_total + s.row_off) * P.attrs + (long)p0 * ch;
    const int grp = threadIdx.x >> 6, q = threadIdx.x & 63;
    for (int c0 = q * 4; c0 < ch; c0 += 256) {                     // one pass for up to 256 channels
        int a[4], attr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = (c0 + u) / P.attrs; attr[u] = (c0 + u) - a[u] * P.attrs; }
        for (int i0 = grp; i0 < np; i0 += 16) {                    // cells i0, i0+4, i0+8, i0+12 of this group
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 4 * u;
                t[u] = i < np ? __ldg(reinterpret_cast<const float4*>(in + (long)i * s.ld + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 4 * u;
                if (i >= np) break;
                const int p = p0 + i;
                const int y = p / s.w, x = p - y * s.w;
                float* o = out + (long)i * ch + c0;
                const float v[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c0 + e < ch) o[e] = decode_one(s, a[e], attr[e], x, y, v[e]);
            }
        }
    }
}

#ifdef YB_EXPERIMENTS
// Variant with staged output (YB_DECODE_STAGED=1, experiment builds only; same-box A/B in profiles/r02t_decode_staged_ab.txt:
// 0.181 ms against 0.167 ms for the direct kernel at 608x608 batch 32 -- the kernel is bound by its instructions per
// element, full-precision expf and a correctly rounded reciprocal, not by the 4-byte store pattern): phase 1
// parks the decoded values in shared memory in output order (thread q of a 64-thread group owns channels q, q+64, ...), phase
// 2 copies the block's contiguous output run with 16-byte stores, the shared-memory image shifted by the run's misalignment.
__global__ void __launch_bounds__(256) decode_nhwc_staged_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ det,
                                                                 int blocks_s0, int blocks_s1, int cpb) {
    extern __shared__ __align__(16) float out_s[];                 // [mis + cpb * ch]
    const int ch = 3 * P.attrs;
    int blk = blockIdx.x;
    const int si = blk >= blocks_s0 + blocks_s1 ? 2 : (blk >= blocks_s0 ? 1 : 0);
    blk -= si == 2 ? blocks_s0 + blocks_s1 : (si == 1 ? blocks_s0 : 0);
    const DecodeScale& s = P.sc[si];
    const int hw = s.h * s.w;
    const int per_img = (hw + cpb - 1) / cpb;
    const int b = blk / per_img, p0 = (blk - b * per_img) * cpb;
    const int np = min(cpb, hw - p0);
    const float* in = s.logits + ((long)b * hw + p0) * s.ld;
    float* out = det + ((long)b * P.n_total + s.row_off) * P.attrs + (long)p0 * ch;
    const int n = np * ch;
    const int mis = (int)((reinterpret_cast<uintptr_t>(out) >> 2) & 3);   // floats past a 16-byte boundary
    const int grp = threadIdx.x >> 6, q = threadIdx.x & 63;
    for (int cb = 0; cb < ch; cb += 256) {
        int a[4], attr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int c = cb + q + 64 * u; a[u] = c / P.attrs; attr[u] = c - a[u] * P.attrs; }
        int p = p0 + grp;
        int y = p / s.w, x = p - y * s.w;
        for (int i = grp; i < np; i += 4) {
            const float* ci = in + (long)i * s.ld + cb + q;
            float t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = cb + q + 64 * u < ch ? __ldg(ci + 64 * u) : 0.f;
            float* o = out_s + mis + i * ch + cb + q;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cb + q + 64 * u < ch) o[64 * u] = decode_one(s, a[u], attr[u], x, y, t[u]);
            x += 4;
            while (x >= s.w) { x -= s.w; ++y; }
        }
    }
    __syncthreads();
    const int head = min(n, (4 - mis) & 3);
    if ((int)threadIdx.x < head) out[threadIdx.x] = out_s[mis + threadIdx.x];
    const int nv = (n - head) >> 2;
    const float4* sv = reinterpret_cast<const float4*>(out_s + mis + head);     // mis + head is a multiple of 4
    float4* gv = reinterpret_cast<float4*>(out + head);
    for (int v = threadIdx.x; v < nv; v += 256) gv[v] = sv[v];
    for (int j = head + 4 * nv + (int)threadIdx.x; j < n; j += 256) out[j] = out_s[mis + j];
}
#endif

// NCHW (reference layout, API boundary) input: a block owns 32 consecutive cells of one image and
// scale; reads each channel row coalesced along the cells, transposes through shared memory, and
// writes the 32*3*attrs contiguous output floats coalesced.
__global__ void __launch_bounds__(256) decode_nchw_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ det,
                                                          int blocks_s0, int blocks_s1) {
    extern __shared__ float tile[];   // [ch][33]
    const int ch = 3 * P.attrs;
    int blk = blockIdx.x;
    const int si = blk >= blocks_s0 + blocks_s1 ? 2 : (blk >= blocks_s0 ? 1 : 0);
    blk -= si == 2 ? blocks_s0 + blocks_s1 : (si == 1 ? blocks_s0 : 0);
    const DecodeScale& s = P.sc[si];
    const int hw = s.h * s.w;
    const int per_img = (hw + 31) / 32;
    const int b = blk / per_img, p0 = (blk % per_img) * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int np = min(32, hw - p0);
    for (int c = warp; c < ch; c += 8) {
        float v = 0.f;
        if (lane < np) {
            const float t = s.logits[((long)b * ch + c) * hw + p0 + lane];
            const int p = p0 + lane;
            v = decode_one(s, c / P.attrs, c % P.attrs, p % s.w, p / s.w, t);
        }
        tile[c * 33 + lane] = v;
    }
    __syncthreads();
    float* o = det + ((long)b * P.n_total + s.row_off) * P.attrs + (long)p0 * ch;
    const int n = np * ch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = tile[(i % ch) * 33 + i / ch];
}

// ---- fused decode + score (the yb_detect path) ---------------------------------------------------------
// One warp per grid cell.  The cell's 3*(5+C) raw logits are one aligned, padded run of the NHWC head map:
// the warp loads it with 16-byte loads, parks it in shared memory, and then
//   * (kWriteDet) decodes every element and writes the cell's three rows of det_cat with coalesced stores
//     -- the standalone decode, without the strided scalar stores of decode_nhwc_kernel;
//   * (kScore) does what pp_score does on those rows without the [B,N,5+C] tensor ever being written or
//     re-read: score_c = cls_c * obj, row max with the lowest-index tie-break, strict '>' threshold, box
//     convert, -> rowcount / rowcand for pp_scan / pp_scatter.  Since cls_c <= 1, score_c <= obj in fp32 as
//     well (rounding is monotonic), so an anchor whose objectness does not pass cannot pass: its class
//     sigmoids are never evaluated (at conf 0.5 that is >99 % of the anchors; at conf 0.001 almost none).
// Arithmetic is decode_one / __fmul_rn exactly as in the two-kernel path, so the results are bit-identical.
constexpr int kFusedWarps = 8;

template <bool kWriteDet, bool kScore>
__global__ void __launch_bounds__(kFusedWarps * 32, 4)
decode_cells_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ det, float thr, int* __restrict__ rowcount,
                    float* __restrict__ rowcand, int ch_pad) {
    extern __shared__ __align__(16) float cell_smem[];      // [kFusedWarps][ch_pad]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = cell_smem + warp * ch_pad;
    const int attrs = P.attrs, ch = 3 * attrs;
    const long total = P.cells_before[3];
    const long wstride = (long)gridDim.x * kFusedWarps;
    int ca[8], cattr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ca[k] = (lane + 32 * k) / attrs; cattr[k] = (lane + 32 * k) - ca[k] * attrs; }
    for (long g = (long)blockIdx.x * kFusedWarps + warp; g < total; g += wstride) {
        const int si = g >= P.cells_before[2] ? 2 : (g >= P.cells_before[1] ? 1 : 0);
        const DecodeScale& s = P.sc[si];
        const int hw = s.h * s.w;
        const long gl = g - P.cells_before[si];
        const int b = (int)(gl / hw), p = (int)(gl - (long)b * hw);
        const int y = p / s.w, x = p - y * s.w;
        const float* in = s.logits + gl * s.ld;              // (b*hw + p) * ld
        __syncwarp();                                        // previous cell's readers are done with sm
        for (int c0 = lane * 4; c0 < ch_pad; c0 += 128)
            *reinterpret_cast<float4*>(sm + c0) = __ldg(reinterpret_cast<const float4*>(in + c0));
        __syncwarp();
        const long row0 = (long)b * P.n_total + s.row_off + (long)p * 3;
        bool need[3] = {true, true, true};
        float obj[3];
        if (kScore && !kWriteDet) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                obj[a] = decode_one(s, a, 4, x, y, sm[a * attrs + 4]);
                need[a] = obj[a] > thr;
                if (!need[a] && lane == 0) rowcount[row0 + a] = 0;
            }
            if (!(need[0] || need[1] || need[2])) continue;
        }
        __syncwarp();                                        // everyone has read the raw objectness logits
        // decode in place: lane takes channels lane, lane+32, ... (anchor / attribute of the first eight are
        // per-thread constants, computed once outside the cell loop)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + 32 * k;
            if (c < ch) {
                const float v = decode_one(s, ca[k], cattr[k], x, y, sm[c]);
                sm[c] = v;
                if (kWriteDet) det[row0 * attrs + c] = v;
            }
        }
        for (int c = lane + 256; c < ch; c += 32) {
            const int a = c / attrs, attr = c - a * attrs;
            const float v = decode_one(s, a, attr, x, y, sm[c]);
            sm[c] = v;
            if (kWriteDet) det[row0 * attrs + c] = v;
        }
        if (!kScore) continue;
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!need[a]) continue;
            const float* r = sm + a * attrs;
            const float o = r[4];
            float best = -INFINITY;
            int bidx = 0x7fffffff;
            for (int e = 5 + lane; e < attrs; e += 32) {
                const float sc = __fmul_rn(r[e], o);
                if (sc > best) { best = sc; bidx = e - 5; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, d);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, d);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            const bool pass = best > thr;
            const long row = row0 + a;
            if (lane == 0) rowcount[row] = pass ? 1 : 0;
            if (pass && lane < 8) {                          // same 8 floats pp_score emits (postprocess.cu: pp_emit_row)
                const float cx = r[0], cy = r[1], hw2 = __fdiv_rn(r[2], 2.f), hh2 = __fdiv_rn(r[3], 2.f);
                float v = 0.f;
                switch (lane) {
                    case 0: v = __fsub_rn(cx, hw2); break;
                    case 1: v = __fsub_rn(cy, hh2); break;
                    case 2: v = __fadd_rn(cx, hw2); break;
                    case 3: v = __fadd_rn(cy, hh2); break;
                    case 4: v = o; break;
                    case 5: v = best; break;
                    case 6: v = (float)bidx; break;
                    default: v = __int_as_float((int)(row - (long)b * P.n_total)); break;
                }
                rowcand[row * 8 + lane] = v;
            }
        }
    }
}

// ---- objectness-first scoring (what yb_detect launches) ---------------------------------------------------
// Phase 1: lane <-> cell.  Each lane reads only the three objectness logits of its cell (3 x 4 bytes out of the
// 1 KB the cell occupies), and rows whose objectness fails the threshold are settled right there
// (rowcount = 0).  Phase 2: the warp walks the cells that still have a live anchor and treats each one as
// decode_cells_kernel does (16-byte loads -> shared memory -> decode -> class arg-max -> candidate row).
// At conf 0.5 almost no cell reaches phase 2 and the kernel reads ~10 % of the head maps; at conf 0.001 nearly
// every cell does and the cost is that of decode_cells_kernel plus the three probes.  Results are bit-identical
// to decode + pp_score in both regimes (same decode_one / __fmul_rn arithmetic, same tie-break).
// Decode + score one cell with the whole warp (phase 2 of the scoring kernels): nd = bit mask of the anchors
// whose objectness passed.
__device__ __forceinline__ void score_cell(const DecodeParams& P, float* sm, int lane, long gc, int nd, float thr,
                                           int* __restrict__ rowcount, float* __restrict__ rowcand, int ch_pad,
                                           const int (&ca)[8], const int (&cattr)[8]) {
    const int attrs = P.attrs, ch = 3 * attrs;
    const int si = gc >= P.cells_before[2] ? 2 : (gc >= P.cells_before[1] ? 1 : 0);
    const DecodeScale& s = P.sc[si];
    const long gl = gc - P.cells_before[si];
    const int hw = s.h * s.w;
    const int b = (int)(gl / hw), p = (int)(gl - (long)b * hw);
    const int y = p / s.w, x = p - y * s.w;
    const float* in = s.logits + gl * s.ld;
    const long row0 = (long)b * P.n_total + s.row_off + (long)p * 3;
    __syncwarp();                                        // previous cell's readers are done with sm
    for (int c0 = lane * 4; c0 < ch_pad; c0 += 128)
        *reinterpret_cast<float4*>(sm + c0) = __ldg(reinterpret_cast<const float4*>(in + c0));
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = lane + 32 * k;
        if (c < ch) sm[c] = decode_one(s, ca[k], cattr[k], x, y, sm[c]);
    }
    for (int c = lane + 256; c < ch; c += 32) {
        const int a = c / attrs;
        sm[c] = decode_one(s, a, c - a * attrs, x, y, sm[c]);
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!((nd >> a) & 1)) continue;
        const float* r = sm + a * attrs;
        const float o = r[4];
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        for (int e = 5 + lane; e < attrs; e += 32) {
            const float sc = __fmul_rn(r[e], o);
            if (sc > best) { best = sc; bidx = e - 5; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, d);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        const bool pass = best > thr;
        const long row = row0 + a;
        if (lane == 0) rowcount[row] = pass ? 1 : 0;
        if (pass && lane < 8) {                          // same 8 floats pp_score emits (postprocess.cu: pp_emit_row)
            const float cx = r[0], cy = r[1], hw2 = __fdiv_rn(r[2], 2.f), hh2 = __fdiv_rn(r[3], 2.f);
            float v = 0.f;
            switch (lane) {
                case 0: v = __fsub_rn(cx, hw2); break;
                case 1: v = __fsub_rn(cy, hh2); break;
                case 2: v = __fadd_rn(cx, hw2); break;
                case 3: v = __fadd_rn(cy, hh2); break;
                case 4: v = o; break;
                case 5: v = best; break;
                case 6: v = (float)bidx; break;
                default: v = __int_as_float((int)(row - (long)b * P.n_total)); break;
            }
            rowcand[row * 8 + lane] = v;
        }
    }
}

// Phase 1 for one cell (one lane): the three objectness probes.  Returns the mask of live anchors and settles
// the dead ones (rowcount = 0).
__device__ __forceinline__ int probe_cell(const DecodeParams& P, long g, float thr, int* __restrict__ rowcount) {
    const int si = g >= P.cells_before[2] ? 2 : (g >= P.cells_before[1] ? 1 : 0);
    const long gl = g - P.cells_before[si];
    const int hw = P.sc[si].h * P.sc[si].w;
    const int b = (int)(gl / hw), p = (int)(gl - (long)b * hw);
    const float* in = P.sc[si].logits + gl * P.sc[si].ld;
    const long row0 = (long)b * P.n_total + P.sc[si].row_off + (long)p * 3;
    float t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) t[a] = __ldg(in + a * P.attrs + 4);
    int need = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // decode_one(attr = 4) is (sigmoid + 0) * 1 == sigmoid exactly
        if (sigmoidf_rn(t[a]) > thr) need |= 1 << a;
        else rowcount[row0 + a] = 0;
    }
    return need;
}

// Single-kernel form: a warp probes 32 cells, then walks its own live cells.
__global__ void __launch_bounds__(kFusedWarps * 32, 4)
score_cells_kernel(const __grid_constant__ DecodeParams P, float thr, int* __restrict__ rowcount, float* __restrict__ rowcand,
                   int ch_pad) {
    extern __shared__ __align__(16) float cell_smem[];      // [kFusedWarps][ch_pad]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = cell_smem + warp * ch_pad;
    const int attrs = P.attrs;
    const long total = P.cells_before[3];
    const long wstride = (long)gridDim.x * kFusedWarps * 32;
    int ca[8], cattr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ca[k] = (lane + 32 * k) / attrs; cattr[k] = (lane + 32 * k) - ca[k] * attrs; }
    for (long g0 = ((long)blockIdx.x * kFusedWarps + warp) * 32; g0 < total; g0 += wstride) {
        const long g = g0 + lane;
        const int need = g < total ? probe_cell(P, g, thr, rowcount) : 0;
        unsigned live = __ballot_sync(0xffffffffu, need != 0);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            score_cell(P, sm, lane, g0 + src, __shfl_sync(0xffffffffu, need, src), thr, rowcount, rowcand, ch_pad, ca, cattr);
        }
    }
}

// Two-kernel form (default): the probe appends live cells to a list, and the scoring kernel spreads that list over
// every warp of the GPU -- in the single-kernel form a warp walks its own live cells one after another, and the
// slowest warp sets the time.  list[0] = count (zeroed by the host), entries (cell << 3 | anchor mask) follow.
__global__ void __launch_bounds__(256) probe_cells_kernel(const __grid_constant__ DecodeParams P, float thr,
                                                          int* __restrict__ rowcount, int* __restrict__ list) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int need = g < P.cells_before[3] ? probe_cell(P, g, thr, rowcount) : 0;
    const unsigned live = __ballot_sync(0xffffffffu, need != 0);
    if (!live) return;
    int base = 0;
    if (lane == __ffs(live) - 1) base = atomicAdd(list, __popc(live));
    base = __shfl_sync(0xffffffffu, base, __ffs(live) - 1);
    if (need) list[1 + base + __popc(live & ((1u << lane) - 1))] = (int)(g << 3) | need;
}

__global__ void __launch_bounds__(kFusedWarps * 32, 4)
score_list_kernel(const __grid_constant__ DecodeParams P, float thr, int* __restrict__ rowcount, float* __restrict__ rowcand,
                  const int* __restrict__ list, int ch_pad) {
    extern __shared__ __align__(16) float cell_smem[];      // [kFusedWarps][ch_pad]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = cell_smem + warp * ch_pad;
    const int attrs = P.attrs;
    const int count = list[0];
    int ca[8], cattr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ca[k] = (lane + 32 * k) / attrs; cattr[k] = (lane + 32 * k) - ca[k] * attrs; }
    for (int i = blockIdx.x * kFusedWarps + warp; i < count; i += gridDim.x * kFusedWarps) {
        const int e = list[1 + i];
        score_cell(P, sm, lane, (long)(e >> 3), e & 7, thr, rowcount, rowcand, ch_pad, ca, cattr);
    }
}

}  // namespace

// mode bit 0: write det_cat, bit 1: score into rowcount / rowcand (non-eval post-process front end).
// Requires the NHWC head maps with pixel pitch ld == ch_pad (a multiple of 4 floats, 16-byte aligned cells).
cudaError_t launch_decode_cells(const DecodeScale sc[3], int B, int attrs, int n_total, int mode, float* det, float thr,
                                int* rowcount, float* rowcand, int* list, int* extra_launches, int num_sms, cudaStream_t s) {
    if (extra_launches) *extra_launches = 0;
    DecodeParams P;
    P.B = B; P.attrs = attrs; P.n_total = n_total;
    P.cells_before[0] = 0;
    for (int i = 0; i < 3; ++i) {
        P.sc[i] = sc[i];
        P.cells_before[i + 1] = P.cells_before[i] + (long)B * sc[i].h * sc[i].w;
    }
    const int ch_pad = (int)sc[0].ld;
    const size_t smem = (size_t)kFusedWarps * ch_pad * sizeof(float);
    const long blocks_needed = (P.cells_before[3] + kFusedWarps - 1) / kFusedWarps;
    const unsigned grid = (unsigned)std::min<long>(blocks_needed, (long)num_sms * 4);
    if (smem > 48 * 1024) {        // more than ~500 classes: the cells of a block no longer fit the default dynamic smem limit
        if (smem > 128 * 1024) return cudaErrorInvalidValue;
        static PerDeviceOnce attr_once;
        cudaError_t e = attr_once.run([] {
            cudaError_t r = cudaFuncSetAttribute(decode_cells_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(decode_cells_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(decode_cells_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(score_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(score_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            return r;
        });
        if (e != cudaSuccess) return e;
    }
    if (mode == 1) decode_cells_kernel<true, false><<<grid, kFusedWarps * 32, smem, s>>>(P, det, thr, rowcount, rowcand, ch_pad);
    else if (mode == 2) {
        // YB_SCORE_MODE: 0 one cell per warp at a time, everything decoded; 1 objectness-first in one kernel;
        // 2 (default) objectness probe -> live-cell list -> scoring kernel over the list
        static const int score_mode = tune_env("YB_SCORE_MODE") ? atoi(tune_env("YB_SCORE_MODE")) : 2;
        const long cells = P.cells_before[3];
        const long groups = (cells + 31) / 32;
        const unsigned grid2 = (unsigned)std::min<long>((groups + kFusedWarps - 1) / kFusedWarps, (long)num_sms * 4);
        if (score_mode == 2 && list && cells < (1L << 28)) {
            cudaMemsetAsync(list, 0, sizeof(int), s);
            probe_cells_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(P, thr, rowcount, list);
            score_list_kernel<<<num_sms * 4, kFusedWarps * 32, smem, s>>>(P, thr, rowcount, rowcand, list, ch_pad);
            if (extra_launches) *extra_launches = 1;
        } else if (score_mode >= 1) {
            score_cells_kernel<<<grid2, kFusedWarps * 32, smem, s>>>(P, thr, rowcount, rowcand, ch_pad);
        } else {
            decode_cells_kernel<false, true><<<grid, kFusedWarps * 32, smem, s>>>(P, det, thr, rowcount, rowcand, ch_pad);
        }
    }
    else decode_cells_kernel<true, true><<<grid, kFusedWarps * 32, smem, s>>>(P, det, thr, rowcount, rowcand, ch_pad);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeScale sc[3], int nchw, int B, int attrs, int n_total, float* det, cudaStream_t s) {
    DecodeParams P;
    P.B = B; P.attrs = attrs; P.n_total = n_total;
    P.cells_before[0] = 0;
    for (int i = 0; i < 3; ++i) {
        P.sc[i] = sc[i];
        P.cells_before[i + 1] = P.cells_before[i] + (long)B * sc[i].h * sc[i].w;
    }
    if (!nchw) {
        int nb[3];
#ifdef YB_EXPERIMENTS
        static const bool staged = tune_env("YB_DECODE_STAGED") && atoi(tune_env("YB_DECODE_STAGED")) != 0;
        const int ch = 3 * attrs;
        if (staged && ch * 4 <= 40 * 1024) {
            const int cpb = std::max(1, std::min(kCellsPerBlock, (40 * 1024) / (ch * 4)));
            for (int i = 0; i < 3; ++i) nb[i] = B * ((sc[i].h * sc[i].w + cpb - 1) / cpb);
            decode_nhwc_staged_kernel<<<nb[0] + nb[1] + nb[2], 256, ((size_t)cpb * ch + 4) * sizeof(float), s>>>(P, det, nb[0], nb[1], cpb);
            return cudaGetLastError();
        }
#endif
        for (int i = 0; i < 3; ++i) nb[i] = B * ((sc[i].h * sc[i].w + kCellsPerBlock - 1) / kCellsPerBlock);
        decode_nhwc_kernel<<<nb[0] + nb[1] + nb[2], 256, 0, s>>>(P, det, nb[0], nb[1]);
    } else {
        int nb[3];
        for (int i = 0; i < 3; ++i) nb[i] = B * ((sc[i].h * sc[i].w + 31) / 32);
        const size_t smem = (size_t)3 * attrs * 33 * sizeof(float);
        static PerDeviceOnce attr_once;
        if (smem > 48 * 1024) {
            cudaError_t e = attr_once.run([] { return cudaFuncSetAttribute(decode_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
            if (e != cudaSuccess) return e;
        }
        decode_nchw_kernel<<<nb[0] + nb[1] + nb[2], 256, smem, s>>>(P, det, nb[0], nb[1]);
    }
    return cudaGetLastError();
}

}  // namespace yb
