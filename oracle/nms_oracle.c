/* CPU oracle, C leg  --  TEST INFRASTRUCTURE ONLY (see oracle/yolo_oracle.py header).
 *
 * Plain-C restatement of the reference post-process so that the large parity cases
 * (10k-20k candidates per image) finish in seconds instead of minutes:
 *
 *   boundingbox.py:25-29   bbox_cxcywh_to_x1y1x2y2
 *   utils.py:226-258       postprocessing   (score = cls*obj, max / threshold, "[]" convention)
 *   utils.py:148-202       get_nms_detections (class-ascending, score-descending, greedy rows)
 *   utils.py:98-119        iou_vectorized   (every fp32 op rounded separately; 0/0 = NaN rule)
 *   utils.py:204-224       get_raw_detections
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (oracle/Makefile).
 * -ffp-contract=off matters: the reference rounds the product and the sums separately.
 * Tie-break: score descending, then candidate order ascending (stable sort), as in the
 * Python leg.  Pinned against tests/golden/postprocess_*.npz.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { float score; long order; } skey_t;

static int cmp_desc(const void* a, const void* b) {
    const skey_t* x = (const skey_t*)a; const skey_t* y = (const skey_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->order > y->order) - (x->order < y->order);
}

static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

/* utils.py:109-118, element (i,j) */
static inline float iou_pair(const float* a, const float* b) {
    float ltx = fmaxf_(a[0], b[0]), lty = fmaxf_(a[1], b[1]);
    float rbx = fminf_(a[2], b[2]), rby = fminf_(a[3], b[3]);
    float iw = rbx - ltx; if (iw < 0.f) iw = 0.f;
    float ih = rby - lty; if (ih < 0.f) ih = 0.f;
    float inter = iw * ih;
    float area_a = (a[2] - a[0]) * (a[3] - a[1]);
    float area_b = (b[2] - b[0]) * (b[3] - b[1]);
    float uni = (area_b + area_a) - inter;
    return inter / uni;
}

/* det: [B,N,5+C] fp32 (cx,cy,w,h,obj,cls...).  rows: [B,cap,7], src: [B,cap], counts: [B].
 * Returns total rows written, or -1 when no (box[,class]) in the whole batch passes the
 * threshold (the reference then returns [] -- utils.py:247-251), or -2 on cap overflow. */
long oracle_postprocess(const float* det, long B, long N, long C, float conf_thr, float nms_thr,
                        int is_eval, int use_nms, float* rows, long* src, long* counts, long cap) {
    const long A = 5 + C;
    long maxcand = N * (is_eval ? C : 1);
    float* cbox = (float*)malloc(sizeof(float) * 6 * (size_t)maxcand);   /* x1,y1,x2,y2,obj,score */
    long* cidx = (long*)malloc(sizeof(long) * (size_t)maxcand);
    int* ccls = (int*)malloc(sizeof(int) * (size_t)maxcand);
    skey_t* keys = (skey_t*)malloc(sizeof(skey_t) * (size_t)maxcand);
    float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)maxcand);
    char* alive = (char*)malloc((size_t)maxcand);
    long* members = (long*)malloc(sizeof(long) * (size_t)maxcand);
    long total = 0, any = 0;
    long ret = 0;

    for (long b = 0; b < B; ++b) {
        long nc = 0;
        for (long n = 0; n < N; ++n) {
            const float* r = det + (b * N + n) * A;
            float hw = r[2] / 2, hh = r[3] / 2;
            float x1 = r[0] - hw, x2 = r[0] + hw, y1 = r[1] - hh, y2 = r[1] + hh;
            float obj = r[4];
            if (is_eval) {
                for (long c = 0; c < C; ++c) {
                    float s = r[5 + c] * obj;
                    if (s > conf_thr) {
                        float* o = cbox + 6 * nc;
                        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = obj; o[5] = s;
                        cidx[nc] = n; ccls[nc] = (int)c; ++nc;
                    }
                }
            } else {
                float best = r[5] * obj; long bc = 0;
                for (long c = 1; c < C; ++c) {
                    float s = r[5 + c] * obj;
                    if (s > best) { best = s; bc = c; }        /* first occurrence wins ties */
                }
                if (best > conf_thr) {
                    float* o = cbox + 6 * nc;
                    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = obj; o[5] = best;
                    cidx[nc] = n; ccls[nc] = (int)bc; ++nc;
                }
            }
        }
        any += nc;
        long k = 0;
        float* out = rows + b * cap * 7;
        long* osrc = src + b * cap;
        if (!use_nms) {
            if (nc > cap) { ret = -2; goto done; }
            for (long i = 0; i < nc; ++i) {
                memcpy(out + 7 * k, cbox + 6 * i, 6 * sizeof(float));
                out[7 * k + 6] = (float)ccls[i]; osrc[k] = cidx[i]; ++k;
            }
        } else {
            for (long c = 0; c < C; ++c) {
                long m = 0;
                for (long i = 0; i < nc; ++i) if (ccls[i] == c) members[m++] = i;
                if (!m) continue;
                for (long i = 0; i < m; ++i) { keys[i].score = cbox[6 * members[i] + 5]; keys[i].order = i; }
                qsort(keys, (size_t)m, sizeof(skey_t), cmp_desc);
                for (long i = 0; i < m; ++i) {
                    memcpy(sb + 4 * i, cbox + 6 * members[keys[i].order], 4 * sizeof(float));
                    alive[i] = iou_pair(sb + 4 * i, sb + 4 * i) > nms_thr;     /* diagonal, utils.py:182 */
                }
                for (long i = 0; i < m; ++i) {
                    if (!alive[i]) continue;
                    for (long j = i + 1; j < m; ++j)
                        if (alive[j] && iou_pair(sb + 4 * i, sb + 4 * j) > nms_thr) alive[j] = 0;
                }
                for (long i = 0; i < m; ++i) {
                    if (!alive[i]) continue;
                    if (k >= cap) { ret = -2; goto done; }
                    long ci = members[keys[i].order];
                    memcpy(out + 7 * k, cbox + 6 * ci, 6 * sizeof(float));
                    out[7 * k + 6] = (float)c; osrc[k] = cidx[ci]; ++k;
                }
            }
        }
        counts[b] = k;
        total += k;
    }
    ret = any ? total : -1;
done:
    free(cbox); free(cidx); free(ccls); free(keys); free(sb); free(alive); free(members);
    return ret;
}
