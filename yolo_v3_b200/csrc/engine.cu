// Engine + C ABI of libyolo_b200.so: weight store, per-(B,H,W) execution plan, and the entry points
// declared in include/yolo_b200.h.  The network wiring restates YoloNet.__init__/forward
// (reference darknet.py:167-231): Darknet-53 (blocks [1,2,8,8,4], routes after mlist[14] and
// mlist[23]), three PreDetectionConvGroups, two UpsampleGroups (upsampled half first in the concat,
// darknet.py:162) -- as a static schedule of 75 fused convolution launches over NHWC buffers:
//   * BN(eval)+LeakyReLU+residual are the conv epilogue;
//   * route tails are written by their producer directly into the channel slice of the concat
//     buffer, and the 1x1 "up" convs write each pixel to its 2x2 block of the other slice, so
//     interpolate() and cat() (darknet.py:161-162) cost no pass of their own;
//   * the three head maps are decoded by one kernel straight into the concatenated [B,N,5+C] tensor.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

#include <cstdlib>

#include "yb_internal.h"

namespace yb {

static thread_local std::string g_last_error;

struct Op {
    int layer = 0;
    bool stem = false;
    ConvArgs a{};
    bool use_tc = false, use_halo = false;
    TcPlan tc;
    HaloPlan halo;
    StemHaloPlan stem_halo;
    bool stem_split_tc = false;       // YB_MODE_FP32_TC: the split halo stem (stem_halo_split.cu) instead of the CUDA-core kernel
    bool fused01 = false;             // the stem op also runs layer 1 (stem_block.cu) and writes layer 1's output
    StemBlockPlan stem_block;
    bool upcopy = false;              // YB_MODE_FP32_TC: nearest x2 copy of up_src into the concat slice up_dst
    TView up_src, up_dst;
};

struct Plan {
    int B = 0, H = 0, W = 0, mode = -1;
    std::vector<void*> allocs;
    std::vector<Op> ops;
    float* logits[3] = {nullptr, nullptr, nullptr};
    int gh[3], gw[3];
    TView backbone_out;
    int n_backbone_ops = 0;
    size_t bytes = 0;
    // CUDA graphs of ops [1, n_ops) -- everything after the stem works on plan-owned buffers only, so the launches can be
    // replayed as they are.  One executable per n_ops value in use (whole network, backbone only); captured on the
    // second call with a shape (the first call runs eagerly and sets the kernels' function attributes).
    struct GraphSlot { int n_ops = 0; cudaGraphExec_t exec = nullptr; int calls = 0; bool failed = false; };
    GraphSlot graphs[2];
};

}  // namespace yb

using namespace yb;

struct yb_ctx {
    int device = 0;
    int num_classes = 80;
    int attrs = 85;
    float anchors[18];
    int num_sms = 148;
    int cc_major = 0;
    std::vector<Layer> layers;
    struct Slot { int layer; int field; size_t numel; };   // field: 0 w,1 bn.w,2 bn.b,3 mean,4 var,5 nbt,6 bias
    std::vector<std::pair<std::string, Slot>> keys;         // registration order
    std::map<std::string, int> key_index;
    int mode = -1;
    int input_f16 = 0;                // element type of the images handed to yb_forward / yb_detect / ... (yb_set_input_dtype)
    int graph_mode = 2;               // yb_set_graph_mode: 0 never, 1 always, 2 auto (launch-bound shapes only)
    long long graph_replays = 0;      // calls whose convolution launches were replayed from a captured graph
    cudaStream_t cap_stream = nullptr; // capture happens on a private stream: the caller's may be the legacy default stream,
                                      // which cannot be captured (PyTorch's current stream usually is)
    bool finalized = false;
    unsigned char* d_blob = nullptr;
    size_t blob_bytes = 0;
    float stem_sb[64] = {};           // host mirror of the stem's scale[32] | bias[32] (kernel parameters of the halo stem):
                                      // re-read from the blob by yb_finalize and after yb_bcast_weights
    float stem_sc_eff[32] = {};       // fp32-grade mode: scale * 2^-(shift + 8) and the per-channel weight shift of the split
    int stem_shift[32] = {};          // halo stem (stem_halo_split.cu), derived from the blob's fp32 stem weights
    std::vector<std::unique_ptr<Plan>> plans;
    PostBuffers post;
    float* box_params = nullptr;      // [B][6] per-image parameters of yb_correct_boxes
    int box_params_cap = 0;
    LbImage* lb_params = nullptr;     // [B] per-image parameters of yb_letterbox
    int lb_params_cap = 0;
    float* det_scratch = nullptr;     // for yb_detect
    size_t det_scratch_bytes = 0;
    int* dbg = nullptr;               // device alias of dbg_host (mapped pinned memory): watchdog words of the
    int* dbg_host = nullptr;          // tensor-core kernel, readable by the host even after a device trap
    mutable std::string err;
    long long launches = 0;
    bool profiling = false;
    bool profile_layers = false;      // profiling level 2: an event after every convolution (perturbs PDL overlap)
    // profiling level 3 (yb_detect only): section events go to a rotating bank of kProfBank sets and nothing is
    // synchronised per call; yb_get_section_ms averages the sets recorded since the last query
    bool profile_deferred = false;
    std::vector<cudaEvent_t> bank;    // [kProfBank][4]: start, after conv stack, after decode/score, after post-process
    int bank_next = 0, bank_count = 0;
    std::vector<cudaEvent_t> ev;
    float sec_ms[3] = {0, 0, 0};
    std::vector<float> layer_ms;
    // NCCL (comm.cu)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
};

namespace yb {
int comm_unique_id(uint8_t* id, std::string& err);
int comm_init(void** comm, const uint8_t* id, int rank, int world, std::string& err);
int comm_bcast(void* comm, void* buf, size_t bytes, int root, cudaStream_t s, std::string& err);
int comm_allgather(void* comm, const void* send, void* recv, size_t bytes, cudaStream_t s, std::string& err);
void comm_destroy(void* comm);
}  // namespace yb

namespace {

int fail(const yb_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    g_last_error = msg;
    return code;
}
#define YB_CUDA(c, expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(c, YB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));            \
    } while (0)

// Per-call parameter blocks (per-image letterbox / box-correction parameters, a few KB) go to the device with a plain
// cudaMemcpyAsync on the caller's stream.  (A ring of pinned staging slots was measured in round 2 and lost:
// profiles/r02a_pinned_bench.txt.)
int stage_params(yb_ctx* c, const void* src, size_t bytes, void* dst, cudaStream_t s) {
    YB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return YB_OK;
}

const float kDefaultAnchors[18] = {10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326};
const int kBlocks[5] = {1, 2, 8, 8, 4};
const int kMasks[3][3] = {{6, 7, 8}, {3, 4, 5}, {0, 1, 2}};   // yolo1, yolo2, yolo3 (darknet.py:184,189,194)

void build_layers(yb_ctx* c) {
    auto add = [&](const std::string& key, int cin, int cout, int ks, int stride, bool bn) {
        Layer L;
        L.key = key; L.cin = cin; L.cout = cout; L.ks = ks; L.stride = stride; L.bn = bn;
        L.cout_pad = (cout + 15) / 16 * 16;
        L.w.assign((size_t)cout * cin * ks * ks, 0.f);
        if (bn) {
            L.g.assign(cout, 1.f); L.b.assign(cout, 0.f); L.mean.assign(cout, 0.f); L.var.assign(cout, 1.f);
        } else {
            L.bias.assign(cout, 0.f);
        }
        c->layers.push_back(std::move(L));
    };
    add("feature.mlist.0", 3, 32, 3, 1, true);
    int idx = 1, ch = 32;
    for (int s = 0; s < 5; ++s) {
        add("feature.mlist." + std::to_string(idx), ch, ch * 2, 3, 2, true);
        ++idx; ch *= 2;
        for (int j = 0; j < kBlocks[s]; ++j) {
            add("feature.mlist." + std::to_string(idx) + ".conv1", ch, ch / 2, 1, 1, true);
            add("feature.mlist." + std::to_string(idx) + ".conv2", ch / 2, ch, 3, 1, true);
            ++idx;
        }
    }
    auto predet = [&](const std::string& name, int nin, int nout) {
        for (int i = 0; i < 3; ++i) {
            add(name + ".mlist." + std::to_string(2 * i), nin, nout, 1, 1, true);
            add(name + ".mlist." + std::to_string(2 * i + 1), nout, nout * 2, 3, 1, true);
            nin = nout * 2;
        }
        add(name + ".mlist.6", nin, 3 * c->attrs, 1, 1, false);
    };
    predet("pre_det1", 1024, 512);
    add("up1.conv", 512, 256, 1, 1, true);
    predet("pre_det2", 768, 256);
    add("up2.conv", 256, 128, 1, 1, true);
    predet("pre_det3", 384, 128);

    for (int i = 0; i < (int)c->layers.size(); ++i) {
        const Layer& L = c->layers[i];
        auto reg = [&](const std::string& k, int field, size_t n) {
            c->key_index[k] = (int)c->keys.size();
            c->keys.push_back({k, {i, field, n}});
        };
        if (L.bn) {
            reg(L.key + ".conv.weight", 0, L.w.size());
            reg(L.key + ".bn.weight", 1, L.cout);
            reg(L.key + ".bn.bias", 2, L.cout);
            reg(L.key + ".bn.running_mean", 3, L.cout);
            reg(L.key + ".bn.running_var", 4, L.cout);
            reg(L.key + ".bn.num_batches_tracked", 5, 1);
        } else {
            reg(L.key + ".weight", 0, L.w.size());
            reg(L.key + ".bias", 6, L.cout);
        }
    }
}

std::vector<float>* field_vec(Layer& L, int field) {
    switch (field) {
        case 0: return &L.w;
        case 1: return &L.g;
        case 2: return &L.b;
        case 3: return &L.mean;
        case 4: return &L.var;
        case 6: return &L.bias;
        default: return nullptr;
    }
}

void free_plans(yb_ctx* c) {
    for (auto& p : c->plans) {
        for (auto& g : p->graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        for (void* q : p->allocs) cudaFree(q);
    }
    c->plans.clear();
}

void free_post(yb_ctx* c) {
    PostBuffers& b = c->post;
    cudaFree(b.rowcount); cudaFree(b.rowoff); cudaFree(b.cand_total); cudaFree(b.rowcand); cudaFree(b.cand);
    cudaFree(b.keys); cudaFree(b.sbox); cudaFree(b.keep); cudaFree(b.seg);
    b = PostBuffers();
}

int ensure_post(yb_ctx* c, int B, int N, int is_eval) {
    PostBuffers& b = c->post;
    const long want = (long)N * (is_eval ? c->num_classes : 1);
    if (want >= (1L << 22)) return fail(c, YB_E_ARG, "post-process: more than 4M candidates per image");
    if (b.B >= B && b.N == N && b.cand_cap >= want && b.C == c->num_classes) return YB_OK;
    const long old_cap = b.N == N ? b.cand_cap : 0;      // (free_post resets the struct: read the old capacity first)
    free_post(c);
    const int cand_cap = (int)std::max<long>(want, old_cap);
    int sort_cap = 2;
    while (sort_cap < cand_cap) sort_cap <<= 1;
    const int C = c->num_classes;
    YB_CUDA(c, cudaMalloc(&b.rowcount, sizeof(int) * (size_t)B * N));
    YB_CUDA(c, cudaMalloc(&b.rowoff, sizeof(int) * (size_t)B * N));
    YB_CUDA(c, cudaMalloc(&b.cand_total, sizeof(int) * (size_t)B));
    YB_CUDA(c, cudaMalloc(&b.rowcand, sizeof(float) * 8 * (size_t)B * N));
    YB_CUDA(c, cudaMalloc(&b.cand, sizeof(float) * 8 * (size_t)B * cand_cap));
    YB_CUDA(c, cudaMalloc(&b.keys, sizeof(unsigned long long) * (size_t)B * sort_cap));
    YB_CUDA(c, cudaMalloc(&b.sbox, sizeof(float4) * (size_t)B * cand_cap));
    YB_CUDA(c, cudaMalloc(&b.keep, (size_t)B * cand_cap));
    YB_CUDA(c, cudaMalloc(&b.seg, sizeof(int) * 2 * (size_t)B * C));
    b.B = B; b.N = N; b.cand_cap = cand_cap; b.sort_cap = sort_cap; b.C = C;
    return YB_OK;
}

// ---- plan ---------------------------------------------------------------------------------------

struct PlanBuilder {
    yb_ctx* c;
    Plan* p;
    size_t elem;
    std::string err;
    bool split = false;   // YB_MODE_FP32_TC

    void* alloc(size_t bytes) {
        void* q = nullptr;
        bytes = (bytes + 255) / 256 * 256;
        if (cudaMalloc(&q, bytes) != cudaSuccess) { err = "cudaMalloc of " + std::to_string(bytes) + " bytes failed"; return nullptr; }
        p->allocs.push_back(q);
        p->bytes += bytes;
        return q;
    }
    // `ld` = channels of the whole buffer (C for a plain tensor, 768 / 384 for the concat buffers).  Split mode stores
    // hi[ld] | lo[ld] per pixel: the pitch doubles and the lo half of any channel slice sits `ld` elements after its hi half.
    TView view(void* base, int B, int H, int W, int C, long ld, int ch_off = 0) const {
        TView v;
        v.p = static_cast<unsigned char*>(base) + (size_t)ch_off * elem;
        v.B = B; v.H = H; v.W = W; v.C = C;
        v.ld = split ? 2 * ld : ld;
        v.lo = split ? ld : 0;
        return v;
    }
    // conv `li`: in -> out (+res), returns false on error
    bool conv(int li, const TView& in, const TView& out, const TView* res, bool upsample, bool head) {
        const Layer& L = c->layers[li];
        Op op;
        op.layer = li;
        ConvArgs& a = op.a;
        a.in = in.p; a.in_ld = in.ld;
        a.out = out.p; a.out_ld = out.ld;
        a.res = res ? res->p : nullptr; a.res_ld = res ? res->ld : 0;
        a.scale = L.d_scale; a.bias = L.d_bias;
        a.B = in.B; a.H = in.H; a.W = in.W; a.Cin = L.cin;
        a.Ho = in.H / L.stride; a.Wo = in.W / L.stride;
        a.Cout = head ? L.cout_pad : L.cout;
        a.ks = L.ks; a.stride = L.stride; a.pad = (L.ks - 1) / 2;
        a.leaky = L.bn ? 1 : 0;
        a.upsample = upsample ? 1 : 0;
        a.out_f32 = head ? 1 : 0;
        a.split = split ? 1 : 0;
        a.in_lo = in.lo; a.out_lo = head ? 0 : out.lo; a.res_lo = res ? res->lo : 0;
        if (in.C != L.cin) { err = "plan: channel mismatch at layer " + L.key; return false; }
        if (p->mode == YB_MODE_FP32_TC) {
            if (!tc_supported(a)) { err = "plan: layer " + L.key + " not supported by the tensor-core kernel"; return false; }
            op.use_tc = true;
            std::string e = tc_make_plan(op.tc, a, L.d_w16, L.cout_pad, 2 * L.ks * L.ks * L.cin, c->num_sms);
            if (!e.empty()) { err = "plan: layer " + L.key + ": " + e; return false; }
        } else if (p->mode == YB_MODE_FP16) {
            if (!tc_supported(a)) { err = "plan: layer " + L.key + " not supported by the tensor-core kernel"; return false; }
            op.use_tc = true;
            std::string e;
            if (halo_supported(a)) {
                op.use_halo = true;
                e = halo_make_plan(op.halo, a, L.d_w16, L.cout_pad, L.ks * L.ks * L.cin, c->num_sms);
            } else {
                e = tc_make_plan(op.tc, a, L.d_w16, L.cout_pad, L.ks * L.ks * L.cin, c->num_sms);
            }
            if (!e.empty()) { err = "plan: layer " + L.key + ": " + e; return false; }
        }
        p->ops.push_back(op);
        return true;
    }
};

// The stem kernels take the stem's scale / bias (and, in the fp32-grade mode, the per-channel weight shift) as kernel
// parameters: the host mirror is read back from the device blob, which is what the kernels' other operands come from and
// what yb_bcast_weights overwrites.
int refresh_stem_mirror(yb_ctx* c) {
    const Layer& L = c->layers[0];
    YB_CUDA(c, cudaMemcpy(c->stem_sb, L.d_scale, 32 * sizeof(float), cudaMemcpyDeviceToHost));
    YB_CUDA(c, cudaMemcpy(c->stem_sb + 32, L.d_bias, 32 * sizeof(float), cudaMemcpyDeviceToHost));
    if (L.d_w32) {
        float w[27 * 32], bi[32];
        YB_CUDA(c, cudaMemcpy(w, L.d_w32, sizeof(w), cudaMemcpyDeviceToHost));
        stem_split_host_params(w, c->stem_sb, c->stem_sb + 32, c->stem_sc_eff, bi, c->stem_shift);
    }
    return YB_OK;
}

// A/B switch of experiment builds: the first two layers as separate kernels even where the fused kernel applies
bool stem_unfused() {
    static const bool v = tune_env("YB_STEM_UNFUSED") && atoi(tune_env("YB_STEM_UNFUSED")) != 0;
    return v;
}

int build_plan(yb_ctx* c, int B, int H, int W, Plan** out) {
    for (auto& q : c->plans)
        if (q->B == B && q->H == H && q->W == W && q->mode == c->mode) { *out = q.get(); return YB_OK; }
    if (c->plans.size() >= 4) free_plans(c);   // keep the arena bounded when shapes keep changing
    std::unique_ptr<Plan> plan(new Plan());
    Plan* p = plan.get();
    p->B = B; p->H = H; p->W = W; p->mode = c->mode;
    const bool split = c->mode == YB_MODE_FP32_TC;
    PlanBuilder pb{c, p, c->mode == YB_MODE_FP32 ? sizeof(float) : sizeof(__half), "", split};
    const size_t elem = pb.elem * (split ? 2 : 1);      // bytes per stored channel (hi + lo in split mode)
    auto bail = [&](const std::string& m) {
        for (void* q : p->allocs) cudaFree(q);
        return fail(c, m.find("cudaMalloc") != std::string::npos ? YB_E_NOMEM : YB_E_CUDA, m);
    };

    void* buf[2];
    buf[0] = pb.alloc((size_t)B * H * W * 32 * elem);
    buf[1] = pb.alloc((size_t)B * (H / 2) * (W / 2) * 64 * elem);
    void* cat1 = pb.alloc((size_t)B * (H / 16) * (W / 16) * 768 * elem);
    void* cat2 = pb.alloc((size_t)B * (H / 8) * (W / 8) * 384 * elem);
    for (int i = 0; i < 3; ++i) {
        const int s = 32 >> i;
        p->gh[i] = H / s; p->gw[i] = W / s;
        const int cp = c->layers.back().cout_pad;
        p->logits[i] = static_cast<float*>(pb.alloc((size_t)B * p->gh[i] * p->gw[i] * cp * sizeof(float)));
    }
    if (!pb.err.empty()) return bail(pb.err);

    // stem
    bool fused01 = false;
    {
        Op op;
        op.layer = 0; op.stem = true;
        op.a.out = buf[0];
        if (p->mode == YB_MODE_FP16) {
            std::string e;
            // stem + first stride-2 convolution in one kernel (the 32-channel tensor never reaches HBM); W % 8 covers both
            // image element types
            fused01 = stem_block_supported(H, W, 1) && !stem_unfused();
            if (fused01) {
                op.fused01 = true;
                op.a.out = buf[1];
                e = stem_block_make_plan(op.stem_block, c->layers[1].d_w16, static_cast<__half*>(buf[1]), 64, B, H, W, c->num_sms);
            } else {
                e = stem_halo_make_plan(op.stem_halo, static_cast<__half*>(buf[0]), 32, B, H, W, c->num_sms);
            }
            if (!e.empty()) return bail("plan: stem: " + e);
        } else if (p->mode == YB_MODE_FP32_TC && stem_split_supported(W)) {
            op.stem_split_tc = true;
            std::string e = stem_split_make_plan(op.stem_halo, static_cast<__half*>(buf[0]), 64, B, H, W, c->num_sms);
            if (!e.empty()) return bail("plan: stem: " + e);
        }
        p->ops.push_back(op);
    }
    int li = 1, ch = 32, h = H, w = W;
    int curb = 0;                       // physical buffer holding `cur` (-1: concat slice)
    TView cur = pb.view(buf[0], B, h, w, 32, 32);
    for (int s = 0; s < 5; ++s) {
        // down-sampling conv (darknet.py:69)
        const int ob = curb == 0 ? 1 : 0;
        TView o = pb.view(buf[ob], B, h / 2, w / 2, ch * 2, ch * 2);
        if (s == 0 && fused01) ++li;    // layer 1 ran inside the stem kernel
        else if (!pb.conv(li++, cur, o, nullptr, false, false)) return bail(pb.err);
        cur = o; curb = ob; h /= 2; w /= 2; ch *= 2;
        for (int j = 0; j < kBlocks[s]; ++j) {
            const int tb = curb == 0 ? 1 : 0;
            TView t = pb.view(buf[tb], B, h, w, ch / 2, ch / 2);
            if (!pb.conv(li++, cur, t, nullptr, false, false)) return bail(pb.err);
            TView o2 = cur;             // in-place residual: out aliases res element for element
            int ob2 = curb;
            const bool last = j == kBlocks[s] - 1;
            if (last && s == 2) { o2 = pb.view(cat2, B, h, w, 256, 384, 128); ob2 = -1; }   // route 36 (darknet.py:181)
            if (last && s == 3) { o2 = pb.view(cat1, B, h, w, 512, 768, 256); ob2 = -1; }   // route 61 (darknet.py:180)
            if (!pb.conv(li++, t, o2, &cur, false, false)) return bail(pb.err);
            cur = o2; curb = ob2;
        }
    }
    p->backbone_out = cur;
    p->n_backbone_ops = (int)p->ops.size();

    int route_buf = 0;                  // physical buffer holding the route (the other one is free once the head conv ran)
    auto predet = [&](TView x, int xb, int nout, int scale_i, TView* route) -> bool {
        // PreDetectionConvGroup.forward (darknet.py:121-126); route = mlist[4] output
        for (int i = 0; i < 6; ++i) {
            const int ob = xb == 0 ? 1 : 0;
            const int co = i % 2 == 0 ? nout : nout * 2;
            TView o = pb.view(buf[ob], B, x.H, x.W, co, co);
            if (!pb.conv(li++, x, o, nullptr, false, false)) return false;
            x = o; xb = ob;
            if (i == 4) { *route = o; route_buf = ob; }
        }
        const int cp = c->layers[li].cout_pad;
        TView lg;
        lg.p = p->logits[scale_i]; lg.B = B; lg.H = x.H; lg.W = x.W; lg.C = cp; lg.ld = cp;
        return pb.conv(li++, x, lg, nullptr, false, true);
    };
    TView route;
    if (!predet(cur, curb, 512, 0, &route)) return bail(pb.err);
    // UpsampleGroup (darknet.py:159-162): 1x1 conv, nearest x2 into channels [0, C) of the concat buffer.  The tensor-core
    // modes write a plain tensor into the free ping-pong buffer through the staged TMA-store epilogue and replicate it with a
    // copy kernel: the direct 2x2-replicating stores of round 1 (32-byte stores one pixel pitch apart per lane) ran the two
    // layers at 6-10 % tensor-pipe activity, 27 + 34 us (profiles/r02_conv_metrics.csv).  The CUDA-core mode replicates in
    // its epilogue.
    auto up = [&](void* cat, int C, int Ctot) -> bool {
        TView o = pb.view(cat, B, route.H, route.W, C, Ctot, 0);
        static const bool up_direct = tune_env("YB_UP_DIRECT") && atoi(tune_env("YB_UP_DIRECT")) != 0;   // A/B (experiment builds)
        if (p->mode == YB_MODE_FP32 || (up_direct && !split)) return pb.conv(li++, route, o, nullptr, true, false);
        TView t = pb.view(buf[route_buf == 0 ? 1 : 0], B, route.H, route.W, C, C);
        if (!pb.conv(li++, route, t, nullptr, false, false)) return false;
        Op op;
        op.layer = li - 1; op.upcopy = true; op.up_src = t; op.up_dst = o;
        p->ops.push_back(op);
        return true;
    };
    if (!up(cat1, 256, 768)) return bail(pb.err);
    TView c1 = pb.view(cat1, B, H / 16, W / 16, 768, 768);
    if (!predet(c1, -1, 256, 1, &route)) return bail(pb.err);
    if (!up(cat2, 128, 384)) return bail(pb.err);
    TView c2 = pb.view(cat2, B, H / 8, W / 8, 384, 384);
    if (!predet(c2, -1, 128, 2, &route)) return bail(pb.err);
    if (li != (int)c->layers.size()) return bail("plan: internal layer count mismatch");

    *out = p;
    c->plans.push_back(std::move(plan));
    return YB_OK;
}

constexpr int kProfBank = 16;



cudaError_t launch_op(yb_ctx* c, Plan* p, Op& op, cudaStream_t s) {
    const Layer& L = c->layers[op.layer];
    if (op.upcopy)
        return launch_upsample2x(static_cast<const __half*>(op.up_src.p), op.up_src.ld, op.up_src.lo, static_cast<__half*>(op.up_dst.p),
                                 op.up_dst.ld, op.up_dst.lo, op.up_src.C, p->B, op.up_src.H, op.up_src.W, p->mode == YB_MODE_FP32_TC ? 2 : 1, s);
    if (op.use_halo) return halo_launch(op.halo, op.a, c->dbg, s);
    if (op.use_tc) return tc_launch(op.tc, op.a, c->dbg, s);
    return launch_conv_simt<float>(op.a, L.d_w32, L.cout_pad, s);
}

// ops [1, n_ops) with plain stream launches
int run_ops_tail(yb_ctx* c, Plan* p, int n_ops, cudaStream_t s) {
    for (int i = 1; i < n_ops; ++i) {
        const cudaError_t e = launch_op(c, p, p->ops[i], s);
        if (e != cudaSuccess) return fail(c, YB_E_CUDA, "launch of layer " + c->layers[p->ops[i].layer].key + ": " + cudaGetErrorString(e));
        ++c->launches;
    }
    return YB_OK;
}

int run_ops(yb_ctx* c, Plan* p, const float* x, int n_ops, cudaStream_t s) {
    const bool deferred = c->profiling && c->profile_deferred;
    const bool prof = c->profiling && !deferred;
    if (deferred) {
        while ((int)c->bank.size() < kProfBank * 4) {
            cudaEvent_t e;
            YB_CUDA(c, cudaEventCreate(&e));
            c->bank.push_back(e);
        }
        YB_CUDA(c, cudaEventRecord(c->bank[c->bank_next * 4 + 0], s));
    }
    if (prof) {
        while ((int)c->ev.size() < (int)p->ops.size() + 4) {
            cudaEvent_t e;
            YB_CUDA(c, cudaEventCreate(&e));
            c->ev.push_back(e);
        }
        YB_CUDA(c, cudaEventRecord(c->ev[0], s));
    }
    // Launch-bound shapes (a single 416x416 image: 80 dependent kernels of a few microseconds each) replay ops [1, n_ops)
    // as one CUDA graph: the host then issues two launches instead of eighty.  Large batches keep the stream launches --
    // the GPU is the bottleneck there and programmatic dependent launch already overlaps the kernels' prologues.
    const bool want_graph = !c->profiling && n_ops > 1 &&
                            (c->graph_mode == 1 || (c->graph_mode == 2 && (long)p->B * p->H * p->W <= (1L << 20)));
    Plan::GraphSlot* slot = nullptr;
    if (want_graph) {
        for (auto& g : p->graphs)
            if (g.n_ops == n_ops || g.n_ops == 0) { slot = &g; break; }
        if (slot) { slot->n_ops = n_ops; ++slot->calls; }
        if (slot && slot->failed) slot = nullptr;
    }
    bool capturing = false;
    for (int i = 0; i < n_ops; ++i) {
        Op& op = p->ops[i];
        const Layer& L = c->layers[op.layer];
        cudaError_t e;
        if (i == 1 && slot) {
            if (slot->exec) {                              // replay
                e = cudaGraphLaunch(slot->exec, s);
                if (e != cudaSuccess) return fail(c, YB_E_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
                c->launches += n_ops - 1;                  // kernels of this library executed by the graph
                ++c->graph_replays;
                break;
            }
            if (slot->calls >= 2) {                        // capture the remaining launches of this call
                e = c->cap_stream ? cudaSuccess : cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal);
                if (e == cudaSuccess) capturing = true;
                else { cudaGetLastError(); slot->failed = true; }
            }
        }
        if (op.stem) {
            if (c->input_f16 && p->mode != YB_MODE_FP16)
                return fail(c, YB_E_UNSUPPORTED, "fp16 input images need YB_MODE_FP16");
            if (p->mode == YB_MODE_FP16 && op.fused01)
                e = stem_block_launch(op.stem_block, x, c->input_f16, p->B, p->H, p->W, L.d_w16, c->stem_sb, c->layers[1].d_scale,
                                      c->layers[1].d_bias, c->dbg, s);
            else if (p->mode == YB_MODE_FP16)
                e = stem_halo_launch(op.stem_halo, x, c->input_f16, p->B, p->H, p->W, L.d_w16, c->stem_sb, c->dbg, s);
            else if (p->mode == YB_MODE_FP32_TC && op.stem_split_tc)
                e = stem_split_launch(op.stem_halo, x, p->B, p->H, p->W, L.d_w32, c->stem_sc_eff, c->stem_sb + 32, c->stem_shift, c->dbg, s);
            else if (p->mode == YB_MODE_FP32_TC)     // widths TMA cannot read: exact fp32 FMAs on the CUDA cores, output written as hi | lo
                e = launch_stem_split(x, static_cast<__half*>(op.a.out), L.d_w32, L.d_scale, L.d_bias, p->B, p->H, p->W, s);
            else
                e = launch_stem<float>(x, static_cast<float*>(op.a.out), L.d_w32, L.d_scale, L.d_bias, p->B, p->H, p->W, s);
        } else {
            e = launch_op(c, p, op, capturing ? c->cap_stream : s);     // (nothing executes while capturing)
        }
        if (e != cudaSuccess) {
            if (capturing) { cudaGraph_t g = nullptr; cudaStreamEndCapture(c->cap_stream, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); }
            return fail(c, YB_E_CUDA, "launch of layer " + L.key + ": " + cudaGetErrorString(e));
        }
        if (!capturing) ++c->launches;
        if (prof && (c->profile_layers || i == n_ops - 1)) YB_CUDA(c, cudaEventRecord(c->ev[i + 1], s));
        if (deferred && i == n_ops - 1) YB_CUDA(c, cudaEventRecord(c->bank[c->bank_next * 4 + 1], s));
    }
    if (capturing) {
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->cap_stream, &g);
        if (e == cudaSuccess && g) e = cudaGraphInstantiate(&slot->exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e == cudaSuccess && slot->exec) e = cudaGraphLaunch(slot->exec, s);
        if (e != cudaSuccess || !slot->exec) {
            // capture is an optimisation: on any failure fall back to eager launches for good and run this call eagerly
            cudaGetLastError();
            if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
            slot->failed = true;
            const int saved = c->graph_mode;
            c->graph_mode = 0;
            const int rc = run_ops_tail(c, p, n_ops, s);
            c->graph_mode = saved;
            return rc;
        }
        c->launches += n_ops - 1;
    }
    return YB_OK;
}

int check_shape(yb_ctx* c, int B, int H, int W) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!c->finalized) return fail(c, YB_E_STATE, "yb_finalize must be called before the forward path");
    if (B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32)
        return fail(c, YB_E_ARG, "B must be > 0 and H, W positive multiples of 32 (got B=" + std::to_string(B) + " H=" +
                                     std::to_string(H) + " W=" + std::to_string(W) + ")");
    return YB_OK;
}

void fill_decode(yb_ctx* c, int H, int W, const float* l0, const float* l1, const float* l2, long ld, DecodeScale sc[3]) {
    const float* ls[3] = {l0, l1, l2};
    int row = 0;
    for (int i = 0; i < 3; ++i) {
        const int s = 32 >> i;
        sc[i].logits = ls[i]; sc[i].ld = ld;
        sc[i].h = H / s; sc[i].w = W / s;
        sc[i].row_off = row;
        row += 3 * sc[i].h * sc[i].w;
        // stride = img_dim[1] / nH as a python float, anchors/stride in fp32 (yololayer.py:36-38)
        const double stride = (double)H / sc[i].h;
        sc[i].stride = (float)stride;
        for (int a = 0; a < 3; ++a) {
            sc[i].aw[a] = c->anchors[2 * kMasks[i][a]] / (float)stride;
            sc[i].ah[a] = c->anchors[2 * kMasks[i][a] + 1] / (float)stride;
        }
    }
}

int total_rows(int H, int W) { return 3 * ((H / 32) * (W / 32) + (H / 16) * (W / 16) + (H / 8) * (W / 8)); }

int record_sections(yb_ctx* c, int n_ops, bool have_decode, bool have_post, cudaStream_t s) {
    // events: [0]=start, [1..n_ops]=after each op, [n_ops+1]=after decode, [n_ops+2]=after post
    YB_CUDA(c, cudaStreamSynchronize(s));
    c->layer_ms.assign(c->profile_layers ? n_ops : 0, 0.f);
    if (c->profile_layers)
        for (int i = 0; i < n_ops; ++i) YB_CUDA(c, cudaEventElapsedTime(&c->layer_ms[i], c->ev[i], c->ev[i + 1]));
    YB_CUDA(c, cudaEventElapsedTime(&c->sec_ms[0], c->ev[0], c->ev[n_ops]));
    c->sec_ms[1] = c->sec_ms[2] = 0.f;
    if (have_decode) YB_CUDA(c, cudaEventElapsedTime(&c->sec_ms[1], c->ev[n_ops], c->ev[n_ops + 1]));
    if (have_post) YB_CUDA(c, cudaEventElapsedTime(&c->sec_ms[2], c->ev[n_ops + 1], c->ev[n_ops + 2]));
    return YB_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int yb_create(yb_ctx** out, int device, int num_classes, const float anchors[18]) {
    if (!out || num_classes <= 0 || num_classes >= 1024) return fail(nullptr, YB_E_ARG, "yb_create: bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, YB_E_CUDA, std::string("yb_create: no CUDA device (") + cudaGetErrorString(e) +
                                            "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, YB_E_ARG, "yb_create: device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, YB_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    yb_ctx* c = new yb_ctx();
    c->device = device;
    c->num_classes = num_classes;
    c->attrs = 5 + num_classes;
    std::memcpy(c->anchors, anchors ? anchors : kDefaultAnchors, sizeof(c->anchors));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->num_sms = prop.multiProcessorCount;
    c->cc_major = prop.major;
    build_layers(c);
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->dbg_host), 64 * sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
        std::memset(c->dbg_host, 0, 64 * sizeof(int));
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->dbg), c->dbg_host, 0) != cudaSuccess) c->dbg = nullptr;
    }
    cudaGetLastError();
    *out = c;
    return YB_OK;
}

void yb_destroy(yb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_plans(c);
    free_post(c);
    cudaFree(c->d_blob);
    cudaFree(c->det_scratch);
    cudaFree(c->box_params);
    cudaFree(c->lb_params);
    cudaFreeHost(c->dbg_host);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->bank) cudaEventDestroy(e);
    if (c->nccl_comm) comm_destroy(c->nccl_comm);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    delete c;
}

const char* yb_last_error(const yb_ctx* c) { return c ? c->err.c_str() : g_last_error.c_str(); }

int yb_num_tensors(const yb_ctx* c) { return c ? (int)c->keys.size() : 0; }

const char* yb_tensor_key(const yb_ctx* c, int i, size_t* numel) {
    if (!c || i < 0 || i >= (int)c->keys.size()) return nullptr;
    if (numel) *numel = c->keys[i].second.numel;
    return c->keys[i].first.c_str();
}

int yb_set_tensor(yb_ctx* c, const char* key, const float* data, size_t n, int on_host) {
    if (!c || !key || !data) return fail(c, YB_E_ARG, "yb_set_tensor: null argument");
    auto it = c->key_index.find(key);
    if (it == c->key_index.end()) return fail(c, YB_E_KEY, std::string("yb_set_tensor: unknown key '") + key + "'");
    const auto& slot = c->keys[it->second].second;
    if (slot.field == 5) return YB_OK;    // num_batches_tracked: accepted, unused in eval
    if (n != slot.numel)
        return fail(c, YB_E_KEY, std::string("yb_set_tensor: '") + key + "' expects " + std::to_string(slot.numel) +
                                     " elements, got " + std::to_string(n));
    std::vector<float>* v = field_vec(c->layers[slot.layer], slot.field);
    if (on_host) {
        std::memcpy(v->data(), data, n * sizeof(float));
    } else {
        // the caller's tensor may have been produced on any stream (PyTorch side streams are non-blocking: the legacy
        // default stream this copy runs on does not order against them), so wait for the device first.  Not a hot path.
        YB_CUDA(c, cudaSetDevice(c->device));
        YB_CUDA(c, cudaDeviceSynchronize());
        YB_CUDA(c, cudaMemcpy(v->data(), data, n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    c->finalized = false;
    return YB_OK;
}

int yb_get_tensor(const yb_ctx* c, const char* key, float* host_out, size_t n) {
    if (!c || !key || !host_out) return fail(c, YB_E_ARG, "yb_get_tensor: null argument");
    auto it = c->key_index.find(key);
    if (it == c->key_index.end()) return fail(c, YB_E_KEY, std::string("yb_get_tensor: unknown key '") + key + "'");
    const auto& slot = c->keys[it->second].second;
    if (slot.field == 5) { if (n) host_out[0] = 0.f; return YB_OK; }
    if (n != slot.numel) return fail(c, YB_E_KEY, std::string("yb_get_tensor: size mismatch for '") + key + "'");
    const std::vector<float>* v = field_vec(const_cast<Layer&>(c->layers[slot.layer]), slot.field);
    std::memcpy(host_out, v->data(), n * sizeof(float));
    return YB_OK;
}

int yb_load_darknet_blob(yb_ctx* c, const float* host, size_t nfloats, int backbone_only, size_t* consumed) {
    if (!c || !host) return fail(c, YB_E_ARG, "yb_load_darknet_blob: null argument");
    size_t ptr = 0;
    auto take = [&](std::vector<float>& v) -> bool {
        if (ptr + v.size() > nfloats) return false;
        std::memcpy(v.data(), host + ptr, v.size() * sizeof(float));
        ptr += v.size();
        return true;
    };
    for (Layer& L : c->layers) {
        if (backbone_only && L.key.compare(0, 8, "feature.") != 0) break;
        bool ok;
        if (L.bn) ok = take(L.b) && take(L.g) && take(L.mean) && take(L.var) && take(L.w);   // darknet.py:279-285
        else ok = take(L.bias) && take(L.w);                                                  // darknet.py:287-290
        if (!ok) {
            if (consumed) *consumed = ptr;
            return fail(c, YB_E_ARG, "yb_load_darknet_blob: stream ends inside layer " + L.key);
        }
    }
    if (consumed) *consumed = ptr;
    c->finalized = false;
    return YB_OK;
}

int yb_save_darknet_blob(const yb_ctx* c, float* host_out, size_t capacity, int backbone_only, size_t* written) {
    if (!c) return fail(c, YB_E_ARG, "yb_save_darknet_blob: null ctx");
    size_t ptr = 0;
    bool overflow = false;
    auto put = [&](const std::vector<float>& v) {
        if (host_out) {
            if (ptr + v.size() > capacity) overflow = true;
            else std::memcpy(host_out + ptr, v.data(), v.size() * sizeof(float));
        }
        ptr += v.size();
    };
    for (const Layer& L : c->layers) {
        if (backbone_only && L.key.compare(0, 8, "feature.") != 0) break;
        if (L.bn) { put(L.b); put(L.g); put(L.mean); put(L.var); put(L.w); }
        else { put(L.bias); put(L.w); }
    }
    if (written) *written = ptr;
    if (overflow) return fail(c, YB_E_CAP, "yb_save_darknet_blob: capacity too small");
    return YB_OK;
}

int yb_finalize(yb_ctx* c, int mode) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (mode != YB_MODE_FP32 && mode != YB_MODE_FP16 && mode != YB_MODE_FP32_TC) return fail(c, YB_E_ARG, "yb_finalize: unknown precision mode");
    YB_CUDA(c, cudaSetDevice(c->device));
    if (mode != YB_MODE_FP32 && c->cc_major != 10)
        return fail(c, YB_E_UNSUPPORTED, "the tensor-core modes need an sm_100 (B200) device: the tcgen05/TMA kernels have no fallback");
    const bool split = mode == YB_MODE_FP32_TC;
    if (mode != c->mode) free_plans(c);
    // one contiguous device blob (so multi-GPU replication is a single broadcast): per layer
    // scale[cout_pad], bias[cout_pad], then the packed weights of the mode (the stem always keeps fp32)
    size_t off = 0;
    auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    struct Off { size_t scale, bias, w32, w16; };
    std::vector<Off> offs(c->layers.size());
    for (size_t i = 0; i < c->layers.size(); ++i) {
        const Layer& L = c->layers[i];
        const size_t K = (size_t)L.ks * L.ks * L.cin;
        offs[i].scale = place(sizeof(float) * L.cout_pad);
        offs[i].bias = place(sizeof(float) * L.cout_pad);
        const bool need32 = mode == YB_MODE_FP32 || i == 0;
        offs[i].w32 = need32 ? place(sizeof(float) * K * L.cout_pad) : (size_t)-1;
        // the stem's fp16 weights are [32][K padded to 32] for the tensor-core stem kernel
        offs[i].w16 = mode == YB_MODE_FP16 ? place(sizeof(__half) * (i == 0 ? 32 : K) * L.cout_pad)
                      : (split && i != 0) ? place(sizeof(__half) * 2 * K * L.cout_pad) : (size_t)-1;
    }
    if (off != c->blob_bytes || !c->d_blob) {
        cudaFree(c->d_blob);
        c->d_blob = nullptr;
        YB_CUDA(c, cudaMalloc(&c->d_blob, off));
        c->blob_bytes = off;
        free_plans(c);   // plans hold pointers into the blob
    }
    std::vector<unsigned char> host(off, 0);
    for (size_t i = 0; i < c->layers.size(); ++i) {
        Layer& L = c->layers[i];
        const int K = L.ks * L.ks * L.cin, taps = L.ks * L.ks;
        float* sc = reinterpret_cast<float*>(host.data() + offs[i].scale);
        float* bi = reinterpret_cast<float*>(host.data() + offs[i].bias);
        for (int n = 0; n < L.cout_pad; ++n) { sc[n] = n < L.cout ? 1.f : 0.f; bi[n] = 0.f; }
        for (int n = 0; n < L.cout; ++n) {
            if (L.bn) {
                // eval BatchNorm as ATen's CPU path folds it: alpha = invstd*weight, beta = bias - mean*alpha
                const float invstd = 1.0f / std::sqrt(L.var[n] + kBnEps);
                const float alpha = invstd * L.g[n];
                sc[n] = alpha;
                bi[n] = L.b[n] - L.mean[n] * alpha;
            } else {
                bi[n] = L.bias[n];
            }
        }
        if (offs[i].w32 != (size_t)-1) {
            float* w = reinterpret_cast<float*>(host.data() + offs[i].w32);   // [tap][cin][cout_pad]
            for (int n = 0; n < L.cout; ++n)
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int t = 0; t < taps; ++t)
                        w[((size_t)t * L.cin + ci) * L.cout_pad + n] = L.w[((size_t)n * L.cin + ci) * taps + t];
        }
        if (offs[i].w16 != (size_t)-1 && split) {
            // [cout_pad][tap][channel block][wh | wl][bke] (bke = 64 channels, 32 when Cin = 32: the k-block unit order of
            // conv_tc_kernel's split mode): row n is scaled by 2^s(n) so that max|w'| lies in [2048, 4096) -- the lo part of
            // every weight down to 2^-23 of the row maximum is then a NORMAL fp16 number -- and the epilogue scale undoes it
            // exactly (a power of two).  wh = RN16(w'), wl = RN16(w' - wh).
            const int bke = L.cin == 32 ? 32 : 64;
            __half* w = reinterpret_cast<__half*>(host.data() + offs[i].w16);
            for (int n = 0; n < L.cout; ++n) {
                float mx = 0.f;
                for (size_t k = 0; k < (size_t)L.cin * taps; ++k) mx = std::max(mx, std::fabs(L.w[(size_t)n * L.cin * taps + k]));
                int sh = 0;
                if (mx > 0.f && std::isfinite(mx)) {
                    int e2 = 0;
                    std::frexp(mx, &e2);                         // mx = f * 2^e2, f in [0.5, 1)
                    sh = std::max(-24, std::min(40, 12 - e2));
                }
                sc[n] = std::ldexp(sc[n], -sh);
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int t = 0; t < taps; ++t) {
                        const float ws = std::ldexp(L.w[((size_t)n * L.cin + ci) * taps + t], sh);
                        const __half wh = __float2half_rn(ws);
                        const __half wl = __float2half_rn(ws - __half2float(wh));
                        __half* row = w + (size_t)n * 2 * K + ((size_t)t * (L.cin / bke) + ci / bke) * 2 * bke + ci % bke;
                        row[0] = wh; row[bke] = wl;
                    }
            }
        } else if (offs[i].w16 != (size_t)-1) {
            __half* w = reinterpret_cast<__half*>(host.data() + offs[i].w16);  // [cout_pad][tap][cin]
            const int Kp = i == 0 ? 32 : K;
            for (int n = 0; n < L.cout; ++n)
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int t = 0; t < taps; ++t)
                        w[(size_t)n * Kp + (size_t)t * L.cin + ci] = __float2half_rn(L.w[((size_t)n * L.cin + ci) * taps + t]);
        }
        L.d_scale = reinterpret_cast<float*>(c->d_blob + offs[i].scale);
        L.d_bias = reinterpret_cast<float*>(c->d_blob + offs[i].bias);
        L.d_w32 = offs[i].w32 != (size_t)-1 ? reinterpret_cast<float*>(c->d_blob + offs[i].w32) : nullptr;
        L.d_w16 = offs[i].w16 != (size_t)-1 ? reinterpret_cast<__half*>(c->d_blob + offs[i].w16) : nullptr;
    }
    // a forward enqueued on a non-blocking stream may still be reading the blob (weights, scale / bias tables): drain the
    // device before overwriting it, and after, so that no later launch on any stream can race the upload
    YB_CUDA(c, cudaDeviceSynchronize());
    YB_CUDA(c, cudaMemcpy(c->d_blob, host.data(), off, cudaMemcpyHostToDevice));
    YB_CUDA(c, cudaDeviceSynchronize());
    int rc_mirror = YB_OK;
    rc_mirror = refresh_stem_mirror(c);
    if (rc_mirror) return rc_mirror;
    c->mode = mode;
    c->finalized = true;
    return YB_OK;
}

int yb_forward(yb_ctx* c, const float* x, int B, int H, int W, float* det, void* stream) {
    int rc = check_shape(c, B, H, W);
    if (rc) return rc;
    if (!x || !det) return fail(c, YB_E_ARG, "yb_forward: null tensor");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    YB_CUDA(c, cudaSetDevice(c->device));
    Plan* p;
    if ((rc = build_plan(c, B, H, W, &p))) return rc;
    if ((rc = run_ops(c, p, x, (int)p->ops.size(), s))) return rc;
    DecodeScale sc[3];
    fill_decode(c, H, W, p->logits[0], p->logits[1], p->logits[2], c->layers.back().cout_pad, sc);
    // decode: the flat-map kernel; YB_DECODE_V2=1 selects the one-warp-per-cell kernel (coalesced row stores, but one
    // cell in flight per warp: measured 0.194 ms against 0.170 ms at 608x608 batch 32, profiles/README.md)
    static const bool decode_v2 = tune_env("YB_DECODE_V2") && atoi(tune_env("YB_DECODE_V2")) != 0;
    if (decode_v2)
        YB_CUDA(c, launch_decode_cells(sc, B, c->attrs, total_rows(H, W), 1, det, 0.f, nullptr, nullptr, nullptr, nullptr, c->num_sms, s));
    else
        YB_CUDA(c, launch_decode(sc, 0, B, c->attrs, total_rows(H, W), det, s));
    ++c->launches;
    if (c->profiling && !c->profile_deferred) {
        YB_CUDA(c, cudaEventRecord(c->ev[p->ops.size() + 1], s));
        return record_sections(c, (int)p->ops.size(), true, false, s);
    }
    return YB_OK;
}

int yb_forward_logits(yb_ctx* c, const float* x, int B, int H, int W, float* l32, float* l16, float* l8, void* stream) {
    int rc = check_shape(c, B, H, W);
    if (rc) return rc;
    if (!x || !l32 || !l16 || !l8) return fail(c, YB_E_ARG, "yb_forward_logits: null tensor");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    YB_CUDA(c, cudaSetDevice(c->device));
    Plan* p;
    if ((rc = build_plan(c, B, H, W, &p))) return rc;
    if ((rc = run_ops(c, p, x, (int)p->ops.size(), s))) return rc;
    float* outs[3] = {l32, l16, l8};
    const int cp = c->layers.back().cout_pad;
    for (int i = 0; i < 3; ++i) {
        YB_CUDA(c, launch_nhwc_to_nchw_f32<float>(p->logits[i], cp, 3 * c->attrs, B, p->gh[i] * p->gw[i], outs[i], s));
        ++c->launches;
    }
    if (c->profiling && !c->profile_deferred) {
        YB_CUDA(c, cudaEventRecord(c->ev[p->ops.size() + 1], s));
        return record_sections(c, (int)p->ops.size(), false, false, s);
    }
    return YB_OK;
}

int yb_backbone(yb_ctx* c, const float* x, int B, int H, int W, float* feat, void* stream) {
    int rc = check_shape(c, B, H, W);
    if (rc) return rc;
    if (!x || !feat) return fail(c, YB_E_ARG, "yb_backbone: null tensor");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    YB_CUDA(c, cudaSetDevice(c->device));
    Plan* p;
    if ((rc = build_plan(c, B, H, W, &p))) return rc;
    if ((rc = run_ops(c, p, x, p->n_backbone_ops, s))) return rc;
    const TView& v = p->backbone_out;
    if (p->mode == YB_MODE_FP32_TC)
        YB_CUDA(c, launch_split_to_nchw_f32(static_cast<const __half*>(v.p), v.ld, v.lo, v.C, B, v.H * v.W, feat, s));
    else if (p->mode == YB_MODE_FP16)
        YB_CUDA(c, launch_nhwc_to_nchw_f32<__half>(static_cast<const __half*>(v.p), v.ld, v.C, B, v.H * v.W, feat, s));
    else
        YB_CUDA(c, launch_nhwc_to_nchw_f32<float>(static_cast<const float*>(v.p), v.ld, v.C, B, v.H * v.W, feat, s));
    ++c->launches;
    if (c->profiling && !c->profile_deferred) {
        YB_CUDA(c, cudaEventRecord(c->ev[p->n_backbone_ops + 1], s));
        return record_sections(c, p->n_backbone_ops, false, false, s);
    }
    return YB_OK;
}

int yb_decode(yb_ctx* c, const float* l32, const float* l16, const float* l8, int B, int H, int W, float* det, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32) return fail(c, YB_E_ARG, "yb_decode: H and W must be multiples of 32, B > 0");
    if (!l32 || !l16 || !l8 || !det) return fail(c, YB_E_ARG, "yb_decode: null tensor");
    YB_CUDA(c, cudaSetDevice(c->device));
    DecodeScale sc[3];
    fill_decode(c, H, W, l32, l16, l8, 0, sc);
    YB_CUDA(c, launch_decode(sc, 1, B, c->attrs, total_rows(H, W), det, static_cast<cudaStream_t>(stream)));
    ++c->launches;
    return YB_OK;
}

int yb_postprocess(yb_ctx* c, const float* det, int B, int N, float conf, float nms, int is_eval, int use_nms,
                   float* rows7, int* counts, int* src_index, int* cand_counts, int cap, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!det || !rows7 || !counts || B <= 0 || N <= 0 || cap <= 0) return fail(c, YB_E_ARG, "yb_postprocess: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_post(c, B, N, is_eval);
    if (rc) return rc;
    PostArgs a{det, B, N, c->num_classes, conf, nms, is_eval, use_nms, rows7, counts, src_index, cand_counts, cap};
    YB_CUDA(c, launch_postprocess(a, c->post, &c->launches, static_cast<cudaStream_t>(stream)));
    return YB_OK;
}

int yb_postprocess_notebook(yb_ctx* c, const float* det, int B, int N, float conf, float nms, float* rows7, int* counts,
                            int* src_index, int* cand_counts, int cap, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!det || !rows7 || !counts || B <= 0 || N <= 0 || cap <= 0) return fail(c, YB_E_ARG, "yb_postprocess_notebook: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_post(c, B, N, 0);
    if (rc) return rc;
    PostArgs a{det, B, N, c->num_classes, conf, nms, 0, 1, rows7, counts, src_index, cand_counts, cap};
    a.variant = 1;
    YB_CUDA(c, launch_postprocess(a, c->post, &c->launches, static_cast<cudaStream_t>(stream)));
    return YB_OK;
}

int yb_detect(yb_ctx* c, const float* x, int B, int H, int W, float conf, float nms, int is_eval, int use_nms,
              float* rows7, int* counts, int* src_index, int* cand_counts, int cap, void* stream) {
    int rc = check_shape(c, B, H, W);
    if (rc) return rc;
    if (!x || !rows7 || !counts || cap <= 0) return fail(c, YB_E_ARG, "yb_detect: bad arguments");
    const int N = total_rows(H, W);
    // Non-eval mode never materialises the [B,N,5+C] tensor: the decode kernel scores each cell's three rows while
    // they are in shared memory and hands pp_scan / pp_scatter the per-row candidates directly.  Eval mode (one
    // candidate per passing (box, class) pair, gathered from det by pp_scatter) keeps the two-kernel path.
    static const bool fused_ok = !(tune_env("YB_FUSED_DETECT") && atoi(tune_env("YB_FUSED_DETECT")) == 0);
    const bool fused = fused_ok && !is_eval;
    if (!fused) {
        const size_t need = sizeof(float) * (size_t)B * N * c->attrs;
        if (need > c->det_scratch_bytes) {
            cudaFree(c->det_scratch);
            c->det_scratch = nullptr;
            c->det_scratch_bytes = 0;
            YB_CUDA(c, cudaMalloc(&c->det_scratch, need));
            c->det_scratch_bytes = need;
        }
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    YB_CUDA(c, cudaSetDevice(c->device));
    Plan* p = nullptr;
    if ((rc = build_plan(c, B, H, W, &p))) return rc;
    if ((rc = ensure_post(c, B, N, is_eval))) return rc;
    const int n_ops = (int)p->ops.size();
    if ((rc = run_ops(c, p, x, n_ops, s))) return rc;       // records events [0] and [n_ops] when profiling
    DecodeScale sc[3];
    fill_decode(c, H, W, p->logits[0], p->logits[1], p->logits[2], c->layers.back().cout_pad, sc);
    if (fused) {
        // the live-cell list borrows rowoff (B*N ints >= 1 + cells; pp_scan overwrites it afterwards)
        int extra = 0;
        YB_CUDA(c, launch_decode_cells(sc, B, c->attrs, N, 2, nullptr, conf, c->post.rowcount, c->post.rowcand, c->post.rowoff, &extra,
                                       c->num_sms, s));
        c->launches += extra;
    } else {
        YB_CUDA(c, launch_decode(sc, 0, B, c->attrs, N, c->det_scratch, s));
    }
    ++c->launches;
    const bool deferred = c->profiling && c->profile_deferred;
    if (deferred) YB_CUDA(c, cudaEventRecord(c->bank[c->bank_next * 4 + 2], s));
    else if (c->profiling) YB_CUDA(c, cudaEventRecord(c->ev[n_ops + 1], s));
    PostArgs a{fused ? nullptr : c->det_scratch, B, N, c->num_classes, conf, nms, is_eval, use_nms, rows7, counts, src_index,
               cand_counts, cap};
    a.pre_scored = fused ? 1 : 0;
    YB_CUDA(c, launch_postprocess(a, c->post, &c->launches, s));
    if (deferred) {
        YB_CUDA(c, cudaEventRecord(c->bank[c->bank_next * 4 + 3], s));
        c->bank_next = (c->bank_next + 1) % kProfBank;
        c->bank_count = std::min(c->bank_count + 1, kProfBank);
        return YB_OK;
    }
    if (c->profiling) {
        YB_CUDA(c, cudaEventRecord(c->ev[n_ops + 2], s));
        return record_sections(c, n_ops, true, true, s);
    }
    return YB_OK;
}

int yb_correct_boxes(yb_ctx* c, const float* boxes, int row_stride, const int* counts, int B, int cap, const int* org_wh,
                     int img_w, int img_h, int is_letterbox, float* out_xywh, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!boxes || !org_wh || !out_xywh || B <= 0 || cap <= 0 || row_stride < 4 || img_w <= 0 || img_h <= 0)
        return fail(c, YB_E_ARG, "yb_correct_boxes: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    if (B > c->box_params_cap) {
        cudaFree(c->box_params);
        c->box_params = nullptr;
        c->box_params_cap = 0;
        YB_CUDA(c, cudaMalloc(&c->box_params, sizeof(float) * 6 * (size_t)B));
        c->box_params_cap = B;
    }
    // host arithmetic exactly as the reference does it in python floats (doubles) and ints
    std::vector<float> prm((size_t)B * 6);
    for (int b = 0; b < B; ++b) {
        const int ow = org_wh[2 * b], oh = org_wh[2 * b + 1];
        if (ow <= 0 || oh <= 0) return fail(c, YB_E_ARG, "yb_correct_boxes: original sizes must be positive");
        double rx, ry;
        long xpad = 0, ypad = 0;
        if (is_letterbox) {                                               // boundingbox.py:106-108
            const double ratio = std::min((double)img_w / ow, (double)img_h / oh);
            const long rw = (long)(ow * ratio), rh = (long)(oh * ratio);
            xpad = (img_w - rw) / 2;
            ypad = (img_h - rh) / 2;
            rx = ry = ratio;
        } else {                                                          // boundingbox.py:130
            rx = (double)img_w / ow;
            ry = (double)img_h / oh;
        }
        float* p = &prm[(size_t)b * 6];
        p[0] = (float)rx; p[1] = (float)ry; p[2] = (float)xpad; p[3] = (float)ypad; p[4] = (float)ow; p[5] = (float)oh;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = stage_params(c, prm.data(), prm.size() * sizeof(float), c->box_params, s)) return rc;
    YB_CUDA(c, launch_correct_boxes(boxes, row_stride, counts, B, cap, c->box_params, out_xywh, s));
    ++c->launches;
    return YB_OK;
}

int yb_resize(yb_ctx* c, const uint8_t* const* imgs_dev, const int* hw_host, int B, int dim_w, int dim_h, float* out_nchw,
              uint8_t* out_hwc, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!imgs_dev || !hw_host || (!out_nchw && !out_hwc) || B <= 0 || dim_w <= 0 || dim_h <= 0) return fail(c, YB_E_ARG, "yb_resize: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    if (B > c->lb_params_cap) {
        cudaFree(c->lb_params);
        c->lb_params = nullptr;
        c->lb_params_cap = 0;
        YB_CUDA(c, cudaMalloc(&c->lb_params, sizeof(LbImage) * (size_t)B));
        c->lb_params_cap = B;
    }
    std::vector<LbImage> prm((size_t)B);
    for (int b = 0; b < B; ++b) {
        const int sh = hw_host[2 * b], sw = hw_host[2 * b + 1];
        if (sh <= 0 || sw <= 0 || !imgs_dev[b]) return fail(c, YB_E_ARG, "yb_resize: image " + std::to_string(b) + " is empty");
        LbImage& q = prm[b];
        q.src = imgs_dev[b]; q.sh = sh; q.sw = sw;
        q.box_w = dim_w; q.box_h = dim_h; q.box_x = 0; q.box_y = 0;      // cv2.resize(img, dim): dsize = (w, h)
        q.scale_x = 1.0 / ((double)dim_w / sw);
        q.scale_y = 1.0 / ((double)dim_h / sh);
        q.interp = 1;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = stage_params(c, prm.data(), prm.size() * sizeof(LbImage), c->lb_params, s)) return rc;
    YB_CUDA(c, launch_letterbox(c->lb_params, B, dim_h, dim_w, out_nchw, out_hwc, s));
    ++c->launches;
    return YB_OK;
}

int yb_letterbox(yb_ctx* c, const uint8_t* const* imgs_dev, const int* hw_host, int B, int dim_w, int dim_h, int canvas_h,
                 int canvas_w, int offset_rule, float* out_nchw, uint8_t* canvas_hwc, float* trans_host, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!imgs_dev || !hw_host || (!out_nchw && !canvas_hwc) || B <= 0 || dim_w <= 0 || dim_h <= 0 || canvas_h <= 0 || canvas_w <= 0)
        return fail(c, YB_E_ARG, "yb_letterbox: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    if (B > c->lb_params_cap) {
        cudaFree(c->lb_params);
        c->lb_params = nullptr;
        c->lb_params_cap = 0;
        YB_CUDA(c, cudaMalloc(&c->lb_params, sizeof(LbImage) * (size_t)B));
        c->lb_params_cap = B;
    }
    // The reference's canvas is np.full(dim + (3,), 128) (utils.py:46), i.e. canvas_h = dim[0] = outer_w and canvas_w =
    // dim[1] = outer_h, while the box is placed with letterbox_transforms' (w, h) reading of dim (utils.py:34-42):
    // identical for the square sizes the reference uses; for other sizes the paste must fit or numpy raises, and so
    // does this call.  The caller states the canvas explicitly.
    std::vector<LbImage> prm((size_t)B);
    for (int b = 0; b < B; ++b) {
        const int sh = hw_host[2 * b], sw = hw_host[2 * b + 1];
        if (sh <= 0 || sw <= 0 || !imgs_dev[b]) return fail(c, YB_E_ARG, "yb_letterbox: image " + std::to_string(b) + " is empty");
        // letterbox_transforms, python-float (double) arithmetic and int() truncation
        const double ratio = std::min((double)dim_w / sw, (double)dim_h / sh);
        const int box_w = (int)(sw * ratio), box_h = (int)(sh * ratio);
        // utils.letterbox_transforms (utils.py:40-41) centres with w//2 - bw//2, IaaLetterbox._compute_height_width_pad
        // (transforms.py:203-210) with (w - bw)//2: one pixel apart when the box size is odd
        const int box_x = offset_rule ? (dim_w - box_w) / 2 : dim_w / 2 - box_w / 2;
        const int box_y = offset_rule ? (dim_h - box_h) / 2 : dim_h / 2 - box_h / 2;
        if (box_w <= 0 || box_h <= 0) return fail(c, YB_E_ARG, "yb_letterbox: image " + std::to_string(b) + " collapses to an empty box");
        if (box_x < 0 || box_y < 0 || box_y + box_h > canvas_h || box_x + box_w > canvas_w)
            return fail(c, YB_E_ARG, "yb_letterbox: the resized image does not fit the canvas (non-square dim, see utils.py:46)");
        LbImage& q = prm[b];
        q.src = imgs_dev[b]; q.sh = sh; q.sw = sw;
        q.box_w = box_w; q.box_h = box_h; q.box_x = box_x; q.box_y = box_y;
        q.scale_x = 1.0 / ((double)box_w / sw);                           // cv::resize: inv_scale = dsize/ssize; scale = 1/inv_scale
        q.scale_y = 1.0 / ((double)box_h / sh);
        q.interp = 0;
        if (trans_host) {
            float* t = trans_host + 5 * (size_t)b;
            t[0] = (float)box_w; t[1] = (float)box_h; t[2] = (float)box_x; t[3] = (float)box_y; t[4] = (float)ratio;
        }
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = stage_params(c, prm.data(), prm.size() * sizeof(LbImage), c->lb_params, s)) return rc;
    YB_CUDA(c, launch_letterbox(c->lb_params, B, canvas_h, canvas_w, out_nchw, canvas_hwc, s));
    ++c->launches;
    return YB_OK;
}

int yb_comm_unique_id(uint8_t id_out[128]) {
    std::string err;
    int rc = comm_unique_id(id_out, err);
    return rc ? fail(nullptr, rc, err) : YB_OK;
}

int yb_comm_init(yb_ctx* c, const uint8_t id[128], int rank, int world) {
    if (!c || !id || world <= 0 || rank < 0 || rank >= world) return fail(c, YB_E_ARG, "yb_comm_init: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    std::string err;
    if (c->nccl_comm) { comm_destroy(c->nccl_comm); c->nccl_comm = nullptr; }   // re-initialisation replaces the communicator
    int rc = comm_init(&c->nccl_comm, id, rank, world, err);
    if (rc) return fail(c, rc, err);
    c->rank = rank; c->world = world;
    return YB_OK;
}

int yb_bcast_weights(yb_ctx* c, int root, void* stream) {
    if (!c || !c->nccl_comm) return fail(c, YB_E_STATE, "yb_bcast_weights: yb_comm_init first");
    if (!c->finalized) return fail(c, YB_E_STATE, "yb_bcast_weights: yb_finalize first (all ranks, same mode)");
    std::string err;
    int rc = comm_bcast(c->nccl_comm, c->d_blob, c->blob_bytes, root, static_cast<cudaStream_t>(stream), err);
    if (rc) return fail(c, rc, err);
    // the halo stem takes its scale / bias as kernel parameters: refresh the host mirror from the received blob
    YB_CUDA(c, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return refresh_stem_mirror(c);
}

int yb_allgather_dets(yb_ctx* c, const float* rows7, const int* counts, int B_local, int cap, float* all_rows,
                      int* all_counts, void* stream) {
    if (!c || !c->nccl_comm) return fail(c, YB_E_STATE, "yb_allgather_dets: yb_comm_init first");
    if (!rows7 || !counts || !all_rows || !all_counts || B_local <= 0 || cap <= 0) return fail(c, YB_E_ARG, "yb_allgather_dets: bad arguments");
    std::string err;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = comm_allgather(c->nccl_comm, rows7, all_rows, sizeof(float) * 7 * (size_t)B_local * cap, s, err);
    if (!rc) rc = comm_allgather(c->nccl_comm, counts, all_counts, sizeof(int) * (size_t)B_local, s, err);
    return rc ? fail(c, rc, err) : YB_OK;
}

int yb_set_input_dtype(yb_ctx* c, int dtype) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (dtype != YB_INPUT_F32 && dtype != YB_INPUT_F16) return fail(c, YB_E_ARG, "yb_set_input_dtype: unknown element type");
    c->input_f16 = dtype == YB_INPUT_F16;
    return YB_OK;
}

int yb_set_graph_mode(yb_ctx* c, int mode) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (mode < 0 || mode > 2) return fail(c, YB_E_ARG, "yb_set_graph_mode: 0 (never), 1 (always) or 2 (auto)");
    c->graph_mode = mode;
    return YB_OK;
}

long long yb_launch_count(const yb_ctx* c) { return c ? c->launches : 0; }

long long yb_graph_replays(const yb_ctx* c) { return c ? c->graph_replays : 0; }

int yb_debug_words(const yb_ctx* c, int* out, int n) {
    if (!c || !out || !c->dbg_host) return 0;
    n = std::min(n, 64);
    for (int i = 0; i < n; ++i) out[i] = c->dbg_host[i];
    return n;
}

int yb_set_profiling(yb_ctx* c, int enabled) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    c->profiling = enabled != 0;
    c->profile_layers = enabled == 2;
    c->profile_deferred = enabled == 3;
    c->bank_next = c->bank_count = 0;
    return YB_OK;
}

int yb_get_section_ms(yb_ctx* c, float* conv_ms, float* decode_ms, float* post_ms) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (c->profile_deferred && c->bank_count > 0) {
        // average over the yb_detect calls recorded since the last query (at most kProfBank), one synchronisation here
        const int last = (c->bank_next + kProfBank - 1) % kProfBank;
        YB_CUDA(c, cudaEventSynchronize(c->bank[last * 4 + 3]));
        float sum[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < c->bank_count; ++k) {
            const int set = (c->bank_next + kProfBank - 1 - k) % kProfBank;
            for (int j = 0; j < 3; ++j) {
                float ms = 0.f;
                YB_CUDA(c, cudaEventElapsedTime(&ms, c->bank[set * 4 + j], c->bank[set * 4 + j + 1]));
                sum[j] += ms;
            }
        }
        for (int j = 0; j < 3; ++j) c->sec_ms[j] = sum[j] / c->bank_count;
        c->bank_count = 0;
    }
    if (conv_ms) *conv_ms = c->sec_ms[0];
    if (decode_ms) *decode_ms = c->sec_ms[1];
    if (post_ms) *post_ms = c->sec_ms[2];
    return YB_OK;
}

int yb_get_layer_ms(yb_ctx* c, float* ms, int capacity) {
    if (!c) return 0;
    const int n = std::min<int>(capacity, (int)c->layer_ms.size());
    for (int i = 0; i < n; ++i) ms[i] = c->layer_ms[i];
    return (int)c->layer_ms.size();
}

int yb_run_layer(yb_ctx* c, int li, const void* in, int B, int H, int W, const void* res, void* out, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!c->finalized) return fail(c, YB_E_STATE, "yb_run_layer: yb_finalize first");
    if (li < 0 || li >= (int)c->layers.size() || !in || !out || B <= 0 || H <= 0 || W <= 0) return fail(c, YB_E_ARG, "yb_run_layer: bad arguments");
    YB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Layer& L = c->layers[li];
    if (c->mode == YB_MODE_FP32_TC) {
        // Unit-test entry of the split mode: fp32 NHWC tensors at the boundary (what the mode stands in for), converted to /
        // from the hi | lo fp16 pairs the kernels work on.  Scratch is allocated per call -- this is not a hot path.
        const bool head = !L.bn;
        const int Ho = H / L.stride, Wo = W / L.stride;
        if (li != 0 && (H % L.stride || W % L.stride)) return fail(c, YB_E_ARG, "yb_run_layer: H, W must be multiples of the stride");
        const size_t Min = (size_t)B * H * W, Mout = (size_t)B * Ho * Wo;
        __half *xin = nullptr, *xout = nullptr, *xres = nullptr;
        auto cleanup = [&]() { cudaFree(xin); cudaFree(xout); cudaFree(xres); };
        if (li != 0) YB_CUDA(c, cudaMalloc(&xin, Min * 2 * L.cin * sizeof(__half)));
        if (!head && cudaMalloc(&xout, Mout * 2 * L.cout * sizeof(__half)) != cudaSuccess) { cleanup(); return fail(c, YB_E_NOMEM, "yb_run_layer: scratch"); }
        if (res && cudaMalloc(&xres, Mout * 2 * L.cout * sizeof(__half)) != cudaSuccess) { cleanup(); return fail(c, YB_E_NOMEM, "yb_run_layer: scratch"); }
        cudaError_t e = cudaSuccess;
        if (li == 0 && stem_split_supported(W)) {
            StemHaloPlan sp;
            std::string err = stem_split_make_plan(sp, xout, 64, B, H, W, c->num_sms);
            if (!err.empty()) { cleanup(); return fail(c, YB_E_CUDA, "yb_run_layer: " + err); }
            e = stem_split_launch(sp, static_cast<const float*>(in), B, H, W, L.d_w32, c->stem_sc_eff, c->stem_sb + 32, c->stem_shift, c->dbg, s);
        } else if (li == 0) {
            e = launch_stem_split(static_cast<const float*>(in), xout, L.d_w32, L.d_scale, L.d_bias, B, H, W, s);
        } else {
            e = launch_f32_to_split(static_cast<const float*>(in), xin, Min, L.cin, s);
            if (e == cudaSuccess && res) e = launch_f32_to_split(static_cast<const float*>(res), xres, Mout, L.cout, s);
            ConvArgs a{};
            a.in = xin; a.in_ld = 2 * L.cin; a.in_lo = L.cin;
            a.out = head ? out : xout; a.out_ld = head ? L.cout_pad : 2 * L.cout; a.out_lo = head ? 0 : L.cout;
            a.res = xres; a.res_ld = 2 * L.cout; a.res_lo = res ? L.cout : 0;
            a.scale = L.d_scale; a.bias = L.d_bias;
            a.B = B; a.H = H; a.W = W; a.Cin = L.cin;
            a.Ho = Ho; a.Wo = Wo; a.Cout = head ? L.cout_pad : L.cout;
            a.ks = L.ks; a.stride = L.stride; a.pad = (L.ks - 1) / 2;
            a.leaky = L.bn ? 1 : 0; a.upsample = 0; a.out_f32 = head ? 1 : 0; a.split = 1;
            if (e == cudaSuccess) {
                if (!tc_supported(a)) { cleanup(); return fail(c, YB_E_UNSUPPORTED, "yb_run_layer: layer not supported by the tensor-core kernel"); }
                TcPlan tp;
                std::string err = tc_make_plan(tp, a, L.d_w16, L.cout_pad, 2 * L.ks * L.ks * L.cin, c->num_sms);
                if (!err.empty()) { cleanup(); return fail(c, YB_E_CUDA, "yb_run_layer: " + err); }
                e = tc_launch(tp, a, c->dbg, s);
            }
        }
        if (e == cudaSuccess && !head) e = launch_split_to_f32(xout, static_cast<float*>(out), Mout, L.cout, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cleanup();
        ++c->launches;
        YB_CUDA(c, e);
        return YB_OK;
    }
    if (li == 0) {
        cudaError_t e;
        if (c->input_f16 && c->mode != YB_MODE_FP16)
            return fail(c, YB_E_UNSUPPORTED, "fp16 input images need YB_MODE_FP16");
        if (c->mode == YB_MODE_FP16) {
            // the image is read by TMA: its row pitch must be a multiple of 16 bytes (the network path has W % 32 == 0)
            if (!stem_halo_supported(W, c->input_f16)) return fail(c, YB_E_ARG, "yb_run_layer: the stem needs W % 4 == 0 (fp32 images) / W % 8 == 0 (fp16)");
            StemHaloPlan sp;
            std::string err = stem_halo_make_plan(sp, static_cast<__half*>(out), 32, B, H, W, c->num_sms);
            if (!err.empty()) return fail(c, YB_E_CUDA, "yb_run_layer: " + err);
            e = stem_halo_launch(sp, in, c->input_f16, B, H, W, L.d_w16, c->stem_sb, c->dbg, s);
        } else {
            e = launch_stem<float>(static_cast<const float*>(in), static_cast<float*>(out), L.d_w32, L.d_scale, L.d_bias, B, H, W, s);
        }
        ++c->launches;
        YB_CUDA(c, e);
        return YB_OK;
    }
    if (H % L.stride || W % L.stride) return fail(c, YB_E_ARG, "yb_run_layer: H, W must be multiples of the stride");
    ConvArgs a{};
    const bool head = !L.bn;
    a.in = in; a.in_ld = L.cin;
    a.out = out; a.out_ld = head ? L.cout_pad : L.cout;
    a.res = res; a.res_ld = L.cout;
    a.scale = L.d_scale; a.bias = L.d_bias;
    a.B = B; a.H = H; a.W = W; a.Cin = L.cin;
    a.Ho = H / L.stride; a.Wo = W / L.stride;
    a.Cout = head ? L.cout_pad : L.cout;
    a.ks = L.ks; a.stride = L.stride; a.pad = (L.ks - 1) / 2;
    a.leaky = L.bn ? 1 : 0; a.upsample = 0; a.out_f32 = head ? 1 : 0;
    if (c->mode == YB_MODE_FP16) {
        if (!tc_supported(a)) return fail(c, YB_E_UNSUPPORTED, "yb_run_layer: layer not supported by the tensor-core kernel");
        if (halo_supported(a)) {
            HaloPlan hp;
            std::string e = halo_make_plan(hp, a, L.d_w16, L.cout_pad, L.ks * L.ks * L.cin, c->num_sms);
            if (!e.empty()) return fail(c, YB_E_CUDA, "yb_run_layer: " + e);
            YB_CUDA(c, halo_launch(hp, a, c->dbg, s));
        } else {
            TcPlan tp;
            std::string e = tc_make_plan(tp, a, L.d_w16, L.cout_pad, L.ks * L.ks * L.cin, c->num_sms);
            if (!e.empty()) return fail(c, YB_E_CUDA, "yb_run_layer: " + e);
            YB_CUDA(c, tc_launch(tp, a, c->dbg, s));
        }
    } else {
        YB_CUDA(c, launch_conv_simt<float>(a, L.d_w32, L.cout_pad, s));
    }
    ++c->launches;
    return YB_OK;
}

int yb_run_stem_block(yb_ctx* c, const void* x, int B, int H, int W, void* out, void* stream) {
    if (!c) return fail(nullptr, YB_E_ARG, "null ctx");
    if (!c->finalized || c->mode != YB_MODE_FP16) return fail(c, YB_E_STATE, "yb_run_stem_block: yb_finalize(YB_MODE_FP16) first");
    if (!x || !out || B <= 0 || H <= 0 || W <= 0) return fail(c, YB_E_ARG, "yb_run_stem_block: bad arguments");
    if (!stem_block_supported(H, W, c->input_f16)) return fail(c, YB_E_UNSUPPORTED, "yb_run_stem_block: shape not supported by the fused kernel");
    YB_CUDA(c, cudaSetDevice(c->device));
    StemBlockPlan sp;
    std::string err = stem_block_make_plan(sp, c->layers[1].d_w16, static_cast<__half*>(out), 64, B, H, W, c->num_sms);
    if (!err.empty()) return fail(c, YB_E_CUDA, "yb_run_stem_block: " + err);
    YB_CUDA(c, stem_block_launch(sp, x, c->input_f16, B, H, W, c->layers[0].d_w16, c->stem_sb, c->layers[1].d_scale, c->layers[1].d_bias,
                                 c->dbg, static_cast<cudaStream_t>(stream)));
    ++c->launches;
    return YB_OK;
}

}  // extern "C"
