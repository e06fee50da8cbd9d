// Host side of libsin3dm_b200: handle, checkpoint ingestion / operand packing, launch plan, C ABI.
// See include/sin3dm_b200.h for the contract and the reference file:line each entry point replaces.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sin3dm_b200.h"
#include "boundary.cuh"
#include "conv_tc.cuh"
#include "host_util.cuh"
#include "kernels.cuh"
#include "train_kernels.cuh"
#include "wgrad_tc.cuh"

using namespace s3d;

// ------------------------------------------------------------------------------------ model description
static const char* kPlane[3] = {"xy", "xz", "yz"};

struct TensorSpec {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> host;
    bool loaded = false;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

struct BlockSpec {       // one TriplaneResBlock (unet_triplane.py:175-311)
    std::string name;
    int cin, cout, level;
    bool has_skip;
    int film_off;        // offset of this block's emb_layers output inside a film row
};
struct OpSpec {
    int kind;            // 0 res, 1 down, 2 up
    int block;           // index into blocks for kind 0
};

struct DevConv3 {        // one 3x3 TriplaneConv (+ optional fused 1x1 skip)
    int C = 0, Cout = 0, Cw = 0, Cs = 0, Ktot = 0;
    float* w_orig[3] = {};      // [Cout][Cw][3][3]
    float* wskip_orig[3] = {};  // [Cout][Cs]
    __half* w_pack[3] = {};     // [2][Cout][Ktot]
    float* wr[3][2] = {};       // rollout 1-D weights [3C][4Cout] (border classes pre-summed) per (plane, group)
    __half* wr16[3][2] = {};    // same as fp16 (hi, lo) [2][4Cout][3C] (K-major B operand of k_roll_tc)
    float* bias[3] = {};        // conv bias (+ skip bias)
    // training: operands of the backward GEMMs, re-packed on the device from w_orig / wskip_orig (k_pack_dgrad)
    __half* wd_pack[3] = {};    // dgrad of the 3x3: [2][C][9*Cout], K = tap'*Cout + co, value W[co][c][2-kh'][2-kw']
    __half* wsd_pack[3] = {};   // dgrad of the 1x1 skip: [2][Cs][Cout]
    float* wrv[3][2] = {};      // rollout adjoint: the broadcast channels' weights as [along*3 + across][Cout][C] (k_pack_rollv)
};
struct DevNorm {
    float* gamma[3] = {};
    float* beta[3] = {};
};
struct DevBlock {
    DevNorm n1, n2;
    DevConv3 c1, c2;
};

struct NamedBuf {
    std::string name;
    TriF p;
    int C;
    TriDims d;
};

struct Plan {
    int B = 0, H = 0, W = 0, D = 0;
    std::vector<std::function<void(cudaStream_t)>> ops;
    std::vector<std::string> op_names;   // kernel:layer, parallel to ops
    std::vector<double> op_flops;        // dense algorithmic FLOPs of the op as the reference executes it (convs only)
    std::vector<Trace> op_trace;         // per-op phase stamps (buf == nullptr unless s3d_unet_trace_enable)
    Trace sched_trace{}, fused_trace{};
    // The fused step boundary runs the NEXT step's in_conv, so after a fused loop the accumulators of in_conv's output hold
    // one contribution too many: whoever launches in_conv stand-alone next clears them first.
    unsigned long long* in_acc = nullptr;
    size_t in_acc_n = 0;
    bool in_acc_stale = false;
    size_t alloc_bytes = 0;
    std::vector<void*> allocs;
    std::vector<NamedBuf> named;
    std::function<void(cudaStream_t, const SchedArgs&)> fused_boundary;   // out head + scheduler + next in_conv (sampling loop)
    float* film_own = nullptr;      // [B][film_dim] used by s3d_unet_forward
    float* emb_tmp[3] = {};         // scratch for the embedding MLP (grown on demand)
    int emb_rows = 0;
    // per-call bindings read by the op lambdas at launch time
    const float* x = nullptr;
    float* out = nullptr;
    const float* film = nullptr;
    const int* film_row = nullptr;
    // sampler-loop state
    int* t_idx = nullptr;           // [B]
    unsigned int* ticket = nullptr;
    float* model_out = nullptr;     // [B][Cout][Hc][Wc]
    // training (s3d_unet_backward)
    bool train = false;
    std::vector<std::pair<void*, size_t>> fwd_zero;     // statistics accumulators cleared before every training forward
    std::vector<std::pair<void*, size_t>> bwd_zero;     // reduction buffers cleared before every backward
    std::vector<std::function<void(cudaStream_t)>> bwd_ops;
    std::vector<std::string> bwd_names;
    std::vector<double> bwd_flops;
    const float* grad_out = nullptr;    // dL/d(out) bound by s3d_unet_backward
    float* grads_own = nullptr;         // flat parameter gradients (state_dict order), copied out at the end
    float* dfilm_own = nullptr;         // [B][film_dim]
    unsigned int* amax = nullptr;       // max |grad_out| (float bits): the loss scale
    unsigned long long* dyn = nullptr;   // [2]: seed, sample_base of the running loop
    std::map<std::string, cudaGraphExec_t> graphs;      // captured steps by (sampler options, buffer pointers)
};

struct s3d_unet {
    s3d_unet_config cfg;
    int device = 0;
    int emb_dim = 0, film_dim = 0;
    std::vector<TensorSpec> tensors;
    std::map<std::string, int> index;
    std::vector<BlockSpec> blocks;
    std::vector<std::vector<OpSpec>> downs, ups;
    std::vector<int> level_ch;      // channels of the stream leaving each level (skip widths)
    bool finalized = false;
    // device weights
    std::vector<void*> wallocs;
    float *te_w0 = nullptr, *te_b0 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr, *film_w = nullptr, *film_b = nullptr;
    float* freqs = nullptr;
    float *in_w[3] = {}, *in_b[3] = {}, *out_w[3] = {}, *out_b[3] = {};
    DevNorm out_norm;
    std::vector<DevBlock> dblocks;
    std::unique_ptr<Plan> plan;
    int last_launches = 0;
    int num_sms = 148;
    bool fuse_roll = true;   // S3D_FUSE_ROLL=0 launches the rollout 1-D GEMM separately
    bool halo_bo_kw = false;
    int roll_tm = kRollTmMax;   // S3D_ROLL_TM: most positions per roll tile (shorter tiles = a shorter roll chain in front of the conv tiles)
    bool fuse_pool = true;      // S3D_FUSE_POOL=0: stand-alone k_avgpool2 instead of pooling in the conv epilogue
    bool trace_on = false;   // s3d_unet_trace_enable
    int profile_mode = -1;   // last s3d_unet_profile_ops: 1 = graph replay with event nodes, 0 = eager launches
    int graph_builds = 0;    // graphs captured + instantiated by s3d_sample_loop
    bool training = false;   // s3d_unet_set_training: plans keep every activation and carry a backward op list
    std::vector<int64_t> grad_off;   // offset of tensor i inside the flat gradient buffer (multiples of 4 floats)
    int64_t grad_numel = 0;
    bool bwd_wgrad_ffma = true;
    void* pack_jobs_dev = nullptr;       // job table of the last s3d_unet_refresh_dev (k_pack_jobs)
    std::vector<char> pack_jobs_host;
};

static bool mode_has_blo(int precision) { return precision == 2 || precision == 3 || precision == 5; }
static int ch_of(const s3d_unet_config& c, int level) { return c.channel_mult[level] * c.model_channels; }

static void add_tensor(s3d_unet* u, const std::string& name, std::vector<int64_t> shape) {
    u->index[name] = static_cast<int>(u->tensors.size());
    TensorSpec t;
    t.name = name;
    t.shape = std::move(shape);
    u->tensors.push_back(std::move(t));
}
static void add_linear(s3d_unet* u, const std::string& n, int i, int o) {
    add_tensor(u, n + ".weight", {o, i});
    add_tensor(u, n + ".bias", {o});
}
static void add_conv(s3d_unet* u, const std::string& n, int i, int o, int k) {
    for (int p = 0; p < 3; ++p) {
        add_tensor(u, n + ".conv_" + kPlane[p] + ".weight", {o, i, k, k});
        add_tensor(u, n + ".conv_" + kPlane[p] + ".bias", {o});
    }
}
static void add_norm(s3d_unet* u, const std::string& n, int c) {
    for (int p = 0; p < 3; ++p) {
        add_tensor(u, n + ".norm_" + kPlane[p] + ".weight", {c});
        add_tensor(u, n + ".norm_" + kPlane[p] + ".bias", {c});
    }
}

// Mirrors the constructor loops of the reference (unet_triplane.py:377-445) for num_res_blocks == 1.
static void build_structure(s3d_unet* u) {
    const auto& c = u->cfg;
    const int L = c.n_levels, mc = c.model_channels, r = c.rollout ? 3 : 1;
    u->emb_dim = 4 * mc;
    add_linear(u, "time_embed.0", mc, u->emb_dim);
    add_linear(u, "time_embed.2", u->emb_dim, u->emb_dim);
    int ch = ch_of(c, 0);
    add_conv(u, "in_conv.0", c.in_channels, ch, 1);
    std::vector<int> chans{ch};
    int film_off = 0;
    auto add_block = [&](const std::string& name, int cin, int cout, int level) {
        BlockSpec b{name, cin, cout, level, cin != cout, film_off};
        film_off += c.use_scale_shift_norm ? 2 * cout : cout;
        add_norm(u, name + ".in_layers.0", cin);
        add_conv(u, name + ".in_layers.2", cin * r, cout, 3);
        add_linear(u, name + ".emb_layers.1", u->emb_dim, c.use_scale_shift_norm ? 2 * cout : cout);
        add_norm(u, name + ".out_layers.0", cout);
        add_conv(u, name + ".out_layers.2", cout * r, cout, 3);
        if (cin != cout) add_conv(u, name + ".skip_connection", cin, cout, 1);
        u->blocks.push_back(b);
        return static_cast<int>(u->blocks.size()) - 1;
    };
    for (int level = 0; level < L; ++level) {
        std::vector<OpSpec> ops;
        int idx = 0;
        if (level != 0) {
            ops.push_back({1, -1});
            idx = 1;
        }
        const int cout = ch_of(c, level);
        ops.push_back({0, add_block("input_blocks." + std::to_string(level) + "." + std::to_string(idx), ch, cout, level)});
        ch = cout;
        chans.push_back(ch);
        u->downs.push_back(ops);
    }
    for (int j = 0; j < L; ++j) {
        const int level = L - 1 - j;
        std::vector<OpSpec> ops;
        int ich = chans.back();
        chans.pop_back();
        if (level == L - 1) ich = 0;
        const int cout = ch_of(c, level);
        ops.push_back({0, add_block("output_blocks." + std::to_string(j) + ".0", ch + ich, cout, level)});
        ch = cout;
        if (level > 0) ops.push_back({2, -1});
        u->ups.push_back(ops);
    }
    u->film_dim = film_off;
    const int c0 = ch_of(c, 0);
    add_norm(u, "out.0", c0);
    add_conv(u, "out.2", c0, c.out_channels, 1);
}

// ------------------------------------------------------------------------------------ device memory helpers
static size_t* g_alloc_counter = nullptr;   // set while a plan is being built
template <typename T>
static T* dev_alloc(std::vector<void*>& owner, size_t n) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    owner.push_back(p);
    if (g_alloc_counter) *g_alloc_counter += n * sizeof(T);
    return static_cast<T*>(p);
}
template <typename T>
static T* dev_upload(std::vector<void*>& owner, const std::vector<T>& h) {
    T* p = dev_alloc<T>(owner, h.size());
    CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}
// Activation arena.  Everything a step touches should stay inside the L2 (126 MB), so activations are carved out of a few
// big chunks with first-fit and handed back as soon as their last consumer has been planned (launches are stream ordered, so
// a later op may overwrite a buffer whose readers were enqueued before it).  S3D_KEEP_ACTS=1 turns the reuse off (bring-up:
// s3d_unet_debug_read then sees every intermediate activation).
struct Arena {
    struct Blk {
        size_t off, size;
        bool free;
    };
    struct Chunk {
        char* base;
        size_t size, peak;
        std::vector<Blk> blks;
    };
    static constexpr size_t kChunk = static_cast<size_t>(48) << 20, kAlign = 1024;
    std::vector<Chunk> chunks;
    bool reuse = true;
    void* alloc(size_t bytes, std::vector<void*>& owner) {
        const size_t need = (std::max<size_t>(bytes, 1) + kAlign - 1) / kAlign * kAlign;
        for (;;) {
            for (auto& c : chunks)
                for (size_t i = 0; i < c.blks.size(); ++i) {
                    if (!c.blks[i].free || c.blks[i].size < need) continue;
                    const Blk b = c.blks[i];
                    c.blks[i] = Blk{b.off, need, false};
                    if (b.size > need) c.blks.insert(c.blks.begin() + i + 1, Blk{b.off + need, b.size - need, true});
                    c.peak = std::max(c.peak, b.off + need);
                    return c.base + b.off;
                }
            Chunk c{};
            c.size = std::max(kChunk, need);
            void* p = nullptr;
            CUDA_TRY(cudaMalloc(&p, c.size));
            owner.push_back(p);
            c.base = static_cast<char*>(p);
            c.blks.push_back(Blk{0, c.size, true});
            chunks.push_back(c);
        }
    }
    void release(void* p) {
        if (!reuse || !p) return;
        for (auto& c : chunks) {
            if (p < c.base || p >= c.base + c.size) continue;
            const size_t off = static_cast<size_t>(static_cast<char*>(p) - c.base);
            for (size_t i = 0; i < c.blks.size(); ++i) {
                if (c.blks[i].off != off) continue;
                if (c.blks[i].free) throw S3dError{"internal: activation released twice"};
                c.blks[i].free = true;
                if (i + 1 < c.blks.size() && c.blks[i + 1].free) {
                    c.blks[i].size += c.blks[i + 1].size;
                    c.blks.erase(c.blks.begin() + i + 1);
                }
                if (i > 0 && c.blks[i - 1].free) {
                    c.blks[i - 1].size += c.blks[i].size;
                    c.blks.erase(c.blks.begin() + i);
                }
                return;
            }
        }
        throw S3dError{"internal: release of an unknown activation"};
    }
    size_t touched() const {
        size_t n = 0;
        for (const auto& c : chunks) n += c.peak;
        return n;
    }
};

static const TensorSpec& T_(const s3d_unet* u, const std::string& n) {
    auto it = u->index.find(n);
    if (it == u->index.end()) throw S3dError{"internal: unknown tensor " + n};
    const TensorSpec& t = u->tensors[it->second];
    if (!t.loaded) throw S3dError{"checkpoint tensor not loaded: " + n};
    return t;
}

static uint16_t f2h_bits(float f) {
    __half h = __float2half_rn(f);   // host-callable
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
static float h2f(uint16_t b) {
    __half h;
    memcpy(&h, &b, 2);
    return __half2float(h);
}

// Pack one plane of a 3x3 conv (+ optional 1x1 skip) into [2][Cout][Ktot] fp16 (hi, lo*2048), K = tap*C + c.
static std::vector<uint16_t> pack_conv(const std::vector<float>& w, int Cout, int Cw, int C, const std::vector<float>* wskip,
                                       int Cs) {
    const int Ktot = 9 * C + Cs;
    std::vector<uint16_t> out(static_cast<size_t>(2) * Cout * Ktot);
    auto put = [&](int co, int k, float v) {
        float vc = std::min(std::max(v, -65504.f), 65504.f);
        uint16_t hi = f2h_bits(vc);
        uint16_t lo = f2h_bits((vc - h2f(hi)) * 2048.f);
        out[static_cast<size_t>(co) * Ktot + k] = hi;
        out[(static_cast<size_t>(Cout) + co) * Ktot + k] = lo;
    };
    for (int co = 0; co < Cout; ++co) {
        for (int tap = 0; tap < 9; ++tap)
            for (int c = 0; c < C; ++c) put(co, tap * C + c, w[(static_cast<size_t>(co) * Cw + c) * 9 + tap]);
        for (int c = 0; c < Cs; ++c) put(co, 9 * C + c, (*wskip)[static_cast<size_t>(co) * Cs + c]);
    }
    return out;
}
// Rollout 1-D weights of group g (1 or 2) of one plane, pre-summed per border class:
//   wc[(along*C + c)][(cls*Cout + co)] = sum_{across kept by cls} W[co][g*C + c][kh][kw]
// row_varying: along = kh, across = kw;  col_varying: along = kw, across = kh.
// cls: 0 interior keeps {0,1,2}, 1 first keeps {1,2}, 2 last keeps {0,1}, 3 single keeps {1}  (zero padding)
static std::vector<float> pack_roll(const std::vector<float>& w, int Cout, int C, int g, bool row_varying) {
    const int Cw = 3 * C, N = 4 * Cout;
    static const bool keep[4][3] = {{true, true, true}, {false, true, true}, {true, true, false}, {false, true, false}};
    std::vector<float> out(static_cast<size_t>(3) * C * N, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int c = 0; c < C; ++c)
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = w[((static_cast<size_t>(co) * Cw + g * C + c) * 3 + kh) * 3 + kw];
                    const int along = row_varying ? kh : kw, across = row_varying ? kw : kh;
                    for (int cls = 0; cls < 4; ++cls)
                        if (keep[cls][across]) out[(static_cast<size_t>(along) * C + c) * N + cls * Cout + co] += v;
                }
    return out;
}
// Which rollout group of which plane varies along rows (see kernels.cuh k_roll1d and unet_triplane.py:37-46):
//   xy: g1 = mean_D(yz)^T  -> varies with column (W);  g2 = mean_D(xz) -> varies with row (H)
//   xz: g1 = mean_W(xy)    -> row (H);                 g2 = mean_W(yz) -> column (D)
//   yz: g1 = mean_H(xy)^T  -> row (W);                 g2 = mean_H(xz) -> column (D)
static bool roll_row_varying(int plane, int g) { return plane == 0 ? g == 2 : g == 1; }

static void upload_conv3(s3d_unet* u, DevConv3& d, const std::string& name, int C, int Cout, const std::string& skip_name,
                         int Cs) {
    const bool ro = u->cfg.rollout;
    d.C = C;
    d.Cout = Cout;
    d.Cw = ro ? 3 * C : C;
    d.Cs = Cs;
    d.Ktot = 9 * C + Cs;
    for (int p = 0; p < 3; ++p) {
        const auto& w = T_(u, name + ".conv_" + kPlane[p] + ".weight").host;
        std::vector<float> bias = T_(u, name + ".conv_" + kPlane[p] + ".bias").host;
        const std::vector<float>* ws = nullptr;
        if (Cs) {
            ws = &T_(u, skip_name + ".conv_" + kPlane[p] + ".weight").host;
            const auto& bs = T_(u, skip_name + ".conv_" + kPlane[p] + ".bias").host;
            for (int i = 0; i < Cout; ++i) bias[i] += bs[i];
            d.wskip_orig[p] = dev_upload(u->wallocs, *ws);
        }
        d.w_orig[p] = dev_upload(u->wallocs, w);
        d.bias[p] = dev_upload(u->wallocs, bias);
        auto packed = pack_conv(w, Cout, d.Cw, C, ws, Cs);
        d.w_pack[p] = reinterpret_cast<__half*>(dev_upload(u->wallocs, packed));
        if (ro)
            for (int g = 1; g <= 2; ++g) {
                const std::vector<float> wc = pack_roll(w, Cout, C, g, roll_row_varying(p, g));
                d.wr[p][g - 1] = dev_upload(u->wallocs, wc);
                const int K = 3 * C, N = 4 * Cout;
                std::vector<uint16_t> w16(static_cast<size_t>(2) * N * K);
                for (int k = 0; k < K; ++k)
                    for (int n = 0; n < N; ++n) {
                        const float v = std::min(std::max(wc[static_cast<size_t>(k) * N + n], -65504.f), 65504.f);
                        const uint16_t hi = f2h_bits(v);
                        w16[static_cast<size_t>(n) * K + k] = hi;
                        w16[(static_cast<size_t>(N) + n) * K + k] = f2h_bits((v - h2f(hi)) * 2048.f);
                    }
                d.wr16[p][g - 1] = reinterpret_cast<__half*>(dev_upload(u->wallocs, w16));
            }
    }
}
static void upload_norm(s3d_unet* u, DevNorm& n, const std::string& name) {
    for (int p = 0; p < 3; ++p) {
        n.gamma[p] = dev_upload(u->wallocs, T_(u, name + ".norm_" + kPlane[p] + ".weight").host);
        n.beta[p] = dev_upload(u->wallocs, T_(u, name + ".norm_" + kPlane[p] + ".bias").host);
    }
}

static void free_all(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}
static void destroy_plan(s3d_unet* u) {
    if (!u->plan) return;
    for (auto& kv : u->plan->graphs) cudaGraphExecDestroy(kv.second);
    free_all(u->plan->allocs);
    u->plan.reset();
}

static void build_dgrad_packs(s3d_unet* u);      // train_host.cuh

static void finalize(s3d_unet* u) {
    CUDA_TRY(cudaSetDevice(u->device));
    destroy_plan(u);
    free_all(u->wallocs);
    u->pack_jobs_host.clear();           // the cached refresh table points into the buffers just freed
    const auto& c = u->cfg;
    u->te_w0 = dev_upload(u->wallocs, T_(u, "time_embed.0.weight").host);
    u->te_b0 = dev_upload(u->wallocs, T_(u, "time_embed.0.bias").host);
    u->te_w2 = dev_upload(u->wallocs, T_(u, "time_embed.2.weight").host);
    u->te_b2 = dev_upload(u->wallocs, T_(u, "time_embed.2.bias").host);
    // sinusoid frequencies (nn.py:113-116); the host mirror may override them through the "__freqs" pseudo tensor
    const int half = c.model_channels / 2;
    std::vector<float> fr(half);
    auto it = u->index.find("__freqs");
    if (it != u->index.end() && u->tensors[it->second].loaded) fr = u->tensors[it->second].host;
    else
        for (int i = 0; i < half; ++i) fr[i] = expf(static_cast<float>(-std::log(10000.0)) * static_cast<float>(i) / half);
    u->freqs = dev_upload(u->wallocs, fr);
    // concatenated emb_layers projection: film[n] = b[n] + W[n][:] . silu(emb)
    std::vector<float> fw(static_cast<size_t>(u->film_dim) * u->emb_dim), fb(u->film_dim);
    for (const auto& b : u->blocks) {
        const auto& w = T_(u, b.name + ".emb_layers.1.weight").host;
        const auto& bb = T_(u, b.name + ".emb_layers.1.bias").host;
        std::copy(w.begin(), w.end(), fw.begin() + static_cast<size_t>(b.film_off) * u->emb_dim);
        std::copy(bb.begin(), bb.end(), fb.begin() + b.film_off);
    }
    u->film_w = dev_upload(u->wallocs, fw);
    u->film_b = dev_upload(u->wallocs, fb);
    for (int p = 0; p < 3; ++p) {
        u->in_w[p] = dev_upload(u->wallocs, T_(u, std::string("in_conv.0.conv_") + kPlane[p] + ".weight").host);
        u->in_b[p] = dev_upload(u->wallocs, T_(u, std::string("in_conv.0.conv_") + kPlane[p] + ".bias").host);
        u->out_w[p] = dev_upload(u->wallocs, T_(u, std::string("out.2.conv_") + kPlane[p] + ".weight").host);
        u->out_b[p] = dev_upload(u->wallocs, T_(u, std::string("out.2.conv_") + kPlane[p] + ".bias").host);
    }
    upload_norm(u, u->out_norm, "out.0");
    u->dblocks.assign(u->blocks.size(), DevBlock{});      // (also forgets the dgrad packs freed with wallocs above)
    for (size_t i = 0; i < u->blocks.size(); ++i) {
        const auto& b = u->blocks[i];
        DevBlock& d = u->dblocks[i];
        upload_norm(u, d.n1, b.name + ".in_layers.0");
        upload_norm(u, d.n2, b.name + ".out_layers.0");
        upload_conv3(u, d.c1, b.name + ".in_layers.2", b.cin, b.cout, "", 0);
        upload_conv3(u, d.c2, b.name + ".out_layers.2", b.cout, b.cout, b.name + ".skip_connection", b.has_skip ? b.cin : 0);
    }
    CUDA_TRY(cudaFuncSetAttribute(k_conv_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<3>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<1>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_conv_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<2>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_conv_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<4>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_conv_tc<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<5>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_gn_silu, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_roll_bwd_vec, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_roll1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_roll_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<3>::kRollSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_roll_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<1>::kRollSmemBytes));
    CUDA_TRY(cudaDeviceSynchronize());
    if (u->training) build_dgrad_packs(u);
    u->finalized = true;
}

// Device-side refresh of every packed operand from checkpoint tensors that already live on the device (state_dict order): what a
// training step calls after the optimizer has changed the weights.  Same values as finalize() produces from host copies; needs
// the buffers finalize() allocated.  Enqueued on `s`; no host synchronisation.
static void refresh_from_device(s3d_unet* u, const float* const* src, int n_src, cudaStream_t s) {
    S3D_CHECK(u->finalized, "s3d_unet_refresh_dev needs one host-side load + s3d_unet_finalize first (it allocates the operand buffers)");
    std::map<std::string, const float*> by_name;
    int k = 0;
    for (const auto& t : u->tensors) {
        if (t.name == "__freqs") continue;
        S3D_CHECK(k < n_src && src[k] != nullptr, "s3d_unet_refresh_dev: one device pointer per checkpoint tensor, in state_dict order");
        by_name[t.name] = src[k++];
    }
    S3D_CHECK(k == n_src, "s3d_unet_refresh_dev: tensor count mismatch");
    auto numel = [&](const std::string& n) { return static_cast<long long>(u->tensors[u->index.at(n)].numel()); };
    // everything goes into ONE job table (k_pack_jobs): ~140 copies + ~170 re-packs would otherwise be ~300 launches per step
    std::vector<PackJob> jobs;
    auto job = [&](int kind, const float* src, const float* src2, void* dst, void* dst2, long long n, int Cout = 0, int Cw = 0, int C = 0,
                   int Cs = 0, int g = 0, int rowv = 0) {
        PackJob J{};
        J.kind = kind; J.src = src; J.src2 = src2; J.dst = dst; J.dst2 = dst2; J.n = n;
        J.Cout = Cout; J.Cw = Cw; J.C = C; J.Cs = Cs; J.g = g; J.rowv = rowv;
        jobs.push_back(J);
    };
    auto copy = [&](float* dst, const std::string& n) { job(0, by_name.at(n), nullptr, dst, nullptr, numel(n)); };
    copy(u->te_w0, "time_embed.0.weight");
    copy(u->te_b0, "time_embed.0.bias");
    copy(u->te_w2, "time_embed.2.weight");
    copy(u->te_b2, "time_embed.2.bias");
    for (const auto& b : u->blocks) {
        copy(u->film_w + static_cast<size_t>(b.film_off) * u->emb_dim, b.name + ".emb_layers.1.weight");
        copy(u->film_b + b.film_off, b.name + ".emb_layers.1.bias");
    }
    auto copy_norm = [&](DevNorm& nm, const std::string& name) {
        for (int p = 0; p < 3; ++p) {
            copy(nm.gamma[p], name + ".norm_" + kPlane[p] + ".weight");
            copy(nm.beta[p], name + ".norm_" + kPlane[p] + ".bias");
        }
    };
    for (int p = 0; p < 3; ++p) {
        copy(u->in_w[p], std::string("in_conv.0.conv_") + kPlane[p] + ".weight");
        copy(u->in_b[p], std::string("in_conv.0.conv_") + kPlane[p] + ".bias");
        copy(u->out_w[p], std::string("out.2.conv_") + kPlane[p] + ".weight");
        copy(u->out_b[p], std::string("out.2.conv_") + kPlane[p] + ".bias");
    }
    copy_norm(u->out_norm, "out.0");
    const bool ro = u->cfg.rollout;
    auto refresh_conv = [&](DevConv3& d, const std::string& name, const std::string& skip_name) {
        for (int p = 0; p < 3; ++p) {
            const float* w = by_name.at(name + ".conv_" + kPlane[p] + ".weight");
            const float* ws = nullptr;
            const float* sb = nullptr;
            copy(d.w_orig[p], name + ".conv_" + kPlane[p] + ".weight");
            if (d.Cs) {
                ws = by_name.at(skip_name + ".conv_" + kPlane[p] + ".weight");
                copy(d.wskip_orig[p], skip_name + ".conv_" + kPlane[p] + ".weight");
                sb = by_name.at(skip_name + ".conv_" + kPlane[p] + ".bias");
            }
            job(1, by_name.at(name + ".conv_" + kPlane[p] + ".bias"), sb, d.bias[p], nullptr, d.Cout);
            job(2, w, ws, d.w_pack[p], nullptr, static_cast<long long>(d.Cout) * d.Ktot, d.Cout, d.Cw, d.C, d.Cs);
            if (ro)
                for (int g = 1; g <= 2; ++g)
                    job(3, w, nullptr, d.wr[p][g - 1], d.wr16[p][g - 1], static_cast<long long>(3) * d.C * 4 * d.Cout, d.Cout, d.Cw, d.C, d.Cs, g,
                        roll_row_varying(p, g) ? 1 : 0);
            if (d.wd_pack[p]) job(4, w, nullptr, d.wd_pack[p], nullptr, static_cast<long long>(d.C) * 9 * d.Cout, d.Cout, d.Cw, d.C, d.Cs);
            for (int g = 1; g <= 2 && ro; ++g)
                if (d.wrv[p][g - 1])
                    job(6, w, nullptr, d.wrv[p][g - 1], nullptr, static_cast<long long>(9) * d.Cout * d.C, d.Cout, d.Cw, d.C, d.Cs, g,
                        roll_row_varying(p, g) ? 1 : 0);
            if (d.Cs && d.wsd_pack[p]) job(5, ws, nullptr, d.wsd_pack[p], nullptr, static_cast<long long>(d.Cs) * d.Cout, d.Cout, d.Cw, d.C, d.Cs);
        }
    };
    for (size_t i = 0; i < u->blocks.size(); ++i) {
        const auto& b = u->blocks[i];
        DevBlock& d = u->dblocks[i];
        copy_norm(d.n1, b.name + ".in_layers.0");
        copy_norm(d.n2, b.name + ".out_layers.0");
        refresh_conv(d.c1, b.name + ".in_layers.2", "");
        refresh_conv(d.c2, b.name + ".out_layers.2", b.name + ".skip_connection");
    }
    // the table only changes when the caller's pointers (or the set of training packs) do: upload it then, replay it otherwise
    const size_t bytes = jobs.size() * sizeof(PackJob);
    if (u->pack_jobs_host.size() != bytes || memcmp(u->pack_jobs_host.data(), jobs.data(), bytes) != 0) {
        if (u->pack_jobs_dev) cudaFree(u->pack_jobs_dev);
        CUDA_TRY(cudaMalloc(&u->pack_jobs_dev, bytes));
        CUDA_TRY(cudaMemcpy(u->pack_jobs_dev, jobs.data(), bytes, cudaMemcpyHostToDevice));
        u->pack_jobs_host.assign(reinterpret_cast<const char*>(jobs.data()), reinterpret_cast<const char*>(jobs.data()) + bytes);
    }
    launch_plain(k_pack_jobs, dim3(24, static_cast<unsigned>(jobs.size())), dim3(256), 0, s, static_cast<const PackJob*>(u->pack_jobs_dev));
    LAUNCH_CHECK("k_pack_jobs");
}

// ------------------------------------------------------------------------------------ launch plan
// GroupNorm statistics attached to a tensor by its producer kernel.  The buffers exist from the start; the consumer
// arms the box with its norm parameters (attach_norm) before the plan runs, and an un-armed box is skipped at launch.
struct SinkBox {
    StatsSink s{};       // producer view
    StatsSrc src{};      // consumer view (norm parameters filled in by the consumer)
    bool armed = false;
    bool use_film = false;
};
struct ActF {
    TriF p;
    int C = 0;
    int level = 0;
    std::shared_ptr<SinkBox> sink;   // null: producer cannot emit statistics -> stand-alone k_gn_stats
};
struct Act16 {
    TriH p;
    int C = 0;
};

struct PlanBuilder {
    s3d_unet* u;
    Plan* P;
    std::vector<TriDims> dims;   // per level
    int B;
    Arena arena;
    bool train = false;          // keep every activation, no in-step re-zeroing chains, record the tape for the backward
    Arena barena;                // backward temporaries (always with reuse)

    void add_op(const char* name, double flops, std::function<void(cudaStream_t)> fn, Trace tr = Trace{}) {
        P->ops.push_back(std::move(fn));
        P->op_names.push_back(name);
        P->op_flops.push_back(flops);
        P->op_trace.push_back(tr);
    }
    static constexpr int kTraceCtas = 2048;
    Trace new_trace() {
        Trace t{};
        if (!u->trace_on) return t;
        const size_t n = static_cast<size_t>(kTraceCtas) * 2 * kTraceSlots;
        t.buf = dev_alloc<unsigned long long>(P->allocs, n);
        CUDA_TRY(cudaMemset(t.buf, 0, n * sizeof(unsigned long long)));
        t.max_ctas = kTraceCtas;
        if (const char* e = getenv("S3D_TRACE_LT0")) t.lt0 = atoi(e);
        return t;
    }
    // dense FLOPs of one TriplaneConv 3x3 (+ its 1x1 skip) as the reference executes it: rollout channels counted
    double conv_flops(int level, const DevConv3& cv) const {
        double pxs = static_cast<double>(px(level, 0) + px(level, 1) + px(level, 2));
        return 2.0 * B * pxs * cv.Cout * (9.0 * cv.Cw + cv.Cs);
    }
    size_t px(int level, int plane) const { return static_cast<size_t>(dims[level].rows[plane]) * dims[level].cols[plane]; }
    int max_px(int level) const { return static_cast<int>(std::max({px(level, 0), px(level, 1), px(level, 2)})); }

    ActF allocF(int level, int C, const std::string& name) {
        ActF a;
        a.C = C;
        a.level = level;
        for (int p = 0; p < 3; ++p)
            a.p.p[p] = static_cast<float*>(arena.alloc(sizeof(float) * B * px(level, p) * C, P->allocs));
        if (!name.empty() && !arena.reuse) P->named.push_back({name, a.p, C, dims[level]});
        return a;
    }
    Act16 alloc16(int level, int C) {
        Act16 a;
        a.C = C;
        for (int p = 0; p < 3; ++p)
            a.p.p[p] = static_cast<__half*>(arena.alloc(sizeof(__half) * 2 * B * px(level, p) * C, P->allocs));
        return a;
    }
    void release(const ActF& a) {
        for (int p = 0; p < 3; ++p) arena.release(a.p.p[p]);
    }
    void release(const Act16& a) {
        for (int p = 0; p < 3; ++p) arena.release(a.p.p[p]);
    }
    static TriCF cf(const TriF& t) { return TriCF{{t.p[0], t.p[1], t.p[2]}}; }
    static TriCF cf3(float* const* t) { return TriCF{{t[0], t[1], t[2]}}; }

    // ---- rollout axis-sum accumulators / means of one GN+SiLU site (see k_gn_silu)
    struct Sums {
        unsigned long long* buf = nullptr;   // [B][total_len][C] 64-bit fixed point, zero between launches
        __half* means16 = nullptr;           // [2][B][total_len][C]
        unsigned int* ticket = nullptr;      // [B][total_tickets]
        int tick_off[3] = {};
        int total_tickets = 0;
        int seg_off[6] = {};                 // plane*2 + kind (0: indexed by row, 1: indexed by column)
        int total_len = 0;
    };
    Sums alloc_sums(int level, int C) {
        Sums S;
        const TriDims d = dims[level];
        int off = 0;
        for (int p = 0; p < 3; ++p) {
            S.seg_off[p * 2 + 0] = off;
            off += d.rows[p];
            S.seg_off[p * 2 + 1] = off;
            off += d.cols[p];
        }
        S.total_len = off;
        const size_t n = static_cast<size_t>(B) * off * C;
        S.buf = dev_alloc<unsigned long long>(P->allocs, n);
        CUDA_TRY(cudaMemset(S.buf, 0, sizeof(unsigned long long) * n));
        if (train) P->fwd_zero.push_back({S.buf, sizeof(unsigned long long) * n});
        S.means16 = dev_alloc<__half>(P->allocs, 2 * n);
        const int ny = std::max(1, 256 / (C / 4));
        int tk = 0;
        for (int p = 0; p < 3; ++p) {
            S.tick_off[p] = tk;
            tk += (d.rows[p] + kGsRows - 1) / kGsRows + (d.cols[p] + ny - 1) / ny;
        }
        S.total_tickets = tk;
        S.ticket = dev_alloc<unsigned int>(P->allocs, static_cast<size_t>(B) * tk);
        CUDA_TRY(cudaMemset(S.ticket, 0, sizeof(unsigned int) * B * tk));
        return S;
    }

    // ---- GroupNorm statistics (fixed-point group-sum accumulators; see StatsSink / StatsSrc in kernels.cuh)
    std::shared_ptr<SinkBox> make_box(int C) {
        auto bx = std::make_shared<SinkBox>();
        StatsSink& S = bx->s;
        const size_t n = static_cast<size_t>(B) * 3 * kGnRep * 64;
        S.acc = dev_alloc<unsigned long long>(P->allocs, n);
        CUDA_TRY(cudaMemset(S.acc, 0, sizeof(unsigned long long) * n));
        if (train) P->fwd_zero.push_back({S.acc, sizeof(unsigned long long) * n});
        S.C = C;
        bx->src.acc = S.acc;
        bx->src.film_dim = u->film_dim;
        bx->src.film_off = -1;
        return bx;
    }
    // Re-zeroing chain: the accumulators a consumer has read are cleared by the NEXT consumer kernel of the step (the first
    // consumer clears the last one's, which is a step old by then).
    std::shared_ptr<SinkBox> first_consumed, last_consumed;
    void chain_zero(const std::shared_ptr<SinkBox>& bx) {
        if (last_consumed) {
            bx->src.zero = last_consumed->s.acc;
            bx->src.zero_n = B * 3 * kGnRep * 64;
        }
        if (!first_consumed) first_consumed = bx;
        last_consumed = bx;
    }
    struct ZeroJob {
        unsigned long long* p = nullptr;
        long long n = 0;
    };
    ZeroJob last_sums{};
    std::shared_ptr<ZeroJob> first_zero_job;
    void close_zero_chain() {
        if (first_consumed && last_consumed) {
            first_consumed->src.zero = last_consumed->s.acc;
            first_consumed->src.zero_n = B * 3 * kGnRep * 64;
        }
        if (first_zero_job && last_sums.p) *first_zero_job = last_sums;
    }
    // sink as seen by a producer launch: disabled unless a consumer armed it
    static StatsSink live_sink(const std::shared_ptr<SinkBox>& bx) {
        StatsSink S{};
        if (bx && bx->armed) S = bx->s;
        return S;
    }
    // consumer view at launch time (binds the conditioning rows of this call)
    static StatsSrc live_src(const std::shared_ptr<SinkBox>& bx, const Plan* Pp) {
        StatsSrc S = bx->src;
        if (bx->use_film) {
            S.film = Pp->film;
            S.film_row = Pp->film_row;
        }
        return S;
    }
    // group sums of x for the norm layer `n` (+FiLM at film_off): emitted by x's producer when it supports it,
    // else by a stand-alone k_gn_stats pass
    std::shared_ptr<SinkBox> stats(const ActF& x, const DevNorm& n, int film_off) {
        std::shared_ptr<SinkBox> bx = x.sink;
        const bool standalone = !bx;
        const int level = x.level, C = x.C;
        S3D_CHECK(C % kGroups == 0 && C % 4 == 0 && C / 4 <= 128, "unsupported channel count for GroupNorm32");
        const int nslots = std::max(1, std::min(128, max_px(level) / 48));
        if (standalone) bx = make_box(C);
        S3D_CHECK(!bx->armed, "tensor normalised twice");
        if (!train) chain_zero(bx);        // training: the sums stay for the backward and are cleared before the next forward
        bx->src.gamma = cf3(n.gamma);
        bx->src.beta = cf3(n.beta);
        bx->src.film_off = film_off;
        bx->use_film = film_off >= 0;
        bx->armed = true;
        if (standalone) {
            TriCF xc = cf(x.p);
            TriDims d = dims[level];
            const int Bv = B;
            StatsSink S = bx->s;
            add_op("k_gn_stats", 0.0, [=](cudaStream_t s) {
                const int ny = std::max(1, std::min(16, 1024 / (C / 4)));
                dim3 grid(nslots, 3, Bv), block(C / 4, ny);
                launch(k_gn_stats, dim3(grid), dim3(block), sizeof(float) * (ny * 2 + 2) * C, s, xc, d, S, nslots);
                LAUNCH_CHECK("k_gn_stats");
            });
        }
        return bx;
    }

    // ---- GN apply + SiLU (+FiLM) -> fp16 operands (+ raw x16) (+ axis means)
    // x: fp32 input, or (xh != nullptr) the (hi, lo) fp16 input written by k_upcat
    void gn_silu(const ActF& x, const Act16* xh, const std::shared_ptr<SinkBox>& st, const Act16& a, const Act16* x16, const Sums* S) {
        const int level = x.level, C = x.C;
        const TriDims d = dims[level];
        const int bx = C / 4;
        S3D_CHECK(bx <= 256, "channel count too large for k_gn_silu");
        const int ny = std::max(1, 256 / bx);
        // column groups per CTA: the smallest tile width that keeps one sample's CTAs in one wave (2 CTAs / SM)
        int ncg = 1, gx = 0;
        for (;; ++ncg) {
            int active = 0;
            gx = 0;
            for (int p = 0; p < 3; ++p) {
                const int n = ((d.rows[p] + kGsRows - 1) / kGsRows) * ((d.cols[p] + ny * ncg - 1) / (ny * ncg));
                gx = std::max(gx, n);
                active += n;
            }
            if (active <= 2 * u->num_sms || ncg == 4) break;     // per sample: the geometry (hence every rounding) must not depend on B
        }
        GnSiluArgs A{};
        A.ncg = ncg;
        if (xh) A.xh = TriCH{{xh->p.p[0], xh->p.p[1], xh->p.p[2]}};
        else A.x = cf(x.p);
        A.d = d;
        A.C = C;
        A.a = a.p;
        A.finalize = (u->cfg.conv_impl == 1 || !u->fuse_roll) ? 1 : 0;
        if (x16) A.x16 = x16->p;
        if (S) {
            A.sums = S->buf;
            A.means16 = S->means16;
            A.ticket = S->ticket;
            A.total_tickets = S->total_tickets;
            for (int i = 0; i < 3; ++i) A.tick_off[i] = S->tick_off[i];
            for (int i = 0; i < 6; ++i) A.seg_off[i] = S->seg_off[i];
            A.total_len = S->total_len;
        }
        const size_t smem = sizeof(float) * (2 + static_cast<size_t>(ny) * kGsRows) * C;
        S3D_CHECK(smem <= 100 * 1024, "k_gn_silu shared memory");
        A.tr = new_trace();
        // Axis sums that a conv's roll tiles read raw are cleared by the NEXT k_gn_silu of the step (all its CTAs share the
        // work); the first one clears the last one's, which is a step old by then.  (With A.finalize the kernel converts and
        // clears its own sums.)
        auto zj = std::make_shared<ZeroJob>();
        if (S && !A.finalize && !train) {
            if (last_sums.p) *zj = last_sums;
            if (!first_zero_job) first_zero_job = zj;
            last_sums = ZeroJob{S->buf, static_cast<long long>(B) * S->total_len * C};
        }
        const int Bv = B;
        Plan* Pp = P;
        add_op("k_gn_silu", 0.0, [=](cudaStream_t s) {
            GnSiluArgs Al = A;
            Al.st = live_src(st, Pp);
            Al.zero_sums = zj->p;
            Al.zero_sums_n = zj->n;
            dim3 grid(gx, 3, Bv), block(bx, ny);
            launch(k_gn_silu, dim3(grid), dim3(block), smem, s, Al, Bv);
            LAUNCH_CHECK("k_gn_silu");
        }, A.tr);
    }

    struct TBuf {
        TriF Trow, Tcol;
        std::shared_ptr<RollTcMaps> roll_maps;   // set when the 1-D GEMM tiles are to be fused into the conv launch
        RollTcArgs roll_args{};
        int roll_mtiles = 0, roll_ntn = 0;
        Sums sums{};
    };
    // ---- rollout 1-D terms (tensor-core GEMM; SIMT cross-check kernel when conv_impl == 1)
    TBuf roll1d(const Sums& S, int level, const DevConv3& cv) {
        const TriDims d = dims[level];
        const int C = cv.C, Cout = cv.Cout;
        TBuf T{};
        for (int p = 0; p < 3; ++p) {
            T.Trow.p[p] = dev_alloc<float>(P->allocs, static_cast<size_t>(B) * 4 * d.rows[p] * Cout);
            T.Tcol.p[p] = dev_alloc<float>(P->allocs, static_cast<size_t>(B) * 4 * d.cols[p] * Cout);
        }
        // (plane, group) -> source plane / which of its means; see roll_row_varying() and unet_triplane.py:37-46
        //   kind 0 = the source plane's mean over its columns (indexed by its row), kind 1 = mean over rows (by column)
        struct SrcDef {
            int sp, kind;
        };
        const SrcDef def[3][2] = {{{2, 0}, {1, 0}}, {{0, 0}, {2, 1}}, {{0, 1}, {1, 1}}};
        int L[6], ncls[6], soff[6], Lmax = 0, ncls_max = 3;
        float* Tp[6];
        for (int p = 0; p < 3; ++p)
            for (int g = 0; g < 2; ++g) {
                const SrcDef sd = def[p][g];
                const bool rowv = roll_row_varying(p, g + 1);
                const int i = p * 2 + g;
                L[i] = rowv ? d.rows[p] : d.cols[p];
                const int across = rowv ? d.cols[p] : d.rows[p];
                ncls[i] = across == 1 ? 4 : 3;
                ncls_max = std::max(ncls_max, ncls[i]);
                soff[i] = S.seg_off[sd.sp * 2 + sd.kind];
                const int src_len = sd.kind == 0 ? d.rows[sd.sp] : d.cols[sd.sp];
                S3D_CHECK(src_len == L[i], "rollout geometry");
                Tp[i] = rowv ? T.Trow.p[p] : T.Tcol.p[p];
                Lmax = std::max(Lmax, L[i]);
            }
        const int Bv = B;
        const int ntn = ncls_max * Cout / 64;
        S3D_CHECK(C % 64 == 0 && Cout % 64 == 0, "rollout tiling");
        if (u->cfg.conv_impl == 1) {
            Roll1dArgs A{};
            A.C = C;
            A.Cout = Cout;
            A.means16 = S.means16;
            A.total_len = S.total_len;
            A.B = B;
            A.ntn = ntn;
            for (int i = 0; i < 6; ++i) {
                A.s[i].sum_off = soff[i];
                A.s[i].L = L[i];
                A.s[i].ncls = ncls[i];
                A.s[i].wc = cv.wr[i / 2][i % 2];
                A.s[i].T = Tp[i];
            }
            const size_t smem = sizeof(float) * static_cast<size_t>(18) * (C + 4);
            add_op("k_roll1d", 0.0, [=](cudaStream_t s) {
                dim3 grid((Lmax + 15) / 16, 6 * ntn, Bv);
                launch(k_roll1d, dim3(grid), dim3(128), smem, s, A);
                LAUNCH_CHECK("k_roll1d");
            });
            return T;
        }
        auto maps = std::make_shared<RollTcMaps>();
        memset(maps.get(), 0, sizeof(RollTcMaps));
        RollTcArgs A{};
        A.C = C;
        A.Cout = Cout;
        int total = 0;
        const uint64_t seg_bytes = static_cast<uint64_t>(S.total_len) * C * 2;
        for (int i = 0; i < 6; ++i) {
            A.L[i] = L[i];
            A.ncls[i] = ncls[i];
            A.T[i] = Tp[i];
            A.soff[i] = soff[i];
            {
                const SrcDef sd = def[i / 2][i % 2];
                const int avg_len = sd.kind == 0 ? d.cols[sd.sp] : d.rows[sd.sp];     // length of the axis the mean runs over
                A.scale[i] = static_cast<float>(1.0 / 16777216.0 / static_cast<double>(avg_len));
            }
            A.tile_start[i] = total;
            {
                const int tmax = std::max(8, std::min(kRollTmMax, u->roll_tm));
                const int nt = (L[i] + tmax - 1) / tmax;                  // equal tiles of at most `tmax` positions
                A.tm[i] = (L[i] + nt - 1) / nt;
                total += nt;
            }
            const uint64_t adims[5] = {static_cast<uint64_t>(C), static_cast<uint64_t>(L[i]), 1, static_cast<uint64_t>(B), 2};
            const uint64_t astr[4] = {static_cast<uint64_t>(C) * 2, seg_bytes, seg_bytes, seg_bytes * B};
            const uint32_t abox[5] = {kBK, kBM, 1, 1, 1};
            make_tmap_strided(&maps->a[i], S.means16 + static_cast<size_t>(soff[i]) * C, 5, adims, astr, abox);
            const uint64_t wdims[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(4 * Cout), 2};
            const bool blo = u->fuse_roll ? mode_has_blo(u->cfg.precision) : u->cfg.precision != 1;
            const uint32_t wbox[3] = {kBK, kBN, blo ? 2u : 1u};      // hi and lo tile in ONE box (lo lands behind hi)
            make_tmap(&maps->w[i], cv.wr16[i / 2][i % 2], 3, wdims, wbox);
        }
        A.tile_start[6] = total;
        const int nsplit = u->cfg.precision == 1 ? 1 : 3;
        if (u->fuse_roll) {
            T.roll_maps = maps;
            T.roll_args = A;
            T.roll_mtiles = total;
            T.roll_ntn = ntn;
            T.sums = S;
            return T;
        }
        add_op("k_roll_tc", 0.0, [=](cudaStream_t s) {
            dim3 grid(total, ntn, Bv);
            if (nsplit == 3) launch(k_roll_tc<3>, dim3(grid), dim3(kRollThreads), ConvTcCfg<3>::kRollSmemBytes, s, *maps, A);
            else launch(k_roll_tc<1>, dim3(grid), dim3(kRollThreads), ConvTcCfg<1>::kRollSmemBytes, s, *maps, A);
            LAUNCH_CHECK("k_roll_tc");
        });
        return T;
    }

    // ---- 3x3 conv (+ fused 1x1 skip)
    // pool_out != nullptr (tcgen05 path only): the epilogue also writes the 2x2-averaged output and its GroupNorm sums
    void conv(const Act16& a, int level, const DevConv3& cv, const TBuf* T, const Act16* x16, const ActF* resid, int emb_off,
              ActF& out, ActF* pool_out = nullptr) {
        const TriDims d = dims[level];
        ConvEpi e{};
        e.bias = cf3(cv.bias);
        if (T) {
            e.Trow = cf(T->Trow);
            e.Tcol = cf(T->Tcol);
        }
        if (resid) e.resid = cf(resid->p);
        e.film_dim = u->film_dim;
        e.film_off = emb_off;
        e.out = out.p;
        const bool use_emb = emb_off >= 0;
        Plan* Pp = P;
        const int Bv = B;
        if (u->cfg.conv_impl == 1) {
            ConvFfmaArgs A{};
            A.a = TriCH{{a.p.p[0], a.p.p[1], a.p.p[2]}};
            if (x16) A.x16 = TriCH{{x16->p.p[0], x16->p.p[1], x16->p.p[2]}};
            A.d = d;
            A.C = cv.C;
            A.Cout = cv.Cout;
            A.Cw = cv.Cw;
            A.Cs = cv.Cs;
            A.w = cf3(cv.w_orig);
            A.wskip = cf3(cv.wskip_orig);
            A.e = e;
            A.single = u->cfg.precision == 1;
            const int mp = max_px(level);
            add_op("k_conv_ffma", conv_flops(level, cv), [=](cudaStream_t s) {
                ConvFfmaArgs Al = A;
                if (use_emb) {
                    Al.e.embadd = Pp->film;
                    Al.e.film_row = Pp->film_row;
                }
                dim3 grid((mp + 3) / 4, 3, Bv), block(64, 4);
                launch(k_conv_ffma, dim3(grid), dim3(block), 0, s, Al, Bv);
                LAUNCH_CHECK("k_conv_ffma");
            });
            return;
        }
        S3D_CHECK(cv.C % kBK == 0 && cv.Cout % kBN == 0 && cv.Cs % kBK == 0 && cv.C + cv.Cs > 0, "tcgen05 conv needs channel counts % 64 == 0");
        auto maps = std::make_shared<ConvTcMaps>();
        memset(maps.get(), 0, sizeof(ConvTcMaps));
        ConvTcArgs A{};
        A.d = d;
        A.C = cv.C;
        A.Cout = cv.Cout;
        A.Cs = cv.Cs;
        A.e = e;
        A.bo_kw = u->halo_bo_kw ? 1 : 0;
        A.tr = new_trace();
        int total = 0;
        for (int p = 0; p < 3; ++p) {
            const uint64_t adims[5] = {static_cast<uint64_t>(cv.C), static_cast<uint64_t>(d.cols[p]),
                                       static_cast<uint64_t>(d.rows[p]), static_cast<uint64_t>(B), 2};
            const uint32_t abox[5] = {kBK, kHaloW, kHaloH, 1, 1};      // halo patch (cols w0-1.., rows h0-1..)
            const uint32_t xbox[5] = {kBK, kTileW, kTileH, 1, 1};      // plain tile for the 1x1 skip chunks
            if (cv.C) make_tmap(&maps->a[p], a.p.p[p], 5, adims, abox);
            if (cv.Cs) {
                const uint64_t xdims[5] = {static_cast<uint64_t>(cv.Cs), static_cast<uint64_t>(d.cols[p]),
                                           static_cast<uint64_t>(d.rows[p]), static_cast<uint64_t>(B), 2};
                make_tmap(&maps->x[p], x16->p.p[p], 5, xdims, xbox);
                if (!cv.C) maps->a[p] = maps->x[p];       // pure 1x1 GEMM (training: the skip conv's dgrad): no 3x3 part
            } else {
                maps->x[p] = maps->a[p];
            }
            const uint64_t wdims[3] = {static_cast<uint64_t>(cv.Ktot), static_cast<uint64_t>(cv.Cout), 2};
            const uint32_t wbox[3] = {kBK, kBN, mode_has_blo(u->cfg.precision) ? 2u : 1u};       // hi and lo tile in ONE box
            make_tmap(&maps->w[p], cv.w_pack[p], 3, wdims, wbox);
            A.tiles_x[p] = (d.cols[p] + kTileW - 1) / kTileW;
            const int tiles_y = (d.rows[p] + kTileH - 1) / kTileH;
            A.tile_start[p] = total;
            total += A.tiles_x[p] * tiles_y;
        }
        A.tile_start[3] = total;
        const int mode = u->cfg.precision;
        const int ntile_n = cv.Cout / kBN;
        const int num_sms = u->num_sms;
        // the epilogue emits the output's GroupNorm group sums when a group is 2, 4 or 8 channels wide (Cout = 64, 128, 256)
        std::shared_ptr<SinkBox> box;
        const int cpg_out = cv.Cout / kGroups;
        if (cv.Cout % kGroups == 0 && (cpg_out == 2 || cpg_out == 4 || cpg_out == 8)) {
            box = make_box(cv.Cout);
            out.sink = box;
        }
        std::shared_ptr<SinkBox> pbox;
        if (pool_out) {
            S3D_CHECK(box != nullptr, "fused pooling needs the statistics epilogue");
            for (int p = 0; p < 3; ++p)
                S3D_CHECK(dims[level + 1].rows[p] == d.rows[p] / 2 && dims[level + 1].cols[p] == d.cols[p] / 2, "pooled plane size");
            A.pool = pool_out->p;
            pbox = make_box(cv.Cout);
            pool_out->sink = pbox;
        }
        FusedRoll F{};
        auto rmaps = std::make_shared<RollTcMaps>();
        memset(rmaps.get(), 0, sizeof(RollTcMaps));
        if (T && T->roll_maps) {
            rmaps = T->roll_maps;
            F.R = T->roll_args;
            F.ntn = T->roll_ntn;
            F.n_roll = T->roll_mtiles * T->roll_ntn * B;
            F.counters = dev_alloc<unsigned int>(P->allocs, 2);
            CUDA_TRY(cudaMemset(F.counters, 0, 2 * sizeof(unsigned int)));
            // the roll tiles read the raw axis sums k_gn_silu accumulated (re-zeroed by the next k_gn_silu of the step)
            F.sums = T->sums.buf;
            F.total_len = T->sums.total_len;
        }
        add_op("k_conv_tc", conv_flops(level, cv), [=](cudaStream_t s) {
            ConvTcArgs Al = A;
            Al.sink = live_sink(box);
            if (pbox) Al.pool_sink = live_sink(pbox);
            if (use_emb) {
                Al.e.embadd = Pp->film;
                Al.e.film_row = Pp->film_row;
            }
            const int total_tiles = total * ntile_n * Bv + F.n_roll;
            dim3 grid(std::min(total_tiles, num_sms));
            if (mode == 3) launch(k_conv_tc<3>, dim3(grid), dim3(kConvThreads), ConvTcCfg<3>::kSmemBytes, s, *maps, *rmaps, Al, F, total_tiles);
            else if (mode == 2) launch(k_conv_tc<2>, dim3(grid), dim3(kConvThreads), ConvTcCfg<2>::kSmemBytes, s, *maps, *rmaps, Al, F, total_tiles);
            else if (mode == 4) launch(k_conv_tc<4>, dim3(grid), dim3(kConvThreads), ConvTcCfg<4>::kSmemBytes, s, *maps, *rmaps, Al, F, total_tiles);
            else if (mode == 5) launch(k_conv_tc<5>, dim3(grid), dim3(kConvThreads), ConvTcCfg<5>::kSmemBytes, s, *maps, *rmaps, Al, F, total_tiles);
            else launch(k_conv_tc<1>, dim3(grid), dim3(kConvThreads), ConvTcCfg<1>::kSmemBytes, s, *maps, *rmaps, Al, F, total_tiles);
            LAUNCH_CHECK("k_conv_tc");
        }, A.tr);
    }

    // One TriplaneResBlock (unet_triplane.py:269-311).  The input is either the fp32 residual stream `x`, or — for the decoder
    // blocks behind a concat — the (hi, lo) fp16 tensor `xh` that k_upcat wrote (it doubles as the skip GEMM's operand).
    // Buffers go back to the arena as soon as their last reader is planned; the caller releases the block's input.
    struct BlockTape {               // what one TriplaneResBlock's backward reads (training plans)
        int bi = -1, level = 0;
        bool x_is_pair = false;
        ActF x;                      // block input (fp32), or
        Act16 xh;                    // the (hi, lo) concat input
        std::shared_ptr<SinkBox> st1, st2;
        Act16 a1, a2, x16;
        Sums s1, s2;
        ActF h1, out;
    };
    struct UpTape {
        ActF low, skip;
        int out_level = 0;
        bool do_up = false;
    };
    std::vector<BlockTape> enc_tape, dec_tape;      // in forward order
    std::vector<UpTape> up_tape;                    // up_tape[j] feeds dec_tape[j] (j >= 1)
    std::vector<BlockTape>* cur_tape = nullptr;

    ActF res_block(int bi, const ActF* x, const Act16* xh, const std::shared_ptr<SinkBox>& xh_sink, int level, ActF* pool_out = nullptr) {
        const BlockSpec& b = u->blocks[bi];
        const DevBlock& w = u->dblocks[bi];
        const bool ro = u->cfg.rollout, ssn = u->cfg.use_scale_shift_norm;
        S3D_CHECK((x ? x->C : xh->C) == b.cin, "res block input width");
        S3D_CHECK(x || b.has_skip, "a (hi, lo) input needs the 1x1 skip GEMM");
        Sums s1{}, s2{};
        if (ro) {
            s1 = alloc_sums(level, b.cin);
            s2 = alloc_sums(level, b.cout);
        }
        ActF xin{};
        if (x) xin = *x;
        else {
            xin.C = xh->C;
            xin.level = level;
            xin.sink = xh_sink;
        }
        auto st1 = stats(xin, w.n1, -1);
        Act16 a1 = alloc16(level, b.cin);
        Act16 x16{};
        if (b.has_skip) x16 = x ? alloc16(level, b.cin) : *xh;
        gn_silu(xin, x ? nullptr : xh, st1, a1, (b.has_skip && x) ? &x16 : nullptr, ro ? &s1 : nullptr);
        TBuf t1{};
        if (ro) t1 = roll1d(s1, level, w.c1);
        ActF h1 = allocF(level, b.cout, b.name + ".h1");
        conv(a1, level, w.c1, ro ? &t1 : nullptr, nullptr, nullptr, ssn ? -1 : b.film_off, h1);
        release(a1);
        auto st2 = stats(h1, w.n2, ssn ? b.film_off : -1);
        Act16 a2 = alloc16(level, b.cout);
        gn_silu(h1, nullptr, st2, a2, nullptr, ro ? &s2 : nullptr);
        release(h1);
        TBuf t2{};
        if (ro) t2 = roll1d(s2, level, w.c2);
        ActF out = allocF(level, b.cout, b.name + ".out");
        conv(a2, level, w.c2, ro ? &t2 : nullptr, b.has_skip ? &x16 : nullptr, b.has_skip ? nullptr : x, -1, out, pool_out);
        release(a2);
        if (b.has_skip && x) release(x16);
        if (train && cur_tape) {
            BlockTape T;
            T.bi = bi;
            T.level = level;
            T.x_is_pair = x == nullptr;
            if (x) T.x = *x;
            else T.xh = *xh;
            T.st1 = st1;
            T.st2 = st2;
            T.a1 = a1;
            T.a2 = a2;
            T.x16 = x16;
            T.s1 = s1;
            T.s2 = s2;
            T.h1 = h1;
            T.out = out;
            cur_tape->push_back(T);
        }
        return out;
    }

    ActF down(const ActF& x, int li) {
        ActF o = allocF(x.level + 1, x.C, "down." + std::to_string(li));
        const TriDims di = dims[x.level], dd = dims[x.level + 1];
        const int C = x.C, Bv = B;
        TriCF xc = cf(x.p);
        TriF op = o.p;
        const int nslots = std::max(1, std::min(128, max_px(x.level + 1) / 16));
        auto box = make_box(C);
        o.sink = box;
        S3D_CHECK(C / 4 <= 256, "channel count too large for k_avgpool2");
        const Trace tr = new_trace();
        add_op("k_avgpool2", 0.0, [=](cudaStream_t s) {
            const int ny = std::max(1, 256 / (C / 4));
            dim3 grid(nslots, 3, Bv), block(C / 4, ny);
            launch(k_avgpool2, dim3(grid), dim3(block), sizeof(float) * (ny * 2 + 2) * C, s, xc, di, dd, C, op, live_sink(box), nslots, tr);
            LAUNCH_CHECK("k_avgpool2");
        }, tr);
        return o;
    }

    // x2 bilinear of `low` (+ resize to the skip's size) and concat with `skip`; the result is written once, as the (hi, lo)
    // fp16 pair that both of its readers want (the 1x1 skip GEMM directly, GroupNorm + SiLU after re-joining the halves)
    Act16 upcat(const ActF& low, const ActF& skip, int out_level, bool do_up, std::shared_ptr<SinkBox>& sink_out) {
        const int Cs = skip.C;
        Act16 o = alloc16(out_level, low.C + Cs);
        const TriDims dl = dims[low.level], dout = dims[out_level];
        TriCF lc = cf(low.p), sc = cf(skip.p);
        TriH op = o.p;
        const int Cu = low.C, Bv = B, Ct = Cu + Cs;
        S3D_CHECK(Ct / 4 <= 256 && Cu % 4 == 0 && Cs % 4 == 0, "channel counts unsupported by k_upcat");
        // one wave at 2 CTAs / SM (the gather loop is latency bound: fewer, longer CTAs with 4 pixels in flight per thread)
        const int nslots = std::max(1, std::min(2 * u->num_sms / 3, max_px(out_level) / 32));
        auto box = make_box(Ct);
        sink_out = box;
        const Trace tr = new_trace();
        add_op("k_upcat", 0.0, [=](cudaStream_t s) {
            const int ny = std::max(1, 256 / (Ct / 4));
            dim3 grid(nslots, 3, Bv), block(Ct / 4, ny);
            launch(k_upcat, dim3(grid), dim3(block), sizeof(float) * (ny * 2 + 2) * Ct, s, lc, dl, Cu, sc, Cs, dout, op, Bv, do_up ? 1 : 0,
                   live_sink(box), nslots, tr);
            LAUNCH_CHECK("k_upcat");
        }, tr);
        return o;
    }
};

#include "train_host.cuh"

static void build_plan(s3d_unet* u, int B, int H, int W, int D) {
    CUDA_TRY(cudaSetDevice(u->device));
    destroy_plan(u);
    u->plan.reset(new Plan());
    Plan* P = u->plan.get();
    P->B = B; P->H = H; P->W = W; P->D = D;
    struct CounterGuard {
        CounterGuard(size_t* c) { g_alloc_counter = c; }
        ~CounterGuard() { g_alloc_counter = nullptr; }
    } guard(&P->alloc_bytes);
    const auto& c = u->cfg;
    PlanBuilder pb{u, P, {}, B};
    if (const char* e = getenv("S3D_KEEP_ACTS")) pb.arena.reuse = atoi(e) == 0;
    if (u->training) {
        S3D_CHECK(u->cfg.conv_impl == 0 && u->fuse_roll, "training needs the tcgen05 conv path with fused roll tiles");
        pb.train = P->train = true;
        pb.arena.reuse = false;        // every activation is read again by the backward
    }
    TriDims d{{H, H, W}, {W, D, D}};
    for (int l = 0; l < c.n_levels; ++l) {
        S3D_CHECK(d.rows[0] >= 1 && d.rows[2] >= 1 && d.cols[1] >= 1, "triplane too small for the number of levels");
        pb.dims.push_back(d);
        for (int p = 0; p < 3; ++p) {
            d.rows[p] /= 2;
            d.cols[p] /= 2;
        }
    }
    P->film_own = dev_alloc<float>(P->allocs, static_cast<size_t>(B) * u->film_dim);
    // ---- in_conv
    const int c0 = ch_of(c, 0);
    ActF h = pb.allocF(0, c0, "in_conv");
    S3D_CHECK(c0 == 64 || c0 == 128, "channel_mult[0] * model_channels must be 64 or 128 (boundary kernels: one lane per channel quad of a pixel)");
    S3D_CHECK(c.in_channels <= kMaxCf && c.out_channels <= kMaxCf, "at most 16 triplane channels are supported");
    BoundaryArgs bnd{};      // fields shared by the three modes
    bnd.d = pb.dims[0];
    bnd.C0 = c0;
    bnd.H = H; bnd.W = W; bnd.Dd = D;
    {
        // 4 x 32-pixel tiles, the 32 along the axis that is contiguous in the composed tensor (rows for the transposed yz plane)
        int total = 0;
        for (int p = 0; p < 4; ++p) {
            const int rows = p < 3 ? bnd.d.rows[p] : D, cols = p < 3 ? bnd.d.cols[p] : D;
            const int fast = p == 2 ? rows : cols, slow = p == 2 ? cols : rows;
            bnd.tiles_fast[p] = (fast + kBndFast - 1) / kBndFast;
            const int n = bnd.tiles_fast[p] * ((slow + kBndSlow - 1) / kBndSlow);
            bnd.tile_start[p] = total;
            total += n;
        }
        bnd.tile_start[4] = total;
    }
    bnd.w_in = PlanBuilder::cf3(u->in_w);
    bnd.b_in = PlanBuilder::cf3(u->in_b);
    bnd.h0 = h.p;
    auto in_box = pb.make_box(c0);
    h.sink = in_box;
    P->in_acc = in_box->s.acc;
    P->in_acc_n = static_cast<size_t>(B) * 3 * kGnRep * 64;
    {
        const int Cin = c.in_channels;
        const Trace tr = pb.new_trace();
        pb.add_op("k_boundary<in_conv>", 2.0 * B * (pb.px(0, 0) + pb.px(0, 1) + pb.px(0, 2)) * Cin * c0, [=](cudaStream_t s) {
            BoundaryArgs Al = bnd;
            Al.Cf = Cin;
            Al.tr = tr;
            Al.x_in = P->x;
            Al.sink = PlanBuilder::live_sink(in_box);
            const dim3 grid(bnd.tile_start[3], B);
            if (c0 == 64) launch(k_boundary<MODE_INCONV, 16>, dim3(grid), dim3(256), 0, s, Al);
            else launch(k_boundary<MODE_INCONV, 32>, dim3(grid), dim3(256), 0, s, Al);
            LAUNCH_CHECK("k_boundary<in_conv>");
        }, tr);
    }
    // ---- encoder
    pb.cur_tape = &pb.enc_tape;
    std::vector<ActF> stack;
    ActF pooled{};
    bool have_pooled = false;     // the Downsample2x of the next level was produced by the previous block's last conv
    for (int l = 0; l < c.n_levels; ++l) {
        const auto& ops = u->downs[l];
        for (size_t oi = 0; oi < ops.size(); ++oi) {
            const auto& op = ops[oi];
            if (op.kind == 1) {
                if (have_pooled) {
                    h = pooled;               // (the block output stays alive: it is the previous level's skip)
                    have_pooled = false;
                } else {
                    h = pb.down(h, l);        // its input stays alive: it is the previous level's skip
                }
            } else {
                // last block of the level, next level opens with a Downsample2x: fuse the pooling into this block's last conv
                const bool fuse = u->cfg.conv_impl == 0 && u->fuse_pool && oi + 1 == ops.size() && l + 1 < c.n_levels &&
                                  !u->downs[l + 1].empty() && u->downs[l + 1][0].kind == 1 &&
                                  (u->blocks[op.block].cout / kGroups == 2 || u->blocks[op.block].cout / kGroups == 4 ||
                                   u->blocks[op.block].cout / kGroups == 8) && u->blocks[op.block].cout % 64 == 0;
                if (fuse) {
                    pooled = pb.allocF(h.level + 1, u->blocks[op.block].cout, "down." + std::to_string(l + 1));
                    have_pooled = true;
                }
                ActF o = pb.res_block(op.block, &h, nullptr, nullptr, h.level, fuse ? &pooled : nullptr);
                pb.release(h);
                h = o;
            }
        }
        stack.push_back(h);
    }
    // ---- decoder (unet_triplane.py:488-505)
    pb.cur_tape = &pb.dec_tape;
    pb.up_tape.assign(c.n_levels, PlanBuilder::UpTape{});
    bool pending_up = false;    // an Upsample2x closed the previous output block
    for (int j = 0; j < c.n_levels; ++j) {
        const int level = c.n_levels - 1 - j;
        Act16 hcat{};
        std::shared_ptr<SinkBox> hcat_sink;
        if (j == 0) {
            h = stack.back();
            stack.pop_back();
        } else {
            ActF skip = stack.back();
            stack.pop_back();
            if (!c.rollout) {
                // ...SmallRaw concatenates without resizing; mismatched (odd) sizes fail in the reference too
                for (int p = 0; p < 3; ++p)
                    S3D_CHECK(2 * pb.dims[level + 1].rows[p] == pb.dims[level].rows[p] &&
                                  2 * pb.dims[level + 1].cols[p] == pb.dims[level].cols[p],
                              "TriplaneUNetModelSmallRaw needs even plane sizes at every level (torch.cat would fail)");
            }
            hcat = pb.upcat(h, skip, level, pending_up, hcat_sink);
            pb.up_tape[j] = PlanBuilder::UpTape{h, skip, level, pending_up};
            pb.release(h);
            pb.release(skip);
            pending_up = false;
        }
        for (const auto& op : u->ups[j]) {
            if (op.kind == 2) {
                pending_up = true;     // fused into the next level's upcat
            } else if (j == 0) {
                ActF o = pb.res_block(op.block, &h, nullptr, nullptr, level);
                pb.release(h);
                h = o;
            } else {
                h = pb.res_block(op.block, nullptr, &hcat, hcat_sink, level);
                pb.release(hcat);
            }
        }
    }
    S3D_CHECK(!pending_up && h.level == 0, "decoder structure");
    // ---- out head (stand-alone forward) and the fused step boundary (sampling loop)
    auto st = pb.stats(h, u->out_norm, -1);
    {
        const int Cout = c.out_channels;
        S3D_CHECK(h.C == c0, "decoder output width");
        BoundaryArgs bo = bnd;
        bo.h = PlanBuilder::cf(h.p);
        bo.w_out = PlanBuilder::cf3(u->out_w);
        bo.b_out = PlanBuilder::cf3(u->out_b);
        const Trace tr = pb.new_trace();
        pb.add_op("k_boundary<head>", 2.0 * B * (pb.px(0, 0) + pb.px(0, 1) + pb.px(0, 2)) * c0 * Cout, [=](cudaStream_t s) {
            BoundaryArgs Al = bo;
            Al.Cf = Cout;
            Al.tr = tr;
            Al.st = PlanBuilder::live_src(st, P);
            Al.model_out = P->out;
            const dim3 grid(bo.tile_start[4], B);
            if (c0 == 64) launch(k_boundary<MODE_HEAD, 16>, dim3(grid), dim3(256), 0, s, Al);
            else launch(k_boundary<MODE_HEAD, 32>, dim3(grid), dim3(256), 0, s, Al);
            LAUNCH_CHECK("k_boundary<head>");
        }, tr);
        if (c.in_channels == c.out_channels) {
            const Trace trf = pb.new_trace();
            P->fused_trace = trf;
            P->fused_boundary = [=](cudaStream_t s, const SchedArgs& sch) {
                BoundaryArgs Al = bo;
                Al.Cf = Cout;
                Al.tr = trf;
                Al.st = PlanBuilder::live_src(st, P);
                Al.sink = PlanBuilder::live_sink(in_box);
                Al.sch = sch;
                const dim3 grid(bo.tile_start[4], B);
                if (c0 == 64) launch(k_boundary<MODE_FUSED, 16>, dim3(grid), dim3(256), 0, s, Al);
                else launch(k_boundary<MODE_FUSED, 32>, dim3(grid), dim3(256), 0, s, Al);
                LAUNCH_CHECK("k_boundary<fused>");
            };
        }
    }
    pb.close_zero_chain();
    if (pb.train) build_backward(u, pb, h, st, bnd);
    P->alloc_bytes += pb.arena.touched() + pb.barena.touched();
    P->sched_trace = pb.new_trace();
    CUDA_TRY(cudaDeviceSynchronize());
}

static Plan* get_plan(s3d_unet* u, int B, int H, int W, int D) {
    if (!u->finalized) throw S3dError{"s3d_unet_finalize() has not been called"};
    S3D_CHECK(B >= 1 && H >= 1 && W >= 1 && D >= 1, "bad shape");
    if (!u->plan || u->plan->B != B || u->plan->H != H || u->plan->W != W || u->plan->D != D) build_plan(u, B, H, W, D);
    return u->plan.get();
}

// film rows for n timesteps: sinusoid -> Linear -> SiLU -> Linear -> SiLU -> concatenated emb_layers
static void run_film(s3d_unet* u, Plan* P, const float* t_dev, int n, float* film_dev, cudaStream_t s) {
    if (P->emb_rows < n) {
        for (int i = 0; i < 3; ++i) P->emb_tmp[i] = dev_alloc<float>(P->allocs, static_cast<size_t>(n) * u->emb_dim);
        P->emb_rows = n;
    }
    const int mc = u->cfg.model_channels, half = mc / 2, E = u->emb_dim;
    S3D_CHECK(mc % 2 == 0, "odd model_channels");
    launch(k_sinusoid, dim3((n * half + 255) / 256), dim3(256), 0, s, t_dev, u->freqs, half, P->emb_tmp[0], n);
    LAUNCH_CHECK("k_sinusoid");
    launch(k_linear, dim3(dim3((E + 7) / 8, n)), dim3(256), 0, s, P->emb_tmp[0], u->te_w0, u->te_b0, P->emb_tmp[1], mc, E, 0);
    LAUNCH_CHECK("k_linear");
    launch(k_linear, dim3(dim3((E + 7) / 8, n)), dim3(256), 0, s, P->emb_tmp[1], u->te_w2, u->te_b2, P->emb_tmp[2], E, E, 1);
    LAUNCH_CHECK("k_linear");
    launch(k_linear, dim3(dim3((u->film_dim + 7) / 8, n)), dim3(256), 0, s, P->emb_tmp[2], u->film_w, u->film_b, film_dev, E, u->film_dim, 1);
    LAUNCH_CHECK("k_linear");
}

static void clear_stale_in_acc(Plan* P, cudaStream_t s) {
    if (P->in_acc_stale) {
        CUDA_TRY(cudaMemsetAsync(P->in_acc, 0, P->in_acc_n * sizeof(unsigned long long), s));
        P->in_acc_stale = false;
    }
}
static void train_forward_prologue(Plan* P, cudaStream_t s) {
    if (!P->train) return;
    for (auto& z : P->fwd_zero) CUDA_TRY(cudaMemsetAsync(z.first, 0, z.second, s));
}
static void run_ops(s3d_unet* u, Plan* P, cudaStream_t s) {
    // experiment switch (timing only, results are wrong): S3D_DUP_OPS=1 launches every op twice back to back, so that the
    // trace of the second launch shows what a warm instruction / constant cache is worth
    static const bool dup = getenv("S3D_DUP_OPS") && atoi(getenv("S3D_DUP_OPS")) != 0;
    if (dup) {
        for (auto& op : P->ops) {
            op(s);
            op(s);
        }
        u->last_launches = static_cast<int>(P->ops.size());
        return;
    }
    for (auto& op : P->ops) op(s);
    u->last_launches = static_cast<int>(P->ops.size());
}

static void launch_sched(const SchedArgs& A, cudaStream_t s) {
    const long long n4 = A.hw * ((A.C + 3) / 4);
    int gx = static_cast<int>(std::min<long long>((n4 + 255) / 256, 148LL * 8));
    gx = std::max(gx, 1);
    launch(k_sched_step, dim3(dim3(gx, A.B)), dim3(256), 0, s, A);
    LAUNCH_CHECK("k_sched_step");
}

// ------------------------------------------------------------------------------------ C ABI
#define API_BEGIN try {
#define API_END                                   \
    }                                             \
    catch (const S3dError& e) { return fail(e.msg); } \
    catch (const std::exception& e) { return fail(e.what()); } \
    return 0;

extern "C" {

int s3d_abi_version(void) { return 4; }
const char* s3d_last_error(void) { return g_err.c_str(); }

int s3d_unet_create(const s3d_unet_config* cfg, int device, s3d_unet** out) {
    API_BEGIN
    S3D_CHECK(cfg && out, "null argument");
    S3D_CHECK(cfg->num_res_blocks == 1,
              "num_res_blocks must be 1: the reference constructor (unet_triplane.py:389-405) cannot build a runnable "
              "model otherwise");
    S3D_CHECK(cfg->n_levels >= 1 && cfg->n_levels <= S3D_MAX_LEVELS, "n_levels out of range");
    S3D_CHECK(cfg->model_channels > 0 && cfg->model_channels % 64 == 0, "model_channels must be a multiple of 64");
    S3D_CHECK(cfg->in_channels >= 1 && cfg->out_channels >= 1 && cfg->out_channels <= 64, "channel counts out of range");
    for (int l = 0; l < cfg->n_levels; ++l) S3D_CHECK(cfg->channel_mult[l] >= 1, "channel_mult must be >= 1");
    S3D_CHECK(cfg->precision >= 1 && cfg->precision <= 5, "precision must be 1 .. 5");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    S3D_CHECK(device >= 0 && device < ndev, "no such CUDA device");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    S3D_CHECK(prop.major == 10, "sin3dm_b200 kernels are built for sm_100a (B200) only");
    std::unique_ptr<s3d_unet> u(new s3d_unet());
    u->cfg = *cfg;
    u->device = device;
    u->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("S3D_FUSE_ROLL")) u->fuse_roll = atoi(e) != 0;
    if (const char* e = getenv("S3D_PDL")) g_pdl = atoi(e) != 0;
    if (const char* e = getenv("S3D_HALO_BO_KW")) u->halo_bo_kw = atoi(e) != 0;
    if (const char* e = getenv("S3D_FUSE_POOL")) u->fuse_pool = atoi(e) != 0;
    if (const char* e = getenv("S3D_ROLL_TM")) u->roll_tm = atoi(e);
    u->bwd_wgrad_ffma = false;
    if (const char* e = getenv("S3D_WGRAD")) u->bwd_wgrad_ffma = std::string(e) == "ffma";      // CUDA-core cross-check kernel
    build_structure(u.get());
    *out = u.release();
    API_END
}

int s3d_unet_destroy(s3d_unet* u) {
    if (!u) return 0;
    cudaSetDevice(u->device);
    destroy_plan(u);
    free_all(u->wallocs);
    if (u->pack_jobs_dev) cudaFree(u->pack_jobs_dev);
    delete u;
    return 0;
}

int s3d_unet_num_tensors(const s3d_unet* u) { return u ? static_cast<int>(u->tensors.size()) : 0; }

int s3d_unet_tensor_info(const s3d_unet* u, int index, const char** name, int* ndim, int64_t shape[4]) {
    API_BEGIN
    S3D_CHECK(u && index >= 0 && index < static_cast<int>(u->tensors.size()), "bad index");
    const TensorSpec& t = u->tensors[index];
    if (name) *name = t.name.c_str();
    if (ndim) *ndim = static_cast<int>(t.shape.size());
    if (shape)
        for (size_t i = 0; i < 4; ++i) shape[i] = i < t.shape.size() ? t.shape[i] : 1;
    API_END
}

int s3d_unet_load_tensor(s3d_unet* u, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    API_BEGIN
    S3D_CHECK(u && name && host_data && shape, "null argument");
    std::string n(name);
    if (n == "__freqs") {
        S3D_CHECK(ndim == 1 && shape[0] == u->cfg.model_channels / 2, "__freqs shape");
        if (!u->index.count(n)) add_tensor(u, n, {shape[0]});
    }
    auto it = u->index.find(n);
    if (it == u->index.end()) throw S3dError{"unexpected key in state_dict: " + n};
    TensorSpec& t = u->tensors[it->second];
    bool same = static_cast<int>(t.shape.size()) == ndim;
    for (int i = 0; same && i < ndim; ++i) same = t.shape[i] == shape[i];
    if (!same) throw S3dError{"size mismatch for " + n};
    t.host.assign(host_data, host_data + t.numel());
    t.loaded = true;
    u->finalized = false;
    API_END
}

int s3d_unet_finalize(s3d_unet* u) {
    API_BEGIN
    S3D_CHECK(u, "null handle");
    finalize(u);
    API_END
}

int s3d_unet_film_dim(const s3d_unet* u) { return u ? u->film_dim : 0; }
int s3d_unet_last_launches(const s3d_unet* u) { return u ? u->last_launches : 0; }

int s3d_unet_film(s3d_unet* u, const float* t_dev, int n, float* film_dev, void* stream) {
    API_BEGIN
    S3D_CHECK(u && t_dev && film_dev && n >= 1, "bad argument");
    if (!u->finalized) throw S3dError{"s3d_unet_finalize() has not been called"};
    CUDA_TRY(cudaSetDevice(u->device));
    if (!u->plan) build_plan(u, 1, 8, 8, 8);   // scratch owner
    run_film(u, u->plan.get(), t_dev, n, film_dev, static_cast<cudaStream_t>(stream));
    API_END
}

int s3d_unet_forward_film(s3d_unet* u, const float* x_dev, const float* film_dev, const int* row_dev, float* out_dev, int B,
                          int H, int W, int D, void* stream) {
    API_BEGIN
    S3D_CHECK(u && x_dev && film_dev && out_dev, "null argument");
    CUDA_TRY(cudaSetDevice(u->device));
    Plan* P = get_plan(u, B, H, W, D);
    P->x = x_dev;
    P->out = out_dev;
    P->film = film_dev;
    P->film_row = row_dev;
    clear_stale_in_acc(P, static_cast<cudaStream_t>(stream));
    train_forward_prologue(P, static_cast<cudaStream_t>(stream));
    run_ops(u, P, static_cast<cudaStream_t>(stream));
    API_END
}

int s3d_unet_forward(s3d_unet* u, const float* x_dev, const float* t_dev, float* out_dev, int B, int H, int W, int D,
                     void* stream) {
    API_BEGIN
    S3D_CHECK(u && x_dev && t_dev && out_dev, "null argument");
    CUDA_TRY(cudaSetDevice(u->device));
    Plan* P = get_plan(u, B, H, W, D);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    run_film(u, P, t_dev, B, P->film_own, s);
    P->x = x_dev;
    P->out = out_dev;
    P->film = P->film_own;
    P->film_row = nullptr;
    clear_stale_in_acc(P, s);
    train_forward_prologue(P, s);
    run_ops(u, P, s);
    u->last_launches += 4;
    API_END
}

static SchedArgs to_sched(const s3d_sched_args* a) {
    SchedArgs A{};
    A.kind = a->kind;
    A.mean_type = a->mean_type;
    A.clip = a->clip_denoised;
    A.is_mask_t0 = a->is_mask_t0;
    A.B = a->B;
    A.n = a->n_per_sample;
    A.C = a->C;
    A.hw = a->n_per_sample / a->C;
    A.model_out = a->model_out;
    A.x = a->x;
    A.noise = a->noise;
    A.y0 = a->y0;
    A.mask = a->mask;
    A.sample = a->sample;
    A.x0_out = a->pred_xstart;
    A.coef = a->coef_dev;
    A.t_idx = const_cast<int*>(a->t_idx_dev);
    A.seed = a->seed;
    A.sample_base = a->sample_base;
    return A;
}

int s3d_sched_step(const s3d_sched_args* a, void* stream) {
    API_BEGIN
    S3D_CHECK(a && a->model_out && a->x && a->sample && a->coef_dev && a->t_idx_dev, "null argument");
    S3D_CHECK(a->kind >= 0 && a->kind <= 2 && (a->mean_type == 0 || a->mean_type == 1), "bad kind / mean_type");
    S3D_CHECK((a->y0 == nullptr) == (a->mask == nullptr), "y0 and mask go together");
    S3D_CHECK(a->B >= 1 && a->n_per_sample >= 1 && a->C >= 1 && a->n_per_sample % a->C == 0, "bad shape");
    launch_sched(to_sched(a), static_cast<cudaStream_t>(stream));
    API_END
}

int s3d_q_sample(const float* x0_dev, const float* noise_dev, float* out_dev, const float* coef_dev, const int* t_idx_dev, int B,
                 int64_t n, void* stream) {
    API_BEGIN
    S3D_CHECK(x0_dev && noise_dev && out_dev && coef_dev && t_idx_dev && B >= 1 && n >= 1, "bad argument");
    int gx = static_cast<int>(std::min<long long>((n / 4 + 255) / 256 + 1, 148LL * 8));
    launch(k_q_sample, dim3(dim3(gx, B)), dim3(256), 0, static_cast<cudaStream_t>(stream), x0_dev, noise_dev, out_dev, coef_dev, t_idx_dev, n);
    LAUNCH_CHECK("k_q_sample");
    API_END
}

// CTAs per sample (the batch is the grid's y dimension): about eight float4 per thread so that the block reduction is amortised, but
// never fewer CTAs in total than four per SM
static int vb_grid(int64_t n, int B) {
    const long long quads = (n / 4 + 255) / 256;                          // one float4 per thread
    const long long want = std::max<long long>((quads + 7) / 8, (148LL * 4 + B - 1) / std::max(B, 1));
    return static_cast<int>(std::max<long long>(1, std::min<long long>(want, quads)));
}
int64_t s3d_vb_workspace_bytes(int B, int64_t n) { return static_cast<int64_t>(B) * vb_grid(n, B) * 3 * sizeof(double); }
int s3d_vb_terms(const s3d_vb_args* a, void* stream) {
    API_BEGIN
    S3D_CHECK(a && a->x_start && a->x_t && a->model_out && a->coef_dev && a->logvar_dev && a->t_idx_dev && a->workspace && a->out &&
                  a->B >= 1 && a->n_per_sample >= 1, "bad argument");
    S3D_CHECK(a->mean_type == S3D_START_X || a->mean_type == S3D_EPSILON, "mean_type");
    VbArgs A{};
    A.mean_type = a->mean_type;
    A.clip = a->clip_denoised;
    A.B = a->B;
    A.n = a->n_per_sample;
    A.x_start = a->x_start;
    A.x_t = a->x_t;
    A.model_out = a->model_out;
    A.noise = a->noise;
    A.x0_out = a->pred_xstart;
    A.coef = a->coef_dev;
    A.logvar = a->logvar_dev;
    A.t_idx = a->t_idx_dev;
    A.partial = static_cast<double*>(a->workspace);
    A.out = a->out;
    const int gx = vb_grid(A.n, A.B);
    launch_plain(k_vb_terms, dim3(gx, A.B), dim3(256), 0, static_cast<cudaStream_t>(stream), A);
    LAUNCH_CHECK("k_vb_terms");
    launch_plain(k_vb_finalize, dim3(A.B), dim3(96), 0, static_cast<cudaStream_t>(stream), A, gx);
    LAUNCH_CHECK("k_vb_finalize");
    API_END
}

int s3d_plane_mse(const float* target_dev, const float* output_dev, int B, int C, int H, int W, int D, void* workspace, float* out_dev,
                  void* stream) {
    API_BEGIN
    S3D_CHECK(target_dev && output_dev && workspace && out_dev && B >= 1 && C >= 1 && H >= 1 && W >= 1 && D >= 1, "bad argument");
    PlaneMseArgs A{};
    A.target = target_dev;
    A.output = output_dev;
    A.C = C; A.H = H; A.W = W; A.D = D;
    A.n = static_cast<long long>(C) * (H + D) * (W + D);
    A.partial = static_cast<double*>(workspace);
    A.out = out_dev;
    const int gx = vb_grid(A.n, B);
    launch_plain(k_plane_mse, dim3(gx, B), dim3(256), 0, static_cast<cudaStream_t>(stream), A);
    LAUNCH_CHECK("k_plane_mse");
    launch_plain(k_plane_mse_finalize, dim3(B), dim3(96), 0, static_cast<cudaStream_t>(stream), A, gx);
    LAUNCH_CHECK("k_plane_mse_finalize");
    API_END
}

int s3d_adamw_ema_step(const s3d_adamw_args* a, void* stream) {
    API_BEGIN
    S3D_CHECK(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq && a->n >= 1 && a->step >= 1, "bad argument");
    S3D_CHECK(a->n_ema >= 0 && a->n_ema <= 4, "at most 4 EMA buffers");
    S3D_CHECK(((reinterpret_cast<uintptr_t>(a->param) | reinterpret_cast<uintptr_t>(a->grad) | reinterpret_cast<uintptr_t>(a->exp_avg) |
                reinterpret_cast<uintptr_t>(a->exp_avg_sq)) & 15) == 0, "buffers must be 16-byte aligned");
    AdamWArgs A{};
    A.p = a->param;
    A.g = a->grad;
    A.m = a->exp_avg;
    A.v = a->exp_avg_sq;
    A.n_ema = a->n_ema;
    for (int k = 0; k < a->n_ema; ++k) {
        S3D_CHECK(a->ema[k] && (reinterpret_cast<uintptr_t>(a->ema[k]) & 15) == 0, "EMA buffer missing or misaligned");
        A.ema[k] = a->ema[k];
        A.ema_rate[k] = a->ema_rate[k];
    }
    A.n = a->n;
    // scalars exactly as torch forms them: python floats (fp64), rounded to fp32 where they meet the tensors
    const double lr = a->lr, b1 = a->beta1, b2 = a->beta2;
    A.decay = static_cast<float>(1.0 - lr * a->weight_decay);
    A.w1 = static_cast<float>(1.0 - b1);
    A.beta2 = static_cast<float>(b2);
    A.w2 = static_cast<float>(1.0 - b2);
    A.bc2_sqrt = static_cast<float>(std::sqrt(1.0 - std::pow(b2, a->step)));
    A.step_size = static_cast<float>(lr / (1.0 - std::pow(b1, a->step)));
    A.eps = static_cast<float>(a->eps);
    const int gx = static_cast<int>(std::min<long long>((A.n / 4 + 255) / 256 + 1, 148LL * 8));
    launch_plain(k_adamw_ema, dim3(gx), dim3(256), 0, static_cast<cudaStream_t>(stream), A);
    LAUNCH_CHECK("k_adamw_ema");
    API_END
}

int s3d_philox_normal(float* out_dev, int B, int C, int64_t hw, uint64_t seed, uint32_t sample_base, uint32_t step, void* stream) {
    API_BEGIN
    S3D_CHECK(out_dev && B >= 1 && C >= 1 && hw >= 1, "bad argument");
    int gx = static_cast<int>(std::min<long long>((hw * ((C + 3) / 4) + 255) / 256, 148LL * 8));
    launch(k_philox_normal, dim3(dim3(gx, B)), dim3(256), 0, static_cast<cudaStream_t>(stream), out_dev, C, hw, seed, sample_base, step);
    LAUNCH_CHECK("k_philox_normal");
    API_END
}

// loop state in device memory: step index of every sample, Philox seed and global index of sample 0
__global__ void k_set_loop_state(int* t_idx, int n, int t_start, unsigned long long* dyn, unsigned long long seed, unsigned int sample_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t_idx[i] = t_start;
    if (i == 0) {
        dyn[0] = seed;
        dyn[1] = sample_base;
    }
}

int s3d_sample_loop(s3d_unet* u, const s3d_loop_args* a, void* stream) {
    API_BEGIN
    S3D_CHECK(u && a && a->x_dev && a->coef_dev && a->film_dev, "null argument");
    S3D_CHECK((a->kind == S3D_DDPM || a->kind == S3D_DDIM) && (a->mean_type == 0 || a->mean_type == 1), "bad kind / mean_type");
    S3D_CHECK((a->y0_dev == nullptr) == (a->mask_dev == nullptr), "y0 and mask go together");
    S3D_CHECK(a->n_steps >= 1 && a->t_start >= a->n_steps - 1, "n_steps / t_start");
    CUDA_TRY(cudaSetDevice(u->device));
    const int Hc = a->H + a->D, Wc = a->W + a->D;
    S3D_CHECK(u->cfg.in_channels == u->cfg.out_channels, "sampling needs in_channels == out_channels");
    const long long n = static_cast<long long>(u->cfg.out_channels) * Hc * Wc;
    S3D_CHECK(a->n_per_sample == n, "x_dev does not have out_channels * (H+D) * (W+D) elements per sample");
    S3D_CHECK(!u->training, "the sampling loop needs an inference plan: s3d_unet_set_training(u, 0) first");
    Plan* P = get_plan(u, a->B, a->H, a->W, a->D);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!P->t_idx) {
        P->t_idx = dev_alloc<int>(P->allocs, a->B);
        P->ticket = dev_alloc<unsigned int>(P->allocs, 1);
        CUDA_TRY(cudaMemset(P->ticket, 0, sizeof(unsigned int)));
        P->model_out = dev_alloc<float>(P->allocs, static_cast<size_t>(a->B) * n);
        P->dyn = dev_alloc<unsigned long long>(P->allocs, 2);
    }
    launch_plain(k_set_loop_state, dim3((a->B + 255) / 256), dim3(256), 0, s, P->t_idx, a->B, a->t_start, P->dyn,
                 static_cast<unsigned long long>(a->seed), a->sample_base);
    LAUNCH_CHECK("k_set_loop_state");
    P->x = a->x_dev;
    P->out = P->model_out;
    P->film = a->film_dev;
    P->film_row = P->t_idx;
    SchedArgs A{};
    A.kind = a->kind;
    A.mean_type = a->mean_type;
    A.clip = a->clip_denoised;
    A.is_mask_t0 = a->is_mask_t0;
    A.B = a->B;
    A.n = n;
    A.C = u->cfg.out_channels;
    A.hw = static_cast<long long>(Hc) * Wc;
    A.model_out = P->model_out;
    A.x = a->x_dev;
    A.noise = a->step_noise_dev;
    A.noise_step_stride = a->step_noise_dev ? static_cast<long long>(a->B) * n : 0;
    A.y0 = a->y0_dev;
    A.mask = a->mask_dev;
    A.sample = a->x_dev;
    A.x0_out = a->pred_xstart_dev;
    A.coef = a->coef_dev;
    A.t_idx = P->t_idx;
    A.dyn = P->dyn;
    A.advance = 1;
    A.ticket = P->ticket;
    A.tr = P->sched_trace;
    // Loop structure: in_conv once, then per step [blocks ..., fused (out head + scheduler + next step's in_conv)].
    // S3D_FUSED_BOUNDARY=0 runs head, scheduler and in_conv as three launches instead (same device code, identical values)
    const char* fb = getenv("S3D_FUSED_BOUNDARY");
    const bool fused = static_cast<bool>(P->fused_boundary) && !(fb && atoi(fb) == 0);
    const size_t nops = P->ops.size();
    auto one_step = [&](cudaStream_t st) {
        if (fused) {
            for (size_t i = 1; i + 1 < nops; ++i) P->ops[i](st);
            P->fused_boundary(st, A);
        } else {
            run_ops(u, P, st);
            launch_sched(A, st);
        }
    };
    clear_stale_in_acc(P, s);
    if (fused) P->ops[0](s);
    if (!a->use_graph) {
        for (int i = 0; i < a->n_steps; ++i) one_step(s);
    } else {
        // The graph bakes the pointers and options of one step; what changes from call to call with the same buffers — step
        // index, seed, sample_base — lives in device memory, so the cached graph is replayed as is.
        char key[512];
        snprintf(key, sizeof(key), "%d|%d|%d|%d|%p|%p|%p|%p|%p|%p|%p|%d", a->kind, a->mean_type, a->clip_denoised,
                 a->is_mask_t0, (void*)a->x_dev, (void*)a->pred_xstart_dev, (void*)a->coef_dev, (void*)a->film_dev,
                 (void*)a->step_noise_dev, (void*)a->y0_dev, (void*)a->mask_dev, fused ? 1 : 0);
        auto it = P->graphs.find(key);
        if (it == P->graphs.end()) {
            if (P->graphs.size() >= 8) {                    // bounded cache: callers that keep changing buffers just re-capture
                for (auto& kv : P->graphs) cudaGraphExecDestroy(kv.second);
                P->graphs.clear();
            }
            cudaStream_t cs;
            CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                try {
                    one_step(cs);
                } catch (...) {
                    cudaStreamEndCapture(cs, &g);
                    if (g) cudaGraphDestroy(g);
                    cudaStreamDestroy(cs);
                    throw;
                }
                e = cudaStreamEndCapture(cs, &g);
            }
            cudaStreamDestroy(cs);
            if (e != cudaSuccess) throw S3dError{std::string("graph capture: ") + cudaGetErrorString(e)};
            cudaGraphExec_t ge = nullptr;
            e = cudaGraphInstantiate(&ge, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) throw S3dError{std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)};
            it = P->graphs.emplace(key, ge).first;
            ++u->graph_builds;
        }
        for (int i = 0; i < a->n_steps; ++i) CUDA_TRY(cudaGraphLaunch(it->second, s));
    }
    if (fused) P->in_acc_stale = true;
    u->last_launches = fused ? static_cast<int>(nops) - 2 : static_cast<int>(nops);
    u->last_launches += 1;   // scheduler kernel
    API_END
}

int s3d_unet_graph_builds(const s3d_unet* u) { return u ? u->graph_builds : 0; }

// ---------------------------------------------------------------------------------- training
int s3d_unet_set_training(s3d_unet* u, int on) {
    API_BEGIN
    S3D_CHECK(u, "null handle");
    const bool want = on != 0;
    if (u->training != want) {
        CUDA_TRY(cudaSetDevice(u->device));
        CUDA_TRY(cudaDeviceSynchronize());
        destroy_plan(u);              // the next forward rebuilds it with (or without) the tape and the backward list
        u->training = want;
        if (want) {
            if (u->grad_numel == 0) build_grad_layout(u);
            if (u->finalized) build_dgrad_packs(u);
        }
    }
    API_END
}

int s3d_unet_refresh_dev(s3d_unet* u, const float* const* tensors_dev, int n, void* stream) {
    API_BEGIN
    S3D_CHECK(u && tensors_dev && n >= 1, "bad argument");
    CUDA_TRY(cudaSetDevice(u->device));
    refresh_from_device(u, tensors_dev, n, static_cast<cudaStream_t>(stream));
    API_END
}

int64_t s3d_unet_grad_numel(s3d_unet* u) {
    if (!u) return 0;
    if (u->grad_numel == 0) build_grad_layout(u);
    return u->grad_numel;
}

int64_t s3d_unet_grad_offset(s3d_unet* u, int index) {
    if (!u || index < 0 || index >= static_cast<int>(u->tensors.size())) return -1;
    if (u->grad_numel == 0) build_grad_layout(u);
    return u->grad_off[index];
}

int s3d_unet_backward(s3d_unet* u, const float* grad_out_dev, float* grads_dev, float* dfilm_dev, void* stream) {
    API_BEGIN
    S3D_CHECK(u && grad_out_dev && grads_dev, "null argument");
    S3D_CHECK(u->training && u->plan && u->plan->train, "no training forward to differentiate: s3d_unet_set_training(u, 1), then a forward");
    Plan* P = u->plan.get();
    S3D_CHECK(P->x && P->film, "run the forward first");
    CUDA_TRY(cudaSetDevice(u->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (auto& z : P->bwd_zero) CUDA_TRY(cudaMemsetAsync(z.first, 0, z.second, s));
    P->grad_out = grad_out_dev;
    for (auto& op : P->bwd_ops) op(s);
    CUDA_TRY(cudaMemcpyAsync(grads_dev, P->grads_own, sizeof(float) * static_cast<size_t>(u->grad_numel), cudaMemcpyDeviceToDevice, s));
    if (dfilm_dev)
        CUDA_TRY(cudaMemcpyAsync(dfilm_dev, P->dfilm_own, sizeof(float) * static_cast<size_t>(P->B) * u->film_dim, cudaMemcpyDeviceToDevice, s));
    u->last_launches = static_cast<int>(P->bwd_ops.size());
    API_END
}

int s3d_unet_bwd_op_count(const s3d_unet* u) { return (u && u->plan) ? static_cast<int>(u->plan->bwd_ops.size()) : 0; }

int s3d_unet_bwd_op_info(const s3d_unet* u, int index, const char** kernel, double* dense_flops) {
    API_BEGIN
    S3D_CHECK(u && u->plan && index >= 0 && index < static_cast<int>(u->plan->bwd_ops.size()), "bad index");
    if (kernel) *kernel = u->plan->bwd_names[index].c_str();
    if (dense_flops) *dense_flops = u->plan->bwd_flops[index];
    API_END
}

// Mean device time (ms) of every backward op: `iters` eager passes with a CUDA-event pair around each launch.  Needs a completed
// forward + backward (uses their bindings); synchronises.
int s3d_unet_profile_bwd_ops(s3d_unet* u, int iters, float* ms_out, void* stream) {
    API_BEGIN
    S3D_CHECK(u && u->plan && u->plan->train && ms_out && iters >= 1, "bad argument");
    Plan* P = u->plan.get();
    S3D_CHECK(P->grad_out, "run a backward first");
    CUDA_TRY(cudaSetDevice(u->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = P->bwd_ops.size();
    std::vector<double> acc(n, 0.0);
    std::vector<cudaEvent_t> ev(2 * n);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    for (int it = 0; it < iters + 1; ++it) {          // first pass is a warm-up
        for (auto& z : P->bwd_zero) CUDA_TRY(cudaMemsetAsync(z.first, 0, z.second, s));
        for (size_t i = 0; i < n; ++i) {
            CUDA_TRY(cudaEventRecord(ev[2 * i], s));
            P->bwd_ops[i](s);
            CUDA_TRY(cudaEventRecord(ev[2 * i + 1], s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        for (size_t i = 0; i < n && it > 0; ++i) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
            acc[i] += ms;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    for (size_t i = 0; i < n; ++i) ms_out[i] = static_cast<float>(acc[i] / iters);
    API_END
}

int64_t s3d_unet_workspace_bytes(const s3d_unet* u) { return (u && u->plan) ? static_cast<int64_t>(u->plan->alloc_bytes) : 0; }

int s3d_unet_op_count(const s3d_unet* u) { return (u && u->plan) ? static_cast<int>(u->plan->ops.size()) : 0; }

int s3d_unet_op_info(const s3d_unet* u, int index, const char** kernel, double* dense_flops) {
    API_BEGIN
    S3D_CHECK(u && u->plan && index >= 0 && index < static_cast<int>(u->plan->ops.size()), "bad index");
    if (kernel) *kernel = u->plan->op_names[index].c_str();
    if (dense_flops) *dense_flops = u->plan->op_flops[index];
    API_END
}

// Re-runs the UNet ops of the current plan `iters` times with a CUDA-event pair around every launch (on `stream`)
// and returns the mean device time per op in milliseconds.  Synchronises.  Uses the bindings of the last forward / loop.
int s3d_unet_profile_ops(s3d_unet* u, int iters, float* ms_out, void* stream) {
    API_BEGIN
    S3D_CHECK(u && u->plan && ms_out && iters >= 1, "bad argument");
    Plan* P = u->plan.get();
    S3D_CHECK(P->x && P->out && P->film, "run a forward or a sampling loop first");
    CUDA_TRY(cudaSetDevice(u->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    clear_stale_in_acc(P, s);
    const size_t n = P->ops.size();
    std::vector<double> acc(n, 0.0);
    // Preferred: the ops are captured into ONE graph with an external event-record node between consecutive launches and
    // the graph is replayed: op i's time = event[i+1] - event[i] in steady state (warm caches, graph launch latencies),
    // which is what the sampling loop sees.  Eager launches mostly measure the host's launch rate at these kernel sizes;
    // they are the fallback when the driver refuses timing on graph-recorded events.
    bool graph_ok = false;
    {
        std::vector<cudaEvent_t> ev(n + 1);
        for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
        cudaStream_t cs = nullptr;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        bool ok = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess;
        bool capturing = false;
        try {
            if (ok) {
                CUDA_TRY(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
                capturing = true;
                for (size_t i = 0; i < n && ok; ++i) {
                    ok = cudaEventRecordWithFlags(ev[i], cs, cudaEventRecordExternal) == cudaSuccess;
                    if (ok) P->ops[i](cs);
                }
                ok = ok && cudaEventRecordWithFlags(ev[n], cs, cudaEventRecordExternal) == cudaSuccess;
                capturing = false;
                ok = (cudaStreamEndCapture(cs, &g) == cudaSuccess) && ok;
                ok = ok && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess;
            }
            for (int it = 0; ok && it < iters + 3; ++it) {                          // 3 warm-up replays
                ok = cudaGraphLaunch(ge, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
                for (size_t i = 0; ok && i < n && it >= 3; ++i) {
                    float ms = 0.f;
                    ok = cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess;
                    acc[i] += ms;
                }
            }
        } catch (...) {
            if (capturing) cudaStreamEndCapture(cs, &g);
            ok = false;
        }
        (void)cudaGetLastError();
        if (ge) cudaGraphExecDestroy(ge);
        if (g) cudaGraphDestroy(g);
        if (cs) cudaStreamDestroy(cs);
        for (auto& e : ev) cudaEventDestroy(e);
        graph_ok = ok;
    }
    if (!graph_ok) {
        std::fill(acc.begin(), acc.end(), 0.0);
        std::vector<cudaEvent_t> ev(2 * n);
        for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
        for (int it = 0; it < iters + 1; ++it) {          // first pass is a warm-up
            for (size_t i = 0; i < n; ++i) {
                CUDA_TRY(cudaEventRecord(ev[2 * i], s));
                P->ops[i](s);
                CUDA_TRY(cudaEventRecord(ev[2 * i + 1], s));
            }
            CUDA_TRY(cudaStreamSynchronize(s));
            for (size_t i = 0; i < n && it > 0; ++i) {
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
                acc[i] += ms;
            }
        }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    for (size_t i = 0; i < n; ++i) ms_out[i] = static_cast<float>(acc[i] / iters);
    u->profile_mode = graph_ok ? 1 : 0;
    API_END
}

int s3d_unet_profile_mode(const s3d_unet* u) { return u ? u->profile_mode : -1; }

int s3d_unet_trace_enable(s3d_unet* u, int on) {
    API_BEGIN
    S3D_CHECK(u, "null handle");
    if (u->trace_on != (on != 0)) {
        CUDA_TRY(cudaSetDevice(u->device));
        CUDA_TRY(cudaDeviceSynchronize());
        destroy_plan(u);          // the next forward / loop rebuilds it with (or without) stamp buffers
        u->trace_on = on != 0;
    }
    API_END
}

int s3d_trace_slots(void) { return kTraceSlots; }

int s3d_unet_trace_read(s3d_unet* u, int op_index, uint64_t* host_out, int max_ctas) {
    API_BEGIN
    S3D_CHECK(u && u->plan && host_out && max_ctas >= 1, "bad argument");
    const Plan* P = u->plan.get();
    S3D_CHECK(op_index >= -2 && op_index < static_cast<int>(P->ops.size()), "bad op index");
    const Trace t = op_index == -2 ? P->fused_trace : (op_index < 0 ? P->sched_trace : P->op_trace[op_index]);
    S3D_CHECK(t.buf != nullptr, "tracing is off (s3d_unet_trace_enable)");
    CUDA_TRY(cudaSetDevice(u->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const int n = std::min(max_ctas, t.max_ctas);
    CUDA_TRY(cudaMemcpy(host_out, t.buf, sizeof(uint64_t) * n * 2 * kTraceSlots, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(t.buf, 0, sizeof(uint64_t) * static_cast<size_t>(t.max_ctas) * 2 * kTraceSlots));
    API_END
}

int s3d_unet_debug_count(const s3d_unet* u) { return (u && u->plan) ? static_cast<int>(u->plan->named.size()) : 0; }

int s3d_unet_debug_info(const s3d_unet* u, int index, const char** name, int* channels, int rows[3], int cols[3],
                        int* batch) {
    API_BEGIN
    S3D_CHECK(u && u->plan && index >= 0 && index < static_cast<int>(u->plan->named.size()), "bad index");
    const NamedBuf& nb = u->plan->named[index];
    if (name) *name = nb.name.c_str();
    if (channels) *channels = nb.C;
    if (batch) *batch = u->plan->B;
    for (int p = 0; p < 3; ++p) {
        if (rows) rows[p] = nb.d.rows[p];
        if (cols) cols[p] = nb.d.cols[p];
    }
    API_END
}

int s3d_unet_debug_read(s3d_unet* u, int index, int plane, float* host_out, int64_t n_floats) {
    API_BEGIN
    S3D_CHECK(u && u->plan && index >= 0 && index < static_cast<int>(u->plan->named.size()) && plane >= 0 && plane < 3,
              "bad index");
    const NamedBuf& nb = u->plan->named[index];
    const int64_t want = static_cast<int64_t>(u->plan->B) * nb.d.rows[plane] * nb.d.cols[plane] * nb.C;
    S3D_CHECK(n_floats == want, "buffer size mismatch");
    CUDA_TRY(cudaSetDevice(u->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(host_out, nb.p.p[plane], sizeof(float) * want, cudaMemcpyDeviceToHost));
    API_END
}

}  // extern "C"
