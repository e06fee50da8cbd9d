// Host side of the triplane decoder: handle, checkpoint ingestion / weight packing, C ABI.
// See include/sin3dm_b200.h (s3d_decoder_*) for the contract and the reference file:line each entry point replaces.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sin3dm_b200.h"
#include "dec_kernels.cuh"
#include "host_util.cuh"

using namespace s3d;

namespace {

const char* kPlaneName[3] = {"xy", "xz", "yz"};

struct DTensor {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> host;
    bool loaded = false, needed = true;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

// One TriplaneGroupResnetBlock (blocks.py:189-256) of a feature branch
struct ConvBlk {
    std::string prefix;                     // "geo_convs." / "tex_convs." / "tex_convs.0." ...
    int cin = 0, ks = 0;
    bool in_norm = false;                   // input_norm + input_act (PBR's second texture block): identity shortcut on the normed input
    float *w_in = nullptr, *b_in = nullptr, *w_out = nullptr, *b_out = nullptr, *ws = nullptr, *bs = nullptr;
    float *gamma = nullptr, *beta = nullptr;
};
// feature branch: geo (latent channels [0, geo)) or tex (latent channels [geo, geo + tex)); its planes are F[..][b * 64 ...]
struct Branch {
    int c0 = 0, cin = 0;
    std::vector<ConvBlk> blocks;
};
// One decoder MLP (DecoderMLPSkipConcat / DecoderMLP) reading the planes of branch `feat`
struct Head {
    std::string prefix;                     // "geo_decoder." ...
    int feat = 0, n_out = 1, out_col = 0, sigmoid = 0;
    DecMlpF32 mf{};                         // CUDA-core kernel: transposed fp32 weights
    uint16_t* w16[kDecLayers] = {};         // tensor-core kernel: [2][256][K] fp16 (hi, lo) of scale * W
    float inv_scale[kDecLayers] = {};
    std::vector<float> bias_h[kDecLayers], w_last_h, b_last_h;     // host copies: they travel as kernel parameters
    CUtensorMap maps[kDecLayers];
};

}  // namespace

struct s3d_decoder {
    s3d_decoder_config cfg{};
    int device = 0;
    std::vector<DTensor> tensors;
    std::map<std::string, int> index;
    bool finalized = false, tc_ok = false;
    std::vector<void*> owned;
    Branch br[2];
    std::vector<Head> heads;                // evaluation / output order
    int nb = 1, n_sm = 148;
    int out_channels = 1;
    // feature planes of the current latent
    std::vector<void*> plane_owned;
    float* F[3] = {nullptr, nullptr, nullptr};
    float* h1[3] = {nullptr, nullptr, nullptr};
    float* T1[3] = {nullptr, nullptr, nullptr};       // output of a branch's first block when a second one follows (PBR)
    double* partial = nullptr;
    float* coef = nullptr;
    int rows[3] = {0, 0, 0}, cols[3] = {0, 0, 0}, max_strips = 0;
    bool planes_set = false;
    int last_launches = 0;
    // encoder: fixed-point axis sums of the current volume
    unsigned long long* enc_sums = nullptr;
    size_t enc_sums_count = 0;
    float* enc_w[4] = {nullptr, nullptr, nullptr, nullptr};     // generic-encoder weights (freed with `owned`; reset by finalize)
};

namespace {

void add_tensor(s3d_decoder* d, const std::string& name, std::vector<int64_t> shape, bool needed = true) {
    DTensor t;
    t.name = name;
    t.shape = std::move(shape);
    t.needed = needed;
    d->index[name] = static_cast<int>(d->tensors.size());
    d->tensors.push_back(std::move(t));
}

// first_layers / second_layers of DecoderMLPSkipConcat in evaluation order (blocks.py:65-83)
struct LinName {
    std::string key;
    int cin, cout;
};
std::vector<LinName> mlp_layers(const s3d_decoder_config& c, const std::string& prefix, int n_out, int* n_first) {
    std::vector<LinName> v;
    const int nh = c.mlp_hidden_layers, up = c.feat_channel_up, hid = c.mlp_hidden_channels;
    if (c.mlp_kind == 1) {
        // DecoderMLP (blocks.py:46-62): one Sequential `layers`: Linear(up, hid), nh x Linear(hid, hid), Linear(hid, out); no concat
        for (int j = 0; j < nh + 2; ++j) v.push_back({prefix + "layers." + std::to_string(2 * j), j == 0 ? up : hid, j == nh + 1 ? n_out : hid});
        *n_first = -1;
        return v;
    }
    const int nf = 1 + nh / 2, ns = 1 + std::max(nh / 2 - 1, 0) + 1;
    for (int j = 0; j < nf; ++j) v.push_back({prefix + "first_layers." + std::to_string(2 * j), j == 0 ? up : hid, hid});
    for (int j = 0; j < ns; ++j)
        v.push_back({prefix + "second_layers." + std::to_string(2 * j), j == 0 ? up + hid : hid, j == ns - 1 ? n_out : hid});
    *n_first = nf;
    return v;
}

// Blocks and heads of the configured network:
//   net_kind 0  AutoEncoderGroupSkip / V3 (networks.py:134-162, 21-60): one ks x ks block per branch, heads geo(1) and tex(tex_channels, sigmoid)
//   net_kind 1  AutoEncoderGroupPBR (networks.py:227-262): geo block ks 5; tex blocks ks 3 (plain) + ks 3 (input norm + act);
//               heads geo(1), rgb(3), mr(2), normal(3), no sigmoid
void describe(s3d_decoder* d) {
    const auto& c = d->cfg;
    d->br[0] = Branch{};
    d->br[1] = Branch{};
    d->heads.clear();
    d->br[0].c0 = 0;
    d->br[0].cin = c.geo_feat_channels;
    d->br[1].c0 = c.geo_feat_channels;
    d->br[1].cin = c.tex_feat_channels;
    auto blk = [](const std::string& p, int cin, int ks, bool in_norm) {
        ConvBlk b;
        b.prefix = p;
        b.cin = cin;
        b.ks = ks;
        b.in_norm = in_norm;
        return b;
    };
    auto head = [](const std::string& p, int feat, int n_out, int col, int sig) {
        Head h;
        h.prefix = p;
        h.feat = feat;
        h.n_out = n_out;
        h.out_col = col;
        h.sigmoid = sig;
        return h;
    };
    if (c.net_kind == 1) {
        d->br[0].blocks.push_back(blk("geo_convs.", c.geo_feat_channels, 5, false));
        d->heads.push_back(head("geo_decoder.", 0, 1, 0, 0));
        if (c.use_tex) {
            d->br[1].blocks.push_back(blk("tex_convs.0.", c.tex_feat_channels, 3, false));
            d->br[1].blocks.push_back(blk("tex_convs.1.", c.feat_channel_up, 3, true));
            d->heads.push_back(head("rgb_decoder.", 1, 3, 1, 0));
            d->heads.push_back(head("mr_decoder.", 1, 2, 4, 0));
            d->heads.push_back(head("normal_decoder.", 1, 3, 6, 0));
        }
        d->out_channels = c.use_tex ? 9 : 1;
    } else {
        d->br[0].blocks.push_back(blk("geo_convs.", c.geo_feat_channels, c.ks, false));
        d->heads.push_back(head("geo_decoder.", 0, 1, 0, 0));
        if (c.use_tex) {
            d->br[1].blocks.push_back(blk("tex_convs.", c.tex_feat_channels, c.ks, false));
            d->heads.push_back(head("tex_decoder.", 1, c.tex_channels, 1, 1));
        }
        d->out_channels = 1 + (c.use_tex ? c.tex_channels : 0);
    }
}

// state_dict() order of the reference module (networks.py:134-162 / 227-262): encoders, then per branch its blocks followed by its heads
void build_structure(s3d_decoder* d) {
    const auto& c = d->cfg;
    describe(d);
    add_tensor(d, "aabb", {6}, false);
    add_tensor(d, "geo_encoder.weight", {c.geo_feat_channels, 1, 4, 4, 4}, false);
    add_tensor(d, "geo_encoder.bias", {c.geo_feat_channels}, false);
    if (c.use_tex) {
        add_tensor(d, "tex_encoder.weight", {c.tex_feat_channels, c.tex_channels + 1, 4, 4, 4}, false);
        add_tensor(d, "tex_encoder.bias", {c.tex_feat_channels}, false);
    }
    const int up = c.feat_channel_up;
    for (int b = 0; b < d->nb; ++b) {
        for (const ConvBlk& K : d->br[b].blocks) {
            const std::string& p = K.prefix;
            // in_layers = Sequential(conv) without input activation, Sequential(SiLU, conv) with it (blocks.py:199-216)
            const std::string in_conv = p + (K.in_norm ? "in_layers.1" : "in_layers.0");
            add_tensor(d, in_conv + ".weight", {3 * up, K.cin, K.ks, K.ks});
            add_tensor(d, in_conv + ".bias", {3 * up});
            for (int pl = 0; pl < 3; ++pl) {
                add_tensor(d, p + "norm_" + kPlaneName[pl] + ".weight", {up});
                add_tensor(d, p + "norm_" + kPlaneName[pl] + ".bias", {up});
            }
            add_tensor(d, p + "out_layers.1.weight", {3 * up, up, K.ks, K.ks});
            add_tensor(d, p + "out_layers.1.bias", {3 * up});
            if (!K.in_norm) {
                add_tensor(d, p + "shortcut.weight", {3 * up, K.cin, 1, 1});
                add_tensor(d, p + "shortcut.bias", {3 * up});
            }
        }
        for (const Head& Hd : d->heads) {
            if (Hd.feat != b) continue;
            int nf;
            for (const auto& l : mlp_layers(c, Hd.prefix, Hd.n_out, &nf)) {
                add_tensor(d, l.key + ".weight", {l.cout, l.cin});
                add_tensor(d, l.key + ".bias", {l.cout});
            }
        }
    }
}

const DTensor& T_(const s3d_decoder* d, const std::string& n) {
    auto it = d->index.find(n);
    if (it == d->index.end()) throw S3dError{"internal: unknown tensor " + n};
    const DTensor& t = d->tensors[it->second];
    if (!t.loaded) throw S3dError{"checkpoint tensor not loaded: " + n};
    return t;
}

template <typename T>
T* dev_alloc(std::vector<void*>& owner, size_t n) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    owner.push_back(p);
    return static_cast<T*>(p);
}
template <typename T>
T* dev_upload(std::vector<void*>& owner, const std::vector<T>& h) {
    T* p = dev_alloc<T>(owner, h.size());
    CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}
uint16_t f2h_bits(float f) {
    __half h = __float2half_rn(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
float h2f(uint16_t b) {
    __half h;
    memcpy(&h, &b, 2);
    return __half2float(h);
}

// grouped conv weight (3*up, cin, ks, ks) -> [plane][tap][ci][co]
std::vector<float> pack_gconv(const std::vector<float>& w, int up, int cin, int ks) {
    std::vector<float> o(static_cast<size_t>(3) * ks * ks * cin * up);
    for (int p = 0; p < 3; ++p)
        for (int co = 0; co < up; ++co)
            for (int ci = 0; ci < cin; ++ci)
                for (int t = 0; t < ks * ks; ++t)
                    o[((static_cast<size_t>(p) * ks * ks + t) * cin + ci) * up + co] =
                        w[((static_cast<size_t>(p) * up + co) * cin + ci) * ks * ks + t];
    return o;
}

void finalize(s3d_decoder* d) {
    const auto& c = d->cfg;
    const int up = c.feat_channel_up, hid = c.mlp_hidden_channels;
    for (const auto& t : d->tensors)
        if (t.needed && !t.loaded) throw S3dError{"checkpoint tensor not loaded: " + t.name};
    for (void* p : d->owned) cudaFree(p);
    d->owned.clear();
    for (auto& p : d->enc_w) p = nullptr;
    d->tc_ok = up == kDecUp && hid == kDecHid && c.mlp_hidden_layers == 4;
    for (const Head& Hd : d->heads) d->tc_ok = d->tc_ok && Hd.n_out <= 4;
    for (int b = 0; b < d->nb; ++b) {
        for (ConvBlk& K : d->br[b].blocks) {
            const std::string& p = K.prefix;
            const std::string in_conv = p + (K.in_norm ? "in_layers.1" : "in_layers.0");
            K.w_in = dev_upload(d->owned, pack_gconv(T_(d, in_conv + ".weight").host, up, K.cin, K.ks));
            K.b_in = dev_upload(d->owned, T_(d, in_conv + ".bias").host);
            K.w_out = dev_upload(d->owned, pack_gconv(T_(d, p + "out_layers.1.weight").host, up, up, K.ks));
            K.b_out = dev_upload(d->owned, T_(d, p + "out_layers.1.bias").host);
            if (!K.in_norm) {
                K.ws = dev_upload(d->owned, pack_gconv(T_(d, p + "shortcut.weight").host, up, K.cin, 1));
                K.bs = dev_upload(d->owned, T_(d, p + "shortcut.bias").host);
            }
            std::vector<float> g(3 * up), be(3 * up);
            for (int pl = 0; pl < 3; ++pl) {
                const auto& gw = T_(d, p + "norm_" + kPlaneName[pl] + ".weight").host;
                const auto& gb = T_(d, p + "norm_" + kPlaneName[pl] + ".bias").host;
                std::copy(gw.begin(), gw.end(), g.begin() + pl * up);
                std::copy(gb.begin(), gb.end(), be.begin() + pl * up);
            }
            K.gamma = dev_upload(d->owned, g);
            K.beta = dev_upload(d->owned, be);
        }
    }
    for (Head& B : d->heads) {
        int nf;
        const auto layers = mlp_layers(c, B.prefix, B.n_out, &nf);
        S3D_CHECK(layers.size() <= 8, "too many MLP layers");
        B.mf = DecMlpF32{};
        B.mf.n_first = nf;
        B.mf.n_layers = static_cast<int>(layers.size());
        for (size_t l = 0; l < layers.size(); ++l) {
            const auto& w = T_(d, layers[l].key + ".weight").host;       // [cout][cin]
            const int K = layers[l].cin, N = layers[l].cout;
            std::vector<float> wt(static_cast<size_t>(K) * N);
            for (int nn = 0; nn < N; ++nn)
                for (int k = 0; k < K; ++k) wt[static_cast<size_t>(k) * N + nn] = w[static_cast<size_t>(nn) * K + k];
            B.mf.wt[l] = dev_upload(d->owned, wt);
            B.mf.b[l] = dev_upload(d->owned, T_(d, layers[l].key + ".bias").host);
            B.mf.K[l] = K;
            B.mf.N[l] = N;
        }
        if (d->tc_ok) {
            for (int l = 0; l < kDecLayers; ++l) {
                const auto& w = T_(d, layers[l].key + ".weight").host;
                const int K = layers[l].cin;
                float mx = 0.f;
                for (float v : w) mx = std::max(mx, std::fabs(v));
                // power-of-two scale: the largest weight lands in [2048, 4096), so the (unscaled) lo halves of all but
                // vanishing weights stay in fp16 normal range
                const int e = mx > 0.f ? static_cast<int>(std::floor(std::log2(4096.0 / mx))) : 0;
                const float sc = std::ldexp(1.f, std::min(std::max(e, -20), 30));
                std::vector<uint16_t> w16(static_cast<size_t>(2) * kDecHid * K);
                for (int nn = 0; nn < kDecHid; ++nn)
                    for (int k = 0; k < K; ++k) {
                        const float v = std::min(std::max(w[static_cast<size_t>(nn) * K + k] * sc, -65504.f), 65504.f);
                        const uint16_t hi = f2h_bits(v);
                        w16[static_cast<size_t>(nn) * K + k] = hi;
                        w16[(static_cast<size_t>(kDecHid) + nn) * K + k] = f2h_bits(v - h2f(hi));
                    }
                B.w16[l] = dev_upload(d->owned, w16);
                B.inv_scale[l] = 1.f / sc;
                B.bias_h[l] = T_(d, layers[l].key + ".bias").host;
                S3D_CHECK(B.bias_h[l].size() == static_cast<size_t>(kDecHid), "hidden bias size");
                const uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(kDecHid), 2};
                const uint32_t box[3] = {64, 128, 1};
                make_tmap(&B.maps[l], B.w16[l], 3, dims, box);
            }
            B.w_last_h = T_(d, layers[kDecLayers].key + ".weight").host;      // [n_out][256]
            B.b_last_h = T_(d, layers[kDecLayers].key + ".bias").host;
            S3D_CHECK(B.n_out >= 1 && B.n_out <= 4 && B.w_last_h.size() == static_cast<size_t>(B.n_out) * kDecHid, "output layer shape");
        }
    }
    CUDA_TRY(cudaDeviceSynchronize());
    d->finalized = true;
}

void set_planes(s3d_decoder* d, const float* xy, const float* xz, const float* yz, int H, int W, int D, cudaStream_t st) {
    S3D_CHECK(d->finalized, "s3d_decoder_finalize has not run");
    S3D_CHECK(H > 0 && W > 0 && D > 0, "plane sizes must be positive");
    const int rows[3] = {H, H, W}, cols[3] = {W, D, D};
    const int CF = kDecUp * d->nb;
    bool two_blocks = false;
    for (int b = 0; b < d->nb; ++b) two_blocks = two_blocks || d->br[b].blocks.size() > 1;
    bool same = d->F[0] != nullptr;
    for (int p = 0; p < 3; ++p) same = same && rows[p] == d->rows[p] && cols[p] == d->cols[p];
    if (!same) {
        for (void* p : d->plane_owned) cudaFree(p);
        d->plane_owned.clear();
        int mx = 0;
        for (int p = 0; p < 3; ++p) {
            d->rows[p] = rows[p];
            d->cols[p] = cols[p];
            const size_t n = static_cast<size_t>(rows[p]) * cols[p];
            d->F[p] = dev_alloc<float>(d->plane_owned, n * CF);
            d->h1[p] = dev_alloc<float>(d->plane_owned, n * kDecUp);
            d->T1[p] = two_blocks ? dev_alloc<float>(d->plane_owned, n * kDecUp) : nullptr;
            mx = std::max(mx, static_cast<int>((n + kDecStripPix - 1) / kDecStripPix));
        }
        d->max_strips = mx;
        d->partial = dev_alloc<double>(d->plane_owned, static_cast<size_t>(3) * mx * 64 * 2);
        d->coef = dev_alloc<float>(d->plane_owned, 3 * 64 * 2);
    }
    const float* x[3] = {xy, xz, yz};
    const int n0 = rows[0] * cols[0], n1 = rows[1] * cols[1], n2 = rows[2] * cols[2];
    d->last_launches = 0;
    // InstanceNorm statistics of a [rows][cols][64] tensor -> per-channel (scale, shift) with the block's affine parameters
    auto in_coef = [&](float* const t[3], const ConvBlk& K) {
        launch_plain(k_dec_in_stats, dim3(d->max_strips, 3), dim3(256), 0, st, t[0], t[1], t[2], n0, n1, n2, d->partial, d->max_strips);
        launch_plain(k_dec_in_finalize, dim3(3), dim3(64), 0, st, d->partial, d->max_strips, n0, n1, n2, K.gamma, K.beta, d->coef);
        d->last_launches += 2;
    };
    for (int b = 0; b < d->nb; ++b) {
        const Branch& B = d->br[b];
        for (size_t bi = 0; bi < B.blocks.size(); ++bi) {
            const ConvBlk& K = B.blocks[bi];
            const bool last_block = bi + 1 == B.blocks.size();
            const int ks = K.ks, hw = kDecTile + ks - 1, hstride = (hw * hw) | 1;
            DecConvArgs a{};
            int ts = 0;
            for (int p = 0; p < 3; ++p) {
                a.rows[p] = rows[p];
                a.cols[p] = cols[p];
                a.tiles_x[p] = (cols[p] + kDecTile - 1) / kDecTile;
                a.tile_start[p] = ts;
                ts += a.tiles_x[p] * ((rows[p] + kDecTile - 1) / kDecTile);
                a.h1[p] = d->h1[p];
                a.F[p] = last_block ? d->F[p] : d->T1[p];          // the block's output: the branch's slice of F, or the next block's input
            }
            a.tile_start[3] = ts;
            a.CF = last_block ? CF : kDecUp;
            a.foff = last_block ? b * kDecUp : 0;
            a.ks = ks;
            const size_t sm1 = static_cast<size_t>(kDecUp) * (hstride + 64) * sizeof(float);
            DecConvArgs a0 = a;
            a0.w = K.w_in;
            a0.bias = K.b_in;
            if (!K.in_norm) {
                // conv ks x ks (c -> 64) on the raw latent + 1x1 shortcut
                S3D_CHECK(bi == 0, "a block without input norm reads the latent");
                for (int p = 0; p < 3; ++p) a0.x[p] = x[p];
                a0.c0 = B.c0;
                a0.cin = K.cin;
                a0.ws = K.ws;
                a0.bs = K.bs;
                const size_t sm0 = static_cast<size_t>(K.cin) * (hstride + 64) * sizeof(float);
                CUDA_TRY(cudaFuncSetAttribute(k_dec_conv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm0)));
                launch_plain(k_dec_conv<0>, dim3(ts), dim3(256), sm0, st, a0);
            } else {
                // input InstanceNorm (the block's own norm_* modules, blocks.py:233-234) + SiLU + conv; identity shortcut on the normed input
                S3D_CHECK(bi > 0 && K.cin == kDecUp, "a block with input norm follows another block");
                in_coef(d->T1, K);
                for (int p = 0; p < 3; ++p) a0.x[p] = d->T1[p];
                a0.cin = kDecUp;
                a0.coef = d->coef;
                CUDA_TRY(cudaFuncSetAttribute(k_dec_conv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm1)));
                launch_plain(k_dec_conv<2>, dim3(ts), dim3(256), sm1, st, a0);
            }
            // norm + SiLU + conv ks x ks (64 -> 64), added onto the shortcut
            in_coef(d->h1, K);
            DecConvArgs a1 = a;
            for (int p = 0; p < 3; ++p) a1.x[p] = d->h1[p];
            a1.cin = kDecUp;
            a1.w = K.w_out;
            a1.bias = K.b_out;
            a1.coef = d->coef;
            CUDA_TRY(cudaFuncSetAttribute(k_dec_conv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm1)));
            launch_plain(k_dec_conv<1>, dim3(ts), dim3(256), sm1, st, a1);
            d->last_launches += 2;
        }
    }
    d->planes_set = true;
}

void decode(s3d_decoder* d, const DecPoints& P, int clamp_tex, float* out, cudaStream_t st) {
    S3D_CHECK(d->planes_set, "s3d_decoder_set_planes has not run");
    if (P.n <= 0) return;
    DecArgs a{};
    for (int p = 0; p < 3; ++p) {
        a.G.F[p] = d->F[p];
        a.G.rows[p] = d->rows[p];
        a.G.cols[p] = d->cols[p];
    }
    a.G.CF = kDecUp * d->nb;
    a.P = P;
    a.oc = d->out_channels;
    a.tex_channels = d->cfg.tex_channels;
    a.clamp_tex = clamp_tex;
    a.out = out;
    if (d->cfg.mlp_impl == 0)
        S3D_CHECK(d->tc_ok, "the tensor-core decoder is specialised for feat_channel_up=64, hidden_dim=256, n_hidden_layers=4 "
                            "(the reference defaults); use mlp_impl=1 for other shapes");
    d->last_launches = 0;
    // the heads two at a time: (geo, tex) for AutoEncoderGroupSkip / V3, (geo, rgb) then (mr, normal) for AutoEncoderGroupPBR
    for (size_t h0 = 0; h0 < d->heads.size(); h0 += 2) {
        const int nh = static_cast<int>(std::min<size_t>(2, d->heads.size() - h0));
        const Head* Hh[2] = {&d->heads[h0], &d->heads[h0 + nh - 1]};
        a.nb = nh;
        for (int b = 0; b < nh; ++b) {
            a.feat_off[b] = Hh[b]->feat * kDecUp;
            a.out_col[b] = Hh[b]->out_col;
            a.sigmoid[b] = Hh[b]->sigmoid;
        }
        if (d->cfg.mlp_impl == 1) {
            const int hid = d->cfg.mlp_hidden_channels;
            const size_t sm = (static_cast<size_t>(32) * kDecUp + static_cast<size_t>(64) * hid) * sizeof(float);
            CUDA_TRY(cudaFuncSetAttribute(k_dec_mlp_ffma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)));
            const long long blocks = (P.n + 31) / 32;
            S3D_CHECK(blocks < (1LL << 31), "too many points for one launch");
            launch_plain(k_dec_mlp_ffma, dim3(static_cast<unsigned>(blocks)), dim3(256), sm, st, a, Hh[0]->mf, Hh[1]->mf, hid);
            ++d->last_launches;
            continue;
        }
        DecTcMaps maps{};
        DecTcArgs ta{};
        static thread_local DecTcConst tc;          // 18.5 KB: kept off the stack
        std::memset(&tc, 0, sizeof(tc));
        ta.D = a;
        for (int b = 0; b < nh; ++b) {
            const Head& B = *Hh[b];
            for (int l = 0; l < kDecLayers; ++l) {
                maps.w[b][l] = B.maps[l];
                ta.inv_scale[b][l] = B.inv_scale[l];
                std::copy(B.bias_h[l].begin(), B.bias_h[l].end(), tc.bias[b][l]);
            }
            std::copy(B.w_last_h.begin(), B.w_last_h.end(), &tc.w_last[b][0][0]);
            std::copy(B.b_last_h.begin(), B.b_last_h.end(), tc.b_last[b]);
            ta.n_out[b] = B.n_out;
        }
        ta.skip = d->cfg.mlp_kind == 0 ? 1 : 0;
        ta.n_tiles = (P.n + kDecPts - 1) / kDecPts;
        const unsigned grid = static_cast<unsigned>(std::min<long long>(ta.n_tiles, d->n_sm));
        if (d->cfg.precision == 1) {
            using Cfg = DecTcCfg<1>;
            CUDA_TRY(cudaFuncSetAttribute(k_dec_mlp_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
            launch_plain(k_dec_mlp_tc<1>, dim3(grid), dim3(kDecTcThreads), Cfg::kSmemBytes, st, maps, ta, tc);
        } else {
            using Cfg = DecTcCfg<3>;
            CUDA_TRY(cudaFuncSetAttribute(k_dec_mlp_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
            launch_plain(k_dec_mlp_tc<3>, dim3(grid), dim3(kDecTcThreads), Cfg::kSmemBytes, st, maps, ta, tc);
        }
        ++d->last_launches;
    }
}

template <int GEO, int TEX, int CT>
void launch_encode(s3d_decoder* d, const EncArgs& a, cudaStream_t st) {
    static thread_local EncW<GEO, TEX, CT> w;        // ~9 KB of kernel parameters
    std::memset(&w, 0, sizeof(w));
    const auto& wg = T_(d, "geo_encoder.weight").host;            // [GEO][1][4][4][4]
    const auto& bg = T_(d, "geo_encoder.bias").host;
    for (int g = 0; g < GEO; ++g) {
        std::copy(wg.begin() + g * 64, wg.begin() + (g + 1) * 64, w.wg[g]);
        w.bias[g] = bg[g];
    }
    if (TEX > 0) {
        const auto& wt = T_(d, "tex_encoder.weight").host;        // [TEX][CT][4][4][4]
        const auto& bt = T_(d, "tex_encoder.bias").host;
        for (int t = 0; t < TEX; ++t) {
            for (int c = 0; c < CT; ++c) std::copy(wt.begin() + (t * CT + c) * 64, wt.begin() + (t * CT + c + 1) * 64, w.wt[t][c]);
            w.bias[GEO + t] = bt[t];
        }
    }
    constexpr int CV = CT > 0 ? CT : 1;
    const size_t smem = (static_cast<size_t>(CV) * kEncIH * kEncIW * 2 * kEncZP + static_cast<size_t>(GEO + TEX) * 8 * 32) * sizeof(float);
    CUDA_TRY(cudaFuncSetAttribute(k_enc_conv3d<GEO, TEX, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const dim3 grid((a.D + kEncTD - 1) / kEncTD, (a.W + kEncTW - 1) / kEncTW, (a.H + kEncTH - 1) / kEncTH);
    S3D_CHECK(grid.y < 65536 && grid.z < 65536, "volume too large for one launch");
    if (a.Z % 4 == 0 && (reinterpret_cast<uintptr_t>(a.vol) & 15) == 0) {
        // 16-byte global strides: the input tile is one TMA box
        CUtensorMap vmap;
        const uint64_t dims[4] = {static_cast<uint64_t>(a.Z), static_cast<uint64_t>(a.Y), static_cast<uint64_t>(a.X), static_cast<uint64_t>(CV)};
        const uint32_t box[4] = {kEncZPitch, kEncIW, kEncIH, CV};
        make_tmap_f32(&vmap, a.vol, 4, dims, box);
        const size_t smem_t = (static_cast<size_t>(CV) * kEncIH * kEncIW * kEncZPitch + static_cast<size_t>(GEO + TEX) * 8 * 32) * sizeof(float) + 16;
        CUDA_TRY(cudaFuncSetAttribute(k_enc_conv3d_tma<GEO, TEX, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_t)));
        launch_plain(k_enc_conv3d_tma<GEO, TEX, CT>, grid, dim3(128), smem_t, st, vmap, w, a);
    } else {
        launch_plain(k_enc_conv3d<GEO, TEX, CT>, grid, dim3(256), smem, st, w, a);
    }
    launch_plain(k_enc_finalize, dim3(GEO + TEX, 3), dim3(kEncFinThreads), 0, st, a);
}

// AutoEncoderGroupSkip.encode (networks.py:164-180)
void encode(s3d_decoder* d, const float* vol, int X, int Y, int Z, float* xy, float* xz, float* yz, cudaStream_t st) {
    S3D_CHECK(d->finalized, "s3d_decoder_finalize has not run");
    S3D_CHECK(X >= 2 && Y >= 2 && Z >= 2, "volume must be at least 2 voxels along every axis");
    const auto& c = d->cfg;
    const int C = c.geo_feat_channels + (c.use_tex ? c.tex_feat_channels : 0);
    EncArgs a{};
    a.vol = vol;
    a.X = X; a.Y = Y; a.Z = Z;
    a.H = (X - 2) / 2 + 1; a.W = (Y - 2) / 2 + 1; a.D = (Z - 2) / 2 + 1;       // Conv3d(k 4, stride 2, pad 1)
    const size_t n[3] = {static_cast<size_t>(a.H) * a.W, static_cast<size_t>(a.H) * a.D, static_cast<size_t>(a.W) * a.D};
    const size_t total = C * (n[0] + n[1] + n[2]);
    if (d->enc_sums_count < total) {
        if (d->enc_sums) CUDA_TRY(cudaFree(d->enc_sums));
        d->enc_sums = nullptr;
        CUDA_TRY(cudaMalloc(&d->enc_sums, total * sizeof(unsigned long long)));
        d->enc_sums_count = total;
    }
    CUDA_TRY(cudaMemsetAsync(d->enc_sums, 0, total * sizeof(unsigned long long), st));
    a.sums[0] = d->enc_sums;
    a.sums[1] = a.sums[0] + C * n[0];
    a.sums[2] = a.sums[1] + C * n[1];
    a.out[0] = xy; a.out[1] = xz; a.out[2] = yz;
    if (c.geo_feat_channels == 4 && c.use_tex && c.tex_feat_channels == 8 && c.tex_channels == 3) launch_encode<4, 8, 4>(d, a, st);
    else if (c.geo_feat_channels == 4 && !c.use_tex) launch_encode<4, 0, 0>(d, a, st);
    else {
        // other channel configurations: generic direct convolution (weights uploaded on first use)
        S3D_CHECK(C <= kEncMaxC, "the encoder supports at most 32 latent channels");
        S3D_CHECK(a.W < 65536 && a.H < 65536, "volume too large for one launch");
        if (!d->enc_w[0]) {
            d->enc_w[0] = dev_upload(d->owned, T_(d, "geo_encoder.weight").host);
            d->enc_w[1] = dev_upload(d->owned, T_(d, "geo_encoder.bias").host);
            if (c.use_tex) {
                d->enc_w[2] = dev_upload(d->owned, T_(d, "tex_encoder.weight").host);
                d->enc_w[3] = dev_upload(d->owned, T_(d, "tex_encoder.bias").host);
            }
        }
        EncGenericW gw{};
        gw.wg = d->enc_w[0]; gw.bg = d->enc_w[1]; gw.wt = d->enc_w[2]; gw.bt = d->enc_w[3];
        gw.GEO = c.geo_feat_channels;
        gw.TEX = c.use_tex ? c.tex_feat_channels : 0;
        gw.CT = c.use_tex ? c.tex_channels + 1 : 0;
        launch_plain(k_enc_conv3d_generic, dim3((a.D + 127) / 128, a.W, a.H), dim3(128), 0, st, gw, a);
        launch_plain(k_enc_finalize, dim3(C, 3), dim3(kEncFinThreads), 0, st, a);
    }
    d->last_launches = 2;
}

void fill_aabb(DecPoints& P, const float* aabb) {
    for (int i = 0; i < 3; ++i) {
        P.amin[i] = aabb[i];
        P.asize[i] = aabb[3 + i] - aabb[i];      // fp32 subtraction, as aabb[3:] - aabb[:3] (networks.py:196)
    }
}

}  // namespace

#define DEC_API_BEGIN try {
#define DEC_API_END                   \
    }                                 \
    catch (const S3dError& e) {       \
        return fail(e.msg);           \
    }                                 \
    catch (const std::exception& e) { \
        return fail(e.what());        \
    }                                 \
    return 0;

extern "C" {

int s3d_decoder_create(const s3d_decoder_config* cfg, int device, s3d_decoder** out) {
    DEC_API_BEGIN
    S3D_CHECK(cfg && out, "null argument");
    S3D_CHECK(cfg->feat_channel_up == kDecUp, "feat_channel_up must be 64 (the reference default)");
    S3D_CHECK(cfg->geo_feat_channels >= 1 && cfg->geo_feat_channels <= 32, "geo_feat_channels out of range");
    S3D_CHECK(!cfg->use_tex || (cfg->tex_feat_channels >= 1 && cfg->tex_feat_channels <= 32), "tex_feat_channels out of range");
    S3D_CHECK(cfg->ks >= 1 && cfg->ks <= kDecMaxKs && (cfg->ks & 1), "ks must be odd and <= 7");
    S3D_CHECK(cfg->mlp_hidden_channels >= 32 && cfg->mlp_hidden_channels <= 256 && cfg->mlp_hidden_channels % 32 == 0,
              "mlp_hidden_channels must be a multiple of 32 in [32, 256]");
    S3D_CHECK(cfg->mlp_hidden_layers >= 2 && cfg->mlp_hidden_layers % 2 == 0 && cfg->mlp_hidden_layers <= 8,
              "mlp_hidden_layers must be even, 2..8");
    S3D_CHECK(cfg->tex_channels >= 1 && cfg->tex_channels <= 8, "tex_channels out of range");
    S3D_CHECK(cfg->precision == 1 || cfg->precision == 3, "precision must be 1 or 3");
    S3D_CHECK(cfg->mlp_impl == 0 || cfg->mlp_impl == 1, "mlp_impl must be 0 or 1");
    S3D_CHECK(cfg->mlp_kind == 0 || cfg->mlp_kind == 1, "mlp_kind must be 0 or 1");
    S3D_CHECK(cfg->net_kind == 0 || cfg->net_kind == 1, "net_kind must be 0 or 1");
    S3D_CHECK(cfg->net_kind == 0 || (cfg->mlp_kind == 0 && (!cfg->use_tex || cfg->tex_channels == 8)),
              "AutoEncoderGroupPBR has skip-concat heads and 8 texture channels (rgb 3, metallic-roughness 2, normal 3)");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    S3D_CHECK(device >= 0 && device < ndev, "no such CUDA device");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    S3D_CHECK(prop.major == 10, "libsin3dm_b200 is built for sm_100a only (no fallback)");
    CUDA_TRY(cudaSetDevice(device));
    auto* d = new s3d_decoder();
    d->cfg = *cfg;
    d->device = device;
    d->nb = cfg->use_tex ? 2 : 1;
    d->n_sm = prop.multiProcessorCount;
    build_structure(d);
    *out = d;
    DEC_API_END
}

int s3d_decoder_destroy(s3d_decoder* d) {
    if (!d) return 0;
    cudaSetDevice(d->device);
    for (void* p : d->owned) cudaFree(p);
    for (void* p : d->plane_owned) cudaFree(p);
    if (d->enc_sums) cudaFree(d->enc_sums);
    delete d;
    return 0;
}

int s3d_decoder_num_tensors(const s3d_decoder* d) { return d ? static_cast<int>(d->tensors.size()) : 0; }

int s3d_decoder_tensor_info(const s3d_decoder* d, int index, const char** name, int* ndim, int64_t shape[5]) {
    DEC_API_BEGIN
    S3D_CHECK(d && index >= 0 && index < static_cast<int>(d->tensors.size()), "tensor index out of range");
    const DTensor& t = d->tensors[index];
    if (name) *name = t.name.c_str();
    if (ndim) *ndim = static_cast<int>(t.shape.size());
    if (shape)
        for (size_t i = 0; i < t.shape.size(); ++i) shape[i] = t.shape[i];
    DEC_API_END
}

int s3d_decoder_load_tensor(s3d_decoder* d, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    DEC_API_BEGIN
    S3D_CHECK(d && name && host_data && shape, "null argument");
    auto it = d->index.find(name);
    if (it == d->index.end()) throw S3dError{std::string("unexpected checkpoint tensor: ") + name};
    DTensor& t = d->tensors[it->second];
    bool ok = ndim == static_cast<int>(t.shape.size());
    for (int i = 0; ok && i < ndim; ++i) ok = shape[i] == t.shape[i];
    if (!ok) throw S3dError{std::string("shape mismatch for ") + name};
    t.host.assign(host_data, host_data + t.numel());
    t.loaded = true;
    d->finalized = false;
    DEC_API_END
}

int s3d_decoder_finalize(s3d_decoder* d) {
    DEC_API_BEGIN
    S3D_CHECK(d, "null handle");
    CUDA_TRY(cudaSetDevice(d->device));
    finalize(d);
    DEC_API_END
}

int s3d_decoder_set_planes(s3d_decoder* d, const float* xy_dev, const float* xz_dev, const float* yz_dev, int H, int W, int D,
                           void* stream) {
    DEC_API_BEGIN
    S3D_CHECK(d && xy_dev && xz_dev && yz_dev, "null argument");
    CUDA_TRY(cudaSetDevice(d->device));
    set_planes(d, xy_dev, xz_dev, yz_dev, H, W, D, static_cast<cudaStream_t>(stream));
    DEC_API_END
}

int s3d_decoder_decode(s3d_decoder* d, const float* pts_dev, int64_t n, const float aabb[6], int clamp_tex, float* out_dev,
                       void* stream) {
    DEC_API_BEGIN
    S3D_CHECK(d && aabb && (n == 0 || (pts_dev && out_dev)), "null argument");
    CUDA_TRY(cudaSetDevice(d->device));
    DecPoints P{};
    P.pts = pts_dev;
    P.n = n;
    P.ny = P.nz = 1;
    fill_aabb(P, aabb);
    decode(d, P, clamp_tex, out_dev, static_cast<cudaStream_t>(stream));
    DEC_API_END
}

int s3d_decoder_decode_grid(s3d_decoder* d, const float* xs_dev, const float* ys_dev, const float* zs_dev, int nx, int ny, int nz,
                            const float aabb[6], int clamp_tex, float* out_dev, void* stream) {
    DEC_API_BEGIN
    S3D_CHECK(d && aabb && xs_dev && ys_dev && zs_dev && out_dev, "null argument");
    S3D_CHECK(nx > 0 && ny > 0 && nz > 0, "grid sizes must be positive");
    CUDA_TRY(cudaSetDevice(d->device));
    DecPoints P{};
    P.pts = nullptr;
    P.xs = xs_dev;
    P.ys = ys_dev;
    P.zs = zs_dev;
    P.ny = ny;
    P.nz = nz;
    P.n = static_cast<long long>(nx) * ny * nz;
    fill_aabb(P, aabb);
    decode(d, P, clamp_tex, out_dev, static_cast<cudaStream_t>(stream));
    DEC_API_END
}

int s3d_decoder_encode(s3d_decoder* d, const float* vol_dev, int X, int Y, int Z, float* xy_dev, float* xz_dev, float* yz_dev,
                       void* stream) {
    DEC_API_BEGIN
    S3D_CHECK(d && vol_dev && xy_dev && xz_dev && yz_dev, "null argument");
    CUDA_TRY(cudaSetDevice(d->device));
    encode(d, vol_dev, X, Y, Z, xy_dev, xz_dev, yz_dev, static_cast<cudaStream_t>(stream));
    DEC_API_END
}

int s3d_decoder_last_launches(const s3d_decoder* d) { return d ? d->last_launches : 0; }

int s3d_decoder_planes_read(s3d_decoder* d, int plane, float* host_out, int64_t n_floats) {
    DEC_API_BEGIN
    S3D_CHECK(d && d->planes_set && plane >= 0 && plane < 3 && host_out, "bad argument");
    CUDA_TRY(cudaSetDevice(d->device));
    const int64_t n = static_cast<int64_t>(d->rows[plane]) * d->cols[plane] * kDecUp * d->nb;
    S3D_CHECK(n_floats == n, "buffer size mismatch");
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(host_out, d->F[plane], n * sizeof(float), cudaMemcpyDeviceToHost));
    DEC_API_END
}

}  // extern "C"
