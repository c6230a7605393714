// Encoder object + C ABI (include/samvit_b200.h).  Host-side orchestration of the forward path
// ImageEncoderViT.forward (image_encoder.py:107-120): PatchEmbed -> +pos_embed -> depth x Block -> SimpleFPN.
#include "../../include/samvit_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace svb {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

enum ParamKind { P_VEC, P_GEMM_W, P_CONVT_W, P_CONVT_B, P_CONV22_W, P_REL_H, P_REL_W, P_IGNORED };

struct Param {
    ParamKind kind = P_VEC;
    int64_t numel = 0;
    int a = 0, b = 0;           // kind-specific dims: GEMM_W (N,K); CONVT_W (Cin,Cout); CONV22_W (Cin,Cout); CONVT_B (C)
    float* f32 = nullptr;       // packed fp32 copy (all kinds)
    bf16* b16 = nullptr;        // packed bf16 copy (GEMM operands)
    float* fold_c = nullptr;    // LayerNorm-fold column sums / folded bias (qkv and lin1 weights, bf16 path): see Epilogue::ln_stats
    float* fold_b = nullptr;
    bool loaded = false;
};

struct Arena {
    char* base = nullptr;
    size_t off = 0;
    void* alloc(size_t bytes) {
        off = (off + 1023) & ~size_t(1023);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
};

}  // namespace
}  // namespace svb

using namespace svb;

struct svb_encoder {
    svb_config_t cfg;
    std::map<std::string, Param> params;
    int D, depth, heads, hd, grid, T, mlp, d4, d8, d32;
    bool taps_enabled = false;
    float* taps = nullptr;          // [(depth+1)][T*D] fp32, first image of the last chunk
    int attn_impl_bf16 = 0;         // 0: SIMT kernel, 1: tcgen05 kernel
    bool ln_fold = false;           // bf16 path: norm1 / norm2 folded into the qkv / lin1 GEMMs (no LayerNorm pass over HBM)
    bool fold_dirty = true;         // a parameter was (re)loaded since the folded weights were last derived
    bool gn_fold = false;           // bf16 path: GroupNorm(1,C) -> 1x1 conv links of the neck folded the same way (down_8, down_4, down_32)
    int grid_pad = 0;               // window-padded token grid (70 for 64 / 14)
    std::vector<bf16*> relpack;     // per block: bf16 rel-pos table block of the tcgen05 attention kernel (attention_tc.cu)
    // scope row N3: tables resized for token grids other than the trained one (dropped whenever a parameter is reloaded)
    std::map<std::pair<int, int>, float*> pos_cache;               // (gh, gw) -> bicubic pos_embed [gh*gw, D]
    std::map<std::tuple<int, int, int>, float*> rel_cache;         // (block, is_w, length) -> linear rel_pos table [length, hd]
    std::map<std::tuple<int, int, int>, bf16*> rel16_cache;        // the same tables as bf16 GEMM operands (tcgen05 path)
    void drop_resized() {
        for (auto& kv : pos_cache) cudaFree(kv.second);
        for (auto& kv : rel_cache) cudaFree(kv.second);
        for (auto& kv : rel16_cache) cudaFree(kv.second);
        pos_cache.clear();
        rel_cache.clear();
        rel16_cache.clear();
    }
    // host path resources
    struct HostPath {
        int chunk = 0, mode = -1, out_dtype = -1;
        float* xin[2] = {nullptr, nullptr};
        void* outs[2][4] = {{nullptr}};
        void* ws = nullptr;
        size_t ws_bytes = 0;
        cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
        cudaEvent_t in_done[2], comp_done[2], out_done[2];
        bool events = false;
    } hp;

    const Param& P(const std::string& k) const { return params.at(k); }
    bool is_global(int i) const {
        for (int j = 0; j < cfg.num_global; ++j)
            if (cfg.global_idx[j] == i) return true;
        return false;
    }
};

namespace {

void add_param(svb_encoder* e, const std::string& key, ParamKind kind, int64_t numel, int a = 0, int b = 0) {
    Param p;
    p.kind = kind;
    p.numel = numel;
    p.a = a;
    p.b = b;
    e->params[key] = p;
}

int alloc_param_storage(Param& p) {
    if (p.kind == P_IGNORED) return 0;
    int64_t n = p.numel;
    if (p.kind == P_CONVT_B) n = 4 * p.numel;
    SVB_CHECK_CUDA(cudaMalloc(&p.f32, sizeof(float) * n));
    if (p.kind == P_GEMM_W || p.kind == P_CONVT_W || p.kind == P_CONV22_W) SVB_CHECK_CUDA(cudaMalloc(&p.b16, sizeof(bf16) * n));
    return 0;
}

int alloc_fold_storage(Param& p) {
    SVB_CHECK_CUDA(cudaMalloc(&p.fold_c, sizeof(float) * p.a));
    SVB_CHECK_CUDA(cudaMalloc(&p.fold_b, sizeof(float) * p.a));
    return 0;
}

struct Buffers {
    void *A0, *Xn, *QKV, *O, *Hid;
    float* X;
    // neck
    void *Xb, *A32, *Gn, *G2n;
    float *G, *G2, *G3;
    double* stats;
    float2 *st1, *st2;          // LayerNorm-fold partial row statistics of the residual stream (norm1 / norm2 inputs)
    float *c1, *c2;             // their per-row shifts (centred hand-over, Epilogue::shift_out)
    float *biasH, *biasW;       // row N3: rel-pos term tables of the global blocks on other token grids
    size_t total;
};

// gh x gw = token grid of the canvas; 0 = the trained grid
Buffers plan(const svb_encoder* e, int chunk, int mode, void* base, int gh = 0, int gw = 0) {
    const size_t es = (mode == SVB_MODE_BF16) ? 2 : 4;
    if (gh == 0) { gh = e->grid; gw = e->grid; }
    const bool native = (gh == e->grid && gw == e->grid);
    const int T = gh * gw;
    const size_t M = (size_t)chunk * T;
    const int D = e->D;
    const int kpe = e->cfg.in_chans * e->cfg.patch_size * e->cfg.patch_size;
    Buffers b;
    Arena ar;
    ar.base = (char*)base;
    b.X = (float*)ar.alloc(M * D * 4);
    b.stats = (double*)ar.alloc(sizeof(double) * 2 * 8 * chunk);
    const size_t parts = (size_t)(D + 127) / 128;
    b.st1 = (float2*)ar.alloc(M * parts * sizeof(float2));
    b.st2 = (float2*)ar.alloc(M * parts * sizeof(float2));
    b.c1 = (float*)ar.alloc(M * sizeof(float));
    b.c2 = (float*)ar.alloc(M * sizeof(float));
    const size_t mark = ar.off;
    b.A0 = ar.alloc(M * kpe * es);
    b.Xn = ar.alloc(M * D * es);
    const int ws_ = e->cfg.window_size;
    const size_t Mp = (mode == SVB_MODE_BF16 && e->attn_impl_bf16 == 1)
                          ? (size_t)chunk * (((gh + ws_ - 1) / ws_) * ws_) * (((gw + ws_ - 1) / ws_) * ws_) : M;   // window-padded grid
    b.QKV = ar.alloc(Mp * 3 * D * es);
    // other token grids on the tcgen05 path (row N3): the rel-pos terms of the global blocks, [M][heads][2 g - 1 rounded up to 8] fp32
    b.biasH = b.biasW = nullptr;
    if (mode == SVB_MODE_BF16 && e->attn_impl_bf16 == 1 && !native) {
        b.biasH = (float*)ar.alloc(M * e->heads * (size_t)((2 * gh - 1 + 7) / 8 * 8) * 4);
        b.biasW = (float*)ar.alloc(M * e->heads * (size_t)((2 * gw - 1 + 7) / 8 * 8) * 4);
    }
    b.O = ar.alloc(M * D * es);
    b.Hid = ar.alloc(M * e->mlp * es);
    const size_t trunk_end = ar.off;
    ar.off = mark;   // the neck reuses the trunk's activation buffers
    b.Xb = ar.alloc(M * D * es);
    b.A32 = ar.alloc(M * D * es);
    const size_t dmax = (size_t)(e->d4 > e->d8 ? e->d4 : e->d8);
    size_t g_elems = M * 4 * dmax;                       // ConvT output [M, 4*d]
    if (g_elems < M * (size_t)e->cfg.fpn_dims[2]) g_elems = M * (size_t)e->cfg.fpn_dims[2];
    if (g_elems < (M / 4) * (size_t)e->d32) g_elems = (M / 4) * (size_t)e->d32;
    b.G = (float*)ar.alloc(g_elems * 4);
    b.Gn = ar.alloc(g_elems * es);
    size_t g2_elems = 4 * M * 4 * (size_t)(e->d4 / 2);    // second ConvT output [4M, 4*d4/2]
    if (g2_elems < 4 * M * (size_t)e->cfg.fpn_dims[1]) g2_elems = 4 * M * (size_t)e->cfg.fpn_dims[1];
    if (g2_elems < (M / 4) * (size_t)e->cfg.fpn_dims[3]) g2_elems = (M / 4) * (size_t)e->cfg.fpn_dims[3];
    b.G2 = (float*)ar.alloc(g2_elems * 4);
    b.G2n = ar.alloc(g2_elems * es);
    b.G3 = (float*)ar.alloc(16 * M * (size_t)e->cfg.fpn_dims[0] * 4);
    b.total = (ar.off > trunk_end ? ar.off : trunk_end) + 1024;
    return b;
}

int linear(int mode, const void* A, int lda, const Param& W, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
    if (mode == SVB_MODE_BF16) return gemm_bf16_tc((const bf16*)A, lda, W.b16, K, M, N, K, ep, st);
    return gemm_f32_simt((const float*)A, lda, W.f32, K, M, N, K, ep, st);
}

// raw uint8 input of a chunk (scope row N2): per image a device pointer to (C,h,w) uint8 and its size; normalisation constants
struct U8Input {
    const uint8_t* const* images = nullptr;
    const int* hs = nullptr;
    const int* ws = nullptr;
    const float* mean = nullptr;
    const float* stdv = nullptr;
};

// resized tables of scope row N3 (cached per encoder)
int resized_pos(svb_encoder* e, int gh, int gw, cudaStream_t st, const float** out) {
    auto it = e->pos_cache.find({gh, gw});
    if (it == e->pos_cache.end()) {
        float* p = nullptr;
        SVB_CHECK_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)gh * gw * e->D));
        int rc = resize_pos_embed(e->P("pos_embed").f32, p, e->grid, e->grid, gh, gw, e->D, st);
        if (rc) { cudaFree(p); return rc; }
        it = e->pos_cache.emplace(std::make_pair(gh, gw), p).first;
    }
    *out = it->second;
    return 0;
}
int resized_rel(svb_encoder* e, int block, bool is_w, int L, cudaStream_t st, const float** out) {
    const Param& t = e->P("blocks." + std::to_string(block) + (is_w ? ".attn.rel_pos_w" : ".attn.rel_pos_h"));
    if (t.a == L) { *out = t.f32; return 0; }
    auto key = std::make_tuple(block, (int)is_w, L);
    auto it = e->rel_cache.find(key);
    if (it == e->rel_cache.end()) {
        float* p = nullptr;
        SVB_CHECK_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)L * e->hd));
        int rc = resize_rel_pos(t.f32, p, t.a, L, e->hd, st);
        if (rc) { cudaFree(p); return rc; }
        it = e->rel_cache.emplace(key, p).first;
    }
    *out = it->second;
    return 0;
}

// the same table as a bf16 GEMM operand [L][hd] (cached per encoder)
int resized_rel_bf16(svb_encoder* e, int block, bool is_w, int L, cudaStream_t st, const bf16** out) {
    auto key = std::make_tuple(block, (int)is_w, L);
    auto it = e->rel16_cache.find(key);
    if (it == e->rel16_cache.end()) {
        const float* src = nullptr;
        int rc = resized_rel(e, block, is_w, L, st, &src);
        if (rc) return rc;
        bf16* p = nullptr;
        SVB_CHECK_CUDA(cudaMalloc(&p, sizeof(bf16) * (size_t)L * e->hd));
        rc = add_cast(src, nullptr, p, true, (long)L * e->hd, st);
        if (rc) { cudaFree(p); return rc; }
        it = e->rel16_cache.emplace(key, p).first;
    }
    *out = it->second;
    return 0;
}

// One pass over B images of img_h x img_w pixels (0 = the trained img_size).  Token grids other than the trained one (scope row
// N3) take the reference's fallbacks: bicubic pos_embed (image_encoder.py:111-114,124-132), linearly resized rel_pos tables in
// the global blocks (:319-330); their attention runs on the fp32-math kernel (the tcgen05 kernels implement the 64 x 64 grid).
int forward_chunk(svb_encoder* e, const void* x, int B, void* const outs[4], int out_dtype, int mode, const Buffers& bf,
                  cudaStream_t st, const U8Input* u8 = nullptr, int img_h = 0, int img_w = 0, int x_dtype = SVB_DTYPE_F32) {
    const bool h = (mode == SVB_MODE_BF16);
    if (img_h == 0) { img_h = e->cfg.img_size; img_w = e->cfg.img_size; }
    const int gh = img_h / e->cfg.patch_size, gw = img_w / e->cfg.patch_size;
    const bool native = (gh == e->grid && gw == e->grid);
    const int D = e->D, T = gh * gw, g = e->grid;
    const int M = B * T;
    const int kpe = e->cfg.in_chans * e->cfg.patch_size * e->cfg.patch_size;
    int rc;
    // bf16 path: norm1 / norm2 are folded into the qkv / lin1 GEMMs.  Every GEMM that writes the fp32 residual stream X
    // (patch embedding, proj, lin2) also writes bf16(X) and the partial row sums of X; the consuming GEMM normalises in its
    // epilogue (Epilogue::ln_stats).  The fp32 validation mode keeps the explicit LayerNorm kernels.
    const bool fold = h && e->ln_fold;
    const int parts = (D + 127) / 128;
    if (fold && e->fold_dirty) {
        for (int i = 0; i < e->depth; ++i) {
            const std::string p = "blocks." + std::to_string(i) + ".";
            for (int k = 0; k < 2; ++k) {
                const Param& W = e->P(p + (k ? "mlp.lin1.weight" : "attn.qkv.weight"));
                const Param& bW = e->P(p + (k ? "mlp.lin1.bias" : "attn.qkv.bias"));
                const std::string nk = p + (k ? "norm2." : "norm1.");
                if ((rc = fold_layernorm(W.f32, bW.f32, e->P(nk + "weight").f32, e->P(nk + "bias").f32, W.b16, W.fold_c, W.fold_b, W.a, W.b, st)))
                    return rc;
            }
        }
        if (e->gn_fold) {
            // GroupNorm(1,C) directly followed by a 1x1 conv (image_encoder.py:430-432, 422-424, 443-445): same algebra, gamma / beta
            // are per INPUT channel of the conv, the statistics per sample
            const char* links[3][3] = {{"neck.down_8.2.", "neck.down_8.1.", nullptr}, {"neck.down_4.5.", "neck.down_4.4.", nullptr},
                                       {"neck.down_32.2.", "neck.down_32.1.", nullptr}};
            for (auto& l : links) {
                const Param& W = e->P(std::string(l[0]) + "weight");
                if ((rc = fold_layernorm(W.f32, e->P(std::string(l[0]) + "bias").f32, e->P(std::string(l[1]) + "weight").f32,
                                         e->P(std::string(l[1]) + "bias").f32, W.b16, W.fold_c, W.fold_b, W.a, W.b, st)))
                    return rc;
            }
        }
        e->fold_dirty = false;
    }
    // SVB_LN_CENTRE=0 switches the centring of the hand-over off (A/B comparisons; the shifts are then all zero)
    static const bool centre = [] { const char* v = getenv("SVB_LN_CENTRE"); return !(v && atoi(v) == 0); }();
    auto produce = [&](Epilogue& ep, float2* stat, bool first = false) {      // epilogue of a GEMM that writes the residual stream
        ep.out = bf.X; ep.ldo = D;
        if (fold) {
            ep.out2 = bf.Xn; ep.ldo2 = D; ep.stat_out = stat;
            // the statistics / bf16 copy are of x - c[row], c = the row's mean before this update (from the other statistics array)
            ep.shift_out = (stat == bf.st1) ? bf.c1 : bf.c2;
            if (!first && centre) {
                ep.shift_stats = (stat == bf.st1) ? bf.st2 : bf.st1;
                ep.shift_in = (stat == bf.st1) ? bf.c2 : bf.c1;
                ep.shift_parts = parts; ep.shift_dim = D;
            }
        }
    };
    auto consume = [&](Epilogue& ep, const Param& W, const Param& b, const float2* stat) {   // epilogue of qkv / lin1
        if (fold) {
            ep.bias = W.fold_b; ep.ln_c = W.fold_c; ep.ln_stats = stat; ep.ln_parts = parts; ep.ln_dim = D; ep.ln_eps = e->cfg.ln_eps;
        } else {
            ep.bias = b.f32;
        }
    };
    // ---- PatchEmbed (image_encoder.py:402-410) + pos_embed (:109-114), fused in the GEMM epilogue ----
    if (u8) {
        if ((rc = stage_u8_patch(u8->images, u8->hs, u8->ws, u8->mean, u8->stdv, bf.A0, h, B, e->cfg.in_chans, e->cfg.img_size,
                                 e->cfg.patch_size, st)))
            return rc;
    } else if ((rc = im2col_patch(x, x_dtype, bf.A0, h, B, e->cfg.in_chans, img_h, img_w, e->cfg.patch_size, st))) {
        return rc;
    }
    {
        const float* pos = e->P("pos_embed").f32;
        if (!native && (rc = resized_pos(e, gh, gw, st, &pos))) return rc;
        Epilogue ep;
        ep.bias = e->P("patch_embed.proj.bias").f32;
        ep.resid = pos;
        ep.resid_mod = T;
        ep.ldr = D;
        produce(ep, bf.st1, true);
        if (fold && centre) {
            // the first producer centres by the mean of what it ADDS to every row of token t: pos_embed[t, :] + the conv bias
            // (c2 is free until proj of block 0 writes it, and only its first T entries are used here)
            if ((rc = row_means(pos, e->P("patch_embed.proj.bias").f32, bf.c2, T, D, st))) return rc;
            ep.shift_in = bf.c2; ep.shift_in_mod = T;
        }
        if ((rc = linear(mode, bf.A0, kpe, e->P("patch_embed.proj.weight"), M, D, kpe, ep, st))) return rc;
    }
    const bool taps = e->taps_enabled && native;        // the tap buffer is sized for the trained grid
    if (taps) SVB_CHECK_CUDA(cudaMemcpyAsync(e->taps, bf.X, sizeof(float) * T * D, cudaMemcpyDeviceToDevice, st));

    // ---- Blocks (image_encoder.py:181-197) ----
    for (int i = 0; i < e->depth; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        if (!fold && (rc = layernorm_rows(bf.X, nullptr, e->P(p + "norm1.weight").f32, e->P(p + "norm1.bias").f32, bf.Xn, h, M, D,
                                          e->cfg.ln_eps, st)))
            return rc;
        const bool glob = e->is_global(i);
        const int ws = glob ? g : e->cfg.window_size;
        // SVB_ATTN_EXT=0: other token grids on the fp32-math kernel as in round 1 (A/B, bisecting)
        static const bool ext_off = [] { const char* v = getenv("SVB_ATTN_EXT"); return v && atoi(v) == 0; }();
        const bool tc = h && e->attn_impl_bf16 == 1 && (native || !ext_off);
        const bool padded = tc && !glob;        // windowed blocks of the tcgen05 path keep qkv on the window-padded grid (70 x 70)
        const int wsz = e->cfg.window_size;
        const int gph = ((gh + wsz - 1) / wsz) * wsz, gpw = ((gw + wsz - 1) / wsz) * wsz;
        {   // qkv (image_encoder.py:242)
            Epilogue ep;
            consume(ep, e->P(p + "attn.qkv.weight"), e->P(p + "attn.qkv.bias"), bf.st1);
            ep.out = bf.QKV;
            ep.out_bf16 = h;
            ep.ldo = 3 * D;
            // pad tokens are zero after norm1 (image_encoder.py:183-187,271-275): their qkv rows are the (unfolded) bias, written by
            // otherwise idle warps of the same GEMM
            // SVB_PAD_IN_GEMM=1: idle warps of the qkv GEMM write the pad rows instead of a separate launch — measured neutral
            // (same-box A/B, 16 images: 170.6 / 167.2 vs 171.6 / 169.7 images/s), so the separate 18 us launch stays the default
            static const bool pad_in_gemm = [] { const char* v = getenv("SVB_PAD_IN_GEMM"); return v && atoi(v) == 1; }();
            if (padded) {
                ep.remap_g = gw; ep.remap_gp = gpw;
                if (!native) { ep.remap_h = gh; ep.remap_hp = gph; }
                else if (pad_in_gemm) ep.pad_bias = e->P(p + "attn.qkv.bias").f32;
            }
            if ((rc = linear(mode, bf.Xn, D, e->P(p + "attn.qkv.weight"), M, 3 * D, D, ep, st))) return rc;
            if (padded && !(native && pad_in_gemm) &&
                (rc = fill_pad_rows((bf16*)bf.QKV, e->P(p + "attn.qkv.bias").f32, B, gh, gw, gph, gpw, 3 * D, st))) return rc;
        }
        {   // windowed / global attention with decomposed rel-pos (image_encoder.py:246-253, 258-304, 340-376)
            if (tc) {
                AttnTcParams ap;
                ap.qkv = (const bf16*)bf.QKV; ap.out = (bf16*)bf.O;
                ap.rel_pack = e->relpack[i];
                ap.batch = B; ap.grid = g; ap.ws = ws; ap.heads = e->heads; ap.hd = e->hd;
                if (!native) { ap.grid_h = gh; ap.grid_w = gw; }
                if (native || !glob) {
                    if ((rc = attention_tc(ap, st))) return rc;
                } else {
                    // global block on another token grid (row N3): the decomposed rel-pos terms q . rel_pos'[j] of every token and
                    // head from the linearly resized tables (get_rel_pos, image_encoder.py:319-330) as tcgen05 GEMMs — per head
                    // [M, hd] x [2 g - 1, hd]^T with fp32 output rows [M][heads][ld] — then the generic-grid attention kernel
                    const int Lh = 2 * gh - 1, Lw = 2 * gw - 1, ldh = (Lh + 7) / 8 * 8, ldw = (Lw + 7) / 8 * 8;
                    for (int w = 0; w < 2; ++w) {
                        const bf16* tab = nullptr;
                        if ((rc = resized_rel_bf16(e, i, w == 1, w ? Lw : Lh, st, &tab))) return rc;
                        const int L = w ? Lw : Lh, ld = w ? ldw : ldh;
                        float* dst = w ? bf.biasW : bf.biasH;
                        for (int hh = 0; hh < e->heads; ++hh) {
                            Epilogue ep;
                            ep.out = dst + (size_t)hh * ld; ep.ldo = e->heads * ld;
                            if ((rc = gemm_bf16_tc((const bf16*)bf.QKV + (size_t)hh * e->hd, 3 * D, tab, e->hd, M, L, e->hd, ep, st))) return rc;
                        }
                    }
                    ap.bias_h = bf.biasH; ap.bias_w = bf.biasW; ap.bias_ld_h = ldh; ap.bias_ld_w = ldw;
                    if ((rc = attention_global_ext(ap, st))) return rc;
                }
            } else {
                AttnParams ap;
                ap.qkv = bf.QKV; ap.out = bf.O;
                ap.rel_h = e->P(p + "attn.rel_pos_h").f32;
                ap.rel_w = e->P(p + "attn.rel_pos_w").f32;
                ap.qkv_bias = e->P(p + "attn.qkv.bias").f32;
                ap.batch = B; ap.grid = g; ap.ws = ws; ap.heads = e->heads; ap.hd = e->hd;
                if (!native) {
                    ap.grid_h = gh; ap.grid_w = gw;
                    ap.ws_h = glob ? gh : e->cfg.window_size;
                    ap.ws_w = glob ? gw : e->cfg.window_size;
                    if (glob) {     // get_rel_pos resizes the table when its length is not 2 * size - 1 (image_encoder.py:319-330)
                        if ((rc = resized_rel(e, i, false, 2 * gh - 1, st, &ap.rel_h))) return rc;
                        if ((rc = resized_rel(e, i, true, 2 * gw - 1, st, &ap.rel_w))) return rc;
                    }
                }
                if ((rc = attention_simt(ap, h, st))) return rc;
            }
        }
        if (fold) {
            // proj (image_encoder.py:253) + shortcut add (:194) in place on X; bf16(X) and the row sums for norm2 come with it
            Epilogue ep;
            ep.bias = e->P(p + "attn.proj.bias").f32;
            ep.resid = bf.X; ep.ldr = D;
            produce(ep, bf.st2);
            if ((rc = linear(mode, bf.O, D, e->P(p + "attn.proj.weight"), M, D, D, ep, st))) return rc;
        } else {
            // proj (image_encoder.py:253).  Its output stays in the activation dtype; the shortcut add x = shortcut + x (:194) is
            // fused into the norm2 kernel below, which is HBM-bound anyway.
            Epilogue ep;
            ep.bias = e->P(p + "attn.proj.bias").f32;
            ep.out = bf.Xn; ep.out_bf16 = h; ep.ldo = D;
            if ((rc = linear(mode, bf.O, D, e->P(p + "attn.proj.weight"), M, D, D, ep, st))) return rc;
            // x += proj(attn);  O = norm2(x)   (image_encoder.py:194-195).  O is free again and receives the normalised rows.
            if ((rc = layernorm_rows(bf.X, bf.Xn, e->P(p + "norm2.weight").f32, e->P(p + "norm2.bias").f32, bf.O, h, M, D, e->cfg.ln_eps, st)))
                return rc;
        }
        {   // MLPBlock lin1 + GELU (common.py:25-26)
            Epilogue ep;
            consume(ep, e->P(p + "mlp.lin1.weight"), e->P(p + "mlp.lin1.bias"), bf.st2);
            ep.act = 1;
            ep.out = bf.Hid; ep.out_bf16 = h; ep.ldo = e->mlp;
            if ((rc = linear(mode, fold ? bf.Xn : bf.O, D, e->P(p + "mlp.lin1.weight"), M, e->mlp, D, ep, st))) return rc;
        }
        {   // lin2 + residual (image_encoder.py:195)
            Epilogue ep;
            ep.bias = e->P(p + "mlp.lin2.bias").f32;
            ep.resid = bf.X; ep.ldr = D;
            produce(ep, bf.st1);
            if ((rc = linear(mode, bf.Hid, e->mlp, e->P(p + "mlp.lin2.weight"), M, D, e->mlp, ep, st))) return rc;
        }
        if (taps)
            SVB_CHECK_CUDA(cudaMemcpyAsync(e->taps + (size_t)(i + 1) * T * D, bf.X, sizeof(float) * T * D, cudaMemcpyDeviceToDevice, st));
    }

    // ---- SimpleFPN neck (image_encoder.py:413-466).  ConvTranspose2d(k=2,s=2) and Conv2d(k=2,s=2) do not overlap, so
    // each is one GEMM; the 2x2 sub-pixel index stays folded in the row index until the final NCHW write. ----
    SVB_CHECK_CUDA(cudaMemsetAsync(bf.stats, 0, sizeof(double) * 2 * 8 * B, st));
    if ((rc = cast_and_space2depth(bf.X, h ? bf.Xb : nullptr, bf.A32, h, B, gh, gw, D, st))) return rc;
    const void* Xb = h ? bf.Xb : (const void*)bf.X;
    const float geps = e->cfg.gn_eps;
    const int* od = e->cfg.fpn_dims;
    auto S = [&](int k) { return bf.stats + (size_t)2 * B * k; };
    auto gemm_stats = [&](const void* A, int lda, const std::string& wkey, const std::string& bkey, int m, int n, int k,
                          float* out, double* stats, int rps) {
        Epilogue ep;
        ep.bias = e->P(bkey).f32;
        ep.out = out; ep.ldo = n;
        ep.stats = stats; ep.rows_per_sample = rps;
        return linear(mode, A, lda, e->P(wkey), m, n, k, ep, st);
    };
    const std::string n = "neck.";
    // down_16: Conv1x1 -> GN -> GELU  (:435-439)
    if ((rc = gemm_stats(Xb, D, n + "down_16.0.weight", n + "down_16.0.bias", M, od[2], D, bf.G, S(0), T))) return rc;
    if ((rc = groupnorm_apply_nchw(bf.G, S(0), e->P(n + "down_16.1.weight").f32, e->P(n + "down_16.1.bias").f32, outs[2], out_dtype,
                                   B, gh, gw, 0, od[2], geps, 1, st))) return rc;
    // GroupNorm -> 1x1 conv links in the folded form (bf16 path): the producer writes the RAW conv output in bf16 (+ its per-sample
    // sums), the consumer GEMM normalises in its epilogue; no GroupNorm pass, no fp32 intermediate.
    const bool gfold = fold && e->gn_fold;
    auto gemm_raw_bf16 = [&](const void* A, int lda, const std::string& wkey, const std::string& bkey, int m, int nn, int k, void* out,
                             double* stats, int rps) {
        Epilogue ep;
        ep.bias = e->P(bkey).f32;
        ep.out = out; ep.ldo = nn; ep.out_bf16 = 1;
        ep.stats = stats; ep.rows_per_sample = rps;
        return linear(mode, A, lda, e->P(wkey), m, nn, k, ep, st);
    };
    auto gemm_gnfold = [&](const void* A, int lda, const std::string& wkey, int m, int nn, int k, float* out, const double* in_stats,
                           int in_rows, double* stats, int rps) {
        const Param& W = e->P(wkey);
        Epilogue ep;
        ep.bias = W.fold_b; ep.ln_c = W.fold_c; ep.ln_eps = geps;
        ep.gn_in_stats = in_stats; ep.gn_in_rows = in_rows;
        ep.out = out; ep.ldo = nn;
        ep.stats = stats; ep.rows_per_sample = rps;
        return linear(mode, A, lda, W, m, nn, k, ep, st);
    };
    // down_8: ConvT -> GN -> Conv1x1 -> GN -> GELU  (:428-434)
    if (gfold) {
        if ((rc = gemm_raw_bf16(Xb, D, n + "down_8.0.weight", n + "down_8.0.bias", M, 4 * e->d8, D, bf.Gn, S(1), T))) return rc;
        if ((rc = gemm_gnfold(bf.Gn, e->d8, n + "down_8.2.weight", 4 * M, od[1], e->d8, bf.G2, S(1), 4 * T, S(2), 4 * T))) return rc;
    } else {
        if ((rc = gemm_stats(Xb, D, n + "down_8.0.weight", n + "down_8.0.bias", M, 4 * e->d8, D, bf.G, S(1), T))) return rc;
        if ((rc = groupnorm_apply(bf.G, S(1), e->P(n + "down_8.1.weight").f32, e->P(n + "down_8.1.bias").f32, bf.Gn, h, (long)4 * M, e->d8,
                                  (long)4 * T, geps, 0, st))) return rc;
        if ((rc = gemm_stats(bf.Gn, e->d8, n + "down_8.2.weight", n + "down_8.2.bias", 4 * M, od[1], e->d8, bf.G2, S(2), 4 * T))) return rc;
    }
    if ((rc = groupnorm_apply_nchw(bf.G2, S(2), e->P(n + "down_8.3.weight").f32, e->P(n + "down_8.3.bias").f32, outs[1], out_dtype, B,
                                   gh, gw, 1, od[1], geps, 1, st))) return rc;
    // down_4: ConvT -> GN -> GELU -> ConvT -> GN -> Conv1x1 -> GN -> GELU  (:417-426)
    if ((rc = gemm_stats(Xb, D, n + "down_4.0.weight", n + "down_4.0.bias", M, 4 * e->d4, D, bf.G, S(3), T))) return rc;
    if ((rc = groupnorm_apply(bf.G, S(3), e->P(n + "down_4.1.weight").f32, e->P(n + "down_4.1.bias").f32, bf.Gn, h, (long)4 * M, e->d4,
                              (long)4 * T, geps, 1, st))) return rc;
    if (gfold) {
        if ((rc = gemm_raw_bf16(bf.Gn, e->d4, n + "down_4.3.weight", n + "down_4.3.bias", 4 * M, 4 * (e->d4 / 2), e->d4, bf.G2n, S(4), 4 * T)))
            return rc;
        if ((rc = gemm_gnfold(bf.G2n, e->d4 / 2, n + "down_4.5.weight", 16 * M, od[0], e->d4 / 2, bf.G3, S(4), 16 * T, S(5), 16 * T))) return rc;
    } else {
        if ((rc = gemm_stats(bf.Gn, e->d4, n + "down_4.3.weight", n + "down_4.3.bias", 4 * M, 4 * (e->d4 / 2), e->d4, bf.G2, S(4), 4 * T)))
            return rc;
        if ((rc = groupnorm_apply(bf.G2, S(4), e->P(n + "down_4.4.weight").f32, e->P(n + "down_4.4.bias").f32, bf.G2n, h, (long)16 * M,
                                  e->d4 / 2, (long)16 * T, geps, 0, st))) return rc;
        if ((rc = gemm_stats(bf.G2n, e->d4 / 2, n + "down_4.5.weight", n + "down_4.5.bias", 16 * M, od[0], e->d4 / 2, bf.G3, S(5), 16 * T)))
            return rc;
    }
    if ((rc = groupnorm_apply_nchw(bf.G3, S(5), e->P(n + "down_4.6.weight").f32, e->P(n + "down_4.6.bias").f32, outs[0], out_dtype, B,
                                   gh, gw, 2, od[0], geps, 1, st))) return rc;
    // down_32: Conv(k2,s2) -> GN -> Conv1x1 -> GN -> GELU  (:441-447)
    if (gfold) {
        if ((rc = gemm_raw_bf16(bf.A32, 4 * D, n + "down_32.0.weight", n + "down_32.0.bias", M / 4, e->d32, 4 * D, bf.Gn, S(6), T / 4))) return rc;
        if ((rc = gemm_gnfold(bf.Gn, e->d32, n + "down_32.2.weight", M / 4, od[3], e->d32, bf.G2, S(6), T / 4, S(7), T / 4))) return rc;
    } else {
        if ((rc = gemm_stats(bf.A32, 4 * D, n + "down_32.0.weight", n + "down_32.0.bias", M / 4, e->d32, 4 * D, bf.G, S(6), T / 4))) return rc;
        if ((rc = groupnorm_apply(bf.G, S(6), e->P(n + "down_32.1.weight").f32, e->P(n + "down_32.1.bias").f32, bf.Gn, h, (long)M / 4, e->d32,
                                  (long)T / 4, geps, 0, st))) return rc;
        if ((rc = gemm_stats(bf.Gn, e->d32, n + "down_32.2.weight", n + "down_32.2.bias", M / 4, od[3], e->d32, bf.G2, S(7), T / 4))) return rc;
    }
    if ((rc = groupnorm_apply_nchw(bf.G2, S(7), e->P(n + "down_32.3.weight").f32, e->P(n + "down_32.3.bias").f32, outs[3], out_dtype, B,
                                   gh / 2, gw / 2, 0, od[3], geps, 1, st))) return rc;
    return 0;
}

// Split `batch` images into passes of at most `max_chunk` images.  The block GEMMs run as persistent kernels over
// ceil(B*T/256) x ceil(N/256) tiles on num_sms/2 CTA pairs, so a pass costs ceil(tiles / pairs) rounds per GEMM: e.g. for ViT-H
// 8 images fill 97.9 % of the rounds' slots, 12 images 99.8 %, 13 images 96.4 %.  Dynamic programme over the batch with that
// cost (weighted by K, the work per tile) plus a small per-pass charge.
struct SchedGeom {          // what the schedules depend on (no device state: svb_pass_schedule_model runs them on a CPU-only host)
    long D, mlp, T;
    int depth, sms;
};
int device_sms() {
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    else cudaGetLastError();
    return sms > 0 ? sms : 148;
}
SchedGeom geom_of(const svb_encoder* e) { return SchedGeom{e->D, e->mlp, e->T, e->depth, device_sms()}; }

std::vector<int> chunk_schedule(const SchedGeom& gq, int batch, int max_chunk) {
    static const int fixed = [] { const char* v = getenv("SVB_FIXED_CHUNKS"); return v ? atoi(v) : 0; }();   // 1: equal passes of max_chunk
    std::vector<int> out;
    if (fixed || max_chunk >= batch) {
        for (int b0 = 0; b0 < batch; b0 += max_chunk) out.push_back(std::min(max_chunk, batch - b0));
        return out;
    }
    const long pairs = std::max(1, gq.sms / 2);
    const long D = gq.D, mlp = gq.mlp;
    const long shp[4][2] = {{3 * D, D}, {D, D}, {mlp, D}, {D, mlp}};
    auto cost = [&](int B) {
        double c = 0;
        for (auto& nk : shp) {
            const long tiles = (((long)B * gq.T + 255) / 256) * ((nk[0] + 255) / 256);
            c += (double)nk[1] * (double)((tiles + pairs - 1) / pairs);
        }
        return c;
    };
    // per pass: ~6 launches per block with ~4 us of fill / drain each, in cost units (7.45 ns x depth x 1.4 per unit: depth cancels).  The
    // first version charged 0.002 x cost(8) and split 32 ViT-L images into 15 + 15 + 2; measured 321.6 vs 323.1 images/s for 16 + 16
    const double per_pass = 6.0 * 4e-6 / (7.45e-9 * 1.4);
    std::vector<double> best(batch + 1, 1e300);
    std::vector<int> pick(batch + 1, 0);
    best[0] = 0;
    for (int n = 1; n <= batch; ++n)
        for (int c = 1; c <= std::min(max_chunk, n); ++c) {
            const double v = best[n - c] + cost(c) + per_pass;
            if (v < best[n]) { best[n] = v; pick[n] = c; }
        }
    for (int n = batch; n > 0; n -= pick[n]) out.push_back(pick[n]);
    std::sort(out.begin(), out.end(), [](int a, int b) { return a > b; });
    return out;
}
std::vector<int> chunk_schedule(const svb_encoder* e, int batch, int max_chunk) { return chunk_schedule(geom_of(e), batch, max_chunk); }

size_t out_elems_per_image(const svb_encoder* e, int k, int img_h = 0, int img_w = 0) {
    if (img_h == 0) { img_h = e->cfg.img_size; img_w = e->cfg.img_size; }
    const int strides[4] = {4, 8, 16, 32};
    const size_t hw = (size_t)(img_h / strides[k]) * (img_w / strides[k]);
    return hw * e->cfg.fpn_dims[k];
}

// The host path (svb_encoder_forward_host) pipelines H2D -> compute -> D2H over the passes, so only the FIRST pass's upload and the
// LAST pass's download are exposed.  Same dynamic programme as chunk_schedule with those two terms added (one cost unit = one
// 256 x 256 tile over one unit of K on a CTA pair, ~7.45 ns at 1.3 PFLOP/s; the four block GEMMs x depth x 1.4 for the rest of the
// block; PCIe at ~50 GB/s): ViT-H, 64 images: 5 x 12 + 4 instead of 16 + 4 x 12 (measured end to end 175.1 -> 175.9 images/s); ViT-B,
// 16 images: 4 + 8 + 4 instead of one pass, whose copies nothing overlapped (511.9 -> 664.8 images/s).  SVB_HOST_SCHEDULE=0 keeps the
// device schedule (A/B).
std::vector<int> chunk_schedule_host(const SchedGeom& gq, int batch, int max_chunk, double in_bytes_per_image, double out_bytes_per_image) {
    static const int on = [] { const char* v = getenv("SVB_HOST_SCHEDULE"); return v ? atoi(v) : 1; }();
    static const int fixed = [] { const char* v = getenv("SVB_FIXED_CHUNKS"); return v ? atoi(v) : 0; }();
    if (!on || fixed || batch < 2) return chunk_schedule(gq, batch, max_chunk);
    const long pairs = std::max(1, gq.sms / 2);
    const long D = gq.D, mlp = gq.mlp;
    const long shp[4][2] = {{3 * D, D}, {D, D}, {mlp, D}, {D, mlp}};
    auto cost = [&](int B) {
        double c = 0;
        for (auto& nk : shp) {
            const long tiles = (((long)B * gq.T + 255) / 256) * ((nk[0] + 255) / 256);
            c += (double)nk[1] * (double)((tiles + pairs - 1) / pairs);
        }
        return c;
    };
    const int cmax = std::min(max_chunk, batch);
    const int cmin = std::min(4, cmax);                                // first / last pass: at least 4 images where the chunk allows (smaller passes leave SMs idle in the attention kernels)
    const double unit_s = 7.45e-9 * (double)gq.depth * 1.4;            // seconds per cost unit (whole pass)
    const double per_pass = 6.0 * gq.depth * 4e-6 / unit_s;            // ~6 launches per block, ~4 us of fill / drain each
    const double h2d_units = in_bytes_per_image / 50e9 / unit_s, d2h_units = out_bytes_per_image / 50e9 / unit_s;
    // best[n] = cheapest way to run n images in the MIDDLE of the sequence (no exposure)
    std::vector<double> best(batch + 1, 1e300);
    std::vector<int> pick(batch + 1, 0);
    best[0] = 0;
    for (int n = 1; n <= batch; ++n)
        for (int c = 1; c <= std::min(cmax, n); ++c) {
            const double v = best[n - c] + cost(c) + per_pass;
            if (v < best[n]) { best[n] = v; pick[n] = c; }
        }
    double top = 1e300;
    int bf = 0, bl = 0;
    for (int f = cmin; f <= cmax; ++f)
        for (int l = 0; l <= cmax && f + l <= batch; ++l) {           // l = 0: a single pass (first == last)
            if (l == 0 ? f != batch : l < cmin) continue;
            if (cmin == 4 && ((f % 4 && f != batch) || l % 4)) continue;      // end passes in whole groups of 4 images (the sizes the parity tests and the A/B runs cover)
            const int mid = batch - f - l;
            const double v = cost(f) + per_pass + (l ? cost(l) + per_pass : 0.0) + best[mid] + h2d_units * f + d2h_units * (l ? l : f);
            if (v < top) { top = v; bf = f; bl = l; }
        }
    if (bf <= 0) return chunk_schedule(gq, batch, max_chunk);          // (no candidate, e.g. 3 images in passes of 2: the device schedule)
    std::vector<int> out;
    out.push_back(bf);
    std::vector<int> midv;
    for (int n = batch - bf - bl; n > 0; n -= pick[n]) midv.push_back(pick[n]);
    std::sort(midv.begin(), midv.end(), [](int a, int b) { return a > b; });
    out.insert(out.end(), midv.begin(), midv.end());
    if (bl) out.push_back(bl);
    return out;
}
std::vector<int> chunk_schedule_host(const svb_encoder* e, int batch, int max_chunk, double in_bytes_per_image, double out_bytes_per_image) {
    return chunk_schedule_host(geom_of(e), batch, max_chunk, in_bytes_per_image, out_bytes_per_image);
}

}  // namespace

extern "C" {

const char* svb_last_error(void) { return g_err; }
int svb_version(void) { return 100; }

int svb_encoder_create(const svb_config_t* cfg, svb_encoder_t** out) {
    SVB_REQUIRE(cfg && out, "svb_encoder_create: null argument");
    SVB_REQUIRE(cfg->embed_dim % cfg->num_heads == 0, "embed_dim %d not divisible by num_heads %d", cfg->embed_dim, cfg->num_heads);
    SVB_REQUIRE(cfg->img_size % cfg->patch_size == 0, "img_size %d not divisible by patch_size %d", cfg->img_size, cfg->patch_size);
    SVB_REQUIRE(cfg->embed_dim % 8 == 0, "embed_dim %d must be a multiple of 8", cfg->embed_dim);
    SVB_REQUIRE((cfg->img_size / cfg->patch_size) % 32 == 0, "token grid must be a multiple of 32 (img_size 1024, patch 16 -> 64)");
    SVB_REQUIRE(cfg->num_global >= 0 && cfg->num_global <= 16, "num_global out of range");
    svb_encoder* e = new svb_encoder();
    e->cfg = *cfg;
    e->D = cfg->embed_dim;
    e->depth = cfg->depth;
    e->heads = cfg->num_heads;
    e->hd = e->D / e->heads;
    e->grid = cfg->img_size / cfg->patch_size;
    e->T = e->grid * e->grid;
    e->mlp = cfg->mlp_dim;
    // SimpleFPN widths (image_encoder.py:416,427,440)
    e->d4 = std::max(cfg->fpn_dims[0] * 2, e->D / 2);
    e->d8 = std::max(cfg->fpn_dims[1], e->D / 2);
    e->d32 = std::max(cfg->fpn_dims[3], e->D * 2);
    const int D = e->D, p = cfg->patch_size, kpe = cfg->in_chans * p * p;
    add_param(e, "pos_embed", P_VEC, (int64_t)e->T * D);
    add_param(e, "patch_embed.proj.weight", P_GEMM_W, (int64_t)D * kpe, D, kpe);
    add_param(e, "patch_embed.proj.bias", P_VEC, D);
    for (int i = 0; i < e->depth; ++i) {
        const std::string b = "blocks." + std::to_string(i) + ".";
        const int L = 2 * (e->is_global(i) ? e->grid : cfg->window_size) - 1;
        add_param(e, b + "norm1.weight", P_VEC, D);
        add_param(e, b + "norm1.bias", P_VEC, D);
        add_param(e, b + "attn.rel_pos_h", P_REL_H, (int64_t)L * e->hd, L, i);
        add_param(e, b + "attn.rel_pos_w", P_REL_W, (int64_t)L * e->hd, L, i);
        add_param(e, b + "attn.qkv.weight", P_GEMM_W, (int64_t)3 * D * D, 3 * D, D);
        add_param(e, b + "attn.qkv.bias", P_VEC, 3 * D);
        add_param(e, b + "attn.proj.weight", P_GEMM_W, (int64_t)D * D, D, D);
        add_param(e, b + "attn.proj.bias", P_VEC, D);
        add_param(e, b + "norm2.weight", P_VEC, D);
        add_param(e, b + "norm2.bias", P_VEC, D);
        add_param(e, b + "mlp.lin1.weight", P_GEMM_W, (int64_t)e->mlp * D, e->mlp, D);
        add_param(e, b + "mlp.lin1.bias", P_VEC, e->mlp);
        add_param(e, b + "mlp.lin2.weight", P_GEMM_W, (int64_t)D * e->mlp, D, e->mlp);
        add_param(e, b + "mlp.lin2.bias", P_VEC, D);
    }
    for (const char* k : {"orig_neck.0.weight", "orig_neck.1.weight", "orig_neck.1.bias", "orig_neck.2.weight", "orig_neck.3.weight",
                          "orig_neck.3.bias"})
        add_param(e, k, P_IGNORED, 0);
    const int* od = cfg->fpn_dims;
    const int d4 = e->d4, d8 = e->d8, d32 = e->d32;
    add_param(e, "neck.down_4.0.weight", P_CONVT_W, (int64_t)D * d4 * 4, D, d4);
    add_param(e, "neck.down_4.0.bias", P_CONVT_B, d4);
    add_param(e, "neck.down_4.1.weight", P_VEC, d4);
    add_param(e, "neck.down_4.1.bias", P_VEC, d4);
    add_param(e, "neck.down_4.3.weight", P_CONVT_W, (int64_t)d4 * (d4 / 2) * 4, d4, d4 / 2);
    add_param(e, "neck.down_4.3.bias", P_CONVT_B, d4 / 2);
    add_param(e, "neck.down_4.4.weight", P_VEC, d4 / 2);
    add_param(e, "neck.down_4.4.bias", P_VEC, d4 / 2);
    add_param(e, "neck.down_4.5.weight", P_GEMM_W, (int64_t)od[0] * (d4 / 2), od[0], d4 / 2);
    add_param(e, "neck.down_4.5.bias", P_VEC, od[0]);
    add_param(e, "neck.down_4.6.weight", P_VEC, od[0]);
    add_param(e, "neck.down_4.6.bias", P_VEC, od[0]);
    add_param(e, "neck.down_8.0.weight", P_CONVT_W, (int64_t)D * d8 * 4, D, d8);
    add_param(e, "neck.down_8.0.bias", P_CONVT_B, d8);
    add_param(e, "neck.down_8.1.weight", P_VEC, d8);
    add_param(e, "neck.down_8.1.bias", P_VEC, d8);
    add_param(e, "neck.down_8.2.weight", P_GEMM_W, (int64_t)od[1] * d8, od[1], d8);
    add_param(e, "neck.down_8.2.bias", P_VEC, od[1]);
    add_param(e, "neck.down_8.3.weight", P_VEC, od[1]);
    add_param(e, "neck.down_8.3.bias", P_VEC, od[1]);
    add_param(e, "neck.down_16.0.weight", P_GEMM_W, (int64_t)od[2] * D, od[2], D);
    add_param(e, "neck.down_16.0.bias", P_VEC, od[2]);
    add_param(e, "neck.down_16.1.weight", P_VEC, od[2]);
    add_param(e, "neck.down_16.1.bias", P_VEC, od[2]);
    add_param(e, "neck.down_32.0.weight", P_CONV22_W, (int64_t)d32 * D * 4, D, d32);
    add_param(e, "neck.down_32.0.bias", P_VEC, d32);
    add_param(e, "neck.down_32.1.weight", P_VEC, d32);
    add_param(e, "neck.down_32.1.bias", P_VEC, d32);
    add_param(e, "neck.down_32.2.weight", P_GEMM_W, (int64_t)od[3] * d32, od[3], d32);
    add_param(e, "neck.down_32.2.bias", P_VEC, od[3]);
    add_param(e, "neck.down_32.3.weight", P_VEC, od[3]);
    add_param(e, "neck.down_32.3.bias", P_VEC, od[3]);
    for (auto& kv : e->params) {
        int rc = alloc_param_storage(kv.second);
        if (rc) { svb_encoder_destroy(e); return rc; }
    }
    // LayerNorm folding (bf16 path): needs whole 32-column epilogue chunks and the CTA-pair GEMM.  SVB_LN_FOLD=0 keeps the
    // explicit LayerNorm kernels (A/B comparisons: 96.0 vs 99.7 ms per 16 ViT-H images on the same box).
    {
        const char* env = getenv("SVB_LN_FOLD");
        e->ln_fold = (D % 32 == 0) && gemm_bf16_tc_supports_fold() && !(env && atoi(env) == 0);
        const char* genv = getenv("SVB_GN_FOLD");
        e->gn_fold = e->ln_fold && !(genv && atoi(genv) == 0) && (e->d8 % 32 == 0) && ((e->d4 / 2) % 32 == 0) && (e->d32 % 32 == 0);
        if (e->gn_fold) {
            for (const char* k : {"neck.down_8.2.weight", "neck.down_4.5.weight", "neck.down_32.2.weight"}) {
                int rc = alloc_fold_storage(e->params[k]);
                if (rc) { svb_encoder_destroy(e); return rc; }
            }
        }
        if (e->ln_fold) {
            for (int i = 0; i < e->depth; ++i) {
                const std::string b = "blocks." + std::to_string(i) + ".";
                for (const char* k : {"attn.qkv.weight", "mlp.lin1.weight"}) {
                    int rc = alloc_fold_storage(e->params[b + k]);
                    if (rc) { svb_encoder_destroy(e); return rc; }
                }
            }
        }
    }
    // tcgen05 attention covers the geometry _build_sam instantiates (64x64 grid, 14x14 windows, head_dim 64 / 80);
    // anything else runs the SIMT kernel.  SVB_ATTN_IMPL=0 forces the SIMT kernel (bisecting aid).
    e->grid_pad = ((e->grid + cfg->window_size - 1) / cfg->window_size) * cfg->window_size;
    e->attn_impl_bf16 = (e->grid == 64 && cfg->window_size == 14 && (e->hd == 64 || e->hd == 80)) ? 1 : 0;
    const char* env = getenv("SVB_ATTN_IMPL");
    if (env && atoi(env) == 0) e->attn_impl_bf16 = 0;
    if (e->attn_impl_bf16 == 1) {
        e->relpack.assign(e->depth, nullptr);
        for (int i = 0; i < e->depth; ++i) {
            const size_t bytes = sizeof(bf16) * (size_t)attention_tc_rel_rows(e->is_global(i) ? e->grid : cfg->window_size, e->grid) * e->hd;
            if (cudaMalloc(&e->relpack[i], bytes) != cudaSuccess || cudaMemset(e->relpack[i], 0, bytes) != cudaSuccess) {
                set_error("svb_encoder_create: cannot allocate the rel-pos table block of block %d", i);
                svb_encoder_destroy(e);
                return 1;
            }
        }
    }
    *out = e;
    return 0;
}

void svb_encoder_destroy(svb_encoder_t* e) {
    if (!e) return;
    for (auto& kv : e->params) {
        if (kv.second.f32) cudaFree(kv.second.f32);
        if (kv.second.b16) cudaFree(kv.second.b16);
        if (kv.second.fold_c) cudaFree(kv.second.fold_c);
        if (kv.second.fold_b) cudaFree(kv.second.fold_b);
    }
    if (e->taps) cudaFree(e->taps);
    e->drop_resized();
    for (bf16* r : e->relpack)
        if (r) cudaFree(r);
    auto& hp = e->hp;
    for (int i = 0; i < 2; ++i) {
        if (hp.xin[i]) cudaFree(hp.xin[i]);
        for (int k = 0; k < 4; ++k)
            if (hp.outs[i][k]) cudaFree(hp.outs[i][k]);
    }
    if (hp.ws) cudaFree(hp.ws);
    if (hp.events)
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(hp.in_done[i]); cudaEventDestroy(hp.comp_done[i]); cudaEventDestroy(hp.out_done[i]); }
    if (hp.s_in) cudaStreamDestroy(hp.s_in);
    if (hp.s_comp) cudaStreamDestroy(hp.s_comp);
    if (hp.s_out) cudaStreamDestroy(hp.s_out);
    delete e;
}

int svb_encoder_load_param(svb_encoder_t* e, const char* key, const float* data, int64_t numel, svb_stream_t stream) {
    SVB_REQUIRE(e && key && data, "svb_encoder_load_param: null argument");
    auto it = e->params.find(key);
    SVB_REQUIRE(it != e->params.end(), "unexpected state_dict key '%s'", key);
    Param& p = it->second;
    if (p.kind == P_IGNORED) { p.loaded = true; return 0; }
    SVB_REQUIRE(numel == p.numel, "size mismatch for '%s': got %lld elements, expected %lld", key, (long long)numel, (long long)p.numel);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    switch (p.kind) {
        case P_VEC:
            SVB_CHECK_CUDA(cudaMemcpyAsync(p.f32, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice, st));
            break;
        case P_REL_H:
        case P_REL_W:
            SVB_CHECK_CUDA(cudaMemcpyAsync(p.f32, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice, st));
            if (e->attn_impl_bf16 == 1) rc = pack_rel_table(data, e->relpack[p.b], p.a, e->hd, p.kind == P_REL_W, st);
            break;
        case P_GEMM_W:
            SVB_CHECK_CUDA(cudaMemcpyAsync(p.f32, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice, st));
            rc = pack_cast(data, p.b16, true, numel, st);
            break;
        case P_CONVT_W:
            rc = pack_convT(data, p.f32, false, p.a, p.b, st);
            if (!rc) rc = pack_convT(data, p.b16, true, p.a, p.b, st);
            break;
        case P_CONVT_B:
            rc = pack_bias4(data, p.f32, (int)numel, st);
            break;
        case P_CONV22_W:
            rc = pack_conv2x2(data, p.f32, false, p.a, p.b, st);
            if (!rc) rc = pack_conv2x2(data, p.b16, true, p.a, p.b, st);
            break;
        default: break;
    }
    if (rc) return rc;
    p.loaded = true;
    e->fold_dirty = true;       // the folded qkv / lin1 operands are re-derived from the fp32 masters at the next bf16 forward
    if (!e->pos_cache.empty() || !e->rel_cache.empty()) {
        cudaStreamSynchronize((cudaStream_t)stream);   // a forward still reading the resized tables must finish before they are freed
        e->drop_resized();
    }
    return 0;
}

int svb_encoder_missing_params(const svb_encoder_t* e) {
    int n = 0;
    for (auto& kv : e->params)
        if (kv.second.kind != P_IGNORED && !kv.second.loaded) ++n;
    return n;
}

size_t svb_encoder_workspace_bytes(const svb_encoder_t* e, int chunk, int mode) {
    if (!e || chunk <= 0) return 0;
    return plan(e, chunk, mode, nullptr).total;
}

int svb_encoder_forward(svb_encoder_t* e, const float* x, int batch, void* res2, void* res3, void* res4, void* res5, int out_dtype,
                        int mode, int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream) {
    SVB_REQUIRE(e && x && res2 && res3 && res4 && res5 && workspace, "svb_encoder_forward: null argument");
    SVB_REQUIRE(mode == SVB_MODE_BF16 || mode == SVB_MODE_FP32, "svb_encoder_forward: bad mode %d", mode);
    SVB_REQUIRE(out_dtype == SVB_DTYPE_F32 || out_dtype == SVB_DTYPE_BF16, "svb_encoder_forward: bad out_dtype %d", out_dtype);
    SVB_REQUIRE(batch > 0 && chunk > 0, "svb_encoder_forward: batch %d / chunk %d must be positive", batch, chunk);
    const int missing = svb_encoder_missing_params(e);
    SVB_REQUIRE(missing == 0, "svb_encoder_forward: %d parameters have not been loaded", missing);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    if (chunk > batch) chunk = batch;
    const size_t need = plan(e, chunk, mode, nullptr).total;
    SVB_REQUIRE(workspace_bytes >= need, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    const Buffers bf = plan(e, chunk, mode, workspace);
    const size_t osz = out_dtype == SVB_DTYPE_BF16 ? 2 : 4;
    const size_t in_per_img = (size_t)e->cfg.in_chans * e->cfg.img_size * e->cfg.img_size;
    char* res[4] = {(char*)res2, (char*)res3, (char*)res4, (char*)res5};
    int b0 = 0;
    for (int B : chunk_schedule(e, batch, chunk)) {
        void* outs[4];
        for (int k = 0; k < 4; ++k) outs[k] = res[k] + (size_t)b0 * out_elems_per_image(e, k) * osz;
        int rc = forward_chunk(e, x + (size_t)b0 * in_per_img, B, outs, out_dtype, mode, bf, (cudaStream_t)stream);
        if (rc) return rc;
        b0 += B;
    }
    return 0;
}

size_t svb_encoder_workspace_bytes_hw(const svb_encoder_t* e, int chunk, int mode, int img_h, int img_w) {
    if (!e || chunk <= 0 || img_h <= 0 || img_w <= 0) return 0;
    const int p = e->cfg.patch_size;
    return plan(e, chunk, mode, nullptr, img_h / p, img_w / p).total;
}

int svb_encoder_forward_x(svb_encoder_t* e, const void* x, int x_dtype, int batch, int img_h, int img_w, void* res2, void* res3, void* res4,
                          void* res5, int out_dtype, int mode, int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream) {
    SVB_REQUIRE(e && x && res2 && res3 && res4 && res5 && workspace, "svb_encoder_forward_x: null argument");
    SVB_REQUIRE(mode == SVB_MODE_BF16 || mode == SVB_MODE_FP32, "svb_encoder_forward_x: bad mode %d", mode);
    SVB_REQUIRE(out_dtype == SVB_DTYPE_F32 || out_dtype == SVB_DTYPE_BF16, "svb_encoder_forward_x: bad out_dtype %d", out_dtype);
    SVB_REQUIRE(x_dtype == SVB_DTYPE_F32 || x_dtype == SVB_DTYPE_BF16 || x_dtype == SVB_DTYPE_F16, "svb_encoder_forward_x: bad input dtype %d", x_dtype);
    SVB_REQUIRE(batch > 0 && chunk > 0, "svb_encoder_forward_x: batch %d / chunk %d must be positive", batch, chunk);
    const int p = e->cfg.patch_size;
    SVB_REQUIRE(img_h > 0 && img_w > 0 && img_h % (32 * p) == 0 && img_w % (32 * p) == 0,
                "svb_encoder_forward_x: image %d x %d: both sides must be multiples of %d (token grid in multiples of 32)", img_h, img_w, 32 * p);
    const int missing = svb_encoder_missing_params(e);
    SVB_REQUIRE(missing == 0, "svb_encoder_forward_x: %d parameters have not been loaded", missing);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    if (chunk > batch) chunk = batch;
    const bool native = (img_h == e->cfg.img_size && img_w == e->cfg.img_size);
    const int T = (img_h / p) * (img_w / p);
    const size_t need = plan(e, chunk, mode, nullptr, img_h / p, img_w / p).total;
    SVB_REQUIRE(workspace_bytes >= need, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    const Buffers bf = plan(e, chunk, mode, workspace, img_h / p, img_w / p);
    const size_t osz = out_dtype == SVB_DTYPE_BF16 ? 2 : 4;
    const size_t in_per_img = (size_t)e->cfg.in_chans * img_h * img_w * (x_dtype == SVB_DTYPE_F32 ? 4 : 2);     // bytes
    char* res[4] = {(char*)res2, (char*)res3, (char*)res4, (char*)res5};
    std::vector<int> passes;
    if (native) passes = chunk_schedule(e, batch, chunk);
    else for (int b0 = 0; b0 < batch; b0 += chunk) passes.push_back(std::min(chunk, batch - b0));
    int b0 = 0;
    for (int B : passes) {
        void* outs[4];
        for (int k = 0; k < 4; ++k) outs[k] = res[k] + (size_t)b0 * out_elems_per_image(e, k, img_h, img_w) * osz;
        int rc = forward_chunk(e, (const char*)x + (size_t)b0 * in_per_img, B, outs, out_dtype, mode, bf, (cudaStream_t)stream, nullptr, img_h, img_w,
                               x_dtype);
        if (rc) return rc;
        b0 += B;
    }
    return 0;
}

int svb_encoder_forward_hw(svb_encoder_t* e, const float* x, int batch, int img_h, int img_w, void* res2, void* res3, void* res4, void* res5,
                           int out_dtype, int mode, int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream) {
    return svb_encoder_forward_x(e, x, SVB_DTYPE_F32, batch, img_h, img_w, res2, res3, res4, res5, out_dtype, mode, chunk, workspace, workspace_bytes,
                                 stream);
}

int svb_resize_pos_embed(const float* src, float* dst, int h0, int w0, int h1, int w1, int dim, svb_stream_t stream) {
    SVB_REQUIRE(src && dst, "svb_resize_pos_embed: null argument");
    return resize_pos_embed(src, dst, h0, w0, h1, w1, dim, (cudaStream_t)stream);
}
int svb_resize_rel_pos(const float* src, float* dst, int len0, int len1, int head_dim, svb_stream_t stream) {
    SVB_REQUIRE(src && dst, "svb_resize_rel_pos: null argument");
    return resize_rel_pos(src, dst, len0, len1, head_dim, (cudaStream_t)stream);
}

int svb_encoder_forward_u8(svb_encoder_t* e, const uint8_t* const* images, const int* heights, const int* widths, const float* pixel_mean,
                           const float* pixel_std, int batch, void* res2, void* res3, void* res4, void* res5, int out_dtype, int mode,
                           int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream) {
    SVB_REQUIRE(e && images && heights && widths && pixel_mean && pixel_std && res2 && res3 && res4 && res5 && workspace,
                "svb_encoder_forward_u8: null argument");
    SVB_REQUIRE(mode == SVB_MODE_BF16 || mode == SVB_MODE_FP32, "svb_encoder_forward_u8: bad mode %d", mode);
    SVB_REQUIRE(out_dtype == SVB_DTYPE_F32 || out_dtype == SVB_DTYPE_BF16, "svb_encoder_forward_u8: bad out_dtype %d", out_dtype);
    SVB_REQUIRE(batch > 0 && chunk > 0, "svb_encoder_forward_u8: batch %d / chunk %d must be positive", batch, chunk);
    const int missing = svb_encoder_missing_params(e);
    SVB_REQUIRE(missing == 0, "svb_encoder_forward_u8: %d parameters have not been loaded", missing);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    for (int c = 0; c < e->cfg.in_chans; ++c) SVB_REQUIRE(pixel_std[c] != 0.f, "svb_encoder_forward_u8: pixel_std[%d] is zero", c);
    if (chunk > batch) chunk = batch;
    const size_t need = plan(e, chunk, mode, nullptr).total;
    SVB_REQUIRE(workspace_bytes >= need, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    const Buffers bf = plan(e, chunk, mode, workspace);
    const size_t osz = out_dtype == SVB_DTYPE_BF16 ? 2 : 4;
    char* res[4] = {(char*)res2, (char*)res3, (char*)res4, (char*)res5};
    int b0 = 0;
    for (int B : chunk_schedule(e, batch, chunk)) {
        void* outs[4];
        for (int k = 0; k < 4; ++k) outs[k] = res[k] + (size_t)b0 * out_elems_per_image(e, k) * osz;
        U8Input u8;
        u8.images = images + b0; u8.hs = heights + b0; u8.ws = widths + b0; u8.mean = pixel_mean; u8.stdv = pixel_std;
        int rc = forward_chunk(e, nullptr, B, outs, out_dtype, mode, bf, (cudaStream_t)stream, &u8);
        if (rc) return rc;
        b0 += B;
    }
    return 0;
}

int svb_stage_images_u8(const uint8_t* const* images, const int* heights, const int* widths, int batch, int chans, int img, int patch,
                        const float* pixel_mean, const float* pixel_std, void* out, int out_dtype, svb_stream_t stream) {
    SVB_REQUIRE(images && heights && widths && pixel_mean && pixel_std && out, "svb_stage_images_u8: null argument");
    return stage_u8_patch(images, heights, widths, pixel_mean, pixel_std, out, out_dtype == SVB_DTYPE_BF16, batch, chans, img, patch,
                          (cudaStream_t)stream);
}

int svb_encoder_forward_host(svb_encoder_t* e, const float* x_host, int batch, void* res2_host, void* res3_host, void* res4_host,
                             void* res5_host, int out_dtype, int mode, int chunk) {
    SVB_REQUIRE(e && x_host && res2_host && res3_host && res4_host && res5_host, "svb_encoder_forward_host: null argument");
    SVB_REQUIRE(batch > 0 && chunk > 0, "svb_encoder_forward_host: batch/chunk must be positive");
    const int missing = svb_encoder_missing_params(e);
    SVB_REQUIRE(missing == 0, "svb_encoder_forward_host: %d parameters have not been loaded", missing);
    if (chunk > batch) chunk = batch;
    auto& hp = e->hp;
    const size_t osz = out_dtype == SVB_DTYPE_BF16 ? 2 : 4;
    const size_t in_per_img = (size_t)e->cfg.in_chans * e->cfg.img_size * e->cfg.img_size;
    if (hp.chunk != chunk || hp.mode != mode || hp.out_dtype != out_dtype) {
        for (int i = 0; i < 2; ++i) {
            if (hp.xin[i]) { cudaFree(hp.xin[i]); hp.xin[i] = nullptr; }
            for (int k = 0; k < 4; ++k)
                if (hp.outs[i][k]) { cudaFree(hp.outs[i][k]); hp.outs[i][k] = nullptr; }
        }
        if (hp.ws) { cudaFree(hp.ws); hp.ws = nullptr; }
        for (int i = 0; i < 2; ++i) {
            SVB_CHECK_CUDA(cudaMalloc(&hp.xin[i], sizeof(float) * in_per_img * chunk));
            for (int k = 0; k < 4; ++k) SVB_CHECK_CUDA(cudaMalloc(&hp.outs[i][k], osz * out_elems_per_image(e, k) * chunk));
        }
        hp.ws_bytes = plan(e, chunk, mode, nullptr).total;
        SVB_CHECK_CUDA(cudaMalloc(&hp.ws, hp.ws_bytes));
        if (!hp.s_in) {
            SVB_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
            SVB_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.s_comp, cudaStreamNonBlocking));
            SVB_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                SVB_CHECK_CUDA(cudaEventCreateWithFlags(&hp.in_done[i], cudaEventDisableTiming));
                SVB_CHECK_CUDA(cudaEventCreateWithFlags(&hp.comp_done[i], cudaEventDisableTiming));
                SVB_CHECK_CUDA(cudaEventCreateWithFlags(&hp.out_done[i], cudaEventDisableTiming));
            }
            hp.events = true;
        }
        hp.chunk = chunk; hp.mode = mode; hp.out_dtype = out_dtype;
    }
    const Buffers bf = plan(e, chunk, mode, hp.ws);
    char* res[4] = {(char*)res2_host, (char*)res3_host, (char*)res4_host, (char*)res5_host};
    int it = 0, b0 = 0;
    double out_bytes_per_image = 0;
    for (int k = 0; k < 4; ++k) out_bytes_per_image += (double)osz * out_elems_per_image(e, k);
    for (int B : chunk_schedule_host(e, batch, chunk, sizeof(float) * (double)in_per_img, out_bytes_per_image)) {
        const int s = it & 1;
        // H2D of this chunk may start once the compute that last read xin[s] (two chunks ago) has finished
        if (it >= 2) SVB_CHECK_CUDA(cudaStreamWaitEvent(hp.s_in, hp.comp_done[s], 0));
        SVB_CHECK_CUDA(cudaMemcpyAsync(hp.xin[s], x_host + (size_t)b0 * in_per_img, sizeof(float) * in_per_img * B, cudaMemcpyHostToDevice, hp.s_in));
        SVB_CHECK_CUDA(cudaEventRecord(hp.in_done[s], hp.s_in));
        SVB_CHECK_CUDA(cudaStreamWaitEvent(hp.s_comp, hp.in_done[s], 0));
        if (it >= 2) SVB_CHECK_CUDA(cudaStreamWaitEvent(hp.s_comp, hp.out_done[s], 0));   // outs[s] drained to the host
        int rc = forward_chunk(e, hp.xin[s], B, hp.outs[s], out_dtype, mode, bf, hp.s_comp);
        if (rc) return rc;
        SVB_CHECK_CUDA(cudaEventRecord(hp.comp_done[s], hp.s_comp));
        SVB_CHECK_CUDA(cudaStreamWaitEvent(hp.s_out, hp.comp_done[s], 0));
        for (int k = 0; k < 4; ++k) {
            const size_t bytes = osz * out_elems_per_image(e, k);
            SVB_CHECK_CUDA(cudaMemcpyAsync(res[k] + (size_t)b0 * bytes, hp.outs[s][k], bytes * B, cudaMemcpyDeviceToHost, hp.s_out));
        }
        SVB_CHECK_CUDA(cudaEventRecord(hp.out_done[s], hp.s_out));
        b0 += B;
        ++it;
    }
    SVB_CHECK_CUDA(cudaStreamSynchronize(hp.s_in));
    SVB_CHECK_CUDA(cudaStreamSynchronize(hp.s_comp));
    SVB_CHECK_CUDA(cudaStreamSynchronize(hp.s_out));
    return 0;
}

int svb_pass_schedule_model(int embed_dim, int mlp_dim, int depth, int tokens, int sms, int batch, int chunk, int host_path,
                            double in_bytes_per_image, double out_bytes_per_image, int* passes, int max_passes) {
    SVB_REQUIRE(passes && embed_dim > 0 && mlp_dim > 0 && depth > 0 && tokens > 0 && sms > 1 && batch > 0 && chunk > 0 && max_passes > 0,
                "svb_pass_schedule_model: bad argument");
    const SchedGeom gq{embed_dim, mlp_dim, tokens, depth, sms};
    const std::vector<int> sch = host_path ? chunk_schedule_host(gq, batch, std::min(chunk, batch), in_bytes_per_image, out_bytes_per_image)
                                           : chunk_schedule(gq, batch, chunk);
    SVB_REQUIRE((int)sch.size() <= max_passes, "svb_pass_schedule_model: %d passes do not fit %d slots", (int)sch.size(), max_passes);
    for (size_t i = 0; i < sch.size(); ++i) passes[i] = sch[i];
    return (int)sch.size() + 1000;
}

int svb_encoder_pass_schedule(svb_encoder_t* e, int batch, int chunk, int host_path, int out_dtype, int* passes, int max_passes) {
    SVB_REQUIRE(e && passes && batch > 0 && chunk > 0 && max_passes > 0, "svb_encoder_pass_schedule: bad argument");
    std::vector<int> sch;
    if (host_path) {
        const size_t osz = out_dtype == SVB_DTYPE_BF16 ? 2 : 4;
        double out_bytes = 0;
        for (int k = 0; k < 4; ++k) out_bytes += (double)osz * out_elems_per_image(e, k);
        sch = chunk_schedule_host(e, batch, std::min(chunk, batch), sizeof(float) * (double)e->cfg.in_chans * e->cfg.img_size * e->cfg.img_size, out_bytes);
    } else {
        sch = chunk_schedule(e, batch, chunk);
    }
    SVB_REQUIRE((int)sch.size() <= max_passes, "svb_encoder_pass_schedule: %d passes do not fit %d slots", (int)sch.size(), max_passes);
    for (size_t i = 0; i < sch.size(); ++i) passes[i] = sch[i];
    return (int)sch.size() + 1000;        // 1000 + number of passes (0 would read as "ok, nothing written"; errors are 1..)
}

int svb_encoder_enable_taps(svb_encoder_t* e, int enable) {
    SVB_REQUIRE(e, "null encoder");
    if (enable && !e->taps) SVB_CHECK_CUDA(cudaMalloc(&e->taps, sizeof(float) * (size_t)(e->depth + 1) * e->T * e->D));
    e->taps_enabled = enable != 0;
    return 0;
}

int svb_encoder_read_tap(svb_encoder_t* e, int block, float* dst, int64_t numel, svb_stream_t stream) {
    SVB_REQUIRE(e && e->taps, "taps are not enabled");
    SVB_REQUIRE(block >= -1 && block < e->depth, "tap index %d out of range", block);
    SVB_REQUIRE(numel == (int64_t)e->T * e->D, "tap size mismatch: %lld vs %lld", (long long)numel, (long long)e->T * e->D);
    SVB_CHECK_CUDA(cudaMemcpyAsync(dst, e->taps + (size_t)(block + 1) * e->T * e->D, sizeof(float) * numel, cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream));
    return 0;
}

int svb_linear(int mode, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act_gelu,
               const float* resid, int ldr, int resid_mod, void* out, int out_dtype, int ldo, double* gn_stats, int rows_per_sample,
               int remap_grid, int remap_grid_pad, svb_stream_t stream) {
    SVB_REQUIRE(A && W && out, "svb_linear: null argument");
    Epilogue ep;
    ep.bias = bias;
    ep.act = (act_gelu == 1 || act_gelu == 2) ? act_gelu : 0;     // 1 GELU (erf), 2 ReLU
    ep.resid = resid; ep.ldr = ldr; ep.resid_mod = resid_mod;
    ep.out = out; ep.out_bf16 = out_dtype == SVB_DTYPE_BF16; ep.ldo = ldo;
    ep.stats = gn_stats; ep.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
    if (remap_grid > 0) {
        SVB_REQUIRE(remap_grid_pad >= remap_grid && M % (remap_grid * remap_grid) == 0,
                    "svb_linear: row remap needs M to be a multiple of grid^2 and grid_pad >= grid");
        ep.remap_g = remap_grid; ep.remap_gp = remap_grid_pad;
    }
    if (mode == SVB_MODE_BF16) return gemm_bf16_tc((const bf16*)A, lda, (const bf16*)W, ldw, M, N, K, ep, (cudaStream_t)stream);
    if (mode == SVB_MODE_FP32) return gemm_f32_simt((const float*)A, lda, (const float*)W, ldw, M, N, K, ep, (cudaStream_t)stream);
    SVB_REQUIRE(false, "svb_linear: bad mode %d", mode);
}

int svb_linear_nt(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, float* out_t, int ldo,
                  svb_stream_t stream) {
    SVB_REQUIRE(A && W && out_t, "svb_linear_nt: null argument");
    SVB_REQUIRE(ldo >= M && (ldo % 8) == 0, "svb_linear_nt: ldo (%d) must be >= M (%d) and a multiple of 8", ldo, M);
    Epilogue ep;
    ep.bias = bias;
    ep.out = out_t; ep.out_bf16 = 0; ep.ldo = ldo;
    ep.out_t = 1;
    return gemm_bf16_tc((const bf16*)A, lda, (const bf16*)W, ldw, M, N, K, ep, (cudaStream_t)stream);
}

int svb_linear_fused(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act_gelu,
                     const float* resid, int ldr, int resid_mod, void* out, int out_dtype, int ldo, const float* ln_stats,
                     const float* ln_colsum, int ln_dim, float ln_eps, void* out_bf16_copy, int ldo2, float* stat_out,
                     int remap_grid, int remap_grid_pad, svb_stream_t stream) {
    SVB_REQUIRE(A && W && out, "svb_linear_fused: null argument");
    Epilogue ep;
    ep.bias = bias;
    ep.act = (act_gelu == 1 || act_gelu == 2) ? act_gelu : 0;     // 1 GELU (erf), 2 ReLU
    ep.resid = resid; ep.ldr = ldr; ep.resid_mod = resid_mod;
    ep.out = out; ep.out_bf16 = out_dtype == SVB_DTYPE_BF16; ep.ldo = ldo;
    if (ln_stats) {
        ep.ln_stats = reinterpret_cast<const float2*>(ln_stats); ep.ln_c = ln_colsum; ep.ln_dim = ln_dim; ep.ln_parts = (ln_dim + 127) / 128;
        ep.ln_eps = ln_eps;
    }
    ep.out2 = out_bf16_copy; ep.ldo2 = ldo2;
    ep.stat_out = reinterpret_cast<float2*>(stat_out);
    if (remap_grid > 0) {
        SVB_REQUIRE(remap_grid_pad >= remap_grid && M % (remap_grid * remap_grid) == 0,
                    "svb_linear_fused: row remap needs M to be a multiple of grid^2 and grid_pad >= grid");
        ep.remap_g = remap_grid; ep.remap_gp = remap_grid_pad;
    }
    return gemm_bf16_tc((const bf16*)A, lda, (const bf16*)W, ldw, M, N, K, ep, (cudaStream_t)stream);
}

int svb_fold_layernorm(const float* W, const float* bias, const float* gamma, const float* beta, void* Wg_bf16, float* colsum,
                       float* bias_f, int N, int K, svb_stream_t stream) {
    SVB_REQUIRE(W && gamma && beta && Wg_bf16 && colsum && bias_f, "svb_fold_layernorm: null argument");
    return fold_layernorm(W, bias, gamma, beta, (bf16*)Wg_bf16, colsum, bias_f, N, K, (cudaStream_t)stream);
}

int svb_add_cast(const float* a, const float* b, void* out, int out_dtype, int64_t numel, svb_stream_t stream) {
    SVB_REQUIRE(a && out && numel >= 0, "svb_add_cast: bad argument");
    if (numel == 0) return 0;
    return add_cast(a, b, out, out_dtype == SVB_DTYPE_BF16, (size_t)numel, (cudaStream_t)stream);
}

int svb_layernorm(float* x, const void* add, const float* weight, const float* bias, void* out, int out_dtype, int rows, int dim, float eps,
                  svb_stream_t stream) {
    SVB_REQUIRE(x && weight && bias && out, "svb_layernorm: null argument");
    return layernorm_rows(x, add, weight, bias, out, out_dtype == SVB_DTYPE_BF16, rows, dim, eps, (cudaStream_t)stream);
}

int svb_layernorm_post(const float* x, const float* add, const float* weight, const float* bias, float* out, void* out_bf16,
                       const float* pos, int pos_rows, void* out_q_bf16, int rows, int dim, float eps, svb_stream_t stream) {
    SVB_REQUIRE(x && weight && bias, "svb_layernorm_post: null argument");
    SVB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(out) |
                  reinterpret_cast<uintptr_t>(pos)) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(out_bf16) | reinterpret_cast<uintptr_t>(out_q_bf16)) & 7) == 0,
                "svb_layernorm_post: the fp32 operands must be 16-byte aligned, the bf16 outputs 8-byte aligned");
    return layernorm_post_rows(x, add, weight, bias, out, (bf16*)out_bf16, pos, pos_rows, (bf16*)out_q_bf16, rows, dim, eps, (cudaStream_t)stream);
}

int svb_attention(int impl, int dtype, const void* qkv, void* out, const float* rel_pos_h, const float* rel_pos_w, const float* qkv_bias,
                  int batch, int grid, int ws, int heads, int head_dim, svb_stream_t stream) {
    SVB_REQUIRE(qkv && out && rel_pos_h && rel_pos_w && qkv_bias, "svb_attention: null argument");
    if (impl == 0) {
        AttnParams ap;
        ap.qkv = qkv; ap.out = out; ap.rel_h = rel_pos_h; ap.rel_w = rel_pos_w; ap.qkv_bias = qkv_bias;
        ap.batch = batch; ap.grid = grid; ap.ws = ws; ap.heads = heads; ap.hd = head_dim;
        return attention_simt(ap, dtype == SVB_DTYPE_BF16, (cudaStream_t)stream);
    }
    SVB_REQUIRE(false, "svb_attention: impl %d is not available in this build", impl);
}

int svb_fill_pad_rows_hw(void* qkv_padded, const float* qkv_bias, int batch, int grid_h, int grid_w, int row_len, svb_stream_t stream) {
    SVB_REQUIRE(qkv_padded && qkv_bias && batch > 0 && grid_h > 0 && grid_w > 0, "svb_fill_pad_rows_hw: bad argument");
    return fill_pad_rows((bf16*)qkv_padded, qkv_bias, batch, grid_h, grid_w, (grid_h + 13) / 14 * 14, (grid_w + 13) / 14 * 14, row_len, (cudaStream_t)stream);
}

int svb_attention_window_hw(const void* qkv_padded, void* out, const void* rel_pack, int batch, int grid_h, int grid_w, int heads, int head_dim,
                            svb_stream_t stream) {
    SVB_REQUIRE(qkv_padded && out && rel_pack, "svb_attention_window_hw: null argument");
    AttnTcParams ap;
    ap.qkv = (const bf16*)qkv_padded; ap.out = (bf16*)out; ap.rel_pack = (const bf16*)rel_pack;
    ap.batch = batch; ap.grid = 0; ap.grid_h = grid_h; ap.grid_w = grid_w; ap.ws = 14; ap.heads = heads; ap.hd = head_dim;
    return attention_tc(ap, (cudaStream_t)stream);
}

size_t svb_attention_global_hw_workspace(int batch, int grid_h, int grid_w, int heads, int head_dim) {
    const size_t M = (size_t)batch * grid_h * grid_w;
    const size_t ldh = (2 * grid_h - 1 + 7) / 8 * 8, ldw = (2 * grid_w - 1 + 7) / 8 * 8;
    return M * heads * (ldh + ldw) * sizeof(float) + (ldh + ldw) * (size_t)head_dim * sizeof(bf16) + 4096;
}

int svb_attention_global_hw(const void* qkv, void* out, const float* rel_pos_h, const float* rel_pos_w, int batch, int grid_h, int grid_w,
                            int heads, int head_dim, void* workspace, size_t workspace_bytes, svb_stream_t stream) {
    SVB_REQUIRE(qkv && out && rel_pos_h && rel_pos_w && workspace, "svb_attention_global_hw: null argument");
    SVB_REQUIRE(workspace_bytes >= svb_attention_global_hw_workspace(batch, grid_h, grid_w, heads, head_dim) &&
                (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "svb_attention_global_hw: workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int D = heads * head_dim, M = batch * grid_h * grid_w;
    const int Lh = 2 * grid_h - 1, Lw = 2 * grid_w - 1, ldh = (Lh + 7) / 8 * 8, ldw = (Lw + 7) / 8 * 8;
    float* bh = (float*)workspace;
    float* bw = bh + (size_t)M * heads * ldh;
    bf16* th = (bf16*)(((uintptr_t)(bw + (size_t)M * heads * ldw) + 255) & ~(uintptr_t)255);
    bf16* tw = th + (size_t)ldh * head_dim;
    int rc;
    if ((rc = add_cast(rel_pos_h, nullptr, th, true, (size_t)Lh * head_dim, st))) return rc;
    if ((rc = add_cast(rel_pos_w, nullptr, tw, true, (size_t)Lw * head_dim, st))) return rc;
    for (int w = 0; w < 2; ++w)
        for (int hh = 0; hh < heads; ++hh) {
            Epilogue ep;
            ep.out = (w ? bw : bh) + (size_t)hh * (w ? ldw : ldh); ep.ldo = heads * (w ? ldw : ldh);
            if ((rc = gemm_bf16_tc((const bf16*)qkv + (size_t)hh * head_dim, 3 * D, w ? tw : th, head_dim, M, w ? Lw : Lh, head_dim, ep, st))) return rc;
        }
    AttnTcParams ap;
    ap.qkv = (const bf16*)qkv; ap.out = (bf16*)out; ap.rel_pack = nullptr;
    ap.batch = batch; ap.grid = 0; ap.grid_h = grid_h; ap.grid_w = grid_w; ap.ws = 0; ap.heads = heads; ap.hd = head_dim;
    ap.bias_h = bh; ap.bias_w = bw; ap.bias_ld_h = ldh; ap.bias_ld_w = ldw;
    return attention_global_ext(ap, st);
}

int svb_attention_tc_phases(const void* qkv, void* out, const void* rel_pack, int batch, int grid, int ws, int heads, int head_dim,
                            long long* phase_clocks, svb_stream_t stream) {
    SVB_REQUIRE(qkv && out && rel_pack, "svb_attention_tc_phases: null argument");
    AttnTcParams ap;
    ap.qkv = (const bf16*)qkv; ap.out = (bf16*)out; ap.rel_pack = (const bf16*)rel_pack;
    ap.batch = batch; ap.grid = grid; ap.ws = ws; ap.heads = heads; ap.hd = head_dim;
    ap.phase_clocks = phase_clocks;
    return attention_tc(ap, (cudaStream_t)stream);
}

int svb_attention_tc(const void* qkv, void* out, const void* rel_pack, int batch, int grid, int ws, int heads, int head_dim,
                     svb_stream_t stream) {
    SVB_REQUIRE(qkv && out && rel_pack, "svb_attention_tc: null argument");
    AttnTcParams ap;
    ap.qkv = (const bf16*)qkv; ap.out = (bf16*)out; ap.rel_pack = (const bf16*)rel_pack;
    ap.batch = batch; ap.grid = grid; ap.ws = ws; ap.heads = heads; ap.hd = head_dim;
    return attention_tc(ap, (cudaStream_t)stream);
}

int svb_rel_pack_rows(int ws, int grid) { return attention_tc_rel_rows(ws, grid); }

int svb_attention_debug_buffer(void* mapped_device_ptr) { return attention_tc_set_debug_buffer(mapped_device_ptr); }

int svb_pack_rel_table(const float* table, void* rel_pack, int table_len, int head_dim, int is_w, svb_stream_t stream) {
    SVB_REQUIRE(table && rel_pack, "svb_pack_rel_table: null argument");
    return pack_rel_table(table, (bf16*)rel_pack, table_len, head_dim, is_w != 0, (cudaStream_t)stream);
}

int svb_fill_pad_rows(void* qkv_padded, const float* qkv_bias, int batch, int grid, int grid_pad, int row_len, svb_stream_t stream) {
    SVB_REQUIRE(qkv_padded && qkv_bias, "svb_fill_pad_rows: null argument");
    SVB_REQUIRE(grid > 0 && grid_pad >= grid, "svb_fill_pad_rows: bad grid %d / padded grid %d", grid, grid_pad);
    return fill_pad_rows((bf16*)qkv_padded, qkv_bias, batch, grid, grid, grid_pad, grid_pad, row_len, (cudaStream_t)stream);
}

int svb_im2col(const float* x, void* out, int out_dtype, int batch, int chans, int img, int patch, svb_stream_t stream) {
    SVB_REQUIRE(x && out, "svb_im2col: null argument");
    return im2col_patch(x, SVB_DTYPE_F32, out, out_dtype == SVB_DTYPE_BF16, batch, chans, img, img, patch, (cudaStream_t)stream);
}

int svb_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out, int out_dtype,
                        int64_t rows, int C, int64_t rows_per_sample, float eps, int gelu, svb_stream_t stream) {
    SVB_REQUIRE(x && stats && gamma && beta && out, "svb_groupnorm_apply: null argument");
    return groupnorm_apply(x, stats, gamma, beta, out, out_dtype == SVB_DTYPE_BF16, (long)rows, C, (long)rows_per_sample, eps, gelu,
                           (cudaStream_t)stream);
}

int svb_groupnorm_apply_nchw(const float* x, const double* stats, const float* gamma, const float* beta, void* out, int out_dtype,
                             int batch, int grid, int levels, int C, float eps, int gelu, svb_stream_t stream) {
    SVB_REQUIRE(x && stats && gamma && beta && out, "svb_groupnorm_apply_nchw: null argument");
    return groupnorm_apply_nchw(x, stats, gamma, beta, out, out_dtype, batch, grid, grid, levels, C, eps, gelu, (cudaStream_t)stream);
}

}  // extern "C"

