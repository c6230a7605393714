// Shared declarations for the sm_100a SAM ViT encoder kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace svb {

typedef __nv_bfloat16 bf16;

// ---- error plumbing: every C-ABI entry returns 0 on success; message via svb_last_error() ----
void set_error(const char* fmt, ...);
#define SVB_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            svb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
#define SVB_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            svb::set_error(__VA_ARGS__);       \
            return 2;                          \
        }                                      \
    } while (0)

// ---- GEMM epilogue description shared by the tcgen05 GEMM and the fp32 SIMT GEMM ----
// out[r, c] = act(acc[r, c] + bias[c]) + resid[(resid_mod ? r % resid_mod : r), c]
// optionally accumulating sum / sum-of-squares of (acc + bias) per sample for GroupNorm(1, C).
struct Epilogue {
    const float* bias = nullptr;     // [N] fp32
    const float* resid = nullptr;    // fp32, leading dim ldr; may alias `out` when out is fp32 (in-place residual)
    int resid_mod = 0;
    int ldr = 0;
    int act = 0;                     // 0 none, 1 GELU (erf form)
    void* out = nullptr;
    int out_bf16 = 0;                // 1: bf16 output, 0: fp32 output
    int ldo = 0;
    double* stats = nullptr;         // [samples][2] (sum, sumsq) or null
    int rows_per_sample = 1;
    // optional output-row remap (token grid g x g -> padded grid gp x gp, the window_partition padding of
    // image_encoder.py:271-275 expressed as a store address): out row = (r / g^2) * gp^2 + ((r % g^2) / g) * gp + r % g
    int remap_g = 0, remap_gp = 0;   // grid WIDTH / padded width
    int remap_h = 0, remap_hp = 0;   // grid height / padded height (0: square, = remap_g / remap_gp)
    // with the remap: also write the PAD rows of the padded grid (= bf16(pad_bias), the qkv bias: pad tokens are zero after norm1,
    // image_encoder.py:183-187,271-275) from otherwise idle warps of the GEMM, instead of a separate fill_pad_rows launch
    const float* pad_bias = nullptr;
    // ---- LayerNorm folded into the GEMM (bf16 path; tcgen05 pair kernel only) ----
    // consumer side: A holds the UN-normalised rows x (bf16), W = bf16(gamma (.) W0), bias = b0 + W0 beta; with the row
    // statistics mu, rstd of x:  LayerNorm(x) W0^T + b0 = rstd * (x W^T - mu * ln_c) + bias,  ln_c[n] = sum_k W[n, k].
    // ln_stats = [M][ln_parts] partial (sum, sum of squares) of each row of x over ln_dim columns (written by the producer).
    const float2* ln_stats = nullptr;
    const float* ln_c = nullptr;     // [N]
    int ln_parts = 0, ln_dim = 0;
    float ln_eps = 0.f;
    // the same fold for GroupNorm(1, C) -> 1x1 conv (SimpleFPN, image_encoder.py:428-447): the statistics are per SAMPLE —
    // gn_in_stats[sample] = (sum, sumsq) in fp64 over gn_in_rows rows x K channels, accumulated by the producing GEMM (`stats`).
    // When set it replaces ln_stats as the source of mu / rstd (ln_c, ln_eps and the folded bias are used as above).
    const double* gn_in_stats = nullptr;
    int gn_in_rows = 0;
    // producer side (fp32 output + residual): a bf16 copy of the final rows (the next GEMM's A operand) and the partial
    // row statistics of the final rows: stat_out[row][(column / 128)] = (sum, sumsq) over that 128-column slab.
    void* out2 = nullptr;            // bf16 [M, ldo2]
    int ldo2 = 0;
    float2* stat_out = nullptr;      // [M][ceil(N / 128)]
    // CENTRED form of that hand-over (LayerNorm is invariant under a per-row shift of its input): the bf16 copy and the partial
    // statistics are taken of  x - c[row],  c[row] = the mean of the row BEFORE this GEMM's update = shift_in[row] + the mean the
    // previous producer's (already centred) statistics give.  The rounding error of the bf16 operand then scales with the row's
    // deviation from its previous mean instead of with |x| (a common offset per row costs nothing), and the single-pass variance
    // sum(y^2)/n - mean(y)^2 no longer cancels.  The consumer is unchanged (it only ever sees y = x - c).  shift_out[row] = c[row].
    // ---- implicit GEMM of a 3x3 / stride 1 / pad 1 convolution on rows (MSDeformAttnPixelDecoder.output_conv,
    // transformer_encoder_deform.py:259-268): A is the ZERO-PADDED bf16 map [batch, conv_h + 2, conv_w + 2, conv_c]; K = 9 conv_c is
    // ordered (ky, kx, c); the K block kb = (tap, channel block) of the output pixels m0.. is the box at padded row
    // (b (h + 2) + y + ky) (w + 2) + x0 + kx, so the TMA producer shifts rows per tap instead of reading a 9x larger im2col operand.
    // conv_w must be a multiple of the CTA's 128 rows (a tile never crosses an image row), conv_c of 64. ----
    int conv_w = 0, conv_h = 0, conv_c = 0;
    // ---- transposed fp32 output (streamlined pair kernel, EPI_F32 only): out[n * ldo + m] = acc[m, n] + bias[n].  For C = A W^T with
    // few W rows and many A rows (the mask logits `einsum("bqc,bchw->bqhw")`, xdecoder.py:459: 101 queries x 65536 positions) the
    // LONG dimension runs along M — no tile rows wasted on the 101-of-256 queries — and the result still lands query-major. ----
    int out_t = 0;
    const float2* shift_stats = nullptr;   // [M][shift_parts] statistics of the rows before the update (null: c = shift_in)
    const float* shift_in = nullptr;       // [M] their shifts (null: zeros); [shift_in_mod] indexed row % shift_in_mod when that is set
    int shift_in_mod = 0;
    float* shift_out = nullptr;            // [M] receives c (null: centring off, c = 0)
    int shift_parts = 0, shift_dim = 0;
};
__host__ __device__ __forceinline__ int epilogue_remap_tokens(const Epilogue& ep) { return ep.remap_g * (ep.remap_h ? ep.remap_h : ep.remap_g); }
__host__ __device__ __forceinline__ int epilogue_remap_padded(const Epilogue& ep) { return ep.remap_gp * (ep.remap_hp ? ep.remap_hp : ep.remap_gp); }
__host__ __device__ __forceinline__ size_t epilogue_out_row(const Epilogue& ep, int row) {
    if (ep.remap_g == 0) return (size_t)row;
    const int t = epilogue_remap_tokens(ep);
    const int b = row / t, r = row % t;
    return (size_t)b * epilogue_remap_padded(ep) + (size_t)(r / ep.remap_g) * ep.remap_gp + (r % ep.remap_g);
}

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue two fp32 operations per lane per instruction) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Exact-erf GELU of two values for the bf16 path (nn.GELU(), common.py:21-26): erf by Abramowitz & Stegun 7.1.25
// (|error| <= 2.5e-5, two orders below the 2^-9 relative rounding of the bf16 store that follows), packed fp32x2
// arithmetic, MUFU rcp / ex2 — about 9 issue slots per element instead of ~24 for erff().
//   a = |x|, t = 1 / (1 + p a / sqrt2), e = 0.5 (a1 t + a2 t^2 + a3 t^3) exp(-x^2 / 2) = 0.5 (1 - erf(a / sqrt2))
//   gelu(x) = 0.5 x + a (0.5 - e)
__device__ __forceinline__ void gelu_pair(float& x0, float& x1) {
    const f32x2 x = f2_pack(x0, x1);
    const f32x2 a = f2_pack(fabsf(x0), fabsf(x1));
    const f32x2 d = f2_fma(a, f2_pack(0.47047f * 0.70710678f, 0.47047f * 0.70710678f), f2_pack(1.0f, 1.0f));
    float d0, d1;
    f2_unpack(d, d0, d1);
    const f32x2 t = f2_pack(rcp_approx(d0), rcp_approx(d1));
    f32x2 p = f2_fma(f2_pack(0.5f * 0.7478556f, 0.5f * 0.7478556f), t, f2_pack(0.5f * -0.0958798f, 0.5f * -0.0958798f));
    p = f2_fma(p, t, f2_pack(0.5f * 0.3480242f, 0.5f * 0.3480242f));
    p = f2_mul(p, t);
    const f32x2 arg = f2_mul(f2_mul(x, f2_pack(-0.5f * 1.44269504f, -0.5f * 1.44269504f)), x);
    float g0, g1;
    f2_unpack(arg, g0, g1);
    const f32x2 e = f2_mul(p, f2_pack(ex2_approx_f(g0), ex2_approx_f(g1)));
    const f32x2 h = f2_add(f2_pack(0.5f, 0.5f), f2_mul(e, f2_pack(-1.0f, -1.0f)));   // 0.5 - e
    const f32x2 r = f2_fma(a, h, f2_mul(x, f2_pack(0.5f, 0.5f)));
    f2_unpack(r, x0, x1);
}
__device__ __forceinline__ float gelu_fast(float x) {
    float y = x, z = 0.f;
    gelu_pair(y, z);
    return y;
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- launch accounting / optional event profiler (profile.cu) ----
enum ProfCat { PC_GEMM = 0, PC_ATTN_WIN = 1, PC_ATTN_GLOBAL = 2, PC_NORM = 3, PC_OTHER = 4, PC_COUNT = 5 };
void count_launch(int n = 1);
void prof_begin(int cat, double flops, double bytes, cudaStream_t st);
void prof_end(cudaStream_t st);
// SVB_PDL (bit mask, A/B): which kernels of the encoder are launched with programmatic stream serialization (their set-up overlaps the
// previous kernel's tail; griddepcontrol.wait orders every memory access behind it): 1 the GEMMs, 2 the attention kernels, 4 the
// pad-row kernel; 0 = plain stream order.  Measured inside the ViT-H step (64 images, alternating runs on one box, profiles/r02_pdl):
// 0: 347.8 / 349.0 ms, 1: 345.4 / 346.8, 3: 345.2 / 346.3, 7: 349.5 / 348.9 (the small pad-row kernel as a dependent launch gives
// the whole gain back) -> default 3.
inline bool pdl_enabled(int which = 1) {
    static const int v = [] { const char* e = getenv("SVB_PDL"); return e ? atoi(e) : 3; }();
    return (v & which) != 0;
}

struct ProfScope {
    cudaStream_t st;
    ProfScope(int cat, double flops, double bytes, cudaStream_t s, int launches = 1) : st(s) {
        count_launch(launches);
        prof_begin(cat, flops, bytes, s);
    }
    ~ProfScope() { prof_end(st); }
};

// ---- launchers implemented in the .cu files (all asynchronous on `stream`) ----
// tcgen05 GEMM: C[M,N] = A[M,K] (bf16, row-major, lda) * W[N,K]^T (bf16, row-major, ldw), fp32 accumulate.
int gemm_bf16_tc(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep,
                 cudaStream_t stream);
// fp32 SIMT GEMM of the validation mode, same contract with fp32 operands.
int gemm_f32_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const Epilogue& ep,
                  cudaStream_t stream);

struct AttnParams {
    const void* qkv;        // [B*grid*grid, 3*D], element type T (token order y*grid+x)
    void* out;              // [B*grid*grid, D], element type T
    const float* rel_h;     // [2*ws-1, hd] fp32
    const float* rel_w;     // [2*ws-1, hd] fp32
    const float* qkv_bias;  // [3*D] fp32: pad tokens of edge windows have k = b_k, v = b_v
    int batch, grid, ws, heads, hd;   // square form: grid x grid tokens, ws x ws windows; ws == grid -> global attention
    // general form (scope row N3), used when grid_h > 0: grid_h x grid_w tokens, ws_h x ws_w windows (= the grid for global attention);
    // rel_h has 2*ws_h-1 rows, rel_w 2*ws_w-1 rows
    int grid_h = 0, grid_w = 0, ws_h = 0, ws_w = 0;
};
// fp32-math SIMT attention with decomposed rel-pos; T = float (validation) or bf16.
int attention_simt(const AttnParams& p, bool is_bf16, cudaStream_t stream);
// tcgen05 attention kernels (bf16 in/out), attention_tc.cu.  `qkv` is the qkv GEMM output: token order [B*64*64, 3D] for
// global attention (ws == grid), the PADDED window grid [B, 70, 70, 3D] (pad rows = qkv bias, see fill_pad_rows) for ws == 14.
// `rel_pack` is the bf16 table block written by pack_rel_table: 64 rows (h at 0, w at 32) for windows, 272 rows (h at 0,
// w at 144) for global attention, hd columns.
struct AttnTcParams {
    const bf16* qkv; bf16* out;
    const bf16* rel_pack;
    int batch, grid, ws, heads, hd;
    long long* phase_clocks = nullptr;   // optional [2][8] device counters: per-phase cycles of the softmax groups (debug)
    // scope row N3: token grids other than the trained one (both sides multiples of 32; 0 = square `grid`).  The windowed kernel only
    // needs the sizes; the global kernel then takes its rel-pos terms from tables in global memory (see attention_tc.cu):
    // bias_h [batch * grid_h * grid_w][heads][bias_ld] fp32 = q . rel_pos_h[j] for j < 2 grid_h - 1 (resized table), bias_w likewise.
    int grid_h = 0, grid_w = 0;
    const float* bias_h = nullptr;
    const float* bias_w = nullptr;
    int bias_ld_h = 0, bias_ld_w = 0;
};
int attention_tc_rel_rows(int ws, int grid);
int attention_tc_set_debug_buffer(void* mapped_device_ptr);
int pack_rel_table(const float* src, bf16* dst, int L, int hd, bool is_w, cudaStream_t stream);
// pad rows of the padded [B, gph, gpw, ld] tensor = bias (gh x gw real tokens per image)
int fill_pad_rows(bf16* qkv, const float* bias, int B, int gh, int gw, int gph, int gpw, int ld, cudaStream_t stream);
int attention_tc(const AttnTcParams& p, cudaStream_t stream);
int attention_window3(const AttnTcParams& p, cudaStream_t stream);   // experiments/attention_win3.cu
int attention_window5(const AttnTcParams& p, cudaStream_t stream);   // experiments/attention_win5.cu: two independent chains per SM
int attention_window6(const AttnTcParams& p, cudaStream_t stream);   // experiments/attention_win6.cu: split-row windowed kernel
int attention_global_ext(const AttnTcParams& p, cudaStream_t stream); // attention_ext.cu: global attention on other token grids
// A = zero-padded bf16 map, see Epilogue::conv_w (gemm_tc2.cu)
int gemm_conv3x3_bf16_tc(const bf16* padded, const bf16* W, int ldw, int batch, int h, int w, int cin, int cout, const Epilogue& ep, cudaStream_t stream);
// tcgen05 masked cross-attention of the X-Decoder layers (xattn_tc.cu)
int xattn_tc_supported(int dtype_bf16, int queries, int keys, int head_dim, const void* q, const void* k, const void* v, const void* mask,
                       const void* out, int batch, int heads);
int xattn_tc_launch(const bf16* q, const bf16* k, const bf16* v, const uint8_t* mask, bf16* out, float* workspace, int64_t workspace_floats,
                    int queries, int keys, int batch, int heads, cudaStream_t stream);

// elementwise / normalisation kernels (elementwise.cu)
int im2col_patch(const void* x, int x_dtype, void* out, bool out_bf16, int B, int C, int img_h, int img_w, int patch, cudaStream_t s);
// scope row N3: bicubic resize of pos_embed (image_encoder.py:124-132), linear resize of a rel_pos table (:319-330)
int resize_pos_embed(const float* src, float* dst, int h0, int w0, int h1, int w1, int D, cudaStream_t s);
// out[t] = mean_d(a[t, d]) + mean_d(b[d]): the expected row mean of patch_embed + pos_embed, the first producer's centring shift
int row_means(const float* a, const float* b, float* out, int rows, int D, cudaStream_t s);
int resize_rel_pos(const float* src, float* dst, int L0, int L1, int hd, cudaStream_t s);
// uint8 (C,h,w) images -> normalised, zero-padded to img x img, patch rows (the patch-embedding GEMM's A operand); `images`, `hs`,
// `ws` are HOST arrays (device pointers / sizes per image)
int stage_u8_patch(const uint8_t* const* images, const int* hs, const int* ws, const float* mean, const float* stdv, void* out, bool out_bf16,
                   int B, int C, int img, int patch, cudaStream_t s);
// out = LayerNorm(x [+ add]); when `add` (same element type as out) is given, x += add is written back first (fused residual)
int layernorm_rows(float* x, const void* add, const float* w, const float* b, void* out, bool out_bf16, int rows, int D, float eps,
                   cudaStream_t s);
int layernorm_post_rows(const float* x, const float* add, const float* w, const float* b, float* out, bf16* out_b, const float* pos,
                        int pos_rows, bf16* out_q, int rows, int D, float eps, cudaStream_t s);
int cast_and_space2depth(const float* x, void* xb, void* a32, bool out_bf16, int B, int gh, int gw, int D, cudaStream_t s);
int groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out, bool out_bf16,
                    long rows, int C, long rows_per_sample, float eps, int gelu, cudaStream_t s);
// final stage: GroupNorm(1,C)+GELU and (pixel-unshuffle) NHWC -> NCHW.  `levels` = number of 2x2 ConvT stages whose
// sub-pixel index is still folded into the row index (0, 1 or 2); base token grid gh x gw.
int groupnorm_apply_nchw(const float* x, const double* stats, const float* gamma, const float* beta, void* out, int out_dtype,
                         int B, int gh, int gw, int levels, int C, float eps, int gelu, cudaStream_t s);
// weight packing
int pack_cast(const float* src, void* dst, bool dst_bf16, long n, cudaStream_t s);
int pack_convT(const float* w /*Cin,Cout,2,2*/, void* dst /*[4*Cout, Cin]*/, bool dst_bf16, int Cin, int Cout, cudaStream_t s);
int pack_conv2x2(const float* w /*Cout,Cin,2,2*/, void* dst /*[Cout, 4*Cin]*/, bool dst_bf16, int Cin, int Cout, cudaStream_t s);
int pack_bias4(const float* b, float* dst, int C, cudaStream_t s);
// out = T(a [+ b]) elementwise, fp32 in
int add_cast(const float* a, const float* b, void* out, bool out_bf16, size_t n, cudaStream_t s);
// LayerNorm -> Linear fold (see Epilogue::ln_stats): Wg bf16 [N,K], colsum [N], bias_f [N] from the fp32 masters
int fold_layernorm(const float* W, const float* bias, const float* gamma, const float* beta, bf16* Wg, float* colsum, float* bias_f, int N,
                   int K, cudaStream_t s);
// true when gemm_bf16_tc dispatches to the CTA-pair kernel (the only one implementing the folded-LayerNorm epilogues)
bool gemm_bf16_tc_supports_fold();

}  // namespace svb
