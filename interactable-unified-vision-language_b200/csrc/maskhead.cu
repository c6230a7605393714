// Mask branch of the X-Decoder prediction heads (scope row N4, first slice): `XDecoder.forward_prediction_heads`,
// /root/reference/modeling/interface/xdecoder.py:429-470 — class-token recompute, (mask_embed MLP and the "bqc,bchw->bqhw" mask
// logits are GEMMs: svb_linear), antialiased bicubic resize of the logits to the attention-mask size, threshold + head repeat.
#include "../../include/samvit_b200.h"
#include "common.cuh"
#include <cstdlib>

namespace svb {
namespace {

// ---- xdecoder.py:440-446: cls = sum_j softmax_j(<x_cls / |x_cls|, x_j / |x_j|>) x_j over the object tokens j < Q-1; row Q-1 <- cls ----
__global__ void cls_token_kernel(float* __restrict__ x, int Q, int C) {
    extern __shared__ float sh[];                          // [Q] similarities, then softmax weights
    float* xb = x + (size_t)blockIdx.x * Q * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* xc = xb + (size_t)(Q - 1) * C;
    float ncls = 0.f;
    for (int c = lane; c < C; c += 32) ncls += xc[c] * xc[c];
    ncls = sqrtf(warp_sum(ncls)) + 1e-7f;
    for (int j = warp; j < Q - 1; j += nw) {
        const float* xj = xb + (size_t)j * C;
        float dot = 0.f, nj = 0.f;
        for (int c = lane; c < C; c += 32) { const float v = xj[c]; dot += v * xc[c]; nj += v * v; }
        dot = warp_sum(dot);
        nj = sqrtf(warp_sum(nj)) + 1e-7f;
        if (lane == 0) sh[j] = dot / (nj * ncls);
    }
    __syncthreads();
    if (warp == 0) {                                       // softmax over the Q-1 object tokens
        float m = -INFINITY;
        for (int j = lane; j < Q - 1; j += 32) m = fmaxf(m, sh[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int j = lane; j < Q - 1; j += 32) { const float e = expf(sh[j] - m); sh[j] = e; s += e; }
        s = warp_sum(s);
        for (int j = lane; j < Q - 1; j += 32) sh[j] /= s;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < Q - 1; ++j) acc += sh[j] * xb[(size_t)j * C + c];
        xb[(size_t)(Q - 1) * C + c] = acc;                 // written after every read of the old class row (the barrier above)
    }
}

// ---- F.interpolate(mode="bicubic", align_corners=False, antialias=True): separable, one pass per axis ----
// PyTorch's `upsample_bicubic2d_aa` weights: scale = in / out, support = 2 * max(scale, 1), centre = scale * (i + 0.5), taps
// [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the input, w = cubic((j - centre + 0.5) / max(scale, 1)),
// normalised to sum 1; cubic convolution filter with a = -0.5.
__device__ __forceinline__ float cubic_aa(float x) {
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
    if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
    return 0.f;
}
// out[m, o, i] = sum_j w_j src[m, j, i]  (AXIS_ROWS) or out[m, r, o] = sum_j w_j src[m, r, j]  (columns); `in` -> `out` along the axis
template <bool AXIS_ROWS>
__global__ void resize_aa_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int in, int out) {
    const int m = blockIdx.y;
    const int orows = AXIS_ROWS ? out : rows, ocols = AXIS_ROWS ? cols : out;
    const float scale = (float)in / (float)out;
    const float support = scale >= 1.f ? 2.f * scale : 2.f, invscale = scale >= 1.f ? 1.f / scale : 1.f;
    const float* s = src + (size_t)m * rows * cols;
    float* d = dst + (size_t)m * orows * ocols;
    const size_t total = (size_t)orows * ocols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int oc = (int)(i % ocols), orow = (int)(i / ocols);
        const int o = AXIS_ROWS ? orow : oc;
        const float center = scale * (o + 0.5f);
        const int lo = max((int)(center - support + 0.5f), 0), hi = min((int)(center + support + 0.5f), in);
        float acc = 0.f, wsum = 0.f;
        for (int j = lo; j < hi; ++j) {
            const float w = cubic_aa((j - center + 0.5f) * invscale);
            const float v = AXIS_ROWS ? __ldg(s + (size_t)j * cols + oc) : __ldg(s + (size_t)orow * cols + j);
            acc += w * v;
            wsum += w;
        }
        d[i] = acc / wsum;
    }
}

// The same two passes with the filter taken out of the inner loop (the kernel above evaluates the cubic for every tap of every output
// element: 33 evaluations per output at the 8x reduction, identical for all rows of a map and all maps; it remains for shapes whose
// tiles do not fit shared memory).  A block first writes the normalised taps of the output indexes it needs into shared memory
// (lo / count / weights, as PyTorch's `upsample_bicubic2d_aa` normalises them before the sum), then:
//   columns (w -> out): a block owns 32 rows of one map, staged once with coalesced 16-byte loads into a padded tile; lane = row,
//     a warp walks the output columns, so the weights are shared-memory broadcasts and the data reads are conflict-free; the
//     32 x out results leave through a second padded tile as whole rows;
//   rows (h -> out): thread = (output row, column) with consecutive lanes on consecutive columns (coalesced, the weights broadcast).
// Shared by both: taps of output index o along an axis of `in` -> `out` elements.
__device__ __forceinline__ void aa_taps(int o, int in, float scale, float support, float invscale, int& lo, int& n, float& center) {
    center = scale * (o + 0.5f);
    lo = max((int)(center - support + 0.5f), 0);
    n = min((int)(center + support + 0.5f), in) - lo;
}
// table for output indexes [o0, o0 + no): wt[(o - o0) * taps_max + t], lo_s / n_s[o - o0]; call with all threads of the block
__device__ __forceinline__ void aa_build_table(float* wt, int* lo_s, int* n_s, int o0, int no, int in, int out, int taps_max) {
    const float scale = (float)in / (float)out;
    const float support = scale >= 1.f ? 2.f * scale : 2.f, invscale = scale >= 1.f ? 1.f / scale : 1.f;
    for (int i = threadIdx.x; i < no * taps_max; i += blockDim.x) {
        const int ol = i / taps_max, t = i - ol * taps_max;
        int lo, n;
        float center;
        aa_taps(o0 + ol, in, scale, support, invscale, lo, n, center);
        wt[i] = t < n ? cubic_aa((lo + t - center + 0.5f) * invscale) : 0.f;
        if (t == 0) { lo_s[ol] = lo; n_s[ol] = min(n, taps_max); }
    }
    __syncthreads();
    for (int ol = threadIdx.x; ol < no; ol += blockDim.x) {
        float wsum = 0.f;
        for (int t = 0; t < n_s[ol]; ++t) wsum += wt[ol * taps_max + t];
        const float inv = 1.f / wsum;
        for (int t = 0; t < n_s[ol]; ++t) wt[ol * taps_max + t] *= inv;
    }
    __syncthreads();
}

constexpr int RA_ROWS = 32;
__global__ void __launch_bounds__(256)
resize_aa_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int out, int taps_max) {
    extern __shared__ float ra_sm[];
    float* tile = ra_sm;                                   // [32][cols + 1]
    float* so = tile + RA_ROWS * (cols + 1);               // [32][out + 1]
    float* wt = so + RA_ROWS * (out + 1);                  // [out][taps_max]
    int* lo_s = reinterpret_cast<int*>(wt + out * taps_max);
    int* n_s = lo_s + out;
    const int m = blockIdx.y, r0 = blockIdx.x * RA_ROWS;
    const float* s = src + ((size_t)m * rows + r0) * cols;
    const int nr = min(RA_ROWS, rows - r0);
    if ((cols & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int c4 = cols >> 2;
        for (int i = threadIdx.x; i < nr * c4; i += blockDim.x) {
            const int r = i / c4, c = (i - r * c4) * 4;
            const float4 v = __ldcs(reinterpret_cast<const float4*>(s + (size_t)r * cols + c));
            float* t = tile + r * (cols + 1) + c;
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        }
    } else {
        for (int i = threadIdx.x; i < nr * cols; i += blockDim.x) {
            const int r = i / cols, c = i - r * cols;
            tile[r * (cols + 1) + c] = __ldcs(s + (size_t)r * cols + c);
        }
    }
    aa_build_table(wt, lo_s, n_s, 0, out, cols, out, taps_max);      // (its barriers also cover the tile)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < nr) {
        const float* trow = tile + lane * (cols + 1);
        for (int o = warp; o < out; o += 8) {
            const float* w = wt + o * taps_max;
            const float* tp = trow + lo_s[o];
            const int n = n_s[o];
            float acc = 0.f;
            for (int t = 0; t < n; ++t) acc = fmaf(w[t], tp[t], acc);
            so[lane * (out + 1) + o] = acc;
        }
    }
    __syncthreads();
    float* d = dst + ((size_t)m * rows + r0) * out;
    for (int i = threadIdx.x; i < nr * out; i += blockDim.x) {
        const int r = i / out, o = i - r * out;
        d[i] = so[r * (out + 1) + o];
    }
}

constexpr int RA_PER_BLOCK = 2048;                         // outputs per block of the row pass
__global__ void __launch_bounds__(256)
resize_aa_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int out, int taps_max) {
    extern __shared__ float ra_sm[];
    const int m = blockIdx.y;
    const int total = out * cols;
    const int i0 = blockIdx.x * RA_PER_BLOCK, i1 = min(total, i0 + RA_PER_BLOCK);
    const int o0 = i0 / cols, no = (i1 - 1) / cols - o0 + 1;
    float* wt = ra_sm;                                     // [no][taps_max]
    int* lo_s = reinterpret_cast<int*>(wt + no * taps_max);
    int* n_s = lo_s + no;
    aa_build_table(wt, lo_s, n_s, o0, no, rows, out, taps_max);
    const float* s = src + (size_t)m * rows * cols;
    float* d = dst + (size_t)m * total;
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int o = i / cols, c = i - o * cols, ol = o - o0;
        const float* w = wt + ol * taps_max;
        const float* sp = s + (size_t)lo_s[ol] * cols + c;
        const int n = n_s[ol];
        float acc = 0.f;
        for (int t = 0; t < n; ++t) acc = fmaf(w[t], __ldg(sp + (size_t)t * cols), acc);
        d[i] = acc;
    }
}

// ---- xdecoder.py:467: (sigmoid(v) < 0.5) == (v < 0), repeated over the heads: out[b, h, :] = v[b, :] < 0 ----
__global__ void mask_threshold_kernel(const float* __restrict__ v, uint8_t* __restrict__ out, int heads, size_t per_sample) {
    const int b = blockIdx.y;
    const float* vb = v + (size_t)b * per_sample;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_sample; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t bit = vb[i] < 0.f ? 1 : 0;
        for (int h = 0; h < heads; ++h) out[((size_t)b * heads + h) * per_sample + i] = bit;
    }
}

// The same with xdecoder.py:267 folded in (`attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False` at the top of the
// next decoder layer): one block per (image, query) row; a row whose every key is masked is written as all-False.  The threshold
// kernel above + svb_mask_clear_full_rows write the 8 per-head copies and read them back; here the row's logits are read twice (the
// second time out of L1 / L2) and the copies written once.
__global__ void __launch_bounds__(256)
mask_threshold_clear_kernel(const float* __restrict__ v, uint8_t* __restrict__ out, int heads, int queries, int keys) {
    __shared__ int any_open;
    const int b = blockIdx.y, q = blockIdx.x;
    const float* row = v + ((size_t)b * queries + q) * keys;
    if (threadIdx.x == 0) any_open = 0;
    __syncthreads();
    int open = 0;
    for (int i = threadIdx.x; i < keys; i += blockDim.x) open |= (__ldg(row + i) < 0.f) ? 0 : 1;      // (NaN counts as open, as `<` does)
    if (__any_sync(0xffffffffu, open) && (threadIdx.x & 31) == 0) atomicOr(&any_open, 1);
    __syncthreads();
    const bool full = any_open == 0;
    const bool vec = (keys & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (vec) {
        for (int i = threadIdx.x * 16; i < keys; i += blockDim.x * 16) {
            uint32_t w[4] = {0, 0, 0, 0};
            if (!full) {
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j >> 2] |= (uint32_t)(__ldg(row + i + j) < 0.f ? 1 : 0) << (8 * (j & 3));
            }
            const uint4 u = make_uint4(w[0], w[1], w[2], w[3]);
            for (int h = 0; h < heads; ++h) *reinterpret_cast<uint4*>(out + (((size_t)b * heads + h) * queries + q) * keys + i) = u;
        }
    } else {
        for (int i = threadIdx.x; i < keys; i += blockDim.x) {
            const uint8_t bit = (!full && __ldg(row + i) < 0.f) ? 1 : 0;
            for (int h = 0; h < heads; ++h) out[(((size_t)b * heads + h) * queries + q) * keys + i] = bit;
        }
    }
}

inline int grid_cap(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    const size_t cap = 148 * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace
}  // namespace svb

using namespace svb;

extern "C" int svb_cls_token_recompute(float* x, int batch, int queries, int channels, svb_stream_t stream) {
    SVB_REQUIRE(x && batch > 0 && queries > 1 && channels > 0, "svb_cls_token_recompute: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OTHER, 0, (double)batch * queries * channels * 8, s);
    cls_token_kernel<<<batch, 256, sizeof(float) * queries, s>>>(x, queries, channels);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_resize_bicubic_aa(const float* src, float* tmp, float* dst, int maps, int h, int w, int out_h, int out_w, svb_stream_t stream) {
    SVB_REQUIRE(src && tmp && dst && maps > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, "svb_resize_bicubic_aa: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OTHER, 0, (double)maps * ((double)h * w + 2.0 * h * out_w + (double)out_h * out_w) * 4, s, 2);
    // columns first (w -> out_w), then rows (h -> out_h), as the reference's separable CPU kernel orders them
    auto taps_of = [](int in, int out) {
        const float scale = (float)in / (float)out;
        return (int)ceilf(2.f * (scale >= 1.f ? 2.f * scale : 2.f)) + 2;
    };
    static const int legacy = [] { const char* e = getenv("SVB_RESIZE_LEGACY"); return e ? atoi(e) : 0; }();   // 1: the per-tap kernels (A/B)
    const int t1 = taps_of(w, out_w), t2 = taps_of(h, out_h);
    const size_t sm1 = sizeof(float) * ((size_t)RA_ROWS * (w + 1) + (size_t)RA_ROWS * (out_w + 1) + (size_t)out_w * t1) + sizeof(int) * 2 * (size_t)out_w;
    const int no2 = (RA_PER_BLOCK + out_w - 1) / out_w + 1;
    const size_t sm2 = sizeof(float) * (size_t)no2 * t2 + sizeof(int) * 2 * (size_t)no2;
    if (!legacy && sm1 <= 96 * 1024 && maps <= 65535) {
        static bool attr = false;
        if (!attr) {
            SVB_CHECK_CUDA(cudaFuncSetAttribute(resize_aa_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr = true;
        }
        dim3 g1((h + RA_ROWS - 1) / RA_ROWS, maps);
        resize_aa_cols_kernel<<<g1, 256, sm1, s>>>(src, tmp, h, w, out_w, t1);
    } else {
        dim3 g1(grid_cap((size_t)h * out_w, 256), maps);
        resize_aa_kernel<false><<<g1, 256, 0, s>>>(src, tmp, h, w, w, out_w);
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    if (!legacy && sm2 <= 48 * 1024 && maps <= 65535) {
        dim3 g2(((size_t)out_h * out_w + RA_PER_BLOCK - 1) / RA_PER_BLOCK, maps);
        resize_aa_rows_kernel<<<g2, 256, sm2, s>>>(tmp, dst, h, out_w, out_h, t2);
    } else {
        dim3 g2(grid_cap((size_t)out_h * out_w, 256), maps);
        resize_aa_kernel<true><<<g2, 256, 0, s>>>(tmp, dst, h, out_w, h, out_h);
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_mask_threshold_heads(const float* v, void* out_bool, int batch, int heads, int64_t per_sample, svb_stream_t stream) {
    SVB_REQUIRE(v && out_bool && batch > 0 && heads > 0 && per_sample > 0, "svb_mask_threshold_heads: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OTHER, 0, (double)batch * per_sample * (4 + heads), s);
    dim3 grid(grid_cap((size_t)per_sample, 256), batch);
    mask_threshold_kernel<<<grid, 256, 0, s>>>(v, (uint8_t*)out_bool, heads, (size_t)per_sample);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_mask_threshold_heads_clear(const float* v, void* out_bool, int batch, int heads, int queries, int keys, svb_stream_t stream) {
    SVB_REQUIRE(v && out_bool && batch > 0 && heads > 0 && queries > 0 && keys > 0 && batch <= 65535, "svb_mask_threshold_heads_clear: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OTHER, 0, (double)batch * queries * keys * (4 + heads), s);
    dim3 grid(queries, batch);
    mask_threshold_clear_kernel<<<grid, 256, 0, s>>>(v, (uint8_t*)out_bool, heads, queries, keys);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================================================
// Masked cross-attention core (scope row N4, second slice): the scaled-dot-product attention inside `nn.MultiheadAttention` as
// `CrossAttentionLayer.forward_post` calls it (/root/reference/modeling/interface/modules.py:95-106): Q <= 128 query tokens
// attend to HW image positions under a boolean mask (True = not allowed, the `attn_mask` of xdecoder.py:467).  ~100 queries per
// image make this a small, key-parallel problem (27 GFLOP per layer at 8 images x 128^2 positions): one thread per query with
// its 64-channel q and output rows in registers, K / V chunks broadcast from shared memory, keys split over blocks
// (flash-decoding style partial results) and a combine kernel.
// ================================================================================================================
namespace svb {
namespace {
constexpr int XA_HD = 64, XA_CHUNK = 64;

template <typename T>
__global__ void __launch_bounds__(128)
xattn_partial_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const uint8_t* __restrict__ mask,
                     float* __restrict__ part, int Q, int HW, int B, int heads, int keys_per_block, float scale) {
    __shared__ float sk[XA_CHUNK][XA_HD], sv[XA_CHUNK][XA_HD];
    const int blk = blockIdx.x, head = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
    const int C = heads * XA_HD;
    const int k0 = blk * keys_per_block, k1 = min(HW, k0 + keys_per_block);
    float qr[XA_HD], o[XA_HD];
    const bool live = t < Q;
#pragma unroll
    for (int d = 0; d < XA_HD; ++d) { qr[d] = live ? to_float(q[((size_t)t * B + b) * C + head * XA_HD + d]) * scale : 0.f; o[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    const uint8_t* mrow = mask ? mask + ((size_t)(b * heads + head) * Q + (live ? t : 0)) * HW : nullptr;
    for (int c0 = k0; c0 < k1; c0 += XA_CHUNK) {
        const int nk = min(XA_CHUNK, k1 - c0);
        __syncthreads();
        for (int i = t; i < XA_CHUNK * XA_HD; i += 128) {
            const int kk = i / XA_HD, d = i % XA_HD;
            const size_t src = ((size_t)(c0 + kk) * B + b) * C + head * XA_HD + d;
            sk[kk][d] = kk < nk ? to_float(k[src]) : 0.f;
            sv[kk][d] = kk < nk ? to_float(v[src]) : 0.f;
        }
        __syncthreads();
        if (!live) continue;
        // 16 keys at a time, fully unrolled: the scores stay in registers (a 64-entry score array went to local memory)
#pragma unroll 1
        for (int k16 = 0; k16 < XA_CHUNK; k16 += 16) {
            if (k16 >= nk) break;
            float s[16];
            float cmax = -INFINITY;
            uint4 mb = make_uint4(0, 0, 0, 0);
            if (mrow) {
                if ((reinterpret_cast<uintptr_t>(mrow + c0 + k16) & 15) == 0 && k16 + 16 <= nk) mb = *reinterpret_cast<const uint4*>(mrow + c0 + k16);   // 16 mask bytes at once (an aligned address, whatever the view's storage offset)
                else {
                    uint8_t tmpb[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) tmpb[j] = (k16 + j < nk) ? mrow[c0 + k16 + j] : 1;
                    mb = *reinterpret_cast<uint4*>(tmpb);
                }
            }
            const uint32_t mw[4] = {mb.x, mb.y, mb.z, mb.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < XA_HD; ++d) acc = fmaf(qr[d], sk[k16 + j][d], acc);
                const bool off = (k16 + j >= nk) || ((mw[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                s[j] = off ? -INFINITY : acc;
                cmax = fmaxf(cmax, s[j]);
            }
            if (cmax == -INFINITY) continue;               // every key of this group is masked for this query
            const float m_new = fmaxf(m, cmax);
            if (m_new > m) {
                const float alpha = __expf(m - m_new);     // 0 when m was -inf
                l *= alpha;
#pragma unroll
                for (int d = 0; d < XA_HD; ++d) o[d] *= alpha;
                m = m_new;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float p = __expf(s[j] - m);          // exp(-inf) = 0 for masked keys
                l += p;
#pragma unroll
                for (int d = 0; d < XA_HD; ++d) o[d] = fmaf(p, sv[k16 + j][d], o[d]);
            }
        }
    }
    if (live) {
        float* dst = part + ((((size_t)b * heads + head) * gridDim.x + blk) * Q + t) * (XA_HD + 2);
        dst[0] = m;
        dst[1] = l;
#pragma unroll
        for (int d = 0; d < XA_HD; ++d) dst[2 + d] = o[d];
    }
}

template <typename T>
__global__ void xattn_combine_kernel(const float* __restrict__ part, T* __restrict__ out, int Q, int B, int heads, int nblk) {
    const int head = blockIdx.y, b = blockIdx.z;
    const int t = blockIdx.x * (blockDim.x / XA_HD) + threadIdx.x / XA_HD, d = threadIdx.x % XA_HD;
    if (t >= Q) return;
    const float* p0 = part + (((size_t)b * heads + head) * nblk * Q + t) * (XA_HD + 2);
    const size_t stride = (size_t)Q * (XA_HD + 2);
    float m = -INFINITY;
    for (int j = 0; j < nblk; ++j) m = fmaxf(m, p0[j * stride]);
    float l = 0.f, o = 0.f;
    for (int j = 0; j < nblk; ++j) {
        const float w = __expf(p0[j * stride] - m);        // NaN only if every key of the row is masked (m = -inf), as in torch
        l += w * p0[j * stride + 1];
        o += w * p0[j * stride + 2 + d];
    }
    out[((size_t)t * B + b) * (heads * XA_HD) + head * XA_HD + d] = from_float<T>(o / l);
}
}  // namespace
}  // namespace svb

extern "C" int svb_masked_cross_attention(const void* q, const void* k, const void* v, int dtype, const void* mask_bool, void* out, float* workspace,
                                          int64_t workspace_floats, int queries, int keys, int batch, int heads, int head_dim, svb_stream_t stream) {
    SVB_REQUIRE(q && k && v && out && workspace, "svb_masked_cross_attention: null argument");
    SVB_REQUIRE(head_dim == XA_HD && queries >= 1 && queries <= 128 && keys >= 1 && batch >= 1 && heads >= 1,
                "svb_masked_cross_attention: head_dim %d (64 supported), %d queries (<= 128)", head_dim, queries);
    cudaStream_t s = (cudaStream_t)stream;
    // bf16, keys in multiples of 64: the tcgen05 kernel (xattn_tc.cu).  SVB_XATTN_IMPL=0 forces the fp32-FMA kernel below (A/B, fp32 mode).
    static const bool tc_off = [] { const char* e = getenv("SVB_XATTN_IMPL"); return e && atoi(e) == 0; }();
    if (!tc_off && xattn_tc_supported(dtype == SVB_DTYPE_BF16, queries, keys, head_dim, q, k, v, mask_bool, out, batch, heads))
        return xattn_tc_launch((const bf16*)q, (const bf16*)k, (const bf16*)v, (const uint8_t*)mask_bool, (bf16*)out, workspace, workspace_floats,
                               queries, keys, batch, heads, s);
    int kpb = 512;
    while (kpb > XA_CHUNK && (long)((keys + kpb - 1) / kpb) * heads * batch < 296) kpb /= 2;      // at least two blocks per SM where possible
    const int nblk = (keys + kpb - 1) / kpb;
    const int64_t need = (int64_t)batch * heads * nblk * queries * (XA_HD + 2);
    SVB_REQUIRE(workspace_floats >= need, "svb_masked_cross_attention: workspace of %lld floats needed, %lld given", (long long)need,
                (long long)workspace_floats);
    ProfScope prof(PC_OTHER, 4.0 * batch * heads * (double)queries * keys * XA_HD, (double)batch * keys * heads * XA_HD * 2 * (dtype == SVB_DTYPE_BF16 ? 2 : 4), s, 2);
    dim3 g1(nblk, heads, batch);
    const float scale = 1.0f / sqrtf((float)head_dim);
    dim3 g2((queries + 3) / 4, heads, batch);
    if (dtype == SVB_DTYPE_BF16) {
        xattn_partial_kernel<bf16><<<g1, 128, 0, s>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const uint8_t*)mask_bool, workspace, queries, keys, batch, heads, kpb, scale);
        SVB_CHECK_CUDA(cudaGetLastError());
        xattn_combine_kernel<bf16><<<g2, 256, 0, s>>>(workspace, (bf16*)out, queries, batch, heads, nblk);
    } else {
        xattn_partial_kernel<float><<<g1, 128, 0, s>>>((const float*)q, (const float*)k, (const float*)v, (const uint8_t*)mask_bool, workspace, queries, keys, batch, heads, kpb, scale);
        SVB_CHECK_CUDA(cudaGetLastError());
        xattn_combine_kernel<float><<<g2, 256, 0, s>>>(workspace, (float*)out, queries, batch, heads, nblk);
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int64_t svb_masked_cross_attention_workspace(int queries, int keys, int batch, int heads) {
    return (int64_t)batch * heads * ((keys + XA_CHUNK - 1) / XA_CHUNK) * queries * (XA_HD + 2);
}

// LanguageEncoder.compute_similarity's normalisation (modeling/language/vlpencoder.py:242, 244): out = scale * x / (|x| + eps) per row
namespace svb {
namespace {
template <typename T>
__global__ void l2_normalize_rows_kernel(const float* __restrict__ x, T* __restrict__ out, int rows, int dim, float eps, float scale) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + (size_t)row * dim;
    float s = 0.f;
    for (int d = lane; d < dim; d += 32) s = fmaf(xr[d], xr[d], s);
    s = warp_sum(s);
    const float f = scale / (sqrtf(s) + eps);
    for (int d = lane; d < dim; d += 32) out[(size_t)row * dim + d] = from_float<T>(xr[d] * f);
}
}  // namespace
}  // namespace svb

extern "C" int svb_l2_normalize_rows(const float* x, void* out, int out_dtype, int rows, int dim, float eps, float scale, svb_stream_t stream) {
    SVB_REQUIRE(x && out && rows > 0 && dim > 0, "svb_l2_normalize_rows: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OTHER, 0, (double)rows * dim * (4 + (out_dtype == SVB_DTYPE_BF16 ? 2 : 4)), s);
    if (out_dtype == SVB_DTYPE_BF16) l2_normalize_rows_kernel<bf16><<<(rows + 7) / 8, 256, 0, s>>>(x, (bf16*)out, rows, dim, eps, scale);
    else l2_normalize_rows_kernel<float><<<(rows + 7) / 8, 256, 0, s>>>(x, (float*)out, rows, dim, eps, scale);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// xdecoder.py:258: attn_mask[where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False — a query whose every key is masked attends to all
namespace svb {
namespace {
__global__ void mask_clear_full_rows_kernel(uint8_t* __restrict__ mask, int keys) {
    uint8_t* row = mask + (size_t)blockIdx.x * keys;
    int all = 1;
    for (int i = threadIdx.x; i < keys; i += blockDim.x) all &= (row[i] != 0);
    all = __syncthreads_and(all);
    if (all) for (int i = threadIdx.x; i < keys; i += blockDim.x) row[i] = 0;
}
}  // namespace
}  // namespace svb

extern "C" int svb_mask_clear_full_rows(void* mask_bool, int64_t rows, int keys, svb_stream_t stream) {
    SVB_REQUIRE(mask_bool && rows > 0 && keys > 0, "svb_mask_clear_full_rows: bad argument");
    ProfScope prof(PC_OTHER, 0, (double)rows * keys, (cudaStream_t)stream);
    mask_clear_full_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((uint8_t*)mask_bool, keys);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
