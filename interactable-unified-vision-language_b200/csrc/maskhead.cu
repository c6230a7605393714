// Mask branch of the X-Decoder prediction heads (scope row N4, first slice): `XDecoder.forward_prediction_heads`,
// /root/reference/modeling/interface/xdecoder.py:429-470 — class-token recompute, (mask_embed MLP and the "bqc,bchw->bqhw" mask
// logits are GEMMs: svb_linear), antialiased bicubic resize of the logits to the attention-mask size, threshold + head repeat.
#include "../../include/samvit_b200.h"
#include "common.cuh"

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

// ---- xdecoder.py:467: (sigmoid(v) < 0.5) == (v < 0), repeated over the heads: out[b, h, :] = v[b, :] < 0 ----
__global__ void mask_threshold_kernel(const float* __restrict__ v, uint8_t* __restrict__ out, int heads, size_t per_sample) {
    const int b = blockIdx.y;
    const float* vb = v + (size_t)b * per_sample;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_sample; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t bit = vb[i] < 0.f ? 1 : 0;
        for (int h = 0; h < heads; ++h) out[((size_t)b * heads + h) * per_sample + i] = bit;
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
    dim3 g1(grid_cap((size_t)h * out_w, 256), maps);
    resize_aa_kernel<false><<<g1, 256, 0, s>>>(src, tmp, h, w, w, out_w);
    SVB_CHECK_CUDA(cudaGetLastError());
    dim3 g2(grid_cap((size_t)out_h * out_w, 256), maps);
    resize_aa_kernel<true><<<g2, 256, 0, s>>>(tmp, dst, h, out_w, h, out_h);
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
