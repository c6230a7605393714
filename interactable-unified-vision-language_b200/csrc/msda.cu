// Multi-scale deformable attention, forward sampling core (scope row N1: the first consumer of res3..res5).
//
// Restates MSDA.ms_deform_attn_forward — the reference's only native kernel, ms_deformable_im2col_gpu_kernel
// (modeling/vision/encoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:242-303, bilinear tap :18-69; sm_86 SIMT, fp32 only, one
// thread per output CHANNEL with scalar loads) — as it is called by MSDeformAttn.forward (ops/modules/ms_deform_attn.py:117-119):
//   out[n, q, m, :] = sum_{l, p} attn[n, q, m, l, p] * bilinear(value[n, level l, :, m, :], loc[n, q, m, l, p])
// with grid_sample(align_corners=False, padding_mode='zeros') semantics (ops/functions/ms_deform_attn_func.py:52-72).
//
// B200 design: the op is a gather out of `value` (22 MB per 1024^2 image in bf16 at the step1.yaml geometry: 3 levels, 8 heads x 64
// channels — L2-resident).  One thread owns 16 BYTES of channels (8 bf16 / 4 fp32) of one (n, q, head): the 4 taps of a sample are
// four 16-byte loads, consecutive threads cover consecutive channels (a 64-channel head = 8 lanes = one 128-byte line per tap),
// accumulation is fp32, and bf16 values halve the gather traffic of the reference's fp32-only kernel (which forces the caller to
// up-cast all levels, transformer_encoder_deform.py:314-345).  ncu showed the first form of the kernel bound by INSTRUCTIONS, not by
// the gather (issue slots 70 % busy, L1 29 %, L2 19 %, L1 / L2 hit rates 49 / 79 %): every lane of a unit repeated the softmax, the
// sampling-location and the bilinear-weight arithmetic for itself.  msda_fused_shared_kernel shares it (each lane prepares its share
// of the unit's points into shared memory; the accumulation loop reads a point as two 16-byte broadcasts): 0.85 -> 0.53 ms per
// MSDeformAttn layer and 10.6-11.2 -> 9.1 ms per pixel-decoder forward for 4 images; the one-lane-does-all kernels remain for head
// widths that are not a power-of-two number of 16-byte groups.
#include "../../include/samvit_b200.h"
#include "common.cuh"

#include <cstdlib>

namespace svb {
namespace {

constexpr int MSDA_MAX_LEVELS = 8;
struct MsdaLevels {
    int h[MSDA_MAX_LEVELS], w[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS];
};

template <typename T> struct Pack16;
template <> struct Pack16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Pack16<bf16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);               // bf16 -> fp32 is a 16-bit shift
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};

// PT = number of points when known at compile time (4: the reference's configuration, fully unrolled), 0 = run-time loop.
// The body is branch-free: taps outside the map get weight 0 and a clamped (always valid) address, so the 4 x PT loads of a
// level are independent and can all be in flight together (the gather is latency-bound otherwise).
template <typename T, int PT>
__global__ void __launch_bounds__(256)
msda_forward_kernel(const T* __restrict__ value, const float* __restrict__ loc, const float* __restrict__ attn, T* __restrict__ out,
                    const __grid_constant__ MsdaLevels lv, long total, int S, int M, int D, int L, int Q, int P) {
    constexpr int V = Pack16<T>::N;
    const int groups = D / V;                                     // 16-byte channel groups per head
    const int np = PT ? PT : P;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % groups);
        const long unit = idx / groups;                           // (n, q, m) flattened: the reference's `sampling_index`
        const int m = (int)(unit % M);
        const long nq = unit / M;
        const int n = (int)(nq / Q);
        const float* wp = attn + unit * L * np;
        const float2* lp = reinterpret_cast<const float2*>(loc) + unit * L * np;
        const size_t head_stride = (size_t)M * D;                 // elements between consecutive spatial positions
        const T* vbase = value + (size_t)n * S * head_stride + (size_t)m * D + (size_t)cg * V;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        for (int l = 0; l < L; ++l) {
            const int H = lv.h[l], W = lv.w[l];
            const T* vl = vbase + (size_t)lv.start[l] * head_stride;
#pragma unroll
            for (int p = 0; p < (PT ? PT : 1); ++p) {
                for (int pr = 0; pr < (PT ? 1 : np); ++pr) {      // run-time point loop only in the generic instantiation
                    const float2 xy = __ldg(lp);
                    const float wt = __ldg(wp);
                    ++lp;
                    ++wp;
                    const float h_im = xy.y * H - 0.5f, w_im = xy.x * W - 0.5f;      // ms_deform_im2col_cuda.cuh:286-287
                    const bool inside = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;     // :289
                    const float hf = floorf(h_im), wf = floorf(w_im);
                    const int h_low = (int)hf, w_low = (int)wf;
                    const int h_high = h_low + 1, w_high = w_low + 1;
                    const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
                    const bool hl = inside && h_low >= 0, hhi = inside && h_high <= H - 1;
                    const bool wl = w_low >= 0, whi = w_high <= W - 1;
                    const float w1 = (hl && wl) ? hh * hw * wt : 0.f, w2 = (hl && whi) ? hh * lw * wt : 0.f;
                    const float w3 = (hhi && wl) ? lh * hw * wt : 0.f, w4 = (hhi && whi) ? lh * lw * wt : 0.f;
                    const int y0 = min(max(h_low, 0), H - 1), y1 = min(max(h_high, 0), H - 1);
                    const int x0 = min(max(w_low, 0), W - 1), x1 = min(max(w_high, 0), W - 1);
                    float t1[V], t2[V], t3[V], t4[V];
                    Pack16<T>::load(vl + ((size_t)y0 * W + x0) * head_stride, t1);
                    Pack16<T>::load(vl + ((size_t)y0 * W + x1) * head_stride, t2);
                    Pack16<T>::load(vl + ((size_t)y1 * W + x0) * head_stride, t3);
                    Pack16<T>::load(vl + ((size_t)y1 * W + x1) * head_stride, t4);
#pragma unroll
                    for (int i = 0; i < V; ++i) acc[i] = fmaf(w1, t1[i], fmaf(w2, t2[i], fmaf(w3, t3[i], fmaf(w4, t4[i], acc[i]))));
                }
            }
        }
        Pack16<T>::store(out + unit * D + (size_t)cg * V, acc);
    }
}

// The same gather with the head of MSDeformAttn.forward fused in (ops/modules/ms_deform_attn.py:97-112): instead of materialised
// sampling_locations / attention_weights tensors it takes `raw` = the output row of the two Linear layers on the query
// ([N*Lq, M*L*P*3]: M*L*P*2 sampling offsets, then M*L*P attention logits) and the reference points, and computes per
// (n, q, head) the softmax over the L*P logits (:99) and  loc = ref + offset / (W_l, H_l)  (:102-105; boxes, ref_dim 4: ref_xy +
// offset / P * ref_wh * 0.5, :106-108) on the fly.  LP = L * P must be <= 32.
// PT as above (4 points fully unrolled: the 16 loads of a level are in flight together; 0 = run-time loop).
template <typename T, int PT>
__global__ void __launch_bounds__(256)
msda_fused_kernel(const T* __restrict__ value, const float* __restrict__ raw, const float* __restrict__ ref, T* __restrict__ out,
                  const __grid_constant__ MsdaLevels lv, long total, int S, int M, int D, int L, int Q, int P, int ref_dim) {
    constexpr int V = Pack16<T>::N;
    const int groups = D / V;
    const int np = PT ? PT : P;
    const int LP = L * np, MLP = M * LP;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % groups);
        const long unit = idx / groups;
        const int m = (int)(unit % M);
        const long nq = unit / M;                                  // n * Q + q
        const int n = (int)(nq / Q);
        const float* row = raw + nq * (size_t)(3 * MLP);
        const float2* offp = reinterpret_cast<const float2*>(row + (size_t)m * LP * 2);
        const float* logp = row + 2 * (size_t)MLP + (size_t)m * LP;
        const float* rp = ref + nq * (size_t)(L * ref_dim);
        // softmax over the L*P logits of this head (F.softmax(..., -1), fp32): maximum, then ONE exponential per logit (the
        // normalisation is applied to the accumulated sum at the end)
        float mx = -INFINITY;
        for (int j = 0; j < LP; ++j) mx = fmaxf(mx, __ldg(logp + j));
        float den = 0.f;
        const size_t head_stride = (size_t)M * D;
        const T* vbase = value + (size_t)n * S * head_stride + (size_t)m * D + (size_t)cg * V;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        for (int l = 0; l < L; ++l) {
            const int H = lv.h[l], W = lv.w[l];
            const T* vl = vbase + (size_t)lv.start[l] * head_stride;
            const float rx = __ldg(rp + l * ref_dim), ry = __ldg(rp + l * ref_dim + 1);
#pragma unroll
            for (int p = 0; p < (PT ? PT : 1); ++p) {
                for (int pr = 0; pr < (PT ? 1 : np); ++pr) {      // run-time point loop only in the generic instantiation
                    const int j = l * np + (PT ? p : pr);
                    const float2 off = __ldg(offp + j);
                    const float wt = __expf(__ldg(logp + j) - mx);
                    den += wt;
                    const float loc_x = ref_dim == 2 ? rx + off.x / (float)W : rx + off.x / (float)np * __ldg(rp + l * ref_dim + 2) * 0.5f;
                    const float loc_y = ref_dim == 2 ? ry + off.y / (float)H : ry + off.y / (float)np * __ldg(rp + l * ref_dim + 3) * 0.5f;
                    const float h_im = loc_y * H - 0.5f, w_im = loc_x * W - 0.5f;
                    const bool inside = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
                    const float hf = floorf(h_im), wf = floorf(w_im);
                    const int h_low = (int)hf, w_low = (int)wf;
                    const int h_high = h_low + 1, w_high = w_low + 1;
                    const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
                    const bool hl = inside && h_low >= 0, hhi = inside && h_high <= H - 1;
                    const bool wl = w_low >= 0, whi = w_high <= W - 1;
                    const float w1 = (hl && wl) ? hh * hw * wt : 0.f, w2 = (hl && whi) ? hh * lw * wt : 0.f;
                    const float w3 = (hhi && wl) ? lh * hw * wt : 0.f, w4 = (hhi && whi) ? lh * lw * wt : 0.f;
                    const int y0 = min(max(h_low, 0), H - 1), y1 = min(max(h_high, 0), H - 1);
                    const int x0 = min(max(w_low, 0), W - 1), x1 = min(max(w_high, 0), W - 1);
                    float t1[V], t2[V], t3[V], t4[V];
                    Pack16<T>::load(vl + ((size_t)y0 * W + x0) * head_stride, t1);
                    Pack16<T>::load(vl + ((size_t)y0 * W + x1) * head_stride, t2);
                    Pack16<T>::load(vl + ((size_t)y1 * W + x0) * head_stride, t3);
                    Pack16<T>::load(vl + ((size_t)y1 * W + x1) * head_stride, t4);
#pragma unroll
                    for (int i = 0; i < V; ++i) acc[i] = fmaf(w1, t1[i], fmaf(w2, t2[i], fmaf(w3, t3[i], fmaf(w4, t4[i], acc[i]))));
                }
            }
        }
        const float inv_den = 1.f / den;
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] *= inv_den;
        Pack16<T>::store(out + unit * D + (size_t)cg * V, acc);
    }
}

// The fused kernel with the per-(query, head) arithmetic SHARED by the unit's lanes.  Measured on msda_fused_kernel (ncu, step1.yaml
// geometry, profiles/r02_final2/ncu_full_heads.csv): issue slots 70 % busy, L1 29 %, L2 19 % — the gather was bound by instructions,
// and half of them were the softmax / sampling-location / bilinear-weight arithmetic that each of the 8 (bf16) or 16 (fp32) lanes
// of a unit repeated for itself.  Here lane c of a unit prepares points c, c + G, ... (weights incl. the un-normalised softmax weight
// and the four element offsets of the taps) into shared memory, the maximum / denominator of the softmax are reduced with shuffles
// inside the unit, and the accumulation loop reads a point as two 16-byte broadcasts.  G = lanes per unit (a power of two <= 32).
// FUSED = false: the plain sampling core (svb_ms_deform_attn_forward) — `raw` = sampling_locations [N, Lq, M, L, P, 2], `ref` =
// attention_weights [N, Lq, M, L, P] (already normalised), ref_dim unused.
template <typename T, bool FUSED>
__global__ void __launch_bounds__(256)
msda_fused_shared_kernel(const T* __restrict__ value, const float* __restrict__ raw, const float* __restrict__ ref, T* __restrict__ out,
                         const __grid_constant__ MsdaLevels lv, long total, int S, int M, int D, int L, int Q, int np, int ref_dim) {
    constexpr int V = Pack16<T>::N;
    extern __shared__ uint4 msda_pts[];                            // [units per block][LP][2]: (w1..w4), (o1..o4)
    const int G = D / V;
    const int LP = L * np, MLP = M * LP;
    const int lane = threadIdx.x & 31;
    const int cg = lane & (G - 1);
    uint4* mine = msda_pts + (size_t)(threadIdx.x / G) * LP * 2;
    const size_t head_stride = (size_t)M * D;
    const long first = blockIdx.x * (long)blockDim.x + (threadIdx.x & ~31);          // the warp's first index: uniform loop bounds
    for (long wbase = first; wbase < total; wbase += (long)gridDim.x * blockDim.x) {
        const long idx = wbase + lane;
        const bool valid = idx < total;                            // whole units are valid or not (total is a multiple of G)
        const long unit = valid ? idx / G : 0;
        const int m = (int)(unit % M);
        const long nq = unit / M;                                  // n * Q + q
        const int n = (int)(nq / Q);
        const float* row = FUSED ? raw + nq * (size_t)(3 * MLP) : raw + (size_t)unit * LP * 2;
        const float2* offp = reinterpret_cast<const float2*>(FUSED ? row + (size_t)m * LP * 2 : row);
        const float* logp = FUSED ? row + 2 * (size_t)MLP + (size_t)m * LP : ref + (size_t)unit * LP;
        const float* rp = ref + nq * (size_t)(L * ref_dim);
        // ---- this lane's points: j = cg, cg + G, ... ----
        float mx = -INFINITY;
        if (FUSED) {
            for (int j = cg; j < LP; j += G) mx = fmaxf(mx, __ldg(logp + j));
            for (int o = 1; o < G; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        float den = 0.f;
        for (int j = cg; j < LP; j += G) {
            const int l = j / np;
            const int H = lv.h[l], W = lv.w[l];
            const float2 off = __ldg(offp + j);
            const float wt = FUSED ? __expf(__ldg(logp + j) - mx) : __ldg(logp + j);
            den += wt;
            float loc_x = off.x, loc_y = off.y;                   // the sampling location itself when it is materialised
            if (FUSED) {
                const float rx = __ldg(rp + l * ref_dim), ry = __ldg(rp + l * ref_dim + 1);
                loc_x = ref_dim == 2 ? rx + off.x / (float)W : rx + off.x / (float)np * __ldg(rp + l * ref_dim + 2) * 0.5f;
                loc_y = ref_dim == 2 ? ry + off.y / (float)H : ry + off.y / (float)np * __ldg(rp + l * ref_dim + 3) * 0.5f;
            }
            const float h_im = loc_y * H - 0.5f, w_im = loc_x * W - 0.5f;
            const bool inside = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int h_low = (int)hf, w_low = (int)wf;
            const int h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
            const bool hl = inside && h_low >= 0, hhi = inside && h_high <= H - 1;
            const bool wl = w_low >= 0, whi = w_high <= W - 1;
            float4 w4;
            w4.x = (hl && wl) ? hh * hw * wt : 0.f; w4.y = (hl && whi) ? hh * lw * wt : 0.f;
            w4.z = (hhi && wl) ? lh * hw * wt : 0.f; w4.w = (hhi && whi) ? lh * lw * wt : 0.f;
            const int y0 = min(max(h_low, 0), H - 1), y1 = min(max(h_high, 0), H - 1);
            const int x0 = min(max(w_low, 0), W - 1), x1 = min(max(w_high, 0), W - 1);
            const uint32_t hs = (uint32_t)head_stride, st = (uint32_t)lv.start[l];
            uint4 o4;                                              // element offsets of the four taps inside this sample's value map
            o4.x = (st + (uint32_t)(y0 * W + x0)) * hs; o4.y = (st + (uint32_t)(y0 * W + x1)) * hs;
            o4.z = (st + (uint32_t)(y1 * W + x0)) * hs; o4.w = (st + (uint32_t)(y1 * W + x1)) * hs;
            if (valid) {
                mine[2 * j] = make_uint4(__float_as_uint(w4.x), __float_as_uint(w4.y), __float_as_uint(w4.z), __float_as_uint(w4.w));
                mine[2 * j + 1] = o4;
            }
        }
        if (FUSED) for (int o = 1; o < G; o <<= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        __syncwarp();
        // ---- the gather: every lane its 16 bytes of channels ----
        const T* vb = value + (size_t)n * S * head_stride + (size_t)m * D + (size_t)cg * V;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        if (valid) {
#pragma unroll 4
            for (int j = 0; j < LP; ++j) {
                const uint4 wq = mine[2 * j], oq = mine[2 * j + 1];
                float t1[V], t2[V], t3[V], t4[V];
                Pack16<T>::load(vb + oq.x, t1);
                Pack16<T>::load(vb + oq.y, t2);
                Pack16<T>::load(vb + oq.z, t3);
                Pack16<T>::load(vb + oq.w, t4);
                const float w1 = __uint_as_float(wq.x), w2 = __uint_as_float(wq.y), w3 = __uint_as_float(wq.z), w4 = __uint_as_float(wq.w);
#pragma unroll
                for (int i = 0; i < V; ++i) acc[i] = fmaf(w1, t1[i], fmaf(w2, t2[i], fmaf(w3, t3[i], fmaf(w4, t4[i], acc[i]))));
            }
            if (FUSED) {
                const float inv_den = 1.f / den;
#pragma unroll
                for (int i = 0; i < V; ++i) acc[i] *= inv_den;
            }
            Pack16<T>::store(out + unit * D + (size_t)cg * V, acc);
        }
        __syncwarp();                                              // the unit's points are rewritten in the next round
    }
}

int msda_levels(MsdaLevels& lv, const int32_t* spatial_shapes, const int32_t* level_start_index, int num_levels, int spatial_size,
                const char* who) {
    SVB_REQUIRE(num_levels >= 1 && num_levels <= MSDA_MAX_LEVELS, "%s: %d levels (1..%d supported)", who, num_levels, MSDA_MAX_LEVELS);
    long covered = 0;
    for (int l = 0; l < num_levels; ++l) {
        lv.h[l] = spatial_shapes[2 * l];
        lv.w[l] = spatial_shapes[2 * l + 1];
        lv.start[l] = level_start_index[l];
        SVB_REQUIRE(lv.h[l] > 0 && lv.w[l] > 0 && lv.start[l] >= 0 && (long)lv.start[l] + (long)lv.h[l] * lv.w[l] <= spatial_size,
                    "%s: level %d (%d x %d at %d) does not fit in %d positions", who, l, lv.h[l], lv.w[l], lv.start[l], spatial_size);
        covered += (long)lv.h[l] * lv.w[l];
    }
    SVB_REQUIRE(covered == spatial_size, "%s: the levels cover %ld positions, value has %d", who, covered, spatial_size);
    return 0;
}

}  // namespace
}  // namespace svb

using namespace svb;

extern "C" int svb_ms_deform_attn_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                                          const float* sampling_locations, const float* attention_weights, void* out, int dtype, int batch,
                                          int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_points,
                                          svb_stream_t stream) {
    SVB_REQUIRE(value && spatial_shapes && level_start_index && sampling_locations && attention_weights && out,
                "svb_ms_deform_attn_forward: null argument");
    SVB_REQUIRE(dtype == SVB_DTYPE_F32 || dtype == SVB_DTYPE_BF16, "svb_ms_deform_attn_forward: bad dtype %d", dtype);
    const int vec = dtype == SVB_DTYPE_BF16 ? 8 : 4;
    SVB_REQUIRE(channels > 0 && channels % vec == 0, "svb_ms_deform_attn_forward: channels per head (%d) must be a multiple of %d", channels, vec);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(value) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "svb_ms_deform_attn_forward: value / out must be 16-byte aligned");
    SVB_REQUIRE(batch >= 0 && num_query >= 0 && num_heads > 0 && num_points > 0 && spatial_size >= 0, "svb_ms_deform_attn_forward: bad sizes");
    MsdaLevels lv;
    {
        int rc = msda_levels(lv, spatial_shapes, level_start_index, num_levels, spatial_size, "svb_ms_deform_attn_forward");
        if (rc) return rc;
    }
    const long total = (long)batch * num_query * num_heads * (channels / vec);
    if (total == 0) return 0;
    const long blocks_needed = (total + 255) / 256;
    const int blocks = (int)(blocks_needed < 148L * 32 ? blocks_needed : 148L * 32);
    const double bytes = (double)batch * num_query * num_heads *
                         ((double)num_levels * num_points * (12.0 + 4.0 * channels * (vec == 8 ? 2 : 4)) + (double)channels * (vec == 8 ? 2 : 4));
    ProfScope prof(PC_OTHER, 0, bytes, (cudaStream_t)stream);
#define SVB_MSDA(TT, PT)                                                                                                         \
    msda_forward_kernel<TT, PT><<<blocks, 256, 0, (cudaStream_t)stream>>>((const TT*)value, sampling_locations, attention_weights, (TT*)out, \
                                                                           lv, total, spatial_size, num_heads, channels, num_levels,        \
                                                                           num_query, num_points)
    const int G = channels / vec, LP = num_levels * num_points;
    static const int shared_on = [] { const char* e = getenv("SVB_MSDA_SHARED"); return e ? atoi(e) : 1; }();
    const bool can_share = shared_on && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && (size_t)(256 / G) * LP * 2 * sizeof(uint4) <= 48 * 1024 &&
                           (double)spatial_size * num_heads * channels < 4.0e9 && (reinterpret_cast<uintptr_t>(sampling_locations) & 7) == 0;
    if (can_share) {
        const size_t smem = (size_t)(256 / G) * LP * 2 * sizeof(uint4);
        if (dtype == SVB_DTYPE_BF16)
            msda_fused_shared_kernel<bf16, false><<<blocks, 256, smem, (cudaStream_t)stream>>>((const bf16*)value, sampling_locations, attention_weights,
                                                                                           (bf16*)out, lv, total, spatial_size, num_heads, channels,
                                                                                           num_levels, num_query, num_points, 2);
        else
            msda_fused_shared_kernel<float, false><<<blocks, 256, smem, (cudaStream_t)stream>>>((const float*)value, sampling_locations,
                                                                                            attention_weights, (float*)out, lv, total, spatial_size,
                                                                                            num_heads, channels, num_levels, num_query, num_points, 2);
    } else if (dtype == SVB_DTYPE_BF16) {
        if (num_points == 4) SVB_MSDA(bf16, 4);
        else SVB_MSDA(bf16, 0);
    } else {
        if (num_points == 4) SVB_MSDA(float, 4);
        else SVB_MSDA(float, 0);
    }
#undef SVB_MSDA
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_ms_deform_attn_fused_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                                                const float* reference_points, int ref_dim, const float* offsets_and_logits, void* out,
                                                int dtype, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                                int num_query, int num_points, svb_stream_t stream) {
    SVB_REQUIRE(value && spatial_shapes && level_start_index && reference_points && offsets_and_logits && out,
                "svb_ms_deform_attn_fused_forward: null argument");
    SVB_REQUIRE(dtype == SVB_DTYPE_F32 || dtype == SVB_DTYPE_BF16, "svb_ms_deform_attn_fused_forward: bad dtype %d", dtype);
    SVB_REQUIRE(ref_dim == 2 || ref_dim == 4, "svb_ms_deform_attn_fused_forward: last dim of reference_points must be 2 or 4, got %d", ref_dim);
    const int vec = dtype == SVB_DTYPE_BF16 ? 8 : 4;
    SVB_REQUIRE(channels > 0 && channels % vec == 0, "svb_ms_deform_attn_fused_forward: channels per head (%d) must be a multiple of %d", channels,
                vec);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(value) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(offsets_and_logits) & 7) == 0, "svb_ms_deform_attn_fused_forward: misaligned pointer");
    SVB_REQUIRE(batch >= 0 && num_query >= 0 && num_heads > 0 && num_points > 0 && spatial_size >= 0 && ((num_heads * num_levels * num_points) % 2) == 0,
                "svb_ms_deform_attn_fused_forward: bad sizes");
    MsdaLevels lv;
    {
        int rc = msda_levels(lv, spatial_shapes, level_start_index, num_levels, spatial_size, "svb_ms_deform_attn_fused_forward");
        if (rc) return rc;
    }
    const long total = (long)batch * num_query * num_heads * (channels / vec);
    if (total == 0) return 0;
    const long blocks_needed = (total + 255) / 256;
    const int blocks = (int)(blocks_needed < 148L * 32 ? blocks_needed : 148L * 32);
    const double bytes = (double)batch * num_query * num_heads *
                         ((double)num_levels * num_points * (12.0 + 4.0 * channels * (vec == 8 ? 2 : 4)) + (double)channels * (vec == 8 ? 2 : 4));
    ProfScope prof(PC_OTHER, 0, bytes, (cudaStream_t)stream);
#define SVB_FUSED(TT, PT)                                                                                                   \
    msda_fused_kernel<TT, PT><<<blocks, 256, 0, (cudaStream_t)stream>>>((const TT*)value, offsets_and_logits, reference_points, (TT*)out, lv, \
                                                                         total, spatial_size, num_heads, channels, num_levels, num_query,  \
                                                                         num_points, ref_dim)
    // lanes of a unit share the point arithmetic when they sit inside one warp (channel groups per head a power of two <= 32) and a
    // sample's value map is addressable with 32-bit element offsets; SVB_MSDA_SHARED=0 keeps the one-lane-does-all kernel for A/B
    const int G = channels / vec, LP = num_levels * num_points;
    static const int shared_on = [] { const char* e = getenv("SVB_MSDA_SHARED"); return e ? atoi(e) : 1; }();
    const bool can_share = shared_on && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && (size_t)(256 / G) * LP * 2 * sizeof(uint4) <= 48 * 1024 &&
                           (double)spatial_size * num_heads * channels < 4.0e9;
    if (can_share) {
        const size_t smem = (size_t)(256 / G) * LP * 2 * sizeof(uint4);
        if (dtype == SVB_DTYPE_BF16)
            msda_fused_shared_kernel<bf16, true><<<blocks, 256, smem, (cudaStream_t)stream>>>((const bf16*)value, offsets_and_logits, reference_points, (bf16*)out,
                                                                                        lv, total, spatial_size, num_heads, channels, num_levels,
                                                                                        num_query, num_points, ref_dim);
        else
            msda_fused_shared_kernel<float, true><<<blocks, 256, smem, (cudaStream_t)stream>>>((const float*)value, offsets_and_logits, reference_points,
                                                                                         (float*)out, lv, total, spatial_size, num_heads, channels,
                                                                                         num_levels, num_query, num_points, ref_dim);
    } else if (dtype == SVB_DTYPE_BF16) { if (num_points == 4) SVB_FUSED(bf16, 4); else SVB_FUSED(bf16, 0); }
    else { if (num_points == 4) SVB_FUSED(float, 4); else SVB_FUSED(float, 0); }
#undef SVB_FUSED
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
