// HBM-bound stages of the encoder: patch im2col, LayerNorm, GroupNorm(1,C) apply (+GELU, + NHWC->NCHW un-shuffle),
// the neck's cast / 2x2 space-to-depth gather, and the one-time weight packing kernels.
// All are sized for coalesced 16-byte accesses; their roofline is HBM bandwidth.
#include <cuda_fp16.h>
#include "common.cuh"

namespace svb {
namespace {

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
    static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void store(bf16* p, float a, float b, float c, float d) {
        uint2 u;
        u.x = pack_bf16x2(a, b);
        u.y = pack_bf16x2(c, d);
        *reinterpret_cast<uint2*>(p) = u;
    }
    static __device__ __forceinline__ float4 load(const bf16* p) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
        return make_float4(a.x, a.y, b.x, b.y);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// im2col for the patch embedding (PatchEmbed, image_encoder.py:379-410): x NCHW fp32 -> rows = patches (b,py,px),
// cols = (c, ky, kx), matching patch_embed.proj.weight.reshape(D, C*p*p).  One thread = 4 consecutive kx.
// 4 consecutive input elements of type TI (fp32 / fp16 / bf16: the reference's pipeline feeds fp16 images after
// cast_batch_to_half, pipeline/XDecoderPipeline.py:93-95) as fp32
template <typename TI> struct Load4;
template <> struct Load4<float> {
    static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
};
template <> struct Load4<__half> {
    static __device__ __forceinline__ float4 load(const __half* p) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
};
template <> struct Load4<bf16> {
    static __device__ __forceinline__ float4 load(const bf16* p) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
};

template <typename T, typename TI>
__global__ void im2col_kernel(const TI* __restrict__ x, T* __restrict__ out, int B, int C, int img_h, int img_w, int patch) {
    const int gh = img_h / patch, gw = img_w / patch;
    const int K = C * patch * patch;
    const size_t plane = (size_t)img_h * img_w;
    const size_t total4 = (size_t)B * C * plane / 4;
    const unsigned plane4 = (unsigned)(plane / 4), w4 = (unsigned)img_w / 4;      // (64-bit divisions per element made this kernel issue-bound)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;                      // flat NCHW element index (input-major -> coalesced reads)
        const unsigned pl = (unsigned)(i / plane4), in4 = (unsigned)(i - (size_t)pl * plane4);   // (image, channel) plane; vector inside it
        const int Y = (int)(in4 / w4), X = (int)(in4 - (unsigned)Y * w4) * 4;
        const int c = (int)(pl % (unsigned)C), b = (int)(pl / (unsigned)C);
        const float4 v = Load4<TI>::load(x + e);
        const int px = X / patch, kx = X % patch, py = Y / patch, ky = Y % patch;
        const size_t row = ((size_t)b * gh + py) * gw + px;
        Vec4<T>::store(out + row * K + (size_t)c * patch * patch + ky * patch + kx, v.x, v.y, v.z, v.w);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Input staging fused with the patch im2col (row N2 of the scope table): what the caller of the encoder does per batch —
// `(x - pixel_mean) / pixel_std` on each uint8 (C,h,w) image, then ImageList.from_tensors(images, 1024): bottom/right zero
// padding to the 1024 x 1024 canvas (xdecoder_model.py:481-484) — written straight into the patch-embedding GEMM's A operand
// (rows = patches, cols = (c, ky, kx)).  One thread = 4 consecutive kx of one patch row; the normalised fp32 canvas never exists.
struct U8Batch {
    const uint8_t* img[16];
    int h[16], w[16];
    float mean[4], std[4];
};
template <typename T, int VEC>      // VEC consecutive kx per thread: 16 (one 16-pixel patch row, patch % 16 == 0) or 4
__global__ void stage_u8_kernel(U8Batch bt, T* __restrict__ out, int B, int C, int img, int patch) {
    const int g = img / patch;                       // (square canvases only: larger canvases go through the fp32 entry point)
    const int pp = patch * patch;
    const int K = C * pp;
    const size_t totalv = (size_t)B * g * g * K / VEC;
    // (x - mean) / std takes only 256 values per channel: the correctly rounded divisions (bit-exact with the callers' fp32 expression,
    // ~10 instructions each) are done ONCE per block into a table, the pixels then cost one shared-memory lookup
    __shared__ float lut[4 * 256];
    for (int t = threadIdx.x; t < 256 * C && t < 4 * 256; t += blockDim.x) lut[t] = __fdiv_rn((float)(t & 255) - bt.mean[t >> 8], bt.std[t >> 8]);
    __syncthreads();
    // (32-bit index arithmetic: at most 16 images per launch, so every index below fits; the 64-bit divisions of the first version were
    // most of the kernel's 24 instructions per pixel — ncu: issue slots 66 %, ALU pipe 60 %, DRAM 24 %)
    const unsigned vpr = (unsigned)K / VEC, ug = (unsigned)g;        // vectors per output row
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)totalv; i += gridDim.x * blockDim.x) {
        const size_t e = (size_t)i * VEC;            // flat output index (row-major [B*g*g, K]): coalesced stores
        const unsigned row = i / vpr;
        const int k = (int)((i - row * vpr) * VEC);
        const unsigned rg = row / ug;
        const int px = (int)(row - rg * ug), b = (int)(rg / ug), py = (int)(rg - (unsigned)b * ug);
        const int c = k / pp, ky = (k - c * pp) / patch, kx = k - c * pp - ky * patch;
        const int y = py * patch + ky, x = px * patch + kx;
        const int h = bt.h[b], w = bt.w[b];
        float v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = 0.f;
        if (y < h && x < w) {
            const uint8_t* src = bt.img[b] + ((size_t)c * h + y) * w + x;
            const float* tab = lut + 256 * c;
            uint8_t px8[VEC];
            if (x + VEC <= w && (reinterpret_cast<uintptr_t>(src) & (VEC - 1)) == 0) {       // whole vector inside the row and aligned
                if constexpr (VEC == 16) *reinterpret_cast<uint4*>(px8) = __ldg(reinterpret_cast<const uint4*>(src));
                else *reinterpret_cast<uint32_t*>(px8) = __ldg(reinterpret_cast<const uint32_t*>(src));
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) px8[j] = (x + j < w) ? src[j] : (uint8_t)0;
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (x + j < w) v[j] = tab[px8[j]];
        }
        if constexpr (sizeof(T) == 2 && VEC == 16) {
            // 32 bytes per thread as two 16-byte stores (four 8-byte stores made every warp store touch each 32-byte sector
            // four times); streaming: the rows are read next by the patch-embedding GEMM's TMA, not by this SM
#pragma unroll
            for (int j = 0; j < VEC; j += 8) {
                uint4 u;
                u.x = pack_bf16x2(v[j], v[j + 1]); u.y = pack_bf16x2(v[j + 2], v[j + 3]);
                u.z = pack_bf16x2(v[j + 4], v[j + 5]); u.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(out + e + j) = u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; j += 4) Vec4<T>::store(out + e + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (norm1 / norm2, image_encoder.py:166,176; eps 1e-6 from build_sam.py:65).
// fp32 residual stream in, T out.  One warp per row, row cached in registers, two-pass variance.
constexpr int LN_MAXV = 16;   // float4 per lane -> D <= 2048
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_kernel(float* __restrict__ x, const T* __restrict__ add, const float* __restrict__ w, const float* __restrict__ bia,
                 T* __restrict__ out, int rows, int D, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int n4 = D >> 2;
    float4* xr = reinterpret_cast<float4*>(x + (size_t)warp * D);
    float4 v[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            v[i] = xr[idx];
            if (add) {      // fused residual add (x += add) ahead of the norm: the sum goes back to the fp32 stream
                const float4 a = Vec4<T>::load(add + (size_t)warp * D + 4 * idx);
                v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
                xr[idx] = v[i];
            }
            sum += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
    T* orow = out + (size_t)warp * D;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
            const float4 b = __ldg(reinterpret_cast<const float4*>(bia) + idx);
            Vec4<T>::store(orow + 4 * idx, (v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                           (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Neck input staging: the SimpleFPN consumes the raw residual stream (no norm, image_encoder.py:120).  Writes
//   xb  [B*g*g, D]      = cast(x)                        (A operand of down_4.0 / down_8.0 / down_16.0)
//   a32 [B*(g/2)^2, 4D] = 2x2 space-to-depth gather     (A operand of the k=2,s=2 conv down_32.0, image_encoder.py:442)
//        a32[(b,Y,X), (dy*2+dx)*D + c] = x[(b,2Y+dy,2X+dx), c]
template <typename T>
__global__ void cast_s2d_kernel(const float* __restrict__ x, T* __restrict__ xb, T* __restrict__ a32, int B, int gh, int gw, int D) {
    const size_t total4 = (size_t)B * gh * gw * D / 4;
    const int hh = gh / 2, hw = gw / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;
        const int c = (int)(e % D);
        const size_t tok = e / D;
        const int xx = (int)(tok % gw), yy = (int)((tok / gw) % gh), b = (int)(tok / ((size_t)gh * gw));
        const float4 v = *reinterpret_cast<const float4*>(x + e);
        if (xb) Vec4<T>::store(xb + e, v.x, v.y, v.z, v.w);
        const size_t row = ((size_t)b * hh + (yy >> 1)) * hw + (xx >> 1);
        const int sub = (yy & 1) * 2 + (xx & 1);
        Vec4<T>::store(a32 + row * (size_t)(4 * D) + (size_t)sub * D + c, v.x, v.y, v.z, v.w);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm(1, C) apply (image_encoder.py:419-446): statistics (sum, sumsq per sample, fp64) were accumulated by the
// producing GEMM's epilogue.  NHWC rows in, NHWC rows out (T) for the next GEMM.
// GELU: exact erf for fp32 outputs (validation mode); the packed-fp32x2 A&S 7.1.25 form (|erf error| <= 2.5e-5, below the
// bf16 rounding of the store) for bf16 outputs, which keeps the kernel HBM-bound instead of issue-bound.
template <typename T> __device__ __forceinline__ void gelu4(float& a, float& b, float& c, float& d);
template <> __device__ __forceinline__ void gelu4<float>(float& a, float& b, float& c, float& d) {
    a = gelu_erf(a); b = gelu_erf(b); c = gelu_erf(c); d = gelu_erf(d);
}
template <> __device__ __forceinline__ void gelu4<bf16>(float& a, float& b, float& c, float& d) {
    gelu_pair(a, b);
    gelu_pair(c, d);
}
__device__ __forceinline__ float2 gn_mean_rstd(const double* __restrict__ stats, int sample, double n, float eps) {
    const double mu = stats[2 * sample] / n;
    const double var = fmax(stats[2 * sample + 1] / n - mu * mu, 0.0);
    return make_float2((float)mu, (float)(1.0 / sqrt(var + (double)eps)));
}

template <typename T>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, T* __restrict__ out, long rows, int C, long rows_per_sample, float eps, int gelu,
                int nsamples) {
    extern __shared__ float2 ms_s[];                 // (mean, rstd) per sample: the fp64 math runs once per block
    for (int sidx = threadIdx.x; sidx < nsamples; sidx += blockDim.x)
        ms_s[sidx] = gn_mean_rstd(stats, sidx, (double)rows_per_sample * C, eps);
    __syncthreads();
    const unsigned cpr = (unsigned)C / 4;            // float4 per row
    const size_t total4 = (size_t)rows * cpr;
    const size_t per_sample4 = (size_t)rows_per_sample * cpr;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const int sample = (int)(i / per_sample4);
        const unsigned c4 = (unsigned)(i % cpr);
        const float2 ms = ms_s[sample];
        const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float a = (v.x - ms.x) * ms.y * g4.x + b4.x, b = (v.y - ms.x) * ms.y * g4.y + b4.y;
        float cc = (v.z - ms.x) * ms.y * g4.z + b4.z, d = (v.w - ms.x) * ms.y * g4.w + b4.w;
        if (gelu) gelu4<T>(a, b, cc, d);
        Vec4<T>::store(out + i * 4, a, b, cc, d);
    }
}

// Final stage of every branch: GroupNorm(1,C) + GELU, then NHWC (with `levels` folded 2x2 sub-pixel indexes in the row
// index) -> NCHW.  Input rows are ordered (b, y, x, s_1, ..., s_levels) with s_l = dy_l*2+dx_l; output pixel
// Y = y*2^L + sum dy_l 2^(L-l), X likewise.  One block = TP output pixels of one output row x 64 channels, transposed
// through shared memory: 16-byte reads along C (256 B per pixel), 16-byte writes along X (whole 128-byte lines for bf16 at
// TP = 64).
template <typename T, int TP>
__global__ void __launch_bounds__(256)
gn_apply_nchw_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ out, int B, int gh, int g, int levels, int C, float eps, int gelu) {
    constexpr int TC = 64;
    __shared__ float tile[TC][TP + 1];
    __shared__ float2 ms_s;
    const int Wout = g << levels, Hout = gh << levels;      // g = token-grid width, gh = its height
    const int xt = Wout / TP;
    const int X0 = (blockIdx.x % xt) * TP;
    const int Y = (blockIdx.x / xt) % Hout;
    const int b = blockIdx.x / (xt * Hout);
    const int c0 = blockIdx.y * TC;
    if (threadIdx.x == 0) ms_s = gn_mean_rstd(stats, b, (double)((long)gh * g << (2 * levels)) * C, eps);
    __syncthreads();
    const float mean = ms_s.x, rstd = ms_s.y;
    {   // ---- phase 1: 16 threads per pixel (4 channels each), 16 pixels per pass ----
        const int cq = threadIdx.x & 15, pl = threadIdx.x >> 4;
        const int c = c0 + 4 * cq;
        const bool c_ok = c < C;                      // C is a multiple of 4
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
        if (c_ok) { g4 = __ldg(reinterpret_cast<const float4*>(gamma + c)); b4 = __ldg(reinterpret_cast<const float4*>(beta + c)); }
        float4 v[TP / 16];
#pragma unroll
        for (int k = 0; k < TP / 16; ++k) {
            const int X = X0 + pl + 16 * k;
            long row = ((long)b * gh + (Y >> levels)) * g + (X >> levels);
            for (int l = 1; l <= levels; ++l) {
                const int sh = levels - l;
                row = row * 4 + (((Y >> sh) & 1) * 2 + ((X >> sh) & 1));
            }
            v[k] = c_ok ? __ldcs(reinterpret_cast<const float4*>(x + row * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < TP / 16; ++k) {
            float a0 = (v[k].x - mean) * rstd * g4.x + b4.x, a1 = (v[k].y - mean) * rstd * g4.y + b4.y;
            float a2 = (v[k].z - mean) * rstd * g4.z + b4.z, a3 = (v[k].w - mean) * rstd * g4.w + b4.w;
            if (gelu) gelu4<T>(a0, a1, a2, a3);
            const int px = pl + 16 * k;
            tile[4 * cq + 0][px] = a0; tile[4 * cq + 1][px] = a1; tile[4 * cq + 2][px] = a2; tile[4 * cq + 3][px] = a3;
        }
    }
    __syncthreads();
    // ---- phase 2: 16 bytes of consecutive X per thread ----
    constexpr int EPT = 16 / (int)sizeof(T);          // elements per thread: 8 (bf16) or 4 (fp32)
    constexpr int TPR = TP / EPT;                     // threads per channel row
    for (int idx = threadIdx.x; idx < TC * TPR; idx += 256) {
        const int cj = idx / TPR, pg = idx % TPR;
        const int c = c0 + cj;
        if (c >= C) continue;
        T* dst = out + (((size_t)b * C + c) * Hout + Y) * Wout + X0 + pg * EPT;
        const float* src = &tile[cj][pg * EPT];
        if constexpr (sizeof(T) == 2) {
            uint4 u;
            u.x = pack_bf16x2(src[0], src[1]); u.y = pack_bf16x2(src[2], src[3]);
            u.z = pack_bf16x2(src[4], src[5]); u.w = pack_bf16x2(src[6], src[7]);
            __stcs(reinterpret_cast<uint4*>(dst), u);
        } else {
            __stcs(reinterpret_cast<float4*>(dst), make_float4(src[0], src[1], src[2], src[3]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing (run once per load_state_dict)
template <typename T>
__global__ void pack_cast_kernel(const float* __restrict__ s, T* __restrict__ d, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) d[i] = from_float<T>(s[i]);
}
// ConvTranspose2d weight (Cin, Cout, 2, 2) -> GEMM weight [N = (dy*2+dx)*Cout + co, K = ci]
template <typename T>
__global__ void pack_convT_kernel(const float* __restrict__ w, T* __restrict__ d, int Cin, int Cout) {
    const long n = (long)4 * Cout * Cin;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int nrow = (int)(i / Cin);
        const int co = nrow % Cout, sub = nrow / Cout;
        d[i] = from_float<T>(w[((long)ci * Cout + co) * 4 + sub]);
    }
}
// Conv2d k=2,s=2 weight (Cout, Cin, 2, 2) -> GEMM weight [N = co, K = (dy*2+dx)*Cin + ci]
template <typename T>
__global__ void pack_conv2x2_kernel(const float* __restrict__ w, T* __restrict__ d, int Cin, int Cout) {
    const long n = (long)4 * Cout * Cin;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % (4 * Cin));
        const int co = (int)(i / (4 * Cin));
        const int ci = k % Cin, sub = k / Cin;
        d[i] = from_float<T>(w[((long)co * Cin + ci) * 4 + sub]);
    }
}
__global__ void pack_bias4_kernel(const float* __restrict__ b, float* __restrict__ d, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * C) d[i] = b[i % C];
}

// LayerNorm folded into the following Linear (bf16 path):  LN(x) W^T + b = rstd (x Wg^T - mu c) + bf  with
//   Wg[n,k] = bf16(gamma[k] W[n,k]),  c[n] = sum_k Wg[n,k] (of the ROUNDED values: the mean term then cancels exactly),
//   bf[n] = b[n] + sum_k beta[k] W[n,k].   One warp per output row n.
__global__ void fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                               const float* __restrict__ beta, bf16* __restrict__ Wg, float* __restrict__ c, float* __restrict__ bfold,
                               int N, int K) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float sc = 0.f, sb = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = W[(size_t)n * K + k];
        const bf16 wg = __float2bfloat16_rn(w * gamma[k]);
        Wg[(size_t)n * K + k] = wg;
        sc += __bfloat162float(wg);
        sb = fmaf(beta[k], w, sb);
    }
    sc = warp_sum(sc);
    sb = warp_sum(sb);
    if (lane == 0) { c[n] = sc; bfold[n] = (bias ? bias[n] : 0.f) + sb; }
}

// ---------------------------------------------------------------------------------------------------------------
// Fallbacks of the reference for token grids other than the one the tables were trained on (scope row N3):
//   * pos_embed (1,h0,w0,D) -> (1,h1,w1,D), bicubic, align_corners=False, A = -0.75, border taps clamped
//     (ImageEncoderViT.interpolate_pos_encoding, image_encoder.py:124-132: F.interpolate(..., mode='bicubic'));
//   * rel_pos table (L0,hd) -> (L1,hd), linear, align_corners=False (get_rel_pos, image_encoder.py:319-330).
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.f, x3 = 2.f - t, u = 1.f - t;
    w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
    w[2] = ((A + 2.f) * u - (A + 3.f)) * u * u + 1.f;
    w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__global__ void bicubic_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int h0, int w0, int h1, int w1, int D) {
    const float sy = (float)h0 / (float)h1, sx = (float)w0 / (float)w1;
    const long total = (long)h1 * w1 * D;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const int x = (int)((i / D) % w1), y = (int)(i / ((long)D * w1));
        const float fy = sy * (y + 0.5f) - 0.5f, fx = sx * (x + 0.5f) - 0.5f;      // no clamp at 0 for cubic (ATen upsample)
        const int iy = (int)floorf(fy), ix = (int)floorf(fx);
        float wy[4], wx[4];
        cubic_coeffs(fy - iy, wy);
        cubic_coeffs(fx - ix, wx);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = min(max(iy - 1 + a, 0), h0 - 1);
            float r = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int xx = min(max(ix - 1 + b, 0), w0 - 1);
                r += wx[b] * __ldg(src + ((long)yy * w0 + xx) * D + c);
            }
            acc += wy[a] * r;
        }
        dst[i] = acc;
    }
}
__global__ void linear_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int L0, int L1, int hd) {
    const float sc = (float)L0 / (float)L1;
    const int total = L1 * hd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % hd, j = i / hd;
        const float f = fmaxf(sc * (j + 0.5f) - 0.5f, 0.f);
        const int i0 = min((int)f, L0 - 1), i1 = min(i0 + 1, L0 - 1);
        const float lam = f - i0;
        dst[i] = (1.f - lam) * __ldg(src + (long)i0 * hd + c) + lam * __ldg(src + (long)i1 * hd + c);
    }
}

// out = T(a [+ b]): the cast of an fp32 stream (optionally plus a positional embedding, `with_pos_embed` of
// transformer_encoder_deform.py:112-114) to the GEMM operand type
template <typename T>
__global__ void add_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, T* __restrict__ out, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
        if (b) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(b) + i);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        Vec4<T>::store(out + i * 4, v.x, v.y, v.z, v.w);
    }
}

inline int grid_for(size_t n, int block) {
    size_t gsz = (n + block - 1) / block;
    const size_t cap = 148 * 16;
    return (int)(gsz < cap ? (gsz ? gsz : 1) : cap);
}

}  // namespace

// x_dtype: 0 fp32, 1 bf16, 2 fp16 (SVB_DTYPE_*)
int im2col_patch(const void* x, int x_dtype, void* out, bool out_bf16, int B, int C, int img_h, int img_w, int patch, cudaStream_t s) {
    SVB_REQUIRE(img_h % patch == 0 && img_w % patch == 0 && patch % 4 == 0, "im2col_patch: image %d x %d / patch %d unsupported", img_h, img_w,
                patch);
    SVB_REQUIRE(x_dtype >= 0 && x_dtype <= 2, "im2col_patch: input dtype %d is not fp32 (0) / bf16 (1) / fp16 (2)", x_dtype);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(x) & (x_dtype == 0 ? 15 : 7)) == 0, "im2col_patch: the input must be 16-byte (fp32) / 8-byte (half) aligned");
    const size_t total4 = (size_t)B * C * img_h * img_w / 4;
    ProfScope prof(PC_OTHER, 0, (double)total4 * 4 * ((x_dtype == 0 ? 4 : 2) + (out_bf16 ? 2 : 4)), s);
    const int grid = grid_for(total4, 256);
#define SVB_IM2COL(TO, TI) im2col_kernel<TO, TI><<<grid, 256, 0, s>>>((const TI*)x, (TO*)out, B, C, img_h, img_w, patch)
    if (out_bf16) { if (x_dtype == 0) SVB_IM2COL(bf16, float); else if (x_dtype == 1) SVB_IM2COL(bf16, bf16); else SVB_IM2COL(bf16, __half); }
    else { if (x_dtype == 0) SVB_IM2COL(float, float); else if (x_dtype == 1) SVB_IM2COL(float, bf16); else SVB_IM2COL(float, __half); }
#undef SVB_IM2COL
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int stage_u8_patch(const uint8_t* const* images, const int* hs, const int* ws, const float* mean, const float* stdv, void* out, bool out_bf16,
                   int B, int C, int img, int patch, cudaStream_t s) {
    SVB_REQUIRE(img % patch == 0 && patch % 4 == 0 && C >= 1 && C <= 4, "stage_u8_patch: img %d / patch %d / chans %d unsupported", img, patch, C);
    const size_t row_elems = (size_t)(img / patch) * (img / patch) * C * patch * patch;
    for (int b0 = 0; b0 < B; b0 += 16) {
        const int nb = B - b0 < 16 ? B - b0 : 16;
        U8Batch bt;
        for (int i = 0; i < 16; ++i) { bt.img[i] = nullptr; bt.h[i] = 0; bt.w[i] = 0; }
        for (int i = 0; i < nb; ++i) {
            SVB_REQUIRE(images[b0 + i] && hs[b0 + i] >= 1 && ws[b0 + i] >= 1 && hs[b0 + i] <= img && ws[b0 + i] <= img,
                        "stage_u8_patch: image %d is %d x %d; sizes 1..%d are implemented (larger canvases are scope row N3)", b0 + i,
                        hs[b0 + i], ws[b0 + i], img);
            bt.img[i] = images[b0 + i]; bt.h[i] = hs[b0 + i]; bt.w[i] = ws[b0 + i];
        }
        for (int c = 0; c < 4; ++c) { bt.mean[c] = c < C ? mean[c] : 0.f; bt.std[c] = c < C ? stdv[c] : 1.f; }
        const int vec = (patch % 16 == 0) ? 16 : 4;
        const size_t totalv = (size_t)nb * row_elems / vec;
        ProfScope prof(PC_OTHER, 0, (double)nb * C * img * img + (double)nb * row_elems * (out_bf16 ? 2 : 4), s);
        bf16* ob = (bf16*)out + (size_t)b0 * row_elems;
        float* of = (float*)out + (size_t)b0 * row_elems;
        if (vec == 16) {
            if (out_bf16) stage_u8_kernel<bf16, 16><<<grid_for(totalv, 256), 256, 0, s>>>(bt, ob, nb, C, img, patch);
            else stage_u8_kernel<float, 16><<<grid_for(totalv, 256), 256, 0, s>>>(bt, of, nb, C, img, patch);
        } else {
            if (out_bf16) stage_u8_kernel<bf16, 4><<<grid_for(totalv, 256), 256, 0, s>>>(bt, ob, nb, C, img, patch);
            else stage_u8_kernel<float, 4><<<grid_for(totalv, 256), 256, 0, s>>>(bt, of, nb, C, img, patch);
        }
        SVB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

// D = 128 * NV exactly (768 / 1024 / 1280): no bounds tests, the row lives in NV float4 registers per lane, small blocks
// (4 rows) at high occupancy so that enough bytes are in flight to cover the HBM latency (HBM-bound: 4 + sizeof(T) B/element).
template <typename T, int NV>
__global__ void __launch_bounds__(128, 8)
layernorm_fixed_kernel(float* __restrict__ x, const T* __restrict__ add, const float* __restrict__ w, const float* __restrict__ bia,
                       T* __restrict__ out, int rows, float eps) {
    constexpr int D = 128 * NV;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4* xr = reinterpret_cast<float4*>(x + (size_t)row * D);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
    if (add) {              // fused residual add (x += add) ahead of the norm: the sum goes back to the fp32 stream
        float4 a[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) a[i] = Vec4<T>::load(add + (size_t)row * D + 4 * (lane + 32 * i));
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x += a[i].x; v[i].y += a[i].y; v[i].z += a[i].z; v[i].w += a[i].w;
            xr[lane + 32 * i] = v[i];
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    T* orow = out + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
        const float4 b = __ldg(reinterpret_cast<const float4*>(bia) + idx);
        Vec4<T>::store(orow + 4 * idx, (v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                       (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    }
}

// Post-norm LayerNorm of the deformable encoder layers (transformer_encoder_deform.py:126-127, 119), one pass for everything
// the next GEMMs read:  y = LayerNorm(x [+ add])  ->  out (fp32 residual stream), out_b = bf16(y) (operand of linear1 / value_proj),
// out_q = bf16(y + pos[row mod pos_rows]) (`with_pos_embed`, :112-114: operand of the sampling-offset / attention-weight Linear).
// x and add are only read (the pre-norm sum is dead after a post-norm layer).  Replaces LayerNorm + one or two cast passes:
// 4 [+ 4] bytes read and 4 + 2 [+ 2] written per element instead of 12 [+ 12] read and 6 [+ 10] written.
template <int NV>
__global__ void __launch_bounds__(128, 8)
layernorm_post_kernel(const float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ w, const float* __restrict__ bia,
                      float* __restrict__ out, bf16* __restrict__ out_b, const float* __restrict__ pos, int pos_rows,
                      bf16* __restrict__ out_q, int rows, float eps) {
    constexpr int D = 128 * NV;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcs(xr + lane + 32 * i);
    if (add) {
        const float4* ar = reinterpret_cast<const float4*>(add + (size_t)row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 a = __ldcs(ar + lane + 32 * i);
            v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    const float4* pr = (pos && out_q) ? reinterpret_cast<const float4*>(pos + (size_t)(row % pos_rows) * D) : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
        const float4 b = __ldg(reinterpret_cast<const float4*>(bia) + idx);
        const float4 y = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                     (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        if (out) reinterpret_cast<float4*>(out + (size_t)row * D)[idx] = y;
        if (out_b) Vec4<bf16>::store(out_b + (size_t)row * D + 4 * idx, y.x, y.y, y.z, y.w);
        if (out_q) {
            float4 q = y;
            if (pr) { const float4 pp = __ldg(pr + idx); q.x += pp.x; q.y += pp.y; q.z += pp.z; q.w += pp.w; }
            Vec4<bf16>::store(out_q + (size_t)row * D + 4 * idx, q.x, q.y, q.z, q.w);
        }
    }
}

template <typename T>
static bool launch_ln_fixed(float* x, const T* add, const float* w, const float* b, T* out, int rows, int D, float eps, cudaStream_t s) {
    const int blocks = (rows + 3) / 4;
    switch (D) {
        case 256: layernorm_fixed_kernel<T, 2><<<blocks, 128, 0, s>>>(x, add, w, b, out, rows, eps); return true;     // deformable encoder widths
        case 512: layernorm_fixed_kernel<T, 4><<<blocks, 128, 0, s>>>(x, add, w, b, out, rows, eps); return true;
        case 768: layernorm_fixed_kernel<T, 6><<<blocks, 128, 0, s>>>(x, add, w, b, out, rows, eps); return true;
        case 1024: layernorm_fixed_kernel<T, 8><<<blocks, 128, 0, s>>>(x, add, w, b, out, rows, eps); return true;
        case 1280: layernorm_fixed_kernel<T, 10><<<blocks, 128, 0, s>>>(x, add, w, b, out, rows, eps); return true;
        default: return false;
    }
}

int layernorm_rows(float* x, const void* add, const float* w, const float* b, void* out, bool out_bf16, int rows, int D, float eps,
                   cudaStream_t s) {
    SVB_REQUIRE(D % 4 == 0 && D <= LN_MAXV * 128, "layernorm_rows: D=%d unsupported (multiple of 4, <= %d)", D, LN_MAXV * 128);
    const int blocks = (rows + 7) / 8;
    const int es = out_bf16 ? 2 : 4;
    ProfScope prof(PC_NORM, 0, (double)rows * D * (4 + es + (add ? 4 + es : 0)), s);
    const bool fixed = out_bf16 ? launch_ln_fixed<bf16>(x, (const bf16*)add, w, b, (bf16*)out, rows, D, eps, s)
                                : launch_ln_fixed<float>(x, (const float*)add, w, b, (float*)out, rows, D, eps, s);
    if (!fixed) {
        if (out_bf16) layernorm_kernel<bf16><<<blocks, 256, 0, s>>>(x, (const bf16*)add, w, b, (bf16*)out, rows, D, eps);
        else layernorm_kernel<float><<<blocks, 256, 0, s>>>(x, (const float*)add, w, b, (float*)out, rows, D, eps);
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int layernorm_post_rows(const float* x, const float* add, const float* w, const float* b, float* out, bf16* out_b, const float* pos,
                        int pos_rows, bf16* out_q, int rows, int D, float eps, cudaStream_t s) {
    SVB_REQUIRE(D == 256 || D == 512 || D == 768 || D == 1024 || D == 1280, "layernorm_post_rows: width %d (256 / 512 / 768 / 1024 / 1280 are built)", D);
    SVB_REQUIRE(out || out_b || out_q, "layernorm_post_rows: no output");
    SVB_REQUIRE(!pos || pos_rows > 0, "layernorm_post_rows: pos needs pos_rows > 0");
    const int blocks = (rows + 3) / 4;
    ProfScope prof(PC_NORM, 0, (double)rows * D * (4 + (add ? 4 : 0) + (out ? 4 : 0) + (out_b ? 2 : 0) + (out_q ? 2 : 0)), s);
    switch (D) {
        case 256: layernorm_post_kernel<2><<<blocks, 128, 0, s>>>(x, add, w, b, out, out_b, pos, pos_rows, out_q, rows, eps); break;
        case 512: layernorm_post_kernel<4><<<blocks, 128, 0, s>>>(x, add, w, b, out, out_b, pos, pos_rows, out_q, rows, eps); break;
        case 768: layernorm_post_kernel<6><<<blocks, 128, 0, s>>>(x, add, w, b, out, out_b, pos, pos_rows, out_q, rows, eps); break;
        case 1024: layernorm_post_kernel<8><<<blocks, 128, 0, s>>>(x, add, w, b, out, out_b, pos, pos_rows, out_q, rows, eps); break;
        default: layernorm_post_kernel<10><<<blocks, 128, 0, s>>>(x, add, w, b, out, out_b, pos, pos_rows, out_q, rows, eps); break;
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int cast_and_space2depth(const float* x, void* xb, void* a32, bool out_bf16, int B, int gh, int gw, int D, cudaStream_t s) {
    SVB_REQUIRE(D % 4 == 0 && gh % 2 == 0 && gw % 2 == 0, "cast_and_space2depth: D=%d grid=%d x %d unsupported", D, gh, gw);
    const size_t total4 = (size_t)B * gh * gw * D / 4;
    ProfScope prof(PC_OTHER, 0, (double)total4 * 4 * (4 + (out_bf16 ? 2 : 4) * (xb ? 2 : 1)), s);
    if (out_bf16) cast_s2d_kernel<bf16><<<grid_for(total4, 256), 256, 0, s>>>(x, (bf16*)xb, (bf16*)a32, B, gh, gw, D);
    else cast_s2d_kernel<float><<<grid_for(total4, 256), 256, 0, s>>>(x, (float*)xb, (float*)a32, B, gh, gw, D);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out, bool out_bf16,
                    long rows, int C, long rows_per_sample, float eps, int gelu, cudaStream_t s) {
    SVB_REQUIRE(C % 4 == 0, "groupnorm_apply: C=%d must be a multiple of 4", C);
    SVB_REQUIRE(rows_per_sample > 0 && rows % rows_per_sample == 0, "groupnorm_apply: rows %ld not a multiple of rows_per_sample %ld", rows,
                rows_per_sample);
    const int nsamples = (int)(rows / rows_per_sample);
    SVB_REQUIRE(nsamples <= 4096, "groupnorm_apply: %d samples per call (max 4096)", nsamples);
    const size_t total4 = (size_t)rows * C / 4;
    const size_t smem = sizeof(float2) * (size_t)nsamples;
    ProfScope prof(PC_NORM, 0, (double)rows * C * (4 + (out_bf16 ? 2 : 4)), s);
    if (out_bf16)
        gn_apply_kernel<bf16><<<grid_for(total4, 256), 256, smem, s>>>(x, stats, gamma, beta, (bf16*)out, rows, C, rows_per_sample, eps, gelu, nsamples);
    else
        gn_apply_kernel<float><<<grid_for(total4, 256), 256, smem, s>>>(x, stats, gamma, beta, (float*)out, rows, C, rows_per_sample, eps, gelu, nsamples);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int groupnorm_apply_nchw(const float* x, const double* stats, const float* gamma, const float* beta, void* out, int out_dtype,
                         int B, int gh, int gw, int levels, int C, float eps, int gelu, cudaStream_t s) {
    const int Wout = gw << levels, Hout = gh << levels;
    SVB_REQUIRE(Wout % 16 == 0 && levels >= 0 && levels <= 2 && C % 4 == 0, "groupnorm_apply_nchw: grid %d x %d levels %d C %d unsupported", gh,
                gw, levels, C);
    const int TP = (Wout % 64 == 0) ? 64 : ((Wout % 32 == 0) ? 32 : 16);
    dim3 grid((unsigned)((size_t)B * Hout * (Wout / TP)), (C + 63) / 64);
    ProfScope prof(PC_NORM, 0, (double)B * Hout * Wout * C * (4 + (out_dtype == 1 ? 2 : 4)), s);
#define SVB_GN_NCHW(TT, TPV) gn_apply_nchw_kernel<TT, TPV><<<grid, 256, 0, s>>>(x, stats, gamma, beta, (TT*)out, B, gh, gw, levels, C, eps, gelu)
    if (out_dtype == 1) {
        if (TP == 64) SVB_GN_NCHW(bf16, 64);
        else if (TP == 32) SVB_GN_NCHW(bf16, 32);
        else SVB_GN_NCHW(bf16, 16);
    } else {
        if (TP == 64) SVB_GN_NCHW(float, 64);
        else if (TP == 32) SVB_GN_NCHW(float, 32);
        else SVB_GN_NCHW(float, 16);
    }
#undef SVB_GN_NCHW
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int pack_cast(const float* src, void* dst, bool dst_bf16, long n, cudaStream_t s) {
    if (dst_bf16) pack_cast_kernel<bf16><<<grid_for(n, 256), 256, 0, s>>>(src, (bf16*)dst, n);
    else pack_cast_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(src, (float*)dst, n);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int pack_convT(const float* w, void* dst, bool dst_bf16, int Cin, int Cout, cudaStream_t s) {
    const long n = (long)4 * Cin * Cout;
    if (dst_bf16) pack_convT_kernel<bf16><<<grid_for(n, 256), 256, 0, s>>>(w, (bf16*)dst, Cin, Cout);
    else pack_convT_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(w, (float*)dst, Cin, Cout);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int pack_conv2x2(const float* w, void* dst, bool dst_bf16, int Cin, int Cout, cudaStream_t s) {
    const long n = (long)4 * Cin * Cout;
    if (dst_bf16) pack_conv2x2_kernel<bf16><<<grid_for(n, 256), 256, 0, s>>>(w, (bf16*)dst, Cin, Cout);
    else pack_conv2x2_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(w, (float*)dst, Cin, Cout);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
namespace {
__global__ void row_means_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int rows, int D) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += a[(size_t)row * D + d] + (b ? b[d] : 0.f);
    s = warp_sum(s);
    if (lane == 0) out[row] = s / (float)D;
}
}  // namespace

int row_means(const float* a, const float* b, float* out, int rows, int D, cudaStream_t s) {
    ProfScope prof(PC_OTHER, 0, (double)rows * D * 4, s);
    row_means_kernel<<<(rows + 7) / 8, 256, 0, s>>>(a, b, out, rows, D);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int resize_pos_embed(const float* src, float* dst, int h0, int w0, int h1, int w1, int D, cudaStream_t s) {
    SVB_REQUIRE(h0 > 0 && w0 > 0 && h1 > 0 && w1 > 0 && D > 0, "resize_pos_embed: bad sizes");
    const long total = (long)h1 * w1 * D;
    bicubic_resize_kernel<<<grid_for((size_t)total, 256), 256, 0, s>>>(src, dst, h0, w0, h1, w1, D);
    count_launch();
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int resize_rel_pos(const float* src, float* dst, int L0, int L1, int hd, cudaStream_t s) {
    SVB_REQUIRE(L0 > 0 && L1 > 0 && hd > 0, "resize_rel_pos: bad sizes");
    linear_resize_kernel<<<(L1 * hd + 255) / 256, 256, 0, s>>>(src, dst, L0, L1, hd);
    count_launch();
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int add_cast(const float* a, const float* b, void* out, bool out_bf16, size_t n, cudaStream_t s) {
    SVB_REQUIRE(n % 4 == 0, "add_cast: element count must be a multiple of 4");
    ProfScope prof(PC_OTHER, 0, (double)n * (4 + (b ? 4 : 0) + (out_bf16 ? 2 : 4)), s);
    if (out_bf16) add_cast_kernel<bf16><<<grid_for(n / 4, 256), 256, 0, s>>>(a, b, (bf16*)out, n / 4);
    else add_cast_kernel<float><<<grid_for(n / 4, 256), 256, 0, s>>>(a, b, (float*)out, n / 4);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int fold_layernorm(const float* W, const float* bias, const float* gamma, const float* beta, bf16* Wg, float* colsum, float* bias_f, int N,
                   int K, cudaStream_t s) {
    fold_ln_kernel<<<(N * 32 + 255) / 256, 256, 0, s>>>(W, bias, gamma, beta, Wg, colsum, bias_f, N, K);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int pack_bias4(const float* b, float* dst, int C, cudaStream_t s) {
    pack_bias4_kernel<<<(4 * C + 255) / 256, 256, 0, s>>>(b, dst, C);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace svb
