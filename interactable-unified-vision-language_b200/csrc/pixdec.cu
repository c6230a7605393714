// Convolutional part of the MSDeformAttn pixel decoder (scope row N1, `MSDeformAttnPixelDecoder.forward`,
// /root/reference/modeling/vision/encoder/transformer_encoder_deform.py:315-359) around the GEMMs and the deformable encoder:
// everything runs on "rows" = [sample][pixel y*W+x][channel] (NHWC), the layout the tcgen05 GEMM reads and writes, so a 1x1
// convolution is one GEMM and the NCHW <-> rows changes happen once at the module boundary.  All kernels are HBM-bound streaming
// kernels: 16-byte accesses along the contiguous dimension, grids capped at 148 x 16 blocks.
#include "../../include/samvit_b200.h"
#include "common.cuh"

namespace svb {
namespace {

inline int grid_cap(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    const size_t cap = 148 * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

__device__ __forceinline__ void store8(float* p, float4 a, float4 b) { reinterpret_cast<float4*>(p)[0] = a; reinterpret_cast<float4*>(p)[1] = b; }
__device__ __forceinline__ void store8(bf16* p, float4 a, float4 b) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void store4(bf16* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

// ---- NCHW -> rows: 32 x 32 tiles through shared memory (coalesced on both sides); casts to the GEMM operand type ----
template <typename TI, typename TO>
__global__ void nchw_to_rows_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int C, int HW, long long dst_sample_stride,
                                    long long dst_pixel_stride, const float* __restrict__ chan_add) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const TI* s = src + (size_t)b * C * HW;
    TO* d = dst + (size_t)b * dst_sample_stride;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? to_float(s[(size_t)c * HW + p]) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + threadIdx.x;
        if (p < HW && c < C) d[(size_t)p * dst_pixel_stride + c] = from_float<TO>(tile[threadIdx.x][j] + (chan_add ? __ldg(chan_add + c) : 0.f));
    }
}

// ---- rows -> NCHW (fp32): the module's outputs (`mask_features`, `multi_scale_features`) ----
__global__ void rows_to_nchw_kernel(const float* __restrict__ src, long long src_sample_stride, float* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const float* s = src + (size_t)b * src_sample_stride;
    float* d = dst + (size_t)b * C * HW;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (p < HW && c < C) ? s[(size_t)p * C + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + threadIdx.x;
        if (c < C && p < HW) d[(size_t)c * HW + p] = tile[threadIdx.x][j];
    }
}

// ---- GroupNorm(groups, C) on rows (torch.nn.GroupNorm: statistics over (C / groups) channels x all pixels of a sample) ----
// pass 1: per (sample, group) sum / sum of squares in fp64 (fp32 partials over at most 64 rows per thread)
__global__ void gn_rows_stats_kernel(const float* __restrict__ x, long long sample_stride, int HW, int C, int groups, int rows_per_block,
                                     double* __restrict__ stats) {
    extern __shared__ double sacc[];                       // [groups][2]
    const int b = blockIdx.y, r0 = blockIdx.x * rows_per_block;
    const int r1 = min(HW, r0 + rows_per_block), cg = C / groups;
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) sacc[i] = 0.0;
    __syncthreads();
    const float* xb = x + (size_t)b * sample_stride;
    for (int c4 = threadIdx.x * 4; c4 < C; c4 += blockDim.x * 4) {      // cg is a multiple of 2; 4 channels may span two groups
        float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = r0; r < r1; ++r) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)r * C + c4));
            s[0] += v.x; q[0] += v.x * v.x; s[1] += v.y; q[1] += v.y * v.y;
            s[2] += v.z; q[2] += v.z * v.z; s[3] += v.w; q[3] += v.w * v.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = (c4 + j) / cg;
            atomicAdd(&sacc[2 * g], (double)s[j]);
            atomicAdd(&sacc[2 * g + 1], (double)q[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) atomicAdd(&stats[(size_t)b * 2 * groups + i], sacc[i]);
}
// pass 2: y = (x - mean_g) * rstd_g * gamma_c + beta_c [ReLU]; mean / rstd of the sample's groups once per block (shared memory)
template <typename TO>
__global__ void gn_rows_apply_kernel(const float* __restrict__ x, long long x_sample_stride, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, TO* __restrict__ out, long long out_sample_stride, int HW, int C,
                                     int groups, float eps, int relu, const double* __restrict__ stats) {
    extern __shared__ float2 mr[];                         // [groups] (mean, rstd)
    const int b = blockIdx.y, cg = C / groups;
    const double n = (double)HW * cg;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const double mu = stats[(size_t)b * 2 * groups + 2 * g] / n;
        const double var = stats[(size_t)b * 2 * groups + 2 * g + 1] / n - mu * mu;
        mr[g] = make_float2((float)mu, (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + (double)eps)));
    }
    __syncthreads();
    const size_t n4 = (size_t)HW * C / 4;
    const float* xb = x + (size_t)b * x_sample_stride;
    TO* ob = out + (size_t)b * out_sample_stride;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)((i * 4) % C);
        const float4 v = __ldg(reinterpret_cast<const float4*>(xb) + i);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c4)), bt = __ldg(reinterpret_cast<const float4*>(beta + c4));
        const float2 m0 = mr[c4 / cg], m1 = mr[(c4 + 1) / cg], m2 = mr[(c4 + 2) / cg], m3 = mr[(c4 + 3) / cg];
        float y0 = (v.x - m0.x) * m0.y * gm.x + bt.x, y1 = (v.y - m1.x) * m1.y * gm.y + bt.y;
        float y2 = (v.z - m2.x) * m2.y * gm.z + bt.z, y3 = (v.w - m3.x) * m3.y * gm.w + bt.w;
        if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
        store4(ob + i * 4, y0, y1, y2, y3);
    }
}

// ---- dst[b, oy, ox, :] += bilinear(src[b])(oy, ox), F.interpolate(mode="bilinear", align_corners=False) to (OH, OW) ----
__global__ void upsample_add_rows_kernel(const float* __restrict__ src, long long src_sample_stride, float* __restrict__ dst, int H, int W,
                                         int OH, int OW, int C) {
    const int b = blockIdx.y, c4n = C / 4;
    const size_t total = (size_t)OH * OW * c4n;
    const float sh = (float)H / OH, sw = (float)W / OW;
    const float* sb = src + (size_t)b * src_sample_stride;
    float* db = dst + (size_t)b * OH * OW * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        const size_t p = i / c4n;
        const int ox = (int)(p % OW), oy = (int)(p / OW);
        const float fy = fmaxf((oy + 0.5f) * sh - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sw - 0.5f, 0.f);
        const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float ly = fy - y0, lx = fx - x0;
        const float4 a = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y0 * W + x0) * C) + c4);
        const float4 bq = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y0 * W + x1) * C) + c4);
        const float4 c = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y1 * W + x0) * C) + c4);
        const float4 d = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y1 * W + x1) * C) + c4);
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        float4* o = reinterpret_cast<float4*>(db + p * C) + c4;
        float4 v = *o;
        v.x += w00 * a.x + w01 * bq.x + w10 * c.x + w11 * d.x;
        v.y += w00 * a.y + w01 * bq.y + w10 * c.y + w11 * d.y;
        v.z += w00 * a.z + w01 * bq.z + w10 * c.z + w11 * d.z;
        v.w += w00 * a.w + w01 * bq.w + w10 * c.w + w11 * d.w;
        *o = v;
    }
}

// ---- im2col of a 3x3 / stride 1 / zero-pad 1 convolution on rows: dst[b, y, x, (ky, kx, c)] = src[b, y+ky-1, x+kx-1, c] ----
template <typename TO>
__global__ void im2col3x3_rows_kernel(const float* __restrict__ src, TO* __restrict__ dst, int H, int W, int C) {
    const int b = blockIdx.y, c8n = C / 8;                 // 8 channels per thread: 16-byte stores of the bf16 operand
    const size_t total = (size_t)H * W * 9 * c8n;
    const float* sb = src + (size_t)b * H * W * C;
    TO* db = dst + (size_t)b * H * W * 9 * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % c8n);
        const size_t t = i / c8n;
        const int tap = (int)(t % 9);
        const size_t p = t / 9;
        const int x = (int)(p % W), y = (int)(p / W);
        const int sy = y + tap / 3 - 1, sx = x + tap % 3 - 1;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
            const float4* q = reinterpret_cast<const float4*>(sb + ((size_t)sy * W + sx) * C) + 2 * c8;
            v0 = __ldg(q);
            v1 = __ldg(q + 1);
        }
        TO* o = db + (p * 9 + tap) * C + c8 * 8;
        store8(o, v0, v1);
    }
}

// ---- the operand of the implicit-GEMM 3x3 convolution: dst[b, y + 1, x + 1, c] = bf16(src[b, y, x, c]) on a zero border ----
__global__ void pad_rows_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int H, int W, int C) {
    const int b = blockIdx.y, c8n = C / 8;
    const size_t total = (size_t)(H + 2) * (W + 2) * c8n;
    const float* sb = src + (size_t)b * H * W * C;
    bf16* db = dst + (size_t)b * (H + 2) * (W + 2) * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % c8n);
        const size_t p = i / c8n;
        const int px = (int)(p % (W + 2)), py = (int)(p / (W + 2));
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (py >= 1 && py <= H && px >= 1 && px <= W) {
            const float4* q = reinterpret_cast<const float4*>(sb + ((size_t)(py - 1) * W + (px - 1)) * C) + 2 * c8;
            v0 = __ldg(q);
            v1 = __ldg(q + 1);
        }
        store8(db + p * C + c8 * 8, v0, v1);
    }
}

// ---- one pass for what sits between the lateral convolution and the 3x3 output convolution of an FPN level
// (transformer_encoder_deform.py:346-349): dst[b, y + 1, x + 1, c] = bf16(GroupNorm(lat)[b, y, x, c] + bilinear(cur[b])(y, x)[c]) on a
// zero border — GroupNorm apply, `cur_fpn + F.interpolate(out[-1], ..., "bilinear")` and the zero-padded bf16 operand of the
// implicit-GEMM convolution, without the two fp32 maps in between (at 256^2 x 512 channels: 1.6 GB moved per 8 images instead of 5.9).
__global__ void fpn_gn_up_pad_kernel(const float* __restrict__ lat, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const double* __restrict__ stats, int groups, float eps, const float* __restrict__ cur,
                                     long long cur_sample_stride, int H, int W, bf16* __restrict__ dst, int OH, int OW, int C) {
    extern __shared__ float2 mr[];                         // [groups] (mean, rstd); unused without a norm
    const int b = blockIdx.y, c8n = C / 8, cg = stats ? C / groups : C;
    if (stats) {
        const double n = (double)OH * OW * cg;
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const double mu = stats[(size_t)b * 2 * groups + 2 * g] / n;
            const double var = stats[(size_t)b * 2 * groups + 2 * g + 1] / n - mu * mu;
            mr[g] = make_float2((float)mu, (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + (double)eps)));
        }
        __syncthreads();
    }
    // (32-bit index arithmetic per image — the host checks (OH + 2) (OW + 2) C / 8 < 2^31: the 64-bit divisions of the first version made
    // the kernel issue-bound, ncu: issue 73 %, DRAM 43 %)
    const unsigned total = (unsigned)(OH + 2) * (unsigned)(OW + 2) * (unsigned)c8n, pw = (unsigned)(OW + 2);
    const float sh = (float)H / OH, sw = (float)W / OW;
    const float* lb = lat + (size_t)b * OH * OW * C;
    const float* sb = cur + (size_t)b * cur_sample_stride;
    bf16* db = dst + (size_t)b * (OH + 2) * (OW + 2) * C;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned p = i / (unsigned)c8n;
        const int c8 = (int)(i - p * (unsigned)c8n);
        const int py = (int)(p / pw), px = (int)(p - (unsigned)py * pw);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (py >= 1 && py <= OH && px >= 1 && px <= OW) {
            const int oy = py - 1, ox = px - 1, c0 = c8 * 8;
            const float4* q = reinterpret_cast<const float4*>(lb + ((size_t)oy * OW + ox) * C + c0);
            const float4 l0 = __ldcs(q), l1 = __ldcs(q + 1);
            v[0] = l0.x; v[1] = l0.y; v[2] = l0.z; v[3] = l0.w; v[4] = l1.x; v[5] = l1.y; v[6] = l1.z; v[7] = l1.w;
            if (stats) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
                const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                if ((cg & 7) == 0) {                   // the thread's 8 channels lie in one group (one index division instead of eight)
                    const float2 m = mr[c0 / cg];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = (v[j] - m.x) * m.y * gm[j] + bt[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 m = mr[(c0 + j) / cg];
                        v[j] = (v[j] - m.x) * m.y * gm[j] + bt[j];
                    }
                }
            }
            const float fy = fmaxf((oy + 0.5f) * sh - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sw - 0.5f, 0.f);
            const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
            const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
            const float ly = fy - y0, lx = fx - x0;
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
            const float4* pa = reinterpret_cast<const float4*>(sb + ((size_t)y0 * W + x0) * C + c0);
            const float4* pb = reinterpret_cast<const float4*>(sb + ((size_t)y0 * W + x1) * C + c0);
            const float4* pc = reinterpret_cast<const float4*>(sb + ((size_t)y1 * W + x0) * C + c0);
            const float4* pd = reinterpret_cast<const float4*>(sb + ((size_t)y1 * W + x1) * C + c0);
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
                const float4 a = __ldg(pa + hq), bq = __ldg(pb + hq), c = __ldg(pc + hq), d = __ldg(pd + hq);
                v[4 * hq + 0] += w00 * a.x + w01 * bq.x + w10 * c.x + w11 * d.x;
                v[4 * hq + 1] += w00 * a.y + w01 * bq.y + w10 * c.y + w11 * d.y;
                v[4 * hq + 2] += w00 * a.z + w01 * bq.z + w10 * c.z + w11 * d.z;
                v[4 * hq + 3] += w00 * a.w + w01 * bq.w + w10 * c.w + w11 * d.w;
            }
        }
        store8(db + (size_t)p * C + c8 * 8, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
    }
}

// out = T(a + b[i mod b_n]): a positional embedding shared by every sample of the batch
template <typename T>
__global__ void add_cast_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b, T* __restrict__ out, size_t n4, size_t b_n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
        const float4 w = __ldg(reinterpret_cast<const float4*>(b) + (i % b_n4));
        out[i * 4] = from_float<T>(v.x + w.x); out[i * 4 + 1] = from_float<T>(v.y + w.y);
        out[i * 4 + 2] = from_float<T>(v.z + w.z); out[i * 4 + 3] = from_float<T>(v.w + w.w);
    }
}

}  // namespace
}  // namespace svb

using namespace svb;

static int nchw_to_rows_launch(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels, int pixels, int64_t sample_stride,
                               int64_t pixel_stride, const float* chan_add, cudaStream_t s) {
    dim3 grid((pixels + 31) / 32, (channels + 31) / 32, batch), block(32, 8);
    const bool ib = src_dtype == SVB_DTYPE_BF16, ob = dst_dtype == SVB_DTYPE_BF16;
    ProfScope prof(PC_OTHER, 0, (double)batch * channels * pixels * ((ib ? 2 : 4) + (ob ? 2 : 4)), s);
    if (ib && ob) nchw_to_rows_kernel<bf16, bf16><<<grid, block, 0, s>>>((const bf16*)src, (bf16*)dst, channels, pixels, sample_stride, pixel_stride, chan_add);
    else if (ib) nchw_to_rows_kernel<bf16, float><<<grid, block, 0, s>>>((const bf16*)src, (float*)dst, channels, pixels, sample_stride, pixel_stride, chan_add);
    else if (ob) nchw_to_rows_kernel<float, bf16><<<grid, block, 0, s>>>((const float*)src, (bf16*)dst, channels, pixels, sample_stride, pixel_stride, chan_add);
    else nchw_to_rows_kernel<float, float><<<grid, block, 0, s>>>((const float*)src, (float*)dst, channels, pixels, sample_stride, pixel_stride, chan_add);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_nchw_to_rows(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels, int pixels,
                                int64_t dst_sample_stride, svb_stream_t stream) {
    SVB_REQUIRE(src && dst && batch > 0 && channels > 0 && pixels > 0, "svb_nchw_to_rows: bad argument");
    if (dst_sample_stride <= 0) dst_sample_stride = (int64_t)pixels * channels;
    return nchw_to_rows_launch(src, src_dtype, dst, dst_dtype, batch, channels, pixels, dst_sample_stride, channels, nullptr, (cudaStream_t)stream);
}

extern "C" int svb_nchw_to_seq(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels, int pixels, const float* chan_add,
                               svb_stream_t stream) {
    SVB_REQUIRE(src && dst && batch > 0 && channels > 0 && pixels > 0, "svb_nchw_to_seq: bad argument");
    return nchw_to_rows_launch(src, src_dtype, dst, dst_dtype, batch, channels, pixels, channels, (int64_t)batch * channels, chan_add,
                               (cudaStream_t)stream);
}

extern "C" int svb_rows_to_nchw(const float* src, int64_t src_sample_stride, float* dst, int batch, int channels, int pixels,
                                svb_stream_t stream) {
    SVB_REQUIRE(src && dst && batch > 0 && channels > 0 && pixels > 0, "svb_rows_to_nchw: bad argument");
    if (src_sample_stride <= 0) src_sample_stride = (int64_t)pixels * channels;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((pixels + 31) / 32, (channels + 31) / 32, batch), block(32, 8);
    ProfScope prof(PC_OTHER, 0, (double)batch * channels * pixels * 8, s);
    rows_to_nchw_kernel<<<grid, block, 0, s>>>(src, src_sample_stride, dst, channels, pixels);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_groupnorm_rows(const float* x, int64_t x_sample_stride, const float* gamma, const float* beta, void* out, int out_dtype,
                                  int64_t out_sample_stride, int batch, int pixels, int channels, int groups, float eps, int relu,
                                  double* stats_ws, svb_stream_t stream) {
    SVB_REQUIRE(x && gamma && beta && out && stats_ws, "svb_groupnorm_rows: null argument");
    SVB_REQUIRE(groups > 0 && channels % groups == 0 && channels % 4 == 0 && (channels / groups) % 2 == 0,
                "svb_groupnorm_rows: channels %d / groups %d unsupported", channels, groups);
    if (x_sample_stride <= 0) x_sample_stride = (int64_t)pixels * channels;
    if (out_sample_stride <= 0) out_sample_stride = (int64_t)pixels * channels;
    SVB_REQUIRE(x_sample_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "svb_groupnorm_rows: x must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_NORM, 0, (double)batch * pixels * channels * (8 + (out_dtype == SVB_DTYPE_BF16 ? 2 : 4)), s, 2);
    SVB_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * groups * batch, s));
    const int rpb = 64;
    dim3 g1((pixels + rpb - 1) / rpb, batch);
    const int threads = channels / 4 >= 256 ? 256 : (channels / 4 >= 32 ? ((channels / 4 + 31) / 32) * 32 : 32);
    gn_rows_stats_kernel<<<g1, threads, sizeof(double) * 2 * groups, s>>>(x, x_sample_stride, pixels, channels, groups, rpb, stats_ws);
    SVB_CHECK_CUDA(cudaGetLastError());
    dim3 g2(grid_cap((size_t)pixels * channels / 4, 256), batch);
    if (out_dtype == SVB_DTYPE_BF16)
        gn_rows_apply_kernel<bf16><<<g2, 256, sizeof(float2) * groups, s>>>(x, x_sample_stride, gamma, beta, (bf16*)out, out_sample_stride, pixels, channels, groups, eps, relu, stats_ws);
    else
        gn_rows_apply_kernel<float><<<g2, 256, sizeof(float2) * groups, s>>>(x, x_sample_stride, gamma, beta, (float*)out, out_sample_stride, pixels, channels, groups, eps, relu, stats_ws);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_upsample_add_rows(const float* src, int64_t src_sample_stride, float* dst, int batch, int h, int w, int out_h, int out_w,
                                     int channels, svb_stream_t stream) {
    SVB_REQUIRE(src && dst && batch > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0 && channels % 4 == 0, "svb_upsample_add_rows: bad argument");
    if (src_sample_stride <= 0) src_sample_stride = (int64_t)h * w * channels;
    SVB_REQUIRE(src_sample_stride % 4 == 0, "svb_upsample_add_rows: sample stride must be a multiple of 4 elements");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OTHER, 0, (double)batch * out_h * out_w * channels * 12, s);
    dim3 grid(grid_cap((size_t)out_h * out_w * channels / 4, 256), batch);
    upsample_add_rows_kernel<<<grid, 256, 0, s>>>(src, src_sample_stride, dst, h, w, out_h, out_w, channels);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_im2col3x3_rows(const float* src, void* dst, int dst_dtype, int batch, int h, int w, int channels, svb_stream_t stream) {
    SVB_REQUIRE(src && dst && batch > 0 && h > 0 && w > 0 && channels % 8 == 0, "svb_im2col3x3_rows: bad argument (channels must be a multiple of 8)");
    cudaStream_t s = (cudaStream_t)stream;
    const bool ob = dst_dtype == SVB_DTYPE_BF16;
    ProfScope prof(PC_OTHER, 0, (double)batch * h * w * channels * (4 + 9 * (ob ? 2 : 4)), s);
    dim3 grid(grid_cap((size_t)h * w * 9 * channels / 8, 256), batch);
    if (ob) im2col3x3_rows_kernel<bf16><<<grid, 256, 0, s>>>(src, (bf16*)dst, h, w, channels);
    else im2col3x3_rows_kernel<float><<<grid, 256, 0, s>>>(src, (float*)dst, h, w, channels);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_conv3x3_rows(const float* src, const void* weight_bf16, const float* bias, float* out, void* padded_ws, int batch, int h, int w,
                                int cin, int cout, int relu, svb_stream_t stream) {
    SVB_REQUIRE(src && weight_bf16 && out && padded_ws && batch > 0 && h > 0 && w > 0 && cin % 8 == 0, "svb_conv3x3_rows: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    {
        ProfScope prof(PC_OTHER, 0, (double)batch * h * w * cin * 4 + (double)batch * (h + 2) * (w + 2) * cin * 2, s);
        dim3 grid(grid_cap((size_t)(h + 2) * (w + 2) * cin / 8, 256), batch);
        pad_rows_bf16_kernel<<<grid, 256, 0, s>>>(src, (bf16*)padded_ws, h, w, cin);
        SVB_CHECK_CUDA(cudaGetLastError());
    }
    Epilogue ep;
    ep.bias = bias;
    ep.out = out; ep.ldo = cout;
    ep.act = relu ? 2 : 0;
    return gemm_conv3x3_bf16_tc((const bf16*)padded_ws, (const bf16*)weight_bf16, 9 * cin, batch, h, w, cin, cout, ep, s);
}

extern "C" int svb_fpn_conv3x3_rows(const float* lateral, const float* gn_gamma, const float* gn_beta, int gn_groups, float gn_eps,
                                    const float* cur, int64_t cur_sample_stride, int cur_h, int cur_w, const void* weight_bf16, const float* bias,
                                    float* out, void* padded_ws, double* stats_ws, int batch, int h, int w, int cin, int cout, int relu,
                                    svb_stream_t stream) {
    SVB_REQUIRE(lateral && cur && weight_bf16 && out && padded_ws && batch > 0 && h > 0 && w > 0 && cur_h > 0 && cur_w > 0 && cin % 8 == 0,
                "svb_fpn_conv3x3_rows: bad argument");
    SVB_REQUIRE(((reinterpret_cast<uintptr_t>(lateral) | reinterpret_cast<uintptr_t>(cur) | reinterpret_cast<uintptr_t>(padded_ws)) & 15) == 0,
                "svb_fpn_conv3x3_rows: 16-byte aligned operands");
    if (cur_sample_stride <= 0) cur_sample_stride = (int64_t)cur_h * cur_w * cin;
    SVB_REQUIRE(cur_sample_stride % 4 == 0, "svb_fpn_conv3x3_rows: sample stride must be a multiple of 4 elements");
    SVB_REQUIRE((int64_t)(h + 2) * (w + 2) * (cin / 8) < (1ll << 31), "svb_fpn_conv3x3_rows: map of %d x %d x %d too large for the per-image index", h, w, cin);
    cudaStream_t s = (cudaStream_t)stream;
    const bool norm = gn_gamma != nullptr;
    if (norm) {
        SVB_REQUIRE(gn_beta && stats_ws && gn_groups > 0 && cin % gn_groups == 0 && (cin / gn_groups) % 2 == 0 && gn_groups <= 1024,
                    "svb_fpn_conv3x3_rows: channels %d / groups %d unsupported", cin, gn_groups);
        ProfScope prof(PC_NORM, 0, (double)batch * h * w * cin * 4, s);
        SVB_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * gn_groups * batch, s));
        const int rpb = 64, pixels = h * w;
        dim3 g1((pixels + rpb - 1) / rpb, batch);
        const int threads = cin / 4 >= 256 ? 256 : (cin / 4 >= 32 ? ((cin / 4 + 31) / 32) * 32 : 32);
        gn_rows_stats_kernel<<<g1, threads, sizeof(double) * 2 * gn_groups, s>>>(lateral, (long long)pixels * cin, pixels, cin, gn_groups, rpb, stats_ws);
        SVB_CHECK_CUDA(cudaGetLastError());
    }
    {
        ProfScope prof(PC_OTHER, 0, (double)batch * h * w * cin * 4 + (double)batch * cur_h * cur_w * cin * 4 + (double)batch * (h + 2) * (w + 2) * cin * 2, s);
        dim3 grid(grid_cap((size_t)(h + 2) * (w + 2) * cin / 8, 256), batch);
        fpn_gn_up_pad_kernel<<<grid, 256, sizeof(float2) * (norm ? gn_groups : 1), s>>>(lateral, gn_gamma, gn_beta, norm ? stats_ws : nullptr, gn_groups,
                                                                                      gn_eps, cur, cur_sample_stride, cur_h, cur_w, (bf16*)padded_ws,
                                                                                      h, w, cin);
        SVB_CHECK_CUDA(cudaGetLastError());
    }
    Epilogue ep;
    ep.bias = bias;
    ep.out = out; ep.ldo = cout;
    ep.act = relu ? 2 : 0;
    return gemm_conv3x3_bf16_tc((const bf16*)padded_ws, (const bf16*)weight_bf16, 9 * cin, batch, h, w, cin, cout, ep, s);
}

extern "C" int svb_add_cast_bcast(const float* a, const float* b, int64_t b_numel, void* out, int out_dtype, int64_t numel, svb_stream_t stream) {
    SVB_REQUIRE(a && b && out && numel >= 0 && b_numel > 0 && numel % 4 == 0 && b_numel % 4 == 0 && numel % b_numel == 0,
                "svb_add_cast_bcast: bad argument");
    if (numel == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool ob = out_dtype == SVB_DTYPE_BF16;
    ProfScope prof(PC_OTHER, 0, (double)numel * (8 + (ob ? 2 : 4)), s);
    if (ob) add_cast_bcast_kernel<bf16><<<grid_cap(numel / 4, 256), 256, 0, s>>>(a, b, (bf16*)out, numel / 4, b_numel / 4);
    else add_cast_bcast_kernel<float><<<grid_cap(numel / 4, 256), 256, 0, s>>>(a, b, (float*)out, numel / 4, b_numel / 4);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
