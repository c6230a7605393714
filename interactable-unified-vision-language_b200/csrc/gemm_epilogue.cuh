// Fused GEMM epilogue shared by the tcgen05 GEMM kernels: one thread owns one output row and 32 consecutive columns.
//   out[r, c] = act(acc[r, c] + bias[c]) + resid[(resid_mod ? r % resid_mod : r), c], optional GroupNorm(1,C) statistics.
#pragma once
#include "common.cuh"

namespace svb {

// residual values of one full 32-column chunk, fetched ahead of the accumulator so the HBM latency overlaps the
// previous chunk's work (the in-place fp32 residual stream of proj / lin2 is the epilogue's longest-latency access)
struct ResidChunk {
    float4 r[8];
    bool valid = false;
};
__device__ __forceinline__ void prefetch_resid(const Epilogue& ep, int row, int col0, int N, bool row_ok, ResidChunk& rc) {
    rc.valid = false;
    if (!ep.resid || !row_ok || col0 + 32 > N) return;
    const int rr = ep.resid_mod ? (row % ep.resid_mod) : row;
    const float4* r4 = reinterpret_cast<const float4*>(ep.resid + (size_t)rr * ep.ldr + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) rc.r[j] = r4[j];
    rc.valid = true;
}

__device__ __forceinline__ void epilogue_chunk(const Epilogue& ep, int row, int col0, int M, int N, bool row_ok,
                                               uint32_t (&raw)[32], float& s_sum, float& s_sq, const ResidChunk* pre = nullptr,
                                               uint32_t bias_smem = 0) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    const bool full = (col0 + 32 <= N);
    if (!row_ok) return;
    const size_t orow = epilogue_out_row(ep, row);
    if (full) {
        if (ep.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b;
                if (bias_smem) {   // this chunk's 32 bias values staged in shared memory by the epilogue warps (broadcast read)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_smem + 16 * j));
                } else {
                    b = __ldg(b4 + j);
                }
                v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
        }
        if (ep.stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { s_sum += v[j]; s_sq += v[j] * v[j]; }
        }
        if (ep.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_pair(v[j], v[j + 1]);
        } else if (ep.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ep.resid) {
            const int rr = ep.resid_mod ? (row % ep.resid_mod) : row;
            const float4* r4 = reinterpret_cast<const float4*>(ep.resid + (size_t)rr * ep.ldr + col0);
            const bool use_pre = pre && pre->valid;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 r = use_pre ? pre->r[j] : r4[j];
                v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
            }
        }
        if (ep.out_bf16) {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + orow * ep.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                o[j] = u;
            }
        } else {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow * ep.ldo + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    } else {
        const int rr = ep.resid ? (ep.resid_mod ? (row % ep.resid_mod) : row) : 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = col0 + j;
            if (c < N) {
                float x = v[j] + (ep.bias ? __ldg(ep.bias + c) : 0.f);
                if (ep.stats) { s_sum += x; s_sq += x * x; }
                if (ep.act == 1) x = gelu_fast(x);
                else if (ep.act == 2) x = fmaxf(x, 0.f);
                if (ep.resid) x += ep.resid[(size_t)rr * ep.ldr + c];
                if (ep.out_bf16) reinterpret_cast<bf16*>(ep.out)[orow * ep.ldo + c] = __float2bfloat16_rn(x);
                else reinterpret_cast<float*>(ep.out)[orow * ep.ldo + c] = x;
            }
        }
    }
}

}  // namespace svb
