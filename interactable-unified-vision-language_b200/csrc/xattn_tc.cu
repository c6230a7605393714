// tcgen05 / TMEM masked cross-attention core of the X-Decoder layers (scope row N4): nn.MultiheadAttention's
// softmax((q / sqrt(d)) k^T + mask) v inside CrossAttentionLayer.forward_post (modeling/interface/modules.py:95-106) and
// SelfAttentionLayer (:37-47), for <= 128 query tokens (101 in step1.yaml) over `keys` image positions, head_dim 64, bf16.
//
//   * one CTA = one 128-row query tile of one (key split, head, image): Q (queries, batch, heads*64) and K / V (keys, batch,
//     heads*64) are read sequence-first AS THE REFERENCE PASSES THEM, by TMA boxes of 64 columns at column (image, head) — no
//     permute / head-split copy; rows past `queries` are zero-filled by the TMA;
//   * S = Q K^T (SS-mode tcgen05.mma, 64-key tiles, two S buffers so that QK(h+2) runs behind PV(h) while the softmax group works on
//     S(h+1)) and O += P V (P written back over S as bf16, the TMEM A operand; V as the MN-major B operand);
//   * the boolean attention mask (batch*heads, queries, keys), non-zero = not allowed, is read 16 bytes at a time per row and
//     applied to S in registers (masked -> -inf); exact-maximum online softmax in fp32 with lazy rescaling of O (as the global
//     encoder kernel);
//   * keys are split over CTAs flash-decoding style when (image, head) pairs alone cannot fill the GPU; partial (m, l, O) go to a
//     workspace and a combine kernel normalises (one split: the kernel writes the output itself).
// A row whose keys are all masked yields NaN (0 / 0), as torch does.
#include "attention_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace svb {
namespace {

struct XCfg {
    static constexpr int NST = 4;
    static constexpr int Q_BYTES = 128 * 128;                 // 128 rows x 64 bf16, 128B swizzle
    static constexpr int KV_BYTES = 64 * 128;                 // 64 keys x 64 bf16
    static constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + NST * KV_BYTES, OFF_BAR = OFF_V + NST * KV_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int B_QFULL = 0, B_FULL = 1, B_EMPTY = B_FULL + NST, B_SFULL = B_EMPTY + NST, B_PFULL = B_SFULL + 2,
                         B_PVDONE = B_PFULL + 2, B_ODONE = B_PVDONE + 1, B_COUNT = B_ODONE + 1;
    static constexpr int TM_S = 0, TM_O = 128, TM_COLS = 256;
};

// partial results of one key split: [m (log2 units), l, o[64]] per (image, head, split, query)
constexpr int XPART = 66;

template <bool PARTIAL>
__global__ void __launch_bounds__(192, 2)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                const uint8_t* __restrict__ mask, bf16* __restrict__ out, float* __restrict__ part, int Q, int HW, int B, int heads,
                int tiles_per_split, float scale_log2) {
    using C = XCfg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int split = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int Cc = heads * 64;
    const int col = b * Cc + head * 64;                         // column of this (image, head) in the [tokens][batch * C] views
    const int total_tiles = HW / 64;
    const int t0 = split * tiles_per_split;
    const int nt = min(tiles_per_split, total_tiles - t0);      // key tiles of this CTA (>= 1 by construction of the grid)

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tm_q);
        ptx::prefetch_tmap(&tm_k);
        ptx::prefetch_tmap(&tm_v);
        for (int s = 0; s < C::B_COUNT; ++s) {
            const bool by_threads = (s >= C::B_PFULL && s < C::B_PFULL + 2);
            ptx::mbar_init(&bars[s], by_threads ? 128 : 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, C::TM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            ptx::mbar_expect_tx(&bars[C::B_QFULL], C::Q_BYTES);
            ptx::tma_load_2d(sm, &tm_q, &bars[C::B_QFULL], col, 0);
            for (int h = 0; h < nt; ++h) {
                const int st = h % C::NST;
                ptx::mbar_wait(&bars[C::B_EMPTY + st], ((h / C::NST) & 1) ^ 1);
                ptx::mbar_expect_tx(&bars[C::B_FULL + st], 2 * C::KV_BYTES);
                ptx::tma_load_2d(sm + C::OFF_K + st * C::KV_BYTES, &tm_k, &bars[C::B_FULL + st], col, (t0 + h) * 64);
                ptx::tma_load_2d(sm + C::OFF_V + st * C::KV_BYTES, &tm_v, &bars[C::B_FULL + st], col, (t0 + h) * 64);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (all lanes run the loop, one elected lane issues) =====================
        constexpr uint32_t id_s = ptx::make_idesc_bf16(128, 64, 0, 0);
        const uint32_t o_tm = tmem + C::TM_O;
        ptx::mbar_wait(&bars[C::B_QFULL], 0);
        auto issue_s = [&](int h) {
            const int st = h % C::NST;
            ptx::mbar_wait(&bars[C::B_FULL + st], (h / C::NST) & 1);
            ptx::tc_fence_after();
            issue_qk<64>(tmem + C::TM_S + 64 * (h & 1), base, 0, base + C::OFF_K + st * C::KV_BYTES, 0, id_s);
            ptx::mma_commit_e(&bars[C::B_SFULL + (h & 1)]);
        };
        issue_s(0);
        if (nt > 1) issue_s(1);
#pragma unroll 1
        for (int h = 0; h < nt; ++h) {
            const int st = h % C::NST;
            ptx::mbar_wait(&bars[C::B_PFULL + (h & 1)], (h >> 1) & 1);      // P(h) is in TMEM
            ptx::tc_fence_after();
            issue_pv<64>(o_tm, tmem + C::TM_S + 64 * (h & 1), base + C::OFF_V + st * C::KV_BYTES, 0, 4, h > 0);
            ptx::mma_commit_e(&bars[C::B_PVDONE]);
            ptx::mma_commit_e(&bars[C::B_EMPTY + st]);                     // K(h) (read by QK(h), earlier) and V(h) are consumed
            if (h == nt - 1) ptx::mma_commit_e(&bars[C::B_ODONE]);
            if (h + 2 < nt) issue_s(h + 2);                                // overwrites S(h) / P(h) behind the PV above (in order)
        }
    } else {
        // ===================== softmax group: one query row per thread =====================
        const int t = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + C::TM_S, o_tmem = tmem + lane_off + C::TM_O;
        const bool live = t < Q;
        const uint8_t* mrow = mask ? mask + ((size_t)(b * heads + head) * Q + (live ? t : 0)) * HW + (size_t)t0 * 64 : nullptr;
        float m_ref = -INFINITY;
        f32x2 l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);
        const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
#pragma unroll 1
        for (int h = 0; h < nt; ++h) {
            const uint32_t s_h = s_tmem + 64 * (h & 1);
            // this tile's 64 mask bytes (issued before the wait for S: the loads overlap it)
            uint4 mk[4];
            if (mrow) {
#pragma unroll
                for (int j = 0; j < 4; ++j) mk[j] = __ldg(reinterpret_cast<const uint4*>(mrow + (size_t)h * 64) + j);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) mk[j] = make_uint4(0, 0, 0, 0);
            }
            uint32_t va[32], vb[32];
            ptx::mbar_wait(&bars[C::B_SFULL + (h & 1)], (h >> 1) & 1);
            ptx::tc_fence_after();
            ptx::tmem_ld_x32(s_h, va);
            ptx::tmem_ld_x32(s_h + 32, vb);
            if (h > 0) {                                           // hand P(h-1) over (its stores were issued last iteration)
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&bars[C::B_PFULL + ((h - 1) & 1)]);
            }
            ptx::tmem_ld_wait_dep(va);
            ptx::tmem_ld_wait_dep(vb);
            // pass 1: x = s * scale (log2 units), masked keys -> -inf; exact tile maximum
            float x[64];
            float mx0 = -INFINITY, mx1 = -INFINITY;
            const uint32_t mw[16] = {mk[0].x, mk[0].y, mk[0].z, mk[0].w, mk[1].x, mk[1].y, mk[1].z, mk[1].w,
                                     mk[2].x, mk[2].y, mk[2].z, mk[2].w, mk[3].x, mk[3].y, mk[3].z, mk[3].w};
#pragma unroll
            for (int e = 0; e < 64; e += 2) {
                const uint32_t r0 = e < 32 ? va[e] : vb[e - 32], r1 = e < 32 ? va[e + 1] : vb[e - 31];
                float a0, a1;
                f2_unpack(f2_mul(f2_pack(__uint_as_float(r0), __uint_as_float(r1)), sc2), a0, a1);
                const uint32_t w = mw[e >> 2] >> (8 * (e & 3));    // bytes e, e + 1 of the row's mask
                x[e] = (w & 0xFFu) ? -INFINITY : a0;
                x[e + 1] = (w & 0xFF00u) ? -INFINITY : a1;
                if (e & 2) mx1 = fmax3(mx1, x[e], x[e + 1]); else mx0 = fmax3(mx0, x[e], x[e + 1]);
            }
            const float tmax = fmaxf(mx0, mx1);
            const bool need = (tmax > m_ref + RESCALE_THRESHOLD) || (m_ref == -INFINITY && tmax > -INFINITY);
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? tmax : m_ref;
                if (h > 0) {
                    const float alpha = need ? ptx::ex2_approx(m_ref - m_new) : 1.0f;      // 0 when nothing was unmasked so far
                    ptx::mbar_wait(&bars[C::B_PVDONE], (h - 1) & 1);                      // O holds tiles 0..h-1
                    ptx::tc_fence_after();
                    uint32_t r[16];
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) {
                        ptx::tmem_ld_x16(o_tmem + c0, r);
                        ptx::tmem_ld_wait_dep(r);
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        ptx::tmem_st_x16(o_tmem + c0, r);
                    }
                    const f32x2 al2 = f2_pack(alpha, alpha);
                    l01 = f2_mul(l01, al2);
                    l23 = f2_mul(l23, al2);
                }
                m_ref = m_new;
            }
            // pass 2: P = exp2(x - m_ref) (0 for masked keys; nothing unmasked yet: reference 0 keeps -inf - ref = -inf)
            const float ref = (m_ref == -INFINITY) ? 0.f : m_ref;
            const f32x2 nr2 = f2_pack(-ref, -ref);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float a0, a1, a2, a3;
                    f2_unpack(f2_add(f2_pack(x[16 * c + e], x[16 * c + e + 1]), nr2), a0, a1);
                    f2_unpack(f2_add(f2_pack(x[16 * c + e + 2], x[16 * c + e + 3]), nr2), a2, a3);
                    const float p0 = ptx::ex2_approx(a0), p1 = ptx::ex2_approx(a1), p2 = ptx::ex2_approx(a2), p3 = ptx::ex2_approx(a3);
                    l01 = f2_add(l01, f2_pack(p0, p1));
                    l23 = f2_add(l23, f2_pack(p2, p3));
                    pk[e / 2] = pack_bf16x2(p0, p1);
                    pk[e / 2 + 1] = pack_bf16x2(p2, p3);
                }
                ptx::tmem_st_x8(s_h + 8 * c, pk);
            }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[C::B_PFULL + ((nt - 1) & 1)]);         // P(nt-1)
        // ---- epilogue ----
        ptx::mbar_wait(&bars[C::B_ODONE], 0);
        ptx::tc_fence_after();
        float l0, l1, l2, l3;
        f2_unpack(l01, l0, l1);
        f2_unpack(l23, l2, l3);
        const float l = (l0 + l1) + (l2 + l3);
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_x32(o_tmem, v0);
        ptx::tmem_ld_x32(o_tmem + 32, v1);
        ptx::tmem_ld_wait_dep(v0);
        ptx::tmem_ld_wait_dep(v1);
        if (live) {
            if (PARTIAL) {
                float* dst = part + ((((size_t)b * heads + head) * gridDim.x + split) * Q + t) * XPART;
                dst[0] = m_ref;
                dst[1] = l;
#pragma unroll
                for (int e = 0; e < 32; ++e) { dst[2 + e] = __uint_as_float(v0[e]); dst[34 + e] = __uint_as_float(v1[e]); }
            } else {
                const float inv = 1.0f / l;                        // l == 0 (every key masked): 0 * inf = NaN, as torch
                uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)t * B + b) * Cc + head * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(v0[8 * j]) * inv, __uint_as_float(v0[8 * j + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(v0[8 * j + 2]) * inv, __uint_as_float(v0[8 * j + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(v0[8 * j + 4]) * inv, __uint_as_float(v0[8 * j + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(v0[8 * j + 6]) * inv, __uint_as_float(v0[8 * j + 7]) * inv);
                    dst[j] = u;
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(v1[8 * j]) * inv, __uint_as_float(v1[8 * j + 1]) * inv);
                    w.y = pack_bf16x2(__uint_as_float(v1[8 * j + 2]) * inv, __uint_as_float(v1[8 * j + 3]) * inv);
                    w.z = pack_bf16x2(__uint_as_float(v1[8 * j + 4]) * inv, __uint_as_float(v1[8 * j + 5]) * inv);
                    w.w = pack_bf16x2(__uint_as_float(v1[8 * j + 6]) * inv, __uint_as_float(v1[8 * j + 7]) * inv);
                    dst[4 + j] = w;
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}

// out = sum_j 2^(m_j - m) o_j / sum_j 2^(m_j - m) l_j over the key splits; one thread per (query, channel)
__global__ void xattn_tc_combine_kernel(const float* __restrict__ part, bf16* __restrict__ out, int Q, int B, int heads, int nsplit) {
    const int head = blockIdx.y, b = blockIdx.z;
    const int t = blockIdx.x * (blockDim.x / 64) + threadIdx.x / 64, d = threadIdx.x % 64;
    if (t >= Q) return;
    const float* p0 = part + (((size_t)b * heads + head) * nsplit * Q + t) * XPART;
    const size_t stride = (size_t)Q * XPART;
    float m = -INFINITY;
    for (int j = 0; j < nsplit; ++j) m = fmaxf(m, p0[j * stride]);
    float l = 0.f, o = 0.f;
    for (int j = 0; j < nsplit; ++j) {
        const float mj = p0[j * stride];
        const float w = (mj == -INFINITY) ? 0.f : exp2f(mj - m);   // a split with every key masked contributes nothing
        l += w * p0[j * stride + 1];
        o += w * p0[j * stride + 2 + d];
    }
    out[((size_t)t * B + b) * (heads * 64) + head * 64 + d] = __float2bfloat16_rn(o / l);   // 0 / 0 = NaN for a fully masked row
}

}  // namespace

// bf16, head_dim 64, queries <= 128, keys a multiple of 64, mask rows 16-byte aligned: the tcgen05 path.  Returns 1 if it launched.
int xattn_tc_supported(int dtype_bf16, int queries, int keys, int head_dim, const void* q, const void* k, const void* v, const void* mask,
                       const void* out, int batch, int heads) {
    if (!dtype_bf16 || head_dim != 64 || queries < 1 || queries > 128 || keys < 64 || (keys % 64) != 0) return 0;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) return 0;
    if (mask && (reinterpret_cast<uintptr_t>(mask) & 15)) return 0;
    return 1;
}

int xattn_tc_splits(int keys, int batch, int heads) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = keys / 64;
    int nsplit = 1;
    // two CTAs fit an SM: split the keys until the grid fills the GPU, keeping at least 4 key tiles per split
    while ((long)nsplit * batch * heads < 2L * sms && tiles / (nsplit * 2) >= 4) nsplit *= 2;
    return nsplit;
}

int xattn_tc_launch(const bf16* q, const bf16* k, const bf16* v, const uint8_t* mask, bf16* out, float* workspace, int64_t workspace_floats,
                    int queries, int keys, int batch, int heads, cudaStream_t stream) {
    using C = XCfg;
    const int Cc = heads * 64;
    const int tiles = keys / 64;
    int nsplit = xattn_tc_splits(keys, batch, heads);
    while (nsplit > 1 && (int64_t)batch * heads * nsplit * queries * XPART > workspace_floats) nsplit /= 2;
    const int tps = (tiles + nsplit - 1) / nsplit;
    nsplit = (tiles + tps - 1) / tps;
    CUtensorMap mq, mk, mv;
    int rc;
    {
        const uint64_t dq[2] = {(uint64_t)batch * Cc, (uint64_t)queries}, dk[2] = {(uint64_t)batch * Cc, (uint64_t)keys};
        const uint64_t st[1] = {(uint64_t)batch * Cc * 2};
        const uint32_t bq[2] = {64, 128}, bk[2] = {64, 64};
        if ((rc = encode_tmap_nd_bf16(&mq, q, 2, dq, st, bq, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&mk, k, 2, dk, st, bk, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&mv, v, 2, dk, st, bk, 128))) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(xattn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        SVB_CHECK_CUDA(cudaFuncSetAttribute(xattn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    const float scale_log2 = LOG2E / sqrtf(64.0f);
    dim3 grid(nsplit, heads, batch);
    ProfScope prof(PC_OTHER, 4.0 * batch * heads * (double)queries * keys * 64, (double)batch * keys * Cc * 2 * 2, stream, nsplit > 1 ? 2 : 1);
    if (nsplit > 1) {
        xattn_tc_kernel<true><<<grid, 192, C::SMEM, stream>>>(mq, mk, mv, mask, out, workspace, queries, keys, batch, heads, tps, scale_log2);
        SVB_CHECK_CUDA(cudaGetLastError());
        dim3 g2((queries + 3) / 4, heads, batch);
        xattn_tc_combine_kernel<<<g2, 256, 0, stream>>>(workspace, out, queries, batch, heads, nsplit);
    } else {
        xattn_tc_kernel<false><<<grid, 192, C::SMEM, stream>>>(mq, mk, mv, mask, out, workspace, queries, keys, batch, heads, tps, scale_log2);
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace svb
