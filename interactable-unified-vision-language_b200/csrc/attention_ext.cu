// tcgen05 / TMEM global attention for token grids OTHER than the trained 64 x 64 one (scope row N3: the reference's live evaluation
// path pads COCO images to 1024 x 2048, configs/step1.yaml:211-212 -> a 64 x 128 grid; image_encoder.py:111-114, 319-330).
//
// Attention.forward + add_decomposed_rel_pos (image_encoder.py:239-255, 340-376) for a (gh x gw) grid, both sides multiples of 32:
//   * one CTA = 128 consecutive query tokens of one (image, head); S = Q K^T (SS-mode tcgen05.mma) and O += P V over 64-key tiles with
//     two S buffers (QK(h+2) runs behind PV(h) while the softmax group works on S(h+1)), P written back over S as bf16 (TMEM A
//     operand), exact-maximum online softmax with lazy rescaling — the pipeline of the X-Decoder cross-attention kernel (xattn_tc.cu);
//   * Q / K / V come straight out of the token-order qkv tensor by TMA; the key tiles are boxes of a 3-D view [3D, gw, B * gh]:
//     min(gw, 64) key columns x (64 / that) key rows, visited column block by column block, so that the w term of the bias is ONE
//     set of 64 registers per sweep and the h term one or two scalars per tile (softmax does not care about the key order);
//   * the decomposed rel-pos terms are read from tables in global memory, bias_h[token][head][j] = q . rel_pos_h'[j] and bias_w
//     likewise (rel_pos' = the table linearly resized to 2 g - 1 rows as get_rel_pos does, :319-330), which the encoder produces
//     with the tcgen05 GEMM per head (the in-kernel Q . R^T products of the 64 x 64 kernel need 2 g - 1 <= 128 TMEM columns).
// Off the benchmarked path; two CTAs per SM.
#include "attention_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace svb {
namespace {

template <int HD> struct ECfg {
    static constexpr int NST = 4;
    static constexpr int TAIL = HD - 64;
    static constexpr int Q_MAIN = 128 * 128, Q_TAIL = TAIL ? 128 * 32 : 0;
    static constexpr int KV_MAIN = 64 * 128, KV_TAIL = TAIL ? 64 * 32 : 0;
    static constexpr int KV = KV_MAIN + KV_TAIL;
    static constexpr int OFF_K = Q_MAIN + Q_TAIL, OFF_V = OFF_K + NST * KV, OFF_BAR = OFF_V + NST * KV;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static_assert(2 * (SMEM + 1024) <= 233472, "two CTAs per SM");
    static constexpr int B_QFULL = 0, B_FULL = 1, B_EMPTY = B_FULL + NST, B_SFULL = B_EMPTY + NST, B_PFULL = B_SFULL + 2,
                         B_PVDONE = B_PFULL + 2, B_ODONE = B_PVDONE + 1, B_COUNT = B_ODONE + 1;
    static constexpr int TM_S = 0, TM_O = 128, TM_COLS = 256;
};

struct ExtMaps {
    CUtensorMap q, qt;           // 2-D [3D, B*T]: boxes (64 | 16, 128)
    CUtensorMap kv, kvt;         // 3-D [3D, gw, B*gh]: boxes (64 | 16, BW, 64 / BW)
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <int HD>
__global__ void __launch_bounds__(192, 2)
attn_global_ext_kernel(const __grid_constant__ ExtMaps maps, const float* __restrict__ bias_h, const float* __restrict__ bias_w, int ld_h,
                       int ld_w, bf16* __restrict__ out, int D, int gh, int gw, int heads, float scale_log2) {
    using C = ECfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int T = gh * gw;
    const int bw = gw < 64 ? gw : 64;                          // key columns per tile
    const int rpt = 64 / bw;                                    // key rows per tile
    const int nblk = (gw + 63) / 64;                            // column blocks (sweeps)
    const int tps = gh / rpt;                                   // tiles per sweep
    const int nt = nblk * tps;                                  // key tiles
    const int colq = head * HD, colk = D + head * HD, colv = 2 * D + head * HD;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&maps.q);
        ptx::prefetch_tmap(&maps.kv);
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
            ptx::mbar_expect_tx(&bars[C::B_QFULL], C::Q_MAIN + C::Q_TAIL);
            ptx::tma_load_2d(sm, &maps.q, &bars[C::B_QFULL], colq, b * T + qt * 128);
            if (HD > 64) ptx::tma_load_2d(sm + C::Q_MAIN, &maps.qt, &bars[C::B_QFULL], colq + 64, b * T + qt * 128);
            for (int h = 0; h < nt; ++h) {
                const int st = h % C::NST;
                const int wb = h / tps, kr = h % tps;
                ptx::mbar_wait(&bars[C::B_EMPTY + st], ((h / C::NST) & 1) ^ 1);
                ptx::mbar_expect_tx(&bars[C::B_FULL + st], 2 * C::KV);
                uint8_t* k = sm + C::OFF_K + st * C::KV;
                uint8_t* v = sm + C::OFF_V + st * C::KV;
                // (columns past gw in the last block of a grid like 96 are out of bounds: zero-filled, and masked by the w terms)
                tma_load_3d(k, &maps.kv, &bars[C::B_FULL + st], colk, 64 * wb, b * gh + kr * rpt);
                tma_load_3d(v, &maps.kv, &bars[C::B_FULL + st], colv, 64 * wb, b * gh + kr * rpt);
                if (HD > 64) {
                    tma_load_3d(k + C::KV_MAIN, &maps.kvt, &bars[C::B_FULL + st], colk + 64, 64 * wb, b * gh + kr * rpt);
                    tma_load_3d(v + C::KV_MAIN, &maps.kvt, &bars[C::B_FULL + st], colv + 64, 64 * wb, b * gh + kr * rpt);
                }
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
            const uint32_t k = base + C::OFF_K + st * C::KV;
            issue_qk<HD>(tmem + C::TM_S + 64 * (h & 1), base, base + C::Q_MAIN, k, k + C::KV_MAIN, id_s);
            ptx::mma_commit_e(&bars[C::B_SFULL + (h & 1)]);
        };
        issue_s(0);
        if (nt > 1) issue_s(1);
#pragma unroll 1
        for (int h = 0; h < nt; ++h) {
            const int st = h % C::NST;
            ptx::mbar_wait(&bars[C::B_PFULL + (h & 1)], (h >> 1) & 1);      // P(h) is in TMEM
            ptx::tc_fence_after();
            const uint32_t v = base + C::OFF_V + st * C::KV;
            issue_pv<HD>(o_tm, tmem + C::TM_S + 64 * (h & 1), v, v + C::KV_MAIN, 4, h > 0);
            ptx::mma_commit_e(&bars[C::B_PVDONE]);
            ptx::mma_commit_e(&bars[C::B_EMPTY + st]);
            if (h == nt - 1) ptx::mma_commit_e(&bars[C::B_ODONE]);
            if (h + 2 < nt) issue_s(h + 2);
        }
    } else {
        // ===================== softmax group: one query token per thread =====================
        const int t = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + C::TM_S, o_tmem = tmem + lane_off + C::TM_O;
        const int tok = qt * 128 + t;
        const int qh = tok / gw, qw = tok % gw;
        const float* bh_row = bias_h + ((size_t)(b * T + tok) * heads + head) * ld_h + (qh + gh - 1);      // entry for key row kh: [-kh]
        const float* bw_row = bias_w + ((size_t)(b * T + tok) * heads + head) * ld_w + (qw + gw - 1);      // entry for key column kw: [-kw]
        float m_ref = -INFINITY;
        f32x2 l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);
        const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
        float bwl[64];
#pragma unroll 1
        for (int h = 0; h < nt; ++h) {
            const int wb = h / tps, kr = h % tps;
            if (kr == 0) {
                // a new column block: this row's 64 w terms (log2 units); columns past the grid -> -inf (their keys are zero-filled)
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int kw = 64 * wb + (bw == 64 ? j : (j % bw));
                    bwl[j] = (kw < gw) ? __ldg(bw_row - kw) * LOG2E : -INFINITY;
                }
            }
            // the h terms of this tile's key rows
            const float d0 = __ldg(bh_row - kr * rpt) * LOG2E;
            const float d1 = (rpt > 1) ? __ldg(bh_row - (kr * rpt + 1)) * LOG2E : d0;
            const uint32_t s_h = s_tmem + 64 * (h & 1);
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
            // pass 1: x = s * scale + w term + h term (key rows of 32 columns: the second half of the tile is the next key row)
            f32x2 xa[16], xb[16];
            float mx0 = -INFINITY, mx1 = -INFINITY;
            const f32x2 da2 = f2_pack(d0, d0), db2 = f2_pack(d1, d1);
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                xa[e / 2] = f2_add(f2_fma(f2_pack(__uint_as_float(va[e]), __uint_as_float(va[e + 1])), sc2, f2_pack(bwl[e], bwl[e + 1])), da2);
                xb[e / 2] = f2_add(f2_fma(f2_pack(__uint_as_float(vb[e]), __uint_as_float(vb[e + 1])), sc2, f2_pack(bwl[32 + e], bwl[33 + e])), db2);
                float a0, a1, b0, b1;
                f2_unpack(xa[e / 2], a0, a1);
                f2_unpack(xb[e / 2], b0, b1);
                mx0 = fmax3(mx0, a0, a1);
                mx1 = fmax3(mx1, b0, b1);
            }
            const float tmax = fmaxf(mx0, mx1);
            const bool need = (h == 0) || (tmax - m_ref > RESCALE_THRESHOLD);
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? tmax : m_ref;
                if (h > 0) {
                    const float alpha = need ? ptx::ex2_approx(m_ref - m_new) : 1.0f;
                    ptx::mbar_wait(&bars[C::B_PVDONE], (h - 1) & 1);                      // O holds tiles 0..h-1
                    ptx::tc_fence_after();
                    uint32_t r[16];
#pragma unroll
                    for (int c0 = 0; c0 < HD; c0 += 16) {
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
            // pass 2: P = exp2(x - m_ref) -> bf16 -> TMEM over the consumed S columns
            const f32x2 nr2 = f2_pack(-m_ref, -m_ref);
#define SVB_EXT_PASS(X, CHUNK)                                                                         \
            {                                                                                            \
                uint32_t pk[16];                                                                         \
                _Pragma("unroll") for (int e = 0; e < 16; e += 2) {                                      \
                    float a0, a1, a2, a3;                                                                \
                    f2_unpack(f2_add(X[e], nr2), a0, a1);                                                \
                    f2_unpack(f2_add(X[e + 1], nr2), a2, a3);                                            \
                    const float p0 = ptx::ex2_approx(a0), p1 = ptx::ex2_approx(a1), p2 = ptx::ex2_approx(a2), p3 = ptx::ex2_approx(a3); \
                    l01 = f2_add(l01, f2_pack(p0, p1));                                                  \
                    l23 = f2_add(l23, f2_pack(p2, p3));                                                  \
                    pk[e] = pack_bf16x2(p0, p1);                                                         \
                    pk[e + 1] = pack_bf16x2(p2, p3);                                                     \
                }                                                                                        \
                ptx::tmem_st_x16(s_h + 16 * (CHUNK), pk);                                                \
            }
            SVB_EXT_PASS(xa, 0)
            SVB_EXT_PASS(xb, 1)
#undef SVB_EXT_PASS
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[C::B_PFULL + ((nt - 1) & 1)]);         // P(nt-1)
        // ---- epilogue: O / l at the token's own position ----
        ptx::mbar_wait(&bars[C::B_ODONE], 0);
        ptx::tc_fence_after();
        float l0, l1, l2, l3;
        f2_unpack(l01, l0, l1);
        f2_unpack(l23, l2, l3);
        const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
        store_row<HD>(out + (size_t)(b * T + tok) * D + head * HD, o_tmem, inv);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}

template <int HD>
int launch_ext(const AttnTcParams& p, cudaStream_t stream) {
    using C = ECfg<HD>;
    const int D = p.heads * p.hd, gh = p.grid_h, gw = p.grid_w, T = gh * gw;
    const int bw = gw < 64 ? gw : 64, rpt = 64 / bw;
    ExtMaps m;
    int rc;
    {
        const uint64_t d2[2] = {(uint64_t)3 * D, (uint64_t)p.batch * T};
        const uint64_t s2[1] = {(uint64_t)3 * D * 2};
        const uint32_t bq[2] = {64, 128}, bqt[2] = {16, 128};
        if ((rc = encode_tmap_nd_bf16(&m.q, p.qkv, 2, d2, s2, bq, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&m.qt, p.qkv, 2, d2, s2, bqt, 32))) return rc;
        const uint64_t d3[3] = {(uint64_t)3 * D, (uint64_t)gw, (uint64_t)p.batch * gh};
        const uint64_t s3[2] = {(uint64_t)3 * D * 2, (uint64_t)gw * 3 * D * 2};
        const uint32_t bk[3] = {64, (uint32_t)bw, (uint32_t)rpt}, bkt[3] = {16, (uint32_t)bw, (uint32_t)rpt};
        if ((rc = encode_tmap_nd_bf16(&m.kv, p.qkv, 3, d3, s3, bk, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&m.kvt, p.qkv, 3, d3, s3, bkt, 32))) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(attn_global_ext_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    const float scale_log2 = LOG2E / sqrtf((float)HD);
    dim3 grid(T / 128, p.heads, p.batch);
    attn_global_ext_kernel<HD><<<grid, 192, C::SMEM, stream>>>(m, p.bias_h, p.bias_w, p.bias_ld_h, p.bias_ld_w, p.out, D, gh, gw, p.heads, scale_log2);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// global attention on a (grid_h x grid_w) token grid with the rel-pos terms in global tables (see the header of this file)
int attention_global_ext(const AttnTcParams& p, cudaStream_t stream) {
    SVB_REQUIRE(p.grid_h > 0 && p.grid_w > 0 && p.grid_h % 32 == 0 && p.grid_w % 32 == 0,
                "attention_global_ext: token grid %d x %d: both sides must be multiples of 32", p.grid_h, p.grid_w);
    SVB_REQUIRE(p.hd == 64 || p.hd == 80, "attention_global_ext: head_dim %d is not 64 or 80", p.hd);
    SVB_REQUIRE(p.bias_h && p.bias_w && p.bias_ld_h >= 2 * p.grid_h - 1 && p.bias_ld_w >= 2 * p.grid_w - 1, "attention_global_ext: rel-pos term tables missing");
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0, "attention_global_ext: pointers must be 16-byte aligned");
    const double D_ = (double)p.heads * p.hd, S = (double)p.grid_h * p.grid_w;
    ProfScope prof(PC_ATTN_GLOBAL, (double)p.batch * (4.0 * S * S * D_), (double)p.batch * S * 4.0 * D_ * 2, stream);
    return p.hd == 64 ? launch_ext<64>(p, stream) : launch_ext<80>(p, stream);
}

}  // namespace svb
