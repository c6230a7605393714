// Device helpers shared by the tcgen05 attention kernels (attention_tc.cu: global + two-group windowed kernel; attention_win3.cu:
// the pipelined windowed kernel): MMA issue sequences, the FMA-pipe exp2, the window softmax tile, small utilities.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace svb {
int encode_tmap_nd_bf16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes);

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units: P stays below 2^8 relative to the reference maximum

__device__ __forceinline__ uint64_t desc_k128(uint32_t addr) { return ptx::make_smem_desc(addr, 0, 1024, ptx::LAYOUT_SW128); }
__device__ __forceinline__ uint64_t desc_k32(uint32_t addr) { return ptx::make_smem_desc(addr, 0, 256, ptx::LAYOUT_SW32); }

// The issue helpers below are called by ALL lanes of an issuer warp under warp-uniform control flow; one elected lane issues
// (ptx::mma_f16_*_e: uniform-register operands, no per-MMA waterfall loop).
// D[128 x N] (+)= A[128 x HD] * B[N x HD]^T, both K-major: 64 columns in a 128B-swizzled tile (+ 16 in a 32B-swizzled tile)
template <int HD>
__device__ __forceinline__ void issue_qk(uint32_t d_tmem, uint32_t a_main, uint32_t a_tail, uint32_t b_main, uint32_t b_tail,
                                         uint32_t idesc) {
    const uint64_t da = desc_k128(a_main), db = desc_k128(b_main);
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::mma_f16_ss_e(d_tmem, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
    if (HD > 64) ptx::mma_f16_ss_e(d_tmem, desc_k32(a_tail), desc_k32(b_tail), idesc, 1u);
}

// O[128 x HD] (+)= P[128 x 16*ksteps] (bf16 in TMEM, two per column) * V[keys x HD] (MN-major in smem)
template <int HD>
__device__ __forceinline__ void issue_pv(uint32_t o_tmem, uint32_t p_tmem, uint32_t v_main, uint32_t v_tail, int ksteps,
                                         bool accumulate) {
    constexpr uint32_t id_main = ptx::make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t id_tail = ptx::make_idesc_bf16(128, 16, 0, 1);
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t acc = (accumulate || k) ? 1u : 0u;
        ptx::mma_f16_ts_e(o_tmem, p_tmem + 8 * k, ptx::make_smem_desc(v_main + k * 2048, 0, 1024, ptx::LAYOUT_SW128), id_main, acc);
        if (HD > 64)
            ptx::mma_f16_ts_e(o_tmem + 64, p_tmem + 8 * k, ptx::make_smem_desc(v_tail + k * 512, 0, 256, ptx::LAYOUT_SW32), id_tail, acc);
    }
}
// The same product as ONE MMA of N = HD per K step: V in two 64-element atoms along N, `atom_stride` bytes apart (the leading-
// dimension byte offset of an MN-major operand; pinned by tests/test_gpu_probe.py).  Measured (tools/mma_rate.py): a tcgen05.mma
// with its A operand in TMEM costs >= 44.5 cycles whatever N is, so the N = 64 + N = 16 pair costs 89 cycles per K step against
// 44.5 for one N = 80 MMA (arithmetic floor 40).
template <int HD>
__device__ __forceinline__ void issue_pv_wide(uint32_t o_tmem, uint32_t p_tmem, uint32_t v_main, uint32_t atom_stride, int ksteps,
                                              bool accumulate) {
    constexpr uint32_t id = ptx::make_idesc_bf16(128, HD, 0, 1);
    const uint64_t dv = ptx::make_smem_desc(v_main, HD > 64 ? atom_stride : 0, 1024, ptx::LAYOUT_SW128);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < ksteps) ptx::mma_f16_ts_e(o_tmem, p_tmem + 8 * k, dv + 128 * k, id, (accumulate || k) ? 1u : 0u);     // 16 keys = 2048 B
    }
}

// 2^x for a pair of exponents on the FMA / ALU pipes instead of the MUFU (16 ex2 per clock and SM is what bounds both softmax
// loops): round-to-nearest split x = j + f with |f| <= 0.5 by the 1.5 * 2^23 trick, degree-3 minimax polynomial of 2^f (maximum
// relative error 7.5e-5, far below the bf16 rounding of P), the exponent added by integer arithmetic.  x is clamped at -120.
__device__ __forceinline__ void exp2_poly_pair(float a0, float a1, float& p0, float& p1) {
    const f32x2 x = f2_pack(fmaxf(a0, -120.f), fmaxf(a1, -120.f));
    const f32x2 t = f2_add(x, f2_pack(12582912.f, 12582912.f));
    const f32x2 j = f2_add(t, f2_pack(-12582912.f, -12582912.f));
    const f32x2 f = f2_fma(j, f2_pack(-1.f, -1.f), x);
    f32x2 q = f2_fma(f2_pack(0.05517163872718811f, 0.05517163872718811f), f, f2_pack(0.2426111251115799f, 0.2426111251115799f));
    q = f2_fma(q, f, f2_pack(0.6932609677314758f, 0.6932609677314758f));
    q = f2_fma(q, f, f2_pack(0.9999280571937561f, 0.9999280571937561f));
    float t0, t1, q0, q1;
    f2_unpack(t, t0, t1);
    f2_unpack(q, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// which of every 8 pairs of exponentials go to the FMA pipe (exp2_poly_pair) instead of the MUFU: K8 of 8, spread evenly
__host__ __device__ constexpr bool poly_pair(int pair, int K8) {
    return (((K8 == 2 ? 0x88 : K8 == 3 ? 0xA4 : K8 == 4 ? 0xAA : K8 == 5 ? 0xDA : K8 == 6 ? 0xEE : K8 >= 8 ? 0xFF : 0) >> (pair & 7)) & 1) != 0;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));   // FMNMX3: two comparisons per issue slot
    return d;
}
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
    float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int e = 0; e < 32; e += 8) {
        m0 = fmax3(m0, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
        m1 = fmax3(m1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        m2 = fmax3(m2, __uint_as_float(v[e + 4]), __uint_as_float(v[e + 5]));
        m3 = fmax3(m3, __uint_as_float(v[e + 6]), __uint_as_float(v[e + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// write HD normalised outputs of one query row
template <int HD>
__device__ __forceinline__ void store_row(bf16* dst, uint32_t o_tmem, float inv) {
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
        ptx::tmem_ld_x32(o_tmem + c, v);
        ptx::tmem_ld_wait_dep(v);
        if (dst) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
                u.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
                u.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
                u.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
                reinterpret_cast<uint4*>(dst + c)[j] = u;
            }
        }
    }
    if (HD > 64) {
        uint32_t w[16];
        ptx::tmem_ld_x16(o_tmem + 64, w);
        ptx::tmem_ld_wait_dep(w);
        if (dst) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(w[8 * j + 0]) * inv, __uint_as_float(w[8 * j + 1]) * inv);
                u.y = pack_bf16x2(__uint_as_float(w[8 * j + 2]) * inv, __uint_as_float(w[8 * j + 3]) * inv);
                u.z = pack_bf16x2(__uint_as_float(w[8 * j + 4]) * inv, __uint_as_float(w[8 * j + 5]) * inv);
                u.w = pack_bf16x2(__uint_as_float(w[8 * j + 6]) * inv, __uint_as_float(w[8 * j + 7]) * inv);
                reinterpret_cast<uint4*>(dst + 64)[j] = u;
            }
        }
    }
}

// One 128-query x 196-key window tile: S (fp32, TMEM columns [0,196) of s_tmem) -> P (bf16, TMEM columns [0,104)), returns
// the row sum.  bhm / bwl are the row's rel-pos terms (log2 units) per key row / key column; bhm is consumed (shifted by
// the reference maximum).  Two passes over TMEM: the exact row maximum, then exp2 / sum / pack with packed fp32x2 arithmetic.
// KO: knock-out diagnostics (SVB_BUILD_KO=1, tools/call14.sh; timing only, results wrong): 1 = no maximum pass, 2 = no exponentials,
// 8 = no P stores, 1024 / 2048 = the maximum / exponential pass WITHOUT its TMEM loads (arithmetic on stale registers)
template <int POLY, int KO = 0>
__device__ __forceinline__ float window_softmax_tile(uint32_t s_tmem, float (&bhm)[14], const float (&bwl)[14], float scale_log2) {
    uint32_t va[32], vb[32], vt[4];
    if constexpr ((KO & 3072) != 0) {               // diagnostics that skip loads: defined register contents
#pragma unroll
        for (int e = 0; e < 32; ++e) { va[e] = 0x3f800000u + e; vb[e] = 0x3f000000u + e; }
    }
    const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
    // ---- pass A: the row's EXACT maximum exponent.  Per key row kh the maximum of s * scale + w term (an FMA and half a
    // three-input maximum per element: the pairs (k, k+1) never straddle a key row, 14 is even), then + h term.  A reference that
    // was really attained cannot make the whole row underflow, whatever the spread of the bias terms (an upper bound built from
    // the largest bias terms could: P = exp2(x - bound) flushes to zero once the bound overshoots by ~126). ----
    float mk[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) mk[k] = -INFINITY;
#define SVB_WIN_A(V, CHUNK)                                                                              \
    _Pragma("unroll") for (int e = 0; e < 32; e += 2) {                                                  \
        const int k0 = 32 * (CHUNK) + e;                                                                 \
        const f32x2 x = f2_fma(f2_pack(__uint_as_float(V[e]), __uint_as_float(V[e + 1])), sc2, f2_pack(bwl[k0 % 14], bwl[(k0 + 1) % 14])); \
        float a0, a1;                                                                                    \
        f2_unpack(x, a0, a1);                                                                            \
        mk[k0 / 14] = fmax3(mk[k0 / 14], a0, a1);                                                        \
    }
    if constexpr (!(KO & 1)) {
    if constexpr (!(KO & 1024)) ptx::tmem_ld_x32(s_tmem, va);
    if constexpr (!(KO & 1024)) ptx::tmem_ld_wait_dep(va);
    if constexpr (!(KO & 1024)) ptx::tmem_ld_x32(s_tmem + 32, vb);
    SVB_WIN_A(va, 0)
    if constexpr (!(KO & 1024)) ptx::tmem_ld_wait_dep(vb);
    if constexpr (!(KO & 1024)) ptx::tmem_ld_x32(s_tmem + 64, va);
    SVB_WIN_A(vb, 1)
    if constexpr (!(KO & 1024)) ptx::tmem_ld_wait_dep(va);
    if constexpr (!(KO & 1024)) ptx::tmem_ld_x32(s_tmem + 96, vb);
    SVB_WIN_A(va, 2)
    if constexpr (!(KO & 1024)) ptx::tmem_ld_wait_dep(vb);
    if constexpr (!(KO & 1024)) ptx::tmem_ld_x32(s_tmem + 128, va);
    SVB_WIN_A(vb, 3)
    if constexpr (!(KO & 1024)) ptx::tmem_ld_wait_dep(va);
    if constexpr (!(KO & 1024)) ptx::tmem_ld_x32(s_tmem + 160, vb);
    SVB_WIN_A(va, 4)
    if constexpr (!(KO & 1024)) ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x4(s_tmem + 192, vt);
    SVB_WIN_A(vb, 5)
    ptx::tmem_ld_wait_dep(vt);
    } else {
        ptx::tmem_ld_x4(s_tmem + 192, vt);
        ptx::tmem_ld_wait_dep(vt);
        mk[0] = 0.f;
    }
#undef SVB_WIN_A
    mk[13] = fmaxf(mk[13], fmaxf(fmaxf(fmaf(__uint_as_float(vt[0]), scale_log2, bwl[192 % 14]), fmaf(__uint_as_float(vt[1]), scale_log2, bwl[193 % 14])),
                                 fmaxf(fmaf(__uint_as_float(vt[2]), scale_log2, bwl[194 % 14]), fmaf(__uint_as_float(vt[3]), scale_log2, bwl[195 % 14]))));
    float m_ref = mk[0] + bhm[0];
#pragma unroll
    for (int k = 1; k < 14; ++k) m_ref = fmaxf(m_ref, mk[k] + bhm[k]);
#pragma unroll
    for (int k = 0; k < 14; ++k) bhm[k] -= m_ref;
    // ---- pass B ----
    f32x2 l01 = f2_pack(0.f, 0.f);
#define SVB_WIN_B(V, CHUNK)                                                                              \
    {                                                                                                \
        uint32_t pk[16];                                                                             \
        _Pragma("unroll") for (int e = 0; e < 32; e += 2) {                                          \
            const int k0 = 32 * (CHUNK) + e, k1 = k0 + 1;                                            \
            const f32x2 x = f2_add(f2_fma(f2_pack(__uint_as_float(V[e]), __uint_as_float(V[e + 1])), sc2,               \
                                          f2_pack(bwl[k0 % 14], bwl[k1 % 14])), f2_pack(bhm[k0 / 14], bhm[k1 / 14]));   \
            float a0, a1;                                                                            \
            f2_unpack(x, a0, a1);                                                                    \
            float p0, p1;                                                                            \
            if constexpr (KO & 2) { p0 = a0; p1 = a1; }                                              \
            else if (poly_pair(e / 2, POLY)) exp2_poly_pair(a0, a1, p0, p1);   /* POLY of every 8 pairs on the FMA pipe */ \
            else { p0 = ptx::ex2_approx(a0); p1 = ptx::ex2_approx(a1); }                             \
            l01 = f2_add(l01, f2_pack(p0, p1));                                                      \
            pk[e / 2] = pack_bf16x2(p0, p1);                                                         \
        }                                                                                            \
        if constexpr (!(KO & 8)) ptx::tmem_st_x16(s_tmem + 16 * (CHUNK), pk);                        \
        else asm volatile("" :: "r"(pk[0] ^ pk[1] ^ pk[2] ^ pk[3] ^ pk[4] ^ pk[5] ^ pk[6] ^ pk[7] ^ pk[8] ^ pk[9] ^ pk[10] ^ pk[11] ^ pk[12] ^ pk[13] ^ pk[14] ^ pk[15])); \
    }
    if constexpr (!(KO & 2048)) ptx::tmem_ld_x32(s_tmem, va);
    if constexpr (!(KO & 2048)) ptx::tmem_ld_wait_dep(va);
    if constexpr (!(KO & 2048)) ptx::tmem_ld_x32(s_tmem + 32, vb);
    SVB_WIN_B(va, 0)
    if constexpr (!(KO & 2048)) ptx::tmem_ld_wait_dep(vb);
    if constexpr (!(KO & 2048)) ptx::tmem_ld_x32(s_tmem + 64, va);
    SVB_WIN_B(vb, 1)
    if constexpr (!(KO & 2048)) ptx::tmem_ld_wait_dep(va);
    if constexpr (!(KO & 2048)) ptx::tmem_ld_x32(s_tmem + 96, vb);
    SVB_WIN_B(va, 2)
    if constexpr (!(KO & 2048)) ptx::tmem_ld_wait_dep(vb);
    if constexpr (!(KO & 2048)) ptx::tmem_ld_x32(s_tmem + 128, va);
    SVB_WIN_B(vb, 3)
    if constexpr (!(KO & 2048)) ptx::tmem_ld_wait_dep(va);
    if constexpr (!(KO & 2048)) ptx::tmem_ld_x32(s_tmem + 160, vb);
    SVB_WIN_B(va, 4)
    if constexpr (!(KO & 2048)) ptx::tmem_ld_wait_dep(vb);
    SVB_WIN_B(vb, 5)
#undef SVB_WIN_B
    {
        // keys 192..195 (vt was loaded in pass A and is still live) + zero columns for keys 196..207
        uint32_t pk[8];
        const float p0 = ptx::ex2_approx(fmaf(__uint_as_float(vt[0]), scale_log2, bwl[192 % 14]) + bhm[192 / 14]);
        const float p1 = ptx::ex2_approx(fmaf(__uint_as_float(vt[1]), scale_log2, bwl[193 % 14]) + bhm[193 / 14]);
        const float p2 = ptx::ex2_approx(fmaf(__uint_as_float(vt[2]), scale_log2, bwl[194 % 14]) + bhm[194 / 14]);
        const float p3 = ptx::ex2_approx(fmaf(__uint_as_float(vt[3]), scale_log2, bwl[195 % 14]) + bhm[195 / 14]);
        l01 = f2_add(l01, f2_pack(p0 + p2, p1 + p3));
        pk[0] = pack_bf16x2(p0, p1);
        pk[1] = pack_bf16x2(p2, p3);
#pragma unroll
        for (int e = 2; e < 8; ++e) pk[e] = 0u;
        ptx::tmem_st_x8(s_tmem + 96, pk);
    }
    ptx::tmem_st_wait();
    float l0, l1;
    f2_unpack(l01, l0, l1);
    return l0 + l1;
}

// The same tile with ONE pass over tensor memory.  Measured (tools/tmem_rate.py, profiles/r02_attnw_analysis): tcgen05.ld delivers
// 56 bytes per clock PER SM however many warps issue it (one x32 load of 4 KB = 72 cycles), and the two-pass form above reads 532
// columns per query row (S twice, the rel-pos products, O) — 8700 cycles per (window, head) item: that, not the MUFU or the issue
// slots, is what the windowed kernel costs.  Here S is read once: the reference of the exponentials is the exact maximum of the row's
// FIRST 32 keys (a value that is attained, so the row cannot underflow as a whole); every later chunk's maximum is checked before its
// exponentials and, if it exceeds the reference by more than 2^64 (never on real activations; tests/test_gpu_ops.py provokes it), the
// reference moves there and the P columns already written are rescaled in place — P <= 2^64 always, so neither bf16 P, the fp32 row
// sum nor the fp32 O accumulator can overflow (torch.softmax has no limit either, image_encoder.py:246-252).
template <int POLY>
__device__ __forceinline__ float window_softmax_tile_1p(uint32_t s_tmem, float (&bhm)[14], const float (&bwl)[14], float scale_log2) {
    constexpr float LIMIT = 64.0f;
    uint32_t va[32], vb[32], vt[4];
    const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
    f32x2 l01 = f2_pack(0.f, 0.f);
    // maximum of chunk CHUNK's exponents (relative to the current reference, which bhm carries)
#define SVB_W1_MAX(V, CHUNK, MX)                                                                         \
    {                                                                                                    \
        float m0_ = -INFINITY, m1_ = -INFINITY;                                                          \
        _Pragma("unroll") for (int e = 0; e < 32; e += 4) {                                              \
            const int k0 = 32 * (CHUNK) + e;                                                             \
            const f32x2 x0 = f2_add(f2_fma(f2_pack(__uint_as_float(V[e]), __uint_as_float(V[e + 1])), sc2,              \
                                           f2_pack(bwl[k0 % 14], bwl[(k0 + 1) % 14])), f2_pack(bhm[k0 / 14], bhm[(k0 + 1) / 14])); \
            const f32x2 x1 = f2_add(f2_fma(f2_pack(__uint_as_float(V[e + 2]), __uint_as_float(V[e + 3])), sc2,          \
                                           f2_pack(bwl[(k0 + 2) % 14], bwl[(k0 + 3) % 14])), f2_pack(bhm[(k0 + 2) / 14], bhm[(k0 + 3) / 14])); \
            float a0, a1, a2, a3;                                                                        \
            f2_unpack(x0, a0, a1);                                                                       \
            f2_unpack(x1, a2, a3);                                                                       \
            m0_ = fmax3(m0_, a0, a1);                                                                    \
            m1_ = fmax3(m1_, a2, a3);                                                                    \
        }                                                                                                \
        MX = fmaxf(m0_, m1_);                                                                            \
    }
    // exp2 / sum / pack / store of chunk CHUNK
#define SVB_W1_EXP(V, CHUNK)                                                                             \
    {                                                                                                    \
        uint32_t pk[16];                                                                                 \
        _Pragma("unroll") for (int e = 0; e < 32; e += 2) {                                              \
            const int k0 = 32 * (CHUNK) + e, k1 = k0 + 1;                                                \
            const f32x2 x = f2_add(f2_fma(f2_pack(__uint_as_float(V[e]), __uint_as_float(V[e + 1])), sc2,               \
                                          f2_pack(bwl[k0 % 14], bwl[k1 % 14])), f2_pack(bhm[k0 / 14], bhm[k1 / 14]));   \
            float a0, a1;                                                                                \
            f2_unpack(x, a0, a1);                                                                        \
            float p0, p1;                                                                                \
            if (poly_pair(e / 2, POLY)) exp2_poly_pair(a0, a1, p0, p1);   /* POLY of every 8 pairs on the FMA pipe */ \
            else { p0 = ptx::ex2_approx(a0); p1 = ptx::ex2_approx(a1); }                                 \
            l01 = f2_add(l01, f2_pack(p0, p1));                                                          \
            pk[e / 2] = pack_bf16x2(p0, p1);                                                             \
        }                                                                                                \
        ptx::tmem_st_x16(s_tmem + 16 * (CHUNK), pk);                                                     \
    }
    // the reference moves up by the chunk maximum where that exceeds LIMIT: rescale what has been written (chunks 0 .. CHUNK-1)
#define SVB_W1_CHECK(CHUNK, MX)                                                                          \
    if (__any_sync(0xffffffffu, MX > LIMIT)) {                                                           \
        const float delta = MX > LIMIT ? MX : 0.f;                                                       \
        const float f = ptx::ex2_approx(-delta);                                                         \
        ptx::tmem_st_wait();                                                                             \
        _Pragma("unroll 1") for (int c = 0; c < (CHUNK); ++c) {                                          \
            uint32_t r[16];                                                                              \
            ptx::tmem_ld_x16(s_tmem + 16 * c, r);                                                        \
            ptx::tmem_ld_wait_dep(r);                                                                    \
            _Pragma("unroll") for (int e = 0; e < 16; ++e)                                               \
                r[e] = pack_bf16x2(__uint_as_float(r[e] << 16) * f, __uint_as_float(r[e] & 0xffff0000u) * f);           \
            ptx::tmem_st_x16(s_tmem + 16 * c, r);                                                        \
        }                                                                                                \
        l01 = f2_mul(l01, f2_pack(f, f));                                                                \
        _Pragma("unroll") for (int k = 0; k < 14; ++k) bhm[k] -= delta;                                  \
    }
    float mx;
    ptx::tmem_ld_x32(s_tmem, va);
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x32(s_tmem + 32, vb);
    SVB_W1_MAX(va, 0, mx)                                          // the reference: exact maximum of keys 0..31
#pragma unroll
    for (int k = 0; k < 14; ++k) bhm[k] -= mx;
    SVB_W1_EXP(va, 0)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x32(s_tmem + 64, va);
    SVB_W1_MAX(vb, 1, mx)
    SVB_W1_CHECK(1, mx)
    SVB_W1_EXP(vb, 1)
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x32(s_tmem + 96, vb);
    SVB_W1_MAX(va, 2, mx)
    SVB_W1_CHECK(2, mx)
    SVB_W1_EXP(va, 2)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x32(s_tmem + 128, va);
    SVB_W1_MAX(vb, 3, mx)
    SVB_W1_CHECK(3, mx)
    SVB_W1_EXP(vb, 3)
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x32(s_tmem + 160, vb);
    SVB_W1_MAX(va, 4, mx)
    SVB_W1_CHECK(4, mx)
    SVB_W1_EXP(va, 4)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x4(s_tmem + 192, vt);
    SVB_W1_MAX(vb, 5, mx)
    SVB_W1_CHECK(5, mx)
    SVB_W1_EXP(vb, 5)
    ptx::tmem_ld_wait_dep(vt);
    {
        // keys 192..195 + zero columns for keys 196..207
        float a0 = fmaf(__uint_as_float(vt[0]), scale_log2, bwl[192 % 14]) + bhm[192 / 14];
        float a1 = fmaf(__uint_as_float(vt[1]), scale_log2, bwl[193 % 14]) + bhm[193 / 14];
        float a2 = fmaf(__uint_as_float(vt[2]), scale_log2, bwl[194 % 14]) + bhm[194 / 14];
        float a3 = fmaf(__uint_as_float(vt[3]), scale_log2, bwl[195 % 14]) + bhm[195 / 14];
        mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        if (__any_sync(0xffffffffu, mx > LIMIT)) {
            const float delta = mx > LIMIT ? mx : 0.f;
            const float f = ptx::ex2_approx(-delta);
            ptx::tmem_st_wait();
#pragma unroll 1
            for (int c = 0; c < 6; ++c) {
                uint32_t r[16];
                ptx::tmem_ld_x16(s_tmem + 16 * c, r);
                ptx::tmem_ld_wait_dep(r);
#pragma unroll
                for (int e = 0; e < 16; ++e) r[e] = pack_bf16x2(__uint_as_float(r[e] << 16) * f, __uint_as_float(r[e] & 0xffff0000u) * f);
                ptx::tmem_st_x16(s_tmem + 16 * c, r);
            }
            l01 = f2_mul(l01, f2_pack(f, f));
            a0 -= delta; a1 -= delta; a2 -= delta; a3 -= delta;
        }
        uint32_t pk[8];
        const float p0 = ptx::ex2_approx(a0), p1 = ptx::ex2_approx(a1), p2 = ptx::ex2_approx(a2), p3 = ptx::ex2_approx(a3);
        l01 = f2_add(l01, f2_pack(p0 + p2, p1 + p3));
        pk[0] = pack_bf16x2(p0, p1);
        pk[1] = pack_bf16x2(p2, p3);
#pragma unroll
        for (int e = 2; e < 8; ++e) pk[e] = 0u;
        ptx::tmem_st_x8(s_tmem + 96, pk);
    }
#undef SVB_W1_MAX
#undef SVB_W1_EXP
#undef SVB_W1_CHECK
    ptx::tmem_st_wait();
    float l0, l1;
    f2_unpack(l01, l0, l1);
    return l0 + l1;
}

// r[j] <- r[j + sh] for a per-thread shift sh in [0, 13]: four conditional-move stages
__device__ __forceinline__ void barrel_shift27(float (&r)[27], int sh) {
#pragma unroll
    for (int bit = 1; bit <= 8; bit <<= 1) {
        const bool on = (sh & bit) != 0;
#pragma unroll
        for (int j = 0; j + bit < 27; ++j) r[j] = on ? r[j + bit] : r[j];
    }
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}


}  // namespace
}  // namespace svb
