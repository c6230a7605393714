// EXPERIMENT (built only with SVB_BUILD_EXPERIMENTAL=1, selected by SVB_ATTNW_IMPL=6) — correct (test_attention_tcgen05 green) and
// measured SLOWER than the kernels it was meant to beat (tools/call18.sh, profiles/r02_attnw_analysis/c18_*): 240 us per 16 ViT-H
// images (209 with every exponential on the MUFU) against 208 for the two-group kernel, 36.7 against 34.1-35.0 ms per 64 images
// inside the encoder step.  Doubling the softmax warps cannot help a kernel that is bound by the TMEM read port (56 bytes per clock
// and SM, tools/tmem_rate.py) — the measurement that led to the one-pass softmax of attention_win5.cu.
// Windowed attention (14 x 14 windows of the padded token grid) with FOUR softmax warps per scheduler (sm_100a).  Same arithmetic and
// the same shared-memory / tensor-memory layout as attn_window_persistent_kernel (attention_tc.cu): S = Q K^T and O = P V on tcgen05
// with fp32 accumulators in TMEM, decomposed rel-pos bias from two extra MMAs against the rel_pos tables skewed in registers,
// exact-maximum base-2 softmax in fp32 (image_encoder.py:239-304, 340-376).  What the measurements of profiles/r02_attnw_analysis
// left: the eight softmax warps of that kernel — two per scheduler — issue one instruction per ~3.8 cycles and carry 8100 of its
// 9070 cycles per item however the two chains are arranged, and the ring adds (TMA latency + that) / 2.  Here
//   * a query row's 196 keys are split over TWO threads (keys 0..111 = 8 key rows / keys 112..195 = 6 key rows + the 12 zero columns
//     of the padded contraction; warps w and w + 4 of a chain work on the same 32 rows): 16 softmax warps at 104 registers
//     (setmaxnreg), four per scheduler.  Each half writes its P over S columns only IT reads: keys 0..111 over the first 56 columns
//     of the S tile, keys 112..207 into 48 of the 96 tensor-memory columns beside the two S tiles (the PV MMAs of the last six K
//     steps take their A operand from there).  The two halves exchange the row maximum and the row sum through one shared-memory
//     word per row and a 64-thread named barrier; each normalises and stages half of the output columns;
//   * the CTA runs two INDEPENDENT chains (chain c owns stage c of the ring, S tile c of tensor memory, 8 softmax warps, an MMA
//     issuer and a load / store thread, and walks its own items, query tile 0 then query tile 1), and the stage is refilled piece
//     by piece as the item lets go of it (Q0 after the store of O0, K after the last S MMA, V after the last PV, Q1 after the
//     store of O1), so the TMA latency of the next item hides under the second tile;
//   * the warps of tile 1 whose 32 rows are all padding (rows 96..127 of the 70-row tile) keep the barrier protocol and skip the
//     arithmetic; Q1 is not loaded for windows whose tile 1 lies outside the token grid.
#include "attention_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace svb {
namespace {

// shared-memory / tensor-map layout of the production windowed kernel (attention_tc.cu)
template <int HD> struct WPCfg {
    static constexpr int TAIL = HD - 64;
    static constexpr int Q_MAIN = 128 * 128, Q_TAIL = TAIL ? 128 * 32 : 0;   // one query tile (126 / 70 rows used)
    static constexpr int QT = Q_MAIN + Q_TAIL;
    static constexpr int K_MAIN = 208 * 128, K_TAIL = TAIL ? 7168 : 0;       // 196 keys (+12 pad rows)
    static constexpr int KT = K_MAIN + K_TAIL;
    static constexpr int OFF_K = 2 * QT, OFF_V = OFF_K + KT;
    static constexpr int STAGE = 2 * QT + 2 * KT;
    static constexpr int R_MAIN = 64 * 128, R_TAIL = TAIL ? 64 * 32 : 0;
    static constexpr int OFF_R = 2 * STAGE;
    static constexpr int OFF_BAR = OFF_R + R_MAIN + R_TAIL;
    static constexpr int OFF_XCH = OFF_BAR + 256;                            // [2 chains][2 halves][128 rows] fp32: row maximum, then row sum
    static constexpr int SMEM = OFF_XCH + 2048 + 1024;
    static_assert(SMEM <= 232448, "shared memory budget");
    static constexpr int ROWB = 128 + (TAIL ? 32 : 0);
    static constexpr int V_TX = 196 * ROWB;
    static constexpr int R_TX = 64 * ROWB;
    static constexpr int V_ATOM = 208 * 32;
    static constexpr int TM_COLS = 512;                                       // S_c at 208 c; O_c inside S_c at +112
};

struct WinPMaps {
    CUtensorMap q0, q1, kv, r;           // loads: boxes (64,14,9,1) / (64,14,5,1) / (64,14,14,1) of the padded qkv; (64,64) of the table
    CUtensorMap q0t, q1t, kvt, rt;       // their 16-column tails (32B swizzle)
    CUtensorMap o0, o1, o0t, o1t;        // stores: boxes (64|16,14,9|5,1) of out viewed as [B,gh,gw,D]
};

// u[j] <- u[j + sh] for j < NOUT, sh in [0, 2^STAGES): conditional-move stages from the high bit down, each only as wide as the later
// stages still need (entries past the 27 products are never selected: sh + j <= 26 by construction)
template <int NOUT, int STAGES>
__device__ __forceinline__ void barrel_pick(float (&u)[27], int sh) {
#pragma unroll
    for (int s = STAGES - 1; s >= 0; --s) {
        const int bit = 1 << s;
        const bool on = (sh & bit) != 0;
#pragma unroll
        for (int j = 0; j < NOUT + bit - 1 && j < 27; ++j) u[j] = on ? (j + bit < 27 ? u[j + bit] : 0.f) : u[j];
    }
}

// One half of a query row of a 128-query x 196-key window tile.  HALF 0: keys 0..111 (8 key rows), S columns [0,112) of the tile at
// s_half, P (bf16 pairs) into 56 columns at p_half.  HALF 1: keys 112..195 (6 key rows) + zeros for keys 196..207, S columns
// [112,208) at s_half, P into 48 columns at p_half.  bh / bw: the row's rel-pos terms of this half's key rows / of the 14 key columns
// (log2 units; bh is consumed).  Two passes over TMEM: the exact maximum (exchanged with the partner thread through xch_mine /
// xch_other and the pair's named barrier), then exp2 / sum / pack.  Returns the half's row sum.
template <int POLY, int HALF>
__device__ __forceinline__ float window_softmax_half(uint32_t s_half, uint32_t p_half, float (&bh)[8], const float (&bw)[14],
                                                     float scale_log2, float* xch_mine, float* xch_other, int pair_bar) {
    constexpr int NCH = HALF ? 6 : 7, NROWS = HALF ? 6 : 8, REAL = HALF ? 84 : 112;     // 16-column chunks, key rows, real keys
    uint32_t va[16], vb[16];
    const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
    float mk[NROWS];
#pragma unroll
    for (int k = 0; k < NROWS; ++k) mk[k] = -INFINITY;
#define SVB_H_A(V, CHUNK)                                                                                \
    if ((CHUNK) < NCH) {                                                                                 \
        _Pragma("unroll") for (int e = 0; e < 16; e += 2) {                                              \
            const int k0 = 16 * (CHUNK) + e;                                                             \
            if (k0 < REAL) {                                                                             \
                const f32x2 x = f2_fma(f2_pack(__uint_as_float(V[e]), __uint_as_float(V[e + 1])), sc2, f2_pack(bw[k0 % 14], bw[(k0 + 1) % 14])); \
                float a0, a1;                                                                            \
                f2_unpack(x, a0, a1);                                                                    \
                mk[k0 / 14] = fmax3(mk[k0 / 14], a0, a1);                                                \
            }                                                                                            \
        }                                                                                                \
    }
    ptx::tmem_ld_x16(s_half, va);
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x16(s_half + 16, vb);
    SVB_H_A(va, 0)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x16(s_half + 32, va);
    SVB_H_A(vb, 1)
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x16(s_half + 48, vb);
    SVB_H_A(va, 2)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x16(s_half + 64, va);
    SVB_H_A(vb, 3)
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x16(s_half + 80, vb);
    SVB_H_A(va, 4)
    ptx::tmem_ld_wait_dep(vb);
    if (NCH > 6) ptx::tmem_ld_x16(s_half + 96, va);
    SVB_H_A(vb, 5)
    if (NCH > 6) {
        ptx::tmem_ld_wait_dep(va);
        SVB_H_A(va, 6)
    }
#undef SVB_H_A
    float m_half = mk[0] + bh[0];
#pragma unroll
    for (int k = 1; k < NROWS; ++k) m_half = fmaxf(m_half, mk[k] + bh[k]);
    *xch_mine = m_half;
    ptx::named_bar_sync(pair_bar, 64);
    const float m_ref = fmaxf(m_half, *xch_other);
#pragma unroll
    for (int k = 0; k < NROWS; ++k) bh[k] -= m_ref;
    // ---- pass B ----
    f32x2 l01 = f2_pack(0.f, 0.f);
#define SVB_H_B(V, CHUNK)                                                                                \
    if ((CHUNK) < NCH) {                                                                                 \
        uint32_t pk[8];                                                                                  \
        _Pragma("unroll") for (int e = 0; e < 16; e += 2) {                                              \
            const int k0 = 16 * (CHUNK) + e, k1 = k0 + 1;                                                \
            if (k0 < REAL) {                                                                             \
                const f32x2 x = f2_add(f2_fma(f2_pack(__uint_as_float(V[e]), __uint_as_float(V[e + 1])), sc2,           \
                                              f2_pack(bw[k0 % 14], bw[k1 % 14])), f2_pack(bh[k0 / 14 < NROWS ? k0 / 14 : 0], bh[k1 / 14 < NROWS ? k1 / 14 : 0])); \
                float a0, a1;                                                                            \
                f2_unpack(x, a0, a1);                                                                    \
                float p0, p1;                                                                            \
                if (poly_pair(e / 2 + 8 * ((CHUNK) & 1), POLY)) exp2_poly_pair(a0, a1, p0, p1);          \
                else { p0 = ptx::ex2_approx(a0); p1 = ptx::ex2_approx(a1); }                             \
                l01 = f2_add(l01, f2_pack(p0, p1));                                                      \
                pk[e / 2] = pack_bf16x2(p0, p1);                                                         \
            } else {                                                                                     \
                pk[e / 2] = 0u;                   /* keys 196..207 of the padded contraction */          \
            }                                                                                            \
        }                                                                                                \
        ptx::tmem_st_x8(p_half + 8 * (CHUNK), pk);                                                       \
    }
    ptx::tmem_ld_x16(s_half, va);
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x16(s_half + 16, vb);
    SVB_H_B(va, 0)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x16(s_half + 32, va);
    SVB_H_B(vb, 1)
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x16(s_half + 48, vb);
    SVB_H_B(va, 2)
    ptx::tmem_ld_wait_dep(vb);
    ptx::tmem_ld_x16(s_half + 64, va);
    SVB_H_B(vb, 3)
    ptx::tmem_ld_wait_dep(va);
    ptx::tmem_ld_x16(s_half + 80, vb);
    SVB_H_B(va, 4)
    ptx::tmem_ld_wait_dep(vb);
    if (NCH > 6) ptx::tmem_ld_x16(s_half + 96, va);
    SVB_H_B(vb, 5)
    if (NCH > 6) {
        ptx::tmem_ld_wait_dep(va);
        SVB_H_B(va, 6)
    }
#undef SVB_H_B
    ptx::tmem_st_wait();
    float l0, l1;
    f2_unpack(l01, l0, l1);
    return l0 + l1;
}

template <int HD> struct W6Cfg : WPCfg<HD> {
    using P = WPCfg<HD>;
    // barriers: RFULL, then per chain c at 1 + 10 c
    static constexpr int B_RFULL = 0, B_QF0 = 0, B_QF1 = 1, B_KF = 2, B_VF = 3, B_BIAS = 4, B_BREAD = 5, B_SFULL = 6, B_PFULL = 7, B_PVDONE = 8,
                         B_OSTAGED = 9, B_PER_CHAIN = 10, B_COUNT = 21;
    static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");
    static constexpr int Q0_TX = 126 * P::ROWB, Q1_TX = 70 * P::ROWB, K_TX = 196 * P::ROWB;
    static constexpr int TM_SPARE = 416;         // 48 columns per chain beside the two S tiles: P of keys 112..207
};

template <int HD, int POLY>
__global__ void __launch_bounds__(640, 1)
attn_window_split_kernel(const __grid_constant__ WinPMaps maps, int D, int g, int nwy, int nwx, int heads, int num_items, float scale_log2,
                          int l2_ahead) {
    // g = token-grid HEIGHT (the last window row's padding decides whether query tile 1 exists); nwy x nwx windows per image
    using C = W6Cfg<HD>;
    constexpr int WS = 14;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // provably warp-uniform

    if (warp == 16 && lane == 0) {
        ptx::prefetch_tmap(&maps.kv);
        ptx::prefetch_tmap(&maps.q0);
        ptx::prefetch_tmap(&maps.q1);
        ptx::prefetch_tmap(&maps.r);
        ptx::prefetch_tmap(&maps.o0);
        ptx::prefetch_tmap(&maps.o1);
        ptx::mbar_init(&bars[C::B_RFULL], 1);
        for (int c = 0; c < 2; ++c)
            for (int s = 0; s < C::B_PER_CHAIN; ++s)
                ptx::mbar_init(&bars[1 + C::B_PER_CHAIN * c + s], (s == C::B_BREAD || s == C::B_PFULL || s == C::B_OSTAGED) ? 256 : 1);
        ptx::fence_barrier_init();
    }
    if (warp == 18) ptx::tmem_alloc(tmem_slot, C::TM_COLS);
    // keys 196..207 of the PV contraction multiply P = 0: their V rows (never written by TMA) must be finite in both stages
    for (int st = 0; st < 2; ++st) {
        uint8_t* v = sm + st * C::STAGE + C::OFF_V;
        if (HD > 64) {
            for (int i = threadIdx.x; i < 5 * 24; i += blockDim.x)
                *reinterpret_cast<uint4*>(v + (i / 24) * C::V_ATOM + 196 * 32 + (i % 24) * 16) = make_uint4(0, 0, 0, 0);
        } else {
            for (int i = threadIdx.x; i < 96; i += blockDim.x) *reinterpret_cast<uint4*>(v + 196 * 128 + i * 16) = make_uint4(0, 0, 0, 0);
        }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto decode = [&](int item, int& b, int& wy, int& wx, int& head) {
        head = item % heads;
        const int bw = item / heads;
        const int win = bw % (nwy * nwx);
        b = bw / (nwy * nwx);
        wy = win / nwx;
        wx = win % nwx;
    };
    // query tile 1 (window rows 9..13) is entirely padding in the last window row of a grid whose height is 8 (mod 14)
    auto two_tiles = [&](int item) {
        const int wy = (item / heads) % (nwy * nwx) / nwx;
        return wy * WS + 9 < g;
    };
    const int stride = 2 * gridDim.x;

    if (warp >= 16) {
    ptx::setmaxnreg_dec<56>();                                     // ONE instruction for the whole auxiliary warp group
    if (warp == 16 || warp == 17) {
        // ===================== load / store thread of chain c =====================
        const int c = warp - 16;
        uint64_t* cb = bars + 1 + C::B_PER_CHAIN * c;
        uint8_t* q0 = sm + c * C::STAGE;
        uint8_t* q1 = q0 + C::QT;
        uint8_t* kk = q0 + C::OFF_K;
        uint8_t* vv = q0 + C::OFF_V;
        if (lane == 0) {
            const int dbg = l2_ahead >> 8;                           // DIAGNOSTIC (timing only): 1 no V loads, 2 no K, 4 no Q, 8 no stores
            l2_ahead &= 0xff;
            if (c == 0) {
                ptx::mbar_expect_tx(&bars[C::B_RFULL], C::R_TX);
                ptx::tma_load_2d(sm + C::OFF_R, &maps.r, &bars[C::B_RFULL], 0, 0);
                if (HD > 64) ptx::tma_load_2d(sm + C::OFF_R + C::R_MAIN, &maps.rt, &bars[C::B_RFULL], 64, 0);
            }
            auto load_q = [&](int item, int tile) {
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                uint8_t* dst = tile ? q1 : q0;
                uint64_t* bar = &cb[tile ? C::B_QF1 : C::B_QF0];
                if (dbg & 4) { ptx::mbar_arrive(bar); return; }
                ptx::mbar_expect_tx(bar, tile ? C::Q1_TX : C::Q0_TX);
                ptx::tma_load_4d(dst, tile ? &maps.q1 : &maps.q0, bar, head * HD, wx * WS, wy * WS + 9 * tile, b);
                if (HD > 64) ptx::tma_load_4d(dst + C::Q_MAIN, tile ? &maps.q1t : &maps.q0t, bar, head * HD + 64, wx * WS, wy * WS + 9 * tile, b);
            };
            auto load_k = [&](int item) {
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                if (dbg & 2) { ptx::mbar_arrive(&cb[C::B_KF]); return; }
                ptx::mbar_expect_tx(&cb[C::B_KF], C::K_TX);
                ptx::tma_load_4d(kk, &maps.kv, &cb[C::B_KF], D + head * HD, wx * WS, wy * WS, b);
                if (HD > 64) ptx::tma_load_4d(kk + C::K_MAIN, &maps.kvt, &cb[C::B_KF], D + head * HD + 64, wx * WS, wy * WS, b);
            };
            auto load_v = [&](int item) {
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                const int cv = 2 * D + head * HD;
                if (dbg & 1) { ptx::mbar_arrive(&cb[C::B_VF]); return; }
                ptx::mbar_expect_tx(&cb[C::B_VF], C::V_TX);
                if (HD > 64) {
                    for (int a = 0; a < 5; ++a) ptx::tma_load_4d(vv + a * C::V_ATOM, &maps.kvt, &cb[C::B_VF], cv + 16 * a, wx * WS, wy * WS, b);
                } else {
                    ptx::tma_load_4d(vv, &maps.kv, &cb[C::B_VF], cv, wx * WS, wy * WS, b);
                }
            };
            auto prefetch_l2 = [&](int item) {                       // HBM -> L2 for an item further ahead (its loads then hit L2)
                if (item >= num_items) return;
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                for (int part = 0; part < 3; ++part) {
                    const int cc = part * D + head * HD;
                    ptx::tma_prefetch_l2_4d(&maps.kv, cc, wx * WS, wy * WS, b);
                    if (HD > 64) ptx::tma_prefetch_l2_4d(&maps.kvt, cc + 64, wx * WS, wy * WS, b);
                }
            };
            auto store_o = [&](int item, int tile) {
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                const uint8_t* ob = tile ? q1 : q0;
                if (dbg & 8) return;
                const int x0 = wx * WS, y0 = wy * WS + 9 * tile;
                tma_store_4d(tile ? &maps.o1 : &maps.o0, ob, head * HD, x0, y0, b);       // rows / columns past the grid are clipped by the TMA
                if (HD > 64) tma_store_4d(tile ? &maps.o1t : &maps.o0t, ob + C::Q_MAIN, head * HD + 64, x0, y0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");            // the store has finished reading the Q buffer
            };
            int item = 2 * blockIdx.x + c;
            if (item < num_items) {
                load_q(item, 0);
                load_k(item);
                if (two_tiles(item)) load_q(item, 1);
                load_v(item);
                for (int a = 1; a <= l2_ahead; ++a) prefetch_l2(item + a * stride);
            }
            uint32_t nt = 0;                                         // tiles of this chain so far (phases of SFULL / PVDONE / OSTAGED)
            for (; item < num_items; item += stride) {
                const int nitem = item + stride;
                const bool next = nitem < num_items, t1 = two_tiles(item);
                ptx::mbar_wait(&cb[C::B_OSTAGED], nt & 1);           // O0 staged in the Q0 buffer (S0, PV0 retired)
                store_o(item, 0);
                if (next) load_q(nitem, 0);
                if (t1) {
                    ptx::mbar_wait(&cb[C::B_SFULL], (nt + 1) & 1);   // S of tile 1 retired: K is free
                    if (next) load_k(nitem);
                    ptx::mbar_wait(&cb[C::B_PVDONE], (nt + 1) & 1);  // PV of tile 1 retired: V is free
                    if (next) load_v(nitem);
                    ptx::mbar_wait(&cb[C::B_OSTAGED], (nt + 1) & 1);
                    store_o(item, 1);
                    if (next && two_tiles(nitem)) load_q(nitem, 1);
                    nt += 2;
                } else {
                    if (next) {
                        load_k(nitem);
                        if (two_tiles(nitem)) load_q(nitem, 1);
                        load_v(nitem);
                    }
                    nt += 1;
                }
                if (next) prefetch_l2(nitem + l2_ahead * stride);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");                     // stores complete before the CTA exits
        }
    } else {
        // ===================== MMA issuer of chain c (all 32 lanes run the loop, one elected lane issues) =====================
        constexpr uint32_t id_r = ptx::make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t id_s = ptx::make_idesc_bf16(128, 208, 0, 0);
        const int c = warp - 18;
        uint64_t* cb = bars + 1 + C::B_PER_CHAIN * c;
        const uint32_t s_tm = tmem + 208 * c;
        const uint32_t k = base + c * C::STAGE + C::OFF_K, v = base + c * C::STAGE + C::OFF_V;
        ptx::mbar_wait(&bars[C::B_RFULL], 0);
        uint32_t nt = 0, ni = 0, n1 = 0;                             // tiles, items, tile-1 loads of this chain so far
#pragma unroll 1
        for (int item = 2 * blockIdx.x + c; item < num_items; item += stride, ++ni) {
            const int ntiles = two_tiles(item) ? 2 : 1;
#pragma unroll 1
            for (int tile = 0; tile < ntiles; ++tile, ++nt) {
                const uint32_t q = base + c * C::STAGE + tile * C::QT;
                ptx::mbar_wait(&cb[tile ? C::B_QF1 : C::B_QF0], (tile ? n1 : ni) & 1);
                ptx::tc_fence_after();
                // columns [0,64) of S: the previous tile's P there was consumed by its PV (same issuer, in-order tensor pipe); its O
                // (columns 112..191) is only overwritten by the S MMA below, issued after the group has loaded it
                issue_qk<HD>(s_tm, q, q + C::Q_MAIN, base + C::OFF_R, base + C::OFF_R + C::R_MAIN, id_r);
                ptx::mma_commit_e(&cb[C::B_BIAS]);
                ptx::mbar_wait(&cb[C::B_BREAD], nt & 1);             // rel-pos products consumed and the previous O loaded
                if (tile == 0) ptx::mbar_wait(&cb[C::B_KF], ni & 1);
                ptx::tc_fence_after();
                issue_qk<HD>(s_tm, q, q + C::Q_MAIN, k, k + C::K_MAIN, id_s);
                ptx::mma_commit_e(&cb[C::B_SFULL]);
                if (tile == 0) ptx::mbar_wait(&cb[C::B_VF], ni & 1);
                ptx::mbar_wait(&cb[C::B_PFULL], nt & 1);             // P is in TMEM
                ptx::tc_fence_after();
                const uint32_t p_hi = tmem + C::TM_SPARE + 48 * c;  // P of keys 112..207 (the second half's columns)
                if (HD > 64) {
                    constexpr uint32_t id_pv = ptx::make_idesc_bf16(128, HD, 0, 1);
                    const uint64_t dv = ptx::make_smem_desc(v, C::V_ATOM, 256, ptx::LAYOUT_SW32);
#pragma unroll
                    for (int kq = 0; kq < 13; ++kq)                  // 16 keys = 512 B inside an atom
                        ptx::mma_f16_ts_e(s_tm + 112, kq < 7 ? s_tm + 8 * kq : p_hi + 8 * (kq - 7), dv + 32 * kq, id_pv, kq ? 1u : 0u);
                } else {
                    issue_pv<HD>(s_tm + 112, s_tm, v, v + C::K_MAIN, 7, false);
                    issue_pv<HD>(s_tm + 112, p_hi, v + 7 * 2048, v + C::K_MAIN + 7 * 512, 6, true);
                }
                ptx::mma_commit_e(&cb[C::B_PVDONE]);
            }
            if (ntiles == 2) ++n1;
        }
    }
    } else {
        // ===================== softmax warps of chain c: one query row per PAIR of threads =====================
        ptx::setmaxnreg_inc<104>();
        const int c = warp >> 3, half = (warp >> 2) & 1, q4 = warp & 3;
        uint64_t* cb = bars + 1 + C::B_PER_CHAIN * c;
        const int t = q4 * 32 + lane;                              // query row inside the tile
        const int xi = t % WS, yi0 = t / WS;                       // window coordinates in tile 0 (tile 1: row + 9; rows past the tile are discarded)
        const uint32_t lane_off = static_cast<uint32_t>(q4 * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + 208 * c;         // S row (columns 0..207), P over its first 104, O at +112
        // this half's S columns; its P: keys 0..111 over the first 56 S columns, keys 112..207 in the chain's 48 spare columns
        const uint32_t s_half = s_tmem + 112 * half, p_half = half ? tmem + lane_off + C::TM_SPARE + 48 * c : s_tmem;
        const uint32_t o_half = s_tmem + 112 + (HD / 2) * half;
        float* xch_mine = reinterpret_cast<float*>(sm + C::OFF_XCH) + (c * 2 + half) * 128 + t;
        float* xch_other = reinterpret_cast<float*>(sm + C::OFF_XCH) + (c * 2 + (half ^ 1)) * 128 + t;
        const int pair_bar = 1 + 4 * c + q4;                       // the two warps that share these 32 rows

        // rel-pos products of the tile whose bias MMA is the `nt`-th of this chain -> this half's 8 / 6 h terms and the 14 w terms
        float bh[8], bw[14];
        auto read_bias = [&](uint32_t nt, int tile) {
            ptx::mbar_wait(&cb[C::B_BIAS], nt & 1);
            ptx::tc_fence_after();
            uint32_t v[32];
            float u[27];
            const int yi = yi0 + 9 * tile;
            // the term of key row kh is products[yi + 13 - kh]: key rows 0..7 -> bh[kl] = products[yi + 6 + (7 - kl)], key rows 8..13
            // -> bh[kl] = products[yi + (5 - kl)]
            ptx::tmem_ld_x32(s_tmem, v);
            ptx::tmem_ld_wait_dep(v);
#pragma unroll
            for (int j = 0; j < 27; ++j) u[j] = __uint_as_float(v[j]) * LOG2E;
            const int yc = yi < 13 ? yi : 13;
            if (half) {
                barrel_pick<6, 4>(u, yc);
#pragma unroll
                for (int kl = 0; kl < 6; ++kl) bh[kl] = u[5 - kl];
                bh[6] = bh[7] = 0.f;
            } else {
                barrel_pick<8, 5>(u, yc + 6);
#pragma unroll
                for (int kl = 0; kl < 8; ++kl) bh[kl] = u[7 - kl];
            }
            ptx::tmem_ld_x32(s_tmem + 32, v);
            ptx::tmem_ld_wait_dep(v);
#pragma unroll
            for (int j = 0; j < 27; ++j) u[j] = __uint_as_float(v[j]) * LOG2E;
            barrel_pick<14, 4>(u, xi);
#pragma unroll
            for (int kk = 0; kk < 14; ++kk) bw[kk] = u[13 - kk];
            ptx::tc_fence_before();
            ptx::mbar_arrive(&cb[C::B_BREAD]);
        };
        // the same hand-shake for a warp whose rows are all padding in the coming tile
        auto skip_bias = [&](uint32_t nt) {
            ptx::mbar_wait(&cb[C::B_BIAS], nt & 1);
            ptx::mbar_arrive(&cb[C::B_BREAD]);
        };
        int item = 2 * blockIdx.x + c;
        int tile = 0;
        uint32_t nt = 0;
        if (item < num_items) read_bias(0, 0);
        while (item < num_items) {
            // the tile after this one
            int nitem = item, ntile = 1;
            if (tile == 1 || !two_tiles(item)) { nitem = item + stride; ntile = 0; }
            // next tile's rel-pos terms are read BEFORE this tile's output is stored (its S then runs during the store) — only when
            // the next tile's Q sits in the OTHER Q buffer: a one-tile item is followed by tile 0 of the next item in THIS buffer,
            // whose reload waits for the store (waiting for its bias first would deadlock)
            const bool more = nitem < num_items, early = more && (ntile != tile);
            if (q4 == 3 && tile == 1) {
                // rows 96..127 of the 70-row tile 1 are all padding: keep the barrier protocol, skip the arithmetic (the rows of P
                // and O these warps would have written are never stored)
                ptx::mbar_wait(&cb[C::B_SFULL], nt & 1);
                ptx::mbar_arrive(&cb[C::B_PFULL]);
                ptx::mbar_wait(&cb[C::B_PVDONE], nt & 1);
                ptx::mbar_arrive(&cb[C::B_OSTAGED]);
                if (more) read_bias(nt + 1, ntile);                // the tile after a tile 1 is a tile 0
            } else {
                // ---- softmax over this half's keys ----
                ptx::mbar_wait(&cb[C::B_SFULL], nt & 1);
                ptx::tc_fence_after();
                const float lmine = half ? window_softmax_half<POLY, 1>(s_half, p_half, bh, bw, scale_log2, xch_mine, xch_other, pair_bar)
                                         : window_softmax_half<POLY, 0>(s_half, p_half, bh, bw, scale_log2, xch_mine, xch_other, pair_bar);
                ptx::tc_fence_before();
                ptx::mbar_arrive(&cb[C::B_PFULL]);
                *xch_other = lmine;                                // the slot only the partner reads (its own maximum, consumed before the barrier inside)
                // ---- this half's columns of O: TMEM -> normalised bf16 in registers ----
                ptx::mbar_wait(&cb[C::B_PVDONE], nt & 1);
                ptx::tc_fence_after();
                ptx::named_bar_sync(pair_bar, 64);
                const float inv = 1.0f / (lmine + *xch_mine);
                uint32_t o[HD / 4];
                {
                    uint32_t v[32];
                    ptx::tmem_ld_x32(o_half, v);
                    ptx::tmem_ld_wait_dep(v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                    if (HD > 64) {
                        uint32_t w[8];
                        ptx::tmem_ld_x8(o_half + 32, w);
                        ptx::tmem_ld_wait_dep(w);
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[16 + j] = pack_bf16x2(__uint_as_float(w[2 * j]) * inv, __uint_as_float(w[2 * j + 1]) * inv);
                    }
                }
                const bool next_dead = q4 == 3 && ntile == 1;
                if (early) { if (next_dead) skip_bias(nt + 1); else read_bias(nt + 1, ntile); }
                // ---- O -> the dead Q buffer of this tile in the TMA layout -> one tensor store per tile ----
                {
                    uint8_t* ob = sm + c * C::STAGE + tile * C::QT;
#pragma unroll
                    for (int j = 0; j < HD / 16; ++j) {            // this half's 16-byte pieces: global piece g of the row
                        const int g8 = (HD / 16) * half + j;
                        const uint4 val = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                        if (g8 < 8) *reinterpret_cast<uint4*>(ob + t * 128 + ((g8 ^ (t & 7)) << 4)) = val;          // 128B swizzle
                        else *reinterpret_cast<uint4*>(ob + C::Q_MAIN + t * 32 + (((g8 - 8) ^ ((t >> 2) & 1)) << 4)) = val;   // 32B swizzle (tail)
                    }
                    ptx::fence_proxy_async_smem();                 // generic writes -> visible to the TMA (async proxy) read
                    ptx::mbar_arrive(&cb[C::B_OSTAGED]);           // the chain's load / store thread issues the tensor store and refills the buffer
                }
                if (!early && more) read_bias(nt + 1, ntile);      // (a tile 0 follows: never a padding-only warp)
            }
            item = nitem;
            tile = ntile;
            ++nt;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 18) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}

}  // namespace

// same contract as attention_tc() for ws = 14
int attention_window6(const AttnTcParams& p, cudaStream_t stream) {
    SVB_REQUIRE(p.hd == 64 || p.hd == 80, "attention_window6: head_dim %d", p.hd);
    auto run = [&](auto hd_tag) -> int {
        constexpr int HD = decltype(hd_tag)::value;
        using C = WPCfg<HD>;
        const int D = p.heads * p.hd;
        const int gh = p.grid_h ? p.grid_h : p.grid, gw = p.grid_w ? p.grid_w : p.grid;
        const int nwy = (gh + 13) / 14, nwx = (gw + 13) / 14, gph = nwy * 14, gpw = nwx * 14;
        WinPMaps wm;
        int rc;
        const uint64_t dims[4] = {(uint64_t)3 * D, (uint64_t)gpw, (uint64_t)gph, (uint64_t)p.batch};
        const uint64_t str[3] = {(uint64_t)3 * D * 2, (uint64_t)gpw * 3 * D * 2, (uint64_t)gph * gpw * 3 * D * 2};
        const uint32_t q0[4] = {64, 14, 9, 1}, q1[4] = {64, 14, 5, 1}, kv[4] = {64, 14, 14, 1};
        const uint32_t q0t[4] = {16, 14, 9, 1}, q1t[4] = {16, 14, 5, 1}, kvt[4] = {16, 14, 14, 1};
        if ((rc = encode_tmap_nd_bf16(&wm.q0, p.qkv, 4, dims, str, q0, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.q1, p.qkv, 4, dims, str, q1, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.kv, p.qkv, 4, dims, str, kv, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.q0t, p.qkv, 4, dims, str, q0t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.q1t, p.qkv, 4, dims, str, q1t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.kvt, p.qkv, 4, dims, str, kvt, 32))) return rc;
        const uint64_t rd[2] = {(uint64_t)HD, 64};
        const uint64_t rs[1] = {(uint64_t)HD * 2};
        const uint32_t rm[2] = {64, 64}, rt[2] = {16, 64};
        if ((rc = encode_tmap_nd_bf16(&wm.r, p.rel_pack, 2, rd, rs, rm, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.rt, p.rel_pack, 2, rd, rs, rt, 32))) return rc;
        const uint64_t od[4] = {(uint64_t)D, (uint64_t)gw, (uint64_t)gh, (uint64_t)p.batch};
        const uint64_t os[3] = {(uint64_t)D * 2, (uint64_t)gw * D * 2, (uint64_t)gh * gw * D * 2};
        if ((rc = encode_tmap_nd_bf16(&wm.o0, p.out, 4, od, os, q0, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o1, p.out, 4, od, os, q1, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o0t, p.out, 4, od, os, q0t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o1t, p.out, 4, od, os, q1t, 32))) return rc;
        const float scale_log2 = LOG2E / sqrtf((float)HD);
        const int items = p.batch * nwy * nwx * p.heads;
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = std::min(sms, (items + 1) / 2);
        static const int l2_ahead = [] { const char* e = getenv("SVB_ATTNW_L2AHEAD"); return e ? atoi(e) : 1; }();
        static const int poly = [] { const char* e = getenv("SVB_ATTNW_POLY"); return e ? atoi(e) : 2; }();
        auto launch = [&](auto kern) -> int {
            SVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
            kern<<<grid, 640, C::SMEM, stream>>>(wm, D, gh, nwy, nwx, p.heads, items, scale_log2, l2_ahead);
            return 0;
        };
        rc = poly == 0 ? launch(attn_window_split_kernel<HD, 0>) : poly == 4 ? launch(attn_window_split_kernel<HD, 4>) : launch(attn_window_split_kernel<HD, 2>);
        if (rc) return rc;
        SVB_CHECK_CUDA(cudaGetLastError());
        return 0;
    };
    return p.hd == 64 ? run(std::integral_constant<int, 64>{}) : run(std::integral_constant<int, 80>{});
}

}  // namespace svb
