// EXPERIMENT (built only with SVB_BUILD_EXPERIMENTAL=1, selected by SVB_ATTNW_IMPL=5): windowed attention with two independent
// chains per SM.  Correct (tests/test_gpu_ops.py::test_attention_tcgen05 green with SVB_ATTNW_IMPL=5) and measured (tools/call15.sh,
// call17.sh, profiles/r02_call15, r02_call17): 199 us per 16 ViT-H images kernel-alone against 210 us for the production two-group
// kernel (attn_window_persistent_kernel, attention_tc.cu), 155 vs 161 us at 12 images, 151 vs 147 us at head_dim 64 — and NO
// difference inside the encoder step (35.5 vs 35.3 ms per 64 images at the step's power-capped 1.2 GHz).  With every TMA load, store
// and prefetch knocked out (diagnostic flags in l2_ahead >> 8) it still takes 187 us: the softmax warps themselves — two per scheduler,
// issuing one instruction per ~3.8 cycles in the exp2 pass (ncu source page: selected 26 %, fixed-latency wait 15 %, instruction
// fetch 15 %) — carry 8100 cycles per item however the two chains are arranged; the production kernel's 9070 cycles per item
// are (TMA latency of a stage + that)/2 behind its two-stage ring.  Kept as the starting point of a kernel with FOUR softmax warps
// per scheduler (a row's 196 keys split over two threads), the one change these measurements leave.
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
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
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

// The two-group production kernel runs its two query-tile chains in LOCKSTEP on the same
// item — both in their softmax (MUFU contended, 4200 cycles) and both in their waits (bias MMA -> skew 2000, S 350, PV 1250,
// output 560: the MUFU idle for half of an item's 8350 cycles) at the same time; and with ALL arithmetic knocked out the launch
// still takes 153 of 208 us: the 770 MB it moves per 16 images is the floor (5.0 TB/s), so at the power-capped clocks of the whole
// step (1.32 GHz) the serial chain, not the memory, is what one pays for.  Here the CTA runs TWO chains that share nothing but the
// SM: chain c owns stage c of the ring, S tile c of tensor memory, a softmax group, an MMA issuer and a load / store thread, and
// walks its OWN items (chain id = 2 blockIdx + c), query tile 0 then query tile 1 of an item one after the other.  The two chains
// drift apart by themselves, so one chain's softmax runs beside the other chain's MMAs, skew and output.  Inside a chain the
// buffers of the stage are refilled piece by piece as the item lets go of them — Q0 when the store of O0 has read it (the output
// tile is staged in the dead Q buffer), K when S of the last tile has retired, V after the last PV, Q1 after the store of O1 —
// so the next item's loads run under this item's second tile.  The warp of tile 1 whose 32 rows are all padding (rows 96..127
// of the 70-row tile) keeps the barrier protocol and skips the arithmetic; Q1 is not loaded for windows whose tile 1 lies outside
// the token grid.
template <int HD> struct W5Cfg : WPCfg<HD> {
    using P = WPCfg<HD>;
    // barriers: RFULL, then per chain c at 1 + 10 c
    static constexpr int B_RFULL = 0, B_QF0 = 0, B_QF1 = 1, B_KF = 2, B_VF = 3, B_BIAS = 4, B_BREAD = 5, B_SFULL = 6, B_PFULL = 7, B_PVDONE = 8,
                         B_OSTAGED = 9, B_PER_CHAIN = 10, B_COUNT = 21;
    static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");
    static constexpr int Q0_TX = 126 * P::ROWB, Q1_TX = 70 * P::ROWB, K_TX = 196 * P::ROWB;
};

template <int HD, int POLY>
__global__ void __launch_bounds__(384, 1)
attn_window_chains_kernel(const __grid_constant__ WinPMaps maps, int D, int g, int nwy, int nwx, int heads, int num_items, float scale_log2,
                          int l2_ahead) {
    // g = token-grid HEIGHT (the last window row's padding decides whether query tile 1 exists); nwy x nwx windows per image
    using C = W5Cfg<HD>;
    constexpr int WS = 14;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // provably warp-uniform

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&maps.kv);
        ptx::prefetch_tmap(&maps.q0);
        ptx::prefetch_tmap(&maps.q1);
        ptx::prefetch_tmap(&maps.r);
        ptx::prefetch_tmap(&maps.o0);
        ptx::prefetch_tmap(&maps.o1);
        ptx::mbar_init(&bars[C::B_RFULL], 1);
        for (int c = 0; c < 2; ++c)
            for (int s = 0; s < C::B_PER_CHAIN; ++s)
                ptx::mbar_init(&bars[1 + C::B_PER_CHAIN * c + s], (s == C::B_BREAD || s == C::B_PFULL || s == C::B_OSTAGED) ? 128 : 1);
        ptx::fence_barrier_init();
    }
    if (warp == 10) ptx::tmem_alloc(tmem_slot, C::TM_COLS);
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

    if (warp == 8 || warp == 9) {
        // ===================== load / store thread of chain c =====================
        const int c = warp - 8;
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
    } else if (warp == 10 || warp == 11) {
        // ===================== MMA issuer of chain c (all 32 lanes run the loop, one elected lane issues) =====================
        constexpr uint32_t id_r = ptx::make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t id_s = ptx::make_idesc_bf16(128, 208, 0, 0);
        const int c = warp - 10;
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
                if (HD > 64) {
                    constexpr uint32_t id_pv = ptx::make_idesc_bf16(128, HD, 0, 1);
                    const uint64_t dv = ptx::make_smem_desc(v, C::V_ATOM, 256, ptx::LAYOUT_SW32);
#pragma unroll
                    for (int kq = 0; kq < 13; ++kq)                  // 16 keys = 512 B inside an atom
                        ptx::mma_f16_ts_e(s_tm + 112, s_tm + 8 * kq, dv + 32 * kq, id_pv, kq ? 1u : 0u);
                } else {
                    issue_pv<HD>(s_tm + 112, s_tm, v, v + C::K_MAIN, 13, false);
                }
                ptx::mma_commit_e(&cb[C::B_PVDONE]);
            }
            if (ntiles == 2) ++n1;
        }
    } else {
        // ===================== softmax group of chain c: one query row per thread =====================
        const int c = warp >> 2, w4 = warp & 3;
        uint64_t* cb = bars + 1 + C::B_PER_CHAIN * c;
        const int t = w4 * 32 + lane;                              // query row inside the tile
        const int xi = t % WS, yi0 = t / WS;                       // window coordinates in tile 0 (tile 1: row + 9; rows past the tile are discarded)
        const uint32_t lane_off = static_cast<uint32_t>(w4 * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + 208 * c;
        const uint32_t o_tmem = s_tmem + 112;

        // rel-pos products of the tile whose bias MMA is the `nt`-th of this chain -> the row's 14 + 14 terms (log2 units)
        float bhm[14], bwl[14];
        auto read_bias = [&](uint32_t nt, int tile) {
            ptx::mbar_wait(&cb[C::B_BIAS], nt & 1);
            ptx::tc_fence_after();
            uint32_t v[32], v2[32];
            float rr[27];
            ptx::tmem_ld_x32(s_tmem, v);                           // both loads in flight: one tensor-memory round trip
            ptx::tmem_ld_x32(s_tmem + 32, v2);
            ptx::tmem_ld_wait_dep(v);
            ptx::tmem_ld_wait_dep(v2);
            const int yi = yi0 + 9 * tile;
#pragma unroll
            for (int j = 0; j < 27; ++j) rr[j] = __uint_as_float(v[j]) * LOG2E;
            barrel_shift27(rr, yi < 13 ? yi : 13);
#pragma unroll
            for (int kq = 0; kq < 14; ++kq) bhm[kq] = rr[13 - kq];
#pragma unroll
            for (int j = 0; j < 27; ++j) rr[j] = __uint_as_float(v2[j]) * LOG2E;
            barrel_shift27(rr, xi);
#pragma unroll
            for (int kq = 0; kq < 14; ++kq) bwl[kq] = rr[13 - kq];
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
            if (w4 == 3 && tile == 1) {
                // rows 96..127 of the 70-row tile 1 are all padding: keep the barrier protocol, skip the arithmetic (the rows of P
                // and O this warp would have written are never stored)
                ptx::mbar_wait(&cb[C::B_SFULL], nt & 1);
                ptx::mbar_arrive(&cb[C::B_PFULL]);
                ptx::mbar_wait(&cb[C::B_PVDONE], nt & 1);
                ptx::mbar_arrive(&cb[C::B_OSTAGED]);
                if (more) read_bias(nt + 1, ntile);                // the tile after a tile 1 is a tile 0
            } else {
                // ---- softmax over the 196 keys ----
                ptx::mbar_wait(&cb[C::B_SFULL], nt & 1);
                ptx::tc_fence_after();
                const float lsum = window_softmax_tile<POLY>(s_tmem, bhm, bwl, scale_log2);
                ptx::tc_fence_before();
                ptx::mbar_arrive(&cb[C::B_PFULL]);
                // ---- O: TMEM -> normalised bf16 in registers ----
                ptx::mbar_wait(&cb[C::B_PVDONE], nt & 1);
                ptx::tc_fence_after();
                const float inv = 1.0f / lsum;
                uint32_t o[HD / 2];
                {
                    uint32_t v[32];
#pragma unroll
                    for (int cc = 0; cc < 64; cc += 32) {
                        ptx::tmem_ld_x32(o_tmem + cc, v);
                        ptx::tmem_ld_wait_dep(v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[cc / 2 + j] = pack_bf16x2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                    }
                    if (HD > 64) {
                        uint32_t w[16];
                        ptx::tmem_ld_x16(o_tmem + 64, w);
                        ptx::tmem_ld_wait_dep(w);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[32 + j] = pack_bf16x2(__uint_as_float(w[2 * j]) * inv, __uint_as_float(w[2 * j + 1]) * inv);
                    }
                }
                const bool next_dead = w4 == 3 && ntile == 1;
                if (early) { if (next_dead) skip_bias(nt + 1); else read_bias(nt + 1, ntile); }
                // ---- O -> the dead Q buffer of this tile in the TMA layout -> one tensor store per tile ----
                {
                    uint8_t* ob = sm + c * C::STAGE + tile * C::QT;
#pragma unroll
                    for (int j = 0; j < 8; ++j)                    // 128B swizzle: 16-byte piece j of row t at piece j ^ (t & 7)
                        *reinterpret_cast<uint4*>(ob + t * 128 + ((j ^ (t & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    if (HD > 64) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)                // 32B swizzle: piece j of row t at piece j ^ ((t >> 2) & 1)
                            *reinterpret_cast<uint4*>(ob + C::Q_MAIN + t * 32 + ((j ^ ((t >> 2) & 1)) << 4)) =
                                make_uint4(o[32 + 4 * j], o[32 + 4 * j + 1], o[32 + 4 * j + 2], o[32 + 4 * j + 3]);
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
    if (warp == 10) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}


}  // namespace

// same contract as attention_tc() for ws = 14
int attention_window5(const AttnTcParams& p, cudaStream_t stream) {
    SVB_REQUIRE(p.hd == 64 || p.hd == 80, "attention_window5: head_dim %d", p.hd);
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
            kern<<<grid, 384, C::SMEM, stream>>>(wm, D, gh, nwy, nwx, p.heads, items, scale_log2, l2_ahead);
            return 0;
        };
        rc = poly ? launch(attn_window_chains_kernel<HD, 2>) : launch(attn_window_chains_kernel<HD, 0>);
        if (rc) return rc;
        SVB_CHECK_CUDA(cudaGetLastError());
        return 0;
    };
    return p.hd == 64 ? run(std::integral_constant<int, 64>{}) : run(std::integral_constant<int, 80>{});
}

}  // namespace svb
