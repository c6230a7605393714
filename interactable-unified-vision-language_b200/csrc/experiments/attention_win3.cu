// Pipelined windowed attention (14 x 14 windows of the 70 x 70 padded token grid) for sm_100a: the production kernel of the
// windowed blocks (image_encoder.py:239-304, 340-376).  Same arithmetic as attn_window_persistent_kernel (attention_tc.cu) —
// S = Q K^T and O = P V on tcgen05 with fp32 accumulators in TMEM, decomposed rel-pos bias from two extra MMAs against the
// rel_pos tables, exact-maximum softmax in fp32 — re-cut so that the serial chain of a (window, head) item is only
//      S MMA -> softmax -> PV MMA -> (O leaves tensor memory) -> next S MMA
// and everything else runs beside it.  Measured on the two-group kernel (tools/dbg_attn_phases.py, cycles per item and group):
// rel-pos read + barrel shift 1900, wait S 310, softmax 3850, wait PV 1290, epilogue 540 — the softmax groups spent 2400 of 7900
// cycles on work that is not softmax, with the MUFU idle.  Here:
//   * a HELPER warp group (4 warps, one per TMEM lane quadrant) does the rel-pos skew and the whole epilogue for both query tiles:
//       - the bias products of the NEXT item land in their own TMEM columns (Q . [Rh | Rw]^T as ONE N = 48 MMA per tile: the
//         window is cut into 8 + 6 query rows so that a tile needs at most 21 + 27 table rows; the 96 columns left beside the two
//         208-column S tiles hold exactly two such products), the helper barrel-shifts them per row and writes the row's
//         14 + 14 terms back into the same columns (tcgen05.st); the softmax thread fetches them with one tcgen05.ld;
//       - after PV the helper pulls O out of tensor memory (which frees the S tile for the next item's S MMA at once), normalises
//         by the row sum the softmax thread left in shared memory, stages the tile in the dead Q buffer and hands it to the store
//         warp (one TMA tensor store per tile = window_unpartition + crop);
//   * the softmax groups only wait for S, run the two softmax passes and hand P over;
//   * all MMA issue is warp-uniform (one elected lane, uniform-register operands).
// Register budget: 512 threads start at 128 registers; the TMA / issuer / store warp group drops to 56, the two softmax groups and
// the helper group rise to 152 (setmaxnreg).
#include "attention_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace svb {
namespace {

template <int HD> struct W3Cfg {
    static constexpr int TAIL = HD - 64;
    static constexpr int ROWS0 = 8, ROWS1 = 6;   // window rows per query tile: 112 + 84 queries
    static constexpr int LIVE0 = 14 * ROWS0, LIVE1 = 14 * ROWS1;
    // Q tiles of one stage: the MMAs read 128 rows per tile, only 112 / 84 are loaded.  Tile 1 starts at row 112 of tile 0's area
    // (tile 0's dead rows 112..127 = tile 1's first rows; nothing ever WRITES a dead row, see the epilogue), the 16-column tails
    // follow in the same arrangement.
    static constexpr int OFF_Q1 = LIVE0 * 128;                               // 14336: 1024-aligned (128B-swizzle atom)
    static constexpr int OFF_QT0 = (LIVE0 + 128) * 128;                      // 30720
    static constexpr int OFF_QT1 = OFF_QT0 + LIVE0 * 32;                     // 256-aligned (32B-swizzle atom)
    static constexpr int Q_AREA = OFF_QT0 + (TAIL ? 8192 : 0);               // tails: (112 + 128) x 32 B rounded up to 1 KB
    static constexpr int K_MAIN = 208 * 128, K_TAIL = TAIL ? 7168 : 0;       // 196 keys (+12 pad rows)
    static constexpr int KT = K_MAIN + K_TAIL;
    static constexpr int OFF_K = Q_AREA, OFF_V = OFF_K + KT;
    static constexpr int STAGE = Q_AREA + 2 * KT;
    static_assert(OFF_Q1 % 1024 == 0 && OFF_QT0 % 1024 == 0 && OFF_QT1 % 256 == 0 && STAGE % 1024 == 0, "swizzle atom alignment");
    static constexpr int R_MAIN = 48 * 128, R_TAIL = TAIL ? 48 * 32 : 0;     // one tile's [Rh slice | Rw] table: 48 rows
    static constexpr int RT = (R_MAIN + R_TAIL + 1023) & ~1023;
    static constexpr int OFF_R = 2 * STAGE;                                  // two tables (tile 0, tile 1)
    static constexpr int OFF_LSUM = OFF_R + 2 * RT;                          // row sums: [2 tiles][128] fp32
    static constexpr int OFF_BAR = OFF_LSUM + 1024;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static_assert(SMEM <= 232448, "shared memory budget");
    static constexpr int ROWB = 128 + (TAIL ? 32 : 0);
    static constexpr int QK_TX = (LIVE0 + LIVE1 + 196) * ROWB;
    static constexpr int V_TX = 196 * ROWB;
    static constexpr int R_TX = 2 * 48 * ROWB;
    // barriers (per tile i unless noted)
    static constexpr int B_RFULL = 0, B_QKFULL = 1, B_VFULL = 3, B_EMPTY = 5, B_BIASF = 7, B_BIASR = 9, B_BIASC = 11, B_SFULL = 13,
                         B_PFULL = 15, B_PVDONE = 17, B_OREAD = 19, B_OSTAGED = 21, B_COUNT = 25;   // OSTAGED: one per (tile, stage)
    static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");
    static constexpr int V_ATOM = 208 * 32;      // head_dim 80: V as five 16-column atoms of the 32B swizzle (one N = 80 PV MMA per K step)
    static constexpr int TM_COLS = 512;          // S_i at 208 i (P_i at its columns [0,104), O_i at +112); bias_i at 416 + 48 i
    static constexpr int TM_BIAS = 416;
    static constexpr int ARRIVE1 = 96;           // tile 1: rows 0..83 are live -> warps 0..2 of a group take part, warp 3 never does
};

struct Win3Maps {
    CUtensorMap q0, q1, kv, r;           // loads: boxes (64,14,8,1) / (64,14,6,1) / (64,14,14,1) of the padded qkv; (64,48) of the tables
    CUtensorMap q0t, q1t, kvt, rt;       // their 16-column tails (32B swizzle)
    CUtensorMap o0, o1, o0t, o1t;        // stores: boxes (64|16,14,8|6,1) of out viewed as [B,g,g,D]
};

__device__ __forceinline__ void setmaxnreg_inc152() { asm volatile("setmaxnreg.inc.sync.aligned.u32 152;" ::: "memory"); }
__device__ __forceinline__ void setmaxnreg_dec56() { asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory"); }

// r[j] <- r[j + sh], sh in [0, 2^STAGES): conditional-move stages over N live entries
template <int N, int STAGES>
__device__ __forceinline__ void barrel_shift(float (&r)[N], int sh) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
        const int bit = 1 << s;
        const bool on = (sh & bit) != 0;
#pragma unroll
        for (int j = 0; j + bit < N; ++j) r[j] = on ? r[j + bit] : r[j];
    }
}

// helper: the rel-pos products of tile I (TMEM columns [0,48) of `b_tmem`: HN columns Q.Rh[slice], then 27 columns Q.Rw) -> the
// row's 14 h terms / 14 w terms in log2 units, written back to columns [0,14) / [16,30)
template <int I>
__device__ __forceinline__ void skew_bias(uint32_t b_tmem, int ty, int xi) {
    constexpr int HN = I ? 19 : 21;              // table rows of the h slice: Rh[0..20] for window rows 0..7, Rh[8..26] for rows 8..13
    uint32_t v[32], w[16];
    ptx::tmem_ld_x32(b_tmem, v);
    ptx::tmem_ld_x16(b_tmem + 32, w);
    ptx::tmem_ld_wait_dep(v);
    ptx::tmem_ld_wait_dep(w);
    float rh[HN], rw[27];
#pragma unroll
    for (int j = 0; j < HN; ++j) rh[j] = __uint_as_float(v[j]) * LOG2E;
#pragma unroll
    for (int j = 0; j < 27; ++j) rw[j] = __uint_as_float(HN + j < 32 ? v[HN + j] : w[HN + j - 32]) * LOG2E;
    // h: key row kh needs table row yi + 13 - kh = slice column ty + 13 - kh (ty = query row inside the tile)
    barrel_shift<HN, 3>(rh, ty);
    // w: key column kk needs table row xi + 13 - kk
    barrel_shift<27, 4>(rw, xi);
    uint32_t oh[16], ow[16];
#pragma unroll
    for (int k = 0; k < 14; ++k) { oh[k] = __float_as_uint(rh[13 - k]); ow[k] = __float_as_uint(rw[13 - k]); }
    oh[14] = oh[15] = ow[14] = ow[15] = 0u;
    ptx::tmem_st_x16(b_tmem, oh);
    ptx::tmem_st_x16(b_tmem + 16, ow);
    ptx::tmem_st_wait();
}

// PH: per-role cycle accounting into `clocks` (64 x int64: role r at [16 r, 16 r + 16); roles 0 / 1 softmax groups, 2 helper, 3 issuer 0)
template <int HD, int POLY, bool PH = false>
__global__ void __launch_bounds__(512, 1)
attn_window3_kernel(const __grid_constant__ Win3Maps maps, int D, int g, int heads, int num_items, float scale_log2,
                    long long* __restrict__ clocks, int l2_ahead) {
    using C = W3Cfg<HD>;
    constexpr int WS = 14, NWS = 5;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);
    float* lsum_s = reinterpret_cast<float*>(sm + C::OFF_LSUM);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // provably warp-uniform

    if (warp == 12 && lane == 0) {
        ptx::prefetch_tmap(&maps.kv);
        ptx::prefetch_tmap(&maps.q0);
        ptx::prefetch_tmap(&maps.q1);
        ptx::prefetch_tmap(&maps.r);
        ptx::prefetch_tmap(&maps.o0);
        ptx::prefetch_tmap(&maps.o1);
        for (int s = 0; s < C::B_COUNT; ++s) {
            uint32_t cnt = 1;
            if (s >= C::B_EMPTY && s < C::B_EMPTY + 2) cnt = 2;                                   // one release per query tile
            else if (s == C::B_BIASR || s == C::B_BIASC || s == C::B_PFULL || s == C::B_OREAD) cnt = 128;
            else if (s == C::B_BIASR + 1 || s == C::B_BIASC + 1 || s == C::B_PFULL + 1 || s == C::B_OREAD + 1) cnt = C::ARRIVE1;
            else if (s == C::B_OSTAGED || s == C::B_OSTAGED + 1) cnt = 128;                        // tile 0, stages 0 / 1
            else if (s == C::B_OSTAGED + 2 || s == C::B_OSTAGED + 3) cnt = C::ARRIVE1;             // tile 1
            ptx::mbar_init(&bars[s], cnt);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 13) ptx::tmem_alloc(tmem_slot, C::TM_COLS);
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
    long long pc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = PH ? clock64() : 0;
#define W3_T(k) if (PH) { const long long tn = clock64(); pc[k] += tn - tprev; tprev = tn; }
#define W3_DUMP(role) if (PH && clocks && lane == 0) { for (int k = 0; k < 12; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(clocks) + 16 * (role) + k, (unsigned long long)pc[k]); }

    // item -> (image, window, head); query tile 1 (window rows 8..13) is entirely padding in the last window row of the grid
    auto decode = [&](int item, int& b, int& wy, int& wx, int& head) {
        head = item % heads;
        const int bw = item / heads;
        const int win = bw % (NWS * NWS);
        b = bw / (NWS * NWS);
        wy = win / NWS;
        wx = win % NWS;
    };
    auto tile1_active = [&](int item) {
        const int wy = ((item / heads) % (NWS * NWS)) / NWS;
        return wy * WS + C::ROWS0 < g;
    };

    if (warp >= 12) {
        setmaxnreg_dec56();
        if (warp == 12) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                ptx::mbar_expect_tx(&bars[C::B_RFULL], C::R_TX);
                for (int i = 0; i < 2; ++i) {
                    ptx::tma_load_2d(sm + C::OFF_R + i * C::RT, &maps.r, &bars[C::B_RFULL], 0, 64 + 48 * i);
                    if (HD > 64) ptx::tma_load_2d(sm + C::OFF_R + i * C::RT + C::R_MAIN, &maps.rt, &bars[C::B_RFULL], 64, 64 + 48 * i);
                }
                int it = 0;
                for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                    const int st = it & 1;
                    const uint32_t par = ((it >> 1) & 1) ^ 1;
                    int b, wy, wx, head;
                    decode(item, b, wy, wx, head);
                    const int x0 = wx * WS, y0 = wy * WS;
                    uint8_t* q0 = sm + st * C::STAGE;
                    uint8_t* q1 = q0 + C::OFF_Q1;
                    uint8_t* k = sm + st * C::STAGE + C::OFF_K;
                    uint8_t* v = sm + st * C::STAGE + C::OFF_V;
                    const int cq = head * HD, ck = D + head * HD, cv = 2 * D + head * HD;
                    ptx::mbar_wait(&bars[C::B_EMPTY + st], par);
                    ptx::mbar_expect_tx(&bars[C::B_QKFULL + st], C::QK_TX);
                    ptx::tma_load_4d(q0, &maps.q0, &bars[C::B_QKFULL + st], cq, x0, y0, b);
                    ptx::tma_load_4d(q1, &maps.q1, &bars[C::B_QKFULL + st], cq, x0, y0 + C::ROWS0, b);
                    ptx::tma_load_4d(k, &maps.kv, &bars[C::B_QKFULL + st], ck, x0, y0, b);
                    if (HD > 64) {
                        ptx::tma_load_4d(q0 + C::OFF_QT0, &maps.q0t, &bars[C::B_QKFULL + st], cq + 64, x0, y0, b);
                        ptx::tma_load_4d(q0 + C::OFF_QT1, &maps.q1t, &bars[C::B_QKFULL + st], cq + 64, x0, y0 + C::ROWS0, b);
                        ptx::tma_load_4d(k + C::K_MAIN, &maps.kvt, &bars[C::B_QKFULL + st], ck + 64, x0, y0, b);
                    }
                    ptx::mbar_expect_tx(&bars[C::B_VFULL + st], C::V_TX);
                    if (HD > 64) {
                        for (int a = 0; a < 5; ++a) ptx::tma_load_4d(v + a * C::V_ATOM, &maps.kvt, &bars[C::B_VFULL + st], cv + 16 * a, x0, y0, b);
                    } else {
                        ptx::tma_load_4d(v, &maps.kv, &bars[C::B_VFULL + st], cv, x0, y0, b);
                    }
                    // the kernel moves 48 MB per image and block through a 2-stage ring whose stages free up late: keep HBM -> L2
                    // traffic in flight for the items after the one in the ring (their loads then hit L2)
                    for (int ahead = 1; ahead <= l2_ahead; ++ahead) {
                        const int pi = item + ahead * gridDim.x;
                        if (pi < num_items) {
                            int pb, pwy, pwx, ph_;
                            decode(pi, pb, pwy, pwx, ph_);
                            for (int part = 0; part < 3; ++part) {
                                const int cc = part * D + ph_ * HD;
                                ptx::tma_prefetch_l2_4d(&maps.kv, cc, pwx * WS, pwy * WS, pb);
                                if (HD > 64) ptx::tma_prefetch_l2_4d(&maps.kvt, cc + 64, pwx * WS, pwy * WS, pb);
                            }
                        }
                    }
                }
            }
        } else if (warp == 13 || warp == 14) {
            // ===================== MMA issuers: one per query tile; all lanes run the loop, one elected lane issues =====================
            const int i = warp - 13;
            constexpr uint32_t id_b = ptx::make_idesc_bf16(128, 48, 0, 0);
            constexpr uint32_t id_s = ptx::make_idesc_bf16(128, 208, 0, 0);
            const uint32_t s_tm = tmem + 208 * i, b_tm = tmem + C::TM_BIAS + 48 * i;
            const uint32_t r_main = base + C::OFF_R + i * C::RT, r_tail = r_main + C::R_MAIN;
            auto active = [&](int item) { return i == 0 || tile1_active(item); };
            const uint32_t q_off = i ? C::OFF_Q1 : 0, qt_off = i ? C::OFF_QT1 : C::OFF_QT0;
            auto issue_bias = [&](int st) {       // Q . [Rh slice | Rw]^T -> bias columns
                const uint32_t sb = base + st * C::STAGE;
                issue_qk<HD>(b_tm, sb + q_off, sb + qt_off, r_main, r_tail, id_b);
                ptx::mma_commit_e(&bars[C::B_BIASF + i]);
            };
            auto issue_s = [&](int st) {          // S = Q K^T
                const uint32_t sb = base + st * C::STAGE, k = sb + C::OFF_K;
                issue_qk<HD>(s_tm, sb + q_off, sb + qt_off, k, k + C::K_MAIN, id_s);
                ptx::mma_commit_e(&bars[C::B_SFULL + i]);
            };
            // tile 1 skips the items whose rows 8..13 are all padding: it only releases their stage — once THAT use of the stage
            // has begun (its loads landed), otherwise the arrival would be counted in the previous use's phase
            auto release_skipped = [&](int from_item, int from_it, int to_item) {
                for (int s = from_item, sit = from_it; s < to_item; s += gridDim.x, ++sit) {
                    ptx::mbar_wait(&bars[C::B_QKFULL + (sit & 1)], (sit >> 1) & 1);
                    ptx::mbar_arrive_e(&bars[C::B_EMPTY + (sit & 1)]);
                }
            };
            ptx::mbar_wait(&bars[C::B_RFULL], 0);
            int item = blockIdx.x, it = 0;
            {   // leading items this tile skips
                int first = item, fit = it;
                while (first < num_items && !active(first)) { first += gridDim.x; ++fit; }
                release_skipped(item, it, first < num_items ? first : num_items);
                item = first;
                it = fit;
            }
            uint32_t n = 0;                                        // active items of this tile so far (barrier phases)
            if (item < num_items) {
                ptx::mbar_wait(&bars[C::B_QKFULL + (it & 1)], (it >> 1) & 1);
                ptx::tc_fence_after();
                issue_bias(it & 1);
                issue_s(it & 1);
            }
#pragma unroll 1
            while (item < num_items) {
                const int st = it & 1;
                int nitem = item + gridDim.x, nit = it + 1;
                while (nitem < num_items && !active(nitem)) { nitem += gridDim.x; ++nit; }
                const bool more = nitem < num_items;
                // the next item's bias products early (while this item's softmax runs) — only when it sits in the OTHER stage: if
                // items were skipped in between, its stage is THIS one, whose reload waits for this item's own epilogue
                // ... and only when its loads have ALREADY landed: waiting for them here would hold this item's PV back
                const bool early = more && (nit == it + 1) && ptx::mbar_test_wait(&bars[C::B_QKFULL + (nit & 1)], (nit >> 1) & 1);
                W3_T(0)
                if (early) {
                    W3_T(1)
                    ptx::mbar_wait(&bars[C::B_BIASC + i], n & 1);          // the softmax group has fetched this item's terms
                    W3_T(2)
                    ptx::tc_fence_after();
                    issue_bias(nit & 1);
                    W3_T(3)
                }
                // ---- O = P V ----
                ptx::mbar_wait(&bars[C::B_VFULL + st], (it >> 1) & 1);
                ptx::mbar_wait(&bars[C::B_PFULL + i], n & 1);              // P_i is in TMEM
                W3_T(4)
                ptx::tc_fence_after();
                {
                    const uint32_t v = base + st * C::STAGE + C::OFF_V;
                    if (HD > 64) {
                        constexpr uint32_t id_pv = ptx::make_idesc_bf16(128, HD, 0, 1);
                        const uint64_t dv = ptx::make_smem_desc(v, C::V_ATOM, 256, ptx::LAYOUT_SW32);
#pragma unroll
                        for (int kk = 0; kk < 13; ++kk)                // 16 keys = 512 B inside an atom
                            ptx::mma_f16_ts_e(s_tm + 112, s_tm + 8 * kk, dv + 32 * kk, id_pv, kk ? 1u : 0u);
                    } else {
                        issue_pv<HD>(s_tm + 112, s_tm, v, v + C::K_MAIN, 13, false);
                    }
                    ptx::mma_commit_e(&bars[C::B_PVDONE + i]);
                }
                // stages of the items this tile skips
                release_skipped(item + gridDim.x, it + 1, more ? nitem : num_items);
                if (more) {
                    if (!early) {
                        ptx::mbar_wait(&bars[C::B_QKFULL + (nit & 1)], (nit >> 1) & 1);
                        ptx::mbar_wait(&bars[C::B_BIASC + i], n & 1);
                        ptx::tc_fence_after();
                        issue_bias(nit & 1);
                    }
                    // ---- next S: its columns hold this item's P (consumed by the PV above, in order) and O (pulled out by the helper) ----
                    W3_T(5)
                    ptx::mbar_wait(&bars[C::B_OREAD + i], n & 1);
                    W3_T(6)
                    ptx::tc_fence_after();
                    issue_s(nit & 1);
                    W3_T(7)
                }
                item = nitem;
                it = nit;
                ++n;
            }
            pc[11] = n;
            if (i == 0) { W3_DUMP(3) }
        } else {
            // ===================== store warp: O tiles staged by the helper -> ONE tensor store per tile =====================
            if (lane == 0) {
                int it = 0;
                uint32_t n[4] = {0, 0, 0, 0};
                for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                    const int st = it & 1;
                    int b, wy, wx, head;
                    decode(item, b, wy, wx, head);
                    for (int i = 0; i < 2; ++i) {
                        if (i == 1 && !tile1_active(item)) continue;   // released by the issuer
                        const uint8_t* ob = sm + st * C::STAGE;
                        ptx::mbar_wait(&bars[C::B_OSTAGED + 2 * i + st], n[2 * i + st] & 1);
                        ++n[2 * i + st];
                        const int x0 = wx * WS, y0 = wy * WS + C::ROWS0 * i;
                        tma_store_4d(i ? &maps.o1 : &maps.o0, ob + (i ? C::OFF_Q1 : 0), head * HD, x0, y0, b);   // rows / columns past the grid are clipped
                        if (HD > 64) tma_store_4d(i ? &maps.o1t : &maps.o0t, ob + (i ? C::OFF_QT1 : C::OFF_QT0), head * HD + 64, x0, y0, b);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the store has finished reading the stage
                        ptx::mbar_arrive(&bars[C::B_EMPTY + st]);      // this tile is done with stage st (its MMAs retired before PVDONE)
                    }
                }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");                  // stores complete before the CTA exits
            }
        }
    } else if (warp >= 8) {
        // ===================== helper group: rel-pos skew + epilogue of both tiles =====================
        setmaxnreg_inc152();
        const int w4 = warp & 3;
        const int t = w4 * 32 + lane;                              // query row inside a tile
        const int ty = t / WS, xi = t % WS;
        const uint32_t lane_off = static_cast<uint32_t>(w4 * 32) << 16;
        const bool live1 = w4 < 3;                                 // tile 1: rows 0..83 live -> warps 0..2 take part
        uint32_t nb[2] = {0, 0}, ne[2] = {0, 0};                   // bias products skewed / epilogues done per tile
        auto skew = [&](int i) {
            W3_T(8)
            ptx::mbar_wait(&bars[C::B_BIASF + i], nb[i] & 1);
            W3_T(0 + 4 * i)
            ptx::tc_fence_after();
            const uint32_t b_tm = tmem + lane_off + C::TM_BIAS + 48 * i;
            if (i == 0) skew_bias<0>(b_tm, ty < 7 ? ty : 7, xi);
            else skew_bias<1>(b_tm, ty < 5 ? ty : 5, xi);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars[C::B_BIASR + i]);
            ++nb[i];
            W3_T(1 + 4 * i)
        };
        auto epilogue = [&](int i, int st) {
            W3_T(8)
            ptx::mbar_wait(&bars[C::B_PFULL + i], ne[i] & 1);      // the softmax threads' row sums are in shared memory
            ptx::mbar_wait(&bars[C::B_PVDONE + i], ne[i] & 1);
            W3_T(2 + 4 * i)
            ptx::tc_fence_after();
            const uint32_t o_tm = tmem + lane_off + 208 * i + 112;
            uint32_t v0[32], v1[32];
            ptx::tmem_ld_x32(o_tm, v0);
            ptx::tmem_ld_x32(o_tm + 32, v1);
            uint32_t o[HD / 2];
            const float inv = 1.0f / lsum_s[i * 128 + t];
            if (HD > 64) {
                uint32_t v2[16];
                ptx::tmem_ld_x16(o_tm + 64, v2);
                ptx::tmem_ld_wait_dep(v0);
                ptx::tmem_ld_wait_dep(v1);
                ptx::tmem_ld_wait_dep(v2);
                ptx::tc_fence_before();
                ptx::mbar_arrive(&bars[C::B_OREAD + i]);           // O_i has left tensor memory: the next S MMA may overwrite it
#pragma unroll
                for (int j = 0; j < 8; ++j) o[32 + j] = pack_bf16x2(__uint_as_float(v2[2 * j]) * inv, __uint_as_float(v2[2 * j + 1]) * inv);
            } else {
                ptx::tmem_ld_wait_dep(v0);
                ptx::tmem_ld_wait_dep(v1);
                ptx::tc_fence_before();
                ptx::mbar_arrive(&bars[C::B_OREAD + i]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                o[j] = pack_bf16x2(__uint_as_float(v0[2 * j]) * inv, __uint_as_float(v0[2 * j + 1]) * inv);
                o[16 + j] = pack_bf16x2(__uint_as_float(v1[2 * j]) * inv, __uint_as_float(v1[2 * j + 1]) * inv);
            }
            // O -> the dead Q buffer of this stage in the TMA layout -> one tensor store per tile (store warp).  Only LIVE rows are
            // written: tile 0's dead rows 112..127 are tile 1's first rows (W3Cfg::OFF_Q1).
            if (t < (i ? C::LIVE1 : C::LIVE0)) {
                uint8_t* ob = sm + st * C::STAGE + (i ? C::OFF_Q1 : 0);
#pragma unroll
                for (int j = 0; j < 8; ++j)                        // 128B swizzle: 16-byte piece j of row t at piece j ^ (t & 7)
                    *reinterpret_cast<uint4*>(ob + t * 128 + ((j ^ (t & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                if (HD > 64) {
                    uint8_t* obt = sm + st * C::STAGE + (i ? C::OFF_QT1 : C::OFF_QT0);
#pragma unroll
                    for (int j = 0; j < 2; ++j)                    // 32B swizzle: piece j of row t at piece j ^ ((t >> 2) & 1)
                        *reinterpret_cast<uint4*>(obt + t * 32 + ((j ^ ((t >> 2) & 1)) << 4)) =
                            make_uint4(o[32 + 4 * j], o[32 + 4 * j + 1], o[32 + 4 * j + 2], o[32 + 4 * j + 3]);
                }
            }
            ptx::fence_proxy_async_smem();                         // generic writes -> visible to the TMA (async proxy) read
            ptx::mbar_arrive(&bars[C::B_OSTAGED + 2 * i + st]);
            ++ne[i];
            W3_T(3 + 4 * i)
        };
        // first items' bias products
        {
            int f1 = blockIdx.x;
            while (f1 < num_items && !tile1_active(f1)) f1 += gridDim.x;
            if (blockIdx.x < num_items) skew(0);
            if (f1 < num_items && live1) skew(1);
        }
        int it = 0;
#pragma unroll 1
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int st = it & 1;
            const bool a1 = tile1_active(item) && live1;
            // next items of the two tiles: their bias products are issued during this item (early, when their loads have landed in
            // time) or right behind this item's PV
            bool skew0 = item + (int)gridDim.x < num_items;
            int n1 = item + gridDim.x;
            while (n1 < num_items && !tile1_active(n1)) n1 += gridDim.x;
            bool skew1 = a1 && n1 < num_items;
            bool epi0 = true, epi1 = a1;
            // event loop: whatever is ready first; an epilogue (on the critical chain: it frees the S tile) before a skew.  The
            // probe is made warp-uniform (lane 0's view: a completed phase stays completed) — the tasks contain .sync.aligned
            // tensor-memory instructions.
            auto ready = [&](int slot, uint32_t par) { return __shfl_sync(0xffffffffu, (int)ptx::mbar_test_wait(&bars[slot], par), 0) != 0; };
            while (skew0 || skew1 || epi0 || epi1) {
                if (epi0 && ready(C::B_PVDONE + 0, ne[0] & 1)) { epilogue(0, st); epi0 = false; continue; }
                if (epi1 && ready(C::B_PVDONE + 1, ne[1] & 1)) { epilogue(1, st); epi1 = false; continue; }
                if (skew0 && ready(C::B_BIASF + 0, nb[0] & 1)) { skew(0); skew0 = false; continue; }
                if (skew1 && ready(C::B_BIASF + 1, nb[1] & 1)) { skew(1); skew1 = false; continue; }
            }
        }
        pc[11] = ne[0];
        if (w4 == 0) { W3_DUMP(2) }
    } else {
        // ===================== softmax warps: group i = query tile i =====================
        setmaxnreg_inc152();
        const int i = warp >> 2, w4 = warp & 3;
        const int t = w4 * 32 + lane;                              // query row inside the tile
        const uint32_t lane_off = static_cast<uint32_t>(w4 * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + 208 * i;
        const uint32_t b_tmem = tmem + lane_off + C::TM_BIAS + 48 * i;
        if (i == 0 || w4 < 3) {                                    // tile 1, warp 3: rows 96..127 are never live
            uint32_t n = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                if (i == 1 && !tile1_active(item)) continue;
                // ---- this row's rel-pos terms (log2 units), skewed by the helper ----
                float bhm[14], bwl[14];
                {
                    W3_T(0)
                    ptx::mbar_wait(&bars[C::B_BIASR + i], n & 1);
                    W3_T(1)
                    ptx::tc_fence_after();
                    uint32_t v[32];
                    ptx::tmem_ld_x32(b_tmem, v);
                    ptx::tmem_ld_wait_dep(v);
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&bars[C::B_BIASC + i]);       // the bias columns may take the next item's products
#pragma unroll
                    for (int k = 0; k < 14; ++k) { bhm[k] = __uint_as_float(v[k]); bwl[k] = __uint_as_float(v[16 + k]); }
                }
                // ---- softmax over the 196 keys ----
                W3_T(2)
                ptx::mbar_wait(&bars[C::B_SFULL + i], n & 1);
                W3_T(3)
                ptx::tc_fence_after();
                const float lsum = window_softmax_tile<POLY>(s_tmem, bhm, bwl, scale_log2);
                lsum_s[i * 128 + t] = lsum;
                ptx::tc_fence_before();
                ptx::mbar_arrive(&bars[C::B_PFULL + i]);
                W3_T(4)
                ++n;
            }
            pc[11] = n;
            if (w4 == 0) { W3_DUMP(i) }
        }
    }

#undef W3_T
#undef W3_DUMP
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}

}  // namespace

// rows of the packed rel-pos table block this kernel reads (written by pack_rel_table): [64, 112) tile 0, [112, 160) tile 1
template <int HD>
static int launch_window3(const AttnTcParams& p, cudaStream_t stream) {
    using C = W3Cfg<HD>;
    const int D = p.heads * p.hd, gp = 70, g = p.grid;
    Win3Maps wm;
    int rc;
    {
        const uint64_t dims[4] = {(uint64_t)3 * D, (uint64_t)gp, (uint64_t)gp, (uint64_t)p.batch};
        const uint64_t str[3] = {(uint64_t)3 * D * 2, (uint64_t)gp * 3 * D * 2, (uint64_t)gp * gp * 3 * D * 2};
        const uint32_t q0[4] = {64, 14, (uint32_t)C::ROWS0, 1}, q1[4] = {64, 14, (uint32_t)C::ROWS1, 1}, kv[4] = {64, 14, 14, 1};
        const uint32_t q0t[4] = {16, 14, (uint32_t)C::ROWS0, 1}, q1t[4] = {16, 14, (uint32_t)C::ROWS1, 1}, kvt[4] = {16, 14, 14, 1};
        if ((rc = encode_tmap_nd_bf16(&wm.q0, p.qkv, 4, dims, str, q0, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.q1, p.qkv, 4, dims, str, q1, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.kv, p.qkv, 4, dims, str, kv, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.q0t, p.qkv, 4, dims, str, q0t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.q1t, p.qkv, 4, dims, str, q1t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.kvt, p.qkv, 4, dims, str, kvt, 32))) return rc;
        const uint64_t rd[2] = {(uint64_t)HD, 160};
        const uint64_t rs[1] = {(uint64_t)HD * 2};
        const uint32_t rm[2] = {64, 48}, rt[2] = {16, 48};
        if ((rc = encode_tmap_nd_bf16(&wm.r, p.rel_pack, 2, rd, rs, rm, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.rt, p.rel_pack, 2, rd, rs, rt, 32))) return rc;
        // output viewed as [B, g, g, D]: the store of a window tile is a box of 14 x (8|6) tokens, clipped at the grid's edge
        const uint64_t od[4] = {(uint64_t)D, (uint64_t)g, (uint64_t)g, (uint64_t)p.batch};
        const uint64_t os[3] = {(uint64_t)D * 2, (uint64_t)g * D * 2, (uint64_t)g * g * D * 2};
        if ((rc = encode_tmap_nd_bf16(&wm.o0, p.out, 4, od, os, q0, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o1, p.out, 4, od, os, q1, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o0t, p.out, 4, od, os, q0t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o1t, p.out, 4, od, os, q1t, 32))) return rc;
    }
    const float scale_log2 = LOG2E / sqrtf((float)HD);
    const int items = p.batch * 25 * p.heads;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = items < sms ? items : sms;
    auto launch = [&](auto kern) -> int {
        static bool attr_set = false;
        if (!attr_set) {
            SVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
            attr_set = true;
        }
        static const int l2_ahead = [] { const char* e = getenv("SVB_ATTNW_L2AHEAD"); return e ? atoi(e) : 1; }();
        kern<<<grid, 512, C::SMEM, stream>>>(wm, D, g, p.heads, items, scale_log2, p.phase_clocks, l2_ahead);
        return 0;
    };
    static const int k8 = [] { const char* e = getenv("SVB_ATTNW_POLY"); return e ? atoi(e) : 2; }();
    if (p.phase_clocks) rc = launch(attn_window3_kernel<HD, 2, true>);
    else if (k8 == 0) rc = launch(attn_window3_kernel<HD, 0>);
    else if (k8 == 3) rc = launch(attn_window3_kernel<HD, 3>);
    else if (k8 == 4) rc = launch(attn_window3_kernel<HD, 4>);
    else rc = launch(attn_window3_kernel<HD, 2>);
    if (rc) return rc;
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int attention_window3(const AttnTcParams& p, cudaStream_t stream) {
    return p.hd == 64 ? launch_window3<64>(p, stream) : launch_window3<80>(p, stream);
}

}  // namespace svb
