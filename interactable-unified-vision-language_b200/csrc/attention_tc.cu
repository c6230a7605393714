// tcgen05 / TMEM attention with the decomposed relative-position bias built in-kernel (sm_100a, bf16 in / bf16 out).
//
// Restates Attention.forward + add_decomposed_rel_pos (image_encoder.py:239-255, 340-376) fused with window_partition /
// window_unpartition (image_encoder.py:258-304) for the two geometries _build_sam instantiates (build_sam.py:60-73):
// a 64x64 token grid with 14x14 windows (25 windows of 196 keys on the 70x70 padded grid) and global 64x64 attention.
//
//   * Q, K, V tiles come straight out of the qkv GEMM output by TMA (cp.async.bulk.tensor, 128B swizzle for the first
//     64 head-dim columns, 32B swizzle for the 16-column tail of head_dim 80).  A window is ONE 4-D TMA box
//     (64 cols x 14 x 14) of the padded [B,70,70,3D] tensor, so window_partition never materialises; the output is written
//     at the token's own position, so window_unpartition + crop never materialise either.
//   * S = Q K^T and O = P V run on the tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM).  P is written
//     back to TMEM as bf16 over the S columns and fed to the second MMA as the TMEM A operand — it never touches
//     shared or global memory.
//   * rel-pos: bias(q,k) = q.Rh[qh-kh+ws-1] + q.Rw[qw-kw+ws-1] (unscaled q, image_encoder.py:369-374).  The products
//     Q.Rh^T and Q.Rw^T against the rel_pos tables are two more tcgen05 MMAs per query tile; each softmax thread (one
//     query row) pulls its own skewed slice into REGISTERS (w term) / a per-warp column of shared memory (h term) and
//     adds it in the exp2 argument.  Neither the gathered (S,S,hd) tables nor the (S,S) bias exist anywhere.
//   * softmax in fp32, base-2 with the scale folded in, lazy rescaling of O (only when the running bound grows by 2^8).
//
// Global kernel:   CTA = 256 queries (4 grid rows) of one (image, head); 8 softmax warps (2 query tiles ping-pong on the
//                  tensor core), 1 TMA warp, 1 MMA warp; K/V tiles of 128 keys through an NST-deep ring.
// Windowed kernel: CTA = one query tile (9 or 5 window rows) of one (image, window, head); 4 softmax warps + TMA + MMA;
//                  2 CTAs per SM.
#include "attention_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace svb {
int encode_tmap_nd_bf16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes);

namespace {

// ================================================================================================================
//                                              GLOBAL ATTENTION (64 x 64 keys)
// ================================================================================================================
template <int HD, int NST_> struct GCfg {
    static constexpr int NST = NST_;
    static constexpr int TAIL = HD - 64;
    static constexpr int T_MAIN = 128 * 128;                  // 128 rows x 64 bf16, 128B swizzle
    static constexpr int T_TAIL = TAIL ? 128 * 32 : 0;        // 128 rows x 16 bf16, 32B swizzle
    static constexpr int TILE = T_MAIN + T_TAIL;              // Q and K tiles (K-major operands: main + tail MMAs along K)
    static constexpr int VTILE = TAIL ? 2 * T_MAIN : T_MAIN;  // V tile (MN-major operand): two 64-column atoms, the second one holds
                                                              // columns 64..79 (+ 48 columns of whatever follows, never read)
    static constexpr int OFF_Q = 0;                           // 2 query tiles
    static constexpr int OFF_K = 2 * TILE;                    // NST key tiles
    static constexpr int OFF_V = OFF_K + NST * TILE;          // NST value tiles
    static constexpr int OFF_STG = OFF_V + NST * VTILE;       // 8 warps x 8 KB: w-term staging, then the h-term columns
    static constexpr int STG_BYTES = 8 * 8192;
    static constexpr int OFF_RW = OFF_STG;                    // rel tables alias the staging area (dead after the bias MMAs)
    static constexpr int RH_MAIN = 80 * 128;
    static constexpr int RH_TAIL = TAIL ? 3072 : 0;           // 80 x 32 B rounded up to 1 KB
    static constexpr int OFF_RH = OFF_RW + TILE;
    static constexpr int OFF_BAR = OFF_STG + STG_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int Q_TX = 2 * (128 * 128 + (TAIL ? 128 * 32 : 0)) + (128 * 128 + (TAIL ? 128 * 32 : 0)) +
                                (80 * 128 + (TAIL ? 80 * 32 : 0));
    static constexpr int KV_TX = 128 * 128 + (TAIL ? 128 * 32 : 0);
    static constexpr int V_TX = VTILE;
    static_assert(OFF_RH + RH_MAIN + RH_TAIL <= OFF_STG + STG_BYTES, "rel tables must fit in the staging area");
    static_assert(SMEM <= 232448, "shared memory budget");
    // barrier slots
    static constexpr int B_QFULL = 0, B_BIAS = 1, B_KFULL = 2, B_KEMPTY = B_KFULL + NST, B_VFULL = B_KEMPTY + NST,
                         B_VEMPTY = B_VFULL + NST, B_SFULL = B_VEMPTY + NST, B_PFULL = B_SFULL + 2, B_PVDONE = B_PFULL + 2,
                         B_COUNT = B_PVDONE + 2;
    static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");
    // TMEM columns
    static constexpr int TM_S = 0, TM_O = 256, TM_COLS = 512;   // S_i at 128*i (P_i aliases its first 64), O_i at 256 + 128*i
};

// ================================================================================================================
//                        GLOBAL ATTENTION, half-tile pipeline (64-key S tiles, double-buffered per query tile)
// ================================================================================================================
// Measured on the kernel above (tools/dbg_attn_g_phases.py): a softmax group spends 1925 of 4400 cycles per 128-key tile
// WAITING for S — the chain P(j) -> PV(j) -> QK(j+1) -> S(j+1) is serial per query tile because P(j) aliases S(j).  Here one S
// tile is one KEY ROW of the image (64 keys, 64 fp32 columns) and every query tile owns two of them: QK(h+2) is issued right
// after PV(h), while the group is still working on S(h+1), so S is always resident when the group asks for it.  TMEM:
// 4 x 64 (S) + 2 x 80 (O) = 416 columns (two 128-key S buffers per query tile would need 672).  The K / V ring keeps 128-key
// TMA tiles; a half tile is a descriptor offset of 64 rows.
template <int HD, int NST_> struct G2Cfg : GCfg<HD, NST_> {
    static constexpr int NST = NST_;
    static constexpr int B_QFULL = 0, B_BIAS = 1, B_BREAD = 3, B_KFULL = 5, B_KEMPTY = B_KFULL + NST, B_VFULL = B_KEMPTY + NST,
                         B_VEMPTY = B_VFULL + NST, B_SFULL = B_VEMPTY + NST, B_PFULL = B_SFULL + 4, B_PVDONE = B_PFULL + 4,
                         B_ODONE = B_PVDONE + 2, B_COUNT = B_ODONE + 2;
    static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");
    // S_i^b at 64*(2i+b) (P_i^b aliases its first 32), O_i at 256 + 80*i, Q_i (bf16 pairs, the A operand of QK) at 416 + 48*i
    static constexpr int TM_S = 0, TM_O = 256, O_STRIDE = 80, TM_Q = 416, Q_STRIDE = 48, TM_COLS = 512;
};

// D[128 x N] (+)= Q[128 x HD] (bf16 pairs in TMEM) * B[N x HD]^T (K-major in smem): no shared-memory read of the A operand
template <int HD>
__device__ __forceinline__ void issue_qk_ts(uint32_t d_tmem, uint32_t q_tmem, uint32_t b_main, uint32_t b_tail, uint32_t idesc) {
    const uint64_t db = desc_k128(b_main);
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::mma_f16_ts_e(d_tmem, q_tmem + 8 * k, db + 2 * k, idesc, k ? 1u : 0u);
    if (HD > 64) ptx::mma_f16_ts_e(d_tmem, q_tmem + 32, desc_k32(b_tail), idesc, 1u);
}

template <int HD, int NST, bool PH, int POLY>
__global__ void __launch_bounds__(352, 1)
attn_global2_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_tail,
                    const __grid_constant__ CUtensorMap tm_rw_main, const __grid_constant__ CUtensorMap tm_rw_tail,
                    const __grid_constant__ CUtensorMap tm_rh_main, const __grid_constant__ CUtensorMap tm_rh_tail,
                    const bf16* __restrict__ qkv, bf16* __restrict__ out, int D, int T, float scale_log2,
                    long long* __restrict__ phase_clocks, int order) {
    using C = G2Cfg<HD, NST>;
    constexpr int NKT = 32;                                     // 128-key TMA tiles
    constexpr int NH = 64;                                      // 64-key half tiles = key rows of the image
    constexpr int HALF_MAIN = 64 * 128, HALF_TAIL = 64 * 32;    // byte offset of keys 64.. inside a K / V tile
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);

    // warp index as a provably warp-uniform value: the issuer warps' address arithmetic then stays in uniform registers
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int pair = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int row0 = b * T + pair * 256;
    // the proj GEMM behind this kernel is launched with programmatic stream serialization: let its CTAs become resident as ours exit
    // (it waits for this grid's completion before it reads `out`)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int colq = head * HD, colk = D + head * HD, colv = 2 * D + head * HD;

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&tm_main);
        ptx::prefetch_tmap(&tm_rw_main);
        ptx::prefetch_tmap(&tm_rh_main);
        if (HD > 64) { ptx::prefetch_tmap(&tm_tail); ptx::prefetch_tmap(&tm_rw_tail); ptx::prefetch_tmap(&tm_rh_tail); }
        ptx::mbar_init(&bars[C::B_QFULL], 1);
        for (int s = 0; s < NST; ++s) {
            ptx::mbar_init(&bars[C::B_KFULL + s], 1);
            ptx::mbar_init(&bars[C::B_KEMPTY + s], 2);             // one release per issuer (query tile)
            ptx::mbar_init(&bars[C::B_VFULL + s], 1);
            ptx::mbar_init(&bars[C::B_VEMPTY + s], 2);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&bars[C::B_BIAS + i], 1);
            ptx::mbar_init(&bars[C::B_BREAD + i], 128);
            ptx::mbar_init(&bars[C::B_PVDONE + i], 1);
            ptx::mbar_init(&bars[C::B_ODONE + i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            ptx::mbar_init(&bars[C::B_SFULL + i], 1);
            ptx::mbar_init(&bars[C::B_PFULL + i], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 9) ptx::tmem_alloc(tmem_slot, C::TM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");          // dependent launch: the set-up above overlapped the qkv GEMM's tail

    if (warp == 8) {
        // ===================== TMA producer (as in attn_global_kernel) =====================
        if (lane == 0) {
            ptx::mbar_expect_tx(&bars[C::B_QFULL], C::Q_TX);
            for (int i = 0; i < 2; ++i) {
                ptx::tma_load_2d(sm + C::OFF_Q + i * C::TILE, &tm_main, &bars[C::B_QFULL], colq, row0 + 128 * i);
                if (HD > 64) ptx::tma_load_2d(sm + C::OFF_Q + i * C::TILE + C::T_MAIN, &tm_tail, &bars[C::B_QFULL], colq + 64, row0 + 128 * i);
            }
            ptx::tma_load_2d(sm + C::OFF_RW, &tm_rw_main, &bars[C::B_QFULL], 0, 144);
            ptx::tma_load_2d(sm + C::OFF_RH, &tm_rh_main, &bars[C::B_QFULL], 0, 4 * pair);
            if (HD > 64) {
                ptx::tma_load_2d(sm + C::OFF_RW + C::T_MAIN, &tm_rw_tail, &bars[C::B_QFULL], 64, 144);
                ptx::tma_load_2d(sm + C::OFF_RH + C::RH_MAIN, &tm_rh_tail, &bars[C::B_QFULL], 64, 4 * pair);
            }
            for (int j = 0; j < NKT; ++j) {
                const int st = j % NST;
                const uint32_t par = ((j / NST) & 1) ^ 1;
                const int krow = b * T + j * 128;
                ptx::mbar_wait(&bars[C::B_KEMPTY + st], par);
                ptx::mbar_expect_tx(&bars[C::B_KFULL + st], C::KV_TX);
                ptx::tma_load_2d(sm + C::OFF_K + st * C::TILE, &tm_main, &bars[C::B_KFULL + st], colk, krow);
                if (HD > 64) ptx::tma_load_2d(sm + C::OFF_K + st * C::TILE + C::T_MAIN, &tm_tail, &bars[C::B_KFULL + st], colk + 64, krow);
                ptx::mbar_wait(&bars[C::B_VEMPTY + st], par);
                ptx::mbar_expect_tx(&bars[C::B_VFULL + st], C::V_TX);
                ptx::tma_load_2d(sm + C::OFF_V + st * C::VTILE, &tm_main, &bars[C::B_VFULL + st], colv, krow);
                if (HD > 64) ptx::tma_load_2d(sm + C::OFF_V + st * C::VTILE + C::T_MAIN, &tm_main, &bars[C::B_VFULL + st], colv + 64, krow);
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ===================== MMA issuers: one per query tile =====================
        // All 32 lanes run this loop (warp-uniform control flow); every tcgen05.mma / commit is issued by one elected lane inside
        // the asm block, with operands in uniform registers.  (Under `if (lane == 0)` each MMA cost a ~14-instruction waterfall
        // loop, 60-110 cycles of issue per MMA for a 44.5-cycle MMA: 52 MMAs + 10 commits per 128 keys made ONE issuer the
        // bottleneck of this pipeline, and two still spent ~2600 of 4150 cycles per 128 keys issuing.)
        {
            const int i = warp - 9;
            constexpr uint32_t id_w = ptx::make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t id_s = ptx::make_idesc_bf16(128, 64, 0, 0);
            constexpr uint32_t id_rh = ptx::make_idesc_bf16(128, 80, 0, 0);
            const uint32_t q_main = base + C::OFF_Q + i * C::TILE, q_tail = q_main + C::T_MAIN;
            const uint32_t q_tm = tmem + C::TM_Q + C::Q_STRIDE * i;
            const uint32_t o_tm = tmem + C::TM_O + C::O_STRIDE * i;
            ptx::mbar_wait(&bars[C::B_QFULL], 0);
            ptx::tc_fence_after();
            // decomposed rel-pos products: Q.Rw^T -> both S buffers of the tile (128 columns), Q.Rh[4*pair ..]^T -> O_i columns
            issue_qk<HD>(tmem + C::TM_S + 128 * i, q_main, q_tail, base + C::OFF_RW, base + C::OFF_RW + C::T_MAIN, id_w);
            issue_qk<HD>(o_tm, q_main, q_tail, base + C::OFF_RH, base + C::OFF_RH + C::RH_MAIN, id_rh);
            ptx::mma_commit_e(&bars[C::B_BIAS + i]);
            // S_i(0), S_i(1): the two halves of key tile 0
            ptx::mbar_wait(&bars[C::B_KFULL + 0], 0);
            ptx::mbar_wait(&bars[C::B_BREAD + i], 0);              // bias products consumed (S_i / O_i columns are free), Q_i is in TMEM
            ptx::tc_fence_after();
            for (int hb = 0; hb < 2; ++hb) {
                issue_qk_ts<HD>(tmem + C::TM_S + 64 * (2 * i + hb), q_tm, base + C::OFF_K + hb * HALF_MAIN,
                                base + C::OFF_K + C::T_MAIN + hb * HALF_TAIL, id_s);
                ptx::mma_commit_e(&bars[C::B_SFULL + 2 * i + hb]);
            }
            ptx::mma_commit_e(&bars[C::B_KEMPTY + 0]);
            long long ipc[3] = {0, 0, 0};
            long long itp = PH ? clock64() : 0;
#define SVB_IPH(k) if (PH) { const long long tn = clock64(); ipc[k] += tn - itp; itp = tn; }
#pragma unroll 1
            for (int h = 0; h < NH; ++h) {
                const int hb = h & 1, jt = h >> 1, st = jt % NST;
                const bool more = (h + 2 < NH);
                const int jt2 = (h + 2) >> 1, st2 = jt2 % NST;
                SVB_IPH(2)
                if (hb == 0) {
                    ptx::mbar_wait(&bars[C::B_VFULL + st], (jt / NST) & 1);
                    if (more) ptx::mbar_wait(&bars[C::B_KFULL + st2], (jt2 / NST) & 1);
                }
                SVB_IPH(0)
                const uint32_t s_i = tmem + C::TM_S + 64 * (2 * i + hb);
                ptx::mbar_wait(&bars[C::B_PFULL + 2 * i + hb], jt & 1);      // P_i(h) is in TMEM
                SVB_IPH(1)
                ptx::tc_fence_after();
                issue_pv_wide<HD>(o_tm, s_i, base + C::OFF_V + st * C::VTILE + hb * HALF_MAIN, C::T_MAIN, 4, h > 0);
                ptx::mma_commit_e(&bars[C::B_PVDONE + i]);
                if (h == NH - 1) ptx::mma_commit_e(&bars[C::B_ODONE + i]);     // O_i is complete
                if (hb == 1) ptx::mma_commit_e(&bars[C::B_VEMPTY + st]);
                if (more) {
                    // in-order execution of this warp's MMAs: the overwrite of S_i^hb / P_i^hb follows the PV above
                    issue_qk_ts<HD>(s_i, q_tm, base + C::OFF_K + st2 * C::TILE + hb * HALF_MAIN,
                                    base + C::OFF_K + st2 * C::TILE + C::T_MAIN + hb * HALF_TAIL, id_s);
                    ptx::mma_commit_e(&bars[C::B_SFULL + 2 * i + hb]);
                    if (hb == 1) ptx::mma_commit_e(&bars[C::B_KEMPTY + st2]);
                }
            }
#undef SVB_IPH
            if (PH && phase_clocks && i == 0 && lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + 3, (unsigned long long)ipc[0]);       // wait K / V
                atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + 4, (unsigned long long)ipc[1]);       // wait P
                atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + 5, (unsigned long long)ipc[2]);       // issue
            }
        }
    } else {
        // ===================== softmax warps: 2 groups of 128 query rows =====================
        const int i = warp >> 2;                                   // query tile of this warp group
        const int w4 = warp & 3;                                   // TMEM lane quadrant
        const int t = w4 * 32 + lane;                              // query row inside the tile
        const int qr = 2 * i + (t >> 6);                           // grid row of the query inside the CTA (0..3), warp-uniform
        const int qw = t & 63;                                     // grid column of the query
        const uint32_t lane_off = static_cast<uint32_t>(w4 * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + C::TM_S + 128 * i;
        const uint32_t o_tmem = tmem + lane_off + C::TM_O + C::O_STRIDE * i;
        float* stg = reinterpret_cast<float*>(sm + C::OFF_STG + warp * 8192);   // [64][32] fp32, private to this warp

        // ---- this thread's query row -> TMEM (bf16 pairs, the layout P has): the A operand of every QK product ----
        {
            const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)(row0 + 128 * i + t) * (3 * D) + colq);
            const uint32_t q_tmem = tmem + lane_off + C::TM_Q + C::Q_STRIDE * i;
            uint32_t qa[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u = __ldg(src + c);
                qa[4 * c] = u.x; qa[4 * c + 1] = u.y; qa[4 * c + 2] = u.z; qa[4 * c + 3] = u.w;
            }
            ptx::tmem_st_x32(q_tmem, qa);
            if (HD > 64) {
                uint32_t qb[8];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint4 u = __ldg(src + 8 + c);
                    qb[4 * c] = u.x; qb[4 * c + 1] = u.y; qb[4 * c + 2] = u.z; qb[4 * c + 3] = u.w;
                }
                ptx::tmem_st_x8(q_tmem + 32, qb);
            }
            ptx::tmem_st_wait();
        }
        // ---- rel-pos prologue: w term into registers, h term into this warp's smem column block ----
        float bwl[64];
        ptx::mbar_wait(&bars[C::B_BIAS + i], 0);
        ptx::mbar_wait(&bars[C::B_BIAS + (i ^ 1)], 0);             // the staging area below aliases the tables BOTH tiles' bias MMAs read
        ptx::tc_fence_after();
#pragma unroll
        for (int p = 0; p < 2; ++p) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[32];
                ptx::tmem_ld_x32(s_tmem + 64 * p + 32 * hh, v);
                ptx::tmem_ld_wait_dep(v);
#pragma unroll
                for (int e = 0; e < 32; ++e) stg[(32 * hh + e) * 32 + lane] = __uint_as_float(v[e]) * LOG2E;
            }
            __syncwarp();
#pragma unroll
            for (int kw = 0; kw < 64; ++kw) {
                const int c = qw + 63 - kw;                        // table row qw - kw + 63
                if ((c >> 6) == p) bwl[kw] = stg[(c & 63) * 32 + lane];
            }
            __syncwarp();
        }
        {
            // column c of the h product = Q . Rh[4*pair + c]; key row kh needs table row (4*pair + qr) - kh + 63
            uint32_t v[32];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                ptx::tmem_ld_x32(o_tmem + c0, v);
                ptx::tmem_ld_wait_dep(v);
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int kh = qr + 63 - (c0 + e);
                    if (kh >= 0 && kh < 64) stg[kh * 32 + lane] = __uint_as_float(v[e]) * LOG2E;
                }
            }
            uint32_t w[16];
            ptx::tmem_ld_x16(o_tmem + 64, w);
            ptx::tmem_ld_wait_dep(w);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int kh = qr + 63 - (64 + e);
                if (kh >= 0 && kh < 64) stg[kh * 32 + lane] = __uint_as_float(w[e]) * LOG2E;
            }
        }
        __syncwarp();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[C::B_BREAD + i]);                   // S_i / O_i columns may be overwritten

        float m_ref = -INFINITY;
        f32x2 l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);      // row sum, four partial accumulators
        const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
        long long pc[5] = {0, 0, 0, 0, 0};
        long long tprev = PH ? clock64() : 0;
#define SVB_GPH(k) if (PH) { const long long tn = clock64(); pc[k] += tn - tprev; tprev = tn; }
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            const int hb = h & 1;
            const uint32_t s_h = s_tmem + 64 * hb;
            // S(h) sits in the other buffer than P(h-1) and is normally already there: its loads are started BEFORE the hand-over
            // of P(h-1), so their latency overlaps the store drain (loads and their wait stay in straight-line code: the
            // destination registers of an in-flight tcgen05.ld must not cross a loop edge, where the compiler may copy them)
            uint32_t va[32], vb[32];
            ptx::mbar_wait(&bars[C::B_SFULL + 2 * i + hb], (h >> 1) & 1);
            SVB_GPH(0)
            ptx::tc_fence_after();
            ptx::tmem_ld_x32(s_h, va);
            ptx::tmem_ld_x32(s_h + 32, vb);
            // (deferring this hand-over into the tile's arithmetic, where the stores have long drained, was measured SLOWER — 912 ->
            // 1083 us per 8 images: the chain PV(h-1) -> QK(h+1) needs the whole tile of slack)
            if (h > 0) {
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&bars[C::B_PFULL + 2 * i + (hb ^ 1)]);
            }
            SVB_GPH(3)
            const float bh = stg[h * 32 + lane];
            float m_new;
            bool need;
            ptx::tmem_ld_wait_dep(va);
            ptx::tmem_ld_wait_dep(vb);
            SVB_GPH(4)
            // SAFE single pass over tensor memory: the 64 scores are in registers, so the tile's EXACT maximum exponent is known
            // before the first exponential — pass 1 forms x = s * scale + w term in place and takes its maximum, pass 2 the
            // exponentials.  The reference maximum moves to it whenever it exceeds the reference by more than 2^8, so P <= 2^8
            // ALWAYS and no input can overflow exp2, the row sum or O (torch.softmax has no such limit either,
            // image_encoder.py:246-252); and because the reference is a maximum that was really attained (not a bound built from
            // the largest bias term), a row cannot underflow as a whole.  Same instruction count as tracking the maximum alongside.
            f32x2 xa[16], xb[16];
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                xa[e / 2] = f2_fma(f2_pack(__uint_as_float(va[e]), __uint_as_float(va[e + 1])), sc2, f2_pack(bwl[e], bwl[e + 1]));
                xa[e / 2 + 1] = f2_fma(f2_pack(__uint_as_float(va[e + 2]), __uint_as_float(va[e + 3])), sc2, f2_pack(bwl[e + 2], bwl[e + 3]));
                xb[e / 2] = f2_fma(f2_pack(__uint_as_float(vb[e]), __uint_as_float(vb[e + 1])), sc2, f2_pack(bwl[32 + e], bwl[33 + e]));
                xb[e / 2 + 1] = f2_fma(f2_pack(__uint_as_float(vb[e + 2]), __uint_as_float(vb[e + 3])), sc2, f2_pack(bwl[34 + e], bwl[35 + e]));
                float a0, a1, a2, a3, b0, b1, b2, b3;
                f2_unpack(xa[e / 2], a0, a1); f2_unpack(xa[e / 2 + 1], a2, a3);
                f2_unpack(xb[e / 2], b0, b1); f2_unpack(xb[e / 2 + 1], b2, b3);
                mx0 = fmax3(mx0, a0, a1); mx1 = fmax3(mx1, a2, a3);
                mx2 = fmax3(mx2, b0, b1); mx3 = fmax3(mx3, b2, b3);
            }
            {
                const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) + bh;     // exact maximum exponent of this tile, log2 units
                need = (h == 0) || (tmax - m_ref > RESCALE_THRESHOLD);
                m_new = need ? tmax : m_ref;
            }
            if (__any_sync(0xffffffffu, need)) {
                if (h > 0) {
                    const float alpha = need ? ptx::ex2_approx(m_ref - m_new) : 1.0f;
                    ptx::mbar_wait(&bars[C::B_PVDONE + i], (h - 1) & 1);   // O_i holds tiles 0..h-1
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
                SVB_GPH(2)
            }
            const f32x2 dd2 = f2_pack(bh - m_ref, bh - m_ref);
            // ---- P = exp2(x + h term - m_ref) -> bf16 -> TMEM (over the consumed S columns) ----
#define SVB_PASS_H(X, CHUNK)                                                                  \
            {                                                                                            \
                uint32_t pk[16];                                                                         \
                _Pragma("unroll") for (int e = 0; e < 32; e += 4) {                                      \
                    const f32x2 x01 = f2_add(X[e / 2], dd2);                                             \
                    const f32x2 x23 = f2_add(X[e / 2 + 1], dd2);                                         \
                    float a0, a1, a2, a3;                                                                \
                    f2_unpack(x01, a0, a1);                                                              \
                    f2_unpack(x23, a2, a3);                                                              \
                    float p0, p1, p2, p3;                                                                \
                    if (poly_pair(e / 2, POLY)) exp2_poly_pair(a0, a1, p0, p1);   /* POLY of every 8 pairs on the FMA pipe */ \
                    else { p0 = ptx::ex2_approx(a0); p1 = ptx::ex2_approx(a1); }                         \
                    if (poly_pair(e / 2 + 1, POLY)) exp2_poly_pair(a2, a3, p2, p3);                      \
                    else { p2 = ptx::ex2_approx(a2); p3 = ptx::ex2_approx(a3); }                         \
                    l01 = f2_add(l01, f2_pack(p0, p1));                                                  \
                    l23 = f2_add(l23, f2_pack(p2, p3));                                                  \
                    pk[e / 2] = pack_bf16x2(p0, p1);                                                     \
                    pk[e / 2 + 1] = pack_bf16x2(p2, p3);                                                 \
                }                                                                                        \
                ptx::tmem_st_x16(s_h + 16 * (CHUNK), pk);                                                \
            }
            SVB_PASS_H(xa, 0)
            SVB_PASS_H(xb, 1)
#undef SVB_PASS_H
            SVB_GPH(1)
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[C::B_PFULL + 2 * i + 1]);             // P(NH-1)
        SVB_GPH(3)
#undef SVB_GPH
        if (PH && phase_clocks && w4 == 0 && lane == 0) {
            for (int k = 0; k < 3; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + i * 8 + k, (unsigned long long)pc[k]);
            atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + i * 8 + 6, (unsigned long long)(NH / 2));
            if (i == 0) { atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + 7, (unsigned long long)pc[3]); atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + 15, (unsigned long long)pc[4]); }
        }
        // ---- epilogue: O / l at the query's own token position (its own barrier: the per-tile PVDONE phases may be up to two
        // ahead of this group here, which a parity wait cannot tell apart) ----
        ptx::mbar_wait(&bars[C::B_ODONE + i], 0);
        ptx::tc_fence_after();
        float l0, l1, l2, l3;
        f2_unpack(l01, l0, l1);
        f2_unpack(l23, l2, l3);
        const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
        bf16* dst = out + (size_t)(row0 + 128 * i + t) * D + head * HD;
        store_row<HD>(dst, o_tmem, inv);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}

// ================================================================================================================
//                     WINDOWED ATTENTION, persistent variant (the production kernel for 14 x 14 windows)
// ================================================================================================================
// One CTA per SM loops over (image, window, head) items.
//   * The item's Q / K / V window boxes are prefetched by TMA into a 2-stage ring while the previous item computes.
//   * Two query tiles per window — window rows 0..8 (126 queries) and 9..13 (70 queries) — each with its own softmax warp group
//     AND its own MMA issuer thread, so the two chains never wait on each other.
//   * The rel-pos skew (row t needs entries yi+13-ki / xi+13-kk of ITS OWN 27-entry products) is a register barrel shift.
//   * The output tile is staged in the (dead) Q buffer in the TMA layout and written with ONE tensor store per tile
//     (cp.async.bulk.tensor, box = 14 x 9|5 tokens): window_unpartition and the crop to 64 x 64 are the store's coordinates and
//     its out-of-bounds clipping.  (Row-per-thread stores were LSU-bound: 3000 of 12400 cycles per item.)
//   * Per group the loop is software-pipelined: after PV(k) the group first hands the rel-pos terms of item k+1 to its issuer,
//     so S(k+1) runs on the tensor core while the group stores O(k).
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
    static constexpr int QK_TX = (126 + 70 + 196) * ROWB;
    static constexpr int V_TX = 196 * ROWB;
    static constexpr int R_TX = 64 * ROWB;
    static constexpr int B_RFULL = 0, B_QKFULL = 1, B_VFULL = 3, B_EMPTY = 5, B_BIAS = 7, B_BREAD = 9, B_SFULL = 11, B_PFULL = 13,
                         B_PVDONE = 15, B_OSTAGED = 17, B_COUNT = 21;   // OSTAGED: one per (tile, stage)
    // V tile at head_dim 80: FIVE 16-column atoms of the 32B swizzle, V_ATOM bytes apart (an MN-major operand's leading-dimension
    // byte offset), so that P.V is ONE N = 80 MMA per K step (a tcgen05.mma with A in TMEM costs >= 44.5 cycles whatever N is:
    // the 64 + 16 pair cost 89).  At head_dim 64: one 64-column atom of the 128B swizzle.
    static constexpr int V_ATOM = 208 * 32;
    static constexpr int TM_COLS = 512;                                       // S_i at 208*i; O_i inside S_i at +112
};

struct WinPMaps {
    CUtensorMap q0, q1, kv, r;           // loads: boxes (64,14,9,1) / (64,14,5,1) / (64,14,14,1) of the padded qkv; (64,64) of the table
    CUtensorMap q0t, q1t, kvt, rt;       // their 16-column tails (32B swizzle)
    CUtensorMap o0, o1, o0t, o1t;        // stores: boxes (64|16,14,9|5,1) of out viewed as [B,64,64,D]
};

template <int HD, int POLY, bool PH, int KO = 0>
__global__ void __launch_bounds__(384, 1)
attn_window_persistent_kernel(const __grid_constant__ WinPMaps maps, int D, int g, int nwy, int nwx, int heads, int num_items, float scale_log2,
                              long long* __restrict__ phase_clocks, int l2_ahead) {
    // g = token-grid HEIGHT (the last window row's padding decides whether query tile 1 exists); nwy x nwx windows per image
    using C = WPCfg<HD>;
    constexpr int WS = 14;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::B_COUNT);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // provably warp-uniform
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // (see attn_global2_kernel: the proj GEMM is a dependent launch)

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&maps.kv);
        ptx::prefetch_tmap(&maps.q0);
        ptx::prefetch_tmap(&maps.q1);
        ptx::prefetch_tmap(&maps.r);
        ptx::prefetch_tmap(&maps.o0);
        ptx::prefetch_tmap(&maps.o1);
        for (int s = 0; s < C::B_COUNT; ++s) {
            const bool by_threads = (s >= C::B_BREAD && s < C::B_BREAD + 2) || (s >= C::B_PFULL && s < C::B_PFULL + 2) ||
                                    (s >= C::B_OSTAGED && s < C::B_OSTAGED + 4);
            const bool by_tiles = (s >= C::B_EMPTY && s < C::B_EMPTY + 2);            // one release per query tile
            ptx::mbar_init(&bars[s], by_threads ? 128 : (by_tiles ? 2 : 1));
        }
        ptx::fence_barrier_init();
    }
    if (warp == 9) ptx::tmem_alloc(tmem_slot, C::TM_COLS);
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
    asm volatile("griddepcontrol.wait;" ::: "memory");          // dependent launch: the set-up above overlapped the previous kernel's tail

    // item -> (image, window, head); query tile 1 (window rows 9..13) is entirely padding in the last window row
    auto decode = [&](int item, int& b, int& wy, int& wx, int& head) {
        if constexpr (KO & 512) item %= heads;                     // diagnostic: every item is window 0 of image 0 (L2-resident operands)
        head = item % heads;
        const int bw = item / heads;
        const int win = bw % (nwy * nwx);
        b = bw / (nwy * nwx);
        wy = win / nwx;
        wx = win % nwx;
    };

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            ptx::mbar_expect_tx(&bars[C::B_RFULL], C::R_TX);
            ptx::tma_load_2d(sm + C::OFF_R, &maps.r, &bars[C::B_RFULL], 0, 0);
            if (HD > 64) ptx::tma_load_2d(sm + C::OFF_R + C::R_MAIN, &maps.rt, &bars[C::B_RFULL], 64, 0);
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t par = ((it >> 1) & 1) ^ 1;
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                const int x0 = wx * WS, y0 = wy * WS;
                uint8_t* q0 = sm + st * C::STAGE;
                uint8_t* q1 = q0 + C::QT;
                uint8_t* k = sm + st * C::STAGE + C::OFF_K;
                uint8_t* v = sm + st * C::STAGE + C::OFF_V;
                const int cq = head * HD, ck = D + head * HD, cv = 2 * D + head * HD;
                ptx::mbar_wait(&bars[C::B_EMPTY + st], par);
                ptx::mbar_expect_tx(&bars[C::B_QKFULL + st], C::QK_TX);
                ptx::tma_load_4d(q0, &maps.q0, &bars[C::B_QKFULL + st], cq, x0, y0, b);
                ptx::tma_load_4d(q1, &maps.q1, &bars[C::B_QKFULL + st], cq, x0, y0 + 9, b);
                ptx::tma_load_4d(k, &maps.kv, &bars[C::B_QKFULL + st], ck, x0, y0, b);
                if (HD > 64) {
                    ptx::tma_load_4d(q0 + C::Q_MAIN, &maps.q0t, &bars[C::B_QKFULL + st], cq + 64, x0, y0, b);
                    ptx::tma_load_4d(q1 + C::Q_MAIN, &maps.q1t, &bars[C::B_QKFULL + st], cq + 64, x0, y0 + 9, b);
                    ptx::tma_load_4d(k + C::K_MAIN, &maps.kvt, &bars[C::B_QKFULL + st], ck + 64, x0, y0, b);
                }
                ptx::mbar_expect_tx(&bars[C::B_VFULL + st], C::V_TX);
                if (HD > 64) {
                    for (int a = 0; a < 5; ++a) ptx::tma_load_4d(v + a * C::V_ATOM, &maps.kvt, &bars[C::B_VFULL + st], cv + 16 * a, x0, y0, b);
                } else {
                    ptx::tma_load_4d(v, &maps.kv, &bars[C::B_VFULL + st], cv, x0, y0, b);
                }
                // the kernel moves 48 MB per image and block through a 2-stage ring whose stages free up late: keep HBM -> L2
                // traffic in flight for the items after the one in the ring (the loads above then hit L2)
                for (int ahead = 1; ahead <= l2_ahead; ++ahead) {
                    const int pi = item + (ahead + 0) * gridDim.x;
                    if (pi < num_items) {
                        int pb, pwy, pwx, ph_;
                        decode(pi, pb, pwy, pwx, ph_);
                        const int px0 = pwx * WS, py0 = pwy * WS;
                        for (int part = 0; part < 3; ++part) {
                            const int cc = part * D + ph_ * HD;
                            ptx::tma_prefetch_l2_4d(&maps.kv, cc, px0, py0, pb);
                            if (HD > 64) ptx::tma_prefetch_l2_4d(&maps.kvt, cc + 64, px0, py0, pb);
                        }
                    }
                }
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ===================== MMA issuers: one per query tile =====================
        // all 32 lanes run the loop, one elected lane issues (uniform-register operands: see ptx::mma_f16_ss_e)
        {
            constexpr uint32_t id_r = ptx::make_idesc_bf16(128, 64, 0, 0);
            constexpr uint32_t id_s = ptx::make_idesc_bf16(128, 208, 0, 0);
            const int i = warp - 9;
            ptx::mbar_wait(&bars[C::B_RFULL], 0);
            int it = 0;
            uint32_t n = 0;                                        // active items of this tile so far (barrier phases)
#pragma unroll 1
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                if (i == 1 && wy * WS + 9 >= g) {
                    // tile 1 is all padding here: just release the stage — but only once THIS use of the stage has begun (its
                    // loads landed), otherwise the arrival would be counted in the previous use's phase
                    ptx::mbar_wait(&bars[C::B_QKFULL + st], ph);
                    ptx::mbar_arrive_e(&bars[C::B_EMPTY + st]);
                    continue;
                }
                const uint32_t q = base + st * C::STAGE + i * C::QT, k = base + st * C::STAGE + C::OFF_K, v = base + st * C::STAGE + C::OFF_V;
                ptx::mbar_wait(&bars[C::B_QKFULL + st], ph);
                ptx::tc_fence_after();
                // columns [0,64) of S_i: the previous item's P_i there was consumed by its PV (same issuer, in-order tensor pipe);
                // its O_i (columns 112..191) is only overwritten by the S MMA below, issued after the group has loaded O_i
                if constexpr (!(KO & 256)) issue_qk<HD>(tmem + 208 * i, q, q + C::Q_MAIN, base + C::OFF_R, base + C::OFF_R + C::R_MAIN, id_r);
                ptx::mma_commit_e(&bars[C::B_BIAS + i]);
                ptx::mbar_wait(&bars[C::B_BREAD + i], n & 1);      // rel-pos products consumed (and the previous O_i loaded)
                ptx::tc_fence_after();
                if constexpr (!(KO & 128)) issue_qk<HD>(tmem + 208 * i, q, q + C::Q_MAIN, k, k + C::K_MAIN, id_s);
                ptx::mma_commit_e(&bars[C::B_SFULL + i]);
                ptx::mbar_wait(&bars[C::B_VFULL + st], ph);
                ptx::mbar_wait(&bars[C::B_PFULL + i], n & 1);      // P_i is in TMEM
                ptx::tc_fence_after();
                if constexpr (KO & 64) {
                } else if (HD > 64) {
                    constexpr uint32_t id_pv = ptx::make_idesc_bf16(128, HD, 0, 1);
                    const uint64_t dv = ptx::make_smem_desc(v, C::V_ATOM, 256, ptx::LAYOUT_SW32);
#pragma unroll
                    for (int kk = 0; kk < 13; ++kk)                // 16 keys = 512 B inside an atom
                        ptx::mma_f16_ts_e(tmem + 208 * i + 112, tmem + 208 * i + 8 * kk, dv + 32 * kk, id_pv, kk ? 1u : 0u);
                } else {
                    issue_pv<HD>(tmem + 208 * i + 112, tmem + 208 * i, v, v + C::K_MAIN, 13, false);
                }
                ptx::mma_commit_e(&bars[C::B_PVDONE + i]);
                ++n;
            }
        }
    } else if (warp == 11) {
        // ===================== store warp: O tiles staged by the groups -> ONE tensor store per tile =====================
        // (the wait for the store to finish reading the stage sat on the critical path of the group's first warp)
        if (lane == 0) {
            // one "staged" barrier per (tile, stage): a group may stage the NEXT item (other stage) before this warp has looked at
            // the previous one, and a single barrier two phases ahead cannot be told from one that has not moved (parity wait)
            int it = 0;
            uint32_t n[4] = {0, 0, 0, 0};
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int st = it & 1;
                int b, wy, wx, head;
                decode(item, b, wy, wx, head);
                for (int i = 0; i < 2; ++i) {
                    if (i == 1 && wy * WS + 9 >= g) continue;      // released by the issuer
                    const uint8_t* ob = sm + st * C::STAGE + i * C::QT;
                    ptx::mbar_wait(&bars[C::B_OSTAGED + 2 * i + st], n[2 * i + st] & 1);
                    ++n[2 * i + st];
                    const int x0 = wx * WS, y0 = wy * WS + 9 * i;
                    tma_store_4d(i ? &maps.o1 : &maps.o0, ob, head * HD, x0, y0, b);   // rows / columns past 64 are clipped by the TMA
                    if (HD > 64) tma_store_4d(i ? &maps.o1t : &maps.o0t, ob + C::Q_MAIN, head * HD + 64, x0, y0, b);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the store has finished reading the stage
                    ptx::mbar_arrive(&bars[C::B_EMPTY + st]);      // this tile is done with stage st (its MMAs retired before PVDONE)
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");                  // stores complete before the CTA exits
        }
    } else {
        // ===================== softmax warps: group i = query tile i =====================
        const int i = warp >> 2, w4 = warp & 3;
        const int t = w4 * 32 + lane;                              // query row inside the tile
        const int yi = 9 * i + t / WS, xi = t % WS;                // window coordinates (rows past the tile are discarded)
        const uint32_t lane_off = static_cast<uint32_t>(w4 * 32) << 16;
        const uint32_t s_tmem = tmem + lane_off + 208 * i;
        const uint32_t o_tmem = s_tmem + 112;
        long long pc[6] = {0, 0, 0, 0, 0, 0};
        long long tprev = (PH && phase_clocks) ? clock64() : 0;
#define SVB_PHASE(k) if (PH && phase_clocks) { const long long tnow = clock64(); pc[k] += tnow - tprev; tprev = tnow; }

        // rel-pos products of the item whose bias MMA is `ph`-th for this tile -> the row's 14 + 14 terms (log2 units)
        float bhm[14], bwl[14];
        auto read_bias = [&](uint32_t ph) {
            ptx::mbar_wait(&bars[C::B_BIAS + i], ph);
            ptx::tc_fence_after();
            uint32_t v[32], v2[32];
            float rr[27];
            ptx::tmem_ld_x32(s_tmem, v);                           // both loads in flight: one tensor-memory round trip
            ptx::tmem_ld_x32(s_tmem + 32, v2);
            ptx::tmem_ld_wait_dep(v);
            ptx::tmem_ld_wait_dep(v2);
#pragma unroll
            for (int j = 0; j < 27; ++j) rr[j] = __uint_as_float(v[j]) * LOG2E;
            if constexpr (!(KO & 4)) barrel_shift27(rr, yi < 13 ? yi : 13);
#pragma unroll
            for (int kk = 0; kk < 14; ++kk) bhm[kk] = rr[13 - kk];
#pragma unroll
            for (int j = 0; j < 27; ++j) rr[j] = __uint_as_float(v2[j]) * LOG2E;
            if constexpr (!(KO & 4)) barrel_shift27(rr, xi);
#pragma unroll
            for (int kk = 0; kk < 14; ++kk) bwl[kk] = rr[13 - kk];
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars[C::B_BREAD + i]);
        };
        auto active = [&](int item) {
            int b, wy, wx, head;
            decode(item, b, wy, wx, head);
            return !(i == 1 && wy * WS + 9 >= g);
        };
        // first active item of this group
        int item = blockIdx.x, it = 0;
        while (item < num_items && !active(item)) { item += gridDim.x; ++it; }
        uint32_t n = 0;
        if (item < num_items) read_bias(0);
        SVB_PHASE(1)
        while (item < num_items) {
            int b, wy, wx, head;
            decode(item, b, wy, wx, head);
            const int st = it & 1;
            const uint32_t ph = n & 1;
            // ---- softmax over the 196 keys ----
            ptx::mbar_wait(&bars[C::B_SFULL + i], ph);
            SVB_PHASE(2)
            ptx::tc_fence_after();
            float lsum = 1.f;
            if constexpr (!(KO & 32)) lsum = window_softmax_tile<POLY, KO>(s_tmem, bhm, bwl, scale_log2);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars[C::B_PFULL + i]);
            SVB_PHASE(3)
            // next active item of this group
            int nitem = item + gridDim.x, nit = it + 1;
            while (nitem < num_items && !active(nitem)) { nitem += gridDim.x; ++nit; }
            // ---- O(k): TMEM -> normalised bf16 in registers ----
            ptx::mbar_wait(&bars[C::B_PVDONE + i], ph);
            SVB_PHASE(4)
            ptx::tc_fence_after();
            const float inv = 1.0f / lsum;
            uint32_t o[HD / 2];
            if constexpr (KO & 16) {
#pragma unroll
                for (int j = 0; j < HD / 2; ++j) o[j] = __float_as_uint(inv);
            } else {
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    ptx::tmem_ld_x32(o_tmem + c, v);
                    ptx::tmem_ld_wait_dep(v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[c / 2 + j] = pack_bf16x2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                }
                if (HD > 64) {
                    uint32_t w[16];
                    ptx::tmem_ld_x16(o_tmem + 64, w);
                    ptx::tmem_ld_wait_dep(w);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[32 + j] = pack_bf16x2(__uint_as_float(w[2 * j]) * inv, __uint_as_float(w[2 * j + 1]) * inv);
                }
            }
            // ---- hand the next item's rel-pos terms over first: S(k+1) then runs while O(k) is being stored.  Only when the
            // next item of this tile sits in the OTHER stage: if the tile skips an item, its next one reuses THIS stage, whose
            // reload waits for the release below (reading its bias first would deadlock). ----
            const bool early = (nitem < num_items) && (nit == it + 1);
            if (early) read_bias((n + 1) & 1);
            SVB_PHASE(1)
            // ---- O(k) -> the dead Q buffer of this stage in the TMA layout -> one tensor store per tile ----
            {
                uint8_t* ob = sm + st * C::STAGE + i * C::QT;
                if constexpr (!(KO & 16)) {
#pragma unroll
                for (int j = 0; j < 8; ++j)                        // 128B swizzle: 16-byte piece j of row t at piece j ^ (t & 7)
                    *reinterpret_cast<uint4*>(ob + t * 128 + ((j ^ (t & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                if (HD > 64) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)                    // 32B swizzle: piece j of row t at piece j ^ ((t >> 2) & 1)
                        *reinterpret_cast<uint4*>(ob + C::Q_MAIN + t * 32 + ((j ^ ((t >> 2) & 1)) << 4)) =
                            make_uint4(o[32 + 4 * j], o[32 + 4 * j + 1], o[32 + 4 * j + 2], o[32 + 4 * j + 3]);
                }
                }
                ptx::fence_proxy_async_smem();                     // generic writes -> visible to the TMA (async proxy) read
                ptx::mbar_arrive(&bars[C::B_OSTAGED + 2 * i + st]); // the store warp issues the tensor store and releases the stage
            }
            if (!early && nitem < num_items) read_bias((n + 1) & 1);
            SVB_PHASE(5)
            item = nitem;
            it = nit;
            ++n;
        }
#undef SVB_PHASE
        if (PH && phase_clocks && w4 == 0 && lane == 0) {
            for (int k = 0; k < 6; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + i * 8 + k, (unsigned long long)pc[k]);
            atomicAdd(reinterpret_cast<unsigned long long*>(phase_clocks) + i * 8 + 6, (unsigned long long)n);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::TM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pad rows of the padded [B, gp, gp, ld] qkv tensor = the qkv bias (pad tokens are zero AFTER norm1, image_encoder.py:183-187,
// 271-275, so their k / v are b_k / b_v); written once per windowed block before the attention kernel reads it.
__global__ void __launch_bounds__(256)
fill_pad_rows_kernel(bf16* __restrict__ qkv, const float* __restrict__ bias, int B, int gh, int gw, int gph, int gpw, int ld) {
    // the bias row is converted to bf16 ONCE per block into shared memory; then one warp per pad row copies it out with 16-byte
    // stores (measured before: every row re-read and re-converted the fp32 bias, 23 us per launch at 1.8 TB/s for 42 MB)
    extern __shared__ uint4 brow[];                                  // ld / 8 pieces of 8 bf16
    const int v8 = ld / 8;
    for (int c8 = threadIdx.x; c8 < v8; c8 += blockDim.x) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
        uint4 u;
        u.x = pack_bf16x2(b0.x, b0.y); u.y = pack_bf16x2(b0.z, b0.w);
        u.z = pack_bf16x2(b1.x, b1.y); u.w = pack_bf16x2(b1.z, b1.w);
        brow[c8] = u;
    }
    __syncthreads();
    // dependent launch (the bias row above overlapped the qkv GEMM's tail); the attention kernel behind this one is one as well
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int npad = gph * gpw - gh * gw, strip = gh * (gpw - gw);
    const int rows = B * npad;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
        const int pi = r % npad, b = r / npad;
        // pad index -> (y, x): first the right-hand strip of the gh real rows, then the full bottom rows
        int y, x;
        if (pi < strip) { y = pi / (gpw - gw); x = gw + pi % (gpw - gw); }
        else { y = gh + (pi - strip) / gpw; x = (pi - strip) % gpw; }
        uint4* orow = reinterpret_cast<uint4*>(qkv + (((size_t)b * gph + y) * gpw + x) * ld);
        for (int c8 = lane; c8 < v8; c8 += 32) orow[c8] = brow[c8];
    }
}

// rel_pos table (L, hd) fp32 rows [src0, src0 + n) -> rows [row_off, row_off + n) of the packed bf16 table
__global__ void pack_rel_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int n_rows, int hd, int src0, int row_off) {
    const int n = n_rows * hd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[(size_t)row_off * hd + i] = __float2bfloat16_rn(src[(size_t)src0 * hd + i]);
}

// A quarter of the softmax exponentials on the FMA pipe (exp2_poly_pair).  Measured: windowed kernel 118.5 -> 114.5 us per 8 images
// (its two softmax groups are MUFU-bound while they overlap), global kernel 920 -> 947 us (issue / latency-bound: the extra
// instructions cost more than the MUFU slots they free).  Defaults: on for the windowed kernel (SVB_ATTNW_POLY=0 turns it off),
// off for the global kernel (SVB_ATTNG_POLY=1 turns it on).
static int exp2_poly(bool windowed) {
    static const int w = [] { const char* e = getenv("SVB_ATTNW_POLY"); return e ? atoi(e) : 2; }();
    static const int g = [] { const char* e = getenv("SVB_ATTNG_POLY"); return e ? atoi(e) : 2; }();
    return windowed ? w : g;
}

template <int HD, int NST>
int launch_global_nst(const AttnTcParams& p, cudaStream_t stream) {
    using C = G2Cfg<HD, NST>;
    const int D = p.heads * p.hd, T = p.grid * p.grid;
    CUtensorMap m[6];
    int rc;
    {
        const uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)p.batch * T};
        const uint64_t str[1] = {(uint64_t)3 * D * 2};
        const uint32_t bm[2] = {64, 128}, bt[2] = {16, 128};
        if ((rc = encode_tmap_nd_bf16(&m[0], p.qkv, 2, dims, str, bm, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&m[1], p.qkv, 2, dims, str, bt, 32))) return rc;
        const uint64_t rd[2] = {(uint64_t)HD, 272};
        const uint64_t rs[1] = {(uint64_t)HD * 2};
        const uint32_t wm[2] = {64, 128}, wt[2] = {16, 128}, hm[2] = {64, 80}, ht[2] = {16, 80};
        if ((rc = encode_tmap_nd_bf16(&m[2], p.rel_pack, 2, rd, rs, wm, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&m[3], p.rel_pack, 2, rd, rs, wt, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&m[4], p.rel_pack, 2, rd, rs, hm, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&m[5], p.rel_pack, 2, rd, rs, ht, 32))) return rc;
    }
    const float scale_log2 = LOG2E / sqrtf((float)HD);
    dim3 grid(T / 256, p.heads, p.batch);
    auto launch = [&](auto kern, long long* clocks) -> int {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled(2) ? 1 : 0;
        SVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], (const bf16*)p.qkv, p.out, D, T, scale_log2, clocks, 1));
        return 0;
    };
    const int k8 = exp2_poly(false);
    if (p.phase_clocks) rc = launch(attn_global2_kernel<HD, NST, true, 2>, p.phase_clocks);
    else if (k8 == 2) rc = launch(attn_global2_kernel<HD, NST, false, 2>, nullptr);
    else if (k8 == 3) rc = launch(attn_global2_kernel<HD, NST, false, 3>, nullptr);
    else if (k8 == 4) rc = launch(attn_global2_kernel<HD, NST, false, 4>, nullptr);
    else rc = launch(attn_global2_kernel<HD, NST, false, 0>, nullptr);
    if (rc) return rc;
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int HD>
int launch_global(const AttnTcParams& p, cudaStream_t stream) {
    // K/V ring depth: 3 stages at head_dim 64, 2 at 80 (V tiles of two full atoms; 2 vs 3 stages measured identical)
    return launch_global_nst<HD, (HD > 64) ? 2 : 3>(p, stream);
}

template <int HD>
int launch_window_persistent(const AttnTcParams& p, cudaStream_t stream) {
    using C = WPCfg<HD>;
    const int D = p.heads * p.hd;
    const int gh = p.grid_h ? p.grid_h : p.grid, gw = p.grid_w ? p.grid_w : p.grid;
    const int nwy = (gh + 13) / 14, nwx = (gw + 13) / 14, gph = nwy * 14, gpw = nwx * 14;
    WinPMaps wm;
    int rc;
    {
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
        // output viewed as [B, gh, gw, D]: the store of a window tile is a box of 14 x (9|5) tokens, clipped at the grid's edge
        const uint64_t od[4] = {(uint64_t)D, (uint64_t)gw, (uint64_t)gh, (uint64_t)p.batch};
        const uint64_t os[3] = {(uint64_t)D * 2, (uint64_t)gw * D * 2, (uint64_t)gh * gw * D * 2};
        if ((rc = encode_tmap_nd_bf16(&wm.o0, p.out, 4, od, os, q0, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o1, p.out, 4, od, os, q1, 128))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o0t, p.out, 4, od, os, q0t, 32))) return rc;
        if ((rc = encode_tmap_nd_bf16(&wm.o1t, p.out, 4, od, os, q1t, 32))) return rc;
    }
    const float scale_log2 = LOG2E / sqrtf((float)HD);
    const int items = p.batch * nwy * nwx * p.heads;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = items < sms ? items : sms;
    // the phase-clock instrumentation is its own instantiation: dormant run-time branches in the hot loop are not free (the same
    // lesson as the global kernel's removed ping-pong option)
    auto launch = [&](auto kern, long long* clocks) -> int {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        static const int l2_ahead = [] { const char* e = getenv("SVB_ATTNW_L2AHEAD"); return e ? atoi(e) : 1; }();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled(2) ? 1 : 0;
        SVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, wm, D, gh, nwy, nwx, p.heads, items, scale_log2, clocks, l2_ahead));
        return 0;
    };
    const int k8 = exp2_poly(true);
#ifdef SVB_ATTN_KO
    static const int ko = [] { const char* e = getenv("SVB_ATTNW_KO"); return e ? atoi(e) : 0; }();
    if (ko && HD == 80) {
        switch (ko) {
#define SVB_KO_CASE(K) case K: rc = launch(attn_window_persistent_kernel<HD, 2, false, K>, nullptr); break;
            SVB_KO_CASE(1) SVB_KO_CASE(2) SVB_KO_CASE(3) SVB_KO_CASE(4) SVB_KO_CASE(8) SVB_KO_CASE(16) SVB_KO_CASE(32) SVB_KO_CASE(64) SVB_KO_CASE(128)
            SVB_KO_CASE(256) SVB_KO_CASE(448) SVB_KO_CASE(36) SVB_KO_CASE(52) SVB_KO_CASE(500) SVB_KO_CASE(11) SVB_KO_CASE(15) SVB_KO_CASE(31)
            SVB_KO_CASE(512) SVB_KO_CASE(513) SVB_KO_CASE(514) SVB_KO_CASE(515) SVB_KO_CASE(516) SVB_KO_CASE(520) SVB_KO_CASE(528) SVB_KO_CASE(544)
            SVB_KO_CASE(576) SVB_KO_CASE(640) SVB_KO_CASE(768) SVB_KO_CASE(960) SVB_KO_CASE(1012) SVB_KO_CASE(523) SVB_KO_CASE(527) SVB_KO_CASE(543)
            SVB_KO_CASE(548) SVB_KO_CASE(564) SVB_KO_CASE(1536) SVB_KO_CASE(2560) SVB_KO_CASE(3584) SVB_KO_CASE(3592) SVB_KO_CASE(1024) SVB_KO_CASE(2048)
            SVB_KO_CASE(3072)
#undef SVB_KO_CASE
            default: SVB_REQUIRE(false, "SVB_ATTNW_KO=%d is not compiled", ko);
        }
        if (rc) return rc;
        SVB_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
#endif
    if (p.phase_clocks) rc = launch(attn_window_persistent_kernel<HD, 2, true>, p.phase_clocks);
    else if (k8 == 0) rc = launch(attn_window_persistent_kernel<HD, 0, false>, nullptr);
    else if (k8 == 3) rc = launch(attn_window_persistent_kernel<HD, 3, false>, nullptr);
    else if (k8 == 4) rc = launch(attn_window_persistent_kernel<HD, 4, false>, nullptr);
    else rc = launch(attn_window_persistent_kernel<HD, 2, false>, nullptr);
    if (rc) return rc;
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// Rows of the packed bf16 rel-pos block of one attention module.  Global (64 x 64): rel_pos_h at rows [0,127), rel_pos_w at
// [144,271).  Windowed (14 x 14): rows [0,27) rel_pos_h and [32,59) rel_pos_w (two-group kernel); rows [64,112) and [112,160) the
// two 48-row tables of the pipelined kernel (attention_win3.cu): query tile 0 (window rows 0..7) = rel_pos_h[0..20] then
// rel_pos_w[0..26], query tile 1 (rows 8..13) = rel_pos_h[8..26] then rel_pos_w[0..26] (+ two unused rows).  Unwritten rows must
// be finite: the block is zero-filled when it is allocated.
int attention_tc_rel_rows(int ws, int grid) { return ws == grid ? 272 : 160; }

// registers a host-mapped buffer (64 x u64, zeroed) that receives mbarrier-timeout records of the attention kernels
int attention_tc_set_debug_buffer(void* mapped_device_ptr) {
    unsigned long long* p = (unsigned long long*)mapped_device_ptr;
    SVB_CHECK_CUDA(cudaMemcpyToSymbol(ptx::svb_dbg_buf, &p, sizeof(p)));
    return 0;
}

int pack_rel_table(const float* src, bf16* dst, int L, int hd, bool is_w, cudaStream_t stream) {
    SVB_REQUIRE(L == 27 || L == 127, "pack_rel_table: table length %d is not 27 (14x14 windows) or 127 (64x64 global)", L);
    auto pack = [&](int n_rows, int src0, int row_off) {
        pack_rel_kernel<<<(n_rows * hd + 255) / 256, 256, 0, stream>>>(src, dst, n_rows, hd, src0, row_off);
        count_launch();
    };
    if (L == 127) {
        pack(L, 0, is_w ? 144 : 0);
    } else if (is_w) {
        pack(27, 0, 32);
        pack(27, 0, 64 + 21);          // tile 0 table: rows 21..47
        pack(27, 0, 112 + 19);         // tile 1 table: rows 19..45
    } else {
        pack(27, 0, 0);
        pack(21, 0, 64);               // tile 0 table: rel_pos_h[0..20]
        pack(19, 8, 112);              // tile 1 table: rel_pos_h[8..26]
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int fill_pad_rows(bf16* qkv, const float* bias, int B, int gh, int gw, int gph, int gpw, int ld, cudaStream_t stream) {
    SVB_REQUIRE(ld % 8 == 0, "fill_pad_rows: row length must be a multiple of 8");
    const long rows = (long)B * (gph * gpw - gh * gw);
    if (rows == 0) return 0;
    const long total = rows * (ld / 8);
    const int blocks = (int)std::min<long>((rows + 7) / 8, 148 * 8);
    ProfScope prof(PC_OTHER, 0.0, (double)total * 16.0, stream);
    SVB_REQUIRE(ld * 2 <= 48 * 1024, "fill_pad_rows: row length %d too large", ld);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = (size_t)ld * 2; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled(4) ? 1 : 0;
    SVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fill_pad_rows_kernel, qkv, bias, B, gh, gw, gph, gpw, ld));
    return 0;
}

int attention_tc(const AttnTcParams& p, cudaStream_t stream) {
    const int gh = p.grid_h ? p.grid_h : p.grid, gw = p.grid_w ? p.grid_w : p.grid;
    const bool native = (gh == 64 && gw == 64);
    SVB_REQUIRE(gh > 0 && gw > 0 && gh % 32 == 0 && gw % 32 == 0, "attention_tc: token grid %d x %d: both sides must be multiples of 32", gh, gw);
    SVB_REQUIRE(p.ws == 14 || (native && p.ws == 64), "attention_tc: window size %d is not 14 (windowed) or the 64 x 64 grid (global; other grids: "
                "attention_global_ext)", p.ws);
    SVB_REQUIRE(p.hd == 64 || p.hd == 80, "attention_tc: head_dim %d is not 64 or 80", p.hd);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(p.rel_pack) & 15) == 0, "attention_tc: pointers must be 16-byte aligned");
    const double D_ = (double)p.heads * p.hd;
    const int S = p.ws * p.ws, nwin = p.ws == 64 ? 1 : ((gh + 13) / 14) * ((gw + 13) / 14);
    ProfScope prof(p.ws == 64 ? PC_ATTN_GLOBAL : PC_ATTN_WIN,
                   (double)p.batch * nwin * (4.0 * S * (double)S * D_ + 2.0 * S * 2.0 * p.ws * D_),
                   (double)p.batch * gh * gw * 4.0 * D_ * 2, stream);
    if (p.ws == 64) return p.hd == 64 ? launch_global<64>(p, stream) : launch_global<80>(p, stream);
#ifdef SVB_EXPERIMENTAL_WIN3
    // experimental builds only (SVB_BUILD_EXPERIMENTAL=1), all measured correct and NOT faster than the two-group kernel below inside the
    // encoder step (csrc/experiments/, profiles/r02_attnw_analysis): SVB_ATTNW_IMPL = 3 helper-group pipeline, 5 two independent chains
    // per SM (SVB_ATTNW_ONEPASS=1: S read from TMEM once), 6 split rows with four softmax warps per scheduler
    static const int impl = [] { const char* e = getenv("SVB_ATTNW_IMPL"); return e ? atoi(e) : 2; }();
    if (impl == 3) return attention_window3(p, stream);
    if (impl == 5 && !p.phase_clocks) return attention_window5(p, stream);
    if (impl == 6 && !p.phase_clocks) return attention_window6(p, stream);
#endif
    return p.hd == 64 ? launch_window_persistent<64>(p, stream) : launch_window_persistent<80>(p, stream);
}

}  // namespace svb
