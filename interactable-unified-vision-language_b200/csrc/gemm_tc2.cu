// CTA-pair tcgen05 GEMM (cta_group::2) fed by TMA:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue), bf16 in, fp32 accumulate.
//
// The production GEMM of the encoder (patch embedding, qkv / proj, lin1 / lin2, SimpleFPN convolutions — image_encoder.py:
// 402-410, 227-228, common.py:21-22, image_encoder.py:417-447).  Two CTAs on the two SMs of a TPC form a cluster and work on
// one 256 x BN output tile: each CTA stages ITS 128 rows of A and ITS half of the BN weight rows, the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory and writes 128 accumulator rows into each CTA's
// tensor memory.  Versus the single-CTA kernel (gemm_tc.cu) this halves the weight-tile traffic per SM (64 instead of 96
// bytes per SM clock at full rate) and leaves room for a 6-deep ring of 32 KB stages, which is what keeps the tensor pipe
// fed through L2 latency spikes (ncu of the single-CTA kernel: tensor pipe 70 % active, L2->SM at 54 % of peak).
//
//   warp 0     : TMA producer (each CTA loads its own halves; completion bytes of BOTH CTAs land on the leader's barrier)
//   warp 1     : MMA issuer (leader CTA only); tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs
//   warps 2..9 : epilogue — two warps per TMEM lane quadrant, each owning half of the tile's columns; tcgen05.ld 32x32b
//                (one accumulator row per thread), bias / GELU / GroupNorm statistics in registers, then a per-warp
//                32 x 32 transpose through swizzled shared memory so that every global access of the warp (residual
//                loads, output stores) covers whole 128-byte lines of 4 (fp32) or 8 (bf16) rows instead of one 16-byte
//                piece of 32 different rows (the row-per-thread pattern cost 32 LSU wavefronts per instruction and made
//                the in-place fp32 residual epilogue of proj / lin2 the bottleneck: ncu tensor pipe 32 %).
//   Two accumulator stages in TMEM (2 x BN columns): the epilogue of tile i overlaps the MMAs of tile i + 1.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace svb {
int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
                      uint32_t box_rows, int swizzle_bytes);
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
                 uint32_t box_rows, int swizzle_bytes);
int num_sms();

namespace {

constexpr int BM_CTA = 128;       // rows per CTA; the pair covers 256
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;
constexpr int EPI_WARPS = 8;

template <int BN> struct Cfg2 {
    static constexpr int A_BYTES = BM_CTA * BK * 2;
    static constexpr int B_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 5 : 6;
    static constexpr int BIAS_BYTES = 2 * 2 * BN * 4;             // per accumulator stage: the tile's BN bias values + BN LayerNorm-fold column sums
    static constexpr int EPI_BYTES = EPI_WARPS * 4096;            // per epilogue warp: one 32 x 32 fp32 transpose tile
    static constexpr int OFF_BARS = STAGES * STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_BARS + 256;
    static constexpr int OFF_EPI = OFF_BIAS + BIAS_BYTES + (1024 - (256 + BIAS_BYTES) % 1024) % 1024;
    static constexpr int SMEM_BYTES = OFF_EPI + EPI_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // relaxed: only the TMEM reads must be ordered before this arrive, and tcgen05.wait::ld + fence::before_thread_sync did that;
    // a release at cluster scope would drain the epilogue's global stores (MEMBAR.ALL.GPU + L1 invalidate) once per tile and warp
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of this CTA's tile into its own smem; the completion bytes are credited to the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// Same, multicast: the box lands at the same CTA-relative offset in every CTA of `mask`, and each destination's share of the
// completion bytes is credited to the barrier at the same offset in that destination's pair leader.
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                    uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// warp-uniform issue (all 32 lanes reach the call, one elected lane issues; operands stay in uniform registers — see
// ptx::mma_f16_ss_e): no per-MMA R2UR waterfall loop in the issuer
__device__ __forceinline__ void mma_f16_ss_pair_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred P, q;\n\telect.sync _|P, 0xffffffff;\n\tsetp.ne.b32 q, %4, 0;\n\t"
        "@P tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_pair_e(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\t"
        "@P tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n"
        ::"r"(ptx::smem_u32(bar)), "h"(mask)
        : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// CL = CTAs per cluster: 2 (one pair) or 4 (two pairs on consecutive 256-row blocks of the same BN columns: the weight tile
// is fetched ONCE from L2 and multicast to both pairs, which cuts the L2->SM bytes per MMA by a quarter — the L2 slices'
// ~6300 B/clk were the limit of the single-pair kernel, ncu: lts 10 TB/s, tensor pipe 68 %).
template <int BN, bool RESID, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N, int K,
                Epilogue ep, int dbg) {
    using C = Cfg2<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base_u32 - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::OFF_BARS);
    uint64_t* full_bar = bars;                       // leader's copy is the one in use
    uint64_t* empty_bar = bars + C::STAGES;          // per CTA
    uint64_t* tmem_full = bars + 2 * C::STAGES;      // per CTA
    uint64_t* tmem_empty = bars + 2 * C::STAGES + 2; // leader's copy is the one in use
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    float* bias_s = reinterpret_cast<float*>(tiles + C::OFF_BIAS);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // provably warp-uniform (the issuer's operands stay in uniform registers)
    const int lane = threadIdx.x & 31;
    constexpr int PAIRS = CL / 2;
    const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
    const uint32_t rank = crank & 1;                 // position inside the CTA pair (0 = leader: issues the MMAs)
    const uint32_t pidx = crank >> 1;                // which pair of the cluster
    const uint32_t lead = crank & ~1u;               // cluster rank of this pair's leader
    const int pair = blockIdx.x / CL, num_pairs = gridDim.x / CL;     // cluster index / number of clusters
    const int num_m = (M + 2 * BM_CTA * PAIRS - 1) / (2 * BM_CTA * PAIRS);
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;             // cluster tiles of (256 * PAIRS) x BN
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_w);
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], PAIRS);             // one tcgen05.commit per pair: the stage is free in every CTA
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 2 * EPI_WARPS);   // one arrival per epilogue warp of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    ptx::tc_fence_before();
    cluster_sync_all();                                       // barriers of both CTAs initialised, TMEM allocated
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                const int m0 = ((t / num_n) * PAIRS + pidx) * (2 * BM_CTA) + rank * BM_CTA;
                const int n0 = (t % num_n) * BN + rank * (BN / 2) + pidx * (BN / 2 / PAIRS);
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::A_BYTES;
                    const uint32_t full_leader = map_to_cta(ptx::smem_u32(&full_bar[stage]), lead);
                    if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                    tma_load_2d_pair(sa, &map_a, full_leader, kb * BK, m0);
                    if constexpr (PAIRS == 1) {
                        tma_load_2d_pair(sb, &map_w, full_leader, kb * BK, n0);
                    } else {
                        // this CTA fetches 1/PAIRS of the pair-rank's weight rows and multicasts them to the CTA of the same
                        // pair-rank in every pair
                        uint16_t mask = 0;
#pragma unroll
                        for (int q = 0; q < PAIRS; ++q) mask |= (uint16_t)(1u << (2 * q + rank));
                        tma_load_2d_pair_mc(sb + pidx * (C::B_BYTES / PAIRS), &map_w, full_leader, kb * BK, n0, mask);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (rank == 0) {
            // all lanes run the loop, one elected lane issues (uniform-register operands); SVB_GEMM2_DBG & 32 selects the former
            // single-thread issue (a per-MMA R2UR waterfall loop) for A/B runs
            auto issue_loop = [&](auto uniform) {
            constexpr bool U = decltype(uniform)::value;
            constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM_CTA, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = base_u32 + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    const uint64_t da = ptx::make_smem_desc(sa, 0, 1024, ptx::LAYOUT_SW128);
                    const uint64_t db = ptx::make_smem_desc(sb, 0, 1024, ptx::LAYOUT_SW128);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) (U ? mma_f16_ss_pair_e : mma_f16_ss_pair)(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    (U ? mma_commit_pair_e : mma_commit_pair)(&empty_bar[stage], (uint16_t)((1u << CL) - 1));   // this pair is done with the stage: tell every CTA
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                (U ? mma_commit_pair_e : mma_commit_pair)(&tmem_full[as], (uint16_t)(3u << lead));   // accumulators ready in both CTAs of this pair
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            };
            if (dbg & 32) { if (lane == 0) issue_loop(std::false_type{}); }
            else issue_loop(std::true_type{});
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                     // column half of the tile
        constexpr int NCH = BN / 64;                          // 32-column chunks per warp
        const int etid = threadIdx.x - 64;                    // 0..255 among the epilogue threads
        const uint32_t tile_s = base_u32 + C::OFF_EPI + (warp - 2) * 4096;   // this warp's transpose tile (1024-aligned)
        const bool out_bf16 = ep.out_bf16 != 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            const int m0 = ((t / num_n) * PAIRS + pidx) * (2 * BM_CTA) + rank * BM_CTA;
            const int nt0 = (t % num_n) * BN;
            const int n0 = nt0 + half * (BN / 2);
            const int r0 = m0 + quad * 32;                    // first row of this warp's 32-row slab
            const int row = r0 + lane;
            const bool row_ok = row < M;
            if (dbg & 8) {                                    // measurement aid: MMA-only rate (no epilogue work, output not written)
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_after();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(ptx::smem_u32(&tmem_empty[as]), lead));
                if (++as == 2) { as = 0; aphase ^= 1; }
                continue;
            }
            // ---- while the MMAs of this tile run: stage the tile's bias in smem, start the residual loads ----
            float* bs = bias_s + as * 2 * BN;
            float* cs = bs + BN;
            if (ep.bias) {
                for (int c = etid; c < BN; c += EPI_WARPS * 32) bs[c] = (nt0 + c < N) ? __ldg(ep.bias + nt0 + c) : 0.f;
            }
            // LayerNorm fold (consumer): this row's mu / rstd from the producer's partial sums, the tile's column sums in smem
            float ln_nmu = 0.f, ln_r = 1.f;
            if (ep.ln_stats) {
                for (int c = etid; c < BN; c += EPI_WARPS * 32) cs[c] = (nt0 + c < N) ? __ldg(ep.ln_c + nt0 + c) : 0.f;
                if (row_ok) {
                    float s1 = 0.f, s2 = 0.f;
                    const float2* sp = ep.ln_stats + (size_t)row * ep.ln_parts;
                    for (int p = 0; p < ep.ln_parts; ++p) { const float2 t = __ldg(sp + p); s1 += t.x; s2 += t.y; }
                    const float inv = 1.0f / (float)ep.ln_dim;
                    const float mu = s1 * inv;
                    ln_r = rsqrtf(fmaxf(s2 * inv - mu * mu, 0.f) + ep.ln_eps);
                    ln_nmu = -mu;
                }
            }
            // residual in the COALESCED mapping of the write-out: instruction i covers rows 4i..4i+3, lane -> (row 4i + lane/8,
            // 16-byte piece lane%8); two chunks in flight
            float4 q[2][8];
            auto load_resid = [&](int c, float4 (&dst)[8]) {
                const int col0 = n0 + c * 32;
                if (col0 + 32 > N) return;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int gr = r0 + 4 * i + (lane >> 3);
                    if (gr < M) {
                        const int rr = ep.resid_mod ? (gr % ep.resid_mod) : gr;
                        dst[i] = *reinterpret_cast<const float4*>(ep.resid + (size_t)rr * ep.ldr + col0 + (lane & 7) * 4);
                    }
                }
            };
            // centred hand-over (Epilogue::shift_out): the shifts of the 8 rows this lane writes out (row 4i + lane / 8)
            float csh[RESID ? 8 : 1];
            if constexpr (RESID) {
                load_resid(0, q[0]);
                if (NCH > 1) load_resid(1, q[1]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    csh[i] = 0.f;
                    const int gr = r0 + 4 * i + (lane >> 3);
                    if (ep.shift_out && gr < M && n0 < N) {
                        float cv = ep.shift_in ? __ldg(ep.shift_in + (ep.shift_in_mod ? gr % ep.shift_in_mod : gr)) : 0.f;
                        if (ep.shift_stats) {
                            float s1 = 0.f;
                            const float2* sp = ep.shift_stats + (size_t)gr * ep.shift_parts;
                            for (int p = 0; p < ep.shift_parts; ++p) s1 += __ldg(sp + p).x;
                            cv += s1 / (float)ep.shift_dim;
                        }
                        csh[i] = cv;
                        if (n0 == 0 && (lane & 7) == 0) ep.shift_out[gr] = cv;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // bias visible to all epilogue warps
            const uint32_t bsm = ptx::smem_u32(bs + half * (BN / 2));
            const uint32_t csm = ptx::smem_u32(cs + half * (BN / 2));
            float rs_sum = 0.f, rs_sq = 0.f;                  // producer side: this lane's row (r0 + 4 (lane & 7) + (lane >> 3)) statistics
            ptx::mbar_wait(&tmem_full[as], aphase);
            ptx::tc_fence_after();
            float s_sum = 0.f, s_sq = 0.f;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * (BN / 2);
            // accumulator chunks: double-buffered TMEM loads, except in the residual variant where the registers go to the
            // two residual chunks in flight (the TMEM load is short and waits in place)
            uint32_t ra[32], rb[RESID ? 1 : 32];
            if (!RESID && n0 < N) ptx::tmem_ld_x32(taddr, ra);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= N) break;
                const bool more = (c + 1 < NCH) && (col0 + 32 < N);
                uint32_t (&raw)[32] = *reinterpret_cast<uint32_t (*)[32]>((!RESID && (c & 1)) ? rb : ra);
                if constexpr (RESID) {
                    ptx::tmem_ld_x32(taddr + c * 32, ra);
                    ptx::tmem_ld_wait_dep(ra);
                } else {
                    ptx::tmem_ld_wait_dep(raw);
                    if (more) ptx::tmem_ld_x32(taddr + (c + 1) * 32, *reinterpret_cast<uint32_t (*)[32]>((c & 1) ? ra : rb));
                }
                if (col0 + 32 > N) {                          // ragged last chunk: element-wise path
                    epilogue_chunk(ep, row, col0, M, N, row_ok, raw, s_sum, s_sq);
                    continue;
                }
                // ---- row-per-thread math: + bias, statistics, GELU ----
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                if (ep.ln_stats) {                            // v = rstd * (acc - mu * c) + bias'
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b, cc;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cc.x), "=f"(cc.y), "=f"(cc.z), "=f"(cc.w) : "r"(csm + c * 128 + 16 * j));
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bsm + c * 128 + 16 * j));
                        v[4 * j + 0] = fmaf(ln_r, fmaf(ln_nmu, cc.x, v[4 * j + 0]), b.x);
                        v[4 * j + 1] = fmaf(ln_r, fmaf(ln_nmu, cc.y, v[4 * j + 1]), b.y);
                        v[4 * j + 2] = fmaf(ln_r, fmaf(ln_nmu, cc.z, v[4 * j + 2]), b.z);
                        v[4 * j + 3] = fmaf(ln_r, fmaf(ln_nmu, cc.w, v[4 * j + 3]), b.w);
                    }
                } else if (ep.bias) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bsm + c * 128 + 16 * j));
                        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                    }
                }
                if (ep.stats && row_ok) {
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
                if (out_bf16) {
                    // ---- transpose tile [32 rows][64 B]: piece j of row r at physical piece j ^ ((r >> 1) & 3) ----
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t a = tile_s + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[8 * j + 0], v[8 * j + 1])),
                                     "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                                     "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7])) : "memory");
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {             // instruction i: rows 8i..8i+7, 4 lanes (64 B) per row
                        const int rr = 8 * i + (lane >> 2), pc = lane & 3;
                        uint4 u;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                     : "r"(tile_s + rr * 64 + ((pc ^ ((rr >> 1) & 3)) << 4)));
                        const int gr = r0 + rr;
                        if (gr < M)
                            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + epilogue_out_row(ep, gr) * ep.ldo + col0 + pc * 8) = u;
                    }
                    __syncwarp();
                } else {
                    // ---- transpose tile [32 rows][128 B]: piece j of row r at physical piece j ^ (r & 7) ----
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t a = tile_s + lane * 128 + ((j ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j + 0]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]),
                                     "f"(v[4 * j + 3]) : "memory");
                    }
                    __syncwarp();
                    float st_s[RESID ? 8 : 1], st_q[RESID ? 8 : 1];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {             // instruction i: rows 4i..4i+3, 8 lanes (128 B) per row
                        const int rr = 4 * i + (lane >> 3), pc = lane & 7;
                        float4 x;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                     : "r"(tile_s + rr * 128 + ((pc ^ (rr & 7)) << 4)));
                        const int gr = r0 + rr;
                        if constexpr (RESID) {
                            if (gr < M) {
                                const float4 r = q[c & 1][i];
                                x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
                            }
                        }
                        if (gr < M) {
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + epilogue_out_row(ep, gr) * ep.ldo + col0 + pc * 4) = x;
                        }
                        if constexpr (RESID) {                // statistics and bf16 copy of the CENTRED row (Epilogue::shift_out)
                            x.x -= csh[i]; x.y -= csh[i]; x.z -= csh[i]; x.w -= csh[i];
                            st_s[i] = (x.x + x.y) + (x.z + x.w);
                            st_q[i] = fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w);
                        }
                        if (gr < M) {
                            if constexpr (RESID) {
                                if (ep.out2) {                // bf16 copy of the final rows: the next GEMM's A operand
                                    uint2 u;
                                    u.x = pack_bf16x2(x.x, x.y); u.y = pack_bf16x2(x.z, x.w);
                                    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out2) + (size_t)gr * ep.ldo2 + col0 + pc * 4) = u;
                                }
                            }
                        }
                    }
                    if constexpr (RESID) {
                        if (ep.stat_out) {
                            // 8 rows x 8 lanes -> one row per lane: butterfly over the 8 lanes of a row group (7 shuffles per
                            // quantity); lane with piece index p ends up with row-instruction i = p
                            const int p = lane & 7;
                            float a4[4], b4[4], a2[2], b2[2];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float ss = (p & 4) ? st_s[k] : st_s[k + 4], sq = (p & 4) ? st_q[k] : st_q[k + 4];
                                a4[k] = ((p & 4) ? st_s[k + 4] : st_s[k]) + __shfl_xor_sync(0xffffffffu, ss, 4);
                                b4[k] = ((p & 4) ? st_q[k + 4] : st_q[k]) + __shfl_xor_sync(0xffffffffu, sq, 4);
                            }
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const float ss = (p & 2) ? a4[k] : a4[k + 2], sq = (p & 2) ? b4[k] : b4[k + 2];
                                a2[k] = ((p & 2) ? a4[k + 2] : a4[k]) + __shfl_xor_sync(0xffffffffu, ss, 2);
                                b2[k] = ((p & 2) ? b4[k + 2] : b4[k]) + __shfl_xor_sync(0xffffffffu, sq, 2);
                            }
                            const float ss = (p & 1) ? a2[0] : a2[1], sq = (p & 1) ? b2[0] : b2[1];
                            rs_sum += ((p & 1) ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, ss, 1);
                            rs_sq += ((p & 1) ? b2[1] : b2[0]) + __shfl_xor_sync(0xffffffffu, sq, 1);
                        }
                    }
                    __syncwarp();
                    if constexpr (RESID) {
                        if (c + 2 < NCH) load_resid(c + 2, q[c & 1]);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(ptx::smem_u32(&tmem_empty[as]), lead));
            if constexpr (RESID) {
                if (ep.stat_out && n0 < N) {
                    const int gr = r0 + 4 * (lane & 7) + (lane >> 3);
                    if (gr < M) ep.stat_out[(size_t)gr * ((N + 127) / 128) + n0 / 128] = make_float2(rs_sum, rs_sq);
                }
            }
            if (ep.stats) {
                s_sum = warp_sum(s_sum);
                s_sq = warp_sum(s_sq);
                if (lane == 0 && r0 < M && n0 < N) {
                    const int sample = r0 / ep.rows_per_sample;
                    atomicAdd(ep.stats + 2 * sample, (double)s_sum);
                    atomicAdd(ep.stats + 2 * sample + 1, (double)s_sq);
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    cluster_sync_all();                                       // the peer may still be reading our smem / signalling our barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    }
}

// =====================================================================================================================
// Streamlined kernel: same TMA / tcgen05 main loop, lean epilogue.
//
// Measured on B200 (tools/gemm_bench.py, ViT-H shapes, 8 images): with the epilogue switched off the main loop runs at
// 1640 TF/s, with the generic epilogue above at 1350 TF/s; ncu of the generic kernel showed 15.6 K SASS instructions,
// ~2600 executed per warp and tile (row predicates, per-row address / remap arithmetic, constant reloads) and 35 % of the
// warp stalls on instruction fetch.  Here each epilogue warp does, per 32-column chunk: tcgen05.ld -> packed fp32x2 math
// with the thread's own row (bias or the folded LayerNorm, GELU, GroupNorm sums) -> 4 or 8 swizzled 16-byte shared stores
// into a [32 rows x 32 columns] staging box -> ONE lane issues a TMA store (or a TMA f32 reduce-add for the in-place
// residual `X += ...` of lin2), which also clips the M / N edges.  No per-row address arithmetic, no row predicates,
// no second pass through registers.
//   MODE 1: bf16 output (qkv, lin1, proj; neck GEMMs feeding another GEMM)      box 32 x 64 B, 64-byte swizzle
//   MODE 2: fp32 output (neck GEMMs with GroupNorm statistics)                   box 32 x 128 B, 128-byte swizzle
//   MODE 3: fp32 reduce-add into the output (out += acc + bias: lin2 residual)   same box, cp.reduce.async.bulk .add.f32
//   MODE 4: LayerNorm-fold producer: out = out + acc + bias in place (fp32), plus a bf16 copy of the new rows and their
//           partial row sums.  The residual box is FETCHED by TMA into the staging box (two boxes per warp, the first two
//           chunks of a tile prefetched while its MMAs run), each thread adds its own row in place (so the row sums are
//           thread-local), and both the fp32 box and a bf16 box go back out by TMA store.
enum { EPI_BF16 = 1, EPI_F32 = 2, EPI_F32_REDADD = 3, EPI_F32_RMW = 4 };

template <int BN, int MODE, bool ONEBOX = false> struct Cfg3 {
    static constexpr bool OUTF32 = MODE != EPI_BF16;
    static constexpr int NBOX = ONEBOX ? 1 : 2;
    static constexpr int A_BYTES = BM_CTA * BK * 2;
    static constexpr int B_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int CHUNK_BYTES = OUTF32 ? 4096 : 2048;       // one staging box: 32 rows x 32 columns
    static constexpr int WARP_EPI_BYTES = NBOX * CHUNK_BYTES + (MODE == 4 ? 2048 : 0);   // boxes per epilogue warp (+ one bf16 box)
    static constexpr int EPI_BYTES = EPI_WARPS * WARP_EPI_BYTES;
    // BN = 256: fp32 boxes (4 KB) leave room for 4 stages with two boxes per warp or 5 with one; bf16 boxes (2 KB) for 5 / 6
    static constexpr int STAGES = (BN == 256) ? (OUTF32 ? (ONEBOX ? 5 : 4) : (ONEBOX ? 6 : 5)) : 6;
    static constexpr int BIAS_BYTES = 2 * 2 * BN * 4;
    static constexpr int OFF_BARS = STAGES * STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_BARS + 256;
    static constexpr int OFF_EPI = OFF_BIAS + BIAS_BYTES + (1024 - (256 + BIAS_BYTES) % 1024) % 1024;
    static constexpr int SMEM_BYTES = OFF_EPI + EPI_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

template <int BN, int CL, int MODE, bool GELU, bool LNF, bool ONEBOX>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc2s_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_o2, int M, int N, int K,
                 Epilogue ep, int dbg) {
    using C = Cfg3<BN, MODE, ONEBOX>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base_u32 - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::OFF_BARS);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    uint64_t* rbar = bars + 2 * C::STAGES + 5;       // MODE 4: [epilogue warp][2] "residual box landed"
    float* bias_s = reinterpret_cast<float*>(tiles + C::OFF_BIAS);
    static_assert((2 * C::STAGES + 5 + (MODE == 4 ? 2 * EPI_WARPS : 0)) * 8 <= 256, "barrier block overflows");

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // provably warp-uniform (the issuer's operands stay in uniform registers)
    const int lane = threadIdx.x & 31;
    constexpr int PAIRS = CL / 2;
    const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
    const uint32_t rank = crank & 1, pidx = crank >> 1, lead = crank & ~1u;
    const int pair = blockIdx.x / CL, num_pairs = gridDim.x / CL;
    const int num_m = (M + 2 * BM_CTA * PAIRS - 1) / (2 * BM_CTA * PAIRS);
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_w);
        ptx::prefetch_tmap(&map_o);
        if constexpr (MODE == EPI_F32_RMW) {
            ptx::prefetch_tmap(&map_o2);
            for (int s = 0; s < 2 * EPI_WARPS; ++s) ptx::mbar_init(&rbar[s], 1);
        }
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], PAIRS);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 2 * EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    ptx::tc_fence_before();
    cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: this grid may have been made resident while the previous kernel of the stream was still
    // draining (its CTAs' set-up above overlaps that tail); everything below reads what that kernel wrote.  The trigger lets the
    // NEXT kernel do the same with this one (it still waits for this grid's completion before it touches memory).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===================== TMA producer (every CTA) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                const int m0 = ((t / num_n) * PAIRS + pidx) * (2 * BM_CTA) + rank * BM_CTA;
                const int n0 = (t % num_n) * BN + rank * (BN / 2) + pidx * (BN / 2 / PAIRS);
                // implicit 3x3 convolution (Epilogue::conv_w): row of output pixel m0 in the zero-padded map, channel blocks per tap
                int conv_row0 = 0, conv_cb = 1;
                if (ep.conv_w) {
                    const int hw = ep.conv_h * ep.conv_w, img = m0 / hw, rem = m0 % hw;
                    conv_row0 = (img * (ep.conv_h + 2) + rem / ep.conv_w) * (ep.conv_w + 2) + rem % ep.conv_w;
                    conv_cb = ep.conv_c / BK;
                }
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::A_BYTES;
                    const uint32_t full_leader = map_to_cta(ptx::smem_u32(&full_bar[stage]), lead);
                    if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                    if (ep.conv_w) {
                        const int tap = kb / conv_cb;
                        tma_load_2d_pair(sa, &map_a, full_leader, (kb % conv_cb) * BK, conv_row0 + (tap / 3) * (ep.conv_w + 2) + tap % 3);
                    } else {
                        tma_load_2d_pair(sa, &map_a, full_leader, kb * BK, m0);
                    }
                    if constexpr (PAIRS == 1) {
                        tma_load_2d_pair(sb, &map_w, full_leader, kb * BK, n0);
                    } else {
                        uint16_t mask = 0;
#pragma unroll
                        for (int q = 0; q < PAIRS; ++q) mask |= (uint16_t)(1u << (2 * q + rank));
                        tma_load_2d_pair_mc(sb + pidx * (C::B_BYTES / PAIRS), &map_w, full_leader, kb * BK, n0, mask);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair leader) =====================
        if (rank == 0) {
            // all lanes run the loop, one elected lane issues (uniform-register operands); SVB_GEMM2_DBG & 32 selects the former
            // single-thread issue (a per-MMA R2UR waterfall loop) for A/B runs
            auto issue_loop = [&](auto uniform) {
            constexpr bool U = decltype(uniform)::value;
            constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM_CTA, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = base_u32 + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    const uint64_t da = ptx::make_smem_desc(sa, 0, 1024, ptx::LAYOUT_SW128);
                    const uint64_t db = ptx::make_smem_desc(sb, 0, 1024, ptx::LAYOUT_SW128);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) (U ? mma_f16_ss_pair_e : mma_f16_ss_pair)(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    (U ? mma_commit_pair_e : mma_commit_pair)(&empty_bar[stage], (uint16_t)((1u << CL) - 1));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                (U ? mma_commit_pair_e : mma_commit_pair)(&tmem_full[as], (uint16_t)(3u << lead));
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            };
            if (dbg & 32) { if (lane == 0) issue_loop(std::false_type{}); }
            else issue_loop(std::true_type{});
        } else if (rank == 1 && ep.pad_bias && ep.remap_g && !ep.remap_h) {      // (square grids only: the SVB_PAD_IN_GEMM experiment)
            // this warp has nothing to issue in the non-leader CTA: it writes the pad rows of the window-padded output (one slice
            // per CTA pair).  pad index -> (y, x): first the right-hand strip of the g real rows, then the full bottom rows.
            const int g = ep.remap_g, gp = ep.remap_gp;
            const int npad = gp * gp - g * g, strip = g * (gp - g);
            const int v8 = N / 8;
            const int rows = (M / (g * g)) * npad;            // pad rows of this pass; one row per iteration and CTA pair,
            bf16* o = reinterpret_cast<bf16*>(ep.out);        // the index arithmetic once per row, the lanes stride over its columns
            for (int r = pair; r < rows; r += num_pairs) {
                const int pi = r % npad, b = r / npad;
                int y, x;
                if (pi < strip) { y = pi / (gp - g); x = g + pi % (gp - g); }
                else { y = g + (pi - strip) / gp; x = (pi - strip) % gp; }
                bf16* orow = o + (((size_t)b * gp + y) * gp + x) * ep.ldo;
                for (int c8 = lane; c8 < v8; c8 += 32) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.pad_bias + c8 * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.pad_bias + c8 * 8 + 4));
                    uint4 u;
                    u.x = pack_bf16x2(b0.x, b0.y); u.y = pack_bf16x2(b0.z, b0.w);
                    u.z = pack_bf16x2(b1.x, b1.y); u.w = pack_bf16x2(b1.z, b1.w);
                    *reinterpret_cast<uint4*>(orow + c8 * 8) = u;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 of every CTA) =====================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int NCH = BN / 64;
        const int etid = threadIdx.x - 64;
        const uint32_t stg = base_u32 + C::OFF_EPI + (warp - 2) * C::WARP_EPI_BYTES;
        // this lane's row of a staging box: 16-byte piece j lives at piece j ^ swz (TMA 64-byte / 128-byte swizzle)
        const uint32_t srow = stg + lane * (C::OUTF32 ? 128 : 64);
        const uint32_t swz = C::OUTF32 ? (lane & 7) : ((lane >> 1) & 3);
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * (BN / 2);
        const uint32_t empty_leader = map_to_cta(ptx::smem_u32(&tmem_empty[0]), lead);
        const int rps = ep.rows_per_sample;
        int as = 0;
        uint32_t aphase = 0;
        uint32_t sbuf = 0;                                    // staging box in use (alternates per chunk, across tiles)
        if constexpr (MODE == EPI_F32_RMW) {
            uint64_t* mybar = rbar + 2 * (warp - 2);
            uint8_t* stg_g = tiles + C::OFF_EPI + (warp - 2) * C::WARP_EPI_BYTES;      // generic address of the two fp32 boxes
            const uint32_t hbox = stg + C::NBOX * C::CHUNK_BYTES;                      // the bf16 box
            const uint32_t hrow = hbox + lane * 64;
            const uint32_t hswz = (lane >> 1) & 3;
            uint32_t rph = 0;                                 // parity bits of the two residual barriers
            const int parts = (N + 127) / 128;
            // column offset (inside the tile) of this warp's chunk c.  Default: the warp owns one 128-column half.  Interleaved
            // (SVB_GEMM2_DBG & 64, A/B): the two warps of a lane quadrant take alternate 32-column chunks, so that the boxes they
            // fetch / store at the same time are 256 contiguous bytes of the fp32 rows (128 of the bf16 copy) instead of 128 (64)
            const bool ilv = (dbg & 64) != 0 && (N % BN) == 0;     // (whole tiles only: the row-sum slots are per 128-column half)
            const uint32_t cstep = ilv ? 64u : 32u, cbase = ilv ? half * 32u : half * (BN / 2);
            const uint32_t tmem_q = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            for (int t = pair; t < num_tiles; t += num_pairs) {
                const int m0 = ((t / num_n) * PAIRS + pidx) * (2 * BM_CTA) + rank * BM_CTA;
                const int nt0 = (t % num_n) * BN;
                const int n0 = nt0 + (int)cbase;                // first column of this warp's first chunk
                const int r0 = m0 + quad * 32;
                const bool slab = (r0 < M) && (n0 < N) && !(dbg & 8);
                const int nch = slab ? min(NCH, (N - n0 + (int)cstep - 1) / (int)cstep) : 0;      // chunks of this slab (N % 32 == 0)
                // ---- while this tile's MMAs run: fetch the residual boxes of the first two chunks, stage the bias ----
                if (lane == 0) {
                    if (nch > 0) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // last tile's stores have left the boxes
                        for (int c = 0; c < C::NBOX && c < nch; ++c) {
                            ptx::mbar_expect_tx(&mybar[c], C::CHUNK_BYTES);
                            ptx::tma_load_2d(stg_g + c * C::CHUNK_BYTES, &map_o, &mybar[c], n0 + (int)cstep * c, r0);
                        }
                    }
                    if (dbg & 16) {                           // experiment: L2 prefetch of the later chunks / the next tile's rows
                        for (int c = 2; c < nch; ++c) tma_prefetch_l2_2d(&map_o, n0 + (int)cstep * c, r0);
                        const int tn = t + num_pairs;
                        if (tn < num_tiles) {
                            const int rn = ((tn / num_n) * PAIRS + pidx) * (2 * BM_CTA) + rank * BM_CTA + quad * 32;
                            const int nn = (tn % num_n) * BN + (int)cbase;
                            if (rn < M)
                                for (int c = 0; c < NCH && nn + (int)cstep * c < N; ++c) tma_prefetch_l2_2d(&map_o, nn + (int)cstep * c, rn);
                        }
                    }
                }
                float* bs = bias_s + as * 2 * BN;
                for (int c = etid; c < BN; c += EPI_WARPS * 32) bs[c] = (ep.bias && nt0 + c < N) ? __ldg(ep.bias + nt0 + c) : 0.f;
                // centred hand-over (Epilogue::shift_out): this row's shift = its mean before the update
                float cshift = 0.f;
                if (ep.shift_out && nch > 0 && r0 + lane < M) {
                    const int row = r0 + lane;
                    if (ep.shift_in) cshift = __ldg(ep.shift_in + (ep.shift_in_mod ? row % ep.shift_in_mod : row));
                    if (ep.shift_stats) {
                        float s1 = 0.f;
                        const float2* sp = ep.shift_stats + (size_t)row * ep.shift_parts;
                        for (int p = 0; p < ep.shift_parts; ++p) s1 += __ldg(sp + p).x;
                        cshift += s1 / (float)ep.shift_dim;
                    }
                    if (n0 == 0) ep.shift_out[row] = cshift;
                }
                const f32x2 nc2 = f2_pack(-cshift, -cshift);
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                const uint32_t bsm = ptx::smem_u32(bs) + cbase * 4;
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_q + as * BN + cbase;
                f32x2 st_s = f2_pack(0.f, 0.f), st_q = st_s;  // this thread's row: sum / sum of squares (two interleaved lanes)
                uint32_t ra[32], rb[32];
                bool released = false;
                if (nch > 0) ptx::tmem_ld_x32(taddr, ra);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (c >= nch) break;
                    const int col0 = n0 + c * (int)cstep;
                    uint32_t (&raw)[32] = *reinterpret_cast<uint32_t (*)[32]>((c & 1) ? rb : ra);
                    ptx::tmem_ld_wait_dep(raw);
                    if (c + 1 < nch) {
                        ptx::tmem_ld_x32(taddr + (c + 1) * cstep, *reinterpret_cast<uint32_t (*)[32]>((c & 1) ? ra : rb));
                    } else {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(empty_leader + as * 8);
                        released = true;
                    }
                    const uint32_t b = c % C::NBOX;
                    ptx::mbar_wait(&mybar[b], (rph >> b) & 1);                 // the residual box of this chunk has landed
                    rph ^= 1u << b;
                    const uint32_t rowa = srow + b * C::CHUNK_BYTES;
                    f32x2 v[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = lds128(bsm + c * (cstep * 4) + 16 * j);
                        const float4 rr = lds128(rowa + ((j ^ swz) << 4));
                        f32x2 x0 = f2_pack(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]));
                        f32x2 x1 = f2_pack(__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
                        x0 = f2_add(f2_add(x0, f2_pack(bb.x, bb.y)), f2_pack(rr.x, rr.y));
                        x1 = f2_add(f2_add(x1, f2_pack(bb.z, bb.w)), f2_pack(rr.z, rr.w));
                        // the fp32 residual stream goes back in place (this thread's own piece of its own row) ...
                        asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(rowa + ((j ^ swz) << 4)), "l"(x0), "l"(x1) : "memory");
                        // ... the statistics and the bf16 copy are taken of the centred row
                        x0 = f2_add(x0, nc2);
                        x1 = f2_add(x1, nc2);
                        st_s = f2_add(st_s, f2_add(x0, x1));
                        st_q = f2_fma(x0, x0, f2_fma(x1, x1, st_q));
                        v[2 * j] = x0; v[2 * j + 1] = x1;
                    }
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the bf16 box is free again
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float a0, a1, a2, a3, a4, a5, a6, a7;
                        f2_unpack(v[4 * j], a0, a1); f2_unpack(v[4 * j + 1], a2, a3);
                        f2_unpack(v[4 * j + 2], a4, a5); f2_unpack(v[4 * j + 3], a6, a7);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + ((j ^ hswz) << 4)), "r"(pack_bf16x2(a0, a1)),
                                     "r"(pack_bf16x2(a2, a3)), "r"(pack_bf16x2(a4, a5)), "r"(pack_bf16x2(a6, a7)) : "memory");
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_o, stg + b * C::CHUNK_BYTES, col0, r0);
                        if (ep.out2) tma_store_2d(&map_o2, hbox, col0, r0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        if (c + C::NBOX < nch) {              // refill this box with the residual of chunk c + NBOX
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            ptx::mbar_expect_tx(&mybar[b], C::CHUNK_BYTES);
                            ptx::tma_load_2d(stg_g + b * C::CHUNK_BYTES, &map_o, &mybar[b], col0 + (int)cstep * C::NBOX, r0);
                        }
                    }
                }
                if (!released) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(empty_leader + as * 8);
                }
                if (ep.stat_out && nch > 0 && r0 + lane < M) {
                    float s0, s1, q0, q1;
                    f2_unpack(st_s, s0, s1);
                    f2_unpack(st_q, q0, q1);
                    ep.stat_out[(size_t)(r0 + lane) * parts + nt0 / 128 + half] = make_float2(s0 + s1, q0 + q1);
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else
        for (int t = pair; t < num_tiles; t += num_pairs) {
            const int m0 = ((t / num_n) * PAIRS + pidx) * (2 * BM_CTA) + rank * BM_CTA;
            const int nt0 = (t % num_n) * BN;
            const int n0 = nt0 + half * (BN / 2);
            const int r0 = m0 + quad * 32;
            const int row = r0 + lane;
            const bool row_ok = row < M;
            // ---- while this tile's MMAs run: stage its bias (and LayerNorm-fold column sums), fetch the row statistics ----
            float* bs = bias_s + as * 2 * BN;
            float* cs = bs + BN;
            for (int c = etid; c < BN; c += EPI_WARPS * 32) bs[c] = (ep.bias && nt0 + c < N) ? __ldg(ep.bias + nt0 + c) : 0.f;
            float ln_nmu = 0.f, ln_r = 1.f;
            if constexpr (LNF) {
                for (int c = etid; c < BN; c += EPI_WARPS * 32) cs[c] = (nt0 + c < N) ? __ldg(ep.ln_c + nt0 + c) : 0.f;
                if (row_ok) {
                    if (ep.gn_in_stats) {             // GroupNorm(1,C) fold: statistics of the row's SAMPLE (fp64 sums)
                        const int sample = row / ep.gn_in_rows;
                        const double n = (double)ep.gn_in_rows * (double)K;
                        const double mu = ep.gn_in_stats[2 * sample] / n;
                        const double var = fmax(ep.gn_in_stats[2 * sample + 1] / n - mu * mu, 0.0);
                        ln_r = (float)(1.0 / sqrt(var + (double)ep.ln_eps));
                        ln_nmu = -(float)mu;
                    } else {
                        float s1 = 0.f, s2 = 0.f;
                        const float2* sp = ep.ln_stats + (size_t)row * ep.ln_parts;
                        for (int p = 0; p < ep.ln_parts; ++p) { const float2 q = __ldg(sp + p); s1 += q.x; s2 += q.y; }
                        const float inv = 1.0f / (float)ep.ln_dim;
                        const float mu = s1 * inv;
                        ln_r = rsqrtf(fmaxf(s2 * inv - mu * mu, 0.f) + ep.ln_eps);
                        ln_nmu = -mu;
                    }
                }
            }
            const int orow0 = (int)epilogue_out_row(ep, r0);  // a 32-row slab is contiguous in the (window-padded) output too
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            const uint32_t bsm = ptx::smem_u32(bs + half * (BN / 2));
            const uint32_t csm = ptx::smem_u32(cs + half * (BN / 2));
            ptx::mbar_wait(&tmem_full[as], aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_lane + as * BN;
            float s_sum = 0.f, s_sq = 0.f;
            uint32_t ra[32], rb[32];
            const bool slab = (r0 < M) && (n0 < N) && !(dbg & 8);
            bool released = false;
            if (slab) ptx::tmem_ld_x32(taddr, ra);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col0 = n0 + c * 32;
                if (!slab || col0 >= N) break;
                uint32_t (&raw)[32] = *reinterpret_cast<uint32_t (*)[32]>((c & 1) ? rb : ra);
                ptx::tmem_ld_wait_dep(raw);
                if (c + 1 < NCH && col0 + 32 < N) {
                    ptx::tmem_ld_x32(taddr + (c + 1) * 32, *reinterpret_cast<uint32_t (*)[32]>((c & 1) ? ra : rb));
                } else {
                    // the last TMEM read of this warp has completed: release the accumulator stage before the math / store
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(empty_leader + as * 8);
                    released = true;
                }
                // ---- this thread's row, 32 columns: bias / folded LayerNorm, GroupNorm sums, GELU (packed fp32x2) ----
                f32x2 v[16];
                const f32x2 nm2 = f2_pack(ln_nmu, ln_nmu), r2 = f2_pack(ln_r, ln_r);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = lds128(bsm + c * 128 + 16 * j);
                    f32x2 x0 = f2_pack(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]));
                    f32x2 x1 = f2_pack(__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
                    if constexpr (LNF) {
                        const float4 cc = lds128(csm + c * 128 + 16 * j);
                        x0 = f2_fma(r2, f2_fma(nm2, f2_pack(cc.x, cc.y), x0), f2_pack(b.x, b.y));
                        x1 = f2_fma(r2, f2_fma(nm2, f2_pack(cc.z, cc.w), x1), f2_pack(b.z, b.w));
                    } else {
                        x0 = f2_add(x0, f2_pack(b.x, b.y));
                        x1 = f2_add(x1, f2_pack(b.z, b.w));
                    }
                    v[2 * j] = x0; v[2 * j + 1] = x1;
                }
                if ((MODE == EPI_F32 || MODE == EPI_BF16) && ep.stats && row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a, b;
                        f2_unpack(v[j], a, b);
                        s_sum += a + b;
                        s_sq = fmaf(a, a, fmaf(b, b, s_sq));
                    }
                }
                if constexpr (GELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a, b;
                        f2_unpack(v[j], a, b);
                        gelu_pair(a, b);
                        v[j] = f2_pack(a, b);
                    }
                } else if (ep.act == 2) {                     // ReLU
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a, b;
                        f2_unpack(v[j], a, b);
                        v[j] = f2_pack(fmaxf(a, 0.f), fmaxf(b, 0.f));
                    }
                }
                // ---- stage the box, hand it to the TMA ----
                const uint32_t box = stg + sbuf * C::CHUNK_BYTES;
                if (lane == 0) {                              // the store that last read this box is done
                    if constexpr (ONEBOX) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                __syncwarp();
                const uint32_t rowa = srow + sbuf * C::CHUNK_BYTES;
                if constexpr (C::OUTF32) {
                    if (MODE == EPI_F32 && ep.out_t) {
                        // transposed box [32 columns][32 rows of the tile]: element (row = lane, column c) goes to box row c, 16-byte
                        // piece (lane / 4) ^ (c % 8), word lane % 4 — the 32 lanes of one store hit 32 different banks
                        const uint32_t tb = box + ((lane & 3) << 2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float a, b;
                            f2_unpack(v[j], a, b);
                            const uint32_t c0 = 2 * j, c1 = 2 * j + 1;
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(tb + c0 * 128 + ((((uint32_t)lane >> 2) ^ (c0 & 7)) << 4)), "f"(a) : "memory");
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(tb + c1 * 128 + ((((uint32_t)lane >> 2) ^ (c1 & 7)) << 4)), "f"(b) : "memory");
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(rowa + ((j ^ swz) << 4)), "l"(v[2 * j]), "l"(v[2 * j + 1]) : "memory");
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float a0, a1, a2, a3, a4, a5, a6, a7;
                        f2_unpack(v[4 * j], a0, a1); f2_unpack(v[4 * j + 1], a2, a3);
                        f2_unpack(v[4 * j + 2], a4, a5); f2_unpack(v[4 * j + 3], a6, a7);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((j ^ swz) << 4)), "r"(pack_bf16x2(a0, a1)),
                                     "r"(pack_bf16x2(a2, a3)), "r"(pack_bf16x2(a4, a5)), "r"(pack_bf16x2(a6, a7)) : "memory");
                    }
                }
                ptx::fence_proxy_async_smem();                // generic-proxy writes -> visible to the TMA read
                __syncwarp();
                if (lane == 0) {
                    if constexpr (MODE == EPI_F32_REDADD) tma_reduce_add_2d(&map_o, box, col0, orow0);
                    else if (MODE == EPI_F32 && ep.out_t) tma_store_2d(&map_o, box, orow0, col0);      // (inner = tile row, outer = column)
                    else tma_store_2d(&map_o, box, col0, orow0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if constexpr (!ONEBOX) sbuf ^= 1;
            }
            if (!released) {                                  // nothing to read for this slab (outside M / N)
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(empty_leader + as * 8);
            }
            if ((MODE == EPI_F32 || MODE == EPI_BF16) && ep.stats) {
                s_sum = warp_sum(s_sum);
                s_sq = warp_sum(s_sq);
                if (lane == 0 && slab) {
                    const int sample = r0 / rps;
                    atomicAdd(ep.stats + 2 * sample, (double)s_sum);
                    atomicAdd(ep.stats + 2 * sample + 1, (double)s_sq);
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
    }

    ptx::tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    }
}

template <int BN, bool RESID, int CL>
int launch_gemm2_cl(const CUtensorMap& ma, const CUtensorMap& mw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream) {
    using C = Cfg2<BN>;
    auto kern = gemm_tc2_kernel<BN, RESID, CL>;
    static int max_clusters = 0;                // co-resident clusters of this kernel (GPC sizes decide; <= num_sms / CL)
    if (!max_clusters) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(CL * (num_sms() / CL));
        qc.blockDim = dim3(NUM_THREADS);
        qc.dynamicSmemBytes = C::SMEM_BYTES;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        qc.attrs = qa; qc.numAttrs = 1;
        int n = 0;
        SVB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &qc));
        SVB_REQUIRE(n > 0, "gemm_tc2: no cluster of %d CTAs fits on this device", CL);
        max_clusters = std::min(n, num_sms() / CL);
        if (getenv("SVB_GEMM2_VERBOSE")) fprintf(stderr, "gemm_tc2<%d,%d,%d>: %d co-resident clusters\n", BN, (int)RESID, CL, max_clusters);
    }
    const int tiles = ((M + BM_CTA * CL - 1) / (BM_CTA * CL)) * ((N + BN - 1) / BN);
    const int clusters = std::min(tiles, max_clusters);
    static const int dbg = [] { const char* e = getenv("SVB_GEMM2_DBG"); return e ? atoi(e) : 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ProfScope prof(PC_GEMM, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + (ep.out_bf16 ? 2.0 : 4.0) * M * N, stream);
    SVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mw, M, N, K, ep, dbg));
    return 0;
}

template <int BN, int MODE, bool GELU, bool LNF, bool ONEBOX = false>
int launch_gemm2s(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const CUtensorMap& mo2, int M, int N, int K,
                  const Epilogue& ep, cudaStream_t stream) {
    using C = Cfg3<BN, MODE, ONEBOX>;
    constexpr int CL = 2;
    auto kern = gemm_tc2s_kernel<BN, CL, MODE, GELU, LNF, ONEBOX>;
    static bool attr_set = false;
    if (!attr_set) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = ((M + BM_CTA * CL - 1) / (BM_CTA * CL)) * ((N + BN - 1) / BN);
    const int clusters = std::min(tiles, num_sms() / CL);
    static const int dbg = [] { const char* e = getenv("SVB_GEMM2_DBG"); return e ? atoi(e) : 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    // SVB_PDL=0: plain stream order (A/B).  Default: programmatic stream serialization — the grid is scheduled as soon as the previous
    // kernel's CTAs have triggered / exited and blocks in griddepcontrol.wait until that kernel has completed
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    const double out_bytes = (MODE == EPI_BF16 ? 2.0 : (MODE == EPI_F32 ? 4.0 : (MODE == EPI_F32_REDADD ? 8.0 : 10.0))) * M * N;
    ProfScope prof(PC_GEMM, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + out_bytes, stream);
    SVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mw, mo, mo2, M, N, K, ep, dbg));
    return 0;
}

// The streamlined kernel covers: bf16 output without residual (bias or folded LayerNorm, optional GELU, optional row remap),
// fp32 output without residual (optional GroupNorm statistics), and the in-place fp32 residual out += acc + bias.
// Everything else (broadcast residual of the patch embedding, the LayerNorm-fold producer outputs) runs the generic kernel.
template <int BN>
int try_launch_streamlined(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream,
                           bool* done) {
    *done = false;
    static const int off = [] { const char* e = getenv("SVB_GEMM_EPI"); return e ? atoi(e) == 0 : 0; }();   // SVB_GEMM_EPI=0: generic epilogue only
    if (off) return 0;
    int mode = 0;
    if (ep.out_t && (ep.out_bf16 || ep.resid || ep.stats || ep.act == 1 || ep.ln_stats || ep.gn_in_stats || ep.remap_g || ep.out2 || ep.stat_out ||
                     ep.conv_w || (M % 4) != 0))
        return 0;                                                        // (the caller reports the refusal)
    if (ep.out2 || ep.stat_out) {
        // LayerNorm-fold producer: in-place fp32 residual + bf16 copy + row sums (BN = 256 only: 128-column slabs)
        if (BN != 256 || ep.out_bf16 || ep.resid != ep.out || ep.ldr != ep.ldo || ep.resid_mod || ep.stats || ep.act || ep.ln_stats ||
            ep.remap_g || (N % 32) != 0)
            return 0;
        if (ep.out2 && ((reinterpret_cast<uintptr_t>(ep.out2) & 15) != 0 || (ep.ldo2 % 8) != 0)) return 0;
        mode = EPI_F32_RMW;
    } else if (ep.out_bf16) {
        if (ep.resid) return 0;
        if (ep.stats && (ep.act || ep.ln_stats || ep.gn_in_stats)) return 0;     // statistics are taken before the activation only
        mode = EPI_BF16;
    } else {
        if (ep.act == 1 || ep.ln_stats || ep.remap_g) return 0;          // (ReLU, act == 2, is a run-time branch of every mode)
        if (ep.act && ep.resid) return 0;
        if (ep.gn_in_stats && ep.resid) return 0;
        if (!ep.resid) mode = EPI_F32;
        else if (ep.resid == ep.out && ep.ldr == ep.ldo && ep.resid_mod == 0 && !ep.stats) mode = EPI_F32_REDADD;
        else return 0;
    }
    if (ep.remap_g && ((ep.remap_g % 32) != 0 || (M % epilogue_remap_tokens(ep)) != 0)) return 0;
    const int esz = ep.out_bf16 ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(ep.out) & 15) != 0 || ((size_t)ep.ldo * esz) % 16 != 0) return 0;
    CUtensorMap ma, mw, mo;
    int rc;
    if (ep.conv_w) {      // implicit 3x3 convolution: A = the zero-padded map [batch (h + 2) (w + 2), conv_c]
        if (ep.conv_w % BM_CTA || ep.conv_c % BK || K != 9 * ep.conv_c || M % (ep.conv_h * ep.conv_w) || ep.remap_g) return 0;
        const uint64_t prow = (uint64_t)(M / (ep.conv_h * ep.conv_w)) * (ep.conv_h + 2) * (ep.conv_w + 2);
        rc = make_tmap_2d_bf16(&ma, A, (uint64_t)ep.conv_c, prow, (uint64_t)lda, BK, BM_CTA, 128);
    } else {
        rc = make_tmap_2d_bf16(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM_CTA, 128);
    }
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&mw, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, BN / 2, 128);
    if (rc) return rc;
    const uint64_t out_rows = ep.remap_g ? (uint64_t)(M / epilogue_remap_tokens(ep)) * epilogue_remap_padded(ep) : (uint64_t)M;
    if (ep.out_t) rc = make_tmap_2d(&mo, ep.out, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldo, 32, 32, 128);      // out^T [N rows][M]
    else rc = make_tmap_2d(&mo, ep.out, esz, (uint64_t)N, out_rows, (uint64_t)ep.ldo, 32, 32, ep.out_bf16 ? 64 : 128);
    if (rc) return rc;
    CUtensorMap mo2 = mo;
    if (mode == EPI_F32_RMW && ep.out2) {
        rc = make_tmap_2d(&mo2, ep.out2, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldo2, 32, 32, 64);
        if (rc) return rc;
    }
    *done = true;
    const bool lnf = ep.ln_stats != nullptr || ep.gn_in_stats != nullptr, gelu = ep.act == 1;
    if constexpr (BN == 256) {
        // producer mode: 5 stages + 1 residual box per warp (default) or 4 stages + 2 boxes.  Measured (interleaved A/B, 8 images):
        // lin2 351 vs 391 us, proj 133 vs 140 us — the deeper operand ring is worth more than the residual prefetch.
        // SVB_GEMM_RMW_ONEBOX: 0 never, 1 always (default), 2 when K >= 2048
        static const int rmw1 = [] { const char* e = getenv("SVB_GEMM_RMW_ONEBOX"); return e ? atoi(e) : 1; }();
        if (mode == EPI_F32_RMW) {
            if (rmw1 == 1 || (rmw1 == 2 && K >= 2048)) return launch_gemm2s<BN, EPI_F32_RMW, false, false, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
            return launch_gemm2s<BN, EPI_F32_RMW, false, false>(ma, mw, mo, mo2, M, N, K, ep, stream);
        }
    }
    // fp32 modes: 5 pipeline stages + 1 staging box per warp (default; measured 337 vs 358 us for lin2) or 4 stages + 2 boxes
    static const int onebox = [] { const char* e = getenv("SVB_GEMM_ONEBOX"); return e ? atoi(e) : 1; }();
    if (mode == EPI_F32 && lnf) {                                   // GroupNorm-fold consumer with fp32 output (+ its own statistics)
        if (BN == 256) return launch_gemm2s<BN, EPI_F32, false, true, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
        return launch_gemm2s<BN, EPI_F32, false, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
    }
    if (onebox && BN == 256) {
        if (mode == EPI_F32) return launch_gemm2s<BN, EPI_F32, false, false, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
        if (mode == EPI_F32_REDADD) return launch_gemm2s<BN, EPI_F32_REDADD, false, false, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
    }
    if (mode == EPI_F32) return launch_gemm2s<BN, EPI_F32, false, false>(ma, mw, mo, mo2, M, N, K, ep, stream);
    if (mode == EPI_F32_REDADD) return launch_gemm2s<BN, EPI_F32_REDADD, false, false>(ma, mw, mo, mo2, M, N, K, ep, stream);
    // bf16 mode without GELU (qkv, proj): 6 stages + one staging box (measured 218 vs 222 us for qkv); with GELU the longer
    // epilogue prefers two boxes (lin1 311 vs 314 us).  SVB_GEMM_BF16_ONEBOX=0 keeps 5 stages + two boxes everywhere.
    static const int bf16_onebox = [] { const char* e = getenv("SVB_GEMM_BF16_ONEBOX"); return e ? atoi(e) : 1; }();
    if (bf16_onebox && BN == 256 && !gelu) {
        if (lnf) return launch_gemm2s<BN, EPI_BF16, false, true, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
        return launch_gemm2s<BN, EPI_BF16, false, false, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
    }
    if (gelu && lnf) return launch_gemm2s<BN, EPI_BF16, true, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
    if (gelu) return launch_gemm2s<BN, EPI_BF16, true, false>(ma, mw, mo, mo2, M, N, K, ep, stream);
    if (lnf) return launch_gemm2s<BN, EPI_BF16, false, true>(ma, mw, mo, mo2, M, N, K, ep, stream);
    return launch_gemm2s<BN, EPI_BF16, false, false>(ma, mw, mo, mo2, M, N, K, ep, stream);
}

template <int BN>
int launch_gemm2(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream) {
    // SVB_GEMM_CLUSTER = 2 (default): one CTA pair per cluster; 4: two pairs sharing the weight tile by TMA multicast.
    // Measured on B200 (tools/gemm_bench.py, ViT-H shapes, 8 images): the multicast variant is 8 % faster PER SM (9.6 vs
    // 8.9 TF/s/SM) but only 33 clusters of 4 are co-resident (132 of 148 SMs: GPCs of 18 SMs hold 4 such clusters), so the
    // whole GEMM is 4 % slower (1268 vs 1316 TF/s).
    static const int cl = [] { const char* e = getenv("SVB_GEMM_CLUSTER"); return e ? atoi(e) : 2; }();
    if (cl != 4) {
        bool done = false;
        const int rc0 = try_launch_streamlined<BN>(A, lda, W, ldw, M, N, K, ep, stream, &done);
        if (rc0 || done) return rc0;
    }
    SVB_REQUIRE(!ep.gn_in_stats, "gemm_tc2: the GroupNorm-fold epilogue exists in the streamlined kernel only (unset SVB_GEMM_EPI / SVB_GEMM_CLUSTER)");
    SVB_REQUIRE(!ep.out_t, "gemm_tc2: the transposed-output epilogue needs the streamlined kernel, fp32 output without residual / statistics / GELU and M %% 4 == 0");
    if (ep.pad_bias && ep.remap_g) {       // the generic kernel does not write the pad rows itself
        int rc = fill_pad_rows((bf16*)ep.out, ep.pad_bias, M / epilogue_remap_tokens(ep), ep.remap_h ? ep.remap_h : ep.remap_g, ep.remap_g,
                               ep.remap_hp ? ep.remap_hp : ep.remap_gp, ep.remap_gp, ep.ldo, stream);
        if (rc) return rc;
    }
    const int pairs = (cl == 4 && M > 2 * BM_CTA) ? 2 : 1;
    CUtensorMap ma, mw;
    int rc = make_tmap_2d_bf16(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM_CTA, 128);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&mw, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, BN / 2 / pairs, 128);
    if (rc) return rc;
    if (pairs == 2) {
        if (ep.resid) return launch_gemm2_cl<BN, true, 4>(ma, mw, M, N, K, ep, stream);
        return launch_gemm2_cl<BN, false, 4>(ma, mw, M, N, K, ep, stream);
    }
    if (ep.resid) return launch_gemm2_cl<BN, true, 2>(ma, mw, M, N, K, ep, stream);
    return launch_gemm2_cl<BN, false, 2>(ma, mw, M, N, K, ep, stream);
}

}  // namespace

int gemm_bf16_tc_pair(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream) {
    // (the row-statistics output is laid out in 128-column slabs = one epilogue warp of the BN = 256 tile)
    if (N <= 128 && !ep.stat_out) return launch_gemm2<128>(A, lda, W, ldw, M, N, K, ep, stream);
    return launch_gemm2<256>(A, lda, W, ldw, M, N, K, ep, stream);
}

// 3x3 / stride 1 / pad 1 convolution on rows as an implicit GEMM (Epilogue::conv_w): out [batch h w, cout] = act(conv + bias)
int gemm_conv3x3_bf16_tc(const bf16* padded, const bf16* W, int ldw, int batch, int h, int w, int cin, int cout, const Epilogue& ep_in,
                         cudaStream_t stream) {
    SVB_REQUIRE(batch > 0 && h > 0 && w > 0 && w % 128 == 0 && cin % 64 == 0 && cout % 32 == 0,
                "gemm_conv3x3: %d x %d map with %d -> %d channels: the width must be a multiple of 128, cin of 64, cout of 32", h, w, cin, cout);
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(padded) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (ldw % 8) == 0 &&
                !ep_in.resid && !ep_in.out2 && !ep_in.stat_out && !ep_in.ln_stats && !ep_in.gn_in_stats && ep_in.act != 1,
                "gemm_conv3x3: unsupported epilogue / alignment");
    Epilogue ep = ep_in;
    ep.conv_w = w; ep.conv_h = h; ep.conv_c = cin;
    const int M = batch * h * w, K = 9 * cin;
    bool done = false;
    const int rc = (cout <= 128) ? try_launch_streamlined<128>(padded, cin, W, ldw, M, cout, K, ep, stream, &done)
                                 : try_launch_streamlined<256>(padded, cin, W, ldw, M, cout, K, ep, stream, &done);
    if (rc) return rc;
    SVB_REQUIRE(done, "gemm_conv3x3: the streamlined GEMM kernel refused this problem");
    return 0;
}

}  // namespace svb
