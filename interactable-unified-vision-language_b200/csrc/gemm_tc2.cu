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

namespace svb {
int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
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
    static constexpr int BIAS_BYTES = 2 * BN * 4;                 // per accumulator stage: the tile's BN bias values
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

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int PAIRS = CL / 2;
    const uint32_t crank = cluster_ctarank();
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
        if (rank == 0 && lane == 0) {
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
                    for (int k = 0; k < BK / UMMA_K; ++k) mma_f16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    mma_commit_pair(&empty_bar[stage], (uint16_t)((1u << CL) - 1));   // this pair is done with the stage: tell every CTA
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                mma_commit_pair(&tmem_full[as], (uint16_t)(3u << lead));   // accumulators ready in both CTAs of this pair
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
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
            // ---- while the MMAs of this tile run: stage the tile's bias in smem, start the residual loads ----
            float* bs = bias_s + as * BN;
            if (ep.bias) {
                for (int c = etid; c < BN; c += EPI_WARPS * 32) bs[c] = (nt0 + c < N) ? __ldg(ep.bias + nt0 + c) : 0.f;
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
            if constexpr (RESID) {
                load_resid(0, q[0]);
                if (NCH > 1) load_resid(1, q[1]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // bias visible to all epilogue warps
            const uint32_t bsm = ptx::smem_u32(bs + half * (BN / 2));
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
                if (ep.bias) {
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
#pragma unroll
                    for (int i = 0; i < 8; ++i) {             // instruction i: rows 4i..4i+3, 8 lanes (128 B) per row
                        const int rr = 4 * i + (lane >> 3), pc = lane & 7;
                        float4 x;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                     : "r"(tile_s + rr * 128 + ((pc ^ (rr & 7)) << 4)));
                        const int gr = r0 + rr;
                        if (gr < M) {
                            if constexpr (RESID) {
                                const float4 r = q[c & 1][i];
                                x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
                            }
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + epilogue_out_row(ep, gr) * ep.ldo + col0 + pc * 4) = x;
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

template <int BN>
int launch_gemm2(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream) {
    // SVB_GEMM_CLUSTER = 2 (default): one CTA pair per cluster; 4: two pairs sharing the weight tile by TMA multicast.
    // Measured on B200 (tools/gemm_bench.py, ViT-H shapes, 8 images): the multicast variant is 8 % faster PER SM (9.6 vs
    // 8.9 TF/s/SM) but only 33 clusters of 4 are co-resident (132 of 148 SMs: GPCs of 18 SMs hold 4 such clusters), so the
    // whole GEMM is 4 % slower (1268 vs 1316 TF/s).
    static const int cl = [] { const char* e = getenv("SVB_GEMM_CLUSTER"); return e ? atoi(e) : 2; }();
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
    if (N <= 128) return launch_gemm2<128>(A, lda, W, ldw, M, N, K, ep, stream);
    return launch_gemm2<256>(A, lda, W, ldw, M, N, K, ep, stream);
}

}  // namespace svb
