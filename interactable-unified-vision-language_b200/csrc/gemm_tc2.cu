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
//   warps 2..9 : epilogue — two warps per TMEM lane quadrant, each owning half of the tile's columns; tcgen05.ld 32x32b,
//                bias / GELU / residual / GroupNorm statistics in registers, 16-byte global stores.
//   Two accumulator stages in TMEM (2 x BN columns): the epilogue of tile i overlaps the MMAs of tile i + 1.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

#include <algorithm>

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
    static constexpr int STAGES = (BN == 256) ? 6 : 8;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
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
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of this CTA's tile into its own smem; the completion bytes are credited to the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
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
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
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

template <int BN, bool RESID>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(200)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N, int K,
                Epilogue ep) {
    using C = Cfg2<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base_u32 - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // leader's copy is the one in use
    uint64_t* empty_bar = bars + C::STAGES;          // per CTA
    uint64_t* tmem_full = bars + 2 * C::STAGES;      // per CTA
    uint64_t* tmem_empty = bars + 2 * C::STAGES + 2; // leader's copy is the one in use
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_m = (M + 2 * BM_CTA - 1) / (2 * BM_CTA);
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_w);
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
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
                const int m0 = (t / num_n) * (2 * BM_CTA) + rank * BM_CTA;
                const int n0 = (t % num_n) * BN + rank * (BN / 2);
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::A_BYTES;
                    const uint32_t full_leader = map_to_cta(ptx::smem_u32(&full_bar[stage]), 0);
                    if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                    tma_load_2d_pair(sa, &map_a, full_leader, kb * BK, m0);
                    tma_load_2d_pair(sb, &map_w, full_leader, kb * BK, n0);
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
                    mma_commit_pair(&empty_bar[stage]);       // frees this stage in both CTAs
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                mma_commit_pair(&tmem_full[as]);              // accumulators ready in both CTAs
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                     // column half of the tile
        constexpr int NCH = BN / 64;                          // 32-column chunks per warp
        int as = 0;
        uint32_t aphase = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            const int m0 = (t / num_n) * (2 * BM_CTA) + rank * BM_CTA;
            const int n0 = (t % num_n) * BN + half * (BN / 2);
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < M;
            ptx::mbar_wait(&tmem_full[as], aphase);
            ptx::tc_fence_after();
            float s_sum = 0.f, s_sq = 0.f;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * (BN / 2);
            if constexpr (RESID) {
                // residual stream (HBM latency) prefetched one chunk ahead; the accumulator load is short and waits in place
                uint32_t ra[32];
                ResidChunk qa, qb;
                if (n0 < N) prefetch_resid(ep, row, n0, N, row_ok, qa);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (n0 + c * 32 >= N) break;
                    const bool more = (c + 1 < NCH) && (n0 + (c + 1) * 32 < N);
                    ptx::tmem_ld_x32(taddr + c * 32, ra);
                    if (c & 1) {
                        if (more) prefetch_resid(ep, row, n0 + (c + 1) * 32, N, row_ok, qa);
                        ptx::tmem_ld_wait_dep(ra);
                        epilogue_chunk(ep, row, n0 + c * 32, M, N, row_ok, ra, s_sum, s_sq, &qb);
                    } else {
                        if (more) prefetch_resid(ep, row, n0 + (c + 1) * 32, N, row_ok, qb);
                        ptx::tmem_ld_wait_dep(ra);
                        epilogue_chunk(ep, row, n0 + c * 32, M, N, row_ok, ra, s_sum, s_sq, &qa);
                    }
                }
            } else {
                uint32_t ra[32], rb[32];
                if (n0 < N) ptx::tmem_ld_x32(taddr, ra);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (n0 + c * 32 >= N) break;
                    const bool more = (c + 1 < NCH) && (n0 + (c + 1) * 32 < N);
                    if (c & 1) {
                        ptx::tmem_ld_wait_dep(rb);
                        if (more) ptx::tmem_ld_x32(taddr + (c + 1) * 32, ra);
                        epilogue_chunk(ep, row, n0 + c * 32, M, N, row_ok, rb, s_sum, s_sq);
                    } else {
                        ptx::tmem_ld_wait_dep(ra);
                        if (more) ptx::tmem_ld_x32(taddr + (c + 1) * 32, rb);
                        epilogue_chunk(ep, row, n0 + c * 32, M, N, row_ok, ra, s_sum, s_sq);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(ptx::smem_u32(&tmem_empty[as]), 0));
            if (ep.stats) {
                s_sum = warp_sum(s_sum);
                s_sq = warp_sum(s_sq);
                if (lane == 0 && (m0 + quad * 32) < M && n0 < N) {
                    const int sample = (m0 + quad * 32) / ep.rows_per_sample;
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

template <int BN>
int launch_gemm2(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream) {
    using C = Cfg2<BN>;
    CUtensorMap ma, mw;
    int rc = make_tmap_2d_bf16(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM_CTA, 128);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&mw, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, BN / 2, 128);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        SVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = ((M + 2 * BM_CTA - 1) / (2 * BM_CTA)) * ((N + BN - 1) / BN);
    const int pairs = std::min(tiles, num_sms() / 2);
    ProfScope prof(PC_GEMM, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + (ep.out_bf16 ? 2.0 : 4.0) * M * N, stream);
    if (ep.resid) gemm_tc2_kernel<BN, true><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ma, mw, M, N, K, ep);
    else gemm_tc2_kernel<BN, false><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ma, mw, M, N, K, ep);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int gemm_bf16_tc_pair(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream) {
    if (N <= 128) return launch_gemm2<128>(A, lda, W, ldw, M, N, K, ep, stream);
    return launch_gemm2<256>(A, lda, W, ldw, M, N, K, ep, stream);
}

}  // namespace svb
