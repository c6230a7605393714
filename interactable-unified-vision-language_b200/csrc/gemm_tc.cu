// tcgen05 / TMEM GEMM fed by TMA (sm_100a):  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// Used for every contraction of the encoder: patch embedding (image_encoder.py:402-410 as an im2col GEMM),
// qkv / proj (image_encoder.py:227-228), lin1 / lin2 (common.py:21-22) and the SimpleFPN convolutions
// (image_encoder.py:417-447, all of which are non-overlapping and therefore plain GEMMs).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0     : TMA producer  — cp.async.bulk.tensor 128x64 (A) and BNx64 (W) bf16 boxes, 128B swizzle,
//                STAGES-deep smem ring, full/empty mbarriers
//   warp 1     : MMA issuer    — one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//                fp32 accumulators in TMEM, two accumulator stages (2*BN <= 512 columns) so the epilogue of
//                tile i overlaps the MMAs of tile i+1
//   warps 2..5 : epilogue      — tcgen05.ld 32x32b (one output row per thread), bias / GELU / residual /
//                GroupNorm statistics in registers, 16-byte global stores
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace svb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 bytes = one 128B-swizzle atom along K
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

template <int BN> struct Cfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N, int K,
               Epilogue ep) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles must sit on 1024-byte boundaries
    const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (base_u32 - ptx::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM;
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
            ptx::mbar_init(&tmem_empty[s], 4);     // one arrival per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::A_BYTES;
                    ptx::mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    ptx::tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m0);
                    ptx::tma_load_2d(sb, &map_w, &full_bar[stage], kb * BK, n0);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = base_u32 + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    // K-major, 128B swizzle: 8-row groups are 1024 B apart (SBO); LBO unused
                    const uint64_t da = ptx::make_smem_desc(sa, 0, 1024, ptx::LAYOUT_SW128);
                    const uint64_t db = ptx::make_smem_desc(sb, 0, 1024, ptx::LAYOUT_SW128);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing 16 elements (32 B) along K inside the swizzle atom = +2 in the >>4 address field
                        ptx::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    }
                    ptx::mma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::mma_commit(&tmem_full[as]);                 // accumulator ready for the epilogue
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quad = warp & 3;                               // TMEM lane quadrant this warp may access
        int as = 0;
        uint32_t aphase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < M;
            ptx::mbar_wait(&tmem_full[as], aphase);
            ptx::tc_fence_after();
            float s_sum = 0.f, s_sq = 0.f;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                if (n0 + c * 32 >= N) break;
                uint32_t raw[32];
                ptx::tmem_ld_32x32b_x32(taddr + c * 32, raw);
                ptx::tmem_ld_wait();
                epilogue_chunk(ep, row, n0 + c * 32, M, N, row_ok, raw, s_sum, s_sq);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
            if (ep.stats) {
                s_sum = warp_sum(s_sum);
                s_sq = warp_sum(s_sq);
                if (lane == 0 && (m0 + quad * 32) < M) {
                    const int sample = (m0 + quad * 32) / ep.rows_per_sample;
                    atomicAdd(ep.stats + 2 * sample, (double)s_sum);
                    atomicAdd(ep.stats + 2 * sample + 1, (double)s_sq);
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace

// 2-D bf16 row-major [rows, inner] with leading dimension ld (elements): box = [box_rows, box_inner], 128B swizzle
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
                 uint32_t box_rows, int swizzle_bytes);
int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
                      uint32_t box_rows, int swizzle_bytes) {
    return make_tmap_2d(map, ptr, 2, inner, rows, ld, box_inner, box_rows, swizzle_bytes);
}

// elem_bytes 2 = bf16, 4 = fp32
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
                 uint32_t box_rows, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    SVB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {ld * (cuuint64_t)elem_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu ld=%llu)", (int)r,
                (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)ld);
    return 0;
}

int encode_tmap_nd_bf16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    SVB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) st[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, st, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %d) failed with CUresult %d", rank, (int)r);
    return 0;
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BN>
static int launch_gemm(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep,
                       cudaStream_t stream) {
    using C = Cfg<BN>;
    CUtensorMap ma, mw;
    int rc = make_tmap_2d_bf16(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, 128);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&mw, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, BN, 128);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    ProfScope prof(PC_GEMM, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + (ep.out_bf16 ? 2.0 : 4.0) * M * N, stream);
    gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ma, mw, M, N, K, ep);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int gemm_bf16_tc_pair(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream);

static int gemm_impl() {
    static const int impl = [] { const char* e = getenv("SVB_GEMM_IMPL"); return e ? atoi(e) : 2; }();
    return impl;
}
bool gemm_bf16_tc_supports_fold() { return gemm_impl() != 1; }

int gemm_bf16_tc(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep,
                 cudaStream_t stream) {
    SVB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16_tc: empty problem M=%d N=%d K=%d", M, N, K);
    SVB_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "gemm_bf16_tc: leading dimensions must be multiples of 8 (16 bytes)");
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                "gemm_bf16_tc: operands must be 16-byte aligned");
    SVB_REQUIRE((ep.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0, "gemm_bf16_tc: output must be 16-byte aligned");
    SVB_REQUIRE(!ep.stats || (ep.rows_per_sample % 32) == 0, "gemm_bf16_tc: rows_per_sample must be a multiple of 32");
    // CTA-pair kernel (gemm_tc2.cu) by default; SVB_GEMM_IMPL=1 selects the single-CTA kernel below (A/B comparisons)
    const bool fold = ep.ln_stats || ep.gn_in_stats || ep.out2 || ep.stat_out;
    if (fold) {
        SVB_REQUIRE(gemm_impl() != 1, "gemm_bf16_tc: the folded-LayerNorm epilogues need the CTA-pair kernel (unset SVB_GEMM_IMPL)");
        SVB_REQUIRE(N % 32 == 0, "gemm_bf16_tc: folded-LayerNorm epilogues need N %% 32 == 0 (N = %d)", N);
        SVB_REQUIRE(!ep.ln_stats || (ep.ln_c && ep.bias && ep.ln_parts > 0 && ep.ln_dim > 0), "gemm_bf16_tc: incomplete LayerNorm-fold arguments");
        SVB_REQUIRE(!ep.gn_in_stats || (ep.ln_c && ep.bias && ep.gn_in_rows > 0 && !ep.ln_stats && !ep.act && !ep.resid && !ep.remap_g),
                    "gemm_bf16_tc: incomplete / unsupported GroupNorm-fold arguments");
        SVB_REQUIRE(!(ep.out2 || ep.stat_out) || (ep.resid && !ep.out_bf16 && ep.remap_g == 0),
                    "gemm_bf16_tc: the bf16 copy / row statistics outputs need the fp32 residual epilogue");
        SVB_REQUIRE(!ep.out2 || ((ep.ldo2 % 4) == 0 && (reinterpret_cast<uintptr_t>(ep.out2) & 7) == 0), "gemm_bf16_tc: out2 must be 8-byte aligned");
    }
    if (gemm_impl() != 1) return gemm_bf16_tc_pair(A, lda, W, ldw, M, N, K, ep, stream);
    if (ep.pad_bias && ep.remap_g) {       // this kernel does not write the pad rows itself
        int rc = fill_pad_rows((bf16*)ep.out, ep.pad_bias, M / (ep.remap_g * ep.remap_g), ep.remap_g, ep.remap_gp, ep.ldo, stream);
        if (rc) return rc;
    }
    if (N <= 128) return launch_gemm<128>(A, lda, W, ldw, M, N, K, ep, stream);
    return launch_gemm<256>(A, lda, W, ldw, M, N, K, ep, stream);
}

}  // namespace svb
