// Host side of the tcgen05 GEMM (sm_100a):  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue) — TMA tensor-map encoding and the entry
// point `gemm_bf16_tc`, which validates the problem and launches the CTA-pair kernels of gemm_tc2.cu.
//
// Used for every contraction of the encoder: patch embedding (image_encoder.py:402-410 as an im2col GEMM),
// qkv / proj (image_encoder.py:227-228), lin1 / lin2 (common.py:21-22) and the SimpleFPN convolutions
// (image_encoder.py:417-447, all of which are non-overlapping and therefore plain GEMMs).
// (The single-CTA M = 128 predecessor kernel that used to live here was measured and rejected in round 1; it is in the history.)
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace svb {

namespace {

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

int gemm_bf16_tc_pair(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t stream);

bool gemm_bf16_tc_supports_fold() { return true; }

int gemm_bf16_tc(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep,
                 cudaStream_t stream) {
    SVB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16_tc: empty problem M=%d N=%d K=%d", M, N, K);
    SVB_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "gemm_bf16_tc: leading dimensions must be multiples of 8 (16 bytes)");
    SVB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                "gemm_bf16_tc: operands must be 16-byte aligned");
    SVB_REQUIRE((ep.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0, "gemm_bf16_tc: output must be 16-byte aligned");
    SVB_REQUIRE(!ep.stats || (ep.rows_per_sample % 32) == 0, "gemm_bf16_tc: rows_per_sample must be a multiple of 32");
    const bool fold = ep.ln_stats || ep.gn_in_stats || ep.out2 || ep.stat_out;
    if (fold) {
        SVB_REQUIRE(N % 32 == 0, "gemm_bf16_tc: folded-LayerNorm epilogues need N %% 32 == 0 (N = %d)", N);
        SVB_REQUIRE(!ep.ln_stats || (ep.ln_c && ep.bias && ep.ln_parts > 0 && ep.ln_dim > 0), "gemm_bf16_tc: incomplete LayerNorm-fold arguments");
        SVB_REQUIRE(!ep.gn_in_stats || (ep.ln_c && ep.bias && ep.gn_in_rows > 0 && !ep.ln_stats && !ep.act && !ep.resid && !ep.remap_g),
                    "gemm_bf16_tc: incomplete / unsupported GroupNorm-fold arguments");
        SVB_REQUIRE(!(ep.out2 || ep.stat_out) || (ep.resid && !ep.out_bf16 && ep.remap_g == 0),
                    "gemm_bf16_tc: the bf16 copy / row statistics outputs need the fp32 residual epilogue");
        SVB_REQUIRE(!ep.out2 || ((ep.ldo2 % 4) == 0 && (reinterpret_cast<uintptr_t>(ep.out2) & 7) == 0), "gemm_bf16_tc: out2 must be 8-byte aligned");
    }
    return gemm_bf16_tc_pair(A, lda, W, ldw, M, N, K, ep, stream);
}

}  // namespace svb
