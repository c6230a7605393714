// fp32 SIMT GEMM of the VALIDATION mode:  C[M,N] = A[M,K] * W[N,K]^T with fp32 operands and fp32 FMA accumulation.
// It exists so that the encoder can be checked against the fp32 reference at 1e-4 relative L2 (BASELINE.json
// north_star: "fp32-accumulate validation mode"); bf16 operands (eps 3.9e-3) cannot meet that bar.  Same fused
// epilogue as the tcgen05 GEMM.  128x128x8 tiles, 256 threads, 8x8 outputs per thread, register-prefetched.
#include "common.cuh"

namespace svb {
namespace {

constexpr int TM = 128, TN = 128, TK = 8;

__device__ __forceinline__ void store_out(const Epilogue& ep, int row, int col, float x, float& s_sum, float& s_sq) {
    if (ep.bias) x += __ldg(ep.bias + col);
    if (ep.stats) { s_sum += x; s_sq += x * x; }
    if (ep.act == 1) x = gelu_erf(x);
    else if (ep.act == 2) x = fmaxf(x, 0.f);     // ReLU (FFN of the deformable encoder layer)
    if (ep.resid) {
        const int rr = ep.resid_mod ? (row % ep.resid_mod) : row;
        x += ep.resid[(size_t)rr * ep.ldr + col];
    }
    if (ep.out_bf16) reinterpret_cast<bf16*>(ep.out)[epilogue_out_row(ep, row) * ep.ldo + col] = __float2bfloat16_rn(x);
    else reinterpret_cast<float*>(ep.out)[epilogue_out_row(ep, row) * ep.ldo + col] = x;
}

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, int M, int N, int K, Epilogue ep) {
    __shared__ float As[2][TK][TM + 4];
    __shared__ float Ws[2][TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    // loader mapping: 128 rows x 8 k per tile = 1024 floats; thread loads 4 (one row, 4 consecutive... K is contiguous)
    const int lrow = tid >> 1;            // 0..127
    const int lk = (tid & 1) * 4;         // 0 or 4
    const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads, each 8x8 outputs (strided by 16 for conflict-free reads)

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_tile = [&](int k0, float (&ra)[4], float (&rw)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + lk + i;
            const int ar = m0 + lrow, wr = n0 + lrow;
            ra[i] = (ar < M && k < K) ? A[(size_t)ar * lda + k] : 0.f;
            rw[i] = (wr < N && k < K) ? W[(size_t)wr * ldw + k] : 0.f;
        }
    };
    float ra[4], rw[4];
    load_tile(0, ra, rw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[0][lk + i][lrow] = ra[i]; Ws[0][lk + i][lrow] = rw[i]; }
    __syncthreads();
    const int nk = (K + TK - 1) / TK;
    for (int kb = 0; kb < nk; ++kb) {
        const int cur = kb & 1;
        if (kb + 1 < nk) load_tile((kb + 1) * TK, ra, rw);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = As[cur][k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = Ws[cur][k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { As[cur ^ 1][lk + i][lrow] = ra[i]; Ws[cur ^ 1][lk + i][lrow] = rw[i]; }
        }
        __syncthreads();
    }
    float s_sum = 0.f, s_sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + ty + 16 * i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + tx + 16 * j;
            if (col < N) store_out(ep, row, col, acc[i][j], s_sum, s_sq);
        }
    }
    if (ep.stats) {
        // a 128-row tile lies inside one sample (rows_per_sample is a multiple of 128 for every neck stage)
        __shared__ float red[2][8];
        s_sum = warp_sum(s_sum);
        s_sq = warp_sum(s_sq);
        if ((tid & 31) == 0) { red[0][tid >> 5] = s_sum; red[1][tid >> 5] = s_sq; }
        __syncthreads();
        if (tid == 0) {
            double a = 0, b = 0;
            for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
            const int sample = m0 / ep.rows_per_sample;
            atomicAdd(ep.stats + 2 * sample, a);
            atomicAdd(ep.stats + 2 * sample + 1, b);
        }
    }
}

}  // namespace

int gemm_f32_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const Epilogue& ep,
                  cudaStream_t stream) {
    SVB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_f32_simt: empty problem");
    SVB_REQUIRE(!ep.stats || (ep.rows_per_sample % TM) == 0, "gemm_f32_simt: rows_per_sample must be a multiple of 128");
    dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
    ProfScope prof(PC_GEMM, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
    gemm_f32_kernel<<<grid, 256, 0, stream>>>(A, lda, W, ldw, M, N, K, ep);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace svb
