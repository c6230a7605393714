// Descriptor probe (test infrastructure, exported as svb_probe_mma): runs ONE CTA that loads an A tile [128 x K] and a
// B tile with TMA (or writes A by hand with the 128B swizzle, the path the attention kernel uses for P), issues
// K/16 tcgen05.mma with caller-supplied shared-memory descriptor fields, and dumps the 128 x N fp32 accumulator.
// tests/test_gpu_probe.py uses it to pin the MN-major and 32B-swizzle descriptor encodings against torch.
#include "../../include/samvit_b200_probe.h"
#include "common.cuh"
#include "ptx.cuh"

namespace svb {
int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner,
                      uint32_t box_rows, int swizzle_bytes);

namespace {
struct ProbeArgs {
    int K, N;
    int a_sw, b_sw;                 // swizzle bytes 128 / 32
    int a_bytes, b_bytes;           // TMA transaction bytes
    uint32_t a_lbo, a_sbo, a_kstep; // descriptor byte offsets, per-K=16 start-address advance (bytes)
    uint32_t b_lbo, b_sbo, b_kstep;
    uint32_t idesc;
    int b_atoms;                    // MN-major B: 64-element atoms along N (1 or 2)
    int a_manual;                   // 1: A written by threads (row-major [128][K] in global, K = 64*n) with manual SW128
                                    // 2: A written by threads into TMEM columns [128, 128 + K/2) as packed bf16 pairs (the P path)
    const bf16* a_gl;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, ProbeArgs pa, float* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint8_t* sA = sm;                 // up to 64 KB
    uint8_t* sB = sm + 65536;         // up to 64 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 131072);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 131072 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        ptx::mbar_init(&bar[0], 1);
        ptx::mbar_init(&bar[1], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) ptx::tmem_alloc(slot, 256);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    if (pa.a_manual == 2) {
        // thread t owns row t = TMEM lane t: column 128 + j holds (A[t][2j], A[t][2j+1]) as one 32-bit word
        const uint32_t* arow = reinterpret_cast<const uint32_t*>(pa.a_gl + (size_t)tid * pa.K);
        for (int c0 = 0; c0 < pa.K / 2; c0 += 16) {
            uint32_t w[16];
            for (int j = 0; j < 16; ++j) w[j] = arow[c0 + j];
            ptx::tmem_st_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 128 + c0, w);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
    } else if (pa.a_manual) {
        // thread t owns row t: K-major, 64-element (128 B) atoms along K, 16-byte chunk c of row r lands at chunk c ^ (r & 7)
        const int r = tid;
        for (int atom = 0; atom < pa.K / 64; ++atom)
            for (int c = 0; c < 8; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(pa.a_gl + (size_t)r * pa.K + atom * 64 + c * 8);
                *reinterpret_cast<uint4*>(sA + atom * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
            }
        ptx::fence_proxy_async_smem();
    }
    __syncthreads();
    if (tid == 0) {
        ptx::tc_fence_after();
        ptx::mbar_expect_tx(&bar[0], (pa.a_manual ? 0 : pa.a_bytes) + pa.b_bytes);
        if (!pa.a_manual) {
            if (pa.a_sw == 128) for (int k0 = 0; k0 < pa.K; k0 += 64) ptx::tma_load_2d(sA + (k0 / 64) * 16384, &map_a, &bar[0], k0, 0);
            else ptx::tma_load_2d(sA, &map_a, &bar[0], 0, 0);
        }
        ptx::tma_load_2d(sB, &map_b, &bar[0], 0, 0);
        // MN-major B wider than one 64-element atom: second atom = columns 64.. (out-of-range columns are zero-filled by the TMA)
        if (pa.b_sw == 128 && pa.b_atoms == 2) ptx::tma_load_2d(sB + pa.K * 128, &map_b, &bar[0], 64, 0);
        // ... or than one 16-element atom of the 32B swizzle: atom a = columns 16a.., K * 32 bytes apart
        if (pa.b_sw == 32) for (int a = 1; a < pa.b_atoms; ++a) ptx::tma_load_2d(sB + a * pa.K * 32, &map_b, &bar[0], 16 * a, 0);
        ptx::mbar_wait(&bar[0], 0);
        ptx::tc_fence_after();
        const uint32_t a_layout = pa.a_sw == 128 ? ptx::LAYOUT_SW128 : ptx::LAYOUT_SW32;
        const uint32_t b_layout = pa.b_sw == 128 ? ptx::LAYOUT_SW128 : ptx::LAYOUT_SW32;
        for (int k = 0; k < pa.K / 16; ++k) {
            const uint32_t a_off = pa.a_sw == 128 ? (k >> 2) * 16384 + (k & 3) * pa.a_kstep : k * pa.a_kstep;
            const uint64_t da = ptx::make_smem_desc(base + a_off, pa.a_lbo, pa.a_sbo, a_layout);
            const uint64_t db = ptx::make_smem_desc(base + 65536 + k * pa.b_kstep, pa.b_lbo, pa.b_sbo, b_layout);
            if (pa.a_manual == 2) ptx::mma_f16_ts(tmem, tmem + 128 + 8 * k, db, pa.idesc, k ? 1u : 0u);
            else ptx::mma_f16_ss(tmem, da, db, pa.idesc, k ? 1u : 0u);
        }
        ptx::mma_commit(&bar[1]);
    }
    __syncwarp();
    ptx::mbar_wait(&bar[1], 0);
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < pa.N; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(size_t)tid * pa.N + c0 + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}

// MMA issue-rate microbenchmark (svb_probe_mma_rate): ONE thread issues `reps` tcgen05.mma of one shape back to back (operands:
// zeroed shared memory / TMEM), commits and waits; out[0] = cycles from the first issue to the completion, out[1] = cycles spent
// in the issue loop.  variant: 0 SS N=128 K-major B | 1 SS N=64 | 2 TS N=64 K-major B | 3 TS N=64 MN-major B | 4 TS N=16 MN-major
// SW32 | 5 PV step (3 then 4, counted as one) | 6 TS N=128 K-major B | 7 SS N=208 | 8 TS N=80 as 64+16 K-major.  alt_d != 0: the
// accumulator alternates between two column ranges (no read-modify-write chain on one range).
template <int V>
__global__ void __launch_bounds__(128, 1) probe_rate_kernel(int reps, int alt_d, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 131072);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 131072 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    volatile int* stop = reinterpret_cast<volatile int*>(sm + 131072 + 128);
    if (tid == 0) { for (int b = 0; b < 5; ++b) ptx::mbar_init(&bar[b], 1); *stop = 0; ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc(slot, 512);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    {
        uint32_t z[16];
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        for (int c0 = 0; c0 < 512; c0 += 16) ptx::tmem_st_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, z);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (tid == 0) {
        ptx::tc_fence_after();
        const uint32_t sA = base, sB = base + 65536;
        const uint32_t a_tm = tmem + 448;                    // A operand columns (TS variants)
        uint64_t da[4], dbk[4], dbm[4], dbt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            da[k] = ptx::make_smem_desc(sA, 0, 1024, ptx::LAYOUT_SW128) + 2 * k;
            dbk[k] = ptx::make_smem_desc(sB, 0, 1024, ptx::LAYOUT_SW128) + 2 * k;
            dbm[k] = ptx::make_smem_desc(sB + k * 2048, 0, 1024, ptx::LAYOUT_SW128);
            dbt[k] = ptx::make_smem_desc(sB + 32768 + k * 512, 0, 256, ptx::LAYOUT_SW32);
        }
        const uint32_t d1 = tmem + ((alt_d & 1) ? 224 : 0);
        const bool commit_each = (alt_d & 2) != 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
            if (commit_each && r) ptx::mma_commit(&bar[1 + ((r >> 2) & 3)]);      // a commit every 4 (8) MMAs, as the attention issuers do
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t d = (k & 1) ? d1 : tmem;
                if (V == 0) ptx::mma_f16_ss(d, da[k], dbk[k], ptx::make_idesc_bf16(128, 128, 0, 0), 1u);
                if (V == 1) ptx::mma_f16_ss(d, da[k], dbk[k], ptx::make_idesc_bf16(128, 64, 0, 0), 1u);
                if (V == 2) ptx::mma_f16_ts(d, a_tm + 8 * k, dbk[k], ptx::make_idesc_bf16(128, 64, 0, 0), 1u);
                if (V == 3) ptx::mma_f16_ts(d, a_tm + 8 * k, dbm[k], ptx::make_idesc_bf16(128, 64, 0, 1), 1u);
                if (V == 4) ptx::mma_f16_ts(d + 64, a_tm + 8 * k, dbt[k], ptx::make_idesc_bf16(128, 16, 0, 1), 1u);
                if (V == 5) {
                    ptx::mma_f16_ts(d, a_tm + 8 * k, dbm[k], ptx::make_idesc_bf16(128, 64, 0, 1), 1u);
                    ptx::mma_f16_ts(d + 64, a_tm + 8 * k, dbt[k], ptx::make_idesc_bf16(128, 16, 0, 1), 1u);
                }
                if (V == 6) ptx::mma_f16_ts(d, a_tm + 8 * k, dbk[k], ptx::make_idesc_bf16(128, 128, 0, 0), 1u);
                if (V == 7) ptx::mma_f16_ss(d, da[k], dbk[k], ptx::make_idesc_bf16(128, 208, 0, 0), 1u);
                if (V == 9) ptx::mma_f16_ts(d, a_tm + 8 * k, ptx::make_smem_desc(sB + k * 2048, 16384, 1024, ptx::LAYOUT_SW128), ptx::make_idesc_bf16(128, 80, 0, 1), 1u);
                if (V == 10) ptx::mma_f16_ts(d, a_tm + 8 * k, ptx::make_smem_desc(sB + k * 512, 6656, 256, ptx::LAYOUT_SW32), ptx::make_idesc_bf16(128, 80, 0, 1), 1u);
                if (V == 8) {
                    ptx::mma_f16_ts(d, a_tm + 8 * k, dbk[k], ptx::make_idesc_bf16(128, 64, 0, 0), 1u);
                    ptx::mma_f16_ts(d + 64, a_tm + 8 * k, dbk[k], ptx::make_idesc_bf16(128, 16, 0, 0), 1u);
                }
            }
        }
        const long long t1 = clock64();
        ptx::mma_commit(&bar[0]);
        ptx::mbar_wait(&bar[0], 0);
        const long long t2 = clock64();
        out[0] = t2 - t0;
        out[1] = t1 - t0;
        *stop = 1;
    } else if ((alt_d & 4) && warp > 0) {
        // background tensor-memory traffic from the other warps (what the softmax groups do): x32 loads + x16 stores on columns
        // the MMAs do not touch
        uint32_t v[32], w[16];
        const uint32_t col = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 300;
        for (int j = 0; j < 16; ++j) w[j] = 0u;
        while (*stop == 0) {
            ptx::tmem_ld_x32(col, v);
            ptx::tmem_ld_wait_dep(v);
            ptx::tmem_ld_x32(col + 32, v);
            ptx::tmem_ld_wait_dep(v);
            w[0] = v[3];
            ptx::tmem_st_x16(col, w);
            ptx::tmem_st_x16(col + 16, w);
            ptx::tmem_st_wait();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// TMEM load / store rate microbenchmark (svb_probe_tmem_rate): `nwarps` warps (warp w works on lane quadrant w % 4) each issue `reps`
// tcgen05.ld (mode 0 / 1: one / two in flight per warp) or tcgen05.st (mode 2) of 32 lanes x 32 columns x 4 bytes back to back;
// out[w] = cycles of warp w.  Bytes per clock and SM = nwarps * reps * 4096 / max_w out[w].
__global__ void __launch_bounds__(512, 1) probe_tmem_rate_kernel(int nwarps, int reps, int mode, long long* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) ptx::tmem_alloc(&slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    if (warp < 4) {
        uint32_t z[16];
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        for (int c0 = 0; c0 < 512; c0 += 16) ptx::tmem_st_x16(lane_base + c0, z);
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp < nwarps) {
        uint32_t acc = 0;
        uint32_t va[32], vb[32];
        for (int j = 0; j < 32; ++j) { va[j] = j; vb[j] = 2 * j; }
        const long long t0 = clock64();
        if (mode == 0) {
            for (int r = 0; r < reps; ++r) {
                ptx::tmem_ld_x32(lane_base + ((r * 32) & 255), va);
                ptx::tmem_ld_wait_dep(va);
                acc ^= va[r & 31];
            }
        } else if (mode == 1) {
            for (int r = 0; r < reps; r += 2) {
                ptx::tmem_ld_x32(lane_base + ((r * 32) & 255), va);
                ptx::tmem_ld_x32(lane_base + 256 + ((r * 32) & 223), vb);
                ptx::tmem_ld_wait_dep(va);
                ptx::tmem_ld_wait_dep(vb);
                acc ^= va[r & 31] ^ vb[r & 31];
            }
        } else if (mode == 2) {
            for (int r = 0; r < reps; ++r) {
                va[0] = r;
                ptx::tmem_st_x32(lane_base + ((warp >> 2) * 128 + (r * 32)) % 480, va);
            }
            ptx::tmem_st_wait();
        } else {
            // the other load shapes, 4 KB per instruction as well (32 registers per thread): 16 lanes x 256 / 128 / 64 bits, repeated
            // along the columns 8 / 16 / 32 times — is the 56 B/clk of the 32x32b shape a property of the shape or of the port?
#define SVB_LD_SHAPE(SHAPE)                                                                                                             \
            for (int r = 0; r < reps; ++r) {                                                                                           \
                asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "    \
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                  \
                             : "=r"(va[0]), "=r"(va[1]), "=r"(va[2]), "=r"(va[3]), "=r"(va[4]), "=r"(va[5]), "=r"(va[6]), "=r"(va[7]),    \
                               "=r"(va[8]), "=r"(va[9]), "=r"(va[10]), "=r"(va[11]), "=r"(va[12]), "=r"(va[13]), "=r"(va[14]), "=r"(va[15]), \
                               "=r"(va[16]), "=r"(va[17]), "=r"(va[18]), "=r"(va[19]), "=r"(va[20]), "=r"(va[21]), "=r"(va[22]), "=r"(va[23]), \
                               "=r"(va[24]), "=r"(va[25]), "=r"(va[26]), "=r"(va[27]), "=r"(va[28]), "=r"(va[29]), "=r"(va[30]), "=r"(va[31]) \
                             : "r"(lane_base + ((r * 64) & 255)) : "memory");                                                          \
                ptx::tmem_ld_wait_dep(va);                                                                                              \
                acc ^= va[r & 31];                                                                                                      \
            }
            if (mode == 3) { SVB_LD_SHAPE("16x256b.x8") }
            else if (mode == 4) { SVB_LD_SHAPE("16x128b.x16") }
            else { SVB_LD_SHAPE("16x64b.x32") }
#undef SVB_LD_SHAPE
        }
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[warp] = (t1 - t0) + (acc == 0x12345678u ? 1 : 0);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}
}  // namespace
}  // namespace svb

using namespace svb;

extern "C" int svb_probe_mma(const void* a, const void* b, float* out, int K, int N, int a_sw, int b_sw, int b_mn_major,
                             int a_manual, unsigned a_lbo, unsigned a_sbo, unsigned a_kstep, unsigned b_lbo, unsigned b_sbo,
                             unsigned b_kstep, svb_stream_t stream) {
    SVB_REQUIRE(K % 16 == 0 && N % 16 == 0 && N <= 256, "probe: bad K/N");
    CUtensorMap ma, mb;
    ProbeArgs pa{};
    pa.K = K; pa.N = N; pa.a_sw = a_sw; pa.b_sw = b_sw;
    pa.a_lbo = a_lbo; pa.a_sbo = a_sbo; pa.a_kstep = a_kstep;
    pa.b_lbo = b_lbo; pa.b_sbo = b_sbo; pa.b_kstep = b_kstep;
    pa.idesc = ptx::make_idesc_bf16(128, N, 0, b_mn_major);
    pa.a_manual = a_manual; pa.a_gl = (const bf16*)a; pa.b_atoms = 1;
    int rc;
    // A: global [128][K] row-major; SW128 -> boxes of 64 columns (one per 64-wide atom), SW32 -> K must be 16
    if (a_sw == 128) rc = make_tmap_2d_bf16(&ma, a, K, 128, K, 64, 128, 128);
    else rc = make_tmap_2d_bf16(&ma, a, K, 128, K, 16, 128, 32);
    if (rc) return rc;
    pa.a_bytes = 128 * K * 2;
    if (!b_mn_major) {   // B global [N][K] (K contiguous)
        if (b_sw == 128) rc = make_tmap_2d_bf16(&mb, b, K, N, K, 64, N, 128);
        else rc = make_tmap_2d_bf16(&mb, b, K, N, K, 16, N, 32);
        pa.b_bytes = N * (b_sw == 128 ? 64 : 16) * 2;
        SVB_REQUIRE(K == (b_sw == 128 ? 64 : 16), "probe: K-major B supports a single atom along K");
    } else {             // B global [K][N] (N contiguous): box {N_inner, K rows}
        if (b_sw == 128) rc = make_tmap_2d_bf16(&mb, b, N, K, N, 64, K, 128);
        else rc = make_tmap_2d_bf16(&mb, b, N, K, N, 16, K, 32);
        pa.b_atoms = b_sw == 128 ? ((N > 64) ? 2 : 1) : N / 16;
        pa.b_bytes = K * (b_sw == 128 ? 64 : 16) * 2 * pa.b_atoms;
        SVB_REQUIRE(b_sw == 128 ? (N <= 128) : (N <= 128), "probe: MN-major B supports up to two 64-element atoms (SW128) or eight 16-element atoms (SW32) along N");
    }
    if (rc) return rc;
    const int smem = 131072 + 1024 + 256;
    SVB_CHECK_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, pa, out);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_probe_mma_rate(int variant, int reps, int alt_d, long long* cycles_out, svb_stream_t stream) {
    SVB_REQUIRE(cycles_out && reps > 0 && reps % 4 == 0 && variant >= 0 && variant <= 10, "probe_mma_rate: bad argument");
    const int smem = 131072 + 1024 + 256;
#define SVB_RATE(V) case V: SVB_CHECK_CUDA(cudaFuncSetAttribute(probe_rate_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
                        probe_rate_kernel<V><<<1, 128, smem, (cudaStream_t)stream>>>(reps, alt_d, cycles_out); break;
    switch (variant) { SVB_RATE(0) SVB_RATE(1) SVB_RATE(2) SVB_RATE(3) SVB_RATE(4) SVB_RATE(5) SVB_RATE(6) SVB_RATE(7) SVB_RATE(8) SVB_RATE(9) SVB_RATE(10) }
#undef SVB_RATE
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int svb_probe_tmem_rate(int nwarps, int reps, int mode, long long* cycles_out, svb_stream_t stream) {
    SVB_REQUIRE(cycles_out && nwarps >= 1 && nwarps <= 16 && reps > 0 && reps % 2 == 0 && mode >= 0 && mode <= 5, "probe_tmem_rate: bad argument");
    probe_tmem_rate_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(nwarps, reps, mode, cycles_out);
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
