/* samvit_b200_probe — hardware probes of the tcgen05 path, built as a SEPARATE shared library (libsamvit_probe.so): test and
 * measurement infrastructure, not part of the product library libsamvit_b200.so.
 *   svb_probe_mma       pins the shared-memory descriptor encodings the attention kernels rely on (tests/test_gpu_probe.py)
 *   svb_probe_mma_rate  tcgen05.mma issue-rate microbenchmark (tools/mma_rate.py)
 */
#ifndef SAMVIT_B200_PROBE_H
#define SAMVIT_B200_PROBE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* svb_stream_t; /* cudaStream_t */

const char* svb_probe_last_error(void);

/* test hook: one CTA, K/16 tcgen05.mma with caller-supplied smem-descriptor fields; dumps the 128 x N accumulator.
 * Pins the MN-major / 32B-swizzle descriptor encodings the attention kernel relies on (tests/test_gpu_probe.py). */
/* MMA issue-rate microbenchmark: cycles_out[0] = cycles for `reps` back-to-back tcgen05.mma of one shape (see probe.cu) incl.
 * completion, [1] = cycles in the issue loop (device pointers). */
int svb_probe_mma_rate(int variant, int reps, int alt_d, long long* cycles_out, svb_stream_t stream);
/* TMEM load / store rate microbenchmark: `nwarps` (1..16) warps each issue `reps` tcgen05.ld / st of 32 lanes x 32 columns back to back
 * (mode 0 / 1: loads, one / two in flight per warp; 2: stores); cycles_out[w] = cycles of warp w (device pointer, 16 entries). */
int svb_probe_tmem_rate(int nwarps, int reps, int mode, long long* cycles_out, svb_stream_t stream);
int svb_probe_mma(const void* a, const void* b, float* out, int K, int N, int a_sw, int b_sw, int b_mn_major, int a_manual,
                  unsigned a_lbo, unsigned a_sbo, unsigned a_kstep, unsigned b_lbo, unsigned b_sbo, unsigned b_kstep,
                  svb_stream_t stream);


#ifdef __cplusplus
}
#endif
#endif /* SAMVIT_B200_PROBE_H */
