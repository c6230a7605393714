// Support code of the separate probe library (libsamvit_probe.so): the error channel and the tensor-map encoder the probe kernels use,
// so that the library links without the product's translation units.
#include "../../include/samvit_b200_probe.h"
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace svb {
static thread_local char g_probe_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_probe_err, sizeof(g_probe_err), fmt, ap);
    va_end(ap);
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_inner, uint32_t box_rows,
                      int swizzle_bytes) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    SVB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}
}  // namespace svb

extern "C" const char* svb_probe_last_error(void) { return svb::g_probe_err; }
