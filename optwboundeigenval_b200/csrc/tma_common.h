// tma_common.h -- tensor-map encoder (resolved from the driver at run time: the library does not link libcuda) and the
// TMA load wrapper shared by conv_tma.cu and conv_wgrad_tma.cu.
#pragma once
#include <cuda.h>            // CUtensorMap and its enums
#include <cuda_runtime.h>

#include "tc_common.cuh"

namespace b2s {

typedef CUresult (*TmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmEncodeFn tm_encoder();     // conv_tma.cu; NULL when the driver has no cuTensorMapEncodeTiled

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

}  // namespace b2s
