// tma_map.cuh — tensor-map TMA (cp.async.bulk.tensor): host-side descriptor encoding and the device-side copies.
#pragma once
#include <cuda.h>
#include "tc_ptx.cuh"

namespace vst {

// fp32 tensor of `rank` dimensions (dims[0] innermost, strides_bytes[i] = stride of dimension i+1), box = tile copied
// per instruction; no swizzle, out-of-bounds elements read as zero.  Encoded on the host per launch (a pure host
// function of a few hundred nanoseconds) and passed to the kernel as a __grid_constant__ parameter.
int make_tensor_map_f32(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box);

#ifdef __CUDACC__
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace vst
