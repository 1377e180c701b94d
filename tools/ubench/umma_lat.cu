// Developer microbenchmark: cost of small tcgen05.mma kind::f16 instructions (M=128, K=16) as a function of N and of
// the accumulator dependency pattern.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../vstnet_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace vst;
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// nacc accumulators used round-robin; group = consecutive UMMAs into the same accumulator
template <int NACC, int GROUP, int CE>
__global__ void k(int N, int count, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t A = smem_u32(sm), B = A + 16384;
        uint32_t ph = 0;
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            const uint64_t bd = make_desc(B, N * 16, 128);
#pragma unroll 8
            for (int i = 0; i < count; ++i) {
                const int acc = (i / GROUP) % NACC;
                umma_f16(slot + acc * N, make_desc(A + (i & 3) * 4096, 2048, 128), bd, idesc, 1u);
                if (CE && (i + 1) % CE == 0 && i + 1 < count) {
                    umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
                }
            }
            long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1;
            long long t2 = clock64();
            out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(slot, 512);
}
int main() {
    long long* d; cudaMalloc(&d, 64); long long h[6];
    int Ns[] = {16, 48, 64, 128, 192, 256};
    const char* names[] = {"same acc", "2 acc alternating", "2 acc pairs", "4 acc alternating", "same acc, commit+wait each", "same acc, commit+wait every 8"};
    for (int N : Ns) for (int p = 0; p < 6; ++p) {
        if ((p == 3 ? 4 : 2) * N > 512) continue;
        cudaFuncSetAttribute(k<1,1,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k<2,1,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k<2,2,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k<4,1,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k<1,1,1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k<1,1,8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        switch (p) {
            case 0: k<1,1,0><<<1, 128, 64 * 1024>>>(N, 96, d); break;
            case 1: k<2,1,0><<<1, 128, 64 * 1024>>>(N, 96, d); break;
            case 2: k<2,2,0><<<1, 128, 64 * 1024>>>(N, 96, d); break;
            case 3: k<4,1,0><<<1, 128, 64 * 1024>>>(N, 96, d); break;
            case 4: k<1,1,1><<<1, 128, 64 * 1024>>>(N, 96, d); break;
            case 5: k<1,1,8><<<1, 128, 64 * 1024>>>(N, 96, d); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        printf("N=%3d %-32s issue %6lld total %6lld cycles for 96 UMMAs = %.1f / UMMA  (%s)\n", N, names[p], h[4], h[5], h[5] / 96.0, cudaGetErrorString(e));
    }
    return 0;
}
