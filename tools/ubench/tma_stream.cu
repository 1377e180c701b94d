// tma_stream.cu — developer microbenchmark: how fast does a bare tensor-map TMA stream (a ring of 16 KB boxes per SM,
// nothing consuming the data) read 265 MB laid out as the path's P4 planes, against the same bytes laid out contiguously?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_stream tma_stream.cu -lcuda && ./tma_stream
// Variants: planes {128 fl x 1 row x 32 groups} (the Gram's box: 512-byte runs over 32 planes), {256 x 1 x 16}, {256 x 2 x 8},
// linear {256 x 16 rows of 1 KB} (one contiguous 16 KB run per box).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 100000;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}

constexpr int BOX_BYTES = 16384;
struct Geo { int b0, b1, b2; int n0, n1, n2; int stages; };   // box dims (elements / rows / planes), boxes per dimension

// wr > 0: every wr-th box is also written back (a 16 KB bulk store to a linear buffer): wr = 2 is the 2 : 1 read : write mix
// of a coupling block (read x, read the coupling operand, write the result)
template <int NR>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, Geo g, int stages_per_cta, float* out, int wr) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* ring = sm + ((1024u - (s32(sm) & 1023u)) & 1023u);
    __shared__ uint64_t full[NR], empty[NR];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NR; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int st0 = blockIdx.x * stages_per_cta, st1 = min(st0 + stages_per_cta, g.stages);
    if (threadIdx.x == 0) {
        for (int st = st0, i = 0; st < st1; ++st, ++i) {
            const int s = i % NR;
            mbar_wait(&empty[s], ((i / NR) & 1) ^ 1);
            mbar_expect(&full[s], BOX_BYTES);
            const int i0 = st % g.n0, r = st / g.n0, i2 = r % g.n2, i1 = r / g.n2;       // innermost fastest, then planes, then rows
            tma3(ring + (size_t)s * BOX_BYTES, &tm, i0 * g.b0, i1 * g.b1, i2 * g.b2, &full[s]);
        }
    } else if (threadIdx.x == 32) {
        for (int st = st0, i = 0; st < st1; ++st, ++i) {
            const int s = i % NR;
            mbar_wait(&full[s], (i / NR) & 1);
            if (wr > 0 && i % wr == wr - 1) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (size_t)st * (BOX_BYTES / 4)),
                             "r"(s32(ring + (size_t)s * BOX_BYTES)), "n"(BOX_BYTES) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            mbar_arrive(&empty[s]);
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    const int G = 128, H = 272, W = 480 + 2;                       // 128 P4 groups of a 270 x 480 map (+ border): 268 MB
    const size_t plane = (size_t)H * W * 4, total = (size_t)G * plane;
    float* buf; CK(cudaMalloc(&buf, total * 4 + (1 << 20))); CK(cudaMemset(buf, 0, total * 4));
    float* flush; CK(cudaMalloc(&flush, 256 << 20));
    float* out; CK(cudaMalloc(&out, total * 4 + (1 << 20)));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    struct V { const char* name; int rank3; cuuint64_t d[3]; cuuint64_t s[2]; cuuint32_t b[3]; } vs[] = {
        {"planes 512 B x 32 groups ", 1, {(cuuint64_t)W * 4, (cuuint64_t)H, G}, {(cuuint64_t)W * 16, plane * 4}, {128, 1, 32}},
        {"planes 1 KB x 16 groups  ", 1, {(cuuint64_t)W * 4, (cuuint64_t)H, G}, {(cuuint64_t)W * 16, plane * 4}, {256, 1, 16}},
        {"planes 1 KB x 2 rows x 8 ", 1, {(cuuint64_t)W * 4, (cuuint64_t)H, G}, {(cuuint64_t)W * 16, plane * 4}, {256, 2, 8}},
        {"planes 1 KB x 16 rows x 1", 1, {(cuuint64_t)W * 4, (cuuint64_t)H, G}, {(cuuint64_t)W * 16, plane * 4}, {256, 16, 1}},
        {"linear 16 KB             ", 1, {256, total / 256 / 16 * 16, 1}, {1024, (cuuint64_t)total * 4}, {256, 16, 1}},
    };
    for (auto& v : vs) {
        CUtensorMap tm;
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, v.d, v.s, v.b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", v.name, (int)r); continue; }
        Geo g;
        g.b0 = v.b[0]; g.b1 = v.b[1]; g.b2 = v.b[2];
        g.n0 = (int)(v.d[0] / v.b[0]); g.n1 = (int)(v.d[1] / v.b[1]); g.n2 = (int)(v.d[2] / v.b[2]);
        g.stages = g.n0 * g.n1 * g.n2;
        const int spc = (g.stages + sms - 1) / sms;
        for (int cfg = 0; cfg < 3; ++cfg) {
            const int nr = cfg == 1 ? 12 : 6, wr = cfg == 2 ? 2 : 0;
            auto kern = nr == 6 ? stream_kernel<6> : stream_kernel<12>;
            const size_t smem = (size_t)nr * BOX_BYTES + 1024;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaMemsetAsync(flush, rep, 256 << 20));          // evict the tensor from L2
                cudaEventRecord(e0);
                kern<<<sms, 64, smem>>>(tm, g, spc, out, wr);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double bytes = (double)g.stages * BOX_BYTES * (wr ? 1.0 + 1.0 / wr : 1.0);
            printf("%s ring %2d x 16 KB%s: %.3f ms  %.0f MB  %.2f TB/s\n", v.name, nr, wr ? ", every 2nd box written back" : "", best, bytes / 1e6,
                   bytes / best / 1e9);
        }
    }
    return 0;
}
