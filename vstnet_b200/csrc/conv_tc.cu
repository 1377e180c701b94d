// conv_tc.cu — 3x3 reflection-padded convolution as an implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), with the same fused epilogues as conv_direct.cu.
//
// Replaces nn.ReflectionPad2d(1)+nn.Conv2d(3x3)+(ReLU | additive coupling) of the reference's
// residual_block (models/RevResNet.py:79-88, :96-116) for the stride-1 layers with Cin % 8 == 0.
//
// GEMM view:  D[pixel, cout] += A[pixel, (tap, cin)] * B[(tap, cin), cout]
//   * CTA tile: R image rows x 128 pixels (one UMMA M=128 block per row) x N couts.
//   * K loop: chunks of 8 input channels; per chunk 9 taps x R rows x TERMS UMMA instructions (K=8).
//   * A operand: the activation halo tile of the chunk, (R+2) rows x 130 px, staged ONCE in shared
//     memory in the no-swizzle K-major canonical layout [cin/4][row][pixel][4 floats].  A pixel is
//     one 16-byte unit, 8 consecutive pixels form a core matrix, so the im2col window of tap
//     (ky,kx) is the same buffer with the descriptor start address moved by (ky*PW+kx)*16 bytes:
//     no im2col copy, no 9x re-read.  Reflection padding is index arithmetic in the loader.
//   * B operand: weights pre-packed per (cout tile, chunk) as [term][tap][cin/4][cout][4 floats];
//     one cp.async.bulk (TMA engine, UBLKCP) per stage, completion on the stage's mbarrier.
//   * Precision: kind::tf32 reads fp32 containers and ignores the low 13 mantissa bits.  The loader
//     splits every activation x = hi + lo (hi = tf32-rounded, lo exact remainder) and stages both;
//     TERMS selects  1: Ah*Wh | 2: + Al*Wh | 3: + Ah*Wl  (error-compensated, fp32 accumulate).
//   * Warp roles: warps 0-3 load/split activations then run the epilogue (TMEM lanes 32w..32w+31),
//     warp 4 issues the UMMAs (one elected lane) and owns the TMEM allocation, warp 5 streams weights.
#include "kernels.cuh"

namespace vst {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must trap (reported as a launch failure), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, M=128, K=8
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued UMMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor, version 1):
//   bits [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 16-byte K chunks) |
//   [32,46) SBO>>4 (stride between 8-row groups) | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ float tf32_round(float x) {   // round-to-nearest onto the tf32 grid
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ int reflect_clamp_tc(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// ------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------
template <int N, int R, int TERMS>
struct TcCfg {
    static constexpr int TA = TERMS >= 2 ? 2 : 1;   // activation terms staged (hi [, lo])
    static constexpr int TW = TERMS >= 3 ? 2 : 1;   // weight terms staged (hi [, lo])
    static constexpr int PW = 132;                  // smem pixels per halo row (130 used)
    static constexpr int ROWS = R + 2;
    static constexpr int A_TERM_BYTES = 2 * ROWS * PW * 16;
    static constexpr int A_BYTES = TA * A_TERM_BYTES;
    static constexpr int B_TERM_BYTES = 9 * 2 * N * 16;
    static constexpr int B_BYTES = TW * B_TERM_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NS = (2 * STAGE_BYTES <= 100 * 1024) ? 2 : ((3 * STAGE_BYTES <= 200 * 1024) ? 3 : 2);
    static constexpr int TMEM_COLS = (R * N <= 32) ? 32 : (R * N <= 64) ? 64 : (R * N <= 128) ? 128 : (R * N <= 256) ? 256 : 512;
    static constexpr size_t SMEM = (size_t)NS * STAGE_BYTES + 1024;
    static_assert(R * N <= 512, "accumulators exceed TMEM");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA M=128 needs N % 16 == 0");
};

size_t tc_packed_floats(int Cin, int Cout, int N, int terms) {
    const int TW = terms >= 3 ? 2 : 1;
    return (size_t)(Cout / N) * (Cin / 8) * TW * 9 * 2 * N * 4;
}

// raw OIHW -> [cout tile][chunk][term][tap][cin/4 (2)][n][4]
__global__ void pack_tc_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int N,
                                       int TW) {
    const size_t total = (size_t)(Cout / N) * (Cin / 8) * TW * 9 * 2 * N * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = (int)(r % 4); r /= 4;
        const int n = (int)(r % N); r /= N;
        const int g = (int)(r % 2); r /= 2;
        const int tap = (int)(r % 9); r /= 9;
        const int term = (int)(r % TW); r /= TW;
        const int chunk = (int)(r % (Cin / 8)); r /= (Cin / 8);
        const int tile = (int)r;
        const int co = tile * N + n, ci = chunk * 8 + g * 4 + e;
        const float v = w[((size_t)co * Cin + ci) * 9 + tap];
        const float hi = tf32_round(v);
        wp[i] = term == 0 ? hi : tf32_round(v - hi);
    }
}

int launch_pack_tc_weights(const float* w, float* wp, int Cin, int Cout, int N, int terms, cudaStream_t st) {
    const int TW = terms >= 3 ? 2 : 1;
    const size_t total = tc_packed_floats(Cin, Cout, N, terms);
    pack_tc_weights_kernel<<<(int)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(w, wp, Cin, Cout, N, TW);
    return check_launch("pack_tc_weights");
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int N, int R, int TERMS>
__global__ void __launch_bounds__(192, 1) conv3x3_tc_kernel(ConvArgs a) {
    using Cfg = TcCfg<N, R, TERMS>;
    constexpr int NS = Cfg::NS, PW = Cfg::PW, ROWS = Cfg::ROWS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][A | B] then barriers
    uint8_t* stage_base = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* bars = (uint64_t*)(stage_base + (size_t)NS * Cfg::STAGE_BYTES);
    uint64_t* full = bars;            // [NS]  loaders (128) + weight producer (1, with tx bytes)
    uint64_t* empty = bars + NS;      // [NS]  released by tcgen05.commit
    uint64_t* acc_full = bars + 2 * NS;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NS + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * R, tile_n = blockIdx.z;
    const int n_chunks = a.Cin / 8;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 129); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ================= activation loaders =================
        const size_t plane = (size_t)a.Hin * a.Win;
        constexpr int ITEMS = 2 * ROWS * 130;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NS, it = c / NS;
            mbar_wait(&empty[s], (it & 1) ^ 1);
            uint8_t* A = stage_base + (size_t)s * Cfg::STAGE_BYTES;
            const float* src = a.in + (size_t)c * 8 * plane;
            for (int i0 = tid; i0 < ITEMS; i0 += 128 * 4) {
                float v[4][4];
                int dst[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 128;
                    const int px = i % 130, rg = i / 130;
                    const int row = rg % ROWS, g = rg / ROWS;
                    dst[u] = (i < ITEMS) ? ((g * ROWS + row) * PW + px) * 16 : -1;
                    const int gy = reflect_clamp_tc(y0 - 1 + row, a.Hin);
                    const int gx = reflect_clamp_tc(x0 - 1 + px, a.Win);
                    const float* p = src + (size_t)(g * 4) * plane + (size_t)gy * a.Win + gx;
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[u][e] = (i < ITEMS) ? __ldg(p + (size_t)e * plane) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (dst[u] < 0) continue;
                    float4 hi = make_float4(tf32_round(v[u][0]), tf32_round(v[u][1]), tf32_round(v[u][2]), tf32_round(v[u][3]));
                    *reinterpret_cast<float4*>(A + dst[u]) = hi;
                    if (Cfg::TA == 2) {
                        float4 lo = make_float4(v[u][0] - hi.x, v[u][1] - hi.y, v[u][2] - hi.z, v[u][3] - hi.w);
                        *reinterpret_cast<float4*>(A + Cfg::A_TERM_BYTES + dst[u]) = lo;
                    }
                }
            }
            fence_proxy_async();          // make the generic-proxy smem writes visible to the tensor core
            mbar_arrive(&full[s]);
        }
    } else if (warp == 5) {
        // ================= weight producer (TMA bulk copy) =================
        if (lane == 0) {
            const float* wsrc = a.w + (size_t)tile_n * n_chunks * (Cfg::B_BYTES / 4);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % NS, it = c / NS;
                mbar_wait(&empty[s], (it & 1) ^ 1);
                uint8_t* B = stage_base + (size_t)s * Cfg::STAGE_BYTES + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full[s], Cfg::B_BYTES);
                bulk_g2s(B, wsrc + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, &full[s]);
            }
        }
    } else {
        // ================= UMMA issuer =================
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t A_LBO = ROWS * PW * 16, B_LBO = N * 16, SBO = 128;
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % NS, it = c / NS;
                mbar_wait(&full[s], it & 1);
                tc_fence_after();
                const uint32_t Aaddr = smem_u32(stage_base + (size_t)s * Cfg::STAGE_BYTES);
                const uint32_t Baddr = Aaddr + Cfg::A_BYTES;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    const uint64_t bh = make_desc(Baddr + tap * 2 * N * 16, B_LBO, SBO);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint32_t aoff = ((r + ky) * PW + kx) * 16;
                        const uint32_t d = tmem_base + r * N;
                        const uint32_t first = (c > 0 || tap > 0) ? 1u : 0u;
                        umma_tf32(d, make_desc(Aaddr + aoff, A_LBO, SBO), bh, IDESC, first);
                        if (TERMS >= 2)
                            umma_tf32(d, make_desc(Aaddr + Cfg::A_TERM_BYTES + aoff, A_LBO, SBO), bh, IDESC, 1u);
                        if (TERMS >= 3)
                            umma_tf32(d, make_desc(Aaddr + aoff, A_LBO, SBO),
                                      make_desc(Baddr + Cfg::B_TERM_BYTES + tap * 2 * N * 16, B_LBO, SBO), IDESC, 1u);
                    }
                }
                umma_commit(&empty[s]);      // smem stage reusable once these UMMAs have read it
            }
            umma_commit(acc_full);           // accumulators complete
        }
        __syncwarp();
    }

    if (warp < 4) {
        // ================= epilogue: TMEM -> registers -> global =================
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int x = x0 + warp * 32 + lane;
        const size_t out_plane = (size_t)a.Hout * a.Wout;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const int y = y0 + r;
#pragma unroll 1
            for (int cb = 0; cb < N / 16; ++cb) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(r * N + cb * 16), v);   // warp-collective
                if (y >= a.Hout || x >= a.Wout) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int co = tile_n * N + cb * 16 + j;
                    const float val = v[j] + __ldg(a.bias + co);
                    const size_t o = (size_t)co * out_plane + (size_t)y * a.Wout + x;
                    switch (a.epi) {
                        case EPI_RELU: a.out[o] = fmaxf(val, 0.f); break;
                        case EPI_NONE: a.out[o] = val; break;
                        case EPI_ADD: a.out[o] = val + a.res[o]; break;
                        case EPI_SUB: a.out[o] = a.res[o] - val; break;
                        case EPI_ADD_SQZ: {
                            int cq = a.Cout >> 2, k = co / cq, ch = co - k * cq;
                            size_t q = (size_t)ch * (out_plane * 4) + (size_t)(2 * y + (k >> 1)) * (2 * a.Wout) + (2 * x + (k & 1));
                            a.out[o] = val + a.res[q];
                        } break;
                        case EPI_SUB_UNSQZ: {
                            int cq = a.Cout >> 2, k = co / cq, ch = co - k * cq;
                            size_t q = (size_t)ch * (out_plane * 4) + (size_t)(2 * y + (k >> 1)) * (2 * a.Wout) + (2 * x + (k & 1));
                            a.out[q] = a.res[o] - val;
                        } break;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int N, int R, int TERMS>
static int launch_tc_cfg(const ConvArgs& a, cudaStream_t st) {
    using Cfg = TcCfg<N, R, TERMS>;
    static bool attr_set = false;
    auto kern = conv3x3_tc_kernel<N, R, TERMS>;
    if (!attr_set) {
        VST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    dim3 grid(cdiv(a.Wout, 128), cdiv(a.Hout, R), a.Cout / N);
    char cls[40];
    snprintf(cls, sizeof(cls), "conv3x3_tc%d %d>%d", TERMS, a.Cin, a.Cout);
    const double px = (double)a.Hout * a.Wout;
    const bool coupled = a.epi >= EPI_ADD;
    ProfScope prof(st, cls, 2.0 * 9 * a.Cin * a.Cout * px,
                   4.0 * ((double)a.Cin * a.Hin * a.Win + (coupled ? 2.0 : 1.0) * a.Cout * px));
    kern<<<grid, 192, Cfg::SMEM, st>>>(a);
    return check_launch("conv3x3_tc");
}

int tc_tile_n(int Cout) { return (Cout % 64 == 0) ? 64 : 16; }

bool tc_eligible(int Cin, int Cout, int stride) {
    return stride == 1 && Cin % 8 == 0 && Cin >= 16 && (Cout % 64 == 0 || Cout == 16 || Cout == 32 || Cout == 48);
}

// a.w must point at weights packed by launch_pack_tc_weights with N = tc_tile_n(Cout) and `terms`
int launch_conv3x3_tc(const ConvArgs& a, int terms, cudaStream_t st) {
    VST_REQUIRE(tc_eligible(a.Cin, a.Cout, 1), "conv3x3_tc: shape %d>%d not eligible", a.Cin, a.Cout);
    VST_REQUIRE(a.Hin >= 2 && a.Win >= 2, "conv3x3: reflection pad needs H,W >= 2");
    const int N = tc_tile_n(a.Cout);
    if (N == 64) {
        if (terms == 1) return launch_tc_cfg<64, 4, 1>(a, st);
        if (terms == 2) return launch_tc_cfg<64, 4, 2>(a, st);
        return launch_tc_cfg<64, 4, 3>(a, st);
    }
    if (terms == 1) return launch_tc_cfg<16, 4, 1>(a, st);
    if (terms == 2) return launch_tc_cfg<16, 4, 2>(a, st);
    return launch_tc_cfg<16, 4, 3>(a, st);
}

}  // namespace vst
