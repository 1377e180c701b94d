// conv_tc.cu — 3x3 reflection-padded convolution as an implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), fed by the TMA engine, with the fused epilogues
// of kernels.cuh (ReLU | additive coupling | squeeze / unsqueeze coupling).
//
// Replaces nn.ReflectionPad2d(1)+nn.Conv2d(3x3)+(ReLU | additive coupling) of the reference's
// residual_block (models/RevResNet.py:79-88, :96-116) for the stride-1 layers with Cin % 8 == 0.
//
// GEMM view:  D[pixel, cout] += A[pixel, (tap, cin)] * B[(tap, cin), cout]
//   * Tile: R image rows x 128 pixels (one UMMA M=128 block per row) x N couts; a persistent CTA
//     (one per SM) walks tiles round-robin and double-buffers the accumulators in TMEM, so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
//   * K loop: chunks of 8 input channels; per chunk 9 taps x R rows x TERMS UMMA instructions (K=8).
//   * A operand: activations are stored in HBM in the P4 layout [C/4][H+2][W+2][4] with the
//     reflection border materialised by the producer's epilogue.  A chunk's halo tile, (R+2) rows x
//     130 px x 2 groups, is therefore 2*(R+2) contiguous 2080-byte segments: one cp.async.bulk
//     (TMA engine, UBLKCP) each, landing directly in the no-swizzle K-major canonical UMMA layout
//     [cin/4][row][pixel][4 floats] (a pixel = one 16-byte unit, 8 pixels = one core matrix).  The
//     im2col window of tap (ky,kx) is the same buffer with the descriptor start address moved by
//     (ky*PW+kx)*16 bytes: no im2col copy, no 9x re-read, no edge cases.
//   * B operand: weights pre-packed per (cout tile, chunk) as [term][tap][cin/4][cout][4 floats];
//     one cp.async.bulk per stage.
//   * Precision: kind::tf32 keeps 10 mantissa bits of each fp32 container.  Converter warps split
//     the staged activations in shared memory, x = hi + lo (hi = tf32-rounded, lo = exact
//     remainder); TERMS selects  1: Ah*Wh | 2: + Al*Wh | 3: + Ah*Wl  (fp32 accumulate in TMEM).
//   * Warp roles (448 threads): warps 0-7 epilogue (TMEM lanes 32(w%4).., column half w/4 ->
//     registers -> P4 global stores, residual loads issued ahead of the accumulator wait),
//     warps 8-11 converters, warp 12 TMA producer, warp 13 UMMA issuer + TMEM owner.
#include <stdlib.h>
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
// TMEM -> registers, 32 lanes x NC consecutive 32-bit columns (one column block per thread/lane)
template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <>
__device__ __forceinline__ void tmem_ld<64>(uint32_t taddr, float* v) {
    tmem_ld<32>(taddr, v);
    tmem_ld<32>(taddr + 32, v + 32);
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


// ------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------
template <int N, int R, int TERMS>
struct TcCfg {
    static constexpr int TA = TERMS >= 2 ? 2 : 1;   // activation terms staged (hi [, lo])
    static constexpr int TW = TERMS >= 3 ? 2 : 1;   // weight terms staged (hi [, lo])
    static constexpr int PW = 130;                  // staged pixels per halo row (128 + 2)
    static constexpr int ROWS = R + 2;
    static constexpr int ROW_BYTES = PW * 16;
    static constexpr int A_TERM_BYTES = 2 * ROWS * ROW_BYTES;
    static constexpr int A_BYTES = TA * A_TERM_BYTES;
    static constexpr int B_TERM_BYTES = 9 * 2 * N * 16;
    static constexpr int B_BYTES = TW * B_TERM_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NS_FIT = (220 * 1024) / STAGE_BYTES;
    static constexpr int NS = NS_FIT > 4 ? 4 : NS_FIT;
    static constexpr int ACC_COLS = R * N;          // one accumulator buffer
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr size_t SMEM = (size_t)NS * STAGE_BYTES + 256;
    static_assert(2 * ACC_COLS <= 512, "double-buffered accumulators exceed TMEM");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA M=128 needs N % 16 == 0");
    static_assert(NS >= 2, "need at least a double-buffered operand pipeline");
    static_assert(A_TERM_BYTES % 128 == 0 && B_TERM_BYTES % 128 == 0, "operand blocks must stay 128-byte aligned");
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
struct TcTiles {
    int n_xt, n_yt, n_ct, n_tiles;
};

template <int N, int R, int TERMS>
__global__ void __launch_bounds__(448, 1) conv3x3_tc_kernel(ConvArgs a, TcTiles tl) {
    using Cfg = TcCfg<N, R, TERMS>;
    constexpr int NS = Cfg::NS, PW = Cfg::PW, ROWS = Cfg::ROWS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stage_base = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* bars = (uint64_t*)(stage_base + (size_t)NS * Cfg::STAGE_BYTES);
    uint64_t* loaded = bars;               // [NS]  producer arrive.expect_tx + TMA bytes
    uint64_t* ready = bars + NS;           // [NS]  128 converter threads
    uint64_t* empty = bars + 2 * NS;       // [NS]  tcgen05.commit
    uint64_t* acc_full = bars + 3 * NS;    // [2]   tcgen05.commit
    uint64_t* acc_empty = bars + 3 * NS + 2;   // [2]   128 epilogue threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * NS + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = a.Cin / 8;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&loaded[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 12) {
        // ================= TMA producer =================
        if (lane == 0) {
            const int Hp = a.Hin + 2, Wp = a.Win + 2;
            const float4* in4 = reinterpret_cast<const float4*>(a.in);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
                const int ct = t % tl.n_ct, rest = t / tl.n_ct;
                const int x0 = (rest % tl.n_xt) * 128, y0 = (rest / tl.n_xt) * R;
                const float* wsrc = a.w + (size_t)ct * n_chunks * (Cfg::B_BYTES / 4);
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int s = it % NS;
                    mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
                    uint8_t* A = stage_base + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&loaded[s], Cfg::A_TERM_BYTES + Cfg::B_BYTES);
#pragma unroll
                    for (int g = 0; g < 2; ++g)
#pragma unroll
                        for (int row = 0; row < ROWS; ++row) {
                            const int py = min(y0 + row, Hp - 1);     // padded row of image row y0-1+row
                            bulk_g2s(A + (g * ROWS + row) * Cfg::ROW_BYTES,
                                     in4 + ((size_t)(2 * c + g) * Hp + py) * Wp + x0, Cfg::ROW_BYTES, &loaded[s]);
                        }
                    bulk_g2s(A + Cfg::A_BYTES, wsrc + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, &loaded[s]);
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ================= converters: hi/lo split of the staged activations =================
        const int ctid = tid - 256;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c, ++it) {
                const int s = it % NS;
                mbar_wait(&loaded[s], (it / NS) & 1);
                float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * Cfg::STAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * Cfg::STAGE_BYTES + Cfg::A_TERM_BYTES);
#pragma unroll 4
                for (int i = ctid; i < 2 * ROWS * PW; i += 128) {
                    const float4 v = hi[i];
                    const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                    hi[i] = h;
                    if (Cfg::TA == 2) lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
                fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core
                mbar_arrive(&ready[s]);
            }
        }
    } else if (warp == 13) {
        // ================= UMMA issuer =================
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t A_LBO = ROWS * Cfg::ROW_BYTES, B_LBO = N * 16, SBO = 128;
            uint32_t it = 0, tcount = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
                const uint32_t b = tcount & 1;
                mbar_wait(&acc_empty[b], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + b * Cfg::ACC_COLS;
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int s = it % NS;
                    mbar_wait(&ready[s], (it / NS) & 1);
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
                            const uint32_t d = acc + r * N;
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
                umma_commit(&acc_full[b]);       // this tile's accumulators are complete
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: TMEM -> registers -> P4 global =================
        // warp w reads TMEM lanes 32*(w%4).. (hardware rule) and the column half w/4 of each row.
        constexpr int NG = N / 8;                      // cout groups per thread per row
        const int q = warp & 3, half = warp >> 2;
        const bool coupled = a.epi >= EPI_ADD;
        uint32_t tcount = 0;
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
            const int ct = t % tl.n_ct, rest = t / tl.n_ct;
            const int x0 = (rest % tl.n_xt) * 128, y0 = (rest / tl.n_xt) * R;
            const uint32_t b = tcount & 1;
            const int x = x0 + q * 32 + lane;
            const bool xin = x < a.Wout;
            const int g0 = ct * (N / 4) + half * NG;
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + b * Cfg::ACC_COLS + half * (N / 2);
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const int y = y0 + r;
                if (y >= a.Hout) break;                                   // warp-uniform
                float4 res[NG];
                if (coupled && xin) {                                     // issued before the accumulators are needed
#pragma unroll
                    for (int j = 0; j < NG; ++j) res[j] = conv_epi_res(a, g0 + j, y, x);
                }
                if (r == 0) {
                    mbar_wait(&acc_full[b], (tcount >> 1) & 1);
                    tc_fence_after();
                }
                float v[NG * 4];
                tmem_ld<NG * 4>(tbase + (uint32_t)(r * N), v);           // warp-collective
                if (xin) {
#pragma unroll
                    for (int j = 0; j < NG; ++j) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias) + g0 + j);
                        conv_epi_store(a, g0 + j, y, x, make_float4(v[4 * j] + bv.x, v[4 * j + 1] + bv.y,
                                                                     v[4 * j + 2] + bv.z, v[4 * j + 3] + bv.w),
                                       coupled ? res[j] : make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[b]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
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
    TcTiles tl;
    tl.n_xt = cdiv(a.Wout, 128); tl.n_yt = cdiv(a.Hout, R); tl.n_ct = a.Cout / N;
    tl.n_tiles = tl.n_xt * tl.n_yt * tl.n_ct;
    const int grid = std::min(tl.n_tiles, num_sms());
    char cls[40];
    snprintf(cls, sizeof(cls), "conv3x3_tc%d %d>%d", TERMS, a.Cin, a.Cout);
    const double px = (double)a.Hout * a.Wout;
    const bool coupled = a.epi >= EPI_ADD;
    ProfScope prof(st, cls, 2.0 * 9 * a.Cin * a.Cout * px,
                   4.0 * ((double)a.Cin * a.Hin * a.Win + (coupled ? 2.0 : 1.0) * a.Cout * px));
    kern<<<grid, 448, Cfg::SMEM, st>>>(a, tl);
    return check_launch("conv3x3_tc");
}

static int tc_wide_n() {   // experiment knob: VST_TC_WIDE=128 -> N=128,R=2 tiles for Cout % 128 == 0
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VST_TC_WIDE");
        v = e ? atoi(e) : 0;
    }
    return v;
}
int tc_tile_n(int Cout) {
    if (tc_wide_n() == 128 && Cout % 128 == 0) return 128;
    return (Cout % 64 == 0) ? 64 : 16;
}

bool tc_eligible(int Cin, int Cout, int stride) {
    return stride == 1 && Cin % 8 == 0 && Cin >= 16 && (Cout % 64 == 0 || Cout == 16 || Cout == 32 || Cout == 48);
}

// a.w must point at weights packed by launch_pack_tc_weights with N = tc_tile_n(Cout) and `terms`
int launch_conv3x3_tc(const ConvArgs& a, int terms, cudaStream_t st) {
    VST_REQUIRE(tc_eligible(a.Cin, a.Cout, 1), "conv3x3_tc: shape %d>%d not eligible", a.Cin, a.Cout);
    VST_REQUIRE(a.Hin >= 2 && a.Win >= 2, "conv3x3: reflection pad needs H,W >= 2");
    VST_REQUIRE(a.Hin == a.Hout && a.Win == a.Wout, "conv3x3_tc is stride 1");
    const int N = tc_tile_n(a.Cout);
    if (N == 128) {
        if (terms == 1) return launch_tc_cfg<128, 2, 1>(a, st);
        if (terms == 2) return launch_tc_cfg<128, 2, 2>(a, st);
        return launch_tc_cfg<128, 2, 3>(a, st);
    }
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
