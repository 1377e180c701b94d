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
#include <cuda_fp16.h>
#include "tc_ptx.cuh"

namespace vst {

// ------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------
// WST ("weights stationary", HALF only, Cin = 64): a persistent CTA always works on the same cout tile (the grid is a
// multiple of the number of cout tiles), so all of its fp16 weights — 4 chunks x 36 KB for N = 128 — are fetched ONCE
// and stay resident in shared memory; a pipeline stage then carries activations only.  The 64 -> 256 coupling conv was
// bound by L2 -> SM traffic (69 KB per chunk and tile, more than half of it weights that every tile re-read).
template <int N, int R, int TERMS, bool HALF = false, bool WST = false>
struct TcCfg {
    static constexpr int KCH = HALF ? 16 : 8;       // input channels per pipeline stage (one UMMA K)
    static constexpr int TA = TERMS >= 2 ? 2 : 1;   // activation terms staged (hi [, lo])
    static constexpr int TW = TERMS >= 3 ? 2 : 1;   // weight terms staged (hi [, lo])
    static constexpr int PW = 130;                  // staged pixels per halo row (128 + 2)
    static constexpr int ROWS = R + 2;
    static constexpr int ROW_BYTES = PW * 16;
    static constexpr int A_TERM_BYTES = 2 * ROWS * ROW_BYTES;
    static constexpr int A_BYTES = TA * A_TERM_BYTES;
    static constexpr int B_TERM_BYTES = 9 * 2 * N * 16;
    static constexpr int B_BYTES = TW * B_TERM_BYTES;
    static constexpr int WST_CHUNKS = 4;            // resident weight chunks (Cin = 64)
    static constexpr int W_RES_BYTES = WST ? WST_CHUNKS * B_BYTES : 0;
    // TSPLIT (WST with two activation terms): a pipeline stage carries ONE term (hi or lo) of a chunk, i.e. the 66 KB
    // that remain beside the resident weights form 4 stages of 16.6 KB instead of 2 of 33 KB.  A stage's TMA fill takes
    // ~3.4 k cycles against ~2.2 k cycles of UMMAs per chunk (role trace), so two stages cannot hide it; four do.
    static constexpr bool TSPLIT = WST && TA == 2;
    static constexpr int NSUB = TSPLIT ? 2 : 1;     // pipeline stages per chunk
    static constexpr int STAGE_BYTES = TSPLIT ? A_TERM_BYTES : WST ? A_BYTES : A_BYTES + B_BYTES;
    static constexpr int AUX_BYTES = 2048;          // barriers (1 KB) + bias (<= 256 floats)
    static constexpr int NS_FIT = (226 * 1024 - AUX_BYTES - W_RES_BYTES) / STAGE_BYTES;
    static constexpr int NS = NS_FIT > 4 ? 4 : NS_FIT;
    static constexpr int ACC_COLS = R * N;          // one accumulator buffer
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr size_t SMEM = (size_t)NS * STAGE_BYTES + W_RES_BYTES + AUX_BYTES + 128;
    static_assert(!WST || HALF, "resident weights are for the fp16 operand path");
    static_assert(2 * ACC_COLS <= 512, "double-buffered accumulators exceed TMEM");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA M=128 needs N % 16 == 0");
    static_assert(NS >= 2, "need at least a double-buffered operand pipeline");
    static_assert(A_TERM_BYTES % 16 == 0 && A_BYTES % 128 == 0 && B_TERM_BYTES % 128 == 0, "operand blocks must stay aligned");
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
    long long* trace;   // developer aid (VST_TC_TRACE): CTA 0 stamps clock64() per role event, 4096 slots per role
};
#define TC_TRACE(role, idx) do { if (tl.trace && blockIdx.x == 0 && (idx) < 4096) tl.trace[(role) * 4096 + (idx)] = clock64(); } while (0)

constexpr int TC_THREADS = 512;   // 4 warpgroups: 2 x epilogue, converters, {operand producer, UMMA issuer, 2 idle}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 inputs, fp32 accumulate), M=128, K=16
__device__ __forceinline__ void umma_f16_tc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// HALF = true: the input is an H8 split-half tensor (kernels.cuh): hi and lo rows arrive by TMA already in
// the K-major fp16 operand layout, there is no converter stage, and the UMMAs are kind::f16 with K = 16.
template <int N, int R, int TERMS, bool HALF, bool SQZ, bool WST = false>
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_tc_kernel(ConvArgs a, TcTiles tl) {
    using Cfg = TcCfg<N, R, TERMS, HALF, WST>;
    constexpr int NS = Cfg::NS, PW = Cfg::PW, ROWS = Cfg::ROWS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stage_base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // by offset: keeps the __shared__ address space (LDS / STS)
    uint8_t* wres = stage_base + (size_t)NS * Cfg::STAGE_BYTES;        // WST: [chunk][tap][k-half][n][8 halfs], resident
    uint64_t* bars = (uint64_t*)(wres + Cfg::W_RES_BYTES);
    uint64_t* loaded = bars;                    // [NS]  operand producer arrive.expect_tx + TMA bytes
    uint64_t* ready = bars + NS;                // [NS]  128 converter threads
    uint64_t* empty = bars + 2 * NS;            // [NS]  tcgen05.commit
    uint64_t* acc_full = bars + 3 * NS;         // [2]   tcgen05.commit
    uint64_t* acc_empty = bars + 3 * NS + 2;    // [2]   256 epilogue threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * NS + 4);
    uint64_t* wbar = bars + 3 * NS + 5;         // WST: the resident weights have landed
    float* bias_s = (float*)((uint8_t*)bars + 1024);
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < a.Cout && i < 256; i += blockDim.x) bias_s[i] = __ldg(a.bias + i);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = a.Cin / Cfg::KCH;
    // SQZ (compile time): the instantiation of the stride-2 blocks' coupling modes — the coupling operand is read
    // through the squeeze addressing (EPI_ADD_SQZ) or the result stored through the unsqueeze addressing
    // (EPI_SUB_UNSQZ); their address arithmetic cannot cost the plain instantiation registers
    const bool coupled = a.epi >= EPI_ADD;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&loaded[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                          // prologue above touches only constants (bias) and on-chip state

    // Roles are dispatched per warpgroup, each branch starting with its register-file rebalancing
    // (setmaxnreg): the epilogue keeps a whole tile of coupling operands in registers — that is the
    // memory-level parallelism of the kernel — everybody else needs very few.
    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 12 && lane == 0) {
            // ================= operand producer (TMA) =================
            const int Hp = a.Hin + 2, Wp = a.Win + 2;
            const float4* in4 = reinterpret_cast<const float4*>(a.in);
            uint32_t it = 0;
            if (WST) {
                // every tile of this CTA has cout tile blockIdx.x % n_ct (gridDim.x is a multiple of n_ct)
                const float* wsrc = a.w + (size_t)(blockIdx.x % tl.n_ct) * n_chunks * (Cfg::B_BYTES / 4);
                mbar_arrive_expect_tx(wbar, (uint32_t)(n_chunks * Cfg::B_BYTES));
                for (int c = 0; c < n_chunks; ++c)
                    bulk_g2s(wres + (size_t)c * Cfg::B_BYTES, wsrc + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, wbar);
            }
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
                const int ct = t % tl.n_ct, rest = t / tl.n_ct;
                const int x0 = (rest % tl.n_xt) * 128, y0 = (rest / tl.n_xt) * R;
                const float* wsrc = a.w + (size_t)ct * n_chunks * (Cfg::B_BYTES / 4);
                for (int c = 0; c < n_chunks; ++c)
#pragma unroll
                for (int sub = 0; sub < Cfg::NSUB; ++sub, ++it) {
                    const int s = it % NS;
                    mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
                    TC_TRACE(0, it);
                    uint8_t* A = stage_base + (size_t)s * Cfg::STAGE_BYTES;
                    constexpr int NT = Cfg::TSPLIT ? 1 : HALF ? Cfg::TA : 1;   // tensors per stage: one term | raw fp32 | hi, lo
                    mbar_arrive_expect_tx(&loaded[s], NT * Cfg::A_TERM_BYTES + (WST ? 0 : Cfg::B_BYTES));
#pragma unroll
                    for (int tt = 0; tt < NT; ++tt) {
                        const int term = Cfg::TSPLIT ? sub : tt;
                        // H8: the lo planes follow the Cin/8 hi planes; both are addressed like P4 groups
                        const float4* src = in4 + (size_t)term * (a.Cin / 8) * Hp * Wp;
#pragma unroll
                        for (int g = 0; g < 2; ++g)
#pragma unroll
                            for (int row = 0; row < ROWS; ++row) {
                                const int py = min(y0 + row, Hp - 1);     // padded row of image row y0-1+row
                                bulk_g2s(A + tt * Cfg::A_TERM_BYTES + (g * ROWS + row) * Cfg::ROW_BYTES,
                                         src + ((size_t)(2 * c + g) * Hp + py) * Wp + x0, Cfg::ROW_BYTES, &loaded[s]);
                            }
                    }
                    if (!WST) bulk_g2s(A + Cfg::A_BYTES, wsrc + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, &loaded[s]);
                }
            }
        } else if (warp == 13) {
            // ================= UMMA issuer =================
            // The whole warp walks the (warp-uniform) schedule and one elected lane issues the tcgen05 instructions:
            // issued from a divergent single-lane branch, every UTCHMMA is wrapped in an election loop and its
            // descriptors are rebuilt through the vector registers — ~12 instructions per UMMA, which paced this role at
            // 70-95 cycles per UMMA against the tensor pipe's 64 (role trace).  Descriptors are constant high bits +
            // (shared address >> 4): one 64-bit add per operand.
            // instruction descriptor: D = F32 (bit 4); A, B = TF32 (format 2) | F16 (format 0); N at bit 17, M at bit 24
            constexpr uint32_t IDESC = (1u << 4) | (HALF ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t A_LBO = ROWS * Cfg::ROW_BYTES, B_LBO = N * 16, SBO = 128;
            uint64_t* const opbar = HALF ? loaded : ready;      // no converter stage for pre-split operands
            const bool el = tc_elect_one();
            const uint64_t dA0 = make_desc(smem_u32(stage_base), A_LBO, SBO);
            const uint64_t dB0 = make_desc(WST ? smem_u32(wres) : smem_u32(stage_base) + Cfg::A_BYTES, B_LBO, SBO);
            uint32_t it = 0, tcount = 0;
            if (WST) mbar_wait(wbar, 0);
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
                const uint32_t b = tcount & 1;
                mbar_wait(&acc_empty[b], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) TC_TRACE(5, tcount);
                const uint32_t acc = tmem_base + b * Cfg::ACC_COLS;
                for (int c = 0; c < n_chunks; ++c)
#pragma unroll
                for (int sub = 0; sub < Cfg::NSUB; ++sub, ++it) {
                    const int s = it % NS;
                    mbar_wait(&opbar[s], (it / NS) & 1);
                    tc_fence_after();
                    if (lane == 0) TC_TRACE(3, it);
                    const uint64_t dA = dA0 + (uint64_t)((uint32_t)s * (Cfg::STAGE_BYTES >> 4));
                    const uint64_t dB = WST ? dB0 + (uint64_t)((uint32_t)c * (Cfg::B_BYTES >> 4)) : dB0 + (uint64_t)((uint32_t)s * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        const uint64_t bh = dB + (uint64_t)((tap * 2 * N * 16) >> 4);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint64_t ah = dA + (uint64_t)((((r + ky) * PW + kx) * 16) >> 4);
                            const uint32_t d = acc + r * N;
                            const uint32_t first = (c > 0 || sub > 0 || tap > 0) ? 1u : 0u;
                            if (el) {
                                if (Cfg::TSPLIT) {
                                    umma_f16_tc(d, ah, bh, IDESC, first);     // this stage's term
                                } else if (HALF) {
                                    umma_f16_tc(d, ah, bh, IDESC, first);
                                    if (TERMS >= 2) umma_f16_tc(d, ah + (uint64_t)(Cfg::A_TERM_BYTES >> 4), bh, IDESC, 1u);
                                } else {
                                    umma_tf32(d, ah, bh, IDESC, first);
                                    if (TERMS >= 2) umma_tf32(d, ah + (uint64_t)(Cfg::A_TERM_BYTES >> 4), bh, IDESC, 1u);
                                    if (TERMS >= 3) umma_tf32(d, ah, bh + (uint64_t)(Cfg::B_TERM_BYTES >> 4), IDESC, 1u);
                                }
                            }
                        }
                    }
                    if (el) umma_commit(&empty[s]);      // smem stage reusable once these UMMAs have read it
                    if (lane == 0) TC_TRACE(4, it);
                }
                if (el) umma_commit(&acc_full[b]);       // this tile's accumulators are complete
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ================= converters: hi/lo split of the staged activations =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int ctid = tid - 256;
        uint32_t it = 0;
        for (int t = blockIdx.x; !HALF && t < tl.n_tiles; t += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c, ++it) {
                const int s = it % NS;
                mbar_wait(&loaded[s], (it / NS) & 1);
                if (ctid == 0) TC_TRACE(1, it);
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
                if (ctid == 0) TC_TRACE(2, it);
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> P4 global =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        // warp w reads TMEM lanes 32*(w%4).. (hardware rule) and the column half w/4 of each row.  The
        // coupling operand of the WHOLE tile is requested up front (R x NG 16-byte loads per thread,
        // 128 KB in flight per SM) — normally while the tile's UMMAs are still running — so the
        // per-element work below is branch-free register arithmetic and fire-and-forget stores:
        //   v = (acc + bias) * sgn ;  v = max(v, floor) ;  out = res + v
        // (exact re-statements of relu(v) | v | res + v | res - v).
        constexpr int NG = N / 8;                                         // cout groups per thread per row
        constexpr int CH = (NG * 4 >= 16) ? 16 : NG * 4;                  // TMEM columns per tcgen05.ld
        const int q = warp & 3, half = warp >> 2;
        const float sgn = (a.epi == EPI_SUB || a.epi == EPI_SUB_UNSQZ) ? -1.f : 1.f;
        const float flo = (a.epi == EPI_RELU) ? 0.f : -INFINITY;
        constexpr float unscale = HALF ? 1.0f / VST_HALF_SCALE : 1.0f;    // H8 operands carry VST_HALF_SCALE * x
        const int H = a.Hout, W = a.Wout, Wp = W + 2;
        const size_t plane = p4_plane_px(H, W);
        // squeeze modes: cout group g = k * gq + gs is group gs of the un-squeezed tensor [Cout/16][2H+2][2W+2][4] at
        // pixel (2y + k/2, 2x + k%2)   (models/RevResNet.py:34-43)
        const bool res_sq = SQZ && a.epi == EPI_ADD_SQZ, out_us = SQZ && a.epi == EPI_SUB_UNSQZ;
        const int gq = a.Cout >> 4, lgq = __ffs(gq) - 1;
        const int Hh = 2 * H, Wh = 2 * W, Wph = Wh + 2;
        const size_t plane_h = p4_plane_px(Hh, Wh);
        uint32_t tcount = 0;
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
            const int ct = t % tl.n_ct, rest = t / tl.n_ct;
            const int x0 = (rest % tl.n_xt) * 128, y0 = (rest / tl.n_xt) * R;
            const uint32_t b = tcount & 1;
            const int x = x0 + q * 32 + lane;
            const bool xin = x < W;
            const int rows = min(R, H - y0);
            // cout groups of this thread: NG consecutive groups (one column half).  (Squeeze modes: giving a thread both
            // horizontal phases of 8 source groups, so that its consecutive accesses share 32-byte sectors, was measured
            // 14 % slower than this mapping, where the two phases are fetched by the two column-half warps.)
            auto g_of = [&](int j) { return ct * (N / 4) + half * NG + j; };
            auto col_of = [&](int tc) { return half * (N / 2) + tc * CH; };
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + b * Cfg::ACC_COLS;
            {
                const bool lf = (x == 1), rt = (x == W - 2);  // this pixel also feeds border column -1 / W
                float4* outp = reinterpret_cast<float4*>(a.out) + (size_t)(y0 + 1) * Wp + (x + 1);
                const float4* resp = reinterpret_cast<const float4*>(a.res) + (size_t)(y0 + 1) * Wp + (x + 1);
                float4 res[R][NG];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int j = 0; j < NG; ++j)
                        if (res_sq) {
                            const int g = g_of(j), k = g >> lgq, gs = g & (gq - 1);
                            res[r][j] = (xin && r < rows) ? reinterpret_cast<const float4*>(a.res)[(size_t)gs * plane_h + (size_t)(2 * (y0 + r) + (k >> 1) + 1) * Wph +
                                                                                                    (2 * x + (k & 1) + 1)]
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
                            res[r][j] = (coupled && xin && r < rows) ? resp[(size_t)g_of(j) * plane + (size_t)r * Wp]
                                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                if (!SQZ && coupled && (lane & 7) == 0) {
                    // the coupling operand of this CTA's NEXT tile: pulled HBM -> L2 now (one 128-byte line per 8
                    // lanes), so that its register loads at the top of the next tile find it in L2 — the epilogue
                    // runs behind the UMMAs, so nothing else hides that latency
                    const int tn = t + (int)gridDim.x;
                    if (tn < tl.n_tiles) {
                        const int ctn = tn % tl.n_ct, restn = tn / tl.n_ct;
                        const int xn = (restn % tl.n_xt) * 128 + q * 32 + lane, yn = (restn / tl.n_xt) * R;
                        if (xn < W) {
                            const float4* rn = reinterpret_cast<const float4*>(a.res) + (size_t)(ctn * (N / 4) + half * NG) * plane +
                                               (size_t)(yn + 1) * Wp + (xn + 1);
#pragma unroll
                            for (int r = 0; r < R; ++r)
#pragma unroll
                                for (int j = 0; j < NG; ++j)
                                    if (yn + r < H) asm volatile("prefetch.global.L2 [%0];" ::"l"(rn + (size_t)j * plane + (size_t)r * Wp));
                        }
                    }
                }
                mbar_wait(&acc_full[b], (tcount >> 1) & 1);
                tc_fence_after();
                if (tid == 0) TC_TRACE(6, tcount);
                // TMEM -> registers is software-pipelined: the load of chunk t+1 is in flight while chunk t is
                // combined and stored (a tcgen05.ld takes ~200 cycles while the tensor pipe is busy)
                constexpr int NCHK = NG * 4 / CH;
                uint32_t ub[2][CH];
                tmem_ld_nowait<CH>(trow + (uint32_t)col_of(0), ub[0]);        // warp-collective
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < rows) {                                    // warp-uniform
                        const int y = y0 + r;
                        const bool up = (y == 1), dn = (y == H - 2);   // row also feeds border row -1 / H
#pragma unroll
                        for (int c0 = 0; c0 < NG * 4; c0 += CH) {
                            constexpr int dummy = 0; (void)dummy;
                            const int t = r * NCHK + c0 / CH;
                            tmem_ld_wait();
                            tmem_ld_fence_regs<CH>(ub[t & 1]);
                            float v[CH];
#pragma unroll
                            for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(ub[t & 1][i]);
                            if (t + 1 < R * NCHK && (c0 + CH < NG * 4 || r + 1 < rows))
                                tmem_ld_nowait<CH>(trow + (uint32_t)(((t + 1) / NCHK) * N + col_of((t + 1) % NCHK)), ub[(t + 1) & 1]);
                            if (xin) {
#pragma unroll
                                for (int jj = 0; jj < CH / 4; ++jj) {
                                    const int j = c0 / 4 + jj;
                                    const int g = g_of(j);
                                    const float4 bv = *reinterpret_cast<const float4*>(bias_s + 4 * g);
                                    float4 o = res[r][j];
                                    o.x += fmaxf((v[4 * jj] * unscale + bv.x) * sgn, flo);
                                    o.y += fmaxf((v[4 * jj + 1] * unscale + bv.y) * sgn, flo);
                                    o.z += fmaxf((v[4 * jj + 2] * unscale + bv.z) * sgn, flo);
                                    o.w += fmaxf((v[4 * jj + 3] * unscale + bv.w) * sgn, flo);
                                    if (out_us) {
                                        const int k = g >> lgq, gs = g & (gq - 1);
                                        p4_store(reinterpret_cast<float4*>(a.out) + (size_t)gs * plane_h, Hh, Wh, 2 * y + (k >> 1), 2 * x + (k & 1), o);
                                        continue;
                                    }
                                    float4* p = outp + (size_t)g * plane + (size_t)r * Wp;
                                    *p = o;
                                    if (lf) p[-2] = o;                 // reflection border, inline and predicated
                                    if (rt) p[2] = o;
                                    if (up) {
                                        float4* qq = p - 2 * (size_t)Wp;
                                        *qq = o;
                                        if (lf) qq[-2] = o;
                                        if (rt) qq[2] = o;
                                    }
                                    if (dn) {
                                        float4* qq = p + 2 * (size_t)Wp;
                                        *qq = o;
                                        if (lf) qq[-2] = o;
                                        if (rt) qq[2] = o;
                                    }
                                }
                            }
                        }
                    }
                    if (tid == 0) TC_TRACE(7, tcount * R + r);
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

static long long* g_tc_trace = nullptr;

// developer aid: VST_TC_TRACE="cin,cout" hands a zeroed stamp buffer to launches of that shape
long long* tc_trace_buffer(int Cin, int Cout, cudaStream_t st) {
    static int tr_cin = -1, tr_cout = -1;
    if (tr_cin < 0) {
        const char* e = getenv("VST_TC_TRACE");
        tr_cin = 0;
        if (e) sscanf(e, "%d,%d", &tr_cin, &tr_cout);
    }
    if (tr_cin != Cin || tr_cout != Cout) return nullptr;
    static long long* buf = nullptr;
    if (!buf) cudaMalloc(&buf, 8 * 4096 * sizeof(long long));
    cudaMemsetAsync(buf, 0, 8 * 4096 * sizeof(long long), st);
    g_tc_trace = buf;
    return buf;
}

template <int N, int R, int TERMS, bool HALF, bool SQZ, bool WST = false>
static int launch_tc_cfg2(const ConvArgs& a, cudaStream_t st);
template <int N, int R, int TERMS, bool HALF = false>
static int launch_tc_cfg(const ConvArgs& a, cudaStream_t st) {
    return a.epi <= EPI_SUB ? launch_tc_cfg2<N, R, TERMS, HALF, false>(a, st) : launch_tc_cfg2<N, R, TERMS, HALF, true>(a, st);
}
template <int N, int R, int TERMS, bool HALF, bool SQZ, bool WST>
static int launch_tc_cfg2(const ConvArgs& a, cudaStream_t st) {
    using Cfg = TcCfg<N, R, TERMS, HALF, WST>;
    static PerDeviceOnce smem_once;
    auto kern = conv3x3_tc_kernel<N, R, TERMS, HALF, SQZ, WST>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    TcTiles tl;
    tl.n_xt = cdiv(a.Wout, 128); tl.n_yt = cdiv(a.Hout, R); tl.n_ct = a.Cout / N;
    tl.n_tiles = tl.n_xt * tl.n_yt * tl.n_ct;
    tl.trace = (a.epi <= EPI_SUB) ? tc_trace_buffer(a.Cin, a.Cout, st) : nullptr;
    int grid = std::min(tl.n_tiles, num_sms());
    if (WST) grid -= grid % tl.n_ct;             // a CTA's tiles must all have the same cout tile
    char cls[40];
    snprintf(cls, sizeof(cls), HALF ? "conv3x3_tcH%d %d>%d%s" : "conv3x3_tc%d %d>%d%s", TERMS, a.Cin, a.Cout, SQZ ? " sqz" : "");
    const double px = (double)a.Hout * a.Wout;
    const bool coupled = a.epi >= EPI_ADD;
    ProfScope prof(st, cls, 2.0 * 9 * a.Cin * a.Cout * px,
                   4.0 * ((double)a.Cin * a.Hin * a.Win + (coupled ? 2.0 : 1.0) * a.Cout * px));
    VST_CUDA_OK(launch_pdl(kern, grid, TC_THREADS, Cfg::SMEM, st, a, tl));
    return check_launch("conv3x3_tc");
}

// cout tile width: 128 (R = 2 rows) where Cout allows — a K=8 UMMA costs the same ~66 cycles for N = 64 and
// N = 128 (it is bound by fetching A), so wider tiles halve the tensor time of the 64>256 coupling convs —
// else 64 (R = 4) or 16.  VST_TC_WIDE=0 forces the narrow tiles (developer knob).
static int tc_wide_n() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VST_TC_WIDE"); v = e ? atoi(e) : 128; }
    return v;
}
int tc_tile_n(int Cout) {
    if (tc_wide_n() >= 128 && Cout % 128 == 0) return 128;
    return (Cout % 64 == 0) ? 64 : 16;
}

bool tc_eligible(int Cin, int Cout, int stride) {
    return stride == 1 && Cin % 8 == 0 && Cin >= 16 && Cout <= 256 && (Cout % 64 == 0 || Cout == 16);
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

bool tc_half_eligible(int Cin, int Cout, int stride) { return tc_eligible(Cin, Cout, stride) && Cin % 16 == 0; }

// raw OIHW fp32 -> fp16 [cout tile][chunk16][tap][k-half (2)][n][8 halfs]   (one weight term: w = fp16(w))
__global__ void pack_tc_half_weights_kernel(const float* __restrict__ w, __half* __restrict__ wp, int Cin, int Cout, int N) {
    const size_t total = (size_t)(Cout / N) * (Cin / 16) * 9 * 2 * N * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = (int)(r % 8); r /= 8;
        const int n = (int)(r % N); r /= N;
        const int kh = (int)(r % 2); r /= 2;
        const int tap = (int)(r % 9); r /= 9;
        const int chunk = (int)(r % (Cin / 16)); r /= (Cin / 16);
        const int tile = (int)r;
        const int co = tile * N + n, ci = chunk * 16 + kh * 8 + e;
        wp[i] = __float2half_rn(w[((size_t)co * Cin + ci) * 9 + tap]);
    }
}

int launch_pack_tc_half_weights(const float* w, float* wp, int Cin, int Cout, int N, cudaStream_t st) {
    const size_t total = (size_t)(Cout / N) * (Cin / 16) * 9 * 2 * N * 8;
    pack_tc_half_weights_kernel<<<(int)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(
        w, reinterpret_cast<__half*>(wp), Cin, Cout, N);
    return check_launch("pack_tc_half_weights");
}

// a.in is an H8 split-half tensor; a.w packed by launch_pack_tc_half_weights with N = tc_tile_n(Cout)
int launch_conv3x3_tc_half(const ConvArgs& a, cudaStream_t st) {
    VST_REQUIRE(tc_half_eligible(a.Cin, a.Cout, 1), "conv3x3_tc_half: shape %d>%d not eligible", a.Cin, a.Cout);
    VST_REQUIRE(a.Hin == a.Hout && a.Win == a.Wout && a.Hin >= 2 && a.Win >= 2, "conv3x3_tc_half is stride 1, H,W >= 2");
    if (conv_pair_eligible(a)) return launch_conv3x3_pair(a, st);        // CTA-pair UMMAs (conv_pair.cu)
    const int N = tc_tile_n(a.Cout);
    static int wst = -1;
    if (wst < 0) { const char* e = getenv("VST_TC_WST"); wst = e ? atoi(e) : 1; }
    if (N == 128 && wst && a.Cin == 64 && (a.Cout / N) <= num_sms())      // resident weights (the 64 -> 256 coupling convs)
        return a.epi <= EPI_SUB ? launch_tc_cfg2<128, 2, 2, true, false, true>(a, st) : launch_tc_cfg2<128, 2, 2, true, true, true>(a, st);
    if (N == 128) return launch_tc_cfg<128, 2, 2, true>(a, st);
    if (N == 64) return launch_tc_cfg<64, 4, 2, true>(a, st);
    return launch_tc_cfg<16, 4, 2, true>(a, st);
}

}  // namespace vst

// developer hook (not declared in the public header): copy the last traced launch's stamps to the host
extern "C" int vst_debug_tc_trace(long long* host, int n) {
    if (!vst::g_tc_trace) return 1;
    cudaDeviceSynchronize();
    return cudaMemcpy(host, vst::g_tc_trace, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
