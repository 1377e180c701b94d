// conv_tch.cu — the kx-folded tcgen05 convolution of conv_tcx.cu with HALF-PRECISION SPLIT operands:
//     x = hi + lo,  hi = fp16(x), lo = fp16(x - hi)   (22 significand bits, like the tf32 hi/lo split)
//     w = fp16(w)                                     (11 significand bits, like tf32)
//     D += hi * w ; D += lo * w                       (kind::f16 UMMA, fp32 accumulate in TMEM)
// A kind::f16 UMMA consumes K = 16 channels for the same ~66-cycle operand-fetch-bound cost at which a
// kind::tf32 UMMA consumes K = 8 (profiles/), and its operands are half as many bytes in shared memory
// and in L2 (weights), so the tensor time and the weight traffic of these layers halve at equal accuracy.
// Range: activations and weights of this network are O(1e-4 .. 1e2); fp16 subnormals (< 6e-5) lose
// relative precision but stay within 3e-8 absolute, far below the fp32 noise of the outputs.
//
// Pipeline (per 16-channel chunk): TMA lands the raw fp32 P4 rows in a RAW ring; converter warps turn
// them into the canonical K-major fp16 operand layout [k-half][row][pixel][8 halfs] (hi and lo) inside
// an OPERAND ring slot, where a second TMA producer drops the chunk's fp16 weights; the UMMA issuer
// consumes operand slots.  Two rings decouple the L2 latency (deep RAW ring) from operand storage.
// Epilogue, tiling and the kx fold are those of conv_tcx.cu.
#include <stdlib.h>
#include "kernels.cuh"
#include <cuda_fp16.h>
#include "tma_map.cuh"

namespace vst {

// Tile geometry (round 2): the 128 rows of a UMMA are FOUR IMAGE ROWS x 32 staged pixels (30 outputs + 2 halo pixels),
// stacked vertically, instead of four 32-pixel windows of ONE image row.  Operand rows live as [k-half][image row][32 px]
// [16 B] (row pitch 512 B), so the rows y..y+3 of a block are 128 consecutive "pixels" and tap ky is a start-address
// shift of 512 B; the kx fold of the epilogue still never leaves a warp's 32 TMEM lanes (warp q = image row q of the
// block).  A tile of R blocks is 4R rows x 30 pixels and stages (4R + 2) x 32 pixels: 1.33x its outputs (R = 2) instead
// of 2.03x for two full-width rows — the L2 -> SM operand traffic per output pixel and 16-channel chunk drops from
// 207 B to 158 B (activations 130 -> 85 B, weights 77 B), which is what bounded the 256 -> 64 conv.
// The RAW tile of a chunk (4 groups x (4R + 2) rows x 32 pixels) arrives as ONE tensor-map TMA copy (3-D box).
template <int NC, int R, int TERMS>
struct TchCfg {
    static constexpr int NP = 3 * NC;               // UMMA N: (kx, cout)
    static constexpr int TA = TERMS >= 2 ? 2 : 1;   // activation terms (hi [, lo])
    static constexpr int TW = TERMS >= 3 ? 2 : 1;   // weight terms (hi [, lo])
    static constexpr int PW = 128;                  // UMMA M: 4 image rows x WIN staged pixels
    static constexpr int WIN = 32;                  // staged pixels per image row
    static constexpr int XS = 30;                   // outputs per tile row
    static constexpr int TR = 4 * R;                // output rows per tile
    static constexpr int ROWS = TR + 2;             // staged image rows
    static constexpr int ROW_BYTES = WIN * 16;      // one image row of one 8-channel fp16 k-half (operand layout)
    static constexpr int RAW_ROW_BYTES = WIN * 16;                  // one image row of one 4-channel fp32 group
    static constexpr int RAW_BYTES = 4 * ROWS * RAW_ROW_BYTES;      // 16 channels fp32: [group][row][pixel][4 floats]
    static constexpr int A_TERM_BYTES = 2 * ROWS * ROW_BYTES;       // 16 channels fp16: [k-half][row][pixel][8 halfs]
    static constexpr int A_BYTES = TA * A_TERM_BYTES;
    static constexpr int B_TERM_BYTES = 3 * 2 * NP * 16;            // [ky][k-half][n'][8 halfs]
    static constexpr int B_BYTES = TW * B_TERM_BYTES;
    static constexpr int OP_BYTES = A_BYTES + B_BYTES;
    static constexpr int ACC_COLS = R * NP;
    static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (NACC * ACC_COLS <= 128) ? 128 : (NACC * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int AUX_BYTES = 2048;                          // barriers (1 KB) + bias (1 KB)
    // operand ring depth.  The chunk's weights (18 KB at NC = 64) ride in the operand slot and take 3-4 k cycles to land,
    // like the raw box: with three slots the few-chunk R = 1 convs ran at fill latency / 3 = ~1.2 k cycles per chunk
    static constexpr int NO = (R == 1 && NC == 64 && TERMS <= 2) ? 5 : 3;
    // converter teams.  A chunk's conversion is a latency chain (LDS -> cvt -> STS, ~1 k cycles with the barrier
    // traffic around it) however few items a thread has; the few-chunk R = 1 convs were paced by it (role trace: ~1.2 k
    // cycles per chunk vs ~400 of UMMAs).  Two teams of four warps convert alternate chunks concurrently.
    static constexpr int TEAMS = (R == 1) ? 2 : 1;
    // raw ring depth (TMA boxes in flight).  A box takes 3-4 k cycles to land under load (role trace); the few-chunk
    // R = 1 convs spend only ~400 cycles of UMMAs per chunk, so four boxes in flight paced them at ~1.2 k cycles per chunk
    static constexpr int NR_MAX = (R == 1) ? 8 : 4;
    static constexpr int NR_FIT = (226 * 1024 - AUX_BYTES - NO * OP_BYTES) / RAW_BYTES;
    static constexpr int NR = NR_FIT > NR_MAX ? NR_MAX : NR_FIT;
    static constexpr size_t SMEM = (size_t)NR * RAW_BYTES + (size_t)NO * OP_BYTES + AUX_BYTES + 128;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(NP % 16 == 0 && NP <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
    static_assert(NR >= 2, "need at least a double-buffered raw ring");
};

// raw OIHW fp32 -> fp16 [cout tile][chunk16][term][ky][k-half (2)][n' = kx*NC + co][8 halfs]
__global__ void pack_tch_weights_kernel(const float* __restrict__ w, __half* __restrict__ wp, int Cin, int Cout, int NC,
                                        int TW) {
    const int NP = 3 * NC;
    const size_t total = (size_t)(Cout / NC) * (Cin / 16) * TW * 3 * 2 * NP * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = (int)(r % 8); r /= 8;
        const int n = (int)(r % NP); r /= NP;
        const int kh = (int)(r % 2); r /= 2;
        const int ky = (int)(r % 3); r /= 3;
        const int term = (int)(r % TW); r /= TW;
        const int chunk = (int)(r % (Cin / 16)); r /= (Cin / 16);
        const int tile = (int)r;
        const int kx = n / NC, co = tile * NC + (n - kx * NC), ci = chunk * 16 + kh * 8 + e;
        const float v = w[((size_t)co * Cin + ci) * 9 + ky * 3 + kx];
        const __half hi = __float2half_rn(v);
        wp[i] = term == 0 ? hi : __float2half_rn(v - __half2float(hi));
    }
}

int launch_pack_tch_weights(const float* w, float* wp, int Cin, int Cout, int NC, int terms, cudaStream_t st) {
    const int TW = terms >= 3 ? 2 : 1;
    const size_t total = (size_t)(Cout / NC) * (Cin / 16) * TW * 3 * 2 * (3 * NC) * 8;
    pack_tch_weights_kernel<<<(int)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(
        w, reinterpret_cast<__half*>(wp), Cin, Cout, NC, TW);
    return check_launch("pack_tch_weights");
}

// The stride-2 conv of a transition block (RevResNet.py:79-81 with stride 2) as a stride-1 conv on the
// SQUEEZED input: with xs[(dy*2+dx)*C + ci][h][w] = x[ci][2h+dy][2w+dx] (RevResNet.py:34-37),
//     out(h,w) = sum_{ky,kx,ci} W[ky][kx][ci] x[ci][2h+ky-1][2w+kx-1]
//              = sum over taps (ky',kx') in {0,1}^2 (offsets -1, 0) and squeezed channels (dy,dx,ci) of
//                W[ky][kx][ci] xs[(dy,dx,ci)][h+ky'-1][w+kx'-1]   with  (ky',dy) -> ky: (0,1)->0, (1,0)->1, (1,1)->2
// (same for x); every other (tap, sub-position) pair gets a zero weight.  4/9 of the MACs are on zeros, but
// they run on the tensor cores in the kx-folded kernel above instead of on CUDA cores.  The reflection row
// x[-1] = x[1] is xs[dy=1][h=0], i.e. the squeezed tensor needs its row 0 / column 0 REPLICATED into the
// border (p4_replicate_topleft); the +1 taps have zero weights, so the other borders do not matter.
// Output layout is that of pack_tch_weights_kernel with Cin' = 4C.
__global__ void pack_tch_s2_weights_kernel(const float* __restrict__ w, __half* __restrict__ wp, int C, int Cout) {
    const int NC = Cout, NP = 3 * NC, Cs = 4 * C;
    const size_t total = (size_t)(Cs / 16) * 3 * 2 * NP * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = (int)(r % 8); r /= 8;
        const int n = (int)(r % NP); r /= NP;
        const int kh = (int)(r % 2); r /= 2;
        const int kyp = (int)(r % 3); r /= 3;
        const int chunk = (int)r;
        const int kxp = n / NC, co = n - kxp * NC, cs = chunk * 16 + kh * 8 + e;
        const int k = cs / C, ci = cs - k * C, dy = k >> 1, dx = k & 1;
        const int ky = (kyp == 0 && dy == 1) ? 0 : (kyp == 1 ? 1 + dy : -1);
        const int kx = (kxp == 0 && dx == 1) ? 0 : (kxp == 1 ? 1 + dx : -1);
        const float v = (ky >= 0 && kx >= 0) ? w[((size_t)co * C + ci) * 9 + ky * 3 + kx] : 0.f;
        wp[i] = __float2half_rn(v);
    }
}

int launch_pack_tch_s2_weights(const float* w, float* wp, int C, int Cout, cudaStream_t st) {
    const size_t total = (size_t)(4 * C / 16) * 3 * 2 * (3 * Cout) * 8;
    pack_tch_s2_weights_kernel<<<(int)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(
        w, reinterpret_cast<__half*>(wp), C, Cout);
    return check_launch("pack_tch_s2_weights");
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 inputs, fp32 accumulate), M=128, K=16
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

struct TchTiles {
    int n_xt, n_yt, n_ct, n_tiles;
    long long* trace;   // developer aid (VST_TC_TRACE): CTA 0 stamps clock64() per role event, 4096 slots per role
};
#define TCH_TRACE(role, idx) do { if (tl.trace && blockIdx.x == 0 && (idx) < 4096) tl.trace[(role) * 4096 + (idx)] = clock64(); } while (0)

constexpr int TCH_NCW = 8;         // converter warps: with 4 the fp32 -> fp16 hi/lo conversion paces the MMA phase (traces)
constexpr int TCH_THREADS = (8 + TCH_NCW + 3) * 32;   // 8 epilogue + converter warps, activation producer, UMMA issuer, weight producer

template <int NC, int R, int TERMS, bool SPLIT>   // SPLIT: H8 split-half output (compile time: each variant keeps its own registers)
__global__ void __launch_bounds__(TCH_THREADS, 1) conv3x3_tch_kernel(ConvArgs a, TchTiles tl,
                                                                     const __grid_constant__ CUtensorMap tm_in) {
    using Cfg = TchCfg<NC, R, TERMS>;
    constexpr int NR = Cfg::NR, NO = Cfg::NO, WIN = Cfg::WIN, ROWS = Cfg::ROWS, NP = Cfg::NP, NACC = Cfg::NACC, XS = Cfg::XS;
    constexpr int TR = Cfg::TR;
    pdl_launch_dependents();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* raw_base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // by offset: keeps the __shared__ address space (LDS / STS)
    uint8_t* op_base = raw_base + (size_t)NR * Cfg::RAW_BYTES;
    uint64_t* bars = (uint64_t*)(op_base + (size_t)NO * Cfg::OP_BYTES);
    uint64_t* raw_full = bars;                  // [NR <= 8]  activation producer arrive.expect_tx + TMA bytes
    uint64_t* raw_empty = bars + 8;             // [NR <= 8]  converter threads
    uint64_t* op_ready = bars + 16;             // [NO <= 8]  converter threads + weight producer arrive.expect_tx + TMA bytes
    uint64_t* op_empty = bars + 24;             // [NO <= 8]  tcgen05.commit
    uint64_t* acc_full = bars + 32;             // [NACC]  tcgen05.commit
    uint64_t* acc_empty = bars + 34;            // [NACC]  256 epilogue threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 36);
    float* bias_s = (float*)((uint8_t*)bars + 1024);
    for (int i = threadIdx.x; i < a.Cout && i < 256; i += blockDim.x) bias_s[i] = __ldg(a.bias + i);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = a.Cin / 16;

    if (tid == 0) {
        for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 32 * TCH_NCW / Cfg::TEAMS); }
        for (int s = 0; s < NO; ++s) { mbar_init(&op_ready[s], 32 * TCH_NCW / Cfg::TEAMS + 1); mbar_init(&op_empty[s], 1); }
        for (int b = 0; b < NACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
        fence_barrier_init();
    }
    constexpr int W_PROD = 8 + TCH_NCW, W_MMA = W_PROD + 1, W_WGT = W_PROD + 2;
    if (warp == W_MMA) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                          // prologue above touches only constants (bias) and on-chip state

    if (warp == W_PROD) {
        // ================= activation producer (TMA): raw fp32 P4 rows of 16 channels -> RAW ring =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
                const int rest = t / tl.n_ct;
                const int xs = (rest % tl.n_xt) * XS, y0 = (rest / tl.n_xt) * TR;
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int s = it % NR;
                    mbar_wait(&raw_empty[s], ((it / NR) & 1) ^ 1);
                    TCH_TRACE(0, it);
                    uint8_t* A = raw_base + (size_t)s * Cfg::RAW_BYTES;
                    mbar_arrive_expect_tx(&raw_full[s], Cfg::RAW_BYTES);
                    // box {32 pixels x 4 floats, ROWS rows, 4 groups}: tile pixel p <-> padded column xs + p (image
                    // x = xs - 1 + p), staged row r <-> padded row y0 + r (image row y0 - 1 + r); rows / columns
                    // past the padded plane arrive as zeros and only feed outputs that are never stored
                    tma_load_3d(A, &tm_in, xs * 4, y0, 4 * c, &raw_full[s]);
                }
            }
        }
    } else if (warp == W_WGT) {
        // ================= weight producer (TMA): fp16 weights of the chunk -> operand slot =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
                const int ct = t % tl.n_ct;
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + (size_t)ct * n_chunks * Cfg::B_BYTES;
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int o = it % NO;
                    mbar_wait(&op_empty[o], ((it / NO) & 1) ^ 1);
                    uint8_t* B = op_base + (size_t)o * Cfg::OP_BYTES + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(&op_ready[o], Cfg::B_BYTES);
                    bulk_g2s(B, wsrc + (size_t)c * Cfg::B_BYTES, Cfg::B_BYTES, &op_ready[o]);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================= UMMA issuer =================
        // the whole warp walks the warp-uniform schedule, one elected lane issues (see conv_tc.cu); descriptors are
        // constant high bits + (shared address >> 4)
        {
            // kind::f16: D = F32 (bit 4), A = B = F16 (format 0), N at bit 17, M at bit 24
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t A_LBO = ROWS * Cfg::ROW_BYTES, B_LBO = NP * 16, SBO = 128;     // A: 4 image rows = 128 contiguous operand rows
            uint32_t elp;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elp));
            const bool el = elp != 0;
            const uint64_t dA0 = make_desc(smem_u32(op_base), A_LBO, SBO);
            const uint64_t dB0 = make_desc(smem_u32(op_base) + Cfg::A_BYTES, B_LBO, SBO);
            uint32_t it = 0, tcount = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
                const uint32_t b = tcount % NACC;
                mbar_wait(&acc_empty[b], ((tcount / NACC) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) TCH_TRACE(5, tcount);
                const uint32_t acc = tmem_base + b * Cfg::ACC_COLS;
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int o = it % NO;
                    mbar_wait(&op_ready[o], (it / NO) & 1);
                    tc_fence_after();
                    if (lane == 0) TCH_TRACE(3, it);
                    const uint64_t so = (uint64_t)((uint32_t)o * (Cfg::OP_BYTES >> 4));
                    const uint64_t dA = dA0 + so, dB = dB0 + so;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint64_t bh = dB + (uint64_t)((ky * 2 * NP * 16) >> 4);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint64_t ah = dA + (uint64_t)(((4 * r + ky) * Cfg::ROW_BYTES) >> 4);
                            const uint32_t d = acc + r * NP;
                            const uint32_t first = (c > 0 || ky > 0) ? 1u : 0u;
                            if (el) {
                                umma_f16(d, ah, bh, IDESC, first);
                                if (TERMS >= 2) umma_f16(d, ah + (uint64_t)(Cfg::A_TERM_BYTES >> 4), bh, IDESC, 1u);
                                if (TERMS >= 3) umma_f16(d, ah, bh + (uint64_t)(Cfg::B_TERM_BYTES >> 4), IDESC, 1u);
                            }
                        }
                    }
                    if (el) umma_commit(&op_empty[o]);
                    if (lane == 0) TCH_TRACE(4, it);
                }
                if (el) umma_commit(&acc_full[b]);
            }
        }
        __syncwarp();
    } else if (warp >= 8 && warp < 8 + TCH_NCW) {
        // ================= converters: raw fp32 -> fp16 hi / lo in the K-major operand layout =================
        // one item = 8 channels (two P4 groups) of one pixel: 32 bytes in, 16 (hi) + 16 (lo) bytes out
        constexpr int TEAMS = Cfg::TEAMS, TEAM_THREADS = 32 * TCH_NCW / TEAMS;
        const int team = (tid - 256) / TEAM_THREADS, ctid = (tid - 256) % TEAM_THREADS;
        uint32_t it = 0;
        uint32_t hmax = 0u;                    // fp16 range guard (tc_ptx.cuh)
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c, ++it) {
                if (TEAMS > 1 && (int)(it % TEAMS) != team) continue;      // the other team's chunk
                const int s = it % NR, o = it % NO;
                mbar_wait(&raw_full[s], (it / NR) & 1);
                mbar_wait(&op_empty[o], ((it / NO) & 1) ^ 1);          // the UMMAs that read this slot are done
                if (tid == 256) TCH_TRACE(1, it);
                const float4* raw = reinterpret_cast<const float4*>(raw_base + (size_t)s * Cfg::RAW_BYTES);
                uint4* hi = reinterpret_cast<uint4*>(op_base + (size_t)o * Cfg::OP_BYTES);
                uint4* lo = reinterpret_cast<uint4*>(op_base + (size_t)o * Cfg::OP_BYTES + Cfg::A_TERM_BYTES);
#pragma unroll 4
                for (int i = ctid; i < 2 * ROWS * WIN; i += TEAM_THREADS) {
                    const int kh = i / (ROWS * WIN), rp = i - kh * (ROWS * WIN);    // k-half, (image row, pixel): same index in RAW
                    const float4 u = raw[(2 * kh) * (ROWS * WIN) + rp];
                    const float4 v = raw[(2 * kh + 1) * (ROWS * WIN) + rp];
                    const float x[8] = {u.x * VST_HALF_SCALE, u.y * VST_HALF_SCALE, u.z * VST_HALF_SCALE, u.w * VST_HALF_SCALE,
                                        v.x * VST_HALF_SCALE, v.y * VST_HALF_SCALE, v.z * VST_HALF_SCALE, v.w * VST_HALF_SCALE};
                    uint32_t hw[4];
                    float l[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 hh = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
                        const float2 hf = __half22float2(hh);
                        hw[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        hmax = range_fold(hmax, hw[e]);
                        l[2 * e] = x[2 * e] - hf.x;
                        l[2 * e + 1] = x[2 * e + 1] - hf.y;
                    }
                    hi[i] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    if (Cfg::TA == 2)
                        lo[i] = make_uint4(pack_half2(l[0], l[1]), pack_half2(l[2], l[3]), pack_half2(l[4], l[5]), pack_half2(l[6], l[7]));
                }
                fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core
                mbar_arrive(&op_ready[o]);
                mbar_arrive(&raw_empty[s]);
                if (tid == 256) TCH_TRACE(2, it);
            }
        }
        range_report(hmax, a.status);
    } else if (warp < 8) {
        // ================= epilogue: TMEM -> (lane-shifted sum over kx) -> ReLU -> P4 global =================
        // lane l of warp (q, half) owns tile pixel m = 32q + l and the cout half `half`.
        constexpr int HC = NC / 2;                         // couts per thread
        constexpr int CH = HC >= 16 ? 16 : HC;             // couts per TMEM load
        const int q = warp & 3, half = warp >> 2;          // q: image row of the block held by this warp's 32 TMEM lanes
        const float flo = (a.epi == EPI_RELU) ? 0.f : -INFINITY;
        const int H = a.Hout, W = a.Wout, Wp = W + 2;
        const size_t plane = p4_plane_px(H, W);
        uint32_t tcount = 0;
        uint32_t hmax = 0u;                    // fp16 range guard of the H8 output (tc_ptx.cuh)
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
            const int ct = t % tl.n_ct, rest = t / tl.n_ct;
            const int xs = (rest % tl.n_xt) * XS, y0 = (rest / tl.n_xt) * TR;
            const uint32_t b = tcount % NACC;
            const int x = xs - 1 + lane;
            const bool xok = (lane >= 1) && (lane <= 30) && (x < W);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + b * Cfg::ACC_COLS + half * HC;
            mbar_wait(&acc_full[b], (tcount / NACC) & 1);
            tc_fence_after();
            if (tid == 0) TCH_TRACE(6, tcount);
            // out[m] = D[m-1][kx=0] + D[m][kx=1] + D[m+1][kx=2], all within the warp
            const bool lf = xok && (x == 1), rt = xok && (x == W - 2);
            const int r_last = min(R, (H - y0 + 3) / 4) - 1;        // last block of this tile with rows inside the image
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const int y = y0 + 4 * r + q;
                if (y0 + 4 * r >= H) break;                       // (warp-uniform) no rows of this block inside the image
                const bool xin = xok && (y < H);
                const bool up = (y == 1), dn = (y == H - 2);
#pragma unroll 1
                for (int c0 = 0; c0 < HC; c0 += CH) {
                    float v0[CH], v1[CH], v2[CH];
                    tmem_ld3<CH>(trow + (uint32_t)(r * NP + 0 * NC + c0), v0, trow + (uint32_t)(r * NP + 1 * NC + c0), v1,
                                 trow + (uint32_t)(r * NP + 2 * NC + c0), v2);
                    if (r == r_last && c0 + CH >= HC) {
                        // the tile's last TMEM read has completed: the accumulator goes back to the issuer now, the fold,
                        // conversion and stores below run beside the next tile's UMMAs
                        tc_fence_before();
                        mbar_arrive(&acc_empty[b]);
                    }
                    if (tid == 0 && tcount == 1) TCH_TRACE(7, 16 + (r * (HC / CH) + c0 / CH) * 3 + 0);
                    const int cb = half * HC + c0;                  // first cout of this chunk (within the tile)
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const float l = __shfl_up_sync(0xffffffffu, v0[i], 1);
                        const float rr = __shfl_down_sync(0xffffffffu, v2[i], 1);
                        v1[i] = ((l + v1[i]) + rr) * (1.0f / VST_HALF_SCALE);
                    }
                    if (tid == 0 && tcount == 1) TCH_TRACE(7, 16 + (r * (HC / CH) + c0 / CH) * 3 + 1);
                    if (SPLIT && xin) {
                        // H8 split-half output (feeds a kind::f16 conv): 8 channels = one 16-byte unit of hi and of lo
#pragma unroll
                        for (int j = 0; j < CH / 8; ++j) {
                            const int g8 = (ct * NC + cb) / 8 + j;
                            float o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                o[e] = fmaxf(v1[8 * j + e] + bias_s[ct * NC + cb + 8 * j + e], flo) * VST_HALF_SCALE;
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __half2 hh = __floats2half2_rn(o[2 * e], o[2 * e + 1]);
                                const float2 hf = __half22float2(hh);
                                hw[e] = *reinterpret_cast<const uint32_t*>(&hh);
                                hmax = range_fold(hmax, hw[e]);
                                lw[e] = pack_half2(o[2 * e] - hf.x, o[2 * e + 1] - hf.y);
                            }
                            const uint4 hv = make_uint4(hw[0], hw[1], hw[2], hw[3]), lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            const size_t lo_off = (size_t)(a.Cout / 8) * plane;          // lo planes follow the hi planes
                            uint4* p = reinterpret_cast<uint4*>(a.out) + (size_t)g8 * plane + (size_t)(y + 1) * Wp + (x + 1);
#pragma unroll
                            for (int part = 0; part < 2; ++part) {
                                uint4* pp = p + (part ? lo_off : 0);
                                const uint4 val = part ? lv : hv;
                                *pp = val;
                                if (lf) pp[-2] = val;
                                if (rt) pp[2] = val;
                                if (up) {
                                    uint4* qq = pp - 2 * (size_t)Wp;
                                    *qq = val;
                                    if (lf) qq[-2] = val;
                                    if (rt) qq[2] = val;
                                }
                                if (dn) {
                                    uint4* qq = pp + 2 * (size_t)Wp;
                                    *qq = val;
                                    if (lf) qq[-2] = val;
                                    if (rt) qq[2] = val;
                                }
                            }
                        }
                    } else if (!SPLIT && xin) {
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) {
                            const int g = (ct * NC + cb) / 4 + j;
                            const float4 bv = *reinterpret_cast<const float4*>(bias_s + 4 * g);
                            float4 o;
                            o.x = fmaxf(v1[4 * j] + bv.x, flo);
                            o.y = fmaxf(v1[4 * j + 1] + bv.y, flo);
                            o.z = fmaxf(v1[4 * j + 2] + bv.z, flo);
                            o.w = fmaxf(v1[4 * j + 3] + bv.w, flo);
                            float4* p = reinterpret_cast<float4*>(a.out) + (size_t)g * plane + (size_t)(y + 1) * Wp + (x + 1);
                            *p = o;
                            if (lf) p[-2] = o;                     // reflection border, inline and predicated
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
                    if (tid == 0 && tcount == 1) TCH_TRACE(7, 16 + (r * (HC / CH) + c0 / CH) * 3 + 2);
                }
            }
            if (tid == 0) TCH_TRACE(7, 2 * tcount + 1);
        }
        if (SPLIT) range_report(hmax, a.status);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int NC, int R, int TERMS, bool SPLIT>
static int launch_tch_cfg2(const ConvArgs& a, cudaStream_t st);
template <int NC, int R, int TERMS>
static int launch_tch_cfg(const ConvArgs& a, cudaStream_t st) {
    return a.out_split ? launch_tch_cfg2<NC, R, TERMS, true>(a, st) : launch_tch_cfg2<NC, R, TERMS, false>(a, st);
}
template <int NC, int R, int TERMS, bool SPLIT>
static int launch_tch_cfg2(const ConvArgs& a, cudaStream_t st) {
    using Cfg = TchCfg<NC, R, TERMS>;
    static PerDeviceOnce smem_once;
    auto kern = conv3x3_tch_kernel<NC, R, TERMS, SPLIT>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    TchTiles tl;
    tl.n_xt = cdiv(a.Wout, Cfg::XS); tl.n_yt = cdiv(a.Hout, Cfg::TR); tl.n_ct = a.Cout / NC;
    tl.n_tiles = tl.n_xt * tl.n_yt * tl.n_ct;
    tl.trace = tc_trace_buffer(a.Cin, a.Cout, st);
    const int grid = std::min(tl.n_tiles, num_sms());
    // the input as a rank-3 fp32 tensor {(W+2) * 4 floats, H+2 rows, Cin/4 groups}; one box = the RAW tile of a chunk
    CUtensorMap tm;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)(a.Win + 2) * 4, (cuuint64_t)(a.Hin + 2), (cuuint64_t)(a.Cin / 4)};
        const cuuint64_t strides[2] = {(cuuint64_t)(a.Win + 2) * 16, (cuuint64_t)(a.Hin + 2) * (a.Win + 2) * 16};
        const cuuint32_t box[3] = {(cuuint32_t)Cfg::WIN * 4, (cuuint32_t)Cfg::ROWS, 4};
        if (make_tensor_map_f32(&tm, 3, a.in, dims, strides, box)) return 2;
    }
    char cls[40];
    snprintf(cls, sizeof(cls), "conv3x3_tch%d %d>%d", TERMS, a.Cin, a.Cout);
    const double px = (double)a.Hout * a.Wout;
    ProfScope prof(st, cls, 2.0 * 9 * a.Cin * a.Cout * px, 4.0 * ((double)a.Cin * a.Hin * a.Win + a.Cout * px));
    VST_CUDA_OK(launch_pdl(kern, grid, TCH_THREADS, Cfg::SMEM, st, a, tl, tm));
    return check_launch("conv3x3_tch");
}

bool tch_eligible(int Cin, int Cout, int stride) {
    return stride == 1 && Cin % 16 == 0 && (Cout == 64 || Cout == 16);
}

// a.w must point at weights packed by launch_pack_tch_weights with NC = Cout and `terms`; epi is RELU or NONE
int launch_conv3x3_tch(const ConvArgs& a, int terms, cudaStream_t st) {
    VST_REQUIRE(tch_eligible(a.Cin, a.Cout, 1), "conv3x3_tch: shape %d>%d not eligible", a.Cin, a.Cout);
    VST_REQUIRE(a.epi == EPI_RELU || a.epi == EPI_NONE, "conv3x3_tch has no coupling epilogue");
    VST_REQUIRE(a.Hin == a.Hout && a.Win == a.Wout && a.Hin >= 2 && a.Win >= 2, "conv3x3_tch is stride 1, H,W >= 2");
    if (a.Cout == 64) {
        // Few input chunks (Cin <= 64: the 64 -> 64 convs): a tile's UMMA phase (4 chunks, ~5 k cycles) is as short as
        // its epilogue (~5 k cycles, role traces), so the two must overlap — one 4-row block per tile (R = 1) leaves room
        // for a double-buffered accumulator (2 x 192 TMEM columns).  Cin = 256 keeps two blocks per tile (R = 2, single
        // accumulator): its 16-chunk UMMA phase dominates and the larger tile halves the weight traffic from L2.
        // (Releasing the accumulator block by block before the store phase was measured: +7 % on the 256 -> 64 conv.)
        static int r1 = -1;
        if (r1 < 0) { const char* e = getenv("VST_TCH_R1"); r1 = e ? atoi(e) : 1; }
        if (a.Cin <= 64 && r1) {
            if (terms == 1) return launch_tch_cfg<64, 1, 1>(a, st);
            if (terms == 2) return launch_tch_cfg<64, 1, 2>(a, st);
            return launch_tch_cfg<64, 1, 3>(a, st);
        }
        if (terms == 1) return launch_tch_cfg<64, 2, 1>(a, st);
        if (terms == 2) return launch_tch_cfg<64, 2, 2>(a, st);
        return launch_tch_cfg<64, 2, 3>(a, st);
    }
    if (terms == 1) return launch_tch_cfg<16, 4, 1>(a, st);
    if (terms == 2) return launch_tch_cfg<16, 4, 2>(a, st);
    return launch_tch_cfg<16, 4, 3>(a, st);
}

}  // namespace vst
