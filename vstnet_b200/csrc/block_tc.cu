// block_tc.cu — a whole stride-1 reversible block on the tensor cores, as ONE row-streaming kernel:
//     out = res +/- conv3(relu(conv2(relu(conv1(x)))))        C -> C/4 -> C/4 -> C channels, C = 16 or 64
// (models/RevResNet.py:79-88 the three ReflectionPad2d(1)+Conv2d(3x3), :96-104 forward coupling, :106-116 inverse).
// The quarter-width intermediates never leave the SM; HBM sees exactly: read x (+ halo rows), read res, write out.
//
// Geometry.  A CTA owns a vertical strip of NB x 128 staged pixels (NB = 2 for C = 16; 122 output columns per block:
// three 3x3 convs eat 3 columns per side) and a segment of output rows [ya, yb); it streams DOWN the strip one image row per step, so there is no
// halo recompute in y except 3 rows at each end of a segment.  Tile pixel m <-> image column x0 - 3 + m.
//
// GEMM view (kx folded into N as in conv_tch.cu; every UMMA is M = 128 pixels of one row, kind::f16, K = 16):
//     D[m, (kx, co)] += A[m, (row, ci)] * W[ky][ci, (kx, co)],       conv(x)[m] = D[m-1, kx=0] + D[m, kx=1] + D[m+1, kx=2]
//   conv1 is INPUT-stationary: when x row r has been converted it is multiplied into the three t1 rows r+1, r, r-1
//     (ky = 0, 1, 2) which accumulate in a 4-slot TMEM ring — an x row is staged once and freed at once.
//   conv2 / conv3 are OUTPUT-stationary over 4-slot shared-memory rings of t1 / t2 rows; the ReflectionPad2d rows
//     of the intermediates (row -1 = row 1, row H = row H-2) are a choice of ring slot, the padded columns are two
//     extra 16-byte stores by the thread that owns column 1 / W-2.
// Arithmetic is that of the f16x2 mode (conv_tch.cu): activations 64*x = hi + lo in fp16, weights fp16, fp32 accumulate.
// For C = 16 the 4-channel intermediates pack hi|lo into one K = 16 operand (weights duplicated), halving their UMMAs.
//
// Warp roles (C = 64: 672 threads; C = 16 has 16 E3 warps, 928 threads): 0-7 E3 (acc3 -> +bias, coupling with res -> global P4 stores), 8-11 E1 (acc1 -> ReLU ->
// t1 ring), 12-15 E2 (acc2 -> ReLU -> t2 ring), 16-19 converters (global fp32 P4 rows -> fp16 hi/lo operand rows),
// 20 weights TMA + UMMA issuer + TMEM owner.  All hand-offs are mbarriers; waits are bounded (tc_ptx.cuh).  Per step the
// issuer runs conv3(y = r-7), conv2(i = r-4), conv1(x row r) in that order, which makes the "accumulator drained"
// hand-offs of acc1 / acc2 implicit in the t-row barriers it waits for anyway.  Launched with programmatic dependent
// launch: everything up to pdl_wait() (rings zeroed, barriers, TMEM, weight copy) overlaps the previous kernel's tail.
#include <stdlib.h>
#include "kernels.cuh"
#include <cuda_fp16.h>
#include "tc_ptx.cuh"

namespace vst {

template <int C>
struct BtcCfg {
    static constexpr int M = C / 4;                      // bottleneck channels
    static constexpr int G = C / 4;                      // P4 groups of the half-state
    static constexpr int KS = C / 16;                    // conv1 K steps per ky
    static constexpr int N1 = (3 * M < 16) ? 16 : 3 * M; // UMMA N of conv1 / conv2: (kx, co), padded to 16
    static constexpr int N3 = 3 * C;                     // UMMA N of conv3
    static constexpr int TT = (M == 4) ? 1 : 2;          // operand terms of a t row (M = 4: hi|lo packed into K)
    static constexpr int NB = (C == 16) ? 2 : 1;         // 128-pixel blocks per CTA step (amortises the per-step barrier traffic)
    static constexpr int PW = 128, XO = 120;               // 4 windows of 30 outputs (btc_win_rows)
    static constexpr int CHUNK = PW * 16;                // one 8-half K chunk of one row: [pixel][16 B]
    static constexpr int XT = (C / 8) * CHUNK;           // one term of an x row: [C/8 chunks][pixel][16 B]
    static constexpr int X_BLK = 2 * XT;
    static constexpr int X_SLOT = NB * X_BLK;
    static constexpr int NX = (C == 16) ? 4 : 3;         // x ring depth (C = 64: 3 x 32 KB is what fits)
    static constexpr int T_TERM = 2 * CHUNK;
    static constexpr int T_BLK = TT * T_TERM;
    static constexpr int T_SLOT = NB * T_BLK;
    static constexpr int NT = 4;
    static constexpr int W1_BYTES = 3 * KS * 2 * N1 * 16;    // [ky][k step][chunk][n][8 halfs]
    static constexpr int W2_BYTES = 3 * 2 * N1 * 16;         // [ky][chunk][n][8 halfs]
    static constexpr int W3_BYTES = 3 * 2 * N3 * 16;
    static constexpr int BIAS_FLOATS = 2 * M + C;            // b1 | b2 | b3
    static constexpr int WPACK_BYTES = W1_BYTES + W2_BYTES + W3_BYTES + ((BIAS_FLOATS * 4 + 15) / 16) * 16;
    static constexpr int NA1 = 4, NA2 = 2, NA3 = (C == 64) ? 1 : 2;   // powers of two
    static constexpr int A1 = 0, A2 = A1 + NA1 * N1, A3 = A2 + NA2 * N1, ACOLS_BLK = A3 + NA3 * N3, ACOLS = NB * ACOLS_BLK;
    static constexpr int TMEM_COLS = ACOLS <= 128 ? 128 : ACOLS <= 256 ? 256 : 512;
    static constexpr int EXCH_FLOATS = 2 * 2 * (4 * 2 * M);                     // E1, E2 (double-buffered)
    static constexpr int AUX_BYTES = 512 + EXCH_FLOATS * 4;
    static constexpr size_t SMEM = (size_t)NX * X_SLOT + 2 * (size_t)NT * T_SLOT + WPACK_BYTES + AUX_BYTES + 1024;
    static constexpr int E3W = 16;                           // E3 warps: 4 lane quarters x NPART cout parts (more parallel roles beat
                                                             // fatter ones here: every role is latency-bound per warp)
    static constexpr int NPART = E3W / 4;
    static constexpr int CPT = C / NPART;                    // couts per E3 thread
    static constexpr int CH = CPT < 8 ? CPT : 8;             // couts per TMEM load / register chunk
    static constexpr int W_E1 = E3W, W_E2 = E3W + 4, W_CONV = E3W + 8, W_MMA = E3W + 12;   // first warp of each role
    static constexpr int THREADS = (E3W + 13) * 32;
    static_assert(ACOLS <= 512, "accumulators exceed TMEM");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

constexpr int BTC_PREFETCH_ROWS = 3;     // L2 prefetch distance of the x / res row streams (A/B: 3 and 1 equal, -4 % vs 6 or none; 12 is worse)

struct BlockTcArgs {
    int ko;                    // developer builds (-DBTC_KNOCKOUT): bit 0 no x loads, bit 1 no res loads, bit 2 no stores
    const float* x;        // P4 [C/4][H+2][W+2][4]  half-state F is evaluated on
    const float* res;      // P4, coupling operand (may alias out)
    float* out;            // P4
    const uint8_t* wpack;  // pack_block_tc_kernel output
    int H, W, sub;         // sub: 0 out = res + F(x), 1 out = res - F(x)
    int* status;           // status word of the call (fp16 range guard), may be null
    int n_strips, rows_per_seg;
    long long* trace;      // developer aid (VST_TC_TRACE="1016,0" / "1064,0"): clock64 stamps of CTA trace_cta, 4096 per role
    int trace_cta;
};
#ifdef BTC_TRACING
#define BTC_TRACE(role, idx) do { if (a.trace && (int)blockIdx.x == a.trace_cta && (idx) < 4096) a.trace[(role) * 4096 + (idx)] = clock64(); } while (0)
#else
#define BTC_TRACE(role, idx) do { } while (0)
#endif
// wait accounting (tracing builds): per role, cycles spent inside barrier waits vs in the whole row loop
#ifdef BTC_TRACING
#define BTC_WAIT(addr, par) do { const long long _t = clock64(); mbar_wait_a(addr, par); btc_wait_cycles += clock64() - _t; } while (0)
#define BTC_ACC_BEGIN() long long btc_wait_cycles = 0; const long long btc_t0 = clock64()
#define BTC_ACC_END(tr, role, steps) do { if (tr) { (tr)[7 * 4096 + (role) * 8] = clock64() - btc_t0; (tr)[7 * 4096 + (role) * 8 + 1] = btc_wait_cycles; (tr)[7 * 4096 + (role) * 8 + 2] = (steps); } } while (0)
#else
#define BTC_WAIT(addr, par) mbar_wait_a(addr, par)
#define BTC_ACC_BEGIN() do { } while (0)
#define BTC_ACC_END(tr, role, steps) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// weights: raw OIHW fp32 of the three convs -> one contiguous block pack (fp16 operands + fp32 biases)
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void pack_block_tc_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                     const float* __restrict__ w2, const float* __restrict__ b2,
                                     const float* __restrict__ w3, const float* __restrict__ b3, uint8_t* __restrict__ pk) {
    using Cfg = BtcCfg<C>;
    constexpr int M = Cfg::M, N1 = Cfg::N1, N3 = Cfg::N3, KS = Cfg::KS;
    __half* p1 = reinterpret_cast<__half*>(pk);
    __half* p2 = reinterpret_cast<__half*>(pk + Cfg::W1_BYTES);
    __half* p3 = reinterpret_cast<__half*>(pk + Cfg::W1_BYTES + Cfg::W2_BYTES);
    float* pb = reinterpret_cast<float*>(pk + Cfg::W1_BYTES + Cfg::W2_BYTES + Cfg::W3_BYTES);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < Cfg::W1_BYTES / 2; i += nth) {
        int r = i;
        const int e = r % 8; r /= 8;
        const int n = r % N1; r /= N1;
        const int ch = r % 2; r /= 2;
        const int ks = r % KS; r /= KS;
        const int ky = r;
        const int kx = n / M, co = n % M, ci = ks * 16 + ch * 8 + e;
        p1[i] = __float2half_rn(n < 3 * M ? w1[((size_t)co * C + ci) * 9 + ky * 3 + kx] : 0.f);
    }
    // bottleneck operand K layout: M = 16: k = channel;  M = 4: k = [hi c0..3 | lo c0..3 | 8 zeros]
    for (int i = tid; i < Cfg::W2_BYTES / 2; i += nth) {
        int r = i;
        const int e = r % 8; r /= 8;
        const int n = r % N1; r /= N1;
        const int ch = r % 2; r /= 2;
        const int ky = r;
        const int kx = n / M, co = n % M;
        const int ci = (M == 4) ? (ch == 0 ? (e & 3) : -1) : ch * 8 + e;
        p2[i] = __float2half_rn((n < 3 * M && ci >= 0) ? w2[((size_t)co * M + ci) * 9 + ky * 3 + kx] : 0.f);
    }
    for (int i = tid; i < Cfg::W3_BYTES / 2; i += nth) {
        int r = i;
        const int e = r % 8; r /= 8;
        const int n = r % N3; r /= N3;
        const int ch = r % 2; r /= 2;
        const int ky = r;
        const int kx = n / C, co = n % C;
        const int ci = (M == 4) ? (ch == 0 ? (e & 3) : -1) : ch * 8 + e;
        p3[i] = __float2half_rn(ci >= 0 ? w3[((size_t)co * M + ci) * 9 + ky * 3 + kx] : 0.f);
    }
    for (int i = tid; i < Cfg::BIAS_FLOATS; i += nth) pb[i] = i < M ? b1[i] : (i < 2 * M ? b2[i - M] : b3[i - 2 * M]);
}

size_t block_tc_pack_floats(int C) {
    return (size_t)((C == 16 ? BtcCfg<16>::WPACK_BYTES : BtcCfg<64>::WPACK_BYTES) + 15) / 16 * 4;
}
bool block_tc_eligible(int C, int mult) { return (C == 16 || C == 64) && mult == 4; }

int launch_pack_block_tc(int C, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                         const float* b3, float* pk, cudaStream_t st) {
    if (C == 16) pack_block_tc_kernel<16><<<16, 256, 0, st>>>(w1, b1, w2, b2, w3, b3, reinterpret_cast<uint8_t*>(pk));
    else pack_block_tc_kernel<64><<<64, 256, 0, st>>>(w1, b1, w2, b2, w3, b3, reinterpret_cast<uint8_t*>(pk));
    return check_launch("pack_block_tc");
}

// ------------------------------------------------------------------------------------------
// device helpers.  Code size matters: every role is one warp per scheduler that walks its per-row code once per
// step, so the per-step code of all roles together must stay inside the 32 KB L1.5 instruction cache
// (B300_MICROARCH.md, I-cache) — no inlined slow paths, shared E1/E2 body, descriptors built by one add.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void btc_umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ bool btc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t btc_pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ int btc_mir(int y, int H) { return y < 0 ? -y : (y >= H ? 2 * H - 2 - y : y); }
__device__ __forceinline__ float btc_lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 btc_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void btc_sts(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void btc_sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 8 scaled fp32 values -> 16 bytes of fp16 hi and 16 bytes of fp16 lo
__device__ __forceinline__ void btc_split8(const float* x, uint4& hv, uint4& lv, uint32_t& hmax) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
        const float2 hf = __half22float2(hh);
        hw[e] = *reinterpret_cast<const uint32_t*>(&hh);
        hmax = range_fold(hmax, hw[e]);
        lw[e] = btc_pack_half2(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
    }
    hv = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

struct BtcSeg {
    int x0;                    // first output column of the strip
    int ya, yb;                // output rows [ya, yb)
    int t2a, t2b, t1a, t1b;    // inclusive row ranges of the intermediates this segment needs
    int xa, xb;                // inclusive image-row range of x (padded row = +1)
};

// E1 / E2 (one shared, out-of-line body): accumulator ring -> (kx fold, bias, ReLU, fp16 split) -> t ring.
// Shared-memory objects are passed as 32-bit shared addresses.
struct BtcMidArgs {
    uint32_t tacc;                 // TMEM column of accumulator slot 0
    uint32_t acc_full;             // mbarrier array (8 bytes per slot)
    uint32_t tring, t_full, t_empty;
    uint32_t bias, exch;
    int nacc, nacc_log2, rows, bar_id, x0, W;
    int windowed;                  // store the t row in window form (btc_win_rows): E2
    int* status;
    long long* trace;              // tracing builds: wait accounting of this role (nullptr otherwise)
};
// Window form of a t2 row (conv3's A operand): operand row 32q + l holds t2 pixel 2 + 30q + l, i.e. every warp of E3
// (TMEM lanes 32q .. 32q+31) sees a self-contained 32-pixel window whose 30 interior lanes fold kx with shuffles
// alone — no exchange through shared memory, no barrier.  Pixels on a window seam live in two rows.
__device__ __forceinline__ void btc_win_rows(int p, int& r0, int& r1) {
    r0 = r1 = -1;
    const int w = p - 2;
    if (w < 0 || w >= 122) return;
    const int q = min(w / 30, 3), l = w - 30 * q;
    r0 = 32 * q + l;
    if (q > 0 && l < 2) r1 = 32 * (q - 1) + 30 + l;
}

template <int C>
static __device__ __noinline__ void btc_mid_epilogue(const BtcMidArgs g) {
    using Cfg = BtcCfg<C>;
    constexpr int M = Cfg::M, N = Cfg::N1;
    constexpr int MP = M > 8 ? 8 : M, NPASS = M / MP;     // channels per pass: M = 16 runs two passes of 8 (register budget: 64)
    const int lane = threadIdx.x & 31, q = (threadIdx.x >> 5) & 3;
    const int m = q * 32 + lane;
    const uint32_t ex_pub = g.exch + (uint32_t)((q * 2 + (lane == 0 ? 1 : 0)) * M * 4);      // where lane 31 / lane 0 publish
    const uint32_t ex_l = g.exch + (uint32_t)(((q > 0 ? q - 1 : 0) * 2 + 0) * M * 4);          // left warp's lane 31, kx = 0
    const uint32_t ex_r = g.exch + (uint32_t)(((q < 3 ? q + 1 : 3) * 2 + 1) * M * 4);          // right warp's lane 0, kx = 2
    const uint32_t trow = g.tacc + ((uint32_t)(q * 32) << 16);
    uint32_t par = 0;
    uint32_t hmax = 0u;                         // fp16 range guard (tc_ptx.cuh)
    int o0 = m, o1 = -1;                        // operand rows of this thread's own pixel
    if (g.windowed) btc_win_rows(m, o0, o1);
    // store one pixel's 16-byte operand chunk at operand rows r0 / r1 (negative: none)
    auto put = [](uint32_t base, int r0, int r1, uint4 v) {
        if (r0 >= 0) btc_sts128(base + 16 * r0, v);
        if (r1 >= 0) btc_sts128(base + 16 * r1, v);
    };
    BTC_ACC_BEGIN();
#pragma unroll 1
    for (int l = 0; l < g.rows; ++l) {
        const int st = l & 3, sa = l & (g.nacc - 1);
        BTC_WAIT(g.acc_full + 8 * sa, (uint32_t)((l >> g.nacc_log2) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int blk = 0; blk < Cfg::NB; ++blk, par ^= (uint32_t)(4 * 2 * M * 4)) {
            const int x = g.x0 + blk * Cfg::XO - 3 + m;
            const bool own = (x >= 0) && (x < g.W);
            const bool mir_l = (x == 1) && (m >= 2), mir_r = (x == g.W - 2) && (m + 2 < Cfg::PW);
            const uint32_t slot = g.tring + (uint32_t)(st * Cfg::T_SLOT + blk * Cfg::T_BLK);
            int ml0 = m - 2, ml1 = -1, mr0 = m + 2, mr1 = -1;            // operand rows of the reflected copies (image edge)
            if (g.windowed && (mir_l || mir_r)) { btc_win_rows(m - 2, ml0, ml1); btc_win_rows(m + 2, mr0, mr1); }
#pragma unroll 1
            for (int ps = 0; ps < NPASS; ++ps) {
                // D columns of this pass: [kx][MP channels]
                uint32_t du[3 * MP < 16 ? 16 : 3 * MP];
                if (M == 4) {
                    tmem_ld16_nowait(trow + (uint32_t)(blk * Cfg::ACOLS_BLK + sa * N), du);      // 12 of the 16 padded columns
                } else {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        tmem_ld8_nowait(trow + (uint32_t)(blk * Cfg::ACOLS_BLK + sa * N + kx * M + ps * MP), du + kx * MP);
                }
                float bs[MP];
#pragma unroll
                for (int c = 0; c < MP; c += 4) {
                    const float4 b4 = btc_lds128(g.bias + 4 * (ps * MP + c));
                    bs[c] = b4.x; bs[c + 1] = b4.y; bs[c + 2] = b4.z; bs[c + 3] = b4.w;
                }
                tmem_ld_wait();
                tmem_ld_fence_regs<(3 * MP < 16 ? 16 : 3 * MP)>(du);
                float d[3 * MP];
#pragma unroll
                for (int c = 0; c < 3 * MP; ++c) d[c] = __uint_as_float(du[c]);
                if (lane == 31 || lane == 0) {
#pragma unroll
                    for (int c = 0; c < MP; ++c) btc_sts(ex_pub + par + 4 * (ps * MP + c), lane == 0 ? d[2 * MP + c] : d[c]);
                }
                named_barrier(g.bar_id, 128);
                // neighbour-warp partial sums: warp-uniform addresses (broadcast loads), all issued before any use
                float el[MP], er[MP];
#pragma unroll
                for (int c = 0; c < MP; c += 4) {
                    const float4 a4 = btc_lds128(ex_l + par + 4 * (ps * MP + c)), b4 = btc_lds128(ex_r + par + 4 * (ps * MP + c));
                    el[c] = a4.x; el[c + 1] = a4.y; el[c + 2] = a4.z; el[c + 3] = a4.w;
                    er[c] = b4.x; er[c + 1] = b4.y; er[c + 2] = b4.z; er[c + 3] = b4.w;
                }
                float o[MP];
#pragma unroll
                for (int c = 0; c < MP; ++c) {
                    const float ls = __shfl_up_sync(0xffffffffu, d[c], 1);
                    const float rs = __shfl_down_sync(0xffffffffu, d[2 * MP + c], 1);
                    const float lv = (lane == 0) ? el[c] : ls;
                    const float rv = (lane == 31) ? er[c] : rs;
                    o[c] = fmaxf(((lv + d[MP + c]) + rv) * (1.0f / VST_HALF_SCALE) + bs[c], 0.f) * VST_HALF_SCALE;
                }
                if (blk == 0 && ps == 0) BTC_WAIT(g.t_empty + 8 * st, (uint32_t)(((l >> 2) & 1) ^ 1));
                if (M == 4) {
                    const __half2 h01 = __floats2half2_rn(o[0], o[1]), h23 = __floats2half2_rn(o[2], o[3]);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    hmax = range_fold(range_fold(hmax, *reinterpret_cast<const uint32_t*>(&h01)), *reinterpret_cast<const uint32_t*>(&h23));
                    const uint4 v = make_uint4(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23),
                                               btc_pack_half2(o[0] - f01.x, o[1] - f01.y), btc_pack_half2(o[2] - f23.x, o[3] - f23.y));
                    if (own) put(slot, o0, o1, v);
                    if (mir_l) put(slot, ml0, ml1, v);
                    if (mir_r) put(slot, mr0, mr1, v);
                } else {
                    uint4 hv, lv;
                    btc_split8(o, hv, lv, hmax);
                    const uint32_t ph = slot + ps * Cfg::CHUNK, pl = ph + Cfg::T_TERM;     // term 0 (hi) / term 1 (lo), K chunk ps
                    if (own) { put(ph, o0, o1, hv); put(pl, o0, o1, lv); }
                    if (mir_l) { put(ph, ml0, ml1, hv); put(pl, ml0, ml1, lv); }
                    if (mir_r) { put(ph, mr0, mr1, hv); put(pl, mr0, mr1, lv); }
                }
            }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive_a(g.t_full + 8 * st);
    }
    range_report(hmax, g.status);
    if (m == 0) BTC_ACC_END(g.trace, g.bar_id, g.rows);
}

// MODE of a launch (compile time: the hot instantiation keeps its registers):
//   BTC_PLAIN      out and res are P4 tensors at the block's own resolution
//   BTC_OUT_SQZ    the block in front of a stride-2 block stores its result SQUEEZED (space-to-depth,
//                  models/RevResNet.py:34-37: plane (dy*2+dx)*G + g at (y/2, x/2)), with the border the squeezed-input
//                  form of the stride-2 conv needs (top row / left column replicated, conv_tch.cu; bottom / right
//                  reflected so that every border value is finite) — the standalone space_to_depth and
//                  p4_replicate_topleft launches of the transition disappear
//   BTC_RES_UNSQZ  the block behind a stride-2 block in the INVERSE pass reads its coupling operand through the
//                  unsqueeze (depth-to-space, :40-43) addressing from the still squeezed tensor — no depth_to_space launch
constexpr int BTC_PLAIN = 0, BTC_OUT_SQZ = 1, BTC_RES_UNSQZ = 2;

// store pixel (ys, xs) of a squeezed plane [Hs+2][Ws+2] and the border positions that depend on it
static __device__ __noinline__ void btc_sqz_border(float4* plane, int Hs, int Ws, int ys, int xs, float4 v) {
    const int Wps = Ws + 2;
    // padded rows / columns this pixel feeds: itself, the replicated top / left border (from row / column 0), the
    // reflected bottom / right border (from row Hs-2 / column Ws-2); a 2-pixel map feeds both borders from pixel 0
    const int rows[3] = {ys + 1, ys == 0 ? 0 : -1, ys == Hs - 2 ? Hs + 1 : -1};
    const int cols[3] = {xs + 1, xs == 0 ? 0 : -1, xs == Ws - 2 ? Ws + 1 : -1};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if ((i | j) && rows[i] >= 0 && cols[j] >= 0) plane[(size_t)rows[i] * Wps + cols[j]] = v;
}
__device__ __forceinline__ void btc_store_sqz(float4* out, int G, int H, int W, int g, int y, int x, float4 v) {
    const int Hs = H >> 1, Ws = W >> 1, ys = y >> 1, xs = x >> 1, k = ((y & 1) << 1) | (x & 1);
    float4* plane = out + (size_t)(k * G + g) * ((size_t)(Hs + 2) * (Ws + 2));
    plane[(size_t)(ys + 1) * (Ws + 2) + xs + 1] = v;
    if ((ys == 0) | (ys == Hs - 2) | (xs == 0) | (xs == Ws - 2)) btc_sqz_border(plane, Hs, Ws, ys, xs, v);
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int C, int MODE>
__global__ void __launch_bounds__(BtcCfg<C>::THREADS, 1) rev_block_tc_kernel(BlockTcArgs a) {
    using Cfg = BtcCfg<C>;
    constexpr int M = Cfg::M, G = Cfg::G, KS = Cfg::KS, N1 = Cfg::N1, N3 = Cfg::N3, TT = Cfg::TT, NX = Cfg::NX, NT = Cfg::NT;
    constexpr int NA1 = Cfg::NA1, NA2 = Cfg::NA2, NA3 = Cfg::NA3;

    pdl_launch_dependents();             // the next kernel of the stream may start its prologue on SMs that drain
    const int H = a.H, W = a.W, Hp = H + 2, Wp = W + 2;
    BtcSeg sg;
    {
        const int sx = blockIdx.x % a.n_strips, sy = blockIdx.x / a.n_strips;
        sg.x0 = sx * (Cfg::NB * Cfg::XO);
        sg.ya = sy * a.rows_per_seg;
        sg.yb = min(H, sg.ya + a.rows_per_seg);
        if (sg.ya >= H) return;
        sg.t2a = max(0, sg.ya - 1); sg.t2b = min(H - 1, sg.yb);
        sg.t1a = max(0, sg.ya - 2); sg.t1b = min(H - 1, sg.yb + 1);
        sg.xa = sg.t1a - 1; sg.xb = sg.t1b + 1;
    }

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* xring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* t1ring = xring + (size_t)NX * Cfg::X_SLOT;
    uint8_t* t2ring = t1ring + (size_t)NT * Cfg::T_SLOT;
    uint8_t* wsm = t2ring + (size_t)NT * Cfg::T_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + Cfg::WPACK_BYTES);
    uint64_t* x_full = bars + 40;         // [NX <= 4]  128 converter threads
    uint64_t* x_empty = bars + 44;        // [NX <= 4]  tcgen05.commit
    uint64_t* a1_full = bars + 6;         // [4]   tcgen05.commit   (no a1/a2 "empty" barriers: implied, see the issuer)
    uint64_t* t1_full = bars + 12;        // [4]   128 E1 threads
    uint64_t* t1_empty = bars + 16;       // [4]   tcgen05.commit
    uint64_t* a2_full = bars + 20;        // [2]
    uint64_t* t2_full = bars + 24;        // [4]
    uint64_t* t2_empty = bars + 28;       // [4]
    uint64_t* a3_full = bars + 32;        // [2]
    uint64_t* a3_empty = bars + 34;       // [2]   256 E3 threads
    uint64_t* w_bar = bars + 36;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);
    float* exch1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
    float* exch2 = exch1 + 2 * (4 * 2 * M);
    const float* bias_s = reinterpret_cast<const float*>(wsm + Cfg::W1_BYTES + Cfg::W2_BYTES + Cfg::W3_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // operand rings start out zero: never-written positions (columns outside the image, the K padding of M = 4)
    // must be finite, and the padding must be zero
    {
        uint4* z = reinterpret_cast<uint4*>(xring);
        const int n16 = (NX * Cfg::X_SLOT + 2 * NT * Cfg::T_SLOT) / 16;
        for (int i = tid; i < n16; i += Cfg::THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int s = 0; s < NX; ++s) { mbar_init(&x_full[s], 128); mbar_init(&x_empty[s], 1); }
        for (int s = 0; s < 4; ++s) mbar_init(&a1_full[s], 1);
        for (int s = 0; s < 4; ++s) {
            mbar_init(&t1_full[s], 128); mbar_init(&t1_empty[s], 1);
            mbar_init(&t2_full[s], 128); mbar_init(&t2_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a2_full[s], 1);
            mbar_init(&a3_full[s], 1); mbar_init(&a3_empty[s], 32 * Cfg::E3W);
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == Cfg::W_MMA) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                          // everything above is independent of the previous kernel's output

    if (warp == Cfg::W_MMA) {
        // ================= weights (one TMA copy) + UMMA issuer =================
        // the whole warp walks the (warp-uniform) schedule; one elected lane issues the tcgen05 instructions
        {
            if (lane == 0) {
                mbar_arrive_expect_tx(w_bar, Cfg::WPACK_BYTES);
                bulk_g2s(wsm, a.wpack, Cfg::WPACK_BYTES, w_bar);
            }
            mbar_wait_a(smem_u32(w_bar), 0u);
            constexpr uint32_t IDESC1 = (1u << 4) | ((uint32_t)(N1 >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t IDESC3 = (1u << 4) | ((uint32_t)(N3 >> 3) << 17) | ((128u >> 4) << 24);
            // descriptors = constant high bits + (shared address >> 4): one 64-bit add per operand
            const uint64_t dX = make_desc(smem_u32(xring), Cfg::CHUNK, 128);
            const uint64_t dT1 = make_desc(smem_u32(t1ring), Cfg::CHUNK, 128), dT2 = make_desc(smem_u32(t2ring), Cfg::CHUNK, 128);
            const uint64_t dW1 = make_desc(smem_u32(wsm), N1 * 16, 128);
            const uint64_t dW2 = make_desc(smem_u32(wsm) + Cfg::W1_BYTES, N1 * 16, 128);
            const uint64_t dW3 = make_desc(smem_u32(wsm) + Cfg::W1_BYTES + Cfg::W2_BYTES, N3 * 16, 128);
            const uint32_t bx_full = smem_u32(x_full), bx_empty = smem_u32(x_empty), b1_full = smem_u32(a1_full),
                           bt1_full = smem_u32(t1_full), bt1_empty = smem_u32(t1_empty), b2_full = smem_u32(a2_full),
                           bt2_full = smem_u32(t2_full), bt2_empty = smem_u32(t2_empty), b3_full = smem_u32(a3_full),
                           b3_empty = smem_u32(a3_empty);
            int sx = 0;                      // x ring slot of row rx, its phase
            uint32_t px = 0;
            // Step rx (order matters — it makes the accumulator-drained barriers of acc1 / acc2 implicit):
            //   wait t2 row rx-6  -> conv3(y = rx-7)   [t2 rows <= y+1 ready]
            //   wait t1 row rx-3  -> conv2(i = rx-4)   [t1 rows <= i+1 ready; acc2 slot of row i-2 = rx-6 was drained
            //                                           before E2 published t2 row rx-6]
            //   wait x row rx     -> conv1(rx)         [starts t1 row rx+1 in the acc1 slot of row rx-3, drained before
            //                                           E1 published t1 row rx-3]
            BTC_ACC_BEGIN();
#pragma unroll 1
            for (int rx = sg.xa;; ++rx) {
                const bool el = btc_elect_one();
                BTC_TRACE(0, 6 * (rx - sg.xa));
                // ---- conv3: out row y from t2 rows mir(y-1), y, mir(y+1)
                {
                    const int r2 = rx - 6 - sg.t2a;
                    if (r2 >= 0 && rx - 6 <= sg.t2b) BTC_WAIT(bt2_full + 8 * (r2 & 3), (uint32_t)((r2 >> 2) & 1));
                }
                const int y = rx - 7;
                if (y >= sg.ya && y < sg.yb) {
                    const int ly = y - sg.ya, sa = ly & (NA3 - 1);
                    BTC_WAIT(b3_empty + 8 * sa, (uint32_t)(((ly >> (NA3 - 1)) & 1) ^ 1));
                    tc_fence_after();
                    BTC_TRACE(0, 6 * (rx - sg.xa) + 1);
                    const uint32_t dcol = tmem_base + Cfg::A3 + sa * N3;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int st = (btc_mir(y + ky - 1, H) - sg.t2a) & 3;
                        const uint64_t ad = dT2 + (uint64_t)(st * (Cfg::T_SLOT >> 4));
                        const uint64_t bd = dW3 + (uint64_t)((ky * 2 * N3 * 16) >> 4);
#pragma unroll
                        for (int blk = 0; blk < Cfg::NB; ++blk)
#pragma unroll
                            for (int t = 0; t < TT; ++t)
                                if (el) btc_umma_f16(dcol + blk * Cfg::ACOLS_BLK, ad + (uint64_t)((blk * Cfg::T_BLK + t * Cfg::T_TERM) >> 4), bd,
                                                     IDESC3, ky > 0 || t > 0);
                    }
                    if (el) umma_commit_a(b3_full + 8 * sa);
                    if (el && y - 1 >= sg.t2a) umma_commit_a(bt2_empty + 8 * ((y - 1 - sg.t2a) & 3));
                }
                BTC_TRACE(0, 6 * (rx - sg.xa) + 2);
                // ---- conv2: t2 row i from t1 rows mir(i-1), i, mir(i+1)
                {
                    const int r1 = rx - 3 - sg.t1a;
                    if (r1 >= 0 && rx - 3 <= sg.t1b) BTC_WAIT(bt1_full + 8 * (r1 & 3), (uint32_t)((r1 >> 2) & 1));
                }
                const int i = rx - 4;
                if (i >= sg.t2a && i <= sg.t2b) {
                    tc_fence_after();
                    BTC_TRACE(0, 6 * (rx - sg.xa) + 3);
                    const int li = i - sg.t2a, sa = li & (NA2 - 1);
                    const uint32_t dcol = tmem_base + Cfg::A2 + sa * N1;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int st = (btc_mir(i + ky - 1, H) - sg.t1a) & 3;
                        const uint64_t ad = dT1 + (uint64_t)(st * (Cfg::T_SLOT >> 4));
                        const uint64_t bd = dW2 + (uint64_t)((ky * 2 * N1 * 16) >> 4);
#pragma unroll
                        for (int blk = 0; blk < Cfg::NB; ++blk)
#pragma unroll
                            for (int t = 0; t < TT; ++t)
                                if (el) btc_umma_f16(dcol + blk * Cfg::ACOLS_BLK, ad + (uint64_t)((blk * Cfg::T_BLK + t * Cfg::T_TERM) >> 4), bd,
                                                     IDESC1, ky > 0 || t > 0);
                    }
                    if (el) umma_commit_a(b2_full + 8 * sa);
                    if (el && i - 1 >= sg.t1a) umma_commit_a(bt1_empty + 8 * ((i - 1 - sg.t1a) & 3));
                }
                // ---- conv1: x row rx -> t1 rows rx-1 (ky=2, completes it), rx (ky=1), rx+1 (ky=0, starts it)
                if (rx <= sg.xb) {
                    BTC_WAIT(bx_full + 8 * sx, px);
                    tc_fence_after();
                    BTC_TRACE(0, 6 * (rx - sg.xa) + 4);
                    const uint64_t dXs = dX + (uint64_t)(sx * (Cfg::X_SLOT >> 4));
#pragma unroll
                    for (int ky = 2; ky >= 0; --ky) {
                        const int j = rx + 1 - ky;
                        if (j >= sg.t1a && j <= sg.t1b) {
                            const int sa = (j - sg.t1a) & (NA1 - 1);
                            const uint32_t dcol = tmem_base + Cfg::A1 + sa * N1;
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                const uint64_t bd = dW1 + (uint64_t)(((ky * KS + ks) * 2 * N1 * 16) >> 4);
#pragma unroll
                                for (int blk = 0; blk < Cfg::NB; ++blk) {
                                    const uint64_t ax = dXs + (uint64_t)((blk * Cfg::X_BLK + ks * 2 * Cfg::CHUNK) >> 4);
                                    if (el) btc_umma_f16(dcol + blk * Cfg::ACOLS_BLK, ax, bd, IDESC1, ky > 0 || ks > 0);
                                    if (el) btc_umma_f16(dcol + blk * Cfg::ACOLS_BLK, ax + (uint64_t)(Cfg::XT >> 4), bd, IDESC1, true);
                                }
                            }
                            if (ky == 2 && el) umma_commit_a(b1_full + 8 * sa);
                        }
                    }
                    if (el) umma_commit_a(bx_empty + 8 * sx);
                    if (++sx == NX) { sx = 0; px ^= 1u; }
                }
                BTC_TRACE(0, 6 * (rx - sg.xa) + 5);
                if (y >= sg.yb - 1) break;
            }
            if (lane == 0) BTC_ACC_END((a.trace && (int)blockIdx.x == a.trace_cta) ? a.trace : nullptr, 0, sg.yb + 7 - sg.xa);
        }
        __syncwarp();
    } else if (warp >= Cfg::W_CONV) {
        // ================= converters: fp32 P4 rows (global / L2) -> fp16 hi | lo operand rows =================
        const int m = tid - Cfg::W_CONV * 32;                         // staged pixel (of every block)
        constexpr int NB = Cfg::NB;
        const int pc0 = max(sg.x0 - 2, 0);
        int pcb[NB];                                                  // padded column per block (clamped: garbage pixels only)
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) pcb[blk] = min(max(sg.x0 + blk * Cfg::XO - 2 + m, 0), Wp - 1);
        const size_t plane = (size_t)Hp * Wp;
        const float4* in4 = reinterpret_cast<const float4*>(a.x);
        const uint32_t bx_full = smem_u32(x_full), bx_empty = smem_u32(x_empty);
        int s = 0;
        uint32_t pe = 1;
        uint32_t hmax = 0u;                                           // fp16 range guard (tc_ptx.cuh)
        constexpr int GB = G < 8 ? G : 8;                             // groups per batch of loads in flight
        constexpr int NBATCH = G / GB, ITEMS = NB * NBATCH;           // item = (block, batch) of one row
        // software pipeline: the next item (possibly of the next row) is requested before the current one is converted
        float4 v[GB];
        {
            const float4* src = in4 + (size_t)(sg.xa + 1) * Wp + pcb[0];
#pragma unroll
            for (int g = 0; g < GB; ++g) v[g] = __ldg(src + (size_t)g * plane);
        }
#ifdef BTC_KNOCKOUT
        const bool ko_x = a.ko & 1;
#endif
        BTC_ACC_BEGIN();
#pragma unroll 1
        for (int rx = sg.xa; rx <= sg.xb; ++rx) {
            if (m < G && rx + BTC_PREFETCH_ROWS <= sg.xb) {
#pragma unroll
                for (int blk = 0; blk < NB; ++blk)
                    l2_prefetch(in4 + (size_t)m * plane + (size_t)(rx + 1 + BTC_PREFETCH_ROWS) * Wp + min(pc0 + blk * Cfg::PW, Wp - 1), Cfg::CHUNK);
            }
            if (m == 0) BTC_TRACE(1, 3 * (rx - sg.xa));
            BTC_WAIT(bx_empty + 8 * s, pe);
            if (m == 0) BTC_TRACE(1, 3 * (rx - sg.xa) + 1);
            const uint32_t hi0 = smem_u32(xring) + (uint32_t)(s * Cfg::X_SLOT + m * 16);
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const int blk = it / NBATCH, b = it % NBATCH;
                float xs[4 * GB];
#pragma unroll
                for (int g = 0; g < GB; ++g) {
                    xs[4 * g] = v[g].x * VST_HALF_SCALE; xs[4 * g + 1] = v[g].y * VST_HALF_SCALE;
                    xs[4 * g + 2] = v[g].z * VST_HALF_SCALE; xs[4 * g + 3] = v[g].w * VST_HALF_SCALE;
                }
                {   // next item: same row, or item 0 of the next row (clamped at the end: a harmless re-read)
                    const int nit = (it + 1 == ITEMS) ? 0 : it + 1;
                    const int nrow = min((it + 1 == ITEMS) ? rx + 1 : rx, sg.xb);
                    const float4* src = in4 + (size_t)((nit % NBATCH) * GB) * plane + (size_t)(nrow + 1) * Wp + pcb[nit / NBATCH];
#ifdef BTC_KNOCKOUT
                    if (!ko_x)
#endif
#pragma unroll
                    for (int g = 0; g < GB; ++g) v[g] = __ldg(src + (size_t)g * plane);
                }
                const uint32_t hi = hi0 + blk * Cfg::X_BLK;
#pragma unroll
                for (int k = 0; k < GB / 2; ++k) {
                    uint4 hv, lv;
                    btc_split8(xs + 8 * k, hv, lv, hmax);
                    btc_sts128(hi + (b * (GB / 2) + k) * Cfg::CHUNK, hv);
                    btc_sts128(hi + Cfg::XT + (b * (GB / 2) + k) * Cfg::CHUNK, lv);
                }
            }
            fence_proxy_async();
            mbar_arrive_a(bx_full + 8 * s);
            if (m == 0) BTC_TRACE(1, 3 * (rx - sg.xa) + 2);
            if (++s == NX) { s = 0; pe ^= 1u; }
        }
        range_report(hmax, a.status);
        if (m == 0) BTC_ACC_END((a.trace && (int)blockIdx.x == a.trace_cta) ? a.trace : nullptr, 3, sg.xb - sg.xa + 1);
    } else if (warp >= Cfg::W_E1) {
        // ================= E1 (warps 8-11) / E2 (warps 12-15) =================
        mbar_wait_a(smem_u32(w_bar), 0u);
        const bool second = warp >= Cfg::W_E2;
        BtcMidArgs g;
        g.tacc = tmem_base + (second ? Cfg::A2 : Cfg::A1);
        g.acc_full = smem_u32(second ? a2_full : a1_full);
        g.tring = smem_u32(second ? t2ring : t1ring);
        g.t_full = smem_u32(second ? t2_full : t1_full);
        g.t_empty = smem_u32(second ? t2_empty : t1_empty);
        g.bias = smem_u32(bias_s) + (second ? 4 * M : 0);
        g.exch = smem_u32(second ? exch2 : exch1);
        g.nacc = second ? NA2 : NA1;
        g.nacc_log2 = second ? 1 : 2;
        g.rows = second ? sg.t2b - sg.t2a + 1 : sg.t1b - sg.t1a + 1;
        g.bar_id = second ? 2 : 1;
        g.windowed = second ? 1 : 0;
        g.x0 = sg.x0; g.W = W;
        g.status = a.status;
        g.trace = (a.trace && (int)blockIdx.x == a.trace_cta) ? a.trace : nullptr;
        btc_mid_epilogue<C>(g);
    } else {
        // ================= E3: acc3 -> kx fold, bias, coupling with res -> global P4 (+ reflection border) =================
        constexpr int CPT = Cfg::CPT, CH = Cfg::CH, NB = Cfg::NB;
        // conv3's operand rows are in window form (btc_win_rows): lane l of lane quarter q holds t2 pixel 2 + 30q + l, and
        // lanes 1 .. 30 produce output pixel 30q + l - 1 of the block from their own and their neighbours' partial sums
        const int q = warp & 3, half = warp >> 2;                            // half = cout part of this warp (0 .. NPART-1)
        const int xb0 = sg.x0 + 30 * q + lane - 1;                           // image column of this thread in block 0
        const bool min_ok = (lane >= 1) && (lane <= 30);
        const size_t plane = (size_t)Hp * Wp;
#ifdef BTC_KNOCKOUT
        const bool has_res = a.res != nullptr && !(a.ko & 2);
        const bool ko_st = a.ko & 4;
#else
        const bool has_res = a.res != nullptr;                             // null: the coupling operand is zero (out = +/- F(x))
#endif
        const bool res_ok = min_ok && has_res;
        const float4* resp = reinterpret_cast<const float4*>(a.res) + (size_t)(half * (CPT / 4)) * plane + (xb0 + 1);
        float4* outp = reinterpret_cast<float4*>(a.out) + (size_t)(half * (CPT / 4)) * plane;
        const int g_first = half * (CPT / 4);                               // first P4 group of this thread's couts
        // coupling operand of group g_first + j at image pixel (yy, xb0 + dx)
        auto res_ld = [&](int j, int yy, int dx) -> float4 {
            if (MODE == BTC_RES_UNSQZ) {
                const int xx = xb0 + dx, Hs = H >> 1, Ws = W >> 1, k = ((yy & 1) << 1) | (xx & 1);
                return reinterpret_cast<const float4*>(a.res)[(size_t)(k * G + g_first + j) * ((size_t)(Hs + 2) * (Ws + 2)) +
                                                              (size_t)((yy >> 1) + 1) * (Ws + 2) + (xx >> 1) + 1];
            }
            return resp[(size_t)j * plane + (size_t)(yy + 1) * Wp + dx];
        };
        auto out_st = [&](int j, int yy, int xx, float4 o) {
#ifdef BTC_KNOCKOUT
            if (ko_st && o.x != 12345.678f) return;
#endif
            if (MODE == BTC_OUT_SQZ) btc_store_sqz(reinterpret_cast<float4*>(a.out), G, H, W, g_first + j, yy, xx, o);
            else p4_store(outp + (size_t)j * plane, H, W, yy, xx, o);
        };
        const float sgn = a.sub ? -1.f : 1.f;
        mbar_wait_a(smem_u32(w_bar), 0u);
        // shared-memory objects as 32-bit shared addresses (explicit ld/st.shared: no generic-address path)
        const uint32_t b3p = smem_u32(bias_s + 2 * M + half * CPT);          // this thread's couts
        const uint32_t b3_full = smem_u32(a3_full), b3_empty = smem_u32(a3_empty);
        const uint32_t trow0 = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::A3 + half * CPT;
        float4 rs[CH / 4], rn[CH / 4];                                       // coupling operand: current / next item
#pragma unroll
        for (int j = 0; j < CH / 4; ++j)
            rs[j] = (res_ok && xb0 < W) ? res_ld(j, sg.ya, 0) : make_float4(0.f, 0.f, 0.f, 0.f);
        BTC_ACC_BEGIN();
#pragma unroll 1
        for (int y = sg.ya; y < sg.yb; ++y) {
            const int ly = y - sg.ya, sa = ly & (NA3 - 1);
            if (tid == 0) BTC_TRACE(4, 4 * ly);
            if (MODE != BTC_RES_UNSQZ && has_res && tid < G && y + BTC_PREFETCH_ROWS < sg.yb)
                l2_prefetch(reinterpret_cast<const float4*>(a.res) + (size_t)tid * plane + (size_t)(y + 1 + BTC_PREFETCH_ROWS) * Wp + sg.x0 + 1,
                            (uint32_t)(max(min(NB * Cfg::XO, W - sg.x0), 1) * 16));
            if (CPT == CH) {
                // ---- all of this thread's couts fit in registers: one pass per block.  The coupling operand of the
                //      NEXT (row, block) item is requested before the current one is processed (rs/rn live across rows)
                BTC_WAIT(b3_full + 8 * sa, (uint32_t)((ly >> (NA3 - 1)) & 1));
                tc_fence_after();
                if (tid == 0) BTC_TRACE(4, 4 * ly + 1);
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) {
                    const int x = xb0 + blk * Cfg::XO;
                    const bool xin = min_ok && (x < W);
                    {   // next item: block blk+1 of this row, or block 0 of the next row
                        const int nblk = (blk + 1 < NB) ? blk + 1 : 0;
                        const int ny = (blk + 1 < NB) ? y : y + 1;
                        const bool nin = res_ok && (xb0 + nblk * Cfg::XO < W) && (ny < sg.yb);
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j)
                            rn[j] = nin ? res_ld(j, ny, nblk * Cfg::XO) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const uint32_t trow = trow0 + blk * Cfg::ACOLS_BLK + sa * N3;
                    uint32_t u0[CH], u1[CH], u2[CH];
                    tmem_ld_nowait<CH>(trow + 0 * C, u0);
                    tmem_ld_nowait<CH>(trow + 1 * C, u1);
                    tmem_ld_nowait<CH>(trow + 2 * C, u2);
                    tmem_ld_wait();
                    tmem_ld_fence_regs<CH>(u0); tmem_ld_fence_regs<CH>(u1); tmem_ld_fence_regs<CH>(u2);
                    float v0[CH], v1[CH], v2[CH];
#pragma unroll
                    for (int i = 0; i < CH; ++i) { v0[i] = __uint_as_float(u0[i]); v1[i] = __uint_as_float(u1[i]); v2[i] = __uint_as_float(u2[i]); }
                    if (tid == 0 && blk == 0) BTC_TRACE(5, 8 * ly);
                    if (blk == NB - 1) {
                        tc_fence_before();
                        mbar_arrive_a(b3_empty + 8 * sa);
                    }
                    if (tid == 0) BTC_TRACE(4, 4 * ly + 2);
                    float bb[CH];
#pragma unroll
                    for (int i = 0; i < CH; i += 4) {       // warp-uniform addresses (broadcast)
                        const float4 d4 = btc_lds128(b3p + 4 * i);
                        bb[i] = d4.x; bb[i + 1] = d4.y; bb[i + 2] = d4.z; bb[i + 3] = d4.w;
                    }
#pragma unroll
                    for (int i = 0; i < CH; ++i) {          // lanes 0 / 31 (window halo) compute values nobody stores
                        const float lv = __shfl_up_sync(0xffffffffu, v0[i], 1);
                        const float rv = __shfl_down_sync(0xffffffffu, v2[i], 1);
                        v1[i] = sgn * (((lv + v1[i]) + rv) * (1.0f / VST_HALF_SCALE) + bb[i]);
                    }
                    if (tid == 0 && blk == 0) BTC_TRACE(5, 8 * ly + 3);
                    if (xin) {
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) {
                            const float4 rr = rs[j];
                            float4 o;
                            o.x = rr.x + v1[4 * j]; o.y = rr.y + v1[4 * j + 1]; o.z = rr.z + v1[4 * j + 2]; o.w = rr.w + v1[4 * j + 3];
                            out_st(j, y, x, o);
                        }
                    }
                    if (tid == 0 && blk == 0) BTC_TRACE(5, 8 * ly + 4);
#pragma unroll
                    for (int j = 0; j < CH / 4; ++j) rs[j] = rn[j];
                }
            } else {
                static_assert(CPT == CH || NB == 1, "the chunked E3 path handles one block per step");
                const int x = xb0;
                const bool xin = min_ok && (x < W);
#pragma unroll
                for (int j = 0; j < CH / 4; ++j) rs[j] = (xin && has_res) ? res_ld(j, y, 0) : make_float4(0.f, 0.f, 0.f, 0.f);
                BTC_WAIT(b3_full + 8 * sa, (uint32_t)((ly >> (NA3 - 1)) & 1));
                tc_fence_after();
                if (tid == 0) BTC_TRACE(4, 4 * ly + 1);
                const uint32_t trow = trow0 + sa * N3;
                if (tid == 0) BTC_TRACE(4, 4 * ly + 2);
#pragma unroll
                for (int c0 = 0; c0 < CPT; c0 += CH) {
                    if (c0 + CH < CPT) {        // next cout chunk of this row
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j)
                            rn[j] = (xin && has_res) ? res_ld((c0 + CH) / 4 + j, y, 0) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    float bb[CH];
#pragma unroll
                    for (int i = 0; i < CH; i += 4) {
                        const float4 d4 = btc_lds128(b3p + 4 * (c0 + i));
                        bb[i] = d4.x; bb[i + 1] = d4.y; bb[i + 2] = d4.z; bb[i + 3] = d4.w;
                    }
                    uint32_t u0[CH], u1[CH], u2[CH];
                    tmem_ld8_nowait(trow + (uint32_t)(0 * C + c0), u0);
                    tmem_ld8_nowait(trow + (uint32_t)(1 * C + c0), u1);
                    tmem_ld8_nowait(trow + (uint32_t)(2 * C + c0), u2);
                    tmem_ld_wait();
                    tmem_ld_fence_regs<CH>(u0); tmem_ld_fence_regs<CH>(u1); tmem_ld_fence_regs<CH>(u2);
                    float v0[CH], v1[CH], v2[CH];
#pragma unroll
                    for (int i = 0; i < CH; ++i) { v0[i] = __uint_as_float(u0[i]); v1[i] = __uint_as_float(u1[i]); v2[i] = __uint_as_float(u2[i]); }
                    if (c0 + CH >= CPT) {           // last TMEM read of this accumulator
                        tc_fence_before();
                        mbar_arrive_a(b3_empty + 8 * sa);
                    }
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const float lv = __shfl_up_sync(0xffffffffu, v0[i], 1);
                        const float rv = __shfl_down_sync(0xffffffffu, v2[i], 1);
                        v1[i] = sgn * (((lv + v1[i]) + rv) * (1.0f / VST_HALF_SCALE) + bb[i]);
                    }
                    if (xin) {
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) {
                            const float4 rr = rs[j];
                            float4 o;
                            o.x = rr.x + v1[4 * j]; o.y = rr.y + v1[4 * j + 1]; o.z = rr.z + v1[4 * j + 2]; o.w = rr.w + v1[4 * j + 3];
                            out_st(c0 / 4 + j, y, x, o);
                        }
                    }
                    if (c0 + CH < CPT) {
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) rs[j] = rn[j];
                    }
                }
            }
            if (tid == 0) BTC_TRACE(4, 4 * ly + 3);
        }
        if (tid == 0) BTC_ACC_END((a.trace && (int)blockIdx.x == a.trace_cta) ? a.trace : nullptr, 4, sg.yb - sg.ya);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == Cfg::W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int C, int MODE>
static int launch_block_tc_cfg(BlockTcArgs a, cudaStream_t st) {
    using Cfg = BtcCfg<C>;
    static PerDeviceOnce smem_once;
    auto kern = rev_block_tc_kernel<C, MODE>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    a.n_strips = cdiv(a.W, Cfg::NB * Cfg::XO);
    int nseg = std::max(1, num_sms() / a.n_strips);
    a.rows_per_seg = cdiv(a.H, nseg);
    if (a.rows_per_seg < 8) a.rows_per_seg = std::min(8, a.H);
    nseg = cdiv(a.H, a.rows_per_seg);
    a.trace = tc_trace_buffer(1000 + C, 0, st);
#ifdef BTC_KNOCKOUT
    { const char* e = getenv("VST_BTC_KO"); a.ko = e ? atoi(e) : 0; }
#endif
    a.trace_cta = std::min(a.n_strips * nseg - 1, a.n_strips * (nseg / 2) + a.n_strips / 2);
    const double px = (double)a.H * a.W;
    ProfScope prof(st, C == 16 ? "rev_block_tc 16>4>4>16" : "rev_block_tc 64>16>16>64",
                   2.0 * 9 * (2.0 * C * Cfg::M + Cfg::M * Cfg::M) * px, (a.res ? 3.0 : 2.0) * 4.0 * C * px);
    VST_CUDA_OK(launch_pdl(kern, a.n_strips * nseg, Cfg::THREADS, Cfg::SMEM, st, a));
    return check_launch("rev_block_tc");
}

int launch_rev_block_tc(int C, const float* x, const float* res, float* out, const float* wpack, int H, int W, int sub,
                        int* status, cudaStream_t st, int mode) {
    VST_REQUIRE(C == 16 || C == 64, "rev_block_tc: C = %d not supported", C);
    VST_REQUIRE(H >= 2 && W >= 4, "rev_block_tc: map %dx%d too small", H, W);
    VST_REQUIRE(mode == BTC_PLAIN || (H % 2 == 0 && W % 2 == 0 && H >= 4 && W >= 4 && res != nullptr && res != out),
                "rev_block_tc: squeeze modes need even H, W >= 4 and a coupling operand that does not alias out");
    BlockTcArgs a;
    a.x = x; a.res = res; a.out = out; a.wpack = reinterpret_cast<const uint8_t*>(wpack);
    a.ko = 0; a.H = H; a.W = W; a.sub = sub; a.status = status; a.n_strips = 0; a.rows_per_seg = 0; a.trace = nullptr; a.trace_cta = 0;
    if (mode == BTC_OUT_SQZ) return C == 16 ? launch_block_tc_cfg<16, BTC_OUT_SQZ>(a, st) : launch_block_tc_cfg<64, BTC_OUT_SQZ>(a, st);
    if (mode == BTC_RES_UNSQZ) return C == 16 ? launch_block_tc_cfg<16, BTC_RES_UNSQZ>(a, st) : launch_block_tc_cfg<64, BTC_RES_UNSQZ>(a, st);
    return C == 16 ? launch_block_tc_cfg<16, BTC_PLAIN>(a, st) : launch_block_tc_cfg<64, BTC_PLAIN>(a, st);
}

}  // namespace vst
