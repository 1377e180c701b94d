// gram_tc.cu — cWCT statistics of an unmasked feature map on the tensor cores:
//   sum[c] = sum_p (x[c,p] - pivot[c])          gram[c,c'] = sum_p (x[c,p] - pivot[c]) (x[c',p] - pivot[c'])
// Replaces mean / centring / x @ x.T of models/cWCT.py:138-144, :153-157 (and :223-226, :241-244) for the
// unmasked paths; the per-label variant stays on the CUDA-core kernel of cwct.cu.
//
// Two sources (template SRC):
//   SRC_NCHW  — the latent as the reference API exchanges it, [C][n] fp32 rows;
//   SRC_STATE — the network's own P4 half-states (x1 | x2) BEFORE the channel_reduction spread
//               (models/RevResNet.py:140-146): latent channel c of sub-position j = state channel j*C + c, so the
//               statistics of z are those of the state read as 4^sp_steps * h * w pixels of C channels; the fused video
//               path never materialises z.
//
// The Gram is a GEMM with M = N = C and K = pixels.  A tcgen05 UMMA is M = 128 wide, so for C < 128 the M (and N)
// dimension is filled with SEGS = 128 / C SEGMENTS — pixel runs (NCHW) or sub-positions (STATE): operand row (s, c)
// holds channel c of segment s, and the diagonal C x C blocks of the 128 x 128 accumulator are the partial Grams of
// the segments (the off-diagonal blocks mix segments and are ignored; the pass stays HBM-bound, profiles/).
//
// Pipeline per CTA (persistent, a contiguous range of stages):
//   * producer (1 lane): ONE tensor-map TMA copy per stage (cp.async.bulk.tensor.2d / .3d, 16 KB) into a 6-deep RAW
//     ring: up to 96 KB in flight per SM — the memory-level parallelism an HBM-bound pass needs (the register-staged
//     loads of the first version kept 16 KB per SM in flight and reached 2.0 TB/s).
//   * converters (8 warps): RAW -> subtract the pivot (fused mean subtraction; the exact mean correction is applied
//     in the factor kernel from `sum`), split x = hi + lo (tf32 + remainder), transpose P4 units if needed, store both
//     terms in the no-swizzle K-major canonical layout [k-chunk][row][4 floats] (row pitch padded: conflict-free).
//   * UMMA issuer (1 thread): per stage 4 k-steps x 2 UMMAs, kind::tf32, fp32 accumulators in TMEM, two buffers:
//     D = H H^T + H (2L)^T.  The consumer (factor kernel) averages the two triangles of the accumulated matrix, and
//     (D + D^T) / 2 = H H^T + H L^T + L H^T — the fp32-equivalent three-term product — so the third UMMA of the
//     symmetric pair is never issued (a UMMA is bound by fetching its operands from shared memory: -1/3 of that traffic).
//   * drain warps (4): every FL stages (split-K) pull the finished buffer out of TMEM and fold it into fp64
//     (registers for C <= 64, global atomics for C = 128); at the end one fp64 atomicAdd per entry and CTA.
#include <stdlib.h>
#include "kernels.cuh"
#include "tma_map.cuh"

namespace vst {

namespace gtc {
constexpr int KT = 32;                       // pixels per segment and stage
constexpr int ROWP = 129;                    // padded rows per k-chunk (128 + 1): conflict-free 16-byte stores
constexpr int TILE_BYTES = (KT / 4) * ROWP * 16;         // one term (hi or lo) of one stage
constexpr int OP_BYTES = 2 * TILE_BYTES;
constexpr int RAW_BYTES = 128 * KT * 4;      // 128 rows x 32 pixels fp32 (any segment count)
constexpr int NO = 3;                        // operand ring depth
constexpr int NR = 6;                        // RAW ring depth (TMA copies in flight)
constexpr size_t SMEM = (size_t)NO * OP_BYTES + (size_t)NR * RAW_BYTES + 1024 + 256;
constexpr int THREADS = 448;                 // 8 converter warps, 4 drain warps, UMMA issuer, TMA producer
constexpr int W_DRAIN = 8, W_MMA = 12, W_TMA = 13;
constexpr int SRC_NCHW = 0, SRC_STATE = 1;
}  // namespace gtc

struct GramTcArgs {
    const float* pivot;      // [C]
    double* count;           // [1]
    double* sum;             // [C]
    double* gram;            // [C*C]
    int C;
    long long n;             // pixels the statistics run over
    int raw_tx_bytes;        // bytes one TMA box delivers (the full 16 KB unless C < CP)
    // SRC_NCHW: stage = pixel run [st * SEGS*KT, +SEGS*KT)
    // SRC_STATE: stage = (row y, x-block of KT pixels, sub-position block); rows of w interior pixels
    int h, w, n_xb, n_jb, groups_per_half;
    int stages_per_cta, n_stages;
};

template <int SEGS, int SRC>
__global__ void __launch_bounds__(gtc::THREADS, 1)
gram_tc_kernel(GramTcArgs a, const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1) {
    using namespace gtc;
    constexpr int CP = 128 / SEGS;               // padded channels per segment
    constexpr int FL = (CP == 128) ? 32 : 8;     // stages per accumulator flush (split-K granularity)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // round up by OFFSET, not through an integer cast: pointer arithmetic on the __shared__ array keeps the address space,
    // so the converters' accesses compile to LDS / STS instead of generic LD / ST
    uint8_t* op_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* raw_base = op_base + (size_t)NO * OP_BYTES;
    uint64_t* bars = (uint64_t*)(raw_base + (size_t)NR * RAW_BYTES);
    uint64_t* raw_full = bars;                 // [NR] producer arrive.expect_tx + TMA bytes
    uint64_t* raw_empty = bars + 8;            // [NR] 256 converter threads
    uint64_t* op_full = bars + 16;             // [NO] 256 converter threads
    uint64_t* op_empty = bars + 20;            // [NO] tcgen05.commit
    uint64_t* acc_full = bars + 24;            // [2]  tcgen05.commit
    uint64_t* acc_empty = bars + 26;           // [2]  128 drain threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 28);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int st_begin = blockIdx.x * a.stages_per_cta;
    const int st_end = min(st_begin + a.stages_per_cta, a.n_stages);
    const int n_stages = max(st_end - st_begin, 0);

    if (tid == 0) {
        for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 256); }
        for (int s = 0; s < NO; ++s) { mbar_init(&op_full[s], 256); mbar_init(&op_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ================= producer: one tensor-map TMA copy per stage =================
        if (lane == 0) {
            for (int i = 0; i < n_stages; ++i) {
                const int st = st_begin + i, s = i % NR;
                mbar_wait(&raw_empty[s], ((i / NR) & 1) ^ 1);
                uint8_t* dst = raw_base + (size_t)s * RAW_BYTES;
                mbar_arrive_expect_tx(&raw_full[s], (uint32_t)a.raw_tx_bytes);
                if (SRC == SRC_NCHW) {
                    // box {SEGS*KT pixels, min(C, CP) rows} at pixel st*SEGS*KT (pixels >= n arrive as zeros)
                    tma_load_2d(dst, &tm0, st * (SEGS * KT), 0, &raw_full[s]);
                } else {
                    // box {KT pixels x 4 floats, 1 row, 32 groups} of the half-state that holds sub-position block jb
                    // x-block fastest: consecutive boxes continue the same 32 DRAM streams.  Knock-out builds (no converter
                    // work / no UMMAs / neither): 0.084 / 0.102 / 0.079 ms against 0.119 ms; ring depth 8 and this order
                    // change it by < 3 %.  (tools/ubench/tma_stream.cu: the same boxes with no consumer warps at all take
                    // 0.058 ms, whatever the box shape — the plane layout is not what limits the stream.)
                    const int xb = st % a.n_xb, r = st / a.n_xb;
                    const int jb = r % a.n_jb, y = r / a.n_jb;
                    const int g0 = jb * 32;                                 // first state group of the block
                    const bool second = g0 >= a.groups_per_half;
                    tma_load_3d(dst, second ? &tm1 : &tm0, (xb * KT + 1) * 4, y + 1, second ? g0 - a.groups_per_half : g0,
                                &raw_full[s]);
                }
            }
        }
    } else if (warp < W_DRAIN) {
        // ================= converters: RAW fp32 -> (x - pivot) = hi + lo, K-major operand rows =================
        float ssum[4] = {0.f, 0.f, 0.f, 0.f};
        float piv[4];
        int crow[4];                                   // STATE: operand row of each of this thread's 4 channels
        if (SRC == SRC_NCHW) {
            // item i: channel c = (i * 256 + tid) / (SEGS * 8), 16-byte unit u = ... % (SEGS * 8) of the channel's run
            constexpr int UPC = SEGS * (KT / 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = (i * 256 + tid) / UPC;
                crow[i] = c;
                piv[i] = c < a.C ? __ldg(a.pivot + c) : 0.f;
            }
        } else {
            // one item: group g = tid / 8 (of the 32 staged groups), pixel block j4 = tid % 8; channels 4g .. 4g+3
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                crow[e] = 4 * (tid >> 3) + e;
                piv[e] = __ldg(a.pivot + (crow[e] % CP));
            }
        }
        for (int i = 0; i < n_stages; ++i) {
            const int st = st_begin + i, rs = i % NR, os = i % NO;
            mbar_wait(&raw_full[rs], (i / NR) & 1);
            mbar_wait(&op_empty[os], ((i / NO) & 1) ^ 1);
            const float4* raw = reinterpret_cast<const float4*>(raw_base + (size_t)rs * RAW_BYTES);
            float4* hi = reinterpret_cast<float4*>(op_base + (size_t)os * OP_BYTES);
            float4* lo = reinterpret_cast<float4*>(op_base + (size_t)os * OP_BYTES + TILE_BYTES);
            if (SRC == SRC_NCHW) {
                constexpr int UPC = SEGS * (KT / 4);
                const long long p_stage = (long long)st * (SEGS * KT);
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int idx = it * 256 + tid, c = idx / UPC, u = idx % UPC;
                    const int seg = u / (KT / 4), j = u % (KT / 4);
                    const bool ok = c < a.C && p_stage + 4 * u < a.n;      // n % 4 == 0: a unit is inside or outside
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) {                                              // rows >= C were not written by the copy
                        v = raw[c * UPC + u];
                        const float pv = piv[it];
                        v.x -= pv; v.y -= pv; v.z -= pv; v.w -= pv;
                    }
                    ssum[it] += (v.x + v.y) + (v.z + v.w);
                    const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                    const int dst = j * ROWP + seg * CP + c;
                    hi[dst] = h;
                    lo[dst] = make_float4(2.f * (v.x - h.x), 2.f * (v.y - h.y), 2.f * (v.z - h.z), 2.f * (v.w - h.w));
                }
            } else {
                const int g = tid >> 3, j4 = tid & 7;
                const int xb = st % a.n_xb;
                const int x0 = xb * KT + 4 * j4;                           // image column of this item's first pixel
                float4 p[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) p[q] = raw[g * KT + 4 * j4 + q];
                const float in[4][4] = {{p[0].x, p[1].x, p[2].x, p[3].x}, {p[0].y, p[1].y, p[2].y, p[3].y},
                                        {p[0].z, p[1].z, p[2].z, p[3].z}, {p[0].w, p[1].w, p[2].w, p[3].w}};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = (x0 + q < a.w) ? in[e][q] - piv[e] : 0.f;   // past the row: border / other data
                    ssum[e] += (v[0] + v[1]) + (v[2] + v[3]);
                    const float4 h = make_float4(tf32_round(v[0]), tf32_round(v[1]), tf32_round(v[2]), tf32_round(v[3]));
                    const int dst = j4 * ROWP + crow[e];
                    hi[dst] = h;
                    lo[dst] = make_float4(2.f * (v[0] - h.x), 2.f * (v[1] - h.y), 2.f * (v[2] - h.z), 2.f * (v[3] - h.w));
                }
            }
            fence_proxy_async();
            mbar_arrive(&op_full[os]);
            mbar_arrive(&raw_empty[rs]);
        }
        // per-channel sums
        if (SRC == SRC_NCHW) {
            constexpr int UPC = SEGS * (KT / 4);       // lanes that share a channel: UPC consecutive lanes (UPC <= 32)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float v = ssum[it];
#pragma unroll
                for (int o = UPC / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int idx = it * 256 + tid, c = idx / UPC, u = idx % UPC;
                if (u == 0 && c < a.C && n_stages > 0) atomicAdd(a.sum + c, (double)v);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = ssum[e];
#pragma unroll
                for (int o = 4; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);       // the 8 pixel blocks of a group
                if ((tid & 7) == 0 && n_stages > 0) atomicAdd(a.sum + (crow[e] % CP), (double)v);
            }
        }
        if (tid == 0 && blockIdx.x == 0) atomicAdd(a.count, (double)a.n);
    } else if (warp == W_MMA) {
        // ================= UMMA issuer =================
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t LBO = ROWP * 16, SBO = 128;
            for (int st = 0; st < n_stages; ++st) {
                const int s = st % NO, grp = st / FL;
                const uint32_t b = grp & 1;
                if (st % FL == 0) {
                    mbar_wait(&acc_empty[b], ((grp >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                mbar_wait(&op_full[s], (st / NO) & 1);
                tc_fence_after();
                const uint32_t Hi = smem_u32(op_base + (size_t)s * OP_BYTES), Lo = Hi + TILE_BYTES;
                const uint32_t d = tmem_base + b * 128;
#pragma unroll
                for (int ks = 0; ks < KT / 8; ++ks) {
                    const uint64_t dh = make_desc(Hi + ks * 2 * LBO, LBO, SBO), dl = make_desc(Lo + ks * 2 * LBO, LBO, SBO);
                    umma_tf32(d, dh, dh, IDESC, (st % FL != 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(d, dh, dl, IDESC, 1u);       // H (2L)^T: with the symmetrisation below = H L^T + L H^T
                }
                umma_commit(&op_empty[s]);
                if (st % FL == FL - 1 || st == n_stages - 1) umma_commit(&acc_full[b]);
            }
        }
        __syncwarp();
    } else {
        // ================= drain: TMEM -> fp64 =================
        const int q = warp & 3;                          // TMEM lane quarter of this warp (warps 8..11)
        const int row = q * 32 + lane;                   // operand row = (segment, channel)
        const int seg = row / CP, c = row % CP;
        constexpr bool REGS = CP <= 64;                  // fp64 partials in registers, else straight to global
        double dacc[REGS ? CP : 1];
#pragma unroll
        for (int i = 0; i < (REGS ? CP : 1); ++i) dacc[i] = 0.0;
        const int n_groups = (n_stages + FL - 1) / FL;
        for (int grp = 0; grp < n_groups; ++grp) {
            const uint32_t b = grp & 1;
            mbar_wait(&acc_full[b], (grp >> 1) & 1);
            tc_fence_after();
            const uint32_t t = tmem_base + ((uint32_t)(q * 32) << 16) + b * 128 + seg * CP;   // this segment's diagonal block
#pragma unroll
            for (int c0 = 0; c0 < CP; c0 += 32) {
                float v[32];
                tmem_ld<32>(t + c0, v);
                if (REGS) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) dacc[c0 + i] += (double)v[i];
                } else if (c < a.C) {
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i < a.C) atomicAdd(a.gram + (size_t)c * a.C + c0 + i, (double)v[i]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[b]);
        }
        if (REGS && c < a.C && n_groups > 0) {
#pragma unroll
            for (int i = 0; i < CP; ++i)
                if (i < a.C) atomicAdd(a.gram + (size_t)c * a.C + i, dacc[i]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------
// Per-label (masked) statistics on the tensor cores — models/cWCT.py:87-95 (np.where + index_select per label) and
// :138-144 for all labels in ONE pass over the feature map.
//
// Same pipeline as above (tensor-map TMA -> RAW ring -> converters -> UMMA -> drain), with three differences:
//   * a stage is 128 pixels of one image row; a CTA owns a contiguous run of stages in row-major order (column-group
//     orders were measured slower: memory locality of consecutive boxes matters more than label locality);
//   * a stage is multiplied once per label PRESENT in it (one pass for almost every stage): the converters zero the
//     pixels of the other labels, so every pass is the masked Gram of one label — "each label gets its own
//     masked-covariance segment";
//   * every pass has its own TMEM accumulator (two, alternating) and is folded into fp64 by the drain warps right
//     away (fp32 accumulation never spans more than 32 pixels per segment); the drain keeps the running fp64 Gram of
//     the CURRENT label in registers and flushes it with atomics only when the label changes.
// ------------------------------------------------------------------------------------------
#ifdef GTC_TRACE
#define GTC_STAMP(role, idx) do { if (blockIdx.x == 3 && (idx) < 40) gtc_trace[(role) * 40 + (idx)] = clock64(); } while (0)
__device__ long long gtc_trace[8 * 40];
#else
#define GTC_STAMP(role, idx) do { } while (0)
#endif
struct GramTcMaskedArgs {
    const float* pivot;       // [C]
    const uint8_t* labels;    // [H*W]
    double* count;            // [L]
    double* sum;              // [L*C]
    double* gram;             // [L*C*C]
    int C, L, H, W, n_xb;     // n_xb: strips of SEGS*KT pixels per row
    int gw;                   // strips per column group (traversal: group by group, row by row, strip by strip)
    int raw_tx_bytes, stages_per_cta, n_stages;
};

template <int SEGS>
__global__ void __launch_bounds__(gtc::THREADS, 1)
gram_tc_masked_kernel(GramTcMaskedArgs a, const __grid_constant__ CUtensorMap tm0) {
    using namespace gtc;
    // ring depths here (shadow gtc::NO / gtc::NR): a strip-major box (32 channel rows 7.7 KB apart from the previous stage's)
    // takes ~8 us to arrive, so the RAW ring is what sets the stage rate: 7 boxes in flight, 2 operand slots
    constexpr int NO = 2, NR = 7;
    constexpr int CP = 128 / SEGS, BOXP = SEGS * KT, UPC = SEGS * (KT / 4);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // round up by OFFSET, not through an integer cast: pointer arithmetic on the __shared__ array keeps the address space,
    // so the converters' accesses compile to LDS / STS instead of generic LD / ST
    uint8_t* op_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* raw_base = op_base + (size_t)NO * OP_BYTES;
    uint64_t* bars = (uint64_t*)(raw_base + (size_t)NR * RAW_BYTES);
    uint64_t* raw_full = bars;                 // [NR]
    uint64_t* raw_empty = bars + 8;            // [NR] 256 converter threads (after the LAST pass of the stage)
    uint64_t* op_full = bars + 16;             // [NO] 256 converter threads
    uint64_t* op_empty = bars + 20;            // [NO] tcgen05.commit
    uint64_t* acc_full = bars + 24;            // [2]  tcgen05.commit (every pass)
    uint64_t* acc_empty = bars + 26;           // [2]  128 drain threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 28);
    __shared__ int pass_label[8];              // label of pass i (ring), -1 terminates the pass stream
    __shared__ __align__(16) unsigned int pres[3 * 8];     // labels present in a stage (three sets, see the converters)
    __shared__ __align__(16) int stage_lab[3 * 128];       // label of every pixel of a stage (-1: outside)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int st_begin = blockIdx.x * a.stages_per_cta;
    const int st_end = min(st_begin + a.stages_per_cta, a.n_stages);
    const int n_stages = max(st_end - st_begin, 0);

    if (tid == 0) {
        for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 256); }
        for (int s = 0; s < NO; ++s) { mbar_init(&op_full[s], 256); mbar_init(&op_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // stage st -> (column group g, row y, strip xb): pixel run [y * W + xb * BOXP, + BOXP), the first min(BOXP, W - xb*BOXP) count
    if (warp == W_TMA) {
        if (lane == 0) {
            for (int i = 0; i < n_stages; ++i) {
                const int st = st_begin + i, s = i % NR;
                const int g = st / (a.gw * a.H), r = st - g * (a.gw * a.H), y = r / a.gw;
                const int xb = min(g * a.gw + (r - y * a.gw), a.n_xb - 1);      // (a stage past the last strip re-reads it: no pixels)
                mbar_wait(&raw_empty[s], ((i / NR) & 1) ^ 1);
                mbar_arrive_expect_tx(&raw_full[s], (uint32_t)a.raw_tx_bytes);
                // two copies of 16 channel rows each (the rows of a box are fetched one after the other)
                tma_load_2d(raw_base + (size_t)s * RAW_BYTES, &tm0, y * a.W + xb * BOXP, 0, &raw_full[s]);
                tma_load_2d(raw_base + (size_t)s * RAW_BYTES + RAW_BYTES / 2, &tm0, y * a.W + xb * BOXP, 16, &raw_full[s]);
            }
        }
    } else if (warp < W_DRAIN) {
        // ================= converters =================
        float ssum[4] = {0.f, 0.f, 0.f, 0.f};
        float piv[4];
        int cnt = 0, cur = -1, pass = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = (i * 256 + tid) / UPC;
            piv[i] = c < a.C ? __ldg(a.pivot + c) : 0.f;
        }
        auto flush_sums = [&](int l) {
            if (l < 0 || l >= a.L) return;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float v = ssum[it];
#pragma unroll
                for (int o = UPC / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int idx = it * 256 + tid, c = idx / UPC, u = idx % UPC;
                if (u == 0 && c < a.C) atomicAdd(a.sum + (size_t)l * a.C + c, (double)v);
                ssum[it] = 0.f;
            }
            int cv = cnt;                                   // pixels counted by the threads of channel 0 (tid < UPC)
#pragma unroll
            for (int o = 16; o; o >>= 1) cv += __shfl_xor_sync(0xffffffffu, cv, o);
            if (tid == 0 && cv > 0) atomicAdd(a.count + l, (double)cv);
            cnt = 0;
        };
        // label of pixel `tid` of a stage (threads 0 .. BOXP-1), fetched ONE STAGE AHEAD: the global-load latency hides
        // behind the current stage's passes
        auto load_label = [&](int i) -> int {
            if (tid >= BOXP || i >= n_stages) return -1;
            const int st = st_begin + i;
            const int g = st / (a.gw * a.H), r = st - g * (a.gw * a.H), y = r / a.gw, xb = g * a.gw + (r - y * a.gw);
            const int x0 = xb * BOXP;
            if (xb >= a.n_xb || tid >= a.W - x0) return -1;
            return (int)__ldg(a.labels + (long long)y * a.W + x0 + tid);     // consumed (and range-checked) one stage later
        };
        int lab_next = load_label(0);
        if (tid < 24) pres[tid] = 0u;
        named_barrier(1, 256);
        for (int i = 0; i < n_stages; ++i) {
            const int rs = i % NR;
            // labels of the stage and the set present (all 256 converter threads agree through shared memory).  Three
            // sets, ONE barrier per stage: set i % 3 is written here, read by this stage's passes after the barrier; set
            // (i + 2) % 3 — last read by stage i - 1, which every thread has left once it is past this barrier — is
            // cleared right after it, two barriers before stage i + 2 writes it.
            unsigned int* pres_i = pres + 8 * (i % 3);
            int* lab_i = stage_lab + 128 * (i % 3);
            if (tid < BOXP) {
                const int l = lab_next < a.L ? lab_next : -1;
                // one shared-memory atomic per distinct label and warp (128 same-address atomics would serialise)
                const unsigned peers = __match_any_sync(0xffffffffu, l);
                if (l >= 0 && (__ffs(peers) - 1) == lane) atomicOr(&pres_i[l >> 5], 1u << (l & 31));
                lab_i[tid] = l;
            }
            lab_next = load_label(i + 1);
            if (tid == 0) GTC_STAMP(6, i);
            if (tid == 255) GTC_STAMP(7, i);
            named_barrier(1, 256);
            if (tid < 8) pres[8 * ((i + 2) % 3) + tid] = 0u;
            if (tid == 0) GTC_STAMP(1, i);
            mbar_wait(&raw_full[rs], (i / NR) & 1);
            if (tid == 0) GTC_STAMP(2, i);
            const float4* raw = reinterpret_cast<const float4*>(raw_base + (size_t)rs * RAW_BYTES);
            const uint4 pa = *reinterpret_cast<const uint4*>(pres_i), pb = *reinterpret_cast<const uint4*>(pres_i + 4);
            const unsigned int pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) {
                unsigned int bits = pw[w8];
                while (bits) {
                    const int lbl = w8 * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (lbl != cur) { flush_sums(cur); cur = lbl; }
                    const int os = pass % NO;
                    mbar_wait(&op_empty[os], ((pass / NO) & 1) ^ 1);
                    float4* hi = reinterpret_cast<float4*>(op_base + (size_t)os * OP_BYTES);
                    float4* lo = reinterpret_cast<float4*>(op_base + (size_t)os * OP_BYTES + TILE_BYTES);
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int idx = it * 256 + tid, c = idx / UPC, u = idx % UPC;
                        const int seg = u / (KT / 4), j = u % (KT / 4);
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < a.C) {
                            const float4 r4 = raw[c * UPC + u];
                            const float pv = piv[it];
                            const int4 lb = *reinterpret_cast<const int4*>(lab_i + 4 * u);
                            v.x = lb.x == lbl ? r4.x - pv : 0.f; v.y = lb.y == lbl ? r4.y - pv : 0.f;
                            v.z = lb.z == lbl ? r4.z - pv : 0.f; v.w = lb.w == lbl ? r4.w - pv : 0.f;
                            if (c == 0) cnt += (lb.x == lbl) + (lb.y == lbl) + (lb.z == lbl) + (lb.w == lbl);
                        }
                        ssum[it] += (v.x + v.y) + (v.z + v.w);
                        const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                        const int dst = j * ROWP + seg * CP + c;
                        hi[dst] = h;
                        lo[dst] = make_float4(2.f * (v.x - h.x), 2.f * (v.y - h.y), 2.f * (v.z - h.z), 2.f * (v.w - h.w));
                    }
                    if (tid == 0) pass_label[pass & 7] = lbl;
                    fence_proxy_async();
                    mbar_arrive(&op_full[os]);
                    if (tid == 0) GTC_STAMP(0, pass);
                    if (tid == 255) GTC_STAMP(5, pass);
                    ++pass;
                }
            }
            mbar_arrive(&raw_empty[rs]);
        }
        flush_sums(cur);
        // terminate the pass stream for the issuer and the drain warps
        {
            const int os = pass % NO;
            mbar_wait(&op_empty[os], ((pass / NO) & 1) ^ 1);
            if (tid == 0) pass_label[pass & 7] = -1;
            fence_proxy_async();
            mbar_arrive(&op_full[os]);
        }
    } else if (warp == W_MMA) {
        // ================= UMMA issuer: one accumulator per pass =================
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t LBO = ROWP * 16, SBO = 128;
            for (int pass = 0;; ++pass) {
                const int s = pass % NO;
                const uint32_t b = pass & 1;
                mbar_wait(&op_full[s], (pass / NO) & 1);
                if (*(volatile int*)&pass_label[pass & 7] < 0) break;
                GTC_STAMP(3, pass);
                mbar_wait(&acc_empty[b], ((pass >> 1) & 1) ^ 1);
                tc_fence_after();
                GTC_STAMP(4, pass);
                const uint32_t Hi = smem_u32(op_base + (size_t)s * OP_BYTES), Lo = Hi + TILE_BYTES;
                const uint32_t d = tmem_base + b * 128;
#pragma unroll
                for (int ks = 0; ks < KT / 8; ++ks) {
                    const uint64_t dh = make_desc(Hi + ks * 2 * LBO, LBO, SBO), dl = make_desc(Lo + ks * 2 * LBO, LBO, SBO);
                    umma_tf32(d, dh, dh, IDESC, ks > 0 ? 1u : 0u);
                    umma_tf32(d, dh, dl, IDESC, 1u);
                }
                umma_commit(&op_empty[s]);
                umma_commit(&acc_full[b]);
            }
            // wake the drain warps with the terminator: they read pass_label of the pass they wait for
            umma_commit(&acc_full[0]);
            umma_commit(&acc_full[1]);
        }
        __syncwarp();
    } else {
        // ================= drain: TMEM -> fp64 per label =================
        // The running fp64 Grams of up to NSLOT labels live in shared memory (a stage where two regions meet alternates
        // between two labels pass by pass: register accumulators of ONE label would be flushed with 4096 global atomics
        // on every pass).  The four drain warps are the four segments of a pass and add their diagonal blocks into the
        // label's slot with shared-memory atomics; a slot is flushed to global memory (8 entries per thread) only when a
        // label outside the cache shows up, least recently used first, and at the end.
        static_assert(CP == 32, "the per-label drain keeps 32 x 32 fp64 slots");
        constexpr int NSLOT = 3;
        double* lacc = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 2048);     // [NSLOT][32][32] fp64
        float* stg = reinterpret_cast<float*>(lacc + NSLOT * 1024);                            // [4 segments][32][32] fp32
        const int q = warp & 3;
        const int dt = tid - W_DRAIN * 32;               // 0..127: owns entries dt*8 .. dt*8+7 of every slot
        const int c = lane;                              // operand row = (segment q, channel c)
        int slot_label[NSLOT], slot_age[NSLOT];
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) { slot_label[k] = -1; slot_age[k] = 0; }
        for (int i = dt; i < NSLOT * 1024; i += 128) lacc[i] = 0.0;
        auto flush_slot = [&](int k) {                   // every thread flushes the entries it owns: no barrier needed
            const int l = slot_label[k];
            if (l >= 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int idx = dt * 8 + e, i = idx >> 5, j = idx & 31;
                    const double v = lacc[k * 1024 + idx];
                    if (i < a.C && j < a.C && v != 0.0) atomicAdd(a.gram + ((size_t)l * a.C + i) * a.C + j, v);
                    lacc[k * 1024 + idx] = 0.0;
                }
            }
        };
        for (int pass = 0;; ++pass) {
            const uint32_t b = pass & 1;
            mbar_wait(&acc_full[b], (pass >> 1) & 1);
            const int lbl = *(volatile int*)&pass_label[pass & 7];
            if (lbl < 0) break;
            tc_fence_after();
            // slot of this label (every drain thread runs the same deterministic LRU cache)
            int k = -1;
#pragma unroll
            for (int s2 = 0; s2 < NSLOT; ++s2) if (slot_label[s2] == lbl) k = s2;
            if (k < 0) {
                k = 0;
#pragma unroll
                for (int s2 = 1; s2 < NSLOT; ++s2) if (slot_age[s2] < slot_age[k]) k = s2;
                flush_slot(k);
                slot_label[k] = lbl;
            }
            slot_age[k] = pass + 1;
            const uint32_t t = tmem_base + ((uint32_t)(q * 32) << 16) + b * 128 + q * CP;     // this segment's diagonal block
            float v[32];
            tmem_ld<32>(t, v);
            tc_fence_before();
            mbar_arrive(&acc_empty[b]);                  // the accumulator is in registers: the issuer may reuse it
            // the four segments (warps) of the pass meet in shared memory; every thread then folds the 8 entries it
            // owns into the label's fp64 slot (no atomics, no contention)
            named_barrier(2, 128);                       // the previous pass's staging has been consumed
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(stg + (q * 32 + c) * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            named_barrier(2, 128);
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
                const int idx = dt * 8 + e;
                float4 s4 = *reinterpret_cast<const float4*>(stg + idx);
#pragma unroll
                for (int sg = 1; sg < 4; ++sg) {
                    const float4 o4 = *reinterpret_cast<const float4*>(stg + sg * 1024 + idx);
                    s4.x += o4.x; s4.y += o4.y; s4.z += o4.z; s4.w += o4.w;
                }
                double* dst = lacc + k * 1024 + idx;
                dst[0] += (double)s4.x; dst[1] += (double)s4.y; dst[2] += (double)s4.z; dst[3] += (double)s4.w;
            }
        }
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) flush_slot(k);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        tried = true;
    }
    return fn;
}

int make_tensor_map_f32(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box) {
    EncodeTiledFn fn = encode_tiled_fn();
    VST_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

template <int SEGS, int SRC>
static int launch_gram_tc_cfg(GramTcArgs a, const CUtensorMap& tm0, const CUtensorMap& tm1, cudaStream_t st) {
    static PerDeviceOnce smem_once;
    auto kern = gram_tc_kernel<SEGS, SRC>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)gtc::SMEM));
    int grid = std::min(a.n_stages, num_sms());
    a.stages_per_cta = cdiv(a.n_stages, grid);
    grid = cdiv(a.n_stages, a.stages_per_cta);
    kern<<<grid, gtc::THREADS, gtc::SMEM, st>>>(a, tm0, tm1);
    return check_launch("cwct_gram_tc");
}

// small maps stay on the CUDA-core kernels: exact fp32 products (a two-term tf32 split carries 22 bits), and a
// persistent tensor-core pipeline has nothing to amortise there
bool gram_tc_eligible(int C, long long n) { return C >= 1 && C <= 128 && n % 4 == 0 && n >= 16384 && n < (1ll << 31); }

// sums and Gram of the pivot-shifted features into zero-initialised fp64 buffers (one label)
int launch_gram_tc(const float* feat, const float* pivot, double* count, double* sum, double* gram, int C, long long n,
                   cudaStream_t st) {
    VST_REQUIRE(gram_tc_eligible(C, n) && (((uintptr_t)feat) & 15) == 0, "gram_tc: C=%d n=%lld not eligible", C, n);
    GramTcArgs a = {};
    a.pivot = pivot; a.count = count; a.sum = sum; a.gram = gram; a.C = C; a.n = n;
    const int segs = C <= 32 ? 4 : C <= 64 ? 2 : 1;
    a.n_stages = (int)((n + segs * gtc::KT - 1) / (segs * gtc::KT));
    a.raw_tx_bytes = C * segs * gtc::KT * 4;             // the box has C rows (rows C .. CP-1 of the slot stay untouched)
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)C};
    const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
    const cuuint32_t box[2] = {(cuuint32_t)(segs * gtc::KT), (cuuint32_t)C};
    if (make_tensor_map_f32(&tm, 2, feat, dims, strides, box)) return 2;
    if (segs == 4) return launch_gram_tc_cfg<4, gtc::SRC_NCHW>(a, tm, tm, st);
    if (segs == 2) return launch_gram_tc_cfg<2, gtc::SRC_NCHW>(a, tm, tm, st);
    return launch_gram_tc_cfg<1, gtc::SRC_NCHW>(a, tm, tm, st);
}

// per-label statistics of feat [C][H*W] with a uint8 label map [H*W] (labels >= L are ignored) into zero-initialised
// fp64 blocks count [L] | sum [L][C] | gram [L][C][C]
// C = 32 (the photorealistic latent, the mode masks are used in): wider latents would fold every pass with global
// atomics (their accumulator rows do not fit the drain warps' registers) and stay on the CUDA-core kernel
bool gram_tc_masked_eligible(int C, int H, int W) {
    return C == 32 && W % 4 == 0 && (long long)H * W >= 16384 && (long long)H * W < (1ll << 31);
}
template <int SEGS>
static int launch_gram_tc_masked_cfg(GramTcMaskedArgs a, const CUtensorMap& tm, cudaStream_t st) {
    static PerDeviceOnce smem_once;
    auto kern = gram_tc_masked_kernel<SEGS>;
    // 2 operand + 7 RAW slots; barriers (2 KB); 3 fp64 label slots; fp32 staging of the 4 segments
    constexpr size_t SMEM_MASKED = 2 * (size_t)gtc::OP_BYTES + 7 * (size_t)gtc::RAW_BYTES + 1024 + 256 + 2048 + 3 * 1024 * 8 + 4 * 1024 * 4;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)SMEM_MASKED));
    a.n_xb = cdiv(a.W, SEGS * gtc::KT);
    // Traversal: column groups of `gw` strips, each walked row by row; default = whole rows (row-major).  Measured on a
    // 1080p latent with a 2 x 4 grid of regions: strip by strip (gw = 1) 0.56 ms — a stage then jumps a whole image row and
    // its boxes take ~8 us to arrive; gw = 4: 0.36 ms; row-major: 0.31 ms, the drain's three label slots absorbing most of
    // the label alternation along a row (the CUDA-core kernel: 0.50 ms).  VST_GRAM_MASKED_GW overrides.
    { static int gw = -1; if (gw < 0) { const char* e = getenv("VST_GRAM_MASKED_GW"); gw = e ? std::max(1, atoi(e)) : (1 << 20); } a.gw = std::min(gw, a.n_xb); }
    a.n_stages = cdiv(a.n_xb, a.gw) * a.gw * a.H;
    int grid = std::min(a.n_stages, num_sms());
    a.stages_per_cta = cdiv(a.n_stages, grid);
    grid = cdiv(a.n_stages, a.stages_per_cta);
    kern<<<grid, gtc::THREADS, SMEM_MASKED, st>>>(a, tm);
#ifdef GTC_TRACE
    {
        cudaDeviceSynchronize();
        long long h[8 * 40];
        cudaMemcpyFromSymbol(h, gtc_trace, sizeof(h));
        long long t0 = h[1 * 40];
        const char* names[8] = {"conv arrive", "conv stage begin", "conv raw ok", "mma op ok", "mma acc ok", "t255 arrive", "t0 pre-barrier", "t255 pre-barrier"};
        for (int r = 0; r < 8; ++r) {
            printf("%-18s", names[r]);
            for (int i = 0; i < 20; ++i) printf(" %7lld", h[r * 40 + i] ? h[r * 40 + i] - t0 : -1);
            printf("\n");
        }
    }
#endif
    return check_launch("cwct_gram_tc_masked");
}
int launch_gram_tc_masked(const float* feat, const uint8_t* labels, const float* pivot, double* count, double* sum, double* gram,
                          int C, int L, int H, int W, cudaStream_t st) {
    VST_REQUIRE(gram_tc_masked_eligible(C, H, W) && (((uintptr_t)feat) & 15) == 0, "gram_tc_masked: C=%d %dx%d not eligible", C, H, W);
    GramTcMaskedArgs a = {};
    a.pivot = pivot; a.labels = labels; a.count = count; a.sum = sum; a.gram = gram; a.C = C; a.L = L; a.H = H; a.W = W;
    const int segs = 128 / C;
    const long long n = (long long)H * W;
    a.raw_tx_bytes = C * segs * gtc::KT * 4;
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)C};
    const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
    const cuuint32_t box[2] = {(cuuint32_t)(segs * gtc::KT), 16};        // half of the channel rows per copy
    if (make_tensor_map_f32(&tm, 2, feat, dims, strides, box)) return 2;
    return launch_gram_tc_masked_cfg<4>(a, tm, st);
}

// the same over the P4 half-states x1 | x2 ([Ch/4 groups][h+2][w+2][4] each) of a latent with C channels and
// 2*Ch / C sub-positions per state pixel
int launch_gram_tc_state(const float* x1, const float* x2, const float* pivot, double* count, double* sum, double* gram, int C,
                         int Ch, int h, int w, cudaStream_t st) {
    VST_REQUIRE(C == 32 || C == 64 || C == 128, "gram_tc_state: C = %d not supported", C);
    VST_REQUIRE((2 * Ch) % C == 0 && Ch % 128 == 0, "gram_tc_state: state width %d incompatible with C = %d", Ch, C);
    GramTcArgs a = {};
    a.pivot = pivot; a.count = count; a.sum = sum; a.gram = gram; a.C = C;
    const int nsub = 2 * Ch / C;                         // sub-positions per state pixel
    a.n = (long long)nsub * h * w;
    a.h = h; a.w = w;
    a.n_xb = cdiv(w, gtc::KT);
    a.n_jb = (2 * Ch / 4) / 32;                          // blocks of 32 state groups (= 128 / C sub-positions)
    a.groups_per_half = Ch / 4;
    a.n_stages = h * a.n_xb * a.n_jb;
    a.raw_tx_bytes = gtc::RAW_BYTES;
    CUtensorMap tm[2];
    const cuuint64_t dims[3] = {(cuuint64_t)(w + 2) * 4, (cuuint64_t)(h + 2), (cuuint64_t)(Ch / 4)};
    const cuuint64_t strides[2] = {(cuuint64_t)(w + 2) * 16, (cuuint64_t)(h + 2) * (w + 2) * 16};
    const cuuint32_t box[3] = {(cuuint32_t)gtc::KT * 4, 1, 32};
    if (make_tensor_map_f32(&tm[0], 3, x1, dims, strides, box) || make_tensor_map_f32(&tm[1], 3, x2, dims, strides, box)) return 2;
    const int segs = 128 / C;
    if (segs == 4) return launch_gram_tc_cfg<4, gtc::SRC_STATE>(a, tm[0], tm[1], st);
    if (segs == 2) return launch_gram_tc_cfg<2, gtc::SRC_STATE>(a, tm[0], tm[1], st);
    return launch_gram_tc_cfg<1, gtc::SRC_STATE>(a, tm[0], tm[1], st);
}

}  // namespace vst
