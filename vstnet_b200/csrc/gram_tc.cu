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
    uint8_t* op_base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
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
                    const int jb = st % a.n_jb, r = st / a.n_jb;
                    const int xb = r % a.n_xb, y = r / a.n_xb;
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
                const int xb = (st / a.n_jb) % a.n_xb;
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
