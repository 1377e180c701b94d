// gram_tc.cu — cWCT statistics of an unmasked feature map on the tensor cores:
//   sum[c] = sum_p (x[c,p] - pivot[c])          gram[c,c'] = sum_p (x[c,p] - pivot[c]) (x[c',p] - pivot[c'])
// Replaces mean / centring / x @ x.T of models/cWCT.py:138-144, :153-157 (and :223-226, :241-244) for the
// unmasked paths; the per-label variant stays on the CUDA-core kernel of cwct.cu.
//
// The Gram is a GEMM with M = N = C and K = pixels, both operands the same K-major matrix (NCHW rows).
// A tcgen05 UMMA is M = 128 wide, so for C < 128 the M (and N) dimension is filled with SEGS = 128 / C
// pixel SEGMENTS: operand row (s, c) holds channel c over the s-th quarter of the staged pixel run, and the
// diagonal C x C blocks of the 128 x 128 accumulator are the partial Grams of the segments (the off-diagonal
// blocks mix segments and are ignored: 1/SEGS of the tensor work is useful, which is still enough to keep the
// pass HBM-bound, profiles/).  C = 128 (artistic latent) uses the full tile.
//
//   * loaders (8 warps): 16-byte coalesced loads of 32 pixels per row and stage, subtract the pivot (fused
//     mean subtraction: the pivot is a cheap estimate of the mean, the exact mean correction is applied in
//     the factor kernel from `sum`), split x = hi + lo (tf32 + remainder) and store both into the
//     no-swizzle K-major canonical layout [k-chunk][row][4 floats] (row pitch padded: conflict-free stores).
//   * UMMA issuer (1 thread): per stage 4 k-steps x 3 terms (hi.hi + hi.lo + lo.hi: fp32-equivalent products),
//     kind::tf32, fp32 accumulators in TMEM, two accumulator buffers.
//   * drain warps (4): every FL stages (split-K) pull the finished buffer out of TMEM and fold it into fp64
//     (registers for C <= 64, global atomics for C = 128); at the end one fp64 atomicAdd per entry and CTA.
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vst {

namespace gtc {
constexpr int KT = 32;                       // pixels per segment and stage
constexpr int ROWP = 129;                    // padded rows per k-chunk (128 + 1): conflict-free 16-byte stores
constexpr int TILE_BYTES = (KT / 4) * ROWP * 16;         // one term (hi or lo) of one stage
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int NS = 4;
constexpr size_t SMEM = (size_t)NS * STAGE_BYTES + 1024 + 128;
constexpr int THREADS = 416;                 // 8 loader warps, 4 drain warps, UMMA issuer
}  // namespace gtc

struct GramTcArgs {
    const float* feat;       // [C][n]
    const float* pivot;      // [C]
    double* count;           // [1]
    double* sum;             // [C]
    double* gram;            // [C*C]
    int C;
    long long n;
    long long px_per_cta;    // multiple of SEGS * KT
};

template <int SEGS>
__global__ void __launch_bounds__(gtc::THREADS, 1) gram_tc_kernel(GramTcArgs a) {
    using namespace gtc;
    constexpr int CP = 128 / SEGS;               // padded channels per segment
    constexpr int RUN = SEGS * KT;               // pixels per stage
    constexpr int FL = (CP == 128) ? 32 : 8;     // stages per accumulator flush (split-K granularity)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stage_base = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* bars = (uint64_t*)(stage_base + (size_t)NS * STAGE_BYTES);
    uint64_t* full = bars;                 // [NS] 256 loader threads
    uint64_t* empty = bars + NS;           // [NS] tcgen05.commit
    uint64_t* acc_full = bars + 2 * NS;    // [2]  tcgen05.commit
    uint64_t* acc_empty = bars + 2 * NS + 2;   // [2]  128 drain threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NS + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long p_begin = (long long)blockIdx.x * a.px_per_cta;
    const long long p_end = p_begin + a.px_per_cta < a.n ? p_begin + a.px_per_cta : a.n;
    const int n_stages = p_begin < p_end ? (int)((p_end - p_begin + RUN - 1) / RUN) : 0;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 256); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        fence_barrier_init();
    }
    if (warp == 12) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ================= loaders =================
        // one warp-load covers 32 consecutive 16-byte units: (4 / SEGS) channels x (SEGS * 8) units per channel
        constexpr int UPC = SEGS * (KT / 4);            // 16-byte units per channel and stage
        constexpr int CPW = 32 / UPC;                   // channels per warp-load
        const int u = lane % UPC, dc = lane / UPC;      // unit within the channel's run, channel within the load
        const int seg = u / (KT / 4), j = u % (KT / 4); // segment, k-chunk
        float ssum[CP / (8 * CPW)];
#pragma unroll
        for (int i = 0; i < CP / (8 * CPW); ++i) ssum[i] = 0.f;
        for (int st = 0; st < n_stages; ++st) {
            const int s = st % NS;
            mbar_wait(&empty[s], ((st / NS) & 1) ^ 1);
            float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * STAGE_BYTES + TILE_BYTES);
            const long long p0 = p_begin + (long long)st * RUN + 4 * u;
#pragma unroll
            for (int i = 0; i < CP / (8 * CPW); ++i) {
                const int c = (i * 8 + warp) * CPW + dc;          // channel handled by this lane in this pass
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < a.C && p0 < p_end) {                      // n % 4 == 0: a unit is entirely inside or outside
                    v = __ldg(reinterpret_cast<const float4*>(a.feat + (size_t)c * a.n + p0));
                    const float pv = __ldg(a.pivot + c);
                    v.x -= pv; v.y -= pv; v.z -= pv; v.w -= pv;
                }
                ssum[i] += (v.x + v.y) + (v.z + v.w);
                const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                const int dst = j * ROWP + seg * CP + c;
                hi[dst] = h;
                lo[dst] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
        }
        // per-channel sums: lanes with the same dc hold the same channel
#pragma unroll
        for (int i = 0; i < CP / (8 * CPW); ++i) {
            float v = ssum[i];
#pragma unroll
            for (int o = UPC / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int c = (i * 8 + warp) * CPW + dc;
            if (u == 0 && c < a.C && n_stages > 0) atomicAdd(a.sum + c, (double)v);
        }
        if (tid == 0 && p_end > p_begin) atomicAdd(a.count, (double)(p_end - p_begin));
    } else if (warp == 12) {
        // ================= UMMA issuer =================
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t LBO = ROWP * 16, SBO = 128;
            for (int st = 0; st < n_stages; ++st) {
                const int s = st % NS, grp = st / FL;
                const uint32_t b = grp & 1;
                if (st % FL == 0) {
                    mbar_wait(&acc_empty[b], ((grp >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                mbar_wait(&full[s], (st / NS) & 1);
                tc_fence_after();
                const uint32_t Hi = smem_u32(stage_base + (size_t)s * STAGE_BYTES), Lo = Hi + TILE_BYTES;
                const uint32_t d = tmem_base + b * 128;
#pragma unroll
                for (int ks = 0; ks < KT / 8; ++ks) {
                    const uint64_t dh = make_desc(Hi + ks * 2 * LBO, LBO, SBO), dl = make_desc(Lo + ks * 2 * LBO, LBO, SBO);
                    umma_tf32(d, dh, dh, IDESC, (st % FL != 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(d, dh, dl, IDESC, 1u);
                    umma_tf32(d, dl, dh, IDESC, 1u);
                }
                umma_commit(&empty[s]);
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
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

template <int SEGS>
static int launch_gram_tc_cfg(const GramTcArgs& a0, cudaStream_t st) {
    static PerDeviceOnce smem_once;
    auto kern = gram_tc_kernel<SEGS>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)gtc::SMEM));
    GramTcArgs a = a0;
    const long long run = (long long)SEGS * gtc::KT;
    const long long runs = (a.n + run - 1) / run;
    int grid = (int)std::min<long long>(runs, num_sms());
    a.px_per_cta = (runs + grid - 1) / grid * run;
    grid = (int)((a.n + a.px_per_cta - 1) / a.px_per_cta);
    kern<<<grid, gtc::THREADS, gtc::SMEM, st>>>(a);
    return check_launch("cwct_gram_tc");
}

// small maps stay on the CUDA-core kernels: exact fp32 products (a two-term tf32 split carries 22 bits), and a
// persistent tensor-core pipeline has nothing to amortise there
bool gram_tc_eligible(int C, long long n) { return C >= 1 && C <= 128 && n % 4 == 0 && n >= 16384; }

// sums and Gram of the pivot-shifted features into zero-initialised fp64 buffers (one label)
int launch_gram_tc(const float* feat, const float* pivot, double* count, double* sum, double* gram, int C, long long n,
                   cudaStream_t st) {
    VST_REQUIRE(gram_tc_eligible(C, n) && (((uintptr_t)feat) & 15) == 0, "gram_tc: C=%d n=%lld not eligible", C, n);
    GramTcArgs a;
    a.feat = feat; a.pivot = pivot; a.count = count; a.sum = sum; a.gram = gram; a.C = C; a.n = n; a.px_per_cta = 0;
    if (C <= 32) return launch_gram_tc_cfg<4>(a, st);
    if (C <= 64) return launch_gram_tc_cfg<2>(a, st);
    return launch_gram_tc_cfg<1>(a, st);
}

}  // namespace vst
