// conv_tcx.cu — "kx-folded" tcgen05 convolution for the convs without a coupling operand
// (conv1 / conv2 of every reversible block: refpad + conv3x3 + ReLU, models/RevResNet.py:80-85).
//
// Why a second formulation: measured on B200, one kind::tf32 UMMA (M=128, K=8) from no-swizzle shared
// memory costs ~66 cycles for ANY N <= 128 and ~0.43 cycles per column above that (profiles/):
// the instruction is bound by fetching its 4 KB A operand, so the tensor pipe is only efficient when
// every A fetch feeds many output columns.  The tap-shifted implicit GEMM of conv_tc.cu fetches
// each activation tile 9x (once per tap) for N = Cout <= 64 columns.  Here the three kx taps are
// folded into the N dimension instead:
//     D[pixel m, (kx, cout)] += A[pixel m, (ky-shifted row, cin)] * W[ky][cin, (kx, cout)]
//     out[x, cout] = D[x-1, (0,cout)] + D[x, (1,cout)] + D[x+1, (2,cout)]
// so A is fetched 3x (once per ky, a whole-row descriptor shift) for N = 3*Cout columns, and the kx
// shift becomes a lane shift in the epilogue (warp shuffles; the two lanes per warp whose neighbour
// lives in another warp go through a 4 KB shared-memory exchange).  Tensor work per output drops 3x.
// A tile spans 128 input pixels and produces 126 outputs (x-stride 126).
//
// Everything else follows conv_tc.cu: P4 activations, one cp.async.bulk per (group, row) segment,
// converter warps for the hi/lo tf32 split, mbarrier stage ring, persistent CTAs, accumulators in
// TMEM (double-buffered when 2 x R x 3Cout columns fit), branch-free ReLU epilogue with inline
// reflection-border stores.
#include <stdlib.h>
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vst {

template <int NC, int R, int TERMS>
struct TcxCfg {
    static constexpr int NP = 3 * NC;               // UMMA N: (kx, cout)
    static constexpr int TA = TERMS >= 2 ? 2 : 1;
    static constexpr int TW = TERMS >= 3 ? 2 : 1;
    static constexpr int PW = 128;                  // staged pixels per row = UMMA M
    static constexpr int XS = 126;                  // outputs per tile row
    static constexpr int ROWS = R + 2;
    static constexpr int ROW_BYTES = PW * 16;
    static constexpr int A_TERM_BYTES = 2 * ROWS * ROW_BYTES;
    static constexpr int A_BYTES = TA * A_TERM_BYTES;
    static constexpr int B_TERM_BYTES = 3 * 2 * NP * 16;
    static constexpr int B_BYTES = TW * B_TERM_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int ACC_COLS = R * NP;
    static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (NACC * ACC_COLS <= 128) ? 128 : (NACC * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int EXCH_FLOATS = 2 * R * 4 * 2 * NC;          // [tile parity][row][warp quarter][side][cout]
    static constexpr int AUX_BYTES = 2048 + EXCH_FLOATS * 4;        // barriers (1 KB) + bias (1 KB) + exchange
    static constexpr int NS_FIT = (226 * 1024 - AUX_BYTES) / STAGE_BYTES;
    static constexpr int NS = NS_FIT > 4 ? 4 : NS_FIT;
    static constexpr size_t SMEM = (size_t)NS * STAGE_BYTES + AUX_BYTES + 128;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(NP % 16 == 0 && NP <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
    static_assert(NS >= 2, "need at least a double-buffered operand pipeline");
};

// raw OIHW -> [cout tile][chunk][term][ky][cin/4 (2)][n' = kx*NC + co][4]
__global__ void pack_tcx_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int NC,
                                        int TW) {
    const int NP = 3 * NC;
    const size_t total = (size_t)(Cout / NC) * (Cin / 8) * TW * 3 * 2 * NP * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = (int)(r % 4); r /= 4;
        const int n = (int)(r % NP); r /= NP;
        const int g = (int)(r % 2); r /= 2;
        const int ky = (int)(r % 3); r /= 3;
        const int term = (int)(r % TW); r /= TW;
        const int chunk = (int)(r % (Cin / 8)); r /= (Cin / 8);
        const int tile = (int)r;
        const int kx = n / NC, co = tile * NC + (n - kx * NC), ci = chunk * 8 + g * 4 + e;
        const float v = w[((size_t)co * Cin + ci) * 9 + ky * 3 + kx];
        const float hi = tf32_round(v);
        wp[i] = term == 0 ? hi : tf32_round(v - hi);
    }
}

int launch_pack_tcx_weights(const float* w, float* wp, int Cin, int Cout, int NC, int terms, cudaStream_t st) {
    const int TW = terms >= 3 ? 2 : 1;
    const size_t total = (size_t)(Cout / NC) * (Cin / 8) * TW * 3 * 2 * (3 * NC) * 4;
    pack_tcx_weights_kernel<<<(int)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(w, wp, Cin, Cout, NC, TW);
    return check_launch("pack_tcx_weights");
}

struct TcxTiles {
    int n_xt, n_yt, n_ct, n_tiles;
    long long* trace;   // developer aid (VST_TC_TRACE): CTA 0 stamps clock64() per role event, 4096 slots per role
};
#define TCX_TRACE(role, idx) do { if (tl.trace && blockIdx.x == 0 && (idx) < 4096) tl.trace[(role) * 4096 + (idx)] = clock64(); } while (0)

constexpr int TCX_THREADS = 448;   // 8 epilogue + 4 converter warps, operand producer, UMMA issuer

template <int NC, int R, int TERMS>
__global__ void __launch_bounds__(TCX_THREADS, 1) conv3x3_tcx_kernel(ConvArgs a, TcxTiles tl) {
    using Cfg = TcxCfg<NC, R, TERMS>;
    constexpr int NS = Cfg::NS, PW = Cfg::PW, ROWS = Cfg::ROWS, NP = Cfg::NP, NACC = Cfg::NACC, XS = Cfg::XS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stage_base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // by offset: keeps the __shared__ address space (LDS / STS)
    uint64_t* bars = (uint64_t*)(stage_base + (size_t)NS * Cfg::STAGE_BYTES);
    uint64_t* loaded = bars;                    // [NS]    operand producer arrive.expect_tx + TMA bytes
    uint64_t* ready = bars + NS;                // [NS]    128 converter threads
    uint64_t* empty = bars + 2 * NS;            // [NS]    tcgen05.commit
    uint64_t* acc_full = bars + 3 * NS;         // [NACC]  tcgen05.commit
    uint64_t* acc_empty = bars + 3 * NS + 2;    // [NACC]  256 epilogue threads
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * NS + 4);
    float* bias_s = (float*)((uint8_t*)bars + 1024);
    float* exch = bias_s + 256;                 // [2][R][4][2][NC]
    for (int i = threadIdx.x; i < a.Cout && i < 256; i += blockDim.x) bias_s[i] = __ldg(a.bias + i);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = a.Cin / 8;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&loaded[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
        for (int b = 0; b < NACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 12) {
        // ================= operand producer (TMA) =================
        if (lane == 0) {
            const int Hp = a.Hin + 2, Wp = a.Win + 2;
            const float4* in4 = reinterpret_cast<const float4*>(a.in);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
                const int ct = t % tl.n_ct, rest = t / tl.n_ct;
                const int xs = (rest % tl.n_xt) * XS, y0 = (rest / tl.n_xt) * R;
                const float* wsrc = a.w + (size_t)ct * n_chunks * (Cfg::B_BYTES / 4);
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int s = it % NS;
                    mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
                    TCX_TRACE(0, it);
                    uint8_t* A = stage_base + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&loaded[s], Cfg::A_TERM_BYTES + Cfg::B_BYTES);
#pragma unroll
                    for (int g = 0; g < 2; ++g)
#pragma unroll
                        for (int row = 0; row < ROWS; ++row) {
                            const int py = min(y0 + row, Hp - 1);     // padded row of image row y0-1+row
                            // tile pixel p <-> padded column xs + p (image x = xs - 1 + p)
                            bulk_g2s(A + (g * ROWS + row) * Cfg::ROW_BYTES,
                                     in4 + ((size_t)(2 * c + g) * Hp + py) * Wp + xs, Cfg::ROW_BYTES, &loaded[s]);
                        }
                    bulk_g2s(A + Cfg::A_BYTES, wsrc + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, &loaded[s]);
                }
            }
        }
    } else if (warp == 13) {
        // ================= UMMA issuer =================
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t A_LBO = ROWS * Cfg::ROW_BYTES, B_LBO = NP * 16, SBO = 128;
            uint32_t it = 0, tcount = 0;
            for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
                const uint32_t b = tcount % NACC;
                mbar_wait(&acc_empty[b], ((tcount / NACC) & 1) ^ 1);
                tc_fence_after();
                TCX_TRACE(5, tcount);
                const uint32_t acc = tmem_base + b * Cfg::ACC_COLS;
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    const int s = it % NS;
                    mbar_wait(&ready[s], (it / NS) & 1);
                    tc_fence_after();
                    TCX_TRACE(3, it);
                    const uint32_t Aaddr = smem_u32(stage_base + (size_t)s * Cfg::STAGE_BYTES);
                    const uint32_t Baddr = Aaddr + Cfg::A_BYTES;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint64_t bh = make_desc(Baddr + ky * 2 * NP * 16, B_LBO, SBO);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint32_t aoff = (r + ky) * Cfg::ROW_BYTES;
                            const uint32_t d = acc + r * NP;
                            const uint32_t first = (c > 0 || ky > 0) ? 1u : 0u;
                            umma_tf32(d, make_desc(Aaddr + aoff, A_LBO, SBO), bh, IDESC, first);
                            if (TERMS >= 2)
                                umma_tf32(d, make_desc(Aaddr + Cfg::A_TERM_BYTES + aoff, A_LBO, SBO), bh, IDESC, 1u);
                            if (TERMS >= 3)
                                umma_tf32(d, make_desc(Aaddr + aoff, A_LBO, SBO),
                                          make_desc(Baddr + Cfg::B_TERM_BYTES + ky * 2 * NP * 16, B_LBO, SBO), IDESC, 1u);
                        }
                    }
                    umma_commit(&empty[s]);
                    TCX_TRACE(4, it);
                }
                umma_commit(&acc_full[b]);
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ================= converters: hi/lo split of the staged activations =================
        const int ctid = tid - 256;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c, ++it) {
                const int s = it % NS;
                mbar_wait(&loaded[s], (it / NS) & 1);
                if (ctid == 0) TCX_TRACE(1, it);
                float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * Cfg::STAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * Cfg::STAGE_BYTES + Cfg::A_TERM_BYTES);
#pragma unroll 4
                for (int i = ctid; i < 2 * ROWS * PW; i += 128) {
                    const float4 v = hi[i];
                    const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                    hi[i] = h;
                    if (Cfg::TA == 2) lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
                fence_proxy_async();
                mbar_arrive(&ready[s]);
                if (ctid == 0) TCX_TRACE(2, it);
            }
        }
    } else {
        // ================= epilogue: TMEM -> (lane-shifted sum over kx) -> ReLU -> P4 global =================
        // lane l of warp (q, half) owns tile pixel m = 32q + l and the cout half `half`.
        constexpr int HC = NC / 2;                         // couts per thread
        constexpr int CH = HC >= 16 ? 16 : HC;             // couts per TMEM load
        const int q = warp & 3, half = warp >> 2;
        const int m = q * 32 + lane;
        const float flo = (a.epi == EPI_RELU) ? 0.f : -INFINITY;
        const int H = a.Hout, W = a.Wout, Wp = W + 2;
        const size_t plane = p4_plane_px(H, W);
        uint32_t tcount = 0;
        for (int t = blockIdx.x; t < tl.n_tiles; t += gridDim.x, ++tcount) {
            const int ct = t % tl.n_ct, rest = t / tl.n_ct;
            const int xs = (rest % tl.n_xt) * XS, y0 = (rest / tl.n_xt) * R;
            const uint32_t b = tcount % NACC;
            const int x = xs - 1 + m;
            const bool xin = (m >= 1) && (m <= XS) && (x < W);
            const int rows = min(R, H - y0);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + b * Cfg::ACC_COLS + half * HC;
            float* ex = exch + (size_t)(tcount & 1) * (R * 4 * 2 * NC);
            mbar_wait(&acc_full[b], (tcount / NACC) & 1);
            tc_fence_after();
            if (tid == 0) TCX_TRACE(6, tcount);
            // ---- phase 1: publish the partial sums a neighbouring warp needs (lane 31's kx=0, lane 0's kx=2)
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
#pragma unroll 1
                for (int c0 = 0; c0 < HC; c0 += CH) {
                    float v0[CH], v2[CH];
                    tmem_ld2<CH>(trow + (uint32_t)(r * NP + 0 * NC + c0), v0, trow + (uint32_t)(r * NP + 2 * NC + c0), v2);
                    if (lane == 31) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) ex[((r * 4 + q) * 2 + 0) * NC + half * HC + c0 + i] = v0[i];
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) ex[((r * 4 + q) * 2 + 1) * NC + half * HC + c0 + i] = v2[i];
                    }
                }
            }
            named_barrier(1, 256);
            if (tid == 0) TCX_TRACE(7, 2 * tcount);
            // ---- phase 2: out[m] = D[m-1][kx=0] + D[m][kx=1] + D[m+1][kx=2]
            const bool lf = xin && (x == 1), rt = xin && (x == W - 2);
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int y = y0 + r;
                const bool up = (y == 1), dn = (y == H - 2);
#pragma unroll 1
                for (int c0 = 0; c0 < HC; c0 += CH) {
                    float v0[CH], v1[CH], v2[CH];
                    tmem_ld3<CH>(trow + (uint32_t)(r * NP + 0 * NC + c0), v0, trow + (uint32_t)(r * NP + 1 * NC + c0), v1,
                                 trow + (uint32_t)(r * NP + 2 * NC + c0), v2);
                    if (tid == 0 && tcount == 1) TCX_TRACE(7, 16 + (r * (HC / CH) + c0 / CH) * 3 + 0);
                    const int cb = half * HC + c0;                  // first cout of this chunk (within the tile)
                    const float* exl = ex + ((r * 4 + (q > 0 ? q - 1 : 0)) * 2 + 0) * NC + cb;   // left neighbour warp, lane 31
                    const float* exr = ex + ((r * 4 + (q < 3 ? q + 1 : 3)) * 2 + 1) * NC + cb;   // right neighbour warp, lane 0
                    // neighbour-warp values: warp-uniform addresses (broadcast loads), selected without branching
                    float el[CH], er[CH];
#pragma unroll
                    for (int i = 0; i < CH; i += 4) {
                        const float4 a4 = *reinterpret_cast<const float4*>(exl + i);
                        const float4 b4 = *reinterpret_cast<const float4*>(exr + i);
                        el[i] = a4.x; el[i + 1] = a4.y; el[i + 2] = a4.z; el[i + 3] = a4.w;
                        er[i] = b4.x; er[i + 1] = b4.y; er[i + 2] = b4.z; er[i + 3] = b4.w;
                    }
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const float ls = __shfl_up_sync(0xffffffffu, v0[i], 1);
                        const float rs = __shfl_down_sync(0xffffffffu, v2[i], 1);
                        const float l = (lane == 0) ? el[i] : ls;
                        const float rr = (lane == 31) ? er[i] : rs;
                        v1[i] = (l + v1[i]) + rr;
                    }
                    if (tid == 0 && tcount == 1) TCX_TRACE(7, 16 + (r * (HC / CH) + c0 / CH) * 3 + 1);
                    if (xin) {
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
                    if (tid == 0 && tcount == 1) TCX_TRACE(7, 16 + (r * (HC / CH) + c0 / CH) * 3 + 2);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[b]);
            if (tid == 0) TCX_TRACE(7, 2 * tcount + 1);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int NC, int R, int TERMS>
static int launch_tcx_cfg(const ConvArgs& a, cudaStream_t st) {
    using Cfg = TcxCfg<NC, R, TERMS>;
    static PerDeviceOnce smem_once;
    auto kern = conv3x3_tcx_kernel<NC, R, TERMS>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    TcxTiles tl;
    tl.n_xt = cdiv(a.Wout, Cfg::XS); tl.n_yt = cdiv(a.Hout, R); tl.n_ct = a.Cout / NC;
    tl.n_tiles = tl.n_xt * tl.n_yt * tl.n_ct;
    tl.trace = tc_trace_buffer(a.Cin, a.Cout, st);
    const int grid = std::min(tl.n_tiles, num_sms());
    char cls[40];
    snprintf(cls, sizeof(cls), "conv3x3_tcx%d %d>%d", TERMS, a.Cin, a.Cout);
    const double px = (double)a.Hout * a.Wout;
    ProfScope prof(st, cls, 2.0 * 9 * a.Cin * a.Cout * px, 4.0 * ((double)a.Cin * a.Hin * a.Win + a.Cout * px));
    kern<<<grid, TCX_THREADS, Cfg::SMEM, st>>>(a, tl);
    return check_launch("conv3x3_tcx");
}

bool tcx_eligible(int Cin, int Cout, int stride) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("VST_TCX_OFF"); off = e ? atoi(e) : 0; }
    return !off && stride == 1 && Cin % 8 == 0 && Cin >= 16 && (Cout == 64 || Cout == 16);
}

// a.w must point at weights packed by launch_pack_tcx_weights with NC = Cout and `terms`; epi is RELU or NONE
int launch_conv3x3_tcx(const ConvArgs& a, int terms, cudaStream_t st) {
    VST_REQUIRE(tcx_eligible(a.Cin, a.Cout, 1), "conv3x3_tcx: shape %d>%d not eligible", a.Cin, a.Cout);
    VST_REQUIRE(a.epi == EPI_RELU || a.epi == EPI_NONE, "conv3x3_tcx has no coupling epilogue");
    VST_REQUIRE(a.Hin == a.Hout && a.Win == a.Wout && a.Hin >= 2 && a.Win >= 2, "conv3x3_tcx is stride 1, H,W >= 2");
    if (a.Cout == 64) {
        if (terms == 1) return launch_tcx_cfg<64, 2, 1>(a, st);
        if (terms == 2) return launch_tcx_cfg<64, 2, 2>(a, st);
        return launch_tcx_cfg<64, 2, 3>(a, st);
    }
    if (terms == 1) return launch_tcx_cfg<16, 4, 1>(a, st);
    if (terms == 2) return launch_tcx_cfg<16, 4, 2>(a, st);
    return launch_tcx_cfg<16, 4, 3>(a, st);
}

}  // namespace vst
