// conv_pair.cu — the 64 -> 256 coupling convolution (conv3 of a stage-3 / channel_reduction block,
// models/RevResNet.py:86-88,101-116) on a CTA PAIR: tcgen05.mma.cta_group::2, M = 256 x N = 256 x K = 16.
//
// Why a pair.  conv_tc.cu's single-CTA version issues M = 128 x N = 128 UMMAs: every instruction fetches 4 KB of A
// and 4 KB of B from shared memory for 64 cycles of math — all of the SM's 128 B/clk shared-memory datapath, which the
// TMA fills of the operand stages and the epilogue's global loads / stores (they share that datapath through L1) also
// need; the role trace shows 72-140 cycles per UMMA.  With cta_group::2 the two SMs of a TPC execute one UMMA on
// 2 x 128 pixels x 256 couts: each SM fetches its own 128 rows of A and only ITS HALF of B (the 128 couts whose
// weights it keeps resident), i.e. 8 KB per 128 cycles of math = half the operand bandwidth per flop.
//
// Geometry.  A CTA owns one image row x 128 pixels of all 256 couts per tile (accumulator 256 TMEM columns, double
// buffered = the whole TMEM); the pair walks tiles 2p, 2p+1 in lockstep.  Operands as in conv_tc.cu (HALF, WST):
// the input is an H8 split-half tensor (hi / lo fp16 rows already in the K-major operand layout), the halo tile of a
// 16-channel chunk and term is 2 x 3 rows x 130 pixels x 16 B = 12.5 KB = one pipeline stage (6 stages), the fp16
// weights of the CTA's cout half (4 chunks x 36 KB) are fetched once and stay resident.
//
// Synchronisation across the pair (rank 0 = leader issues every UMMA):
//   loaded[s]       local TMA bytes of stage s                      (each CTA, its own producer)
//   peer_loaded[s]  rank 1's relay warp forwards its loaded[s] to the leader with a remote mbarrier arrive
//   empty[s]        tcgen05.commit multicast to both CTAs: stage s may be refilled
//   acc_full[b]     tcgen05.commit multicast: accumulator b complete in both CTAs' TMEM
//   acc_empty[b]    on the leader: one arrive per epilogue warp of BOTH CTAs (8 local + 8 remote)
#include <stdlib.h>
#include "kernels.cuh"
#include <cuda_fp16.h>
#include "tc_ptx.cuh"

namespace vst {

struct PairCfg {
    static constexpr int KCH = 16, PW = 130, ROWS = 3;
    static constexpr int ROW_BYTES = PW * 16;
    static constexpr int STAGE_BYTES = 2 * ROWS * ROW_BYTES;       // one term of one chunk: [k-half][row][pixel][8 halfs]
    static constexpr int NLOC = 128, N = 256;                      // couts resident per CTA / per UMMA
    static constexpr int B_BYTES = 9 * 2 * NLOC * 16;              // one chunk of the CTA's weights: [tap][k-half][n][8 halfs]
    static constexpr int CHUNKS = 4;                               // Cin = 64
    static constexpr int W_RES_BYTES = CHUNKS * B_BYTES;
    static constexpr int NS = 6;
    static constexpr int AUX_BYTES = 2048;                         // barriers (1 KB) + bias (256 floats)
    static constexpr size_t SMEM = (size_t)NS * STAGE_BYTES + W_RES_BYTES + AUX_BYTES + 128;
    static constexpr int THREADS = 512;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    static_assert(STAGE_BYTES % 16 == 0 && B_BYTES % 128 == 0, "operand blocks must stay aligned");
};

struct PairTiles {
    int n_xt, n_tiles, n_ptiles;
    long long* trace;
};
#define PAIR_TRACE1(role, idx) do { if (tl.trace && blockIdx.x == 1 && (idx) < 2048) tl.trace[(role) * 4096 + 2048 + (idx)] = clock64(); } while (0)   // peer CTA of pair 0
#define PAIR_TRACE(role, idx) do { if (tl.trace && blockIdx.x == 0 && (idx) < 4096) tl.trace[(role) * 4096 + (idx)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(rank)
                 : "memory");
}
// polling wait (mbarrier.test_wait): for barriers completed by a REMOTE arrive — a thread parked by try_wait's
// suspend-time hint was observed to wake up to several thousand cycles after a remote arrive had completed the phase
__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows per CTA] * B[128 couts per CTA], kind::f16, M = 256, K = 16 (leader only)
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once all previously issued pair UMMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ bool pair_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg::THREADS, 1) conv3x3_pair_kernel(ConvArgs a, PairTiles tl) {
    using Cfg = PairCfg;
    constexpr int NS = Cfg::NS, PW = Cfg::PW, ROWS = Cfg::ROWS, N = Cfg::N;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stage_base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // by offset: keeps the __shared__ address space
    uint8_t* wres = stage_base + (size_t)NS * Cfg::STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(wres + Cfg::W_RES_BYTES);
    uint64_t* loaded = bars;                     // [NS]
    uint64_t* empty = bars + NS;                 // [NS]
    uint64_t* peer_loaded = bars + 2 * NS;       // [NS]  (used on the leader)
    uint64_t* acc_full = bars + 3 * NS;          // [2]
    uint64_t* acc_empty = bars + 3 * NS + 2;     // [2]   (used on the leader)
    uint64_t* wbar = bars + 3 * NS + 4;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * NS + 5);
    float* bias_s = (float*)((uint8_t*)bars + 1024);
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bias_s[i] = __ldg(a.bias + i);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&loaded[s], 1); mbar_init(&empty[s], 1); mbar_init(&peer_loaded[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 16); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 12 && lane == 0) {
            // ================= operand producer (TMA), both CTAs =================
            const int Hp = a.Hin + 2, Wp = a.Win + 2;
            const float4* in4 = reinterpret_cast<const float4*>(a.in);
            {   // this CTA's cout half: weights packed per 128-cout tile (launch_pack_tc_half_weights, N = 128)
                const float* wsrc = a.w + (size_t)rank * Cfg::CHUNKS * (Cfg::B_BYTES / 4);
                mbar_arrive_expect_tx(wbar, (uint32_t)Cfg::W_RES_BYTES);
                for (int c = 0; c < Cfg::CHUNKS; ++c)
                    bulk_g2s(wres + (size_t)c * Cfg::B_BYTES, wsrc + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, wbar);
            }
            uint32_t it = 0;
            for (int pt = pair; pt < tl.n_ptiles; pt += npairs) {
                const int t = min(2 * pt + (int)rank, tl.n_tiles - 1);      // an odd tile count: the last tile is computed twice
                const int x0 = (t % tl.n_xt) * 128, y0 = t / tl.n_xt;
                for (int c = 0; c < Cfg::CHUNKS; ++c)
#pragma unroll
                    for (int term = 0; term < 2; ++term, ++it) {
                        const int s = it % NS;
                        mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
                        PAIR_TRACE(0, it);
                        PAIR_TRACE1(0, it);
                        uint8_t* A = stage_base + (size_t)s * Cfg::STAGE_BYTES;
                        mbar_arrive_expect_tx(&loaded[s], Cfg::STAGE_BYTES);
                        // H8: the lo planes follow the Cin/8 hi planes; both are addressed like P4 groups
                        const float4* src = in4 + (size_t)term * (a.Cin / 8) * Hp * Wp;
#pragma unroll
                        for (int g = 0; g < 2; ++g)
#pragma unroll
                            for (int row = 0; row < ROWS; ++row) {
                                const int py = min(y0 + row, Hp - 1);     // padded row of image row y0-1+row
                                bulk_g2s(A + (g * ROWS + row) * Cfg::ROW_BYTES, src + ((size_t)(2 * c + g) * Hp + py) * Wp + x0,
                                         Cfg::ROW_BYTES, &loaded[s]);
                            }
                    }
            }
        } else if (warp == 13 && rank == 0) {
            // ================= UMMA issuer (leader): warp-uniform schedule, one elected lane issues =================
            // instruction descriptor: D = F32 (bit 4), A = B = F16, N at bit 17, M = 256 at bit 24
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
            constexpr uint32_t A_LBO = ROWS * Cfg::ROW_BYTES, B_LBO = Cfg::NLOC * 16, SBO = 128;
            const bool el = pair_elect_one();
            const uint64_t dA0 = make_desc(smem_u32(stage_base), A_LBO, SBO);
            const uint64_t dB0 = make_desc(smem_u32(wres), B_LBO, SBO);
            uint32_t it = 0, tcount = 0;
            mbar_wait(wbar, 0);
            for (int pt = pair; pt < tl.n_ptiles; pt += npairs, ++tcount) {
                const uint32_t b = tcount & 1;
                mbar_wait_poll(&acc_empty[b], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) PAIR_TRACE(5, tcount);
                const uint32_t acc = tmem_base + b * N;
                for (int c = 0; c < Cfg::CHUNKS; ++c)
#pragma unroll
                    for (int term = 0; term < 2; ++term, ++it) {
                        const int s = it % NS;
                        if (lane == 0) PAIR_TRACE(1, it);
                        mbar_wait(&loaded[s], (it / NS) & 1);
                        if (lane == 0) PAIR_TRACE(2, it);
                        mbar_wait_poll(&peer_loaded[s], (it / NS) & 1);
                        tc_fence_after();
                        if (lane == 0) PAIR_TRACE(3, it);
                        const uint64_t dA = dA0 + (uint64_t)((uint32_t)s * (Cfg::STAGE_BYTES >> 4));
                        const uint64_t dB = dB0 + (uint64_t)((uint32_t)c * (Cfg::B_BYTES >> 4));
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int ky = tap / 3, kx = tap % 3;
                            if (el)
                                umma_f16_pair(acc, dA + (uint64_t)(((ky * PW + kx) * 16) >> 4), dB + (uint64_t)((tap * 2 * Cfg::NLOC * 16) >> 4),
                                              IDESC, (c > 0 || term > 0 || tap > 0) ? 1u : 0u);
                        }
                        if (el) umma_commit_pair(&empty[s]);
                        if (lane == 0) PAIR_TRACE(4, it);
                    }
                if (el) umma_commit_pair(&acc_full[b]);
            }
        } else if (warp == 13) {
            // ================= relay (rank 1): forwards "stage loaded" to the leader =================
            uint32_t it = 0;
            if (lane == 0) {
                mbar_wait(wbar, 0);          // the first forward also vouches for this CTA's resident weights
                for (int pt = pair; pt < tl.n_ptiles; pt += npairs)
                    for (int k = 0; k < 2 * Cfg::CHUNKS; ++k, ++it) {
                        const int s = it % NS;
                        mbar_wait(&loaded[s], (it / NS) & 1);
                        PAIR_TRACE1(1, it);
                        mbar_arrive_remote(&peer_loaded[s], 0);
                    }
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");      // idle warpgroup: hands its registers to the epilogue
    } else {
        // ================= epilogue (both CTAs): TMEM -> registers -> + coupling operand -> P4 global =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        // as conv_tc.cu's hot epilogue with one row and 128 couts per thread column half: the coupling operand of the
        // whole tile (32 x 16-byte loads per thread) is requested before the accumulator wait
        constexpr int NG = N / 8;                                         // cout groups per thread
        constexpr int CH = 16;                                            // TMEM columns per tcgen05.ld
        constexpr int NCHK = NG * 4 / CH;
        const int q = warp & 3, half = warp >> 2;
        const bool coupled = a.epi == EPI_ADD || a.epi == EPI_SUB;
        const float sgn = (a.epi == EPI_SUB) ? -1.f : 1.f;
        const float flo = (a.epi == EPI_RELU) ? 0.f : -INFINITY;
        constexpr float unscale = 1.0f / VST_HALF_SCALE;                  // H8 operands carry VST_HALF_SCALE * x
        const int H = a.Hout, W = a.Wout, Wp = W + 2;
        const size_t plane = p4_plane_px(H, W);
        const int g0 = half * NG;
        uint32_t tcount = 0;
        // The coupling operand of a tile lives in registers (32 x 16 bytes per thread).  It is requested one tile AHEAD and
        // piecewise: as soon as a 16-cout chunk of tile i has been combined and stored, its registers are re-targeted at the
        // same chunk of tile i+1.  The loads thus have a whole tile period to arrive, and they reach the memory system as
        // eight 16 KB pieces instead of one 128 KB burst that the operand stages' TMA fills would queue behind.
        auto tile_of = [&](int pt, bool& valid, int& x, int& y) {
            const int tt = 2 * pt + (int)rank;
            valid = pt < tl.n_ptiles && tt < tl.n_tiles;
            const int t = min(tt, tl.n_tiles - 1);
            x = (t % tl.n_xt) * 128 + q * 32 + lane;
            y = t / tl.n_xt;
        };
        float4 res[NG];
        {
            bool v0; int x, y;
            tile_of(pair, v0, x, y);
            const float4* resp = reinterpret_cast<const float4*>(a.res) + (size_t)g0 * plane + (size_t)(y + 1) * Wp + (x + 1);
#pragma unroll
            for (int j = 0; j < NG; ++j) res[j] = (coupled && v0 && x < W) ? resp[(size_t)j * plane] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int pt = pair; pt < tl.n_ptiles; pt += npairs, ++tcount) {
            bool valid; int x, y;
            tile_of(pt, valid, x, y);
            const uint32_t b = tcount & 1;
            const bool xin = valid && x < W;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + b * N + half * (N / 2);
            const bool lf = (x == 1), rt = (x == W - 2);      // this pixel also feeds border column -1 / W
            const bool up = (y == 1), dn = (y == H - 2);      // row also feeds border row -1 / H
            float4* outp = reinterpret_cast<float4*>(a.out) + (size_t)g0 * plane + (size_t)(y + 1) * Wp + (x + 1);
            bool nvalid; int xn, yn;
            tile_of(pt + npairs, nvalid, xn, yn);
            const bool nin = coupled && nvalid && xn < W;
            const float4* resn = reinterpret_cast<const float4*>(a.res) + (size_t)g0 * plane + (size_t)(yn + 1) * Wp + (xn + 1);
            if (coupled && (lane & 7) == 0) {
                // the coupling operand of the tile after the next: pulled HBM -> L2 now (one 128-byte line per 8 lanes)
                bool v2; int x2, y2;
                tile_of(pt + 2 * npairs, v2, x2, y2);
                if (v2 && x2 < W) {
                    const float4* rn = reinterpret_cast<const float4*>(a.res) + (size_t)g0 * plane + (size_t)(y2 + 1) * Wp + (x2 + 1);
#pragma unroll
                    for (int j = 0; j < NG; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(rn + (size_t)j * plane));
                }
            }
            mbar_wait(&acc_full[b], (tcount >> 1) & 1);
            tc_fence_after();
            if (tid == 0) PAIR_TRACE(6, tcount);
            uint32_t ub[2][CH];
            tmem_ld_nowait<CH>(trow, ub[0]);
#pragma unroll
            for (int k = 0; k < NCHK; ++k) {
                tmem_ld_wait();
                tmem_ld_fence_regs<CH>(ub[k & 1]);
                float v[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(ub[k & 1][i]);
                if (k + 1 < NCHK) tmem_ld_nowait<CH>(trow + (uint32_t)((k + 1) * CH), ub[(k + 1) & 1]);
                if (k + 1 == NCHK) {
                    // every TMEM read of this accumulator has completed: hand it back to the leader's issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(&acc_empty[b]);
                        else mbar_arrive_remote(&acc_empty[b], 0);
                    }
                }
                if (xin) {
#pragma unroll
                    for (int jj = 0; jj < CH / 4; ++jj) {
                        const int j = k * (CH / 4) + jj;
                        const float4 bv = *reinterpret_cast<const float4*>(bias_s + 4 * (g0 + j));
                        float4 o = res[j];
                        o.x += fmaxf((v[4 * jj] * unscale + bv.x) * sgn, flo);
                        o.y += fmaxf((v[4 * jj + 1] * unscale + bv.y) * sgn, flo);
                        o.z += fmaxf((v[4 * jj + 2] * unscale + bv.z) * sgn, flo);
                        o.w += fmaxf((v[4 * jj + 3] * unscale + bv.w) * sgn, flo);
                        float4* p = outp + (size_t)j * plane;
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
#pragma unroll
                for (int jj = 0; jj < CH / 4; ++jj) {          // this chunk's registers now fetch the next tile's operand
                    const int j = k * (CH / 4) + jj;
                    res[j] = nin ? resn[(size_t)j * plane] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (tid == 0) PAIR_TRACE(7, tcount);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // no CTA leaves (or frees TMEM) while its peer can still signal it
    if (warp == 13) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

long long* tc_trace_buffer(int Cin, int Cout, cudaStream_t st);

// Off by default (VST_TC_PAIR=1 enables it): measured 0.091 ms per launch against 0.084 ms for conv_tc.cu's single-CTA
// weight-stationary kernel at 1080p.  Knock-out runs show why the halved operand bandwidth does not pay here: with the
// epilogue's coupling loads and stores switched off the single-CTA kernel drops to 0.069 ms, with half of the operand
// fills switched off it stays at 0.080 ms — the conv is bound by its epilogue's global traffic and by per-launch fixed
// costs (147 KB of resident weights per CTA, 7.3 tiles per CTA), and the pair adds lock-step waits on the slower CTA.
bool conv_pair_eligible(const ConvArgs& a) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("VST_TC_PAIR"); on = e ? atoi(e) : 0; }
    return on && a.Cin == 64 && a.Cout == 256 && a.epi <= EPI_SUB && a.Hin == a.Hout && a.Win == a.Wout && a.Hin >= 2 && a.Win >= 2 &&
           num_sms() >= 2;
}

// a.in is an H8 split-half tensor; a.w packed by launch_pack_tc_half_weights with N = 128
int launch_conv3x3_pair(const ConvArgs& a, cudaStream_t st) {
    using Cfg = PairCfg;
    VST_REQUIRE(conv_pair_eligible(a), "conv3x3_pair: shape %d>%d not eligible", a.Cin, a.Cout);
    static PerDeviceOnce smem_once;
    auto kern = conv3x3_pair_kernel;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    PairTiles tl;
    tl.n_xt = cdiv(a.Wout, 128);
    tl.n_tiles = tl.n_xt * a.Hout;
    tl.n_ptiles = cdiv(tl.n_tiles, 2);
    tl.trace = tc_trace_buffer(a.Cin, a.Cout, st);
    const int grid = 2 * std::min(tl.n_ptiles, num_sms() / 2);
    const double px = (double)a.Hout * a.Wout;
    const bool coupled = a.epi >= EPI_ADD;
    ProfScope prof(st, "conv3x3_pair 64>256", 2.0 * 9 * a.Cin * a.Cout * px,
                   4.0 * ((double)a.Cin * a.Hin * a.Win + (coupled ? 2.0 : 1.0) * a.Cout * px));
    VST_CUDA_OK(launch_pdl(kern, grid, Cfg::THREADS, Cfg::SMEM, st, a, tl));
    return check_launch("conv3x3_pair");
}

}  // namespace vst
