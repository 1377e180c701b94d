// block16.cu — one whole reversible block of the full-resolution stage in ONE kernel (CUDA-core fp32).
//
// Replaces residual_block.forward / .inverse (models/RevResNet.py:96-116) for channel = 16,
// mult = 4, stride = 1:   out = res +/- conv3(relu(conv2(relu(conv1(refpad(x))))))
// with 16 -> 4 -> 4 -> 16 channels, a ReflectionPad2d(1) before every conv (:79-88).
//
// Why not tensor cores: N = 4 output channels is below the tcgen05 minimum and the stage is
// HBM-bound anyway (13.5 FLOP/B); what matters is that the 4-channel intermediates
// never touch HBM.  A CTA owns a 28x28 output tile: it streams the 34x34x16 input halo tile
// through shared memory, computes conv1 on 32x32, conv2 on 30x30, conv3 on 28x28, and reads/writes
// the coupling operand once.  HBM traffic per block: x read once (+halo), res read once, out written
// once — 3 half-states instead of the 4+ of three separate conv launches.
//
// Reflection inside the fused tile: intermediates exist only at in-image positions; after each
// conv the tile positions that stand for image row/col -1 or H/W are overwritten with their
// mirror (row 1 / H-2), which is exactly what the next ReflectionPad2d(1) would read.
//
// Data path: the input halo tile arrives by TMA bulk copies (one 544-byte segment per group and
// row, straight from the P4 tensor, completion on an mbarrier), one 4-channel group at a time
// through a 2-slot ring so that loads overlap conv1 and three CTAs fit on an SM; it stays
// 4-channel interleaved in shared memory, as do the two intermediates, so every shared-memory
// access is a conflict-free 16-byte vector.  Thread mapping: 4 warps; lane = tile column, each warp owns a band of rows and
// every thread a vertical strip of that band (a 10-row register window per kx serves the three
// ky taps), all output channels of a pass in registers.  Weights sit in shared memory in the
// [cin][tap][cout] layout of conv_direct.cu's pack and are read as warp-uniform float4 broadcasts.
#include "kernels.cuh"

namespace vst {

namespace b16 {
constexpr int TO = 28;   // output tile edge
constexpr int T1 = 32;   // conv1 tile edge (TO + 4) == warp width
constexpr int T2 = 30;   // conv2 tile edge
constexpr int XT = 34;   // input tile edge
constexpr int T2_PITCH = 32;
constexpr int XG_F4 = XT * XT;         // one 4-channel group of the input halo tile, [row][col] float4
constexpr int T1_F4 = T1 * T1;         // [row][col] float4 (the 4 bottleneck channels)
constexpr int T2_F4 = T2 * T2_PITCH;
constexpr int W_FLOATS = 16 * 9 * 4 + 4 + 4 * 9 * 4 + 4 + 4 * 9 * 16 + 16;
constexpr size_t SMEM = (size_t)(2 * XG_F4 + T1_F4 + T2_F4) * 16 + (size_t)W_FLOATS * 4 + 32;
}  // namespace b16

__device__ __forceinline__ uint32_t b16_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// overwrite the tile rows/cols that stand for image row/col -1 and H/W with their mirrors.
// tile(r, c) <-> image(oy + r, ox + c); square tile of edge n, row pitch `pitch` (float4 elements).
__device__ __forceinline__ void tile_reflect(float4* t, int n, int pitch, int oy, int ox, int H, int W, int tid) {
    // CTA-uniform early out: interior tiles hold no out-of-image positions
    if (oy >= 0 && ox >= 0 && oy + n <= H && ox + n <= W) return;
    {
        const int rm = -1 - oy, rs = 1 - oy;            // row -1 <- row 1
        if (rm >= 0 && rs < n)
            for (int c = tid; c < n; c += 128) t[rm * pitch + c] = t[rs * pitch + c];
        const int rH = H - oy, rS = H - 2 - oy;         // row H <- row H-2
        if (rH < n && rS >= 0)
            for (int c = tid; c < n; c += 128) t[rH * pitch + c] = t[rS * pitch + c];
    }
    __syncthreads();
    {
        const int cm = -1 - ox, cc = 1 - ox;
        if (cm >= 0 && cc < n)
            for (int r = tid; r < n; r += 128) t[r * pitch + cm] = t[r * pitch + cc];
        const int cW = W - ox, cS = W - 2 - ox;
        if (cW < n && cS >= 0)
            for (int r = tid; r < n; r += 128) t[r * pitch + cW] = t[r * pitch + cS];
    }
    __syncthreads();
}

// Packed fp32 FMA (Blackwell FFMA2): (d0, d1) += (a, a) * (b0, b1), IEEE fma per half — bit-identical to two
// fmaf().  A three-register FFMA issues every other cycle per SM sub-partition (register-file read ports);
// the packed form does two FMAs in the same issue slot, which is what reaches the fp32 peak.
__device__ __forceinline__ void b16_fma2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}
#define B16_FMA4(ACC, XV, WV)                          \
    b16_fma2(ACC[0], ACC[1], XV, WV.x, WV.y);          \
    b16_fma2(ACC[2], ACC[3], XV, WV.z, WV.w);

__device__ __forceinline__ void b16_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t ba = b16_smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(ba), "r"(parity) : "memory");
    if (!done) __trap();
}

__global__ void __launch_bounds__(128, 3) rev_block16_kernel(Block16Args a) {
    using namespace b16;
    extern __shared__ __align__(16) float4 sm4[];
    float4* xs = sm4;                     // [2 slots][34][34]  ring of input-tile channel groups
    float4* t1 = sm4 + 2 * XG_F4;         // [32][32]
    float4* t2 = t1 + T1_F4;              // [30][32]
    float* ws = reinterpret_cast<float*>(t2 + T2_F4);
    float* w1s = ws;                      // [16][9][4]
    float* b1s = w1s + 16 * 9 * 4;
    float* w2s = b1s + 4;                 // [4][9][4]
    float* b2s = w2s + 4 * 9 * 4;
    float* w3s = b2s + 4;                 // [4][9][16]
    float* b3s = w3s + 4 * 9 * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b3s + 16);   // [3] one per ring slot + weights

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, Hp = H + 2, Wp = W + 2;
    // persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the input ring runs ahead across tile
    // boundaries, so the next tile's first channel groups stream in while this tile's conv2 / conv3 run
    const int n_tx = (W + TO - 1) / TO, n_tiles = n_tx * ((H + TO - 1) / TO);
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_items = 4 * my_tiles;                     // ring items: (tile, channel group)
    const bool tr0 = a.trace && tid == 0 && blockIdx.x == gridDim.x / 2;
    int trn = 0;
#define B16_STAMP() do { if (tr) a.trace[trn++] = clock64(); } while (0)

    // ---- input halo tile (image rows y0-3 .. y0+30 = padded rows y0-2 ..), one channel group at a
    //      time through a 2-slot ring: one TMA bulk copy per row segment straight from the P4
    //      tensor, completion on the slot's mbarrier; group g+2 streams in while g+1 is computed.
    auto issue_item = [&](int item) {                      // executed by all lanes of warp 0
        const int tile = (int)blockIdx.x + (item >> 2) * (int)gridDim.x, g = item & 3, slot = item & 1;
        const int tx0 = (tile % n_tx) * TO, ty0 = (tile / n_tx) * TO;
        const int shift = tx0 < 2 ? 2 - tx0 : 0;           // tile columns left of the padded tensor
        const uint32_t seg_bytes = (uint32_t)(XT - shift) * 16;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b16_smem_u32(&bars[slot])),
                         "r"(seg_bytes * XT) : "memory");
        __syncwarp();
        const float4* x4 = reinterpret_cast<const float4*>(a.x);
        for (int iy = lane; iy < XT; iy += 32) {
            const int py = min(max(ty0 - 2 + iy, 0), Hp - 1);
            const float4* src = x4 + ((size_t)g * Hp + py) * Wp + (tx0 - 2 + shift);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             b16_smem_u32(xs + slot * XG_F4 + iy * XT + shift)),
                         "l"(src), "r"(seg_bytes), "r"(b16_smem_u32(&bars[slot]))
                         : "memory");
        }
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b16_smem_u32(&bars[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b16_smem_u32(&bars[1])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b16_smem_u32(&bars[2])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {   // weights: the block's three packs are one contiguous 5280-byte run (w1 b1 w2 b2 w3 b3)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b16_smem_u32(&bars[2])),
                         "r"((uint32_t)(W_FLOATS * 4)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             b16_smem_u32(ws)), "l"(a.w1), "r"((uint32_t)(W_FLOATS * 4)), "r"(b16_smem_u32(&bars[2]))
                         : "memory");
        }
        issue_item(0);
        issue_item(1);
    }
    b16_wait(&bars[2], 0);                     // weights landed (once per CTA)

  for (int k = 0; k < my_tiles; ++k) {
    const int tile = (int)blockIdx.x + k * (int)gridDim.x;
    const int x0 = (tile % n_tx) * TO, y0 = (tile / n_tx) * TO;
    const bool tr = tr0 && k == 1;
    B16_STAMP();
    // ---- conv1 16 -> 4 on the 32x32 tile: t1(r, c) <-> image (y0-2+r, x0-2+c); warp = 8 rows, lane = col
    {
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = b1s[c];
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            const int item = 4 * k + g, slot = item & 1;
            b16_wait(&bars[slot], (uint32_t)((item >> 1) & 1));
#pragma unroll 1
            for (int kx = 0; kx < 3; ++kx) {
                float4 v[10];
                const float4* ip = xs + slot * XG_F4 + (warp * 8) * XT + lane + kx;
#pragma unroll
                for (int r = 0; r < 10; ++r) v[r] = ip[r * XT];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float* wp = w1s + ((4 * g) * 9 + ky * 3 + kx) * 4;
                    const float4 wa = *reinterpret_cast<const float4*>(wp);
                    const float4 wb = *reinterpret_cast<const float4*>(wp + 36);
                    const float4 wc = *reinterpret_cast<const float4*>(wp + 72);
                    const float4 wd = *reinterpret_cast<const float4*>(wp + 108);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        B16_FMA4(acc[r], v[r + ky].x, wa)
                        B16_FMA4(acc[r], v[r + ky].y, wb)
                        B16_FMA4(acc[r], v[r + ky].z, wc)
                        B16_FMA4(acc[r], v[r + ky].w, wd)
                    }
                }
            }
            B16_STAMP();
            __syncthreads();                           // every warp is done reading this slot
            if (warp == 0 && item + 2 < n_items) issue_item(item + 2);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            t1[(warp * 8 + r) * T1 + lane] = make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f),
                                                         fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    }
    __syncthreads();
    tile_reflect(t1, T1, T1, y0 - 2, x0 - 2, H, W, tid);
    B16_STAMP();

    // ---- conv2 4 -> 4 on the 30x30 tile: t2(r, c) <-> image (y0-1+r, x0-1+c); reads t1 rows r..r+2
    {
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = b2s[c];
#pragma unroll 1
        for (int kx = 0; kx < 3; ++kx) {
            float4 v[10];
#pragma unroll
            for (int r = 0; r < 10; ++r) v[r] = t1[min(warp * 8 + r, T1 - 1) * T1 + min(lane + kx, T1 - 1)];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* wp = w2s + (ky * 3 + kx) * 4;
                const float4 wa = *reinterpret_cast<const float4*>(wp);
                const float4 wb = *reinterpret_cast<const float4*>(wp + 36);
                const float4 wc = *reinterpret_cast<const float4*>(wp + 72);
                const float4 wd = *reinterpret_cast<const float4*>(wp + 108);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    B16_FMA4(acc[r], v[r + ky].x, wa)
                    B16_FMA4(acc[r], v[r + ky].y, wb)
                    B16_FMA4(acc[r], v[r + ky].z, wc)
                    B16_FMA4(acc[r], v[r + ky].w, wd)
                }
            }
        }
        if (lane < T2)
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (warp * 8 + r < T2)
                    t2[(warp * 8 + r) * T2_PITCH + lane] = make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f),
                                                                       fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    }
    __syncthreads();
    tile_reflect(t2, T2, T2_PITCH, y0 - 1, x0 - 1, H, W, tid);
    B16_STAMP();

    // ---- conv3 4 -> 16 on the 28x28 tile, two passes of 8 output channels; warp = 7 rows
    const int x = x0 + lane;
    const bool xin = lane < TO && x < W;
    const float4* res4 = reinterpret_cast<const float4*>(a.res);
    float4* out4 = reinterpret_cast<float4*>(a.out);
    const size_t plane = p4_plane_px(H, W);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        // coupling operand first: its global latency hides behind the FMAs below
        float4 rv[7][2];
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int y = y0 + warp * 7 + r;
#pragma unroll
            for (int j = 0; j < 2; ++j)
                rv[r][j] = (xin && y < H) ? res4[(size_t)(half * 2 + j) * plane + (size_t)(y + 1) * Wp + x + 1]
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float acc[7][8];
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = b3s[half * 8 + c];
#pragma unroll 1
        for (int kx = 0; kx < 3; ++kx) {
            float4 v[9];
#pragma unroll
            for (int r = 0; r < 9; ++r) v[r] = t2[(warp * 7 + r) * T2_PITCH + min(lane + kx, T2 - 1)];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const float* wp = w3s + (ci * 9 + ky * 3 + kx) * 16 + half * 8;
                    const float4 wa = *reinterpret_cast<const float4*>(wp);
                    const float4 wb = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const float xv = ci == 0 ? v[r + ky].x : ci == 1 ? v[r + ky].y : ci == 2 ? v[r + ky].z : v[r + ky].w;
                        B16_FMA4(acc[r], xv, wa)
                        B16_FMA4((acc[r] + 4), xv, wb)
                    }
                }
            }
        }
        B16_STAMP();
        // ---- additive coupling + store (RevResNet.py:103, :110-111)
        if (xin) {
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const int y = y0 + warp * 7 + r;
                if (y >= H) continue;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 q = rv[r][j];
                    float4 o;
                    if (a.sub) o = make_float4(q.x - acc[r][4 * j], q.y - acc[r][4 * j + 1], q.z - acc[r][4 * j + 2], q.w - acc[r][4 * j + 3]);
                    else o = make_float4(q.x + acc[r][4 * j], q.y + acc[r][4 * j + 1], q.z + acc[r][4 * j + 2], q.w + acc[r][4 * j + 3]);
                    p4_store(out4 + (size_t)(half * 2 + j) * plane, H, W, y, x, o);
                }
            }
        }
        B16_STAMP();
    }
  }   // tile loop
}

int launch_rev_block16(const Block16Args& a, cudaStream_t st) {
    static PerDeviceOnce smem_once;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, rev_block16_kernel, (int)b16::SMEM));
    VST_REQUIRE(a.H >= 4 && a.W >= 4, "rev_block16: image too small (%dx%d)", a.H, a.W);
    VST_REQUIRE(a.b1 == a.w1 + 576 && a.w2 == a.b1 + 4 && a.b2 == a.w2 + 144 && a.w3 == a.b2 + 4 && a.b3 == a.w3 + 576 &&
                    ((uintptr_t)a.w1 & 15) == 0,
                "rev_block16: the block's weight packs must be one contiguous 16-byte aligned run");
    const int n_tiles = cdiv(a.W, b16::TO) * cdiv(a.H, b16::TO);
    const int grid = std::min(n_tiles, 3 * num_sms());
    Block16Args aa = a;
    aa.trace = tc_trace_buffer(16, 4, st);
    const double px = (double)a.H * a.W;
    ProfScope prof(st, "rev_block16 (16>4>4>16)", 2.0 * 9 * (16 * 4 + 4 * 4 + 4 * 16) * px, 3.0 * 64.0 * px);
    rev_block16_kernel<<<grid, 128, b16::SMEM, st>>>(aa);
    return check_launch("rev_block16");
}

}  // namespace vst
