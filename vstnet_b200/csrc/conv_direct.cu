// conv_direct.cu — fp32 (CUDA-core FFMA) 3x3 reflection-padded convolution with fused epilogues,
// plus the pure data-movement kernels of the reversible network (space<->depth, latent spread).
//
// Replaces, per launch: nn.ReflectionPad2d(1) + nn.Conv2d(3x3, bias) (+ nn.ReLU | + additive
// coupling) of the reference's residual_block (models/RevResNet.py:79-88, :96-116).
//
// Layout: activations are P4 (kernels.cuh): [C/4][H+2][W+2][4] fp32 with the reflection border
// already in place, so staging a tile is a plain rectangular read.  Weights are repacked once to
// [Cin][9][CoutPad] (cout innermost) so a warp reads a contiguous cout vector per (cin, tap).
// This kernel serves the layers the tcgen05 kernel does not take (stride 2, Cin or Cout of 4).
#include "kernels.cuh"

namespace vst {

// ------------------------------------------------------------------------------------------
// weight repack: OIHW [Cout][Cin][3][3] -> [Cin][9][CoutPad], bias -> [CoutPad] (zero padded)
// ------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                         float* __restrict__ wp, float* __restrict__ bp, int Cin, int Cout,
                                         int CoutPad) {
    int total = Cin * 9 * CoutPad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int co = i % CoutPad;
        int t = (i / CoutPad) % 9;
        int ci = i / (CoutPad * 9);
        wp[i] = (co < Cout) ? w[((size_t)co * Cin + ci) * 9 + t] : 0.f;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CoutPad; i += gridDim.x * blockDim.x)
        bp[i] = (i < Cout) ? b[i] : 0.f;
}

int launch_pack_conv_weights(const float* w, const float* b, float* wp, float* bp, int Cin, int Cout, int CoutPad,
                             cudaStream_t st) {
    int total = Cin * 9 * CoutPad;
    pack_conv_weights_kernel<<<cdiv(total, 256), 256, 0, st>>>(w, b, wp, bp, Cin, Cout, CoutPad);
    return check_launch("pack_conv_weights");
}

// ------------------------------------------------------------------------------------------
// direct convolution
// ------------------------------------------------------------------------------------------
template <int S, int CO, int WC, int PT, int KC>
struct ConvCfg {
    static constexpr int WY = 8 / WC;
    static constexpr int TW = 32;
    static constexpr int TH = WY * PT;
    static constexpr int CT = WC * CO;
    static constexpr int IH = (TH - 1) * S + 3;
    static constexpr int IW = (TW - 1) * S + 3;
    static constexpr int IN_FLOATS = ((KC * IH * IW + 3) / 4) * 4;
    static constexpr int W_FLOATS = KC * 9 * CT;
    static constexpr size_t SMEM = (size_t)(IN_FLOATS + W_FLOATS) * sizeof(float);
};

template <int S, int CO, int WC, int PT, int KC>
__global__ void __launch_bounds__(256) conv3x3_ffma_kernel(ConvArgs a) {
    using Cfg = ConvCfg<S, CO, WC, PT, KC>;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, CT = Cfg::CT, IH = Cfg::IH, IW = Cfg::IW;
    constexpr int ROWS = (PT - 1) * S + 3;
    extern __shared__ __align__(16) float smem[];
    float* in_s = smem;
    float* w_s = smem + Cfg::IN_FLOATS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wc = warp % WC, wy = warp / WC;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int co_tile = blockIdx.z * CT;
    const float4* in4 = reinterpret_cast<const float4*>(a.in);

    float acc[PT][CO];
#pragma unroll
    for (int j = 0; j < PT; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;

    for (int c0 = 0; c0 < a.Cin; c0 += KC) {
        // ---- stage the input tile (border already reflected in the P4 tensor)
        for (int i = tid; i < (KC / 4) * IH * IW; i += 256) {
            int g = i / (IH * IW);
            int r = i - g * (IH * IW);
            int iy = r / IW, ix = r - iy * IW;
            int py = min(y0 * S + iy, a.Hin + 1);       // padded coords; clamp only matters in tile overhang
            int px = min(x0 * S + ix, a.Win + 1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + 4 * g < a.Cin)
                v = __ldg(in4 + ((size_t)(c0 / 4 + g) * (a.Hin + 2) + py) * (a.Win + 2) + px);
            float* d = in_s + (4 * g) * (IH * IW) + r;
            d[0] = v.x; d[IH * IW] = v.y; d[2 * IH * IW] = v.z; d[3 * IH * IW] = v.w;
        }
        // ---- stage the weights of this cin chunk / cout tile
        for (int i = tid; i < KC * 9 * (CT / 4); i += 256) {
            int row = i / (CT / 4);      // c*9 + tap
            int j = i - row * (CT / 4);
            int c = row / 9;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + c < a.Cin)
                v = __ldg(reinterpret_cast<const float4*>(a.w + ((size_t)c0 * 9 + row) * a.CoutPad + co_tile) + j);
            reinterpret_cast<float4*>(w_s)[i] = v;
        }
        __syncthreads();

#pragma unroll 2
        for (int c = 0; c < KC; ++c) {
            float v[ROWS][3];
            const float* ip = in_s + (c * IH + wy * PT * S) * IW + lane * S;
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) v[r][kx] = ip[r * IW + kx];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float wv[CO];
                    const float4* wp = reinterpret_cast<const float4*>(w_s + (c * 9 + ky * 3 + kx) * CT + wc * CO);
#pragma unroll
                    for (int q = 0; q < CO / 4; ++q) {
                        float4 t = wp[q];
                        wv[4 * q + 0] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
                    }
#pragma unroll
                    for (int j = 0; j < PT; ++j)
#pragma unroll
                        for (int q = 0; q < CO; ++q) acc[j][q] = fmaf(v[j * S + ky][kx], wv[q], acc[j][q]);
                }
        }
        __syncthreads();
    }

    // ---- epilogue
    const int x = x0 + lane;
    if (x >= a.Wout) return;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int y = y0 + wy * PT + j;
        if (y >= a.Hout) continue;
#pragma unroll
        for (int q = 0; q < CO / 4; ++q) {
            const int co = co_tile + wc * CO + 4 * q;
            if (co >= a.Cout) continue;
            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + co));
            conv_epilogue(a, co >> 2, y, x, make_float4(acc[j][4 * q] + bv.x, acc[j][4 * q + 1] + bv.y,
                                                        acc[j][4 * q + 2] + bv.z, acc[j][4 * q + 3] + bv.w));
        }
    }
}

template <int S, int CO, int WC, int PT, int KC>
static int launch_cfg(const ConvArgs& a, cudaStream_t st) {
    using Cfg = ConvCfg<S, CO, WC, PT, KC>;
    static PerDeviceOnce smem_once;
    auto kern = conv3x3_ffma_kernel<S, CO, WC, PT, KC>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    dim3 grid(cdiv(a.Wout, Cfg::TW), cdiv(a.Hout, Cfg::TH), cdiv(a.Cout, Cfg::CT));
    char cls[40];
    snprintf(cls, sizeof(cls), "conv3x3_ffma %d>%d s%d", a.Cin, a.Cout, S);
    const double px = (double)a.Hout * a.Wout;
    const bool coupled = a.epi >= EPI_ADD;
    ProfScope prof(st, cls, 2.0 * 9 * a.Cin * a.Cout * px,
                   4.0 * ((double)a.Cin * a.Hin * a.Win + (coupled ? 2.0 : 1.0) * a.Cout * px));
    kern<<<grid, 256, Cfg::SMEM, st>>>(a);
    return check_launch("conv3x3_ffma");
}

int conv_cout_pad(int Cout) {
    if (Cout <= 4) return 4;
    if (Cout <= 16) return 16;
    return (Cout + 63) / 64 * 64;
}

template <int S>
static int launch_s(const ConvArgs& a, cudaStream_t st) {
    constexpr int KCB = (S == 1) ? 8 : 4;
    if (a.CoutPad == 4) {
        if (a.Cin <= 4) return launch_cfg<S, 4, 1, 2, 4>(a, st);
        return launch_cfg<S, 4, 1, 2, KCB>(a, st);
    }
    if (a.CoutPad == 16) {
        if (a.Cin <= 4) return launch_cfg<S, 16, 1, 2, 4>(a, st);
        return launch_cfg<S, 16, 1, 2, KCB>(a, st);
    }
    if (a.Cin <= 4) return launch_cfg<S, 16, 4, 4, 4>(a, st);
    return launch_cfg<S, 16, 4, 4, KCB>(a, st);
}

int launch_conv3x3_ffma(const ConvArgs& a, int stride, cudaStream_t st) {
    VST_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride %d unsupported", stride);
    VST_REQUIRE(a.Hin >= 2 && a.Win >= 2, "conv3x3: reflection pad needs H,W >= 2 (got %dx%d)", a.Hin, a.Win);
    return stride == 1 ? launch_s<1>(a, st) : launch_s<2>(a, st);
}

}  // namespace vst
