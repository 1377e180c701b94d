// conv_direct.cu — fp32 (CUDA-core FFMA) 3x3 reflection-padded convolution with fused epilogues,
// plus the pure data-movement kernels of the reversible network (space<->depth, latent spread).
//
// Replaces, per launch: nn.ReflectionPad2d(1) + nn.Conv2d(3x3, bias) (+ nn.ReLU | + additive
// coupling) of the reference's residual_block (models/RevResNet.py:79-88, :96-116).
//
// Layout: activations are planar fp32 [C][H][W] (one sample).  Weights are repacked once to
// [Cin][9][CoutPad] (cout innermost) so a warp reads a contiguous cout vector per (cin, tap).
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
__device__ __forceinline__ int reflect_clamp(int i, int n) {
    // ReflectionPad2d(1): -1 -> 1, n -> n-2 ; positions further out only occur in tile overhang
    // whose results are discarded, so they are clamped to stay in bounds.
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

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
    const size_t in_plane = (size_t)a.Hin * a.Win;

    float acc[PT][CO];
#pragma unroll
    for (int j = 0; j < PT; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;

    for (int c0 = 0; c0 < a.Cin; c0 += KC) {
        // ---- stage the reflection-padded input tile
        for (int i = tid; i < KC * IH * IW; i += 256) {
            int c = i / (IH * IW);
            int r = i - c * (IH * IW);
            int iy = r / IW, ix = r - iy * IW;
            int gy = reflect_clamp(y0 * S - 1 + iy, a.Hin);
            int gx = reflect_clamp(x0 * S - 1 + ix, a.Win);
            float v = 0.f;
            if (c0 + c < a.Cin) v = __ldg(a.in + (size_t)(c0 + c) * in_plane + (size_t)gy * a.Win + gx);
            in_s[i] = v;
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
    const size_t out_plane = (size_t)a.Hout * a.Wout;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int y = y0 + wy * PT + j;
        if (y >= a.Hout) continue;
#pragma unroll
        for (int q = 0; q < CO; ++q) {
            const int co = co_tile + wc * CO + q;
            if (co >= a.Cout) continue;
            float val = acc[j][q] + __ldg(a.bias + co);
            const size_t o = (size_t)co * out_plane + (size_t)y * a.Wout + x;
            switch (a.epi) {
                case EPI_RELU: a.out[o] = fmaxf(val, 0.f); break;
                case EPI_NONE: a.out[o] = val; break;
                case EPI_ADD: a.out[o] = val + a.res[o]; break;
                case EPI_SUB: a.out[o] = a.res[o] - val; break;
                case EPI_ADD_SQZ: {   // res is [Cout/4][2Hout][2Wout], squeezed on the fly
                    int cq = a.Cout >> 2, k = co / cq, ch = co - k * cq;
                    size_t r = (size_t)ch * (out_plane * 4) + (size_t)(2 * y + (k >> 1)) * (2 * a.Wout) + (2 * x + (k & 1));
                    a.out[o] = val + a.res[r];
                } break;
                case EPI_SUB_UNSQZ: { // out is [Cout/4][2Hout][2Wout], unsqueezed on the fly
                    int cq = a.Cout >> 2, k = co / cq, ch = co - k * cq;
                    size_t r = (size_t)ch * (out_plane * 4) + (size_t)(2 * y + (k >> 1)) * (2 * a.Wout) + (2 * x + (k & 1));
                    a.out[r] = a.res[o] - val;
                } break;
            }
        }
    }
}

template <int S, int CO, int WC, int PT, int KC>
static int launch_cfg(const ConvArgs& a, cudaStream_t st) {
    using Cfg = ConvCfg<S, CO, WC, PT, KC>;
    static bool attr_set = false;
    auto kern = conv3x3_ffma_kernel<S, CO, WC, PT, KC>;
    if (!attr_set) {
        VST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
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

// ------------------------------------------------------------------------------------------
// space <-> depth  (models/RevResNet.py:34-43)   out[(dy*2+dx)*C + c][h][w] = in[c][2h+dy][2w+dx]
// ------------------------------------------------------------------------------------------
__global__ void space_to_depth_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int Ho, int Wo) {
    size_t total = (size_t)4 * C * Ho * Wo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int w = (int)(i % Wo);
        int h = (int)((i / Wo) % Ho);
        int co = (int)(i / ((size_t)Wo * Ho));
        int k = co / C, c = co - k * C;
        out[i] = __ldg(in + ((size_t)c * (2 * Ho) + (2 * h + (k >> 1))) * (2 * Wo) + 2 * w + (k & 1));
    }
}
__global__ void depth_to_space_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int Hi, int Wi) {
    // in [4C][Hi][Wi] -> out [C][2Hi][2Wi]; iterate over output for coalesced writes
    size_t total = (size_t)4 * C * Hi * Wi;
    int Wo = 2 * Wi, Ho = 2 * Hi;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int X = (int)(i % Wo);
        int Y = (int)((i / Wo) % Ho);
        int c = (int)(i / ((size_t)Wo * Ho));
        int k = (Y & 1) * 2 + (X & 1);
        out[i] = __ldg(in + ((size_t)(k * C + c) * Hi + (Y >> 1)) * Wi + (X >> 1));
    }
}

static int ew_grid(size_t total) { return (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16); }

int launch_space_to_depth(const float* in, float* out, int C, int Hin, int Win, cudaStream_t st) {
    size_t total = (size_t)C * Hin * Win;
    ProfScope prof(st, "space_to_depth", 0.0, 8.0 * total);
    space_to_depth_kernel<<<ew_grid(total), 256, 0, st>>>(in, out, C, Hin / 2, Win / 2);
    return check_launch("space_to_depth");
}
int launch_depth_to_space(const float* in, float* out, int Cout, int Hin, int Win, cudaStream_t st) {
    size_t total = (size_t)4 * Cout * Hin * Win;
    ProfScope prof(st, "depth_to_space", 0.0, 8.0 * total);
    depth_to_space_kernel<<<ew_grid(total), 256, 0, st>>>(in, out, Cout, Hin, Win);
    return check_launch("depth_to_space");
}

// ------------------------------------------------------------------------------------------
// latent spread / gather  (models/RevResNet.py:140-144, :149-152): merge(x1,x2) followed by
// sp_steps depth-to-space levels, in one pass;   z [Cz][h<<L][w<<L]  <->  x1,x2 [Ch][h][w]
// ------------------------------------------------------------------------------------------
template <bool TO_LATENT>
__global__ void latent_spread_kernel(float* __restrict__ x1, float* __restrict__ x2, float* __restrict__ z, int Ch,
                                     int h, int w, int L) {
    const int Cz = (2 * Ch) >> (2 * L);
    const int H = h << L, W = w << L;
    size_t total = (size_t)Cz * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int X = (int)(i % W);
        int Y = (int)((i / W) % H);
        int c = (int)(i / ((size_t)W * H));
        int D = Cz, y = Y, x = X;
        for (int l = 0; l < L; ++l) {
            c = ((y & 1) * 2 + (x & 1)) * D + c;
            D *= 4; y >>= 1; x >>= 1;
        }
        float* src = (c < Ch) ? (x1 + ((size_t)c * h + y) * w + x) : (x2 + ((size_t)(c - Ch) * h + y) * w + x);
        if (TO_LATENT) z[i] = *src; else *src = z[i];
    }
}

int launch_latent_spread(const float* x1, const float* x2, float* z, int Ch, int h, int w, int L, cudaStream_t st) {
    size_t total = (size_t)2 * Ch * h * w;
    ProfScope prof(st, "latent_spread", 0.0, 8.0 * total);
    latent_spread_kernel<true><<<ew_grid(total), 256, 0, st>>>(const_cast<float*>(x1), const_cast<float*>(x2), z, Ch, h, w, L);
    return check_launch("latent_spread");
}
int launch_latent_gather(const float* z, float* x1, float* x2, int Ch, int h, int w, int L, cudaStream_t st) {
    size_t total = (size_t)2 * Ch * h * w;
    ProfScope prof(st, "latent_gather", 0.0, 8.0 * total);
    latent_spread_kernel<false><<<ew_grid(total), 256, 0, st>>>(x1, x2, const_cast<float*>(z), Ch, h, w, L);
    return check_launch("latent_gather");
}

// ------------------------------------------------------------------------------------------
// frame format conversion (video_transfer.py:188, :211-214)
// ------------------------------------------------------------------------------------------
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ hwc, float* __restrict__ chw, int n, int bgr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int sc = bgr ? 2 - c : c;
            chw[(size_t)c * n + i] = (float)hwc[(size_t)i * 3 + sc] / 255.f;   // ToTensor: byte / 255
        }
    }
}
__global__ void f32_to_u8_kernel(const float* __restrict__ chw, uint8_t* __restrict__ hwc, int n, int bgr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int dc = bgr ? 2 - c : c;
            float v = chw[(size_t)c * n + i] * 255.f;
            v = fminf(fmaxf(v, 0.f), 255.f);                                 // mul(255).clamp(0,255)
            hwc[(size_t)i * 3 + dc] = (uint8_t)v;                             // .byte() truncates
        }
    }
}

}  // namespace vst

extern "C" int vst_frame_u8_to_f32(const uint8_t* hwc, float* chw, int H, int W, int bgr, void* stream) {
    using namespace vst;
    VST_REQUIRE(hwc && chw && H > 0 && W > 0, "vst_frame_u8_to_f32: bad arguments");
    int n = H * W;
    u8_to_f32_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(hwc, chw, n, bgr);
    return check_launch("u8_to_f32");
}
extern "C" int vst_frame_f32_to_u8(const float* chw, uint8_t* hwc, int H, int W, int bgr, void* stream) {
    using namespace vst;
    VST_REQUIRE(hwc && chw && H > 0 && W > 0, "vst_frame_f32_to_u8: bad arguments");
    int n = H * W;
    f32_to_u8_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(chw, hwc, n, bgr);
    return check_launch("f32_to_u8");
}
