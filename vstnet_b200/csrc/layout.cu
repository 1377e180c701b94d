// layout.cu — the pure data-movement steps of the reversible network, on P4 tensors (kernels.cuh):
// injective_pad (RevResNet.py:19-31), squeeze / unsqueeze (:34-43), the channel_reduction spread
// loops fused with split / merge (:8-16, :140-152), and the frame format conversion of the video
// entry point.  Every kernel that writes a P4 tensor also writes its reflection border (p4_store).
#include "kernels.cuh"

namespace vst {

static int ew_grid(size_t total) { return (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16); }

// ------------------------------------------------------------------------------------------
// injective_pad.forward + split (RevResNet.py:24-28, :8-12): x NCHW [Cimg][H][W] -> s0 P4 with C0
// channels, channels >= Cimg zero.        injective_pad.inverse (:30-31): first Cimg channels of s0.
// ------------------------------------------------------------------------------------------
// `add` (optional): per-channel constant added to every pixel — F(0) of the first block when that block is folded
// into this kernel (revnet.cu, fold_first_block): float4 per group, read from pixel (add_off) of a tiny P4 tensor.
__device__ __forceinline__ float4 state_add_const(const float4* add, size_t add_plane, size_t add_off, int g) {
    return add ? __ldg(add + (size_t)g * add_plane + add_off) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void image_to_state_kernel(const float* __restrict__ x, float4* __restrict__ s0, int Cimg, int G, int H,
                                      int W, int* __restrict__ status_clear, const float4* __restrict__ add,
                                      size_t add_plane, size_t add_off) {
    if (status_clear && blockIdx.x == 0 && threadIdx.x == 0) *status_clear = 0;     // first kernel of an encode
    const size_t n = (size_t)H * W, total = (size_t)G * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / n);
        const size_t p = i - (size_t)g * n;
        const int y = (int)(p / W), xx = (int)(p - (size_t)y * W);
        const float4 c = state_add_const(add, add_plane, add_off, g);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (4 * g + e < Cimg) ? __ldg(x + (size_t)(4 * g + e) * n + p) : 0.f;
        p4_store(s0 + (size_t)g * p4_plane_px(H, W), H, W, y, xx, make_float4(v[0] + c.x, v[1] + c.y, v[2] + c.z, v[3] + c.w));
    }
}
__global__ void state_to_image_kernel(const float4* __restrict__ s0, float* __restrict__ x, int Cimg, int H, int W) {
    const size_t n = (size_t)H * W, total = (size_t)((Cimg + 3) / 4) * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / n);
        const size_t p = i - (size_t)g * n;
        const int y = (int)(p / W), xx = (int)(p - (size_t)y * W);
        const float4 v = __ldg(s0 + (size_t)g * p4_plane_px(H, W) + (size_t)(y + 1) * (W + 2) + xx + 1);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * g + k < Cimg) x[(size_t)(4 * g + k) * n + p] = e[k];
    }
}

// The same straight from / to the video frame format, uint8 HWC (RGB or BGR): ToTensor's byte / 255 and the
// mul(255).clamp(0,255).byte() truncation of video_transfer.py:188, :211-214 folded into the first / last kernel of the
// pass (bit-identical to vst_frame_u8_to_f32 / vst_frame_f32_to_u8 around the fp32 kernels).
__global__ void image_u8_to_state_kernel(const uint8_t* __restrict__ hwc, float4* __restrict__ s0, int G, int H, int W, int bgr,
                                         int* __restrict__ status_clear, const float4* __restrict__ add, size_t add_plane,
                                         size_t add_off) {
    if (status_clear && blockIdx.x == 0 && threadIdx.x == 0) *status_clear = 0;
    const size_t n = (size_t)H * W, total = (size_t)G * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / n);
        const size_t p = i - (size_t)g * n;
        const int y = (int)(p / W), xx = (int)(p - (size_t)y * W);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g == 0) {
            const uint8_t* px = hwc + p * 3;
            const float r = (float)px[bgr ? 2 : 0] / 255.f, gg = (float)px[1] / 255.f, b = (float)px[bgr ? 0 : 2] / 255.f;
            v = make_float4(r, gg, b, 0.f);
        }
        const float4 c = state_add_const(add, add_plane, add_off, g);
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
        p4_store(s0 + (size_t)g * p4_plane_px(H, W), H, W, y, xx, v);
    }
}
__global__ void state_to_image_u8_kernel(const float4* __restrict__ s0, uint8_t* __restrict__ hwc, int H, int W, int bgr) {
    const size_t n = (size_t)H * W;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), xx = (int)(p - (size_t)y * W);
        const float4 v = __ldg(s0 + (size_t)(y + 1) * (W + 2) + xx + 1);
        const float e[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float q = fminf(fmaxf(e[c] * 255.f, 0.f), 255.f);        // mul(255).clamp(0,255)
            hwc[p * 3 + (bgr ? 2 - c : c)] = (uint8_t)q;                     // .byte() truncates
        }
    }
}
int launch_image_u8_to_state(const uint8_t* hwc, float* s0, int C0, int H, int W, int bgr, int* status_clear,
                             const float* add, int add_h, int add_w, cudaStream_t st) {
    const size_t total = (size_t)(C0 / 4) * H * W;
    ProfScope prof(st, "image_u8_to_state", 0.0, 3.0 * H * W + 16.0 * total);
    image_u8_to_state_kernel<<<ew_grid(total), 256, 0, st>>>(hwc, reinterpret_cast<float4*>(s0), C0 / 4, H, W, bgr, status_clear,
                                                            reinterpret_cast<const float4*>(add), p4_plane_px(add_h, add_w),
                                                            (size_t)(add_h / 2 + 1) * (add_w + 2) + add_w / 2 + 1);
    return check_launch("image_u8_to_state");
}
int launch_state_to_image_u8(const float* s0, uint8_t* hwc, int H, int W, int bgr, cudaStream_t st) {
    const size_t total = (size_t)H * W;
    ProfScope prof(st, "state_to_image_u8", 0.0, 16.0 * total + 3.0 * total);
    state_to_image_u8_kernel<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float4*>(s0), hwc, H, W, bgr);
    return check_launch("state_to_image_u8");
}

int launch_image_to_state(const float* x, float* s0, int Cimg, int C0, int H, int W, int* status_clear, const float* add,
                          int add_h, int add_w, cudaStream_t st) {
    const size_t total = (size_t)(C0 / 4) * H * W;
    ProfScope prof(st, "image_to_state", 0.0, 4.0 * Cimg * H * W + 16.0 * total);
    image_to_state_kernel<<<ew_grid(total), 256, 0, st>>>(x, reinterpret_cast<float4*>(s0), Cimg, C0 / 4, H, W, status_clear,
                                                         reinterpret_cast<const float4*>(add), p4_plane_px(add_h, add_w),
                                                         (size_t)(add_h / 2 + 1) * (add_w + 2) + add_w / 2 + 1);
    return check_launch("image_to_state");
}
int launch_state_to_image(const float* s0, float* x, int Cimg, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)((Cimg + 3) / 4) * H * W;
    ProfScope prof(st, "state_to_image", 0.0, 16.0 * total + 4.0 * Cimg * H * W);
    state_to_image_kernel<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float4*>(s0), x, Cimg, H, W);
    return check_launch("state_to_image");
}

// ------------------------------------------------------------------------------------------
// space <-> depth  (RevResNet.py:34-43)   out[(dy*2+dx)*C + c][h][w] = in[c][2h+dy][2w+dx]
// With C % 4 == 0 a 4-channel group moves as one 16-byte unit: out group k*(C/4)+g <- in group g.
// ------------------------------------------------------------------------------------------
__global__ void space_to_depth_kernel(const float4* __restrict__ in, float4* __restrict__ out, int G, int Ho, int Wo) {
    const size_t n = (size_t)Ho * Wo, total = (size_t)4 * G * n;
    const int Hi = 2 * Ho, Wi = 2 * Wo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int go = (int)(i / n);
        const size_t p = i - (size_t)go * n;
        const int h = (int)(p / Wo), w = (int)(p - (size_t)h * Wo);
        const int k = go / G, g = go - k * G;
        const float4 v = __ldg(in + (size_t)g * p4_plane_px(Hi, Wi) + (size_t)(2 * h + (k >> 1) + 1) * (Wi + 2) + 2 * w + (k & 1) + 1);
        p4_store(out + (size_t)go * p4_plane_px(Ho, Wo), Ho, Wo, h, w, v);
    }
}
__global__ void depth_to_space_kernel(const float4* __restrict__ in, float4* __restrict__ out, int G, int Hi, int Wi) {
    // in [4G groups][Hi][Wi] -> out [G groups][2Hi][2Wi]; iterate over the output
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    const size_t n = (size_t)Ho * Wo, total = (size_t)G * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / n);
        const size_t p = i - (size_t)g * n;
        const int Y = (int)(p / Wo), X = (int)(p - (size_t)Y * Wo);
        const int k = (Y & 1) * 2 + (X & 1);
        const float4 v = __ldg(in + (size_t)(k * G + g) * p4_plane_px(Hi, Wi) + (size_t)((Y >> 1) + 1) * (Wi + 2) + (X >> 1) + 1);
        p4_store(out + (size_t)g * p4_plane_px(Ho, Wo), Ho, Wo, Y, X, v);
    }
}

// row -1 <- row 0 and column -1 <- column 0 (corner included) for every group: the border the squeezed-input
// form of the stride-2 conv needs (conv_tch.cu, pack_tch_s2_weights_kernel).  O(perimeter).
__global__ void p4_replicate_topleft_kernel(float4* __restrict__ t, int G, int H, int W) {
    const int Wp = W + 2, per = (W + 1) + H;          // top row incl. corner, then the left column
    const size_t total = (size_t)G * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / per), k = (int)(i - (size_t)g * per);
        float4* p = t + (size_t)g * p4_plane_px(H, W);
        if (k <= W) {                                  // padded row 0, padded columns 0..W  <- image row 0, column max(k-1, 0)
            p[k] = p[(size_t)Wp + (k == 0 ? 1 : k)];
        } else {                                       // padded column 0, padded rows 1..H  <- image column 0
            const int r = k - W;
            p[(size_t)r * Wp] = p[(size_t)r * Wp + 1];
        }
    }
}
int launch_p4_replicate_topleft(float* t, int C, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)(C / 4) * (W + 1 + H);
    p4_replicate_topleft_kernel<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<float4*>(t), C / 4, H, W);
    return check_launch("p4_replicate_topleft");
}

int launch_space_to_depth(const float* in, float* out, int C, int Hin, int Win, cudaStream_t st) {
    VST_REQUIRE(C % 4 == 0, "space_to_depth: C must be a multiple of 4");
    const size_t total = (size_t)C * Hin * Win / 4;
    ProfScope prof(st, "space_to_depth", 0.0, 32.0 * total);
    space_to_depth_kernel<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                         C / 4, Hin / 2, Win / 2);
    return check_launch("space_to_depth");
}
int launch_depth_to_space(const float* in, float* out, int Cout, int Hin, int Win, cudaStream_t st) {
    VST_REQUIRE(Cout % 4 == 0, "depth_to_space: C must be a multiple of 4");
    const size_t total = (size_t)Cout * Hin * Win;
    ProfScope prof(st, "depth_to_space", 0.0, 32.0 * total);
    depth_to_space_kernel<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                         Cout / 4, Hin, Win);
    return check_launch("depth_to_space");
}

// ------------------------------------------------------------------------------------------
// latent spread / gather  (RevResNet.py:140-144, :149-152): merge(x1,x2) followed by sp_steps
// depth-to-space levels, in one pass, converting between the network's P4 state and the NCHW
// latent the cWCT API exchanges:   z NCHW [Cz][h<<L][w<<L]  <->  x1,x2 P4 [Ch][h][w]
// Through every level a 4-channel group stays together (Cz % 4 == 0) and lands in latent channels
// c..c+3 at one pixel; the 2^L x-neighbours of a latent row come from 2^L different groups of the
// same state pixel.  A thread therefore moves 2^L groups (16-byte loads, coalesced across the warp),
// transposes them in registers and writes 2^L-wide vectors per latent channel: both sides coalesced.
// ------------------------------------------------------------------------------------------
template <int L, bool TO_LATENT>
__global__ void latent_spread_kernel(float4* __restrict__ x1, float4* __restrict__ x2, float* __restrict__ z, int Ch,
                                     int h, int w, int* __restrict__ status_clear) {
    if (status_clear && blockIdx.x == 0 && threadIdx.x == 0) *status_clear = 0;     // first kernel of a decode
    constexpr int S = 1 << L;                               // sub-positions per axis
    const int Gh = Ch / 4, Cz = (2 * Ch) >> (2 * L), Gz = Cz / 4;
    const int H = h << L, W = w << L;
    const size_t n = (size_t)h * w, total = (size_t)Gz * S * n, zplane = (size_t)H * W, splane = p4_plane_px(h, w);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const int yy = (int)((i / n) % S);                  // sub-row inside the S x S block of this state pixel
        const int cg = (int)(i / (n * S));                  // latent channel group
        float4 v[S];
        float4* src[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {                       // s = sub-column
            int cc = 4 * cg;
#pragma unroll
            for (int l = 0; l < L; ++l) {                   // level l consumes bit (L-1-l) of yy (row) and s (column)
                const int k = (((yy >> (L - 1 - l)) & 1) << 1) | ((s >> (L - 1 - l)) & 1);
                cc += k * ((2 * Ch) >> (2 * (l + 1)));
            }
            const int gg = cc >> 2;
            src[s] = (gg < Gh ? x1 + (size_t)gg * splane : x2 + (size_t)(gg - Gh) * splane);
        }
        float* zp = z + (size_t)(4 * cg) * zplane + (size_t)((y << L) + yy) * W + ((size_t)x << L);
        if (TO_LATENT) {
#pragma unroll
            for (int s = 0; s < S; ++s) v[s] = src[s][(size_t)(y + 1) * (w + 2) + x + 1];
            if constexpr (S == 4) {
                *reinterpret_cast<float4*>(zp) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
                *reinterpret_cast<float4*>(zp + zplane) = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
                *reinterpret_cast<float4*>(zp + 2 * zplane) = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
                *reinterpret_cast<float4*>(zp + 3 * zplane) = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
            } else {
                *reinterpret_cast<float2*>(zp) = make_float2(v[0].x, v[S - 1].x);
                *reinterpret_cast<float2*>(zp + zplane) = make_float2(v[0].y, v[S - 1].y);
                *reinterpret_cast<float2*>(zp + 2 * zplane) = make_float2(v[0].z, v[S - 1].z);
                *reinterpret_cast<float2*>(zp + 3 * zplane) = make_float2(v[0].w, v[S - 1].w);
            }
        } else {
            float r[4][S];
            if constexpr (S == 4) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(zp + e * zplane));
                    r[e][0] = t.x; r[e][1] = t.y; r[e][2] = t.z; r[e][3] = t.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(zp + e * zplane));
                    r[e][0] = t.x; r[e][S - 1] = t.y;
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) p4_store(src[s], h, w, y, x, make_float4(r[0][s], r[1][s], r[2][s], r[3][s]));
        }
    }
}

// any other depth (not used by the reference's two modes): one 4-channel unit per thread
template <bool TO_LATENT>
__global__ void latent_spread_generic_kernel(float4* __restrict__ x1, float4* __restrict__ x2, float* __restrict__ z,
                                             int Ch, int h, int w, int L, int* __restrict__ status_clear) {
    if (status_clear && blockIdx.x == 0 && threadIdx.x == 0) *status_clear = 0;
    const int Gh = Ch / 4;
    const size_t n = (size_t)h * w, total = (size_t)2 * Gh * n;
    const int H = h << L, W = w << L;
    const size_t zplane = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int gg = (int)(i / n);
        const size_t p = i - (size_t)gg * n;
        const int y = (int)(p / w), x = (int)(p - (size_t)y * w);
        int cc = 4 * gg, D = 2 * Ch, Y = y, X = x;
        for (int l = 0; l < L; ++l) {
            D >>= 2;
            const int k = cc / D;
            cc -= k * D;
            Y = 2 * Y + (k >> 1);
            X = 2 * X + (k & 1);
        }
        float4* sp = (gg < Gh ? x1 + (size_t)gg * p4_plane_px(h, w) : x2 + (size_t)(gg - Gh) * p4_plane_px(h, w));
        float* zp = z + (size_t)cc * zplane + (size_t)Y * W + X;
        if (TO_LATENT) {
            const float4 v = sp[(size_t)(y + 1) * (w + 2) + x + 1];
            zp[0] = v.x; zp[zplane] = v.y; zp[2 * zplane] = v.z; zp[3 * zplane] = v.w;
        } else {
            p4_store(sp, h, w, y, x, make_float4(__ldg(zp), __ldg(zp + zplane), __ldg(zp + 2 * zplane), __ldg(zp + 3 * zplane)));
        }
    }
}

template <bool TO_LATENT>
static int launch_spread_any(float* x1, float* x2, float* z, int Ch, int h, int w, int L, int* status_clear, cudaStream_t st) {
    const size_t units = (size_t)2 * Ch * h * w / 4;
    float4 *a = reinterpret_cast<float4*>(x1), *b = reinterpret_cast<float4*>(x2);
    ProfScope prof(st, TO_LATENT ? "latent_spread" : "latent_gather", 0.0, 32.0 * units);
    if (L == 2) latent_spread_kernel<2, TO_LATENT><<<ew_grid(units / 4), 256, 0, st>>>(a, b, z, Ch, h, w, status_clear);
    else if (L == 1) latent_spread_kernel<1, TO_LATENT><<<ew_grid(units / 2), 256, 0, st>>>(a, b, z, Ch, h, w, status_clear);
    else latent_spread_generic_kernel<TO_LATENT><<<ew_grid(units), 256, 0, st>>>(a, b, z, Ch, h, w, L, status_clear);
    return check_launch(TO_LATENT ? "latent_spread" : "latent_gather");
}

int launch_latent_spread(const float* x1, const float* x2, float* z, int Ch, int h, int w, int L, cudaStream_t st) {
    return launch_spread_any<true>(const_cast<float*>(x1), const_cast<float*>(x2), z, Ch, h, w, L, nullptr, st);
}
int launch_latent_gather(const float* z, float* x1, float* x2, int Ch, int h, int w, int L, int* status_clear,
                         cudaStream_t st) {
    return launch_spread_any<false>(x1, x2, const_cast<float*>(z), Ch, h, w, L, status_clear, st);
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

// label map resize, nearest neighbour with PIL's Image.NEAREST sampling: the reference's cWCT.resize
// (models/cWCT.py:191-197), which this fork leaves commented out at :72-73 — masked transfer then only works when
// the masks already have the latent resolution (never in artistic mode).  PIL walks each axis with an accumulated
// double, xo = 0.5 * s; index = (int)xo; xo += s  (s = src / dst); the index tables are built the same way (one
// thread per axis: the accumulation order is part of the result at exact-integer sample points).
__global__ void mask_resize_tables_kernel(int Hs, int Ws, int Hd, int Wd, int* __restrict__ iy, int* __restrict__ ix) {
    const int axis = threadIdx.x;                       // 0: rows, 1: columns
    if (axis > 1) return;
    const int ns = axis ? Ws : Hs, nd = axis ? Wd : Hd;
    int* t = axis ? ix : iy;
    const double sc = (double)ns / (double)nd;
    double xo = sc * 0.5;
    for (int k = 0; k < nd; ++k) {
        t[k] = min((int)xo, ns - 1);
        xo += sc;
    }
}
__global__ void mask_resize_nearest_kernel(const uint8_t* __restrict__ src, int Ws, uint8_t* __restrict__ dst, int Hd, int Wd,
                                           const int* __restrict__ iy, const int* __restrict__ ix) {
    const int n = Hd * Wd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / Wd, x = i - y * Wd;
        dst[i] = src[(size_t)iy[y] * Ws + ix[x]];
    }
}
extern "C" int vst_mask_resize_nearest(const uint8_t* src, int Hs, int Ws, uint8_t* dst, int Hd, int Wd, int* scratch,
                                       void* stream) {
    using namespace vst;
    VST_REQUIRE(src && dst && scratch && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, "vst_mask_resize_nearest: bad arguments");
    const int n = Hd * Wd;
    mask_resize_tables_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(Hs, Ws, Hd, Wd, scratch, scratch + Hd);
    if (check_launch("mask_resize_tables")) return 1;
    mask_resize_nearest_kernel<<<std::min(cdiv(n, 256), 4096), 256, 0, (cudaStream_t)stream>>>(src, Ws, dst, Hd, Wd, scratch,
                                                                                              scratch + Hd);
    return check_launch("mask_resize_nearest");
}

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

// ------------------------------------------------------------------------------------------
// Mask preparation on the device: label histogram, the label re-mapping of
// models/segmentation/SegReMapping.py:19-76 (self_remapping / cross_remapping) as a 256-entry look-up table built by
// one small CTA, and the colour -> label rule of utils/utils.py:105-137 (load_segment).  No host round trip: the
// reference's np.unique / per-label boolean masks / per-pixel Python loop become three streaming kernels.
// ------------------------------------------------------------------------------------------
namespace vst {

__global__ void seg_hist_kernel(const uint8_t* __restrict__ seg, long long n, unsigned int* __restrict__ counts) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;                                     // blockDim.x == 256
    __syncthreads();
    const long long n16 = n / 16;
    const uint4* s16 = reinterpret_cast<const uint4*>(seg);
    const bool aligned = ((uintptr_t)seg & 15) == 0;
    if (aligned) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
            const uint4 v = __ldg(s16 + i);
            const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b) atomicAdd(&h[(w[k] >> (8 * b)) & 255u], 1u);
        }
    }
    for (long long i = (aligned ? n16 * 16 : 0) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        atomicAdd(&h[seg[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

// mode 0: self_remapping  — a label whose pixel ratio is below min_ratio moves to the first label of its column of the
//         relation table that is present with a ratio >= min_ratio (SegReMapping.py:49-76)
// mode 1: cross_remapping — a content label absent from the style moves to the first label of its column that the
//         style has (SegReMapping.py:19-46).  counts_a: the map being re-labelled, counts_b: the style map (mode 1).
__global__ void seg_lut_kernel(const unsigned int* __restrict__ counts_a, const unsigned int* __restrict__ counts_b,
                               long long n_a, const int* __restrict__ mapping, int rows, int n_classes, float min_ratio,
                               int mode, uint8_t* __restrict__ lut) {
    const int l = threadIdx.x;                              // one thread per label, 256 threads
    int out = l;
    const unsigned int ca = counts_a[l];
    if (ca > 0 && l < n_classes) {
        if (mode == 0) {
            const float nf = (float)n_a;
            if ((float)ca / nf < min_ratio) {
                for (int j = 0; j < rows; ++j) {
                    const int nl = mapping[(size_t)j * n_classes + l];
                    if (nl >= 0 && nl < 256 && counts_a[nl] > 0 && (float)counts_a[nl] / nf >= min_ratio) { out = nl; break; }
                }
            }
        } else if (counts_b[l] == 0) {
            for (int j = 0; j < rows; ++j) {
                const int nl = mapping[(size_t)j * n_classes + l];
                if (nl >= 0 && nl < 256 && counts_b[nl] > 0) { out = nl; break; }
            }
        }
    }
    lut[l] = (uint8_t)out;
}

__global__ void seg_apply_lut_kernel(const uint8_t* __restrict__ seg, long long n, const uint8_t* __restrict__ lut_g,
                                     uint8_t* __restrict__ out) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = lut_g[threadIdx.x];                  // blockDim.x == 256
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = lut[seg[i]];
}

// nearest table colour in L1, the first strict minimum in the reference's dict order winning ties
// (utils/utils.py:106-137: the tie branch there always raises inside its try and leaves the first minimum in place)
__global__ void seg_labels_from_colors_kernel(const uint8_t* __restrict__ rgb, long long n, uint8_t* __restrict__ labels) {
    const int tab[9][4] = {{0, 0, 255, 3}, {0, 255, 0, 2}, {0, 0, 0, 0}, {255, 255, 255, 1}, {255, 0, 0, 4}, {255, 255, 0, 5},
                           {128, 128, 128, 6}, {0, 255, 255, 7}, {255, 0, 255, 8}};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        int best = 99999, lab = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int d = abs(r - tab[k][0]) + abs(g - tab[k][1]) + abs(b - tab[k][2]);
            if (d < best) { best = d; lab = tab[k][3]; }
        }
        labels[i] = (uint8_t)lab;
    }
}

static int seg_grid(long long n) { return (int)std::min<long long>((n + 4095) / 4096 + 1, (long long)num_sms() * 8); }

}  // namespace vst

extern "C" size_t vst_seg_scratch_bytes(void) { return 2 * 256 * sizeof(unsigned int) + 256; }

extern "C" int vst_seg_remap(const uint8_t* seg, long long n, const uint8_t* style_seg, long long n_style, const int* mapping,
                             int rows, int n_classes, float min_ratio, uint8_t* out, void* scratch, void* stream) {
    using namespace vst;
    VST_REQUIRE(seg && out && mapping && scratch && n > 0 && rows > 0 && n_classes > 0 && n_classes <= 256,
                "vst_seg_remap: bad arguments");
    VST_REQUIRE(((uintptr_t)scratch & 3) == 0, "vst_seg_remap: scratch must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* ca = (unsigned int*)scratch;
    unsigned int* cb = ca + 256;
    uint8_t* lut = (uint8_t*)(cb + 256);
    VST_CUDA_OK(cudaMemsetAsync(scratch, 0, 2 * 256 * sizeof(unsigned int), st));
    count_launch();
    seg_hist_kernel<<<seg_grid(n), 256, 0, st>>>(seg, n, ca);
    if (check_launch("seg_hist")) return 1;
    if (style_seg) {
        VST_REQUIRE(n_style > 0, "vst_seg_remap: empty style map");
        seg_hist_kernel<<<seg_grid(n_style), 256, 0, st>>>(style_seg, n_style, cb);
        if (check_launch("seg_hist")) return 1;
    }
    seg_lut_kernel<<<1, 256, 0, st>>>(ca, cb, n, mapping, rows, n_classes, min_ratio, style_seg ? 1 : 0, lut);
    if (check_launch("seg_lut")) return 1;
    seg_apply_lut_kernel<<<seg_grid(n), 256, 0, st>>>(seg, n, lut, out);
    return check_launch("seg_apply_lut");
}

extern "C" int vst_seg_labels_from_colors(const uint8_t* rgb_hwc, long long n, uint8_t* labels, void* stream) {
    using namespace vst;
    VST_REQUIRE(rgb_hwc && labels && n > 0, "vst_seg_labels_from_colors: bad arguments");
    seg_labels_from_colors_kernel<<<seg_grid(n), 256, 0, (cudaStream_t)stream>>>(rgb_hwc, n, labels);
    return check_launch("seg_labels_from_colors");
}
