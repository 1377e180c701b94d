// kernels.cuh — internal launch interfaces shared by the translation units of libvstb200.
#pragma once
#include <algorithm>
#include "common.cuh"

namespace vst {

// ------------------------------------------------------------------------------------------
// "P4" activation layout (every tensor inside the reversible network lives in it):
//     [C/4][H+2][W+2][4] fp32
// i.e. groups of 4 channels interleaved per pixel (one pixel of one group = one 16-byte unit)
// with a 1-pixel border that already holds the ReflectionPad2d(1) values (RevResNet.py:80,83,86).
// Whoever WRITES a tensor also writes its border (p4_store), so every consumer — in particular
// the TMA loader of the tcgen05 convolution — reads plain rectangular boxes with no edge cases,
// and the im2col window of tap (ky,kx) is a 16-byte-granular address shift.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t p4_plane_px(int H, int W) { return (size_t)(H + 2) * (size_t)(W + 2); }
// floats needed for a C-channel tensor (+ slack so that tile over-reads past the last row stay in bounds)
__host__ __device__ inline size_t p4_floats(int C, int H, int W) {
    return ((size_t)((C + 3) / 4) * p4_plane_px(H, W) + 1024) * 4;
}

#ifdef __CUDACC__
// the border positions that mirror interior pixel (y, x): row -1 <- row 1, row H <- row H-2, same for
// columns (and the corners).  Out of line on purpose: it runs for O(perimeter) pixels only and must
// not bloat the hot epilogues.
static __device__ __noinline__ void p4_border_fix(float4* plane, int H, int W, int y, int x, float4 v) {
    const int Wp = W + 2;
    const bool up = (y == 1), dn = (y == H - 2), lf = (x == 1), rt = (x == W - 2);
    float4* rows[3] = {plane + (size_t)(y + 1) * Wp, plane, plane + (size_t)(H + 1) * Wp};   // self, row -1, row H
    const bool rowon[3] = {true, up, dn};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!rowon[a]) continue;
        if (a) rows[a][x + 1] = v;
        if (lf) rows[a][0] = v;            // column -1
        if (rt) rows[a][W + 1] = v;        // column W
    }
}
__device__ __forceinline__ bool p4_is_edge(int H, int W, int y, int x) {
    return (y == 1) | (y == H - 2) | (x == 1) | (x == W - 2);
}
// store one 4-channel pixel of group plane `plane` (interior coords y in [0,H), x in [0,W)) and
// every border position that reflects onto it.
__device__ __forceinline__ void p4_store(float4* plane, int H, int W, int y, int x, float4 v) {
    plane[(size_t)(y + 1) * (W + 2) + (x + 1)] = v;
    if (p4_is_edge(H, W, y, x)) p4_border_fix(plane, H, W, y, x, v);
}
#endif

#define VST_HALF_SCALE 64.0f
// VST_STATUS_F16_RANGE (vstb200.h): some |VST_HALF_SCALE * x| rounded to +-inf in fp16 (|x| >= 1023.75)

enum ConvEpi {
    EPI_RELU = 0,       // out = relu(conv + bias)
    EPI_NONE = 1,       // out = conv + bias
    EPI_ADD = 2,        // out = res + (conv + bias)                     forward coupling  (RevResNet.py:103)
    EPI_SUB = 3,        // out = res - (conv + bias)                     inverse coupling  (RevResNet.py:110-111)
    EPI_ADD_SQZ = 4,    // out = squeeze(res) + (conv + bias)            stride-2 forward  (RevResNet.py:100-103)
    EPI_SUB_UNSQZ = 5,  // unsqueeze(out) = res - (conv + bias)          stride-2 inverse  (RevResNet.py:112-113)
};

struct ConvArgs {
    const float* in;    // P4 [Cin/4][Hin+2][Win+2][4]
    const float* w;     // packed weights (layout depends on the kernel)
    const float* bias;  // [CoutPad]
    const float* res;   // coupling operand, P4 (may alias out for EPI_ADD / EPI_SUB)
    float* out;         // P4
    int Cin, Cout, CoutPad, Hin, Win, Hout, Wout;
    int epi;
    int out_split;      // conv_tch only: write the result as split fp16 (hi planes, then lo planes; see below)
    int* status;        // device status word of the call (bit 0: fp16 operand range exceeded), may be null
};

// "H8" split-half layout of a bottleneck tensor that only feeds a tensor-core conv (precision f16x2):
//     hi [C/8][H+2][W+2][8 halfs]  followed by  lo [C/8][H+2][W+2][8 halfs],   x = hi + lo
// fp16 has only 5 exponent bits: for |x| < 0.1 the remainder x - hi (~ x * 2^-12) would fall into the fp16
// subnormals and lose its low bits.  The split therefore works on VST_HALF_SCALE * x (a power of two:
// exact), which keeps lo normal down to |x| ~ 4e-3 and still leaves |x| < 1000 representable; the
// consuming kernel multiplies its fp32 accumulators by 1 / VST_HALF_SCALE (exact) in the epilogue.
// One pixel of one 8-channel group is again a 16-byte unit, so a tensor in this layout is addressed
// exactly like a P4 tensor with C/8 "groups" (border included) and occupies the same number of bytes as
// the fp32 P4 tensor it replaces.

#ifdef __CUDACC__
// Fused epilogue shared by the FFMA and tcgen05 kernels, in two halves so that callers can issue
// the residual loads of many units before they consume any (memory-level parallelism):
//   conv_epi_res   — the coupling operand for output channels 4g..4g+3 at output pixel (y, x)
//   conv_epi_store — combine `v` (= conv + bias) with it and store (+ reflection border)
__device__ __forceinline__ float4 conv_epi_res(const ConvArgs& a, int g, int y, int x) {
    const float4* resp = reinterpret_cast<const float4*>(a.res);
    if (a.epi == EPI_ADD_SQZ) {   // res is the un-squeezed tensor [Cout/16][2Hout+2][2Wout+2][4]
        const int gq = a.Cout >> 4, k = g / gq, gs = g - k * gq;
        const int Hh = 2 * a.Hout, Wh = 2 * a.Wout;
        return resp[(size_t)gs * p4_plane_px(Hh, Wh) + (size_t)(2 * y + (k >> 1) + 1) * (Wh + 2) + (2 * x + (k & 1) + 1)];
    }
    return resp[(size_t)g * p4_plane_px(a.Hout, a.Wout) + (size_t)(y + 1) * (a.Wout + 2) + (x + 1)];
}
__device__ __forceinline__ void conv_epi_store(const ConvArgs& a, int g, int y, int x, float4 v, float4 r) {
    float4* outp = reinterpret_cast<float4*>(a.out);
    switch (a.epi) {
        case EPI_RELU:
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            break;
        case EPI_NONE: break;
        case EPI_ADD:
        case EPI_ADD_SQZ:
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            break;
        case EPI_SUB:
            v.x = r.x - v.x; v.y = r.y - v.y; v.z = r.z - v.z; v.w = r.w - v.w;
            break;
        case EPI_SUB_UNSQZ: { // out is the un-squeezed tensor; res is at the conv's own resolution
            const int gq = a.Cout >> 4, k = g / gq, gs = g - k * gq;
            const int Hh = 2 * a.Hout, Wh = 2 * a.Wout;
            v.x = r.x - v.x; v.y = r.y - v.y; v.z = r.z - v.z; v.w = r.w - v.w;
            p4_store(outp + (size_t)gs * p4_plane_px(Hh, Wh), Hh, Wh, 2 * y + (k >> 1), 2 * x + (k & 1), v);
            return;
        }
    }
    p4_store(outp + (size_t)g * p4_plane_px(a.Hout, a.Wout), a.Hout, a.Wout, y, x, v);
}
__device__ __forceinline__ void conv_epilogue(const ConvArgs& a, int g, int y, int x, float4 v) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.epi >= EPI_ADD) r = conv_epi_res(a, g, y, x);
    conv_epi_store(a, g, y, x, v, r);
}
#endif

int conv_cout_pad(int Cout);
int launch_pack_conv_weights(const float* w, const float* b, float* wp, float* bp, int Cin, int Cout, int CoutPad,
                             cudaStream_t st);
int launch_conv3x3_ffma(const ConvArgs& a, int stride, cudaStream_t st);

// fused reversible block for the full-resolution stage (16 -> 4 -> 4 -> 16 channels), block16.cu:
//   out = res +/- conv3(relu(conv2(relu(conv1(x)))))      weights in the conv3x3_ffma pack layout
struct Block16Args {
    const float* x;      // P4 [4 groups][H+2][W+2][4]   (the half-state F is evaluated on)
    const float* res;    // P4, coupling operand (may alias out)
    float* out;          // P4
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    int H, W;
    int sub;             // 0: out = res + F(x)   1: out = res - F(x)
    long long* trace;    // developer aid (VST_TC_TRACE="16,4"): clock64() stamps of one interior CTA
};
int launch_rev_block16(const Block16Args& a, cudaStream_t st);

// the same for C = 16 / 64 as ONE row-streaming tcgen05 kernel (precision f16x2), block_tc.cu
bool block_tc_eligible(int C, int mult);
size_t block_tc_pack_floats(int C);
int launch_pack_block_tc(int C, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                         const float* b3, float* pk, cudaStream_t st);
// mode 0: plain; 1: store the result squeezed (the block in front of a stride-2 block, forward); 2: read the coupling
// operand through the unsqueeze addressing (the block behind a stride-2 block, inverse) — block_tc.cu
int launch_rev_block_tc(int C, const float* x, const float* res, float* out, const float* wpack, int H, int W, int sub,
                        int* status, cudaStream_t st, int mode = 0);

// tensor-core (tcgen05) path, conv_tc.cu
bool tc_eligible(int Cin, int Cout, int stride);
int tc_tile_n(int Cout);
size_t tc_packed_floats(int Cin, int Cout, int N, int terms);
int launch_pack_tc_weights(const float* w, float* wp, int Cin, int Cout, int N, int terms, cudaStream_t st);
int launch_conv3x3_tc(const ConvArgs& a, int terms, cudaStream_t st);
// ... with fp16 split operands: a.in is an H8 tensor, weights packed by launch_pack_tch3_weights (2 terms)
bool tc_half_eligible(int Cin, int Cout, int stride);
int launch_pack_tc_half_weights(const float* w, float* wp, int Cin, int Cout, int N, cudaStream_t st);
bool conv_pair_eligible(const ConvArgs& a);
int launch_conv3x3_pair(const ConvArgs& a, cudaStream_t st);     // conv_pair.cu: 64 -> 256 on a CTA pair (cta_group::2)
int launch_conv3x3_tc_half(const ConvArgs& a, cudaStream_t st);
long long* tc_trace_buffer(int Cin, int Cout, cudaStream_t st);   // developer aid, conv_tc.cu

// kx-folded tensor-core path for the convs without a coupling operand (ReLU / plain epilogue), conv_tcx.cu
bool tcx_eligible(int Cin, int Cout, int stride);
int launch_pack_tcx_weights(const float* w, float* wp, int Cin, int Cout, int NC, int terms, cudaStream_t st);
int launch_conv3x3_tcx(const ConvArgs& a, int terms, cudaStream_t st);

// the same with half-precision split operands (kind::f16, K = 16 per UMMA), conv_tch.cu
bool tch_eligible(int Cin, int Cout, int stride);
int launch_pack_tch_weights(const float* w, float* wp, int Cin, int Cout, int NC, int terms, cudaStream_t st);
int launch_conv3x3_tch(const ConvArgs& a, int terms, cudaStream_t st);
// stride-2 conv [Cout][C][3][3] as a stride-1 kx-folded conv on the squeezed (4C-channel) input
int launch_pack_tch_s2_weights(const float* w, float* wp, int C, int Cout, cudaStream_t st);

// tensor-core Gram + sums of an unmasked feature map (cWCT statistics), gram_tc.cu
bool gram_tc_eligible(int C, long long n);
int launch_gram_tc(const float* feat, const float* pivot, double* count, double* sum, double* gram, int C, long long n,
                   cudaStream_t st);
// per-label statistics on the tensor cores (label map [H*W] uint8, strip-major traversal), gram_tc.cu
bool gram_tc_masked_eligible(int C, int H, int W);
int launch_gram_tc_masked(const float* feat, const uint8_t* labels, const float* pivot, double* count, double* sum, double* gram,
                          int C, int L, int H, int W, cudaStream_t st);
// ... of the latent that the P4 half-states x1 | x2 (Ch channels each, h x w) spread to, without materialising it
int launch_gram_tc_state(const float* x1, const float* x2, const float* pivot, double* count, double* sum, double* gram, int C,
                         int Ch, int h, int w, cudaStream_t st);

// cWCT on the network's own state (fused video path), cwct.cu
int launch_stats_state(const float* x1, const float* x2, int C, int Ch, int h, int w, void* stats, cudaStream_t st);
int launch_apply_state(float* x1, float* x2, int C, int Ch, int h, int w, const float* T, const float* mu, const float* beta,
                       const int* valid, cudaStream_t st);
int launch_factor(int mode, const void* content_stats, const void* const* style_stats, const float* alpha_s, int n_styles,
                  float alpha_c, float eps, int C, int n_labels, int masked, int use_double, float* T, float* mu, float* beta,
                  int* valid, int* status, cudaStream_t st);
size_t cwct_stats_bytes(int C, int n_labels);

// layout / rearrangement kernels, layout.cu  (all tensors P4 unless stated)
// NCHW -> P4; `add` (optional): a tiny P4 tensor [C0/4][add_h+2][add_w+2][4] whose centre pixel is added to every pixel
int launch_image_to_state(const float* x, float* s0, int Cimg, int C0, int H, int W, int* status_clear, const float* add,
                          int add_h, int add_w, cudaStream_t st);
int launch_state_to_image(const float* s0, float* x, int Cimg, int H, int W, cudaStream_t st);           // P4 -> NCHW
int launch_image_u8_to_state(const uint8_t* hwc, float* s0, int C0, int H, int W, int bgr, int* status_clear,
                             const float* add, int add_h, int add_w, cudaStream_t st);
int launch_state_to_image_u8(const float* s0, uint8_t* hwc, int H, int W, int bgr, cudaStream_t st);
int launch_space_to_depth(const float* in, float* out, int C, int Hin, int Win, cudaStream_t st);
int launch_p4_replicate_topleft(float* t, int C, int H, int W, cudaStream_t st);
int launch_depth_to_space(const float* in, float* out, int Cout, int Hin, int Win, cudaStream_t st);
int launch_latent_spread(const float* x1, const float* x2, float* z, int Ch, int h, int w, int L, cudaStream_t st);
int launch_latent_gather(const float* z, float* x1, float* x2, int Ch, int h, int w, int L, int* status_clear, cudaStream_t st);

}  // namespace vst
