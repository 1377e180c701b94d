// revnet.cu — host-side plan and dispatch of the reversible network (encode / decode).
//
// Replaces models/RevResNet.py:166-239 (RevResNet), :68-116 (residual_block), :119-163
// (channel_reduction).  The plan holds no device memory; every buffer is carved out of the
// caller's workspace, in the P4 layout (kernels.cuh).  State convention: the two half-states (s0,s1) live in two of three
// rotating planar buffers; an additive-coupling block updates one half IN PLACE
//     forward : s0 <- s0 + F(s1) ; swap       (RevResNet.py:96-104)
//     inverse : s1 <- s1 - F(s0) ; swap       (RevResNet.py:106-116)
// so split()/merge() (RevResNet.py:8-16) are pointer bookkeeping, not copies.
#include <vector>
#include <new>
#include <stdlib.h>
#include "kernels.cuh"

namespace vst {

struct ConvDesc {
    int Cin, Cout, CoutPad, stride;
    size_t raw_w, raw_b;  // float offsets into the flat raw parameter buffer
    size_t pk_w, pk_b;    // float offsets into the packed buffer
    bool tc;              // eligible for the tcgen05 implicit-GEMM kernel
    bool tcx;             // ... in its kx-folded form (no coupling operand: conv1 / conv2 of a block)
    bool tch;             // ... kx-folded with fp16 split operands (precision f16x2)
    bool s2tc;            // stride-2 conv runnable as a stride-1 tensor-core conv on the squeezed input (f16x2)
    size_t pk_s2;         // float offset of that conv's remapped fp16 weight pack
    size_t pk_tc;         // float offset of the tensor-core weight pack (sized for hi+lo terms)
};
struct BlockDesc {
    int channel, stride;
    ConvDesc conv[3];
    bool btc;             // whole block runnable as one row-streaming tensor-core kernel (block_tc.cu, f16x2)
    size_t pk_blk;        // float offset of that kernel's block pack (16-byte aligned)
};

}  // namespace vst

struct vst_revnet {
    vst_revnet_config cfg;
    std::vector<vst::BlockDesc> stack, cr;
    size_t raw_floats = 0, packed_floats = 0;
    int precision = VST_CONV_FP32;
    int c0 = 0;          // half-state channels at full resolution
    int c_last = 0;      // half-state channels after the stack
    int cr_channel = 0;  // channel_reduction block width (hidden_dim * 4^sp_steps)
    int down = 1;        // prod(strides)
};

namespace vst {

static void add_block(vst_revnet* n, std::vector<BlockDesc>& dst, int channel, int stride) {
    BlockDesc b;
    b.channel = channel;
    b.stride = stride;
    const int mid = channel / n->cfg.mult;
    const int in_ch = (stride == 1) ? channel : channel / 4;   // RevResNet.py:74-77
    const int cin[3] = {in_ch, mid, mid}, cout[3] = {mid, mid, channel}, st[3] = {stride, 1, 1};
    for (int k = 0; k < 3; ++k) {
        ConvDesc& c = b.conv[k];
        c.Cin = cin[k]; c.Cout = cout[k]; c.stride = st[k];
        c.CoutPad = conv_cout_pad(c.Cout);
        c.raw_w = n->raw_floats; n->raw_floats += (size_t)c.Cout * c.Cin * 9;
        c.raw_b = n->raw_floats; n->raw_floats += (size_t)c.Cout;
        c.pk_w = n->packed_floats; n->packed_floats += (size_t)c.Cin * 9 * c.CoutPad;
        c.pk_b = n->packed_floats; n->packed_floats += (size_t)c.CoutPad;
        c.tc = tc_eligible(c.Cin, c.Cout, c.stride);
        c.tcx = k < 2 && tcx_eligible(c.Cin, c.Cout, c.stride);
        c.tch = k < 2 && tch_eligible(c.Cin, c.Cout, c.stride);
        c.pk_tc = n->packed_floats;
        if (c.tc) n->packed_floats += tc_packed_floats(c.Cin, c.Cout, tc_tile_n(c.Cout), 3);
        c.s2tc = k == 0 && c.stride == 2 && tch_eligible(4 * c.Cin, c.Cout, 1);
        c.pk_s2 = n->packed_floats;
        if (c.s2tc) n->packed_floats += tc_packed_floats(4 * c.Cin, c.Cout, c.Cout, 1);
    }
    b.btc = stride == 1 && block_tc_eligible(channel, n->cfg.mult);
    b.pk_blk = align_up(n->packed_floats, 4);
    if (b.btc) n->packed_floats = b.pk_blk + block_tc_pack_floats(channel);
    dst.push_back(b);
}

// VST_BLOCK_TC (developer knob): bit 0 = fused tensor-core block for C = 16, bit 1 = for C = 64; default both
static int block_tc_mask() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VST_BLOCK_TC"); v = e ? atoi(e) : 3; }
    return v;
}
static bool block_fused_tc(const vst_revnet* n, const BlockDesc& b) {
    return n->precision == VST_CONV_F16X2 && b.btc && (block_tc_mask() & (b.channel == 16 ? 1 : 2));
}

struct Shape { int c, h, w; };

// VST_SQZ_FUSE=0 (developer knob) keeps the standalone space_to_depth / depth_to_space launches of the transitions

// VST_FOLD0=0 (developer knob) keeps the first block as an ordinary launch
static bool fold_first_block(const vst_revnet* n, int W) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VST_FOLD0"); v = e ? atoi(e) : 1; }
    return v && n->stack.size() >= 2 && n->stack[0].stride == 1 && n->stack[1].stride == 1 && W >= 4 &&
           block_fused_tc(n, n->stack[0]) && block_fused_tc(n, n->stack[1]);
}

struct Workspace {
    int* status;     // first 64 bytes of the workspace: [0] status bits of the last call (VST_STATUS_*),
                     // [1] cWCT Cholesky retries (-1: failed) and [2] cWCT validity of the last fused stylize call
    uint8_t* cstats; // cWCT scratch of the fused path: content statistics block (one label) ...
    float *T, *mu, *beta;   // ... and the transform
    float* tiny;     // two 8x8 P4 tensors of c0 channels (zeros | F(0) of the first block), see fold_first_block
    float* P[3];
    float* T1;
    float* T2;
};
constexpr size_t WS_HEADER_FLOATS = 16;
constexpr int TINY = 8;
static size_t tiny_floats(const vst_revnet* n) { return align_up(p4_floats(n->c0, TINY, TINY), 64); }
static size_t cwct_scratch_floats(const vst_revnet* n) {
    const int C = 2 * n->cfg.hidden_dim;
    return align_up(cwct_stats_bytes(C, 1) / 4 + (size_t)C * C + 2 * (size_t)C + 16, 64) + 2 * tiny_floats(n);
}

static size_t half_state_floats(const vst_revnet* n, int H, int W) {
    // largest half-state over all stages, in the P4 layout (border + slack included)
    size_t m = p4_floats(n->c0, H, W);
    int h = H, w = W;
    for (const BlockDesc& b : n->stack) {
        if (b.stride == 2) { h /= 2; w /= 2; }
        m = std::max(m, p4_floats(b.channel, h, w));
    }
    m = std::max(m, p4_floats(n->cr_channel, h, w));
    return align_up(m, 64);
}
static size_t temp_floats(const vst_revnet* n, int H, int W) {
    // largest bottleneck tensor: (channel/mult) x h x w over all blocks
    size_t m = 0;
    int h = H, w = W;
    for (const BlockDesc& b : n->stack) {
        if (b.stride == 2) { h /= 2; w /= 2; }
        m = std::max(m, p4_floats(b.channel / n->cfg.mult, h, w));
    }
    for (const BlockDesc& b : n->cr) m = std::max(m, p4_floats(b.channel / n->cfg.mult, h, w));
    return align_up(m, 64);
}

static int carve(const vst_revnet* n, int H, int W, void* ws, size_t ws_bytes, Workspace* out) {
    size_t hs = half_state_floats(n, H, W), ts = temp_floats(n, H, W);
    size_t need = (WS_HEADER_FLOATS + cwct_scratch_floats(n) + 3 * hs + 2 * ts) * sizeof(float);
    VST_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: have %zu bytes, need %zu", ws_bytes, need);
    VST_REQUIRE(((uintptr_t)ws & 15) == 0, "workspace must be 16-byte aligned");
    float* p = (float*)ws;
    out->status = (int*)p; p += WS_HEADER_FLOATS;
    {
        const int C = 2 * n->cfg.hidden_dim;
        float* q = p;
        out->cstats = (uint8_t*)q; q += cwct_stats_bytes(C, 1) / 4;
        out->T = q; q += (size_t)C * C;
        out->mu = q; q += C;
        out->beta = q;
        p += cwct_scratch_floats(n);
        out->tiny = p - 2 * tiny_floats(n);
    }
    for (int i = 0; i < 3; ++i) { out->P[i] = p; p += hs; }
    out->T1 = p; p += ts;
    out->T2 = p;
    return 0;
}

static ConvArgs conv_args(const ConvDesc& c, const float* packed, const float* in, int Hin, int Win, float* out,
                          const float* res, int epi, int* status) {
    ConvArgs a;
    a.in = in; a.w = packed + c.pk_w; a.bias = packed + c.pk_b; a.res = res; a.out = out;
    a.Cin = c.Cin; a.Cout = c.Cout; a.CoutPad = c.CoutPad;
    a.Hin = Hin; a.Win = Win; a.Hout = Hin / c.stride; a.Wout = Win / c.stride;
    a.epi = epi;
    a.out_split = 0;
    a.status = status;
    return a;
}

// precision f16x2: conv2 (kx-folded fp16 kernel) hands its result to conv3 as an H8 split-half tensor, and
// conv3 runs its kind::f16 variant on it (no converter stage, half the weight bytes)
static bool block_split_t2(const vst_revnet* n, const BlockDesc& b) {
    return n->precision == VST_CONV_F16X2 && b.conv[1].tch && b.conv[2].tc && tc_half_eligible(b.conv[2].Cin, b.conv[2].Cout, 1);
}

static int tc_terms(int precision) {
    return precision == VST_CONV_TF32 ? 1 : (precision == VST_CONV_TF32X2 || precision == VST_CONV_F16X2) ? 2
           : precision == VST_CONV_TF32X3 ? 3 : 0;
}

static int run_conv(const vst_revnet* n, const ConvDesc& c, const float* packed, const float* in, int Hin, int Win,
                    float* out, const float* res, int epi, int* status, cudaStream_t st, bool out_split = false,
                    bool in_split = false) {
    ConvArgs a = conv_args(c, packed, in, Hin, Win, out, res, epi, status);
    const int terms = tc_terms(n->precision);
    a.out_split = out_split ? 1 : 0;
    if (in_split) {
        a.w = packed + c.pk_tc;
        return launch_conv3x3_tc_half(a, st);
    }
    if (n->precision == VST_CONV_F16X2 && c.tch && (epi == EPI_RELU || epi == EPI_NONE)) {
        a.w = packed + c.pk_tc;
        return launch_conv3x3_tch(a, terms, st);
    }
    if (terms > 0 && c.tcx && (epi == EPI_RELU || epi == EPI_NONE)) {
        a.w = packed + c.pk_tc;
        return launch_conv3x3_tcx(a, terms, st);
    }
    if (terms > 0 && c.tc) {
        a.w = packed + c.pk_tc;
        return launch_conv3x3_tc(a, terms, st);
    }
    return launch_conv3x3_ffma(a, c.stride, st);
}

// F(x) = conv3(relu(conv2(relu(conv1(x)))))  with the coupling fused into conv3's epilogue
static bool block_s2tc(const vst_revnet* n, const BlockDesc& b) {
    return n->precision == VST_CONV_F16X2 && b.stride == 2 && b.conv[0].s2tc;
}

// x_sq: for stride-2 blocks in f16x2 mode, squeeze(x) with its top/left border replicated (see conv_tch.cu)
// btc_mode: layout mode of the fused tensor-core block (block_tc.cu): 0 plain, 1 squeezed output, 2 squeezed coupling operand
static int run_F(const vst_revnet* n, const BlockDesc& b, const float* packed, const float* x, int Hin, int Win,
                 const Workspace& ws, const float* res, float* out, int epi, cudaStream_t st,
                 const float* x_sq = nullptr, int btc_mode = 0) {
    const int Ho = Hin / b.stride, Wo = Win / b.stride;
    if (block_fused_tc(n, b) && (epi == EPI_ADD || epi == EPI_SUB) && Win >= 4)
        return launch_rev_block_tc(b.channel, x, res, out, packed + b.pk_blk, Hin, Win, epi == EPI_SUB ? 1 : 0, ws.status, st,
                                   btc_mode);
    VST_REQUIRE(btc_mode == 0, "internal: squeeze modes need the fused tensor-core block");
    if (b.stride == 1 && b.conv[0].Cin == 16 && b.conv[0].Cout == 4 && b.conv[2].Cout == 16 &&
        (epi == EPI_ADD || epi == EPI_SUB)) {
        // full-resolution stage: the whole block in one fused CUDA-core kernel (block16.cu)
        Block16Args a;
        a.x = x; a.res = res; a.out = out;
        a.w1 = packed + b.conv[0].pk_w; a.b1 = packed + b.conv[0].pk_b;
        a.w2 = packed + b.conv[1].pk_w; a.b2 = packed + b.conv[1].pk_b;
        a.w3 = packed + b.conv[2].pk_w; a.b3 = packed + b.conv[2].pk_b;
        a.H = Hin; a.W = Win; a.sub = (epi == EPI_SUB) ? 1 : 0; a.trace = nullptr;
        return launch_rev_block16(a, st);
    }
    if (x_sq && block_s2tc(n, b)) {
        const ConvDesc& c = b.conv[0];
        ConvArgs a = conv_args(c, packed, x_sq, Ho, Wo, ws.T1, nullptr, EPI_RELU, ws.status);
        a.Cin = 4 * c.Cin; a.Hout = Ho; a.Wout = Wo;           // stride-1 conv on the squeezed tensor
        a.w = packed + c.pk_s2;
        if (launch_conv3x3_tch(a, 2, st)) return 1;
    } else if (run_conv(n, b.conv[0], packed, x, Hin, Win, ws.T1, nullptr, EPI_RELU, ws.status, st)) return 1;
    const bool split = block_split_t2(n, b);
    if (run_conv(n, b.conv[1], packed, ws.T1, Ho, Wo, ws.T2, nullptr, EPI_RELU, ws.status, st, split, false)) return 1;
    if (run_conv(n, b.conv[2], packed, ws.T2, Ho, Wo, out, res, epi, ws.status, st, false, split)) return 1;
    return 0;
}

static bool sqz_fuse_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VST_SQZ_FUSE"); v = e ? atoi(e) : 1; }
    return v != 0;
}
// the stride-2 block `s2` has a fused tensor-core stride-1 neighbour `nb` working at the un-squeezed resolution h x w:
// that neighbour can write (forward) / read (inverse) the squeezed layout itself
static bool sqz_fusable(const vst_revnet* n, const BlockDesc& s2, const BlockDesc& nb, int h, int w) {
    return sqz_fuse_enabled() && block_s2tc(n, s2) && nb.stride == 1 && block_fused_tc(n, nb) && h % 2 == 0 && w % 2 == 0 &&
           h >= 4 && w >= 4;
}

// the two half-states of the network between encode and decode (what channel_reduction's spread loops turn into z)
struct StatePair { float *s0, *s1, *spare; int h, w; };

// frame source / sink of a pass: fp32 NCHW (the reference API) or uint8 HWC (the video frame format)
struct FrameIO { const float* f32_in; const uint8_t* u8_in; float* f32_out; uint8_t* u8_out; int bgr; };

static int encode_state(const vst_revnet* n, const float* packed, const FrameIO& io, int H, int W, const Workspace& ws,
                        bool first, cudaStream_t st, StatePair* sp) {
    float *s0 = ws.P[0], *s1 = ws.P[1], *spare = ws.P[2];
    // injective_pad + split (RevResNet.py:24-28, :8-12): s0 = [x, 0...], s1 = 0.
    // fold_first_block: the first block evaluates F on s1 = 0, and F(0) is the same vector at every pixel (zeros reflect
    // to zeros).  It is evaluated once with the block's own kernel on an 8x8 zero tile (bit-identical arithmetic), added
    // by the image -> state kernel, and the second block runs without coupling operand: the 133 MB memset, one full
    // block launch and one coupling-operand read per encode disappear (block 0: (x, 0) -> (0, x + F(0));
    // block 1: (0, y) -> (y, F(y)), RevResNet.py:96-104).
    const bool fold = fold_first_block(n, W);
    const float* add = nullptr;
    if (fold) {
        const size_t tf = tiny_floats(n);
        float *tz = ws.tiny, *tout = ws.tiny + tf;
        VST_CUDA_OK(cudaMemsetAsync(tz, 0, tf * sizeof(float), st));
        count_launch(1);
        if (run_F(n, n->stack[0], packed, tz, TINY, TINY, ws, tz, tout, EPI_ADD, st)) return 1;
        add = tout;
        std::swap(s0, s1);                           // the image (+ F(0)) is the "s1" the second block evaluates F on
    }
    float* img = fold ? s1 : s0;
    if (io.u8_in) {
        VST_REQUIRE(n->cfg.in_channel == 3, "uint8 frames have 3 channels");
        if (launch_image_u8_to_state(io.u8_in, img, n->c0, H, W, io.bgr, first ? ws.status : nullptr, add, TINY, TINY, st)) return 1;
    } else if (launch_image_to_state(io.f32_in, img, n->cfg.in_channel, n->c0, H, W, first ? ws.status : nullptr, add, TINY,
                                     TINY, st)) return 1;
    size_t first_block = 0;
    if (fold) {
        if (run_F(n, n->stack[1], packed, s1, H, W, ws, nullptr, s0, EPI_ADD, st)) return 1;     // s0 <- 0 + F(s1)
        std::swap(s0, s1);
        first_block = 2;
    } else {
        VST_CUDA_OK(cudaMemsetAsync(s1, 0, p4_floats(n->c0, H, W) * sizeof(float), st));
        count_launch(1);
    }

    int c = n->c0, h = H, w = W;
    bool s1_squeezed = false;                       // s1 already holds squeeze(s1) (written so by the previous block)
    for (size_t bi = first_block; bi < n->stack.size(); ++bi) {
        const BlockDesc& b = n->stack[bi];
        if (b.stride == 1) {
            if (bi + 1 < n->stack.size() && n->stack[bi + 1].stride == 2 && sqz_fusable(n, n->stack[bi + 1], b, h, w) &&
                bi >= first_block) {
                // the block in front of a transition stores y = s0 + F(s1) SQUEEZED into the spare buffer (it cannot be in
                // place: the layouts differ); s0's old buffer becomes the spare one
                if (run_F(n, b, packed, s1, h, w, ws, s0, spare, EPI_ADD, st, nullptr, 1)) return 1;
                float* old_s0 = s0;
                s0 = s1; s1 = spare; spare = old_s0;       // (s0, s1) = (x2, squeeze(y))
                s1_squeezed = true;
                continue;
            }
            if (run_F(n, b, packed, s1, h, w, ws, s0, s0, EPI_ADD, st)) return 1;
            std::swap(s0, s1);
        } else {
            if (s1_squeezed) {
                // s1 is already squeeze(x2) with the border the stride-2 conv needs: y1 = F(x2) + squeeze(s0) -> spare
                if (run_F(n, b, packed, nullptr, h, w, ws, s0, spare, EPI_ADD_SQZ, st, s1)) return 1;
                float* old_s0 = s0;
                s0 = s1; s1 = spare; spare = old_s0;       // (s0, s1) = (squeeze(x2), y1)
                s1_squeezed = false;
            } else if (block_s2tc(n, b)) {
                // new x1 = squeeze(s1) -> spare first: the stride-2 conv then runs on it (tensor cores);
                // y1 = F(s1) + squeeze(s0) -> the now dead s1 buffer
                if (launch_space_to_depth(s1, spare, c, h, w, st)) return 1;
                if (launch_p4_replicate_topleft(spare, 4 * c, h / 2, w / 2, st)) return 1;
                if (run_F(n, b, packed, s1, h, w, ws, s0, s1, EPI_ADD_SQZ, st, spare)) return 1;
                float* old_s0 = s0;
                s0 = spare; spare = old_s0;  // (s0, s1) = (squeeze(x2), y1)
            } else {
                // y1 = F(s1) + squeeze(s0) -> spare ; new x1 = squeeze(s1) -> old s0 buffer
                if (run_F(n, b, packed, s1, h, w, ws, s0, spare, EPI_ADD_SQZ, st)) return 1;
                if (launch_space_to_depth(s1, s0, c, h, w, st)) return 1;
                float* old_s1 = s1;
                s1 = spare; spare = old_s1;      // (s0, s1) = (squeeze(x2), y1)
            }
            c *= 4; h /= 2; w /= 2;
        }
    }
    if (n->cr_channel > c) {   // channel_reduction's injective pad (RevResNet.py:133-134): zero groups c/4..
        const size_t off = (size_t)(c / 4) * p4_plane_px(h, w) * 4;
        const size_t cnt = (size_t)((n->cr_channel - c) / 4) * p4_plane_px(h, w) * 4;
        VST_CUDA_OK(cudaMemsetAsync(s0 + off, 0, cnt * sizeof(float), st));
        VST_CUDA_OK(cudaMemsetAsync(s1 + off, 0, cnt * sizeof(float), st));
        count_launch(2);
    }
    for (const BlockDesc& b : n->cr) {
        if (run_F(n, b, packed, s1, h, w, ws, s0, s0, EPI_ADD, st)) return 1;
        std::swap(s0, s1);
    }
    sp->s0 = s0; sp->s1 = s1; sp->spare = spare; sp->h = h; sp->w = w;
    return 0;
}

static int forward_one(const vst_revnet* n, const float* packed, const float* x, float* z, int H, int W,
                       const Workspace& ws, bool first, cudaStream_t st) {
    StatePair sp;
    FrameIO io = {x, nullptr, nullptr, nullptr, 0};
    if (encode_state(n, packed, io, H, W, ws, first, st, &sp)) return 1;
    return launch_latent_spread(sp.s0, sp.s1, z, n->cr_channel, sp.h, sp.w, n->cfg.sp_steps, st);
}

static int decode_state(const vst_revnet* n, const float* packed, StatePair sp, const FrameIO& io, int H, int W,
                        const Workspace& ws, cudaStream_t st) {
    float *s0 = sp.s0, *s1 = sp.s1, *spare = sp.spare;
    int h = sp.h, w = sp.w;
    for (int i = (int)n->cr.size() - 1; i >= 0; --i) {
        if (run_F(n, n->cr[i], packed, s0, h, w, ws, s1, s1, EPI_SUB, st)) return 1;
        std::swap(s0, s1);
    }
    int c = n->c_last;   // channel_reduction's pad channels are dropped by simply ignoring them
    bool s1_squeezed = false;                // s1 still holds squeeze(x2): the next block reads it through the unsqueeze addressing
    for (int i = (int)n->stack.size() - 1; i >= 0; --i) {
        const BlockDesc& b = n->stack[i];
        if (b.stride == 1) {
            if (s1_squeezed) {
                // x1' = unsqueeze(s1) - F(s0) -> spare (not in place: the layouts differ); the squeezed buffer is then free
                if (run_F(n, b, packed, s0, h, w, ws, s1, spare, EPI_SUB, st, nullptr, 2)) return 1;
                float* old_s1 = s1;
                s1 = s0; s0 = spare; spare = old_s1;       // swap included: (s0, s1) = (x1', old s0)
                s1_squeezed = false;
                continue;
            }
            if (run_F(n, b, packed, s0, h, w, ws, s1, s1, EPI_SUB, st)) return 1;
            std::swap(s0, s1);
        } else {
            const bool sq = block_s2tc(n, b);          // F's stride-2 conv reads the squeezed x2 (= s0) directly
            if (sq && i > 0 && sqz_fusable(n, b, n->stack[i - 1], 2 * h, 2 * w)) {
                // x2 stays squeezed in s0's buffer (the next block un-squeezes it on the fly);
                // x1 = unsqueeze(s1 - F(x2)) -> spare
                if (launch_p4_replicate_topleft(s0, c, h, w, st)) return 1;
                if (run_F(n, b, packed, nullptr, 2 * h, 2 * w, ws, s1, spare, EPI_SUB_UNSQZ, st, s0)) return 1;
                float* old_s1 = s1;
                s1 = s0; s0 = spare; spare = old_s1;       // (s0, s1) = (x1, squeeze(x2))
                s1_squeezed = true;
            } else {
                // x2 = unsqueeze(s0) -> spare ; x1 = unsqueeze(s1 - F(x2)) -> old s0 buffer
                if (launch_depth_to_space(s0, spare, c / 4, h, w, st)) return 1;
                if (sq && launch_p4_replicate_topleft(s0, c, h, w, st)) return 1;
                if (run_F(n, b, packed, spare, 2 * h, 2 * w, ws, s1, s0, EPI_SUB_UNSQZ, st, sq ? s0 : nullptr)) return 1;
                float* old_s1 = s1;
                s1 = spare; spare = old_s1;      // (s0, s1) = (x1, x2)
            }
            c /= 4; h *= 2; w *= 2;
        }
    }
    // merge + injective_pad.inverse (RevResNet.py:30-31): keep the first in_channel channels of x1
    if (io.u8_out) return launch_state_to_image_u8(s0, io.u8_out, H, W, io.bgr, st);
    return launch_state_to_image(s0, io.f32_out, n->cfg.in_channel, H, W, st);
}

static int inverse_one(const vst_revnet* n, const float* packed, const float* z, float* x, int H, int W,
                       const Workspace& ws, bool first, cudaStream_t st) {
    StatePair sp = {ws.P[0], ws.P[1], ws.P[2], H / n->down, W / n->down};
    if (launch_latent_gather(z, sp.s0, sp.s1, n->cr_channel, sp.h, sp.w, n->cfg.sp_steps, first ? ws.status : nullptr, st))
        return 1;
    FrameIO io = {nullptr, nullptr, x, nullptr, 0};
    return decode_state(n, packed, sp, io, H, W, ws, st);
}

__global__ void stylize_status_kernel(const int* __restrict__ valid_status, int* __restrict__ ws_status) {
    ws_status[1] = valid_status[1];      // Cholesky retries (-1: failed)
    ws_status[2] = valid_status[0];      // validity
}

}  // namespace vst

using namespace vst;

extern "C" int vst_revnet_create(const vst_revnet_config* cfg, vst_revnet** out) {
    VST_REQUIRE(cfg && out, "vst_revnet_create: null argument");
    VST_REQUIRE(cfg->n_stages >= 1 && cfg->n_stages <= VST_MAX_STAGES, "n_stages %d out of range", cfg->n_stages);
    VST_REQUIRE(cfg->mult >= 1 && cfg->sp_steps >= 0 && cfg->hidden_dim >= 1 && cfg->n_cr_blocks >= 0,
                "bad mult/sp_steps/hidden_dim/n_cr_blocks");
    vst_revnet* n = new (std::nothrow) vst_revnet();
    VST_REQUIRE(n, "out of host memory");
    n->cfg = *cfg;
    n->c0 = cfg->n_channels[0];
    int c = n->c0;
    for (int s = 0; s < cfg->n_stages; ++s) {
        const int ch = cfg->n_channels[s], st = cfg->n_strides[s];
        bool ok = (st == 1 || st == 2) && cfg->n_blocks[s] >= 1 && ch % (4 * cfg->mult) == 0 &&
                  ((st == 1 && ch == c) || (st == 2 && ch == 4 * c));
        if (!ok) {
            delete n;
            set_error("stage %d: channels %d / stride %d inconsistent with incoming half-state width %d", s, ch, st, c);
            return 2;
        }
        for (int i = 0; i < cfg->n_blocks[s]; ++i) add_block(n, n->stack, ch, i == 0 ? st : 1);
        n->down *= st;
        c = ch;
    }
    n->c_last = c;
    n->cr_channel = cfg->hidden_dim;
    for (int i = 0; i < cfg->sp_steps; ++i) n->cr_channel *= 4;
    if (cfg->in_channel < 1 || cfg->in_channel > n->c0 || n->c0 % 4 != 0 || n->cr_channel < c ||
        n->cr_channel % (4 * cfg->mult) != 0 || ((2 * n->cr_channel) >> (2 * cfg->sp_steps)) % 4 != 0 ||
        (2 * n->cr_channel) % (1 << (2 * cfg->sp_steps)) != 0) {
        delete n;
        set_error("unsupported in_channel=%d / hidden_dim=%d / sp_steps=%d for last width %d", cfg->in_channel,
                  cfg->hidden_dim, cfg->sp_steps, c);
        return 2;
    }
    for (int i = 0; i < cfg->n_cr_blocks; ++i) add_block(n, n->cr, n->cr_channel, 1);
    *out = n;
    return 0;
}

extern "C" void vst_revnet_destroy(vst_revnet* net) { delete net; }

extern "C" int vst_revnet_set_precision(vst_revnet* net, int mode) {
    VST_REQUIRE(net, "null net");
    VST_REQUIRE(mode >= VST_CONV_FP32 && mode <= VST_CONV_F16X2, "unknown precision mode %d", mode);
    net->precision = mode;
    return 0;
}
extern "C" int vst_revnet_latent_channels(const vst_revnet* net) { return net ? 2 * net->cfg.hidden_dim : -1; }
extern "C" int vst_revnet_down_scale(const vst_revnet* net) { return net ? net->down : -1; }
extern "C" size_t vst_revnet_param_floats(const vst_revnet* net) { return net ? net->raw_floats : 0; }
extern "C" size_t vst_revnet_packed_bytes(const vst_revnet* net) { return net ? net->packed_floats * sizeof(float) : 0; }

extern "C" int vst_revnet_pack_weights(const vst_revnet* net, const float* raw, void* packed, void* stream) {
    VST_REQUIRE(net && raw && packed, "vst_revnet_pack_weights: null argument");
    VST_REQUIRE(((uintptr_t)packed & 15) == 0, "packed buffer must be 16-byte aligned");
    float* pk = (float*)packed;
    for (const std::vector<BlockDesc>* lst : {&net->stack, &net->cr})
        for (const BlockDesc& b : *lst) {
            if (net->precision == VST_CONV_F16X2 && b.btc &&
                launch_pack_block_tc(b.channel, raw + b.conv[0].raw_w, raw + b.conv[0].raw_b, raw + b.conv[1].raw_w,
                                     raw + b.conv[1].raw_b, raw + b.conv[2].raw_w, raw + b.conv[2].raw_b, pk + b.pk_blk,
                                     (cudaStream_t)stream))
                return 1;
            for (int k = 0; k < 3; ++k) {
                const ConvDesc& c = b.conv[k];
                if (launch_pack_conv_weights(raw + c.raw_w, raw + c.raw_b, pk + c.pk_w, pk + c.pk_b, c.Cin, c.Cout,
                                             c.CoutPad, (cudaStream_t)stream))
                    return 1;
                const int terms = tc_terms(net->precision);
                if (net->precision == VST_CONV_F16X2 && c.s2tc &&
                    launch_pack_tch_s2_weights(raw + c.raw_w, pk + c.pk_s2, c.Cin, c.Cout, (cudaStream_t)stream))
                    return 1;
                if (k == 2 && block_split_t2(net, b)) {
                    if (launch_pack_tc_half_weights(raw + c.raw_w, pk + c.pk_tc, c.Cin, c.Cout, tc_tile_n(c.Cout),
                                                    (cudaStream_t)stream))
                        return 1;
                } else if (net->precision == VST_CONV_F16X2 && c.tch) {
                    if (launch_pack_tch_weights(raw + c.raw_w, pk + c.pk_tc, c.Cin, c.Cout, c.Cout, terms,
                                                (cudaStream_t)stream))
                        return 1;
                } else if (terms > 0 && c.tcx) {
                    if (launch_pack_tcx_weights(raw + c.raw_w, pk + c.pk_tc, c.Cin, c.Cout, c.Cout, terms,
                                                (cudaStream_t)stream))
                        return 1;
                } else if (terms > 0 && c.tc &&
                           launch_pack_tc_weights(raw + c.raw_w, pk + c.pk_tc, c.Cin, c.Cout, tc_tile_n(c.Cout), terms,
                                                  (cudaStream_t)stream))
                    return 1;
            }
        }
    return 0;
}

extern "C" size_t vst_revnet_workspace_bytes(const vst_revnet* net, int B, int H, int W) {
    (void)B;
    if (!net || H <= 0 || W <= 0) return 0;
    return (WS_HEADER_FLOATS + cwct_scratch_floats(net) + 3 * half_state_floats(net, H, W) + 2 * temp_floats(net, H, W)) * sizeof(float);
}

static int check_hw(const vst_revnet* net, int B, int H, int W) {
    VST_REQUIRE(B >= 1, "batch must be >= 1");
    VST_REQUIRE(H > 0 && W > 0 && H % net->down == 0 && W % net->down == 0,
                "H, W must be positive multiples of down_scale=%d (got %dx%d)", net->down, H, W);
    VST_REQUIRE(H / net->down >= 2 && W / net->down >= 2, "image too small for reflection padding at 1/%d resolution",
                net->down);
    return 0;
}

extern "C" int vst_revnet_forward(const vst_revnet* net, const void* packed, const float* x, float* z, int B, int H,
                                  int W, void* workspace, size_t workspace_bytes, void* stream) {
    VST_REQUIRE(net && packed && x && z, "vst_revnet_forward: null argument");
    if (int r = check_hw(net, B, H, W)) return r;
    Workspace ws;
    if (int r = carve(net, H, W, workspace, workspace_bytes, &ws)) return r;
    const size_t xs = (size_t)net->cfg.in_channel * H * W;
    const size_t zs = (size_t)2 * net->cr_channel * (H / net->down) * (W / net->down);
    for (int b = 0; b < B; ++b)
        if (forward_one(net, (const float*)packed, x + b * xs, z + b * zs, H, W, ws, b == 0, (cudaStream_t)stream)) return 1;
    return 0;
}

static bool stylize_fused_ok(const vst_revnet* net, int H, int W) {
    const int C = 2 * net->cfg.hidden_dim;
    return (C == 32 || C == 128) && net->cr_channel % 128 == 0 && W / net->down >= 32 && net->cfg.in_channel == 3;
}

extern "C" int vst_revnet_stylize_supported(const vst_revnet* net, int H, int W) {
    return net && H > 0 && W > 0 && H % net->down == 0 && W % net->down == 0 && stylize_fused_ok(net, H, W) ? 1 : 0;
}

extern "C" int vst_revnet_stylize(const vst_revnet* net, const void* packed, const void* frame_in, void* frame_out, int io_u8,
                                  int bgr, int H, int W, const void* style_stats, float alpha_c, float eps, int use_double,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    VST_REQUIRE(net && packed && frame_in && frame_out && style_stats, "vst_revnet_stylize: null argument");
    if (int r = check_hw(net, 1, H, W)) return r;
    VST_REQUIRE(stylize_fused_ok(net, H, W), "vst_revnet_stylize: configuration not supported by the fused path "
                                             "(latent channels 32 or 128, latent width >= 32)");
    Workspace ws;
    if (int r = carve(net, H, W, workspace, workspace_bytes, &ws)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const float* pk = (const float*)packed;
    const int C = 2 * net->cfg.hidden_dim;
    FrameIO io = {io_u8 ? nullptr : (const float*)frame_in, io_u8 ? (const uint8_t*)frame_in : nullptr,
                  io_u8 ? nullptr : (float*)frame_out, io_u8 ? (uint8_t*)frame_out : nullptr, bgr};
    StatePair sp;
    if (encode_state(net, pk, io, H, W, ws, true, st, &sp)) return 1;      // (its first kernel clears status word 0)
    // cWCT on the state: statistics (tensor-core Gram over the P4 half-states), factor, in-place apply
    if (launch_stats_state(sp.s0, sp.s1, C, net->cr_channel, sp.h, sp.w, ws.cstats, st)) return 1;
    const void* ss[1] = {style_stats};
    const float one[1] = {1.f};
    int* vs = ws.status + 4;             // [4] valid, [5] status of the factor kernel
    if (launch_factor(0, ws.cstats, ss, one, 1, alpha_c, eps, C, 1, 0, use_double, ws.T, ws.mu, ws.beta, vs, vs + 1, st))
        return 1;
    stylize_status_kernel<<<1, 1, 0, st>>>(vs, ws.status);
    if (check_launch("stylize_status")) return 1;
    if (launch_apply_state(sp.s0, sp.s1, C, net->cr_channel, sp.h, sp.w, ws.T, ws.mu, ws.beta, vs, st)) return 1;
    return decode_state(net, pk, sp, io, H, W, ws, st);
}

extern "C" int vst_revnet_inverse(const vst_revnet* net, const void* packed, const float* z, float* x, int B, int H,
                                  int W, void* workspace, size_t workspace_bytes, void* stream) {
    VST_REQUIRE(net && packed && x && z, "vst_revnet_inverse: null argument");
    if (int r = check_hw(net, B, H, W)) return r;
    Workspace ws;
    if (int r = carve(net, H, W, workspace, workspace_bytes, &ws)) return r;
    const size_t xs = (size_t)net->cfg.in_channel * H * W;
    const size_t zs = (size_t)2 * net->cr_channel * (H / net->down) * (W / net->down);
    for (int b = 0; b < B; ++b)
        if (inverse_one(net, (const float*)packed, z + b * zs, x + b * xs, H, W, ws, b == 0, (cudaStream_t)stream)) return 1;
    return 0;
}
