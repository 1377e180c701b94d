/*
 * vstb200.h — C ABI of libvstb200.so: the B200-native (sm_100a) CAP-VSTNet stylization hot path.
 *
 * The reference (delldu/VSTNet) is pure eager PyTorch and has no native/FFI interface for this
 * path; the "interface each entry point replaces" is therefore the reference's Python call site
 * (file:line cited per function, relative to the reference root).  Host code (Python, ctypes —
 * see INTEGRATION.md) passes raw device pointers, integer shapes and a CUDA stream handle.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; vst_last_error() then returns a
 *    thread-local, human-readable message.  No C++ exceptions cross this boundary.
 *  - all device memory is owned by the caller (torch); the library never allocates or frees device
 *    memory and never synchronises the device.  Work is enqueued on `stream` (a cudaStream_t /
 *    CUstream passed as void*; NULL = legacy default stream).
 *  - feature maps are fp32, NCHW, contiguous.  Masks are uint8 [H*W] label maps, labels 0..254.
 */
#ifndef VSTB200_H
#define VSTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VST_MAX_STAGES 8
#define VST_MAX_LABELS 256 /* uint8 labels; the reference sizes its table max(label)+1 (cWCT.py:173-175) */
#define VST_MAX_STYLES 16

/* -------------------------------------------------------------------------------------------
 * Library
 * ----------------------------------------------------------------------------------------- */
const char* vst_last_error(void);
int vst_version(void);
/* number of device kernels this library has launched in the calling process (for bench.py's
 * "gpu_launches" and for tests that prove the CUDA path ran). */
unsigned long long vst_launch_count(void);

/* Optional per-launch device timing (CUDA events on the launch stream), aggregated by kernel
 * class.  bench.py enables it to report the live per-kernel durations its roofline is built on.
 * vst_profile_collect synchronises on the recorded events, fills up to max_entries entries and
 * clears the log.  flops / bytes are the ALGORITHMIC figures of DESIGN.md, summed over launches. */
typedef struct vst_profile_entry {
    char name[40];
    double ms;
    long long launches;
    double flops;
    double bytes;
} vst_profile_entry;
int vst_profile_enable(int on);
int vst_profile_collect(vst_profile_entry* out, int max_entries, int* n_out);

/* -------------------------------------------------------------------------------------------
 * RevResNet  — replaces models/RevResNet.py:166-239 (class RevResNet, _forward, _inverse),
 *              :68-116 (residual_block), :119-163 (channel_reduction), :19-43 (pad / squeeze).
 * ----------------------------------------------------------------------------------------- */
typedef struct vst_revnet_config {
    int n_stages;                 /* len(nBlocks)                       RevResNet.py:168 */
    int n_blocks[VST_MAX_STAGES]; /* nBlocks                            RevResNet.py:168 */
    int n_strides[VST_MAX_STAGES];/* nStrides (1 or 2)                  RevResNet.py:169 */
    int n_channels[VST_MAX_STAGES];/* nChannels (half-state channels)   RevResNet.py:170 */
    int in_channel;               /* 3                                  RevResNet.py:171 */
    int mult;                     /* bottleneck divisor, 4              RevResNet.py:172 */
    int hidden_dim;               /* 16 photorealistic / 64 artistic    RevResNet.py:173 */
    int sp_steps;                 /* 2 photorealistic / 1 artistic      RevResNet.py:174 */
    int n_cr_blocks;              /* channel_reduction n_blocks, 2      RevResNet.py:120 */
} vst_revnet_config;

typedef struct vst_revnet vst_revnet; /* host-side plan; holds no device memory */

/* precision of the convolution arithmetic */
#define VST_CONV_FP32 0     /* CUDA-core FFMA, fp32 exact products                          */
#define VST_CONV_TF32X3 1   /* tcgen05 kind::tf32, 3-term error-compensated split (fp32-equivalent) */
#define VST_CONV_TF32X2 2   /* tcgen05 kind::tf32, activations split hi+lo, weights rounded to tf32 */
#define VST_CONV_TF32 3     /* tcgen05 kind::tf32, single term                              */
#define VST_CONV_F16X2 4    /* tcgen05 kind::f16 on the non-coupling convs: activations split hi+lo in fp16 (22 bits),
                               weights rounded to fp16 (11 bits, as tf32); coupling convs as VST_CONV_TF32X2 */

int vst_revnet_create(const vst_revnet_config* cfg, vst_revnet** out);   /* RevResNet.__init__ :167-190 */
void vst_revnet_destroy(vst_revnet* net);
int vst_revnet_set_precision(vst_revnet* net, int mode);
int vst_revnet_latent_channels(const vst_revnet* net);                    /* 2*hidden_dim */
int vst_revnet_down_scale(const vst_revnet* net);                         /* prod(nStrides), RevResNet.py:186 */

/* Raw parameters: ONE flat fp32 device buffer holding the reference's state_dict tensors in
 * state_dict order (stack.{i}.conv.{1,4,7}.{weight,bias}, then
 * channel_reduction.block_list.{j}.conv.{1,4,7}.{weight,bias}), each OIHW contiguous
 * (SURVEY.md A.2).  vst_revnet_pack_weights repacks them on the device into the kernels'
 * layouts (incl. the hi/lo tf32 split); `packed` must have vst_revnet_packed_bytes() bytes. */
size_t vst_revnet_param_floats(const vst_revnet* net);
size_t vst_revnet_packed_bytes(const vst_revnet* net);
int vst_revnet_pack_weights(const vst_revnet* net, const float* raw_params, void* packed, void* stream);

/* scratch needed by forward / inverse for a B x 3 x H x W image (H, W multiples of down_scale).
 * The first 4 bytes of the workspace are an int32 STATUS WORD that every forward / inverse call clears in its first
 * kernel and that is valid once the call's stream work has completed (the library never synchronises):
 *   bit 0 (VST_STATUS_F16_RANGE): precision VST_CONV_F16X2 only — an activation left the fp16 operand range
 *   (|x| >= 1023.75), the result of this call is invalid; re-run with VST_CONV_TF32X2. */
#define VST_STATUS_F16_RANGE 1
size_t vst_revnet_workspace_bytes(const vst_revnet* net, int B, int H, int W);

/* encode: x [B,in_channel,H,W] -> z [B, 2*hidden_dim, H*2^sp/ds, W*2^sp/ds]
 * replaces RevResNet.forward(x, forward=True) (RevResNet.py:203-223). */
int vst_revnet_forward(const vst_revnet* net, const void* packed, const float* x, float* z,
                       int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream);
/* decode: z -> x [B,in_channel,H,W]; z is not modified.
 * replaces RevResNet.forward(z, forward=False) (RevResNet.py:225-239). */
int vst_revnet_inverse(const vst_revnet* net, const void* packed, const float* z, float* x,
                       int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* Fused stylization of one frame against hoisted style statistics — the video hot path
 * (video_transfer.py:192-206 with the loop-invariant style encode of :195 hoisted):
 *     encode -> cWCT statistics / factor / apply ON THE NETWORK'S OWN STATE -> decode
 * The latent z is never materialised: the statistics are taken from the two half-states that channel_reduction's
 * spread loops (RevResNet.py:140-146) would turn into z, and T (x - mu) + beta is applied to them in place
 * (SURVEY.md 8(f) rank 1).  frame_in / frame_out are fp32 [3,H,W] in [0,1] (io_u8 == 0) or uint8 [H,W,3]
 * (io_u8 != 0; bgr != 0: channel order B,G,R) with ToTensor's byte/255 and the mul(255).clamp(0,255).byte()
 * truncation of video_transfer.py:188, :211-214 folded into the first and last kernel (8(f) rank 2).
 * style_stats: a one-label block of vst_cwct_stats over the style latent.  Results equal
 * vst_revnet_forward + vst_cwct_stats/_factor/_apply + vst_revnet_inverse up to the summation order of the
 * statistics.  Workspace status words: [1] Cholesky retries (-1: failed, the frame passes through unstylized),
 * [2] validity.  vst_revnet_stylize_supported: 1 if this plan / size can take the fused path. */
int vst_revnet_stylize_supported(const vst_revnet* net, int H, int W);
int vst_revnet_stylize(const vst_revnet* net, const void* packed, const void* frame_in, void* frame_out, int io_u8, int bgr,
                       int H, int W, const void* style_stats, float alpha_c, float eps, int use_double,
                       void* workspace, size_t workspace_bytes, void* stream);

/* -------------------------------------------------------------------------------------------
 * cWCT — replaces models/cWCT.py: whitening :134-149, coloring :152-164, cholesky_dec :111-132,
 *        _transfer :24-47, _transfer_seg :49-109, compute_label_info :166-189, interpolation
 *        :206-262.   out = T_l (x - mu_c,l) + beta_l per label l.
 * ----------------------------------------------------------------------------------------- */

/* A "stats" block describes the first and second moments of one feature map, per label:
 *   count [L] (double), sum [L*C] (double), gram [L*C*C] (double, sum of (x-p)(x-p)^T with the
 *   per-channel pivot p [C] (float) chosen by the library to avoid cancellation).
 * vst_cwct_stats_bytes gives the size of one block; layout is private to the library. */
size_t vst_cwct_stats_bytes(int C, int n_labels);

/* feat [C, n] fp32 (one sample), labels uint8 [n] or NULL (=> one label, 0; n_labels must be 1).
 * Replaces the mean / centre / x@x.T of cWCT.py:138-144, :153-157 and the per-label
 * np.where + index_select of :87-95.  Zero-initialises `stats` itself. */
int vst_cwct_stats(const float* feat, int C, long long n, const uint8_t* labels, int n_labels,
                   void* stats, void* stream);

/* The same for a feature map whose 2-D shape is known, feat [C, H*W], labels [H*W]: large per-label maps of the
 * photorealistic latent (C = 32) then run on the tensor cores (strip-wise traversal, one masked-covariance pass per
 * label present in a stage); everything else behaves as vst_cwct_stats. */
int vst_cwct_stats2d(const float* feat, int C, int H, int W, const uint8_t* labels, int n_labels,
                     void* stats, void* stream);

/* Per label: covariance (n-1 divisor), Cholesky with the cumulative eps*I retry (cWCT.py:111-128),
 * triangular solve, T = (1-alpha_c) * (sum_k alpha_s[k] Ls_k) Lc^-1 + alpha_c I,
 * mu = mu_c,  beta = (1-alpha_c) sum_k alpha_s[k] mu_s,k + alpha_c mu_c   (out = T (x - mu) + beta).
 * masked != 0 applies the validity rule of cWCT.py:178 (n_c>10, n_s>10, ratios < 100) per label;
 * invalid labels get T = I, mu = beta = 0 and valid[l] = 0.
 *   T [L,C,C] fp32, mu [L,C] fp32, beta [L,C] fp32, valid [L] int32, status [L] int32 (#jitter
 *   retries, <0 on failure).  The factor arithmetic is always fp64; use_double (cWCT.py:13) selects the failure
 *   rule of the Cholesky: any positive pivot is accepted (fp64 LAPACK), otherwise a pivot below the fp32 noise floor
 *   of its diagonal counts as a failure and triggers the jitter retry, as fp32 LAPACK would see it. */
int vst_cwct_factor(const void* content_stats, const void* const* style_stats, const float* alpha_s,
                    int n_styles, float alpha_c, float eps, int C, int n_labels, int masked,
                    int use_double, float* T, float* mu, float* beta, int* valid, int* status, void* stream);

/* The two halves of the transform as the reference exposes them (one label, unmasked):
 *   vst_cwct_whiten_factor: T = Lc^-1, mu = mean, beta = 0   => apply gives whitening(x)       (cWCT.py:134-149)
 *   vst_cwct_color_factor : T = Ls,    mu = 0,    beta = mean => apply gives coloring(w, style) (cWCT.py:152-164)
 * `stats` is a one-label block of vst_cwct_stats over the tensor whose covariance is factorised. */
int vst_cwct_whiten_factor(const void* stats, float eps, int C, int use_double, float* T, float* mu, float* beta,
                           int* valid, int* status, void* stream);
int vst_cwct_color_factor(const void* stats, float eps, int C, int use_double, float* T, float* mu, float* beta,
                          int* valid, int* status, void* stream);

/* cholesky_dec (cWCT.py:111-132) of a given C x C matrix on the device: L = chol(cov), on failure cov += eps*I,
 * then 2 eps*I more, ... (cumulative); invert != 0 returns L^-1 (torch.inverse(L)).  cov / out are fp32
 * (is_double == 0) or fp64 [C,C] row-major; only the lower triangle of cov is read; the strict upper triangle of
 * out is zero.  status[0] = #retries, or -1 (out = NaN) if 64 retries did not help.  No host synchronisation. */
int vst_cwct_cholesky(const void* cov, int C, int is_double, float eps, int invert, void* out, int* status,
                      void* stream);

/* out[:,p] = T[l(p)] (feat[:,p] - mu[l(p)]) + beta[l(p)]; labels NULL => l = 0; pixels whose
 * label has valid[l] == 0 are copied through.  `out` may alias `feat`.
 * Replaces inv_L @ x, Ls @ whiten + mean (cWCT.py:147,161-162,256-257) and the index_copy_
 * scatter of :99-101. */
int vst_cwct_apply(const float* feat, float* out, int C, long long n, const uint8_t* labels,
                   int n_labels, const float* T, const float* mu, const float* beta, const int* valid,
                   void* stream);

/* -------------------------------------------------------------------------------------------
 * Frame I/O helpers for the video entry point (video_transfer.py:188, :211-214).
 * ----------------------------------------------------------------------------------------- */
/* uint8 HWC (RGB or BGR) -> fp32 CHW RGB in [0,1]  (ToTensor, video_transfer.py:188) */
int vst_frame_u8_to_f32(const uint8_t* hwc, float* chw, int H, int W, int bgr, void* stream);
/* label map [Hs][Ws] -> [Hd][Wd], nearest neighbour with PIL Image.NEAREST sampling: the reference's cWCT.resize
 * (models/cWCT.py:191-197; its call at :72-73 is commented out in this fork).  SURVEY.md 8(f) rank 3. */
int vst_mask_resize_nearest(const uint8_t* src, int Hs, int Ws, uint8_t* dst, int Hd, int Wd,
                            int* scratch /* Hd + Wd ints, device */, void* stream);
/* fp32 CHW -> uint8 HWC, mul(255).clamp(0,255).byte() truncation (video_transfer.py:211-214) */
int vst_frame_f32_to_u8(const float* chw, uint8_t* hwc, int H, int W, int bgr, void* stream);

/* -------------------------------------------------------------------------------------------
 * Mask preparation on the device.
 * ----------------------------------------------------------------------------------------- */
/* SegReMapping (models/segmentation/SegReMapping.py:19-76) of a uint8 label map [n]:
 *   style_seg == NULL : self_remapping  — labels covering less than min_ratio of the map move to the first label of
 *                       their column of the relation table `mapping` (int32 [rows][n_classes], the reference's
 *                       ade20k_semantic_rel.npy) that is present with a ratio >= min_ratio;
 *   style_seg != NULL : cross_remapping — labels of `seg` that the style map lacks move to the first label of their
 *                       column that the style has.
 * `out` may alias `seg`.  scratch: vst_seg_scratch_bytes() bytes of device memory.  No host synchronisation. */
size_t vst_seg_scratch_bytes(void);
int vst_seg_remap(const uint8_t* seg, long long n, const uint8_t* style_seg, long long n_style, const int* mapping,
                  int rows, int n_classes, float min_ratio, uint8_t* out, void* scratch, void* stream);
/* colour-coded segmentation (uint8 RGB, HWC) -> labels 0..8: exact table colours, else the nearest table colour in
 * L1 (utils/utils.py:105-137, load_segment / change_seg). */
int vst_seg_labels_from_colors(const uint8_t* rgb_hwc, long long n, uint8_t* labels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSTB200_H */
