/*
 * ewvit.h -- C ABI of libewvit.so: the B200 (sm_100a) hot path of Efficient Wavelet ViT.
 *
 * The reference (Sheldon-Xiao9/efficient-wavelet-vit) is pure Python/PyTorch and has no FFI;
 * each entry point below replaces the *library call sequence* that the cited reference lines
 * dispatch to (cuDNN/cuBLAS/ATen eager kernels), and is what a Python `ctypes` binding in the
 * reference's `network/*.py` would bind (see INTEGRATION.md for the stub).
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative EWVIT_ERR_* otherwise; `ewvit_last_error()` returns a
 *     thread-local, NUL-terminated description of the last failure on the calling thread;
 *   - no exceptions cross the ABI; no torch types; plain pointers and sizes;
 *   - every pointer is a DEVICE pointer on the calling thread's current CUDA device unless it
 *     is documented as host; the caller owns every buffer (inputs, outputs, workspaces); the
 *     library never allocates or frees device memory and keeps no reference to caller memory
 *     after the call returns (it caches only TMA descriptors / function attributes);
 *   - `stream` is a `cudaStream_t` passed as `void*`; work is enqueued asynchronously on it and
 *     the call never synchronises the device;
 *   - re-entrant and thread-safe; one process per GPU is the intended deployment;
 *   - tensors are dense, row-major in the index order written in each comment, 16-byte aligned.
 */
#ifndef EWVIT_H_
#define EWVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EWVIT_API __attribute__((visibility("default")))
#else
#define EWVIT_API
#endif

#define EWVIT_OK 0
#define EWVIT_ERR_INVALID_ARG (-1)   /* NULL / misaligned pointer, bad size               */
#define EWVIT_ERR_UNSUPPORTED (-2)   /* shape outside what the kernel family handles      */
#define EWVIT_ERR_CUDA (-3)          /* a CUDA runtime / driver call failed (see message) */
#define EWVIT_ERR_NO_DEVICE (-4)     /* current device is not sm_100                      */

/* Library ABI version (bumped on any signature change). */
EWVIT_API int ewvit_abi_version(void);

/* Thread-local description of the last error on this thread ("" if none). */
EWVIT_API const char *ewvit_last_error(void);

/* Number of kernels this library has launched on the calling process so far (all threads).
 * bench.py uses the difference across the timed region for its `gpu_launches` claim. */
EWVIT_API uint64_t ewvit_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Haar analysis filter bank  
 *
 * Replaces `pytorch_wavelets.DWTForward(J=1, wave='haar', mode='zero').forward`, called at
 * reference network/mwt.py:76 (constructed at mwt.py:20): one level, zero-mode boundary (an odd
 * height / width gets ONE zero row / column appended at the bottom / right; even sizes are not
 * padded).
 *
 *   x  [planes, h, w]            fp32   (planes = N*C of an NCHW tensor)
 *   ll [planes, h2, w2]          fp32   h2 = (h+1)/2, w2 = (w+1)/2
 *   yh [planes, 3, h2, w2]       fp32   subband order as the reference: 0 = (W-low, H-high),
 *                                       1 = (W-high, H-low), 2 = (W-high, H-high)
 * fp32 evaluation order is fixed (see oracle/haar.py) so results are bit-reproducible.
 * ------------------------------------------------------------------------------------------- */
EWVIT_API int ewvit_dwt_haar_fwd(const float *x, int64_t planes, int h, int w,
                       float *ll, float *yh, void *stream);

/* Three chained levels in ONE pass over HBM (the loop at reference network/mwt.py:107-111 taken
 * together with the three DWTForward calls it makes): x is read once, every subband of every
 * level is written once.  Requires h % 8 == 0 and w % 8 == 0 (224 -> 112 -> 56 -> 28), so the
 * zero-mode boundary never fires.  Any of the six outputs may be NULL to skip materialising it.
 *
 *   x   [planes, h, w]
 *   ll1 [planes, h/2, w/2]   hf1 [planes, 3, h/2, w/2]
 *   ll2 [planes, h/4, w/4]   hf2 [planes, 3, h/4, w/4]
 *   ll3 [planes, h/8, w/8]   hf3 [planes, 3, h/8, w/8]
 */
EWVIT_API int ewvit_dwt3_haar_fwd(const float *x, int64_t planes, int h, int w,
                        float *ll1, float *hf1, float *ll2, float *hf2, float *ll3, float *hf3,
                        void *stream);

/* The same three levels straight from uint8 frames (row f-3: the step upstream of the path, reference
 * config/transforms.py:97-98 `ToTensor` + `Normalize`): every sample is converted on load as
 *   v = ((float(u) / 255) - mean[c]) / std[c],   c = plane % channels,  each operation rounded to fp32 separately
 * which is bit-identical to torchvision's to_tensor().sub_(mean).div_(std).  mean, std: [channels] fp32 on the device. */
EWVIT_API int ewvit_dwt3_haar_u8_fwd(const uint8_t *x, const float *mean, const float *stdv, int channels, int64_t planes, int h,
                                     int w, float *ll1, float *hf1, float *ll2, float *hf2, float *ll3, float *hf3, void *stream);

/* ---------------------------------------------------------------------------------------------
 * bf16 tensor-core linear layer  (rows a-5, a-7: nn.Linear call sites network/sfe.py:155 patch_to_embedding,
 * sfe.py:52,54 to_qkv/to_out, sfe.py:31-37 FeedForward, sfe.py:141 feat_map; all `x @ W^T + b`)
 *
 *   out[M, N] = act( (A[M, K] @ W[N, K]^T) * scale[N] + shift[N] + residual[M, N] )
 *
 *   a        [M, K]  bf16 row-major            w  [N, K] bf16 row-major (the nn.Linear weight as stored)
 *   scale    [N] fp32 or NULL (= 1)            shift [N] fp32 or NULL (= 0; the bias)
 *   act      0 none, 1 ReLU, 2 GELU (erf)      residual [M, ldr] fp32 or NULL (added BEFORE act)
 *   out      [M, ldo] bf16 (out_fp32 = 0) or fp32 (out_fp32 = 1)
 *   splits   > 1 selects split-K: `workspace` must hold splits*M*N floats; partial sums are reduced
 *            in a fixed order (deterministic).
 * Needs K % 64 == 0 and N % 128 == 0.  fp32 accumulation in TMEM (tcgen05.mma kind::f16).
 * ------------------------------------------------------------------------------------------- */
EWVIT_API int ewvit_linear_bf16(const void *a, const void *w, int64_t M, int N, int K,
                                const float *scale, const float *shift, int act,
                                const float *residual, int64_t ldr,
                                void *out, int out_fp32, int64_t ldo,
                                int splits, float *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------
 * 3x3 convolution, padding 1, stride 1 or 2, as a bf16 implicit GEMM with a fused per-channel
 * scale/shift (folded conv bias + eval-mode BatchNorm) and ReLU  (rows a-3, a-4: the Conv2d+BatchNorm2d+
 * ReLU triples at reference network/mwt.py:33-36 freq_conv, :40-42 freq_pool, :60-64 hf_conv.fusion,
 * :68-72 multiscale_fusion).
 *
 *   x  NHWC bf16: [n, h, wd, x_ldc], or with in_padded = 1 [n, h+2, wd+2, x_ldc] carrying an explicit zero border; the conv
 *      reads the cin channels starting at channel x_coff of every pixel (x_ldc = cin, x_coff = 0: a dense tensor; a slice of a
 *      wider tensor -- one level of the three-level MWT head -- is implemented for the stride-1 padded-flat path)
 *   w  [cout, 3, 3, cin] bf16 (tap-major K: k = (ky*3 + kx)*cin + c)
 *   y  NHWC bf16 with channel pitch y_ldc, written at channel offset y_coff (lets three producers fill
 *      one concatenated buffer, mwt.py:113): [n, ho, wo, y_ldc], or with out_padded = 1
 *      [n, ho+2, wo+2, y_ldc];  ho = (h-1)/stride + 1
 *   y[.., co] = relu?( conv(x, w)[.., co] * scale[co] + shift[co] )
 * stride 1 with in_padded = out_padded = 1 takes the row-shift path and ALSO writes the zero border of y;
 * every other combination takes the box path (stride-2 via the TMA element stride) and writes interior
 * pixels only (a padded y must have been zeroed once by the caller).  force_tiled = 1 forces the box path.
 * Needs cin % 64 == 0 and cout % 128 == 0.
 * ------------------------------------------------------------------------------------------- */
EWVIT_API int ewvit_conv3x3_bf16(const void *x, int x_ldc, int x_coff, const void *w, int n, int h, int wd, int cin, int cout,
                                 int stride, int in_padded, const float *scale, const float *shift, int relu,
                                 void *y, int y_ldc, int y_coff, int out_padded, int force_tiled, void *stream);

/* ---------------------------------------------------------------------------------------------
 * MWT glue  (rows a-3, a-4)
 * ------------------------------------------------------------------------------------------- */

/* High-frequency head of the three wavelet levels (model.py:35 levels = 3): reference network/mwt.py:77-86 per level --
 * `hf[0].reshape(B, 3C, h, w)` (colour-major), `F.interpolate(..., mode='bilinear')` to the level-1 grid (align_corners=False;
 * identity at level 1), then the three per-colour Conv2d(3->18,3x3,p1)+BN+ReLU of hf_conv['seperate'] and their channel concat --
 * for all levels in two launches (what one MWT.wavelet_transform loop, mwt.py:107-111, computes before hf_conv.fusion):
 *   1. ewvit_mwt_upsample3_fwd: hf1 [n, 9, h, wd], hf2 [n, 9, h/2, wd/2], hf3 [n, 9, h/4, wd/4] fp32 (h x wd = the level-1 grid,
 *      multiples of 4) -> up [n, h+2, wd+2, 32] bf16 padded-flat NHWC: channel 9 l + c = subband c of level l + 1 upsampled to
 *      h x wd, channels 27..31 zero.  Interior pixels only: the one-pixel border must be zero (zero the buffer once).
 *   2. ewvit_mwt_head_conv3_fwd: the convs as one block-diagonal conv on the tensor cores.
 *      w [128, 288] bf16 = per tap (dy*3 + dx) and K step one [128, 16] tile, w[r][((dy*3 + dx)*2 + step)*16 + kk]:
 *      rows [0, 64) of step 0 = level 1, rows [64, 128) of step 0 and rows [0, 64) of step 1 = level 2, rows [64, 128) of step 1 =
 *      level 3; inside a 64-row block row g*18+oc holds seperate[g].weight[oc][ic][dy][dx] where channel 16*step + kk of a pixel is
 *      subband 3g+ic of that level, zero elsewhere (the head is shared by the levels); scale/shift [192] fp32 = the 64-entry
 *      folded bias + eval BatchNorm (zeros past channel 54) repeated three times; y [n, h+2, wd+2, 192] bf16 padded-flat: channels
 *      [64 l, 64 l + 54) = the 54-channel head of level l + 1 (the input of hf_conv.fusion, read in place through x_ldc / x_coff
 *      of ewvit_conv3x3_bf16), everything else and the border zero. */
EWVIT_API int ewvit_mwt_upsample3_fwd(const float *hf1, const float *hf2, const float *hf3, int n, int h, int wd, void *up, void *stream);
EWVIT_API int ewvit_mwt_head_conv3_fwd(const void *up, const void *w, int n, int h, int wd, const float *scale,
                                       const float *shift, void *y, void *stream);

/* nn.MaxPool2d(2, 2) on NHWC bf16 (freq_pool[0], mwt.py:39): x [n,h,w,c] -> y [n,h/2,w/2,c]. */
EWVIT_API int ewvit_maxpool2x2_nhwc_bf16(const void *x, int64_t n, int h, int w, int c, void *y, void *stream);

/* nn.AdaptiveAvgPool2d(1) on NHWC bf16 (freq_pool[4], mwt.py:43): x [n, hw, c] -> y [n, ldy] fp32. */
EWVIT_API int ewvit_gap_nhwc_bf16(const void *x, int64_t n, int hw, int c, float *y, int64_t ldy, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Efficient-ViT token glue  (row a-5)
 * ------------------------------------------------------------------------------------------- */

/* reference network/sfe.py:156-159: x = cat(cls_token, patch_embedding) + pos_embedding[0:N], with the
 * position resolved per frame by the caller (pos_index[f] = position of frame f inside its reference chunk).
 *   emb [n, d] fp32 (patch_to_embedding output incl. bias)   cls [d]   pos [pos_rows, d]   pos_index [n] int32
 *   x   [n, 2, d] fp32 */
EWVIT_API int ewvit_vit_assemble(const float *emb, const float *cls, const float *pos, const int *pos_index,
                                 int64_t n, int d, int pos_rows, float *x, void *stream);

/* nn.LayerNorm (sfe.py:23, dama.py:63-66) on fp32 rows, bf16 result ready to be a GEMM operand.
 * gamma == beta == NULL: plain strided fp32 -> bf16 cast (used to pick token 1 for feat_map, sfe.py:171).
 *   x [rows, ldx] fp32 -> y [rows, ldy] bf16 over the first d columns */
EWVIT_API int ewvit_layernorm_bf16(const float *x, int64_t ldx, const float *gamma, const float *beta, float eps,
                                   void *y, int64_t ldy, int64_t rows, int d, void *stream);

/* Attention core of sfe.py:58-69 for `tokens` tokens per frame (2 at the shipped config):
 *   qkv [n*tokens, 3*heads*dim_head] fp32 (q | k | v, head-major) -> out [n*tokens, heads*dim_head] bf16 */
EWVIT_API int ewvit_vit_attention(const float *qkv, int64_t n, int tokens, int heads, int dim_head, void *out,
                                  void *stream);

/* ---------------------------------------------------------------------------------------------
 * DAMA fusion tail  (rows a-7 .. a-10)
 * ------------------------------------------------------------------------------------------- */

/* Number of floats in the packed weight buffer of ewvit_dama_tail_fwd. */
EWVIT_API int64_t ewvit_dama_wpack_floats(int d, int depth);

/* reference network/dama.py:143-169 (`DAMA._process_frame` after the two branches): BidirectionalCrossTransformer
 * (depth x {space attends freq, freq attends space}, kv_include_self=True => 1 query x 2 keys), fusion_gate
 * (centre tap of the 3x3 conv on the 1x1 map + eval BatchNorm + ReLU), gate_net (Linear-ReLU-Linear-softmax(3)),
 * weighted sum.  fp32 throughout.
 *   space_in, freq_in [n, d]                 fused, space, freq [n, d] (outputs)
 *   wpack: fp32, all matrices TRANSPOSED to [in, out]:
 *     for l in 0..depth-1, for dir in (space<-freq, freq<-space):
 *        ln_w[d] ln_b[d] to_q^T[d,d] to_kv^T[d,2d] to_out^T[d,d] to_out_bias[d]
 *     fusion_gate centre tap^T [2d,d], folded scale[d], folded shift[d],
 *     gate_net.2^T [2d, d/2], gate_net.2 bias [d/2], gate_net.5 weight [3, d/2] (not transposed), gate_net.5 bias [3] */
EWVIT_API int ewvit_dama_tail_fwd(const float *space_in, const float *freq_in, int64_t n, int d, int heads, int depth,
                                  const float *wpack, float ln_eps, float *fused, float *space, float *freq,
                                  void *stream);

/* reference network/dama.py:188-199 (per-video mean over the k frames of each video; frames of a video are
 * consecutive rows) and network/model.py:62-68,92 (classifier Linear(d->hc)+ReLU+Linear(hc->1), eval mode).
 * space/freq (and their means) may be NULL; cw1 == NULL skips the classifier.
 *   fused/space/freq [videos*k, d] -> mean_* [videos, d];  logits [videos]
 *   cw1 [hc, d], cb1 [hc], cw2 [hc], cb2 [1] */
EWVIT_API int ewvit_video_head_fwd(const float *fused, const float *space, const float *freq, int64_t videos, int k,
                                   int d, float *mean_fused, float *mean_space, float *mean_freq, const float *cw1,
                                   const float *cb1, const float *cw2, const float *cb2, int hc, float *logits,
                                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * EfficientNet feature extractor  (row a-6 / f-1; third-party arithmetic: torchvision `efficientnet_v2_s(...).features`
 * called at reference network/sfe.py:150, eval mode, BatchNorm folded into weights/bias by the host)
 * All activations NHWC bf16.
 * ------------------------------------------------------------------------------------------- */

/* Dense NHWC convolution on the tcgen05 implicit-GEMM kernel with a fused epilogue:
 *   y = act(conv(x, w) + bias) + residual          (the skip connection is added after the activation)
 * ksize 1 (stride 1): w [cout, cin] bf16.   ksize 3 (pad 1, stride 1|2; cin <= 64 or cin % 64 == 0):
 * w [cout, Kpad] bf16 with the dense tap-major index k = (ky*3+kx)*cin + c, zero-padded to Kpad = ceil(9*cin/64)*64.
 *   x [n, h, wd, cin] bf16    y, residual [n, ho, wo, cout] bf16 (residual may be NULL)    bias [cout] fp32 or NULL
 *   act: 0 none, 1 ReLU, 3 SiLU, 4 SiLU on pre-halved operands (the caller passes w/2 and bias/2 -- exact in bf16/fp32 --
 *        and the epilogue evaluates silu(v) = h*tanh(h) + h with h = v/2 directly: one multiply less per output).
 *   cin, cout multiples of 8 (tails are zero-filled by TMA and masked). */
EWVIT_API int ewvit_conv_nhwc_bf16(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int ksize,
                                   int stride, const float *bias, int act, const void *residual, void *y, void *stream);

/* The same convolution on "padded-flat" tensors [n, h+2, wd+2, c] (one-pixel zero border around every image):
 *   in_padded / out_padded say which of x / (y, residual) use that layout.  Padded outputs: 1x1 convs and the stride-1
 *   small-channel 3x3 convs write the border as zeros; the stride-2 small-channel 3x3 convs write the interior only (the
 *   caller zeroes the buffer once).  1x1 convs need in_padded == out_padded.  A 3x3 stride-1 conv with cin < 64 on padded
 *   input AND output takes the "overlapping window" path: its weights are w [cout, 3*nsub*64] bf16 with
 *   nsub = ceil(3*cin/64) and k = dy*(nsub*64) + dx*cin + c (zero elsewhere); every other case uses the layouts above. */
EWVIT_API int ewvit_conv_nhwc_bf16_ex(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int ksize,
                                      int stride, const float *bias, int act, const void *residual, void *y, int in_padded,
                                      int out_padded, void *stream);

/* 3x3 / stride 1 / pad 1 convolution with exactly 24 input and 24 output channels + bias + SiLU (+ residual = x):
 * the stage-1 FusedMBConv blocks of EfficientNetV2-S.  Direct convolution on warp-level tensor-core MMAs (a 128-row
 * tcgen05 tile is all overhead at N = 24).  x, y [n, h, wd, 24] bf16; w [24, wk] bf16 in the dense tap-major layout of
 * ewvit_conv_nhwc_bf16 (k = (ky*3+kx)*24 + c, wk >= 216 elements per row); bias [24] fp32. */
EWVIT_API int ewvit_conv3x3_c24_fwd(const void *x, const void *w, int wk, const float *bias, int n, int h, int wd, int residual,
                                    void *y, void *stream);

/* Stem: Conv2d(3 -> cout, 3x3, stride 2, pad 1) + bias + SiLU straight from the fp32 NCHW frames (also the
 * fp32 -> bf16 / NCHW -> NHWC conversion).  x [n,3,h,wd] fp32, w [cout,3,3,3] fp32, y [n,ho,wo,cout] bf16. */
EWVIT_API int ewvit_stem_conv_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                  void *y, void *stream);
/* Same with TensorFlow 'SAME' padding (efficientnet_pytorch's Conv2dStaticSamePadding, EfficientNet-b0 stem 3 -> 32,
 * network/sfe.py:109,148): on even sizes the only padding is one zero row/column at the bottom/right. */
EWVIT_API int ewvit_stem_conv_same_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                       void *y, void *stream);
/* Same as ewvit_stem_conv_fwd, writing the interior of a padded-flat output y [n, ho+2, wo+2, cout] (the caller zeroes the border once). */
EWVIT_API int ewvit_stem_conv_padded_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                         void *y, void *stream);
/* Stem from uint8 frames x [n,3,h,wd] with the same on-load normalisation as ewvit_dwt3_haar_u8_fwd (the conv's zero
 * padding is applied after the normalisation, as in the reference); out_padded selects the padded-flat output layout. */
EWVIT_API int ewvit_stem_conv_u8_fwd(const uint8_t *x, const float *mean, const float *stdv, int n, int h, int wd, const float *w,
                                     const float *bias, int cout, void *y, int out_padded, void *stream);

/* Depthwise 3x3 (pad 1, stride 1|2) + bias + SiLU, plus the squeeze of the SE block: pooled[n, c] = spatial mean of
 * the stored result (NULL to skip).  x [n,h,wd,c] bf16, w [9, c] fp32 (tap-major), y [n,ho,wo,c] bf16; c % 64 == 0. */
EWVIT_API int ewvit_dwconv3x3_nhwc_bf16(const void *x, const float *w, const float *bias, int n, int h, int wd, int c,
                                        int stride, void *y, float *pooled, void *stream);

/* Depthwise k x k (k = 3 | 5, stride 1 | 2) + bias + activation (act: 0 none, 4 SiLU) with explicit top/left zero padding and
 * output size, so both torchvision's symmetric padding (EfficientNetV2-S: pad k/2, ho = (h-1)/s+1; network/sfe.py:111-113)
 * and TensorFlow 'SAME' padding (efficientnet_pytorch b0: pad_top = total/2 with the odd pixel at the bottom/right,
 * ho = ceil(h/s); network/sfe.py:109) are covered.  x [n,h,wd,c] bf16, w [k*k, c] fp32 (tap-major), bias [c], y [n,ho,wo,c]
 * bf16; c even.  pooled (NULL to skip) receives the SE squeeze as ewvit_dwconv_pool_parts(ho, wo, k, s) partial means per
 * frame: pooled[n, parts, c] fp32, whose sum over `parts` is the spatial mean of the stored result. */
EWVIT_API int ewvit_dwconv_nhwc_bf16(const void *x, const float *w, const float *bias, int n, int h, int wd, int c, int ksize,
                                     int stride, int pad_top, int pad_left, int ho, int wo, int act, void *y, float *pooled,
                                     void *stream);
EWVIT_API int ewvit_dwconv_pool_parts(int ho, int wo, int ksize, int stride);

/* Squeeze-excitation gate (torchvision.ops.SqueezeExcitation with SiLU / Sigmoid; efficientnet_pytorch's _se_reduce / _se_expand):
 * gate[n, c] = sigmoid(W2 silu(W1 pooled[n] + b1) + b2) with w1 [sq, c], b1 [sq], w2t [sq, c] (= fc2 weight transposed), b2 [c];
 * the scaling x * gate itself is fused into ewvit_conv1x1_gated_nhwc_bf16.
 * pooled [n, pool_parts, c]: partial means as written by ewvit_dwconv_nhwc_bf16, summed here in index order (pool_parts = 1:
 * plain [n, c] means).  gate_bf16 != 0: the gates are written as bf16 (what the gated conv multiplies fastest), else fp32. */
EWVIT_API int ewvit_se_gate_fwd(const float *pooled, int pool_parts, const float *w1, const float *b1, const float *w2t, const float *b2,
                                int n, int c, int sq, void *gate, int gate_bf16, void *stream);

/* 1x1 convolution behind a squeeze-excitation block (MBConv project conv):
 *   y = act(((x * gate[frame]) W^T) + bias) + residual
 * x [n, hw, cin] bf16, gate [n, cin] fp32 (gate_bf16 == 0) or bf16 (gate_bf16 != 0), w [cout, cin] bf16, y/residual [n, hw, cout] bf16.  The gate is applied while
 * the A operand tile is assembled in shared memory, so the expanded tensor is read exactly once. */
EWVIT_API int ewvit_conv1x1_gated_nhwc_bf16(const void *x, const void *gate, int gate_bf16, const void *w, int n, int hw, int cin,
                                            int cout, const float *bias, int act, const void *residual, void *y, void *stream);

/* Debug aid for kernel bring-up: when non-NULL, CTA 0 of every subsequent tensor-core GEMM/conv launch writes
 * clock64 stamps of its warp roles to this device buffer ([6 roles][64 tiles][4] int64).  NULL switches it off. */
EWVIT_API int ewvit_debug_set_trace(void *device_buffer);

/* Debug / A-B switches of the tensor-core kernel (tools/*.py use them to attribute time).  0 restores normal behaviour; never
 * set in production.  1: generic backbone epilogue skips its global stores; 2: ... skips the activation; 32: MWT 3x3 convs
 * fetch one tile per tap instead of row-shared windows; 64: no resident weights; 256: generic instead of specialised
 * backbone epilogues; 512: MWT epilogue does no work at all. */
EWVIT_API int ewvit_debug_set_flags(int flags);

/* Binary classification metrics of the evaluation loop (eval.py:79-94,174-192; scikit-learn definitions) computed on the device
 * from the gathered per-video scores: scores [n] fp32 probabilities (sigmoid of the logits), labels [n] int32 (0 real / 1 fake),
 * n <= 8192.  out [12] fp32: 0 auc (roc_auc_score), 1 eer, 2 eer_threshold (calculate_eer on roc_curve with its default
 * drop_intermediate), 3 accuracy, 4 precision, 5 recall, 6 f1 (threshold 0.5, 0 when undefined), 7 average precision,
 * 8..11 confusion matrix tn, fp, fn, tp.  auc / eer / ap are NaN when only one class is present. */
EWVIT_API int ewvit_binary_metrics_fwd(const float *scores, const int *labels, int n, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EWVIT_H_ */
