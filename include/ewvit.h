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
 *   x  NHWC bf16: [n, h, wd, cin], or with in_padded = 1 [n, h+2, wd+2, cin] carrying an explicit zero border
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
EWVIT_API int ewvit_conv3x3_bf16(const void *x, const void *w, int n, int h, int wd, int cin, int cout,
                                 int stride, int in_padded, const float *scale, const float *shift, int relu,
                                 void *y, int y_ldc, int y_coff, int out_padded, int force_tiled, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EWVIT_H_ */
