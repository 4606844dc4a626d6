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

#ifdef __cplusplus
}
#endif
#endif /* EWVIT_H_ */
