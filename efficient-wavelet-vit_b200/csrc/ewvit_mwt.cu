// MWT glue kernels around the tensor-core convs.
//
//  * mwt_upsample: high-frequency subbands of one level -> bilinear upsample to the level-1 grid (F.interpolate,
//    mwt.py:79-81) -> bf16 "padded-flat" NHWC [N, H+2, W+2, 16], the input of the block-diagonal head conv
//    (ewvit_mwt_head_conv_fwd in ewvit_gemm.cu: the three per-colour Conv2d(3->18,3x3,p1)+BN+ReLU of mwt.py:84-86).
//  * maxpool2x2 (freq_pool[0], mwt.py:39) and the global average pool (freq_pool[4], mwt.py:43) on NHWC bf16.
#include "ewvit_common.cuh"

namespace {

// PyTorch bilinear source index (align_corners=False): max(0, scale*(dst+0.5)-0.5)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int &i0, int &i1, float &l1) {
    float s = scale * (dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
}


// ---- tensor-core head, step 1: high-frequency subbands of one level -> bilinear upsample (F.interpolate,
//      align_corners=False, mwt.py:79-81; identity at level 1) -> bf16 "padded-flat" NHWC [n, hout+2, wout+2, 16]
//      (channels 0..8 = the nine subbands in the reference's colour-major order, 9..15 zero; the one-pixel border is
//      never written and must be zero).  One thread per output pixel: 32 bytes out, coalesced.  Step 2 is the
//      block-diagonal 9 -> 54 conv on the tensor cores (ewvit_mwt_head_conv_fwd).
__global__ void __launch_bounds__(256) mwt_upsample_kernel(const float *__restrict__ hf, __nv_bfloat16 *__restrict__ y, long long n,
                                                           int hin, int win, int hout, int wout, float ry, float rx) {
    const long long total = n * hout * wout;
    const bool same = (hin == hout) && (win == wout);
    const long long plane = (long long)hin * win;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ox = (int)(i % wout);
        const long long t = i / wout;
        const int oy = (int)(t % hout);
        const long long img = t / hout;
        const float *src = hf + img * 9 * plane;
        float v[9];
        if (same) {
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = __ldg(src + c * plane + (long long)oy * win + ox);
        } else {
            int ya, yb, xa, xb;
            float ly, lx;
            src_index(oy, ry, hin, ya, yb, ly);
            src_index(ox, rx, win, xa, xb, lx);
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const float *pc = src + c * plane;
                const float v00 = __ldg(pc + ya * win + xa), v01 = __ldg(pc + ya * win + xb);
                const float v10 = __ldg(pc + yb * win + xa), v11 = __ldg(pc + yb * win + xb);
                v[c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
            }
        }
        uint4 lo, hi;
        __nv_bfloat162 b;
        b = __floats2bfloat162_rn(v[0], v[1]); lo.x = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[2], v[3]); lo.y = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[4], v[5]); lo.z = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[6], v[7]); lo.w = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[8], 0.f);  hi.x = *reinterpret_cast<uint32_t *>(&b);
        hi.y = hi.z = hi.w = 0u;
        uint4 *dst = reinterpret_cast<uint4 *>(y + ((img * (hout + 2) + oy + 1) * (wout + 2) + ox + 1) * 16);
        dst[0] = lo;
        dst[1] = hi;
    }
}

// Upsampling variant (hin < hout): a CTA owns `rows_out` output rows of one frame, stages the source rows those need for
// all nine planes in shared memory with coalesced loads and gathers the four bilinear taps from there -- the direct kernel
// above issues 36 scattered global loads per output pixel and is L1-gather bound (115 us per level against ~35 us of HBM
// time).  Same arithmetic, same results.
__global__ void __launch_bounds__(256) mwt_upsample_smem_kernel(const float *__restrict__ hf, __nv_bfloat16 *__restrict__ y, int hin,
                                                                int win, int hout, int wout, float ry, float rx, int rows_out,
                                                                int src_rows_max) {
    extern __shared__ __align__(16) float up_s[];            // [9][src_rows_max][win]
    const int parts = (hout + rows_out - 1) / rows_out;
    const long long img = blockIdx.x / parts;
    const int part = blockIdx.x % parts;
    const int oy0 = part * rows_out, oy1 = min(hout, oy0 + rows_out);
    int ya0, yb_, ya1, yb1;
    float l_;
    src_index(oy0, ry, hin, ya0, yb_, l_);
    src_index(oy1 - 1, ry, hin, ya1, yb1, l_);
    const int nrows = yb1 - ya0 + 1;                          // <= src_rows_max (checked on the host)
    const float *src = hf + img * 9 * (long long)hin * win;
    const int row_f4 = win >> 2;                              // win % 4 == 0 (checked on the host)
    for (int i = threadIdx.x; i < 9 * nrows * row_f4; i += 256) {
        const int c = i / (nrows * row_f4), rem = i - c * (nrows * row_f4);
        const int r = rem / row_f4, q = rem - r * row_f4;
        reinterpret_cast<float4 *>(up_s + ((size_t)c * src_rows_max + r) * win)[q] =
            __ldg(reinterpret_cast<const float4 *>(src + ((long long)c * hin + ya0 + r) * win) + q);
    }
    __syncthreads();
    const int npix = (oy1 - oy0) * wout;
    for (int i = threadIdx.x; i < npix; i += 256) {
        const int oy = oy0 + i / wout, ox = i % wout;
        int ya, yb, xa, xb;
        float ly, lx;
        src_index(oy, ry, hin, ya, yb, ly);
        src_index(ox, rx, win, xa, xb, lx);
        float v[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const float *pc = up_s + (size_t)c * src_rows_max * win;
            const float v00 = pc[(ya - ya0) * win + xa], v01 = pc[(ya - ya0) * win + xb];
            const float v10 = pc[(yb - ya0) * win + xa], v11 = pc[(yb - ya0) * win + xb];
            v[c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
        }
        uint4 lo, hi;
        __nv_bfloat162 b;
        b = __floats2bfloat162_rn(v[0], v[1]); lo.x = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[2], v[3]); lo.y = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[4], v[5]); lo.z = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[6], v[7]); lo.w = *reinterpret_cast<uint32_t *>(&b);
        b = __floats2bfloat162_rn(v[8], 0.f);  hi.x = *reinterpret_cast<uint32_t *>(&b);
        hi.y = hi.z = hi.w = 0u;
        uint4 *dst = reinterpret_cast<uint4 *>(y + ((img * (hout + 2) + oy + 1) * (wout + 2) + ox + 1) * 16);
        dst[0] = lo;
        dst[1] = hi;
    }
}

// 2x2/stride-2 max pool on NHWC bf16; 8 channels (16 bytes) per thread.
__global__ void maxpool2x2_kernel(const __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *__restrict__ y, long long n, int h,
                                  int w, int c) {
    const int ho = h / 2, wo = w / 2, c8 = c / 8;
    const long long total = n * ho * wo * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8);
        long long t = i / c8;
        const int ox = (int)(t % wo);
        t /= wo;
        const int oy = (int)(t % ho);
        const long long img = t / ho;
        const __nv_bfloat16 *p00 = x + ((img * h + 2 * oy) * w + 2 * ox) * c + cc * 8;
        uint4 a = *reinterpret_cast<const uint4 *>(p00), b = *reinterpret_cast<const uint4 *>(p00 + c);
        uint4 d = *reinterpret_cast<const uint4 *>(p00 + (long long)w * c), e = *reinterpret_cast<const uint4 *>(p00 + (long long)w * c + c);
        uint4 r;
        const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&a), *pb = reinterpret_cast<const __nv_bfloat162 *>(&b);
        const __nv_bfloat162 *pd = reinterpret_cast<const __nv_bfloat162 *>(&d), *pe = reinterpret_cast<const __nv_bfloat162 *>(&e);
        __nv_bfloat162 *pr = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) pr[k] = __hmax2(__hmax2(pa[k], pb[k]), __hmax2(pd[k], pe[k]));
        *reinterpret_cast<uint4 *>(y + ((img * ho + oy) * wo + ox) * c + cc * 8) = r;
    }
}

// Global average pool over hw pixels: x [n, hw, c] bf16 -> y [n, ldy] fp32 at channel offset.
__global__ void gap_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y, int hw, int c, long long ldy) {
    const long long img = blockIdx.x;
    const __nv_bfloat16 *px = x + img * hw * c;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < hw; ++i) s += __bfloat162float(px[(long long)i * c + ch]);
        y[img * ldy + ch] = s / (float)hw;
    }
}

}  // namespace



extern "C" int ewvit_mwt_upsample_fwd(const float *hf, int n, int hin, int win, int hout, int wout, void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hin > 0 && win > 0 && hout > 0 && wout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_upsample_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(hf && y && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG, "ewvit_mwt_upsample_fwd: NULL or misaligned pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    if ((hin < hout || win < wout) && win % 4 == 0 && ewvit_aligned16(hf) && n <= (1 << 22)) {
        // staged variant: ~half a frame of output rows per CTA when that keeps the nine source planes under ~64 KB
        int rows_out = hout;
        auto src_rows = [&](int ro) { return (int)((long long)ro * hin / hout) + 3; };
        while (rows_out > 8 && (size_t)9 * src_rows(rows_out) * win * 4 > 64 * 1024) rows_out = (rows_out + 1) / 2;
        const int srm = src_rows(rows_out);
        const size_t smem = (size_t)9 * srm * win * 4;
        if (smem <= 96 * 1024) {
            static bool attr_set[64] = {false};
            int dev = 0;
            EWVIT_CUDA_OK(cudaGetDevice(&dev));
            if (dev < 0 || dev >= 64 || !attr_set[dev]) {
                EWVIT_CUDA_OK(cudaFuncSetAttribute(mwt_upsample_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                if (dev >= 0 && dev < 64) attr_set[dev] = true;
            }
            const int parts = (hout + rows_out - 1) / rows_out;
            mwt_upsample_smem_kernel<<<(unsigned)((long long)n * parts), 256, smem, (cudaStream_t)stream>>>(
                hf, static_cast<__nv_bfloat16 *>(y), hin, win, hout, wout, (float)hin / (float)hout, (float)win / (float)wout, rows_out, srm);
            EWVIT_LAUNCH_OK();
            return EWVIT_OK;
        }
    }
    const long long total = (long long)n * hout * wout;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ewvit_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    mwt_upsample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hf, static_cast<__nv_bfloat16 *>(y), n, hin, win, hout, wout,
                                                                           (float)hin / (float)hout, (float)win / (float)wout);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_maxpool2x2_nhwc_bf16(const void *x, int64_t n, int h, int w, int c, void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0, EWVIT_ERR_INVALID_ARG, "ewvit_maxpool2x2_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && y && ewvit_aligned16(x) && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG,
                  "ewvit_maxpool2x2_nhwc_bf16: NULL or misaligned pointer");
    EWVIT_REQUIRE(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_maxpool2x2_nhwc_bf16: needs even h, w and c %% 8 == 0");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const long long total = n * (h / 2) * (w / 2) * (c / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ewvit_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    maxpool2x2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16 *>(x), static_cast<__nv_bfloat16 *>(y), n, h, w, c);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_gap_nhwc_bf16(const void *x, int64_t n, int hw, int c, float *y, int64_t ldy, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hw > 0 && c > 0 && ldy >= c, EWVIT_ERR_INVALID_ARG, "ewvit_gap_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && y, EWVIT_ERR_INVALID_ARG, "ewvit_gap_nhwc_bf16: NULL pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    gap_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x), y, hw, c, ldy);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
