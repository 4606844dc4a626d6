// MWT glue kernels around the tensor-core convs.
//
//  * mwt_upsample3: high-frequency subbands of the three levels -> bilinear upsample to the level-1 grid (F.interpolate,
//    mwt.py:79-81) -> bf16 "padded-flat" NHWC [N, H+2, W+2, 32], the input of the block-diagonal head conv
//    (ewvit_mwt_head_conv3_fwd in ewvit_gemm.cu: the three per-colour Conv2d(3->18,3x3,p1)+BN+ReLU of mwt.py:84-86).
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


// ---- all three levels at once: hf_l [n, 9, h >> (l-1), w >> (l-1)] fp32 (l = 1..3, sizes relative to the level-1 grid h x w)
//      -> up [n, h+2, w+2, 32] bf16 padded-flat, a pixel = one 64-byte row: channel 9 l' + c = subband c of level l' + 1 upsampled
//      to the level-1 grid (identity for level 1), channels 27..31 zero.  This is the operand layout of ewvit_mwt_head_conv3_fwd,
//      which serves the three levels from ONE window fetch.  A CTA owns `rows_out` output rows of one frame and stages the source
//      rows of the three levels in shared memory (coalesced loads).  A lane computes the 27 values of one pixel (warp-uniform
//      control flow); the warp then transposes its 32 pixels x 64 bytes through a swizzled shared-memory buffer so that every
//      store instruction writes 512 contiguous bytes.
struct Up3Params {
    const float *hf[3];
    int hin[3], win[3], rows_max[3], off[3];     // source size, staged rows per plane, float offset of the level in shared memory
    float ry[3], rx[3];
    int tr_off;                                  // float offset of the per-warp transposition buffers (8 x 2 KB)
};
__global__ void __launch_bounds__(256) mwt_upsample3_kernel(const Up3Params p, __nv_bfloat16 *__restrict__ y, int hout, int wout,
                                                            int rows_out) {
    extern __shared__ __align__(16) float up3_s[];
    const int parts = (hout + rows_out - 1) / rows_out;
    const long long img = blockIdx.x / parts;
    const int part = blockIdx.x % parts;
    const int oy0 = part * rows_out, oy1 = min(hout, oy0 + rows_out);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ya0[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        int yb_, ya1, yb1;
        float l_;
        src_index(oy0, p.ry[l], p.hin[l], ya0[l], yb_, l_);
        src_index(oy1 - 1, p.ry[l], p.hin[l], ya1, yb1, l_);
        const int nrows = yb1 - ya0[l] + 1;                   // <= rows_max[l] (sized on the host)
        const int win = p.win[l], rmax = p.rows_max[l];
        const float *src = p.hf[l] + img * 9 * (long long)p.hin[l] * win;
        float *dst = up3_s + p.off[l];
        // the staged rows of a plane are ONE contiguous block in global and in shared memory: asynchronous 16-byte copies, all in
        // flight at once (a load-then-store loop left one request per warp outstanding: 0.5 ms per 512 frames, latency-bound)
        const bool vec = (win & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
        const int blk = nrows * win;                           // floats per plane
        for (int c = 0; c < 9; ++c) {
            const float *sp = src + ((long long)c * p.hin[l] + ya0[l]) * win;
            float *dp = dst + (size_t)c * rmax * win;
            if (vec) {
                const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dp);
                for (int q = threadIdx.x; q < (blk >> 2); q += 256)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)q * 16u), "l"(sp + 4 * q) : "memory");
            } else {
                for (int x = threadIdx.x; x < blk; x += 256) dp[x] = __ldg(sp + x);
            }
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const uint32_t tr = (uint32_t)__cvta_generic_to_shared(up3_s + p.tr_off) + (uint32_t)warp * 2048u;
    const int npix = (oy1 - oy0) * wout;
    for (int base = warp * 32; base < npix; base += 256) {
        const int pix = base + lane;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        if (pix < npix) {
            const int oy = oy0 + pix / wout, ox = pix % wout;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const int win = p.win[l], rmax = p.rows_max[l];
                const float *lb = up3_s + p.off[l];
                if (p.hin[l] == hout && win == wout) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) v[9 * l + c] = lb[((size_t)c * rmax + (oy - ya0[l])) * win + ox];
                } else {
                    int ya, yb, xa, xb;
                    float ly, lx;
                    src_index(oy, p.ry[l], p.hin[l], ya, yb, ly);
                    src_index(ox, p.rx[l], win, xa, xb, lx);
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const float *pc = lb + (size_t)c * rmax * win;
                        const float v00 = pc[(ya - ya0[l]) * win + xa], v01 = pc[(ya - ya0[l]) * win + xb];
                        const float v10 = pc[(yb - ya0[l]) * win + xa], v11 = pc[(yb - ya0[l]) * win + xb];
                        v[9 * l + c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
                    }
                }
            }
        }
        // lane = pixel: chunk k (8 channels) goes to position k ^ ((lane >> 1) & 3) of the pixel's 64-byte row (conflict-free both ways)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                w4[i] = *reinterpret_cast<const uint32_t *>(&b);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tr + (uint32_t)lane * 64u + (uint32_t)((k ^ ((lane >> 1) & 3)) << 4)),
                         "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3]) : "memory");
        }
        __syncwarp();
        // lane = (pixel within a group of 8, chunk): one store instruction writes 8 pixels x 64 bytes = 512 contiguous bytes
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pp = 8 * j + (lane >> 2), k = lane & 3;
            uint4 val;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(tr + (uint32_t)pp * 64u + (uint32_t)((k ^ ((pp >> 1) & 3)) << 4)) : "memory");
            const int gp = base + pp;
            if (gp < npix) {
                const int oy = oy0 + gp / wout, ox = gp % wout;
                *reinterpret_cast<uint4 *>(y + ((img * (hout + 2) + oy + 1) * (wout + 2) + ox + 1) * 32 + k * 8) = val;
            }
        }
        __syncwarp();
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

// Global average pool over hw pixels: x [n, hw, c] bf16 -> y [n, ldy] fp32 at channel offset.  One CTA per image; a thread owns
// 8 channels (one 16-byte load per pixel) of every G-th pixel, G = 256 / (c / 8) pixel groups; the group sums meet in shared
// memory and are added in group order (deterministic).  (One thread per channel walking all pixels was a chain of hw dependent
// 2-byte loads: 73 us per 512 frames for 25 MB.)
__global__ void __launch_bounds__(256) gap_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y, int hw, int c, long long ldy) {
    extern __shared__ float gap_s[];                 // [groups][c]
    const long long img = blockIdx.x;
    const __nv_bfloat16 *px = x + img * hw * c;
    const int c8 = c >> 3;
    const int groups = max(1, 256 / c8);
    const int cc = threadIdx.x % c8, g = threadIdx.x / c8;
    if (g < groups) {
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = 0.f;
        for (int i = g; i < hw; i += groups) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(px + (long long)i * c) + cc);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[2 * j] += __uint_as_float(w4[j] << 16);
                s[2 * j + 1] += __uint_as_float(w4[j] & 0xffff0000u);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) gap_s[g * c + cc * 8 + j] = s[j];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += 256) {
        float t = 0.f;
        for (int k = 0; k < groups; ++k) t += gap_s[k * c + ch];
        y[img * ldy + ch] = t / (float)hw;
    }
}

}  // namespace



extern "C" int ewvit_mwt_upsample3_fwd(const float *hf1, const float *hf2, const float *hf3, int n, int h, int wd, void *up, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0 && h % 4 == 0 && wd % 4 == 0, EWVIT_ERR_INVALID_ARG,
                  "ewvit_mwt_upsample3_fwd: the level-1 grid must be a multiple of 4 in both directions (got %dx%d)", h, wd);
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(hf1 && hf2 && hf3 && up && ewvit_aligned16(up), EWVIT_ERR_INVALID_ARG, "ewvit_mwt_upsample3_fwd: NULL or misaligned pointer");
    EWVIT_REQUIRE(n <= (1 << 22), EWVIT_ERR_UNSUPPORTED, "ewvit_mwt_upsample3_fwd: too many frames");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    Up3Params p;
    p.hf[0] = hf1; p.hf[1] = hf2; p.hf[2] = hf3;
    int rows_out = 8;
    size_t smem = 0;
    for (;;) {
        int off = 0;
        for (int l = 0; l < 3; ++l) {
            p.hin[l] = h >> l; p.win[l] = wd >> l;
            p.ry[l] = (float)p.hin[l] / (float)h; p.rx[l] = (float)p.win[l] / (float)wd;
            p.rows_max[l] = (rows_out >> l) + 3;                      // source rows ya(oy0) .. yb(oy1 - 1) of the band
            p.off[l] = off;
            off += (9 * p.rows_max[l] * p.win[l] + 3) & ~3;          // keep every level 16-byte aligned
        }
        p.tr_off = off;
        off += 8 * 2048 / (int)sizeof(float);                        // transposition buffers: 8 warps x 32 pixels x 64 bytes
        smem = (size_t)off * sizeof(float);
        if (smem <= 96 * 1024 || rows_out == 1) break;
        rows_out >>= 1;
    }
    EWVIT_REQUIRE(smem <= 200 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_mwt_upsample3_fwd: rows of %d pixels do not fit the staging buffer", wd);
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(mwt_upsample3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int parts = (h + rows_out - 1) / rows_out;
    mwt_upsample3_kernel<<<(unsigned)((long long)n * parts), 256, smem, (cudaStream_t)stream>>>(p, static_cast<__nv_bfloat16 *>(up), h, wd, rows_out);
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
    EWVIT_REQUIRE(c % 8 == 0 && c <= 2048 && ewvit_aligned16(x), EWVIT_ERR_UNSUPPORTED, "ewvit_gap_nhwc_bf16: needs c %% 8 == 0, c <= 2048 and a 16-byte aligned input");
    const int groups = (256 / (c / 8)) > 0 ? 256 / (c / 8) : 1;
    gap_kernel<<<(unsigned)n, 256, (size_t)groups * c * sizeof(float), (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x), y, hw, c, ldy);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
