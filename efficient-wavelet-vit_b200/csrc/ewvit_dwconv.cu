// Depthwise k x k convolution (k = 3 | 5, stride 1 | 2, arbitrary top/left zero padding) + bias + SiLU on NHWC bf16, with the
// squeeze-excitation mean as a by-product: the MBConv middle step of EfficientNetV2-S (torchvision, symmetric pad 1;
// network/sfe.py:111-113,150) and of EfficientNet-b0 (TensorFlow 'SAME' padding, 3x3 and 5x5; network/sfe.py:109,148).
//
// HBM-bound (9..25 FMAs per 2-byte element): the kernel is organised so that EVERY global access is a full 128-byte line
// and nothing is staged in shared memory.
//   lane   = one channel pair (one 32-bit load/store per pixel; a warp covers 64 consecutive channels = 128 bytes),
//   warp   = one unit (frame, 64-channel slab, band of output rows, segment of WS output columns),
//   thread = a k-row sliding window of the WIN = (WS-1)*S + k input columns of its segment, kept PACKED (bf16x2) in
//            registers; every output row shifts the window by S rows and loads S new rows (requested before the FMAs of
//            the current row are issued, so the loads of row r+1 fly under the arithmetic of row r).
// Per output row a thread issues S*WIN independent 4-byte loads (9..22), 2*k*k*WS FMAs and 2*WS MUFU.TANH.
// Measured (512 frames, B200): this kernel is bound by instruction issue (~17 instructions per output), not by HBM --
// 14x14x960 stride 1: 145 us against 110 us of the TMA-staged 3x3 kernel of ewvit_backbone.cu, which therefore keeps
// serving the V2-S layers; this one serves what that kernel cannot (5x5, TensorFlow-SAME padding, c % 64 != 0: the b0).
// The per-(frame, channel) means of the outputs (SE squeeze) are accumulated per thread and written as one partial per
// (band, segment): pooled[frame][part][channel], already scaled by 1/(ho*wo), summed by the gate kernel in a fixed order.
#include "ewvit_common.cuh"

namespace {

template <int K, int S, int WS, int MINB>
__global__ void __launch_bounds__(256, MINB) dwconv_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ w,
                                                     const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y,
                                                     float *__restrict__ pooled, int n, int h, int wd, int ho, int wo, int c,
                                                     int pad_t, int pad_l, int rb, int nbands, int nsegs, int nslab, int act,
                                                     long long units) {
    constexpr int WIN = (WS - 1) * S + K;
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (unit >= units) return;
    // unit -> (frame, band, segment, slab); slab fastest so that the warps of a block touch neighbouring lines
    const int slab = (int)(unit % nslab);
    long long t = unit / nslab;
    const int seg = (int)(t % nsegs);
    t /= nsegs;
    const int band = (int)(t % nbands);
    const int f = (int)(t / nbands);
    const int c0 = slab * 64 + lane * 2;
    if (c0 >= c) return;                       // channel tail: no warp-level primitives below

    // weights and bias of this lane's two channels; SiLU is evaluated as h*tanh(h) + h with h = v/2, so halve them here
    const float hs = act == 4 ? 0.5f : 1.f;
    float wr[K * K][2], br[2];
#pragma unroll
    for (int k = 0; k < K * K; ++k) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(w + (long long)k * c + c0));
        wr[k][0] = v.x * hs;
        wr[k][1] = v.y * hs;
    }
    {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(bias + c0));
        br[0] = v.x * hs;
        br[1] = v.y * hs;
    }

    const int ox0 = seg * WS;
    const int ix0 = ox0 * S - pad_l;           // input column of window slot 0
    const int oy0 = band * rb, oy1 = min(ho, oy0 + rb);
    const uint32_t *xf = reinterpret_cast<const uint32_t *>(x + (long long)f * h * wd * c + c0);   // channel pair as one word
    const long long pix = c >> 1;              // words per pixel

    // column validity is the same for every row
    bool colok[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) colok[j] = (unsigned)(ix0 + j) < (unsigned)wd;

    auto load_row = [&](int iy, uint32_t (&r)[WIN]) {
        const bool rowok = (unsigned)iy < (unsigned)h;
        const uint32_t *rp = xf + ((long long)iy * wd + ix0) * pix;
#pragma unroll
        for (int j = 0; j < WIN; ++j) r[j] = (rowok && colok[j]) ? __ldg(rp + j * pix) : 0u;
    };

    uint32_t win[K][WIN];
    int iy = oy0 * S - pad_t;                  // input row held in win[0]
#pragma unroll
    for (int k = 0; k < K; ++k) load_row(iy + k, win[k]);

    float sum0 = 0.f, sum1 = 0.f;
    __nv_bfloat16 *yp = y + (((long long)f * ho + oy0) * wo + ox0) * c + c0;
    for (int oy = oy0; oy < oy1; ++oy) {
        // request the S rows the NEXT output row adds before touching the FMAs of this one
        uint32_t nxt[S][WIN];
        if (oy + 1 < oy1) {
#pragma unroll
            for (int s = 0; s < S; ++s) load_row(iy + K + s, nxt[s]);
        } else {
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int j = 0; j < WIN; ++j) nxt[s][j] = 0u;
        }
        float acc[WS][2];
#pragma unroll
        for (int o = 0; o < WS; ++o) { acc[o][0] = br[0]; acc[o][1] = br[1]; }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                const float x0 = __uint_as_float(win[ky][j] << 16), x1 = __uint_as_float(win[ky][j] & 0xffff0000u);
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    // window slot j feeds output column o = (j - kx) / S when that is an integer inside the segment
                    if ((j - kx) >= 0 && (j - kx) % S == 0 && (j - kx) / S < WS) {
                        acc[(j - kx) / S][0] = fmaf(wr[ky * K + kx][0], x0, acc[(j - kx) / S][0]);
                        acc[(j - kx) / S][1] = fmaf(wr[ky * K + kx][1], x1, acc[(j - kx) / S][1]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < WS; ++o) {
            float a0 = acc[o][0], a1 = acc[o][1];
            if (act == 4) {
                float t0, t1;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
                a0 = fmaf(a0, t0, a0);
                a1 = fmaf(a1, t1, a1);
            }
            if (ox0 + o < wo) {
                sum0 += a0;
                sum1 += a1;
                const __nv_bfloat162 pk = __floats2bfloat162_rn(a0, a1);
                *reinterpret_cast<__nv_bfloat162 *>(yp + (long long)o * c) = pk;
            }
        }
        yp += (long long)wo * c;
        // slide the window down by S rows
#pragma unroll
        for (int k = 0; k + S < K; ++k)
#pragma unroll
            for (int j = 0; j < WIN; ++j) win[k][j] = win[k + S][j];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < WIN; ++j) win[K - S + s][j] = nxt[s][j];
        iy += S;
    }
    if (pooled) {
        const float inv = 1.f / (float)(ho * wo);
        const int part = band * nsegs + seg;
        *reinterpret_cast<float2 *>(pooled + ((long long)f * (nbands * nsegs) + part) * c + c0) = make_float2(sum0 * inv, sum1 * inv);
    }
}

struct DwPlan {
    int ws, rb, nbands, nsegs;
};

// The decomposition is a pure function of the output size, kernel and stride (callers size `pooled` from it).
DwPlan dw_plan(int ho, int wo, int ksize, int stride) {
    DwPlan p;
    p.ws = (ksize == 5 && stride == 2) ? 4 : 7;        // 14-column segments measured 10% slower than 7 (206 registers)
    p.nsegs = (wo + p.ws - 1) / p.ws;
    p.rb = ho <= 28 ? ho : 28;
    p.nbands = (ho + p.rb - 1) / p.rb;
    return p;
}

template <int K, int S, int WS, int MINB>
int launch(const void *x, const float *w, const float *bias, void *y, float *pooled, int n, int h, int wd, int ho, int wo, int c,
           int pad_t, int pad_l, const DwPlan &pl, int act, cudaStream_t stream) {
    const int nslab = (c + 63) / 64;
    const long long units = (long long)n * pl.nbands * pl.nsegs * nslab;
    const long long blocks = (units + 7) / 8;
    EWVIT_REQUIRE(blocks < (1LL << 31), EWVIT_ERR_UNSUPPORTED, "ewvit_dwconv_nhwc_bf16: too many units");
    dwconv_kernel<K, S, WS, MINB><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(x), w, bias, static_cast<__nv_bfloat16 *>(y),
                                                                  pooled, n, h, wd, ho, wo, c, pad_t, pad_l, pl.rb, pl.nbands, pl.nsegs, nslab,
                                                                  act, units);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

}  // namespace

extern "C" int ewvit_dwconv_pool_parts(int ho, int wo, int ksize, int stride) {
    if (ho <= 0 || wo <= 0 || (ksize != 3 && ksize != 5) || (stride != 1 && stride != 2)) return -1;
    const DwPlan p = dw_plan(ho, wo, ksize, stride);
    return p.nbands * p.nsegs;
}

extern "C" int ewvit_dwconv_nhwc_bf16(const void *x, const float *w, const float *bias, int n, int h, int wd, int c, int ksize,
                                      int stride, int pad_top, int pad_left, int ho, int wo, int act, void *y, float *pooled,
                                      void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0 && c > 0 && ho > 0 && wo > 0, EWVIT_ERR_INVALID_ARG, "ewvit_dwconv_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && bias && y, EWVIT_ERR_INVALID_ARG, "ewvit_dwconv_nhwc_bf16: NULL pointer");
    EWVIT_REQUIRE((ksize == 3 || ksize == 5) && (stride == 1 || stride == 2), EWVIT_ERR_UNSUPPORTED,
                  "ewvit_dwconv_nhwc_bf16: supports 3x3 / 5x5, stride 1 | 2 (got k=%d s=%d)", ksize, stride);
    EWVIT_REQUIRE(c % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3u) == 0 && (reinterpret_cast<uintptr_t>(y) & 3u) == 0 &&
                      (reinterpret_cast<uintptr_t>(w) & 7u) == 0 && (reinterpret_cast<uintptr_t>(bias) & 7u) == 0 &&
                      (reinterpret_cast<uintptr_t>(pooled) & 7u) == 0,
                  EWVIT_ERR_INVALID_ARG, "ewvit_dwconv_nhwc_bf16: needs an even channel count and 4/8-byte aligned pointers");
    EWVIT_REQUIRE(pad_top >= 0 && pad_top < ksize && pad_left >= 0 && pad_left < ksize, EWVIT_ERR_INVALID_ARG, "ewvit_dwconv_nhwc_bf16: bad padding");
    EWVIT_REQUIRE((ho - 1) * stride - pad_top < h && (wo - 1) * stride - pad_left < wd, EWVIT_ERR_INVALID_ARG,
                  "ewvit_dwconv_nhwc_bf16: output %dx%d reaches past the %dx%d input", ho, wo, h, wd);
    EWVIT_REQUIRE(act == 0 || act == 4, EWVIT_ERR_INVALID_ARG, "ewvit_dwconv_nhwc_bf16: act must be 0 (none) or 4 (SiLU)");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const DwPlan pl = dw_plan(ho, wo, ksize, stride);
    cudaStream_t st = (cudaStream_t)stream;
#define EWVIT_DW(K_, S_, WS_, B_) launch<K_, S_, WS_, B_>(x, w, bias, y, pooled, n, h, wd, ho, wo, c, pad_top, pad_left, pl, act, st)
    // register budgets measured on B200 (tools/dw_bench.py): variants that spill (3 or 4 blocks per SM) are 1.3-2x slower
    if (ksize == 3 && stride == 1) return EWVIT_DW(3, 1, 7, 2);
    if (ksize == 3) return EWVIT_DW(3, 2, 7, 2);
    if (stride == 1) return EWVIT_DW(5, 1, 7, 1);
    return EWVIT_DW(5, 2, 4, 1);
#undef EWVIT_DW
}
