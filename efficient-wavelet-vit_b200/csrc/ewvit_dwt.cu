// Haar analysis filter bank kernels (include/ewvit.h).
//
// ewvit_dwt3_haar_fwd: three chained levels in ONE HBM pass.  HBM-bandwidth bound:
//   algorithmic bytes per 8x8 input block = 256 (read) + 256 (LL1+HF1) + 64 (LL2+HF2) + 16 (LL3+HF3)
//                                         = 592 B  ->  356 450 304 B for 256x3x224x224.
// Layout / mapping:
//   * an NCHW fp32 tensor is a 1-D array of "bands" (8 consecutive image rows of one plane,
//     32*w contiguous bytes); a tile is up to 8 consecutive bands, fetched by ONE 1-D bulk-TMA
//     copy (cp.async.bulk -> SASS UBLKCP) into a 3-stage shared-memory ring, completion on an
//     mbarrier; CTAs are persistent (one per SM) and stride over the tiles;
//   * one thread owns one 8x8 block: 16 conflict-free LDS.128, the 3-level butterfly in
//     registers, then 128-bit stores of the four level-1 subbands (a warp writes 448
//     contiguous bytes per subband row at w=224), 64-bit stores for level 2, 32-bit for level 3.
//   * arithmetic order is the oracle's (oracle/haar.py): every product and sum rounded
//     separately (__fmul_rn/__fadd_rn, no FMA contraction) -> bit-exact against the oracle.
#include "ewvit_common.cuh"

namespace {

constexpr float kS = 0.70710677f;   // fp32(1/sqrt(2)) -- the reference's filter tap
constexpr int kStages = 3;
constexpr int kMaxBandsPerTile = 8;
constexpr int kThreads = 256;
constexpr int kMaxU8Channels = 4;

struct Haar4 {
    float ll, lh, hl, hh;
};

// a b / c d -> one 2x2 Haar butterfly, the oracle's rounding order.
__device__ __forceinline__ Haar4 haar2x2(float a, float b, float c, float d) {
    const float pa = __fmul_rn(a, kS), pb = __fmul_rn(b, kS), pc = __fmul_rn(c, kS), pd = __fmul_rn(d, kS);
    const float lo_t = __fmul_rn(__fadd_rn(pa, pb), kS), hi_t = __fmul_rn(__fsub_rn(pa, pb), kS);
    const float lo_b = __fmul_rn(__fadd_rn(pc, pd), kS), hi_b = __fmul_rn(__fsub_rn(pc, pd), kS);
    Haar4 r;
    r.ll = __fadd_rn(lo_t, lo_b);
    r.lh = __fsub_rn(lo_t, lo_b);
    r.hl = __fadd_rn(hi_t, hi_b);
    r.hh = __fsub_rn(hi_t, hi_b);
    return r;
}

struct Dwt3Params {
    const void *x;           // fp32 planes, or uint8 planes when the kernel is instantiated with kU8
    const float *mean, *stdv;   // kU8: per-channel normalisation, value = ((u / 255) - mean[c]) / std[c], c = plane % channels
    int channels;
    float *ll1, *hf1, *ll2, *hf2, *ll3, *hf3;
    long long total_bands;   // planes * h/8
    int bands_per_plane;     // h/8
    int w;                   // multiple of 8
    int bw;                  // w/8 blocks per band
    int bands_per_tile;
};

// kU8: the frames arrive as uint8 (what a decoder produces) and the ToTensor + Normalize arithmetic of the reference's input
// pipeline (config/transforms.py:97-98: x/255, then (x - mean) / std, each rounded separately) happens on load -- a quarter
// of the input bytes cross PCIe and HBM.
template <bool kU8>
__global__ void __launch_bounds__(kThreads, 1) dwt3_haar_kernel(const Dwt3Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[kStages];

    // kU8: 256-entry lookup table per channel, built with exactly the reference's arithmetic (two IEEE divisions and a
    // subtraction per entry) so that the per-sample conversion is one shared-memory read and still bit-identical
    __shared__ float s_lut[kU8 ? kMaxU8Channels * 256 : 1];
    const int tid = threadIdx.x;
    if (kU8) {
        for (int i = tid; i < p.channels * 256; i += kThreads) {
            const int ch = i >> 8, u = i & 255;
            s_lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), __ldg(p.mean + ch)), __ldg(p.stdv + ch));
        }
    }
    const uint32_t band_bytes = (kU8 ? 8u : 32u) * (uint32_t)p.w;
    const uint32_t stage_bytes = band_bytes * (uint32_t)p.bands_per_tile;
    const uint32_t smem_base = ewvit::smem_u32(smem_raw);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) ewvit::mbar_init(ewvit::smem_u32(&full_bar[s]), 1);
        ewvit::mbar_fence_init();
    }
    __syncthreads();

    // contiguous, near-equal share of the bands for this CTA (imbalance <= 1 band)
    const long long share = p.total_bands / gridDim.x, rem = p.total_bands % gridDim.x;
    const long long cta_b0 = blockIdx.x * share + (blockIdx.x < rem ? blockIdx.x : rem);
    const long long cta_b1 = cta_b0 + share + (blockIdx.x < rem ? 1 : 0);
    const long long num_tiles = (cta_b1 - cta_b0 + p.bands_per_tile - 1) / p.bands_per_tile;

    auto issue = [&](long long tile, int stage) {
        const long long band0 = cta_b0 + tile * p.bands_per_tile;
        long long nb = cta_b1 - band0;
        if (nb > p.bands_per_tile) nb = p.bands_per_tile;
        const uint32_t bytes = (uint32_t)nb * band_bytes;
        const uint32_t bar = ewvit::smem_u32(&full_bar[stage]);
        ewvit::mbar_expect_tx(bar, bytes);
        ewvit::bulk_g2s(smem_base + stage * stage_bytes,
                        reinterpret_cast<const unsigned char *>(p.x) + band0 * (long long)band_bytes, bytes, bar);
    };

    if (tid == 0) {
        for (int s = 0; s < kStages - 1; ++s)
            if (s < num_tiles) issue(s, s);
    }

    const int w = p.w, w1 = w >> 1, w2 = w >> 2, w3 = w >> 3;
    const int h1 = p.bands_per_plane * 4, h2 = p.bands_per_plane * 2, h3 = p.bands_per_plane;
    const int swap = (tid >> 2) & 1;   // quarter-warp halves read opposite 16-byte halves: conflict-free LDS.128

    for (long long it = 0; it < num_tiles; ++it) {
        const int stage = (int)(it % kStages);
        // everyone is done with the stage consumed in the previous iteration -> refill it
        __syncthreads();
        if (tid == 0) {
            const long long nt = it + (kStages - 1);
            if (nt < num_tiles) {
                ewvit::fence_proxy_async();
                issue(nt, (int)((it + kStages - 1) % kStages));
            }
        }
        ewvit::mbar_wait(ewvit::smem_u32(&full_bar[stage]), (uint32_t)((it / kStages) & 1));

        const long long band0 = cta_b0 + it * p.bands_per_tile;
        long long nbl = cta_b1 - band0;
        const int nb = nbl > p.bands_per_tile ? p.bands_per_tile : (int)nbl;
        const float *tile_s = reinterpret_cast<const float *>(smem_raw + (size_t)stage * stage_bytes);

        for (int t = tid; t < nb * p.bw; t += kThreads) {
            const int bl = t / p.bw, j = t - bl * p.bw;
            const long long g = band0 + bl;
            const long long plane = g / p.bands_per_plane;
            const int bi = (int)(g - plane * p.bands_per_plane);

            // ---- 8x8 block from shared memory
            float v[8][8];
            if (kU8) {
                const unsigned char *blk8 = smem_raw + (size_t)stage * stage_bytes + ((size_t)bl * 8 * w + j * 8);
                const float *lut = s_lut + ((int)(plane % p.channels) << 8);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint2 q = *reinterpret_cast<const uint2 *>(blk8 + r * w);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[r][k] = lut[((k < 4 ? q.x : q.y) >> (8 * (k & 3))) & 0xffu];
                }
            } else {
            const float *blk = tile_s + (size_t)bl * 8 * w + j * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 *rowp = reinterpret_cast<const float4 *>(blk + r * w);
                const float4 q0 = rowp[swap], q1 = rowp[swap ^ 1];
                const float4 lo = swap ? q1 : q0, hi = swap ? q0 : q1;
                v[r][0] = lo.x; v[r][1] = lo.y; v[r][2] = lo.z; v[r][3] = lo.w;
                v[r][4] = hi.x; v[r][5] = hi.y; v[r][6] = hi.z; v[r][7] = hi.w;
            }
            }

            // ---- level 1: 4x4 coefficients per subband
            float l1[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float4 o_ll, o_lh, o_hl, o_hh;
                Haar4 c0 = haar2x2(v[2 * r][0], v[2 * r][1], v[2 * r + 1][0], v[2 * r + 1][1]);
                Haar4 c1 = haar2x2(v[2 * r][2], v[2 * r][3], v[2 * r + 1][2], v[2 * r + 1][3]);
                Haar4 c2 = haar2x2(v[2 * r][4], v[2 * r][5], v[2 * r + 1][4], v[2 * r + 1][5]);
                Haar4 c3 = haar2x2(v[2 * r][6], v[2 * r][7], v[2 * r + 1][6], v[2 * r + 1][7]);
                l1[r][0] = c0.ll; l1[r][1] = c1.ll; l1[r][2] = c2.ll; l1[r][3] = c3.ll;
                o_ll = make_float4(c0.ll, c1.ll, c2.ll, c3.ll);
                o_lh = make_float4(c0.lh, c1.lh, c2.lh, c3.lh);
                o_hl = make_float4(c0.hl, c1.hl, c2.hl, c3.hl);
                o_hh = make_float4(c0.hh, c1.hh, c2.hh, c3.hh);
                const long long row = 4LL * bi + r;
                const long long col = 4LL * j;
                if (p.ll1) *reinterpret_cast<float4 *>(p.ll1 + (plane * h1 + row) * w1 + col) = o_ll;
                if (p.hf1) {
                    float *base = p.hf1 + ((plane * 3) * h1 + row) * w1 + col;
                    const long long sb = (long long)h1 * w1;
                    *reinterpret_cast<float4 *>(base) = o_lh;
                    *reinterpret_cast<float4 *>(base + sb) = o_hl;
                    *reinterpret_cast<float4 *>(base + 2 * sb) = o_hh;
                }
            }

            // ---- level 2: 2x2 coefficients per subband
            float l2[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                Haar4 c0 = haar2x2(l1[2 * r][0], l1[2 * r][1], l1[2 * r + 1][0], l1[2 * r + 1][1]);
                Haar4 c1 = haar2x2(l1[2 * r][2], l1[2 * r][3], l1[2 * r + 1][2], l1[2 * r + 1][3]);
                l2[r][0] = c0.ll; l2[r][1] = c1.ll;
                const long long row = 2LL * bi + r;
                const long long col = 2LL * j;
                if (p.ll2) *reinterpret_cast<float2 *>(p.ll2 + (plane * h2 + row) * w2 + col) = make_float2(c0.ll, c1.ll);
                if (p.hf2) {
                    float *base = p.hf2 + ((plane * 3) * h2 + row) * w2 + col;
                    const long long sb = (long long)h2 * w2;
                    *reinterpret_cast<float2 *>(base) = make_float2(c0.lh, c1.lh);
                    *reinterpret_cast<float2 *>(base + sb) = make_float2(c0.hl, c1.hl);
                    *reinterpret_cast<float2 *>(base + 2 * sb) = make_float2(c0.hh, c1.hh);
                }
            }

            // ---- level 3: one coefficient per subband
            {
                Haar4 c = haar2x2(l2[0][0], l2[0][1], l2[1][0], l2[1][1]);
                if (p.ll3) p.ll3[(plane * h3 + bi) * w3 + j] = c.ll;
                if (p.hf3) {
                    float *base = p.hf3 + ((plane * 3) * h3 + bi) * w3 + j;
                    const long long sb = (long long)h3 * w3;
                    base[0] = c.lh;
                    base[sb] = c.hl;
                    base[2 * sb] = c.hh;
                }
            }
        }
    }
}

// One level, any size, zero-mode boundary.  Edge-case path (odd sizes, tiny images): one thread
// per output coefficient, consecutive threads on consecutive output columns.
__global__ void dwt1_haar_kernel(const float *__restrict__ x, long long planes, int h, int w, int h2, int w2,
                                 float *__restrict__ ll, float *__restrict__ yh) {
    const long long total = planes * h2 * w2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % w2);
        const long long t = idx / w2;
        const int i = (int)(t % h2);
        const long long plane = t / h2;
        const float *px = x + plane * h * w;
        const int r0 = 2 * i, r1 = 2 * i + 1, c0 = 2 * j, c1 = 2 * j + 1;
        const bool rb = r1 < h, cb = c1 < w;
        const float a = px[(long long)r0 * w + c0];
        const float b = cb ? px[(long long)r0 * w + c1] : 0.f;
        const float c = rb ? px[(long long)r1 * w + c0] : 0.f;
        const float d = (rb && cb) ? px[(long long)r1 * w + c1] : 0.f;
        const Haar4 o = haar2x2(a, b, c, d);
        ll[idx] = o.ll;
        float *py = yh + ((plane * 3) * h2 + i) * w2 + j;
        const long long sb = (long long)h2 * w2;
        py[0] = o.lh;
        py[sb] = o.hl;
        py[2 * sb] = o.hh;
    }
}

}  // namespace

extern "C" int ewvit_dwt_haar_fwd(const float *x, int64_t planes, int h, int w, float *ll, float *yh,
                                  void *stream) {
    EWVIT_REQUIRE(planes >= 0 && h >= 0 && w >= 0, EWVIT_ERR_INVALID_ARG,
                  "ewvit_dwt_haar_fwd: negative size (planes=%lld h=%d w=%d)", (long long)planes, h, w);
    if (planes == 0 || h == 0 || w == 0) return EWVIT_OK;   // empty input: nothing to do
    EWVIT_REQUIRE(x && ll && yh, EWVIT_ERR_INVALID_ARG, "ewvit_dwt_haar_fwd: NULL pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
    const long long total = (long long)planes * h2 * w2;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)ewvit_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    dwt1_haar_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(x, planes, h, w, h2, w2, ll, yh);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

static int dwt3_impl(const void *x, bool u8, const float *mean, const float *stdv, int channels, int64_t planes, int h, int w, float *ll1,
                     float *hf1, float *ll2, float *hf2, float *ll3, float *hf3, void *stream) {
    EWVIT_REQUIRE(planes >= 0 && h >= 0 && w >= 0, EWVIT_ERR_INVALID_ARG,
                  "ewvit_dwt3_haar_fwd: negative size (planes=%lld h=%d w=%d)", (long long)planes, h, w);
    EWVIT_REQUIRE(h % 8 == 0 && w % 8 == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_dwt3_haar_fwd: h and w must be multiples of 8 (got %dx%d); use three "
                  "ewvit_dwt_haar_fwd calls for ragged sizes", h, w);
    EWVIT_REQUIRE(w <= 2048, EWVIT_ERR_UNSUPPORTED, "ewvit_dwt3_haar_fwd: w=%d > 2048", w);
    if (planes == 0 || h == 0 || w == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x, EWVIT_ERR_INVALID_ARG, "ewvit_dwt3_haar_fwd: x is NULL");
    EWVIT_REQUIRE(ewvit_aligned16(x) && ewvit_aligned16(ll1) && ewvit_aligned16(hf1) && ewvit_aligned16(ll2) &&
                      ewvit_aligned16(hf2) && ewvit_aligned16(ll3) && ewvit_aligned16(hf3),
                  EWVIT_ERR_INVALID_ARG, "ewvit_dwt3_haar_fwd: pointers must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;

    Dwt3Params p;
    p.x = x; p.ll1 = ll1; p.hf1 = hf1; p.ll2 = ll2; p.hf2 = hf2; p.ll3 = ll3; p.hf3 = hf3;
    p.mean = mean; p.stdv = stdv; p.channels = channels > 0 ? channels : 1;
    p.bands_per_plane = h / 8;
    p.total_bands = (long long)planes * p.bands_per_plane;
    p.w = w;
    p.bw = w / 8;
    const int band_bytes = (u8 ? 8 : 32) * w;
    int bpt = (64 * 1024) / band_bytes;           // <= 64 KiB per stage
    if (bpt > kMaxBandsPerTile) bpt = kMaxBandsPerTile;
    if (bpt < 1) bpt = 1;
    p.bands_per_tile = bpt;
    const size_t smem = (size_t)kStages * bpt * band_bytes;

    static bool attr_set[64] = {false};   // per device; benign race (same value every time)
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwt3_haar_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwt3_haar_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    long long grid = ewvit_num_sms();
    if (grid > p.total_bands) grid = p.total_bands;
    if (u8) dwt3_haar_kernel<true><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    else dwt3_haar_kernel<false><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_dwt3_haar_fwd(const float *x, int64_t planes, int h, int w, float *ll1, float *hf1,
                                   float *ll2, float *hf2, float *ll3, float *hf3, void *stream) {
    return dwt3_impl(x, false, nullptr, nullptr, 1, planes, h, w, ll1, hf1, ll2, hf2, ll3, hf3, stream);
}

extern "C" int ewvit_dwt3_haar_u8_fwd(const uint8_t *x, const float *mean, const float *stdv, int channels, int64_t planes, int h, int w,
                                      float *ll1, float *hf1, float *ll2, float *hf2, float *ll3, float *hf3, void *stream) {
    EWVIT_REQUIRE(mean && stdv && channels > 0 && channels <= kMaxU8Channels, EWVIT_ERR_INVALID_ARG,
                  "ewvit_dwt3_haar_u8_fwd: needs mean/std and 1..4 channels");
    EWVIT_REQUIRE((8 * w) % 16 == 0, EWVIT_ERR_UNSUPPORTED, "ewvit_dwt3_haar_u8_fwd: w must be even");
    return dwt3_impl(x, true, mean, stdv, channels, planes, h, w, ll1, hf1, ll2, hf2, ll3, hf3, stream);
}
