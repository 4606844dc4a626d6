// EfficientNet feature-extractor glue kernels: everything of a FusedMBConv /
// MBConv block that is not a dense contraction.  The dense parts (3x3 FusedMBConv convs, 1x1 expand/project convs,
// the 1x1 head) run on the tcgen05 implicit-GEMM kernel (ewvit_conv_nhwc_bf16) with bias + SiLU + residual fused
// into the epilogue, so an activation tensor is written once and read once -- the eager cuDNN path spends >50 % of
// its device time in separate elementwise passes (profiles/r01_bench_launches.md).  All tensors NHWC bf16,
// BatchNorm (eval) folded into the weights/bias on the host.
#include "ewvit_tc.cuh"

namespace {

__device__ __forceinline__ float silu(float x) { return ewvit::silu_fast(x); }

// ---- stem: Conv2d(3 -> cout<=32, 3x3, stride 2, pad 1) + bias + SiLU, fp32 NCHW frames -> bf16 NHWC
//      (also the fp32 -> bf16 conversion of the input; torchvision features[0], sfe.py:150)
constexpr int kStemMaxC = 32;
constexpr int kStemPx = 4;      // output pixels per thread (consecutive columns): every weight read from shared memory feeds 4 FMAs
// TIn = float: normalised fp32 frames.  TIn = unsigned char: raw frames, ((u / 255) - mean[c]) / std[c] applied on load
// (config/transforms.py:97-98), zero padding applied AFTER the normalisation as in the reference.
template <typename TIn>
__global__ void __launch_bounds__(128) stem_conv_kernel(const TIn *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y,
                                                        int n, int h, int wd, int ho, int wo, int cout, int pad,
                                                        const float *__restrict__ mean, const float *__restrict__ stdv, int in_pad) {
    // weights transposed to [27][cout] so that 4 consecutive output channels are one 16-byte shared-memory read;
    // pre-halved for the h*tanh(h)+h form of SiLU
    __shared__ __align__(16) float s_w[27 * kStemMaxC];
    __shared__ __align__(16) float s_b[kStemMaxC];
    __shared__ float s_lut[sizeof(TIn) == 1 ? 3 * 256 : 1];      // uint8 frames: per-channel value table, the reference's arithmetic
    if (sizeof(TIn) == 1) {
        for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x)
            s_lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.f), __ldg(mean + (i >> 8))), __ldg(stdv + (i >> 8)));
    }
    for (int i = threadIdx.x; i < cout * 27; i += blockDim.x) s_w[(i % 27) * kStemMaxC + i / 27] = 0.5f * w[i];
    if (threadIdx.x < cout) s_b[threadIdx.x] = 0.5f * bias[threadIdx.x];
    __syncthreads();
    const int wq = (wo + kStemPx - 1) / kStemPx;                 // pixel quads per output row
    const long long total = (long long)n * ho * wq;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int qx = (int)(idx % wq);
        const long long t = idx / wq;
        const int oy = (int)(t % ho);
        const long long img = t / ho;
        const int ox0 = qx * kStemPx;
        // input patch: 3 channels x 3 rows x 9 columns (columns 2*ox0 - in_pad .. 2*ox0 - in_pad + 8); in_pad = 1: torchvision's
        // symmetric padding, in_pad = 0: TensorFlow 'SAME' on an even size (the extra zero row/column sits at the bottom/right)
        float in[3][3][2 * kStemPx + 1];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int iy = 2 * oy + dy - in_pad;
                const TIn *row = x + ((img * 3 + c) * h + (iy >= 0 && iy < h ? iy : 0)) * (long long)wd;
#pragma unroll
                for (int j = 0; j < 2 * kStemPx + 1; ++j) {
                    const int ix = 2 * ox0 + j - in_pad;
                    float v = 0.f;
                    if (iy >= 0 && iy < h && ix >= 0 && ix < wd) {
                        if (sizeof(TIn) == 1) v = s_lut[c * 256 + (int)__ldg(row + ix)];
                        else v = (float)__ldg(row + ix);
                    }
                    in[c][dy][j] = v;
                }
            }
        // pad = 1: padded-flat output [n, ho+2, wo+2, cout] (interior written, the zero border belongs to the caller)
        __nv_bfloat16 *py = y + ((img * (ho + 2 * pad) + oy + pad) * (wo + 2 * pad) + ox0 + pad) * cout;
        for (int c0 = 0; c0 < cout; c0 += 8) {
            // packed fp32x2 FMAs (FFMA2: each half rounds exactly like fmaf): a three-register FFMA issues every other cycle on
            // sm_100, so the 648 FMAs per output pixel bound this kernel at half the fp32 rate
            float2 acc[kStemPx][4];
#pragma unroll
            for (int p4 = 0; p4 < kStemPx; ++p4)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[p4][j] = make_float2(s_b[c0 + 2 * j], s_b[c0 + 2 * j + 1]);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int k = c * 9 + dy * 3 + dx;
                        const float4 w0 = *reinterpret_cast<const float4 *>(&s_w[k * kStemMaxC + c0]);
                        const float4 w1 = *reinterpret_cast<const float4 *>(&s_w[k * kStemMaxC + c0 + 4]);
                        const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w);
                        const float2 wc = make_float2(w1.x, w1.y), wd2 = make_float2(w1.z, w1.w);
#pragma unroll
                        for (int p4 = 0; p4 < kStemPx; ++p4) {
                            const float v = in[c][dy][2 * p4 + dx];
                            const float2 vv = make_float2(v, v);
                            acc[p4][0] = __ffma2_rn(wa, vv, acc[p4][0]);
                            acc[p4][1] = __ffma2_rn(wb, vv, acc[p4][1]);
                            acc[p4][2] = __ffma2_rn(wc, vv, acc[p4][2]);
                            acc[p4][3] = __ffma2_rn(wd2, vv, acc[p4][3]);
                        }
                    }
#pragma unroll
            for (int p4 = 0; p4 < kStemPx; ++p4) {
                if (ox0 + p4 >= wo) break;
                float o[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 th;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(acc[p4][j].x));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(acc[p4][j].y));
                    const float2 r = __ffma2_rn(acc[p4][j], th, acc[p4][j]);
                    o[2 * j] = r.x;
                    o[2 * j + 1] = r.y;
                }
                uint4 pk;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(o[0], o[1]), b1 = __floats2bfloat162_rn(o[2], o[3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(o[4], o[5]), b3 = __floats2bfloat162_rn(o[6], o[7]);
                pk.x = *reinterpret_cast<uint32_t *>(&b0);
                pk.y = *reinterpret_cast<uint32_t *>(&b1);
                pk.z = *reinterpret_cast<uint32_t *>(&b2);
                pk.w = *reinterpret_cast<uint32_t *>(&b3);
                *reinterpret_cast<uint4 *>(py + p4 * cout + c0) = pk;
            }
        }
    }
}

// ---- depthwise 3x3 (stride 1|2, pad 1) + bias + SiLU, and the squeeze (spatial mean) of the result.
//      Work unit = (channel slab of SC = 64|32 channels, pass of `fb` consecutive frames).  Persistent CTAs (two per
//      SM) each take a contiguous range of units, slab-major, so the grid has no partial last wave and a CTA reloads
//      its weights at most once or twice.  One TMA box per unit brings the slab's input planes INCLUDING the one-pixel
//      zero border into shared memory (out-of-range box coordinates are zero-filled by the TMA unit), double-buffered
//      so the next unit streams in under the compute of the current one.  A work item = (frame, output column,
//      4-channel group): the thread walks down its column with a sliding 3x3 window of fp32 registers, so every staged
//      value is converted once per column (3 conversions per output instead of 9), the loop has no bounds checks or
//      index divisions, and the 36 weights stay in registers.  The squeeze needs no atomics: column sums meet in
//      shared memory.  SiLU(v) = h*tanh(h) + h with h = v/2: the 1/2 is folded into the weights and bias.
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int SC, int STRIDE>
__global__ void __launch_bounds__(256, 2) dwconv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const float *__restrict__ w,
                                                           const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y,
                                                           float *__restrict__ pooled, int n, int h, int wd, int ho, int wo,
                                                           int c, int fb, int stage_bytes) {
    constexpr int G = SC / 4;            // 4-channel groups per slab
    extern __shared__ __align__(128) unsigned char dw_smem[];
    // layout: [2 stages][fb][h+2][wd+2][SC] bf16 | weights [9][SC] + bias [SC] fp32 | column sums [2][fb][wo][SC] fp32 | 2 mbarriers
    const uint32_t s_in = ewvit::smem_u32(dw_smem);
    float *s_w = reinterpret_cast<float *>(dw_smem + (size_t)2 * stage_bytes);
    float *s_sum = s_w + 10 * SC;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_sum + (size_t)2 * fb * wo * SC);
    const int tid = threadIdx.x;
    const int slabs = c / SC;
    const int passes = (n + fb - 1) / fb;
    const long long units = (long long)slabs * passes;
    const int u0 = (int)(units * blockIdx.x / gridDim.x), u1 = (int)(units * (blockIdx.x + 1) / gridDim.x);
    const int npix = ho * wo;
    const int wp = wd + 2;
    const int plane_bytes = (h + 2) * wp * SC * 2;
    const uint32_t row_bytes = (uint32_t)wp * SC * 2;

    if (tid == 0) {
        if (s_in & 127u) __trap();          // TMA destination alignment
        ewvit::mbar_init(ewvit::smem_u32(&s_bar[0]), 1);
        ewvit::mbar_init(ewvit::smem_u32(&s_bar[1]), 1);
        ewvit::mbar_fence_init();
        ewvit::fence_proxy_async();
        if (u0 < u1) {
            ewvit::mbar_expect_tx(ewvit::smem_u32(&s_bar[0]), (uint32_t)fb * plane_bytes);
            ewvit::tma_load_4d(s_in, &tmX, (u0 / passes) * SC, -1, -1, (u0 % passes) * fb, ewvit::smem_u32(&s_bar[0]));
        }
    }
    __syncthreads();

    const int items = fb * wo * G;
    const float inv_npix = 1.f / (float)npix;
    const long long ystep = (long long)wo * c;
    int buf = 0, cur_slab = -1;
    uint32_t phases = 0u;
    for (int u = u0; u < u1; ++u, buf ^= 1) {
        const int slab = u / passes;
        const int f = (u - slab * passes) * fb;
        const int cs = slab * SC;
        if (tid == 0 && u + 1 < u1) {      // the other buffer was released by the barrier that ended the previous unit
            const uint32_t bar = ewvit::smem_u32(&s_bar[buf ^ 1]);
            ewvit::mbar_expect_tx(bar, (uint32_t)fb * plane_bytes);
            ewvit::tma_load_4d(s_in + (uint32_t)(buf ^ 1) * stage_bytes, &tmX, ((u + 1) / passes) * SC, -1, -1, ((u + 1) % passes) * fb, bar);
        }
        if (slab != cur_slab) {             // CTA-uniform; everyone left the previous unit's item loop at its closing barrier
            for (int i = tid; i < 9 * SC; i += 256) s_w[i] = 0.5f * w[(i / SC) * c + cs + (i % SC)];
            if (tid < SC) s_w[9 * SC + tid] = 0.5f * bias[cs + tid];
            cur_slab = slab;
            __syncthreads();
        }
        ewvit::mbar_wait(ewvit::smem_u32(&s_bar[buf]), (phases >> buf) & 1u);
        phases ^= 1u << buf;
        const uint32_t sb = s_in + (uint32_t)buf * stage_bytes;
        float *ssum = s_sum + (size_t)buf * fb * wo * SC;
        for (int item = tid; item < items; item += 256) {
            const int g = item % G;
            const int t = item / G;
            const int ox = t % wo, fl = t / wo;
            // channel pairs as float2: packed fp32x2 FMAs (FFMA2, each half rounds exactly like fmaf) -- a three-register FFMA issues
            // every other cycle on sm_100, and 36 of them per 4-channel output made the FMA pipe a co-limiter of this kernel
            float2 wr[9][2], br[2], sum[2];
            const uint32_t wbase = ewvit::smem_u32(s_w) + g * 16;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 v = lds128f(wbase + k * SC * 4);
                wr[k][0] = make_float2(v.x, v.y); wr[k][1] = make_float2(v.z, v.w);
            }
            {
                const float4 v = lds128f(wbase + 9 * SC * 4);
                br[0] = make_float2(v.x, v.y); br[1] = make_float2(v.z, v.w);
            }
            sum[0] = sum[1] = make_float2(0.f, 0.f);
            uint32_t rp = sb + (uint32_t)fl * plane_bytes + (uint32_t)(ox * STRIDE) * (SC * 2) + g * 8;   // padded (row 0, first tap column)
            const bool live = f + fl < n;
            __nv_bfloat16 *py = y + ((long long)(f + fl) * npix + ox) * c + cs + g * 4;
            auto load_row = [&](float2 (&r)[3][2]) {      // the next padded row of this column's three taps
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint2 v = lds64(rp + dx * (SC * 2));
                    r[dx][0] = make_float2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
                    r[dx][1] = make_float2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
                }
                rp += row_bytes;
            };
            auto emit = [&](const float2 (&r0)[3][2], const float2 (&r1)[3][2], const float2 (&r2)[3][2]) {
                float2 a[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float2 hv = br[j];                  // h = v/2 (weights and bias are pre-halved)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        hv = __ffma2_rn(wr[dx][j], r0[dx][j], hv);
                        hv = __ffma2_rn(wr[3 + dx][j], r1[dx][j], hv);
                        hv = __ffma2_rn(wr[6 + dx][j], r2[dx][j], hv);
                    }
                    float2 th;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(hv.x));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(hv.y));
                    a[j] = __ffma2_rn(hv, th, hv);
                    sum[j] = __fadd2_rn(sum[j], a[j]);
                }
                if (live) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(a[0].x, a[0].y), p1 = __floats2bfloat162_rn(a[1].x, a[1].y);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t *>(&p0);
                    pk.y = *reinterpret_cast<const uint32_t *>(&p1);
                    *reinterpret_cast<uint2 *>(py) = pk;
                }
                py += ystep;
            };
            float2 ra[3][2], rb[3][2], rc[3][2];
            if (STRIDE == 1) {
                load_row(ra);
                load_row(rb);
                for (int oy = 0; oy < ho; oy += 3) {
                    load_row(rc);
                    emit(ra, rb, rc);
                    if (oy + 1 < ho) { load_row(ra); emit(rb, rc, ra); }
                    if (oy + 2 < ho) { load_row(rb); emit(rc, ra, rb); }
                }
            } else {
                load_row(ra);
                for (int oy = 0; oy < ho; oy += 2) {
                    load_row(rb);
                    load_row(rc);
                    emit(ra, rb, rc);
                    if (oy + 1 < ho) { load_row(rb); load_row(ra); emit(rc, rb, ra); }
                }
            }
            *reinterpret_cast<float4 *>(ssum + (fl * wo + ox) * SC + g * 4) = make_float4(sum[0].x, sum[0].y, sum[1].x, sum[1].y);
        }
        __syncthreads();    // column sums are complete and everyone is done reading this stage buffer
        if (pooled) {
            for (int i = tid; i < fb * SC; i += 256) {
                const int fl = i / SC, ch = i - fl * SC;
                if (f + fl >= n) continue;
                float t = 0.f;
                for (int k = 0; k < wo; ++k) t += ssum[(fl * wo + k) * SC + ch];
                pooled[(long long)(f + fl) * c + cs + ch] = t * inv_npix;
            }
        }
    }
}

// ---- 3x3 / stride 1 / pad 1 convolution with 24 input and 24 output channels + bias + SiLU (+ residual = the input):
//      the two stage-1 FusedMBConv blocks of EfficientNetV2-S at 112x112.  With N = 24 a 128-row tcgen05 tile is all
//      fixed cost (the assembled-A path spent ~2600 cycles per 128x24 tile building operands), so these layers use
//      warp-level mma.sync.m16n8k16 on a direct-convolution tile instead: a CTA owns 16x16 output pixels, the 18x18x24
//      input halo sits in shared memory (80-byte pixel pitch: conflict-free ldmatrix rows), and an A fragment of tap
//      (dy, dx) is simply ldmatrix on the pixel rows shifted by that tap -- no im2col copy exists anywhere.  K = 216
//      dense (27 chunks of 8 channels, chunk 27 reads zeros), weights [24][K] stay in shared memory.
constexpr int kC24 = 24;
constexpr int kC24Tile = 16;
constexpr int kC24Halo = kC24Tile + 2;                 // 18
constexpr int kC24Pitch = 80;                           // bytes per halo pixel (24 channels = 48 bytes + padding)
constexpr int kC24HaloBytes = kC24Halo * kC24Halo * kC24Pitch;   // 25920
constexpr int kC24KSteps = 14;                          // 28 chunks of 8 channels (27 real)
constexpr int kC24WPitch = 464;                         // bytes per weight row: 224 elements + padding (conflict-free ldmatrix)
constexpr int kC24Smem = 2 * kC24HaloBytes + kC24 * kC24WPitch + 128 /*zeros*/ + 128;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 3) conv3x3_c24_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ w, int wk,
                                                             const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y, int n, int h,
                                                             int wd, int residual) {
    extern __shared__ __align__(128) unsigned char c24_smem[];
    const uint32_t s_halo = ewvit::smem_u32(c24_smem);                          // [2][18*18][80 B]
    const uint32_t s_w = s_halo + 2 * kC24HaloBytes;                            // [24][464 B]
    const uint32_t s_zero = s_w + kC24 * kC24WPitch;                            // 128 B of zeros
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (wd + kC24Tile - 1) / kC24Tile, tiles_y = (h + kC24Tile - 1) / kC24Tile;
    const long long tiles = (long long)n * tiles_y * tiles_x;

    // weights (K-major rows, dense k = tap*24 + c) and the zero block
    for (int i = tid; i < kC24 * 28; i += 256) {
        const int row = i / 28, q = i - row * 28;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q * 8 < wk) v = __ldg(reinterpret_cast<const uint4 *>(w + (long long)row * wk + q * 8));
        if (q == 27) v = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s_w + row * kC24WPitch + q * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    if (tid < 8) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(s_zero + tid * 16), "r"(0u) : "memory");

    auto stage = [&](long long t, int buf) {       // halo of tile t -> buffer buf (out-of-image pixels zero-filled)
        const int tx = (int)(t % tiles_x);
        const long long t2 = t / tiles_x;
        const int ty = (int)(t2 % tiles_y);
        const long long img = t2 / tiles_y;
        const int y0 = ty * kC24Tile - 1, x0 = tx * kC24Tile - 1;
        const __nv_bfloat16 *src = x + img * h * wd * kC24;
        const uint32_t dst = s_halo + buf * kC24HaloBytes;
        for (int i = tid; i < kC24Halo * kC24Halo * 3; i += 256) {
            const int px = i / 3, j = i - px * 3;
            const int hy = px / kC24Halo, hx = px - hy * kC24Halo;
            const int gy = y0 + hy, gx = x0 + hx;
            const bool in = gy >= 0 && gy < h && gx >= 0 && gx < wd;
            const __nv_bfloat16 *g = src + ((long long)(in ? gy : 0) * wd + (in ? gx : 0)) * kC24 + j * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + px * kC24Pitch + j * 16), "l"(g), "r"(in ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // per-lane ldmatrix geometry.  A: lane -> (pixel row of the m16 tile, k half); B: lane -> (output channel, k half)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kh = lane >> 4;
    const int b_n = ((lane >> 4) << 3) + (lane & 7), b_kh = (lane >> 3) & 1;
    uint32_t a_off[kC24KSteps];                      // byte offset of this lane's chunk within the halo, relative to the tile pixel
#pragma unroll
    for (int s2 = 0; s2 < kC24KSteps; ++s2) {
        const int q = 2 * s2 + a_kh;
        const int tap = q / 3, part = q - tap * 3;
        const int dy = tap / 3, dx = tap - dy * 3;
        a_off[s2] = q < 27 ? (uint32_t)((dy * kC24Halo + dx) * kC24Pitch + part * 16) : 0xFFFFFFFFu;
    }
    float bv[3][2];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        bv[nt][0] = bias[nt * 8 + 2 * (lane & 3)];
        bv[nt][1] = bias[nt * 8 + 2 * (lane & 3) + 1];
    }

    long long t = blockIdx.x;
    if (t < tiles) stage(t, 0);
    int buf = 0;
    for (; t < tiles; t += gridDim.x, buf ^= 1) {
        const long long tn = t + gridDim.x;
        if (tn < tiles) {
            stage(tn, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int tx = (int)(t % tiles_x);
        const long long t2 = t / tiles_x;
        const int ty = (int)(t2 % tiles_y);
        const long long img = t2 / tiles_y;
        const uint32_t halo = s_halo + buf * kC24HaloBytes;
        // warp -> output rows 2*warp, 2*warp + 1 of the tile (two m16 tiles of 16 pixels each)
        float acc[2][3][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][nt][e] = 0.f;
        const uint32_t a_base0 = halo + ((2 * warp) * kC24Halo + a_row) * kC24Pitch;          // pixel (row 2w, col a_row) + tap offset
        const uint32_t a_base1 = a_base0 + kC24Halo * kC24Pitch;
        const uint32_t b_base = s_w + b_n * kC24WPitch + b_kh * 16;
#pragma unroll
        for (int s2 = 0; s2 < kC24KSteps; ++s2) {
            uint32_t a0[4], a1[4], b01[4], b2[2];
            const bool zero = a_off[s2] == 0xFFFFFFFFu;
            ldsm_x4(zero ? s_zero : a_base0 + a_off[s2], a0[0], a0[1], a0[2], a0[3]);
            ldsm_x4(zero ? s_zero : a_base1 + a_off[s2], a1[0], a1[1], a1[2], a1[3]);
            ldsm_x4(b_base + s2 * 32, b01[0], b01[1], b01[2], b01[3]);                        // output channels 0..15
            ldsm_x2(s_w + (16 + (lane & 7)) * kC24WPitch + ((lane >> 3) & 1) * 16 + s2 * 32, b2[0], b2[1]);   // 16..23
            mma_bf16_16816(acc[0][0], a0[0], a0[1], a0[2], a0[3], b01[0], b01[1]);
            mma_bf16_16816(acc[0][1], a0[0], a0[1], a0[2], a0[3], b01[2], b01[3]);
            mma_bf16_16816(acc[0][2], a0[0], a0[1], a0[2], a0[3], b2[0], b2[1]);
            mma_bf16_16816(acc[1][0], a1[0], a1[1], a1[2], a1[3], b01[0], b01[1]);
            mma_bf16_16816(acc[1][1], a1[0], a1[1], a1[2], a1[3], b01[2], b01[3]);
            mma_bf16_16816(acc[1][2], a1[0], a1[1], a1[2], a1[3], b2[0], b2[1]);
        }
        // epilogue: bias + SiLU (+ the input pixel, read back from the halo) -> bf16
        const int g = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int oy = ty * kC24Tile + 2 * warp + m;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int px = g + 8 * hh, ox = tx * kC24Tile + px;
                if (oy < h && ox < wd) {
                    __nv_bfloat16 *dst = y + ((img * h + oy) * wd + ox) * kC24 + 2 * tq;
                    const uint32_t res_addr = halo + ((2 * warp + m + 1) * kC24Halo + px + 1) * kC24Pitch + 4 * tq;
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt) {
                        float v0 = silu(acc[m][nt][2 * hh] + bv[nt][0]), v1 = silu(acc[m][nt][2 * hh + 1] + bv[nt][1]);
                        if (residual) {
                            uint32_t rv;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rv) : "r"(res_addr + nt * 16) : "memory");
                            v0 += __uint_as_float(rv << 16);
                            v1 += __uint_as_float(rv & 0xffff0000u);
                        }
                        const __nv_bfloat162 o = __floats2bfloat162_rn(v0, v1);
                        *reinterpret_cast<uint32_t *>(dst + nt * 8) = *reinterpret_cast<const uint32_t *>(&o);
                    }
                }
            }
        }
        __syncthreads();      // everyone is done with this halo buffer before the next stage() overwrites it
    }
}

// ---- squeeze-excitation gate: gate[n, c] = sigmoid(W2 * silu(W1 * pooled[n] + b1) + b2).
//      One CTA serves kSeF frames, so each weight matrix (up to 0.4 MB at c = 1536) is pulled from L2 once per kSeF
//      frames; both layers read the weights as float4 with several independent loads in flight per thread (the
//      first version was a chain of dependent L2 round trips: ~50 us per launch for ~1 MFLOP).  Round 2 tried eight frames per
//      CTA (half the L2 traffic: 33 us instead of 21 at c = 1536) and clusters of four CTAs splitting both weight matrices with
//      the hidden units exchanged through distributed shared memory (25 us): neither the L2 bytes nor the weight reads bound it.
constexpr int kSeF = 4;
constexpr int kSeThreads = 512;
constexpr int kSeMaxC = 12 * 128;          // FC1 keeps one weight row (c / 128 float4 per lane) in registers
__global__ void __launch_bounds__(kSeThreads) se_gate_kernel(const float *__restrict__ pooled, const float *__restrict__ w1,
                                                             const float *__restrict__ b1, const float *__restrict__ w2t,
                                                             const float *__restrict__ b2, float *__restrict__ gate, int n, int c, int sq,
                                                             int out_bf16, int parts) {
    extern __shared__ __align__(16) float se_sm[];
    float *s_pool = se_sm;                 // [kSeF][c]
    float *s_hid = se_sm + kSeF * c;       // [kSeF][sq]
    const int f0 = blockIdx.x * kSeF;
    const int nf = min(kSeF, n - f0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c4 = c >> 2;
    for (int i = tid; i < kSeF * c4; i += kSeThreads) {
        const int f = i / c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < nf) {                      // partial means of the depthwise kernel's (band, segment) units, fixed order
            const float4 *pp = reinterpret_cast<const float4 *>(pooled + (long long)(f0 + f) * parts * c) + (i - f * c4);
            for (int q = 0; q < parts; ++q) {
                const float4 u = __ldg(pp + (long long)q * c4);
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
            }
        }
        reinterpret_cast<float4 *>(s_pool)[i] = v;
    }
    __syncthreads();
    for (int j = warp; j < sq; j += kSeThreads / 32) {
        // the whole weight row is requested before the first use: one L2 round trip per row instead of one per few loads
        const float4 *wrow = reinterpret_cast<const float4 *>(w1 + (long long)j * c);
        float4 wv[kSeMaxC / 128];
#pragma unroll
        for (int i = 0; i < kSeMaxC / 128; ++i) {
            const int k = lane + 32 * i;
            wv[i] = k < c4 ? __ldg(wrow + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float acc[kSeF];
#pragma unroll
        for (int f = 0; f < kSeF; ++f) acc[f] = 0.f;
#pragma unroll
        for (int i = 0; i < kSeMaxC / 128; ++i) {
            const int k = lane + 32 * i;
            if (k < c4) {
#pragma unroll
                for (int f = 0; f < kSeF; ++f) {
                    const float4 pv = reinterpret_cast<const float4 *>(s_pool + f * c)[k];
                    acc[f] = fmaf(wv[i].x, pv.x, fmaf(wv[i].y, pv.y, fmaf(wv[i].z, pv.z, fmaf(wv[i].w, pv.w, acc[f]))));
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kSeF; ++f) {
            float v = acc[f];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_hid[f * sq + j] = silu(v + b1[j]);
        }
    }
    __syncthreads();
    for (int i = tid; i < c4; i += kSeThreads) {          // 4 consecutive channels per thread
        float4 acc[kSeF];
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(b2) + i);
#pragma unroll
        for (int f = 0; f < kSeF; ++f) acc[f] = bb;
        for (int j0 = 0; j0 < sq; j0 += 16) {
            float4 wv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u)
                wv[u] = j0 + u < sq ? __ldg(reinterpret_cast<const float4 *>(w2t + (long long)(j0 + u) * c) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (j0 + u < sq) {
#pragma unroll
                    for (int f = 0; f < kSeF; ++f) {
                        const float hv = s_hid[f * sq + j0 + u];
                        acc[f].x = fmaf(wv[u].x, hv, acc[f].x);
                        acc[f].y = fmaf(wv[u].y, hv, acc[f].y);
                        acc[f].z = fmaf(wv[u].z, hv, acc[f].z);
                        acc[f].w = fmaf(wv[u].w, hv, acc[f].w);
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kSeF; ++f)
            if (f < nf) {
                float4 o;
                o.x = 1.f / (1.f + __expf(-acc[f].x));
                o.y = 1.f / (1.f + __expf(-acc[f].y));
                o.z = 1.f / (1.f + __expf(-acc[f].z));
                o.w = 1.f / (1.f + __expf(-acc[f].w));
                if (out_bf16) {
                    const __nv_bfloat162 q0 = __floats2bfloat162_rn(o.x, o.y), q1 = __floats2bfloat162_rn(o.z, o.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t *>(&q0);
                    pk.y = *reinterpret_cast<const uint32_t *>(&q1);
                    reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(gate) + (long long)(f0 + f) * c)[i] = pk;
                } else {
                    reinterpret_cast<float4 *>(gate + (long long)(f0 + f) * c)[i] = o;
                }
            }
    }
}

}  // namespace

static int stem_impl(const void *x, bool u8, const float *mean, const float *stdv, int n, int h, int wd, const float *w, const float *bias,
                     int cout, void *y, int out_padded, void *stream, int same_tf = 0) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0, EWVIT_ERR_INVALID_ARG, "ewvit_stem_conv_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && bias && y && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG, "ewvit_stem_conv_fwd: NULL or misaligned pointer");
    EWVIT_REQUIRE(cout % 8 == 0 && cout <= kStemMaxC, EWVIT_ERR_UNSUPPORTED, "ewvit_stem_conv_fwd: cout must be a multiple of 8, <= 32");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int ho = (h - 1) / 2 + 1, wo = (wd - 1) / 2 + 1;          // = ceil(h / 2): the same for both padding rules
    // TensorFlow 'SAME' (efficientnet_pytorch): total padding max((ho-1)*2 + 3 - h, 0), floor(total/2) of it on top/left
    const int in_pad = same_tf ? (((ho - 1) * 2 + 3 - h) > 0 ? ((ho - 1) * 2 + 3 - h) / 2 : 0) : 1;
    EWVIT_REQUIRE(!same_tf || h == wd, EWVIT_ERR_UNSUPPORTED, "ewvit_stem_conv_same_fwd: square frames only");
    const long long total = (long long)n * ho * ((wo + kStemPx - 1) / kStemPx);
    long long blocks = (total + 127) / 128;
    const long long cap = (long long)ewvit_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    if (u8)
        stem_conv_kernel<unsigned char><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char *>(x), w, bias,
                                                                                           static_cast<__nv_bfloat16 *>(y), n, h, wd, ho, wo, cout,
                                                                                           out_padded ? 1 : 0, mean, stdv, in_pad);
    else
        stem_conv_kernel<float><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(static_cast<const float *>(x), w, bias,
                                                                                   static_cast<__nv_bfloat16 *>(y), n, h, wd, ho, wo, cout,
                                                                                   out_padded ? 1 : 0, nullptr, nullptr, in_pad);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_stem_conv_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                   void *y, void *stream) {
    return stem_impl(x, false, nullptr, nullptr, n, h, wd, w, bias, cout, y, 0, stream);
}

extern "C" int ewvit_stem_conv_same_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                        void *y, void *stream) {
    return stem_impl(x, false, nullptr, nullptr, n, h, wd, w, bias, cout, y, 0, stream, 1);
}

extern "C" int ewvit_stem_conv_padded_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                          void *y, void *stream) {
    return stem_impl(x, false, nullptr, nullptr, n, h, wd, w, bias, cout, y, 1, stream);
}

extern "C" int ewvit_stem_conv_u8_fwd(const uint8_t *x, const float *mean, const float *stdv, int n, int h, int wd, const float *w,
                                      const float *bias, int cout, void *y, int out_padded, void *stream) {
    EWVIT_REQUIRE(mean && stdv, EWVIT_ERR_INVALID_ARG, "ewvit_stem_conv_u8_fwd: mean/std missing");
    return stem_impl(x, true, mean, stdv, n, h, wd, w, bias, cout, y, out_padded, stream);
}

template <int SC, int STRIDE>
static int launch_dw(const CUtensorMap &tm, const float *w, const float *bias, void *y, float *pooled, int n, int h, int wd, int ho,
                     int wo, int c, int fb, int stage_bytes, size_t smem, unsigned grid, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwconv3x3_kernel<SC, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    dwconv3x3_kernel<SC, STRIDE><<<grid, 256, smem, stream>>>(tm, w, bias, static_cast<__nv_bfloat16 *>(y), pooled, n, h, wd, ho, wo, c,
                                                             fb, stage_bytes);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_dwconv3x3_nhwc_bf16(const void *x, const float *w, const float *bias, int n, int h, int wd, int c,
                                         int stride, void *y, float *pooled, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0 && c > 0, EWVIT_ERR_INVALID_ARG, "ewvit_dwconv3x3_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && bias && y && ewvit_aligned16(x) && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG,
                  "ewvit_dwconv3x3_nhwc_bf16: NULL or misaligned pointer");
    EWVIT_REQUIRE(c % 64 == 0 && (stride == 1 || stride == 2), EWVIT_ERR_UNSUPPORTED,
                  "ewvit_dwconv3x3_nhwc_bf16: needs c %% 64 == 0 and stride 1|2 (got c=%d stride=%d)", c, stride);
    EWVIT_REQUIRE(h + 2 <= 256 && wd + 2 <= 256, EWVIT_ERR_UNSUPPORTED, "ewvit_dwconv3x3_nhwc_bf16: plane %dx%d exceeds the TMA box limit", h, wd);
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    // 64-channel slabs when two stages of the padded input plane fit ~100 KB of shared memory (two CTAs per SM), else 32
    const size_t plane64 = (size_t)(h + 2) * (wd + 2) * 64 * 2;
    const bool wide = 2 * plane64 <= 100 * 1024;
    const int sc = wide ? 64 : 32;
    const size_t plane = (size_t)(h + 2) * (wd + 2) * sc * 2;
    EWVIT_REQUIRE(2 * plane <= 190 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_dwconv3x3_nhwc_bf16: %dx%d input plane too large for the staged kernel", h, wd);
    // frames per pass: enough (frame, column, channel group) items for the 256 threads, within ~48 KB per stage
    int fb = 1;
    while (fb < 8 && fb * wo * (sc / 4) < 448 && 2 * (2 * fb) * plane <= 96 * 1024 && fb * 2 <= n) fb *= 2;
    const int stage_bytes = (int)(((size_t)fb * plane + 127) / 128 * 128);
    const size_t smem = (size_t)2 * stage_bytes + (size_t)10 * sc * 4 + (size_t)2 * fb * wo * sc * 4 + 16;
    EWVIT_REQUIRE(smem <= 200 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_dwconv3x3_nhwc_bf16: %dx%d needs %zu bytes of shared memory", h, wd, smem);
    const long long units = (long long)(c / sc) * ((n + fb - 1) / fb);
    long long grid = 2LL * ewvit_num_sms();
    if (grid > units) grid = units;
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[4] = {2, (uint64_t)c * 2, (uint64_t)wd * c * 2, (uint64_t)h * wd * c * 2};
    const uint32_t box[4] = {(uint32_t)sc, (uint32_t)(wd + 2), (uint32_t)(h + 2), (uint32_t)fb};
    rc = ewvit_make_tmap_bf16(&tm, x, 4, dims, strides, box, nullptr, /*swizzle128=*/false);
    if (rc != EWVIT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (wide) {
        if (stride == 1) return launch_dw<64, 1>(tm, w, bias, y, pooled, n, h, wd, ho, wo, c, fb, stage_bytes, smem, (unsigned)grid, st);
        return launch_dw<64, 2>(tm, w, bias, y, pooled, n, h, wd, ho, wo, c, fb, stage_bytes, smem, (unsigned)grid, st);
    }
    if (stride == 1) return launch_dw<32, 1>(tm, w, bias, y, pooled, n, h, wd, ho, wo, c, fb, stage_bytes, smem, (unsigned)grid, st);
    return launch_dw<32, 2>(tm, w, bias, y, pooled, n, h, wd, ho, wo, c, fb, stage_bytes, smem, (unsigned)grid, st);
}

extern "C" int ewvit_conv3x3_c24_fwd(const void *x, const void *w, int wk, const float *bias, int n, int h, int wd, int residual, void *y,
                                     void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0, EWVIT_ERR_INVALID_ARG, "ewvit_conv3x3_c24_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && bias && y && ewvit_aligned16(x) && ewvit_aligned16(w) && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG,
                  "ewvit_conv3x3_c24_fwd: NULL or misaligned pointer");
    EWVIT_REQUIRE(wk >= 216 && wk % 8 == 0, EWVIT_ERR_INVALID_ARG, "ewvit_conv3x3_c24_fwd: weight rows must hold >= 216 elements (got %d)", wk);
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(conv3x3_c24_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC24Smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const long long tiles = (long long)n * ((h + kC24Tile - 1) / kC24Tile) * ((wd + kC24Tile - 1) / kC24Tile);
    long long grid = 3LL * ewvit_num_sms();
    if (grid > tiles) grid = tiles;
    conv3x3_c24_kernel<<<(unsigned)grid, 256, kC24Smem, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x),
                                                                               static_cast<const __nv_bfloat16 *>(w), wk, bias,
                                                                               static_cast<__nv_bfloat16 *>(y), n, h, wd, residual ? 1 : 0);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

// Squeeze-excitation gate only (the scaling is fused into ewvit_conv1x1_gated_nhwc_bf16).
extern "C" int ewvit_se_gate_fwd(const float *pooled, int pool_parts, const float *w1, const float *b1, const float *w2t, const float *b2,
                                 int n, int c, int sq, void *gate, int gate_bf16, void *stream) {
    EWVIT_REQUIRE(n >= 0 && c > 0 && sq > 0 && pool_parts >= 1, EWVIT_ERR_INVALID_ARG, "ewvit_se_gate_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(pooled && w1 && b1 && w2t && b2 && gate && ewvit_aligned16(gate) && ewvit_aligned16(pooled) &&
                      ewvit_aligned16(w1) && ewvit_aligned16(w2t) && ewvit_aligned16(b2), EWVIT_ERR_INVALID_ARG,
                  "ewvit_se_gate_fwd: NULL or misaligned pointer");
    const size_t gate_smem = (size_t)kSeF * (c + sq) * sizeof(float);
    EWVIT_REQUIRE(c % 8 == 0 && c <= kSeMaxC && gate_smem <= 160 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_se_gate_fwd: c=%d sq=%d not supported", c, sq);
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    static bool se_attr[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !se_attr[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(se_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        if (dev >= 0 && dev < 64) se_attr[dev] = true;
    }
    se_gate_kernel<<<(unsigned)((n + kSeF - 1) / kSeF), kSeThreads, gate_smem, (cudaStream_t)stream>>>(pooled, w1, b1, w2t, b2, static_cast<float *>(gate), n, c, sq, gate_bf16 ? 1 : 0, pool_parts);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
