// MWT glue kernels around the tensor-core convs.
//
//  * mwt_head: high-frequency subbands of one level  ->  bilinear upsample to the level-1 grid
//    (F.interpolate, mwt.py:79-81)  ->  the three per-colour Conv2d(3->18,3x3,p1)+BN+ReLU
//    (hf_conv['seperate'], mwt.py:84-86)  ->  concatenated 54 channels, written as the bf16
//    "padded-flat" NHWC tensor [N, H+2, W+2, 64] (channels 54..63 zero) that the 54->128 fusion conv
//    consumes through TMA.  K = 27 per group is far too skinny for tensor cores: CUDA cores, fp32.
//  * maxpool2x2 (freq_pool[0], mwt.py:39) and the global average pool (freq_pool[4], mwt.py:43) on
//    NHWC bf16.
#include "ewvit_common.cuh"

namespace {

constexpr int kTile = 16;               // 16x16 output pixels per CTA
constexpr int kHalo = kTile + 2;        // 18
constexpr int kHaloPitch = 20;
constexpr int kHeadThreads = 192;       // 64 pixel quads x 3 colour groups
constexpr int kOcPerGroup = 18;
constexpr int kWPitch = 28;             // 27 weights per output channel, padded to 7 float4
constexpr int kOutC = 64;               // 54 real + 10 zero channels
constexpr int kUpBytes = 9 * kHalo * kHaloPitch * 4;                    // 12960
constexpr int kWBytes = (3 * kOcPerGroup * kWPitch + 2 * 56) * 4;       // 6496
constexpr int kOutPitchW = 33;          // 32 words (64 bf16) per pixel + 1 pad word: conflict-free column access
constexpr int kHeadSmem = kUpBytes + kWBytes + kTile * kTile * kOutPitchW * 4;   // 53248

struct HeadParams {
    const float *hf;      // [n, 9, hin, win]  (colour-major, subband-minor: the reference's reshape at mwt.py:77)
    const float *w;       // [3 groups][18][3][3][3]
    const float *scale;   // [54] folded BN scale
    const float *shift;   // [54] folded conv bias + BN shift
    __nv_bfloat16 *y;     // [n, hout+2, wout+2, 64]
    int n, hin, win, hout, wout;
    float ry, rx;         // hin/hout, win/wout (PyTorch area_pixel_compute_scale, align_corners=False)
};

// PyTorch bilinear source index (align_corners=False): max(0, scale*(dst+0.5)-0.5)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int &i0, int &i1, float &l1) {
    float s = scale * (dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

__global__ void __launch_bounds__(kHeadThreads) mwt_head_kernel(const HeadParams p) {
    extern __shared__ __align__(16) unsigned char head_smem[];
    float (*s_up)[kHalo][kHaloPitch] = reinterpret_cast<float (*)[kHalo][kHaloPitch]>(head_smem);
    float *s_w = reinterpret_cast<float *>(head_smem + kUpBytes);
    float *s_scale = s_w + 3 * kOcPerGroup * kWPitch;
    float *s_shift = s_scale + 56;
    uint32_t *s_out = reinterpret_cast<uint32_t *>(head_smem + kUpBytes + kWBytes);   // [256 px][33 words]

    const int tid = threadIdx.x;
    const int tiles_x = (p.wout + kTile - 1) / kTile;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
    const int img = blockIdx.y;
    const int y0 = ty * kTile, x0 = tx * kTile;

    for (int i = tid; i < 3 * kOcPerGroup * kWPitch; i += kHeadThreads) {
        const int oc = i / kWPitch, k = i % kWPitch;
        s_w[i] = k < 27 ? p.w[oc * 27 + k] : 0.f;
    }
    if (tid < 54) {
        s_scale[tid] = p.scale[tid];
        s_shift[tid] = p.shift[tid];
    }
    // upsampled halo tile (zero outside the image: the conv's padding)
    const float *src = p.hf + (long long)img * 9 * p.hin * p.win;
    const bool same = (p.hin == p.hout) && (p.win == p.wout);
    // 2916 halo elements over 192 threads: gather in batches of 8 so that up to 32 independent global loads are in
    // flight per thread (a load->store loop left the kernel waiting on one load latency per element: ncu showed
    // 46 % of all stall samples on this staging store)
    constexpr int kBatch = 8;
    for (int i0 = tid; i0 < 9 * kHalo * kHalo; i0 += kHeadThreads * kBatch) {
        float v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kHeadThreads;
            v[u] = 0.f;
            if (i < 9 * kHalo * kHalo) {
                const int c = i / (kHalo * kHalo), rem = i - c * (kHalo * kHalo);
                const int hy = rem / kHalo, hx = rem - hy * kHalo;
                const int oy = y0 + hy - 1, ox = x0 + hx - 1;
                if (oy >= 0 && oy < p.hout && ox >= 0 && ox < p.wout) {
                    const float *pc = src + (long long)c * p.hin * p.win;
                    if (same) {
                        v[u] = __ldg(pc + oy * p.win + ox);
                    } else {
                        int ya, yb, xa, xb;
                        float ly, lx;
                        src_index(oy, p.ry, p.hin, ya, yb, ly);
                        src_index(ox, p.rx, p.win, xa, xb, lx);
                        const float v00 = __ldg(pc + ya * p.win + xa), v01 = __ldg(pc + ya * p.win + xb);
                        const float v10 = __ldg(pc + yb * p.win + xa), v11 = __ldg(pc + yb * p.win + xb);
                        v[u] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kHeadThreads;
            if (i < 9 * kHalo * kHalo) {
                const int c = i / (kHalo * kHalo), rem = i - c * (kHalo * kHalo);
                const int hy = rem / kHalo, hx = rem - hy * kHalo;
                s_up[c][hy][hx] = v[u];
            }
        }
    }
    __syncthreads();

    // ---- 3 -> 18 conv for one colour group on 4 vertically adjacent pixels.  Consecutive lanes own
    //      consecutive columns, so with the 33-word pixel pitch of s_out the packed bf16x2 stores below hit
    //      distinct banks (a 32-word pitch made every lane of a warp collide on one bank).
    {
        const int g = tid / 64, sub = tid % 64;
        const int col = sub % kTile, r0 = (sub / kTile) * 4;
        float in[3][6][3];
#pragma unroll
        for (int ic = 0; ic < 3; ++ic)
#pragma unroll
            for (int dy = 0; dy < 6; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) in[ic][dy][dx] = s_up[g * 3 + ic][r0 + dy][col + dx];
#pragma unroll 1
        for (int oc = 0; oc < kOcPerGroup; oc += 2) {
            float a[2][4];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const float4 *wp = reinterpret_cast<const float4 *>(&s_w[(g * kOcPerGroup + oc + o) * kWPitch]);
                float wv[28];
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const float4 t = wp[i];
                    wv[4 * i] = t.x; wv[4 * i + 1] = t.y; wv[4 * i + 2] = t.z; wv[4 * i + 3] = t.w;
                }
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float wk = wv[ic * 9 + dy * 3 + dx];
                            a0 = fmaf(wk, in[ic][dy][dx], a0);
                            a1 = fmaf(wk, in[ic][dy + 1][dx], a1);
                            a2 = fmaf(wk, in[ic][dy + 2][dx], a2);
                            a3 = fmaf(wk, in[ic][dy + 3][dx], a3);
                        }
                const int ch = g * kOcPerGroup + oc + o;
                const float sc = s_scale[ch], sh = s_shift[ch];
                a[o][0] = fmaxf(fmaf(a0, sc, sh), 0.f);
                a[o][1] = fmaxf(fmaf(a1, sc, sh), 0.f);
                a[o][2] = fmaxf(fmaf(a2, sc, sh), 0.f);
                a[o][3] = fmaxf(fmaf(a3, sc, sh), 0.f);
            }
            const int word = (g * kOcPerGroup + oc) >> 1;      // channels (oc, oc+1) share one 32-bit word
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162 v = __floats2bfloat162_rn(a[0][k], a[1][k]);
                s_out[((r0 + k) * kTile + col) * kOutPitchW + word] = *reinterpret_cast<const uint32_t *>(&v);
            }
        }
        if (g == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int word = 27; word < 32; ++word) s_out[((r0 + k) * kTile + col) * kOutPitchW + word] = 0u;
        }
    }
    __syncthreads();

    // ---- one warp stores one pixel (32 words = 128 contiguous bytes of the NHWC tensor) per instruction
    const int wp_ = p.wout + 2;
    uint32_t *ybase = reinterpret_cast<uint32_t *>(p.y + (long long)img * (p.hout + 2) * wp_ * kOutC);
    const int lane = tid & 31;
#pragma unroll 4
    for (int px = tid >> 5; px < kTile * kTile; px += kHeadThreads / 32) {
        const int oy = y0 + px / kTile, ox = x0 + px % kTile;
        if (oy < p.hout && ox < p.wout)
            ybase[((long long)(oy + 1) * wp_ + (ox + 1)) * (kOutC / 2) + lane] = s_out[px * kOutPitchW + lane];
    }
}

// ---- the head of one wavelet level on warp-level tensor-core MMAs, ONE kernel: bilinear upsample of the nine subband
//      planes into a shared-memory halo (bf16, 16 channel slots per pixel), then the three per-colour 3->18 convs as one
//      block-diagonal 9(16) -> 54(56) direct convolution: an A fragment of tap (dy, dx) is ldmatrix on the halo pixels
//      shifted by the tap (one k16 step per tap, K = 144), B fragments come from [56][144] weights in shared memory,
//      BN + ReLU in registers, the tile leaves through a staging buffer as full 128-byte pixel rows.  Compared with the
//      tcgen05 route (mwt_upsample_kernel + ewvit_mwt_head_conv_fwd) the 16-channel intermediate never touches HBM and
//      the 128-row-tile fixed costs (N = 64) disappear.
constexpr int kHmTile = 16, kHmHalo = 18;
constexpr int kHmPitch = 48;                                  // bytes per halo pixel: 16 bf16 + padding (conflict-free ldmatrix)
constexpr int kHmHaloBytes = kHmHalo * kHmHalo * kHmPitch;    // 15552
constexpr int kHmWPitch = 304;                                // bytes per weight row: 144 bf16 + padding
constexpr int kHmN = 56;                                      // 54 output channels, padded to 7 n8 tiles
constexpr int kHmOutPitchW = 36;                              // words per staged pixel: lane (g, tq) -> bank 4g + tq, conflict-free
constexpr int kHmMisc = 1024;                                  // shift[64] + bilinear index tables
constexpr int kHmSmem = kHmHaloBytes + kHmN * kHmWPitch + kHmMisc + kHmTile * kHmTile * kHmOutPitchW * 4;   // 70464

__device__ __forceinline__ void hm_ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void hm_ldsm_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void hm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 3) mwt_head_mma_kernel(const HeadParams p) {
    extern __shared__ __align__(128) unsigned char hm_smem[];
    unsigned char *s_halo_p = hm_smem;
    const uint32_t s_halo = ewvit::smem_u32(hm_smem);
    const uint32_t s_w = s_halo + kHmHaloBytes;
    float *s_shift = reinterpret_cast<float *>(hm_smem + kHmHaloBytes + kHmN * kHmWPitch);     // [64]
    int *s_ix = reinterpret_cast<int *>(s_shift + 64);                                        // [2][18] (i0 | i1 << 16), x then y
    float *s_l = reinterpret_cast<float *>(s_ix + 2 * kHmHalo);                               // [2][18] lambda
    uint32_t *s_out = reinterpret_cast<uint32_t *>(hm_smem + kHmHaloBytes + kHmN * kHmWPitch + kHmMisc);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (p.wout + kHmTile - 1) / kHmTile, tiles_y = (p.hout + kHmTile - 1) / kHmTile;
    const long long tiles = (long long)p.n * tiles_y * tiles_x;

    // block-diagonal weights W[n = 18 g + oc][k = 16 tap + 3 g + ic] with the folded BatchNorm scale multiplied in before
    // the bf16 rounding, zero elsewhere; the epilogue only adds the shift
    for (int i = tid; i < kHmN * (kHmWPitch / 4); i += 256) reinterpret_cast<uint32_t *>(hm_smem + kHmHaloBytes)[i] = 0u;
    for (int i = tid; i < 64; i += 256) s_shift[i] = i < 54 ? p.shift[i] : 0.f;
    for (int i = tid; i < kHmTile * kHmTile * kHmOutPitchW; i += 256) s_out[i] = 0u;     // words 28..31 (channels 56..63) stay zero
    for (int i = tid; i < kHmHaloBytes / 4; i += 256) reinterpret_cast<uint32_t *>(hm_smem)[i] = 0u;   // channel slots 9..15 stay zero
    __syncthreads();
    for (int i = tid; i < 3 * kOcPerGroup * 27; i += 256) {
        const int g = i / (kOcPerGroup * 27), r = i - g * (kOcPerGroup * 27);
        const int oc = r / 27, r2 = r - oc * 27;
        const int ic = r2 / 9, tap = r2 - ic * 9;
        const int nrow = g * kOcPerGroup + oc;
        reinterpret_cast<__nv_bfloat16 *>(hm_smem + kHmHaloBytes + nrow * kHmWPitch)[tap * 16 + g * 3 + ic] =
            __float2bfloat16_rn(p.w[i] * p.scale[nrow]);
    }

    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kh = lane >> 4;
    const int b_n = ((lane >> 4) << 3) + (lane & 7), b_kh = (lane >> 3) & 1;
    const bool same = (p.hin == p.hout) && (p.win == p.wout);
    const long long plane = (long long)p.hin * p.win;
    const int wp_ = p.wout + 2;
    const int g = lane >> 2, tq = lane & 3;

    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int tx = (int)(t % tiles_x);
        const long long t2 = t / tiles_x;
        const int ty = (int)(t2 % tiles_y);
        const long long img = t2 / tiles_y;
        const int y0 = ty * kHmTile, x0 = tx * kHmTile;
        // ---- bilinear source indices / weights of the 18 halo columns and rows (PyTorch align_corners=False rule);
        //      -1 marks a position outside the image (the conv's zero padding)
        if (tid < 2 * kHmHalo) {
            const int isy = tid / kHmHalo, k = tid - isy * kHmHalo;
            const int o = (isy ? y0 : x0) + k - 1;
            const int osz = isy ? p.hout : p.wout, isz = isy ? p.hin : p.win;
            int i0 = 0, i1 = 0;
            float l1 = 0.f;
            if (o >= 0 && o < osz) {
                if (same) { i0 = i1 = o; }
                else src_index(o, isy ? p.ry : p.rx, isz, i0, i1, l1);
                s_ix[tid] = i0 | (i1 << 16);
            } else {
                s_ix[tid] = -1;
            }
            s_l[tid] = l1;
        }
        __syncthreads();
        // ---- upsampled halo: one work item per (halo pixel, colour) = three subband planes
        const float *src = p.hf + img * 9 * plane;
        for (int it = tid; it < kHmHalo * kHmHalo * 3; it += 256) {
            const int px = it / 3, gcol = it - px * 3;
            const int hy = px / kHmHalo, hx = px - hy * kHmHalo;
            const int ix = s_ix[hx], iy = s_ix[kHmHalo + hy];
            float v[3] = {0.f, 0.f, 0.f};
            if ((ix | iy) >= 0) {
                const float *pc = src + (gcol * 3) * plane;
                const int xa = ix & 0xffff, xb = ix >> 16, ya = iy & 0xffff, yb = iy >> 16;
                if (same) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = __ldg(pc + c * plane + ya * p.win + xa);
                } else {
                    const float lx = s_l[hx], ly = s_l[kHmHalo + hy];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float *pp = pc + c * plane;
                        const float v00 = __ldg(pp + ya * p.win + xa), v01 = __ldg(pp + ya * p.win + xb);
                        const float v10 = __ldg(pp + yb * p.win + xa), v11 = __ldg(pp + yb * p.win + xb);
                        v[c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
                    }
                }
            }
            __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(s_halo_p + px * kHmPitch) + gcol * 3;
            dst[0] = __float2bfloat16_rn(v[0]);
            dst[1] = __float2bfloat16_rn(v[1]);
            dst[2] = __float2bfloat16_rn(v[2]);
        }
        __syncthreads();

        // ---- warp -> output rows 2*warp, 2*warp+1 (two m16 tiles) x 7 n8 tiles, one k16 step per tap
        float acc[2][7][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 7; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][nt][e] = 0.f;
        const uint32_t a_base = s_halo + ((2 * warp) * kHmHalo + a_row) * kHmPitch + a_kh * 16;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            uint32_t a0[4], a1[4];
            hm_ldsm_x4(a_base + (dy * kHmHalo + dx) * kHmPitch, a0[0], a0[1], a0[2], a0[3]);
            hm_ldsm_x4(a_base + ((dy + 1) * kHmHalo + dx) * kHmPitch, a1[0], a1[1], a1[2], a1[3]);
#pragma unroll
            for (int np = 0; np < 3; ++np) {       // n8 tiles 2 np, 2 np + 1
                uint32_t b[4];
                hm_ldsm_x4(s_w + (np * 16 + b_n) * kHmWPitch + b_kh * 16 + tap * 32, b[0], b[1], b[2], b[3]);
                hm_mma(acc[0][2 * np], a0, b[0], b[1]);
                hm_mma(acc[0][2 * np + 1], a0, b[2], b[3]);
                hm_mma(acc[1][2 * np], a1, b[0], b[1]);
                hm_mma(acc[1][2 * np + 1], a1, b[2], b[3]);
            }
            uint32_t b6[2];
            hm_ldsm_x2(s_w + (48 + (lane & 7)) * kHmWPitch + ((lane >> 3) & 1) * 16 + tap * 32, b6[0], b6[1]);
            hm_mma(acc[0][6], a0, b6[0], b6[1]);
            hm_mma(acc[1][6], a1, b6[0], b6[1]);
        }
        // ---- + shift, ReLU -> bf16 pairs -> staging (36-word pixel pitch: lane (g, tq) -> bank 4g + tq)
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {
            const float2 sh = *reinterpret_cast<const float2 *>(s_shift + nt * 8 + 2 * tq);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int px = (2 * warp + m) * kHmTile + g + 8 * hh;
                    const __nv_bfloat162 o = __floats2bfloat162_rn(fmaxf(acc[m][nt][2 * hh] + sh.x, 0.f), fmaxf(acc[m][nt][2 * hh + 1] + sh.y, 0.f));
                    s_out[px * kHmOutPitchW + nt * 4 + tq] = *reinterpret_cast<const uint32_t *>(&o);
                }
        }
        __syncthreads();
        // ---- a warp stores four pixels (4 x 128 contiguous bytes of the padded-flat NHWC tensor) per instruction
        {
            __nv_bfloat16 *ybase = p.y + ((img * (p.hout + 2) + y0 + 1) * wp_ + x0 + 1) * kOutC;      // output pixel (y0, x0)
            const int sub = lane >> 3, q4 = (lane & 7) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int px = warp * 32 + i * 4 + sub;
                const int ry = px >> 4, rx = px & 15;
                if (y0 + ry < p.hout && x0 + rx < p.wout) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(s_out + px * kHmOutPitchW + q4);
                    *reinterpret_cast<uint4 *>(ybase + (ry * wp_ + rx) * kOutC + q4 * 2) = v;
                }
            }
        }
        // no third barrier: the next tile's index table / halo writes are ordered behind this tile's ldmatrix reads by the
        // barrier above, its staging writes behind this tile's staging reads by its own barriers
    }
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

extern "C" int ewvit_mwt_head_fwd(const float *hf, int n, int hin, int win, int hout, int wout, const float *w,
                                  const float *scale, const float *shift, void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hin > 0 && win > 0 && hout > 0 && wout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(hf && w && scale && shift && y, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_fwd: NULL pointer");
    EWVIT_REQUIRE(ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_fwd: y must be 16-byte aligned");
    EWVIT_REQUIRE(n <= 65535, EWVIT_ERR_UNSUPPORTED, "ewvit_mwt_head_fwd: n=%d > 65535 frames per call", n);
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    HeadParams p;
    p.hf = hf; p.w = w; p.scale = scale; p.shift = shift; p.y = static_cast<__nv_bfloat16 *>(y);
    p.n = n; p.hin = hin; p.win = win; p.hout = hout; p.wout = wout;
    p.ry = (float)hin / (float)hout;
    p.rx = (float)win / (float)wout;
    const int tiles = ((hout + kTile - 1) / kTile) * ((wout + kTile - 1) / kTile);
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(mwt_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    mwt_head_kernel<<<dim3(tiles, n), kHeadThreads, kHeadSmem, (cudaStream_t)stream>>>(p);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_mwt_head_mma_fwd(const float *hf, int n, int hin, int win, int hout, int wout, const float *w,
                                      const float *scale, const float *shift, void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hin > 0 && win > 0 && hout > 0 && wout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_mma_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(hf && w && scale && shift && y, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_mma_fwd: NULL pointer");
    EWVIT_REQUIRE(ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_mma_fwd: y must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    HeadParams p;
    p.hf = hf; p.w = w; p.scale = scale; p.shift = shift; p.y = static_cast<__nv_bfloat16 *>(y);
    p.n = n; p.hin = hin; p.win = win; p.hout = hout; p.wout = wout;
    p.ry = (float)hin / (float)hout;
    p.rx = (float)win / (float)wout;
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(mwt_head_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHmSmem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const long long tiles = (long long)n * ((hout + kHmTile - 1) / kHmTile) * ((wout + kHmTile - 1) / kHmTile);
    long long grid = 3LL * ewvit_num_sms();
    if (grid > tiles) grid = tiles;
    mwt_head_mma_kernel<<<(unsigned)grid, 256, kHmSmem, (cudaStream_t)stream>>>(p);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

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
