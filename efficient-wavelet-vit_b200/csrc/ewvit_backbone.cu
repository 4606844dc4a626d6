// EfficientNet feature-extractor glue kernels: everything of a FusedMBConv /
// MBConv block that is not a dense contraction.  The dense parts (3x3 FusedMBConv convs, 1x1 expand/project convs,
// the 1x1 head) run on the tcgen05 implicit-GEMM kernel (ewvit_conv_nhwc_bf16) with bias + SiLU + residual fused
// into the epilogue, so an activation tensor is written once and read once -- the eager cuDNN path spends >50 % of
// its device time in separate elementwise passes (profiles/r01_bench_launches.md).  All tensors NHWC bf16,
// BatchNorm (eval) folded into the weights/bias on the host.
#include "ewvit_common.cuh"

namespace {

__device__ __forceinline__ float silu(float x) { return ewvit::silu_fast(x); }

// ---- stem: Conv2d(3 -> cout<=32, 3x3, stride 2, pad 1) + bias + SiLU, fp32 NCHW frames -> bf16 NHWC
//      (also the fp32 -> bf16 conversion of the input; torchvision features[0], sfe.py:150)
constexpr int kStemMaxC = 32;
__global__ void __launch_bounds__(256) stem_conv_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y,
                                                        int n, int h, int wd, int ho, int wo, int cout) {
    __shared__ float s_w[kStemMaxC * 27];
    __shared__ float s_b[kStemMaxC];
    for (int i = threadIdx.x; i < cout * 27; i += blockDim.x) s_w[i] = w[i];
    if (threadIdx.x < cout) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const long long total = (long long)n * ho * wo;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % wo);
        const long long t = idx / wo;
        const int oy = (int)(t % ho);
        const long long img = t / ho;
        float in[27];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int iy = 2 * oy + dy - 1, ix = 2 * ox + dx - 1;
                    in[c * 9 + dy * 3 + dx] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? __ldg(x + ((img * 3 + c) * h + iy) * wd + ix) : 0.f;
                }
        __nv_bfloat16 *py = y + idx * cout;
        for (int c0 = 0; c0 < cout; c0 += 8) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = s_b[c0 + j];
#pragma unroll
                for (int k = 0; k < 27; ++k) a = fmaf(s_w[(c0 + j) * 27 + k], in[k], a);
                o[j] = silu(a);
            }
            uint4 pk;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(o[0], o[1]), b1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(o[4], o[5]), b3 = __floats2bfloat162_rn(o[6], o[7]);
            pk.x = *reinterpret_cast<uint32_t *>(&b0);
            pk.y = *reinterpret_cast<uint32_t *>(&b1);
            pk.z = *reinterpret_cast<uint32_t *>(&b2);
            pk.w = *reinterpret_cast<uint32_t *>(&b3);
            *reinterpret_cast<uint4 *>(py + c0) = pk;
        }
    }
}

// ---- depthwise 3x3 (stride 1|2, pad 1) + bias + SiLU, and the squeeze (spatial mean) of the result.
//      One CTA owns a channel slab (SC = 64 or 32 channels) of `fpc` consecutive frames.  Each frame's input plane
//      for the slab is staged in shared memory with coalesced 16-byte cp.async copies, double-buffered so the next
//      frame streams in while the current one is computed; thread = (pixel lane, 8-channel group) keeps its 72
//      weights in registers for all frames and reads its 9 taps from shared memory (LDS.128, a quarter-warp reads
//      128 contiguous bytes: conflict-free); the squeeze needs no atomics.
template <int SC>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y,
                                                        float *__restrict__ pooled, int n, int h, int wd, int ho, int wo,
                                                        int c, int stride, int fpc, int fb) {
    // fb = frames staged and processed together (small planes: 4 frames of 7x7 keep all pixel lanes busy and
    // amortise the barriers); fpc = frames per CTA (multiple of fb)
    constexpr int G = SC / 8;            // 8-channel groups per slab
    constexpr int PL = 256 / G;          // pixel lanes
    extern __shared__ __align__(16) unsigned char dw_smem[];
    const int plane = h * wd * SC;                                                      // elements of one frame's slab
    __nv_bfloat16 *s_in = reinterpret_cast<__nv_bfloat16 *>(dw_smem);                  // [2][fb][h*wd][SC]
    float *s_w = reinterpret_cast<float *>(dw_smem + (size_t)2 * fb * plane * 2);       // [9][SC] then bias [SC]
    float *s_sum = s_w + 10 * SC;                                                       // [fb][PL][SC + 1]
    const int slabs = c / SC;
    const int cs = (blockIdx.x % slabs) * SC;
    const int f0 = (blockIdx.x / slabs) * fpc;
    const int f1 = min(n, f0 + fpc);
    const int tid = threadIdx.x;
    const int npix = ho * wo;

    auto stage_in = [&](int frame, int buf) {
        const int nf = min(fb, f1 - frame);
        for (int fl = 0; fl < nf; ++fl) {
            const __nv_bfloat16 *px = x + (long long)(frame + fl) * h * wd * c + cs;
            __nv_bfloat16 *dstb = s_in + ((size_t)buf * fb + fl) * plane;
            for (int i = tid; i < h * wd * G; i += 256) {
                const int p = i / G, g = i - p * G;
                const uint32_t dst = ewvit::smem_u32(dstb + p * SC + g * 8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(px + (long long)p * c + g * 8) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_in(f0, 0);
    for (int i = tid; i < 9 * SC; i += 256) s_w[i] = w[(i / SC) * c + cs + (i % SC)];
    if (tid < SC) s_w[9 * SC + tid] = bias[cs + tid];
    __syncthreads();
    const int g = tid % G, pl = tid / G;
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[t][j] = s_w[t * SC + g * 8 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] = s_w[9 * SC + g * 8 + j];

    int buf = 0;
    for (int f = f0; f < f1; f += fb, buf ^= 1) {
        const int nf = min(fb, f1 - f);
        if (f + fb < f1) {
            stage_in(f + fb, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        float sum[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] = 0.f;
        int cur_fl = 0;
        for (int item = pl; item < nf * npix; item += PL) {
            const int fl = item / npix, o = item - fl * npix;
            if (fl != cur_fl) {      // this thread's items moved on to the next frame: park the finished partial sum
#pragma unroll
                for (int j = 0; j < 8; ++j) { s_sum[(cur_fl * PL + pl) * (SC + 1) + g * 8 + j] = sum[j]; sum[j] = 0.f; }
                cur_fl = fl;
            }
            const int oy = o / wo, ox = o - oy * wo;
            const __nv_bfloat16 *sp = s_in + ((size_t)buf * fb + fl) * plane + g * 8;
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = br[j];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int iy = oy * stride + dy - 1;
                if (iy < 0 || iy >= h) continue;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int ix = ox * stride + dx - 1;
                    if (ix < 0 || ix >= wd) continue;
                    const uint4 v = *reinterpret_cast<const uint4 *>(sp + (iy * wd + ix) * SC);
                    const __nv_bfloat162 *vp = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 fv = __bfloat1622float2(vp[j]);
                        a[2 * j] = fmaf(wr[dy * 3 + dx][2 * j], fv.x, a[2 * j]);
                        a[2 * j + 1] = fmaf(wr[dy * 3 + dx][2 * j + 1], fv.y, a[2 * j + 1]);
                    }
                }
            }
            uint4 pk;
            __nv_bfloat162 *pp = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                pp[j] = __floats2bfloat162_rn(silu(a[2 * j]), silu(a[2 * j + 1]));
                const float2 r = __bfloat1622float2(pp[j]);   // pool what is actually stored
                sum[2 * j] += r.x;
                sum[2 * j + 1] += r.y;
            }
            *reinterpret_cast<uint4 *>(y + ((long long)(f + fl) * npix + o) * c + cs + g * 8) = pk;
        }
        if (pooled) {
            // every (frame, pixel lane) slot is written exactly once: by the flush above, here, or as an explicit zero
            for (int fl = 0; fl < nf; ++fl) {
                const bool mine = fl == cur_fl, later = fl > cur_fl;
                if (mine || later) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) s_sum[(fl * PL + pl) * (SC + 1) + g * 8 + j] = mine ? sum[j] : 0.f;
                }
            }
            __syncthreads();
            for (int i = tid; i < nf * SC; i += 256) {
                const int fl = i / SC, ch = i - fl * SC;
                float t = 0.f;
                for (int k = 0; k < PL; ++k) t += s_sum[(fl * PL + k) * (SC + 1) + ch];
                pooled[(long long)(f + fl) * c + cs + ch] = t / (float)npix;
            }
        }
        __syncthreads();   // everyone is done with this stage buffer (and s_sum) before it is refilled
    }
}

// ---- squeeze-excitation gate: gate[n, c] = sigmoid(W2 * silu(W1 * pooled[n] + b1) + b2).
//      One CTA serves kSeF frames so every weight row is read from L2 once per kSeF frames (the two FC matrices are
//      ~0.8 MB at c = 1536: one-frame CTAs made the 512-frame launch read 400 MB of weights).
constexpr int kSeF = 1;
__global__ void __launch_bounds__(256) se_gate_kernel(const float *__restrict__ pooled, const float *__restrict__ w1,
                                                      const float *__restrict__ b1, const float *__restrict__ w2t,
                                                      const float *__restrict__ b2, float *__restrict__ gate, int n, int c, int sq) {
    extern __shared__ float se_sm[];
    float *s_pool = se_sm;                 // [kSeF][c]
    float *s_hid = se_sm + kSeF * c;       // [kSeF][sq]
    const int f0 = blockIdx.x * kSeF;
    const int nf = min(kSeF, n - f0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kSeF * c; i += 256) {
        const int f = i / c;
        s_pool[i] = f < nf ? pooled[(long long)(f0 + f) * c + (i - f * c)] : 0.f;
    }
    __syncthreads();
    for (int j = warp; j < sq; j += 8) {
        float acc[kSeF];
#pragma unroll
        for (int f = 0; f < kSeF; ++f) acc[f] = 0.f;
#pragma unroll 4
        for (int k = lane; k < c; k += 32) {
            const float wv = w1[(long long)j * c + k];
#pragma unroll
            for (int f = 0; f < kSeF; ++f) acc[f] = fmaf(wv, s_pool[f * c + k], acc[f]);
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
    for (int i = tid; i < c; i += 256) {
        float acc[kSeF];
        const float bb = b2[i];
#pragma unroll
        for (int f = 0; f < kSeF; ++f) acc[f] = bb;
#pragma unroll 8
        for (int j = 0; j < sq; ++j) {
            const float wv = w2t[(long long)j * c + i];
#pragma unroll
            for (int f = 0; f < kSeF; ++f) acc[f] = fmaf(wv, s_hid[f * sq + j], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < kSeF; ++f)
            if (f < nf) gate[(long long)(f0 + f) * c + i] = 1.f / (1.f + __expf(-acc[f]));
    }
}

// ---- x[n, hw, c] *= gate[n, c] in place, 8 channels (16 bytes) per thread
__global__ void __launch_bounds__(256) se_scale_kernel(__nv_bfloat16 *__restrict__ x, const float *__restrict__ gate,
                                                       long long total8, int hw, int c8) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total8; i += (long long)gridDim.x * 256) {
        const int cg = (int)(i % c8);
        const long long img = i / ((long long)hw * c8);
        uint4 v = *reinterpret_cast<uint4 *>(x + i * 8);
        const float4 g0 = *reinterpret_cast<const float4 *>(gate + (img * c8 + cg) * 8);
        const float4 g1 = *reinterpret_cast<const float4 *>(gate + (img * c8 + cg) * 8 + 4);
        __nv_bfloat162 *vp = reinterpret_cast<__nv_bfloat162 *>(&v);
        float2 f;
        f = __bfloat1622float2(vp[0]); vp[0] = __floats2bfloat162_rn(f.x * g0.x, f.y * g0.y);
        f = __bfloat1622float2(vp[1]); vp[1] = __floats2bfloat162_rn(f.x * g0.z, f.y * g0.w);
        f = __bfloat1622float2(vp[2]); vp[2] = __floats2bfloat162_rn(f.x * g1.x, f.y * g1.y);
        f = __bfloat1622float2(vp[3]); vp[3] = __floats2bfloat162_rn(f.x * g1.z, f.y * g1.w);
        *reinterpret_cast<uint4 *>(x + i * 8) = v;
    }
}

}  // namespace

extern "C" int ewvit_stem_conv_fwd(const float *x, int n, int h, int wd, const float *w, const float *bias, int cout,
                                   void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0, EWVIT_ERR_INVALID_ARG, "ewvit_stem_conv_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && bias && y && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG, "ewvit_stem_conv_fwd: NULL or misaligned pointer");
    EWVIT_REQUIRE(cout % 8 == 0 && cout <= kStemMaxC, EWVIT_ERR_UNSUPPORTED, "ewvit_stem_conv_fwd: cout must be a multiple of 8, <= 32");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int ho = (h - 1) / 2 + 1, wo = (wd - 1) / 2 + 1;
    const long long total = (long long)n * ho * wo;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ewvit_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    stem_conv_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, w, bias, static_cast<__nv_bfloat16 *>(y), n, h, wd, ho, wo, cout);
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
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    // 64-channel slabs when the staged input plane fits ~56 KB of shared memory, else 32-channel slabs
    const bool wide = (size_t)h * wd * 64 * 2 <= 50 * 1024;   // two stages of the slab's input plane
    const int sc = wide ? 64 : 32;
    int fb = 1;                                                // frames processed together (small planes)
    while (fb < 4 && (size_t)h * wd * sc * 2 * (fb * 2) <= 26 * 1024 && n % (fb * 2) == 0) fb *= 2;
    const size_t smem = (size_t)2 * fb * h * wd * sc * 2 + (size_t)10 * sc * 4 + (size_t)fb * (256 / (sc / 8)) * (sc + 1) * 4;
    EWVIT_REQUIRE(smem <= 200 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_dwconv3x3_nhwc_bf16: %dx%d input plane too large for the staged kernel", h, wd);
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwconv3x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwconv3x3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // frames per CTA: amortise the weight prologue while keeping >= ~4 CTAs per SM in the grid
    int fpc = 8;
    while (fpc > fb && (long long)((n + fpc - 1) / fpc) * (c / sc) < 4LL * ewvit_num_sms()) fpc /= 2;
    if (fpc < fb) fpc = fb;
    const unsigned grid = (unsigned)((long long)((n + fpc - 1) / fpc) * (c / sc));
    if (wide)
        dwconv3x3_kernel<64><<<grid, 256, smem, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x), w, bias,
                                                                        static_cast<__nv_bfloat16 *>(y), pooled, n, h, wd, ho, wo, c, stride, fpc, fb);
    else
        dwconv3x3_kernel<32><<<grid, 256, smem, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x), w, bias,
                                                                        static_cast<__nv_bfloat16 *>(y), pooled, n, h, wd, ho, wo, c, stride, fpc, fb);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_se_apply_nhwc_bf16(void *x, const float *pooled, const float *w1, const float *b1, const float *w2t,
                                        const float *b2, int n, int hw, int c, int sq, float *gate_ws, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hw > 0 && c > 0 && sq > 0, EWVIT_ERR_INVALID_ARG, "ewvit_se_apply_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && pooled && w1 && b1 && w2t && b2 && gate_ws && ewvit_aligned16(x) && ewvit_aligned16(gate_ws), EWVIT_ERR_INVALID_ARG,
                  "ewvit_se_apply_nhwc_bf16: NULL or misaligned pointer");
    const size_t gate_smem = (size_t)kSeF * (c + sq) * sizeof(float);
    EWVIT_REQUIRE(c % 8 == 0 && gate_smem <= 160 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_se_apply_nhwc_bf16: c=%d sq=%d not supported", c, sq);
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    static bool se_attr[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !se_attr[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(se_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        if (dev >= 0 && dev < 64) se_attr[dev] = true;
    }
    se_gate_kernel<<<(unsigned)((n + kSeF - 1) / kSeF), 256, gate_smem, (cudaStream_t)stream>>>(pooled, w1, b1, w2t, b2, gate_ws, n, c, sq);
    EWVIT_LAUNCH_OK();
    const long long total8 = (long long)n * hw * (c / 8);
    long long blocks = (total8 + 255) / 256;
    const long long cap = (long long)ewvit_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    se_scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<__nv_bfloat16 *>(x), gate_ws, total8, hw, c / 8);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
