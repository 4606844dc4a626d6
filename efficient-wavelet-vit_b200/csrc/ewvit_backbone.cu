// EfficientNet feature-extractor glue kernels: everything of a FusedMBConv /
// MBConv block that is not a dense contraction.  The dense parts (3x3 FusedMBConv convs, 1x1 expand/project convs,
// the 1x1 head) run on the tcgen05 implicit-GEMM kernel (ewvit_conv_nhwc_bf16) with bias + SiLU + residual fused
// into the epilogue, so an activation tensor is written once and read once -- the eager cuDNN path spends >50 % of
// its device time in separate elementwise passes (profiles/r01_bench_launches.md).  All tensors NHWC bf16,
// BatchNorm (eval) folded into the weights/bias on the host.
#include "ewvit_common.cuh"

namespace {

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }

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
//      One CTA owns all output pixels of (frame, channel slab of SC = 64 or 32 channels).  The slab's whole input
//      plane is first staged in shared memory with coalesced 16-byte cp.async copies (<= ~50 KB, several CTAs per
//      SM overlap their loads), then thread = (pixel lane, 8-channel group) reads its 9 taps from shared memory
//      (LDS.128, a quarter-warp reads 128 contiguous bytes: conflict-free), and the squeeze needs no atomics.
template <int SC>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y,
                                                        float *__restrict__ pooled, int h, int wd, int ho, int wo, int c,
                                                        int stride) {
    constexpr int G = SC / 8;            // 8-channel groups per slab
    constexpr int PL = 256 / G;          // pixel lanes
    extern __shared__ __align__(16) unsigned char dw_smem[];
    __nv_bfloat16 *s_in = reinterpret_cast<__nv_bfloat16 *>(dw_smem);                  // [h*wd][SC]
    float *s_w = reinterpret_cast<float *>(dw_smem + (size_t)h * wd * SC * 2);          // [9][SC] then bias [SC]
    float *s_sum = s_w + 10 * SC;                                                       // [PL][SC + 1]
    const int slabs = c / SC;
    const long long img = blockIdx.x / slabs;
    const int cs = (blockIdx.x % slabs) * SC;
    const int tid = threadIdx.x;
    const __nv_bfloat16 *px = x + img * h * wd * c + cs;
    for (int i = tid; i < h * wd * G; i += 256) {
        const int p = i / G, g = i - p * G;
        const uint32_t dst = ewvit::smem_u32(s_in + p * SC + g * 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(px + (long long)p * c + g * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < 9 * SC; i += 256) s_w[i] = w[(i / SC) * c + cs + (i % SC)];
    if (tid < SC) s_w[9 * SC + tid] = bias[cs + tid];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int g = tid % G, pl = tid / G;
    float wr[9][8], br[8], sum[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[t][j] = s_w[t * SC + g * 8 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) { br[j] = s_w[9 * SC + g * 8 + j]; sum[j] = 0.f; }
    __nv_bfloat16 *py = y + img * ho * wo * c + cs + g * 8;
    for (int o = pl; o < ho * wo; o += PL) {
        const int oy = o / wo, ox = o - oy * wo;
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
                const uint4 v = *reinterpret_cast<const uint4 *>(s_in + (iy * wd + ix) * SC + g * 8);
                const __nv_bfloat162 *vp = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(vp[j]);
                    a[2 * j] = fmaf(wr[dy * 3 + dx][2 * j], f.x, a[2 * j]);
                    a[2 * j + 1] = fmaf(wr[dy * 3 + dx][2 * j + 1], f.y, a[2 * j + 1]);
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
        *reinterpret_cast<uint4 *>(py + (long long)o * c) = pk;
    }
    if (pooled) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s_sum[pl * (SC + 1) + g * 8 + j] = sum[j];
        __syncthreads();
        if (tid < SC) {
            float s = 0.f;
            for (int i = 0; i < PL; ++i) s += s_sum[i * (SC + 1) + tid];
            pooled[img * c + cs + tid] = s / (float)(ho * wo);
        }
    }
}

// ---- squeeze-excitation: gate = sigmoid(W2 * silu(W1 * pooled + b1) + b2), then x *= gate in place.
//      grid = (frames, splits): every CTA recomputes the (tiny) gate of its frame and scales its share of pixels.
__global__ void __launch_bounds__(256) se_apply_kernel(__nv_bfloat16 *__restrict__ x, const float *__restrict__ pooled,
                                                       const float *__restrict__ w1, const float *__restrict__ b1,
                                                       const float *__restrict__ w2t, const float *__restrict__ b2, int hw,
                                                       int c, int sq) {
    extern __shared__ float se_sm[];
    float *s_pool = se_sm, *s_hid = se_sm + c, *s_gate = s_hid + sq;
    const long long img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < c; i += 256) s_pool[i] = pooled[img * c + i];
    __syncthreads();
    for (int j = warp; j < sq; j += 8) {
        float s = 0.f;
        for (int k = lane; k < c; k += 32) s = fmaf(w1[(long long)j * c + k], s_pool[k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_hid[j] = silu(s + b1[j]);
    }
    __syncthreads();
    for (int i = tid; i < c; i += 256) {
        float s = b2[i];
        for (int j = 0; j < sq; ++j) s = fmaf(w2t[(long long)j * c + i], s_hid[j], s);
        s_gate[i] = 1.f / (1.f + __expf(-s));
    }
    __syncthreads();
    const int c8 = c / 8;
    const long long per = ((long long)hw * c8 + gridDim.y - 1) / gridDim.y;
    const long long lo = blockIdx.y * per, hi = min((long long)hw * c8, lo + per);
    __nv_bfloat16 *px = x + img * hw * c;
    for (long long i = lo + tid; i < hi; i += 256) {
        const int cg = (int)(i % c8);
        uint4 v = *reinterpret_cast<uint4 *>(px + i * 8);
        __nv_bfloat162 *vp = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(vp[j]);
            f.x *= s_gate[cg * 8 + 2 * j];
            f.y *= s_gate[cg * 8 + 2 * j + 1];
            vp[j] = __floats2bfloat162_rn(f.x, f.y);
        }
        *reinterpret_cast<uint4 *>(px + i * 8) = v;
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
    const bool wide = (size_t)h * wd * 64 * 2 <= 56 * 1024;
    const int sc = wide ? 64 : 32;
    const size_t smem = (size_t)h * wd * sc * 2 + (size_t)10 * sc * 4 + (size_t)(256 / (sc / 8)) * (sc + 1) * 4;
    EWVIT_REQUIRE(smem <= 200 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_dwconv3x3_nhwc_bf16: %dx%d input plane too large for the staged kernel", h, wd);
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwconv3x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dwconv3x3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const unsigned grid = (unsigned)((long long)n * (c / sc));
    if (wide)
        dwconv3x3_kernel<64><<<grid, 256, smem, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x), w, bias,
                                                                        static_cast<__nv_bfloat16 *>(y), pooled, h, wd, ho, wo, c, stride);
    else
        dwconv3x3_kernel<32><<<grid, 256, smem, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(x), w, bias,
                                                                        static_cast<__nv_bfloat16 *>(y), pooled, h, wd, ho, wo, c, stride);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_se_apply_nhwc_bf16(void *x, const float *pooled, const float *w1, const float *b1, const float *w2t,
                                        const float *b2, int n, int hw, int c, int sq, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hw > 0 && c > 0 && sq > 0, EWVIT_ERR_INVALID_ARG, "ewvit_se_apply_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && pooled && w1 && b1 && w2t && b2 && ewvit_aligned16(x), EWVIT_ERR_INVALID_ARG,
                  "ewvit_se_apply_nhwc_bf16: NULL or misaligned pointer");
    EWVIT_REQUIRE(c % 8 == 0 && (2 * c + sq) * 4 <= 48 * 1024, EWVIT_ERR_UNSUPPORTED, "ewvit_se_apply_nhwc_bf16: c=%d sq=%d not supported", c, sq);
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int splits = hw >= 100 ? 4 : 2;
    se_apply_kernel<<<dim3(n, splits), 256, (2 * c + sq) * sizeof(float), (cudaStream_t)stream>>>(
        static_cast<__nv_bfloat16 *>(x), pooled, w1, b1, w2t, b2, hw, c, sq);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
