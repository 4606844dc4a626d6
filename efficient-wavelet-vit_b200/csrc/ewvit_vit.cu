// Token-level glue of the 2-token Efficient-ViT: everything between the
// tensor-core linears.  All of it is per-token vector math (warp-shuffle reductions), fp32.
#include "ewvit_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x[n,0,:] = cls + pos[idx[n]] ; x[n,1,:] = emb[n] + pos[idx[n]]      (sfe.py:156-159)
__global__ void vit_assemble_kernel(const float *__restrict__ emb, const float *__restrict__ cls,
                                    const float *__restrict__ pos, const int *__restrict__ pos_index, float *__restrict__ x,
                                    long long n, int d, int pos_rows) {
    const long long total = n * d;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / d;
        const int c = (int)(i % d);
        const int pi = pos_index[f];
        // an index outside the table (the reference's broadcast raises, sfe.py:158-159; the host wrappers check first)
        // poisons the frame's tokens instead of reading out of bounds
        const float pe = (unsigned)pi < (unsigned)pos_rows ? pos[(long long)pi * d + c] : __int_as_float(0x7fc00000);
        x[(f * 2) * d + c] = cls[c] + pe;
        x[(f * 2 + 1) * d + c] = emb[f * d + c] + pe;
    }
}

// One warp per row: LayerNorm (biased variance, eps inside the sqrt, like nn.LayerNorm) -> bf16.
// gamma == nullptr: plain fp32 -> bf16 cast of the (strided) rows.
__global__ void layernorm_bf16_kernel(const float *__restrict__ x, long long ldx, const float *__restrict__ gamma,
                                      const float *__restrict__ beta, float eps, __nv_bfloat16 *__restrict__ y,
                                      long long ldy, long long rows, int d) {
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float *px = x + row * ldx;
    __nv_bfloat16 *py = y + row * ldy;
    if (gamma == nullptr) {
        for (int c = lane; c < d; c += 32) py[c] = __float2bfloat16_rn(px[c]);
        return;
    }
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += px[c];
    const float mean = warp_sum(s) / (float)d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float t = px[c] - mean;
        v += t * t;
    }
    const float rstd = rsqrtf(warp_sum(v) / (float)d + eps);
    for (int c = lane; c < d; c += 32) py[c] = __float2bfloat16_rn((px[c] - mean) * rstd * gamma[c] + beta[c]);
}

// Self-attention over T (<= 8) tokens per frame; one warp per (frame, head).      (sfe.py:58-69)
// qkv [n*T, 3*heads*dh] fp32 (q | k | v), out [n*T, heads*dh] bf16.
template <int T>
__global__ void vit_attn_kernel(const float *__restrict__ qkv, __nv_bfloat16 *__restrict__ out, long long n, int heads,
                                int dh, float scale) {
    const int lane = threadIdx.x & 31;
    const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= n * heads) return;
    const long long f = wid / heads;
    const int h = (int)(wid % heads);
    const int inner = heads * dh;
    const float *base = qkv + (f * T) * 3LL * inner + h * dh;
    float dots[T][T];
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) {
            float s = 0.f;
            for (int c = lane; c < dh; c += 32) s += base[i * 3LL * inner + c] * base[j * 3LL * inner + inner + c];
            dots[i][j] = warp_sum(s) * scale;
        }
#pragma unroll
    for (int i = 0; i < T; ++i) {
        float m = dots[i][0];
#pragma unroll
        for (int j = 1; j < T; ++j) m = fmaxf(m, dots[i][j]);
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < T; ++j) {
            dots[i][j] = expf(dots[i][j] - m);
            den += dots[i][j];
        }
        const float inv = 1.f / den;
        for (int c = lane; c < dh; c += 32) {
            float o = 0.f;
#pragma unroll
            for (int j = 0; j < T; ++j) o += dots[i][j] * inv * base[j * 3LL * inner + 2 * inner + c];
            out[(f * T + i) * (long long)inner + h * dh + c] = __float2bfloat16_rn(o);
        }
    }
}

}  // namespace

extern "C" int ewvit_vit_assemble(const float *emb, const float *cls, const float *pos, const int *pos_index, int64_t n,
                                  int d, int pos_rows, float *x, void *stream) {
    EWVIT_REQUIRE(n >= 0 && d > 0 && pos_rows > 0, EWVIT_ERR_INVALID_ARG, "ewvit_vit_assemble: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(emb && cls && pos && pos_index && x, EWVIT_ERR_INVALID_ARG, "ewvit_vit_assemble: NULL pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const long long total = n * d;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ewvit_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    vit_assemble_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(emb, cls, pos, pos_index, x, n, d, pos_rows);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_layernorm_bf16(const float *x, int64_t ldx, const float *gamma, const float *beta, float eps,
                                    void *y, int64_t ldy, int64_t rows, int d, void *stream) {
    EWVIT_REQUIRE(rows >= 0 && d > 0 && ldx >= d && ldy >= d, EWVIT_ERR_INVALID_ARG, "ewvit_layernorm_bf16: bad sizes");
    if (rows == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && y && ((gamma == nullptr) == (beta == nullptr)), EWVIT_ERR_INVALID_ARG,
                  "ewvit_layernorm_bf16: NULL pointer (gamma and beta must both be given or both be NULL)");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int wpb = 8;
    layernorm_bf16_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, ldx, gamma, beta, eps, static_cast<__nv_bfloat16 *>(y), ldy, rows, d);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_vit_attention(const float *qkv, int64_t n, int tokens, int heads, int dim_head, void *out,
                                   void *stream) {
    EWVIT_REQUIRE(n >= 0 && heads > 0 && dim_head > 0, EWVIT_ERR_INVALID_ARG, "ewvit_vit_attention: bad sizes");
    EWVIT_REQUIRE(tokens == 2, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_vit_attention: tokens=%d; the shipped config (patch-size 7 on a 7x7 map) has cls + 1 patch = 2", tokens);
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(qkv && out, EWVIT_ERR_INVALID_ARG, "ewvit_vit_attention: NULL pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int wpb = 8;
    const long long warps = n * heads;
    const float scale = 1.0f / sqrtf((float)dim_head);
    vit_attn_kernel<2><<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        qkv, static_cast<__nv_bfloat16 *>(out), n, heads, dim_head, scale);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
