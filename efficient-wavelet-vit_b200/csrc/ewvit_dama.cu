// DAMA fusion tail: bidirectional cross-attention over the
// 1x1 spatial / frequency tokens, centre-tap fusion gate, softmax(3) adaptive gate, weighted sum, per-video
// mean and the classifier.  One fused kernel instead of ~60 tiny eager launches; per frame this is ~0.4 MFLOP
// of 128-wide mat-vecs, so it runs on CUDA cores in fp32 (no precision loss against the fp32 reference) with
// the (pre-transposed) weights streamed from L2 and shared by the frames of a CTA.
#include "ewvit_common.cuh"

namespace {

constexpr int kFpc = 4;        // frames per CTA
constexpr int kMaxGroups = 4;  // thread groups of D threads: independent projections of a cross-attention run side by side (512 threads: 128 registers each, the 64-row weight prefetch does not spill)

struct DamaParams {
    const float *space_in, *freq_in;   // [n, D]
    const float *wpack;                // packed weights, layout in include/ewvit.h
    float *fused, *space, *freq;       // [n, D] outputs
    long long n;
    int d, heads, depth;
    float ln_eps;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[f][j] = sum_k Wt[k*ldw + j] * x[k][f]   for the kFpc frames of the CTA, k ascending, one fmaf per term.
// Activations live in shared memory CHANNEL-major, x[k][kFpc]: one 16-byte broadcast read brings the four frames' values of
// channel k (frame-major rows cost four 4-byte reads per weight and the shared-memory pipe bounded the kernel).
// The first version asked for 16 weight rows at a time from a 128-thread CTA that ran its 28 projections one after the other: ~230 dependent L2
// round trips per CTA with four warps per SM to hide them -- 0.34 ms per 512 frames for 0.17 GFMA.
static_assert(kFpc == 4, "activations are read as float4 = the four frames of a channel");
__device__ __forceinline__ float ldg_nc(const float *ptr) {       // volatile: keeps the batch of loads where it is written
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(ptr));
    return v;
}
// in_dim % 32 == 0 (the launcher requires d % 32 == 0).  Chunks of 32 weight rows, the NEXT chunk requested before the FMAs
// of the current one: 32..64 independent, coalesced L2 loads in flight per thread.
__device__ __forceinline__ void matvec_t(const float *__restrict__ wt, int ldw, int in_dim, const float *x, int j, float (&acc)[kFpc]) {
#pragma unroll
    for (int f = 0; f < kFpc; ++f) acc[f] = 0.f;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float *wp = wt + j;
    float wc[32], wn[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) wc[u] = ldg_nc(wp + (long long)u * ldw);
    for (int k0 = 0; k0 < in_dim; k0 += 32) {
        if (k0 + 32 < in_dim) {
#pragma unroll
            for (int u = 0; u < 32; ++u) wn[u] = ldg_nc(wp + (long long)(k0 + 32 + u) * ldw);
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const float4 xv = x4[k0 + u];
            acc[0] = fmaf(wc[u], xv.x, acc[0]);
            acc[1] = fmaf(wc[u], xv.y, acc[1]);
            acc[2] = fmaf(wc[u], xv.z, acc[2]);
            acc[3] = fmaf(wc[u], xv.w, acc[3]);
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) wc[u] = wn[u];
    }
}

// blockDim.x = G * D: thread (g, j) = group g, output column j.  Independent projections are dealt to the groups round-robin
// (q | k,v of the query tokens | k,v of the context: five per cross-attention; fusion conv | gate MLP at the end), so the
// dependent chain of a CTA is 3 projections per cross-attention instead of 7.  Every output is still ONE thread's fmaf chain
// over k ascending: results are bit-identical to the single-group version whatever G is.
#define AT(c, f) ((c) * kFpc + (f))      // channel-major activations
// kD = 128 (the shipped dama_dim): row pitches of the weight matrices are compile-time constants, so the 64 weight addresses of a
// thread are immediates off one base register; kD = 0: any d % 32 == 0.
template <int kD>
__global__ void __launch_bounds__(kMaxGroups * 128) dama_tail_kernel(const DamaParams p) {
    extern __shared__ __align__(16) float sm[];
    const int D = kD ? kD : p.d, G = blockDim.x / D;
    const int g = threadIdx.x / D, j = threadIdx.x - g * D;
    float *s_s = sm;                 // [D][kFpc] spatial tokens
    float *s_f = s_s + kFpc * D;     // frequency tokens
    float *s_xn = s_f + kFpc * D;    // normalised query tokens
    float *s_q = s_xn + kFpc * D;
    float *s_k0 = s_q + kFpc * D, *s_k1 = s_k0 + kFpc * D, *s_v0 = s_k1 + kFpc * D, *s_v1 = s_v0 + kFpc * D;
    float *s_p0 = s_v1 + kFpc * D, *s_p1 = s_p0 + kFpc * D;   // per-channel q*k products
    float *s_att = s_p1 + kFpc * D;
    float *s_cat = s_att + kFpc * D;   // [2D][kFpc]
    float *s_hid = s_cat + kFpc * 2 * D;   // [kFpc][D/2]
    float *s_gate = s_hid + kFpc * (D / 2);   // [kFpc][4]
    float *s_fus = s_gate + kFpc * 4;          // [D][kFpc] fusion-gate features

    const long long f0 = (long long)blockIdx.x * kFpc;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dh = D / p.heads;
    const float scale = rsqrtf((float)dh);

    for (int i = threadIdx.x; i < kFpc * D; i += blockDim.x) {
        const int f = i / D, c = i - f * D;
        const long long fr = f0 + f;
        s_s[AT(c, f)] = fr < p.n ? p.space_in[fr * D + c] : 0.f;
        s_f[AT(c, f)] = fr < p.n ? p.freq_in[fr * D + c] : 0.f;
    }
    __syncthreads();

    const long long blk = 4LL * D * D + 3LL * D;
    for (int l = 0; l < p.depth; ++l) {
        for (int dir = 0; dir < 2; ++dir) {
            const float *wb = p.wpack + (l * 2 + dir) * blk;
            const float *ln_w = wb, *ln_b = wb + D, *wq_t = wb + 2 * D, *wkv_t = wq_t + (long long)D * D;
            const float *wo_t = wkv_t + 2LL * D * D, *bo = wo_t + (long long)D * D;
            float *xq = dir == 0 ? s_s : s_f;     // query stream (updated in place)
            float *ctx = dir == 0 ? s_f : s_s;    // context: the other stream (dir 1 sees the UPDATED spatial tokens)

            // LayerNorm of the query tokens (dama.py:71,75), one warp per frame
            for (int f = warp; f < kFpc; f += nwarps) {
                float s = 0.f;
                for (int c = lane; c < D; c += 32) s += xq[AT(c, f)];
                const float mean = warp_sum(s) / (float)D;
                float v = 0.f;
                for (int c = lane; c < D; c += 32) {
                    const float t = xq[AT(c, f)] - mean;
                    v += t * t;
                }
                const float rstd = rsqrtf(warp_sum(v) / (float)D + p.ln_eps);
                for (int c = lane; c < D; c += 32) s_xn[AT(c, f)] = (xq[AT(c, f)] - mean) * rstd * ln_w[c] + ln_b[c];
            }
            __syncthreads();

            // q from the normalised tokens; k/v from cat(normalised tokens, raw context) (dama.py:38-42): five independent tasks
            for (int t = g; t < 5; t += G) {
                float acc[kFpc];
                const float *w = t == 0 ? wq_t : (t <= 2 ? wkv_t : wkv_t + D);
                const float *x = (t == 0 || t == 1 || t == 3) ? s_xn : ctx;
                float *dst = t == 0 ? s_q : t == 1 ? s_k0 : t == 2 ? s_k1 : t == 3 ? s_v0 : s_v1;
                matvec_t(w, t == 0 ? D : 2 * D, D, x, j, acc);
                *reinterpret_cast<float4 *>(dst + AT(j, 0)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < kFpc * D; i += blockDim.x) {
                s_p0[i] = s_q[i] * s_k0[i];
                s_p1[i] = s_q[i] * s_k1[i];
            }
            __syncthreads();

            // 1 query x 2 keys per head: softmax over the two dots (dama.py:44-48)
            for (int i = threadIdx.x; i < kFpc * D; i += blockDim.x) {
                const int c0 = i / kFpc, f = i - c0 * kFpc;
                const int h0 = (c0 / dh) * dh;
                float d0 = 0.f, d1 = 0.f;
                for (int c = 0; c < dh; ++c) {
                    d0 += s_p0[AT(h0 + c, f)];
                    d1 += s_p1[AT(h0 + c, f)];
                }
                d0 *= scale;
                d1 *= scale;
                const float m = fmaxf(d0, d1);
                const float e0 = expf(d0 - m), e1 = expf(d1 - m);
                const float inv = 1.f / (e0 + e1);
                s_att[i] = (e0 * inv) * s_v0[i] + (e1 * inv) * s_v1[i];
            }
            __syncthreads();

            if (g == 0) {
                float acc[kFpc];
                matvec_t(wo_t, D, D, s_att, j, acc);
                const float b = bo[j];
#pragma unroll
                for (int f = 0; f < kFpc; ++f) xq[AT(j, f)] += acc[f] + b;   // residual (dama.py:72,76)
            }
            __syncthreads();
        }
    }

    // ---- fusion gate: 3x3 conv on a 1x1 map == its centre tap (dama.py:124-128,153), folded BN, ReLU
    const float *wf_t = p.wpack + (long long)p.depth * 2 * blk;
    const float *f_scale = wf_t + 2LL * D * D, *f_shift = f_scale + D;
    const float *g1_t = f_shift + D, *g1_b = g1_t + 2LL * D * (D / 2);
    const float *g2 = g1_b + D / 2, *g2_b = g2 + 3 * (D / 2);
    for (int i = threadIdx.x; i < kFpc * D; i += blockDim.x) {      // cat(space, freq) along the channel axis
        s_cat[i] = s_s[i];
        s_cat[kFpc * D + i] = s_f[i];
    }
    __syncthreads();
    // two independent tasks: the fusion conv (D outputs) and the first gate_net layer (D/2 outputs; dama.py:105-113,156-157)
    for (int t = g; t < 2; t += G) {
        float acc[kFpc];
        if (t == 0) {
            matvec_t(wf_t, D, 2 * D, s_cat, j, acc);
#pragma unroll
            for (int f = 0; f < kFpc; ++f) s_fus[AT(j, f)] = fmaxf(fmaf(acc[f], f_scale[j], f_shift[j]), 0.f);
        } else if (j < D / 2) {
            matvec_t(g1_t, D / 2, 2 * D, s_cat, j, acc);
#pragma unroll
            for (int f = 0; f < kFpc; ++f) s_hid[f * (D / 2) + j] = fmaxf(acc[f] + g1_b[j], 0.f);
        }
    }
    __syncthreads();
    for (int f = warp; f < kFpc; f += nwarps) {
        float z[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            float s = 0.f;
            for (int c = lane; c < D / 2; c += 32) s += g2[o * (D / 2) + c] * s_hid[f * (D / 2) + c];
            z[o] = warp_sum(s) + g2_b[o];
        }
        const float m = fmaxf(z[0], fmaxf(z[1], z[2]));
        const float e0 = expf(z[0] - m), e1 = expf(z[1] - m), e2 = expf(z[2] - m);
        const float inv = 1.f / (e0 + e1 + e2);
        if (lane == 0) {
            s_gate[f * 4 + 0] = e0 * inv;
            s_gate[f * 4 + 1] = e1 * inv;
            s_gate[f * 4 + 2] = e2 * inv;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kFpc * D; i += blockDim.x) {
        const int f = i / D, c = i - f * D;
        const long long fr = f0 + f;
        if (fr < p.n) {
            const float sv = s_s[AT(c, f)], fv = s_f[AT(c, f)];
            p.fused[fr * D + c] = s_gate[f * 4] * sv + s_gate[f * 4 + 1] * fv + s_gate[f * 4 + 2] * s_fus[AT(c, f)];   // dama.py:159-163
            p.space[fr * D + c] = sv;
            p.freq[fr * D + c] = fv;
        }
    }
}
#undef AT

// Per-video mean over K consecutive frames (dama.py:188-199) and, when cw1 != nullptr, the classifier
// Linear(D->Hc)+ReLU+Linear(Hc->1) on the first feature set (model.py:62-68,92).  One CTA per video.
__global__ void video_head_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ c,
                                  int k, int d, float *__restrict__ ma, float *__restrict__ mb, float *__restrict__ mc,
                                  const float *__restrict__ cw1, const float *__restrict__ cb1, const float *__restrict__ cw2,
                                  const float *__restrict__ cb2, int hc, float *__restrict__ logits) {
    extern __shared__ float sm[];
    float *s_mean = sm, *s_hid = sm + d;
    const long long v = blockIdx.x;
    const int j = threadIdx.x;
    const float *src[3] = {a, b, c};
    float *dst[3] = {ma, mb, mc};
    for (int t = 0; t < 3; ++t) {
        if (!src[t]) continue;
        for (int ch = j; ch < d; ch += blockDim.x) {
            float s = 0.f;
            const float *pv = src[t] + v * k * d + ch;
            for (int i0 = 0; i0 < k; i0 += 16) {        // 16 frames requested at once, summed in frame order as before
                float tmp[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) tmp[u] = i0 + u < k ? __ldg(pv + (long long)(i0 + u) * d) : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (i0 + u < k) s += tmp[u];
            }
            s /= (float)k;
            dst[t][v * d + ch] = s;
            if (t == 0) s_mean[ch] = s;
        }
    }
    if (!cw1) return;
    __syncthreads();
    for (int o = j; o < hc; o += blockDim.x) {
        float s = cb1[o];
        for (int ch = 0; ch < d; ++ch) s = fmaf(cw1[(long long)o * d + ch], s_mean[ch], s);
        s_hid[o] = fmaxf(s, 0.f);
    }
    __syncthreads();
    if (j < 32) {
        float s = 0.f;
        for (int o = j; o < hc; o += 32) s += cw2[o] * s_hid[o];
        s = warp_sum(s);
        if (j == 0) logits[v] = s + cb2[0];
    }
}

}  // namespace

extern "C" int64_t ewvit_dama_wpack_floats(int d, int depth) {
    const int64_t D = d;
    return (int64_t)depth * 2 * (4 * D * D + 3 * D) + 2 * D * D + 2 * D + 2 * D * (D / 2) + D / 2 + 3 * (D / 2) + 3;
}

extern "C" int ewvit_dama_tail_fwd(const float *space_in, const float *freq_in, int64_t n, int d, int heads, int depth,
                                   const float *wpack, float ln_eps, float *fused, float *space, float *freq,
                                   void *stream) {
    EWVIT_REQUIRE(n >= 0 && d > 0 && heads > 0 && depth > 0, EWVIT_ERR_INVALID_ARG, "ewvit_dama_tail_fwd: bad sizes");
    EWVIT_REQUIRE(d % 32 == 0 && d <= 512 && d % heads == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_dama_tail_fwd: dim must be a multiple of 32 (<= 512) and divisible by heads (got dim=%d heads=%d)", d, heads);
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(space_in && freq_in && wpack && fused && space && freq, EWVIT_ERR_INVALID_ARG, "ewvit_dama_tail_fwd: NULL pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    DamaParams p;
    p.space_in = space_in; p.freq_in = freq_in; p.wpack = wpack; p.fused = fused; p.space = space; p.freq = freq;
    p.n = n; p.d = d; p.heads = heads; p.depth = depth; p.ln_eps = ln_eps;
    const size_t smem = (size_t)(kFpc * d * 12 + kFpc * 2 * d + kFpc * (d / 2) + kFpc * 4) * sizeof(float);
    int groups = (kMaxGroups * 128) / d;          // thread groups of d threads within the 512-thread launch bound
    if (groups > kMaxGroups) groups = kMaxGroups;
    if (groups < 1) groups = 1;
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dama_tail_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        EWVIT_CUDA_OK(cudaFuncSetAttribute(dama_tail_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (d == 128) dama_tail_kernel<128><<<(unsigned)((n + kFpc - 1) / kFpc), groups * d, smem, (cudaStream_t)stream>>>(p);
    else dama_tail_kernel<0><<<(unsigned)((n + kFpc - 1) / kFpc), groups * d, smem, (cudaStream_t)stream>>>(p);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

extern "C" int ewvit_video_head_fwd(const float *fused, const float *space, const float *freq, int64_t videos, int k, int d,
                                    float *mean_fused, float *mean_space, float *mean_freq, const float *cw1,
                                    const float *cb1, const float *cw2, const float *cb2, int hc, float *logits,
                                    void *stream) {
    EWVIT_REQUIRE(videos >= 0 && k > 0 && d > 0, EWVIT_ERR_INVALID_ARG, "ewvit_video_head_fwd: bad sizes");
    if (videos == 0) return EWVIT_OK;
    EWVIT_REQUIRE(fused && mean_fused, EWVIT_ERR_INVALID_ARG, "ewvit_video_head_fwd: NULL pointer");
    EWVIT_REQUIRE((space == nullptr) == (mean_space == nullptr) && (freq == nullptr) == (mean_freq == nullptr),
                  EWVIT_ERR_INVALID_ARG, "ewvit_video_head_fwd: each optional input needs its output");
    EWVIT_REQUIRE(!cw1 || (cb1 && cw2 && cb2 && logits && hc > 0), EWVIT_ERR_INVALID_ARG,
                  "ewvit_video_head_fwd: incomplete classifier arguments");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const size_t smem = (size_t)(d + (hc > 0 ? hc : 0)) * sizeof(float);
    video_head_kernel<<<(unsigned)videos, 128, smem, (cudaStream_t)stream>>>(fused, space, freq, k, d, mean_fused, mean_space,
                                                                            mean_freq, cw1, cb1, cw2, cb2, hc, logits);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
