// bf16 tcgen05/TMEM implicit-GEMM kernel shared by the MWT 3x3 convolutions, patch_to_embedding and
// the ViT linears.
//
//   D[128 x 128] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem, K-major, SW128)  *  B[128 x 64]^T
//
// "Tap streaming": the K loop runs over (tap, 64-channel chunk) pairs.  For a plain GEMM there is
// one tap; for a 3x3 convolution over an NHWC activation there are nine, and the A tile of a tap is
// simply the same 128 output pixels shifted by the tap offset -- fetched by TMA straight from the
// activation tensor (no im2col buffer), with out-of-bounds pixels zero-filled by the TMA unit:
//   * A_FLAT   : activation is [rows, C] with an explicit zero border around every image
//                ("padded-flat" NHWC [N, H+2, W+2, C]); a tap is a constant ROW SHIFT of the tile;
//   * A_TILE4D : activation is [N, H, W, C]; the tile is a box_h x box_w pixel patch, a tap is a
//                (dy, dx) shift of the box origin, stride-2 convs use the TMA element stride.
// Warp roles (320 threads, persistent CTAs, one per SM):
//   warp 0 lane 0 : TMA producer      (6-stage smem ring, full/empty mbarriers)
//   warp 1 lane 0 : tcgen05.mma issuer (2 accumulator stages in TMEM, 2 x 128 columns)
//   warps 2..9    : epilogue           (tcgen05.ld -> scale/shift/act/residual -> global); the epilogue
//                   flavour is a template parameter so the conv path stays ~3 instructions/element
#include "ewvit_tc.cuh"

#include <mutex>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr int kTileBytes = BM * BK * 2;        // 16 KiB
constexpr int kStageBytes = 2 * kTileBytes;    // A + B
constexpr int kEpiWarps = 8;                   // 2 warps per TMEM lane quarter, each takes half the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;  // 320
enum { EPI_CONV = 0, EPI_LINEAR = 1, EPI_PARTIAL = 2, EPI_BB = 3 };
constexpr uint32_t kTmemCols = kAccStages * BN;   // 256
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

enum { A_FLAT = 0, A_TILE4D = 1 };

struct GemmParams {
    int a_mode;
    int chunks_per_tap, num_kb;
    int tap_a0[9];   // FLAT: row shift of the tap.  TILE4D: x offset of the tap
    int tap_a1[9];   // TILE4D: y offset of the tap
    long long M;     // FLAT: number of valid rows
    int N;
    int tiles_m, tiles_n, splits, kb_per_split;
    // TILE4D geometry
    int tiles_x, tiles_y, box_w, box_h, in_stride;
    int out_w, out_h;
    long long out_img_rows;
    int out_wp, out_pad;
    // epilogue
    void *out;
    int out_fp32;
    long long ldo;
    int col_off;
    const float *scale;
    const float *shift;
    int act;
    const float *residual;
    long long ldr;
    int pad_hp, pad_wp;
    float *partial;
    const __nv_bfloat16 *residual_bf16;   // EPI_BB: optional bf16 residual, same layout as the output
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    return v;
}

template <int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + kStages * kStageBytes);
    unsigned long long *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages,
                       *tempty = bars + 2 * kStages + kAccStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2 * kAccStages);
    __shared__ __align__(16) float s_scale[BN], s_shift[BN];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ewvit::tma_prefetch_desc(&tmA);
        ewvit::tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) {
            ewvit::mbar_init(ewvit::smem_u32(&full[s]), 1);
            ewvit::mbar_init(ewvit::smem_u32(&empty[s]), 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            ewvit::mbar_init(ewvit::smem_u32(&tfull[a]), 1);
            ewvit::mbar_init(ewvit::smem_u32(&tempty[a]), kEpiWarps);
        }
        ewvit::mbar_fence_init();
    }
    if (warp == 1) {
        ewvit::tmem_alloc(ewvit::smem_u32(tmem_slot), kTmemCols);
        ewvit::tmem_relinquish();
    }
    ewvit::tc_fence_before();
    __syncthreads();
    ewvit::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long total_work = (long long)p.tiles_m * p.tiles_n * p.splits;
    const uint32_t smem_base = ewvit::smem_u32(smem);

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int n_t = (int)(w % p.tiles_n);
                const long long wm = w / p.tiles_n;
                const int m_t = (int)(wm % p.tiles_m);
                const int sp = (int)(wm / p.tiles_m);
                const int kb0 = sp * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                int tx = 0, ty = 0, img = 0;
                if (p.a_mode == A_TILE4D) {
                    tx = m_t % p.tiles_x;
                    const int t2 = m_t / p.tiles_x;
                    ty = t2 % p.tiles_y;
                    img = t2 / p.tiles_y;
                }
                int tap = kb0 / p.chunks_per_tap, chunk = kb0 - tap * p.chunks_per_tap;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ewvit::mbar_wait(ewvit::smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t bar = ewvit::smem_u32(&full[stage]);
                    ewvit::mbar_expect_tx(bar, kStageBytes);
                    const uint32_t a_dst = smem_base + stage * kStageBytes;
                    if (p.a_mode == A_FLAT) {
                        ewvit::tma_load_2d(a_dst, &tmA, chunk * BK, m_t * BM + p.tap_a0[tap], bar);
                    } else {
                        ewvit::tma_load_4d(a_dst, &tmA, chunk * BK, tx * p.box_w * p.in_stride + p.tap_a0[tap],
                                           ty * p.box_h * p.in_stride + p.tap_a1[tap], img, bar);
                    }
                    ewvit::tma_load_2d(a_dst + kTileBytes, &tmB, kb * BK, n_t * BN, bar);
                    if (++chunk == p.chunks_per_tap) { chunk = 0; ++tap; }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int sp = (int)((w / p.tiles_n) / p.tiles_m);
                const int kb0 = sp * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                // columns past N are zero-filled B rows: shrink the MMA's N to the valid part (multiple of 16)
                const int n_valid = min(BN, p.N - (int)(w % p.tiles_n) * BN);
                const uint32_t idesc = ewvit::umma_idesc_bf16(BM, (uint32_t)((n_valid + 15) & ~15));
                ewvit::mbar_wait(ewvit::smem_u32(&tempty[acc]), acc_phase ^ 1);
                ewvit::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ewvit::mbar_wait(ewvit::smem_u32(&full[stage]), phase);
                    ewvit::tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * kStageBytes;
                    const uint32_t b_addr = a_addr + kTileBytes;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        ewvit::umma_bf16(d_tmem, ewvit::umma_desc_sw128(a_addr + k * 32),
                                         ewvit::umma_desc_sw128(b_addr + k * 32), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ewvit::umma_commit(ewvit::smem_u32(&empty[stage]));   // frees the smem slot when the MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                ewvit::umma_commit(ewvit::smem_u32(&tfull[acc]));          // accumulator ready for the epilogue
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..9)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int hc = (warp - 2) >> 2;         // which half of the 128 columns this warp drains
        const int r = q * 32 + lane;            // row of the tile owned by this thread
        const int etid = threadIdx.x - 64;      // 0..255 among the epilogue threads
        int acc = 0;
        uint32_t acc_phase = 0;
        int cur_nt = -1;
        for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int n_t = (int)(w % p.tiles_n);
            const long long wm = w / p.tiles_n;
            const int m_t = (int)(wm % p.tiles_m);
            const int sp = (int)(wm / p.tiles_m);

            if ((kEpi == EPI_CONV || kEpi == EPI_BB) && n_t != cur_nt) {   // (re)stage the per-channel scale/shift of this column tile
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (etid < BN) {
                    const bool in = n_t * BN + etid < p.N;
                    s_scale[etid] = (p.scale && in) ? p.scale[n_t * BN + etid] : 1.f;
                    s_shift[etid] = (p.shift && in) ? p.shift[n_t * BN + etid] : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                cur_nt = n_t;
            }

            bool valid, zero = false;
            long long orow;
            if (p.a_mode == A_FLAT) {
                orow = (long long)m_t * BM + r;
                valid = orow < p.M;
                if (p.pad_hp > 0) {
                    const long long img_rows = (long long)p.pad_hp * p.pad_wp;
                    const int qi = (int)(orow % img_rows);
                    const int y = qi / p.pad_wp, x = qi - y * p.pad_wp;
                    zero = (y == 0) || (y == p.pad_hp - 1) || (x == 0) || (x == p.pad_wp - 1);
                }
            } else {
                const int tx = m_t % p.tiles_x;
                const int t2 = m_t / p.tiles_x;
                const int ty = t2 % p.tiles_y, img = t2 / p.tiles_y;
                const int oy = ty * p.box_h + r / p.box_w, ox = tx * p.box_w + r % p.box_w;
                valid = (oy < p.out_h) && (ox < p.out_w);
                orow = (long long)img * p.out_img_rows + (long long)(oy + p.out_pad) * p.out_wp + ox + p.out_pad;
            }

            ewvit::mbar_wait(ewvit::smem_u32(&tfull[acc]), acc_phase);
            ewvit::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c = hc * 2 + cc;
                if (kEpi == EPI_BB && n_t * BN + c * 32 >= p.N) continue;   // warp-uniform: nothing valid in this chunk
                uint32_t v[32];
                ewvit::tmem_ld_32x32(t_row + c * 32, v);
                ewvit::tmem_ld_wait();
                const int col0 = n_t * BN + c * 32;
                if (!valid) continue;
                if (kEpi == EPI_BB) {
                    // bias (+ SiLU / ReLU) (+ bf16 residual) -> bf16, 8 columns (16 bytes) at a time, masked past N
                    __nv_bfloat16 *orow_p = static_cast<__nv_bfloat16 *>(p.out) + orow * p.ldo + p.col_off;
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = col0 + g8 * 8;
                        if (col >= p.N) break;
                        float f8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) f8[i] = __uint_as_float(v[g8 * 8 + i]) + s_shift[c * 32 + g8 * 8 + i];
                        if (p.act == 3) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f8[i] = __fdividef(f8[i], 1.f + __expf(-f8[i]));
                        } else if (p.act == 1) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f8[i] = fmaxf(f8[i], 0.f);
                        }
                        if (p.residual_bf16) {   // skip connection is added AFTER the activation (FusedMBConv/MBConv)
                            const uint4 rv = *reinterpret_cast<const uint4 *>(p.residual_bf16 + orow * p.ldr + col);
                            const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rv);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 r2 = __bfloat1622float2(rp[i]);
                                f8[2 * i] += r2.x;
                                f8[2 * i + 1] += r2.y;
                            }
                        }
                        uint4 pk;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(f8[0], f8[1]), b1 = __floats2bfloat162_rn(f8[2], f8[3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(f8[4], f8[5]), b3 = __floats2bfloat162_rn(f8[6], f8[7]);
                        pk.x = *reinterpret_cast<uint32_t *>(&b0);
                        pk.y = *reinterpret_cast<uint32_t *>(&b1);
                        pk.z = *reinterpret_cast<uint32_t *>(&b2);
                        pk.w = *reinterpret_cast<uint32_t *>(&b3);
                        *reinterpret_cast<uint4 *>(orow_p + col) = pk;
                    }
                    continue;
                }
                if (kEpi == EPI_PARTIAL) {
                    float4 *dst = reinterpret_cast<float4 *>(p.partial + ((long long)sp * p.M + orow) * p.N + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                             __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    continue;
                }
                float f[32];
                if (kEpi == EPI_CONV) {
                    const float lo = p.act ? 0.f : -INFINITY;
                    const float keep = zero ? 0.f : 1.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sc = *reinterpret_cast<const float4 *>(&s_scale[c * 32 + 4 * i]);
                        const float4 sh = *reinterpret_cast<const float4 *>(&s_shift[c * 32 + 4 * i]);
                        f[4 * i + 0] = fmaxf(fmaf(__uint_as_float(v[4 * i + 0]), sc.x, sh.x), lo) * keep;
                        f[4 * i + 1] = fmaxf(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y), lo) * keep;
                        f[4 * i + 2] = fmaxf(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z), lo) * keep;
                        f[4 * i + 3] = fmaxf(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w), lo) * keep;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float x = __uint_as_float(v[i]);
                        if (p.scale) x *= __ldg(p.scale + col0 + i);
                        if (p.shift) x += __ldg(p.shift + col0 + i);
                        if (p.residual) x += __ldg(p.residual + orow * p.ldr + col0 + i);
                        f[i] = x;
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = 0.5f * f[i] * (1.f + erff(f[i] * 0.70710678118654752f));
                    }
                }
                if (kEpi == EPI_LINEAR && p.out_fp32) {
                    float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(p.out) + orow * p.ldo + p.col_off + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                } else {
                    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(p.out) + orow * p.ldo + p.col_off + col0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 pk;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(f[8 * i + 0], f[8 * i + 1]);
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(f[8 * i + 2], f[8 * i + 3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[8 * i + 4], f[8 * i + 5]);
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(f[8 * i + 6], f[8 * i + 7]);
                        pk.x = *reinterpret_cast<uint32_t *>(&b0);
                        pk.y = *reinterpret_cast<uint32_t *>(&b1);
                        pk.z = *reinterpret_cast<uint32_t *>(&b2);
                        pk.w = *reinterpret_cast<uint32_t *>(&b3);
                        dst[i] = pk;
                    }
                }
            }
            ewvit::tc_fence_before();
            __syncwarp();
            if (lane == 0) ewvit::mbar_arrive(ewvit::smem_u32(&tempty[acc]));
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }

    ewvit::tc_fence_before();
    __syncthreads();
    if (warp == 1) ewvit::tmem_dealloc(tmem_base, kTmemCols);
}

// Split-K second pass: sum the fp32 partials, then the same epilogue as the fused path.
__global__ void splitk_reduce_kernel(const float *__restrict__ partial, int splits, long long M, int N,
                                     const float *__restrict__ scale, const float *__restrict__ shift, int act,
                                     const float *__restrict__ residual, long long ldr, void *out, int out_fp32,
                                     long long ldo) {
    const long long total = M * (N / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (N / 4);
        const int col = (int)(i % (N / 4)) * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < splits; ++k) {
            const float4 v = *reinterpret_cast<const float4 *>(partial + ((long long)k * M + row) * N + col);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        float f[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x = f[j];
            if (scale) x *= scale[col + j];
            if (shift) x += shift[col + j];
            if (residual) x += residual[row * ldr + col + j];
            f[j] = apply_act(x, act);
        }
        if (out_fp32) {
            *reinterpret_cast<float4 *>(static_cast<float *>(out) + row * ldo + col) = make_float4(f[0], f[1], f[2], f[3]);
        } else {
            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t *>(&b0);
            pk.y = *reinterpret_cast<uint32_t *>(&b1);
            *reinterpret_cast<uint2 *>(static_cast<__nv_bfloat16 *>(out) + row * ldo + col) = pk;
        }
    }
}

template <int kEpi>
int launch_gemm_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    long long work = (long long)p.tiles_m * p.tiles_n * p.splits;
    long long grid = ewvit_num_sms();
    if (grid > work) grid = work;
    if (grid <= 0) return EWVIT_OK;
    gemm_tc_kernel<kEpi><<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(tmA, tmB, p);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int epi, cudaStream_t stream) {
    if (epi == EPI_CONV) return launch_gemm_t<EPI_CONV>(tmA, tmB, p, stream);
    if (epi == EPI_BB) return launch_gemm_t<EPI_BB>(tmA, tmB, p, stream);
    if (epi == EPI_PARTIAL) return launch_gemm_t<EPI_PARTIAL>(tmA, tmB, p, stream);
    return launch_gemm_t<EPI_LINEAR>(tmA, tmB, p, stream);
}

}  // namespace

// ------------------------------------------------------------------ tensor-map encoding (host)
ewvit_encode_tiled_fn ewvit_get_encode_tiled() {
    static ewvit_encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<ewvit_encode_tiled_fn>(ptr);
    });
    return fn;
}

int ewvit_make_tmap_bf16(CUtensorMap *out, const void *base, int rank, const uint64_t *dims,
                         const uint64_t *strides_bytes, const uint32_t *box, const uint32_t *estr) {
    ewvit_encode_tiled_fn enc = ewvit_get_encode_tiled();
    EWVIT_REQUIRE(enc != nullptr, EWVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        es[i] = estr ? estr[i] : 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i + 1];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EWVIT_REQUIRE(r == CUDA_SUCCESS, EWVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return EWVIT_OK;
}

// ------------------------------------------------------------------ C ABI
extern "C" int ewvit_linear_bf16(const void *a, const void *w, int64_t M, int N, int K, const float *scale,
                                 const float *shift, int act, const float *residual, int64_t ldr, void *out,
                                 int out_fp32, int64_t ldo, int splits, float *workspace, void *stream) {
    EWVIT_REQUIRE(M >= 0 && N > 0 && K > 0, EWVIT_ERR_INVALID_ARG, "ewvit_linear_bf16: bad sizes M=%lld N=%d K=%d", (long long)M, N, K);
    if (M == 0) return EWVIT_OK;
    EWVIT_REQUIRE(a && w && out, EWVIT_ERR_INVALID_ARG, "ewvit_linear_bf16: NULL pointer");
    EWVIT_REQUIRE(K % BK == 0 && N % BN == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_linear_bf16: needs K %% 64 == 0 and N %% 128 == 0 (got N=%d K=%d)", N, K);
    EWVIT_REQUIRE(act >= 0 && act <= 2, EWVIT_ERR_INVALID_ARG, "ewvit_linear_bf16: act must be 0 (none), 1 (relu) or 2 (gelu)");
    EWVIT_REQUIRE(ewvit_aligned16(a) && ewvit_aligned16(w) && ewvit_aligned16(out) && ewvit_aligned16(workspace) &&
                      ewvit_aligned16(residual), EWVIT_ERR_INVALID_ARG, "ewvit_linear_bf16: pointers must be 16-byte aligned");
    EWVIT_REQUIRE(ldo % 8 == 0 && ldo >= N && (!residual || (ldr % 4 == 0 && ldr >= N)), EWVIT_ERR_INVALID_ARG,
                  "ewvit_linear_bf16: ldo must be a multiple of 8 and >= N, ldr a multiple of 4 and >= N");
    const int num_kb = K / BK;
    if (splits < 1) splits = 1;
    if (splits > num_kb) splits = num_kb;
    EWVIT_REQUIRE(splits == 1 || workspace, EWVIT_ERR_INVALID_ARG, "ewvit_linear_bf16: split-K needs a workspace of splits*M*N floats");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;

    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {2, (uint64_t)K * 2};
        uint32_t box[2] = {BK, BM};
        rc = ewvit_make_tmap_bf16(&tmA, a, 2, dims, str, box, nullptr);
        if (rc != EWVIT_OK) return rc;
        uint64_t dimsb[2] = {(uint64_t)K, (uint64_t)N};
        uint32_t boxb[2] = {BK, BN};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, str, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
    }
    GemmParams p = {};
    p.a_mode = A_FLAT;
    p.chunks_per_tap = num_kb;
    p.num_kb = num_kb;
    p.M = M;
    p.N = N;
    p.tiles_m = (int)((M + BM - 1) / BM);
    p.tiles_n = N / BN;
    p.kb_per_split = (num_kb + splits - 1) / splits;
    p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;   // every split gets >= 1 k-block
    p.out = out; p.out_fp32 = out_fp32; p.ldo = ldo; p.col_off = 0;
    p.scale = scale; p.shift = shift; p.act = act; p.residual = residual; p.ldr = ldr;
    p.partial = p.splits > 1 ? workspace : nullptr;
    rc = launch_gemm(tmA, tmB, p, p.splits > 1 ? EPI_PARTIAL : EPI_LINEAR, (cudaStream_t)stream);
    if (rc != EWVIT_OK) return rc;
    if (p.splits > 1) {
        const long long total = M * (N / 4);
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)ewvit_num_sms() * 8;
        if (blocks > cap) blocks = cap;
        splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(workspace, p.splits, M, N, scale, shift, act,
                                                                                 residual, ldr, out, out_fp32, ldo);
        EWVIT_LAUNCH_OK();
    }
    return EWVIT_OK;
}

extern "C" int ewvit_conv3x3_bf16(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int stride,
                                  int in_padded, const float *scale, const float *shift, int relu, void *y,
                                  int y_ldc, int y_coff, int out_padded, int force_tiled, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_conv3x3_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && y, EWVIT_ERR_INVALID_ARG, "ewvit_conv3x3_bf16: NULL pointer");
    EWVIT_REQUIRE(stride == 1 || stride == 2, EWVIT_ERR_UNSUPPORTED, "ewvit_conv3x3_bf16: stride must be 1 or 2");
    EWVIT_REQUIRE(cin % BK == 0 && cout % BN == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv3x3_bf16: needs cin %% 64 == 0 and cout %% 128 == 0 (got cin=%d cout=%d)", cin, cout);
    EWVIT_REQUIRE(y_ldc % 8 == 0 && y_coff % 8 == 0 && y_coff + cout <= y_ldc, EWVIT_ERR_INVALID_ARG,
                  "ewvit_conv3x3_bf16: bad output channel pitch/offset");
    EWVIT_REQUIRE(ewvit_aligned16(x) && ewvit_aligned16(w) && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG,
                  "ewvit_conv3x3_bf16: pointers must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;

    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    const int hin = in_padded ? h + 2 : h, win = in_padded ? wd + 2 : wd;
    const int chunks = cin / BK;
    GemmParams p = {};
    p.chunks_per_tap = chunks;
    p.num_kb = 9 * chunks;
    p.N = cout;
    p.tiles_n = cout / BN;
    p.splits = 1;
    p.kb_per_split = p.num_kb;
    p.out = y; p.out_fp32 = 0; p.ldo = y_ldc; p.col_off = y_coff;
    p.scale = scale; p.shift = shift; p.act = relu ? 1 : 0;

    CUtensorMap tmA, tmB;
    {
        uint64_t dimsb[2] = {(uint64_t)9 * cin, (uint64_t)cout}, strb[2] = {2, (uint64_t)9 * cin * 2};
        uint32_t boxb[2] = {BK, BN};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
    }
    const bool flat = (stride == 1) && in_padded && out_padded && !force_tiled;
    if (flat) {
        const long long rows = (long long)n * hin * win;
        uint64_t dims[2] = {(uint64_t)cin, (uint64_t)rows}, str[2] = {2, (uint64_t)cin * 2};
        uint32_t box[2] = {BK, BM};
        rc = ewvit_make_tmap_bf16(&tmA, x, 2, dims, str, box, nullptr);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_FLAT;
        p.M = rows;
        p.tiles_m = (int)((rows + BM - 1) / BM);
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) p.tap_a0[dy * 3 + dx] = (dy - 1) * win + (dx - 1);
        p.pad_hp = hin;
        p.pad_wp = win;
    } else {
        const int box_w = 16, box_h = 8;
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)win, (uint64_t)hin, (uint64_t)n};
        uint64_t str[4] = {2, (uint64_t)cin * 2, (uint64_t)win * cin * 2, (uint64_t)hin * win * cin * 2};
        uint32_t box[4] = {BK, (uint32_t)(box_w * stride), (uint32_t)(box_h * stride), 1};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        rc = ewvit_make_tmap_bf16(&tmA, x, 4, dims, str, box, es);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_TILE4D;
        p.box_w = box_w; p.box_h = box_h; p.in_stride = stride;
        p.tiles_x = (wo + box_w - 1) / box_w;
        p.tiles_y = (ho + box_h - 1) / box_h;
        p.tiles_m = p.tiles_x * p.tiles_y * n;
        p.out_w = wo; p.out_h = ho;
        p.out_pad = out_padded ? 1 : 0;
        p.out_wp = wo + 2 * p.out_pad;
        p.out_img_rows = (long long)(ho + 2 * p.out_pad) * p.out_wp;
        p.M = (long long)n * p.out_img_rows;
        const int off = in_padded ? 0 : -1;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                p.tap_a0[dy * 3 + dx] = dx + off;
                p.tap_a1[dy * 3 + dx] = dy + off;
            }
    }
    return launch_gemm(tmA, tmB, p, EPI_CONV, (cudaStream_t)stream);
}


// General NHWC bf16 convolution for the EfficientNet backbone (1x1 or 3x3/pad 1, stride 1 or 2) with a fused
// bias + optional bf16 residual + activation epilogue.  Channel counts only need to be multiples of 8: the K and N
// tails are zero-filled by TMA (boxes may overhang the tensor) and masked in the epilogue.
extern "C" int ewvit_conv_nhwc_bf16(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int ksize,
                                    int stride, const float *bias, int act, const void *residual, void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_conv_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && y, EWVIT_ERR_INVALID_ARG, "ewvit_conv_nhwc_bf16: NULL pointer");
    EWVIT_REQUIRE((ksize == 1 && stride == 1) || (ksize == 3 && (stride == 1 || stride == 2)), EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv_nhwc_bf16: supports 1x1/stride 1 and 3x3/stride 1|2 (got k=%d s=%d)", ksize, stride);
    EWVIT_REQUIRE(cin % 8 == 0 && cout % 8 == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv_nhwc_bf16: channel counts must be multiples of 8 (got cin=%d cout=%d)", cin, cout);
    EWVIT_REQUIRE(act == 0 || act == 1 || act == 3, EWVIT_ERR_INVALID_ARG, "ewvit_conv_nhwc_bf16: act must be 0 (none), 1 (relu) or 3 (silu)");
    EWVIT_REQUIRE(ewvit_aligned16(x) && ewvit_aligned16(w) && ewvit_aligned16(y) && ewvit_aligned16(residual), EWVIT_ERR_INVALID_ARG,
                  "ewvit_conv_nhwc_bf16: pointers must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;

    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    GemmParams p = {};
    p.N = cout;
    p.tiles_n = (cout + BN - 1) / BN;
    p.splits = 1;
    p.out = y; p.out_fp32 = 0; p.ldo = cout; p.col_off = 0;
    p.shift = bias; p.act = act;
    p.residual_bf16 = static_cast<const __nv_bfloat16 *>(residual);
    p.ldr = cout;
    CUtensorMap tmA, tmB;
    if (ksize == 1) {
        const long long rows = (long long)n * h * wd;
        const int num_kb = (cin + BK - 1) / BK;
        uint64_t dims[2] = {(uint64_t)cin, (uint64_t)rows}, str[2] = {2, (uint64_t)cin * 2};
        uint32_t box[2] = {BK, BM};
        rc = ewvit_make_tmap_bf16(&tmA, x, 2, dims, str, box, nullptr);
        if (rc != EWVIT_OK) return rc;
        uint64_t dimsb[2] = {(uint64_t)cin, (uint64_t)cout};
        uint32_t boxb[2] = {BK, BN};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, str, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_FLAT;
        p.chunks_per_tap = num_kb;
        p.num_kb = num_kb;
        p.kb_per_split = num_kb;
        p.M = rows;
        p.tiles_m = (int)((rows + BM - 1) / BM);
    } else {
        const int chunks = (cin + BK - 1) / BK;
        const int kpad = chunks * BK;
        uint64_t dimsb[2] = {(uint64_t)9 * kpad, (uint64_t)cout}, strb[2] = {2, (uint64_t)9 * kpad * 2};
        uint32_t boxb[2] = {BK, BN};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
        const int box_w = 16, box_h = 8;
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
        uint64_t str[4] = {2, (uint64_t)cin * 2, (uint64_t)wd * cin * 2, (uint64_t)h * wd * cin * 2};
        uint32_t box[4] = {BK, (uint32_t)(box_w * stride), (uint32_t)(box_h * stride), 1};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        rc = ewvit_make_tmap_bf16(&tmA, x, 4, dims, str, box, es);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_TILE4D;
        p.chunks_per_tap = chunks;
        p.num_kb = 9 * chunks;
        p.kb_per_split = p.num_kb;
        p.box_w = box_w; p.box_h = box_h; p.in_stride = stride;
        p.tiles_x = (wo + box_w - 1) / box_w;
        p.tiles_y = (ho + box_h - 1) / box_h;
        p.tiles_m = p.tiles_x * p.tiles_y * n;
        p.out_w = wo; p.out_h = ho;
        p.out_pad = 0;
        p.out_wp = wo;
        p.out_img_rows = (long long)ho * wo;
        p.M = (long long)n * p.out_img_rows;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                p.tap_a0[dy * 3 + dx] = dx - 1;
                p.tap_a1[dy * 3 + dx] = dy - 1;
            }
    }
    return launch_gemm(tmA, tmB, p, EPI_BB, (cudaStream_t)stream);
}
