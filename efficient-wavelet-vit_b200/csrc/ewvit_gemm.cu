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
//   warps 2..9    : epilogue           (tcgen05.ld -> scale/shift/act/residual -> global) in two groups of four
//                   warps, one per accumulator stage, so two tiles drain concurrently (short-K tiles are
//                   bounded by the epilogue's latency chain); the flavour is a template parameter
//   warps 10..13  : A-tile builders    (A_IM2COL only: small-channel 3x3 convs assemble K = 9*cin rows from a halo)
#include "ewvit_tc.cuh"

#include <mutex>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 12;          // barrier slots; the ring depth actually used is GemmParams::stages
constexpr int kMaxAccStages = 4;
constexpr int kTileBytes = BM * BK * 2;        // 16 KiB
enum { EPI_CONV = 0, EPI_LINEAR = 1, EPI_PARTIAL = 2, EPI_BB = 3 };
constexpr int kBuilderWarps = 8;                // A_IM2COL / A_SCALED only: 256 threads assemble or rescale the A tiles in shared memory
constexpr int kBuilderSlots = 4;                // ring slots are owned by builder-warp PAIRS (64 rows each)
constexpr int kMaxHalo = 8;                     // A_IM2COL: halo ring depth (TMA latency of the small-row boxes is ~3 us)
constexpr int kFlat3Rows = BM + 8;               // 136 rows = 17 swizzle atoms: rows [m0 - 1, m0 + 135) cover the shifts 0..2
constexpr int kFlat3ABytes = kFlat3Rows * BK * 2; // 17408
constexpr int kStgBytes = 32 * 64;               // one epilogue chunk: 32 rows x 32 bf16, dense, 64B-swizzled (TMA store box)
constexpr int kGateFrames = 4;                   // A_SCALED: SE gates of the (<= 4) frames an M tile touches, staged in shared memory
constexpr int kGateMaxK = 1536;
constexpr int kGateBytes = kGateFrames * kGateMaxK * 2;
constexpr int kGateChunks = kGateBytes / 16 / (32 * kBuilderWarps);   // 16-byte chunks per builder thread
// Epilogue warps come in groups of four (one warp per TMEM lane quarter).  Kernels without builder warps run FOUR
// groups: two per accumulator stage, each draining half of the tile's columns -- short-K layers are bound by the
// epilogue's latency chain (tcgen05.ld -> MUFU -> staging -> TMA store, ~1200 cycles per 32x32 chunk with two warps per
// scheduler), and twice the warps hide twice the latency.  Only the backbone's plain 1x1 convs (EPI_BB without
// builders: K <= 256) do this: their operand ring can be shallow, which pays for the second set of staging buffers.
// Builder kernels keep two groups (register file); the long-K MWT convs and the linears keep the deep ring.
// The three-level MWT head conv (EPI_CONV with a 192-column tile = 3 levels x 64 channels: 27 K = 16 MMAs per tile) is all epilogue
// as well -- 128 x 192 outputs per ~900 cycles of MMA -- and gets the four groups, too.
template <int kEpi, bool kBuilder, int kBN = 128>
struct Cfg {
    static constexpr bool kHead3 = kEpi == EPI_CONV && kBN == 192;
    static constexpr bool kWideEpi = (kEpi == EPI_BB && !kBuilder) || kHead3;
    static constexpr int kEpiWarps = kWideEpi ? 16 : 8;
    static constexpr int kEpiGroups = kEpiWarps / 4;
    static constexpr int kHalves = kEpiGroups / 2;                       // column halves per accumulator stage
    static constexpr int kThreads = 64 + 32 * kEpiWarps + (kBuilder ? 32 * kBuilderWarps : 0);
    // per-warp staging buffers of the epilogue; the MWT conv epilogue double-buffers them (the TMA store of chunk c reads
    // buffer c & 1 while chunk c + 1 is converted and staged: a single buffer made every chunk wait for the previous store's read)
    static constexpr int kStgBufs = kEpi == EPI_CONV ? 2 : 1;
    // TMEM accumulator stages.  The MWT convs (128 columns per tile) use all 512 columns = 4 stages: the two epilogue groups have the
    // throughput (one tile per ~2500 cycles each against one per ~2200 of the issuer) but with two stages the issuer still waited
    // 300-900 cycles per tile for the group two tiles back; four stages absorb that latency
    static constexpr int kAccStages = (kEpi == EPI_CONV && !kHead3) ? 4 : 2;
    static constexpr int kStagingBytes = kEpiWarps * kStgBytes * kStgBufs;
    static constexpr int kOperandBytes = (kHead3 ? 144 : kWideEpi ? 160 : kBuilder ? 192 : kEpi == EPI_CONV ? 184 : 200) * 1024;  // operand stages (+ halo buffers)
    static constexpr int kGateOff = kOperandBytes + kStagingBytes;       // builder kernels: staged SE gates
    static constexpr int kPayloadBytes = kGateOff + (kBuilder ? kGateBytes : 0);  // barriers live right behind
    static constexpr int kSmemBytes = kPayloadBytes + 1024 /*align slack*/ + 512 /*barriers*/;
    // 227 KB opt-in limit covers dynamic + static shared memory (s_scale/s_shift for kBN = 256: 4 KB, + the builder tables)
    static_assert(kSmemBytes + 4096 + 512 <= 227 * 1024, "operand region too large: cudaFuncSetAttribute would fail");
};

enum { A_FLAT = 0, A_TILE4D = 1, A_IM2COL = 2, A_SCALED = 3 };

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
    // A_IM2COL (3x3 conv, cin <= 64): the (box_h*s+2) x (box_w*s+2) input halo of a tile is fetched ONCE by TMA and
    // builder warps assemble the densely packed K = 9*cin operand rows from it (tile geometry = A_TILE4D fields)
    int stages;          // operand ring depth actually used (<= kStages)
    int cin;             // channels per pixel in the halo
    int halo_w, halo_h;  // halo extent in pixels
    int halo_bytes;      // bytes of one halo buffer (TMA transaction size)
    int halo_stride;     // 1024-aligned distance between consecutive halo buffers
    int halo_bufs;       // halo ring depth (<= kMaxHalo)
    // A_SCALED (1x1 conv behind a squeeze-excitation block): the A tile arrives by TMA as for A_FLAT and builder warps
    // multiply it in place by the per-(frame, channel) gate before the MMA consumes it, so the separate x *= gate
    // pass (one read + one write of the expanded tensor) disappears; tile geometry = A_FLAT
    const void *a_gate;    // [frames, K] fp32, or bf16 when a_gate_bf16
    int a_gate_bf16;
    int a_hw;              // rows (pixels) per frame
    int a_k;               // K = row pitch of A and of the gate
    int k16;               // 2 = the three-level MWT head conv: flat3 with 32 channels per pixel, a pixel is ONE 64-byte row = two MMA K steps;
                           // the window of a vertical tap is 136 rows x 64 bytes (64B swizzle), the weight tiles are [128, 16] (32B swizzle)
    int pair;              // CTA pairs on cta_group::2 (MWT flat3 convs, backbone 1x1 convs): every CTA stores b_half filter rows per B tile
    int b_half;            // pair mode: rows of a B tile held by each CTA = half of the MMA's N
    int a_gate_smem;       // the gates of a tile's frames are staged in shared memory once per tile (bf16 gates, <= kGateFrames frames per tile)
    int flat3;           // A_FLAT 3x3 conv, "row-shared" taps: one ring slot = (dy, channel chunk) holds ONE window of
                         // kFlat3Rows activation rows and the three weight tiles of dx = 0,1,2; the three A operands are the
                         // same window read at a start address shifted by 0/1/2 rows, so every activation row crosses
                         // L2 -> shared memory 3 times per tile instead of 9
    int b_tile_bytes;    // bytes of one B tile = rows of the tmB box x 128 (0 = kBN rows)
    int stage_bytes;     // ring slot size: A tile (+ B tile unless the weights are resident)
    int b_res;           // A_IM2COL: all k-blocks of B stay resident in shared memory (loaded once per CTA)
    int bres_off;        // byte offset of the resident B region
    int halo_nb;         // the halo is fetched as halo_nb boxes of halo_ppb pixels x halo_h rows ("planes"): rows of
    int halo_ppb;        // halo_ppb*cin contiguous elements keep the TMA row count ~10x lower than per-pixel rows
    int dbg;             // debug experiments (ewvit_debug_set_flags): 1 = skip epilogue stores, 2 = skip activation, 4 = skip staging transpose
    long long *trace;    // debug: clock64 stamps of CTA 0, [6 roles][64 tiles][4] (ewvit_debug_set_trace)
    int plane_bytes;     // distance between planes in shared memory (128-byte aligned, >= halo_h*halo_ppb*cin*2)
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    return v;
}

// bf16 store of one 32-row x 32-column chunk held row-per-lane (pk = this lane's 32 values packed) through the TMA
// unit: the warp writes the chunk into its shared-memory staging buffer (dense 64-byte rows, 64B-swizzled so the
// 16-byte writes are conflict-free) and one lane issues a bulk tensor store.  The LSU never sees the scattered
// row-per-lane pattern (32 cache lines x 16 bytes per STG measured ~15k cycles per 128x256 tile), stores drain
// asynchronously while the warp computes the next chunk, and rows/columns outside the tensor are clipped by TMA.
template <int kPending = 0>
__device__ __forceinline__ void stage_chunk_bf16(uint32_t stg, int lane, const uint32_t (&pk)[16]) {
    // the store that last used THIS buffer has finished reading it (kPending = stores allowed to be still in flight)
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
    __syncwarp();
    const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 64 + (((uint32_t)i ^ sw) << 4)), "r"(pk[4 * i]),
                     "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                     : "memory");
    ewvit::fence_proxy_async();
    __syncwarp();
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const void *tmap, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(src),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// work item w of this CTA -> (row tile, column tile, K split); always true for the m-fastest order
__device__ __forceinline__ bool decode_tile(const GemmParams &p, int w, int &m_t, int &n_t, int &sp) {
    m_t = w % p.tiles_m;            // m fastest: a CTA keeps its column tile (and its staged bias) for many consecutive tiles
    const int wn = w / p.tiles_m;
    n_t = wn % p.tiles_n;
    sp = wn / p.tiles_n;
    return true;
}

#define EWVIT_TRACE(role, tile, k)                                                                  \
    do {                                                                                            \
        if (p.trace && blockIdx.x == 0 && (tile) < 64) p.trace[((role) * 64 + (tile)) * 4 + (k)] = clock64(); \
    } while (0)

// kFast (EPI_BB only): 0 = generic epilogue (runtime activation / residual switches), 1 = SiLU on pre-halved operands
// without residual, 2 = no activation (+ optional bf16 residual).  The specialised bodies are ONE basic block per 32-column chunk
// (32 independent values per lane for the scheduler) -- the generic one branches every 8 columns and ran at ~1/3 IPC.
// kPair (MWT 3x3 convs on the row-shared-tap path only): CTA pairs on cta_group::2.  The two CTAs of a 2-CTA cluster take two
// consecutive 128-row tiles; each loads its own activation window and HALF of every weight tile (64 of the 128 filters), the
// leader issues M = 256 MMAs for both and the completions are multicast onto both CTAs' barriers.  Per MMA a CTA's tensor core
// reads 4 KB of A and 2 KB of B from its shared memory instead of 4 + 4, and the ring slot shrinks from 65 KB to 41 KB
// (multiscale conv: 3 -> 4 stages; fusion conv with resident weights: 72 KB instead of 144 KB resident, 3 -> 7 windows in flight).
template <int kEpi, bool kBuilder, int kBN, int kFast = 0, bool kPair = false>
__global__ void __launch_bounds__(Cfg<kEpi, kBuilder, kBN>::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    constexpr int kEpiWarps = Cfg<kEpi, kBuilder, kBN>::kEpiWarps, kHalves = Cfg<kEpi, kBuilder, kBN>::kHalves;
    constexpr int kOperandBytes = Cfg<kEpi, kBuilder, kBN>::kOperandBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + Cfg<kEpi, kBuilder, kBN>::kPayloadBytes);
    constexpr int kAccStages = Cfg<kEpi, kBuilder, kBN>::kAccStages;
    unsigned long long *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages,
                       *tempty = bars + 2 * kStages + kMaxAccStages;
    unsigned long long *hfull = bars + 2 * kStages + 2 * kMaxAccStages, *hempty = hfull + kMaxHalo;
    unsigned long long *bres_bar = hempty + kMaxHalo;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bres_bar + 1);
    const uint32_t kBTileB = (uint32_t)p.b_tile_bytes;    // B tile: b_rows x 64 bf16 (b_rows = kBN, or the 16-aligned N of a single column tile)
    constexpr uint32_t kTmemColsT = kAccStages * kBN <= 256 ? 256 : 512;      // allocations are powers of two (2 x 192 -> 512)
    __shared__ __align__(16) float s_scale[2 * kBN], s_shift[2 * kBN];
    __shared__ int s_koff[kBuilder ? 9 * 8 : 1];
    __shared__ __align__(16) uint32_t s_zero[4];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler, too
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = kPair ? ewvit::cluster_ctarank() : 0u;   // 0 = leader of the CTA pair (issues the MMAs)
    const int nstages = p.stages;
    const uint32_t kStageB = (uint32_t)p.stage_bytes;    // A tile + B tile (A only when B is resident)

    if (threadIdx.x == 0) {
        ewvit::tma_prefetch_desc(&tmA);
        ewvit::tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) {
            // [TMA B] + the two warps of the slot's builder pair (A_SCALED on CTA pairs: the builder pairs of BOTH CTAs)
            ewvit::mbar_init(ewvit::smem_u32(&full[s]), !kBuilder ? 1 : p.a_mode == A_SCALED ? (kPair ? 4 : 2) : (p.b_res ? 2 : 3));
            ewvit::mbar_init(ewvit::smem_u32(&empty[s]), 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            ewvit::mbar_init(ewvit::smem_u32(&tfull[a]), 1);
            ewvit::mbar_init(ewvit::smem_u32(&tempty[a]), kPair ? kEpiWarps : kEpiWarps / 2);   // pair: the epilogue warps of BOTH CTAs
        }
        ewvit::mbar_init(ewvit::smem_u32(bres_bar), 1);
        for (int a = 0; a < kMaxHalo; ++a) {
            ewvit::mbar_init(ewvit::smem_u32(&hfull[a]), 1);
            ewvit::mbar_init(ewvit::smem_u32(&hempty[a]), kBuilderWarps);
        }
        ewvit::mbar_fence_init();
    }
    if (warp == 1) {
        if (kPair) {
            ewvit::tmem_alloc_pair(ewvit::smem_u32(tmem_slot), kTmemColsT);
            ewvit::tmem_relinquish_pair();
        } else {
            ewvit::tmem_alloc(ewvit::smem_u32(tmem_slot), kTmemColsT);
            ewvit::tmem_relinquish();
        }
    }
    ewvit::tc_fence_before();
    if (kPair) ewvit::cluster_sync_all();     // both CTAs' barriers are initialised before any remote arrive / multicast commit
    else __syncthreads();
    ewvit::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_work = p.tiles_m * p.tiles_n * p.splits;
    // pair mode: work items are PAIRS of row tiles; cluster c takes items c, c + #clusters, ...; CTA rank r the r-th tile of the pair
    const int tiles_m2 = (p.tiles_m + 1) >> 1;
    const int pair_work = tiles_m2 * p.tiles_n;
    const int w_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int w_total = kPair ? pair_work : total_work;
    auto decode = [&](int w, int &m_t, int &n_t, int &sp) -> bool {
        if (kPair) {
            m_t = 2 * (w % tiles_m2) + (int)cta_rank;      // may be one past the last tile (odd tile count): TMA zero-fills / clips it
            n_t = w / tiles_m2;
            sp = 0;
            return true;
        }
        return decode_tile(p, w, m_t, n_t, sp);
    };
    const uint32_t smem_base = ewvit::smem_u32(smem);

    // The two issuing roles run WARP-UNIFORM loops (all 32 lanes track the same counters and poll the same
    // barriers) and only the instruction issue itself is predicated on one elected lane.  With a lane-0-only loop
    // every operand is a divergent value and the compiler wraps each UTMALDG/UTCHMMA in an ELECT + R2UR
    // "uniformisation" loop (~570 cycles per k-block measured, more than the 256 cycles its MMAs take).
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        int stage = 0, hb = 0;
        uint32_t phase = 0, hphase = 0;
        int tt = 0;
        if (p.b_res) {     // weights are small and identical for every tile: fetch all k-blocks once
            // flat3: num_kb counts (dy, chunk) slots of three taps each; the three-level head conv keeps two K = 16 tiles per tap
            const int nb = p.k16 == 2 ? 6 * p.num_kb : p.flat3 ? 3 * p.num_kb : p.num_kb;
            const int kstep = p.k16 ? 16 : BK;
            if (ewvit::elect_one()) {
                const uint32_t bb = ewvit::smem_u32(bres_bar);
                if (kPair) {           // each CTA keeps ITS half of the filter rows; both halves count on the leader's barrier
                    if (cta_rank == 0) ewvit::mbar_expect_tx(bb, (uint32_t)(2 * nb * kBTileB));
                    for (int kb = 0; kb < nb; ++kb)
                        ewvit::tma_load_2d_pair(smem_base + p.bres_off + kb * kBTileB, &tmB, kb * BK, (int)cta_rank * p.b_half, bb);
                } else {
                    ewvit::mbar_expect_tx(bb, (uint32_t)(nb * kBTileB));
                    for (int kb = 0; kb < nb; ++kb)
                        ewvit::tma_load_2d(smem_base + p.bres_off + kb * kBTileB, &tmB, kb * kstep, 0, bb);
                }
            }
            __syncwarp();
        }
        for (int w = w_first; w < w_total; w += w_step, ++tt) {
            if (lane == 0) EWVIT_TRACE(0, tt, 0);
            int m_t, n_t, sp;
            if (!decode(w, m_t, n_t, sp)) continue;
            const int kb0 = sp * p.kb_per_split;
            const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            int tx = 0, ty = 0, img = 0;
            if (p.a_mode == A_TILE4D || p.a_mode == A_IM2COL) {
                tx = m_t % p.tiles_x;
                const int t2 = m_t / p.tiles_x;
                ty = t2 % p.tiles_y;
                img = t2 / p.tiles_y;
            }
            if (kBuilder && p.a_mode == A_IM2COL) {
                // the tile's input halo (out-of-image pixels zero-filled) as a few wide boxes; then only B per k-block
                ewvit::mbar_wait(ewvit::smem_u32(&hempty[hb]), hphase ^ 1);
                const uint32_t hbar = ewvit::smem_u32(&hfull[hb]);
                const uint32_t hdst = smem_base + nstages * kStageB + hb * p.halo_stride;
                const int hx0 = (tx * p.box_w * p.in_stride - 1) * p.cin, hy0 = ty * p.box_h * p.in_stride - 1;
                if (ewvit::elect_one()) {
                    ewvit::mbar_expect_tx(hbar, (uint32_t)p.halo_bytes);
                    for (int b = 0; b < p.halo_nb; ++b)
                        ewvit::tma_load_3d(hdst + b * p.plane_bytes, &tmA, hx0 + b * p.halo_ppb * p.cin, hy0, img, hbar);
                }
                __syncwarp();
                if (++hb == p.halo_bufs) { hb = 0; hphase ^= 1; }
                for (int kb = kb0; kb < kb1 && !p.b_res; ++kb) {
                    ewvit::mbar_wait(ewvit::smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t bar = ewvit::smem_u32(&full[stage]);
                    const uint32_t b_dst = smem_base + stage * kStageB + kTileBytes;
                    if (ewvit::elect_one()) {
                        ewvit::mbar_expect_tx(bar, kBTileB);
                        ewvit::tma_load_2d(b_dst, &tmB, kb * BK, n_t * kBN, bar);
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (lane == 0) EWVIT_TRACE(0, tt, 1);
                continue;
            }
            int tap = kb0 / p.chunks_per_tap, chunk = kb0 - tap * p.chunks_per_tap;
            const int ax0 = tx * p.box_w * p.in_stride, ay0 = ty * p.box_h * p.in_stride;
            if (p.flat3) {
                for (int kb = kb0; kb < kb1; ++kb) {       // kb = dy * chunks + chunk
                    const int dy = kb / p.chunks_per_tap, ch = kb - dy * p.chunks_per_tap;
                    ewvit::mbar_wait(ewvit::smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t bar = ewvit::smem_u32(&full[stage]);
                    const uint32_t a_dst = smem_base + stage * kStageB;
                    if (kPair) {
                        if (ewvit::elect_one()) {
                            // both CTAs' loads of this slot complete on the LEADER's full barrier (it expects the bytes of the pair)
                            if (cta_rank == 0) ewvit::mbar_expect_tx(bar, 2 * kStageB);
                            ewvit::tma_load_2d_pair(a_dst, &tmA, ch * BK, m_t * BM + p.tap_a0[dy * 3], bar);
                            if (!p.b_res) {
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx)
                                    ewvit::tma_load_2d_pair(a_dst + kFlat3ABytes + dx * kBTileB, &tmB, ((dy * 3 + dx) * p.chunks_per_tap + ch) * BK,
                                                            n_t * kBN + (int)cta_rank * p.b_half, bar);
                            }
                        }
                    } else if (ewvit::elect_one()) {
                        ewvit::mbar_expect_tx(bar, kStageB);
                        ewvit::tma_load_2d(a_dst, &tmA, ch * BK, m_t * BM + p.tap_a0[dy * 3], bar);
                        if (!p.b_res) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                ewvit::tma_load_2d(a_dst + kFlat3ABytes + dx * kBTileB, &tmB, ((dy * 3 + dx) * p.chunks_per_tap + ch) * BK, n_t * kBN, bar);
                        }
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (lane == 0) EWVIT_TRACE(0, tt, 1);
                continue;
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                ewvit::mbar_wait(ewvit::smem_u32(&empty[stage]), phase ^ 1);
                // A_SCALED: the builders post-process the raw tile, so completion goes to the `raw` (halo) barrier
                const uint32_t bar = ewvit::smem_u32((kBuilder && p.a_mode == A_SCALED) ? &hfull[stage] : &full[stage]);
                const uint32_t a_dst = smem_base + stage * kStageB;
                if (kPair && !(kBuilder && p.a_mode == A_SCALED)) {
                    // plain 1x1 conv on a CTA pair: both CTAs' tiles of this slot complete on the LEADER's full barrier
                    if (ewvit::elect_one()) {
                        if (cta_rank == 0) ewvit::mbar_expect_tx(bar, 2 * kStageB);
                        if (p.a_mode == A_TILE4D)      // stride-1|2 box path on a pair: each CTA fetches the pixel box of ITS tile
                            ewvit::tma_load_4d_pair(a_dst, &tmA, chunk * BK, ax0 + p.tap_a0[tap], ay0 + p.tap_a1[tap], img, bar);
                        else
                            ewvit::tma_load_2d_pair(a_dst, &tmA, chunk * BK, m_t * BM + p.tap_a0[tap], bar);
                        if (!p.b_res) ewvit::tma_load_2d_pair(a_dst + kTileBytes, &tmB, kb * BK, n_t * kBN + (int)cta_rank * p.b_half, bar);
                    }
                } else if (ewvit::elect_one()) {
                    // (A_SCALED on a CTA pair: each CTA's raw tile + its half of B complete on its OWN raw barrier; its builders
                    // then arrive on the leader's full barrier)
                    ewvit::mbar_expect_tx(bar, kStageB);
                    if (p.a_mode == A_FLAT || p.a_mode == A_SCALED)
                        ewvit::tma_load_2d(a_dst, &tmA, chunk * BK, m_t * BM + p.tap_a0[tap], bar);
                    else
                        ewvit::tma_load_4d(a_dst, &tmA, chunk * BK, ax0 + p.tap_a0[tap], ay0 + p.tap_a1[tap], img, bar);
                    if (!p.b_res) ewvit::tma_load_2d(a_dst + kTileBytes, &tmB, kb * BK, n_t * kBN + (kPair ? (int)cta_rank * p.b_half : 0), bar);
                }
                __syncwarp();
                if (++chunk == p.chunks_per_tap) { chunk = 0; ++tap; }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) EWVIT_TRACE(0, tt, 1);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (pair mode: the leader CTA only)
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        int tt = 0;
        if (p.b_res && cta_rank == 0) ewvit::mbar_wait(ewvit::smem_u32(bres_bar), 0);
        for (int w = w_first; w < w_total && cta_rank == 0; w += w_step, ++tt) {
            if (lane == 0) EWVIT_TRACE(1, tt, 0);
            int m_t, n_t, sp;
            if (!decode(w, m_t, n_t, sp)) continue;
            (void)m_t;
            const int kb0 = sp * p.kb_per_split;
            const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            // columns past N are zero-filled B rows: shrink the MMA's N to the valid part (multiple of 16)
            const int n_valid = min(kBN, p.N - n_t * kBN);
            // pair mode splits B by HALF OF THE MMA's N: with several column tiles every CTA holds kBN / 2 rows, so N stays kBN
            // (rows past the tensor are zero-filled)
            const uint32_t idesc = ewvit::umma_idesc_bf16(kPair ? 2 * BM : BM, p.k16 == 2 ? 128u : (kPair && p.tiles_n > 1) ? (uint32_t)kBN : (uint32_t)((n_valid + 15) & ~15));
            ewvit::mbar_wait(ewvit::smem_u32(&tempty[acc]), acc_phase ^ 1);
            ewvit::tc_fence_after();
            if (lane == 0) EWVIT_TRACE(1, tt, 1);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBN);
            for (int kb = kb0; kb < kb1; ++kb) {
                ewvit::mbar_wait(ewvit::smem_u32(&full[stage]), phase);
                if (kb == kb0 && lane == 0) EWVIT_TRACE(1, tt, 2);
                ewvit::tc_fence_after();
                const uint32_t a_addr = smem_base + stage * kStageB;
                const uint64_t a_desc = ewvit::umma_desc_sw128(a_addr);
                const uint64_t b_desc = ewvit::umma_desc_sw128(p.b_res ? smem_base + p.bres_off + kb * kBTileB : a_addr + kTileBytes);
                const uint32_t ebar = ewvit::smem_u32(&empty[stage]);
                const uint32_t first = kb > kb0 ? 1u : 0u;
                if (p.flat3 && p.k16 == 2) {
                    // three wavelet levels side by side: a pixel is one 64-byte row of 32 channels = [level 0..2][9 subbands] (+ 5 of
                    // padding), i.e. TWO K steps: level 0 lives in step 0, level 2 in step 1, level 1 straddles both.  Per tap
                    // (dy = kb, dx) TWO K = 16, N = 128 MMAs read the 64B-swizzled window at a start address shifted by dx rows:
                    // step 0 against [level-0 | level-1] copies of the shared head weights into accumulator columns [0, 128), step 1
                    // against [level-1 | level-2] copies into columns [64, 192).  (An MMA costs its operand reads -- 4 KB of A whatever
                    // N is -- so twelve N = 64 MMAs per slot ran at ~70 cycles each and bounded the kernel.)  Columns [64, 128) are
                    // written by both steps: the very first step-1 MMA of a tile is split so that [128, 192) starts from zero.
                    if (ewvit::elect_one()) {
                        const uint32_t idesc64 = ewvit::umma_idesc_bf16(BM, 64u);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t ad = ewvit::umma_desc_sw64(a_addr + dx * 64);
                            const uint32_t bt = smem_base + p.bres_off + (uint32_t)((kb * 3 + dx) * 2) * kBTileB;
                            const bool first = dx == 0 && kb == kb0;
                            ewvit::umma_bf16(d_tmem, ad, ewvit::umma_desc_sw32(bt), idesc, first ? 0u : 1u);
                            if (first) {
                                ewvit::umma_bf16(d_tmem + 64u, ad + 2, ewvit::umma_desc_sw32(bt + kBTileB), idesc64, 1u);
                                ewvit::umma_bf16(d_tmem + 128u, ad + 2, ewvit::umma_desc_sw32(bt + kBTileB + 64u * 32u), idesc64, 0u);
                            } else {
                                ewvit::umma_bf16(d_tmem + 64u, ad + 2, ewvit::umma_desc_sw32(bt + kBTileB), idesc, 1u);
                            }
                        }
                        ewvit::umma_commit(ebar);
                    }
                } else if (p.flat3) {
                    if (ewvit::elect_one()) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            // the operand of tap dx is the window shifted by dx rows (128 bytes): the 128B swizzle is a function
                            // of the shared-memory address, so a row-shifted start address reads what TMA wrote for row r + dx
                            // (verified on hardware; the descriptor's base-offset field must stay 0 for this)
                            const uint64_t ad = ewvit::umma_desc_sw128(a_addr + dx * 128);
                            const int dy = kb / p.chunks_per_tap, ch = kb - dy * p.chunks_per_tap;
                            const uint64_t bd = ewvit::umma_desc_sw128(
                                p.b_res ? smem_base + p.bres_off + ((dy * 3 + dx) * p.chunks_per_tap + ch) * kBTileB : a_addr + kFlat3ABytes + dx * kBTileB);
                            if (kPair) {
                                ewvit::umma_bf16_pair(d_tmem, ad, bd, idesc, (dx > 0 || kb > kb0) ? 1u : 0u);
                                ewvit::umma_bf16_pair(d_tmem, ad + 2, bd + 2, idesc, 1u);
                                ewvit::umma_bf16_pair(d_tmem, ad + 4, bd + 4, idesc, 1u);
                                ewvit::umma_bf16_pair(d_tmem, ad + 6, bd + 6, idesc, 1u);
                            } else {
                                ewvit::umma_bf16(d_tmem, ad, bd, idesc, (dx > 0 || kb > kb0) ? 1u : 0u);
                                ewvit::umma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1u);
                                ewvit::umma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1u);
                                ewvit::umma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1u);
                            }
                        }
                        if (kPair) ewvit::umma_commit_pair(ebar);      // frees this slot in BOTH CTAs
                        else ewvit::umma_commit(ebar);
                    }
                } else if (ewvit::elect_one()) {
                    // advancing 16 K-elements = 32 bytes = +2 in the descriptor's (address >> 4) field
                    if (kPair) {
                        ewvit::umma_bf16_pair(d_tmem, a_desc, b_desc, idesc, first);
                        ewvit::umma_bf16_pair(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
                        ewvit::umma_bf16_pair(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
                        ewvit::umma_bf16_pair(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
                        ewvit::umma_commit_pair(ebar);
                    } else {
                        ewvit::umma_bf16(d_tmem, a_desc, b_desc, idesc, first);
                        ewvit::umma_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
                        ewvit::umma_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
                        ewvit::umma_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
                        ewvit::umma_commit(ebar);   // frees the smem slot when the MMAs retire
                    }
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            if (ewvit::elect_one()) {                                                    // accumulator ready for the epilogue(s)
                if (kPair) ewvit::umma_commit_pair(ewvit::smem_u32(&tfull[acc]));
                else ewvit::umma_commit(ewvit::smem_u32(&tfull[acc]));
            }
            __syncwarp();
            if (lane == 0) EWVIT_TRACE(1, tt, 3);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (kBuilder && warp >= 2 + kEpiWarps) {
        // ---------------------------------------------------------------- A-tile builders (A_IM2COL)
        // thread r owns output pixel r of the tile: for every k-block it gathers eight 16-byte channel chunks
        // (k = tap*cin + c, dense) from the halo and stores them 128B-swizzled, exactly as TMA would have.
        // Builder warp b assembles every 4th k-block on its own (all 128 rows, 4 rows per lane), so four k-blocks
        // are in flight and the proxy fence of one overlaps the copies of the others.
        const int bwarp = warp - (2 + kEpiWarps);
        const int btid = threadIdx.x - 32 * (2 + kEpiWarps);
        if (p.a_mode == A_SCALED) {
            // The producer warp fetches the raw A tile (and B) by TMA exactly as for a plain 1x1 conv, but signals the
            // `raw` barrier (the halo barriers, unused in this mode); the 256 builder threads multiply the tile by the
            // squeeze-excitation gate IN PLACE (swizzled 16-byte chunks, 4 rows per thread) and only then hand the slot
            // to the tensor core.  Global traffic stays asynchronous and the separate x *= gate pass is gone.
            // Ring slots are owned by builder-warp PAIRS (slot & 3 == pair), so up to four k-blocks are rescaled
            // concurrently and the gate loads / fence of one overlap the copies of the others.
            const int pair = bwarp & (kBuilderSlots - 1);
            const int t64 = (bwarp >> 2) * 32 + lane;   // thread within the pair
            const int j = t64 & 7;                      // 16-byte chunk (8 channels) of the 64-channel k-block
            const int rb = t64 >> 3;                    // rows rb + 8 i, i = 0..15  (all have the same swizzle phase rb)
            const uint32_t chunk_off = (uint32_t)rb * 128u + (((uint32_t)j ^ (uint32_t)rb) << 4);
            const int last_frame = (int)((p.M - 1) / p.a_hw);
            uint32_t g = 0;
            // Gates in shared memory (bf16 gates, the engine's case).  Fetching them from global memory per k-block was the
            // stall of this kernel (ncu: issue-active 19 %, 16 dependent L2 round trips per builder thread and k-block): now
            // the [frames of the tile][K] gate block is staged once per tile -- its 16-byte chunks are requested one tile
            // AHEAD into registers, so the copy costs two named barriers and no exposed latency.
            const bool gsm = p.a_gate_smem != 0;
            const uint32_t gate_sm = smem_base + (uint32_t)Cfg<kEpi, kBuilder, kBN>::kGateOff;
            const int kchunks = p.a_k >> 3;                       // 16-byte chunks per gate row
            uint4 gpre[kGateChunks];
            auto gate_prefetch = [&](int w2) {
                int m2 = 0, n2 = 0, s2 = 0;
                if (w2 < w_total) decode(w2, m2, n2, s2);
                const int f0n = (int)(((long long)m2 * BM) / p.a_hw);
#pragma unroll
                for (int q = 0; q < kGateChunks; ++q) {
                    const int ch = btid + q * (32 * kBuilderWarps);
                    const int fr = ch / kchunks, cc = ch - fr * kchunks;
                    gpre[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (fr < kGateFrames && w2 < w_total)
                        gpre[q] = __ldg(reinterpret_cast<const uint4 *>(static_cast<const __nv_bfloat16 *>(p.a_gate) +
                                                                        (long long)min(f0n + fr, last_frame) * p.a_k) + cc);
                }
            };
            if (gsm) gate_prefetch(w_first);
            for (int w = w_first; w < w_total; w += w_step) {
                int m_t, n_t_unused, sp;
                decode(w, m_t, n_t_unused, sp);
                const int kb0 = sp * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                // gate row (frame) of each of this thread's 16 tile rows; rows past M are zero-filled by TMA
                int goff[16];
                {
                    const long long row0 = (long long)m_t * BM + rb;
                    const int f0 = (int)(((long long)m_t * BM) / p.a_hw);
                    int fr = (int)(row0 / p.a_hw);
                    int rem = (int)(row0 - (long long)fr * p.a_hw);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        goff[i] = gsm ? (min(fr, last_frame) - f0) * p.a_k * 2 + j * 16      // byte offset inside the staged block
                                      : min(fr, last_frame) * p.a_k + j * 8;
                        rem += 8;
                        while (rem >= p.a_hw) { rem -= p.a_hw; ++fr; }
                    }
                }
                if (gsm) {
                    asm volatile("bar.sync 3, %0;" ::"n"(32 * kBuilderWarps) : "memory");   // every builder is done with the previous tile's gates
#pragma unroll
                    for (int q = 0; q < kGateChunks; ++q) {
                        const int ch = btid + q * (32 * kBuilderWarps);
                        if (ch < kGateFrames * kchunks)
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(gate_sm + (uint32_t)ch * 16u), "r"(gpre[q].x), "r"(gpre[q].y),
                                         "r"(gpre[q].z), "r"(gpre[q].w) : "memory");
                    }
                    asm volatile("bar.sync 3, %0;" ::"n"(32 * kBuilderWarps) : "memory");
                    gate_prefetch(w + w_step);
                }
                for (int kb = kb0; kb < kb1; ++kb, ++g) {
                    const int stage = (int)(g % (uint32_t)nstages);
                    if ((stage & (kBuilderSlots - 1)) != pair) continue;
                    const uint32_t phase = (g / (uint32_t)nstages) & 1u;
                    const bool kin = kb * BK + j * 8 < p.a_k;               // K tail: the tile holds zeros there
                    const uint32_t a_base = smem_base + stage * kStageB + chunk_off;
                    if (gsm) {
                        ewvit::mbar_wait(ewvit::smem_u32(&hfull[stage]), phase);
                        const uint32_t gk = gate_sm + (uint32_t)kb * (BK * 2);
#pragma unroll
                        for (int b4 = 0; b4 < 2; ++b4) {
                            uint4 gq[8], v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                gq[i] = make_uint4(0u, 0u, 0u, 0u);
                                if (kin)
                                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(gq[i].x), "=r"(gq[i].y), "=r"(gq[i].z), "=r"(gq[i].w)
                                                 : "r"(gk + (uint32_t)goff[b4 * 8 + i]) : "memory");
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                             : "r"(a_base + (b4 * 8 + i) * 1024) : "memory");
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const __nv_bfloat162 *vp = reinterpret_cast<const __nv_bfloat162 *>(&v[i]);
                                const __nv_bfloat162 *gp2 = reinterpret_cast<const __nv_bfloat162 *>(&gq[i]);
                                const __nv_bfloat162 o0 = __hmul2(vp[0], gp2[0]), o1 = __hmul2(vp[1], gp2[1]);
                                const __nv_bfloat162 o2 = __hmul2(vp[2], gp2[2]), o3 = __hmul2(vp[3], gp2[3]);
                                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + (b4 * 8 + i) * 1024),
                                             "r"(*reinterpret_cast<const uint32_t *>(&o0)), "r"(*reinterpret_cast<const uint32_t *>(&o1)),
                                             "r"(*reinterpret_cast<const uint32_t *>(&o2)), "r"(*reinterpret_cast<const uint32_t *>(&o3))
                                             : "memory");
                            }
                        }
                    } else
                    if (p.a_gate_bf16) {
                        // bf16 gates: one 16-byte load and four packed multiplies per 8-channel chunk (the product of two bf16
                        // values is exact in fp32, so HMUL2.BF16 rounds exactly like the fp32 path)
                        const __nv_bfloat16 *gk = static_cast<const __nv_bfloat16 *>(p.a_gate) + kb * BK;
                        bool waited = false;
#pragma unroll
                        for (int b4 = 0; b4 < 2; ++b4) {
                            uint4 gq[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                gq[i] = make_uint4(0u, 0u, 0u, 0u);
                                if (kin) gq[i] = __ldg(reinterpret_cast<const uint4 *>(gk + goff[b4 * 8 + i]));
                            }
                            if (!waited) {
                                ewvit::mbar_wait(ewvit::smem_u32(&hfull[stage]), phase);
                                waited = true;
                            }
                            uint4 v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                             : "r"(a_base + (b4 * 8 + i) * 1024) : "memory");
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const __nv_bfloat162 *vp = reinterpret_cast<const __nv_bfloat162 *>(&v[i]);
                                const __nv_bfloat162 *gp2 = reinterpret_cast<const __nv_bfloat162 *>(&gq[i]);
                                const __nv_bfloat162 o0 = __hmul2(vp[0], gp2[0]), o1 = __hmul2(vp[1], gp2[1]);
                                const __nv_bfloat162 o2 = __hmul2(vp[2], gp2[2]), o3 = __hmul2(vp[3], gp2[3]);
                                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + (b4 * 8 + i) * 1024),
                                             "r"(*reinterpret_cast<const uint32_t *>(&o0)), "r"(*reinterpret_cast<const uint32_t *>(&o1)),
                                             "r"(*reinterpret_cast<const uint32_t *>(&o2)), "r"(*reinterpret_cast<const uint32_t *>(&o3))
                                             : "memory");
                            }
                        }
                    } else {
                    const float *gk = static_cast<const float *>(p.a_gate) + kb * BK;
                    bool waited = false;
#pragma unroll
                    for (int b4 = 0; b4 < 4; ++b4) {
                        float4 g0[4], g1[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            g0[i] = g1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (kin) {
                                g0[i] = __ldg(reinterpret_cast<const float4 *>(gk + goff[b4 * 4 + i]));
                                g1[i] = __ldg(reinterpret_cast<const float4 *>(gk + goff[b4 * 4 + i]) + 1);
                            }
                        }
                        if (!waited) {
                            ewvit::mbar_wait(ewvit::smem_u32(&hfull[stage]), phase);
                            waited = true;
                        }
                        uint4 v[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                         : "r"(a_base + (b4 * 4 + i) * 1024) : "memory");
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const __nv_bfloat162 *vp = reinterpret_cast<const __nv_bfloat162 *>(&v[i]);
                            const float2 f0 = __bfloat1622float2(vp[0]), f1 = __bfloat1622float2(vp[1]);
                            const float2 f2 = __bfloat1622float2(vp[2]), f3 = __bfloat1622float2(vp[3]);
                            const __nv_bfloat162 o0 = __floats2bfloat162_rn(f0.x * g0[i].x, f0.y * g0[i].y);
                            const __nv_bfloat162 o1 = __floats2bfloat162_rn(f1.x * g0[i].z, f1.y * g0[i].w);
                            const __nv_bfloat162 o2 = __floats2bfloat162_rn(f2.x * g1[i].x, f2.y * g1[i].y);
                            const __nv_bfloat162 o3 = __floats2bfloat162_rn(f3.x * g1[i].z, f3.y * g1[i].w);
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + (b4 * 4 + i) * 1024),
                                         "r"(*reinterpret_cast<const uint32_t *>(&o0)), "r"(*reinterpret_cast<const uint32_t *>(&o1)),
                                         "r"(*reinterpret_cast<const uint32_t *>(&o2)), "r"(*reinterpret_cast<const uint32_t *>(&o3))
                                         : "memory");
                        }
                    }
                    }
                    ewvit::fence_proxy_async();      // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        if (kPair) ewvit::mbar_arrive_leader(ewvit::smem_u32(&full[stage]));   // the issuer waits for the builders of both CTAs
                        else ewvit::mbar_arrive(ewvit::smem_u32(&full[stage]));
                    }
                }
            }
        } else {
        // byte offset (relative to a pixel's halo origin) of every 16-byte chunk of the dense K axis, -1 = zero pad;
        // the same for every tile, so the divisions are done once
        for (int i = btid; i < p.num_kb * 8; i += 32 * kBuilderWarps) {
            const int k = i * 8;
            int off = 3 << 28;                                               // dx code 3 = zero padding of the K axis
            if (k < 9 * p.cin) {
                const int tap = k / p.cin, c = k - tap * p.cin;
                const int dy = tap / 3, dx = tap - dy * 3;
                off = ((dy * p.halo_ppb * p.cin + c) * 2) | (dx << 28);      // row part + channel, dx in the top bits
            }
            s_koff[i] = off;
        }
        if (btid < 4) s_zero[btid] = 0u;
        const uint32_t zero_addr = ewvit::smem_u32(s_zero);
        asm volatile("bar.sync 3, %0;" ::"n"(32 * kBuilderWarps) : "memory");   // ids 1,2 belong to the epilogue groups
        uint32_t row_src[2][3], row_dst[2], row_sw[2];   // row_src[rr][dx]: halo byte offset of pixel (py*s, px*s + dx)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int row = (bwarp >> 2) * 64 + lane + 32 * rr;
            const int py = row / p.box_w, px = row - py * p.box_w;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int hx = px * p.in_stride + dx;
                const int b = hx / p.halo_ppb, xw = hx - b * p.halo_ppb;
                row_src[rr][dx] = (uint32_t)(b * p.plane_bytes + ((py * p.in_stride) * p.halo_ppb + xw) * p.cin * 2);
            }
            row_dst[rr] = (uint32_t)(row * 128);
            row_sw[rr] = (uint32_t)(row & 7);
        }
        uint32_t g = 0;      // k-blocks seen so far by this CTA (all builder warps count the same sequence)
        int hb = 0;
        uint32_t hphase = 0;
        int tt = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tt) {
            const int sp = (w / p.tiles_m) / p.tiles_n;
            const int kb0 = sp * p.kb_per_split;
            const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            ewvit::mbar_wait(ewvit::smem_u32(&hfull[hb]), hphase);
            if (bwarp == 0 && lane == 0) EWVIT_TRACE(4, tt, 0);
            const uint32_t halo = smem_base + nstages * kStageB + hb * p.halo_stride;
            for (int kb = kb0; kb < kb1; ++kb, ++g) {
                // every ring slot is owned by ONE builder warp, so consecutive uses of a slot are ordered by that warp's
                // own waits (two warps sharing a slot could run two phases apart, which a parity wait cannot see)
                const int stage = (int)(g % (uint32_t)nstages);
                if ((stage & (kBuilderSlots - 1)) != (bwarp & (kBuilderSlots - 1))) continue;
                const uint32_t phase = (g / (uint32_t)nstages) & 1u;
                const uint32_t a_base = smem_base + stage * kStageB;
                // the 8 chunk offsets of this k-block are the same for all 128 rows: read the table once
                int off[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) off[j] = s_koff[kb * 8 + j];
                {
                    const int half = bwarp >> 2;          // this warp's 64 rows of the tile
                    uint4 v[2][8];
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2) {
                        const int rr = r2;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            // branch-free gather (predicated selects in PTX: the C++ ternaries compiled to ~3 branches
                            // per chunk, ~1300 cycles per k-block); padding chunks read a 16-byte block of zeros instead
                            asm("{\n\t.reg .pred q1, q2, q3;\n\t.reg .u32 a;\n\t"
                                "setp.ge.s32 q1, %4, 1;\n\tsetp.ge.s32 q2, %4, 2;\n\tsetp.eq.s32 q3, %4, 3;\n\t"
                                "selp.u32 a, %6, %5, q1;\n\tselp.u32 a, %7, a, q2;\n\tadd.u32 a, a, %8;\n\tselp.u32 a, %9, a, q3;\n\t"
                                "ld.shared.v4.u32 {%0, %1, %2, %3}, [a];\n\t}"
                                : "=r"(v[r2][j].x), "=r"(v[r2][j].y), "=r"(v[r2][j].z), "=r"(v[r2][j].w)
                                : "r"(off[j] >> 28), "r"(row_src[rr][0]), "r"(row_src[rr][1]), "r"(row_src[rr][2]),
                                  "r"(halo + (uint32_t)(off[j] & 0x0FFFFFFF)), "r"(zero_addr)
                                : "memory");
                        }
                    }
                    // the gathers above are in flight while we wait for the ring slot to drain
                    (void)half;
                    {
                        if (bwarp == 0 && lane == 0) EWVIT_TRACE(4, tt, 1);
                        ewvit::mbar_wait(ewvit::smem_u32(&empty[stage]), phase ^ 1);
                        if (bwarp == 0 && lane == 0) EWVIT_TRACE(4, tt, 2);
                    }
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2) {
                        const int rr = r2;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + row_dst[rr] + (((uint32_t)j ^ row_sw[rr]) << 4)),
                                         "r"(v[r2][j].x), "r"(v[r2][j].y), "r"(v[r2][j].z), "r"(v[r2][j].w)
                                         : "memory");
                    }
                }
                ewvit::fence_proxy_async();      // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) ewvit::mbar_arrive(ewvit::smem_u32(&full[stage]));
                if (bwarp == 0 && lane == 0) EWVIT_TRACE(4, tt, 3);
            }
            __syncwarp();
            if (lane == 0) ewvit::mbar_arrive(ewvit::smem_u32(&hempty[hb]));
            if (++hb == p.halo_bufs) { hb = 0; hphase ^= 1; }
        }
        }   // A_IM2COL
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..9)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;        // groups of 4 warps: group g drains accumulator stage g & 1, i.e. every
                                                // other tile of this CTA (two tiles in flight in the epilogue), and with
                                                // four groups only column half g >> 1 of it
        const int half = grp >> 1;
        const int r = q * 32 + lane;            // row of the tile owned by this thread
        const int gtid = (threadIdx.x - 64) & 127;
        const int grp2 = grp & 1;               // which of every two consecutive tiles this group drains
        float *g_scale = s_scale + grp2 * kBN, *g_shift = s_shift + grp2 * kBN;
        constexpr int kColsPerGroup = kBN / kHalves;
        const uint32_t stg = smem_base + kOperandBytes + (uint32_t)(warp - 2) * kStgBytes * Cfg<kEpi, kBuilder, kBN>::kStgBufs;
        int cur_nt = -1;
        int it = 0;
        for (int w = w_first; w < w_total; w += w_step, ++it) {
            if ((it & 1) != grp2) continue;
            const int acc = it % kAccStages;                                    // accumulator stage of this tile ...
            const uint32_t acc_phase = (uint32_t)(it / kAccStages) & 1u;        // ... and the parity of its current use
            if (q == 2 && lane == 0) EWVIT_TRACE(2 + grp, it, 0);
            int m_t, n_t, sp;
            if (!decode(w, m_t, n_t, sp)) continue;

            if ((kEpi == EPI_CONV || kEpi == EPI_BB) && n_t != cur_nt) {   // (re)stage the per-channel scale/shift of this column tile
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                for (int i = half * kColsPerGroup + gtid; i < (half + 1) * kColsPerGroup; i += 128) {
                    const bool in = n_t * kBN + i < p.N;
                    g_scale[i] = (p.scale && in) ? p.scale[n_t * kBN + i] : 1.f;
                    g_shift[i] = (p.shift && in) ? p.shift[n_t * kBN + i] : 0.f;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                cur_nt = n_t;
            }

            bool valid, zero = false;
            long long orow;
            int st_x = 0, st_y = 0, st_img = 0;     // TMA-store box origin of this warp's 32 rows (2 pixel rows x 16 pixels)
            if (p.a_mode == A_FLAT || p.a_mode == A_SCALED) {
                orow = (long long)m_t * BM + r;
                valid = orow < p.M;
                if (p.pad_hp > 0) {
                    const long long img_rows = (long long)p.pad_hp * p.pad_wp;
                    // 32-bit arithmetic whenever the row index fits (a 64-bit modulo is ~10x the cost and this runs per tile)
                    const int qi = orow < 0x7fffffffLL ? (int)((unsigned)orow % (unsigned)img_rows) : (int)(orow % img_rows);
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
                st_x = tx * p.box_w + p.out_pad;
                st_y = ty * p.box_h + (q * 32) / p.box_w + p.out_pad;
                st_img = img;
            }
            if (kEpi == EPI_BB && p.residual_bf16 && valid) {   // start pulling the skip-connection row while the MMAs run
                const __nv_bfloat16 *rp = p.residual_bf16 + orow * p.ldr + n_t * kBN;
                const int ncol = min(kBN, p.N - n_t * kBN);
                for (int cb = 0; cb < ncol; cb += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + cb));
            }

            ewvit::mbar_wait(ewvit::smem_u32(&tfull[acc]), acc_phase);
            ewvit::tc_fence_after();
            if (q == 2 && lane == 0) EWVIT_TRACE(2 + grp, it, 1);
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kBN);
            if (kEpi == EPI_CONV && (p.dbg & 512)) {
                // debug: no epilogue work at all (isolates the MMA / operand-feed rate of the MWT convs)
            } else if (kEpi == EPI_CONV) {
                // MWT convs: the chunk loop is unrolled and software-pipelined -- the TMEM load of chunk c+1 is in flight while
                // chunk c goes through BN + ReLU, packing, staging and its TMA store (320-thread kernel: registers to spare)
                constexpr int kChunks = kColsPerGroup / 32;
                const float lo = p.act ? 0.f : -INFINITY;
                const float keep = zero ? 0.f : 1.f;
                uint32_t v2[2][32];
                ewvit::tmem_ld_32x32(t_row + half * kColsPerGroup, v2[0]);
#pragma unroll
                for (int ci = 0; ci < kChunks; ++ci) {
                    const int c = half * kChunks + ci;
                    ewvit::tmem_ld_wait();
                    if (ci + 1 < kChunks) ewvit::tmem_ld_32x32(t_row + (c + 1) * 32, v2[(ci + 1) & 1]);
                    if (n_t * kBN + c * 32 >= p.N) continue;          // warp-uniform: nothing valid in this chunk
                    const uint32_t (&v)[32] = v2[ci & 1];
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sc = *reinterpret_cast<const float4 *>(&g_scale[c * 32 + 4 * i]);
                        const float4 sh = *reinterpret_cast<const float4 *>(&g_shift[c * 32 + 4 * i]);
                        const float f0 = fmaxf(fmaf(__uint_as_float(v[4 * i + 0]), sc.x, sh.x), lo) * keep;
                        const float f1 = fmaxf(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y), lo) * keep;
                        const float f2 = fmaxf(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z), lo) * keep;
                        const float f3 = fmaxf(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w), lo) * keep;
                        const __nv_bfloat162 b0 = __floats2bfloat162_rn(f0, f1), b1 = __floats2bfloat162_rn(f2, f3);
                        pk[2 * i] = *reinterpret_cast<const uint32_t *>(&b0);
                        pk[2 * i + 1] = *reinterpret_cast<const uint32_t *>(&b1);
                    }
                    const uint32_t sbuf = stg + (uint32_t)(ci & 1) * kStgBytes;
                    stage_chunk_bf16<1>(sbuf, lane, pk);
                    if (lane == 0) {
                        const int col0 = n_t * kBN + c * 32;
                        if (p.a_mode == A_FLAT || p.a_mode == A_SCALED) tma_store_2d(&tmC, sbuf, p.col_off + col0, m_t * BM + q * 32);
                        else tma_store_4d(&tmC, sbuf, p.col_off + col0, st_x, st_y, st_img);
                    }
                }
            } else
#pragma unroll 1
            for (int c = half * (kColsPerGroup / 32); c < (half + 1) * (kColsPerGroup / 32); ++c) {
                if ((kEpi == EPI_BB || kEpi == EPI_CONV) && n_t * kBN + c * 32 >= p.N) continue;   // warp-uniform: nothing valid in this chunk
                uint32_t v[32];
                ewvit::tmem_ld_32x32(t_row + c * 32, v);
                // skip-connection values of the whole chunk are requested before the accumulator wait (row-per-lane 16-byte
                // loads: as four dependent L2 round trips they made the residual layers' epilogue ~2900 cycles per chunk)
                uint4 rcur[4];
                if (kEpi == EPI_BB && p.residual_bf16) {
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = n_t * kBN + c * 32 + g8 * 8;
                        rcur[g8] = make_uint4(0u, 0u, 0u, 0u);
                        if (valid && col < p.N) rcur[g8] = __ldg(reinterpret_cast<const uint4 *>(p.residual_bf16 + orow * p.ldr + col));
                    }
                }
                ewvit::tmem_ld_wait();
                const int col0 = n_t * kBN + c * 32;
                if (kEpi == EPI_BB && kFast != 0) {
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sh = *reinterpret_cast<const float4 *>(&g_shift[c * 32 + 4 * i]);
                        f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + sh.x;
                        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + sh.y;
                        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + sh.z;
                        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + sh.w;
                    }
                    if (kFast == 1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float th;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(f[i]));
                            f[i] = fmaf(f[i], th, f[i]);
                        }
                    } else if (p.residual_bf16) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rcur[g8]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 r2 = __bfloat1622float2(rp[i]);
                                f[g8 * 8 + 2 * i] += r2.x;
                                f[g8 * 8 + 2 * i + 1] += r2.y;
                            }
                        }
                    }
                    if (zero) {              // padded-flat output: the one-pixel border stays zero
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = 0.f;
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __nv_bfloat162 bb = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                        pk[i] = *reinterpret_cast<const uint32_t *>(&bb);
                    }
                    stage_chunk_bf16(stg, lane, pk);
                    if (lane == 0) {
                        if (p.a_mode == A_FLAT || p.a_mode == A_SCALED) tma_store_2d(&tmC, stg, p.col_off + col0, m_t * BM + q * 32);
                        else tma_store_4d(&tmC, stg, p.col_off + col0, st_x, st_y, st_img);
                    }
                    continue;
                }
                if (kEpi == EPI_BB) {
                    // bias (+ SiLU / ReLU) (+ bf16 residual) -> bf16; columns past N are computed but never stored
                    uint32_t pk[16];
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = col0 + g8 * 8;
                        float f8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) f8[i] = __uint_as_float(v[g8 * 8 + i]) + g_shift[c * 32 + g8 * 8 + i];
                        if (p.act == 4) {        // SiLU on pre-halved operands: f8 = v/2, silu(v) = h*tanh(h) + h  (2 ops per value)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float th;
                                asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(f8[i]));
                                f8[i] = fmaf(f8[i], th, f8[i]);
                            }
                        } else if (p.act == 3 && !(p.dbg & 2)) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f8[i] = ewvit::silu_fast(f8[i]);
                        } else if (p.act == 1) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f8[i] = fmaxf(f8[i], 0.f);
                        }
                        if (zero) {              // padded-flat output: the one-pixel border stays zero
#pragma unroll
                            for (int i = 0; i < 8; ++i) f8[i] = 0.f;
                        } else if (p.residual_bf16) {   // skip connection is added AFTER the activation (zeros past N / past M)
                            const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rcur[g8]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 r2 = __bfloat1622float2(rp[i]);
                                f8[2 * i] += r2.x;
                                f8[2 * i + 1] += r2.y;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const __nv_bfloat162 bb = __floats2bfloat162_rn(f8[2 * i], f8[2 * i + 1]);
                            pk[g8 * 4 + i] = *reinterpret_cast<const uint32_t *>(&bb);
                        }
                    }
                    if (!(p.dbg & 1)) {
                        stage_chunk_bf16(stg, lane, pk);
                        if (lane == 0) {
                            if (p.a_mode == A_FLAT || p.a_mode == A_SCALED) tma_store_2d(&tmC, stg, p.col_off + col0, m_t * BM + q * 32);
                            else tma_store_4d(&tmC, stg, p.col_off + col0, st_x, st_y, st_img);
                        }
                    } else if (pk[0] == 0x12345678u) static_cast<uint32_t *>(p.out)[0] = pk[1];   // keep the math alive
                    continue;
                }
                if (kEpi != EPI_CONV && !valid) continue;
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
                        const float4 sc = *reinterpret_cast<const float4 *>(&g_scale[c * 32 + 4 * i]);
                        const float4 sh = *reinterpret_cast<const float4 *>(&g_shift[c * 32 + 4 * i]);
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
                } else if (kEpi == EPI_CONV) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __nv_bfloat162 bb = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                        pk[i] = *reinterpret_cast<const uint32_t *>(&bb);
                    }
                    stage_chunk_bf16(stg, lane, pk);
                    if (lane == 0) {
                        if (p.a_mode == A_FLAT || p.a_mode == A_SCALED) tma_store_2d(&tmC, stg, p.col_off + col0, m_t * BM + q * 32);
                        else tma_store_4d(&tmC, stg, p.col_off + col0, st_x, st_y, st_img);
                    }
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
            if (q == 2 && lane == 0) EWVIT_TRACE(2 + grp, it, 2);
            ewvit::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kPair) ewvit::mbar_arrive_leader(ewvit::smem_u32(&tempty[acc]));   // the leader's issuer waits for both CTAs' epilogues
                else ewvit::mbar_arrive(ewvit::smem_u32(&tempty[acc]));
            }
            if (q == 2 && lane == 0) EWVIT_TRACE(2 + grp, it, 3);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging buffers are read out before the CTA retires
    }

    ewvit::tc_fence_before();
    if (kPair) ewvit::cluster_sync_all();     // the leader's MMAs read the peer's shared memory and signal its barriers until the very end
    else __syncthreads();
    if (warp == 1) {
        if (kPair) ewvit::tmem_dealloc_pair(tmem_base, kTmemColsT);
        else ewvit::tmem_dealloc(tmem_base, kTmemColsT);
    }
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

// Rows of the B (weight) box: a single column tile only needs the 16-aligned number of output channels -- the MMA reads
// n_valid rows -- which shrinks the resident filter bank / ring slots and so deepens the activation ring.
static inline int b_box_rows(int N, int bn, int tiles_n) {
    if (tiles_n != 1) return bn;
    const int r = (N + 15) & ~15;
    return r < bn ? r : bn;
}

static long long *g_trace = nullptr;
static int g_dbg = 0;

template <int kEpi, bool kBuilder, int kBN, int kFast = 0, bool kPair = false>
int launch_gemm_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, GemmParams p, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    EWVIT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EWVIT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<kEpi, kBuilder, kBN, kFast, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<kEpi, kBuilder, kBN>::kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (p.b_tile_bytes <= 0) p.b_tile_bytes = kBN * BK * 2;
    if (p.stage_bytes <= 0) p.stage_bytes = kTileBytes + p.b_tile_bytes;
    if (p.stages <= 0) {
        p.stages = Cfg<kEpi, kBuilder, kBN>::kOperandBytes / p.stage_bytes;
        if (p.stages > 6) p.stages = 6;
    }
    // plain (one-tap-per-slot) paths with a single column tile and a small weight matrix: keep all of B resident and let
    // the ring carry activation tiles only -- TMA's ~1.5 us latency makes throughput = bytes in flight / latency, and a
    // slot without its B tile is half the size
    if (!kBuilder && !p.flat3 && !p.b_res && p.tiles_n == 1 && p.splits == 1 && !(g_dbg & 64) &&
        p.num_kb * p.b_tile_bytes <= 96 * 1024 && p.stage_bytes == kTileBytes + p.b_tile_bytes) {
        p.b_res = 1;
        p.stage_bytes = kTileBytes;
        p.stages = (Cfg<kEpi, kBuilder, kBN>::kOperandBytes - p.num_kb * p.b_tile_bytes) / kTileBytes;
        if (p.stages > kStages) p.stages = kStages;
        p.bres_off = p.stages * kTileBytes;
    }
    p.trace = g_trace;
    p.dbg = g_dbg;
    long long work = (long long)p.tiles_m * p.tiles_n * p.splits;
    long long grid = ewvit_num_sms();
    if (kPair) {
        // one 2-CTA cluster (= one TPC) per pair of row tiles, persistent: #clusters = min(#SMs / 2, #pair items)
        const long long pairs = (long long)((p.tiles_m + 1) / 2) * p.tiles_n;
        long long clusters = grid / 2;
        if (clusters > pairs) clusters = pairs;
        if (clusters <= 0) return EWVIT_OK;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * clusters));
        cfg.blockDim = dim3(Cfg<kEpi, kBuilder, kBN>::kThreads);
        cfg.dynamicSmemBytes = Cfg<kEpi, kBuilder, kBN>::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        EWVIT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<kEpi, kBuilder, kBN, kFast, kPair>, tmA, tmB, tmC, p));
        ewvit_count_launch();
        return EWVIT_OK;
    }
    if (grid > work) grid = work;
    if (grid <= 0) return EWVIT_OK;
    gemm_tc_kernel<kEpi, kBuilder, kBN, kFast><<<(unsigned)grid, Cfg<kEpi, kBuilder, kBN>::kThreads, Cfg<kEpi, kBuilder, kBN>::kSmemBytes, stream>>>(tmA, tmB, tmC, p);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}

int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const GemmParams &p, int epi, cudaStream_t stream,
                int bn = BN) {
    if (epi == EPI_CONV && p.pair) return launch_gemm_t<EPI_CONV, false, 128, 0, true>(tmA, tmB, tmC, p, stream);
    if (epi == EPI_CONV && bn == 192) return launch_gemm_t<EPI_CONV, false, 192>(tmA, tmB, tmC, p, stream);   // three-level MWT head conv
    if (epi == EPI_CONV) return launch_gemm_t<EPI_CONV, false, 128>(tmA, tmB, tmC, p, stream);
    if (epi == EPI_BB) {
        const int fast = (g_dbg & 256) ? 0 : (p.act == 4 && !p.residual_bf16) ? 1 : p.act == 0 ? 2 : 0;
        const bool builder = p.a_mode == A_IM2COL || p.a_mode == A_SCALED;
#define EWVIT_BB(B_, N_, P_)                                                                          \
        (fast == 1 ? launch_gemm_t<EPI_BB, B_, N_, 1, P_>(tmA, tmB, tmC, p, stream)                       \
         : fast == 2 ? launch_gemm_t<EPI_BB, B_, N_, 2, P_>(tmA, tmB, tmC, p, stream)                     \
                     : launch_gemm_t<EPI_BB, B_, N_, 0, P_>(tmA, tmB, tmC, p, stream))
        if (p.pair) {        // 1x1 convs on CTA pairs (A_FLAT plain, A_SCALED gated)
            if (builder) return bn == 256 ? EWVIT_BB(true, 256, true) : EWVIT_BB(true, 128, true);
            return bn == 256 ? EWVIT_BB(false, 256, true) : EWVIT_BB(false, 128, true);
        }
        if (builder) return bn == 256 ? EWVIT_BB(true, 256, false) : EWVIT_BB(true, 128, false);
        return bn == 256 ? EWVIT_BB(false, 256, false) : EWVIT_BB(false, 128, false);
#undef EWVIT_BB
    }
    if (epi == EPI_PARTIAL) return launch_gemm_t<EPI_PARTIAL, false, 128>(tmA, tmB, tmC, p, stream);
    return launch_gemm_t<EPI_LINEAR, false, 128>(tmA, tmB, tmC, p, stream);
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
                         const uint64_t *strides_bytes, const uint32_t *box, const uint32_t *estr, bool swizzle128, bool swizzle32, bool swizzle64) {
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
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EWVIT_REQUIRE(r == CUDA_SUCCESS, EWVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return EWVIT_OK;
}

// Output tensor map of the TMA-store epilogues: bf16, 64-byte swizzle, box = 32 channels x 32 rows (flat) or
// 32 channels x 16 pixels x 2 pixel rows (tiled NHWC, optionally with a one-pixel border).
static int make_out_tmap(CUtensorMap *out, void *base, bool flat, long long rows, int ldc, int n, int h_total, int w_total,
                         int h_extent = 0, int w_extent = 0) {
    ewvit_encode_tiled_fn enc = ewvit_get_encode_tiled();
    EWVIT_REQUIRE(enc != nullptr, EWVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t bdim[4], es[4] = {1, 1, 1, 1};
    int rank;
    if (flat) {
        rank = 2;
        gdim[0] = (cuuint64_t)ldc; gdim[1] = (cuuint64_t)rows;
        gstr[0] = (cuuint64_t)ldc * 2;
        bdim[0] = 32; bdim[1] = 32;
    } else {
        rank = 4;
        // the extents may stop short of the pitch (a trailing border column/row that must never be written)
        gdim[0] = (cuuint64_t)ldc; gdim[1] = (cuuint64_t)(w_extent ? w_extent : w_total);
        gdim[2] = (cuuint64_t)(h_extent ? h_extent : h_total); gdim[3] = (cuuint64_t)n;
        gstr[0] = (cuuint64_t)ldc * 2; gstr[1] = (cuuint64_t)w_total * ldc * 2; gstr[2] = (cuuint64_t)h_total * w_total * ldc * 2;
        bdim[0] = 32; bdim[1] = 16; bdim[2] = 2; bdim[3] = 1;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gdim, gstr, bdim, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EWVIT_REQUIRE(r == CUDA_SUCCESS, EWVIT_ERR_CUDA, "cuTensorMapEncodeTiled (output map) failed with CUresult %d", (int)r);
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
    rc = launch_gemm(tmA, tmB, tmA /*no TMA store in these epilogues*/, p, p.splits > 1 ? EPI_PARTIAL : EPI_LINEAR, (cudaStream_t)stream);
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

extern "C" int ewvit_conv3x3_bf16(const void *x_base, int x_ldc, int x_coff, const void *w, int n, int h, int wd, int cin, int cout,
                                  int stride, int in_padded, const float *scale, const float *shift, int relu, void *y,
                                  int y_ldc, int y_coff, int out_padded, int force_tiled, void *stream) {
    EWVIT_REQUIRE(x_base && x_ldc >= cin && x_coff >= 0 && x_coff + cin <= x_ldc && x_ldc % 8 == 0 && x_coff % 8 == 0, EWVIT_ERR_INVALID_ARG,
                  "ewvit_conv3x3_bf16: bad input channel pitch/offset (x_ldc=%d x_coff=%d cin=%d)", x_ldc, x_coff, cin);
    EWVIT_REQUIRE(x_ldc == cin || (stride == 1 && in_padded && out_padded && !force_tiled), EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv3x3_bf16: a channel slice of a wider tensor is implemented for the stride-1 padded-flat path only");
    const void *x = static_cast<const __nv_bfloat16 *>(x_base) + x_coff;
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
        uint64_t dims[2] = {(uint64_t)cin, (uint64_t)rows}, str[2] = {2, (uint64_t)x_ldc * 2};
        uint32_t box[2] = {BK, BM};
        rc = ewvit_make_tmap_bf16(&tmA, x, 2, dims, str, box, nullptr);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_FLAT;
        p.M = rows;
        p.tiles_m = (int)((rows + BM - 1) / BM);
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) p.tap_a0[dy * 3 + dx] = (dy - 1) * win + (dx - 1);
        if (!(g_dbg & 32)) {      // row-shared taps (debug flag 32 = the one-tile-per-tap path)
            box[1] = kFlat3Rows;
            rc = ewvit_make_tmap_bf16(&tmA, x, 2, dims, str, box, nullptr);
            if (rc != EWVIT_OK) return rc;
            p.flat3 = 1;
            p.num_kb = 3 * chunks;
            p.kb_per_split = p.num_kb;
            // CTA pairs (cta_group::2) whenever there are at least two row tiles: each CTA stores half of every weight tile
            p.pair = (!(g_dbg & 128) && p.tiles_m >= 2) ? 1 : 0;
            const int bn_rows = p.pair ? BN / 2 : BN;
            if (p.pair) {
                uint64_t dimsb[2] = {(uint64_t)9 * cin, (uint64_t)cout}, strb[2] = {2, (uint64_t)9 * cin * 2};
                uint32_t boxh[2] = {BK, (uint32_t)bn_rows};
                rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxh, nullptr);
                if (rc != EWVIT_OK) return rc;
            }
            p.b_half = bn_rows;
            p.b_tile_bytes = bn_rows * BK * 2;
            p.stage_bytes = kFlat3ABytes + 3 * p.b_tile_bytes;
            p.stages = Cfg<EPI_CONV, false>::kOperandBytes / p.stage_bytes;
            if (p.stages > kStages) p.stages = kStages;
            // small filter banks (the 64 -> 128 fusion conv: 144 KB, 72 KB per CTA of a pair) stay resident; the ring then carries
            // activation windows only
            const int wbytes = 9 * cin * bn_rows * 2;
            if (p.tiles_n == 1 && !(g_dbg & 64) && wbytes + 3 * kFlat3ABytes <= Cfg<EPI_CONV, false>::kOperandBytes) {
                p.b_res = 1;
                p.stage_bytes = kFlat3ABytes;
                p.stages = (Cfg<EPI_CONV, false>::kOperandBytes - wbytes) / kFlat3ABytes;
                if (p.stages > kStages) p.stages = kStages;
                p.bres_off = p.stages * kFlat3ABytes;
            }
        }
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
        if (!(g_dbg & 128) && p.tiles_m >= 2 && p.tiles_n == 1) {
            // CTA pairs here, too (freq_conv / freq_pool, stride 2): half of every weight tile per CTA, ring slot 32 -> 24 KB
            // (7 stages instead of 5 in the 184 KB operand region) and 2 KB less shared-memory traffic per MMA
            p.pair = 1;
            p.b_half = BN / 2;
            uint64_t dimsb[2] = {(uint64_t)9 * cin, (uint64_t)cout}, strb[2] = {2, (uint64_t)9 * cin * 2};
            uint32_t boxh[2] = {BK, (uint32_t)p.b_half};
            rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxh, nullptr);
            if (rc != EWVIT_OK) return rc;
            p.b_tile_bytes = p.b_half * BK * 2;
            p.stage_bytes = kTileBytes + p.b_tile_bytes;
            p.stages = Cfg<EPI_CONV, false>::kOperandBytes / p.stage_bytes;
            if (p.stages > 8) p.stages = 8;
        }
        const int off = in_padded ? 0 : -1;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                p.tap_a0[dy * 3 + dx] = dx + off;
                p.tap_a1[dy * 3 + dx] = dy + off;
            }
    }
    CUtensorMap tmC;
    if (flat)
        rc = make_out_tmap(&tmC, y, true, (long long)n * hin * win, y_ldc, 0, 0, 0);
    else
        rc = make_out_tmap(&tmC, y, false, 0, y_ldc, n, ho + 2 * p.out_pad, wo + 2 * p.out_pad, ho + p.out_pad, wo + p.out_pad);
    if (rc != EWVIT_OK) return rc;
    return launch_gemm(tmA, tmB, tmC, p, EPI_CONV, (cudaStream_t)stream);
}


// General NHWC bf16 convolution for the EfficientNet backbone (1x1 or 3x3/pad 1, stride 1 or 2) with a fused
// bias + optional bf16 residual + activation epilogue.  Channel counts only need to be multiples of 8: the K and N
// tails are zero-filled by TMA (boxes may overhang the tensor) and masked in the epilogue.
static int conv_nhwc_impl(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int ksize, int stride,
                          const float *bias, int act, const void *residual, void *y, int in_padded, int out_padded, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_conv_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && w && y, EWVIT_ERR_INVALID_ARG, "ewvit_conv_nhwc_bf16: NULL pointer");
    EWVIT_REQUIRE((ksize == 1 && stride == 1) || (ksize == 3 && (stride == 1 || stride == 2)), EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv_nhwc_bf16: supports 1x1/stride 1 and 3x3/stride 1|2 (got k=%d s=%d)", ksize, stride);
    EWVIT_REQUIRE(cin % 8 == 0 && cout % 8 == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv_nhwc_bf16: channel counts must be multiples of 8 (got cin=%d cout=%d)", cin, cout);
    EWVIT_REQUIRE(act == 0 || act == 1 || act == 3 || act == 4, EWVIT_ERR_INVALID_ARG, "ewvit_conv_nhwc_bf16: act must be 0 (none), 1 (relu), 3 (silu) or 4 (silu, halved operands)");
    EWVIT_REQUIRE(ewvit_aligned16(x) && ewvit_aligned16(w) && ewvit_aligned16(y) && ewvit_aligned16(residual), EWVIT_ERR_INVALID_ARG,
                  "ewvit_conv_nhwc_bf16: pointers must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;

    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    GemmParams p = {};
    // column tile: 256 wide when cout > 128, so A is fetched (or assembled) once per 256 output channels
    // (the stride-2 halo of the assembled path is 4x larger: keep 128-wide tiles there so two halo buffers still fit)
    const int bn = (cout > 128 && !(ksize == 3 && cin < BK && stride == 2)) ? 256 : 128;
    p.N = cout;
    p.tiles_n = (cout + bn - 1) / bn;
    const int brows = b_box_rows(cout, bn, p.tiles_n);
    p.b_tile_bytes = brows * BK * 2;
    const int stage_bytes = kTileBytes + p.b_tile_bytes;
    p.splits = 1;
    p.out = y; p.out_fp32 = 0; p.ldo = cout; p.col_off = 0;
    p.shift = bias; p.act = act;
    p.residual_bf16 = static_cast<const __nv_bfloat16 *>(residual);
    p.ldr = cout;
    CUtensorMap tmA, tmB;
    const bool ow = ksize == 3 && stride == 1 && cin < BK && in_padded && out_padded;
    EWVIT_REQUIRE(in_padded == out_padded || (ksize == 3 && cin < BK), EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv_nhwc_bf16_ex: only the small-channel 3x3 convs may change the padding of the layout");
    EWVIT_REQUIRE(!(in_padded || out_padded) || ksize == 1 || cin < BK, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv_nhwc_bf16_ex: padded layouts are implemented for 1x1 convs and 3x3 convs with cin < 64");
    long long flat_rows = 0;
    if (ow) {
        // "overlapping window" 3x3 conv on padded-flat tensors: with cin < 64 the three horizontal taps of an output pixel
        // are 3*cin CONTIGUOUS elements, so the im2col row of a vertical tap is a window of a tensor map whose rows are
        // 3*cin long but only cin apart; k-blocks = (dy, 64-element piece of the window), columns past 3*cin zero-filled.
        // No builder warps: this is the plain TMA -> MMA pipeline with the four-group epilogue.
        const int hin = h + 2, win = wd + 2;
        const long long rows = (long long)n * hin * win;
        const int nsub = (3 * cin + BK - 1) / BK;
        uint64_t dims[2] = {(uint64_t)3 * cin, (uint64_t)(rows - 2)}, str[2] = {2, (uint64_t)cin * 2};
        uint32_t box[2] = {BK, BM};
        rc = ewvit_make_tmap_bf16(&tmA, x, 2, dims, str, box, nullptr);
        if (rc != EWVIT_OK) return rc;
        uint64_t dimsb[2] = {(uint64_t)3 * nsub * BK, (uint64_t)cout}, strb[2] = {2, (uint64_t)3 * nsub * BK * 2};
        uint32_t boxb[2] = {BK, (uint32_t)brows};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_FLAT;
        p.chunks_per_tap = nsub;
        p.num_kb = 3 * nsub;
        p.kb_per_split = p.num_kb;
        for (int dy = 0; dy < 3; ++dy) p.tap_a0[dy] = (dy - 1) * win - 1;
        p.M = rows;
        p.tiles_m = (int)((rows + BM - 1) / BM);
        p.pad_hp = hin;
        p.pad_wp = win;
        flat_rows = rows;
    } else if (ksize == 1) {
        const long long rows = in_padded ? (long long)n * (h + 2) * (wd + 2) : (long long)n * h * wd;
        if (in_padded) { p.pad_hp = h + 2; p.pad_wp = wd + 2; }
        flat_rows = rows;
        const int num_kb = (cin + BK - 1) / BK;
        uint64_t dims[2] = {(uint64_t)cin, (uint64_t)rows}, str[2] = {2, (uint64_t)cin * 2};
        uint32_t box[2] = {BK, BM};
        rc = ewvit_make_tmap_bf16(&tmA, x, 2, dims, str, box, nullptr);
        if (rc != EWVIT_OK) return rc;
        uint64_t dimsb[2] = {(uint64_t)cin, (uint64_t)cout};
        uint32_t boxb[2] = {BK, (uint32_t)brows};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, str, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
        p.a_mode = A_FLAT;
        p.chunks_per_tap = num_kb;
        p.num_kb = num_kb;
        p.kb_per_split = num_kb;
        p.M = rows;
        p.tiles_m = (int)((rows + BM - 1) / BM);
        if (!(g_dbg & 128) && p.tiles_m >= 2) {
            // CTA pairs: each CTA stores half of every weight tile (half of the MMA's N: kBN / 2 rows with several column tiles,
            // else half of the 16-aligned channel count) -- a ring slot shrinks by a third and so do the L2 -> shared bytes
            p.pair = 1;
            p.b_half = (p.tiles_n > 1 ? bn : brows) / 2;
            uint32_t boxh[2] = {BK, (uint32_t)p.b_half};
            rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, str, boxh, nullptr);
            if (rc != EWVIT_OK) return rc;
            p.b_tile_bytes = p.b_half * BK * 2;
        }
    } else {
        EWVIT_REQUIRE(cin <= BK || cin % BK == 0, EWVIT_ERR_UNSUPPORTED,
                      "ewvit_conv_nhwc_bf16: 3x3 needs cin <= 64 or cin %% 64 == 0 (got %d)", cin);
        const int kdense = 9 * cin;
        const int num_kb = (kdense + BK - 1) / BK;
        uint64_t dimsb[2] = {(uint64_t)num_kb * BK, (uint64_t)cout}, strb[2] = {2, (uint64_t)num_kb * BK * 2};
        uint32_t boxb[2] = {BK, (uint32_t)brows};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxb, nullptr);
        if (rc != EWVIT_OK) return rc;
        const int box_w = 16, box_h = 8;
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
        uint64_t str[4] = {2, (uint64_t)cin * 2, (uint64_t)wd * cin * 2, (uint64_t)h * wd * cin * 2};
        p.box_w = box_w; p.box_h = box_h; p.in_stride = stride;
        p.tiles_x = (wo + box_w - 1) / box_w;
        p.tiles_y = (ho + box_h - 1) / box_h;
        p.tiles_m = p.tiles_x * p.tiles_y * n;
        p.out_w = wo; p.out_h = ho;
        p.out_pad = out_padded ? 1 : 0;
        p.out_wp = wo + 2 * p.out_pad;
        p.out_img_rows = (long long)(ho + 2 * p.out_pad) * p.out_wp;
        p.M = (long long)n * p.out_img_rows;
        p.num_kb = num_kb;
        p.kb_per_split = num_kb;
        if (cin < BK) {
            // halo fetched once per tile, operand rows assembled in shared memory (dense K = 9*cin)
            p.a_mode = A_IM2COL;
            p.cin = cin;
            p.halo_w = (box_w - 1) * stride + 3;
            p.halo_h = (box_h - 1) * stride + 3;
            p.halo_nb = (p.halo_w * cin + 255) / 256;                   // TMA boxes are at most 256 elements wide
            p.halo_ppb = (p.halo_w + p.halo_nb - 1) / p.halo_nb;
            const int plane_payload = p.halo_h * p.halo_ppb * cin * 2;
            p.plane_bytes = (plane_payload + 127) & ~127;                // TMA destinations must be 128-byte aligned
            p.halo_bytes = p.halo_nb * plane_payload;                   // bytes the TMA unit reports on the mbarrier
            p.halo_stride = (p.halo_nb * p.plane_bytes + 1023) & ~1023;
            p.b_res = (p.tiles_n == 1 && num_kb * p.b_tile_bytes <= 64 * 1024) ? 1 : 0;
            int sbytes = stage_bytes, reserve = 0;
            if (p.b_res) {                      // ring slots hold A only; the weights get their own region
                sbytes = kTileBytes;
                reserve = num_kb * p.b_tile_bytes;
            }
            p.stage_bytes = sbytes;
            p.stages = bn == 256 ? 3 : 4;
            while (p.stages > 2 && (Cfg<EPI_BB, true>::kOperandBytes - reserve - p.stages * sbytes) / p.halo_stride < 2) p.stages -= 1;
            p.halo_bufs = (Cfg<EPI_BB, true>::kOperandBytes - reserve - p.stages * sbytes) / p.halo_stride;
            if (p.halo_bufs > kMaxHalo) p.halo_bufs = kMaxHalo;
            EWVIT_REQUIRE(p.halo_bufs >= 2, EWVIT_ERR_UNSUPPORTED, "ewvit_conv_nhwc_bf16: halo too large");
            p.bres_off = p.stages * sbytes + p.halo_bufs * p.halo_stride;
            p.chunks_per_tap = 1;
            // activation viewed as [n][h][wd*cin]: a run of pixels of one image row is one contiguous TMA row
            // (a padded-flat input is addressed through its interior: same extents, the pitches of the padded tensor)
            const int hpi = in_padded ? h + 2 : h, wpi = in_padded ? wd + 2 : wd;
            const void *xin = in_padded ? static_cast<const void *>(static_cast<const __nv_bfloat16 *>(x) + ((size_t)wpi + 1) * cin) : x;
            EWVIT_REQUIRE(ewvit_aligned16(xin), EWVIT_ERR_UNSUPPORTED, "ewvit_conv_nhwc_bf16_ex: interior of the padded input is not 16-byte aligned");
            uint64_t dims3[3] = {(uint64_t)wd * cin, (uint64_t)h, (uint64_t)n};
            uint64_t str3[3] = {2, (uint64_t)wpi * cin * 2, (uint64_t)hpi * wpi * cin * 2};
            uint32_t box3[3] = {(uint32_t)(p.halo_ppb * cin), (uint32_t)p.halo_h, 1};
            rc = ewvit_make_tmap_bf16(&tmA, xin, 3, dims3, str3, box3, nullptr, /*swizzle128=*/false);
            if (rc != EWVIT_OK) return rc;
        } else {
            p.a_mode = A_TILE4D;
            p.chunks_per_tap = cin / BK;
            uint32_t box[4] = {BK, (uint32_t)(box_w * stride), (uint32_t)(box_h * stride), 1};
            uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
            rc = ewvit_make_tmap_bf16(&tmA, x, 4, dims, str, box, es);
            if (rc != EWVIT_OK) return rc;
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx) {
                    p.tap_a0[dy * 3 + dx] = dx - 1;
                    p.tap_a1[dy * 3 + dx] = dy - 1;
                }
        }
    }
    CUtensorMap tmC;
    if (flat_rows)
        rc = make_out_tmap(&tmC, y, true, flat_rows, cout, 0, 0, 0);
    else
        rc = make_out_tmap(&tmC, y, false, 0, cout, n, ho + 2 * p.out_pad, wo + 2 * p.out_pad, ho + p.out_pad, wo + p.out_pad);
    if (rc != EWVIT_OK) return rc;
    return launch_gemm(tmA, tmB, tmC, p, EPI_BB, (cudaStream_t)stream, bn);
}

extern "C" int ewvit_conv_nhwc_bf16(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int ksize,
                                    int stride, const float *bias, int act, const void *residual, void *y, void *stream) {
    return conv_nhwc_impl(x, w, n, h, wd, cin, cout, ksize, stride, bias, act, residual, y, 0, 0, stream);
}

extern "C" int ewvit_conv_nhwc_bf16_ex(const void *x, const void *w, int n, int h, int wd, int cin, int cout, int ksize,
                                       int stride, const float *bias, int act, const void *residual, void *y, int in_padded,
                                       int out_padded, void *stream) {
    return conv_nhwc_impl(x, w, n, h, wd, cin, cout, ksize, stride, bias, act, residual, y, in_padded ? 1 : 0, out_padded ? 1 : 0, stream);
}

// Tensor-core head of the wavelet levels, step 2 (mwt.py:84-86): the three per-colour Conv2d(3->18, 3x3, p1)+BN+ReLU as ONE
// block-diagonal 9 -> 54(64) conv per level.
// All three wavelet levels in ONE launch.  `up` is the padded-flat [n, h+2, wd+2, 32] bf16 tensor of ewvit_mwt_upsample3_fwd: a pixel
// is one 64-byte row = channels [9 l, 9 l + 9) for level l (+ 5 channels of zero padding) = two MMA K steps.  Per 128-pixel tile and
// vertical tap ONE 64B-swizzled window of 136 pixel rows is fetched and 6 MMAs (3 horizontal taps x 2 K steps, N = 128) read it at
// start addresses shifted by dx rows; level l accumulates in TMEM columns [64 l, 64 l + 64).  The single-level kernel fetched 408
// window rows and issued nine MMAs per tile and LEVEL and was bound by those fixed costs (~1700 cycles per tile); here the same rows and
// eighteen MMAs serve the three levels (versions with 128-byte pixel rows or with one N = 64 MMA per level and K step were slower:
// 2.7 GB of L2 -> shared window traffic / ~70 cycles per MMA whatever its N), and the 128 x 192 epilogue runs on four groups of warps.
//   w [128, 288] bf16, w[r][((dy*3 + dx)*2 + step)*16 + kk]: rows [0, 64) of step 0 = level 0, rows [64, 128) of step 0 and rows
//     [0, 64) of step 1 = level 1, rows [64, 128) of step 1 = level 2; within a block row g*18+oc holds seperate[g].weight[oc][ic][dy][dx]
//     where channel 16*step + kk of a pixel is subband 3g+ic of that level, zero elsewhere (engine.pack_head3_weights);
//   scale/shift [192] fp32 = the 64-entry folded BN repeated per level;
//   y [n, h+2, wd+2, 192] bf16 padded-flat: channels [64 l, 64 l + 54) = head of level l, the rest and the border zero.
extern "C" int ewvit_mwt_head_conv3_fwd(const void *up, const void *w, int n, int h, int wd, const float *scale, const float *shift,
                                        void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && h > 0 && wd > 0, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_conv3_fwd: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(up && w && scale && shift && y, EWVIT_ERR_INVALID_ARG, "ewvit_mwt_head_conv3_fwd: NULL pointer");
    EWVIT_REQUIRE(ewvit_aligned16(up) && ewvit_aligned16(w) && ewvit_aligned16(y), EWVIT_ERR_INVALID_ARG,
                  "ewvit_mwt_head_conv3_fwd: pointers must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int hin = h + 2, win = wd + 2, cpx = 32, cout = 64, levels = 3, tiles_per_tap = 2, brows = 128;
    const long long rows = (long long)n * hin * win;
    GemmParams p = {};
    p.a_mode = A_FLAT;
    p.flat3 = 1;
    p.k16 = 2;                    // three levels per pixel row
    p.chunks_per_tap = 1;
    p.num_kb = 3;
    p.kb_per_split = 3;
    p.splits = 1;
    p.N = levels * cout;
    p.tiles_n = 1;
    p.M = rows;
    p.tiles_m = (int)((rows + BM - 1) / BM);
    for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) p.tap_a0[dy * 3 + dx] = (dy - 1) * win + (dx - 1);
    p.pad_hp = hin;
    p.pad_wp = win;
    p.out = y; p.out_fp32 = 0; p.ldo = levels * cout; p.col_off = 0;
    p.scale = scale; p.shift = shift; p.act = 1;
    CUtensorMap tmA, tmB, tmC;
    {
        uint64_t dims[2] = {(uint64_t)cpx, (uint64_t)rows}, str[2] = {2, (uint64_t)cpx * 2};
        uint32_t box[2] = {(uint32_t)cpx, (uint32_t)kFlat3Rows};
        rc = ewvit_make_tmap_bf16(&tmA, up, 2, dims, str, box, nullptr, false, false, /*swizzle64=*/true);
        if (rc != EWVIT_OK) return rc;
        const uint64_t kw = (uint64_t)9 * tiles_per_tap * 16;
        uint64_t dimsb[2] = {kw, (uint64_t)brows}, strb[2] = {2, kw * 2};
        uint32_t boxb[2] = {16u, (uint32_t)brows};
        rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, strb, boxb, nullptr, false, /*swizzle32=*/true);
        if (rc != EWVIT_OK) return rc;
    }
    p.b_tile_bytes = brows * 16 * 2;                     // 4 KB per (tap, K step)
    p.stage_bytes = kFlat3Rows * cpx * 2;                // 8704 bytes = 17 swizzle atoms of 512 bytes
    p.b_res = 1;                                         // all 18 weight tiles (72 KB) stay resident
    p.stages = (Cfg<EPI_CONV, false, 192>::kOperandBytes - 9 * tiles_per_tap * p.b_tile_bytes) / p.stage_bytes;
    if (p.stages > kStages) p.stages = kStages;
    p.bres_off = (p.stages * p.stage_bytes + 1023) & ~1023;
    rc = make_out_tmap(&tmC, y, true, rows, levels * cout, 0, 0, 0);
    if (rc != EWVIT_OK) return rc;
    return launch_gemm(tmA, tmB, tmC, p, EPI_CONV, (cudaStream_t)stream, 192);
}

// Debug aid: when non-NULL, CTA 0 of every subsequent GEMM/conv launch writes clock64 stamps of its warp roles
// to this device buffer ([6 roles][64 tiles][4] int64).  Not part of the hot path; pass NULL to switch it off.
extern "C" int ewvit_debug_set_flags(int flags) {
    g_dbg = flags;
    return EWVIT_OK;
}

extern "C" int ewvit_debug_set_trace(void *device_buffer) {
    g_trace = static_cast<long long *>(device_buffer);
    return EWVIT_OK;
}


// 1x1 convolution behind a squeeze-excitation block: y = act(((x * gate[frame]) W^T) + bias) + residual, with the gate
// applied while the A operand tile is assembled (no separate scaling pass over the expanded tensor).
extern "C" int ewvit_conv1x1_gated_nhwc_bf16(const void *x, const void *gate, int gate_bf16, const void *w, int n, int hw, int cin,
                                             int cout, const float *bias, int act, const void *residual, void *y, void *stream) {
    EWVIT_REQUIRE(n >= 0 && hw > 0 && cin > 0 && cout > 0, EWVIT_ERR_INVALID_ARG, "ewvit_conv1x1_gated_nhwc_bf16: bad sizes");
    if (n == 0) return EWVIT_OK;
    EWVIT_REQUIRE(x && gate && w && y, EWVIT_ERR_INVALID_ARG, "ewvit_conv1x1_gated_nhwc_bf16: NULL pointer");
    EWVIT_REQUIRE(cin % 8 == 0 && cout % 8 == 0, EWVIT_ERR_UNSUPPORTED,
                  "ewvit_conv1x1_gated_nhwc_bf16: channel counts must be multiples of 8 (got cin=%d cout=%d)", cin, cout);
    EWVIT_REQUIRE(act == 0 || act == 1 || act == 3 || act == 4, EWVIT_ERR_INVALID_ARG, "ewvit_conv1x1_gated_nhwc_bf16: act must be 0, 1, 3 or 4");
    EWVIT_REQUIRE(ewvit_aligned16(x) && ewvit_aligned16(gate) && ewvit_aligned16(w) && ewvit_aligned16(y) && ewvit_aligned16(residual),
                  EWVIT_ERR_INVALID_ARG, "ewvit_conv1x1_gated_nhwc_bf16: pointers must be 16-byte aligned");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    const int bn = cout > 128 ? 256 : 128;
    const long long rows = (long long)n * hw;
    const int num_kb = (cin + BK - 1) / BK;
    GemmParams p = {};
    p.a_mode = A_SCALED;
    p.a_gate = gate;
    p.a_gate_bf16 = gate_bf16 ? 1 : 0;
    p.a_hw = hw;
    p.a_k = cin;
    // a 128-row tile touches at most (127 + hw - 1) / hw + 1 frames
    p.a_gate_smem = (gate_bf16 && cin <= kGateMaxK && (BM - 1 + hw - 1) / hw + 1 <= kGateFrames) ? 1 : 0;
    p.N = cout;
    p.M = rows;
    p.tiles_m = (int)((rows + BM - 1) / BM);
    p.tiles_n = (cout + bn - 1) / bn;
    p.splits = 1;
    p.chunks_per_tap = num_kb;
    p.num_kb = num_kb;
    p.kb_per_split = num_kb;
    p.out = y; p.out_fp32 = 0; p.ldo = cout; p.col_off = 0;
    p.shift = bias; p.act = act;
    p.residual_bf16 = static_cast<const __nv_bfloat16 *>(residual);
    p.ldr = cout;
    int brows = b_box_rows(cout, bn, p.tiles_n);
    if (!(g_dbg & 128) && p.tiles_m >= 2) {      // CTA pairs: half of every weight tile per CTA
        p.pair = 1;
        p.b_half = (p.tiles_n > 1 ? bn : brows) / 2;
        brows = p.b_half;
    }
    p.b_tile_bytes = brows * BK * 2;
    p.stage_bytes = kTileBytes + p.b_tile_bytes;
    p.stages = Cfg<EPI_BB, true>::kOperandBytes / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    CUtensorMap tmA, tmB, tmC;
    uint64_t dimsa[2] = {(uint64_t)cin, (uint64_t)rows}, dimsb[2] = {(uint64_t)cin, (uint64_t)cout}, str[2] = {2, (uint64_t)cin * 2};
    uint32_t boxa[2] = {BK, BM}, boxb[2] = {BK, (uint32_t)brows};
    rc = ewvit_make_tmap_bf16(&tmA, x, 2, dimsa, str, boxa, nullptr);
    if (rc != EWVIT_OK) return rc;
    rc = ewvit_make_tmap_bf16(&tmB, w, 2, dimsb, str, boxb, nullptr);
    if (rc != EWVIT_OK) return rc;
    rc = make_out_tmap(&tmC, y, true, rows, cout, 0, 0, 0);
    if (rc != EWVIT_OK) return rc;
    return launch_gemm(tmA, tmB, tmC, p, EPI_BB, (cudaStream_t)stream, bn);
}
