// First VBM block without its 16-channel full-resolution tensors.
//
// conv1 -> BatchNorm -> ReLU -> MaxPool(2) at 8 x 121x145x121 produces a 543 MB bf16 conv output that the unfused
// pipeline writes once and reads three times (pool forward, BN backward, weight gradient) and a second 543 MB
// gradient tensor that is written and read once: ~2.7 GB of HBM traffic and four bandwidth/instruction bound
// kernels (~0.95 ms per step).  The banded-Toeplitz formulation (the W axis of the volume is the GEMM K axis) makes the convolution so
// cheap on the tensor cores (45 us of tcgen05 time per pass) that it is faster to RECOMPUTE it than to store it:
//
//   MODE_STATS  conv -> per-channel sum / sum of squares of the fp32 accumulators            (nothing else written)
//   MODE_POOL   conv -> BN -> ReLU -> 2x2x2 max  -> pooled bf16 [N,D/2,H/2,W/2,16] + one code byte per pooled value
//               (bits 0-2: argmax position d*4+h*2+w inside the window, bit 3: ReLU active)
//   MODE_BWD    conv (recomputed, bit-identical) -> dy = BN/ReLU/pool backward of the pooled gradient, rounded to
//               bf16 and written to SHARED MEMORY as the MN-major operand of a second tcgen05 GEMM that accumulates
//               the weight gradient dW1 in TMEM across the whole persistent CTA.  dy never exists in global memory.
//
// Geometry.  The padded input matrix XP2[(n, h', d'), w'] (bf16, zero halo, rows ordered with d' fastest) is read by
// TMA in [136 rows][16 w'] windows (32-byte rows, 32B swizzle), one window per (output block of 8 w, input line h').
// A GEMM row tile is one d-line: tile row r <-> output d = r, so the operand of filter tap (kd, kh) is the window of
// line h+kh shifted by kd rows (+kd*32 bytes on the descriptor).  A work unit is a PAIR of output lines (h, h+1):
// the 2x2x2 pooling window is then (two accumulators) x (two adjacent columns of one thread) x (two adjacent lanes).
// The same 32-byte-row window is the K-major A operand of the convolution (M = d rows, K = 16 w') and the MN-major B
// operand of the weight-gradient GEMM (K = d rows, N = 3 kd-shifted blocks of 16 w', LBO = one row):
//
//   dW~[kh][(wl, c), kd*16 + k] += sum_rows dy[row, (wl, c)] * XP2[line h+kh, row + kd, 8j + k],   dW1[c,kd,kh,kw] = sum_wl dW~[kh][(wl,c), kd*16 + wl + kw]
#include "umma.cuh"
#include <cstdlib>

namespace coinn {

enum { C1F_STATS = 0, C1F_POOL = 1, C1F_BWD = 2 };

constexpr int C1F_THREADS = 320;                       // warp 0: TMA, warp 1: MMA, warps 2-9: two epilogue groups
constexpr int C1F_ROWS = 136;                          // 128 + 2 halo rows, rounded up to 8
constexpr uint32_t C1F_SLAB = C1F_ROWS * 32;           // BWD: one window [136 rows][16 w'] per output block (32B swizzle)
constexpr uint32_t C1F_STAGE = 4 * C1F_SLAB;           // input lines h' .. h'+3 of one output block (17 KB)
constexpr uint32_t C1F_WSLAB = C1F_ROWS * 128;         // STATS / POOL: one window [136 rows][64 w'] serves 7 output blocks (128B swizzle)
constexpr uint32_t C1F_WSTAGE = 4 * C1F_WSLAB;         // 68 KB
constexpr uint32_t C1F_T_BYTES = 9 * 4096;             // Toeplitz matrices T_{kd,kh}: [128 n][16 k] bf16, K-major, 32B swizzle
constexpr uint32_t C1F_DY_BYTES = 32768;               // dy tile: 2 M-blocks x [128 rows][128 B], 128B swizzle
constexpr int C1F_MAX_STAGES = 8;

struct C1FParams {
    const float* w;             // [16][27] fp32
    int N, D, H, W;
    int Dp, Hp;                 // D + 2, H + 2
    int nblk, groups, bpg;      // ceil(W / 8) output blocks per line; groups of bpg blocks (8, or 7 per 64-column window)
    int tpl, hpairs;            // 128-row tiles per d-line, ceil(H / 2)
    int num_units;              // N * hpairs * tpl * groups
    int stages;
    float* stats;               // STATS: [32] (zeroed)
    const float* mean;          // POOL / BWD: BatchNorm batch statistics and affine parameters [16]
    const float* invstd;
    const float* gamma;
    const float* beta;
    __nv_bfloat16* pooled;      // POOL out / unused
    uint8_t* code;              // POOL out, BWD in
    const __nv_bfloat16* dpool; // BWD: gradient of the pooled output
    const float* acc;           // BWD: [32] sum(g), sum(g * xhat) over the batch (bn_pool_bwd_stats_pooled)
    float inv_count;            // BWD: 1 / (N*D*H*W)
    float* dw;                  // BWD: [16][27] fp32 (zeroed)
    int dbg;                    // COINN_C1F_DEBUG (profiling only): 1 = STATS epilogue skips the math, 2 = also the TMEM loads
};

// xp2[(n*Hp + h')*Dp + d'][w'] = x[n, d'-1, h'-1, w'-1] (zero outside); one thread per 16-byte chunk
template <typename TX>
__global__ void conv1_pad_input_hd_kernel(const TX* __restrict__ x, __nv_bfloat16* __restrict__ xp, int N, int D, int H, int W,
                                          int Dp, int Hp, int chunks, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % chunks);
    long long r = i / chunks;
    const int dp = (int)(r % Dp); r /= Dp;
    const int hp = (int)(r % Hp);
    const int n = (int)(r / Hp);
    uint32_t out[4] = {0u, 0u, 0u, 0u};
    if (dp >= 1 && dp <= D && hp >= 1 && hp <= H) {
        const TX* row = x + (((long long)n * D + (dp - 1)) * H + (hp - 1)) * W;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int wi = ch * 8 + e - 1;
            v[e] = (wi >= 0 && wi < W) ? (float)row[wi] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
    }
    reinterpret_cast<uint4*>(xp)[i] = make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ void c1f_unit(const C1FParams& p, int u, int& n, int& hp, int& t, int& g) {
    g = u % p.groups; u /= p.groups;
    t = u % p.tpl; u /= p.tpl;
    hp = u % p.hpairs;
    n = u / p.hpairs;
}

__device__ __forceinline__ void st_shared_128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_global_nc_128(const void* ptr) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(C1F_THREADS, 1)
conv1_fused_kernel(const __grid_constant__ CUtensorMap tmap_xp, const C1FParams p) {
    constexpr int NBUF = MODE == C1F_BWD ? 2 : 4;            // conv accumulators (128 TMEM columns each)
    // TMA moves ~one box row per 3-4 cycles whatever its width, and 32-byte rows made the recompute kernels TMA-bound
    // (ncu: tensor pipe 45 % active, epilogue warps starved).  STATS / POOL therefore load 128-byte rows: a window of
    // 64 input columns is the operand of 7 output blocks, block m starting 16*m bytes into the swizzled row.
    constexpr bool WIDE = MODE != C1F_BWD;
    constexpr uint32_t SLAB = WIDE ? C1F_WSLAB : C1F_SLAB;
    constexpr uint32_t STAGE = 4 * SLAB;
    constexpr uint32_t WG_COL = 256;                         // BWD: weight-gradient accumulators (48 columns) at columns 256 + kh*64

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* t_smem = smem;
    uint8_t* stage_base = smem + C1F_T_BYTES;
    uint8_t* dy_smem = stage_base + (size_t)p.stages * STAGE;            // BWD only: 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(dy_smem + (MODE == C1F_BWD ? 2 * C1F_DY_BYTES : 0));
    uint64_t* full_bar = bars;                               // [stages] TMA -> MMA
    uint64_t* empty_bar = bars + C1F_MAX_STAGES;             // [stages] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * C1F_MAX_STAGES;         // [4] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 4;                    // [4] epilogue -> MMA
    uint64_t* dy_ready = tmem_empty + 4;                     // [2] epilogue -> MMA  (BWD)
    uint64_t* dy_free = dy_ready + 2;                        // [2] MMA -> epilogue  (BWD)
    uint64_t* done_bar = dy_free + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);    // [32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = p.stages;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_xp);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 4; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], MODE == C1F_BWD ? 8 : 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(&dy_ready[b], 8); mbar_init(&dy_free[b], 1); }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) red[threadIdx.x] = 0.f;
    // banded Toeplitz matrices: element (n = wl*16 + c, k) of T_s (s = kd*3 + kh) is W1[c, kd, kh, k - wl]
    for (int i = threadIdx.x; i < 9 * 128 * 16; i += C1F_THREADS) {
        const int s = i >> 11, n = (i >> 4) & 127, k = i & 15;
        const int wl = n >> 4, c = n & 15, kw = k - wl;
        const float v = (kw >= 0 && kw <= 2) ? p.w[c * 27 + s * 3 + kw] : 0.f;
        // K-major rows of 32 bytes with the 32B swizzle (16-byte chunk index ^= address bit 7 = bit 2 of n): 8 % faster
        // MMAs than the unswizzled core-matrix planes (84.6 -> 77.6 us for the statistics pass)
        reinterpret_cast<__nv_bfloat16*>(t_smem + s * 4096 + n * 32 + ((((k >> 3) ^ ((n >> 2) & 1))) << 4))[k & 7] = __float2bfloat16(v);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int first = blockIdx.x, step = gridDim.x;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------ TMA producer
        {
            const bool leader = elect_one();                               // uniform loop, one issuing lane (see the MMA warp)
            uint32_t it = 0;
            for (int u = first; u < p.num_units; u += step) {
                int n, hp, t, g;
                c1f_unit(p, u, n, hp, t, g);
                const int nb = (p.nblk - g * p.bpg) < p.bpg ? (p.nblk - g * p.bpg) : p.bpg;
                const int row0 = (n * p.Hp + 2 * hp) * p.Dp + t * 128;
                const int nloads = WIDE ? 1 : nb;                          // one window per unit / per output block
                for (int jj = 0; jj < nloads; ++jj, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* dst = stage_base + (size_t)s * STAGE;
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], STAGE);
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (leader) tma_load_2d(dst + l * SLAB, &tmap_xp, &full_bar[s], (g * p.bpg + jj) * 8, row0 + l * p.Dp);
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------------------------------- MMA issuer
        // The WHOLE warp runs this loop and one elected lane issues.  With the loop under `if (lane == 0)` ptxas cannot
        // prove that the descriptors are warp-uniform and wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST
        // "waterfall" (7 dependent R2UR per MMA): ~150 cycles of issue latency per MMA against a 64-cycle tensor floor,
        // i.e. the tensor pipe sat idle 55 % of the time (profiles/ncu_conv1_fused_v1.txt).  Uniform control flow lets
        // the descriptor arithmetic live in uniform registers.
        {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc_f16(128, p.dbg == 3 ? 64 : 128, 1, 0, 0);      // dbg 3: half-width MMAs (timing experiment)
            const uint32_t kd_rows = p.dbg == 5 ? 0u : (WIDE ? 8u : 2u);                     // dbg 5: no row shift (timing experiment)
            constexpr uint32_t idesc_w = make_idesc_f16(128, 48, 1, 1, 1);
            const uint64_t a_const = WIDE ? make_smem_desc(0, 16, 1024, SMEM_LAYOUT_SW128)  // K-major, 128-byte rows
                                          : make_smem_desc(0, 16, 256, SMEM_LAYOUT_SW32);   // K-major, 32-byte rows
            const uint64_t b_const = make_smem_desc(0, 16, 256, SMEM_LAYOUT_SW32);
            const uint64_t wa_const = make_smem_desc(0, 16384, 1024, SMEM_LAYOUT_SW128);   // dy: MN-major, 2 blocks of 64
            const uint64_t wb_const = make_smem_desc(0, 32, 256, SMEM_LAYOUT_SW32);        // window: MN-major, block = row shift
            const uint32_t t16 = (smem_u32(t_smem) & 0x3FFFFu) >> 4;
            const uint32_t dy16 = (smem_u32(dy_smem) & 0x3FFFFu) >> 4;
            uint32_t it = 0, hb = 0;
            uint32_t prev_slab16 = 0, prev_stage = 0, prev_ab = 0;      // BWD: operands of the half-block awaiting its wgrad
            auto issue_wgrad = [&](uint32_t h) {
                const uint32_t bb = h & 1;
                mbar_wait(&dy_ready[bb], (h >> 1) & 1);
                tcgen05_after_sync();
                const uint32_t a0 = dy16 + bb * (C1F_DY_BYTES / 16);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
                        if (leader) umma_f16(tmem_base + WG_COL + kh * 64, wa_const | (a0 + k * 128), wb_const | (prev_slab16 + kh * (C1F_SLAB / 16) + k * 32),
                                 idesc_w, (h | (uint32_t)k) ? 1u : 0u);
                }
                if (leader) { umma_commit(&dy_free[bb]); if (prev_ab) umma_commit(&empty_bar[prev_stage]); }
                __syncwarp();
            };
            for (int u = first; u < p.num_units; u += step) {
                int n, hp, t, g;
                c1f_unit(p, u, n, hp, t, g);
                const int nb = (p.nblk - g * p.bpg) < p.bpg ? (p.nblk - g * p.bpg) : p.bpg;
                for (int jj = 0; jj < nb; ++jj) {
                    const uint32_t s = it % STAGES;
                    if (!WIDE || jj == 0) mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    const uint32_t st16 = (smem_u32(stage_base + (size_t)s * STAGE) & 0x3FFFFu) >> 4;
#pragma unroll
                    for (int ab = 0; ab < 2; ++ab, ++hb) {
                        const uint32_t b = hb & (NBUF - 1);
                        mbar_wait(&tmem_empty[b], ((hb / NBUF) & 1) ^ 1);
                        tcgen05_after_sync();
                        const uint32_t d_tmem = tmem_base + b * 128;
                        const uint32_t a0 = st16 + ab * (SLAB / 16) + (WIDE ? (uint32_t)jj : 0u);   // block jj: +16 bytes per block
#pragma unroll
                        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh)
                                if (leader) umma_f16(d_tmem, a_const | (a0 + kh * (SLAB / 16) + kd * kd_rows), b_const | (t16 + (kd * 3 + kh) * 256), idesc,
                                         (kd | kh) ? 1u : 0u);
                        }
                        if (leader) umma_commit(&tmem_full[b]);
                        __syncwarp();
                        if (MODE == C1F_BWD) {
                            if (hb > 0) issue_wgrad(hb - 1);
                            prev_slab16 = a0; prev_stage = s; prev_ab = ab;
                        }
                    }
                    if (WIDE ? (jj == nb - 1) : false) { if (leader) umma_commit(&empty_bar[s]); ++it; }
                    if (!WIDE) ++it;
                }
            }
            if (MODE == C1F_BWD) {
                if (hb > 0) issue_wgrad(hb - 1);
                if (leader) umma_commit(done_bar);
            }
        }
    } else {
        // ----------------------------------------------------------------------------------------- epilogue
        const int ew = warp - 2, q = warp & 3, grp = ew >> 2;          // TMEM lane quarter = warp % 4
        const int r = q * 32 + lane;                                    // tile row = output d (within the tile)
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int PD = p.D >> 1, PH = p.H >> 1, PW = p.W >> 1;

        if (MODE == C1F_STATS) {
            float s1[16], s2[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
            uint32_t blk = 0;
            for (int u = first; u < p.num_units; u += step) {
                int n, hp, t, g;
                c1f_unit(p, u, n, hp, t, g);
                const int nb = (p.nblk - g * p.bpg) < p.bpg ? (p.nblk - g * p.bpg) : p.bpg;
                const bool row_ok = t * 128 + r < p.D;
                for (int jj = 0; jj < nb; ++jj, ++blk) {
                    if ((blk & 1u) != (uint32_t)grp) continue;
                    const int w0 = (g * p.bpg + jj) * 8;
#pragma unroll
                    for (int ab = 0; ab < 2; ++ab) {
                        const uint32_t b = 2 * (blk & 1) + ab;
                        mbar_wait(&tmem_full[b], (blk >> 1) & 1);
                        tcgen05_after_sync();
                        const bool line_ok = row_ok && (2 * hp + ab) < p.H;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            if (p.dbg >= 2) break;
                            uint32_t v[4][16];
#pragma unroll
                            for (int i = 0; i < 4; ++i) tmem_ld_32x32b_x16(tmem_base + lane_off + b * 128 + (half * 4 + i) * 16, v[i]);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (line_ok && w0 + half * 4 + i < p.W && (p.dbg == 0 || v[i][0] == 0x7fc01234u)) {
#pragma unroll
                                    for (int c = 0; c < 16; ++c) {
                                        const float f = __uint_as_float(v[i][c]);
                                        s1[c] += f; s2[c] = fmaf(f, f, s2[c]);
                                    }
                                }
                            }
                        }
                        tcgen05_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[b]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float sa = warp_sum(s1[c]), sb = warp_sum(s2[c]);
                if (lane == 0) { atomicAdd(&red[c], sa); atomicAdd(&red[16 + c], sb); }
            }
        } else if (MODE == C1F_POOL) {
            float sc[16], sh[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) { sc[c] = p.gamma[c] * p.invstd[c]; sh[c] = p.beta[c] - p.mean[c] * sc[c]; }
            const bool odd = lane & 1;
            const uint32_t dbit = odd ? 4u : 0u;
            uint32_t blk = 0;
            for (int u = first; u < p.num_units; u += step) {
                int n, hp, t, g;
                c1f_unit(p, u, n, hp, t, g);
                const int nb = (p.nblk - g * p.bpg) < p.bpg ? (p.nblk - g * p.bpg) : p.bpg;
                const int pd = (t * 128 + r) >> 1;
                const bool cell_row_ok = pd < PD && hp < PH;
                const long long cell_row = (((long long)n * PD + pd) * PH + hp) * PW;
                for (int jj = 0; jj < nb; ++jj, ++blk) {
                    if ((blk & 1u) != (uint32_t)grp) continue;
                    const int w0 = (g * p.bpg + jj) * 8;
                    const uint32_t bA = 2 * (blk & 1), bB = bA + 1;
                    mbar_wait(&tmem_full[bA], (blk >> 1) & 1);
                    mbar_wait(&tmem_full[bB], (blk >> 1) & 1);
                    tcgen05_after_sync();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                          // pooled column pw = (w0 >> 1) + i
                        uint32_t v[4][16];                                 // (h even, w even), (h even, w odd), (h odd, w even), (h odd, w odd)
                        tmem_ld_32x32b_x16(tmem_base + lane_off + bA * 128 + (2 * i) * 16, v[0]);
                        tmem_ld_32x32b_x16(tmem_base + lane_off + bA * 128 + (2 * i + 1) * 16, v[1]);
                        tmem_ld_32x32b_x16(tmem_base + lane_off + bB * 128 + (2 * i) * 16, v[2]);
                        tmem_ld_32x32b_x16(tmem_base + lane_off + bB * 128 + (2 * i + 1) * 16, v[3]);
                        tmem_ld_wait();
                        // The window position (3 bits) rides in the low mantissa bits of the candidate, so the whole
                        // 2x2x2 arg-max is FMNMX on single registers (7 fp32 ulps of noise, far below the bf16 output;
                        // exact ties go to the larger position instead of the smaller one).
                        float m[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            float e[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float z = fmaf(__uint_as_float(v[j][c]), sc[c], sh[c]);
                                e[j] = __uint_as_float((__float_as_uint(z) & ~7u) | (dbit | (uint32_t)j));
                            }
                            m[c] = fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3]));
                        }
                        uint32_t pk[4] = {0u, 0u, 0u, 0u};                  // this lane's 8 channels (even lane 0-7, odd lane 8-15), bf16
                        uint32_t ck[2] = {0u, 0u};                          // ... and their code bytes
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            const float send = odd ? m[cc] : m[cc + 8];     // the partner lane (d ^ 1) finalises the other 8 channels
                            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                            const float best = fmaxf(odd ? m[cc + 8] : m[cc], recv);
                            const uint32_t bits = __float_as_uint(best);
                            const float val = __uint_as_float(bits & ~7u);
                            const bool active = val > 0.f;
                            const uint32_t hbits = active ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16(val)) : 0u;
                            pk[cc >> 1] |= hbits << (16 * (cc & 1));
                            ck[cc >> 2] |= ((bits & 7u) | (active ? 8u : 0u)) << (8 * (cc & 3));
                        }
                        const int pw = (w0 >> 1) + i;
                        if (cell_row_ok && pw < PW) {
                            const long long cell = cell_row + pw;
                            *reinterpret_cast<uint4*>(p.pooled + cell * 16 + (odd ? 8 : 0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            *reinterpret_cast<uint2*>(p.code + cell * 16 + (odd ? 8 : 0)) = make_uint2(ck[0], ck[1]);
                        }
                    }
                    tcgen05_before_sync();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&tmem_empty[bA]); mbar_arrive(&tmem_empty[bB]); }
                }
            }
        } else {
            // dy = sc * (g - c1 - xhat * c2) = S*g + (A + B*y)
            float cA[16], cB[16];
            float* cS = red;                                               // S lives in shared memory (16 fewer live registers)
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float is = p.invstd[c];
                const float S = p.gamma[c] * is;
                cB[c] = -S * (p.acc[16 + c] * p.inv_count) * is;
                cA[c] = -S * (p.acc[c] * p.inv_count) - cB[c] * p.mean[c];
                if (ew == 0 && lane == 0) cS[c] = S;
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");               // the 8 epilogue warps
            const uint32_t dy_addr0 = smem_u32(dy_smem) + (uint32_t)r * 128u;
            const uint32_t sw = (uint32_t)(r & 7);
            // Both epilogue groups work on EVERY half-block (group g owns voxel pairs 2g, 2g+1 = M block g of the dy tile):
            // the latency of one half-block's epilogue - which the next wgrad MMAs wait for - is half of what it is when
            // the groups alternate between half-blocks.
            uint32_t hb = 0;
            for (int u = first; u < p.num_units; u += step) {
                int n, hp, t, g;
                c1f_unit(p, u, n, hp, t, g);
                const int nb = (p.nblk - g * p.bpg) < p.bpg ? (p.nblk - g * p.bpg) : p.bpg;
                const int d = t * 128 + r;
                for (int jj = 0; jj < nb; ++jj) {
                    const int w0 = (g * p.bpg + jj) * 8;
#pragma unroll
                    for (int ab = 0; ab < 2; ++ab, ++hb) {
                        const int h = 2 * hp + ab;                         // line A / line B of the pair
                        const bool row_ok = d < p.D && h < p.H;
                        const bool cell_row_ok = row_ok && (d >> 1) < PD && (h >> 1) < PH;
                        const long long cell_row = (((long long)n * PD + (d >> 1)) * PH + (h >> 1)) * PW;
                        const uint32_t pos_dh = (uint32_t)(((d & 1) << 2) | ((h & 1) << 1)) | 8u;
                        const uint32_t b = (uint32_t)ab;
                        // the global loads of this half-block (arg-max codes + pooled gradients of this group's 2 pool cells)
                        // are issued before the barrier waits: their latency hides behind the MMAs
                        uint4 cdv[2], g0v[2], g1v[2];
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {
                            const int pw = (w0 >> 1) + 2 * grp + ii;
                            cdv[ii] = make_uint4(0u, 0u, 0u, 0u); g0v[ii] = cdv[ii]; g1v[ii] = cdv[ii];
                            if (cell_row_ok && pw < PW) {
                                const long long cell = cell_row + pw;
                                cdv[ii] = ld_global_nc_128(p.code + cell * 16);
                                g0v[ii] = ld_global_nc_128(p.dpool + cell * 16);
                                g1v[ii] = ld_global_nc_128(p.dpool + cell * 16 + 8);
                            }
                        }
                        if (ab == 0 && jj + 1 < nb && cell_row_ok) {        // next block's cells -> L1 (lines A and B share them)
#pragma unroll
                            for (int ii = 0; ii < 2; ++ii) {
                                const int pw = (w0 >> 1) + 4 + 2 * grp + ii;
                                if (pw < PW) {
                                    asm volatile("prefetch.global.L1 [%0];" :: "l"(p.code + (cell_row + pw) * 16));
                                    asm volatile("prefetch.global.L1 [%0];" :: "l"(p.dpool + (cell_row + pw) * 16));
                                }
                            }
                        }
                        mbar_wait(&tmem_full[b], (hb >> 1) & 1);
                        mbar_wait(&dy_free[b], ((hb >> 1) & 1) ^ 1);       // wgrad MMAs of the previous tile in this buffer are done
                        tcgen05_after_sync();
                        const uint32_t dy_row = dy_addr0 + b * C1F_DY_BYTES + (uint32_t)grp * 16384u;
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {                   // voxel pair (w0 + 2i, w0 + 2i + 1) shares one pool cell
                            const int i = 2 * grp + ii;
                            uint32_t v[2][16];
                            tmem_ld_32x32b_x16(tmem_base + lane_off + b * 128 + (2 * i) * 16, v[0]);
                            tmem_ld_32x32b_x16(tmem_base + lane_off + b * 128 + (2 * i + 1) * 16, v[1]);
                            const uint4 cd = cdv[ii], g0 = g0v[ii], g1 = g1v[ii];
                            const uint32_t cdw[4] = {cd.x, cd.y, cd.z, cd.w};
                            const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                            float gsc[16];
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float2 f = unpack_bf16x2(gw[c]);
                                gsc[2 * c] = f.x * cS[2 * c]; gsc[2 * c + 1] = f.y * cS[2 * c + 1];
                            }
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int w = w0 + 2 * i + e;
                                const bool vox_ok = row_ok && w < p.W;
                                const uint32_t match = pos_dh | (uint32_t)e;
                                float o[16];
#pragma unroll
                                for (int c = 0; c < 16; ++c) {
                                    const bool hit = ((cdw[c >> 2] >> (8 * (c & 3))) & 0xFFu) == match;
                                    float t0 = fmaf(cB[c], __uint_as_float(v[e][c]), cA[c]);
                                    t0 += hit ? gsc[c] : 0.f;
                                    o[c] = vox_ok ? t0 : 0.f;
                                }
                                // voxel wl = 2i + e -> M block (wl >> 2) == grp, 16-byte chunks ((wl & 3) * 2, +1), 128B swizzle
                                const uint32_t c0 = (uint32_t)((2 * ii + e) * 2);
                                st_shared_128(dy_row + (((c0) ^ sw) << 4), pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                              pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
                                st_shared_128(dy_row + (((c0 + 1u) ^ sw) << 4), pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]),
                                              pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
                            }
                        }
                        fence_proxy_async_smem();
                        tcgen05_before_sync();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&tmem_empty[b]); mbar_arrive(&dy_ready[b]); }
                    }
                }
            }
            // ---- final: extract the 27 diagonals of the three accumulators and merge the CTAs
            if (grp == 0 && first < p.num_units) {
                float* scratch = reinterpret_cast<float*>(dy_smem);                       // [128][49] (padded)
                float* dw_s = reinterpret_cast<float*>(dy_smem + C1F_DY_BYTES);          // [432]
                mbar_wait(done_bar, 0);
                tcgen05_after_sync();
                for (int i = r; i < 432; i += 128) dw_s[i] = 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int wl = r >> 4, c = r & 15;
#pragma unroll 1
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(tmem_base + lane_off + WG_COL + kh * 64 + part * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 16; ++k) scratch[r * 49 + part * 16 + k] = __uint_as_float(v[k]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            atomicAdd(&dw_s[c * 27 + (kd * 3 + kh) * 3 + kw], scratch[r * 49 + kd * 16 + wl + kw]);
                    }
                    __syncwarp();
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = r; i < 432; i += 128) atomicAdd(&p.dw[i], dw_s[i]);
            }
        }
    }
    tcgen05_before_sync();
    __syncthreads();
    if (MODE == C1F_STATS && threadIdx.x < 32) atomicAdd(&p.stats[threadIdx.x], red[threadIdx.x]);
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static inline void c1f_geometry(int N, int D, int H, int W, C1FParams& p, int& Wq) {
    p.N = N; p.D = D; p.H = H; p.W = W;
    p.Dp = D + 2; p.Hp = H + 2;
    p.nblk = (W + 7) / 8;
    p.groups = (p.nblk + p.bpg - 1) / p.bpg;
    p.tpl = (D + 127) / 128;
    p.hpairs = (H + 1) / 2;
    p.num_units = N * p.hpairs * p.tpl * p.groups;
    Wq = 8 * (p.nblk + 1);
}

template <int MODE>
static int c1f_launch(const void* xp, C1FParams& p, int N, int D, int H, int W, cudaStream_t st) {
    int Wq;
    constexpr bool WIDE = MODE != C1F_BWD;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("COINN_C1F_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
    p.bpg = WIDE ? 7 : 8;
    c1f_geometry(N, D, H, W, p, Wq);
    const long long rows = (long long)N * p.Hp * p.Dp;
    if (rows + 4LL * p.Dp + 512 >= (1LL << 31)) return -1;
    p.stages = WIDE ? 2 : 5;
    const int smem_bytes = (int)C1F_T_BYTES + p.stages * (int)(WIDE ? C1F_WSTAGE : C1F_STAGE) + (WIDE ? 0 : 2 * (int)C1F_DY_BYTES) + 1024 + 1024;
    CUtensorMap tx;
    if (make_tmap_2d_bf16(&tx, xp, (uint64_t)rows, (uint64_t)Wq, (uint64_t)Wq * 2, C1F_ROWS, WIDE ? 64 : 16,
                          WIDE ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B) != 0) return -3;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv1_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int grid = p.num_units < B200_SM_COUNT ? p.num_units : B200_SM_COUNT;
    conv1_fused_kernel<MODE><<<grid, C1F_THREADS, smem_bytes, st>>>(tx, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// rows and row length (elements) of the padded bf16 input matrix for an [N, D, H, W] volume:
// (D+2)*(H+2) rows per sample, W rounded up to whole 8-element chunks plus one zero chunk
COINN_API int coinn_conv1_padded_shape(int N, int D, int H, int W, long long* rows, int* cols) {
    *rows = (long long)N * (D + 2) * (H + 2);
    *cols = 8 * ((W + 7) / 8 + 1);
    return 0;
}

// x: [N,D,H,W] fp32 (x_dtype 0) or bf16 (1)  ->  xp: [N*(H+2)*(D+2), Wq] bf16, zero halo, d' fastest
COINN_API int coinn_conv1_pad_input_hd(const void* x, int x_dtype, void* xp, int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    const int Dp = D + 2, Hp = H + 2, chunks = (W + 7) / 8 + 1;
    const long long total = (long long)N * Dp * Hp * chunks;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_dtype == 0)
        conv1_pad_input_hd_kernel<float><<<(unsigned)blocks, threads, 0, st>>>((const float*)x, (__nv_bfloat16*)xp, N, D, H, W, Dp, Hp, chunks, total);
    else
        conv1_pad_input_hd_kernel<__nv_bfloat16><<<(unsigned)blocks, threads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)xp, N, D, H, W, Dp, Hp, chunks, total);
    COINN_CHECK_LAUNCH();
    return 0;
}

// stats[32] (zeroed) += per-channel sum / sum of squares of conv1(x) (fp32 accumulators)
COINN_API int coinn_conv1_fused_stats(const void* xp, const float* w, float* stats, int N, int D, int H, int W, void* stream) {
    coinn::C1FParams p{};
    p.w = w; p.stats = stats;
    return coinn::c1f_launch<coinn::C1F_STATS>(xp, p, N, D, H, W, reinterpret_cast<cudaStream_t>(stream));
}

// pooled [N,D/2,H/2,W/2,16] bf16 = maxpool2(relu(bn(conv1(x)))), code: one byte per pooled value
COINN_API int coinn_conv1_fused_pool(const void* xp, const float* w, const float* mean, const float* invstd, const float* gamma,
                                     const float* beta, void* pooled, void* code, int N, int D, int H, int W, void* stream) {
    coinn::C1FParams p{};
    p.w = w; p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.beta = beta;
    p.pooled = reinterpret_cast<__nv_bfloat16*>(pooled); p.code = reinterpret_cast<uint8_t*>(code);
    return coinn::c1f_launch<coinn::C1F_POOL>(xp, p, N, D, H, W, reinterpret_cast<cudaStream_t>(stream));
}

// dw[16*27] (zeroed) += conv1 weight gradient; acc[32] = (sum g, sum g*xhat) from coinn_bn_pool_bwd_stats_pooled
COINN_API int coinn_conv1_fused_bwd(const void* xp, const float* w, const float* mean, const float* invstd, const float* gamma,
                                    const float* acc, const void* dpool, const void* code, float* dw, int N, int D, int H, int W,
                                    void* stream) {
    coinn::C1FParams p{};
    p.w = w; p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.acc = acc;
    p.dpool = reinterpret_cast<const __nv_bfloat16*>(dpool); p.code = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(code));
    p.dw = dw; p.inv_count = 1.f / ((float)N * D * H * W);
    return coinn::c1f_launch<coinn::C1F_BWD>(xp, p, N, D, H, W, reinterpret_cast<cudaStream_t>(stream));
}
