// First VBM block (C_in = 1) as a banded-Toeplitz GEMM on tcgen05: no im2col is ever built.
//
// With one input channel an im2col row holds only 27 values, so the K dimension of an implicit GEMM is tiny and the
// kernel is bound by the CUDA cores that assemble the rows (conv1_tc.cu: 19 instructions per output voxel, 535 us
// for 8 subjects).  Here the W axis of the volume becomes the GEMM K axis instead:
//
//     Y[(n,d,h), (w, c)] = sum_{kd,kh}  X[(n, d+kd, h+kh), w'] * T_{kd,kh}[w', (w, c)],
//     T_{kd,kh}[w', (w, c)] = W1[c, kd, kh, w' - w]   (zero unless 0 <= w' - w <= 2)
//
// T is banded, so for a block of 8 output columns only 10 input columns matter: one K = 16 tcgen05.mma per
// (kd, kh) with N = 8 w x 16 c = 128.  The input is stored once as a zero-padded bf16 matrix
// XP[(n, d', h'), w'] (d' < D+2, h' < H+2, row length Wq = 8 * (ceil(W/8) + 1)), so that
//   * a GEMM row tile is 128 CONSECUTIVE padded rows, and the operand of (kd, kh) is the same matrix shifted by
//     kd*(H+2) + kh rows - an affine address, no bounds logic (rows that are padding are computed and dropped, 3 %);
//   * TMA drops 16-byte column chunks [row][8 w'] straight into the K-major no-swizzle core-matrix layout
//     (SBO = 128 B between 8-row groups, LBO = one chunk plane), where the operand of output block j starts at
//     chunk j and the kh shift is +kh*16 bytes on the descriptor start address.
// The 9 T matrices (36 KB) are built once per persistent CTA.  Accumulators: 4 x 128 TMEM columns, drained by
// 8 epilogue warps; a thread owns one volume row and therefore writes 8 voxels x 16 channels = 256 contiguous
// bytes of the channels-last output per block, and keeps the BatchNorm sums of its row in registers.
// 144 MMAs of 128x128x16 per 128-row tile = 5.3x the minimal FLOPs, still < 60 us of tensor time per step: the
// kernel is bound by writing y.
#include "umma.cuh"

namespace coinn {

constexpr int C1T_THREADS = 320;                       // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
constexpr int C1T_ROWS = 136;                          // rows per slab: 128 + 2 halo, rounded up to 8
constexpr int C1T_CHUNKS = 9;                          // 8 output blocks need chunks j .. j+1
constexpr uint32_t C1T_CHUNK_BYTES = C1T_ROWS * 16;    // one chunk plane [136 rows][8 bf16]
constexpr uint32_t C1T_SLAB = C1T_CHUNKS * C1T_CHUNK_BYTES;
constexpr uint32_t C1T_STAGE = 3 * C1T_SLAB;           // three d-planes
constexpr uint32_t C1T_B_BYTES = 9 * 4096;             // T_{kd,kh}: [2 k-chunks][128 n][8 k] bf16
constexpr int C1T_MAX_STAGES = 4;

struct C1TParams {
    __nv_bfloat16* y;           // [N, D, H, W, 16]
    float* stats;               // [32]: sum, sum of squares of the stored values
    const float* w;             // [16][27] fp32
    int N, D, H, W;
    int Dp, Hp;                 // D + 2, H + 2
    int nblk, groups;           // ceil(W / 8), ceil(nblk / 8)
    int num_units;              // row tiles x groups
    int stages;
};

// ---------------------------------------------------------------------------------------------- input padding
// xp[(n*Dp + d')*Hp + h'][w'] = x[n, d'-1, h'-1, w'-1] (zero outside); one thread per 16-byte chunk
template <typename TX>
__global__ void conv1_pad_input_kernel(const TX* __restrict__ x, __nv_bfloat16* __restrict__ xp, int N, int D, int H, int W,
                                       int Dp, int Hp, int chunks, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % chunks);
    long long r = i / chunks;
    const int hp = (int)(r % Hp); r /= Hp;
    const int dp = (int)(r % Dp);
    const int n = (int)(r / Dp);
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

// 256-bit store (sm_100: STG.256): a lane writes a whole 32-byte sector, so the row-strided epilogue stores move
// twice the bytes per LSU request of two 16-byte stores and never leave half-written sectors in L2
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// ----------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(C1T_THREADS, 1)
conv1_toeplitz_fwd_kernel(const __grid_constant__ CUtensorMap tmap_xp, const C1TParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_smem = smem;
    uint8_t* stage_base = smem + C1T_B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)p.stages * C1T_STAGE);
    uint64_t* full_bar = bars;                               // [stages] TMA -> MMA
    uint64_t* empty_bar = bars + C1T_MAX_STAGES;             // [stages] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * C1T_MAX_STAGES;         // [4] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 4;                    // [4] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);    // [32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = p.stages;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_xp);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 4; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        fence_mbar_init();
    }
    if (threadIdx.x < 32) red[threadIdx.x] = 0.f;
    // banded Toeplitz matrices: element (n = wl*16 + c, k) of T_s is W1[c, s, k - wl]
    for (int i = threadIdx.x; i < 9 * 128 * 16; i += C1T_THREADS) {
        const int s = i >> 11, n = (i >> 4) & 127, k = i & 15;
        const int wl = n >> 4, c = n & 15, kw = k - wl;
        const float v = (kw >= 0 && kw <= 2) ? p.w[c * 27 + s * 3 + kw] : 0.f;
        reinterpret_cast<__nv_bfloat16*>(b_smem + s * 4096 + (k >> 3) * 2048 + n * 16)[k & 7] = __float2bfloat16(v);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int first = blockIdx.x, step = gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int u = first; u < p.num_units; u += step, ++it) {
                const int tile = u / p.groups, g = u - tile * p.groups;
                const int s = it % STAGES;
                mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                uint8_t* dst = stage_base + (size_t)s * C1T_STAGE;
                mbar_arrive_expect_tx(&full_bar[s], C1T_STAGE);
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                    for (int c = 0; c < C1T_CHUNKS; ++c)
                        tma_load_2d(dst + kd * C1T_SLAB + c * C1T_CHUNK_BYTES, &tmap_xp, &full_bar[s], (g * 8 + c) * 8,
                                    tile * 128 + kd * p.Hp);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, 128, 1, 0, 0);
            const uint64_t a_const = make_smem_desc(0, C1T_CHUNK_BYTES, 128, SMEM_LAYOUT_NONE);
            const uint64_t b_const = make_smem_desc(0, 2048, 128, SMEM_LAYOUT_NONE);
            const uint32_t b16 = (smem_u32(b_smem) & 0x3FFFFu) >> 4;
            uint32_t it = 0, blk = 0;
            for (int u = first; u < p.num_units; u += step, ++it) {
                const int tile = u / p.groups, g = u - tile * p.groups;
                const int nb = (p.nblk - g * 8) < 8 ? (p.nblk - g * 8) : 8;
                const int s = it % STAGES;
                mbar_wait(&full_bar[s], (it / STAGES) & 1);
                const uint32_t st16 = (smem_u32(stage_base + (size_t)s * C1T_STAGE) & 0x3FFFFu) >> 4;
                for (int jj = 0; jj < nb; ++jj, ++blk) {
                    const uint32_t b = blk & 3;
                    mbar_wait(&tmem_empty[b], ((blk >> 2) & 1) ^ 1);
                    tcgen05_after_sync();
                    const uint32_t d_tmem = tmem_base + b * 128;
                    const uint32_t a0 = st16 + jj * (C1T_CHUNK_BYTES / 16);
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh)
                            umma_f16(d_tmem, a_const | (a0 + kd * (C1T_SLAB / 16) + kh), b_const | (b16 + (kd * 3 + kh) * 256), idesc,
                                     (kd | kh) ? 1u : 0u);
                    }
                    umma_commit(&tmem_full[b]);
                }
                umma_commit(&empty_bar[s]);
            }
        }
    } else {
        const int ew = warp - 2, q = warp & 3, grp = ew >> 2;      // TMEM lane quarter = warp % 4
        float s1[16], s2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
        uint32_t blk = 0;
        for (int u = first; u < p.num_units; u += step) {
            const int tile = u / p.groups, g = u - tile * p.groups;
            const int nb = (p.nblk - g * 8) < 8 ? (p.nblk - g * 8) : 8;
            // this thread's row of the padded matrix -> (n, d, h) of the output it produces
            unsigned r = (unsigned)tile * 128u + (unsigned)(q * 32 + lane);
            const int h = (int)(r % (unsigned)p.Hp); r /= (unsigned)p.Hp;
            const int d = (int)(r % (unsigned)p.Dp);
            const int n = (int)(r / (unsigned)p.Dp);
            const bool row_ok = h < p.H && d < p.D && n < p.N;
            __nv_bfloat16* out_row = p.y + (((long long)n * p.D + d) * p.H + h) * p.W * 16;
            for (int jj = 0; jj < nb; ++jj, ++blk) {
                if ((blk & 1u) != (uint32_t)grp) continue;
                const uint32_t b = blk & 3;
                const int w0 = (g * 8 + jj) * 8;
                mbar_wait(&tmem_full[b], (blk >> 2) & 1);
                tcgen05_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 128;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r4[4][16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) tmem_ld_32x32b_x16(taddr + (half * 4 + i) * 16, r4[i]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int w = w0 + half * 4 + i;
                        if (row_ok && w < p.W) {
                            uint32_t pk[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) pk[e] = pack_bf16x2(__uint_as_float(r4[i][2 * e]), __uint_as_float(r4[i][2 * e + 1]));
                            st_global_256(out_row + w * 16, pk);        // one voxel = 16 channels = one 32-byte sector
#pragma unroll
                            for (int e = 0; e < 8; ++e) {                  // statistics of the stored (rounded) values
                                const float2 f = unpack_bf16x2(pk[e]);
                                s1[2 * e] += f.x; s2[2 * e] = fmaf(f.x, f.x, s2[2 * e]);
                                s1[2 * e + 1] += f.y; s2[2 * e + 1] = fmaf(f.y, f.y, s2[2 * e + 1]);
                            }
                        }
                    }
                }
                tcgen05_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[b]);
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float sa = warp_sum(s1[c]), sb = warp_sum(s2[c]);
            if (lane == 0) { atomicAdd(&red[c], sa); atomicAdd(&red[16 + c], sb); }
        }
    }
    tcgen05_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) atomicAdd(&p.stats[threadIdx.x], red[threadIdx.x]);
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static inline void c1t_geometry(int D, int H, int W, int& Dp, int& Hp, int& nblk, int& Wq) {
    Dp = D + 2; Hp = H + 2; nblk = (W + 7) / 8; Wq = 8 * (nblk + 1);
}

}  // namespace coinn

// rows and row length (elements) of the padded bf16 input matrix for an [N, D, H, W] volume
COINN_API int coinn_conv1_padded_shape(int N, int D, int H, int W, long long* rows, int* cols) {
    int Dp, Hp, nblk, Wq;
    coinn::c1t_geometry(D, H, W, Dp, Hp, nblk, Wq);
    *rows = (long long)N * Dp * Hp;
    *cols = Wq;
    return 0;
}

// x: [N,D,H,W] fp32 (x_dtype 0) or bf16 (1)  ->  xp: [N*(D+2)*(H+2), Wq] bf16, zero halo
COINN_API int coinn_conv1_pad_input(const void* x, int x_dtype, void* xp, int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    int Dp, Hp, nblk, Wq;
    c1t_geometry(D, H, W, Dp, Hp, nblk, Wq);
    const int chunks = Wq / 8;
    const long long total = (long long)N * Dp * Hp * chunks;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_dtype == 0)
        conv1_pad_input_kernel<float><<<(unsigned)blocks, threads, 0, st>>>((const float*)x, (__nv_bfloat16*)xp, N, D, H, W, Dp, Hp, chunks, total);
    else
        conv1_pad_input_kernel<__nv_bfloat16><<<(unsigned)blocks, threads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)xp, N, D, H, W, Dp, Hp, chunks, total);
    COINN_CHECK_LAUNCH();
    return 0;
}

// xp: padded input (coinn_conv1_pad_input); w: [16,27] fp32; y: [N,D,H,W,16] bf16; stats: 32 floats (zeroed)
COINN_API int coinn_conv1_fwd_toeplitz(const void* xp, const float* w, void* y, float* stats, int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    C1TParams p;
    int Wq;
    c1t_geometry(D, H, W, p.Dp, p.Hp, p.nblk, Wq);
    p.y = reinterpret_cast<__nv_bfloat16*>(y); p.stats = stats; p.w = w;
    p.N = N; p.D = D; p.H = H; p.W = W;
    p.groups = (p.nblk + 7) / 8;
    const long long rows = (long long)N * p.Dp * p.Hp;
    if (rows + 3LL * p.Hp + 256 >= (1LL << 31)) return -1;
    const int tiles = (int)((rows + 127) / 128);
    p.num_units = tiles * p.groups;
    p.stages = 3;
    const int smem_bytes = (int)C1T_B_BYTES + p.stages * (int)C1T_STAGE + 1024 + 1024;
    CUtensorMap tx;
    if (make_tmap_2d_bf16(&tx, xp, (uint64_t)rows, (uint64_t)Wq, (uint64_t)Wq * 2, C1T_ROWS, 8, CU_TENSOR_MAP_SWIZZLE_NONE) != 0) return -3;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv1_toeplitz_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int grid = p.num_units < B200_SM_COUNT ? p.num_units : B200_SM_COUNT;
    conv1_toeplitz_fwd_kernel<<<grid, C1T_THREADS, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tx, p);
    COINN_CHECK_LAUNCH();
    return 0;
}
