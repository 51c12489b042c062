// Conv3d weight gradient from the zero-padded TMA halo (companion of conv3d_halo.cu).
//
//   dWt[(tap,ci), co] = sum_p x[p + off(tap), ci] * dy[p, co]
//
// A tile is TH complete rows of one (n, d) plane in the padded, flattened pixel order (width Wp = W + 2).  The three
// halo planes of x (whole pixels, CIN*2 bytes each, 32B/64B swizzle) and the dy tile (box width Wp, so the two
// padding columns of every row arrive as zeros) are four TMA loads.  For the filter row (kd, kh) the MN-major A
// operand is the halo plane kd shifted by kh*Wp pixels; its MN blocks (one pixel's CIN channels) are ONE PIXEL apart
// (LBO = pixel size), i.e. block j is the view shifted by j pixels: blocks 0..2 are the taps kw = 0..2, the
// remaining blocks of the 128-row MMA are harmless extra shifts that are simply not written back.  So the whole
// weight gradient is 9 accumulators x 8 MMAs (K = 16 pixels each) per tile, reading every input voxel from L2
// once instead of 27 times (the gather kernel conv3d_wgrad_tcgen05.cu is bound by exactly that traffic).
// Accumulators stay in TMEM across all tiles of the persistent CTA; fp32 atomics merge the CTAs at the end.
#include "umma.cuh"
#include <cstdlib>

namespace coinn {

constexpr int WH_THREADS = 192;

struct WgradHaloParams {
    float* dwt;                 // [27*CIN, COUT] fp32 (zeroed by the caller)
    int N, D, H, W;
    int TH, Wp, tiles_h, num_tiles;
    uint32_t x_region;          // bytes reserved per halo plane (1 KB multiple)
    uint32_t x_tx, dy_tx;       // exact bytes delivered by one x plane box / the dy box
    int stages;
    int mt_begin, mt_count;     // unused (grid.x selects the M-tile group)
    int m64;                    // CIN == 16 only: 64-row MMAs (4 shift blocks instead of 8: half the A-operand smem traffic)
};

__device__ __forceinline__ void tma5(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// MT_MAX: accumulators (filter rows (kd,kh)) per CTA; gridDim.x groups cover all 9
template <int CIN, int COUT, int MT_MAX>
__global__ void __launch_bounds__(WH_THREADS, 1)
conv3d_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WgradHaloParams p) {
    constexpr int PIXB = CIN * 2;                                   // bytes per pixel of x
    constexpr int DYB = COUT * 2;                                   // bytes per pixel of dy
    constexpr uint64_t A_LAYOUT = CIN == 16 ? SMEM_LAYOUT_SW32 : SMEM_LAYOUT_SW64;
    constexpr uint64_t B_LAYOUT = COUT == 32 ? SMEM_LAYOUT_SW64 : SMEM_LAYOUT_SW128;
    constexpr uint32_t DY_REGION = 128 * DYB;                       // 128 pixel rows
    constexpr uint32_t TMEM_COLS = (MT_MAX * COUT) <= 128 ? 128 : ((MT_MAX * COUT) <= 256 ? 256 : 512);
    static_assert(MT_MAX * COUT <= 512, "accumulators exceed TMEM");
    constexpr int MAX_STAGES = 8;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = 3 * p.x_region + 1024 + DY_REGION;  // +1 KB slack: shifted views run past plane 2
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* done_bar = bars + 2 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = p.stages;
    const int mt0 = blockIdx.x * MT_MAX;                             // first filter row (kd*3+kh) of this CTA
    const int mts = min(MT_MAX, 9 - mt0);

    // everything the MMAs may touch beyond the TMA boxes (padding, slack, tail pixel rows of dy) must be finite: zero it
    for (uint32_t i = threadIdx.x; i < (uint32_t)p.stages * stage_bytes / 16; i += WH_THREADS)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_dy);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_proxy_async_smem();
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int first = blockIdx.y, step = gridDim.y;
    const int my_tiles = first < p.num_tiles ? (p.num_tiles - first + step - 1) / step : 0;

    if (warp == 0) {
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            uint32_t it = 0;
            for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                const int plane = tile / p.tiles_h, h0 = (tile % p.tiles_h) * p.TH;
                const int n = plane / p.D, d = plane % p.D;
                const int s = it % STAGES;
                mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                uint8_t* dst = smem + (size_t)s * stage_bytes;
                if (leader) mbar_arrive_expect_tx(&full_bar[s], 3 * p.x_tx + p.dy_tx);
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) if (leader) tma5(dst + kd * p.x_region, &tmap_x, &full_bar[s], 0, -1, h0 - 1, d + kd - 1, n);
                if (leader) tma5(dst + 3 * p.x_region + 1024, &tmap_dy, &full_bar[s], 0, 0, h0, d, n);
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            // both operands MN-major.  With CIN = 16 an M = 128 MMA carries 8 pixel-shift blocks of which only 3 (kw) are
            // used; M = 64 carries 4.  The kernel is bound by the shared-memory reads of the A operand, so M = 64 wins.
            const uint32_t idesc = make_idesc_f16(p.m64 ? 64 : 128, COUT, 1, 1, 1);
            uint32_t it = 0;
            for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full_bar[s], (it / STAGES) & 1);
                tcgen05_after_sync();
                const uint32_t xs = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t dys = xs + 3 * p.x_region + 1024;
                // constant descriptor parts hoisted; per-MMA work is two 32-bit adds (see conv3d_halo.cu)
                const uint64_t a_const = make_smem_desc(0, PIXB, 8 * PIXB, A_LAYOUT);
                const uint64_t b_const = make_smem_desc(0, 8192, 8 * DYB, B_LAYOUT);
                const uint32_t xs16 = (xs & 0x3FFFFu) >> 4, dys16 = (dys & 0x3FFFFu) >> 4;
                const uint32_t region16 = p.x_region >> 4, row16 = (uint32_t)p.Wp * (PIXB / 16);
                int kd = mt0 / 3, kh = mt0 % 3;
#pragma unroll 1
                for (int m = 0; m < mts; ++m) {
                    const uint32_t a0 = xs16 + kd * region16 + kh * row16;
#pragma unroll
                    for (int k = 0; k < 8; ++k)                                      // 16 pixels per MMA
                        if (leader) umma_f16(tmem_base + m * COUT, a_const | (a0 + k * PIXB), b_const | (dys16 + k * DYB), idesc,
                                 (it > 0 || k > 0) ? 1u : 0u);
                    if (++kh == 3) { kh = 0; ++kd; }
                }
                if (leader) umma_commit(&empty_bar[s]);
            }
            if (leader) umma_commit(done_bar);
        }
    } else {
        const int q = warp & 3;
        // accumulator row = (shift j, channel ci).  M = 128: row r sits in TMEM lane r.  M = 64 (cta_group::1): rows
        // 16q .. 16q+15 sit in lanes 32q .. 32q+15, i.e. warp q holds shift j = q in its first 16 lanes.
        const int l = p.m64 ? (lane < 16 ? q * 16 + lane : 3 * CIN) : q * 32 + lane;
        const int j = l / CIN, ci = l % CIN;
        if (my_tiles > 0) {
            mbar_wait(done_bar, 0);
            tcgen05_after_sync();
#pragma unroll 1
            for (int m = 0; m < mts; ++m) {
                const int fr = mt0 + m;
                const int tap = fr * 3 + j;                           // (kd*3 + kh)*3 + kw
#pragma unroll 1
                for (int c = 0; c < COUT; c += 16) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + m * COUT + c, r);
                    tmem_ld_wait();
                    if (j < 3) {
                        float* dst = p.dwt + ((size_t)tap * CIN + ci) * COUT + c;
#pragma unroll
                        for (int e = 0; e < 16; ++e) atomicAdd(dst + e, __uint_as_float(r[e]));
                    }
                }
            }
        }
    }
    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN, int COUT, int MT_MAX>
static int launch_wgrad_halo(const void* x, const void* dy, float* dwt, int N, int D, int H, int W, cudaStream_t st) {
    WgradHaloParams p;
    static int m64_env = -1;
    if (m64_env < 0) { const char* e = getenv("COINN_WGRAD_M64"); m64_env = e ? atoi(e) : 1; }
    p.m64 = (CIN == 16 && m64_env) ? 1 : 0;
    p.dwt = dwt; p.N = N; p.D = D; p.H = H; p.W = W;
    p.Wp = W + 2;
    if (p.Wp > 128) return -1;
    p.TH = 128 / p.Wp;
    if (p.TH > H) p.TH = H;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.num_tiles = N * D * p.tiles_h;
    p.x_tx = (uint32_t)(p.TH + 2) * p.Wp * CIN * 2;
    p.x_region = (p.x_tx + 1023u) & ~1023u;
    p.dy_tx = (uint32_t)p.TH * p.Wp * COUT * 2;
    p.mt_begin = 0; p.mt_count = 9;
    const uint32_t stage_bytes = 3 * p.x_region + 1024 + 128 * COUT * 2;
    int stages = (int)((212 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return -1;
    p.stages = stages;
    const int smem_bytes = stages * (int)stage_bytes + 1024 + 256;

    auto enc = get_tensor_map_encoder();
    if (!enc) return -2;
    CUtensorMap tx, tdy;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    {
        cuuint64_t dims[5] = {(cuuint64_t)CIN, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)CIN * 2, (cuuint64_t)W * CIN * 2, (cuuint64_t)H * W * CIN * 2, (cuuint64_t)D * H * W * CIN * 2};
        cuuint32_t box[5] = {(cuuint32_t)CIN, (cuuint32_t)p.Wp, (cuuint32_t)(p.TH + 2), 1, 1};
        if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CIN == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
    }
    {
        cuuint64_t dims[5] = {(cuuint64_t)COUT, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)COUT * 2, (cuuint64_t)W * COUT * 2, (cuuint64_t)H * W * COUT * 2, (cuuint64_t)D * H * W * COUT * 2};
        cuuint32_t box[5] = {(cuuint32_t)COUT, (cuuint32_t)p.Wp, (cuuint32_t)p.TH, 1, 1};
        if (enc(&tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                COUT == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -4;
    }
    static int configured = 0;
    if (configured < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_halo_kernel<CIN, COUT, MT_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = smem_bytes;
    }
    const int groups = (9 + MT_MAX - 1) / MT_MAX;
    int splits = B200_SM_COUNT / groups;
    if (splits > p.num_tiles) splits = p.num_tiles;
    if (splits < 1) splits = 1;
    dim3 grid(groups, splits);
    conv3d_wgrad_halo_kernel<CIN, COUT, MT_MAX><<<grid, WH_THREADS, smem_bytes, st>>>(tx, tdy, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// same contract as coinn_conv3d_wgrad; -1 when the shape is not covered
COINN_API int coinn_conv3d_wgrad_halo(const void* x, const void* dy, float* dwt, int N, int D, int H, int W, int cin, int cout,
                                      void* stream) {
    using namespace coinn;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cin == 16 && cout == 32) return launch_wgrad_halo<16, 32, 9>(x, dy, dwt, N, D, H, W, st);
    if (cin == 32 && cout == 64) return launch_wgrad_halo<32, 64, 5>(x, dy, dwt, N, D, H, W, st);
    if (cin == 32 && cout == 32) return launch_wgrad_halo<32, 32, 9>(x, dy, dwt, N, D, H, W, st);
    return -1;
}
