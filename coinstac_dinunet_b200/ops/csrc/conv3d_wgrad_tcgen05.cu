// Weight gradient of the 3x3x3 Conv3d as an implicit GEMM on tcgen05:
//
//   dWt[(tap,ci), co] += sum_m  x[m + off(tap), ci] * dy[m, co]          (reduction over the voxels m)
//
// The reduction dimension is the voxel index, so BOTH operands are "MN-major" for the tensor core: for a
// fixed voxel the (tap,ci) values / the co values are contiguous.  Per 64-voxel k-block
//   A  [64 voxels x 128 (tap,ci)]  per M-tile: gathered with cp.async (zero fill outside the volume) into
//                                  the 128B-swizzled MN-major layout (two 64-element blocks, LBO = 8 KB)
//   B  [64 voxels x COUT]          one TMA 2-D load per 64 channels straight out of dy (64B swizzle for 32)
//   D  TM accumulators [128 x COUT] fp32 in TMEM (TM*COUT <= 512 columns)
// The voxel range is split over gridDim.y CTAs (split-K); partial sums are merged with fp32 atomics.
// grid.x enumerates groups of TM consecutive M-tiles of the (tap,ci) axis.
#include "umma.cuh"

namespace coinn {

constexpr int WG_BK = 64;                 // voxels per k-block
constexpr int WG_THREADS = 320;
constexpr int WG_LAG = 1;

struct WgradParams {
    const __nv_bfloat16* x;     // [M_total, CIN]
    float* dwt;                 // [27*CIN, COUT] fp32, zero-initialised by the caller
    int N, D, H, W;
    long long m_total;
    long long range_len;        // voxels per grid.y slice (multiple of 64)
};

template <int COUT, int TM> struct WgradCfg {
    static constexpr int A_TILE = WG_BK * 128 * 2;                   // 16 KB per M-tile
    static constexpr int A_BYTES = TM * A_TILE;
    static constexpr int B_BYTES = WG_BK * COUT * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    __host__ __device__ static constexpr int stages() { return 3; }
    __host__ __device__ static constexpr int smem_bytes() { return stages() * STAGE_BYTES + 1024 + 256; }
};

__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
}

template <int CIN, int COUT, int TM>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv3d_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const WgradParams p) {
    using Cfg = WgradCfg<COUT, TM>;
    constexpr int STAGES = Cfg::stages();
    constexpr int ROWS = 27 * CIN;                                 // valid rows of the (tap,ci) axis
    constexpr uint32_t TMEM_COLS = (TM * COUT) <= 32 ? 32 : ((TM * COUT) <= 64 ? 64 : ((TM * COUT) <= 128 ? 128 :
                                   ((TM * COUT) <= 256 ? 256 : 512)));
    static_assert(TM * COUT <= 512, "accumulators exceed TMEM");
    constexpr bool B_SW64 = (COUT == 32);
    constexpr int B_BLOCKS = B_SW64 ? 1 : COUT / 64;               // 64-channel TMA boxes per k-block
    static_assert(COUT == 32 || COUT % 64 == 0, "COUT must be 32 or a multiple of 64");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt0 = blockIdx.x * TM;                                // first M-tile of this CTA
    const long long m_begin = (long long)blockIdx.y * p.range_len;
    const long long m_end = min(p.m_total, m_begin + p.range_len);
    const int num_kb = m_end > m_begin ? (int)((m_end - m_begin + WG_BK - 1) / WG_BK) : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_dy);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 128 + 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ dy TMA producer ============================
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
                if (leader) mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES);
                uint8_t* b_dst = smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES;
                const int row0 = (int)(m_begin + (long long)kb * WG_BK);
#pragma unroll
                for (int j = 0; j < B_BLOCKS; ++j)
                    if (leader) tma_load_2d(b_dst + j * (WG_BK * 128), &tmap_dy, &full_bar[s], j * 64, row0);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            constexpr uint32_t idesc = make_idesc_f16(128, COUT, 1, 1, 1);      // A and B MN-major
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                tcgen05_after_sync();
                const uint32_t a_base = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
                for (int mt = 0; mt < TM; ++mt) {
#pragma unroll
                    for (int k = 0; k < WG_BK / 16; ++k) {
                        // 16 voxel rows per MMA: 16 * 128 B (A, B with 128B rows) or 16 * 64 B (B with 64B rows)
                        const uint64_t adesc = make_smem_desc(a_base + mt * Cfg::A_TILE + k * 2048, 8192, 1024, SMEM_LAYOUT_SW128);
                        const uint64_t bdesc = B_SW64 ? make_smem_desc(b_base + k * 1024, 4096, 512, SMEM_LAYOUT_SW64)
                                                      : make_smem_desc(b_base + k * 2048, 8192, 1024, SMEM_LAYOUT_SW128);
                        if (leader) umma_f16(tmem_base + mt * COUT, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                }
                if (leader) umma_commit(&empty_bar[s]);
            }
            if (leader) umma_commit(done_bar);
        }
    } else if (warp < 6) {
        // ================================= epilogue =================================
        const int q = warp & 3;
        if (num_kb > 0) {
            mbar_wait(done_bar, 0);
            tcgen05_after_sync();
#pragma unroll 1
            for (int mt = 0; mt < TM; ++mt) {
                const int row = (mt0 + mt) * 128 + q * 32 + lane;            // (tap, ci) index
#pragma unroll 1
                for (int c = 0; c < COUT; c += 16) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + mt * COUT + c, r);
                    tmem_ld_wait();
                    if (row < ROWS) {
                        float* dst = p.dwt + (size_t)row * COUT + c;
#pragma unroll
                        for (int j = 0; j < 16; ++j) atomicAdd(dst + j, __uint_as_float(r[j]));
                    }
                }
            }
        }
    } else {
        // ============================ im2col gather producers ============================
        const int t = threadIdx.x - 6 * 32;
        const int krow = t & 63;                                     // voxel row inside the k-block
        const int half = t >> 6;                                     // which 64-element MN block of every M-tile
        const uint32_t row_off = (uint32_t)krow * 128u + (uint32_t)half * 8192u;
        const uint32_t sw = (uint32_t)(krow & 7);
        const long long HW = (long long)p.H * p.W;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const long long m = m_begin + (long long)kb * WG_BK + krow;
            const bool row_ok = m < m_end;
            int w = 0, h = 0, d = 0;
            if (row_ok) {
                long long tq = m;
                w = (int)(tq % p.W); tq /= p.W;
                h = (int)(tq % p.H); tq /= p.H;
                d = (int)(tq % p.D);
            }
            const uint32_t vd = (d > 0 ? 1u : 0u) | 2u | (d + 1 < p.D ? 4u : 0u);
            const uint32_t vh = (h > 0 ? 1u : 0u) | 2u | (h + 1 < p.H ? 4u : 0u);
            const uint32_t vw = (w > 0 ? 1u : 0u) | 2u | (w + 1 < p.W ? 4u : 0u);
            const __nv_bfloat16* center = p.x + (row_ok ? m : 0) * CIN;

            mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
            const uint32_t a_dst = smem_u32(smem + s * Cfg::STAGE_BYTES) + row_off;
#pragma unroll 1
            for (int mt = 0; mt < TM; ++mt) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = (mt0 + mt) * 128 + half * 64 + c * 8;   // first (tap,ci) element of the chunk
                    const int tap = j / CIN, ci0 = j % CIN;
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    const bool ok = row_ok && tap < 27 && ((vd >> kd) & 1u) && ((vh >> kh) & 1u) && ((vw >> kw) & 1u);
                    const long long voff = (long long)(kd - 1) * HW + (long long)(kh - 1) * p.W + (kw - 1);
                    const __nv_bfloat16* src = ok ? center + voff * CIN + ci0 : p.x;
                    cp16z(a_dst + mt * Cfg::A_TILE + (((uint32_t)c ^ sw) << 4), src, ok);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (kb >= WG_LAG) {
                asm volatile("cp.async.wait_group %0;" :: "n"(WG_LAG) : "memory");
                fence_proxy_async_smem();
                mbar_arrive(&full_bar[(kb - WG_LAG) % STAGES]);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async_smem();
        const int pending = num_kb < WG_LAG ? num_kb : WG_LAG;
        for (int j = num_kb - pending; j < num_kb; ++j) mbar_arrive(&full_bar[j % STAGES]);
    }

    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN, int COUT, int TM>
static int launch_wgrad(const void* x, const void* dy, float* dwt, int N, int D, int H, int W, cudaStream_t st) {
    using Cfg = WgradCfg<COUT, TM>;
    const long long m_total = (long long)N * D * H * W;
    CUtensorMap tdy;
    const bool sw64 = (COUT == 32);
    if (make_tmap_2d_bf16(&tdy, dy, (uint64_t)m_total, (uint64_t)COUT, (uint64_t)COUT * 2, WG_BK, sw64 ? 32 : 64,
                          sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -2;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_kernel<CIN, COUT, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::smem_bytes());
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    constexpr int m_tiles = (27 * CIN + 127) / 128;
    constexpr int groups = (m_tiles + TM - 1) / TM;
    const long long kblocks = (m_total + WG_BK - 1) / WG_BK;
    long long splits = B200_SM_COUNT / groups;
    if (splits < 1) splits = 1;
    if (splits > kblocks) splits = kblocks;
    WgradParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.dwt = dwt;
    p.N = N; p.D = D; p.H = H; p.W = W;
    p.m_total = m_total;
    p.range_len = ((kblocks + splits - 1) / splits) * WG_BK;
    splits = (m_total + p.range_len - 1) / p.range_len;
    dim3 grid(groups, (unsigned)splits);
    conv3d_wgrad_kernel<CIN, COUT, TM><<<grid, WG_THREADS, Cfg::smem_bytes(), st>>>(tdy, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// x: [N,D,H,W,cin] bf16, dy: [N,D,H,W,cout] bf16, dwt: [27*cin, cout] fp32 (zeroed), dwt[(tap*cin+ci), co].
COINN_API int coinn_conv3d_wgrad(const void* x, const void* dy, float* dwt, int N, int D, int H, int W, int cin, int cout,
                                 void* stream) {
    using namespace coinn;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cin == 16 && cout == 32) return launch_wgrad<16, 32, 4>(x, dy, dwt, N, D, H, W, st);
    if (cin == 32 && cout == 64) return launch_wgrad<32, 64, 4>(x, dy, dwt, N, D, H, W, st);
    if (cin == 64 && cout == 128) return launch_wgrad<64, 128, 3>(x, dy, dwt, N, D, H, W, st);
    if (cin == 128 && cout == 256) return launch_wgrad<128, 256, 2>(x, dy, dwt, N, D, H, W, st);
    if (cin == 32 && cout == 32) return launch_wgrad<32, 32, 4>(x, dy, dwt, N, D, H, W, st);
    if (cin == 64 && cout == 64) return launch_wgrad<64, 64, 4>(x, dy, dwt, N, D, H, W, st);
    return -1;
}
