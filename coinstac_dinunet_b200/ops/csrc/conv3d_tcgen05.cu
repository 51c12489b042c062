// 3x3x3 / pad 1 / stride 1 Conv3d as an implicit GEMM on tcgen05 tensor cores.
//
//   y[m, co] = sum_{tap, ci} x[m + off(tap), ci] * Wk[co, tap*CIN + ci]       m = linear voxel index of NDHWC
//
// GEMM view: M = N*D*H*W voxels (tiles of 128), N = COUT (one tile: 16..256), K = 27*CIN (k-blocks of 64 bf16).
// The same kernel computes dgrad: feed dy as the input and the tap-flipped, (ci,co)-swapped weights.
//
// Persistent CTAs (one per SM) loop over voxel tiles; 10 warps with fixed roles:
//   warp 0      TMA producer of the weight k-slices (cp.async.bulk.tensor.2d, 128B swizzle)
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16 (128 x COUT x 16), accumulators in TMEM,
//               double buffered (2 x COUT columns) so the epilogue of tile t overlaps the MMAs of tile t+1
//   warps 2-5   epilogue: tcgen05.ld -> bf16 -> global (channels-last rows are contiguous)
//   warps 6-9   im2col gather producers: thread r owns tile row r; per k-slice it copies the 8 16-byte chunks
//               of its row (taps x channels, zero-filled outside the volume) with cp.async into the
//               128B-swizzled K-major layout the UMMA descriptor expects, then fence.proxy.async + mbarrier
//               arrive so the tensor core (async proxy) sees the data.  No im2col buffer ever touches HBM.
#include "umma.cuh"

namespace coinn {

constexpr int CONV_BM = 128;
constexpr int CONV_BK = 64;
constexpr int CONV_THREADS = 320;
constexpr int CONV_GATHER_LAG = 2;        // cp.async groups kept in flight per gather thread

struct ConvParams {
    const __nv_bfloat16* x;     // [M_total, CIN]
    __nv_bfloat16* y;           // [M_total, COUT]
    int N, D, H, W;
    long long m_total;
    int num_tiles;
};

template <int COUT> struct ConvCfg {
    static constexpr int A_BYTES = CONV_BM * CONV_BK * 2;                 // 16 KB
    static constexpr int B_BYTES = COUT * CONV_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    __host__ __device__ static constexpr int stages() { return COUT >= 256 ? 4 : (COUT >= 128 ? 6 : 8); }
    __host__ __device__ static constexpr int smem_bytes() { return stages() * STAGE_BYTES + 1024 + 512; }
};

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int CIN, int COUT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const ConvParams p) {
    using Cfg = ConvCfg<COUT>;
    constexpr int STAGES = Cfg::stages();
    constexpr int KB = (27 * CIN + CONV_BK - 1) / CONV_BK;       // k-blocks per tile
    constexpr int CHUNKS_PER_TAP = CIN / 8;                       // 16-byte chunks per tap
    constexpr uint32_t TMEM_COLS = (2 * COUT) < 32 ? 32 : 2 * COUT;
    static_assert(CIN % 8 == 0 && (CHUNKS_PER_TAP & (CHUNKS_PER_TAP - 1)) == 0, "CIN must be 8 * 2^k");
    static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "COUT must be a valid UMMA N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;        // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 128 + 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int first_tile = blockIdx.x, tile_step = gridDim.x;

    if (warp == 0) {
        // ============================ weight TMA producer ============================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES);
                    tma_load_2d(smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES, &tmap_w, &full_bar[s], kb * CONV_BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(CONV_BM, COUT, 1, 0, 0);
            uint32_t it = 0, t = 0;
            for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++t) {
                const uint32_t a = t & 1;
                mbar_wait(&tmem_empty[a], ((t >> 1) & 1) ^ 1);          // epilogue drained this accumulator
                tcgen05_after_sync();
                const uint32_t d_tmem = tmem_base + a * COUT;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tcgen05_after_sync();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < CONV_BK / 16; ++k) {
                        umma_f16(d_tmem, make_smem_desc(a_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128),
                                 make_smem_desc(b_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128), idesc,
                                 (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[a]);
            }
        }
    } else if (warp < 6) {
        // ================================= epilogue =================================
        const int q = warp & 3;
        uint32_t t = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++t) {
            const uint32_t a = t & 1;
            mbar_wait(&tmem_full[a], (t >> 1) & 1);
            tcgen05_after_sync();
            const long long m = (long long)tile * CONV_BM + q * 32 + lane;
            __nv_bfloat16* out = p.y + m * COUT;
#pragma unroll 1
            for (int c = 0; c < COUT; c += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + a * COUT + c, r);
                tmem_ld_wait();
                if (m < p.m_total) {
                    uint4 lo = make_uint4(pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1])),
                                          pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3])),
                                          pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5])),
                                          pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7])));
                    uint4 hi = make_uint4(pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9])),
                                          pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11])),
                                          pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13])),
                                          pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15])));
                    *reinterpret_cast<uint4*>(out + c) = lo;
                    *reinterpret_cast<uint4*>(out + c + 8) = hi;
                }
            }
            tcgen05_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
        }
    } else {
        // ============================ im2col gather producers ============================
        const int r = threadIdx.x - 6 * 32;                        // tile row owned by this thread
        const uint32_t row_off = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        const long long HW = (long long)p.H * p.W;
        uint32_t it = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
            const long long m = (long long)tile * CONV_BM + r;
            const bool row_ok = m < p.m_total;
            int w = 0, h = 0, d = 0;
            if (row_ok) {
                long long tq = m;
                w = (int)(tq % p.W); tq /= p.W;
                h = (int)(tq % p.H); tq /= p.H;
                d = (int)(tq % p.D);
            }
            // per-axis validity of the -1 / 0 / +1 neighbours, bit k set <=> offset (k-1) stays inside
            const uint32_t vd = (d > 0 ? 1u : 0u) | 2u | (d + 1 < p.D ? 4u : 0u);
            const uint32_t vh = (h > 0 ? 1u : 0u) | 2u | (h + 1 < p.H ? 4u : 0u);
            const uint32_t vw = (w > 0 ? 1u : 0u) | 2u | (w + 1 < p.W ? 4u : 0u);
            const __nv_bfloat16* center = p.x + (row_ok ? m : 0) * CIN;

            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                const uint32_t a_dst = smem_u32(smem + s * Cfg::STAGE_BYTES) + row_off;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int qidx = kb * 8 + c;                       // 16-byte chunk index along K
                    const int tap = qidx / CHUNKS_PER_TAP;
                    const int ci0 = (qidx % CHUNKS_PER_TAP) * 8;
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    const bool ok = row_ok && tap < 27 && ((vd >> kd) & 1u) && ((vh >> kh) & 1u) && ((vw >> kw) & 1u);
                    const long long voff = (long long)(kd - 1) * HW + (long long)(kh - 1) * p.W + (kw - 1);
                    const __nv_bfloat16* src = ok ? center + voff * CIN + ci0 : p.x;
                    cp_async_16_zfill(a_dst + (((uint32_t)c ^ sw) << 4), src, ok);
                }
                cp_async_commit();
                if (it >= (uint32_t)CONV_GATHER_LAG) {                 // publish the slice issued LAG iterations ago
                    cp_async_wait<CONV_GATHER_LAG>();
                    fence_proxy_async_smem();
                    mbar_arrive(&full_bar[(it - CONV_GATHER_LAG) % STAGES]);
                }
            }
        }
        // drain: publish the last LAG slices
        cp_async_wait<0>();
        fence_proxy_async_smem();
        const uint32_t pending = it < (uint32_t)CONV_GATHER_LAG ? it : (uint32_t)CONV_GATHER_LAG;
        for (uint32_t j = it - pending; j < it; ++j) mbar_arrive(&full_bar[j % STAGES]);
    }

    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN, int COUT>
static int launch_conv(const void* x, const void* wk, void* y, int N, int D, int H, int W, int kpad, cudaStream_t st) {
    using Cfg = ConvCfg<COUT>;
    CUtensorMap tw;
    if (make_tmap_2d_bf16(&tw, wk, (uint64_t)COUT, (uint64_t)kpad, (uint64_t)kpad * 2, COUT, CONV_BK) != 0) return -2;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_igemm_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::smem_bytes());
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    ConvParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.N = N; p.D = D; p.H = H; p.W = W;
    p.m_total = (long long)N * D * H * W;
    p.num_tiles = (int)((p.m_total + CONV_BM - 1) / CONV_BM);
    const int grid = p.num_tiles < B200_SM_COUNT ? p.num_tiles : B200_SM_COUNT;
    conv3d_igemm_kernel<CIN, COUT><<<grid, CONV_THREADS, Cfg::smem_bytes(), st>>>(tw, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// x: [N,D,H,W,cin] bf16, wk: [cout, kpad] bf16 with k = tap*cin + ci (kpad = 27*cin rounded up to 64, zero padded),
// y: [N,D,H,W,cout] bf16.  Returns -1 for unsupported channel combinations.
COINN_API int coinn_conv3d_igemm(const void* x, const void* wk, void* y, int N, int D, int H, int W, int cin, int cout,
                                 int kpad, void* stream) {
    using namespace coinn;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (kpad != ((27 * cin + 63) / 64) * 64) return (int)cudaErrorInvalidValue;
#define CASE(CI, CO) if (cin == CI && cout == CO) return launch_conv<CI, CO>(x, wk, y, N, D, H, W, kpad, st);
    CASE(16, 32) CASE(32, 64) CASE(64, 128) CASE(128, 256)          // fprop of blocks 2..5
    CASE(32, 16) CASE(64, 32) CASE(128, 64) CASE(256, 128)          // dgrad of blocks 2..5
    CASE(16, 16) CASE(32, 32) CASE(64, 64) CASE(128, 128)           // square (tests / other nets)
#undef CASE
    return -1;
}
