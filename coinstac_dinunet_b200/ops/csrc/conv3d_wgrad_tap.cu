// Conv3d weight gradient for the small-volume / many-channel blocks (C_in = 64, 128): one GEMM per filter tap.
//
//   dWt[(tap, ci), co] = sum_p x[p + off(tap), ci] * dy[p, co]
//
// In blocks 4 and 5 of VBMNet the volume has 32 k / 3.5 k voxels and 64->128 / 128->256 channels: the contraction
// (voxels) is short, the output (27 x C_in x C_out) is large.  The gather kernel (conv3d_wgrad_tcgen05.cu) spends 99 us
// and 72 us there for 14.5 and 6.2 GFLOP.  Here both operands of a tap are plain TMA boxes over the SAME voxel box
// {64 channels, W, Hb, Db} - the x box shifted by the tap offset, out-of-bounds voxels zero-filled by TMA (= the padding)
// - landing as [voxel rows][128 B] tiles with the 128B swizzle, i.e. directly the MN-major operands of
// tcgen05.mma (K = voxels).  A CTA owns two 64-row M blocks (two taps for C_in = 64, the two channel halves of one tap
// for C_in = 128) and a slice of the voxel boxes; its 128 x C_out accumulator stays in TMEM for the whole CTA and is
// merged with fp32 atomics at the end.  dy is re-read once per tap group from L2 (it is 2-8 MB).
#include "umma.cuh"

namespace coinn {

constexpr int WT_THREADS = 192;                 // warp 0: TMA, warp 1: MMA, warps 2-5: final epilogue
constexpr uint32_t WT_TILE = 128 * 128;         // one operand tile: [128 voxel rows][64 channels] bf16

struct WgradTapParams {
    float* dwt;                 // [27*CIN, COUT] fp32 (accumulated into)
    int N, D, H, W;
    int Hb, Db;                 // voxel box = W x Hb x Db (<= 128 voxels)
    int boxes_h, boxes_d, num_boxes;
    int ksteps;                 // ceil(W*Hb*Db / 16)
    uint32_t tile_tx;           // bytes one box delivers: W*Hb*Db*128
    int stages;
};

__device__ __forceinline__ void wt_tma5(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(WT_THREADS, 1)
conv3d_wgrad_tap_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WgradTapParams p) {
    constexpr int CI_CHUNKS = CIN / 64;                 // 64-channel chunks per tap
    constexpr int MBLOCKS = 27 * CI_CHUNKS;             // (tap, chunk) blocks of 64 accumulator rows
    constexpr int NB = COUT / 64;                       // dy tiles per stage
    constexpr uint32_t STAGE = (2 + NB) * WT_TILE;
    constexpr int MAX_STAGES = 4;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* done_bar = bars + 2 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = p.stages;
    // rows of a tile beyond the voxel box are never written by TMA: they must read as zero in the last K step
    for (uint32_t i = threadIdx.x; i < (uint32_t)STAGES * STAGE / 16; i += WT_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_dy);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, COUT);
    fence_proxy_async_smem();
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // the two M blocks of this CTA
    const int blk0 = blockIdx.x * 2, blk1 = blk0 + 1;
    const bool has1 = blk1 < MBLOCKS;
    const int first = blockIdx.y, step = gridDim.y;
    const int my_boxes = first < p.num_boxes ? (p.num_boxes - first + step - 1) / step : 0;

    if (warp == 0) {
        const bool leader = elect_one();
        uint32_t it = 0;
        for (int b = first; b < p.num_boxes; b += step, ++it) {
            int r = b;
            const int bh = r % p.boxes_h; r /= p.boxes_h;
            const int bd = r % p.boxes_d;
            const int n = r / p.boxes_d;
            const int h0 = bh * p.Hb, d0 = bd * p.Db;
            const int s = it % STAGES;
            mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            uint8_t* dst = smem + (size_t)s * STAGE;
            if (leader) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)((has1 ? 2 : 1) + NB) * p.tile_tx);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int blk = blk0 + j;
                if (blk < MBLOCKS) {
                    const int tap = blk / CI_CHUNKS, chunk = blk % CI_CHUNKS;
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    if (leader) wt_tma5(dst + j * WT_TILE, &tmap_x, &full_bar[s], chunk * 64, kw - 1, h0 + kh - 1, d0 + kd - 1, n);
                }
            }
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (leader) wt_tma5(dst + (2 + j) * WT_TILE, &tmap_dy, &full_bar[s], j * 64, 0, h0, d0, n);
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16(128, COUT, 1, 1, 1);                     // both operands MN-major
        const uint64_t a_const = make_smem_desc(0, WT_TILE, 1024, SMEM_LAYOUT_SW128);       // M blocks = the two x tiles
        const uint64_t b_const = make_smem_desc(0, WT_TILE, 1024, SMEM_LAYOUT_SW128);       // N blocks = the dy tiles
        uint32_t it = 0;
        for (int b = first; b < p.num_boxes; b += step, ++it) {
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            tcgen05_after_sync();
            const uint32_t a16 = (smem_u32(smem + (size_t)s * STAGE) & 0x3FFFFu) >> 4;
            const uint32_t b16 = a16 + 2 * (WT_TILE / 16);
            for (int k = 0; k < p.ksteps; ++k)                                              // 16 voxels per MMA
                if (leader) umma_f16(tmem_base, a_const | (a16 + k * 128), b_const | (b16 + k * 128), idesc, (it | (uint32_t)k) ? 1u : 0u);
            if (leader) umma_commit(&empty_bar[s]);
        }
        if (leader) umma_commit(done_bar);
    } else if (my_boxes > 0) {
        const int q = warp & 3;
        const int r = q * 32 + lane;                                  // accumulator row: M block r / 64, channel r % 64
        const int blk = blk0 + (r >> 6);
        mbar_wait(done_bar, 0);
        tcgen05_after_sync();
        if (blk < MBLOCKS) {
            const int tap = blk / CI_CHUNKS, chunk = blk % CI_CHUNKS;
            float* dst = p.dwt + ((size_t)tap * CIN + chunk * 64 + (r & 63)) * COUT;
#pragma unroll 1
            for (int c = 0; c < COUT; c += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) atomicAdd(dst + c + e, __uint_as_float(v[e]));
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < COUT; c += 16) { uint32_t v[16]; tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c, v); tmem_ld_wait(); }
        }
    }
    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, COUT);
}

template <int CIN, int COUT>
static int launch_wgrad_tap(const void* x, const void* dy, float* dwt, int N, int D, int H, int W, cudaStream_t st) {
    constexpr int NB = COUT / 64;
    constexpr uint32_t STAGE = (2 + NB) * WT_TILE;
    if (W > 128) return -1;
    WgradTapParams p;
    p.dwt = dwt; p.N = N; p.D = D; p.H = H; p.W = W;
    p.Hb = 128 / W; if (p.Hb > H) p.Hb = H;
    p.Db = 128 / (W * p.Hb); if (p.Db > D) p.Db = D; if (p.Db < 1) p.Db = 1;
    if (p.Hb > 256 || p.Db > 256) return -1;
    p.boxes_h = (H + p.Hb - 1) / p.Hb;
    p.boxes_d = (D + p.Db - 1) / p.Db;
    p.num_boxes = N * p.boxes_d * p.boxes_h;
    const int vox = W * p.Hb * p.Db;
    p.ksteps = (vox + 15) / 16;
    p.tile_tx = (uint32_t)vox * 128u;
    p.stages = (int)((220 * 1024 - 1024) / STAGE);
    if (p.stages > 4) p.stages = 4;
    if (p.stages < 2) return -1;
    const int smem_bytes = p.stages * (int)STAGE + 1024 + 256;

    auto enc = get_tensor_map_encoder();
    if (!enc) return -2;
    CUtensorMap tx, tdy;
    cuuint32_t box[5] = {64, (cuuint32_t)W, (cuuint32_t)p.Hb, (cuuint32_t)p.Db, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    {
        cuuint64_t dims[5] = {(cuuint64_t)CIN, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)CIN * 2, (cuuint64_t)W * CIN * 2, (cuuint64_t)H * W * CIN * 2, (cuuint64_t)D * H * W * CIN * 2};
        if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
    }
    {
        cuuint64_t dims[5] = {(cuuint64_t)COUT, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)COUT * 2, (cuuint64_t)W * COUT * 2, (cuuint64_t)H * W * COUT * 2, (cuuint64_t)D * H * W * COUT * 2};
        if (enc(&tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
    }
    static int configured = 0;
    if (configured < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_tap_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = smem_bytes;
    }
    const int groups = (27 * (CIN / 64) + 1) / 2;
    int ksplit = B200_SM_COUNT / groups;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > p.num_boxes) ksplit = p.num_boxes;
    conv3d_wgrad_tap_kernel<CIN, COUT><<<dim3(groups, ksplit), WT_THREADS, smem_bytes, st>>>(tx, tdy, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// x: [N,D,H,W,cin] bf16, dy: [N,D,H,W,cout] bf16, dwt: [27*cin, cout] fp32 (accumulated into); -1 = shape not covered
COINN_API int coinn_conv3d_wgrad_tap(const void* x, const void* dy, float* dwt, int N, int D, int H, int W, int cin, int cout, void* stream) {
    using namespace coinn;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cin == 64 && cout == 128) return launch_wgrad_tap<64, 128>(x, dy, dwt, N, D, H, W, st);
    if (cin == 128 && cout == 256) return launch_wgrad_tap<128, 256>(x, dy, dwt, N, D, H, W, st);
    if (cin == 64 && cout == 64) return launch_wgrad_tap<64, 64>(x, dy, dwt, N, D, H, W, st);
    return -1;
}
