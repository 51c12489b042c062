// Conv3d (3x3x3, pad 1, stride 1) fprop / dgrad, version 2: the im2col operand is produced by the TMA unit.
//
// Output tile = TH x TW pixels (TH*TW = 128) of one (n, d) slice.  For every filter tap (kd,kh,kw) ONE
// cp.async.bulk.tensor.5d load of the box {CB channels, TW, TH, 1, 1} at coordinates
// (c0, w0+kw-1, h0+kh-1, d+kd-1, n) of the NDHWC input drops the 128 shifted pixels x CB channels into shared
// memory, already in the K-major swizzled layout tcgen05.mma consumes; everything outside the volume is
// zero-filled by the TMA unit (padding costs nothing, no address math, no registers).  The matching weight slice
// Wk[:, tap*CIN + c0 : +CB] arrives with a 2-D TMA load on the same mbarrier.  One pipeline stage = one
// (tap, channel-box) unit = CB/16 MMAs of 128 x COUT x 16.
//
//   warp 0     TMA producer (A box + B slice per unit)
//   warp 1     MMA issuer, fp32 accumulators double-buffered in TMEM (2 x COUT columns)
//   warps 2-5  epilogue: tcgen05.ld -> bf16 -> masked channels-last stores
// Persistent: one CTA per SM walks the tile list.  dgrad = same kernel on dy with flipped/transposed weights.
#include "umma.cuh"

namespace coinn {

constexpr int CT_THREADS = 192;

struct ConvTmaParams {
    __nv_bfloat16* y;           // [N, D, H, W, COUT]
    int N, D, H, W;
    int TW, TH;                 // tile = TH x TW pixels, TH*TW == 128, TW power of two
    int tiles_w, tiles_h;
    int num_tiles;
};

template <int CIN, int COUT> struct ConvTmaCfg {
    static constexpr int CB = CIN < 64 ? CIN : 64;                  // channels per TMA box (<= 128 bytes)
    static constexpr int BOXES = CIN / CB;
    static constexpr int UNITS = 27 * BOXES;
    static constexpr int A_BYTES = 128 * CB * 2;
    static constexpr int B_BYTES_RAW = COUT * CB * 2;
    static constexpr int B_BYTES = (B_BYTES_RAW + 1023) / 1024 * 1024;   // keep every operand 1024-B aligned
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    __host__ __device__ static constexpr int stages() {
        return (200 * 1024) / STAGE_BYTES > 24 ? 24 : (200 * 1024) / STAGE_BYTES;
    }
    __host__ __device__ static constexpr int smem_bytes() { return stages() * STAGE_BYTES + 1024 + 1024; }
    static constexpr uint64_t LAYOUT = CB == 16 ? SMEM_LAYOUT_SW32 : (CB == 32 ? SMEM_LAYOUT_SW64 : SMEM_LAYOUT_SW128);
    static constexpr uint32_t SBO = 8 * CB * 2;                     // 8 rows of CB bf16
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv3d_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvTmaParams p) {
    using Cfg = ConvTmaCfg<CIN, COUT>;
    constexpr int STAGES = Cfg::stages();
    constexpr int CB = Cfg::CB;
    constexpr uint32_t TMEM_COLS = (2 * COUT) < 32 ? 32 : 2 * COUT;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int first_tile = blockIdx.x, tile_step = gridDim.x;
    const int tiles_per_slice = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            uint32_t it = 0;
            for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
                const int slice = tile / tiles_per_slice, rem = tile % tiles_per_slice;
                const int n = slice / p.D, d = slice % p.D;
                const int h0 = (rem / p.tiles_w) * p.TH, w0 = (rem % p.tiles_w) * p.TW;
                for (int tap = 0; tap < 27; ++tap) {
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
#pragma unroll
                    for (int b = 0; b < Cfg::BOXES; ++b, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                        uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
                        if (leader) mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES_RAW);
                        if (leader) tma_load_5d(a_dst, &tmap_x, &full_bar[s], b * CB, w0 + kw - 1, h0 + kh - 1, d + kd - 1, n);
                        if (leader) tma_load_2d(a_dst + Cfg::A_BYTES, &tmap_w, &full_bar[s], tap * CIN + b * CB, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            constexpr uint32_t idesc = make_idesc_f16(128, COUT, 1, 0, 0);
            uint32_t it = 0, t = 0;
            for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++t) {
                const uint32_t a = t & 1;
                mbar_wait(&tmem_empty[a], ((t >> 1) & 1) ^ 1);
                tcgen05_after_sync();
                const uint32_t d_tmem = tmem_base + a * COUT;
                for (int u = 0; u < Cfg::UNITS; ++u, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tcgen05_after_sync();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < CB / 16; ++k) {
                        if (leader) umma_f16(d_tmem, make_smem_desc(a_addr + k * 32, 16, Cfg::SBO, Cfg::LAYOUT),
                                 make_smem_desc(b_addr + k * 32, 16, Cfg::SBO, Cfg::LAYOUT), idesc,
                                 (u > 0 || k > 0) ? 1u : 0u);
                    }
                    if (leader) umma_commit(&empty_bar[s]);
                }
                if (leader) umma_commit(&tmem_full[a]);
            }
        }
    } else {
        const int q = warp & 3;
        const int pix = q * 32 + lane;
        const int hh = pix / p.TW, ww = pix % p.TW;
        uint32_t t = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++t) {
            const uint32_t a = t & 1;
            const int slice = tile / tiles_per_slice, rem = tile % tiles_per_slice;
            const int h = (rem / p.tiles_w) * p.TH + hh, w = (rem % p.tiles_w) * p.TW + ww;
            const bool ok = h < p.H && w < p.W;
            __nv_bfloat16* out = p.y + (((long long)slice * p.H + h) * p.W + w) * COUT;
            mbar_wait(&tmem_full[a], (t >> 1) & 1);
            tcgen05_after_sync();
#pragma unroll 1
            for (int c = 0; c < COUT; c += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + a * COUT + c, r);
                tmem_ld_wait();
                if (ok) {
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
    }

    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// 5-D tiled tensor map over an NDHWC bf16 tensor
static int make_tmap_ndhwc(CUtensorMap* out, const void* base, int N, int D, int H, int W, int C, int box_c, int box_w, int box_h,
                           CUtensorMapSwizzle sw) {
    auto enc = get_tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return (int)enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <int CIN, int COUT>
static int launch_conv_tma(const void* x, const void* wk, void* y, int N, int D, int H, int W, int kpad, cudaStream_t st) {
    using Cfg = ConvTmaCfg<CIN, COUT>;
    ConvTmaParams p;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.N = N; p.D = D; p.H = H; p.W = W;
    int tw = 8;
    while (tw < W && tw < 128) tw <<= 1;
    p.TW = tw; p.TH = 128 / tw;
    p.tiles_w = (W + p.TW - 1) / p.TW;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.num_tiles = N * D * p.tiles_w * p.tiles_h;
    const CUtensorMapSwizzle sw = Cfg::CB == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (Cfg::CB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap tx, tw_;
    if (make_tmap_ndhwc(&tx, x, N, D, H, W, CIN, Cfg::CB, p.TW, p.TH, sw) != 0) return -2;
    if (make_tmap_2d_bf16(&tw_, wk, (uint64_t)COUT, (uint64_t)kpad, (uint64_t)kpad * 2, COUT, Cfg::CB, sw) != 0) return -3;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_tma_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes());
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int grid = p.num_tiles < B200_SM_COUNT ? p.num_tiles : B200_SM_COUNT;
    conv3d_tma_kernel<CIN, COUT><<<grid, CT_THREADS, Cfg::smem_bytes(), st>>>(tx, tw_, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// same contract as coinn_conv3d_igemm (x NDHWC bf16, wk [cout, kpad] with k = tap*cin + ci)
COINN_API int coinn_conv3d_tma(const void* x, const void* wk, void* y, int N, int D, int H, int W, int cin, int cout,
                               int kpad, void* stream) {
    using namespace coinn;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CASE(CI, CO) if (cin == CI && cout == CO) return launch_conv_tma<CI, CO>(x, wk, y, N, D, H, W, kpad, st);
    CASE(16, 32) CASE(32, 64) CASE(64, 128) CASE(128, 256)
    CASE(32, 16) CASE(64, 32) CASE(128, 64) CASE(256, 128)
    CASE(16, 16) CASE(32, 32) CASE(64, 64) CASE(128, 128)
#undef CASE
    return -1;
}
