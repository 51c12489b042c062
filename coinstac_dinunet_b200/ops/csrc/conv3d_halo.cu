// Conv3d (3x3x3, pad 1, stride 1) fprop / dgrad, version 3 ("halo"): every input voxel crosses L2 -> SM once.
//
// v2 (conv3d_tma.cu) loads one shifted 128-pixel box per filter tap: 27x redundant L2 traffic, which is what
// bounds it (profiles/: ~5 TB/s of L2->SM traffic).  Here a tile is TH complete rows of one (n, d) plane and the
// TMA unit materialises the zero-PADDED halo  (TH+2) x (W+2) pixels x 8 channels  per (d-plane, channel chunk)
// directly in shared memory (out-of-bounds coordinates are zero filled).  In that padded, flattened layout the
// operand of tap (kd,kh,kw) is the SAME buffer shifted by (kh*(W+2) + kw) pixels, i.e. 27 (x CIN/16) tcgen05.mma
// instructions read 27 different start addresses of one halo - no im2col is ever built.  Layout: K-major,
// no swizzle; a pixel's 8-channel chunk is 16 bytes, 8 consecutive pixels form one 128-byte core matrix
// (SBO = 128 B), the second chunk of a K=16 step lives one region further (LBO = region stride).
// The 27*CIN x COUT weight matrix is loaded once per (persistent) CTA and stays in shared memory.
// Output rows at the two padding columns of every line are computed and discarded (2/(W+2) waste).
#include "umma.cuh"
#include <cstdlib>

namespace coinn {

constexpr int CH_THREADS = 192;

struct ConvHaloParams {
    __nv_bfloat16* y;           // [N, D, H, W, COUT]
    int N, D, H, W;
    int TH, Wp;                 // rows per tile, padded width (W + 2); TH * Wp <= 128
    int tiles_h, num_tiles;
    uint32_t region_bytes;      // (TH+2) * Wp * 16, rounded up to 128
    uint32_t region_tx;         // exact bytes one halo box delivers
    int stages;
    float* stats;               // optional [2*COUT]: per-channel sum / sum of squares of the stored (bf16) outputs, for BatchNorm
};

__device__ __forceinline__ void tma_load_5d_h(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                              int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int CIN, int COUT, bool FULLPIX, bool STATS>
__global__ void __launch_bounds__(CH_THREADS, STATS ? 1 : 2)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvHaloParams p) {
    // FULLPIX: one box per d-plane holding whole pixels (CIN*2 = 32/64 bytes, 32B/64B swizzle) - 2-4x fewer and
    // 2-4x larger TMA requests than the chunk-plane layout (8 channels = 16 B per request, no swizzle).
    constexpr int CHUNKS = CIN / 8;                         // 16-byte channel chunks per pixel
    constexpr int REGIONS = FULLPIX ? 3 : 3 * CHUNKS;       // per tile: d-planes (x chunk planes)
    constexpr int PIXB = FULLPIX ? CIN * 2 : 16;            // bytes per pixel inside one region
    constexpr uint64_t A_LAYOUT = !FULLPIX ? SMEM_LAYOUT_NONE : (CIN == 16 ? SMEM_LAYOUT_SW32 : SMEM_LAYOUT_SW64);
    constexpr int KCH = 27 * CHUNKS;                        // weight k-chunks
    constexpr uint32_t W_CHUNK_BYTES = COUT * 16;
    constexpr uint32_t W_BYTES = (KCH * W_CHUNK_BYTES + 1023) / 1024 * 1024;
    constexpr uint32_t TMEM_COLS = (2 * COUT) < 32 ? 32 : 2 * COUT;
    constexpr int MAX_STAGES = 8;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w_smem = smem;
    uint8_t* stage_base = smem + W_BYTES;
    const uint32_t stage_bytes = REGIONS * p.region_bytes + 1024;       // + slack: shifted reads run past the last region
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* w_bar = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int STAGES = p.stages;
    __shared__ float stat_red[STATS ? 2 * COUT : 1];
    if (STATS && threadIdx.x < 2 * COUT) stat_red[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        mbar_init(w_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int first_tile = blockIdx.x, tile_step = gridDim.x;

    if (warp == 0) {
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            // weights: one [COUT x 8] box per k-chunk, resident for the lifetime of the CTA
            if (leader) mbar_arrive_expect_tx(w_bar, KCH * W_CHUNK_BYTES);
            for (int kc = 0; kc < KCH; ++kc) if (leader) tma_load_2d(w_smem + kc * W_CHUNK_BYTES, &tmap_w, w_bar, kc * 8, 0);
            uint32_t it = 0;
            for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++it) {
                const int plane = tile / p.tiles_h, h0 = (tile % p.tiles_h) * p.TH;
                const int n = plane / p.D, d = plane % p.D;
                const int s = it % STAGES;
                mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                uint8_t* dst = stage_base + (size_t)s * stage_bytes;
                if (leader) mbar_arrive_expect_tx(&full_bar[s], REGIONS * p.region_tx);
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    if (FULLPIX) {
                        if (leader) tma_load_5d_h(dst + kd * p.region_bytes, &tmap_x, &full_bar[s], 0, -1, h0 - 1, d + kd - 1, n);
                    } else {
#pragma unroll
                        for (int c = 0; c < CHUNKS; ++c)
                            if (leader) tma_load_5d_h(dst + (kd * CHUNKS + c) * p.region_bytes, &tmap_x, &full_bar[s], c * 8, -1, h0 - 1, d + kd - 1, n);
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            constexpr uint32_t idesc = make_idesc_f16(128, COUT, 1, 0, 0);
            mbar_wait(w_bar, 0);
            const uint32_t w_addr = smem_u32(w_smem);
            uint32_t it = 0;
            for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++it) {
                const uint32_t a = it & 1;
                const int s = it % STAGES;
                mbar_wait(&tmem_empty[a], ((it >> 1) & 1) ^ 1);
                mbar_wait(&full_bar[s], (it / STAGES) & 1);
                tcgen05_after_sync();
                const uint32_t d_tmem = tmem_base + a * COUT;
                // Descriptor = constant part | (address >> 4).  All per-tap offsets are affine in (kd, kh, kw), so the
                // fully unrolled nest below costs ~3 integer instructions per MMA (the first version divided tap by
                // 9 and 3 and rebuilt both 64-bit descriptors for every MMA: ~185 cycles of scalar code per MMA in the
                // single issuing thread, which - not the tensor core - bounded the kernel; see profiles/).
                const uint32_t halo16 = (smem_u32(stage_base + (size_t)s * stage_bytes) & 0x3FFFFu) >> 4;
                const uint32_t region16 = p.region_bytes >> 4;
                const uint32_t row16 = (uint32_t)p.Wp * (PIXB / 16);
                const uint64_t a_const = FULLPIX ? make_smem_desc(0, 16, 8 * PIXB, A_LAYOUT)
                                                 : make_smem_desc(0, p.region_bytes, 128, SMEM_LAYOUT_NONE);
                const uint64_t b_const = make_smem_desc(0, W_CHUNK_BYTES, 128, SMEM_LAYOUT_NONE);
                const uint32_t w16 = (w_addr & 0x3FFFFu) >> 4;
                // every offset below is a compile-time multiple of a warp-uniform value: ptxas keeps both descriptors in
                // uniform registers and the 27 * CIN/16 UTCHMMA issue back to back
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                            for (int j = 0; j < CIN / 16; ++j) {
                                const uint32_t a16 = FULLPIX
                                    ? halo16 + kd * region16 + kh * row16 + kw * (PIXB / 16) + j * 2
                                    : halo16 + (kd * CHUNKS + 2 * j) * region16 + kh * row16 + kw;
                                const uint32_t b16 = w16 + (uint32_t)(((kd * 3 + kh) * 3 + kw) * (CIN / 16) + j) * (2 * (W_CHUNK_BYTES / 16));
                                if (leader) umma_f16(d_tmem, a_const | a16, b_const | b16, idesc, (kd | kh | kw | j) ? 1u : 0u);
                            }
                        }
                    }
                }
                if (leader) umma_commit(&empty_bar[s]);
                if (leader) umma_commit(&tmem_full[a]);
            }
        }
    } else {
        const int q = warp & 3;
        const int pix = q * 32 + lane;
        const int hh = pix / p.Wp, ww = pix % p.Wp;
        // BatchNorm batch statistics fused into the producer: the thread owns one pixel, i.e. all COUT channels
        float s1[STATS ? COUT : 1], s2[STATS ? COUT : 1];
        if (STATS) {
#pragma unroll
            for (int c = 0; c < COUT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
        }
        uint32_t it = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++it) {
            const uint32_t a = it & 1;
            const int plane = tile / p.tiles_h, h = (tile % p.tiles_h) * p.TH + hh;
            const bool ok = hh < p.TH && ww < p.W && h < p.H;
            __nv_bfloat16* out = p.y + (((long long)plane * p.H + h) * p.W + ww) * COUT;
            mbar_wait(&tmem_full[a], (it >> 1) & 1);
            tcgen05_after_sync();
#pragma unroll
            for (int c = 0; c < COUT; c += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + a * COUT + c, r);
                tmem_ld_wait();
                if (ok) {
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) pk[e] = pack_bf16x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
                    *reinterpret_cast<uint4*>(out + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(out + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    if (STATS) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float2 f = unpack_bf16x2(pk[e]);
                            s1[c + 2 * e] += f.x; s2[c + 2 * e] = fmaf(f.x, f.x, s2[c + 2 * e]);
                            s1[c + 2 * e + 1] += f.y; s2[c + 2 * e + 1] = fmaf(f.y, f.y, s2[c + 2 * e + 1]);
                        }
                    }
                }
            }
            tcgen05_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
        }
        if (STATS) {
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float sa = warp_sum(s1[c]), sb = warp_sum(s2[c]);
                if (lane == 0) { atomicAdd(&stat_red[c], sa); atomicAdd(&stat_red[COUT + c], sb); }
            }
        }
    }

    tcgen05_before_sync();
    __syncthreads();
    // one global atomic per channel and CTA (same-address atomics serialise in L2: 4 warps x 148 CTAs x 2*COUT of them
    // cost as much as the separate statistics pass they replace)
    if (STATS && threadIdx.x < 2 * COUT) atomicAdd(&p.stats[threadIdx.x], stat_red[threadIdx.x]);
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN, int COUT, bool FULLPIX, bool STATS>
static int launch_conv_halo(const void* x, const void* wk, void* y, float* stats, int N, int D, int H, int W, int kpad, cudaStream_t st) {
    constexpr int CHUNKS = CIN / 8, REGIONS = FULLPIX ? 3 : 3 * CHUNKS, KCH = 27 * CHUNKS;
    constexpr int PIXB = FULLPIX ? CIN * 2 : 16;
    constexpr uint32_t W_BYTES = (KCH * COUT * 16 + 1023) / 1024 * 1024;
    ConvHaloParams p;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.stats = stats;
    p.N = N; p.D = D; p.H = H; p.W = W;
    p.Wp = W + 2;
    if (p.Wp > 128) return -1;
    p.TH = 128 / p.Wp;
    if (p.TH > H) p.TH = H;
    if (p.TH + 2 > 256) return -1;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.num_tiles = N * D * p.tiles_h;
    p.region_tx = (uint32_t)(p.TH + 2) * p.Wp * (uint32_t)PIXB;
    p.region_bytes = (p.region_tx + 1023u) & ~1023u;          // 1 KB: keeps swizzle phases of the regions identical
    const uint32_t stage_bytes = REGIONS * p.region_bytes + 1024;
    // two CTAs per SM (each with its own MMA-issue thread, TMA queue and epilogue) when there are plenty of tiles and
    // the weights + 3 halo stages fit twice: the per-tile latency chains of the two CTAs overlap
    static int ctas_env = -1;
    if (ctas_env < 0) { const char* e = getenv("COINN_HALO_CTAS"); ctas_env = e ? atoi(e) : 1; }   // measured: 2 CTAs/SM 161 us vs 1 CTA 154 us (layer-2 fprop): no gain
    int ctas = (!STATS && ctas_env >= 2 && p.num_tiles >= 4 * B200_SM_COUNT && 2 * COUT * 2 <= 256 &&
                (int)W_BYTES + 3 * (int)stage_bytes + 2048 <= 110 * 1024) ? 2 : 1;
    const int budget = (ctas == 2 ? 110 : 220) * 1024 - (int)W_BYTES - 1024 - 512;
    int stages = budget / (int)stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return -1;
    p.stages = stages;
    const int smem_bytes = (int)W_BYTES + stages * (int)stage_bytes + 1024 + 512;

    auto enc = get_tensor_map_encoder();
    if (!enc) return -2;
    CUtensorMap tx, tw;
    {
        cuuint64_t dims[5] = {(cuuint64_t)CIN, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)CIN * 2, (cuuint64_t)W * CIN * 2, (cuuint64_t)H * W * CIN * 2, (cuuint64_t)D * H * W * CIN * 2};
        cuuint32_t box[5] = {(cuuint32_t)(FULLPIX ? CIN : 8), (cuuint32_t)p.Wp, (cuuint32_t)(p.TH + 2), 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                FULLPIX ? (CIN == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B) : CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
    }
    if (make_tmap_2d_bf16(&tw, wk, (uint64_t)COUT, (uint64_t)kpad, (uint64_t)kpad * 2, COUT, 8, CU_TENSOR_MAP_SWIZZLE_NONE) != 0) return -4;
    static int configured = 0;
    if (configured < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_halo_kernel<CIN, COUT, FULLPIX, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = smem_bytes;
    }
    const int grid = p.num_tiles < ctas * B200_SM_COUNT ? p.num_tiles : ctas * B200_SM_COUNT;
    conv3d_halo_kernel<CIN, COUT, FULLPIX, STATS><<<grid, CH_THREADS, smem_bytes, st>>>(tx, tw, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// same contract as coinn_conv3d_igemm; returns -1 when the shape is outside what the halo kernel covers
// fullpix != 0 selects the whole-pixel (swizzled) halo layout, available for cin in {16, 32}
// stats (optional, [2*cout] fp32, zeroed): per-channel sum / sum of squares of y for BatchNorm, cout <= 64 only (-1 otherwise)
COINN_API int coinn_conv3d_halo_stats(const void* x, const void* wk, void* y, float* stats, int N, int D, int H, int W, int cin, int cout,
                                      int kpad, int fullpix, void* stream) {
    using namespace coinn;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CASE(CI, CO) if (cin == CI && cout == CO) return stats ? launch_conv_halo<CI, CO, false, true>(x, wk, y, stats, N, D, H, W, kpad, st) \
                                                               : launch_conv_halo<CI, CO, false, false>(x, wk, y, nullptr, N, D, H, W, kpad, st);
#define CASEF(CI, CO) if (fullpix && cin == CI && cout == CO) return stats ? launch_conv_halo<CI, CO, true, true>(x, wk, y, stats, N, D, H, W, kpad, st) \
                                                                           : launch_conv_halo<CI, CO, true, false>(x, wk, y, nullptr, N, D, H, W, kpad, st);
    CASEF(16, 32) CASEF(32, 16) CASEF(32, 64) CASEF(16, 16) CASEF(32, 32)
    CASE(16, 32) CASE(32, 16) CASE(32, 64) CASE(64, 32) CASE(16, 16) CASE(32, 32)
#undef CASE
#undef CASEF
    return -1;
}

COINN_API int coinn_conv3d_halo(const void* x, const void* wk, void* y, int N, int D, int H, int W, int cin, int cout,
                                int kpad, int fullpix, void* stream) {
    return coinn_conv3d_halo_stats(x, wk, y, nullptr, N, D, H, W, cin, cout, kpad, fullpix, stream);
}
