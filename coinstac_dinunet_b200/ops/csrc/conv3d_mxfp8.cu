// Conv3d (3x3x3, pad 1, stride 1) fprop / dgrad in MX-FP8: e4m3 activations and weights, one ue8m0 scale per 32 input
// channels, `tcgen05.mma.kind::mxf8f6f4.block_scale` (BASELINE config 4).  Same implicit-GEMM structure as conv3d_tma.cu
// (the im2col operand is a 5-D TMA box per filter tap, zero fill = padding), plus the scale-factor plumbing:
//
//   operands   xq [N,D,H,W,CIN] e4m3 bytes, sfa [voxel][BOXES] words : byte j of word b = scale of channels [b*CB + 32 j, +32)
//              wq [COUT, 27*CIN] e4m3 bytes (k = tap*CIN + ci), sfb [COUT][27*BOXES] words, same byte convention per unit
//   unit       one (tap, channel box of CB = min(CIN,128) bytes) = CB/32 MMAs of 128 x COUT x 32
//   warp 0     TMA producer (A box + B slice per unit)
//   warp 1     MMA issuer: per unit  tcgen05.cp SFA / SFB (512-byte 32x4x4 blocks) smem -> TMEM, then the MMAs whose
//              instruction descriptor picks byte k of the scale words; accumulators (double-)buffered in TMEM
//   warps 2-5  epilogue: tcgen05.ld -> bf16 -> masked channels-last stores
//   warps 6-9  scale-factor loaders: thread = output pixel of the tile; the A scales are gathered at the tap-shifted voxel
//              (out of the volume -> 2^0, the data there is TMA zero fill), thread t also fetches the B scale word of row t
// Persistent: one CTA per SM walks the tile list.  dgrad = same kernel on quantised dy with flipped / transposed weights.
#include "umma.cuh"
#include <cuda_fp8.h>

namespace coinn {

constexpr int CQ_THREADS = 320;

struct ConvQParams {
    __nv_bfloat16* y;           // [N, D, H, W, COUT]
    const uint32_t* sfa;        // [N*D*H*W][BOXES]
    const uint32_t* sfb;        // [COUT][27*BOXES]
    int N, D, H, W;
    int TW, TH, tiles_w, tiles_h, num_tiles;
};

__host__ __device__ constexpr uint32_t pow2_cols(uint32_t need) {
    return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

template <int CIN, int COUT> struct ConvQCfg {
    static constexpr int CB = CIN < 128 ? CIN : 128;                 // bytes (= e4m3 elements) per TMA box row
    static constexpr int BOXES = CIN / CB;
    static constexpr int UNITS = 27 * BOXES;
    static constexpr int KSUB = CB / 32;                            // MMAs (K = 32) per unit
    static constexpr int A_BYTES = 128 * CB;
    static constexpr int B_BYTES_RAW = COUT * CB;
    static constexpr int B_BYTES = (B_BYTES_RAW + 1023) / 1024 * 1024;
    static constexpr int SFB_CHUNKS = (COUT + 127) / 128;
    static constexpr int SF_BYTES = 512 + 512 * SFB_CHUNKS;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int ACC_BUFS = COUT > 128 ? 1 : 2;
    static constexpr uint32_t SF_COL = ACC_BUFS * COUT;              // SFA at SF_COL, SFB at SF_COL + 4
    static constexpr uint32_t TMEM_COLS = pow2_cols(ACC_BUFS * COUT + 4 + 4 * SFB_CHUNKS);
    __host__ __device__ static constexpr int stages() {
        return (190 * 1024) / (STAGE_BYTES + SF_BYTES) > 16 ? 16 : (190 * 1024) / (STAGE_BYTES + SF_BYTES);
    }
    __host__ __device__ static constexpr int smem_bytes() { return stages() * (STAGE_BYTES + SF_BYTES) + 1024 + 1024; }
    static constexpr uint64_t LAYOUT = CB == 32 ? SMEM_LAYOUT_SW32 : (CB == 64 ? SMEM_LAYOUT_SW64 : SMEM_LAYOUT_SW128);
    static constexpr uint32_t SBO = 8 * CB;
};

__host__ __device__ constexpr uint32_t cq_idesc(uint32_t M, uint32_t N, uint32_t a_sf, uint32_t b_sf) {
    return (b_sf << 4) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (a_sf << 29);
}
__device__ __forceinline__ void cq_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                       uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb) : "memory");
}
__device__ __forceinline__ void cq_utccp(uint32_t tmem_dst, uint32_t smem_addr) {
    const uint64_t desc = make_smem_desc(smem_addr, 0, 128, SMEM_LAYOUT_NONE);
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" :: "r"(tmem_dst), "l"(desc) : "memory");
}
__device__ __forceinline__ void cq_tma_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(CQ_THREADS, 1)
conv3d_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvQParams p) {
    using Cfg = ConvQCfg<CIN, COUT>;
    constexpr int STAGES = Cfg::stages();
    constexpr int CB = Cfg::CB;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sf_smem = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sf_smem + STAGES * Cfg::SF_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* sf_bar = empty_bar + STAGES;
    uint64_t* tmem_full = sf_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&sf_bar[s], 128); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int first_tile = blockIdx.x, tile_step = gridDim.x;
    const int tiles_per_slice = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        const bool leader = elect_one();
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
                    if (leader) cq_tma_5d(a_dst, &tmap_x, &full_bar[s], b * CB, w0 + kw - 1, h0 + kh - 1, d + kd - 1, n);
                    if (leader) tma_load_2d(a_dst + Cfg::A_BYTES, &tmap_w, &full_bar[s], tap * CIN + b * CB, 0);
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        uint32_t it = 0, t = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++t) {
            const uint32_t a = Cfg::ACC_BUFS == 2 ? (t & 1) : 0;
            const uint32_t use = Cfg::ACC_BUFS == 2 ? (t >> 1) : t;
            mbar_wait(&tmem_empty[a], (use & 1) ^ 1);
            tcgen05_after_sync();
            const uint32_t d_tmem = tmem_base + a * COUT;
            for (int u = 0; u < Cfg::UNITS; ++u, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                mbar_wait(&sf_bar[s], ph);
                tcgen05_after_sync();
                const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t b_addr = a_addr + Cfg::A_BYTES;
                const uint32_t sf_addr = smem_u32(sf_smem + s * Cfg::SF_BYTES);
                if (leader) cq_utccp(tmem_base + Cfg::SF_COL, sf_addr);
#pragma unroll
                for (int c = 0; c < Cfg::SFB_CHUNKS; ++c)
                    if (leader) cq_utccp(tmem_base + Cfg::SF_COL + 4 + 4 * c, sf_addr + 512 + 512 * c);
#pragma unroll
                for (int k = 0; k < Cfg::KSUB; ++k) {
                    if (leader) cq_mma(d_tmem, make_smem_desc(a_addr + k * 32, 16, Cfg::SBO, Cfg::LAYOUT),
                                       make_smem_desc(b_addr + k * 32, 16, Cfg::SBO, Cfg::LAYOUT), cq_idesc(128, COUT, k, k),
                                       (u > 0 || k > 0) ? 1u : 0u, tmem_base + Cfg::SF_COL, tmem_base + Cfg::SF_COL + 4);
                }
                if (leader) umma_commit(&empty_bar[s]);
            }
            if (leader) umma_commit(&tmem_full[a]);
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        const int pix = q * 32 + lane;
        const int hh = pix / p.TW, ww = pix % p.TW;
        uint32_t t = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++t) {
            const uint32_t a = Cfg::ACC_BUFS == 2 ? (t & 1) : 0;
            const uint32_t use = Cfg::ACC_BUFS == 2 ? (t >> 1) : t;
            const int slice = tile / tiles_per_slice, rem = tile % tiles_per_slice;
            const int h = (rem / p.tiles_w) * p.TH + hh, w = (rem % p.tiles_w) * p.TW + ww;
            const bool ok = h < p.H && w < p.W;
            __nv_bfloat16* out = p.y + (((long long)slice * p.H + h) * p.W + w) * COUT;
            mbar_wait(&tmem_full[a], use & 1);
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
    } else {
        // ---- scale-factor loaders: thread tq = pixel of the tile (A) and weight row (B) ----
        const int tq = threadIdx.x - 192;                      // 0..127
        const int slot = (tq & 31) * 4 + (tq >> 5);
        const int hh = tq / p.TW, ww = tq % p.TW;
        uint32_t it = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
            const int slice = tile / tiles_per_slice, rem = tile % tiles_per_slice;
            const int n = slice / p.D, d = slice % p.D;
            const int h0 = (rem / p.tiles_w) * p.TH, w0 = (rem % p.tiles_w) * p.TW;
            for (int tap = 0; tap < 27; ++tap) {
                const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                const int dd = d + kd - 1, h = h0 + hh + kh - 1, w = w0 + ww + kw - 1;
                const bool inside = dd >= 0 && dd < p.D && h >= 0 && h < p.H && w >= 0 && w < p.W;
                const long long vox = (((long long)n * p.D + dd) * p.H + h) * p.W + w;
#pragma unroll
                for (int b = 0; b < Cfg::BOXES; ++b, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint32_t* dst = reinterpret_cast<uint32_t*>(sf_smem + s * Cfg::SF_BYTES);
                    dst[slot] = inside ? __ldg(p.sfa + vox * Cfg::BOXES + b) : 0x7f7f7f7fu;
                    const int u = tap * Cfg::BOXES + b;
#pragma unroll
                    for (int c = 0; c < Cfg::SFB_CHUNKS; ++c) {
                        const int row = c * 128 + tq;
                        dst[128 + c * 128 + slot] = row < COUT ? __ldg(p.sfb + (long long)row * Cfg::UNITS + u) : 0x7f7f7f7fu;
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&sf_bar[s]);
                }
            }
        }
    }

    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

static int cq_tmap_ndhwc_u8(CUtensorMap* out, const void* base, int N, int D, int H, int W, int C, int box_c, int box_w, int box_h,
                            CUtensorMapSwizzle sw) {
    auto enc = get_tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C, (cuuint64_t)W * C, (cuuint64_t)H * W * C, (cuuint64_t)D * H * W * C};
    cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return (int)enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
static int cq_tmap_2d_u8(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                         CUtensorMapSwizzle sw) {
    auto enc = get_tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return (int)enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <int CIN, int COUT>
static int launch_conv_q(const void* xq, const void* sfa, const void* wq, const void* sfb, void* y, int N, int D, int H, int W,
                         cudaStream_t st) {
    using Cfg = ConvQCfg<CIN, COUT>;
    static_assert(Cfg::stages() >= 2, "not enough shared memory for a pipeline");
    ConvQParams p;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.sfa = reinterpret_cast<const uint32_t*>(sfa);
    p.sfb = reinterpret_cast<const uint32_t*>(sfb);
    p.N = N; p.D = D; p.H = H; p.W = W;
    int tw = 8;
    while (tw < W && tw < 128) tw <<= 1;
    p.TW = tw; p.TH = 128 / tw;
    p.tiles_w = (W + p.TW - 1) / p.TW;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.num_tiles = N * D * p.tiles_w * p.tiles_h;
    const CUtensorMapSwizzle sw = Cfg::CB == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (Cfg::CB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap tx, tw_;
    if (cq_tmap_ndhwc_u8(&tx, xq, N, D, H, W, CIN, Cfg::CB, p.TW, p.TH, sw) != 0) return -2;
    if (cq_tmap_2d_u8(&tw_, wq, (uint64_t)COUT, (uint64_t)27 * CIN, COUT, Cfg::CB, sw) != 0) return -3;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_mxfp8_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes());
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int grid = p.num_tiles < B200_SM_COUNT ? p.num_tiles : B200_SM_COUNT;
    conv3d_mxfp8_kernel<CIN, COUT><<<grid, CQ_THREADS, Cfg::smem_bytes(), st>>>(tx, tw_, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------ grouped quantiser
// x [R, K] (row stride ld elements) -> q [R, K] e4m3 bytes, sf [R][K / (32 G)] words: byte j (< G) of word g = scale of the
// 32-element block g*G + j; bytes >= G stay 0x7f (the caller memsets sf to 0x7f).  G = blocks per word (1, 2 or 4).
template <typename TIn> __device__ __forceinline__ float cq_load(const TIn* p);
template <> __device__ __forceinline__ float cq_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float cq_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TIn>
__global__ void __launch_bounds__(256) quantize_mx_grouped_kernel(const TIn* __restrict__ x, long long ld, uint8_t* __restrict__ q,
                                                                  uint8_t* __restrict__ sf, long long R, int K, int G) {
    const int blocks_per_row = K / 32;
    const int words_per_row = blocks_per_row / G;
    const long long total = R * blocks_per_row;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (long long)gridDim.x * blockDim.x) {
        const long long row = b / blocks_per_row;
        const int kb = (int)(b % blocks_per_row), k0 = kb * 32;
        float v[32];
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = cq_load(x + row * ld + k0 + j); amax = fmaxf(amax, fabsf(v[j])); }
        int e = 127;
        if (amax > 0.f && isfinite(amax)) {
            int ex;
            const float m = frexpf(amax / 448.f, &ex);
            e = ex + 127 - (m == 0.5f ? 1 : 0);
            e = e < 1 ? 1 : (e > 254 ? 254 : e);
        }
        const float inv = exp2f((float)(127 - e));
        uint32_t packed[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t u = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                u |= (uint32_t)__nv_cvt_float_to_fp8(v[w * 4 + j] * inv, __NV_SATFINITE, __NV_E4M3) << (8 * j);
            packed[w] = u;
        }
        uint4* dst = reinterpret_cast<uint4*>(q + row * K + k0);
        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        sf[(row * words_per_row + kb / G) * 4 + (kb % G)] = (uint8_t)e;
    }
}

}  // namespace coinn

using namespace coinn;

// x: [R, K] fp32 (0) / bf16 (1) with row stride ld -> q [R, K] bytes, sf [R, K/(32 G)] words (pre-filled with 0x7f by the caller)
COINN_API int coinn_quantize_mx_grouped(const void* x, int dtype, long long ld, void* q, void* sf, long long R, int K, int G, void* stream) {
    if (K % 32 || (G != 1 && G != 2 && G != 4) || (K / 32) % G) return (int)cudaErrorInvalidValue;
    if (R == 0) return 0;
    const long long total = R * (K / 32);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == 0) quantize_mx_grouped_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, ld, (uint8_t*)q, (uint8_t*)sf, R, K, G);
    else quantize_mx_grouped_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, ld, (uint8_t*)q, (uint8_t*)sf, R, K, G);
    COINN_CHECK_LAUNCH();
    return 0;
}

// xq [N,D,H,W,cin] e4m3, sfa [voxels][boxes] words, wq [cout, 27*cin] e4m3, sfb [cout][27*boxes] words -> y [N,D,H,W,cout] bf16
// returns -1 when there is no instantiation for (cin, cout)
COINN_API int coinn_conv3d_mxfp8(const void* xq, const void* sfa, const void* wq, const void* sfb, void* y, int N, int D, int H, int W,
                                 int cin, int cout, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CASE(CI, CO) if (cin == CI && cout == CO) return launch_conv_q<CI, CO>(xq, sfa, wq, sfb, y, N, D, H, W, st);
    CASE(32, 64) CASE(64, 128) CASE(128, 256)
    CASE(64, 32) CASE(128, 64) CASE(256, 128)
    CASE(32, 32) CASE(64, 64) CASE(128, 128)
#undef CASE
    return -1;
}
