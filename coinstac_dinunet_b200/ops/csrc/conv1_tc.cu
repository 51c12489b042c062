// First VBM block (C_in = 1) on the tensor cores.
//
// conv1_fwd_tc : y[p, co] = sum_tap x[p + off(tap)] * W1[co, tap]  as a 128 x 16 x 32 tcgen05 MMA per 128-pixel
//                row segment (K = 27 taps padded to 32).  The single-channel im2col row (27 values) is read
//                straight from global/L1 by the thread that owns the pixel, converted to bf16 and written into
//                the K-major 64B-swizzled A tile; BatchNorm statistics of the stored (bf16) y are accumulated in
//                registers across all tiles of the persistent CTA.
// conv1_wgrad_tc: dW1[tap, co] = sum_p x[p + off(tap)] * dy[p, co]: the same im2col tile, now used as the
//                MN-major A operand (M = taps, padded to 128; K = 128 pixels = 8 MMAs), dy arrives by TMA
//                (5-D box, 32B swizzle, zero fill beyond the row end), one TMEM accumulator per CTA, fp32 atomics
//                at the very end.
// The CUDA-core versions in vbm_fused.cu are bound by FMA issue (fwd) and by shared-memory broadcast
// bandwidth (wgrad: 1.18 ms per step); here both become bound by writing/reading the 16-channel tensor.
#include "umma.cuh"

namespace coinn {

struct C1Dims { int N, D, H, W; int tiles_w; long long num_tiles; };

__device__ __forceinline__ void c1_tile_coords(const C1Dims& d, long long tile64, int& n, int& dd, int& h, int& w0) {
    unsigned tile = (unsigned)tile64;                      // < 2^31 tiles: 32-bit divisions are ~3x cheaper
    if (d.tiles_w == 1) { w0 = 0; } else { w0 = (int)(tile % (unsigned)d.tiles_w) * 128; tile /= (unsigned)d.tiles_w; }
    h = (int)(tile % (unsigned)d.H); tile /= (unsigned)d.H;
    dd = (int)(tile % (unsigned)d.D);
    n = (int)(tile / (unsigned)d.D);
}

// 27 neighbourhood values of pixel (n, dd, h, w) (zero outside the volume).  One 32-bit centre offset + constant tap
// offsets and three 3-bit validity masks: no per-tap multiplies, all 27 loads independent (in flight together).
__device__ __forceinline__ float c1_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float c1_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TX>
__device__ __forceinline__ void c1_load_taps(const TX* __restrict__ x, const C1Dims& d, int n, int dd, int h, int w, bool active,
                                             float (&v)[27]) {
    const int HW = d.H * d.W;
    const int centre = ((n * d.D + dd) * d.H + h) * d.W + w;           // < 2^31 elements for any volume we support
    const uint32_t vd = active ? ((dd > 0 ? 1u : 0u) | 2u | (dd + 1 < d.D ? 4u : 0u)) : 0u;
    const uint32_t vh = (h > 0 ? 1u : 0u) | 2u | (h + 1 < d.H ? 4u : 0u);
    const uint32_t vw = (w > 0 ? 1u : 0u) | 2u | (w + 1 < d.W ? 4u : 0u);
    const TX* c = x + centre;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const bool row_ok = ((vd >> kd) & 1u) && ((vh >> kh) & 1u);
            const int roff = (kd - 1) * HW + (kh - 1) * d.W;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const bool ok = row_ok && ((vw >> kw) & 1u);
                v[(kd * 3 + kh) * 3 + kw] = ok ? c1_ld(c + roff + (kw - 1)) : 0.f;
            }
        }
    }
}

// write one im2col row (32 bf16 = 64 B, taps 27..31 zero) into a K-major / MN-major swizzled tile
template <int ROW_BYTES>   // 64: SW64 rows (fwd),  128: SW128 rows (wgrad, taps 32..63 stay zero)
__device__ __forceinline__ void c1_store_row(uint8_t* tile, int p, const float (&v)[27]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 13; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    w[13] = pack_bf16x2(v[26], 0.f);
    w[14] = 0u; w[15] = 0u;
    const uint32_t sw = ROW_BYTES == 64 ? (uint32_t)((p >> 1) & 3) : (uint32_t)(p & 7);
    uint8_t* row = tile + (size_t)p * ROW_BYTES;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(row + (((uint32_t)c ^ sw) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int C1F_THREADS = 192;      // warp 0: idle/setup, warp 1: MMA, warps 2-5: build + epilogue

template <typename TX>
__global__ void __launch_bounds__(C1F_THREADS, 4)
conv1_fwd_tc_kernel(const TX* __restrict__ x, const float* __restrict__ w /*[16][27]*/, __nv_bfloat16* __restrict__ y,
                    float* __restrict__ stats, const C1Dims d) {
    __shared__ __align__(1024) uint8_t a_tile[2][128 * 64];      // K-major, 64 B rows, 64B swizzle
    __shared__ __align__(1024) uint8_t b_tile[16 * 64];          // W1 [16 co x 32 taps], same layout
    __shared__ uint64_t a_ready[2], tmem_full[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float red[32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_ready[i], 128); mbar_init(&tmem_full[i], 1); }
        fence_mbar_init();
    }
    if (threadIdx.x < 32) red[threadIdx.x] = 0.f;
    if (threadIdx.x < 16) {                                      // weights -> bf16 K-major swizzled rows
        float v[27];
#pragma unroll
        for (int t = 0; t < 27; ++t) v[t] = w[threadIdx.x * 27 + t];
        c1_store_row<64>(b_tile, threadIdx.x, v);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 32);
    fence_proxy_async_smem();
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const long long first = blockIdx.x, step = gridDim.x;

    if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, 16, 1, 0, 0);
            const uint32_t b_addr = smem_u32(b_tile);
            uint32_t t = 0;
            for (long long tile = first; tile < d.num_tiles; tile += step, ++t) {
                const uint32_t a = t & 1;
                mbar_wait(&a_ready[a], (t >> 1) & 1);            // builders finished tile t (and drained TMEM buffer a)
                tcgen05_after_sync();
                const uint32_t a_addr = smem_u32(a_tile[a]);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_f16(tmem_base + a * 16, make_smem_desc(a_addr + k * 32, 16, 512, SMEM_LAYOUT_SW64),
                             make_smem_desc(b_addr + k * 32, 16, 512, SMEM_LAYOUT_SW64), idesc, k);
                umma_commit(&tmem_full[a]);
            }
        }
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int p = q * 32 + lane;                              // pixel inside the tile == TMEM lane
        float s1[16], s2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { s1[c] = 0.f; s2[c] = 0.f; }

        float v[27];
        int n, dd, h, w0;
        uint32_t t = 0;
        if (first < d.num_tiles) {                                // prologue: build tile 0
            c1_tile_coords(d, first, n, dd, h, w0);
            c1_load_taps(x, d, n, dd, h, w0 + p, w0 + p < d.W, v);
            c1_store_row<64>(a_tile[0], p, v);
            fence_proxy_async_smem();
            mbar_arrive(&a_ready[0]);
        }
        for (long long tile = first; tile < d.num_tiles; tile += step, ++t) {
            const uint32_t a = t & 1;
            const long long nxt = tile + step;
            int n2 = 0, d2 = 0, h2 = 0, w2 = 0;
            const bool has_next = nxt < d.num_tiles;
            if (has_next) {                                       // loads of tile t+1 fly while we drain tile t
                c1_tile_coords(d, nxt, n2, d2, h2, w2);
                c1_load_taps(x, d, n2, d2, h2, w2 + p, w2 + p < d.W, v);
            }
            // ---- epilogue of tile t (its coordinates were decoded one iteration ago)
            mbar_wait(&tmem_full[a], (t >> 1) & 1);
            tcgen05_after_sync();
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + a * 16, r);
            tmem_ld_wait();
            tcgen05_before_sync();
            if (w0 + p < d.W) {
                uint4 lo = make_uint4(pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3])),
                                      pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7])));
                uint4 hi = make_uint4(pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9])), pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11])),
                                      pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13])), pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15])));
                __nv_bfloat16* out = y + ((((long long)n * d.D + dd) * d.H + h) * d.W + w0 + p) * 16;
                reinterpret_cast<uint4*>(out)[0] = lo;
                reinterpret_cast<uint4*>(out)[1] = hi;
                const uint32_t pk[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {                      // statistics of the stored (rounded) values
                    const float2 f = unpack_bf16x2(pk[i]);
                    s1[2 * i] += f.x; s2[2 * i] = fmaf(f.x, f.x, s2[2 * i]);
                    s1[2 * i + 1] += f.y; s2[2 * i + 1] = fmaf(f.y, f.y, s2[2 * i + 1]);
                }
            }
            // ---- build tile t+1 into the other A buffer (its previous MMA, tile t-1, completed before tmem_full[t-1])
            if (has_next) {
                c1_store_row<64>(a_tile[a ^ 1], p, v);
                fence_proxy_async_smem();
                mbar_arrive(&a_ready[a ^ 1]);
                n = n2; dd = d2; h = h2; w0 = w2;
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
    if (threadIdx.x < 32) atomicAdd(&stats[threadIdx.x], red[threadIdx.x]);
    if (warp == 1) tmem_dealloc(tmem_base, 32);
}

// ------------------------------------------------------------------------------------------------- wgrad
constexpr int C1W_THREADS = 192;      // warp 0: dy TMA producer, warp 1: MMA, warps 2-5: builders (+ final epilogue)
constexpr int C1W_STAGES = 3;         // dy tiles in flight

__device__ __forceinline__ void tma_load_5d_c1(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <typename TX>
__global__ void __launch_bounds__(C1W_THREADS, 3)
conv1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const TX* __restrict__ x, float* __restrict__ dw /*[16][27]*/,
                      const C1Dims d) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile0 = smem;                      // [128 px][128 B]  taps 0..63 (only 0..31 ever written)
    uint8_t* a_tile1 = smem + 16384;
    uint8_t* a_zero = smem + 32768;               // second MN block (taps 64..127): all zero, shared by both buffers
    uint8_t* b_base = smem + 49152;               // C1W_STAGES x [128 px][32 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + C1W_STAGES * 4096);
    uint64_t* a_ready = bars;                     // [2]   builders -> MMA
    uint64_t* a_free = bars + 2;                  // [2]   MMA -> builders
    uint64_t* b_full = bars + 4;                  // [C1W_STAGES] TMA -> MMA
    uint64_t* b_empty = b_full + C1W_STAGES;      // [C1W_STAGES] MMA -> TMA
    uint64_t* done_bar = b_empty + C1W_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 49152 / 16; i += C1W_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_dy);
        for (int i = 0; i < 2; ++i) { mbar_init(&a_ready[i], 128); mbar_init(&a_free[i], 1); }
        for (int i = 0; i < C1W_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 32);
    fence_proxy_async_smem();
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const long long first = blockIdx.x, step = gridDim.x;
    const long long my_tiles = first < d.num_tiles ? (d.num_tiles - first + step - 1) / step : 0;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t t = 0;
            for (long long tile = first; tile < d.num_tiles; tile += step, ++t) {
                int n, dd, h, w0;
                c1_tile_coords(d, tile, n, dd, h, w0);
                const int s = t % C1W_STAGES;
                mbar_wait(&b_empty[s], ((t / C1W_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&b_full[s], 4096);
                tma_load_5d_c1(b_base + s * 4096, &tmap_dy, &b_full[s], 0, w0, h, dd, n);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, 16, 1, 1, 1);      // A (taps) and B (co) both MN-major
            uint32_t t = 0;
            for (long long tile = first; tile < d.num_tiles; tile += step, ++t) {
                const uint32_t a = t & 1;
                const int s = t % C1W_STAGES;
                mbar_wait(&a_ready[a], (t >> 1) & 1);
                mbar_wait(&b_full[s], (t / C1W_STAGES) & 1);
                tcgen05_after_sync();
                const uint32_t a_addr = smem_u32(a ? a_tile1 : a_tile0);
                const uint32_t lbo = smem_u32(a_zero) - a_addr;               // second MN block = the shared zero block
                const uint32_t b_addr = smem_u32(b_base + s * 4096);
#pragma unroll
                for (int k = 0; k < 8; ++k)                                    // 16 pixels per MMA
                    umma_f16(tmem_base, make_smem_desc(a_addr + k * 2048, lbo, 1024, SMEM_LAYOUT_SW128),
                             make_smem_desc(b_addr + k * 512, 4096, 256, SMEM_LAYOUT_SW32), idesc, (t > 0 || k > 0) ? 1u : 0u);
                umma_commit(&a_free[a]);
                umma_commit(&b_empty[s]);
            }
            umma_commit(done_bar);
        }
    } else {
        const int q = warp & 3;
        const int p = q * 32 + lane;
        float v[27], vn[27];
        uint32_t t = 0;
        int n, dd, h, w0;
        if (first < d.num_tiles) {
            c1_tile_coords(d, first, n, dd, h, w0);
            c1_load_taps(x, d, n, dd, h, w0 + p, w0 + p < d.W, v);               // zero rows beyond the line end
        }
        for (long long tile = first; tile < d.num_tiles; tile += step, ++t) {
            const uint32_t a = t & 1;
            const long long nxt = tile + step;
            if (nxt < d.num_tiles) {                                             // next tile's loads fly during the store
                c1_tile_coords(d, nxt, n, dd, h, w0);
                c1_load_taps(x, d, n, dd, h, w0 + p, w0 + p < d.W, vn);
            }
            mbar_wait(&a_free[a], ((t >> 1) & 1) ^ 1);
            c1_store_row<128>(a ? a_tile1 : a_tile0, p, v);
            fence_proxy_async_smem();
            mbar_arrive(&a_ready[a]);
#pragma unroll
            for (int i = 0; i < 27; ++i) v[i] = vn[i];
        }
        if (my_tiles > 0) {
            mbar_wait(done_bar, 0);
            tcgen05_after_sync();
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16), r);
            tmem_ld_wait();
            if (p < 27) {
#pragma unroll
                for (int c = 0; c < 16; ++c) atomicAdd(&dw[c * 27 + p], __uint_as_float(r[c]));
            }
        }
    }
    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 32);
}

}  // namespace coinn

// x: [N,D,H,W] fp32 (x_dtype 0) or bf16 (1); w: [16,27] fp32; y: [N,D,H,W,16] bf16; stats: 32 floats (zeroed)
COINN_API int coinn_conv1_fwd_tc(const void* x, int x_dtype, const float* w, void* y, float* stats, int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    C1Dims d{N, D, H, W, (W + 127) / 128, 0};
    d.num_tiles = (long long)N * D * H * d.tiles_w;
    const long long cap = 4LL * B200_SM_COUNT;      // several small CTAs per SM overlap their per-tile barrier chains
    const int grid = (int)(d.num_tiles < cap ? d.num_tiles : cap);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_dtype == 0) conv1_fwd_tc_kernel<float><<<grid, C1F_THREADS, 0, st>>>((const float*)x, w, (__nv_bfloat16*)y, stats, d);
    else conv1_fwd_tc_kernel<__nv_bfloat16><<<grid, C1F_THREADS, 0, st>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, stats, d);
    COINN_CHECK_LAUNCH();
    return 0;
}

// dy: [N,D,H,W,16] bf16; x: [N,D,H,W] fp32 (x_dtype 0) or bf16 (1); dw: [16*27] fp32 (zeroed)
COINN_API int coinn_conv1_wgrad_tc(const void* dy, const void* x, int x_dtype, float* dw, int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    C1Dims d{N, D, H, W, (W + 127) / 128, 0};
    d.num_tiles = (long long)N * D * H * d.tiles_w;
    auto enc = get_tensor_map_encoder();
    if (!enc) return -2;
    CUtensorMap tdy;
    cuuint64_t dims[5] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {32, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32, (cuuint64_t)D * H * W * 32};
    cuuint32_t box[5] = {16, 128, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
    const int smem_bytes = 49152 + C1W_STAGES * 4096 + 256 + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv1_wgrad_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv1_wgrad_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const long long capw = 3LL * B200_SM_COUNT;
    const int grid = (int)(d.num_tiles < capw ? d.num_tiles : capw);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_dtype == 0) conv1_wgrad_tc_kernel<float><<<grid, C1W_THREADS, smem_bytes, st>>>(tdy, (const float*)x, dw, d);
    else conv1_wgrad_tc_kernel<__nv_bfloat16><<<grid, C1W_THREADS, smem_bytes, st>>>(tdy, (const __nv_bfloat16*)x, dw, d);
    COINN_CHECK_LAUNCH();
    return 0;
}
