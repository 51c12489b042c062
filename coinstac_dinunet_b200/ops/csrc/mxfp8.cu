// MX-FP8 (OCP microscaling: e4m3 elements, one ue8m0 scale per 32 elements along K) on the 5th-gen tensor cores.
// BASELINE config 4 / SURVEY §7.2 step 10: `tcgen05.mma.kind::mxf8f6f4.block_scale`, scale factors staged in TMEM.
//
//   quantize_mx_kernel   x[R, K] (bf16 / fp32) -> q[R, Kp] e4m3 bytes + sf[R, Kp/128] words (4 ue8m0 bytes, byte j =
//                        the j-th 32-element block of that 128-element group); Kp = K rounded up to 128, padded with zeros
//   gemm_mxfp8_kernel    C[M, N] = (A_q * 2^sfa)[M, K] . (B_q * 2^sfb)[N, K]^T, fp32 accumulation in TMEM
//
// GEMM structure (one CTA per 128 x 128 output tile, split-K over blockIdx.z):
//   warp 0      TMA producer: 128 x 128-byte A and B k-slices (one 128B-swizzle row = 128 e4m3) into a 4-stage ring
//   warp 1      MMA issuer:   per stage  tcgen05.cp (scale factors smem -> TMEM, 32x128b.warpx4) for A and B, then four
//                             tcgen05.mma.kind::mxf8f6f4.block_scale (M=128, N=128, K=32) whose instruction descriptor
//                             selects byte k of the scale words (a_sf_id / b_sf_id); tcgen05.commit frees the stage
//   warps 2-5   scale-factor loaders during the main loop (thread = tile row: one ld.global of the packed scale word,
//               stored in the 32x4x4 layout tcgen05.cp expects: word (row % 32) * 4 + row / 32), then the epilogue
//               (tcgen05.ld -> bias / ReLU -> global)
// The scale-factor TMEM columns are single-buffered: tcgen05.cp and tcgen05.mma execute in issue order.
#include "umma.cuh"
#include <cuda_fp8.h>

namespace coinn {

constexpr int MX_BM = 128, MX_BN = 128, MX_BK = 128, MX_STAGES = 4;
constexpr int MX_THREADS = 192;
constexpr int MX_A_BYTES = MX_BM * MX_BK, MX_B_BYTES = MX_BN * MX_BK, MX_SF_BYTES = 512;
constexpr int MX_STAGE_BYTES = MX_A_BYTES + MX_B_BYTES;
constexpr int MX_SMEM = MX_STAGES * (MX_STAGE_BYTES + 2 * MX_SF_BYTES) + 1024 + 256;
constexpr uint32_t MX_TMEM_COLS = 256;            // 128 accumulator columns + 4 (SFA) + 4 (SFB), power of two
constexpr uint32_t MX_SFA_COL = 128, MX_SFB_COL = 132;

// block-scaled instruction descriptor (kind::mxf8f6f4): b_sf_id [4,6) | a_format [7,10) | b_format [10,13) (0 = e4m3)
// | a_major 15 | b_major 16 (0 = K-major) | N>>3 [17,23) | scale_format 23 (1 = ue8m0) | M>>4 [24,29) | a_sf_id [29,31)
__host__ __device__ constexpr uint32_t make_idesc_mxf8(uint32_t M, uint32_t N, uint32_t a_sf, uint32_t b_sf) {
    return (b_sf << 4) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (a_sf << 29);
}

__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb) : "memory");
}
// 512 bytes (32 rows x 16 B, no swizzle) of packed scale words -> 4 TMEM columns, replicated to the four lane quarters
__device__ __forceinline__ void utccp_sf(uint32_t tmem_dst, uint32_t smem_addr) {
    const uint64_t desc = make_smem_desc(smem_addr, 0, 128, SMEM_LAYOUT_NONE);
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" :: "r"(tmem_dst), "l"(desc) : "memory");
}

struct MxGemmParams {
    void* C;
    const float* bias;
    const uint32_t* sfa;        // [M, kgroups] packed scale words
    const uint32_t* sfb;        // [N, kgroups]
    int M, N, K, ldc, kgroups;  // K multiple of 128; kgroups = K / 128
    int out_f32, relu, bias_mode, atomic_out, kblocks_per_split;
};

__global__ void __launch_bounds__(MX_THREADS, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const MxGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sf_smem = smem + MX_STAGES * MX_STAGE_BYTES;                    // [stage][A 512 B | B 512 B]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sf_smem + MX_STAGES * 2 * MX_SF_BYTES);
    uint64_t* empty_bar = full_bar + MX_STAGES;
    uint64_t* sf_bar = empty_bar + MX_STAGES;
    uint64_t* tmem_full_bar = sf_bar + MX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * MX_BM, n0 = blockIdx.x * MX_BN;
    const int kb_begin = blockIdx.z * p.kblocks_per_split;
    const int kb_end = min(p.kgroups, kb_begin + p.kblocks_per_split);
    const int num_kb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < MX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&sf_bar[s], 128); }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, MX_TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const bool leader = elect_one();
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % MX_STAGES;
            mbar_wait(&empty_bar[s], ((i / MX_STAGES) & 1) ^ 1);
            uint8_t* a_dst = smem + s * MX_STAGE_BYTES;
            if (leader) mbar_arrive_expect_tx(&full_bar[s], MX_STAGE_BYTES);
            const int k0 = (kb_begin + i) * MX_BK;
            if (leader) tma_load_2d(a_dst, &tmap_a, &full_bar[s], k0, m0);
            if (leader) tma_load_2d(a_dst + MX_A_BYTES, &tmap_b, &full_bar[s], k0, n0);
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % MX_STAGES;
            const uint32_t ph = (i / MX_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            mbar_wait(&sf_bar[s], ph);
            tcgen05_after_sync();
            const uint32_t a_addr = smem_u32(smem + s * MX_STAGE_BYTES);
            const uint32_t b_addr = a_addr + MX_A_BYTES;
            const uint32_t sf_addr = smem_u32(sf_smem + s * 2 * MX_SF_BYTES);
            if (leader) utccp_sf(tmem_base + MX_SFA_COL, sf_addr);
            if (leader) utccp_sf(tmem_base + MX_SFB_COL, sf_addr + MX_SF_BYTES);
#pragma unroll
            for (int k = 0; k < MX_BK / 32; ++k) {
                const uint64_t adesc = make_smem_desc(a_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128);
                const uint64_t bdesc = make_smem_desc(b_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128);
                if (leader) umma_mxf8(tmem_base, adesc, bdesc, make_idesc_mxf8(MX_BM, MX_BN, k, k), (i > 0 || k > 0) ? 1u : 0u,
                                      tmem_base + MX_SFA_COL, tmem_base + MX_SFB_COL);
            }
            if (leader) umma_commit(&empty_bar[s]);
        }
        if (leader) umma_commit(tmem_full_bar);
    } else {
        // ---- scale-factor loaders (thread = tile row) ----
        const int t = threadIdx.x - 64;                     // 0..127
        const int slot = (t & 31) * 4 + (t >> 5);           // word index inside the 512-byte UTCCP block
        const int arow = m0 + t, brow = n0 + t;
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % MX_STAGES;
            mbar_wait(&empty_bar[s], ((i / MX_STAGES) & 1) ^ 1);
            const int g = kb_begin + i;
            const uint32_t wa = arow < p.M ? __ldg(p.sfa + (long long)arow * p.kgroups + g) : 0x7f7f7f7fu;
            const uint32_t wb = brow < p.N ? __ldg(p.sfb + (long long)brow * p.kgroups + g) : 0x7f7f7f7fu;
            uint32_t* dst = reinterpret_cast<uint32_t*>(sf_smem + s * 2 * MX_SF_BYTES);
            dst[slot] = wa;
            dst[128 + slot] = wb;
            fence_proxy_async_smem();                       // generic-proxy stores -> visible to tcgen05.cp (async proxy)
            mbar_arrive(&sf_bar[s]);
        }
        // ---- epilogue ----
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(tmem_full_bar, 0);
            tcgen05_after_sync();
        }
        const float row_bias = (p.bias_mode == 2 && row < p.M) ? p.bias[row] : 0.f;
#pragma unroll 1
        for (int c = 0; c < MX_BN; c += 16) {
            uint32_t r[16];
            if (num_kb > 0) {
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            const int col0 = n0 + c;
            if (row >= p.M || col0 >= p.N) continue;
            const size_t base = (size_t)row * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (col0 + j >= p.N) break;
                float x = __uint_as_float(r[j]) + row_bias;
                if (p.bias_mode == 1) x += p.bias[col0 + j];
                if (p.relu) x = fmaxf(x, 0.f);
                if (p.atomic_out) atomicAdd(reinterpret_cast<float*>(p.C) + base + j, x);
                else if (p.out_f32) reinterpret_cast<float*>(p.C)[base + j] = x;
                else reinterpret_cast<__nv_bfloat16*>(p.C)[base + j] = __float2bfloat16_rn(x);
            }
        }
    }
    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, MX_TMEM_COLS);
}

static int make_tmap_2d_u8(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                           uint32_t box_rows, uint32_t box_cols) {
    auto enc = get_tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return (int)enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// ------------------------------------------------------------------------------------------------ quantisation
template <typename TIn> __device__ __forceinline__ float mx_load(const TIn* p);
template <> __device__ __forceinline__ float mx_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float mx_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// one thread per 32-element block: amax -> ue8m0 exponent e (scale 2^(e-127) >= amax / 448), q = sat_e4m3(x / scale)
template <typename TIn>
__global__ void __launch_bounds__(256) quantize_mx_kernel(const TIn* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf,
                                                          long long R, int K, int Kp) {
    const int blocks_per_row = Kp / 32;
    const long long total = R * blocks_per_row;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (long long)gridDim.x * blockDim.x) {
        const long long row = b / blocks_per_row;
        const int kb = (int)(b % blocks_per_row), k0 = kb * 32;
        float v[32];
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            v[j] = (k0 + j < K) ? mx_load(x + row * K + k0 + j) : 0.f;
            amax = fmaxf(amax, fabsf(v[j]));
        }
        int e = 127;
        if (amax > 0.f && isfinite(amax)) {
            int ex;
            const float m = frexpf(amax / 448.f, &ex);          // amax / 448 = m * 2^ex, m in [0.5, 1)
            e = ex + 127 - (m == 0.5f ? 1 : 0);                 // smallest power of two >= amax / 448
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
        uint4* dst = reinterpret_cast<uint4*>(q + row * Kp + k0);
        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        sf[row * blocks_per_row + kb] = (uint8_t)e;              // byte (kb % 4) of word kb / 4: little endian packing
    }
}

}  // namespace coinn

using namespace coinn;

// x: [R, K] fp32 (dtype 0) or bf16 (1), contiguous -> q: [R, Kp] bytes, sf: [R, Kp/32] bytes (= [R, Kp/128] words), Kp % 128 == 0
COINN_API int coinn_quantize_mx(const void* x, int dtype, void* q, void* sf, long long R, int K, int Kp, void* stream) {
    if (Kp % 128 || Kp < K) return (int)cudaErrorInvalidValue;
    if (R == 0) return 0;
    const long long total = R * (Kp / 32);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == 0) quantize_mx_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, (uint8_t*)q, (uint8_t*)sf, R, K, Kp);
    else quantize_mx_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (uint8_t*)q, (uint8_t*)sf, R, K, Kp);
    COINN_CHECK_LAUNCH();
    return 0;
}

// C[M,N] = dequant(A)[M,K] . dequant(B)[N,K]^T.  A, B: e4m3 bytes, rows of K bytes (K % 128 == 0, 16-byte aligned);
// sfa / sfb: packed ue8m0 words [rows, K/128].  out_dtype 0 bf16 / 1 fp32; bias_mode 0 none / 1 along N / 2 along M;
// split_k > 1: C must be zeroed fp32, accumulated atomically (bias / relu ignored).
COINN_API int coinn_gemm_mxfp8_tn(const void* A, const void* sfa, const void* B, const void* sfb, void* C, const float* bias,
                                  int M, int N, int K, int ldc, int out_dtype, int relu, int bias_mode, int split_k, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if (K % 128 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return (int)cudaErrorInvalidValue;
    MxGemmParams p;
    p.C = C; p.bias = bias; p.sfa = (const uint32_t*)sfa; p.sfb = (const uint32_t*)sfb;
    p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.kgroups = K / 128;
    p.out_f32 = out_dtype; p.relu = relu; p.bias_mode = bias ? bias_mode : 0;
    if (split_k < 1) split_k = 1;
    if (split_k > p.kgroups) split_k = p.kgroups;
    p.kblocks_per_split = (p.kgroups + split_k - 1) / split_k;
    split_k = (p.kgroups + p.kblocks_per_split - 1) / p.kblocks_per_split;
    p.atomic_out = split_k > 1;
    if (p.atomic_out) { p.out_f32 = 1; p.relu = 0; p.bias_mode = 0; }
    CUtensorMap ta, tb;
    if (make_tmap_2d_u8(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, MX_BM, MX_BK) != 0) return -2;
    if (make_tmap_2d_u8(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)K, MX_BN, MX_BK) != 0) return -3;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((N + MX_BN - 1) / MX_BN, (M + MX_BM - 1) / MX_BM, split_k);
    gemm_mxfp8_kernel<<<grid, MX_THREADS, MX_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(ta, tb, p);
    COINN_CHECK_LAUNCH();
    return 0;
}
