// bf16 GEMM on the 5th-gen tensor cores:  C[M,N] = A[M,K] · B[N,K]^T (+bias)(ReLU), fp32 accumulate.
// Used by the Linear layers (forward / dgrad / wgrad of FSNet and of the VBMNet head).
//
// Structure (one CTA per 128 x BLOCK_N output tile, optional split-K over blockIdx.z):
//   warp 0   TMA producer   cp.async.bulk.tensor.2d of the A and B k-slices (64 bf16 = one 128B swizzle row)
//                           into a STAGES-deep shared-memory ring, completion on `full` mbarriers
//   warp 1   MMA issuer     one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16)
//                           four per k-slice, accumulators live in TMEM; tcgen05.commit frees the smem slot
//                           (`empty` mbarrier) and finally signals the epilogue (`tmem_full`)
//   warps 2-5 epilogue      tcgen05.ld 32x32b.x16 (lane == accumulator row) -> bias / ReLU / cast -> global
// Both operands are K-major (row-major [rows, K]); M, N and K tails are handled by TMA out-of-bounds zero
// fill on the loads and by predication on the stores.
#include "umma.cuh"

namespace coinn {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;            // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int GEMM_THREADS = 192;

struct GemmParams {
    void* C;
    const float* bias;
    int M, N, K, ldc;
    int out_f32;        // 0: bf16 output, 1: fp32 output
    int relu;
    int bias_mode;      // 0 none, 1 per column (N), 2 per row (M)
    int atomic_out;     // split-K: fp32 atomicAdd into C (must be zeroed; bias/relu must be off)
    int kblocks_per_split;
};

template <int BLOCK_N>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BLOCK_N * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    __host__ __device__ static constexpr int stages() { return BLOCK_N <= 64 ? 6 : (BLOCK_N <= 128 ? 5 : 4); }
    __host__ __device__ static constexpr int total() { return stages() * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/; }
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using S = GemmSmem<BLOCK_N>;
    constexpr int STAGES = S::stages();
    constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * BLOCK_N;
    const int total_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int kb_begin = blockIdx.z * p.kblocks_per_split;
    const int kb_end = min(total_kb, kb_begin + p.kblocks_per_split);
    const int num_kb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                uint8_t* b_dst = a_dst + S::A_BYTES;
                if (leader) mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
                const int k0 = (kb_begin + i) * GEMM_BK;
                if (leader) tma_load_2d(a_dst, &tmap_a, &full_bar[s], k0, m0);
                if (leader) tma_load_2d(b_dst, &tmap_b, &full_bar[s], k0, n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            const bool leader = elect_one();   // whole warp runs the loop (uniform descriptors), one lane issues
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BLOCK_N, 1, 0, 0);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_after_sync();
                const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    const uint64_t adesc = make_smem_desc(a_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128);
                    const uint64_t bdesc = make_smem_desc(b_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128);
                    if (leader) umma_f16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                if (leader) umma_commit(&empty_bar[s]);          // smem slot reusable once these MMAs retire
            }
            if (leader) umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(tmem_full_bar, 0);
            tcgen05_after_sync();
        }
        const float row_bias = (p.bias_mode == 2 && row < p.M) ? p.bias[row] : 0.f;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 16) {
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
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(r[j]) + row_bias;
                if (p.bias_mode == 1 && col0 + j < p.N) x += p.bias[col0 + j];
                if (p.relu) x = fmaxf(x, 0.f);
                v[j] = x;
            }
            const size_t base = (size_t)row * p.ldc + col0;
            const bool full = (col0 + 16 <= p.N);
            if (p.atomic_out) {
                float* C = reinterpret_cast<float*>(p.C);
#pragma unroll
                for (int j = 0; j < 16; ++j) if (col0 + j < p.N) atomicAdd(C + base + j, v[j]);
            } else if (p.out_f32) {
                float* C = reinterpret_cast<float*>(p.C);
                if (full && ((base & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(C + base + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (col0 + j < p.N) C[base + j] = v[j];
                }
            } else {
                __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C);
                if (full && ((base & 7) == 0)) {
                    uint4 lo = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    uint4 hi = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                    *reinterpret_cast<uint4*>(C + base) = lo;
                    *reinterpret_cast<uint4*>(C + base + 8) = hi;
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (col0 + j < p.N) C[base + j] = __float2bfloat16_rn(v[j]);
                }
            }
        }
    }

    tcgen05_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BLOCK_N>
static int launch_gemm(const void* A, const void* B, const GemmParams& p, int lda, int ldb, int split_k, cudaStream_t st) {
    CUtensorMap ta, tb;
    if (make_tmap_2d_bf16(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda * 2, GEMM_BM, GEMM_BK) != 0) return -2;
    if (make_tmap_2d_bf16(&tb, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb * 2, BLOCK_N, GEMM_BK) != 0) return -3;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             GemmSmem<BLOCK_N>::total());
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((p.N + BLOCK_N - 1) / BLOCK_N, (p.M + GEMM_BM - 1) / GEMM_BM, split_k);
    gemm_bf16_tn_kernel<BLOCK_N><<<grid, GEMM_THREADS, GemmSmem<BLOCK_N>::total(), st>>>(ta, tb, p);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// C[M,N] = A[M,K] · B[N,K]^T.  lda/ldb/ldc in elements; A, B bf16 with 16-byte aligned rows (ld % 8 == 0).
// out_dtype: 0 bf16, 1 fp32.  bias_mode: 0 none, 1 along N, 2 along M.  split_k > 1 => C must be zeroed fp32
// and is accumulated atomically (bias/relu ignored).
COINN_API int coinn_gemm_bf16_tn(const void* A, const void* B, void* C, const float* bias, int M, int N, int K,
                                 int lda, int ldb, int ldc, int out_dtype, int relu, int bias_mode, int split_k,
                                 void* stream) {
    using namespace coinn;
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if ((lda & 7) || (ldb & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return (int)cudaErrorMisalignedAddress;
    GemmParams p;
    p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc;
    p.out_f32 = out_dtype; p.relu = relu; p.bias_mode = bias ? bias_mode : 0;
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    if (split_k < 1) split_k = 1;
    if (split_k > total_kb) split_k = total_kb;
    p.kblocks_per_split = (total_kb + split_k - 1) / split_k;
    split_k = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
    p.atomic_out = split_k > 1;
    if (p.atomic_out) { p.out_f32 = 1; p.relu = 0; p.bias_mode = 0; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (N <= 16) return launch_gemm<16>(A, B, p, lda, ldb, split_k, st);
    if (N <= 32) return launch_gemm<32>(A, B, p, lda, ldb, split_k, st);
    if (N <= 64) return launch_gemm<64>(A, B, p, lda, ldb, split_k, st);
    // 128 x 256 tiles double the MMA work per byte of shared-memory operand traffic (the N=128 kernel sits at 44 % tensor
    // pipe, bound by operand fetch); use them when they still fill the machine
    const long long tiles256 = (long long)((M + GEMM_BM - 1) / GEMM_BM) * ((N + 255) / 256) * split_k;
    if (N >= 256 && tiles256 >= B200_SM_COUNT) return launch_gemm<256>(A, B, p, lda, ldb, split_k, st);
    return launch_gemm<128>(A, B, p, lda, ldb, split_k, st);
}
