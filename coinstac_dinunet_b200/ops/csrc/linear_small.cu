// Linear layers at federated-site batch sizes (M <= 32 rows).
//
// The classifier heads of both reference models run at batch 8-16: a [8 x 9216] x [9216 x 256] product has no use for
// a 128-row tensor-core tile, and expressing it through the GEMM path costs ~15 launches per layer (casts, transposes,
// split-K zero fills, mask, bias sum, .grad accumulation) - ~70 launches and ~8 % of a VBM step for ~20 MFLOP.
// Two CUDA-core kernels per layer instead, both bound by streaming the fp32 weight matrix once:
//   forward : one CTA per output feature, all M rows accumulated together           y = act(x W^T + b)
//   backward: one thread per input feature k and a chunk of output features n:
//             dW[n,k] += sum_m dy[m,n] x[m,k]   (accumulated IN PLACE into the parameter's .grad - no temporary)
//             db[n]   += sum_m dy[m,n]
//             dx[m,k] += sum_n dy[m,n] W[n,k]   (fp32 atomics across the n-chunks)
//             with the ReLU mask (y > 0) applied while dy is staged in shared memory.
#include "common.cuh"

namespace coinn {

constexpr int LS_MAX_M = 32;

__device__ __forceinline__ float ls_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ls_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// ---------------------------------------------------------------------------------------------------- forward
template <typename TX, int MB>
__global__ void __launch_bounds__(256) linear_small_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ W,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               int M, int N, int K, int relu) {
    // one CTA per output feature: 256 threads stride over K (36 iterations for the 9216-wide VBM head)
    __shared__ float s_red[8][MB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x;
    float acc[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) acc[m] = 0.f;
    const float* w = W + (long long)n * K;
#pragma unroll 2
    for (int k = threadIdx.x; k < K; k += 256) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int m = 0; m < MB; ++m)
            if (m < M) acc[m] = fmaf(ls_ld(x + (long long)m * K + k), wv, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) {
        const float s = warp_sum(acc[m]);
        if (lane == 0) s_red[warp][m] = s;
    }
    __syncthreads();
    if (threadIdx.x < M) {
        float v = bias ? bias[n] : 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) v += s_red[wq][threadIdx.x];
        y[(long long)threadIdx.x * N + n] = relu ? fmaxf(v, 0.f) : v;
    }
}

// --------------------------------------------------------------------------------------------------- backward
constexpr int LS_NC = 16;             // output features per CTA (grid.y chunks)

template <typename TX, int MB>
__global__ void __launch_bounds__(256) linear_small_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y_mask,
                                                               const TX* __restrict__ x, const float* __restrict__ W,
                                                               float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dx,
                                                               int M, int N, int K) {
    __shared__ float s_dy[LS_NC][MB];
    const int n0 = blockIdx.y * LS_NC;
    const int nc = (N - n0) < LS_NC ? (N - n0) : LS_NC;
    for (int i = threadIdx.x; i < LS_NC * MB; i += blockDim.x) {
        const int j = i / MB, m = i % MB;
        float v = 0.f;
        if (j < nc && m < M) {
            v = dy[(long long)m * N + n0 + j];
            if (y_mask && !(y_mask[(long long)m * N + n0 + j] > 0.f)) v = 0.f;
        }
        s_dy[j][m] = v;
    }
    __syncthreads();
    if (db && blockIdx.x == 0 && threadIdx.x < nc) {
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < MB; ++m) s += s_dy[threadIdx.x][m];
        db[n0 + threadIdx.x] += s;                          // one writer per n (grid.x == 0 only)
    }
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float xv[MB], dxa[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) { xv[m] = m < M ? ls_ld(x + (long long)m * K + k) : 0.f; dxa[m] = 0.f; }
#pragma unroll 4
    for (int j = 0; j < nc; ++j) {
        const long long off = (long long)(n0 + j) * K + k;
        const float wv = dx ? __ldg(W + off) : 0.f;
        float g = 0.f;
#pragma unroll
        for (int m = 0; m < MB; ++m) {
            const float d = s_dy[j][m];
            g = fmaf(d, xv[m], g);
            dxa[m] = fmaf(d, wv, dxa[m]);
        }
        dW[off] += g;                                       // each (n, k) belongs to exactly one thread
    }
    if (dx) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
            if (m < M) atomicAdd(dx + (long long)m * K + k, dxa[m]);
    }
}

template <typename TX>
static int ls_fwd(const void* x, const float* W, const float* bias, float* y, int M, int N, int K, int relu, cudaStream_t st) {
    if (M <= 8) linear_small_fwd_kernel<TX, 8><<<N, 256, 0, st>>>((const TX*)x, W, bias, y, M, N, K, relu);
    else if (M <= 16) linear_small_fwd_kernel<TX, 16><<<N, 256, 0, st>>>((const TX*)x, W, bias, y, M, N, K, relu);
    else linear_small_fwd_kernel<TX, 32><<<N, 256, 0, st>>>((const TX*)x, W, bias, y, M, N, K, relu);
    COINN_CHECK_LAUNCH();
    return 0;
}

template <typename TX>
static int ls_bwd(const float* dy, const float* ym, const void* x, const float* W, float* dW, float* db, float* dx, int M, int N, int K,
                  cudaStream_t st) {
    const dim3 grid((K + 255) / 256, (N + LS_NC - 1) / LS_NC);
    if (M <= 8) linear_small_bwd_kernel<TX, 8><<<grid, 256, 0, st>>>(dy, ym, (const TX*)x, W, dW, db, dx, M, N, K);
    else if (M <= 16) linear_small_bwd_kernel<TX, 16><<<grid, 256, 0, st>>>(dy, ym, (const TX*)x, W, dW, db, dx, M, N, K);
    else linear_small_bwd_kernel<TX, 32><<<grid, 256, 0, st>>>(dy, ym, (const TX*)x, W, dW, db, dx, M, N, K);
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// x: [M,K] fp32 (x_dtype 0) or bf16 (1); W: [N,K] fp32; bias: [N] or null; y: [M,N] fp32.  M <= 32.
COINN_API int coinn_linear_small_fwd(const void* x, int x_dtype, const float* W, const float* bias, float* y, int M, int N, int K, int relu,
                                     void* stream) {
    using namespace coinn;
    if (M > LS_MAX_M || M < 1) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return x_dtype == 0 ? ls_fwd<float>(x, W, bias, y, M, N, K, relu, st) : ls_fwd<__nv_bfloat16>(x, W, bias, y, M, N, K, relu, st);
}

// dy: [M,N] fp32; y_mask: [M,N] fp32 forward output (ReLU mask) or null; dW [N,K] and db [N] are ACCUMULATED into;
// dx: [M,K] fp32 zeroed by the caller (accumulated with atomics) or null.
COINN_API int coinn_linear_small_bwd(const float* dy, const float* y_mask, const void* x, int x_dtype, const float* W, float* dW, float* db,
                                     float* dx, int M, int N, int K, void* stream) {
    using namespace coinn;
    if (M > LS_MAX_M || M < 1) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return x_dtype == 0 ? ls_bwd<float>(dy, y_mask, x, W, dW, db, dx, M, N, K, st)
                        : ls_bwd<__nv_bfloat16>(dy, y_mask, x, W, dW, db, dx, M, N, K, st);
}
