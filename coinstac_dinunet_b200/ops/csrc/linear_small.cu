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
//
// Linear + BatchNorm1d + ReLU (the FreeSurfer MLP's hidden layers, SURVEY K4 "BN1d+ReLU epilogue"): at M <= 32 a CTA that owns
// an output feature owns its whole batch column, so training-mode BatchNorm needs NO cross-CTA reduction:
//   forward : y = x W^T + b ; mean / biased var over the M rows by warp shuffles ; xhat stored ; z = relu(gamma xhat + beta) ;
//             running statistics (unbiased var, momentum) and num_batches_tracked updated in the same launch
//   backward: the dy staging loop first turns dz into dy = gamma invstd (g - mean(g) - xhat mean(g xhat)), g = dz [z > 0],
//             and accumulates dgamma / dbeta in place; the rest (dW, db, dx) is the plain kernel.
#include "common.cuh"

namespace coinn {

constexpr int LS_MAX_M = 32;

__device__ __forceinline__ float ls_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ls_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// ---------------------------------------------------------------------------------------------------- forward
struct BnFwd {                 // all null / zero for a plain Linear
    const float* gamma; const float* beta;
    float* running_mean; float* running_var; long long* num_batches_tracked;
    float* xhat;               // [M, N] saved for backward (training)
    float* invstd;             // [N]    saved for backward (training)
    float eps, momentum;
    int training;              // 1: batch statistics (+ running update); 0: running statistics
};

template <typename TX, int MB>
__global__ void __launch_bounds__(256) linear_small_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ W,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               int M, int N, int K, int relu, BnFwd bn) {
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
    if (threadIdx.x < 32) {                    // warp 0: lane m owns row m of this feature's column (MB <= 32)
        const bool live = threadIdx.x < M;
        float v = 0.f;
        if (live) {
            v = bias ? bias[n] : 0.f;
#pragma unroll
            for (int wq = 0; wq < 8; ++wq) v += s_red[wq][threadIdx.x < MB ? threadIdx.x : 0];
        }
        if (bn.gamma) {
            float mean, invstd;
            if (bn.training) {
                mean = warp_sum(live ? v : 0.f) / (float)M;
                const float d = live ? v - mean : 0.f;
                const float ss = warp_sum(d * d);
                const float var = ss / (float)M;
                invstd = 1.0f / sqrtf(var + bn.eps);
                if (threadIdx.x == 0) {
                    if (bn.invstd) bn.invstd[n] = invstd;
                    if (bn.running_mean) {
                        const float unbiased = M > 1 ? ss / (float)(M - 1) : var;
                        bn.running_mean[n] = (1.f - bn.momentum) * bn.running_mean[n] + bn.momentum * mean;
                        bn.running_var[n] = (1.f - bn.momentum) * bn.running_var[n] + bn.momentum * unbiased;
                    }
                    if (bn.num_batches_tracked && n == 0) *bn.num_batches_tracked += 1;
                }
            } else {
                mean = bn.running_mean[n];
                invstd = 1.0f / sqrtf(bn.running_var[n] + bn.eps);
            }
            const float xh = (v - mean) * invstd;
            if (live && bn.xhat) bn.xhat[(long long)threadIdx.x * N + n] = xh;
            v = fmaf(bn.gamma[n], xh, bn.beta[n]);
        }
        if (live) y[(long long)threadIdx.x * N + n] = relu ? fmaxf(v, 0.f) : v;
    }
}

// --------------------------------------------------------------------------------------------------- backward
constexpr int LS_NC = 16;             // output features per CTA (grid.y chunks)

struct BnBwd {                 // gamma == null: plain Linear (+ optional ReLU mask)
    const float* gamma; const float* beta; const float* xhat; const float* invstd;
    float* dgamma; float* dbeta;          // accumulated in place (one writer per feature: the blockIdx.x == 0 CTAs)
    int relu;
};

template <typename TX, int MB>
__global__ void __launch_bounds__(256) linear_small_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y_mask,
                                                               const TX* __restrict__ x, const float* __restrict__ W,
                                                               float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dx,
                                                               int M, int N, int K, BnBwd bn) {
    __shared__ float s_dy[LS_NC][MB];
    const int n0 = blockIdx.y * LS_NC;
    const int nc = (N - n0) < LS_NC ? (N - n0) : LS_NC;
    if (bn.gamma) {
        // warp w turns dz into dy for features j = w, w + 8 (lane = batch row): BatchNorm1d backward is column-local
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int j = warp; j < LS_NC; j += 8) {
            float out = 0.f;
            if (j < nc) {                                                  // warp-uniform
                const int n = n0 + j;
                const bool live = lane < M;
                const float xh = live ? bn.xhat[(long long)lane * N + n] : 0.f;
                const float ga = bn.gamma[n];
                float g = live ? dy[(long long)lane * N + n] : 0.f;
                if (bn.relu && !(fmaf(ga, xh, bn.beta[n]) > 0.f)) g = 0.f;
                const float sg = warp_sum(g), sgx = warp_sum(g * xh);
                if (lane == 0 && blockIdx.x == 0) {
                    if (bn.dbeta) bn.dbeta[n] += sg;
                    if (bn.dgamma) bn.dgamma[n] += sgx;
                }
                const float inv_m = 1.f / (float)M;
                out = live ? ga * bn.invstd[n] * (g - sg * inv_m - xh * sgx * inv_m) : 0.f;
            }
            if (lane < MB) s_dy[j][lane] = out;
        }
    } else {
        for (int i = threadIdx.x; i < LS_NC * MB; i += blockDim.x) {
            const int j = i / MB, m = i % MB;
            float v = 0.f;
            if (j < nc && m < M) {
                v = dy[(long long)m * N + n0 + j];
                if (y_mask && !(y_mask[(long long)m * N + n0 + j] > 0.f)) v = 0.f;
            }
            s_dy[j][m] = v;
        }
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
static int ls_fwd(const void* x, const float* W, const float* bias, float* y, int M, int N, int K, int relu, const BnFwd& bn,
                  cudaStream_t st) {
    if (M <= 8) linear_small_fwd_kernel<TX, 8><<<N, 256, 0, st>>>((const TX*)x, W, bias, y, M, N, K, relu, bn);
    else if (M <= 16) linear_small_fwd_kernel<TX, 16><<<N, 256, 0, st>>>((const TX*)x, W, bias, y, M, N, K, relu, bn);
    else linear_small_fwd_kernel<TX, 32><<<N, 256, 0, st>>>((const TX*)x, W, bias, y, M, N, K, relu, bn);
    COINN_CHECK_LAUNCH();
    return 0;
}

template <typename TX>
static int ls_bwd(const float* dy, const float* ym, const void* x, const float* W, float* dW, float* db, float* dx, int M, int N, int K,
                  const BnBwd& bn, cudaStream_t st) {
    const dim3 grid((K + 255) / 256, (N + LS_NC - 1) / LS_NC);
    if (M <= 8) linear_small_bwd_kernel<TX, 8><<<grid, 256, 0, st>>>(dy, ym, (const TX*)x, W, dW, db, dx, M, N, K, bn);
    else if (M <= 16) linear_small_bwd_kernel<TX, 16><<<grid, 256, 0, st>>>(dy, ym, (const TX*)x, W, dW, db, dx, M, N, K, bn);
    else linear_small_bwd_kernel<TX, 32><<<grid, 256, 0, st>>>(dy, ym, (const TX*)x, W, dW, db, dx, M, N, K, bn);
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
    const BnFwd none{};
    return x_dtype == 0 ? ls_fwd<float>(x, W, bias, y, M, N, K, relu, none, st)
                        : ls_fwd<__nv_bfloat16>(x, W, bias, y, M, N, K, relu, none, st);
}

// Linear + BatchNorm1d (+ ReLU) forward, M <= 32.  training = 1: batch statistics, xhat [M,N] and invstd [N] are written
// for the backward pass, running_mean / running_var / num_batches_tracked (int64, may be null) are updated;
// training = 0: running statistics, nothing saved.
COINN_API int coinn_linear_bn_small_fwd(const void* x, int x_dtype, const float* W, const float* bias, const float* gamma,
                                        const float* beta, float* running_mean, float* running_var, long long* nbt,
                                        float* xhat, float* invstd, float* z, int M, int N, int K, int relu, int training,
                                        float eps, float momentum, void* stream) {
    using namespace coinn;
    if (M > LS_MAX_M || M < 1 || !gamma || !beta) return -1;
    if (!training && (!running_mean || !running_var)) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    BnFwd bn{gamma, beta, running_mean, running_var, nbt, xhat, invstd, eps, momentum, training};
    return x_dtype == 0 ? ls_fwd<float>(x, W, bias, z, M, N, K, relu, bn, st)
                        : ls_fwd<__nv_bfloat16>(x, W, bias, z, M, N, K, relu, bn, st);
}

// dy: [M,N] fp32; y_mask: [M,N] fp32 forward output (ReLU mask) or null; dW [N,K] and db [N] are ACCUMULATED into;
// dx: [M,K] fp32 zeroed by the caller (accumulated with atomics) or null.
COINN_API int coinn_linear_small_bwd(const float* dy, const float* y_mask, const void* x, int x_dtype, const float* W, float* dW, float* db,
                                     float* dx, int M, int N, int K, void* stream) {
    using namespace coinn;
    if (M > LS_MAX_M || M < 1) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const BnBwd none{};
    return x_dtype == 0 ? ls_bwd<float>(dy, y_mask, x, W, dW, db, dx, M, N, K, none, st)
                        : ls_bwd<__nv_bfloat16>(dy, y_mask, x, W, dW, db, dx, M, N, K, none, st);
}

// Backward of coinn_linear_bn_small_fwd (training mode).  dz: [M,N] gradient of the block output; xhat / invstd from the
// forward; dW, db, dgamma, dbeta are ACCUMULATED into; dx: zeroed [M,K] (atomics) or null.
COINN_API int coinn_linear_bn_small_bwd(const float* dz, const float* xhat, const float* invstd, const float* gamma, const float* beta,
                                        const void* x, int x_dtype, const float* W, float* dW, float* db, float* dgamma, float* dbeta,
                                        float* dx, int M, int N, int K, int relu, void* stream) {
    using namespace coinn;
    if (M > LS_MAX_M || M < 1 || !gamma || !xhat || !invstd) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    BnBwd bn{gamma, beta, xhat, invstd, dgamma, dbeta, relu};
    return x_dtype == 0 ? ls_bwd<float>(dz, nullptr, x, W, dW, db, dx, M, N, K, bn, st)
                        : ls_bwd<__nv_bfloat16>(dz, nullptr, x, W, dW, db, dx, M, N, K, bn, st);
}
