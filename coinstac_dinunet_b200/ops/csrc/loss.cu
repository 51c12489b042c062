// Fused log-softmax + NLL (mean) + argmax, forward and backward.
// The reference user code runs log_softmax, nll_loss, max as three ATen ops and then reads the loss
// with .item() (README.md:73-84).  Here one warp handles one row of logits [N, C]; the loss is
// accumulated on the device (atomicAdd of row losses pre-divided by N), predictions come out of the
// same pass, and the probabilities are kept for a one-pass backward:  dlogits = (p - onehot) * g / N.
#include "common.cuh"

namespace coinn {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void softmax_nll_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ labels,
                                       float* __restrict__ probs, long long* __restrict__ pred,
                                       float* __restrict__ loss, int N, int C) {
    const int warps_per_block = blockDim.x >> 5;
    const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (row >= N) return;
    const int lane = lane_id();
    const T* x = logits + (size_t)row * C;

    float mx = -INFINITY; int arg = 0;
    for (int c = lane; c < C; c += 32) {
        const float v = to_f(x[c]);
        if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {          // (max, first index) reduction
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > mx || (ov == mx && oa < arg)) { mx = ov; arg = oa; }
    }
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += __expf(to_f(x[c]) - mx);
    sum = warp_sum(sum);
    const float lse = mx + __logf(sum);
    const float inv = 1.f / sum;
    float* pr = probs + (size_t)row * C;
    for (int c = lane; c < C; c += 32) pr[c] = __expf(to_f(x[c]) - mx) * inv;
    if (lane == 0) {
        const long long y = labels[row];
        pred[row] = arg;
        if (y >= 0 && y < C) atomicAdd(loss, (lse - to_f(x[y])) / (float)N);
    }
}

template <typename T>
__global__ void softmax_nll_bwd_kernel(const float* __restrict__ probs, const long long* __restrict__ labels,
                                       const float* __restrict__ gloss, T* __restrict__ dlogits, int N, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * C) return;
    const int row = (int)(i / C), c = (int)(i % C);
    const float g = gloss[0] / (float)N;
    const float onehot = (labels[row] == c) ? 1.f : 0.f;
    dlogits[i] = from_f<T>((probs[i] - onehot) * g);
}

}  // namespace coinn

// dtype: 0 f32, 1 bf16, 2 f16.  `loss` must be zeroed by the caller (a 4-byte memset node).
COINN_API int coinn_softmax_nll_fwd(const void* logits, const long long* labels, float* probs, long long* pred,
                                    float* loss, int N, int C, int dtype, void* stream) {
    using namespace coinn;
    if (N == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int threads = 128, rows_per_block = threads / 32;
    const int grid = (N + rows_per_block - 1) / rows_per_block;
    if (dtype == 0) softmax_nll_fwd_kernel<float><<<grid, threads, 0, st>>>((const float*)logits, labels, probs, pred, loss, N, C);
    else if (dtype == 1) softmax_nll_fwd_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>((const __nv_bfloat16*)logits, labels, probs, pred, loss, N, C);
    else softmax_nll_fwd_kernel<__half><<<grid, threads, 0, st>>>((const __half*)logits, labels, probs, pred, loss, N, C);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_softmax_nll_bwd(const float* probs, const long long* labels, const float* gloss, void* dlogits,
                                    int N, int C, int dtype, void* stream) {
    using namespace coinn;
    if (N == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)N * C;
    const int threads = 256;
    const int grid = (int)((total + threads - 1) / threads);
    if (dtype == 0) softmax_nll_bwd_kernel<float><<<grid, threads, 0, st>>>(probs, labels, gloss, (float*)dlogits, N, C);
    else if (dtype == 1) softmax_nll_bwd_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(probs, labels, gloss, (__nv_bfloat16*)dlogits, N, C);
    else softmax_nll_bwd_kernel<__half><<<grid, threads, 0, st>>>(probs, labels, gloss, (__half*)dlogits, N, C);
    COINN_CHECK_LAUNCH();
    return 0;
}
