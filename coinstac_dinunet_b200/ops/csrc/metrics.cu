// Device-side metric accumulators: no host sync per batch.
//   count_binary     counter[4] (tn, fp, fn, tp) += histogram of 2*true + pred   (255 -> 1 for 8-bit masks)
//   count_confusion  matrix[C*C] [pred*C + true] += 1
// The reference does 4 compare+sum launches and 4 .item() syncs per Prf1a.add (metrics.py:158-170)
// and a sparse->dense add on the CPU for the confusion matrix (metrics.py:243-249).
#include "common.cuh"

namespace coinn {

template <typename TP, typename TT>
__global__ void count_binary_kernel(const TP* __restrict__ pred, const TT* __restrict__ truth,
                                    unsigned long long* __restrict__ counter, long long n) {
    unsigned int c[4] = {0u, 0u, 0u, 0u};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long p = (long long)pred[i], t = (long long)truth[i];
        if (p == 255) p = 1;
        if (t == 255) t = 1;
        const long long code = 2 * t + p;
        if (code >= 0 && code <= 3) c[code]++;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned int v = c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane_id() == 0 && v) atomicAdd(counter + k, (unsigned long long)v);
    }
}

template <typename TP, typename TT>
__global__ void count_confusion_kernel(const TP* __restrict__ pred, const TT* __restrict__ truth,
                                       unsigned long long* __restrict__ matrix, long long n, int C) {
    extern __shared__ unsigned int hist[];
    const int bins = C * C;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long p = (long long)pred[i], t = (long long)truth[i];
        if (p >= 0 && p < C && t >= 0 && t < C) atomicAdd(&hist[p * C + t], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (hist[i]) atomicAdd(matrix + i, (unsigned long long)hist[i]);
}

template <typename TP>
static int dispatch_truth(const void* pred, const void* truth, int truth_dtype, unsigned long long* out, long long n,
                          int C, bool confusion, cudaStream_t st) {
    const int threads = 256;
    long long want = (n + threads * 8 - 1) / (threads * 8);
    const int grid = (int)(want < 1 ? 1 : (want > 4 * B200_SM_COUNT ? 4 * B200_SM_COUNT : want));
#define LAUNCH(TT)                                                                                         \
    if (confusion) count_confusion_kernel<TP, TT><<<grid, threads, (size_t)C * C * sizeof(unsigned int), st>>>( \
        (const TP*)pred, (const TT*)truth, out, n, C);                                                     \
    else count_binary_kernel<TP, TT><<<grid, threads, 0, st>>>((const TP*)pred, (const TT*)truth, out, n);
    switch (truth_dtype) {
        case 0: LAUNCH(long long) break;
        case 1: LAUNCH(int) break;
        case 2: LAUNCH(unsigned char) break;
        case 3: LAUNCH(float) break;
        default: return (int)cudaErrorInvalidValue;
    }
#undef LAUNCH
    COINN_CHECK_LAUNCH();
    return 0;
}

}  // namespace coinn

// integer-like dtypes: 0 int64, 1 int32, 2 uint8, 3 float32
COINN_API int coinn_count(const void* pred, int pred_dtype, const void* truth, int truth_dtype, void* out, long long n,
                          int C, int confusion, void* stream) {
    using namespace coinn;
    if (n == 0) return 0;
    if (confusion && (C <= 0 || (size_t)C * C * sizeof(unsigned int) > 48 * 1024)) return (int)cudaErrorInvalidValue;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
    switch (pred_dtype) {
        case 0: return dispatch_truth<long long>(pred, truth, truth_dtype, o, n, C, confusion != 0, st);
        case 1: return dispatch_truth<int>(pred, truth, truth_dtype, o, n, C, confusion != 0, st);
        case 2: return dispatch_truth<unsigned char>(pred, truth, truth_dtype, o, n, C, confusion != 0, st);
        case 3: return dispatch_truth<float>(pred, truth, truth_dtype, o, n, C, confusion != 0, st);
        default: return (int)cudaErrorInvalidValue;
    }
}
