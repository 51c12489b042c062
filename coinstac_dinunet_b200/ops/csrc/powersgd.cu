// PowerSGD helper: in-place modified Gram-Schmidt of a tall [m, r] fp32 matrix
// (row-major, r <= 32) in ONE launch, one CTA per matrix.  The reference runs O(r^2) tiny ATen
// launches per matrix (powersgd/__init__.py:15-38).
#include "common.cuh"

namespace coinn {

constexpr int kMaxRank = 32;

__global__ void __launch_bounds__(1024) orthogonalize_kernel(float* __restrict__ a, int m, int r, float eps) {
    __shared__ float scratch[32];
    __shared__ float dots[kMaxRank];
    for (int i = 0; i < r; ++i) {
        // 1) normalise column i
        float ss = 0.f;
        for (int row = threadIdx.x; row < m; row += blockDim.x) { const float v = a[(size_t)row * r + i]; ss += v * v; }
        const float nrm = sqrtf(block_sum(ss, scratch));
        const float inv = 1.f / (nrm + eps);
        for (int row = threadIdx.x; row < m; row += blockDim.x) a[(size_t)row * r + i] *= inv;
        __syncthreads();
        if (i + 1 >= r) break;
        // 2) projections of the remaining columns on column i (one pass over the rows)
        float part[kMaxRank];
#pragma unroll
        for (int j = 0; j < kMaxRank; ++j) part[j] = 0.f;
        for (int row = threadIdx.x; row < m; row += blockDim.x) {
            const float q = a[(size_t)row * r + i];
#pragma unroll
            for (int j = 0; j < kMaxRank; ++j)
                if (j > i && j < r) part[j] = fmaf(q, a[(size_t)row * r + j], part[j]);
        }
#pragma unroll
        for (int j = 0; j < kMaxRank; ++j) {
            if (j > i && j < r) {
                const float d = block_sum(part[j], scratch);
                if (threadIdx.x == 0) dots[j] = d;
            }
        }
        __syncthreads();
        // 3) remove them
        for (int row = threadIdx.x; row < m; row += blockDim.x) {
            const float q = a[(size_t)row * r + i];
            for (int j = i + 1; j < r; ++j) a[(size_t)row * r + j] -= dots[j] * q;
        }
        __syncthreads();
    }
}

}  // namespace coinn

COINN_API int coinn_orthogonalize(float* a, int m, int r, float eps, void* stream) {
    using namespace coinn;
    if (m == 0 || r == 0) return 0;
    if (r > kMaxRank) return (int)cudaErrorInvalidValue;
    int threads = m >= 1024 ? 1024 : ((m + 31) / 32) * 32;
    orthogonalize_kernel<<<1, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, m, r, eps);
    COINN_CHECK_LAUNCH();
    return 0;
}
