// Low-rank gradient-compression kernels for the device data plane:
//
//   PowerSGD (coinstac_dinunet/distrib/powersgd/__init__.py:81-83,123-126,166-169) - every compressible matrix of the model
//   in ONE launch per stage, driven by a device descriptor table:
//     psgd_mq           M = G + E (kept in E),  P = M Q                  reads G, E once; writes M
//     psgd_mtp          Q = M^T P
//     psgd_reconstruct  G = P Q^T,  E = M - P Q^T                        the error feedback is the epilogue
//     orthogonalize_batched   Gram-Schmidt of every P (or Q) factor, one CTA per matrix
//     segcopy           gather / scatter of the rank-1 (bias, norm) gradients into / out of the exchange buffer
//   r <= 8: these are GEMV-like, bound by streaming M once at HBM speed - tensor cores have nothing to do here.
//
//   rankDAD (rankdad/spi.py:9-86,190-250; rankdad/__init__.py:63-98):
//     gram_seg          BtB of a (column-block segmented) skinny matrix [rows, n], n <= 96
//     lowrank_eig       top-k singular triplets of B C^T entirely in the n-dimensional coefficient space
//                       (x <- Gc Gb x with B-metric deflation): one CTA, no host sync, no cuSOLVER
//     skinny_gemm_seg   out[rows, k] = B[rows, n] X[n, k]
//     dad_reconstruct   W.grad = delta act^T (+ bias column) written straight into the gradient arena
//   and the symmetric-memory all-gather of the factor buffers lives in xgpu_allgather (below), built on the same
//   flag barrier as the fused reduce.
#include "common.cuh"

namespace coinn {

constexpr int PS_MAX_R = 8;

struct PsgdDesc {           // one compressible matrix: G viewed as [n, m] row-major at arena offset g_off
    long long g_off;        // element offset into the gradient arena AND the error buffer
    long long p_off;        // element offset of P [n, r] in the P buffer
    long long q_off;        // element offset of Q [m, r] in the Q buffer
    int n, m;
};

// ------------------------------------------------------------------------------------------------ M = G + E, P = M Q
// grid.x = row tiles over all matrices (tile table), 256 threads: 8 warps, each warp owns rows; lanes stride the columns.
struct PsgdTile { int mat; int row0; };
constexpr int PS_ROWS = 8;                     // rows per CTA (one per warp)
static_assert(PS_ROWS == 256 / 32, "psgd row tiles: one row per warp of a 256-thread CTA");

__global__ void __launch_bounds__(256) psgd_mq_kernel(const PsgdDesc* __restrict__ desc, const PsgdTile* __restrict__ tiles,
                                                      const float* __restrict__ G, float* __restrict__ E,
                                                      const float* __restrict__ Q, float* __restrict__ P, int r, int use_error) {
    const PsgdTile t = tiles[blockIdx.x];
    const PsgdDesc d = desc[t.mat];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = t.row0 + warp;
    if (row >= d.n) return;
    const float* g = G + d.g_off + (long long)row * d.m;
    float* e = E + d.g_off + (long long)row * d.m;
    const float* q = Q + d.q_off;
    float acc[PS_MAX_R];
#pragma unroll
    for (int k = 0; k < PS_MAX_R; ++k) acc[k] = 0.f;
    for (int j = lane; j < d.m; j += 32) {
        float v = g[j];
        if (use_error) v += e[j];
        e[j] = v;                                  // E now holds M (turned back into the error by psgd_reconstruct)
#pragma unroll
        for (int k = 0; k < PS_MAX_R; ++k)
            if (k < r) acc[k] = fmaf(v, __ldg(q + (long long)j * r + k), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < PS_MAX_R; ++k) {
        if (k < r) {
            const float s = warp_sum(acc[k]);
            if (lane == 0) P[d.p_off + (long long)row * r + k] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ Q = M^T P
// one thread per column j of one matrix, rows split in chunks of PS_CHUNK (grid.y) combined with atomics into zeroed Q
constexpr int PS_CHUNK = 256;
struct PsgdColTile { int mat; int col0; int row0; };

__global__ void __launch_bounds__(256) psgd_mtp_kernel(const PsgdDesc* __restrict__ desc, const PsgdColTile* __restrict__ tiles,
                                                       const float* __restrict__ M, const float* __restrict__ P,
                                                       float* __restrict__ Q, int r) {
    const PsgdColTile t = tiles[blockIdx.x];
    const PsgdDesc d = desc[t.mat];
    __shared__ float sp[PS_CHUNK][PS_MAX_R];
    const int rows = min(PS_CHUNK, d.n - t.row0);
    for (int i = threadIdx.x; i < rows * r; i += blockDim.x)
        sp[i / r][i % r] = P[d.p_off + (long long)(t.row0 + i / r) * r + (i % r)];
    __syncthreads();
    const int j = t.col0 + threadIdx.x;
    if (j >= d.m) return;
    const float* m = M + d.g_off + (long long)t.row0 * d.m + j;
    float acc[PS_MAX_R];
#pragma unroll
    for (int k = 0; k < PS_MAX_R; ++k) acc[k] = 0.f;
    for (int i = 0; i < rows; ++i) {
        const float v = m[(long long)i * d.m];
#pragma unroll
        for (int k = 0; k < PS_MAX_R; ++k)
            if (k < r) acc[k] = fmaf(v, sp[i][k], acc[k]);
    }
    const bool single = d.n <= PS_CHUNK;
#pragma unroll
    for (int k = 0; k < PS_MAX_R; ++k) {
        if (k < r) {
            float* dst = Q + d.q_off + (long long)j * r + k;
            if (single) *dst = acc[k]; else atomicAdd(dst, acc[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ G = P Q^T ; E = M - G
__global__ void __launch_bounds__(256) psgd_reconstruct_kernel(const PsgdDesc* __restrict__ desc, const PsgdTile* __restrict__ tiles,
                                                               float* __restrict__ G, float* __restrict__ E,
                                                               const float* __restrict__ P, const float* __restrict__ Q,
                                                               int r, int use_error) {
    const PsgdTile t = tiles[blockIdx.x];
    const PsgdDesc d = desc[t.mat];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = t.row0 + warp;
    if (row >= d.n) return;
    float p[PS_MAX_R];
#pragma unroll
    for (int k = 0; k < PS_MAX_R; ++k) p[k] = k < r ? P[d.p_off + (long long)row * r + k] : 0.f;
    float* g = G + d.g_off + (long long)row * d.m;
    float* e = E + d.g_off + (long long)row * d.m;
    const float* q = Q + d.q_off;
    for (int j = lane; j < d.m; j += 32) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < PS_MAX_R; ++k)
            if (k < r) a = fmaf(p[k], __ldg(q + (long long)j * r + k), a);
        g[j] = a;
        e[j] = use_error ? e[j] - a : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ batched Gram-Schmidt
// one CTA per matrix; which = 0: the P factor [n, r], 1: the Q factor [m, r]
__global__ void __launch_bounds__(1024) orthogonalize_batched_kernel(const PsgdDesc* __restrict__ desc, float* __restrict__ buf,
                                                                     int which, int r, float eps) {
    __shared__ float scratch[32];
    __shared__ float dots[PS_MAX_R];
    const PsgdDesc d = desc[blockIdx.x];
    float* a = buf + (which == 0 ? d.p_off : d.q_off);
    const int m = which == 0 ? d.n : d.m;
    for (int i = 0; i < r; ++i) {
        float ss = 0.f;
        for (int row = threadIdx.x; row < m; row += blockDim.x) { const float v = a[(size_t)row * r + i]; ss += v * v; }
        const float inv = 1.f / (sqrtf(block_sum(ss, scratch)) + eps);
        for (int row = threadIdx.x; row < m; row += blockDim.x) a[(size_t)row * r + i] *= inv;
        __syncthreads();
        if (i + 1 >= r) break;
        float part[PS_MAX_R];
#pragma unroll
        for (int j = 0; j < PS_MAX_R; ++j) part[j] = 0.f;
        for (int row = threadIdx.x; row < m; row += blockDim.x) {
            const float q = a[(size_t)row * r + i];
#pragma unroll
            for (int j = 0; j < PS_MAX_R; ++j)
                if (j > i && j < r) part[j] = fmaf(q, a[(size_t)row * r + j], part[j]);
        }
#pragma unroll
        for (int j = 0; j < PS_MAX_R; ++j) {
            if (j > i && j < r) {                                    // uniform across the CTA
                const float dsum = block_sum(part[j], scratch);
                if (threadIdx.x == 0) dots[j] = dsum;
            }
        }
        __syncthreads();
        for (int row = threadIdx.x; row < m; row += blockDim.x) {
            const float q = a[(size_t)row * r + i];
            for (int j = i + 1; j < r; ++j) a[(size_t)row * r + j] -= dots[j] * q;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ segmented copy
struct Seg { long long src; long long dst; long long len; };

__global__ void __launch_bounds__(256) segcopy_kernel(const Seg* __restrict__ segs, int nseg, const float* __restrict__ src,
                                                      float* __restrict__ dst, float scale) {
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const Seg sg = segs[s];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sg.len; i += (long long)gridDim.x * blockDim.x)
            dst[sg.dst + i] = src[sg.src + i] * scale;
    }
}

// ================================================================================================ rankDAD
// A "segmented" skinny matrix: logical [rows, n] whose columns come in blocks of kseg, block s starting at
// base + s * seg_stride, each block row-major [rows, kseg]   (all-gathered factors: one block per site).
struct SegMat { const float* base; long long seg_stride; int rows, n, kseg; };

__device__ __forceinline__ float segmat_at(const SegMat& A, int row, int col) {
    const int s = col / A.kseg, c = col - s * A.kseg;
    return A.base[(long long)s * A.seg_stride + (long long)row * A.kseg + c];
}

constexpr int LR_MAX_N = 96;
constexpr int GRAM_ROWS = 64;

// part[chunk][n][n] = A_chunk^T A_chunk for GRAM_ROWS rows per CTA; gram_reduce_kernel then adds the chunks in a FIXED order.
// No atomics: every site re-compresses the same gathered factors redundantly and the replicas must stay bit-identical.
__global__ void __launch_bounds__(256) gram_seg_kernel(SegMat A, float* __restrict__ part) {
    extern __shared__ float tile[];                      // [GRAM_ROWS][n]
    const int n = A.n;
    const int row0 = blockIdx.x * GRAM_ROWS;
    const int rows = min(GRAM_ROWS, A.rows - row0);
    for (int i = threadIdx.x; i < rows * n; i += blockDim.x) tile[i] = segmat_at(A, row0 + i / n, i % n);
    __syncthreads();
    float* out = part + (long long)blockIdx.x * n * n;
    for (int o = threadIdx.x; o < n * n; o += blockDim.x) {
        const int i = o / n, j = o % n;
        if (j < i) continue;                              // symmetric: compute the upper triangle, mirror
        float s = 0.f;
        for (int rr = 0; rr < rows; ++rr) s = fmaf(tile[rr * n + i], tile[rr * n + j], s);
        out[(long long)i * n + j] = s;
        out[(long long)j * n + i] = s;
    }
}

__global__ void __launch_bounds__(256) gram_reduce_kernel(const float* __restrict__ part, int chunks, int nn, float* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nn) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += part[(long long)c * nn + o];
    out[o] = s;
}

// Top-k singular triplets of G = B C^T from the n x n Gram matrices Gb = B^T B, Gc = C^T C.
//   u = B x (left singular vectors live in range(B)), G G^T (B x) = B (Gc Gb x)  ->  iterate x <- Gc Gb x in R^n,
//   normalise / deflate in the B metric (x^T Gb x).   sigma^2 = (Gb x)^T Gc (Gb x).
// Outputs (coefficients, applied by skinny_gemm_seg afterwards):
//   X [n, k]: left * sigma = B X        Y [n, k]: right = C Y        (columns below tol * sigma_0 are zeroed)
__global__ void __launch_bounds__(128) lowrank_eig_kernel(const float* __restrict__ Gb_g, const float* __restrict__ Gc_g, int n, int k,
                                                          int iters, float tol, float* __restrict__ X, float* __restrict__ Y) {
    extern __shared__ float sm[];
    float* Gb = sm;                     // [n][n]
    float* Gc = Gb + n * n;             // [n][n]
    float* x = Gc + n * n;              // [n]
    float* w = x + n;                   // [n]   Gb x
    float* z = w + n;                   // [n]   Gc w
    float* prev = z + n;                // [k][n] previous (B-orthonormal) coefficient vectors
    float* prevw = prev + k * n;        // [k][n] Gb * prev
    __shared__ float red[32];
    __shared__ float sigma0;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n * n; i += nt) { Gb[i] = Gb_g[i]; Gc[i] = Gc_g[i]; }
    __syncthreads();
    auto matvec = [&](const float* Mx, const float* v, float* out) {
        for (int i = tid; i < n; i += nt) {
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(Mx[i * n + j], v[j], s);
            out[i] = s;
        }
        __syncthreads();
    };
    auto dot = [&](const float* a, const float* b) {
        float s = 0.f;
        for (int i = tid; i < n; i += nt) s = fmaf(a[i], b[i], s);
        return block_sum(s, red);
    };
    for (int c = 0; c < k; ++c) {
        // deterministic start vector (same on every site): a fixed quasi-random pattern
        for (int i = tid; i < n; i += nt) x[i] = 0.5f + 0.5f * __sinf(12.9898f * (float)(i + 1) + 78.233f * (float)(c + 1));
        __syncthreads();
        float sig2 = 0.f;
        for (int it = 0; it <= iters; ++it) {
            // B-orthogonalise against the previous vectors, normalise in the B metric
            matvec(Gb, x, w);
            for (int p = 0; p < c; ++p) {
                const float a = dot(prevw + p * n, x);            // u_p . u = x_p^T Gb x
                for (int i = tid; i < n; i += nt) x[i] -= a * prev[p * n + i];
                __syncthreads();
            }
            if (c > 0) matvec(Gb, x, w);
            const float nrm2 = dot(x, w);
            const float inv = nrm2 > 1e-30f ? rsqrtf(nrm2) : 0.f;
            for (int i = tid; i < n; i += nt) { x[i] *= inv; w[i] *= inv; }
            __syncthreads();
            matvec(Gc, w, z);                                     // z = Gc Gb x
            sig2 = dot(w, z);                                     // sigma^2 for the current (normalised) x
            if (it == iters) break;
            for (int i = tid; i < n; i += nt) x[i] = z[i];        // x <- Gc Gb x
            __syncthreads();
        }
        const float sig = sqrtf(fmaxf(sig2, 0.f));
        if (c == 0 && tid == 0) sigma0 = sig;
        __syncthreads();
        const bool keep = sig > tol * fmaxf(sigma0, 1e-30f) && sig > 0.f;
        for (int i = tid; i < n; i += nt) {
            prev[c * n + i] = x[i];
            prevw[c * n + i] = w[i];
            X[(long long)i * k + c] = keep ? x[i] * sig : 0.f;
            Y[(long long)i * k + c] = keep ? w[i] / sig : 0.f;
        }
        __syncthreads();
    }
}

// out[rows, k] = A[rows, n] X[n, k]      (k <= 16; X in shared memory)
__global__ void __launch_bounds__(256) skinny_gemm_seg_kernel(SegMat A, const float* __restrict__ X, int k, float* __restrict__ out,
                                                              float scale) {
    extern __shared__ float sx[];                        // [n][k]
    for (int i = threadIdx.x; i < A.n * k; i += blockDim.x) sx[i] = X[i];
    __syncthreads();
    const long long total = (long long)A.rows * k;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(o / k), c = (int)(o % k);
        float s = 0.f;
        for (int j = 0; j < A.n; ++j) s = fmaf(segmat_at(A, row, j), sx[j * k + c], s);
        out[o] = s * scale;
    }
}

// W.grad[out, in] (+)= sum_c delta[out, c] act[in, c];  act may carry one extra row (the bias column): b.grad[out] gets it
__global__ void __launch_bounds__(256) dad_reconstruct_kernel(const float* __restrict__ delta, const float* __restrict__ act,
                                                              int out_f, int in_f, int act_rows, int k, float* __restrict__ wgrad,
                                                              float* __restrict__ bgrad, float scale) {
    extern __shared__ float sd[];                        // delta rows of this CTA: [16][k]
    const int o0 = blockIdx.y * 16;
    for (int i = threadIdx.x; i < 16 * k; i += blockDim.x) {
        const int oo = o0 + i / k;
        sd[i] = oo < out_f ? delta[(long long)oo * k + (i % k)] * scale : 0.f;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;          // input feature (or the bias column)
    if (j >= act_rows) return;
    float a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = c < k ? act[(long long)j * k + c] : 0.f;
    for (int oo = 0; oo < 16 && o0 + oo < out_f; ++oo) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < k) s = fmaf(sd[oo * k + c], a[c], s);
        if (j < in_f) wgrad[(long long)(o0 + oo) * in_f + j] = s;
        else if (bgrad) bgrad[o0 + oo] = s;
    }
}

// ================================================================================================ symmetric all-gather
constexpr int AG_MAX_RANKS = 8;
constexpr int AG_MAX_BLOCKS = 64;

struct AllGatherArgs {
    const float* src_ptrs[AG_MAX_RANKS];   // every rank's (peer-mapped) send buffer
    uint32_t*    flag_ptrs[AG_MAX_RANKS];  // every rank's flag pad [AG_MAX_BLOCKS][AG_MAX_RANKS]
    float*       dst;                      // local [world][numel]
    uint32_t*    epoch;                    // local [AG_MAX_BLOCKS]
    int*         error;
    long long    numel;                    // floats per rank (multiple of 4)
    int rank, world;
    unsigned timeout_ms;
    int _pad;
};

__device__ __forceinline__ void ag_barrier(const AllGatherArgs& a, int block, uint32_t seq) {
    __syncthreads();
    if (threadIdx.x < a.world) {
        const int peer = threadIdx.x;
        st_release_sys(a.flag_ptrs[peer] + block * AG_MAX_RANKS + a.rank, seq);
        const uint32_t* mine = a.flag_ptrs[a.rank] + block * AG_MAX_RANKS + peer;
        volatile int* err = a.error;
        if (!(err && *err)) {
            unsigned probes = 0;
            unsigned long long t0 = 0;
            while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
                if (a.timeout_ms && ((++probes & 1023u) == 0u)) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > (unsigned long long)a.timeout_ms * 1000000ull) { if (err) atomicCAS(a.error, 0, 1 + peer); break; }
                    if (err && *err) break;
                }
            }
        }
    }
    __syncthreads();
}

// every rank copies every rank's send buffer (its own included) into dst[rank_index]; two flag barriers fence the
// send buffers: peers' data is final before anybody reads it, and nobody overwrites it before everybody has read it
__global__ void __launch_bounds__(512, 1) allgather_kernel(const AllGatherArgs a) {
    const int b = blockIdx.x, G = gridDim.x;
    const uint32_t seq = a.epoch[b] + 1u;
    __threadfence_system();
    ag_barrier(a, b, 2u * seq - 1u);
    const long long nvec = a.numel >> 2;
    const long long chunk = (nvec + G - 1) / G;
    const long long lo = min((long long)b * chunk, nvec), hi = min(lo + chunk, nvec);
    for (int p = 0; p < a.world; ++p) {
        const float4* src = reinterpret_cast<const float4*>(a.src_ptrs[p]);
        float4* dst = reinterpret_cast<float4*>(a.dst + (long long)p * a.numel);
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = ld_stream_f4(src + i);
    }
    __threadfence_system();
    ag_barrier(a, b, 2u * seq);
    if (threadIdx.x == 0) a.epoch[b] = seq;
}

}  // namespace coinn

using namespace coinn;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

COINN_API int coinn_psgd_desc_size() { return (int)sizeof(PsgdDesc); }
COINN_API int coinn_psgd_rows_per_tile() { return PS_ROWS; }
COINN_API int coinn_psgd_row_chunk() { return PS_CHUNK; }

COINN_API int coinn_psgd_mq(const void* desc, const void* tiles, int ntiles, const float* G, float* E, const float* Q, float* P, int r,
                            int use_error, void* stream) {
    if (r < 1 || r > PS_MAX_R) return (int)cudaErrorInvalidValue;
    if (ntiles == 0) return 0;
    psgd_mq_kernel<<<ntiles, 256, 0, ST(stream)>>>((const PsgdDesc*)desc, (const PsgdTile*)tiles, G, E, Q, P, r, use_error);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_psgd_mtp(const void* desc, const void* tiles, int ntiles, const float* M, const float* P, float* Q, int r, void* stream) {
    if (r < 1 || r > PS_MAX_R) return (int)cudaErrorInvalidValue;
    if (ntiles == 0) return 0;
    psgd_mtp_kernel<<<ntiles, 256, 0, ST(stream)>>>((const PsgdDesc*)desc, (const PsgdColTile*)tiles, M, P, Q, r);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_psgd_reconstruct(const void* desc, const void* tiles, int ntiles, float* G, float* E, const float* P, const float* Q,
                                     int r, int use_error, void* stream) {
    if (r < 1 || r > PS_MAX_R) return (int)cudaErrorInvalidValue;
    if (ntiles == 0) return 0;
    psgd_reconstruct_kernel<<<ntiles, 256, 0, ST(stream)>>>((const PsgdDesc*)desc, (const PsgdTile*)tiles, G, E, P, Q, r, use_error);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_orthogonalize_batched(const void* desc, int nmat, float* buf, int which, int r, float eps, void* stream) {
    if (r < 1 || r > PS_MAX_R) return (int)cudaErrorInvalidValue;
    if (nmat == 0) return 0;
    orthogonalize_batched_kernel<<<nmat, 1024, 0, ST(stream)>>>((const PsgdDesc*)desc, buf, which, r, eps);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_segcopy(const void* segs, int nseg, long long max_len, const float* src, float* dst, float scale, void* stream) {
    if (nseg == 0) return 0;
    long long bx = (max_len + 255) / 256;
    dim3 grid((unsigned)(bx < 1 ? 1 : (bx > 64 ? 64 : bx)), (unsigned)(nseg > 1024 ? 1024 : nseg));
    segcopy_kernel<<<grid, 256, 0, ST(stream)>>>((const Seg*)segs, nseg, src, dst, scale);
    COINN_CHECK_LAUNCH();
    return 0;
}

static SegMat mk_seg(const float* base, long long seg_stride, int rows, int n, int kseg) {
    SegMat A; A.base = base; A.seg_stride = seg_stride; A.rows = rows; A.n = n; A.kseg = kseg > 0 ? kseg : n; return A;
}

// out [n, n] = A^T A (deterministic two-stage sum); scratch: >= ceil(rows / 64) * n * n floats
COINN_API int coinn_gram_seg(const float* base, long long seg_stride, int rows, int n, int kseg, float* out, float* scratch, void* stream) {
    if (n < 1 || n > LR_MAX_N || rows < 1) return (int)cudaErrorInvalidValue;
    const int chunks = (rows + GRAM_ROWS - 1) / GRAM_ROWS;
    const int smem = GRAM_ROWS * n * (int)sizeof(float);
    gram_seg_kernel<<<chunks, 256, smem, ST(stream)>>>(mk_seg(base, seg_stride, rows, n, kseg), scratch);
    COINN_CHECK_LAUNCH();
    gram_reduce_kernel<<<(n * n + 255) / 256, 256, 0, ST(stream)>>>(scratch, chunks, n * n, out);
    COINN_CHECK_LAUNCH();
    return 0;
}
COINN_API int coinn_gram_rows_per_chunk() { return GRAM_ROWS; }

COINN_API int coinn_lowrank_eig(const float* Gb, const float* Gc, int n, int k, int iters, float tol, float* X, float* Y, void* stream) {
    if (n < 1 || n > LR_MAX_N || k < 1 || k > 16 || k > n) return (int)cudaErrorInvalidValue;
    const int smem = (2 * n * n + 3 * n + 2 * k * n) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(lowrank_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); attr = true; }
    lowrank_eig_kernel<<<1, 128, smem, ST(stream)>>>(Gb, Gc, n, k, iters, tol, X, Y);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_skinny_gemm_seg(const float* base, long long seg_stride, int rows, int n, int kseg, const float* X, int k, float* out,
                                    float scale, void* stream) {
    if (n < 1 || n > LR_MAX_N || k < 1 || k > 16) return (int)cudaErrorInvalidValue;
    if (rows == 0) return 0;
    long long blocks = ((long long)rows * k + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    skinny_gemm_seg_kernel<<<(unsigned)blocks, 256, n * k * (int)sizeof(float), ST(stream)>>>(mk_seg(base, seg_stride, rows, n, kseg), X, k, out, scale);
    COINN_CHECK_LAUNCH();
    return 0;
}

// delta [out_f, k], act [act_rows, k] with act_rows == in_f (+1: bias column) -> wgrad [out_f, in_f], bgrad [out_f] or null
COINN_API int coinn_dad_reconstruct(const float* delta, const float* act, int out_f, int in_f, int act_rows, int k, float* wgrad,
                                    float* bgrad, float scale, void* stream) {
    if (k < 1 || k > 16 || act_rows < in_f || act_rows > in_f + 1) return (int)cudaErrorInvalidValue;
    dim3 grid((act_rows + 255) / 256, (out_f + 15) / 16);
    dad_reconstruct_kernel<<<grid, 256, 16 * k * (int)sizeof(float), ST(stream)>>>(delta, act, out_f, in_f, act_rows, k, wgrad, bgrad, scale);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_allgather_args_size() { return (int)sizeof(AllGatherArgs); }
COINN_API int coinn_allgather_flag_slots() { return AG_MAX_BLOCKS * AG_MAX_RANKS; }
COINN_API int coinn_allgather_max_blocks() { return AG_MAX_BLOCKS; }

COINN_API int coinn_allgather(const AllGatherArgs* args, void* stream) {
    AllGatherArgs a = *args;
    if (a.world < 1 || a.world > AG_MAX_RANKS || (a.numel & 3)) return (int)cudaErrorInvalidValue;
    if (a.numel == 0) return 0;
    long long want = ((a.numel >> 2) + 2047) / 2048;
    int grid = (int)(want < 1 ? 1 : (want > AG_MAX_BLOCKS ? AG_MAX_BLOCKS : want));
    allgather_kernel<<<grid, 512, 0, ST(stream)>>>(a);
    COINN_CHECK_LAUNCH();
    return 0;
}
