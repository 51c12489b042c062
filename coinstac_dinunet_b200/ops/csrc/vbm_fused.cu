// Fused per-site kernels of the VBM 3-D CNN blocks: channels-last (NDHWC) bf16
// activations, everything elementwise/normalisation/pooling folded into as few passes over HBM as
// training-mode BatchNorm allows.
//
//   (the first block, C_in = 1, lives in conv1_fused.cu: conv recomputed on tcgen05, nothing stored at full resolution)
//   bn_stats_kernel           sum / sumsq of a [M, C] bf16 tensor (statistics for the tcgen05 conv outputs)
//   bn_relu_pool_fwd_kernel   y -> maxpool2(relu(bn(y)))  : reads y once, writes 1/8 of it
//   bn_relu_pool_bwd_stats    dgamma, dbeta of the fused block from (y, dpooled)  [pass A]
//   bn_relu_pool_bwd_apply    dy of the fused block (full resolution)              [pass B]
//
// The PyTorch chain for one block (conv -> BN -> ReLU -> MaxPool, autocast bf16) moves ~10x the bytes of this.
#include "common.cuh"

namespace coinn {

struct Dims { int N, D, H, W; };

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
// stats[0:C] = sum_m y[m,c], stats[C:2C] = sum_m y[m,c]^2      (y: [M, C] bf16, C % 8 == 0, C <= 256)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ stats,
                                                       long long M, int C) {
    extern __shared__ float sm[];                       // [2*C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int chunks = C >> 3;                          // 16-byte chunks per row
    // thread owns one fixed channel chunk (so partial sums stay in registers) and strides over rows
    const int my_chunk = threadIdx.x % chunks;
    const int rows_per_iter = blockDim.x / chunks;
    const int my_row_off = threadIdx.x / chunks;
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    if (my_row_off < rows_per_iter) {
        for (long long m = (long long)blockIdx.x * rows_per_iter + my_row_off; m < M; m += (long long)gridDim.x * rows_per_iter) {
            const uint4 u = ld_stream_u4(reinterpret_cast<const uint4*>(y + m * C) + my_chunk);
            float f[8];
            unpack8(u, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sm[my_chunk * 8 + j], s1[j]);
            atomicAdd(&sm[C + my_chunk * 8 + j], s2[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], sm[i]);
}

// mean / invstd from sums; optionally update running stats (PyTorch semantics: unbiased running var)
__global__ void bn_finalize_kernel(float* __restrict__ stats, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float count, float eps, float momentum, int C, long long* __restrict__ num_batches_tracked,
                                   int zero_stats) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    const float mu = stats[c] / count;
    const float var = fmaxf(stats[C + c] / count - mu * mu, 0.f);
    if (zero_stats) { stats[c] = 0.f; stats[C + c] = 0.f; }   // persistent accumulator: ready for the next step, no memset launch
    mean[c] = mu;
    invstd[c] = rsqrtf(var + eps);
    if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
        const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// ------------------------------------------------------------------------------------------------
// p[n, d/2, h/2, w/2, c] = max over the 2x2x2 window of relu(scale[c] * y + shift[c])
// one thread = one pooled voxel x 8 channels (a 16-byte chunk)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, __nv_bfloat16* __restrict__ p,
                                                               Dims d, int C) {
    const int chunks = C >> 3;
    const int PD = d.D >> 1, PH = d.H >> 1, PW = d.W >> 1;
    const long long total = (long long)d.N * PD * PH * PW * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        long long t = i / chunks;
        const int pw = (int)(t % PW); t /= PW;
        const int ph = (int)(t % PH); t /= PH;
        const int pd = (int)(t % PD);
        const int n = (int)(t / PD);
        float sc[8], sh[8], best[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            sc[j] = gamma[c] * invstd[c];
            sh[j] = beta[c] - mean[c] * sc[j];
            best[j] = 0.f;                               // relu floor
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int zd = 2 * pd + (k >> 2), zh = 2 * ph + ((k >> 1) & 1), zw = 2 * pw + (k & 1);
            const long long vox = (((long long)n * d.D + zd) * d.H + zh) * d.W + zw;
            const uint4 u = ld_stream_u4(reinterpret_cast<const uint4*>(y + vox * C) + ch);
            float f[8];
            unpack8(u, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], fmaf(f[j], sc[j], sh[j]));
        }
        reinterpret_cast<uint4*>(p + (i / chunks) * C)[ch] = pack8(best);
    }
}

// ------------------------------------------------------------------------------------------------
// backward of the fused block.  For every pooled voxel: recompute z = relu(bn(y)) on its window, route dp to the
// first arg-max (PyTorch tie rule), zero if the max is not positive.
//   pass A: dbeta[c] = sum dz ; dgamma[c] = sum dz * xhat            (acc[0:C] = dbeta, acc[C:2C] = dgamma)
//   pass B: dy = gamma*invstd * (dz - dbeta/M - xhat * dgamma/M)     for ALL voxels (incl. odd borders)
// ------------------------------------------------------------------------------------------------
template <bool APPLY>
__global__ void __launch_bounds__(256, APPLY ? 1 : 2) bn_relu_pool_bwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dp,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ acc, __nv_bfloat16* __restrict__ dy,
                                                               Dims d, int C, float inv_count) {
    extern __shared__ float sm[];                        // pass A: [2*C] block partials
    const int chunks = C >> 3;
    if (!APPLY) {
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
    }
    // windows cover ceil(dim/2) cells so that odd border voxels (not pooled) still get their dy in pass B
    const int PD = d.D >> 1, PH = d.H >> 1, PW = d.W >> 1;
    const int CD = (d.D + 1) >> 1, CH = (d.H + 1) >> 1, CW = (d.W + 1) >> 1;
    const int my_chunk = threadIdx.x % chunks;
    const int cells_per_iter = blockDim.x / chunks;
    const int my_cell_off = threadIdx.x / chunks;
    const long long total_cells = (long long)d.N * CD * CH * CW;

    // per-channel constants of this thread's 8 channels
    float sc[8], sh[8], mu[8], is[8], a0[8], a1[8];     // pass A: a0/a1 accumulate dbeta/dgamma; pass B: a0/a1 = dbeta/M, dgamma/M
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = my_chunk * 8 + j;
        mu[j] = mean[c]; is[j] = invstd[c];
        sc[j] = gamma[c] * is[j];
        sh[j] = beta[c] - mu[j] * sc[j];
        a0[j] = APPLY ? acc[c] * inv_count : 0.f;
        a1[j] = APPLY ? acc[C + c] * inv_count : 0.f;
    }
    if (my_cell_off < cells_per_iter) {
        for (long long cell = (long long)blockIdx.x * cells_per_iter + my_cell_off; cell < total_cells;
             cell += (long long)gridDim.x * cells_per_iter) {
            long long t = cell;
            const int pw = (int)(t % CW); t /= CW;
            const int ph = (int)(t % CH); t /= CH;
            const int pd = (int)(t % CD);
            const int n = (int)(t / CD);
            const bool pooled = pd < PD && ph < PH && pw < PW;
            if (!pooled && !APPLY) continue;             // no gradient flows through un-pooled borders
            // issue all loads of the window first (8 x 16 B + dp), values stay packed as bf16
            uint4 raw[8];
            uint4 graw = make_uint4(0u, 0u, 0u, 0u);
            if (pooled) {
                const long long pv = (((long long)n * PD + pd) * PH + ph) * PW + pw;
                graw = ld_stream_u4(reinterpret_cast<const uint4*>(dp + pv * C) + my_chunk);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int zd = 2 * pd + (k >> 2), zh = 2 * ph + ((k >> 1) & 1), zw = 2 * pw + (k & 1);
                raw[k] = make_uint4(0u, 0u, 0u, 0u);
                if (zd < d.D && zh < d.H && zw < d.W) {
                    const long long vox = (((long long)n * d.D + zd) * d.H + zh) * d.W + zw;
                    raw[k] = ld_stream_u4(reinterpret_cast<const uint4*>(y + vox * C) + my_chunk);
                }
            }
            float g[8], best[8], besty[8];
            int arg[8];
            unpack8(graw, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; besty[j] = 0.f; arg[j] = 0; }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float f[8];
                unpack8(raw[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float z = fmaf(f[j], sc[j], sh[j]);
                    if (z > best[j]) { best[j] = z; besty[j] = f[j]; arg[j] = k; }      // first max wins (PyTorch rule)
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) if (!(best[j] > 0.f) || !pooled) g[j] = 0.f;       // relu'(<=0) = 0
            if (!APPLY) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a0[j] += g[j];
                    a1[j] = fmaf(g[j], (besty[j] - mu[j]) * is[j], a1[j]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int zd = 2 * pd + (k >> 2), zh = 2 * ph + ((k >> 1) & 1), zw = 2 * pw + (k & 1);
                    if (!(zd < d.D && zh < d.H && zw < d.W)) continue;
                    float f[8], o[8];
                    unpack8(raw[k], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dz = (arg[j] == k) ? g[j] : 0.f;
                        const float xh = (f[j] - mu[j]) * is[j];
                        o[j] = sc[j] * (dz - a0[j] - xh * a1[j]);
                    }
                    const long long vox = (((long long)n * d.D + zd) * d.H + zh) * d.W + zw;
                    reinterpret_cast<uint4*>(dy + vox * C)[my_chunk] = pack8(o);
                }
            }
        }
    }
    if (!APPLY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sm[my_chunk * 8 + j], a0[j]);
            atomicAdd(&sm[C + my_chunk * 8 + j], a1[j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&acc[i], sm[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// Pass B, 4 channels (8 bytes) per thread: half the per-thread state of the 8-channel version (<= 96 registers,
// 2-3 CTAs per SM instead of 1), which is what a latency-bound streaming kernel needs to approach the copy roofline.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack4(const uint2& u, float (&f)[4]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
}
__global__ void __launch_bounds__(256, 2) bn_relu_pool_bwd_apply4_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dp,
                                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        const float* __restrict__ acc, __nv_bfloat16* __restrict__ dy,
                                                                        Dims d, int C, float inv_count) {
    const int chunks = C >> 2;                            // 8-byte chunks per voxel
    const int PD = d.D >> 1, PH = d.H >> 1, PW = d.W >> 1;
    const int CD = (d.D + 1) >> 1, CH = (d.H + 1) >> 1, CW = (d.W + 1) >> 1;
    const int my_chunk = threadIdx.x % chunks;
    const int cells_per_iter = blockDim.x / chunks;
    const int my_cell_off = threadIdx.x / chunks;
    const long long total_cells = (long long)d.N * CD * CH * CW;
    if (my_cell_off >= cells_per_iter) return;

    float sc[4], sh[4], mu[4], is[4], c1[4], c2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = my_chunk * 4 + j;
        mu[j] = mean[c]; is[j] = invstd[c];
        sc[j] = gamma[c] * is[j];
        sh[j] = beta[c] - mu[j] * sc[j];
        c1[j] = acc[c] * inv_count;
        c2[j] = acc[C + c] * inv_count;
    }
    for (long long cell = (long long)blockIdx.x * cells_per_iter + my_cell_off; cell < total_cells;
         cell += (long long)gridDim.x * cells_per_iter) {
        long long t = cell;
        const int pw = (int)(t % CW); t /= CW;
        const int ph = (int)(t % CH); t /= CH;
        const int pd = (int)(t % CD);
        const int n = (int)(t / CD);
        const bool pooled = pd < PD && ph < PH && pw < PW;
        uint2 raw[8];
        uint2 graw = make_uint2(0u, 0u);
        if (pooled) {
            const long long pv = (((long long)n * PD + pd) * PH + ph) * PW + pw;
            graw = ld_stream_u2(reinterpret_cast<const uint2*>(dp + pv * C) + my_chunk);
        }
        long long vox[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int zd = 2 * pd + (k >> 2), zh = 2 * ph + ((k >> 1) & 1), zw = 2 * pw + (k & 1);
            const bool in = zd < d.D && zh < d.H && zw < d.W;
            vox[k] = in ? (((long long)n * d.D + zd) * d.H + zh) * d.W + zw : -1;
            raw[k] = in ? ld_stream_u2(reinterpret_cast<const uint2*>(y + vox[k] * C) + my_chunk) : make_uint2(0u, 0u);
        }
        float g[4], best[4];
        int arg[4];
        unpack4(graw, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) { best[j] = -INFINITY; arg[j] = 0; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float f[4];
            unpack4(raw[k], f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = fmaf(f[j], sc[j], sh[j]);
                if (z > best[j]) { best[j] = z; arg[j] = k; }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) if (!(best[j] > 0.f) || !pooled) g[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (vox[k] < 0) continue;
            float f[4], o[4];
            unpack4(raw[k], f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dz = (arg[j] == k) ? g[j] : 0.f;
                const float xh = (f[j] - mu[j]) * is[j];
                o[j] = sc[j] * (dz - c1[j] - xh * c2[j]);
            }
            st_stream_u2(reinterpret_cast<uint2*>(dy + vox[k] * C) + my_chunk, make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3])));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pass A without touching y:  the pooled output already carries what the reductions need.
//   z_argmax = relu-input at the arg-max = p (when p > 0)   =>   xhat_argmax = (p - beta) / gamma
//   dbeta[c]  = sum_cells dp * [p > 0]          dgamma[c] = sum_cells dp * [p > 0] * (p - beta) / gamma
// p, dp: [cells, C] bf16 (1/8 of y).  acc[0:C] += dbeta, acc[C:2C] += dgamma.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_pool_bwd_stats_pooled_kernel(const __nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ dp,
                                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                       float* __restrict__ acc, long long cells, int C) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int chunks = C >> 3;
    const int my_chunk = threadIdx.x % chunks;
    const int rows_per_iter = blockDim.x / chunks;
    const int my_row = threadIdx.x / chunks;
    float be[8], ig[8], a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float g = gamma[my_chunk * 8 + j];
        be[j] = beta[my_chunk * 8 + j];
        ig[j] = fabsf(g) > 1e-12f ? 1.f / g : 0.f;
        a0[j] = 0.f; a1[j] = 0.f;
    }
    if (my_row < rows_per_iter) {
        // 4 rows (8 independent 16-byte loads) in flight per thread: the single-row loop was bound by load latency
        // (58 us for the 66 MB of layer 1, 10 us at the copy roofline)
        const long long stride = (long long)gridDim.x * rows_per_iter;
        for (long long m0 = (long long)blockIdx.x * rows_per_iter + my_row; m0 < cells; m0 += 4 * stride) {
            uint4 pr[4], gr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long m = m0 + u * stride;
                pr[u] = make_uint4(0u, 0u, 0u, 0u); gr[u] = pr[u];
                if (m < cells) {
                    pr[u] = ld_stream_u4(reinterpret_cast<const uint4*>(p + m * C) + my_chunk);
                    gr[u] = ld_stream_u4(reinterpret_cast<const uint4*>(dp + m * C) + my_chunk);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float pv[8], gv[8];
                unpack8(pr[u], pv);
                unpack8(gr[u], gv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float g = pv[j] > 0.f ? gv[j] : 0.f;          // rows beyond the end were loaded as p = 0
                    a0[j] += g;
                    a1[j] = fmaf(g, (pv[j] - be[j]) * ig[j], a1[j]);
                }
            }
        }
    }
    // lanes l, l + chunks, l + 2*chunks, ... of a warp hold the same channel chunk: butterfly over those first, so only
    // `chunks` lanes per warp touch shared memory (128 threads hammering 16 addresses serialised the old epilogue), and
    // a small grid keeps the same-address global atomics short (1184 CTAs x 32 atomics cost more than the 66 MB read)
    if (chunks <= 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            for (int off = 16; off >= chunks; off >>= 1) {
                a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], off);
                a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], off);
            }
        }
    }
    if (my_row < rows_per_iter && (chunks > 32 || (threadIdx.x & 31) < chunks)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sm[my_chunk * 8 + j], a0[j]);
            atomicAdd(&sm[C + my_chunk * 8 + j], a1[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&acc[i], sm[i]);
}

// ------------------------------------------------------------------------------------------------
// One launch packs a conv weight [COUT, CIN, 27] (fp32) into both bf16 GEMM operands:
//   wf[co, tap*CIN + ci]          (fprop:  y  = conv(x,  W))
//   wd[ci, (26-tap)*COUT + co]    (dgrad:  dx = conv(dy, flip(W)^T))
// Row pitches kf / kd are the K extents rounded up to 64; the padding is zeroed once when the buffers are created.
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                         int cout, int cin, int kf, int kd) {
    const int total = cout * cin * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % 27, ci = (i / 27) % cin, co = i / (27 * cin);
        const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
        wf[(size_t)co * kf + tap * cin + ci] = v;
        if (wd) wd[(size_t)ci * kd + (26 - tap) * cout + co] = v;
    }
}

static inline int grid_for(long long work_items, int threads, int per_thread = 4) {
    long long want = (work_items + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
    const long long cap = 8LL * B200_SM_COUNT;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace coinn

using coinn::Dims;


COINN_API int coinn_bn_stats(const void* y, float* stats, long long M, int C, void* stream) {
    using namespace coinn;
    if (C % 8 || C > 256 || 256 % (C / 8)) return (int)cudaErrorInvalidValue;
    const int rows_per_iter = 256 / (C / 8);
    long long want = (M + (long long)rows_per_iter * 16 - 1) / ((long long)rows_per_iter * 16);
    const int grid = (int)(want < 1 ? 1 : (want > 4LL * B200_SM_COUNT ? 4LL * B200_SM_COUNT : want));
    bn_stats_kernel<<<grid, 256, 2 * C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __nv_bfloat16*)y, stats, M, C);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_bn_finalize(const float* stats, float* mean, float* invstd, float* running_mean, float* running_var,
                                float count, float eps, float momentum, int C, void* stream) {
    coinn::bn_finalize_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        const_cast<float*>(stats), mean, invstd, running_mean, running_var, count, eps, momentum, C, nullptr, 0);
    COINN_CHECK_LAUNCH();
    return 0;
}

// + num_batches_tracked += 1 (optional) and re-zeroing of the (persistent) stats accumulator in the same launch
COINN_API int coinn_bn_finalize2(float* stats, float* mean, float* invstd, float* running_mean, float* running_var,
                                 float count, float eps, float momentum, int C, long long* num_batches_tracked, int zero_stats,
                                 void* stream) {
    coinn::bn_finalize_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        stats, mean, invstd, running_mean, running_var, count, eps, momentum, C, num_batches_tracked, zero_stats);
    COINN_CHECK_LAUNCH();
    return 0;
}

// End of a conv block's backward, ONE launch instead of a permute copy, three AccumulateGrad adds and two memsets:
//   w_grad[co][ci][tap] += dwt[(tap*cin + ci)*cout + co]   (transposed != 0; the tcgen05 wgrad kernels' layout)
//   w_grad[i]           += dwt[i]                          (transposed == 0; conv1: already [co][tap])
//   gamma_grad += acc[C:2C], beta_grad += acc[0:C]; dwt and acc are zeroed for the next step (persistent accumulators)
namespace coinn {
__global__ void conv_block_grad_finalize_kernel(float* __restrict__ dwt, float* __restrict__ acc, float* __restrict__ w_grad,
                                                float* __restrict__ gamma_grad, float* __restrict__ beta_grad, int cin, int cout,
                                                int transposed) {
    extern __shared__ float tile[];                          // transposed: [27][cout] slice of one input channel
    if (!transposed) {
        const int total = 27 * cin * cout;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            w_grad[i] += dwt[i];
            dwt[i] = 0.f;
        }
    } else {
        // one CTA per input channel ci: the 27 rows dwt[(tap, ci), :] are read (and zeroed) coalesced along co, the
        // gradient is written as 27-float runs w_grad[co][ci][0..26] (a direct permuting loop is a 4-byte scatter with
        // stride 27*cin: 60 us for the 3.5 MB of block 5)
        for (int ci = blockIdx.x; ci < cin; ci += gridDim.x) {
            for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
                const int tap = i / cout, co = i - tap * cout;
                const size_t src = ((size_t)tap * cin + ci) * cout + co;
                tile[tap * (cout + 1) + co] = dwt[src];          // +1: conflict-free column reads below
                dwt[src] = 0.f;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
                const int co = i / 27, tap = i - co * 27;
                w_grad[((size_t)co * cin + ci) * 27 + tap] += tile[tap * (cout + 1) + co];
            }
            __syncthreads();
        }
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < cout; c += blockDim.x) {
            beta_grad[c] += acc[c];
            gamma_grad[c] += acc[cout + c];
            acc[c] = 0.f; acc[cout + c] = 0.f;
        }
    }
}
}  // namespace coinn

COINN_API int coinn_conv_block_grad_finalize(float* dwt, float* acc, float* w_grad, float* gamma_grad, float* beta_grad, int cin, int cout,
                                             int transposed, void* stream) {
    const int total = 27 * cin * cout;
    int grid = transposed ? cin : (total + 255) / 256;
    if (grid > 2 * B200_SM_COUNT) grid = 2 * B200_SM_COUNT;
    const size_t smem = transposed ? (size_t)27 * (cout + 1) * sizeof(float) : 0;
    coinn::conv_block_grad_finalize_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(dwt, acc, w_grad, gamma_grad, beta_grad,
                                                                                                        cin, cout, transposed);
    COINN_CHECK_LAUNCH();
    return 0;
}

COINN_API int coinn_bn_relu_pool_fwd(const void* y, const float* mean, const float* invstd, const float* gamma,
                                     const float* beta, void* p, int N, int D, int H, int W, int C, void* stream) {
    using namespace coinn;
    if (C % 8) return (int)cudaErrorInvalidValue;
    Dims d{N, D, H, W};
    const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return 0;
    bn_relu_pool_fwd_kernel<<<grid_for(total, 256, 2), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __nv_bfloat16*)y, mean, invstd, gamma, beta, (__nv_bfloat16*)p, d, C);
    COINN_CHECK_LAUNCH();
    return 0;
}

// apply == 0: acc[2C] (zeroed) += (dbeta, dgamma).  apply == 1: writes dy using acc.
COINN_API int coinn_bn_relu_pool_bwd(const void* y, const void* dp, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, float* acc, void* dy,
                                     int N, int D, int H, int W, int C, int apply, void* stream) {
    using namespace coinn;
    if (C % 8 || C > 256 || 256 % (C / 8)) return (int)cudaErrorInvalidValue;
    Dims d{N, D, H, W};
    const int cells_per_iter = 256 / (C / 8);
    const long long cells = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
    long long want = (cells + (long long)cells_per_iter * 2 - 1) / ((long long)cells_per_iter * 2);
    const int grid = (int)(want < 1 ? 1 : (want > 8LL * B200_SM_COUNT ? 8LL * B200_SM_COUNT : want));
    const float inv_count = 1.f / ((float)N * D * H * W);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (apply) {
        if (256 % (C / 4) == 0) {                        // 4-channel threads: twice the CTAs per SM
            const int cpi = 256 / (C / 4);
            long long w4 = (cells + (long long)cpi * 2 - 1) / ((long long)cpi * 2);
            const int g4 = (int)(w4 < 1 ? 1 : (w4 > 16LL * B200_SM_COUNT ? 16LL * B200_SM_COUNT : w4));
            bn_relu_pool_bwd_apply4_kernel<<<g4, 256, 0, st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)dp, mean, invstd, gamma, beta,
                                                                acc, (__nv_bfloat16*)dy, d, C, inv_count);
        } else {
            bn_relu_pool_bwd_kernel<true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)dp, mean, invstd,
                                                                gamma, beta, acc, (__nv_bfloat16*)dy, d, C, inv_count);
        }
    }
    else bn_relu_pool_bwd_kernel<false><<<grid, 256, 2 * C * sizeof(float), st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)dp, mean,
                                                                                   invstd, gamma, beta, acc, nullptr, d, C, inv_count);
    COINN_CHECK_LAUNCH();
    return 0;
}

// acc[2C] (zeroed) += (dbeta, dgamma) computed from the pooled output p and its gradient dp only
COINN_API int coinn_bn_pool_bwd_stats_pooled(const void* p, const void* dp, const float* gamma, const float* beta, float* acc,
                                             long long cells, int C, void* stream) {
    using namespace coinn;
    if (C % 8 || C > 256 || 256 % (C / 8)) return (int)cudaErrorInvalidValue;
    if (cells == 0) return 0;
    const int rows_per_iter = 256 / (C / 8);
    long long want = (cells + (long long)rows_per_iter * 8 - 1) / ((long long)rows_per_iter * 8);
    const int grid = (int)(want < 1 ? 1 : (want > 2LL * B200_SM_COUNT ? 2LL * B200_SM_COUNT : want));
    bn_pool_bwd_stats_pooled_kernel<<<grid, 256, 2 * C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __nv_bfloat16*)p, (const __nv_bfloat16*)dp, gamma, beta, acc, cells, C);
    COINN_CHECK_LAUNCH();
    return 0;
}


COINN_API int coinn_pack_conv_weights(const float* w, void* wf, void* wd, int cout, int cin, int kf, int kd, void* stream) {
    const int total = cout * cin * 27;
    const int grid = (total + 255) / 256 > 592 ? 592 : (total + 255) / 256;
    coinn::pack_conv_weights_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, cout, cin, kf, kd);
    COINN_CHECK_LAUNCH();
    return 0;
}
