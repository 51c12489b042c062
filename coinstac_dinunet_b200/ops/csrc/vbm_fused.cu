// Fused per-site kernels of the VBM 3-D CNN blocks: channels-last (NDHWC) bf16
// activations, everything elementwise/normalisation/pooling folded into as few passes over HBM as
// training-mode BatchNorm allows.
//
//   conv1_fwd_kernel          Conv3d(1 -> 16, k3 p1) on CUDA cores (K = 27 is tensor-core hostile) + per-channel
//                             sum / sum-of-squares in the same pass (BatchNorm batch statistics)
//   bn_stats_kernel           sum / sumsq of a [M, C] bf16 tensor (statistics for the tcgen05 conv outputs)
//   bn_relu_pool_fwd_kernel   y -> maxpool2(relu(bn(y)))  : reads y once, writes 1/8 of it
//   bn_relu_pool_bwd_stats    dgamma, dbeta of the fused block from (y, dpooled)  [pass A]
//   bn_relu_pool_bwd_apply    dy of the fused block (full resolution)              [pass B]
//   conv1_wgrad_kernel        dW1[16,27] = sum_vox dy[vox,:] (x) x[vox + tap]
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
// conv1: x [N,D,H,W] (fp32 or bf16, single channel) -> y [N,D,H,W,16] bf16, + stats[0:16]=sum, [16:32]=sumsq
// Each thread produces VOX consecutive-w output voxels x 16 channels, weights broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int C1_OUT = 16;
constexpr int C1_VOX = 4;

template <typename TIn>
__device__ __forceinline__ float load_in(const TIn* p);
template <> __device__ __forceinline__ float load_in<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_in<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TIn>
__global__ void __launch_bounds__(256, 1) conv1_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ w /*[16][27]*/,
                                                        __nv_bfloat16* __restrict__ y, float* __restrict__ stats, Dims d) {
    __shared__ float4 ws[27][4];              // [tap][co/4] -> 4 consecutive output channels
    __shared__ float red[2 * C1_OUT];
    for (int i = threadIdx.x; i < 27 * C1_OUT; i += blockDim.x) {
        const int tap = i / C1_OUT, co = i % C1_OUT;
        reinterpret_cast<float*>(&ws[tap][0])[co] = w[co * 27 + tap];
    }
    if (threadIdx.x < 2 * C1_OUT) red[threadIdx.x] = 0.f;
    __syncthreads();

    const int wgroups = (d.W + C1_VOX - 1) / C1_VOX;
    const long long total = (long long)d.N * d.D * d.H * wgroups;
    float s1[C1_OUT], s2[C1_OUT];
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }

    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int wg = (int)(g % wgroups);
        long long t = g / wgroups;
        const int h = (int)(t % d.H); t /= d.H;
        const int dd = (int)(t % d.D);
        const int n = (int)(t / d.D);
        const int w0 = wg * C1_VOX;

        float acc[C1_VOX][C1_OUT];
#pragma unroll
        for (int v = 0; v < C1_VOX; ++v)
#pragma unroll
            for (int c = 0; c < C1_OUT; ++c) acc[v][c] = 0.f;

        // all 9 x (VOX+2) neighbourhood values first (branch-free: clamped address x validity), so the loads
        // are in flight together and the 27 x 64 FMAs below run without further memory stalls
        float in[9][C1_VOX + 2];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int zd = dd + r / 3 - 1, zh = h + r % 3 - 1;
            const bool row_ok = zd >= 0 && zd < d.D && zh >= 0 && zh < d.H;
            const int cd = min(max(zd, 0), d.D - 1), chh = min(max(zh, 0), d.H - 1);
            const TIn* row = x + (((long long)n * d.D + cd) * d.H + chh) * d.W;
#pragma unroll
            for (int j = 0; j < C1_VOX + 2; ++j) {
                const int zw = w0 + j - 1;
                const int cw = min(max(zw, 0), d.W - 1);
                const float v = load_in<TIn>(row + cw);
                in[r][j] = (row_ok && zw >= 0 && zw < d.W) ? v : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < 9; ++r) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int tap = r * 3 + kw;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = ws[tap][q];
#pragma unroll
                    for (int v = 0; v < C1_VOX; ++v) {
                        const float xv = in[r][v + kw];
                        acc[v][4 * q + 0] = fmaf(xv, wv.x, acc[v][4 * q + 0]);
                        acc[v][4 * q + 1] = fmaf(xv, wv.y, acc[v][4 * q + 1]);
                        acc[v][4 * q + 2] = fmaf(xv, wv.z, acc[v][4 * q + 2]);
                        acc[v][4 * q + 3] = fmaf(xv, wv.w, acc[v][4 * q + 3]);
                    }
                }
            }
        }
        __nv_bfloat16* out = y + ((((long long)n * d.D + dd) * d.H + h) * d.W + w0) * C1_OUT;
#pragma unroll
        for (int v = 0; v < C1_VOX; ++v) {
            if (w0 + v >= d.W) break;
            float r[8], r2[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { r[c] = acc[v][c]; r2[c] = acc[v][8 + c]; }
            const uint4 lo = pack8(r), hi = pack8(r2);
            // statistics of the *stored* (bf16-rounded) values: exactly what BatchNorm normalises later
            float q[8];
            unpack8(lo, q);
#pragma unroll
            for (int c = 0; c < 8; ++c) { s1[c] += q[c]; s2[c] = fmaf(q[c], q[c], s2[c]); }
            unpack8(hi, q);
#pragma unroll
            for (int c = 0; c < 8; ++c) { s1[8 + c] += q[c]; s2[8 + c] = fmaf(q[c], q[c], s2[8 + c]); }
            reinterpret_cast<uint4*>(out + v * C1_OUT)[0] = lo;
            reinterpret_cast<uint4*>(out + v * C1_OUT)[1] = hi;
        }
    }
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) {
        const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
        if (lane_id() == 0) { atomicAdd(&red[c], a); atomicAdd(&red[C1_OUT + c], b); }
    }
    __syncthreads();
    if (threadIdx.x < 2 * C1_OUT) atomicAdd(&stats[threadIdx.x], red[threadIdx.x]);
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
// conv1 wgrad: dW[co, tap] += sum_vox dy[vox, co] * x[vox + tap]      (dy: [N,D,H,W,16] bf16, x: [N,D,H,W])
// lane = tap (27 of 32 lanes active), 16 output channels in registers, dy row broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int WG_TILE_W = 64;            // voxels (along w) staged per step

template <typename TIn>
__global__ void __launch_bounds__(128) conv1_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const TIn* __restrict__ x,
                                                          float* __restrict__ dw /*[16][27]*/, Dims d) {
    __shared__ __align__(16) float s_dy[4][WG_TILE_W][C1_OUT];             // per warp, already fp32
    __shared__ float s_x[4][3][3][WG_TILE_W + 2];                          // per warp: halo rows
    __shared__ float s_acc[C1_OUT * 27];
    for (int i = threadIdx.x; i < C1_OUT * 27; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int nwarps = blockDim.x >> 5;
    const int wtiles = (d.W + WG_TILE_W - 1) / WG_TILE_W;
    const long long total = (long long)d.N * d.D * d.H * wtiles;
    const int tap = lane < 27 ? lane : 26;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    float acc[C1_OUT];
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) acc[c] = 0.f;

    for (long long job = (long long)blockIdx.x * nwarps + warp; job < total; job += (long long)gridDim.x * nwarps) {
        const int wt = (int)(job % wtiles);
        long long t = job / wtiles;
        const int h = (int)(t % d.H); t /= d.H;
        const int dd = (int)(t % d.D);
        const int n = (int)(t / d.D);
        const int w0 = wt * WG_TILE_W;
        const int nw = min(WG_TILE_W, d.W - w0);
        __syncwarp();
        // stage dy rows (nw voxels x 16 ch) as fp32 and the 3x3 halo rows of x
        const uint4* src = reinterpret_cast<const uint4*>(dy + ((((long long)n * d.D + dd) * d.H + h) * d.W + w0) * C1_OUT);
        for (int i = lane; i < nw * 2; i += 32) {
            float f[8];
            unpack8(ld_stream_u4(src + i), f);
            float4* dst = reinterpret_cast<float4*>(&s_dy[warp][0][0]) + i * 2;
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
        for (int r = 0; r < 9; ++r) {
            const int zd = dd + r / 3 - 1, zh = h + r % 3 - 1;
            const bool ok = zd >= 0 && zd < d.D && zh >= 0 && zh < d.H;
            const TIn* row = x + (((long long)n * d.D + (ok ? zd : 0)) * d.H + (ok ? zh : 0)) * d.W;
            for (int j = lane; j < nw + 2; j += 32) {
                const int zw = w0 + j - 1;
                s_x[warp][r / 3][r % 3][j] = (ok && zw >= 0 && zw < d.W) ? load_in<TIn>(row + zw) : 0.f;
            }
        }
        __syncwarp();
        const float* xs = &s_x[warp][kd][kh][kw];
#pragma unroll 4
        for (int v = 0; v < nw; ++v) {
            const float xv = xs[v];
            const float4* g = reinterpret_cast<const float4*>(&s_dy[warp][v][0]);   // same address for all lanes: broadcast
            const float4 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
            acc[0] = fmaf(g0.x, xv, acc[0]);   acc[1] = fmaf(g0.y, xv, acc[1]);
            acc[2] = fmaf(g0.z, xv, acc[2]);   acc[3] = fmaf(g0.w, xv, acc[3]);
            acc[4] = fmaf(g1.x, xv, acc[4]);   acc[5] = fmaf(g1.y, xv, acc[5]);
            acc[6] = fmaf(g1.z, xv, acc[6]);   acc[7] = fmaf(g1.w, xv, acc[7]);
            acc[8] = fmaf(g2.x, xv, acc[8]);   acc[9] = fmaf(g2.y, xv, acc[9]);
            acc[10] = fmaf(g2.z, xv, acc[10]); acc[11] = fmaf(g2.w, xv, acc[11]);
            acc[12] = fmaf(g3.x, xv, acc[12]); acc[13] = fmaf(g3.y, xv, acc[13]);
            acc[14] = fmaf(g3.z, xv, acc[14]); acc[15] = fmaf(g3.w, xv, acc[15]);
        }
    }
    if (lane < 27) {
#pragma unroll
        for (int c = 0; c < C1_OUT; ++c) atomicAdd(&s_acc[c * 27 + lane], acc[c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C1_OUT * 27; i += blockDim.x) atomicAdd(&dw[i], s_acc[i]);
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

// x_dtype: 0 fp32, 1 bf16.  stats must be zeroed (2*16 floats).
COINN_API int coinn_conv1_fwd(const void* x, int x_dtype, const float* w, void* y, float* stats,
                              int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    Dims d{N, D, H, W};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long groups = (long long)N * D * H * ((W + C1_VOX - 1) / C1_VOX);
    const int grid = grid_for(groups, 256, 2);
    if (x_dtype != 0) return (int)cudaErrorInvalidValue;       // callers up-cast bf16 volumes (8.5 MB) first
    conv1_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, w, (__nv_bfloat16*)y, stats, d);
    COINN_CHECK_LAUNCH();
    return 0;
}

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

// dw[16*27] must be zeroed.
COINN_API int coinn_conv1_wgrad(const void* dy, const void* x, int x_dtype, float* dw, int N, int D, int H, int W, void* stream) {
    using namespace coinn;
    Dims d{N, D, H, W};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long jobs = (long long)N * D * H * ((W + WG_TILE_W - 1) / WG_TILE_W);
    long long want = (jobs + 4 * 4 - 1) / (4 * 4);
    const int grid = (int)(want < 1 ? 1 : (want > 8LL * B200_SM_COUNT ? 8LL * B200_SM_COUNT : want));
    if (x_dtype == 0) conv1_wgrad_kernel<float><<<grid, 128, 0, st>>>((const __nv_bfloat16*)dy, (const float*)x, dw, d);
    else conv1_wgrad_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dw, d);
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
