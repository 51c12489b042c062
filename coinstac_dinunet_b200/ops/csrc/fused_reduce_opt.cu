// Fused cross-GPU gradient reduce + scale + cast + optimizer update, the dSGD data plane.
//
// Replaces the reference's per-step round trip  grads.npy -> remote mean -> avg_grads.npy ->
// optimizer.step()  (coinstac_dinunet/distrib/learner.py:20-59, reducer.py:25-54) by
// ONE kernel per bucket that every site (GPU) launches on its own stream:
//
//   start barrier (all peers' gradients are final)            st.release.sys / ld.acquire.sys flags
//   g = (1/S) * sum_p grad_p[i]          peer loads over NVLink (P2P) or multimem.ld_reduce (NVLS)
//   Adam / AdamW / SGD on fp32 master weights (+ optional bf16 shadow copy for the compute path)
//   two-shot / NVLS: store the updated shard into every peer's parameter arena
//   end barrier (peers are done with my memory), then zero my gradient slices for the next step
//
// Variants:
//   ONE_SHOT  every GPU reduces the whole bucket redundantly; ingress (S-1)*B; latency regime
//   TWO_SHOT  GPU r reduces + updates shard r, then broadcasts parameters; 2*(S-1)/S*B; bandwidth regime
//   NVLS      two-shot with in-switch reduction (multimem.ld_reduce) and multicast store (multimem.st)
//   S == 1    degenerates to a fused local optimizer step (no barriers)
// The sum runs in fixed rank order, so replicas stay bit-identical (tests/test_fused_reduce_gpu.py).
// No tensor cores here on purpose: the op is elementwise, it is bound by NVLink/HBM bytes.
#include "common.cuh"

namespace coinn {

constexpr int kMaxRanks = 8;
constexpr int kMaxBlocks = 296;   // flag slots per rank (grid is capped to the resident CTA count)
constexpr int kThreads = 512;

enum Variant { ONE_SHOT = 0, TWO_SHOT = 1, NVLS = 2 };
enum OptKind { ADAM = 0, ADAMW = 1, SGD = 2, NONE = 3 };   // NONE: plain all-reduce(mean) into the parameter arena
enum GradType { G_F32 = 0, G_BF16 = 1, G_F16 = 2 };

struct FusedArgs {
    const void* grad_ptrs[kMaxRanks];   // every rank's gradient arena (peer-mapped), indexed by rank
    float*      param_ptrs[kMaxRanks];  // every rank's fp32 master-parameter arena
    void*       shadow_ptrs[kMaxRanks]; // every rank's bf16 shadow-parameter arena (or null)
    uint32_t*   flag_ptrs[kMaxRanks];   // every rank's flag pad  [kMaxBlocks][kMaxRanks]
    const void* grad_mc;                // multicast alias of the gradient arena (NVLS) or null
    float*      param_mc;               // multicast alias of the parameter arena (NVLS) or null
    void*       shadow_mc;
    float*      m;                      // local optimizer state
    float*      v;
    uint32_t*   epoch;                  // local, [kMaxBlocks]: per-CTA barrier sequence number
    int*        step;                   // local device step counter (Adam bias correction)
    uint32_t*   ticket;                 // local, last-CTA detection
    const float* lr_ptr;                // optional device-side learning rate (graph-safe schedules)
    float*      grad32;                 // grad_dtype != F32: the local fp32 gradient arena autograd accumulates into; it is
                                        // packed into this rank's 16-bit wire buffer (= grad_ptrs[rank]) before the start barrier
    int*        error;                  // local watchdog word: 0 = healthy, 1 + peer = that peer never arrived at a barrier
    long long   offset;                 // first element of the bucket (multiple of 4)
    long long   numel;                  // elements in the bucket (multiple of 4)
    float lr, beta1, beta2, eps, weight_decay, grad_scale, momentum;
    int rank, world, variant, opt_kind, grad_dtype, zero_grads, bump_step, nesterov;
    unsigned timeout_ms;                // bounded spin of the cross-GPU barrier (0 = spin forever)
    int _pad;
};

// ------------------------------------------------------------------------------------------------
template <int GD> struct GradIO;
template <> struct GradIO<G_F32> {
    static __device__ __forceinline__ void store(void*, long long, const float4&) {}
    static __device__ __forceinline__ float4 load(const void* base, long long vec) {
        return ld_stream_f4(reinterpret_cast<const float4*>(base) + vec);
    }
    static __device__ __forceinline__ float4 load_mc(const void* mc, long long vec) {
        return multimem_ld_reduce_f4(reinterpret_cast<const float4*>(mc) + vec);
    }
    static __device__ __forceinline__ void zero(void* base, long long vec) {
        reinterpret_cast<float4*>(base)[vec] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
};
template <> struct GradIO<G_BF16> {
    static __device__ __forceinline__ void store(void* base, long long vec, const float4& g) {
        reinterpret_cast<uint2*>(base)[vec] = make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
    }
    static __device__ __forceinline__ float4 load(const void* base, long long vec) {
        uint2 u = ld_stream_u2(reinterpret_cast<const uint2*>(base) + vec);
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float4 load_mc(const void* mc, long long vec) {
        // 64-bit in-switch reduce of 4 bf16 with fp32 accumulation
        uint2 u;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v2.bf16x2 {%0,%1}, [%2];"
                     : "=r"(u.x), "=r"(u.y) : "l"(reinterpret_cast<const uint2*>(mc) + vec) : "memory");
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ void zero(void* base, long long vec) {
        reinterpret_cast<uint2*>(base)[vec] = make_uint2(0u, 0u);
    }
};
template <> struct GradIO<G_F16> {
    static __device__ __forceinline__ void store(void* base, long long vec, const float4& g) {
        __half2 a = __floats2half2_rn(g.x, g.y), b = __floats2half2_rn(g.z, g.w);
        reinterpret_cast<uint2*>(base)[vec] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
    static __device__ __forceinline__ float4 load(const void* base, long long vec) {
        uint2 u = ld_stream_u2(reinterpret_cast<const uint2*>(base) + vec);
        float2 a = unpack_f16x2(u.x), b = unpack_f16x2(u.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float4 load_mc(const void* mc, long long vec) {
        uint2 u;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v2.f16x2 {%0,%1}, [%2];"
                     : "=r"(u.x), "=r"(u.y) : "l"(reinterpret_cast<const uint2*>(mc) + vec) : "memory");
        float2 a = unpack_f16x2(u.x), b = unpack_f16x2(u.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ void zero(void* base, long long vec) {
        reinterpret_cast<uint2*>(base)[vec] = make_uint2(0u, 0u);
    }
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Cross-GPU barrier between the CTAs with the same blockIdx on every rank.  Flags only ever
// grow (sequence numbers), so there is no reset race; one writer per slot.
// Watchdog: the spin is bounded by `timeout_ms` of wall time (%globaltimer, polled every 1024 probes).  A CTA that
// gives up records `1 + peer` in the local error word; once the word is set every later barrier of this rank falls
// straight through, so a dead site costs one timeout, not a hung box - the host raises after the round
// (DistArena.check_health).  The numerical result of such a step is garbage by construction.
__device__ __forceinline__ void cta_barrier_all_ranks(const FusedArgs& a, int block, uint32_t seq) {
    __syncthreads();
    if (threadIdx.x < a.world) {
        const int peer = threadIdx.x;
        st_release_sys(a.flag_ptrs[peer] + block * kMaxRanks + a.rank, seq);
        const uint32_t* mine = a.flag_ptrs[a.rank] + block * kMaxRanks + peer;
        volatile int* err = a.error;
        if (!(err && *err)) {
            unsigned probes = 0;
            unsigned long long t0 = 0;
            while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
                if (a.timeout_ms && ((++probes & 1023u) == 0u)) {
                    const unsigned long long now = global_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > (unsigned long long)a.timeout_ms * 1000000ull) {
                        if (err) atomicCAS(a.error, 0, 1 + peer);
                        break;
                    }
                    if (err && *err) break;
                }
            }
        }
    }
    __syncthreads();
}

struct Hyper {
    float lr, b1, b2, eps, wd, mom, bc1_inv, bc2_rsqrt;
    int first_step, nesterov;
};

template <int OPT>
__device__ __forceinline__ void opt_update(float g, float& p, float& m, float& v, const Hyper& h) {
    if (OPT == NONE) {
        p = g;                                   // the reduced value itself (PowerSGD P/Q factors, rank-1 grads, ...)
    } else if (OPT == ADAM || OPT == ADAMW) {
        if (OPT == ADAMW) p *= (1.f - h.lr * h.wd);
        else g = fmaf(h.wd, p, g);
        m = fmaf(h.b1, m, (1.f - h.b1) * g);
        v = fmaf(h.b2, v, (1.f - h.b2) * g * g);
        const float denom = sqrtf(v) * h.bc2_rsqrt + h.eps;
        p -= (h.lr * h.bc1_inv) * (m / denom);
    } else {
        g = fmaf(h.wd, p, g);
        if (h.mom != 0.f) {
            m = h.first_step ? g : fmaf(h.mom, m, g);
            g = h.nesterov ? fmaf(h.mom, m, g) : m;
        }
        p -= h.lr * g;
    }
}

template <int GD, int OPT, int VAR>
__global__ void __launch_bounds__(kThreads, 1) fused_reduce_opt_kernel(const FusedArgs a) {
    const int b = blockIdx.x, G = gridDim.x, S = a.world, r = a.rank, tid = threadIdx.x;
    const bool multi = S > 1;

    uint32_t seq = 0;
    if (multi) seq = a.epoch[b] + 1u;

    // ---- hyper-parameters (bias corrections from the device-side step counter) ----
    Hyper h;
    const int t = *a.step + 1;
    h.lr = a.lr_ptr ? *a.lr_ptr : a.lr;
    h.b1 = a.beta1; h.b2 = a.beta2; h.eps = a.eps; h.wd = a.weight_decay; h.mom = a.momentum;
    h.first_step = (t == 1); h.nesterov = a.nesterov;
    if (OPT == ADAM || OPT == ADAMW) {
        const double bc1 = 1.0 - pow((double)a.beta1, (double)t);
        const double bc2 = 1.0 - pow((double)a.beta2, (double)t);
        h.bc1_inv = (float)(1.0 / bc1);
        h.bc2_rsqrt = (float)(1.0 / sqrt(bc2));
    } else { h.bc1_inv = 1.f; h.bc2_rsqrt = 1.f; }

    // ---- which vectors (4 elements each) does this CTA own? ----
    const long long nvec = a.numel >> 2, off = a.offset >> 2;
    long long shard = nvec, s_lo = 0, s_hi = nvec;
    if (multi && VAR != ONE_SHOT) {
        shard = (nvec + S - 1) / S;
        s_lo = min((long long)r * shard, nvec);
        s_hi = min(s_lo + shard, nvec);
    }
    const long long chunk = (shard + G - 1) / G;     // identical on every rank
    const long long lo = min(s_lo + (long long)b * chunk, s_hi), hi = min(lo + chunk, s_hi);

    // ---- 16-bit wire (precision_bits = 16): pack my fp32 gradients into my wire buffer - exactly the slices that the
    //      CTAs with MY block index read on the peers, so the per-block barrier below orders pack -> peer loads ----
    if (GD != G_F32 && multi) {
        float4* g32 = reinterpret_cast<float4*>(a.grad32) + off;
        void* wire = const_cast<void*>(a.grad_ptrs[r]);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int nq = (VAR == ONE_SHOT) ? 1 : S;
        for (int q = 0; q < nq; ++q) {
            long long z_lo = lo, z_hi = hi;
            if (VAR != ONE_SHOT) {
                const long long q_lo = min((long long)q * shard, nvec), q_hi = min(q_lo + shard, nvec);
                z_lo = min(q_lo + (long long)b * chunk, q_hi); z_hi = min(z_lo + chunk, q_hi);
            }
            for (long long i = z_lo + tid; i < z_hi; i += kThreads) {
                GradIO<GD>::store(wire, off + i, g32[i]);
                if (a.zero_grads) g32[i] = z4;         // the fp32 arena is private: re-zero it right here
            }
        }
        __threadfence_system();
    }

    if (multi) cta_barrier_all_ranks(a, b, 2u * seq - 1u);

    float4* __restrict__ P = reinterpret_cast<float4*>(a.param_ptrs[r]) + off;
    float4* __restrict__ M = reinterpret_cast<float4*>(a.m) + off;
    float4* __restrict__ V = reinterpret_cast<float4*>(a.v) + off;
    uint2* SH = a.shadow_ptrs[r] ? reinterpret_cast<uint2*>(a.shadow_ptrs[r]) + off : nullptr;

    for (long long i = lo + tid; i < hi; i += kThreads) {
        float4 g;
        if (multi && VAR == NVLS) {
            g = GradIO<GD>::load_mc(a.grad_mc, off + i);
        } else {
            float4 acc[kMaxRanks];
#pragma unroll
            for (int p = 0; p < kMaxRanks; ++p)
                if (p < S) acc[p] = GradIO<GD>::load(a.grad_ptrs[p], off + i);   // S loads in flight
            g = acc[0];
#pragma unroll
            for (int p = 1; p < kMaxRanks; ++p)
                if (p < S) { g.x += acc[p].x; g.y += acc[p].y; g.z += acc[p].z; g.w += acc[p].w; }
        }
        const float sc = a.grad_scale;
        g.x *= sc; g.y *= sc; g.z *= sc; g.w *= sc;

        float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), m4 = p4, v4 = p4;
        if (OPT != NONE) { p4 = P[i]; m4 = M[i]; }
        if (OPT == ADAM || OPT == ADAMW) v4 = V[i];
        opt_update<OPT>(g.x, p4.x, m4.x, v4.x, h);
        opt_update<OPT>(g.y, p4.y, m4.y, v4.y, h);
        opt_update<OPT>(g.z, p4.z, m4.z, v4.z, h);
        opt_update<OPT>(g.w, p4.w, m4.w, v4.w, h);
        if (OPT != NONE) M[i] = m4;
        if (OPT == ADAM || OPT == ADAMW) V[i] = v4;

        const uint2 sh = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
        if (multi && VAR == NVLS) {
            multimem_st_f4(reinterpret_cast<float4*>(a.param_mc) + off + i, p4);                   // lands on every GPU
            if (a.shadow_mc) multimem_st_u2(reinterpret_cast<uint2*>(a.shadow_mc) + off + i, sh);
        } else {
            P[i] = p4;
            if (SH) SH[i] = sh;
            if (multi && VAR == TWO_SHOT) {
#pragma unroll
                for (int q = 0; q < kMaxRanks; ++q) {
                    if (q < S && q != r) {
                        st_stream_f4(reinterpret_cast<float4*>(a.param_ptrs[q]) + off + i, p4);
                        if (a.shadow_ptrs[q]) st_stream_u2(reinterpret_cast<uint2*>(a.shadow_ptrs[q]) + off + i, sh);
                    }
                }
            }
        }
    }

    if (multi) {
        __threadfence_system();                    // parameter stores visible before the flag
        cta_barrier_all_ranks(a, b, 2u * seq);
        if (tid == 0) a.epoch[b] = seq;
    }

    // ---- my gradient slices have been consumed by every peer: clear them for the next step ----
    if (a.zero_grads && (GD == G_F32 || !multi)) {   // (16-bit wire: the fp32 arena was re-zeroed while packing)
        void* mine = const_cast<void*>(a.grad_ptrs[r]);
        if (!multi || VAR == ONE_SHOT) {
            for (long long i = lo + tid; i < hi; i += kThreads) GradIO<GD>::zero(mine, off + i);
        } else {
            for (int q = 0; q < S; ++q) {          // slice (q, b): read by rank q's CTA b
                const long long q_lo = min((long long)q * shard, nvec), q_hi = min(q_lo + shard, nvec);
                const long long z_lo = min(q_lo + (long long)b * chunk, q_hi), z_hi = min(z_lo + chunk, q_hi);
                for (long long i = z_lo + tid; i < z_hi; i += kThreads) GradIO<GD>::zero(mine, off + i);
            }
        }
    }

    // ---- last CTA of the launch advances the step counter (everyone has read it by now) ----
    if (a.bump_step) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned tk = atomicAdd(a.ticket, 1u);
            if (tk == (unsigned)G - 1u) { *a.ticket = 0u; *a.step = t; }
        }
    }
}

template <int GD, int OPT>
static cudaError_t launch_var(const FusedArgs& a, int grid, cudaStream_t st) {
    switch (a.variant) {
        case ONE_SHOT: fused_reduce_opt_kernel<GD, OPT, ONE_SHOT><<<grid, kThreads, 0, st>>>(a); break;
        case TWO_SHOT: fused_reduce_opt_kernel<GD, OPT, TWO_SHOT><<<grid, kThreads, 0, st>>>(a); break;
        case NVLS:     fused_reduce_opt_kernel<GD, OPT, NVLS><<<grid, kThreads, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
template <int GD>
static cudaError_t launch_opt(const FusedArgs& a, int grid, cudaStream_t st) {
    switch (a.opt_kind) {
        case ADAM:  return launch_var<GD, ADAM>(a, grid, st);
        case ADAMW: return launch_var<GD, ADAMW>(a, grid, st);
        case SGD:   return launch_var<GD, SGD>(a, grid, st);
        case NONE:  return launch_var<GD, NONE>(a, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace coinn

// grid <= 0 -> chosen here: enough CTAs to cover the bucket, never more than the co-resident
// count (1 CTA/SM because the cross-GPU barrier spins).
COINN_API int coinn_fused_reduce_opt(const coinn::FusedArgs* args, int grid, void* stream) {
    using namespace coinn;
    FusedArgs a = *args;
    if (a.world < 1 || a.world > kMaxRanks || (a.numel & 3) || (a.offset & 3)) return (int)cudaErrorInvalidValue;
    if (a.numel == 0) return 0;
    if (a.world == 1) a.grad_dtype = G_F32;                       // no wire without peers
    if (a.grad_dtype != G_F32 && a.grad32 == nullptr) return (int)cudaErrorInvalidValue;
    const long long nvec = a.numel >> 2;
    long long per_rank = (a.world > 1 && a.variant != ONE_SHOT) ? (nvec + a.world - 1) / a.world : nvec;
    if (grid <= 0) {
        long long want = (per_rank + kThreads * 4 - 1) / (kThreads * 4);   // >= 4 vectors per thread
        grid = (int)(want < 1 ? 1 : (want > B200_SM_COUNT ? B200_SM_COUNT : want));
    }
    if (grid > kMaxBlocks) grid = kMaxBlocks;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    switch (a.grad_dtype) {
        case G_F32:  e = launch_opt<G_F32>(a, grid, st); break;
        case G_BF16: e = launch_opt<G_BF16>(a, grid, st); break;
        case G_F16:  e = launch_opt<G_F16>(a, grid, st); break;
        default: e = cudaErrorInvalidValue;
    }
    return (int)e;
}

COINN_API int coinn_fused_flag_slots() { return coinn::kMaxBlocks * coinn::kMaxRanks; }
COINN_API int coinn_fused_max_blocks() { return coinn::kMaxBlocks; }
COINN_API int coinn_fused_args_size() { return (int)sizeof(coinn::FusedArgs); }
