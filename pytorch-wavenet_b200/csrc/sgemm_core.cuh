// sgemm_core.cuh -- the register-tiled fp32 SGEMM machinery shared by the training forward and backward kernels:
// 16x16 threads, (TM/16) x 8 accumulators per thread, K slabs of 16 staged through shared memory with register
// prefetch; A operands come from "loader" functors (tap gathers, relu'd rows, ...), B operands are K-outer
// weight matrices [K][ldw] in global memory.
#pragma once
#include "common.cuh"

namespace wn {

constexpr int NT = 256;   // threads per CTA: 16 (tx, output columns) x 16 (ty, frames)
constexpr int KS = 16;    // K slab
constexpr int NC = 128;   // output columns per chunk: thread owns cols tx*4..+3 and 64+tx*4..+3
constexpr int ZPAD = 4;   // Zs row pitch = TM + ZPAD floats

__host__ __device__ inline int n1p_of(int D) { return ((D + 63) / 64) * 128; }
__host__ __device__ inline int n2p_of(int N) { return ((N + 127) / 128) * 128; }

// ------------------------------------------------------------------------------------------------ A loaders
// An A loader yields element (m, k) of the CTA's [TM x K] left operand; `vec` says whether 4 consecutive k
// starting at a multiple of 4 are contiguous and 16-byte aligned in global memory (and never straddle a tap).
struct TapLoader {              // residual block: A row m = taps of h_in around frame t0+m
    const float* h;             // h_in + b*L*R
    int R, ktaps, dil, t0, L, in_start, K;
    bool vec;
    __device__ __forceinline__ const float* addr(int m, int kidx, bool& ok) const {
        const int j = kidx / R, c = kidx - j * R;
        const int t = t0 + m, ts = t - (ktaps - 1 - j) * dil;
        ok = (kidx < K) && (t < L) && (ts >= in_start);
        return h + (size_t)ts * R + c;
    }
    __device__ __forceinline__ float load1(int m, int kidx) const {
        bool ok; const float* p = addr(m, kidx, ok);
        return ok ? __ldg(p) : 0.f;
    }
    __device__ __forceinline__ float4 load4(int m, int kidx) const {
        bool ok; const float* p = addr(m, kidx, ok);
        return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
};

struct ReluRowLoader {          // head: A row m = relu(skip[frame t0+m])
    const float* s;             // skip + b*Tsk*S, indexed from skip_start
    int S, t0, L, skip_start;
    bool vec;
    __device__ __forceinline__ float load1(int m, int kidx) const {
        const int t = t0 + m;
        if (kidx >= S || t >= L) return 0.f;
        return fmaxf(__ldg(s + (size_t)(t - skip_start) * S + kidx), 0.f);
    }
    __device__ __forceinline__ float4 load4(int m, int kidx) const {
        const int t = t0 + m;
        if (kidx >= S || t >= L) return make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = __ldg(reinterpret_cast<const float4*>(s + (size_t)(t - skip_start) * S + kidx));
        return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    }
};

struct ColumnLoader {           // dense start conv: A[m][k] = x[b][k][t0+m]  (time contiguous)
    const float* x;             // x + b*classes*L
    int classes, t0, L;
    static constexpr bool vec = false;
    __device__ __forceinline__ float load1(int m, int kidx) const {
        const int t = t0 + m;
        return (kidx < classes && t < L) ? __ldg(x + (size_t)kidx * L + t) : 0.f;
    }
    __device__ __forceinline__ float4 load4(int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
};

// ------------------------------------------------------------------------------------------------ SGEMM core
template <int TM>
struct Tile {
    static constexpr int MI = TM / 16;                       // frames per thread
    static constexpr int AV = (TM * 4 + NT - 1) / NT;        // float4 A-vectors per thread per slab
    __device__ static __forceinline__ int row(int ty, int i) {
        if constexpr (MI == 8) return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
        else return ty * MI + i;
    }
};

template <int TM>
__device__ __forceinline__ void load_a_frag(float (&a)[TM / 16], const float* __restrict__ src, int ty) {
    constexpr int MI = TM / 16;
    if constexpr (MI == 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + ty * 4);
        const float4 v1 = *reinterpret_cast<const float4*>(src + 64 + ty * 4);
        a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
    } else if constexpr (MI == 4) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + ty * 4);
        a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
    } else if constexpr (MI == 2) {
        const float2 v0 = *reinterpret_cast<const float2*>(src + ty * 2);
        a[0] = v0.x; a[1] = v0.y;
    } else {
        a[0] = src[ty];
    }
}

// acc[TM/16][8] += A[TM x K] * W_t[K x 128-col chunk at col0].  A comes either from a loader (staged through As)
// or from the resident Zs (K-outer, pitch TM+ZPAD).  As: [2][KS][TM], Bs: [2][KS][NC].
template <int TM, bool A_FROM_Z, class ALoader>
__device__ __forceinline__ void mainloop(float (&acc)[TM / 16][8], const ALoader& al, const float* __restrict__ Zs,
                                         const float* __restrict__ w_t, int ldw, int col0, int K,
                                         float* __restrict__ As, float* __restrict__ Bs) {
    using T = Tile<TM>;
    constexpr int MI = T::MI;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nks = (K + KS - 1) / KS;

    float4 bpre[2];
    float4 apre_v[T::AV];
    float apre_s[MI];

    auto gload = [&](int ks) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int v = tid + i * NT, r = v >> 5, c4 = v & 31, kk = ks * KS + r;
            bpre[i] = (kk < K) ? __ldg(reinterpret_cast<const float4*>(w_t + (size_t)kk * ldw + col0 + c4 * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if constexpr (!A_FROM_Z) {
            if (al.vec) {
#pragma unroll
                for (int i = 0; i < T::AV; ++i) {
                    const int v = tid + i * NT;
                    if (v < TM * 4) apre_v[i] = al.load4(v % TM, ks * KS + (v / TM) * 4);
                }
            } else {
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int e = tid + i * NT;
                    apre_s[i] = al.load1(e % TM, ks * KS + e / TM);
                }
            }
        }
    };
    auto sstore = [&](int buf) {
        float* bs = Bs + buf * KS * NC;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int v = tid + i * NT, r = v >> 5, c4 = v & 31;
            *reinterpret_cast<float4*>(bs + r * NC + c4 * 4) = bpre[i];
        }
        if constexpr (!A_FROM_Z) {
            float* as = As + buf * KS * TM;
            if (al.vec) {
#pragma unroll
                for (int i = 0; i < T::AV; ++i) {
                    const int v = tid + i * NT;
                    if (v < TM * 4) {
                        const int m = v % TM, q = v / TM;
                        as[(q * 4 + 0) * TM + m] = apre_v[i].x;
                        as[(q * 4 + 1) * TM + m] = apre_v[i].y;
                        as[(q * 4 + 2) * TM + m] = apre_v[i].z;
                        as[(q * 4 + 3) * TM + m] = apre_v[i].w;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int e = tid + i * NT;
                    as[(e / TM) * TM + (e % TM)] = apre_s[i];
                }
            }
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    for (int ks = 0; ks < nks; ++ks) {
        const int buf = ks & 1;
        if (ks + 1 < nks) gload(ks + 1);
        const float* bs = Bs + buf * KS * NC;
        const float* as = A_FROM_Z ? (Zs + (size_t)ks * KS * (TM + ZPAD)) : (As + buf * KS * TM);
        constexpr int APITCH = A_FROM_Z ? (TM + ZPAD) : TM;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            float a[MI], b[8];
            load_a_frag<TM>(a, as + kk * APITCH, ty);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * NC + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + kk * NC + 64 + tx * 4);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (ks + 1 < nks) sstore(buf ^ 1);
        __syncthreads();
    }
}


}  // namespace wn
