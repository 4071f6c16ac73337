// gen.cu -- Fast-WaveNet sampling as ONE persistent cooperative kernel.
//
// Replaces (reference file:line): WaveNetModel.generate_fast's warm-up and sampling loops wavenet_model.py:260-302,
// queue_dilate :177-184, the layer loop :131-165 and head :167-169 evaluated on single columns, and
// DilatedQueue.enqueue/dequeue/reset wavenet_modules.py:55-77.
//
// Work decomposition.  One network evaluation is a chain of matrix-vector products with a full dependency
// between stages, so the whole GPU works on every stage and stages are separated by a grid barrier:
//   per layer   stage 1: rows (f_c, g_c) of the k-tap conv for the CTA's dilation channels c -> z_c (exchange buffer)
//               stage 2: rows of residual_conv (-> next layer's ring slot for time t) and of skip_conv
//                        (-> running skip sum, kept in shared memory of the owning CTA across all layers)
//   head        end_conv_1 rows -> end_conv_2 rows -> every CTA redundantly picks the next sample
// Row r of a stage belongs to CTA (r mod G); one warp computes one row for all streams (lanes split K, butterfly
// reduction), so a stream's arithmetic does not depend on how many streams run beside it.
// The rings ("dilated queues") are the exchange medium between layers: ring l has (k-1)*d_l+1 slots of
// [n_streams][R]; slot (t mod len) holds the layer's input at time t, unwritten slots are zero (reset).
// Exchanged vectors are read with ld.global.cg (L2) because L1 is not coherent across SMs.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace wn {

constexpr int GEN_NT = 256;
constexpr int GEN_WARPS = GEN_NT / 32;

struct GenLayer {
    const float *wf, *wg, *wr, *ws, *bf, *bg, *br, *bs;
    long long ring_off;     // in floats, from rings base
    int dil, ring_len;
};

struct GenParams {
    const GenLayer* layers;
    int n_layers, k, R, D, S, E, C, NS;
    const float *start_w, *start_b, *e1w, *e1b, *e2w, *e2b;
    float* rings;
    float *zbuf, *skipbuf, *y1buf, *logitbuf;
    int* cur_idx;
    unsigned* bar;
    // run
    const int* first; int n_given;
    const int* forced; const double* uniforms;
    int* out_idx; float* out_logits;
    int n_samples, t0, n_evals;
    float temperature, regularize;
    // smem carve (floats)
    int regA, regB, pre_n, skacc_n;
    // ---- flag-in-data ("LL") exchange kernel
    uint2* ringLL;          // same geometry as rings, 8-byte {value, tag} elements
    uint2 *zLL, *skipLL, *y1LL, *logitLL;      // [2 parities][...] exchange vectors
    int* err;               // set to 1 by a CTA that timed out waiting for a tag
    int part_n;             // partial-sum scratch (floats)
    int wslot_floats, n_wslots;                // weight prefetch ring in shared memory (0 slots: read weights via L2)
    long long* trace;       // optional: clock64 stamps of CTA 0 / thread 0 during the last evaluation (wn_gen_read_trace)
    const unsigned char* cl8_img;              // batched cluster kernel: fragment-ordered bf16 hi/lo weight images (cl8_pack_kernel)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Row ownership: CTA c owns the contiguous rows [c*per, min(N, (c+1)*per)), per = ceil(N/G).  Contiguous (not
// interleaved) so that what a CTA publishes per stage is one run of {value,tag} pairs: with >= 4 rows per CTA the
// stores of one warp instruction fill whole 32-byte sectors, which measured 2.6x faster to exchange than 16-byte
// partial-sector writes from twice as many CTAs (tools/lat_probe.cu).
__device__ __forceinline__ int own_per(int N, int G) { return (N + G - 1) / G; }
__device__ __forceinline__ int own_cnt(int N, int G, int cta) {
    const int per = own_per(N, G), n = N - cta * per;
    return n < 0 ? 0 : (n < per ? n : per);
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target, unsigned G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += G;
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

// dot products of one weight row (K floats, global) with the vectors xs[s][0..K) (shared) for streams s0..s0+SB
template <int SB>
__device__ __forceinline__ void row_dot(const float* __restrict__ w, const float* __restrict__ xs, int K, int NS,
                                        int s0, int lane, float (&acc)[SB]) {
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = 0.f;
    if ((K & 3) == 0) {
        const float4* w4p = reinterpret_cast<const float4*>(w);
        for (int i4 = lane; i4 < (K >> 2); i4 += 32) {
            const float4 w4 = __ldg(w4p + i4);
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                if (s0 + j < NS) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xs + (size_t)(s0 + j) * K + 4 * i4);
                    float a = acc[j];
                    a = fmaf(w4.x, x4.x, a); a = fmaf(w4.y, x4.y, a); a = fmaf(w4.z, x4.z, a); a = fmaf(w4.w, x4.w, a);
                    acc[j] = a;
                }
            }
        }
    } else {
        for (int i = lane; i < K; i += 32) {
            const float wv = __ldg(w + i);
#pragma unroll
            for (int j = 0; j < SB; ++j)
                if (s0 + j < NS) acc[j] = fmaf(wv, xs[(size_t)(s0 + j) * K + i], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = warp_sum(acc[j]);
}

// Activations of the sampler kernels (all three kernels use these, so they stay bit-identical to each other):
// ex2-based exp and fast division, absolute error ~2e-7 -- the same order as the difference between expf and the
// reference's CPU math, at a fraction of the instructions on the latency-critical path.
__device__ __forceinline__ float sigmoid_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 2.f * sigmoid_(2.f * x) - 1.f; }

template <int SB>
__global__ void __launch_bounds__(GEN_NT, 1) gen_kernel(const GenParams p) {
    extern __shared__ __align__(16) float sm[];
    float* regA = sm;                          // stage-1 inputs [NS][k*R] / head input [NS][S]
    float* regB = regA + p.regA;               // z [NS][D] / y1 [NS][E]
    float* pre = regB + p.regB;                // per-item results [items][NS]
    float* skacc = pre + p.pre_n;              // running skip sums of the rows this CTA owns [rows][NS]
    int* idx_s = reinterpret_cast<int*>(skacc + p.skacc_n);     // current input index per stream [NS]
    float* prob = reinterpret_cast<float*>(idx_s + p.NS);       // [GEN_WARPS][C] softmax scratch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int NS = p.NS, R = p.R, D = p.D, S = p.S, E = p.E, C = p.C, k = p.k;
    const int K1 = k * R;
    unsigned bar_target = 0;

    // owned rows per stage: row r belongs to CTA (r mod G); local index r / G
    const int nD = own_cnt(D, G, cta), oD = cta * own_per(D, G);
    const int nR = own_cnt(R, G, cta), oR = cta * own_per(R, G);
    const int nS = own_cnt(S, G, cta), oS = cta * own_per(S, G);
    const int nE = own_cnt(E, G, cta), oE = cta * own_per(E, G);
    const int nC = own_cnt(C, G, cta), oC = cta * own_per(C, G);

    // the index chosen by the evaluation before t0 (continuation of a previous launch)
    for (int s = tid; s < NS; s += GEN_NT) idx_s[s] = p.cur_idx[s];
    __syncthreads();

    for (int ev = 0; ev < p.n_evals; ++ev) {
        const int t = p.t0 + ev;                           // absolute evaluation counter == time
        const bool want_head = (t >= p.n_given - 1);
        const int samp = t - (p.n_given - 1);              // sample number this evaluation chooses
        // ---- input index of this evaluation
        if (t < p.n_given) {
            for (int s = tid; s < NS; s += GEN_NT) idx_s[s] = p.first[(size_t)s * p.n_given + t];
        } else if (p.forced != nullptr) {
            for (int s = tid; s < NS; s += GEN_NT) idx_s[s] = p.forced[(size_t)s * p.n_samples + (t - p.n_given)];
        }
        for (int i = tid; i < nS * NS; i += GEN_NT) skacc[i] = 0.f;
        __syncthreads();

        for (int l = 0; l < p.n_layers; ++l) {
            const GenLayer L = p.layers[l];
            float* ring = p.rings + L.ring_off;
            const int slot_t = t % L.ring_len;
            // ---- gather stage-1 inputs, interleaved like a conv weight row: regA[s][r*k + j] = tap j of channel r
            for (int i = tid; i < NS * K1; i += GEN_NT) {
                const int s = i / K1, rem = i - s * K1, r = rem / k, j = rem - r * k;
                float v;
                if (j == k - 1 && l == 0) {                 // current input of layer 0 = start conv column
                    int c = idx_s[s];
                    c = c < 0 ? 0 : (c >= C ? C - 1 : c);
                    v = __ldg(p.start_w + (size_t)r * C + c) + (p.start_b ? __ldg(p.start_b + r) : 0.f);
                    if (r >= oR && r < oR + nR) ring[((size_t)slot_t * NS + s) * R + r] = v;       // enqueue (owner writes)
                } else {
                    int tt = t - (k - 1 - j) * L.dil;        // time of tap j
                    int slot = tt % L.ring_len;
                    if (slot < 0) slot += L.ring_len;        // never-written slot: zero history
                    v = __ldcg(ring + ((size_t)slot * NS + s) * R + r);
                }
                regA[i] = v;
            }
            __syncthreads();
            // ---- stage 1: filter / gate rows of the owned dilation channels
            for (int it = warp; it < 2 * nD; it += GEN_WARPS) {
                const int c = oD + (it >> 1);
                const float* w = ((it & 1) ? L.wg : L.wf) + (size_t)c * K1;
                const float* bp = (it & 1) ? L.bg : L.bf;
                const float bias = bp ? __ldg(bp + c) : 0.f;
                for (int s0 = 0; s0 < NS; s0 += SB) {
                    float acc[SB];
                    row_dot<SB>(w, regA, K1, NS, s0, lane, acc);
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < SB; ++j)
                            if (s0 + j < NS) pre[it * NS + s0 + j] = acc[j] + bias;
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < nD * NS; i += GEN_NT) {
                const int ci = i / NS, s = i - ci * NS;
                const float z = tanh_(pre[(2 * ci) * NS + s]) * sigmoid_(pre[(2 * ci + 1) * NS + s]);
                p.zbuf[(size_t)s * D + oD + ci] = z;
            }
            grid_barrier(p.bar, bar_target, G);
            // ---- stage 2: residual rows (not needed after the last layer) and skip rows (not needed in warm-up)
            for (int i = tid; i < NS * D; i += GEN_NT) regB[i] = __ldcg(p.zbuf + i);
            __syncthreads();
            const int nres = (l + 1 < p.n_layers) ? nR : 0;
            const int nskp = want_head ? nS : 0;
            for (int it = warp; it < nres + nskp; it += GEN_WARPS) {
                const bool is_res = it < nres;
                const int li = is_res ? it : it - nres;
                const int row = (is_res ? oR : oS) + li;
                const float* w = (is_res ? L.wr : L.ws) + (size_t)row * D;
                const float* bp = is_res ? L.br : L.bs;
                const float bias = bp ? __ldg(bp + row) : 0.f;
                for (int s0 = 0; s0 < NS; s0 += SB) {
                    float acc[SB];
                    row_dot<SB>(w, regB, D, NS, s0, lane, acc);
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < SB; ++j) {
                            const int s = s0 + j;
                            if (s >= NS) break;
                            const float v = acc[j] + bias;
                            if (is_res) {
                                const GenLayer& Ln = p.layers[l + 1];
                                const float cur = regA[(size_t)s * K1 + row * k + (k - 1)];
                                p.rings[Ln.ring_off + ((size_t)(t % Ln.ring_len) * NS + s) * R + row] = v + cur;
                            } else {
                                skacc[li * NS + s] = v + skacc[li * NS + s];
                            }
                        }
                    }
                }
            }
            if (l + 1 == p.n_layers && want_head) {
                __syncthreads();
                for (int i = tid; i < nS * NS; i += GEN_NT) {
                    const int li = i / NS, s = i - li * NS;
                    p.skipbuf[(size_t)s * S + oS + li] = skacc[i];
                }
            }
            grid_barrier(p.bar, bar_target, G);
        }
        if (!want_head) continue;

        // ---- head A: y1 = relu(W1 relu(skip) + b1)
        for (int i = tid; i < NS * S; i += GEN_NT) regA[i] = fmaxf(__ldcg(p.skipbuf + i), 0.f);
        __syncthreads();
        for (int it = warp; it < nE; it += GEN_WARPS) {
            const int row = oE + it;
            const float bias = __ldg(p.e1b + row);
            for (int s0 = 0; s0 < NS; s0 += SB) {
                float acc[SB];
                row_dot<SB>(p.e1w + (size_t)row * S, regA, S, NS, s0, lane, acc);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < SB; ++j)
                        if (s0 + j < NS) p.y1buf[(size_t)(s0 + j) * E + row] = fmaxf(acc[j] + bias, 0.f);
                }
            }
        }
        grid_barrier(p.bar, bar_target, G);
        // ---- head B: logits = W2 y1 + b2, minus the regularizer (wavenet_model.py:273-274,280)
        for (int i = tid; i < NS * E; i += GEN_NT) regB[i] = __ldcg(p.y1buf + i);
        __syncthreads();
        for (int it = warp; it < nC; it += GEN_WARPS) {
            const int row = oC + it;
            const float bias = __ldg(p.e2b + row);
            const float dc = (float)row - (float)C / 2.f;
            const float reg = (dc * dc) * p.regularize;
            for (int s0 = 0; s0 < NS; s0 += SB) {
                float acc[SB];
                row_dot<SB>(p.e2w + (size_t)row * E, regB, E, NS, s0, lane, acc);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < SB; ++j) {
                        const int s = s0 + j;
                        if (s >= NS) break;
                        const float v = (acc[j] + bias) - reg;
                        p.logitbuf[(size_t)s * C + row] = v;
                        if (p.out_logits) p.out_logits[((size_t)s * p.n_samples + samp) * C + row] = v;
                    }
                }
            }
        }
        grid_barrier(p.bar, bar_target, G);
        // ---- choose (every CTA redundantly, so no broadcast barrier is needed): warp per stream
        float* pw = prob + warp * C;
        for (int s = warp; s < NS; s += GEN_WARPS) {
            const float* lg = p.logitbuf + (size_t)s * C;
            int choice;
            if (p.temperature > 0.f) {
                float m = -INFINITY;
                for (int c = lane; c < C; c += 32) {
                    const float x = __ldcg(lg + c) / p.temperature;
                    pw[c] = x;
                    m = fmaxf(m, x);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float sum = 0.f;
                for (int c = lane; c < C; c += 32) {
                    const float e = expf(pw[c] - m);
                    pw[c] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                __syncwarp();
                choice = 0;
                if (lane == 0) {
                    // numpy.random.choice: float64 cumulative sum, normalised by its last element, searchsorted 'right'
                    double total = 0.0;
                    for (int c = 0; c < C; ++c) total += (double)(pw[c] / sum);
                    const double u = p.uniforms[(size_t)s * p.n_samples + samp];
                    double run = 0.0;
                    int cnt = 0;
                    for (int c = 0; c < C; ++c) {
                        run += (double)(pw[c] / sum);
                        cnt += ((run / total) <= u) ? 1 : 0;
                    }
                    choice = cnt < C ? cnt : C - 1;
                }
                choice = __shfl_sync(0xffffffffu, choice, 0);
            } else {
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int c = lane; c < C; c += 32) {
                    const float x = __ldcg(lg + c);
                    if (x > best || (x == best && c < bi)) { best = x; bi = c; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                choice = bi == 0x7fffffff ? 0 : bi;
            }
            if (lane == 0) {
                idx_s[s] = choice;
                if (cta == 0) p.out_idx[(size_t)s * p.n_samples + samp] = choice;
            }
        }
        __syncthreads();
    }
    if (cta == 0)
        for (int s = tid; s < NS; s += GEN_NT) p.cur_idx[s] = idx_s[s];
}

// ================================================================================================ LL kernel
// Same decomposition as gen_kernel, but no grid barriers: every exchanged float travels as an 8-byte {value, tag}
// pair written with one volatile store; consumers spin on the pair itself until the tag they expect appears
// (the NCCL "LL" idea).  Tags: ring slot of time tau carries tau+1; per-evaluation vectors carry t+1 and are
// double-buffered by the parity of t.  A location is rewritten two evaluations (or one full ring period) after it
// was last read, and a writer can only get there once every CTA has finished the evaluation in between -- each CTA
// owns at least one dilation channel (G <= D) whose z every residual row needs -- so no reader is ever overtaken.
// The weight rows a CTA needs for the next stages are prefetched into shared memory with cp.async.bulk (TMA) while
// it waits for data; they are the same rows every evaluation, in the parameters' own (state_dict) layout.
constexpr long long GEN_TIMEOUT_CYCLES = 6000000000LL;     // ~3 s: a missing tag aborts the launch instead of hanging

// one naturally aligned 64-bit word = single-copy atomic; gpu scope keeps the traffic at the L2
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return make_uint2((unsigned)(w & 0xffffffffull), (unsigned)(w >> 32));
}
__device__ __forceinline__ void st_pair(uint2* p, float val, unsigned tag) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(val);
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ float poll_pair(const uint2* p, unsigned tag, int* err, int* abort_s) {
    uint2 v = ld_pair(p);
    if (v.y != tag) {
        const long long t0 = clock64();
        do {
            __nanosleep(40);                      // keep the polling storm off the L2 slice the producer writes to
            v = ld_pair(p);
            if (clock64() - t0 > GEN_TIMEOUT_CYCLES || *reinterpret_cast<volatile int*>(err) != 0) {
                *reinterpret_cast<volatile int*>(err) = 1;
                *reinterpret_cast<volatile int*>(abort_s) = 1;
                break;
            }
        } while (v.y != tag);
    }
    return __uint_as_float(v.x);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// partial dot products over K-range [k0, k1) (multiples of 4 when K%4==0) -- weights from shared or global memory
template <int SB, bool W_SMEM>
__device__ __forceinline__ void row_dot_part(const float* __restrict__ w, const float* __restrict__ xs, int K, int k0,
                                             int k1, int NS, int s0, int lane, float (&acc)[SB]) {
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = 0.f;
    if ((K & 3) == 0) {
        const float4* w4p = reinterpret_cast<const float4*>(w);
        for (int i4 = (k0 >> 2) + lane; i4 < (k1 >> 2); i4 += 32) {
            float4 w4;
            if constexpr (W_SMEM) w4 = w4p[i4];
            else w4 = __ldg(w4p + i4);
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                if (s0 + j < NS) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xs + (size_t)(s0 + j) * K + 4 * i4);
                    float a = acc[j];
                    a = fmaf(w4.x, x4.x, a); a = fmaf(w4.y, x4.y, a); a = fmaf(w4.z, x4.z, a); a = fmaf(w4.w, x4.w, a);
                    acc[j] = a;
                }
            }
        }
    } else {
        for (int i = k0 + lane; i < k1; i += 32) {
            float wv;
            if constexpr (W_SMEM) wv = w[i];
            else wv = __ldg(w + i);
#pragma unroll
            for (int j = 0; j < SB; ++j)
                if (s0 + j < NS) acc[j] = fmaf(wv, xs[(size_t)(s0 + j) * K + i], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = warp_sum(acc[j]);
}

// Rows of one stage for one CTA.  Stage numbering inside an evaluation: 2l = conv rows of layer l, 2l+1 = 1x1 rows
// of layer l, 2NL = end_conv_1 rows, 2NL+1 = end_conv_2 rows.
struct StageDesc {
    int n, K;                 // rows, row length
    int n_first;              // stage 2: residual rows come first (n_first of them), then skip rows
    int n_nom;                // nominal row count of the stage (fixes the K split, whatever rows are active)
};
__device__ __forceinline__ StageDesc stage_desc(const GenParams& p, int st, bool want_head, int nD, int nR, int nS,
                                                int nE, int nC) {
    StageDesc d;
    const int NL = p.n_layers;
    if (st < 2 * NL) {
        const int l = st >> 1;
        if ((st & 1) == 0) { d.n = 2 * nD; d.K = p.k * p.R; d.n_first = d.n; d.n_nom = d.n; }
        else { d.n_first = (l + 1 < NL) ? nR : 0; d.n = d.n_first + (want_head ? nS : 0); d.K = p.D; d.n_nom = nR + nS; }
    } else if (st == 2 * NL) { d.n = nE; d.K = p.S; d.n_first = d.n; d.n_nom = d.n; }
    else { d.n = nC; d.K = p.E; d.n_first = d.n; d.n_nom = d.n; }
    return d;
}
__device__ __forceinline__ const float* stage_row(const GenParams& p, const GenLayer* layers, int st, const StageDesc& d,
                                                  int i, int cta, int G) {
    const int NL = p.n_layers;
    if (st < 2 * NL) {
        const GenLayer& L = layers[st >> 1];
        if ((st & 1) == 0) return ((i & 1) ? L.wg : L.wf) + (size_t)(cta * own_per(p.D, G) + (i >> 1)) * d.K;
        if (i < d.n_first) return L.wr + (size_t)(cta * own_per(p.R, G) + i) * d.K;
        return L.ws + (size_t)(cta * own_per(p.S, G) + (i - d.n_first)) * d.K;
    }
    if (st == 2 * NL) return p.e1w + (size_t)(cta * own_per(p.E, G) + i) * d.K;
    return p.e2w + (size_t)(cta * own_per(p.C, G) + i) * d.K;
}

template <int SB, bool PREFETCH>
__global__ void __launch_bounds__(GEN_NT, 1) gen_kernel_ll(const GenParams p) {
    extern __shared__ __align__(16) float sm[];
    float* regA = sm;                          // stage-1 inputs [NS][k*R] / head input [NS][S] / logits [NS][C]
    float* regB = regA + p.regA;               // z [NS][D] / y1 [NS][E]
    float* part = regB + p.regB;               // partial sums [item][kpart][NS]
    float* skacc = part + p.part_n;            // running skip sums of the rows this CTA owns [rows][NS]
    float* prob = skacc + p.skacc_n;           // [GEN_WARPS][C] softmax scratch
    double* cdf = reinterpret_cast<double*>(prob + GEN_WARPS * p.C);   // [GEN_WARPS][C]
    float* wbuf = reinterpret_cast<float*>(cdf + GEN_WARPS * p.C);     // [n_wslots][wslot_floats]
    unsigned long long* fullb = reinterpret_cast<unsigned long long*>(wbuf + (size_t)p.n_wslots * p.wslot_floats);
    GenLayer* lay_s = reinterpret_cast<GenLayer*>(fullb + 8);           // per-layer table, copied from global once
    int* idx_s = reinterpret_cast<int*>(lay_s + p.n_layers);            // current input index per stream [NS]
    int* abort_s = idx_s + p.NS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int NS = p.NS, R = p.R, D = p.D, S = p.S, E = p.E, C = p.C, k = p.k, NL = p.n_layers;
    const int K1 = k * R;
    const int nD = own_cnt(D, G, cta), oD = cta * own_per(D, G);
    const int nR = own_cnt(R, G, cta), oR = cta * own_per(R, G);
    const int nS = own_cnt(S, G, cta), oS = cta * own_per(S, G);
    const int nE = own_cnt(E, G, cta), oE = cta * own_per(E, G);
    const int nC = own_cnt(C, G, cta), oC = cta * own_per(C, G);
    const int NSLOT = p.n_wslots;

    for (int s = tid; s < NS; s += GEN_NT) idx_s[s] = p.cur_idx[s];
    {
        const int* src = reinterpret_cast<const int*>(p.layers);
        int* dst = reinterpret_cast<int*>(lay_s);
        for (int i = tid; i < NL * (int)(sizeof(GenLayer) / sizeof(int)); i += GEN_NT) dst[i] = src[i];
    }
    if (tid == 0) {
        *abort_s = 0;
        if (PREFETCH)
            for (int i = 0; i < NSLOT; ++i) mbar_init(fullb + i, 1);
    }
    if (PREFETCH) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // ---- weight prefetch pipeline (thread 0 produces): stage q of the launch lives in slot q % NSLOT
    int pf_ev = 0, pf_st = 0;                 // producer cursor
    long long pf_q = 0, cons_q = 0;
    auto stages_in_eval = [&](int ev) { return (p.t0 + ev >= p.n_given - 1) ? 2 * NL + 2 : 2 * NL; };
    auto produce_one = [&]() {                // the last thread only (it has no epilogue work)
        if (pf_ev >= p.n_evals) return;
        const bool wh = (p.t0 + pf_ev >= p.n_given - 1);
        const StageDesc d = stage_desc(p, pf_st, wh, nD, nR, nS, nE, nC);
        const int slot = (int)(pf_q % NSLOT);
        mbar_expect_tx(fullb + slot, (unsigned)(d.n * d.K * 4));
        float* dst = wbuf + (size_t)slot * p.wslot_floats;
        for (int i = 0; i < d.n; ++i) bulk_g2s(dst + (size_t)i * d.K, stage_row(p, lay_s, pf_st, d, i, cta, G), d.K * 4, fullb + slot);
        ++pf_q;
        if (++pf_st >= stages_in_eval(pf_ev)) { pf_st = 0; ++pf_ev; }
    };
    if (PREFETCH && tid == GEN_NT - 1)
        for (int i = 0; i < NSLOT; ++i) produce_one();

    // One dot-product stage: rows of `st` times the vectors xs[s][0..K) -> part[item][kpart][s]; all 8 warps work,
    // rows are split over nw = 8/items warps along K when there are fewer rows than warps.
    auto run_stage = [&](int st, const StageDesc& d, const float* xs, int& nw_out) {
        int nw = 1;
        while (nw * 2 * d.n_nom <= GEN_WARPS && (d.K / (nw * 2)) % 4 == 0 && d.K / (nw * 2) >= 32) nw *= 2;
        if (d.n_nom == 0) nw = 1;
        nw_out = nw;
        const float* wslot = nullptr;
        if (PREFETCH) {
            const int slot = (int)(cons_q % NSLOT);
            mbar_wait(fullb + slot, (unsigned)((cons_q / NSLOT) & 1));
            wslot = wbuf + (size_t)slot * p.wslot_floats;
        }
        const int kchunk = d.K / nw;
        for (int wi = warp; wi < d.n * nw; wi += GEN_WARPS) {
            const int it = wi / nw, kp = wi - it * nw;
            const int k0 = kp * kchunk, k1 = (kp == nw - 1) ? d.K : k0 + kchunk;
            for (int s0 = 0; s0 < NS; s0 += SB) {
                float acc[SB];
                if (PREFETCH) row_dot_part<SB, true>(wslot + (size_t)it * d.K, xs, d.K, k0, k1, NS, s0, lane, acc);
                else row_dot_part<SB, false>(stage_row(p, lay_s, st, d, it, cta, G), xs, d.K, k0, k1, NS, s0, lane, acc);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < SB; ++j)
                        if (s0 + j < NS) part[(size_t)(it * nw + kp) * NS + s0 + j] = acc[j];
                }
            }
        }
        ++cons_q;
    };
    auto sum_parts = [&](int it, int nw, int s) {
        float v = part[(size_t)(it * nw) * NS + s];
        for (int kp = 1; kp < nw; ++kp) v += part[(size_t)(it * nw + kp) * NS + s];
        return v;
    };

    for (int ev = 0; ev < p.n_evals; ++ev) {
        const int t = p.t0 + ev;
        const unsigned tag = (unsigned)t + 1u;
        const int par = t & 1;
        const bool want_head = (t >= p.n_given - 1);
        const int samp = t - (p.n_given - 1);
        if (t < p.n_given) {
            for (int s = tid; s < NS; s += GEN_NT) idx_s[s] = p.first[(size_t)s * p.n_given + t];
        } else if (p.forced != nullptr) {
            for (int s = tid; s < NS; s += GEN_NT) idx_s[s] = p.forced[(size_t)s * p.n_samples + (t - p.n_given)];
        }
        for (int i = tid; i < nS * NS; i += GEN_NT) skacc[i] = 0.f;
        __syncthreads();
        if (*abort_s) return;

        for (int l = 0; l < NL; ++l) {
            const GenLayer& L = lay_s[l];
            uint2* ring = p.ringLL + L.ring_off;
            const int slot_t = t % L.ring_len;
            // ---- stage-1 inputs: regA[s][r*k + j] = tap j of channel r (tap k-1 = the value just enqueued)
            for (int i = tid; i < NS * K1; i += GEN_NT) {
                const int s = i / K1, rem = i - s * K1, r = rem / k, j = rem - r * k;
                float v;
                if (j == k - 1 && l == 0) {
                    int c = idx_s[s];
                    c = c < 0 ? 0 : (c >= C ? C - 1 : c);
                    v = __ldg(p.start_w + (size_t)r * C + c) + (p.start_b ? __ldg(p.start_b + r) : 0.f);
                    if (r >= oR && r < oR + nR) st_pair(ring + ((size_t)slot_t * NS + s) * R + r, v, tag);       // enqueue
                } else {
                    const int tt = t - (k - 1 - j) * L.dil;
                    if (tt < 0) v = 0.f;                                                             // zero history
                    else v = poll_pair(ring + ((size_t)(tt % L.ring_len) * NS + s) * R + r, (unsigned)tt + 1u, p.err, abort_s);
                }
                regA[i] = v;
            }
            __syncthreads();
            if (*abort_s) return;
            // ---- stage 1: filter / gate rows -> z
            int nw;
            StageDesc d1 = stage_desc(p, 2 * l, want_head, nD, nR, nS, nE, nC);
            run_stage(2 * l, d1, regA, nw);
            __syncthreads();
            if (PREFETCH && tid == GEN_NT - 1) produce_one();
            uint2* zl = p.zLL + ((size_t)(par * NL + l) * NS) * D;
            for (int i = tid; i < nD * NS; i += GEN_NT) {
                const int ci = i / NS, s = i - ci * NS, c = oD + ci;
                const float f = sum_parts(2 * ci, nw, s) + (L.bf ? __ldg(L.bf + c) : 0.f);
                const float g = sum_parts(2 * ci + 1, nw, s) + (L.bg ? __ldg(L.bg + c) : 0.f);
                st_pair(zl + (size_t)s * D + c, tanh_(f) * sigmoid_(g), tag);
            }
            // ---- stage 2: residual rows (-> next layer's ring slot t) and skip rows (-> running sums)
            StageDesc d2 = stage_desc(p, 2 * l + 1, want_head, nD, nR, nS, nE, nC);
            if (d2.n > 0) {
                for (int i = tid; i < NS * D; i += GEN_NT) regB[i] = poll_pair(zl + i, tag, p.err, abort_s);
            }
            __syncthreads();                       // also protects `part` (read above) against the next run_stage
            if (*abort_s) return;
            run_stage(2 * l + 1, d2, regB, nw);
            __syncthreads();
            if (PREFETCH && tid == GEN_NT - 1) produce_one();
            for (int i = tid; i < d2.n * NS; i += GEN_NT) {
                const int it = i / NS, s = i - it * NS;
                if (it < d2.n_first) {
                    const int row = oR + it;
                    const GenLayer& Ln = lay_s[l + 1];
                    const float v = sum_parts(it, nw, s) + (L.br ? __ldg(L.br + row) : 0.f);
                    const float cur = regA[(size_t)s * K1 + row * k + (k - 1)];
                    st_pair(p.ringLL + Ln.ring_off + ((size_t)(t % Ln.ring_len) * NS + s) * R + row, v + cur, tag);
                } else {
                    const int li = it - d2.n_first, row = oS + li;
                    const float v = sum_parts(it, nw, s) + (L.bs ? __ldg(L.bs + row) : 0.f);
                    skacc[li * NS + s] = v + skacc[li * NS + s];
                }
            }
            __syncthreads();                       // the epilogue read regA (cur) and part: done before the next gather
        }
        if (!want_head) continue;

        // ---- head: skip -> end_conv_1 -> end_conv_2 -> choose
        __syncthreads();
        uint2* skl = p.skipLL + (size_t)par * NS * S;
        for (int i = tid; i < nS * NS; i += GEN_NT) {
            const int li = i / NS, s = i - li * NS;
            st_pair(skl + (size_t)s * S + oS + li, skacc[i], tag);
        }
        int nw;
        StageDesc dA = stage_desc(p, 2 * NL, true, nD, nR, nS, nE, nC);
        if (dA.n > 0)
            for (int i = tid; i < NS * S; i += GEN_NT) regA[i] = fmaxf(poll_pair(skl + i, tag, p.err, abort_s), 0.f);
        __syncthreads();
        if (*abort_s) return;
        run_stage(2 * NL, dA, regA, nw);
        __syncthreads();
        if (PREFETCH && tid == GEN_NT - 1) produce_one();
        uint2* yl = p.y1LL + (size_t)par * NS * E;
        for (int i = tid; i < nE * NS; i += GEN_NT) {
            const int it = i / NS, s = i - it * NS, row = oE + it;
            st_pair(yl + (size_t)s * E + row, fmaxf(sum_parts(it, nw, s) + __ldg(p.e1b + row), 0.f), tag);
        }
        StageDesc dB = stage_desc(p, 2 * NL + 1, true, nD, nR, nS, nE, nC);
        if (dB.n > 0)
            for (int i = tid; i < NS * E; i += GEN_NT) regB[i] = poll_pair(yl + i, tag, p.err, abort_s);
        __syncthreads();
        if (*abort_s) return;
        run_stage(2 * NL + 1, dB, regB, nw);
        __syncthreads();
        if (PREFETCH && tid == GEN_NT - 1) produce_one();
        uint2* lgl = p.logitLL + (size_t)par * NS * C;
        for (int i = tid; i < nC * NS; i += GEN_NT) {
            const int it = i / NS, s = i - it * NS, row = oC + it;
            const float dc = (float)row - (float)C / 2.f;
            const float v = (sum_parts(it, nw, s) + __ldg(p.e2b + row)) - (dc * dc) * p.regularize;
            st_pair(lgl + (size_t)s * C + row, v, tag);
            if (p.out_logits) p.out_logits[((size_t)s * p.n_samples + samp) * C + row] = v;
        }
        // ---- every CTA collects all logits and picks the next sample itself (no broadcast needed)
        for (int i = tid; i < NS * C; i += GEN_NT) regA[i] = poll_pair(lgl + i, tag, p.err, abort_s);
        __syncthreads();
        if (*abort_s) return;
        float* pw = prob + warp * C;
        double* cw = cdf + warp * C;
        for (int s = warp; s < NS; s += GEN_WARPS) {
            const float* lg = regA + (size_t)s * C;
            int choice;
            if (p.temperature > 0.f) {
                float m = -INFINITY;
                for (int c = lane; c < C; c += 32) {
                    const float x = lg[c] / p.temperature;
                    pw[c] = x;
                    m = fmaxf(m, x);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float sum = 0.f;
                for (int c = lane; c < C; c += 32) {
                    const float e = expf(pw[c] - m);
                    pw[c] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                for (int c = lane; c < C; c += 32) cw[c] = (double)(pw[c] / sum);
                __syncwarp();
                // numpy.random.choice: sequential float64 cumulative sum, normalised by its last element,
                // searchsorted(side='right') == number of normalised entries <= u
                if (lane == 0) {
                    double run = 0.0;
                    for (int c = 0; c < C; ++c) { run += cw[c]; cw[c] = run; }
                }
                __syncwarp();
                const double total = cw[C - 1];
                const double u = p.uniforms[(size_t)s * p.n_samples + samp];
                int cnt = 0;
                for (int c = lane; c < C; c += 32) cnt += ((cw[c] / total) <= u) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                choice = cnt < C ? cnt : C - 1;
            } else {
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int c = lane; c < C; c += 32) {
                    const float x = lg[c];
                    if (x > best || (x == best && c < bi)) { best = x; bi = c; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                choice = bi == 0x7fffffff ? 0 : bi;
            }
            if (lane == 0) {
                idx_s[s] = choice;
                if (cta == 0) p.out_idx[(size_t)s * p.n_samples + samp] = choice;
            }
        }
        __syncthreads();
    }
    if (cta == 0)
        for (int s = tid; s < NS; s += GEN_NT) p.cur_idx[s] = idx_s[s];
}

// ================================================================================================ fast kernel
// Single stream, k = 2, power-of-two grid, every stage's rows a divisor of 8: the latency-critical special case
// (cfg 2).  Same exchange protocol, same tags, same summation order as gen_kernel_ll, but the per-stage critical
// path is stripped down: every warp owns one (row, K-part) pair for the whole launch, polls exactly the {value,tag}
// pairs it multiplies straight into registers (no staging, no index arithmetic with divisions), and there is ONE
// __syncthreads per stage (partial sums are double buffered).  Ring positions advance incrementally.
struct Pair2 { uint2 a, b; };
__device__ __forceinline__ Pair2 ld_pair2(const uint2* p) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    Pair2 r;
    r.a = make_uint2((unsigned)(w0 & 0xffffffffull), (unsigned)(w0 >> 32));
    r.b = make_uint2((unsigned)(w1 & 0xffffffffull), (unsigned)(w1 >> 32));
    return r;
}
// two consecutive pairs carrying `tag`; spins (bounded) until both are there
// slow path of poll2, out of line: spinning code (with its timeout) must not bloat the per-layer instruction stream
__device__ __noinline__ Pair2 poll2_spin(const uint2* p, unsigned tag, int* err, int* abort_s) {
    Pair2 q;
    const long long t0 = clock64();
    unsigned spins = 0;
    do {
        q = ld_pair2(p);
        if ((++spins & 255u) == 0 &&
            (clock64() - t0 > GEN_TIMEOUT_CYCLES || *reinterpret_cast<volatile int*>(err) != 0)) {
            *reinterpret_cast<volatile int*>(err) = 1;
            *reinterpret_cast<volatile int*>(abort_s) = 1;
            break;
        }
    } while (q.a.y != tag || q.b.y != tag);
    return q;
}
__device__ __forceinline__ void poll2(const uint2* p, unsigned tag, float& v0, float& v1, int* err, int* abort_s) {
    Pair2 q = ld_pair2(p);
    if (q.a.y != tag || q.b.y != tag) q = poll2_spin(p, tag, err, abort_s);
    v0 = __uint_as_float(q.a.x);
    v1 = __uint_as_float(q.b.x);
}

constexpr int FAST_MAXI = 4;      // float4 iterations per lane per row part (K part <= 512)

// All the pairs one lane multiplies in a stage, fetched with every load in flight at once (one L2 round trip when the
// data is already there) and re-fetched as a batch until all tags match.  Iteration `it` covers pairs
// [first + it*128, +4) of `base` (4 consecutive values = one float4 of the weight row).
__device__ __forceinline__ void poll_quads(const uint2* base, int first, int n_iter, unsigned tag, float (&v)[FAST_MAXI][4],
                                           int* err, int* abort_s) {
    Pair2 a[FAST_MAXI], b[FAST_MAXI];
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
#pragma unroll
        for (int it = 0; it < FAST_MAXI; ++it)
            if (it < n_iter) {
                a[it] = ld_pair2(base + first + it * 128);
                b[it] = ld_pair2(base + first + it * 128 + 2);
            }
        bool ok = true;
#pragma unroll
        for (int it = 0; it < FAST_MAXI; ++it)
            if (it < n_iter) ok = ok && a[it].a.y == tag && a[it].b.y == tag && b[it].a.y == tag && b[it].b.y == tag;
        if (ok) break;
        if ((++spins & 255u) == 0 &&
            (clock64() - t0 > GEN_TIMEOUT_CYCLES || *reinterpret_cast<volatile int*>(err) != 0)) {
            *reinterpret_cast<volatile int*>(err) = 1;
            *reinterpret_cast<volatile int*>(abort_s) = 1;
            break;
        }
    }
#pragma unroll
    for (int it = 0; it < FAST_MAXI; ++it)
        if (it < n_iter) {
            v[it][0] = __uint_as_float(a[it].a.x); v[it][1] = __uint_as_float(a[it].b.x);
            v[it][2] = __uint_as_float(b[it].a.x); v[it][3] = __uint_as_float(b[it].b.x);
        }
}
// Same for the 2-tap conv input: iteration `it` needs channels (r0, r0+1), r0 = first_r + it*64, from two ring slots
// (time t-d with tag_old unless the history is still empty, time t with tag_cur unless `cur` comes from elsewhere).
__device__ __forceinline__ void poll_taps(const uint2* old_slot, unsigned tag_old, bool have_old, const uint2* cur_slot,
                                          unsigned tag_cur, bool have_cur, int first_r, int n_iter,
                                          float (&o)[FAST_MAXI][2], float (&c)[FAST_MAXI][2], int* err, int* abort_s) {
    Pair2 a[FAST_MAXI], b[FAST_MAXI];
    const long long t0 = clock64();
    for (;;) {
#pragma unroll
        for (int it = 0; it < FAST_MAXI; ++it)
            if (it < n_iter) {
                if (have_old) a[it] = ld_pair2(old_slot + first_r + it * 64);
                if (have_cur) b[it] = ld_pair2(cur_slot + first_r + it * 64);
            }
        bool ok = true;
#pragma unroll
        for (int it = 0; it < FAST_MAXI; ++it)
            if (it < n_iter) {
                if (have_old) ok = ok && a[it].a.y == tag_old && a[it].b.y == tag_old;
                if (have_cur) ok = ok && b[it].a.y == tag_cur && b[it].b.y == tag_cur;
            }
        if (ok) break;
        if (clock64() - t0 > GEN_TIMEOUT_CYCLES || *reinterpret_cast<volatile int*>(err) != 0) {
            *reinterpret_cast<volatile int*>(err) = 1;
            *reinterpret_cast<volatile int*>(abort_s) = 1;
            break;
        }
    }
#pragma unroll
    for (int it = 0; it < FAST_MAXI; ++it)
        if (it < n_iter) {
            o[it][0] = have_old ? __uint_as_float(a[it].a.x) : 0.f;
            o[it][1] = have_old ? __uint_as_float(a[it].b.x) : 0.f;
            if (have_cur) { c[it][0] = __uint_as_float(b[it].a.x); c[it][1] = __uint_as_float(b[it].b.x); }
        }
}

#define WORKER_SYNC() asm volatile("bar.sync 1, 256;" ::: "memory")
__device__ __forceinline__ void mbar_arrive_(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}

// argmax (first index wins ties) or numpy.random.choice's inverse-CDF draw, by one warp over C logits in shared memory.
// Kept out of line so that its fp64 code does not sit in the instruction stream of the per-layer loop.
__device__ __noinline__ int choose_sample(float* logit_s, double* cdf, int C, int lane, float temperature, const double* u_ptr) {
    int choice;
    if (temperature > 0.f) {
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) {
            const float x = logit_s[c] / temperature;
            logit_s[c] = x;
            m = fmaxf(m, x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float e = expf(logit_s[c] - m);
            logit_s[c] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        // numpy.random.choice: float64 cumulative sum of the float32 probabilities, normalised by its last element,
        // searchsorted(side='right').  The running sum is taken per lane over a contiguous chunk plus a warp scan
        // (equal to the sequential sum up to float64 rounding, i.e. ~1e-16 relative on the CDF).
        const int per = (C + 31) / 32;
        const int c_lo = lane * per, c_hi = min(C, c_lo + per);
        double run = 0.0;
        for (int c = c_lo; c < c_hi; ++c) { run += (double)(logit_s[c] / sum); cdf[c] = run; }
        double incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        const double total = __shfl_sync(0xffffffffu, incl, 31);
        const double u = *u_ptr;
        const double ut = u * total;
        int cnt = 0;
        for (int c = c_lo; c < c_hi; ++c) {
            const double v = cdf[c] + excl;
            bool le = v <= ut;
            if (fabs(v - ut) <= 1e-9 * total) le = (v / total) <= u;       // exact rule only where it can matter
            cnt += le ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        choice = cnt < C ? cnt : C - 1;
    } else {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float x = logit_s[c];
            if (x > best || (x == best && c < bi)) { best = x; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        choice = bi == 0x7fffffff ? 0 : bi;
    }
    return choice;
}

template <bool PREFETCH, bool TRACE>
__global__ void __launch_bounds__(GEN_NT + 32, 1) gen_kernel_fast(const GenParams p) {
    extern __shared__ __align__(16) float sm[];
    float* part = sm;                                   // [2][GEN_WARPS] partial sums, double buffered by stage parity
    float* skacc = part + 2 * GEN_WARPS;                // [nS]
    float* cur_own = skacc + ((p.S / (int)gridDim.x + 3) & ~3);      // [2][nR] layer input at the residual rows this CTA owns
    float* xin = cur_own + 2 * ((p.R / (int)gridDim.x + 3) & ~3);    // [2][XN] the polled input vector of a stage
    const int XN = p.regA;                              // max(R, D, S, E, C) rounded to 4 (set by the host)
    double* cdf = reinterpret_cast<double*>(xin + 2 * XN);                     // [C]
    float* wbuf = reinterpret_cast<float*>(cdf + p.C);                         // [n_wslots][wslot_floats]
    unsigned long long* fullb = reinterpret_cast<unsigned long long*>(wbuf + (size_t)p.n_wslots * p.wslot_floats);
    unsigned long long* emptyb = fullb + 4;             // consumers -> producer: slot may be refilled (8 warps arrive)
    GenLayer* lay_s = reinterpret_cast<GenLayer*>(fullb + 8);
    uint2* old_s = reinterpret_cast<uint2*>(lay_s + p.n_layers);               // [2][R] prefetched old taps (see below)
    int* slot_s = reinterpret_cast<int*>(old_s + 2 * p.R);                     // [NL] ring slot of time t per layer
    int* misc = slot_s + p.n_layers;                                           // [0] current index, [1] abort

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int R = p.R, D = p.D, S = p.S, E = p.E, C = p.C, NL = p.n_layers;
    const int K1 = 2 * R;
    const int nD = D / G, nR = R / G, nS = S / G, nE = E / G, nC = C / G;
    const int oD = cta * nD, oR = cta * nR, oS = cta * nS, oE = cta * nE, oC = cta * nC;
    const int NSLOT = p.n_wslots;
    // fixed warp -> (row, K-part) assignment per stage kind, and this lane's float4 range inside the part
    const int HS1 = GEN_WARPS / (2 * nD), HS2 = GEN_WARPS / (nR + nS), HSA = GEN_WARPS / nE, HSB = GEN_WARPS / nC;
    const int row1 = warp / HS1, g1 = ((warp - row1 * HS1) * (K1 / HS1) >> 2) + lane, n1 = (((K1 / HS1) >> 2) - lane + 31) / 32;
    const int row2 = warp / HS2, g2 = ((warp - row2 * HS2) * (D / HS2) >> 2) + lane, n2 = (((D / HS2) >> 2) - lane + 31) / 32;
    const int rowA = warp / HSA, gA = ((warp - rowA * HSA) * (S / HSA) >> 2) + lane, nA = (((S / HSA) >> 2) - lane + 31) / 32;
    const int rowB = warp / HSB, gB = ((warp - rowB * HSB) * (E / HSB) >> 2) + lane, nB = (((E / HSB) >> 2) - lane + 31) / 32;

    {
        const int* src = reinterpret_cast<const int*>(p.layers);
        int* dst = reinterpret_cast<int*>(lay_s);
        for (int i = tid; i < NL * (int)(sizeof(GenLayer) / sizeof(int)); i += GEN_NT + 32) dst[i] = src[i];
    }
    if (tid == 0) {
        misc[0] = p.cur_idx[0];
        misc[1] = 0;
        if (PREFETCH)
            for (int i = 0; i < NSLOT; ++i) { mbar_init(fullb + i, 1); mbar_init(emptyb + i, GEN_WARPS); }
    }
    if (PREFETCH) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    for (int l = tid; l < NL; l += GEN_NT) {          // slot of time t0-1, advanced at the top of every evaluation
        const int len = lay_s[l].ring_len;
        slot_s[l] = (p.t0 + len - 1) % len;
    }
    // ---- weight prefetch: a dedicated producer warp (warp 8) runs ahead of the 8 worker warps through the same stage
    // sequence as gen_kernel_ll (all rows of a stage are always fetched); full[slot] = bytes landed, empty[slot] = all
    // worker warps are done with the slot.  Keeping the producer off the worker warps matters: every stage needs every
    // worker warp's row, so anything a worker lane does besides its row is on the critical path of the whole GPU.
    int* abort_s = misc + 1;
    const unsigned smask = (unsigned)NSLOT - 1u, sshift = (NSLOT == 4) ? 2u : 1u;
    if (warp == GEN_WARPS) {
        if (PREFETCH && lane == 0) {
            unsigned q = 0;
            for (int ev = 0; ev < p.n_evals; ++ev) {
                const bool wh = (p.t0 + ev >= p.n_given - 1);
                const int n_st = wh ? 2 * NL + 2 : 2 * NL;
                for (int st = 0; st < n_st; ++st, ++q) {
                    StageDesc d = stage_desc(p, st, true, nD, nR, nS, nE, nC);
                    if (st < 2 * NL && (st & 1)) { d.n_first = nR; d.n = nR + nS; }
                    const int slot = (int)(q & smask);
                    if (q >= (unsigned)NSLOT) {                       // wait until the workers released this slot
                        const unsigned par = ((q >> sshift) & 1u) ^ 1u;
                        unsigned done = 0, spins = 0;
                        while (!done) {
                            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(done) : "r"(smem_u32(emptyb + slot)), "r"(par) : "memory");
                            if (!done && (++spins & 1023u) == 0 && *reinterpret_cast<volatile int*>(abort_s)) return;
                        }
                    }
                    mbar_expect_tx(fullb + slot, (unsigned)(d.n * d.K * 4));
                    float* dst = wbuf + (size_t)slot * p.wslot_floats;
                    for (int i = 0; i < d.n; ++i)
                        bulk_g2s(dst + (size_t)i * d.K, stage_row(p, lay_s, st, d, i, cta, G), d.K * 4, fullb + slot);
                }
            }
        }
        return;
    }
    unsigned cons_q = 0;
    // weights of (stage, row): shared-memory slot when prefetching, else the parameter tensor itself
    auto stage_weights = [&](int st, int row, int K) -> const float* {
        if (PREFETCH) {
            const int slot = (int)(cons_q & smask);
            mbar_wait(fullb + slot, (cons_q >> sshift) & 1u);
            return wbuf + (size_t)slot * p.wslot_floats + (size_t)row * K;
        }
        StageDesc d = stage_desc(p, st, true, nD, nR, nS, nE, nC);
        if (st < 2 * NL && (st & 1)) { d.n_first = nR; d.n = nR + nS; }
        return stage_row(p, lay_s, st, d, row, cta, G);
    };
    auto ldw = [&](const float* w, int i4) -> float4 {
        if (PREFETCH) return reinterpret_cast<const float4*>(w)[i4];
        return __ldg(reinterpret_cast<const float4*>(w) + i4);
    };
    auto release_slot = [&]() {                       // this warp is done reading the weight slot of the current stage
        if (PREFETCH) { __syncwarp(); if (lane == 0) mbar_arrive_(emptyb + (cons_q & smask)); }
        ++cons_q;
    };
    // One CTA-wide poll of a published vector: warps 4..7 (which never have epilogue work, so they are free the moment
    // the dot barrier opens) fetch N {value,tag} pairs ONCE per CTA -- 8x fewer L2 requests on the hot lines than every
    // warp polling its own operands, which measured as the dominant cost -- and leave the values in shared memory.
    auto poll_vector = [&](const uint2* src, int N, unsigned tg, float* dst, bool relu) {
        if (warp >= 4) {
            for (int j = tid - 128; 2 * j < N; j += 128) {
                float a, b;
                poll2(src + 2 * j, tg, a, b, p.err, abort_s);
                if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                *reinterpret_cast<float2*>(dst + 2 * j) = make_float2(a, b);
            }
        }
    };
    unsigned stage_par = 0;
    WORKER_SYNC();
    // Old taps (the ring slot of time t-d) were written >= 1 evaluation ago; they are fetched one stage ahead with
    // cp.async into shared memory: during stage 2 of layer l for layer l+1 (or for layer 0 of the next evaluation).
    auto prefetch_old = [&](int ln, int te, int slot_te) {        // slot_te = ring slot of time te in layer ln
        const GenLayer& Lp = lay_s[ln];
        if (te >= Lp.dil && tid < R / 2) {
            const int so = (slot_te + 1 == Lp.ring_len) ? 0 : slot_te + 1;
            const uint2* src = p.ringLL + Lp.ring_off + (size_t)so * R + 2 * tid;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(old_s + (ln & 1) * R + 2 * tid);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    {
        const int len0 = lay_s[0].ring_len;
        prefetch_old(0, p.t0, p.t0 % len0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    WORKER_SYNC();

    for (int ev = 0; ev < p.n_evals; ++ev) {
        const int t = p.t0 + ev;
        const unsigned tag = (unsigned)t + 1u;
        const int par = t & 1;
        const bool want_head = (t >= p.n_given - 1);
        const int samp = t - (p.n_given - 1);
        if (tid == 0) {
            if (t < p.n_given) misc[0] = p.first[t];
            else if (p.forced != nullptr) misc[0] = p.forced[t - p.n_given];
        }
        for (int l = tid; l < NL; l += GEN_NT) {
            const int s1 = slot_s[l] + 1;
            slot_s[l] = (s1 == lay_s[l].ring_len) ? 0 : s1;
        }
        for (int i = tid; i < nS; i += GEN_NT) skacc[i] = 0.f;
        WORKER_SYNC();
        if (*abort_s) return;
        int idx = misc[0];
        idx = idx < 0 ? 0 : (idx >= C ? C - 1 : idx);
        const bool tr_on = TRACE && p.trace != nullptr && cta == 0 && tid == 0 && ev == p.n_evals - 1;
        int tr_n = 0;
#define TR() do { if (TRACE) { if (tr_on && tr_n < 2040) p.trace[tr_n++] = clock64(); } } while (0)
        TR();

        for (int l = 0; l < NL; ++l) {
            const GenLayer& L = lay_s[l];
            uint2* ring = p.ringLL + L.ring_off;
            const int slot_t = slot_s[l];
            const int slot_old = (slot_t + 1 == L.ring_len) ? 0 : slot_t + 1;          // time t-d with len = d+1
            const bool have_old = (t >= L.dil);
            const unsigned tag_old = (unsigned)(t - L.dil) + 1u;
            float* cur_l = cur_own + (l & 1) * ((nR + 3) & ~3);
            float* xc = xin + stage_par * XN;
            // ================= stage 1 inputs: the layer's current input vector (R values) -> xc
            if (warp >= 4) {
                for (int j = tid - 128; 2 * j < R; j += 128) {
                    const int r0 = 2 * j;
                    float c0, c1;
                    if (l == 0) {                                  // start conv column of the current sample index
                        c0 = __ldg(p.start_w + (size_t)r0 * C + idx) + (p.start_b ? __ldg(p.start_b + r0) : 0.f);
                        c1 = __ldg(p.start_w + (size_t)(r0 + 1) * C + idx) + (p.start_b ? __ldg(p.start_b + r0 + 1) : 0.f);
                        if (r0 >= oR && r0 < oR + nR) st_pair(ring + (size_t)slot_t * R + r0, c0, tag);          // enqueue
                        if (r0 + 1 >= oR && r0 + 1 < oR + nR) st_pair(ring + (size_t)slot_t * R + r0 + 1, c1, tag);
                    } else {
                        poll2(ring + (size_t)slot_t * R + r0, tag, c0, c1, p.err, abort_s);
                    }
                    if (r0 >= oR && r0 < oR + nR) cur_l[r0 - oR] = c0;
                    if (r0 + 1 >= oR && r0 + 1 < oR + nR) cur_l[r0 + 1 - oR] = c1;
                    *reinterpret_cast<float2*>(xc + r0) = make_float2(c0, c1);
                }
            }
            const float* w1 = stage_weights(2 * l, row1, K1);
            TR();          // 1: own share of the inputs polled (pollers) / weights ready
            WORKER_SYNC();
            // ================= stage 1: filter/gate rows, K = 2R interleaved (old, cur) per channel
            {
                const uint2* os = old_s + (l & 1) * R;
                float acc = 0.f;
                bool old_ok = true;
#pragma unroll
                for (int it = 0; it < FAST_MAXI; ++it)
                    if (it < n1) {
                        const int g4 = g1 + it * 32, r0 = 2 * g4;
                        float o0 = 0.f, o1 = 0.f;
                        if (have_old) {
                            const uint4 q = *reinterpret_cast<const uint4*>(os + r0);
                            old_ok = old_ok && q.y == tag_old && q.w == tag_old;
                            o0 = __uint_as_float(q.x);
                            o1 = __uint_as_float(q.z);
                        }
                        const float2 c = *reinterpret_cast<const float2*>(xc + r0);
                        const float4 w4 = ldw(w1, g4);
                        acc = fmaf(w4.x, o0, acc); acc = fmaf(w4.y, c.x, acc); acc = fmaf(w4.z, o1, acc); acc = fmaf(w4.w, c.y, acc);
                    }
                if (!__all_sync(0xffffffffu, old_ok)) {            // prefetched copy not there yet (rare): poll the ring itself
                    acc = 0.f;
                    for (int it = 0; it < n1; ++it) {
                        const int g4 = g1 + it * 32, r0 = 2 * g4;
                        float o0, o1;
                        poll2(ring + (size_t)slot_old * R + r0, tag_old, o0, o1, p.err, abort_s);
                        const float2 c = *reinterpret_cast<const float2*>(xc + r0);
                        const float4 w4 = ldw(w1, g4);
                        acc = fmaf(w4.x, o0, acc); acc = fmaf(w4.y, c.x, acc); acc = fmaf(w4.z, o1, acc); acc = fmaf(w4.w, c.y, acc);
                    }
                }
                acc = warp_sum(acc);
                if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
                release_slot();
            }
            TR();          // 2: stage-1 dot done
            WORKER_SYNC();
            TR();          // 3: barrier passed
            if (*abort_s) return;
            uint2* zl = p.zLL + (size_t)(par * NL + l) * D;
            if (tid < nD) {
                const int c = oD + tid;
                const float* pf = part + stage_par * GEN_WARPS + (2 * tid) * HS1;
                float f = pf[0], g = pf[HS1];
                for (int q = 1; q < HS1; ++q) { f += pf[q]; g += pf[HS1 + q]; }
                f += L.bf ? __ldg(L.bf + c) : 0.f;
                g += L.bg ? __ldg(L.bg + c) : 0.f;
                st_pair(zl + c, tanh_(f) * sigmoid_(g), tag);
            }
            TR();          // 4: z published
            stage_par ^= 1;
            // ================= stage 2 inputs: z of all CTAs -> xz; old taps of the next stage 1 start flying now
            float* xz = xin + stage_par * XN;
            const bool active2 = (row2 < nR) ? (l + 1 < NL) : want_head;
            if (l + 1 < NL) prefetch_old(l + 1, t, slot_s[l + 1]);
            else if (ev + 1 < p.n_evals) prefetch_old(0, t + 1, (slot_s[0] + 1 == lay_s[0].ring_len) ? 0 : slot_s[0] + 1);
            poll_vector(zl, D, tag, xz, false);
            const float* w2 = stage_weights(2 * l + 1, row2, D);
            TR();          // 5: own share of z polled
            WORKER_SYNC();
            // ================= stage 2: residual rows (first nR) and skip rows (next nS), K = D
            {
                float acc = 0.f;
                if (active2) {
#pragma unroll
                    for (int it = 0; it < FAST_MAXI; ++it)
                        if (it < n2) {
                            const float4 z4 = *reinterpret_cast<const float4*>(xz + 4 * (g2 + it * 32));
                            const float4 w4 = ldw(w2, g2 + it * 32);
                            acc = fmaf(w4.x, z4.x, acc); acc = fmaf(w4.y, z4.y, acc); acc = fmaf(w4.z, z4.z, acc); acc = fmaf(w4.w, z4.w, acc);
                        }
                    acc = warp_sum(acc);
                }
                if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
                release_slot();
                asm volatile("cp.async.wait_group 0;" ::: "memory");      // the old taps for the next stage 1 have landed
            }
            TR();          // 6: stage-2 dot done
            WORKER_SYNC();
            TR();          // 7: barrier passed
            if (*abort_s) return;
            if (tid < nR + nS) {
                const float* ps = part + stage_par * GEN_WARPS + tid * HS2;
                float v = ps[0];
                for (int q = 1; q < HS2; ++q) v += ps[q];
                if (tid < nR) {
                    if (l + 1 < NL) {
                        const int row = oR + tid;
                        const GenLayer& Ln = lay_s[l + 1];
                        v += L.br ? __ldg(L.br + row) : 0.f;
                        st_pair(p.ringLL + Ln.ring_off + (size_t)slot_s[l + 1] * R + row, v + cur_l[tid], tag);
                    }
                } else if (want_head) {
                    const int li = tid - nR, row = oS + li;
                    v += L.bs ? __ldg(L.bs + row) : 0.f;
                    skacc[li] = v + skacc[li];
                }
            }
            TR();          // 8: h' published
            stage_par ^= 1;
            // no barrier here: part, xin and cur_own are double buffered, and their next writers sit behind a barrier
            // that this epilogue's threads must reach first
        }
        if (!want_head) continue;

        // ================= head
        uint2* skl = p.skipLL + (size_t)par * S;
        if (tid >= nR && tid < nR + nS) st_pair(skl + oS + (tid - nR), skacc[tid - nR], tag);   // same thread that summed it
        uint2* yl = p.y1LL + (size_t)par * E;
        {
            float* xs = xin + stage_par * XN;
            poll_vector(skl, S, tag, xs, true);
            const float* w = stage_weights(2 * NL, rowA, S);
            WORKER_SYNC();
            float acc = 0.f;
#pragma unroll
            for (int it = 0; it < FAST_MAXI; ++it)
                if (it < nA) {
                    const float4 z4 = *reinterpret_cast<const float4*>(xs + 4 * (gA + it * 32));
                    const float4 w4 = ldw(w, gA + it * 32);
                    acc = fmaf(w4.x, z4.x, acc); acc = fmaf(w4.y, z4.y, acc); acc = fmaf(w4.z, z4.z, acc); acc = fmaf(w4.w, z4.w, acc);
                }
            acc = warp_sum(acc);
            if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
            release_slot();
        }
        WORKER_SYNC();
        if (*abort_s) return;
        if (tid < nE) {
            const int row = oE + tid;
            const float* ps = part + stage_par * GEN_WARPS + tid * HSA;
            float v = ps[0];
            for (int q = 1; q < HSA; ++q) v += ps[q];
            st_pair(yl + row, fmaxf(v + __ldg(p.e1b + row), 0.f), tag);
        }
        stage_par ^= 1;
        uint2* lgl = p.logitLL + (size_t)par * C;
        {
            float* xs = xin + stage_par * XN;
            poll_vector(yl, E, tag, xs, false);
            const float* w = stage_weights(2 * NL + 1, rowB, E);
            WORKER_SYNC();
            float acc = 0.f;
#pragma unroll
            for (int it = 0; it < FAST_MAXI; ++it)
                if (it < nB) {
                    const float4 z4 = *reinterpret_cast<const float4*>(xs + 4 * (gB + it * 32));
                    const float4 w4 = ldw(w, gB + it * 32);
                    acc = fmaf(w4.x, z4.x, acc); acc = fmaf(w4.y, z4.y, acc); acc = fmaf(w4.z, z4.z, acc); acc = fmaf(w4.w, z4.w, acc);
                }
            acc = warp_sum(acc);
            if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
            release_slot();
        }
        WORKER_SYNC();
        if (*abort_s) return;
        if (tid < nC) {
            const int row = oC + tid;
            const float* ps = part + stage_par * GEN_WARPS + tid * HSB;
            float v = ps[0];
            for (int q = 1; q < HSB; ++q) v += ps[q];
            const float dc = (float)row - (float)C / 2.f;
            v = (v + __ldg(p.e2b + row)) - (dc * dc) * p.regularize;
            st_pair(lgl + row, v, tag);
            if (p.out_logits) p.out_logits[(size_t)samp * C + row] = v;
        }
        stage_par ^= 1;
        // ---- all logits -> shared memory (cooperative poll), then warp 0 chooses
        float* logit_s = xin + stage_par * XN;
        poll_vector(lgl, C, tag, logit_s, false);
        WORKER_SYNC();
        if (*abort_s) return;
        if (warp == 0) {
            const int choice = choose_sample(logit_s, cdf, C, lane, p.temperature, p.uniforms ? p.uniforms + samp : nullptr);
            if (lane == 0) {
                misc[0] = choice;
                if (cta == 0) p.out_idx[samp] = choice;
            }
        }
        stage_par ^= 1;
        // the top-of-evaluation barrier publishes misc[0]
    }
    WORKER_SYNC();
    if (cta == 0 && tid == 0) p.cur_idx[0] = misc[0];
}

// ================================================================================================ cluster kernel
// One thread-block cluster (16 CTAs) per stream.  The stage-to-stage exchange no longer goes through the L2: a CTA
// that has computed a value stores the {value, tag} pair straight into the shared memory of all 16 CTAs of its
// cluster (distributed shared memory), and consumers spin on their OWN shared memory.  Measured on this part an L2
// all-to-all costs ~1300-1650 cycles per stage (tools/lat_probe.cu); a DSMEM store lands in ~200-250.
//   stage 1 (conv rows)   warp-local: local poll of the layer input -> 4 rows per warp over the full K -> warp reduce ->
//                         every lane recomputes z of the warp's 2 channels and stores it to one of the 16 CTAs
//   stage 2 (1x1 rows)    warps 0-3 residual rows, warps 4-7 skip rows; one CTA barrier per layer before h' is
//                         published keeps slow warps from being overtaken (buffers are double buffered by layer parity)
//   head                  skip -> end_conv_1 -> end_conv_2 -> every CTA holds all logits; rank 0 records the choice
// Clusters do not talk to each other, so a launch may hold any number of streams (they run in waves of co-resident
// clusters) and needs no cooperative launch.  History (ring slots of time t-d) still lives in global memory and is
// fetched one stage ahead with cp.async, as in gen_kernel_fast; weights stream through a TMA ring fed by warp 8.
constexpr int CL = 16;                      // CTAs per cluster
constexpr int CL_ROWS = 4;                  // max rows per warp per stage

__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// {value, tag} into the same shared-memory offset of CTA `dst` of this cluster
__device__ __forceinline__ void st_remote_pair(const void* local_ptr, unsigned dst, float v, unsigned tag) {
    unsigned laddr = smem_u32(local_ptr), raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(dst));
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(raddr), "l"(w) : "memory");
}
// local spin on a shared-memory {value, tag} pair (volatile: the writer is another SM)
__device__ __forceinline__ float wait_local(const uint2* p, unsigned tag, int* abort_s) {
    const volatile unsigned long long* vp = reinterpret_cast<const volatile unsigned long long*>(p);
    unsigned long long w = *vp;
    if ((unsigned)(w >> 32) != tag) {
        const long long t0 = clock64();
        unsigned spins = 0;
        do {
            w = *vp;
            if ((++spins & 1023u) == 0 && clock64() - t0 > GEN_TIMEOUT_CYCLES) asm volatile("trap;");   // never hang the GPU
        } while ((unsigned)(w >> 32) != tag);
    }
    return __uint_as_float((unsigned)w);
}

__global__ void __launch_bounds__(GEN_NT + 32, 1) gen_kernel_cluster(const GenParams p) {
    extern __shared__ __align__(16) float sm[];
    // exchange buffers first: they must sit at the same offset in every CTA (mapa keeps the offset)
    uint2* xcur = reinterpret_cast<uint2*>(sm);                  // [2][R]  layer input (h), by layer parity
    uint2* xz = xcur + 2 * p.R;                                  // [2][D]  gated activation z, by layer parity
    uint2* xhead = xz + 2 * p.D;                                 // [S + E + C] skip, y1, logits of the current evaluation
    uint2* old_s = xhead + (p.S + p.E + p.C);                    // [2][R] prefetched old taps
    float* skacc = reinterpret_cast<float*>(old_s + 2 * p.R);    // [S/CL]
    float* cur_own = skacc + ((p.S / CL + 3) & ~3);              // [2][R/CL]
    float* part = cur_own + 2 * ((p.R / CL + 3) & ~3);           // [E/CL + C/CL + 32] row sums of the head stages
    float* logit_s = part + (((p.E + p.C) / CL + 32 + 3) & ~3);  // [C]
    double* cdf = reinterpret_cast<double*>(logit_s + ((p.C + 3) & ~3));      // [C]
    float* wbuf = reinterpret_cast<float*>(cdf + p.C);                        // [n_wslots][wslot_floats]
    unsigned long long* fullb = reinterpret_cast<unsigned long long*>(wbuf + (size_t)p.n_wslots * p.wslot_floats);
    unsigned long long* emptyb = fullb + 4;
    GenLayer* lay_s = reinterpret_cast<GenLayer*>(fullb + 8);
    int* slot_s = reinterpret_cast<int*>(lay_s + p.n_layers);
    int* misc = slot_s + p.n_layers;                             // [0] current index, [1] abort

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_rank();
    const int stream = blockIdx.x / CL, NS = p.NS;
    const int R = p.R, D = p.D, S = p.S, E = p.E, C = p.C, NL = p.n_layers;
    const int K1 = 2 * R;
    const int nD = D / CL, nR = R / CL, nS = S / CL, nE = E / CL, nC = C / CL;
    const int oD = rank * nD, oR = rank * nR, oS = rank * nS, oE = rank * nE, oC = rank * nC;
    const int NSLOT = p.n_wslots;
    const int rw1 = 2 * nD / GEN_WARPS, rw2 = (nR + nS) / GEN_WARPS;       // rows per warp in stage 1 / stage 2 (<= CL_ROWS)

    {   // zero the exchange buffers (tag 0 = nothing yet), copy the layer table
        unsigned long long* z0 = reinterpret_cast<unsigned long long*>(xcur);
        const int n0 = 2 * R + 2 * D + S + E + C + 2 * R;
        for (int i = tid; i < n0; i += GEN_NT + 32) z0[i] = 0ull;
        const int* src = reinterpret_cast<const int*>(p.layers);
        int* dst = reinterpret_cast<int*>(lay_s);
        for (int i = tid; i < NL * (int)(sizeof(GenLayer) / sizeof(int)); i += GEN_NT + 32) dst[i] = src[i];
    }
    if (tid == 0) {
        misc[0] = p.cur_idx[stream];
        misc[1] = 0;
        for (int i = 0; i < NSLOT; ++i) { mbar_init(fullb + i, 1); mbar_init(emptyb + i, GEN_WARPS); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    for (int l = tid; l < NL; l += GEN_NT) {
        const int len = lay_s[l].ring_len;
        slot_s[l] = (p.t0 + len - 1) % len;
    }
    cluster_sync_all();                                          // nobody may store into a peer before it is zeroed
    int* abort_s = misc + 1;
    const unsigned smask = (unsigned)NSLOT - 1u, sshift = (NSLOT == 4) ? 2u : 1u;

    // ---- producer warp: weight rows of this CTA for every stage, in order, through the TMA ring
    if (warp == GEN_WARPS) {
        if (lane == 0) {
            unsigned q = 0;
            for (int ev = 0; ev < p.n_evals; ++ev) {
                const bool wh = (p.t0 + ev >= p.n_given - 1);
                const int n_st = wh ? 2 * NL + 2 : 2 * NL;
                for (int st = 0; st < n_st; ++st, ++q) {
                    StageDesc d = stage_desc(p, st, true, nD, nR, nS, nE, nC);
                    if (st < 2 * NL && (st & 1)) { d.n_first = nR; d.n = nR + nS; }
                    const int slot = (int)(q & smask);
                    if (q >= (unsigned)NSLOT) {
                        const unsigned par = ((q >> sshift) & 1u) ^ 1u;
                        unsigned done = 0, spins = 0;
                        while (!done) {
                            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(done) : "r"(smem_u32(emptyb + slot)), "r"(par) : "memory");
                            if (!done && ++spins > (1u << 30)) asm volatile("trap;");
                        }
                    }
                    mbar_expect_tx(fullb + slot, (unsigned)(d.n * d.K * 4));
                    float* dst = wbuf + (size_t)slot * p.wslot_floats;
                    for (int i = 0; i < d.n; ++i)
                        bulk_g2s(dst + (size_t)i * d.K, stage_row(p, lay_s, st, d, i, rank, CL), d.K * 4, fullb + slot);
                }
            }
        }
        cluster_sync_all();                                      // matches the workers' final cluster barrier
        return;
    }
    unsigned cons_q = 0;
    auto stage_weights = [&](int row, int K) -> const float* {
        const int slot = (int)(cons_q & smask);
        mbar_wait(fullb + slot, (cons_q >> sshift) & 1u);
        return wbuf + (size_t)slot * p.wslot_floats + (size_t)row * K;
    };
    auto release_slot = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive_(emptyb + (cons_q & smask));
        ++cons_q;
    };
    const size_t ring_stream = (size_t)stream * R;               // ring element (slot, stream, r): (slot*NS + stream)*R + r
    auto prefetch_old = [&](int ln, int te, int slot_te) {
        const GenLayer& Lp = lay_s[ln];
        if (te >= Lp.dil && tid < R / 2) {
            const int so = (slot_te + 1 == Lp.ring_len) ? 0 : slot_te + 1;
            const uint2* src = p.ringLL + Lp.ring_off + (size_t)so * NS * R + ring_stream + 2 * tid;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(old_s + (ln & 1) * R + 2 * tid);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch_old(0, p.t0, p.t0 % lay_s[0].ring_len);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    WORKER_SYNC();
    unsigned seq = (unsigned)p.t0 * (unsigned)(2 * NL + 4);      // exchange tag counter, unique per (evaluation, stage)

    for (int ev = 0; ev < p.n_evals; ++ev) {
        const int t = p.t0 + ev;
        const unsigned rtag = (unsigned)t + 1u;                  // ring tag of time t
        const bool want_head = (t >= p.n_given - 1);
        const int samp = t - (p.n_given - 1);
        if (tid == 0) {
            if (t < p.n_given) misc[0] = p.first[(size_t)stream * p.n_given + t];
            else if (p.forced != nullptr) misc[0] = p.forced[(size_t)stream * p.n_samples + (t - p.n_given)];
        }
        for (int l = tid; l < NL; l += GEN_NT) {
            const int s1 = slot_s[l] + 1;
            slot_s[l] = (s1 == lay_s[l].ring_len) ? 0 : s1;
        }
        for (int i = tid; i < nS; i += GEN_NT) skacc[i] = 0.f;
        asm volatile("cp.async.wait_group 0;" ::: "memory");    // layer 0's history, requested at the end of the last evaluation
        WORKER_SYNC();
        int idx = misc[0];
        idx = idx < 0 ? 0 : (idx >= C ? C - 1 : idx);
        const unsigned etag = seq + 1u;                          // tags of this evaluation: etag + 2l (cur), etag + 2l + 1 (z)
        seq += (unsigned)(2 * NL + 4);

        // layer 0's input: the start-conv column, computed locally by every CTA; owners also enqueue it
        {
            const GenLayer& L0 = lay_s[0];
            for (int r = tid; r < R; r += GEN_NT) {
                const float v = __ldg(p.start_w + (size_t)r * C + idx) + (p.start_b ? __ldg(p.start_b + r) : 0.f);
                reinterpret_cast<unsigned long long*>(xcur)[r] =
                    ((unsigned long long)etag << 32) | (unsigned long long)__float_as_uint(v);
                if (r >= oR && r < oR + nR)
                    st_pair(p.ringLL + L0.ring_off + ((size_t)slot_s[0] * NS) * R + ring_stream + r, v, rtag);
            }
        }
        WORKER_SYNC();

        for (int l = 0; l < NL; ++l) {
            const GenLayer& L = lay_s[l];
            const bool have_old = (t >= L.dil);
            const unsigned tag_old = (unsigned)(t - L.dil) + 1u, tag_cur = etag + 2u * (unsigned)l, tag_z = tag_cur + 1u;
            const uint2* xc = xcur + (l & 1) * R;
            const uint2* os = old_s + (l & 1) * R;
            uint2* zbuf = xz + (l & 1) * D;
            float* cur_l = cur_own + (l & 1) * ((nR + 3) & ~3);
            // ================= stage 1: rows [warp*rw1, +rw1) of this CTA's 2*nD conv rows, full K, warp-local
            {
                const float* w = stage_weights(warp * rw1, K1);
                float acc[CL_ROWS];
#pragma unroll
                for (int j = 0; j < CL_ROWS; ++j) acc[j] = 0.f;
                for (int g4 = lane; g4 < (K1 >> 2); g4 += 32) {
                    const int r0 = 2 * g4;
                    float o0 = 0.f, o1 = 0.f;
                    if (have_old) {
                        uint4 q = *reinterpret_cast<const uint4*>(os + r0);
                        if (q.y != tag_old || q.w != tag_old) {          // prefetched copy not there yet (rare): go to the ring
                            const int so = (slot_s[l] + 1 == L.ring_len) ? 0 : slot_s[l] + 1;
                            poll2(p.ringLL + L.ring_off + (size_t)so * NS * R + ring_stream + r0, tag_old, o0, o1, p.err, abort_s);
                        } else {
                            o0 = __uint_as_float(q.x);
                            o1 = __uint_as_float(q.z);
                        }
                    }
                    const float c0 = wait_local(xc + r0, tag_cur, abort_s), c1 = wait_local(xc + r0 + 1, tag_cur, abort_s);
                    if (r0 >= oR && r0 < oR + nR) cur_l[r0 - oR] = c0;           // every warp writes the same values
                    if (r0 + 1 >= oR && r0 + 1 < oR + nR) cur_l[r0 + 1 - oR] = c1;
#pragma unroll
                    for (int j = 0; j < CL_ROWS; ++j)
                        if (j < rw1) {
                            const float4 w4 = reinterpret_cast<const float4*>(w + (size_t)j * K1)[g4];
                            float a = acc[j];
                            a = fmaf(w4.x, o0, a); a = fmaf(w4.y, c0, a); a = fmaf(w4.z, o1, a); a = fmaf(w4.w, c1, a);
                            acc[j] = a;
                        }
                }
#pragma unroll
                for (int j = 0; j < CL_ROWS; ++j)
                    if (j < rw1) acc[j] = warp_sum(acc[j]);
                release_slot();
                // rows come in (filter, gate) pairs: channel ci = (warp*rw1 + 2*jj)/2; every lane recomputes z and
                // lane -> (channel jj = lane / 16, destination CTA = lane % 16)
                const int dst = lane & 15;
#pragma unroll
                for (int jj = 0; jj < CL_ROWS / 2; ++jj)
                    if (2 * jj < rw1 && ((lane >> 4) == (jj & 1))) {
                        const int ci = (warp * rw1 >> 1) + jj, c = oD + ci;
                        const float f = acc[2 * jj] + (L.bf ? __ldg(L.bf + c) : 0.f);
                        const float g = acc[2 * jj + 1] + (L.bg ? __ldg(L.bg + c) : 0.f);
                        st_remote_pair(zbuf + c, (unsigned)dst, tanh_(f) * sigmoid_(g), tag_z);
                    }
            }
            // ================= stage 2: warps [0, nR/rw2) residual rows, the rest skip rows; K = D
            {
                if (l + 1 < NL) prefetch_old(l + 1, t, slot_s[l + 1]);
                const int row0 = warp * rw2;
                const bool is_res = row0 < nR;
                const bool active = is_res ? (l + 1 < NL) : want_head;
                const float* w = stage_weights(row0, D);
                float acc[CL_ROWS];
#pragma unroll
                for (int j = 0; j < CL_ROWS; ++j) acc[j] = 0.f;
                if (active) {
                    for (int g4 = lane; g4 < (D >> 2); g4 += 32) {
                        const float z0 = wait_local(zbuf + 4 * g4, tag_z, abort_s), z1 = wait_local(zbuf + 4 * g4 + 1, tag_z, abort_s);
                        const float z2 = wait_local(zbuf + 4 * g4 + 2, tag_z, abort_s), z3 = wait_local(zbuf + 4 * g4 + 3, tag_z, abort_s);
#pragma unroll
                        for (int j = 0; j < CL_ROWS; ++j)
                            if (j < rw2) {
                                const float4 w4 = reinterpret_cast<const float4*>(w + (size_t)j * D)[g4];
                                float a = acc[j];
                                a = fmaf(w4.x, z0, a); a = fmaf(w4.y, z1, a); a = fmaf(w4.z, z2, a); a = fmaf(w4.w, z3, a);
                                acc[j] = a;
                            }
                    }
#pragma unroll
                    for (int j = 0; j < CL_ROWS; ++j)
                        if (j < rw2) acc[j] = warp_sum(acc[j]);
                }
                release_slot();
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                // all warps of this CTA are past their reads of z(l) and h(l) before h'(l) leaves (see header comment)
                WORKER_SYNC();
                if (l + 1 == NL && ev + 1 < p.n_evals)       // history of the next evaluation's layer 0 (buffer 0 is free now)
                    prefetch_old(0, t + 1, (slot_s[0] + 1 == lay_s[0].ring_len) ? 0 : slot_s[0] + 1);
                        if (is_res) {
                    if (l + 1 < NL) {
                        const GenLayer& Ln = lay_s[l + 1];
                        uint2* xn = xcur + ((l + 1) & 1) * R;
                        const unsigned tag_n = tag_cur + 2u;
                        // lane -> (row j = lane / 16 + 2*pass, destination CTA = lane % 16)
#pragma unroll
                        for (int ps = 0; ps < CL_ROWS / 2; ++ps) {
                            const int j = (lane >> 4) + 2 * ps;
                            if (j < rw2) {
                                const int li = row0 + j, row = oR + li;
                                float v = (j == 0) ? acc[0] : (j == 1) ? acc[1] : (j == 2) ? acc[2] : acc[3];
                                v += L.br ? __ldg(L.br + row) : 0.f;
                                v += cur_l[li];
                                st_remote_pair(xn + row, (unsigned)(lane & 15), v, tag_n);
                                if ((lane & 15) == 0)                      // history for the taps d steps from now
                                    st_pair(p.ringLL + Ln.ring_off + ((size_t)slot_s[l + 1] * NS) * R + ring_stream + row, v, rtag);
                            }
                        }
                    }
                } else if (want_head && lane == 0) {
#pragma unroll
                    for (int j = 0; j < CL_ROWS; ++j)
                        if (j < rw2) {
                            const int li = row0 + j - nR, row = oS + li;
                            const float v = acc[j] + (L.bs ? __ldg(L.bs + row) : 0.f);
                            skacc[li] = v + skacc[li];
                        }
                }
            }
        }
        if (!want_head) continue;

        // ================= head (3 exchanges per evaluation; rows spread one per warp-iteration)
        const unsigned tag_s = etag + 2u * (unsigned)NL + 1u, tag_y = tag_s + 1u, tag_l = tag_s + 2u;
        uint2* xs = xhead, *xy = xhead + S, *xl = xhead + S + E;
        WORKER_SYNC();                                           // skacc complete (written by the skip warps' lane 0)
        for (int i = tid; i < nS * CL; i += GEN_NT) st_remote_pair(xs + oS + (i >> 4), (unsigned)(i & 15), skacc[i >> 4], tag_s);
        {
            const float* w = stage_weights(0, S);
            for (int it = warp; it < nE; it += GEN_WARPS) {
                float acc = 0.f;
                for (int g4 = lane; g4 < (S >> 2); g4 += 32) {
                    const float4 w4 = reinterpret_cast<const float4*>(w + (size_t)it * S)[g4];
                    acc = fmaf(w4.x, fmaxf(wait_local(xs + 4 * g4, tag_s, abort_s), 0.f), acc);
                    acc = fmaf(w4.y, fmaxf(wait_local(xs + 4 * g4 + 1, tag_s, abort_s), 0.f), acc);
                    acc = fmaf(w4.z, fmaxf(wait_local(xs + 4 * g4 + 2, tag_s, abort_s), 0.f), acc);
                    acc = fmaf(w4.w, fmaxf(wait_local(xs + 4 * g4 + 3, tag_s, abort_s), 0.f), acc);
                }
                acc = warp_sum(acc);
                const int row = oE + it;
                if (lane < CL) st_remote_pair(xy + row, (unsigned)lane, fmaxf(acc + __ldg(p.e1b + row), 0.f), tag_y);
            }
            release_slot();
        }
        {
            const float* w = stage_weights(0, E);
            for (int it = warp; it < nC; it += GEN_WARPS) {
                float acc = 0.f;
                for (int g4 = lane; g4 < (E >> 2); g4 += 32) {
                    const float4 w4 = reinterpret_cast<const float4*>(w + (size_t)it * E)[g4];
                    acc = fmaf(w4.x, wait_local(xy + 4 * g4, tag_y, abort_s), acc);
                    acc = fmaf(w4.y, wait_local(xy + 4 * g4 + 1, tag_y, abort_s), acc);
                    acc = fmaf(w4.z, wait_local(xy + 4 * g4 + 2, tag_y, abort_s), acc);
                    acc = fmaf(w4.w, wait_local(xy + 4 * g4 + 3, tag_y, abort_s), acc);
                }
                acc = warp_sum(acc);
                const int row = oC + it;
                const float dc = (float)row - (float)C / 2.f;
                const float v = (acc + __ldg(p.e2b + row)) - (dc * dc) * p.regularize;
                if (lane < CL) st_remote_pair(xl + row, (unsigned)lane, v, tag_l);
                if (lane == 0 && p.out_logits) p.out_logits[((size_t)stream * p.n_samples + samp) * C + row] = v;
            }
            release_slot();
        }
        for (int c = tid; c < C; c += GEN_NT) logit_s[c] = wait_local(xl + c, tag_l, abort_s);
        WORKER_SYNC();
        if (warp == 0) {
            const int choice = choose_sample(logit_s, cdf, C, lane, p.temperature,
                                             p.uniforms ? p.uniforms + (size_t)stream * p.n_samples + samp : nullptr);
            if (lane == 0) {
                misc[0] = choice;
                if (rank == 0) p.out_idx[(size_t)stream * p.n_samples + samp] = choice;
            }
        }
    }
    WORKER_SYNC();
    if (rank == 0 && tid == 0) p.cur_idx[stream] = misc[0];
    cluster_sync_all();                                          // peers may still be storing into this CTA's shared memory
}

// ================================================================================================ two-level exchange kernel
// Single stream, the grid of gen_kernel_fast (64 CTAs x 4 rows per stage vector) organised as 4 thread-block clusters of 16.
// gen_kernel_fast pays ~1650 cycles per stage for an all-to-all in which 64 CTAs poll all 256 {value, tag} pairs through the
// L2.  Here a value travels two hops instead:
//   * inside a cluster the producer stores it straight into the shared memory of its 16 CTAs (distributed shared memory,
//     ~250 cycles) -- and, once, to the L2 buffer the other kernels use (the ring history needs that anyway);
//   * between clusters ONE CTA per destination cluster polls it in the L2 -- rank r of cluster c fetches the 32-byte sector
//     of the four values that rank r of each other cluster produced (3 sectors per stage instead of 64) -- and forwards it
//     to its 16 cluster peers through DSMEM.
// Every consumer then spins on its OWN shared memory (no polling storm on hot L2 lines, no staging pass and one CTA barrier
// per stage instead of two).  Same row ownership, K split, summation order and activations as gen_kernel_fast /
// gen_kernel_ll: logits and indices are bit-identical to theirs.
// MEASURED (B200, cfg 2, 4000 samples): 299 us/sample against 150 us/sample for gen_kernel_fast.  The all-to-all inside a
// 16-CTA cluster alone costs 910-940 cycles per round (tools/dsmem_probe.cu, variant A) -- no better than the L2 all-to-all
// it replaces -- and the second hop is added on top.  Kept as mode 5 (selectable, tested bit-identical), never the default.
__device__ __forceinline__ void wait_local2(const uint2* p, unsigned tag, float& a, float& b) {        // two pairs, 16-byte aligned
    const unsigned addr = smem_u32(p);
    unsigned x, y, z, w;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
    if (y != tag || w != tag) {
        const long long t0 = clock64();
        unsigned spins = 0;
        do {
            asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
            if ((++spins & 1023u) == 0 && clock64() - t0 > GEN_TIMEOUT_CYCLES) asm volatile("trap;");
        } while (y != tag || w != tag);
    }
    a = __uint_as_float(x);
    b = __uint_as_float(z);
}

__global__ void __launch_bounds__(GEN_NT + 32, 1) gen_kernel_x2(const GenParams p) {
    extern __shared__ __align__(16) float sm[];
    // exchange buffers first: they must sit at the same offset in every CTA (mapa keeps the offset)
    uint2* Xcur = reinterpret_cast<uint2*>(sm);                  // [2][R]  layer input h, by layer parity
    uint2* Xz = Xcur + 2 * p.R;                                  // [2][D]  gated activation z, by layer parity
    uint2* Xhead = Xz + 2 * p.D;                                 // [S + E + C] skip, y1, logits of the current evaluation
    uint2* old_s = Xhead + (p.S + p.E + p.C);                    // [2][R] prefetched old taps
    float* part = reinterpret_cast<float*>(old_s + 2 * p.R);     // [2][GEN_WARPS] partial sums, double buffered by stage parity
    float* skacc = part + 2 * GEN_WARPS;                         // [nS]
    float* logit_s = skacc + 4;                                  // [C]
    double* cdf = reinterpret_cast<double*>(logit_s + ((p.C + 3) & ~3));      // [C]
    float* wbuf = reinterpret_cast<float*>(cdf + p.C);                        // [n_wslots][wslot_floats]
    unsigned long long* fullb = reinterpret_cast<unsigned long long*>(wbuf + (size_t)p.n_wslots * p.wslot_floats);
    unsigned long long* emptyb = fullb + 4;
    GenLayer* lay_s = reinterpret_cast<GenLayer*>(fullb + 8);
    int* slot_s = reinterpret_cast<int*>(lay_s + p.n_layers);
    int* misc = slot_s + p.n_layers;                             // [0] current index

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int rank = (int)cluster_rank(), cl = cta / CL, NCL = G / CL;
    const int R = p.R, D = p.D, S = p.S, E = p.E, C = p.C, NL = p.n_layers;
    const int K1 = 2 * R;
    constexpr int NV = 4;                                        // values per CTA per stage vector (host checks D/G = R/G = ... = 4)
    const int oV = cta * NV;                                     // first index this CTA owns in every stage vector
    const int NSLOT = p.n_wslots;
    // fixed warp -> (row, K-part) assignment per stage kind, as in gen_kernel_fast (same summation order)
    const int HS1 = GEN_WARPS / (2 * NV), HS2 = GEN_WARPS / (2 * NV), HSA = GEN_WARPS / NV, HSB = GEN_WARPS / NV;
    const int row1 = warp / HS1, g1 = ((warp - row1 * HS1) * (K1 / HS1) >> 2) + lane, n1 = (((K1 / HS1) >> 2) - lane + 31) / 32;
    const int row2 = warp / HS2, g2 = ((warp - row2 * HS2) * (D / HS2) >> 2) + lane, n2 = (((D / HS2) >> 2) - lane + 31) / 32;
    const int rowA = warp / HSA, gA = ((warp - rowA * HSA) * (S / HSA) >> 2) + lane, nA = (((S / HSA) >> 2) - lane + 31) / 32;
    const int rowB = warp / HSB, gB = ((warp - rowB * HSB) * (E / HSB) >> 2) + lane, nB = (((E / HSB) >> 2) - lane + 31) / 32;

    {   // zero the exchange buffers (tag 0 = nothing yet), copy the layer table
        unsigned long long* z0 = reinterpret_cast<unsigned long long*>(Xcur);
        const int n0 = 2 * R + 2 * D + S + E + C + 2 * R;
        for (int i = tid; i < n0; i += GEN_NT + 32) z0[i] = 0ull;
        const int* src = reinterpret_cast<const int*>(p.layers);
        int* dst = reinterpret_cast<int*>(lay_s);
        for (int i = tid; i < NL * (int)(sizeof(GenLayer) / sizeof(int)); i += GEN_NT + 32) dst[i] = src[i];
    }
    if (tid == 0) {
        misc[0] = p.cur_idx[0];
        for (int i = 0; i < NSLOT; ++i) { mbar_init(fullb + i, 1); mbar_init(emptyb + i, GEN_WARPS); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    for (int l = tid; l < NL; l += GEN_NT) {
        const int len = lay_s[l].ring_len;
        slot_s[l] = (p.t0 + len - 1) % len;
    }
    cluster_sync_all();                                          // nobody may store into a peer before it is zeroed
    const unsigned smask = (unsigned)NSLOT - 1u, sshift = (NSLOT == 4) ? 2u : 1u;

    // ---- producer warp: weight rows of this CTA for every stage, in order, through the TMA ring (as gen_kernel_fast)
    if (warp == GEN_WARPS) {
        if (lane == 0) {
            unsigned q = 0;
            for (int ev = 0; ev < p.n_evals; ++ev) {
                const bool wh = (p.t0 + ev >= p.n_given - 1);
                const int n_st = wh ? 2 * NL + 2 : 2 * NL;
                for (int st = 0; st < n_st; ++st, ++q) {
                    StageDesc d = stage_desc(p, st, true, NV, NV, NV, NV, NV);
                    if (st < 2 * NL && (st & 1)) { d.n_first = NV; d.n = 2 * NV; }
                    const int slot = (int)(q & smask);
                    if (q >= (unsigned)NSLOT) {
                        const unsigned par = ((q >> sshift) & 1u) ^ 1u;
                        unsigned done = 0, spins = 0;
                        while (!done) {
                            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(done) : "r"(smem_u32(emptyb + slot)), "r"(par) : "memory");
                            if (!done && ++spins > (1u << 30)) asm volatile("trap;");
                        }
                    }
                    mbar_expect_tx(fullb + slot, (unsigned)(d.n * d.K * 4));
                    float* dst = wbuf + (size_t)slot * p.wslot_floats;
                    for (int i = 0; i < d.n; ++i)
                        bulk_g2s(dst + (size_t)i * d.K, stage_row(p, lay_s, st, d, i, cta, G), d.K * 4, fullb + slot);
                }
            }
        }
        cluster_sync_all();                                      // matches the workers' final cluster barrier
        return;
    }
    unsigned cons_q = 0;
    auto stage_weights = [&](int row, int K) -> const float* {
        const int slot = (int)(cons_q & smask);
        mbar_wait(fullb + slot, (cons_q >> sshift) & 1u);
        return wbuf + (size_t)slot * p.wslot_floats + (size_t)row * K;
    };
    auto release_slot = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive_(emptyb + (cons_q & smask));
        ++cons_q;
    };
    auto prefetch_old = [&](int ln, int te, int slot_te) {
        const GenLayer& Lp = lay_s[ln];
        if (te >= Lp.dil && tid < R / 2) {
            const int so = (slot_te + 1 == Lp.ring_len) ? 0 : slot_te + 1;
            const uint2* src = p.ringLL + Lp.ring_off + (size_t)so * R + 2 * tid;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(old_s + (ln & 1) * R + 2 * tid);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Second hop: warps 4 .. 4+NCL-2 each serve one OTHER cluster: fetch the sector (4 pairs) its rank-`rank` CTA published to
    // the L2 vector `gvec` (tag gtag) and store it, re-tagged, into vector `xvec` of all 16 CTAs of this cluster.
    auto forward = [&](const uint2* gvec, unsigned gtag, uint2* xvec, unsigned xtag) {
        const int k = warp - 4;
        if (k < 0 || k >= NCL - 1) return;
        const int cp = (cl + 1 + k) % NCL;
        const int idx = (cp * CL + rank) * NV + 2 * (lane >> 4);        // lanes 0-15: pairs 0,1; lanes 16-31: pairs 2,3
        Pair2 q = ld_pair2(gvec + idx);
        if (q.a.y != gtag || q.b.y != gtag) {
            const long long t0 = clock64();
            unsigned spins = 0;
            do {
                q = ld_pair2(gvec + idx);
                if ((++spins & 255u) == 0 && clock64() - t0 > GEN_TIMEOUT_CYCLES) asm volatile("trap;");
            } while (q.a.y != gtag || q.b.y != gtag);
        }
        st_remote_pair(xvec + idx, (unsigned)(lane & 15), __uint_as_float(q.a.x), xtag);
        st_remote_pair(xvec + idx + 1, (unsigned)(lane & 15), __uint_as_float(q.b.x), xtag);
    };
    {
        const int len0 = lay_s[0].ring_len;
        prefetch_old(0, p.t0, p.t0 % len0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    WORKER_SYNC();
    unsigned stage_par = 0;
    unsigned seq = (unsigned)p.t0 * (unsigned)(2 * NL + 4);      // exchange tag counter, unique per (evaluation, stage)

    for (int ev = 0; ev < p.n_evals; ++ev) {
        const int t = p.t0 + ev;
        const unsigned rtag = (unsigned)t + 1u;                  // tag of time t in the L2 buffers (ring history, exchange vectors)
        const int par = t & 1;
        const bool want_head = (t >= p.n_given - 1);
        const int samp = t - (p.n_given - 1);
        if (tid == 0) {
            if (t < p.n_given) misc[0] = p.first[t];
            else if (p.forced != nullptr) misc[0] = p.forced[t - p.n_given];
        }
        for (int l = tid; l < NL; l += GEN_NT) {
            const int s1 = slot_s[l] + 1;
            slot_s[l] = (s1 == lay_s[l].ring_len) ? 0 : s1;
        }
        if (tid < NV) skacc[tid] = 0.f;
        WORKER_SYNC();
        int idx = misc[0];
        idx = idx < 0 ? 0 : (idx >= C ? C - 1 : idx);
        const unsigned etag = seq + 1u;                          // tags of this evaluation: etag + 2l (h of layer l), etag + 2l + 1 (z)
        seq += (unsigned)(2 * NL + 4);

        // layer 0's input: the start-conv column, computed locally by every CTA; owners also enqueue it in the L2 ring
        {
            const GenLayer& L0 = lay_s[0];
            for (int r = tid; r < R; r += GEN_NT) {
                const float v = __ldg(p.start_w + (size_t)r * C + idx) + (p.start_b ? __ldg(p.start_b + r) : 0.f);
                reinterpret_cast<volatile unsigned long long*>(Xcur)[r] =
                    ((unsigned long long)etag << 32) | (unsigned long long)__float_as_uint(v);
                if (r >= oV && r < oV + NV) st_pair(p.ringLL + L0.ring_off + (size_t)slot_s[0] * R + r, v, rtag);
            }
        }

        for (int l = 0; l < NL; ++l) {
            const GenLayer& L = lay_s[l];
            const int slot_t = slot_s[l];
            const int slot_old = (slot_t + 1 == L.ring_len) ? 0 : slot_t + 1;
            const bool have_old = (t >= L.dil);
            const unsigned tag_old = (unsigned)(t - L.dil) + 1u, tag_cur = etag + 2u * (unsigned)l, tag_z = tag_cur + 1u;
            const uint2* xc = Xcur + (l & 1) * R;
            const uint2* os = old_s + (l & 1) * R;
            uint2* zb = Xz + (l & 1) * D;
            uint2* zl = p.zLL + (size_t)(par * NL + l) * D;
            // ================= stage 1: filter/gate rows, K = 2R interleaved (old, cur) per channel
            {
                const float* w1 = stage_weights(row1, K1);
                float acc = 0.f;
                bool old_ok = true;
                float cc[FAST_MAXI][2];
#pragma unroll
                for (int it = 0; it < FAST_MAXI; ++it)
                    if (it < n1) wait_local2(xc + 2 * (g1 + it * 32), tag_cur, cc[it][0], cc[it][1]);
#pragma unroll
                for (int it = 0; it < FAST_MAXI; ++it)
                    if (it < n1) {
                        const int g4 = g1 + it * 32, r0 = 2 * g4;
                        float o0 = 0.f, o1 = 0.f;
                        if (have_old) {
                            const uint4 q = *reinterpret_cast<const uint4*>(os + r0);
                            old_ok = old_ok && q.y == tag_old && q.w == tag_old;
                            o0 = __uint_as_float(q.x);
                            o1 = __uint_as_float(q.z);
                        }
                        const float4 w4 = reinterpret_cast<const float4*>(w1)[g4];
                        acc = fmaf(w4.x, o0, acc); acc = fmaf(w4.y, cc[it][0], acc); acc = fmaf(w4.z, o1, acc); acc = fmaf(w4.w, cc[it][1], acc);
                    }
                if (!__all_sync(0xffffffffu, old_ok)) {            // prefetched copy not there yet (rare): poll the ring itself
                    acc = 0.f;
                    for (int it = 0; it < n1; ++it) {
                        const int g4 = g1 + it * 32, r0 = 2 * g4;
                        float o0, o1;
                        poll2(p.ringLL + L.ring_off + (size_t)slot_old * R + r0, tag_old, o0, o1, p.err, misc + 1);
                        const float4 w4 = reinterpret_cast<const float4*>(w1)[g4];
                        acc = fmaf(w4.x, o0, acc); acc = fmaf(w4.y, cc[it][0], acc); acc = fmaf(w4.z, o1, acc); acc = fmaf(w4.w, cc[it][1], acc);
                    }
                }
                acc = warp_sum(acc);
                if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
                release_slot();
            }
            WORKER_SYNC();
            if (tid < CL * NV) {                                   // publish z: thread -> (value tid / 16, destination CTA tid % 16)
                const int vi = tid >> 4, c = oV + vi;
                const float* pf = part + stage_par * GEN_WARPS + (2 * vi) * HS1;
                float f = pf[0], g = pf[HS1];
                for (int q = 1; q < HS1; ++q) { f += pf[q]; g += pf[HS1 + q]; }
                f += L.bf ? __ldg(L.bf + c) : 0.f;
                g += L.bg ? __ldg(L.bg + c) : 0.f;
                const float zv = tanh_(f) * sigmoid_(g);
                st_remote_pair(zb + c, (unsigned)(tid & 15), zv, tag_z);
                if ((tid & 15) == 0) st_pair(zl + c, zv, rtag);
            }
            forward(zl, rtag, zb, tag_z);
            stage_par ^= 1;
            // ================= stage 2: residual rows (first NV) and skip rows (next NV), K = D
            if (l + 1 < NL) prefetch_old(l + 1, t, slot_s[l + 1]);
            else if (ev + 1 < p.n_evals) prefetch_old(0, t + 1, (slot_s[0] + 1 == lay_s[0].ring_len) ? 0 : slot_s[0] + 1);
            {
                const bool active2 = (row2 < NV) ? (l + 1 < NL) : want_head;
                const float* w2 = stage_weights(row2, D);
                float acc = 0.f;
                if (active2) {
                    float zz[FAST_MAXI][4];
#pragma unroll
                    for (int it = 0; it < FAST_MAXI; ++it)
                        if (it < n2) {
                            wait_local2(zb + 4 * (g2 + it * 32), tag_z, zz[it][0], zz[it][1]);
                            wait_local2(zb + 4 * (g2 + it * 32) + 2, tag_z, zz[it][2], zz[it][3]);
                        }
#pragma unroll
                    for (int it = 0; it < FAST_MAXI; ++it)
                        if (it < n2) {
                            const float4 w4 = reinterpret_cast<const float4*>(w2)[g2 + it * 32];
                            acc = fmaf(w4.x, zz[it][0], acc); acc = fmaf(w4.y, zz[it][1], acc); acc = fmaf(w4.z, zz[it][2], acc); acc = fmaf(w4.w, zz[it][3], acc);
                        }
                    acc = warp_sum(acc);
                }
                if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
                release_slot();
                asm volatile("cp.async.wait_group 0;" ::: "memory");      // the old taps for the next stage 1 have landed
            }
            WORKER_SYNC();
            if (tid < CL * NV) {
                if (l + 1 < NL) {                                  // publish h' = Wr z + br + h
                    const int vi = tid >> 4, row = oV + vi;
                    const GenLayer& Ln = lay_s[l + 1];
                    const float* ps = part + stage_par * GEN_WARPS + vi * HS2;
                    float v = ps[0];
                    for (int q = 1; q < HS2; ++q) v += ps[q];
                    v += L.br ? __ldg(L.br + row) : 0.f;
                    const float hv = v + wait_local(xc + row, tag_cur, misc + 1);
                    st_remote_pair(Xcur + ((l + 1) & 1) * R + row, (unsigned)(tid & 15), hv, tag_cur + 2u);
                    if ((tid & 15) == 0) st_pair(p.ringLL + Ln.ring_off + (size_t)slot_s[l + 1] * R + row, hv, rtag);
                }
            } else if (tid < CL * NV + NV && want_head) {          // skip rows accumulate locally
                const int li = tid - CL * NV;
                const float* ps = part + stage_par * GEN_WARPS + (NV + li) * HS2;
                float v = ps[0];
                for (int q = 1; q < HS2; ++q) v += ps[q];
                v += L.bs ? __ldg(L.bs + oV + li) : 0.f;
                skacc[li] = v + skacc[li];
            }
            if (l + 1 < NL) {
                const GenLayer& Ln = lay_s[l + 1];
                forward(p.ringLL + Ln.ring_off + (size_t)slot_s[l + 1] * R, rtag, Xcur + ((l + 1) & 1) * R, tag_cur + 2u);
            }
            stage_par ^= 1;
        }
        if (!want_head) continue;

        // ================= head
        const unsigned tag_s = etag + 2u * (unsigned)NL + 1u, tag_y = tag_s + 1u, tag_l = tag_s + 2u;
        uint2* xs = Xhead, *xy = Xhead + S, *xl = Xhead + S + E;
        uint2* skl = p.skipLL + (size_t)par * S;
        uint2* yl = p.y1LL + (size_t)par * E;
        uint2* lgl = p.logitLL + (size_t)par * C;
        WORKER_SYNC();                                           // skacc complete
        if (tid < CL * NV) {
            const int vi = tid >> 4;
            st_remote_pair(xs + oV + vi, (unsigned)(tid & 15), skacc[vi], tag_s);
            if ((tid & 15) == 0) st_pair(skl + oV + vi, skacc[vi], rtag);
        }
        forward(skl, rtag, xs, tag_s);
        {
            const float* w = stage_weights(rowA, S);
            float zz[FAST_MAXI][4];
#pragma unroll
            for (int it = 0; it < FAST_MAXI; ++it)
                if (it < nA) {
                    wait_local2(xs + 4 * (gA + it * 32), tag_s, zz[it][0], zz[it][1]);
                    wait_local2(xs + 4 * (gA + it * 32) + 2, tag_s, zz[it][2], zz[it][3]);
                }
            float acc = 0.f;
#pragma unroll
            for (int it = 0; it < FAST_MAXI; ++it)
                if (it < nA) {
                    const float4 w4 = reinterpret_cast<const float4*>(w)[gA + it * 32];
                    acc = fmaf(w4.x, fmaxf(zz[it][0], 0.f), acc); acc = fmaf(w4.y, fmaxf(zz[it][1], 0.f), acc);
                    acc = fmaf(w4.z, fmaxf(zz[it][2], 0.f), acc); acc = fmaf(w4.w, fmaxf(zz[it][3], 0.f), acc);
                }
            acc = warp_sum(acc);
            if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
            release_slot();
        }
        WORKER_SYNC();
        if (tid < CL * NV) {
            const int vi = tid >> 4, row = oV + vi;
            const float* ps = part + stage_par * GEN_WARPS + vi * HSA;
            float v = ps[0];
            for (int q = 1; q < HSA; ++q) v += ps[q];
            v = fmaxf(v + __ldg(p.e1b + row), 0.f);
            st_remote_pair(xy + row, (unsigned)(tid & 15), v, tag_y);
            if ((tid & 15) == 0) st_pair(yl + row, v, rtag);
        }
        forward(yl, rtag, xy, tag_y);
        stage_par ^= 1;
        {
            const float* w = stage_weights(rowB, E);
            float zz[FAST_MAXI][4];
#pragma unroll
            for (int it = 0; it < FAST_MAXI; ++it)
                if (it < nB) {
                    wait_local2(xy + 4 * (gB + it * 32), tag_y, zz[it][0], zz[it][1]);
                    wait_local2(xy + 4 * (gB + it * 32) + 2, tag_y, zz[it][2], zz[it][3]);
                }
            float acc = 0.f;
#pragma unroll
            for (int it = 0; it < FAST_MAXI; ++it)
                if (it < nB) {
                    const float4 w4 = reinterpret_cast<const float4*>(w)[gB + it * 32];
                    acc = fmaf(w4.x, zz[it][0], acc); acc = fmaf(w4.y, zz[it][1], acc); acc = fmaf(w4.z, zz[it][2], acc); acc = fmaf(w4.w, zz[it][3], acc);
                }
            acc = warp_sum(acc);
            if (lane == 0) part[stage_par * GEN_WARPS + warp] = acc;
            release_slot();
        }
        WORKER_SYNC();
        if (tid < CL * NV) {
            const int vi = tid >> 4, row = oV + vi;
            const float* ps = part + stage_par * GEN_WARPS + vi * HSB;
            float v = ps[0];
            for (int q = 1; q < HSB; ++q) v += ps[q];
            const float dc = (float)row - (float)C / 2.f;
            v = (v + __ldg(p.e2b + row)) - (dc * dc) * p.regularize;
            st_remote_pair(xl + row, (unsigned)(tid & 15), v, tag_l);
            if ((tid & 15) == 0) {
                st_pair(lgl + row, v, rtag);
                if (p.out_logits) p.out_logits[(size_t)samp * C + row] = v;
            }
        }
        forward(lgl, rtag, xl, tag_l);
        stage_par ^= 1;
        for (int c = tid; c < C; c += GEN_NT) logit_s[c] = wait_local(xl + c, tag_l, misc + 1);
        WORKER_SYNC();
        if (warp == 0) {
            const int choice = choose_sample(logit_s, cdf, C, lane, p.temperature, p.uniforms ? p.uniforms + samp : nullptr);
            if (lane == 0) {
                misc[0] = choice;
                if (cta == 0) p.out_idx[samp] = choice;
            }
        }
        // the top-of-evaluation barrier publishes misc[0]
    }
    WORKER_SYNC();
    if (cta == 0 && tid == 0) p.cur_idx[0] = misc[0];
    cluster_sync_all();                                          // peers may still be storing into this CTA's shared memory
}

// ================================================================================================ batched cluster kernel
// Tensor-core sampler for 256-wide nets (R = D = S = E = classes = 256, k = 2), one stream or many: a thread-block cluster
// advances up to CL8_SB = 8 independent streams together, so the weights of a stage enter shared memory ONCE per 8 streams
// and step (gen_kernel_cluster streams all 79 MB per stream and step; only 7 of its clusters are co-resident, so 64 streams
// ran as waves).  Every exchanged vector is 16 blocks of 16 channels x 8 streams; a CTA owns one block (clusters of 16) or
// two (clusters of 8) and computes those rows of every stage for all 8 streams:
//   * the dot products are mma.sync m16n8k16 (M = 16 rows of the stage, N = the 8 streams, K = 16 input channels = one
//     block) with bf16 hi/lo operand pairs -- x = hi + lo up to 2^-17 relative, three MMAs per product (lo.hi, hi.lo, hi.hi
//     on independent accumulators), fp32 accumulation: the scheme of the training kernels (tc_block.cu).  Weights are
//     pre-split ONCE per session into fragment-ordered images (cl8_pack_kernel): a warp's A fragments are two conflict-free
//     LDS.128, a stage's image is one bulk copy.  Activations are exchanged already split, in B-fragment order: a k-step's
//     B operands are one LDS.128.  8 warps = 2 m-tiles x 4 K quarters; the partials meet in shared memory and 128-256
//     finishing threads apply bias / tanh.sigmoid / the residual add and stage the CTA's block(s);
//   * exchange: a pusher warp reads a staged 512-byte block back and issues ONE st.async.v4 per destination CTA, crediting
//     the bytes to an mbarrier there; consumers sleep on their own mbarrier.  893-1 017 cycles per all-to-all round against
//     1 125-1 254 with a bulk copy per destination and 2 624-4 144 with per-lane stores (tools/dsmem_probe.cu: H, F, E);
//   * weights: producer warp(s) keep a 128 KB ring of images full (bulk copies; cp.async for the second block of an 8-CTA
//     cluster's CTA);
//   * history: the {value, tag} fp32 ring of the other kernels (same layout: sessions, queue export and kernel switches keep
//     working), written by the owning CTA, fetched one stage ahead into registers, validated by tag, split on arrival.
// Streams never mix (N is the stream index of the MMA) and both cluster sizes add in the same order: a stream's indices and
// logits do not depend on how many streams run beside it, bit for bit.
constexpr int CL8_SB = 8;               // streams per cluster
constexpr int CL8_W = 256;              // the width this kernel is specialised for
constexpr int CL8_BLK = 512;            // bytes of one (source CTA) block of an exchanged vector: 8 streams x 16 channels x (hi, lo)
constexpr int CL8_VEC = CL * CL8_BLK;   // bytes of one exchanged vector in every CTA
constexpr int CL8_IMG1 = 2 * 32 * 1024, CL8_IMG2 = 2 * 16 * 1024, CL8_IMGH = 16 * 1024;    // weight images per (layer, rank) / head stage
// byte offset of the (hi pair | lo pair) unit of channels (c & ~1, c | 1) of stream s inside a block: per stream 64 bytes =
// 4 x [unit t | unit t+4], the order in which lane (g = stream, t) of an m16n8k16 B fragment consumes them
__device__ __forceinline__ int cl8_unit_off(int s, int c) {
    const int u = c >> 1;
    return s * 64 + (u & 3) * 16 + (u >> 2) * 8;
}
__device__ __forceinline__ void cl8_split(float x, unsigned short& hi, unsigned short& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}
// one value of a block, written as two 16-bit stores (the pair partner is written by another thread)
__device__ __forceinline__ void cl8_put(unsigned char* blk, int s, int c, float x) {
    unsigned short hi, lo;
    cl8_split(x, hi, lo);
    unsigned short* q = reinterpret_cast<unsigned short*>(blk + cl8_unit_off(s, c)) + (c & 1);
    q[0] = hi;
    q[2] = lo;
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint4& a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity) {
    unsigned done = 0, spins = 0;
    const unsigned a = smem_u32(bar);
    while (true) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 26)) asm volatile("trap;");        // never hang the GPU
    }
}
__device__ __forceinline__ unsigned mapa_u32(unsigned laddr, unsigned dst) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(dst));
    return r;
}

// Weight images.  One (m-tile, k-step) = 1 KB: [hi: 32 lanes x 16 B][lo: 32 lanes x 16 B], lane's 16 bytes = the A fragment
// registers a0..a3 of m16n8k16: a_j covers row g + 8*(j&1), k pair 2t + 8*(j>>1) (g = lane>>2, t = lane&3).
//   layout [layer][kind][virtual rank 0-15][32 KB image = [m-tile][k-step 0-15][hi | lo]], kind 0 = stage 1, tap 0 (old;
//   m-tile 0 = filter rows, 1 = gate rows; k-step = channel block), 1 = stage 1, tap 1 (current), 2 = stage 2 (m-tile 0 =
//   residual rows, 1 = skip rows; k-step = z block); then [end_conv_1 | end_conv_2][virtual rank][16 KB: one m-tile].
//   The images of consecutive virtual ranks are adjacent, so a CTA that owns VR of them fetches a stage with ONE copy.
__global__ void cl8_pack_kernel(const GenLayer* layers, int n_layers, const float* e1w, const float* e2w, unsigned* img) {
    const int W = CL8_W;
    constexpr size_t IW = CL8_IMG2 / 4, HW = CL8_IMGH / 4;          // words per layer image / head image
    const size_t per_layer = 3 * CL * IW, total = per_layer * n_layers + 2 * CL * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float* src;      // row-major weight matrix the word comes from
        int row, col, ld, stride = 1;
        size_t w = i;
        if (w < per_layer * n_layers) {
            const int l = (int)(w / per_layer);
            w -= (size_t)l * per_layer;
            const int kind = (int)(w / (CL * IW));          // 0: stage 1, tap 0 (old)   1: stage 1, tap 1 (current)   2: stage 2
            w -= (size_t)kind * CL * IW;
            const int vrank = (int)(w / IW);
            w -= (size_t)vrank * IW;
            const GenLayer& L = layers[l];
            const int j = (int)(w & 3), lane = (int)((w >> 2) & 31), ks = (int)((w >> 8) & 15), mt = (int)(w >> 12);
            const int g = lane >> 2, t = lane & 3;
            row = vrank * 16 + g + 8 * (j & 1);
            const int kk = 2 * t + 8 * (j >> 1);
            if (kind < 2) { src = mt ? L.wg : L.wf; col = (ks * 16 + kk) * 2 + kind; ld = 2 * W; stride = 2; }
            else { src = mt ? L.ws : L.wr; col = ks * 16 + kk; ld = W; }
        } else {
            w -= per_layer * n_layers;
            const int which = (int)(w / (CL * HW));         // 0: end_conv_1   1: end_conv_2
            w -= (size_t)which * CL * HW;
            const int vrank = (int)(w / HW);
            w -= (size_t)vrank * HW;
            const int j = (int)(w & 3), lane = (int)((w >> 2) & 31), ks = (int)(w >> 8);
            const int g = lane >> 2, t = lane & 3;
            row = vrank * 16 + g + 8 * (j & 1);
            col = ks * 16 + 2 * t + 8 * (j >> 1);
            ld = W;
            src = which ? e2w : e1w;
        }
        const int half = (int)((i >> 7) & 1);                // every image is a multiple of 256 words: bit 7 of i is the hi/lo plane
        const float x0 = src[(size_t)row * ld + col], x1 = src[(size_t)row * ld + col + stride];
        unsigned short h0, l0, h1, l1;
        cl8_split(x0, h0, l0);
        cl8_split(x1, h1, l1);
        img[i] = half ? ((unsigned)l1 << 16 | l0) : ((unsigned)h1 << 16 | h0);
    }
}

// CS = CTAs per cluster (16 or 8).  The exchanged vectors always consist of 16 blocks ("virtual ranks" of 16 channels);
// a CTA of a CS-cluster owns VR = 16 / CS consecutive virtual ranks and walks them one after the other in every stage.  At
// most 7 clusters of 16 CTAs are co-resident on a B200 (tools/cluster_occ.cu) but 15 clusters of 8: CS = 8 runs 64 streams
// (8 clusters) in one wave on 64 SMs, at twice the per-CTA work -- the step is bound by the exchange latency, not by it.
template <int CS>
__global__ void __launch_bounds__(GEN_NT + 64 + (CS == 8 ? 32 : 0), 1) gen_kernel_cl8(const GenParams p) {
    extern __shared__ __align__(128) unsigned char smb[];
    constexpr int W = CL8_W, SB = CL8_SB, NV = 16, BLK = CL8_BLK, VEC = CL8_VEC, VR = CL / CS, NVC = NV * VR;
    // exchanged vectors first: same offsets in every CTA (mapa keeps the offset).  Each is 16 blocks of 512 bytes.
    unsigned char* Xcur = smb;                          // [2] layer input h (hi/lo split), by layer parity
    unsigned char* Xz = Xcur + 2 * VEC;                 // [2] gated activation z, by layer parity; Xz[1] doubles as sampling scratch
    unsigned char* Xs = Xz + 2 * VEC;                   // relu(skip sum)           \  Xz[1], Xs, Xy are contiguous: 24 KB that no
    unsigned char* Xy = Xs + VEC;                       // end_conv_1 output        /  peer writes while this CTA samples
    unsigned char* Xl = Xy + VEC;                       // logits, fp32: [virtual rank][stream][16]
    unsigned char* Xold = Xl + VEC;                     // history taps of the coming stage 1 (local)
    unsigned char* stg = Xold + VEC;                    // [2][VR][BLK] this CTA's contribution of a stage, staged for the pusher
    float* part = reinterpret_cast<float*>(stg + 2 * VR * BLK);               // [2][VR][8 warps][32 lanes][4] partial C fragments
    float* hown = part + 2 * VR * 8 * 128;              // [2][VR][16][SB] fp32 layer input at the channels this CTA owns
    constexpr int SLOTB = VR * CL8_IMG2, NSLOT = 4 / VR;         // weight ring: 4 x 32 KB (VR = 1) or 2 x 64 KB (VR = 2)
    constexpr int NTHR = GEN_NT + 64 + (VR == 2 ? 32 : 0);      // workers + producer + pusher (+ second producer for VR = 2)
    unsigned char* wbuf = reinterpret_cast<unsigned char*>(hown + 2 * NVC * SB);         // [NSLOT][SLOTB] weight image ring
    unsigned long long* fullb = reinterpret_cast<unsigned long long*>(wbuf + (size_t)NSLOT * SLOTB);
    unsigned long long* emptyb = fullb + 4;
    unsigned long long* xbar = emptyb + 4;              // [0,1] h by layer parity, [2,3] z by layer parity, [4] skip, [5] y1, [6] logits
    GenLayer* lay_s = reinterpret_cast<GenLayer*>(xbar + 8);
    int* slot_s = reinterpret_cast<int*>(lay_s + p.n_layers);
    int* idx_s = slot_s + p.n_layers;                   // [SB] current class index per stream, [SB] abort flag
    float* logit_s = reinterpret_cast<float*>(Xz + VEC);                      // [SB][W] sampling scratch (aliases Xz[1])
    double* cdf = reinterpret_cast<double*>(Xs);                              // [SB][W]                   (aliases Xs, Xy)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_rank(), cl = blockIdx.x / CS, NS = p.NS, NL = p.n_layers;
    const int v0 = rank * VR;                           // first virtual rank of this CTA
    const int o0 = v0 * NV;                             // first channel this CTA owns in every stage vector

    {
        unsigned* z0 = reinterpret_cast<unsigned*>(smb);
        const int n0 = (int)((reinterpret_cast<unsigned char*>(wbuf) - smb) / 4);
        for (int i = tid; i < n0; i += NTHR) z0[i] = 0u;
        const int* src = reinterpret_cast<const int*>(p.layers);
        int* dst = reinterpret_cast<int*>(lay_s);
        for (int i = tid; i < NL * (int)(sizeof(GenLayer) / sizeof(int)); i += NTHR) dst[i] = src[i];
    }
    if (tid < 2 * SB) idx_s[tid] = (tid < SB && cl * SB + tid < NS) ? p.cur_idx[cl * SB + tid] : 0;
    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(fullb + i, VR == 2 ? 33 : 1); mbar_init(emptyb + i, GEN_WARPS); }
        for (int i = 0; i < 7; ++i) mbar_init(xbar + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < 7; ++i) mbar_expect_tx(xbar + i, VEC);             // arm phase 0 of every exchange barrier
    for (int l = tid; l < NL; l += GEN_NT) {
        const int len = lay_s[l].ring_len;
        slot_s[l] = (p.t0 + len - 1) % len;
    }
    cluster_sync_all();                                  // nobody may store into a peer before its barriers exist
    constexpr unsigned smask = (unsigned)NSLOT - 1u, sshift = (NSLOT == 4) ? 2u : 1u;
    const size_t img_kind = (size_t)CL * CL8_IMG2;       // bytes of one (layer, kind): 16 virtual ranks
    const unsigned char* img_mine = p.cl8_img + (size_t)v0 * CL8_IMG2;
    const unsigned char* img_head = p.cl8_img + 3 * img_kind * NL + (size_t)v0 * CL8_IMGH;

    // push the staged blocks `sb` into blocks v0 .. v0+VR-1 of vector `vec` of every CTA of the cluster: the pusher warp
    // (warp 9) reads each 512-byte block back (16 bytes per lane) and issues ONE st.async.v4 per destination -- a whole
    // block per instruction, its bytes credited to the destination's mbarrier.  Measured per exchange round
    // (tools/dsmem_probe.cu, 8 clusters of 16): 1 017 cycles this way (H), 1 254 with one cp.async.bulk per destination (F,
    // which also occupies the SM's bulk-copy engine and needs a proxy fence after staging), 4 144 with per-lane 8-byte
    // stores from the worker warps (E), and ~1 200 more for a multicast copy from a global staging slot (the fence after the
    // global stores).  The workers only signal "staged" (bar.arrive on barrier 2) and move on to the next stage.
    const unsigned sm_base = smem_u32(smb);
    unsigned rdelta[CS];                                 // shared::cluster address of CTA d minus the local address (pusher warp)
#pragma unroll
    for (int d = 0; d < CS; ++d) rdelta[d] = (warp == GEN_WARPS + 1) ? mapa_u32(sm_base, (unsigned)d) - sm_base : 0u;
    auto push = [&](int sb, unsigned char* vec, int bar_i) {
#pragma unroll
        for (int vr = 0; vr < VR; ++vr) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + (sb * VR + vr) * BLK + lane * 16);
            const unsigned la = smem_u32(vec + (v0 + vr) * BLK + lane * 16), lb = smem_u32(xbar + bar_i);
#pragma unroll
            for (int d = 0; d < CS; ++d)
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                                 la + rdelta[d]),
                             "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(lb + rdelta[d])
                             : "memory");
        }
    };
#define CL8_STAGED_SYNC() asm volatile("bar.sync 2, 288;" ::: "memory")
#define CL8_STAGED_ARRIVE() asm volatile("bar.arrive 2, 288;" ::: "memory")
    if (warp == GEN_WARPS + 1) {
        int sb = 0;
        for (int ev = 0; ev < p.n_evals; ++ev) {
            const bool wh = (p.t0 + ev >= p.n_given - 1);
            for (int l = 0; l < NL; ++l) {
                CL8_STAGED_SYNC();
                push(sb, Xz + (l & 1) * VEC, 2 + (l & 1));
                sb ^= 1;
                if (l + 1 < NL) {
                    CL8_STAGED_SYNC();
                    push(sb, Xcur + ((l + 1) & 1) * VEC, (l + 1) & 1);
                    sb ^= 1;
                }
            }
            if (wh) {
                CL8_STAGED_SYNC(); push(sb, Xs, 4); sb ^= 1;
                CL8_STAGED_SYNC(); push(sb, Xy, 5); sb ^= 1;
                CL8_STAGED_SYNC(); push(sb, Xl, 6); sb ^= 1;
            }
        }
        cluster_sync_all();
        return;
    }
    // ---- producer warp: the weight images of this CTA for every stage, in order: ONE bulk copy per stage kind (the images
    // of the CTA's VR virtual ranks are adjacent).  One lane issues; the issuing thread is held ~100 cycles + bytes/84 per
    // instruction, so whole images it is: 8 KB pieces cap an SM at 45 B/cycle, 32 KB images reach 42 and 64 KB images 84
    // with two in flight (tools/bulk_bw_probe.cu).  The exchange no longer uses the bulk-copy engine (st.async), which had
    // throttled these copies to ~22 B/cycle.
    if (warp == GEN_WARPS) {
        if (lane == 0) {
            unsigned q = 0;
            for (int ev = 0; ev < p.n_evals; ++ev) {
                const bool wh = (p.t0 + ev >= p.n_given - 1);
                const int n_st = wh ? 3 * NL + 2 : 3 * NL;      // per layer: old taps, current input, stage 2; then the two head stages
                for (int st = 0; st < n_st; ++st, ++q) {
                    const unsigned char* src;
                    unsigned bytes;
                    if (st < 3 * NL) {
                        src = img_mine + (size_t)st * img_kind;
                        bytes = VR * CL8_IMG2;
                    } else {
                        src = img_head + (size_t)(st - 3 * NL) * CL * CL8_IMGH;
                        bytes = VR * CL8_IMGH;
                    }
                    const int slot = (int)(q & smask);
                    if (q >= (unsigned)NSLOT) {
                        const unsigned par = ((q >> sshift) & 1u) ^ 1u;
                        unsigned done = 0, spins = 0;
                        while (!done) {
                            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(done) : "r"(smem_u32(emptyb + slot)), "r"(par) : "memory");
                            if (!done && ++spins > (1u << 30)) asm volatile("trap;");
                        }
                    }
                    // VR = 2: the bulk copy brings the first virtual rank's image, warp 10 the second one (below)
                    mbar_expect_tx(fullb + slot, bytes / VR);
                    bulk_g2s(wbuf + (size_t)slot * SLOTB, src, bytes / VR, fullb + slot);
                }
            }
        }
        cluster_sync_all();                              // matches the workers' final cluster barrier
        return;
    }
    // ---- second producer (VR = 2 only): one bulk copy at a time streams ~31 bytes/cycle into the 2-slot ring, which made
    // the 192 KB of a layer the bottleneck; this warp fetches the second virtual rank's image of every stage as 16-byte
    // cp.async copies (load/store path, ~27 bytes/cycle for one warp) in parallel with the bulk copy of the first.
    if (VR == 2 && warp == GEN_WARPS + 2) {
        unsigned q = 0;
        for (int ev = 0; ev < p.n_evals; ++ev) {
            const bool wh = (p.t0 + ev >= p.n_given - 1);
            const int n_st = wh ? 3 * NL + 2 : 3 * NL;
            for (int st = 0; st < n_st; ++st, ++q) {
                const unsigned half = (st < 3 * NL) ? CL8_IMG2 : CL8_IMGH;
                const unsigned char* src = (st < 3 * NL) ? img_mine + (size_t)st * img_kind + half
                                                         : img_head + (size_t)(st - 3 * NL) * CL * CL8_IMGH + half;
                const int slot = (int)(q & smask);
                if (q >= (unsigned)NSLOT) {
                    const unsigned par = ((q >> sshift) & 1u) ^ 1u;
                    unsigned done = 0, spins = 0;
                    while (!done) {
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done) : "r"(smem_u32(emptyb + slot)), "r"(par) : "memory");
                        if (!done && ++spins > (1u << 30)) asm volatile("trap;");
                    }
                }
                const unsigned dst = smem_u32(wbuf + (size_t)slot * SLOTB + half) + lane * 16;
                const unsigned char* sp = src + lane * 16;
#pragma unroll 8
                for (unsigned o = 0; o < half; o += 512)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(sp + o) : "memory");
                // this lane's arrival on the slot's "full" barrier fires when its copies above have landed
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(fullb + slot)) : "memory");
            }
        }
        cluster_sync_all();
        return;
    }
    unsigned cons_q = 0;
    auto stage_weights = [&]() -> const unsigned char* {
        const int slot = (int)(cons_q & smask);
        mbar_wait(fullb + slot, (cons_q >> sshift) & 1u);
        return wbuf + (size_t)slot * SLOTB;
    };
    auto release_slot = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive_(emptyb + (cons_q & smask));
        ++cons_q;
    };
    // exchange barrier i: wait for the phase all threads are at, then thread 0 arms the next phase (it has seen this one end)
    unsigned xpar = 0;                                   // bit i = parity of the phase of xbar[i] to wait for next
    auto xwait = [&](int i) {
        mbar_wait_bounded(xbar + i, (xpar >> i) & 1u);
        xpar ^= 1u << i;
        if (tid == 0) mbar_expect_tx(xbar + i, VEC);
    };
    // one k-step of this warp's m-tile: A fragments (hi, lo) from the stage image, B fragments of the 8 streams from a block
    const int mt = warp & 1, kq = warp >> 1;
    // three independent accumulation chains (lo.hi, hi.lo, hi.hi) so that consecutive MMAs do not wait for each other; they
    // are added in a fixed order when the partial is stored
    struct Acc3 { float lh[4], hl[4], hh[4]; };
    auto acc_zero = [](Acc3& d) {
#pragma unroll
        for (int i = 0; i < 4; ++i) d.lh[i] = d.hl[i] = d.hh[i] = 0.f;
    };
    auto acc_store = [&](const Acc3& d, float* dst) {
        *reinterpret_cast<float4*>(dst) = make_float4((d.lh[0] + d.hl[0]) + d.hh[0], (d.lh[1] + d.hl[1]) + d.hh[1],
                                                      (d.lh[2] + d.hl[2]) + d.hh[2], (d.lh[3] + d.hl[3]) + d.hh[3]);
    };
    auto mma_step = [&](const unsigned char* a, const unsigned char* xblk, Acc3& d) {      // a: this lane's hi fragment; lo at +512
        const uint4 ah = *reinterpret_cast<const uint4*>(a), al = *reinterpret_cast<const uint4*>(a + 512);
        const uint4 b = *reinterpret_cast<const uint4*>(xblk + (lane >> 2) * 64 + (lane & 3) * 16);
        mma_bf16_16816(d.lh, al, b.x, b.z);
        mma_bf16_16816(d.hl, ah, b.y, b.w);
        mma_bf16_16816(d.hh, ah, b.x, b.z);
    };
    // this warp's four k-steps (4*kq .. 4*kq+3) of its m-tile of the 32 KB layer image of every virtual rank (images VR apart
    // by CL8_IMG2) against blocks 4*kq.. of vector x.  The virtual ranks' chains are interleaved k-step by k-step: they are
    // independent, share the B fragments, and hide each other's MMA latency (nothing is stored until all are issued).
    auto mma_quarter = [&](const unsigned char* wimg, const unsigned char* x, Acc3 (&d)[VR]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int vr = 0; vr < VR; ++vr)
                mma_step(wimg + vr * CL8_IMG2 + ((size_t)(mt * 16 + 4 * kq + i) * 2) * 512 + lane * 16, x + (4 * kq + i) * BLK, d[vr]);
    };
    // sum over the 4 K quarters of output (m-tile m, row r of the tile, stream s): fragment element (lane', j) of each partial
    auto part_sum = [&](const float* pb, int m, int r, int s) {
        const float* q = pb + ((m * 32 + (r & 7) * 4 + (s >> 1)) << 2) + ((r >> 3) << 1) + (s & 1);
        return ((q[0] + q[256]) + q[512]) + q[768];
    };
    // history taps: thread -> (stream = warp, channels 2*lane + 64*j + {0,1}), fetched into registers one stage ahead
    const int hs_g = cl * SB + warp;                     // global stream whose taps this thread fetches (and which it samples)
    Pair2 hq[4];
    const uint2* hsrc = nullptr;
    unsigned htag = 0;
    auto issue_old = [&](int ln, int te, int slot_te) {  // slot_te = ring slot of time te in layer ln
        const GenLayer& Lp = lay_s[ln];
        hsrc = nullptr;
        if (te >= Lp.dil && hs_g < NS) {
            const int so = (slot_te + 1 == Lp.ring_len) ? 0 : slot_te + 1;
            hsrc = p.ringLL + Lp.ring_off + ((size_t)so * NS + hs_g) * W + 2 * lane;
            htag = (unsigned)(te - Lp.dil) + 1u;
#pragma unroll
            for (int j = 0; j < 4; ++j) hq[j] = ld_pair2(hsrc + 64 * j);
        }
    };
    auto commit_old = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = 0.f, b = 0.f;
            if (hsrc != nullptr) {
                if (hq[j].a.y != htag || hq[j].b.y != htag) hq[j] = poll2_spin(hsrc + 64 * j, htag, p.err, idx_s + SB);
                a = __uint_as_float(hq[j].a.x);
                b = __uint_as_float(hq[j].b.x);
            }
            const int c0 = 2 * lane + 64 * j;             // channels c0, c0 + 1: one (hi pair | lo pair) unit
            unsigned short h0, l0, h1, l1;
            cl8_split(a, h0, l0);
            cl8_split(b, h1, l1);
            *reinterpret_cast<uint2*>(Xold + (c0 >> 4) * BLK + cl8_unit_off(warp, c0 & 15)) =
                make_uint2((unsigned)h1 << 16 | h0, (unsigned)l1 << 16 | l0);
        }
    };
    issue_old(0, p.t0, p.t0 % lay_s[0].ring_len);
    commit_old();
    WORKER_SYNC();
    // finishing threads: output (virtual rank fvr of this CTA, channel fc of its 16, stream fs).  VR = 2: every thread finishes
    // one conv / residual output AND one skip output; VR = 1: threads 0-127 the conv / residual ones, 128-255 the skip ones.
    const int fvr = (VR == 2) ? (tid >> 7) : 0, fc = (tid >> 3) & 15, fs = tid & 7;
    const bool fin_a = (VR == 2) || tid < NV * SB, fin_s = (VR == 2) || tid >= NV * SB;
    const int fch = o0 + fvr * NV + fc;                   // the channel (= row of the stage's weight matrix) this thread finishes
    const int fsg = cl * SB + fs;
    const bool fs_on = fsg < NS;
    unsigned pb_i = 0, sb_i = 0;                          // partial-sum / staging double-buffer indices
    auto part_of = [&](unsigned pb, int vr) { return part + (pb * VR + vr) * 8 * 128; };
    auto stg_of = [&](unsigned sb, int vr) { return stg + (sb * VR + vr) * BLK; };

    for (int ev = 0; ev < p.n_evals; ++ev) {
        const int t = p.t0 + ev;
        const unsigned rtag = (unsigned)t + 1u;          // ring tag of time t
        const bool want_head = (t >= p.n_given - 1);
        const int samp = t - (p.n_given - 1);
        const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && tid == 0 && ev == p.n_evals - 1;
        int tr_n = 0;
#define TR8() do { if (tr_on && tr_n < 2040) p.trace[tr_n++] = clock64(); } while (0)
        if (tr_on) p.trace[2040] = clock64();             // whole-evaluation stamps live at [2040..2047]
        if (tid < SB && cl * SB + tid < NS) {
            const int g = cl * SB + tid;
            if (t < p.n_given) idx_s[tid] = p.first[(size_t)g * p.n_given + t];
            else if (p.forced != nullptr) idx_s[tid] = p.forced[(size_t)g * p.n_samples + (t - p.n_given)];
        }
        for (int l = tid; l < NL; l += GEN_NT) {
            const int s1 = slot_s[l] + 1;
            slot_s[l] = (s1 == lay_s[l].ring_len) ? 0 : s1;
        }
        WORKER_SYNC();
        // layer 0's input: the start-conv column of every stream (warp = stream), computed locally by every CTA; the
        // owners of a channel also enqueue it in the ring and keep the fp32 value for the residual add
        {
            int idx = idx_s[warp];
            idx = idx < 0 ? 0 : (idx >= W ? W - 1 : idx);
            const GenLayer& L0 = lay_s[0];
            uint2* ring0 = p.ringLL + L0.ring_off + ((size_t)slot_s[0] * NS + hs_g) * W;
#pragma unroll
            for (int j = 0; j < W / 32; ++j) {
                const int r = lane + 32 * j;
                const float v = __ldg(p.start_w + (size_t)r * W + idx) + (p.start_b ? __ldg(p.start_b + r) : 0.f);
                cl8_put(Xcur + (r >> 4) * BLK, warp, r & 15, v);
                if (r >= o0 && r < o0 + NVC) {
                    hown[(r - o0) * SB + warp] = v;
                    if (hs_g < NS) st_pair(ring0 + r, v, rtag);
                }
            }
        }
        WORKER_SYNC();
        float skr = 0.f;                                  // skip sum of (channel fch, stream fs) in the fin_s threads
        TR8();             // layer stamps start here (index 0)

        for (int l = 0; l < NL; ++l) {
            const GenLayer& L = lay_s[l];
            const bool more = (l + 1 < NL);
            unsigned char* xc = Xcur + (l & 1) * VEC;
            unsigned char* zb = Xz + (l & 1) * VEC;
            // biases of this thread's outputs: requested now, used one or two stages later
            const float b_f = (fin_a && L.bf) ? __ldg(L.bf + fch) : 0.f, b_g = (fin_a && L.bg) ? __ldg(L.bg + fch) : 0.f;
            const float b_r = (fin_a && L.br) ? __ldg(L.br + fch) : 0.f, b_s = (fin_s && L.bs) ? __ldg(L.bs + fch) : 0.f;
            // ================= stage 1: m-tile 0 = filter rows, 1 = gate rows; k-steps 4*kq+i of the old taps, then of h
            {
                // history taps of the NEXT stage 1 start their trip through the L2 now; they are used a whole stage later
                if (more) issue_old(l + 1, t, slot_s[l + 1]);
                else if (ev + 1 < p.n_evals) issue_old(0, t + 1, (slot_s[0] + 1 == lay_s[0].ring_len) ? 0 : slot_s[0] + 1);
                else hsrc = nullptr;
                Acc3 d[VR];
                const unsigned char* wimg = stage_weights();
                TR8();         // 1: stage-1 weights (old tap) landed
#pragma unroll
                for (int vr = 0; vr < VR; ++vr) acc_zero(d[vr]);
                mma_quarter(wimg, Xold, d);
                release_slot();
                wimg = stage_weights();
                if (l > 0) xwait(l & 1);
                TR8();         // 2: old-tap MMAs done, h arrived
                mma_quarter(wimg, xc, d);
                release_slot();
#pragma unroll
                for (int vr = 0; vr < VR; ++vr) acc_store(d[vr], part_of(pb_i, vr) + ((kq * 2 + mt) * 32 + lane) * 4);
                TR8();         // 3: MMAs done, partials stored
                WORKER_SYNC();
                TR8();         // 4: barrier
                if (fin_a) {
                    const float* pb = part_of(pb_i, fvr);
                    const float f = part_sum(pb, 0, fc, fs) + b_f;
                    const float g = part_sum(pb, 1, fc, fs) + b_g;
                    cl8_put(stg_of(sb_i, fvr), fs, fc, tanh_(f) * sigmoid_(g));
                }
                pb_i ^= 1;
                TR8();         // 5: z computed and staged
                CL8_STAGED_ARRIVE();
                sb_i ^= 1;
            }
            // ================= stage 2: m-tile 0 = residual rows, 1 = skip rows; k-steps 4*kq+i of z
            {
                // every warp is past its stage-1 reads of Xold (two barriers ago): refill it while z is in flight
                commit_old();
                TR8();         // 6: next history taps in place
                const bool active = (mt == 0) ? more : want_head;
                const unsigned char* wimg = stage_weights();
                TR8();         // 7: stage-2 weights landed
                xwait(2 + (l & 1));
                TR8();         // 8: z arrived
                Acc3 d[VR];
#pragma unroll
                for (int vr = 0; vr < VR; ++vr) acc_zero(d[vr]);
                if (active) mma_quarter(wimg, zb, d);
                release_slot();
#pragma unroll
                for (int vr = 0; vr < VR; ++vr) acc_store(d[vr], part_of(pb_i, vr) + ((kq * 2 + mt) * 32 + lane) * 4);
                TR8();         // 9: MMAs done, partials stored
                WORKER_SYNC();
                TR8();         // 10: barrier
                const float* pb = part_of(pb_i, fvr);
                if (fin_a && more) {
                    const GenLayer& Ln = lay_s[l + 1];
                    float v = part_sum(pb, 0, fc, fs) + b_r;
                    v += hown[(l & 1) * NVC * SB + (fvr * NV + fc) * SB + fs];
                    hown[((l + 1) & 1) * NVC * SB + (fvr * NV + fc) * SB + fs] = v;
                    if (fs_on) st_pair(p.ringLL + Ln.ring_off + ((size_t)slot_s[l + 1] * NS + fsg) * W + fch, v, rtag);
                    cl8_put(stg_of(sb_i, fvr), fs, fc, v);
                }
                if (fin_s && want_head) {
                    const float v = part_sum(pb, 1, fc, fs) + b_s;
                    skr = v + skr;
                }
                pb_i ^= 1;
                TR8();         // 11: h' computed and staged
                if (more) {
                    CL8_STAGED_ARRIVE();
                    sb_i ^= 1;
                }
            }
        }
        if (tr_on) p.trace[2041] = clock64();             // layers done
        if (!want_head) continue;

        // ================= head: relu(skip) -> end_conv_1 -> relu -> end_conv_2; one m-tile per virtual rank, 16 k-steps
        // over the 8 warps (2 each)
        if (fin_s) cl8_put(stg_of(sb_i, fvr), fs, fc, fmaxf(skr, 0.f));
        CL8_STAGED_ARRIVE();
        sb_i ^= 1;
        auto head_stage = [&](const unsigned char* x) {     // this warp's 2 k-steps of each virtual rank's m-tile -> part
            const unsigned char* wimg = stage_weights();
            Acc3 d[VR];
#pragma unroll
            for (int vr = 0; vr < VR; ++vr) acc_zero(d[vr]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int vr = 0; vr < VR; ++vr) {
                    const int ks = 2 * warp + i;
                    mma_step(wimg + vr * CL8_IMGH + ((size_t)ks * 2) * 512 + lane * 16, x + ks * BLK, d[vr]);
                }
            release_slot();
#pragma unroll
            for (int vr = 0; vr < VR; ++vr) acc_store(d[vr], part_of(pb_i, vr) + (warp * 32 + lane) * 4);
        };
        auto head_sum = [&](const float* pb, int r, int s) {  // 8 partials, one per warp
            const float* q = pb + (((r & 7) * 4 + (s >> 1)) << 2) + ((r >> 3) << 1) + (s & 1);
            float v = q[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) v += q[i * 128];
            return v;
        };
        xwait(4);
        head_stage(Xs);
        WORKER_SYNC();
        if (fin_a) {
            const float y = fmaxf(head_sum(part_of(pb_i, fvr), fc, fs) + __ldg(p.e1b + fch), 0.f);
            cl8_put(stg_of(sb_i, fvr), fs, fc, y);
        }
        pb_i ^= 1;
        CL8_STAGED_ARRIVE();
        sb_i ^= 1;
        xwait(5);
        head_stage(Xy);
        WORKER_SYNC();
        if (fin_a) {
            const float dc = (float)fch - (float)W / 2.f;
            const float v = (head_sum(part_of(pb_i, fvr), fc, fs) + __ldg(p.e2b + fch)) - (dc * dc) * p.regularize;
            if (fs_on && p.out_logits) p.out_logits[((size_t)fsg * p.n_samples + samp) * W + fch] = v;
            reinterpret_cast<float*>(stg_of(sb_i, fvr))[fs * NV + fc] = v;               // logits travel as fp32: [stream][16]
        }
        pb_i ^= 1;
        CL8_STAGED_ARRIVE();
        sb_i ^= 1;
        xwait(6);
        if (tr_on) p.trace[2042] = clock64();             // head done, logits everywhere
        // every CTA holds all logits of its 8 streams: warp = stream draws the next index (all CTAs agree)
        if (hs_g < NS) {
            float* lg = logit_s + warp * W;
            const float* xl = reinterpret_cast<const float*>(Xl);
            for (int c = lane; c < W; c += 32) lg[c] = xl[(c >> 4) * (BLK / 4) + warp * NV + (c & 15)];
            __syncwarp();
            const int choice = choose_sample(lg, cdf + warp * W, W, lane, p.temperature,
                                             p.uniforms ? p.uniforms + (size_t)hs_g * p.n_samples + samp : nullptr);
            if (lane == 0) {
                idx_s[warp] = choice;
                if (rank == 0) p.out_idx[(size_t)hs_g * p.n_samples + samp] = choice;
            }
        }
        if (tr_on) p.trace[2043] = clock64();             // sampled
        // the top-of-evaluation barrier publishes idx_s
    }
    WORKER_SYNC();
    if (rank == 0 && tid < SB && cl * SB + tid < NS) p.cur_idx[cl * SB + tid] = idx_s[tid];
    cluster_sync_all();                                  // peers may still be storing into this CTA's shared memory
}

// ------------------------------------------------------------------------------------------------ host side
struct ScratchLayout {
    size_t bar, cur_idx, layers, zbuf, skipbuf, y1buf, logitbuf, err, zLL, skipLL, y1LL, logitLL, ll_end, trace, cl8_img, cl8_bytes, total;
};
// the batched cluster kernel's shape: a k = 2 net whose five widths are all 256 (any number of streams)
static bool cl8_shape_ok(const wn_gen_shape& s) {
    return s.n_streams >= 1 && s.k == 2 && s.n_layers >= 2 && s.R == CL8_W && s.D == CL8_W && s.S == CL8_W && s.E == CL8_W &&
           s.classes == CL8_W;
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static ScratchLayout scratch_layout(const wn_gen_shape& s) {
    ScratchLayout o;
    size_t off = 0;
    o.bar = off; off += 256;
    o.cur_idx = off; off = align_up(off + sizeof(int) * s.n_streams, 256);
    o.layers = off; off = align_up(off + sizeof(GenLayer) * s.n_layers, 256);
    o.zbuf = off; off = align_up(off + sizeof(float) * (size_t)s.n_streams * s.D, 256);
    o.skipbuf = off; off = align_up(off + sizeof(float) * (size_t)s.n_streams * s.S, 256);
    o.y1buf = off; off = align_up(off + sizeof(float) * (size_t)s.n_streams * s.E, 256);
    o.logitbuf = off; off = align_up(off + sizeof(float) * (size_t)s.n_streams * s.classes, 256);
    o.err = off; off += 256;
    o.zLL = off; off = align_up(off + sizeof(uint2) * 2 * (size_t)s.n_layers * s.n_streams * s.D, 256);
    o.skipLL = off; off = align_up(off + sizeof(uint2) * 2 * (size_t)s.n_streams * s.S, 256);
    o.y1LL = off; off = align_up(off + sizeof(uint2) * 2 * (size_t)s.n_streams * s.E, 256);
    o.logitLL = off; off = align_up(off + sizeof(uint2) * 2 * (size_t)s.n_streams * s.classes, 256);
    o.ll_end = off;
    o.trace = off; off += 8 * 2048;
    o.cl8_img = off;
    o.cl8_bytes = cl8_shape_ok(s) ? (size_t)CL * ((size_t)s.n_layers * (CL8_IMG1 + CL8_IMG2) + 2 * CL8_IMGH) : 0;
    off = align_up(off + o.cl8_bytes, 256);
    o.total = off;
    return o;
}
static size_t ring_floats(const wn_gen_shape& s, std::vector<long long>* offs) {
    size_t total = 0;
    for (int l = 0; l < s.n_layers; ++l) {
        if (offs) offs->push_back((long long)total);
        total += (size_t)((s.k - 1) * s.dilations[l] + 1) * s.n_streams * s.R;
    }
    return total;
}

}  // namespace wn

using namespace wn;

struct wn_gen_handle {
    wn_gen_shape shape;
    std::vector<int> dil;
    std::vector<GenLayer> layers;
    GenParams base;
    ScratchLayout lay;
    char* scratch;
    size_t ring_bytes;
    int grid, sm_count;
    size_t smem;
    bool tables_uploaded;
    int cur_t;
    int mode;               // 0 = best kernel for the shape (default), 1 = grid barrier, 2 = generic flag-in-data, 3-5 see wn_gen_set_mode
    size_t smem_ll;
    bool fast_ok;           // single stream, k=2, power-of-two grid, rows per stage divide 8: gen_kernel_fast applies
    size_t smem_fast;
    int n_wslots_fast;
    int xn_fast;
    bool cluster_ok;        // k=2, 256-class nets whose rows split over 16 CTAs x 8 warps: gen_kernel_cluster applies
    size_t smem_cluster;
    int n_wslots_cluster, wslot_cluster;
    bool cl8_ok, cl8_packed;   // gen_kernel_cl8 applies (cl8_shape_ok and the shared memory fits); its weight images are built
    bool generic_ok, ll_ok;    // the grid-barrier / the generic flag-exchange kernel fit in shared memory for this stream count
    bool cl8_8_ok;             // ... and so does its 8-CTA-cluster instantiation
    int cl8_cs;                // cluster size picked at the first launch (0 = not yet)
    size_t smem_cl8, smem_cl8_8;
    bool x2_ok;             // fast_ok on a 64-CTA grid with 4 rows per stage vector per CTA: gen_kernel_x2 applies
    size_t smem_x2;
    int n_wslots_x2;
};

static int validate_shape(const wn_gen_shape* s) {
    WN_REQUIRE(s && s->dilations, WN_E_BADARG, "wn_gen: null shape");
    WN_REQUIRE(s->n_layers > 0 && s->k >= 1 && s->R > 0 && s->D > 0 && s->S > 0 && s->E > 0 && s->classes > 0 &&
                   s->n_streams > 0,
               WN_E_BADARG, "wn_gen: bad shape");
    for (int l = 0; l < s->n_layers; ++l) WN_REQUIRE(s->dilations[l] >= 1, WN_E_BADARG, "wn_gen: bad dilation");
    return 0;
}

extern "C" int wn_gen_workspace_bytes(const wn_gen_shape* s, size_t* ring_bytes, size_t* scratch_bytes) {
    if (int rc = validate_shape(s)) return rc;
    if (ring_bytes) *ring_bytes = sizeof(uint2) * ring_floats(*s, nullptr);   // {value, tag} pairs
    if (scratch_bytes) *scratch_bytes = scratch_layout(*s).total;
    return 0;
}

extern "C" int wn_gen_create(const wn_gen_shape* s, const wn_gen_weights* w, float* d_rings, void* d_scratch,
                             wn_gen_handle** out) {
    if (int rc = validate_shape(s)) return rc;
    WN_REQUIRE(w && d_rings && d_scratch && out, WN_E_BADARG, "wn_gen_create: null pointer");
    WN_REQUIRE(w->d_start_w && w->d_wf && w->d_wg && w->d_wr && w->d_ws && w->d_end1_w && w->d_end1_b && w->d_end2_w &&
                   w->d_end2_b,
               WN_E_BADARG, "wn_gen_create: null weight pointer");
    int dev = 0, sms = 0, smem_optin = 0;
    WN_CUDA(cudaGetDevice(&dev));
    WN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WN_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    wn_gen_handle* h = new (std::nothrow) wn_gen_handle();
    WN_REQUIRE(h, WN_E_BADARG, "wn_gen_create: out of host memory");
    h->shape = *s;
    h->dil.assign(s->dilations, s->dilations + s->n_layers);
    h->shape.dilations = h->dil.data();
    h->lay = scratch_layout(h->shape);
    h->scratch = (char*)d_scratch;
    std::vector<long long> offs;
    h->ring_bytes = sizeof(uint2) * ring_floats(h->shape, &offs);
    h->layers.resize(s->n_layers);
    for (int l = 0; l < s->n_layers; ++l) {
        GenLayer& L = h->layers[l];
        L.wf = w->d_wf[l]; L.wg = w->d_wg[l]; L.wr = w->d_wr[l]; L.ws = w->d_ws[l];
        L.bf = w->d_bf ? w->d_bf[l] : nullptr; L.bg = w->d_bg ? w->d_bg[l] : nullptr;
        L.br = w->d_br ? w->d_br[l] : nullptr; L.bs = w->d_bs ? w->d_bs[l] : nullptr;
        if (!(L.wf && L.wg && L.wr && L.ws)) {
            delete h;
            return set_err(WN_E_BADARG, "wn_gen_create: null weight pointer in layer %d", l);
        }
        L.ring_off = offs[l];
        L.dil = s->dilations[l];
        L.ring_len = (s->k - 1) * s->dilations[l] + 1;
    }
    GenParams& p = h->base;
    p.layers = reinterpret_cast<const GenLayer*>(h->scratch + h->lay.layers);
    p.n_layers = s->n_layers; p.k = s->k; p.R = s->R; p.D = s->D; p.S = s->S; p.E = s->E; p.C = s->classes;
    p.NS = s->n_streams;
    p.start_w = w->d_start_w; p.start_b = w->d_start_b;
    p.e1w = w->d_end1_w; p.e1b = w->d_end1_b; p.e2w = w->d_end2_w; p.e2b = w->d_end2_b;
    p.rings = d_rings;
    p.zbuf = reinterpret_cast<float*>(h->scratch + h->lay.zbuf);
    p.skipbuf = reinterpret_cast<float*>(h->scratch + h->lay.skipbuf);
    p.y1buf = reinterpret_cast<float*>(h->scratch + h->lay.y1buf);
    p.logitbuf = reinterpret_cast<float*>(h->scratch + h->lay.logitbuf);
    p.cur_idx = reinterpret_cast<int*>(h->scratch + h->lay.cur_idx);
    p.bar = reinterpret_cast<unsigned*>(h->scratch + h->lay.bar);

    // grid: as many CTAs as keep the per-stage row count per CTA minimal, at most one per SM
    // grid: a power of two, at most one CTA per SM, and -- when the net is wide enough -- at least 4 rows of every
    // exchanged vector per CTA so that a CTA's published pairs fill whole 32-byte sectors (see own_per above)
    int mind = s->D;
    if (s->R < mind) mind = s->R;
    if (s->E < mind) mind = s->E;
    if (s->classes < mind) mind = s->classes;
    int G = 1;
    while (G * 2 <= sms && G * 2 * 4 <= mind) G *= 2;
    if (G == 1)
        while (G * 2 <= sms && G * 2 <= s->D) G *= 2;
    if (const char* e = getenv("WN_GEN_GRID")) {            // tuning knob: fewer, fatter CTAs
        const int v = atoi(e);
        if (v >= 1 && v <= G) G = v;
    }
    if (G < 1) G = 1;
    h->grid = G;
    h->sm_count = sms;
    const int NS = s->n_streams;
    auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
    const int mx1 = (s->k * s->R > s->S) ? s->k * s->R : s->S;
    const int mx2 = (s->D > s->E) ? s->D : s->E;
    p.regA = NS * mx1;
    p.regB = NS * mx2;
    int items = 2 * cdiv(s->D, G);
    p.pre_n = items * NS;
    p.skacc_n = cdiv(s->S, G) * NS;
    if (p.skacc_n < 4) p.skacc_n = 4;
    p.regA = (p.regA + 3) / 4 * 4; p.regB = (p.regB + 3) / 4 * 4; p.pre_n = (p.pre_n + 3) / 4 * 4;
    p.skacc_n = (p.skacc_n + 3) / 4 * 4;
    h->smem = sizeof(float) * ((size_t)p.regA + p.regB + p.pre_n + p.skacc_n + NS + (size_t)GEN_WARPS * s->classes);
    // the grid-barrier / generic kernels stage all streams' vectors in every CTA; the cluster kernels do not
    h->generic_ok = h->smem <= (size_t)smem_optin;
    const bool cluster_shape = s->k == 2 && s->n_layers >= 2 && s->D % CL == 0 && s->R % CL == 0 && s->S % CL == 0 &&
                               s->E % CL == 0 && s->classes % CL == 0;
    if (!h->generic_ok && !cluster_shape) {
        const size_t need = h->smem;
        delete h;
        return set_err(WN_E_UNSUPP, "wn_gen_create: %zu bytes of shared memory needed for %d streams, %d available", need,
                       NS, smem_optin);
    }
    // ---- LL kernel: exchange regions, shared-memory carve, weight prefetch ring
    p.ringLL = reinterpret_cast<uint2*>(d_rings);
    p.zLL = reinterpret_cast<uint2*>(h->scratch + h->lay.zLL);
    p.skipLL = reinterpret_cast<uint2*>(h->scratch + h->lay.skipLL);
    p.y1LL = reinterpret_cast<uint2*>(h->scratch + h->lay.y1LL);
    p.logitLL = reinterpret_cast<uint2*>(h->scratch + h->lay.logitLL);
    p.err = reinterpret_cast<int*>(h->scratch + h->lay.err);
    p.trace = getenv("WN_GEN_TRACE") ? reinterpret_cast<long long*>(h->scratch + h->lay.trace) : nullptr;
    {
        const int nDm = cdiv(s->D, G), nRm = cdiv(s->R, G), nSm = cdiv(s->S, G), nEm = cdiv(s->E, G), nCm = cdiv(s->classes, G);
        int mxA = mx1 > s->classes ? mx1 : s->classes;
        const int regA_ll = (NS * mxA + 3) / 4 * 4;
        int items_max = 2 * nDm;
        if (nRm + nSm > items_max) items_max = nRm + nSm;
        if (nEm > items_max) items_max = nEm;
        if (nCm > items_max) items_max = nCm;
        p.part_n = ((items_max > GEN_WARPS ? items_max : GEN_WARPS) * NS + 3) / 4 * 4;
        long long slot = (long long)2 * nDm * s->k * s->R;
        if ((long long)(nRm + nSm) * s->D > slot) slot = (long long)(nRm + nSm) * s->D;
        if ((long long)nEm * s->S > slot) slot = (long long)nEm * s->S;
        if ((long long)nCm * s->E > slot) slot = (long long)nCm * s->E;
        slot = (slot + 3) / 4 * 4;
        const size_t base = sizeof(float) * ((size_t)regA_ll + p.regB + p.part_n + p.skacc_n + (size_t)GEN_WARPS * s->classes) +
                            sizeof(double) * (size_t)GEN_WARPS * s->classes + 64 + sizeof(GenLayer) * (size_t)s->n_layers +
                            sizeof(int) * (size_t)(NS + 4);
        const bool k_ok = ((s->k * s->R) % 4 == 0) && (s->D % 4 == 0) && (s->S % 4 == 0) && (s->E % 4 == 0);
        int nslots = 0;
        if (k_ok && base < (size_t)smem_optin) {
            long long fit = ((long long)smem_optin - (long long)base) / (slot * 4);
            nslots = fit >= 4 ? 4 : (fit >= 2 ? (int)fit : 0);
        }
        if (getenv("WN_GEN_NOPREFETCH")) nslots = 0;        // tuning knob: read weights through L2 instead
        p.wslot_floats = (int)slot;
        p.n_wslots = nslots;
        h->smem_ll = base + (size_t)nslots * slot * 4;
        // regA of the LL kernel also stages the logits; keep one GenParams for both kernels
        if (regA_ll > p.regA) {
            p.regA = regA_ll;
            h->smem = sizeof(float) * ((size_t)p.regA + p.regB + p.pre_n + p.skacc_n + NS + (size_t)GEN_WARPS * s->classes);
        }
        h->mode = (s->n_layers >= 2 && h->smem_ll <= (size_t)smem_optin) ? 0 : 1;
        // ---- fast kernel eligibility (same K split as the generic kernel so both sum in the same order)
        auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
        auto split_ok = [&](int rows, int K) {
            if (!(rows == 1 || rows == 2 || rows == 4 || rows == 8)) return false;
            const int hs = GEN_WARPS / rows;
            return K % hs == 0 && (K / hs) % 4 == 0 && (K / hs) >= 32 && (K / hs) <= FAST_MAXI * 128;
        };
        bool ok = NS == 1 && s->k == 2 && pow2(G) && G >= 2 && s->n_layers >= 2 && s->D % G == 0 && s->R % G == 0 &&
                  s->S % G == 0 && s->E % G == 0 && s->classes % G == 0;
        if (ok)
            ok = split_ok(2 * (s->D / G), 2 * s->R) && split_ok((s->R + s->S) / G, s->D) && split_ok(s->E / G, s->S) &&
                 split_ok(s->classes / G, s->E);
        h->fast_ok = false;
        if (ok) {
            int xn = s->R;
            if (s->D > xn) xn = s->D;
            if (s->S > xn) xn = s->S;
            if (s->E > xn) xn = s->E;
            if (s->classes > xn) xn = s->classes;
            xn = (xn + 3) & ~3;
            h->xn_fast = xn;
            const size_t fbase = sizeof(float) * (2 * GEN_WARPS + ((s->S / G + 3) & ~3) + 2 * ((s->R / G + 3) & ~3) + 2 * xn) +
                                 sizeof(double) * s->classes + 64 + sizeof(GenLayer) * (size_t)s->n_layers +
                                 sizeof(uint2) * 2 * (size_t)s->R + sizeof(int) * (size_t)(s->n_layers + 4);
            int fs = 0;
            if (fbase < (size_t)smem_optin) {
                long long fit = ((long long)smem_optin - (long long)fbase) / (slot * 4);
                fs = fit >= 4 ? 4 : (fit >= 2 ? 2 : 0);
            }
            if (getenv("WN_GEN_NOPREFETCH")) fs = 0;
            h->n_wslots_fast = fs;
            h->smem_fast = fbase + (size_t)fs * slot * 4;
            h->fast_ok = h->smem_fast <= (size_t)smem_optin;
        }
        // ---- two-level exchange kernel: the fast kernel's grid as clusters of 16, exactly 4 values per CTA per stage vector
        h->x2_ok = false;
        if (ok && G % CL == 0 && G / CL >= 1 && G / CL <= 5 && s->D / G == 4 && s->R / G == 4 && s->S / G == 4 &&
            s->E / G == 4 && s->classes / G == 4 && !getenv("WN_GEN_NOX2")) {
            const size_t xbase = sizeof(uint2) * (size_t)(2 * s->R + 2 * s->D + s->S + s->E + s->classes + 2 * s->R) +
                                 sizeof(float) * (size_t)(2 * GEN_WARPS + 4 + ((s->classes + 3) & ~3)) +
                                 sizeof(double) * s->classes + 64 + sizeof(GenLayer) * (size_t)s->n_layers +
                                 sizeof(int) * (size_t)(s->n_layers + 4);
            int xs = 0;
            if (xbase < (size_t)smem_optin) {
                long long fit = ((long long)smem_optin - (long long)xbase) / (slot * 4);
                xs = fit >= 4 ? 4 : (fit >= 2 ? 2 : 0);
            }
            h->n_wslots_x2 = xs;
            h->smem_x2 = xbase + (size_t)xs * slot * 4;
            h->x2_ok = xs >= 2 && h->smem_x2 <= (size_t)smem_optin;
        }
    }
    // ---- cluster kernel eligibility: rows of every stage split evenly over 16 CTAs x 8 warps, <= 4 rows per warp
    {
        auto div_ok = [&](int N) { return N % CL == 0; };
        bool ok = s->k == 2 && s->n_layers >= 2 && div_ok(s->D) && div_ok(s->R) && div_ok(s->S) && div_ok(s->E) &&
                  div_ok(s->classes) && (s->R % 4 == 0) && (s->D % 4 == 0) && (s->S % 4 == 0) && (s->E % 4 == 0);
        if (ok) {
            const int nD = s->D / CL, nR = s->R / CL, nS = s->S / CL;
            const int rw1 = 2 * nD / GEN_WARPS, rw2 = (nR + nS) / GEN_WARPS;
            ok = (2 * nD) % GEN_WARPS == 0 && (nR + nS) % GEN_WARPS == 0 && (rw1 == 2 || rw1 == 4) && rw2 >= 1 && rw2 <= CL_ROWS &&
                 nR % rw2 == 0 && (rw2 == 2 || rw2 == 4 || rw2 == 1);
        }
        h->cluster_ok = false;
        if (ok && !getenv("WN_GEN_NOCLUSTER")) {
            long long slot = (long long)2 * (s->D / CL) * 2 * s->R;
            if ((long long)(s->R / CL + s->S / CL) * s->D > slot) slot = (long long)(s->R / CL + s->S / CL) * s->D;
            if ((long long)(s->E / CL) * s->S > slot) slot = (long long)(s->E / CL) * s->S;
            if ((long long)(s->classes / CL) * s->E > slot) slot = (long long)(s->classes / CL) * s->E;
            slot = (slot + 3) / 4 * 4;
            const size_t cbase = sizeof(uint2) * (size_t)(2 * s->R + 2 * s->D + s->S + s->E + s->classes + 2 * s->R) +
                                 sizeof(float) * (size_t)(((s->S / CL + 3) & ~3) + 2 * ((s->R / CL + 3) & ~3) +
                                                          (((s->E + s->classes) / CL + 32 + 3) & ~3) + ((s->classes + 3) & ~3)) +
                                 sizeof(double) * s->classes + 64 + sizeof(GenLayer) * (size_t)s->n_layers +
                                 sizeof(int) * (size_t)(s->n_layers + 4);
            int cs = 0;
            if (cbase < (size_t)smem_optin) {
                long long fit = ((long long)smem_optin - (long long)cbase) / (slot * 4);
                cs = fit >= 4 ? 4 : (fit >= 2 ? 2 : 0);
            }
            h->n_wslots_cluster = cs;
            h->wslot_cluster = (int)slot;
            h->smem_cluster = cbase + (size_t)cs * slot * 4;
            h->cluster_ok = cs >= 2 && h->smem_cluster <= (size_t)smem_optin;
        }
    }
    // ---- batched cluster kernel (8 streams per cluster, tensor cores)
    {
        auto smem_for = [&](int vr) {
            const size_t fixed = (size_t)8 * CL8_VEC + (size_t)2 * vr * CL8_BLK + sizeof(float) * (size_t)(2 * vr * 8 * 128 + 2 * 16 * vr * CL8_SB) +
                                 16 * 8 + sizeof(GenLayer) * (size_t)s->n_layers + sizeof(int) * (size_t)(s->n_layers + 2 * CL8_SB);
            return align_up(fixed, 16) + 4 * (size_t)CL8_IMG2;
        };
        h->smem_cl8 = smem_for(1);
        h->smem_cl8_8 = smem_for(2);
        h->cl8_ok = cl8_shape_ok(*s) && h->smem_cl8 <= (size_t)smem_optin && !getenv("WN_GEN_NOCL8");
        h->cl8_8_ok = h->cl8_ok && h->smem_cl8_8 <= (size_t)smem_optin;
        h->cl8_cs = 0;
        h->cl8_packed = false;
        p.cl8_img = reinterpret_cast<const unsigned char*>(h->scratch + h->lay.cl8_img);
    }
    h->generic_ok = h->smem <= (size_t)smem_optin;
    h->ll_ok = h->smem_ll <= (size_t)smem_optin && s->n_layers >= 2;
    if (h->mode == 1 && (h->cl8_ok || h->cluster_ok)) h->mode = 0;       // many streams: only the cluster kernels fit
    if (!h->generic_ok && !h->cl8_ok && !h->cluster_ok) {
        const size_t need = h->smem > h->smem_ll ? h->smem : h->smem_ll;
        delete h;
        return set_err(WN_E_UNSUPP, "wn_gen_create: %zu bytes of shared memory needed for %d streams, %d available", need,
                       s->n_streams, smem_optin);
    }
    h->tables_uploaded = false;
    h->cur_t = 0;
    *out = h;
    return 0;
}

extern "C" int wn_gen_reset(wn_gen_handle* h, void* stream) {
    WN_REQUIRE(h, WN_E_STATE, "wn_gen_reset: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    WN_CUDA(cudaMemsetAsync(h->base.rings, 0, h->ring_bytes, st));
    WN_CUDA(cudaMemsetAsync(h->scratch + h->lay.cur_idx, 0, sizeof(int) * h->shape.n_streams, st));
    WN_CUDA(cudaMemsetAsync(h->scratch + h->lay.err, 0, h->lay.ll_end - h->lay.err, st));
    if (!h->tables_uploaded) {
        WN_CUDA(cudaMemcpyAsync(h->scratch + h->lay.layers, h->layers.data(), sizeof(GenLayer) * h->layers.size(),
                                cudaMemcpyHostToDevice, st));
        WN_CUDA(cudaStreamSynchronize(st));       // h->layers is pageable host memory owned by the handle
        h->tables_uploaded = true;
    }
    if (h->cl8_ok && !h->cl8_packed) {               // weights are constant for the life of a handle: split them once
        cl8_pack_kernel<<<1184, 256, 0, st>>>(h->base.layers, h->shape.n_layers, h->base.e1w, h->base.e2w,
                                             reinterpret_cast<unsigned*>(h->scratch + h->lay.cl8_img));
        WN_CUDA(cudaGetLastError());
        h->cl8_packed = true;
    }
    h->cur_t = 0;
    return 0;
}

template <int SB>
static int launch_gen(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    WN_CUDA(cudaFuncSetAttribute(gen_kernel<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    int per_sm = 0;
    WN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gen_kernel<SB>, GEN_NT, h->smem));
    WN_REQUIRE(per_sm * h->sm_count >= h->grid, WN_E_UNSUPP, "wn_gen_run: %d CTAs cannot be co-resident", h->grid);
    void* args[] = {(void*)&p};
    WN_CUDA(cudaLaunchCooperativeKernel((const void*)gen_kernel<SB>, dim3(h->grid), dim3(GEN_NT), args, h->smem, st));
    return 0;
}

template <int SB, bool PF>
static int launch_gen_ll(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_ll<SB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ll));
    int per_sm = 0;
    WN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gen_kernel_ll<SB, PF>, GEN_NT, h->smem_ll));
    WN_REQUIRE(per_sm * h->sm_count >= h->grid, WN_E_UNSUPP, "wn_gen_run: %d CTAs cannot be co-resident", h->grid);
    void* args[] = {(void*)&p};
    WN_CUDA(cudaLaunchCooperativeKernel((const void*)gen_kernel_ll<SB, PF>, dim3(h->grid), dim3(GEN_NT), args, h->smem_ll, st));
    return 0;
}

static int launch_gen_cluster(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_cluster));
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(h->shape.n_streams * CL));
    cfg.blockDim = dim3(GEN_NT + 32);
    cfg.dynamicSmemBytes = h->smem_cluster;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    WN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, gen_kernel_cluster, &cfg));
    WN_REQUIRE(max_clusters >= 1, WN_E_UNSUPP, "wn_gen_run: a %d-CTA cluster cannot be scheduled on this device", CL);
    WN_CUDA(cudaLaunchKernelEx(&cfg, gen_kernel_cluster, p));
    return 0;
}

template <int CS>
static int launch_gen_cl8_cs(wn_gen_handle* h, GenParams& p, cudaStream_t st, int* max_clusters_out, bool launch) {
    const size_t smem = (CS == 16) ? h->smem_cl8 : h->smem_cl8_8;
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_cl8<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_cl8<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((h->shape.n_streams + CL8_SB - 1) / CL8_SB * CS));
    cfg.blockDim = dim3(GEN_NT + 64 + (CS == 8 ? 32 : 0));    // 8 worker warps, the weight producer warp(s), the pusher warp
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    WN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, gen_kernel_cl8<CS>, &cfg));
    if (max_clusters_out) *max_clusters_out = max_clusters;
    if (!launch) return 0;
    WN_REQUIRE(max_clusters >= 1, WN_E_UNSUPP, "wn_gen_run: a %d-CTA cluster cannot be scheduled on this device", CS);
    WN_CUDA(cudaLaunchKernelEx(&cfg, gen_kernel_cl8<CS>, p));   // clusters are independent: more than fit run in waves
    return 0;
}
// Cluster size: 16 CTAs (least work per CTA) while all clusters are co-resident, else 8 (15 clusters fit instead of 7).
static int launch_gen_cl8(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    if (h->cl8_cs == 0) {
        int fit16 = 0;
        const int rc = launch_gen_cl8_cs<16>(h, p, st, &fit16, false);
        if (rc) return rc;
        const int need = (h->shape.n_streams + CL8_SB - 1) / CL8_SB;
        h->cl8_cs = (need <= fit16 || !h->cl8_8_ok) ? 16 : 8;
        if (const char* e = getenv("WN_GEN_CL8_CS")) {
            const int v = atoi(e);
            if (v == 16 || (v == 8 && h->cl8_8_ok)) h->cl8_cs = v;
        }
    }
    return h->cl8_cs == 16 ? launch_gen_cl8_cs<16>(h, p, st, nullptr, true) : launch_gen_cl8_cs<8>(h, p, st, nullptr, true);
}

// 64 CTAs as 4 clusters of 16, all co-resident (the clusters exchange through the L2 while they run): launched with the
// cluster dimension AND the cooperative attribute, which makes the driver refuse the launch unless every CTA fits at once.
static int launch_gen_x2(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_x2));
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_x2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)h->grid);
    cfg.blockDim = dim3(GEN_NT + 32);
    cfg.dynamicSmemBytes = h->smem_x2;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    WN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, gen_kernel_x2, &cfg));
    WN_REQUIRE(max_clusters >= h->grid / CL, WN_E_UNSUPP, "wn_gen_run: %d clusters of %d CTAs cannot be co-resident (%d fit)",
               h->grid / CL, CL, max_clusters);
    cfg.numAttrs = 2;
    WN_CUDA(cudaLaunchKernelEx(&cfg, gen_kernel_x2, p));
    return 0;
}

template <bool PF, bool TRACE>
static int launch_gen_fast_t(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    WN_CUDA(cudaFuncSetAttribute(gen_kernel_fast<PF, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_fast));
    int per_sm = 0;
    WN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gen_kernel_fast<PF, TRACE>, GEN_NT + 32, h->smem_fast));
    WN_REQUIRE(per_sm * h->sm_count >= h->grid, WN_E_UNSUPP, "wn_gen_run: %d CTAs cannot be co-resident", h->grid);
    void* args[] = {(void*)&p};
    WN_CUDA(cudaLaunchCooperativeKernel((const void*)gen_kernel_fast<PF, TRACE>, dim3(h->grid), dim3(GEN_NT + 32), args,
                                        h->smem_fast, st));
    return 0;
}
template <bool PF>
static int launch_gen_fast(wn_gen_handle* h, GenParams& p, cudaStream_t st) {
    // the stamping variant is a separate instantiation so that the production kernel's hot loop carries no trace code
    return p.trace ? launch_gen_fast_t<PF, true>(h, p, st) : launch_gen_fast_t<PF, false>(h, p, st);
}

extern "C" int wn_gen_set_mode(wn_gen_handle* h, int mode) {
    WN_REQUIRE(h, WN_E_STATE, "wn_gen_set_mode: null handle");
    WN_REQUIRE(mode >= 0 && mode <= 6, WN_E_BADARG,
               "wn_gen_set_mode: mode must be 0 (auto), 1 (grid barrier), 2 (generic flag exchange), 3 (single-stream L2 kernel), "
               "4 (cluster / DSMEM kernel), 5 (single-stream two-level exchange kernel) or 6 (batched tensor-core cluster kernel)");
    if (mode == 3) WN_REQUIRE(h->fast_ok, WN_E_UNSUPP, "wn_gen_set_mode: the single-stream L2 kernel does not apply to this shape");
    if (mode == 4) WN_REQUIRE(h->cluster_ok, WN_E_UNSUPP, "wn_gen_set_mode: the cluster kernel does not apply to this shape");
    if (mode == 5) WN_REQUIRE(h->x2_ok, WN_E_UNSUPP, "wn_gen_set_mode: the two-level exchange kernel does not apply to this shape");
    if (mode == 6) WN_REQUIRE(h->cl8_ok, WN_E_UNSUPP, "wn_gen_set_mode: the batched cluster kernel does not apply to this shape");
    WN_REQUIRE(h->cur_t == 0, WN_E_STATE, "wn_gen_set_mode: switch kernels only right after wn_gen_reset");
    if (mode != 1) WN_REQUIRE(h->shape.n_layers >= 2, WN_E_UNSUPP, "wn_gen_set_mode: flag exchange needs >= 2 layers");
    h->mode = mode;
    return 0;
}

extern "C" int wn_gen_weights_changed(wn_gen_handle* h) {
    WN_REQUIRE(h, WN_E_STATE, "wn_gen_weights_changed: null handle");
    h->cl8_packed = false;          // the next wn_gen_reset splits the weights again
    return 0;
}

extern "C" int wn_gen_check(wn_gen_handle* h, void* stream) {
    WN_REQUIRE(h, WN_E_STATE, "wn_gen_check: null handle");
    int flag = 0;
    WN_CUDA(cudaMemcpyAsync(&flag, h->scratch + h->lay.err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    WN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    WN_REQUIRE(flag == 0, WN_E_STATE, "wn_gen_check: the sampler timed out waiting for an exchange tag (launch aborted)");
    return 0;
}

extern "C" int wn_gen_run(wn_gen_handle* h, const wn_gen_run_args* a, void* stream) {
    WN_REQUIRE(h, WN_E_STATE, "wn_gen_run: null handle");
    WN_REQUIRE(h->tables_uploaded, WN_E_STATE, "wn_gen_run: call wn_gen_reset first");
    WN_REQUIRE(a && a->d_first && a->d_out_idx, WN_E_BADARG, "wn_gen_run: null pointer");
    WN_REQUIRE(a->n_given >= 1 && a->n_samples >= 0 && a->n_evals >= 0 && a->t0 >= 0, WN_E_BADARG, "wn_gen_run: bad counts");
    WN_REQUIRE(a->t0 == h->cur_t, WN_E_STATE, "wn_gen_run: t0=%d does not continue the previous call (expected %d)", a->t0,
               h->cur_t);
    WN_REQUIRE(a->t0 + a->n_evals <= a->n_given - 1 + a->n_samples, WN_E_BADARG,
               "wn_gen_run: evaluations [%d,%d) exceed the schedule of %d given + %d samples", a->t0, a->t0 + a->n_evals,
               a->n_given, a->n_samples);
    WN_REQUIRE(!(a->temperature > 0.f) || a->d_uniforms, WN_E_BADARG, "wn_gen_run: temperature > 0 needs d_uniforms");
    if (a->n_evals == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int bars_per_eval = 2 * h->shape.n_layers + 2;
    int done = 0;
    while (done < a->n_evals) {
        // keep the barrier counter below 2^31: chunk very long runs
        long long max_evals = (long long)0x7fffffff / ((long long)bars_per_eval * h->grid);
        int n = a->n_evals - done;
        if ((long long)n > max_evals) n = (int)max_evals;
        GenParams p = h->base;
        p.first = a->d_first; p.n_given = a->n_given; p.forced = a->d_forced; p.uniforms = a->d_uniforms;
        p.out_idx = a->d_out_idx; p.out_logits = a->d_out_logits; p.n_samples = a->n_samples;
        p.t0 = a->t0 + done; p.n_evals = n; p.temperature = a->temperature; p.regularize = a->regularize;
        WN_CUDA(cudaMemsetAsync(p.bar, 0, sizeof(unsigned), st));
        int rc;
        const bool auto_cluster = h->mode == 0 && h->cluster_ok && (h->shape.n_streams > 1 || !h->fast_ok);
        if ((h->mode == 0 || h->mode == 6) && h->cl8_ok) {       // several streams of a 256-wide net: 8 streams per cluster
            rc = launch_gen_cl8(h, p, st);
        } else if ((auto_cluster || h->mode == 4) && h->cluster_ok) {
            p.n_wslots = h->n_wslots_cluster;
            p.wslot_floats = h->wslot_cluster;
            rc = launch_gen_cluster(h, p, st);
        } else if (h->mode == 5 && h->x2_ok) {                // never picked automatically: measured 2x slower than kernel 3
            p.n_wslots = h->n_wslots_x2;
            rc = launch_gen_x2(h, p, st);
        } else if ((h->mode == 0 || h->mode == 3) && h->fast_ok) {
            p.n_wslots = h->n_wslots_fast;
            p.regA = h->xn_fast;                    // the fast kernel reads its input-vector pitch from regA
            rc = p.n_wslots ? launch_gen_fast<true>(h, p, st) : launch_gen_fast<false>(h, p, st);
        } else if (!(((h->mode == 0 || h->mode == 2) && h->ll_ok) || h->generic_ok)) {
            return set_err(WN_E_UNSUPP, "wn_gen_run: %d streams need a cluster kernel (modes 4, 6) for this net; mode %d does not fit in "
                           "shared memory", h->shape.n_streams, h->mode);
        } else if ((h->mode == 0 || h->mode == 2) && h->ll_ok) {
            if (h->shape.n_streams == 1)
                rc = p.n_wslots ? launch_gen_ll<1, true>(h, p, st) : launch_gen_ll<1, false>(h, p, st);
            else
                rc = p.n_wslots ? launch_gen_ll<8, true>(h, p, st) : launch_gen_ll<8, false>(h, p, st);
        } else {
            rc = (h->shape.n_streams == 1) ? launch_gen<1>(h, p, st) : launch_gen<8>(h, p, st);
        }
        if (rc) return rc;
        done += n;
    }
    h->cur_t = a->t0 + a->n_evals;
    return 0;
}

extern "C" int wn_gen_read_trace(wn_gen_handle* h, long long* host_out, int n, void* stream) {
    WN_REQUIRE(h && host_out && n > 0 && n <= 2048, WN_E_BADARG, "wn_gen_read_trace: bad arguments");
    WN_REQUIRE(h->base.trace, WN_E_STATE, "wn_gen_read_trace: tracing is off (set WN_GEN_TRACE=1 before wn_gen_create)");
    WN_CUDA(cudaMemcpyAsync(host_out, h->base.trace, sizeof(long long) * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    WN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

extern "C" int wn_gen_destroy(wn_gen_handle* h) {
    delete h;
    return 0;
}

/* the kernel wn_gen_run picks in the handle's current mode: the mode number (1-6) of wn_gen_set_mode */
extern "C" int wn_gen_kernel_id(const wn_gen_handle* h) {
    if (!h) return 0;
    const bool auto_cluster = h->mode == 0 && h->cluster_ok && (h->shape.n_streams > 1 || !h->fast_ok);
    if ((h->mode == 0 || h->mode == 6) && h->cl8_ok) return 6;
    if ((auto_cluster || h->mode == 4) && h->cluster_ok) return 4;
    if (h->mode == 5 && h->x2_ok) return 5;
    if ((h->mode == 0 || h->mode == 3) && h->fast_ok) return 3;
    if ((h->mode == 0 || h->mode == 2) && h->ll_ok) return 2;
    return 1;
}

extern "C" int wn_gen_launch_info(const wn_gen_handle* h, int* grid, int* block, int* barriers_per_eval) {
    WN_REQUIRE(h, WN_E_STATE, "wn_gen_launch_info: null handle");
    // the kernel wn_gen_run would pick in the handle's current mode (same selection as in wn_gen_run)
    const bool auto_cluster = h->mode == 0 && h->cluster_ok && (h->shape.n_streams > 1 || !h->fast_ok);
    const bool cl8 = (h->mode == 0 || h->mode == 6) && h->cl8_ok;
    const bool cluster = !cl8 && (auto_cluster || h->mode == 4) && h->cluster_ok;
    const bool fast = !cluster && ((h->mode == 5 && h->x2_ok) || ((h->mode == 0 || h->mode == 3) && h->fast_ok));
    if (grid) *grid = cl8 ? (h->shape.n_streams + CL8_SB - 1) / CL8_SB * (h->cl8_cs ? h->cl8_cs : CL) : cluster ? h->shape.n_streams * CL : h->grid;
    if (block) *block = cl8 ? GEN_NT + 64 + (h->cl8_cs == 8 ? 32 : 0) : (cluster || fast) ? GEN_NT + 32 : GEN_NT;   // + helper warps
    if (barriers_per_eval) *barriers_per_eval = 2 * h->shape.n_layers + 2;      // exchange stages per evaluation
    return 0;
}
