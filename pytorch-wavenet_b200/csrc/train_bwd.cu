// train_bwd.cu -- data-gradient kernels of the training path (exact fp32), the mirror of train_fwd.cu.
//
// The reference gets its backward from autograd over dilate/conv1d/tanh/sigmoid (wavenet_training.py:71 calls
// loss.backward() on the graph built by wavenet_model.py:125-171).  Here each residual block's data gradient is two
// launches on the absolute time axis (no fold/un-fold of gradients either):
//
//   dz kernel   dZ[TM x D]  = [dh_out(t) | dskip(t)] * [Wr; Ws]            (K = R+S)
//               dF = dZ * g * (1 - f^2),  dG = dZ * f * g * (1 - g)         f, g saved by the forward
//               -> dfg (B,L,2D) and z = f*g (B,L,D) for the weight gradients
//   dh kernel   dh_in[t] = dh_out[t] + sum_j [dF|dG](t + (k-1-j) d) * [Wf;Wg][:,:,j]      (anti-causal taps)
//
// and the head's data gradient is three row-GEMMs with relu masks.  Frames where a gradient is structurally zero
// (left of the receptive cone of the last `out_len` outputs) are neither read nor written: every buffer carries
// a "first valid frame" and loaders return zero left of it.
// Weight gradients are plain GEMMs over these buffers ( dW = X^T Y ) and are taken by the host with library GEMMs.
#include "sgemm_core.cuh"
#include <type_traits>

namespace wn {

// ------------------------------------------------------------------------------------------------ loaders
struct ConcatGradLoader {       // A row = [dh_out(t) (R) | dskip(t) (S)]
    const float* dh;            // dh_out + b*L*R or nullptr
    const float* ds;            // dskip + b*Tds*S
    int R, S, t0, L, gs_out, ds_start;
    bool vec;
    __device__ __forceinline__ float load1(int m, int kidx) const {
        const int t = t0 + m;
        if (t >= L) return 0.f;
        if (kidx < R) return (dh != nullptr && t >= gs_out) ? __ldg(dh + (size_t)t * R + kidx) : 0.f;
        if (kidx < R + S) return (t >= ds_start) ? __ldg(ds + (size_t)(t - ds_start) * S + (kidx - R)) : 0.f;
        return 0.f;
    }
    __device__ __forceinline__ float4 load4(int m, int kidx) const {
        const int t = t0 + m;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= L) return z4;
        if (kidx < R) return (dh != nullptr && t >= gs_out) ? __ldg(reinterpret_cast<const float4*>(dh + (size_t)t * R + kidx)) : z4;
        if (kidx < R + S)
            return (t >= ds_start) ? __ldg(reinterpret_cast<const float4*>(ds + (size_t)(t - ds_start) * S + (kidx - R))) : z4;
        return z4;
    }
};

struct FutureTapLoader {        // A row = [dfg(t + (k-1)d) | ... | dfg(t)]  (tap j pairs with weight [:,:,j])
    const float* dfg;           // dfg + b*L*N2
    int N2, ktaps, dil, t0, L, gz, K;
    bool vec;
    __device__ __forceinline__ const float* addr(int m, int kidx, bool& ok) const {
        const int j = kidx / N2, n2 = kidx - j * N2;
        const int ts = t0 + m + (ktaps - 1 - j) * dil;
        ok = (kidx < K) && (ts < L) && (ts >= gz);
        return dfg + (size_t)ts * N2 + n2;
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

struct RowLoader {              // A row = rows[frame t - row0] (optionally relu'd), a plain (frames, K) matrix per sequence
    const float* rows;          // + b * n_rows * K
    int K, t0, L, row0;
    bool relu, vec;
    __device__ __forceinline__ float load1(int m, int kidx) const {
        const int t = t0 + m;
        if (kidx >= K || t >= L) return 0.f;
        const float v = __ldg(rows + (size_t)(t - row0) * K + kidx);
        return relu ? fmaxf(v, 0.f) : v;
    }
    __device__ __forceinline__ float4 load4(int m, int kidx) const {
        const int t = t0 + m;
        if (kidx >= K || t >= L) return make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = __ldg(reinterpret_cast<const float4*>(rows + (size_t)(t - row0) * K + kidx));
        if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        return v;
    }
};

// ------------------------------------------------------------------------------------------------ epilogues
// store(acc value, sequence b, frame t, output column n)
struct GateBwdEpi {
    const float* fg; float* dfg; float* z;
    int L, D;
    __device__ __forceinline__ void operator()(float dz, int b, int t, int n) const {
        if (n >= D) return;
        const size_t base = ((size_t)b * L + t) * (2 * D);
        const float f = __ldg(fg + base + n), g = __ldg(fg + base + D + n);
        dfg[base + n] = dz * g * (1.f - f * f);
        dfg[base + D + n] = dz * f * g * (1.f - g);
        z[((size_t)b * L + t) * D + n] = f * g;
    }
};
struct ResidualAddEpi {
    const float* dh_out; float* dh_in;
    int L, R, id_start;         // identity path: frames >= id_start carry dh_out straight through
    __device__ __forceinline__ void operator()(float v, int b, int t, int n) const {
        if (n >= R) return;
        const size_t i = ((size_t)b * L + t) * R + n;
        if (dh_out != nullptr && t >= id_start) v += __ldg(dh_out + i);
        dh_in[i] = v;
    }
};
struct MaskRowsEpi {            // out[b][t-row0][n] = (bias ? relu(v + bias) : v * (act[b][t-act_row0][n] > 0))
    const float* act; const float* bias; float* out;
    int N, rows_per_seq, row0, act_rows_per_seq, act_row0;
    __device__ __forceinline__ void operator()(float v, int b, int t, int n) const {
        if (n >= N) return;
        const size_t i = ((size_t)b * rows_per_seq + (t - row0)) * N + n;
        if (bias != nullptr) out[i] = fmaxf(v + __ldg(bias + n), 0.f);
        else out[i] = (__ldg(act + ((size_t)b * act_rows_per_seq + (t - act_row0)) * N + n) > 0.f) ? v : 0.f;
    }
};

// ------------------------------------------------------------------------------------------------ generic row GEMM
template <int TM, class AL, class EP>
__global__ void __launch_bounds__(NT, 1) rows_gemm_kernel(AL al, const EP ep, const float* __restrict__ w_t, int ldw,
                                                          int K, int N, int t_begin, int L, size_t a_seq_stride,
                                                          size_t a2_seq_stride) {
    using T = Tile<TM>;
    constexpr int MI = T::MI;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Bs = As + 2 * KS * TM;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y, t0 = t_begin + blockIdx.x * TM;
    al.t0 = t0;
    if constexpr (std::is_same<AL, ConcatGradLoader>::value) {
        if (al.dh) al.dh += (size_t)b * a_seq_stride;
        al.ds += (size_t)b * a2_seq_stride;
    } else if constexpr (std::is_same<AL, FutureTapLoader>::value) {
        al.dfg += (size_t)b * a_seq_stride;
    } else {
        al.rows += (size_t)b * a_seq_stride;
    }
    const int n_chunks = (N + NC - 1) / NC;
    for (int ch = 0; ch < n_chunks; ++ch) {
        float acc[MI][8];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        mainloop<TM, false>(acc, al, nullptr, w_t, ldw, ch * NC, K, As, Bs);
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int t = t0 + T::row(ty, i);
                if (t >= L) continue;
#pragma unroll
                for (int q = 0; q < 4; ++q) ep(acc[i][g * 4 + q], b, t, ch * NC + g * 64 + tx * 4 + q);
            }
    }
}

template <class AL, class EP>
static int launch_rows(const AL& al, const EP& ep, const float* w_t, int ldw, int K, int N, int t_begin, int L, int B,
                       size_t s1, size_t s2, cudaStream_t st) {
    if (t_begin >= L) return 0;
    constexpr int TM = 128;
    const size_t smem = sizeof(float) * (2 * KS * TM + 2 * KS * NC);
    rows_gemm_kernel<TM, AL, EP><<<dim3(ceil_div(L - t_begin, TM), B), NT, smem, st>>>(al, ep, w_t, ldw, K, N, t_begin, L,
                                                                                   s1, s2);
    WN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace wn

using namespace wn;

extern "C" int wn_block_bwd_data(const wn_block_bwd_args* a, void* stream) {
    WN_REQUIRE(a, WN_E_BADARG, "wn_block_bwd_data: null args");
    WN_REQUIRE(a->d_dskip && a->d_fg && a->d_dfg && a->d_z && a->d_dh_in && a->d_wrs_rows && a->d_wfg_bwd, WN_E_BADARG,
               "wn_block_bwd_data: null pointer");
    WN_REQUIRE(a->B > 0 && a->L > 0 && a->R > 0 && a->D > 0 && a->S > 0 && a->k >= 1 && a->dilation >= 1, WN_E_BADARG,
               "wn_block_bwd_data: bad shape");
    WN_REQUIRE(a->gz >= a->out_start && a->gs_in >= a->in_start && a->ds_start >= a->out_start && a->ds_start <= a->L,
               WN_E_BADARG, "wn_block_bwd_data: bad gradient frame ranges");
    cudaStream_t st = (cudaStream_t)stream;
    const int R = a->R, D = a->D, S = a->S, L = a->L, B = a->B;
    // ---- dz + gate backward
    ConcatGradLoader cl;
    cl.dh = a->d_dh_out; cl.ds = a->d_dskip; cl.R = R; cl.S = S; cl.t0 = 0; cl.L = L; cl.gs_out = a->gs_out;
    cl.ds_start = a->ds_start;
    cl.vec = (R % KS == 0) && (S % KS == 0);
    GateBwdEpi ge{a->d_fg, a->d_dfg, a->d_z, L, D};
    if (int rc = launch_rows(cl, ge, a->d_wrs_rows, n2p_of(D), R + S, D, a->gz, L, B, (size_t)L * R,
                             (size_t)(L - a->ds_start) * S, st))
        return rc;
    // ---- dh_in: anti-causal taps of dfg + identity
    FutureTapLoader fl;
    fl.dfg = a->d_dfg; fl.N2 = 2 * D; fl.ktaps = a->k; fl.dil = a->dilation; fl.t0 = 0; fl.L = L; fl.gz = a->gz;
    fl.K = a->k * 2 * D;
    fl.vec = ((2 * D) % KS == 0);
    const int id_start = a->gs_out > a->out_start ? a->gs_out : a->out_start;
    ResidualAddEpi re{a->d_dh_out, a->d_dh_in, L, R, id_start};
    return launch_rows(fl, re, a->d_wfg_bwd, n2p_of(R), fl.K, R, a->gs_in, L, B, (size_t)L * 2 * D, 0, st);
}

extern "C" int wn_head_bwd_data(const wn_head_bwd_args* a, void* stream) {
    WN_REQUIRE(a, WN_E_BADARG, "wn_head_bwd_data: null args");
    WN_REQUIRE(a->d_dlogits && a->d_skip && a->d_y1 && a->d_dy1 && a->d_dskip && a->d_w1_t && a->d_b1 && a->d_w2_rows &&
                   a->d_w1_rows,
               WN_E_BADARG, "wn_head_bwd_data: null pointer");
    WN_REQUIRE(a->B > 0 && a->L > 0 && a->S > 0 && a->E > 0 && a->classes > 0 && a->out_len > 0 &&
                   a->out_len <= a->L - a->skip_start,
               WN_E_BADARG, "wn_head_bwd_data: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = a->L, B = a->B, OL = a->out_len, S = a->S, E = a->E, C = a->classes;
    const int t_first = L - OL;
    // y1 = relu(W1 relu(skip) + b1) for the last OL frames (recomputed: cheaper than saving it in the forward)
    RowLoader sl;
    sl.rows = a->d_skip; sl.K = S; sl.t0 = 0; sl.L = L; sl.row0 = a->skip_start; sl.relu = true; sl.vec = (S % KS == 0);
    MaskRowsEpi e1{nullptr, a->d_b1, a->d_y1, E, OL, t_first, OL, t_first};
    if (int rc = launch_rows(sl, e1, a->d_w1_t, n2p_of(E), S, E, t_first, L, B, (size_t)(L - a->skip_start) * S, 0, st))
        return rc;
    // dy1 = (dlogits W2) * (y1 > 0)
    RowLoader gl;
    gl.rows = a->d_dlogits; gl.K = C; gl.t0 = 0; gl.L = L; gl.row0 = t_first; gl.relu = false; gl.vec = (C % KS == 0);
    MaskRowsEpi e2{a->d_y1, nullptr, a->d_dy1, E, OL, t_first, OL, t_first};
    if (int rc = launch_rows(gl, e2, a->d_w2_rows, n2p_of(E), C, E, t_first, L, B, (size_t)OL * C, 0, st)) return rc;
    // dskip = (dy1 W1) * (skip > 0)   -- mask read from the saved skip rows of the same frames
    RowLoader yl;
    yl.rows = a->d_dy1; yl.K = E; yl.t0 = 0; yl.L = L; yl.row0 = t_first; yl.relu = false; yl.vec = (E % KS == 0);
    MaskRowsEpi e3{a->d_skip, nullptr, a->d_dskip, S, OL, t_first, L - a->skip_start, a->skip_start};
    return launch_rows(yl, e3, a->d_w1_rows, n2p_of(S), E, S, t_first, L, B, (size_t)OL * E, 0, st);
}
