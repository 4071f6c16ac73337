// wgrad.cu -- weight gradients of the 1x1 / tap convolutions, exact fp32.
//
// The reference leaves these to autograd (conv1d backward-weight of wavenet_model.py:145-165 and :167-169): for every
// convolution  dW[n][c] = sum over sequences b and frames t of  g[b][t][n] * x[b][t][c],  with g the gradient of the
// convolution output and x its (tap-shifted) input.  In the frames layout both operands are (frames x channels) with
// channels contiguous, i.e. the contraction runs over the SLOW axis of both: a tall-skinny "A^T B" product with
// M = N_out <= 512, N = C_in <= 256 and K = B*L ~ 128 000.  Library SGEMMs handle that shape poorly (few output
// tiles, enormous K); here the frames axis is cut into `splits` ranges so that (output tiles x splits) fills the
// 148 SMs about twice, every CTA accumulates a 128x128 partial over its range by rank-1 updates straight from the
// row-major operands (coalesced 512-byte rows, no transposes), and a second small kernel adds the partials in a
// fixed order (deterministic result) and scatters them into the (out, in, k) weight-gradient tensor.
#include "common.cuh"

namespace {

constexpr int WG_TILE = 128;      // output tile: 128 gradient channels x 128 input channels
constexpr int WG_KS   = 16;       // frames per shared-memory slab
constexpr int WG_NT   = 256;      // 16 x 16 threads, 8 x 8 accumulators each
constexpr int WG_TARGET_CTAS = 296;

struct WgradParams {
    const float* g; const float* x; float* work;
    long long g_seq, x_seq;
    int ldg, ldx, B, rows, N, C;
    int tiles_c, splits, chunk;           // chunk = frames per split (multiple of WG_KS)
    bool vec_g, vec_x;
};

// one 16-frame x 128-channel slab: thread -> (frame r and r+8, four channels).  The (sequence, frame) position of
// the thread's two rows is carried along incrementally (no division in the loop).
struct SlabRegs { float4 v[2]; };
struct RowPos { int kk[2], b[2], t[2]; };

__device__ __forceinline__ RowPos row_pos_init(int k_beg, int rows, int tid) {
    RowPos q;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        q.kk[h] = k_beg + (tid >> 5) + 8 * h;
        q.b[h] = q.kk[h] / rows;
        q.t[h] = q.kk[h] - q.b[h] * rows;
    }
    return q;
}
__device__ __forceinline__ void row_pos_advance(RowPos& q, int rows) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        q.kk[h] += WG_KS;
        q.t[h] += WG_KS;
        while (q.t[h] >= rows) { q.t[h] -= rows; ++q.b[h]; }
    }
}

__device__ __forceinline__ SlabRegs load_slab(const float* base, long long seq, int ld, int nch, int ch0,
                                              const RowPos& q, int k_end, bool vec, int tid) {
    SlabRegs s;
    const int c = ch0 + ((tid & 31) << 2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q.kk[h] < k_end && c < nch) {
            const float* p = base + q.b[h] * seq + (long long)q.t[h] * ld + c;
            if (vec) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                v.x = __ldg(p);
                if (c + 1 < nch) v.y = __ldg(p + 1);
                if (c + 2 < nch) v.z = __ldg(p + 2);
                if (c + 3 < nch) v.w = __ldg(p + 3);
            }
        }
        s.v[h] = v;
    }
    return s;
}

__global__ void __launch_bounds__(WG_NT, 2) wgrad_partial_kernel(const WgradParams p) {
    __shared__ __align__(16) float Gs[2][WG_KS][WG_TILE];
    __shared__ __align__(16) float Xs[2][WG_KS][WG_TILE];
    const int tid = threadIdx.x;
    const int tile_n = blockIdx.x / p.tiles_c, tile_c = blockIdx.x % p.tiles_c;
    const int n0 = tile_n * WG_TILE, c0 = tile_c * WG_TILE;
    const int k_total = p.B * p.rows;                       // host checks B*rows < 2^30
    const int k_beg = (int)blockIdx.y * p.chunk;
    const int k_end = (k_total - k_beg > p.chunk) ? k_beg + p.chunk : k_total;
    const int ty = tid >> 4, tx = tid & 15;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int n_slabs = (k_end > k_beg) ? (k_end - k_beg + WG_KS - 1) / WG_KS : 0;
    const int sr = tid >> 5, sc = (tid & 31) << 2;
    RowPos q = row_pos_init(k_beg, p.rows, tid);
    if (n_slabs > 0) {
        SlabRegs g = load_slab(p.g, p.g_seq, p.ldg, p.N, n0, q, k_end, p.vec_g, tid);
        SlabRegs x = load_slab(p.x, p.x_seq, p.ldx, p.C, c0, q, k_end, p.vec_x, tid);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            *reinterpret_cast<float4*>(&Gs[0][sr + 8 * h][sc]) = g.v[h];
            *reinterpret_cast<float4*>(&Xs[0][sr + 8 * h][sc]) = x.v[h];
        }
    }
    __syncthreads();
    for (int s = 0; s < n_slabs; ++s) {
        const int cur = s & 1;
        SlabRegs g, x;
        const bool more = s + 1 < n_slabs;
        if (more) {
            row_pos_advance(q, p.rows);
            g = load_slab(p.g, p.g_seq, p.ldg, p.N, n0, q, k_end, p.vec_g, tid);
            x = load_slab(p.x, p.x_seq, p.ldx, p.C, c0, q, k_end, p.vec_x, tid);
        }
#pragma unroll
        for (int k = 0; k < WG_KS; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&Gs[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Gs[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Xs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Xs[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                *reinterpret_cast<float4*>(&Gs[cur ^ 1][sr + 8 * h][sc]) = g.v[h];
                *reinterpret_cast<float4*>(&Xs[cur ^ 1][sr + 8 * h][sc]) = x.v[h];
            }
        }
        __syncthreads();
    }

    // partial tile -> work[split][n][c]
    float* out = p.work + (size_t)blockIdx.y * p.N * p.C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (n >= p.N) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c < p.C) out[(size_t)n * p.C + c] = acc[i][j];
        }
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ work, float* __restrict__ dw, int N, int C, int splits,
                                    long long n_stride, long long c_stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += work[(size_t)k * N * C + idx];
    const int n = idx / C, c = idx - n * C;
    dw[n * n_stride + c * c_stride] = s;
}

int max_splits(int N, int C) {
    const int tiles = wn::ceil_div(N, WG_TILE) * wn::ceil_div(C, WG_TILE);
    return wn::ceil_div(WG_TARGET_CTAS, tiles);
}

}  // namespace

extern "C" size_t wn_wgrad_workspace_bytes(int N, int C) {
    if (N <= 0 || C <= 0) return 0;
    return (size_t)max_splits(N, C) * (size_t)N * (size_t)C * sizeof(float);
}

extern "C" int wn_wgrad(const wn_wgrad_args* a, void* stream) {
    WN_REQUIRE(a != nullptr, WN_E_BADARG, "wn_wgrad: null argument block");
    WN_REQUIRE(a->N > 0 && a->C > 0 && a->B > 0 && a->rows >= 0, WN_E_BADARG, "wn_wgrad: bad sizes N=%d C=%d B=%d rows=%d",
               a->N, a->C, a->B, a->rows);
    WN_REQUIRE(a->d_dw && a->d_work && (a->rows == 0 || (a->d_g && a->d_x)), WN_E_BADARG, "wn_wgrad: null device pointer");
    WN_REQUIRE(a->ldg >= a->N && a->ldx >= a->C, WN_E_BADARG, "wn_wgrad: row pitch smaller than the channel count");
    WN_REQUIRE((long long)a->B * a->rows < (1ll << 30), WN_E_UNSUPP, "wn_wgrad: B*rows = %lld frames exceeds 2^30",
               (long long)a->B * a->rows);
    cudaStream_t st = (cudaStream_t)stream;
    const long long k_total = (long long)a->B * a->rows;
    const int tiles_n = wn::ceil_div(a->N, WG_TILE), tiles_c = wn::ceil_div(a->C, WG_TILE);
    int splits = max_splits(a->N, a->C);
    // at least four slabs of frames per split, at least one split
    const long long by_k = k_total / (4 * WG_KS);
    if ((long long)splits > by_k) splits = by_k > 0 ? (int)by_k : 1;
    long long chunk = (k_total + splits - 1) / splits;
    chunk = (chunk + WG_KS - 1) / WG_KS * WG_KS;
    if (chunk <= 0) chunk = WG_KS;
    splits = k_total > 0 ? (int)((k_total + chunk - 1) / chunk) : 1;

    WgradParams p;
    p.g = a->d_g; p.x = a->d_x; p.work = a->d_work;
    p.g_seq = a->g_seq_stride; p.x_seq = a->x_seq_stride;
    p.ldg = a->ldg; p.ldx = a->ldx; p.B = a->B; p.rows = a->rows > 0 ? a->rows : 1; p.N = a->N; p.C = a->C;
    if (a->rows == 0) p.B = 0;
    p.tiles_c = tiles_c; p.splits = splits; p.chunk = (int)chunk;
    auto vec_ok = [](const float* ptr, long long seq, int ld, int n) {
        return ((uintptr_t)ptr % 16 == 0) && (seq % 4 == 0) && (ld % 4 == 0) && (n % 4 == 0);
    };
    p.vec_g = a->rows > 0 && vec_ok(a->d_g, a->g_seq_stride, a->ldg, a->N);
    p.vec_x = a->rows > 0 && vec_ok(a->d_x, a->x_seq_stride, a->ldx, a->C);
    dim3 grid(tiles_n * tiles_c, splits);
    wgrad_partial_kernel<<<grid, WG_NT, 0, st>>>(p);
    WN_CUDA(cudaGetLastError());
    const int total = a->N * a->C;
    wgrad_reduce_kernel<<<wn::ceil_div(total, 256), 256, 0, st>>>(a->d_work, a->d_dw, a->N, a->C, splits,
                                                                 a->dw_n_stride, a->dw_c_stride);
    WN_CUDA(cudaGetLastError());
    return 0;
}
