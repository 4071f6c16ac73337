// tc_gemm.cu -- tensor-core (tcgen05 / TMEM / TMA) form of the training block: exact-fp32-class numerics through
// 3xTF32 (hi/lo split of both operands, fp32 accumulation in tensor memory).
//
// Same mathematics and frames layout as train_fwd.cu (reference wavenet_model.py:142-165), split in two launches
// per block because the gated activation cannot stay on chip in split form (128 frames x 256 ch x {hi,lo} = 256 KB):
//   pass A   FG[128 frames x 256] per tile = A[128 x kR] * Wa^T      A rows = taps of h_in (TMA boxes shifted by
//            the dilation; frames left of in_start come back as zeros from the TMA out-of-bounds fill)
//            epilogue: z = tanh(F+bf) * sigmoid(G+bg)  -> z (B,L,D)  [+ optional f,g for the backward]
//   pass B   [O|S][128 x 256] per tile = z[128 x D] * Wb^T ;  h_out = O + br + h_in,  skip (+)= S + bs
// Kernel anatomy (one CTA per SM, persistent over (sequence, 128-frame tile) items, 448 threads):
//   warp 0        TMA producer: per K slab (16 fp32 = one 64-byte swizzle row) loads A raw, W_hi, W_lo (4-stage ring)
//   warp 1        allocates TMEM, issues tcgen05.mma kind::tf32 (M128 N256 K8): hi*hi + lo*hi + hi*lo per k-step
//   warps 2,3,8,9 splitter: rewrite the landed A slab as hi = rna_tf32(x) in place and lo = x - hi in a second buffer
//   warps 4-7,10-13 epilogue (2 groups x 4 TMEM lane quadrants): tcgen05.ld the finished accumulator (2 x 256 TMEM
//                 columns, double buffered), apply gate / residual / skip, store whole sectors
// mbarriers: full (TMA landed), split (lo ready), empty (MMAs of the stage retired), acc_full / acc_empty.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>

namespace wn {
namespace tc {

constexpr int BM = 128;            // frames per tile (UMMA M)
constexpr int BN = 256;            // output columns per tile (UMMA N)
constexpr int BK = 16;             // fp32 per K slab = 64 bytes = one 64B-swizzle row (two k-steps of 8).  Measured on B200:
                                   // 128B rows x 2 stages 36.4 ms, 64B x 4 stages 27.7 ms, 32B x 8 stages 33.5 ms per cfg-3
                                   // forward -- a stage slot needs ~3500 cycles to come round (TMA ~1900, split ~700), so
                                   // depth matters, but TMA latency grows again when the boxes get too small
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;          // 8 KB
constexpr int W_BYTES = BN * BK * 4;          // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;      // A_hi | A_lo | W_hi | W_lo = 48 KB
// bf16x2 precision (PREC_BF16X2): the operand tiles are bf16 (hi, lo) pairs with 32-byte rows; the raw fp32 A slab keeps
// its own buffer, so a stage is A_raw 8K | A_hi 4K | A_lo 4K | W_hi 8K | W_lo 8K = 32 KB and six stages fit
constexpr int STAGES_BF = 6;
constexpr int ABF_BYTES = BM * BK * 2;        // 4 KB
constexpr int WBF_BYTES = BN * BK * 2;        // 8 KB
constexpr int STAGE_BYTES_BF = A_BYTES + 2 * ABF_BYTES + 2 * WBF_BYTES;     // 32 KB
enum { PREC_TF32X3 = 0, PREC_TF32X1 = 1, PREC_BF16X2 = 2 };
constexpr int NTHREADS = 448;             // 14 warps: TMA, MMA, 4 split (2,3,8,9), 2 x 4 epilogue (4-7 and 10-13)
constexpr int EPI_THREADS = 256;
constexpr int SPLIT_THREADS = 128;
constexpr int CS = 2;                       // CTAs per cluster sharing every weight slab through TMA multicast
constexpr int TP = 20;                      // pitch of the 32x16 epilogue transpose tile (16-byte aligned rows)
constexpr unsigned SPIN_LIMIT = 1u << 28;     // a barrier that never completes traps instead of hanging the GPU

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done, spins = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
        if (!done && ++spins > SPIN_LIMIT) asm volatile("trap;");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar,
                                               unsigned short mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(s32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(unsigned long long* bar, unsigned short mask) {   // arrive on `bar` of every CTA in mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(s32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_rank_() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(unsigned* slot_in_smem, unsigned cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr, unsigned cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                         unsigned accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {       // arrives when all prior MMAs of this thread retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major rows of BK*4 bytes with the matching swizzle
// (64B rows -> SWIZZLE_64B, 128B rows -> SWIZZLE_128B); 8-row atoms are 8*row_bytes apart
template <unsigned row_bytes = BK * 4>
__device__ __forceinline__ unsigned long long smem_desc(unsigned saddr) {
    constexpr unsigned long long layout = (row_bytes == 128) ? 2ull : (row_bytes == 64 ? 4ull : 6ull);   // SWIZZLE_128B/64B/32B
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);            // start address, 16-byte units
    d |= (unsigned long long)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)((8 * row_bytes) >> 4) << 32;       // stride byte offset between 8-row atoms
    d |= (unsigned long long)1 << 46;                            // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B tf32 (format 2, kind::tf32) or bf16 (format 1,
// kind::f16), both K-major, M=128, N=256
__host__ __device__ constexpr unsigned make_idesc(bool bf16 = false) {
    return (1u << 4) | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- kernel
enum { EPI_GATE = 0, EPI_RES_SKIP = 1, EPI_GATE_BWD = 2, EPI_ADD = 3 };

struct TcParams {
    int B, L, t_begin;            // frames [t_begin, L) of every sequence are produced
    int taps, dil, C;             // A row = taps x C channels; tap j reads frame t - (taps-1-j)*dil  (dil < 0: future frames)
    int a_origin;                 // frame that coordinate 0 of the A tensor map corresponds to
    // optional second A source appended along K (backward dz: A row = [dh_out(t) | dskip(t)]): C2 channels read through
    // mapA2 at frame t - a2_origin; when C2 > 0 and C == 0 the first source is absent (last layer: no dh_out)
    int C2, a2_origin;
    int n_tiles, n_total;         // output columns = n_tiles * BN; the W map holds hi rows [0,n_total) then lo rows
    // epilogue
    const float* bias;            // [n_total] in tile column order
    float* out0;                  // GATE: z (B,L,D)            RES_SKIP: h_out (B,L,R)
    float* out1;                  // GATE: fg_save (B,L,2D)|0   RES_SKIP: skip (B,L-skip_start,S)
    const float* res;             // RES_SKIP: h_in (B,L,R)     GATE_BWD: fg (B,L,2D)      ADD: dh_out (B,L,R) or null
    float* out2;                  // GATE_BWD: z (B,L,D)
    int id_start;                 // ADD: frames >= id_start carry `res` straight through
    int D, R, S, in_start, skip_start, skip_init;
    long long* dbg;               // optional trace buffer (WN_TC_TRACE): CTA 0 stamps clock64 per stage, see tools/tc_trace.py
};

__device__ __forceinline__ float sigmoid_tc(float x) { return 1.f / (1.f + expf(-x)); }

// PREC_TF32X3: 3xTF32 (hi*hi + lo*hi + hi*lo, hi = rna_tf32(x)): ~6e-7 on the logits after 50 layers.
// PREC_BF16X2: both operands as bf16 pairs (hi = bf16(x), lo = bf16(x - hi)), the same three products on kind::f16 at
//              twice the tf32 rate: 16 mantissa bits per operand, ~3e-6 on the logits after 50 layers (inside the 1e-4 bar).
// PREC_TF32X1: one TF32 MMA per k-step on the raw fp32 operands (the tensor core drops the low mantissa bits):
//              ~1e-3 relative on the logits after 50 layers, i.e. outside the parity bar; opt-in, reported separately.
template <int EPI, int PREC>
__global__ void __launch_bounds__(NTHREADS, 1)
frames_gemm_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
               const __grid_constant__ CUtensorMap mapW, const TcParams p) {
    constexpr bool EXACT = PREC == PREC_TF32X3;
    constexpr bool BF = PREC == PREC_BF16X2;
    constexpr int ST = BF ? STAGES_BF : STAGES;                // ring depth
    constexpr int SB = BF ? STAGE_BYTES_BF : STAGE_BYTES;      // bytes per stage
    // mapW's box is BN/CS rows: every CTA of the cluster fetches its share of a weight slab and multicasts it to all
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* stage_mem = base;                                           // ST * SB
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + ST * SB);
    unsigned long long* full = bars;                 // [ST]
    unsigned long long* split = bars + ST;           // [ST]
    unsigned long long* empty = bars + 2 * ST;       // [ST]
    unsigned long long* acc_full = bars + 3 * ST;          // [2]
    unsigned long long* acc_empty = bars + 3 * ST + 2;     // [2]
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 3 * ST + 4);
    float* bias_s = reinterpret_cast<float*>(base + ST * SB + 512);                                // [n_total]
    float* stage_t = bias_s + ((p.n_total + 3) & ~3);                          // [4 warps][32][TP] epilogue transpose tiles

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_tiles = (p.L - p.t_begin + BM - 1) / BM;
    const int items = p.B * m_tiles;
    // Work: item (sequence b, 128-frame tile) number (cluster + round * n_clusters) * CS + rank.  The CTAs of a cluster walk
    // the same (output tile, K slab) sequence in lockstep because they share the weight slabs; a CTA whose item does not
    // exist runs a ghost tile (frames beyond L: TMA zero fill, no stores) so that it keeps feeding its peers.
    const int crank = (int)cluster_rank_(), n_clusters = gridDim.x / CS, cluster_id = blockIdx.x / CS;
    const int rounds = (items + n_clusters * CS - 1) / (n_clusters * CS);
    const unsigned short mc_mask = (unsigned short)((1u << CS) - 1u);
    const int slabs_per_tap = p.C / BK;
    const int slabs1 = p.taps * slabs_per_tap;                  // K slabs of the first A source
    const int slabs = slabs1 + p.C2 / BK;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < ST; ++i) { mbar_init(full + i, 1); mbar_init(split + i, SPLIT_THREADS); mbar_init(empty + i, CS); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    for (int i = tid; i < p.n_total; i += NTHREADS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                              // peers' barriers are initialised before anything is multicast at them
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (elect_one()) {
            unsigned it = 0;
            for (int rd = 0; rd < rounds; ++rd) {
                const int item = (cluster_id + rd * n_clusters) * CS + crank;
                const bool ghost = item >= items;
                const int b = ghost ? 0 : item / m_tiles, t0 = ghost ? p.L : p.t_begin + (item % m_tiles) * BM;
                for (int nt = 0; nt < p.n_tiles; ++nt)
                    for (int sl = 0; sl < slabs; ++sl, ++it) {
                        const int st = it % ST;
                        const unsigned ph = (it / ST) & 1;
                        if (p.dbg && blockIdx.x == 0 && it < 512) p.dbg[it * 8 + 0] = clock64();
                        mbar_wait(empty + st, ph ^ 1);           // every CTA of the cluster is done with this stage
                        if (p.dbg && blockIdx.x == 0 && it < 512) p.dbg[it * 8 + 1] = clock64();
                        unsigned char* sm = stage_mem + st * SB;
                        mbar_expect_tx(full + st, BF ? A_BYTES + 2 * WBF_BYTES : A_BYTES + (EXACT ? 2 : 1) * W_BYTES);
                        if (sl < slabs1) {
                            const int j = sl / slabs_per_tap, c0 = (sl % slabs_per_tap) * BK;
                            tma_load_3d(sm, &mapA, c0, t0 - (p.taps - 1 - j) * p.dil - p.a_origin, b, full + st);
                        } else {
                            tma_load_3d(sm, &mapA2, (sl - slabs1) * BK, t0 - p.a2_origin, b, full + st);
                        }
                        constexpr int WR = BN / CS;                          // this CTA's rows of the slab
                        constexpr int WT = BF ? WBF_BYTES : W_BYTES;        // bytes of one weight tile (hi or lo)
                        constexpr int W0 = BF ? A_BYTES + 2 * ABF_BYTES : 2 * A_BYTES;      // offset of W_hi in the stage
                        tma_load_2d_mc(sm + W0 + crank * (WT / CS), &mapW, sl * BK, nt * BN + crank * WR, full + st, mc_mask);
                        if (EXACT || BF)
                            tma_load_2d_mc(sm + W0 + WT + crank * (WT / CS), &mapW, sl * BK, p.n_total + nt * BN + crank * WR,
                                           full + st, mc_mask);
                    }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr unsigned idesc = make_idesc(BF);
        unsigned it = 0, tile = 0;
        for (int rd = 0; rd < rounds; ++rd)
            for (int nt = 0; nt < p.n_tiles; ++nt, ++tile) {
                const unsigned ab = tile & 1, aph = (tile >> 1) & 1;
                mbar_wait(acc_empty + ab, aph ^ 1);
                tc_fence_after();
                const unsigned d_tmem = tmem_base + ab * BN;
                for (int sl = 0; sl < slabs; ++sl, ++it) {
                    const int st = it % ST;
                    const unsigned ph = (it / ST) & 1;
                    if (p.dbg && blockIdx.x == 0 && it < 512 && lane == 0) p.dbg[it * 8 + 2] = clock64();
                    mbar_wait(split + st, ph);                   // TMA landed and the splitter produced hi/lo
                    if (p.dbg && blockIdx.x == 0 && it < 512 && lane == 0) p.dbg[it * 8 + 3] = clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const unsigned sa = s32(stage_mem + st * SB);
                        if constexpr (BF) {
                            // one k-step per slab: 16 bf16 = one 32-byte swizzle row
                            const unsigned long long a_hi = smem_desc<32>(sa + A_BYTES), a_lo = smem_desc<32>(sa + A_BYTES + ABF_BYTES);
                            const unsigned long long w_hi = smem_desc<32>(sa + A_BYTES + 2 * ABF_BYTES);
                            const unsigned long long w_lo = smem_desc<32>(sa + A_BYTES + 2 * ABF_BYTES + WBF_BYTES);
                            umma_f16(d_tmem, a_hi, w_hi, idesc, sl != 0);
                            umma_f16(d_tmem, a_lo, w_hi, idesc, 1);
                            umma_f16(d_tmem, a_hi, w_lo, idesc, 1);
                        } else {
                            const unsigned long long a_hi = smem_desc(sa), a_lo = smem_desc(sa + A_BYTES);
                            const unsigned long long w_hi = smem_desc(sa + 2 * A_BYTES), w_lo = smem_desc(sa + 2 * A_BYTES + W_BYTES);
#pragma unroll
                            for (int kk = 0; kk < BK / 8; ++kk) {    // 8 tf32 = 32 bytes = 2 descriptor units per k-step
                                const unsigned long long o = (unsigned long long)(kk * 2);
                                umma_tf32(d_tmem, a_hi + o, w_hi + o, idesc, (sl | kk) != 0);
                                if (EXACT) {
                                    umma_tf32(d_tmem, a_lo + o, w_hi + o, idesc, 1);
                                    umma_tf32(d_tmem, a_hi + o, w_lo + o, idesc, 1);
                                }
                            }
                        }
                        umma_commit_mc(empty + st, mc_mask);     // stage reusable (in every CTA) once these MMAs retire
                        if (sl == slabs - 1) umma_commit(acc_full + ab);
                    }
                    __syncwarp();
                    if (p.dbg && blockIdx.x == 0 && it < 512 && lane == 0) p.dbg[it * 8 + 4] = clock64();
                }
            }
    } else if (warp == 2 || warp == 3 || warp == 8 || warp == 9) {
        // ================================================================= splitter (warps 2,3,8,9 = 128 threads)
        const int st_tid = warp < 4 ? tid - 64 : tid - 192;
        unsigned it = 0;
        for (int rd = 0; rd < rounds; ++rd)
            for (int nt = 0; nt < p.n_tiles; ++nt)
                for (int sl = 0; sl < slabs; ++sl, ++it) {
                    const int st = it % ST;
                    const unsigned ph = (it / ST) & 1;
                    mbar_wait(full + st, ph);
                    if (p.dbg && blockIdx.x == 0 && it < 512 && st_tid == 0) p.dbg[it * 8 + 5] = clock64();
                    float4* hi = reinterpret_cast<float4*>(stage_mem + st * SB);
                    float4* lo = reinterpret_cast<float4*>(stage_mem + st * SB + A_BYTES);
                    if constexpr (BF) {
                        // raw slab: 128 rows x 64 bytes, 64B-swizzled by the TMA (16-byte chunk c of row r sits at c ^ ((r>>1)&3));
                        // operand tiles: 128 rows x 32 bytes of bf16, 32B swizzle (chunk c of row r sits at c ^ ((r>>2)&1))
                        unsigned char* ahi = stage_mem + st * SB + A_BYTES;
                        unsigned char* alo = ahi + ABF_BYTES;
#pragma unroll
                        for (int i = st_tid; i < A_BYTES / 16; i += SPLIT_THREADS) {
                            const float4 x = hi[i];
                            const int row = i >> 2, lch = (i & 3) ^ ((row >> 1) & 3);      // logical chunk: fp32 k = 4*lch .. 4*lch+3
                            const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
                            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                            const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - f01.x, x.y - f01.y);
                            const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - f23.x, x.w - f23.y);
                            const unsigned off = (unsigned)row * 32u + ((unsigned)((lch >> 1) ^ ((row >> 2) & 1)) << 4) + (unsigned)(lch & 1) * 8u;
                            uint2 hv, lv;
                            hv.x = *reinterpret_cast<const unsigned*>(&h01); hv.y = *reinterpret_cast<const unsigned*>(&h23);
                            lv.x = *reinterpret_cast<const unsigned*>(&l01); lv.y = *reinterpret_cast<const unsigned*>(&l23);
                            *reinterpret_cast<uint2*>(ahi + off) = hv;
                            *reinterpret_cast<uint2*>(alo + off) = lv;
                        }
                    }
#pragma unroll
                    for (int i = st_tid; EXACT && i < A_BYTES / 16; i += SPLIT_THREADS) {
                        const float4 x = hi[i];
                        float4 h, l;
                        unsigned u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
                        hi[i] = h;
                        lo[i] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA
                    if (p.dbg && blockIdx.x == 0 && it < 512 && st_tid == 0) p.dbg[it * 8 + 6] = clock64();
                    if (p.dbg && blockIdx.x == 0 && it < 512 && warp == 9 && lane == 0) p.dbg[it * 8 + 7] = clock64();   // last splitter warp
                    mbar_arrive(split + st);
                }
    } else if ((warp >= 4 && warp < 8) || warp >= 10) {
        // ================================================================= epilogue: two groups of 4 warps (4-7 and 10-13); warp w may
        // touch TMEM lanes 32*(w%4)..+31, so each group covers all 128 accumulator rows and takes half of the tile's columns.
        // A 16-column chunk is read row-per-thread, turned through a 32x16 shared-memory tile and leaves as 64-byte row pieces
        // (instruction i: lane -> frame 8i + lane/4, 16-byte piece lane%4): whole sectors, 8 rows per warp instruction.
        const int grp = warp >= 10 ? 1 : 0, q = warp & 3;
        float* tt = stage_t + (grp * 4 + q) * 32 * TP;
        const int sub_r = lane >> 2, sub_c = (lane & 3) * 4;
        unsigned tile = 0;
        for (int rd = 0; rd < rounds; ++rd) {
            const int item = (cluster_id + rd * n_clusters) * CS + crank;
            const bool ghost = item >= items;
            const int b = ghost ? 0 : item / m_tiles, t0 = ghost ? p.L : p.t_begin + (item % m_tiles) * BM;
            const int tbase = t0 + q * 32;                                   // first frame of this warp's rows
            for (int nt = 0; nt < p.n_tiles; ++nt, ++tile) {
                const unsigned ab = tile & 1, aph = (tile >> 1) & 1;
                mbar_wait(acc_full + ab, aph);
                tc_fence_after();
                const unsigned taddr = tmem_base + ab * BN + ((unsigned)(q * 32) << 16);
                if (EPI == EPI_GATE) {
                    // tile columns: [0,128) = F of channels 128*nt.., [128,256) = G of the same channels; group g takes 64 channels
                    const float* bt = bias_s + nt * BN;
#pragma unroll 1
                    for (int c = grp * 64; c < grp * 64 + 64; c += 16) {
                        float f[16], g[16];
                        tmem_ld16(taddr + c, f);
                        tmem_ld16(taddr + 128 + c, g);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            f[i] = tanhf(f[i] + bt[c + i]);
                            g[i] = sigmoid_tc(g[i] + bt[128 + c + i]);
                        }
                        const int n_pass = p.out1 ? 3 : 1;                   // z, then (optionally) f and g for the backward
                        for (int ps = 0; ps < n_pass; ++ps) {
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                float4 v;
                                if (ps == 0) v = make_float4(f[i] * g[i], f[i + 1] * g[i + 1], f[i + 2] * g[i + 2], f[i + 3] * g[i + 3]);
                                else if (ps == 1) v = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                                else v = make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]);
                                *reinterpret_cast<float4*>(tt + lane * TP + i) = v;
                            }
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int fr = tbase + 8 * i + sub_r;
                                if (fr < p.L) {
                                    const float4 v = *reinterpret_cast<const float4*>(tt + (8 * i + sub_r) * TP + sub_c);
                                    float* dst = (ps == 0) ? p.out0 + ((size_t)b * p.L + fr) * p.D + nt * 128 + c + sub_c
                                                           : p.out1 + ((size_t)b * p.L + fr) * (2 * p.D) + (ps == 2 ? p.D : 0) + nt * 128 + c + sub_c;
                                    *reinterpret_cast<float4*>(dst) = v;
                                }
                            }
                        }
                    }
                } else if (EPI == EPI_GATE_BWD) {
                    // dz tile (columns = dilation channels nt*256 + c): dF = dz*g*(1-f^2), dG = dz*f*g*(1-g), z = f*g
                    const int n0 = nt * BN;
#pragma unroll 1
                    for (int c = grp * 128; c < grp * 128 + 128; c += 16) {
                        float v[16];
                        tmem_ld16(taddr + c, v);
                        float4 fq[4], gq[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int fr = tbase + 8 * i + sub_r;
                            fq[i] = gq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (fr < p.L) {
                                const float* fg = p.res + ((size_t)b * p.L + fr) * (2 * p.D) + n0 + c + sub_c;
                                fq[i] = __ldg(reinterpret_cast<const float4*>(fg));
                                gq[i] = __ldg(reinterpret_cast<const float4*>(fg + p.D));
                            }
                        }
                        tmem_ld_wait();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(tt + lane * TP + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int fr = tbase + 8 * i + sub_r;
                            if (fr >= p.L) continue;
                            const float4 dz = *reinterpret_cast<const float4*>(tt + (8 * i + sub_r) * TP + sub_c);
                            const float4 f = fq[i], g = gq[i];
                            float* dfg = p.out0 + ((size_t)b * p.L + fr) * (2 * p.D) + n0 + c + sub_c;
                            *reinterpret_cast<float4*>(dfg) = make_float4(dz.x * g.x * (1.f - f.x * f.x), dz.y * g.y * (1.f - f.y * f.y),
                                                                          dz.z * g.z * (1.f - f.z * f.z), dz.w * g.w * (1.f - f.w * f.w));
                            *reinterpret_cast<float4*>(dfg + p.D) = make_float4(dz.x * f.x * g.x * (1.f - g.x), dz.y * f.y * g.y * (1.f - g.y),
                                                                                dz.z * f.z * g.z * (1.f - g.z), dz.w * f.w * g.w * (1.f - g.w));
                            *reinterpret_cast<float4*>(p.out2 + ((size_t)b * p.L + fr) * p.D + n0 + c + sub_c) =
                                make_float4(f.x * g.x, f.y * g.y, f.z * g.z, f.w * g.w);
                        }
                    }
                } else if (EPI == EPI_ADD) {
                    // dh_in tile (columns = residual channels nt*256 + c) = acc + dh_out(t) for t >= id_start
                    const int n0 = nt * BN;
#pragma unroll 1
                    for (int c = grp * 128; c < grp * 128 + 128; c += 16) {
                        float v[16];
                        tmem_ld16(taddr + c, v);
                        float4 x[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int fr = tbase + 8 * i + sub_r;
                            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (fr < p.L && p.res != nullptr && fr >= p.id_start)
                                x[i] = __ldg(reinterpret_cast<const float4*>(p.res + ((size_t)b * p.L + fr) * p.R + n0 + c + sub_c));
                        }
                        tmem_ld_wait();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(tt + lane * TP + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int fr = tbase + 8 * i + sub_r;
                            if (fr >= p.L) continue;
                            float4 o = *reinterpret_cast<const float4*>(tt + (8 * i + sub_r) * TP + sub_c);
                            o.x += x[i].x; o.y += x[i].y; o.z += x[i].z; o.w += x[i].w;
                            *reinterpret_cast<float4*>(p.out0 + ((size_t)b * p.L + fr) * p.R + n0 + c + sub_c) = o;
                        }
                    }
                } else {
                    // tile columns: global output column n = nt*256 + c; n < R residual, else skip channel n - R; group g takes 128
                    const int n0 = nt * BN;
                    const bool is_res = n0 < p.R;
                    const float* bt = bias_s + n0;
                    const int Tsk = p.L - p.skip_start;
#pragma unroll 1
                    for (int c = grp * 128; c < grp * 128 + 128; c += 16) {
                        float v[16];
                        tmem_ld16(taddr + c, v);
                        // the values this lane will add in the coalesced domain (residual h_in(t) or the running skip)
                        float4 x[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int fr = tbase + 8 * i + sub_r;
                            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (fr < p.L) {
                                if (is_res) {
                                    if (fr >= p.in_start)
                                        x[i] = __ldg(reinterpret_cast<const float4*>(p.res + ((size_t)b * p.L + fr) * p.R + n0 + c + sub_c));
                                } else if (!p.skip_init && fr >= p.skip_start) {
                                    x[i] = *reinterpret_cast<const float4*>(p.out1 + ((size_t)b * Tsk + (fr - p.skip_start)) * p.S + (n0 - p.R) + c + sub_c);
                                }
                            }
                        }
                        tmem_ld_wait();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(tt + lane * TP + i) =
                                make_float4(v[i] + bt[c + i], v[i + 1] + bt[c + i + 1], v[i + 2] + bt[c + i + 2], v[i + 3] + bt[c + i + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int fr = tbase + 8 * i + sub_r;
                            if (fr >= p.L) continue;
                            float4 o = *reinterpret_cast<const float4*>(tt + (8 * i + sub_r) * TP + sub_c);
                            o.x += x[i].x; o.y += x[i].y; o.z += x[i].z; o.w += x[i].w;
                            if (is_res)
                                *reinterpret_cast<float4*>(p.out0 + ((size_t)b * p.L + fr) * p.R + n0 + c + sub_c) = o;
                            else if (fr >= p.skip_start)
                                *reinterpret_cast<float4*>(p.out1 + ((size_t)b * Tsk + (fr - p.skip_start)) * p.S + (n0 - p.R) + c + sub_c) = o;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(acc_empty + ab);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                              // no peer may still signal this CTA's barriers after it is gone
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- weight packing
// pass A: rows in tile order (tile p: F channels 128p.., then G channels 128p..), columns kk = j*R + r; hi then lo copy
__global__ void pack_a_kernel(const float* __restrict__ wf, const float* __restrict__ wg, const float* __restrict__ bf,
                              const float* __restrict__ bg, int R, int D, int k, float* __restrict__ wa, float* __restrict__ ba) {
    const int K = k * R, N = 2 * D;
    const long long total = (long long)N * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / K), kk = (int)(i % K);
        const int tile = n / 256, w = n % 256, ch = tile * 128 + (w & 127);
        const bool is_g = w >= 128;
        const int j = kk / R, r = kk % R;
        const float v = (is_g ? wg : wf)[((size_t)ch * R + r) * k + j];
        unsigned u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float hi = __uint_as_float(u);
        wa[i] = hi;
        wa[total + i] = v - hi;
        if (kk == 0) {
            const float* bsrc = is_g ? bg : bf;
            ba[n] = bsrc ? bsrc[ch] : 0.f;
        }
    }
}
// pass B: rows = residual outputs then skip outputs, columns = dilation channel; hi then lo copy
__global__ void pack_b_kernel(const float* __restrict__ wr, const float* __restrict__ ws, const float* __restrict__ br,
                              const float* __restrict__ bs, int R, int D, int S, float* __restrict__ wb, float* __restrict__ bb) {
    const int N = R + S;
    const long long total = (long long)N * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / D), c = (int)(i % D);
        const float v = n < R ? wr[(size_t)n * D + c] : ws[(size_t)(n - R) * D + c];
        unsigned u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float hi = __uint_as_float(u);
        wb[i] = hi;
        wb[total + i] = v - hi;
        if (c == 0) bb[n] = n < R ? (br ? br[n] : 0.f) : (bs ? bs[n - R] : 0.f);
    }
}

// backward dz: rows = dilation channels c, columns k: [0,R) = residual_conv.weight[k][c], [R,R+S) = skip_conv.weight[k-R][c]
__global__ void pack_dz_kernel(const float* __restrict__ wr, const float* __restrict__ ws, int R, int D, int S,
                               float* __restrict__ w) {
    const int K = R + S;
    const long long total = (long long)D * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / K), kk = (int)(i % K);
        const float v = kk < R ? wr[(size_t)kk * D + c] : ws[(size_t)(kk - R) * D + c];
        unsigned u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float hi = __uint_as_float(u);
        w[i] = hi;
        w[total + i] = v - hi;
    }
}
// backward dh_in: rows = residual channels r, columns j*2D + n: [filter;gate].weight[n][r][j]
__global__ void pack_dh_kernel(const float* __restrict__ wf, const float* __restrict__ wg, int R, int D, int k,
                               float* __restrict__ w) {
    const int K = k * 2 * D;
    const long long total = (long long)R * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / K), kk = (int)(i % K);
        const int j = kk / (2 * D), n = kk % (2 * D);
        const float v = n < D ? wf[((size_t)n * R + r) * k + j] : wg[((size_t)(n - D) * R + r) * k + j];
        unsigned u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float hi = __uint_as_float(u);
        w[i] = hi;
        w[total + i] = v - hi;
    }
}

// fp32 (hi | lo) tf32-split pair arrays -> bf16 (hi | lo) pair arrays of the same shape.  hi + lo is the original
// weight exactly (lo = x - rna_tf32(x) is exact in fp32), so x is recovered and re-split for bf16.
__global__ void convert_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = src[i] + src[n + i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        dst[i] = h;
        dst[n + i] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}

// ---------------------------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// activations (B, L, C) fp32: dims {C, L - origin, B}, box {32, 128, 1}; frames left of `origin` are out of bounds -> zeros
static int make_act_map(CUtensorMap* m, const float* base, int B, int L, int C, int origin) {
    EncodeTiledFn fn = encode_fn();
    WN_REQUIRE(fn, WN_E_UNSUPP, "cuTensorMapEncodeTiled is not available from this driver");
    WN_REQUIRE(L - origin >= 1, WN_E_BADARG, "empty activation range");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)(L - origin), (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)L * C * 4};
    cuuint32_t box[3] = {BK, BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(base + (size_t)origin * C), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, (BK * 4 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : (BK * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WN_REQUIRE(r == CUDA_SUCCESS, WN_E_UNSUPP, "cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
    return 0;
}
// weights (rows, K) fp32 K-major: dims {K, rows}, box {32, 256}
// `col0`: first K column (element offset into every row); bf16 = true: the array holds bf16 pairs (32-byte slab rows)
static int make_w_map(CUtensorMap* m, const void* base, int rows, int K, int col0 = 0, bool bf16 = false) {
    EncodeTiledFn fn = encode_fn();
    WN_REQUIRE(fn, WN_E_UNSUPP, "cuTensorMapEncodeTiled is not available from this driver");
    const int es = bf16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * es};
    cuuint32_t box[2] = {BK, BN / CS};          // one CTA's share of a slab; the multicast assembles the rest
    cuuint32_t estr[2] = {1, 1};
    const int row_bytes = BK * es;
    CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    (void*)((const unsigned char*)base + (size_t)col0 * es), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WN_REQUIRE(r == CUDA_SUCCESS, WN_E_UNSUPP, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    return 0;
}

static size_t tc_smem_bytes(int n_total, int prec) {
    const size_t ring = prec == PREC_BF16X2 ? (size_t)STAGES_BF * STAGE_BYTES_BF : (size_t)STAGES * STAGE_BYTES;
    return 1024 + ring + 512 + sizeof(float) * ((n_total + 3) & ~3) + sizeof(float) * 8 * 32 * TP;
}

template <int EPI, int PREC>
static int launch_tc(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mW, const TcParams& p, cudaStream_t st) {
    int dev = 0, sms = 0;
    WN_CUDA(cudaGetDevice(&dev));
    WN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = tc_smem_bytes(p.n_total, PREC);
    WN_CUDA(cudaFuncSetAttribute(frames_gemm_tc<EPI, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int items = p.B * ((p.L - p.t_begin + BM - 1) / BM);
    int grid = ((items + CS - 1) / CS) * CS;
    const int max_grid = (sms / CS) * CS;
    if (grid > max_grid) grid = max_grid;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WN_CUDA(cudaLaunchKernelEx(&cfg, frames_gemm_tc<EPI, PREC>, mA, mA2, mW, p));
    WN_CUDA(cudaGetLastError());
    return 0;
}
template <int EPI>
static int launch_tc_prec(int prec, const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mW, const TcParams& p,
                          cudaStream_t st) {
    if (prec == PREC_BF16X2) return launch_tc<EPI, PREC_BF16X2>(mA, mA2, mW, p, st);
    if (prec == PREC_TF32X1) return launch_tc<EPI, PREC_TF32X1>(mA, mA2, mW, p, st);
    return launch_tc<EPI, PREC_TF32X3>(mA, mA2, mW, p, st);
}


// ============================================================================================== weight gradients
// dW[n][c] = sum over sequences b and frames t of g[b][t][n] * x[b][t][c]  (what wn_wgrad computes on the FMA pipe,
// wgrad.cu) on the tensor cores.  The contraction runs over frames, the SLOW axis of both row-major operands, so the
// tiles arrive "MN-major"; instead of MN-major descriptors the splitter -- which has to rewrite every element as a
// bf16 (hi, lo) pair anyway -- writes the operand tiles TRANSPOSED, i.e. in the same K-major, 32-byte-swizzled form the
// block kernels use.  One CTA = one 128-row tile of n (UMMA M) x all C = 256 columns (UMMA N) x one range of K slabs
// (16 frames each); its fp32 partial goes to the split-frames workspace and wgrad_tc_reduce_kernel adds the partials.
//   warp 0      TMA producer: raw fp32 tiles g[16 frames][128 ch] and x[16 frames][256 ch] (no swizzle), 4-stage ring
//   warp 1      TMEM (256 columns), tcgen05.mma kind::f16: hi*hi + lo*hi + hi*lo per slab
//   warps 2-9   splitter: thread = (channel, 8 frames): 8 conflict-free LDS.32 down a column, bf16 hi/lo split,
//               one 16-byte store per operand tile row chunk
//   warps 2-9   after the last slab: TMEM -> registers -> workspace
constexpr int WG_THREADS = 320;
constexpr int WG_SPLIT_THREADS = 256;
constexpr int WG_STAGES = 4;
constexpr int WG_RAW_A = BK * BM * 4;            // 8 KB   [16 frames][128 ch] fp32
constexpr int WG_RAW_B = BK * BN * 4;            // 16 KB  [16 frames][256 ch] fp32
constexpr int WG_STAGE_BYTES = WG_RAW_A + WG_RAW_B + 2 * ABF_BYTES + 2 * WBF_BYTES;      // 48 KB

struct WgTcParams {
    int slabs_per_seq, total_slabs, slabs_per_split, m_tiles;
    int N, C;
    float* work;                  // [splits][N][C]
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX, const WgTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + WG_STAGES * WG_STAGE_BYTES);
    unsigned long long* full = bars;                       // [WG_STAGES] TMA landed
    unsigned long long* split = bars + WG_STAGES;          // [WG_STAGES] operand tiles written
    unsigned long long* empty = bars + 2 * WG_STAGES;      // [WG_STAGES] MMAs of the stage retired
    unsigned long long* acc_full = bars + 3 * WG_STAGES;   // [1]
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 3 * WG_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_tile = blockIdx.x % p.m_tiles, sp = blockIdx.x / p.m_tiles;
    const int s_beg = sp * p.slabs_per_split;
    const int s_end = (s_beg + p.slabs_per_split < p.total_slabs) ? s_beg + p.slabs_per_split : p.total_slabs;
    const int n_slabs = s_end > s_beg ? s_end - s_beg : 0;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(split + i, WG_SPLIT_THREADS); mbar_init(empty + i, 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapG) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < n_slabs; ++i) {
                const int st = i % WG_STAGES;
                const unsigned ph = (i / WG_STAGES) & 1;
                mbar_wait(empty + st, ph ^ 1);
                unsigned char* sm = base + st * WG_STAGE_BYTES;
                const int s = s_beg + i, b = s / p.slabs_per_seq, t0 = (s % p.slabs_per_seq) * BK;
                mbar_expect_tx(full + st, WG_RAW_A + WG_RAW_B);
                tma_load_3d(sm, &mapG, m_tile * BM, t0, b, full + st);           // frames past the sequence end: zero fill
                tma_load_3d(sm + WG_RAW_A, &mapX, 0, t0, b, full + st);
            }
        }
    } else if (warp == 1) {
        constexpr unsigned idesc = make_idesc(true);
        for (int i = 0; i < n_slabs; ++i) {
            const int st = i % WG_STAGES;
            const unsigned ph = (i / WG_STAGES) & 1;
            mbar_wait(split + st, ph);
            tc_fence_after();
            if (elect_one()) {
                const unsigned sa = s32(base + st * WG_STAGE_BYTES) + WG_RAW_A + WG_RAW_B;
                const unsigned long long a_hi = smem_desc<32>(sa), a_lo = smem_desc<32>(sa + ABF_BYTES);
                const unsigned long long b_hi = smem_desc<32>(sa + 2 * ABF_BYTES), b_lo = smem_desc<32>(sa + 2 * ABF_BYTES + WBF_BYTES);
                umma_f16(tmem_base, a_hi, b_hi, idesc, i != 0);
                umma_f16(tmem_base, a_lo, b_hi, idesc, 1);
                umma_f16(tmem_base, a_hi, b_lo, idesc, 1);
                umma_commit(empty + st);
                if (i == n_slabs - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- splitter (256 threads), then epilogue
        const int stid = tid - 64;
        for (int i = 0; i < n_slabs; ++i) {
            const int st = i % WG_STAGES;
            const unsigned ph = (i / WG_STAGES) & 1;
            mbar_wait(full + st, ph);
            unsigned char* sm = base + st * WG_STAGE_BYTES;
            const float* rawA = reinterpret_cast<const float*>(sm);
            const float* rawB = reinterpret_cast<const float*>(sm + WG_RAW_A);
            unsigned char* tiles = sm + WG_RAW_A + WG_RAW_B;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                // groups 0..255: A (128 channels x 2 halves of 8 frames); groups 256..767: B (256 channels x 2 halves)
                const int g = stid + WG_SPLIT_THREADS * j;
                const bool isA = g < 2 * BM;
                const int gg = isA ? g : g - 2 * BM;
                const int nch = isA ? BM : BN;
                const int ch = gg % nch, half = gg / nch;
                const float* src = (isA ? rawA : rawB) + (half * 8) * nch + ch;
                float x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = src[k * nch];
                unsigned hv[4], lv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
                    const float2 hf = __bfloat1622float2(h);
                    const __nv_bfloat162 l = __floats2bfloat162_rn(x[2 * k] - hf.x, x[2 * k + 1] - hf.y);
                    hv[k] = *reinterpret_cast<const unsigned*>(&h);
                    lv[k] = *reinterpret_cast<const unsigned*>(&l);
                }
                // K-major tile, 32-byte rows (16 bf16), 32B swizzle: 16-byte chunk `half` of row `ch` sits at half ^ ((ch>>2)&1)
                const unsigned off = (unsigned)ch * 32u + ((unsigned)(half ^ ((ch >> 2) & 1)) << 4);
                unsigned char* hi_t = tiles + (isA ? 0 : 2 * ABF_BYTES);
                unsigned char* lo_t = hi_t + (isA ? ABF_BYTES : WBF_BYTES);
                *reinterpret_cast<uint4*>(hi_t + off) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<uint4*>(lo_t + off) = make_uint4(lv[0], lv[1], lv[2], lv[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(split + st);
        }
        // epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (= rows of the n tile); warps 2-5 take columns [0,128), 6-9 [128,256)
        const int q = warp & 3, grp = (warp - 2) >> 2;
        const int n = m_tile * BM + q * 32 + lane;
        float* out = p.work + ((size_t)sp * p.N + n) * p.C;
        if (n_slabs > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16);
#pragma unroll 1
            for (int c = grp * 128; c < grp * 128 + 128; c += 16) {
                float v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(out + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        } else {
            for (int c = grp * 128; c < grp * 128 + 128; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ work, float* __restrict__ dw, int N, int C, int splits,
                                       long long n_stride, long long c_stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += work[(size_t)k * N * C + idx];
    const int n = idx / C, c = idx - n * C;
    dw[n * n_stride + c * c_stride] = s;
}

// raw (B, rows, ld) fp32 rows as a 3D map {channels, rows, B}, box {box_ch, 16, 1}, no swizzle, zero fill past `rows`
static int make_rows_map(CUtensorMap* m, const float* base, int channels, int rows, int B, int ld, long long seq, int box_ch) {
    EncodeTiledFn fn = encode_fn();
    WN_REQUIRE(fn, WN_E_UNSUPP, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {(cuuint64_t)channels, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)seq * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_ch, BK, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WN_REQUIRE(r == CUDA_SUCCESS, WN_E_UNSUPP, "cuTensorMapEncodeTiled(wgrad rows) failed with %d", (int)r);
    return 0;
}

}  // namespace tc
}  // namespace wn

using namespace wn;

static long long* g_tc_dbg = nullptr;
extern "C" int wn_tc_read_trace(long long* host_out, int n) {
    WN_REQUIRE(g_tc_dbg && host_out && n > 0 && n <= 512 * 8, WN_E_STATE, "wn_tc_read_trace: tracing is off (WN_TC_TRACE=1) or bad n");
    WN_CUDA(cudaDeviceSynchronize());
    WN_CUDA(cudaMemcpy(host_out, g_tc_dbg, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int wn_tc_supported(int R, int D, int S, int k) {
    return (R % 256 == 0) && (S % 256 == 0) && (D % 128 == 0) && k >= 1 && (R + S) <= 2048 && 2 * D <= 2048;
}

extern "C" int wn_tc_pack_block_weights(const float* d_wf, const float* d_wg, const float* d_bf, const float* d_bg,
                                        const float* d_wr, const float* d_ws, const float* d_br, const float* d_bs, int R,
                                        int D, int S, int k, float* d_wa, float* d_ba, float* d_wb, float* d_bb, void* stream) {
    WN_REQUIRE(d_wf && d_wg && d_wr && d_ws && d_wa && d_ba && d_wb && d_bb, WN_E_BADARG, "wn_tc_pack_block_weights: null pointer");
    WN_REQUIRE(wn_tc_supported(R, D, S, k), WN_E_UNSUPP, "wn_tc_pack_block_weights: shape R=%d D=%d S=%d not supported", R, D, S);
    cudaStream_t st = (cudaStream_t)stream;
    tc::pack_a_kernel<<<1024, 256, 0, st>>>(d_wf, d_wg, d_bf, d_bg, R, D, k, d_wa, d_ba);
    tc::pack_b_kernel<<<512, 256, 0, st>>>(d_wr, d_ws, d_br, d_bs, R, D, S, d_wb, d_bb);
    WN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wn_tc_block_fwd(const wn_tc_block_args* a, void* stream) {
    WN_REQUIRE(a, WN_E_BADARG, "wn_tc_block_fwd: null args");
    WN_REQUIRE(a->d_h_in && a->d_h_out && a->d_skip && a->d_z && a->d_wa && a->d_ba && a->d_wb && a->d_bb, WN_E_BADARG,
               "wn_tc_block_fwd: null pointer");
    WN_REQUIRE(wn_tc_supported(a->R, a->D, a->S, a->k), WN_E_UNSUPP, "wn_tc_block_fwd: shape not supported by the tensor-core path");
    WN_REQUIRE(a->B > 0 && a->L > 0 && a->dilation >= 1 && a->in_start >= 0 && a->out_start >= a->in_start &&
                   a->out_start < a->L && a->skip_start >= a->out_start && a->skip_start < a->L,
               WN_E_BADARG, "wn_tc_block_fwd: bad frame ranges");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap mA, mWa, mZ, mWb;
    if (int rc = tc::make_act_map(&mA, a->d_h_in, a->B, a->L, a->R, a->in_start)) return rc;
    const int prec = a->fast_tf32;            // 0: 3xTF32, 1: single TF32, 2: bf16 pairs (d_wa / d_wb from wn_tc_convert_weights_bf16)
    WN_REQUIRE(prec >= 0 && prec <= 2, WN_E_BADARG, "wn_tc_block_fwd: fast_tf32 (precision mode) must be 0, 1 or 2");
    const bool bf = prec == tc::PREC_BF16X2;
    if (int rc = tc::make_w_map(&mWa, a->d_wa, 2 * 2 * a->D, a->k * a->R, 0, bf)) return rc;
    if (int rc = tc::make_act_map(&mZ, a->d_z, a->B, a->L, a->D, a->out_start)) return rc;
    if (int rc = tc::make_w_map(&mWb, a->d_wb, 2 * (a->R + a->S), a->D, 0, bf)) return rc;
    tc::TcParams p;
    memset(&p, 0, sizeof(p));
    if (getenv("WN_TC_TRACE")) {
        if (!g_tc_dbg) WN_CUDA(cudaMalloc(&g_tc_dbg, sizeof(long long) * 512 * 8));
        p.dbg = g_tc_dbg;
    }
    p.B = a->B; p.L = a->L; p.t_begin = a->out_start;
    p.D = a->D; p.R = a->R; p.S = a->S; p.in_start = a->in_start; p.skip_start = a->skip_start; p.skip_init = a->skip_init;
    // pass A: conv taps + gate
    p.taps = a->k; p.dil = a->dilation; p.C = a->R; p.a_origin = a->in_start;
    p.n_total = 2 * a->D; p.n_tiles = p.n_total / tc::BN;
    p.bias = a->d_ba; p.out0 = a->d_z; p.out1 = a->d_fg_save; p.res = nullptr;
    if (int rc = tc::launch_tc_prec<tc::EPI_GATE>(prec, mA, mA, mWa, p, st)) return rc;
    // pass B: residual + skip 1x1
    if (const char* which = getenv("WN_TC_TRACE_PASS")) { if (which[0] == 'A') p.dbg = nullptr; }     // keep pass A's stamps
    p.taps = 1; p.dil = 0; p.C = a->D; p.a_origin = a->out_start;
    p.n_total = a->R + a->S; p.n_tiles = p.n_total / tc::BN;
    p.bias = a->d_bb; p.out0 = a->d_h_out; p.out1 = a->d_skip; p.res = a->d_h_in;
    return tc::launch_tc_prec<tc::EPI_RES_SKIP>(prec, mZ, mZ, mWb, p, st);
}

extern "C" int wn_tc_bwd_supported(int R, int D, int S, int k) {
    return (R % 256 == 0) && (S % 256 == 0) && (D % 256 == 0) && k >= 1 && (R + S) <= 2048 && k * 2 * D <= 4096;
}

extern "C" int wn_tc_pack_block_bwd_weights(const float* d_wf, const float* d_wg, const float* d_wr, const float* d_ws, int R,
                                            int D, int S, int k, float* d_wdz, float* d_wdh, void* stream) {
    WN_REQUIRE(d_wf && d_wg && d_wr && d_ws && d_wdz && d_wdh, WN_E_BADARG, "wn_tc_pack_block_bwd_weights: null pointer");
    WN_REQUIRE(wn_tc_bwd_supported(R, D, S, k), WN_E_UNSUPP, "wn_tc_pack_block_bwd_weights: shape not supported");
    cudaStream_t st = (cudaStream_t)stream;
    tc::pack_dz_kernel<<<512, 256, 0, st>>>(d_wr, d_ws, R, D, S, d_wdz);
    tc::pack_dh_kernel<<<1024, 256, 0, st>>>(d_wf, d_wg, R, D, k, d_wdh);
    WN_CUDA(cudaGetLastError());
    return 0;
}

// Tensor-core form of wn_block_bwd_data (same arguments; d_wrs_rows / d_wfg_bwd are replaced by the packed, pre-split
// d_wdz [2][D][R+S] and d_wdh [2][R][k*2D] of wn_tc_pack_block_bwd_weights).
extern "C" int wn_tc_block_bwd_data(const wn_block_bwd_args* a, const float* d_wdz, const float* d_wdh, void* stream) {
    return wn_tc_block_bwd_data_prec(a, d_wdz, d_wdh, 0, stream);
}

// precision: 0 = 3xTF32 on the fp32 pair arrays, 2 = bf16 pairs (arrays converted by wn_tc_convert_weights_bf16)
extern "C" int wn_tc_block_bwd_data_prec(const wn_block_bwd_args* a, const void* d_wdz, const void* d_wdh, int precision,
                                         void* stream) {
    WN_REQUIRE(a && d_wdz && d_wdh, WN_E_BADARG, "wn_tc_block_bwd_data: null args");
    WN_REQUIRE(precision == 0 || precision == 2, WN_E_BADARG, "wn_tc_block_bwd_data: precision must be 0 (3xTF32) or 2 (bf16 pairs)");
    const bool bf = precision == 2;
    WN_REQUIRE(a->d_dskip && a->d_fg && a->d_dfg && a->d_z && a->d_dh_in, WN_E_BADARG, "wn_tc_block_bwd_data: null pointer");
    WN_REQUIRE(wn_tc_bwd_supported(a->R, a->D, a->S, a->k), WN_E_UNSUPP, "wn_tc_block_bwd_data: shape not supported");
    WN_REQUIRE(a->gz >= a->out_start && a->gz < a->L && a->gs_in >= a->in_start && a->gs_in <= a->gz && a->ds_start >= a->out_start &&
                   a->ds_start < a->L,
               WN_E_BADARG, "wn_tc_block_bwd_data: bad gradient frame ranges");
    cudaStream_t st = (cudaStream_t)stream;
    const int R = a->R, D = a->D, S = a->S, L = a->L, B = a->B;
    const bool have_dh = a->d_dh_out != nullptr && a->gs_out < L;
    CUtensorMap mDh, mDs, mW1, mDfg, mW2;
    if (int rc = tc::make_act_map(&mDs, a->d_dskip, B, L - a->ds_start, S, 0)) return rc;          // (B, L-ds_start, S): own frame axis
    if (have_dh) { if (int rc = tc::make_act_map(&mDh, a->d_dh_out, B, L, R, a->gs_out)) return rc; }
    else mDh = mDs;
    if (int rc = tc::make_w_map(&mW1, d_wdz, 2 * D, R + S, have_dh ? 0 : R, bf)) return rc;       // without dh_out: start at column R
    tc::TcParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.L = L; p.D = D; p.R = R; p.S = S;
    // ---- dz + gate backward for frames [gz, L)
    p.t_begin = a->gz;
    p.taps = have_dh ? 1 : 0; p.dil = 0; p.C = have_dh ? R : 0; p.a_origin = a->gs_out;
    p.C2 = S; p.a2_origin = a->ds_start;
    p.n_total = D; p.n_tiles = D / tc::BN;
    p.bias = nullptr; p.out0 = a->d_dfg; p.out2 = a->d_z; p.res = a->d_fg;
    if (int rc = tc::launch_tc_prec<tc::EPI_GATE_BWD>(precision, mDh, mDs, mW1, p, st)) return rc;
    // ---- dh_in = dh_out (identity) + anti-causal taps of dfg, for frames [gs_in, L)
    if (int rc = tc::make_act_map(&mDfg, a->d_dfg, B, L, 2 * D, a->gz)) return rc;
    if (int rc = tc::make_w_map(&mW2, d_wdh, 2 * R, a->k * 2 * D, 0, bf)) return rc;
    p.t_begin = a->gs_in;
    p.taps = a->k; p.dil = -a->dilation; p.C = 2 * D; p.a_origin = a->gz;
    p.C2 = 0; p.a2_origin = 0;
    p.n_total = R; p.n_tiles = R / tc::BN;
    p.out0 = a->d_dh_in; p.out2 = nullptr; p.res = a->d_dh_out;
    p.id_start = a->gs_out > a->out_start ? a->gs_out : a->out_start;
    return tc::launch_tc_prec<tc::EPI_ADD>(precision, mDfg, mDfg, mW2, p, st);
}

extern "C" int wn_tc_convert_weights_bf16(const float* d_pairs, void* d_out, long long n_per_half, void* stream) {
    WN_REQUIRE(d_pairs && d_out && n_per_half > 0, WN_E_BADARG, "wn_tc_convert_weights_bf16: null pointer or empty array");
    tc::convert_bf16_kernel<<<512, 256, 0, (cudaStream_t)stream>>>(d_pairs, (__nv_bfloat16*)d_out, n_per_half);
    WN_CUDA(cudaGetLastError());
    return 0;
}

// Tensor-core form of wn_wgrad (same argument block and workspace): C must be 256 and N a multiple of 128, pitches and
// strides multiples of 4 floats, pointers 16-byte aligned; bf16-pair split, fp32 accumulation.
extern "C" int wn_tc_wgrad_supported(int N, int C) { return C == tc::BN && N > 0 && N % tc::BM == 0; }

extern "C" int wn_tc_wgrad(const wn_wgrad_args* a, void* stream) {
    WN_REQUIRE(a != nullptr, WN_E_BADARG, "wn_tc_wgrad: null argument block");
    WN_REQUIRE(wn_tc_wgrad_supported(a->N, a->C), WN_E_UNSUPP, "wn_tc_wgrad: needs C == 256 and N %% 128 == 0 (got N=%d C=%d)", a->N, a->C);
    WN_REQUIRE(a->B > 0 && a->rows >= 1, WN_E_BADARG, "wn_tc_wgrad: bad sizes B=%d rows=%d", a->B, a->rows);
    WN_REQUIRE(a->d_g && a->d_x && a->d_dw && a->d_work, WN_E_BADARG, "wn_tc_wgrad: null device pointer");
    WN_REQUIRE(a->ldg >= a->N && a->ldx >= a->C && a->ldg % 4 == 0 && a->ldx % 4 == 0 && a->g_seq_stride % 4 == 0 &&
                   a->x_seq_stride % 4 == 0 && (uintptr_t)a->d_g % 16 == 0 && (uintptr_t)a->d_x % 16 == 0,
               WN_E_BADARG, "wn_tc_wgrad: pitches / strides must be multiples of 4 floats and pointers 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    WN_CUDA(cudaGetDevice(&dev));
    WN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    tc::WgTcParams p;
    p.N = a->N; p.C = a->C; p.work = a->d_work;
    p.m_tiles = a->N / tc::BM;
    p.slabs_per_seq = (a->rows + tc::BK - 1) / tc::BK;
    const long long total = (long long)a->B * p.slabs_per_seq;
    WN_REQUIRE(total < (1ll << 30), WN_E_UNSUPP, "wn_tc_wgrad: too many frames");
    p.total_slabs = (int)total;
    // the workspace holds wn_wgrad_workspace_bytes(N, C) = ceil(296 / (N/128 * C/128)) partials of N x C
    const int max_splits = (296 + 2 * p.m_tiles - 1) / (2 * p.m_tiles);
    int splits = sms / p.m_tiles;
    if (splits > max_splits) splits = max_splits;
    if (splits > p.total_slabs / 4) splits = p.total_slabs / 4;
    if (splits < 1) splits = 1;
    p.slabs_per_split = (p.total_slabs + splits - 1) / splits;
    splits = (p.total_slabs + p.slabs_per_split - 1) / p.slabs_per_split;
    CUtensorMap mG, mX;
    if (int rc = tc::make_rows_map(&mG, a->d_g, a->N, a->rows, a->B, a->ldg, a->g_seq_stride, tc::BM)) return rc;
    if (int rc = tc::make_rows_map(&mX, a->d_x, a->C, a->rows, a->B, a->ldx, a->x_seq_stride, tc::BN)) return rc;
    const size_t smem = 1024 + (size_t)tc::WG_STAGES * tc::WG_STAGE_BYTES + 256;
    WN_CUDA(cudaFuncSetAttribute(tc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::wgrad_tc_kernel<<<p.m_tiles * splits, tc::WG_THREADS, smem, st>>>(mG, mX, p);
    WN_CUDA(cudaGetLastError());
    const int total_out = a->N * a->C;
    tc::wgrad_tc_reduce_kernel<<<(total_out + 255) / 256, 256, 0, st>>>(a->d_work, a->d_dw, a->N, a->C, splits, a->dw_n_stride,
                                                                         a->dw_c_stride);
    WN_CUDA(cudaGetLastError());
    return 0;
}
