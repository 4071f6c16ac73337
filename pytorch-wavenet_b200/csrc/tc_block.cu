// tc_block.cu -- the training-time residual block (reference wavenet_model.py:142-165) as ONE tensor-core kernel per layer.
//
//   z[t]     = tanh(Wf0 h[t-d] + Wf1 h[t] + bf) * sigmoid(Wg0 h[t-d] + Wg1 h[t] + bg)          (h[t] = 0 left of in_start)
//   h_out[t] = Wr z[t] + br + h[t]                                                             t in [out_start, L)
//   skip[t]  (+)= Ws z[t] + bs                                                                 t in [skip_start, L)
//
// Numerics: fp32-class through bf16 PAIRS -- every operand x is carried as hi = bf16(x), lo = bf16(x - hi) and a product is
// hi*hi + lo*hi + hi*lo on tcgen05 kind::f16 with fp32 accumulation in tensor memory (16 mantissa bits per operand; measured
// 6.5e-6 on the logits of the 50-layer cfg-2/3 net, tools/split_sim.py; the parity bar is 1e-4).
//
// Data layout ("chunked pair layout"): an activation tensor of C channels over L frames is stored as
//   [sequence b][plane: hi, lo][channel chunk c/8][frame t][8 channels] bf16            (same bytes as fp32 (B,L,C))
// so that  (1) a TMA box {128 frames, 4 chunks, 2 planes} lands in shared memory exactly as the SWIZZLE_NONE K-major
// core-matrix image tcgen05 consumes ([k-chunk][row][16 B], LBO = 2048, SBO = 128): no splitter, no swizzle, no conflicts;
//          (2) an epilogue thread (= one frame, tcgen05.ld 32x32b) reads/writes 16-byte pieces that are CONTIGUOUS across the
// lanes of its warp (consecutive frames): fully coalesced global accesses with no shared-memory transpose;
//          (3) the same image is the MN-major operand of the weight-gradient GEMM (contraction over frames).
// skip is [b][channel chunk c/4][frame][4 channels] fp32 for the same reason.  Weights are pre-split and pre-tiled into the
// shared-memory image of every (n-tile, k-slab, CTA half), so a weight slab is one contiguous 16 KB TMA box.
//
// Kernel: clusters of 2 CTAs (one per SM of a TPC), cta_group::2 MMAs M256 N256 K16: CTA r owns frames t0+128r..+127 of a
// 256-frame item (its A rows, its 128 accumulator lanes) and stages half of every weight slab (its 128 of the 256 B rows).
//   warp 0      TMA producer (both CTAs): ring of six 16 KB slots; a k-slab of pass A = one activation slot (32 channels of
//               one tap, hi+lo) + one weight slot; pass B = weight slots only (its A operand z is resident)
//   warp 1      tensor memory (512 columns = two 256-column accumulators) and, in the leader CTA, the MMA issue loop
//   warps 2-9   epilogue: two groups x four TMEM lane quadrants
// Per item the MMA warp runs  A0 -> acc0, A1 -> acc1  (filter|gate of channels 0..127 / 128..255, K = 2 taps x 256),
// B0 -> acc0 (residual), B1 -> acc1 (skip; skipped for items left of skip_start).  The gate epilogue writes z as a bf16 pair
// image into shared memory (128 KB: the A operand of pass B never leaves the SM) and signals it per 32-channel slab, so pass
// B starts on the first slabs while the rest of the gate is still being computed.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace wn {
namespace tb {
using namespace px;

constexpr int BM = 128;                   // frames per CTA
constexpr int PM = 256;                   // frames per item (CTA pair)
constexpr int SLOT = 16384;               // one ring slot
constexpr int NSLOT = 6;
constexpr int NTHREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr unsigned LBO = BM * 16, SBO = 128;

// Shapes and operand precision.  PAIR: every MMA operand is a bf16 (hi, lo) pair and a product costs three MMAs
// (fp32-class: the parity path).  !PAIR: single-pass bf16 operands (the hi planes only) with fp32 accumulation -- the
// "bf16 training" mode of BASELINE.json configs[4]; the residual stream h and skip stay fp32-class (h is still stored and
// updated as a pair, skip as fp32), only the matrix products see bf16.
template <int CH_, bool PAIR_>
struct Cfg {
    static constexpr int CH = CH_;
    static constexpr bool PAIR = PAIR_;
    static constexpr int PLANES = PAIR ? 2 : 1;              // planes of an MMA operand image
    static constexpr int KC = PAIR ? 4 : 8;                  // 8-channel chunks per k-slab slot: [plane][chunk][row 128][16 B] = 16 KB
    static constexpr int KS = KC * 8;                        // channels per k-slab
    static constexpr int SLABS_A = 2 * CH / KS;              // k-slabs per pass-A n-tile (2 taps)
    static constexpr int SLABS_B = CH / KS;
    static constexpr int NT_A = 2 * CH / 256;                // pass-A n-tiles: [tanh | sigmoid] pre-activations of 128 channels each
    static constexpr int NT_R = CH / 256;                    // pass-B n-tiles: residual, then as many skip tiles
    static constexpr int NT_B = 2 * CH / 256;
    static constexpr int ZPLANE = (CH / 8) * BM * 16;        // one plane of the resident z image
    static constexpr int ZBYTES = PLANES * ZPLANE;
    static constexpr int NZ = CH / KS;                       // z slabs = z_ready barriers
    static constexpr int WROWS_A = NT_A * SLABS_A * 2 * 8;   // 2 KB rows of one layer's packed pass-A weights
    static constexpr int WROWS_LAYER = WROWS_A + NT_B * SLABS_B * 2 * 8;
    static constexpr size_t W_LAYER_BYTES = (size_t)WROWS_LAYER * 2048;
    static constexpr size_t SMEM_BYTES = 128 + ZBYTES + NSLOT * SLOT + 256;
    static_assert(ZBYTES <= 131072, "the z image must fit beside the ring");
    static_assert(NZ <= 8, "z_ready barriers");
};

// One layer of a launch.  A single-layer launch carries it as a kernel parameter; the whole-stack launch reads an array of
// them from global memory (the tensor map must then be 64-byte aligned there).
struct alignas(128) LayerDesc {
    CUtensorMap mapH;              // this layer's input pair tensor, origin = in_start
    int t_begin, in_start, dil, skip_init;
    int tiles_per_seq, n_items, item_base, w_row0;     // items of this layer are global items [item_base, item_base + n_items)
    const float* bias;             // [bf CH | bg CH | br CH | bs CH]
    const uint4* h_in;             // chunked pair (B, 2, CH/8, L, 8) bf16, viewed as 16-byte pieces
    uint4* h_out;
    float4* fg_save;               // optional chunked (B, 2CH/4, L, 4) fp32: tanh outputs in chunks [0,CH/4), sigmoid after
    int war_layer;                 // >= 0: every item of that earlier layer must be complete before this layer writes h_out
    int pad[7];
};
static_assert(sizeof(LayerDesc) % 128 == 0, "LayerDesc array elements must keep the tensor map aligned");

struct BlockParams {
    int B, L, skip_start;
    int n_layers, total_items;
    float4* skip;                  // chunked (B, CH/4, L - skip_start, 4) fp32
    const LayerDesc* layers;       // whole-stack launch: [n_layers] in global memory
    unsigned* item_done;           // whole-stack launch: [total_items] arrival counters (2 CTAs x 8 epilogue warps = 16 when complete)
    unsigned* layer_done;          //                     [n_layers] completed items per layer
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
    unsigned spins = 0;
    while (ld_acquire_gpu(p) < target) {
        __nanosleep(64);
        if (++spins > (1u << 24)) asm volatile("trap;");          // a dependency that never completes must not hang the GPU
    }
}

// MULTI = false: one layer (`single`), items are independent.  MULTI = true: ALL layers of a forward in one persistent launch:
// the global item list is layer-major and dealt round-robin to the clusters, an item waits (in its producer) for the items of the
// previous layer that wrote the frames it reads, and announces itself when its epilogues have stored -- no launch gaps and no
// idle tail between layers (a layer of cfg 3 is 456 items for 74 clusters: 6.16 rounds that cost 7 as separate launches).
template <typename C, bool MULTI>
__global__ void __launch_bounds__(NTHREADS, 1)
block_fused_kernel(const __grid_constant__ LayerDesc single, const __grid_constant__ CUtensorMap mapW, const BlockParams p) {
    constexpr int CH = C::CH;
    auto LD = [&](int l) -> const LayerDesc& { return MULTI ? p.layers[l] : single; };
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    unsigned char* zbuf = base;
    unsigned char* ring = base + C::ZBYTES;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + NSLOT * SLOT);
    unsigned long long* full = bars;                   // [NSLOT]  leader: both CTAs' bytes of the slot have landed
    unsigned long long* empty = bars + NSLOT;          // [NSLOT]  per CTA: the MMAs reading the slot have retired
    unsigned long long* acc_full = bars + 2 * NSLOT;   // [2]      per CTA: accumulator complete
    unsigned long long* acc_empty = acc_full + 2;      // [2]      leader: both CTAs' epilogues are done with the accumulator
    unsigned long long* z_ready = acc_empty + 2;       // [8]      leader: both CTAs wrote z slab s of the current item
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(z_ready + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned rank = cluster_rank();
    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 2 * EPI_WARPS); }
        for (int i = 0; i < 8; ++i) mbar_init(z_ready + i, 2 * EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    }
    if (warp == 1) tmem2_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer (one lane, both CTAs)
        if (elect_one()) {
            unsigned it = 0;
            auto acquire = [&](unsigned& bar_addr) -> unsigned char* {
                const unsigned s = it % NSLOT, ph = (it / NSLOT) & 1;
                mbar_wait(empty + s, ph ^ 1);
                if (rank == 0) mbar_expect_tx(full + s, 2 * SLOT);
                bar_addr = mapa(s32(full + s), 0);
                ++it;
                return ring + s * SLOT;
            };
            int l = 0;
            for (int n = cluster_id; n < p.total_items; n += n_clusters) {
                while (n >= LD(l).item_base + LD(l).n_items) ++l;
                const LayerDesc& Ld = LD(l);
                const int item = n - Ld.item_base, dil = Ld.dil, w_row0 = Ld.w_row0;
                const int b = item / Ld.tiles_per_seq, t0 = Ld.t_begin + (item % Ld.tiles_per_seq) * PM;
                const bool need_skip = t0 + PM > p.skip_start;
                const int tc = t0 + (int)rank * BM - Ld.in_start;          // this CTA's first frame, relative to the map origin
                const CUtensorMap* mapH = &Ld.mapH;
                if (MULTI && l > 0) {
                    // the previous layer's items that produced frames [t0 - d, t0 - d + 255] and [t0, t0 + 255] of this sequence
                    // (they also performed the previous accumulation into the skip frames this item updates)
                    const LayerDesc& Lp = LD(l - 1);
                    const int pt = Lp.t_begin, ptiles = Lp.tiles_per_seq;
                    const unsigned* flags = p.item_done + Lp.item_base + b * ptiles;
                    for (int rg = 0; rg < 2; ++rg) {
                        int lo = t0 - (rg == 0 ? dil : 0), hi = lo + PM - 1;
                        lo = lo < pt ? pt : lo;
                        hi = hi >= p.L ? p.L - 1 : hi;
                        if (lo > hi) continue;
                        for (int tl = (lo - pt) / PM; tl <= (hi - pt) / PM; ++tl) wait_counter(flags + tl, 2 * EPI_WARPS);
                    }
                    if (Ld.war_layer >= 0) wait_counter(p.layer_done + Ld.war_layer, (unsigned)LD(Ld.war_layer).n_items);
                    asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy writes of other SMs -> this thread's TMA reads
                }
                for (int j = 0; j < C::NT_A; ++j)
                    for (int sl = 0; sl < C::SLABS_A; ++sl) {
                        unsigned bar;
                        unsigned char* dst = acquire(bar);
                        const int tap = sl / (C::SLABS_A / 2);              // tap 0 reads h[t - d], tap 1 reads h[t]
                        tma2_load_4d(dst, mapH, 2 * (tc - (1 - tap) * dil), (sl % (C::SLABS_A / 2)) * C::KC, 0, b, bar);
                        dst = acquire(bar);
                        tma2_load_2d(dst, &mapW, 0, w_row0 + ((j * C::SLABS_A + sl) * 2 + (int)rank) * 8, bar);
                    }
                for (int j = 0; j < (need_skip ? C::NT_B : C::NT_R); ++j)
                    for (int s8 = 0; s8 < C::SLABS_B; ++s8) {
                        unsigned bar;
                        unsigned char* dst = acquire(bar);
                        tma2_load_2d(dst, &mapW, 0, w_row0 + C::WROWS_A + ((j * C::SLABS_B + s8) * 2 + (int)rank) * 8, bar);
                    }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA)
        if (rank == 0) {
            constexpr unsigned idesc = make_idesc_bf16(PM, 256);
            unsigned it = 0, q = 0, n_item = 0;                  // slot counter, n-tile counter (accumulator = q & 1), item counter
            // the products of one k-step (16 channels = 2 chunks): a / bw = byte addresses of the hi planes; the lo planes sit
            // a_lo / SLOT/2 bytes further (pairs only)
            auto mma_step = [&](unsigned d, unsigned a, unsigned a_lo, unsigned bw, unsigned accumulate) {
                const unsigned long long ah = smem_desc(a, LBO, SBO), bh = smem_desc(bw, LBO, SBO);
                umma2_f16(d, ah, bh, idesc, accumulate);
                if constexpr (C::PAIR) {
                    umma2_f16(d, smem_desc(a + a_lo, LBO, SBO), bh, idesc, 1);
                    umma2_f16(d, ah, smem_desc(bw + SLOT / 2, LBO, SBO), idesc, 1);
                }
            };
            auto next_acc = [&]() -> unsigned {                  // claim the next accumulator buffer (waits for its epilogue)
                const unsigned ab = q & 1, u = q >> 1;
                ++q;
                if (u > 0) mbar_wait_cluster(acc_empty + ab, (u - 1) & 1);
                tc_fence_after();
                return ab;
            };
            int l = 0;
            for (int n = cluster_id; n < p.total_items; n += n_clusters, ++n_item) {
                while (n >= LD(l).item_base + LD(l).n_items) ++l;
                const int item = n - LD(l).item_base;
                const int t0 = LD(l).t_begin + (item % LD(l).tiles_per_seq) * PM;
                const bool need_skip = t0 + PM > p.skip_start;
                // ---------------- pass A: n-tiles of [tanh | sigmoid] pre-activations, K = 2 taps x CH
                for (int j = 0; j < C::NT_A; ++j) {
                    const unsigned ab = next_acc(), d = tmem_base + ab * 256;
                    for (int sl = 0; sl < C::SLABS_A; ++sl) {
                        const unsigned sa = it % NSLOT, pa = (it / NSLOT) & 1; ++it;
                        const unsigned sw = it % NSLOT, pw = (it / NSLOT) & 1; ++it;
                        mbar_wait_cluster(full + sa, pa);
                        mbar_wait_cluster(full + sw, pw);
                        tc_fence_after();
                        if (elect_one()) {
                            const unsigned a = s32(ring + sa * SLOT), w = s32(ring + sw * SLOT);
#pragma unroll
                            for (int ks = 0; ks < C::KS / 16; ++ks)
                                mma_step(d, a + ks * 2 * LBO, SLOT / 2, w + ks * 2 * LBO, (sl | ks) != 0);
                            umma2_commit(empty + sa);
                            umma2_commit(empty + sw);
                            if (sl == C::SLABS_A - 1) umma2_commit(acc_full + ab);
                        }
                        __syncwarp();
                    }
                }
                // ---------------- pass B: residual tiles, then skip tiles, from the resident z image
                for (int j = 0; j < (need_skip ? C::NT_B : C::NT_R); ++j) {
                    const unsigned ab = next_acc(), d = tmem_base + ab * 256;
                    for (int s8 = 0; s8 < C::SLABS_B; ++s8) {
                        const unsigned sw = it % NSLOT, pw = (it / NSLOT) & 1; ++it;
                        mbar_wait_cluster(z_ready + s8, n_item & 1);
                        mbar_wait_cluster(full + sw, pw);
                        tc_fence_after();
                        if (elect_one()) {
                            const unsigned a = s32(zbuf) + (unsigned)s8 * C::KC * LBO, w = s32(ring + sw * SLOT);
#pragma unroll
                            for (int ks = 0; ks < C::KS / 16; ++ks)
                                mma_step(d, a + ks * 2 * LBO, C::ZPLANE, w + ks * 2 * LBO, (s8 | ks) != 0);
                            umma2_commit(empty + sw);
                            if (s8 == C::SLABS_B - 1) umma2_commit(acc_full + ab);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===================================================================== epilogue: warps 2..9
        const int qd = warp & 3, grp = (warp - 2) >> 2;               // TMEM lane quadrant (= warp id mod 4), column group
        const int row = qd * 32 + lane;
        const unsigned lane_addr = tmem_base + ((unsigned)(qd * 32) << 16);
        const unsigned acc_empty_addr[2] = {mapa(s32(acc_empty), 0), mapa(s32(acc_empty + 1), 0)};
        const unsigned z_ready_addr = mapa(s32(z_ready), 0);
        const size_t plane_stride = (size_t)(CH / 8) * p.L;            // 16-byte pieces per plane of a pair tensor
        const int Tsk = p.L - p.skip_start;
        unsigned q = 0;
        int l = 0;
        for (int n = cluster_id; n < p.total_items; n += n_clusters) {
            while (n >= LD(l).item_base + LD(l).n_items) ++l;
            const LayerDesc& Ld = LD(l);
            const int item = n - Ld.item_base;
            const int b = item / Ld.tiles_per_seq, t0 = Ld.t_begin + (item % Ld.tiles_per_seq) * PM;
            const bool need_skip = t0 + PM > p.skip_start;
            const int t = t0 + (int)rank * BM + row;                   // this thread's frame
            const bool live = t < p.L;
            const float* bias = Ld.bias;
            float4* fg_save = Ld.fg_save;
            const int skip_init = Ld.skip_init;
            // ---------------- gate: z = tanh(F + bf) * sigmoid(G + bg) -> shared-memory operand image (+ optional saves)
            for (int j = 0; j < C::NT_A; ++j) {
                const unsigned ab = q & 1, u = q >> 1;
                ++q;
                mbar_wait(acc_full + ab, u & 1);
                tc_fence_after();
                const unsigned ta = lane_addr + ab * 256;
#pragma unroll 1
                for (int sb = 0; sb < 128 / C::KS; ++sb) {
                    // slab sb of this n-tile = KS dilation channels; the two column groups take half each, so the slabs become
                    // ready in the order pass B consumes them
#pragma unroll 1
                    for (int c = sb * C::KS + grp * (C::KS / 2); c < sb * C::KS + (grp + 1) * (C::KS / 2); c += 16) {
                        float f[16], g[16];
                        tmem_ld16(ta + c, f);
                        tmem_ld16(ta + 128 + c, g);
                        tmem_ld_wait();
                        const int ch = j * 128 + c;                    // first of the 16 dilation channels
                        const float4* bf4 = reinterpret_cast<const float4*>(bias + ch);
                        const float4* bg4 = reinterpret_cast<const float4*>(bias + CH + ch);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 x = __ldg(bf4 + i), y = __ldg(bg4 + i);
                            f[4 * i] = tanh_fast(f[4 * i] + x.x); f[4 * i + 1] = tanh_fast(f[4 * i + 1] + x.y);
                            f[4 * i + 2] = tanh_fast(f[4 * i + 2] + x.z); f[4 * i + 3] = tanh_fast(f[4 * i + 3] + x.w);
                            g[4 * i] = sigmoid_fast(g[4 * i] + y.x); g[4 * i + 1] = sigmoid_fast(g[4 * i + 1] + y.y);
                            g[4 * i + 2] = sigmoid_fast(g[4 * i + 2] + y.z); g[4 * i + 3] = sigmoid_fast(g[4 * i + 3] + y.w);
                        }
                        if (fg_save != nullptr && live) {
                            float4* fs = fg_save + ((size_t)b * (2 * CH / 4) + ch / 4) * p.L + t;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                fs[(size_t)i * p.L] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                                fs[(size_t)(CH / 4 + i) * p.L] = make_float4(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
                            }
                        }
                        unsigned hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) split2(f[2 * i] * g[2 * i], f[2 * i + 1] * g[2 * i + 1], hi[i], lo[i]);
                        unsigned char* zr = zbuf + (ch / 8) * (BM * 16) + row * 16;
                        *reinterpret_cast<uint4*>(zr) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(zr + BM * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                        if constexpr (C::PAIR) {
                            *reinterpret_cast<uint4*>(zr + C::ZPLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            *reinterpret_cast<uint4*>(zr + C::ZPLANE + BM * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                        }
                    }
                    // this warp's 32 rows of its half of z slab (128j / KS + sb) are in shared memory
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(z_ready_addr + 8u * (unsigned)(j * (128 / C::KS) + sb));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_addr[ab]);
            }
            // ---------------- pass B tiles: residual h_out = acc + br + h_in -> pair; then skip (+)= acc + bs
            for (int j = 0; j < (need_skip ? C::NT_B : C::NT_R); ++j) {
                const unsigned ab = q & 1, u = q >> 1;
                ++q;
                mbar_wait(acc_full + ab, u & 1);
                tc_fence_after();
                const unsigned ta = lane_addr + ab * 256;
                if (j < C::NT_R) {
                    const int n0 = j * 256;                                // first residual channel of this tile
                    const uint4* hin = Ld.h_in + (size_t)b * 2 * plane_stride + t;
                    uint4* hout = Ld.h_out + (size_t)b * 2 * plane_stride + t;
                    uint4 nx[4];
                    auto load_res = [&](int c) {
                        nx[0] = nx[1] = nx[2] = nx[3] = make_uint4(0, 0, 0, 0);
                        if (live) {
                            const uint4* s = hin + (size_t)((n0 + c) / 8) * p.L;
                            nx[0] = __ldg(s); nx[1] = __ldg(s + p.L); nx[2] = __ldg(s + plane_stride); nx[3] = __ldg(s + plane_stride + p.L);
                        }
                    };
                    load_res(grp * 128);
#pragma unroll 1
                    for (int c = grp * 128; c < grp * 128 + 128; c += 16) {
                        float v[16];
                        tmem_ld16(ta + c, v);
                        const uint4 xh0 = nx[0], xh1 = nx[1], xl0 = nx[2], xl1 = nx[3];
                        if (c + 16 < grp * 128 + 128) load_res(c + 16);        // next chunk's loads fly under this chunk's math
                        tmem_ld_wait();
                        const float4* b4 = reinterpret_cast<const float4*>(bias + 2 * CH + n0 + c);
                        const unsigned xh[8] = {xh0.x, xh0.y, xh0.z, xh0.w, xh1.x, xh1.y, xh1.z, xh1.w};
                        const unsigned xl[8] = {xl0.x, xl0.y, xl0.z, xl0.w, xl1.x, xl1.y, xl1.z, xl1.w};
                        unsigned hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 bb = __ldg(b4 + i);
                            const float2 h0 = unpack_bf16x2(xh[2 * i]), l0 = unpack_bf16x2(xl[2 * i]);
                            const float2 h1 = unpack_bf16x2(xh[2 * i + 1]), l1 = unpack_bf16x2(xl[2 * i + 1]);
                            split2(v[4 * i] + bb.x + (h0.x + l0.x), v[4 * i + 1] + bb.y + (h0.y + l0.y), hi[2 * i], lo[2 * i]);
                            split2(v[4 * i + 2] + bb.z + (h1.x + l1.x), v[4 * i + 3] + bb.w + (h1.y + l1.y), hi[2 * i + 1], lo[2 * i + 1]);
                        }
                        if (live) {
                            uint4* o = hout + (size_t)((n0 + c) / 8) * p.L;
                            o[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            o[p.L] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                            o[plane_stride] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            o[plane_stride + p.L] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                        }
                    }
                } else {
                    const int n0 = (j - C::NT_R) * 256;                    // first skip channel of this tile
                    const bool on = live && t >= p.skip_start;
                    float4* sk = p.skip + (size_t)b * (CH / 4) * Tsk + (t - p.skip_start);
                    float4 nx[4];
                    const bool rmw = on && !skip_init;
                    auto load_skip = [&](int c) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            nx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rmw) nx[i] = sk[(size_t)((n0 + c) / 4 + i) * Tsk];
                        }
                    };
                    load_skip(grp * 128);
#pragma unroll 1
                    for (int c = grp * 128; c < grp * 128 + 128; c += 16) {
                        float v[16];
                        tmem_ld16(ta + c, v);
                        const float4 x[4] = {nx[0], nx[1], nx[2], nx[3]};
                        if (c + 16 < grp * 128 + 128) load_skip(c + 16);
                        tmem_ld_wait();
                        const float4* b4 = reinterpret_cast<const float4*>(bias + 3 * CH + n0 + c);
                        if (on) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 bb = __ldg(b4 + i);
                                sk[(size_t)((n0 + c) / 4 + i) * Tsk] = make_float4(v[4 * i] + bb.x + x[i].x, v[4 * i + 1] + bb.y + x[i].y,
                                                                                   v[4 * i + 2] + bb.z + x[i].z, v[4 * i + 3] + bb.w + x[i].w);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_addr[ab]);
            }
            if (MULTI) {
                // this warp's stores of h_out / skip are done: publish (release at gpu scope); the 16th arrival completes the item
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    const unsigned old = atomicAdd(p.item_done + n, 1u);
                    if (old == 2 * EPI_WARPS - 1) { __threadfence(); atomicAdd(p.layer_done + l, 1u); }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                               // the peer may still multicast commits at this CTA's barriers until here
    if (warp == 1) tmem2_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- weight packing
// Packed image of one layer (bf16): pass A blocks [n-tile j][k-slab sl][half r], pass B blocks [j][s][r]; a block is the 16 KB
// slot image [plane][chunk KC][row 128][8].  Pass A: N row n = r*128 + row is filter (r = 0) or gate (r = 1) channel 128j + row;
// K index sl*KS + ck*8 + e = tap*CH + input channel (tap 0 = weight[:, :, 0], the older frame).  Pass B: tiles j < NT_R are
// residual rows, the rest skip rows, output channel 256*(j mod NT_R) + r*128 + row; K = dilation channel.
// All layers in one launch: ptrs[layer] = {wf, wg, bf, bg, wr, ws, br, bs} (biases may be null); blockIdx.y = layer.
template <typename C>
__global__ void pack_block_all_kernel(const float* const* __restrict__ ptrs, __nv_bfloat16* __restrict__ out_all, float* __restrict__ bias_all) {
    constexpr int CH = C::CH;
    const float* const* q = ptrs + (size_t)blockIdx.y * 8;
    const float* wf = q[0]; const float* wg = q[1]; const float* wr = q[4]; const float* ws = q[5];
    __nv_bfloat16* out = out_all + (size_t)blockIdx.y * (C::W_LAYER_BYTES / 2);
    constexpr int n_a = C::NT_A * C::SLABS_A * 2, n_blocks = n_a + C::NT_B * C::SLABS_B * 2;
    constexpr int per_plane = SLOT / 2 / C::PLANES;                        // bf16 elements of one plane of a slot image
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n_blocks * per_plane;
         i += (long long)gridDim.x * blockDim.x) {
        const int blk = (int)(i / per_plane), w = (int)(i % per_plane);   // w: element index inside one plane of the block
        const int ck = w / (BM * 8), row = (w / 8) % BM, e = w % 8;
        float v;
        if (blk < n_a) {
            const int j = blk / (C::SLABS_A * 2), sl = (blk / 2) % C::SLABS_A, r = blk % 2;
            const int kk = sl * C::KS + ck * 8 + e, tap = kk / CH, cin = kk % CH, cout = j * 128 + row;
            v = (r == 0 ? wf : wg)[((size_t)cout * CH + cin) * 2 + tap];
        } else {
            const int bb = blk - n_a, j = bb / (C::SLABS_B * 2), s8 = (bb / 2) % C::SLABS_B, r = bb % 2;
            const int cin = s8 * C::KS + ck * 8 + e, cout = (j % C::NT_R) * 256 + r * 128 + row;
            v = (j < C::NT_R ? wr : ws)[(size_t)cout * CH + cin];
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        __nv_bfloat16* o = out + (size_t)blk * (SLOT / 2) + w;
        o[0] = h;
        if constexpr (C::PAIR) o[per_plane] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 4 * CH; i += blockDim.x) {
            const float* src = i < CH ? q[2] : (i < 2 * CH ? q[3] : (i < 3 * CH ? q[6] : q[7]));
            bias_all[(size_t)blockIdx.y * 4 * CH + i] = src ? src[i % CH] : 0.f;
        }
}

// ---------------------------------------------------------------------------------------------- layout converters
// fp32 frames (B, L, C) -> chunked pair (B, 2, C/8, L, 8) for frames [t_begin, L)
__global__ void pair_from_frames_kernel(const float* __restrict__ x, uint4* __restrict__ out, int B, int L, int C, int t_begin) {
    const int chunks = C / 8;
    const long long total = (long long)B * chunks * (L - t_begin);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = t_begin + (int)(i % (L - t_begin)), ck = (int)((i / (L - t_begin)) % chunks), b = (int)(i / ((long long)(L - t_begin) * chunks));
        const float4* s = reinterpret_cast<const float4*>(x + ((size_t)b * L + t) * C + ck * 8);
        const float4 a = s[0], c = s[1];
        unsigned hi[4], lo[4];
        split2(a.x, a.y, hi[0], lo[0]); split2(a.z, a.w, hi[1], lo[1]); split2(c.x, c.y, hi[2], lo[2]); split2(c.z, c.w, hi[3], lo[3]);
        uint4* o = out + ((size_t)b * 2 * chunks + ck) * L + t;
        o[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        o[(size_t)chunks * L] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}
// chunked pair -> fp32 frames (hi + lo) for frames [t_begin, L)
__global__ void frames_from_pair_kernel(const uint4* __restrict__ in, float* __restrict__ x, int B, int L, int C, int t_begin) {
    const int chunks = C / 8;
    const long long total = (long long)B * chunks * (L - t_begin);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ck = (int)(i % chunks), t = t_begin + (int)((i / chunks) % (L - t_begin)), b = (int)(i / ((long long)(L - t_begin) * chunks));
        const uint4* s = in + ((size_t)b * 2 * chunks + ck) * L + t;
        const uint4 h = s[0], l = s[(size_t)chunks * L];
        const unsigned hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 a = unpack_bf16x2(hw[k]), c = unpack_bf16x2(lw[k]);
            o[2 * k] = a.x + c.x; o[2 * k + 1] = a.y + c.y;
        }
        float4* d = reinterpret_cast<float4*>(x + ((size_t)b * L + t) * C + ck * 8);
        d[0] = make_float4(o[0], o[1], o[2], o[3]);
        d[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}
// chunked fp32 (B, C/4, T, 4) <-> frames (B, T, C) for frames [t_first, t_first + n) of the chunked tensor
__global__ void frames_from_chunks4_kernel(const float4* __restrict__ in, float* __restrict__ x, int B, int T, int C, int t_first, int n) {
    const int chunks = C / 4;
    const long long total = (long long)B * chunks * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ck = (int)(i % chunks), tt = (int)((i / chunks) % n), b = (int)(i / ((long long)n * chunks));
        reinterpret_cast<float4*>(x + ((size_t)b * n + tt) * C)[ck] = in[((size_t)b * chunks + ck) * T + t_first + tt];
    }
}
// start conv on class indices (reference wavenet_model.py:65-68,127 on one-hot input == a gather of one weight column):
// h0 pair <- table[idx[b][t]] where table (classes, ldt) holds start_conv.weight^T (+ bias) as packed by wn_pack_1x1_weights
template <typename IDX>
__global__ void start_pair_kernel(const IDX* __restrict__ idx, const float* __restrict__ table, const float* __restrict__ bias,
                                  uint4* __restrict__ out, int B, int L, int classes, int ldt, int R, int* __restrict__ err) {
    const int chunks = R / 8;
    const long long total = (long long)B * chunks * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % L), ck = (int)((i / L) % chunks), b = (int)(i / ((long long)L * chunks));
        long long cls = (long long)idx[(size_t)b * L + t];
        if (cls < 0 || cls >= classes) { if (err) atomicExch(err, 1); cls = cls < 0 ? 0 : classes - 1; }
        const float4* s = reinterpret_cast<const float4*>(table + (size_t)cls * ldt + ck * 8);
        const float4* bb = reinterpret_cast<const float4*>(bias + ck * 8);
        const float4 a = __ldg(s), c = __ldg(s + 1), ba = __ldg(bb), bc = __ldg(bb + 1);
        unsigned hi[4], lo[4];
        split2(a.x + ba.x, a.y + ba.y, hi[0], lo[0]); split2(a.z + ba.z, a.w + ba.w, hi[1], lo[1]);
        split2(c.x + bc.x, c.y + bc.y, hi[2], lo[2]); split2(c.z + bc.z, c.w + bc.w, hi[3], lo[3]);
        uint4* o = out + ((size_t)b * 2 * chunks + ck) * L + t;
        o[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        o[(size_t)chunks * L] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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
// chunked pair tensor (B, 2, C/8, L, 8) bf16 as 8-byte elements: dims {2(L - origin), C/8, 2, B}; box {2 box_frames, box_chunks,
// box_planes, 1} (box_planes = 1: the hi plane only, for single-pass bf16 operands).  Frames left of `origin` (and right of L)
// are out of bounds -> zeros.
int make_pair_map(CUtensorMap* m, const void* base, int B, int L, int C, int origin, int box_frames, int box_chunks, int box_planes) {
    EncodeTiledFn fn = encode_fn();
    WN_REQUIRE(fn, WN_E_UNSUPP, "cuTensorMapEncodeTiled is not available from this driver");
    WN_REQUIRE(L - origin >= 1, WN_E_BADARG, "empty activation range");
    const cuuint64_t chunks = (cuuint64_t)(C / 8);
    cuuint64_t dims[4] = {(cuuint64_t)2 * (L - origin), chunks, 2, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)L * 16, chunks * L * 16, 2 * chunks * L * 16};
    cuuint32_t box[4] = {(cuuint32_t)(2 * box_frames), (cuuint32_t)box_chunks, (cuuint32_t)box_planes, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void*)((const unsigned char*)base + (size_t)origin * 16), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WN_REQUIRE(r == CUDA_SUCCESS, WN_E_UNSUPP, "cuTensorMapEncodeTiled(pair activations) failed with %d", (int)r);
    return 0;
}
// packed weights: rows of 2 KB (256 8-byte elements); box = 8 rows = one 16 KB slot image
int make_wrows_map(CUtensorMap* m, const void* base, long long rows) {
    EncodeTiledFn fn = encode_fn();
    WN_REQUIRE(fn, WN_E_UNSUPP, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {256, (cuuint64_t)rows};
    cuuint64_t strides[1] = {2048};
    cuuint32_t box[2] = {256, 8};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WN_REQUIRE(r == CUDA_SUCCESS, WN_E_UNSUPP, "cuTensorMapEncodeTiled(packed weights) failed with %d", (int)r);
    return 0;
}

}  // namespace tb
}  // namespace wn

using namespace wn;

extern "C" int wn_tb_supported(int R, int D, int S, int k) { return R == D && D == S && (R == 256 || R == 512) && k == 2; }
// precision: WN_PREC_BF16_PAIRS (fp32-class, channels == 256 only: the z image of 512 channels does not fit as a pair) or
// WN_PREC_BF16 (single-pass bf16 operands, 256 or 512 channels)
extern "C" int wn_tb_precision_supported(int channels, int precision) {
    return (precision == WN_PREC_BF16_PAIRS && channels == 256) || (precision == WN_PREC_BF16 && (channels == 256 || channels == 512));
}
extern "C" size_t wn_tb_weight_bytes_per_layer(int channels, int precision) {
    if (!wn_tb_precision_supported(channels, precision)) return 0;
    if (precision == WN_PREC_BF16_PAIRS) return tb::Cfg<256, true>::W_LAYER_BYTES;
    return channels == 256 ? tb::Cfg<256, false>::W_LAYER_BYTES : tb::Cfg<512, false>::W_LAYER_BYTES;
}

extern "C" int wn_tb_pack_all_weights(const float* const* d_ptrs, int n_layers, int channels, int precision, void* d_w_all,
                                      float* d_bias_all, void* stream) {
    WN_REQUIRE(d_ptrs && d_w_all && d_bias_all && n_layers > 0, WN_E_BADARG, "wn_tb_pack_all_weights: bad arguments");
    WN_REQUIRE(wn_tb_precision_supported(channels, precision), WN_E_UNSUPP, "wn_tb_pack_all_weights: %d channels with precision %d is not supported",
               channels, precision);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(74, n_layers);
    if (precision == WN_PREC_BF16_PAIRS) tb::pack_block_all_kernel<tb::Cfg<256, true>><<<grid, 256, 0, st>>>(d_ptrs, (__nv_bfloat16*)d_w_all, d_bias_all);
    else if (channels == 256) tb::pack_block_all_kernel<tb::Cfg<256, false>><<<grid, 256, 0, st>>>(d_ptrs, (__nv_bfloat16*)d_w_all, d_bias_all);
    else tb::pack_block_all_kernel<tb::Cfg<512, false>><<<grid, 256, 0, st>>>(d_ptrs, (__nv_bfloat16*)d_w_all, d_bias_all);
    WN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wn_pair_from_frames(const float* d_frames, void* d_pair, int B, int L, int C, int t_begin, void* stream) {
    WN_REQUIRE(d_frames && d_pair && B > 0 && L > 0 && C > 0 && C % 8 == 0 && t_begin >= 0 && t_begin < L, WN_E_BADARG,
               "wn_pair_from_frames: bad arguments");
    tb::pair_from_frames_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(d_frames, (uint4*)d_pair, B, L, C, t_begin);
    WN_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int wn_frames_from_pair(const void* d_pair, float* d_frames, int B, int L, int C, int t_begin, void* stream) {
    WN_REQUIRE(d_frames && d_pair && B > 0 && L > 0 && C > 0 && C % 8 == 0 && t_begin >= 0 && t_begin < L, WN_E_BADARG,
               "wn_frames_from_pair: bad arguments");
    tb::frames_from_pair_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>((const uint4*)d_pair, d_frames, B, L, C, t_begin);
    WN_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int wn_frames_from_chunks4(const float* d_chunked, float* d_frames, int B, int T, int C, int t_first, int n, void* stream) {
    WN_REQUIRE(d_chunked && d_frames && B > 0 && T > 0 && C > 0 && C % 4 == 0 && t_first >= 0 && n >= 1 && t_first + n <= T, WN_E_BADARG,
               "wn_frames_from_chunks4: bad arguments");
    tb::frames_from_chunks4_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>((const float4*)d_chunked, d_frames, B, T, C, t_first, n);
    WN_CUDA(cudaGetLastError());
    return 0;
}

static int start_pair(const void* d_idx, bool u8, const float* d_w_t, const float* d_b_p, void* d_h_pair, int B, int classes, int L,
                      int R, int* d_err, void* stream) {
    WN_REQUIRE(d_idx && d_w_t && d_b_p && d_h_pair && B > 0 && L > 0 && classes > 0, WN_E_BADARG, "wn_tb_start_index: bad arguments");
    WN_REQUIRE(R > 0 && R % 8 == 0, WN_E_UNSUPP, "wn_tb_start_index: R must be a multiple of 8");
    const int ldt = wn_n2p(R);
    cudaStream_t st = (cudaStream_t)stream;
    if (u8) tb::start_pair_kernel<uint8_t><<<1184, 256, 0, st>>>((const uint8_t*)d_idx, d_w_t, d_b_p, (uint4*)d_h_pair, B, L, classes, ldt, R, d_err);
    else tb::start_pair_kernel<long long><<<1184, 256, 0, st>>>((const long long*)d_idx, d_w_t, d_b_p, (uint4*)d_h_pair, B, L, classes, ldt, R, d_err);
    WN_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int wn_tb_start_index_u8(const uint8_t* d_idx, const float* d_w_t, const float* d_b_p, void* d_h_pair, int B, int classes,
                                    int L, int R, int* d_err, void* stream) {
    return start_pair(d_idx, true, d_w_t, d_b_p, d_h_pair, B, classes, L, R, d_err, stream);
}
extern "C" int wn_tb_start_index_i64(const int64_t* d_idx, const float* d_w_t, const float* d_b_p, void* d_h_pair, int B, int classes,
                                     int L, int R, int* d_err, void* stream) {
    return start_pair(d_idx, false, d_w_t, d_b_p, d_h_pair, B, classes, L, R, d_err, stream);
}

static int launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, size_t smem, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tb::NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return 0;
}

template <typename C>
static int fill_layer(tb::LayerDesc& d, const void* h_in, void* h_out, const float* bias4, float* fg_save, int layer, int B, int L,
                      int dilation, int in_start, int out_start, int skip_init, int item_base) {
    memset(&d, 0, sizeof(d));
    if (int rc = tb::make_pair_map(&d.mapH, h_in, B, L, C::CH, in_start, tb::BM, C::KC, C::PLANES)) return rc;
    d.t_begin = out_start; d.in_start = in_start; d.dil = dilation; d.skip_init = skip_init;
    d.tiles_per_seq = (L - out_start + tb::PM - 1) / tb::PM;
    d.n_items = B * d.tiles_per_seq;
    d.item_base = item_base;
    d.w_row0 = layer * C::WROWS_LAYER;
    d.bias = bias4; d.h_in = (const uint4*)h_in; d.h_out = (uint4*)h_out; d.fg_save = (float4*)fg_save;
    d.war_layer = -1;
    return 0;
}

template <typename C>
static int launch_block(const wn_tb_block_args* a, cudaStream_t st) {
    int dev = 0, sms = 0;
    WN_CUDA(cudaGetDevice(&dev));
    WN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    tb::LayerDesc d;
    CUtensorMap mW;
    if (int rc = fill_layer<C>(d, a->d_h_in, a->d_h_out, a->d_bias4, a->d_fg_save, a->layer, a->B, a->L, a->dilation, a->in_start,
                               a->out_start, a->skip_init, 0)) return rc;
    if (int rc = tb::make_wrows_map(&mW, a->d_w_all, (long long)a->n_layers * C::WROWS_LAYER)) return rc;
    tb::BlockParams p;
    memset(&p, 0, sizeof(p));
    p.B = a->B; p.L = a->L; p.skip_start = a->skip_start; p.n_layers = 1; p.total_items = d.n_items;
    p.skip = (float4*)a->d_skip;
    WN_CUDA(cudaFuncSetAttribute(tb::block_fused_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    int grid = 2 * p.total_items;
    const int max_grid = (sms / 2) * 2;
    if (grid > max_grid) grid = max_grid;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, grid, C::SMEM_BYTES, st);
    WN_CUDA(cudaLaunchKernelEx(&cfg, tb::block_fused_kernel<C, false>, d, mW, p));
    WN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wn_tb_block_fwd(const wn_tb_block_args* a, void* stream) {
    WN_REQUIRE(a, WN_E_BADARG, "wn_tb_block_fwd: null args");
    WN_REQUIRE(a->d_h_in && a->d_h_out && a->d_skip && a->d_w_all && a->d_bias4, WN_E_BADARG, "wn_tb_block_fwd: null pointer");
    WN_REQUIRE(wn_tb_precision_supported(a->channels, a->precision), WN_E_UNSUPP, "wn_tb_block_fwd: %d channels with precision %d is not supported",
               a->channels, a->precision);
    WN_REQUIRE(a->B > 0 && a->L > 0 && a->dilation >= 1 && a->in_start >= 0 && a->out_start >= a->in_start && a->out_start < a->L &&
                   a->skip_start >= a->out_start && a->skip_start < a->L && a->layer >= 0 && a->layer < a->n_layers,
               WN_E_BADARG, "wn_tb_block_fwd: bad frame ranges or layer index");
    WN_REQUIRE(((uintptr_t)a->d_h_in | (uintptr_t)a->d_h_out | (uintptr_t)a->d_skip | (uintptr_t)a->d_w_all) % 16 == 0, WN_E_BADARG,
               "wn_tb_block_fwd: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->precision == WN_PREC_BF16_PAIRS) return launch_block<tb::Cfg<256, true>>(a, st);
    if (a->channels == 256) return launch_block<tb::Cfg<256, false>>(a, st);
    return launch_block<tb::Cfg<512, false>>(a, st);
}

// ---------------------------------------------------------------------------------------------- the whole stack in one launch
extern "C" size_t wn_tb_stack_desc_bytes(void) { return sizeof(tb::LayerDesc); }
extern "C" long long wn_tb_stack_items(int n_layers, int B, int L, const int* out_start) {
    long long n = 0;
    for (int i = 0; i < n_layers; ++i) n += (long long)B * ((L - out_start[i] + tb::PM - 1) / tb::PM);
    return n;
}

template <typename C>
static int launch_stack(const wn_tb_stack_args* a, cudaStream_t st) {
    int dev = 0, sms = 0;
    WN_CUDA(cudaGetDevice(&dev));
    WN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int nl = a->n_layers;
    std::vector<tb::LayerDesc> desc((size_t)nl);
    int base = 0;
    for (int i = 0; i < nl; ++i) {
        float* fg = a->d_fg_all ? a->d_fg_all + (size_t)i * a->B * (2 * C::CH) * a->L : nullptr;
        if (int rc = fill_layer<C>(desc[i], a->h_ptrs[i], const_cast<void*>(a->h_ptrs[i + 1]), a->d_bias_all + (size_t)i * 4 * C::CH, fg, i, a->B, a->L,
                                   a->dilations[i], a->in_start[i], a->out_start[i], i == 0, base)) return rc;
        WN_REQUIRE(a->in_start[i] >= 0 && a->out_start[i] >= a->in_start[i] && a->out_start[i] < a->L && a->skip_start >= a->out_start[i],
                   WN_E_BADARG, "wn_tb_stack_fwd: bad frame ranges of layer %d", i);
        WN_REQUIRE(a->h_ptrs[i] && a->h_ptrs[i + 1] && a->h_ptrs[i] != a->h_ptrs[i + 1] && (uintptr_t)a->h_ptrs[i + 1] % 16 == 0,
                   WN_E_BADARG, "wn_tb_stack_fwd: layer %d needs distinct 16-byte aligned input and output buffers", i);
        WN_REQUIRE(i == 0 || a->h_ptrs[i + 1] != a->h_ptrs[i - 1], WN_E_BADARG,
                   "wn_tb_stack_fwd: layer %d may not write the buffer layer %d is reading (rotate three buffers)", i, i - 1);
        // write-after-read: the latest earlier layer that READS the buffer this layer writes must be complete first
        for (int j = i - 2; j >= 0; --j)
            if (a->h_ptrs[j] == a->h_ptrs[i + 1]) { desc[i].war_layer = j; break; }
        base += desc[i].n_items;
    }
    CUtensorMap mW;
    if (int rc = tb::make_wrows_map(&mW, a->d_w_all, (long long)nl * C::WROWS_LAYER)) return rc;
    WN_CUDA(cudaMemcpyAsync(a->d_desc, desc.data(), sizeof(tb::LayerDesc) * nl, cudaMemcpyHostToDevice, st));
    WN_CUDA(cudaMemsetAsync(a->d_flags, 0, sizeof(unsigned) * ((size_t)base + nl), st));
    tb::BlockParams p;
    memset(&p, 0, sizeof(p));
    p.B = a->B; p.L = a->L; p.skip_start = a->skip_start; p.n_layers = nl; p.total_items = base;
    p.skip = (float4*)a->d_skip;
    p.layers = (const tb::LayerDesc*)a->d_desc;
    p.item_done = a->d_flags; p.layer_done = a->d_flags + base;
    WN_CUDA(cudaFuncSetAttribute(tb::block_fused_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    // every cluster must be resident (items wait for items of other clusters): one CTA per SM, at most sms/2 clusters
    int grid = 2 * base;
    const int max_grid = (sms / 2) * 2;
    if (grid > max_grid) grid = max_grid;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, grid, C::SMEM_BYTES, st);
    WN_CUDA(cudaLaunchKernelEx(&cfg, tb::block_fused_kernel<C, true>, desc[0], mW, p));
    WN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wn_tb_stack_fwd(const wn_tb_stack_args* a, void* stream) {
    WN_REQUIRE(a, WN_E_BADARG, "wn_tb_stack_fwd: null args");
    WN_REQUIRE(a->h_ptrs && a->d_skip && a->d_w_all && a->d_bias_all && a->d_desc && a->d_flags && a->dilations && a->in_start && a->out_start,
               WN_E_BADARG, "wn_tb_stack_fwd: null pointer");
    WN_REQUIRE(wn_tb_precision_supported(a->channels, a->precision), WN_E_UNSUPP, "wn_tb_stack_fwd: %d channels with precision %d is not supported",
               a->channels, a->precision);
    WN_REQUIRE(a->B > 0 && a->L > 0 && a->n_layers > 0 && a->skip_start >= 0 && a->skip_start < a->L && (uintptr_t)a->d_desc % 128 == 0,
               WN_E_BADARG, "wn_tb_stack_fwd: bad sizes (d_desc must be 128-byte aligned)");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->precision == WN_PREC_BF16_PAIRS) return launch_stack<tb::Cfg<256, true>>(a, st);
    if (a->channels == 256) return launch_stack<tb::Cfg<256, false>>(a, st);
    return launch_stack<tb::Cfg<512, false>>(a, st);
}
