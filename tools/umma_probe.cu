// umma_probe.cu -- hardware probe for the round-2 block kernel design (NOT product code; a measuring tool like lat_probe.cu).
//
// Questions it answers on a B200 in ONE run ("chunked pair layout"):
//   T1  K-major SWIZZLE_NONE shared-memory descriptors on the "channel-chunked" tile image [k-chunk of 8][row][8 bf16]
//       (LBO = stride between the two 16-byte K chunks of a k-step, SBO = stride between 8-row groups), filled by TMA from
//       a (C/8, L, 8) bf16 array described as UINT64 elements with a 2 KB inner box, incl. out-of-bounds zero fill on both
//       sides; tcgen05.mma cta_group::1 (M128) and cta_group::2 (M256: 2SM TMA signalling the leader's barrier,
//       commit multicast, accumulator rows 128r.. in CTA r).
//   T2  MN-major SWIZZLE_NONE descriptors on the same image (contraction over frames: the weight-gradient GEMM).
//   T3  MMA issue rate: cycles per M128/M256 x N256 x K16 MMA from resident operands, cta_group 1 vs 2, alone and with a
//       concurrent TMA stream into other shared-memory slots (is shared-memory bandwidth the limiter?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* b, unsigned parity) {
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try(b, parity)) { if (++spins > (1u << 26)) { printf("probe: barrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); asm volatile("trap;"); } }
}
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
template <int CG>
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned bar_cluster_addr) {
    if constexpr (CG == 2)
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(unsigned* slot, unsigned cols) {
    if constexpr (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(unsigned addr, unsigned cols) {
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_f16(unsigned d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    if constexpr (CG == 2)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {       // arrive on `bar` (same offset) in every CTA of the group
    if constexpr (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(s32(bar)), "h"((unsigned short)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
// SWIZZLE_NONE descriptor: start, LBO, SBO in bytes (all multiples of 16), version 1
__device__ __forceinline__ unsigned long long smem_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= (unsigned long long)1 << 46;
    return d;
}

struct ProbeParams {
    int a_c0, a_c1_base, a_c1_per_rank;       // TMA coordinates of this CTA's A tile: (a_c0 + rank*a_c0_per_rank, ...)
    int a_c0_per_rank;
    int b_c0, b_c1_base, b_c1_per_rank, b_c0_per_rank;
    unsigned a_bytes, b_bytes;                // bytes each CTA's box delivers
    unsigned a_lbo, a_sbo, b_lbo, b_sbo;      // descriptor fields
    unsigned a_kstep, b_kstep;                // start-address advance per k-step (bytes)
    int ksteps;
    unsigned idesc;
    float* out;                               // [CG*128][256]
};

// one cluster (CG CTAs) computes D[CG*128 x 256] with `ksteps` MMAs of K=16
template <int CG>
__global__ void __launch_bounds__(128, 1)
gemm_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ProbeParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                    // up to 64 KB
    unsigned char* sB = base + 65536;            // up to 64 KB
    unsigned long long* full = reinterpret_cast<unsigned long long*>(base + 131072);
    unsigned long long* done = full + 1;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = CG == 2 ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<CG>(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {
        const unsigned bar = CG == 2 ? mapa(s32(full), 0) : s32(full);         // the leader's barrier counts both CTAs' bytes
        if (rank == 0) mbar_expect_tx(full, CG * (p.a_bytes + p.b_bytes));
        tma_load_2d<CG>(sA, &mapA, p.a_c0 + (int)rank * p.a_c0_per_rank, p.a_c1_base + (int)rank * p.a_c1_per_rank, bar);
        tma_load_2d<CG>(sB, &mapB, p.b_c0 + (int)rank * p.b_c0_per_rank, p.b_c1_base + (int)rank * p.b_c1_per_rank, bar);
    }
    if (warp == 1 && rank == 0) {
        mbar_wait(full, 0);
        tc_fence_after();
        if (elect_one()) {
            for (int k = 0; k < p.ksteps; ++k)
                umma_f16<CG>(tmem_base, smem_desc(s32(sA) + k * p.a_kstep, p.a_lbo, p.a_sbo),
                             smem_desc(s32(sB) + k * p.b_kstep, p.b_lbo, p.b_sbo), p.idesc, k != 0);
            umma_commit<CG>(done);
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = (int)rank * 128 + warp * 32 + lane;
    for (int c = 0; c < 256; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((unsigned)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 16; ++i) p.out[(size_t)row * 256 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------- T3: issue-rate benchmark
// Every cluster issues `n_mma` MMAs (M = CG*128, N = 256, K = 16) from resident operands (contents irrelevant), in batches of
// `batch` followed by a commit the issuer waits on two batches later (so the pipe never drains); optionally warp 0 streams
// `tma_bytes_per_batch` of TMA loads into four spare 16 KB slots for every batch.  out[cluster] = cycles for all MMAs.
struct RateParams { int n_mma, batch, tma_per_batch; long long* cycles; };
template <int CG>
__global__ void __launch_bounds__(128, 1)
rate_probe(const __grid_constant__ CUtensorMap mapA, const RateParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                    // 32 KB of "A" (4 k-steps x hi/lo), never reloaded
    unsigned char* sB = base + 32768;            // 32 KB of "B"
    unsigned char* slots = base + 65536;         // 4 x 16 KB TMA landing slots
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + 65536 + 65536);
    unsigned long long* batch_done = bars;       // [4]
    unsigned long long* slot_full = bars + 4;    // [4]
    unsigned long long* stop = bars + 8;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 9);
    const int warp = threadIdx.x >> 5;
    const unsigned rank = CG == 2 ? cluster_rank() : 0;
    for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<unsigned*>(base)[i] = 0x3c003c00u;     // bf16 ~0.0078 pairs
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(batch_done + i, 1); mbar_init(slot_full + i, 1); }
        mbar_init(stop, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) tmem_alloc<CG>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;
    const int n_batches = p.n_mma / p.batch;
    if (warp == 0 && elect_one() && p.tma_per_batch > 0) {
        // TMA stream paced by the clock: tma_per_batch loads of 16 KB per nominal batch time (batch x 128 cycles)
        const long long interval = (long long)p.batch * 128 / p.tma_per_batch;
        const unsigned total = (unsigned)n_batches * (unsigned)p.tma_per_batch;
        const long long t_start = clock64();
        for (unsigned it = 0; it < total; ++it) {
            const int s = it & 3;
            if (it >= 4) mbar_wait(slot_full + s, ((it - 4) >> 2) & 1);
            while (clock64() - t_start < (long long)it * interval) { }
            mbar_expect_tx(slot_full + s, 16384);
            tma_load_2d<1>(slots + s * 16384, &mapA, (int)((it * 256) % 4096), (int)((blockIdx.x * 8) % 64), s32(slot_full + s));
        }
        for (unsigned j = total >= 4 ? total - 4 : 0; j < total; ++j) mbar_wait(slot_full + (j & 3), (j >> 2) & 1);
    }
    if (warp == 1 && rank == 0) {
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((unsigned)((CG * 128) >> 4) << 24);
        long long t0 = 0;
        for (int b = 0; b < n_batches; ++b) {
            if (b >= 2) mbar_wait(batch_done + ((b - 2) & 3), ((b - 2) >> 2) & 1);
            if (b == 2) t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < p.batch; ++i) {
                    const unsigned off = (unsigned)(i & 3) * 4096u;
                    umma_f16<CG>(tmem_base + (unsigned)((i & 1) * 256), smem_desc(s32(sA) + off, 2048, 128),
                                 smem_desc(s32(sB) + off, 2048, 128), idesc, 1);
                }
                umma_commit<CG>(batch_done + (b & 3));
            }
            __syncwarp();
        }
        for (int b = n_batches >= 2 ? n_batches - 2 : 0; b < n_batches; ++b) mbar_wait(batch_done + (b & 3), (b >> 2) & 1);
        const long long t1 = clock64();
        if (threadIdx.x == 32) p.cycles[blockIdx.x / CG] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    return reinterpret_cast<EncodeTiledFn>(p);
}
// chunked array (n_chunks, L, 8) bf16 seen as UINT64 elements: dims {2*(L - origin), n_chunks}; box {2*box_frames, box_chunks}
static CUtensorMap make_chunk_map(const void* base, int n_chunks, int L, int origin, int box_frames, int box_chunks) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)2 * (L - origin), (cuuint64_t)n_chunks};
    cuuint64_t strides[1] = {(cuuint64_t)L * 16};
    cuuint32_t box[2] = {(cuuint32_t)(2 * box_frames), (cuuint32_t)box_chunks};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)((const char*)base + (size_t)origin * 16), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box %d x %d)\n", (int)r, 2 * box_frames, box_chunks); exit(3); }
    return m;
}
static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int CG>
static std::vector<float> run_gemm(const CUtensorMap& mA, const CUtensorMap& mB, ProbeParams p) {
    float* d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * CG * 128 * 256));
    CK(cudaMemset(d_out, 0xff, sizeof(float) * CG * 128 * 256));
    p.out = d_out;
    const size_t smem = 1024 + 131072 + 64;
    CK(cudaFuncSetAttribute(gemm_probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, gemm_probe<CG>, mA, mB, p));
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(CG * 128 * 256, NAN);
    if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(4); }
    CK(cudaMemcpy(out.data(), d_out, sizeof(float) * out.size(), cudaMemcpyDeviceToHost));
    CK(cudaFree(d_out));
    return out;
}
static double max_diff(const std::vector<float>& a, const std::vector<float>& b) {
    double m = 0;
    for (size_t i = 0; i < a.size(); ++i) { double d = std::isfinite(a[i]) ? fabs((double)a[i] - b[i]) : 1e30; if (d > m) m = d; }
    return m;
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
    // ---------------------------------------------------------------- data: activations A (C=64 channels, L=300 frames) and
    // weights W (256 rows x K=64), small integers so fp32 accumulation is exact
    const int C = 64, L = 300, N = 256;
    std::vector<float> a((size_t)L * C), w((size_t)N * C);
    srand(7);
    for (auto& x : a) x = (float)(rand() % 9 - 4);
    for (auto& x : w) x = (float)(rand() % 7 - 3);
    // chunked activation image (C/8, L, 8)
    std::vector<__nv_bfloat16> a_ch((size_t)L * C);
    for (int t = 0; t < L; ++t) for (int c = 0; c < C; ++c) a_ch[((size_t)(c / 8) * L + t) * 8 + c % 8] = __float2bfloat16_rn(a[(size_t)t * C + c]);
    // packed weight image: [half h][chunk][128 rows][8]  == chunked array with "frames" = rows: (2*C/8, 128, 8)
    std::vector<__nv_bfloat16> w_pk((size_t)N * C);
    for (int n = 0; n < N; ++n) for (int c = 0; c < C; ++c)
        w_pk[(((size_t)(n / 128) * (C / 8) + c / 8) * 128 + n % 128) * 8 + c % 8] = __float2bfloat16_rn(w[(size_t)n * C + c]);
    __nv_bfloat16 *d_a, *d_w;
    CK(cudaMalloc(&d_a, a_ch.size() * 2)); CK(cudaMalloc(&d_w, w_pk.size() * 2));
    CK(cudaMemcpy(d_a, a_ch.data(), a_ch.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w_pk.data(), w_pk.size() * 2, cudaMemcpyHostToDevice));
    const int origin = 20;                       // frames left of `origin` must read as zero
    const unsigned idesc_k = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17);             // + M field; both K-major

    // ================================================================ T1: K-major, chunked image
    for (int cg = 1; cg <= 2; ++cg) {
        const int M = cg * 128;
        const int t0 = cg == 2 ? 60 : -30;       // cg2: rows 60..315 (>= L zero fill); cg1: rows -30..97 relative to... see below
        // frames t0 .. t0+M-1 (absolute); TMA coordinate = 2*(t - origin)
        std::vector<float> ref((size_t)M * N, 0.f);
        for (int m = 0; m < M; ++m) {
            const int t = t0 + m;
            if (t < origin || t >= L) continue;
            for (int n = 0; n < N; ++n) { float s = 0; for (int c = 0; c < C; ++c) s += a[(size_t)t * C + c] * w[(size_t)n * C + c]; ref[(size_t)m * N + n] = s; }
        }
        CUtensorMap mA = make_chunk_map(d_a, C / 8, L, origin, 128, C / 8);                  // box: 128 frames x 8 chunks = 16 KB
        CUtensorMap mB = make_chunk_map(d_w, 2 * C / 8, 128, 0, 128, C / 8);                 // a weight half: 128 rows x 8 chunks
        for (int swap = 0; swap < 2; ++swap) {
            ProbeParams p; memset(&p, 0, sizeof(p));
            p.a_c0 = 2 * (t0 - origin); p.a_c0_per_rank = 2 * 128; p.a_c1_base = 0; p.a_c1_per_rank = 0;
            p.b_c0 = 0; p.b_c0_per_rank = 0; p.b_c1_base = 0; p.b_c1_per_rank = C / 8;
            p.a_bytes = 128 * C * 2; p.b_bytes = 128 * C * 2;
            const unsigned kchunk = 128 * 16, rowgrp = 128;
            p.a_lbo = p.b_lbo = swap ? rowgrp : kchunk;
            p.a_sbo = p.b_sbo = swap ? kchunk : rowgrp;
            p.a_kstep = p.b_kstep = 2 * kchunk;
            p.ksteps = C / 16;
            p.idesc = idesc_k | ((unsigned)(M >> 4) << 24);
            std::vector<float> got;
            if (cg == 1) {
                // cta_group::1: one weight half (rows 0..127) -> N = 128
                p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((unsigned)(M >> 4) << 24);
                got = run_gemm<1>(mA, mB, p);
                double m = 0;
                for (int r = 0; r < 128; ++r) for (int n = 0; n < 128; ++n) {
                    const float g = got[(size_t)r * 256 + n];
                    const double d = std::isfinite(g) ? fabs((double)g - ref[(size_t)r * N + n]) : 1e30;
                    if (d > m) m = d;
                }
                printf("T1 K-major  cta_group::1 M128 N128 K64  %s : max|diff| = %.3g  -> %s\n", swap ? "LBO=rowgroup SBO=kchunk" : "LBO=kchunk SBO=rowgroup", m, m == 0 ? "PASS" : "fail");
            } else {
                got = run_gemm<2>(mA, mB, p);
                const double m = max_diff(got, ref);
                printf("T1 K-major  cta_group::2 M256 N256 K64  %s : max|diff| = %.3g  -> %s\n", swap ? "LBO=rowgroup SBO=kchunk" : "LBO=kchunk SBO=rowgroup", m, m == 0 ? "PASS" : "fail");
                if (m != 0 && !swap) {
                    int shown = 0;
                    for (int r = 0; r < M && shown < 6; r += 37) { printf("    row %d: got %g %g %g  want %g %g %g\n", r, got[(size_t)r * N], got[(size_t)r * N + 1], got[(size_t)r * N + 200], ref[(size_t)r * N], ref[(size_t)r * N + 1], ref[(size_t)r * N + 200]); ++shown; }
                }
            }
        }
    }

    // ================================================================ T2: MN-major (contraction over frames), cta_group::2
    // D[n][c] = sum_{t in [t0, t0+64)} G[t][n] * X[t][c]; G (256 ch) and X (256 ch) chunked (32, L2, 8); CTA r holds channels 128r..
    {
        const int C2 = 256, L2 = 200, KT = 64, t0 = 150;      // frames 150..213: the last 14 are out of bounds -> zero fill
        std::vector<float> g((size_t)L2 * C2), x((size_t)L2 * C2);
        for (auto& v : g) v = (float)(rand() % 5 - 2);
        for (auto& v : x) v = (float)(rand() % 7 - 3);
        std::vector<__nv_bfloat16> g_ch(g.size()), x_ch(x.size());
        for (int t = 0; t < L2; ++t) for (int c = 0; c < C2; ++c) {
            g_ch[((size_t)(c / 8) * L2 + t) * 8 + c % 8] = __float2bfloat16_rn(g[(size_t)t * C2 + c]);
            x_ch[((size_t)(c / 8) * L2 + t) * 8 + c % 8] = __float2bfloat16_rn(x[(size_t)t * C2 + c]);
        }
        __nv_bfloat16 *d_g, *d_x;
        CK(cudaMalloc(&d_g, g_ch.size() * 2)); CK(cudaMalloc(&d_x, x_ch.size() * 2));
        CK(cudaMemcpy(d_g, g_ch.data(), g_ch.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_x, x_ch.data(), x_ch.size() * 2, cudaMemcpyHostToDevice));
        std::vector<float> ref((size_t)256 * 256, 0.f);
        for (int n = 0; n < 256; ++n) for (int c = 0; c < 256; ++c) {
            float s = 0;
            for (int t = t0; t < t0 + KT && t < L2; ++t) s += g[(size_t)t * C2 + n] * x[(size_t)t * C2 + c];
            ref[(size_t)n * 256 + c] = s;
        }
        CUtensorMap mG = make_chunk_map(d_g, C2 / 8, L2, 0, KT, 16);          // box: 64 frames x 16 chunks (128 channels) = 16 KB
        CUtensorMap mX = make_chunk_map(d_x, C2 / 8, L2, 0, KT, 16);
        for (int swap = 0; swap < 2; ++swap) {
            ProbeParams p; memset(&p, 0, sizeof(p));
            p.a_c0 = 2 * t0; p.a_c0_per_rank = 0; p.a_c1_base = 0; p.a_c1_per_rank = 16;
            p.b_c0 = 2 * t0; p.b_c0_per_rank = 0; p.b_c1_base = 0; p.b_c1_per_rank = 16;
            p.a_bytes = p.b_bytes = KT * 128 * 2;
            const unsigned mn_group = KT * 16, k_group = 128;                  // next 8 channels: +KT*16 bytes; next 8 frames: +128 bytes
            p.a_lbo = p.b_lbo = swap ? mn_group : k_group;
            p.a_sbo = p.b_sbo = swap ? k_group : mn_group;
            p.a_kstep = p.b_kstep = 2 * k_group;
            p.ksteps = KT / 16;
            p.idesc = idesc_k | (1u << 15) | (1u << 16) | ((256u >> 4) << 24);
            std::vector<float> got = run_gemm<2>(mG, mX, p);
            const double m = max_diff(got, ref);
            printf("T2 MN-major cta_group::2 M256 N256 K64  %s : max|diff| = %.3g  -> %s\n", swap ? "LBO=mn-group SBO=k-group" : "LBO=k-group SBO=mn-group", m, m == 0 ? "PASS" : "fail");
        }
    }

    // ================================================================ T3: MMA issue rate
    {
        const int sms = prop.multiProcessorCount;
        long long* d_cyc;
        CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
        // a big chunked array to stream from (64 chunks x 4096 frames x 16 B = 4 MB, L2 resident)
        void* d_big;
        CK(cudaMalloc(&d_big, (size_t)64 * 4096 * 16));
        CK(cudaMemset(d_big, 0, (size_t)64 * 4096 * 16));
        CUtensorMap mS = make_chunk_map(d_big, 64, 4096, 0, 128, 8);          // 16 KB boxes
        const size_t smem = 1024 + 131072 + 128;
        CK(cudaFuncSetAttribute(rate_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(rate_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int n_mma = 24 * 400, batch = 24;
        for (int cg = 1; cg <= 2; ++cg)
            for (int tma = 0; tma <= 6; tma += 2) {
                RateParams rp; rp.n_mma = n_mma; rp.batch = batch; rp.tma_per_batch = tma; rp.cycles = d_cyc;
                cudaLaunchConfig_t cfg = {};
                const int grid = (sms / cg) * cg;
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (cg == 1) CK(cudaLaunchKernelEx(&cfg, rate_probe<1>, mS, rp)); else CK(cudaLaunchKernelEx(&cfg, rate_probe<2>, mS, rp));
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                std::vector<long long> cyc(sms);
                CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * (grid / cg), cudaMemcpyDeviceToHost));
                double avg = 0; for (int i = 0; i < grid / cg; ++i) avg += (double)cyc[i]; avg /= (grid / cg);
                const double timed_mma = (double)(n_mma - 2 * batch);
                const double flop = 2.0 * (cg * 128) * 256 * 16 * (double)n_mma * (grid / cg);
                printf("T3 rate cta_group::%d  M%d N256 K16, %d MMAs/cluster, TMA stream %d x 16 KB per %d MMAs (%.0f B per MMA): "
                       "%.1f cycles/MMA, kernel %.3f ms -> %.0f TFLOP/s dense bf16\n", cg, cg * 128, n_mma, tma, batch,
                       tma * 16384.0 / batch, avg / timed_mma, best, flop / (best * 1e-3) / 1e12);
            }
    }
    printf("probe done\n");
    return 0;
}
