// tc_ptx.cuh -- inline-PTX wrappers for the cta_group::2 tensor-core kernels (tc_block.cu, tc_bwd2.cu): mbarrier, TMA with
// the barrier in the pair's leader CTA, tcgen05 alloc / mma / commit / ld, SWIZZLE_NONE shared-memory descriptors.
// Every form used here was verified on a B200 by tools/umma_probe.cu (profiles/umma_probe_r2_a.txt).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace wn {
namespace px {

constexpr unsigned SPIN_LIMIT = 1u << 28;     // a barrier that never completes traps instead of hanging the GPU

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
// arrive (release at cluster scope) on a barrier given by its shared::cluster address (any CTA of the cluster)
__device__ __forceinline__ void mbar_arrive_cluster(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* b, unsigned parity) {
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try(b, parity)) if (++spins > SPIN_LIMIT) asm volatile("trap;");
}
// the same with acquire at cluster scope: the arrivals came from the peer CTA (mbar_arrive_cluster)
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* b, unsigned parity) {
    unsigned done, spins = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
        if (!done && ++spins > SPIN_LIMIT) asm volatile("trap;");
    } while (!done);
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}

// ---- TMA loads executed by both CTAs of a pair; `bar` is the shared::cluster address of the LEADER's barrier
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// ---- tensor memory, cta_group::2 (the same warp of BOTH CTAs executes alloc / dealloc)
__device__ __forceinline__ void tmem2_alloc(unsigned* slot_in_smem, unsigned cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(unsigned addr, unsigned cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 operands, fp32 accumulate), M = 256 over the CTA pair; leader thread only
__device__ __forceinline__ void umma2_f16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on `bar` (same shared-memory offset) in BOTH CTAs once all MMAs issued so far by this thread have retired
__device__ __forceinline__ void umma2_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(s32(bar)), "h"((unsigned short)3) : "memory");
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

// SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).  K-major operand on the chunked
// tile image [k-chunk of 8 elements][row][16 bytes]: lbo = bytes between the two 16-byte K chunks of a k-step (= rows * 16),
// sbo = bytes between 8-row groups (= 128).  MN-major operand on the same image: lbo = bytes between 8-frame K groups (128),
// sbo = bytes between 8-channel MN groups (= frames * 16).
__device__ __forceinline__ unsigned long long smem_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= (unsigned long long)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, M x N, majors (0 = K, 1 = MN)
__host__ __device__ constexpr unsigned make_idesc_bf16(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) | ((unsigned)(N >> 3) << 17) |
           ((unsigned)(M >> 4) << 24);
}

// ---- activations for the gate: ex2 / rcp approximations (relative error ~2e-7; the parity bar is 1e-4 on the logits)
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcpf(1.f + ex2f(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcpf(1.f + ex2f(2.8853900817779268f * x)); }

// fp32 -> bf16 (hi, lo) pair; packs two values per 32-bit word, first value in the low half
__device__ __forceinline__ void split2(float a, float b, unsigned& hi, unsigned& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const unsigned*>(&h);
    lo = *reinterpret_cast<const unsigned*>(&l);
}
__device__ __forceinline__ float2 unpack_bf16x2(unsigned w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

}  // namespace px
}  // namespace wn
