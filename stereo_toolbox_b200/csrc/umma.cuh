// Thin inline-PTX layer for the Blackwell (sm_100a) data path used by conv3d_umma.cu:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (TMEM alloc, UMMA issue/commit, TMEM loads) and the
// shared-memory / instruction descriptors.  Bit layouts follow the PTX ISA tcgen05 descriptors
// (cross-checked against cute/arch/mma_sm100_desc.hpp in the vendored CUTLASS headers).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (uint32_t i = 1;; ++i) {
        if (mbar_try_wait(bar, parity)) return;
        if ((i & 255u) == 0 && clock64() - t0 > 6000000000LL) break;   // ~3 s at 2 GHz
    }
    asm volatile("trap;");
}

// Waits used by warps that do not sit on the MMA critical path.  (A sleeping back-off was tried for the epilogue: it does
// not change the tcgen05.mma rate -- the issue rate is bounded elsewhere, see profiles/umma_issue_r01.md -- but
// __nanosleep's granularity added 0.3-2K clk to every accumulator hand-over, so the epilogue spins; only the TMA
// producer, which runs many planes ahead, sleeps between polls.)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (uint32_t i = 1;; ++i) {
        __nanosleep(ns);
        if (mbar_try_wait(bar, parity)) return;
        if ((i & 255u) == 0 && clock64() - t0 > 6000000000LL) break;
    }
    asm volatile("trap;");
}
// One lane polls, the rest of the warp parks on __syncwarp.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ---- thread-block clusters: the CTAs that hold the output-channel slices of one tile read the SAME input planes, so a plane
// box is fetched from L2 once and written into the shared memory of every CTA of the cluster (and its byte count signalled
// on the mbarrier at the same offset in each of them).
__device__ __forceinline__ void tma_load_5d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                               int c4, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, "
        "%6, %7}], [%2], %8;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// every thread of every CTA of the cluster
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One lane of a converged warp: the compiler treats an elect.sync predicate as single-thread, so uniform-
// datapath instructions (UTCHMMA, UTMALDG) issue without the ELECT/BRA.U.ANY serialisation loop that a plain
// `if (lane == 0)` produces.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* holder_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16/fp16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, descriptors passed as (low word, shared high word) pairs so the 64-bit values are assembled in PTX.
__device__ __forceinline__ void mma_f16_ss2(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Predicated form (always accumulates): the MMA is skipped when `enable` is 0 -- a uniform predicate on the
// instruction instead of a branch around it.
__device__ __forceinline__ void mma_f16_ss2p(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t desc_hi, uint32_t idesc,
                                             bool enable) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(desc_hi), "r"(idesc), "r"((uint32_t)enable)
        : "memory");
}
// Fully predicated form: issued only when `enable` is set; accumulates when `accumulate` is set.
__device__ __forceinline__ void mma_f16_ss3(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t accumulate, bool enable) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"((uint32_t)enable)
        : "memory");
}
// A value held by ONE active thread (mask = that thread's lane bit), returned through REDUX: ptxas keeps the result in
// a uniform register, which is what the UTCHMMA operands and uniform branches want.
__device__ __forceinline__ uint32_t uni(uint32_t self_mask, uint32_t v) { return __reduce_or_sync(self_mask, v); }
// mbarrier arrives once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// the same arrival on the mbarrier at this offset in EVERY CTA of cta_mask (ring slots are refilled by multicast, so a slot
// is free only when the consumers of all CTAs of the cluster have released it)
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of warp w reads lane (32*(w%4)+t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// one fp32 column of 32 lanes (the single-channel classifier epilogue)
__device__ __forceinline__ void tmem_ld_32x32_x1(uint32_t taddr, uint32_t& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ descriptors
enum Swizzle : uint32_t { SW_NONE = 0, SW_128B = 2, SW_64B = 4, SW_32B = 6 };

// K-major operand tile in shared memory. start/lbo/sbo in BYTES (16-byte granular).
__host__ __device__ __forceinline__ uint64_t smem_desc(uint32_t start_addr, uint32_t lbo, uint32_t sbo,
                                                       uint32_t layout, uint32_t base_offset = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((start_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

// kind::f16 instruction descriptor: D=f32, A/B = bf16 (fmt 1) or f16 (fmt 0), both K-major.
__host__ __device__ __forceinline__ uint32_t instr_desc_f16(uint32_t M, uint32_t N, uint32_t ab_fmt = 1) {
    uint32_t d = 0;
    d |= 1u << 4;               // c_format = F32
    d |= (ab_fmt & 7) << 7;     // a_format
    d |= (ab_fmt & 7) << 10;    // b_format
    d |= (N >> 3) << 17;        // n_dim
    d |= (M >> 4) << 24;        // m_dim
    return d;
}

}  // namespace umma

// ---------------------------------------------------------------------------------------------- host
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
namespace umma_host {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// dims/strides innermost first; strides in bytes for dims 1..rank-1.
inline bool make_tmap(CUtensorMap* m, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw,
                      const uint32_t* elem_strides = nullptr, int promo_bytes = 256) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i > 0) gs[i - 1] = strides_bytes[i - 1];
    }
    // L2 promotion widens every TMA request to 64 / 128 / 256 B sectors-groups: a box whose inner extent covers only PART of a
    // row (one K-chunk of a K-split pass) must not promote beyond its own bytes, or every pass fetches the whole tensor
    const CUtensorMapL2promotion promo = promo_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                         : promo_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                         : promo_bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r = fn(m, dt, (cuuint32_t)rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
}  // namespace umma_host
