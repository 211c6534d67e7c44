// Tensor-core 3-D convolution family for sm_100a: implicit GEMM on tcgen05 with TMEM accumulators,
// WITHOUT materialising im2col -- every filter tap is a row-shifted *view* of one TMA-staged,
// zero-padded activation tile in shared memory.
//
//   activations : channels-last NDHWC, 16-bit (bf16 or fp16)  x[b][d][h][w][c]
//   weights     : tap-major, K-major 16-bit  wt[tile][kchunk][co][KC]  (eval BatchNorm scale folded in)
//   accumulate  : fp32 in TMEM; epilogue = (+ fp32 partial) + shift[co] (+ residual) -> act -> 16-bit (or fp32)
//
// Work item (one CTA): an (h-tile, w-tile) column of the volume marching along depth.  Each input
// depth-plane tile  [(TH+span_h) x 32 voxels x KC]  is loaded ONCE by TMA (5-D box, out-of-bounds
// = zero = conv padding) into a ring of shared-memory slots.  Rows of the UMMA A operand are the
// flattened (h, w) positions of the *padded* tile (row pitch 32 voxels), so tap (dz, dh, dw) is the same
// tile with its descriptor start address advanced by (dh*32 + dw) rows (verified on hardware by
// probe/umma_probe.cu: the 128B/64B/32B swizzles are functions of the absolute smem address, base_offset
// stays 0): the 128 rows of an M-tile are 4 padded rows; columns >= TW of each padded row are junk outputs
// that the epilogue skips.  Flavours, all expressed as tap tables built on the host:
//   Conv3d k3 s1 p1 (27 taps), k1 (1 tap)                       PSMNet/submodule.py:16-19
//   Conv3d k3 s2 p1: the input plane is staged as 4 (h,w)-parity sub-tiles (TMA element strides 2), so a
//     stride-2 tap is again a unit-stride row shift inside one sub-tile; depth advances 2 planes per step
//   ConvTranspose3d k3 s2 p1 op1 / k4 s2 p1: 8 output-parity classes share the staged tile; each class is
//     an accumulator round of 1..8 taps stored with stride 2    PSMNet/stackhourglass.py:25-29
// Input channels beyond one K-chunk (64, or 32 for stride 2) are handled by the host as K-split passes that
// chain through an fp32 partial buffer; output channels whose weight tiles do not fit in smem as N-split passes.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane, also owns
// TMEM alloc/dealloc), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Pipelines:
// plane ring (full/empty mbarriers, released by tcgen05.commit) and a 2-deep TMEM accumulator ring.
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

using namespace umma;

namespace {

constexpr int TWP = 32;            // padded tile row pitch in voxels (one warp = one padded row in the epilogue)
constexpr int MAX_UTAPS = 64;
constexpr int MAX_UCLASS = 8;
constexpr int MAX_RING = 12;
constexpr int UMMA_THREADS = 352;      // warp 0 TMA, warps 1-2 MMA issuers, warps 3-10 epilogue (two groups of 4)
constexpr size_t SMEM_CAP = 227 * 1024;

struct UTap { int8_t dz; uint8_t sub; int16_t rowoff; uint16_t widx; uint8_t nblk; uint8_t cls0; };
struct UClass { uint16_t tap_begin, tap_end; int8_t od0, oh0, ow0, pad; };

struct UArgs {
    const void* residual;            // NDHWC 16-bit, Cout_total channels (nullable)
    const float* partial;            // fp32 [.., Cout_total] partial sums of earlier K-split passes (nullable)
    void* out;                       // 16-bit NDHWC (Cout_total) or fp32
    const float* shift;              // [Cout_total] (nullable)
    int B, Do, Ho, Wo;
    int Cn, Cn_valid, cout_off, Cout_total, w_rows, w_tile_stride, w_kc_off, nwtiles;
    int TH, TW, nM;
    int sd_in, dzmin, dzmax, R, in_stride, nsub;
    int out_stride, nclass;
    int nsteps, dchunk, nchunks, tiles_h, tiles_w;
    int in_h_off, in_w_off, cin_off;
    int act, out_fp32, f16;
    int ROWB, layout, bo_mode, merge, ntaps_total;
    uint32_t tab_bytes;              // issue table bytes (multiple of 1024)
    int merge_step;                  // lane distance between kw-merged column blocks (= dilation along w)
    int cblocks;                     // accumulator column blocks (of Cn) per M-tile: 1, 3 (kw-merge) or 8 (merged transposed conv)
    int nclass_h, nclass_w;          // class output extents per step in h/w (positions)
    uint32_t plane_bytes, chunk_bytes, wtile_bytes, w_bytes_total, tmem_cols;
    UClass cls[MAX_UCLASS];
    UTap taps[MAX_UTAPS];
};

// Optional in-kernel timeline (debug aid, see tools/umma_trace.py): when armed through stb_conv3d_umma_set_trace,
// CTA 0 records globaltimer-free clock64() stamps per accumulator round: [0] issuer-0 start of issue, [1] issuer-0
// after commit, [2] epilogue group 0 woke (accumulator ready), [3] epilogue group 0 released the buffer
// (8 int64 slots per round: step top, planes resident, issue start, commit, epilogue wake, epilogue end).
__device__ long long* g_umma_trace = nullptr;
__device__ int g_umma_trace_rounds = 0;

__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
    if (f16) {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack16(uint32_t w, int f16) {
    if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ float load16(const uint16_t* p, int f16) {
    return f16 ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
}
__device__ __forceinline__ void store16(uint16_t* p, float v, int f16) {
    if (f16) *reinterpret_cast<__half*>(p) = __float2half_rn(v);
    else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(v);
}

// ACT / F16 are compile-time so the epilogue stays a few hundred instructions: with a runtime activation switch
// unrolled over 32 channels the kernel was 83 KB of SASS and the epilogue warps stalled on instruction fetch
// (ncu: stall_no_inst on every epilogue line, profiles/ncu_umma_r01_*.txt).
template <int ACT, bool F16>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv3d_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ UArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    // [0,2048): barriers + tmem holder + issue tables ; then weights ; then plane ring
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);
    uint64_t* plane_full = bar_w + 1;
    uint64_t* plane_empty = plane_full + MAX_RING;
    uint64_t* tmem_full = plane_empty + MAX_RING;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint32_t* tapA = tmem_holder + 2;            // [MAX_UTAPS] A byte offset of the tap inside a plane slot
    uint32_t* tapB = tapA + MAX_UTAPS;           // [MAX_UTAPS] low word of the tap's B (weight tile) descriptor
    uint32_t* tapZ = tapB + MAX_UTAPS;           // [MAX_UTAPS] plane index of the tap relative to dzmin
    uint32_t* tapI = tapZ + MAX_UTAPS;           // [MAX_UTAPS] instruction descriptor (N = nblk*Cn differs per tap)
    uint32_t* tapD = tapI + MAX_UTAPS;           // [MAX_UTAPS] accumulator column offset (cls0*Cn)
    uint32_t* slotTab = tapD + MAX_UTAPS;        // [MAX_RING] encoded (addr >> 4) of every ring slot
    uint4* issueTab = reinterpret_cast<uint4*>(smem + 2048);      // [R][ntaps] {A desc lo, B desc lo, idesc, D column offset}
    uint8_t* sW = smem + 2048 + a.tab_bytes;
    uint8_t* sP = sW + ((a.w_bytes_total + 1023) & ~1023u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- work item decode
    int t = blockIdx.x;
    const int tw_i = t % a.tiles_w; t /= a.tiles_w;
    const int th_i = t % a.tiles_h; t /= a.tiles_h;
    const int ch_i = t % a.nchunks;
    const int b = t / a.nchunks;
    const int s_lo = ch_i * a.dchunk, s_hi = min(a.nsteps, s_lo + a.dchunk);
    const int jh0 = th_i * a.TH, jw0 = tw_i * a.TW;          // class-position origin of this tile
    const int p_first = s_lo * a.sd_in + a.dzmin;
    const int p_last = (s_hi - 1) * a.sd_in + a.dzmax;
    const int nplanes = p_last - p_first + 1;
    const int nouts = (s_hi - s_lo) * a.nclass;               // accumulator rounds

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        const int n_iss = nouts >= 2 ? 2 : 1;                                    // MMA issuers take alternate rounds
        const int n_grp = a.nM * (a.cblocks == 8 ? 8 : 1) >= 2 ? 2 : 1;           // active epilogue groups
        for (int i = 0; i < a.R; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], n_iss); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4 * n_grp); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, a.tmem_cols);
    // issue tables (see the MMA issuer): per tap the A offset inside a ring slot, the B descriptor low word and the
    // plane index relative to dzmin; per ring slot its encoded base address
    for (int tp = threadIdx.x; tp < a.ntaps_total; tp += UMMA_THREADS) {
        tapA[tp] = ((uint32_t)a.taps[tp].sub * a.chunk_bytes + (uint32_t)a.taps[tp].rowoff * a.ROWB) >> 4;
        tapB[tp] = (((smem_u32(sW) + (uint32_t)a.taps[tp].widx * a.wtile_bytes) >> 4) & 0x3FFFu) | (1u << 16);
        tapZ[tp] = (uint32_t)(a.taps[tp].dz - a.dzmin);
        tapI[tp] = instr_desc_f16(128, (uint32_t)a.taps[tp].nblk * a.Cn, F16 ? 0 : 1);
        tapD[tp] = (uint32_t)a.taps[tp].cls0 * a.Cn;
    }
    for (int i = threadIdx.x; i < a.R; i += UMMA_THREADS) slotTab[i] = (smem_u32(sP) + (uint32_t)i * a.plane_bytes) >> 4;
    // one 16-byte record per (ring phase, tap): everything the issuer needs for the tap's first MMA of M-tile 0
    for (int i = threadIdx.x; i < a.R * a.ntaps_total; i += UMMA_THREADS) {
        const int ph = i / a.ntaps_total, tp = i - ph * a.ntaps_total;
        int sl = ph + (int)(a.taps[tp].dz - a.dzmin);
        sl -= (sl >= a.R) ? a.R : 0;
        const uint32_t a16 = ((smem_u32(sP) + (uint32_t)sl * a.plane_bytes) >> 4) +
                             (((uint32_t)a.taps[tp].sub * a.chunk_bytes + (uint32_t)a.taps[tp].rowoff * a.ROWB) >> 4);
        uint4 e;
        e.x = a16 & 0x3FFFu;                                                    // (masked + flagged at issue, after + M-tile offset)
        e.y = (((smem_u32(sW) + (uint32_t)a.taps[tp].widx * a.wtile_bytes) >> 4) & 0x3FFFu) | (1u << 16);
        e.z = instr_desc_f16(128, (uint32_t)a.taps[tp].nblk * a.Cn, F16 ? 0 : 1);
        e.w = (uint32_t)a.taps[tp].cls0 * a.Cn;
        issueTab[i] = e;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            prefetch_tmap(&tm_x);
            prefetch_tmap(&tm_w);
            mbar_arrive_expect_tx(bar_w, a.w_bytes_total);
            for (int i = 0; i < a.nwtiles; ++i)
                tma_load_2d(sW + (size_t)i * a.wtile_bytes, &tm_w, bar_w, 0,
                            (i * a.w_tile_stride + a.w_kc_off) * a.w_rows + a.cout_off);
            const int ih0 = (jh0 + a.in_h_off) * a.in_stride, iw0 = (jw0 + a.in_w_off) * a.in_stride;
            for (int n = 0; n < nplanes; ++n) {
                const int slot = n % a.R;
                mbar_wait(&plane_empty[slot], ((n / a.R) & 1) ^ 1);
                mbar_arrive_expect_tx(&plane_full[slot], (uint32_t)a.nsub * a.chunk_bytes);
                uint8_t* dst = sP + (size_t)slot * a.plane_bytes;
                for (int sb = 0; sb < a.nsub; ++sb)       // sub-tile sb = (h parity, w parity) for stride 2
                    tma_load_5d(dst + (size_t)sb * a.chunk_bytes, &tm_x, &plane_full[slot], a.cin_off,
                                iw0 + (sb & 1), ih0 + (sb >> 1), p_first + n, b);
            }
        }
    } else if (warp <= 2) {
        // ================================ MMA issuers (2) ================================
        // The two issuing warps take ALTERNATE accumulator rounds: while one is blocked on a full MMA queue, the
        // other performs the waits / bookkeeping of the next round, so the tensor pipe sees back-to-back work
        // (in-kernel timeline, profiles/umma_trace_r01.txt: with a single issuing sequence ~1.6K of every 4.2K
        // clocks were barrier waits with the pipe idle).  The issue loop itself carries no division, no descriptor
        // construction and no constant-bank traffic (smem tables above).
        const int issuer = warp - 1;
        const int n_iss = nouts >= 2 ? 2 : 1;
        if (issuer < n_iss && elect_one()) {
            const uint64_t desc_hi = (uint64_t)((((8u * a.ROWB) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)a.layout << 29)) << 32;
            const int ksteps = a.ROWB / 32;
            const uint32_t mtile16 = (128u * a.ROWB) >> 4;
            const int nM = a.nM, R = a.R, sd = a.sd_in, nclass = a.nclass;
            const uint32_t ncol = (uint32_t)(a.Cn * a.cblocks);
            mbar_wait(bar_w, 0);
            int waited = 0;                       // planes [0, waited) are known to be resident (this issuer's view)
            int released = 0;                     // planes [0, released) have been handed back by this issuer
            int wslot = 0, wphase = 0;            // ring slot / phase of plane `waited`
            int rslot = 0;                        // ring slot of plane `released`
            for (int round = issuer; round < nouts; round += n_iss) {
                const int si = round / nclass, c = round - si * nclass;      // step index within the chunk, class
                const int s = s_lo + si;
                const int need = (s * sd + a.dzmax) - p_first + 1;
                const bool trace = g_umma_trace && blockIdx.x == 0 && round < g_umma_trace_rounds;
                if (trace) g_umma_trace[round * 8 + 0] = clock64();
                // Ring bookkeeping.  Both issuers observe (wait for) EVERY plane in order and hand back every plane
                // below their current window, also planes only the other issuer read: a plane slot is recycled when both
                // have released it.  An issuer releases a plane only after it has seen that plane land, so its arrival can
                // never fall into the slot's previous phase (that race corrupted the ring once -- flaky deadlock/trap).
                const int dead_now = (s * sd + a.dzmin) - p_first;
                while (waited < need) {
                    while (released < dead_now && released < waited) {
                        mma_commit(&plane_empty[rslot]);
                        ++released;
                        if (++rslot == R) rslot = 0;
                    }
                    mbar_wait(&plane_full[wslot], wphase);
                    ++waited;
                    if (++wslot == R) { wslot = 0; wphase ^= 1; }
                }
                while (released < dead_now) {
                    mma_commit(&plane_empty[rslot]);
                    ++released;
                    if (++rslot == R) rslot = 0;
                }
                if (trace) g_umma_trace[round * 8 + 1] = clock64();
                const int buf = round & 1;
                mbar_wait(&tmem_empty[buf], ((round >> 1) & 1) ^ 1);
                tc_fence_after();
                if (trace) g_umma_trace[round * 8 + 2] = clock64();
                const int phase = ((s * sd + a.dzmin) - p_first) % R;         // one division per round
                const uint4* tab = issueTab + phase * a.ntaps_total;
                const int t0 = a.cls[c].tap_begin, t1 = a.cls[c].tap_end;
                for (int m = 0; m < nM; ++m) {
                    const uint32_t dcol = tmem_base + (uint32_t)(buf * nM + m) * ncol;
                    const uint32_t moff = (uint32_t)m * mtile16;
                    uint32_t acc = 0;
#pragma unroll 3
                    for (int tp = t0; tp < t1; ++tp) {
                        const uint4 e = tab[tp];                                   // one LDS.128 per tap
                        const uint32_t alo = ((e.x + moff) & 0x3FFFu) | (1u << 16);
                        const uint32_t blo = e.y;
                        const uint32_t idesc = e.z;
                        const uint32_t dcol_t = dcol + e.w;
                        mma_f16_ss(dcol_t, desc_hi | (uint64_t)alo, desc_hi | (uint64_t)blo, idesc, acc);
                        if (ksteps >= 2)
                            mma_f16_ss(dcol_t, desc_hi | (uint64_t)(alo + 2u), desc_hi | (uint64_t)(blo + 2u), idesc, 1u);
                        if (ksteps == 4) {
                            mma_f16_ss(dcol_t, desc_hi | (uint64_t)(alo + 4u), desc_hi | (uint64_t)(blo + 4u), idesc, 1u);
                            mma_f16_ss(dcol_t, desc_hi | (uint64_t)(alo + 6u), desc_hi | (uint64_t)(blo + 6u), idesc, 1u);
                        }
                        acc = 1;
                    }
                }
                mma_commit(&tmem_full[buf]);
                if (trace) g_umma_trace[round * 8 + 3] = clock64();
                // hand back every plane this issuer will not read again: its next round is `round + n_iss`
                const int nxt = round + n_iss;
                int dead_upto = nxt < nouts ? ((s_lo + nxt / nclass) * sd + a.dzmin) - p_first : nplanes;
                if (dead_upto > waited) dead_upto = waited;          // only planes this issuer has seen land
                while (released < dead_upto) {
                    mma_commit(&plane_empty[rslot]);
                    ++released;
                    if (++rslot == R) rslot = 0;
                }
            }
        }
    } else {
        // ================================ epilogue warps ================================
        const int q4 = warp & 3;                    // TMEM lane quarter this warp may touch
        const int egroup = (warp - 3) >> 2;         // epilogue group 0/1 drains M-tiles m = egroup, egroup+2, ...
        const int nblk_e = a.cblocks == 8 ? 8 : 1;     // merged transposed conv: all 8 parity classes in one round
        const int items = a.nM * nblk_e;               // (M-tile, class block) work items per round, dealt to 2 groups
        const int my_rounds = egroup < (items >= 2 ? 2 : 1) ? nouts : 0;
        const size_t ostride_w = (size_t)a.Cout_total;
        constexpr int f16 = F16 ? 1 : 0;
        const bool full32 = (a.Cn_valid & 31) == 0;  // every 32-column block is complete: vector path
        float sh0[32];                              // folded-BN shift of the first 32 channels stays in registers
#pragma unroll
        for (int i = 0; i < 32; ++i) sh0[i] = (a.shift && i < a.Cn_valid) ? __ldg(a.shift + a.cout_off + i) : 0.f;
        const int merge = a.merge, Cn = a.Cn, nM = a.nM;
        for (int round = 0; round < my_rounds; ++round) {
            const int buf = round & 1;
            const int s = s_lo + round / a.nclass;
            const UClass cl = a.cls[round % a.nclass];
            mbar_wait(&tmem_full[buf], (round >> 1) & 1);
            tc_fence_after();
            const bool trace = g_umma_trace && blockIdx.x == 0 && warp == 3 && lane == 0 && round < g_umma_trace_rounds;
            if (trace) g_umma_trace[round * 8 + 4] = clock64();
            for (int item = egroup; item < items; item += 2) {
              {
                const int m = item / nblk_e, blk = item - m * nblk_e;
                const int cd = nblk_e == 8 ? (blk >> 2) : cl.od0, chh = nblk_e == 8 ? ((blk >> 1) & 1) : cl.oh0,
                          cww = nblk_e == 8 ? (blk & 1) : cl.ow0;
                const int od = s * a.out_stride + cd;
                const int q = 128 * m + q4 * 32 + lane;
                const int jh_l = q / TWP, jw_l = q % TWP;
                const int jh = jh0 + jh_l, jw = jw0 + jw_l;
                const bool valid = (jh_l < a.TH) && (jw_l < a.TW) && (jh < a.nclass_h) && (jw < a.nclass_w) && (od < a.Do);
                const int oh = jh * a.out_stride + chh, ow = jw * a.out_stride + cww;
                const bool inb = valid && oh < a.Ho && ow < a.Wo;
                const size_t vox = (((size_t)b * a.Do + od) * a.Ho + oh) * a.Wo + ow;
                for (int c0 = 0; c0 < Cn; c0 += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) +
                                           (uint32_t)((buf * nM + m) * Cn * a.cblocks + (nblk_e == 8 ? blk * Cn : 0) + c0);
                    __syncwarp();                      // tcgen05.ld is .sync.aligned: whole warp, converged
                    tmem_ld_32x32(taddr, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    // kw-merged accumulators: column block k holds P_k[q] = sum over (kd,kh,ci) for filter column
                    // kw = k evaluated WITHOUT the w shift; out[q] = P_0[q] + P_1[q+1] + P_2[q+2], and q+k is
                    // lane+k of the same warp (one warp = one padded tile row).
                    for (int k = 1; k < merge; ++k) {
                        __syncwarp();
                        tmem_ld_32x32(taddr + (uint32_t)(k * Cn), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] += __shfl_down_sync(0xffffffffu, __uint_as_float(v[i]), k * a.merge_step);
                    }
                    const size_t eoff = vox * ostride_w + a.cout_off + c0;
                    if (inb && full32) {
                        // ---------------- vector path: 32 complete channels
                        if (a.partial) {
                            const float4* pp = reinterpret_cast<const float4*>(a.partial + eoff);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 pv = __ldg(pp + i);
                                f[i * 4] += pv.x; f[i * 4 + 1] += pv.y; f[i * 4 + 2] += pv.z; f[i * 4 + 3] += pv.w;
                            }
                        }
                        if (c0 == 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] += sh0[i];
                        } else if (a.shift) {
                            const float4* sp = reinterpret_cast<const float4*>(a.shift + a.cout_off + c0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 sv = __ldg(sp + i);
                                f[i * 4] += sv.x; f[i * 4 + 1] += sv.y; f[i * 4 + 2] += sv.z; f[i * 4 + 3] += sv.w;
                            }
                        }
                        if (a.residual) {
                            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.residual) + eoff);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 rv = __ldg(rp + i);
                                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 h2 = unpack16(rw[j], f16);
                                    f[i * 8 + j * 2] += h2.x;
                                    f[i * 8 + j * 2 + 1] += h2.y;
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = stb_act(f[i], ACT);
                        if (a.out_fp32) {
                            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + eoff);
#pragma unroll
                            for (int i = 0; i < 8; ++i) op[i] = make_float4(f[i * 4], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
                        } else {
                            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) + eoff);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 o;
                                o.x = pack16(f[i * 8 + 0], f[i * 8 + 1], f16);
                                o.y = pack16(f[i * 8 + 2], f[i * 8 + 3], f16);
                                o.z = pack16(f[i * 8 + 4], f[i * 8 + 5], f16);
                                o.w = pack16(f[i * 8 + 6], f[i * 8 + 7], f16);
                                op[i] = o;
                            }
                        }
                    } else if (inb) {
                        // ---------------- ragged path (Cout not a multiple of 32: the 32->1 classifier, IGEV widths)
                        const int nch = min(32, a.Cn_valid - c0);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < nch) {
                                float x = f[i];
                                if (a.partial) x += __ldg(a.partial + eoff + i);
                                if (a.shift) x += __ldg(a.shift + a.cout_off + c0 + i);
                                if (a.residual) x += load16(reinterpret_cast<const uint16_t*>(a.residual) + eoff + i, f16);
                                x = stb_act(x, ACT);
                                if (a.out_fp32) reinterpret_cast<float*>(a.out)[eoff + i] = x;
                                else store16(reinterpret_cast<uint16_t*>(a.out) + eoff + i, x, f16);
                            }
                        }
                    }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            if (trace) g_umma_trace[round * 8 + 5] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

template <int ACT, bool F16>
int launch_one(unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tx, const CUtensorMap& tw, const UArgs& a) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv3d_umma_kernel<ACT, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP);
        attr_set = true;
    }
    conv3d_umma_kernel<ACT, F16><<<grid, UMMA_THREADS, smem, st>>>(tx, tw, a);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

int launch_umma(int act, int f16, unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tx,
                const CUtensorMap& tw, const UArgs& a) {
#define STB_LAUNCH_ACT(A)                                                         \
    case A: return f16 ? launch_one<A, true>(grid, smem, st, tx, tw, a) : launch_one<A, false>(grid, smem, st, tx, tw, a);
    switch (act) {
        STB_LAUNCH_ACT(STB_ACT_NONE)
        STB_LAUNCH_ACT(STB_ACT_RELU)
        STB_LAUNCH_ACT(STB_ACT_LEAKY)
        STB_LAUNCH_ACT(STB_ACT_MISH)
        default: return STB_E_BADARG;
    }
#undef STB_LAUNCH_ACT
}

}  // namespace

extern "C" int stb_conv3d_umma_set_trace(long long* dev_buf, int rounds) {
    if (cudaMemcpyToSymbol(g_umma_trace, &dev_buf, sizeof(dev_buf)) != cudaSuccess) return STB_E_BADARG;
    if (cudaMemcpyToSymbol(g_umma_trace_rounds, &rounds, sizeof(rounds)) != cudaSuccess) return STB_E_BADARG;
    return STB_OK;
}

// Host entry -- see include/stb200.h for the argument contract.
extern "C" int stb_conv3d_umma(const void* x, const void* wt, const float* shift, const void* residual, void* out,
                               float* ws, int f16, int B, int Cin, int KC, int Di, int Hi, int Wi, int Cout_total,
                               int Cout_valid, int Do, int Ho, int Wo, int ntaps, const int* dz, const int* dh,
                               const int* dw, const int* sub, const int* widx, const int* nblk, const int* cls0,
                               int nwtiles, int nclass,
                               const int* tap_begin, const int* tap_end, const int* od0, const int* oh0,
                               const int* ow0, int in_stride, int out_stride, int nsteps, int nclass_h, int nclass_w,
                               int in_h_off, int in_w_off, int act, int out_fp32, int flags, int dchunk,
                               void* stream) {
    if (!x || !wt || !out || !dz || !dh || !dw || !sub || !widx || !tap_begin || !tap_end || !od0 || !oh0 || !ow0)
        return STB_E_BADARG;
    if (B <= 0 || ntaps <= 0 || ntaps > MAX_UTAPS || nclass <= 0 || nclass > MAX_UCLASS || nsteps <= 0) return STB_E_BADARG;
    if (KC != 16 && KC != 32 && KC != 64) return STB_E_UNSUPPORTED;      // one swizzle row = one voxel's K-chunk
    if (Cin % KC) return STB_E_UNSUPPORTED;
    if (in_stride < 1 || in_stride > 2 || out_stride < 1 || out_stride > 2) return STB_E_UNSUPPORTED;
    if (in_stride == 2 && out_stride != 1) return STB_E_UNSUPPORTED;
    const int nk = Cin / KC;                       // K-split passes
    if (nk > 1 && !ws) return STB_E_BADARG;        // needs the fp32 partial workspace [B,Do,Ho,Wo,Cout_total]
    UArgs a;
    memset(&a, 0, sizeof(a));
    a.f16 = f16;
    a.ROWB = KC * 2;
    a.layout = a.ROWB == 128 ? SW_128B : a.ROWB == 64 ? SW_64B : SW_32B;
    const CUtensorMapSwizzle cusw = a.ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : a.ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUtensorMapDataType cudt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    a.bo_mode = flags & 1;
    const int es_variant = (flags >> 1) & 1;
    a.merge = (flags & 4) ? 3 : 1;            // kw-merged taps: N = 3*Cn, weight tiles (kd,kh,0..2) are contiguous
    a.merge_step = ((flags >> 8) & 7) ? ((flags >> 8) & 7) : 1;   // flags bits 8..10: dilation along w of the merged taps
    if (a.merge == 3 && (in_stride != 1 || out_stride != 1 || nclass != 1)) return STB_E_UNSUPPORTED;
    a.in_stride = in_stride;
    a.nsub = in_stride == 2 ? 4 : 1;
    a.sd_in = (flags & 16) ? 1 : in_stride;     // flags bit4: 2-D convolution, the depth axis (image index) is never strided
    int maxdh = 0, maxdw = 0, dzmin = 127, dzmax = -127;
    for (int t = 0; t < ntaps; ++t) {
        if (dh[t] < 0 || dw[t] < 0 || dh[t] > 6 || dw[t] > 6 || widx[t] < 0 || widx[t] >= nwtiles || sub[t] < 0 ||
            sub[t] >= a.nsub)
            return STB_E_BADARG;
        maxdh = dh[t] > maxdh ? dh[t] : maxdh;
        maxdw = dw[t] > maxdw ? dw[t] : maxdw;
        dzmin = dz[t] < dzmin ? dz[t] : dzmin;
        dzmax = dz[t] > dzmax ? dz[t] : dzmax;
        a.taps[t].dz = (int8_t)dz[t];
        a.taps[t].sub = (uint8_t)sub[t];
        a.taps[t].rowoff = (int16_t)(dh[t] * TWP + dw[t]);
        a.taps[t].widx = (uint16_t)widx[t];
        a.taps[t].nblk = (uint8_t)(nblk ? nblk[t] : a.merge);
        a.taps[t].cls0 = (uint8_t)(cls0 ? cls0[t] : 0);
        if (a.taps[t].nblk < 1 || a.taps[t].nblk + a.taps[t].cls0 > 8) return STB_E_BADARG;
    }
    // accumulator column blocks per M-tile: 8 when taps address parity-class blocks (merged transposed conv),
    // 3 for kw-merge, else 1.  The first tap of every class must cover all blocks (it zero-initialises them).
    a.cblocks = (flags & 8) ? 8 : a.merge;
    for (int c = 0; c < nclass; ++c)
        if (a.taps[tap_begin[c]].nblk != a.cblocks || a.taps[tap_begin[c]].cls0 != 0) return STB_E_BADARG;
    if (a.cblocks == 8 && (nclass != 1 || out_stride != 2 || in_stride != 1)) return STB_E_UNSUPPORTED;
    a.ntaps_total = ntaps;
    for (int c = 0; c < nclass; ++c) {
        a.cls[c].tap_begin = (uint16_t)tap_begin[c];
        a.cls[c].tap_end = (uint16_t)tap_end[c];
        a.cls[c].od0 = (int8_t)od0[c]; a.cls[c].oh0 = (int8_t)oh0[c]; a.cls[c].ow0 = (int8_t)ow0[c];
    }
    a.nclass = nclass;
    a.TW = TWP - (a.merge == 3 ? 2 * a.merge_step : maxdw);
    a.dzmin = dzmin; a.dzmax = dzmax;
    const int window = dzmax - dzmin + 1;
    a.B = B; a.Do = Do; a.Ho = Ho; a.Wo = Wo;
    a.Cout_total = Cout_total;
    a.out_stride = out_stride; a.nsteps = nsteps; a.nclass_h = nclass_h; a.nclass_w = nclass_w;
    a.in_h_off = in_h_off; a.in_w_off = in_w_off;
    const int Cpad = (Cout_total + 15) / 16 * 16;     // weight tiles are padded to a multiple of 16 rows by the packer
    // ---- tiling: pick (Cn, TH, R) that fit shared memory; prefer tall tiles, then wide N
    int Cn = Cpad > 256 ? 256 : Cpad, TH = 0, R = 0;
    // A single in-flight TMA box sustains only a few GB/s per SM (measured: 20-32 KB per ~7-13 us), so the plane
    // ring must keep many planes in flight: choose the tile height whose ring holds the most prefetched bytes
    // (capped at 64 KB), ties to the taller tile.
    auto ring_for = [&](int cn, int th) {
        const int nM = th * TWP / 128;
        const int slack = (cn & 31) ? 32 : 0;                                     // epilogue reads 32-column blocks
        if (2 * nM * cn * a.cblocks + slack > 512 || cn * a.cblocks > 256) return 0;
        const size_t wbytes = (((size_t)nwtiles * cn * a.ROWB) + 1023) & ~(size_t)1023;
        const size_t plane = (size_t)a.nsub * (th + maxdh) * TWP * a.ROWB;
        if (3072 + wbytes + 2048 >= SMEM_CAP) return 0;
        // the per-(phase, tap) issue table grows with the ring: r*(plane + 16*ntaps) + 1 KB rounding slack
        int r = (int)((SMEM_CAP - 3072 - wbytes - 2048) / (plane + (size_t)16 * ntaps));
        if (r > MAX_RING) r = MAX_RING;
        return r >= window + 1 ? r : 0;
    };
    if (window > 6) return STB_E_UNSUPPORTED;
    for (;;) {
        size_t best_score = 0;
        for (int th = 16; th >= 4; th -= 4) {
            const int r = ring_for(Cn, th);
            if (!r) continue;
            const size_t plane = (size_t)a.nsub * (th + maxdh) * TWP * a.ROWB;
            size_t inflight = (size_t)(r - window) * plane;
            if (inflight > 64 * 1024) inflight = 64 * 1024;   // enough to cover TMA latency; beyond that prefer tall tiles
                                                              // (>= 2 M-tiles per round keep both issuers / epilogue groups busy)
            if (inflight > best_score) { best_score = inflight; TH = th; R = r; }
        }
        if (best_score) break;
        if (Cn <= 16) return STB_E_SMEM;
        Cn = (Cn / 2 + 15) / 16 * 16;
    }
    while (TH > 4 && TH - 4 >= nclass_h) TH -= 4;      // do not stage rows that do not exist
    a.R = R;
    a.tab_bytes = (uint32_t)((((size_t)R * ntaps * 16) + 1023) & ~(size_t)1023);
    a.TH = TH; a.nM = TH * TWP / 128;
    const int box_h = TH + maxdh;
    a.chunk_bytes = (uint32_t)(box_h * TWP * a.ROWB);
    a.plane_bytes = (uint32_t)a.nsub * a.chunk_bytes;
    a.tiles_h = stb_ceil_div(nclass_h, a.TH);
    a.tiles_w = stb_ceil_div(nclass_w, a.TW);
    if (dchunk <= 0) {
        // Depth chunking: CTAs = columns x chunks, one CTA per SM at a time.  Pick the chunk count that best
        // fills whole waves of SMs while keeping the redundant halo planes (window-1 per chunk) small.
        static int num_sms = 0;
        if (!num_sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
        }
        const long long cols = (long long)B * a.tiles_h * a.tiles_w;
        double best = -1.0;
        for (int nch = 1; nch <= nsteps; ++nch) {
            const int dch = stb_ceil_div(nsteps, nch);
            if (dch < 4 && nch > 1) break;
            const long long ctas = cols * stb_ceil_div(nsteps, dch);
            const long long waves = (ctas + num_sms - 1) / num_sms;
            const double eff = (double)ctas / (double)(waves * num_sms) * (double)dch / (double)(dch + window - 1);
            if (eff > best + 1e-3) { best = eff; dchunk = dch; }
        }
    }
    a.dchunk = dchunk;
    a.nchunks = stb_ceil_div(nsteps, dchunk);
    int tmem_cols = 32;
    while (tmem_cols < 2 * a.nM * Cn * a.cblocks + ((Cn & 31) ? 32 : 0)) tmem_cols <<= 1;
    a.tmem_cols = (uint32_t)tmem_cols;

    CUtensorMap tm_x, tm_w;
    {
        uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)Di, (uint64_t)B};
        uint64_t str[4] = {(uint64_t)Cin * 2, (uint64_t)Wi * Cin * 2, (uint64_t)Hi * Wi * Cin * 2,
                           (uint64_t)Di * Hi * Wi * Cin * 2};
        const uint32_t bm = (in_stride == 2 && !es_variant) ? 2u : 1u;     // box extent convention under element strides
        uint32_t box[5] = {(uint32_t)KC, (uint32_t)TWP * bm, (uint32_t)box_h * bm, 1, 1};
        uint32_t es[5] = {1, (uint32_t)in_stride, (uint32_t)in_stride, 1, 1};
        if (!umma_host::make_tmap(&tm_x, cudt, 5, const_cast<void*>(x), dims, str, box, cusw, es)) return STB_E_DRIVER;
    }
    a.w_rows = Cpad;
    a.w_tile_stride = nk;
    a.nwtiles = nwtiles;
    const long long ncta = (long long)B * a.nchunks * a.tiles_h * a.tiles_w;
    if (ncta > 2147483647LL) return STB_E_BADARG;
    for (int kp = 0; kp < nk; ++kp) {
        const bool last = kp == nk - 1;
        a.cin_off = kp * KC;
        a.w_kc_off = kp;
        a.partial = kp > 0 ? ws : nullptr;
        a.out = last ? out : (void*)ws;
        a.out_fp32 = last ? out_fp32 : 1;
        a.shift = last ? shift : nullptr;
        a.residual = last ? residual : nullptr;
        a.act = last ? act : STB_ACT_NONE;
        for (int co = 0; co < Cpad; co += Cn) {
            const int cn = (Cpad - co) < Cn ? (Cpad - co) : Cn;
            a.Cn = cn;
            a.cout_off = co;
            a.Cn_valid = (Cout_valid - co) < cn ? (Cout_valid - co) : cn;
            if (a.Cn_valid <= 0) break;
            a.wtile_bytes = (uint32_t)(cn * a.ROWB);
            a.w_bytes_total = (uint32_t)((size_t)nwtiles * a.wtile_bytes);
            uint64_t dims[2] = {(uint64_t)KC, (uint64_t)nwtiles * nk * Cpad};
            uint64_t str[1] = {(uint64_t)KC * 2};
            uint32_t box[2] = {(uint32_t)KC, (uint32_t)cn};
            if (!umma_host::make_tmap(&tm_w, cudt, 2, const_cast<void*>(wt), dims, str, box, cusw)) return STB_E_DRIVER;
            size_t smem = 3072 + a.tab_bytes + (((size_t)a.w_bytes_total + 1023) & ~(size_t)1023) + (size_t)a.R * a.plane_bytes + 1024;
            if (smem > SMEM_CAP) return STB_E_SMEM;
            const int rc = launch_umma(a.act, f16, (unsigned)ncta, smem, (cudaStream_t)stream, tm_x, tm_w, a);
            if (rc != STB_OK) return rc;
        }
    }
    return STB_OK;
}
