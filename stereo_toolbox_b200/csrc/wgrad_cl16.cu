// Weight gradient of the 3-D convolution family on channels-last 16-bit tensors (training path, BASELINE config 3).
//
//   dW[t][cp][cq] = sum over (b, q)  P[b][s*q + k(t) - pad][cp] * Q[b][q][cq]        t = (kd,kh,kw)
//
// Conv3d (PSMNet/submodule.py:16-19):   P = layer input x, Q = grad of the layer output, s = conv stride
// ConvTranspose3d(k3,s2,p1,op1) (PSMNet/stackhourglass.py:25-29): P = grad of the output, Q = layer input, s = 2
// (the transposed conv is the adjoint of the strided conv, so the same index relation holds with the roles swapped).
//
// The contraction runs over POSITIONS, which are the slow axis of NDHWC tensors, so both operands are "MN-major";
// this version uses the warp-level tensor path (mma.sync m16n8k16, ldmatrix.trans from padded smem rows) --
// SASS: HMMA + LDSM.  A persistent CTA keeps up to 32 output blocks of 32x32 (tap, cp-block, cq-block) in registers
// (4 per warp), streams (P halo tile, Q tile) pairs through a cp.async double buffer and adds its partial sums to the
// fp32 result with atomics once at the end.  Why not tcgen05 here: with K = positions both operands are MN-major, the
// 32-channel layers (most of PSMNet's FLOPs) give M = 32 < the UMMA minimum of 64, and the junk columns of the padded
// position tiles would have to be zeroed in shared memory before every MMA.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

// 8 warps x 4 output blocks.  (16 warps x 2 blocks at 128 registers was measured too, on the hypothesis that two warps per
// scheduler cannot hide the fixed HMMA / LDSM latencies -- ncu: issue slots 36 % busy, "wait" 2.4 per issue, HMMA pipe 30 % --:
// 11.3 ms instead of 10.9 ms for the 28 launches of a PSMNet step, so occupancy is not what limits it.)
constexpr int WG_THREADS = 256;
constexpr int WG_TW = 32;        // q positions per tile row
constexpr int WG_UPW = 4;        // 32x32 output blocks per warp
constexpr int WG_PADB = 16;      // smem row padding in bytes (conflict-free ldmatrix)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WgradArgs {
    const uint16_t* P;
    const uint16_t* Q;
    float* dW;
    int B, Dp, Hp, Wp, Cp;       // P tensor [B][Dp][Hp][Wp][Cp]
    int Dq, Hq, Wq, Cq;          // Q tensor [B][Dq][Hq][Wq][Cq]
    int K, pad, s;               // kernel size per dim, padding, stride (p = s*q + k - pad)
    int TH;                      // q rows per tile
    int PH, PW;                  // halo tile rows / cols
    int cps, cqs;                // channel slice widths handled by one CTA (<= 64, multiples of 32)
    int nps, nqs;                // slices per tensor
    int taps_per_cta;            // taps per CTA group
    int ntapgrp;
    int tiles_h, tiles_w;
    long long ntiles;
    int nstage;
};

template <bool F16>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const WgradArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ppitch = a.cps * 2 + WG_PADB, qpitch = a.cqs * 2 + WG_PADB;       // bytes per smem row
    const int prow = a.K * a.PH * a.PW, qrow = a.TH * WG_TW;
    const size_t pbytes = (size_t)prow * ppitch, qbytes = (size_t)qrow * qpitch;
    const size_t stage_bytes = (pbytes + qbytes + 127) / 128 * 128;
    // ---- which output blocks this CTA / warp owns
    int grp = blockIdx.y;
    const int tapgrp = grp % a.ntapgrp; grp /= a.ntapgrp;
    const int qsl = grp % a.nqs, psl = grp / a.nqs;
    const int ntaps = a.K * a.K * a.K;
    const int tap0 = tapgrp * a.taps_per_cta;
    const int tap1 = min(ntaps, tap0 + a.taps_per_cta);
    const int pbn = a.cps / 32, qbn = a.cqs / 32;
    const int nunits = (tap1 - tap0) * pbn * qbn;
    int u_tap[WG_UPW], u_pb[WG_UPW], u_qb[WG_UPW], u_poff[WG_UPW];
    bool u_ok[WG_UPW];
#pragma unroll
    for (int i = 0; i < WG_UPW; ++i) {
        const int u = warp * WG_UPW + i;
        u_ok[i] = u < nunits;
        const int uu = u_ok[i] ? u : 0;
        u_tap[i] = tap0 + uu / (pbn * qbn);
        u_pb[i] = (uu / qbn) % pbn;
        u_qb[i] = uu % qbn;
        const int t = u_tap[i];
        const int kd = t / (a.K * a.K), kh = (t / a.K) % a.K, kw = t % a.K;
        u_poff[i] = (kd * a.PH + kh) * a.PW + kw;            // halo-tile row of the tap's (0, 0) position
    }
    float acc[WG_UPW][2][4][4];
#pragma unroll
    for (int i = 0; i < WG_UPW; ++i)
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][m][n][e] = 0.f;

    const int pchunks = a.cps / 8, qchunks = a.cqs / 8;      // 16-byte chunks per row
    const uint16_t* Pg = a.P + (size_t)psl * a.cps;
    const uint16_t* Qg = a.Q + (size_t)qsl * a.cqs;

    auto load_tile = [&](long long tile, int stage) {
        unsigned char* sp = smem + (size_t)stage * stage_bytes;
        unsigned char* sq = sp + pbytes;
        long long t = tile;
        const int tw = (int)(t % a.tiles_w); t /= a.tiles_w;
        const int th = (int)(t % a.tiles_h); t /= a.tiles_h;
        const int dq = (int)(t % a.Dq);
        const int b = (int)(t / a.Dq);
        const int h0 = th * a.TH, w0 = tw * WG_TW;
        for (int i = tid; i < prow * pchunks; i += WG_THREADS) {
            const int row = i / pchunks, ch = i - row * pchunks;
            const int pc = row % a.PW;
            const int r2 = row / a.PW;
            const int pr = r2 % a.PH, kd = r2 / a.PH;
            const int dp = a.s * dq + kd - a.pad, hp = a.s * h0 + pr - a.pad, wp = a.s * w0 + pc - a.pad;
            const bool ok = dp >= 0 && dp < a.Dp && hp >= 0 && hp < a.Hp && wp >= 0 && wp < a.Wp;
            const uint16_t* src = ok ? Pg + ((((size_t)b * a.Dp + dp) * a.Hp + hp) * a.Wp + wp) * a.Cp + ch * 8 : Pg;
            cp_async16(s_u32(sp + (size_t)row * ppitch + ch * 16), src, ok);
        }
        for (int i = tid; i < qrow * qchunks; i += WG_THREADS) {
            const int row = i / qchunks, ch = i - row * qchunks;
            const int qr = row / WG_TW, qc = row - qr * WG_TW;
            const int hq = h0 + qr, wq = w0 + qc;
            const bool ok = hq < a.Hq && wq < a.Wq;
            const uint16_t* src = ok ? Qg + ((((size_t)b * a.Dq + dq) * a.Hq + hq) * a.Wq + wq) * a.Cq + ch * 8 : Qg;
            cp_async16(s_u32(sq + (size_t)row * qpitch + ch * 16), src, ok);
        }
    };

    // per-lane ldmatrix row/column selectors (see the fragment layout notes in the header comment of mma16816)
    const int a_krow = (lane & 7) + ((lane >> 4) << 3);      // k offset supplied by this lane (A operand)
    const int a_mcol = ((lane >> 3) & 1) << 3;               // m offset (channels)
    const int b_krow = (lane & 7) + (((lane >> 3) & 1) << 3);
    const int b_ncol = (lane >> 4) << 3;

    long long tile = blockIdx.x;
    int stage = 0;
    if (tile < a.ntiles) load_tile(tile, 0);
    cp_async_commit();
    for (; tile < a.ntiles; tile += gridDim.x) {
        const long long next = tile + gridDim.x;
        if (a.nstage > 1) {
            if (next < a.ntiles) load_tile(next, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t sp = s_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sq = sp + (uint32_t)pbytes;
        for (int ks = 0; ks < a.TH * 2; ++ks) {
            const int r = ks >> 1, c0 = (ks & 1) << 4;
#pragma unroll
            for (int i = 0; i < WG_UPW; ++i) {
                if (!u_ok[i]) continue;                       // warp-uniform
                // B fragments: Q rows r*32 + c0 + k, channels qb*32 + n
                uint32_t bfr[4][2];
                {
                    const uint32_t base = sq + (uint32_t)((r * WG_TW + c0 + b_krow) * qpitch + (u_qb[i] * 32 + b_ncol) * 2);
                    ldsm_x4_t(base, bfr[0][0], bfr[0][1], bfr[1][0], bfr[1][1]);
                    ldsm_x4_t(base + 32, bfr[2][0], bfr[2][1], bfr[3][0], bfr[3][1]);
                }
                // A fragments: P halo rows (kd, s*r + kh, s*(c0 + k) + kw), channels pb*32 + m
                const int prow_i = u_poff[i] + a.s * r * a.PW + a.s * (c0 + a_krow);
                const uint32_t abase = sp + (uint32_t)(prow_i * ppitch + (u_pb[i] * 32 + a_mcol) * 2);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    uint32_t afr[4];
                    ldsm_x4_t(abase + m * 32, afr[0], afr[1], afr[2], afr[3]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma16816<F16>(acc[i][m][n], afr, bfr[n][0], bfr[n][1]);
                }
            }
        }
        __syncthreads();
        if (a.nstage > 1) stage ^= 1;
        else {
            if (next < a.ntiles) load_tile(next, 0);
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
    // ---- partial sums -> dW[t][cp][cq] (fp32 atomics; the caller zeroed dW)
#pragma unroll
    for (int i = 0; i < WG_UPW; ++i) {
        if (!u_ok[i]) continue;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = psl * a.cps + u_pb[i] * 32 + m * 16 + (lane >> 2) + ((e >> 1) << 3);
                    const int col = qsl * a.cqs + u_qb[i] * 32 + n * 8 + ((lane & 3) << 1) + (e & 1);
                    atomicAdd(a.dW + ((size_t)u_tap[i] * a.Cp + row) * a.Cq + col, acc[i][m][n][e]);
                }
    }
}

}  // namespace

extern "C" int stb_conv3d_wgrad_cl16(const void* P, const void* Q, float* dW, int f16, int B, int Dp, int Hp, int Wp,
                                     int Cp, int Dq, int Hq, int Wq, int Cq, int K, int pad, int stride, int max_ctas,
                                     void* stream) {
    if (!P || !Q || !dW || B <= 0 || Cp <= 0 || Cq <= 0) return STB_E_BADARG;
    if (Cp % 32 || Cq % 32 || K < 1 || K > 3 || stride < 1 || stride > 2) return STB_E_UNSUPPORTED;
    WgradArgs a;
    a.P = (const uint16_t*)P; a.Q = (const uint16_t*)Q; a.dW = dW;
    a.B = B; a.Dp = Dp; a.Hp = Hp; a.Wp = Wp; a.Cp = Cp; a.Dq = Dq; a.Hq = Hq; a.Wq = Wq; a.Cq = Cq;
    a.K = K; a.pad = pad; a.s = stride;
    a.cps = Cp % 64 == 0 ? 64 : 32;
    a.cqs = Cq % 64 == 0 ? 64 : 32;
    a.nps = Cp / a.cps; a.nqs = Cq / a.cqs;
    const int blocks_per_tap = (a.cps / 32) * (a.cqs / 32);
    const int ntaps = K * K * K;
    a.taps_per_cta = (WG_THREADS / 32) * WG_UPW / blocks_per_tap;
    if (a.taps_per_cta > ntaps) a.taps_per_cta = ntaps;
    a.ntapgrp = (ntaps + a.taps_per_cta - 1) / a.taps_per_cta;
    a.taps_per_cta = (ntaps + a.ntapgrp - 1) / a.ntapgrp;           // balance the groups
    const size_t limit = 220 * 1024;
    int best_th = 0, best_stage = 0;
    for (int st = 2; st >= 1 && !best_th; --st)          // a double buffer first (loads overlap the MMAs), then the tallest tile
        for (int th = 4; th >= 1; th >>= 1) {
            const int ph = stride * (th - 1) + K, pw = stride * (WG_TW - 1) + K;
            size_t pb = (size_t)K * ph * pw * (a.cps * 2 + WG_PADB), qb = (size_t)th * WG_TW * (a.cqs * 2 + WG_PADB);
            size_t sb = (pb + qb + 127) / 128 * 128;
            if (sb * st <= limit) { best_th = th; best_stage = st; break; }
        }
    if (!best_th) return STB_E_SMEM;
    a.TH = best_th; a.nstage = best_stage;
    a.PH = stride * (a.TH - 1) + K; a.PW = stride * (WG_TW - 1) + K;
    a.tiles_h = stb_ceil_div(Hq, a.TH); a.tiles_w = stb_ceil_div(Wq, WG_TW);
    a.ntiles = (long long)B * Dq * a.tiles_h * a.tiles_w;
    const size_t stage_bytes = ((size_t)K * a.PH * a.PW * (a.cps * 2 + WG_PADB) +
                                (size_t)a.TH * WG_TW * (a.cqs * 2 + WG_PADB) + 127) / 128 * 128;
    const size_t smem = stage_bytes * a.nstage;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int ngroups = a.nps * a.nqs * a.ntapgrp;
    long long gx = max_ctas > 0 ? max_ctas : nsm;
    if (ngroups > 1 && max_ctas <= 0) gx = (nsm + ngroups - 1) / ngroups * 2;    // keep every SM busy, few atomics
    if (gx > a.ntiles) gx = a.ntiles;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)ngroups);
    if (f16) {
        cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wgrad_kernel<true><<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(a);
    } else {
        cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wgrad_kernel<false><<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(a);
    }
    STB_CHECK_LAUNCH();
    return STB_OK;
}
