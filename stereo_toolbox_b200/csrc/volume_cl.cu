// Cost-volume builder for the tensor-core path: gwc (+ concat) volume written ONCE, channels-last
// bf16  vol[b][d][h][w][Ct]  (Ct = G + 2*Cc, zero-padded to Ct_pad), i.e. directly in the layout the
// first 3x3x3 conv TMA-loads.  Replaces build_gwc_volume + build_concat_volume + torch.cat
// (GwcNet/submodule.py:30-63, GwcNet/gwcnet.py:175-180; inline loop PSMNet/stackhourglass.py:111-120).
// Products and the group mean are fp32; only the stored value is rounded to bf16.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t pk16(float a, float b, int f16) {
    if (f16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint16_t cv16(float a, int f16) {
    if (f16) { __half v = __float2half_rn(a); return *reinterpret_cast<uint16_t*>(&v); }
    __nv_bfloat16 v = __float2bfloat16(a);
    return *reinterpret_cast<uint16_t*>(&v);
}
__device__ __forceinline__ float ld16(uint16_t w, int f16) {
    if (f16) return __half2float(*reinterpret_cast<const __half*>(&w));
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&w));
}

__device__ __forceinline__ int vsplit_idx(int c) { return ((c >> 4) << 5) + (c & 15); }   // storage index of channel c in a split row

constexpr int VT = 256;      // threads
constexpr int VW = 32;       // w positions per CTA

// grid (tiles_w, H, B). dynamic smem: NR*(VW+1) + NR*(VW+D) floats, NR = max(8, 8*cpg) staged rows
// CPG > 0: channels per correlation group known at compile time -> the left features of a chunk live in
// registers and every (d, w) costs one shared-memory load per MAC instead of two.
template <int CPG>
__global__ void __launch_bounds__(VT)
volume_cl_kernel(const float* __restrict__ gl, const float* __restrict__ gr, const float* __restrict__ cl,
                 const float* __restrict__ cr, uint16_t* __restrict__ vol, int Cg, int G, int Cc, int H, int W,
                 int D, int Ct, int Ct_pad, int mask_left, int NR, int f16) {
    extern __shared__ __align__(16) float sm[];
    const int LP = VW + 1, RP = VW + D;           // row pitches; R window covers w in [w0-D+1, w0+VW)
    float* Ls = sm;                                // [NR][LP]
    float* Rs = sm + NR * LP;                      // [NR][RP]
    const int w0 = blockIdx.x * VW, h = blockIdx.y, b = blockIdx.z;
    const int cpg = G > 0 ? Cg / G : 1;
    const size_t plane = (size_t)H * W;
    const int lw = threadIdx.x & 31, dgrp = threadIdx.x >> 5;   // 8 depth groups
    const int w = w0 + lw;
    const int nchunk = Ct_pad / 8;
    for (int ch = 0; ch < nchunk; ++ch) {
        const int oc0 = ch * 8;
        __syncthreads();
        // ---- stage the source rows of this chunk
        const bool gwc_chunk = oc0 + 8 <= G;       // all 8 output channels are correlation groups
        const int nrow = gwc_chunk ? 8 * cpg : 8;
        if (gwc_chunk) {
            if (nrow > NR) return;                 // guarded on the host
            const float* lsrc = gl + ((size_t)b * Cg + (size_t)oc0 * cpg) * plane + (size_t)h * W;
            const float* rsrc = gr + ((size_t)b * Cg + (size_t)oc0 * cpg) * plane + (size_t)h * W;
            for (int i = threadIdx.x; i < nrow * VW; i += VT) {
                int r = i / VW, x = i - r * VW;
                Ls[r * LP + x] = (w0 + x < W) ? __ldg(lsrc + (size_t)r * plane + w0 + x) : 0.f;
            }
            for (int i = threadIdx.x; i < nrow * (VW + D - 1); i += VT) {
                int r = i / (VW + D - 1), x = i - r * (VW + D - 1);
                int ww = w0 - (D - 1) + x;
                Rs[r * RP + x] = (ww >= 0 && ww < W) ? __ldg(rsrc + (size_t)r * plane + ww) : 0.f;
            }
        } else {
            // mixed chunk: per output channel pick its source row (gwc group rows are not mixed with concat
            // rows by construction: G % 8 == 0 is required on the host)
            for (int i = threadIdx.x; i < 8 * (VW + D - 1); i += VT) {
                int r = i / (VW + D - 1), x = i - r * (VW + D - 1);
                int ww = w0 - (D - 1) + x;
                int j = oc0 + r - G;
                float v = 0.f;
                if (j >= 0 && j < 2 * Cc && ww >= 0 && ww < W) {
                    const float* src = j < Cc ? cl + ((size_t)b * Cc + j) * plane : cr + ((size_t)b * Cc + (j - Cc)) * plane;
                    v = __ldg(src + (size_t)h * W + ww);
                }
                Rs[r * RP + x] = v;
            }
        }
        __syncthreads();
        if (w >= W) continue;
        if (CPG > 0 && gwc_chunk) {
            float lreg[8 * (CPG > 0 ? CPG : 1)];
#pragma unroll
            for (int r = 0; r < 8 * CPG; ++r) lreg[r] = Ls[r * LP + lw];
            const float inv = 1.f / (float)CPG;
            for (int d = dgrp; d < D; d += VT / 32) {
                const float* rrow = Rs + lw + (D - 1) - d;
                float o[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < CPG; ++c) acc = fmaf(lreg[g * CPG + c], rrow[(g * CPG + c) * RP], acc);
                    o[g] = acc * inv;
                }
                uint4 pk;
                pk.x = pk16(o[0], o[1], f16); pk.y = pk16(o[2], o[3], f16);
                pk.z = pk16(o[4], o[5], f16); pk.w = pk16(o[6], o[7], f16);
                uint16_t* dst = vol + ((((size_t)b * D + d) * H + h) * W + w) * Ct_pad + oc0;
                *reinterpret_cast<uint4*>(dst) = pk;
            }
            continue;
        }
        for (int d = dgrp; d < D; d += VT / 32) {
            float o[8];
            if (gwc_chunk) {
                const float inv = 1.f / (float)cpg;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float acc = 0.f;
                    for (int c = 0; c < cpg; ++c) {
                        const int r = g * cpg + c;
                        acc = fmaf(Ls[r * LP + lw], Rs[r * RP + lw + (D - 1) - d], acc);
                    }
                    o[g] = acc * inv;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int j = oc0 + r - G;
                    float v = 0.f;
                    if (j >= 0 && j < Cc) v = (!mask_left || w >= d) ? Rs[r * RP + lw + (D - 1)] : 0.f;
                    else if (j >= Cc && j < 2 * Cc) v = Rs[r * RP + lw + (D - 1) - d];   // zero-filled where w-d < 0
                    o[r] = v;
                }
            }
            uint4 pk;
            pk.x = pk16(o[0], o[1], f16); pk.y = pk16(o[2], o[3], f16);
            pk.z = pk16(o[4], o[5], f16); pk.w = pk16(o[6], o[7], f16);
            uint16_t* dst = vol + ((((size_t)b * D + d) * H + h) * W + w) * Ct_pad + oc0;
            *reinterpret_cast<uint4*>(dst) = pk;
        }
    }
}


// ---- v2 builder: full-sector stores + LDS.128 register tiles -------------------------------------------------
// The v1 kernel above stores 16 B per voxel per pass (half a 32-B sector, 128 B apart between lanes) and issues one
// scalar LDS per MAC; ncu (profiles/ncu_launches_r01.txt) had it at 1.87 ms for 2.1 GB of compulsory traffic.
// v2: CTA = (32-wide w tile, h, b).  Output channels are produced 16 at a time (32 B per voxel = one full sector);
// their source rows (16*CPG left rows of 32 floats, 16*CPG right rows covering w0-D4 .. w0+31) are staged in smem.
// A thread owns a 4(w) x 4(d) register tile of TWO adjacent output channels: per source row it needs 7 consecutive
// right values (two aligned LDS.128) for 16 FMAs, because R[w-d] is constant along the (w,d) diagonal.  Results are
// packed to 16-bit pairs into a padded smem tile and written out as 32-B runs per voxel.
constexpr int V2_TW = 32;      // w per CTA
constexpr int V2_DP = 16;      // depths per pass (4 e-blocks of 4)

__device__ __forceinline__ void vcl_cp16(float* dst_smem, const float* src, bool valid) {
    const int sz = valid ? 16 : 0;                   // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src), "r"(sz)
                 : "memory");
}

// Channels-last 16-bit feature sources (CLSRC): the tensor-core extractor's own outputs [2B][H][W][C_i] (left = image b,
// right = image B + b) are read directly -- no NCHW fp32 copy of the 320-channel feature, no torch.cat of layer2/3/4.
struct ClSrc {
    const uint16_t* f[4];     // up to 4 source tensors whose channels concatenate to the Cg correlation channels
    int c[4], c0[4];          // channels of tensor i, first concatenated channel it holds (multiples of 8)
    int n;
    const uint16_t* cat;      // concat feature [2B][H][W][cat_c] (Cc real channels), nullable
    int cat_c, Bn;
};

__device__ __forceinline__ void unpack8(const uint4& q, int f16, float (&v)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t;
        if (f16) t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        else t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
}

// SPLIT: the volume (and, with CLSRC, the sources) is operand-split fp16 -- fp16 hi + fp16 lo per value, interleaved per 16
// channels (conv3d_umma.cu) -- so a 16-channel pass writes 64 contiguous bytes per voxel [hi 0..15 | lo 0..15].
template <int CPG, bool CACHE_L, bool CLSRC, bool SPLIT>
__global__ void __launch_bounds__(256, 2)
volume_cl2_kernel(const float* __restrict__ gl, const float* __restrict__ gr, const float* __restrict__ cl,
                  const float* __restrict__ cr, uint16_t* __restrict__ vol, int Cg, int G, int Cc, int H, int W,
                  int D, int Ct_pad, int mask_left, int f16, const ClSrc cs) {
    extern __shared__ __align__(16) float sm[];
    const int E = (D + 3) >> 2, D4 = E * 4, RP = D4 + V2_TW;
    constexpr int NR = 16 * CPG;
    // Channels-last sources: the 16 groups of a pass are contiguous in a voxel's row, so consecutive lanes of the staging loop
    // take consecutive GROUPS of one voxel (8 lines per warp load instead of 32: ncu counted as many LSU wavefronts for these
    // scattered 16-byte loads as for all shared-memory traffic).  The staged rows are then channel-major (row = c*16 + group)
    // with a pitch of +4 floats, so that the 16 lanes of a voxel scatter to 8 different banks.
    constexpr int LPIT = CLSRC ? V2_TW + 4 : V2_TW;   // pitch of the left rows
    const int RPIT = CLSRC ? RP + 4 : RP;             // pitch of the right rows
#define VROW(ol, c) (CLSRC ? (c) * 16 + (ol) : (ol) * CPG + (c))
    float* Ls = sm;                                   // [NR][LPIT]
    float* Rs = sm + NR * LPIT;                       // [NR][RPIT]   x <-> ww = w0 - D4 + x
    constexpr int OP = SPLIT ? 17 : 9;                // words per voxel in the output tile (+1 pad)
    uint32_t* Os = reinterpret_cast<uint32_t*>(Rs + (size_t)NR * RPIT);   // [V2_DP][32][OP] packed channel pairs (SPLIT: 8 hi words, 8 lo words)
    const int w0 = blockIdx.x * V2_TW, h = blockIdx.y, b = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane & 7;                           // w block: w = w0 + 4j + ww
    const int gp = (warp & 1) * 4 + (lane >> 3);      // channel pair inside the 16-channel group
    const int el = warp >> 1;                         // e block inside a pass
    const float inv = 1.f / (float)CPG;

    for (int oc0 = 0; oc0 < Ct_pad; oc0 += 16) {
        __syncthreads();
        // ---- stage source rows: one warp per row.  W % 4 == 0: 16-byte cp.async (zero-fill outside the image), all
        // rows of the group in flight at once -- the scalar LDG->STS loop of the first v2 draft exposed one global
        // latency per row (ncu: 47 % of the stall samples on the staging STS, long scoreboard).
        const bool vec = (W & 3) == 0;
        if (CLSRC) {
            // one 16-byte load = the 8 channels of one correlation group at one voxel; scattered (as fp32) to the 8 source
            // rows of that group.  Loads of a batch are issued before the first store.
            static_assert(!CLSRC || CPG == 8, "channels-last sources: 8 channels per group");
            constexpr int VB = SPLIT ? 2 : 4;       // loads in flight per thread (split rows need two 16-byte loads each: 4 spilled)
            const int span = V2_TW + RP, total = 16 * span;
            for (int i0 = threadIdx.x; i0 < total; i0 += 256 * VB) {
                uint4 q[VB], ql[SPLIT ? VB : 1];
                int kind_[VB];
#pragma unroll
                for (int u = 0; u < VB; ++u) {
                    const int i = i0 + u * 256;
                    q[u] = make_uint4(0, 0, 0, 0);
                    if (SPLIT) ql[SPLIT ? u : 0] = make_uint4(0, 0, 0, 0);
                    kind_[u] = -1;
                    if (i >= total) continue;
                    const int x = i >> 4, ol = i & 15, o = oc0 + ol;          // group fastest: a voxel's 16 groups are contiguous
                    const bool isL = x < V2_TW;
                    const int ww = isL ? w0 + x : w0 - D4 + (x - V2_TW);
                    const bool ok = ww >= 0 && ww < W;
                    const size_t vox = ((size_t)(isL ? b : cs.Bn + b) * H + h) * W + (ok ? ww : 0);
                    if (o < G) {
                        kind_[u] = 0;
                        if (ok) {
                            const int ch0 = o * 8;
                            int t = 0;
#pragma unroll
                            for (int k = 1; k < 4; ++k) t = (k < cs.n && ch0 >= cs.c0[k]) ? k : t;
                            if (SPLIT) {
                                const uint16_t* src = cs.f[t] + vox * (2 * cs.c[t]) + vsplit_idx(ch0 - cs.c0[t]);
                                q[u] = __ldg(reinterpret_cast<const uint4*>(src));
                                ql[SPLIT ? u : 0] = __ldg(reinterpret_cast<const uint4*>(src + 16));
                            } else
                            q[u] = __ldg(reinterpret_cast<const uint4*>(cs.f[t] + vox * cs.c[t] + (ch0 - cs.c0[t])));
                        }
                    } else if (o < G + Cc ? isL : (o < G + 2 * Cc && !isL)) {
                        kind_[u] = 1;
                        const int jc = o < G + Cc ? o - G : o - G - Cc;
                        if (ok && SPLIT) {
                            const uint16_t* src = cs.cat + vox * cs.cat_c + vsplit_idx(jc);
                            q[u].x = __ldg(src);
                            q[u].y = __ldg(src + 16);
                        } else if (ok) q[u].x = __ldg(cs.cat + vox * cs.cat_c + jc);
                    }
                }
#pragma unroll
                for (int u = 0; u < VB; ++u) {
                    const int i = i0 + u * 256;
                    if (kind_[u] < 0) continue;
                    const int x = i >> 4, ol = i & 15;
                    const bool isL = x < V2_TW;
                    float* dst = isL ? Ls + (size_t)VROW(ol, 0) * LPIT + x : Rs + (size_t)VROW(ol, 0) * RPIT + (x - V2_TW);
                    const int pitch = (isL ? LPIT : RPIT) * (CLSRC ? 16 : 1);          // distance between the rows of channel c and c + 1
                    if (kind_[u] == 0) {
                        float v[8];
                        unpack8(q[u], f16, v);
                        if (SPLIT) {
                            float vl[8];
                            unpack8(ql[SPLIT ? u : 0], 1, vl);
#pragma unroll
                            for (int c = 0; c < 8; ++c) v[c] += vl[c];
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) dst[c * pitch] = v[c];
                    } else {
                        dst[0] = ld16((uint16_t)q[u].x, f16) + (SPLIT ? ld16((uint16_t)q[u].y, 1) : 0.f);
                    }
                }
            }
        }
        for (int r = warp; r < NR && !CLSRC; r += 8) {
            const int ol = r / CPG, c = r - ol * CPG, o = oc0 + ol;
            const float* lsrc = nullptr;
            const float* rsrc = nullptr;
            if (o < G) {
                lsrc = gl + ((size_t)b * Cg + (size_t)o * CPG + c) * plane + (size_t)h * W;
                rsrc = gr + ((size_t)b * Cg + (size_t)o * CPG + c) * plane + (size_t)h * W;
            } else if (c == 0 && o < G + Cc) {
                lsrc = cl + ((size_t)b * Cc + (o - G)) * plane + (size_t)h * W;
            } else if (c == 0 && o < G + 2 * Cc) {
                rsrc = cr + ((size_t)b * Cc + (o - G - Cc)) * plane + (size_t)h * W;
            }
            if (vec) {
                if (lsrc && lane < 8) {
                    const int ww = w0 + 4 * lane;
                    vcl_cp16(Ls + r * LPIT + 4 * lane, lsrc + (ww < W ? ww : 0), ww < W);
                }
                if (rsrc)
                    for (int x4 = lane; x4 < (RP >> 2); x4 += 32) {
                        const int ww = w0 - D4 + 4 * x4;
                        const bool ok = ww >= 0 && ww < W;
                        vcl_cp16(Rs + (size_t)r * RPIT + 4 * x4, rsrc + (ok ? ww : 0), ok);
                    }
            } else {
                if (lsrc) Ls[r * LPIT + lane] = (w0 + lane < W) ? __ldg(lsrc + w0 + lane) : 0.f;
                if (rsrc)
                    for (int x = lane; x < RP; x += 32) {
                        const int ww = w0 - D4 + x;
                        Rs[(size_t)r * RPIT + x] = (ww >= 0 && ww < W) ? __ldg(rsrc + ww) : 0.f;
                    }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        int kind[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int o = oc0 + 2 * gp + q;
            kind[q] = o < G ? 0 : (o < G + Cc ? 1 : (o < G + 2 * Cc ? 2 : 3));
        }
        float4 Lc[2][CACHE_L ? CPG : 1];
        if (CACHE_L) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int c = 0; c < CPG; ++c)
                    if (kind[q] == 0)
                        Lc[q][CACHE_L ? c : 0] = *reinterpret_cast<const float4*>(Ls + VROW(2 * gp + q, c) * LPIT + 4 * j);
        }
        for (int d0 = 0; d0 < D; d0 += V2_DP) {
            const int e = (d0 >> 2) + el;
            uint32_t pk[4][4], pl[SPLIT ? 4 : 1][4];
            if (e < E) {
                float acc[2][4][4];
                const int x0 = D4 + 4 * (j - e) - 4;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                        for (int ww = 0; ww < 4; ++ww) acc[q][dd][ww] = 0.f;
                    const int og = 2 * gp + q;               // output channel (group) inside the pass
                    if (kind[q] == 0) {
#pragma unroll
                        for (int c = 0; c < CPG; ++c) {
                            const float4 l4 = CACHE_L ? Lc[q][CACHE_L ? c : 0]
                                                      : *reinterpret_cast<const float4*>(Ls + VROW(og, c) * LPIT + 4 * j);
                            const float* rr = Rs + (size_t)VROW(og, c) * RPIT + x0;
                            const float4 ra = *reinterpret_cast<const float4*>(rr);
                            const float4 rb = *reinterpret_cast<const float4*>(rr + 4);
                            const float v[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                            const float l[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                            for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                                for (int ww = 0; ww < 4; ++ww) acc[q][dd][ww] = fmaf(l[ww], v[4 + ww - dd], acc[q][dd][ww]);
                        }
#pragma unroll
                        for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                            for (int ww = 0; ww < 4; ++ww) acc[q][dd][ww] *= inv;
                    } else if (kind[q] == 1) {
                        const float4 l4 = *reinterpret_cast<const float4*>(Ls + VROW(og, 0) * LPIT + 4 * j);
                        const float l[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                        for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                            for (int ww = 0; ww < 4; ++ww)
                                acc[q][dd][ww] = (mask_left && (w0 + 4 * j + ww) < (4 * e + dd)) ? 0.f : l[ww];
                    } else if (kind[q] == 2) {
                        const float* rr = Rs + (size_t)VROW(og, 0) * RPIT + x0;
                        const float4 ra = *reinterpret_cast<const float4*>(rr);
                        const float4 rb = *reinterpret_cast<const float4*>(rr + 4);
                        const float v[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                        for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                            for (int ww = 0; ww < 4; ++ww) acc[q][dd][ww] = v[4 + ww - dd];
                    }
                }
#pragma unroll
                for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                    for (int ww = 0; ww < 4; ++ww) {
                        pk[dd][ww] = pk16(acc[0][dd][ww], acc[1][dd][ww], f16);
                        if (SPLIT) {
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&pk[dd][ww]));
                            pl[SPLIT ? dd : 0][ww] = pk16(acc[0][dd][ww] - hf.x, acc[1][dd][ww] - hf.y, 1);
                        }
                    }
            }
            __syncthreads();          // previous pass's write-out has finished reading Os
            if (e < E) {
#pragma unroll
                for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                    for (int ww = 0; ww < 4; ++ww) {
                        Os[((4 * el + dd) * V2_TW + 4 * j + ww) * OP + gp] = pk[dd][ww];
                        if (SPLIT) Os[((4 * el + dd) * V2_TW + 4 * j + ww) * OP + 8 + gp] = pl[SPLIT ? dd : 0][ww];
                    }
            }
            __syncthreads();
            // ---- write-out: 8 words (32 B) per voxel, consecutive lanes -> consecutive words; a thread keeps its
            // (w, word) and walks the 16 depths of the pass with pointer increments
            if (SPLIT) {
                // 16 words (64 B: [hi 0..15 | lo 0..15]) per voxel
                const int k = threadIdx.x & 15;
                const size_t dstride = (size_t)H * W * Ct_pad;           // words per depth plane (2 halves per channel)
                const int nd = min(V2_DP, D - d0);
                for (int wl = threadIdx.x >> 4; wl < V2_TW; wl += 16) {
                    if (w0 + wl >= W) continue;
                    uint32_t* dst = reinterpret_cast<uint32_t*>(vol + ((((size_t)b * D + d0) * H + h) * W + w0 + wl) * (2 * (size_t)Ct_pad) + 2 * oc0) + k;
                    const uint32_t* src = Os + wl * OP + k;
#pragma unroll 4
                    for (int dl = 0; dl < nd; ++dl) dst[dl * dstride] = src[dl * V2_TW * OP];
                }
            } else {
                const int k = threadIdx.x & 7, wl = threadIdx.x >> 3;
                if (w0 + wl < W) {
                    const size_t dstride = (size_t)H * W * Ct_pad / 2;       // words per depth plane
                    uint32_t* dst = reinterpret_cast<uint32_t*>(vol + ((((size_t)b * D + d0) * H + h) * W + w0 + wl) * Ct_pad + oc0) + k;
                    const uint32_t* src = Os + wl * OP + k;
                    const int nd = min(V2_DP, D - d0);
#pragma unroll 4
                    for (int dl = 0; dl < nd; ++dl) dst[dl * dstride] = src[dl * V2_TW * OP];
                }
            }
        }
    }
}
#undef VROW

// NCDHW fp32 -> NDHWC bf16 and back (layout boundary of the tensor-core path; used by tests and by
// models that enter / leave the path with reference-layout tensors)
__global__ void ncdhw_to_cl_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int C, size_t S,
                                   int Cpad, int f16) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const size_t s0 = (size_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i;
        const size_t sidx = s0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && sidx < S) ? __ldg(src + ((size_t)b * C + c) * S + sidx) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const size_t sidx = s0 + i;
        const int c = c0 + threadIdx.x;
        if (sidx < S && c < Cpad) {
            if (f16 == 2) {                 // operand-split fp16: hi at vsplit_idx(c), lo 16 elements later
                const float x = tile[threadIdx.x][i];
                const __half hv = __float2half_rn(x);
                uint16_t* d = dst + ((size_t)b * S + sidx) * (2 * (size_t)Cpad) + vsplit_idx(c);
                *reinterpret_cast<__half*>(d) = hv;
                *reinterpret_cast<__half*>(d + 16) = __float2half_rn(x - __half2float(hv));
            } else
            dst[((size_t)b * S + sidx) * Cpad + c] = cv16(tile[threadIdx.x][i], f16);
        }
    }
}

// Few input channels (the RGB image padded to 16): one thread per position gathers its C planes (coalesced along S)
// and writes one Cpad-channel row; the 32x32 tile transpose above would move 32 channel rows to use 3 of them.
template <int CPAD>
__global__ void ncdhw_to_cl_small_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int C, size_t S, int f16) {
    const int b = blockIdx.y;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (size_t)gridDim.x * blockDim.x) {
        float v[CPAD];
#pragma unroll
        for (int c = 0; c < CPAD; ++c) v[c] = c < C ? __ldg(src + ((size_t)b * C + c) * S + s) : 0.f;
        if (f16 == 2) {                     // operand-split fp16 (CPAD == 16): [hi 0..15 | lo 0..15]
            uint4* o = reinterpret_cast<uint4*>(dst + ((size_t)b * S + s) * (2 * CPAD));
            float r[CPAD];
#pragma unroll
            for (int c = 0; c < CPAD; ++c) { const __half hv = __float2half_rn(v[c]); r[c] = v[c] - __half2float(hv); }
#pragma unroll
            for (int i = 0; i < CPAD / 8; ++i) {
                o[i] = make_uint4(pk16(v[8 * i], v[8 * i + 1], 1), pk16(v[8 * i + 2], v[8 * i + 3], 1),
                                  pk16(v[8 * i + 4], v[8 * i + 5], 1), pk16(v[8 * i + 6], v[8 * i + 7], 1));
                o[CPAD / 8 + i] = make_uint4(pk16(r[8 * i], r[8 * i + 1], 1), pk16(r[8 * i + 2], r[8 * i + 3], 1),
                                             pk16(r[8 * i + 4], r[8 * i + 5], 1), pk16(r[8 * i + 6], r[8 * i + 7], 1));
            }
            continue;
        }
        uint4* o = reinterpret_cast<uint4*>(dst + ((size_t)b * S + s) * CPAD);
#pragma unroll
        for (int i = 0; i < CPAD / 8; ++i)
            o[i] = make_uint4(pk16(v[8 * i], v[8 * i + 1], f16), pk16(v[8 * i + 2], v[8 * i + 3], f16),
                              pk16(v[8 * i + 4], v[8 * i + 5], f16), pk16(v[8 * i + 6], v[8 * i + 7], f16));
    }
}

__global__ void cl_to_ncdhw_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, int C, size_t S,
                                   int Cpad, int f16) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const size_t s0 = (size_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const size_t sidx = s0 + i;
        const int c = c0 + threadIdx.x;
        float v = 0.f;
        if (sidx < S && c < C) {
            if (f16 == 2) {
                const uint16_t* p = src + ((size_t)b * S + sidx) * (2 * (size_t)Cpad) + vsplit_idx(c);
                v = ld16(p[0], 1) + ld16(p[16], 1);
            } else v = ld16(src[((size_t)b * S + sidx) * Cpad + c], f16);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i;
        const size_t sidx = s0 + threadIdx.x;
        if (c < C && sidx < S) dst[((size_t)b * C + c) * S + sidx] = tile[threadIdx.x][i];
    }
}

}  // namespace

extern "C" int stb_volume_cl16(const float* gwc_l, const float* gwc_r, const float* cat_l, const float* cat_r,
                               void* vol, int f16, int B, int Cg, int G, int Cc, int H, int W, int D, int Ct_pad,
                               int mask_left, void* stream) {
    if (!vol || B <= 0 || H <= 0 || W <= 0 || D <= 0 || G < 0 || Cc < 0) return STB_E_BADARG;
    if (G > 0 && (!gwc_l || !gwc_r || Cg % G)) return STB_E_BADARG;
    if (Cc > 0 && (!cat_l || !cat_r)) return STB_E_BADARG;
    const int Ct = G + 2 * Cc;
    if (Ct <= 0 || Ct_pad < Ct || Ct_pad % 8 || G % 8) return STB_E_UNSUPPORTED;
    const int NR = G > 0 ? (8 * (Cg / G) > 8 ? 8 * (Cg / G) : 8) : 8;
    if (NR > 256) return STB_E_UNSUPPORTED;
    if (H > 65535 || B > 65535) return STB_E_BADARG;
    size_t smem = (size_t)(NR * (VW + 1) + NR * (VW + D)) * sizeof(float);
    if (smem > 200 * 1024) return STB_E_SMEM;
    dim3 grid(stb_ceil_div(W, VW), H, B);
    const int cpg = G > 0 ? Cg / G : 0;
#define STB_VOL_LAUNCH(K)                                                                                          \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            cudaFuncSetAttribute(volume_cl_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        volume_cl_kernel<K><<<grid, VT, smem, (cudaStream_t)stream>>>(gwc_l, gwc_r, cat_l, cat_r, (uint16_t*)vol, Cg, G, \
                                                                   Cc, H, W, D, Ct, Ct_pad, mask_left, NR, f16);   \
    } while (0)
    // (register-blocked CPG variants measured SLOWER on B200 -- 2.43 vs 1.86 ms at B=8 K-shape: occupancy drops
    //  to 2 CTAs/SM and the kernel is bound by staging latency + 16-byte strided stores, not by LDS)
    (void)cpg;
    const bool split = f16 == 2;             // operand-split fp16 volume (f16 = 2): v2 kernel only
    if (Ct_pad % 16 == 0 && W <= 0x7fffffff - 64 && (split || !getenv("STB_VOLUME_V1"))) {
        const int E = (D + 3) / 4, RP = 4 * E + V2_TW;
        const int cpg2 = G > 0 ? Cg / G : 1;
        size_t smem2 = (size_t)16 * cpg2 * (V2_TW + RP) * sizeof(float) + (size_t)V2_DP * V2_TW * (split ? 17 : 9) * sizeof(uint32_t);
        dim3 grid2(stb_ceil_div(W, V2_TW), H, B);
#define STB_VOL2_LAUNCH_S(K, CL, S)                                                                                 \
    do {                                                                                                           \
        cudaFuncSetAttribute(volume_cl2_kernel<K, CL, false, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2); \
        volume_cl2_kernel<K, CL, false, S><<<grid2, 256, smem2, (cudaStream_t)stream>>>(gwc_l, gwc_r, cat_l, cat_r, \
                                                                            (uint16_t*)vol, Cg, G, Cc, H, W, D,   \
                                                                            Ct_pad, mask_left, split ? 1 : f16, ClSrc()); \
        STB_CHECK_LAUNCH();                                                                                        \
        return STB_OK;                                                                                             \
    } while (0)
#define STB_VOL2_LAUNCH(K, CL)                                                                                      \
    do {                                                                                                           \
        if (split) STB_VOL2_LAUNCH_S(K, CL, true);                                                                 \
        STB_VOL2_LAUNCH_S(K, CL, false);                                                                           \
    } while (0)
        if (smem2 <= 200 * 1024) {
            if (cpg2 == 1) STB_VOL2_LAUNCH(1, true);
            if (cpg2 == 4) STB_VOL2_LAUNCH(4, true);
            if (cpg2 == 8) STB_VOL2_LAUNCH(8, true);
            if (cpg2 == 12) STB_VOL2_LAUNCH(12, false);
        }
#undef STB_VOL2_LAUNCH
#undef STB_VOL2_LAUNCH_S
    }
    if (split) return STB_E_UNSUPPORTED;
    STB_VOL_LAUNCH(0);
#undef STB_VOL_LAUNCH
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_volume_cl16_from_cl16(const void* const* feats, const int* feat_ch, int nfeat, const void* cat, int cat_c,
                                         void* vol, int f16, int B, int G, int Cc, int H, int W, int D, int Ct_pad,
                                         int mask_left, void* stream) {
    if (!feats || !feat_ch || nfeat < 1 || nfeat > 4 || !vol || B <= 0 || H <= 0 || W <= 0 || D <= 0 || G <= 0 || Cc < 0)
        return STB_E_BADARG;
    if (Cc > 0 && (!cat || cat_c < Cc)) return STB_E_BADARG;
    ClSrc cs;
    memset(&cs, 0, sizeof(cs));
    int ctot = 0;
    for (int i = 0; i < nfeat; ++i) {
        if (!feats[i] || feat_ch[i] <= 0 || feat_ch[i] % 8) return STB_E_UNSUPPORTED;
        cs.f[i] = (const uint16_t*)feats[i];
        cs.c[i] = feat_ch[i];
        cs.c0[i] = ctot;
        ctot += feat_ch[i];
    }
    cs.n = nfeat;
    cs.cat = (const uint16_t*)cat;
    cs.cat_c = cat_c;
    cs.Bn = B;
    if (ctot != 8 * G) return STB_E_UNSUPPORTED;              // 8 channels per correlation group (GwcNet / ACVNet / PCWNet)
    const int Ct = G + 2 * Cc;
    if (Ct_pad < Ct || Ct_pad % 16 || G % 8) return STB_E_UNSUPPORTED;
    if (H > 65535 || B > 65535) return STB_E_BADARG;
    const int E = (D + 3) / 4, RP = 4 * E + V2_TW;
    const bool split = f16 == 2;             // sources and volume operand-split fp16; feat_ch / cat_c / Ct_pad stay LOGICAL channel counts
    const size_t smem2 = (size_t)16 * 8 * (V2_TW + 4 + RP + 4) * sizeof(float) + (size_t)V2_DP * V2_TW * (split ? 17 : 9) * sizeof(uint32_t);
    if (smem2 > 200 * 1024) return STB_E_SMEM;
    dim3 grid2(stb_ceil_div(W, V2_TW), H, B);
    if (split) {
        for (int i = 0; i < nfeat; ++i)
            if (feat_ch[i] % 16) return STB_E_UNSUPPORTED;
        if (Cc > 0 && cat_c % 16) return STB_E_UNSUPPORTED;
        cs.cat_c = 2 * cat_c;                // storage elements per voxel
        // CACHE_L keeps the left values of a thread's two channels in 64 registers over the depth passes: with the split epilogue's
        // extra live values the kernel then spills 160 bytes per thread at its 128-register cap; STB_VOLUME_CACHEL selects
        // the register-resident variant (measured: 1.73 -> 1.48 ms per batch-8 step without it, profiles/bench_r02_progress.md)
        static const bool cache_l = getenv("STB_VOLUME_CACHEL") ? atoi(getenv("STB_VOLUME_CACHEL")) != 0 : false;
        if (cache_l) {
            cudaFuncSetAttribute(volume_cl2_kernel<8, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            volume_cl2_kernel<8, true, true, true><<<grid2, 256, smem2, (cudaStream_t)stream>>>(nullptr, nullptr, nullptr, nullptr, (uint16_t*)vol,
                                                                                              8 * G, G, Cc, H, W, D, Ct_pad, mask_left, 1, cs);
        } else {
            cudaFuncSetAttribute(volume_cl2_kernel<8, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            volume_cl2_kernel<8, false, true, true><<<grid2, 256, smem2, (cudaStream_t)stream>>>(nullptr, nullptr, nullptr, nullptr, (uint16_t*)vol,
                                                                                               8 * G, G, Cc, H, W, D, Ct_pad, mask_left, 1, cs);
        }
        STB_CHECK_LAUNCH();
        return STB_OK;
    }
    cudaFuncSetAttribute(volume_cl2_kernel<8, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    volume_cl2_kernel<8, true, true, false><<<grid2, 256, smem2, (cudaStream_t)stream>>>(nullptr, nullptr, nullptr, nullptr, (uint16_t*)vol,
                                                                                8 * G, G, Cc, H, W, D, Ct_pad, mask_left, f16, cs);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_ncdhw_to_cl16(const float* src, void* dst, int f16, int B, int C, long long S, int Cpad, void* stream) {
    if (!src || !dst || B <= 0 || C <= 0 || S <= 0 || Cpad < C) return STB_E_BADARG;
    if (f16 == 2 && Cpad % 16) return STB_E_UNSUPPORTED;
    if (C <= 8 && Cpad == 16 && B <= 65535) {
        long long gx = (S + 255) / 256;
        if (gx > 148 * 8) gx = 148 * 8;
        ncdhw_to_cl_small_kernel<16><<<dim3((unsigned)gx, B), 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, C, (size_t)S, f16);
        STB_CHECK_LAUNCH();
        return STB_OK;
    }
    dim3 grid((unsigned)((S + 31) / 32), stb_ceil_div(Cpad, 32), B);
    ncdhw_to_cl_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, C, (size_t)S, Cpad, f16);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_cl16_to_ncdhw(const void* src, float* dst, int f16, int B, int C, long long S, int Cpad, void* stream) {
    if (!src || !dst || B <= 0 || C <= 0 || S <= 0 || Cpad < C) return STB_E_BADARG;
    dim3 grid((unsigned)((S + 31) / 32), stb_ceil_div(C, 32), B);
    cl_to_ncdhw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const uint16_t*)src, dst, C, (size_t)S, Cpad, f16);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
