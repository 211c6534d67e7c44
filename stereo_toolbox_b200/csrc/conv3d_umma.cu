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
// Persistent CTAs (one per SM) stride over the work items.  Warp roles (352 threads): warp 0 = TMA producer, warps 1-2 =
// MMA issuers (one elected lane each, alternate accumulator rounds, issuer i owns TMEM buffer i; warp 1 also owns TMEM
// alloc/dealloc), warps 3..10 = epilogue in two groups (TMEM lane quarter = warp_id % 4).  Pipelines: plane ring with
// per-plane full/empty mbarriers (released by tcgen05.commit right after a plane's last use) and the 2-deep TMEM
// accumulator ring.  What shaped the issue path is measured in profiles/umma_issue_r01.md.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
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
// (32-bit fields throughout: sub-word constant loads have no uniform-datapath form and drag the issue loop's
//  counters into vector registers)
struct UClass { uint32_t tap_begin, tap_end, grp_begin, grp_end; int32_t od0, oh0, ow0, pad; };
// A run of taps of one class that read the same plane (z = dz - dzmin); rel = 1 when no later group of the class
// reads that plane, i.e. the plane may be handed back after this group if no later STEP needs it either.
struct UGroup { uint32_t tap_begin, tap_end, z, rel; };
// Per-tap issue record, read by the issuer straight from the kernel-parameter constant bank into UNIFORM registers
// (LDCU): A start offset inside a plane slot, B start offset inside the weight block (both >> 4), instruction
// descriptor (N = nblk*Cn differs per tap) and accumulator column offset (cls0*Cn).
struct UIss { uint32_t a16, b16, idesc, dcol; };

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
    uint32_t cperm;                  // merged transposed conv: output-parity class (cd*4 + ch*2 + cw) held by column block p = (cperm >> 4p) & 7
    int kdepth;                      // > 0: the K-chunks of the input lie along a PSEUDO-depth axis (plane P = image*kdepth + chunk), 2-D convs
    uint32_t desc_hi;                // high word of every smem descriptor (SBO = 8 rows, version, swizzle mode)
    uint32_t smem_base;              // shared-window address of the 1024-aligned dynamic smem base (queried once per kernel instance)
    int debug;                       // timing experiments only (STB_UMMA_DEBUG): 1 = no TMA plane traffic, 2 = epilogue skips its work
    int ntiles;                      // work items (b, depth chunk, h tile, w tile); CTAs are persistent and stride over them
    int ngroups;
    int nslices;                     // output-channel slices (of Cn) handled inside ONE launch: CTA c takes slice c % nslices
    int cluster;                     // > 1: the nslices CTAs of a tile form a thread-block cluster and share every plane load (TMA multicast)
    float oscale;                    // SPLIT: the packed weights carry a factor 2^s (keeps their lo halves out of the fp16
                                     // subnormals); accumulators are multiplied by oscale = 2^-s before shift / residual
    UClass cls[MAX_UCLASS];
    UTap taps[MAX_UTAPS];
    UGroup grp[MAX_UTAPS];
    UIss iss[MAX_UTAPS];
};

// One work item: an (h-tile, w-tile) column of the volume over a chunk of depth steps.
struct UTile { int b, s_lo, nst, jh0, jw0, p_first, nplanes, nouts; };
__device__ __forceinline__ UTile decode_tile(const UArgs& a, int t) {
    UTile u;
    const int tw_i = t % a.tiles_w; t /= a.tiles_w;
    const int th_i = t % a.tiles_h; t /= a.tiles_h;
    const int ch_i = t % a.nchunks;
    u.b = t / a.nchunks;
    u.s_lo = ch_i * a.dchunk;
    const int s_hi = min(a.nsteps, u.s_lo + a.dchunk);
    u.nst = s_hi - u.s_lo;
    u.jh0 = th_i * a.TH; u.jw0 = tw_i * a.TW;          // class-position origin of this tile
    u.p_first = u.s_lo * a.sd_in + a.dzmin;
    u.nplanes = (s_hi - 1) * a.sd_in + a.dzmax - u.p_first + 1;
    u.nouts = u.nst * a.nclass;                        // accumulator rounds
    return u;
}

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

// 256-bit global accesses (sm_100: LDG.256 / STG.256).  An epilogue thread owns a whole output row (one voxel x 32 channels), so
// the 32 lanes of a warp access 32 different rows: with 128-bit accesses every instruction touched HALF of 32 sectors (ncu on
// the transposed convs: 2.0 store sectors per sector of data, l1tex LSU data pipe at 69-84 % -- the limiter of the memory-heavy
// flavours); with 256-bit accesses an instruction moves whole 32-byte sectors and half as many wavefronts.
__device__ __forceinline__ void stg256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                       uint32_t a6, uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4),
                 "r"(a5), "r"(a6), "r"(a7)
                 : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}

// ---- 32 consecutive logical channels of one voxel <-> storage.  Plain 16-bit: 64 contiguous bytes.  SPLIT (fp16 hi + fp16 lo,
// interleaved per 16 channels): 128 contiguous bytes  [hi 0..15 | lo 0..15 | hi 16..31 | lo 16..31].
template <bool F16, bool SPLIT>
__device__ __forceinline__ void add_residual32(float (&f)[32], const uint16_t* rp) {
    constexpr int f16 = F16 ? 1 : 0;
    if constexpr (SPLIT) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {        // 16-channel block: 32 B of hi halves, then 32 B of lo halves
            uint32_t hw[8], lw[8];
            ldg256(rp + blk * 32, hw);
            ldg256(rp + blk * 32 + 16, lw);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 h2 = unpack16(hw[j], 1), l2 = unpack16(lw[j], 1);
                f[blk * 16 + j * 2] += h2.x + l2.x;          // hi + lo is exact in fp32
                f[blk * 16 + j * 2 + 1] += h2.y + l2.y;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint32_t rw[8];
            ldg256(rp + i * 16, rw);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 h2 = unpack16(rw[j], f16);
                f[i * 16 + j * 2] += h2.x;
                f[i * 16 + j * 2 + 1] += h2.y;
            }
        }
    }
}
template <bool F16, bool SPLIT>
__device__ __forceinline__ void store32(const float (&f)[32], uint16_t* outp) {
    constexpr int f16 = F16 ? 1 : 0;
    if constexpr (SPLIT) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x0 = f[blk * 16 + j * 2], x1 = f[blk * 16 + j * 2 + 1];
                const __half2 h = __floats2half2_rn(x0, x1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                hw[j] = *reinterpret_cast<const uint32_t*>(&h);
                lw[j] = *reinterpret_cast<const uint32_t*>(&l);
            }
            stg256(outp + blk * 32, hw[0], hw[1], hw[2], hw[3], hw[4], hw[5], hw[6], hw[7]);
            stg256(outp + blk * 32 + 16, lw[0], lw[1], lw[2], lw[3], lw[4], lw[5], lw[6], lw[7]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            stg256(outp + i * 16, pack16(f[i * 16 + 0], f[i * 16 + 1], f16), pack16(f[i * 16 + 2], f[i * 16 + 3], f16),
                   pack16(f[i * 16 + 4], f[i * 16 + 5], f16), pack16(f[i * 16 + 6], f[i * 16 + 7], f16),
                   pack16(f[i * 16 + 8], f[i * 16 + 9], f16), pack16(f[i * 16 + 10], f[i * 16 + 11], f16),
                   pack16(f[i * 16 + 12], f[i * 16 + 13], f16), pack16(f[i * 16 + 14], f[i * 16 + 15], f16));
    }
}
// ---- 16 consecutive logical channels (one storage block): 32 bytes plain, 64 bytes split [hi 0..15 | lo 0..15]
template <bool F16, bool SPLIT>
__device__ __forceinline__ void add_residual16(float (&f)[16], const uint16_t* rp) {
    constexpr int f16 = F16 ? 1 : 0;
    uint32_t hw[8];
    ldg256(rp, hw);
    if constexpr (SPLIT) {
        uint32_t lw[8];
        ldg256(rp + 16, lw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 h2 = unpack16(hw[j], 1), l2 = unpack16(lw[j], 1);
            f[j * 2] += h2.x + l2.x;
            f[j * 2 + 1] += h2.y + l2.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 h2 = unpack16(hw[j], f16);
            f[j * 2] += h2.x;
            f[j * 2 + 1] += h2.y;
        }
    }
}
template <bool F16, bool SPLIT>
__device__ __forceinline__ void store16v(const float (&f)[16], uint16_t* outp) {
    constexpr int f16 = F16 ? 1 : 0;
    uint32_t hw[8];
    if constexpr (SPLIT) {
        uint32_t lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x0 = f[j * 2], x1 = f[j * 2 + 1];
            const __half2 h = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            hw[j] = *reinterpret_cast<const uint32_t*>(&h);
            lw[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
        stg256(outp, hw[0], hw[1], hw[2], hw[3], hw[4], hw[5], hw[6], hw[7]);
        stg256(outp + 16, lw[0], lw[1], lw[2], lw[3], lw[4], lw[5], lw[6], lw[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) hw[j] = pack16(f[j * 2], f[j * 2 + 1], f16);
        stg256(outp, hw[0], hw[1], hw[2], hw[3], hw[4], hw[5], hw[6], hw[7]);
    }
}
// storage index of logical channel c inside a split row
__device__ __forceinline__ int split_idx(int c) { return ((c >> 4) << 5) + (c & 15); }

// ---- MMA issue, straight-line.  A branch controlled by a vector register between two UTCHMMAs costs about as much as
// an MMA (probe mmarate3 variant 2: 99 vs 56 clk per N=96 MMA), and every counter of the issuer lives in a vector
// register because it is live across mbarrier spin loops.  So the unit of issue is a chunk of NT taps whose
// NT x NM x KS MMAs are fully unrolled; per-tap records come from the constant bank (LDCU -> uniform registers).
//
// SPLIT (operand-split fp16, "fp16x2"): every activation and weight is stored as fp16 hi + fp16 lo (x = hi + lo, 22
// mantissa bits), interleaved per 16 channels -- a 32-byte k-slice of hi values is followed by the 32-byte k-slice of the
// lo values of the same 16 channels.  One logical K=16 step is three kind::f16 MMAs into the same fp32 accumulator:
// hi*hi, hi*lo_w, lo_a*hi_w (lo*lo is below fp32 rounding).  The tap / plane machinery is unchanged: the lo values are
// just further k-slices of the same TMA-staged rows.
template <int NT, int NM, int KS, bool SPLIT>
__device__ __forceinline__ void issue_taps(const UArgs& a, int tq, uint32_t abase, uint32_t dbase) {
    const uint32_t mt16 = (128u * (uint32_t)a.ROWB) >> 4, dhi = a.desc_hi, ncol = (uint32_t)(a.Cn * a.cblocks);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const UIss e = a.iss[tq + j];
        const uint32_t acc = (e.dcol >> 31) ^ 1u;                  // bit 31: first tap of its class (overwrite)
        const uint32_t dc = dbase + (e.dcol & 0xffffu);
        const uint32_t a0 = abase + e.a16;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const uint32_t dcol_t = dc + (uint32_t)m * ncol;
            const uint32_t alo = a0 + (uint32_t)m * mt16;
            if constexpr (SPLIT) {
#pragma unroll
                for (int k = 0; k < KS; k += 2) {           // k-slice k = hi, k + 1 = lo of the same 16 channels
                    mma_f16_ss2(dcol_t, alo + 2u * k, e.b16 + 2u * k, dhi, e.idesc, k ? 1u : acc);
                    mma_f16_ss2(dcol_t, alo + 2u * k, e.b16 + 2u * k + 2u, dhi, e.idesc, 1u);
                    mma_f16_ss2(dcol_t, alo + 2u * k + 2u, e.b16 + 2u * k, dhi, e.idesc, 1u);
                }
            } else {
#pragma unroll
                for (int k = 0; k < KS; ++k)
                    mma_f16_ss2(dcol_t, alo + 2u * k, e.b16 + 2u * k, dhi, e.idesc, k ? 1u : acc);
            }
        }
    }
}
template <int NT, int KS, bool SPLIT>
__device__ __forceinline__ void issue_dispatch_m(const UArgs& a, int tq, uint32_t abase, uint32_t dbase, int nM) {
    if (nM == 2) issue_taps<NT, 2, KS, SPLIT>(a, tq, abase, dbase);
    else if (nM == 1) issue_taps<NT, 1, KS, SPLIT>(a, tq, abase, dbase);
    else if (nM == 4) issue_taps<NT, 4, KS, SPLIT>(a, tq, abase, dbase);
    else issue_taps<NT, 3, KS, SPLIT>(a, tq, abase, dbase);
}
template <int NT, bool SPLIT>
__device__ __forceinline__ void issue_dispatch(const UArgs& a, int tq, uint32_t abase, uint32_t dbase, int nM, int ks) {
    if (ks == 2) issue_dispatch_m<NT, 2, SPLIT>(a, tq, abase, dbase, nM);
    else if (ks == 4) issue_dispatch_m<NT, 4, SPLIT>(a, tq, abase, dbase, nM);
    else if constexpr (!SPLIT) issue_dispatch_m<NT, 1, SPLIT>(a, tq, abase, dbase, nM);      // split rows hold >= one (hi, lo) slice pair
}

// ACT / F16 are compile-time so the epilogue stays a few hundred instructions: with a runtime activation switch
// unrolled over 32 channels the kernel was 83 KB of SASS and the epilogue warps stalled on instruction fetch
// (ncu: stall_no_inst on every epilogue line, profiles/ncu_umma_r01_*.txt).
// LEAN instantiation: compiled WITHOUT the rarely used paths (ragged channel counts, fp32 output, K-split partial sums,
// timeline trace, timing-experiment switches).  The kernel sits at its 168-register cap with spills, and every line added
// to the epilogue was measured to cost ALL layers (profiles/umma_issue_r01.md section 4), so the common layers get a
// kernel that does not carry code they never execute.
// LEAN: 0 = generic, 1 = lean / plain taps (one column block), 2 = lean / kw-merged (3 column blocks + lane realignment),
// 4 = kw-merged single-channel classifier with fp32 output (opt-in, STB_UMMA_CLS1),
// 5 = merged transposed conv with the two w-parity classes stored as one contiguous pair (opt-in, STB_UMMA_T2PAIR),
// 3 = lean / merged transposed conv (8 parity-class blocks).
template <int ACT, bool F16, int LEAN, bool SPLIT>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv3d_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ UArgs a) {
    // LV: the epilogue flavour; LEAN 7 = the generic flavour (0) plus the stride-2 pair merge, kept out of LEAN 0 so that the generic
    // instantiation (K-split passes of the big layers) does not carry it
    // LEAN 9 = flavour 6 (merged transposed conv on 16-channel slices) as a K-split pass: adds the fp32 partial of the earlier
    // passes and / or writes its own (grouped pseudo-depth chunks, flags bits 11..13 of the host entry)
    constexpr int LV = LEAN == 7 ? 0 : (LEAN == 9 ? 6 : LEAN);
    constexpr bool PIO = LEAN == 9;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    // [0,2048): barriers + tmem holder + issue tables ; then weights ; then plane ring
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);
    uint64_t* plane_full = bar_w + 1;
    uint64_t* plane_empty = plane_full + MAX_RING;
    uint64_t* tmem_full = plane_empty + MAX_RING;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint8_t* sW = smem + 2048;
    uint8_t* sP = sW + ((a.w_bytes_total + 1023) & ~1023u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Output-channel slices inside one launch (weights of Cpad / nslices channels fit in shared memory, not all of them):
    // CTA c owns slice c % nslices and walks every work item with stride gridDim.x / nslices, so the CTAs of all slices work
    // on the same tile at about the same time and the input planes are read from HBM once (L2 serves the others).  Before:
    // one launch per slice (profiles/op_bench_r01.md: 87 launches for 30 layers).
    const int nsl = a.nslices, slice = (int)blockIdx.x % nsl;
    const int tile0 = (int)blockIdx.x / nsl, tstride = (int)gridDim.x / nsl;
    const int cout_off = a.cout_off + slice * a.Cn;
    if (a.ntiles < 0) {                            // base-address query launch (host: smem_base_of)
        if (threadIdx.x == 0) *reinterpret_cast<uint32_t*>(a.out) = smem_u32(smem);
        return;
    }

    // ---- persistent CTA: work items blockIdx.x, blockIdx.x + gridDim.x, ... ; every role walks the same list and
    // carries its pipeline state (ring slot / phase, accumulator round) across items, so the producer is already
    // filling the ring for the next item while the tensor pipe and the epilogue finish the current one.
    long long* const trace_buf = LV ? nullptr : g_umma_trace;   // read ONCE: a global load per round sat on the issuer's critical path
    const int trace_rounds = (!LV && trace_buf && blockIdx.x == 0) ? g_umma_trace_rounds : 0;
    const int dbg = LV ? 0 : a.debug;

    // Cluster mode: the CTAs of a cluster (= the output-channel slices of one tile sequence) walk identical plane sequences; every
    // plane box is issued by ONE of them, round robin, and multicast into the ring slot of all of them.  A slot may be refilled
    // once the issuers of ALL CTAs have released it: plane_empty counts 2 * ncl arrivals, delivered by multicast commits.
    if (threadIdx.x == 0) {
        const int ncl = a.cluster > 1 ? a.cluster : 1;
        mbar_init(bar_w, 1);
        // active epilogue groups: two when a round holds >= 2 (M-tile, class block) items, or (column split) when its single
        // item is >= 64 channels wide -- the groups then take alternate 32-column blocks
#ifdef STB_NO_CSPLIT
        const bool csplit_i = false;
#else
        const bool csplit_i = ((LV == 1 || LV == 2) && a.nM == 1 && a.cblocks != 8 && a.Cn >= 64) || LV == 8;
#endif
        const int n_grp = (a.nM * (a.cblocks == 8 ? 8 : 1) >= 2 || csplit_i) ? 2 : 1;
        for (int i = 0; i < a.R; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], 2 * ncl); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4 * n_grp); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, a.tmem_cols);
    tc_fence_before();
    if (a.cluster > 1) cluster_sync_all();    // no CTA may multicast into a peer whose barriers are not initialised yet
    else __syncthreads();
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
                            (i * a.w_tile_stride + a.w_kc_off) * a.w_rows + cout_off);
            int slot = 0;
            uint32_t eph = 1;                      // parity to wait for on plane_empty[slot] (fresh barrier: passes)
            uint32_t gbox = 0;                     // running box counter, identical in every CTA of a cluster: box j is issued by CTA j % ncl
            const int ncl = a.cluster > 1 ? a.cluster : 1;
            const uint32_t crank = ncl > 1 ? cluster_ctarank() : 0u;
            const uint16_t cmask = (uint16_t)((1u << ncl) - 1u);
            for (int tile = tile0; tile < a.ntiles && !(dbg & 1); tile += tstride) {
                const UTile u = decode_tile(a, tile);
                const int ih0 = (u.jh0 + a.in_h_off) * a.in_stride, iw0 = (u.jw0 + a.in_w_off) * a.in_stride;
                for (int n = 0; n < u.nplanes; ++n) {
                    mbar_wait_backoff(&plane_empty[slot], eph, 64);
                    mbar_arrive_expect_tx(&plane_full[slot], (uint32_t)a.nsub * a.chunk_bytes);
                    uint8_t* dst = sP + (size_t)slot * a.plane_bytes;
                    if (a.kdepth > 0) {
                        // 2-D conv with Cin = kdepth K-chunks: pseudo-plane P = image*kdepth + chunk; the taps of chunk c
                        // carry dz = c, so all chunks accumulate in TMEM inside one launch (no fp32 workspace round trip)
                        const int P = u.p_first + n, img = (P + 64 * a.kdepth) / a.kdepth - 64, chunk = P - img * a.kdepth;   // floor division: 3-D planes may be negative (padding)
                        if (ncl == 1) tma_load_5d(dst, &tm_x, &plane_full[slot], a.cin_off + chunk * (a.ROWB >> 1), iw0, ih0, img, u.b);
                        else if (gbox % (uint32_t)ncl == crank)
                            tma_load_5d_mc(dst, &tm_x, &plane_full[slot], a.cin_off + chunk * (a.ROWB >> 1), iw0, ih0, img, u.b, cmask);
                        ++gbox;
                    } else {
                        for (int sb = 0; sb < a.nsub; ++sb, ++gbox) {     // sub-tile sb = (h parity, w parity) for stride 2
                            if (ncl == 1)
                                tma_load_5d(dst + (size_t)sb * a.chunk_bytes, &tm_x, &plane_full[slot], a.cin_off,
                                            iw0 + (sb & 1), ih0 + (sb >> 1), u.p_first + n, u.b);
                            else if (gbox % (uint32_t)ncl == crank)
                                tma_load_5d_mc(dst + (size_t)sb * a.chunk_bytes, &tm_x, &plane_full[slot], a.cin_off,
                                               iw0 + (sb & 1), ih0 + (sb >> 1), u.p_first + n, u.b, cmask);
                        }
                    }
                    if (++slot == a.R) { slot = 0; eph ^= 1; }
                }
            }
        }
    } else if (warp <= 2) {
        // ================================ MMA issuers (2) ================================
        // One elected thread per warp; the two issuers take ALTERNATE accumulator rounds (issuer i owns TMEM buffer i).
        // What bounds the tensor pipe here is not the UTCHMMA stream itself but the scalar bookkeeping between the
        // streams (waits, ring arithmetic, record loads: ~600 clk per group of latency-bound single-thread code, while
        // the MMA queue is only a few entries deep -- in-kernel timeline, profiles/umma_issue_r01.md): with two issuers
        // one does its bookkeeping while the other's MMAs run.  The MMAs of a group are issued as straight-line blocks
        // (issue_taps).  Taps are sorted by the plane they read, so an issuer consumes planes in increasing order: it
        // waits for a plane right before its first tap and hands it back (tcgen05.commit -> plane_empty, count 2: a slot
        // is recycled once BOTH issuers released it) right after its last use, i.e. when its next round no longer needs it.
        const int issuer = warp - 1;
        if (elect_one()) {
            const uint32_t plane16 = a.plane_bytes >> 4;
            const int ksteps = a.ROWB / 32;
            // The per-tap records hold ABSOLUTE encoded addresses (the host knows this kernel's dynamic-smem base, see
            // launch_one): everything the issue blocks need is an LDCU away, never an R2UR.
            if ((smem_u32(smem) & 0xFFFFFFu) != a.smem_base) asm volatile("trap;");      // low 24 bits: CTA-local offset (cluster rank above)
            const int nM = a.nM, R = a.R, sd = a.sd_in, nclass = a.nclass;
            const int ncl = a.cluster > 1 ? a.cluster : 1;
            const uint16_t cmask = (uint16_t)((1u << ncl) - 1u);
            const uint32_t ncol = (uint32_t)(a.Cn * a.cblocks);
            mbar_wait(bar_w, 0);
            int waited = 0, released = 0;         // planes of the CURRENT item known resident / handed back (this issuer)
            int wslot = 0, rslot = 0;             // ring slots of plane `waited` / `released`
            uint32_t wphase = 0;
            int base_slot = 0;                    // ring slot of the current item's plane 0
            uint32_t ground = 0;                  // accumulator round counter over all items
            const bool no_planes = dbg & 1;
            auto wait_upto = [&](int n) {
                while (waited < n) {
                    if (!no_planes) mbar_wait(&plane_full[wslot], wphase);
                    ++waited;
                    if (++wslot == R) { wslot = 0; wphase ^= 1; }
                }
            };
            auto release_upto = [&](int n) {      // each arrival fires when every MMA this thread issued so far has completed
                while (released < n) {
                    wait_upto(released + 1);      // never hand back a plane that has not landed (phase safety)
                    if (ncl > 1) mma_commit_mc(&plane_empty[rslot], cmask);
                    else mma_commit(&plane_empty[rslot]);
                    ++released;
                    if (++rslot == R) rslot = 0;
                }
            };
            for (int tile = tile0; tile < a.ntiles; tile += tstride) {
                const UTile u = decode_tile(a, tile);
                int step_slot = base_slot;        // ring slot of plane (s*sd + dzmin) for the current step
                for (int si = 0; si < u.nst; ++si) {
                    const int dead_now = si * sd;                 // item-relative index of the step's first plane
                    for (int c = 0; c < nclass; ++c, ++ground) {
                        if ((int)(ground & 1u) != issuer) continue;
                        const bool trace = !LV && (int)ground < trace_rounds;
                        if (trace) trace_buf[ground * 8 + 0] = clock64();
                        // first plane this issuer still needs in ITS next round (ground + 2)
                        int nsi = si, nc = c + 2;
                        while (nc >= nclass) { nc -= nclass; ++nsi; }
                        const int next_dead = nsi >= u.nst ? u.nplanes : nsi * sd;
                        release_upto(dead_now);                   // planes between my previous window and this one
                        const int buf = issuer;
                        if (!(dbg & 4)) {
                            mbar_wait(&tmem_empty[buf], ((ground >> 1) & 1) ^ 1);
                            tc_fence_after();
                        }
                        if (trace) trace_buf[ground * 8 + 1] = clock64();
                        const uint32_t dbase = tmem_base + (uint32_t)(buf * nM) * ncol;
                        const int gb = (int)a.cls[c].grp_begin, ge = (int)a.cls[c].grp_end;
                        for (int g = gb; g < ge; ++g) {
                            const UGroup gr = a.grp[g];
                            const int z = (int)gr.z, g0 = (int)gr.tap_begin, g1 = (int)gr.tap_end;
                            int sl = step_slot + z;
                            while (sl >= R) sl -= R;          // (kdepth rings may be shorter than a round's chunk count)
                            const uint32_t abase = (uint32_t)sl * plane16;
                            wait_upto(dead_now + z + 1);
                            if (trace && g == gb) trace_buf[ground * 8 + 2] = clock64();
                            int tq = g0;
                            if constexpr (SPLIT) {          // 3x the MMAs per tap: unroll one tap at a time
                                for (; tq < g1; ++tq) issue_dispatch<1, true>(a, tq, abase, dbase, nM, ksteps);
                            } else {
                                for (; tq + 3 <= g1; tq += 3) issue_dispatch<3, false>(a, tq, abase, dbase, nM, ksteps);
                                for (; tq < g1; ++tq) issue_dispatch<1, false>(a, tq, abase, dbase, nM, ksteps);
                            }
                            if (gr.rel && dead_now + z < next_dead) release_upto(dead_now + z + 1);
                        }
                        if (!(dbg & 4)) mma_commit(&tmem_full[buf]);
                        if (trace) trace_buf[ground * 8 + 3] = clock64();
                    }
                    step_slot += sd;
                    while (step_slot >= R) step_slot -= R;
                }
                release_upto(u.nplanes);          // everything of this item (also planes only the other issuer read)
                base_slot = (base_slot + u.nplanes) % R;
                waited -= u.nplanes;              // item-relative counters restart; wslot / rslot / wphase carry on
                released -= u.nplanes;
            }
        }
    } else if (warp >= 3) {
        // ================================ epilogue warps ================================
        const int q4 = warp & 3;                    // TMEM lane quarter this warp may touch
        const int egroup = (warp - 3) >> 2;         // epilogue group 0/1 drains M-tiles m = egroup, egroup+2, ...
        const int nblk_e = LV == 3 ? 8 : (LV == 5 || LV == 6) ? 4 : (LV ? 1 : (a.cblocks == 8 ? 8 : 1));     // merged transposed conv: all 8 parity classes in one round (LV 5: as 4 w-pairs)
        const int items = a.nM * nblk_e;               // (M-tile, class block) work items per round, dealt to 2 groups
        // Column split: with one M-tile per round (the split-storage tiling) the second epilogue group used to idle; for items
        // >= 64 channels wide the two groups take alternate 32-column blocks (2-D 64-channel layers: the epilogue, not the MMAs,
        // bounded them -- 8.1 k clk per item for 3.5 k clk of MMAs)
#ifdef STB_NO_CSPLIT
        constexpr bool csplit = false;
        const int item0 = egroup;
        constexpr int item_step = 2, cblk0 = 0, cblk_step = 32;
#else
        const bool csplit = (LV == 1 || LV == 2) && items == 1 && a.Cn >= 64;
        const int item0 = (csplit || LV == 8) ? 0 : egroup, item_step = csplit ? 1 : 2;
        const int cblk0 = csplit ? 32 * egroup : 0, cblk_step = csplit ? 64 : 32;
#endif
        const bool active = egroup < ((items >= 2 || csplit || LV == 8) ? 2 : 1);
        const size_t ostride_w = (size_t)a.Cout_total;
        constexpr int f16 = F16 ? 1 : 0;
        constexpr size_t K16 = SPLIT ? 2 : 1;       // 16-bit storage elements per logical channel
        const bool full32 = LV || (a.Cn_valid & 31) == 0;  // every 32-column block is complete: vector path
        const float* const partial = (LV && !PIO) ? nullptr : a.partial;
        const bool out_fp32 = (!LV || PIO) && a.out_fp32;
        float sh0[32];                              // folded-BN shift of the first 32 channels stays in registers
#pragma unroll
        for (int i = 0; i < 32; ++i) sh0[i] = (a.shift && i < a.Cn_valid) ? __ldg(a.shift + cout_off + i) : 0.f;
        float sh16[16];                             // LEAN 8: this group's 16 channels
#pragma unroll
        for (int i = 0; i < 16; ++i) sh16[i] = (LV == 8 && a.shift) ? __ldg(a.shift + cout_off + 16 * egroup + i) : 0.f;
        const float oscale = SPLIT ? a.oscale : 1.f;
        const int merge = (LV == 2 || LV == 4 || LV == 8) ? 3 : (LV ? 1 : a.merge), Cn = a.Cn, nM = a.nM;
        uint32_t ground = 0;                        // accumulator round counter over all items (same order as the issuer)
        // The memory-heavy flavours (transposed convs: 8 output blocks per M-tile, each with a residual row to read; K-split
        // passes reading their fp32 partial sums) were bound by the LATENCY of those loads: a block loads, waits, computes,
        // stores, and only then the next block's loads are issued (fp16 64->32 s2T: 0.66 ms for an MMA time of 0.14 ms and a
        // HBM floor of 0.26 ms).  Registers for a software pipeline do not exist (168-register cap), so the rows a round will
        // read are pulled into L2 one round ahead with prefetch.global.L2 -- no registers, no completion to wait for.
        constexpr bool PREFETCH = LV == 0 || LV == 1 || LV == 3 || LV == 5 || LV == 6;
        const bool want_pf = PREFETCH && (a.residual != nullptr || partial != nullptr);
        auto prefetch_round = [&](const UTile& u, int s, const UClass& cl) {
            for (int item = egroup; item < items; item += 2) {
                int m, od, oh, ow, jh_l, jw_l;
                bool ok;
                if (LV == 5 || LV == 6) {
                    m = item >> 2;
                    const int pb = item & 3, q = 128 * m + q4 * 32 + lane;
                    jh_l = q / TWP; jw_l = q % TWP;
                    const int pc = (int)((a.cperm >> (8 * pb)) & 7u);          // class of the pair's first column block
                    od = s * a.out_stride + (pc >> 2); oh = (u.jh0 + jh_l) * 2 + ((pc >> 1) & 1); ow = (u.jw0 + jw_l) * 2;
                } else {
                    m = nblk_e == 8 ? (item >> 3) : item;
                    const int blk = nblk_e == 8 ? (int)((a.cperm >> (4 * (item & 7))) & 7u) : 0, q = 128 * m + q4 * 32 + lane;
                    jh_l = q / TWP; jw_l = q % TWP;
                    od = s * a.out_stride + (nblk_e == 8 ? (blk >> 2) : cl.od0);
                    oh = (u.jh0 + jh_l) * a.out_stride + (nblk_e == 8 ? ((blk >> 1) & 1) : cl.oh0);
                    ow = (u.jw0 + jw_l) * a.out_stride + (nblk_e == 8 ? (blk & 1) : cl.ow0);
                }
                ok = jh_l < a.TH && jw_l < a.TW && (u.jh0 + jh_l) < a.nclass_h && (u.jw0 + jw_l) < a.nclass_w && od < a.Do &&
                     oh < a.Ho && ow < a.Wo;
                if (!ok) continue;
                const size_t vox = (((size_t)u.b * a.Do + od) * a.Ho + oh) * a.Wo + ow;
                for (int c0 = 0; c0 < Cn; c0 += 32) {
                    const size_t eoff = vox * ostride_w + cout_off + c0;
                    if (a.residual) {
                        const uint16_t* rp = reinterpret_cast<const uint16_t*>(a.residual) + eoff * K16;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                        if (LV == 5 && SPLIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 64));   // the pair's second voxel
                        if (LV == 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + ostride_w * K16));
                    }
                    if (partial) asm volatile("prefetch.global.L2 [%0];" ::"l"(partial + eoff));
                    if (PIO && partial) asm volatile("prefetch.global.L2 [%0];" ::"l"(partial + eoff + ostride_w));   // the pair's second voxel
                }
            }
        };
        for (int tile = tile0; active && tile < a.ntiles && !(dbg & 4); tile += tstride) {
          const UTile u = decode_tile(a, tile);
          const int b = u.b, jh0 = u.jh0, jw0 = u.jw0;
          if (want_pf) prefetch_round(u, u.s_lo, a.cls[0]);
          int e_si = 0, e_c = 0;                      // (step, class) of the round, kept incrementally (no division)
          for (int round = 0; round < u.nouts; ++round, ++ground) {
            const int buf = ground & 1;
            const int s = u.s_lo + e_si;
            const UClass cl = a.cls[e_c];
            if (++e_c == a.nclass) { e_c = 0; ++e_si; }
            if (want_pf && round + 1 < u.nouts) prefetch_round(u, u.s_lo + e_si, a.cls[e_c]);   // (e_si, e_c) already name the NEXT round
            mbar_wait_warp(&tmem_full[buf], (ground >> 1) & 1);
            tc_fence_after();
            const bool trace = !LV && (int)ground < trace_rounds && warp == 3 && lane == 0;
            if (trace) trace_buf[ground * 8 + 4] = clock64();
            for (int item = item0; item < items && !(dbg & 2); item += item_step) {
              if (LV == 6) {
                // Merged transposed conv on a 16-channel output slice (Cn == 16: all K-chunks of the layer accumulate in TMEM because
                // the weights of a 16-channel slice fit next to the chunk ring, no K-split pass through an fp32 partial).  The two
                // w-parity classes of one (d,h) parity are ADJACENT 16-column blocks, so one 32-column TMEM read holds the thread's
                // two output voxels (ow, ow + 1) x 16 channels; each is one whole 16-channel storage block of its voxel row.
                // (class order of the column blocks = a.cperm: the two blocks of a pair always share (cd, ch); with the Gray order
                //  -- fewer, wider MMAs, see the host plan -- every other pair holds cw = 1 FIRST: `rev`)
                const int m = item >> 2, pb = item & 3;
                const int pc = (int)((a.cperm >> (8 * pb)) & 7u);
                const bool rev = pc & 1;
                const int od = s * a.out_stride + (pc >> 2);
                const int q = 128 * m + q4 * 32 + lane;
                const int jh_l = q / TWP, jw_l = q % TWP;
                const int jh = jh0 + jh_l, jw = jw0 + jw_l;
                const bool valid = (jh_l < a.TH) && (jw_l < a.TW) && (jh < a.nclass_h) && (jw < a.nclass_w) && (od < a.Do);
                const int oh = jh * 2 + ((pc >> 1) & 1), ow = jw * 2;
                const bool in0 = valid && oh < a.Ho && ow < a.Wo, in1 = in0 && ow + 1 < a.Wo;
                const size_t eoff = ((((size_t)b * a.Do + od) * a.Ho + oh) * a.Wo + ow) * ostride_w + cout_off;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((buf * nM + m) * Cn * 8 + pb * 2 * Cn);
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(taddr, v);
                tmem_ld_wait();
                if (item + 2 >= items) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                if (rev) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { const uint32_t t = v[i]; v[i] = v[16 + i]; v[16 + i] = t; }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!(h ? in1 : in0)) continue;
                    const size_t off = eoff + (size_t)h * ostride_w;
                    float f[16];
                    if constexpr (PIO) {
                        // K-split pass: (+ the fp32 partial of the earlier passes) and, unless it is the last pass, the raw sums back
                        // out as the next pass's partial -- 16 floats = two whole sectors per voxel
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[h * 16 + i]);
                        if (partial) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                uint32_t pv[8];
                                ldg256(partial + off + i * 8, pv);
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[i * 8 + j] += __uint_as_float(pv[j]);
                            }
                        }
                        if (out_fp32) {
                            float* op = reinterpret_cast<float*>(a.out) + off;
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                stg256(op + i * 8, __float_as_uint(f[i * 8]), __float_as_uint(f[i * 8 + 1]), __float_as_uint(f[i * 8 + 2]),
                                       __float_as_uint(f[i * 8 + 3]), __float_as_uint(f[i * 8 + 4]), __float_as_uint(f[i * 8 + 5]),
                                       __float_as_uint(f[i * 8 + 6]), __float_as_uint(f[i * 8 + 7]));
                            continue;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = SPLIT ? fmaf(f[i], oscale, sh0[i]) : f[i] + sh0[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = SPLIT ? fmaf(__uint_as_float(v[h * 16 + i]), oscale, sh0[i]) : __uint_as_float(v[h * 16 + i]) + sh0[i];
                    }
                    if (a.residual) add_residual16<F16, SPLIT>(f, reinterpret_cast<const uint16_t*>(a.residual) + off * K16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = stb_act(f[i], ACT);
                    store16v<F16, SPLIT>(f, reinterpret_cast<uint16_t*>(a.out) + off * K16);
                }
                continue;
              }
              if (LV == 8) {
                // kw-merged conv, ONE 32-channel M-tile per round (the 2-D layers on split storage: extractor 128-channel layers in
                // 32-channel slices, the update block's convs): the round's single item is split by COLUMNS -- group g takes
                // channels [16g, 16g + 16), one whole (hi, lo) storage block -- instead of leaving the second group idle; these
                // layers have a third of the taps per output of a 3-D conv, so the epilogue bounds them.
                const int q = q4 * 32 + lane;
                const int jh_l = q / TWP, jw_l = q % TWP;
                const int jh = jh0 + jh_l, jw = jw0 + jw_l;
                const int od = s + cl.od0;
                const bool inb = (jh_l < a.TH) && (jw_l < a.TW) && (jh < a.nclass_h) && (jw < a.nclass_w) && (od < a.Do) && jh < a.Ho && jw < a.Wo;
                const size_t eoff = ((((size_t)b * a.Do + od) * a.Ho + jh) * a.Wo + jw) * ostride_w + cout_off + 16 * egroup;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * Cn * 3 + 16 * egroup);
                uint32_t v0[16], v1[16], v2[16];
                __syncwarp();
                tmem_ld_32x32_x16(taddr, v0);
                tmem_ld_32x32_x16(taddr + (uint32_t)Cn, v1);
                tmem_ld_32x32_x16(taddr + (uint32_t)(2 * Cn), v2);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                const int ms = a.merge_step;
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float r = __uint_as_float(v0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v1[i]), ms) +
                                    __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i]), 2 * ms);
                    f[i] = SPLIT ? fmaf(r, oscale, sh16[i]) : r + sh16[i];
                }
                if (inb) {
                    if (a.residual) add_residual16<F16, SPLIT>(f, reinterpret_cast<const uint16_t*>(a.residual) + eoff * K16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = stb_act(f[i], ACT);
                    store16v<F16, SPLIT>(f, reinterpret_cast<uint16_t*>(a.out) + eoff * K16);
                }
                break;                                  // the round's only item
              }
              if (LV == 5) {
                // Merged transposed conv, the two w-parity classes of one (d,h) parity handled together: a thread's two output
                // voxels (ow, ow + 1) are adjacent in memory, so it writes 2 x Cout_total contiguous channels (128 B at 32
                // channels) and a warp a fully contiguous span, instead of 64-byte pieces at a 128-byte stride; the address /
                // validity arithmetic is done once per pair.  Cn == 32 (one column block per class).
                const int m = item >> 2, pb = item & 3;
                const int od = s * a.out_stride + (pb >> 1);
                const int q = 128 * m + q4 * 32 + lane;
                const int jh_l = q / TWP, jw_l = q % TWP;
                const int jh = jh0 + jh_l, jw = jw0 + jw_l;
                const bool valid = (jh_l < a.TH) && (jw_l < a.TW) && (jh < a.nclass_h) && (jw < a.nclass_w) && (od < a.Do);
                const int oh = jh * 2 + (pb & 1), ow = jw * 2;
                const bool in0 = valid && oh < a.Ho && ow < a.Wo, in1 = in0 && ow + 1 < a.Wo;
                const size_t eoff = ((((size_t)b * a.Do + od) * a.Ho + oh) * a.Wo + ow) * ostride_w + cout_off;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((buf * nM + m) * Cn * 8 + pb * 2 * Cn);
                uint32_t v0[32], v1[32];
                __syncwarp();
                tmem_ld_32x32(taddr, v0);
                tmem_ld_32x32(taddr + (uint32_t)Cn, v1);
                tmem_ld_wait();
                if (item + 2 >= items) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                auto finish = [&](const uint32_t (&v)[32], size_t off, bool ok) {
                    if (!ok) return;
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = SPLIT ? fmaf(__uint_as_float(v[i]), oscale, sh0[i]) : __uint_as_float(v[i]) + sh0[i];
                    if (a.residual) add_residual32<F16, SPLIT>(f, reinterpret_cast<const uint16_t*>(a.residual) + off * K16);
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = stb_act(f[i], ACT);
                    store32<F16, SPLIT>(f, reinterpret_cast<uint16_t*>(a.out) + off * K16);
                };
                finish(v0, eoff, in0);
                finish(v1, eoff + ostride_w, in1);
                continue;
              }
              {
                const int m = nblk_e == 8 ? (item >> 3) : item, blk = nblk_e == 8 ? (item & 7) : 0;      // blk: column-block position
                const int pcls = (int)((a.cperm >> (4 * blk)) & 7u);                                       // ... and the class it holds
                const int cd = nblk_e == 8 ? (pcls >> 2) : cl.od0, chh = nblk_e == 8 ? ((pcls >> 1) & 1) : cl.oh0,
                          cww = nblk_e == 8 ? (pcls & 1) : cl.ow0;
                const int od = s * a.out_stride + cd;
                const int q = 128 * m + q4 * 32 + lane;
                const int jh_l = q / TWP, jw_l = q % TWP;
                const int jh = jh0 + jh_l, jw = jw0 + jw_l;
                const bool valid = (jh_l < a.TH) && (jw_l < a.TW) && (jh < a.nclass_h) && (jw < a.nclass_w) && (od < a.Do);
                const int oh = jh * a.out_stride + chh, ow = jw * a.out_stride + cww;
                const bool inb = valid && oh < a.Ho && ow < a.Wo;
                const size_t vox = (((size_t)b * a.Do + od) * a.Ho + oh) * a.Wo + ow;
                for (int c0 = cblk0; c0 < Cn; c0 += cblk_step) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) +
                                           (uint32_t)((buf * nM + m) * Cn * a.cblocks + (nblk_e == 8 ? blk * Cn : 0) + c0);
                    // All TMEM reads of the block are issued back to back and waited for once; after the LAST block of
                    // the round the accumulator buffer is handed back to its issuer BEFORE the arithmetic and the
                    // stores (the data is in registers): the buffer is busy ~0.3K instead of ~2.5K clocks per round,
                    // which was what the issuers were waiting for (profiles/umma_issue_r01.md).
                    if (LV == 4) {
                        // 32->1 classifier (kw-merged, fp32 out, no residual): ONE accumulator column per kw block is
                        // live, so read three single columns instead of three 32-column blocks and realign with two
                        // shuffles instead of 64.  Same additions in the same order as the generic path below.
                        uint32_t v0, v1, v2;
                        __syncwarp();
                        tmem_ld_32x32_x1(taddr, v0);
                        tmem_ld_32x32_x1(taddr + (uint32_t)Cn, v1);
                        tmem_ld_32x32_x1(taddr + (uint32_t)(2 * Cn), v2);
                        tmem_ld_wait();
                        if (item + 2 >= items) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                        }
                        const int ms1 = a.merge_step;
                        const float r = __uint_as_float(v0) + __shfl_down_sync(0xffffffffu, __uint_as_float(v1), ms1) +
                                        __shfl_down_sync(0xffffffffu, __uint_as_float(v2), 2 * ms1);
                        if (inb) reinterpret_cast<float*>(a.out)[vox * ostride_w + cout_off] = stb_act(SPLIT ? fmaf(r, oscale, sh0[0]) : r + sh0[0], ACT);
                        break;                         // Cn = 16: a single column block per item
                    }
                    float f[32];
                    __syncwarp();                      // tcgen05.ld is .sync.aligned: whole warp, converged
                    if (merge == 3) {
                        // kw-merged accumulators: column block k holds P_k[q] = sum over (kd,kh,ci) for filter column
                        // kw = k evaluated WITHOUT the w shift; out[q] = P_0[q] + P_1[q+1] + P_2[q+2], and q+k is
                        // lane+k of the same warp (one warp = one padded tile row).
                        uint32_t v0[32], v1[32], v2[32];
                        tmem_ld_32x32(taddr, v0);
                        tmem_ld_32x32(taddr + (uint32_t)Cn, v1);
                        tmem_ld_32x32(taddr + (uint32_t)(2 * Cn), v2);
                        tmem_ld_wait();
                        if (trace && item == egroup && c0 == 0) trace_buf[ground * 8 + 6] = clock64();     // TMEM data in registers
                        if (item + item_step >= items && c0 + cblk_step >= Cn) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                        }
                        const int ms = a.merge_step;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            f[i] = __uint_as_float(v0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v1[i]), ms) +
                                   __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i]), 2 * ms);
                    } else if (LEAN == 7 && merge == 2) {
                        // stride-2 pair merge: block 0 = kw 0 (+ the un-merged kw 1 taps), block 1 = kw 2 evaluated one padded
                        // position early: out[q] = P_0[q] + P_1[q+1]
                        uint32_t v0[32], v1[32];
                        tmem_ld_32x32(taddr, v0);
                        tmem_ld_32x32(taddr + (uint32_t)Cn, v1);
                        tmem_ld_wait();
                        if (item + item_step >= items && c0 + cblk_step >= Cn) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            f[i] = __uint_as_float(v0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v1[i]), 1);
                    } else {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr, v);
                        tmem_ld_wait();
                        if (item + item_step >= items && c0 + cblk_step >= Cn) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    }
                    const size_t eoff = vox * ostride_w + cout_off + c0;
                    if (inb && full32) {
                        // ---------------- vector path: 32 complete channels
                        if (partial) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint32_t pv[8];
                                ldg256(partial + eoff + i * 8, pv);
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[i * 8 + j] += __uint_as_float(pv[j]);
                            }
                        }
                        if constexpr (SPLIT) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] *= oscale;
                        }
                        if (c0 == 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] += sh0[i];
                        } else if (a.shift) {
                            const float4* sp = reinterpret_cast<const float4*>(a.shift + cout_off + c0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 sv = __ldg(sp + i);
                                f[i * 4] += sv.x; f[i * 4 + 1] += sv.y; f[i * 4 + 2] += sv.z; f[i * 4 + 3] += sv.w;
                            }
                        }
                        if (a.residual) add_residual32<F16, SPLIT>(f, reinterpret_cast<const uint16_t*>(a.residual) + eoff * K16);
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = stb_act(f[i], ACT);
                        if (trace && item == egroup && c0 == 0) trace_buf[ground * 8 + 7] = clock64();     // arithmetic done, stores next
                        if (out_fp32) {
                            float* op = reinterpret_cast<float*>(a.out) + eoff;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                stg256(op + i * 8, __float_as_uint(f[i * 8]), __float_as_uint(f[i * 8 + 1]), __float_as_uint(f[i * 8 + 2]),
                                       __float_as_uint(f[i * 8 + 3]), __float_as_uint(f[i * 8 + 4]), __float_as_uint(f[i * 8 + 5]),
                                       __float_as_uint(f[i * 8 + 6]), __float_as_uint(f[i * 8 + 7]));
                        } else {
                            store32<F16, SPLIT>(f, reinterpret_cast<uint16_t*>(a.out) + eoff * K16);
                        }
                    } else if (inb) {
                        // ---------------- ragged path (Cout not a multiple of 32: the 32->1 classifier, IGEV widths)
                        const int nch = min(32, a.Cn_valid - c0);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < nch) {
                                float x = f[i];
                                if (partial) x += __ldg(partial + eoff + i);
                                if (SPLIT) x *= oscale;
                                if (a.shift) x += __ldg(a.shift + cout_off + c0 + i);
                                // split rows: logical channel -> (hi, lo) slots of its 16-channel block
                                const size_t e16 = SPLIT ? vox * (2 * ostride_w) + split_idx(cout_off + c0 + i) : eoff + i;
                                if (a.residual) {
                                    const uint16_t* rp = reinterpret_cast<const uint16_t*>(a.residual) + e16;
                                    x += SPLIT ? load16(rp, 1) + load16(rp + 16, 1) : load16(rp, f16);
                                }
                                x = stb_act(x, ACT);
                                if (out_fp32) reinterpret_cast<float*>(a.out)[eoff + i] = x;
                                else if (SPLIT) {
                                    const __half h = __float2half_rn(x);
                                    uint16_t* op = reinterpret_cast<uint16_t*>(a.out) + e16;
                                    *reinterpret_cast<__half*>(op) = h;
                                    *reinterpret_cast<__half*>(op + 16) = __float2half_rn(x - __half2float(h));
                                } else store16(reinterpret_cast<uint16_t*>(a.out) + eoff + i, x, f16);
                            }
                        }
                    }
                }
              }
            }
            if (dbg & 2) {                   // (timing experiment: no epilogue work, just hand the buffer back)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
            if (trace) trace_buf[ground * 8 + 5] = clock64();
          }
        }
    }
    tc_fence_before();
    if (a.cluster > 1) cluster_sync_all();    // peers may still multicast plane data / slot releases into this CTA until they are done
    else __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

template <int ACT, bool F16, int LEAN, bool SPLIT = false>
int launch_one_impl(unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tx, const CUtensorMap& tw, UArgs& a) {
    static bool attr_set = false;
    static uint32_t smem_base = 0;      // per kernel instance: address of the aligned dynamic-smem base in the shared window
    if (!attr_set) {
        cudaFuncSetAttribute(conv3d_umma_kernel<ACT, F16, LEAN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP);
        // one-off query launch (same kernel, same dynamic-smem attribute; the base does not depend on the size)
        uint32_t* dev = nullptr;
        if (cudaMalloc(&dev, sizeof(uint32_t)) != cudaSuccess) return STB_E_DRIVER;
        UArgs q;
        memset(&q, 0, sizeof(q));
        q.ntiles = -1;
        q.nslices = 1;
        q.out = dev;
        conv3d_umma_kernel<ACT, F16, LEAN, SPLIT><<<1, UMMA_THREADS, 4096, st>>>(tx, tw, q);
        cudaError_t e = cudaMemcpyAsync(&smem_base, dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(dev);
        if (e != cudaSuccess || smem_base == 0) return STB_E_DRIVER;
        attr_set = true;
    }
    // (in a cluster launch cvta-to-shared returns shared::cluster window addresses: CTA rank in bits 24+, the same CTA-local
    //  offset below -- the kernel compares the low 24 bits)
    const uint32_t base_used = smem_base;
    // absolute encoded operand addresses: weights at base + 2048, plane ring after the (1 KB-rounded) weight block
    const uint32_t w0 = base_used + 2048;
    const uint32_t p0 = w0 + ((a.w_bytes_total + 1023) & ~1023u);
    a.smem_base = base_used;
    for (int t = 0; t < a.ntaps_total; ++t) {
        a.iss[t].a16 = ((p0 + (uint32_t)a.taps[t].sub * a.chunk_bytes + (uint32_t)a.taps[t].rowoff * a.ROWB) >> 4) | (1u << 16);
        a.iss[t].b16 = ((w0 + (uint32_t)a.taps[t].widx * a.wtile_bytes) >> 4) | (1u << 16);
        a.iss[t].idesc = instr_desc_f16(128, (uint32_t)a.taps[t].nblk * a.Cn, F16 ? 0 : 1);
        a.iss[t].dcol = (uint32_t)a.taps[t].cls0 * a.Cn;
    }
    for (int c = 0; c < a.nclass; ++c) a.iss[a.cls[c].tap_begin].dcol |= 1u << 31;      // overwrite instead of accumulate
    a.desc_hi = (((8u * a.ROWB) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)a.layout << 29);
    if (a.cluster > 1) {
        // the nslices CTAs of a tile as one thread-block cluster (consecutive blockIdx.x): plane loads are multicast
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(UMMA_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)a.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // persistent CTAs: no more clusters than can be resident at once (a cluster must fit into one GPC)
        static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int& mc = max_clusters[a.cluster];
        if (mc == 0) {
            int n = 0;
            cudaLaunchConfig_t probe = cfg;
            probe.dynamicSmemBytes = SMEM_CAP;
            if (cudaOccupancyMaxActiveClusters(&n, conv3d_umma_kernel<ACT, F16, LEAN, SPLIT>, &probe) != cudaSuccess || n <= 0) {
                cudaGetLastError();
                n = -1;
            }
            mc = n;
        }
        if (mc > 0 && grid > (unsigned)(mc * a.cluster)) cfg.gridDim.x = (unsigned)(mc * a.cluster);
        if (cudaLaunchKernelEx(&cfg, conv3d_umma_kernel<ACT, F16, LEAN, SPLIT>, tx, tw, a) != cudaSuccess) {
            STB_CHECK_LAUNCH();
            return STB_E_DRIVER;
        }
        STB_CHECK_LAUNCH();
        return STB_OK;
    }
    conv3d_umma_kernel<ACT, F16, LEAN, SPLIT><<<grid, UMMA_THREADS, smem, st>>>(tx, tw, a);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

bool g_trace_armed = false;      // host mirror of g_umma_trace != nullptr (stb_conv3d_umma_set_trace)

template <int ACT, bool F16, bool SPLIT>
int launch_one(unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tx, const CUtensorMap& tw, UArgs& a) {
    // own instantiation for the single-channel classifiers (STB_UMMA_CLS1=0 falls back to the generic epilogue)
    static const bool cls1 = getenv("STB_UMMA_CLS1") == nullptr || atoi(getenv("STB_UMMA_CLS1")) != 0;
    if constexpr (ACT == STB_ACT_NONE) {
        if (cls1 && !a.debug && !g_trace_armed && !a.partial && a.out_fp32 && a.Cn_valid == 1 && !a.residual &&
            a.cblocks == 3 && a.merge == 3 && a.Cn <= 32)
            return launch_one_impl<ACT, F16, 4, SPLIT>(grid, smem, st, tx, tw, a);
    }
    const bool lean = !a.debug && !g_trace_armed && !a.partial && !a.out_fp32 && (a.Cn_valid & 31) == 0;
    // merged transposed conv with w-paired stores (LEAN 5; STB_UMMA_T2PAIR=0 falls back to LEAN 3)
    static const bool t2pair = getenv("STB_UMMA_T2PAIR") == nullptr || atoi(getenv("STB_UMMA_T2PAIR")) != 0;
    if constexpr (ACT == STB_ACT_RELU) {
        if (t2pair && lean && a.cblocks == 8 && a.merge == 1 && a.Cn == 32 && a.shift && a.cperm == 0x76543210u)
            return launch_one_impl<ACT, F16, 5, SPLIT>(grid, smem, st, tx, tw, a);
    }
    // merged transposed conv on 16-channel output slices (all K-chunks in TMEM, see LEAN 6 in the epilogue)
    if constexpr (ACT == STB_ACT_RELU && SPLIT) {
        if (!a.debug && !g_trace_armed && !a.partial && !a.out_fp32 && a.cblocks == 8 && a.merge == 1 && a.Cn == 16 &&
            a.Cn_valid == 16 && a.shift)
            return launch_one_impl<ACT, F16, 6, SPLIT>(grid, smem, st, tx, tw, a);
    }
    // ... and the same flavour as a K-split pass (grouped pseudo-depth chunks: fp32 partial in and / or out, LEAN 9)
    if constexpr ((ACT == STB_ACT_RELU || ACT == STB_ACT_NONE) && SPLIT) {
        if (!a.debug && !g_trace_armed && (a.partial || a.out_fp32) && a.kdepth > 0 && a.cblocks == 8 && a.merge == 1 && a.Cn == 16 &&
            a.Cn_valid == 16 && (a.Cout_total & 7) == 0)
            return launch_one_impl<ACT, F16, 9, SPLIT>(grid, smem, st, tx, tw, a);
    }
    if (a.merge == 2) return launch_one_impl<ACT, F16, 7, SPLIT>(grid, smem, st, tx, tw, a);      // stride-2 pair merge (generic + 2-block realignment)
    if (lean && a.cblocks == 8 && a.merge == 1 && a.cperm == 0x76543210u) return launch_one_impl<ACT, F16, 3, SPLIT>(grid, smem, st, tx, tw, a);
    // one 32-channel M-tile per round: epilogue split by columns across the two groups (LEAN 8; STB_UMMA_CSPLIT16=0 -> LEAN 2)
    static const bool csplit16 = getenv("STB_UMMA_CSPLIT16") == nullptr || atoi(getenv("STB_UMMA_CSPLIT16")) != 0;
    if constexpr (SPLIT) {
        if (csplit16 && lean && a.cblocks == 3 && a.merge == 3 && a.nM == 1 && a.Cn == 32 && a.nclass == 1 && a.out_stride == 1 && (a.kdepth > 0 || a.sd_in == 1 && a.dzmin == 0 && a.dzmax == 0))
            return launch_one_impl<ACT, F16, 8, SPLIT>(grid, smem, st, tx, tw, a);
    }
    if (lean && a.cblocks == 3 && a.merge == 3) return launch_one_impl<ACT, F16, 2, SPLIT>(grid, smem, st, tx, tw, a);
    if (lean && a.cblocks == 1 && a.merge == 1) return launch_one_impl<ACT, F16, 1, SPLIT>(grid, smem, st, tx, tw, a);
    return launch_one_impl<ACT, F16, 0, SPLIT>(grid, smem, st, tx, tw, a);
}

int launch_umma(int act, int f16, int split, unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tx,
                const CUtensorMap& tw, UArgs& a) {
#define STB_LAUNCH_ACT(A)                                                                          \
    case A:                                                                                        \
        if (split) return launch_one<A, true, true>(grid, smem, st, tx, tw, a);                    \
        return f16 ? launch_one<A, true, false>(grid, smem, st, tx, tw, a) : launch_one<A, false, false>(grid, smem, st, tx, tw, a);
    switch (act) {
        STB_LAUNCH_ACT(STB_ACT_NONE)
        STB_LAUNCH_ACT(STB_ACT_RELU)
        STB_LAUNCH_ACT(STB_ACT_LEAKY)
        STB_LAUNCH_ACT(STB_ACT_MISH)
        // gate activations of the ConvGRU (update block of the iterative models): split storage only
        case STB_ACT_SIGMOID: return split ? launch_one<STB_ACT_SIGMOID, true, true>(grid, smem, st, tx, tw, a) : STB_E_UNSUPPORTED;
        case STB_ACT_TANH: return split ? launch_one<STB_ACT_TANH, true, true>(grid, smem, st, tx, tw, a) : STB_E_UNSUPPORTED;
        default: return STB_E_BADARG;
    }
#undef STB_LAUNCH_ACT
}

}  // namespace

extern "C" int stb_conv3d_umma_set_trace(long long* dev_buf, int rounds) {
    if (cudaMemcpyToSymbol(g_umma_trace, &dev_buf, sizeof(dev_buf)) != cudaSuccess) return STB_E_BADARG;
    if (cudaMemcpyToSymbol(g_umma_trace_rounds, &rounds, sizeof(rounds)) != cudaSuccess) return STB_E_BADARG;
    g_trace_armed = dev_buf != nullptr;
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
    // flags bit6: operand-split fp16 ("fp16x2").  x, wt, residual and a 16-bit out hold fp16 (hi, lo) pairs interleaved per
    // 16 channels; Cin / KC count STORAGE elements (2 per logical channel), Cout_total / Cout_valid logical channels.
    const int split = (flags >> 6) & 1;
    // flags bits 16..22 (split only): the packed weights were multiplied by 2^s so that their lo halves stay normal fp16
    // numbers; the epilogue multiplies the accumulators by 2^-s.
    const float wscale_inv = ldexpf(1.f, -((flags >> 16) & 127));
    if (split && (!f16 || (KC != 32 && KC != 64) || (Cout_total % 16 && !out_fp32))) return STB_E_UNSUPPORTED;
    // flags bits 11..13: chunks per pass G of the pseudo-depth form (0 = all of them, one pass).  With 0 < G < Cin/KC the layer
    // runs as (Cin/KC)/G K-split passes of G chunks each: pass p reads the chunks [p*G, p*G + G) (taps carry dz*G + chunk-in-pass,
    // weight tiles ordered [pass][chunk in pass][tile], nwtiles = tiles of ONE pass) and chains through the fp32 partial.
    const int kgroup = (flags >> 11) & 7;
    const int kdepth = (flags & 32) ? (kgroup ? kgroup : Cin / KC) : 0; // flags bit5: K-chunks along a pseudo-depth axis, accumulated in TMEM (2-D convs)
    // (3-D layers too: pseudo-plane = depth*kdepth + chunk, taps carry dz*kdepth + chunk; unit input stride only)
    if (kdepth && in_stride != 1) return STB_E_UNSUPPORTED;
    if (kgroup && (!(flags & 32) || (Cin / KC) % kgroup)) return STB_E_BADARG;
    const bool kdepth2d = kdepth && (flags & 16);
    const int nk = kdepth ? (Cin / KC) / kdepth : Cin / KC;          // K-split passes
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
    a.merge = (flags & 4) ? 3 : ((flags & 128) ? 2 : 1);   // kw-merged taps: N = 3*Cn, weight tiles (kd,kh,0..2) are contiguous;
                                              // flags bit7: stride-2 pair merge, N = 2*Cn: the kw = 0 / 2 taps read the same parity
                                              // sub-tile one row apart, block 1 is realigned by one lane in the epilogue
    a.merge_step = ((flags >> 8) & 7) ? ((flags >> 8) & 7) : 1;   // flags bits 8..10: dilation along w of the merged taps
    if (a.merge == 3 && (in_stride != 1 || out_stride != 1 || nclass != 1)) return STB_E_UNSUPPORTED;
    if (a.merge == 2 && (in_stride != 2 || out_stride != 1 || nclass != 1 || (flags & 8))) return STB_E_UNSUPPORTED;
    a.in_stride = in_stride;
    a.nsub = in_stride == 2 ? 4 : 1;
    a.sd_in = kdepth ? kdepth : ((flags & 16) ? 1 : in_stride);     // flags bit4: 2-D convolution, the depth axis (image index) is never strided
    a.kdepth = kdepth;
    // flags bit14 (merged transposed conv): the 8 parity-class column blocks are in Gray order 0,1,3,2,6,7,5,4 instead of 0..7
    // (w-pairs stay adjacent; the class sets of the input shifts become 10 runs instead of 14 = 10 MMAs per K-step instead of 14)
    a.cperm = (flags & (1 << 14)) ? 0x45762310u : 0x76543210u;
    int maxdh = 0, maxdw = 0, dzmin = 127, dzmax = -127;
    // Taps of a class are issued in the order of the plane they read (stable sort by dz): the issuer then consumes
    // planes monotonically and can wait for / hand back each plane individually.
    int order[MAX_UTAPS];
    {
        int n = 0;
        for (int c = 0; c < nclass; ++c) {
            if (tap_begin[c] < 0 || tap_end[c] > ntaps || tap_begin[c] >= tap_end[c] || tap_begin[c] != n) return STB_E_BADARG;
            order[n++] = tap_begin[c];          // the caller's first tap initialises the accumulator blocks: keep it first
            for (int z = -127; z <= 127; ++z)
                for (int t = tap_begin[c] + 1; t < tap_end[c]; ++t)
                    if (dz[t] == z) order[n++] = t;
            if (n != tap_end[c]) return STB_E_BADARG;       // a dz outside [-127, 127]
        }
        if (n != ntaps) return STB_E_BADARG;
    }
    for (int i = 0; i < ntaps; ++i) {
        const int t = order[i];
        if (dh[t] < 0 || dw[t] < 0 || dh[t] > 6 || dw[t] > 6 || widx[t] < 0 || widx[t] >= nwtiles || sub[t] < 0 ||
            sub[t] >= a.nsub)
            return STB_E_BADARG;
        maxdh = dh[t] > maxdh ? dh[t] : maxdh;
        maxdw = dw[t] > maxdw ? dw[t] : maxdw;
        dzmin = dz[t] < dzmin ? dz[t] : dzmin;
        dzmax = dz[t] > dzmax ? dz[t] : dzmax;
        a.taps[i].dz = (int8_t)dz[t];
        a.taps[i].sub = (uint8_t)sub[t];
        a.taps[i].rowoff = (int16_t)(dh[t] * TWP + dw[t]);
        a.taps[i].widx = (uint16_t)widx[t];
        a.taps[i].nblk = (uint8_t)(nblk ? nblk[t] : (a.merge == 2 ? 1 : a.merge));
        a.taps[i].cls0 = (uint8_t)(cls0 ? cls0[t] : 0);
        if (a.taps[i].nblk < 1 || a.taps[i].nblk + a.taps[i].cls0 > 8) return STB_E_BADARG;
    }
    // plane groups: maximal runs of taps of one class with the same dz
    a.ngroups = 0;
    for (int c = 0; c < nclass; ++c) {
        a.cls[c].grp_begin = (uint32_t)a.ngroups;
        for (int t = tap_begin[c]; t < tap_end[c];) {
            int e = t;
            while (e < tap_end[c] && a.taps[e].dz == a.taps[t].dz) ++e;
            a.grp[a.ngroups].tap_begin = (uint32_t)t;
            a.grp[a.ngroups].tap_end = (uint32_t)e;
            a.grp[a.ngroups].z = (uint32_t)(a.taps[t].dz - dzmin);
            a.grp[a.ngroups].rel = 1;
            for (int l = e; l < tap_end[c]; ++l)
                if (a.taps[l].dz == a.taps[t].dz) a.grp[a.ngroups].rel = 0;
            ++a.ngroups;
            t = e;
        }
        a.cls[c].grp_end = (uint32_t)a.ngroups;
    }
    // accumulator column blocks per M-tile: 8 when taps address parity-class blocks (merged transposed conv),
    // 3 for kw-merge, else 1.  The first tap of every class must cover all blocks (it zero-initialises them).
    a.cblocks = (flags & 8) ? 8 : a.merge;
    for (int c = 0; c < nclass; ++c)
        if (a.taps[tap_begin[c]].nblk != a.cblocks || a.taps[tap_begin[c]].cls0 != 0) return STB_E_BADARG;
    if (a.cblocks == 8 && (nclass != 1 || out_stride != 2 || in_stride != 1)) return STB_E_UNSUPPORTED;
    a.ntaps_total = ntaps;
    for (int c = 0; c < nclass; ++c) {
        a.cls[c].tap_begin = (uint32_t)tap_begin[c];
        a.cls[c].tap_end = (uint32_t)tap_end[c];
        a.cls[c].od0 = od0[c]; a.cls[c].oh0 = oh0[c]; a.cls[c].ow0 = ow0[c];
    }
    a.nclass = nclass;
    a.TW = TWP - (a.merge == 3 ? 2 * a.merge_step : (a.merge == 2 && maxdw < 1 ? 1 : maxdw));
    a.dzmin = dzmin; a.dzmax = dzmax;
    const int window = dzmax - dzmin + 1;
    // Planes an accumulator round needs RESIDENT AT ONCE.  3-D convs: the whole dz window.  K-chunks along the pseudo-depth
    // axis (kdepth): the chunk planes of a round are consumed strictly one after the other and handed back right after
    // their group, so one resident plane suffices and the ring may be shorter than the round (kdepth up to 10 chunks).
    const int need = kdepth2d ? 1 : window;
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
        int r = (int)((SMEM_CAP - 3072 - wbytes - 1024) / plane);
        if (r > MAX_RING) r = MAX_RING;
        return r >= need + 1 ? r : 0;
    };
    if (need > 6 || window > 16) return STB_E_UNSUPPORTED;
    for (;;) {
        size_t best_score = 0;
        for (int th = 16; th >= 4; th -= 4) {
            const int r = ring_for(Cn, th);
            if (!r) continue;
            const size_t plane = (size_t)a.nsub * (th + maxdh) * TWP * a.ROWB;
            size_t inflight = (size_t)(r - need) * plane;
            if (inflight > 64 * 1024) inflight = 64 * 1024;   // enough to cover TMA latency; beyond that prefer tall tiles
                                                              // (>= 2 M-tiles per round keep both issuers / epilogue groups busy)
            if (inflight > best_score) { best_score = inflight; TH = th; R = r; }
        }
        if (best_score) break;
        if (Cn <= 16) return STB_E_SMEM;
        Cn = (Cn / 2 + 15) / 16 * 16;
    }
    {   // tuning overrides for sweeps with tools/layer_bench.py (host-side only; ignored when they do not fit)
        static const int force_th = getenv("STB_UMMA_TH") ? atoi(getenv("STB_UMMA_TH")) : 0;
        static const int force_ring = getenv("STB_UMMA_RING") ? atoi(getenv("STB_UMMA_RING")) : 0;
        if (force_th >= 4 && force_th <= 16 && force_th % 4 == 0) {
            const int r = ring_for(Cn, force_th);
            if (r) { TH = force_th; R = r; }
        }
        if (force_ring >= need + 1 && force_ring < R) R = force_ring;
    }
    while (TH > 4 && TH - 4 >= nclass_h) TH -= 4;      // do not stage rows that do not exist
    a.R = R;
    a.tab_bytes = 0;
    a.TH = TH; a.nM = TH * TWP / 128;
    const int box_h = TH + maxdh;
    a.chunk_bytes = (uint32_t)(box_h * TWP * a.ROWB);
    a.plane_bytes = (uint32_t)a.nsub * a.chunk_bytes;
    a.tiles_h = stb_ceil_div(nclass_h, a.TH);
    a.tiles_w = stb_ceil_div(nclass_w, a.TW);
    // all output-channel slices in one launch when they are equal and complete (see the kernel); else one launch per slice
    static const bool slice_mode = getenv("STB_UMMA_NSLICE") == nullptr || atoi(getenv("STB_UMMA_NSLICE")) != 0;
    const int ns = (slice_mode && Cpad % Cn == 0 && Cout_valid == Cpad && Cpad / Cn > 1 && Cpad / Cn <= 16) ? Cpad / Cn : 1;
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
            const int sms = num_sms / ns;                      // SMs that work on one output-channel slice
            const long long waves = (ctas + sms - 1) / sms;
            // redundant depth steps per depth chunk (none between the images of a 2-D conv)
            const int halo = kdepth ? (kdepth2d ? 0 : (window - kdepth) / kdepth) : window - 1;
            const double eff = (double)ctas / (double)(waves * sms) * (double)dch / (double)(dch + halo);
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
        // K-split passes read ONE K-chunk of every row per launch: promote no wider than the chunk (STB_TMA_PROMO overrides)
        static const int promo_env = getenv("STB_TMA_PROMO") ? atoi(getenv("STB_TMA_PROMO")) : 0;
        const int promo = promo_env > 0 ? promo_env : (nk > 1 ? KC * 2 : 256);
        if (!umma_host::make_tmap(&tm_x, cudt, 5, const_cast<void*>(x), dims, str, box, cusw, es, promo)) return STB_E_DRIVER;
    }
    a.w_rows = Cpad;
    a.w_tile_stride = kdepth ? 1 : nk;       // pseudo-depth form: the tiles of a pass are contiguous ([pass][chunk][tile])
    a.nwtiles = nwtiles;
    const long long ncta = (long long)B * a.nchunks * a.tiles_h * a.tiles_w;
    if (ncta > 2147483647LL) return STB_E_BADARG;
    a.ntiles = (int)ncta;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
    }
    static const bool verbose = getenv("STB_UMMA_VERBOSE") != nullptr;
    static const int debug = getenv("STB_UMMA_DEBUG") ? atoi(getenv("STB_UMMA_DEBUG")) : 0;
    a.debug = debug;
    for (int kp = 0; kp < nk; ++kp) {
        const bool last = kp == nk - 1;
        a.cin_off = kp * KC * (kdepth ? kdepth : 1);
        a.w_kc_off = kdepth ? kp * nwtiles : kp;
        a.partial = kp > 0 ? ws : nullptr;
        a.out = last ? out : (void*)ws;
        a.out_fp32 = last ? out_fp32 : 1;
        a.shift = last ? shift : nullptr;
        a.residual = last ? residual : nullptr;
        a.act = last ? act : STB_ACT_NONE;
        a.oscale = (split && last) ? wscale_inv : 1.f;
        for (int co = 0; co < Cpad; co += Cn * ns) {
            const int cn = (Cpad - co) < Cn ? (Cpad - co) : Cn;
            a.Cn = cn;
            a.cout_off = co;
            a.nslices = ns;
            // STB_UMMA_CLUSTER: 0 (default) = off, 1 = clusters of 2 / 4 slices, 2 = also 8.  Measured (profiles/bench_r02_progress.md,
            // trip 36): bit-identical results, no gain for 2-CTA clusters and +20 % time for 4-CTA clusters (lockstep rings, 8-14 %
            // of the SMs idle because a cluster must fit into a GPC) -- the sliced layers are not bound by L2 -> smem traffic
            static const int cluster_mode = getenv("STB_UMMA_CLUSTER") ? atoi(getenv("STB_UMMA_CLUSTER")) : 0;
            a.cluster = (cluster_mode > 0 && !debug && ns >= 2 && ns <= (cluster_mode > 1 ? 8 : 4) && ncta * ns >= 2 * ns) ? ns : 0;
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
            const long long want = ncta * ns;
            const unsigned grid = (unsigned)(want < sm_count ? want : (sm_count / ns) * ns);     // persistent: one CTA per SM
            if (verbose)
                fprintf(stderr, "[stb_conv3d_umma] Cin=%d KC=%d kpass=%d/%d Cn=%d/%d x%d taps=%d groups=%d cblocks=%d in_stride=%d out_stride=%d "
                        "TH=%d TW=%d nM=%d R=%d window=%d plane=%uB weights=%uB smem=%zuB dchunk=%d items=%d grid=%u\n",
                        Cin, KC, kp, nk, cn, Cpad, ns, ntaps, a.ngroups, a.cblocks, in_stride, out_stride, a.TH, a.TW, a.nM, a.R,
                        window, a.plane_bytes, a.w_bytes_total, smem, a.dchunk, a.ntiles, grid);
            const int rc = launch_umma(a.act, f16, split, grid, smem, (cudaStream_t)stream, tm_x, tm_w, a);
            if (rc != STB_OK) return rc;
        }
    }
    return STB_OK;
}
