// Exact-fp32 3-D convolution family on CUDA cores ("tap-list" direct convolution, NCDHW).
//
// Every conv flavour of the stacked hourglass is expressed as a list of taps
// (input offset dd/dh/dw, weight slice wt[t][ci][co]) over a class of output positions:
//   Conv3d k3 s1/s2 p1, k1        PSMNet/submodule.py:16-19, GwcNet/gwcnet.py:72-93
//   ConvTranspose3d k3 s2 p1 op1  PSMNet/stackhourglass.py:25-29  (8 output-parity classes, 1..8 taps)
//   ConvTranspose3d k4 s2 p1      IGEVStereo/igev_stereo.py:43-50 (8 classes x 8 taps)
// Epilogue: + shift (folded eval BatchNorm3d) + residual, then ReLU / LeakyReLU / Mish.
//
// CTA = 128 threads = 4 warps.  Tile: 32 output channels x (1 x 4 x 32) positions; warp w owns
// output channels [8w, 8w+8) (its weight loads are warp-uniform broadcasts), lane = (row, quad):
// 4 consecutive w positions.  Input channels are staged in chunks of CK through shared memory
// together with the [tap][ck][32] weight slab.  This is the bit-faithful path (fp32 FMA, same
// summation structure per output up to ordering); the tensor-core path lives in conv3d_umma.cu.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ float ld16(const void* base, size_t i, int f16) {
    return f16 ? __half2float(reinterpret_cast<const __half*>(base)[i])
               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
}
__device__ __forceinline__ uint32_t pk16(float a, float b, int f16) {
    if (f16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st16(void* base, size_t i, float v, int f16) {
    if (f16) reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16(v);
}

constexpr int CT_THREADS = 128;
constexpr int CO_TILE = 32;
constexpr int TH = 4, TW = 32;
constexpr int MAX_TAPS = 64;

struct ConvArgs {
    const void* x; const float* wt; const float* shift; const void* residual; void* out;
    int out_fp32, f16;
    int B, Cin, Di, Hi, Wi, Cout, Do, Ho, Wo;
    int ntaps, is, os, od0, oh0, ow0, nd, nh, nw, act;
    int min_d, min_h, min_w, ED, EH, EW, EWp, CK;   // staged input extent per channel
    int tiles_h, tiles_w;
    signed char dd[MAX_TAPS], dh[MAX_TAPS], dw[MAX_TAPS];
};

// CL = false: fp32 NCDHW tensors.  CL = true: bf16 channels-last NDHWC tensors (the layout of the tcgen05
// path; this kernel is its CUDA-core companion for layer shapes conv3d_umma.cu does not take yet).
template <bool CL>
__global__ void __launch_bounds__(CT_THREADS)
conv3d_taps_kernel(const __grid_constant__ ConvArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int CK = a.CK;
    float* xs = smem;                                   // [CK][ED][EH][EWp]
    float* ws = smem + ((CK * a.ED * a.EH * a.EWp + 3) & ~3);   // [ntaps][CK][CO_TILE], 16-byte aligned (read as float4)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = lane >> 3, quad = lane & 7;

    int t = blockIdx.x;
    const int tw_i = t % a.tiles_w; t /= a.tiles_w;
    const int th_i = t % a.tiles_h; t /= a.tiles_h;
    const int jd = t % a.nd;
    const int b = t / a.nd;
    const int co0 = blockIdx.y * CO_TILE;
    const int jh0 = th_i * TH, jw0 = tw_i * TW;

    // input-space origin of the staged tile
    const int id0 = jd * a.is + a.min_d, ih0 = jh0 * a.is + a.min_h, iw0 = jw0 * a.is + a.min_w;
    const size_t in_plane = (size_t)a.Hi * a.Wi, in_vol = (size_t)a.Di * in_plane;
    const int tile_elems = a.ED * a.EH * a.EWp;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < a.Cin; c0 += CK) {
        __syncthreads();
        // stage input tile (zero padded)
        for (int i = threadIdx.x; i < CK * tile_elems; i += CT_THREADS) {
            int ck, r;
            if (CL) { ck = i % CK; r = i / CK; } else { ck = i / tile_elems; r = i - ck * tile_elems; }
            const int r0 = r;
            int ed = r / (a.EH * a.EWp); r -= ed * a.EH * a.EWp;
            int eh = r / a.EWp, ew = r - eh * a.EWp;
            int ci = c0 + ck, id = id0 + ed, ih = ih0 + eh, iw = iw0 + ew;
            float v = 0.f;
            if (ci < a.Cin && ew < a.EW && (unsigned)id < (unsigned)a.Di && (unsigned)ih < (unsigned)a.Hi &&
                (unsigned)iw < (unsigned)a.Wi) {
                if (CL)
                    v = ld16(a.x, ((((size_t)b * a.Di + id) * a.Hi + ih) * a.Wi + iw) * a.Cin + ci, a.f16);
                else
                    v = __ldg(reinterpret_cast<const float*>(a.x) + ((size_t)b * a.Cin + ci) * in_vol +
                              (size_t)id * in_plane + (size_t)ih * a.Wi + iw);
            }
            xs[ck * tile_elems + r0] = v;
        }
        // stage weights [tap][ck][co]
        for (int i = threadIdx.x; i < a.ntaps * CK * CO_TILE; i += CT_THREADS) {
            int co = i % CO_TILE, r = i / CO_TILE;
            int ck = r % CK, tp = r / CK;
            int ci = c0 + ck;
            float v = 0.f;
            if (ci < a.Cin && co0 + co < a.Cout) v = __ldg(a.wt + ((size_t)tp * a.Cin + ci) * a.Cout + co0 + co);
            ws[i] = v;
        }
        __syncthreads();
        for (int tp = 0; tp < a.ntaps; ++tp) {
            const int off = ((a.dd[tp] - a.min_d) * a.EH + (row * a.is + a.dh[tp] - a.min_h)) * a.EWp +
                            (quad * 4 * a.is + a.dw[tp] - a.min_w);
            const float* wrow = ws + (size_t)tp * CK * CO_TILE + warp * 8;
            const float* xrow = xs + off;
            for (int ck = 0; ck < CK; ++ck) {
                const float4 w0 = *reinterpret_cast<const float4*>(wrow + ck * CO_TILE);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + ck * CO_TILE + 4);
                const float* xp = xrow + ck * tile_elems;
                float xv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] = xp[j * a.is];
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
            }
        }
    }

    // epilogue
    const int jh = jh0 + row;
    if (jh >= a.nh) return;
    const int od = jd * a.os + a.od0, oh = jh * a.os + a.oh0;
    const size_t out_plane = (size_t)a.Ho * a.Wo, out_vol = (size_t)a.Do * out_plane;
    if (CL) {
        const int cob = co0 + warp * 8;
        if (cob >= a.Cout) return;
        const int nch = min(8, a.Cout - cob);
        const void* res = a.residual;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jw = jw0 + quad * 4 + j;
            if (jw >= a.nw) continue;
            const size_t vox = (((size_t)b * a.Do + od) * a.Ho + oh) * a.Wo + (size_t)(jw * a.os + a.ow0);
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = acc[i][j];
                if (i < nch) {
                    if (a.shift) v += __ldg(a.shift + cob + i);
                    if (res) v += ld16(res, vox * a.Cout + cob + i, a.f16);
                }
                f[i] = stb_act(v, a.act);
            }
            if (a.out_fp32) {
                float* op = reinterpret_cast<float*>(a.out) + vox * a.Cout + cob;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < nch) op[i] = f[i];
            } else {
                uint16_t* op = reinterpret_cast<uint16_t*>(a.out) + vox * a.Cout + cob;
                if (nch == 8 && (a.Cout & 7) == 0) {
                    uint4 o;
                    o.x = pk16(f[0], f[1], a.f16); o.y = pk16(f[2], f[3], a.f16);
                    o.z = pk16(f[4], f[5], a.f16); o.w = pk16(f[6], f[7], a.f16);
                    *reinterpret_cast<uint4*>(op) = o;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (i < nch) st16(op, i, f[i], a.f16);
                }
            }
        }
        return;
    }
    const float* resf = reinterpret_cast<const float*>(a.residual);
    float* outf = reinterpret_cast<float*>(a.out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + warp * 8 + i;
        if (co >= a.Cout) break;
        const float sh = a.shift ? __ldg(a.shift + co) : 0.f;
        const size_t obase = ((size_t)b * a.Cout + co) * out_vol + (size_t)od * out_plane + (size_t)oh * a.Wo;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jw = jw0 + quad * 4 + j;
            if (jw < a.nw) {
                const size_t o = obase + (size_t)(jw * a.os + a.ow0);
                float v = acc[i][j] + sh;
                if (resf) v += __ldg(resf + o);
                outf[o] = stb_act(v, a.act);
            }
        }
    }
}

}  // namespace

static int conv3d_taps_launch(bool cl, int out_fp32, int f16, const void* x, const float* wt, const float* shift,
                              const void* residual, void* out, int B, int Cin, int Di, int Hi, int Wi, int Cout,
                              int Do, int Ho, int Wo, int ntaps, const int* dd, const int* dh, const int* dw,
                              int in_stride, int out_stride, int od0, int oh0, int ow0, int nd, int nh, int nw,
                              int act, void* stream) {
    if (!x || !wt || !out || !dd || !dh || !dw) return STB_E_BADARG;
    if (B <= 0 || Cin <= 0 || Cout <= 0 || ntaps <= 0 || ntaps > MAX_TAPS) return STB_E_BADARG;
    if (in_stride < 1 || in_stride > 2 || out_stride < 1 || out_stride > 2) return STB_E_UNSUPPORTED;
    if (nd <= 0 || nh <= 0 || nw <= 0) return STB_E_BADARG;
    if ((nd - 1) * out_stride + od0 >= Do || (nh - 1) * out_stride + oh0 >= Ho || (nw - 1) * out_stride + ow0 >= Wo)
        return STB_E_BADARG;
    ConvArgs a;
    a.x = x; a.wt = wt; a.shift = shift; a.residual = residual; a.out = out; a.out_fp32 = out_fp32; a.f16 = f16;
    a.B = B; a.Cin = Cin; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.Cout = Cout; a.Do = Do; a.Ho = Ho; a.Wo = Wo;
    a.ntaps = ntaps; a.is = in_stride; a.os = out_stride; a.od0 = od0; a.oh0 = oh0; a.ow0 = ow0;
    a.nd = nd; a.nh = nh; a.nw = nw; a.act = act;
    int mn[3] = {127, 127, 127}, mx[3] = {-127, -127, -127};
    for (int t = 0; t < ntaps; ++t) {
        const int o[3] = {dd[t], dh[t], dw[t]};
        for (int k = 0; k < 3; ++k) {
            if (o[k] < -8 || o[k] > 8) return STB_E_UNSUPPORTED;
            mn[k] = o[k] < mn[k] ? o[k] : mn[k];
            mx[k] = o[k] > mx[k] ? o[k] : mx[k];
        }
        a.dd[t] = (signed char)dd[t]; a.dh[t] = (signed char)dh[t]; a.dw[t] = (signed char)dw[t];
    }
    a.min_d = mn[0]; a.min_h = mn[1]; a.min_w = mn[2];
    a.ED = mx[0] - mn[0] + 1;
    a.EH = (TH - 1) * in_stride + mx[1] - mn[1] + 1;
    a.EW = (TW - 1) * in_stride + mx[2] - mn[2] + 1;
    a.EWp = a.EW | 1;                       // odd pitch: the 4 rows of a warp hit distinct banks
    if ((a.EWp & 3) == 1 && in_stride == 1) a.EWp += 2;  // prefer pitch = 3 (mod 4)
    a.tiles_h = stb_ceil_div(nh, TH);
    a.tiles_w = stb_ceil_div(nw, TW);
    // channel chunk: as large as fits ~96 KB so that 2 CTAs share an SM
    int CK = 8;
    auto smem_for = [&](int ck) {
        return (size_t)(((ck * a.ED * a.EH * a.EWp + 3) & ~3) + ntaps * ck * CO_TILE) * sizeof(float);
    };
    while (CK > 1 && smem_for(CK) > 96 * 1024) CK >>= 1;
    if (CK > Cin) { CK = 1; while (CK * 2 <= Cin && CK < 8) CK <<= 1; }
    a.CK = CK;
    size_t smem = smem_for(CK);
    if (smem > 200 * 1024) return STB_E_SMEM;
    long long nblk = (long long)B * nd * a.tiles_h * a.tiles_w;
    if (nblk > 2147483647LL) return STB_E_BADARG;
    dim3 grid((unsigned)nblk, stb_ceil_div(Cout, CO_TILE), 1);
    if (cl) {
        cudaFuncSetAttribute(conv3d_taps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv3d_taps_kernel<true><<<grid, CT_THREADS, smem, (cudaStream_t)stream>>>(a);
    } else {
        cudaFuncSetAttribute(conv3d_taps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv3d_taps_kernel<false><<<grid, CT_THREADS, smem, (cudaStream_t)stream>>>(a);
    }
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_conv3d_taps_f32(const float* x, const float* wt, const float* shift, const float* residual,
                                   float* out, int B, int Cin, int Di, int Hi, int Wi, int Cout, int Do, int Ho,
                                   int Wo, int ntaps, const int* dd, const int* dh, const int* dw, int in_stride,
                                   int out_stride, int od0, int oh0, int ow0, int nd, int nh, int nw, int act,
                                   void* stream) {
    return conv3d_taps_launch(false, 1, 0, x, wt, shift, residual, out, B, Cin, Di, Hi, Wi, Cout, Do, Ho, Wo, ntaps, dd,
                              dh, dw, in_stride, out_stride, od0, oh0, ow0, nd, nh, nw, act, stream);
}

extern "C" int stb_conv3d_taps_cl16(const void* x, const float* wt, const float* shift, const void* residual,
                                    void* out, int out_fp32, int f16, int B, int Cin, int Di, int Hi, int Wi, int Cout,
                                       int Do, int Ho, int Wo, int ntaps, const int* dd, const int* dh, const int* dw,
                                       int in_stride, int out_stride, int od0, int oh0, int ow0, int nd, int nh,
                                       int nw, int act, void* stream) {
    return conv3d_taps_launch(true, out_fp32, f16, x, wt, shift, residual, out, B, Cin, Di, Hi, Wi, Cout, Do, Ho, Wo,
                              ntaps, dd, dh, dw, in_stride, out_stride, od0, oh0, ow0, nd, nh, nw, act, stream);
}
