// ACVNet-specific pieces of the cost-volume path (SURVEY.md section 8a row a5/a7 callers):
//   stb_patch_dw_f32     depthwise (1,3,3) dilated "patch" convolution      ACVNet/acv.py:109-112, 169-173
//   stb_block_attention  multi-head self-attention inside (b0,b1,b2) blocks ACVNet/submodule.py:381-428
// Both are tiny next to the 3-D convolutions (the attention runs on the 1/16-scale volume); they are plain
// CUDA-core kernels sized for coalesced traffic, not tensor-core work.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// out[b,c0+c,d,h,w] = sum_{a,b} w[c,a,b] * x[b,c0+c,d,h+(a-1)*dil,w+(b-1)*dil]    (zero padding)
__global__ void __launch_bounds__(256)
patch_dw_kernel(const float* __restrict__ x, const float* __restrict__ wgt, float* __restrict__ out,
                int Ctot, int c0, int C, int D, int H, int W, int dil, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % W);
    size_t r = i / W;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D); r /= D;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const size_t base = (((size_t)b * Ctot + c0 + c) * D + d) * H * (size_t)W;
    const float* wc = wgt + c * 9;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int hh = h + (a - 1) * dil;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const int ww = w + (bb - 1) * dil;
            if (ww < 0 || ww >= W) continue;
            acc = fmaf(__ldg(wc + a * 3 + bb), __ldg(x + base + (size_t)hh * W + ww), acc);
        }
    }
    out[base + (size_t)h * W + w] = acc;
}

// Four consecutive outputs per thread (W % 4 == 0, dilation <= 4): per filter row one aligned 16-byte load left, centre and
// right of the thread's quad covers every tap -- 9 loads for 4 outputs instead of 36 (the scalar kernel was LSU-bound at
// 1.55 TB/s: ACVNet 1152x1920 D=256 spent 7.2 ms in its four launches).  Same products, same summation order per output.
template <int DIL>
__global__ void __launch_bounds__(256)
patch_dw_v4_kernel(const float* __restrict__ x, const float* __restrict__ wgt, float* __restrict__ out,
                   int Ctot, int c0, int C, int D, int H, int W, size_t total4) {
    constexpr int dil = DIL;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int W4 = W >> 2;
    const int w = (int)(i % W4) << 2;
    size_t r = i / W4;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D); r /= D;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const size_t base = (((size_t)b * Ctot + c0 + c) * D + d) * H * (size_t)W;
    const float* wc = wgt + c * 9;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int hh = h + (a - 1) * dil;
        if (hh < 0 || hh >= H) continue;
        const float* row = x + base + (size_t)hh * W + w;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 lf = w >= 4 ? __ldg(reinterpret_cast<const float4*>(row - 4)) : z;
        const float4 ce = __ldg(reinterpret_cast<const float4*>(row));
        const float4 rt = w + 4 < W ? __ldg(reinterpret_cast<const float4*>(row + 4)) : z;
        const float v[12] = {lf.x, lf.y, lf.z, lf.w, ce.x, ce.y, ce.z, ce.w, rt.x, rt.y, rt.z, rt.w};
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const float wv = __ldg(wc + a * 3 + bb);
            const int off = 4 + (bb - 1) * dil;            // compile time; dil <= 4: 0 <= off + j <= 11
            // (columns outside the image are zero in v: the left / right quads are zero-filled, and an out-of-range column can
            //  only occur there)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv, v[off + j], acc[j]);
        }
    }
    *reinterpret_cast<float4*>(out + base + (size_t)h * W + w) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_f<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct AttArgs {
    int C, heads, D, H0, W0, b0, b1, b2, nd, nh, nw;
    int mask_on, mask_row0, mask_col0;           // token (y,x) is "padded-class" when y >= row0 || x >= col0
    long long qs[5], os[5];                      // element strides (b, c, d, h, w) of qkv / out
    float scale;
};

constexpr int ATT_MAX_TOK = 128;

// One CTA per (block, head, batch); one thread per query token.  K and V of the block live in shared memory.
template <typename T, int HD>
__global__ void __launch_bounds__(ATT_MAX_TOK)
block_attention_kernel(const T* __restrict__ qkv, const float* __restrict__ bias, T* __restrict__ out, AttArgs a) {
    __shared__ float Ks[ATT_MAX_TOK][HD + 1];
    __shared__ float Vs[ATT_MAX_TOK][HD + 1];
    __shared__ unsigned char Ms[ATT_MAX_TOK];
    const int nt = a.b0 * a.b1 * a.b2;
    const int head = blockIdx.x % a.heads;
    int blk = blockIdx.x / a.heads;
    const int bw = blk % a.nw; blk /= a.nw;
    const int bh = blk % a.nh;
    const int bd = blk / a.nh;
    const int b = blockIdx.y;
    const int t = threadIdx.x;
    const bool live = t < nt;
    int z = 0, y = 0, x = 0;
    bool inside = false;
    float q[HD];
    if (live) {
        const int t2 = t % a.b2, t1 = (t / a.b2) % a.b1, t0 = t / (a.b2 * a.b1);
        z = bd * a.b0 + t0; y = bh * a.b1 + t1; x = bw * a.b2 + t2;
        inside = (y < a.H0) && (x < a.W0);
        Ms[t] = (unsigned char)(a.mask_on && (y >= a.mask_row0 || x >= a.mask_col0));
        const int cq = head * HD;
        if (inside) {
            const T* p = qkv + (size_t)b * a.qs[0] + (size_t)z * a.qs[2] + (size_t)y * a.qs[3] + (size_t)x * a.qs[4];
#pragma unroll
            for (int e = 0; e < HD; ++e) {
                q[e] = ld_f(p + (size_t)(cq + e) * a.qs[1]);
                Ks[t][e] = ld_f(p + (size_t)(a.C + cq + e) * a.qs[1]);
                Vs[t][e] = ld_f(p + (size_t)(2 * a.C + cq + e) * a.qs[1]);
            }
        } else {        // zero-padded token: the Linear leaves only its bias
#pragma unroll
            for (int e = 0; e < HD; ++e) {
                q[e] = __ldg(bias + cq + e);
                Ks[t][e] = __ldg(bias + a.C + cq + e);
                Vs[t][e] = __ldg(bias + 2 * a.C + cq + e);
            }
        }
    }
    __syncthreads();
    if (!live || !inside) return;
    const int mt = Ms[t];
    float mx = -INFINITY;
    for (int j = 0; j < nt; ++j) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < HD; ++e) s = fmaf(q[e], Ks[j][e], s);
        s = s * a.scale + (Ms[j] != mt ? -1000.f : 0.f);
        mx = fmaxf(mx, s);
    }
    float den = 0.f, o[HD];
#pragma unroll
    for (int e = 0; e < HD; ++e) o[e] = 0.f;
    for (int j = 0; j < nt; ++j) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < HD; ++e) s = fmaf(q[e], Ks[j][e], s);
        s = s * a.scale + (Ms[j] != mt ? -1000.f : 0.f);
        const float p = expf(s - mx);
        den += p;
#pragma unroll
        for (int e = 0; e < HD; ++e) o[e] = fmaf(p, Vs[j][e], o[e]);
    }
    const float inv = 1.f / den;
    T* po = out + (size_t)b * a.os[0] + (size_t)z * a.os[2] + (size_t)y * a.os[3] + (size_t)x * a.os[4];
#pragma unroll
    for (int e = 0; e < HD; ++e) st_f(po + (size_t)(head * HD + e) * a.os[1], o[e] * inv);
}

template <typename T>
int launch_attention(const void* qkv, const float* bias, void* out, const AttArgs& a, int hd, int nblocks, int B,
                     cudaStream_t st) {
    dim3 grid(nblocks * a.heads, B);
    const T* q = static_cast<const T*>(qkv);
    T* o = static_cast<T*>(out);
    switch (hd) {
        case 4: block_attention_kernel<T, 4><<<grid, ATT_MAX_TOK, 0, st>>>(q, bias, o, a); break;
        case 8: block_attention_kernel<T, 8><<<grid, ATT_MAX_TOK, 0, st>>>(q, bias, o, a); break;
        case 16: block_attention_kernel<T, 16><<<grid, ATT_MAX_TOK, 0, st>>>(q, bias, o, a); break;
        case 32: block_attention_kernel<T, 32><<<grid, ATT_MAX_TOK, 0, st>>>(q, bias, o, a); break;
        default: return STB_E_UNSUPPORTED;
    }
    STB_CHECK_LAUNCH();
    return STB_OK;
}

}  // namespace

extern "C" int stb_patch_dw_f32(const float* x, const float* weight, float* out, int B, int c_total, int c_off, int C,
                                int D, int H, int W, int dilation, void* stream) {
    if (!x || !weight || !out) return STB_E_BADARG;
    if (B <= 0 || C <= 0 || c_off < 0 || c_off + C > c_total || D <= 0 || H <= 0 || W <= 0 || dilation < 1)
        return STB_E_BADARG;
    if (x == out) return STB_E_BADARG;          // taps read neighbours: not an in-place operation
    const size_t total = (size_t)B * C * D * H * W;
    if ((W & 3) == 0 && dilation <= 4 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0) {
        const size_t total4 = total >> 2;
        const unsigned g = (unsigned)((total4 + 255) / 256);
        cudaStream_t st = (cudaStream_t)stream;
        switch (dilation) {
            case 1: patch_dw_v4_kernel<1><<<g, 256, 0, st>>>(x, weight, out, c_total, c_off, C, D, H, W, total4); break;
            case 2: patch_dw_v4_kernel<2><<<g, 256, 0, st>>>(x, weight, out, c_total, c_off, C, D, H, W, total4); break;
            case 3: patch_dw_v4_kernel<3><<<g, 256, 0, st>>>(x, weight, out, c_total, c_off, C, D, H, W, total4); break;
            default: patch_dw_v4_kernel<4><<<g, 256, 0, st>>>(x, weight, out, c_total, c_off, C, D, H, W, total4); break;
        }
        STB_CHECK_LAUNCH();
        return STB_OK;
    }
    patch_dw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, weight, out, c_total, c_off, C,
                                                                                      D, H, W, dilation, total);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_block_attention(const void* qkv, const float* qkv_bias, void* out, int dtype, int B, int C, int heads,
                                   int D, int H0, int W0, int b0, int b1, int b2, const long long* qkv_strides,
                                   const long long* out_strides, void* stream) {
    if (!qkv || !qkv_bias || !out || !qkv_strides || !out_strides) return STB_E_BADARG;
    if (B <= 0 || C <= 0 || heads <= 0 || C % heads || D <= 0 || H0 <= 0 || W0 <= 0 || b0 <= 0 || b1 <= 0 || b2 <= 0)
        return STB_E_BADARG;
    if (D % b0) return STB_E_BADARG;            // the reference's view() fails as well: D is never padded
    if (b0 * b1 * b2 > ATT_MAX_TOK) return STB_E_UNSUPPORTED;
    if (dtype < 0 || dtype > 2) return STB_E_UNSUPPORTED;
    AttArgs a;
    a.C = C; a.heads = heads; a.D = D; a.H0 = H0; a.W0 = W0; a.b0 = b0; a.b1 = b1; a.b2 = b2;
    const int pad_r = (b2 - W0 % b2) % b2, pad_b = (b1 - H0 % b1) % b1;
    const int H = H0 + pad_b, W = W0 + pad_r;
    a.nd = D / b0; a.nh = H / b1; a.nw = W / b2;
    // ACVNet/submodule.py:403-406: mask[:, -pad_b:, :] = 1; mask[:, :, -pad_r:] = 1 -- with pad == 0 the slice
    // "-0:" is the whole axis, so a single zero pad marks every token and the mask cancels out.
    a.mask_on = (pad_r > 0 || pad_b > 0) ? 1 : 0;
    a.mask_row0 = pad_b > 0 ? H - pad_b : 0;
    a.mask_col0 = pad_r > 0 ? W - pad_r : 0;
    for (int i = 0; i < 5; ++i) { a.qs[i] = qkv_strides[i]; a.os[i] = out_strides[i]; }
    const int hd = C / heads;
    a.scale = (float)(1.0 / sqrt((double)hd));   // python: head_dim ** -0.5 in double, then fp32
    const int nblocks = a.nd * a.nh * a.nw;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == 0) return launch_attention<float>(qkv, qkv_bias, out, a, hd, nblocks, B, st);
    if (dtype == 1) return launch_attention<__half>(qkv, qkv_bias, out, a, hd, nblocks, B, st);
    return launch_attention<__nv_bfloat16>(qkv, qkv_bias, out, a, hd, nblocks, B, st);
}
