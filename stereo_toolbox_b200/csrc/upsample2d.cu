// Learned convex 9-tap upsampling of the iterative models, one pass each, HBM-bound:
//   * RAFT-Stereo  upsample_flow   (RAFTStereo/raft_stereo.py:81-93): softmax over the 9 mask logits of every fine pixel,
//     convex combination of the 3x3 coarse neighbourhood of factor*flow.  The reference materialises the softmax
//     [N,1,9,f,f,H,W], the unfold [N,D*9,H*W] and the product before the sum and the 6-D permute.
//   * IGEV         context_upsample (IGEVStereo/submodule.py:243-255): weights [B,9,4h,4w] (optionally soft-maxed here:
//     igev_stereo.py:164 applies F.softmax right before) times the 3x3 neighbourhood of the coarse parent.
#include "common.cuh"

namespace {

constexpr int UP_THREADS = 128;

// grid (ceil(W/128), H*f, N): thread = (coarse x, fine row fy of coarse row y); it produces the f fine pixels
// (fx = 0..f-1) of that row for every flow channel and stores them contiguously.
template <int F>
__global__ void __launch_bounds__(UP_THREADS)
convex_upsample_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ out, int D,
                       int H, int W) {
    const int x = blockIdx.x * UP_THREADS + threadIdx.x;
    const int yf = blockIdx.y, y = yf / F, fy = yf - y * F, n = blockIdx.z;
    if (x >= W) return;
    const size_t plane = (size_t)H * W;
    const float* mp = mask + (size_t)n * 9 * F * F * plane + (size_t)y * W + x;      // channel = tap*F*F + fy*F + fx
    float wgt[F][9];
#pragma unroll
    for (int fx = 0; fx < F; ++fx) {
        float m = -INFINITY;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            wgt[fx][t] = __ldg(mp + (size_t)(t * F * F + fy * F + fx) * plane);
            m = fmaxf(m, wgt[fx][t]);
        }
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            wgt[fx][t] = __expf(wgt[fx][t] - m);
            s += wgt[fx][t];
        }
        const float inv = 1.f / s;
#pragma unroll
        for (int t = 0; t < 9; ++t) wgt[fx][t] *= inv;
    }
    for (int d = 0; d < D; ++d) {
        const float* fp = flow + ((size_t)n * D + d) * plane;
        float nb[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            nb[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)F * __ldg(fp + (size_t)yy * W + xx) : 0.f;   // unfold pads zeros
        }
        float* op = out + (((size_t)n * D + d) * H * F + yf) * ((size_t)W * F) + (size_t)x * F;
        float r[F];
#pragma unroll
        for (int fx = 0; fx < F; ++fx) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) acc = fmaf(wgt[fx][t], nb[t], acc);
            r[fx] = acc;
        }
        if (F % 4 == 0) {
#pragma unroll
            for (int q = 0; q < F / 4; ++q) reinterpret_cast<float4*>(op)[q] = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        } else {
#pragma unroll
            for (int fx = 0; fx < F; ++fx) op[fx] = r[fx];
        }
    }
}

// thread per fine pixel (coalesced along X): 9 weights of its own, 3x3 neighbourhood of the coarse parent (L1/L2 hits)
__global__ void __launch_bounds__(256)
context_upsample_kernel(const float* __restrict__ disp, const float* __restrict__ wts, float* __restrict__ out, int h,
                        int w, int f, float scale, int softmax) {
    const int X = blockIdx.x * 256 + threadIdx.x, Y = blockIdx.y, b = blockIdx.z;
    const int HW = h * f, WW = w * f;
    if (X >= WW) return;
    const size_t fine = (size_t)HW * WW;
    const float* wp = wts + (size_t)b * 9 * fine + (size_t)Y * WW + X;
    float wv[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t] = __ldg(wp + (size_t)t * fine);
    if (softmax) {
        float m = wv[0];
#pragma unroll
        for (int t = 1; t < 9; ++t) m = fmaxf(m, wv[t]);
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) { wv[t] = __expf(wv[t] - m); s += wv[t]; }
        const float inv = 1.f / s;
#pragma unroll
        for (int t = 0; t < 9; ++t) wv[t] *= inv;
    }
    const int y = Y / f, x = X / f;
    const float* dp = disp + (size_t)b * h * w;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? scale * __ldg(dp + (size_t)yy * w + xx) : 0.f;
        acc = fmaf(wv[t], v, acc);
    }
    out[(size_t)b * fine + (size_t)Y * WW + X] = acc;
}

}  // namespace

extern "C" int stb_convex_upsample_f32(const float* flow, const float* mask, float* out, int N, int D, int H, int W,
                                       int factor, void* stream) {
    if (!flow || !mask || !out || N <= 0 || D <= 0 || H <= 0 || W <= 0) return STB_E_BADARG;
    if (N > 65535 || (long long)H * factor > 65535) return STB_E_BADARG;
    dim3 grid(stb_ceil_div(W, UP_THREADS), H * factor, N);
    switch (factor) {
        case 2: convex_upsample_kernel<2><<<grid, UP_THREADS, 0, (cudaStream_t)stream>>>(flow, mask, out, D, H, W); break;
        case 4: convex_upsample_kernel<4><<<grid, UP_THREADS, 0, (cudaStream_t)stream>>>(flow, mask, out, D, H, W); break;
        case 8: convex_upsample_kernel<8><<<grid, UP_THREADS, 0, (cudaStream_t)stream>>>(flow, mask, out, D, H, W); break;
        default: return STB_E_UNSUPPORTED;
    }
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_context_upsample_f32(const float* disp_low, const float* weights, float* out, int B, int h, int w,
                                        int factor, float scale, int apply_softmax, void* stream) {
    if (!disp_low || !weights || !out || B <= 0 || h <= 0 || w <= 0 || factor <= 0) return STB_E_BADARG;
    if (B > 65535 || (long long)h * factor > 65535) return STB_E_BADARG;
    dim3 grid(stb_ceil_div(w * factor, 256), h * factor, B);
    context_upsample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(disp_low, weights, out, h, w, factor, scale, apply_softmax);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
