// Backward kernels of the hot path, exact fp32 / reference layouts (training path, BASELINE config 3).
//
//   stb_conv3d_wgrad_f32            weight gradient of Conv3d / ConvTranspose3d (PSMNet/submodule.py:16-19,
//                                   PSMNet/stackhourglass.py:25-29) -- the data gradient needs no kernel of its own: it is the
//                                   transposed (resp. strided) convolution of the same family, run through stb_conv3d_taps_f32
//   stb_concat_volume_bwd_f32       adjoint of build_concat_volume (GwcNet/submodule.py:30-41, PSMNet/stackhourglass.py:111-120)
//   stb_gwc_volume_bwd_f32          adjoint of build_gwc_volume (GwcNet/submodule.py:44-63)
//   stb_upsample_softargmin_bwd_f32 adjoint of F.upsample(trilinear) + softmax + disparity_regression
//                                   (PSMNet/stackhourglass.py:139-156, GwcNet/gwcnet.py:196-216)
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ weight gradient
//   dW[kd][kh][kw][cp][cq] += sum_{b, q} P[b][cp][S*q + k - pad] * Q[b][cq][q]
// Conv3d:          P = layer input, Q = grad of the output, S = stride   (weight.grad[co,ci,k] = dW[k][ci][co])
// ConvTranspose3d: P = grad of the output, Q = layer input, S = stride   (weight.grad[ci,co,k] = dW[k][co][ci])
// Persistent CTAs: blockIdx.y = (cp block, cq block) of 32x32, blockIdx.z = kd, blockIdx.x strides over row tiles
// (b, qd, qh, 32-wide w chunk).  A thread owns a 2x2 (cp, cq) block for all K*K (kh, kw) taps in registers and flushes
// them with atomics once at the end.
constexpr int WG_T = 256, WG_C = 32, WG_W = 32;

template <int K, int S>
__global__ void __launch_bounds__(WG_T)
wgrad_f32_kernel(const float* __restrict__ P, const float* __restrict__ Q, float* __restrict__ dW, int B, int Cp, int Dp,
                 int Hp, int Wp, int Cq, int Dq, int Hq, int Wq, int pad, int wchunks, long long ntiles) {
    constexpr int PW = S * (WG_W - 1) + K;                 // P columns needed by 32 consecutive q
    __shared__ float Ps[WG_C][K][PW + 1];
    __shared__ float Qs[WG_C][WG_W + 1];
    const int cqb = (Cq + WG_C - 1) / WG_C;
    const int cp0 = (blockIdx.y / cqb) * WG_C, cq0 = (blockIdx.y % cqb) * WG_C;
    const int kd = blockIdx.z;
    const int tp = threadIdx.x >> 4, tq = threadIdx.x & 15;       // thread owns cp {tp, tp+16}, cq {tq, tq+16}
    float acc[K][K][4];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b2 = 0; b2 < K; ++b2)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[a][b2][i] = 0.f;
    const size_t pplane = (size_t)Hp * Wp, qplane = (size_t)Hq * Wq;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long t = tile;
        const int wc = (int)(t % wchunks); t /= wchunks;
        const int qh = (int)(t % Hq); t /= Hq;
        const int qd = (int)(t % Dq);
        const int b = (int)(t / Dq);
        const int pd = S * qd + kd - pad;
        if (pd < 0 || pd >= Dp) continue;                        // the whole P plane is zero padding (block-uniform)
        const int qw0 = wc * WG_W, pw0 = S * qw0 - pad;
        __syncthreads();
        for (int i = threadIdx.x; i < WG_C * WG_W; i += WG_T) {
            const int c = i / WG_W, x = i - c * WG_W;
            const int cq = cq0 + c, qw = qw0 + x;
            Qs[c][x] = (cq < Cq && qw < Wq) ? __ldg(Q + (((size_t)b * Cq + cq) * Dq + qd) * qplane + (size_t)qh * Wq + qw) : 0.f;
        }
        for (int i = threadIdx.x; i < WG_C * K * PW; i += WG_T) {
            const int c = i / (K * PW), r = (i / PW) % K, x = i % PW;
            const int cp = cp0 + c, ph = S * qh + r - pad, pw = pw0 + x;
            float v = 0.f;
            if (cp < Cp && ph >= 0 && ph < Hp && pw >= 0 && pw < Wp)
                v = __ldg(P + (((size_t)b * Cp + cp) * Dp + pd) * pplane + (size_t)ph * Wp + pw);
            Ps[c][r][x] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int x = 0; x < WG_W; ++x) {
            const float q0 = Qs[tq][x], q1 = Qs[tq + 16][x];
#pragma unroll
            for (int r = 0; r < K; ++r)
#pragma unroll
                for (int kw = 0; kw < K; ++kw) {
                    const float p0 = Ps[tp][r][S * x + kw], p1 = Ps[tp + 16][r][S * x + kw];
                    acc[r][kw][0] = fmaf(p0, q0, acc[r][kw][0]);
                    acc[r][kw][1] = fmaf(p0, q1, acc[r][kw][1]);
                    acc[r][kw][2] = fmaf(p1, q0, acc[r][kw][2]);
                    acc[r][kw][3] = fmaf(p1, q1, acc[r][kw][3]);
                }
        }
    }
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
        for (int kw = 0; kw < K; ++kw)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cp = cp0 + tp + (i >> 1) * 16, cq = cq0 + tq + (i & 1) * 16;
                if (cp < Cp && cq < Cq && acc[r][kw][i] != 0.f)
                    atomicAdd(dW + ((((size_t)kd * K + r) * K + kw) * Cp + cp) * Cq + cq, acc[r][kw][i]);
            }
}

// ------------------------------------------------------------------------------------------------ volume adjoints
// dvol [B, c_total, D, H, W]; channels [c_off, c_off+C) left copy, [c_off+C, c_off+2C) right copy shifted by d.
__global__ void concat_bwd_kernel(const float* __restrict__ dvol, float* __restrict__ dL, float* __restrict__ dR, int C, int H,
                                  int W, int D, int c_total, int c_off, int mask_left, long long total) {
    const size_t plane = (size_t)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        long long t = i / W;
        const int h = (int)(t % H); t /= H;
        const int c = (int)(t % C);
        const int b = (int)(t / C);
        const float* gl = dvol + (((size_t)b * c_total + c_off + c) * D) * plane + (size_t)h * W;
        const float* gr = dvol + (((size_t)b * c_total + c_off + C + c) * D) * plane + (size_t)h * W;
        float sl = 0.f, sr = 0.f;
        for (int d = 0; d < D; ++d) {
            if (!mask_left || d <= w) sl += __ldg(gl + (size_t)d * plane + w);
            if (w + d < W) sr += __ldg(gr + (size_t)d * plane + w + d);
        }
        dL[i] = sl;
        dR[i] = sr;
    }
}

// CTA = (h, group, b): the D x W slab of dvol and the k left/right feature rows are staged in shared memory.
__global__ void __launch_bounds__(256)
gwc_bwd_kernel(const float* __restrict__ dvol, const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ dL,
               float* __restrict__ dR, int C, int H, int W, int D, int G, int c_total, int c_off) {
    extern __shared__ float sm[];
    const int h = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int k = C / G;
    float* dv = sm;                          // [D][W]
    float* Ls = dv + (size_t)D * W;          // [k][W]
    float* Rs = Ls + (size_t)k * W;          // [k][W]
    const size_t plane = (size_t)H * W;
    const float* src = dvol + (((size_t)b * c_total + c_off + g) * D) * plane + (size_t)h * W;
    for (int i = threadIdx.x; i < D * W; i += blockDim.x) {
        const int d = i / W, w = i - d * W;
        dv[i] = __ldg(src + (size_t)d * plane + w);
    }
    for (int i = threadIdx.x; i < k * W; i += blockDim.x) {
        const int c = i / W, w = i - c * W;
        const size_t off = (((size_t)b * C + (size_t)g * k + c) * H + h) * W + w;
        Ls[i] = __ldg(L + off);
        Rs[i] = __ldg(R + off);
    }
    __syncthreads();
    const float inv = 1.f / (float)k;
    for (int i = threadIdx.x; i < k * W; i += blockDim.x) {
        const int c = i / W, w = i - c * W;
        float sl = 0.f, sr = 0.f;
        const int dmax_l = min(D - 1, w), dmax_r = min(D - 1, W - 1 - w);
        for (int d = 0; d <= dmax_l; ++d) sl = fmaf(dv[d * W + w], Rs[c * W + w - d], sl);
        for (int d = 0; d <= dmax_r; ++d) sr = fmaf(dv[d * W + w + d], Ls[c * W + w + d], sr);
        const size_t off = (((size_t)b * C + (size_t)g * k + c) * H + h) * W + w;
        dL[off] = sl * inv;
        dR[off] = sr * inv;
    }
}

// ------------------------------------------------------------------------------------------------ head adjoint
constexpr int HB_THREADS = 128;

__device__ __forceinline__ void src_index_b(int dst, int in_size, float scale, bool align, int& i0, int& i1, float& t) {
    float src = align ? scale * dst : fmaxf(scale * (dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    t = src - (float)i0;
}

// Same thread mapping as upsample_softargmin_kernel (head.cu): one thread per output pixel recomputes its interpolated
// column and the softmax, then scatters d cost through the transposed interpolation (atomics into the low-res volume).
__global__ void __launch_bounds__(HB_THREADS)
upsample_softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ gdisp, float* __restrict__ dcost, int D,
                               int H, int W, int outD, int outH, int outW, float sd, float sh, float sw, int align) {
    extern __shared__ float smh[];           // col [D][T] then gcol [D][T]
    float* col = smh;
    float* gcol = smh + (size_t)D * HB_THREADS;
    const int x = blockIdx.x * HB_THREADS + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= outW) return;
    int h0, h1, w0, w1;
    float th, tw;
    src_index_b(y, H, sh, align, h0, h1, th);
    src_index_b(x, W, sw, align, w0, w1, tw);
    const size_t plane = (size_t)H * W;
    const float* base = cost + (size_t)b * D * plane;
    const float a00 = (1.f - th) * (1.f - tw), a01 = (1.f - th) * tw, a10 = th * (1.f - tw), a11 = th * tw;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) {
        const float* p = base + (size_t)d * plane;
        const float v = a00 * __ldg(p + h0 * W + w0) + a01 * __ldg(p + h0 * W + w1) + a10 * __ldg(p + h1 * W + w0) +
                        a11 * __ldg(p + h1 * W + w1);
        col[d * HB_THREADS + threadIdx.x] = v;
        gcol[d * HB_THREADS + threadIdx.x] = 0.f;
    }
    for (int od = 0; od < outD; ++od) {          // softmax shift = maximum over the interpolated bins (see head.cu)
        int d0, d1;
        float td;
        src_index_b(od, D, sd, align, d0, d1, td);
        m = fmaxf(m, (1.f - td) * col[d0 * HB_THREADS + threadIdx.x] + td * col[d1 * HB_THREADS + threadIdx.x]);
    }
    float s = 0.f, acc = 0.f;
    for (int od = 0; od < outD; ++od) {
        int d0, d1;
        float td;
        src_index_b(od, D, sd, align, d0, d1, td);
        const float v = (1.f - td) * col[d0 * HB_THREADS + threadIdx.x] + td * col[d1 * HB_THREADS + threadIdx.x];
        const float e = __expf(v - m);
        s += e;
        acc = fmaf(e, (float)od, acc);
    }
    const float disp = acc / s;
    const float g = __ldg(gdisp + ((size_t)b * outH + y) * outW + x) / s;
    for (int od = 0; od < outD; ++od) {
        int d0, d1;
        float td;
        src_index_b(od, D, sd, align, d0, d1, td);
        const float v = (1.f - td) * col[d0 * HB_THREADS + threadIdx.x] + td * col[d1 * HB_THREADS + threadIdx.x];
        const float du = g * __expf(v - m) * ((float)od - disp);       // d loss / d upsampled cost bin
        gcol[d0 * HB_THREADS + threadIdx.x] += (1.f - td) * du;
        gcol[d1 * HB_THREADS + threadIdx.x] += td * du;
    }
    float* gb = dcost + (size_t)b * D * plane;
    for (int d = 0; d < D; ++d) {
        const float gc = gcol[d * HB_THREADS + threadIdx.x];
        float* p = gb + (size_t)d * plane;
        atomicAdd(p + h0 * W + w0, a00 * gc);
        atomicAdd(p + h0 * W + w1, a01 * gc);
        atomicAdd(p + h1 * W + w0, a10 * gc);
        atomicAdd(p + h1 * W + w1, a11 * gc);
    }
}

float lin_scale(int in_size, int out_size, int align) {
    if (align) return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    return (float)in_size / (float)out_size;
}

template <int K, int S>
int launch_wgrad(const float* P, const float* Q, float* dW, int B, int Cp, int Dp, int Hp, int Wp, int Cq, int Dq, int Hq,
                 int Wq, int pad, cudaStream_t st) {
    const int wchunks = stb_ceil_div(Wq, WG_W);
    const long long ntiles = (long long)B * Dq * Hq * wchunks;
    const int cblocks = stb_ceil_div(Cp, WG_C) * stb_ceil_div(Cq, WG_C);
    long long gx = (148LL * 4 + (long long)cblocks * K - 1) / ((long long)cblocks * K);     // ~4 CTAs per SM in total
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    dim3 grid((unsigned)gx, (unsigned)cblocks, (unsigned)K);
    wgrad_f32_kernel<K, S><<<grid, WG_T, 0, st>>>(P, Q, dW, B, Cp, Dp, Hp, Wp, Cq, Dq, Hq, Wq, pad, wchunks, ntiles);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

}  // namespace

extern "C" int stb_conv3d_wgrad_f32(const float* P, const float* Q, float* dW, int B, int Cp, int Dp, int Hp, int Wp, int Cq,
                                    int Dq, int Hq, int Wq, int K, int pad, int stride, void* stream) {
    if (!P || !Q || !dW || B <= 0 || Cp <= 0 || Cq <= 0 || Dq <= 0 || Hq <= 0 || Wq <= 0) return STB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
#define STB_WG(KK, SS) \
    if (K == KK && stride == SS) return launch_wgrad<KK, SS>(P, Q, dW, B, Cp, Dp, Hp, Wp, Cq, Dq, Hq, Wq, pad, st);
    STB_WG(1, 1) STB_WG(3, 1) STB_WG(3, 2) STB_WG(4, 2)
#undef STB_WG
    return STB_E_UNSUPPORTED;
}

extern "C" int stb_concat_volume_bwd_f32(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W, int D,
                                         int mask_left, int c_total, int c_off, void* stream) {
    if (!dvol || !dleft || !dright || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0 || c_off < 0 || c_off + 2 * C > c_total)
        return STB_E_BADARG;
    const long long total = (long long)B * C * H * W;
    long long g = (total + 255) / 256;
    if (g > 148LL * 16) g = 148LL * 16;
    concat_bwd_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(dvol, dleft, dright, C, H, W, D, c_total, c_off, mask_left,
                                                                     total);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_gwc_volume_bwd_f32(const float* dvol, const float* left, const float* right, float* dleft, float* dright,
                                      int B, int C, int H, int W, int D, int G, int c_total, int c_off, void* stream) {
    if (!dvol || !left || !right || !dleft || !dright || B <= 0 || C <= 0 || G <= 0 || C % G || H <= 0 || W <= 0 || D <= 0 ||
        c_off < 0 || c_off + G > c_total)
        return STB_E_BADARG;
    if (H > 65535 || G > 65535 || B > 65535) return STB_E_BADARG;
    const size_t smem = ((size_t)D * W + 2 * (size_t)(C / G) * W) * sizeof(float);
    if (smem > 200 * 1024) return STB_E_SMEM;
    if (smem > 48 * 1024) cudaFuncSetAttribute(gwc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gwc_bwd_kernel<<<dim3(H, G, B), 256, smem, (cudaStream_t)stream>>>(dvol, left, right, dleft, dright, C, H, W, D, G, c_total,
                                                                        c_off);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_upsample_softargmin_bwd_f32(const float* cost, const float* gdisp, float* dcost, int B, int D, int H, int W,
                                               int outD, int outH, int outW, int align_corners, void* stream) {
    if (!cost || !gdisp || !dcost || B <= 0 || D <= 0 || H <= 0 || W <= 0 || outD <= 0 || outH <= 0 || outW <= 0)
        return STB_E_BADARG;
    if (B > 65535 || outH > 65535) return STB_E_BADARG;
    const size_t smem = 2 * (size_t)D * HB_THREADS * sizeof(float);
    if (smem > 200 * 1024) return STB_E_SMEM;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(upsample_softargmin_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(stb_ceil_div(outW, HB_THREADS), outH, B);
    upsample_softargmin_bwd_kernel<<<grid, HB_THREADS, smem, (cudaStream_t)stream>>>(
        cost, gdisp, dcost, D, H, W, outD, outH, outW, lin_scale(D, outD, align_corners), lin_scale(H, outH, align_corners),
        lin_scale(W, outW, align_corners), align_corners);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
