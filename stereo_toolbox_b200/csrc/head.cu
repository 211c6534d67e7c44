// Fused head: trilinear upsample (ATen upsample_trilinear3d semantics) + softmax over D +
// expected value.  Replaces GwcNet/gwcnet.py:220-223 / PSMNet/stackhourglass.py:150-156, which
// materialise a [B,192,H,W] tensor and make ~6 passes over it.
//
// One thread owns one output pixel.  Phase 1 interpolates the D input planes at (y,x) -- 4 taps
// each, served by L1/L2 because neighbouring pixels share taps -- into a shared-memory column;
// phase 2 walks the outD upsampled bins with a running softmax (max is taken over the D knots,
// an upper bound of every interpolated value).  After fusion the op is MUFU/ALU-bound.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int HEAD_THREADS = 128;

__device__ __forceinline__ void src_index(int dst, int in_size, float scale, bool align, int& i0, int& i1, float& t) {
    // at::native::area_pixel_compute_source_index + guard, float arithmetic like ATen
    float src = align ? scale * dst : fmaxf(scale * (dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    t = src - (float)i0;
}

__global__ void __launch_bounds__(HEAD_THREADS)
upsample_softargmin_kernel(const float* __restrict__ cost, float* __restrict__ disp, int D, int H, int W,
                           int outD, int outH, int outW, float sd, float sh, float sw, int align) {
    extern __shared__ float col[];   // [D][HEAD_THREADS]
    const int x = blockIdx.x * HEAD_THREADS + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= outW) return;
    int h0, h1, w0, w1;
    float th, tw;
    src_index(y, H, sh, align, h0, h1, th);
    src_index(x, W, sw, align, w0, w1, tw);
    const size_t plane = (size_t)H * W;
    const float* base = cost + (size_t)b * D * plane;
    const float a00 = (1.f - th) * (1.f - tw), a01 = (1.f - th) * tw, a10 = th * (1.f - tw), a11 = th * tw;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) {
        const float* p = base + (size_t)d * plane;
        float v = a00 * __ldg(p + h0 * W + w0) + a01 * __ldg(p + h0 * W + w1) + a10 * __ldg(p + h1 * W + w0) +
                  a11 * __ldg(p + h1 * W + w1);
        col[d * HEAD_THREADS + threadIdx.x] = v;
    }
    // softmax shift = maximum over the INTERPOLATED bins (what F.softmax subtracts), not over the knots: interior knots are
    // never sampled exactly (td >= 0.125), so for an isolated peak > ~700 above its neighbours every exp(v - max_knot) would
    // flush to zero and the result would be 0/0
    for (int od = 0; od < outD; ++od) {
        int d0, d1;
        float td;
        src_index(od, D, sd, align, d0, d1, td);
        m = fmaxf(m, (1.f - td) * col[d0 * HEAD_THREADS + threadIdx.x] + td * col[d1 * HEAD_THREADS + threadIdx.x]);
    }
    float s = 0.f, acc = 0.f;
    for (int od = 0; od < outD; ++od) {
        int d0, d1;
        float td;
        src_index(od, D, sd, align, d0, d1, td);
        float v = (1.f - td) * col[d0 * HEAD_THREADS + threadIdx.x] + td * col[d1 * HEAD_THREADS + threadIdx.x];
        float e = __expf(v - m);
        s += e;
        acc = fmaf(e, (float)od, acc);
    }
    disp[((size_t)b * outH + y) * outW + x] = acc / s;
}

// Specialisation for the shape every 3-D-conv model uses: outD == 4*D, align_corners == false.  The source index of
// upsampled bin od is max(0.25*od - 0.375, 0): bins 0,1 sit on knot 0; bin 4k+2+j (j = 0..3) lies between knots k and
// min(k+1, D-1) with weight td = 0.125 + 0.25*j -- all exact in fp32, so this kernel evaluates the SAME expressions in
// the SAME order as the generic one (same result up to the compiler's choice of FMA contraction, which is pinned here
// to the generic kernel's: FMUL (1-td)*c0, then FFMA td*c1 + that) without the per-bin index arithmetic and with each knot read
// from shared memory once: ~7 instead of ~18 instructions per exp.
__global__ void __launch_bounds__(HEAD_THREADS)
upsample_softargmin_x4_kernel(const float* __restrict__ cost, float* __restrict__ disp, int D, int H, int W,
                              int outH, int outW, float sh, float sw) {
    extern __shared__ float col[];   // [D][HEAD_THREADS]
    const int x = blockIdx.x * HEAD_THREADS + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= outW) return;
    int h0, h1, w0, w1;
    float th, tw;
    src_index(y, H, sh, false, h0, h1, th);
    src_index(x, W, sw, false, w0, w1, tw);
    const size_t plane = (size_t)H * W;
    const float* base = cost + (size_t)b * D * plane;
    const float a00 = (1.f - th) * (1.f - tw), a01 = (1.f - th) * tw, a10 = th * (1.f - tw), a11 = th * tw;
    float m = -INFINITY, pv = 0.f;
    for (int d = 0; d < D; ++d) {
        const float* p = base + (size_t)d * plane;
        float v = a00 * __ldg(p + h0 * W + w0) + a01 * __ldg(p + h0 * W + w1) + a10 * __ldg(p + h1 * W + w0) +
                  a11 * __ldg(p + h1 * W + w1);
        col[d * HEAD_THREADS + threadIdx.x] = v;
        // softmax shift = maximum over the INTERPOLATED bins (see the generic kernel): between two knots the bins are linear
        // in td, so only the outermost two (td = 0.125, 0.875) can be the largest; bins 0, 1 sit on knot 0 and the last two
        // on knot D-1.  Tracked here, while the knots pass through registers (no second pass over the column).
        m = d == 0 ? v : fmaxf(m, fmaxf(fmaf(0.125f, v, __fmul_rn(0.875f, pv)), fmaf(0.875f, v, __fmul_rn(0.125f, pv))));
        pv = v;
    }
    m = fmaxf(m, pv);
    float c0 = col[threadIdx.x];
    // bins 0 and 1: td = 0, the generic kernel computes 1*c0 + 0*c1 = c0
    float e = __expf(c0 - m);
    float s = e + e, acc = fmaf(e, 1.f, fmaf(e, 0.f, 0.f));
    float od = 2.f;
    for (int k = 0; k < D; ++k) {
        const float c1 = col[min(k + 1, D - 1) * HEAD_THREADS + threadIdx.x];
        const int nj = (k == D - 1) ? 2 : 4;           // the last knot only owns bins 4D-2 and 4D-1
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nj) {
                const float td = 0.125f + 0.25f * (float)j;
                const float v = fmaf(td, c1, __fmul_rn(1.f - td, c0));   // the generic kernel's contraction (SASS: FMUL + FFMA)
                e = __expf(v - m);
                s += e;
                acc = fmaf(e, od + (float)j, acc);
            }
        }
        od += 4.f;
        c0 = c1;
    }
    disp[((size_t)b * outH + y) * outW + x] = acc / s;
}

// disparity_regression on an explicit probability volume (GwcNet/submodule.py:23-27)
__global__ void disparity_regression_kernel(const float* __restrict__ prob, float* __restrict__ disp, int D,
                                            size_t plane, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t b = i / plane, p = i - b * plane;
    const float* s = prob + b * D * plane + p;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(__ldg(s + (size_t)d * plane), (float)d, acc);
    disp[i] = acc;
}

}  // namespace

extern "C" int stb_disparity_regression_f32(const float* prob, float* disp, int B, int D, long long plane,
                                            void* stream) {
    if (!prob || !disp || B <= 0 || D <= 0 || plane <= 0) return STB_E_BADARG;
    size_t total = (size_t)B * (size_t)plane;
    disparity_regression_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        prob, disp, D, (size_t)plane, total);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

static float linear_scale(int in_size, int out_size, int align) {
    if (align) return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    return (float)in_size / (float)out_size;
}

extern "C" int stb_upsample_softargmin_f32(const float* cost, float* disp, int B, int D, int H, int W, int outD,
                                           int outH, int outW, int align_corners, void* stream) {
    if (!cost || !disp || B <= 0 || D <= 0 || H <= 0 || W <= 0 || outD <= 0 || outH <= 0 || outW <= 0)
        return STB_E_BADARG;
    if (B > 65535 || outH > 65535) return STB_E_BADARG;
    size_t smem = (size_t)D * HEAD_THREADS * sizeof(float);
    if (smem > 200 * 1024) return STB_E_SMEM;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(upsample_softargmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(stb_ceil_div(outW, HEAD_THREADS), outH, B);
    // opt-in until it has been confirmed bit-identical on hardware (tests/test_lowp_model_gpu.py::test_head_x4_*)
    static const bool use_x4 = getenv("STB_HEAD_X4") == nullptr || atoi(getenv("STB_HEAD_X4")) != 0;     // 0 = generic kernel
    if (use_x4 && !align_corners && outD == 4 * D && D >= 2) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(upsample_softargmin_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        upsample_softargmin_x4_kernel<<<grid, HEAD_THREADS, smem, (cudaStream_t)stream>>>(
            cost, disp, D, H, W, outH, outW, linear_scale(H, outH, 0), linear_scale(W, outW, 0));
        STB_CHECK_LAUNCH();
        return STB_OK;
    }
    upsample_softargmin_kernel<<<grid, HEAD_THREADS, smem, (cudaStream_t)stream>>>(
        cost, disp, D, H, W, outD, outH, outW, linear_scale(D, outD, align_corners),
        linear_scale(H, outH, align_corners), linear_scale(W, outW, align_corners), align_corners);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
