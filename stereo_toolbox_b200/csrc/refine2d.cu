// PCWNet's full-resolution refinement inputs (reference: models/PCWNet/submodule.py:122-152
// `warp`, :104-120 `build_corrleation_volume(..., num_groups=1)`, used by PCWNet/pcwnet.py:491-506):
//
//   stb_warp_disp_f32        right features bilinearly resampled at x - disp with the reference's own grid arithmetic (grid
//                            normalised with W-1 / H-1 but sampled with align_corners=False, so BOTH axes are slightly rescaled),
//                            zero padding, multiplied by the validity mask (sum of in-bounds bilinear weights >= 0.999)
//   stb_corr_volume_1d_f32   [B, 2*md+1, H, W]: plane i+md = mean_c L[x] * R[x-i] for i >= 0 (x >= i); for i < 0 the
//                            reference's slices pair the FIRST k = -i columns of the left with the LAST k columns of the right,
//                            written to columns [0, k) -- reproduced as is; everything else is zero
//
// The reference runs the second as 2*md+1 = 49 slice-multiply-mean-assign sequences over [B,32,H,W] tensors at full
// resolution; here the left / warped-right row tiles are staged in shared memory once and all 49 planes written in one pass.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
warp_disp_kernel(const float* __restrict__ x, const float* __restrict__ disp, float* __restrict__ out, int C, int H, int W,
                 long long npix) {
    const float inv_w = __frcp_rn((float)max(W - 1, 1)), inv_h = __frcp_rn((float)max(H - 1, 1));
    const size_t plane = (size_t)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const long long t = i / W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        // grid (reference arithmetic, fp32): gx = 2*(x - d)/(W-1) - 1, gy = 2*y/(H-1) - 1, as torch evaluates it -- one rounding
        // per elementwise op, the division by a Python scalar as a multiplication with its fp32 reciprocal -- so that the
        // sampling positions (ulp 1.2e-4 at W = 1248) agree bit for bit; then unnormalise with align_corners=False
        const float gx = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, __fsub_rn((float)w, __ldg(disp + i))), inv_w), 1.0f);
        const float gy = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)h), inv_h), 1.0f);
        const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
        const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;     // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
        const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
        float m = 0.f;
        if (vy0 && vx0) m += w00;
        if (vy0 && vx1) m += w01;
        if (vy1 && vx0) m += w10;
        if (vy1 && vx1) m += w11;
        const float mask = m >= 0.999f ? 1.f : 0.f;
        const float* xb = x + (size_t)b * C * plane;
        float* ob = out + (size_t)b * C * plane + (size_t)h * W + w;
        for (int c = 0; c < C; ++c) {
            const float* xp = xb + (size_t)c * plane;
            float v = 0.f;
            if (vy0 && vx0) v += __ldg(xp + (size_t)y0 * W + x0) * w00;
            if (vy0 && vx1) v += __ldg(xp + (size_t)y0 * W + x1) * w01;
            if (vy1 && vx0) v += __ldg(xp + (size_t)y1 * W + x0) * w10;
            if (vy1 && vx1) v += __ldg(xp + (size_t)y1 * W + x1) * w11;
            ob[(size_t)c * plane] = v * mask;
        }
    }
}

constexpr int CV_TW = 128;      // x positions per CTA
constexpr int CV_MAXC = 32;     // channels staged per pass

// grid (w tiles, H, B), 128 threads: thread = one x; L / R row tiles staged per 32-channel pass.  MD compile-time: the md + 1
// running sums of a thread stay in registers.
template <int MD>
__global__ void __launch_bounds__(CV_TW)
corr_volume_1d_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ vol, int C, int H, int W) {
    extern __shared__ float sm[];
    constexpr int md = MD;
    const int halo = md;
    float* Ls = sm;                                  // [CV_MAXC][CV_TW]
    float* Rs = sm + CV_MAXC * CV_TW;                // [CV_MAXC][CV_TW + halo]   column j <-> x = x0 - halo + j
    const int x0 = blockIdx.x * CV_TW, h = blockIdx.y, b = blockIdx.z;
    const int x = x0 + threadIdx.x;
    const size_t plane = (size_t)H * W;
    const int RP = CV_TW + halo;
    float acc[MD + 1];                               // planes i = 0..md
#pragma unroll
    for (int i = 0; i <= MD; ++i) acc[i] = 0.f;
    for (int c0 = 0; c0 < C; c0 += CV_MAXC) {
        const int cn = min(CV_MAXC, C - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn * CV_TW; i += CV_TW) {
            const int c = i / CV_TW, j = i - c * CV_TW;
            const int xx = x0 + j;
            Ls[c * CV_TW + j] = xx < W ? __ldg(L + ((size_t)b * C + c0 + c) * plane + (size_t)h * W + xx) : 0.f;
        }
        for (int i = threadIdx.x; i < cn * RP; i += CV_TW) {
            const int c = i / RP, j = i - c * RP;
            const int xx = x0 - halo + j;
            Rs[c * RP + j] = (xx >= 0 && xx < W) ? __ldg(R + ((size_t)b * C + c0 + c) * plane + (size_t)h * W + xx) : 0.f;
        }
        __syncthreads();
        for (int c = 0; c < cn; ++c) {
            const float l = Ls[c * CV_TW + threadIdx.x];
            const float* rr = Rs + c * RP + threadIdx.x + halo;       // rr[-i] = R[x - i]
#pragma unroll
            for (int i = 0; i <= MD; ++i) acc[i] = fmaf(l, rr[-i], acc[i]);
        }
    }
    if (x >= W) return;
    const float inv = 1.f / (float)C;
    float* vb = vol + ((size_t)b * (2 * md + 1)) * plane + (size_t)h * W + x;
#pragma unroll
    for (int i = 0; i <= MD; ++i) vb[(size_t)(md + i) * plane] = x >= i ? acc[i] * inv : 0.f;
    // i < 0 (k = -i): columns [0, k) hold mean_c L[x] * R[W - k + x]; rare (x < md), read straight from global memory
    for (int k = 1; k <= md; ++k) {
        float v = 0.f;
        if (x < k && W - k + x >= 0) {
            float s = 0.f;
            for (int c = 0; c < C; ++c)
                s = fmaf(__ldg(L + ((size_t)b * C + c) * plane + (size_t)h * W + x),
                         __ldg(R + ((size_t)b * C + c) * plane + (size_t)h * W + (W - k + x)), s);
            v = s * inv;
        }
        vb[(size_t)(md - k) * plane] = v;
    }
}

}  // namespace

extern "C" int stb_warp_disp_f32(const float* x, const float* disp, float* out, int B, int C, int H, int W, void* stream) {
    if (!x || !disp || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return STB_E_BADARG;
    const long long npix = (long long)B * H * W;
    long long g = (npix + 255) / 256;
    if (g > 148LL * 16) g = 148LL * 16;
    warp_disp_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, disp, out, C, H, W, npix);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_corr_volume_1d_f32(const float* left, const float* right, float* vol, int B, int C, int H, int W, int maxdisp,
                                      void* stream) {
    if (!left || !right || !vol || B <= 0 || C <= 0 || H <= 0 || W <= 0) return STB_E_BADARG;
    if (maxdisp >= W) return STB_E_UNSUPPORTED;
    const size_t smem = sizeof(float) * (size_t)CV_MAXC * (2 * CV_TW + maxdisp);
    dim3 grid((unsigned)stb_ceil_div(W, CV_TW), (unsigned)H, (unsigned)B);
    if (maxdisp == 24) corr_volume_1d_kernel<24><<<grid, CV_TW, smem, (cudaStream_t)stream>>>(left, right, vol, C, H, W);   // PCWNet's refinement
    else if (maxdisp == 8) corr_volume_1d_kernel<8><<<grid, CV_TW, smem, (cudaStream_t)stream>>>(left, right, vol, C, H, W);
    else if (maxdisp == 4) corr_volume_1d_kernel<4><<<grid, CV_TW, smem, (cudaStream_t)stream>>>(left, right, vol, C, H, W);
    else return STB_E_UNSUPPORTED;
    STB_CHECK_LAUNCH();
    return STB_OK;
}
