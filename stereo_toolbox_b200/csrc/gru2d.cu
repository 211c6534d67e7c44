// Elementwise / resampling companions of the iterative models' update block on the tensor-core 2-D conv path
// (reference: models/RAFTStereo/update.py:29-44 ConvGRU, :91-95 pool2x / interp, :115-138
// BasicMultiUpdateBlock.forward).  The ConvGRU's convolutions run on csrc/conv3d_umma.cu with channels-last operand-split
// fp16 activations ("fp16x2": every value = fp16 hi + fp16 lo, interleaved per 16 channels); these kernels do what lies
// between the convolutions on the SAME storage, so the recurrent state never leaves that layout:
//   stb_gru_rh_split      r * h                    (r = channels [C, 2C) of the fused sigmoid(convz | convr) output)
//   stb_gru_blend_split   (1 - z) * h + z * q      (z = channels [0, C) of the same tensor)
//   stb_pool2x_split      F.avg_pool2d(x, 3, stride=2, padding=1)                       (count_include_pad: always / 9)
//   stb_interp_split      F.interpolate(x, (Ho, Wo), mode="bilinear", align_corners=True)
// All are HBM/L2-bound single passes; a thread owns 8 consecutive logical channels of one pixel (two 16-byte accesses per
// tensor: the hi and the lo halves).
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

// 8 logical channels [8*c8, 8*c8 + 8) of a split row (2*C halves): 16-channel block b = c8 >> 1 occupies halves
// [32*b, 32*b + 32) = [hi 0..15 | lo 0..15]; the 8-channel half h = c8 & 1 sits at +8*h (hi) and +16 + 8*h (lo).
__device__ __forceinline__ void load8_split(const uint16_t* row, int c8, float (&v)[8]) {
    const uint16_t* p = row + (c8 >> 1) * 32 + (c8 & 1) * 8;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + 16));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
        v[2 * j] = a.x + b.x;          // hi + lo is exact in fp32
        v[2 * j + 1] = a.y + b.y;
    }
}
__device__ __forceinline__ void store8_split(uint16_t* row, int c8, const float (&v)[8]) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        hw[j] = *reinterpret_cast<const uint32_t*>(&h);
        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    uint16_t* p = row + (c8 >> 1) * 32 + (c8 & 1) * 8;
    *reinterpret_cast<uint4*>(p) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(p + 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// mode 0: out = r * h ; mode 1: out = (1 - z) * h + z * q.   zr rows hold 2C logical channels (z | r), h / q / out rows C.
__global__ void gru_gate_kernel(const uint16_t* __restrict__ zr, const uint16_t* __restrict__ h, const uint16_t* __restrict__ q,
                                uint16_t* __restrict__ out, long long npix, int C, int mode) {
    const int chunks = C >> 3;
    const long long total = npix * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / chunks;
        const int c8 = (int)(i - pix * chunks);
        float hv[8], gv[8], o[8];
        load8_split(h + pix * (2 * (long long)C), c8, hv);
        load8_split(zr + pix * (4 * (long long)C), mode == 0 ? chunks + c8 : c8, gv);
        if (mode == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = gv[k] * hv[k];
        } else {
            float qv[8];
            load8_split(q + pix * (2 * (long long)C), c8, qv);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (1.f - gv[k]) * hv[k] + gv[k] * qv[k];
        }
        store8_split(out + pix * (2 * (long long)C), c8, o);
    }
}

// avg_pool2d(3, stride 2, padding 1), zero padding counted: x [N,H,W,C] -> out [N,Ho,Wo,C]
__global__ void pool2x_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ out, int N, int H, int W, int Ho, int Wo,
                              int C) {
    const int chunks = C >> 3;
    const long long total = (long long)N * Ho * Wo * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % chunks);
        long long t = i / chunks;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hh = 2 * ho - 1 + dy;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ww = 2 * wo - 1 + dx;
                if (ww < 0 || ww >= W) continue;
                float v[8];
                load8_split(x + (((long long)n * H + hh) * W + ww) * (2 * (long long)C), c8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += v[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = acc[k] / 9.f;
        store8_split(out + (((long long)n * Ho + ho) * Wo + wo) * (2 * (long long)C), c8, acc);
    }
}

// bilinear, align_corners = True (torch upsample_bilinear2d: src = dst * (in - 1) / (out - 1), lambda weights in fp32)
__global__ void interp_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ out, int N, int H, int W, int Ho, int Wo,
                              int C, float sh, float sw) {
    const int chunks = C >> 3;
    const long long total = (long long)N * Ho * Wo * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % chunks);
        long long t = i / chunks;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float fh = sh * (float)ho, fw = sw * (float)wo;
        const int h0 = min((int)fh, H - 1), w0 = min((int)fw, W - 1);
        const int h1 = min(h0 + 1, H - 1), w1 = min(w0 + 1, W - 1);
        const float lh1 = fh - (float)h0, lw1 = fw - (float)w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
        const long long rowb = (long long)n * H;
        float a[8], b[8], c[8], d[8], o[8];
        load8_split(x + ((rowb + h0) * W + w0) * (2 * (long long)C), c8, a);
        load8_split(x + ((rowb + h0) * W + w1) * (2 * (long long)C), c8, b);
        load8_split(x + ((rowb + h1) * W + w0) * (2 * (long long)C), c8, c);
        load8_split(x + ((rowb + h1) * W + w1) * (2 * (long long)C), c8, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = lh0 * (lw0 * a[k] + lw1 * b[k]) + lh1 * (lw0 * c[k] + lw1 * d[k]);
        store8_split(out + (((long long)n * Ho + ho) * Wo + wo) * (2 * (long long)C), c8, o);
    }
}

inline unsigned grid_for(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int stb_gru_rh_split(const void* zr, const void* h, void* out, long long npix, int C, void* stream) {
    if (!zr || !h || !out || npix <= 0 || C <= 0 || C % 16) return STB_E_BADARG;
    gru_gate_kernel<<<grid_for(npix * (C >> 3), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)zr, (const uint16_t*)h, nullptr, (uint16_t*)out, npix, C, 0);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_gru_blend_split(const void* zr, const void* h, const void* q, void* out, long long npix, int C, void* stream) {
    if (!zr || !h || !q || !out || npix <= 0 || C <= 0 || C % 16) return STB_E_BADARG;
    gru_gate_kernel<<<grid_for(npix * (C >> 3), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)zr, (const uint16_t*)h, (const uint16_t*)q, (uint16_t*)out, npix, C, 1);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_pool2x_split(const void* x, void* out, int N, int H, int W, int C, void* stream) {
    if (!x || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 16) return STB_E_BADARG;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    pool2x_kernel<<<grid_for((long long)N * Ho * Wo * (C >> 3), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)x, (uint16_t*)out, N, H, W, Ho, Wo, C);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_interp_split(const void* x, void* out, int N, int H, int W, int Ho, int Wo, int C, void* stream) {
    if (!x || !out || N <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || C <= 0 || C % 16) return STB_E_BADARG;
    const float sh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    interp_kernel<<<grid_for((long long)N * Ho * Wo * (C >> 3), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)x, (uint16_t*)out, N, H, W, Ho, Wo, C, sh, sw);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
