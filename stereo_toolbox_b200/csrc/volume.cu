// Cost-volume builders (fp32 NCDHW), one launch instead of the reference's 48 x (mul, mean, copy).
//   gwc   : GwcNet/submodule.py:44-63      concat: GwcNet/submodule.py:30-41, ACVNet/submodule.py:180-191
// HBM-bound: each (b, group, h) CTA stages its 2 x cpg feature rows in shared memory once and
// streams D output rows with 16-byte stores; the right row is kept as a sliding register window
// so that one shared-memory load feeds 4 outputs.
#include "common.cuh"

namespace {

constexpr int GWC_THREADS = 128;

// grid: (H, G, B); dynamic smem: cpg*(W) + cpg*(W+Dpad) floats
template <int CPG>
__global__ void __launch_bounds__(GWC_THREADS)
gwc_volume_kernel(const float* __restrict__ left, const float* __restrict__ right, float* __restrict__ vol,
                  int C, int H, int W, int D, int G, int c_total, int c_off) {
    extern __shared__ __align__(16) float smem[];
    const int h = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int Wp = (W + 3) & ~3;            // padded row pitch for L
    const int Rl = Wp + D;                  // logical R row: D leading zeros (the w<d region) + the feature row
    // R rows are stored SKEWED: logical index x lives at x + (x >> 5).  A lane owns 4 consecutive w, so the lanes of a
    // warp read logical indices 16 bytes apart: unskewed that is a 4-way bank conflict on every sliding-window load
    // (8 per depth step -- the shared-memory pipe, not HBM, bounded the kernel at 0.37 of the roofline).
    const int Rp = Rl + (Rl >> 5) + 1;
    float* Ls = smem;                       // [CPG][Wp]
    float* Rs = smem + CPG * Wp;            // [CPG][Rp] skewed
#define RSK(x) ((x) + ((x) >> 5))
    const size_t plane = (size_t)H * W;
    const float* lsrc = left + ((size_t)b * C + (size_t)g * CPG) * plane + (size_t)h * W;
    const float* rsrc = right + ((size_t)b * C + (size_t)g * CPG) * plane + (size_t)h * W;
    for (int i = threadIdx.x; i < CPG * Wp; i += GWC_THREADS) {
        int c = i / Wp, w = i - c * Wp;
        Ls[i] = (w < W) ? __ldg(lsrc + (size_t)c * plane + w) : 0.f;
    }
    for (int i = threadIdx.x; i < CPG * Rl; i += GWC_THREADS) {
        int c = i / Rl, x = i - c * Rl, w = x - D;
        Rs[c * Rp + RSK(x)] = (w >= 0 && w < W) ? __ldg(rsrc + (size_t)c * plane + w) : 0.f;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = GWC_THREADS / 32;
    const int dper = (D + NW - 1) / NW;
    const int d_lo = warp * dper, d_hi = min(D, d_lo + dper);
    const float inv = 1.f / (float)CPG;
    const bool vec_ok = (W & 3) == 0;
    float* obase = vol + (((size_t)b * c_total + c_off + g) * D) * plane + (size_t)h * W;
    const int nquads = Wp >> 2;
    for (int q = lane; q < nquads; q += 32) {
        const int w0 = q * 4;
        float l[CPG][4], r[CPG][4];
#pragma unroll
        for (int c = 0; c < CPG; ++c) {
            float4 v = *reinterpret_cast<const float4*>(Ls + c * Wp + w0);
            l[c][0] = v.x; l[c][1] = v.y; l[c][2] = v.z; l[c][3] = v.w;
        }
        if (d_lo < d_hi) {
#pragma unroll
            for (int c = 0; c < CPG; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) r[c][j] = Rs[c * Rp + RSK(D + w0 + j - d_lo)];
        }
        for (int d = d_lo; d < d_hi; ++d) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CPG; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(l[c][j], r[c][j], acc[j]);
            float* o = obase + (size_t)d * plane + w0;
            if (vec_ok) {
                __stcs(reinterpret_cast<float4*>(o), make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (w0 + j < W) o[j] = acc[j] * inv;
            }
            // slide the window: next d needs R[w0+j-(d+1)]
            if (d + 1 < d_hi) {
#pragma unroll
                for (int c = 0; c < CPG; ++c) {
                    r[c][3] = r[c][2]; r[c][2] = r[c][1]; r[c][1] = r[c][0];
                    r[c][0] = Rs[c * Rp + RSK(D + w0 - (d + 1))];
                }
            }
        }
    }
}

#undef RSK

// concat volume: grid (H, C, B); each CTA copies one left row and one (shifted) right row into D planes.
__global__ void __launch_bounds__(128)
concat_volume_kernel(const float* __restrict__ left, const float* __restrict__ right,
                     const float* __restrict__ att, float* __restrict__ vol,
                     int C, int H, int W, int D, int mask_left, int c_total, int c_off) {
    extern __shared__ __align__(16) float smem[];
    const int h = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    const int Wp = (W + 3) & ~3;
    float* Ls = smem;            // [Wp]
    float* Rs = smem + Wp;       // [D + Wp], D leading zeros
    const size_t plane = (size_t)H * W;
    const float* lsrc = left + ((size_t)b * C + c) * plane + (size_t)h * W;
    const float* rsrc = right + ((size_t)b * C + c) * plane + (size_t)h * W;
    for (int w = threadIdx.x; w < Wp; w += blockDim.x) Ls[w] = w < W ? __ldg(lsrc + w) : 0.f;
    for (int i = threadIdx.x; i < D + Wp; i += blockDim.x) {
        int w = i - D;
        Rs[i] = (w >= 0 && w < W) ? __ldg(rsrc + w) : 0.f;
    }
    __syncthreads();
    float* ol = vol + (((size_t)b * c_total + c_off + c) * D) * plane + (size_t)h * W;
    float* orr = vol + (((size_t)b * c_total + c_off + C + c) * D) * plane + (size_t)h * W;
    const bool vec_ok = (W & 3) == 0;
    const int nquads = Wp >> 2;
    const float* attb = att ? att + (size_t)b * D * plane + (size_t)h * W : nullptr;
    for (int i = threadIdx.x; i < D * nquads; i += blockDim.x) {
        const int d = i / nquads, w0 = (i - d * nquads) * 4;
        float lv[4], rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = w0 + j;
            lv[j] = (!mask_left || w >= d) ? Ls[w] : 0.f;
            rv[j] = Rs[D + w - d];
        }
        if (attb) {
            // ACVNet/acv.py:196: volume *= softmax_d(att); attb already holds the probabilities
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = w0 + j;
                if (w < W) {
                    const float p = __ldg(attb + (size_t)d * plane + w);
                    lv[j] *= p; rv[j] *= p;
                }
            }
        }
        float* pl = ol + (size_t)d * plane + w0;
        float* pr = orr + (size_t)d * plane + w0;
        if (vec_ok) {
            __stcs(reinterpret_cast<float4*>(pl), make_float4(lv[0], lv[1], lv[2], lv[3]));
            __stcs(reinterpret_cast<float4*>(pr), make_float4(rv[0], rv[1], rv[2], rv[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (w0 + j < W) { pl[j] = lv[j]; pr[j] = rv[j]; }
        }
    }
}

template <int CPG>
int launch_gwc(const float* l, const float* r, float* v, int B, int C, int H, int W, int D, int G,
               int c_total, int c_off, cudaStream_t st) {
    const int Wp = (W + 3) & ~3;
    const int Rl = Wp + D;
    size_t smem = (size_t)(CPG * Wp + CPG * (Rl + (Rl >> 5) + 1)) * sizeof(float);       // skewed R rows, see the kernel
    if (smem > 200 * 1024) return STB_E_SMEM;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(gwc_volume_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(H, G, B);
    gwc_volume_kernel<CPG><<<grid, GWC_THREADS, smem, st>>>(l, r, v, C, H, W, D, G, c_total, c_off);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

// softmax over the D axis of [B,D,HW] (one thread per (b, hw) column)
__global__ void softmax_d_kernel(const float* __restrict__ x, float* __restrict__ y, int D, size_t plane, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t b = i / plane, p = i - b * plane;
    const float* xs = x + b * D * plane + p;
    float* ys = y + b * D * plane + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(xs + (size_t)d * plane));
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(__ldg(xs + (size_t)d * plane) - m);
    const float inv = 1.f / s;
    for (int d = 0; d < D; ++d) ys[(size_t)d * plane] = expf(__ldg(xs + (size_t)d * plane) - m) * inv;
}

}  // namespace

extern "C" int stb_softmax_d_f32(const float* x, float* y, int B, int D, long long plane, void* stream) {
    if (!x || !y || B <= 0 || D <= 0 || plane <= 0) return STB_E_BADARG;
    size_t total = (size_t)B * (size_t)plane;
    softmax_d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, D, (size_t)plane, total);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_gwc_volume_f32(const float* left, const float* right, float* vol, int B, int C, int H, int W,
                                  int D, int G, int c_total, int c_off, void* stream) {
    if (!left || !right || !vol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0 || G <= 0) return STB_E_BADARG;
    if (C % G) return STB_E_BADARG;   // reference: assert C % num_groups == 0 (GwcNet/submodule.py:46)
    if (c_off < 0 || c_off + G > c_total || B > 65535 || G > 65535) return STB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C / G) {
        case 1: return launch_gwc<1>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 2: return launch_gwc<2>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 3: return launch_gwc<3>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 4: return launch_gwc<4>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 6: return launch_gwc<6>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 8: return launch_gwc<8>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 12: return launch_gwc<12>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        case 16: return launch_gwc<16>(left, right, vol, B, C, H, W, D, G, c_total, c_off, st);
        default: return STB_E_UNSUPPORTED;
    }
}

extern "C" int stb_concat_volume_f32(const float* left, const float* right, const float* att, float* vol, int B,
                                     int C, int H, int W, int D, int mask_left, int c_total, int c_off,
                                     void* stream) {
    if (!left || !right || !vol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return STB_E_BADARG;
    if (c_off < 0 || c_off + 2 * C > c_total || B > 65535 || C > 65535) return STB_E_BADARG;
    const int Wp = (W + 3) & ~3;
    size_t smem = (size_t)(2 * Wp + D) * sizeof(float);
    if (smem > 48 * 1024) return STB_E_SMEM;
    dim3 grid(H, C, B);
    concat_volume_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(left, right, att, vol, C, H, W, D, mask_left,
                                                                    c_total, c_off);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
