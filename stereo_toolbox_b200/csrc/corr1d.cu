// 1-D all-pairs correlation, its 1x2 average pyramid and the multi-level 9-tap lookup.
//   CorrBlock1D            RAFTStereo/corr.py:110-156 (+ bilinear_sampler RAFTStereo/utils/utils.py:59-74)
//   Combined_Geo_Encoding_Volume   IGEVStereo/geometry.py:7-70
// The reference issues 4 grid_sample launches + cat + permute per GRU iteration; here one launch
// produces the [B, levels*(2r+1), H, W] tensor with stores coalesced along W.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// corr[b,h,w1,w2] = scale * sum_c f1[b,c,h,w1] * f2[b,c,h,w2]      (fp32 FMA, 64x64x16 tiles)
constexpr int CB = 64, CKK = 16;

__global__ void __launch_bounds__(256)
corr1d_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ corr,
              int C, int H, int W1, int W2, float scale) {
    __shared__ __align__(16) float As[CKK][CB + 4];
    __shared__ __align__(16) float Bs[CKK][CB + 4];
    const int bh = blockIdx.z, b = bh / H, h = bh - b * H;
    const int m0 = blockIdx.y * CB, n0 = blockIdx.x * CB;
    const size_t cstride1 = (size_t)H * W1, cstride2 = (size_t)H * W2;
    const float* a = f1 + (size_t)b * C * cstride1 + (size_t)h * W1;
    const float* bb = f2 + (size_t)b * C * cstride2 + (size_t)h * W2;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;    // 16 x 16 threads, 4x4 outputs each
    float acc[4][4] = {};
    for (int k0 = 0; k0 < C; k0 += CKK) {
        for (int i = threadIdx.x; i < CKK * CB; i += 256) {
            int k = i / CB, m = i - k * CB;
            As[k][m] = (k0 + k < C && m0 + m < W1) ? __ldg(a + (size_t)(k0 + k) * cstride1 + m0 + m) : 0.f;
            Bs[k][m] = (k0 + k < C && n0 + m < W2) ? __ldg(bb + (size_t)(k0 + k) * cstride2 + n0 + m) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CKK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* o = corr + (size_t)bh * W1 * W2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= W1) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < W2) o[(size_t)m * W2 + n] = acc[i][j] * scale;
        }
    }
}

__global__ void avgpool_last_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t total, int Ws, int Wd) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t r = i / Wd;
    int w = (int)(i - r * Wd);
    const float* s = src + r * Ws + 2 * w;
    dst[i] = 0.5f * (__ldg(s) + __ldg(s + 1));
}

// ---------------------------------------------------------------------------------------------
constexpr int MAX_LEVELS = 4;
struct LookupArgs {
    const float* pyr[MAX_LEVELS];
    const float* coords;
    float* out;
    long long coords_bstride;
    int B, H, W1, W2, levels, radius;
};

// zero-padded linear interpolation of row[0..W) at x (grid_sample, align_corners=True, zeros)
__device__ __forceinline__ void sample_taps(const float* __restrict__ row, int W, float x, int ntap,
                                            float* __restrict__ dst, size_t dst_stride) {
    const float xf = floorf(x);
    const float f = x - xf;
    const int i0 = (int)xf;
    float prev = ((unsigned)i0 < (unsigned)W) ? __ldg(row + i0) : 0.f;
    for (int k = 0; k < ntap; ++k) {
        const int i1 = i0 + k + 1;
        const float nxt = ((unsigned)i1 < (unsigned)W) ? __ldg(row + i1) : 0.f;
        dst[(size_t)k * dst_stride] = prev * (1.f - f) + nxt * f;
        prev = nxt;
    }
}

__global__ void __launch_bounds__(128)
corr1d_lookup_kernel(const __grid_constant__ LookupArgs a) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    if (w >= a.W1) return;
    const float cx = __ldg(a.coords + (size_t)b * a.coords_bstride + (size_t)h * a.W1 + w);
    const int ntap = 2 * a.radius + 1;
    const size_t plane = (size_t)a.H * a.W1;
    float* o = a.out + (size_t)b * a.levels * ntap * plane + (size_t)h * a.W1 + w;
    const size_t pix = ((size_t)b * a.H + h) * a.W1 + w;
    for (int l = 0; l < a.levels; ++l) {
        const int Wl = a.W2 >> l;
        const float x = cx / (float)(1 << l) - (float)a.radius;
        sample_taps(a.pyr[l] + pix * Wl, Wl, x, ntap, o + (size_t)l * ntap * plane, plane);
    }
}

struct GeoArgs {
    const float* geo[MAX_LEVELS];
    const float* corr[MAX_LEVELS];
    const float* disp; const float* coords; float* out;
    int B, H, W, C, D, W2, levels, radius;
};

__global__ void __launch_bounds__(128)
geo_lookup_kernel(const __grid_constant__ GeoArgs a) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    if (w >= a.W) return;
    const size_t plane = (size_t)a.H * a.W;
    const size_t pix = ((size_t)b * a.H + h) * a.W + w;
    const float dsp = __ldg(a.disp + pix), cx = __ldg(a.coords + pix);
    const int ntap = 2 * a.radius + 1;
    const int per_level = (a.C + 1) * ntap;
    float* o = a.out + (size_t)b * a.levels * per_level * plane + (size_t)h * a.W + w;
    for (int l = 0; l < a.levels; ++l) {
        const int Dl = a.D >> l, Wl = a.W2 >> l;
        const float s = (float)(1 << l);
        const float xg = dsp / s - (float)a.radius;
        float* ol = o + (size_t)l * per_level * plane;
        for (int c = 0; c < a.C; ++c)
            sample_taps(a.geo[l] + (pix * a.C + c) * Dl, Dl, xg, ntap, ol + (size_t)c * ntap * plane, plane);
        const float xc = cx / s - dsp / s - (float)a.radius;
        sample_taps(a.corr[l] + pix * Wl, Wl, xc, ntap, ol + (size_t)a.C * ntap * plane, plane);
    }
}

// [B][R][P] -> [B][P][R] tiled transpose (R = C*D rows, P = H*W pixels)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const float* s = src + (size_t)b * R * P;
    float* d = dst + (size_t)b * R * P;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int r = r0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && p < P) ? __ldg(s + (size_t)r * P + p) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int p = p0 + i, r = r0 + threadIdx.x;
        if (r < R && p < P) d[(size_t)p * R + r] = tile[threadIdx.x][i];
    }
}

}  // namespace

extern "C" int stb_corr1d_f32(const float* f1, const float* f2, float* corr, int B, int C, int H, int W1, int W2,
                              float scale, void* stream) {
    if (!f1 || !f2 || !corr || B <= 0 || C <= 0 || H <= 0 || W1 <= 0 || W2 <= 0) return STB_E_BADARG;
    if ((long long)B * H > 65535) return STB_E_BADARG;
    dim3 grid(stb_ceil_div(W2, CB), stb_ceil_div(W1, CB), B * H);
    corr1d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(f1, f2, corr, C, H, W1, W2, scale);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_avgpool_last_f32(const float* src, float* dst, long long rows, int Wsrc, void* stream) {
    if (!src || !dst || rows <= 0 || Wsrc < 2) return STB_E_BADARG;
    const int Wd = Wsrc / 2;
    size_t total = (size_t)rows * Wd;
    avgpool_last_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, total, Wsrc, Wd);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_corr1d_lookup_f32(const float* const* pyr, const float* coords_x, long long coords_bstride,
                                     float* out, int B, int H, int W1, int W2, int levels, int radius, void* stream) {
    if (!pyr || !coords_x || !out || B <= 0 || H <= 0 || W1 <= 0 || W2 <= 0) return STB_E_BADARG;
    if (levels < 1 || levels > MAX_LEVELS || radius < 0 || radius > 16 || (W2 >> (levels - 1)) < 1) return STB_E_BADARG;
    if (B > 65535 || H > 65535) return STB_E_BADARG;
    LookupArgs a;
    for (int l = 0; l < MAX_LEVELS; ++l) a.pyr[l] = l < levels ? pyr[l] : nullptr;
    a.coords = coords_x; a.out = out; a.coords_bstride = coords_bstride;
    a.B = B; a.H = H; a.W1 = W1; a.W2 = W2; a.levels = levels; a.radius = radius;
    dim3 grid(stb_ceil_div(W1, 128), H, B);
    corr1d_lookup_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_geo_lookup_f32(const float* const* geo, const float* const* corr, const float* disp,
                                  const float* coords_x, float* out, int B, int H, int W, int C, int D, int W2,
                                  int levels, int radius, void* stream) {
    if (!geo || !corr || !disp || !coords_x || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || D <= 0 || W2 <= 0)
        return STB_E_BADARG;
    if (levels < 1 || levels > MAX_LEVELS || radius < 0 || radius > 16) return STB_E_BADARG;
    if (B > 65535 || H > 65535) return STB_E_BADARG;
    GeoArgs a;
    for (int l = 0; l < MAX_LEVELS; ++l) {
        a.geo[l] = l < levels ? geo[l] : nullptr;
        a.corr[l] = l < levels ? corr[l] : nullptr;
    }
    a.disp = disp; a.coords = coords_x; a.out = out;
    a.B = B; a.H = H; a.W = W; a.C = C; a.D = D; a.W2 = W2; a.levels = levels; a.radius = radius;
    dim3 grid(stb_ceil_div(W, 128), H, B);
    geo_lookup_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_geo_permute_f32(const float* src, float* dst, int B, int C, int D, int H, int W, void* stream) {
    if (!src || !dst || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return STB_E_BADARG;
    const int R = C * D, P = H * W;
    dim3 grid(stb_ceil_div(P, 32), stb_ceil_div(R, 32), B);
    if (grid.y > 65535 || B > 65535) return STB_E_BADARG;
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, R, P);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
