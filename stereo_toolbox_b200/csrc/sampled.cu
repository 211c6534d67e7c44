// CFNet sampled cost volume (cascade stages at 1/4 and 1/2 resolution), one launch instead of the reference's
// expand + gather + multiply + mean + cat chain (CFNet/cfnet.py:472-496 cost_volume_generator x 2 + cat :545-550,
// CFNet/submodule.py:302-349 SpatialTransformer, :162-168 groupwise_correlation_4D), which materialises a
// [B, 320, S, H, W] gather (613 MB per pair at 1/4 resolution) before reducing it to 40 groups.
//
//   vol[b, g,        s, h, w] = valid * mean_c  L_gw[b, g*k+c, h, w] * R_gw[b, g*k+c, h, x]      g < G
//   vol[b, G+c,      s, h, w] = L_cat[b, c, h, w]                                                 (left: broadcast over s)
//   vol[b, G+Cc+c,   s, h, w] = valid * R_cat[b, c, h, x]
//   vol[b, G+2Cc,    s, h, w] = samples[b, s, h, w]
//   with x = clamp(w - samples[b,s,h,w], 0, W-1) and valid = !(w - sample < 0 || w - sample > W-1).
//
// HBM-bound (the same staging as gwc_volume_kernel): one CTA per (b, output channel, h) keeps its 2 x cpg feature rows
// in shared memory, walks the S x W slab with w fastest (coalesced sample reads and stores); the right-row read is a
// shared-memory gather.  Written after the round-1 GPU budget was spent: opt-in (STB_CFNET_SAMPLED=1), not yet run.
#include "common.cuh"

namespace {

constexpr int SV_THREADS = 128;

__global__ void __launch_bounds__(SV_THREADS)
sampled_volume_kernel(const float* __restrict__ gw_l, const float* __restrict__ gw_r, const float* __restrict__ cat_l,
                      const float* __restrict__ cat_r, const float* __restrict__ samples, float* __restrict__ vol,
                      int Cg, int G, int Cc, int S, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    const int h = blockIdx.x, o = blockIdx.y, b = blockIdx.z;
    const int Ct = G + 2 * Cc + 1;
    const size_t plane = (size_t)H * W;
    float* out = vol + (((size_t)b * Ct + o) * S) * plane + (size_t)h * W;             // + s*plane + w
    const float* smp = samples + ((size_t)b * S) * plane + (size_t)h * W;              // + s*plane + w
    const float wmax = (float)(W - 1);
    const int n = S * W;
    if (o < G) {
        const int cpg = Cg / G;
        float* Ls = smem;                 // [cpg][W]
        float* Rs = smem + cpg * W;       // [cpg][W]
        const float* lsrc = gw_l + ((size_t)b * Cg + (size_t)o * cpg) * plane + (size_t)h * W;
        const float* rsrc = gw_r + ((size_t)b * Cg + (size_t)o * cpg) * plane + (size_t)h * W;
        for (int i = threadIdx.x; i < cpg * W; i += SV_THREADS) {
            const int c = i / W, w = i - c * W;
            Ls[i] = __ldg(lsrc + (size_t)c * plane + w);
            Rs[i] = __ldg(rsrc + (size_t)c * plane + w);
        }
        __syncthreads();
        const float inv = 1.f / (float)cpg;
        for (int i = threadIdx.x; i < n; i += SV_THREADS) {
            const int s = i / W, w = i - s * W;
            const float coord = (float)w - __ldg(smp + (size_t)s * plane + w);
            const bool valid = !(coord < 0.f) && !(coord > wmax);
            const int x = (int)fminf(fmaxf(coord, 0.f), wmax);
            float acc = 0.f;
            for (int c = 0; c < cpg; ++c) acc = fmaf(Ls[c * W + w], Rs[c * W + x], acc);
            __stcs(out + (size_t)s * plane + w, valid ? acc * inv : 0.f);
        }
    } else if (o < G + Cc) {
        const float* lsrc = cat_l + ((size_t)b * Cc + (o - G)) * plane + (size_t)h * W;
        for (int i = threadIdx.x; i < n; i += SV_THREADS) {
            const int s = i / W, w = i - s * W;
            __stcs(out + (size_t)s * plane + w, __ldg(lsrc + w));
        }
    } else if (o < G + 2 * Cc) {
        float* Rs = smem;                 // [W]
        const float* rsrc = cat_r + ((size_t)b * Cc + (o - G - Cc)) * plane + (size_t)h * W;
        for (int w = threadIdx.x; w < W; w += SV_THREADS) Rs[w] = __ldg(rsrc + w);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SV_THREADS) {
            const int s = i / W, w = i - s * W;
            const float coord = (float)w - __ldg(smp + (size_t)s * plane + w);
            const bool valid = !(coord < 0.f) && !(coord > wmax);
            const int x = (int)fminf(fmaxf(coord, 0.f), wmax);
            __stcs(out + (size_t)s * plane + w, valid ? Rs[x] : 0.f);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += SV_THREADS) {
            const int s = i / W, w = i - s * W;
            __stcs(out + (size_t)s * plane + w, __ldg(smp + (size_t)s * plane + w));
        }
    }
}

}  // namespace

extern "C" int stb_sampled_volume_f32(const float* gw_left, const float* gw_right, const float* cat_left,
                                      const float* cat_right, const float* samples, float* vol, int B, int Cg, int G,
                                      int Cc, int S, int H, int W, void* stream) {
    if (!gw_left || !gw_right || !samples || !vol || B <= 0 || Cg <= 0 || G <= 0 || Cc < 0 || S <= 0 || H <= 0 || W <= 0)
        return STB_E_BADARG;
    if (Cg % G) return STB_E_BADARG;                       // CFNet/submodule.py:164: assert C % num_groups == 0
    if (Cc > 0 && (!cat_left || !cat_right)) return STB_E_BADARG;
    if (B > 65535 || G + 2 * Cc + 1 > 65535) return STB_E_BADARG;
    const int cpg = Cg / G;
    const size_t smem = (size_t)2 * cpg * W * sizeof(float);
    if (smem > 200 * 1024) return STB_E_SMEM;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(sampled_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(H, G + 2 * Cc + 1, B);
    sampled_volume_kernel<<<grid, SV_THREADS, smem, (cudaStream_t)stream>>>(gw_left, gw_right, cat_left, cat_right, samples,
                                                                            vol, Cg, G, Cc, S, H, W);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
