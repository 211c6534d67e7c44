// Elementwise companions of the conv family (HBM-bound, one pass):
//   * FeatureAtt gate of IGEV (IGEVStereo/submodule.py:228-241): cv * sigmoid(feat_att)[:, :, None] -- the 2-D gate
//     logits [B,C,H,W] are broadcast over the disparity axis of the cost volume
//   * disparity_variance of CFNet (CFNet/submodule.py:127-133): sum_d p_d * (d - disp)^2
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// x, out [B,C,D,H,W] fp32; gate [B,C,H,W] fp32 logits.  blockIdx.x = (b*C + c)*D + d (one volume plane), blockIdx.y strides
// over the plane: no per-element index decode, float4 accesses when the plane size allows.
__global__ void gate_f32_kernel(const float* __restrict__ x, const float* __restrict__ gate, float* __restrict__ out,
                                int D, long long hw) {
    const long long plane = blockIdx.x;
    const long long bc = plane / D;
    const float* xp = x + plane * hw;
    float* op = out + plane * hw;
    const float* gp = gate + bc * hw;
    if ((hw & 3) == 0) {
        const long long n4 = hw >> 2;
        for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.y * blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xp) + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(gp) + i);
            reinterpret_cast<float4*>(op)[i] = make_float4(v.x * sigmoidf_(g.x), v.y * sigmoidf_(g.y), v.z * sigmoidf_(g.z),
                                                           v.w * sigmoidf_(g.w));
        }
    } else {
        for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.y * blockDim.x)
            op[i] = xp[i] * sigmoidf_(__ldg(gp + i));
    }
}

// x, out [B,D,H,W,Cpad] 16-bit; gate [B,C,H,W] fp32 logits; channels >= C are copied (they are zero padding).
__global__ void gate_cl16_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gate, uint16_t* __restrict__ out,
                                 int C, int Cpad, int D, int H, int W, long long nvox, int f16) {
    const int chunks = Cpad >> 3;
    const long long total = nvox * chunks;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long vox = i / chunks;
        const int c0 = (int)(i - vox * chunks) * 8;
        const long long p = vox % hw;
        const long long b = vox / (hw * D);
        const uint4 v = *reinterpret_cast<const uint4*>(x + vox * Cpad + c0);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f;
            if (f16) f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
            else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
            const int c = c0 + 2 * j;
            if (c < C) f.x *= sigmoidf_(__ldg(gate + (b * C + c) * hw + p));
            if (c + 1 < C) f.y *= sigmoidf_(__ldg(gate + (b * C + c + 1) * hw + p));
            if (f16) { __half2 h = __floats2half2_rn(f.x, f.y); w[j] = *reinterpret_cast<uint32_t*>(&h); }
            else { __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y); w[j] = *reinterpret_cast<uint32_t*>(&h); }
        }
        *reinterpret_cast<uint4*>(out + vox * Cpad + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Operand-split storage ("fp16x2": [B,D,H,W,2*Cpad] halves, (hi, lo) interleaved per 16 channels): a thread owns 8 logical
// channels of a voxel -- the hi and the lo 16-byte pieces of its 16-channel block -- joins them, gates in fp32, splits again.
__global__ void gate_split_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gate, uint16_t* __restrict__ out,
                                  int C, int Cpad, int D, int H, int W, long long nvox) {
    const int chunks = Cpad >> 3;
    const long long total = nvox * chunks;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long vox = i / chunks;
        const int c8 = (int)(i - vox * chunks), c0 = c8 * 8;
        const long long p = vox % hw;
        const long long b = vox / (hw * D);
        const size_t off = (size_t)vox * (2 * (size_t)Cpad) + (c8 >> 1) * 32 + (c8 & 1) * 8;
        const uint4 hq = *reinterpret_cast<const uint4*>(x + off), lq = *reinterpret_cast<const uint4*>(x + off + 16);
        const uint32_t hw_[4] = {hq.x, hq.y, hq.z, hq.w}, lw_[4] = {lq.x, lq.y, lq.z, lq.w};
        uint32_t ho[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw_[j]));
            const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lw_[j]));
            float2 f = make_float2(a.x + l.x, a.y + l.y);
            const int c = c0 + 2 * j;
            if (c < C) f.x *= sigmoidf_(__ldg(gate + (b * C + c) * hw + p));
            if (c + 1 < C) f.y *= sigmoidf_(__ldg(gate + (b * C + c + 1) * hw + p));
            const __half2 h = __floats2half2_rn(f.x, f.y);
            const float2 hf = __half22float2(h);
            const __half2 r = __floats2half2_rn(f.x - hf.x, f.y - hf.y);
            ho[j] = *reinterpret_cast<const uint32_t*>(&h);
            lo[j] = *reinterpret_cast<const uint32_t*>(&r);
        }
        *reinterpret_cast<uint4*>(out + off) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
        *reinterpret_cast<uint4*>(out + off + 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// prob [B,D,plane], disp [B,plane] -> var [B,plane]
__global__ void variance_kernel(const float* __restrict__ prob, const float* __restrict__ disp, float* __restrict__ var,
                                int D, long long plane, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, p = i - b * plane;
        const float m = disp[i];
        const float* src = prob + b * D * plane + p;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) {
            const float t = (float)d - m;
            acc = fmaf(__ldg(src + (long long)d * plane), t * t, acc);
        }
        var[i] = acc;
    }
}

inline unsigned grid_for(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 148LL * 32;          // a few waves of resident CTAs; the kernels are grid-stride
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int stb_feature_gate_f32(const float* x, const float* gate, float* out, int B, int C, int D, int H, int W,
                                    void* stream) {
    if (!x || !gate || !out || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return STB_E_BADARG;
    const long long planes = (long long)B * C * D, hw = (long long)H * W;
    if (planes > 2147483647LL) return STB_E_BADARG;
    long long gy = (hw / 4 + 255) / 256;
    gy = gy < 1 ? 1 : (gy > 64 ? 64 : gy);
    gate_f32_kernel<<<dim3((unsigned)planes, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(x, gate, out, D, hw);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_feature_gate_cl16(const void* x, const float* gate, void* out, int f16, int B, int C, int Cpad, int D,
                                     int H, int W, void* stream) {
    if (!x || !gate || !out || B <= 0 || C <= 0 || Cpad < C || (Cpad & 7) || D <= 0 || H <= 0 || W <= 0) return STB_E_BADARG;
    const long long nvox = (long long)B * D * H * W;
    if (f16 == 2) {       // operand-split fp16 storage
        if (Cpad & 15) return STB_E_BADARG;
        gate_split_kernel<<<grid_for(nvox * (Cpad >> 3), 256), 256, 0, (cudaStream_t)stream>>>(
            (const uint16_t*)x, gate, (uint16_t*)out, C, Cpad, D, H, W, nvox);
        STB_CHECK_LAUNCH();
        return STB_OK;
    }
    gate_cl16_kernel<<<grid_for(nvox * (Cpad >> 3), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)x, gate, (uint16_t*)out, C, Cpad, D, H, W, nvox, f16);
    STB_CHECK_LAUNCH();
    return STB_OK;
}

extern "C" int stb_disparity_variance_f32(const float* prob, const float* disp, float* var, int B, int D, long long plane,
                                          void* stream) {
    if (!prob || !disp || !var || B <= 0 || D <= 0 || plane <= 0) return STB_E_BADARG;
    const long long total = (long long)B * plane;
    variance_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(prob, disp, var, D, plane, total);
    STB_CHECK_LAUNCH();
    return STB_OK;
}
