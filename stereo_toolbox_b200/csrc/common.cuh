// Shared helpers for libstb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/stb200.h"

#define STB_CHECK_LAUNCH()                                   \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return -(1000 + (int)e__);   \
    } while (0)

static inline int stb_ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float stb_act(float v, int act) {
    switch (act) {
        case STB_ACT_RELU: return fmaxf(v, 0.f);
        case STB_ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
        case STB_ACT_MISH: {
            // x * tanh(softplus(x)); softplus threshold 20 like torch
            float sp = v > 20.f ? v : log1pf(expf(v));
            return v * tanhf(sp);
        }
        case STB_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case STB_ACT_TANH: return tanhf(v);
        default: return v;
    }
}
