// Library-level entry points of the C ABI (include/stb200.h).
#include "common.cuh"

extern "C" const char* stb_error_string(int code) {
    switch (code) {
        case STB_OK: return "ok";
        case STB_E_BADARG: return "bad argument (null pointer, non-positive size or inconsistent shape)";
        case STB_E_UNSUPPORTED: return "configuration not supported by this kernel";
        case STB_E_SMEM: return "shared-memory footprint exceeds the sm_100a limit";
        case STB_E_DRIVER: return "CUDA driver entry point unavailable (cuTensorMapEncodeTiled)";
        default:
            if (code <= -1000) return cudaGetErrorString((cudaError_t)(-code - 1000));
            return "unknown error";
    }
}

extern "C" int stb_version(void) { return 100; }
