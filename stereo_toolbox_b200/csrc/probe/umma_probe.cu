// Hardware probe for the facts conv3d_umma.cu depends on (run on a B200; one variant per process):
//   * K-major UMMA tiles written by TMA with SWIZZLE_128B / 64B / 32B rows (128/64/32-byte channel rows)
//   * ROW-SHIFTED operand views: start address = tile + j * row_bytes (the im2col-free 3x3x3 tap trick),
//     with base_offset = 0 and base_offset = (addr >> 7) & 7
//   * 5-D TMA box loads with out-of-bounds (negative) coordinates -> zero fill (conv padding)
// usage: umma_probe <swizzle:128|64|32> <shift_rows> <base_offset_mode:0|1> [N=32]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../umma.cuh"

using namespace umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int M = 128;
constexpr int ROWS_LOADED = 160;   // rows staged in smem (>= M + max shift)

struct Params {
    int K;            // elements per row (row_bytes = 2K = swizzle span)
    int N;
    int shift;        // rows
    int bo_mode;
    uint32_t layout;  // umma::Swizzle
};

__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_holder;
    const int row_bytes = p.K * 2;
    uint8_t* sA = smem;                                   // ROWS_LOADED rows
    uint8_t* sB = smem + ((ROWS_LOADED * row_bytes + 1023) / 1024) * 1024;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_holder, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_holder;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_tma, (ROWS_LOADED + p.N) * row_bytes);
        tma_load_2d(sA, &tmA, &bar_tma, 0, 0);
        tma_load_2d(sB, &tmB, &bar_tma, 0, 0);
        mbar_wait(&bar_tma, 0);
        tc_fence_after();
        const uint32_t idesc = instr_desc_f16(M, p.N, 1);
        const uint32_t a0 = smem_u32(sA) + p.shift * row_bytes;
        const uint32_t b0 = smem_u32(sB);
        const uint32_t sbo = 8 * row_bytes;
        for (int k = 0; k < p.K / 16; ++k) {
            const uint32_t aaddr = a0 + k * 32, baddr = b0 + k * 32;
            const uint32_t abo = p.bo_mode ? ((aaddr >> 7) & 7) : 0;
            uint64_t ad = smem_desc(aaddr, 16, sbo, p.layout, abo);
            uint64_t bd = smem_desc(baddr, 16, sbo, p.layout, 0);
            mma_f16_ss(tmem, ad, bd, idesc, k > 0);
        }
        mma_commit(&bar_mma);
    }
    __syncthreads();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + (threadIdx.x & 31)) * p.N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// 5-D halo load check: NDHWC tensor, box with negative origin
__global__ void halo_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* dump, int nelem, int c0, int w0,
                            int h0, int d0, int n0) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, nelem * 2);
        tma_load_5d(smem, &tm, &bar, c0, w0, h0, d0, n0);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) dump[i] = reinterpret_cast<__nv_bfloat16*>(smem)[i];
}

// TMA box-load throughput: every CTA streams `nbox` boxes [rows x 32 voxels x C ch] of an NDHWC tensor through
// a ring of `R` smem slots (re-issuing a slot as soon as its box has landed); no compute.
__global__ void __launch_bounds__(64)
tmabw_kernel(const __grid_constant__ CUtensorMap tm, int R, int nbox, uint32_t box_bytes, int D, int tiles_w, int tiles_h) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < R; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = blockIdx.x;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h; t /= tiles_h;
        const int b = t;
        for (int n = 0; n < nbox + R; ++n) {
            const int slot = n % R;
            if (n >= R) mbar_wait(&full[slot], ((n / R) - 1) & 1);      // box n-R has landed -> slot reusable
            if (n < nbox) {
                mbar_arrive_expect_tx(&full[slot], box_bytes);
                tma_load_5d(smem + (size_t)slot * box_bytes, &tm, &full[slot], 0, tw * 30 - 1, th * 8 - 1, (n % D), b);
            }
        }
    }
}

// Same streaming loop with explicit tile origins (tmabw2: unit-stride vs element-strided vs parity-folded tensor maps).
__global__ void __launch_bounds__(64)
tmabw2_kernel(const __grid_constant__ CUtensorMap tm, int R, int nbox, uint32_t box_bytes, int D, int tiles_w, int tiles_h,
              int step_w, int step_h) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < R; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = blockIdx.x;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h; t /= tiles_h;
        const int b = t;
        for (int n = 0; n < nbox + R; ++n) {
            const int slot = n % R;
            if (n >= R) mbar_wait(&full[slot], ((n / R) - 1) & 1);
            if (n < nbox) {
                mbar_arrive_expect_tx(&full[slot], box_bytes);
                tma_load_5d(smem + (size_t)slot * box_bytes, &tm, &full[slot], 0, tw * step_w, th * step_h, (n % D), b);
            }
        }
    }
}

// Raw tcgen05.mma issue/execute rate: `nissue` warps each issue `iters` x `per_iter` MMAs (M=128, N, K=16) on
// garbage smem operands (no TMA, no epilogue), operands K-major with the given swizzle / row shift.
__global__ void __launch_bounds__(128)
mmarate_kernel(int N, uint32_t layout, int rowb, int shift, int iters, int per_iter, int nissue, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    long long t0 = 0, t1 = 0;
    if (warp < nissue && elect_one()) {
        const uint32_t idesc = instr_desc_f16(128, N, 1);
        const uint64_t hi = (uint64_t)((((8u * rowb) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29)) << 32;
        const uint32_t a0 = ((smem_u32(smem) + warp * 32768 + shift * rowb) >> 4 & 0x3FFFu) | (1u << 16);
        const uint32_t b0 = ((smem_u32(smem) + 65536 + warp * 32768) >> 4 & 0x3FFFu) | (1u << 16);
        const uint32_t d = tmem + warp * 256;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int j = 0; j < per_iter; ++j) {
                const uint32_t off = (uint32_t)((j & 15) * (rowb * 8 / 16));   // walk over row groups like taps do
                mma_f16_ss(d, hi | (uint64_t)(a0 + off), hi | (uint64_t)(b0 + (j & 1) * 2), idesc, 1u);
            }
        }
        mma_commit(&bar[warp]);
        mbar_wait(&bar[warp], 0);
        t1 = clock64();
        if (blockIdx.x == 0) cycles[warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// mmarate2: as mmarate with more knobs -- operand format (0 fp16 / 1 bf16), smem filled with zeros or random
// normal-range values (data-dependent power throttling), `nacc` accumulators used round-robin (dependent chains),
// and `ldwarps` extra warps that keep reading TMEM with tcgen05.ld like the conv epilogue does.
__global__ void __launch_bounds__(256)
mmarate2_kernel(int N, uint32_t layout, int rowb, int fmt, int fill, int nacc, int ldwarps, int iters, int per_iter,
                long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t holder;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); done = 0; fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&holder, 512);
    uint16_t* s16 = reinterpret_cast<uint16_t*>(smem);
    for (int i = threadIdx.x; i < 130 * 1024 / 2; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const float v = fill ? ((float)(h & 0xffff) / 32768.f - 1.f) : 0.f;
        if (fmt) { __nv_bfloat16 b = __float2bfloat16(v); s16[i] = *reinterpret_cast<uint16_t*>(&b); }
        else { __half b = __float2half(v); s16[i] = *reinterpret_cast<uint16_t*>(&b); }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = instr_desc_f16(128, N, fmt);
        const uint64_t hi = (uint64_t)((((8u * rowb) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29)) << 32;
        const uint32_t a0 = ((smem_u32(smem)) >> 4 & 0x3FFFu) | (1u << 16);
        const uint32_t b0 = ((smem_u32(smem) + 65536) >> 4 & 0x3FFFu) | (1u << 16);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int j = 0; j < per_iter; ++j) {
                const uint32_t off = (uint32_t)((j & 15) * (rowb * 8 / 16));
                const uint32_t d = tmem + (uint32_t)((j & (nacc - 1)) * N);
                mma_f16_ss(d, hi | (uint64_t)(a0 + off), hi | (uint64_t)(b0 + (j & 1) * 2), idesc, 1u);
            }
        }
        mma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
        done = 1;
    } else if (warp >= 4 && warp < 4 + ldwarps) {
        uint32_t v[32];
        uint32_t acc = 0;
        while (!done) {
            __syncwarp();
            tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16), v);
            tmem_ld_wait();
            acc += v[0];
        }
        if (acc == 0x12345678u) cycles[1] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// mmarate3: the exact MMA address sequence of one accumulator round of the 32->32 k3 kw-merged layer (TH=8: 2 M-tiles,
// 3 planes x 3 kh taps x 2 K-steps, N=96, SW64): variant 0 = addresses as compile-time immediates (fully unrolled),
// variant 1 = per-tap records read from the kernel-parameter constant bank like conv3d_umma does.
struct Iss3 { uint32_t a16, b16, idesc, dcol; };
struct Tab3 { Iss3 e[9]; };
template <int VARIANT>
__global__ void __launch_bounds__(352)
mmarate3_kernel(const __grid_constant__ Tab3 tab, int iters, long long* cycles, int commit_groups, int iwarp, int nz, int nm) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t holder;
    __shared__ volatile int vbound;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); vbound = 3; fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&holder, 512);
    uint16_t* s16 = reinterpret_cast<uint16_t*>(smem);
    for (int i = threadIdx.x; i < 200 * 1024 / 2; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        __half b = __float2half((float)(h & 0xffff) / 32768.f - 1.f);
        s16[i] = *reinterpret_cast<uint16_t*>(&b);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (warp == iwarp && elect_one()) {
        const uint32_t idesc = instr_desc_f16(128, 96, 0);
        const uint32_t hi32 = (((8u * 64) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)SW_64B << 29);
        const uint64_t hi = (uint64_t)hi32 << 32;
        const uint32_t sW16 = (smem_u32(smem) >> 4) | (1u << 16);                 // weights: 9 tiles x 6144 B
        const uint32_t sP16 = ((smem_u32(smem) + 56 * 1024) >> 4) | (1u << 16);   // 3 plane slots x 20480 B
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (VARIANT == 0) {
#pragma unroll
                for (int z = 0; z < 3; ++z)
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint32_t alo = sP16 + (uint32_t)((z * 20480 + m * 8192 + kh * 2048) >> 4) + 2 * k;
                                const uint32_t blo = sW16 + (uint32_t)(((z * 3 + kh) * 6144) >> 4) + 2 * k;
                                mma_f16_ss(tmem + m * 96, hi | alo, hi | blo, idesc, (z | kh | k) ? 1u : 0u);
                                if (m == 1 && kh == 2 && k == 1 && z < commit_groups) mma_commit(&bar[1]);
                            }
            } else {
                for (int z = 0; z < nz; ++z) {
                    const uint32_t abase = sP16 + (uint32_t)z * (20480 >> 4);
                    for (int m = 0; m < nm; ++m) {
                        const uint32_t am = abase + (uint32_t)m * (8192 >> 4);
                        const uint32_t dcol = tmem + m * 96;
#pragma unroll 1
                        const int tend = z * 3 + (VARIANT == 2 ? vbound : 3);     // variant 2: bound in a VECTOR register
                        for (int tp = z * 3; tp < tend; ++tp) {
                            const Iss3 e = tab.e[tp];
                            const uint32_t alo = am + e.a16, blo = sW16 + e.b16;
                            mma_f16_ss2(dcol + e.dcol, alo, blo, hi32, e.idesc, tp != 0);
                            mma_f16_ss2(dcol + e.dcol, alo + 2u, blo + 2u, hi32, e.idesc, 1u);
                        }
                    }
                }
            }
        }
        mma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "halo")) {
        // tensor [N=2][D=4][H=5][W=6][C=32] bf16, box (C=32, W=8, H=4, D=3, N=1) at (0,-1,-1,-1,1), SWIZZLE_64B
        const int N = 2, D = 4, H = 5, W = 6, C = 32;
        std::vector<__nv_bfloat16> h((size_t)N * D * H * W * C);
        for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)(i % 4093) * 0.25f + 1.f);
        __nv_bfloat16 *d, *dump;
        CK(cudaMalloc(&d, h.size() * 2));
        CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        const int bw = 8, bh = 4, bd = 3, nelem = C * bw * bh * bd;
        CK(cudaMalloc(&dump, nelem * 2));
        CUtensorMap tm;
        uint64_t dims[5] = {C, W, H, D, N};
        uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
        uint32_t box[5] = {C, bw, bh, bd, 1};
        if (!umma_host::make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) {
            printf("halo: tensor map encode FAILED\n");
            return 1;
        }
        halo_kernel<<<1, 128, nelem * 2 + 1024>>>(tm, dump, nelem, 0, -1, -1, -1, 1);
        CK(cudaDeviceSynchronize());
        std::vector<__nv_bfloat16> o(nelem);
        CK(cudaMemcpy(o.data(), dump, nelem * 2, cudaMemcpyDeviceToHost));
        // expected: row r = ((dd*bh)+hh)*bw+ww ; 64-byte rows, 16B chunk index ^= (byte_addr>>7)&3
        int bad = 0;
        for (int r = 0; r < bw * bh * bd; ++r) {
            int ww = r % bw, hh = (r / bw) % bh, dd = r / (bw * bh);
            int gw = ww - 1, gh = hh - 1, gd = dd - 1;
            bool in = gw >= 0 && gw < W && gh >= 0 && gh < H && gd >= 0 && gd < D;
            for (int c = 0; c < C; ++c) {
                float want = in ? __bfloat162float(h[((((size_t)1 * D + gd) * H + gh) * W + gw) * C + c]) : 0.f;
                int chunk = c / 8, within = c % 8;
                int byte_row = r * 64;
                int pchunk = chunk ^ ((byte_row >> 7) & 3);
                float got = __bfloat162float(o[(size_t)r * 32 + pchunk * 8 + within]);
                if (got != want) { if (bad < 5) printf("halo mismatch r=%d c=%d want %g got %g\n", r, c, want, got); ++bad; }
            }
        }
        printf("halo 5D OOB zero-fill + SW64 layout: %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
        return bad ? 1 : 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "mmarate")) {
        // usage: mmarate <N> <rowb:64|128> <shift_rows> <nissue:1|2>
        const int N = atoi(argv[2]), rowb = atoi(argv[3]), shift = atoi(argv[4]), nissue = atoi(argv[5]);
        const uint32_t layout = rowb == 128 ? SW_128B : SW_64B;
        long long* dc;
        CK(cudaMalloc(&dc, 16));
        CK(cudaMemset(dc, 0, 16));
        const int iters = 200, per_iter = 36;
        size_t smem = 140 * 1024;
        CK(cudaFuncSetAttribute(mmarate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mmarate_kernel<<<148, 128, smem>>>(N, layout, rowb, shift, iters, per_iter, nissue, dc);
        CK(cudaDeviceSynchronize());
        mmarate_kernel<<<148, 128, smem>>>(N, layout, rowb, shift, iters, per_iter, nissue, dc);
        CK(cudaDeviceSynchronize());
        long long hc[2];
        CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
        const double per = (double)hc[0] / (iters * per_iter);
        printf("mmarate N=%d rowb=%d shift=%d issuers=%d : %.1f clk per MMA per issuer (%.1f clk per MMA overall), ideal math %.1f\n",
               N, rowb, shift, nissue, per, per / nissue, 128.0 * N / 256.0);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "mmarate2")) {
        // usage: mmarate2 <N> <rowb:64|128> <fmt:0 fp16|1 bf16> <fill:0|1> <nacc> <ldwarps>
        const int N = atoi(argv[2]), rowb = atoi(argv[3]), fmt = atoi(argv[4]), fill = atoi(argv[5]), nacc = atoi(argv[6]),
                  ldw = atoi(argv[7]);
        const uint32_t layout = rowb == 128 ? SW_128B : SW_64B;
        long long* dc;
        CK(cudaMalloc(&dc, 16));
        CK(cudaMemset(dc, 0, 16));
        const int iters = 200, per_iter = 36;
        size_t smem = 140 * 1024;
        CK(cudaFuncSetAttribute(mmarate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int rep = 0; rep < 2; ++rep) {
            mmarate2_kernel<<<148, 256, smem>>>(N, layout, rowb, fmt, fill, nacc, ldw, iters, per_iter, dc);
            CK(cudaDeviceSynchronize());
        }
        long long hc[2];
        CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
        printf("mmarate2 N=%d rowb=%d fmt=%s fill=%s nacc=%d ldwarps=%d : %.1f clk per MMA (ideal math %.1f, smem operand bytes/128 = %.1f)\n",
               N, rowb, fmt ? "bf16" : "fp16", fill ? "random" : "zeros", nacc, ldw, (double)hc[0] / (iters * per_iter),
               128.0 * N / 256.0, (128.0 * 32 + N * 32.0) / 128.0);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "mmarate3")) {
        const int variant = atoi(argv[2]);
        long long* dc;
        CK(cudaMalloc(&dc, 16));
        CK(cudaMemset(dc, 0, 16));
        Tab3 tab;
        for (int z = 0; z < 3; ++z)
            for (int kh = 0; kh < 3; ++kh) {
                Iss3& e = tab.e[z * 3 + kh];
                e.a16 = (kh * 2048) >> 4;
                e.b16 = ((z * 3 + kh) * 6144) >> 4;
                e.idesc = instr_desc_f16(128, 96, 0);
                e.dcol = 0;
            }
        const int iters = argc > 7 ? atoi(argv[7]) : 400;
        const int nthr = argc > 4 ? atoi(argv[4]) : 128, iwarp = argc > 5 ? atoi(argv[5]) : 0;
        size_t smem = (argc > 6 ? atoi(argv[6]) : 210) * 1024;
        CK(cudaFuncSetAttribute(mmarate3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(mmarate3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(mmarate3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t ev0, ev1;
        cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(ev0);
            if (variant == 0) mmarate3_kernel<0><<<148, nthr, smem>>>(tab, iters, dc, argc > 3 ? atoi(argv[3]) : 0, iwarp, 3, 2);
            else if (variant == 1) mmarate3_kernel<1><<<148, nthr, smem>>>(tab, iters, dc, 0, iwarp, 3, 2);
            else mmarate3_kernel<2><<<148, nthr, smem>>>(tab, iters, dc, 0, iwarp, 3, 2);
            cudaEventRecord(ev1);
            CK(cudaDeviceSynchronize());
        }
        float wall_ms = 0.f;
        cudaEventElapsedTime(&wall_ms, ev0, ev1);
        printf("  wall %.1f us for %d MMAs per SM = %.2f ns per MMA (x1.965 GHz = %.1f clk)\n", wall_ms * 1e3, iters * 36,
               wall_ms * 1e6 / (iters * 36.0), wall_ms * 1e6 / (iters * 36.0) * 1.965);
        long long hc[2];
        CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
        printf("mmarate3 threads=%d issuer_warp=%d smem=%zuKB variant=%d commits/36=%d (%s): %.1f clk per MMA (N=96 K=16; smem-operand bound 56)\n", nthr, iwarp, smem / 1024, variant, argc > 3 ? atoi(argv[3]) : 0,
               variant ? "constant-bank records" : "immediates", (double)hc[0] / (iters * 36.0));
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "tmabw")) {
        // usage: tmabw <C:32|64> <rows> <R>
        const int C = atoi(argv[2]), rows = atoi(argv[3]), R = atoi(argv[4]);
        const int B = 8, D = 48, H = 96, W = 312;
        size_t n = (size_t)B * D * H * W * C;
        __nv_bfloat16* d;
        CK(cudaMalloc(&d, n * 2));
        CK(cudaMemset(d, 0, n * 2));
        CUtensorMap tm;
        uint64_t dims[5] = {(uint64_t)C, W, H, D, B};
        uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
        uint32_t box[5] = {(uint32_t)C, 32, (uint32_t)rows, 1, 1};
        if (!umma_host::make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box,
                                  C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) { printf("tmap fail\n"); return 1; }
        const uint32_t box_bytes = (uint32_t)C * 2 * 32 * rows;
        const int tiles_w = 11, tiles_h = 12, nbox = 96;
        size_t smem = (size_t)R * box_bytes + 2048;
        CK(cudaFuncSetAttribute(tmabw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int grid = B * tiles_w * tiles_h;
        tmabw_kernel<<<grid, 64, smem>>>(tm, R, nbox, box_bytes, D, tiles_w, tiles_h);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        tmabw_kernel<<<grid, 64, smem>>>(tm, R, nbox, box_bytes, D, tiles_w, tiles_h);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)grid * nbox * box_bytes / 1e9;
        printf("tmabw C=%d rows=%d (box %u B) ring=%d smem=%zu KB: %.3f ms, %.1f GB/s (%d CTAs x %d boxes)\n", C, rows, box_bytes, R,
               smem / 1024, ms, gb / (ms * 1e-3), grid, nbox);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "tmabw2")) {
        // usage: tmabw2 <C:32|64> <rows> <R> <mode>   every box delivers rows x 32 voxels x C channels to shared memory
        //   mode 0: unit stride (the stride-1 conv's plane load)
        //   mode 1: every second voxel in h and w through TMA elementStrides (1,2,2,1,1) -- what the stride-2 conv does today
        //   mode 2: the same voxels through a parity-folded tensor map: dims (C, W/2, H/2, D, B) with DOUBLED w / h byte
        //           strides and unit element strides (one map per (h,w) parity, base pointer offset by the parity)
        const int C = atoi(argv[2]), rows = atoi(argv[3]), R = atoi(argv[4]), mode = argc > 5 ? atoi(argv[5]) : 0;
        const int B = 8, D = 48, H = 96, W = 312;
        size_t n = (size_t)B * D * H * W * C;
        __nv_bfloat16* d;
        CK(cudaMalloc(&d, n * 2));
        CK(cudaMemset(d, 0, n * 2));
        CUtensorMap tm;
        const CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        bool ok;
        int step_w, step_h;
        if (mode == 2) {
            uint64_t dims[5] = {(uint64_t)C, W / 2, H / 2, D, B};
            uint64_t str[4] = {(uint64_t)2 * C * 2, (uint64_t)2 * W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
            uint32_t box[5] = {(uint32_t)C, 32, (uint32_t)rows, 1, 1};
            ok = umma_host::make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, sw);
            step_w = 30; step_h = rows - 2;
        } else {
            uint64_t dims[5] = {(uint64_t)C, W, H, D, B};
            uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
            const uint32_t m = mode == 1 ? 2u : 1u;
            uint32_t box[5] = {(uint32_t)C, 32 * m, (uint32_t)rows * m, 1, 1};
            uint32_t es[5] = {1, m, m, 1, 1};
            ok = umma_host::make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, sw, es);
            step_w = 30 * (int)m; step_h = (rows - 2) * (int)m;
        }
        if (!ok) { printf("tmap fail\n"); return 1; }
        const uint32_t box_bytes = (uint32_t)C * 2 * 32 * rows;
        const int tiles_w = 5, tiles_h = mode == 0 ? 96 / (rows - 2) / 2 : 48 / (rows - 2), nbox = 96;   // same CTA count in all modes
        size_t smem = (size_t)R * box_bytes + 2048;
        CK(cudaFuncSetAttribute(tmabw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int grid = B * tiles_w * tiles_h;
        tmabw2_kernel<<<grid, 64, smem>>>(tm, R, nbox, box_bytes, D, tiles_w, tiles_h, step_w, step_h);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        tmabw2_kernel<<<grid, 64, smem>>>(tm, R, nbox, box_bytes, D, tiles_w, tiles_h, step_w, step_h);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)grid * nbox * box_bytes / 1e9;
        printf("tmabw2 mode=%d (%s) C=%d rows=%d (box %u B delivered) ring=%d: %.3f ms, %.1f GB/s delivered (%d CTAs x %d boxes)\n", mode,
               mode == 0 ? "unit stride" : mode == 1 ? "elementStrides 2,2" : "parity-folded map", C, rows, box_bytes, R, ms,
               gb / (ms * 1e-3), grid, nbox);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "halo2")) {
        // stride-2 deinterleave by TMA elementStrides: tensor [N=1][D=2][H=9][W=13][C=32], element strides
        // (1,2,2,1,1); box extents given in the ORIGINAL index space (2*8, 2*4) -> 8 x 4 voxels; start (-1,-1)
        const int N = 1, D = 2, H = 9, W = 13, C = 32;
        std::vector<__nv_bfloat16> h((size_t)N * D * H * W * C);
        for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)(i % 4093) * 0.25f + 1.f);
        __nv_bfloat16 *d, *dump;
        CK(cudaMalloc(&d, h.size() * 2));
        CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        const int bw = 8, bh = 4, nelem = C * bw * bh;
        CK(cudaMalloc(&dump, nelem * 2));
        CK(cudaMemset(dump, 0xff, nelem * 2));
        int variant = argc > 2 ? atoi(argv[2]) : 0;       // 0: boxDim = 2*n ; 1: boxDim = n
        CUtensorMap tm;
        uint64_t dims[5] = {C, W, H, D, N};
        uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
        uint32_t box[5] = {C, (uint32_t)(variant ? bw : 2 * bw), (uint32_t)(variant ? bh : 2 * bh), 1, 1};
        uint32_t es[5] = {1, 2, 2, 1, 1};
        if (!umma_host::make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, es)) {
            printf("halo2 variant %d: tensor map encode FAILED\n", variant);
            return 1;
        }
        halo_kernel<<<1, 128, nelem * 2 + 1024>>>(tm, dump, nelem, 0, -1, -1, 1, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("halo2 variant %d: kernel failed: %s (byte count mismatch -> trap)\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<__nv_bfloat16> o(nelem);
        CK(cudaMemcpy(o.data(), dump, nelem * 2, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r = 0; r < bw * bh; ++r) {
            int ww = r % bw, hh = r / bw;
            int gw = -1 + 2 * ww, gh = -1 + 2 * hh, gd = 1;
            bool in = gw >= 0 && gw < W && gh >= 0 && gh < H;
            for (int c = 0; c < C; ++c) {
                float want = in ? __bfloat162float(h[((((size_t)0 * D + gd) * H + gh) * W + gw) * C + c]) : 0.f;
                int pchunk = (c / 8) ^ (((r * 64) >> 7) & 3);
                float got = __bfloat162float(o[(size_t)r * 32 + pchunk * 8 + c % 8]);
                if (got != want) { if (bad < 5) printf("halo2 mismatch r=%d c=%d want %g got %g\n", r, c, want, got); ++bad; }
            }
        }
        printf("halo2 (elementStrides=2, variant %d: boxDim=%s): %s (%d mismatches)\n", variant, variant ? "n" : "2n", bad ? "FAIL" : "PASS", bad);
        return bad ? 1 : 0;
    }
    if (argc < 4) { printf("usage\n"); return 1; }
    const int sw = atoi(argv[1]);
    Params p;
    p.shift = atoi(argv[2]);
    p.bo_mode = atoi(argv[3]);
    p.N = argc > 4 ? atoi(argv[4]) : 32;
    p.K = sw / 2;
    p.layout = sw == 128 ? SW_128B : sw == 64 ? SW_64B : SW_32B;
    CUtensorMapSwizzle cusw = sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const int K = p.K, N = p.N;
    std::vector<float> A((size_t)ROWS_LOADED * K), B((size_t)N * K);
    std::vector<__nv_bfloat16> Ab(A.size()), Bb(B.size());
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = bf((rand() % 2001 - 1000) / 500.f); Ab[i] = __float2bfloat16(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = bf((rand() % 2001 - 1000) / 500.f); Bb[i] = __float2bfloat16(B[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dO;
    CK(cudaMalloc(&dA, Ab.size() * 2));
    CK(cudaMalloc(&dB, Bb.size() * 2));
    CK(cudaMalloc(&dO, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0, (size_t)M * N * 4));
    CUtensorMap tmA, tmB;
    uint64_t dimsA[2] = {(uint64_t)K, ROWS_LOADED}, dimsB[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t boxA[2] = {(uint32_t)K, ROWS_LOADED}, boxB[2] = {(uint32_t)K, (uint32_t)N};
    if (!umma_host::make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, str, boxA, cusw) ||
        !umma_host::make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, boxB, cusw)) {
        printf("tensor map encode FAILED\n");
        return 1;
    }
    size_t smem = ((ROWS_LOADED * K * 2 + 1023) / 1024) * 1024 + N * K * 2 + 2048;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(tmA, tmB, dO, p);
    CK(cudaDeviceSynchronize());
    std::vector<float> O((size_t)M * N);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    // expected with row shift; also find, for the first rows, which source row each output row actually used
    double maxerr = 0;
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            float acc = 0;
            for (int k = 0; k < K; ++k) acc += A[(size_t)(r + p.shift) * K + k] * B[(size_t)n * K + k];
            maxerr = fmax(maxerr, fabs(acc - O[(size_t)r * N + n]));
        }
    printf("sw=%d shift=%d bo_mode=%d N=%d : maxerr %.4g -> %s\n", sw, p.shift, p.bo_mode, N, maxerr, maxerr < 1e-2 ? "PASS" : "FAIL");
    if (maxerr >= 1e-2) {
        printf("  row mapping (out row -> best source row, err): ");
        for (int r = 0; r < 16; ++r) {
            int best = -1; double be = 1e30;
            for (int s = 0; s < ROWS_LOADED; ++s) {
                double e = 0;
                for (int n = 0; n < N; ++n) {
                    float acc = 0;
                    for (int k = 0; k < K; ++k) acc += A[(size_t)s * K + k] * B[(size_t)n * K + k];
                    e = fmax(e, fabs(acc - O[(size_t)r * N + n]));
                }
                if (e < be) { be = e; best = s; }
            }
            printf("%d->%d(%.2g) ", r, best, be);
        }
        printf("\n");
    }
    return maxerr < 1e-2 ? 0 : 1;
}
