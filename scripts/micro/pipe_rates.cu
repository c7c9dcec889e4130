// Developer microbenchmark: issue rate of the integer / half min-max and shift instructions the tensor Hamming
// epilogue is made of.  One CTA of 4 warps per SM (one warp per SMSP... x WARPS), 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
template <int OP> __device__ __forceinline__ uint32_t step(uint32_t a, uint32_t b, uint32_t c) {
    if (OP == 0) return (uint32_t)max(max((int32_t)a, (int32_t)b), (int32_t)c);            // VIMNMX3
    if (OP == 1) return (uint32_t)max((int32_t)a, (int32_t)b);                              // VIMNMX
    if (OP == 2) { __half2 r = __hmax2(__hmax2(*(__half2 *)&a, *(__half2 *)&b), *(__half2 *)&c); return *(uint32_t *)&r; }   // VHMNMX
    if (OP == 3) { __half2 r = __hmax2(*(__half2 *)&a, *(__half2 *)&b); return *(uint32_t *)&r; }                           // HMNMX2
    if (OP == 4) return a * 0x200u + b;                                                    // IMAD.SHL
    if (OP == 5) return (a & b) ^ c;                                                        // LOP3
    if (OP == 6) return a + b + c;                                                          // IADD3
    if (OP == 7) return a * b + c;                                                          // IMAD
    if (OP == 8) return __vimax3_s16x2(a, b, c);                                            // VIMNMX3.S16x2
    if (OP == 9) return __float_as_uint(fmaxf(__uint_as_float(a), __uint_as_float(b)));     // FMNMX
    if (OP == 10) return __vimin3_s16x2(a, b, c);                                           // VIMNMX3.S16x2 min
    if (OP == 11) return __vimin3_s16x2(a * 512u, b * 512u, c);                             // 2 IMAD.SHL + min
    if (OP == 12) return __vimax3_s16x2(a, b ^ 0x80008000u, c);                             // negative halves
    if (OP == 13) { __half2 r = __hfma2(*(__half2 *)&a, *(__half2 *)&b, *(__half2 *)&c); return *(uint32_t *)&r; }        // HFMA2 (fma pipe)
    if (OP == 14) { __half2 r = __hfma2_relu(*(__half2 *)&a, *(__half2 *)&b, *(__half2 *)&c); return *(uint32_t *)&r; }   // HFMA2.RELU
    return a;
}
template <int OP> __global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, unsigned long long *cyc) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = seed * (threadIdx.x + 17 * j + 1);
    
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = step<OP>(v[j], v[(j + 3 + r) & 7], v[(j + 5) & 7]);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}
// Two instruction kinds in one loop, four independent chains each: time per instruction ~ the mean of the two rates when they share a
// pipe, ~ half the slower one (but >= 1 issue slot) when they do not.
template <int OPA, int OPB> __global__ void __launch_bounds__(1024) kmix(uint32_t *out, uint32_t seed, unsigned long long *cyc) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = seed * (threadIdx.x + 17 * j + 1);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = (j & 1) ? step<OPB>(v[j], v[(j + 2 + 2 * (r & 1)) & 7], v[(j + 4) & 7]) : step<OPA>(v[j], v[(j + 2 + 2 * (r & 1)) & 7], v[(j + 4) & 7]);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}
template <int OPA, int OPB> void runmix(const char *name, uint32_t *out, unsigned long long *cyc) {
    for (int warps : {8, 16}) {
        kmix<OPA, OPB><<<148, warps * 32>>>(out, 12345u, cyc);
        cudaDeviceSynchronize();
        unsigned long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double inst_per_smsp = (double)ITER * 32 * (warps / 4);
        printf("%-28s warps/SM %2d: %.2f cycles per warp-instruction per SMSP\n", name, warps, (double)h / inst_per_smsp);
    }
}
template <int OP> void run(const char *name, uint32_t *out, unsigned long long *cyc) {
    for (int warps : {4, 8, 16}) {
        k<OP><<<148, warps * 32>>>(out, 12345u, cyc);
        cudaDeviceSynchronize();
        unsigned long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double inst_per_smsp = (double)ITER * 32 * (warps / 4);
        printf("%-14s warps/SM %2d: %.2f cycles per warp-instruction per SMSP\n", name, warps, (double)h / inst_per_smsp);
    }
}
int main() {
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<0>("VIMNMX3", out, cyc); run<1>("VIMNMX", out, cyc); run<2>("VHMNMX(3in)", out, cyc); run<3>("HMNMX2", out, cyc);
    run<4>("IMAD.SHL", out, cyc); run<5>("LOP3", out, cyc); run<6>("IADD3", out, cyc); run<7>("IMAD", out, cyc);
    run<8>("VIMNMX3.S16x2", out, cyc); run<9>("FMNMX", out, cyc); run<10>("S16x2 min", out, cyc); run<11>("2SHL+S16x2min", out, cyc); run<12>("S16x2 mixedsign", out, cyc);
    run<13>("HFMA2", out, cyc); run<14>("HFMA2.RELU", out, cyc);
    runmix<8, 4>("VIMNMX3.S16x2 + IMAD.SHL", out, cyc); runmix<8, 3>("VIMNMX3.S16x2 + HMNMX2", out, cyc); runmix<8, 2>("VIMNMX3.S16x2 + HMNMX2(3in)", out, cyc);
    runmix<8, 13>("VIMNMX3.S16x2 + HFMA2", out, cyc); runmix<8, 14>("VIMNMX3.S16x2 + HFMA2.RELU", out, cyc); runmix<8, 9>("VIMNMX3.S16x2 + FMNMX", out, cyc);
    runmix<8, 5>("VIMNMX3.S16x2 + LOP3", out, cyc); runmix<4, 13>("IMAD.SHL + HFMA2", out, cyc);
    return 0;
}
