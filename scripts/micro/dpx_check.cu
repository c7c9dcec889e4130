// Developer check: do __vimax3_s16x2 / __vimin3_s16x2 (VIMNMX3.S16x2) match a scalar per-halfword emulation on sm_100a?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__host__ __device__ inline int16_t lo(uint32_t x) { return (int16_t)(x & 0xFFFF); }
__host__ __device__ inline int16_t hi(uint32_t x) { return (int16_t)(x >> 16); }
__host__ __device__ inline uint32_t pk(int16_t l, int16_t h) { return (uint32_t)(uint16_t)l | ((uint32_t)(uint16_t)h << 16); }
__global__ void k(const uint32_t *in, uint32_t *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
    out[2 * i] = __vimax3_s16x2(a, b, c);
    out[2 * i + 1] = __vimin3_s16x2(a, b, c);
}
int main() {
    const int n = 1 << 16;
    uint32_t *h = new uint32_t[3 * n], *o = new uint32_t[2 * n], *din, *dout;
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < 3 * n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint32_t v = (uint32_t)(s >> 16);
        if (i % 3 == 0 && (i / 3) % 4 == 0) v = pk((int16_t)((int)(v % 8321) - 4160), (int16_t)((int)((v >> 13) % 8321) - 4160));   // accumulator-like values
        h[i] = v; }
    cudaMalloc(&din, 3 * n * 4); cudaMalloc(&dout, 2 * n * 4);
    cudaMemcpy(din, h, 3 * n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(din, dout, n);
    cudaMemcpy(o, dout, 2 * n * 4, cudaMemcpyDeviceToHost);
    int bad_max = 0, bad_min = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t a = h[3 * i], b = h[3 * i + 1], c = h[3 * i + 2];
        auto mx = [](int16_t x, int16_t y, int16_t z) { int16_t m = x > y ? x : y; return m > z ? m : z; };
        auto mn = [](int16_t x, int16_t y, int16_t z) { int16_t m = x < y ? x : y; return m < z ? m : z; };
        uint32_t emx = pk(mx(lo(a), lo(b), lo(c)), mx(hi(a), hi(b), hi(c))), emn = pk(mn(lo(a), lo(b), lo(c)), mn(hi(a), hi(b), hi(c)));
        if (o[2 * i] != emx) { if (bad_max++ < 5) printf("max mismatch a=%08x b=%08x c=%08x got %08x want %08x\n", a, b, c, o[2 * i], emx); }
        if (o[2 * i + 1] != emn) { if (bad_min++ < 5) printf("min mismatch a=%08x b=%08x c=%08x got %08x want %08x\n", a, b, c, o[2 * i + 1], emn); }
    }
    printf("checked %d: %d max mismatches, %d min mismatches\n", n, bad_max, bad_min);
    return 0;
}
