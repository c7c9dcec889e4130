// Developer microbenchmark: what bounds an epilogue that has to LOOK at every accumulator of an int8 tcgen05.mma?
// One CTA per SM, no data dependence between the roles, so each figure is a pure throughput:
//   mode 0  MMA only      warp 0 issues M128 x N256 x K64 kind::i8 tiles back to back into two 256-column TMEM stages
//   mode 1  loads only    16 warps each read "their" 64 columns of a stage over and over (tcgen05.ld 32x32b.x32.pack::16b)
//   mode 2  both at once  the two streams run unsynchronised; if the TMEM port is shared the times add
//   mode 3  loads only, unpacked (32x32b.x32 on 32 columns, twice per 64 columns)
//   mode 4  loads only, x16 packed (two per 64 columns)
//   mode 5  both at once, unpacked loads
// Output: clocks per 128 x 256 accumulator tile for the MMA stream and for the load stream (16 warps x 64 columns = one tile).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_port tmem_port.cu && ./tmem_port
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void umma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
#define R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
    "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define L32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}"
#define R16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define L16 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}"

constexpr int kTiles = 4096;     // accumulator tiles per CTA and stream
constexpr int kThreads = 32 * 17;

__global__ void __launch_bounds__(kThreads, 1) tmem_port_kernel(int mode, unsigned long long *out) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    unsigned char *sA = smem, *sB = smem + 16384;            // 128 x 128 B query tile, 256 x 128 B operand rows (zeros)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384 + 32768);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 1);
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const bool do_mma = mode == 0 || mode == 2 || mode == 5, do_ld = mode != 0;
    const long long t0 = clock64();
    long long t1 = t0;
    if (warp == 0) {
        if (do_mma) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t ad = desc_sw128(smem_u32(sA)), bd = desc_sw128(smem_u32(sB));
            if (lane == 0) {
                for (int t = 0; t < kTiles; ++t) {
                    umma_i8(tmem + (t & 1) * 256, ad, bd, idesc, 0u);
                    umma_i8(tmem + (t & 1) * 256, ad + 2, bd + 2, idesc, 1u);
                }
                umma_commit(&bar[0]);
            }
            __syncwarp();
            mbar_wait(&bar[0], 0);
            t1 = clock64();
        }
    } else if (do_ld) {
        const uint32_t quad = warp & 3, part = (uint32_t)(warp - 1) >> 2;
        const uint32_t taddr = tmem + ((quad * 32u) << 16) + part * 64;
        uint32_t acc = 0;
        for (int t = 0; t < kTiles; ++t) {
            const uint32_t ta = taddr + (t & 1) * 256;
            if (mode == 1 || mode == 2) {
                uint32_t v[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32 ", [%32];" : R32(v) : "r"(ta));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= v[0] ^ v[31];
            } else if (mode == 3 || mode == 5) {
                uint32_t v[32], w[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : R32(v) : "r"(ta));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : R32(w) : "r"(ta + 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= v[0] ^ w[31];
            } else {
                uint32_t v[16], w[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 " L16 ", [%16];" : R16(v) : "r"(ta));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 " L16 ", [%16];" : R16(w) : "r"(ta + 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= v[0] ^ w[15];
            }
        }
        t1 = clock64();
        if (acc == 0x12345u) out[63] = acc;
    }
    if (lane == 0 && blockIdx.x == 0) out[warp] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    unsigned long long *out, h[64];
    cudaMalloc(&out, 64 * 8);
    const size_t smem = 16384 + 32768 + 1024 + 256;
    cudaFuncSetAttribute(tmem_port_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char *names[] = {"MMA only", "ld x32 pack16 only", "MMA + ld x32 pack16", "ld x32 unpacked only", "ld 2 x (x16 pack16) only", "MMA + ld x32 unpacked"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(out, 0, 64 * 8);
            tmem_port_kernel<<<148, kThreads, smem>>>(mode, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost);
        unsigned long long ld_max = 0;
        for (int w = 1; w < 17; ++w) ld_max = h[w] > ld_max ? h[w] : ld_max;
        printf("mode %d %-26s: MMA stream %7.1f clk/tile   load stream %7.1f clk/tile (16 warps x 64 columns)\n", mode, names[mode],
               (double)h[0] / kTiles, (double)ld_max / kTiles);
    }
    return 0;
}
