// Microbenchmark 2: add the real kernel's features one at a time (developer tool)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ uint4 ldg_nc(const uint4*p){uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p)); return r;}
__device__ __forceinline__ uint32_t xor_or(uint32_t a, uint32_t b, uint32_t c){uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xBE;":"=r"(r):"r"(a),"r"(b),"r"(c)); return r;}
__global__ void fill(uint64_t* p, uint64_t n){ uint64_t i=(uint64_t)blockIdx.x*blockDim.x+threadIdx.x, s=(uint64_t)gridDim.x*blockDim.x; for(;i<n;i+=s){uint64_t z=(i+1)*0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; p[i]=z^(z>>31);} }
struct __align__(16) QSlot{uint32_t lo,hi,thr,pad;};
template<bool NOINLINE> struct Surv;
__device__ __noinline__ void surv_noinline(const uint32_t (&lo)[8], const uint32_t (&hi)[8], uint4 s, uint32_t* out){
  #pragma unroll
  for(int c=0;c<8;c++){ uint32_t d=__popc(lo[c]^s.x)+__popc(hi[c]^s.y); if(d<=s.z) atomicAdd(out,1);} }
// MODE bit0: guards like the real kernel; bit1: smem query loop; bit2: noinline survivors w/ arrays by reference
__device__ __noinline__ void surv_byval(uint4 a, uint4 b, uint4 c, uint4 d, uint4 s, uint32_t* out){
  const uint32_t lo[8]={a.x,a.z,b.x,b.z,c.x,c.z,d.x,d.z}, hi[8]={a.y,a.w,b.y,b.w,c.y,c.w,d.y,d.w};
  #pragma unroll
  for(int i=0;i<8;i++){ uint32_t dd=__popc(lo[i]^s.x)+__popc(hi[i]^s.y); if(dd<=s.z) atomicAdd(out,1);} }
template<int MODE> __global__ void __launch_bounds__(256) kR(const uint64_t* __restrict__ codes, uint64_t nrows, const QSlot* slots, uint32_t nq, uint32_t* out){
  extern __shared__ uint4 sq[];
  if(MODE&2){ for(uint32_t i=threadIdx.x;i<nq;i+=256) sq[i]=reinterpret_cast<const uint4*>(slots)[i]; __syncthreads(); }
  const uint64_t ntiles=(nrows+2047)/2048; uint32_t macc=64;
  for(uint64_t tile=blockIdx.x; tile<ntiles; tile+=gridDim.x){
    const uint64_t tile_row=tile*2048; uint32_t lo[8],hi[8];
    #pragma unroll
    for(int j=0;j<4;j++){ uint64_t r=tile_row+2ull*(j*256+threadIdx.x); uint4 v=make_uint4(0,0,0,0);
      if(MODE&1){ if(r+1<nrows) v=ldg_nc(reinterpret_cast<const uint4*>(codes+r)); else if(r<nrows){uint64_t c=codes[r]; v.x=(uint32_t)c; v.y=(uint32_t)(c>>32);} }
      else v=ldg_nc(reinterpret_cast<const uint4*>(codes+r));
      lo[2*j]=v.x;hi[2*j]=v.y;lo[2*j+1]=v.z;hi[2*j+1]=v.w; }
    if(MODE&2){
      for(uint32_t q=0;q<nq;q++){ const uint4 s=sq[q]; uint32_t m=64;
        #pragma unroll
        for(int c=0;c<8;c++) m=min(m,(uint32_t)__popc(xor_or(hi[c],s.y,lo[c]^s.x)));
        if(m<=s.z){ if(MODE&8) surv_byval(make_uint4(lo[0],hi[0],lo[1],hi[1]),make_uint4(lo[2],hi[2],lo[3],hi[3]),make_uint4(lo[4],hi[4],lo[5],hi[5]),make_uint4(lo[6],hi[6],lo[7],hi[7]),s,out); else if(MODE&4) surv_noinline(lo,hi,s,out); else atomicAdd(out,1); } }
    } else {
      #pragma unroll
      for(int c=0;c<8;c++) macc=min(macc,(uint32_t)__popc(xor_or(hi[c],2u,lo[c]^1u)));
    }
  }
  if(!(MODE&2) && macc==0) atomicAdd(out,1);
}
int main(){
  uint64_t nrows=1ull<<30, bytes=nrows*8; uint64_t* p; uint32_t* out; QSlot* slots; CK(cudaMalloc(&p,bytes+256)); CK(cudaMalloc(&out,4)); CK(cudaMalloc(&slots,16*16));
  fill<<<148*8,256>>>(p,nrows); QSlot h[16]; for(int i=0;i<16;i++) h[i]=QSlot{0x12345678u*(i+1),0x9abcdef0u+i,11,0}; CK(cudaMemcpy(slots,h,sizeof(h),cudaMemcpyHostToDevice)); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run=[&](const char* name, auto launch){ for(int i=0;i<2;i++) launch(); cudaEventRecord(e0); for(int i=0;i<5;i++) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5; printf("%-44s %.3f ms  %.0f GB/s  (%s)\n",name,ms,bytes/ms/1e6,cudaGetErrorString(cudaGetLastError())); };
  int g=148*4;
  run("random data, plain",            [&]{ kR<0><<<g,256,256>>>(p,nrows,slots,1,out); });
  run("+guards",                       [&]{ kR<1><<<g,256,256>>>(p,nrows,slots,1,out); });
  run("+smem query loop nq=1",         [&]{ kR<3><<<g,256,256>>>(p,nrows,slots,1,out); });
  run("+noinline survivors nq=1",      [&]{ kR<7><<<g,256,256>>>(p,nrows,slots,1,out); });
  run("+noinline survivors nq=2",      [&]{ kR<7><<<g,256,256>>>(p,nrows,slots,2,out); });
  run("+noinline survivors nq=4",      [&]{ kR<7><<<g,256,256>>>(p,nrows,slots,4,out); });
  run("smem loop nq=4 (no noinline)",  [&]{ kR<3><<<g,256,256>>>(p,nrows,slots,4,out); });
  run("by-value noinline nq=1",        [&]{ kR<11><<<g,256,256>>>(p,nrows,slots,1,out); });
  run("by-value noinline nq=2",        [&]{ kR<11><<<g,256,256>>>(p,nrows,slots,2,out); });
  run("by-value noinline nq=4",        [&]{ kR<11><<<g,256,256>>>(p,nrows,slots,4,out); });
  run("by-value noinline nq=16",       [&]{ kR<11><<<g,256,256>>>(p,nrows,slots,16,out); });
  run("+noinline nq=1 grid 8/SM",      [&]{ kR<7><<<148*8,256,256>>>(p,nrows,slots,1,out); });
  return 0;
}
