// Microbenchmark: what does it take to stream 8 GB of u64 codes at HBM speed on B200?  (developer tool)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ uint4 ldg_nc(const uint4*p){uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p)); return r;}
__device__ __forceinline__ uint32_t work(uint4 v, uint32_t ql, uint32_t qh){ return min(__popc((v.y^qh)|(v.x^ql)), __popc((v.w^qh)|(v.z^ql))); }
// A: persistent grid-stride, CPT codes per thread per tile
template<int V> __global__ void kA(const uint4* __restrict__ p, uint64_t n16, uint32_t ql, uint32_t qh, uint32_t* out){
  uint32_t m=64; uint64_t tile=(uint64_t)blockDim.x*V;
  for(uint64_t base=(uint64_t)blockIdx.x*tile; base<n16; base+=(uint64_t)gridDim.x*tile){
    uint4 v[V];
    #pragma unroll
    for(int j=0;j<V;j++){ uint64_t i=base+(uint64_t)j*blockDim.x+threadIdx.x; v[j]= i<n16? ldg_nc(p+i): make_uint4(0,0,0,0);}
    #pragma unroll
    for(int j=0;j<V;j++) m=min(m,work(v[j],ql,qh));
  }
  if(m==0) atomicAdd(out,1);
}
// B: one tile per CTA
template<int V> __global__ void kB(const uint4* __restrict__ p, uint64_t n16, uint32_t ql, uint32_t qh, uint32_t* out){
  uint32_t m=64; uint64_t base=(uint64_t)blockIdx.x*blockDim.x*V;
  uint4 v[V];
  #pragma unroll
  for(int j=0;j<V;j++){ uint64_t i=base+(uint64_t)j*blockDim.x+threadIdx.x; v[j]= i<n16? ldg_nc(p+i): make_uint4(0,0,0,0);}
  #pragma unroll
  for(int j=0;j<V;j++) m=min(m,work(v[j],ql,qh));
  if(m==0) atomicAdd(out,1);
}
int main(){
  uint64_t bytes=8ull<<30, n16=bytes/16; uint4* p; uint32_t* out; CK(cudaMalloc(&p,bytes)); CK(cudaMalloc(&out,4)); CK(cudaMemset(p,0x5a,bytes));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run=[&](const char* name, auto launch){ for(int i=0;i<2;i++) launch(); cudaEventRecord(e0); for(int i=0;i<5;i++) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5; printf("%-28s %.3f ms  %.0f GB/s  (%s)\n",name,ms,bytes/ms/1e6,cudaGetErrorString(cudaGetLastError())); };
  int sms=148;
  run("A V=4 256thr grid 4/SM", [&]{ kA<4><<<sms*4,256>>>(p,n16,1,2,out); });
  run("A V=4 256thr grid 8/SM", [&]{ kA<4><<<sms*8,256>>>(p,n16,1,2,out); });
  run("A V=8 256thr grid 4/SM", [&]{ kA<8><<<sms*4,256>>>(p,n16,1,2,out); });
  run("A V=2 512thr grid 4/SM", [&]{ kA<2><<<sms*4,512>>>(p,n16,1,2,out); });
  run("A V=4 512thr grid 4/SM", [&]{ kA<4><<<sms*4,512>>>(p,n16,1,2,out); });
  run("A V=4 1024thr grid 2/SM", [&]{ kA<4><<<sms*2,1024>>>(p,n16,1,2,out); });
  run("B V=4 256thr", [&]{ kB<4><<<(unsigned)((n16+1023)/1024),256>>>(p,n16,1,2,out); });
  run("B V=8 256thr", [&]{ kB<8><<<(unsigned)((n16+2047)/2048),256>>>(p,n16,1,2,out); });
  run("B V=4 512thr", [&]{ kB<4><<<(unsigned)((n16+2047)/2048),512>>>(p,n16,1,2,out); });
  run("B V=2 256thr", [&]{ kB<2><<<(unsigned)((n16+511)/512),256>>>(p,n16,1,2,out); });
  return 0;
}
