// Microbenchmark: tcgen05.mma.cta_group::2 (M=256 over a CTA pair) issue/execute rate vs N.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|(1ull<<16)|(64ull<<32)|(1ull<<46)|(2ull<<61); }
__global__ void __cluster_dims__(2,1,1) k(int N, int nmma, long long* out, int mode){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar; __shared__ __align__(8) uint64_t sb[8]; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  for (int i=threadIdx.x;i<(16384+32768)/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  int warp=threadIdx.x>>5, lane=threadIdx.x&31;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;":"=r"(rank));
  if (threadIdx.x==0){ for(int i=0;i<8;++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&sb[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;");
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc=(1u<<4)|(1u<<7)|(1u<<10)|((uint32_t)(N>>3)<<17)|((uint32_t)(256>>4)<<24);
  uint64_t da=desc(s32(base)), db=desc(s32(base+16384));
  if (warp==1 && rank==0){
    long long t0=clock64();
    for(int i=0;i<nmma;++i){ uint32_t acc=i>0; int kk=i&3;
      asm volatile("{\n.reg .pred p, e;\nelect.sync _|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(da+2*kk),"l"(db+2*kk),"r"(idesc),"r"(acc):"memory");
      if (kk==3 && mode==1) asm volatile("{\n.reg .pred e;\n.reg .b16 m;\nelect.sync _|e, 0xffffffff;\nmov.b16 m, 3;\n@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n}"::"r"(s32(&sb[(i>>2)%5])):"memory");
      if (kk==3 && mode==2) asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"::"r"(s32(&sb[(i>>2)%5])):"memory");
    }
    asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"::"r"(s32(&bar)):"memory");
    uint32_t ok=0; while(!ok){ asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(&bar)),"r"(0u):"memory"); }
    long long t1=clock64();
    if (lane==0 && blockIdx.x==0){ out[0]=t1-t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;");
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(512u));
}
int main(){
  long long* d; cudaMalloc(&d,8); int smem=16384+32768+1024;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  for (int mode : {0,1,2}) for (int grid : {148}) for (int N : {16,224}){
    int nmma=2048; long long h=0;
    k<<<grid,128,smem>>>(N,nmma,d,mode); cudaDeviceSynchronize();
    k<<<grid,128,smem>>>(N,nmma,d,mode); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(&h,d,8,cudaMemcpyDeviceToHost);
    printf("mode=%d grid=%3d N=%3d: %8.1f cycles/MMA(M=256) (floor %d)  %s\n",mode,grid,N,(double)h/nmma,256*N/512,cudaGetErrorString(e));
  }
  return 0;
}
