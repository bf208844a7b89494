// Microbenchmark: back-to-back tcgen05.mma kind::tf32 issue rate (M=128, K=8, SS mode) vs N; K-major SWIZZLE_128B
// operands (mode 1) and MN-major SWIZZLE_128B_BASE32B operands (mode 2, the NCHW conv's layout).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|(1ull<<16)|(64ull<<32)|(1ull<<46)|(2ull<<61); }
__global__ void k(int N, int nmma, int mode, long long* out){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  for (int i=threadIdx.x;i<(16384+32768)/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  int warp=threadIdx.x>>5, lane=threadIdx.x&31;
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc=(1u<<4)|(2u<<7)|(2u<<10)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  uint64_t da=desc(s32(base)), db=desc(s32(base+16384));
  if (mode==2){ idesc|=(1u<<15)|(1u<<16); auto mk=[](uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|(64ull<<16)|(32ull<<32)|(1ull<<46)|(1ull<<61); }; da=mk(s32(base)); db=mk(s32(base+16384)); }
  long long t0=0,t1=0;
  if (warp==1){
    if (mode==99){ // divergent single lane (like `if (lane == 0)`)
      if (lane==0){
        t0=clock64();
        for(int i=0;i<nmma;++i){ uint32_t acc=i>0; int k=i&3;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(da+(mode==2?0:2*k)),"l"(db+(mode==2?0:2*k)),"r"(idesc),"r"(acc):"memory"); }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(s32(&bar)):"memory");
      }
    } else { // whole warp converged, elect.sync picks the issuer (CUTLASS style)
      t0=clock64();
      for(int i=0;i<nmma;++i){ uint32_t acc=i>0; int k=i&3;
        asm volatile("{\n.reg .pred p, e;\nelect.sync _|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(da+(mode==2?0:2*k)),"l"(db+(mode==2?0:2*k)),"r"(idesc),"r"(acc):"memory"); }
      asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"::"r"(s32(&bar)):"memory");
    }
    uint32_t ok=0; while(!ok){ asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(&bar)),"r"(0u):"memory"); }
    t1=clock64();
    if (lane==0 && blockIdx.x==0){ out[0]=t1-t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(512u));
}
int main(){
  long long* d; cudaMalloc(&d,8); int smem=16384+32768+1024;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  for (int grid : {148}) for (int mode=1;mode<3;++mode) for (int N : {32,64,96,128,224,256}){
    int nmma=2048; long long h=0;
    k<<<grid,128,smem>>>(N,nmma,mode,d); cudaDeviceSynchronize();
    k<<<grid,128,smem>>>(N,nmma,mode,d); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(&h,d,8,cudaMemcpyDeviceToHost);
    printf("grid=%3d mode=%d N=%3d: %8.1f cycles/MMA (bf16-rate floor %d)  %s\n",grid,mode,N,(double)h/nmma,128*N/256,cudaGetErrorString(e));
  }
  return 0;
}
