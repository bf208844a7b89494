// Microbenchmark 2: what does a "stage" cost the MMA issuer?  4 MMAs + optional commit / barrier wait / fence,
// with other warps of the CTA optionally spin-waiting on an mbarrier (like epilogue warps do).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|(1ull<<16)|(64ull<<32)|(1ull<<46)|(2ull<<61); }
__device__ __forceinline__ bool tryw(uint64_t* b, uint32_t par){ uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(b)),"r"(par):"memory"); return ok; }
__global__ void k(int N, int nstage, int flags, long long* out){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[8]; __shared__ __align__(8) uint64_t done; __shared__ __align__(8) uint64_t never; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  for (int i=threadIdx.x;i<(16384+32768)/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  int warp=threadIdx.x>>5, lane=threadIdx.x&31;
  if (threadIdx.x==0){ for(int i=0;i<8;++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar[i])));
     asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&done))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&never)));
     asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc=(1u<<4)|(1u<<7)|(1u<<10)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  uint64_t da=desc(s32(base)), db=desc(s32(base+16384));
  if (warp==1){
    if (lane==0){
      long long t0=clock64();
      for(int st=0;st<nstage;++st){
        int s = st % 5;
        if (flags&2){ uint32_t par=((st/5)&1)^1; while(!tryw(&bar[s], par)){} }   // wait "empty"-like barrier previously committed
        if (flags&4) asm volatile("tcgen05.fence::after_thread_sync;":::"memory");
        for(int k=0;k<4;++k){ uint32_t acc=(st|k)>0;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(da+2*k),"l"(db+2*k),"r"(idesc),"r"(acc):"memory"); }
        if (flags&1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(s32(&bar[s])):"memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(s32(&done)):"memory");
      while(!tryw(&done,0)){}
      long long t1=clock64();
      if (blockIdx.x==0) out[0]=t1-t0;
    }
  } else if (warp>=2 && (flags&8)) {
    // epilogue-like warps: spin on a barrier that completes only at the end
    while(!tryw(&done,0)){}
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(512u));
}
int main(){
  long long* d; cudaMalloc(&d,8); int smem=16384+32768+1024;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  const char* names[]={"mma only","+commit/stage","+wait","+commit+wait","+fence","+commit+wait+fence"};
  int fl[]={0,1,2,3,4,7};
  for (int spin=0;spin<2;++spin) for (int N : {16,224}) for (int i=0;i<6;++i){
    if (fl[i]&2 && !(fl[i]&1)) continue;
    int flags=fl[i]|(spin?8:0); int nstage=512; long long h=0;
    k<<<148,192,smem>>>(N,nstage,flags,d); cudaDeviceSynchronize();
    k<<<148,192,smem>>>(N,nstage,flags,d); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(&h,d,8,cudaMemcpyDeviceToHost);
    printf("spin_warps=%d N=%3d %-22s: %8.1f cycles/stage (floor %d)  %s\n",spin,N,names[i],(double)h/nstage,4*(128*N/256>40?128*N/256:40),cudaGetErrorString(e));
  }
  return 0;
}
