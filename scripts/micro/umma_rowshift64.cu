// Micro-test (SWIZZLE_64B variant of umma_rowshift.cu): K-major SWIZZLE_64B A operand (kind::tf32, 16 channels = 64-byte rows) whose descriptor start address is
// shifted by whole rows (128 B) inside a larger tile written with the address-based 128B swizzle TMA uses.  If the MMA
// swizzles on absolute shared-memory address bits, a 3x3 conv's kw taps can read one halo tile at +0/+1/+2 pixels.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint64_t* b, uint32_t par){ uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(b)),"r"(par):"memory"); return ok; }
__host__ __device__ inline float aval(int m, int k){ return (float)(((m*3 + k*7) % 11) - 5); }
__host__ __device__ inline float bval(int k, int n){ return (float)(((k*5 + n*3) % 7) - 3); }
constexpr int N = 32, ROWS = 160;
__global__ void k(int shift_rows, int base_offset, float* out){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  unsigned char* A = base; unsigned char* B = base + ROWS*64;   // B: 32 rows (n) x 64 B (k = 16 tf32), K-major SW64
  for (int i=threadIdx.x;i<(ROWS*64+4096)/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  __syncthreads();
  for (int i=threadIdx.x;i<ROWS*16;i+=blockDim.x){ int m=i/16,kk=i%16;
    uint32_t off=m*64 + kk*4; off ^= ((off>>7)&3u)<<4;               // Swizzle<2,4,3> on the address (tile base is 1024-aligned)
    *(float*)(A+off) = aval(m,kk); }
  for (int i=threadIdx.x;i<N*16;i+=blockDim.x){ int n=i/16,kk=i%16;
    uint32_t off=n*64 + kk*4; off ^= ((off>>7)&3u)<<4;
    *(float*)(B+off) = bval(kk,n); }
  int warp=threadIdx.x>>5;
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&done))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(32u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc=(1u<<4)|(2u<<7)|(2u<<10)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  auto mk=[&](uint32_t a, int bo){ return (uint64_t)((a>>4)&0x3FFF)|(1ull<<16)|(32ull<<32)|(1ull<<46)|((uint64_t)(bo&7)<<49)|(4ull<<61); };
  if (threadIdx.x==32){
    for (int ks=0; ks<2; ++ks){   // K = 16 = 2 steps of 8 tf32 (32 bytes each)
      uint64_t da=mk(s32(A)+shift_rows*64+ks*32, base_offset), db=mk(s32(B)+ks*32, 0);
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(da),"l"(db),"r"(idesc),"r"((uint32_t)(ks>0)):"memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(s32(&done)):"memory");
  }
  while(!tryw(&done,0)){}
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32]; uint32_t taddr = tm + ((uint32_t)(warp*32)<<16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    :"=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
     "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
    :"r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j=0;j<32;++j) out[threadIdx.x*32+j]=__uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(32u));
}
int main(int argc, char** argv){
  int shift = argc>1?atoi(argv[1]):0, bo = argc>2?atoi(argv[2]):0;
  float* d; cudaMalloc(&d,128*32*4); static float h[128*32];
  int smem = ROWS*64+4096+1024;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  k<<<1,128,smem>>>(shift,bo,d); cudaError_t e=cudaDeviceSynchronize();
  cudaMemcpy(h,d,sizeof(h),cudaMemcpyDeviceToHost);
  int bad=0; for(int m=0;m<128;++m) for(int n=0;n<N;++n){ float r=0; for(int kk=0;kk<16;++kk) r+=aval(m+shift,kk)*bval(kk,n); if (r!=h[m*32+n]) ++bad; }
  printf("A start shifted by %d rows (128 B each), descriptor base_offset %d: %d / %d mismatches (%s)\n",shift,bo,bad,128*N,cudaGetErrorString(e));
  return 0;
}
