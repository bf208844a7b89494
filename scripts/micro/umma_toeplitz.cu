// Micro-test for the "pixels on N" formulation: A = weights, K-major NO-swizzle ([8-row group][k half][8 rows][16 B],
// LBO = 128 B between the two K halves, SBO = 256 B between 8-row groups), B = image tile MN-major SWIZZLE_128B_BASE32B
// with N = 96 (three 32-pixel atoms, LBO = atom stride), kind::tf32.  Also: A start shifted by whole 8-row groups
// (the block-Toeplitz trick) and an even D column offset.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint64_t* b, uint32_t par){ uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(b)),"r"(par):"memory"); return ok; }
__host__ __device__ inline float aval(int m, int k){ return (float)(((m*3 + k*7) % 11) - 5); }
__host__ __device__ inline float bval(int k, int n){ return (float)(((k*5 + n*3) % 7) - 3); }
constexpr int N = 96, ROWS = 192;           // A buffer has 192 rows so that shifted starts stay inside
__global__ void k(int shift_groups, int dcol, int lbo, int sbo, float* out, int ldoff){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  unsigned char* A = base; unsigned char* B = base + 8192;
  for (int i=threadIdx.x;i<(8192+4096)/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  __syncthreads();
  for (int i=threadIdx.x;i<ROWS*8;i+=blockDim.x){ int m=i/8,kk=i%8;
    uint32_t off=(m/8)*sbo + (kk/4)*lbo + (m%8)*16 + (kk%4)*4;
    *(float*)(A+off) = aval(m,kk); }
  for (int i=threadIdx.x;i<N*8;i+=blockDim.x){ int n=i%N,kk=i/N;
    uint32_t off=(n/32)*1024 + (kk/4)*512 + (kk%4)*128 + (n%32)*4;
    off ^= ((off>>7)&3u)<<5;
    *(float*)(B+off) = bval(kk,n); }
  int warp=threadIdx.x>>5;
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&done))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(128u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc=(1u<<4)|(2u<<7)|(2u<<10)|(0u<<15)|(1u<<16)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  uint32_t a_addr = s32(A) + shift_groups*sbo;
  uint64_t da=(uint64_t)((a_addr>>4)&0x3FFF)|((uint64_t)((lbo>>4)&0x3FFF)<<16)|((uint64_t)((sbo>>4)&0x3FFF)<<32)|(1ull<<46)|(0ull<<61);
  uint64_t db=(uint64_t)((s32(B)>>4)&0x3FFF)|(64ull<<16)|(32ull<<32)|(1ull<<46)|(1ull<<61);
  if (threadIdx.x==32){
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"::"r"(tm+(uint32_t)dcol),"l"(da),"l"(db),"r"(idesc),"r"(0u):"memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(s32(&done)):"memory");
  }
  while(!tryw(&done,0)){}
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0=0;c0<96;c0+=32){
    uint32_t v[32]; uint32_t taddr = tm + ((uint32_t)(warp*32)<<16) + c0 + ldoff;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      :"=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
       "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
      :"r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j=0;j<32;++j) out[threadIdx.x*128+c0+j]=__uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(128u));
}
int main(int argc, char** argv){
  int shift = argc>1?atoi(argv[1]):0, dcol = argc>2?atoi(argv[2]):0, lbo = argc>3?atoi(argv[3]):128, sbo = argc>4?atoi(argv[4]):256, ldoff = argc>5?atoi(argv[5]):0;
  float* d; cudaMalloc(&d,128*128*4); static float h[128*128];
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,16384);
  k<<<1,128,16384>>>(shift,dcol,lbo,sbo,d,ldoff); cudaError_t e=cudaDeviceSynchronize();
  cudaMemcpy(h,d,sizeof(h),cudaMemcpyDeviceToHost);
  int bad=0; for(int m=0;m<128;++m) for(int n=ldoff;n<N-32;++n){ float r=0; for(int kk=0;kk<8;++kk) r+=aval(m+8*shift,kk)*bval(kk,n); if (r!=h[m*128+dcol+n-ldoff]) ++bad; }
  printf("A K-major no-swizzle (LBO %d, SBO %d), start shifted by %d groups, D column offset %d, tcgen05.ld column offset %d: %d mismatches (%s)\n",lbo,sbo,shift,dcol,ldoff,bad,cudaGetErrorString(e));
  return 0;
}
