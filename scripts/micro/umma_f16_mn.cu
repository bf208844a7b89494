// Micro-test: tcgen05.mma kind::f16 with MN-major fp16 A and B whose MN atoms are 32 elements (64 B) -- the layout of the thin
// NCHW conv's fp16 correction operand (32 pixels per image row, 8 channels per K atom).  Checks which (layout_type, swizzle,
// LBO, SBO) hypothesis reproduces A*B exactly for M = 128 (4 image rows), K = 16 (2 K atoms), N = 32, and that a kind::tf32
// MMA and a kind::f16 MMA may accumulate into the same TMEM columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/micro/umma_f16_mn scripts/micro/umma_f16_mn.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint64_t* b, uint32_t par){ uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(b)),"r"(par):"memory"); return ok; }
__host__ __device__ inline float aval(int m, int k){ return (float)(((m*3 + k*7) % 11) - 5); }
__host__ __device__ inline float bval(int k, int n){ return (float)(((k*5 + n*3) % 7) - 3); }
// swz: 0 none, 1 = Swizzle<2,4,3> (bits 4..5 ^= bits 7..8: 64B), 2 = Swizzle<3,4,3> (128B), 3 = Swizzle<1,4,3> (32B), 4 = Swizzle<2,5,2>
__device__ __forceinline__ uint32_t swizzle(uint32_t off, int swz){
  if (swz==1) return off ^ (((off>>7)&3u)<<4);
  if (swz==2) return off ^ (((off>>7)&7u)<<4);
  if (swz==3) return off ^ (((off>>7)&1u)<<4);
  if (swz==4) return off ^ (((off>>7)&3u)<<5);
  return off;
}
struct Hyp { int layout_type, swz, mn_stride, k_stride, dlbo, dsbo, mix; };
__global__ void k(Hyp h, int N, float* out){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  unsigned char* A = base; unsigned char* B = base + 8192; unsigned char* A32 = base + 16384; unsigned char* B32 = base + 24576;
  for (int i=threadIdx.x;i<32768/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  __syncthreads();
  // fp16 element (mn, k): MN atom = 32 elements (64 B), K atom = 8 rows of 64 B
  for (int i=threadIdx.x;i<128*16;i+=blockDim.x){ int m=i%128,kk=i/128;
    uint32_t off=(m/32)*h.mn_stride + (kk/8)*h.k_stride + (kk%8)*64 + (m%32)*2;
    *(__half*)(A+swizzle(off,h.swz)) = __float2half(aval(m,kk)); }
  for (int i=threadIdx.x;i<N*16;i+=blockDim.x){ int n=i%N,kk=i/N;
    uint32_t off=(n/32)*h.mn_stride + (kk/8)*h.k_stride + (kk%8)*64 + (n%32)*2;
    *(__half*)(B+swizzle(off,h.swz)) = __float2half(bval(kk,n)); }
  // tf32 MN-major operands (SWIZZLE_128B_BASE32B, validated in umma_tf32_mn.cu): K = 8, values aval(m, 16+k) / bval(16+k, n)
  for (int i=threadIdx.x;i<128*8;i+=blockDim.x){ int m=i%128,kk=i/128;
    uint32_t off=(m/32)*1024 + (kk/4)*512 + (kk%4)*128 + (m%32)*4;
    *(float*)(A32+swizzle(off,4)) = aval(m,16+kk); }
  for (int i=threadIdx.x;i<N*8;i+=blockDim.x){ int n=i%N,kk=i/N;
    uint32_t off=(n/32)*1024 + (kk/4)*512 + (kk%4)*128 + (n%32)*4;
    *(float*)(B32+swizzle(off,4)) = bval(16+kk,n); }
  int warp=threadIdx.x>>5;
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&done))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(64u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc16=(1u<<4)|(1u<<15)|(1u<<16)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  uint32_t idesc32=(1u<<4)|(2u<<7)|(2u<<10)|(1u<<15)|(1u<<16)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  auto mk=[&](uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|((uint64_t)((h.dlbo>>4)&0x3FFF)<<16)|((uint64_t)((h.dsbo>>4)&0x3FFF)<<32)|(1ull<<46)|((uint64_t)h.layout_type<<61); };
  auto mk32=[&](uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|(64ull<<16)|(32ull<<32)|(1ull<<46)|(1ull<<61); };
  if (threadIdx.x==32){
    uint32_t acc = 0;
    if (h.mix){
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(mk32(s32(A32))),"l"(mk32(s32(B32))),"r"(idesc32),"r"(0u):"memory");
      acc = 1;
    }
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"::"r"(tm),"l"(mk(s32(A))),"l"(mk(s32(B))),"r"(idesc16),"r"(acc):"memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(s32(&done)):"memory");
  }
  while(!tryw(&done,0)){}
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  uint32_t taddr = tm + ((uint32_t)(warp*32)<<16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    :"=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
     "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
    :"r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j=0;j<32;++j) out[threadIdx.x*32+j]=__uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(64u));
}
int main(){
  float* d; cudaMalloc(&d,128*32*4); int smem=32768+1024;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  const int N=32;
  Hyp hyps[] = {
    {4,1,1024,512,1024,512,0},   // SWIZZLE_64B, Swizzle<2,4,3>, LBO = mn-atom stride, SBO = k-atom stride
    {4,1,1024,512,512,1024,0},   // the same with LBO/SBO swapped
    {4,1,1024,512,1024,512,1},   // ... plus a tf32 MMA into the same accumulator first
    {0,0,1024,512,1024,512,0},   // no swizzle
    {2,2,1024,512,1024,512,0},   // 128B swizzle pattern on 64-byte rows
    {6,3,1024,512,1024,512,0},   // 32B
    {4,1,512,2048,512,2048,0},   // mn atoms packed (512 B apart), k atoms 2 KB apart
  };
  static float h[128*32];
  for (auto& hy : hyps){
    cudaMemset(d,0,sizeof(h));
    k<<<1,128,smem>>>(hy,N,d); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(h,d,sizeof(h),cudaMemcpyDeviceToHost);
    int bad=0; for(int m=0;m<128;++m) for(int n=0;n<N;++n){ float r=0; for(int kk=0;kk<(hy.mix?24:16);++kk) r+=aval(m,kk)*bval(kk,n); if (r!=h[m*32+n]) ++bad; }
    printf("layout_type=%d swz=%d mn_stride=%d k_stride=%d dLBO=%d dSBO=%d mix=%d : %d / %d mismatches  (%s)  d[0][0..3]=%g %g %g %g\n",hy.layout_type,hy.swz,hy.mn_stride,hy.k_stride,hy.dlbo,hy.dsbo,hy.mix,bad,128*N,cudaGetErrorString(e),h[0],h[1],h[2],h[3]);
    if (e!=cudaSuccess) return 1;
  }
  return 0;
}
