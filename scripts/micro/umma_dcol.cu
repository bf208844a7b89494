// Micro-test: tcgen05.mma kind::tf32 with MN-major A and B (the layout an NCHW fp32 image gives when pixels
// are the GEMM M dimension).  Checks which (layout_type, swizzle, LBO, SBO) hypothesis reproduces A*B exactly,
// and dumps what TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes so the two can be matched.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_tf32_mn scripts/micro/umma_tf32_mn.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint64_t* b, uint32_t par){ uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(b)),"r"(par):"memory"); return ok; }

__host__ __device__ inline float aval(int m, int k){ return (float)(((m*3 + k*7) % 11) - 5); }
__host__ __device__ inline float bval(int k, int n){ return (float)(((k*5 + n*3) % 7) - 3); }

// swz: 0 none, 1 = Swizzle<2,5,2> (bits 5..6 ^= bits 7..8), 2 = Swizzle<3,4,3> (bits 4..6 ^= bits 7..9)
__device__ __forceinline__ uint32_t swizzle(uint32_t off, int swz){
  if (swz==1) return off ^ (((off>>7)&3u)<<5);
  if (swz==2) return off ^ (((off>>7)&7u)<<4);
  return off;
}

struct Hyp { int layout_type, swz, krows, lbo, sbo, dlbo, dsbo; };

__global__ void k(Hyp h, int N, float* out, int dcol){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done; __shared__ uint32_t tslot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  unsigned char* A = base; unsigned char* B = base + 8192;
  for (int i=threadIdx.x;i<16384/4;i+=blockDim.x) ((uint32_t*)base)[i]=0;
  __syncthreads();
  // element (mn, k): atom along MN = 32 elements (128 B), atom along K = h.krows rows of 128 B
  for (int i=threadIdx.x;i<128*8;i+=blockDim.x){ int m=i%128,kk=i/128;
    uint32_t off=(m/32)*h.lbo + (kk/h.krows)*h.sbo + (kk%h.krows)*128 + (m%32)*4;
    *(float*)(A+swizzle(off,h.swz)) = aval(m,kk); }
  for (int i=threadIdx.x;i<N*8;i+=blockDim.x){ int n=i%N,kk=i/N;
    uint32_t off=(n/32)*h.lbo + (kk/h.krows)*h.sbo + (kk%h.krows)*128 + (n%32)*4;
    *(float*)(B+swizzle(off,h.swz)) = bval(kk,n); }
  int warp=threadIdx.x>>5;
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&done))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp==0){ asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"::"r"(s32(&tslot)),"r"(64u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t idesc=(1u<<4)|(2u<<7)|(2u<<10)|(1u<<15)|(1u<<16)|((uint32_t)(N>>3)<<17)|((uint32_t)(128>>4)<<24);
  auto mk=[&](uint32_t a){ return (uint64_t)((a>>4)&0x3FFF)|((uint64_t)((h.dlbo>>4)&0x3FFF)<<16)|((uint64_t)((h.dsbo>>4)&0x3FFF)<<32)|(1ull<<46)|((uint64_t)h.layout_type<<61); };
  uint64_t da=mk(s32(A)), db=mk(s32(B));
  if (threadIdx.x==32){
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"::"r"(tm+(uint32_t)dcol),"l"(da),"l"(db),"r"(idesc),"r"(0u):"memory");
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
  for (int j=0;j<32;++j) out[threadIdx.x*64+j]=__uint_as_float(v[j]);
  taddr += 32;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    :"=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
     "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
    :"r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j=0;j<32;++j) out[threadIdx.x*64+32+j]=__uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp==0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"::"r"(tm),"r"(64u));
}

// TMA dump: loads an [rows][32] fp32 tile with the given swizzle mode and writes raw smem back.
__global__ void tma_dump(const __grid_constant__ CUtensorMap tmap, int rows, float* out){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&bar)),"r"(rows*128):"memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(s32(base)),"l"(reinterpret_cast<uint64_t>(&tmap)),"r"(0),"r"(0),"r"(s32(&bar)):"memory");
  }
  while(!tryw(&bar,0)){}
  for (int i=threadIdx.x;i<rows*32;i+=blockDim.x) out[i]=((float*)base)[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv){ int only = argc>1 ? atoi(argv[1]) : -1;
  float* d; cudaMalloc(&d,128*64*4); int smem=16384+1024;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  const int N=32;
  Hyp hyps[] = {
    {1,1,4,1024,512,1024,512},   // BASE32B, Swizzle<2,5,2>, k-atom 4 rows, LBO=mn-atom stride, SBO=k-atom stride
    {1,1,4,1024,512,512,1024},   // same layout, LBO/SBO swapped in the descriptor
    {2,2,8,1024,1024,1024,1024}, // plain SWIZZLE_128B with 8-row atoms
    {0,0,8,1024,1024,1024,1024}, // no swizzle, 8-row atoms
    {1,1,8,1024,1024,1024,1024}, // BASE32B descriptor, 8-row atoms
    {1,1,4,2048,512,2048,512},   // BASE32B with a larger mn-atom stride (as with 16 channels per row)
  };
  static float h[128*64];
  for (auto& hy : hyps){
    cudaMemset(d,0,sizeof(h));
    k<<<1,128,smem>>>(hy,N,d,0); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(h,d,sizeof(h),cudaMemcpyDeviceToHost);
    int bad=0; for(int m=0;m<128;++m) for(int n=0;n<N;++n){ float r=0; for(int kk=0;kk<8;++kk) r+=aval(m,kk)*bval(kk,n); if (r!=h[m*64+n]) ++bad; }
    printf("layout_type=%d swz=%d krows=%d mn_stride=%d k_stride=%d dLBO=%d dSBO=%d : %d / %d mismatches  (%s)  d[0][0..3]=%g %g %g %g\n",hy.layout_type,hy.swz,hy.krows,hy.lbo,hy.sbo,hy.dlbo,hy.dsbo,bad,128*N,cudaGetErrorString(e),h[0],h[1],h[2],h[3]);
    if (e!=cudaSuccess) return 1;
  }
  for (int dcol : {0,1,2,3,4,5,7,8,13,16,31}){ if (only>=0 && dcol!=only) continue;
    cudaMemset(d,0,sizeof(h));
    k<<<1,128,smem>>>(hyps[0],N,d,dcol); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(h,d,sizeof(h),cudaMemcpyDeviceToHost);
    int bad=0; for(int m=0;m<128;++m) for(int n=0;n<N;++n){ float r=0; for(int kk=0;kk<8;++kk) r+=aval(m,kk)*bval(kk,n); if (r!=h[m*64+dcol+n]) ++bad; }
    printf("D column offset %2d: %d / %d mismatches (%s)\n",dcol,bad,128*N,cudaGetErrorString(e));
    if (e!=cudaSuccess) return 1;
  }
  // ---- TMA swizzle dump
  void* fn=nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fn,cudaEnableDefault,&q);
  EncodeFn enc=(EncodeFn)fn;
  const int rows=8; static float src[rows*32]; for(int i=0;i<rows*32;++i) src[i]=(float)i;
  float* g; cudaMalloc(&g,sizeof(src)); cudaMemcpy(g,src,sizeof(src),cudaMemcpyHostToDevice);
  for (int mode : {(int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,(int)CU_TENSOR_MAP_SWIZZLE_128B}){
    CUtensorMap tm; cuuint64_t dims[2]={32,(cuuint64_t)rows}; cuuint64_t strides[1]={128}; cuuint32_t box[2]={32,(cuuint32_t)rows}; cuuint32_t es[2]={1,1};
    CUresult r=enc(&tm,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,2,g,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,(CUtensorMapSwizzle)mode,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("TMA swizzle mode %d encode rc=%d\n",mode,(int)r); if (r) continue;
    cudaMemset(d,0,sizeof(h)); tma_dump<<<1,128,rows*128+1024>>>(tm,rows,d); cudaError_t e=cudaDeviceSynchronize();
    cudaMemcpy(h,d,rows*128,cudaMemcpyDeviceToHost);
    printf("  (%s) smem word -> source index, one row per 128 B; listing every 4th word (16-byte chunk starts)\n",cudaGetErrorString(e));
    for(int r2=0;r2<rows;++r2){ printf("  row %d:",r2); for(int c=0;c<32;c+=4) printf(" %4g",h[r2*32+c]); printf("\n"); }
  }
  return 0;
}
