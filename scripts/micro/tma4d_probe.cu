// Probe: which 4-D TMA box loads of an NCHW fp32 image (view W,C,H,B) are legal with SWIZZLE_128B_ATOM_32B?
//   tma4d_probe <swizzle enum> <x> <y> <RH> <promo enum> <W> <H>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint64_t* b, uint32_t par){ uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(b)),"r"(par):"memory"); return ok; }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int rh, float* out){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem+1023)&~(uintptr_t)1023);
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&bar)),"r"(rh*1024):"memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(s32(base)),"l"(reinterpret_cast<uint64_t>(&tmap)),"r"(x),"r"(0),"r"(y),"r"(0),"r"(s32(&bar)):"memory");
  }
  while(!tryw(&bar,0)){}
  for (int i=threadIdx.x;i<rh*256;i+=blockDim.x) out[i]=((float*)base)[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv){
  int mode=atoi(argv[1]), x=atoi(argv[2]), y=atoi(argv[3]), rh=atoi(argv[4]), promo=atoi(argv[5]), W=atoi(argv[6]), H=atoi(argv[7]);
  const int C=8;
  float* g; cudaMalloc(&g,(size_t)W*H*C*4); float* hsrc=(float*)malloc((size_t)W*H*C*4);
  for (int i=0;i<W*H*C;++i) hsrc[i]=(float)i; cudaMemcpy(g,hsrc,(size_t)W*H*C*4,cudaMemcpyHostToDevice);
  float* d; cudaMalloc(&d,rh*1024);
  void* fn=nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fn,cudaEnableDefault,&q);
  CUtensorMap tm; cuuint64_t dims[4]={(cuuint64_t)W,C,(cuuint64_t)H,1}; cuuint64_t strides[3]={(cuuint64_t)H*W*4,(cuuint64_t)W*4,(cuuint64_t)C*H*W*4};
  cuuint32_t box[4]={32,8,(cuuint32_t)rh,1}; cuuint32_t es[4]={1,1,1,1};
  CUresult r=((EncodeFn)fn)(&tm,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,4,g,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,(CUtensorMapSwizzle)mode,(CUtensorMapL2promotion)promo,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r){ printf("mode=%d x=%d y=%d rh=%d promo=%d W=%d H=%d: encode rc=%d\n",mode,x,y,rh,promo,W,H,(int)r); return 0; }
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,rh*1024+1024);
  k<<<1,128,rh*1024+1024>>>(tm,x,y,rh,d); cudaError_t e=cudaDeviceSynchronize();
  float h4[4]={0,0,0,0}; if (e==cudaSuccess) cudaMemcpy(h4,d,16,cudaMemcpyDeviceToHost);
  printf("mode=%d x=%d y=%d rh=%d promo=%d W=%d H=%d: %s  first words %g %g %g %g\n",mode,x,y,rh,promo,W,H,cudaGetErrorString(e),h4[0],h4[1],h4[2],h4[3]);
  return 0;
}
