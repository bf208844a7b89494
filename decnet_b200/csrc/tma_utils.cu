// decnet_b200/csrc/tma_utils.cu -- host-side cuTensorMapEncodeTiled through
// cudaGetDriverEntryPoint (keeps libdecnet_b200.so loadable on machines without libcuda).
#include "tma_utils.cuh"
#include <mutex>
#include <cstring>

namespace decnet {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static std::mutex mu;
    static EncodeTiledFn fn = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Encoding costs ~10 us of host time per map; the same (pointer, shape, box) recurs every step
// (ping-pong activation buffers, static weights), so a small thread-local cache removes it.
struct MapKey {
    const void *addr; int dtype, rank, swizzle, promo;
    uint64_t dims[5], strides[4]; uint32_t box[5];
    bool operator==(const MapKey &o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapEntry { MapKey key; CUtensorMap map; bool used; };
static thread_local MapEntry g_cache[64];
static thread_local int g_cache_next = 0;

static int encode_uncached(CUtensorMap *out, CUtensorMapDataType dtype, int rank, const void *gaddr,
                           const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
                           CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo);

int encode_tensor_map(CUtensorMap *out, CUtensorMapDataType dtype, int rank, const void *gaddr,
                      const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
                      CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo)
{
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.addr = gaddr; k.dtype = (int)dtype; k.rank = rank; k.swizzle = (int)swizzle; k.promo = (int)promo;
    for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides_bytes[i];
    for (int i = 0; i < 64; ++i)
        if (g_cache[i].used && g_cache[i].key == k) { *out = g_cache[i].map; return 0; }
    int rc = encode_uncached(out, dtype, rank, gaddr, dims, strides_bytes, box, swizzle, promo);
    if (rc) return rc;
    MapEntry &e = g_cache[g_cache_next];
    g_cache_next = (g_cache_next + 1) % 64;
    e.key = k; e.map = *out; e.used = true;
    return 0;
}

static int encode_uncached(CUtensorMap *out, CUtensorMapDataType dtype, int rank, const void *gaddr,
                           const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
                           CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DECNET_ERR_UNSUPPORTED; }
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void *>(gaddr), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
        return DECNET_ERR_CUDA_BASE + 999;
    }
    return 0;
}

}  // namespace decnet
