// decnet_b200/csrc/conv3d_tcgen05.cu -- coarse 3-D cost aggregation as a bf16 implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Replaces one Conv3dUnit of CostRegNetNoDown (modules/submodule.py:90-123, 608-662):
//   Conv3d 3x3x3, pad 1, no bias  ->  BatchNorm3d (eval)  ->  ReLU  [-> + residual]
// with the BN scale folded into the weights and the BN shift applied in the epilogue.
//
// GEMM view (per layer):  Out[m, n] = sum_{tap, k} A_tap[m, k] * W[tap][n, k]
//   m   : output voxel (b, d, h, w)                      M = B*D*H*W      (5760 per SceneFlow pair)
//   n   : output channel, padded to NP (multiple of 16)  N = 224 (216)    or 16 (last layer, 1)
//   k   : input channel, padded to CP (multiple of 16)   K = 27 * 224
// Layouts: activations bf16 channels-last [B, D, H, W, CP]; weights bf16 [27][NP][CP] (K-major).
//
// One CTA = one 128-voxel tile, shaped as a (bw x bh x bd) box of the volume (bw*bh*bd = 128) so
// that the A operand of tap (kd,kh,kw) is ONE 5-D TMA box {64 ch, bw, bh, bd, 1} of the un-padded
// activation tensor at the shifted (signed) coordinates: the conv's zero padding is TMA's
// out-of-bounds zero fill, no im2col buffer, no halo copies.  K is walked as 27 taps x ceil(CP/64)
// chunks of 64 channels (128-byte rows, SWIZZLE_128B on both the TMA and the UMMA descriptor);
// the channel tail (224 = 3*64 + 32) is zero-filled by TMA and only its valid K-steps are issued.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2-5 = epilogue (TMEM -> registers -> bias/ReLU/residual ->
// bf16 global stores).  4-stage smem ring with full/empty mbarriers; tcgen05.commit releases a
// stage when the MMAs that read it have retired and signals the epilogue at the end.
#include "common.cuh"
#include "tma_utils.cuh"
#include <cuda_bf16.h>
#include <cstring>

namespace decnet {
namespace conv3d {

constexpr int kStages = 4;
constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 channels per stage row = 128 bytes
constexpr int kABytes = kTileM * kChunkK * 2;     // 16 KB

__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    // K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart
    // (cute::UMMA::SmemDescriptor: start>>4 | LBO=1<<16 | SBO=64<<32 | version=1<<46 | layout=2<<61)
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t make_idesc_bf16_f32(int M, int N) {
    // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major A and B,
    // n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
    const float *bias;                 // [NP] BN shift (0 in the padding)
    const __nv_bfloat16 *residual;     // [M][NP] or null
    __nv_bfloat16 *out_bf16;           // [M][NP]          (mode 0)
    float *out_f32;                    // [M] channel 0     (mode 1: the 216->1 layer)
    int B, D, H, W;
    int cp, np;                        // padded in / out channels
    int nchunks, last_ksteps;          // ceil(cp/64), K-steps of 16 in the last chunk
    int bw, bh, bd, tw, th, td;        // tile box and tiles per axis
    int relu, mode, tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
conv3d_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.np * kChunkK * 2;
    const int stage_bytes = kABytes + b_bytes;

    // tile -> origin in the volume
    int t = blockIdx.x;
    const int twi = t % p.tw; t /= p.tw;
    const int thi = t % p.th; t /= p.th;
    const int tdi = t % p.td; t /= p.td;
    const int b = t;
    const int w0 = twi * p.bw, h0 = thi * p.bh, d0 = tdi * p.bd;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int total_iters = 27 * p.nchunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int tap = 0; tap < 27; ++tap) {
                const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                for (int ck = 0; ck < p.nchunks; ++ck, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    unsigned char *sa = base + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    tma_load_5d(sa, &tmA, ck * kChunkK, w0 + kw - 1, h0 + kh - 1, d0 + kd - 1, b, &full_bar[s]);
                    tma_load_3d(sa + kABytes, &tmB, ck * kChunkK, 0, tap, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16_f32(kTileM, p.np);
            int it = 0;
            for (int tap = 0; tap < 27; ++tap) {
                for (int ck = 0; ck < p.nchunks; ++ck, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
                    const uint64_t da = make_smem_desc_sw128(sa);
                    const uint64_t db = make_smem_desc_sw128(sa + kABytes);
                    const int ksteps = (ck == p.nchunks - 1) ? p.last_ksteps : (kChunkK / 16);
                    for (int k = 0; k < ksteps; ++k) {
                        // +32 bytes along K inside the 128-byte swizzle row = +2 in the 16-byte address field
                        umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);           // frees the stage once these MMAs have read it
                }
            }
            umma_commit(&tmem_full_bar);                  // accumulator complete
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                           // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                      // accumulator row = voxel within the tile
        const int dw = r % p.bw, dh = (r / p.bw) % p.bh, dd = r / (p.bw * p.bh);
        const int w = w0 + dw, h = h0 + dh, d = d0 + dd;
        const bool valid = (w < p.W) && (h < p.H) && (d < p.D);
        const size_t m = (((size_t)b * p.D + d) * p.H + h) * p.W + w;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        if (p.mode == 1) {
            float v[16];
            tmem_ld16(trow, v);
            if (valid) { float x = v[0] + p.bias[0]; if (p.relu) x = fmaxf(x, 0.f); p.out_f32[m] = x; }
        } else {
            for (int c0 = 0; c0 < p.np; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);        // warp-collective: every lane takes part
                if (valid) {
                    __align__(16) __nv_bfloat16 o[16];
                    __align__(16) __nv_bfloat16 rs[16];
                    if (p.residual) {
                        const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + m * p.np + c0);
                        *reinterpret_cast<uint4 *>(rs) = __ldg(rp);
                        *reinterpret_cast<uint4 *>(rs + 8) = __ldg(rp + 1);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x = v[i] + __ldg(p.bias + c0 + i);
                        if (p.relu) x = fmaxf(x, 0.f);
                        if (p.residual) x += __bfloat162float(rs[i]);
                        o[i] = __float2bfloat16(x);
                    }
                    uint4 *op = reinterpret_cast<uint4 *>(p.out_bf16 + m * p.np + c0);
                    op[0] = *reinterpret_cast<const uint4 *>(o);
                    op[1] = *reinterpret_cast<const uint4 *>(o + 8);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
    (void)total_iters;
}

// pick the (bw, bh, bd) power-of-two box with bw*bh*bd = 128 that wastes the fewest voxels
static void pick_tile(int W, int H, int D, int &bw, int &bh, int &bd) {
    double best = 1e30;
    bw = 128; bh = 1; bd = 1;
    for (int a = 1; a <= 128; a <<= 1)
        for (int c = 1; a * c <= 128; c <<= 1) {
            const int e = 128 / (a * c);
            const double cover = (double)((W + a - 1) / a * a) * ((H + c - 1) / c * c) * ((D + e - 1) / e * e);
            // prefer wider boxes on ties (longer contiguous runs for TMA and the epilogue stores)
            const double score = cover - 1e-3 * a;
            if (score < best) { best = score; bw = a; bh = c; bd = e; }
        }
}

}  // namespace conv3d
}  // namespace decnet

using namespace decnet;
using namespace decnet::conv3d;

extern "C" {

int decnet_conv3d_bf16(const void *x_ndhwc, const void *w_packed, const float *bias, const void *residual,
                       void *out, int out_mode, int B, int D, int H, int W, int cp, int np, int relu, void *stream)
{
    DECNET_REQUIRE(x_ndhwc && w_packed && bias && out, "null pointer");
    DECNET_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "non-positive size");
    DECNET_REQUIRE(cp % 16 == 0 && cp >= 16 && cp <= 1024, "cp=%d must be a multiple of 16", cp);
    DECNET_REQUIRE(np % 16 == 0 && np >= 16 && np <= 256, "np=%d must be a multiple of 16 in [16,256]", np);
    DECNET_REQUIRE(out_mode == 0 || out_mode == 1, "out_mode must be 0 (bf16 [M][np]) or 1 (fp32 [M], channel 0)");
    DECNET_REQUIRE((reinterpret_cast<uintptr_t>(x_ndhwc) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(residual) & 15u) == 0, "pointers must be 16-byte aligned");
    Params p{};
    p.bias = bias;
    p.residual = static_cast<const __nv_bfloat16 *>(residual);
    p.out_bf16 = out_mode == 0 ? static_cast<__nv_bfloat16 *>(out) : nullptr;
    p.out_f32 = out_mode == 1 ? static_cast<float *>(out) : nullptr;
    p.B = B; p.D = D; p.H = H; p.W = W; p.cp = cp; p.np = np;
    p.nchunks = (cp + kChunkK - 1) / kChunkK;
    p.last_ksteps = (cp - (p.nchunks - 1) * kChunkK) / 16;
    pick_tile(W, H, D, p.bw, p.bh, p.bd);
    p.tw = (W + p.bw - 1) / p.bw; p.th = (H + p.bh - 1) / p.bh; p.td = (D + p.bd - 1) / p.bd;
    p.relu = relu; p.mode = out_mode;
    p.tmem_cols = np <= 32 ? 32 : np <= 64 ? 64 : np <= 128 ? 128 : 256;

    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[5] = {(uint64_t)cp, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)B};
        const uint64_t strides[4] = {(uint64_t)cp * 2, (uint64_t)W * cp * 2, (uint64_t)H * W * cp * 2,
                                     (uint64_t)D * H * W * cp * 2};
        const uint32_t box[5] = {(uint32_t)kChunkK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, 1u};
        int rc = encode_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x_ndhwc, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)cp, (uint64_t)np, 27ull};
        const uint64_t strides[2] = {(uint64_t)cp * 2, (uint64_t)np * cp * 2};
        const uint32_t box[3] = {(uint32_t)kChunkK, (uint32_t)np, 1u};
        int rc = encode_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_packed, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    const size_t smem = (size_t)kStages * (kABytes + (size_t)np * kChunkK * 2) + 1024;
    DECNET_CUDA(cudaFuncSetAttribute(conv3d_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (long long)B * p.tw * p.th * p.td;
    DECNET_REQUIRE(tiles < (1ll << 31), "too many tiles");
    conv3d_tcgen05_kernel<<<(unsigned)tiles, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
    return after_launch("conv3d_tcgen05_kernel");
}

}  // extern "C"
