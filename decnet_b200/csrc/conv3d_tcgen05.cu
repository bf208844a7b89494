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
#include <mutex>

namespace decnet {
namespace conv3d {

constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kRowBytes = 128;                    // one stage row = 128 bytes = 64 bf16 or 32 fp32 (tf32) channels
constexpr int kABytes = kTileM * kRowBytes;       // 16 KB

__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    // K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart
    // (cute::UMMA::SmemDescriptor: start>>4 | LBO=1<<16 | SBO=64<<32 | version=1<<46 | layout=2<<61)
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t make_idesc(int fmt, int M, int N) {
    // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format (BF16 = 1, TF32 = 2) at bits 7 / 10,
    // K-major A and B, n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The whole MMA warp stays converged and elect.sync picks the single issuing lane INSIDE the asm block:
// the descriptors stay in uniform registers.  Issuing from a divergent `if (lane == 0)` region makes
// nvcc wrap every UTCHMMA in an ELECT/BRA.U.ANY loop with R2UR moves (~140 cycles per MMA instead of ~40).
// One stage = `KSTEPS` K-steps of 16 (32 bytes each inside the 128-byte swizzle row = +2 in the
// descriptor's 16-byte address field) followed by the commit that releases the stage: a single asm
// block with ONE elect.sync keeps the issuing warp's dependent instruction chain per stage short.
#define DECNET_MMA_NEXT(OFF)                                                       \
        "add.u64 a, %1, " #OFF ";\n\t"                                              \
        "add.u64 b, %2, " #OFF ";\n\t"                                              \
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
#define DECNET_STAGE_HEAD                                                          \
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b;\n\t"                            \
        "elect.sync _|e, 0xffffffff;\n\t"                                           \
        "setp.ne.b32 p, %4, 0;\n\t"                                                 \
        "setp.eq.b32 t, 0, 0;\n\t"                                                  \
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
#define DECNET_STAGE_TAIL                                                          \
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
template <int KSTEPS>
__device__ __forceinline__ void umma_stage_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate_first, uint32_t bar_addr) {
    static_assert(KSTEPS >= 1 && KSTEPS <= 4, "1..4 K-steps per stage");
    if constexpr (KSTEPS == 4) {
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_MMA_NEXT(4) DECNET_MMA_NEXT(6) DECNET_STAGE_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    } else if constexpr (KSTEPS == 3) {
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_MMA_NEXT(4) DECNET_STAGE_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    } else if constexpr (KSTEPS == 2) {
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_STAGE_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    } else {
        asm volatile(DECNET_STAGE_HEAD DECNET_STAGE_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    }
}
// Two consecutive stages (4 K-steps, then K2 K-steps) behind ONE election: eight MMAs and both commits in one block, so the
// issuing warp pays its wait / fence / descriptor set-up once per 128 channels instead of once per 64.
#define DECNET_MMA_NEXT2(OFF)                                                      \
        "add.u64 a, %6, " #OFF ";\n\t"                                              \
        "add.u64 b, %7, " #OFF ";\n\t"                                              \
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
template <int K2>
__device__ __forceinline__ void umma_stage_pair_elect(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t idesc,
                                                      uint32_t accumulate_first, uint32_t bar0, uint64_t da1, uint64_t db1,
                                                      uint32_t bar1) {
    static_assert(K2 >= 1 && K2 <= 4, "1..4 K-steps in the second stage");
#define DECNET_PAIR_MID                                                            \
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t" \
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %6, %7, %3, t;\n\t"
#define DECNET_PAIR_TAIL                                                           \
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
#define DECNET_PAIR_ARGS ::"r"(tmem_d), "l"(da0), "l"(db0), "r"(idesc), "r"(accumulate_first), "r"(bar0), "l"(da1), "l"(db1), "r"(bar1) : "memory"
    if constexpr (K2 == 4)
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_MMA_NEXT(4) DECNET_MMA_NEXT(6) DECNET_PAIR_MID
                     DECNET_MMA_NEXT2(2) DECNET_MMA_NEXT2(4) DECNET_MMA_NEXT2(6) DECNET_PAIR_TAIL DECNET_PAIR_ARGS);
    else if constexpr (K2 == 3)
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_MMA_NEXT(4) DECNET_MMA_NEXT(6) DECNET_PAIR_MID
                     DECNET_MMA_NEXT2(2) DECNET_MMA_NEXT2(4) DECNET_PAIR_TAIL DECNET_PAIR_ARGS);
    else if constexpr (K2 == 2)
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_MMA_NEXT(4) DECNET_MMA_NEXT(6) DECNET_PAIR_MID
                     DECNET_MMA_NEXT2(2) DECNET_PAIR_TAIL DECNET_PAIR_ARGS);
    else
        asm volatile(DECNET_STAGE_HEAD DECNET_MMA_NEXT(2) DECNET_MMA_NEXT(4) DECNET_MMA_NEXT(6) DECNET_PAIR_MID
                     DECNET_PAIR_TAIL DECNET_PAIR_ARGS);
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}
// kind::tf32 reads fp32 words and drops the low 13 mantissa bits; rounding to nearest beforehand
// (what cuDNN's TF32 path does with cvt.rna) halves the error.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
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
    int relu, mode, tmem_cols;         // mode 0 bf16 [M][np], 1 fp32 [M] (channel 0), 2 fp32 [M][np]; 2 TMEM slots
    int fmt, chunk_ch;                 // operand format (1 bf16 / 2 tf32) and channels per 128-byte stage row
    int taps_d;                        // 3: 3x3x3 taps (Conv3d), 1: 3x3 taps on a D=1 volume (Conv2d)
    int skip_tma;                      // tuning only (variant 3): after the first ring fill, signal stages without loading
    float *out_f32_full;               // [M][np] (mode 2)
    int round_tf32;                    // mode 2: round the stored activations to TF32 (nearest) for the next tf32 conv
    int num_tiles, stages;
    // Wave-quantisation tail: the persistent CTAs walk num_items work items; items [0, tail_first) are whole tiles, the rest
    // are the tail_tiles tiles of the last, partial round split in two along N (np/2 output channels each) when they then
    // still fit one round -- 360 tiles on 148 SMs become 2 full rounds + 128 half tiles instead of 3 rounds.
    int num_items, tail_first, tail_tiles;
    int skip_h_edges;                  // row-band mode: rows h = 0 and h = H-1 are halo slots a NEIGHBOUR rank fills over NVLink
                                       // (or that stay zero at the image edge); this launch must not store into them
    long long *dbg;                    // optional per-CTA timing (decnet_conv3d_debug_timing), else null
    float *pred;                       // mode 1, td == 1: soft-argmin over the tile's disparities fused into the epilogue
                                       // ([B][H][W]; submodule.py:766-777), else null
};

// work item -> (tile, first output channel, number of output channels)
__device__ __forceinline__ void decode_item(const Params &p, int item, int &tile, int &n0, int &nlen) {
    if (item < p.tail_first) { tile = item; n0 = 0; nlen = p.np; return; }
    const int k = item - p.tail_first;
    const int half = k >= p.tail_tiles ? 1 : 0;
    tile = p.tail_first + k - half * p.tail_tiles;
    nlen = p.np >> 1;
    n0 = half * nlen;
}

__global__ void __launch_bounds__(kThreads, 1)
conv3d_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmBh, const Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float cost_s[2][128];   // fused soft-argmin: the tile's costs, double-buffered over tiles

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.np * kRowBytes;
    const int stage_bytes = kABytes + b_bytes;
    const int kStages = p.stages;
    const int acc_stride = p.tmem_cols >> 1;              // two accumulator slots (double buffering)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBh);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
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
    pdl_trigger();                                        // see common.cuh: the next kernel may be scheduled from here on
    const uint32_t tmem_base = tmem_base_slot;

    // Persistent: this CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the smem ring and the
    // two TMEM accumulator slots keep rolling across tiles, so the epilogue of tile j overlaps the
    // main loop of tile j+1.
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            pdl_wait();                                   // first access to the previous kernel's output: the activation tiles
            int s = 0; uint32_t ph = 0;                   // ring slot / phase, advanced without divisions
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int tile, n0, nlen;
                decode_item(p, item, tile, n0, nlen);
                const CUtensorMap *tb = nlen == p.np ? &tmB : &tmBh;
                const uint32_t tx = (uint32_t)(kABytes + nlen * kRowBytes);
                int t = tile;
                const int w0 = (t % p.tw) * p.bw; t /= p.tw;
                const int h0 = (t % p.th) * p.bh; t /= p.th;
                const int d0 = (t % p.td) * p.bd; t /= p.td;
                const int b = t;
                int tap = 0;
                for (int kd = (p.taps_d == 3 ? 0 : 1); kd < (p.taps_d == 3 ? 3 : 2); ++kd)
                    for (int kh = 0; kh < 3; ++kh)
                        for (int kw = 0; kw < 3; ++kw, ++tap)
                            for (int ck = 0; ck < p.nchunks; ++ck) {
                                mbar_wait(&empty_bar[s], ph ^ 1u);
                                unsigned char *sa = base + (size_t)s * stage_bytes;
                                if (p.skip_tma && (ph || item != (int)blockIdx.x)) { mbar_arrive(&full_bar[s]); }
                                else {
                                mbar_arrive_expect_tx(&full_bar[s], tx);
                                tma_load_5d(sa, &tmA, ck * p.chunk_ch, w0 + kw - 1, h0 + kh - 1, d0 + kd - 1, b, &full_bar[s]);
                                tma_load_3d(sa + kABytes, tb, ck * p.chunk_ch, n0, tap, &full_bar[s]);
                                }
                                if (++s == kStages) { s = 0; ph ^= 1u; }
                            }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        {
            const uint32_t idesc_full = make_idesc(p.fmt, kTileM, p.np), idesc_half = make_idesc(p.fmt, kTileM, p.np >> 1);
            int it = 0, j = 0;
            int s = 0; uint32_t ph = 0;                   // ring slot / phase, advanced without divisions
            const uint32_t smem_base = smem_u32(base);
            const uint32_t empty_base = smem_u32(&empty_bar[0]);
            bool ready = false;
            long long t_begin = clock64(), t_wait = 0;
            unsigned long long ns_begin; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
                const uint32_t idesc = item < p.tail_first ? idesc_full : idesc_half;
                const int slot = j & 1;
                mbar_wait(&tmem_empty_bar[slot], ((uint32_t)(j >> 1) & 1u) ^ 1u);   // epilogue drained this slot
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(slot * acc_stride);
                uint32_t first = 1u;
                const int last_ck = p.nchunks - 1;
                if ((p.nchunks & 1) == 0 && kStages >= 4 && !p.dbg) {
                    // even number of chunks: two stages per issue block
                    const int npairs = 9 * p.taps_d * (p.nchunks >> 1);
                    const int pairs_per_tap = p.nchunks >> 1;
                    int pc = 0;                                        // pair index within the tap
                    for (int pr = 0; pr < npairs; ++pr, it += 2) {
                        int s1 = s + 1; uint32_t ph1 = ph;
                        if (s1 == kStages) { s1 = 0; ph1 ^= 1u; }
                        if (!ready) mbar_wait(&full_bar[s], ph);
                        mbar_wait(&full_bar[s1], ph1);
                        tc_fence_after();
                        const uint32_t sa0 = smem_base + (uint32_t)(s * stage_bytes), sa1 = smem_base + (uint32_t)(s1 * stage_bytes);
                        const uint64_t da0 = make_smem_desc_sw128(sa0), db0 = make_smem_desc_sw128(sa0 + kABytes);
                        const uint64_t da1 = make_smem_desc_sw128(sa1), db1 = make_smem_desc_sw128(sa1 + kABytes);
                        const uint32_t e0 = empty_base + (uint32_t)(s * 8), e1 = empty_base + (uint32_t)(s1 * 8);
                        int sn = s1 + 1; uint32_t phn = ph1;
                        if (sn == kStages) { sn = 0; phn ^= 1u; }
                        ready = mbar_test_wait(&full_bar[sn], phn);   // probe the stage after the pair now
                        const bool last_pair = ++pc == pairs_per_tap;
                        if (last_pair) pc = 0;
                        if (!last_pair) umma_stage_pair_elect<4>(acc, da0, db0, idesc, first ^ 1u, e0, da1, db1, e1);
                        else switch (p.last_ksteps) {
                            case 4: umma_stage_pair_elect<4>(acc, da0, db0, idesc, first ^ 1u, e0, da1, db1, e1); break;
                            case 3: umma_stage_pair_elect<3>(acc, da0, db0, idesc, first ^ 1u, e0, da1, db1, e1); break;
                            case 2: umma_stage_pair_elect<2>(acc, da0, db0, idesc, first ^ 1u, e0, da1, db1, e1); break;
                            default: umma_stage_pair_elect<1>(acc, da0, db0, idesc, first ^ 1u, e0, da1, db1, e1); break;
                        }
                        first = 0u;
                        s = sn; ph = phn;
                    }
                } else
                for (int tap = 0; tap < 9 * p.taps_d; ++tap) {
                    for (int ck = 0; ck < p.nchunks; ++ck, ++it) {
                        // `ready` was probed one stage ahead (the probe's latency hides behind the MMA issue)
                        if (!ready) {
                            if (p.dbg) { const long long w0_ = clock64(); mbar_wait(&full_bar[s], ph); t_wait += clock64() - w0_; }
                            else mbar_wait(&full_bar[s], ph);
                        }
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)(s * stage_bytes);
                        const uint64_t da = make_smem_desc_sw128(sa);
                        const uint64_t db = make_smem_desc_sw128(sa + kABytes);
                        const uint32_t ebar = empty_base + (uint32_t)(s * 8);
                        int sn = s + 1; uint32_t phn = ph;
                        if (sn == kStages) { sn = 0; phn ^= 1u; }
                        ready = mbar_test_wait(&full_bar[sn], phn);       // probe the NEXT stage now
                        // K-steps of the stage + the commit that frees it: one asm block, one elect.sync
                        if (ck != last_ck) umma_stage_elect<4>(acc, da, db, idesc, first ^ 1u, ebar);
                        else switch (p.last_ksteps) {
                            case 4: umma_stage_elect<4>(acc, da, db, idesc, first ^ 1u, ebar); break;
                            case 3: umma_stage_elect<3>(acc, da, db, idesc, first ^ 1u, ebar); break;
                            case 2: umma_stage_elect<2>(acc, da, db, idesc, first ^ 1u, ebar); break;
                            default: umma_stage_elect<1>(acc, da, db, idesc, first ^ 1u, ebar); break;
                        }
                        first = 0u;
                        s = sn; ph = phn;
                    }
                }
                umma_commit_elect(&tmem_full_bar[slot]);  // accumulator of this tile complete
            }
            if (p.dbg && lane == 0) {
                unsigned long long ns_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
                p.dbg[4 * blockIdx.x + 0] = clock64() - t_begin;     // cycles the issuer was alive
                p.dbg[4 * blockIdx.x + 1] = t_wait;                  // cycles blocked on full barriers
                p.dbg[4 * blockIdx.x + 2] = (long long)(ns_end - ns_begin);
                p.dbg[4 * blockIdx.x + 3] = it;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                           // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                      // accumulator row = voxel within the tile
        const int dw = r % p.bw, dh = (r / p.bw) % p.bh, dd = r / (p.bw * p.bh);
        int j = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
            int tile, n0, nlen;
            decode_item(p, item, tile, n0, nlen);
            int t = tile;
            const int w = (t % p.tw) * p.bw + dw; t /= p.tw;
            const int h = (t % p.th) * p.bh + dh; t /= p.th;
            const int d = (t % p.td) * p.bd + dd; t /= p.td;
            const int b = t;
            const bool valid = (w < p.W) && (h < p.H) && (d < p.D) && !(p.skip_h_edges && (h == 0 || h == p.H - 1));
            const size_t m = (((size_t)b * p.D + d) * p.H + h) * p.W + w;
            const int slot = j & 1;
            mbar_wait(&tmem_full_bar[slot], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * acc_stride);
            if (p.mode == 1) {
                float v[16];
                tmem_ld16(trow, v);
                float x = v[0] + p.bias[0];
                if (p.relu) x = fmaxf(x, 0.f);
                if (valid) p.out_f32[m] = x;
                if (p.pred) {
                    // a4 in the epilogue: the tile box spans all D disparities (td == 1), so the 128 accumulator rows hold
                    // the whole cost column of bw*bh pixels.  Same arithmetic as softargmin_kernel (max-shifted expf, sums
                    // in disparity order), so the fused and the two-launch routes agree bit for bit.
                    float *cs = cost_s[j & 1];
                    cs[r] = d < p.D ? x : -INFINITY;
                    asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps
                    const int npix = p.bw * p.bh;
                    if (r < npix && w < p.W && h < p.H) {                   // r < npix: dd == 0, (h, w) is the pixel
                        float mx = -INFINITY;
                        for (int e = 0; e < p.D; ++e) mx = fmaxf(mx, cs[e * npix + r]);
                        float s0 = 0.f, s1 = 0.f;
                        for (int e = 0; e < p.D; ++e) {
                            const float ex = expf(cs[e * npix + r] - mx);
                            s0 += ex; s1 += ex * (float)e;
                        }
                        p.pred[((size_t)b * p.H + h) * p.W + w] = s1 / s0;
                    }
                }
            } else if (p.mode == 2) {
                for (int cc = 0; cc < nlen; cc += 16) {
                    const int c0 = n0 + cc;
                    float v[16];
                    tmem_ld16(trow + (uint32_t)cc, v);
                    if (valid) {
                        const float4 *bp = reinterpret_cast<const float4 *>(p.bias + c0);
                        float4 *op = reinterpret_cast<float4 *>(p.out_f32_full + m * p.np + c0);
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 bv = __ldg(bp + i4);
                            float4 o = make_float4(v[4 * i4] + bv.x, v[4 * i4 + 1] + bv.y, v[4 * i4 + 2] + bv.z, v[4 * i4 + 3] + bv.w);
                            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                            if (p.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                            op[i4] = o;
                        }
                    }
                }
            } else {
                for (int cc = 0; cc < nlen; cc += 16) {
                    const int c0 = n0 + cc;
                    float v[16];
                    tmem_ld16(trow + (uint32_t)cc, v);    // warp-collective: every lane takes part
                    if (valid) {
                        __align__(16) __nv_bfloat16 o[16];
                        __align__(16) __nv_bfloat16 rs[16];
                        if (p.residual) {
                            const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + m * p.np + c0);
                            *reinterpret_cast<uint4 *>(rs) = __ldg(rp);
                            *reinterpret_cast<uint4 *>(rs + 8) = __ldg(rp + 1);
                        }
                        const float4 *bp = reinterpret_cast<const float4 *>(p.bias + c0);
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 bv = __ldg(bp + i4);
                            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float x = v[4 * i4 + i] + bb[i];
                                if (p.relu) x = fmaxf(x, 0.f);
                                if (p.residual) x += __bfloat162float(rs[4 * i4 + i]);
                                o[4 * i4 + i] = __float2bfloat16(x);
                            }
                        }
                        uint4 *op = reinterpret_cast<uint4 *>(p.out_bf16 + m * p.np + c0);
                        op[0] = *reinterpret_cast<const uint4 *>(o);
                        op[1] = *reinterpret_cast<const uint4 *>(o + 8);
                    }
                }
            }
            // this warp has read its lanes of the slot: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// =============================================================================================
// 2-CTA variant (cta_group::2): a CTA PAIR (cluster of 2 SMs) computes two 128-voxel tiles with ONE
// M=256 MMA per K-step.  Each CTA stages its own A tile (128 voxels x 64 ch) and HALF of the weight
// tile (np/2 output channels x 64 ch); the tensor cores of both SMs read the two B halves, so per
// CTA the shared-memory fill and operand traffic per stage drops from 16+28 KB to 16+14 KB and the
// single issuing thread launches half as many instructions per FLOP.
//   * TMA loads are issued by both CTAs with .cta_group::2 and complete_tx on the LEADER's full barrier
//     (count 2: leader arrive.expect_tx for both CTAs' bytes + the peer's remote arrive);
//   * the leader's tcgen05.commit is multicast to both CTAs' empty / tmem_full barriers;
//   * epilogue warps of both CTAs arrive on the leader's tmem_empty barrier (count 8).
// =============================================================================================
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4,
                                                uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
          "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void *dst, const CUtensorMap *tm, int c0, int c1, int c2,
                                                uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar) : "memory");
}
#define DECNET_MMA2_NEXT(OFF)                                                      \
        "add.u64 a, %1, " #OFF ";\n\t"                                              \
        "add.u64 b, %2, " #OFF ";\n\t"                                              \
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, t;\n\t"
#define DECNET_STAGE2_HEAD                                                         \
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b;\n\t.reg .b16 m;\n\t"            \
        "elect.sync _|e, 0xffffffff;\n\t"                                           \
        "setp.ne.b32 p, %4, 0;\n\t"                                                 \
        "setp.eq.b32 t, 0, 0;\n\t"                                                  \
        "mov.b16 m, 3;\n\t"                                                         \
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
#define DECNET_STAGE2_TAIL                                                         \
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%5], m;\n\t}"
template <int KSTEPS>
__device__ __forceinline__ void umma2_stage_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate_first, uint32_t bar_addr) {
    if constexpr (KSTEPS == 4) {
        asm volatile(DECNET_STAGE2_HEAD DECNET_MMA2_NEXT(2) DECNET_MMA2_NEXT(4) DECNET_MMA2_NEXT(6) DECNET_STAGE2_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    } else if constexpr (KSTEPS == 3) {
        asm volatile(DECNET_STAGE2_HEAD DECNET_MMA2_NEXT(2) DECNET_MMA2_NEXT(4) DECNET_STAGE2_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    } else if constexpr (KSTEPS == 2) {
        asm volatile(DECNET_STAGE2_HEAD DECNET_MMA2_NEXT(2) DECNET_STAGE2_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    } else {
        asm volatile(DECNET_STAGE2_HEAD DECNET_STAGE2_TAIL
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(bar_addr) : "memory");
    }
}
__device__ __forceinline__ void umma2_commit_mc_elect(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b16 m, 3;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(bar_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3d_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];       // this CTA's operands landed (local TMA completion)
    __shared__ __align__(8) uint64_t peer_full_bar[kMaxStages];  // leader only: the peer's operands landed (relayed)
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int half_n = p.np >> 1;
    const int b_bytes = half_n * kRowBytes;               // this CTA's half of the weight tile
    const int stage_bytes = kABytes + b_bytes;
    const int kStages = p.stages;
    const int acc_stride = p.tmem_cols >> 1;
    const int num_units = (p.num_tiles + 1) >> 1;         // a unit = two consecutive tiles (one per CTA)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&peer_full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // barriers of both CTAs initialised before any remote use
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                int t = min(2 * u + (int)cta_rank, p.num_tiles - 1);
                const int w0 = (t % p.tw) * p.bw; t /= p.tw;
                const int h0 = (t % p.th) * p.bh; t /= p.th;
                const int d0 = (t % p.td) * p.bd; t /= p.td;
                const int b = t;
                int tap = 0;
                for (int kd = (p.taps_d == 3 ? 0 : 1); kd < (p.taps_d == 3 ? 3 : 2); ++kd)
                    for (int kh = 0; kh < 3; ++kh)
                        for (int kw = 0; kw < 3; ++kw, ++tap)
                            for (int ck = 0; ck < p.nchunks; ++ck) {
                                mbar_wait(&empty_bar[s], ph ^ 1u);
                                unsigned char *sa = base + (size_t)s * stage_bytes;
                                // every CTA completes its OWN barrier (signalling the leader's barrier from the peer's TMA
                                // made the loads 2-3x slower); the peer's idle MMA warp relays "landed" to the leader
                                const bool la = !(p.skip_tma & 2), lb = !(p.skip_tma & 4);
                                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)((la ? kABytes : 0) + (lb ? b_bytes : 0)));
                                if (la) tma_load_5d(sa, &tmA, ck * p.chunk_ch, w0 + kw - 1, h0 + kh - 1, d0 + kd - 1, b, &full_bar[s]);
                                if (lb) tma_load_3d(sa + kABytes, &tmB, ck * p.chunk_ch, (int)cta_rank * half_n, tap, &full_bar[s]);
                                if (++s == kStages) { s = 0; ph ^= 1u; }
                            }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only; whole warp converged) =====================
        if (leader) {
            const uint32_t idesc = make_idesc(p.fmt, 2 * kTileM, p.np);
            int j = 0;
            int s = 0; uint32_t ph = 0;
            const uint32_t smem_base = smem_u32(base);
            const uint32_t empty_base = smem_u32(&empty_bar[0]);
            bool ready = false;
            long long t_begin = clock64(), t_wait = 0, t_wait_acc = 0; int it = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters, ++j) {
                const int slot = j & 1;
                { const long long w0_ = clock64(); mbar_wait(&tmem_empty_bar[slot], ((uint32_t)(j >> 1) & 1u) ^ 1u); t_wait_acc += clock64() - w0_; }
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(slot * acc_stride);
                uint32_t first = 1u;
                const int last_ck = p.nchunks - 1;
                for (int tap = 0; tap < 9 * p.taps_d; ++tap) {
                    for (int ck = 0; ck < p.nchunks; ++ck, ++it) {
                        if (!ready) { const long long w0_ = clock64(); mbar_wait(&full_bar[s], ph); mbar_wait(&peer_full_bar[s], ph); t_wait += clock64() - w0_; }
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)(s * stage_bytes);
                        const uint64_t da = make_smem_desc_sw128(sa);
                        const uint64_t db = make_smem_desc_sw128(sa + kABytes);
                        const uint32_t ebar = empty_base + (uint32_t)(s * 8);
                        int sn = s + 1; uint32_t phn = ph;
                        if (sn == kStages) { sn = 0; phn ^= 1u; }
                        ready = mbar_test_wait(&full_bar[sn], phn) && mbar_test_wait(&peer_full_bar[sn], phn);
                        if (ck != last_ck) umma2_stage_elect<4>(acc, da, db, idesc, first ^ 1u, ebar);
                        else switch (p.last_ksteps) {
                            case 4: umma2_stage_elect<4>(acc, da, db, idesc, first ^ 1u, ebar); break;
                            case 3: umma2_stage_elect<3>(acc, da, db, idesc, first ^ 1u, ebar); break;
                            case 2: umma2_stage_elect<2>(acc, da, db, idesc, first ^ 1u, ebar); break;
                            default: umma2_stage_elect<1>(acc, da, db, idesc, first ^ 1u, ebar); break;
                        }
                        first = 0u;
                        s = sn; ph = phn;
                    }
                }
                umma2_commit_mc_elect(smem_u32(&tmem_full_bar[slot]));     // both CTAs' epilogues
            }
            if (p.dbg && lane == 0) {
                p.dbg[4 * blockIdx.x + 0] = clock64() - t_begin;
                p.dbg[4 * blockIdx.x + 1] = t_wait;
                p.dbg[4 * blockIdx.x + 2] = t_wait_acc;
                p.dbg[4 * blockIdx.x + 3] = it;
            }
        } else if (lane == 0) {
            // ===================== peer CTA: relay "my operands landed" to the leader, stage by stage =====================
            int s = 0; uint32_t ph = 0;
            const int per_unit = 9 * p.taps_d * p.nchunks;
            for (int u = cluster_id; u < num_units; u += num_clusters)
                for (int i = 0; i < per_unit; ++i) {
                    mbar_wait(&full_bar[s], ph);
                    remote_arrive(map_to_cta(smem_u32(&peer_full_bar[s]), 0));
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
        }
    } else {
        // ===================== epilogue (warps 2..5 of both CTAs) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int dw = r % p.bw, dh = (r / p.bw) % p.bh, dd = r / (p.bw * p.bh);
        int j = 0;
        for (int u = cluster_id; u < num_units; u += num_clusters, ++j) {
            const int tile = 2 * u + (int)cta_rank;
            const bool tile_ok = tile < p.num_tiles;
            int t = min(tile, p.num_tiles - 1);
            const int w = (t % p.tw) * p.bw + dw; t /= p.tw;
            const int h = (t % p.th) * p.bh + dh; t /= p.th;
            const int d = (t % p.td) * p.bd + dd; t /= p.td;
            const int b = t;
            const bool valid = tile_ok && (w < p.W) && (h < p.H) && (d < p.D);
            const size_t m = (((size_t)b * p.D + d) * p.H + h) * p.W + w;
            const int slot = j & 1;
            mbar_wait(&tmem_full_bar[slot], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * acc_stride);
            if (p.mode == 1) {
                float v[16];
                tmem_ld16(trow, v);
                if (valid) { float x = v[0] + p.bias[0]; if (p.relu) x = fmaxf(x, 0.f); p.out_f32[m] = x; }
            } else {
                for (int c0 = 0; c0 < p.np; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + (uint32_t)c0, v);
                    if (valid) {
                        __align__(16) __nv_bfloat16 o[16];
                        __align__(16) __nv_bfloat16 rs[16];
                        if (p.residual) {
                            const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + m * p.np + c0);
                            *reinterpret_cast<uint4 *>(rs) = __ldg(rp);
                            *reinterpret_cast<uint4 *>(rs + 8) = __ldg(rp + 1);
                        }
                        const float4 *bp = reinterpret_cast<const float4 *>(p.bias + c0);
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 bv = __ldg(bp + i4);
                            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float x = v[4 * i4 + i] + bb[i];
                                if (p.relu) x = fmaxf(x, 0.f);
                                if (p.residual) x += __bfloat162float(rs[4 * i4 + i]);
                                o[4 * i4 + i] = __float2bfloat16(x);
                            }
                        }
                        uint4 *op = reinterpret_cast<uint4 *>(p.out_bf16 + m * p.np + c0);
                        op[0] = *reinterpret_cast<const uint4 *>(o);
                        op[1] = *reinterpret_cast<const uint4 *>(o + 8);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tmem_empty_bar[slot]);
                else remote_arrive(map_to_cta(smem_u32(&tmem_empty_bar[slot]), 0));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // the peer may still be signalling my barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
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

static thread_local long long *g_conv3d_dbg = nullptr;
static thread_local int g_conv3d_variant = 0;      // 0 = auto (1-CTA kernel), 1 = force 1-CTA, 2 = CTA-pair kernel
void decnet_conv3d_set_variant(int v) { g_conv3d_variant = v; }
void decnet_conv3d_debug_timing(void *dbg_buffer) { g_conv3d_dbg = static_cast<long long *>(dbg_buffer); }

// Common launcher.  esize 2 -> bf16 operands (kind::f16, K-step 16, 64 channels per stage row),
//                  esize 4 -> fp32 operands read as tf32 (kind::tf32, K-step 8, 32 channels per row).
static int launch_conv(const void *x, const void *w_packed, const float *bias, const void *residual, void *out,
                       int out_mode, int esize, int taps_d, int B, int D, int H, int W, int cp, int np, int relu,
                       void *stream, int round_out = 0, int skip_h_edges = 0, float *pred = nullptr, int *pred_fused = nullptr)
{
    DECNET_REQUIRE(x && w_packed && bias && out, "null pointer");
    DECNET_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "non-positive size");
    const int kstep = esize == 2 ? 16 : 8;
    const int chunk_ch = kRowBytes / esize;
    DECNET_REQUIRE(cp % kstep == 0 && cp >= kstep && cp <= 4096, "cp=%d must be a multiple of %d", cp, kstep);
    DECNET_REQUIRE(np % 16 == 0 && np >= 16 && np <= 256, "np=%d must be a multiple of 16 in [16,256]", np);
    DECNET_REQUIRE(out_mode >= 0 && out_mode <= 2, "out_mode must be 0 (bf16 [M][np]), 1 (fp32 [M]) or 2 (fp32 [M][np])");
    DECNET_REQUIRE(taps_d == 3 || (taps_d == 1 && D == 1), "taps_d=1 needs a D=1 volume");
    DECNET_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(residual) & 15u) == 0, "pointers must be 16-byte aligned");
    Params p{};
    p.bias = bias;
    p.residual = static_cast<const __nv_bfloat16 *>(residual);
    p.out_bf16 = out_mode == 0 ? static_cast<__nv_bfloat16 *>(out) : nullptr;
    p.out_f32 = out_mode == 1 ? static_cast<float *>(out) : nullptr;
    p.out_f32_full = out_mode == 2 ? static_cast<float *>(out) : nullptr;
    p.B = B; p.D = D; p.H = H; p.W = W; p.cp = cp; p.np = np;
    p.fmt = esize == 2 ? 1 : 2; p.chunk_ch = chunk_ch; p.taps_d = taps_d;
    p.nchunks = (cp + chunk_ch - 1) / chunk_ch;
    p.last_ksteps = (cp - (p.nchunks - 1) * chunk_ch) / kstep;
    pick_tile(W, H, D, p.bw, p.bh, p.bd);
    p.tw = (W + p.bw - 1) / p.bw; p.th = (H + p.bh - 1) / p.bh; p.td = (D + p.bd - 1) / p.bd;
    p.relu = relu; p.mode = out_mode; p.round_tf32 = round_out; p.skip_h_edges = skip_h_edges;
    p.tmem_cols = np <= 16 ? 32 : np <= 32 ? 64 : np <= 64 ? 128 : np <= 128 ? 256 : 512;   // 2 slots
    // fused soft-argmin: needs the whole disparity axis inside one tile box (single-CTA kernel, mode 1)
    p.pred = (pred && out_mode == 1 && p.td == 1 && (g_conv3d_variant % 10) != 2) ? pred : nullptr;
    if (pred_fused) *pred_fused = p.pred != nullptr;

    const long long tiles = (long long)B * p.tw * p.th * p.td;
    DECNET_REQUIRE(tiles < (1ll << 31), "too many tiles");
    p.num_tiles = (int)tiles;
    p.dbg = g_conv3d_dbg;
    p.skip_tma = g_conv3d_variant == 3 ? 1 : (g_conv3d_variant >= 10 ? (g_conv3d_variant / 10) * 2 : 0);
    const int sms = sm_count_cached();
    // The CTA-pair kernel is correct (same tests) but measured 2x slower than the single-CTA one in
    // round 1 (MMAs slow down 3x while TMA fills run): opt-in only.
    const bool two_cta = (g_conv3d_variant % 10) == 2 && tiles >= 2 && sms >= 2 && out_mode != 2 && !skip_h_edges;
    const size_t stage_bytes = kABytes + (size_t)(two_cta ? np / 2 : np) * kRowBytes;
    p.stages = (int)((226 * 1024 - 1024) / stage_bytes);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    DECNET_REQUIRE(p.stages >= 2, "stage too large");
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    const CUtensorMapDataType dt = esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    // split the last, partial round of tiles in two along N when that round then still fits the machine (single-CTA kernel)
    p.num_items = p.num_tiles; p.tail_first = p.num_tiles; p.tail_tiles = 0;
    if (!two_cta && g_conv3d_variant % 10 != 1 && np % 32 == 0 && out_mode != 1) {
        const int tail = p.num_tiles % sms;
        if (tail > 0 && 2 * tail <= sms) {
            p.tail_tiles = tail; p.tail_first = p.num_tiles - tail; p.num_items = p.num_tiles + tail;
        }
    }
    CUtensorMap tmA, tmB, tmBh;
    {
        const uint64_t dims[5] = {(uint64_t)cp, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)B};
        const uint64_t strides[4] = {(uint64_t)cp * esize, (uint64_t)W * cp * esize, (uint64_t)H * W * cp * esize,
                                     (uint64_t)D * H * W * cp * esize};
        const uint32_t box[5] = {(uint32_t)chunk_ch, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, 1u};
        int rc = encode_tensor_map(&tmA, dt, 5, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)cp, (uint64_t)np, (uint64_t)(9 * taps_d)};
        const uint64_t strides[2] = {(uint64_t)cp * esize, (uint64_t)np * cp * esize};
        const uint32_t box[3] = {(uint32_t)chunk_ch, (uint32_t)(two_cta ? np / 2 : np), 1u};
        int rc = encode_tensor_map(&tmB, dt, 3, w_packed, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
        const uint32_t boxh[3] = {(uint32_t)chunk_ch, (uint32_t)(p.tail_tiles ? np / 2 : np), 1u};
        rc = encode_tensor_map(&tmBh, dt, 3, w_packed, dims, strides, boxh, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    if (two_cta) {
        static std::mutex mu2;
        static size_t set_for2[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        {
            std::lock_guard<std::mutex> lk(mu2);
            if (dev < 0 || dev >= 64 || set_for2[dev] < smem) {
                DECNET_CUDA(cudaFuncSetAttribute(conv3d_tcgen05_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                if (dev >= 0 && dev < 64) set_for2[dev] = smem;
            }
        }
        const long long units = (tiles + 1) / 2;
        const long long clusters = units < sms / 2 ? units : sms / 2;
        conv3d_tcgen05_2cta_kernel<<<(unsigned)(2 * clusters), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
        return after_launch("conv3d_tcgen05_2cta_kernel");
    }
    {   // raise the dynamic smem limit once per device and size (the call costs host time on every launch)
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < smem) {
            DECNET_CUDA(cudaFuncSetAttribute(conv3d_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) set_for[dev] = smem;
        }
    }
    const unsigned grid = (unsigned)(p.num_items < sms ? p.num_items : sms);       // persistent: one CTA per SM
    DECNET_CUDA(launch_pdl(conv3d_tcgen05_kernel, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), tmA, tmB, tmBh, p));
    return after_launch("conv3d_tcgen05_kernel");
}

int decnet_conv3d_bf16(const void *x_ndhwc, const void *w_packed, const float *bias, const void *residual,
                       void *out, int out_mode, int B, int D, int H, int W, int cp, int np, int relu, void *stream)
{
    DECNET_REQUIRE(out_mode == 0 || out_mode == 1, "out_mode must be 0 (bf16 [M][np]) or 1 (fp32 [M], channel 0)");
    return launch_conv(x_ndhwc, w_packed, bias, residual, out, out_mode, 2, 3, B, D, H, W, cp, np, relu, stream);
}

// a3 (last layer) + a4 in one launch: cost fp32 [B,D,H,W] (channel 0) and pred [B,H,W] = soft-argmin_d(cost).  The
// soft-argmin runs in the conv epilogue when one tile box spans the D axis (D <= 8 at every published stage-0 size);
// otherwise the standalone kernel follows on the same stream.  Same result either way.
int decnet_conv3d_bf16_softargmin(const void *x_ndhwc, const void *w_packed, const float *bias, float *cost, float *pred,
                                  int B, int D, int H, int W, int cp, int np, int relu, void *stream)
{
    DECNET_REQUIRE(pred, "null pointer");
    int fused = 0;
    const int st = launch_conv(x_ndhwc, w_packed, bias, nullptr, cost, 1, 2, 3, B, D, H, W, cp, np, relu, stream, 0, 0, pred, &fused);
    if (st != 0 || fused) return st;
    return decnet_softargmin(cost, pred, B, D, H, W, stream);
}

int decnet_conv3d_bf16_band(const void *x_ndhwc, const void *w_packed, const float *bias, const void *residual,
                            void *out, int B, int D, int H, int W, int cp, int np, int relu, void *stream)
{
    return launch_conv(x_ndhwc, w_packed, bias, residual, out, 0, 2, 3, B, D, H, W, cp, np, relu, stream, 0, 1);
}

int decnet_conv2d_tf32_nhwc(const float *x_nhwc, const float *w_packed, const float *bias, float *out,
                            int B, int H, int W, int cp, int np, int relu, int round_out_tf32, void *stream)
{
    return launch_conv(x_nhwc, w_packed, bias, nullptr, out, 2, 4, 1, B, 1, H, W, cp, np, relu, stream, round_out_tf32);
}

}  // extern "C"
