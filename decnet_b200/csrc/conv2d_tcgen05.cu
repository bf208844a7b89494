// decnet_b200/csrc/conv2d_tcgen05.cu -- the small 3x3 Conv2d layers of the fine stages (detail detection,
// soft attention, refinement: modules/submodule.py:347-372, 593-604, 666-762) as TF32 implicit GEMMs on the
// 5th-generation tensor cores, reading and writing the reference's own NCHW fp32 tensors.
//
// Why not the channels-last kernel of conv3d_tcgen05.cu: these layers have 1..24 channels, their
// neighbours (pack / warp / blend kernels, SpaMat) are NCHW, and at 8 channels a channels-last row is
// 32 bytes.  In NCHW the pixels of one image row are contiguous, so pixels become the GEMM M
// dimension of an **MN-major** A operand:
//
//   D[m = pixel, n = (kw, cout)] += A[m, k = cin] * Wt[k, n]          per row tap kh
//
//   * A tile: ONE 4-D TMA box {32 px, 8 ch, RH rows, 1} of the view (W, C, H, B) of x, at signed
//     coordinates (zero padding = TMA out-of-bounds fill).  TMA writes [row][ch][32 px] = one 1 KB
//     block per image row, which is exactly two UMMA "MN-major, 128-byte swizzle with 32-byte
//     atoms" k-atoms (4 ch x 128 B each; SBO = 512 B), image rows being the m-atoms (LBO = 1 KB).
//     fp32 MN-major operands exist only in that layout (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B
//     = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side; checked in scripts/micro/umma_tf32_mn.cu).
//   * One MMA = M 128 (4 image rows x 32 px), K 8 (one channel chunk), N = 3*CP rounded up to 32.
//     The row tap kh is a row offset into the same smem tile (dilation d: kh*d rows), so a tile of
//     R = 4G output rows needs its R+2d input rows loaded ONCE.
//   * The column tap kw cannot be an address offset (a 4-byte shift is below the descriptor's 16-byte
//     granularity), so the three kw taps are three groups of N columns: D_kw[px] = x[px] * w[., kw]
//     at the UNSHIFTED pixel, and the epilogue -- one warp per image row, lane = pixel -- adds
//     D_0[lane-d] + D_1[lane] + D_2[lane+d] with two warp shuffles.  Lanes d..31-d are valid.  TMA needs
//     the box's first column 16-byte aligned (measured: scripts/micro/tma4d_probe.cu), so tiles advance by
//     S = (32-2d) rounded down to 4 columns (28 of 32 lanes useful at d = 1) from column -roundup4(d).
//   * kind::tf32 truncates fp32 operands; four converter warps round every landed tile to TF32
//     (cvt.rna, what cuDNN's TF32 path does) in place before the MMA warp reads it.  Weights are
//     rounded on the host.
//
//   * split = 1 ("3xTF32"): hi = trunc_tf32(x) is what the MMA reads from the raw tile; the converters write
//     lo = rna_tf32(x - hi) into the second half of the stage, the weights arrive as hi and lo parts, and every (row group, row tap) issues three MMAs into
//     the same accumulator: lo*hi + hi*lo + hi*hi.  The dropped lo*lo term and the rounding of the lo parts are
//     ~2^-22 relative, so the result is fp32-class (the parity gates of tests/test_glue_gpu.py run on this mode);
//     the kernel is bound by its epilogue, so the extra MMAs on K = 8 are nearly free.
//   * split = 2 (the default of the Python side): the two correction products as ONE kind::f16 MMA per tap.  The converters
//     write the fp16 operand [fp16(2^11 * lo(x)) | fp16(x)] (K = 16: 8 + 8 channels) into the second half of the stage in the
//     MN-major SWIZZLE_64B layout, the second half of the weights holds [fp16(hi(w) * 2^sw) ; fp16(lo(w) * 2^(11+sw))], the TF32
//     weights are hi(w) * 2^(11+sw), and the epilogue multiplies the accumulator by 2^-(11+sw) (bias[CP]): two MMAs per tap, a
//     third fewer operand bytes through shared memory, the same error budget.
//
// Warp roles (704 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2-5 = converters,
// 6-21 = epilogue.  Persistent CTAs, smem ring of RH-KB stages, two TMEM accumulator slots, the
// weights of the whole layer resident in smem.
#include "common.cuh"
#include "tma_utils.cuh"
#include <mutex>

namespace decnet {
namespace conv2dtc {

constexpr int kMaxStages = 8;
constexpr int kConvWarps = 4;
constexpr int kEpiWarps = 16;              // 4 per TMEM lane quarter
constexpr int kThreads = 32 * (2 + kConvWarps + kEpiWarps);
constexpr int kRowBlock = 1024;            // one image row of a stage: 8 channels x 32 pixels x 4 bytes
constexpr int kWBox = 64;                  // weight rows (128 B each) per TMA box

struct Params {
    const float *bias;                     // [CP]
    const float *addend;                   // [B][H][W] added to output channel 0 after the activation (Cout == 1: the
                                           // refinement's disp + residual, submodule.py:761), or null
    float *out;                            // [B][Cout][H][W]
    int B, Cout, H, W;
    int w_valid;                           // columns >= w_valid (<= W) are written as zeros: row-pitch padding of a W % 4 != 0 image
    int dil;                               // dilation = padding
    int nck;                               // channel chunks of 8
    int ck1, ck2;                          // chunks [0,ck1) come from source 0, [ck1,ck2) from source 1, [ck2,nck) from source 2
    int CP, N, natoms;                     // Cout rounded up to 4; MMA N = 3*CP rounded up to 32; N/32
    int G, RH, vw, pad;                    // row groups per tile (R = 4G), input rows per stage, columns per tile, first box column = -pad
    int tw, th, num_tiles, stages;
    int relu;
    int w_rows, w_bytes;                   // packed weight rows (of 128 B) and the smem reserved for them
    int split;                             // 1: error-compensated 3xTF32 (hi/lo operand split, fp32-class result)
    int lo_off, w_lo_off;                  // split: byte offset of the lo tile inside a stage / of the lo weights
    int tmem_cols;
    long long *prof;                       // tuning only: per-role cycle counters of CTA 0 (16 int64), or null
    int dbg;                               // tuning only: 1 skip rounding, 2 skip epilogue math/stores, 4 skip MMAs
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// MN-major, SWIZZLE_128B_BASE32B: start>>4 | LBO (mn-atom stride 1024 B)>>4 at 16 | SBO (k-atom stride 512 B)>>4
// at 32 | version 1 at 46 | layout type 1 at 61
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (64ull << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc_tf32_mn(int N) {
    // c_format F32, a/b format TF32 (2), a_major = b_major = MN (bits 15, 16), n_dim N>>3 at 17, m_dim 128>>4 at 24
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}
// whole warp converged; one elected lane issues (see conv3d_tcgen05.cu on why the election lives inside the asm)
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// The three row taps kh of one row group behind ONE election: descriptors advance by a_step (kh*dil image rows) and
// b_step (one row tap of the weights), both in 16-byte units of the descriptor's address field.
__device__ __forceinline__ void umma_tf32_kh3(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t a_step, uint32_t b_step,
                                              uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b, sa, sb;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "cvt.u64.u32 sa, %3;\n\t"
        "cvt.u64.u32 sb, %4;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %5, p;\n\t"
        "add.u64 a, %1, sa;\n\t"
        "add.u64 b, %2, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %5, t;\n\t"
        "add.u64 a, a, sa;\n\t"
        "add.u64 b, b, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %5, t;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(a_step), "r"(b_step), "r"(idesc), "r"(accumulate_first) : "memory");
}
// 3xTF32: per row tap lo(x)*hi(w) + hi(x)*lo(w) + hi(x)*hi(w), small terms first; a_lo / b_lo are the descriptor offsets
// of the lo tile / lo weights (16-byte units).
__device__ __forceinline__ void umma_tf32_kh3_split(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t a_step, uint32_t b_step,
                                                    uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b, al, bl, sa, sb, la, lb;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %8, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "cvt.u64.u32 sa, %3;\n\t"
        "cvt.u64.u32 sb, %4;\n\t"
        "cvt.u64.u32 la, %5;\n\t"
        "cvt.u64.u32 lb, %6;\n\t"
        "add.u64 al, %1, la;\n\t"
        "add.u64 bl, %2, lb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], al, %2, %7, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, bl, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %7, t;\n\t"
        "add.u64 a, %1, sa;\n\t"
        "add.u64 b, %2, sb;\n\t"
        "add.u64 al, al, sa;\n\t"
        "add.u64 bl, bl, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], al, b, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, bl, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %7, t;\n\t"
        "add.u64 a, a, sa;\n\t"
        "add.u64 b, b, sb;\n\t"
        "add.u64 al, al, sa;\n\t"
        "add.u64 bl, bl, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], al, b, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, bl, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %7, t;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(a_step), "r"(b_step), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first)
        : "memory");
}
// split = 2: per row tap ONE kind::f16 MMA [lo16(x) | fp16(x)] * [hi16(w) ; lo16(w)] (both correction terms, K = 16) and the
// kind::tf32 hi(x)*hi(w).  The fp16 operands sit where the TF32 lo parts would (same 1 KB blocks, same LBO / SBO), in the
// MN-major SWIZZLE_64B layout (scripts/micro/umma_f16_mn.cu): their descriptors are the lo descriptors with the layout type
// 1 -> 4 (+3 at bit 61).  idesc16: the kind::f16 instruction descriptor.
__device__ __forceinline__ void umma_kh3_split16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t a_step, uint32_t b_step,
                                                 uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t idesc16,
                                                 uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b, al, bl, sa, sb, la, lb;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "cvt.u64.u32 sa, %3;\n\t"
        "cvt.u64.u32 sb, %4;\n\t"
        "cvt.u64.u32 la, %5;\n\t"
        "cvt.u64.u32 lb, %6;\n\t"
        "add.u64 al, %1, la;\n\t"
        "add.u64 bl, %2, lb;\n\t"
        "add.u64 al, al, 0x6000000000000000;\n\t"
        "add.u64 bl, bl, 0x6000000000000000;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bl, %8, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %7, t;\n\t"
        "add.u64 a, %1, sa;\n\t"
        "add.u64 b, %2, sb;\n\t"
        "add.u64 al, al, sa;\n\t"
        "add.u64 bl, bl, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bl, %8, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %7, t;\n\t"
        "add.u64 a, a, sa;\n\t"
        "add.u64 b, b, sb;\n\t"
        "add.u64 al, al, sa;\n\t"
        "add.u64 bl, bl, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bl, %8, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %7, t;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(a_step), "r"(b_step), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(idesc16),
          "r"(accumulate_first)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t rna_tf32(uint32_t x) {
    // == cvt.rna.tf32.f32 for finite values (round to nearest, ties away), as two full-rate integer ops
    return (x + 0x1000u) & 0xFFFFE000u;
}

template <int NC> struct TmemLd;
template <> struct TmemLd<4> {
    static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[4]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    }
};
template <> struct TmemLd<8> {
    static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[8]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    }
};

// Epilogue of one warp.  kEpiWarps/4 warps share each TMEM lane quarter (row groups g = part, part + 4, ...).
// A work item is (row group g, NC output channels from c0): TMEM -> registers, the column taps by warp
// shuffle (left neighbour's kw=0 product, own kw=1, right neighbour's kw=2), bias, ReLU, store.  This
// warp's own instruction latency is the budget, so no divisions and no recomputed addresses in the loop.
template <int NC>
__device__ __forceinline__ void epilogue(const Params &p, const float *bias_s, uint64_t *tmem_full_bar,
                                         uint64_t *tmem_empty_bar, uint32_t tmem_base, int acc_stride,
                                         int warp, int lane, long long (&prof_acc)[4])
{
    const int q = warp & 3;                                   // TMEM lane quarter = row within the group
    const int half = (warp - (2 + kConvWarps)) >> 2;          // row groups half, half + kEpiWarps/4, ...
    const int d = p.dil;
    const bool lane_ok = lane >= d && lane < d + p.vw;        // d + vw <= 32 - d
    const size_t plane = (size_t)p.H * p.W;
    const int G = (p.dbg & 2) ? 0 : p.G, CP = p.CP, N = p.N;
    const size_t gstride = (size_t)4 * p.W;
    const float floor_ = p.relu ? 0.f : -INFINITY;            // ReLU as an unconditional max
    // split = 2: every product carries the weights' power-of-two scale 2^(11+sw) (ops.pack_conv2d_tf32_nchw_weights); its inverse
    // sits behind the bias
    const bool scaled = p.split == 2;
    const float inv_s = scaled ? bias_s[p.CP] : 1.f;
    // tile cursor without divisions: (tx, ty, b) advance by gridDim.x tiles
    int tx, ty, tb;
    { int t = blockIdx.x; tx = t % p.tw; t /= p.tw; ty = t % p.th; tb = t / p.th; }
    const int dx = gridDim.x % p.tw, dy = (gridDim.x / p.tw) % p.th, db = gridDim.x / (p.tw * p.th);
    int j = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
        const int col = tx * p.vw - p.pad + lane;
        const int h0 = ty * (4 * p.G);
        const bool col_ok = lane_ok && col >= 0 && col < p.W;
        const bool keep = col < p.w_valid;                       // pitch-padding columns stay exactly zero
        const int slot = j & 1;
        const long long tq0 = clock64();
        mbar_wait_relaxed(&tmem_full_bar[slot], (uint32_t)(j >> 1) & 1u);
        prof_acc[0] += clock64() - tq0;
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * acc_stride);
        float *orow0 = p.out + (size_t)tb * p.Cout * plane + (size_t)(h0 + q) * p.W + col;
        const int rows_left = p.H - (h0 + q);                 // row 4g+q is inside the image iff 4g < rows_left
        uint32_t A[3][NC];
        auto issue = [&](uint32_t (&v)[3][NC], int g, int c0) {
            const uint32_t ta = trow + (uint32_t)(g * N + c0);
            TmemLd<NC>::ld(ta, v[0]);
            TmemLd<NC>::ld(ta + (uint32_t)CP, v[1]);
            TmemLd<NC>::ld(ta + (uint32_t)(2 * CP), v[2]);
        };
        auto finish = [&](const uint32_t (&v)[3][NC], int g, int c0) {
            float x[NC];
#pragma unroll
            for (int i4 = 0; i4 < NC; i4 += 4) {
                const float4 bv = *reinterpret_cast<const float4 *>(bias_s + c0 + i4);
                const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v[0][i4 + i]), d);
                    const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v[2][i4 + i]), d);
                    const float mid = __uint_as_float(v[1][i4 + i]);
                    const float y = scaled ? fmaf((left + mid) + right, inv_s, bb[i]) : (left + mid) + (right + bb[i]);
                    x[i4 + i] = keep ? fmaxf(y, floor_) : 0.f;
                }
            }
            if (col_ok && 4 * g < rows_left) {
                float *op = orow0 + (size_t)g * gstride + (size_t)c0 * plane;
                if (p.addend && c0 == 0 && keep)
                    x[0] += __ldg(p.addend + (size_t)tb * plane + (size_t)(h0 + q) * p.W + col + (size_t)g * gstride);
                if (c0 + NC <= p.Cout) {                      // full chunk: straight-line stores
#pragma unroll
                    for (int i = 0; i < NC; ++i) { *op = x[i]; op += plane; }
                } else {
#pragma unroll
                    for (int i = 0; i < NC; ++i) { if (c0 + i < p.Cout) *op = x[i]; op += plane; }
                }
            }
        };
        // four warps per lane quarter: the other warps' items hide this warp's TMEM-load latency
        for (int g = half; g < G; g += kEpiWarps / 4)
            for (int c0 = 0; c0 < CP; c0 += NC) {
                issue(A, g, c0);
                tmem_ld_wait();
                finish(A, g, c0);
            }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[slot]);
        tx += dx; if (tx >= p.tw) { tx -= p.tw; ++ty; }
        ty += dy; if (ty >= p.th) { ty -= p.th; ++tb; }
        tb += db;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
conv2d_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX1,
                      const __grid_constant__ CUtensorMap tmX2, const __grid_constant__ CUtensorMap tmW, const Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];       // TMA landed
    __shared__ __align__(8) uint64_t ready_bar[kMaxStages];      // rounded to TF32
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];      // MMAs that read the stage retired
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[100];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x < p.CP + (p.split == 2 ? 1 : 0)) bias_s[threadIdx.x] = p.bias[threadIdx.x];
    unsigned char *wsm = base;                                   // resident weights
    unsigned char *ring = base + p.w_bytes;                      // stages
    const int tile_bytes = p.RH * kRowBlock;                     // what TMA delivers per stage
    const int stage_bytes = p.split ? 2 * tile_bytes : tile_bytes;
    const int kStages = p.stages;
    const int acc_stride = p.G * p.N;                            // TMEM columns per accumulator slot

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmX1); tma_prefetch_desc(&tmX2); tma_prefetch_desc(&tmW);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], kConvWarps); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], kEpiWarps); }
        mbar_init(&w_bar, 1);
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
    pdl_trigger();                                               // the next kernel of the stream may be scheduled from here on
    const uint32_t tmem_base = tmem_base_slot;
    long long prof_acc[4] = {0, 0, 0, 0};
    const long long prof_t0 = clock64();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // the layer's weights, once: rows of 128 B = [kh][chunk][n-atom][8 k][32 n], already TF32-rounded
            const int nbox = (p.w_rows + kWBox - 1) / kWBox;
            mbar_arrive_expect_tx(&w_bar, (uint32_t)(nbox * kWBox * 128));
            for (int i = 0; i < nbox; ++i) tma_load_2d(wsm + (size_t)i * kWBox * 128, &tmW, 0, i * kWBox, &w_bar);
            // everything above (barriers, TMEM, bias, the layer's weights) is independent of the previous kernel; the
            // activations are not.  All other roles touch global memory only after data this thread loads below has
            // arrived (epilogue stores, the addend), so this one wait orders the whole CTA behind the predecessor.
            pdl_wait();
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int t = tile;
                const int x0 = (t % p.tw) * p.vw - p.pad; t /= p.tw;      // multiple of 4 columns (TMA alignment)
                const int h0 = (t % p.th) * (4 * p.G); t /= p.th;
                const int b = t;
                for (int ck = 0; ck < p.nck; ++ck) {
                    const long long tq0 = clock64();
                    mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);
                    prof_acc[0] += clock64() - tq0;
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)tile_bytes);
                    // the input is the channel concatenation of up to three tensors (each padded to whole chunks)
                    const CUtensorMap *tm = ck < p.ck1 ? &tmX : (ck < p.ck2 ? &tmX1 : &tmX2);
                    const int cl = ck < p.ck1 ? ck : (ck < p.ck2 ? ck - p.ck1 : ck - p.ck2);
                    tma_load_4d(ring + (size_t)s * stage_bytes, tm, x0, cl * 8, h0 - p.dil, b, &full_bar[s]);
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_tf32_mn(p.N);
        // kind::f16: a/b format F16 (0), both MN-major, fp32 accumulate
        const uint32_t idesc16 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.N >> 3) << 17) | (8u << 24);
        const uint32_t ring_base = smem_u32(ring);
        const uint32_t w_base = smem_u32(wsm);
        const uint32_t empty_base = smem_u32(&empty_bar[0]);
        // descriptor steps in 16-byte units: kh*dil image rows of the tile / one row tap of the weights / lo parts
        const uint32_t a_step = (uint32_t)(p.dil * kRowBlock) >> 4, b_step = (uint32_t)(p.nck * p.natoms * kRowBlock) >> 4;
        const uint32_t a_lo = (uint32_t)p.lo_off >> 4, b_lo = (uint32_t)p.w_lo_off >> 4;
        mbar_wait(&w_bar, 0);
        int s = 0; uint32_t ph = 0; int j = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
            const int slot = j & 1;
            long long tq0 = clock64();
            mbar_wait(&tmem_empty_bar[slot], ((uint32_t)(j >> 1) & 1u) ^ 1u);
            prof_acc[0] += clock64() - tq0;
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(slot * acc_stride);
            for (int ck = 0; ck < p.nck; ++ck) {
                tq0 = clock64();
                mbar_wait(&ready_bar[s], ph);
                prof_acc[1] += clock64() - tq0;
                tc_fence_after();
                const uint32_t sa = ring_base + (uint32_t)(s * stage_bytes);
                // per row group: the three row taps (x3 in split mode) behind one election
                const uint64_t da0 = make_desc_mn(sa);
                const uint64_t db0 = make_desc_mn(w_base + (uint32_t)(ck * p.natoms * kRowBlock));
                const uint32_t accum = ck != 0 ? 1u : 0u;
                const int Gn = (p.dbg & 4) ? 0 : p.G;
                if (p.split == 2) {
#pragma unroll 1
                    for (int g = 0; g < Gn; ++g)
                        umma_kh3_split16(acc + (uint32_t)(g * p.N), da0 + (uint64_t)(g * (4 * kRowBlock >> 4)), db0, a_step, b_step,
                                         a_lo, b_lo, idesc, idesc16, accum);
                } else if (p.split) {
#pragma unroll 1
                    for (int g = 0; g < Gn; ++g)
                        umma_tf32_kh3_split(acc + (uint32_t)(g * p.N), da0 + (uint64_t)(g * (4 * kRowBlock >> 4)), db0, a_step, b_step,
                                            a_lo, b_lo, idesc, accum);
                } else {
#pragma unroll 1
                    for (int g = 0; g < Gn; ++g)
                        umma_tf32_kh3(acc + (uint32_t)(g * p.N), da0 + (uint64_t)(g * (4 * kRowBlock >> 4)), db0, a_step, b_step,
                                      idesc, accum);
                }
                umma_commit_elect(empty_base + (uint32_t)(s * 8));
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
            umma_commit_elect(smem_u32(&tmem_full_bar[slot]));
        }
    } else if (warp < 2 + kConvWarps) {
        // ===================== converters: round the landed tile to TF32 in place =====================
        const int ctid = threadIdx.x - 64;
        const int n16 = tile_bytes >> 4;                          // 16-byte words in a tile (RH * 64)
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (int ck = 0; ck < p.nck; ++ck) {
                const long long tq0 = clock64();
                mbar_wait(&full_bar[s], ph);
                prof_acc[0] += clock64() - tq0;
                uint4 *st = reinterpret_cast<uint4 *>(ring + (size_t)s * stage_bytes);
                if (p.split == 2) {
                    // fp16 correction operand beside the raw tile, one image row (1 KB) per warp instruction: lane = (channel c,
                    // pixel octet u).  Raw row: [8 ch][32 px] fp32, the 32-byte unit u of channel c at unit u ^ (c & 3)
                    // (TMA's SWIZZLE_128B_ATOM_32B); fp16 row: K atom 0 = 2^11 * lo(x), K atom 1 = fp16(x), each [8 ch][32 px]
                    // halves, the 16-byte unit u of channel c at unit u ^ ((c >> 1) & 3) (SWIZZLE_64B).  The two 16-byte
                    // halves of the 32-byte read are taken in the order (c & 1, c & 1 ^ 1), which keeps a quarter warp's
                    // two channels on different banks; batches of three rows, loads before stores.
                    const int c = lane >> 2, u = lane & 3, h0 = (c & 1) << 4;
                    const unsigned char *raw = reinterpret_cast<const unsigned char *>(st) + c * 128 + ((u ^ (c & 3)) << 5);
                    unsigned char *a2 = reinterpret_cast<unsigned char *>(st) + p.lo_off + c * 64 + ((u ^ ((c >> 1) & 3)) << 4);
                    const int cw = warp - 2;
                    for (int r0 = cw; r0 < ((p.dbg & 1) ? 0 : p.RH); r0 += 3 * kConvWarps) {
                        uint4 va[3], vb[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (r0 + k * kConvWarps < p.RH) {
                                va[k] = *reinterpret_cast<const uint4 *>(raw + (r0 + k * kConvWarps) * kRowBlock + h0);
                                vb[k] = *reinterpret_cast<const uint4 *>(raw + (r0 + k * kConvWarps) * kRowBlock + (h0 ^ 16));
                            }
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (r0 + k * kConvWarps < p.RH) {
                                const uint4 v0 = h0 ? vb[k] : va[k], v1 = h0 ? va[k] : vb[k];    // pixels 0-3, 4-7 of the octet
                                uint4 l, h;
                                l.x = lo16x2(v0.x, v0.y); l.y = lo16x2(v0.z, v0.w); l.z = lo16x2(v1.x, v1.y); l.w = lo16x2(v1.z, v1.w);
                                h.x = hi16x2(v0.x, v0.y); h.y = hi16x2(v0.z, v0.w); h.z = hi16x2(v1.x, v1.y); h.w = hi16x2(v1.z, v1.w);
                                *reinterpret_cast<uint4 *>(a2 + (r0 + k * kConvWarps) * kRowBlock) = l;
                                *reinterpret_cast<uint4 *>(a2 + (r0 + k * kConvWarps) * kRowBlock + 512) = h;
                            }
                    }
                } else if (p.split) {
                    // hi in place, lo = rna(x - hi) (the difference is exact in fp32) at the same swizzled offset of the lo tile
                    uint4 *sl = reinterpret_cast<uint4 *>(ring + (size_t)s * stage_bytes + p.lo_off);
#pragma unroll 4
                    for (int i = ctid; i < n16; i += kConvWarps * 32) {
                        // hi = the raw tile as the MMA reads it (kind::tf32 drops the low 13 mantissa bits: hi = trunc(x), no
                        // store needed); lo = rna(x - trunc(x)), exact difference, stored beside it
                        const uint4 v = st[i];
                        uint4 l;
                        l.x = rna_tf32(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u)));
                        l.y = rna_tf32(__float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u)));
                        l.z = rna_tf32(__float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u)));
                        l.w = rna_tf32(__float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u)));
                        sl[i] = l;
                    }
                } else {
#pragma unroll 4
                    for (int i = ctid; i < ((p.dbg & 1) ? 0 : n16); i += kConvWarps * 32) {
                        uint4 v = st[i];
                        v.x = rna_tf32(v.x); v.y = rna_tf32(v.y); v.z = rna_tf32(v.z); v.w = rna_tf32(v.w);
                        st[i] = v;
                    }
                }
                fence_proxy_async_smem();                         // generic-proxy writes -> visible to the MMA's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_bar[s]);
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 6..21): one warp per image row of a row group =====================
        if (p.CP & 4) epilogue<4>(p, bias_s, tmem_full_bar, tmem_empty_bar, tmem_base, acc_stride, warp, lane, prof_acc);
        else epilogue<8>(p, bias_s, tmem_full_bar, tmem_empty_bar, tmem_base, acc_stride, warp, lane, prof_acc);
    }
    if (p.prof && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6)) {
        const int r = warp == 0 ? 0 : warp == 1 ? 1 : warp == 2 ? 2 : 3;      // producer, MMA, converter, epilogue
        p.prof[4 * r + 0] = clock64() - prof_t0;       // cycles alive
        p.prof[4 * r + 1] = prof_acc[0];               // waiting: empty stage | TMEM slot | landed stage | full accumulator
        p.prof[4 * r + 2] = prof_acc[1];               // MMA warp: waiting for a rounded stage
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// Shape support and tile plan (host).  Returns false when the layer does not fit this kernel.
static bool plan(int Cin, int Cout, int H, int W, int dil, int B, int split, Params &p, size_t &smem)
{
    if (split < 0 || split > 2) return false;
    p.split = split;
    if (Cin < 1 || Cout < 1 || dil < 1 || dil > 12 || (W & 3) != 0) return false;   // TMA strides: multiples of 16 B
    p.nck = (Cin + 7) / 8;
    p.CP = Cout <= 4 ? 4 : (Cout + 7) / 8 * 8;
    p.N = (3 * p.CP + 31) / 32 * 32;
    if (p.N > 256) return false;
    p.natoms = p.N / 32;
    p.w_lo_off = 3 * p.nck * p.natoms * 8 * 128;                 // hi part: a whole number of 1 KB blocks
    p.w_rows = 3 * p.nck * p.natoms * 8 * (p.split ? 2 : 1);
    p.w_bytes = (int)round_up((size_t)p.w_rows * 128, (size_t)kWBox * 128);
    if (p.w_bytes > 160 * 1024) return false;
    p.dil = dil;
    p.vw = (32 - 2 * dil) / 4 * 4;
    p.pad = (dil + 3) / 4 * 4;
    // tile t writes columns [t*vw - pad + dil, t*vw - pad + dil + vw)
    p.tw = (W + p.pad - dil + p.vw - 1) / p.vw;
    const int sms = sm_count_cached();
    int gmax = 256 / p.N;                                        // two slots of G*N columns in 512
    if (gmax > 8) gmax = 8;
    if (gmax < 1) return false;
    // tallest tile that still leaves every SM at least two tiles (small images: shorter tiles, more of them)
    int G = gmax;
    while (G > 1 && (long long)B * p.tw * ((H + 4 * G - 1) / (4 * G)) < 2ll * sms) G >>= 1;
    p.G = G;
    // the ring must hold at least two stages beside the resident weights (split: hi + lo tile per stage)
    const size_t avail = 226 * 1024 - 1024 - (size_t)p.w_bytes;
    while (G > 1 && (size_t)(4 * G + 2 * dil) * kRowBlock * (p.split ? 2 : 1) * 2 > avail) G >>= 1;
    p.G = G;
    p.RH = 4 * G + 2 * dil;
    if (p.RH > 256) return false;
    p.lo_off = p.RH * kRowBlock;
    p.th = (H + 4 * G - 1) / (4 * G);
    const long long tiles = (long long)B * p.tw * p.th;
    if (tiles >= (1ll << 31)) return false;
    p.num_tiles = (int)tiles;
    const size_t stage = (size_t)p.RH * kRowBlock * (p.split ? 2 : 1);
    p.stages = (int)(avail / stage);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages < 2) return false;
    const int cols = 2 * G * p.N;
    p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    smem = (size_t)p.w_bytes + (size_t)p.stages * stage + 1024;
    return true;
}

}  // namespace conv2dtc
}  // namespace decnet

using namespace decnet;
using namespace decnet::conv2dtc;

extern "C" {

static thread_local int g_conv2dtc_dbg = 0;
static thread_local long long *g_conv2dtc_prof = nullptr;
void decnet_conv2d_tf32_debug(int flags, void *prof16) { g_conv2dtc_dbg = flags; g_conv2dtc_prof = static_cast<long long *>(prof16); }

int decnet_conv2d_tc_supported(int Cin, int Cout, int H, int W, int dilation, int split)
{
    Params p{};
    size_t smem = 0;
    return plan(Cin, Cout, H, W, dilation, 1, split, p, smem) ? 1 : 0;
}

int decnet_conv2d_tc_packed_floats(int Cin, int Cout, int split)
{
    const int nck = (Cin + 7) / 8, CP = Cout <= 4 ? 4 : (Cout + 7) / 8 * 8, N = (3 * CP + 31) / 32 * 32;
    return 3 * nck * (N / 32) * 8 * 32 * (split ? 2 : 1);
}

int decnet_conv2d_tf32_supported(int Cin, int Cout, int H, int W, int dilation)
{
    return decnet_conv2d_tc_supported(Cin, Cout, H, W, dilation, 0);
}

int decnet_conv2d_tf32_packed_floats(int Cin, int Cout) { return decnet_conv2d_tc_packed_floats(Cin, Cout, 0); }

int decnet_conv2d_tc_nchw_cat(const float *const *srcs, const int *src_channels, int nsrc, const float *w_packed,
                              const float *bias_padded, float *out, int B, int Cout, int H, int W, int dilation,
                              int relu, int w_valid, int split, void *stream)
{
    return decnet_conv2d_tc_nchw_cat_add(srcs, src_channels, nsrc, w_packed, bias_padded, nullptr, out, B, Cout, H, W, dilation,
                                         relu, w_valid, split, stream);
}

int decnet_conv2d_tc_nchw_cat_add(const float *const *srcs, const int *src_channels, int nsrc, const float *w_packed,
                                  const float *bias_padded, const float *addend, float *out, int B, int Cout, int H, int W,
                                  int dilation, int relu, int w_valid, int split, void *stream)
{
    DECNET_REQUIRE(!addend || Cout == 1, "addend only for single-channel outputs");
    DECNET_REQUIRE(w_valid >= 0 && w_valid <= W, "w_valid=%d outside [0, W=%d]", w_valid, W);
    DECNET_REQUIRE(srcs && src_channels && w_packed && bias_padded && out, "null pointer");
    DECNET_REQUIRE(nsrc >= 1 && nsrc <= 3, "1..3 concatenated sources, got %d", nsrc);
    DECNET_REQUIRE(B > 0 && H > 0 && W > 0, "non-positive size");
    int cin_pad = 0, cks[3] = {0, 0, 0};
    for (int i = 0; i < nsrc; ++i) {
        DECNET_REQUIRE(srcs[i] && src_channels[i] >= 1, "source %d: null pointer or no channels", i);
        DECNET_REQUIRE((reinterpret_cast<uintptr_t>(srcs[i]) & 15u) == 0, "source %d must be 16-byte aligned", i);
        cks[i] = (src_channels[i] + 7) / 8;
        cin_pad += 8 * cks[i];
    }
    DECNET_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0, "w_packed must be 16-byte aligned");
    Params p{};
    size_t smem = 0;
    DECNET_REQUIRE(plan(cin_pad, Cout, H, W, dilation, B, split, p, smem),
                   "conv2d_tc_nchw: unsupported shape Cin=%d (padded per source) Cout=%d H=%d W=%d dilation=%d split=%d "
                   "(see decnet_conv2d_tc_supported)", cin_pad, Cout, H, W, dilation, split);
    p.ck1 = cks[0]; p.ck2 = cks[0] + cks[1];
    p.dbg = g_conv2dtc_dbg; p.prof = g_conv2dtc_prof;
    p.bias = bias_padded; p.addend = addend; p.out = out; p.B = B; p.Cout = Cout; p.H = H; p.W = W; p.relu = relu;
    p.w_valid = w_valid > 0 ? w_valid : W;
    CUtensorMap tmX[3], tmW;
    for (int i = 0; i < 3; ++i) {
        if (i >= nsrc) { tmX[i] = tmX[0]; continue; }
        const int Ci = src_channels[i];
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)Ci, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)H * W * 4, (uint64_t)W * 4, (uint64_t)Ci * H * W * 4};
        const uint32_t box[4] = {32u, 8u, (uint32_t)p.RH, 1u};
        int rc = encode_tensor_map(&tmX[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, srcs[i], dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {32u, (uint64_t)p.w_rows};
        const uint64_t strides[1] = {128u};
        const uint32_t box[2] = {32u, (uint32_t)kWBox};
        int rc = encode_tensor_map(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w_packed, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    {
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < smem) {
            DECNET_CUDA(cudaFuncSetAttribute(conv2d_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) set_for[dev] = smem;
        }
    }
    const int sms = sm_count_cached();
    const unsigned grid = (unsigned)(p.num_tiles < sms ? p.num_tiles : sms);
    DECNET_CUDA(launch_pdl(conv2d_tcgen05_kernel, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream),
                           tmX[0], tmX[1], tmX[2], tmW, p));
    return after_launch("conv2d_tcgen05_kernel");
}

int decnet_conv2d_tf32_nchw_cat(const float *const *srcs, const int *src_channels, int nsrc, const float *w_packed,
                                const float *bias_padded, float *out, int B, int Cout, int H, int W, int dilation,
                                int relu, int w_valid, void *stream)
{
    return decnet_conv2d_tc_nchw_cat(srcs, src_channels, nsrc, w_packed, bias_padded, out, B, Cout, H, W, dilation, relu,
                                     w_valid, 0, stream);
}

int decnet_conv2d_tf32_nchw(const float *x, const float *w_packed, const float *bias_padded, float *out,
                            int B, int Cin, int Cout, int H, int W, int dilation, int relu, void *stream)
{
    return decnet_conv2d_tc_nchw_cat(&x, &Cin, 1, w_packed, bias_padded, out, B, Cout, H, W, dilation, relu, 0, 0, stream);
}

}  // extern "C"
