// decnet_b200/csrc/conv2d_rows_tcgen05.cu -- second formulation of the thin 3x3 Conv2d layers (C_out <= 8,
// dilation <= 4) on NCHW fp32 with TF32 tensor cores: **pixels on the N dimension, block-Toeplitz weights on M**.
//
// conv2d_tcgen05.cu puts pixels on M (TMEM lanes), so its column taps cost three TMEM reads and two warp
// shuffles per output and its epilogue is bound by the LSU/MIO pipe.  Here
//
//   D[m = (row r, c_out), n = pixel] += A_{kw,r'}[m, k = c_in] * X_{r'}[k, n]      per input row r' and column tap kw
//
//   * X_{r'} is the B operand: input row r' of the SAME smem tile conv2d_tcgen05.cu uses ([x-block][row][8 ch][32 px],
//     MN-major SWIZZLE_128B_BASE32B, n-atom stride = RH KB), N = 96 pixels, K = 8 channels.
//   * A_{kw,r'} (128 x 8) holds w[c_out][c_in][kh][kw] in the row blocks r with r' = r + kh*d and zeros elsewhere: a
//     window of one band matrix Z_kw (K-major, no swizzle, 8-row groups 256 B apart), so "input row r'" is just a
//     start-address shift of Z by whole row groups (scripts/micro/umma_toeplitz.cu).  16 output rows x 8 channels
//     fill the 128 lanes; 13 of 16 row blocks multiply zeros -- the tensor pipe has the headroom (56 cycles per MMA,
//     3*RH MMAs per 16 x 88 output tile and channel chunk).
//   * The column tap is a column offset of the ACCUMULATOR: the product of input pixel j belongs to output pixel
//     j - (kw-1)*d.  A TMEM accumulator's column base must be even (scripts/micro/umma_dcol.cu), so kw = 0 and
//     kw = 2 (2d apart) share accumulator P02 and kw = 1 has P1:  out[j] = P1[j] + P02[j + d].
//   * Epilogue: thread = (row, c_out), TMEM columns = consecutive pixels: two tcgen05.ld.x32, adds, bias, ReLU and
//     eight 128-bit stores per 32 pixels -- no shuffles, 2x (not 3x) the output bytes read from TMEM.
//
// TMA alignment (first box column multiple of 4 pixels): tiles read columns [88t-4, 88t+92) and write [88t, 88t+88).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-5 converters (TF32 rounding in place), 6-17 epilogue (3 per TMEM lane
// quarter, one 32-column chunk each).
#include "common.cuh"
#include "tma_utils.cuh"
#include <mutex>

namespace decnet {
namespace conv2drows {

constexpr int kMaxStages = 4;
constexpr int kConvWarps = 4;
constexpr int kEpiWarps = 12;
constexpr int kThreads = 32 * (2 + kConvWarps + kEpiWarps);
constexpr int kRows = 16;                  // output rows per tile (16 rows x 8 channels = 128 accumulator lanes)
constexpr int kNI = 96;                    // input columns per tile (three 32-pixel atoms)
constexpr int kXO = 88;                    // output columns per tile
constexpr int kP02 = 96;                   // column of accumulator P02 inside a slot
constexpr int kSlot = 200;                 // TMEM columns per slot: P1 [0,96) + P02 [96, 96+96+2d <= 200)
constexpr int kGroupBytes = 256;           // one 8-row group of the band matrix: [k half][8 rows][16 B]

struct Params {
    const float *w;                        // compact weights [kh][kw][chunk][c_out 8][c 8], TF32-rounded, zero padded
    const float *bias;                     // [8]
    float *out;                            // [B][Cout][H][W]
    int B, Cout, H, W;
    int dil, nck, ck1, ck2;
    int RH, U;                             // input rows per tile (16 + 2d); row groups of one band matrix (RH + 15)
    int tw, th, num_tiles, stages;
    int relu;
    int z_bytes;                           // shared memory of the band matrices: 3 * nck * U * 256
    int dbg;                               // tuning only: 1 skip rounding, 2 skip epilogue stores, 4 skip MMAs
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// the three column taps of one input row: one election, three MMAs (the third always accumulates)
__device__ __forceinline__ void umma_tf32_x3_elect(uint32_t d2, uint32_t d1, uint32_t d0, uint64_t a2, uint64_t a1, uint64_t a0,
                                                   uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %8, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %3, %6, %7, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], %4, %6, %7, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%2], %5, %6, %7, t;\n\t}"
        ::"r"(d2), "r"(d1), "r"(d0), "l"(a2), "l"(a1), "l"(a0), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

__global__ void __launch_bounds__(kThreads, 1)
conv2d_rows_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX1,
                           const __grid_constant__ CUtensorMap tmX2, const Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t ready_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *zsm = base;                                   // band matrices Z[kw][chunk][U groups]
    unsigned char *ring = base + ((p.z_bytes + 1023) & ~1023);
    const int stage_bytes = 3 * p.RH * 1024;
    const int kStages = p.stages;

    // ---- band matrices: zero, then the three non-zero row groups (kh = 0,1,2 at group RH-1-kh*d) of each (kw, chunk)
    for (int i = threadIdx.x; i < p.z_bytes / 16; i += kThreads) reinterpret_cast<uint4 *>(zsm)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * p.nck * 64; i += kThreads) {
        const int c = i & 7, co = (i >> 3) & 7;
        int t = i >> 6;
        const int ck = t % p.nck; t /= p.nck;
        const int kw = t % 3, kh = t / 3;
        const int u = p.RH - 1 - kh * p.dil;
        float *dst = reinterpret_cast<float *>(zsm + (size_t)((kw * p.nck + ck) * p.U + u) * kGroupBytes +
                                               (c >> 2) * 128 + co * 16 + (c & 3) * 4);
        *dst = __ldg(p.w + i);
    }
    fence_proxy_async_smem();                                     // generic-proxy writes -> the MMA's async proxy

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmX1); tma_prefetch_desc(&tmX2);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], kConvWarps); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int t = tile;
                const int x0 = (t % p.tw) * kXO - 4; t /= p.tw;
                const int h0 = (t % p.th) * kRows; t /= p.th;
                const int b = t;
                for (int ck = 0; ck < p.nck; ++ck) {
                    mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    const CUtensorMap *tm = ck < p.ck1 ? &tmX : (ck < p.ck2 ? &tmX1 : &tmX2);
                    const int cl = ck < p.ck1 ? ck : (ck < p.ck2 ? ck - p.ck1 : ck - p.ck2);
                    unsigned char *dst = ring + (size_t)s * stage_bytes;
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                        tma_load_4d(dst + (size_t)a * p.RH * 1024, tm, x0 + 32 * a, cl * 8, h0 - p.dil, b, &full_bar[s]);
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // c_format F32, a/b TF32, A K-major (bit 15 = 0), B MN-major (bit 16), N = 96, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kNI >> 3) << 17) | (8u << 24);
        // A: K-major, no swizzle: LBO 128 B (K halves), SBO 256 B (8-row groups).  B: MN-major SWIZZLE_128B_BASE32B:
        // LBO = n-atom stride RH KB, SBO 512 B (k-atoms).
        const uint64_t da_hi = (8ull << 16) | (16ull << 32) | (1ull << 46);
        const uint64_t db_hi = ((uint64_t)(p.RH * 64) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
        const uint32_t ring_base = smem_u32(ring);
        const uint32_t z_base = smem_u32(zsm);
        const uint32_t empty_base = smem_u32(&empty_bar[0]);
        int s = 0; uint32_t ph = 0; int j = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
            const int slot = j & 1;
            mbar_wait(&tmem_empty_bar[slot], ((uint32_t)(j >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t acc1 = tmem_base + (uint32_t)(slot * kSlot);
            const uint32_t acc02 = acc1 + kP02;
            for (int ck = 0; ck < p.nck; ++ck) {
                mbar_wait(&ready_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = ring_base + (uint32_t)(s * stage_bytes);
                // descriptors advance by constants: B one image row (1 KB) up, the band-matrix windows one row group down
                uint64_t db = db_hi | (uint64_t)((sa >> 4) & 0x3FFFu);
                const uint32_t ztop = (uint32_t)(p.RH - 1) * kGroupBytes;
                uint64_t a2 = da_hi | (uint64_t)(((z_base + (uint32_t)((2 * p.nck + ck) * p.U) * kGroupBytes + ztop) >> 4) & 0x3FFFu);
                uint64_t a1 = da_hi | (uint64_t)(((z_base + (uint32_t)((1 * p.nck + ck) * p.U) * kGroupBytes + ztop) >> 4) & 0x3FFFu);
                uint64_t a0 = da_hi | (uint64_t)(((z_base + (uint32_t)((0 * p.nck + ck) * p.U) * kGroupBytes + ztop) >> 4) & 0x3FFFu);
                uint32_t first = ck == 0 ? 0u : 1u;
                for (int r = 0; r < ((p.dbg & 4) ? 0 : p.RH); ++r) {
                    // kw = 2 first: it is the MMA that initialises P02 (columns [0,96)); kw = 0 lands 2d columns higher
                    umma_tf32_x3_elect(acc02, acc1, acc02 + (uint32_t)(2 * p.dil), a2, a1, a0, db, idesc, first);
                    first = 1u;
                    db += 64; a2 -= 16; a1 -= 16; a0 -= 16;
                }
                umma_commit_elect(empty_base + (uint32_t)(s * 8));
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
            umma_commit_elect(smem_u32(&tmem_full_bar[slot]));
        }
    } else if (warp < 2 + kConvWarps) {
        // ===================== converters: round the landed tile to TF32 (nearest) in place =====================
        const int ctid = threadIdx.x - 64;
        const int n16 = stage_bytes >> 4;
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (int ck = 0; ck < p.nck; ++ck) {
                mbar_wait(&full_bar[s], ph);
                uint4 *st = reinterpret_cast<uint4 *>(ring + (size_t)s * stage_bytes);
#pragma unroll 4
                for (int i = ctid; i < ((p.dbg & 1) ? 0 : n16); i += kConvWarps * 32) {
                    uint4 v = st[i];
                    v.x = (v.x + 0x1000u) & 0xFFFFE000u; v.y = (v.y + 0x1000u) & 0xFFFFE000u;
                    v.z = (v.z + 0x1000u) & 0xFFFFE000u; v.w = (v.w + 0x1000u) & 0xFFFFE000u;
                    st[i] = v;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_bar[s]);
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue: thread = (output row, channel), TMEM columns = pixels =====================
        const int e = warp - (2 + kConvWarps);
        const int q = warp & 3;                                   // TMEM lane quarter: output rows 4q .. 4q+3
        // the three warps of a quarter take one 32-column chunk each (the last one has 24 useful columns)
        int chunk = 0;
        {   // rank of this warp among the warps with the same (warp & 3)
            const int first = 2 + kConvWarps;
            for (int wq = first; wq < warp; ++wq) chunk += ((wq & 3) == q) ? 1 : 0;
        }
        (void)e;
        const int r = 4 * q + (lane >> 3), co = lane & 7;
        const int j0 = 4 + 32 * chunk;                            // first input-column index of this chunk
        const int ncols = chunk == 2 ? kXO - 64 : 32;
        const float bias = co < p.Cout ? __ldg(p.bias + co) : 0.f;
        const float floor_ = p.relu ? 0.f : -INFINITY;
        const size_t plane = (size_t)p.H * p.W;
        int tx, ty, tb;
        { int t = blockIdx.x; tx = t % p.tw; t /= p.tw; ty = t % p.th; tb = t / p.th; }
        const int dx = gridDim.x % p.tw, dy = (gridDim.x / p.tw) % p.th, db = gridDim.x / (p.tw * p.th);
        int j = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
            const int col0 = tx * kXO - 4 + j0;                   // image column of this chunk's first output
            const int row = ty * kRows + r;
            const int slot = j & 1;
            mbar_wait_relaxed(&tmem_full_bar[slot], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * kSlot + j0);
            uint32_t a[32], c[32];
            tmem_ld32(t1, a);
            tmem_ld32(t1 + (uint32_t)(kP02 + p.dil), c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // the accumulators are in registers: hand the slot back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[slot]);
            if (row < p.H && co < p.Cout && !(p.dbg & 2)) {
                float *op = p.out + ((size_t)tb * p.Cout + co) * plane + (size_t)row * p.W + col0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (i < ncols && col0 + i < p.W) {
                        float4 o;
                        o.x = fmaxf(__uint_as_float(a[i]) + __uint_as_float(c[i]) + bias, floor_);
                        o.y = fmaxf(__uint_as_float(a[i + 1]) + __uint_as_float(c[i + 1]) + bias, floor_);
                        o.z = fmaxf(__uint_as_float(a[i + 2]) + __uint_as_float(c[i + 2]) + bias, floor_);
                        o.w = fmaxf(__uint_as_float(a[i + 3]) + __uint_as_float(c[i + 3]) + bias, floor_);
                        *reinterpret_cast<float4 *>(op + i) = o;
                    }
                }
            }
            tx += dx; if (tx >= p.tw) { tx -= p.tw; ++ty; }
            ty += dy; if (ty >= p.th) { ty -= p.th; ++tb; }
            tb += db;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

static bool plan(int cin_pad, int Cout, int H, int W, int dil, int B, Params &p, size_t &smem)
{
    if (cin_pad < 8 || (cin_pad & 7) || Cout < 1 || Cout > 8 || dil < 1 || dil > 4 || (W & 3) != 0 || H < 1) return false;
    p.nck = cin_pad / 8;
    p.dil = dil;
    p.RH = kRows + 2 * dil;
    p.U = p.RH + kRows - 1;
    p.z_bytes = 3 * p.nck * p.U * kGroupBytes;
    const size_t z_al = ((size_t)p.z_bytes + 1023) & ~(size_t)1023;
    const size_t stage = (size_t)3 * p.RH * 1024;
    if (z_al + 2 * stage > 225 * 1024) return false;
    p.stages = (int)((225 * 1024 - z_al) / stage);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.tw = (W + kXO - 1) / kXO;
    p.th = (H + kRows - 1) / kRows;
    const long long tiles = (long long)B * p.tw * p.th;
    if (tiles >= (1ll << 31)) return false;
    p.num_tiles = (int)tiles;
    smem = z_al + (size_t)p.stages * stage + 1024;
    return true;
}

}  // namespace conv2drows
}  // namespace decnet

using namespace decnet;
using namespace decnet::conv2drows;

extern "C" {

static thread_local int g_rows_dbg = 0;
void decnet_conv2d_tf32_rows_debug(int flags) { g_rows_dbg = flags; }

int decnet_conv2d_tf32_rows_supported(int Cin_padded, int Cout, int H, int W, int dilation)
{
    Params p{};
    size_t smem = 0;
    return plan(Cin_padded, Cout, H, W, dilation, 1, p, smem) ? 1 : 0;
}

int decnet_conv2d_tf32_rows_nchw_cat(const float *const *srcs, const int *src_channels, int nsrc, const float *w_compact,
                                     const float *bias8, float *out, int B, int Cout, int H, int W, int dilation,
                                     int relu, void *stream)
{
    DECNET_REQUIRE(srcs && src_channels && w_compact && bias8 && out, "null pointer");
    DECNET_REQUIRE(nsrc >= 1 && nsrc <= 3, "1..3 concatenated sources, got %d", nsrc);
    DECNET_REQUIRE(B > 0 && H > 0 && W > 0, "non-positive size");
    int cin_pad = 0, cks[3] = {0, 0, 0};
    for (int i = 0; i < nsrc; ++i) {
        DECNET_REQUIRE(srcs[i] && src_channels[i] >= 1, "source %d: null pointer or no channels", i);
        DECNET_REQUIRE((reinterpret_cast<uintptr_t>(srcs[i]) & 15u) == 0, "source %d must be 16-byte aligned", i);
        cks[i] = (src_channels[i] + 7) / 8;
        cin_pad += 8 * cks[i];
    }
    DECNET_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "out must be 16-byte aligned");
    Params p{};
    size_t smem = 0;
    DECNET_REQUIRE(plan(cin_pad, Cout, H, W, dilation, B, p, smem),
                   "conv2d_tf32_rows: unsupported shape Cin=%d (padded per source) Cout=%d H=%d W=%d dilation=%d", cin_pad,
                   Cout, H, W, dilation);
    p.ck1 = cks[0]; p.ck2 = cks[0] + cks[1];
    p.dbg = g_rows_dbg;
    p.w = w_compact; p.bias = bias8; p.out = out; p.B = B; p.Cout = Cout; p.H = H; p.W = W; p.relu = relu;
    CUtensorMap tmX[3];
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
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < smem) {
            DECNET_CUDA(cudaFuncSetAttribute(conv2d_rows_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) set_for[dev] = smem;
        }
    }
    const int sms = sm_count_cached();
    const unsigned grid = (unsigned)(p.num_tiles < sms ? p.num_tiles : sms);
    conv2d_rows_tcgen05_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmX[0], tmX[1], tmX[2], p);
    return after_launch("conv2d_rows_tcgen05_kernel");
}

}  // extern "C"
