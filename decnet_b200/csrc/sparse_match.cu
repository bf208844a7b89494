// decnet_b200/csrc/sparse_match.cu -- SparseMatching / SparseVar for sm_100a.
//
// Replaces the reference's SM_kernel.cu / SV_kernel.cu.
// Design (not a translation; the reference is thread-per-pixel with every cost
// recomputed by four separate kernels):
//
//   * Both ops are ROW-LOCAL: pixel (b,h,w) only touches row (b,h) of the two
//     feature maps.  One CTA owns one row.
//   * Mask compaction: per 32-column chunk one __ballot_sync gives the mask bits,
//     __popc the chunk count, a warp scan over chunk counts the offsets; the
//     sorted list of valid right columns makes the candidate set of pixel w the
//     CONTIGUOUS list range [prefix(max(0,w-D+1)), prefix(w+1)) -- bit-exact with
//     the reference's `tar_mask[w-d] != 0` scan, and work is proportional to
//     density^2 instead of D.
//   * The listed columns of both views are transposed into compact [j][C] operand
//     buffers in shared memory.  Default (sparse_row_gather_kernel): gathered straight
//     from global memory through the sorted lists -- only the sectors that hold a listed
//     column are read, four CTAs per SM.  Alternative (sparse_row_kernel): the whole
//     [C,W] rows staged first by TMA or cp.async.  DESIGN.md section 3.1 has the
//     measurements of both and of the persistent / pipelined forms kept behind
//     decnet_set_sparse_variant().
//   * A group of G lanes (G picked per row from the candidate density) owns one
//     masked pixel; lanes own candidates; the channel dot product is the same
//     sequential FMA chain as the reference (bit-identical costs); softmax is one
//     pass (online max) with fp64 moments, so SpaVar around the final mean costs no
//     second pass; max / sum reductions are warp shuffles.
//   * Outputs are written completely (the row is zero-filled with 128-bit stores,
//     masked pixels stored on top), so callers need no memset.
#include "common.cuh"
#include "sparse_core.cuh"
#include "tma_utils.cuh"
#include <algorithm>
#include <cstring>
#include <mutex>

namespace decnet {
namespace sparse {

// -------------------------------------------------------------------------------------
// Forward row kernel: a CTA walks rows  blockIdx.x, blockIdx.x + gridDim.x, ...
//   <NT = 256, NB = 3>, grid = B*H : one row per CTA, 3 CTAs/SM (62 KB of slabs + 8 KB of lists per
//                    CTA at the SceneFlow shapes), so that while one CTA evaluates its row two more
//                    have their slabs in flight.
//   <NT = 384, NB = 2>, grid = 2 * #SM : PERSISTENT.  Half an SM's shared memory per CTA leaves room
//                    to compact BOTH operands of every realistic row, so the raw slabs are dead as
//                    soon as the two gathers are done: the next row's slabs (and its two mask rows,
//                    cp.async into a small staging buffer) are put in flight right there, BEFORE the
//                    cost / softmax phase, which then runs entirely under the next row's load.
//   USE_TMA = true : each operand's [C, W] slab arrives as ceil(W/256) TMA boxes
//                    {bw, 1, C} of the 3-D view (W, H, B*C); one elected thread issues them
//                    and the CTA waits on a single mbarrier (complete_tx::bytes).
//   USE_TMA = false: cp.async (16 B when W % 4 == 0 and pointers are 16-B aligned, else
//                    4 B) for shapes TMA cannot describe (e.g. KITTI's odd widths).
// -------------------------------------------------------------------------------------
struct RowArgs {
    const float *L, *R, *lmask, *rmask, *disp_in;
    float *out_a;      // MAT: out   VAR: var   FUSED: out
    float *out_b;      // FUSED: var, else unused
    float *sum_sim, *max_cost;
    int C, H, W, D, nrows;
    int vec_ok, mvec_ok;              // 128-bit access allowed on features+outputs / on the mask rows
    int bw, nchunks, chunk_stride;    // TMA box geometry
    uint32_t bw_magic;
    int rc_cap;                       // floats available for the compacted operands
    int stage_masks;                  // masks go through the shared staging buffer (persistent mode)
};

template <bool USE_TMA, int NT>
__device__ __forceinline__ void issue_slabs(const CUtensorMap &tmL, const CUtensorMap &tmR, const RowArgs &a,
                                            float *Ls, float *Rs, uint64_t *mbar, int row, int tid)
{
    const int b = row / a.H, h = row - b * a.H;
    if (USE_TMA) {
        if (tid == 0) {
            mbar_arrive_expect_tx(mbar, (uint32_t)(2 * a.nchunks * a.C * a.bw * 4));
            for (int ch = 0; ch < a.nchunks; ++ch) {
                tma_load_3d(Ls + ch * a.chunk_stride, &tmL, ch * a.bw, h, b * a.C, mbar);
                tma_load_3d(Rs + ch * a.chunk_stride, &tmR, ch * a.bw, h, b * a.C, mbar);
            }
        }
    } else {
        const int lane = tid & 31, warp = tid >> 5;
        const int W = a.W, Wp = (W + 3) & ~3;
        const size_t plane = (size_t)a.H * W;
        const float *Lrow = a.L + (size_t)b * a.C * plane + (size_t)h * W;
        const float *Rrow = a.R + (size_t)b * a.C * plane + (size_t)h * W;
        if (a.vec_ok) {
            const int w4n = W >> 2;
            for (int c = warp; c < a.C; c += NT / 32) {
                const float *ls = Lrow + (size_t)c * plane, *rs = Rrow + (size_t)c * plane;
                float *ld = Ls + c * Wp, *rd = Rs + c * Wp;
                for (int w4 = lane; w4 < w4n; w4 += 32) {
                    cp_async_16(ld + 4 * w4, ls + 4 * w4);
                    cp_async_16(rd + 4 * w4, rs + 4 * w4);
                }
            }
        } else {
            for (int c = warp; c < a.C; c += NT / 32) {
                const float *ls = Lrow + (size_t)c * plane, *rs = Rrow + (size_t)c * plane;
                float *ld = Ls + c * Wp, *rd = Rs + c * Wp;
                for (int w = lane; w < W; w += 32) { cp_async_4(ld + w, ls + w); cp_async_4(rd + w, rs + w); }
            }
        }
    }
}

// both mask rows of `row` -> Ms[0..Wp) (left), Ms[Wp..2Wp) (right)
template <int NT>
__device__ __forceinline__ void issue_masks(const RowArgs &a, float *Ms, int row, int tid)
{
    const int W = a.W, Wp = (W + 3) & ~3;
    const float *lm = a.lmask + (size_t)row * W, *rm = a.rmask + (size_t)row * W;
    if (a.mvec_ok) {
        for (int w4 = tid; w4 < (W >> 2); w4 += NT) {
            cp_async_16(Ms + 4 * w4, lm + 4 * w4);
            cp_async_16(Ms + Wp + 4 * w4, rm + 4 * w4);
        }
    } else {
        for (int w = tid; w < W; w += NT) { cp_async_4(Ms + w, lm + w); cp_async_4(Ms + Wp + w, rm + w); }
    }
}

template <int MODE, bool USE_TMA, int NT, int NB>
__global__ void __launch_bounds__(NT, NB)
sparse_row_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ RowArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x;
    const int C = a.C, W = a.W, D = a.D;
    const int Wp = (W + 3) & ~3;
    const int Cp = (C + 3) & ~3;
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int tile_floats = USE_TMA ? a.nchunks * a.chunk_stride : C * Wp;
    float *Ls = reinterpret_cast<float *>(base);
    float *Rs = Ls + tile_floats;
    RowSmem s;
    float *Ms = reinterpret_cast<float *>(carve_lists(s, reinterpret_cast<unsigned char *>(Rs + tile_floats), W));
    float *Rc = a.stage_masks ? Ms + 2 * Wp : Ms;
    const int cs = USE_TMA ? a.bw : Wp;

    if (USE_TMA) {
        if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
        __syncthreads();
    }
    uint32_t parity = 0;
    bool inflight = false;            // this row's slabs (and staged masks) were issued during the previous row

    for (int row = blockIdx.x; row < a.nrows; row += gridDim.x) {
        const size_t m0 = (size_t)row * W;
        // 1. put the whole [C,W] slabs of both views in flight (masks first: they are needed first)
        if (!inflight) {
            if (a.stage_masks) { issue_masks<NT>(a, Ms, row, tid); cp_async_commit(); }
            issue_slabs<USE_TMA, NT>(tmL, tmR, a, Ls, Rs, &mbar, row, tid);
            if (!USE_TMA) cp_async_commit();
        }
        // 2. while they fly: compact both masks, zero the output rows
        if (a.stage_masks) {
            if (USE_TMA) cp_async_wait_all();
            else asm volatile("cp.async.wait_group 1;\n" ::: "memory");
            __syncthreads();
            compact_row_masks<true>(s, Ms, Ms + Wp, W, tid, NT, USE_TMA ? a.bw : 0, a.chunk_stride, a.bw_magic, D, C);
        } else {
            compact_row_masks<false>(s, a.lmask + m0, a.rmask + m0, W, tid, NT, USE_TMA ? a.bw : 0, a.chunk_stride,
                                     a.bw_magic, D, C);
        }
        zero_rows(a.out_a + m0, a.sum_sim + m0, a.max_cost + m0, MODE == MODE_FUSED ? a.out_b + m0 : nullptr,
                  W, a.vec_ok, tid, NT);
        // The slabs must have landed before this CTA may retire or reuse the buffers, even when the
        // row has no masked pixel and the zeros are already the answer.
        if (USE_TMA) { if (tid == 0) mbar_wait(&mbar, parity); parity ^= 1u; }
        else cp_async_wait_all();
        __syncthreads();   // slabs visible to all; also orders the zero fill before the result stores
        inflight = false;
        const int nL = s.counts[0], nR = s.counts[1];
        const int next = row + (int)gridDim.x;
        if (nL != 0) {
            // 3. costs / softmax regression / variance for every masked pixel, stored straight to global
            const float *disp_row = MODE == MODE_VAR ? a.disp_in + m0 : nullptr;
            if ((nR + nL) * Cp <= a.rc_cap) {
                // both operands compacted: Rc[j][Cp] for the valid right columns, Lc[i][Cp] for the masked left pixels
                float *Lc = Rc + nR * Cp;
                gather_columns(s.rlist, nR, Rs, cs, C, Cp, Rc, tid, NT);
                gather_columns(s.llist, nL, Ls, cs, C, Cp, Lc, tid, NT);
                __syncthreads();
                if (a.stage_masks && next < a.nrows) {
                    // slabs and mask staging are dead: the next row loads under this row's arithmetic
                    issue_masks<NT>(a, Ms, next, tid); cp_async_commit();
                    if (USE_TMA && tid == 0) fence_proxy_async_smem();   // generic reads above -> async-proxy writes
                    issue_slabs<USE_TMA, NT>(tmL, tmR, a, Ls, Rs, &mbar, next, tid);
                    if (!USE_TMA) cp_async_commit();
                    inflight = true;
                }
                process_row<MODE, 2>(s, Lc, Rc, cs, C, Cp, D, disp_row,
                                     a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
            } else if (nR * Cp <= a.rc_cap) {
                gather_columns(s.rlist, nR, Rs, cs, C, Cp, Rc, tid, NT);
                __syncthreads();
                process_row<MODE, 1>(s, Ls, Rc, cs, C, Cp, D, disp_row,
                                     a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
            } else {
                process_row<MODE, 0>(s, Ls, Rs, cs, C, Cp, D, disp_row,
                                     a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
            }
        }
        if (next < a.nrows) __syncthreads();   // lists / compacted operands / slabs are free for the next row
    }
}

// -------------------------------------------------------------------------------------
// Sector-gather row kernel (variant 3, the default): no slabs at all.
// The row kernel above stages the whole [C,W] rows of both views although only the masked left pixels and
// the valid right columns are ever read, and its 62 KB of slabs limit an SM to three rows in flight while a
// row is a chain of short latency-bound phases.  Here the compacted operands Rc[j][Cp] / Lc[i][Cp] are
// gathered STRAIGHT from global memory through the sorted column lists (neighbouring list entries share
// 32-byte sectors; the read-only path coalesces them), so
//   * DRAM traffic drops with the mask density (a sector is fetched only if one of its 8 columns is listed),
//   * a CTA needs only its lists and the operand buffer: four CTAs of 256 threads per SM (64 registers) instead
//     of three, which is what hides the per-row latency chain (measured: 128-thread CTAs, 6 or 8 per SM, lose
//     more in the cost phase -- one lane per pixel -- than they gain).
// Rows whose operands do not fit the buffer (density above ~50 %) are evaluated from global memory directly.
// -------------------------------------------------------------------------------------
__device__ inline void gather_columns_global(const uint32_t *__restrict__ list, int n, const float *__restrict__ row,
                                             size_t plane, int C, int Cp, float *__restrict__ dst, int tid, int nthreads)
{
    const int q4 = Cp >> 2;
    const uint32_t magic = 0xffffffffu / (uint32_t)q4 + 1u;      // idx / q4, exact for idx < 2^16 * q4
    const int total = n * q4;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int j = (q4 == 1) ? idx : (int)__umulhi((uint32_t)idx, magic);
        const int q = idx - j * q4;
        const int c = 4 * q;
        const float *src = row + (size_t)c * plane + (list[j] & 0xffffu);
        float4 v;
        v.x = __ldg(src);
        v.y = (c + 1 < C) ? __ldg(src + plane) : 0.f;
        v.z = (c + 2 < C) ? __ldg(src + 2 * plane) : 0.f;
        v.w = (c + 3 < C) ? __ldg(src + 3 * plane) : 0.f;
        *reinterpret_cast<float4 *>(dst + j * Cp + c) = v;
    }
}

template <int MODE, int NT>
__device__ __forceinline__ void gather_row(const RowArgs &a, const int row, unsigned char *smem_raw)
{
    const int tid = threadIdx.x;
    const int C = a.C, W = a.W, D = a.D;
    const int Cp = (C + 3) & ~3;
    const int b = row / a.H, h = row - b * a.H;
    const size_t plane = (size_t)a.H * W;
    const size_t m0 = (size_t)row * W;
    const float *Lrow = a.L + (size_t)b * C * plane + (size_t)h * W;
    const float *Rrow = a.R + (size_t)b * C * plane + (size_t)h * W;
    RowSmem s;
    float *Rc = reinterpret_cast<float *>(carve_lists(s, smem_raw, W));

    compact_row_masks<false, 0, (NT <= 128 ? 8 : 4)>(s, a.lmask + m0, a.rmask + m0, W, tid, NT, 0, 0, 0, D, C);
    zero_rows(a.out_a + m0, a.sum_sim + m0, a.max_cost + m0, MODE == MODE_FUSED ? a.out_b + m0 : nullptr,
              W, a.vec_ok, tid, NT);
    const int nL = s.counts[0], nR = s.counts[1];
    if (nL == 0) return;
    const float *disp_row = MODE == MODE_VAR ? a.disp_in + m0 : nullptr;
    if ((nR + nL) * Cp <= a.rc_cap) {
        float *Lc = Rc + nR * Cp;
        gather_columns_global(s.rlist, nR, Rrow, plane, C, Cp, Rc, tid, NT);
        gather_columns_global(s.llist, nL, Lrow, plane, C, Cp, Lc, tid, NT);
        __syncthreads();   // operands complete; also orders the zero fill before the result stores
        process_row<MODE, 2>(s, Lc, Rc, 0, C, Cp, D, disp_row,
                             a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
    } else if (nR * Cp <= a.rc_cap) {
        // only the right columns fit (wide C or a dense row): each masked left pixel reads its own channel vector
        // from global memory (list entries carry the column in both halves, the plane is the channel stride)
        gather_columns_global(s.rlist, nR, Rrow, plane, C, Cp, Rc, tid, NT);
        __syncthreads();
        process_row<MODE, 1>(s, Lrow, Rc, (int)plane, C, Cp, D, disp_row,
                             a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
    } else {
        __syncthreads();
        // nothing fits: the slab addressing of mode 0 on global memory
        process_row<MODE, 0>(s, Lrow, Rrow, (int)plane, C, Cp, D, disp_row,
                             a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
    }
}

template <int MODE, int NT, int NB>
__global__ void __launch_bounds__(NT, NB)
sparse_row_gather_kernel(const __grid_constant__ RowArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();
    gather_row<MODE, NT>(a, blockIdx.x, smem_raw);
}

// The rows of SEVERAL pyramid levels in one launch (SparseDenseNetRefinementMask.py:183-192 runs the ops level after
// level; their inputs -- features and masks -- do not depend on each other).  C*W is the same at every level, so a row is
// the same amount of data everywhere, but the coarse levels have few rows (480 and 1440 against 4320 at B = 8): as launches
// of their own they are one or two waves of latency (0.20 / 0.55 of the HBM roofline against 0.63 for the finest level).
// Here the blocks of the finest level come first and the coarse rows fill the machine behind them.
constexpr int kMaxLevels = 4;
struct MultiArgs {
    RowArgs lv[kMaxLevels];
    int first_row[kMaxLevels + 1];        // block index of each level's first row (levels in launch order)
    int nlev;
};

template <int MODE, int NT, int NB>
__global__ void __launch_bounds__(NT, NB)
sparse_row_gather_multi_kernel(const __grid_constant__ MultiArgs m)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();                            // masks and features come from the kernels before this one
    int lvl = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i) lvl += (i < m.nlev && (int)blockIdx.x >= m.first_row[i]) ? 1 : 0;
    gather_row<MODE, NT>(m.lv[lvl], (int)blockIdx.x - m.first_row[lvl], smem_raw);
}

// -------------------------------------------------------------------------------------
// Software-pipelined sector-gather kernel (variant 4, opt-in: measured slower than one row per CTA).
// Same operands as sparse_row_gather_kernel, but a persistent CTA walks its rows with every global read one
// row ahead of its use, all of them cp.async so no thread ever waits for the data it has just asked for:
//   iteration k:  wait(gathers of row k, masks of row k+1)            <- issued one cost phase ago
//                 compact row k+1 from the staged masks, zero its output rows,
//                 cp.async-gather its operands into the OTHER list/operand buffer,
//                 cp.async the masks of row k+2,
//                 costs / softmax regression / variance of row k       <- runs under all of the above loads
// Rows whose operands do not fit a buffer are evaluated from global memory directly (mode 0).
// -------------------------------------------------------------------------------------
constexpr int kStreamThreads = 256;

// 4-byte cp.async gather of the listed columns: warp = channel, lane = list entry (sorted columns: neighbouring
// lanes share sectors)
__device__ __forceinline__ void gather_columns_async(const uint32_t *__restrict__ list, int n,
                                                     const float *__restrict__ row, size_t plane, int C, int Cp,
                                                     float *__restrict__ dst, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    for (int j = lane; j < n; j += 32) {
        const float *src = row + (list[j] & 0xffffu);
        float *d = dst + j * Cp;
        for (int c = warp; c < C; c += kStreamThreads / 32) cp_async_4(d + c, src + (size_t)c * plane);
    }
}

template <int MODE, int NB>
__global__ void __launch_bounds__(kStreamThreads, NB)
sparse_row_stream_kernel(const __grid_constant__ RowArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = kStreamThreads;
    const int tid = threadIdx.x;
    const int C = a.C, W = a.W, D = a.D, H = a.H;
    const int Wp = (W + 3) & ~3;
    const int Cp = (C + 3) & ~3;
    const size_t plane = (size_t)H * W;
    const size_t list_bytes = list_smem_bytes(W);
    float *Ms = reinterpret_cast<float *>(smem_raw);                      // staged mask rows [2][Wp]
    unsigned char *lists0 = reinterpret_cast<unsigned char *>(Ms + 2 * Wp);
    float *ob0 = reinterpret_cast<float *>(lists0 + 2 * list_bytes);      // operand buffers [2][rc_cap]
    const int nmine = (a.nrows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int stride = (int)gridDim.x;

    // compaction + zero fill + operand gather of the CTA's k-th row into buffer k & 1 (masks already staged)
    auto front = [&](int k) {
        const int row = (int)blockIdx.x + k * stride;
        const int b = row / H, h = row - b * H;
        const size_t m0 = (size_t)row * W;
        RowSmem s;
        carve_lists(s, lists0 + (k & 1) * list_bytes, W);
        float *Rc = ob0 + (size_t)(k & 1) * a.rc_cap;
        compact_row_masks<true>(s, Ms, Ms + Wp, W, tid, NT, 0, 0, 0, D, C);
        zero_rows(a.out_a + m0, a.sum_sim + m0, a.max_cost + m0, MODE == MODE_FUSED ? a.out_b + m0 : nullptr,
                  W, a.vec_ok, tid, NT);
        const int nL = s.counts[0], nR = s.counts[1];
        if (nL != 0 && (nR + nL) * Cp <= a.rc_cap) {
            const float *Lrow = a.L + (size_t)b * C * plane + (size_t)h * W;
            const float *Rrow = a.R + (size_t)b * C * plane + (size_t)h * W;
            gather_columns_async(s.rlist, nR, Rrow, plane, C, Cp, Rc, tid);
            gather_columns_async(s.llist, nL, Lrow, plane, C, Cp, Rc + nR * Cp, tid);
        }
        cp_async_commit();
    };

    if (nmine <= 0) return;
    issue_masks<NT>(a, Ms, (int)blockIdx.x, tid);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    front(0);
    if (nmine > 1) { issue_masks<NT>(a, Ms, (int)blockIdx.x + stride, tid); cp_async_commit(); }

    for (int k = 0; k < nmine; ++k) {
        cp_async_wait_all();
        __syncthreads();          // operands of row k and masks of row k+1 landed; everyone is done with row k-1
        if (k + 1 < nmine) {
            front(k + 1);         // its barriers also order the reads of the mask staging before the refill below
            if (k + 2 < nmine) { issue_masks<NT>(a, Ms, (int)blockIdx.x + (k + 2) * stride, tid); cp_async_commit(); }
        }
        const int row = (int)blockIdx.x + k * stride;
        const size_t m0 = (size_t)row * W;
        RowSmem s;
        carve_lists(s, lists0 + (k & 1) * list_bytes, W);
        const int nL = s.counts[0], nR = s.counts[1];
        if (nL == 0) continue;
        const float *disp_row = MODE == MODE_VAR ? a.disp_in + m0 : nullptr;
        if ((nR + nL) * Cp <= a.rc_cap) {
            const float *Rc = ob0 + (size_t)(k & 1) * a.rc_cap;
            process_row<MODE, 2>(s, Rc + nR * Cp, Rc, 0, C, Cp, D, disp_row,
                                 a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
        } else {
            const int b = row / H, h = row - b * H;
            const float *Lrow = a.L + (size_t)b * C * plane + (size_t)h * W;
            const float *Rrow = a.R + (size_t)b * C * plane + (size_t)h * W;
            process_row<MODE, 0>(s, Lrow, Rrow, (int)plane, C, Cp, D, disp_row,
                                 a.out_a + m0, a.out_b + m0, a.sum_sim + m0, a.max_cost + m0, tid, NT);
        }
    }
}

// -------------------------------------------------------------------------------------
// Candidate signature (test hook): same compaction + prefix code, no features.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
candidate_signature_kernel(const float *__restrict__ lmask, const float *__restrict__ rmask,
                           int32_t *__restrict__ count, unsigned long long *__restrict__ hash,
                           int W, int D)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const size_t m0 = (size_t)blockIdx.x * W;
    RowSmem s;
    carve_lists(s, smem_raw, W);
    compact_row_masks(s, lmask + m0, rmask + m0, W, tid, kThreads);
    for (int w = tid; w < W; w += kThreads) { count[m0 + w] = 0; hash[m0 + w] = 0ull; }
    __syncthreads();
    const int nL = s.counts[0];
    for (int i = tid; i < nL; i += kThreads) {
        const int w = (int)(s.llist[i] & 0xffffu);
        const int x0 = max(0, w - D + 1);
        const int lo = row_prefix(s, x0), hi = row_prefix(s, w + 1);
        unsigned long long hs = 0ull;
        for (int j = lo; j < hi; ++j) {
            const int d = w - (int)(s.rlist[j] & 0xffffu);
            unsigned long long x = (unsigned long long)(d + 1) * 0x9E3779B97F4A7C15ull;
            x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
            hs += x;
        }
        count[m0 + w] = max(hi - lo, 0);
        hash[m0 + w] = hs;
    }
}

// -------------------------------------------------------------------------------------
// Backward (API parity with SM_kernel.cu:143-195,300-355 and SV_kernel.cu:142-325).
// One CTA per row, one warp per masked pixel; the per-candidate weight
//   p_d = e_d * q_d   (q = d - out   or   (d-disp)^2 - var)
// is computed once per (pixel, candidate) instead of once per channel thread.
// Gradients arrive zero-filled; only masked positions are written (reference contract).
// -------------------------------------------------------------------------------------
template <int VARMODE>
__global__ void __launch_bounds__(kThreads)
sparse_row_backward_kernel(const float *__restrict__ L, const float *__restrict__ R,
                           const float *__restrict__ lmask, const float *__restrict__ rmask,
                           const float *__restrict__ disp, const float *__restrict__ outv,
                           const float *__restrict__ sum_sim, const float *__restrict__ max_cost,
                           const float *__restrict__ gout,
                           float *__restrict__ dL, float *__restrict__ dR, float *__restrict__ ddisp,
                           int C, int H, int W, int D, int vec_ok)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nWarps = kThreads / 32;
    const int row = blockIdx.x;
    const int b = row / H, h = row - b * H;
    const int Wp = (W + 3) & ~3;
    float *Ls = reinterpret_cast<float *>(smem_raw);
    float *Rs = Ls + C * Wp;
    RowSmem s;
    carve_lists(s, reinterpret_cast<unsigned char *>(Rs + C * Wp), W);
    const size_t plane = (size_t)H * W;
    const size_t f0 = (size_t)b * C * plane + (size_t)h * W;
    const float *Lrow = L + f0, *Rrow = R + f0;
    if (vec_ok) {
        const int w4n = W >> 2;
        for (int c = warp; c < C; c += nWarps) {
            for (int w4 = lane; w4 < w4n; w4 += 32) {
                cp_async_16(Ls + c * Wp + 4 * w4, Lrow + (size_t)c * plane + 4 * w4);
                cp_async_16(Rs + c * Wp + 4 * w4, Rrow + (size_t)c * plane + 4 * w4);
            }
        }
    } else {
        for (int c = warp; c < C; c += nWarps) {
            for (int w = lane; w < W; w += 32) {
                cp_async_4(Ls + c * Wp + w, Lrow + (size_t)c * plane + w);
                cp_async_4(Rs + c * Wp + w, Rrow + (size_t)c * plane + w);
            }
        }
    }
    cp_async_commit();
    const size_t m0 = (size_t)row * W;
    compact_row_masks(s, lmask + m0, rmask + m0, W, tid, kThreads);
    cp_async_wait_all();
    __syncthreads();
    // per-pixel scalars of the row come straight from global memory
    const float *o_a = outv + m0, *o_b = VARMODE ? disp + m0 : outv + m0;
    const float *o_ssim = sum_sim + m0, *o_max = max_cost + m0;
    const int nL = s.counts[0], nR = s.counts[1];

    // ---- left gradient (and SpaVar's disparity gradient): warp per masked left pixel.
    // Lanes own candidates in ascending d (descending list index); the first KB*32
    // candidate weights p = e*q stay in registers, the (rare) remainder is recomputed.
    constexpr int KB = 8;
    for (int i = warp; i < nL; i += nWarps) {
        const int w = (int)(s.llist[i] & 0xffffu);
        const int lo = row_prefix(s, max(0, w - D + 1)), hi = row_prefix(s, w + 1);
        const float mx = o_max[w], ov = o_a[w], mu = o_b[w];
        const float gw = gout[m0 + w], ss = o_ssim[w];
        auto weight = [&](int col, float &dpart) -> float {
            float cost = 0.f;
            for (int cc = 0; cc < C; ++cc) cost = fmaf(Ls[cc * Wp + w], Rs[cc * Wp + col], cost);
            const float e = expf(cost - mx);
            const float df = (float)(w - col);
            if (VARMODE) { const float dd = df - mu; dpart = e * dd; return e * (dd * dd - ov); }
            dpart = 0.f;
            return e * (df - ov);
        };
        float p[KB]; int ro[KB];
        float dacc = 0.f;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const int j = hi - 1 - (k * 32 + lane);
            p[k] = 0.f; ro[k] = 0;
            if (j >= lo) {
                ro[k] = (int)(s.rlist[j] & 0xffffu);
                float dp; p[k] = weight(ro[k], dp); dacc += dp;
            }
        }
        for (int j = hi - 1 - (KB * 32 + lane); j >= lo; j -= 32) {
            float dp; (void)weight((int)(s.rlist[j] & 0xffffu), dp); dacc += dp;
        }
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < KB; ++k) acc += p[k] * Rs[c * Wp + ro[k]];
            for (int j = hi - 1 - (KB * 32 + lane); j >= lo; j -= 32) {
                const int col = (int)(s.rlist[j] & 0xffffu);
                float dp; acc += weight(col, dp) * Rs[c * Wp + col];
            }
            acc = group_sum(acc, 32);
            if (lane == 0) dL[f0 + (size_t)c * plane + w] = gw * acc / ss;
        }
        if (VARMODE) {
            dacc = group_sum(dacc, 32);
            if (lane == 0) ddisp[m0 + w] = -2 * gw * dacc / ss;
        }
    }
    // ---- right gradient: warp per masked right pixel, gather over masked left pixels
    // wl = w + d, d in [0, min(D, W - w))  (SM_kernel.cu:327-346)
    for (int i = warp; i < nR; i += nWarps) {
        const int w = (int)(s.rlist[i] & 0xffffu);
        const int lo = left_prefix(s, w), hi = left_prefix(s, min(W, w + D));
        auto weight = [&](int wl) -> float {
            float cost = 0.f;
            for (int cc = 0; cc < C; ++cc) cost = fmaf(Ls[cc * Wp + wl], Rs[cc * Wp + w], cost);
            const float e = expf(cost - o_max[wl]);
            const float df = (float)(wl - w);
            float q;
            if (VARMODE) { const float dd = df - o_b[wl]; q = dd * dd - o_a[wl]; }
            else q = df - o_a[wl];
            return gout[m0 + wl] * e * q / o_ssim[wl];
        };
        float p[KB]; int ro[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const int j = lo + k * 32 + lane;
            p[k] = 0.f; ro[k] = 0;
            if (j < hi) { ro[k] = (int)(s.llist[j] & 0xffffu); p[k] = weight(ro[k]); }
        }
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < KB; ++k) acc += p[k] * Ls[c * Wp + ro[k]];
            for (int j = lo + KB * 32 + lane; j < hi; j += 32) {
                const int wl = (int)(s.llist[j] & 0xffffu);
                acc += weight(wl) * Ls[c * Wp + wl];
            }
            acc = group_sum(acc, 32);
            if (lane == 0) dR[f0 + (size_t)c * plane + w] = acc;
        }
    }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
static thread_local int g_last_path = 0;
static thread_local int g_forced_path = 0;
static thread_local int g_last_variant = 0;
static thread_local int g_forced_variant = 0;

static int validate_common(const void *L, const void *R, const void *ml, const void *mr,
                           int B, int C, int H, int W)
{
    DECNET_REQUIRE(L && R && ml && mr, "null input pointer");
    DECNET_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "non-positive size B=%d C=%d H=%d W=%d", B, C, H, W);
    DECNET_REQUIRE((long long)B * H < (1ll << 31) && B <= 65535, "B*H too large");
    if (W > 65535) { set_error("W=%d exceeds 65535 columns", W); return DECNET_ERR_UNSUPPORTED; }
    return 0;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct TileGeom { int use_tma, bw, nchunks, chunk_stride, rc_cap, persistent; uint32_t bw_magic; size_t smem; };

// TMA needs 16-B aligned rows (W % 4 == 0, aligned bases), box dims <= 256 and tile offsets
// that fit the 16-bit field of the packed column lists.  The leftover shared memory of the
// chosen occupancy tier becomes the buffer of the compacted operands: 3, 2 or 1 CTAs per SM for
// the one-row-per-CTA launch, half an SM for the persistent one (which also stages the mask rows).
static TileGeom plan_tiles(const float *L, const float *R, int C, int W, bool allow_tma, bool persistent)
{
    TileGeom g{};
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t Cp = (size_t)((C + 3) & ~3);
    bool tma_ok = allow_tma && (W % 4 == 0) && aligned16(L) && aligned16(R) && C <= 256;
    if (tma_ok) {
        g.nchunks = (W + 255) / 256;
        g.bw = ((W + g.nchunks - 1) / g.nchunks + 3) & ~3;
        g.chunk_stride = (int)(round_up((size_t)C * g.bw * 4, 128) / 4);
        g.bw_magic = 0xffffffffu / (uint32_t)g.bw + 1u;
        if ((size_t)g.nchunks * g.chunk_stride > 65535) tma_ok = false;
    }
    size_t fixed;
    if (tma_ok) {
        g.use_tma = 1;
        fixed = 2 * (size_t)g.nchunks * g.chunk_stride * 4 + list_smem_bytes(W) + 128;
    } else {
        g.use_tma = 0; g.bw = 0; g.nchunks = 0; g.chunk_stride = 0; g.bw_magic = 0;
        fixed = 2 * (size_t)C * Wp * 4 + list_smem_bytes(W) + 128;
    }
    const size_t want = 2 * Wp * Cp * 4;                  // every column of both views listed
    size_t budget = kMaxSmem;
    const size_t half = (size_t)(228 * 1024) / 2 - 1024 - 64;
    if (persistent && fixed + 2 * Wp * 4 + 2048 <= half) {
        g.persistent = 1;
        fixed += 2 * Wp * 4;                              // mask staging rows
        budget = half;
    } else {
        for (int n = 3; n >= 1; --n) {
            const size_t b = (size_t)(228 * 1024) / n - 1024 - 64;   // per-CTA share minus the reserved KB
            if (fixed + 2048 <= b || n == 1) { budget = b < kMaxSmem ? b : kMaxSmem; break; }
        }
    }
    size_t rc = budget > fixed ? budget - fixed : 0;
    if (rc > want) rc = want;
    rc &= ~(size_t)15;
    g.rc_cap = (int)(rc / 4);
    g.smem = fixed + rc;
    return g;
}

// Operand buffer of the sector-gather kernel: what the CTA's share of the SM leaves after the lists, at most
// half of the row's columns of both views and 32 KB.
static size_t gather_smem(int C, int W, int nb, int &rc_cap)
{
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t Cp = (size_t)((C + 3) & ~3);
    const size_t lists = list_smem_bytes(W);
    size_t cap = Wp * Cp * 4;
    if (cap > 32 * 1024) cap = 32 * 1024;                 // more would push the SM to its largest carve-out: the gather's
                                                          // read-only loads lose their L1 (measured: C = 32, 353 -> 475 us)
    const size_t share = (size_t)(228 * 1024) / nb - 1280;
    if (lists + cap > share) cap = share > lists + 4096 ? share - lists : 4096;
    cap &= ~(size_t)15;
    rc_cap = (int)(cap / 4);
    return lists + cap;
}

template <int MODE, int NT, int NB>
static int launch_gather(const float *L, const float *R, const float *ml, const float *mr,
                         const float *disp, float *out_a, float *out_b, float *ssim, float *mx,
                         int B, int C, int H, int W, int D, cudaStream_t st)
{
    RowArgs a{};
    const size_t smem = gather_smem(C, W, NB, a.rc_cap);
    auto kern = sparse_row_gather_kernel<MODE, NT, NB>;
    {
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < smem) {
            DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) set_for[dev] = smem;
        }
    }
    a.L = L; a.R = R; a.lmask = ml; a.rmask = mr; a.disp_in = disp;
    a.out_a = out_a; a.out_b = out_b; a.sum_sim = ssim; a.max_cost = mx;
    a.C = C; a.H = H; a.W = W; a.D = D; a.nrows = B * H;
    a.vec_ok = (W % 4 == 0) && aligned16(out_a) && aligned16(ssim) && aligned16(mx) &&
               (MODE != MODE_FUSED || aligned16(out_b));
    DECNET_CUDA(launch_pdl(kern, dim3(a.nrows), dim3(NT), smem, st, a));
    return after_launch("sparse_row_gather_kernel");
}

// Fused SpaMat + SpaVar over the rows of up to kMaxLevels levels, one launch (levels given finest first).
static int launch_gather_multi(int nlev, const float *const *L, const float *const *R, const float *const *ml,
                               const float *const *mr, float *const *out, float *const *var, float *const *ssim,
                               float *const *mx, const int *B, const int *C, const int *H, const int *W, const int *D,
                               cudaStream_t st)
{
    constexpr int NT = 256, NB = 4;
    MultiArgs m{};
    m.nlev = nlev;
    size_t smem = 0;
    int rows = 0;
    for (int i = 0; i < nlev; ++i) {
        RowArgs &a = m.lv[i];
        const size_t need = gather_smem(C[i], W[i], NB, a.rc_cap);
        smem = need > smem ? need : smem;
        a.L = L[i]; a.R = R[i]; a.lmask = ml[i]; a.rmask = mr[i]; a.disp_in = nullptr;
        a.out_a = out[i]; a.out_b = var[i]; a.sum_sim = ssim[i]; a.max_cost = mx[i];
        a.C = C[i]; a.H = H[i]; a.W = W[i]; a.D = D[i] < 0 ? 0 : D[i]; a.nrows = B[i] * H[i];
        a.vec_ok = (W[i] % 4 == 0) && aligned16(out[i]) && aligned16(ssim[i]) && aligned16(mx[i]) && aligned16(var[i]);
        m.first_row[i] = rows;
        rows += a.nrows;
    }
    m.first_row[nlev] = rows;
    auto kern = sparse_row_gather_multi_kernel<MODE_FUSED, NT, NB>;
    {
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < smem) {
            DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) set_for[dev] = smem;
        }
    }
    DECNET_CUDA(launch_pdl(kern, dim3(rows), dim3(NT), smem, st, m));
    return after_launch("sparse_row_gather_multi_kernel");
}

// Shared memory of the pipelined kernel: mask staging + two list sets + two operand buffers in a 1/NB share of the SM.
static size_t stream_smem(int C, int W, int nb, int &rc_cap)
{
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t Cp = (size_t)((C + 3) & ~3);
    const size_t fixed = 2 * Wp * 4 + 2 * list_smem_bytes(W);
    const size_t share = (size_t)(228 * 1024) / nb - 1280;
    size_t cap = Wp * Cp * 4;                              // half of the row's columns of both views
    if (fixed + 2 * cap > share) cap = share > fixed + 2 * 2048 ? (share - fixed) / 2 : 2048;
    cap &= ~(size_t)15;
    rc_cap = (int)(cap / 4);
    return fixed + 2 * cap;
}

template <int MODE, int NB>
static int launch_stream(const float *L, const float *R, const float *ml, const float *mr,
                         const float *disp, float *out_a, float *out_b, float *ssim, float *mx,
                         int B, int C, int H, int W, int D, cudaStream_t st)
{
    RowArgs a{};
    const size_t smem = stream_smem(C, W, NB, a.rc_cap);
    auto kern = sparse_row_stream_kernel<MODE, NB>;
    {
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < smem) {
            DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) set_for[dev] = smem;
        }
    }
    a.L = L; a.R = R; a.lmask = ml; a.rmask = mr; a.disp_in = disp;
    a.out_a = out_a; a.out_b = out_b; a.sum_sim = ssim; a.max_cost = mx;
    a.C = C; a.H = H; a.W = W; a.D = D; a.nrows = B * H;
    a.vec_ok = (W % 4 == 0) && aligned16(out_a) && aligned16(ssim) && aligned16(mx) &&
               (MODE != MODE_FUSED || aligned16(out_b));
    a.mvec_ok = (W % 4 == 0) && aligned16(ml) && aligned16(mr);
    const int grid = std::min(a.nrows, NB * sm_count_cached());
    kern<<<grid, kStreamThreads, smem, st>>>(a);
    return after_launch("sparse_row_stream_kernel");
}

template <int MODE, bool USE_TMA, int NT, int NB>
static int launch_forward(const TileGeom &g, const float *L, const float *R, const float *ml, const float *mr,
                          const float *disp, float *out_a, float *out_b, float *ssim, float *mx,
                          int B, int C, int H, int W, int D, cudaStream_t st)
{
    CUtensorMap tmL, tmR;
    memset(&tmL, 0, sizeof(tmL)); memset(&tmR, 0, sizeof(tmR));
    if (USE_TMA) {
        const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)B * C};
        const uint64_t strides[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
        const uint32_t box[3] = {(uint32_t)g.bw, 1u, (uint32_t)C};
        int rc = encode_tensor_map(&tmL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, L, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        rc = encode_tensor_map(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, R, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    auto kern = sparse_row_kernel<MODE, USE_TMA, NT, NB>;
    {   // once per (instantiation, device, size): the attribute call costs host time on every launch
        static std::mutex mu;
        static size_t set_for[64] = {0};
        int dev = 0;
        DECNET_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || set_for[dev] < g.smem) {
            DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
            if (dev >= 0 && dev < 64) set_for[dev] = g.smem;
        }
    }
    RowArgs a{};
    a.L = L; a.R = R; a.lmask = ml; a.rmask = mr; a.disp_in = disp;
    a.out_a = out_a; a.out_b = out_b; a.sum_sim = ssim; a.max_cost = mx;
    a.C = C; a.H = H; a.W = W; a.D = D; a.nrows = B * H;
    a.vec_ok = (W % 4 == 0) && aligned16(L) && aligned16(R) && aligned16(out_a) && aligned16(ssim) &&
               aligned16(mx) && (MODE != MODE_FUSED || aligned16(out_b));
    a.mvec_ok = (W % 4 == 0) && aligned16(ml) && aligned16(mr);
    a.bw = g.bw; a.nchunks = g.nchunks; a.chunk_stride = g.chunk_stride; a.bw_magic = g.bw_magic;
    a.rc_cap = g.rc_cap;
    a.stage_masks = g.persistent;
    const int grid = g.persistent ? std::min(a.nrows, NB * sm_count_cached()) : a.nrows;
    kern<<<grid, NT, g.smem, st>>>(tmL, tmR, a);
    return after_launch("sparse_row_kernel");
}

static int forward_dispatch(int mode, const float *L, const float *R, const float *ml, const float *mr,
                            const float *disp, float *out_a, float *out_b, float *ssim, float *mx,
                            int B, int C, int H, int W, int D, void *stream)
{
    int st_ = validate_common(L, R, ml, mr, B, C, H, W);
    if (st_) return st_;
    DECNET_REQUIRE(out_a && ssim && mx, "null output pointer");
    DECNET_REQUIRE(mode != MODE_VAR || disp, "null disparity pointer");
    DECNET_REQUIRE(mode != MODE_FUSED || out_b, "null variance output pointer");
    if (D < 0) D = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Default: the sector-gather kernel (any W / alignment; the int indexing of its direct-from-global fallback
    // needs C*H*W < 2^31).  Variant 4 is its software-pipelined persistent form (measured slower, opt-in);
    // forcing a staged load path or variant 1 / 2 selects the staged row kernel.
    if ((g_forced_variant == 0 || g_forced_variant == 3 || g_forced_variant == 4) && g_forced_path == 0 &&
        (long long)C * H * W < (1ll << 31)) {
        const int v = g_forced_variant == 4 ? 4 : 3;
        g_last_path = 3; g_last_variant = v;
#define DECNET_GATHER(M)                                                                                    \
    (v == 3 ? launch_gather<M, 256, 4>(L, R, ml, mr, disp, out_a, out_b, ssim, mx, B, C, H, W, D, st)       \
            : launch_stream<M, 3>(L, R, ml, mr, disp, out_a, out_b, ssim, mx, B, C, H, W, D, st))
        switch (mode) {
            case MODE_MAT: return DECNET_GATHER(MODE_MAT);
            case MODE_VAR: return DECNET_GATHER(MODE_VAR);
            default:       return DECNET_GATHER(MODE_FUSED);
        }
#undef DECNET_GATHER
    }
    const bool persistent = g_forced_variant == 2;
    TileGeom g = plan_tiles(L, R, C, W, g_forced_path != 1, persistent);
    if (g_forced_path == 2 && !g.use_tma) {
        set_error("TMA path forced but shape/alignment not eligible (W=%d)", W);
        return DECNET_ERR_UNSUPPORTED;
    }
    if (g.smem > kMaxSmem) {
        set_error("row slab needs %zu B of shared memory (> %d): C*W=%d too large", g.smem, (int)kMaxSmem, C * W);
        return DECNET_ERR_UNSUPPORTED;
    }
    g_last_path = g.use_tma ? 2 : 1;
    g_last_variant = g.persistent ? 2 : 1;
#define DECNET_FWD2(M, T)                                                                                          \
    (g.persistent ? launch_forward<M, T, 384, 2>(g, L, R, ml, mr, disp, out_a, out_b, ssim, mx, B, C, H, W, D, st)  \
                  : launch_forward<M, T, 256, 3>(g, L, R, ml, mr, disp, out_a, out_b, ssim, mx, B, C, H, W, D, st))
#define DECNET_FWD(M) (g.use_tma ? DECNET_FWD2(M, true) : DECNET_FWD2(M, false))
    switch (mode) {
        case MODE_MAT: return DECNET_FWD(MODE_MAT);
        case MODE_VAR: return DECNET_FWD(MODE_VAR);
        default:       return DECNET_FWD(MODE_FUSED);
    }
#undef DECNET_FWD2
#undef DECNET_FWD
}

template <int VARMODE>
static int backward_dispatch(const float *L, const float *R, const float *ml, const float *mr,
                             const float *disp, const float *outv, const float *ssim, const float *mx,
                             const float *g, float *dL, float *dR, float *ddisp,
                             int B, int C, int H, int W, int D, void *stream)
{
    int st_ = validate_common(L, R, ml, mr, B, C, H, W);
    if (st_) return st_;
    DECNET_REQUIRE(outv && ssim && mx && g && dL && dR, "null pointer in backward");
    DECNET_REQUIRE(!VARMODE || (disp && ddisp), "null disparity pointer in backward");
    if (D < 0) D = 0;
    const size_t smem = 2 * (size_t)C * ((W + 3) & ~3) * 4 + list_smem_bytes(W);
    if (smem > kMaxSmem) {
        set_error("row slab needs %zu B of shared memory (> %d)", smem, (int)kMaxSmem);
        return DECNET_ERR_UNSUPPORTED;
    }
    auto kern = sparse_row_backward_kernel<VARMODE>;
    DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int vec_ok = (W % 4 == 0) && aligned16(L) && aligned16(R);
    kern<<<B * H, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        L, R, ml, mr, disp, outv, ssim, mx, g, dL, dR, ddisp, C, H, W, D, vec_ok);
    return after_launch("sparse_row_backward_kernel");
}

}  // namespace sparse
}  // namespace decnet

using namespace decnet::sparse;

extern "C" {

int decnet_spamat_fwd(const float *L, const float *R, const float *ml, const float *mr,
                      float *out, float *ssim, float *mx, int B, int C, int H, int W, int D, void *stream) {
    return forward_dispatch(MODE_MAT, L, R, ml, mr, nullptr, out, nullptr, ssim, mx, B, C, H, W, D, stream);
}

int decnet_spavar_fwd(const float *L, const float *R, const float *ml, const float *mr, const float *disp,
                      float *var, float *ssim, float *mx, int B, int C, int H, int W, int D, void *stream) {
    return forward_dispatch(MODE_VAR, L, R, ml, mr, disp, var, nullptr, ssim, mx, B, C, H, W, D, stream);
}

int decnet_spamat_spavar_fwd(const float *L, const float *R, const float *ml, const float *mr,
                             float *out, float *var, float *ssim, float *mx,
                             int B, int C, int H, int W, int D, void *stream) {
    return forward_dispatch(MODE_FUSED, L, R, ml, mr, nullptr, out, var, ssim, mx, B, C, H, W, D, stream);
}

int decnet_spamat_spavar_fwd_levels(int nlev, const float *const *L, const float *const *R, const float *const *ml,
                                    const float *const *mr, float *const *out, float *const *var, float *const *ssim,
                                    float *const *mx, const int *B, const int *C, const int *H, const int *W, const int *D,
                                    void *stream) {
    DECNET_REQUIRE(nlev >= 1 && nlev <= kMaxLevels, "1..%d levels, got %d", kMaxLevels, nlev);
    DECNET_REQUIRE(L && R && ml && mr && out && var && ssim && mx && B && C && H && W && D, "null pointer");
    long long rows = 0;
    for (int i = 0; i < nlev; ++i) {
        DECNET_REQUIRE(L[i] && R[i] && ml[i] && mr[i] && out[i] && var[i] && ssim[i] && mx[i], "level %d: null pointer", i);
        DECNET_REQUIRE(B[i] > 0 && C[i] > 0 && H[i] > 0 && W[i] > 0 && W[i] <= 65535, "level %d: bad size", i);
        DECNET_REQUIRE((long long)C[i] * H[i] * W[i] < (1ll << 31), "level %d: C*H*W too large for the gather kernel", i);
        rows += (long long)B[i] * H[i];
    }
    DECNET_REQUIRE(rows < (1ll << 31), "too many rows");
    g_last_path = 3; g_last_variant = 3;
    return launch_gather_multi(nlev, L, R, ml, mr, out, var, ssim, mx, B, C, H, W, D, static_cast<cudaStream_t>(stream));
}

int decnet_spamat_bwd(const float *L, const float *R, const float *ml, const float *mr, const float *out,
                      const float *ssim, const float *mx, const float *g, float *dL, float *dR,
                      int B, int C, int H, int W, int D, void *stream) {
    return backward_dispatch<0>(L, R, ml, mr, nullptr, out, ssim, mx, g, dL, dR, nullptr, B, C, H, W, D, stream);
}

int decnet_spavar_bwd(const float *L, const float *R, const float *ml, const float *mr, const float *disp,
                      const float *var, const float *ssim, const float *mx, const float *g,
                      float *dL, float *dR, float *ddisp, int B, int C, int H, int W, int D, void *stream) {
    return backward_dispatch<1>(L, R, ml, mr, disp, var, ssim, mx, g, dL, dR, ddisp, B, C, H, W, D, stream);
}

int decnet_candidate_signature(const float *ml, const float *mr, int32_t *count, uint64_t *hash,
                               int B, int H, int W, int D, void *stream) {
    DECNET_REQUIRE(ml && mr && count && hash, "null pointer");
    DECNET_REQUIRE(B > 0 && H > 0 && W > 0 && W <= 65535, "bad size B=%d H=%d W=%d", B, H, W);
    if (D < 0) D = 0;
    const size_t smem = list_smem_bytes(W);
    DECNET_CUDA(cudaFuncSetAttribute(candidate_signature_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    candidate_signature_kernel<<<B * H, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        ml, mr, count, reinterpret_cast<unsigned long long *>(hash), W, D);
    return decnet::after_launch("candidate_signature_kernel");
}

int decnet_last_sparse_path(void) { return g_last_path; }
void decnet_set_sparse_path(int path) { g_forced_path = path; }
int decnet_last_sparse_variant(void) { return g_last_variant; }
void decnet_set_sparse_variant(int variant) { g_forced_variant = variant; }

}  // extern "C"
