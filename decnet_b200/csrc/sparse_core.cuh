// decnet_b200/csrc/sparse_core.cuh -- row-level building blocks shared by the cp.async and
// TMA sparse-matching kernels: shared-memory carve-up, ballot/popc mask compaction,
// candidate-range lookup, the per-pixel cost / softmax-regression / variance evaluation,
// and coalesced row stores.  Semantics: SURVEY.md appendix A (reference
// SM_kernel.cu:22-125, SV_kernel.cu:76-124).
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace decnet {
namespace sparse {

constexpr int kThreads = 256;          // threads that evaluate a row
constexpr int KU = 4;                  // candidates per lane per round (held in registers)
constexpr size_t kMaxSmem = 227 * 1024;
constexpr float kEps6 = 0.000001f;     // SM_kernel.cu:45,104

enum { MODE_MAT = 0, MODE_VAR = 1, MODE_FUSED = 2 };

struct RowSmem {
    float *Ls, *Rs;                    // [C][Wp] slabs (cp.async layout)
    float *o_a, *o_b, *o_ssim, *o_max; // output rows
    uint32_t *rlist, *llist;           // compacted columns: (smem_offset << 16) | column
    uint32_t *rbits, *lbits;           // mask bits per 32-column chunk (+1 sentinel)
    int *roff, *loff;                  // exclusive chunk offsets (+1 total)
    int *counts;                       // [0]=nL [1]=nR
};

__host__ __device__ inline size_t list_smem_bytes(int W) {
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t nch = (size_t)((W + 31) / 32) + 1;
    return 4 * Wp * 4        // 4 output rows
           + 2 * Wp * 4      // rlist, llist
           + 4 * ((nch + 3) & ~(size_t)3) * 4   // rbits lbits roff loff
           + 16;             // counts
}
__host__ __device__ inline size_t row_smem_bytes(int C, int W) {
    const size_t Wp = (size_t)((W + 3) & ~3);
    return 2 * (size_t)C * Wp * 4 + list_smem_bytes(W);
}

// lists / outputs carved from `p` (16-byte aligned)
__device__ inline unsigned char *carve_lists(RowSmem &s, unsigned char *p, int W) {
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t nch = (((size_t)((W + 31) / 32) + 1) + 3) & ~(size_t)3;
    s.o_a = reinterpret_cast<float *>(p);    p += Wp * 4;
    s.o_b = reinterpret_cast<float *>(p);    p += Wp * 4;
    s.o_ssim = reinterpret_cast<float *>(p); p += Wp * 4;
    s.o_max = reinterpret_cast<float *>(p);  p += Wp * 4;
    s.rlist = reinterpret_cast<uint32_t *>(p); p += Wp * 4;
    s.llist = reinterpret_cast<uint32_t *>(p); p += Wp * 4;
    s.rbits = reinterpret_cast<uint32_t *>(p); p += nch * 4;
    s.lbits = reinterpret_cast<uint32_t *>(p); p += nch * 4;
    s.roff = reinterpret_cast<int *>(p);       p += nch * 4;
    s.loff = reinterpret_cast<int *>(p);       p += nch * 4;
    s.counts = reinterpret_cast<int *>(p);     p += 16;
    return p;
}

__device__ inline RowSmem carve_row_smem(unsigned char *p, int C, int W) {
    RowSmem s;
    const size_t Wp = (size_t)((W + 3) & ~3);
    s.Ls = reinterpret_cast<float *>(p); p += (size_t)C * Wp * 4;
    s.Rs = reinterpret_cast<float *>(p); p += (size_t)C * Wp * 4;
    carve_lists(s, p, W);
    return s;
}

// number of valid right columns strictly below x, x in [0, W]
__device__ __forceinline__ int row_prefix(const RowSmem &s, int x) {
    const int k = x >> 5;
    return s.roff[k] + __popc(s.rbits[k] & ((1u << (x & 31)) - 1u));
}
// number of masked left columns strictly below x
__device__ __forceinline__ int left_prefix(const RowSmem &s, int x) {
    const int k = x >> 5;
    return s.loff[k] + __popc(s.lbits[k] & ((1u << (x & 31)) - 1u));
}

// Warp-ballot / popc / prefix-sum compaction of both masks of one row.
//   tile_bw == 0 : slab layout [C][Wp]            -> smem offset of column w is w
//   tile_bw  > 0 : TMA box layout [chunk][C][bw]  -> (w / bw) * chunk_stride + w % bw
// Mask test is `!= 0` on fp32 exactly as the reference's `== 0` early-outs
// (SM_kernel.cu:33,49): -0.0 is unmasked, NaN is masked.
__device__ inline void compact_row_masks(RowSmem &s, const float *lmask_row, const float *rmask_row,
                                         int W, int tid, int nthreads,
                                         int tile_bw = 0, int chunk_stride = 0, int bar_id = 0)
{
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int nch = (W + 31) >> 5;
    auto sync = [&]() {
        if (bar_id == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    };
    for (int k = warp; k < nch; k += nwarps) {
        const int w = (k << 5) + lane;
        const bool lv = (w < W) && (lmask_row[w] != 0.f);
        const bool rv = (w < W) && (rmask_row[w] != 0.f);
        const uint32_t lb = __ballot_sync(0xffffffffu, lv);
        const uint32_t rb = __ballot_sync(0xffffffffu, rv);
        if (lane == 0) { s.lbits[k] = lb; s.rbits[k] = rb; }
    }
    if (tid == 0) { s.lbits[nch] = 0u; s.rbits[nch] = 0u; }
    sync();
    if (warp == 0) {
        int carryL = 0, carryR = 0;
        for (int base = 0; base < nch; base += 32) {
            const int k = base + lane;
            const int cl = (k < nch) ? __popc(s.lbits[k]) : 0;
            const int cr = (k < nch) ? __popc(s.rbits[k]) : 0;
            int il = cl, ir = cr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tl = __shfl_up_sync(0xffffffffu, il, o);
                const int tr = __shfl_up_sync(0xffffffffu, ir, o);
                if (lane >= o) { il += tl; ir += tr; }
            }
            if (k < nch) { s.loff[k] = carryL + il - cl; s.roff[k] = carryR + ir - cr; }
            carryL += __shfl_sync(0xffffffffu, il, 31);
            carryR += __shfl_sync(0xffffffffu, ir, 31);
        }
        if (lane == 0) { s.loff[nch] = carryL; s.roff[nch] = carryR; s.counts[0] = carryL; s.counts[1] = carryR; }
    }
    sync();
    for (int k = warp; k < nch; k += nwarps) {
        const int w = (k << 5) + lane;
        const uint32_t lb = s.lbits[k], rb = s.rbits[k];
        const uint32_t below = (1u << lane) - 1u;
        uint32_t off = (uint32_t)w;
        if (tile_bw > 0) { const int ch = w / tile_bw; off = (uint32_t)(ch * chunk_stride + (w - ch * tile_bw)); }
        const uint32_t packed = (off << 16) | (uint32_t)w;
        if ((lb >> lane) & 1u) s.llist[s.loff[k] + __popc(lb & below)] = packed;
        if ((rb >> lane) & 1u) s.rlist[s.roff[k] + __popc(rb & below)] = packed;
    }
    sync();
}

// lanes per masked pixel, from the expected number of candidates per pixel of this row
__device__ __forceinline__ int pick_group_size(int nR, int W, int D) {
    const float avg = 1.25f * (float)nR * (float)min(D, W) / (float)W;
    int G = 1;
    while (G < 32 && (float)(G * KU) < avg) G <<= 1;
    return G;
}

// Evaluate every masked pixel of the row.  `tid` in [0, nthreads), nthreads % 32 == 0;
// all threads of every participating warp must call (warp shuffles inside).
//   Ls/Rs  : staged slabs, element (c, col) at  c*cs + offset(col)
//   MODE_MAT   o_a = out
//   MODE_VAR   o_a = var around disp_row[w]
//   MODE_FUSED o_a = out, o_b = var around out
template <int MODE>
__device__ inline void process_row(RowSmem &s, const float *__restrict__ Ls, const float *__restrict__ Rs,
                                   int cs, int C, int W, int D, const float *disp_row,
                                   int G, int tid, int nthreads)
{
    const int t = tid & (G - 1);
    const int gid = tid / G, nG = nthreads / G;
    const int nL = s.counts[0];
    const int slots = G * KU;

    for (int i0 = 0; i0 < nL; i0 += nG) {
        const int i = i0 + gid;
        const bool act = i < nL;
        int w = 0, lw = 0, lo = 0, hi = 0;
        if (act) {
            const uint32_t le = s.llist[i];
            w = (int)(le & 0xffffu); lw = (int)(le >> 16);
            lo = row_prefix(s, max(0, w - D + 1));
            hi = row_prefix(s, w + 1);
        }
        const int cnt = max(hi - lo, 0);
        const int nrounds = (cnt + slots - 1) / slots;

        // costs of one round: slot k of lane t is list entry jbase + k*G + t
        auto round_costs = [&](int jbase, float (&cost)[KU], float (&df)[KU], bool (&val)[KU]) {
            int ro[KU];
            const int kcnt = min(KU, (hi - jbase + G - 1) / G);   // group-uniform
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                const int j = jbase + k * G + t;
                val[k] = j < hi;
                const uint32_t e = val[k] ? s.rlist[j] : 0u;
                ro[k] = (int)(e >> 16);
                df[k] = (float)(w - (int)(e & 0xffffu));
                cost[k] = 0.f;
            }
            const float *lp = Ls + lw;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                const float l = lp[c * cs];
#pragma unroll
                for (int k = 0; k < KU; ++k)
                    if (k < kcnt) cost[k] = fmaf(l, Rs[c * cs + ro[k]], cost[k]);
            }
        };

        // ---- pass 1: maximum cost (floor 1e-6, SM_kernel.cu:45-59)
        float cost0[KU], df0[KU]; bool v0[KU];
#pragma unroll
        for (int k = 0; k < KU; ++k) { cost0[k] = 0.f; df0[k] = 0.f; v0[k] = false; }
        float mx = -CUDART_INF_F;
        if (nrounds > 0) {
            round_costs(lo, cost0, df0, v0);
#pragma unroll
            for (int k = 0; k < KU; ++k) if (v0[k]) mx = fmaxf(mx, cost0[k]);
        }
        for (int r = 1; r < nrounds; ++r) {
            float cr[KU], dr[KU]; bool vr[KU];
            round_costs(lo + r * slots, cr, dr, vr);
#pragma unroll
            for (int k = 0; k < KU; ++k) if (vr[k]) mx = fmaxf(mx, cr[k]);
        }
        mx = fmaxf(group_max(mx, G), kEps6);

        // ---- pass 2: softmax sums
        const float mu_in = (MODE == MODE_VAR && act) ? disp_row[w] : 0.f;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        float e0[KU];
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            e0[k] = v0[k] ? expf(cost0[k] - mx) : 0.f;
            s0 += e0[k];
            if (MODE == MODE_VAR) { const float dd = df0[k] - mu_in; s2 += e0[k] * dd * dd; }
            else s1 += e0[k] * df0[k];
        }
        for (int r = 1; r < nrounds; ++r) {
            float cr[KU], dr[KU]; bool vr[KU];
            round_costs(lo + r * slots, cr, dr, vr);
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                const float e = vr[k] ? expf(cr[k] - mx) : 0.f;
                s0 += e;
                if (MODE == MODE_VAR) { const float dd = dr[k] - mu_in; s2 += e * dd * dd; }
                else s1 += e * dr[k];
            }
        }
        s0 = group_sum(s0, G);
        const float ssim = kEps6 + s0;
        float outv = 0.f;
        if (MODE != MODE_VAR) { s1 = group_sum(s1, G); outv = (kEps6 + s1) / ssim; }

        // ---- pass 3 (fused): variance around the FINAL mean, exp() of round 0 reused
        if (MODE == MODE_FUSED) {
#pragma unroll
            for (int k = 0; k < KU; ++k) { const float dd = df0[k] - outv; s2 += e0[k] * dd * dd; }
            for (int r = 1; r < nrounds; ++r) {
                float cr[KU], dr[KU]; bool vr[KU];
                round_costs(lo + r * slots, cr, dr, vr);
#pragma unroll
                for (int k = 0; k < KU; ++k) {
                    const float e = vr[k] ? expf(cr[k] - mx) : 0.f;
                    const float dd = dr[k] - outv;
                    s2 += e * dd * dd;
                }
            }
        }
        float varv = 0.f;
        if (MODE != MODE_MAT) { s2 = group_sum(s2, G); varv = (kEps6 + s2) / ssim; }

        if (act && t == 0) {
            s.o_ssim[w] = ssim;
            s.o_max[w] = mx;
            if (MODE == MODE_MAT) s.o_a[w] = outv;
            else if (MODE == MODE_VAR) s.o_a[w] = varv;
            else { s.o_a[w] = outv; s.o_b[w] = varv; }
        }
    }
}

__device__ __forceinline__ void store_row(float *__restrict__ dst, const float *src, int W, int vec_ok,
                                          int tid, int nthreads)
{
    if (vec_ok) {
        const int n4 = W >> 2;
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int i = tid; i < n4; i += nthreads) __stcs(d4 + i, s4[i]);
    } else {
        for (int i = tid; i < W; i += nthreads) dst[i] = src[i];
    }
}

}  // namespace sparse
}  // namespace decnet
