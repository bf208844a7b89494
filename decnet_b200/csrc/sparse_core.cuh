// decnet_b200/csrc/sparse_core.cuh -- row-level building blocks shared by the cp.async and
// TMA sparse-matching kernels: shared-memory carve-up, ballot/popc mask compaction,
// candidate-range lookup and the per-pixel cost / softmax-regression / variance evaluation.
// Semantics: SURVEY.md appendix A (reference SM_kernel.cu:22-125, SV_kernel.cu:76-124).
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace decnet {
namespace sparse {

constexpr int kThreads = 256;          // threads that evaluate a row
constexpr int KU = 4;                  // candidates per lane evaluated together (independent FMA chains)
constexpr size_t kMaxSmem = 227 * 1024;
constexpr float kEps6 = 0.000001f;     // SM_kernel.cu:45,104

enum { MODE_MAT = 0, MODE_VAR = 1, MODE_FUSED = 2 };

struct RowSmem {
    uint32_t *rlist, *llist;           // compacted columns: (smem_offset << 16) | column
    uint32_t *rbits, *lbits;           // mask bits per 32-column chunk (+1 sentinel)
    int *roff, *loff;                  // exclusive chunk offsets (+1 total)
    int *counts;                       // [0]=nL [1]=nR
};

__host__ __device__ inline size_t list_smem_bytes(int W) {
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t nch = (((size_t)((W + 31) / 32) + 1) + 3) & ~(size_t)3;
    return 2 * Wp * 4 + 4 * nch * 4 + 16;
}

// lists carved from `p` (16-byte aligned)
__device__ inline unsigned char *carve_lists(RowSmem &s, unsigned char *p, int W) {
    const size_t Wp = (size_t)((W + 3) & ~3);
    const size_t nch = (((size_t)((W + 31) / 32) + 1) + 3) & ~(size_t)3;
    s.rlist = reinterpret_cast<uint32_t *>(p); p += Wp * 4;
    s.llist = reinterpret_cast<uint32_t *>(p); p += Wp * 4;
    s.rbits = reinterpret_cast<uint32_t *>(p); p += nch * 4;
    s.lbits = reinterpret_cast<uint32_t *>(p); p += nch * 4;
    s.roff = reinterpret_cast<int *>(p);       p += nch * 4;
    s.loff = reinterpret_cast<int *>(p);       p += nch * 4;
    s.counts = reinterpret_cast<int *>(p);     p += 16;
    return p;
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
// Lanes per masked pixel.  Tiny cost model (warp instructions for the row) evaluated for
// G = 1..32: waves(G) * (chunks(G) * per_chunk + per_pixel + per_shuffle_step * log2 G).
// Called by a whole warp (it sits between two barriers of the row, on every thread's critical path): lane lg
// prices G = 2^lg, a shuffle arg-min picks the cheapest (ties: the smaller group, as a sequential scan would).
__device__ __forceinline__ int pick_group_log2(int nL, int nR, int W, int D, int C, int nthreads, int lane) {
    const float avg = (float)nR * (float)min(D, W) / (float)W;
    const float hi = avg + 2.f * sqrtf(avg) + 1.f;        // ~max candidates over the lanes of a warp
    const float per_chunk = (float)(KU * (2 * C + 28));
    const int lg = lane;
    float cost = 3.0e38f;
    if (lg <= 5) {
        const int G = 1 << lg;
        const float waves = ceilf((float)nL * (float)G / (float)nthreads);
        const float chunks = ceilf(hi / (float)(KU * G));
        cost = waves * (chunks * per_chunk + 160.f + 45.f * (float)lg);
    }
    int best = lg;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {                     // lanes 0..7 hold the six candidates
        const float oc = __shfl_down_sync(0xffffffffu, cost, o);
        const int ob = __shfl_down_sync(0xffffffffu, best, o);
        if (oc < cost || (oc == cost && ob < best)) { cost = oc; best = ob; }
    }
    return __shfl_sync(0xffffffffu, best, 0);
}

// Barrier of the threads that work on one row: the whole CTA (N == 0) or a named barrier `id` of N threads.
template <int N> __device__ __forceinline__ void role_sync(int id) {
    if (N == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}

// (see compact_row_masks doc above)  SMEM_SRC: the mask rows were staged in shared memory (plain loads);
// BAR_N / bar_id: the barrier of the `nthreads` threads running it (0 = __syncthreads); PRE: mask chunks per
// thread loaded up front.
template <bool SMEM_SRC = false, int BAR_N = 0, int PRE = 4>
__device__ inline void compact_row_masks(RowSmem &s, const float *__restrict__ lmask_row,
                                         const float *__restrict__ rmask_row,
                                         int W, int tid, int nthreads,
                                         int tile_bw = 0, int chunk_stride = 0, uint32_t bw_magic = 0,
                                         int D = 0, int C = 0, int bar_id = 0)
{
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int nch = (W + 31) >> 5;
    // all mask loads of this thread are issued before the first ballot (one memory round trip)
    float lm[PRE], rm[PRE];
#pragma unroll
    for (int it = 0; it < PRE; ++it) {
        const int w = ((warp + it * nwarps) << 5) + lane;
        const bool in = w < W;
        lm[it] = in ? (SMEM_SRC ? lmask_row[w] : __ldg(lmask_row + w)) : 0.f;
        rm[it] = in ? (SMEM_SRC ? rmask_row[w] : __ldg(rmask_row + w)) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < PRE; ++it) {
        const int k = warp + it * nwarps;
        if (k < nch) {
            const uint32_t lb = __ballot_sync(0xffffffffu, lm[it] != 0.f);
            const uint32_t rb = __ballot_sync(0xffffffffu, rm[it] != 0.f);
            if (lane == 0) { s.lbits[k] = lb; s.rbits[k] = rb; }
        }
    }
    for (int k = warp + PRE * nwarps; k < nch; k += nwarps) {
        const int w = (k << 5) + lane;
        const bool lv = (w < W) && ((SMEM_SRC ? lmask_row[w] : __ldg(lmask_row + w)) != 0.f);
        const bool rv = (w < W) && ((SMEM_SRC ? rmask_row[w] : __ldg(rmask_row + w)) != 0.f);
        const uint32_t lb = __ballot_sync(0xffffffffu, lv);
        const uint32_t rb = __ballot_sync(0xffffffffu, rv);
        if (lane == 0) { s.lbits[k] = lb; s.rbits[k] = rb; }
    }
    if (tid == 0) { s.lbits[nch] = 0u; s.rbits[nch] = 0u; }
    role_sync<BAR_N>(bar_id);
    // Every warp scans the chunk counts itself (32 chunks per step: five shuffles) and takes the offsets of ITS
    // chunks from its own registers, so no barrier separates the scan from the list build and no warp waits
    // for a scanning warp; warp 0 also publishes the offsets that row_prefix() reads later.
    {
        int carryL = 0, carryR = 0;
        const uint32_t below = (1u << lane) - 1u;
        for (int base = 0; base < nch; base += 32) {
            const int kk = base + lane;
            const uint32_t lbv = (kk < nch) ? s.lbits[kk] : 0u;
            const uint32_t rbv = (kk < nch) ? s.rbits[kk] : 0u;
            const int cl = __popc(lbv), cr = __popc(rbv);
            int il = cl, ir = cr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tl = __shfl_up_sync(0xffffffffu, il, o);
                const int tr = __shfl_up_sync(0xffffffffu, ir, o);
                if (lane >= o) { il += tl; ir += tr; }
            }
            const int exL = carryL + il - cl, exR = carryR + ir - cr;
            if (warp == 0 && kk < nch) { s.loff[kk] = exL; s.roff[kk] = exR; }
            const int kend = min(base + 32, nch);
            for (int k = base + warp; k < kend; k += nwarps) {           // warp-uniform bounds
                const int src = k - base;
                const uint32_t lb = __shfl_sync(0xffffffffu, lbv, src), rb = __shfl_sync(0xffffffffu, rbv, src);
                const int oL = __shfl_sync(0xffffffffu, exL, src), oR = __shfl_sync(0xffffffffu, exR, src);
                const int w = (k << 5) + lane;
                uint32_t off = (uint32_t)w;
                if (tile_bw > 0) {
                    const int ch = (int)__umulhi((uint32_t)w, bw_magic);     // w / tile_bw (exact for w < 2^16)
                    off = (uint32_t)(ch * chunk_stride + (w - ch * tile_bw));
                }
                const uint32_t packed = (off << 16) | (uint32_t)w;
                if ((lb >> lane) & 1u) s.llist[oL + __popc(lb & below)] = packed;
                if ((rb >> lane) & 1u) s.rlist[oR + __popc(rb & below)] = packed;
            }
            carryL += __shfl_sync(0xffffffffu, il, 31);
            carryR += __shfl_sync(0xffffffffu, ir, 31);
        }
        if (warp == 0) {
            const int lg = pick_group_log2(carryL, carryR, W, D, C, nthreads, lane);
            if (lane == 0) {
                s.loff[nch] = carryL; s.roff[nch] = carryR; s.counts[0] = carryL; s.counts[1] = carryR;
                s.counts[2] = lg;
            }
        }
    }
    role_sync<BAR_N>(bar_id);
}

template <int G> __device__ __forceinline__ float gmax(float v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int G> __device__ __forceinline__ double gsum(double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Evaluate every masked pixel of the row and store the results straight to global memory
// (the rows were zero-filled earlier by the same CTA, ordered by a __syncthreads()).
//
// Single pass per candidate ("online softmax"): running maximum m (starts at the 1e-6
// floor of SM_kernel.cu:45), sums rescaled by exp(m_old - m_new) when m grows (once per
// chunk of KU candidates).  The three moments S0 = sum e, S1 = sum e*d, S2 = sum e*d^2 are
// kept in fp64 so that
//     var = (1e-6 + S2 - 2*mu*S1 + mu^2*S0) / (1e-6 + S0)
// (identical to sum e*(d-mu)^2 around the FINAL mean) has no fp32 cancellation.
//   Ls     : staged left slab, element (c, col) at c*cs + offset(col)
//   RC     : true  -> Rc holds the VALID right columns transposed to [j][Cp] (j = index in
//                     the compacted list), read with 128-bit loads;
//            false -> Rs is the staged right slab addressed like Ls.
//   MODE_MAT   g_a = out
//   MODE_VAR   g_a = var around disp_row[w]
//   MODE_FUSED g_a = out, g_b = var around out
template <int MODE, int G, int VAR>
__device__ __forceinline__ void process_row_g(const RowSmem &s, const float *__restrict__ Ls,
                                              const float *__restrict__ Rs, int cs, int C, int Cp, int D,
                                              const float *__restrict__ disp_row,
                                              float *__restrict__ g_a, float *__restrict__ g_b,
                                              float *__restrict__ g_ssim, float *__restrict__ g_max,
                                              int tid, int nthreads)
{
    constexpr int LG = (G == 1) ? 0 : (G == 2) ? 1 : (G == 4) ? 2 : (G == 8) ? 3 : (G == 16) ? 4 : 5;
    const int t = tid & (G - 1);
    const int gid = tid >> LG, nG = nthreads >> LG;
    constexpr bool RC = VAR >= 1;          // right columns compacted to Rc[j][Cp]
    constexpr bool LC = VAR >= 2;          // masked left pixels compacted to Lc[i][Cp] (passed in Ls)
    const int nL = s.counts[0];
    const int C4 = C >> 2;

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
        float m = kEps6;
        double S0 = 0.0, S1 = 0.0, S2 = 0.0;
        const float *lp = LC ? Ls + (act ? i : 0) * Cp : Ls + lw;
        const int lcs = LC ? 1 : cs;          // channel stride of the left operand

        for (int jb = lo + t; jb < hi; jb += KU * G) {
            int ro[KU], dk[KU];
            float cost[KU];
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                const int j = jb + k * G;
                const bool ok = j < hi;
                const uint32_t e = ok ? s.rlist[j] : 0u;
                ro[k] = RC ? (ok ? j * Cp : 0) : (int)(e >> 16);
                dk[k] = w - (int)(e & 0xffffu);
                cost[k] = 0.f;
            }
            // the reference's sequential FMA chain over channels, KU independent chains
            if (RC) {
                for (int c4 = 0; c4 < C4; ++c4) {
                    float l0, l1, l2, l3;
                    if (LC) {
                        const float4 lv = *reinterpret_cast<const float4 *>(lp + 4 * c4);
                        l0 = lv.x; l1 = lv.y; l2 = lv.z; l3 = lv.w;
                    } else {
                        l0 = lp[(4 * c4 + 0) * cs]; l1 = lp[(4 * c4 + 1) * cs];
                        l2 = lp[(4 * c4 + 2) * cs]; l3 = lp[(4 * c4 + 3) * cs];
                    }
#pragma unroll
                    for (int k = 0; k < KU; ++k) {
                        const float4 r = *reinterpret_cast<const float4 *>(Rs + ro[k] + 4 * c4);
                        cost[k] = fmaf(l0, r.x, cost[k]);
                        cost[k] = fmaf(l1, r.y, cost[k]);
                        cost[k] = fmaf(l2, r.z, cost[k]);
                        cost[k] = fmaf(l3, r.w, cost[k]);
                    }
                }
                for (int c = 4 * C4; c < C; ++c) {
                    const float l = lp[c * lcs];
#pragma unroll
                    for (int k = 0; k < KU; ++k) cost[k] = fmaf(l, Rs[ro[k] + c], cost[k]);
                }
            } else {
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    const float l = lp[c * cs];
                    const float *rp = Rs + c * cs;
#pragma unroll
                    for (int k = 0; k < KU; ++k) cost[k] = fmaf(l, rp[ro[k]], cost[k]);
                }
            }
            // one rescale per chunk, then one exp per candidate
            float mk = m;
#pragma unroll
            for (int k = 0; k < KU; ++k) if (jb + k * G < hi) mk = fmaxf(mk, cost[k]);
            if (mk > m) {
                const double sc = (double)__expf(m - mk);
                S0 *= sc; S1 *= sc; if (MODE != MODE_MAT) S2 *= sc;
                m = mk;
            }
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                if (jb + k * G < hi) {
                    // ex2.approx(x*log2e): rel. error ~2e-7 (+6e-8*|x|), far inside the 1e-3 abs gate on disparities
                    const double e = (double)__expf(cost[k] - m);
                    const double dd = (double)dk[k];
                    S0 += e;
                    const double ed = e * dd;
                    S1 += ed;
                    if (MODE != MODE_MAT) S2 = fma(ed, dd, S2);
                }
            }
        }
        // merge the G lanes of the group
        if (G > 1) {
            const float mg = gmax<G>(m);
            const double sc = (double)__expf(m - mg);
            S0 = gsum<G>(S0 * sc);
            S1 = gsum<G>(S1 * sc);
            if (MODE != MODE_MAT) S2 = gsum<G>(S2 * sc);
            m = mg;
        }
        if (act && t == 0) {
            // the moments are exact to fp64; the two quotients are fp32 divisions of the rounded sums (1.5 ulp:
            // < 4e-5 px at 216 disparities, against the 1e-3 gate)
            const double den = (double)kEps6 + S0;
            const float ssim = (float)den;
            const float outv = (float)((double)kEps6 + S1) / ssim;
            g_ssim[w] = ssim;
            g_max[w] = m;
            if (MODE == MODE_MAT) {
                g_a[w] = outv;
            } else {
                const double mu = (MODE == MODE_VAR) ? (double)disp_row[w] : (double)outv;
                const double cen = fma(mu, fma(mu, S0, -2.0 * S1), S2);   // sum e*(d-mu)^2
                const float varv = (float)((double)kEps6 + cen) / ssim;
                if (MODE == MODE_VAR) g_a[w] = varv;
                else { g_a[w] = outv; g_b[w] = varv; }
            }
        }
    }
}

template <int MODE, int VAR>
__device__ inline void process_row(const RowSmem &s, const float *Ls, const float *Rs, int cs, int C, int Cp,
                                   int D, const float *disp_row, float *g_a, float *g_b,
                                   float *g_ssim, float *g_max, int tid, int nthreads)
{
#define DECNET_PR(GG) process_row_g<MODE, GG, VAR>(s, Ls, Rs, cs, C, Cp, D, disp_row, g_a, g_b, g_ssim, g_max, tid, nthreads)
    switch (s.counts[2]) {
        case 0: DECNET_PR(1); break;
        case 1: DECNET_PR(2); break;
        case 2: DECNET_PR(4); break;
        case 3: DECNET_PR(8); break;
        case 4: DECNET_PR(16); break;
        default: DECNET_PR(32); break;
    }
#undef DECNET_PR
}

// Transpose the listed (valid right / masked left) columns of a staged slab into Rc[j][Cp] (j = list index).
__device__ inline void gather_columns(const uint32_t *__restrict__ list, int nR, const float *__restrict__ Rs,
                                      int cs, int C, int Cp, float *__restrict__ Rc, int tid, int nthreads)
{
    const int q4 = Cp >> 2;
    const uint32_t magic = 0xffffffffu / (uint32_t)q4 + 1u;      // idx / q4, exact for idx < 2^16 * q4
    const int total = nR * q4;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int j = (q4 == 1) ? idx : (int)__umulhi((uint32_t)idx, magic);
        const int q = idx - j * q4;
        const int off = (int)(list[j] >> 16);
        const int c = 4 * q;
        float4 v;
        v.x = Rs[(c + 0) * cs + off];
        v.y = (c + 1 < C) ? Rs[(c + 1) * cs + off] : 0.f;
        v.z = (c + 2 < C) ? Rs[(c + 2) * cs + off] : 0.f;
        v.w = (c + 3 < C) ? Rs[(c + 3) * cs + off] : 0.f;
        *reinterpret_cast<float4 *>(Rc + j * Cp + c) = v;
    }
}

// coalesced zero fill of the output rows of one image row (n rows of W floats)
__device__ __forceinline__ void zero_rows(float *__restrict__ r0, float *__restrict__ r1, float *__restrict__ r2,
                                          float *__restrict__ r3, int W, int vec_ok, int tid, int nthreads) {
    if (vec_ok) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < (W >> 2); i += nthreads) {
            reinterpret_cast<float4 *>(r0)[i] = z;
            reinterpret_cast<float4 *>(r1)[i] = z;
            reinterpret_cast<float4 *>(r2)[i] = z;
            if (r3) reinterpret_cast<float4 *>(r3)[i] = z;
        }
    } else {
        for (int i = tid; i < W; i += nthreads) { r0[i] = 0.f; r1[i] = 0.f; r2[i] = 0.f; if (r3) r3[i] = 0.f; }
    }
}

}  // namespace sparse
}  // namespace decnet
