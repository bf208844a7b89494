/*
 * oracle/sparse_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32) of DecNet's two native ops, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under decnet_b200/ may call into it.
 *
 * Follows (semantics, not text) the reference CUDA kernels:
 *   forward   modules/SparseMatching/src/SM_kernel.cu:22-60   (get_max_cost)
 *             modules/SparseMatching/src/SM_kernel.cu:76-125  (sparse_matching_forward)
 *             modules/SparseVar/src/SV_kernel.cu:76-124       (sparse_var_forward)
 *   backward  modules/SparseMatching/src/SM_kernel.cu:143-195, 300-355
 *             modules/SparseVar/src/SV_kernel.cu:142-325
 *   output ownership / zero-fill contract: functions/SpaMat.py:25-27,
 *             functions/SpaVar.py:25-27 -- the CALLER zero-fills every output;
 *             these routines write only where the relevant mask is non-zero.
 *
 * Arithmetic notes kept identical to the device code:
 *   - the channel dot product is a sequential chain from c=0 contracted to FMA
 *     (ptxas emits FFMA for `cost += a*b`), restated here with fmaf();
 *   - max_cost starts at 1e-6f (a floor, not -inf);
 *   - accumulators start at 1e-6f and add candidates in ascending disparity;
 *   - the disparity is converted int->float before use.
 * expf() is glibc's here and CUDA's on the device; both are within 2 ulp, far
 * inside the 1e-3 abs tolerance the north star states.
 *
 * Parity pinning: the reference ships no golden vectors for these ops
 * (SURVEY.md section 4).  This restatement is pinned on the GPU box against
 * the UNMODIFIED reference kernels compiled into oracle/_ref/ (see
 * oracle/Makefile, tests/test_ref_cuda_gpu.py).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define EPS6 0.000001f

static inline float row_cost(const float *L, const float *R, int C, size_t step,
                             size_t lpos, size_t rpos)
{
    float cost = 0.f;
    for (int c = 0; c < C; ++c)
        cost = fmaf(L[lpos + (size_t)c * step], R[rpos + (size_t)c * step], cost);
    return cost;
}

/* ---- forward: max pass (SM_kernel.cu:22-60, identical copy in SV_kernel.cu) ---- */
static float pixel_max_cost(const float *L, const float *R, const float *mr_row,
                            int C, size_t step, size_t base3d, int w, int D)
{
    int nd = (w - D + 1 >= 0) ? D : (w + 1);
    float mx = EPS6;
    for (int d = 0; d < nd; ++d) {
        if (mr_row[w - d] == 0.f) continue;
        float cost = row_cost(L, R, C, step, base3d, base3d - (size_t)d);
        if (mx < cost) mx = cost;
    }
    return mx;
}

void oracle_spamat_forward(const float *L, const float *R, const float *ml,
                           const float *mr, float *out, float *sum_sim,
                           float *max_cost, int B, int C, int H, int W, int D)
{
    const size_t step = (size_t)H * W;
    const long rows = (long)B * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (long r = 0; r < rows; ++r) {
        const int b = (int)(r / H), h = (int)(r % H);
        const size_t m0 = (size_t)r * W;
        const size_t f0 = (size_t)b * C * step + (size_t)h * W;
        for (int w = 0; w < W; ++w) {
            if (ml[m0 + w] == 0.f) continue;
            const size_t base3d = f0 + w;
            float mx = pixel_max_cost(L, R, mr + m0, C, step, base3d, w, D);
            max_cost[m0 + w] = mx;
            int nd = (w - D + 1 >= 0) ? D : (w + 1);
            float ssim = EPS6, sdisp = EPS6;
            for (int d = 0; d < nd; ++d) {
                if (mr[m0 + w - d] == 0.f) continue;
                float cost = row_cost(L, R, C, step, base3d, base3d - (size_t)d);
                float e = expf(cost - mx);
                sdisp += e * (float)d;
                ssim += e;
            }
            sum_sim[m0 + w] = ssim;
            out[m0 + w] = sdisp / ssim;
        }
    }
}

void oracle_spavar_forward(const float *L, const float *R, const float *ml,
                           const float *mr, const float *disp, float *var,
                           float *sum_sim, float *max_cost, int B, int C, int H,
                           int W, int D)
{
    const size_t step = (size_t)H * W;
    const long rows = (long)B * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (long r = 0; r < rows; ++r) {
        const int b = (int)(r / H), h = (int)(r % H);
        const size_t m0 = (size_t)r * W;
        const size_t f0 = (size_t)b * C * step + (size_t)h * W;
        for (int w = 0; w < W; ++w) {
            if (ml[m0 + w] == 0.f) continue;
            const size_t base3d = f0 + w;
            float mx = pixel_max_cost(L, R, mr + m0, C, step, base3d, w, D);
            max_cost[m0 + w] = mx;
            int nd = (w - D + 1 >= 0) ? D : (w + 1);
            float ssim = EPS6, sacc = EPS6;
            const float mu = disp[m0 + w];
            for (int d = 0; d < nd; ++d) {
                if (mr[m0 + w - d] == 0.f) continue;
                float cost = row_cost(L, R, C, step, base3d, base3d - (size_t)d);
                float e = expf(cost - mx);
                float dd = (float)d - mu;
                sacc += e * dd * dd;
                ssim += e;
            }
            sum_sim[m0 + w] = ssim;
            var[m0 + w] = sacc / ssim;
        }
    }
}

/*
 * Backward.  The reference launches one thread per (b,c,h,w) and each thread
 * recomputes the same channel dot product; here exp(cost-max) is computed once
 * per (pixel, candidate) and applied to every channel, which yields the same
 * values because every channel thread of the reference sees the same `cost`.
 *   mode 0 (SpaMat, SM_kernel.cu:143-195,300-355): q(w,d) = d - out[w]
 *   mode 1 (SpaVar, SV_kernel.cu:142-271):         q(w,d) = (d-disp[w])^2 - var[w]
 * dL[c,w] = g[w] * (sum_d e_d * R[c,w-d] * q) / sum_sim[w]          (ml[w] != 0)
 * dR[c,w] = sum_d g[w+d] * e * L[c,w+d] * q(w+d,d) / sum_sim[w+d]   (mr[w] != 0, ml[w+d] != 0)
 */
static void backward_common(int mode, const float *L, const float *R,
                            const float *ml, const float *mr, const float *disp,
                            const float *outv, const float *sum_sim,
                            const float *max_cost, const float *g, float *dL,
                            float *dR, float *ddisp, int B, int C, int H, int W,
                            int D)
{
    const size_t step = (size_t)H * W;
    const long rows = (long)B * H;
#pragma omp parallel
    {
    float *ebuf = (float *)malloc(sizeof(float) * (size_t)(D > 0 ? D : 1));
#pragma omp for schedule(dynamic, 2)
    for (long r = 0; r < rows; ++r) {
        const int b = (int)(r / H), h = (int)(r % H);
        const size_t m0 = (size_t)r * W;
        const size_t f0 = (size_t)b * C * step + (size_t)h * W;
        /* ---- ref (left) gradient, and SpaVar's disparity gradient ---- */
        for (int w = 0; w < W; ++w) {
            if (ml[m0 + w] == 0.f) continue;
            const size_t base3d = f0 + w;
            const float mx = max_cost[m0 + w];
            int nd = (w - D + 1 >= 0) ? D : (w + 1);
            for (int d = 0; d < nd; ++d) {
                if (mr[m0 + w - d] == 0.f) continue;
                float cost = row_cost(L, R, C, step, base3d, base3d - (size_t)d);
                ebuf[d] = expf(cost - mx);
            }
            for (int c = 0; c < C; ++c) {
                float acc = 0.f;
                for (int d = 0; d < nd; ++d) {
                    if (mr[m0 + w - d] == 0.f) continue;
                    float q;
                    if (mode == 0) {
                        q = (float)d - outv[m0 + w];
                    } else {
                        float dd = (float)d - disp[m0 + w];
                        q = dd * dd - outv[m0 + w];
                    }
                    acc += ebuf[d] * R[base3d + (size_t)c * step - (size_t)d] * q;
                }
                dL[base3d + (size_t)c * step] = g[m0 + w] * acc / sum_sim[m0 + w];
            }
            if (mode == 1 && ddisp) {
                float acc = 0.f;
                for (int d = 0; d < nd; ++d) {
                    if (mr[m0 + w - d] == 0.f) continue;
                    acc += ebuf[d] * ((float)d - disp[m0 + w]);
                }
                /* SV_kernel.cu:324 */
                ddisp[m0 + w] = -2 * g[m0 + w] * acc / sum_sim[m0 + w];
            }
        }
        /* ---- tar (right) gradient: gather over w+d ---- */
        for (int w = 0; w < W; ++w) {
            if (mr[m0 + w] == 0.f) continue;
            const size_t base3d = f0 + w;
            int nd = (w + D <= W) ? D : (W - w);
            for (int d = 0; d < nd; ++d) {
                const size_t p = m0 + w + d;
                if (ml[p] == 0.f) continue;
                float cost = row_cost(L, R, C, step, base3d + (size_t)d, base3d);
                ebuf[d] = expf(cost - max_cost[p]);
            }
            for (int c = 0; c < C; ++c) {
                float acc = 0.f;
                for (int d = 0; d < nd; ++d) {
                    const size_t p = m0 + w + d;
                    if (ml[p] == 0.f) continue;
                    float q;
                    if (mode == 0) {
                        q = (float)d - outv[p];
                    } else {
                        float dd = (float)d - disp[p];
                        q = dd * dd - outv[p];
                    }
                    acc += g[p] * ebuf[d] * L[base3d + (size_t)c * step + (size_t)d] * q / sum_sim[p];
                }
                dR[base3d + (size_t)c * step] = acc;
            }
        }
    }
    free(ebuf);
    }
}

void oracle_spamat_backward(const float *L, const float *R, const float *ml,
                            const float *mr, const float *out,
                            const float *sum_sim, const float *max_cost,
                            const float *g, float *dL, float *dR, int B, int C,
                            int H, int W, int D)
{
    backward_common(0, L, R, ml, mr, NULL, out, sum_sim, max_cost, g, dL, dR,
                    NULL, B, C, H, W, D);
}

void oracle_spavar_backward(const float *L, const float *R, const float *ml,
                            const float *mr, const float *disp, const float *var,
                            const float *sum_sim, const float *max_cost,
                            const float *g, float *dL, float *dR, float *ddisp,
                            int B, int C, int H, int W, int D)
{
    backward_common(1, L, R, ml, mr, disp, var, sum_sim, max_cost, g, dL, dR,
                    ddisp, B, C, H, W, D);
}

/*
 * Candidate index sets (bit-exact gate).  For every (b,h,w) with ml != 0 emits
 * the number of valid candidates and an order-independent 64-bit hash of the
 * set {d : 0 <= d < min(D, w+1), mr[w-d] != 0}; unmasked pixels get 0 / 0.
 * The CUDA path exposes the same quantity through decnet_candidate_signature.
 */
void oracle_candidate_signature(const float *ml, const float *mr, int32_t *count,
                                uint64_t *hash, int B, int H, int W, int D)
{
    const long rows = (long)B * H;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
        const size_t m0 = (size_t)r * W;
        for (int w = 0; w < W; ++w) {
            int32_t n = 0;
            uint64_t hsh = 0;
            if (ml[m0 + w] != 0.f) {
                int nd = (w - D + 1 >= 0) ? D : (w + 1);
                for (int d = 0; d < nd; ++d) {
                    if (mr[m0 + w - d] == 0.f) continue;
                    ++n;
                    uint64_t x = (uint64_t)(d + 1) * 0x9E3779B97F4A7C15ull;
                    x ^= x >> 29;
                    x *= 0xBF58476D1CE4E5B9ull;
                    x ^= x >> 32;
                    hsh += x;
                }
            }
            count[m0 + w] = n;
            hash[m0 + w] = hsh;
        }
    }
}
