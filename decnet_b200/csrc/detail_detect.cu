// decnet_b200/csrc/detail_detect.cu -- SURVEY.md section 8f rank 3: the image-space lost-detail detector
// `detailDetection` (utils/utils.py:447-534; caller demo.py:161-162), which the reference runs per image on
// the CPU with cv2, as device kernels on NCHW fp32 images.  One pyramid level per call:
//
//   down = resize_linear(GaussianBlur(data, 3x3, sigma 1), 1/3)   == blur evaluated at (3y+1, 3x+1) only
//   up   = GaussianBlur(resize_linear(down, x3), 5x5, sigma 1)
//   r    = sum_c |data - up| ;  mask = (r - min_img) / (max_img - min_img) >= thold ;  next level: data = down
//
// The arithmetic restates cv2's float32 pipeline the way oracle/detail.py does (which reproduces the reference's
// masks bit for bit): separable filters rows-then-columns with BORDER_REFLECT_101, products and sums rounded
// separately (no FMA contraction), INTER_LINEAR coordinates fx = float((dx+0.5)/3 - 0.5) with cv2's edge clamps.
#include "common.cuh"
#include <algorithm>

namespace decnet {
namespace detail {

constexpr int kBlock = 256;
__constant__ float kG3[3] = {0.274068624f, 0.451862752f, 0.274068624f};        // float32(exp(-x^2/2) / sum)
__constant__ float kG5[5] = {0.054488685f, 0.244201347f, 0.402619958f, 0.244201347f, 0.054488685f};

__device__ __forceinline__ int reflect101(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}
__device__ __forceinline__ float mac(float acc, float k, float v) { return __fadd_rn(acc, __fmul_rn(k, v)); }

// down[b,c,y,x] = (blur3 of data)(3y+1, 3x+1): the 3x3 window is always inside the image (H, W multiples of 3)
__global__ void __launch_bounds__(kBlock)
blur3_down3_kernel(const float *__restrict__ data, float *__restrict__ down, int H, int W, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int h = H / 3, w = W / 3;
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const long long bc = i / ((long long)w * h);
    const float *p = data + bc * H * W + (long long)(3 * y) * W + 3 * x;
    float rows[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float t = 0.f;
        t = mac(t, kG3[0], __ldg(p + j * W)); t = mac(t, kG3[1], __ldg(p + j * W + 1)); t = mac(t, kG3[2], __ldg(p + j * W + 2));
        rows[j] = t;
    }
    float o = 0.f;
    o = mac(o, kG3[0], rows[0]); o = mac(o, kG3[1], rows[1]); o = mac(o, kG3[2], rows[2]);
    down[i] = o;
}

__device__ __forceinline__ void lin_coeff(int d, int n_src, int &s0, int &s1, float &f) {
    float fx = (float)(((double)d + 0.5) * (1.0 / 3.0) - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= n_src - 1) { sx = n_src - 1; fx = 0.f; }
    s0 = sx; s1 = min(sx + 1, n_src - 1); f = fx;
}

// U = resize_linear(down, x3): horizontal interpolation of the two source rows, then vertical
__global__ void __launch_bounds__(kBlock)
up3_kernel(const float *__restrict__ down, float *__restrict__ U, int H, int W, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int h = H / 3, w = W / 3;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long bc = i / ((long long)W * H);
    int sx0, sx1, sy0, sy1; float fx, fy;
    lin_coeff(x, w, sx0, sx1, fx);
    lin_coeff(y, h, sy0, sy1, fy);
    const float *p = down + bc * h * w;
    const float ax = 1.f - fx, ay = 1.f - fy;
    const float h0 = __fadd_rn(__fmul_rn(__ldg(p + (long long)sy0 * w + sx0), ax), __fmul_rn(__ldg(p + (long long)sy0 * w + sx1), fx));
    const float h1 = __fadd_rn(__fmul_rn(__ldg(p + (long long)sy1 * w + sx0), ax), __fmul_rn(__ldg(p + (long long)sy1 * w + sx1), fx));
    U[i] = __fadd_rn(__fmul_rn(h0, ay), __fmul_rn(h1, fy));
}

__global__ void __launch_bounds__(kBlock)
blur5_rows_kernel(const float *__restrict__ U, float *__restrict__ T, int W, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % W);
    const float *row = U + (i - x);
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) t = mac(t, kG5[j], __ldg(row + reflect101(x + j - 2, W)));
    T[i] = t;
}

// r[b,y,x] = sum_c |data - (column blur of T)| ; per-image min / max through order-preserving uint atomics (r >= 0)
__global__ void __launch_bounds__(kBlock)
blur5_cols_residual_kernel(const float *__restrict__ T, const float *__restrict__ data, float *__restrict__ r,
                           unsigned int *__restrict__ minmax, int H, int W)
{
    const int b = blockIdx.y;
    const long long HW = (long long)H * W;
    float lmin = INFINITY, lmax = 0.f;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < HW; i += (long long)gridDim.x * kBlock) {
        const int x = (int)(i % W), y = (int)(i / W);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *t = T + ((long long)b * 3 + c) * HW + x;
            float u = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) u = mac(u, kG5[j], __ldg(t + (long long)reflect101(y + j - 2, H) * W));
            s = __fadd_rn(s, fabsf(__fsub_rn(__ldg(data + ((long long)b * 3 + c) * HW + i), u)));
        }
        r[(long long)b * HW + i] = s;
        lmin = fminf(lmin, s); lmax = fmaxf(lmax, s);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[2 * b], __float_as_uint(lmin));
        atomicMax(&minmax[2 * b + 1], __float_as_uint(lmax));
    }
}

__global__ void __launch_bounds__(kBlock)
residual_mask_kernel(const float *__restrict__ r, const unsigned int *__restrict__ minmax, float thold,
                     float *__restrict__ mask, long long HW)
{
    const int b = blockIdx.y;
    const float mn = __uint_as_float(minmax[2 * b]), mx = __uint_as_float(minmax[2 * b + 1]);
    const float span = __fsub_rn(mx, mn);
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < HW; i += (long long)gridDim.x * kBlock) {
        const float t = __fdiv_rn(__fsub_rn(r[(long long)b * HW + i], mn), span);
        mask[(long long)b * HW + i] = t >= thold ? 1.f : 0.f;
    }
}

}  // namespace detail
}  // namespace decnet

using namespace decnet;
using namespace decnet::detail;

extern "C" {

long long decnet_detail_level_scratch_floats(int B, int H, int W)
{
    return (long long)B * H * W * 7 + 2 * (long long)B;       // U, T (3 planes each), r, min/max words
}

int decnet_detail_level(const float *data, float *down, float *mask, float *scratch, float thold, int B, int H, int W,
                        void *stream)
{
    DECNET_REQUIRE(data && down && mask && scratch, "null pointer");
    DECNET_REQUIRE(B > 0 && B <= 65535 && H >= 3 && W >= 3 && H % 3 == 0 && W % 3 == 0,
                   "H=%d, W=%d must be positive multiples of 3 (pad the image to a multiple of 27 first, demo.py:75-81)", H, W);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long HW = (long long)H * W, n_full = 3ll * B * HW, n_down = n_full / 9;
    float *U = scratch, *T = scratch + n_full, *r = scratch + 2 * n_full;
    unsigned int *minmax = reinterpret_cast<unsigned int *>(r + (long long)B * HW);
    DECNET_CUDA(cudaMemsetAsync(minmax, 0, (size_t)B * 2 * sizeof(unsigned int), st));
    DECNET_CUDA(cudaMemset2DAsync(minmax, 2 * sizeof(unsigned int), 0x7f, sizeof(unsigned int), B, st));   // min slots: large
    blur3_down3_kernel<<<(unsigned)((n_down + kBlock - 1) / kBlock), kBlock, 0, st>>>(data, down, H, W, n_down);
    int rc = after_launch("blur3_down3_kernel");
    if (rc) return rc;
    up3_kernel<<<(unsigned)((n_full + kBlock - 1) / kBlock), kBlock, 0, st>>>(down, U, H, W, n_full);
    if ((rc = after_launch("up3_kernel"))) return rc;
    blur5_rows_kernel<<<(unsigned)((n_full + kBlock - 1) / kBlock), kBlock, 0, st>>>(U, T, W, n_full);
    if ((rc = after_launch("blur5_rows_kernel"))) return rc;
    const unsigned gx = (unsigned)std::min<long long>((HW + kBlock - 1) / kBlock, 148 * 8);
    blur5_cols_residual_kernel<<<dim3(gx, B), kBlock, 0, st>>>(T, data, r, minmax, H, W);
    if ((rc = after_launch("blur5_cols_residual_kernel"))) return rc;
    residual_mask_kernel<<<dim3(gx, B), kBlock, 0, st>>>(r, minmax, thold, mask, HW);
    return after_launch("residual_mask_kernel");
}

}  // extern "C"
