// decnet_b200/csrc/conv2d_small.cu -- direct fp32 convolutions for the tiny-channel 2-D stacks around
// the hot path (the Conv2d stacks of a5 / a13 / a14 at the fine levels,
// modules/submodule.py:351-364, 596-600, 677-716).  At 4.2 M pixels with <= 17 input and <= 8 output
// channels cuDNN runs 6x off the HBM roofline (layout transposes, separate bias / ReLU passes); a
// direct kernel that keeps all output channels of a pixel in registers is FMA/L1-bound instead.
//
//   conv2d_small : 3x3 (dilation d, padding d) or 1x1, stride 1, NCHW fp32, folded-BN bias, optional
//                  ReLU, optional single-channel addend (the refinement's `disp + residual`).
//                  Each thread owns P pixels (warp-strided along x, coalesced loads) x COUT channels;
//                  weights [Cin][taps][COUTP] sit in shared memory and are read as broadcast float4.
//   deconv3x3s3  : ConvTranspose2d(k=3, s=3) + bias + ReLU (GenerateSparseMask.deconv.0, :350-351):
//                  stride == kernel, so every output pixel is one Cin-long dot product.
#include "common.cuh"

namespace decnet {
namespace conv2d {

constexpr int kTX = 32, kTY = 8;

template <int COUT, int P>
__global__ void __launch_bounds__(kTX * kTY)
conv2d_small_kernel(const float *__restrict__ x, const float *__restrict__ wpk, const float *__restrict__ bias,
                    const float *__restrict__ addend, float *__restrict__ out,
                    int Cin, int H, int W, int taps, int dil, int relu)
{
    constexpr int COUTP = (COUT + 3) & ~3;
    extern __shared__ __align__(16) float ws[];          // [Cin][taps][COUTP]
    const int nw = Cin * taps * COUTP;
    for (int i = threadIdx.y * kTX + threadIdx.x; i < nw; i += kTX * kTY) ws[i] = wpk[i];
    __syncthreads();

    const int b = blockIdx.z;
    const int y = blockIdx.y * kTY + threadIdx.y;
    const int x0 = blockIdx.x * (kTX * P) + threadIdx.x;
    if (y >= H) return;
    float acc[P][COUT];
#pragma unroll
    for (int k = 0; k < P; ++k)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[k][co] = __ldg(bias + co);

    const size_t plane = (size_t)H * W;
    const float *xb = x + (size_t)b * Cin * plane;
    const int t0 = (taps == 9) ? 0 : 1, t1 = (taps == 9) ? 3 : 2;     // 1x1: centre tap only
    for (int ci = 0; ci < Cin; ++ci) {
        const float *xc = xb + (size_t)ci * plane;
        for (int ty = t0; ty < t1; ++ty) {
            const int yy = y + (ty - 1) * dil;
            const bool row_ok = yy >= 0 && yy < H;
            const float *xr = xc + (size_t)(row_ok ? yy : 0) * W;
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
                if (taps != 9 && tx != 1) continue;
                const int tap = (taps == 9) ? ty * 3 + tx : 0;
                const float4 *wv = reinterpret_cast<const float4 *>(ws + (ci * taps + tap) * COUTP);
                float w[COUTP];
#pragma unroll
                for (int q = 0; q < COUTP / 4; ++q) {
                    const float4 t = wv[q];
                    w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const int xx = x0 + k * kTX + (tx - 1) * dil;
                    const float v = (row_ok && xx >= 0 && xx < W) ? __ldg(xr + xx) : 0.f;
#pragma unroll
                    for (int co = 0; co < COUT; ++co) acc[k][co] = fmaf(v, w[co], acc[k][co]);
                }
            }
        }
    }
    float *ob = out + (size_t)b * COUT * plane + (size_t)y * W;
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int xx = x0 + k * kTX;
        if (xx < W) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float v = acc[k][co];
                if (relu) v = fmaxf(v, 0.f);
                if (COUT == 1 && addend) v += addend[(size_t)b * plane + (size_t)y * W + xx];
                ob[(size_t)co * plane + xx] = v;
            }
        }
    }
}

// ConvTranspose2d k=3 s=3 (+bias, ReLU): out[b,co,3y+i,3x+j] = relu(bias[co] + sum_ci in[b,ci,y,x] * w[ci,co,i,j])
template <int COUT>
__global__ void __launch_bounds__(256)
deconv3x3s3_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                   float *__restrict__ out, int Cin, int h, int wd, int relu)
{
    extern __shared__ __align__(16) float ws[];          // [Cin][COUT][9] (PyTorch's ConvTranspose2d layout)
    for (int i = threadIdx.x; i < Cin * COUT * 9; i += 256) ws[i] = w[i];
    __syncthreads();
    const int b = blockIdx.z;
    const int H = 3 * h, W = 3 * wd;
    const int Y = blockIdx.y;
    const int X = blockIdx.x * 256 + threadIdx.x;
    if (X >= W) return;
    const int yy = Y / 3, i = Y - 3 * yy, xx = X / 3, j = X - 3 * xx;
    const int tap = i * 3 + j;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = __ldg(bias + co);
    const size_t cplane = (size_t)h * wd;
    const float *xp = x + (size_t)b * Cin * cplane + (size_t)yy * wd + xx;
    for (int ci = 0; ci < Cin; ++ci) {
        const float v = __ldg(xp + (size_t)ci * cplane);
        const float *wr = ws + ci * COUT * 9 + tap;
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = fmaf(v, wr[co * 9], acc[co]);
    }
    const size_t plane = (size_t)H * W;
    float *ob = out + (size_t)b * COUT * plane + (size_t)Y * W + X;
#pragma unroll
    for (int co = 0; co < COUT; ++co) ob[(size_t)co * plane] = relu ? fmaxf(acc[co], 0.f) : acc[co];
}

template <int COUT, int P>
static int launch(const float *x, const float *wpk, const float *bias, const float *addend, float *out,
                  int B, int Cin, int H, int W, int taps, int dil, int relu, cudaStream_t st)
{
    constexpr int COUTP = (COUT + 3) & ~3;
    const size_t smem = (size_t)Cin * taps * COUTP * sizeof(float);
    auto kern = conv2d_small_kernel<COUT, P>;
    if (smem > 48 * 1024) DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + kTX * P - 1) / (kTX * P), (H + kTY - 1) / kTY, B);
    kern<<<grid, dim3(kTX, kTY), smem, st>>>(x, wpk, bias, addend, out, Cin, H, W, taps, dil, relu);
    return after_launch("conv2d_small_kernel");
}

}  // namespace conv2d
}  // namespace decnet

using namespace decnet;
using namespace decnet::conv2d;

extern "C" {

int decnet_conv2d_small_supported(int Cin, int Cout, int ksize) {
    if (ksize != 1 && ksize != 3) return 0;
    if (Cin < 1 || Cin > 160) return 0;
    // Cout = 24 is instantiated but left to cuDNN's TF32 tensor-core kernels, which win from ~24 channels up
    switch (Cout) { case 1: case 3: case 4: case 8: case 12: return 1; default: return 0; }
}

int decnet_conv2d_small(const float *x, const float *w_packed, const float *bias, const float *addend, float *out,
                        int B, int Cin, int H, int W, int Cout, int ksize, int dilation, int relu, void *stream)
{
    DECNET_REQUIRE(x && w_packed && bias && out, "null pointer");
    DECNET_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && H > 0 && W > 0 && dilation >= 1, "bad size");
    DECNET_REQUIRE(decnet_conv2d_small_supported(Cin, Cout, ksize), "unsupported conv shape Cin=%d Cout=%d k=%d", Cin, Cout, ksize);
    DECNET_REQUIRE(!addend || Cout == 1, "addend only for single-channel outputs");
    DECNET_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0, "weights must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int taps = ksize * ksize;
    switch (Cout) {
        case 1:  return launch<1, 4>(x, w_packed, bias, addend, out, B, Cin, H, W, taps, dilation, relu, st);
        case 3:  return launch<3, 4>(x, w_packed, bias, addend, out, B, Cin, H, W, taps, dilation, relu, st);
        case 4:  return launch<4, 4>(x, w_packed, bias, addend, out, B, Cin, H, W, taps, dilation, relu, st);
        case 8:  return launch<8, 4>(x, w_packed, bias, addend, out, B, Cin, H, W, taps, dilation, relu, st);
        case 12: return launch<12, 2>(x, w_packed, bias, addend, out, B, Cin, H, W, taps, dilation, relu, st);
        default: return launch<24, 2>(x, w_packed, bias, addend, out, B, Cin, H, W, taps, dilation, relu, st);
    }
}

int decnet_deconv3x3s3(const float *x, const float *w, const float *bias, float *out,
                       int B, int Cin, int h, int w_in, int Cout, int relu, void *stream)
{
    DECNET_REQUIRE(x && w && bias && out, "null pointer");
    DECNET_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && h > 0 && w_in > 0, "bad size");
    DECNET_REQUIRE(Cout == 8, "deconv3x3s3 is instantiated for 8 output channels (GenerateSparseMask.deconv.0)");
    const size_t smem = (size_t)Cin * Cout * 9 * sizeof(float);
    DECNET_REQUIRE(smem <= 200 * 1024, "Cin too large");
    auto kern = deconv3x3s3_kernel<8>;
    if (smem > 48 * 1024) DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((3 * w_in + 255) / 256, 3 * h, B);
    kern<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, w, bias, out, Cin, h, w_in, relu);
    return after_launch("deconv3x3s3_kernel");
}

}  // extern "C"
