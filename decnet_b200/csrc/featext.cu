// decnet_b200/csrc/featext.cu -- data-movement kernels that put the coarse half of the feature extractor
// (FeatExtNetChannelPlus, modules/submodule.py:245-343: the stride-3 convs, the 72 / 216-channel layers, the ASPP and the
// 216 -> 72 transposed conv at 1/9 and 1/27 resolution) on the tensor-core GEMM / conv kernel of conv2d_nhwc_tcgen05.cu:
//
//   im2col3x3            rows [pixel][tap*C + c] of a 3x3 window (any stride / dilation, zero padding = dilation) from a
//                        source with arbitrary element strides (NCHW, flat or zero-bordered channels-last), so that a strided
//                        or dilated conv becomes one GEMM (decnet_gemm_tc_nhwc)
//   deconv3x3s3_shuffle  pixel shuffle behind the GEMM form of ConvTranspose2d(k 3, s 3): [p][tap*Cout + co] -> the interior
//                        of a zero-bordered channels-last tensor at 3x the resolution (channel slice of a wider tensor)
//   nhwc_to_nchw         channels-last (flat or zero-bordered, row stride ld) -> NCHW, tiled through shared memory
//   conv3x3s3_nchw       the first, 8 -> 24 channel stride-3 conv at full resolution as a direct fp32 kernel (NCHW in/out:
//                        too few channels for a GEMM, too many pixels for an im2col buffer)
#include "common.cuh"

namespace decnet {
namespace featext {

constexpr int kBlock = 256;

// ---- im2col, channels-last source (sc == 1): one thread per output element, k fastest -> coalesced both ways
__global__ void __launch_bounds__(kBlock)
im2col3x3_cl_kernel(const float *__restrict__ src, float *__restrict__ out, int C, int H, int W, long long sb, long long sy,
                    long long sx, int stride, int dil, int Ho, int Wo, int Kp, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;      // over P * Kp
    if (i >= n) return;
    const int k = (int)(i % Kp);
    long long p = i / Kp;
    const int xo = (int)(p % Wo); p /= Wo;
    const int yo = (int)(p % Ho);
    const long long b = p / Ho;
    float v = 0.f;
    if (k < 9 * C) {
        const int tap = k / C, c = k - tap * C;
        const int y = yo * stride + (tap / 3 - 1) * dil, x = xo * stride + (tap % 3 - 1) * dil;
        if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(src + b * sb + (long long)y * sy + (long long)x * sx + c);
    }
    out[i] = v;
}

// the same with four channels per thread (C, Kp, strides and base all multiples of 4 floats): 128-bit loads and stores
__global__ void __launch_bounds__(kBlock)
im2col3x3_cl4_kernel(const float *__restrict__ src, float *__restrict__ out, int C, int H, int W, long long sb, long long sy,
                     long long sx, int stride, int dil, int Ho, int Wo, int Kp, long long n4)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;      // over P * Kp / 4
    if (i >= n4) return;
    const int kq = Kp >> 2;
    const int k = (int)(i % kq) * 4;
    long long p = i / kq;
    const int xo = (int)(p % Wo); p /= Wo;
    const int yo = (int)(p % Ho);
    const long long b = p / Ho;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < 9 * C) {
        const int tap = k / C, c = k - tap * C;
        const int y = yo * stride + (tap / 3 - 1) * dil, x = xo * stride + (tap % 3 - 1) * dil;
        if (y >= 0 && y < H && x >= 0 && x < W)
            v = __ldg(reinterpret_cast<const float4 *>(src + b * sb + (long long)y * sy + (long long)x * sx + c));
    }
    reinterpret_cast<float4 *>(out)[i] = v;
}

// ---- im2col, x-contiguous source (NCHW: sx == 1): 32 output pixels of one row x 32 k per block, transposed through smem
__global__ void __launch_bounds__(1024)
im2col3x3_xc_kernel(const float *__restrict__ src, float *__restrict__ out, int C, int H, int W, long long sb, long long sc,
                    long long sy, int stride, int dil, int Ho, int Wo, int Kp)
{
    __shared__ float tile[32][33];
    const int xo0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int yo = blockIdx.z % Ho;
    const long long b = blockIdx.z / Ho;
    {
        const int k = k0 + threadIdx.y, xo = xo0 + threadIdx.x;               // read: pixels fastest
        float v = 0.f;
        if (k < 9 * C && xo < Wo) {
            const int tap = k / C, c = k - tap * C;
            const int y = yo * stride + (tap / 3 - 1) * dil, x = xo * stride + (tap % 3 - 1) * dil;
            if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(src + b * sb + (long long)c * sc + (long long)y * sy + x);
        }
        tile[threadIdx.y][threadIdx.x] = v;
    }
    __syncthreads();
    {
        const int k = k0 + threadIdx.x, xo = xo0 + threadIdx.y;               // write: k fastest
        if (k < Kp && xo < Wo) out[((b * Ho + yo) * Wo + xo) * Kp + k] = tile[threadIdx.x][threadIdx.y];
    }
}

// ---- pixel shuffle of the GEMM-form transposed conv
__global__ void __launch_bounds__(kBlock)
deconv3x3s3_shuffle_kernel(const float *__restrict__ in, float *__restrict__ out, int h, int w, int Cout, int ld_in, int ldc,
                           int c_off, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;      // over B * 3h * 3w * Cout, co fastest
    if (i >= n) return;
    const int co = (int)(i % Cout);
    long long p = i / Cout;
    const int X = (int)(p % (3 * w)); p /= 3 * w;
    const int Y = (int)(p % (3 * h));
    const long long b = p / (3 * h);
    const int y = Y / 3, ky = Y - 3 * y, x = X / 3, kx = X - 3 * x;
    const float v = __ldg(in + ((b * h + y) * w + x) * ld_in + (ky * 3 + kx) * Cout + co);
    out[((b * (3 * h + 2) + Y + 1) * (long long)(3 * w + 2) + X + 1) * ldc + c_off + co] = v;
}

// ---- channels-last -> NCHW: 32 pixels x 32 channels per block through smem (coalesced reads over c, writes over x)
__global__ void __launch_bounds__(1024)
nhwc_to_nchw_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int ld, int h, int w, int pad)
{
    __shared__ float tile[32][33];
    const long long hw = (long long)h * w;
    const long long p0 = (long long)blockIdx.x * 32;                       // pixel index within the image
    const int c0 = blockIdx.y * 32;
    const long long b = blockIdx.z;
    {
        const long long p = p0 + threadIdx.y;
        const int c = c0 + threadIdx.x;
        float v = 0.f;
        if (p < hw && c < C) {
            const int y = (int)(p / w), x = (int)(p - (long long)y * w);
            const long long row = pad ? (b * (h + 2) + y + 1) * (long long)(w + 2) + x + 1 : b * hw + p;
            v = __ldg(in + row * ld + c);
        }
        tile[threadIdx.y][threadIdx.x] = v;
    }
    __syncthreads();
    {
        const long long p = p0 + threadIdx.x;
        const int c = c0 + threadIdx.y;
        if (p < hw && c < C) out[(b * C + c) * hw + p] = tile[threadIdx.x][threadIdx.y];
    }
}

// ---- direct 3x3 stride-3 pad-1 conv + bias + ReLU, NCHW, COUT output channels per thread (one output pixel per thread)
template <int COUT>
__global__ void __launch_bounds__(kBlock)
conv3x3s3_nchw_kernel(const float *__restrict__ x, const float *__restrict__ wpk, const float *__restrict__ bias,
                      float *__restrict__ out, int Cin, int H, int W, int Ho, int Wo, int relu)
{
    extern __shared__ __align__(16) float ws[];                             // [Cin][9][COUT]
    for (int i = threadIdx.x; i < Cin * 9 * COUT; i += kBlock) ws[i] = wpk[i];
    __syncthreads();
    const int xo = blockIdx.x * kBlock + threadIdx.x;
    const int yo = blockIdx.y;
    const long long b = blockIdx.z;
    if (xo >= Wo) return;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
    const long long plane = (long long)H * W;
    const float *xb = x + b * Cin * plane;
    for (int c = 0; c < Cin; ++c) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int y = 3 * yo + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = 3 * xo + kx - 1;
                const float v = (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(xb + c * plane + (long long)y * W + xx) : 0.f;
                const float4 *wv = reinterpret_cast<const float4 *>(ws + (c * 9 + ky * 3 + kx) * COUT);
#pragma unroll
                for (int q = 0; q < COUT / 4; ++q) {
                    const float4 t = wv[q];
                    acc[4 * q] = fmaf(v, t.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, t.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, t.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, t.w, acc[4 * q + 3]);
                }
            }
        }
    }
    const long long oplane = (long long)Ho * Wo;
    float *ob = out + b * COUT * oplane + (long long)yo * Wo + xo;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float v = acc[co] + __ldg(bias + co);
        if (relu) v = fmaxf(v, 0.f);
        ob[co * oplane] = v;
    }
}

}  // namespace featext
}  // namespace decnet

using namespace decnet;
using namespace decnet::featext;

extern "C" {

int decnet_im2col3x3(const float *src, float *out, int B, int C, int H, int W, long long sb, long long sc, long long sy,
                     long long sx, int stride, int dilation, int Ho, int Wo, int Kp, void *stream)
{
    DECNET_REQUIRE(src && out, "null pointer");
    DECNET_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && stride >= 1 && dilation >= 1, "bad size");
    DECNET_REQUIRE(Kp >= 9 * C, "Kp=%d < 9*C=%d", Kp, 9 * C);
    DECNET_REQUIRE(sc == 1 || sx == 1, "the source must be channels-last (sc = 1) or x-contiguous (sx = 1)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (sc == 1 && C % 4 == 0 && Kp % 4 == 0 && sb % 4 == 0 && sy % 4 == 0 && sx % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        const long long n4 = (long long)B * Ho * Wo * (Kp / 4);
        im2col3x3_cl4_kernel<<<(unsigned)((n4 + kBlock - 1) / kBlock), kBlock, 0, st>>>(src, out, C, H, W, sb, sy, sx, stride, dilation,
                                                                                    Ho, Wo, Kp, n4);
        return after_launch("im2col3x3_cl4_kernel");
    }
    if (sc == 1) {
        const long long n = (long long)B * Ho * Wo * Kp;
        im2col3x3_cl_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, st>>>(src, out, C, H, W, sb, sy, sx, stride, dilation,
                                                                                  Ho, Wo, Kp, n);
        return after_launch("im2col3x3_cl_kernel");
    }
    DECNET_REQUIRE((long long)B * Ho <= 65535, "B*Ho too large for grid.z");
    dim3 grid((Wo + 31) / 32, (Kp + 31) / 32, (unsigned)(B * Ho));
    im2col3x3_xc_kernel<<<grid, dim3(32, 32), 0, st>>>(src, out, C, H, W, sb, sc, sy, stride, dilation, Ho, Wo, Kp);
    return after_launch("im2col3x3_xc_kernel");
}

int decnet_deconv3x3s3_shuffle(const float *in, float *out_pad, int B, int h, int w, int Cout, int ld_in, int ldc, int c_off,
                               void *stream)
{
    DECNET_REQUIRE(in && out_pad, "null pointer");
    DECNET_REQUIRE(B > 0 && h > 0 && w > 0 && Cout > 0 && ld_in >= 9 * Cout && ldc >= c_off + Cout && c_off >= 0, "bad size");
    const long long n = (long long)B * 9 * h * w * Cout;
    deconv3x3s3_shuffle_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out_pad, h, w, Cout, ld_in, ldc, c_off, n);
    return after_launch("deconv3x3s3_shuffle_kernel");
}

int decnet_nhwc_to_nchw(const float *in, float *out, int B, int C, int ld, int h, int w, int pad, void *stream)
{
    DECNET_REQUIRE(in && out, "null pointer");
    DECNET_REQUIRE(B > 0 && B <= 65535 && C > 0 && C <= ld && h > 0 && w > 0, "bad size");
    dim3 grid((unsigned)(((long long)h * w + 31) / 32), (C + 31) / 32, B);
    nhwc_to_nchw_kernel<<<grid, dim3(32, 32), 0, static_cast<cudaStream_t>(stream)>>>(in, out, C, ld, h, w, pad ? 1 : 0);
    return after_launch("nhwc_to_nchw_kernel");
}

int decnet_conv3x3s3_nchw(const float *x, const float *w_packed, const float *bias, float *out, int B, int Cin, int H, int W,
                          int Cout, int relu, void *stream)
{
    DECNET_REQUIRE(x && w_packed && bias && out, "null pointer");
    DECNET_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cin <= 48 && H > 0 && W > 0, "bad size");
    DECNET_REQUIRE(Cout == 24, "conv3x3s3_nchw is instantiated for 24 output channels (FeatExtNetChannelPlus.conv1.0), got %d", Cout);
    DECNET_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0, "weights must be 16-byte aligned");
    const int Ho = (H + 2 - 3) / 3 + 1, Wo = (W + 2 - 3) / 3 + 1;
    DECNET_REQUIRE(Ho <= 65535, "too many rows");
    const size_t smem = (size_t)Cin * 9 * 24 * sizeof(float);
    dim3 grid((Wo + kBlock - 1) / kBlock, Ho, B);
    conv3x3s3_nchw_kernel<24><<<grid, kBlock, smem, static_cast<cudaStream_t>(stream)>>>(x, w_packed, bias, out, Cin, H, W, Ho, Wo, relu);
    return after_launch("conv3x3s3_nchw_kernel");
}

}  // extern "C"
