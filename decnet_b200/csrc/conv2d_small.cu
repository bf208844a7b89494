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
    const int halo = (taps == 9) ? dil : 0;
    // Interior threads (every tap of every owned pixel inside the image: > 98 % of a 540x972 map) take a
    // path without bounds checks whose 9 tap addresses are one 64-bit add each and whose P loads per tap
    // use immediate offsets (the first version executed 226 M warp instructions for 76 M warp-FFMA).
    const bool interior = (y - halo >= 0) && (y + halo < H) && (x0 - halo >= 0) && (x0 + (P - 1) * kTX + halo < W);
    if (interior && taps == 9) {
        int toff[9];
#pragma unroll
        for (int ty = 0; ty < 3; ++ty)
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) toff[ty * 3 + tx] = (ty - 1) * dil * W + (tx - 1) * dil;
        const float *pc = xb + (size_t)y * W + x0;
        const float *wc = ws;
        for (int ci = 0; ci < Cin; ++ci, pc += plane, wc += 9 * COUTP) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float *pt = pc + toff[t];
                float v[P];
#pragma unroll
                for (int k = 0; k < P; ++k) v[k] = __ldg(pt + k * kTX);
                const float4 *wv = reinterpret_cast<const float4 *>(wc + t * COUTP);
                float w[COUTP];
#pragma unroll
                for (int q = 0; q < COUTP / 4; ++q) {
                    const float4 tt = wv[q];
                    w[4 * q] = tt.x; w[4 * q + 1] = tt.y; w[4 * q + 2] = tt.z; w[4 * q + 3] = tt.w;
                }
#pragma unroll
                for (int k = 0; k < P; ++k)
#pragma unroll
                    for (int co = 0; co < COUT; ++co) acc[k][co] = fmaf(v[k], w[co], acc[k][co]);
            }
        }
    } else {
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

// ---------------------------------------------------------------------------------------------
// Shared-memory tiled variant for the large (fine-level) maps.  A CTA owns a 64 x 16 output tile; the
// input tile (+ dilation halo, zero filled outside the image) of a few channels at a time is staged
// in shared memory with coalesced loads, so the 9 taps are conflict-free LDS with immediate offsets
// (no per-tap bounds checks or address arithmetic) and the weights are broadcast float4 LDS.
// Thread (tx, ty) computes the 4 pixels (tx, ty + 4*r) x COUT channels.
// ---------------------------------------------------------------------------------------------
constexpr int kTW = 64, kTH = 16, kRows = 4;

template <int COUT, int DIL>
__global__ void __launch_bounds__(256)
conv2d_tiled_kernel(const float *__restrict__ x, const float *__restrict__ wpk, const float *__restrict__ bias,
                    const float *__restrict__ addend, float *__restrict__ out,
                    int Cin, int H, int W, int cc /*channels per round*/, int relu)
{
    constexpr int COUTP = (COUT + 3) & ~3;
    constexpr int PW = kTW + 2 * DIL;                     // tile pitch (floats)
    constexpr int PH = kTH + 2 * DIL;
    extern __shared__ __align__(16) float sm[];
    float *ws = sm;                                       // [Cin][9][COUTP]
    float *tile = sm + Cin * 9 * COUTP;                   // [cc][PH][PW]
    const int tid = threadIdx.x;
    for (int i = tid; i < Cin * 9 * COUTP; i += 256) ws[i] = wpk[i];

    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int tx = tid & 63, ty = tid >> 6;
    float acc[kRows][COUT];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[r][co] = __ldg(bias + co);

    const size_t plane = (size_t)H * W;
    const float *xb = x + (size_t)b * Cin * plane;
    for (int c0 = 0; c0 < Cin; c0 += cc) {
        const int nc = min(cc, Cin - c0);
        __syncthreads();                                  // previous round's tile fully consumed (and ws written)
        for (int i = tid; i < nc * PH * PW; i += 256) {
            const int c = i / (PH * PW), rem = i - c * (PH * PW);
            const int yy = rem / PW, xx = rem - yy * PW;
            const int gy = y0 + yy - DIL, gx = x0 + xx - DIL;
            float v = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(xb + (size_t)(c0 + c) * plane + (size_t)gy * W + gx);
            tile[i] = v;
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            const float *tc = tile + c * PH * PW + ty * PW + tx;
            const float *wc = ws + (c0 + c) * 9 * COUTP;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float w[COUTP];
                    const float4 *wv = reinterpret_cast<const float4 *>(wc + (ky * 3 + kx) * COUTP);
#pragma unroll
                    for (int q = 0; q < COUTP / 4; ++q) {
                        const float4 t = wv[q];
                        w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
                    }
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        const float v = tc[(4 * r + ky * DIL) * PW + kx * DIL];
#pragma unroll
                        for (int co = 0; co < COUT; ++co) acc[r][co] = fmaf(v, w[co], acc[r][co]);
                    }
                }
        }
    }
    const int gx = x0 + tx;
    if (gx < W) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int gy = y0 + ty + 4 * r;
            if (gy < H) {
                float *ob = out + (size_t)b * COUT * plane + (size_t)gy * W + gx;
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    float v = acc[r][co];
                    if (relu) v = fmaxf(v, 0.f);
                    if (COUT == 1 && addend) v += addend[(size_t)b * plane + (size_t)gy * W + gx];
                    ob[(size_t)co * plane] = v;
                }
            }
        }
    }
}

template <int COUT, int DIL>
static int launch_tiled(const float *x, const float *wpk, const float *bias, const float *addend, float *out,
                        int B, int Cin, int H, int W, int relu, cudaStream_t st)
{
    constexpr int COUTP = (COUT + 3) & ~3;
    constexpr int PW = kTW + 2 * DIL, PH = kTH + 2 * DIL;
    const size_t wbytes = (size_t)Cin * 9 * COUTP * sizeof(float);
    // channels per round: keep the CTA around 48 KB so that several CTAs share an SM
    int cc = (int)((48 * 1024 - (long long)wbytes) / (long long)(PH * PW * sizeof(float)));
    if (cc < 1) cc = 1;
    if (cc > Cin) cc = Cin;
    const size_t smem = wbytes + (size_t)cc * PH * PW * sizeof(float);
    auto kern = conv2d_tiled_kernel<COUT, DIL>;
    if (smem > 48 * 1024) DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, B);
    kern<<<grid, 256, smem, st>>>(x, wpk, bias, addend, out, Cin, H, W, cc, relu);
    return after_launch("conv2d_tiled_kernel");
}

template <int COUT>
static int launch_tiled_dil(int dil, const float *x, const float *wpk, const float *bias, const float *addend, float *out,
                            int B, int Cin, int H, int W, int relu, cudaStream_t st, bool *handled)
{
    *handled = true;
    switch (dil) {
        case 1: return launch_tiled<COUT, 1>(x, wpk, bias, addend, out, B, Cin, H, W, relu, st);
        case 2: return launch_tiled<COUT, 2>(x, wpk, bias, addend, out, B, Cin, H, W, relu, st);
        case 3: return launch_tiled<COUT, 3>(x, wpk, bias, addend, out, B, Cin, H, W, relu, st);
        case 4: return launch_tiled<COUT, 4>(x, wpk, bias, addend, out, B, Cin, H, W, relu, st);
        case 6: return launch_tiled<COUT, 6>(x, wpk, bias, addend, out, B, Cin, H, W, relu, st);
        case 9: return launch_tiled<COUT, 9>(x, wpk, bias, addend, out, B, Cin, H, W, relu, st);
        default: *handled = false; return 0;
    }
}

// ConvTranspose2d k=3 s=3 (+bias, ReLU): out[b,co,3y+i,3x+j] = relu(bias[co] + sum_ci in[b,ci,y,x] * w[ci,co,i,j])
template <int COUT, int UNROLL>
__global__ void __launch_bounds__(128)
deconv3x3s3_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                   float *__restrict__ out, int Cin, int h, int wd, int relu)
{
    // One thread = one coarse pixel of output row Y = 3*yy + i: its 3 fine pixels x COUT channels (24
    // accumulators) from ONE load of x per input channel; the row's weights [ci][j][co] are broadcast
    // reads from shared memory (6 LDS.128 per 24 FMAs).
    extern __shared__ __align__(16) float ws[];          // [Cin][3 j][COUT] for this block's row phase i
    const int Y = blockIdx.y, yy = Y / 3, i = Y - 3 * yy;
    {
        // thread -> fixed (j, co), input channels ci0, ci0 + nci, ...: no division inside the loop
        constexpr int kPer = 3 * COUT, kCiStep = 128 / kPer > 0 ? 128 / kPer : 1;
        if (kPer <= 128) {
            const int ci0 = threadIdx.x / kPer, r = threadIdx.x - ci0 * kPer, j = r / COUT, co = r - j * COUT;
            if (ci0 < kCiStep)
                for (int ci = ci0; ci < Cin; ci += kCiStep)
                    ws[ci * kPer + r] = __ldg(w + (ci * COUT + co) * 9 + i * 3 + j);   // PyTorch ConvTranspose2d layout [Cin][Cout][3][3]
        } else {
            for (int t = threadIdx.x; t < Cin * kPer; t += 128) {
                const int ci = t / kPer, r = t - ci * kPer, j = r / COUT, co = r - j * COUT;
                ws[t] = __ldg(w + (ci * COUT + co) * 9 + i * 3 + j);
            }
        }
    }
    __syncthreads();
    const int b = blockIdx.z;
    const int xx = blockIdx.x * 128 + threadIdx.x;
    if (xx >= wd) return;
    float acc[3 * COUT];
#pragma unroll
    for (int k = 0; k < 3 * COUT; ++k) acc[k] = __ldg(bias + (k % COUT));
    const size_t cplane = (size_t)h * wd;
    const float *xp = x + (size_t)b * Cin * cplane + (size_t)yy * wd + xx;
    auto step = [&](int ci) {
        const float v = __ldg(xp + (size_t)ci * cplane);
        const float4 *wr = reinterpret_cast<const float4 *>(ws + ci * 3 * COUT);
#pragma unroll
        for (int k4 = 0; k4 < 3 * COUT / 4; ++k4) {
            const float4 wv = wr[k4];
            acc[4 * k4 + 0] = fmaf(v, wv.x, acc[4 * k4 + 0]);
            acc[4 * k4 + 1] = fmaf(v, wv.y, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(v, wv.z, acc[4 * k4 + 2]);
            acc[4 * k4 + 3] = fmaf(v, wv.w, acc[4 * k4 + 3]);
        }
    };
    // UNROLL = 8 for the coarse levels (few threads, long channel loops: bound by the latency of the loads, keep 8 in flight),
    // 4 where the grid fills the machine (fewer registers, more resident blocks)
#pragma unroll UNROLL
    for (int ci = 0; ci < Cin; ++ci) step(ci);
    const int W = 3 * wd;
    const size_t plane = (size_t)(3 * h) * W;
    float *ob = out + (size_t)b * COUT * plane + (size_t)Y * W + 3 * xx;
#pragma unroll
    for (int co = 0; co < COUT; ++co)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float a = acc[j * COUT + co];
            ob[(size_t)co * plane + j] = relu ? fmaxf(a, 0.f) : a;
        }
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

static thread_local int g_conv2d_variant = 0;     // 0 = auto (register/L1 kernel), 2 = shared-memory tiled kernel

extern "C" {

void decnet_conv2d_set_variant(int v) { g_conv2d_variant = v; }

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
    // Shared-memory tiled kernel: opt-in (variant 2).  Measured SLOWER than the register/L1 kernel on B200 in
    // round 1 (219 vs 200 us for 8->8, 726 vs ~350 us for 17->8 dil 3 at 4.2 M pixels: the halo re-loads and the
    // two block barriers per channel round cost more than the per-tap bounds checks they remove).
    if (ksize == 3 && (long long)H * W >= 64 * 64 && Cin * 9 * ((Cout + 3) & ~3) * 4 <= 24 * 1024 && g_conv2d_variant == 2) {
        bool handled = false;
        int rc = 0;
        switch (Cout) {
            case 1:  rc = launch_tiled_dil<1>(dilation, x, w_packed, bias, addend, out, B, Cin, H, W, relu, st, &handled); break;
            case 3:  rc = launch_tiled_dil<3>(dilation, x, w_packed, bias, addend, out, B, Cin, H, W, relu, st, &handled); break;
            case 4:  rc = launch_tiled_dil<4>(dilation, x, w_packed, bias, addend, out, B, Cin, H, W, relu, st, &handled); break;
            case 8:  rc = launch_tiled_dil<8>(dilation, x, w_packed, bias, addend, out, B, Cin, H, W, relu, st, &handled); break;
            case 12: rc = launch_tiled_dil<12>(dilation, x, w_packed, bias, addend, out, B, Cin, H, W, relu, st, &handled); break;
            default: break;
        }
        if (handled) return rc;
    }
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
    DECNET_REQUIRE(Cout == 8 || Cout == 24, "deconv3x3s3 is instantiated for 8 and 24 output channels "
                   "(GenerateSparseMask.deconv.0, feature extractor deconv1 / deconv2)");
    const size_t smem = (size_t)Cin * Cout * 3 * sizeof(float);
    DECNET_REQUIRE(smem <= 200 * 1024, "Cin too large");
    DECNET_REQUIRE(3 * h <= 65535, "too many rows");
    dim3 grid((w_in + 127) / 128, 3 * h, B);
    if (Cout == 8 && Cin >= 48) {
        auto kern = deconv3x3s3_kernel<8, 8>;
        if (smem > 48 * 1024) DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(x, w, bias, out, Cin, h, w_in, relu);
    } else if (Cout == 8) {
        auto kern = deconv3x3s3_kernel<8, 4>;
        if (smem > 48 * 1024) DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(x, w, bias, out, Cin, h, w_in, relu);
    } else {
        auto kern = deconv3x3s3_kernel<24, 4>;
        if (smem > 48 * 1024) DECNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(x, w, bias, out, Cin, h, w_in, relu);
    }
    return after_launch("deconv3x3s3_kernel");
}

}  // extern "C"
