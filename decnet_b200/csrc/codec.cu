// decnet_b200/csrc/codec.cu -- SURVEY.md section 8f rank 4: the data formats either side of the path, on the device.
//   image_prepare_u8 : uint8 RGB HWC image -> top/left zero pad to the network size (demo.py:75-81 `padding`), /255
//                      (the input of detailDetection, demo.py:158-162) and ToTensor + Normalize(mean, std)
//                      (demo.py:82-88 `transform`), NCHW fp32 -- one pass, a quarter of the upload bytes of fp32 images.
//   disp_to_u16      : the KITTI-style 16-bit disparity image demo.py:191-197 writes: clamp(pred * 256, 0, 65535)
//                      truncated to uint16, cropped to the last ori_h rows / ori_w columns (the PNG deflate stays on the host).
//   epe_3px          : modules/loss.py:427-437 `test_loss_func`: EPE and 3-px / 5 % error over 0 < gt < max_disp.
#include "common.cuh"
#include <algorithm>

namespace decnet {
namespace codec {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock)
image_prepare_u8_kernel(const unsigned char *__restrict__ img, float *__restrict__ out01, float *__restrict__ norm,
                        float m0, float m1, float m2, float s0, float s1, float s2, int h, int w, int H, int W, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;      // over B*H*W padded pixels
    if (i >= n) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    const int ry = H - h, rx = W - w;                                      // residual rows / columns come FIRST
    float v[3] = {0.f, 0.f, 0.f};
    if (y >= ry && x >= rx) {
        const unsigned char *p = img + ((b * h + (y - ry)) * (long long)w + (x - rx)) * 3;
        v[0] = __fdiv_rn((float)p[0], 255.f); v[1] = __fdiv_rn((float)p[1], 255.f); v[2] = __fdiv_rn((float)p[2], 255.f);
    }
    const long long plane = (long long)H * W, o = b * 3 * plane + (long long)y * W + x;
    if (out01) { out01[o] = v[0]; out01[o + plane] = v[1]; out01[o + 2 * plane] = v[2]; }
    if (norm) {
        norm[o] = __fdiv_rn(__fsub_rn(v[0], m0), s0);
        norm[o + plane] = __fdiv_rn(__fsub_rn(v[1], m1), s1);
        norm[o + 2 * plane] = __fdiv_rn(__fsub_rn(v[2], m2), s2);
    }
}

__global__ void __launch_bounds__(kBlock)
disp_to_u16_kernel(const float *__restrict__ pred, unsigned short *__restrict__ out, int H, int W, int oh, int ow, long long n)
{
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;      // over B*oh*ow cropped pixels
    if (i >= n) return;
    const int x = (int)(i % ow), y = (int)((i / ow) % oh);
    const long long b = i / ((long long)ow * oh);
    float v = __fmul_rn(pred[(b * H + (H - oh + y)) * (long long)W + (W - ow + x)], 256.f);
    v = v < 0.f ? 0.f : v;                                                 // also sends NaN comparisons' false branch through
    v = v > 65535.f ? 65535.f : v;
    out[i] = (unsigned short)(v == v ? v : 0.f);                           // truncation like numpy's astype('uint16')
}

__global__ void __launch_bounds__(kBlock)
epe_3px_kernel(const float *__restrict__ pred, const float *__restrict__ gt, float max_disp, double *__restrict__ sums, long long n)
{
    double e = 0.0, ok = 0.0, cnt = 0.0;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const float g = gt[i];
        if (g < max_disp && g > 0.f) {
            const float err = fabsf(pred[i] - g);
            e += err; cnt += 1.0;
            ok += (err < 3.f || err < 0.05f * g) ? 1.0 : 0.0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o); ok += __shfl_xor_sync(0xffffffffu, ok, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], e); atomicAdd(&sums[1], ok); atomicAdd(&sums[2], cnt); }
}

}  // namespace codec
}  // namespace decnet

using namespace decnet;
using namespace decnet::codec;

extern "C" {

int decnet_image_prepare_u8(const unsigned char *img_hwc, float *out01, float *out_norm, const float *mean3_host,
                            const float *std3_host, int B, int h, int w, int H, int W, void *stream)
{
    DECNET_REQUIRE(img_hwc && (out01 || out_norm) && mean3_host && std3_host, "null pointer");
    DECNET_REQUIRE(B > 0 && h > 0 && w > 0 && H >= h && W >= w, "padded size %dx%d must cover the image %dx%d", H, W, h, w);
    const long long n = (long long)B * H * W;
    image_prepare_u8_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        img_hwc, out01, out_norm, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2], h, w, H, W, n);
    return after_launch("image_prepare_u8_kernel");
}

int decnet_disp_to_u16(const float *pred, unsigned short *out, int B, int H, int W, int ori_h, int ori_w, void *stream)
{
    DECNET_REQUIRE(pred && out, "null pointer");
    DECNET_REQUIRE(B > 0 && ori_h > 0 && ori_w > 0 && ori_h <= H && ori_w <= W, "crop %dx%d outside the %dx%d map", ori_h, ori_w, H, W);
    const long long n = (long long)B * ori_h * ori_w;
    disp_to_u16_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(pred, out, H, W, ori_h, ori_w, n);
    return after_launch("disp_to_u16_kernel");
}

int decnet_epe_3px(const float *pred, const float *gt, float max_disp, double *sums3, long long n, void *stream)
{
    DECNET_REQUIRE(pred && gt && sums3 && n > 0, "null pointer or empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DECNET_CUDA(cudaMemsetAsync(sums3, 0, 3 * sizeof(double), st));
    epe_3px_kernel<<<(unsigned)std::min<long long>((n + kBlock - 1) / kBlock, 148 * 8), kBlock, 0, st>>>(pred, gt, max_disp, sums3, n);
    return after_launch("epe_3px_kernel");
}

}  // extern "C"
