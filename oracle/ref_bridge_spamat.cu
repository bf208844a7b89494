// oracle/ref_bridge_spamat.cu -- TEST INFRASTRUCTURE.
// Raw-pointer C entry points around the reference's own host launchers
// (sparse_matching_kernel_forward / _backward, SM_kernel.cu:359-429), linked with
// the unmodified reference SM_kernel.cu.  Used on the GPU box as the on-device
// parity oracle and as the "reference kernels recompiled for sm_100a" timing bar.
#include <torch/extension.h>

extern "C" {
void sparse_matching_kernel_forward(at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor,
                                    at::Tensor, at::Tensor, const int);
void sparse_matching_kernel_backward(at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor,
                                     at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor,
                                     const int);
}

static at::Tensor t4(const void *p, int B, int C, int H, int W) {
    at::Tensor t; t.ptr = const_cast<void *>(p); t.ndim = 4;
    t.dims[0] = B; t.dims[1] = C; t.dims[2] = H; t.dims[3] = W; return t;
}
static at::Tensor t3(const void *p, int B, int H, int W) {
    at::Tensor t; t.ptr = const_cast<void *>(p); t.ndim = 3;
    t.dims[0] = B; t.dims[1] = H; t.dims[2] = W; return t;
}

// The reference launches on the legacy default stream and never checks errors;
// the bridge adds the check so a failed launch is visible to the tests.
extern "C" int ref_spamat_forward(const float *L, const float *R, const float *ml, const float *mr,
                                  float *out, float *sum_sim, float *max_cost,
                                  int B, int C, int H, int W, int D) {
    sparse_matching_kernel_forward(t4(L, B, C, H, W), t4(R, B, C, H, W), t3(ml, B, H, W),
                                   t3(mr, B, H, W), t3(out, B, H, W), t3(sum_sim, B, H, W),
                                   t3(max_cost, B, H, W), D);
    return (int)cudaGetLastError();
}

extern "C" int ref_spamat_backward(const float *L, const float *R, const float *ml, const float *mr,
                                   const float *out, const float *sum_sim, const float *max_cost,
                                   const float *g, float *dL, float *dR,
                                   int B, int C, int H, int W, int D) {
    sparse_matching_kernel_backward(t4(L, B, C, H, W), t4(R, B, C, H, W), t3(ml, B, H, W),
                                    t3(mr, B, H, W), t3(out, B, H, W), t3(sum_sim, B, H, W),
                                    t3(max_cost, B, H, W), t3(g, B, H, W), t4(dL, B, C, H, W),
                                    t4(dR, B, C, H, W), D);
    return (int)cudaGetLastError();
}
