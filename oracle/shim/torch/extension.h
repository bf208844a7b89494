// oracle/shim/torch/extension.h -- TEST INFRASTRUCTURE.
// Minimal stand-in for <torch/extension.h>, just enough for the reference's
// SM_kernel.cu / SV_kernel.cu host launchers (they only use at::Tensor::size(),
// ::numel() and ::data_ptr<float>(); SM_kernel.cu:359-387, SV_kernel.cu:329-410)
// so the UNMODIFIED reference kernels compile in seconds without torch headers
// and can be driven with raw device pointers from oracle/ref_bridge_*.cu.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

namespace at {
struct Tensor {
    void *ptr = nullptr;
    int64_t dims[4] = {1, 1, 1, 1};
    int ndim = 0;
    int64_t size(int i) const { return dims[i]; }
    int64_t numel() const {
        int64_t n = 1;
        for (int i = 0; i < ndim; ++i) n *= dims[i];
        return n;
    }
    template <typename T> T *data_ptr() const { return static_cast<T *>(ptr); }
};
}  // namespace at
