// decnet_b200/csrc/sparse_match_tma.cu -- TMA-staged persistent variant of the sparse row kernel.
// (placeholder until the cp.async kernel is validated on hardware: reports "not handled")
#include "common.cuh"
#include "sparse_core.cuh"

namespace decnet {
namespace sparse {

int tma_forward(int, const float *, const float *, const float *, const float *, const float *,
                float *, float *, float *, float *, int, int, int, int, int, cudaStream_t, bool *handled)
{
    *handled = false;
    return 0;
}

}  // namespace sparse
}  // namespace decnet
