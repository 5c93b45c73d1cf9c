// fused3d.cu — fused multi-axis f32 kernel (placeholder until the first GPU parity run).
#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

bool fused3d_supported(const FusedVolume&, const F32Taps[3], bool) { return false; }

cudaError_t launch_fused3d(const FusedVolume&, const F32Taps[3], const F32Taps[3], bool, cudaStream_t)
{
    return cudaErrorNotSupported;
}

}  // namespace sepfilt
