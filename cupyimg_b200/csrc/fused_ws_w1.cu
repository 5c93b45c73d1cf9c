// fused_ws_w1.cu — instantiations of the warp-specialised fused kernel (fused_ws.cuh) for one group of radii
#include "fused_ws.cuh"

namespace sepfilt {
namespace ws {

cudaError_t launch_wide_r9_12(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius)
{
    switch (radius) {
    case 9: return launch_wide<9>(v, p, sms, s);
    case 10: return launch_wide<10>(v, p, sms, s);
    case 11: return launch_wide<11>(v, p, sms, s);
    case 12: return launch_wide<12>(v, p, sms, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ws
}  // namespace sepfilt
