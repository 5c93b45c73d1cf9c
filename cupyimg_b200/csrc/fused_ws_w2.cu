// fused_ws_w2.cu — instantiations of the warp-specialised fused kernel (fused_ws.cuh) for one group of radii
#include "fused_ws.cuh"

namespace sepfilt {
namespace ws {

cudaError_t launch_wide_r13_16(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius)
{
    switch (radius) {
    case 13: return launch_wide<13>(v, p, sms, s);
    case 14: return launch_wide<14>(v, p, sms, s);
    case 15: return launch_wide<15>(v, p, sms, s);
    case 16: return launch_wide<16>(v, p, sms, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ws
}  // namespace sepfilt
