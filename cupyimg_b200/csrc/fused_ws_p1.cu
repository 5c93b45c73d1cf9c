// fused_ws_p1.cu — instantiations of the warp-specialised fused kernel (fused_ws.cuh) for one group of radii
#include "fused_ws.cuh"

namespace sepfilt {
namespace ws {

cudaError_t launch_plain_r1_4(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius)
{
    switch (radius) {
    case 1: return launch_plain<1, true>(v, p, sms, s);
    case 2: return launch_plain<2, true>(v, p, sms, s);
    case 3: return launch_plain<3, true>(v, p, sms, s);
    case 4: return launch_plain<4, true>(v, p, sms, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ws
}  // namespace sepfilt
