// fused_ws_p2.cu — instantiations of the warp-specialised fused kernel (fused_ws.cuh) for one group of radii
#include "fused_ws.cuh"

namespace sepfilt {
namespace ws {

cudaError_t launch_plain_r5_8(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius)
{
    switch (radius) {
    case 5: return launch_plain<5, true>(v, p, sms, s);
    case 6: return launch_plain<6, true>(v, p, sms, s);
    case 7: return launch_plain<7, true>(v, p, sms, s);
    case 8: return launch_plain<8, true>(v, p, sms, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ws
}  // namespace sepfilt
