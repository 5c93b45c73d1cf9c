// fused_ws_g.cu — instantiations of the warp-specialised fused kernel (fused_ws.cuh) for one group of radii
#include "fused_ws.cuh"

namespace sepfilt {
namespace ws {

cudaError_t launch_grad_r1_6(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius)
{
    switch (radius) {
    case 1: return launch_grad<1>(v, p, sms, s);
    case 2: return launch_grad<2>(v, p, sms, s);
    case 3: return launch_grad<3>(v, p, sms, s);
    case 4: return launch_grad<4>(v, p, sms, s);
    case 5: return launch_grad<5>(v, p, sms, s);
    case 6: return launch_grad<6>(v, p, sms, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ws
}  // namespace sepfilt
