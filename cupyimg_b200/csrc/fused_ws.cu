// fused_ws.cu — dispatch of the warp-specialised fused separable f32 filter (kernel and design notes: fused_ws.cuh;
// instantiations: fused_ws_{p1,p2,w1,w2,g}.cu)
#include "fused_ws.cuh"

namespace sepfilt {

using namespace ws;

bool fused_ws_supported(const FusedVolume& v, const F32Taps taps[3], const F32Taps dtaps[3], bool gradmag)
{
    if (getenv("SEPFILT_NO_WS")) return false;                  // A/B aid
    const bool has_z = !(taps[0].radius == 0 && taps[0].w[0] == 1.0f);
    const int r = taps[2].radius;
    if (taps[1].radius != r || (has_z && taps[0].radius != r)) return false;   // one exact radius for all axes
    if (gradmag) {
        // one launch for the whole gradient magnitude: three z accumulator sets per column fit the
        // register file up to radius 6 (sigma <= 1.5 at truncate 4)
        if (!has_z || r < 1 || r > 6) return false;
        for (int a = 0; a < 3; ++a)
            if (dtaps[a].radius != r) return false;
    } else {
        // plain filters, every radius 1 .. 16.  Radius <= 8: 14- / 16-row tiles, 8 columns per XZ thread — 512^3 reflect
        // sigma 1 / 1.5 / 2: 0.192 / 0.237 / 0.282 ms against 0.228 / 0.289 / 0.330 ms on fused3d.  Radius 9 .. 16 (sigma
        // 2.25 .. 4): 8-row tiles, 4 columns per thread — sigma 2.5 / 3 / 4 0.531 / 0.599 / 0.738 ms against 0.595 / 0.641 /
        // 0.761 ms for three single-axis passes; those tiles re-read a (8 + 2R)-row box per 8 output rows, so with
        // neighbour halos the z-slab sharding hands them LOCAL pads (sharded.py, "pull1"), never planes across NVLink
        if (!has_z || r < 1 || r > WS_MAXR) return false;
    }
    if (v.nx % 4 != 0 || (reinterpret_cast<uintptr_t>(v.in) & 15) || (reinterpret_cast<uintptr_t>(v.out) & 15))
        return false;
    if (v.nx < 16 || v.ny < r + 1 || v.nx < r + 1 || v.nz_in < 1 || v.nz_out < 1) return false;
    if (v.cval != 0.f)
        for (int a = 0; a < 3; ++a)
            if (v.mode[a] == SEPFILT_CONSTANT && taps[a].radius > 0) return false;
    for (int a = 1; a < 3; ++a)
        if (v.mode[a] == SEPFILT_WRAP && taps[a].radius > 0) return false;
    const long long tiles = (long long)((v.nx + 127) / 128) * ((v.ny + 13) / 14);
    if (tiles * 64 > 2147483647LL) return false;
    if (v.halo) {
        if (!has_z || v.z_offset < 0 || v.z_offset + v.nz_out > v.nz_in || v.nz_in < r) return false;
        if ((v.halo->lo && v.halo->planes_lo < r) || (v.halo->hi && v.halo->planes_hi < r)) return false;
    }
    return true;
}

cudaError_t launch_fused_ws(const FusedVolume& v, const F32Taps taps[3], const F32Taps dtaps[3], bool gradmag,
                            cudaStream_t s)
{
    WsParams p;
    std::memset(&p, 0, sizeof p);
    p.in = v.in; p.out = v.out;
    p.nz_in = v.nz_in; p.nz_out = v.nz_out; p.ny = v.ny; p.nx = v.nx; p.z_offset = v.z_offset;
    p.mode_z = v.mode[0]; p.mode_y = v.mode[1]; p.mode_x = v.mode[2];
    put_taps(taps[0], p.wz); put_taps(taps[1], p.wy); put_taps(taps[2], p.wx);
    if (gradmag) { put_taps(dtaps[0], p.dz); put_taps(dtaps[1], p.dy); put_taps(dtaps[2], p.dx); }
    const int sms = device_sms();
    const int r = taps[2].radius;
    if (gradmag) return launch_grad_r1_6(v, p, sms, s, r);
    if (r <= 4) return launch_plain_r1_4(v, p, sms, s, r);
    if (r <= 8) return launch_plain_r5_8(v, p, sms, s, r);
    if (r <= 12) return launch_wide_r9_12(v, p, sms, s, r);
    return launch_wide_r13_16(v, p, sms, s, r);
}

}  // namespace sepfilt
