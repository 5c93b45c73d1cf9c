// kernels.h — host-side launcher interface between api.cu and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sepfilt.h"

namespace sepfilt {

// ---- exact path (exact.cu): float64, scipy order, any dtype pair, any strides ----
struct ExactParams {
    const char* in;
    char*       out;
    int32_t     in_dtype, out_dtype;
    int32_t     ndim;                    // collapsed rank (>= 1), dims in C order
    int32_t     axis;
    int64_t     shape[SEPFILT_MAX_NDIM]; // OUTPUT extents
    int64_t     istride[SEPFILT_MAX_NDIM];
    int64_t     ostride[SEPFILT_MAX_NDIM];
    int64_t     n_in;                    // input extent along axis
    int64_t     in_offset;               // out position p <-> in position p + in_offset
    int64_t     total;                   // number of output elements
    int32_t     K, before;               // before = K/2 + origin
    int32_t     mode;
    int32_t     symmetric;               // +1 / -1 / 0 (scipy's probe), 2 = uniform window sum, 3 / 4 = window minimum / maximum
    double      cval;
    const double* wdev;                  // taps in device memory (K > SEPFILT_PARAM_TAPS) or nullptr
    double      w[SEPFILT_PARAM_TAPS];
};
cudaError_t launch_exact_corr1d(const ExactParams& p, cudaStream_t s);
cudaError_t launch_gradmag_step(void* acc, const void* a, int64_t n, int dtype, int op, cudaStream_t s);

// ---- dense N-d correlation, exact arithmetic (correlate_nd.cu) ----
struct CorrNdParams {
    const char* in;
    char*       out;
    int32_t     in_dtype, out_dtype;
    int32_t     ndim, K;                    // K = product of wshape
    int64_t     shape[SEPFILT_MAX_NDIM];    // input == output extents
    int64_t     istride[SEPFILT_MAX_NDIM];
    int64_t     ostride[SEPFILT_MAX_NDIM];
    int32_t     wshape[SEPFILT_MAX_NDIM];
    int32_t     before[SEPFILT_MAX_NDIM];   // wshape / 2 + origin per axis
    int64_t     total;
    int32_t     mode;
    double      cval;
    const double* wdev;                     // weights in device memory (K > SEPFILT_PARAM_TAPS) or nullptr
    double      w[SEPFILT_PARAM_TAPS];
};
cudaError_t launch_correlate_nd(const CorrNdParams& p, cudaStream_t s);

// ---- exact path, tiled (exact_tiled.cu): C-contiguous arrays, odd symmetric / anti-symmetric taps ----
struct ExactTiledGeom {
    const void* in;
    void*       out;
    int32_t     in_dtype, out_dtype;
    int64_t     outer, n_in, n_out, inner;   // (outer, n, inner) view
    int64_t     shift;                       // source index of the filter centre = output position + shift
};
bool        exact_tiled_supported(const ExactTiledGeom& g, int K, int symmetric);
cudaError_t launch_exact_tiled(const ExactTiledGeom& g, const double* taps, int K, int symmetric, int mode,
                               double cval, cudaStream_t s);

// ---- exact path, register streaming (exact_stream.cu): dtype-preserving u8 / i16 / u16 / f32 / f64,
//      radius <= 8, vector-aligned geometry; same bits as the two kernels above ----
bool        exact_stream_supported(const ExactTiledGeom& g, int K, int symmetric);
cudaError_t launch_exact_stream(const ExactTiledGeom& g, const double* taps, int K, int symmetric, int mode,
                                double cval, cudaStream_t s);

// ---- minimum / maximum window passes, register streaming (minmax_stream.cu) ----
bool        minmax_stream_supported(const ExactTiledGeom& g, int S, int origin, double cval);
cudaError_t launch_minmax_stream(const ExactTiledGeom& g, int S, int origin, int mode, double cval, bool is_max,
                                 cudaStream_t s);

// ---- f32 tiled 1-D passes (f32_1d.cu): C-contiguous (outer, n, inner) view ----
struct F32Taps {
    int32_t radius;                       // R: taps cover offsets -R..R (zero padded)
    float   w[2 * SEPFILT_FAST_MAX_RADIUS + 1];
};
struct F32Line {
    const float* in;
    float*       out;
    int64_t      outer;      // product of extents before the axis
    int32_t      n_in, n_out;
    int64_t      inner;      // product of extents after the axis (1 = contiguous axis)
    int32_t      in_offset;
    int32_t      mode;
    float        cval;
};
bool        f32_line_supported(const F32Line& g, int radius);
// register-streaming versions (f32_stream.cu) for vector-aligned geometries; same contract
bool        f32_stream_supported(const F32Line& g, int radius);
cudaError_t launch_f32_stream(const F32Line& g, const F32Taps& t, cudaStream_t s);
cudaError_t launch_f32_corr1d(const F32Line& g, const F32Taps& t, cudaStream_t s);

// ---- fused multi-axis f32 (fused3d.cu) ----
struct FusedVolume {
    const float* in;
    float*       out;
    int32_t      nz_in, nz_out, ny, nx;   // 2-D inputs use nz = 1 with an identity z pass
    int32_t      z_offset;                // out plane z <-> in plane z + z_offset
    int32_t      mode[3];                 // per axis (z, y, x)
    float        cval;
    const sepfilt_halo* halo;             // neighbour planes beyond the slab's ends (multi-GPU), or nullptr
};
bool        fused3d_supported(const FusedVolume& v, const F32Taps taps[3], bool gradmag);
cudaError_t launch_fused3d(const FusedVolume& v, const F32Taps taps[3], const F32Taps dtaps[3],
                           bool gradmag, cudaStream_t s);

// ---- warp-specialised fused f32 (fused_ws.cu): same contract as fused3d, tried first ----
bool        fused_ws_supported(const FusedVolume& v, const F32Taps taps[3], const F32Taps dtaps[3], bool gradmag);
cudaError_t launch_fused_ws(const FusedVolume& v, const F32Taps taps[3], const F32Taps dtaps[3],
                            bool gradmag, cudaStream_t s);

// ---- elementwise stages of the skimage-level consumers (consumers.cu) ----
cudaError_t launch_multiply(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t s);
cudaError_t launch_ssim_map(const void* ux, const void* uy, const void* uxx, const void* uyy, const void* uxy,
                            void* S, double* sum, int ndim, const int64_t* shape, int pad, double cov_norm,
                            double C1, double C2, int dtype, cudaStream_t s);

}  // namespace sepfilt
