// correlate_nd.cu — dense N-d correlation for small kernels: one launch of the
// reference's generated correlate kernel for N-d weights (filters.py:65-210 -> _correlate_or_convolve
// :441-495, kernel body _filters_core.py:239-312).
//
// scipy's NI_Correlate arithmetic: float64, `tmp += x * w` over the taps with |w| > DBL_EPSILON in C order of
// the weights array (explicit __dmul_rn / __dadd_rn, never contracted), boundary extension per axis
// (_util.py:170-228), then the C-cast store — bit-exact integer outputs, bit-identical float64.  Any (in, out)
// dtype pair, any byte strides, up to SEPFILT_MAX_NDIM dimensions.  One output element per thread
// (grid-stride); the boundary rule is evaluated only for threads whose footprint leaves the array.
#include <type_traits>
#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

template <typename InT>
__global__ void __launch_bounds__(256)
correlate_nd_kernel(const __grid_constant__ CorrNdParams p)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
        int64_t start[SEPFILT_MAX_NDIM];                      // source coordinate of tap (0, .., 0)
        int64_t rem = idx, ooff = 0;
        bool interior = true;
#pragma unroll 1
        for (int d = p.ndim - 1; d >= 0; --d) {
            const int64_t ext = p.shape[d];
            const int64_t c = rem % ext;
            rem /= ext;
            ooff += c * p.ostride[d];
            start[d] = c - p.before[d];
            interior = interior && start[d] >= 0 && start[d] + p.wshape[d] <= ext;
        }
        int kk[SEPFILT_MAX_NDIM];
#pragma unroll 1
        for (int d = 0; d < p.ndim; ++d) kk[d] = 0;
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < p.K; ++k) {
            const double w = p.wdev ? p.wdev[k] : p.w[k];
            if (fabs(w) > 2.220446049250313e-16) {             // scipy's footprint: |w| > DBL_EPSILON
                int64_t ioff = 0;
                bool outside = false;
#pragma unroll 1
                for (int d = 0; d < p.ndim; ++d) {
                    int64_t s = start[d] + kk[d];
                    if (!interior) {
                        s = remap_index(p.mode, s, p.shape[d]);
                        if (s < 0) { outside = true; break; }
                    }
                    ioff += s * p.istride[d];
                }
                const double v = outside ? p.cval : load_as_double<InT>(p.in + ioff);
                acc = __dadd_rn(acc, __dmul_rn(v, w));
            }
#pragma unroll 1
            for (int d = p.ndim - 1; d >= 0; --d) {            // odometer over the weights shape, C order
                if (++kk[d] < p.wshape[d]) break;
                kk[d] = 0;
            }
        }
        store_cast(p.out + ooff, p.out_dtype, acc);
    }
}

cudaError_t launch_correlate_nd(const CorrNdParams& p, cudaStream_t s)
{
    if (p.total <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks64 = (p.total + threads - 1) / threads;
    const int64_t cap = 148 * 32;
    int blocks = (int)(blocks64 < cap ? blocks64 : cap);
    switch (p.in_dtype) {
#define CASE(T, C) case T: correlate_nd_kernel<C><<<blocks, threads, 0, s>>>(p); break;
        CASE(SEPFILT_I8, int8_t) CASE(SEPFILT_U8, uint8_t) CASE(SEPFILT_BOOL, uint8_t)
        CASE(SEPFILT_I16, int16_t) CASE(SEPFILT_U16, uint16_t)
        CASE(SEPFILT_I32, int32_t) CASE(SEPFILT_U32, uint32_t)
        CASE(SEPFILT_I64, int64_t) CASE(SEPFILT_U64, uint64_t)
        CASE(SEPFILT_F32, float) CASE(SEPFILT_F64, double)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace sepfilt
