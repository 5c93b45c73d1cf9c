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

// ---- kernels that act on the last two axes of C-contiguous arrays (3x3, 5x5, 1xK, Kx1 ... on images or
// stacks of images): the common case of this entry point.  Same arithmetic and tap order; what goes away is
// the N-d bookkeeping — 32-bit plane indexing, no coordinate arrays in local memory, and threads whose
// footprint lies inside the plane never touch the boundary rule.
template <typename InT>
__global__ void __launch_bounds__(256)
correlate_2d_kernel(const __grid_constant__ CorrNdParams p, const int ny, const int nx, const int kh, const int kw,
                    const int by, const int bx, const int64_t planes)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= nx || y >= ny) return;
    const int sy = y - by, sx = x - bx;                       // source coordinate of tap (0, 0)
    const bool interior = sy >= 0 && sy + kh <= ny && sx >= 0 && sx + kw <= nx;
    const int osize = dtype_size(p.out_dtype);
    for (int64_t z = blockIdx.z; z < planes; z += gridDim.z) {
        const InT* plane = reinterpret_cast<const InT*>(p.in) + z * (int64_t)ny * nx;
        double acc = 0.0;
        if (interior) {
            const InT* src = plane + (int64_t)sy * nx + sx;
            for (int ky = 0, k = 0; ky < kh; ++ky, src += nx)
                for (int kx = 0; kx < kw; ++kx, ++k) {
                    const double w = p.wdev ? p.wdev[k] : p.w[k];
                    if (fabs(w) > 2.220446049250313e-16) acc = __dadd_rn(acc, __dmul_rn((double)src[kx], w));
                }
        } else {
            for (int ky = 0, k = 0; ky < kh; ++ky) {
                const int my = remap_index32(p.mode, sy + ky, ny);
                for (int kx = 0; kx < kw; ++kx, ++k) {
                    const double w = p.wdev ? p.wdev[k] : p.w[k];
                    if (!(fabs(w) > 2.220446049250313e-16)) continue;
                    const int mx = remap_index32(p.mode, sx + kx, nx);
                    const double v = (my < 0 || mx < 0) ? p.cval : (double)plane[(int64_t)my * nx + mx];
                    acc = __dadd_rn(acc, __dmul_rn(v, w));
                }
            }
        }
        store_cast(p.out + (z * (int64_t)ny * nx + (int64_t)y * nx + x) * osize, p.out_dtype, acc);
    }
}

static bool contiguous_c(const CorrNdParams& p, const int64_t* stride, int esize)
{
    int64_t expect = esize;
    for (int d = p.ndim - 1; d >= 0; --d) {
        if (p.shape[d] != 1 && stride[d] != expect) return false;
        expect *= p.shape[d];
    }
    return true;
}

template <typename InT>
static bool try_launch_2d(const CorrNdParams& p, cudaStream_t s, cudaError_t* err)
{
    if (p.ndim < 2) return false;
    for (int d = 0; d < p.ndim - 2; ++d)
        if (p.wshape[d] != 1) return false;
    if (!contiguous_c(p, p.istride, dtype_size(p.in_dtype)) || !contiguous_c(p, p.ostride, dtype_size(p.out_dtype)))
        return false;
    const int64_t ny = p.shape[p.ndim - 2], nx = p.shape[p.ndim - 1];
    if (ny * nx > 2147483647LL || ny > (1 << 30) || nx > (1 << 30)) return false;
    int64_t planes = 1;
    for (int d = 0; d < p.ndim - 2; ++d) planes *= p.shape[d];
    if ((ny + 7) / 8 > 65535) return false;
    dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((ny + 7) / 8), (unsigned)(planes < 65535 ? planes : 65535));
    correlate_2d_kernel<InT><<<grid, 256, 0, s>>>(p, (int)ny, (int)nx, p.wshape[p.ndim - 2], p.wshape[p.ndim - 1],
                                                  p.before[p.ndim - 2], p.before[p.ndim - 1], planes);
    *err = cudaGetLastError();
    return true;
}

cudaError_t launch_correlate_nd(const CorrNdParams& p, cudaStream_t s)
{
    if (p.total <= 0) return cudaSuccess;
    {
        cudaError_t e = cudaSuccess;
        bool done = false;
        switch (p.in_dtype) {
#define CASE(T, C) case T: done = try_launch_2d<C>(p, s, &e); break;
            CASE(SEPFILT_I8, int8_t) CASE(SEPFILT_U8, uint8_t) CASE(SEPFILT_BOOL, uint8_t)
            CASE(SEPFILT_I16, int16_t) CASE(SEPFILT_U16, uint16_t)
            CASE(SEPFILT_I32, int32_t) CASE(SEPFILT_U32, uint32_t)
            CASE(SEPFILT_I64, int64_t) CASE(SEPFILT_U64, uint64_t)
            CASE(SEPFILT_F32, float) CASE(SEPFILT_F64, double)
#undef CASE
        default: break;
        }
        if (done) return e;
    }
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
