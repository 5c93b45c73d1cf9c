// consumers.cu — the elementwise stages of the skimage-level callers of the separable filters:
//   * products of two arrays for the filters-of-products of SSIM (reference
//     skimage/metrics/_structural_similarity.py:203-207: uxx, uyy, uxy) and of the structure tensor
//     (skimage/feature/corner.py:131-134: der0 * der1);
//   * the SSIM map and its cropped mean (_structural_similarity.py:208-233) from the five filtered arrays in ONE
//     pass — the reference runs ~20 cupy elementwise kernels with as many temporaries.
// Every operation is a separately rounded multiply / add / subtract / divide in the array dtype (__f*_rn /
// __d*_rn, no FMA contraction), in the reference's order, so the float64 map has the bits of the numpy
// evaluation of the same formula.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

namespace {

__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  sub_rn(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float  div_rn(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <class T>
__global__ void __launch_bounds__(256)
multiply_kernel(const T* a, const T* b, T* out, int64_t n)     // out may alias a or b
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = mul_rn(a[i], b[i]);
}

struct SsimParams {
    const void *ux, *uy, *uxx, *uyy, *uxy;
    void*   S;                       // full SSIM map (may be nullptr)
    double* sum;                     // += sum of S over the cropped region (float64)
    int64_t n;
    int32_t ndim;
    int64_t shape[3];
    int32_t pad;
    double  cov_norm, C1, C2;
};

template <class T>
__global__ void __launch_bounds__(256)
ssim_kernel(const SsimParams p)
{
    const T cov = (T)p.cov_norm, C1 = (T)p.C1, C2 = (T)p.C2, two = (T)2;
    const T* ux = static_cast<const T*>(p.ux);
    const T* uy = static_cast<const T*>(p.uy);
    const T* uxx = static_cast<const T*>(p.uxx);
    const T* uyy = static_cast<const T*>(p.uyy);
    const T* uxy = static_cast<const T*>(p.uxy);
    T* S = static_cast<T*>(p.S);
    double local = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const T mx = ux[i], my = uy[i];
        // _structural_similarity.py:208-210
        const T vx = mul_rn(cov, sub_rn(uxx[i], mul_rn(mx, mx)));
        const T vy = mul_rn(cov, sub_rn(uyy[i], mul_rn(my, my)));
        const T vxy = mul_rn(cov, sub_rn(uxy[i], mul_rn(mx, my)));
        // :216-223
        const T A1 = add_rn(mul_rn(mul_rn(two, mx), my), C1);
        const T A2 = add_rn(mul_rn(two, vxy), C2);
        const T B1 = add_rn(add_rn(mul_rn(mx, mx), mul_rn(my, my)), C1);
        const T B2 = add_rn(add_rn(vx, vy), C2);
        const T D = mul_rn(B1, B2);
        const T s = div_rn(mul_rn(A1, A2), D);
        if (S) S[i] = s;
        // crop(S, pad): every coordinate in [pad, extent - pad)
        int64_t r = i;
        bool in = true;
        for (int d = p.ndim - 1; d >= 0; --d) {
            const int64_t c = r % p.shape[d];
            r /= p.shape[d];
            in = in && c >= p.pad && c < p.shape[d] - p.pad;
        }
        if (in) local += (double)s;
    }
    // block reduction, one float64 atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(p.sum, local);
}

int grid_for(int64_t n)
{
    const int64_t b = (n + 255) / 256;
    return (int)(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace

cudaError_t launch_multiply(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    if (dtype == SEPFILT_F32)
        multiply_kernel<float><<<grid_for(n), 256, 0, s>>>((const float*)a, (const float*)b, (float*)out, n);
    else if (dtype == SEPFILT_F64)
        multiply_kernel<double><<<grid_for(n), 256, 0, s>>>((const double*)a, (const double*)b, (double*)out, n);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_ssim_map(const void* ux, const void* uy, const void* uxx, const void* uyy, const void* uxy,
                            void* S, double* sum, int ndim, const int64_t* shape, int pad, double cov_norm,
                            double C1, double C2, int dtype, cudaStream_t s)
{
    SsimParams p;
    p.ux = ux; p.uy = uy; p.uxx = uxx; p.uyy = uyy; p.uxy = uxy; p.S = S; p.sum = sum;
    p.ndim = ndim; p.pad = pad; p.cov_norm = cov_norm; p.C1 = C1; p.C2 = C2;
    p.n = 1;
    for (int d = 0; d < 3; ++d) { p.shape[d] = d < ndim ? shape[d] : 1; if (d < ndim) p.n *= shape[d]; }
    if (p.n <= 0) return cudaSuccess;
    if (dtype == SEPFILT_F32) ssim_kernel<float><<<grid_for(p.n), 256, 0, s>>>(p);
    else if (dtype == SEPFILT_F64) ssim_kernel<double><<<grid_for(p.n), 256, 0, s>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace sepfilt
