// exact.cu — the "exact" 1-D correlation pass: scipy's arithmetic on the GPU.
//
// float64 accumulation in scipy's NI_Correlate1D summation order (symmetric /
// anti-symmetric / generic) with explicit __dmul_rn / __dadd_rn so
// that nvcc never contracts into FMA, then the C-cast store rules (App. C.4).  This is
// what makes integer outputs bit-exact with scipy.ndimage and float64 outputs
// bit-identical.  Any (in, out) dtype pair, any byte strides of either sign, 64-bit
// indexing, any filter length (K > array length = multi-reflection included).
//
// Replaces one launch of the reference's generated ElementwiseKernel
// (_filters_core.py:190-348, body in SURVEY.md App. A).  Unlike that kernel the
// boundary rule is evaluated only for threads whose window leaves the array.
#include <cmath>
#include <type_traits>
#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

template <typename InT>
__global__ void __launch_bounds__(256)
exact_corr1d_kernel(const __grid_constant__ ExactParams p)
{
    const int K = p.K;
    const int size1 = K / 2, size2 = K - size1 - 1;
    const int64_t astride = p.istride[p.axis];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
        // unravel the output index (C order); dims were collapsed on the host
        int64_t rem = idx, ioff = 0, ooff = 0, pos = 0;
#pragma unroll 1
        for (int d = p.ndim - 1; d >= 0; --d) {
            const int64_t ext = p.shape[d];
            const int64_t c = rem % ext;
            rem /= ext;
            if (d == p.axis) pos = c; else ioff += c * p.istride[d];
            ooff += c * p.ostride[d];
        }
        const char* line = p.in + ioff;
        const int64_t start = pos + p.in_offset - p.before;       // source index of tap 0
        const bool interior = start >= 0 && start + K <= p.n_in;
        // il(j): the boundary-extended input at offset j from the filter centre (j = -size1..size2)
        auto il = [&](int j) -> double {
            int64_t src = start + size1 + j;
            if (!interior) {
                src = remap_index(p.mode, src, p.n_in);
                if (src < 0) return p.cval;
            }
            return load_as_double<InT>(line + src * astride);
        };
        auto fw = [&](int j) -> double { return p.wdev ? p.wdev[size1 + j] : p.w[size1 + j]; };
        double acc;
        if (p.symmetric == 1) {
            acc = __dmul_rn(il(0), fw(0));
            for (int j = -size1; j < 0; ++j)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(il(j), il(-j)), fw(j)));
        } else if (p.symmetric == -1) {
            acc = __dmul_rn(il(0), fw(0));
            for (int j = -size1; j < 0; ++j)
                acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(il(j), il(-j)), fw(j)));
        } else if (p.symmetric == 0) {
            acc = __dmul_rn(il(size2), fw(size2));
            for (int j = -size1; j < size2; ++j)
                acc = __dadd_rn(acc, __dmul_rn(il(j), fw(j)));
        } else if (p.symmetric == 2) {
            // uniform_filter1d: exact window sum, one division
            acc = 0.0;
            for (int j = -size1; j <= size2; ++j) acc = __dadd_rn(acc, il(j));
            acc = __ddiv_rn(acc, (double)K);
        } else {
            // minimum_filter1d (3) / maximum_filter1d (4): the reference's generated kernel compares in
            // double (filters.py:1511-1557, "value = min(cast<double>(x), value)"), C comparison semantics
            acc = il(-size1);
            if (p.symmetric == 3) {
                for (int j = -size1 + 1; j <= size2; ++j) { const double v = il(j); acc = v < acc ? v : acc; }
            } else {
                for (int j = -size1 + 1; j <= size2; ++j) { const double v = il(j); acc = v > acc ? v : acc; }
            }
        }
        store_cast(p.out + ooff, p.out_dtype, acc);
    }
}

// uniform_filter1d with scipy's running sum, one thread per line:
//     tmp = sum of the first window;  out[0] = tmp / K;  tmp += x[l + K - 1] - x[l - 1];  out[l] = tmp / K
// The window-sum kernel above gives the same bits whenever every partial sum is exact (integer data, integer
// cval).  It does not when the sums round — float input, or constant mode with a fractional cval — and the
// output is an integer type, where one ulp can flip a truncation; those calls take this sequential kernel.
template <typename InT>
__global__ void __launch_bounds__(128)
uniform_running_kernel(const __grid_constant__ ExactParams p, const int64_t lines)
{
    const int K = p.K;
    const int64_t astride = p.istride[p.axis], ostr = p.ostride[p.axis], n_out = p.shape[p.axis];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ln < lines; ln += stride) {
        int64_t rem = ln, ioff = 0, ooff = 0;
#pragma unroll 1
        for (int d = p.ndim - 1; d >= 0; --d) {
            if (d == p.axis) continue;
            const int64_t ext = p.shape[d];
            const int64_t c = rem % ext;
            rem /= ext;
            ioff += c * p.istride[d];
            ooff += c * p.ostride[d];
        }
        const char* line = p.in + ioff;
        const int64_t first = p.in_offset - p.before;            // source index of extended element 0
        auto ext_at = [&](int64_t l) -> double {
            const int64_t src = remap_index(p.mode, first + l, p.n_in);
            return src < 0 ? p.cval : load_as_double<InT>(line + src * astride);
        };
        double tmp = 0.0;
        for (int l = 0; l < K; ++l) tmp = __dadd_rn(tmp, ext_at(l));
        store_cast(p.out + ooff, p.out_dtype, __ddiv_rn(tmp, (double)K));
        for (int64_t l = 1; l < n_out; ++l) {
            tmp = __dadd_rn(tmp, __dsub_rn(ext_at(l + K - 1), ext_at(l - 1)));
            store_cast(p.out + ooff + l * ostr, p.out_dtype, __ddiv_rn(tmp, (double)K));
        }
    }
}

static bool needs_running_sum(const ExactParams& p)
{
    if (p.symmetric != 2) return false;
    if (p.out_dtype == SEPFILT_F32 || p.out_dtype == SEPFILT_F64) return false;   // float outputs: rtol contract
    if (p.in_dtype == SEPFILT_F32 || p.in_dtype == SEPFILT_F64) return true;
    return p.mode == SEPFILT_CONSTANT && p.cval != floor(p.cval);
}

cudaError_t launch_exact_corr1d(const ExactParams& p, cudaStream_t s)
{
    if (p.total <= 0) return cudaSuccess;
    if (needs_running_sum(p)) {
        const int64_t lines = p.total / p.shape[p.axis];
        int64_t b64 = (lines + 127) / 128;
        const int blocks = (int)(b64 < 148 * 16 ? b64 : 148 * 16);
        switch (p.in_dtype) {
#define CASE(T, C) case T: uniform_running_kernel<C><<<blocks, 128, 0, s>>>(p, lines); break;
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
    const int threads = 256;
    int64_t blocks64 = (p.total + threads - 1) / threads;
    const int64_t cap = 148 * 32;   // grid-stride beyond 32 CTAs per SM
    int blocks = (int)(blocks64 < cap ? blocks64 : cap);
    switch (p.in_dtype) {
#define CASE(T, C) case T: exact_corr1d_kernel<C><<<blocks, threads, 0, s>>>(p); break;
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

// ---- generic_gradient_magnitude epilogue in the output dtype (filters.py:1187-1201) ----
template <typename T>
__global__ void __launch_bounds__(256)
gradmag_step_kernel(T* acc, const T* a, int64_t n, int dtype, int op)   // acc may alias a (ops 0 / 2 are unary)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if constexpr (std::is_same<T, float>::value) {
            if (op == 0) acc[i] = __fmul_rn(a[i], a[i]);
            else if (op == 1) acc[i] = __fadd_rn(acc[i], __fmul_rn(a[i], a[i]));
            else if (op == 2) acc[i] = __fsqrt_rn(acc[i]);
            else if (op == 4) acc[i] = __fsub_rn(acc[i], a[i]);
            else acc[i] = __fadd_rn(acc[i], a[i]);
        } else if constexpr (std::is_same<T, double>::value) {
            if (op == 0) acc[i] = __dmul_rn(a[i], a[i]);
            else if (op == 1) acc[i] = __dadd_rn(acc[i], __dmul_rn(a[i], a[i]));
            else if (op == 2) acc[i] = __dsqrt_rn(acc[i]);
            else if (op == 4) acc[i] = __dsub_rn(acc[i], a[i]);
            else acc[i] = __dadd_rn(acc[i], a[i]);
        } else {                                                        // integers wrap
            const uint64_t x = (uint64_t)(int64_t)a[i];
            const uint64_t xx = x * x;
            if (op == 0) acc[i] = (T)xx;
            else if (op == 1) acc[i] = (T)((uint64_t)(int64_t)acc[i] + xx);
            else if (op == 2) store_cast(reinterpret_cast<char*>(acc + i), dtype, __dsqrt_rn((double)acc[i]));
            else if (op == 4) acc[i] = (T)((uint64_t)(int64_t)acc[i] - x);
            else acc[i] = (T)((uint64_t)(int64_t)acc[i] + x);
        }
    }
}

cudaError_t launch_gradmag_step(void* acc, const void* a, int64_t n, int dtype, int op, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t b64 = (n + threads - 1) / threads;
    int blocks = (int)(b64 < 148 * 32 ? b64 : 148 * 32);
    switch (dtype) {
#define CASE(T, C) case T: gradmag_step_kernel<C><<<blocks, threads, 0, s>>>((C*)acc, (const C*)a, n, dtype, op); break;
        CASE(SEPFILT_I8, int8_t) CASE(SEPFILT_U8, uint8_t) CASE(SEPFILT_I16, int16_t)
        CASE(SEPFILT_U16, uint16_t) CASE(SEPFILT_I32, int32_t) CASE(SEPFILT_U32, uint32_t)
        CASE(SEPFILT_I64, int64_t) CASE(SEPFILT_U64, uint64_t)
        CASE(SEPFILT_F32, float) CASE(SEPFILT_F64, double)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace sepfilt
