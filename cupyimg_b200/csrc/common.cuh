// common.cuh — shared device helpers for libsepfilt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sepfilt.h"

namespace sepfilt {

// ---- boundary index remapping -------------------------------------------------
// Same rules as the reference's _util._generate_boundary_condition_ops
// (_util.py:170-228) but evaluated ONLY for out-of-range indices, i.e. at the
// line / tile edge, never per tap in the interior.  Returns -1 for "use cval".
__host__ __device__ __forceinline__ int64_t remap_index(int mode, int64_t ix, int64_t n)
{
    if (ix >= 0 && ix < n) return ix;
    switch (mode) {
    case SEPFILT_REFLECT: {
        if (ix < 0) ix = -1 - ix;
        ix %= 2 * n;
        int64_t m = 2 * n - 1 - ix;
        return ix < m ? ix : m;
    }
    case SEPFILT_MIRROR: {
        if (n == 1) return 0;
        if (ix < 0) ix = -ix;
        ix = 1 + (ix - 1) % (2 * n - 2);
        int64_t m = 2 * n - 2 - ix;
        return ix < m ? ix : m;
    }
    case SEPFILT_NEAREST:
        return ix < 0 ? 0 : n - 1;
    case SEPFILT_WRAP:
        ix %= n;
        return ix < 0 ? ix + n : ix;
    default:
        return -1;
    }
}

// 32-bit variant for the tiled kernels (extents < 2^31)
__device__ __forceinline__ int remap_index32(int mode, int ix, int n)
{
    if (ix >= 0 && ix < n) return ix;
    switch (mode) {
    case SEPFILT_REFLECT: {
        if (ix < 0) ix = -1 - ix;
        if (ix >= 2 * n) ix %= 2 * n;
        int m = 2 * n - 1 - ix;
        return ix < m ? ix : m;
    }
    case SEPFILT_MIRROR: {
        if (n == 1) return 0;
        if (ix < 0) ix = -ix;
        ix = 1 + (ix - 1) % (2 * n - 2);
        int m = 2 * n - 2 - ix;
        return ix < m ? ix : m;
    }
    case SEPFILT_NEAREST:
        return ix < 0 ? 0 : n - 1;
    case SEPFILT_WRAP:
        ix %= n;
        return ix < 0 ? ix + n : ix;
    default:
        return -1;
    }
}

// ---- element load / store under scipy's cast rules -----------
template <typename T> __device__ __forceinline__ double load_as_double(const char* p)
{
    return (double)(*reinterpret_cast<const T*>(p));
}

// x86 cvttsd2si semantics: truncation toward zero, "integer indefinite" when out of range
__device__ __forceinline__ int32_t cvt_x86_i32(double v)
{
    return (v > -2147483649.0 && v < 2147483648.0) ? __double2int_rz(v) : INT32_MIN;
}
__device__ __forceinline__ int64_t cvt_x86_i64(double v)
{
    return (v >= -9223372036854775808.0 && v < 9223372036854775808.0) ? __double2ll_rz(v)
                                                                       : INT64_MIN;
}

__device__ __forceinline__ void store_cast(char* p, int dtype, double v)
{
    switch (dtype) {
    case SEPFILT_I8:  *reinterpret_cast<int8_t*>(p) = (int8_t)cvt_x86_i32(v); break;
    case SEPFILT_U8:  *reinterpret_cast<uint8_t*>(p) = (uint8_t)cvt_x86_i32(v); break;
    case SEPFILT_I16: *reinterpret_cast<int16_t*>(p) = (int16_t)cvt_x86_i32(v); break;
    case SEPFILT_U16: *reinterpret_cast<uint16_t*>(p) = (uint16_t)cvt_x86_i32(v); break;
    case SEPFILT_I32: *reinterpret_cast<int32_t*>(p) = cvt_x86_i32(v); break;
    case SEPFILT_U32: *reinterpret_cast<uint32_t*>(p) = (uint32_t)cvt_x86_i64(v); break;
    case SEPFILT_I64: *reinterpret_cast<int64_t*>(p) = cvt_x86_i64(v); break;
    case SEPFILT_U64:
        *reinterpret_cast<uint64_t*>(p) =
            (v < 9223372036854775808.0)
                ? (uint64_t)cvt_x86_i64(v)
                : ((uint64_t)cvt_x86_i64(v - 9223372036854775808.0) ^ 0x8000000000000000ull);
        break;
    case SEPFILT_F32: *reinterpret_cast<float*>(p) = __double2float_rn(v); break;
    default:          *reinterpret_cast<double*>(p) = v; break;
    }
}

__host__ __device__ __forceinline__ int dtype_size(int t)
{
    switch (t) {
    case SEPFILT_I8: case SEPFILT_U8: case SEPFILT_BOOL: return 1;
    case SEPFILT_I16: case SEPFILT_U16: return 2;
    case SEPFILT_I32: case SEPFILT_U32: case SEPFILT_F32: return 4;
    case SEPFILT_I64: case SEPFILT_U64: case SEPFILT_F64: return 8;
    default: return 0;
    }
}

}  // namespace sepfilt
