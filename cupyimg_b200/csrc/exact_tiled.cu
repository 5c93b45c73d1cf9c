// exact_tiled.cu — tiled version of the exact (scipy-arithmetic) 1-D pass for the case that
// covers the named integer workloads: C-contiguous arrays, an odd-length SYMMETRIC or
// ANTI-SYMMETRIC filter (Gaussian and its odd derivatives, [1 2 1], [1 1 1], [-1 0 1], [1 -2 1]).
//
// Same bits as exact.cu / scipy's NI_Correlate1D symmetric branches:
//     acc = x[c] * w[0];  for j = -size1 .. -1:  acc += (x[c+j] +/- x[c-j]) * w[j]
// in float64 with __dmul_rn / __dadd_rn (never contracted), then the C-cast store.  Taps are
// zero-padded to a compile-time radius bucket: the extra outermost pairs add (a +/- b) * 0.0,
// which leaves every finite accumulator unchanged, so the result is bit-identical.  (A generic,
// non-symmetric filter cannot be padded without changing scipy's summation order; it stays on the
// per-element kernel in exact.cu, as do strided arrays and uniform windows.)
//
// Layout: the (outer, n, inner) view of the array.  The tile is staged ONCE into shared memory as
// float64 — each element is converted once, boundary remapping (_util.py:170-228) is resolved once
// per staged cell — then every thread keeps a register window and produces 4 outputs.
//   exact_sym_row_kernel  inner == 1: 4 lines x 256 outputs per CTA, 4 adjacent outputs per thread.
//   exact_sym_col_kernel  inner  > 1: (64 + 2R) x 64 tile, 2 adjacent columns x 4 outputs per thread.
// HBM-bound by bytes (sizeof(in) + sizeof(out) per element); the FP64 pipe and the int<->f64
// conversions are the next limits.
#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

namespace {

// run f.template operator()<T>() with T = the element type of `dtype`: the switch is taken once per
// thread, outside the staging / store loops, instead of once per element
template <class F>
__device__ __forceinline__ void dispatch_dtype(int dtype, F&& f)
{
    switch (dtype) {
    case SEPFILT_I8:  f.template operator()<int8_t>(); break;
    case SEPFILT_U8: case SEPFILT_BOOL: f.template operator()<uint8_t>(); break;
    case SEPFILT_I16: f.template operator()<int16_t>(); break;
    case SEPFILT_U16: f.template operator()<uint16_t>(); break;
    case SEPFILT_I32: f.template operator()<int32_t>(); break;
    case SEPFILT_U32: f.template operator()<uint32_t>(); break;
    case SEPFILT_I64: f.template operator()<int64_t>(); break;
    case SEPFILT_U64: f.template operator()<uint64_t>(); break;
    case SEPFILT_F32: f.template operator()<float>(); break;
    default:          f.template operator()<double>(); break;
    }
}

// typed store under the same cast rules as store_cast (common.cuh), without the per-element switch
template <class T> __device__ __forceinline__ T cast_out(double v);
template <> __device__ __forceinline__ int8_t   cast_out<int8_t>(double v)   { return (int8_t)cvt_x86_i32(v); }
template <> __device__ __forceinline__ uint8_t  cast_out<uint8_t>(double v)  { return (uint8_t)cvt_x86_i32(v); }
template <> __device__ __forceinline__ int16_t  cast_out<int16_t>(double v)  { return (int16_t)cvt_x86_i32(v); }
template <> __device__ __forceinline__ uint16_t cast_out<uint16_t>(double v) { return (uint16_t)cvt_x86_i32(v); }
template <> __device__ __forceinline__ int32_t  cast_out<int32_t>(double v)  { return cvt_x86_i32(v); }
template <> __device__ __forceinline__ uint32_t cast_out<uint32_t>(double v) { return (uint32_t)cvt_x86_i64(v); }
template <> __device__ __forceinline__ int64_t  cast_out<int64_t>(double v)  { return cvt_x86_i64(v); }
template <> __device__ __forceinline__ uint64_t cast_out<uint64_t>(double v)
{
    return (v < 9223372036854775808.0) ? (uint64_t)cvt_x86_i64(v)
                                       : ((uint64_t)cvt_x86_i64(v - 9223372036854775808.0) ^ 0x8000000000000000ull);
}
template <> __device__ __forceinline__ float    cast_out<float>(double v)    { return __double2float_rn(v); }
template <> __device__ __forceinline__ double   cast_out<double>(double v)   { return v; }

struct SymParams {
    const char* in;
    char*       out;
    int32_t     in_dtype, out_dtype, out_size;   // out_size = bytes per output element
    int64_t     outer, inner;
    int32_t     n_in, n_out;
    int32_t     shift;        // source index of the filter centre = output position + shift
    int32_t     mode;
    double      cval;
    double      w[SEPFILT_FAST_MAX_RADIUS + 1];   // w[d] = tap at distance d left of the centre (fw[-d]); w[0] centre
};

// scipy's symmetric / anti-symmetric accumulation for one output whose window is win[0 .. 2R]
template <int R, int SGN>
__device__ __forceinline__ double sym_acc(const double* win, const SymParams& p)
{
    double acc = __dmul_rn(win[R], p.w[0]);
#pragma unroll
    for (int d = R; d >= 1; --d) {           // j = -d: outermost pair first
        const double pair = SGN > 0 ? __dadd_rn(win[R - d], win[R + d]) : __dsub_rn(win[R - d], win[R + d]);
        acc = __dadd_rn(acc, __dmul_rn(pair, p.w[d]));
    }
    return acc;
}

constexpr int XR_W = 256, XR_ROWS = 4;

template <int R, int SGN>
__global__ void __launch_bounds__(256)
exact_sym_row_kernel(const __grid_constant__ SymParams p)
{
    constexpr int PITCH = XR_W + 2 * R;
    __shared__ double tile[XR_ROWS][PITCH];
    const int64_t row0 = (int64_t)blockIdx.x * XR_ROWS;
    const int x0 = blockIdx.y * XR_W;
    const int tid = threadIdx.x;
    const int src0 = x0 + p.shift - R;
    const bool interior = src0 >= 0 && src0 + PITCH <= p.n_in;     // no cell of this tile needs remapping
    dispatch_dtype(p.in_dtype, [&]<class T>() {
        const T* in = reinterpret_cast<const T*>(p.in);
#pragma unroll 2
        for (int i = tid; i < XR_ROWS * PITCH; i += 256) {
            const int r = i / PITCH, s = i - r * PITCH;
            const int64_t row = row0 + r;
            if (row >= p.outer) continue;
            int m = src0 + s;
            if (!interior) m = remap_index32(p.mode, m, p.n_in);
            tile[r][s] = m < 0 ? p.cval : (double)in[row * p.n_in + m];
        }
    });
    __syncthreads();
    const int r = tid >> 6, c = (tid & 63) * 4;
    const int64_t row = row0 + r;
    const int x = x0 + c;
    if (row >= p.outer || x >= p.n_out) return;
    double win[4 + 2 * R];
#pragma unroll
    for (int i = 0; i < 4 + 2 * R; ++i) win[i] = tile[r][c + i];
    double res[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) res[o] = sym_acc<R, SGN>(win + o, p);
    dispatch_dtype(p.out_dtype, [&]<class T>() {
        T* dst = reinterpret_cast<T*>(p.out) + row * p.n_out + x;
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (x + o < p.n_out) dst[o] = cast_out<T>(res[o]);
    });
}

constexpr int XC_TI = 64, XC_RN = 4;

template <int R, int SGN>
__global__ void __launch_bounds__(256)
exact_sym_col_kernel(const __grid_constant__ SymParams p, const int n_itiles)
{
    // 256 threads = 32 (pairs of inner columns) x 8 (groups of XC_RN outputs): tile of 32 output rows x 64 columns
    constexpr int TN = 8 * XC_RN;
    constexpr int ROWS = TN + 2 * R;
    __shared__ __align__(16) double tile[ROWS][XC_TI];
    const int64_t bx = blockIdx.x;
    const int64_t o = bx / n_itiles;
    const int64_t i0 = (bx - o * n_itiles) * XC_TI;
    const int n0 = blockIdx.y * TN;
    const int tid = threadIdx.x;
    const int64_t in_base = o * (int64_t)p.n_in * p.inner;
    const int src0 = n0 + p.shift - R;
    const int rows_needed = min(ROWS, p.n_out - n0 + 2 * R);
    dispatch_dtype(p.in_dtype, [&]<class T>() {
        const int lane = tid & 63;
        const int64_t ii = i0 + lane;
        const T* col = reinterpret_cast<const T*>(p.in) + in_base + ii;
        const bool inside = ii < p.inner;
#pragma unroll 4
        for (int e = tid >> 6; e < rows_needed; e += 4) {
            const int m = remap_index32(p.mode, src0 + e, p.n_in);
            tile[e][lane] = (m >= 0 && inside) ? (double)col[(int64_t)m * p.inner] : p.cval;
        }
    });
    __syncthreads();
    const int ti = tid & 31, tn = tid >> 5;
    const int64_t ii = i0 + 2 * ti;
    const int p0 = n0 + tn * XC_RN;
    if (ii >= p.inner || p0 >= p.n_out) return;
    double wa[XC_RN + 2 * R], wb[XC_RN + 2 * R];
#pragma unroll
    for (int j = 0; j < XC_RN + 2 * R; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(&tile[tn * XC_RN + j][2 * ti]);
        wa[j] = v.x; wb[j] = v.y;
    }
    double ra[XC_RN], rb[XC_RN];
#pragma unroll
    for (int q = 0; q < XC_RN; ++q) { ra[q] = sym_acc<R, SGN>(wa + q, p); rb[q] = sym_acc<R, SGN>(wb + q, p); }
    dispatch_dtype(p.out_dtype, [&]<class T>() {
        T* dst = reinterpret_cast<T*>(p.out) + (o * (int64_t)p.n_out + p0) * p.inner + ii;
#pragma unroll
        for (int q = 0; q < XC_RN; ++q) {
            if (p0 + q >= p.n_out) break;
            dst[(int64_t)q * p.inner] = cast_out<T>(ra[q]);
            if (ii + 1 < p.inner) dst[(int64_t)q * p.inner + 1] = cast_out<T>(rb[q]);
        }
    });
}

int sym_bucket(int r)
{
    static const int buckets[] = {1, 2, 4, 8, 16};
    for (int b : buckets) if (r <= b) return b;
    return -1;
}

template <int R, int SGN>
cudaError_t launch_sym(const SymParams& p, cudaStream_t s)
{
    if (p.inner == 1) {
        dim3 grid((unsigned)((p.outer + XR_ROWS - 1) / XR_ROWS), (unsigned)((p.n_out + XR_W - 1) / XR_W));
        exact_sym_row_kernel<R, SGN><<<grid, 256, 0, s>>>(p);
    } else {
        constexpr int TN = 8 * XC_RN;
        const int64_t n_itiles = (p.inner + XC_TI - 1) / XC_TI;
        dim3 grid((unsigned)(p.outer * n_itiles), (unsigned)((p.n_out + TN - 1) / TN));
        exact_sym_col_kernel<R, SGN><<<grid, 256, 0, s>>>(p, (int)n_itiles);
    }
    return cudaGetLastError();
}

}  // namespace

bool exact_tiled_supported(const ExactTiledGeom& g, int K, int symmetric)
{
    if (!(K & 1) || (symmetric != 1 && symmetric != -1)) return false;
    if (sym_bucket(K / 2) < 0) return false;
    // zero padding is bit-neutral for finite data only (0 * NaN): float inputs need the radius to be the bucket
    if ((g.in_dtype == SEPFILT_F32 || g.in_dtype == SEPFILT_F64) && sym_bucket(K / 2) != K / 2) return false;
    if (g.outer <= 0 || g.inner <= 0 || g.n_in <= 0 || g.n_out <= 0) return false;
    if (g.n_in > 1073741824LL /* 2n must fit an int in the boundary fold */ || g.n_out > 1073741824LL /* 2n must fit an int in the boundary fold */) return false;
    if (g.shift > 1073741824LL || g.shift < -1073741824LL) return false;
    if (g.inner == 1) {
        if ((g.outer + XR_ROWS - 1) / XR_ROWS > 2147483647LL || (g.n_out + XR_W - 1) / XR_W > 65535) return false;
    } else {
        const int64_t n_itiles = (g.inner + XC_TI - 1) / XC_TI;
        if (g.outer * n_itiles > 2147483647LL || (g.n_out + 8 * XC_RN - 1) / (8 * XC_RN) > 65535) return false;
    }
    return true;
}

cudaError_t launch_exact_tiled(const ExactTiledGeom& g, const double* taps, int K, int symmetric, int mode,
                               double cval, cudaStream_t s)
{
    SymParams p;
    p.in = static_cast<const char*>(g.in);
    p.out = static_cast<char*>(g.out);
    p.in_dtype = g.in_dtype;
    p.out_dtype = g.out_dtype;
    p.out_size = dtype_size(g.out_dtype);
    p.outer = g.outer;
    p.inner = g.inner;
    p.n_in = (int32_t)g.n_in;
    p.n_out = (int32_t)g.n_out;
    p.shift = (int32_t)g.shift;
    p.mode = mode;
    p.cval = cval;
    const int r = K / 2;
    for (int d = 0; d <= SEPFILT_FAST_MAX_RADIUS; ++d) p.w[d] = d <= r ? taps[r - d] : 0.0;   // fw[-d]
    const int R = sym_bucket(r);
    const int key = R * 2 + (symmetric > 0 ? 1 : 0);
    switch (key) {
#define CASE(RR) case RR * 2 + 1: return launch_sym<RR, 1>(p, s); case RR * 2: return launch_sym<RR, -1>(p, s);
        CASE(1) CASE(2) CASE(4) CASE(8) CASE(16)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace sepfilt
