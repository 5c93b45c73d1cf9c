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

__device__ __forceinline__ double load_any(const char* base, int dtype, int64_t idx)
{
    switch (dtype) {
    case SEPFILT_I8:  return (double)reinterpret_cast<const int8_t*>(base)[idx];
    case SEPFILT_U8: case SEPFILT_BOOL: return (double)reinterpret_cast<const uint8_t*>(base)[idx];
    case SEPFILT_I16: return (double)reinterpret_cast<const int16_t*>(base)[idx];
    case SEPFILT_U16: return (double)reinterpret_cast<const uint16_t*>(base)[idx];
    case SEPFILT_I32: return (double)reinterpret_cast<const int32_t*>(base)[idx];
    case SEPFILT_U32: return (double)reinterpret_cast<const uint32_t*>(base)[idx];
    case SEPFILT_I64: return (double)reinterpret_cast<const int64_t*>(base)[idx];
    case SEPFILT_U64: return (double)reinterpret_cast<const uint64_t*>(base)[idx];
    case SEPFILT_F32: return (double)reinterpret_cast<const float*>(base)[idx];
    default:          return reinterpret_cast<const double*>(base)[idx];
    }
}

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
    for (int i = tid; i < XR_ROWS * PITCH; i += 256) {
        const int r = i / PITCH, s = i - r * PITCH;
        const int64_t row = row0 + r;
        if (row >= p.outer) continue;
        const int m = remap_index32(p.mode, src0 + s, p.n_in);
        tile[r][s] = m < 0 ? p.cval : load_any(p.in, p.in_dtype, row * p.n_in + m);
    }
    __syncthreads();
    const int r = tid >> 6, c = (tid & 63) * 4;
    const int64_t row = row0 + r;
    const int x = x0 + c;
    if (row >= p.outer || x >= p.n_out) return;
    double win[4 + 2 * R];
#pragma unroll
    for (int i = 0; i < 4 + 2 * R; ++i) win[i] = tile[r][c + i];
    char* dst = p.out + (row * p.n_out + x) * p.out_size;
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (x + o < p.n_out) store_cast(dst + o * p.out_size, p.out_dtype, sym_acc<R, SGN>(win + o, p));
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
    {
        const int lane = tid & 63;
        const int64_t ii = i0 + lane;
        for (int e = tid >> 6; e < rows_needed; e += 4) {
            const int m = remap_index32(p.mode, src0 + e, p.n_in);
            double v = p.cval;
            if (m >= 0 && ii < p.inner) v = load_any(p.in, p.in_dtype, in_base + (int64_t)m * p.inner + ii);
            tile[e][lane] = v;
        }
    }
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
#pragma unroll
    for (int q = 0; q < XC_RN; ++q) {
        const int pp = p0 + q;
        if (pp >= p.n_out) break;
        char* dst = p.out + ((o * (int64_t)p.n_out + pp) * p.inner + ii) * p.out_size;
        store_cast(dst, p.out_dtype, sym_acc<R, SGN>(wa + q, p));
        if (ii + 1 < p.inner) store_cast(dst + p.out_size, p.out_dtype, sym_acc<R, SGN>(wb + q, p));
    }
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
    if (g.outer <= 0 || g.inner <= 0 || g.n_in <= 0 || g.n_out <= 0) return false;
    if (g.n_in > 2147483647LL - 4096 || g.n_out > 2147483647LL - 4096) return false;
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
