// minmax_stream.cu — register-streaming minimum_filter1d / maximum_filter1d passes
// for dtype-preserving u8 / i16 / u16 / f32 / f64 arrays, window sizes 2..9.
//
// Same results as the general kernel in exact.cu (symmetric = 3 / 4): the reference compares in double
// (filters.py:1511-1557) and stores with the C cast; for these types the comparison of two array elements in T
// is the comparison of their exact double images, and `cval` is folded to T on the host only when that does not
// change any outcome (api.cu).  What changes is the data movement, as in f32_stream.cu:
//   minmax_stream_col_kernel  strided axis: a thread owns C adjacent columns (one 4..16 byte load per row) and
//       marches a segment of the axis; every input row is folded at once into S per-column running extrema that
//       shift by one output per step (loop unrolled by S: logical slot j at step s is physical (j + s) mod S).
//   minmax_stream_row_kernel  contiguous axis, centred windows (origin 0): a thread loads the 16-byte chunks of
//       its outputs plus the halo chunks from global memory and folds S - 1 neighbours per output.
// Neither uses shared memory or a CTA barrier.
#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

namespace {

struct MStreamParams {
    const char* in;
    char*       out;
    int64_t     outer, inner;
    int32_t     n;                   // extent of the filtered axis (input == output)
    int32_t     before;              // window of output p = [p - before, p - before + S)
    int32_t     mode;
    int32_t     seg, xblocks, gpr, row_aligned;
    double      cval;                // exactly representable in T (checked on the host)
};

__device__ __noinline__ int mremap_outside(int mode, int ix, int n) { return remap_index32(mode, ix, n); }
__device__ __forceinline__ int mremap_fast(int mode, int ix, int n)
{
    if ((unsigned)ix < (unsigned)n) return ix;
    if (mode == SEPFILT_CONSTANT) return -1;
    if (mode == SEPFILT_NEAREST) return ix < 0 ? 0 : n - 1;
    if (ix > -n && ix < 2 * n - 1) {
        const bool low = ix < 0;
        if (mode == SEPFILT_REFLECT) return low ? -1 - ix : 2 * n - 1 - ix;
        if (mode == SEPFILT_MIRROR) return low ? -ix : 2 * n - 2 - ix;
        return low ? ix + n : ix - n;
    }
    return mremap_outside(mode, ix, n);
}

template <class T, int C> struct alignas(sizeof(T) * C) MPack { T v[C]; };

template <bool MAX, class T> __device__ __forceinline__ T fold(T a, T b)
{
    return MAX ? (b > a ? b : a) : (b < a ? b : a);          // C comparison semantics, like the general kernel
}

// ---- column kernel ----
template <class T, int S, bool MAX>
__global__ void __launch_bounds__(128)
minmax_stream_col_kernel(const __grid_constant__ MStreamParams p)
{
    constexpr int C = sizeof(T) >= 4 ? 16 / (int)sizeof(T) : 4;      // 4 x f32, 2 x f64, 4 x u8 / u16
    typedef MPack<T, C> V;
    const int64_t bx = blockIdx.x;
    const int64_t o = bx / p.xblocks;
    const int64_t col = ((bx - o * p.xblocks) * 128 + threadIdx.x) * C;
    if (col >= p.inner) return;
    const int p0 = blockIdx.y * p.seg;
    const int p_end = min(p0 + p.seg, p.n);
    const T* __restrict__ in = reinterpret_cast<const T*>(p.in) + o * (int64_t)p.n * p.inner + col;
    T* __restrict__ out = reinterpret_cast<T*>(p.out) + o * (int64_t)p.n * p.inner + col;
    const T cv = (T)p.cval;
    const int q0 = p0 - p.before;                             // first input row of the segment
    const int n_steps = (p_end - p0) + S - 1;

    auto fetch = [&](int q) -> V {
        V v;
        const int m = mremap_fast(p.mode, q, p.n);
        if (m < 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) v.v[c] = cv;
        } else {
            v = *reinterpret_cast<const V*>(in + (int64_t)m * p.inner);
        }
        return v;
    };
    V pre[S];                                                 // rows of the next S steps in flight
#pragma unroll
    for (int i = 0; i < S; ++i) pre[i] = fetch(q0 + i);
    T acc[S][C];
#pragma unroll
    for (int j = 0; j < S; ++j)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[j][c] = cv;

    for (int base = 0; base < n_steps; base += S) {
        const bool interior = q0 + base >= 0 && q0 + base + 2 * S <= p.n && base + S <= n_steps;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int t = base + s;
            if (!interior && t >= n_steps) break;
            const V v = pre[s];
            if (interior) pre[s] = *reinterpret_cast<const V*>(in + (int64_t)(q0 + t + S) * p.inner);
            else if (t + S < n_steps) pre[s] = fetch(q0 + t + S);
            // logical slot j (the output p0 + t - (S-1) + j) lives in physical slot (j + s) % S
#pragma unroll
            for (int j = 0; j < S - 1; ++j)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[(j + 1 + s) % S][c] = fold<MAX>(acc[(j + 1 + s) % S][c], v.v[c]);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[s % S][c] = v.v[c];           // a new output starts with this row
            if (t >= S - 1) {
                V r;
#pragma unroll
                for (int c = 0; c < C; ++c) r.v[c] = acc[(s + 1) % S][c];
                *reinterpret_cast<V*>(out + (int64_t)(p0 + t - (S - 1)) * p.inner) = r;
            }
        }
    }
}

// ---- row kernel (centred windows: before == S / 2) ----
constexpr int MROW_THREADS = 128;

template <class T, int S, bool MAX>
__global__ void __launch_bounds__(MROW_THREADS)
minmax_stream_row_kernel(const __grid_constant__ MStreamParams p)
{
    constexpr int E = 16 / (int)sizeof(T);                    // elements per 16-byte chunk
    constexpr int NCH = sizeof(T) >= 4 ? 2 : 1;               // output chunks per thread
    constexpr int P = NCH * E;                                // outputs per thread
    constexpr int BEFORE = S / 2, AFTER = S - 1 - BEFORE;
    constexpr int HL = (BEFORE + E - 1) / E, HR = (AFTER + E - 1) / E;
    constexpr int NW = HL + NCH + HR;
    typedef MPack<T, E> V;
    const int64_t gid = (int64_t)blockIdx.x * MROW_THREADS + threadIdx.x;
    const int64_t row = gid / p.gpr;
    if (row >= p.outer) return;
    const int x = (int)(gid - row * p.gpr) * P;
    const T* __restrict__ src = reinterpret_cast<const T*>(p.in) + row * p.n;
    const T cv = (T)p.cval;
    const int g0 = x - HL * E;
    T win[NW * E];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const int g = g0 + j * E;
        if (p.row_aligned && g >= 0 && g + E <= p.n) {
            const V v = *reinterpret_cast<const V*>(src + g);
#pragma unroll
            for (int e = 0; e < E; ++e) win[j * E + e] = v.v[e];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int m = mremap_fast(p.mode, g + e, p.n);
                win[j * E + e] = m < 0 ? cv : src[m];
            }
        }
    }
    T* __restrict__ dst = reinterpret_cast<T*>(p.out) + row * p.n + x;
    const bool vec_st = p.row_aligned && x + P <= p.n;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        V r;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int o = HL * E + k * E + e - BEFORE;        // window of this output starts at win[o]
            T a = win[o];
#pragma unroll
            for (int j = 1; j < S; ++j) a = fold<MAX>(a, win[o + j]);
            r.v[e] = a;
        }
        if (vec_st) {
            *reinterpret_cast<V*>(dst + k * E) = r;
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (x + k * E + e < p.n) dst[k * E + e] = r.v[e];
        }
    }
}

template <class T> constexpr int mcols() { return sizeof(T) >= 4 ? 16 / (int)sizeof(T) : 4; }

template <class T, int S, bool MAX>
cudaError_t launch_ms(MStreamParams& p, cudaStream_t s)
{
    if (p.inner == 1) {
        constexpr int P = (sizeof(T) >= 4 ? 2 : 1) * (16 / (int)sizeof(T));
        p.gpr = (p.n + P - 1) / P;
        const int64_t groups = p.outer * p.gpr;
        minmax_stream_row_kernel<T, S, MAX><<<(unsigned)((groups + MROW_THREADS - 1) / MROW_THREADS), MROW_THREADS, 0, s>>>(p);
        return cudaGetLastError();
    }
    static const int per_sm = [] {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, minmax_stream_col_kernel<T, S, MAX>, 128, 0) != cudaSuccess || n < 1)
            n = 8;
        return n;
    }();
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.xblocks = (int32_t)((p.inner / mcols<T>() + 127) / 128);
    const int64_t slots = (int64_t)sms * per_sm, cols = p.outer * p.xblocks;
    double best = 1e300;
    int best_seg = p.n;
    for (int seg = 8; ; seg += 8) {
        const int sg = seg < p.n ? seg : p.n;
        const int64_t nseg = (p.n + sg - 1) / sg;
        const int64_t waves = (cols * nseg + slots - 1) / slots;
        const double cost = (double)waves * (sg + S + 8);
        if (cost < best) { best = cost; best_seg = sg; }
        if (seg >= p.n || seg >= 4096) break;
    }
    p.seg = best_seg;
    if ((p.n + p.seg - 1) / p.seg > 65535) p.seg = (int32_t)((p.n + 65534) / 65535);
    dim3 grid((unsigned)cols, (unsigned)((p.n + p.seg - 1) / p.seg));
    minmax_stream_col_kernel<T, S, MAX><<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}

template <class T, bool MAX>
cudaError_t launch_ms_size(MStreamParams& p, int S, cudaStream_t s)
{
    switch (S) {
    case 2: return launch_ms<T, 2, MAX>(p, s);
    case 3: return launch_ms<T, 3, MAX>(p, s);
    case 4: return launch_ms<T, 4, MAX>(p, s);
    case 5: return launch_ms<T, 5, MAX>(p, s);
    case 6: return launch_ms<T, 6, MAX>(p, s);
    case 7: return launch_ms<T, 7, MAX>(p, s);
    case 8: return launch_ms<T, 8, MAX>(p, s);
    case 9: return launch_ms<T, 9, MAX>(p, s);
    default: return cudaErrorInvalidValue;
    }
}

int mcols_rt(int dtype) { return dtype == SEPFILT_F64 ? 2 : 4; }

}  // namespace

bool minmax_stream_supported(const ExactTiledGeom& g, int S, int origin, double cval)
{
    if (S < 2 || S > 9 || g.in_dtype != g.out_dtype || g.n_in != g.n_out || g.shift != 0) return false;
    double lo, hi;
    switch (g.in_dtype) {
    case SEPFILT_U8:  lo = 0; hi = 255; break;
    case SEPFILT_I16: lo = -32768; hi = 32767; break;
    case SEPFILT_U16: lo = 0; hi = 65535; break;
    case SEPFILT_F32: case SEPFILT_F64: lo = hi = 0; break;
    default: return false;
    }
    // cval folded to T: integers need an in-range value (truncation toward zero then preserves every
    // comparison outcome after the cast); floats always do (rounding is monotonic); NaN never reaches here
    if (!(cval == cval)) return false;
    if (g.in_dtype != SEPFILT_F32 && g.in_dtype != SEPFILT_F64 && (cval <= lo - 1.0 || cval >= hi + 1.0)) return false;
    if (g.in_dtype == SEPFILT_F32 && (cval > 3.4028234663852886e38 || cval < -3.4028234663852886e38)) return false;
    if (g.outer <= 0 || g.inner <= 0 || g.n_in <= 0 || g.n_in > 1073741824LL /* 2n must fit an int in the boundary fold */) return false;
    const int es = dtype_size(g.in_dtype);
    const uintptr_t a = reinterpret_cast<uintptr_t>(g.in) | reinterpret_cast<uintptr_t>(g.out);
    if (g.inner == 1) {
        if (origin != 0 || (a & (es - 1))) return false;
        if (g.outer * ((g.n_in + 3) / 4) / MROW_THREADS > 2147483647LL) return false;
    } else {
        const int C = mcols_rt(g.in_dtype);
        if (g.inner % C != 0 || (a & (uintptr_t)(es * C - 1))) return false;
        if (g.outer * ((g.inner / C + 127) / 128) > 2147483647LL) return false;
    }
    return true;
}

cudaError_t launch_minmax_stream(const ExactTiledGeom& g, int S, int origin, int mode, double cval, bool is_max,
                                 cudaStream_t s)
{
    MStreamParams p;
    p.in = static_cast<const char*>(g.in);
    p.out = static_cast<char*>(g.out);
    p.outer = g.outer; p.inner = g.inner;
    p.n = (int32_t)g.n_in;
    p.before = S / 2 + origin;
    p.mode = mode;
    p.seg = 0; p.xblocks = 0; p.gpr = 0;
    p.cval = cval;
    const int es = dtype_size(g.in_dtype);
    const uintptr_t a = reinterpret_cast<uintptr_t>(g.in) | reinterpret_cast<uintptr_t>(g.out);
    p.row_aligned = ((a & 15) == 0 && (g.n_in * es) % 16 == 0) ? 1 : 0;
#define DISPATCH(T) (is_max ? launch_ms_size<T, true>(p, S, s) : launch_ms_size<T, false>(p, S, s))
    switch (g.in_dtype) {
    case SEPFILT_U8:  return DISPATCH(uint8_t);
    case SEPFILT_I16: return DISPATCH(int16_t);
    case SEPFILT_U16: return DISPATCH(uint16_t);
    case SEPFILT_F32: return DISPATCH(float);
    case SEPFILT_F64: return DISPATCH(double);
    default: return cudaErrorInvalidValue;
    }
#undef DISPATCH
}

}  // namespace sepfilt
