// exact_stream.cu — register-streaming version of the exact (scipy-arithmetic) 1-D pass for the
// dtype-preserving integer / float workloads (C3: convolve1d on uint16 image stacks).
//
// Same bits as exact.cu / exact_tiled.cu / scipy's NI_Correlate1D symmetric branches (SURVEY.md
// App. C.2):   acc = x[c] * w[0];  for d = R .. 1:  acc += (x[c-d] +/- x[c+d]) * w[d]
// in float64 with __dmul_rn / __dadd_rn (never contracted), then the C-cast store.  What changes
// is the data movement: exact_tiled.cu stages a float64 tile in shared memory with scalar loads
// (115 thread-instructions per pixel, FP64 pipe 12-16 % busy — profiles/r1 C3 capture); here the
// FP64 pipe is the intended limit (3R + 1 FP64 ops per output at 64 / clk / SM):
//   exact_stream_col_kernel  axis with inner > 1: each thread owns C adjacent columns (one 4..16
//       byte vector load per row), marches a segment of the filtered axis and keeps the 2R+1 row
//       window as float64 in REGISTERS (rotating slots, fully unrolled: no shared memory, no
//       register moves).  Every element is loaded once per segment and converted once; the rows
//       of the next 2R+1 steps are in flight while the current ones are filtered.
//   exact_stream_row_kernel  contiguous axis: a thread loads the 16-byte chunks of its 16 outputs plus
//       the halo chunks on either side straight from global memory (its neighbours' chunks: L1 hits),
//       converts the 16 + 2R window once, and stores 16-byte vectors.  On aligned rows the halo of a
//       row's first / last thread is filled by register selects from elements the thread already
//       holds; otherwise element-wise through the boundary rule (_util.py:170-228).
// Neither kernel uses shared memory or a CTA barrier.
// Out-of-range float64 -> integer casts follow x86 cvttsd2si ("integer indefinite"), detected
// with an integer compare on the exponent instead of FP64 compares.
#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

namespace {

constexpr int SMAXR = 8;   // largest radius bucket of the streaming kernels

struct StreamParams {
    const char* in;
    char*       out;
    int64_t     outer, inner;        // (outer, n, inner) view, extents in elements
    int32_t     n_in, n_out;
    int32_t     shift;               // source index of the filter centre = output position + shift
    int32_t     mode;
    int32_t     seg;                 // col kernel: output rows per segment
    int32_t     xblocks;             // col kernel: CTAs across `inner`; row kernel: CTAs across n_out
    int32_t     tpr;                 // row kernel: threads (16-output groups) per row
    int32_t     row_aligned;         // row kernel: every row start is 16-byte aligned
    double      cval;
    double      w[SMAXR + 1];        // w[d] = tap at distance d from the centre (fw[-d]); w[0] centre
};

template <class T, int C> struct alignas(sizeof(T) * C) Pack { T v[C]; };

// ---- float64 -> T under scipy's C casts, |v| >= 2^31 / NaN -> INT32_MIN ----
__device__ __forceinline__ int32_t cvt_x86_i32_fast(double v)
{
    const int hi = __double2hiint(v);
    return ((hi & 0x7ff00000) >= 0x41e00000) ? INT32_MIN : __double2int_rz(v);
}
template <class T> __device__ __forceinline__ T cast_fast(double v) { return (T)cvt_x86_i32_fast(v); }
template <> __device__ __forceinline__ float  cast_fast<float>(double v)  { return __double2float_rn(v); }
template <> __device__ __forceinline__ double cast_fast<double>(double v) { return v; }

// boundary remap of an index that left the array: rare (tile edges only) and full of integer
// divisions, so it is kept out of line — inlined into the unrolled row loops it multiplied the
// kernel's instruction count
__device__ __noinline__ int remap_outside(int mode, int ix, int n) { return remap_index32(mode, ix, n); }
// in range: identity.  Less than one array length outside (the case of every halo when n > R): one
// fold without divisions, cheap enough to inline.  Further out (multi-reflection, n == 1): the call.
__device__ __forceinline__ int remap_fast(int mode, int ix, int n)
{
    if ((unsigned)ix < (unsigned)n) return ix;
    if (mode == SEPFILT_CONSTANT) return -1;
    if (mode == SEPFILT_NEAREST) return ix < 0 ? 0 : n - 1;
    if (ix > -n && ix < 2 * n - 1) {
        const bool low = ix < 0;
        if (mode == SEPFILT_REFLECT) return low ? -1 - ix : 2 * n - 1 - ix;
        if (mode == SEPFILT_MIRROR) return low ? -ix : 2 * n - 2 - ix;
        return low ? ix + n : ix - n;                        // wrap
    }
    return remap_outside(mode, ix, n);
}

template <int R, int SGN>
__device__ __forceinline__ double pair_term(double a, double b)
{
    return SGN > 0 ? __dadd_rn(a, b) : __dsub_rn(a, b);
}

// =====================================================================================
// column kernel
// =====================================================================================
// registers: the float64 window (2 per cell) + two raw row blocks + ~28 for addressing / accumulation
// Measured on B200, 8 x 2048 x 2048, 9 taps (tools/time_c3.py, build-time variants of these two macros):
// uint16 column pass 0.061 ms (2 columns, no register prefetch: 96 registers, 5 CTAs / SM), 0.061 (2, prefetch),
// 0.058 (4 columns, none: 168 registers, 3 CTAs / SM), 0.062 (4, prefetch) — within 5 % of each other, the
// FP64 pipe sits at 47 % in all of them; the register double buffer of the next block's rows pays where the
// pass is HBM-bound (float64: 0.090 vs 0.113 ms, 91 % of the measured copy bandwidth).
#ifndef STREAM_COLS
#define STREAM_COLS 2
#endif
#ifndef STREAM_PREFETCH
#define STREAM_PREFETCH (sizeof(T) >= 4 ? 1 : 0)
#endif
template <class T, int R> struct ColGeom {
    static constexpr int C = (R <= 4 && sizeof(T) <= 2) ? STREAM_COLS : 2;
    static constexpr int W = 2 * R + 1;
    static constexpr int RAW = (C * (int)sizeof(T) + 3) / 4;
    static constexpr int PF = STREAM_PREFETCH;
    static constexpr int REGS = 2 * W * C + (1 + PF) * W * RAW + 40 + (R >= 8 ? 16 : 0);
    #ifdef STREAM_FORCE_CTAS
    static constexpr int CTAS = STREAM_FORCE_CTAS;
#else
    static constexpr int CTAS = REGS <= 84 ? 6 : (REGS <= 100 ? 5 : (REGS <= 128 ? 4 : (REGS <= 168 ? 3 : 2)));
#endif     // 128-thread CTAs per SM
};

template <class T, int R, int SGN>
__global__ void __launch_bounds__(128, (ColGeom<T, R>::CTAS))
exact_stream_col_kernel(const __grid_constant__ StreamParams p)
{
    constexpr int C = ColGeom<T, R>::C;
    constexpr int W = ColGeom<T, R>::W;
    constexpr bool PF = ColGeom<T, R>::PF != 0;
    typedef Pack<T, C> P;
    const int64_t bx = blockIdx.x;
    const int64_t o = bx / p.xblocks;
    const int64_t col = ((bx - o * p.xblocks) * 128 + threadIdx.x) * C;
    if (col >= p.inner) return;
    const int p0 = blockIdx.y * p.seg;
    const int p_end = min(p0 + p.seg, p.n_out);
    const T* __restrict__ in = reinterpret_cast<const T*>(p.in) + o * (int64_t)p.n_in * p.inner + col;
    T* __restrict__ out = reinterpret_cast<T*>(p.out) + o * (int64_t)p.n_out * p.inner + col;

    // raw row fetch; rows outside the array are remapped (uniform over the CTA) or flagged for cval
    auto fetch = [&](int q, P& raw) -> bool {
        const int m = remap_fast(p.mode, q, p.n_in);
        if (m < 0) { raw = P{}; return false; }
        raw = *reinterpret_cast<const P*>(in + (int64_t)m * p.inner);
        return true;
    };
    double win[W][C];
    {
        P raw[2 * R];
        unsigned ok = 0;
#pragma unroll
        for (int j = 0; j < 2 * R; ++j) ok |= (fetch(p0 + p.shift - R + j, raw[j]) ? 1u : 0u) << j;
#pragma unroll
        for (int j = 0; j < 2 * R; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) win[j][c] = ((ok >> j) & 1u) ? (double)raw[j].v[c] : p.cval;
    }
    P cur[W];
    unsigned cur_ok = 0;
#pragma unroll
    for (int s = 0; s < W; ++s)
        if (p0 + s < p_end) cur_ok |= (fetch(p0 + s + p.shift + R, cur[s]) ? 1u : 0u) << s;

    // one block of W outputs: convert the block's rows into their rotating window slots, filter, store.
    // INTERIOR blocks (every row of this and of the next block inside the array, full block) carry no
    // remap, no cval selects and no bounds tests.
    auto filter_step = [&](int s, int q) {
        P res;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double acc = __dmul_rn(win[(s + R) % W][c], p.w[0]);
#pragma unroll
            for (int d = R; d >= 1; --d)
                acc = __dadd_rn(acc, __dmul_rn(pair_term<R, SGN>(win[(s + R - d) % W][c], win[(s + R + d) % W][c]), p.w[d]));
            res.v[c] = cast_fast<T>(acc);
        }
        *reinterpret_cast<P*>(out + (int64_t)q * p.inner) = res;
    };
    for (int base = p0; base < p_end; base += W) {
        P nxt[W];
        const int first_next = base + W + p.shift + R;       // first input row of the next block
        if (cur_ok == (1u << W) - 1u && base + 2 * W <= p_end && first_next >= 0 && first_next + W <= p.n_in) {
            const T* __restrict__ src = in + (int64_t)first_next * p.inner;
            // (an additional prefetch.global.L2 of the rows four blocks ahead was measured: no gain for the
            //  FP64-bound types, 0.090 -> 0.106 ms for float64)
            if constexpr (PF) {
#pragma unroll
                for (int s = 0; s < W; ++s) nxt[s] = *reinterpret_cast<const P*>(src + (int64_t)s * p.inner);
            }
#pragma unroll
            for (int s = 0; s < W; ++s) {
#pragma unroll
                for (int c = 0; c < C; ++c) win[(s + 2 * R) % W][c] = (double)cur[s].v[c];
                filter_step(s, base + s);
            }
            if constexpr (PF) {
#pragma unroll
                for (int s = 0; s < W; ++s) cur[s] = nxt[s];
            } else {
                // no register double buffer: the next block's rows are requested now and awaited at the
                // top of the next iteration; the other resident warps cover the latency
#pragma unroll
                for (int s = 0; s < W; ++s) cur[s] = *reinterpret_cast<const P*>(src + (int64_t)s * p.inner);
            }
            continue;
        }
        // edge block: rows of the NEXT block still in flight while this block is filtered
        unsigned nxt_ok = 0;
#pragma unroll
        for (int s = 0; s < W; ++s)
            if (base + W + s < p_end) nxt_ok |= (fetch(base + W + s + p.shift + R, nxt[s]) ? 1u : 0u) << s;
#pragma unroll
        for (int s = 0; s < W; ++s) {
            const int q = base + s;
            if (q < p_end) {
                // newest row of the window of output q lands in slot (s + 2R) % W
#pragma unroll
                for (int c = 0; c < C; ++c)
                    win[(s + 2 * R) % W][c] = ((cur_ok >> s) & 1u) ? (double)cur[s].v[c] : p.cval;
                filter_step(s, q);
            }
        }
#pragma unroll
        for (int s = 0; s < W; ++s) cur[s] = nxt[s];
        cur_ok = nxt_ok;
    }
}

// =====================================================================================
// row kernel
// =====================================================================================
constexpr int ROW_P = 16;          // outputs per thread
constexpr int ROW_THREADS = 128;

template <class T, int R> struct RowGeom {
    static constexpr int E = 16 / sizeof(T);                // elements per 16-byte chunk
    static constexpr int H = (R + E - 1) / E;               // halo chunks per side
    static constexpr int NCH = ROW_P / E;                   // output chunks per thread
};

// No shared memory and no barrier: a thread loads the NCH chunks of its 16 outputs plus H halo chunks
// on either side straight from global memory (the halo chunks are its neighbours' own chunks: L1
// hits), converts the 16 + 2HE window once and filters it.  Chunks that are not fully inside the row
// (row ends, unaligned rows, origins that break chunk alignment) are gathered element by element with
// the boundary rule; that path is taken by the two edge threads of a row only.
// Measured and rejected (8 x 2048 x 2048 uint16, 9 taps): a shared-memory staged tile with a CTA
// barrier (0.082 ms against 0.051 ms) and a persistent grid that walks down the rows with the next
// row's chunks prefetched into registers (0.059 ms: 128 registers, a third of the warps).
template <class T, int R, int SGN>
__global__ void __launch_bounds__(ROW_THREADS)
exact_stream_row_kernel(const __grid_constant__ StreamParams p)
{
    typedef RowGeom<T, R> G;
    constexpr int E = G::E, H = G::H, NCH = G::NCH;
    typedef Pack<T, E> P;
    const int64_t gid = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
    const int64_t row = gid / p.tpr;                         // tpr = 16-output groups per row
    if (row >= p.outer) return;
    const int x = (int)(gid - row * p.tpr) * ROW_P;
    const T* __restrict__ src = reinterpret_cast<const T*>(p.in) + row * p.n_in;
    const int g0 = x + p.shift - H * E;                      // source index of win[0]
    const bool vec_ok = p.row_aligned && (p.shift % E) == 0;
    // Aligned rows: the halo of the first / last thread of a row mirrors elements the thread already holds,
    // so it is filled by register selects after the conversion instead of an element-wise gather (a second
    // round of dependent loads in every warp that holds an edge thread).
    constexpr int HE = H * E;
    const bool reg_edges = vec_ok && p.shift == 0 && p.n_in == p.n_out && (p.n_in % ROW_P) == 0 &&
                           p.n_in >= 2 * ROW_P && p.mode != SEPFILT_WRAP && HE <= ROW_P;

    double win[(NCH + 2 * H) * E];
#pragma unroll
    for (int j = 0; j < NCH + 2 * H; ++j) {
        const int g = g0 + j * E;
        if (vec_ok && g >= 0 && g + E <= p.n_in) {
            const P v = *reinterpret_cast<const P*>(src + g);
#pragma unroll
            for (int e = 0; e < E; ++e) win[j * E + e] = (double)v.v[e];
        } else if (reg_edges) {
#pragma unroll
            for (int e = 0; e < E; ++e) win[j * E + e] = p.cval;      // constant mode; the others below
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int m = remap_fast(p.mode, g + e, p.n_in);
                win[j * E + e] = m < 0 ? p.cval : (double)src[m];
            }
        }
    }
    if (reg_edges && p.mode != SEPFILT_CONSTANT) {
        const bool refl = p.mode == SEPFILT_REFLECT, near = p.mode == SEPFILT_NEAREST;
        if (x == 0) {                                        // win[i], i < HE, is array column i - HE
#pragma unroll
            for (int i = 0; i < HE; ++i) {
                const double a = win[2 * HE - 1 - i], b = win[2 * HE - i];
                win[i] = near ? win[HE] : (refl ? a : b);
            }
        }
        if (x + ROW_P == p.n_in) {                           // win[i], i >= HE + ROW_P, is array column n + (i - HE - ROW_P)
#pragma unroll
            for (int i = HE + ROW_P; i < 2 * HE + ROW_P; ++i) {
                const double a = win[2 * (HE + ROW_P) - 1 - i], b = win[2 * (HE + ROW_P) - 2 - i];
                win[i] = near ? win[HE + ROW_P - 1] : (refl ? a : b);
            }
        }
    }
    constexpr int OFF = H * E - R;                           // win index of the leftmost tap of output 0
    T* __restrict__ dst = reinterpret_cast<T*>(p.out) + row * p.n_out + x;
    const bool vec_st = p.row_aligned && x + ROW_P <= p.n_out;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        P res;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int o = k * E + e;
            double acc = __dmul_rn(win[OFF + o + R], p.w[0]);
#pragma unroll
            for (int d = R; d >= 1; --d)
                acc = __dadd_rn(acc, __dmul_rn(pair_term<R, SGN>(win[OFF + o + R - d], win[OFF + o + R + d]), p.w[d]));
            res.v[e] = cast_fast<T>(acc);
        }
        if (vec_st) {
            *reinterpret_cast<P*>(dst + k * E) = res;
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (x + k * E + e < p.n_out) dst[k * E + e] = res.v[e];
        }
    }
}

int stream_bucket(int r)
{
    static const int buckets[] = {1, 2, 4, 8};
    for (int b : buckets) if (r <= b) return b;
    return -1;
}

// segment length of the column kernel: whole waves of the resident CTA slots with the least halo
// re-reading (a segment re-reads 2R rows and pays a fixed prologue)
void plan_segments(StreamParams& p, int R, int64_t slots)
{
    const int W = 2 * R + 1;
    const int64_t cols = p.outer * p.xblocks;
    double best = 1e300;
    int best_seg = p.n_out;
    for (int m = 1; m <= 256; ++m) {
        const int seg = W * m;
        const int64_t nseg = (p.n_out + seg - 1) / seg;
        const int64_t waves = (cols * nseg + slots - 1) / slots;
        const double cost = (double)waves * (seg + 2 * R + 6);
        if (cost < best) { best = cost; best_seg = seg; }
        if (seg >= p.n_out) break;
    }
    p.seg = best_seg;
    if ((p.n_out + p.seg - 1) / p.seg > 65535) p.seg = (int32_t)((p.n_out + 65534) / 65535);
}

template <class T, int R, int SGN>
cudaError_t launch_stream(StreamParams& p, cudaStream_t s)
{
    if (p.inner != 1) {
        static const int per_sm = [] {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, exact_stream_col_kernel<T, R, SGN>, 128, 0) != cudaSuccess || n < 1)
                n = ColGeom<T, R>::CTAS;
            return n;
        }();
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        plan_segments(p, R, (int64_t)sms * per_sm);
    }
    if (p.inner == 1) {
        const int64_t groups = p.outer * p.tpr;
        exact_stream_row_kernel<T, R, SGN><<<(unsigned)((groups + ROW_THREADS - 1) / ROW_THREADS), ROW_THREADS, 0, s>>>(p);
    } else {
        const unsigned nseg = (unsigned)((p.n_out + p.seg - 1) / p.seg);
        dim3 grid((unsigned)(p.outer * p.xblocks), nseg);
        exact_stream_col_kernel<T, R, SGN><<<grid, 128, 0, s>>>(p);
    }
    return cudaGetLastError();
}

template <class T>
cudaError_t launch_stream_t(StreamParams& p, int R, int sgn, cudaStream_t s)
{
    switch (R * 2 + (sgn > 0 ? 1 : 0)) {
#define CASE(RR) case RR * 2 + 1: return launch_stream<T, RR, 1>(p, s); case RR * 2: return launch_stream<T, RR, -1>(p, s);
        CASE(1) CASE(2) CASE(4) CASE(8)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

int cols_per_thread(int R, int esize) { return (R <= 4 && esize <= 2) ? STREAM_COLS : 2; }

}  // namespace

bool exact_stream_supported(const ExactTiledGeom& g, int K, int symmetric)
{
    if (!(K & 1) || (symmetric != 1 && symmetric != -1)) return false;
    const int R = stream_bucket(K / 2);
    if (R < 0) return false;
    if (g.in_dtype != g.out_dtype) return false;
    // zero padding to the bucket is bit-neutral for finite data only: 0 * NaN / 0 * Inf would poison outputs
    // beyond the true footprint, so float arrays take this kernel only when the radius IS the bucket
    if ((g.in_dtype == SEPFILT_F32 || g.in_dtype == SEPFILT_F64) && R != K / 2) return false;
    switch (g.in_dtype) {
    case SEPFILT_U8: case SEPFILT_I16: case SEPFILT_U16: case SEPFILT_F32: case SEPFILT_F64: break;
    default: return false;
    }
    if (g.outer <= 0 || g.inner <= 0 || g.n_in <= 0 || g.n_out <= 0) return false;
    if (g.n_in > 1073741824LL /* 2n must fit an int in the boundary fold */ || g.n_out > 1073741824LL /* 2n must fit an int in the boundary fold */) return false;
    if (g.shift > 1073741824LL || g.shift < -1073741824LL) return false;
    const int es = dtype_size(g.in_dtype);
    const uintptr_t a = reinterpret_cast<uintptr_t>(g.in) | reinterpret_cast<uintptr_t>(g.out);
    if (g.inner == 1) {
        if (a & (es - 1)) return false;
        const int64_t groups = g.outer * ((g.n_out + ROW_P - 1) / ROW_P);
        if ((groups + ROW_THREADS - 1) / ROW_THREADS > 2147483647LL) return false;
    } else {
        const int C = cols_per_thread(R, es);
        if (g.inner % C != 0 || (a & (uintptr_t)(es * C - 1))) return false;
        const int64_t xb = (g.inner / C + 127) / 128;
        if (g.outer * xb > 2147483647LL) return false;
    }
    return true;
}

cudaError_t launch_exact_stream(const ExactTiledGeom& g, const double* taps, int K, int symmetric, int mode,
                                double cval, cudaStream_t s)
{
    StreamParams p;
    p.in = static_cast<const char*>(g.in);
    p.out = static_cast<char*>(g.out);
    p.outer = g.outer;
    p.inner = g.inner;
    p.n_in = (int32_t)g.n_in;
    p.n_out = (int32_t)g.n_out;
    p.shift = (int32_t)g.shift;
    p.mode = mode;
    p.cval = cval;
    const int r = K / 2;
    const int R = stream_bucket(r);
    for (int d = 0; d <= SMAXR; ++d) p.w[d] = d <= r ? taps[r - d] : 0.0;   // fw[-d]; zero padding is bit-neutral
    const int es = dtype_size(g.in_dtype);
    p.seg = 0; p.tpr = 0; p.row_aligned = 0;
    if (g.inner == 1) {
        p.tpr = (int32_t)((g.n_out + ROW_P - 1) / ROW_P);
        p.xblocks = 0;
        const uintptr_t a = reinterpret_cast<uintptr_t>(g.in) | reinterpret_cast<uintptr_t>(g.out);
        p.row_aligned = ((a & 15) == 0 && (g.n_in * es) % 16 == 0 && (g.n_out * es) % 16 == 0) ? 1 : 0;
    } else {
        const int C = cols_per_thread(R, es);
        p.xblocks = (int32_t)((g.inner / C + 127) / 128);
    }
    switch (g.in_dtype) {
    case SEPFILT_U8:  return launch_stream_t<uint8_t>(p, R, symmetric, s);
    case SEPFILT_I16: return launch_stream_t<int16_t>(p, R, symmetric, s);
    case SEPFILT_U16: return launch_stream_t<uint16_t>(p, R, symmetric, s);
    case SEPFILT_F32: return launch_stream_t<float>(p, R, symmetric, s);
    case SEPFILT_F64: return launch_stream_t<double>(p, R, symmetric, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace sepfilt
