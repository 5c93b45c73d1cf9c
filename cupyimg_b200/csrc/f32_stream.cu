// f32_stream.cu — register-streaming contiguous-axis float32 pass (float32 in, float32 FMA accumulate, float32
// out): correlate1d / convolve1d / gaussian_filter1d along the last axis of float32 arrays, and the x pass of
// 3-D filters the fused kernels decline (radius > 8 with a z pass, mixed radii, wrap along y / x, cval != 0).
// Same arithmetic contract as f32_1d.cu (rtol 1e-5 of scipy); no shared memory, no CTA barrier:
//   f32_stream_row_kernel  a thread loads the two 16-byte chunks of its 8 outputs plus the halo chunks on either
//       side straight from global memory (its neighbours' chunks: L1 hits) and stores two 16-byte vectors; on
//       aligned rows the out-of-array entries of the threads near a row end are filled by register selects from
//       entries the thread already holds, otherwise element-wise through the boundary rule (_util.py:170-228).
// Strided-axis passes and unaligned geometries run on the shared-memory tiles of f32_1d.cu.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace sepfilt {

namespace {

struct FStreamParams {
    const float* in;
    float*       out;
    int64_t      outer, inner;
    int32_t      n_in, n_out;
    int32_t      shift;               // source index of the filter centre = output position + shift
    int32_t      mode;
    int32_t      seg;                 // col kernel: output rows per segment
    int32_t      xblocks;             // col kernel: CTAs across `inner`
    int32_t      gpr;                 // row kernel: 8-output groups per row
    int32_t      row_aligned;         // row kernel: every row start is 16-byte aligned (in and out)
    float        cval;
    float        w[2 * SEPFILT_FAST_MAX_RADIUS + 1];   // taps at offsets -R..R
};

__device__ __noinline__ int fremap_outside(int mode, int ix, int n) { return remap_index32(mode, ix, n); }
__device__ __forceinline__ int fremap_fast(int mode, int ix, int n)
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
    return fremap_outside(mode, ix, n);
}

// (Column kernels lived here until round 2: per-thread rotating accumulators with a register prefetch ring, 4 or 2
//  columns per thread.  Once the shared-memory tile kernel of f32_1d.cu issued all its staging loads before the
//  first store and filtered with packed FFMA2, the tile won at every radius — 512^3 column pass, 9 / 17 / 33 taps:
//  0.171 / 0.181 / 0.232 ms against 0.19 / 0.251 / 0.59 ms here — and the streaming column kernels were removed.)

// ---- row kernel ----
constexpr int FROW_P = 8;            // outputs per thread
constexpr int FROW_THREADS = 128;

template <int R>
__global__ void __launch_bounds__(FROW_THREADS)
f32_stream_row_kernel(const __grid_constant__ FStreamParams p)
{
    constexpr int H = (R + 3) / 4;                           // halo chunks per side
    constexpr int NW = FROW_P / 4 + 2 * H;                   // chunks in the window
    constexpr int OFF = 4 * H - R;                           // window index of the leftmost tap of output 0
    const int64_t gid = (int64_t)blockIdx.x * FROW_THREADS + threadIdx.x;
    const int64_t row = gid / p.gpr;
    if (row >= p.outer) return;
    const int x = (int)(gid - row * p.gpr) * FROW_P;
    const float* __restrict__ src = p.in + row * p.n_in;
    const int g0 = x + p.shift - 4 * H;                      // source index of win[0]
    const bool vec_ok = p.row_aligned && (p.shift & 3) == 0;
    // Aligned rows (the common case): the halo of the first / last thread of a row is the mirror image of
    // elements the thread already holds, so it is filled by register selects instead of an element-wise
    // gather (which cost a second round of dependent loads in every warp holding an edge thread).
    const bool reg_edges = vec_ok && p.shift == 0 && p.n_in == p.n_out && (p.n_in & 7) == 0 && p.n_in >= 16 * H &&
                           p.mode != SEPFILT_WRAP;
    float win[4 * NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const int g = g0 + 4 * j;
        if (vec_ok && g >= 0 && g + 4 <= p.n_in) {
            const float4 v = *reinterpret_cast<const float4*>(src + g);
            win[4 * j] = v.x; win[4 * j + 1] = v.y; win[4 * j + 2] = v.z; win[4 * j + 3] = v.w;
        } else if (reg_edges) {
            win[4 * j] = win[4 * j + 1] = win[4 * j + 2] = win[4 * j + 3] = p.cval;   // constant mode; others below
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = fremap_fast(p.mode, g + e, p.n_in);
                win[4 * j + e] = m < 0 ? p.cval : src[m];
            }
        }
    }
    if (reg_edges && p.mode != SEPFILT_CONSTANT) {
        const int e = p.mode == SEPFILT_REFLECT ? 1 : 0;     // reflect: d c b a | a b c d ; mirror: d c b | a b c d
        // a thread 8 d columns away from a row end (d = 0 .. (4H-1)/8) holds out-of-array window entries whose
        // sources it also holds: win[i] is array column x + i - 4H
#pragma unroll
        for (int d = 0; 8 * d < 4 * H; ++d) {
            if (x == 8 * d) {                                // entries i < 4H - 8d lie left of column 0
#pragma unroll
                for (int i = 0; i < 4 * H - 8 * d; ++i) {
                    const float refl = win[8 * H - 16 * d - 1 - i], mirr = win[8 * H - 16 * d - i];
                    win[i] = p.mode == SEPFILT_NEAREST ? win[4 * H - 8 * d] : (e ? refl : mirr);
                }
            }
            if (x + FROW_P + 8 * d == p.n_in) {              // entries i >= 4H + 8 + 8d lie right of column n - 1
#pragma unroll
                for (int i = 4 * H + 8 + 8 * d; i < 8 * H + 8; ++i) {
                    const float refl = win[8 * H + 15 + 16 * d - i], mirr = win[8 * H + 14 + 16 * d - i];
                    win[i] = p.mode == SEPFILT_NEAREST ? win[4 * H + 7 + 8 * d] : (e ? refl : mirr);
                }
            }
        }
    }
    // (packed fma.rn.f32x2 on output pairs with re-paired window copies was measured here: 17 taps 0.189 -> 0.225 ms,
    // 33 taps 0.298 -> 0.386 ms on 512^3 — the second copy of the window costs the occupancy the halved issue
    // slots would have bought)
    float acc[FROW_P];
#pragma unroll
    for (int o = 0; o < FROW_P; ++o) acc[o] = win[OFF + o] * p.w[0];
#pragma unroll
    for (int k = 1; k <= 2 * R; ++k) {
        const float w = p.w[k];
#pragma unroll
        for (int o = 0; o < FROW_P; ++o) acc[o] = fmaf(w, win[OFF + o + k], acc[o]);
    }
    float* __restrict__ dst = p.out + row * p.n_out + x;
    if (p.row_aligned && x + FROW_P <= p.n_out) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
        for (int o = 0; o < FROW_P; ++o)
            if (x + o < p.n_out) dst[o] = acc[o];
    }
}

// one instantiation per radius: a tap set is never zero-padded to a wider bucket (0 * NaN / 0 * Inf would
// spread a non-finite input sample beyond the true footprint of the filter, unlike scipy)
int fstream_bucket(int r)
{
    return (r >= 1 && r <= 16) ? r : -1;
}

template <int R>
cudaError_t launch_fstream(FStreamParams& p, const F32Taps& t, cudaStream_t s)
{
    for (int k = 0; k <= 2 * SEPFILT_FAST_MAX_RADIUS; ++k) p.w[k] = 0.f;
    for (int k = 0; k <= 2 * t.radius; ++k) p.w[k + R - t.radius] = t.w[k];
    if (p.inner == 1) {
        const int64_t groups = p.outer * p.gpr;
        f32_stream_row_kernel<R><<<(unsigned)((groups + FROW_THREADS - 1) / FROW_THREADS), FROW_THREADS, 0, s>>>(p);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;       // column passes run on the tile kernel (f32_1d.cu)
}

}  // namespace

bool f32_stream_supported(const F32Line& g, int radius)
{
    const int R = fstream_bucket(radius);
    if (R < 0) return false;
    // Measured on 512^3 (tools/time_f32_1d.py, ms, streaming vs shared-memory tile): row pass 5 / 9 / 17 taps
    // 0.156 / 0.160 / 0.189 vs 0.245 / 0.253 / 0.296 (87-105 % of the measured copy bandwidth; with an
    // element-wise halo gather in the edge threads the 17-tap row pass took 0.352), column pass 0.170 / 0.19 /
    // 0.252 vs 0.23 / 0.24 / 0.27; 25 / 33 taps 0.34-0.67 vs 0.30-0.35: radius 12 / 16 stay on the tiles.
    if (g.n_in <= 0 || g.n_out <= 0 || g.outer <= 0 || g.inner <= 0) return false;
    // contiguous-axis passes only: the row kernel wins over the shared-memory tile at every radius (33 taps on 512^3:
    // 0.299 vs 0.356 ms); strided-axis passes run on the tile kernel of f32_1d.cu
    if (g.inner != 1) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(g.in) | reinterpret_cast<uintptr_t>(g.out);
    if (a & 3) return false;
    const int64_t groups = g.outer * ((g.n_out + FROW_P - 1) / FROW_P);
    if ((groups + FROW_THREADS - 1) / FROW_THREADS > 2147483647LL) return false;
    return true;
}

cudaError_t launch_f32_stream(const F32Line& g, const F32Taps& t, cudaStream_t s)
{
    FStreamParams p;
    p.in = g.in; p.out = g.out;
    p.outer = g.outer; p.inner = g.inner;
    p.n_in = g.n_in; p.n_out = g.n_out;
    p.shift = g.in_offset;
    p.mode = g.mode;
    p.cval = g.cval;
    p.seg = 0; p.xblocks = 0; p.gpr = 0; p.row_aligned = 0;
    const int R = fstream_bucket(t.radius);
    p.gpr = (g.n_out + FROW_P - 1) / FROW_P;
    const uintptr_t a = reinterpret_cast<uintptr_t>(g.in) | reinterpret_cast<uintptr_t>(g.out);
    p.row_aligned = ((a & 15) == 0 && g.n_in % 4 == 0 && g.n_out % 4 == 0) ? 1 : 0;
    switch (R) {
    case 1: return launch_fstream<1>(p, t, s);
    case 2: return launch_fstream<2>(p, t, s);
    case 3: return launch_fstream<3>(p, t, s);
    case 4: return launch_fstream<4>(p, t, s);
    case 5: return launch_fstream<5>(p, t, s);
    case 6: return launch_fstream<6>(p, t, s);
    case 7: return launch_fstream<7>(p, t, s);
    case 8: return launch_fstream<8>(p, t, s);
    case 9: return launch_fstream<9>(p, t, s);
    case 10: return launch_fstream<10>(p, t, s);
    case 11: return launch_fstream<11>(p, t, s);
    case 12: return launch_fstream<12>(p, t, s);
    case 13: return launch_fstream<13>(p, t, s);
    case 14: return launch_fstream<14>(p, t, s);
    case 15: return launch_fstream<15>(p, t, s);
    case 16: return launch_fstream<16>(p, t, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace sepfilt
