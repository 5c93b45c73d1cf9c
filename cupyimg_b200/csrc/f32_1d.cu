// f32_1d.cu — tiled single-axis f32 correlation passes (float32 in, float32 FMA
// accumulate, float32 out) for a C-contiguous array viewed as (outer, n, inner).
//
//   corr1d_f32_row_kernel  inner == 1: the filtered axis is the contiguous one.
//       16-byte coalesced loads of a row segment + halo into shared memory, each lane
//       keeps a register window of 4 + 2R values and produces 4 adjacent outputs.
//   corr1d_f32_col_kernel  inner  > 1: the filtered axis is strided.
//       (TN + 2R) x TI tile with halo ROWS staged in shared memory, each lane owns a
//       float4 of the contiguous inner axis and marches down the filtered axis producing
//       RN outputs from RN + 2R shared-memory rows.
//
// In both, boundary remapping (_util.py:170-228) is resolved while staging the tile —
// only for cells that fall outside the array — never per tap.  Taps live in kernel
// parameters (constant bank), zero-padded to the compile-time radius bucket R.
// These replace one `_call_kernel` launch (_filters_core.py:152) each; the fused
// multi-axis kernel in fused3d.cu replaces a whole per-axis loop.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace sepfilt {

__host__ __device__ constexpr int roundup4(int r) { return (r + 3) & ~3; }

// =============================== contiguous axis ===============================
constexpr int ROW_W = 256;      // outputs per tile row (2 warps x 32 lanes x 4)
constexpr int ROW_ROWS = 4;     // lines per CTA

template <int R>
__global__ void __launch_bounds__(256)
corr1d_f32_row_kernel(const __grid_constant__ F32Line g, const __grid_constant__ F32Taps t,
                      const int vec_in, const int vec_out)
{
    constexpr int HL = roundup4(R);
    constexpr int PITCH = ROW_W + 2 * HL;
    constexpr int NV = 2 * HL / 4 + 1;
    __shared__ __align__(16) float tile[ROW_ROWS][PITCH];

    const int64_t row0 = (int64_t)blockIdx.x * ROW_ROWS;
    const int x0 = blockIdx.y * ROW_W;
    const int tid = threadIdx.x;

    // ---- stage: tile column s <-> source coordinate x0 - HL + s + in_offset ----
    const int src0 = x0 - HL + g.in_offset;
    if (vec_in) {
        constexpr int V_PER_ROW = PITCH / 4;
        for (int i = tid; i < ROW_ROWS * V_PER_ROW; i += 256) {
            const int r = i / V_PER_ROW, v = i - r * V_PER_ROW;
            const int64_t row = row0 + r;
            if (row >= g.outer) continue;
            const float* line = g.in + row * g.n_in;
            const int sx = src0 + 4 * v;
            float4 val;
            if (sx >= 0 && sx + 3 < g.n_in) {
                val = __ldg(reinterpret_cast<const float4*>(line + sx));
            } else {
                float e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int m = remap_index32(g.mode, sx + q, g.n_in);
                    e[q] = m < 0 ? g.cval : __ldg(line + m);
                }
                val = make_float4(e[0], e[1], e[2], e[3]);
            }
            *reinterpret_cast<float4*>(&tile[r][4 * v]) = val;
        }
    } else {
        for (int i = tid; i < ROW_ROWS * PITCH; i += 256) {
            const int r = i / PITCH, s = i - r * PITCH;
            const int64_t row = row0 + r;
            if (row >= g.outer) continue;
            const float* line = g.in + row * g.n_in;
            const int m = remap_index32(g.mode, src0 + s, g.n_in);
            tile[r][s] = m < 0 ? g.cval : __ldg(line + m);
        }
    }
    __syncthreads();

    // ---- compute: lane -> 4 adjacent outputs, register window of 4*NV values ----
    const int r = tid >> 6;                 // 64 threads per row
    const int c = (tid & 63) * 4;
    const int64_t row = row0 + r;
    const int x = x0 + c;
    if (row >= g.outer || x >= g.n_out) return;
    float win[4 * NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(&tile[r][c + 4 * i]);
        win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) {
        const float w = t.w[k];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(w, win[o + HL - R + k], acc[o]);
    }
    float* dst = g.out + row * g.n_out + x;
    if (vec_out && x + 3 < g.n_out) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (x + o < g.n_out) dst[o] = acc[o];
    }
}

// ================================ strided axis =================================
constexpr int COL_TI = 64;      // contiguous inner elements per tile (16 lanes x float4)
constexpr int COL_RN = 8;       // outputs per lane along the filtered axis
constexpr int COL_TN = 16 * COL_RN;   // 128 output rows per tile

template <int R>
__global__ void __launch_bounds__(256)
corr1d_f32_col_kernel(const __grid_constant__ F32Line g, const __grid_constant__ F32Taps t,
                      const int vec, const int n_itiles)
{
    constexpr int ROWS = COL_TN + 2 * R;
    __shared__ __align__(16) float tile[ROWS][COL_TI];

    const int64_t bx = blockIdx.x;
    const int64_t o = bx / n_itiles;
    const int64_t i0 = (bx - o * n_itiles) * COL_TI;
    const int n0 = blockIdx.y * COL_TN;
    const int tid = threadIdx.x;
    const float* base_in = g.in + o * (int64_t)g.n_in * g.inner;
    float* base_out = g.out + o * (int64_t)g.n_out * g.inner;

    // ---- stage: tile row e <-> source row n0 - R + e + in_offset (remapped once per row) ----
    const int src0 = n0 - R + g.in_offset;
    const int rows_needed = min(ROWS, g.n_out - n0 + 2 * R);
    if (vec) {
        const int lane16 = tid & 15;
        const int64_t ii = i0 + 4 * lane16;
        // all loads of a thread are issued before its first store: the staging phase costs one memory round
        // trip instead of one per 16 rows (the rolled loop exposed ROWS / 16 dependent latencies per CTA)
        constexpr int NIT = (ROWS + 15) / 16;
        float4 v[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = (tid >> 4) + 16 * it;
            v[it] = make_float4(g.cval, g.cval, g.cval, g.cval);
            if (e < rows_needed) {
                const int m = remap_index32(g.mode, src0 + e, g.n_in);
                if (m >= 0 && ii < g.inner)
                    v[it] = __ldg(reinterpret_cast<const float4*>(base_in + (int64_t)m * g.inner + ii));
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = (tid >> 4) + 16 * it;
            if (e < rows_needed) *reinterpret_cast<float4*>(&tile[e][4 * lane16]) = v[it];
        }
    } else {
        const int lane64 = tid & 63;
        const int64_t ii = i0 + lane64;
        for (int e = tid >> 6; e < rows_needed; e += 4) {
            const int m = remap_index32(g.mode, src0 + e, g.n_in);
            float v = g.cval;
            if (m >= 0 && ii < g.inner) v = __ldg(base_in + (int64_t)m * g.inner + ii);
            tile[e][lane64] = v;
        }
    }
    __syncthreads();

    // ---- compute: lane -> float4 of inner x RN outputs, marching down the axis ----
    const int ti = tid & 15, tn = tid >> 4;
    const int64_t ii = i0 + 4 * ti;
    const int p0 = n0 + tn * COL_RN;
    if (ii >= g.inner || p0 >= g.n_out) return;
    // packed fma.rn.f32x2 on the two column pairs of the lane's float4, scalar-broadcast tap: half the issue
    // slots of four scalar FFMA per row and tap (same products, same summation order)
    ptx::u64 acc2[COL_RN][2];
#pragma unroll
    for (int q = 0; q < COL_RN; ++q) acc2[q][0] = acc2[q][1] = 0ull;
#pragma unroll
    for (int j = 0; j < COL_RN + 2 * R; ++j) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&tile[tn * COL_RN + j][4 * ti]);
#pragma unroll
        for (int q = 0; q < COL_RN; ++q) {
            const int k = j - q;
            if (k >= 0 && k <= 2 * R) {
                acc2[q][0] = ptx::fma2s(v.x, t.w[k], acc2[q][0]);
                acc2[q][1] = ptx::fma2s(v.y, t.w[k], acc2[q][1]);
            }
        }
    }
    float4 acc[COL_RN];
#pragma unroll
    for (int q = 0; q < COL_RN; ++q) {
        ptx::unpack2(acc2[q][0], acc[q].x, acc[q].y);
        ptx::unpack2(acc2[q][1], acc[q].z, acc[q].w);
    }
#pragma unroll
    for (int q = 0; q < COL_RN; ++q) {
        const int p = p0 + q;
        if (p >= g.n_out) break;
        float* dst = base_out + (int64_t)p * g.inner + ii;
        if (vec) {
            *reinterpret_cast<float4*>(dst) = acc[q];
        } else {
            const float e[4] = {acc[q].x, acc[q].y, acc[q].z, acc[q].w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ii + u < g.inner) dst[u] = e[u];
        }
    }
}

// =================================== host side ===================================
static int radius_bucket(int r)
{
    // one instantiation per radius: no zero padding (0 * NaN would spread non-finite samples, unlike scipy)
    return (r >= 1 && r <= 16) ? r : -1;
}

bool f32_line_supported(const F32Line& g, int radius)
{
    if (radius_bucket(radius) < 0) return false;
    if (g.n_in <= 0 || g.n_out <= 0 || g.outer <= 0 || g.inner <= 0) return false;
    if (g.inner == 1) {
        if ((g.outer + ROW_ROWS - 1) / ROW_ROWS > 2147483647LL) return false;
        if ((g.n_out + ROW_W - 1) / ROW_W > 65535) return false;
    } else {
        const int64_t n_itiles = (g.inner + COL_TI - 1) / COL_TI;
        if (n_itiles > 2147483647LL || g.outer * n_itiles > 2147483647LL) return false;
        if ((g.n_out + COL_TN - 1) / COL_TN > 65535) return false;
    }
    return true;
}

template <int R>
static cudaError_t launch_bucket(const F32Line& g, const F32Taps& t, cudaStream_t s)
{
    // re-centre the taps inside the bucket: kernel expects offsets -R..R
    F32Taps tb;
    tb.radius = R;
    for (int k = 0; k <= 2 * R; ++k) tb.w[k] = 0.f;
    for (int k = 0; k <= 2 * t.radius; ++k) tb.w[k + R - t.radius] = t.w[k];
    const bool al_in = (reinterpret_cast<uintptr_t>(g.in) & 15) == 0;
    const bool al_out = (reinterpret_cast<uintptr_t>(g.out) & 15) == 0;
    if (g.inner == 1) {
        const int vec_in = al_in && (g.n_in % 4 == 0) && (g.in_offset % 4 == 0);
        const int vec_out = al_out && (g.n_out % 4 == 0);
        dim3 grid((unsigned)((g.outer + ROW_ROWS - 1) / ROW_ROWS), (unsigned)((g.n_out + ROW_W - 1) / ROW_W));
        corr1d_f32_row_kernel<R><<<grid, 256, 0, s>>>(g, tb, vec_in, vec_out);
    } else {
        const int vec = al_in && al_out && (g.inner % 4 == 0);
        const int64_t n_itiles = (g.inner + COL_TI - 1) / COL_TI;
        dim3 grid((unsigned)(g.outer * n_itiles), (unsigned)((g.n_out + COL_TN - 1) / COL_TN));
        corr1d_f32_col_kernel<R><<<grid, 256, 0, s>>>(g, tb, vec, (int)n_itiles);
    }
    return cudaGetLastError();
}

cudaError_t launch_f32_corr1d(const F32Line& g, const F32Taps& t, cudaStream_t s)
{
    static const bool no_stream = getenv("SEPFILT_NO_F32_STREAM") != nullptr;   // A/B aid
    if (!no_stream && f32_stream_supported(g, t.radius)) return launch_f32_stream(g, t, s);
    switch (radius_bucket(t.radius)) {
    case 1: return launch_bucket<1>(g, t, s);
    case 2: return launch_bucket<2>(g, t, s);
    case 3: return launch_bucket<3>(g, t, s);
    case 4: return launch_bucket<4>(g, t, s);
    case 5: return launch_bucket<5>(g, t, s);
    case 6: return launch_bucket<6>(g, t, s);
    case 7: return launch_bucket<7>(g, t, s);
    case 8: return launch_bucket<8>(g, t, s);
    case 9: return launch_bucket<9>(g, t, s);
    case 10: return launch_bucket<10>(g, t, s);
    case 11: return launch_bucket<11>(g, t, s);
    case 12: return launch_bucket<12>(g, t, s);
    case 13: return launch_bucket<13>(g, t, s);
    case 14: return launch_bucket<14>(g, t, s);
    case 15: return launch_bucket<15>(g, t, s);
    case 16: return launch_bucket<16>(g, t, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace sepfilt
