// correlate_nd.cu — dense N-d correlation for small kernels: one launch of the
// reference's generated correlate kernel for N-d weights (filters.py:65-210 -> _correlate_or_convolve
// :441-495, kernel body _filters_core.py:239-312).
//
// scipy's NI_Correlate arithmetic: float64, `tmp += x * w` over the taps with |w| > DBL_EPSILON in C order of
// the weights array (explicit __dmul_rn / __dadd_rn, never contracted), boundary extension per axis
// (_util.py:170-228), then the C-cast store — bit-exact integer outputs, bit-identical float64.  Any (in, out)
// dtype pair, any byte strides, up to SEPFILT_MAX_NDIM dimensions.  One output element per thread
// (grid-stride); the boundary rule is evaluated only for threads whose footprint leaves the array.
#include <cmath>
#include <cstdlib>
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

// ---- tile kernel for small 2-D weights (kh, kw <= 7: 3x3, 5x5, 7x7 ... on images or stacks of images) ----
// The per-pixel kernel above converts every element to float64 once per TAP (9 .. 49 times) and walks the taps
// with runtime loops.  Here a CTA stages the (32 + kh - 1) x (128 + KW - 1) footprint of a 128 x 32 output tile
// in shared memory AS FLOAT64 — every element is loaded, boundary-mapped and converted exactly once — and a
// thread then owns a 4 x 4 block of outputs: one row of its window (4 + KW - 1 doubles, LDS.128) feeds up to
// 4 output rows x KW taps x 4 columns of DMUL + DADD from registers.  Same arithmetic and the same order of
// operations per output as NI_Correlate (ky outer, kx inner, taps with |w| <= DBL_EPSILON skipped): bit-exact.
// The kernel width is rounded up to KW = 3 / 5 / 7 with zero taps on the right, which the skip rule drops.
constexpr int CT_W = 128, CT_H = 32, CT_MAXK = 7;
constexpr int CT_ODD = 36;                  // 16-byte chunk (2 doubles) c lives at (c >> 1) + (c & 1) * CT_ODD: a lane's
                                            // window starts 2 chunks after its neighbour's, and with even / odd chunks
                                            // in separate halves (4 mod 8 apart) the LDS.128 and the staging STS.64 of a
                                            // warp touch consecutive chunks
constexpr int CT_PITCH = 2 * (CT_ODD + (CT_W + CT_MAXK - 1) / 4 + 1);     // doubles per staged row
constexpr double CT_EPS = 2.220446049250313e-16;

// taps as kernel parameters in the padded [ky][KW] layout: every tap address and every skip test is a compile-time
// offset into the constant bank (bit ky * KW + kx of `mask`: |w| > DBL_EPSILON, scipy's footprint)
struct CtWeights {
    double w[CT_MAXK * CT_MAXK];
    unsigned long long mask;
};

__device__ __forceinline__ int ct_off(int ty, int tx)
{
    const int c = tx >> 1;
    return ty * CT_PITCH + 2 * ((c >> 1) + (c & 1) * CT_ODD) + (tx & 1);
}

template <typename InT, int KW>
__global__ void __launch_bounds__(256, 3)          // 3 CTAs / SM: 0.162 ms on 8 x 2048^2 f32 3x3 against 0.205 ms at 2
correlate_2d_tile_kernel(const __grid_constant__ CorrNdParams p, const __grid_constant__ CtWeights cw, const int ny, const int nx,
                         const int kh, const int by, const int bx, const int64_t planes)
{
    constexpr int SW = CT_W + KW - 1;                         // staged columns
    constexpr int NWC = (4 + KW - 1 + 1) / 2;                 // 16-byte chunks of a thread's window row
    __shared__ __align__(16) double tile[(CT_H + CT_MAXK - 1) * CT_PITCH];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const int sh = CT_H + kh - 1;                             // staged rows
    const int ys0 = y0 - by, xs0 = x0 - bx;                   // source coordinate of staged element (0, 0)
    const bool interior = ys0 >= 0 && ys0 + sh <= ny && xs0 >= 0 && xs0 + SW <= nx;
    const int total = sh * SW;
    const int tx = tid & 31, tg = tid >> 5;
    const int ox = x0 + 4 * tx, oy = y0 + 4 * tg;
    const int osize = dtype_size(p.out_dtype);
    constexpr int VEC = 16 / (int)sizeof(InT);                // elements per 16-byte staging load
    constexpr int NVR = (SW + VEC - 1) / VEC + 1;             // aligned vectors that cover a staged row at any phase
    constexpr int NITV = ((CT_H + CT_MAXK - 1) * NVR + 255) / 256;
    const bool vec_in = interior && nx % VEC == 0 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
    const bool vec_out = (nx & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && ox + 3 < nx;

    for (int64_t z = blockIdx.z; z < planes; z += gridDim.z) {
        const InT* plane = reinterpret_cast<const InT*>(p.in) + z * (int64_t)ny * nx;
        __syncthreads();                                      // the previous plane's tile is consumed
        // ---- stage.  Interior tiles: every load of the thread is in flight before its first conversion (a tile is
        //      ~18 elements per thread; no branch sits between the loads, so ptxas keeps them batched — with the
        //      boundary rule inline every F2F waited for its own load: 53 % long-scoreboard stalls, 0.378 ms)
        constexpr int NIT = ((CT_H + CT_MAXK - 1) * SW + 255) / 256;
        if (vec_in) {
            // 16-byte loads from the aligned range that covers the staged columns: a quarter of the load instructions
            // and of the index arithmetic of the element-wise path below (VEC elements per load)
            const int a0 = xs0 & ~(VEC - 1), shift = xs0 - a0;
            const InT* src = plane + (int64_t)ys0 * nx + a0;
            uint4 raw[NITV];
#pragma unroll
            for (int u = 0; u < NITV; ++u) {
                const int e = tid + 256 * u;
                const int ty = e / NVR, vi = e - ty * NVR;
                raw[u] = make_uint4(0u, 0u, 0u, 0u);
                if (ty < sh && vi * VEC < shift + SW) raw[u] = *reinterpret_cast<const uint4*>(src + ty * nx + vi * VEC);
            }
#pragma unroll
            for (int u = 0; u < NITV; ++u) {
                const int e = tid + 256 * u;
                const int ty = e / NVR, vi = e - ty * NVR;
                alignas(16) InT el[VEC];
                *reinterpret_cast<uint4*>(el) = raw[u];
                if (ty < sh) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const int c = vi * VEC + v - shift;
                        if (c >= 0 && c < SW) tile[ct_off(ty, c)] = (double)el[v];
                    }
                }
            }
        } else if (interior) {
            const InT* src = plane + (int64_t)ys0 * nx + xs0;
            InT raw[NIT];
#pragma unroll
            for (int u = 0; u < NIT; ++u) {
                const int e = tid + 256 * u;
                const int ty = e / SW, txs = e - ty * SW;
                raw[u] = e < total ? src[ty * nx + txs] : InT(0);
            }
#pragma unroll
            for (int u = 0; u < NIT; ++u) {
                const int e = tid + 256 * u;
                const int ty = e / SW, txs = e - ty * SW;
                if (e < total) tile[ct_off(ty, txs)] = (double)raw[u];
            }
        } else {
#pragma unroll 1
            for (int e = tid; e < total; e += 256) {
                const int ty = e / SW, txs = e - ty * SW;
                const int sy = remap_index32(p.mode, ys0 + ty, ny), sx = remap_index32(p.mode, xs0 + txs, nx);
                tile[ct_off(ty, txs)] = (sy < 0 || sx < 0) ? p.cval : (double)plane[(int64_t)sy * nx + sx];
            }
        }
        __syncthreads();
        // ---- compute: 4 x 4 outputs per thread; input row r of the block feeds output rows i with ky = r - i
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        const double* wrow = tile + (4 * tg) * CT_PITCH + 2 * tx;     // chunk 2 tx + m -> tx + (m >> 1) + (m & 1) CT_ODD
        const unsigned long long mask = cw.mask;
#pragma unroll
        for (int r = 0; r < 4 + CT_MAXK - 1; ++r) {
            if (r < kh + 3) {                                 // rows beyond the kernel height are not staged
                double win[2 * NWC];
#pragma unroll
                for (int m = 0; m < NWC; ++m) {
                    const double2 q = *reinterpret_cast<const double2*>(wrow + r * CT_PITCH + 2 * ((m >> 1) + (m & 1) * CT_ODD));
                    win[2 * m] = q.x; win[2 * m + 1] = q.y;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ky = r - i;                     // compile time; rows ky >= kh have no mask bits
                    if (ky >= 0 && ky < CT_MAXK) {
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx) {
                            if ((mask >> (ky * KW + kx)) & 1ull) {
                                const double w = cw.w[ky * KW + kx];
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(win[j + kx], w));
                            }
                        }
                    }
                }
            }
        }
        // ---- store under scipy's cast rules
        char* obase = p.out + (z * (int64_t)ny * nx) * osize;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = oy + i;
            if (y >= ny) break;
            const int64_t eo = (int64_t)y * nx + ox;
            if (vec_out && p.out_dtype == SEPFILT_F32) {
                *reinterpret_cast<float4*>(obase + eo * 4) = make_float4(__double2float_rn(acc[i][0]), __double2float_rn(acc[i][1]),
                                                                         __double2float_rn(acc[i][2]), __double2float_rn(acc[i][3]));
            } else if (vec_out && p.out_dtype == SEPFILT_F64) {
                *reinterpret_cast<double2*>(obase + eo * 8) = make_double2(acc[i][0], acc[i][1]);
                *reinterpret_cast<double2*>(obase + eo * 8 + 16) = make_double2(acc[i][2], acc[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (ox + j < nx) store_cast(obase + (eo + j) * osize, p.out_dtype, acc[i][j]);
            }
        }
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
    const int kh = p.wshape[p.ndim - 2], kw = p.wshape[p.ndim - 1];
    static const bool no_tile = getenv("SEPFILT_NO_CORR_TILE") != nullptr;      // A/B aid
    if (!no_tile && !p.wdev && kh <= CT_MAXK && kw <= CT_MAXK && (ny + CT_H - 1) / CT_H <= 65535) {
        dim3 tgrid((unsigned)((nx + CT_W - 1) / CT_W), (unsigned)((ny + CT_H - 1) / CT_H), (unsigned)(planes < 65535 ? planes : 65535));
        const int by = p.before[p.ndim - 2], bx = p.before[p.ndim - 1];
        const int KWT = kw <= 3 ? 3 : kw <= 5 ? 5 : 7;
        CtWeights cw;
        cw.mask = 0;
        for (int ky = 0; ky < CT_MAXK; ++ky)
            for (int kx = 0; kx < KWT; ++kx) {
                const double w = (ky < kh && kx < kw) ? p.w[ky * kw + kx] : 0.0;
                cw.w[ky * KWT + kx] = w;
                if (std::fabs(w) > CT_EPS) cw.mask |= 1ull << (ky * KWT + kx);
            }
        if (KWT == 3) correlate_2d_tile_kernel<InT, 3><<<tgrid, 256, 0, s>>>(p, cw, (int)ny, (int)nx, kh, by, bx, planes);
        else if (KWT == 5) correlate_2d_tile_kernel<InT, 5><<<tgrid, 256, 0, s>>>(p, cw, (int)ny, (int)nx, kh, by, bx, planes);
        else correlate_2d_tile_kernel<InT, 7><<<tgrid, 256, 0, s>>>(p, cw, (int)ny, (int)nx, kh, by, bx, planes);
        *err = cudaGetLastError();
        return true;
    }
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
