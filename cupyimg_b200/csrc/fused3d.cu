// fused3d.cu — fused multi-axis separable f32 filter: the volume crosses HBM once in and
// once out instead of once per axis (the reference launches one ElementwiseKernel per axis
// plus copy-backs: filters.py:651-662 / :777-789, _filters_core.py:148-155).
//
// One CTA owns a TX x ty column of the volume and marches along z.  Per batch of PZ input
// planes:
//   stage   (TX+2HL) x (ty+2R) raw tiles -> shared memory by TMA (cp.async.bulk.tensor.3d,
//           one box per plane, mbarrier complete_tx), issued by one thread: no address
//           arithmetic in the other 511.  z is remapped per plane through the box
//           coordinate; x / y cells outside the array arrive zero-filled and — only in
//           tiles that touch the array edge, only for non-constant modes — are patched
//           from the remapped source (_util.py:170-228): boundary handling costs nothing
//           in interior tiles and is never evaluated per tap.
//   y pass  shared -> shared, register tile of 4 columns x 8 rows per thread (packed
//           fma.rn.f32x2: two columns per instruction, scalar tap broadcast).
//   x + z   per plane each thread reads a 4+2HL window of y-filtered values, produces 4
//           x-filtered outputs in registers, and scatters them into 2R+1 per-column
//           z accumulators that shift by one plane per step (acc[j] = fma(w, v, acc[j+1]),
//           no register moves); acc[0] is a finished output voxel and is stored with a
//           16-byte store.
// The next batch's TMA traffic overlaps the x + z phase.  Tensor cores are
// deliberately not used: this is a bandwidth / FP32-issue bound stencil.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace sepfilt {

namespace {

constexpr int TX = 128;     // tile width (contiguous axis)
constexpr int TYM = 16;     // max tile rows
constexpr int RY = 8;       // rows per y-pass register tile
constexpr int NT = 512;     // threads per CTA
constexpr int MAXR = SEPFILT_FAST_MAX_RADIUS;

struct FusedParams {
    const float* in;
    float*       out;
    int nz_in, nz_out, ny, nx, z_offset;
    int mode_z, mode_y, mode_x;
    float cval;
    int tiles_x, tiles_y, ty;   // ty = rows per tile (<= TYM)
    int box_rows;               // ty + 2R: height of the TMA box
    int zseg, nzseg;            // output planes per z segment, number of segments
    float wz[2 * MAXR + 1], wy[2 * MAXR + 1], wx[2 * MAXR + 1];   // taps at offsets -R..R
};

__host__ __device__ constexpr int rup4(int r) { return (r + 3) & ~3; }

template <int R> struct Cfg {
    static constexpr int HL = rup4(R);                 // x halo staged (multiple of 4 floats)
    static constexpr int PITCH = TX + 2 * HL;          // floats per staged row
    static constexpr int NCG = PITCH / 4;              // float4 column groups per row
    static constexpr int RROWS = TYM + 2 * R;          // staged rows per plane
    static constexpr int PZ = NT / (2 * NCG);          // planes per batch: 2*NCG y-items per plane
    static constexpr int NV = 2 * HL / 4 + 1;          // float4 loads of the x window
    static constexpr int RSLOT = (RROWS * PITCH + 31) & ~31;   // floats per raw plane slot (128 B multiple: TMA dst)
    static constexpr size_t SMEM = sizeof(float) * ((size_t)PZ * RSLOT + (size_t)PZ * TYM * PITCH + RROWS + PITCH) + 16;
};

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
// d = a * (w, w) + c on two packed floats; ptxas emits FFMA2 with a scalar-broadcast operand
__device__ __forceinline__ u64 fma2s(u64 a, float w, u64 c)
{
    u64 d, ww = pack2(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(ww), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2s(u64 a, float w)
{
    u64 d, ww = pack2(w, w);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(ww));
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void* ptr)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one (PITCH x rows x 1) box of the input volume -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* map, int cx, int cy, int cz, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(bar))
        : "memory");
}

template <int R, bool HAS_Z>
__global__ void __launch_bounds__(NT, 1)
fused3d_kernel(const __grid_constant__ FusedParams p, const __grid_constant__ CUtensorMap tmap)
{
    using C = Cfg<R>;
    constexpr int HL = C::HL, PITCH = C::PITCH, NCG = C::NCG, PZ = C::PZ, NV = C::NV, RSLOT = C::RSLOT;
    extern __shared__ __align__(128) float smem[];
    float* raw = smem;                                   // [PZ] slots of RSLOT floats, rows x PITCH dense inside
    float* ybuf = smem + (size_t)PZ * RSLOT;             // [PZ][TYM][PITCH]
    uint64_t& full_bar = *reinterpret_cast<uint64_t*>(ybuf + (size_t)PZ * TYM * PITCH);

    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const int tile_x = b % p.tiles_x; b /= p.tiles_x;
    const int tile_y = b % p.tiles_y; b /= p.tiles_y;
    const int seg = b;
    const int x0 = tile_x * TX, y0 = tile_y * p.ty;
    const int ty = min(p.ty, p.ny - y0);
    const int zb = seg * p.zseg, ze = min(zb + p.zseg, p.nz_out);
    // input planes this CTA consumes, in input coordinates
    const int p_first = zb + p.z_offset - (HAS_Z ? R : 0);
    const int n_planes = (ze - zb) + (HAS_Z ? 2 * R : 0);
    const int n_batches = (n_planes + PZ - 1) / PZ;
    const size_t plane_elems = (size_t)p.ny * p.nx;

    // ---- stage one batch: one TMA box per plane, issued by a single thread
    auto stage = [&](int batch) {
        const int planes = min(PZ, n_planes - batch * PZ);
        mbar_expect_tx(&full_bar, (uint32_t)planes * (uint32_t)(p.box_rows * PITCH * sizeof(float)));
        for (int q = 0; q < planes; ++q) {
            int pz = p_first + batch * PZ + q;
            if (p.mode_z != SEPFILT_CONSTANT) pz = remap_index32(p.mode_z, pz, p.nz_in);   // constant: OOB box -> zeros
            tma_load_plane(raw + (size_t)q * RSLOT, &tmap, x0 - HL, y0 - R, pz, &full_bar);
        }
    };
    // ---- patch the zero-filled cells of an edge tile from their remapped sources.
    // Source tables are built once per CTA: for a staged row / column that lies outside the
    // array, the staged row / column holding its remapped source (>= 0), or -2 when the source
    // is not inside this tile (wrap, tiny arrays: fetched from global memory instead).
    const int oob_top = min(p.box_rows, max(0, R - y0)), oob_bot = min(p.box_rows, max(0, y0 + p.box_rows - R - p.ny));
    const int oob_left = min(PITCH, max(0, HL - x0)), oob_right = min(PITCH, max(0, x0 + TX + HL - p.nx));
    const bool patch_rows = (oob_top + oob_bot) > 0 && p.mode_y != SEPFILT_CONSTANT;
    const bool patch_cols = (oob_left + oob_right) > 0 && p.mode_x != SEPFILT_CONSTANT;
    int* ysrc = reinterpret_cast<int*>(&full_bar + 1);        // [RROWS]
    int* xsrc = ysrc + C::RROWS;                               // [PITCH]
    if (patch_rows || patch_cols) {
        for (int i = tid; i < p.box_rows; i += NT) {
            const int m = remap_index32(p.mode_y, y0 - R + i, p.ny) - (y0 - R);
            ysrc[i] = (m >= 0 && m < p.box_rows) ? m : -2;
        }
        for (int i = tid; i < PITCH; i += NT) {
            const int m = remap_index32(p.mode_x, x0 - HL + i, p.nx) - (x0 - HL);
            xsrc[i] = (m >= 0 && m < PITCH) ? m : -2;
        }
    }
    auto global_cell = [&](int batch, int q, int yy, int c) -> float {
        const int pz = remap_index32(p.mode_z, p_first + batch * PZ + q, p.nz_in);
        const int gy = remap_index32(p.mode_y, y0 - R + yy, p.ny);
        const int gx = remap_index32(p.mode_x, x0 - HL + c, p.nx);
        if (pz < 0 || gy < 0 || gx < 0) return 0.f;
        return __ldg(p.in + (size_t)pz * plane_elems + (size_t)gy * p.nx + gx);
    };
    auto patch = [&](int batch, int planes) {
        if (patch_rows) {            // whole rows: float4 copies of the in-range column groups
            const int nr = oob_top + oob_bot;
            const int g0 = oob_left / 4, ng = NCG - (oob_left + oob_right) / 4;
            for (int i = tid; i < planes * nr * ng; i += NT) {
                const int q = i / (nr * ng), rem = i - q * (nr * ng);
                const int ra = rem / ng, g = g0 + rem - ra * ng;
                const int yy = ra < oob_top ? ra : p.box_rows - oob_bot + (ra - oob_top);
                float* slot = raw + (size_t)q * RSLOT;
                const int sy = ysrc[yy];
                float4 v;
                if (sy >= 0) {
                    v = *reinterpret_cast<const float4*>(slot + sy * PITCH + 4 * g);
                } else {
                    v.x = global_cell(batch, q, yy, 4 * g);     v.y = global_cell(batch, q, yy, 4 * g + 1);
                    v.z = global_cell(batch, q, yy, 4 * g + 2); v.w = global_cell(batch, q, yy, 4 * g + 3);
                }
                *reinterpret_cast<float4*>(slot + yy * PITCH + 4 * g) = v;
            }
        }
        if (patch_cols) {            // out-of-array columns of every staged row
            const int nc = oob_left + oob_right;
            for (int i = tid; i < planes * p.box_rows * nc; i += NT) {
                const int q = i / (p.box_rows * nc), rem = i - q * (p.box_rows * nc);
                const int yy = rem / nc, cb = rem - yy * nc;
                const int c = cb < oob_left ? cb : PITCH - oob_right + (cb - oob_left);
                const bool row_oob = yy < oob_top || yy >= p.box_rows - oob_bot;
                if (row_oob && p.mode_y == SEPFILT_CONSTANT) continue;        // stays zero
                float* slot = raw + (size_t)q * RSLOT;
                const int sy = row_oob ? ysrc[yy] : yy, sx = xsrc[c];
                slot[yy * PITCH + c] = (sy >= 0 && sx >= 0) ? slot[sy * PITCH + sx] : global_cell(batch, q, yy, c);
            }
        }
    };

    // per-thread z accumulators: 2R+1 shifting partial sums for 4 adjacent columns
    u64 zacc[HAS_Z ? 2 * R + 1 : 1][2];
#pragma unroll
    for (int j = 0; j < (HAS_Z ? 2 * R + 1 : 1); ++j) zacc[j][0] = zacc[j][1] = 0ull;

    const int cg_o = tid & 31, row_o = tid >> 5;          // x+z phase ownership
    const bool owner = row_o < ty && x0 + 4 * cg_o < p.nx;
    float* out_col = p.out + (size_t)(y0 + row_o) * p.nx + x0 + 4 * cg_o;

    if (tid == 0) mbar_init(&full_bar, 1);
    __syncthreads();
    if (tid == 0) stage(0);
    uint32_t parity = 0;
    for (int batch = 0; batch < n_batches; ++batch) {
        const int planes = min(PZ, n_planes - batch * PZ);
        mbar_wait(&full_bar, parity);
        parity ^= 1;
        if (patch_rows || patch_cols) patch(batch, planes);
        __syncthreads();          // patches visible; everyone is done reading ybuf of the previous batch

        // ---- y pass: raw -> ybuf, 4 columns x RY rows per thread
        {
            const int q = tid / (2 * NCG);
            const int rem = tid - q * (2 * NCG);
            const int half = rem / NCG, cg = rem - half * NCG;
            if (q < planes && half * RY < ty) {
                u64 acc[RY][2];
#pragma unroll
                for (int o = 0; o < RY; ++o) acc[o][0] = acc[o][1] = 0ull;
                const float* src = raw + (size_t)q * RSLOT + (half * RY) * PITCH + 4 * cg;
#pragma unroll
                for (int j = 0; j < RY + 2 * R; ++j) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + j * PITCH);
#pragma unroll
                    for (int o = 0; o < RY; ++o) {
                        const int k = j - o;
                        if (k >= 0 && k <= 2 * R) {
                            acc[o][0] = fma2s(v.x, p.wy[k], acc[o][0]);
                            acc[o][1] = fma2s(v.y, p.wy[k], acc[o][1]);
                        }
                    }
                }
                float* dst = ybuf + ((size_t)q * TYM + half * RY) * PITCH + 4 * cg;
#pragma unroll
                for (int o = 0; o < RY; ++o)
                    *reinterpret_cast<ulonglong2*>(dst + o * PITCH) = make_ulonglong2(acc[o][0], acc[o][1]);
            }
        }
        __syncthreads();
        if (tid == 0 && batch + 1 < n_batches) {           // raw is free: next batch overlaps the x + z phase
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            stage(batch + 1);
        }

        // ---- x pass + z scatter, plane by plane
        if (owner) {
            for (int q = 0; q < planes; ++q) {
                const float* src = ybuf + ((size_t)q * TYM + row_o) * PITCH + 4 * cg_o;
                float win[4 * NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
                    win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
                }
                float xo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k <= 2 * R; ++k) {
                    const float w = p.wx[k];
#pragma unroll
                    for (int o = 0; o < 4; ++o) xo[o] = fmaf(w, win[o + HL - R + k], xo[o]);
                }
                const int idx = batch * PZ + q;            // plane index within this CTA's march
                if (HAS_Z) {
                    const u64 v0 = pack2(xo[0], xo[1]), v1 = pack2(xo[2], xo[3]);
#pragma unroll
                    for (int j = 0; j < 2 * R; ++j) {
                        zacc[j][0] = fma2s(v0, p.wz[2 * R - j], zacc[j + 1][0]);
                        zacc[j][1] = fma2s(v1, p.wz[2 * R - j], zacc[j + 1][1]);
                    }
                    zacc[2 * R][0] = mul2s(v0, p.wz[0]);
                    zacc[2 * R][1] = mul2s(v1, p.wz[0]);
                    const int zo = zb + idx - 2 * R;       // finished output plane
                    if (zo >= zb)
                        *reinterpret_cast<ulonglong2*>(out_col + (size_t)zo * plane_elems) =
                            make_ulonglong2(zacc[0][0], zacc[0][1]);
                } else {
                    *reinterpret_cast<float4*>(out_col + (size_t)(zb + idx) * plane_elems) =
                        make_float4(xo[0], xo[1], xo[2], xo[3]);
                }
            }
        }
    }
}

int radius_bucket(int r)
{
    static const int buckets[] = {1, 2, 4, 8};
    for (int b : buckets) if (r <= b) return b;
    return -1;
}

void recentre(const F32Taps& t, int R, float* w)
{
    for (int k = 0; k <= 2 * MAXR; ++k) w[k] = 0.f;
    for (int k = 0; k <= 2 * t.radius; ++k) w[k + R - t.radius] = t.w[k];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

template <int R, bool HAS_Z>
cudaError_t launch_t(FusedParams& p, cudaStream_t s)
{
    using C = Cfg<R>;
    static_assert(C::PZ >= 1, "tile too wide");
    p.box_rows = p.ty + 2 * R;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.nx, (cuuint64_t)p.ny, (cuuint64_t)p.nz_in};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.nx * 4, (cuuint64_t)p.nx * (cuuint64_t)p.ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)C::PITCH, (cuuint32_t)p.box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.in), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    auto kern = fused3d_kernel<R, HAS_Z>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.nzseg;
    kern<<<(unsigned)blocks, NT, C::SMEM, s>>>(p, tmap);
    return cudaGetLastError();
}

}  // namespace

bool fused3d_supported(const FusedVolume& v, const F32Taps taps[3], bool gradmag)
{
    if (gradmag) return false;      // fused gradient-magnitude epilogue: next step
    int r = 0;
    for (int a = 0; a < 3; ++a) r = taps[a].radius > r ? taps[a].radius : r;
    if (radius_bucket(r) < 0) return false;
    if (v.nx % 4 != 0 || (reinterpret_cast<uintptr_t>(v.in) & 15) || (reinterpret_cast<uintptr_t>(v.out) & 15))
        return false;
    if (v.nx < 1 || v.ny < 1 || v.nz_in < 1 || v.nz_out < 1) return false;
    const bool has_z = !(taps[0].radius == 0 && taps[0].w[0] == 1.0f);
    if (!has_z && v.nz_out < 4) return false;   // a lone 2-D image cannot fill a plane batch: per-axis passes
    // constant mode with cval != 0 is affine, not linear: the passes no longer commute and the
    // halo of an intermediate is cval, not filtered cval -> per-axis passes in scipy's order
    if (v.cval != 0.f)
        for (int a = 0; a < 3; ++a)
            if (v.mode[a] == SEPFILT_CONSTANT && taps[a].radius > 0) return false;
    const long long tiles = (long long)((v.nx + TX - 1) / TX) * v.ny;
    if (tiles * v.nz_out > 2147483647LL) return false;
    return true;
}

cudaError_t launch_fused3d(const FusedVolume& v, const F32Taps taps[3], const F32Taps[3], bool gradmag,
                           cudaStream_t s)
{
    if (gradmag) return cudaErrorNotSupported;
    int r = 0;
    for (int a = 0; a < 3; ++a) r = taps[a].radius > r ? taps[a].radius : r;
    const int R = radius_bucket(r);
    if (R < 0) return cudaErrorInvalidValue;
    const bool has_z = !(taps[0].radius == 0 && taps[0].w[0] == 1.0f);

    FusedParams p;
    p.in = v.in; p.out = v.out;
    p.nz_in = v.nz_in; p.nz_out = v.nz_out; p.ny = v.ny; p.nx = v.nx; p.z_offset = v.z_offset;
    p.mode_z = v.mode[0]; p.mode_y = v.mode[1]; p.mode_x = v.mode[2];
    p.cval = v.cval;
    recentre(taps[0], R, p.wz);
    recentre(taps[1], R, p.wy);
    recentre(taps[2], R, p.wx);

    // tile rows / z segments: fill the 148 SMs with whole waves where the shape allows
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.tiles_x = (v.nx + TX - 1) / TX;
    int best_ty = TYM, best_seg = 1;
    double best_cost = 1e300;
    for (int ty = TYM; ty >= 8; --ty) {
        const int tiles_y = (v.ny + ty - 1) / ty;
        for (int nseg = 1; nseg <= 64 && nseg <= v.nz_out; ++nseg) {
            const int zseg = (v.nz_out + nseg - 1) / nseg;
            const long long ctas = (long long)p.tiles_x * tiles_y * nseg;
            const long long waves = (ctas + sms - 1) / sms;
            // per-CTA time ~ planes marched x (y-pass on a full 16-row tile + x/z on ty rows)
            const double per_cta = (double)(zseg + (has_z ? 2 * R : 0)) * (0.36 * TYM + 0.64 * ty);
            const double cost = waves * per_cta;
            if (cost < best_cost) { best_cost = cost; best_ty = ty; best_seg = nseg; }
            if (!has_z) break;
        }
    }
    if (v.ny <= TYM) best_ty = v.ny < 1 ? 1 : (v.ny < TYM ? v.ny : TYM);
    p.ty = best_ty;
    p.tiles_y = (v.ny + p.ty - 1) / p.ty;
    p.nzseg = best_seg;
    p.zseg = (v.nz_out + best_seg - 1) / best_seg;
    p.nzseg = (v.nz_out + p.zseg - 1) / p.zseg;

    switch (R * 2 + (has_z ? 1 : 0)) {
    case 1 * 2 + 0: return launch_t<1, false>(p, s);
    case 1 * 2 + 1: return launch_t<1, true>(p, s);
    case 2 * 2 + 0: return launch_t<2, false>(p, s);
    case 2 * 2 + 1: return launch_t<2, true>(p, s);
    case 4 * 2 + 0: return launch_t<4, false>(p, s);
    case 4 * 2 + 1: return launch_t<4, true>(p, s);
    case 8 * 2 + 0: return launch_t<8, false>(p, s);
    case 8 * 2 + 1: return launch_t<8, true>(p, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace sepfilt
