// fused3d.cu — fused multi-axis separable f32 filter: the volume crosses HBM once in and
// once out instead of once per axis (the reference launches one ElementwiseKernel per axis
// plus copy-backs: filters.py:651-662 / :777-789, _filters_core.py:148-155).
//
// One CTA owns a TX x ty column of the volume and marches along z.  Per group of G input
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
// Raw tiles and y-filtered tiles are double buffered: the y pass of group k+1, the x + z pass of
// group k and the TMA traffic of groups k+2 / k+3 overlap, with one CTA barrier per group.  Tensor cores are
// deliberately not used: this is a bandwidth / FP32-issue bound stencil.
#include <cuda.h>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace sepfilt {

namespace {

constexpr int TYM = 16;     // max tile rows
constexpr int MAXR = SEPFILT_FAST_MAX_RADIUS;

struct FusedParams {
    const float* in;
    float*       out;
    int nz_in, nz_out, ny, nx, z_offset;
    int mode_z, mode_y, mode_x;
    float cval;
    int tiles_x, tiles_y, ty;   // ty = rows per tile (<= TYM)
    int pad_;                   // keeps the tap arrays at 4 (mod 8) bytes, see the static_assert below
    int yshift;                 // tile row t covers rows [t * ty - yshift, (t + 1) * ty - yshift): the rows a whole
                                // number of tiles overshoots the array by are split between the first and the last
                                // tile row — the two that pay for y-edge patching — instead of the last alone
    int box_rows;               // ty + 2R: height of the TMA box
    int zseg, nzseg;            // output planes per z segment, number of segments
    float wz[2 * MAXR + 1], wy[2 * MAXR + 1], wx[2 * MAXR + 1];   // taps at offsets -R..R
    int epilogue;               // 0: out = v   1: out = v*v   2: out += v*v   3: out = sqrt(out + v*v)
};

// Measured on B200 (nvcc 12.9): with the tap arrays at an offset of 4 (mod 8) in the kernel parameter
// block ptxas pairs the constant-bank taps with the packed FFMA2 operands without extra moves; at
// 0 (mod 8) the same source compiles to ~70 more MOV/IMAD per loop body and runs 8 % slower.
static_assert(offsetof(FusedParams, wz) % 8 == 4, "keep the fused kernel's tap arrays at 4 (mod 8) bytes");

__host__ __device__ constexpr int rup4(int r) { return (r + 3) & ~3; }

// TX: tile width along the contiguous axis; NT = TX * TYM / 4 threads (one float4 column group x one
// row each in the x+z pass); G: planes per pipeline group; CTAS: co-resident CTAs per SM.
// RY: rows per y-pass register tile; a tile of ty <= 2 * RY rows is covered by two row halves.
// (Three 5-row parts, i.e. 432 y-pass items spread over 13.5 warps so that every warp carries both x+z and
//  y work — the CTA barrier is 17.5 % of the warp time because the 7 warps without y items idle — were
//  measured: 0.297 -> 0.310 ms constant, 0.329 -> 0.373 ms reflect on 512^3: 37 % more shared-memory loads
//  and no idle warp left for the edge patch cost more than the balance gains.)
template <int R, int TX_, int G_, int CTAS_, int RY_ = 8> struct Cfg {
    static constexpr int TX = TX_, NT = TX_ * TYM / 4, CTAS = CTAS_, RY = RY_;
    static constexpr int HL = rup4(R);                 // x halo staged (multiple of 4 floats)
    static constexpr int PITCH = TX + 2 * HL;          // floats per staged row
    static constexpr int NCG = PITCH / 4;              // float4 column groups per row
    static constexpr int RROWS = TYM + 2 * R;          // staged rows per plane
    static constexpr int G = G_;                       // planes per group (the pipeline granule)
    static constexpr int NV = 2 * HL / 4 + 1;          // float4 loads of the x window
    static constexpr int RSLOT = (RROWS * PITCH + 31) & ~31;   // floats per raw plane slot (128 B multiple: TMA dst)
    static constexpr int YSLOT = TYM * PITCH;          // floats per y-filtered plane
    static constexpr int ITEMS = 2 * NCG * G;          // y-pass items (4 cols x RY rows) per group
    static constexpr int MAXCELL = RROWS * 2 * HL;     // out-of-array column cells of one plane slot
    // raw[2][G][RSLOT] + ybuf[2][G][YSLOT] + patch tables (cells + rows) + 3 mbarriers + patch parameters
    static constexpr size_t SMEM = sizeof(float) * (2 * (size_t)G * (RSLOT + YSLOT)) + sizeof(int) * ((MAXCELL + RROWS + 1) & ~1) + 96;
    static_assert(ITEMS <= NT, "one y-pass item per thread");
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// wait (one thread, once per CTA, before the plane loop) until both neighbours' slabs are complete
__device__ __noinline__ void halo_wait_ready(const ptx::HaloMaps& hm)
{
    if (hm.planes_lo && hm.ready_lo) ptx::wait_flag_geq(hm.ready_lo, hm.epoch);
    if (hm.planes_hi && hm.ready_hi) ptx::wait_flag_geq(hm.ready_hi, hm.epoch);
    ptx::fence_proxy_async_all();
}

// generic_gradient_magnitude staging in the output dtype (filters.py:1187-1201), fused into the store:
// 1: out = v*v   2: out += v*v   3: out = sqrt(out + v*v); explicit _rn ops, no FMA contraction, so
// the float32 roundings are the ones of the reference's separate multiply / add / sqrt kernels.
__device__ __forceinline__ float4 epilogue(int mode, float4 v, const float4 old)
{
    float4 sq = make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w));
    if (mode == 1) return sq;
    sq = make_float4(__fadd_rn(old.x, sq.x), __fadd_rn(old.y, sq.y), __fadd_rn(old.z, sq.z), __fadd_rn(old.w, sq.w));
    if (mode == 2) return sq;
    return make_float4(__fsqrt_rn(sq.x), __fsqrt_rn(sq.y), __fsqrt_rn(sq.z), __fsqrt_rn(sq.w));
}

// Cold path of the edge patch: a staged cell whose remapped source is not inside the tile (wrap, or
// arrays smaller than the halo).  Deliberately NOT inlined: inlined, its three modulo remaps were
// if-converted into the hot patch loop and cost ~4000 cycles per group in every edge tile.
__device__ __noinline__ float fetch_remapped_cell(const FusedParams& p, int pz, int gy, int gx)
{
    pz = remap_index32(p.mode_z, pz, p.nz_in);
    gy = remap_index32(p.mode_y, gy, p.ny);
    gx = remap_index32(p.mode_x, gx, p.nx);
    if (pz < 0 || gy < 0 || gx < 0) return 0.f;
    return __ldg(p.in + ((size_t)pz * p.ny + gy) * p.nx + gx);
}

#ifdef SEPFILT_DEBUG_CYCLES
} __device__ long long g_dbg_cycles[4096]; namespace {
#endif

// HALO: the planes beyond the slab's ends come from the neighbours' arrays (multi-GPU z-slabs); a separate
// instantiation so that the single-GPU kernel keeps its register allocation
template <int R, bool HAS_Z, class C, bool EPI, bool HALO>
__global__ void __launch_bounds__(C::NT, C::CTAS)
fused3d_kernel(const __grid_constant__ FusedParams p, const __grid_constant__ CUtensorMap tmap,
               const __grid_constant__ ptx::HaloMaps hm)
{
    constexpr int TX = C::TX, NT = C::NT, CGW = TX / 4, RY = C::RY;
    constexpr int HL = C::HL, PITCH = C::PITCH, NCG = C::NCG, G = C::G, NV = C::NV;
    constexpr int RSLOT = C::RSLOT, YSLOT = C::YSLOT;
    extern __shared__ __align__(128) float smem[];
    float* raw = smem;                                    // [2][G] plane slots, box_rows x PITCH dense in each
    float* ybuf = smem + 2 * G * RSLOT;                   // [2][G][TYM][PITCH]
    int* meta = reinterpret_cast<int*>(ybuf + 2 * G * YSLOT);          // patch tables: [MAXCELL] cells + [RROWS] rows
    uint64_t* rfull = reinterpret_cast<uint64_t*>(meta + ((C::MAXCELL + C::RROWS + 1) & ~1));   // [2] raw slot landed

    const int tid = threadIdx.x;
#ifdef SEPFILT_DEBUG_CYCLES
    const long long t_start = clock64();
    if (tid == 0) { g_dbg_cycles[1024 + blockIdx.x] = 0; g_dbg_cycles[2048 + blockIdx.x] = 0; g_dbg_cycles[3072 + blockIdx.x] = 0; }
#endif
    int b = blockIdx.x;
    const int tile_x = b % p.tiles_x; b /= p.tiles_x;
    const int tile_y = b % p.tiles_y; b /= p.tiles_y;
    const int seg = b;
    const int x0 = tile_x * TX, y0 = tile_y * p.ty - p.yshift;   // y0 < 0 in the first tile row when yshift > 0
    const int ty = min(p.ty, p.ny - y0);
    const int zb = seg * p.zseg, ze = min(zb + p.zseg, p.nz_out);
    // input planes this CTA consumes, in input coordinates
    const int p_first = zb + p.z_offset - (HAS_Z ? R : 0);
    const int n_planes = (ze - zb) + (HAS_Z ? 2 * R : 0);
    const int n_groups = (n_planes + G - 1) / G;
    const size_t plane_elems = (size_t)p.ny * p.nx;

    // ---- stage one group: one TMA box per plane into raw slot (g & 1), issued by a single thread
    int* iinfo = reinterpret_cast<int*>(rfull + 4) + 8;    // the issuing thread's coordinates, kept out of registers
    if (tid == 0) { iinfo[0] = x0 - HL; iinfo[1] = y0 - R; iinfo[2] = p_first; iinfo[3] = n_planes; if (HALO) halo_wait_ready(hm); }
    auto issue = [&](int g) {                               // thread 0 only
        const int cx = iinfo[0], cy = iinfo[1], first = iinfo[2];
        const int planes = min(G, iinfo[3] - g * G);
        uint64_t* bar = &rfull[g & 1];
        mbar_expect_tx(bar, (uint32_t)planes * (uint32_t)(p.box_rows * PITCH * sizeof(float)));
        for (int q = 0; q < planes; ++q) {
            int pz = first + g * G + q;
            if (HALO) {
                // planes beyond the slab's ends: the neighbours' slabs, read in place over NVLink (their ready
                // flags were awaited once, before the plane loop)
                const CUtensorMap* map = &tmap;
                if (pz < 0 && hm.planes_lo) { map = &hm.lo; pz += hm.planes_lo; }
                else if (pz >= p.nz_in && hm.planes_hi) { map = &hm.hi; pz -= p.nz_in; }
                else if (p.mode_z != SEPFILT_CONSTANT) pz = remap_index32(p.mode_z, pz, p.nz_in);
                tma_load_plane(raw + ((g & 1) * G + q) * RSLOT, map, cx, cy, pz, bar);
                continue;
            }
            if (p.mode_z != SEPFILT_CONSTANT) pz = remap_index32(p.mode_z, pz, p.nz_in);   // constant: OOB box -> zeros
            tma_load_plane(raw + ((g & 1) * G + q) * RSLOT, &tmap, cx, cy, pz, bar);
        }
    };

    // ---- edge tiles: the cells TMA zero-filled are patched from their remapped sources.
    // The (destination, source) offsets of every out-of-array column cell and row of a plane slot
    // are tabulated once per CTA.  The patch of group k+2 is executed during phase k by the warps
    // that have NO y-pass items: they finish their x+z work first and would otherwise idle at the
    // group barrier, so boundary handling hides in that slack.  Source 0xffff = not inside this tile
    // (wrap, arrays smaller than the halo): fetched from global memory.
    // (rows / columns further than the halo beyond the array end are never read by a stored voxel)
    const int oob_top = min(p.box_rows, max(0, R - y0)), oob_bot = min(p.box_rows, max(0, y0 + p.box_rows - R - p.ny));
    const int oob_left = min(PITCH, max(0, HL - x0)), oob_right = min(PITCH, max(0, x0 + TX + HL - p.nx));
    const int need_bot = min(oob_bot, R), need_right = min(oob_right, HL);
    const int need_top = min(oob_top, R);                             // staged rows further above the array are never read
    const bool patch_rows = (oob_top + oob_bot) > 0 && p.mode_y != SEPFILT_CONSTANT;
    // Measured and rejected for the x edges (512^3, sigma 2, reflect; per-CTA cycle counters): letting the
    // y-pass thread that produces a source column also store it to its mirror position in the y-filtered
    // tile (no raw column cells, no patch team in x-edge tiles) — the extra dependent LDS/STS chain on
    // the y-pass warps, which are the critical path of a phase, cost an x-edge CTA +21 % against +19 %
    // for the table-driven patch below (0.337 -> 0.403 ms for the launch).  Also measured: TMA prefetch
    // of the planes 2 / 4 / 8 groups ahead into L2 (cp.async.bulk.prefetch.tensor): no gain in constant
    // mode, 0.336 -> 0.375 ms in reflect mode (the issuing thread belongs to the patch team).
    const bool patch_cols = (oob_left + oob_right) > 0 && p.mode_x != SEPFILT_CONSTANT;
    const bool patching = patch_rows || patch_cols;
    constexpr int NONE = -1;
    constexpr int NW = NT / 32, NYW = (C::ITEMS + 31) / 32, NPW = NW - NYW;   // warps: all, y-pass, patch
    static_assert(NPW >= 1, "no warp left for patching");
    const int rg0 = oob_left / 4, rg1 = NCG - oob_right / 4;           // in-range column groups [rg0, rg1)
    const int ncols = patch_cols ? oob_left + need_right : 0;          // out-of-array columns per staged row
    const int rows_used = p.box_rows - (oob_bot - need_bot);           // staged rows a stored voxel can read
    const int ncell = ncols * rows_used;                               // column cells per plane (<= C::MAXCELL)
    const int nrow = patch_rows ? need_top + need_bot : 0;             // out-of-array rows per plane
    int* pcell = meta;                                                 // [MAXCELL] dst | src << 16
    int* prow = meta + C::MAXCELL;                                     // [RROWS]   dst | src << 16 (row offsets)
    // ncell / nrow / rg0 / rg1 live in shared memory (pinfo), not in registers: the y and x+z passes are
    // register-bound, and every value kept live across them costs ptxas scheduling freedom (measured:
    // 0.348 -> 0.318 ms on 512^3 in constant mode, which never even runs a patch)
    int* pinfo = reinterpret_cast<int*>(rfull + 4);
    if (tid == 0) {
        pinfo[0] = ncell; pinfo[1] = nrow; pinfo[2] = rg0; pinfo[3] = rg1;
        pinfo[4] = (ty * (TX / 4) + 31) / 32;               // warps that own output rows
    }
    if (patching) {
        auto staged_row = [&](int yy) {      // staged row holding the source of row yy; -1 stays zero; -2 not in tile
            const int gy = remap_index32(p.mode_y, y0 - R + yy, p.ny);
            if (gy < 0) return -1;
            const int m = gy - (y0 - R);
            return (m >= 0 && m < p.box_rows) ? m : -2;
        };
        auto staged_col = [&](int c) {
            const int gx = remap_index32(p.mode_x, x0 - HL + c, p.nx);
            if (gx < 0) return -1;
            const int m = gx - (x0 - HL);
            return (m >= 0 && m < PITCH) ? m : -2;
        };
        for (int i = tid; i < ncell; i += NT) {
            const int yy = i / ncols, cb = i - yy * ncols;
            const int c = cb < oob_left ? cb : PITCH - oob_right + (cb - oob_left);
            const bool row_oob = yy < oob_top || yy >= p.box_rows - oob_bot;
            const int sy = row_oob ? staged_row(yy) : yy, sx = staged_col(c);
            int m = NONE;
            if (sy != -1 && sx != -1)                    // -1: constant-mode zero stays
                m = (yy * PITCH + c) | (((sy >= 0 && sx >= 0) ? sy * PITCH + sx : 0xffff) << 16);
            pcell[i] = m;
        }
        for (int i = tid; i < nrow; i += NT) {
            const int ry = i < need_top ? oob_top - need_top + i : p.box_rows - oob_bot + (i - need_top);
            const int sy = staged_row(ry);
            prow[i] = sy == -1 ? NONE : ((ry * PITCH) | ((sy >= 0 ? sy * PITCH : 0xffff) << 16));
        }
    }
    auto global_cell = [&](int g, int q, int off) -> float {
        const int yy = off / PITCH, c = off - yy * PITCH;
        return fetch_remapped_cell(p, p_first + g * G + q, y0 - R + yy, x0 - HL + c);
    };
    // patch team: the NPW warps that have no y-pass items plus the warps that own no output row of this
    // tile (ty < 16): both idle at the start of a phase.  pt = index inside the team, npt = team size.
    auto patch = [&](int g) {
        const int ncell = pinfo[0], nrow = pinfo[1], rg0 = pinfo[2], rg1 = pinfo[3];
        const int xw = pinfo[4];                            // warps owning output rows: [0, xw)
        const int first_free = max(xw, NPW);
        const int pt = (tid >> 5) < NPW ? tid : tid - (first_free - NPW) * 32;
        const int npt = (NPW + NW - first_free) * 32;
        int* pcell = meta;
        int* prow = meta + C::MAXCELL;
        // The patch warps share their SM sub-partition with FMA-saturated warps, so every dependent
        // instruction costs tens of cycles: all loads of a cell (its G planes) are issued before its
        // stores, and the cold global path is a separate loop.
        const int planes = min(G, n_planes - g * G);
        float* base = raw + (g & 1) * G * RSLOT;
        // rows: float4 copies of the in-range column groups; work item = (row, group) in slots of 64 groups
        for (int i = pt; i < nrow * 64; i += npt) {
            const int ra = i >> 6, gq = rg0 + (i & 63);
            if (gq >= rg1) continue;
            const int mr = prow[ra];
            if (mr == NONE) continue;
            const int dst = mr & 0xffff, src = (mr >> 16) & 0xffff;
            {
                if (src != 0xffff) {
                    float4 v[G];
#pragma unroll
                    for (int q = 0; q < G; ++q)
                        v[q] = *reinterpret_cast<const float4*>(base + q * RSLOT + src + 4 * gq);
#pragma unroll
                    for (int q = 0; q < G; ++q)
                        if (q < planes) *reinterpret_cast<float4*>(base + q * RSLOT + dst + 4 * gq) = v[q];
                } else {
                    for (int q = 0; q < planes; ++q) {
                        float4 v;
                        v.x = global_cell(g, q, dst + 4 * gq);     v.y = global_cell(g, q, dst + 4 * gq + 1);
                        v.z = global_cell(g, q, dst + 4 * gq + 2); v.w = global_cell(g, q, dst + 4 * gq + 3);
                        *reinterpret_cast<float4*>(base + q * RSLOT + dst + 4 * gq) = v;
                    }
                }
            }
        }
        for (int i = pt; i < ncell; i += npt) {            // column cells
            const int mc = pcell[i];
            if (mc == NONE) continue;
            const int dst = mc & 0xffff, src = (mc >> 16) & 0xffff;
            if (src != 0xffff) {
                float v[G];
#pragma unroll
                for (int q = 0; q < G; ++q) v[q] = base[q * RSLOT + src];
#pragma unroll
                for (int q = 0; q < G; ++q)
                    if (q < planes) base[q * RSLOT + dst] = v[q];
            } else {
                for (int q = 0; q < planes; ++q) base[q * RSLOT + dst] = global_cell(g, q, dst);
            }
        }
    };

    // ---- y pass of one group: raw slot -> ybuf slot, 4 columns x RY rows per thread
    // (the y-pass items live on the LAST warps of the CTA: when ty < 16 those warps own no output
    //  rows, which evens out the issue load of the four SM sub-partitions)
    constexpr int Y_TID0 = NT - ((C::ITEMS + 31) / 32) * 32;
    const int yi = tid - Y_TID0;
    const int yq = yi / (2 * NCG), yrem = yi - yq * (2 * NCG);
    const int yhalf = yrem / NCG, ycg = yrem - yhalf * NCG;
    const bool y_item = yi >= 0 && yi < C::ITEMS && yhalf * RY < ty;
    uint64_t* pdone = rfull + 2;                            // raw slot of this phase patched (edge tiles)
    auto ypass = [&](int g) {
        if (yi < 0) return;                                // not a y-pass warp
#ifdef SEPFILT_DEBUG_CYCLES
        const long long tw0 = clock64();
#endif
        if (patching) mbar_wait(pdone, (uint32_t)g & 1u);  // patched (implies landed)
        else mbar_wait(&rfull[g & 1], (uint32_t)(g >> 1) & 1u);   // TMA landed
#ifdef SEPFILT_DEBUG_CYCLES
        if (tid == NT - 1) g_dbg_cycles[1024 + blockIdx.x] += clock64() - tw0;
#endif
        if (!y_item || g * G + yq >= n_planes) return;
        u64 acc[RY][2];
        const float* src = raw + ((g & 1) * G + yq) * RSLOT + (yhalf * RY) * PITCH + 4 * ycg;
#pragma unroll
        for (int j = 0; j < RY + 2 * R; ++j) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + j * PITCH);
#pragma unroll
            for (int o = 0; o < RY; ++o) {
                const int k = j - o;
                if (k == 0) {                               // first tap initialises: no zeroing of 2 RY register pairs
                    acc[o][0] = mul2s(v.x, p.wy[0]);
                    acc[o][1] = mul2s(v.y, p.wy[0]);
                } else if (k > 0 && k <= 2 * R) {
                    acc[o][0] = fma2s(v.x, p.wy[k], acc[o][0]);
                    acc[o][1] = fma2s(v.y, p.wy[k], acc[o][1]);
                }
            }
        }
        float* dst = ybuf + ((g & 1) * G + yq) * YSLOT + (yhalf * RY) * PITCH + 4 * ycg;
#pragma unroll
        for (int o = 0; o < RY; ++o)
            *reinterpret_cast<ulonglong2*>(dst + o * PITCH) = make_ulonglong2(acc[o][0], acc[o][1]);
    };

    // ---- x pass + z scatter of one group, plane by plane; 2R+1 shifting z accumulators per column
    // Z accumulators: 2R+1 logical accumulators in 2R+1 + (G-1) physical slots.  Within a G-plane group
    // every plane updates in place (logical j of plane q lives in slot j + q: the shift costs nothing);
    // the LAST plane of the group writes logical j back to slot j, in ascending j so that no live value
    // is overwritten.  The loop-carried naming is therefore the same at the top of every group.  (With
    // 2R+1 slots the compiler had to rotate all of them back at the end of each group: ~17 MOV /
    // IMAD.MOV per thread and plane at R = 8, 13 % of the executed instructions, IMAD.MOV on the FMA
    // pipe.  17 phase-specialised loop bodies with zero moves were measured too: 207 KB of code thrash
    // the 32 KB instruction cache, 0.346 -> 0.568 ms.)
    constexpr int ZW = HAS_Z ? 2 * R + 1 : 1, ZS = HAS_Z ? ZW + G - 1 : 1;
    u64 zacc[ZS][2];
#pragma unroll
    for (int j = 0; j < ZS; ++j) zacc[j][0] = zacc[j][1] = 0ull;
    const int cg_o = tid % CGW, row_o = tid / CGW;
    const bool owner = row_o < ty && y0 + row_o >= 0 && x0 + 4 * cg_o < p.nx;
    // the output pointer of the NEXT finished plane, advanced by one plane per step (no 64-bit multiply per plane)
    float* out_ptr = p.out + (size_t)zb * plane_elems + (size_t)max(y0 + row_o, 0) * p.nx + x0 + 4 * cg_o;
    constexpr int EPI_AHEAD = 8;
    float4 old_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI && !HAS_Z && p.epilogue >= 2 && owner) old_next = __ldcg(reinterpret_cast<const float4*>(out_ptr));
    auto xzpass = [&](int g) {
        if (!owner) return;
        const int planes = min(G, n_planes - g * G);
#pragma unroll
        for (int q = 0; q < G; ++q) {
            if (q >= planes) break;
            const float* src = ybuf + ((g & 1) * G + q) * YSLOT + row_o * PITCH + 4 * cg_o;
            // gradient-magnitude accumulation: the running sum of squares of the voxel a step finishes is
            // requested one step ahead into a register and EPI_AHEAD planes ahead into L2 — loaded at the
            // store it exposed a DRAM round trip per plane (0.50 / 0.55 ms per accumulating launch against
            // 0.33 ms for the first on 512^3)
            float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI && p.epilogue >= 2) {
                const int fin = g * G + q - (HAS_Z ? 2 * R : 0);      // output plane (relative to zb) this step finishes
                old = old_next;
                if (fin + 1 >= 0 && fin + 1 < ze - zb)
                    old_next = __ldcg(reinterpret_cast<const float4*>(out_ptr + (fin >= 0 ? plane_elems : 0)));
                if (fin + EPI_AHEAD >= 0 && fin + EPI_AHEAD < ze - zb)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(out_ptr + (size_t)(fin >= 0 ? EPI_AHEAD : EPI_AHEAD + fin) * plane_elems));
            }
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
            const int idx = g * G + q;                     // plane index within this CTA's march
            if (HAS_Z) {
                const u64 v0 = pack2(xo[0], xo[1]), v1 = pack2(xo[2], xo[3]);
                constexpr int LASTQ = G - 1;
                const int fs = q == LASTQ ? 0 : q + 1;        // slot of the voxel this plane finishes
#pragma unroll
                for (int j = 0; j < 2 * R; ++j) {
                    const int src_slot = j + 1 + q, dst_slot = q == LASTQ ? j : j + 1 + q;
                    zacc[dst_slot][0] = fma2s(v0, p.wz[2 * R - j], zacc[src_slot][0]);
                    zacc[dst_slot][1] = fma2s(v1, p.wz[2 * R - j], zacc[src_slot][1]);
                }
                zacc[q == LASTQ ? 2 * R : 2 * R + 1 + q][0] = mul2s(v0, p.wz[0]);
                zacc[q == LASTQ ? 2 * R : 2 * R + 1 + q][1] = mul2s(v1, p.wz[0]);
                const int zo = zb + idx - 2 * R;           // finished output plane
                if (zo >= zb) {
                    float4* dst = reinterpret_cast<float4*>(out_ptr);
                    out_ptr += plane_elems;
                    if (!EPI) {
                        *reinterpret_cast<ulonglong2*>(dst) = make_ulonglong2(zacc[fs][0], zacc[fs][1]);
                    } else {
                        float4 v;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(zacc[fs][0]));
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(zacc[fs][1]));
                        *dst = epilogue(p.epilogue, v, old);
                    }
                }
            } else {
                float4* dst = reinterpret_cast<float4*>(out_ptr);
                out_ptr += plane_elems;
                const float4 v = make_float4(xo[0], xo[1], xo[2], xo[3]);
                *dst = EPI ? epilogue(p.epilogue, v, old) : v;
            }
        }
    };

    // ---- software pipeline over groups: raw and ybuf are double buffered, ONE CTA barrier per group.
    // phase k:  [patch warps: wait TMA (k+1), patch it, signal pdone] | x+z pass (k) | y pass (k+1) |
    //           barrier | issue TMA (k+3)
    // The TMA of a group is issued two barriers before its first use, the patch runs on the warps that
    // have no y-pass items while the others are busy with the x+z pass, and the y pass comes last in
    // the phase: edge handling and load latency both hide behind arithmetic.
    auto wait_and_patch = [&](int g) {                      // patch team, edge tiles only
        if (!patching || g >= n_groups) return;
        if ((tid >> 5) >= NPW && (tid >> 5) < pinfo[4]) return;   // a warp with both y items and output rows
#ifdef SEPFILT_DEBUG_CYCLES
        const long long tw0 = clock64();
#endif
        mbar_wait(&rfull[g & 1], (uint32_t)(g >> 1) & 1u);
#ifdef SEPFILT_DEBUG_CYCLES
        const long long tw1 = clock64();
#endif
        patch(g);
#ifdef SEPFILT_DEBUG_CYCLES
        if (tid == 0) { g_dbg_cycles[2048 + blockIdx.x] += tw1 - tw0; g_dbg_cycles[3072 + blockIdx.x] += clock64() - tw1; }
#endif
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(pdone);
    };
    if (tid == 0) { mbar_init(&rfull[0], 1); mbar_init(&rfull[1], 1); mbar_init(pdone, (uint32_t)(NPW + NW - max((ty * (TX / 4) + 31) / 32, NPW))); }
    __syncthreads();                                       // barriers + patch tables visible
    if (tid == 0) { issue(0); if (n_groups > 1) issue(1); }
    wait_and_patch(0);
    ypass(0);
    __syncthreads();
    if (tid == 0 && n_groups > 2) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(2); }
    for (int k = 0; k < n_groups; ++k) {
        wait_and_patch(k + 1);
        xzpass(k);
        if (k + 1 < n_groups) ypass(k + 1);
        __syncthreads();
        if (tid == 0 && k + 3 < n_groups) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(k + 3); }
    }
    if (HALO && tid == 0) ptx::halo_signal_done(hm, gridDim.x);   // every plane this CTA staged has landed and been read
#ifdef SEPFILT_DEBUG_CYCLES
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_dbg_cycles[blockIdx.x] = (clock64() - t_start) | ((long long)smid << 48);
    }
#endif
}

int radius_bucket(int r, bool has_z)
{
    // With a z pass, radius > 8 is NOT fused: 33 z accumulators x 4 columns force 256-thread CTAs on
    // 64-wide tiles (1.5x y-pass halo work, 2 warps per sub-partition) and measured 1.25 ms on 512^3
    // against 1.02 ms for three tiled passes.  Without one (stacks of 2-D images) the kernel carries no
    // z state and serves radius 12 and 16 with 2-plane groups.  (Tiled z pass + this kernel for the
    // y / x part of a 3-D filter: 1.04 ms on 512^3 sigma 4, no better than three tiled passes.)
    // one instantiation per radius, never a zero-padded wider one: 0 * NaN / 0 * Inf would spread a
    // non-finite sample beyond the true footprint of the filter (scipy keeps those neighbours finite)
    if (r >= 1 && r <= 8) return r;
    // radius 12 / 16 without a z pass (stacks of 2-D images) used to run here (launch_wide_xy): 0.706 ms for
    // sigma (0, 4, 4) on 512^3, against 0.242 + 0.299 ms for the two single-axis passes it replaces
    (void)has_z;
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

template <int R, bool HAS_Z, class C, bool EPI, bool HALO = false>
cudaError_t launch_e(FusedParams& p, cudaStream_t s);

// neighbour halos of the call in flight on this thread (set by launch_fused3d around its launches)
thread_local const sepfilt_halo* t_halo = nullptr;

template <int R, bool HAS_Z, class C>
cudaError_t launch_c(FusedParams& p, cudaStream_t s)
{
    if (t_halo) {
        if constexpr (HAS_Z) {
            if (!p.epilogue) return launch_e<R, HAS_Z, C, false, true>(p, s);
        }
        return cudaErrorInvalidValue;       // fused3d_supported admits halos only for plain filters with a z pass
    }
    return p.epilogue ? launch_e<R, HAS_Z, C, true>(p, s) : launch_e<R, HAS_Z, C, false>(p, s);
}

template <int R, bool HAS_Z, class C, bool EPI, bool HALO>
cudaError_t launch_e(FusedParams& p, cudaStream_t s)
{
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
    ptx::HaloMaps hm;
    memset(&hm, 0, sizeof hm);
    if (HALO && t_halo) {
        const sepfilt_halo& h = *t_halo;
        const void* src[2] = {h.lo, h.hi};
        const int planes[2] = {h.lo ? h.planes_lo : 0, h.hi ? h.planes_hi : 0};
        CUtensorMap* maps[2] = {&hm.lo, &hm.hi};
        for (int i = 0; i < 2; ++i) {
            if (!planes[i]) continue;
            const cuuint64_t hdim[3] = {(cuuint64_t)p.nx, (cuuint64_t)p.ny, (cuuint64_t)planes[i]};
            if (enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(src[i]), hdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return cudaErrorInvalidValue;
        }
        hm.planes_lo = planes[0]; hm.planes_hi = planes[1];
        hm.ready_lo = h.ready_lo; hm.ready_hi = h.ready_hi; hm.epoch = h.epoch;
        hm.done_lo = h.done_lo; hm.done_hi = h.done_hi; hm.counter = h.cta_counter;
    }
    auto kern = fused3d_kernel<R, HAS_Z, C, EPI, HALO>;
    // the attribute is per (function, device): set once per device, not on every call
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.nzseg;
    kern<<<(unsigned)blocks, C::NT, C::SMEM, s>>>(p, tmap, hm);
    return cudaGetLastError();
}

}  // namespace

bool fused3d_supported(const FusedVolume& v, const F32Taps taps[3], bool gradmag)
{
    int r = 0;
    for (int a = 0; a < 3; ++a) r = taps[a].radius > r ? taps[a].radius : r;
    const bool z_pass = !(taps[0].radius == 0 && taps[0].w[0] == 1.0f);
    if (radius_bucket(r, z_pass || gradmag) < 0) return false;   // gradient magnitude: one launch per axis, epilogue variants exist up to radius 8
    // every filtered axis must have exactly the kernel's radius (no zero-padded taps, see radius_bucket); an
    // unfiltered z axis has its own instantiation (HAS_Z = false), an unfiltered y or x axis has not
    for (int a = 0; a < 3; ++a) {
        const bool identity = taps[a].radius == 0 && taps[a].w[0] == 1.0f;
        if (a == 0 && identity) continue;
        if (taps[a].radius != r) return false;
    }
    if (v.nx % 4 != 0 || (reinterpret_cast<uintptr_t>(v.in) & 15) || (reinterpret_cast<uintptr_t>(v.out) & 15))
        return false;
    if (v.nx < 1 || v.ny < 1 || v.nz_in < 1 || v.nz_out < 1) return false;
    const bool has_z = !(taps[0].radius == 0 && taps[0].w[0] == 1.0f);
    if (!has_z && v.nz_out < 4) return false;   // a lone 2-D image cannot fill a plane group: per-axis passes
    // constant mode with cval != 0 is affine, not linear: the passes no longer commute and the
    // halo of an intermediate is cval, not filtered cval -> per-axis passes in scipy's order
    if (v.cval != 0.f)
        for (int a = 0; a < 3; ++a)
            if (v.mode[a] == SEPFILT_CONSTANT && taps[a].radius > 0) return false;
    // wrap along y or x: the source of a halo cell lies at the far side of the array, outside the tile,
    // and the patch falls back to one global load per cell — 2.7 ms on 512^3 against 0.85 ms for three
    // tiled passes.  (wrap along z is free: the TMA plane coordinate is remapped.)
    for (int a = 1; a < 3; ++a)
        if (v.mode[a] == SEPFILT_WRAP && taps[a].radius > 0) return false;
    const long long tiles = (long long)((v.nx + 63) / 64) * v.ny;
    if (tiles * v.nz_out > 2147483647LL) return false;
    if (v.halo) {
        // neighbour planes come through their own tensor maps: whole slab, a z pass, sources of every
        // x / y halo cell inside the tile, one launch (the accumulating gradient-magnitude launches are not
        // instantiated with halos: fused_ws serves that call)
        if (!has_z || gradmag || v.z_offset < 0 || v.z_offset + v.nz_out > v.nz_in || v.ny < 32 || v.nx < 32) return false;
        if ((v.halo->lo && v.halo->planes_lo < taps[0].radius) || (v.halo->hi && v.halo->planes_hi < taps[0].radius))
            return false;
        if (v.nz_in < taps[0].radius) return false;
    }
    return true;
}

namespace {

// tile rows / z segments for a tile width: fill the SMs with whole waves where the shape allows.
// Returns the modelled cost (arbitrary units) so that the caller can compare tile widths.
double plan_tiles(const FusedVolume& v, int R, bool has_z, int tx, int slots, FusedParams* p)
{
    const int tiles_x = (v.nx + tx - 1) / tx;
    int best_ty = TYM, best_seg = 1;
    double best_cost = 1e300;
    for (int ty = TYM; ty >= 8; --ty) {
        const int tiles_y = (v.ny + ty - 1) / ty;
        for (int nseg = 1; nseg <= 64 && nseg <= v.nz_out; ++nseg) {
            const int zseg = (v.nz_out + nseg - 1) / nseg;
            const long long ctas = (long long)tiles_x * tiles_y * nseg;
            const long long waves = (ctas + slots - 1) / slots;
            // per-CTA time ~ planes marched x (y pass on a full 16-row tile incl. the x halo + x/z on ty rows)
            const double per_cta = (double)(zseg + (has_z ? 2 * R : 0)) *
                                   (0.33 * TYM * (tx + 2 * R) + 0.67 * ty * tx);
            const double cost = waves * per_cta;
            if (cost < best_cost) { best_cost = cost; best_ty = ty; best_seg = nseg; }
            if (!has_z) break;
        }
    }
    static const char* fty = getenv("SEPFILT_FUSED_TY");         // tuning aid (read once): force the tile height
    if (fty) {
        const int t = atoi(fty);
        if (t >= 1 && t <= TYM) { best_ty = t; best_seg = 1; }
    }
    if (v.ny <= TYM) best_ty = v.ny < 1 ? 1 : (v.ny < TYM ? v.ny : TYM);
    if (p) {
        p->tiles_x = tiles_x;
        p->ty = best_ty;
        p->tiles_y = (v.ny + best_ty - 1) / best_ty;
        const int over = p->tiles_y * best_ty - v.ny;
        p->yshift = p->tiles_y >= 2 ? over / 2 : 0;
        p->zseg = (v.nz_out + best_seg - 1) / best_seg;
        p->nzseg = (v.nz_out + p->zseg - 1) / p->zseg;
    }
    return best_cost;
}

template <int R, bool HAS_Z>
cudaError_t launch_r(const FusedVolume& v, FusedParams& p, cudaStream_t s)
{
    const int sms = cached_sm_count();
    {
    using Wide = Cfg<R, 128, 4, 1>;      // 512 threads, one CTA per SM
    using Narrow = Cfg<R, 64, 3, 2>;     // 256 threads, two CTAs per SM: one CTA's barrier waits hide behind the other
    static const char* force = getenv("SEPFILT_FUSED_TILE");
    const double cw = plan_tiles(v, R, HAS_Z, 128, sms, nullptr);
    const double cn = plan_tiles(v, R, HAS_Z, 64, 2 * sms, nullptr) * 2.0;   // two CTAs share an SM
    bool narrow = cn < cw;
    if (force) narrow = force[0] == 'n';
    if (narrow) {
        plan_tiles(v, R, HAS_Z, 64, 2 * sms, &p);
        return launch_c<R, HAS_Z, Narrow>(p, s);
    }
    plan_tiles(v, R, HAS_Z, 128, sms, &p);
    // tiles of <= 14 rows (e.g. 512 rows = 37 x 14 on 148 SMs): 7-row y-pass register tiles, so that the
    // y pass does not filter two rows per tile that nobody reads
    if (p.ty <= 14 && p.ty > 8) return launch_c<R, HAS_Z, Cfg<R, 128, 4, 1, 7>>(p, s);
    return launch_c<R, HAS_Z, Wide>(p, s);
    }
}

}  // namespace

// y + x only, radius 12 / 16: one configuration (128-wide tiles, 2-plane groups, no epilogue variants)
static cudaError_t launch_one(const FusedVolume& v, const F32Taps taps[3], int epilogue_mode, cudaStream_t s)
{
    int r = 0;
    for (int a = 0; a < 3; ++a) r = taps[a].radius > r ? taps[a].radius : r;
    const bool has_z = !(taps[0].radius == 0 && taps[0].w[0] == 1.0f);
    const int R = radius_bucket(r, has_z || epilogue_mode != 0);
    if (R < 0) return cudaErrorInvalidValue;

    FusedParams p;
    p.in = v.in; p.out = v.out;
    p.nz_in = v.nz_in; p.nz_out = v.nz_out; p.ny = v.ny; p.nx = v.nx; p.z_offset = v.z_offset;
    p.mode_z = v.mode[0]; p.mode_y = v.mode[1]; p.mode_x = v.mode[2];
    p.cval = v.cval;
    p.epilogue = epilogue_mode;
    p.pad_ = 0; p.yshift = 0;
    recentre(taps[0], R, p.wz);
    recentre(taps[1], R, p.wy);
    recentre(taps[2], R, p.wx);

    switch (R * 2 + (has_z ? 1 : 0)) {
    case 1 * 2 + 0: return launch_r<1, false>(v, p, s);
    case 1 * 2 + 1: return launch_r<1, true>(v, p, s);
    case 2 * 2 + 0: return launch_r<2, false>(v, p, s);
    case 2 * 2 + 1: return launch_r<2, true>(v, p, s);
    case 3 * 2 + 0: return launch_r<3, false>(v, p, s);
    case 3 * 2 + 1: return launch_r<3, true>(v, p, s);
    case 4 * 2 + 0: return launch_r<4, false>(v, p, s);
    case 4 * 2 + 1: return launch_r<4, true>(v, p, s);
    case 5 * 2 + 0: return launch_r<5, false>(v, p, s);
    case 5 * 2 + 1: return launch_r<5, true>(v, p, s);
    case 6 * 2 + 0: return launch_r<6, false>(v, p, s);
    case 6 * 2 + 1: return launch_r<6, true>(v, p, s);
    case 7 * 2 + 0: return launch_r<7, false>(v, p, s);
    case 7 * 2 + 1: return launch_r<7, true>(v, p, s);
    case 8 * 2 + 0: return launch_r<8, false>(v, p, s);
    case 8 * 2 + 1: return launch_r<8, true>(v, p, s);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fused3d(const FusedVolume& v, const F32Taps taps[3], const F32Taps dtaps[3], bool gradmag,
                           cudaStream_t s)
{
    struct HaloScope {
        explicit HaloScope(const sepfilt_halo* h) { t_halo = h; }
        ~HaloScope() { t_halo = nullptr; }
    } scope(v.halo);
    if (!gradmag) return launch_one(v, taps, 0, s);
    // gradient magnitude: sqrt(sum_a (D_a prod_{b != a} S_b in)^2), one launch per filtered axis
    const int first = (v.nz_in == 1 && v.nz_out == 1 && taps[0].radius == 0 && dtaps[0].radius == 0) ? 1 : 0;
    for (int a = first; a < 3; ++a) {
        F32Taps t[3] = {taps[0], taps[1], taps[2]};
        t[a] = dtaps[a];
        const int mode = a == first ? 1 : (a == 2 ? 3 : 2);
        cudaError_t e = launch_one(v, t, first == 2 ? 3 : mode, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace sepfilt

#ifdef SEPFILT_DEBUG_CYCLES
extern "C" __attribute__((visibility("default"))) int sepfilt_debug_cycles(long long* host, int n)
{
    return (int)cudaMemcpyFromSymbol(host, sepfilt::g_dbg_cycles, sizeof(long long) * n);
}
#endif
