#pragma once
// fused_ws.cuh — warp-specialised fused separable f32 filter (round 2): the volume crosses HBM once in
// and once out (the reference launches one ElementwiseKernel per axis plus copy-backs:
// filters.py:651-662 / :777-789, _filters_core.py:148-155), and — unlike fused3d.cu — no warp ever waits
// at a CTA barrier: the passes run as a producer / consumer pipeline of warps with their own
// register budgets.
//
// One CTA (12 warps) owns a 128 x TY column of the volume and marches along z.
//   Y warps  (warps 0-3, one per SM sub-partition, plus warp 11 on 14-row tiles; 72-104 registers): take the next input plane from a
//            shared counter, wait for its TMA box (cp.async.bulk.tensor.3d, mbarrier complete_tx), run
//            the y pass for ALL rows of the tile from registers (a lane owns 4 columns: every staged
//            row is read from shared memory once; packed fma.rn.f32x2) and publish the y-filtered
//            plane through an mbarrier.  The warp that drained a raw slot re-issues the TMA load of
//            the plane NR steps ahead into it: no issue thread, no CTA-wide barrier.  Tile-edge
//            handling is warp-local: out-of-array rows are copied inside the warp's own raw plane,
//            out-of-array columns inside its y-filtered plane (x extension commutes with the y pass).
//   XZ warps (warps 4-11, 200-208 registers): per plane a thread loads the window of its 8 (or 4)
//            outputs from the y-filtered plane, runs the x pass in registers and scatters the result
//            into 2R+1 shifting per-column z accumulators (packed FFMA2); a finished voxel leaves
//            with a 16-byte store.  Gradient magnitude keeps three accumulator sets (G'z Gy Gx,
//            Gz G'y Gx, Gz Gy G'x share the two y passes and three x passes) and applies the
//            square / sum / sqrt of filters.py:1187-1201 in the store: ONE launch, 8 B / voxel.
// The sub-partition whose XZ warps have less to do lets its Y warp run faster, and that warp then takes
// more planes from the counter: the roles balance themselves.  Tensor cores are deliberately not used
// (FP32-bound stencil, TF32 would break rtol 1e-5).
#include <cuda.h>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace sepfilt {

namespace ws {

using namespace ptx;

constexpr int WS_MAXR = 16;
constexpr int WS_TAPS = 2 * WS_MAXR + 2;        // 34 floats: every tap array starts at the same offset mod 8
constexpr int WS_SMEM_BUDGET = 232448;           // 227 KiB: the per-CTA maximum on sm_100

struct WsParams {
    const float* in;
    float*       out;
    int nz_in, nz_out, ny, nx, z_offset;
    int mode_z, mode_y, mode_x;
    int tiles_x, tiles_y, yshift;               // tile row t covers rows [t * TYC - yshift, (t + 1) * TYC - yshift)
    int zseg, nzseg;                            // output planes per z segment, number of segments
    int pad_[2];                                // keeps the tap arrays at 4 (mod 8) bytes, see the static_assert below
    float wz[WS_TAPS], wy[WS_TAPS], wx[WS_TAPS];    // taps at offsets -R..R (exact radius: no zero padding)
    float dz[WS_TAPS], dy[WS_TAPS], dx[WS_TAPS];    // derivative taps (gradient magnitude)
};

// Same finding as fused3d.cu: with the tap arrays at 4 (mod 8) bytes in the kernel parameter block ptxas feeds the
// packed FFMA2 its scalar tap straight from the constant bank (measured: 0.3206 -> 0.3159 ms on 512^3 sigma 2)
static_assert(offsetof(WsParams, wz) % 8 == 4 && (WS_TAPS * sizeof(float)) % 8 == 0, "keep the tap arrays at 4 (mod 8) bytes");

__host__ __device__ constexpr int rup4(int r) { return (r + 3) & ~3; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }

// R: radius (exact), CPT: columns per XZ thread (8: 16 threads per tile row, 4: 32), TYC: tile rows,
// HAS_Z: z pass present, GRAD: gradient magnitude (2 y-filtered planes, 3 z accumulator sets)
// YSPLIT: the Y warps filter the tile rows in this many chunks (fewer live accumulators, more row loads)
template <int R_, int CPT_, int TYC_, bool HAS_Z_, bool GRAD_, int YREGS_, int XZREGS_, int YSPLIT_ = 1>
struct WsCfg {
    static constexpr int R = R_, CPT = CPT_, TYC = TYC_, YREGS = YREGS_, XZREGS = XZREGS_, YSPLIT = YSPLIT_;
    static constexpr int YCH = TYC_ / YSPLIT_;                  // rows per chunk
    static constexpr bool HAS_Z = HAS_Z_, GRAD = GRAD_;
    static constexpr int TX = 128, TPR = TX / CPT;
    static constexpr int NT = 384, NYW = 4;
    static constexpr int NXZW = (TPR * TYC + 31) / 32;          // XZ warps with work (<= 8)
    // A warp slot without XZ rows (14-row tiles: warp 11) runs as one more Y warp.  ncu (round 2) showed the XZ
    // warps waiting for y-filtered planes 22 % of their time — the Y warps are the critical path — and warp 11 sits
    // on the sub-partition that has one XZ warp less: 512^3 sigma 2 reflect 0.310 -> 0.282 ms.  More warps do not
    // fit: a sub-partition's register file holds 512 registers per lane (Y 104 + 2 x XZ 200 = 504).
    // (Measured and rejected in the same session: an x pass on (row a, row b) pairs read from a row-interleaved
    //  y plane — no re-paired window copies, 17 % fewer instructions overall, but the interleaving MOVs land on
    //  the Y warps: 0.292 ms.  The same x pass in the gradient-magnitude kernel, whose XZ warps bind (140 -> 34 MOV per
    //  two planes): 0.709 -> 0.728 ms on 512^3 sigma 1.5 — the instruction count is not what limits these kernels, the
    //  dependent FFMA2 chains of 2-3 resident warps per sub-partition are.)
    static constexpr int NYTOT = NYW + (8 - NXZW);
    static constexpr int HL = rup4(R), PW = TX + 2 * HL, NCG = PW / 4;
    static constexpr int BOX_ROWS = TYC + 2 * R;
    static constexpr int NF = GRAD ? 2 : 1, NZF = GRAD ? 3 : 1;
    // tail of the y pass: the 2 HL columns beyond the 32 lanes x 4 columns of the main part
    static constexpr int NPART = 32 / HL, RPP = (TYC + NPART - 1) / NPART;
    static constexpr int RAW_ROWS = cmax(TYC, NPART * RPP) + 2 * R;
    static constexpr int RSLOT = (RAW_ROWS * PW + 31) & ~31;    // floats per raw plane slot (128 B multiple)
    // y-filtered plane: with 8 columns per XZ thread the lanes of a quarter warp read float4 groups
    // 2 apart; even groups are stored in the first half of a row and odd groups from ODD_OFF on
    // (ODD_OFF = 4 mod 8 groups), which makes both the XZ loads and the Y stores conflict-free
    static constexpr bool SWZ = CPT == 8;
    static constexpr int ODD_OFF = SWZ ? ((NCG / 2 + 3) / 8) * 8 + 4 : 0;
    static constexpr int YPG = SWZ ? ODD_OFF + NCG / 2 : NCG;   // float4 groups per row
    static constexpr int YP = 4 * YPG;
    static constexpr int YSLOT = NF * TYC * YP;
    static constexpr int NY = 8;
    static constexpr int NR = cmin(12, (WS_SMEM_BUDGET - NY * YSLOT * 4 - 1024) / (RSLOT * 4));
    static constexpr size_t SMEM = sizeof(float) * ((size_t)NR * RSLOT + (size_t)NY * YSLOT) + 1024;
    static constexpr int G = 2;                                 // planes per unrolled group of the XZ loop
    static constexpr int ZW = HAS_Z ? 2 * R + 1 : 1, ZS = HAS_Z ? ZW + G - 1 : 1;
    static_assert(ODD_OFF >= NCG / 2 || !SWZ, "odd groups overlap the even ones");
    static_assert(NR >= 5, "too few raw slots");
    static_assert(NY % G == 0, "slot arithmetic assumes NY % G == 0");
    static_assert(NXZW <= 8 && NXZW >= 1, "XZ warps");
    // setmaxnreg moves registers inside the CTA's launch allocation: 12 warps x 168 (= 65536 / 384 rounded
    // down to 8); a budget beyond it makes setmaxnreg.inc wait forever
    static_assert(NYTOT * YREGS_ + NXZW * XZREGS_ <= 12 * 168, "register budget exceeds the launch allocation");
    // ... and registers only move inside an SM sub-partition (warps w, w + 4, w + 8: one Y and two XZ warps, 3 x 168 at
    // launch): Y 112 / XZ 200 passes the CTA-wide test above and hangs in setmaxnreg.inc (measured: a 15-minute timeout)
    static_assert(NXZW < 5 || YREGS_ + 2 * XZREGS_ <= 3 * 168, "register budget exceeds the sub-partition's launch allocation");
    static_assert(TYC_ % YSPLIT_ == 0, "row chunks");
    static_assert(WS_SMEM_BUDGET >= (int)SMEM, "shared memory");
};

template <bool SWZ, int ODD_OFF> __device__ __forceinline__ int ycol_offset(int c)
{
    // float offset inside a y-filtered row of staged column c
    if (!SWZ) return c;
    const int g = c >> 2;
    return 4 * ((g >> 1) + (g & 1) * ODD_OFF) + (c & 3);
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
fws_kernel(const __grid_constant__ WsParams p, const __grid_constant__ CUtensorMap tmap,
           const __grid_constant__ HaloMaps hm)
{
    constexpr int R = C::R, CPT = C::CPT, TYC = C::TYC, TX = C::TX, TPR = C::TPR;
    constexpr int HL = C::HL, PW = C::PW, NCG = C::NCG, NR = C::NR, NY = C::NY, NF = C::NF, NZF = C::NZF;
    constexpr int RSLOT = C::RSLOT, YSLOT = C::YSLOT, YP = C::YP, G = C::G;
    constexpr bool HAS_Z = C::HAS_Z, GRAD = C::GRAD, SWZ = C::SWZ;
    constexpr uint32_t BOX_BYTES = (uint32_t)(PW * C::BOX_ROWS * sizeof(float));

    extern __shared__ __align__(128) float smem[];
    float* raw = smem;                                          // [NR][RSLOT]
    float* ybuf = smem + NR * RSLOT;                            // [NY][NF][TYC][YP]
    uint64_t* full_raw = reinterpret_cast<uint64_t*>(ybuf + NY * YSLOT);   // [NR] TMA landed
    uint64_t* full_y = full_raw + NR;                           // [NY] y-filtered plane published
    uint64_t* empty_y = full_y + NY;                            // [NY] every XZ warp has read the plane
    int* meta = reinterpret_cast<int*>(empty_y + NY);           // [0] next plane; [4..4+2R) rows; [4+2R..) cols
    int* rowtab = meta + 4;
    int* coltab = rowtab + C::RAW_ROWS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int b = blockIdx.x;
    const int tile_x = b % p.tiles_x; b /= p.tiles_x;
    const int tile_y = b % p.tiles_y; b /= p.tiles_y;
    const int seg = b;
    const int x0 = tile_x * TX, y0 = tile_y * TYC - p.yshift;
    const int zb = seg * p.zseg, ze = min(zb + p.zseg, p.nz_out);
    const int p_first = zb + p.z_offset - (HAS_Z ? R : 0);      // first input plane (input coordinates)
    const int n_planes = (ze - zb) + (HAS_Z ? 2 * R : 0);

    // ---- tile-edge tables (warp 0): staged rows / columns outside the array whose value a stored voxel
    //      can read, with the staged row / column holding their remapped source (_util.py:170-228);
    //      constant mode keeps TMA's zero fill
    const int row_lo = max(y0, 0), row_hi = min(y0 + TYC, p.ny);    // stored rows [row_lo, row_hi)
    int ncol = 0;
    const bool ytab = p.mode_y != SEPFILT_CONSTANT && (y0 - R < 0 || y0 + TYC + R > p.ny);
    if (p.mode_x != SEPFILT_CONSTANT) ncol = max(0, R - x0) + max(0, min(x0 + TX, p.nx) + R - p.nx);
    if (warp == 0) {
        // rowtab[jj]: float offset (inside a raw plane slot) of the staged row the y pass reads in place of staged
        // row jj: the row itself when it lies inside the array (or the mode is constant: TMA's zero fill stays),
        // else the staged row holding its remapped source — out-of-array rows are never patched, they are simply
        // not read.  Only tiles that touch the first / last array row read through the table (ytab).
        for (int jj = lane; jj < C::RAW_ROWS; jj += 32) {
            const int gy = y0 - R + jj;
            int sj = jj;
            if (p.mode_y != SEPFILT_CONSTANT && jj < C::BOX_ROWS && (gy < 0 || gy >= p.ny)) {
                const int m = remap_index32(p.mode_y, gy, p.ny) - (y0 - R);
                if (m >= 0 && m < C::BOX_ROWS) sj = m;
            }
            rowtab[jj] = sj * PW;
        }
        const int left = max(0, R - x0);
        for (int i = lane; i < ncol; i += 32) {
            const int gx = i < left ? i - left : p.nx + (i - left);
            const int sx = remap_index32(p.mode_x, gx, p.nx);
            coltab[i] = ycol_offset<SWZ, C::ODD_OFF>(gx - (x0 - HL)) | (ycol_offset<SWZ, C::ODD_OFF>(sx - (x0 - HL)) << 16);
        }
        if (lane == 0) {
            meta[0] = 0; meta[1] = 0;
            for (int i = 0; i < NR; ++i) mbar_init(&full_raw[i], 1);
            for (int i = 0; i < NY; ++i) { mbar_init(&full_y[i], 1); mbar_init(&empty_y[i], C::NXZW); }
            mbar_fence_init();
            tma_prefetch_desc(&tmap);
        }
    }
    __syncthreads();

    if (warp < C::NYW || warp >= C::NYW + C::NXZW) {
        // =============================== Y warps ===============================
        reg_dealloc<C::YREGS>();
        auto issue = [&](int pl) {                               // one lane
            int pz = p_first + pl;
            const CUtensorMap* map = &tmap;
            if (pz < 0 && hm.planes_lo) {
                // a plane of the lower neighbour's slab, read in place over NVLink once its array is ready
                if (hm.ready_lo) { wait_flag_geq(hm.ready_lo, hm.epoch); fence_proxy_async_all(); }
                map = &hm.lo; pz += hm.planes_lo;
            } else if (pz >= p.nz_in && hm.planes_hi) {
                if (hm.ready_hi) { wait_flag_geq(hm.ready_hi, hm.epoch); fence_proxy_async_all(); }
                map = &hm.hi; pz -= p.nz_in;
            } else if (p.mode_z != SEPFILT_CONSTANT) {
                pz = remap_index32(p.mode_z, pz, p.nz_in);   // constant: OOB box -> zeros
            }
            uint64_t* bar = &full_raw[pl % NR];
            mbar_expect_tx(bar, BOX_BYTES);
            tma_load_3d(raw + (pl % NR) * RSLOT, map, x0 - HL, y0 - R, pz, bar);
        };
        if (tid == 0)
            for (int pl = 0; pl < NR && pl < n_planes; ++pl) issue(pl);
        // column patch list of this lane: destination | source << 16 (float offsets inside a y slot), -1 = none
        // (a tile that touches BOTH x ends of a narrow array can have more cells than 4 per lane: generic loop below)
        static_assert(YSLOT < 32768, "patch offsets are packed into 16 bits");
        const bool cfast = ncol * TYC * NF <= 128;
        int cpatch[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = lane + 32 * u;
            cpatch[u] = -1;
            if (cfast && i < ncol * TYC * NF) {
                const int r = i / ncol, c = i - r * ncol;
                const int e = coltab[c];
                cpatch[u] = (r * YP + (e & 0xffff)) | ((r * YP + (e >> 16)) << 16);
            }
        }
        // (Measured and rejected: reading the plane counter one plane ahead to hide the atomic's latency — a reserved
        //  but unstarted plane stalls the in-order XZ warps: 512^3 sigma 2 0.287 -> 0.298 ms.)
        for (;;) {
            int pl = 0;
            if (lane == 0) pl = atomicAdd(&meta[0], 1);
            pl = __shfl_sync(0xffffffffu, pl, 0);
            if (pl >= n_planes) break;
            const int rs = pl % NR, ys = pl % NY;
            float* rawp = raw + rs * RSLOT;
            float* yp = ybuf + ys * YSLOT;
            mbar_wait(&full_raw[rs], (uint32_t)(pl / NR) & 1u);
            mbar_wait(&empty_y[ys], ((uint32_t)(pl / NY) & 1u) ^ 1u);
            // ---- main part: lane owns staged columns [4 lane, 4 lane + 4), all TYC rows (YSPLIT chunks).
            //      TAB: every row is read through rowtab (tiles touching the first / last array row)
            auto ypass = [&]<bool TAB>() {
#pragma unroll
            for (int h = 0; h < C::YSPLIT; ++h) {
                constexpr int YCH = C::YCH;
                u64 acc[NF][YCH][2];
                const float* src = rawp + 4 * lane + h * YCH * PW;
#pragma unroll
                for (int j = 0; j < YCH + 2 * R; ++j) {
                    const int jj = h * YCH + j;                 // staged row (compile time)
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(TAB ? rawp + 4 * lane + rowtab[jj] : src + j * PW);
#pragma unroll
                    for (int o = 0; o < YCH; ++o) {
                        const int k = j - o;
                        if (k == 0) {                          // first tap initialises: no zeroing
                            acc[0][o][0] = mul2s(v.x, p.wy[0]);
                            acc[0][o][1] = mul2s(v.y, p.wy[0]);
                            if (GRAD) { acc[NF - 1][o][0] = mul2s(v.x, p.dy[0]); acc[NF - 1][o][1] = mul2s(v.y, p.dy[0]); }
                        } else if (k > 0 && k <= 2 * R) {
                            acc[0][o][0] = fma2s(v.x, p.wy[k], acc[0][o][0]);
                            acc[0][o][1] = fma2s(v.y, p.wy[k], acc[0][o][1]);
                            if (GRAD) {
                                acc[NF - 1][o][0] = fma2s(v.x, p.dy[k], acc[NF - 1][o][0]);
                                acc[NF - 1][o][1] = fma2s(v.y, p.dy[k], acc[NF - 1][o][1]);
                            }
                        }
                    }
                }
                float* dst = yp + ycol_offset<SWZ, C::ODD_OFF>(4 * lane) + h * YCH * YP;
#pragma unroll
                for (int f = 0; f < NF; ++f)
#pragma unroll
                    for (int o = 0; o < YCH; ++o)
                        *reinterpret_cast<ulonglong2*>(dst + (f * TYC + o) * YP) = make_ulonglong2(acc[f][o][0], acc[f][o][1]);
            }
            // ---- tail: staged columns [128, 128 + 2 HL): lane = (column pair, row part)
            {
                constexpr int RPP = C::RPP;
                const int cp = lane % HL, part = lane / HL;
                u64 acc[NF][RPP];
                const float* src = rawp + (part * RPP) * PW + 128 + 2 * cp;
#pragma unroll
                for (int j = 0; j < RPP + 2 * R; ++j) {
                    const u64 v = *reinterpret_cast<const u64*>(TAB ? rawp + 128 + 2 * cp + rowtab[part * RPP + j] : src + j * PW);
#pragma unroll
                    for (int o = 0; o < RPP; ++o) {
                        const int k = j - o;
                        if (k == 0) {
                            acc[0][o] = mul2s(v, p.wy[0]);
                            if (GRAD) acc[NF - 1][o] = mul2s(v, p.dy[0]);
                        } else if (k > 0 && k <= 2 * R) {
                            acc[0][o] = fma2s(v, p.wy[k], acc[0][o]);
                            if (GRAD) acc[NF - 1][o] = fma2s(v, p.dy[k], acc[NF - 1][o]);
                        }
                    }
                }
                float* dst = yp + ycol_offset<SWZ, C::ODD_OFF>(128 + 2 * cp);
#pragma unroll
                for (int f = 0; f < NF; ++f)
#pragma unroll
                    for (int o = 0; o < RPP; ++o)
                        if (part * RPP + o < TYC) *reinterpret_cast<u64*>(dst + (f * TYC + part * RPP + o) * YP) = acc[f][o];
            }
            };
            if (ytab) ypass.template operator()<true>(); else ypass.template operator()<false>();
            __syncwarp();
            // the raw slot is drained: load the plane NR steps ahead into it
            if (lane == 0 && pl + NR < n_planes) {
                fence_proxy_async();
                issue(pl + NR);
            }
            if (ncol && !cfast) {
                const int total = ncol * TYC * NF;
                for (int i = lane; i < total; i += 32) {
                    const int r = i / ncol, c = i - r * ncol;
                    const int e = coltab[c];
                    yp[r * YP + (e & 0xffff)] = yp[r * YP + (e >> 16)];
                }
                __syncwarp();
            } else if (ncol) {
                // out-of-array columns of the y-filtered plane(s) <- their source columns: the lane's (destination,
                // source) offsets were tabulated once, so a plane costs four independent LDS + STS per lane
                float t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (cpatch[u] >= 0) t[u] = yp[cpatch[u] >> 16];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (cpatch[u] >= 0) yp[cpatch[u] & 0xffff] = t[u];
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&full_y[ys]);
        }
        // the last Y warp of the CTA to run out of planes: every staged plane has landed and been filtered
        if (lane == 0 && (hm.planes_lo | hm.planes_hi) && atomicAdd(&meta[1], 1) == C::NYTOT - 1)
            halo_signal_done(hm, gridDim.x);
        return;
    }

    // =============================== XZ warps ===============================
    reg_alloc<C::XZREGS>();
    const int t = tid - C::NYW * 32;
    const int row = t / TPR, oct = t - row * TPR;
    const bool row_ok = row < TYC && y0 + row >= 0 && y0 + row < p.ny;
    const bool ok0 = row_ok && x0 + CPT * oct < p.nx;
    const bool ok1 = CPT == 8 && row_ok && x0 + CPT * oct + 4 < p.nx;
    const size_t plane_elems = (size_t)p.ny * p.nx;
    float* out_ptr = p.out + (size_t)zb * plane_elems + (size_t)min(max(y0 + row, 0), p.ny - 1) * p.nx + x0 + CPT * oct;
    const int yoff = (row < TYC ? row : 0) * YP + (SWZ ? 4 * oct : CPT * oct);   // window start inside a y plane

    constexpr int NP = CPT / 2;                                 // packed column pairs per thread
    constexpr int ZS = C::ZS;
    u64 zacc[NZF][ZS][NP];
#pragma unroll
    for (int f = 0; f < NZF; ++f)
#pragma unroll
        for (int j = 0; j < ZS; ++j)
#pragma unroll
            for (int c = 0; c < NP; ++c) zacc[f][j][c] = 0ull;

    constexpr int WIN = CPT + 2 * HL;
    auto load_window = [&](const float* src, float (&win)[WIN]) {
        if (SWZ) {
#pragma unroll
            for (int m = 0; m < WIN / 8; ++m) {
                const float4 e = *reinterpret_cast<const float4*>(src + 4 * m);
                const float4 o = *reinterpret_cast<const float4*>(src + 4 * (C::ODD_OFF + m));
                win[8 * m] = e.x; win[8 * m + 1] = e.y; win[8 * m + 2] = e.z; win[8 * m + 3] = e.w;
                win[8 * m + 4] = o.x; win[8 * m + 5] = o.y; win[8 * m + 6] = o.z; win[8 * m + 7] = o.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < WIN / 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
                win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
            }
        }
    };
    // (Measured and rejected: applying the taps in the order of the window pair they read, so that only one
    //  re-paired copy is live at a time — ncu counts 161 MOV against 136 FFMA2 per pair of planes in this loop,
    //  ptxas re-materialises the copies under register pressure — 0.3159 -> 0.3219 ms: fewer MOVs, worse ILP.)
    // x pass on packed column pairs (FFMA2): output pair j needs the input pairs starting at window index
    // 2j + (HL - R) + k; even starts are the aligned pairs of the window, odd starts are re-paired copies
    // (two MOVs each, once per window, shared by every tap set) — half the issue slots of a scalar FFMA pass,
    // same products, same summation order
    auto pair_window = [&](const float (&win)[WIN], u64 (&pe)[WIN / 2], u64 (&po)[WIN / 2]) {
#pragma unroll
        for (int m = 0; m < WIN / 2; ++m) pe[m] = pack2(win[2 * m], win[2 * m + 1]);
#pragma unroll
        for (int m = 0; m + 1 < WIN / 2; ++m) po[m] = pack2(win[2 * m + 1], win[2 * m + 2]);
        po[WIN / 2 - 1] = 0ull;
    };
    auto xpass = [&](const u64 (&pe)[WIN / 2], const u64 (&po)[WIN / 2], const float* w, u64 (&o)[NP]) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            constexpr int off = HL - R;
            const int i0 = 2 * j + off;
            o[j] = mul2s((i0 & 1) ? po[i0 >> 1] : pe[i0 >> 1], w[0]);
        }
#pragma unroll
        for (int k = 1; k <= 2 * R; ++k)
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = 2 * j + (HL - R) + k;
                o[j] = fma2s((i & 1) ? po[i >> 1] : pe[i >> 1], w[k], o[j]);
            }
    };
    // shifting z accumulators (see fused3d.cu): 2R+1 logical accumulators in 2R+1 + (G-1) slots; inside a
    // G-plane group every update is in place, the last plane of a group writes logical j back to slot j
    auto zscatter = [&](int f, int q, const u64 (&v)[NP], const float* w) {
        constexpr int LASTQ = G - 1;
#pragma unroll
        for (int j = 0; j < 2 * R; ++j) {
            const int s = j + 1 + q, d = q == LASTQ ? j : j + 1 + q;
#pragma unroll
            for (int c = 0; c < NP; ++c) zacc[f][d][c] = fma2s(v[c], w[2 * R - j], zacc[f][s][c]);
        }
#pragma unroll
        for (int c = 0; c < NP; ++c) zacc[f][q == LASTQ ? 2 * R : 2 * R + 1 + q][c] = mul2s(v[c], w[0]);
    };

    int ys0 = 0;
    uint32_t par = 0;
    for (int g = 0; g < n_planes; g += G) {
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int idx = g + q;
            if (idx >= n_planes) break;
            const int ys = ys0 + q;
            const float* yp = ybuf + ys * YSLOT + yoff;
            mbar_wait(&full_y[ys], par);
            u64 v[NZF][NP];
            {
                float win[WIN];
                u64 pe[WIN / 2], po[WIN / 2];
                load_window(yp, win);
                pair_window(win, pe, po);
                if (!GRAD) {
                    xpass(pe, po, p.wx, v[0]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_y[ys]);   // after the x pass: every lane's window is in registers
                } else {
                    xpass(pe, po, p.wx, v[0]);                 // Gy Gx   -> z derivative term
                    xpass(pe, po, p.dx, v[NZF - 1]);           // Gy G'x  -> x term
                }
            }
            if (GRAD) {
                float win[WIN];
                u64 pe[WIN / 2], po[WIN / 2];
                load_window(yp + TYC * YP, win);
                pair_window(win, pe, po);
                xpass(pe, po, p.wx, v[NZF > 1 ? 1 : 0]);       // G'y Gx  -> y term
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_y[ys]);
            }
            if (HAS_Z) {
                constexpr int LASTQ = G - 1;
                const int fs = q == LASTQ ? 0 : q + 1;         // slot of the voxel this plane finishes
                if (!GRAD) {
                    zscatter(0, q, v[0], p.wz);
                } else {
                    zscatter(0, q, v[0], p.dz);
                    zscatter(NZF > 1 ? 1 : 0, q, v[NZF > 1 ? 1 : 0], p.wz);
                    zscatter(NZF - 1, q, v[NZF - 1], p.wz);
                }
                if (idx >= 2 * R) {
                    float res[CPT];
                    if (!GRAD) {
#pragma unroll
                        for (int c = 0; c < NP; ++c) unpack2(zacc[0][fs][c], res[2 * c], res[2 * c + 1]);
                    } else {
                        // filters.py:1187-1201 in the output dtype: d0*d0, += d1*d1, += d2*d2, sqrt — explicit
                        // _rn ops (no FMA contraction): the roundings of the reference's separate ufunc kernels
#pragma unroll
                        for (int c = 0; c < NP; ++c) {
                            float a0, a1, b0, b1, c0, c1;
                            unpack2(zacc[0][fs][c], a0, a1);
                            unpack2(zacc[NZF > 1 ? 1 : 0][fs][c], b0, b1);
                            unpack2(zacc[NZF - 1][fs][c], c0, c1);
                            float s0 = __fmul_rn(a0, a0), s1 = __fmul_rn(a1, a1);
                            s0 = __fadd_rn(s0, __fmul_rn(b0, b0)); s1 = __fadd_rn(s1, __fmul_rn(b1, b1));
                            s0 = __fadd_rn(s0, __fmul_rn(c0, c0)); s1 = __fadd_rn(s1, __fmul_rn(c1, c1));
                            res[2 * c] = __fsqrt_rn(s0); res[2 * c + 1] = __fsqrt_rn(s1);
                        }
                    }
                    if (ok0) *reinterpret_cast<float4*>(out_ptr) = make_float4(res[0], res[1], res[2], res[3]);
                    if (CPT == 8 && ok1)
                        *reinterpret_cast<float4*>(out_ptr + 4) = make_float4(res[CPT - 4], res[CPT - 3], res[CPT - 2], res[CPT - 1]);
                    out_ptr += plane_elems;
                }
            } else {
                float res[CPT];
#pragma unroll
                for (int c = 0; c < NP; ++c) unpack2(v[0][c], res[2 * c], res[2 * c + 1]);
                if (ok0) *reinterpret_cast<float4*>(out_ptr) = make_float4(res[0], res[1], res[2], res[3]);
                if (CPT == 8 && ok1)
                    *reinterpret_cast<float4*>(out_ptr + 4) = make_float4(res[CPT - 4], res[CPT - 3], res[CPT - 2], res[CPT - 1]);
                out_ptr += plane_elems;
            }
        }
        ys0 += G;
        if (ys0 == NY) { ys0 = 0; par ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct WsPlan {
    int tiles_x, tiles_y, yshift, zseg, nzseg;
    double cost;
};

// z segments for a tile height: fill the SMs with whole waves where the shape allows
inline WsPlan plan_tiles(const FusedVolume& v, int R, bool has_z, int tyc, int sms)
{
    WsPlan best{};
    best.cost = 1e300;
    const int tiles_x = (v.nx + 127) / 128;
    const int tiles_y = (v.ny + tyc - 1) / tyc;
    for (int nseg = 1; nseg <= 64 && nseg <= v.nz_out; ++nseg) {
        const int zseg = (v.nz_out + nseg - 1) / nseg;
        const long long ctas = (long long)tiles_x * tiles_y * nseg;
        const long long waves = (ctas + sms - 1) / sms;
        const double per_cta = (double)(zseg + (has_z ? 2 * R : 0) + 6) * tyc;   // + pipeline fill
        const double cost = waves * per_cta;
        if (cost < best.cost) {
            best.cost = cost;
            best.tiles_x = tiles_x; best.tiles_y = tiles_y;
            best.zseg = zseg; best.nzseg = (v.nz_out + zseg - 1) / zseg;
            const int over = tiles_y * tyc - v.ny;
            best.yshift = tiles_y >= 2 ? over / 2 : 0;
        }
        if (!has_z) break;
    }
    return best;
}

inline int device_sms() { return cached_sm_count(); }

template <class C>
cudaError_t launch_cfg(const FusedVolume& v, WsParams& p, const WsPlan& plan, cudaStream_t s)
{
    p.tiles_x = plan.tiles_x; p.tiles_y = plan.tiles_y; p.yshift = plan.yshift;
    p.zseg = plan.zseg; p.nzseg = plan.nzseg;
    CUtensorMap tmap;
    if (!encode_volume_map(&tmap, v.in, v.nx, v.ny, v.nz_in, C::PW, C::BOX_ROWS)) return cudaErrorInvalidValue;
    auto kern = fws_kernel<C>;
    // the attribute is per (function, device): set once per device, not on every call
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !done[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    HaloMaps hm;
    std::memset(&hm, 0, sizeof hm);
    if (v.halo) {
        const sepfilt_halo& h = *v.halo;
        if (h.lo && h.planes_lo) {
            if (!encode_volume_map(&hm.lo, static_cast<const float*>(h.lo), v.nx, v.ny, h.planes_lo, C::PW, C::BOX_ROWS))
                return cudaErrorInvalidValue;
            hm.planes_lo = h.planes_lo;
        }
        if (h.hi && h.planes_hi) {
            if (!encode_volume_map(&hm.hi, static_cast<const float*>(h.hi), v.nx, v.ny, h.planes_hi, C::PW, C::BOX_ROWS))
                return cudaErrorInvalidValue;
            hm.planes_hi = h.planes_hi;
        }
        hm.ready_lo = h.ready_lo; hm.ready_hi = h.ready_hi; hm.epoch = h.epoch;
        hm.done_lo = h.done_lo; hm.done_hi = h.done_hi; hm.counter = h.cta_counter;
    }
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.nzseg;
    kern<<<(unsigned)blocks, C::NT, C::SMEM, s>>>(p, tmap, hm);
    return cudaGetLastError();
}

inline void put_taps(const F32Taps& t, float* w)
{
    for (int k = 0; k < WS_TAPS; ++k) w[k] = 0.f;
    for (int k = 0; k <= 2 * t.radius; ++k) w[k] = t.w[k];
}

// plain filter: radius R (all filtered axes share it), 8 columns per thread, tile rows 16 or 14
template <int R, bool HAS_Z>
cudaError_t launch_plain(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s)
{
    const WsPlan p16 = plan_tiles(v, R, HAS_Z, 16, sms), p14 = plan_tiles(v, R, HAS_Z, 14, sms);
    // cost is in row-planes; a 14-row tile does 14 / 16 of the work of a 16-row tile per plane and has the fifth Y
    // warp (measured ~9 % faster per row-plane): ties go to 14 rows
    if (0.92 * p14.cost < p16.cost) return launch_cfg<WsCfg<R, 8, 14, HAS_Z, false, 104, 200>>(v, p, p14, s);
    return launch_cfg<WsCfg<R, 8, 16, HAS_Z, false, 104, 200>>(v, p, p16, s);
}

// wide plain filters (radius 9 .. 16: sigma 2.5 .. 4 at truncate 4): 2R + 2 z accumulator slots fit the XZ register budget
// only with 4 columns per thread, i.e. 8-row tiles (the geometry of the gradient-magnitude kernel)
template <int R>
cudaError_t launch_wide(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s)
{
    const WsPlan plan = plan_tiles(v, R, true, 8, sms);
    if constexpr (R > 12) return launch_cfg<WsCfg<R, 4, 8, true, false, 88, 208, 1>>(v, p, plan, s);
    else return launch_cfg<WsCfg<R, 4, 8, true, false, 104, 200, 1>>(v, p, plan, s);
}

// gradient magnitude: 4 columns per XZ thread, 8 tile rows (three z accumulator sets per column)
template <int R>
cudaError_t launch_grad(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s)
{
    const WsPlan plan = plan_tiles(v, R, true, 8, sms);
    return launch_cfg<WsCfg<R, 4, 8, true, true, 72, 216, 2>>(v, p, plan, s);
}


// one translation unit per group of instantiations (fused_ws_*.cu: the build compiles them in parallel)
cudaError_t launch_plain_r1_4(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius);
cudaError_t launch_plain_r5_8(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius);
cudaError_t launch_wide_r9_12(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius);
cudaError_t launch_wide_r13_16(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius);
cudaError_t launch_grad_r1_6(const FusedVolume& v, WsParams& p, int sms, cudaStream_t s, int radius);

}  // namespace ws
}  // namespace sepfilt
