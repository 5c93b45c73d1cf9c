// ptx.cuh — sm_100a PTX helpers shared by the fused kernels: packed f32x2 arithmetic, mbarrier,
// TMA tensor loads, register re-allocation between warp roles, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sepfilt {
namespace ptx {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// d = a * (w, w) + c on two packed floats (SASS: FFMA2 with a scalar-broadcast uniform operand)
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
}
__device__ __forceinline__ void mbar_fence_init()
{
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
// generic-proxy accesses to shared memory before, async-proxy (TMA) accesses after
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one (box_x x box_y x 1) box of a 3-D tensor -> shared memory, completion (bytes) on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int cx, int cy, int cz, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// warp-role register budgets: every thread of a 4-warp group executes the same one
template <int N> __device__ __forceinline__ void reg_dealloc()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_alloc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// wait until a flag in global memory (written by a neighbour GPU's stream memory operation) reaches
// `epoch`; one thread.  5 s without progress traps instead of hanging the device.
__device__ __forceinline__ void wait_flag_geq(const uint32_t* flag, uint32_t epoch)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - epoch) >= 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        __nanosleep(200);
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 5000000000ull) __trap();
    }
}
// order the flag read (generic proxy) before the TMA reads (async proxy) of the memory it guards
__device__ __forceinline__ void fence_proxy_async_all()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}

// tensor maps and flags of the neighbour planes beyond a z-slab's ends (sepfilt_halo, include/sepfilt.h)
struct HaloMaps {
    CUtensorMap lo, hi;
    const uint32_t* ready_lo;
    const uint32_t* ready_hi;
    uint32_t* done_lo;
    uint32_t* done_hi;
    unsigned int* counter;
    uint32_t epoch;
    int planes_lo, planes_hi;
    int pad_;
};

// one thread per CTA, after the CTA's last read of neighbour planes: the last CTA of the grid tells the
// neighbours (flags in THEIR memory) that their planes have been read, and re-arms the counter
static __device__ __noinline__ void halo_signal_done(const HaloMaps& hm, unsigned int n_ctas)
{
    if (!hm.counter) return;
    __threadfence();
    if (atomicAdd(hm.counter, 1u) != n_ctas - 1) return;
    *hm.counter = 0;
    __threadfence_system();
    if (hm.done_lo) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hm.done_lo), "r"(hm.epoch) : "memory");
    if (hm.done_hi) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hm.done_hi), "r"(hm.epoch) : "memory");
}

}  // namespace ptx

// SM count of the current device, looked up once per device (not on every launch)
inline int cached_sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = sms;
    }
    return cached[dev];
}

// ---- host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder()
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
// f32 volume (nz, ny, nx) C-contiguous, box (box_x, box_y, 1), out-of-range cells arrive as zeros
inline bool encode_volume_map(CUtensorMap* map, const float* base, int nx, int ny, int nz, int box_x, int box_y)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    const cuuint64_t gstride[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * (cuuint64_t)ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace sepfilt
