/*
 * sepfilt.h — C ABI of libsepfilt_b200.so: separable n-d correlation on B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of mritools/cupyimg:
 * cupyimg.scipy.ndimage.filters — correlate1d / convolve1d and the filters built
 * on them.  The reference has no FFI of its own; its operator boundary is the
 * code-generation seam
 *     kernel = _get_correlate_kernel(mode, w_shape, int_type, offsets, cval)   (filters.py:498-511)
 *     _filters_core._call_kernel(kernel, input, weights, output, weights_dtype) (_filters_core.py:112-156)
 * i.e. "one JIT-compiled ElementwiseKernel launch per 1-D pass".  Every entry
 * point below replaces that seam (or a fixed sequence of such launches) with an
 * ahead-of-time compiled sm_100a kernel.  The Python host layer
 * (cupyimg_b200/scipy/ndimage/filters.py) keeps the reference's signatures and
 * does all argument validation before calling in here.
 *
 * Conventions
 *   - plain C types only; no torch / cupy types cross this boundary;
 *   - the library never allocates or frees device memory: input, output and
 *     scratch are caller-owned; no pointer is retained after return;
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*);
 *     no implicit device synchronisation;
 *   - every function returns SEPFILT_OK (0) or a negative sepfilt_status; the
 *     message for the last failure on the calling thread is sepfilt_last_error();
 *   - re-entrant: no mutable global state except an immutable kernel table and
 *     the thread-local error string (cf. TestThreading, tests/test_filters.py:354-412).
 */
#ifndef SEPFILT_H_
#define SEPFILT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SEPFILT_API __attribute__((visibility("default")))
#else
#define SEPFILT_API
#endif

#define SEPFILT_VERSION 100      /* 0.1.0 */
#define SEPFILT_MAX_NDIM 8
#define SEPFILT_MAX_TAPS 4096    /* taps per 1-D filter accepted by any entry point */
#define SEPFILT_PARAM_TAPS 129   /* taps carried in kernel parameters; longer filters need scratch */
#define SEPFILT_FAST_MAX_RADIUS 16 /* largest tap radius the f32 tiled / fused kernels are built for */

typedef enum {
    SEPFILT_OK = 0,
    SEPFILT_ERR_INVALID = -1,      /* bad argument (NULL, rank, K, mode, dtype) */
    SEPFILT_ERR_UNSUPPORTED = -2,  /* valid request this entry point has no kernel for */
    SEPFILT_ERR_SCRATCH = -3,      /* scratch buffer missing or too small */
    SEPFILT_ERR_CUDA = -4,         /* a CUDA runtime / driver call failed */
    SEPFILT_ERR_VALUE = -5         /* an axis or origin out of range: the argument errors for which the reference raises
                                      ValueError (_util.py / _filters_core.py:63-76) rather than RuntimeError — the class
                                      travels in the status code, not in the message text */
} sepfilt_status;

/* element types: the 10 numeric types the reference tests sweep
 * (tests/test_ndimage.py:62-81) + bool input.  float16 is rejected like scipy. */
typedef enum {
    SEPFILT_I8 = 0, SEPFILT_U8 = 1, SEPFILT_I16 = 2, SEPFILT_U16 = 3,
    SEPFILT_I32 = 4, SEPFILT_U32 = 5, SEPFILT_I64 = 6, SEPFILT_U64 = 7,
    SEPFILT_F32 = 8, SEPFILT_F64 = 9, SEPFILT_BOOL = 10
} sepfilt_dtype;

/* boundary modes (_util.py:170-228; filters treat wrap == grid-wrap, _filters_core.py:224-225;
 * grid-mirror == reflect, grid-constant == constant as in scipy) */
typedef enum {
    SEPFILT_REFLECT = 0,   /* d c b a | a b c d | d c b a */
    SEPFILT_CONSTANT = 1,  /* k k k k | a b c d | k k k k */
    SEPFILT_NEAREST = 2,   /* a a a a | a b c d | d d d d */
    SEPFILT_MIRROR = 3,    /* d c b | a b c d | c b a   */
    SEPFILT_WRAP = 4       /* a b c d | a b c d | a b c d */
} sepfilt_mode;

/* accumulator policy (reference: dtype_mode, _util.py:28-40) */
typedef enum {
    SEPFILT_ACC_F64_EXACT = 0, /* scipy's arithmetic: float64, scipy's summation order, no FMA
                                  contraction; bit-exact integer outputs */
    SEPFILT_ACC_F32 = 1        /* float32 FMA accumulation; only f32 -> f32 */
} sepfilt_acc;

/* a strided device array; stride in BYTES, any sign (cupy allows negative strides) */
typedef struct {
    void*   ptr;
    int32_t dtype;                       /* sepfilt_dtype */
    int32_t ndim;                        /* 0..SEPFILT_MAX_NDIM */
    int64_t shape[SEPFILT_MAX_NDIM];
    int64_t stride_bytes[SEPFILT_MAX_NDIM];
    int32_t device;                      /* CUDA device ordinal that owns ptr */
    int32_t reserved;
} sepfilt_tensor;

/* one 1-D pass of a separable filter */
typedef struct {
    int32_t       axis;      /* normalised, 0 <= axis < ndim */
    int32_t       ntaps;     /* K >= 1 */
    const double* taps;      /* HOST pointer, K correlation weights */
    int32_t       origin;    /* -(K/2) <= origin <= (K-1)/2   (_util.py:98-102) */
    int32_t       mode;      /* sepfilt_mode */
    int32_t       uniform;   /* window filters, taps ignored: 1 = scipy uniform_filter1d semantics (window sum / K);
                              * 2 = minimum_filter1d, 3 = maximum_filter1d (filters.py:1422-1557), exact path only */
    int32_t       reserved;
} sepfilt_pass;

SEPFILT_API int         sepfilt_version(void);
SEPFILT_API const char* sepfilt_last_error(void);

/*
 * One 1-D correlation pass:  out[.., p, ..] = sum_k w[k] * in[.., remap(p + in_offset - (K/2+origin) + k), ..]
 * Replaces one `kernel(input, weights, output)` launch of _filters_core._call_kernel
 * (_filters_core.py:152) for a weights shape that is 1 everywhere except `axis`
 * (_filters_core._convert_1d_args, :51-60).
 *
 *  - in/out: same rank and same extents except along `axis`, where out may be a
 *    window of in: out position p corresponds to in position p + in_offset
 *    (in_offset = 0 and equal extents for the plain filter; the z-slab sharding
 *    uses windows so that halo planes are read but not written).
 *  - boundary remapping applies at the ends of `in` along `axis`.
 *  - in and out must not overlap (the host layer resolves aliasing with a
 *    temporary, like _filters_core.py:148-155).
 *  - acc = SEPFILT_ACC_F64_EXACT: any (in, out) dtype pair.
 *    acc = SEPFILT_ACC_F32: f32 -> f32 only, taps rounded to f32.
 *  - scratch: only read when pass->ntaps > SEPFILT_PARAM_TAPS (needs ntaps*8 bytes, device).
 */
SEPFILT_API int sepfilt_correlate1d(const sepfilt_tensor* in, const sepfilt_tensor* out,
                        const sepfilt_pass* pass, int64_t in_offset, double cval,
                        int acc, void* scratch, size_t scratch_bytes, void* stream);

/*
 * A whole separable filter (up to one pass per axis) on a C-contiguous f32 volume
 * in ONE launch: replaces the per-axis loop of uniform_filter (filters.py:651-662),
 * gaussian_filter (:777-789) and the N _call_kernel launches + copy-backs under it.
 * The volume crosses HBM once in and once out.
 *
 *  - in, out: f32, rank 2 or 3, C-contiguous, non-overlapping, out may be a window
 *    of in along axis 0 (in_offset0) as above.
 *  - passes[i].axis must be distinct; every pass needs
 *    max(K/2+origin, K-1-K/2-origin) <= SEPFILT_FAST_MAX_RADIUS and a radius no larger
 *    than the extent of its axis; otherwise SEPFILT_ERR_UNSUPPORTED (host falls back to
 *    sepfilt_correlate1d per axis).
 *  - gradient_magnitude = 0: out = pass_{n-1}(..pass_0(in)).
 *    gradient_magnitude = 1: `passes` holds ndim smoothing passes and `dpasses` ndim
 *    derivative passes; out = sqrt(sum_a (D_a prod_{b!=a} S_b in)^2)
 *    (generic_gradient_magnitude, filters.py:1175-1201).
 */
SEPFILT_API int sepfilt_separable_f32(const sepfilt_tensor* in, const sepfilt_tensor* out,
                          const sepfilt_pass* passes, int npasses,
                          const sepfilt_pass* dpasses, int gradient_magnitude,
                          int64_t in_offset0, double cval, void* stream);

/*
 * z-slab halos read straight from the neighbours' memory (multi-GPU, one process per GPU).
 * The reference is single-GPU; this is the scale-out of its per-axis loops (filters.py:651-662,
 * :777-789): the volume is sharded along axis 0 and a rank's z pass needs `radius` planes of RAW
 * input from each neighbour.  Instead of exchanging them into local buffers first, the fused kernel
 * loads those planes with TMA directly from the neighbour's array (a peer-mapped pointer: CUDA IPC /
 * VMM symmetric memory over NVLink), so no copy and no communication kernel sits on the data path.
 *   lo: the `planes_lo` planes that precede in-plane 0 (the lower neighbour's LAST planes), C-contiguous
 *       (planes_lo, ny, nx) float32; NULL / 0 planes: the boundary mode applies at the slab's start.
 *   hi: the `planes_hi` planes that follow the last in-plane (the upper neighbour's FIRST planes).
 *   ready_lo / ready_hi: 32-bit flags in THIS device's memory; the kernel reads a neighbour's planes only
 *       after the flag is >= epoch (the neighbour writes it with a stream memory operation once its
 *       array is complete).  NULL: no wait.  A flag that stays below epoch for 5 s traps the kernel.
 *   done_lo / done_hi: 32-bit flags in the NEIGHBOURS' memory (peer-mapped) that the kernel sets to epoch
 *       when its last CTA has finished — "I have read your planes, you may overwrite them" — so that no
 *       stream operation is needed after the launch.  NULL: not signalled.  Needs cta_counter: one 32-bit
 *       word of THIS device's memory, zero before the first launch (the kernel leaves it zero).
 * planes_lo / planes_hi must be 0 or >= the z radius of the filter.
 */
typedef struct {
    const void*     lo;
    const void*     hi;
    int32_t         planes_lo, planes_hi;
    const uint32_t* ready_lo;
    const uint32_t* ready_hi;
    uint32_t*       done_lo;
    uint32_t*       done_hi;
    uint32_t*       cta_counter;
    uint32_t        epoch;
    uint32_t        reserved;
} sepfilt_halo;

/* sepfilt_separable_f32 on a z-slab with neighbour halos (rank-3 float32, a z pass present).  `out` may be a
 * window of `in` along axis 0 as in sepfilt_separable_f32 (out plane z <-> in plane z + in_offset0,
 * 0 <= in_offset0, in_offset0 + out planes <= in planes): the r-plane boundary strips of a slab are filtered
 * from its first / last 2r planes plus one neighbour's halo while the interior runs without any halo.
 * SEPFILT_ERR_UNSUPPORTED when no fused kernel takes the request: the caller then exchanges halos into a
 * local buffer and uses the windowed call. */
SEPFILT_API int sepfilt_separable_f32_halo(const sepfilt_tensor* in, const sepfilt_tensor* out,
                               const sepfilt_pass* passes, int npasses,
                               const sepfilt_pass* dpasses, int gradient_magnitude,
                               const sepfilt_halo* halo, int64_t in_offset0, double cval, void* stream);

/* Stream-ordered 32-bit flag operations (cuStreamWriteValue32 / cuStreamWaitValue32, no kernel, no SM):
 * write `value` to *addr when the stream reaches this point (addr may be peer-mapped), or hold the
 * stream until *addr >= value.  These carry the ready / done flags of the halo protocol above. */
SEPFILT_API int sepfilt_stream_write32(void* stream, void* addr, uint32_t value);
/* the same value to two addresses in ONE batched stream operation (either may be NULL) */
SEPFILT_API int sepfilt_stream_write32x2(void* stream, void* addr_a, void* addr_b, uint32_t value);
SEPFILT_API int sepfilt_stream_wait32_geq(void* stream, void* addr, uint32_t value);

/*
 * Elementwise stages of the skimage-level callers of the separable filters (float32 / float64, contiguous,
 * n elements; every operation separately rounded in the array dtype like the reference's cupy ufunc calls):
 *  sepfilt_multiply: out = a * b — the products SSIM filters (skimage/metrics/_structural_similarity.py:203-207:
 *      im1 * im1, im2 * im2, im1 * im2) and the structure tensor (skimage/feature/corner.py:131-134: der0 * der1);
 *      out may alias a or b.
 *  sepfilt_ssim_map: the SSIM map S = (A1 * A2) / (B1 * B2) of _structural_similarity.py:208-224 from the five
 *      filtered arrays, written to S (may be NULL) and summed in float64 over crop(S, pad) (:226-230) into *sum
 *      (device memory, zeroed by the caller) — one pass instead of ~20 elementwise kernels.  ndim <= 3.
 */
SEPFILT_API int sepfilt_multiply(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
SEPFILT_API int sepfilt_ssim_map(const void* ux, const void* uy, const void* uxx, const void* uyy, const void* uxy,
                     void* S, double* sum, int ndim, const int64_t* shape, int pad,
                     double cov_norm, double C1, double C2, int dtype, void* stream);

/* Kernels enqueued by the calling thread's last successful sepfilt_separable_f32 call (1, or one per
 * axis when the gradient magnitude runs as accumulating launches): lets the host layer report
 * launch counts instead of assuming them. */
SEPFILT_API int sepfilt_last_launch_count(void);

/* Would sepfilt_separable_f32_halo (and therefore sepfilt_separable_f32) accept this request?  1 yes, 0 no (no error
 * is set).  Answers for a launch with neighbour halos: the z-slab sharding asks before it relies on the halo launch. */
SEPFILT_API int sepfilt_separable_f32_supported(const sepfilt_tensor* in, const sepfilt_tensor* out,
                                    const sepfilt_pass* passes, int npasses,
                                    int gradient_magnitude, double cval);

/* Elementwise epilogue helpers used by generic_gradient_magnitude on the exact path
 * (filters.py:1187-1201: multiply / += / sqrt, all in the OUTPUT dtype):
 *   op 0: acc  = a*a          (first axis)
 *   op 1: acc += a*a          (further axes)
 *   op 2: acc  = sqrt(acc)    (final, "unsafe" cast back to the dtype)
 *   op 3: acc += a            (generic_laplace accumulation, filters.py:1024-1035)
 *   op 4: acc -= a            (real part of a complex x complex correlation: re*re - im*im; the
 *                              reference does the complex multiply-add in the kernel, filters.py:467-469)
 * a and acc are C-contiguous arrays of `dtype` with n elements. */
SEPFILT_API int sepfilt_gradmag_step(void* acc, const void* a, int64_t n, int dtype, int op, void* stream);

/*
 * Dense N-d correlation with a small N-d weights array: replaces one
 * `kernel(input, weights, output)` launch of _filters_core._call_kernel (_filters_core.py:152) for the
 * kernel _get_correlate_kernel builds from N-d weights (filters.py:498-511; callers correlate / convolve,
 * filters.py:65-210).   out[i] = sum_k w[k] * in[remap(i + k - (wshape/2 + origin))]  per axis, in scipy's
 * arithmetic (float64, taps with |w| > DBL_EPSILON in C order, no FMA contraction, C-cast store).
 *  - in / out: same shape, any dtype pair, any byte strides, must not overlap;
 *  - weights: HOST doubles, C order, rank = in->ndim, extents wshape[]; origin[] per axis;
 *  - scratch: only read when the number of taps exceeds SEPFILT_PARAM_TAPS (needs taps * 8 bytes, device).
 */
SEPFILT_API int sepfilt_correlate_nd(const sepfilt_tensor* in, const sepfilt_tensor* out,
                         const double* weights, const int32_t* wshape, const int32_t* origin,
                         int mode, double cval, void* scratch, size_t scratch_bytes, void* stream);

/* Strided copy with dtype conversion under the same cast rules as the filter store
 * (used for the "no axes to filter -> output[...] = input[...]" branches, filters.py:663-664). */
SEPFILT_API int sepfilt_copy_cast(const sepfilt_tensor* in, const sepfilt_tensor* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEPFILT_H_ */
