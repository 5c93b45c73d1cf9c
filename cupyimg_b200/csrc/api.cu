// api.cu — the extern "C" boundary declared in include/sepfilt.h: argument validation,
// geometry normalisation and kernel dispatch.  No device allocation, no retained
// pointers, all launches on the caller's stream.
#include <cuda.h>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "kernels.h"

using namespace sepfilt;

namespace {

thread_local std::string g_last_error;
thread_local int g_last_launches = 0;     // kernels enqueued by the last sepfilt_separable_f32 call of this thread

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int fail_cuda(cudaError_t e, const char* what)
{
    return fail(SEPFILT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

bool valid_dtype(int t) { return t >= SEPFILT_I8 && t <= SEPFILT_BOOL; }

int check_tensor(const sepfilt_tensor* t, const char* name, bool is_output)
{
    if (!t) return fail(SEPFILT_ERR_INVALID, "%s tensor is NULL", name);
    if (t->ndim < 0 || t->ndim > SEPFILT_MAX_NDIM)
        return fail(SEPFILT_ERR_INVALID, "%s rank %d out of range", name, t->ndim);
    if (!valid_dtype(t->dtype) || (is_output && t->dtype == SEPFILT_BOOL))
        return fail(SEPFILT_ERR_INVALID, "%s dtype %d not supported", name, t->dtype);
    for (int d = 0; d < t->ndim; ++d)
        if (t->shape[d] < 0) return fail(SEPFILT_ERR_INVALID, "%s has a negative extent", name);
    return SEPFILT_OK;
}

int64_t numel(const sepfilt_tensor* t)
{
    int64_t n = 1;
    for (int d = 0; d < t->ndim; ++d) n *= t->shape[d];
    return n;
}

bool c_contiguous(const sepfilt_tensor* t)
{
    int64_t expect = dtype_size(t->dtype);
    for (int d = t->ndim - 1; d >= 0; --d) {
        if (t->shape[d] != 1 && t->stride_bytes[d] != expect) return false;
        expect *= t->shape[d];
    }
    return true;
}

int check_pass(const sepfilt_pass* p, int ndim)
{
    if (!p) return fail(SEPFILT_ERR_INVALID, "pass is NULL");
    if (p->axis < 0 || p->axis >= ndim) return fail(SEPFILT_ERR_VALUE, "invalid axis %d", p->axis);
    if (p->ntaps < 1 || p->ntaps > SEPFILT_MAX_TAPS)
        return fail(SEPFILT_ERR_INVALID, "filter length %d not in 1..%d", p->ntaps, SEPFILT_MAX_TAPS);
    const int before = p->ntaps / 2 + p->origin;
    if (before < 0 || before >= p->ntaps) return fail(SEPFILT_ERR_VALUE, "invalid origin");
    if (p->mode < SEPFILT_REFLECT || p->mode > SEPFILT_WRAP)
        return fail(SEPFILT_ERR_INVALID, "boundary mode not supported");
    if (!p->uniform && !p->taps) return fail(SEPFILT_ERR_INVALID, "no filter weights given");
    return SEPFILT_OK;
}

// taps of one pass as offsets -R..R around the output position, rounded to f32
bool pass_to_f32_taps(const sepfilt_pass* p, F32Taps* t)
{
    if (p->uniform > 1) return false;            // minimum / maximum windows are not correlations
    const int before = p->ntaps / 2 + p->origin;
    const int after = p->ntaps - 1 - before;
    const int R = before > after ? before : after;
    if (R > SEPFILT_FAST_MAX_RADIUS) return false;
    t->radius = R;
    for (int k = 0; k <= 2 * SEPFILT_FAST_MAX_RADIUS; ++k) t->w[k] = 0.f;
    for (int k = 0; k < p->ntaps; ++k)
        t->w[k - before + R] = p->uniform ? (float)(1.0 / p->ntaps) : (float)p->taps[k];
    return true;
}

// scipy's NI_Correlate1D symmetry probe
int probe_symmetry(const double* w, int K)
{
    if (!(K & 1)) return 0;
    const int s1 = K / 2;
    bool sym = true;
    for (int i = 1; i <= s1; ++i)
        if (std::fabs(w[s1 + i] - w[s1 - i]) > DBL_EPSILON) { sym = false; break; }
    if (sym) return 1;
    for (int i = 1; i <= s1; ++i)
        if (std::fabs(w[s1 + i] + w[s1 - i]) > DBL_EPSILON) return 0;
    return -1;
}

}  // namespace

extern "C" {

int sepfilt_version(void) { return SEPFILT_VERSION; }

const char* sepfilt_last_error(void) { return g_last_error.c_str(); }

int sepfilt_correlate1d(const sepfilt_tensor* in, const sepfilt_tensor* out,
                        const sepfilt_pass* pass, int64_t in_offset, double cval,
                        int acc, void* scratch, size_t scratch_bytes, void* stream)
{
    int rc;
    if ((rc = check_tensor(in, "input", false)) || (rc = check_tensor(out, "output", true))) return rc;
    if (in->ndim != out->ndim) return fail(SEPFILT_ERR_INVALID, "input and output rank differ");
    if (in->ndim < 1) return fail(SEPFILT_ERR_INVALID, "rank must be >= 1");
    if ((rc = check_pass(pass, in->ndim))) return rc;
    for (int d = 0; d < in->ndim; ++d)
        if (d != pass->axis && in->shape[d] != out->shape[d])
            return fail(SEPFILT_ERR_INVALID, "output shape not correct");
    if (in->device != out->device) return fail(SEPFILT_ERR_INVALID, "input and output on different devices");
    if (acc != SEPFILT_ACC_F64_EXACT && acc != SEPFILT_ACC_F32)
        return fail(SEPFILT_ERR_INVALID, "unknown accumulator policy %d", acc);
    const int64_t total = numel(out);
    if (total == 0) return SEPFILT_OK;
    if (in->shape[pass->axis] < 1) return fail(SEPFILT_ERR_VALUE, "input is empty along the filtered axis");
    if (!in->ptr || !out->ptr) return fail(SEPFILT_ERR_INVALID, "NULL data pointer");

    DeviceGuard guard(in->device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int axis = pass->axis;

    if (acc == SEPFILT_ACC_F32) {
        if (in->dtype != SEPFILT_F32 || out->dtype != SEPFILT_F32)
            return fail(SEPFILT_ERR_UNSUPPORTED, "float32 accumulation needs float32 input and output");
        if (!c_contiguous(in) || !c_contiguous(out))
            return fail(SEPFILT_ERR_UNSUPPORTED, "float32 tiled pass needs C-contiguous arrays");
        F32Taps taps;
        if (!pass_to_f32_taps(pass, &taps))
            return fail(SEPFILT_ERR_UNSUPPORTED, "filter radius exceeds %d", SEPFILT_FAST_MAX_RADIUS);
        // the 32-bit boundary fold computes 2n: axes beyond 2^30 elements take the 64-bit exact kernel
        if (in->shape[axis] > 1073741824LL || out->shape[axis] > 1073741824LL ||
            in_offset > 2147483647LL / 2 || in_offset < -2147483647LL / 2)
            return fail(SEPFILT_ERR_UNSUPPORTED, "axis too long for the tiled pass");
        F32Line g;
        g.in = static_cast<const float*>(in->ptr);
        g.out = static_cast<float*>(out->ptr);
        g.outer = 1;
        g.inner = 1;
        for (int d = 0; d < axis; ++d) g.outer *= in->shape[d];
        for (int d = axis + 1; d < in->ndim; ++d) g.inner *= in->shape[d];
        g.n_in = (int32_t)in->shape[axis];
        g.n_out = (int32_t)out->shape[axis];
        g.in_offset = (int32_t)in_offset;
        g.mode = pass->mode;
        g.cval = (float)cval;
        if (!f32_line_supported(g, taps.radius))
            return fail(SEPFILT_ERR_UNSUPPORTED, "geometry not supported by the tiled pass");
        cudaError_t e = launch_f32_corr1d(g, taps, s);
        if (e != cudaSuccess) return fail_cuda(e, "corr1d_f32 launch");
        return SEPFILT_OK;
    }

    // ---- exact path, tiled kernels where the geometry and the taps allow ----
    if ((pass->uniform == 0 || pass->uniform > 1) && c_contiguous(in) && c_contiguous(out)) {
        ExactTiledGeom g;
        g.in = in->ptr;
        g.out = out->ptr;
        g.in_dtype = in->dtype;
        g.out_dtype = out->dtype;
        g.outer = 1;
        g.inner = 1;
        for (int d = 0; d < axis; ++d) g.outer *= in->shape[d];
        for (int d = axis + 1; d < in->ndim; ++d) g.inner *= in->shape[d];
        g.n_in = in->shape[axis];
        g.n_out = out->shape[axis];
        g.shift = in_offset - pass->origin;
        if (pass->uniform > 1) {
            // minimum / maximum window: streaming kernels for the dtype-preserving common cases
            g.shift = in_offset;
            static const bool no_mm = getenv("SEPFILT_NO_MINMAX_STREAM") != nullptr;   // A/B aid
            if (!no_mm && pass->uniform <= 3 && minmax_stream_supported(g, pass->ntaps, pass->origin, cval)) {
                cudaError_t e = launch_minmax_stream(g, pass->ntaps, pass->origin, pass->mode, cval,
                                                     pass->uniform == 3, s);
                if (e != cudaSuccess) return fail_cuda(e, "minmax_stream launch");
                return SEPFILT_OK;
            }
        }
    }
    if (!pass->uniform && c_contiguous(in) && c_contiguous(out)) {
        const int sym = probe_symmetry(pass->taps, pass->ntaps);
        ExactTiledGeom g;
        g.in = in->ptr;
        g.out = out->ptr;
        g.in_dtype = in->dtype;
        g.out_dtype = out->dtype;
        g.outer = 1;
        g.inner = 1;
        for (int d = 0; d < axis; ++d) g.outer *= in->shape[d];
        for (int d = axis + 1; d < in->ndim; ++d) g.inner *= in->shape[d];
        g.n_in = in->shape[axis];
        g.n_out = out->shape[axis];
        g.shift = in_offset - pass->origin;
        static const bool no_stream = getenv("SEPFILT_NO_STREAM") != nullptr;   // test / tuning aid
        if (!no_stream && exact_stream_supported(g, pass->ntaps, sym)) {
            cudaError_t e = launch_exact_stream(g, pass->taps, pass->ntaps, sym, pass->mode, cval, s);
            if (e != cudaSuccess) return fail_cuda(e, "exact_stream launch");
            return SEPFILT_OK;
        }
        if (exact_tiled_supported(g, pass->ntaps, sym)) {
            cudaError_t e = launch_exact_tiled(g, pass->taps, pass->ntaps, sym, pass->mode, cval, s);
            if (e != cudaSuccess) return fail_cuda(e, "exact_tiled launch");
            return SEPFILT_OK;
        }
    }

    // ---- exact path, per-element kernel (any strides, any taps) ----
    ExactParams p;
    std::memset(&p, 0, sizeof p);
    p.in = static_cast<const char*>(in->ptr);
    p.out = static_cast<char*>(out->ptr);
    p.in_dtype = in->dtype;
    p.out_dtype = out->dtype;
    p.n_in = in->shape[axis];
    p.in_offset = in_offset;
    p.total = total;
    p.K = pass->ntaps;
    p.before = pass->ntaps / 2 + pass->origin;
    p.mode = pass->mode;
    p.cval = cval;
    // collapse: drop unit dims, merge neighbours that are jointly contiguous in both arrays
    int nd = 0;
    p.axis = -1;
    for (int d = 0; d < in->ndim; ++d) {
        const bool is_axis = (d == axis);
        if (!is_axis && out->shape[d] == 1) continue;
        if (nd > 0 && !is_axis && p.axis != nd - 1 &&
            p.istride[nd - 1] == in->stride_bytes[d] * out->shape[d] &&
            p.ostride[nd - 1] == out->stride_bytes[d] * out->shape[d]) {
            p.shape[nd - 1] *= out->shape[d];
            p.istride[nd - 1] = in->stride_bytes[d];
            p.ostride[nd - 1] = out->stride_bytes[d];
            continue;
        }
        p.shape[nd] = out->shape[d];
        p.istride[nd] = in->stride_bytes[d];
        p.ostride[nd] = out->stride_bytes[d];
        if (is_axis) p.axis = nd;
        ++nd;
    }
    p.ndim = nd;
    if (pass->uniform) {
        if (pass->uniform < 1 || pass->uniform > 3) return fail(SEPFILT_ERR_INVALID, "unknown window kind %d", pass->uniform);
        p.symmetric = 1 + pass->uniform;         // 2: mean, 3: minimum, 4: maximum
    } else {
        p.symmetric = probe_symmetry(pass->taps, pass->ntaps);
        if (pass->ntaps <= SEPFILT_PARAM_TAPS) {
            std::memcpy(p.w, pass->taps, sizeof(double) * pass->ntaps);
        } else {
            const size_t need = sizeof(double) * (size_t)pass->ntaps;
            if (!scratch || scratch_bytes < need)
                return fail(SEPFILT_ERR_SCRATCH, "filters longer than %d taps need %zu bytes of device scratch",
                            SEPFILT_PARAM_TAPS, need);
            cudaError_t e = cudaMemcpyAsync(scratch, pass->taps, need, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(taps)");
            p.wdev = static_cast<const double*>(scratch);
        }
    }
    cudaError_t e = launch_exact_corr1d(p, s);
    if (e != cudaSuccess) return fail_cuda(e, "exact_corr1d launch");
    return SEPFILT_OK;
}

static int build_fused(const sepfilt_tensor* in, const sepfilt_tensor* out,
                       const sepfilt_pass* passes, int npasses, const sepfilt_pass* dpasses,
                       int gradient_magnitude, int64_t in_offset0, double cval,
                       FusedVolume* v, F32Taps taps[3], F32Taps dtaps[3], bool set_error,
                       const sepfilt_halo* halo = nullptr)
{
#define UNSUP(...) return set_error ? fail(SEPFILT_ERR_UNSUPPORTED, __VA_ARGS__) : SEPFILT_ERR_UNSUPPORTED
    if (!in || !out || !passes) UNSUP("NULL argument");
    if (in->dtype != SEPFILT_F32 || out->dtype != SEPFILT_F32) UNSUP("fused path is float32 only");
    if (in->ndim != out->ndim || (in->ndim != 2 && in->ndim != 3)) UNSUP("fused path needs rank 2 or 3");
    if (!c_contiguous(in) || !c_contiguous(out)) UNSUP("fused path needs C-contiguous arrays");
    const int nd = in->ndim;
    for (int d = 1; d < nd; ++d)
        if (in->shape[d] != out->shape[d]) UNSUP("shape mismatch");
    if (nd == 2 && (in_offset0 != 0 || in->shape[0] != out->shape[0])) UNSUP("windows need rank 3");
    if (npasses < 1 || npasses > nd) UNSUP("bad pass count");
    if (gradient_magnitude && (npasses != nd || !dpasses)) UNSUP("gradient magnitude needs one pass per axis");
    for (int a = 0; a < 3; ++a) {
        taps[a].radius = 0;
        for (int k = 0; k <= 2 * SEPFILT_FAST_MAX_RADIUS; ++k) taps[a].w[k] = dtaps[a].w[k] = 0.f;
        taps[a].w[0] = 1.f;   // identity
        dtaps[a] = taps[a];
        v->mode[a] = SEPFILT_NEAREST;
    }
    bool seen[3] = {false, false, false};
    for (int i = 0; i < npasses; ++i) {
        const sepfilt_pass* p = &passes[i];
        if (p->axis < 0 || p->axis >= nd || p->ntaps < 1) UNSUP("bad pass");
        const int before = p->ntaps / 2 + p->origin;
        if (before < 0 || before >= p->ntaps) UNSUP("invalid origin");
        const int a = p->axis + (3 - nd);      // slot: 0 = z, 1 = y, 2 = x
        if (seen[a]) UNSUP("duplicate axis");
        seen[a] = true;
        if (!pass_to_f32_taps(p, &taps[a])) UNSUP("radius too large");
        if (taps[a].radius > in->shape[p->axis]) UNSUP("radius exceeds the axis extent");
        v->mode[a] = p->mode;
        if (gradient_magnitude) {
            const sepfilt_pass* q = &dpasses[i];
            if (q->axis != p->axis || q->mode != p->mode) UNSUP("derivative passes must mirror the smoothing passes");
            if (!pass_to_f32_taps(q, &dtaps[a])) UNSUP("radius too large");
            if (dtaps[a].radius > in->shape[p->axis]) UNSUP("radius exceeds the axis extent");
        }
    }
    v->in = static_cast<const float*>(in->ptr);
    v->out = static_cast<float*>(out->ptr);
    if (nd == 3) {
        if (in->shape[0] > 1073741824LL || out->shape[0] > 1073741824LL) UNSUP("too many planes");
        v->nz_in = (int32_t)in->shape[0];
        v->nz_out = (int32_t)out->shape[0];
        v->z_offset = (int32_t)in_offset0;
    } else {
        v->nz_in = v->nz_out = 1;
        v->z_offset = 0;
    }
    if (in->shape[nd - 2] > 1073741824LL || in->shape[nd - 1] > 1073741824LL) UNSUP("plane too large");
    v->ny = (int32_t)in->shape[nd - 2];
    v->nx = (int32_t)in->shape[nd - 1];
    v->cval = (float)cval;
    v->halo = halo;
    if (halo) {
        if (nd != 3) UNSUP("neighbour halos need a rank-3 slab");
        if (halo->planes_lo < 0 || halo->planes_hi < 0) UNSUP("negative halo plane count");
        if ((halo->lo && (reinterpret_cast<uintptr_t>(halo->lo) & 15)) || (halo->hi && (reinterpret_cast<uintptr_t>(halo->hi) & 15)))
            UNSUP("halo planes must be 16-byte aligned");
    }
    if (!fused_ws_supported(*v, taps, dtaps, gradient_magnitude != 0) &&
        !fused3d_supported(*v, taps, gradient_magnitude != 0))
        UNSUP("geometry not supported by the fused kernel");
    return SEPFILT_OK;
#undef UNSUP
}

int sepfilt_separable_f32_supported(const sepfilt_tensor* in, const sepfilt_tensor* out,
                                    const sepfilt_pass* passes, int npasses, int gradient_magnitude, double cval)
{
    FusedVolume v;
    F32Taps taps[3], dtaps[3];
    // derivative passes are only needed for their radius == smoothing radius here.  The query is what the z-slab
    // sharding asks before it relies on sepfilt_separable_f32_halo, so it answers for a launch WITH neighbour halos
    sepfilt_halo with_halo;
    std::memset(&with_halo, 0, sizeof with_halo);
    return build_fused(in, out, passes, npasses, passes, gradient_magnitude, 0, cval, &v, taps, dtaps, false, &with_halo) == SEPFILT_OK;
}

int sepfilt_separable_f32(const sepfilt_tensor* in, const sepfilt_tensor* out,
                          const sepfilt_pass* passes, int npasses,
                          const sepfilt_pass* dpasses, int gradient_magnitude,
                          int64_t in_offset0, double cval, void* stream)
{
    FusedVolume v;
    F32Taps taps[3], dtaps[3];
    int rc = build_fused(in, out, passes, npasses, dpasses, gradient_magnitude, in_offset0, cval,
                         &v, taps, dtaps, true);
    if (rc != SEPFILT_OK) return rc;
    if (numel(out) == 0) return SEPFILT_OK;
    DeviceGuard guard(in->device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool ws = fused_ws_supported(v, taps, dtaps, gradient_magnitude != 0);
    cudaError_t e = ws ? launch_fused_ws(v, taps, dtaps, gradient_magnitude != 0, s)
                       : launch_fused3d(v, taps, dtaps, gradient_magnitude != 0, s);
    if (e != cudaSuccess) return fail_cuda(e, "fused launch");
    // the warp-specialised kernel computes the gradient magnitude in one launch, fused3d in one per axis
    g_last_launches = (gradient_magnitude && !ws) ? in->ndim : 1;
    return SEPFILT_OK;
}

int sepfilt_last_launch_count(void) { return g_last_launches; }

int sepfilt_multiply(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream)
{
    if (!a || !b || !out || n < 0) return fail(SEPFILT_ERR_INVALID, "bad multiply arguments");
    if (dtype != SEPFILT_F32 && dtype != SEPFILT_F64) return fail(SEPFILT_ERR_UNSUPPORTED, "multiply takes float32 / float64");
    cudaError_t e = launch_multiply(a, b, out, n, dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "multiply launch");
    return SEPFILT_OK;
}

int sepfilt_ssim_map(const void* ux, const void* uy, const void* uxx, const void* uyy, const void* uxy,
                     void* S, double* sum, int ndim, const int64_t* shape, int pad,
                     double cov_norm, double C1, double C2, int dtype, void* stream)
{
    if (!ux || !uy || !uxx || !uyy || !uxy || !sum || !shape) return fail(SEPFILT_ERR_INVALID, "bad ssim_map arguments");
    if (ndim < 1 || ndim > 3 || pad < 0) return fail(SEPFILT_ERR_INVALID, "ssim_map takes 1 to 3 dimensions");
    if (dtype != SEPFILT_F32 && dtype != SEPFILT_F64) return fail(SEPFILT_ERR_UNSUPPORTED, "ssim_map takes float32 / float64");
    cudaError_t e = launch_ssim_map(ux, uy, uxx, uyy, uxy, S, sum, ndim, shape, pad, cov_norm, C1, C2, dtype,
                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "ssim_map launch");
    return SEPFILT_OK;
}

int sepfilt_separable_f32_halo(const sepfilt_tensor* in, const sepfilt_tensor* out,
                               const sepfilt_pass* passes, int npasses,
                               const sepfilt_pass* dpasses, int gradient_magnitude,
                               const sepfilt_halo* halo, int64_t in_offset0, double cval, void* stream)
{
    if (!halo) return fail(SEPFILT_ERR_INVALID, "halo descriptor is NULL");
    FusedVolume v;
    F32Taps taps[3], dtaps[3];
    int rc = build_fused(in, out, passes, npasses, dpasses, gradient_magnitude, in_offset0, cval, &v, taps, dtaps, true, halo);
    if (rc != SEPFILT_OK) return rc;
    if (numel(out) == 0) return SEPFILT_OK;
    DeviceGuard guard(in->device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool ws = fused_ws_supported(v, taps, dtaps, gradient_magnitude != 0);
    cudaError_t e = ws ? launch_fused_ws(v, taps, dtaps, gradient_magnitude != 0, s)
                       : launch_fused3d(v, taps, dtaps, gradient_magnitude != 0, s);
    if (e != cudaSuccess) return fail_cuda(e, "fused launch (halo)");
    g_last_launches = 1;
    return SEPFILT_OK;
}

namespace {
typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValue32Fn stream_value_fn(const char* name)
{
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<StreamValue32Fn>(f);
}
}  // namespace

int sepfilt_stream_write32(void* stream, void* addr, uint32_t value)
{
    static StreamValue32Fn fn = stream_value_fn("cuStreamWriteValue32");
    if (!fn) return fail(SEPFILT_ERR_CUDA, "cuStreamWriteValue32 is not available");
    if (!addr || (reinterpret_cast<uintptr_t>(addr) & 3)) return fail(SEPFILT_ERR_INVALID, "flag address must be 4-byte aligned");
    const CUresult r = fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) return fail(SEPFILT_ERR_CUDA, "cuStreamWriteValue32 failed (%d)", (int)r);
    return SEPFILT_OK;
}

int sepfilt_stream_write32x2(void* stream, void* addr_a, void* addr_b, uint32_t value)
{
    typedef CUresult (*BatchFn)(CUstream, unsigned int, CUstreamBatchMemOpParams*, unsigned int);
    static BatchFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<BatchFn>(f);
    }();
    if (!fn) return fail(SEPFILT_ERR_CUDA, "cuStreamBatchMemOp is not available");
    CUstreamBatchMemOpParams ops[2];
    memset(ops, 0, sizeof ops);
    unsigned int n = 0;
    void* addrs[2] = {addr_a, addr_b};
    for (void* a : addrs) {
        if (!a) continue;
        if (reinterpret_cast<uintptr_t>(a) & 3) return fail(SEPFILT_ERR_INVALID, "flag address must be 4-byte aligned");
        ops[n].writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
        ops[n].writeValue.address = reinterpret_cast<CUdeviceptr>(a);
        ops[n].writeValue.value = value;
        ops[n].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
        ++n;
    }
    if (!n) return SEPFILT_OK;
    const CUresult r = fn(static_cast<CUstream>(stream), n, ops, 0);
    if (r != CUDA_SUCCESS) return fail(SEPFILT_ERR_CUDA, "cuStreamBatchMemOp failed (%d)", (int)r);
    return SEPFILT_OK;
}

int sepfilt_stream_wait32_geq(void* stream, void* addr, uint32_t value)
{
    static StreamValue32Fn fn = stream_value_fn("cuStreamWaitValue32");
    if (!fn) return fail(SEPFILT_ERR_CUDA, "cuStreamWaitValue32 is not available");
    if (!addr || (reinterpret_cast<uintptr_t>(addr) & 3)) return fail(SEPFILT_ERR_INVALID, "flag address must be 4-byte aligned");
    const CUresult r = fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return fail(SEPFILT_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)r);
    return SEPFILT_OK;
}

int sepfilt_gradmag_step(void* acc, const void* a, int64_t n, int dtype, int op, void* stream)
{
    if (n < 0 || op < 0 || op > 4 || dtype < SEPFILT_I8 || dtype > SEPFILT_F64)
        return fail(SEPFILT_ERR_INVALID, "bad gradmag_step arguments");
    if (n == 0) return SEPFILT_OK;
    if (!acc || !a) return fail(SEPFILT_ERR_INVALID, "NULL data pointer");
    cudaError_t e = launch_gradmag_step(acc, a, n, dtype, op, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "gradmag_step launch");
    return SEPFILT_OK;
}

int sepfilt_correlate_nd(const sepfilt_tensor* in, const sepfilt_tensor* out,
                         const double* weights, const int32_t* wshape, const int32_t* origin,
                         int mode, double cval, void* scratch, size_t scratch_bytes, void* stream)
{
    int rc;
    if ((rc = check_tensor(in, "input", false)) || (rc = check_tensor(out, "output", true))) return rc;
    if (in->ndim != out->ndim) return fail(SEPFILT_ERR_INVALID, "input and output rank differ");
    if (in->ndim < 1) return fail(SEPFILT_ERR_INVALID, "rank must be >= 1");
    if (!weights || !wshape || !origin) return fail(SEPFILT_ERR_INVALID, "no filter weights given");
    if (mode < SEPFILT_REFLECT || mode > SEPFILT_WRAP) return fail(SEPFILT_ERR_INVALID, "boundary mode not supported");
    if (in->device != out->device) return fail(SEPFILT_ERR_INVALID, "input and output on different devices");
    CorrNdParams p;
    std::memset(&p, 0, sizeof p);
    int64_t K = 1;
    for (int d = 0; d < in->ndim; ++d) {
        if (in->shape[d] != out->shape[d]) return fail(SEPFILT_ERR_INVALID, "output shape not correct");
        if (wshape[d] < 1) return fail(SEPFILT_ERR_INVALID, "filter weights array has incorrect shape");
        const int before = wshape[d] / 2 + origin[d];
        if (before < 0 || before >= wshape[d]) return fail(SEPFILT_ERR_VALUE, "invalid origin");
        K *= wshape[d];
        if (K > SEPFILT_MAX_TAPS) return fail(SEPFILT_ERR_INVALID, "more than %d filter weights", SEPFILT_MAX_TAPS);
        p.shape[d] = in->shape[d];
        p.istride[d] = in->stride_bytes[d];
        p.ostride[d] = out->stride_bytes[d];
        p.wshape[d] = wshape[d];
        p.before[d] = before;
    }
    p.total = numel(out);
    if (p.total == 0) return SEPFILT_OK;
    if (!in->ptr || !out->ptr) return fail(SEPFILT_ERR_INVALID, "NULL data pointer");
    DeviceGuard guard(in->device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    p.in = static_cast<const char*>(in->ptr);
    p.out = static_cast<char*>(out->ptr);
    p.in_dtype = in->dtype;
    p.out_dtype = out->dtype;
    p.ndim = in->ndim;
    p.K = (int32_t)K;
    p.mode = mode;
    p.cval = cval;
    if (K <= SEPFILT_PARAM_TAPS) {
        std::memcpy(p.w, weights, sizeof(double) * (size_t)K);
    } else {
        const size_t need = sizeof(double) * (size_t)K;
        if (!scratch || scratch_bytes < need)
            return fail(SEPFILT_ERR_SCRATCH, "more than %d filter weights need %zu bytes of device scratch",
                        SEPFILT_PARAM_TAPS, need);
        cudaError_t e = cudaMemcpyAsync(scratch, weights, need, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(weights)");
        p.wdev = static_cast<const double*>(scratch);
    }
    cudaError_t e = launch_correlate_nd(p, s);
    if (e != cudaSuccess) return fail_cuda(e, "correlate_nd launch");
    return SEPFILT_OK;
}

int sepfilt_copy_cast(const sepfilt_tensor* in, const sepfilt_tensor* out, void* stream)
{
    // a 1-tap identity correlation is a strided copy under the store's cast rules
    static const double one = 1.0;
    int rc;
    if ((rc = check_tensor(in, "input", false)) || (rc = check_tensor(out, "output", true))) return rc;
    sepfilt_tensor i1 = *in, o1 = *out;
    if (i1.ndim == 0) {
        i1.ndim = o1.ndim = 1;
        i1.shape[0] = o1.shape[0] = 1;
        i1.stride_bytes[0] = dtype_size(i1.dtype);
        o1.stride_bytes[0] = dtype_size(o1.dtype);
    }
    sepfilt_pass p;
    std::memset(&p, 0, sizeof p);
    p.axis = i1.ndim - 1;
    p.ntaps = 1;
    p.taps = &one;
    p.mode = SEPFILT_NEAREST;
    return sepfilt_correlate1d(&i1, &o1, &p, 0, 0.0, SEPFILT_ACC_F64_EXACT, nullptr, 0, stream);
}

}  // extern "C"
