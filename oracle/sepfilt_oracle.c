/*
 * sepfilt_oracle.c — CPU restatement of the separable-correlation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cupyimg_b200/ may import, link or call
 * this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg have an independent checker for the CUDA path.
 *
 * What it restates (reference = /root/reference/cupyimg, scipy = scipy.ndimage 1.18):
 *   - index remapping per boundary mode:     scipy/ndimage/_util.py:170-228
 *   - tap placement  off = K//2 + origin:    scipy/ndimage/_filters_core.py:10-11, _util.py:231-239
 *   - per-element sum  y = cast<Y>(sum_k cast<W>(x[remap]) * w[k]):
 *                                            scipy/ndimage/_filters_core.py:239-312, filters.py:498-511
 *   - accumulator dtype W = float64 (dtype_mode="ndimage"):  _util.py:28-40
 *   - constant mode -> cval:                 _filters_core.py:276-293
 * The reference is "a GPU port of scipy.ndimage" and its own tests compare with
 * scipy on the CPU (tests/test_ndimage_vs_scipy.py:24-111), so where the reference
 * leaves the arithmetic to CuPy's generic code (summation order, float->int cast)
 * this file follows scipy's compiled NI_Correlate1D / NI_UniformFilter1D:
 *   - symmetric / anti-symmetric / generic summation order,
 *   - running-sum uniform filter (App. C.3),
 *   - C casts for the store (App. C.4).
 * Pinned by tests/test_oracle.py against (i) the literal known-answer tables of the
 * reference test-suite (tests/golden/reference_kats.json), (ii) 2680 outputs of the reference
 * ITSELF, executed on the CPU from the kernel source its own generator emits
 * (tests/golden/make_reference_vectors.py -> reference_vectors.npz), (iii) scipy.ndimage run
 * in the same process, bit-for-bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).  -ffp-contract=off
 * matters: scipy's binary does not contract a*b+c into FMA.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { M_REFLECT = 0, M_CONSTANT = 1, M_NEAREST = 2, M_MIRROR = 3, M_WRAP = 4 };
enum { T_I8 = 0, T_U8, T_I16, T_U16, T_I32, T_U32, T_I64, T_U64, T_F32, T_F64, T_BOOL };

/* _util.py:170-228 — returns the source index, or -1 for "use cval" */
int64_t oracle_remap(int mode, int64_t ix, int64_t n)
{
    if (ix >= 0 && ix < n) return ix;
    switch (mode) {
    case M_REFLECT: /* _util.py:176-183 */
        if (ix < 0) ix = -1 - ix;
        ix %= 2 * n;
        return ix < 2 * n - 1 - ix ? ix : 2 * n - 1 - ix;
    case M_MIRROR: /* _util.py:184-196 */
        if (n == 1) return 0;
        if (ix < 0) ix = -ix;
        ix = 1 + (ix - 1) % (2 * n - 2);
        return ix < 2 * n - 2 - ix ? ix : 2 * n - 2 - ix;
    case M_NEAREST: /* _util.py:197-201 */
        return ix < 0 ? 0 : n - 1;
    case M_WRAP: /* grid-wrap, _util.py:202-209 */
        ix %= n;
        return ix < 0 ? ix + n : ix;
    default: /* constant, _util.py:219-225 */
        return -1;
    }
}

static size_t tsize(int t)
{
    switch (t) {
    case T_I8: case T_U8: case T_BOOL: return 1;
    case T_I16: case T_U16: return 2;
    case T_I32: case T_U32: case T_F32: return 4;
    default: return 8;
    }
}

static double load(const void* p, int t, int64_t i)
{
    switch (t) {
    case T_I8: return ((const int8_t*)p)[i];
    case T_U8: case T_BOOL: return ((const uint8_t*)p)[i];
    case T_I16: return ((const int16_t*)p)[i];
    case T_U16: return ((const uint16_t*)p)[i];
    case T_I32: return ((const int32_t*)p)[i];
    case T_U32: return ((const uint32_t*)p)[i];
    case T_I64: return (double)((const int64_t*)p)[i];
    case T_U64: return (double)((const uint64_t*)p)[i];
    case T_F32: return ((const float*)p)[i];
    default: return ((const double*)p)[i];
    }
}

/* the store: plain C casts, as scipy's CASE_COPY_LINE_TO_DATA does.
 * In range this is truncation toward zero; negative -> unsigned wraps on x86-64. */
static void store(void* p, int t, int64_t i, double v)
{
    switch (t) {
    case T_I8: ((int8_t*)p)[i] = (int8_t)v; break;
    case T_U8: ((uint8_t*)p)[i] = (uint8_t)(int32_t)v; break;
    case T_I16: ((int16_t*)p)[i] = (int16_t)v; break;
    case T_U16: ((uint16_t*)p)[i] = (uint16_t)(int32_t)v; break;
    case T_I32: ((int32_t*)p)[i] = (int32_t)v; break;
    case T_U32: ((uint32_t*)p)[i] = (uint32_t)(int64_t)v; break;
    case T_I64: ((int64_t*)p)[i] = (int64_t)v; break;
    case T_U64: ((uint64_t*)p)[i] = v < 0 ? (uint64_t)(int64_t)v : (uint64_t)v; break;
    case T_F32: ((float*)p)[i] = (float)v; break;
    default: ((double*)p)[i] = v; break;
    }
}

/*
 * One correlate1d pass on a C-contiguous array viewed as (outer, n, inner).
 * out line length n_out may be a window of the input line: out[p] <-> in[p + in_offset].
 * uniform != 0 selects scipy's uniform_filter1d running sum (taps ignored).
 * Returns 0, or -1 on bad arguments.
 */
int oracle_correlate1d_lines(const void* in, int in_t, void* out, int out_t,
                             int64_t outer, int64_t n_in, int64_t n_out, int64_t inner,
                             const double* w, int K, int origin, int mode, double cval,
                             int64_t in_offset, int uniform, int64_t line_lo, int64_t line_hi);

int oracle_correlate1d(const void* in, int in_t, void* out, int out_t,
                       int64_t outer, int64_t n_in, int64_t n_out, int64_t inner,
                       const double* w, int K, int origin, int mode, double cval,
                       int64_t in_offset, int uniform)
{
    return oracle_correlate1d_lines(in, in_t, out, out_t, outer, n_in, n_out, inner, w, K, origin,
                                    mode, cval, in_offset, uniform, 0, outer * inner);
}

/* Same, restricted to the lines [line_lo, line_hi) of the outer*inner lines, so that the
 * bench's multi-threaded cpu_baseline can give each host thread a disjoint range. */
int oracle_correlate1d_lines(const void* in, int in_t, void* out, int out_t,
                             int64_t outer, int64_t n_in, int64_t n_out, int64_t inner,
                             const double* w, int K, int origin, int mode, double cval,
                             int64_t in_offset, int uniform, int64_t line_lo, int64_t line_hi)
{
    if (K < 1 || K / 2 + origin < 0 || K / 2 + origin >= K) return -1; /* _util.py:98-102 */
    if (outer <= 0 || n_out <= 0 || inner <= 0) return 0;
    if (n_in <= 0) return -1;
    const int size1 = K / 2, size2 = K - size1 - 1;
    const int64_t before = size1 + origin; /* taps left of the output position */
    int symmetric = 0;
    if (!uniform && (K & 1)) { /* scipy NI_Correlate1D symmetry probe */
        symmetric = 1;
        for (int i = 1; i <= size1; i++)
            if (fabs(w[size1 + i] - w[size1 - i]) > DBL_EPSILON) { symmetric = 0; break; }
        if (!symmetric) {
            symmetric = -1;
            for (int i = 1; i <= size1; i++)
                if (fabs(w[size1 + i] + w[size1 - i]) > DBL_EPSILON) { symmetric = 0; break; }
        }
    }
    /* boundary-extended line buffer, like scipy's NI_LineBuffer */
    const int64_t ext = n_out + K - 1;
    double* line = (double*)malloc(sizeof(double) * (size_t)ext);
    if (!line) return -1;
    const double* fw = w ? w + size1 : NULL;
    if (line_lo < 0) line_lo = 0;
    if (line_hi > outer * inner) line_hi = outer * inner;
    for (int64_t ln = line_lo; ln < line_hi; ln++) {
        {
            const int64_t o = ln / inner, c = ln % inner;
            const int64_t ibase = o * n_in * inner + c;
            const int64_t obase = o * n_out * inner + c;
            for (int64_t e = 0; e < ext; e++) {
                int64_t src = oracle_remap(mode, e - before + in_offset, n_in);
                line[e] = src < 0 ? cval : load(in, in_t, ibase + src * inner);
            }
            if (uniform) { /* SURVEY App. C.3 */
                double tmp = 0.0;
                for (int l = 0; l < K; l++) tmp += line[l];
                store(out, out_t, obase, tmp / (double)K);
                for (int64_t l = 1; l < n_out; l++) {
                    tmp += line[l + K - 1] - line[l - 1];
                    store(out, out_t, obase + l * inner, tmp / (double)K);
                }
                continue;
            }
            const double* il = line + size1;
            for (int64_t l = 0; l < n_out; l++, il++) {
                double acc;
                if (symmetric > 0) {
                    acc = il[0] * fw[0];
                    for (int j = -size1; j < 0; j++) acc += (il[j] + il[-j]) * fw[j];
                } else if (symmetric < 0) {
                    acc = il[0] * fw[0];
                    for (int j = -size1; j < 0; j++) acc += (il[j] - il[-j]) * fw[j];
                } else {
                    acc = il[size2] * fw[size2];
                    for (int j = -size1; j < size2; j++) acc += il[j] * fw[j];
                }
                store(out, out_t, obase + l * inner, acc);
            }
        }
    }
    free(line);
    return 0;
}

/* generic_gradient_magnitude epilogue in the output dtype (filters.py:1187-1201):
 * op 0: acc = a*a; op 1: acc += a*a; op 2: acc = sqrt(acc) with an unsafe cast back;
 * op 3: acc += a (generic_laplace, filters.py:1024-1035). */
int oracle_gradmag_step(void* acc, const void* a, int64_t n, int t, int op)
{
    for (int64_t i = 0; i < n; i++) {
        switch (t) {
#define INTCASE(T, C)                                                        \
    case T: {                                                                \
        C* y = (C*)acc; const C* x = (const C*)a;                            \
        uint64_t xx = (uint64_t)(int64_t)x[i] * (uint64_t)(int64_t)x[i];     \
        if (op == 0) y[i] = (C)xx;                                           \
        else if (op == 1) y[i] = (C)((uint64_t)(int64_t)y[i] + xx);          \
        else if (op == 2) store(acc, t, i, sqrt((double)y[i]));              \
        else y[i] = (C)((uint64_t)(int64_t)y[i] + (uint64_t)(int64_t)x[i]);  \
    } break;
            INTCASE(T_I8, int8_t) INTCASE(T_U8, uint8_t) INTCASE(T_I16, int16_t)
            INTCASE(T_U16, uint16_t) INTCASE(T_I32, int32_t) INTCASE(T_U32, uint32_t)
            INTCASE(T_I64, int64_t) INTCASE(T_U64, uint64_t)
#undef INTCASE
        case T_F32: {
            float* y = (float*)acc; const float* x = (const float*)a;
            if (op == 0) y[i] = x[i] * x[i];
            else if (op == 1) y[i] = y[i] + x[i] * x[i];
            else if (op == 2) y[i] = sqrtf(y[i]);
            else y[i] = y[i] + x[i];
        } break;
        case T_F64: {
            double* y = (double*)acc; const double* x = (const double*)a;
            if (op == 0) y[i] = x[i] * x[i];
            else if (op == 1) y[i] = y[i] + x[i] * x[i];
            else if (op == 2) y[i] = sqrt(y[i]);
            else y[i] = y[i] + x[i];
        } break;
        default: return -1;
        }
    }
    return 0;
}

/* dtype conversion under the store rules (the "no axis filtered" copy branch, filters.py:663-664) */
int oracle_copy_cast(const void* in, int in_t, void* out, int out_t, int64_t n)
{
    if (tsize(in_t) == 0 || tsize(out_t) == 0) return -1;
    for (int64_t i = 0; i < n; i++) store(out, out_t, i, load(in, in_t, i));
    return 0;
}
