/*
 * fft.h -- the library's own 3-D real<->complex FFT (replaces dft_r2c_cube / dft_c2r_cube,
 * reference dft.c:18-72, and carries filter_box, filtering.c:308-394, in its load stage).
 *
 * Layout: FFTW's in-place convention, which the reference uses everywhere
 * (indexing.h:84-98): complex box [nx][ny][nzc] (nzc = nz/2+1) aliasing the padded real box
 * [nx][ny][2*nzc] -- except that rows are pitched to a multiple of 8 complex values (Fft3D::pitch)
 * for alignment.  Transforms are unnormalised in both directions, like FFTW.
 *
 * Algorithm: three passes of batched 1-D Stockham autosort FFTs held entirely in shared memory
 * (mixed radix 8/4/2/3/5/7, unrolled primes up to 31, a direct-sum stage for larger primes), one pass per axis:
 *   x, y axes : strided lines; a CTA owns a tile of T adjacent lines so that every global
 *               row access is T*8 contiguous bytes
 *   z axis    : contiguous lines; the real<->complex conversion, scaling, clipping and the
 *               global min/max reduction ride along in the load/store stages.
 */
#pragma once
#include "rt.h"

#define FFT_MAX_FACTORS 16
struct FftFactors {
    int nf;
    int r[FFT_MAX_FACTORS];
};

struct Fft1D {
    int n = 0;
    FftFactors f;
    float2 *tw = nullptr; /* device, n entries: exp(-2 pi i k / n) */
};

struct Fft3D {
    int nx, ny, nz, nzc;
    int pitch; /* complex elements per (x,y) row in device memory: nzc rounded up to a multiple of 8
                  (64-byte rows: the strided passes move aligned 64/128-byte segments); the padded
                  real view of the same box has 2*pitch floats per row.  Internal layout only. */
    Fft1D px, py, pz;
    size_t n_real() const { return (size_t)nx * ny * nz; }
    size_t n_cplx() const { return (size_t)nx * ny * pitch; }
};

/* k-space multipliers applied while the x-pass of a c2r transform loads its lines.  A
   derivative operator (if any) is applied first and rounded to float, then the window (if any),
   which reproduces the reference's separate passes over the complex float box. */
enum { KMUL_NONE = 0, KMUL_FILTER = 1 };
enum {
    KOP_NONE = 0,
    KOP_VELOCITY_F = 1, /* delta_k * (c * i k_a / k^2), k in float  (PerturbedField.c:320-350) */
    KOP_GRADIENT_D = 2, /* delta_k * (i k_a / k^2), k in double, DC -> 0 (InitialConditions.c:240-267) */
    KOP_LAPLACIAN_D = 3 /* delta_k * (-k_a k_b / k^2), DC -> 0          (InitialConditions.c:269-297) */
};
struct KMul {
    int kind = KMUL_NONE;      /* window on/off */
    int filter_type = 0;       /* 0 real-space top-hat, 1 sharp-k, 2 gaussian, 3 exp-mfp, 4 shell */
    float R = 0.f;             /* filter_box takes float R (filtering.c:308) */
    double R_param = 0.;       /* mfp (type 3) or inner radius (type 4) */
    double r_const = 0.;       /* exp(-R/R_param) for type 3 */
    double dk[3] = {0, 0, 0};  /* 2 pi / box length per axis (filtering.c:310-314) */
    int fast = 0;              /* 1: single-precision window (window_value_fast) for the hot sweep */
    const float *wtab = nullptr; /* optional window table indexed by nx^2+ny^2+nz^2 (cubic boxes) */
    int wtab_n = 0;
    const float *wtab3 = nullptr; /* the same window expanded to [|nx|][|ny|][kz] rows (window_table_expand):
                                     the x pass reads it with the lanes along kz, i.e. coalesced */
    int op = KOP_NONE;         /* derivative operator */
    int axis_a = 0, axis_b = 0;
    double op_factor = 1.;     /* c in KOP_VELOCITY_F */
    bool active() const { return kind != KMUL_NONE || op != KOP_NONE; }
};

/* what the last (z) pass of a c2r transform does with each real value */
struct ZEpilogue {
    float scale = 1.f;          /* multiply (e.g. 1/VOLUME) */
    int clip = 0;               /* clamp to [clip_lo, clip_hi] after min/max were taken */
    float clip_lo = 0.f, clip_hi = 0.f;
    int *minmax_keys = nullptr; /* device int[2] = {min key, max key} of the unclipped values, combined with
                                   integer atomics on float_order_key(); caller presets {INT_MAX, INT_MIN} */
    float *dst = nullptr;       /* optional separate real destination (else in place, padded) */
    long long dst_row_stride = 0; /* floats between consecutive (x,y) rows of dst */
};

/* what the first (z) pass of an r2c transform does while loading real rows */
struct ZPrologue {
    const float *src = nullptr;   /* real source (else in place, padded) */
    long long src_row_stride = 0; /* floats between consecutive rows of src */
    float premul = 1.f;           /* value * premul, then clamp if clip */
    int clip = 0;
    float clip_lo = 0.f, clip_hi = 0.f;
    float post_scale = 1.f;       /* applied to the complex output of the whole 3-D transform */
};

Fft3D *fft_plan(int nx, int ny, int nz);
/* table of the window over |n|^2 for cubic boxes (see KMul::wtab) */
int window_table_size(const Fft3D *p);
void window_table_build(const Fft3D *p, int type, float R, double dk, float *out);
/* expanded form for the power-of-two x pass: out3[(ax * (ny/2+1) + ay) * pitch + kz] = tab[ax^2 + ay^2 + kz^2] */
size_t window_table3_size(const Fft3D *p);
/* y_lo <= y_hi: only the rows a y-slab [y_lo, y_hi] of the box reads are filled */
void window_table_expand(const Fft3D *p, const float *tab, float *out3, int y_lo = 0, int y_hi = -1);

/* box *= W(kR) in place (the exact double window of filter_box, rounded to float per mode); with
   nyl > 0 the box is a transposed k-space slab [nx][nyl][pitch] whose first row is global y = y_off */
void fft_apply_window(Fft3D *p, float2 *box, const KMul &km, int nyl = 0, int y_off = 0);
/* forward: real (padded or pro.src) -> complex in `box` */
void fft_r2c(Fft3D *p, float2 *box, const ZPrologue &pro);
/* inverse: complex `src` -> real in `work` (src may equal work); optional k-space multiplier */
void fft_c2r(Fft3D *p, const float2 *src, float2 *work, const KMul &km, const ZEpilogue &epi);

/* Slab-decomposed transforms of ONE box over the ranks of dist.h (see fft.cu): real space and work
   boxes are x-slabs [nxl][ny][...], k space is the transposed y-slab [nx][nyl][pitch]. */
struct FftSlab {
    Fft3D *plan = nullptr;
    int P = 1, rank = 0, nxl = 0, nyl = 0, x0 = 0, y0 = 0;
    float2 *recv[2] = {nullptr, nullptr}; /* symmetric receive buffers of the transpose */
    size_t n_cplx() const { return (size_t)nxl * plan->ny * plan->pitch; } /* = nx * nyl * pitch */
    size_t n_real() const { return (size_t)nxl * plan->ny * plan->nz; }
};
FftSlab fft_slab_setup(Fft3D *plan);
/* real x-slab (pro.src, or padded in tmp) -> transposed k-space slab kT; tmp is an x-slab work box */
void fft_r2c_slab(FftSlab *s, float2 *kT, float2 *tmp, const ZPrologue &pro);
/* transposed k-space slab -> real x-slab in `work` (or epi.dst); multipliers ride on the x-pass load */
void fft_c2r_slab(FftSlab *s, const float2 *kT, float2 *work, const KMul &km, const ZEpilogue &epi);

/* Restatement of the reference's window functions (filtering.c:18-117) with the arithmetic
   types of the -Ofast x86-64 build of filter_box (filtering.c:331-381): |k|^2 is accumulated in
   float, kR = (float)(sqrt((double)k2) * (double)R), the window is evaluated in double and the
   complex float mode is multiplied in double and rounded once. */
HD double window_value(int type, float kmag_sq, float R, double R_param, double r_const) {
    if (type == 0) {  /* real-space top-hat: 3 (sin x - x cos x) / x^3, Taylor below 1e-4 */
        float kRf = (float)(sqrt((double)kmag_sq) * (double)R);
        double x = (double)kRf;
        double x2 = x * x;
        if (x < 1e-4) return 1.0 - x2 * 0.1;
        double s, c;
        sincos(x, &s, &c);
        return (3.0 / (x2 * x)) * (s - x * c);
    } else if (type == 1) {  /* sharp-k: zero above kR = (9 pi / 2)^(1/3) */
        float kRf = (float)(sqrt((double)kmag_sq) * (double)R);
        return ((double)kRf * 0.413566994 > 1.0) ? 0.0 : 1.0;
    } else if (type == 2) {  /* gaussian: the reference passes (kR)^2 held in a float */
        float kR2 = kmag_sq * (R * R); /* -Ofast hoists R*R (float) out of the loop */
        return exp(-0.643 * 0.643 * (double)kR2 / 2.);
    } else if (type == 3) {  /* exponentially attenuated top-hat (Davies & Furlanetto) */
        double k = sqrt((double)kmag_sq);
        double kR = k * (double)R, ratio = R_param / (double)R;
        if (kR < 1e-4) {
            double r2 = ratio * ratio, r3 = r2 * ratio;
            double ts_0 = 6 * r3 - r_const * (6 * r3 + 6 * r2 + 3 * ratio);
            return ts_0 + (r_const * (2 * r2 + 0.5 * ratio) - 2 * ts_0 * r2) * kR * kR;
        }
        double r2 = ratio * ratio, s, c;
        sincos(kR, &s, &c);
        double f = (kR * kR * r2 + 2 * ratio + 1) * ratio * c;
        f += (kR * kR * (r2 - r2 * ratio) + ratio + 1) * s / kR;
        f *= r_const;
        f -= 2 * r2;
        double den = kR * ratio * kR * ratio + 1;
        f *= -3 * ratio / (den * den);
        return f;
    } else if (type == 4) {
        /* spherical shell.  filter_box calls spherical_shell_filter(k, R, R_param) whose
           parameters are (k, R_inner, R_outer) (filtering.c:106-117,376): R is the inner radius */
        double k = sqrt((double)kmag_sq);
        double ki = k * (double)R, ko = k * R_param;
        if (ko < 1e-4) {
            double q = (double)R / R_param;
            return 1. - ko * ko / 10 * (q * q * q * q * q - 1) / (q * q * q - 1);
        }
        return 3.0 / (ko * ko * ko - ki * ki * ki) * (sin(ko) - cos(ko) * ko - sin(ki) + cos(ki) * ki);
    }
    return 1.0;
}

/* Single-precision window for the ionisation sweep (top-hat and gaussian; sharp-k stays on the
   exact path).  |W_fast - W| <= ~3e-7 over the whole kR range, i.e. below the rounding noise the
   float32 FFT itself adds to the filtered field, at ~1/5 of the instruction count of the double
   evaluation.  The double path above remains the one test_filter and the IC code use and can be
   forced everywhere with B200_EXACT_WINDOW=1. */
DEV float window_value_fast(int type, float kmag_sq, float R) {
    if (type == 0) {
        const float x = sqrtf(kmag_sq) * R;
        if (x < 1.5f) { /* Taylor series of 3 (sin x - x cos x) / x^3: no cancellation */
            const float t = x * x;
            float w = 3.2119e-11f;
            w = fmaf(w, t, -5.7813e-9f);
            w = fmaf(w, t, 7.5156325e-7f);
            w = fmaf(w, t, -6.6137566e-5f);
            w = fmaf(w, t, 3.5714286e-3f);
            w = fmaf(w, t, -0.1f);
            return fmaf(w, t, 1.0f);
        }
        float s, c;
        sincosf(x, &s, &c);
        return 3.0f * fmaf(-x, c, s) / (x * x * x);
    }
    /* gaussian */
    const float kR2 = kmag_sq * (R * R);
    return expf(-0.5f * 0.643f * 0.643f * kR2);
}

/* Window as a function of the integer |n|^2 = nx^2 + ny^2 + nz^2 (cubic boxes): the exact double
   formula at k = dk sqrt(n2), rounded to float.  Differs from the per-mode evaluation only through
   the float rounding of the wavenumber components (<= 1 ulp of kR). */
HD float window_of_n2(int type, long long n2, double dk, float R) {
    const double k = dk * sqrt((double)n2);
    const float kmag_sq = (float)(k * k);
    if (type == 0) {
        const double x = (double)(float)(k * (double)R);
        if (x < 1e-4) return (float)(1.0 - x * x * 0.1);
        double sn, cs;
        sincos(x, &sn, &cs);
        return (float)((3.0 / (x * x * x)) * (sn - x * cs));
    }
    const float kR2 = kmag_sq * (R * R);
    return (float)exp(-0.643 * 0.643 * (double)kR2 / 2.);
}

/* float wavenumber of grid index n on an axis of `dim` cells: the reference computes
   (float)(n_signed * delta_k) with delta_k a double (filtering.c:338-352). */
HD float kf_of_index(int n, int dim, double dk) {
    int ns = (n > dim / 2) ? n - dim : n;
    return (float)((double)ns * dk);
}
DEV float kmag_sq_f(float kx, float ky, float kz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
}
