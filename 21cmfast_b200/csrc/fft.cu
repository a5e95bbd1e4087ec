/* fft.cu -- batched shared-memory Stockham FFTs and the 3-D r2c / c2r drivers (see fft.h). */
#include "fft.h"
#include "dist.h"

#include <map>
#include <tuple>
#include <vector>

/* ------------------------------------------------------------------ complex helpers */
DEV float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
/* multiply by -i*sign... : rot(a, s) = a * (s * i), s = +1 or -1 */
DEV float2 crot(float2 a, int s) { return s > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

/* ------------------------------------------------------------------ small DFTs
 * sign = -1 forward (exp(-i...)), +1 inverse. */
DEV void bfly2(float2 *v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
DEV void bfly4(float2 *v, int sign) {
    float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    float2 c = cadd(v[1], v[3]), d = crot(csub(v[1], v[3]), sign);
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
}
DEV void bfly8(float2 *v, int sign) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    bfly4(e, sign);
    bfly4(o, sign);
    /* W8^k * o[k], W8 = exp(sign * 2 pi i / 8) */
    float2 t1 = make_float2(h * (o[1].x - sign * o[1].y), h * (o[1].y + sign * o[1].x));
    float2 t2 = crot(o[2], sign);
    float2 t3 = make_float2(h * (-o[3].x - sign * o[3].y), h * (-o[3].y + sign * o[3].x));
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], t1);   v[5] = csub(e[1], t1);
    v[2] = cadd(e[2], t2);   v[6] = csub(e[2], t2);
    v[3] = cadd(e[3], t3);   v[7] = csub(e[3], t3);
}
DEV void bfly3(float2 *v, int sign) {
    const float s3 = 0.86602540378443864676f;
    float2 t1 = cadd(v[1], v[2]);
    float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    float2 d = csub(v[1], v[2]);
    float2 t3 = crot(make_float2(s3 * d.x, s3 * d.y), sign); /* sign*i*s3*(b-c) */
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
}
/* generic odd radix by direct summation; cs/sn hold cos/sin(2 pi m / R) */
template <int R> DEV void bfly_generic(float2 *v, int sign, const float *cs, const float *sn) {
    float2 out[R];
#pragma unroll
    for (int p = 0; p < R; p++) {
        float2 acc = v[0];
#pragma unroll
        for (int q = 1; q < R; q++) {
            int m = (p * q) % R;
            float2 w = make_float2(cs[m], sign * sn[m]);
            acc = cadd(acc, cmul(v[q], w));
        }
        out[p] = acc;
    }
#pragma unroll
    for (int p = 0; p < R; p++) v[p] = out[p];
}
#ifndef B200_EMU
__device__ const float c5_cs[5] = {1.f, 0.30901699437494742410f, -0.80901699437494742410f,
                                   -0.80901699437494742410f, 0.30901699437494742410f};
__device__ const float c5_sn[5] = {0.f, 0.95105651629515357212f, 0.58778525229247312917f,
                                   -0.58778525229247312917f, -0.95105651629515357212f};
__device__ const float c7_cs[7] = {1.f, 0.62348980185873353053f, -0.22252093395631440429f,
                                   -0.90096886790241912624f, -0.90096886790241912624f,
                                   -0.22252093395631440429f, 0.62348980185873353053f};
__device__ const float c7_sn[7] = {0.f, 0.78183148246802980871f, 0.97492791218182360702f,
                                   0.43388373911755812048f, -0.43388373911755812048f,
                                   -0.97492791218182360702f, -0.78183148246802980871f};
#else
static const float c5_cs[5] = {1.f, 0.30901699437494742410f, -0.80901699437494742410f,
                               -0.80901699437494742410f, 0.30901699437494742410f};
static const float c5_sn[5] = {0.f, 0.95105651629515357212f, 0.58778525229247312917f,
                               -0.58778525229247312917f, -0.95105651629515357212f};
static const float c7_cs[7] = {1.f, 0.62348980185873353053f, -0.22252093395631440429f,
                               -0.90096886790241912624f, -0.90096886790241912624f,
                               -0.22252093395631440429f, 0.62348980185873353053f};
static const float c7_sn[7] = {0.f, 0.78183148246802980871f, 0.97492791218182360702f,
                               0.43388373911755812048f, -0.43388373911755812048f,
                               -0.97492791218182360702f, -0.78183148246802980871f};
#endif

template <int R> DEV void small_dft(float2 *v, int sign);
template <> DEV void small_dft<2>(float2 *v, int) { bfly2(v); }
template <> DEV void small_dft<3>(float2 *v, int sign) { bfly3(v, sign); }
template <> DEV void small_dft<4>(float2 *v, int sign) { bfly4(v, sign); }
template <> DEV void small_dft<5>(float2 *v, int sign) { bfly_generic<5>(v, sign, c5_cs, c5_sn); }
template <> DEV void small_dft<7>(float2 *v, int sign) { bfly_generic<7>(v, sign, c7_cs, c7_sn); }
template <> DEV void small_dft<8>(float2 *v, int sign) { bfly8(v, sign); }

/* ------------------------------------------------------------------ one Stockham stage
 * Tile layout in shared memory: element i of line c lives at [i * Tp + c] (Tp = T + 1 pad).
 * Work item w -> (line c = w % T, butterfly j = w / T) so that a warp touches consecutive
 * addresses.  Ns = product of the radices already applied. */
template <int R>
DEV void stockham_stage(const float2 *src, float2 *dst, int n, int Ns, int T, int Tp,
                        const float2 *__restrict__ tw, int sign) {
    const int nb = n / R;
    const int tws = n / (Ns * R);
    for (int w = threadIdx.x; w < nb * T; w += blockDim.x) {
        const int c = w % T, j = w / T;
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; q++) v[q] = src[(j + q * nb) * Tp + c];
        if (Ns > 1) {
            const int t1 = k * tws;
#pragma unroll
            for (int q = 1; q < R; q++) {
                float2 t = ldg(&tw[q * t1]);
                if (sign > 0) t.y = -t.y;
                v[q] = cmul(v[q], t);
            }
        }
        small_dft<R>(v, sign);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; q++) dst[(j0 + q * Ns) * Tp + c] = v[q];
    }
}

/* generic prime radix (11..31) with local arrays: rare sizes only */
DEV void stockham_stage_any(int R, const float2 *src, float2 *dst, int n, int Ns, int T, int Tp,
                            const float2 *__restrict__ tw, int sign) {
    const int nb = n / R;
    const int tws = n / (Ns * R);
    const int rstep = n / R;
    for (int w = threadIdx.x; w < nb * T; w += blockDim.x) {
        const int c = w % T, j = w / T;
        const int k = j % Ns;
        float2 v[32];
        for (int q = 0; q < R; q++) {
            float2 a = src[(j + q * nb) * Tp + c];
            if (q > 0 && Ns > 1) {
                float2 t = ldg(&tw[q * k * tws]);
                if (sign > 0) t.y = -t.y;
                a = cmul(a, t);
            }
            v[q] = a;
        }
        const int j0 = (j - k) * R + k;
        for (int p = 0; p < R; p++) {
            float2 acc = v[0];
            for (int q = 1; q < R; q++) {
                float2 t = ldg(&tw[((p * q) % R) * rstep]);
                if (sign > 0) t.y = -t.y;
                acc = cadd(acc, cmul(v[q], t));
            }
            dst[(j0 + p * Ns) * Tp + c] = acc;
        }
    }
}

/* any larger prime radix: the direct sum straight out of shared memory (no local array), O(R^2) per butterfly.
   FFTW takes every length, so the reference does too; these sizes are rare and slow there as well. */
DEV void stockham_stage_big(int R, const float2 *src, float2 *dst, int n, int Ns, int T, int Tp,
                            const float2 *__restrict__ tw, int sign) {
    const int nb = n / R;
    const int tws = n / (Ns * R);
    const int rstep = n / R;
    for (int w = threadIdx.x; w < nb * T; w += blockDim.x) {
        const int c = w % T, j = w / T;
        const int k = j % Ns;
        const int j0 = (j - k) * R + k;
        for (int p = 0; p < R; p++) {
            float2 acc = src[j * Tp + c];
            for (int q = 1; q < R; q++) {
                float2 a = src[(j + q * nb) * Tp + c];
                if (Ns > 1) {
                    float2 t = ldg(&tw[q * k * tws]);
                    if (sign > 0) t.y = -t.y;
                    a = cmul(a, t);
                }
                float2 t = ldg(&tw[(int)(((long long)p * q) % R) * rstep]);
                if (sign > 0) t.y = -t.y;
                acc = cadd(acc, cmul(a, t));
            }
            dst[(j0 + p * Ns) * Tp + c] = acc;
        }
    }
}

/* all stages; returns the buffer holding the result.  BIG: the length has a prime factor above 31 -- a separate
   instantiation of the kernels, so that the code generated for every other length is what it was without it */
template <bool BIG>
DEV float2 *fft_tile(float2 *A, float2 *B, int n, int T, int Tp, const FftFactors &f,
                     const float2 *__restrict__ tw, int sign) {
    float2 *src = A, *dst = B;
    int Ns = 1;
    for (int s = 0; s < f.nf; s++) {
        const int R = f.r[s];
        switch (R) {
            case 8: stockham_stage<8>(src, dst, n, Ns, T, Tp, tw, sign); break;
            case 4: stockham_stage<4>(src, dst, n, Ns, T, Tp, tw, sign); break;
            case 2: stockham_stage<2>(src, dst, n, Ns, T, Tp, tw, sign); break;
            case 3: stockham_stage<3>(src, dst, n, Ns, T, Tp, tw, sign); break;
            case 5: stockham_stage<5>(src, dst, n, Ns, T, Tp, tw, sign); break;
            case 7: stockham_stage<7>(src, dst, n, Ns, T, Tp, tw, sign); break;
            default:
                if (BIG && R > 31) stockham_stage_big(R, src, dst, n, Ns, T, Tp, tw, sign);
                else stockham_stage_any(R, src, dst, n, Ns, T, Tp, tw, sign);
                break;
        }
        __syncthreads();
        Ns *= R;
        float2 *t = src; src = dst; dst = t;
    }
    return src;
}

/* ------------------------------------------------------------------ strided-axis kernel */
struct StridedArgs {
    int n;                 /* line length */
    long long line_stride; /* complex elements between consecutive points of a line */
    int ncols;             /* lines per group (adjacent lines are 1 element apart) */
    long long group_stride;
    int T, Tp, sign;
    float scale;
    FftFactors f;
    const float2 *tw;
    /* KMUL geometry: the line axis is x; column m of a group is (y, kz) = (m / nzc, m % nzc) */
    int kmul, filter_type, fast_window, nx, ny, nz, nzc, pitch;
    float R;
    double R_param, r_const, dkx, dky, dkz;
    int op, axis_a, axis_b;
    double op_factor;
    const float *wtab;
    const float *wtab3;      /* expanded [|nx|][|ny|][pitch] form of wtab, or null */
    long long w3_xstride;    /* (ny/2+1) * pitch */
    int y_off;               /* slab-decomposed boxes: global y index of the first local (y, kz) row */
    /* slab-decomposed transforms: the store stage IS the all-to-all transpose.  Point i of a line goes
       to rank i / sc_nl, at sc_base[rank] + (i % sc_nl) * sc_line_stride + group * sc_group_stride + col
       (sc_base[r] is rank r's receive buffer as mapped here: a peer pointer over NVLink, see dist.h) */
    int sc_on, sc_nl;
    float2 *sc_base[8];
    long long sc_line_stride, sc_group_stride;
};
DEV float2 *scatter_dst(const StridedArgs &a, int i, long long group, int col) {
    const int r = i / a.sc_nl;
    return a.sc_base[r] + ((long long)(i - r * a.sc_nl) * a.sc_line_stride + group * a.sc_group_stride + col);
}

/* index_to_k (indexing.h:116-120): double wavenumber of a grid index */
DEV double kd_of_index(int n, int dim, double dk) {
    const double buf = (n <= dim / 2) ? n : (n - dim);
    return buf * dk;
}

/* k-space multipliers of the x-pass load stage: derivative operator first (rounded to float),
   then the window -- see KMul in fft.h.  (i = x index, col = flattened (y, kz) with row pitch) */
DEV float2 apply_kmul(float2 v, int i, int col, const StridedArgs &a) {
    const int iyl = col / a.pitch, iz = col - iyl * a.pitch;
    const int iy = iyl + a.y_off;
    if (iz >= a.nzc) return v; /* pad column */
    if (a.op != KOP_NONE) {
        if (i == 0 && iy == 0 && iz == 0) {
            v = make_float2(0.f, 0.f);
        } else if (a.op == KOP_VELOCITY_F) {
            /* float wavenumbers straight from index_to_k's double value */
            const float kx = (float)kd_of_index(i, a.nx, a.dkx), ky = (float)kd_of_index(iy, a.ny, a.dky),
                        kz = (float)kd_of_index(iz, a.nz, a.dkz);
            const float ksq = kmag_sq_f(kx, ky, kz);
            const float ka = a.axis_a == 0 ? kx : (a.axis_a == 1 ? ky : kz);
            const double g = a.op_factor * (double)ka / (double)ksq;
            v = make_float2((float)(-(double)v.y * g), (float)((double)v.x * g)); /* (re + i im) * (i g) */
        } else {
            const double kx = kd_of_index(i, a.nx, a.dkx), ky = kd_of_index(iy, a.ny, a.dky),
                         kz = kd_of_index(iz, a.nz, a.dkz);
            const double ksq = kx * kx + ky * ky + kz * kz;
            const double ka = a.axis_a == 0 ? kx : (a.axis_a == 1 ? ky : kz);
            if (a.op == KOP_GRADIENT_D) {
                const double g = ka / ksq;
                v = make_float2((float)(-(double)v.y * g), (float)((double)v.x * g));
            } else {
                const double kb = a.axis_b == 0 ? kx : (a.axis_b == 1 ? ky : kz);
                const double g = -ka * kb / ksq;
                v = make_float2((float)((double)v.x * g), (float)((double)v.y * g));
            }
        }
    }
    if (a.kmul == KMUL_FILTER && a.wtab) {
        const int sx = (i > a.nx / 2) ? i - a.nx : i, sy = (iy > a.ny / 2) ? iy - a.ny : iy;
        const float W = ldg(&a.wtab[sx * sx + sy * sy + iz * iz]);
        v.x *= W;
        v.y *= W;
    } else if (a.kmul == KMUL_FILTER) {
        const float kx = kf_of_index(i, a.nx, a.dkx);
        const float ky = kf_of_index(iy, a.ny, a.dky);
        const float kz = (float)((double)iz * a.dkz);
        if (a.fast_window) {
            const float W = window_value_fast(a.filter_type, kmag_sq_f(kx, ky, kz), a.R);
            v.x *= W;
            v.y *= W;
        } else {
            const double W = window_value(a.filter_type, kmag_sq_f(kx, ky, kz), a.R, a.R_param, a.r_const);
            v.x = (float)((double)v.x * W);
            v.y = (float)((double)v.y * W);
        }
    }
    return v;
}

template <bool BIG>
__global__ void __launch_bounds__(256) fft_strided_kernel(const float2 *__restrict__ src,
                                                          float2 *__restrict__ dst,
                                                          StridedArgs a) {
    DYN_SMEM(float2, smem);
    float2 *A = smem, *B = smem + (size_t)a.n * a.Tp;
    const int col0 = blockIdx.x * a.T;
    const long long gbase = (long long)blockIdx.y * a.group_stride;
    for (int w = threadIdx.x; w < a.n * a.T; w += blockDim.x) {
        const int c = w % a.T, i = w / a.T;
        const int col = col0 + c;
        float2 v = make_float2(0.f, 0.f);
        if (col < a.ncols) {
            v = src[gbase + (long long)i * a.line_stride + col];
            if (a.kmul != KMUL_NONE || a.op != KOP_NONE) v = apply_kmul(v, i, col, a);
        }
        A[i * a.Tp + c] = v;
    }
    __syncthreads();
    const float2 *res = fft_tile<BIG>(A, B, a.n, a.T, a.Tp, a.f, a.tw, a.sign);
    for (int w = threadIdx.x; w < a.n * a.T; w += blockDim.x) {
        const int c = w % a.T, i = w / a.T;
        const int col = col0 + c;
        if (col < a.ncols) {
            float2 v = res[i * a.Tp + c];
            if (a.scale != 1.f) { v.x *= a.scale; v.y *= a.scale; }
            if (a.sc_on) *scatter_dst(a, i, blockIdx.y, col) = v;
            else dst[gbase + (long long)i * a.line_stride + col] = v;
        }
    }
}

/* ------------------------------------------------------------------ z-axis kernels */
struct ZArgs {
    int n, nzc, pitch, nrows, L, Tp;
    FftFactors f;
    const float2 *tw;
    float scale;
    int clip;
    float clip_lo, clip_hi;
    int *minmax_keys;           /* {min, max} order keys or null */
    long long real_row_stride;  /* floats between rows of the real side */
    float premul;
};

/* complex rows -> real rows (inverse).  Builds the full Hermitian line in shared memory. */
template <bool BIG>
__global__ void __launch_bounds__(256) fft_c2r_z_kernel(const float2 *__restrict__ src,
                                                        float *__restrict__ dst, ZArgs a) {
    DYN_SMEM(float2, smem);
    float2 *A = smem, *B = smem + (size_t)a.n * a.Tp;
    const long long row0 = (long long)blockIdx.x * a.L;
    const int n = a.n, nzc = a.nzc;
    for (int w = threadIdx.x; w < a.L * nzc; w += blockDim.x) {
        const int k = w % nzc, l = w / nzc;
        const long long row = row0 + l;
        float2 v = make_float2(0.f, 0.f);
        if (row < a.nrows) v = src[row * a.pitch + k];
        if (k == 0 || 2 * k == n) v.y = 0.f;
        A[k * a.Tp + l] = v;
        if (k > 0 && 2 * k < n) A[(n - k) * a.Tp + l] = make_float2(v.x, -v.y);
    }
    __syncthreads();
    const float2 *res = fft_tile<BIG>(A, B, n, a.L, a.Tp, a.f, a.tw, +1);
    float lmin = 3.0e38f, lmax = -3.0e38f;
    for (int w = threadIdx.x; w < a.L * n; w += blockDim.x) {
        const int z = w % n, l = w / n;
        const long long row = row0 + l;
        if (row < a.nrows) {
            float val = res[z * a.Tp + l].x * a.scale;
            lmin = fminf(lmin, val);
            lmax = fmaxf(lmax, val);
            if (a.clip) val = fmaxf(fminf(val, a.clip_hi), a.clip_lo);
            dst[row * a.real_row_stride + z] = val;
        }
    }
    if (a.minmax_keys) {
        /* block reduction through shared memory (reuse A: all reads of res are done), then one
           pair of integer atomics per CTA on order-preserving keys */
        __syncthreads();
        float *red = reinterpret_cast<float *>(smem);
        red[threadIdx.x] = lmin;
        red[blockDim.x + threadIdx.x] = lmax;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) {
                red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + s]);
                red[blockDim.x + threadIdx.x] =
                    fmaxf(red[blockDim.x + threadIdx.x], red[blockDim.x + threadIdx.x + s]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            atomic_min_i32(&a.minmax_keys[0], float_order_key(float_as_int_bits(red[0])));
            atomic_max_i32(&a.minmax_keys[1], float_order_key(float_as_int_bits(red[blockDim.x])));
        }
    }
}

/* real rows -> complex rows (forward) */
template <bool BIG>
__global__ void __launch_bounds__(256) fft_r2c_z_kernel(const float *__restrict__ src,
                                                        float2 *__restrict__ dst, ZArgs a) {
    DYN_SMEM(float2, smem);
    float2 *A = smem, *B = smem + (size_t)a.n * a.Tp;
    const long long row0 = (long long)blockIdx.x * a.L;
    const int n = a.n, nzc = a.nzc;
    for (int w = threadIdx.x; w < a.L * n; w += blockDim.x) {
        const int z = w % n, l = w / n;
        const long long row = row0 + l;
        float val = 0.f;
        if (row < a.nrows) {
            val = src[row * a.real_row_stride + z];
            if (a.premul != 1.f || a.clip) {
                /* prepare_box_for_filtering (IonisationBox.c:343-346): double product, clamp */
                double cc = (double)val * (double)a.premul;
                if (a.clip) cc = fmax(fmin(cc, (double)a.clip_hi), (double)a.clip_lo);
                val = (float)cc;
            }
        }
        A[z * a.Tp + l] = make_float2(val, 0.f);
    }
    __syncthreads();
    const float2 *res = fft_tile<BIG>(A, B, n, a.L, a.Tp, a.f, a.tw, -1);
    for (int w = threadIdx.x; w < a.L * nzc; w += blockDim.x) {
        const int k = w % nzc, l = w / nzc;
        const long long row = row0 + l;
        if (row < a.nrows) {
            float2 v = res[k * a.Tp + l];
            if (a.scale != 1.f) { v.x *= a.scale; v.y *= a.scale; }
            dst[row * a.pitch + k] = v;
        }
    }
}

/* ------------------------------------------------------------------ plans */
static bool has_big_prime(const FftFactors &f) {
    for (int i = 0; i < f.nf; i++)
        if (f.r[i] > 31) return true;
    return false;
}
static bool factorize(int n, FftFactors &f) {
    f.nf = 0;
    int e = 0;
    while (n % 2 == 0) { n /= 2; e++; }
    int n8 = e / 3, rem = e % 3;
    int n4 = 0, n2 = 0;
    if (rem == 1) { if (n8 >= 1) { n8--; n4 = 2; } else n2 = 1; }
    else if (rem == 2) n4 = 1;
    for (int i = 0; i < n8; i++) f.r[f.nf++] = 8;
    for (int i = 0; i < n4; i++) f.r[f.nf++] = 4;
    for (int i = 0; i < n2; i++) f.r[f.nf++] = 2;
    for (int p = 3; (long long)p * p <= n; p += 2)
        while (n % p == 0) {
            if (f.nf >= FFT_MAX_FACTORS) return false;
            f.r[f.nf++] = p;
            n /= p;
        }
    if (n > 1) { /* what is left is one prime (of any size: stockham_stage_big) */
        if (f.nf >= FFT_MAX_FACTORS) return false;
        f.r[f.nf++] = n;
    }
    return true;
}

static std::map<int, Fft1D> g_plans1d;
static std::map<std::tuple<int, int, int>, Fft3D *> g_plans3d;

static Fft1D &plan1d(int n) {
    auto it = g_plans1d.find(n);
    if (it != g_plans1d.end()) return it->second;
    Fft1D p;
    p.n = n;
    if (!factorize(n, p.f))
        b200_throw(B200_ValueError, "FFT length %d has more than %d prime factors", n, FFT_MAX_FACTORS);
    std::vector<float2> tw(n);
    for (int k = 0; k < n; k++) {
        double ph = -2.0 * M_PI * (double)k / (double)n;
        tw[k] = make_float2((float)cos(ph), (float)sin(ph));
    }
    p.tw = (float2 *)dev_alloc(sizeof(float2) * n);
    h2d(p.tw, tw.data(), sizeof(float2) * n);
    dev_sync();
    g_stats.h2d -= (long long)sizeof(float2) * n; /* plan tables are not per-call traffic */
    return g_plans1d.emplace(n, p).first->second;
}

Fft3D *fft_plan(int nx, int ny, int nz) {
    auto key = std::make_tuple(nx, ny, nz);
    auto it = g_plans3d.find(key);
    if (it != g_plans3d.end()) return it->second;
    Fft3D *p = new Fft3D();
    p->nx = nx; p->ny = ny; p->nz = nz; p->nzc = nz / 2 + 1;
    p->pitch = (p->nzc + 7) / 8 * 8;
    p->px = plan1d(nx); p->py = plan1d(ny); p->pz = plan1d(nz);
    g_plans3d[key] = p;
    return p;
}

void fft_plans_drop() {
    for (auto &kv : g_plans3d) delete kv.second;
    g_plans3d.clear();
    for (auto &kv : g_plans1d) dev_free(kv.second.tw);
    g_plans1d.clear();
}

/* tile width: as many lines per CTA as fit ~72 KB (3 CTAs/SM), else 110 KB, else 220 KB */
static int pick_tile(int n, int maxT) {
    const size_t budgets[3] = {72 * 1024, 110 * 1024, 220 * 1024};
    for (int b = 0; b < 3; b++)
        for (int T = maxT; T >= 1; T >>= 1) {
            size_t bytes = 2 * (size_t)n * (T + 1) * sizeof(float2);
            if (bytes <= budgets[b] && (T >= 4 || b == 2 || T == maxT)) return T;
        }
    b200_throw(B200_ValueError, "FFT length %d does not fit in shared memory", n);
}
static size_t tile_smem(int n, int T) {
    size_t b = 2 * (size_t)n * (T + 1) * sizeof(float2);
    return b < 2048 ? 2048 : b;
}

#ifndef B200_EMU
template <typename K> static void allow_smem(K kernel, size_t bytes) {
    static std::map<const void *, size_t> done;
    size_t &cur = done[(const void *)kernel];
    if (bytes > cur && bytes > 32 * 1024) { /* static + dynamic may pass 48 KB before dynamic alone does */
        CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)bytes));
        cur = bytes;
    }
}
#include "fft_pow2.cuh"
#else
template <typename K> static void allow_smem(K, size_t) {}
static bool pow2_strided(const float2 *, float2 *, const StridedArgs &, int) { return false; }
static bool pow2_c2r_z(const float2 *, float *, const ZArgs &) { return false; }
static bool pow2_r2c_z(const float *, float2 *, const ZArgs &) { return false; }
#endif

struct Scatter {
    int nl = 0;               /* line points per destination rank */
    float2 *base[8];
    long long line_stride = 0, group_stride = 0;
};
static void run_strided(const Fft1D &p1, const float2 *src, float2 *dst, long long line_stride,
                        int ncols, long long group_stride, int ngroups, int sign, float scale,
                        const KMul *km, const Fft3D *p3, int y_off = 0, const Scatter *sc = nullptr) {
    StridedArgs a;
    memset(&a, 0, sizeof(a));
    a.y_off = y_off;
    if (sc) {
        a.sc_on = 1; a.sc_nl = sc->nl;
        for (int r = 0; r < 8; r++) a.sc_base[r] = sc->base[r];
        a.sc_line_stride = sc->line_stride; a.sc_group_stride = sc->group_stride;
    }
    a.n = p1.n; a.line_stride = line_stride; a.ncols = ncols; a.group_stride = group_stride;
    a.T = pick_tile(p1.n, 8); a.Tp = a.T + 1; a.sign = sign; a.scale = scale;
    a.f = p1.f; a.tw = p1.tw;
    a.kmul = KMUL_NONE;
    a.op = KOP_NONE;
    a.nx = p3->nx; a.ny = p3->ny; a.nz = p3->nz; a.nzc = p3->nzc; a.pitch = p3->pitch;
    if (km && km->active()) {
        a.kmul = km->kind; a.filter_type = km->filter_type;
        {
            static int exact = -1;
            if (exact < 0) { const char *e = getenv("B200_EXACT_WINDOW"); exact = (e && e[0] == '1') ? 1 : 0; }
            a.fast_window = (km->fast && !exact && (km->filter_type == 0 || km->filter_type == 2)) ? 1 : 0;
        }
        a.wtab = (a.fast_window && km->wtab) ? km->wtab : nullptr;
        a.wtab3 = a.wtab ? km->wtab3 : nullptr;
        a.w3_xstride = (long long)(p3->ny / 2 + 1) * p3->pitch;
        a.R = km->R; a.R_param = km->R_param; a.r_const = km->r_const;
        a.dkx = km->dk[0]; a.dky = km->dk[1]; a.dkz = km->dk[2];
        a.op = km->op; a.axis_a = km->axis_a; a.axis_b = km->axis_b; a.op_factor = km->op_factor;
    }
    if (pow2_strided(src, dst, a, ngroups)) return;
    size_t smem = tile_smem(a.n, a.T);
    dim3 grid((ncols + a.T - 1) / a.T, ngroups, 1);
    if (has_big_prime(a.f)) {
        allow_smem(fft_strided_kernel<true>, smem);
        B200_LAUNCH_T("fft_strided_kernel", fft_strided_kernel<true>, grid, 256, smem, src, dst, a);
    } else {
        allow_smem(fft_strided_kernel<false>, smem);
        B200_LAUNCH_T("fft_strided_kernel", fft_strided_kernel<false>, grid, 256, smem, src, dst, a);
    }
}

/* x-planes per (y pass, z pass) chunk of fft_c2r: the y pass leaves its output dirty in L2 and the z
   pass of the same planes reads it back from there, so the intermediate box never makes the round
   trip through HBM.  B200_FFT_CHUNK_MB sets the chunk size (0 = one chunk = whole box, the default). */
static int c2r_chunk_planes(const Fft3D *p) {
    static int mb = -1;
    if (mb < 0) {
        const char *e = getenv("B200_FFT_CHUNK_MB");
        mb = e ? atoi(e) : 0; /* measured on B200 at 512^3: per-chunk launches cost more than the L2 hits save */
    }
    const size_t plane = (size_t)p->ny * p->pitch * sizeof(float2);
    if (mb <= 0 || plane * p->nx <= (size_t)mb << 20) return p->nx;
    long long planes = ((long long)mb << 20) / (long long)plane;
    if (planes < 1) planes = 1;
    /* equal chunks */
    const int nchunks = (int)((p->nx + planes - 1) / planes);
    return (p->nx + nchunks - 1) / nchunks;
}

void fft_c2r(Fft3D *p, const float2 *src, float2 *work, const KMul &km, const ZEpilogue &epi) {
    const int nx = p->nx, ny = p->ny, pitch = p->pitch;
    /* x: lines over (y,kz) flattened (pad columns included), stride ny*pitch; multipliers ride on the load */
    run_strided(p->px, src, work, (long long)ny * pitch, ny * pitch, 0, 1, +1, 1.f, &km, p);
    ZArgs a;
    memset(&a, 0, sizeof(a));
    a.n = p->nz; a.nzc = p->nzc; a.pitch = pitch;
    a.L = pick_tile(p->nz, 8); a.Tp = a.L + 1;
    a.f = p->pz.f; a.tw = p->pz.tw;
    a.scale = epi.scale; a.clip = epi.clip; a.clip_lo = epi.clip_lo; a.clip_hi = epi.clip_hi;
    float *dst = epi.dst ? epi.dst : reinterpret_cast<float *>(work);
    a.real_row_stride = epi.dst ? epi.dst_row_stride : 2LL * pitch;
    a.minmax_keys = epi.minmax_keys;
    const int xc = c2r_chunk_planes(p);
    for (int x0 = 0; x0 < nx; x0 += xc) {
        const int planes = (x0 + xc <= nx) ? xc : nx - x0;
        float2 *w = work + (long long)x0 * ny * pitch;
        /* y: for each x, lines over kz, stride pitch */
        run_strided(p->py, w, w, pitch, pitch, (long long)ny * pitch, planes, +1, 1.f, nullptr, p);
        /* z: contiguous rows, complex -> real */
        a.nrows = planes * ny;
        float *d = dst + (long long)x0 * ny * a.real_row_stride;
        if (pow2_c2r_z(w, d, a)) continue;
        const int nblocks = (a.nrows + a.L - 1) / a.L;
        size_t smem = tile_smem(a.n, a.L);
        if (has_big_prime(a.f)) {
            allow_smem(fft_c2r_z_kernel<true>, smem);
            B200_LAUNCH_T("fft_c2r_z_kernel", fft_c2r_z_kernel<true>, dim3(nblocks), 256, smem, w, d, a);
        } else {
            allow_smem(fft_c2r_z_kernel<false>, smem);
            B200_LAUNCH_T("fft_c2r_z_kernel", fft_c2r_z_kernel<false>, dim3(nblocks), 256, smem, w, d, a);
        }
    }
}

void fft_r2c(Fft3D *p, float2 *box, const ZPrologue &pro) {
    const int nx = p->nx, ny = p->ny, pitch = p->pitch;
    ZArgs a;
    memset(&a, 0, sizeof(a));
    a.n = p->nz; a.nzc = p->nzc; a.pitch = pitch; a.nrows = nx * ny;
    a.L = pick_tile(p->nz, 8); a.Tp = a.L + 1;
    a.f = p->pz.f; a.tw = p->pz.tw;
    a.scale = 1.f; a.premul = pro.premul; a.clip = pro.clip;
    a.clip_lo = pro.clip_lo; a.clip_hi = pro.clip_hi;
    const float *src = pro.src ? pro.src : reinterpret_cast<const float *>(box);
    a.real_row_stride = pro.src ? pro.src_row_stride : 2LL * pitch;
    if (!pow2_r2c_z(src, box, a)) {
        const int nblocks = (a.nrows + a.L - 1) / a.L;
        size_t smem = tile_smem(a.n, a.L);
        if (has_big_prime(a.f)) {
            allow_smem(fft_r2c_z_kernel<true>, smem);
            B200_LAUNCH_T("fft_r2c_z_kernel", fft_r2c_z_kernel<true>, dim3(nblocks), 256, smem, src, box, a);
        } else {
            allow_smem(fft_r2c_z_kernel<false>, smem);
            B200_LAUNCH_T("fft_r2c_z_kernel", fft_r2c_z_kernel<false>, dim3(nblocks), 256, smem, src, box, a);
        }
    }
    run_strided(p->py, box, box, pitch, pitch, (long long)ny * pitch, nx, -1, 1.f, nullptr, p);
    run_strided(p->px, box, box, (long long)ny * pitch, ny * pitch, 0, 1, -1, pro.post_scale, nullptr, p);
}

/* ------------------------------------------------------------------ slab-decomposed transforms
 * One box over P ranks (replaces dft_r2c_cube / dft_c2r_cube, dft.c:18-72, for a
 * box that is split over GPUs).  Real space and the y/z passes live on x-slabs [nxl][ny][...]; k space
 * lives transposed, on y-slabs [nx][nyl][pitch], where the x pass (and every k-space multiplier) is
 * local.  The all-to-all between the two layouts is not a separate step: the pass in front of it
 * stores every point of a line straight into the owning rank's receive buffer (Scatter), local HBM for
 * the own block, NVLink peer stores for the others, and ONE stream-ordered barrier makes the blocks
 * visible.  Two receive buffers alternate, so the barrier of transform i also licenses the reuse of the
 * buffer of transform i - 1 (every rank has finished reading it before it arrives).  Per-line arithmetic is
 * that of the single-GPU kernels: the result is bit-identical to the undistributed transform. */
FftSlab fft_slab_setup(Fft3D *plan) {
    dist_require();
    FftSlab s;
    s.plan = plan;
    s.P = g_dist.world; s.rank = g_dist.rank;
    if (plan->nx % s.P || plan->ny % s.P)
        b200_throw(B200_ValueError, "slab FFT: HII_DIM = %d must be a multiple of the number of ranks (%d)", plan->nx, s.P);
    s.nxl = plan->nx / s.P; s.nyl = plan->ny / s.P;
    s.x0 = s.rank * s.nxl; s.y0 = s.rank * s.nyl;
    for (int b = 0; b < 2; b++) s.recv[b] = (float2 *)dist_alloc(s.n_cplx() * sizeof(float2));
    return s;
}

void fft_r2c_slab(FftSlab *s, float2 *kT, float2 *tmp, const ZPrologue &pro) {
    Fft3D *p = s->plan;
    const int ny = p->ny, nx = p->nx, pitch = p->pitch;
    /* z: the rank's real rows -> complex rows, x-slab layout */
    ZArgs a;
    memset(&a, 0, sizeof(a));
    a.n = p->nz; a.nzc = p->nzc; a.pitch = pitch; a.nrows = s->nxl * ny;
    a.L = pick_tile(p->nz, 8); a.Tp = a.L + 1;
    a.f = p->pz.f; a.tw = p->pz.tw;
    a.scale = 1.f; a.premul = pro.premul; a.clip = pro.clip;
    a.clip_lo = pro.clip_lo; a.clip_hi = pro.clip_hi;
    const float *src = pro.src ? pro.src : reinterpret_cast<const float *>(tmp);
    a.real_row_stride = pro.src ? pro.src_row_stride : 2LL * pitch;
    if (!pow2_r2c_z(src, tmp, a)) {
        const int nblocks = (a.nrows + a.L - 1) / a.L;
        size_t smem = tile_smem(a.n, a.L);
        if (has_big_prime(a.f)) {
            allow_smem(fft_r2c_z_kernel<true>, smem);
            B200_LAUNCH_T("fft_r2c_z_kernel", fft_r2c_z_kernel<true>, dim3(nblocks), 256, smem, src, tmp, a);
        } else {
            allow_smem(fft_r2c_z_kernel<false>, smem);
            B200_LAUNCH_T("fft_r2c_z_kernel", fft_r2c_z_kernel<false>, dim3(nblocks), 256, smem, src, tmp, a);
        }
    }
    /* y: lines over kz per local x plane; point y goes to rank y / nyl at [x0 + xl][y % nyl][kz] */
    float2 *recv = s->recv[g_dist.transforms & 1];
    g_dist.transforms++;
    Scatter sc;
    sc.nl = s->nyl; sc.line_stride = pitch; sc.group_stride = (long long)s->nyl * pitch;
    for (int r = 0; r < 8; r++) sc.base[r] = r < s->P ? dist_peer(recv, r) + (long long)s->x0 * s->nyl * pitch : nullptr;
    run_strided(p->py, tmp, tmp, pitch, pitch, (long long)ny * pitch, s->nxl, -1, 1.f, nullptr, p, 0, &sc);
    dist_barrier();
    /* x: lines over the local (y, kz) columns, out of the receive buffer into the caller's k-space box */
    run_strided(p->px, recv, kT, (long long)s->nyl * pitch, s->nyl * pitch, 0, 1, -1, pro.post_scale, nullptr, p, s->y0);
    (void)nx;
}

void fft_c2r_slab(FftSlab *s, const float2 *kT, float2 *work, const KMul &km, const ZEpilogue &epi) {
    Fft3D *p = s->plan;
    const int ny = p->ny, pitch = p->pitch;
    /* x (multipliers on the load): point x goes to rank x / nxl at [x % nxl][y0 + yl][kz] */
    float2 *recv = s->recv[g_dist.transforms & 1];
    g_dist.transforms++;
    Scatter sc;
    sc.nl = s->nxl; sc.line_stride = (long long)ny * pitch; sc.group_stride = 0;
    for (int r = 0; r < 8; r++) sc.base[r] = r < s->P ? dist_peer(recv, r) + (long long)s->y0 * pitch : nullptr;
    run_strided(p->px, kT, nullptr, (long long)s->nyl * pitch, s->nyl * pitch, 0, 1, +1, 1.f, &km, p, s->y0, &sc);
    dist_barrier();
    /* y: per local x plane, receive buffer -> work box */
    run_strided(p->py, recv, work, pitch, pitch, (long long)ny * pitch, s->nxl, +1, 1.f, nullptr, p);
    /* z: complex rows -> real rows with the epilogue */
    ZArgs a;
    memset(&a, 0, sizeof(a));
    a.n = p->nz; a.nzc = p->nzc; a.pitch = pitch; a.nrows = s->nxl * ny;
    a.L = pick_tile(p->nz, 8); a.Tp = a.L + 1;
    a.f = p->pz.f; a.tw = p->pz.tw;
    a.scale = epi.scale; a.clip = epi.clip; a.clip_lo = epi.clip_lo; a.clip_hi = epi.clip_hi;
    float *dst = epi.dst ? epi.dst : reinterpret_cast<float *>(work);
    a.real_row_stride = epi.dst ? epi.dst_row_stride : 2LL * pitch;
    a.minmax_keys = epi.minmax_keys;
    if (pow2_c2r_z(work, dst, a)) return;
    const int nblocks = (a.nrows + a.L - 1) / a.L;
    size_t smem = tile_smem(a.n, a.L);
    if (has_big_prime(a.f)) {
        allow_smem(fft_c2r_z_kernel<true>, smem);
        B200_LAUNCH_T("fft_c2r_z_kernel", fft_c2r_z_kernel<true>, dim3(nblocks), 256, smem, work, dst, a);
    } else {
        allow_smem(fft_c2r_z_kernel<false>, smem);
        B200_LAUNCH_T("fft_c2r_z_kernel", fft_c2r_z_kernel<false>, dim3(nblocks), 256, smem, work, dst, a);
    }
}

/* ------------------------------------------------------------------ in-place k-space window */
struct KWinArgs {
    int nx, ny, nz, nzc, pitch;
    int nyl, y_off; /* rows held locally per x and the global y index of the first one (slab layout) */
    int type;
    float R;
    double R_param, r_const, dkx, dky, dkz;
    float2 *box;
};
/* filter_box (filtering.c:308-394) as a separate pass: box *= W(k R), rounded to float */
__global__ void kspace_window_kernel(KWinArgs a) {
    const long long rows = (long long)a.nx * a.nyl;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int ix = (int)(row / a.nyl), iy = (int)(row - (long long)ix * a.nyl) + a.y_off;
        const float kx = kf_of_index(ix, a.nx, a.dkx), ky = kf_of_index(iy, a.ny, a.dky);
        for (int iz = threadIdx.x; iz < a.nzc; iz += blockDim.x) {
            const float kz = (float)((double)iz * a.dkz);
            const double W = window_value(a.type, kmag_sq_f(kx, ky, kz), a.R, a.R_param, a.r_const);
            float2 v = a.box[row * a.pitch + iz];
            v.x = (float)((double)v.x * W);
            v.y = (float)((double)v.y * W);
            a.box[row * a.pitch + iz] = v;
        }
    }
}
void fft_apply_window(Fft3D *p, float2 *box, const KMul &km, int nyl, int y_off) {
    if (nyl <= 0) { nyl = p->ny; y_off = 0; }
    KWinArgs a = {p->nx, p->ny, p->nz, p->nzc, p->pitch, nyl, y_off, km.filter_type, km.R, km.R_param, km.r_const,
                  km.dk[0], km.dk[1], km.dk[2], box};
    const long long rows = (long long)p->nx * nyl;
    const int cap = dev_num_sms() * 16;
    B200_LAUNCH(kspace_window_kernel, (int)(rows < cap ? rows : cap), 128, 0, a);
}

/* ------------------------------------------------------------------ window table */
struct WTabArgs {
    int type, n;
    float R;
    double dk;
    float *out;
};
__global__ void window_table_kernel(WTabArgs a) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
        a.out[i] = window_of_n2(a.type, i, a.dk, a.R);
}
struct WExpandArgs {
    int nax, nay, pitch, nzc;
    const float *tab;
    float *out;
    int ay0, ay1; /* rows |ny| in [ay0, ay1] only (a y-slab of a slab-decomposed box reads no others) */
};
__global__ void window_expand_kernel(WExpandArgs a) {
    /* the table is symmetric in (ax, ay) for a cubic box: gather each row once, write it twice */
    const bool whole = a.ay0 == 0 && a.ay1 == a.nay - 1;
    const bool sym = whole && a.nax == a.nay;
    const int nsel = a.ay1 - a.ay0 + 1;
    const long long rows = (long long)a.nax * nsel;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const int ax = (int)(r / nsel), ay = a.ay0 + (int)(r - (long long)ax * nsel);
        if (sym && ax > ay) continue;
        const int n2 = ax * ax + ay * ay;
        const long long row = (long long)ax * a.nay + ay;
        const long long mirror = (long long)ay * a.nay + ax;
        for (int iz = threadIdx.x; iz < a.pitch; iz += blockDim.x) {
            const float w = iz < a.nzc ? ldg(&a.tab[n2 + iz * iz]) : 1.0f;
            a.out[row * a.pitch + iz] = w;
            if (sym && ax != ay) a.out[mirror * a.pitch + iz] = w;
        }
    }
}
size_t window_table3_size(const Fft3D *p) { return (size_t)(p->nx / 2 + 1) * (p->ny / 2 + 1) * p->pitch; }
void window_table_expand(const Fft3D *p, const float *tab, float *out3, int y_lo, int y_hi) {
    WExpandArgs a = {p->nx / 2 + 1, p->ny / 2 + 1, p->pitch, p->nzc, tab, out3, 0, p->ny / 2};
    if (y_hi >= y_lo) { /* |ny| range of the global rows y_lo .. y_hi */
        int lo = p->ny, hi = 0;
        for (int y = y_lo; y <= y_hi; y++) {
            const int ay = (y > p->ny / 2) ? p->ny - y : y;
            lo = ay < lo ? ay : lo;
            hi = ay > hi ? ay : hi;
        }
        a.ay0 = lo; a.ay1 = hi;
    }
    const long long rows = (long long)a.nax * (a.ay1 - a.ay0 + 1);
    const int cap = dev_num_sms() * 16;
    B200_LAUNCH(window_expand_kernel, (int)(rows < cap ? rows : cap), 256, 0, a);
}
int window_table_size(const Fft3D *p) { return 3 * (p->nx / 2) * (p->nx / 2) + 1; }
void window_table_build(const Fft3D *p, int type, float R, double dk, float *out) {
    WTabArgs a = {type, window_table_size(p), R, dk, out};
    B200_LAUNCH(window_table_kernel, (a.n + 255) / 256, 256, 0, a);
}
