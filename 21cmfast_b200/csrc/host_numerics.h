/*
 * host_numerics.h -- the scalar numerical methods the host glue needs.
 *
 * The reference obtains these from GSL (not vendored, version unpinned); the product restates the
 * published algorithms itself so that the host scalars agree with the reference to rounding:
 *   qag61      QUADPACK QAG with the 61-point Gauss-Kronrod pair (gsl_integration_qag(...,
 *              GSL_INTEG_GAUSS61) at cosmology.c:389,441 and hmf.c:628)
 *   CubicSpline  natural cubic spline (gsl_interp_cspline at heating_helper_progs.c:121-123)
 *   gauss_legendre  Gauss-Legendre nodes on [a,b] (gauleg, hmf.c:660-697)
 *   Mt19937 + polar Gaussian + sequential choose / shuffle (gsl_rng_mt19937, gsl_ran_ugaussian,
 *              gsl_ran_choose, gsl_ran_shuffle at rng.c:33-54, InitialConditions.c:126-127)
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "gk61_table.h"

namespace hostnum {

enum { QAG_OK = 0, QAG_EROUND = 18, QAG_ESING = 21, QAG_EMAXITER = 11, QAG_EBADTOL = 13,
       QAG_EFAILED = 5 };

/* one 61-point Gauss-Kronrod panel on [a,b]; returns Kronrod estimate, writes error terms */
template <typename F>
inline double gk61_panel(F &&f, double a, double b, double *abserr, double *resabs,
                         double *resasc) {
    const int n = 31;
    const double *xgk = GK61_XGK, *wgk = GK61_WGK, *wg = GK61_WG;
    double fv1[31], fv2[31];
    const double center = 0.5 * (a + b), half = 0.5 * (b - a), ahalf = std::fabs(half);
    const double fc = f(center);
    double rg = 0, rk = fc * wgk[n - 1], rabs = std::fabs(rk);
    for (int j = 0; j < (n - 1) / 2; j++) {
        const int jtw = 2 * j + 1;
        const double dx = half * xgk[jtw];
        const double f1 = f(center - dx), f2 = f(center + dx), fs = f1 + f2;
        fv1[jtw] = f1; fv2[jtw] = f2;
        rg += wg[j] * fs;
        rk += wgk[jtw] * fs;
        rabs += wgk[jtw] * (std::fabs(f1) + std::fabs(f2));
    }
    for (int j = 0; j < n / 2; j++) {
        const int jt = 2 * j;
        const double dx = half * xgk[jt];
        const double f1 = f(center - dx), f2 = f(center + dx);
        fv1[jt] = f1; fv2[jt] = f2;
        rk += wgk[jt] * (f1 + f2);
        rabs += wgk[jt] * (std::fabs(f1) + std::fabs(f2));
    }
    const double mean = rk * 0.5;
    double rasc = wgk[n - 1] * std::fabs(fc - mean);
    for (int j = 0; j < n - 1; j++)
        rasc += wgk[j] * (std::fabs(fv1[j] - mean) + std::fabs(fv2[j] - mean));
    double err = std::fabs((rk - rg) * half);
    rk *= half; rabs *= ahalf; rasc *= ahalf;
    if (rasc != 0 && err != 0) {
        const double scale = std::pow(200 * err / rasc, 1.5);
        err = scale < 1 ? rasc * scale : rasc;
    }
    const double eps = 2.2204460492503131e-16, tiny = 2.2250738585072014e-308;
    if (rabs > tiny / (50 * eps)) {
        const double min_err = 50 * eps * rabs;
        if (min_err > err) err = min_err;
    }
    *abserr = err; *resabs = rabs; *resasc = rasc;
    return rk;
}

/* the 61 abscissae of a panel in the order gk61_panel evaluates them */
inline void gk61_abscissae(double a, double b, double *x) {
    const int n = 31;
    const double center = 0.5 * (a + b), half = 0.5 * (b - a);
    int k = 0;
    x[k++] = center;
    for (int j = 0; j < (n - 1) / 2; j++) {
        const double dx = half * GK61_XGK[2 * j + 1];
        x[k++] = center - dx;
        x[k++] = center + dx;
    }
    for (int j = 0; j < n / 2; j++) {
        const double dx = half * GK61_XGK[2 * j];
        x[k++] = center - dx;
        x[k++] = center + dx;
    }
}

/* globally adaptive QAG on a panel evaluator pe(a, b, &abserr, &resabs, &resasc) -> Kronrod estimate
   (lets a caller cache what its integrand needs per panel); returns a QAG_* status */
template <typename PE>
inline int qag61_panels(PE &&pe, double a, double b, double epsabs, double epsrel, size_t limit,
                        double *result, double *abserr);

/* globally adaptive QAG; returns a QAG_* status */
template <typename F>
inline int qag61(F &&f, double a, double b, double epsabs, double epsrel, size_t limit,
                 double *result, double *abserr) {
    return qag61_panels([&](double pa, double pb, double *e, double *ra, double *rs) { return gk61_panel(f, pa, pb, e, ra, rs); },
                        a, b, epsabs, epsrel, limit, result, abserr);
}
template <typename PE>
inline int qag61_panels(PE &&pe, double a, double b, double epsabs, double epsrel, size_t limit,
                        double *result, double *abserr) {
    const double eps = 2.2204460492503131e-16, tiny = 2.2250738585072014e-308;
    *result = 0; *abserr = 0;
    if (epsabs <= 0 && (epsrel < 50 * eps || epsrel < 0.5e-28)) return QAG_EBADTOL;
    struct Seg { double a, b, r, e; };
    std::vector<Seg> segs;
    segs.reserve(64);
    double e0, ra0, rs0;
    const double r0 = pe(a, b, &e0, &ra0, &rs0);
    segs.push_back({a, b, r0, e0});
    double tol = std::fmax(epsabs, epsrel * std::fabs(r0));
    if (e0 <= 50 * eps * ra0 && e0 > tol) { *result = r0; *abserr = e0; return QAG_EROUND; }
    if ((e0 <= tol && e0 != rs0) || e0 == 0.0) { *result = r0; *abserr = e0; return QAG_OK; }
    if (limit == 1) { *result = r0; *abserr = e0; return QAG_EMAXITER; }
    double area = r0, errsum = e0;
    size_t iter = 1;
    int ro1 = 0, ro2 = 0, etype = 0;
    do {
        size_t im = 0;
        for (size_t k = 1; k < segs.size(); k++)
            if (segs[k].e > segs[im].e) im = k;
        const Seg s = segs[im];
        const double a1 = s.a, b1 = 0.5 * (s.a + s.b), a2 = b1, b2 = s.b;
        double e1, e2, ra1, ra2, rs1, rs2;
        const double r1 = pe(a1, b1, &e1, &ra1, &rs1);
        const double r2 = pe(a2, b2, &e2, &ra2, &rs2);
        const double r12 = r1 + r2, e12 = e1 + e2;
        errsum += e12 - s.e;
        area += r12 - s.r;
        if (rs1 != e1 && rs2 != e2) {
            const double delta = s.r - r12;
            if (std::fabs(delta) <= 1.0e-5 * std::fabs(r12) && e12 >= 0.99 * s.e) ro1++;
            if (iter >= 10 && e12 > s.e) ro2++;
        }
        tol = std::fmax(epsabs, epsrel * std::fabs(area));
        if (errsum > tol) {
            if (ro1 >= 6 || ro2 >= 20) etype = 2;
            const double tmp = (1 + 100 * eps) * (std::fabs(a2) + 1000 * tiny);
            if (std::fabs(a1) <= tmp && std::fabs(b2) <= tmp) etype = 3;
        }
        if (e2 > e1) { segs[im] = {a2, b2, r2, e2}; segs.push_back({a1, b1, r1, e1}); }
        else { segs[im] = {a1, b1, r1, e1}; segs.push_back({a2, b2, r2, e2}); }
        iter++;
    } while (iter < limit && !etype && errsum > tol);
    double sum = 0;
    for (const Seg &s : segs) sum += s.r;
    *result = sum; *abserr = errsum;
    if (errsum <= tol) return QAG_OK;
    if (etype == 2) return QAG_EROUND;
    if (etype == 3) return QAG_ESING;
    if (iter == limit) return QAG_EMAXITER;
    return QAG_EFAILED;
}

/* natural cubic spline through strictly increasing x */
class CubicSpline {
  public:
    void init(const std::vector<double> &x, const std::vector<double> &y);
    double eval(double x) const; /* NaN outside [x0, xn] */
    bool ready() const { return !x_.empty(); }
    void clear() { x_.clear(); y_.clear(); c_.clear(); }
    double xmax() const { return x_.back(); }
    const std::vector<double> &coeffs() const { return c_; } /* the c_i of y + t (b + t (c + t d)) */
  private:
    std::vector<double> x_, y_, c_;
};

/* n-point Gauss-Legendre rule on [a,b]; x,w are 1-based like the reference's arrays */
void gauss_legendre(double a, double b, int n, double *x, double *w);

/* MT19937 exactly as GSL seeds and scales it */
class Mt19937 {
  public:
    explicit Mt19937(unsigned long seed = 0) { this->seed(seed); }
    void seed(unsigned long s);
    uint32_t next();
    double uniform() { return next() / 4294967296.0; }
    double uniform_pos() { double x; do { x = uniform(); } while (x == 0); return x; }
    unsigned long uniform_int(unsigned long n);
    double ugaussian();
  private:
    uint32_t mt_[624];
    int mti_;
};

/* The generator each OpenMP thread of the reference owns: thread t gets type t mod 5 of
   {mt19937, gfsr4, cmrg, mrg, taus2} (rng.c:57-83), seeded and scaled to doubles as GSL does
   (gfsr4: Ziff 1998 four-tap shift register; cmrg: L'Ecuyer 1996; mrg: L'Ecuyer, Blouin & Couture
   1993; taus2: L'Ecuyer 1999 with the corrected seeding).  Used for seed parity with N_THREADS > 1. */
class ThreadRng {
  public:
    ThreadRng(int thread_index, unsigned long seed);
    double uniform();
    double uniform_pos() { double x; do { x = uniform(); } while (x == 0); return x; }
    double ugaussian(); /* polar Box-Muller, gsl_ran_ugaussian */
  private:
    int kind_;
    Mt19937 mt_;
    std::vector<uint32_t> ring_; /* gfsr4: 2^14 words */
    int ring_pos_ = 0;
    long lag_[6] = {0, 0, 0, 0, 0, 0}; /* cmrg: x1..x3, y1..y3; mrg: x1..x5 */
    uint32_t taus_[3] = {0, 0, 0};
    unsigned long next_cmrg();
    unsigned long next_mrg();
    uint32_t next_taus();
    uint32_t next_gfsr4();
};

/* [begin, end) of the iterations 0..n-1 that thread t of n_threads runs under the static schedule of
   `#pragma omp for` (libgomp: the first n % n_threads threads take one extra iteration) */
inline void omp_static_range(int n, int n_threads, int t, int *begin, int *end) {
    const int q = n / n_threads, r = n % n_threads;
    *begin = t < r ? t * (q + 1) : r * (q + 1) + (t - r) * q;
    *end = *begin + (t < r ? q + 1 : q);
}

/* the seed the reference derives for thread 0 from the user seed (rng.c:31-54 with N_THREADS=1):
   sequential selection of 1 of INT_MAX/16 integers, then a (trivial) shuffle */
unsigned int derive_thread_seeds(unsigned long long seed, int n_threads, unsigned int *out);

}  // namespace hostnum
