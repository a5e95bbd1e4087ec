/* host_numerics.cpp -- see host_numerics.h */
#include "host_numerics.h"

#include <map>
#include <mutex>

#include <climits>
#include <limits>

namespace hostnum {

void CubicSpline::init(const std::vector<double> &x, const std::vector<double> &y) {
    x_ = x; y_ = y;
    const size_t n = x.size();
    c_.assign(n, 0.0);
    if (n < 3) return;
    const size_t m = n - 2; /* interior unknowns c[1..n-2] */
    std::vector<double> diag(m), off(m), rhs(m);
    for (size_t i = 0; i < m; i++) {
        const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        const double d0 = y[i + 1] - y[i], d1 = y[i + 2] - y[i + 1];
        off[i] = h1;
        diag[i] = 2.0 * (h0 + h1);
        rhs[i] = 3.0 * (d1 / h1 - d0 / h0);
    }
    /* Thomas algorithm on the symmetric tridiagonal system */
    std::vector<double> cp(m), dp(m);
    cp[0] = off[0] / diag[0];
    dp[0] = rhs[0] / diag[0];
    for (size_t i = 1; i < m; i++) {
        const double den = diag[i] - off[i - 1] * cp[i - 1];
        cp[i] = off[i] / den;
        dp[i] = (rhs[i] - off[i - 1] * dp[i - 1]) / den;
    }
    c_[m] = dp[m - 1];
    for (size_t i = m - 1; i-- > 0;) c_[i + 1] = dp[i] - cp[i] * c_[i + 2];
}

double CubicSpline::eval(double x) const {
    const size_t n = x_.size();
    if (n < 2 || x < x_[0] || x > x_[n - 1]) return std::numeric_limits<double>::quiet_NaN();
    size_t lo = 0, hi = n - 1;
    while (hi > lo + 1) {
        const size_t mid = (lo + hi) / 2;
        if (x_[mid] > x) hi = mid; else lo = mid;
    }
    const double dx = x_[lo + 1] - x_[lo], dy = y_[lo + 1] - y_[lo], t = x - x_[lo];
    const double b = dy / dx - dx * (c_[lo + 1] + 2.0 * c_[lo]) / 3.0;
    const double d = (c_[lo + 1] - c_[lo]) / (3.0 * dx);
    return y_[lo] + t * (b + t * (c_[lo] + t * d));
}

void gauss_legendre(double a, double b, int n, double *x, double *w) {
    /* gauleg (hmf.c:660-697).  The roots z_i of P_n and the derivatives pp_i do not depend on the
       interval: they are found once per n (same Newton iteration, same values) and only the final
       affine map is redone per call -- the ionisation ladder calls this once per filter radius. */
    struct Roots { std::vector<double> z, pp; };
    static std::map<int, Roots> cache;
    static std::mutex mu;
    const int m = (n + 1) / 2;
    const Roots *r;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(n);
        if (it == cache.end()) {
            Roots fresh;
            fresh.z.resize(m + 1);
            fresh.pp.resize(m + 1);
            for (int i = 1; i <= m; i++) {
                double z = std::cos(M_PI * (i - 0.25) / (n + 0.5)), z1, pp;
                int iter = 0;
                do {
                    double p1 = 1.0, p2 = 0.0;
                    for (int j = 1; j <= n; j++) {
                        const double p3 = p2;
                        p2 = p1;
                        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
                    }
                    pp = n * (z * p1 - p2) / (z * z - 1.0);
                    z1 = z;
                    z = z1 - p1 / pp;
                } while (std::fabs(z - z1) > 3.0e-11 && ++iter < 100);
                fresh.z[i] = z;
                fresh.pp[i] = pp;
            }
            it = cache.emplace(n, std::move(fresh)).first;
        }
        r = &it->second;
    }
    const double xm = 0.5 * (b + a), xl = 0.5 * (b - a);
    for (int i = 1; i <= m; i++) {
        const double z = r->z[i], pp = r->pp[i];
        x[i] = xm - xl * z;
        x[n + 1 - i] = xm + xl * z;
        w[i] = 2.0 * xl / ((1.0 - z * z) * pp * pp);
        w[n + 1 - i] = w[i];
    }
}

void Mt19937::seed(unsigned long s) {
    if (s == 0) s = 4357;
    mt_[0] = (uint32_t)(s & 0xffffffffUL);
    for (int i = 1; i < 624; i++)
        mt_[i] = 1812433253U * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
    mti_ = 624;
}
uint32_t Mt19937::next() {
    if (mti_ >= 624) {
        for (int k = 0; k < 624; k++) {
            const uint32_t y = (mt_[k] & 0x80000000U) | (mt_[(k + 1) % 624] & 0x7fffffffU);
            mt_[k] = mt_[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
        }
        mti_ = 0;
    }
    uint32_t k = mt_[mti_++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680U;
    k ^= (k << 15) & 0xefc60000U;
    k ^= (k >> 18);
    return k;
}
unsigned long Mt19937::uniform_int(unsigned long n) {
    const unsigned long range = 0xffffffffUL, scale = range / n;
    unsigned long k;
    do { k = next() / scale; } while (k >= n);
    return k;
}
double Mt19937::ugaussian() {
    double x, y, r2;
    do {
        x = -1 + 2 * uniform_pos();
        y = -1 + 2 * uniform_pos();
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return y * std::sqrt(-2.0 * std::log(r2) / r2);
}

/* ---- the other four per-thread generator types (rng.c:57-83) ---- */
namespace {
inline unsigned long lcg69069(unsigned long n) { return (69069UL * n) & 0xffffffffUL; }
/* (a x) mod m without overflow by Schrage's decomposition m = a q + r */
inline long schrage(long a, long q, long r, long x) { const long h = x / q; return a * (x - h * q) - h * r; }
}  // namespace

uint32_t ThreadRng::next_taus() {
    auto step = [](uint32_t s, int a, int b, uint32_t c, int d) -> uint32_t { return ((s & c) << d) ^ (((s << a) ^ s) >> b); };
    taus_[0] = step(taus_[0], 13, 19, 4294967294U, 12);
    taus_[1] = step(taus_[1], 2, 25, 4294967288U, 4);
    taus_[2] = step(taus_[2], 3, 11, 4294967280U, 17);
    return taus_[0] ^ taus_[1] ^ taus_[2];
}
uint32_t ThreadRng::next_gfsr4() {
    const int M = 16383;
    ring_pos_ = (ring_pos_ + 1) & M;
    const uint32_t v = ring_[(ring_pos_ + M + 1 - 471) & M] ^ ring_[(ring_pos_ + M + 1 - 1586) & M] ^
                       ring_[(ring_pos_ + M + 1 - 6988) & M] ^ ring_[(ring_pos_ + M + 1 - 9689) & M];
    ring_[ring_pos_] = v;
    return v;
}
unsigned long ThreadRng::next_cmrg() {
    const long m1 = 2147483647, m2 = 2145483479;
    long *x = lag_, *y = lag_ + 3;
    long p3 = schrage(183326, 11714, 2883, x[2]), p2 = schrage(63308, 33921, 12979, x[1]);
    if (p3 < 0) p3 += m1;
    if (p2 < 0) p2 += m1;
    x[2] = x[1]; x[1] = x[0]; x[0] = p2 - p3;
    if (x[0] < 0) x[0] += m1;
    long q3 = schrage(539608, 3976, 2071, y[2]), q1 = schrage(86098, 24919, 7417, y[0]);
    if (q3 < 0) q3 += m2;
    if (q1 < 0) q1 += m2;
    y[2] = y[1]; y[1] = y[0]; y[0] = q1 - q3;
    if (y[0] < 0) y[0] += m2;
    return (unsigned long)(x[0] < y[0] ? x[0] - y[0] + m1 : x[0] - y[0]);
}
unsigned long ThreadRng::next_mrg() {
    const long m = 2147483647;
    long *x = lag_;
    long p5 = schrage(104480, 20554, 1727, x[4]), p1 = schrage(107374182, 20, 7, x[0]);
    if (p5 > 0) p5 -= m;
    if (p1 < 0) p1 += m;
    x[4] = x[3]; x[3] = x[2]; x[2] = x[1]; x[1] = x[0];
    x[0] = p1 + p5;
    if (x[0] < 0) x[0] += m;
    return (unsigned long)x[0];
}
ThreadRng::ThreadRng(int thread_index, unsigned long s) : kind_(thread_index % 5), mt_(0) {
    switch (kind_) {
    case 0:
        mt_.seed(s);
        break;
    case 1: { /* every word is assembled from the top bits of 32 LCG steps; 32 words are then forced independent */
        if (s == 0) s = 4357;
        ring_.assign(16384, 0);
        for (auto &w : ring_) {
            uint32_t t = 0;
            for (uint32_t bit = 0x80000000U; bit; bit >>= 1) {
                s = lcg69069(s);
                if (s & 0x80000000UL) t |= bit;
            }
            w = t;
        }
        uint32_t msb = 0x80000000U, mask = 0xffffffffU;
        for (int i = 0; i < 32; i++, mask >>= 1, msb >>= 1) ring_[7 + 3 * i] = (ring_[7 + 3 * i] & mask) | msb;
        ring_pos_ = 32;
        break;
    }
    case 2: {
        const long m1 = 2147483647, m2 = 2145483479;
        if (s == 0) s = 1;
        for (int i = 0; i < 6; i++) { s = lcg69069(s); lag_[i] = (long)(s % (unsigned long)(i < 3 ? m1 : m2)); }
        for (int i = 0; i < 7; i++) next_cmrg();
        break;
    }
    case 3: {
        if (s == 0) s = 1;
        for (int i = 0; i < 5; i++) { s = lcg69069(s); lag_[i] = (long)(s % 2147483647UL); }
        for (int i = 0; i < 6; i++) next_mrg();
        break;
    }
    default: {
        if (s == 0) s = 1;
        const uint32_t floor_[3] = {2, 8, 16};
        unsigned long v = s;
        for (int i = 0; i < 3; i++) {
            v = lcg69069(v);
            if (v < floor_[i]) v += floor_[i];
            taus_[i] = (uint32_t)v;
        }
        for (int i = 0; i < 6; i++) next_taus();
    }
    }
}
double ThreadRng::uniform() {
    switch (kind_) {
    case 0: return mt_.uniform();
    case 1: return next_gfsr4() / 4294967296.0;
    case 2: return next_cmrg() / 2147483647.0;
    case 3: return next_mrg() / 2147483647.0;
    default: return next_taus() / 4294967296.0;
    }
}
double ThreadRng::ugaussian() {
    double x, y, r2;
    do {
        x = -1 + 2 * uniform_pos();
        y = -1 + 2 * uniform_pos();
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return y * std::sqrt(-2.0 * std::log(r2) / r2);
}

unsigned int derive_thread_seeds(unsigned long long seed, int n_threads, unsigned int *out) {
    /* rng.c:31-54: choose n_threads of the integers 0..INT_MAX/16-1 by sequential selection
       (an integer i is taken when (n-i) U < k-j), then Fisher-Yates shuffle the picks.  The big
       integer array of the reference is never materialised: src[i] == i. */
    Mt19937 r((unsigned long)seed);
    const size_t n = (size_t)(INT_MAX / 16), k = (size_t)n_threads;
    size_t j = 0;
    for (size_t i = 0; i < n && j < k; i++)
        if ((double)(n - i) * r.uniform() < (double)(k - j)) out[j++] = (unsigned int)i;
    for (size_t i = k - 1; i > 0 && k > 0; i--) {
        const size_t jj = r.uniform_int(i + 1);
        const unsigned int t = out[i]; out[i] = out[jj]; out[jj] = t;
    }
    return out[0];
}

}  // namespace hostnum
