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
