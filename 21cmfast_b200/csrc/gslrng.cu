/*
 * gslrng.cu -- the reference's Gaussian random stream on the device, in stream order.
 *
 * sample_ic_modes (InitialConditions.c:103-139) draws two gsl_ran_ugaussian() per mode from ONE
 * gsl_rng_mt19937 generator (N_THREADS = 1; rng.c:31-90 seeds it), in x-major mode order.  GSL is
 * not part of the reference tree; the algorithms restated here are the published ones (same
 * restatement as hostnum::Mt19937 in host_numerics.cpp, which the oracle shims pin against numpy's
 * legacy MT19937 seeding):
 *   mt19937       : 624-word state, x[k+624] = x[k+397] ^ twist(x[k], x[k+1]), tempering on output;
 *   uniform_pos   : u = x / 2^32, redrawn while u == 0;
 *   ugaussian     : polar Box-Muller -- x = -1 + 2 u1, y = -1 + 2 u2, r2 = x^2 + y^2, redrawn while
 *                   r2 > 1 or r2 == 0; returns y sqrt(-2 ln r2 / r2) (the partner value is discarded).
 *
 * The stream is sequential by construction, but only weakly so:
 *   1. raw words: within one 624-word refresh, words 0..226 depend only on the old state, words
 *      227..453 on the old state and on new words 0..226, words 454..623 on new words 227..396 (and
 *      new word 0): three data-parallel steps per refresh, run by ONE CTA (the generator itself is a
 *      single chain; measured ~0.75 G words/s on one SM, while every other stage uses the whole GPU);
 *   2. every polar attempt consumes exactly two words, accepted or not, so attempt i always owns
 *      words 2i, 2i+1: acceptance is data-parallel, and the j-th Gaussian is the j-th accepted
 *      attempt -- an exclusive prefix sum over the accept flags gives each accepted attempt its place;
 *   3. a zero word (probability 2^-32 each) would shift the pairing: chunks that contain one are
 *      converted by the sequential host routine instead (gaussians_from_raw_host), bit-for-bit the
 *      same rule.
 * Result: the same doubles as the host generator up to the last bit of the device's log().
 */
#include "rt.h"
#include "host_numerics.h"

#include <vector>

#define MT_N 624
#define MT_M 397

struct GslStream {
    unsigned int *d_state = nullptr; /* 624 words */
    unsigned int *d_raw = nullptr;   /* tempered output words not yet consumed: [pos, n_raw) */
    size_t cap_raw = 0, n_raw = 0, pos = 0;
    int *d_counts = nullptr;         /* per-block accept counts / offsets */
    size_t cap_counts = 0;
    long long *d_info = nullptr;     /* [0] attempts consumed (or -1), [1] zero word seen */
    long long total_gaussians = 0;
};

/* ------------------------------------------------------------------ kernels */
struct MtGenArgs {
    unsigned int *state, *out;
    long long nblocks;
};
DEV unsigned int mt_twist(unsigned int a, unsigned int b, unsigned int m) {
    const unsigned int y = (a & 0x80000000U) | (b & 0x7fffffffU);
    return m ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
}
DEV unsigned int mt_temper(unsigned int k) {
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680U;
    k ^= (k << 15) & 0xefc60000U;
    k ^= (k >> 18);
    return k;
}
/* one CTA refreshes the state `nblocks` times, appending 624 tempered words per refresh (three
   barrier-separated steps of <= 227 independent words).  Measured on B200: ~0.75 G words/s, i.e.
   12.8 s of the 20.4 s an exact-stream DIM=1536 box takes; a single warp with __syncwarp instead of
   block barriers was tried and is 2.3x slower (too little latency hiding). */
#define MT_SYNC() __syncthreads()
__global__ void __launch_bounds__(256) mt19937_generate_kernel(MtGenArgs a) {
    __shared__ unsigned int cur[MT_N], nxt[MT_N];
    for (int k = threadIdx.x; k < MT_N; k += blockDim.x) cur[k] = a.state[k];
    MT_SYNC();
    for (long long b = 0; b < a.nblocks; b++) {
        for (int k = threadIdx.x; k < MT_N - MT_M; k += blockDim.x) /* 0..226: old words only */
            nxt[k] = mt_twist(cur[k], cur[k + 1], cur[k + MT_M]);
        MT_SYNC();
        for (int k = (MT_N - MT_M) + threadIdx.x; k < 2 * (MT_N - MT_M); k += blockDim.x) /* 227..453 */
            nxt[k] = mt_twist(cur[k], cur[k + 1], nxt[k - (MT_N - MT_M)]);
        MT_SYNC();
        for (int k = 2 * (MT_N - MT_M) + threadIdx.x; k < MT_N; k += blockDim.x) /* 454..623 */
            nxt[k] = mt_twist(cur[k], k + 1 < MT_N ? cur[k + 1] : nxt[0], nxt[k - (MT_N - MT_M)]);
        MT_SYNC();
        unsigned int *o = a.out + b * MT_N;
        for (int k = threadIdx.x; k < MT_N; k += blockDim.x) {
            const unsigned int v = nxt[k];
            cur[k] = v;
            o[k] = mt_temper(v);
        }
        MT_SYNC();
    }
    for (int k = threadIdx.x; k < MT_N; k += blockDim.x) a.state[k] = cur[k];
}

/* polar attempt on words (w1, w2): returns acceptance, y and r2 in double exactly as GSL computes them */
DEV bool polar_attempt(unsigned int w1, unsigned int w2, double &y, double &r2) {
    const double x = -1 + 2 * ((double)w1 / 4294967296.0);
    y = -1 + 2 * ((double)w2 / 4294967296.0);
    r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)); /* no FMA contraction: r2 decides acceptance and, near 1,
                                                          log(r2) amplifies its last bit */
    return !(r2 > 1.0 || r2 == 0);
}

#define GA_PER_THREAD 16 /* consecutive attempts per thread */
struct GaussArgs {
    const unsigned int *raw; /* first unconsumed word */
    long long n_attempts;    /* attempts available */
    long long want;          /* Gaussians to emit */
    int *counts;             /* [gridDim.x] */
    long long *info;
    double *out;
    int per_block;           /* attempts per block = blockDim.x * GA_PER_THREAD */
};
/* pass 1: accept count of every block of attempts; flags zero words */
__global__ void __launch_bounds__(256) gauss_count_kernel(GaussArgs a) {
    __shared__ int red[256];
    for (long long blk = blockIdx.x; blk * a.per_block < a.n_attempts; blk += gridDim.x) {
        const int per_thread = a.per_block / (int)blockDim.x; /* GA_PER_THREAD on the GPU */
        const long long i0 = blk * a.per_block + (long long)threadIdx.x * per_thread;
        int cnt = 0;
        for (int j = 0; j < per_thread; j++) {
            const long long i = i0 + j;
            if (i >= a.n_attempts) break;
            const unsigned int w1 = a.raw[2 * i], w2 = a.raw[2 * i + 1];
            if (w1 == 0U || w2 == 0U) a.info[1] = 1;
            double y, r2;
            cnt += polar_attempt(w1, w2, y, r2) ? 1 : 0;
        }
        red[threadIdx.x] = cnt;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) a.counts[blk] = red[0];
        __syncthreads();
    }
}
/* pass 2: exclusive scan of the block counts in place (one CTA; the block count is small) */
struct ScanArgs {
    int *counts;
    long long n;
    long long *total;
};
__global__ void __launch_bounds__(256) gauss_scan_kernel(ScanArgs a) {
    __shared__ long long part[256];
    const long long per = (a.n + blockDim.x - 1) / blockDim.x;
    const long long lo = (long long)threadIdx.x * per, hi = lo + per < a.n ? lo + per : a.n;
    long long s = 0;
    for (long long i = lo; i < hi; i++) s += a.counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < (int)blockDim.x; t++) { const long long v = part[t]; part[t] = run; run += v; }
        *a.total = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (long long i = lo; i < hi; i++) { const int v = a.counts[i]; a.counts[i] = (int)run; run += v; }
}
/* pass 3: place every accepted attempt; the attempt that yields Gaussian number `want` - 1 reports how
   many attempts the request consumed */
__global__ void __launch_bounds__(256) gauss_emit_kernel(GaussArgs a) {
    __shared__ int offs[256];
    for (long long blk = blockIdx.x; blk * a.per_block < a.n_attempts; blk += gridDim.x) {
        const int per_thread = a.per_block / (int)blockDim.x;
        const long long i0 = blk * a.per_block + (long long)threadIdx.x * per_thread;
        int cnt = 0;
        for (int j = 0; j < per_thread; j++) {
            const long long i = i0 + j;
            if (i >= a.n_attempts) break;
            double y, r2;
            cnt += polar_attempt(a.raw[2 * i], a.raw[2 * i + 1], y, r2) ? 1 : 0;
        }
        offs[threadIdx.x] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) { /* exclusive scan over the threads of the block */
            int run = 0;
            for (int t = 0; t < (int)blockDim.x; t++) { const int v = offs[t]; offs[t] = run; run += v; }
        }
        __syncthreads();
        long long g = (long long)a.counts[blk] + offs[threadIdx.x];
        for (int j = 0; j < per_thread; j++) {
            const long long i = i0 + j;
            if (i >= a.n_attempts) break;
            double y, r2;
            if (polar_attempt(a.raw[2 * i], a.raw[2 * i + 1], y, r2)) {
                if (g < a.want) {
                    a.out[g] = y * sqrt(-2.0 * log(r2) / r2);
                    if (g == a.want - 1) a.info[0] = i + 1;
                }
                g++;
            }
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------ host side */
/* sequential conversion with GSL's exact rules, including the zero-word redraw; returns the number
   of words consumed or -1 when the words run out */
long long gaussians_from_raw_host(const unsigned int *raw, long long n_raw, long long want, double *out) {
    long long p = 0;
    auto upos = [&](double &u) -> bool {
        while (p < n_raw) {
            const unsigned int w = raw[p++];
            if (w != 0U) { u = (double)w / 4294967296.0; return true; }
        }
        return false;
    };
    for (long long g = 0; g < want; g++) {
        double x, y, r2;
        do {
            double u1, u2;
            if (!upos(u1) || !upos(u2)) return -1;
            x = -1 + 2 * u1;
            y = -1 + 2 * u2;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0);
        out[g] = y * std::sqrt(-2.0 * std::log(r2) / r2);
    }
    return p;
}

GslStream *gsl_stream_create(unsigned long mt_seed) {
    GslStream *s = new GslStream();
    /* gsl_rng_set for mt19937: seed 0 -> 4357, linear congruential fill (hostnum::Mt19937::seed) */
    unsigned int st[MT_N];
    unsigned long sd = mt_seed ? mt_seed : 4357;
    st[0] = (unsigned int)(sd & 0xffffffffUL);
    for (int i = 1; i < MT_N; i++) st[i] = 1812433253U * (st[i - 1] ^ (st[i - 1] >> 30)) + (unsigned int)i;
    s->d_state = (unsigned int *)dev_alloc(sizeof(unsigned int) * MT_N);
    s->d_info = (long long *)dev_alloc(sizeof(long long) * 2);
    h2d(s->d_state, st, sizeof(st));
    dev_sync();
    g_stats.h2d -= (long long)sizeof(st);
    return s;
}
void gsl_stream_destroy(GslStream *s) {
    if (!s) return;
    dev_sync();
    dev_free(s->d_state); dev_free(s->d_raw); dev_free(s->d_counts); dev_free(s->d_info);
    delete s;
}

/* the next `want` Gaussians of the stream into d_out (device) */
void gsl_stream_gaussians(GslStream *s, double *d_out, long long want) {
    if (want <= 0) return;
    /* attempts needed: want / (pi/4) on average, sigma = sqrt(want (1-p)) / p; 12 sigma + slack */
    const double p_acc = 0.78539816339744830962;
    const long long attempts = (long long)((double)want / p_acc + 12.0 * std::sqrt((double)want * (1 - p_acc)) / p_acc) + 2048;
    const size_t need = (size_t)(2 * attempts);
    size_t have = s->n_raw - s->pos;
    if (have < need) {
        /* refill: unconsumed tail to the front, then whole 624-word refreshes up to `need` */
        const size_t cap_needed = need + MT_N;
        if (cap_needed > s->cap_raw) {
            unsigned int *nb = (unsigned int *)dev_alloc(sizeof(unsigned int) * cap_needed);
            if (have) d2d(nb, s->d_raw + s->pos, sizeof(unsigned int) * have);
            dev_sync();
            dev_free(s->d_raw);
            s->d_raw = nb; s->cap_raw = cap_needed;
        } else if (s->pos > 0 && have > 0) {
            if (have <= s->pos) {
                d2d(s->d_raw, s->d_raw + s->pos, sizeof(unsigned int) * have); /* disjoint ranges */
            } else {
                DevBuf<unsigned int> tmp(have);
                d2d(tmp, s->d_raw + s->pos, sizeof(unsigned int) * have);
                d2d(s->d_raw, tmp, sizeof(unsigned int) * have);
                dev_sync();
            }
        }
        s->pos = 0; s->n_raw = have;
        const size_t blocks = (need - have + MT_N - 1) / MT_N;
        MtGenArgs ga = {s->d_state, s->d_raw + s->n_raw, (long long)blocks};
        B200_LAUNCH(mt19937_generate_kernel, 1, 256, 0, ga);
        s->n_raw += blocks * MT_N;
        have = s->n_raw - s->pos;
    }
    const int per_block = 256 * GA_PER_THREAD;
    const long long nblk = (attempts + per_block - 1) / per_block;
    if ((size_t)nblk > s->cap_counts) {
        dev_free(s->d_counts);
        s->d_counts = (int *)dev_alloc(sizeof(int) * (size_t)nblk);
        s->cap_counts = (size_t)nblk;
    }
    long long info[2] = {-1, 0};
    h2d(s->d_info, info, sizeof(info));
    g_stats.h2d -= (long long)sizeof(info);
    GaussArgs a = {s->d_raw + s->pos, attempts, want, s->d_counts, s->d_info, d_out, per_block};
    const int grid = (int)(nblk < (long long)dev_num_sms() * 8 ? nblk : (long long)dev_num_sms() * 8);
    B200_LAUNCH(gauss_count_kernel, grid, 256, 0, a);
    DevBuf<long long> total(1);
    ScanArgs sa = {s->d_counts, nblk, total};
    B200_LAUNCH(gauss_scan_kernel, 1, 256, 0, sa);
    B200_LAUNCH(gauss_emit_kernel, grid, 256, 0, a);
    d2h(info, s->d_info, sizeof(info));
    g_stats.d2h -= (long long)sizeof(info);
    if (info[1] != 0) {
        /* a zero word in the chunk: GSL redraws it, which shifts the pairing -> sequential rule */
        std::vector<unsigned int> raw(need);
        std::vector<double> out((size_t)want);
        d2h(raw.data(), s->d_raw + s->pos, sizeof(unsigned int) * need);
        const long long used = gaussians_from_raw_host(raw.data(), (long long)need, want, out.data());
        if (used < 0) b200_throw(B200_TableGenerationError, "random stream: not enough words for the request");
        h2d(d_out, out.data(), sizeof(double) * (size_t)want);
        dev_sync();
        g_stats.h2d -= (long long)(sizeof(double) * (size_t)want);
        g_stats.d2h -= (long long)(sizeof(unsigned int) * need);
        s->pos += (size_t)used;
    } else {
        if (info[0] < 0) b200_throw(B200_TableGenerationError, "random stream: acceptance margin exhausted");
        s->pos += (size_t)(2 * info[0]);
    }
    s->total_gaussians += want;
}

/* ------------------------------------------------------------------ test hooks (C ABI) */
/* n1 then n2 Gaussians of the mt19937 stream seeded with `mt_seed`, through the device pipeline */
extern "C" int b200_gsl_gaussian_stream(unsigned long mt_seed, long long n1, long long n2, double *host_out) {
    try {
        rt_init();
        GslStream *s = gsl_stream_create(mt_seed);
        DevBuf<double> d((size_t)(n1 + n2 > 0 ? n1 + n2 : 1));
        gsl_stream_gaussians(s, d, n1);
        gsl_stream_gaussians(s, d.p + n1, n2);
        d2h(host_out, d, sizeof(double) * (size_t)(n1 + n2));
        gsl_stream_destroy(s);
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_gsl_gaussian_stream: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
/* the same stream from the sequential host generator (hostnum::Mt19937::ugaussian) */
extern "C" int b200_host_gaussian_stream(unsigned long mt_seed, long long n, double *host_out) {
    hostnum::Mt19937 rng(mt_seed);
    for (long long i = 0; i < n; i++) host_out[i] = rng.ugaussian();
    return 0;
}
extern "C" long long b200_gaussians_from_raw_host(const unsigned int *raw, long long n_raw, long long want, double *out) {
    return gaussians_from_raw_host(raw, n_raw, want, out);
}

/* test hook: the x-plane range of thread t under the reference's static OpenMP schedule */
extern "C" void b200_omp_static_range(int n, int n_threads, int t, int *begin, int *end) {
    hostnum::omp_static_range(n, n_threads, t, begin, end);
}
