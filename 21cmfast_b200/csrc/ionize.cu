/*
 * ionize.cu -- ComputeIonizedBox: excursion-set ionisation field (reference IonisationBox.c).
 *
 * Scope: SOURCE_MODEL in {CONST-ION-EFF, E-INTEGRAL}, USE_TS_FLUCT off,
 * RECOMB_MODEL none, no mini-halos, centre-cell flagging.  Other branches return ValueError.
 *
 * Data flow per call (all grids stay in HBM; the k-space density is transformed once and stays
 * resident across the whole radius ladder):
 *
 *   density (host) --H2D--> r2c with clip / (1/N) fused          [prepare_box_for_filtering]
 *   for R from largest to smallest:
 *     window table for this R (exact double formula per |n|^2, expanded to coalesced rows)
 *     c2r with W(kR) multiplied on load, clip + global min/max in the epilogue
 *                                                   [copy_filter_transform + clip_and_get_extrema]
 *     2 floats D2H -> host builds the 400-point f_coll(delta) table -> 1.6 KB H2D
 *                                                   [setup_integration_tables]
 *     sweep 1: grid sum of f_coll(delta) as deterministic double block sums (the f_coll grid itself
 *              is only materialised for the last radius, where it is the unnormalised_nion output)
 *                                                               [calculate_fcoll_grid]
 *     sweep 2: mean fix + barrier test re-evaluated from the filtered density -> ionised byte mask;
 *              last radius: flags from the f_coll grid + partial ionisation of the other cells
 *                                                               [find_ionised_regions]
 *   xH = 0 / z_reion from the mask, temperatures of ionised cells, D2H of the outputs
 *                                                               [set_ionized_temperatures]
 *
 * The radius loop is software-pipelined: while the host waits for the two min/max keys of radius
 * k and integrates its table, the stream already filters and transforms the next radius (two at
 * 256^3) into other work boxes, so the GPU does not idle on the host-side table build.
 * Multi-GPU: the ladder can be partitioned by radius (IonPartition below).
 */
#include "fft.h"
#include "dist.h"
#include "host_physics.h"
#include "host_recomb.h"

#include <omp.h>
#include <algorithm>
#include <type_traits>
#include <thread>
#include <vector>

#define HII_ROUND_ERR (1e-5)

/* ------------------------------------------------------------------ device side */
struct DevTable {
    double x_min, x_width, inv_width;
    int log_valued;
    float y[N_DENS_INTERP];
};

/* partial sums a CTA of the sum sweep writes per chunk: one per warp (256-thread CTAs), one in the host emulation */
#ifndef B200_EMU
#define SWEEP_PARTIALS 8
#else
#define SWEEP_PARTIALS 1
#endif

struct SweepArgs {
    int nx, ny, nz, nzc;
    const float *filtered;   /* padded real rows, already clipped to [-1, 1e6] */
    const DevTable *table;
    double *partial;         /* [ceil(nx ny / chunk_rows)][SWEEP_PARTIALS] sums of consecutive row chunks, per warp */
    float *fcoll;            /* unpadded f_coll grid of this radius, or null (sum only) */
    int chunk_rows;          /* rows per chunk: divides ny, so that chunks never straddle an x plane */
    /* speculative flags (fcoll_sum_kernel<LOG, true>), see SpecState */
    struct SpecState *spec;
    int j;                   /* index of this radius among the radii of the call */
    double ion_eff_factor;
    unsigned char *mask;
    uint2 *queue;            /* gridDim.x private segments of qcap (cell index, density bits) pairs, one per CTA */
    unsigned int qcap;
    unsigned int *qcounts;   /* [gridDim.x] cells each CTA wanted to queue (> qcap: its segment overflowed) */
};

/* ONE sweep per radius instead of two.  The ionised flag of a cell is mean_fix f_coll(delta) zeta > 1, and
   mean_fix = <f_coll> / (grid mean of f_coll) is only known after the whole grid has been summed -- which is why
   the reference (and ionise_delta_kernel) walk the filtered grid a second time.  The mean fix, however,
   drifts by well under a per cent from one radius of the ladder to the next.  The sum sweep therefore
   classifies every cell against a BRACKET of the mean fix predicted from the two previous radii
   (geometric extrapolation, +- SPEC_EPS or 0.3 SPEC_EPS, see spec_bracket): cells that are ionised for every mean fix inside the bracket
   are flagged at once, cells that are neutral for every mean fix inside it are skipped, and the few cells
   in between are queued (4-byte cell indices; every CTA fills a private segment of the queue behind a
   shared-memory counter -- one global counter serialised the sweep: 0.95 ms per radius against 0.19).  Once
   the grid sum exists, spec_resolve_kernel checks that the
   true mean fix lies inside the bracket and decides the queued cells with the reference arithmetic.  If the
   prediction ever fails, flags may already be wrong: the kernel raises `failed` and the host re-runs the
   ladder of this call without speculation (never observed; the check is what makes the shortcut exact).
   A queue segment that overflows (it holds a sixteenth of the CTA's cells; ~1.3 % are queued) does the same. */
#define SPEC_EPS 0.02
#define MAX_RADII 256 /* filter radii of one ladder: DELTA_R_HII_FACTOR = 1.03 between 0.62 and 50 Mpc gives 149 */
#define SPEC_MAX_RADII MAX_RADII
struct SpecState {
    double eps;                      /* half width of the bracket (SPEC_EPS; B200_SPEC_EPS overrides it for tests) */
    double mean_fix[SPEC_MAX_RADII]; /* true mean fix of every radius processed so far in this call */
    unsigned int qcount[SPEC_MAX_RADII];   /* queued cells of the radius, all segments */
    unsigned int overflow[SPEC_MAX_RADII]; /* a segment of the radius was too small: full flag sweep instead */
    int failed;
};
struct SpecBracket {
    double gain_lo, gain_hi; /* mean_fix zeta at the ends of the bracket */
};
DEV SpecBracket spec_bracket(const SpecState *sp, int j, double ion_eff_factor) {
    /* geometric extrapolation of the mean fix from the previous radii: first order (ratio of the last two) for the
       third radius of a call, second order (the ratio's own ratio) from the fourth on -- the mean fix is a smooth
       function of the radius, and the second-order prediction misses it by 0.08-0.14 % where the first-order one
       misses by 0.55 % (512^3 and 32^3 ladders, both source models), so its bracket is 0.3 eps wide and the queue
       a third as long */
    const double p1 = sp->mean_fix[j - 1], p2 = sp->mean_fix[j - 2];
    double ratio = p1 / p2, half = sp->eps;
    if (!(ratio > 0.5 && ratio < 2.0)) ratio = 1.0;
    if (j >= 3) {
        const double r2 = p2 / sp->mean_fix[j - 3];
        const double accel = ratio / r2;
        if (r2 > 0.5 && r2 < 2.0 && accel > 0.9 && accel < 1.1) { ratio *= accel; half *= 0.3; }
    }
    const double pred = p1 * ratio;
    SpecBracket b;
    b.gain_lo = pred * (1.0 - half) * ion_eff_factor;
    b.gain_hi = pred * (1.0 + half) * ion_eff_factor;
    return b;
}

/* The grid sum is reduced in a FIXED tree that depends on the box shape only -- chunk sums (a CTA's
   deterministic reduction over chunk_rows consecutive rows), then per-x-plane sums of the chunk sums in
   order (plane_sum_kernel), then the fixed-order sum over the planes that every CTA of the flag sweep
   re-adds -- so the mean fix is bit-identical for any grid size and for any split of the box into
   x-slabs over GPUs. */

/* Shared-memory copy of the radius' table in the forms the sweeps use.
 *
 * The sweeps are HBM-bound only if a cell costs ~25 single-precision instructions: on B200 a cell
 * of a 4-byte stream has a budget of ~0.2 SM cycles, i.e. ~13 double-precision or ~3 64-bit
 * conversion instructions.  The reference's arithmetic (double interpolation of a float table,
 * interpolation.c:123-131, double exp) is therefore kept for what is stored or decided, and a
 * single-precision evaluation is used where its error cannot reach the result:
 *   - grid SUM of radii whose f_coll grid is not an output: per-cell relative error ~1e-7, unbiased,
 *     so the mean over >= 32^3 cells moves by < 1e-9 relative;
 *   - ionised FLAG: decided in single precision only when the value is further from the threshold
 *     than a band that covers the single-precision evaluation error (1e-5 + 1e-4 |dy| of the bin);
 *     inside the band the reference arithmetic decides.
 * Log-valued (E-INTEGRAL) table: exp(y0 (1-t) + y1 t) = exp(y0) exp(t dy) with exp(y0) tabulated
 * per bin and exp(u), |u| <= 1/4, a degree-6 Taylor series (remainder < 1.3e-8, typical |u| << 1/4);
 * steeper bins take the reference path. */
struct SweepTable {
    float y[N_DENS_INTERP];        /* the table as uploaded (reference path) */
    float2 *rep;                   /* dynamic shared memory: the bin table the fast path reads, SWEEP_REP copies
                                      interleaved as rep[idx * SWEEP_REP + (lane & (SWEEP_REP - 1))].
                                      Entry = {y0 or exp(y0), y1 - y0}, see sweep_table_load(.., as_exp) */
    double x_min, x_width, inv_width;
    float x_min_f, inv_width_f;
    int log_valued;
};
#define SWEEP_REP 1 /* copies of the bin table; 16 interleaved copies make the loads bank-conflict-free but
                       measured slower on B200 (the sweeps are not shared-memory bound) */
#define SWEEP_REP_BYTES (N_DENS_INTERP * SWEEP_REP * sizeof(float2) + 2 * 4 * 256 * sizeof(float4)) /* table + load ring */
DEV int sweep_rep_lane() { return threadIdx.x & (SWEEP_REP - 1); }
/* as_exp: first component exp(y0) (grid sum of a log-valued table) instead of y0 */
DEV void sweep_table_load(SweepTable *st, const DevTable *t, float2 *rep, bool as_exp) {
    for (int i = threadIdx.x; i < N_DENS_INTERP; i += blockDim.x) {
        const float y0 = t->y[i];
        const float y1 = t->y[i + 1 < N_DENS_INTERP ? i + 1 : i];
        const float dy = (float)((double)y1 - (double)y0);
        st->y[i] = y0;
        const float2 e = make_float2(as_exp ? (float)exp((double)y0) : y0, dy);
#pragma unroll
        for (int r = 0; r < SWEEP_REP; r++) rep[i * SWEEP_REP + r] = e;
    }
    if (threadIdx.x == 0) {
        st->rep = rep;
        st->x_min = t->x_min; st->x_width = t->x_width; st->inv_width = t->inv_width;
        st->x_min_f = (float)t->x_min; st->inv_width_f = (float)t->inv_width;
        st->log_valued = t->log_valued;
    }
}
/* reference arithmetic: EvaluateRGTable1D_f (interpolation.c:123-131), exp for log-valued tables */
#ifndef B200_EMU
__device__ __noinline__
#else
inline
#endif
double fcoll_exact(float d, const SweepTable *h) {
    const double x = (double)d;
    const int idx = (int)floor((x - h->x_min) * h->inv_width);
    const double table_val = h->x_min + h->x_width * (float)idx;
    const double t = (x - table_val) * h->inv_width;
    const double v = (double)h->y[idx] * (1 - t) + (double)h->y[idx + 1] * t;
    return h->log_valued ? exp(v) : v;
}
/* single-precision bin coordinates of TWO cells at a time (packed arithmetic).  pos is clamped into
   the table so that a rounding slip at the table ends cannot index outside it. */
struct SweepConstsF {
    float inv, c0, floor;   /* pos = d * inv + c0,  c0 = -x_min * inv */
    float pos_lo, pos_hi;   /* clamp of pos: the density floor and the table ends in one min/max pair */
};
DEV SweepConstsF sweep_consts(const SweepTable *h, float dens_floor) {
    SweepConstsF k;
    k.inv = h->inv_width_f;
    k.c0 = -h->x_min_f * h->inv_width_f;
    k.floor = dens_floor;
    k.pos_lo = fmaxf(fmaf(dens_floor, k.inv, k.c0), 0.f);
    k.pos_hi = (float)(N_DENS_INTERP - 1) - 1e-3f;
    return k;
}
DEV void table_coords_f2(float2 d, const SweepConstsF &k, int &i0, int &i1, float2 &t) {
    float2 pos = f2_fma(d, make_float2(k.inv, k.inv), make_float2(k.c0, k.c0));
    pos.x = fminf(fmaxf(pos.x, k.pos_lo), k.pos_hi);
    pos.y = fminf(fmaxf(pos.y, k.pos_lo), k.pos_hi);
    i0 = float_to_int_floor(pos.x);
    i1 = float_to_int_floor(pos.y);
    t = f2_add(pos, make_float2(-(float)i0, -(float)i1));
}
DEV float2 exp_small_f2(float2 u) { /* Taylor series of exp, |u| <= 1/4, two lanes */
    float2 p = make_float2(1.0f / 720.0f, 1.0f / 720.0f);
    p = f2_fma(p, u, make_float2(1.0f / 120.0f, 1.0f / 120.0f));
    p = f2_fma(p, u, make_float2(1.0f / 24.0f, 1.0f / 24.0f));
    p = f2_fma(p, u, make_float2(1.0f / 6.0f, 1.0f / 6.0f));
    p = f2_fma(p, u, make_float2(0.5f, 0.5f));
    p = f2_fma(p, u, make_float2(1.0f, 1.0f));
    return f2_fma(p, u, make_float2(1.0f, 1.0f));
}
/* f_coll of two cells for the grid sum only: straight-line code; `steep` is raised when a lane's
   |t dy| leaves the range of the series and the caller must redo the chunk with fcoll_exact */
template <bool LOG> DEV float2 fcoll_fast2(float2 d, const float2 *rep_base, const SweepConstsF &k, bool &steep) {
    /* rep_base is the kernel's own dynamic shared-memory array (not the generic pointer kept in SweepTable): the
       lookups compile to LDS with 32-bit addresses instead of generic LD + 64-bit address arithmetic */
    int i0, i1;
    float2 t;
    table_coords_f2(d, k, i0, i1, t);
    const float2 *rep = rep_base + sweep_rep_lane();
    const float2 e0 = rep[i0 * SWEEP_REP], e1 = rep[i1 * SWEEP_REP];
    const float2 u = f2_mul(t, make_float2(e0.y, e1.y));
    if (!LOG) return f2_add(make_float2(e0.x, e1.x), u);
    steep = steep || (fmaxf(fabsf(u.x), fabsf(u.y)) > 0.25f);
    return f2_mul(make_float2(e0.x, e1.x), exp_small_f2(u));
}

/* Visit every float4 chunk of the padded real box [nrows][2 pitch] (nz valid floats per row, nz a
   multiple of 4): f(d4, row, zc) with zc the chunk index inside the row.  When a 256-thread CTA
   covers a whole number of rows (256 % (nz/4) == 0, i.e. nz a power of two <= 1024) a thread keeps
   its (row offset, zc) for the whole kernel and an iteration is four independent 128-bit loads
   at constant row strides: no per-chunk index arithmetic. */
#define SWEEP_RING_BYTES (2 * 4 * 256 * sizeof(float4)) /* two iterations of four chunks per thread */
template <class R, class F1, class F2>
DEV void for_each_chunk(const float *filtered, long long nrows, int nz, int pitch, float4 *ring, F1 &&fast, F2 &&finish) {
    /* fast(d4) -> R must be straight-line code (the four chunks of an iteration are evaluated
       back to back so that their dependency chains interleave); finish(R, d4, row, zc) holds the
       rare branches and the side effects */
    const int q = nz >> 2;
    const long long rstride = 2LL * pitch; /* floats per padded row */
#ifndef B200_EMU
    if (blockDim.x == 256 && q <= 256 && (256 % q) == 0) {
        /* A thread keeps its (row offset, zc) for the whole kernel.  Its loads travel as cp.async
           copies into a private slot of a shared-memory ring, one iteration ahead of the
           arithmetic: the bytes in flight per SM stay at 16 KB per resident CTA all the time
           without holding registers (ncu: the register-staged version stalled on the loads,
           long_scoreboard ~8 of ~12 cycles per issue).  A thread only ever reads the slots it
           filled itself, so no barrier is involved. */
        const int rows_per_step = 256 / q;
        const int r = threadIdx.x / q, zc = threadIdx.x - r * q;
        const long long step_rows = 4LL * rows_per_step;
        const long long gstride = (long long)gridDim.x * step_rows;
        auto issue = [&](long long row0, int slot) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                long long row = row0 + r + (long long)u * rows_per_step;
                if (row >= nrows) row = nrows - 1; /* harmless duplicate, discarded below */
                cp_async_16(&ring[(slot * 4 + u) * 256 + threadIdx.x], filtered + row * rstride + 4 * zc);
            }
            cp_async_commit();
        };
        long long row0 = (long long)blockIdx.x * step_rows;
        if (row0 < nrows) issue(row0, 0);
        for (int it = 0; row0 < nrows; row0 += gstride, it++) {
            if (row0 + gstride < nrows) issue(row0 + gstride, (it + 1) & 1);
            else cp_async_commit(); /* keep one group per iteration so that wait_group 1 is uniform */
            cp_async_wait_group<1>();
            float4 d4[4];
            R res[4];
#pragma unroll
            for (int u = 0; u < 4; u++) d4[u] = ring[((it & 1) * 4 + u) * 256 + threadIdx.x];
#pragma unroll
            for (int u = 0; u < 4; u++) res[u] = fast(d4[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long row = row0 + r + (long long)u * rows_per_step;
                if (row < nrows) finish(res[u], d4[u], row, zc);
            }
        }
        cp_async_wait_all();
    } else {
        const long long nchunks = nrows * q;
        for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < nchunks;
             id += (long long)gridDim.x * blockDim.x) {
            const long long row = id / q;
            const int zc = (int)(id - row * q);
            const float4 d = *reinterpret_cast<const float4 *>(filtered + row * rstride + 4 * zc);
            finish(fast(d), d, row, zc);
        }
    }
#else
    (void)ring;
    {
        const long long nchunks = nrows * q;
        for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < nchunks;
             id += (long long)gridDim.x * blockDim.x) {
            const long long row = id / q;
            const int zc = (int)(id - row * q);
            const float4 d = *reinterpret_cast<const float4 *>(filtered + row * rstride + 4 * zc);
            finish(fast(d), d, row, zc);
        }
    }
#endif
}

/* The same visit order restricted to chunks of `CH` consecutive rows: CTA b owns chunks b, b + grid, ...
   and end_chunk(chunk) is called by every thread of the CTA after the last cell of a chunk (it may hold
   barriers).  The cp.async ring keeps running across chunk boundaries. */
template <class R, class F1, class F2, class F3>
DEV void for_each_chunk_blocked(const float *filtered, long long nrows, int nz, int pitch, int CH, float4 *ring, F1 &&fast,
                                F2 &&finish, F3 &&end_chunk) {
    const int q = nz >> 2;
    const long long rstride = 2LL * pitch;
    const long long nchunks = (nrows + CH - 1) / CH;
#ifndef B200_EMU
    if (blockDim.x == 256 && q <= 256 && (256 % q) == 0 && (CH % (4 * (256 / q))) == 0 && (nrows % CH) == 0) {
        /* the chunks tile the rows exactly: no row of an iteration lies outside the grid */
        const int rows_per_step = 256 / q;
        const int r = threadIdx.x / q, zc = threadIdx.x - r * q;
        const int step_rows = 4 * rows_per_step;
        const int ipc = CH / step_rows; /* iterations per chunk */
        const float *const mine = filtered + (long long)r * rstride + 4 * zc; /* this thread's column of the first step */
        const long long ustride = (long long)rows_per_step * rstride;
        auto issue = [&](long long row0, int slot) {
            const float *p = mine + row0 * rstride;
#pragma unroll
            for (int u = 0; u < 4; u++) cp_async_16(&ring[(slot * 4 + u) * 256 + threadIdx.x], p + u * ustride);
            cp_async_commit();
        };
        long long chunk = blockIdx.x, next_chunk = blockIdx.x;
        int it_in = 0, next_it = 0; /* position of the current / the prefetched iteration inside its chunk */
        auto advance = [&](long long &c, int &i) { if (++i == ipc) { i = 0; c += gridDim.x; } };
        if (chunk < nchunks) { issue(chunk * CH, 0); advance(next_chunk, next_it); }
        for (int it = 0; chunk < nchunks; it++) {
            if (next_chunk < nchunks) issue(next_chunk * CH + (long long)next_it * step_rows, (it + 1) & 1);
            else cp_async_commit(); /* keep one group per iteration so that wait_group 1 is uniform */
            advance(next_chunk, next_it);
            cp_async_wait_group<1>();
            const long long row0 = chunk * CH + (long long)it_in * step_rows;
            float4 d4[4];
            R res[4];
#pragma unroll
            for (int u = 0; u < 4; u++) d4[u] = ring[((it & 1) * 4 + u) * 256 + threadIdx.x];
#pragma unroll
            for (int u = 0; u < 4; u++) res[u] = fast(d4[u]);
            finish(res, d4, row0 + r, rows_per_step, zc, std::integral_constant<int, 4>()); /* rows row0 + r + u rows_per_step */
            const long long done = chunk;
            advance(chunk, it_in);
            if (it_in == 0) end_chunk(done);
        }
        cp_async_wait_all();
        return;
    }
#else
    (void)ring;
#endif
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long r0 = chunk * CH, r1 = (r0 + CH < nrows) ? r0 + CH : nrows;
        const long long nch = (r1 - r0) * q;
        for (long long id = threadIdx.x; id < nch; id += blockDim.x) {
            const long long row = r0 + id / q;
            const int zc = (int)(id - (row - r0) * q);
            const float4 d1[4] = {*reinterpret_cast<const float4 *>(filtered + row * rstride + 4 * zc)};
            R r1[4];
            r1[0] = fast(d1[0]);
            finish(r1, d1, row, 0, zc, std::integral_constant<int, 1>()); /* only entry 0 is live */
        }
        end_chunk(chunk);
    }
}

/* sweep 1: sum of f_coll over the grid as deterministic double block sums (calculate_fcoll_grid,
   IonisationBox.c:773-962).  With a.fcoll set (last radius: the grid is the unnormalised_nion
   output) every cell takes the reference arithmetic and the float grid is written. */
template <bool LOG, bool SPEC> __global__ void __launch_bounds__(256) fcoll_sum_kernel(SweepArgs a) {
    __shared__ SweepTable st;
    __shared__ unsigned int q_n;
    DYN_SMEM(float2, rep);
    sweep_table_load(&st, a.table, rep, LOG);
    __shared__ int s_first_hi, s_last_lo;
    __shared__ float s_d_lo, s_d_hi;
    if (SPEC && threadIdx.x == 0) { q_n = 0; s_first_hi = N_DENS_INTERP; s_last_lo = -1; }
    __syncthreads();
    /* The flag test of the reference, (float) f_coll(delta) mean_fix zeta > 1, follows the table, which grows
       with the density (up to a few unordered nodes, below), so the bracket of the mean fix becomes a bracket
       [d_lo, d_hi] of the filtered density itself: ionised for every mean fix inside the bracket above d_hi,
       neutral for every one below d_lo.  Two compares per cell instead of a second table evaluation.  The
       crossings are solved on the interpolant of the reference path (fcoll_exact) with a margin of 1e-5 in
       (log) f_coll -- two orders above its rounding -- and rounded outwards. */
    float d_lo = 0.f, d_hi = 0.f;
    if (SPEC) {
        const SpecBracket br = spec_bracket(a.spec, a.j, a.ion_eff_factor);
        const double v_hi = LOG ? -log(br.gain_lo) + 1e-5 : (1.0 / br.gain_lo) * (1.0 + 1e-5);
        const double v_lo = LOG ? -log(br.gain_hi) - 1e-5 : (1.0 / br.gain_hi) * (1.0 - 1e-5);
        /* the table need not be ordered (the reference's integrals switch method at delta = 1.2 and collapse to
           one halo of the condition mass near delta_crit): d_hi lies in the bin after the LAST node below v_hi --
           every node above it, hence the interpolant, is >= v_hi -- and d_lo in the bin before the FIRST node
           above v_lo; what is unordered in between simply falls inside [d_lo, d_hi] and is queued */
        for (int i = threadIdx.x; i < N_DENS_INTERP; i += blockDim.x) {
            const double y = (double)st.y[i];
            if (y < v_hi) atomic_max_i32(&s_last_lo, i);   /* reused: last node below v_hi */
            if (y > v_lo) atomic_min_i32(&s_first_hi, i);  /* reused: first node above v_lo */
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const float inf = 3.0e38f;
            float lo, hi;
            const int a_hi = s_last_lo, b_lo = s_first_hi;
            if (a_hi < 0) hi = -inf;                          /* every node ionises for sure */
            else if (a_hi >= N_DENS_INTERP - 1) hi = inf;     /* no cell can be ionised for sure */
            else {
                const double y0 = (double)st.y[a_hi], y1 = (double)st.y[a_hi + 1];
                const double x = st.x_min + st.x_width * ((double)a_hi + (v_hi - y0) / (y1 - y0));
                hi = (float)x;
                hi += fabsf(hi) * 2.4e-7f + 1e-30f; /* outwards by two units in the last place */
            }
            if (b_lo >= N_DENS_INTERP) lo = inf;              /* every node is neutral for sure */
            else if (b_lo == 0) lo = -inf;
            else {
                const double y0 = (double)st.y[b_lo - 1], y1 = (double)st.y[b_lo];
                const double x = st.x_min + st.x_width * ((double)(b_lo - 1) + (v_lo - y0) / (y1 - y0));
                lo = (float)x;
                lo -= fabsf(lo) * 2.4e-7f + 1e-30f;
            }
            if (!(lo <= hi)) { lo = -inf; hi = inf; } /* cannot happen for v_lo < v_hi; then every cell goes to the queue */
            s_d_lo = lo; s_d_hi = hi;
        }
        __syncthreads();
        d_lo = s_d_lo; d_hi = s_d_hi;
    }
    /* four cells: bits 0..3 = ionised for sure, bits 4..7 = inside the bracket */
    auto classify4 = [&](const float4 &d) -> unsigned {
        const unsigned s = (d.x > d_hi ? 1u : 0u) | (d.y > d_hi ? 2u : 0u) | (d.z > d_hi ? 4u : 0u) | (d.w > d_hi ? 8u : 0u);
        const unsigned n = (d.x < d_lo ? 1u : 0u) | (d.y < d_lo ? 2u : 0u) | (d.z < d_lo ? 4u : 0u) | (d.w < d_lo ? 8u : 0u);
        return s | ((~(s | n) & 15u) << 4);
    };
    uint2 *const seg = SPEC ? a.queue + (size_t)blockIdx.x * a.qcap : nullptr;
    auto push_cells = [&](unsigned bits, unsigned int cell, const float4 &d) { /* bits 0..3: cells cell + i, with their densities */
        const unsigned n = (unsigned)__builtin_popcount(bits);
        const unsigned at = atomic_fetch_add_u32(&q_n, n);
        const float dv[4] = {d.x, d.y, d.z, d.w};
        unsigned k = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (bits & (1u << i)) {
                if (at + k < a.qcap) seg[at + k] = make_uint2(cell + i, (unsigned int)float_as_int_bits(dv[i]));
                k++;
            }
    };
    /* flags of four cells = one 32-bit OR into the byte mask (a fire-and-forget reduction: no branches, and the
       flags of earlier radii in the same word stay) */
    auto set_sure = [&](unsigned m4, unsigned int cell) {
        atomic_or_u32(reinterpret_cast<unsigned int *>(a.mask + cell), (m4 * 0x00204081u) & 0x01010101u);
    };
    const long long nrows = (long long)a.nx * a.ny;
    const int CH = a.chunk_rows;
    const float dens_floor = (float)(-1. + pc::FRACT_FLOAT_ERR);
    const SweepConstsF kf = sweep_consts(&st, dens_floor);
    double acc = 0.;
    /* deterministic reduction of the chunk's thread sums: butterfly over the lanes of each warp (every lane ends
       with the same bits) -> partial[chunk][warp]; plane_sum_kernel adds the warps' entries in order.  No CTA
       barrier: the two __syncthreads per chunk of a CTA-wide reduction were 20 % of the sweep's stall cycles. */
    auto end_chunk = [&](long long chunk) {
#ifndef B200_EMU
        double v = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) a.partial[chunk * SWEEP_PARTIALS + (threadIdx.x >> 5)] = v;
#else
        a.partial[chunk * SWEEP_PARTIALS] = acc;
#endif
        acc = 0.;
    };
    if ((a.nz & 3) == 0 && !a.fcoll) {
        struct SumRes { float sum; bool steep; unsigned code; };
        for_each_chunk_blocked<SumRes>(
            a.filtered, nrows, a.nz, a.nzc, CH, reinterpret_cast<float4 *>(rep + N_DENS_INTERP * SWEEP_REP),
            [&](const float4 &d) -> SumRes {
                SumRes r;
                r.steep = false;
                r.code = 0;
                const float2 f = f2_add(fcoll_fast2<LOG>(make_float2(d.x, d.y), rep, kf, r.steep),
                                        fcoll_fast2<LOG>(make_float2(d.z, d.w), rep, kf, r.steep));
                r.sum = f.x + f.y;
                if (SPEC) r.code = classify4(d);
                return r;
            },
            [&](const SumRes (&r)[4], const float4 (&d)[4], long long row, int rows_per_step, int zc, auto live_c) {
                constexpr int live = decltype(live_c)::value;
                bool steep = false;
                unsigned code = 0;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (u < live) { steep = steep || r[u].steep; code |= r[u].code; }
                if (!steep) { /* one conversion and one double add per 16 cells */
                    float s4 = r[0].sum;
#pragma unroll
                    for (int u = 1; u < 4; u++)
                        if (u < live) s4 += r[u].sum;
                    acc += (double)s4;
                } else { /* a steep bin somewhere: reference arithmetic for the cells of the steep groups */
                    for (int u = 0; u < live; u++) {
                        if (r[u].steep)
                            acc += (fcoll_exact(fmaxf(d[u].x, dens_floor), &st) + fcoll_exact(fmaxf(d[u].y, dens_floor), &st)) +
                                   (fcoll_exact(fmaxf(d[u].z, dens_floor), &st) + fcoll_exact(fmaxf(d[u].w, dens_floor), &st));
                        else
                            acc += (double)r[u].sum;
                    }
                }
                if (SPEC && code) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (u < live && r[u].code) {
                            const unsigned int cell = (unsigned int)(row + (long long)u * rows_per_step) * (unsigned int)a.nz +
                                                      4u * (unsigned int)zc; /* NL < 2^32 */
                            if (r[u].code & 15u) set_sure(r[u].code & 15u, cell);
                            if (r[u].code & 0xf0u) push_cells(r[u].code >> 4, cell, d[u]);
                        }
                    }
                }
            },
            end_chunk);
    } else {
        const long long nchunks = (nrows + CH - 1) / CH;
        for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
            const long long r0 = chunk * CH, r1 = (r0 + CH < nrows) ? r0 + CH : nrows;
            for (long long row = r0; row < r1; row++) {
                const float *src = a.filtered + row * 2 * a.nzc;
                for (int z = threadIdx.x; z < a.nz; z += blockDim.x) {
                    if (a.fcoll) {
                        const double f = fcoll_exact(fmaxf(src[z], dens_floor), &st);
                        acc += f;
                        a.fcoll[row * a.nz + z] = (float)f;
                    } else {
                        bool steep = false;
                        const float f = fcoll_fast2<LOG>(make_float2(src[z], src[z]), rep, kf, steep).x;
                        acc += steep ? fcoll_exact(fmaxf(src[z], dens_floor), &st) : (double)f;
                    }
                }
            }
            end_chunk(chunk);
        }
    }
    if (SPEC) {
        __syncthreads();
        if (threadIdx.x == 0) a.qcounts[blockIdx.x] = q_n;
    }
}

/* second level of the fixed reduction tree: plane[x] = sum of the chunk sums of x-plane x, in order */
struct PlaneSumArgs {
    int nx, chunks_per_plane;
    const double *partial;
    double *plane;
};
__global__ void plane_sum_kernel(PlaneSumArgs a) {
#ifndef B200_EMU
    /* one warp per plane: lane l adds the entries l, l + 32, ... in order, then a butterfly over the lanes -- a fixed
       tree of the plane's entries, the same on every grid size and slab split */
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int x = warp; x < a.nx; x += nwarps) {
        const double *p = a.partial + (long long)x * a.chunks_per_plane;
        double acc = 0.;
        for (int c = lane; c < a.chunks_per_plane; c += 32) acc += p[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) a.plane[x] = acc;
    }
#else
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < a.nx; x += gridDim.x * blockDim.x) {
        const double *p = a.partial + (long long)x * a.chunks_per_plane;
        double acc = 0.;
        for (int c = 0; c < a.chunks_per_plane; c++) acc += p[c];
        a.plane[x] = acc;
    }
#endif
}

struct CritArgs {
    long long n;
    const float *fcoll;      /* f_coll grid written by sweep 1 */
    const double *partial;   /* x-plane sums of sweep 1 (plane_sum_kernel), all planes of the box */
    int n_partial;
    const float *density;    /* unfiltered perturbed density, unpadded */
    const float *prev_zre;   /* previous box z_reion or null (= all -1) */
    unsigned char *mask;     /* 1 = ionised at some radius so far */
    float *xH, *z_reion, *Tk;
    double n_cells, mean_f_coll, f_limit, ion_eff_factor;
    int mass_dep_zeta, R_index;
    double redshift, TK_nofluct, adia_TK_term, T_re;
    /* recombinations (RECOMB_MODEL != none): the barrier becomes 1 + N_rec / (1 + delta_R) and the
       first crossing records Gamma12 and the mean free path (IonisationBox.c:1084-1140) */
    int recomb;              /* 0 none, 1 filtered N_rec grid (padded rows), 2 per-cell previous N_rec, 3 one global value */
    const float *rec_grid;
    double rec_scalar;
    const float *filtered;   /* padded real rows of delta_R (curr_dens for R_index > 0) */
    int nz, nzc;
    float *G12, *mfp;
    double R, gamma_prefactor;
    /* spin-temperature inputs (USE_TS_FLUCT): the residual electron fraction x_e, filtered like the
       density, lowers the barrier to (1 - x_e)(1 + rec) and is taken off the partial ionisations; the
       neutral gas has the TsBox's temperature (IonisationBox.c:1100-1107,1165-1187) */
    const float *xe_grid;    /* padded real rows of x_e at this radius, or null */
    const float *Tk_neutral; /* unpadded, or null */
    /* IONISE_ENTIRE_SPHERE: cells already painted by a sphere of an earlier radius (their x_HI is 0:
       no partial ionisation), or null */
    const unsigned char *paint;
    /* Lagrangian source grids (SOURCE_MODEL = L-INTEGRAL): the filtered photon-output grid of the HaloBox replaces
       f_coll, divided by the local baryon density (IonisationBox.c:1054-1066); sfr_grid feeds Gamma12 (:1126-1132) */
    const float *stars_grid; /* padded real rows, or null */
    const float *sfr_grid;   /* padded real rows, or null */
    double rho_baryon;       /* RHOcrit * OMb */
};

DEV float partially_ionized_temperature(float T_HI, float res_xH, float T_re) { /* thermochem.c:58-63 */
    if (res_xH <= 0.) return T_re;
    if (res_xH >= 1) return T_HI;
    return T_HI * res_xH + T_re * (1. - res_xH);
}

/* one cell of find_ionised_regions (IonisationBox.c:1040-1196), centre-cell method.  Radii above
   the last one only record "ionised" in a byte mask (the reference rewrites xH = 0 and z_reion at
   every radius that ionises the cell; both are functions of the mask alone and are materialised
   once by finalize_kernel).  The last radius also assigns the partial ionisations. */
DEV void ionise_cell(const CritArgs &a, long long idx, float fcoll, double mean_fix) {
    double curr_fcoll = mean_fix * (double)fcoll;
    if (a.mass_dep_zeta && curr_fcoll < a.f_limit) curr_fcoll = a.f_limit;
    if (curr_fcoll * a.ion_eff_factor > 1.0) {
        a.mask[idx] = 1;
    } else if (a.R_index == 0 && !a.mask[idx] && !(a.paint && a.paint[idx]) && (a.xH[idx] > pc::TINY)) {
        double res_xH = 1. - curr_fcoll * a.ion_eff_factor;
        if (a.Tk) {
            const float T_HI = (float)(a.TK_nofluct * (1 + a.adia_TK_term * a.density[idx]));
            a.Tk[idx] = partially_ionized_temperature(T_HI, (float)res_xH, (float)a.T_re);
        }
        if (res_xH < 0) res_xH = 0;
        else if (res_xH > 1) res_xH = 1;
        a.xH[idx] = (float)res_xH;
    }
}

/* the same cell with recombinations and / or x_e in the barrier; runs at every radius of the ladder */
DEV void ionise_cell_recomb(const CritArgs &a, long long idx, float fcoll, double mean_fix) {
    double curr_fcoll = mean_fix * (double)fcoll;
    if (a.mass_dep_zeta && curr_fcoll < a.f_limit) curr_fcoll = a.f_limit;
    const long long row = idx / a.nz;
    const long long idx_f = row * 2 * a.nzc + (idx - row * a.nz);
    const double curr_dens = a.R_index == 0 ? (double)a.density[idx]
                                            : (double)fmaxf(a.filtered[idx_f], (float)(-1. + pc::FRACT_FLOAT_ERR));
    double rec = a.recomb == 1 ? (double)a.rec_grid[idx_f] : a.recomb == 2 ? (double)a.rec_grid[idx] : a.rec_scalar;
    rec /= (1. + curr_dens);
    const double xe = a.xe_grid ? (double)a.xe_grid[idx_f] : 0.;
    if (curr_fcoll * a.ion_eff_factor > (1. - xe) * (1.0 + rec)) {
        if (a.recomb && !a.mask[idx] && (double)a.xH[idx] > pc::FRACT_FLOAT_ERR) { /* first (largest-R) crossing */
            a.G12[idx] = (float)(a.R * (a.gamma_prefactor * curr_fcoll));
            if (a.mfp) a.mfp[idx] = (float)a.R;
        }
        a.mask[idx] = 1;
    } else if (a.R_index == 0 && !a.mask[idx] && (a.xH[idx] > pc::TINY)) {
        double res_xH = 1. - curr_fcoll * a.ion_eff_factor;
        if (a.Tk) {
            const float T_HI = a.Tk_neutral ? a.Tk_neutral[idx] : (float)(a.TK_nofluct * (1 + a.adia_TK_term * a.density[idx]));
            a.Tk[idx] = partially_ionized_temperature(T_HI, (float)res_xH, (float)a.T_re);
        }
        res_xH -= xe;
        if (res_xH < 0) res_xH = 0;
        else if (res_xH > 1) res_xH = 1;
        a.xH[idx] = (float)res_xH;
    }
}

/* the same cell when the sources are a HaloBox grid: no mean fix, ion_eff_factor = 1, the photon output per
   baryon of the (filtered) cell against 1 + recombinations */
DEV void ionise_cell_lagrangian(const CritArgs &a, long long idx) {
    const long long row = idx / a.nz;
    const long long idx_f = row * 2 * a.nzc + (idx - row * a.nz);
    const double curr_dens = a.R_index == 0 ? (double)a.density[idx]
                                            : (double)fmaxf(a.filtered[idx_f], (float)(-1. + pc::FRACT_FLOAT_ERR));
    double curr_fcoll = (double)fmaxf(a.stars_grid[idx_f], 0.0f);
    curr_fcoll *= 1 / (a.rho_baryon * (1 + curr_dens));
    if (curr_fcoll < a.f_limit) curr_fcoll = a.f_limit;
    double rec = 0.;
    if (a.recomb) {
        rec = a.recomb == 1 ? (double)fmaxf(a.rec_grid[idx_f], 0.0f) : a.recomb == 2 ? (double)a.rec_grid[idx] : a.rec_scalar;
        rec /= (1. + curr_dens);
    }
    if (curr_fcoll > (1.0 + rec)) {
        if (a.recomb && !a.mask[idx] && (double)a.xH[idx] > pc::FRACT_FLOAT_ERR) { /* first (largest-R) crossing */
            a.G12[idx] = (float)(a.R * a.gamma_prefactor / (1 + curr_dens) * (double)fmaxf(a.sfr_grid[idx_f], 0.0f));
            if (a.mfp) a.mfp[idx] = (float)a.R;
        }
        a.mask[idx] = 1;
    } else if (a.R_index == 0 && !a.mask[idx] && (a.xH[idx] > pc::TINY)) {
        double res_xH = 1. - curr_fcoll;
        if (a.Tk) {
            const float T_HI = (float)(a.TK_nofluct * (1 + a.adia_TK_term * a.density[idx]));
            a.Tk[idx] = partially_ionized_temperature(T_HI, (float)res_xH, (float)a.T_re);
        }
        if (res_xH < 0) res_xH = 0;
        else if (res_xH > 1) res_xH = 1;
        a.xH[idx] = (float)res_xH;
    }
}
__global__ void __launch_bounds__(256) ionise_lagrangian_kernel(CritArgs a) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.n; idx += (long long)gridDim.x * blockDim.x)
        ionise_cell_lagrangian(a, idx);
}
/* deterministic block sums of a padded real grid clipped at zero (the grid mean the reference reports as
   mean_f_coll for Lagrangian sources, IonisationBox.c:1623-1628) */
struct RowSumArgs {
    long long nrows;
    int nz, nzc;
    const float *grid;
    double *partial; /* [gridDim.x] */
};
__global__ void __launch_bounds__(256) clipped_sum_kernel(RowSumArgs a) {
    __shared__ double red[256];
    double acc = 0.;
    for (long long row = blockIdx.x; row < a.nrows; row += gridDim.x)
        for (int z = threadIdx.x; z < a.nz; z += blockDim.x) acc += (double)fmaxf(a.grid[row * 2 * a.nzc + z], 0.0f);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partial[blockIdx.x] = red[0];
}

/* sweep 2: find_ionised_regions (IonisationBox.c:1008-1201) */
__global__ void __launch_bounds__(256) ionise_kernel(CritArgs a) {
    __shared__ double red[256];
    /* every CTA re-adds the block sums of sweep 1 in the same fixed order: the grid mean is
       bit-reproducible and no separate finishing launch is needed */
    double acc = 0.;
    for (int i = threadIdx.x; i < a.n_partial; i += blockDim.x) acc += a.partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    /* grid mean with the reference's floor (IonisationBox.c:1566-1576), then the mean fix */
    double grid_mean = red[0] / a.n_cells;
    if (a.mass_dep_zeta) {
        if (grid_mean <= a.f_limit) grid_mean = a.f_limit;
    } else {
        if (grid_mean <= pc::FRACT_FLOAT_ERR) grid_mean = pc::FRACT_FLOAT_ERR;
    }
    const double mean_fix = a.mean_f_coll / grid_mean;
    if (a.recomb || a.xe_grid) {
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.n;
             idx += (long long)gridDim.x * blockDim.x)
            ionise_cell_recomb(a, idx, a.fcoll[idx], mean_fix);
    } else if ((a.n & 3) == 0) {
        const long long n4 = a.n >> 2;
        const long long stride = (long long)gridDim.x * blockDim.x;
        const float4 *f4p = reinterpret_cast<const float4 *>(a.fcoll);
        for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += 4 * stride) {
            /* four independent 128-bit loads in flight per thread */
            float4 f4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long j = i4 + u * stride;
                f4[u] = (j < n4) ? f4p[j] : float4{0.f, 0.f, 0.f, 0.f};
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long j = i4 + u * stride;
                if (j < n4) {
                    ionise_cell(a, 4 * j + 0, f4[u].x, mean_fix);
                    ionise_cell(a, 4 * j + 1, f4[u].y, mean_fix);
                    ionise_cell(a, 4 * j + 2, f4[u].z, mean_fix);
                    ionise_cell(a, 4 * j + 3, f4[u].w, mean_fix);
                }
            }
        }
    } else {
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.n;
             idx += (long long)gridDim.x * blockDim.x)
            ionise_cell(a, idx, a.fcoll[idx], mean_fix);
    }
}

struct CritDeltaArgs {
    int nx, ny, nz, nzc;
    const float *filtered;   /* padded real rows of this radius (the sweep-1 input) */
    const DevTable *table;
    const double *partial;   /* x-plane sums of sweep 1 (plane_sum_kernel), all planes of the box */
    int n_partial;
    unsigned char *mask;     /* 1 = ionised at some radius so far */
    double n_cells, mean_f_coll, f_limit, ion_eff_factor;
    int mass_dep_zeta;
    /* speculation bookkeeping: the true mean fix of radius j is recorded in spec->mean_fix[j]; with
       only_if_overflow the kernel is the fallback of a speculative radius and returns at once unless that
       radius' queue overflowed; with resolve it decides the queued cells instead of walking the grid */
    SpecState *spec;
    int j, only_if_overflow;
    const uint2 *queue;         /* n_seg segments of qcap (cell, density) pairs, qcounts[s] of them wanted by segment s */
    const unsigned int *qcounts;
    unsigned int qcap;
    int n_seg;
};
/* sweep 2 for every radius but the last one processed: only the flag "f_coll zeta > 1" is needed
   (find_ionised_regions, IonisationBox.c:1040-1151, centre-cell method), so the f_coll grid is never
   written: the criterion is re-evaluated from the filtered density.  The reference compares
   mean_fix * (double)(float)f_coll * zeta with 1; for the log-valued table that is decided in log
   space whenever the interpolated log f_coll is more than 1e-6 away from the threshold (float
   rounding moves f_coll by < 6e-8 relative), and with the reference's exact arithmetic inside
   that band. */
template <bool LOG> __global__ void __launch_bounds__(256) ionise_delta_kernel(CritDeltaArgs a) {
    __shared__ SweepTable st;
    __shared__ double red[256];
    DYN_SMEM(float2, rep);
    sweep_table_load(&st, a.table, rep, false);
    double acc = 0.;
    for (int i = threadIdx.x; i < a.n_partial; i += blockDim.x) acc += a.partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    double grid_mean = red[0] / a.n_cells;
    if (a.mass_dep_zeta) {
        if (grid_mean <= a.f_limit) grid_mean = a.f_limit;
    } else {
        if (grid_mean <= pc::FRACT_FLOAT_ERR) grid_mean = pc::FRACT_FLOAT_ERR;
    }
    const double mean_fix = a.mean_f_coll / grid_mean;
    const double gain = mean_fix * a.ion_eff_factor;
    if (a.spec) {
        if (a.only_if_overflow) {
            if (!a.spec->overflow[a.j]) return; /* the queue held every bracket cell: nothing to do */
        } else if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.spec->mean_fix[a.j] = mean_fix;
        }
    }
    const bool floor_ionises = a.mass_dep_zeta && (a.f_limit * a.ion_eff_factor > 1.0);
    /* threshold in the table's own units: log f_coll or f_coll */
    const float thr = LOG ? (float)(-log(gain)) : (float)(1.0 / gain);
    const float band0 = LOG ? 1e-5f : 1e-5f * fabsf(thr);
    const float dens_floor = (float)(-1. + pc::FRACT_FLOAT_ERR);
    const SweepConstsF kf = sweep_consts(&st, dens_floor);
    const long long nrows = (long long)a.nx * a.ny;
    /* inside the band: the reference arithmetic on the float-rounded f_coll decides */
    auto exact = [&](float dens) -> bool {
        double curr = mean_fix * (double)(float)fcoll_exact(fmaxf(dens, dens_floor), &st);
        if (a.mass_dep_zeta && curr < a.f_limit) curr = a.f_limit;
        return curr * a.ion_eff_factor > 1.0;
    };
    /* two cells at a time, straight-line: bits 0/1 = ionised by the single-precision test,
       bits 4/5 = inside the band (the reference arithmetic must decide) */
    auto ionised2 = [&](float2 d) -> unsigned {
        int i0, i1;
        float2 t;
        table_coords_f2(d, kf, i0, i1, t);
        const float2 *rep = st.rep + sweep_rep_lane();
        const float2 y0 = rep[i0 * SWEEP_REP], y1 = rep[i1 * SWEEP_REP];
        const float2 diff = f2_add(f2_fma(t, make_float2(y0.y, y1.y), make_float2(y0.x, y1.x)), make_float2(-thr, -thr));
        const float b0 = fmaf(1e-4f, fabsf(y0.y), band0), b1 = fmaf(1e-4f, fabsf(y1.y), band0);
        const bool r0 = diff.x > 0.f ? true : floor_ionises, r1 = diff.y > 0.f ? true : floor_ionises;
        return (r0 ? 1u : 0u) | (r1 ? 2u : 0u) | (fabsf(diff.x) <= b0 ? 16u : 0u) | (fabsf(diff.y) <= b1 ? 32u : 0u);
    };
    auto set_mask = [&](unsigned m4, long long cell) {
        if (!m4) return;
        unsigned char *m = a.mask + cell;
        if (m4 == 15u) {
            *reinterpret_cast<unsigned int *>(m) = 0x01010101u;
        } else {
            if (m4 & 1u) m[0] = 1;
            if (m4 & 2u) m[1] = 1;
            if (m4 & 4u) m[2] = 1;
            if (m4 & 8u) m[3] = 1;
        }
    };
    if ((a.nz & 3) == 0) {
        for_each_chunk<unsigned>(
            a.filtered, nrows, a.nz, a.nzc, reinterpret_cast<float4 *>(rep + N_DENS_INTERP * SWEEP_REP),
            [&](const float4 &d) -> unsigned { return ionised2(make_float2(d.x, d.y)) | (ionised2(make_float2(d.z, d.w)) << 2); },
            [&](unsigned res, const float4 &d, long long row, int zc) {
                unsigned m4 = res & 15u;
                if (res & 0xf0u) { /* some cell of the chunk sits inside the band */
                    const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (res & (16u << i)) m4 = (m4 & ~(1u << i)) | (exact(dd[i]) ? (1u << i) : 0u);
                }
                set_mask(m4, row * a.nz + 4 * zc);
            });
    } else {
        for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
            const float *src = a.filtered + row * 2 * a.nzc;
            for (int z = threadIdx.x; z < a.nz; z += blockDim.x)
            {
                const unsigned res = ionised2(make_float2(src[z], src[z]));
                const bool ion = (res & 16u) ? exact(src[z]) : (res & 1u) != 0;
                if (ion) a.mask[row * a.nz + z] = 1;
            }
        }
    }
}

/* second half of a speculative radius: check the bracket against the true mean fix, then the queued cells */
__global__ void __launch_bounds__(256) spec_resolve_kernel(CritDeltaArgs a) {
    __shared__ SweepTable st;
    __shared__ double red[256];
    DYN_SMEM(float2, rep);
    sweep_table_load(&st, a.table, rep, false);
    double acc = 0.;
    for (int i = threadIdx.x; i < a.n_partial; i += blockDim.x) acc += a.partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    double grid_mean = red[0] / a.n_cells;
    if (a.mass_dep_zeta) {
        if (grid_mean <= a.f_limit) grid_mean = a.f_limit;
    } else {
        if (grid_mean <= pc::FRACT_FLOAT_ERR) grid_mean = pc::FRACT_FLOAT_ERR;
    }
    const double mean_fix = a.mean_f_coll / grid_mean;
    const double gain = mean_fix * a.ion_eff_factor;
    const SpecBracket br = spec_bracket(a.spec, a.j, a.ion_eff_factor);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.spec->mean_fix[a.j] = mean_fix;
        if (!(gain >= br.gain_lo && gain <= br.gain_hi)) a.spec->failed = 1;
    }
    if (!(gain >= br.gain_lo && gain <= br.gain_hi)) return;
    const float dens_floor = (float)(-1. + pc::FRACT_FLOAT_ERR);
    /* a segment that overflowed sends the whole radius to the fallback sweep; flags set here meanwhile are
       ones that sweep would set as well */
    for (int sg = blockIdx.x; sg < a.n_seg; sg += gridDim.x) {
        unsigned int n = a.qcounts[sg];
        if (threadIdx.x == 0) {
            atomic_fetch_add_u32(&a.spec->qcount[a.j], n);
            if (n > a.qcap) { a.spec->overflow[a.j] = 1; a.spec->failed = 1; } /* cells were dropped: the host re-runs the ladder */
        }
        if (n > a.qcap) n = a.qcap;
        const uint2 *seg = a.queue + (size_t)sg * a.qcap;
        for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint2 e = seg[i];
            const unsigned int cell = e.x;
            const float dens = int_bits_as_float((int)e.y);
            double curr = mean_fix * (double)(float)fcoll_exact(fmaxf(dens, dens_floor), &st);
            if (a.mass_dep_zeta && curr < a.f_limit) curr = a.f_limit;
            if (curr * a.ion_eff_factor > 1.0) a.mask[cell] = 1;
        }
    }
}

/* ComputeFullyIonizedTemperature (thermochem.c:31-56).  The two powers that do not depend on the
   cell, pow(T_re, 1.7) and pow(1e4 (1+z)/4, 1.7), are evaluated once per launch (c_Tre17, c_z17);
   the remaining ones are written as exp(p log x), which costs half of a general pow. */
DEV double pow_pos(double x, double p) { return exp(p * log(x)); }
DEV float fully_ionized_temperature(float z_re, float z, float delta, double c_Tre17, double c_z17) {
    float result, delta_re;
    if (fabs(z - z_re) < 1e-4)
        result = 1;
    else {
        delta_re = delta * (1. + z) / (1. + z_re);
        if (delta_re <= -1) delta_re = -1. + 9e-8;
        if (delta <= -1) delta = -1. + 9e-8;
        result = pow_pos((1. + delta) / (1. + delta_re), 1.1333);
        result *= pow_pos((1. + z) / (1. + z_re), 3.4);
        result *= expf(pow_pos((1. + z) / 7.1, 2.5) - pow_pos((1. + z_re) / 7.1, 2.5));
    }
    result *= c_Tre17;
    result += c_z17 * (1 + delta);
    result = pow_pos((double)result, 0.5882);
    return result;
}

struct FinalArgs {
    long long n;
    const unsigned char *mask;
    const float *density, *prev_zre;
    float *xH, *z_reion, *Tk;
    int *nonfinite;
    double redshift, stored_redshift, T_re, TK_nofluct, adia_TK_term;
    double c_Tre17, c_z17; /* pow(T_re, 1.7), pow(1e4 (1 + z) / 4, 1.7) */
    const float *Tk_neutral; /* USE_TS_FLUCT: the floor of the ionised gas temperature, else null */
    const unsigned char *paint; /* IONISE_ENTIRE_SPHERE: cells inside a painted sphere (x_HI = 0 only), else null */
};
/* materialise the flags (xH = 0, z_reion; IonisationBox.c:1142-1151) and set_ionized_temperatures
   (IonisationBox.c:1203-1256) in one pass */
__global__ void __launch_bounds__(256) finalize_kernel(FinalArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        float zre = -1.0f;
        if (a.mask[i]) {
            const float pz = a.prev_zre ? a.prev_zre[i] : -1.f;
            zre = (pz < 0) ? (float)a.redshift : pz;
            a.xH[i] = 0.f;
        } else if (a.paint && a.paint[i]) {
            a.xH[i] = 0.f; /* inside a sphere: ionised, but only the sphere's centre records z_reion and T_k */
        }
        a.z_reion[i] = zre;
        if (a.Tk) {
            float tk = a.Tk[i];
            if (a.mask[i] && zre > 0) {
                const float d = a.density[i];
                tk = fully_ionized_temperature(zre, (float)a.stored_redshift, d, a.c_Tre17, a.c_z17);
                const float thistk = a.Tk_neutral ? a.Tk_neutral[i] : (float)(a.TK_nofluct * (1 + a.adia_TK_term * d));
                if (tk < thistk) tk = thistk;
                a.Tk[i] = tk;
            }
            if (!isfinite(tk)) *a.nonfinite = 1;
        }
    }
}

/* Last radius of the plain ladder and the finalisation in ONE pass (ionise_kernel + finalize_kernel read and wrote
   the mask, x_HI and T_k twice: 3.0 ms of the step at 512^3, a fifth of the HBM rate each, one cell per thread
   and iteration).  Four cells per thread, every array touched once; the same per-cell arithmetic in the same
   order, so the outputs are bit-identical to the two-kernel sequence (B200_FUSED_LAST=0 keeps that one). */
struct LastArgs {
    CritArgs c;
    FinalArgs f;
};
DEV void last_cell(const LastArgs &a, double mean_fix, float fcoll, unsigned char m, float dens, float pz, float xh_in,
                   float tk_in, float &xh_out, float &zre_out, float &tk_out, bool &bad) {
    double curr_fcoll = mean_fix * (double)fcoll;
    if (a.c.mass_dep_zeta && curr_fcoll < a.c.f_limit) curr_fcoll = a.c.f_limit;
    float xh = xh_in, tk = tk_in;
    bool ion = m != 0;
    if (curr_fcoll * a.c.ion_eff_factor > 1.0) {
        ion = true;
    } else if (a.c.R_index == 0 && !ion && (xh_in > pc::TINY)) { /* ionise_cell's partial ionisation */
        double res_xH = 1. - curr_fcoll * a.c.ion_eff_factor;
        const float T_HI = (float)(a.c.TK_nofluct * (1 + a.c.adia_TK_term * dens));
        tk = partially_ionized_temperature(T_HI, (float)res_xH, (float)a.c.T_re);
        if (res_xH < 0) res_xH = 0;
        else if (res_xH > 1) res_xH = 1;
        xh = (float)res_xH;
    }
    float zre = -1.0f;
    if (ion) { /* finalize_kernel */
        zre = (pz < 0) ? (float)a.f.redshift : pz;
        xh = 0.f;
        if (zre > 0) {
            tk = fully_ionized_temperature(zre, (float)a.f.stored_redshift, dens, a.f.c_Tre17, a.f.c_z17);
            const float thistk = (float)(a.f.TK_nofluct * (1 + a.f.adia_TK_term * dens));
            if (tk < thistk) tk = thistk;
        }
    }
    if (!isfinite(tk)) bad = true;
    xh_out = xh; zre_out = zre; tk_out = tk;
}
__global__ void __launch_bounds__(256) ionise_last_fused_kernel(LastArgs a) {
    __shared__ double red[256];
    double acc = 0.;
    for (int i = threadIdx.x; i < a.c.n_partial; i += blockDim.x) acc += a.c.partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    double grid_mean = red[0] / a.c.n_cells;
    if (a.c.mass_dep_zeta) {
        if (grid_mean <= a.c.f_limit) grid_mean = a.c.f_limit;
    } else {
        if (grid_mean <= pc::FRACT_FLOAT_ERR) grid_mean = pc::FRACT_FLOAT_ERR;
    }
    const double mean_fix = a.c.mean_f_coll / grid_mean;
    bool bad = false;
    const long long n4 = a.c.n >> 2;
    const float4 *f4 = reinterpret_cast<const float4 *>(a.c.fcoll), *d4 = reinterpret_cast<const float4 *>(a.c.density);
    const float4 *p4 = reinterpret_cast<const float4 *>(a.c.prev_zre);
    const uchar4 *m4 = reinterpret_cast<const uchar4 *>(a.c.mask);
    float4 *x4 = reinterpret_cast<float4 *>(a.c.xH), *z4 = reinterpret_cast<float4 *>(a.c.z_reion), *t4 = reinterpret_cast<float4 *>(a.c.Tk);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 fc = f4[i], dn = d4[i], xin = x4[i];
        const uchar4 mk = m4[i];
        const float4 pz = p4 ? p4[i] : make_float4(-1.f, -1.f, -1.f, -1.f);
        const float4 tin = t4 ? t4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 xo, zo, to;
        last_cell(a, mean_fix, fc.x, mk.x, dn.x, pz.x, xin.x, tin.x, xo.x, zo.x, to.x, bad);
        last_cell(a, mean_fix, fc.y, mk.y, dn.y, pz.y, xin.y, tin.y, xo.y, zo.y, to.y, bad);
        last_cell(a, mean_fix, fc.z, mk.z, dn.z, pz.z, xin.z, tin.z, xo.z, zo.z, to.z, bad);
        last_cell(a, mean_fix, fc.w, mk.w, dn.w, pz.w, xin.w, tin.w, xo.w, zo.w, to.w, bad);
        x4[i] = xo;
        z4[i] = zo;
        if (t4) t4[i] = to;
    }
    if (bad && a.c.Tk) *a.f.nonfinite = 1;
}

/* set_recombination_rates, inhomogeneous model (IonisationBox.c:1277-1341): per cell, the
   PDF-integrated rate at the cell's effective redshift (1 + z) (1 + delta)^(1/3) - 1 and its own
   Gamma12, evaluated from the host-built table (host_recomb.cpp): index-sampled in redshift, natural
   cubic spline in ln(Gamma12) (splined_recombination_rate, recombinations.c:66-90). */
struct RecombArgs {
    long long n;
    const float *density, *xH, *G12, *prev_rec; /* prev_rec null = zeros */
    float *cum_rec;
    const double *lnGamma, *y, *c; /* [NG], [NZ][NG], [NZ][NG] */
    double lnGamma_max, one_plus_z, dt; /* dt = |dt/dz| dz in 1e15 s */
    int *nonfinite;
};
DEV double recomb_rate_lookup(const RecombArgs &a, double z_eff, double gamma12) {
    int z_ct = (int)(z_eff / RECOMB_DEL_Z + 0.5);
    if (z_ct < 0) z_ct = 0;
    else if (z_ct >= RECOMB_NZ) z_ct = RECOMB_NZ - 1;
    double lnG = log(gamma12);
    if (lnG < RECOMB_LNGAMMA_MIN) return 0;
    if (lnG >= a.lnGamma_max) lnG = a.lnGamma_max - pc::FRACT_FLOAT_ERR;
    int i = (int)((lnG - RECOMB_LNGAMMA_MIN) * 10.0);
    if (i < 0) i = 0;
    if (i > RECOMB_NG - 2) i = RECOMB_NG - 2;
    while (i > 0 && a.lnGamma[i] > lnG) i--;
    while (i < RECOMB_NG - 2 && a.lnGamma[i + 1] <= lnG) i++;
    const double *y = a.y + (size_t)z_ct * RECOMB_NG, *c = a.c + (size_t)z_ct * RECOMB_NG;
    const double dx = a.lnGamma[i + 1] - a.lnGamma[i], dy = y[i + 1] - y[i], t = lnG - a.lnGamma[i];
    const double b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double d = (c[i + 1] - c[i]) / (3.0 * dx);
    return y[i] + t * (b + t * (c[i] + t * d));
}
__global__ void __launch_bounds__(256) recomb_update_kernel(RecombArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        const double curr_dens = 1.0 + (double)a.density[i];
        const double z_eff = pow(curr_dens, 1.0 / 3.0) * a.one_plus_z;
        const double dNrec = recomb_rate_lookup(a, z_eff - 1., (double)a.G12[i]) * a.dt * (1. - (double)a.xH[i]);
        if (!isfinite(dNrec)) *a.nonfinite = 1;
        a.cum_rec[i] = (float)((double)(a.prev_rec ? a.prev_rec[i] : 0.f) + dNrec);
    }
}

/* box means of neutral_fraction and Gamma12 for the homogeneous model (IonisationBox.c:1594-1607):
   deterministic double block sums, finished on the host */
struct Mean2Args {
    long long n;
    const float *a, *b;
    double *partial; /* [gridDim][2] */
};
__global__ void __launch_bounds__(256) mean2_kernel(Mean2Args a) {
    __shared__ double ra[256], rb[256];
    double sa = 0., sb = 0.;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        sa += (double)a.a[i];
        sb += (double)a.b[i];
    }
    ra[threadIdx.x] = sa; rb[threadIdx.x] = sb;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { ra[threadIdx.x] += ra[threadIdx.x + s]; rb[threadIdx.x] += rb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { a.partial[2 * blockIdx.x] = ra[0]; a.partial[2 * blockIdx.x + 1] = rb[0]; }
}

/* IONISE_ENTIRE_SPHERE (update_in_sphere / check_region, bubble_helper_progs.c:263-413): every cell
   whose periodic distance to a centre flagged at this radius is below R (in cells of the x axis) is
   ionised.  The reference scatters a sphere from every flagged centre; here the set is the exact
   squared Euclidean distance transform of the centre flags, built separably -- nearest flagged cell
   along z, then the lower envelope of d_z^2 + dy^2 along y, then of that + dx^2 along x -- and
   thresholded with the reference's float comparison.  Offsets beyond ceil(R) cannot decide a cell, so
   every pass scans a window of 2 ceil(R) + 1 cells (at most the whole periodic axis). */
struct SphereArgs {
    int nx, ny, nz, h;            /* h = min(ceil(R), n/2) per axis is applied in the kernel */
    const unsigned char *centre;  /* flags of this radius */
    unsigned short *dz;           /* pass z out: distance to the nearest flagged cell of the row, capped */
    unsigned int *d2;             /* pass y out: min d_z^2 + dy^2, capped */
    unsigned char *paint, *mask;  /* pass x: paint |= (Rsq > d^2); mask |= centre */
    float Rsq;
};
constexpr unsigned int SPHERE_FAR = 0x3fffffffu;
__global__ void __launch_bounds__(256) sphere_z_kernel(SphereArgs a) {
    const long long n = (long long)a.nx * a.ny * a.nz;
    const int h = a.h < a.nz / 2 ? a.h : a.nz / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / a.nz;
        const int z = (int)(i - row * a.nz);
        const unsigned char *c = a.centre + row * a.nz;
        int best = 0xffff;
        for (int d = 0; d <= h; d++) {
            int zp = z + d; if (zp >= a.nz) zp -= a.nz;
            int zm = z - d; if (zm < 0) zm += a.nz;
            if (c[zp] | c[zm]) { best = d; break; }
        }
        a.dz[i] = (unsigned short)best;
    }
}
__global__ void __launch_bounds__(256) sphere_y_kernel(SphereArgs a) {
    const long long n = (long long)a.nx * a.ny * a.nz;
    const int h = a.h < a.ny / 2 ? a.h : a.ny / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % a.nz);
        const long long xy = i / a.nz;
        const int y = (int)(xy % a.ny);
        const long long x = xy / a.ny;
        unsigned int best = SPHERE_FAR;
        for (int d = -h; d <= h; d++) {
            int yy = y + d; if (yy >= a.ny) yy -= a.ny; else if (yy < 0) yy += a.ny;
            const unsigned int dz = a.dz[(x * a.ny + yy) * a.nz + z];
            if (dz != 0xffffu) {
                const unsigned int v = dz * dz + (unsigned int)(d * d);
                if (v < best) best = v;
            }
        }
        a.d2[i] = best;
    }
}
__global__ void __launch_bounds__(256) sphere_x_kernel(SphereArgs a) {
    const long long n = (long long)a.nx * a.ny * a.nz;
    const long long plane = (long long)a.ny * a.nz;
    const int h = a.h < a.nx / 2 ? a.h : a.nx / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i / plane);
        const long long yz = i - (long long)x * plane;
        unsigned int best = SPHERE_FAR;
        for (int d = -h; d <= h; d++) {
            int xx = x + d; if (xx >= a.nx) xx -= a.nx; else if (xx < 0) xx += a.nx;
            const unsigned int v = a.d2[(long long)xx * plane + yz];
            if (v != SPHERE_FAR && v + (unsigned int)(d * d) < best) best = v + (unsigned int)(d * d);
        }
        if (best != SPHERE_FAR && a.Rsq > (float)best) a.paint[i] = 1;
        if (a.centre[i]) a.mask[i] = 1;
    }
}

struct FillArgs {
    long long n;
    float *p;
    float v;
};
__global__ void fill_kernel(FillArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x)
        a.p[i] = a.v;
}
struct KeyInitArgs {
    int n;
    int *keys;
};
__global__ void minmax_key_init_kernel(KeyInitArgs a) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        a.keys[2 * i] = 2147483647;
        a.keys[2 * i + 1] = -2147483647 - 1;
    }
}

struct NeutralArgs {
    long long n;
    const float *density;
    float *xH, *Tk;
    float xH_val;
    double TK_nofluct, adia_TK_term;
    const float *xe, *Tk_neutral; /* USE_TS_FLUCT: x_HI = 1 - x_e and the TsBox's temperature, else null */
};
/* set_fully_neutral_box (IonisationBox.c:531-565) */
__global__ void neutral_box_kernel(NeutralArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        if (a.xe) {
            a.xH[i] = (float)(1. - (double)a.xe[i]);
            if (a.Tk) a.Tk[i] = a.Tk_neutral[i];
        } else {
            a.xH[i] = a.xH_val;
            if (a.Tk) a.Tk[i] = a.TK_nofluct * (1.0 + a.adia_TK_term * a.density[i]);
        }
    }
}

/* ------------------------------------------------------------------ orchestration */
static void sweep_smem_optin() {
#ifndef B200_EMU
    static bool done = false;
    if (done) return;
    const int bytes = (int)SWEEP_REP_BYTES;
    CUDA_CHECK(cudaFuncSetAttribute(fcoll_sum_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CUDA_CHECK(cudaFuncSetAttribute(fcoll_sum_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CUDA_CHECK(cudaFuncSetAttribute(fcoll_sum_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CUDA_CHECK(cudaFuncSetAttribute(fcoll_sum_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CUDA_CHECK(cudaFuncSetAttribute(spec_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CUDA_CHECK(cudaFuncSetAttribute(ionise_delta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CUDA_CHECK(cudaFuncSetAttribute(ionise_delta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done = true;
#endif
}
static int grid_for(long long n_items, int per_block) {
    long long want = (n_items + per_block - 1) / per_block;
    long long cap = (long long)dev_num_sms() * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

struct IonDeviceIO {
    const float *density;   /* device, N */
    const float *prev_zre;  /* device or null */
    float *xH, *z_reion, *Tk, *nion; /* device, N each; Tk / nion may be null */
    int wait_slot = -1;     /* copy-stream event that must have fired before xH / Tk / prev_zre are read
                               (their upload overlaps the radius ladder), or -1 */
    bool *nion_written = nullptr; /* out: the ladder stored a grid in `nion` (the reference leaves the
                                     caller's zero-initialised array untouched when it exits early) */
    /* recombinations (RECOMB_MODEL != none) */
    const float *prev_rec = nullptr;  /* device, N: previous cumulative recombinations (inhomogeneous), null = zeros */
    double prev_rec_scalar = 0.;      /* the one global value of the homogeneous model */
    float *G12 = nullptr, *mfp = nullptr, *cum_rec = nullptr; /* device, N; mfp optional; cum_rec inhomogeneous only */
    double *cum_rec_scalar_out = nullptr; /* host: new global value (homogeneous) */
    bool *rec_written = nullptr;      /* out: G12 / mfp / cumulative recombinations were updated */
    /* spin-temperature inputs (USE_TS_FLUCT): device, N each */
    const float *xe = nullptr, *Tk_neutral = nullptr;
    /* Lagrangian source grids (HaloBox): device, N each; whalo_sfr only with recombinations */
    const float *halo_nion = nullptr, *halo_wsfr = nullptr;
    double log10_Mcrit_ACG_ave = 0., log10_Mcrit_MCG_ave = 0.;
};

/* pinned staging that outlives a call (cudaMallocHost is too slow to repeat per call) */
struct IonStaging {
    int cap = 0;
    int *h_keys = nullptr;       /* [cap][2] */
    DevTable *h_tables = nullptr; /* [cap] */
    void *events[MAX_RADII];
    ~IonStaging() {}
    void ensure(int n) {
        if (n <= cap) return;
        if (n > MAX_RADII) b200_throw(B200_ValueError, "more than %d filter radii", MAX_RADII);
        host_pinned_free(h_keys);
        host_pinned_free(h_tables);
        const int want = n <= 64 ? 64 : MAX_RADII; /* the usual ladders stay on the small staging buffers */
        h_keys = (int *)host_pinned_alloc(sizeof(int) * 2 * want);
        h_tables = (DevTable *)host_pinned_alloc(sizeof(DevTable) * want);
        for (int i = cap; i < want; i++) events[i] = dev_event_create();
        cap = want;
    }
};
static IonStaging g_stage;

/* Radius-parallel execution of ONE box on several GPUs: given delta_k the radii
   are independent units -- each needs only its own filtered grid, table, grid sum and flags, and
   the flags combine by OR.  phase 0: this rank runs the radii k = part (mod nparts) except the last
   one and leaves their flags in `mask` (N bytes, device); the caller ORs the masks of all ranks
   (all-reduce MAX over bytes, the path's only collective).  phase 1: every rank runs the last
   radius on the merged mask (it assigns the partial ionisations of the never-flagged cells) and
   finalises, so every rank ends with the complete box.  phase -1: the whole ladder, no partition. */
struct IonPartition {
    int part = 0, nparts = 1, phase = -1;
    unsigned char *mask = nullptr;
    /* slab = true: the box is split into x-slabs over the ranks of dist.h (every array of the call is
       this rank's slab [nxl][ny][nz]); the transforms are the slab-decomposed ones of fft.cu, the
       per-radius extrema and plane sums are combined by dist_barrier_minmax / dist_barrier_gather.
       Same arithmetic per cell and the same reduction tree: bit-identical to the single-GPU ladder. */
    bool slab = false;
};

/* rows per chunk of the sum sweep: a multiple of the fast path's rows per iteration where that
   divides ny (so that the cp.async ring applies), grown to >= 32 rows; else the largest divisor of ny <= 32 */
static int sweep_chunk_rows(int ny, int nz) {
    int step_rows = 0;
    if ((nz & 3) == 0) {
        const int q = nz >> 2;
        if (q <= 256 && 256 % q == 0) step_rows = 4 * (256 / q);
    }
    if (step_rows && ny % step_rows == 0) {
        int ch = step_rows;
        while (ch < 32 && ny % (2 * ch) == 0) ch *= 2;
        return ch;
    }
    int ch = 1;
    for (int d = 1; d <= 32 && d <= ny; d++)
        if (ny % d == 0) ch = d;
    return ch;
}
/* CTAs of the sum sweep: the largest divisor of the chunk count that is resident at once (no tail) */
static int sweep_grid(long long nchunks) {
    static int mult = 0; /* resident 256-thread CTAs per SM the sweeps are sized for (B200_SWEEP_CTAS overrides) */
    if (mult == 0) {
        const char *e = getenv("B200_SWEEP_CTAS");
        mult = e ? atoi(e) : 8;
        if (mult < 1 || mult > 16) mult = 8;
    }
    const long long cap = (long long)dev_num_sms() * mult;
    if (nchunks <= cap) return (int)(nchunks > 0 ? nchunks : 1);
    for (long long g = cap; g >= cap / 2; g--)
        if (nchunks % g == 0) return (int)g;
    return (int)cap;
}

/* set_recombination_rates (IonisationBox.c:1258-1342) after the ladder */
static void recomb_update(int recomb, const IonDeviceIO &io, const IonConsts &c, long long N, int *d_flag) {
    if (recomb == 2) { /* set_recombination_rates, inhomogeneous (IonisationBox.c:1277-1341) */
        const RecombTables *rt = recomb_tables();
        const size_t tn = (size_t)RECOMB_NZ * RECOMB_NG;
        DevBuf<double> d_rr(2 * tn + RECOMB_NG);
        h2d(d_rr.p, rt->y.data(), tn * sizeof(double));
        h2d(d_rr.p + tn, rt->c.data(), tn * sizeof(double));
        h2d(d_rr.p + 2 * tn, rt->lnGamma, RECOMB_NG * sizeof(double));
        g_stats.h2d -= (long long)((2 * tn + RECOMB_NG) * sizeof(double));
        RecombArgs ra = {N, io.density, io.xH, io.G12, io.prev_rec, io.cum_rec, d_rr.p + 2 * tn, d_rr.p, d_rr.p + tn,
                         rt->lnGamma_max, 1. + c.stored_redshift, c.fabs_dtdz * c.dz, d_flag};
        B200_LAUNCH(recomb_update_kernel, grid_for(N, 1024), 256, 0, ra);
        int flag = 0;
        d2h(&flag, d_flag, sizeof(int));
        g_stats.d2h -= (long long)sizeof(int);
        if (flag) b200_throw(B200_InfinityorNaNError, "recombinations returned an infinite or NaN value");
    } else if (recomb == 1) { /* homogeneous (IonisationBox.c:1261-1276): one rate from the box means */
        const int nb = grid_for(N, 1024);
        DevBuf<double> d_p(2 * (size_t)nb);
        Mean2Args ma = {N, io.xH, io.G12, d_p};
        B200_LAUNCH(mean2_kernel, nb, 256, 0, ma);
        std::vector<double> hp(2 * (size_t)nb);
        d2h(hp.data(), d_p, hp.size() * sizeof(double));
        g_stats.d2h -= (long long)(hp.size() * sizeof(double));
        double sx = 0., sg = 0.;
        for (int i = 0; i < nb; i++) { sx += hp[2 * i]; sg += hp[2 * i + 1]; }
        const double global_xH = sx / (double)N;
        const float global_G12 = (float)(sg / (double)N); /* a float in the reference */
        const double dNrec = recomb_rate_host(c.stored_redshift, global_G12) * c.fabs_dtdz * c.dz * (1. - global_xH);
        const double cum = io.prev_rec_scalar + dNrec;
        if (!std::isfinite(cum)) b200_throw(B200_InfinityorNaNError, "non-finite cumulative recombinations");
        if (io.cum_rec_scalar_out) *io.cum_rec_scalar_out = cum;
    }
}

static void ionize_core(float redshift_f, float prev_redshift_f, const IonDeviceIO &io, IonizedBox *box,
                        const IonPartition &pt = IonPartition()) {
    const SimulationOptions *so = simulation_options_global;
    const AstroOptions *ao = astro_options_global;
    const MatterOptions *mo = matter_options_global;
    if (mo->SOURCE_MODEL != SRC_CONST_ION_EFF && mo->SOURCE_MODEL != SRC_E_INTEGRAL)
        b200_throw(B200_ValueError, "SOURCE_MODEL=%d: only CONST-ION-EFF and E-INTEGRAL are in scope", mo->SOURCE_MODEL);
    if (ao->USE_MINI_HALOS || ao->PHOTON_CONS_TYPE != 0)
        b200_throw(B200_ValueError, "USE_MINI_HALOS / photon conservation are outside the scoped IonizeBox path");
    /* USE_TS_FLUCT: the caller's TsBox supplies x_e (filtered per radius) and the neutral-gas temperature */
    const bool ts = ao->USE_TS_FLUCT;
    if (ts) {
        if (!io.xe || !io.Tk_neutral)
            b200_throw(B200_ValueError, "USE_TS_FLUCT needs the TsBox's xray_ionised_fraction and kinetic_temp_neutral");
        if (pt.phase >= 0 || pt.slab) b200_throw(B200_ValueError, "the multi-GPU ladders are not built for USE_TS_FLUCT");
    }
    /* recombinations: 1 homogeneous (one global N_rec), 2 inhomogeneous (per cell; filtered with the
       density unless CELL_RECOMB) */
    const int recomb = ao->RECOMB_MODEL;
    const bool filter_rec = recomb != 0 && !ao->CELL_RECOMB;
    if (recomb) {
        if (recomb == 1 && filter_rec)
            b200_throw(B200_ValueError, "RECOMB_MODEL=homogeneous needs CELL_RECOMB (there is no N_rec grid to filter)");
        if (pt.phase >= 0 || pt.slab) b200_throw(B200_ValueError, "the multi-GPU ladders are not built for RECOMB_MODEL != none");
        if (!io.G12 || (recomb == 2 && !io.cum_rec))
            b200_throw(B200_ValueError, "RECOMB_MODEL != none needs ionisation_rate_G12 and cumulative_recombinations");
        if (!recomb_tables()) b200_throw(B200_TableEvaluationError, "RECOMB_MODEL != none needs init_MHR()");
    }
    if (mo->USE_INTERPOLATION_TABLES != 2)
        b200_throw(B200_ValueError, "this build needs USE_INTERPOLATION_TABLES='hmf-interpolation'");

    const double redshift = redshift_f, prev_redshift = prev_redshift_f;
    IonConsts c;
    set_ionbox_constants(redshift, prev_redshift, &c);
    const int nx = so->HII_DIM, ny = so->HII_DIM, nz = hii_d_para();
    const long long N = (long long)nx * ny * nz; /* cells of the whole box */
    Fft3D *plan = fft_plan(nx, ny, nz);
    /* slab decomposition: this call owns x-planes [x0, x0 + nxl) */
    FftSlab slab;
    const bool sl = pt.slab;
    if (sl) {
        if (pt.phase >= 0) b200_throw(B200_ValueError, "slab and radius partitions do not combine");
        dist_reset();
        slab = fft_slab_setup(plan);
    }
    const int nxl = sl ? slab.nxl : nx;
    const long long NL = (long long)nxl * ny * nz; /* cells of this call's arrays */
    const size_t kbox_n = sl ? slab.n_cplx() : plan->n_cplx();

    std::vector<RadiusSpec> radii = setup_radii(c);
    const int n_radii = (int)radii.size();

    /* box-level scalars (IonisationBox.c:1431-1455) */
    double Mturn_avg;
    if (c.mass_dep_zeta) {
        Mturn_avg = astro_params_global->M_TURN;
        box->log10_Mturnover_ave = log10(Mturn_avg);
    } else {
        Mturn_avg = c.M_min;
        box->log10_Mturnover_ave = log10(c.M_min);
    }
    box->log10_Mturnover_MINI_ave = 0.0;
    const int method = ao->INTEGRATION_METHOD_ATOMIC;
    if (method == INTEG_GL) initialise_GL(c.lnMmin, c.lnMmax_gl);

    /* set_mean_fcoll (IonisationBox.c:468-529) */
    double f_limit = 0.;
    if (c.mass_dep_zeta) {
        box->mean_f_coll = Nion_General(redshift, c.lnMmin, c.lnMmax_gl, Mturn_avg, &c.sc);
        f_limit = Nion_General(so->Z_HEAT_MAX, c.lnMmin, c.lnMmax_gl, Mturn_avg, &c.sc);
    } else {
        box->mean_f_coll = Fcoll_General(redshift, c.lnMmin, c.lnMmax_gl);
        f_limit = Fcoll_General(so->Z_HEAT_MAX, c.lnMmin, c.lnMmax_gl);
    }
    box->mean_f_coll_MINI = 0.;
    if (!std::isfinite(box->mean_f_coll) || box->mean_f_coll < 0)
        b200_throw(B200_InfinityorNaNError, "Mean collapse fraction is invalid");

    const double exp_global_hii = box->mean_f_coll * c.ion_eff_factor_gl;
    if (exp_global_hii < HII_ROUND_ERR) {
        if (pt.phase == 0) { dev_zero(pt.mask, (size_t)N); dev_sync(); return; }
        if (io.wait_slot >= 0) main_wait_copy_event(io.wait_slot);
        { FillArgs f = {NL, io.z_reion, -1.0f}; B200_LAUNCH(fill_kernel, grid_for(NL, 1024), 256, 0, f); }
        NeutralArgs na = {NL, io.density, io.xH, io.Tk, (float)(1. - xion_RECFAST(redshift)), c.TK_nofluct, c.adia_TK_term,
                          ts ? io.xe : nullptr, ts ? io.Tk_neutral : nullptr};
        B200_LAUNCH(neutral_box_kernel, grid_for(NL, 1024), 256, 0, na);
        if (sl) { dist_barrier(); dist_check(); }
        return;
    }

    /* radii that will actually be processed, largest first (IonisationBox.c:1531-1541) */
    std::vector<int> todo;
    for (int R_ct = n_radii; R_ct--;) {
        if (c.M_min > RtoM(radii[R_ct].R)) break;
        todo.push_back(R_ct);
    }
    const int n_todo = (int)todo.size();
    /* the ladder steps this call runs, in order */
    std::vector<int> mine;
    for (int k = 0; k < n_todo; k++) {
        const bool last_k = (k == n_todo - 1);
        if (pt.phase < 0 || (pt.phase == 0 && !last_k && k % pt.nparts == pt.part) || (pt.phase == 1 && last_k))
            mine.push_back(k);
    }
    const int n_mine = (int)mine.size();

    /* The stream runs the filter + transform of up to `ahead` later radii while the host waits for
       the extrema of radius j and integrates its table.  One radius ahead hides the host at 512^3
       (1.1 ms of kernels per radius against ~0.1 ms of table build); at 256^3 (0.2 ms of kernels per
       radius) two ahead is measurably better (12.0 -> 11.3 ms per step), at 128^3 and below the
       host is the serial resource and running further ahead only delays the sweeps (7.2 -> 7.5 ms).
       B200_IONIZE_AHEAD = 1..3 overrides. */
    const size_t work_bytes = kbox_n * sizeof(float2);
    int ahead = (work_bytes >= ((size_t)32 << 20) && work_bytes <= ((size_t)256 << 20)) ? 2 : 1;
    if (const char *e = getenv("B200_IONIZE_AHEAD")) { ahead = atoi(e); if (ahead < 1) ahead = 1; if (ahead > 3) ahead = 3; }
    const int NW = ahead + 1;
    DevBuf<float2> k_unfiltered(kbox_n);
    DevBuf<float2> work_ring[4];
    float2 *work[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < NW; i++) { work_ring[i].alloc(kbox_n); work[i] = work_ring[i].p; }
    DevBuf<float> d_fcoll;
    const bool general = recomb != 0 || ts; /* per-cell barrier: the reference's arithmetic at every radius */
    if (!io.nion || general) d_fcoll.alloc(NL);
    /* x_e in k space and its filtered copies, one per work box */
    DevBuf<float2> k_xe, xe_ring[4];
    float2 *work_xe[4] = {nullptr, nullptr, nullptr, nullptr};
    if (ts) {
        k_xe.alloc(plan->n_cplx());
        for (int i = 0; i < NW; i++) { xe_ring[i].alloc(plan->n_cplx()); work_xe[i] = xe_ring[i].p; }
    }
    /* N_rec of the previous snapshot in k space and its filtered copies, one per work box */
    const bool rec_grid_filtered = filter_rec && io.prev_rec;
    DevBuf<float2> k_nrec, rec_ring[4];
    float2 *work_rec[4] = {nullptr, nullptr, nullptr, nullptr};
    if (rec_grid_filtered) {
        k_nrec.alloc(plan->n_cplx());
        for (int i = 0; i < NW; i++) { rec_ring[i].alloc(plan->n_cplx()); work_rec[i] = rec_ring[i].p; }
    }
    DevBuf<int> d_keys(2 * (size_t)(n_todo > 0 ? n_todo : 1));
    DevBuf<DevTable> d_tables((size_t)(n_todo > 0 ? n_todo : 1));
    /* fixed reduction tree of the grid sum: chunk sums -> x-plane sums -> sum over all planes of the box */
    const int chunk_rows = sweep_chunk_rows(ny, nz);
    const int chunks_per_plane = ny / chunk_rows;
    const long long nchunks = (long long)nxl * chunks_per_plane;
    const int sum_blocks = sweep_grid(nchunks);
    const int sweep_blocks = grid_for((long long)nxl * ny, 1);
    DevBuf<double> d_partial((size_t)nchunks * SWEEP_PARTIALS);
    DevBuf<double> d_plane((size_t)nx);
    /* slab mode: this rank's extrema keys and plane sums are published in the symmetric heap, one slot per radius */
    int *keys_sym = nullptr;
    double *plane_sym = nullptr;
    float *tab_sym = nullptr; /* this rank's share of every radius' table */
    unsigned long long *fail_sym = nullptr; /* this rank's SpecState::failed, so that the ranks re-run the ladder together */
    const int tab_per = sl ? (N_DENS_INTERP + slab.P - 1) / slab.P : 0;
    if (sl) {
        keys_sym = (int *)dist_alloc(sizeof(int) * 2 * MAX_RADII);
        plane_sym = (double *)dist_alloc(sizeof(double) * MAX_RADII * (size_t)nxl);
        if (!(getenv("B200_SPLIT_TABLES") && getenv("B200_SPLIT_TABLES")[0] == '0'))
            tab_sym = (float *)dist_alloc(sizeof(float) * MAX_RADII * (size_t)tab_per);
        fail_sym = (unsigned long long *)dist_alloc(sizeof(unsigned long long));
    }
    DevBuf<unsigned long long> d_fail_all(sl ? (size_t)slab.P : 0);
    DevBuf<int> d_flag(1);
    dev_zero(d_flag, sizeof(int));
    DevBuf<unsigned char> own_mask;
    unsigned char *d_mask = pt.mask;
    if (pt.phase < 0) { own_mask.alloc((size_t)NL); d_mask = own_mask; }
    if (pt.phase <= 0) dev_zero(d_mask, (size_t)NL); /* phase 1 continues on the merged mask */
    /* IONISE_ENTIRE_SPHERE: the flags of each radius are kept apart (d_centre), dilated by the radius'
       sphere into d_paint and merged into d_mask (the centres, which alone record z_reion and T_k) */
    const bool sphere = ao->IONISE_ENTIRE_SPHERE;
    auto sphere_radius_cells = [&](double R) { return (float)(R / so->BOX_LEN) * (float)so->HII_DIM; };
    if (sphere) {
        if (general || pt.phase >= 0 || sl)
            b200_throw(B200_ValueError, "IONISE_ENTIRE_SPHERE is built for the plain ladder only (no recombinations, "
                                        "spin temperature or radius partition: the reference's result then depends "
                                        "on its cell visiting order)");
        /* at the unfiltered radius the partial ionisations of the same sweep would depend on the visiting
           order unless the sphere is the centre cell alone */
        if (n_todo > 0 && radii[todo.back()].R_index == 0) {
            const float rc = sphere_radius_cells(radii[todo.back()].R);
            if ((float)pow((double)rc, 2) > 1.0f)
                b200_throw(B200_ValueError, "IONISE_ENTIRE_SPHERE with R_BUBBLE_MIN above one cell is order-dependent "
                                            "in the reference and not built");
        }
    }
    DevBuf<unsigned char> d_centre(sphere ? (size_t)N : 0), d_paint(sphere ? (size_t)N : 0);
    DevBuf<unsigned short> d_sdz(sphere ? (size_t)N : 0);
    DevBuf<unsigned int> d_sd2(sphere ? (size_t)N : 0);
    if (sphere) dev_zero(d_paint, (size_t)N);
    auto paint_spheres = [&](double R) {
        const float rc = sphere_radius_cells(R);
        SphereArgs sa = {nx, ny, nz, (int)ceil((double)rc), d_centre, d_sdz, d_sd2, d_paint, d_mask, (float)pow((double)rc, 2)};
        const int nb = grid_for(N, 256);
        B200_LAUNCH(sphere_z_kernel, nb, 256, 0, sa);
        B200_LAUNCH(sphere_y_kernel, nb, 256, 0, sa);
        B200_LAUNCH(sphere_x_kernel, nb, 256, 0, sa);
    };
    /* window tables over |n|^2 (cubic boxes, top-hat / gaussian): one slot per work box, stream-ordered reuse */
    const bool cubic = nx == ny && ny == nz && so->NON_CUBIC_FACTOR == 1.0f;
    const bool use_wtab = cubic && (c.hii_filter == 0 || c.hii_filter == 2);
    const int wtab_n = use_wtab ? window_table_size(plan) : 0;
    DevBuf<float> d_wtab(use_wtab ? (size_t)NW * wtab_n : 0);
    /* expanded [|nx|][|ny|][kz] copy for the coalesced x-pass lookup (power-of-two grids, <= 1 GB per slot) */
    const size_t wtab3_n = use_wtab ? window_table3_size(plan) : 0;
    const bool use_wtab3 = use_wtab && (nx & (nx - 1)) == 0 && nx >= 16 && wtab3_n * sizeof(float) <= ((size_t)1 << 30);
    DevBuf<float> d_wtab3(use_wtab3 ? (size_t)NW * wtab3_n : 0);
#ifndef B200_EMU
    const bool overlap_tables = use_wtab && !(getenv("B200_TABLE_OVERLAP") && getenv("B200_TABLE_OVERLAP")[0] == '0');
#else
    const bool overlap_tables = false;
#endif
    g_stage.ensure(n_todo);
    if (n_todo > 0) {
        KeyInitArgs ka = {n_todo, sl ? keys_sym : d_keys.p};
        B200_LAUNCH(minmax_key_init_kernel, 1, 64, 0, ka);
    }

    /* prepare_box_for_filtering (IonisationBox.c:323-360) */
    ZPrologue pro;
    pro.src = io.density; pro.src_row_stride = nz; pro.premul = 1.f;
    pro.clip = 1; pro.clip_lo = -1.f; pro.clip_hi = 1e6f;
    pro.post_scale = 1.f / (float)N;
    if (sl) fft_r2c_slab(&slab, k_unfiltered, work[0], pro);
    else fft_r2c(plan, k_unfiltered, pro);
    if (general && io.wait_slot >= 0) main_wait_copy_event(io.wait_slot); /* xH / G12 / N_rec / x_e are read at every radius */
    if (ts) { /* prepare_box_for_filtering(xray_ionised_fraction, 0, 1) (IonisationBox.c:1510-1513) */
        ZPrologue px = pro;
        px.src = io.xe; px.clip_lo = 0.f; px.clip_hi = 1.f;
        fft_r2c(plan, k_xe, px);
    }
    if (rec_grid_filtered) {
        ZPrologue pr = pro;
        pr.src = io.prev_rec; pr.clip_lo = 0.f; pr.clip_hi = 1e20f;
        fft_r2c(plan, k_nrec, pr);
    }

    const double dk0 = 2.0 * M_PI / so->BOX_LEN;
    const double dkz = 2.0 * M_PI / (so->BOX_LEN * so->NON_CUBIC_FACTOR);

    /* stage A of radius k: filter + c2r + clip + min/max keys; the two keys travel to pinned host
       memory behind an event, so the host can wait for exactly this radius while the stream
       already runs the next one (copy_filter_transform + clip_and_get_extrema) */
    auto enqueue_transform = [&](int j) {
        const int k = mine[j];
        const RadiusSpec &rs = radii[todo[k]];
        KMul km;
        if (rs.R_index > 0) {
            km.kind = KMUL_FILTER; km.filter_type = c.hii_filter; km.R = (float)rs.R; km.fast = 1;
            km.dk[0] = dk0; km.dk[1] = dk0; km.dk[2] = dkz;
            if (use_wtab) {
                /* the two small table kernels run on the auxiliary stream beside the main stream's
                   passes of earlier radii; the slot is free once the x pass that read it NW radii
                   ago has finished (event 64 + slot), the x pass of this radius waits for event j */
                float *slot = d_wtab.p + (size_t)(j % NW) * wtab_n;
                struct AuxGuard { bool on; ~AuxGuard() { if (on) rt_use_aux(false); } } guard{overlap_tables};
                if (overlap_tables) {
                    rt_use_aux(true);
                    if (j >= NW) rt_stream_wait(64 + (j % NW));
                }
                window_table_build(plan, c.hii_filter, km.R, dk0, slot);
                km.wtab = slot; km.wtab_n = wtab_n;
                if (use_wtab3) {
                    float *slot3 = d_wtab3.p + (size_t)(j % NW) * wtab3_n;
                    if (sl) window_table_expand(plan, slot, slot3, slab.y0, slab.y0 + slab.nyl - 1);
                    else window_table_expand(plan, slot, slot3);
                    km.wtab3 = slot3;
                }
                if (overlap_tables) {
                    rt_event_record(j % 64);
                    rt_use_aux(false);
                    guard.on = false;
                    rt_stream_wait(j % 64);
                }
            }
        }
        ZEpilogue epi;
        epi.scale = 1.f; epi.clip = 1; epi.clip_lo = -1.f; epi.clip_hi = 1e6f;
        epi.minmax_keys = (sl ? keys_sym : d_keys.p) + 2 * k;
        if (sl) {
            fft_c2r_slab(&slab, k_unfiltered, work[j % NW], km, epi);
            dist_barrier_minmax(keys_sym + 2 * k, d_keys.p + 2 * k); /* extrema of the whole box on every rank */
        } else {
            fft_c2r(plan, k_unfiltered, work[j % NW], km, epi);
        }
        if (rec_grid_filtered) { /* <N_rec> over the same window, floored at zero (IonisationBox.c:806-809) */
            ZEpilogue er;
            er.scale = 1.f; er.clip = 1; er.clip_lo = 0.f; er.clip_hi = 3.0e38f;
            fft_c2r(plan, k_nrec, work_rec[j % NW], km, er);
        }
        if (ts) { /* x_e over the same window, kept inside [0, 0.999] (IonisationBox.c:812-818) */
            ZEpilogue ex;
            ex.scale = 1.f; ex.clip = 1; ex.clip_lo = 0.f; ex.clip_hi = 0.999f;
            fft_c2r(plan, k_xe, work_xe[j % NW], km, ex);
        }
        if (overlap_tables) rt_event_record(64 + (j % NW)); /* this radius' table slot is free again */
        d2h_async(g_stage.h_keys + 2 * k, d_keys.p + 2 * k, 2 * sizeof(int));
        dev_event_record(g_stage.events[k]);
    };

    /* speculative single-sweep radii (SpecState): plain ladder on one GPU or on slabs, never with a per-cell
       barrier, sphere painting or a floor that ionises everything; B200_SPEC=0 switches it off */
    bool use_spec = !general && !sphere && pt.phase < 0 && (nz & 3) == 0 && NL < ((long long)1 << 32) && n_mine > 3 &&
                    !(c.mass_dep_zeta && f_limit * c.ion_eff_factor > 1.0) &&
                    !(getenv("B200_SPEC") && getenv("B200_SPEC")[0] == '0');
    DevBuf<SpecState> d_spec(1);
    /* a private queue segment per CTA of the sum sweep, a sixteenth of the CTA's cells (a per cent or two of the
       cells sit inside the bracket) */
    unsigned int qcap = use_spec ? (unsigned int)((NL / 16) / sum_blocks > 256 ? (NL / 16) / sum_blocks : 256) : 0;
    if (use_spec && getenv("B200_SPEC_QCAP")) qcap = (unsigned int)atoi(getenv("B200_SPEC_QCAP")); /* tests: force overflows */
    DevBuf<uint2> d_queue((size_t)qcap * (use_spec ? sum_blocks : 0));
    DevBuf<unsigned int> d_qcounts(use_spec ? sum_blocks : 0);

    FcollTable htab;
    bool fused_last = false; /* the last radius' kernel also finalised the box */
    double t_wait = 0, t_table = 0, t_launch = 0;
    const bool verbose = getenv("B200_TIMING") != nullptr;
    for (int attempt = 0; attempt < 2; attempt++) {
    bool restart = false;
    dev_zero(d_spec, sizeof(SpecState));
    if (use_spec) {
        const double eps = getenv("B200_SPEC_EPS") ? atof(getenv("B200_SPEC_EPS")) : SPEC_EPS;
        h2d(d_spec.p, &eps, sizeof(double));
        g_stats.h2d -= (long long)sizeof(double);
    }
    if (attempt > 0) { /* the prediction of the mean fix failed somewhere: the same ladder, two sweeps per radius */
        use_spec = false;
        if (pt.phase <= 0) dev_zero(d_mask, (size_t)NL);
        KeyInitArgs ka = {n_todo, sl ? keys_sym : d_keys.p};
        B200_LAUNCH(minmax_key_init_kernel, 1, 64, 0, ka);
        if (verbose) fprintf(stderr, "[21cmfast_b200] ionize: mean-fix prediction left its bracket (or a queue segment overflowed), ladder re-run without speculation\n");
    }
    int next_enq = 0;
    if (n_mine > 0) enqueue_transform(next_enq++);
    for (int j = 0; j < n_mine; j++) {
        const int k = mine[j];
        const RadiusSpec &rs = radii[todo[k]];
        double t0 = omp_get_wtime();
        while (next_enq < n_mine && next_enq <= j + ahead) enqueue_transform(next_enq++);
        double t1 = omp_get_wtime();
        dev_event_wait_host(g_stage.events[k]);
        double t2 = omp_get_wtime();
        const double min_density = (double)float_from_order_key(g_stage.h_keys[2 * k]) - 0.001;
        const double max_density = (double)float_from_order_key(g_stage.h_keys[2 * k + 1]) + 0.001;

        /* setup_integration_tables (IonisationBox.c:702-768) on the host while the stream works.  On slabs every
           rank sees the same extrema, so the ranks share the 400 quadratures of the table round-robin and
           exchange their entries through the symmetric heap (at 8 GPUs the per-radius table, not the kernels,
           bounded the ladder: 0.13-0.24 ms of host time against 0.2 ms of GPU time per radius) */
        const bool split_table = sl && slab.P > 1 && c.mass_dep_zeta && tab_sym;
        if (c.mass_dep_zeta) {
            if (method == INTEG_GL) initialise_GL(c.lnMmin, rs.ln_M_max_R);
            build_nion_table(&htab, c.redshift, min_density, max_density, c.M_min, rs.M_max_R, &c.sc, method, so->N_THREADS,
                             split_table ? slab.rank : 0, split_table ? slab.P : 1);
        } else {
            build_fgtrm_table(&htab, min_density, max_density, c.growth_factor, c.sigma_minmass, rs.sigma_maxmass);
        }
        double t3 = omp_get_wtime();
        t_launch += t1 - t0; t_wait += t2 - t1; t_table += t3 - t2;
        DevTable *st = &g_stage.h_tables[k];
        st->x_min = htab.x_min; st->x_width = htab.x_width; st->inv_width = 1.0 / htab.x_width;
        st->log_valued = htab.log_valued;
        if (split_table) {
            /* this rank's entries, packed, to its slot of the symmetric heap; the gather kernel assembles y[] */
            for (int q = 0, i = slab.rank; i < N_DENS_INTERP; i += slab.P, q++) st->y[q] = htab.y[i];
            h2d_async(tab_sym + (size_t)k * tab_per, st->y, tab_per * sizeof(float));
            h2d_async(d_tables.p + k, st, offsetof(DevTable, y));
            dist_barrier_gather_interleaved32(reinterpret_cast<const unsigned int *>(tab_sym + (size_t)k * tab_per),
                                              reinterpret_cast<unsigned int *>(d_tables.p[k].y), N_DENS_INTERP);
        } else {
            memcpy(st->y, htab.y, sizeof(st->y));
            h2d_async(d_tables.p + k, st, sizeof(DevTable));
        }

        /* the reference leaves the last processed radius' f_coll in unnormalised_nion; every
           other radius only needs the grid sum and the ionised flags, so its f_coll grid is never
           materialised */
        const bool last = (k == n_todo - 1);
        float *fc = last ? (io.nion ? io.nion : d_fcoll.p) : (general ? d_fcoll.p : nullptr);
        if (last && io.nion && io.nion_written) *io.nion_written = true;
        const float *filtered = reinterpret_cast<const float *>(work[j % NW]);
        if (last && use_spec && j >= 3) {
            /* every speculative radius has been resolved once the stream has drained: a failed prediction
               must be known before the last radius turns the mask into the outputs */
            SpecState hs;
            d2h(&hs, d_spec, sizeof(SpecState));
            g_stats.d2h -= (long long)sizeof(SpecState);
            if (verbose) {
                fprintf(stderr, "[21cmfast_b200] ionize speculation: mean fix / queued cells per radius:");
                for (int q = 0; q < j; q++) fprintf(stderr, " %.6f/%u", hs.mean_fix[q], hs.qcount[q]);
                fprintf(stderr, " (queue: %d segments of %u)\n", sum_blocks, qcap);
            }
            if (sl) {
                /* a queue segment can overflow on one rank only: the ranks must leave the ladder together, or their
                   barrier sequences part ways */
                dev_zero(fail_sym, sizeof(unsigned long long));
                d2d(fail_sym, reinterpret_cast<const char *>(d_spec.p) + offsetof(SpecState, failed), sizeof(int));
                dist_barrier_gather(fail_sym, d_fail_all.p, 1);
                unsigned long long all[DIST_MAX_RANKS];
                d2h(all, d_fail_all.p, sizeof(unsigned long long) * (size_t)slab.P);
                g_stats.d2h -= (long long)(sizeof(unsigned long long) * (size_t)slab.P);
                for (int r = 0; r < slab.P; r++) hs.failed |= (all[r] != 0);
            }
            if (hs.failed) { restart = true; break; }
        }
        const bool spec_j = use_spec && j >= 2 && !last;
        SweepArgs sa = {nxl, ny, nz, plan->pitch, filtered, d_tables.p + k, d_partial, fc, chunk_rows,
                        d_spec.p, j, c.ion_eff_factor, d_mask, d_queue.p, qcap, d_qcounts.p};
        sweep_smem_optin();
        {
            auto ks = spec_j ? (htab.log_valued ? &fcoll_sum_kernel<true, true> : &fcoll_sum_kernel<false, true>)
                             : (htab.log_valued ? &fcoll_sum_kernel<true, false> : &fcoll_sum_kernel<false, false>);
            B200_LAUNCH_T(spec_j ? "fcoll_sum_classify_kernel" : "fcoll_sum_kernel", ks, sum_blocks, 256, SWEEP_REP_BYTES, sa);
        }
        {
            PlaneSumArgs ps = {nxl, chunks_per_plane * SWEEP_PARTIALS, d_partial, sl ? plane_sym + (size_t)k * nxl : d_plane.p};
            B200_LAUNCH(plane_sum_kernel, (nxl + 7) / 8, 256, 0, ps); /* a warp per plane (a thread per plane in the emulation) */
            if (sl) dist_barrier_gather(reinterpret_cast<const unsigned long long *>(plane_sym + (size_t)k * nxl),
                                        reinterpret_cast<unsigned long long *>(d_plane.p), nxl);
        }

        if (!last && !general) {
            CritDeltaArgs cd;
            memset(&cd, 0, sizeof(cd));
            cd.nx = nxl; cd.ny = ny; cd.nz = nz; cd.nzc = plan->pitch;
            cd.filtered = filtered; cd.table = d_tables.p + k;
            cd.partial = d_plane; cd.n_partial = nx; cd.mask = d_mask;
            if (sphere) { dev_zero(d_centre, (size_t)N); cd.mask = d_centre; }
            cd.n_cells = (double)N; cd.mean_f_coll = box->mean_f_coll; cd.f_limit = f_limit;
            cd.ion_eff_factor = c.ion_eff_factor; cd.mass_dep_zeta = c.mass_dep_zeta ? 1 : 0;
            if (use_spec) { cd.spec = d_spec.p; cd.j = j; cd.queue = d_queue.p; cd.qcounts = d_qcounts.p; cd.qcap = qcap; cd.n_seg = sum_blocks; }
            if (spec_j) { /* the queued cells; an overflowing segment raises `failed` like a missed bracket (an empty
                             fallback launch per radius cost 0.7 ms per step for a case that has not been observed) */
                B200_LAUNCH(spec_resolve_kernel, dev_num_sms() * 2, 256, SWEEP_REP_BYTES, cd);
            } else if (htab.log_valued) {
                B200_LAUNCH(ionise_delta_kernel<true>, sweep_blocks, 256, SWEEP_REP_BYTES, cd);
            } else {
                B200_LAUNCH(ionise_delta_kernel<false>, sweep_blocks, 256, SWEEP_REP_BYTES, cd);
            }
            if (sphere) paint_spheres(rs.R);
        } else {
            if (io.wait_slot >= 0) main_wait_copy_event(io.wait_slot);
            CritArgs ca;
            memset(&ca, 0, sizeof(ca));
            ca.n = NL; ca.fcoll = fc; ca.partial = d_plane; ca.n_partial = nx;
            ca.density = io.density; ca.prev_zre = io.prev_zre;
            ca.mask = d_mask; ca.xH = io.xH; ca.z_reion = io.z_reion; ca.Tk = io.Tk;
            ca.n_cells = (double)N; ca.mean_f_coll = box->mean_f_coll; ca.f_limit = f_limit;
            ca.ion_eff_factor = c.ion_eff_factor; ca.mass_dep_zeta = c.mass_dep_zeta ? 1 : 0;
            ca.R_index = rs.R_index; ca.redshift = c.redshift;
            ca.TK_nofluct = c.TK_nofluct; ca.adia_TK_term = c.adia_TK_term; ca.T_re = c.T_re;
            if (general) {
                ca.filtered = filtered; ca.nz = nz; ca.nzc = plan->pitch;
                if (ts) { ca.xe_grid = reinterpret_cast<const float *>(work_xe[j % NW]); ca.Tk_neutral = io.Tk_neutral; }
            }
            if (recomb) {
                ca.recomb = rec_grid_filtered ? 1 : (recomb == 2 && io.prev_rec) ? 2 : 3;
                ca.rec_grid = rec_grid_filtered ? reinterpret_cast<const float *>(work_rec[j % NW]) : io.prev_rec;
                ca.rec_scalar = recomb == 1 ? io.prev_rec_scalar : 0.;
                ca.G12 = io.G12; ca.mfp = io.mfp;
                ca.R = rs.R; ca.gamma_prefactor = c.gamma_prefactor;
            }
            const bool dilate_last = sphere && rs.R_index != 0; /* at R_index 0 the sphere is the centre cell itself */
            if (sphere) ca.paint = d_paint;
            if (dilate_last) { dev_zero(d_centre, (size_t)N); ca.mask = d_centre; }
            /* plain ladder: the last radius and the finalisation are one pass (ionise_last_fused_kernel) */
            fused_last = last && !general && !sphere && pt.phase < 0 && (NL & 3) == 0 &&
                         !(getenv("B200_FUSED_LAST") && getenv("B200_FUSED_LAST")[0] == '0');
            if (fused_last) {
                const float zf = (float)c.stored_redshift, Tref = (float)c.T_re;
                LastArgs la;
                la.c = ca;
                la.f = FinalArgs{NL, d_mask, io.density, io.prev_zre, io.xH, io.z_reion, io.Tk, d_flag,
                                 c.redshift, c.stored_redshift, c.T_re, c.TK_nofluct, c.adia_TK_term,
                                 pow((double)Tref, 1.7), pow(1e4 * ((1. + zf) / 4.), 1.7), nullptr, nullptr};
                B200_LAUNCH(ionise_last_fused_kernel, grid_for(NL, 1024), 256, 0, la);
            } else {
                B200_LAUNCH(ionise_kernel, grid_for(NL, 1024), 256, 0, ca);
            }
            if (dilate_last) paint_spheres(rs.R);
        }
    }
    if (!restart) break;
    }

    if (verbose)
        fprintf(stderr, "[21cmfast_b200] ionize host: enqueue %.3f ms, event wait %.3f ms, tables %.3f ms (%d radii)\n",
                1e3 * t_launch, 1e3 * t_wait, 1e3 * t_table, n_mine);
    if (pt.phase == 0) {
        dev_sync(); /* drain the stream before the work boxes are released */
        return;
    }
    {
        if (io.wait_slot >= 0) main_wait_copy_event(io.wait_slot);
        const float zf = (float)c.stored_redshift, Tref = (float)c.T_re;
        FinalArgs fa = {NL, d_mask, io.density, io.prev_zre, io.xH, io.z_reion, io.Tk, d_flag,
                        c.redshift, c.stored_redshift, c.T_re, c.TK_nofluct, c.adia_TK_term,
                        pow((double)Tref, 1.7), pow(1e4 * ((1. + zf) / 4.), 1.7), ts ? io.Tk_neutral : nullptr,
                        sphere ? d_paint.p : nullptr};
        if (!fused_last) B200_LAUNCH(finalize_kernel, grid_for(NL, 1024), 256, 0, fa);
        if (sl) dist_barrier(); /* no rank leaves (and reuses the symmetric heap) before every rank is done */
        int flag = 0;
        d2h(&flag, d_flag, sizeof(int)); /* also drains the stream before the work boxes are released */
        g_stats.d2h -= (long long)sizeof(int);
        if (sl) dist_check();
        if (flag) b200_throw(B200_InfinityorNaNError, "Tk after full ionisation is infinite or NaN");
    }
    recomb_update(recomb, io, c, N, d_flag);
    if (recomb && io.rec_written) *io.rec_written = true;
}

/* The radius ladder on Lagrangian source grids (SOURCE_MODEL = L-INTEGRAL; the `lagrangian_source_grids`
   branches of IonisationBox.c:587-593,615-621,636-642,819-835,1054-1066,1126-1132,1425-1430,1482-1488,1623-1628).
   The HaloBox already holds the ionising photons emitted per cell, so no per-radius table, extrema or mean fix
   exist: every radius is filter + c2r of the density (window HII_FILTER) and of the photon grid (the
   exponential mean-free-path window with USE_EXP_FILTER) and one criterion sweep, all enqueued without a
   host round trip. */
static void ionize_core_lagrangian(float redshift_f, float prev_redshift_f, const IonDeviceIO &io, IonizedBox *box) {
    const SimulationOptions *so = simulation_options_global;
    const AstroOptions *ao = astro_options_global;
    const MatterOptions *mo = matter_options_global;
    if (ao->USE_MINI_HALOS || ao->PHOTON_CONS_TYPE != 0 || ao->USE_TS_FLUCT || ao->IONISE_ENTIRE_SPHERE)
        b200_throw(B200_ValueError, "L-INTEGRAL ladder: mini-halos, photon conservation, USE_TS_FLUCT and IONISE_ENTIRE_SPHERE are not built");
    if (!io.halo_nion) b200_throw(B200_ValueError, "SOURCE_MODEL = L-INTEGRAL needs a computed HaloBox (n_ion)");
    const int recomb = ao->RECOMB_MODEL;
    const bool filter_rec = recomb != 0 && !ao->CELL_RECOMB;
    if (recomb) {
        if (recomb == 1 && filter_rec)
            b200_throw(B200_ValueError, "RECOMB_MODEL=homogeneous needs CELL_RECOMB (there is no N_rec grid to filter)");
        if (!io.G12 || (recomb == 2 && !io.cum_rec) || !io.halo_wsfr)
            b200_throw(B200_ValueError, "RECOMB_MODEL != none needs ionisation_rate_G12, cumulative_recombinations and the HaloBox's whalo_sfr");
        if (!recomb_tables()) b200_throw(B200_TableEvaluationError, "RECOMB_MODEL != none needs init_MHR()");
    }
    if (mo->USE_INTERPOLATION_TABLES != 2)
        b200_throw(B200_ValueError, "this build needs USE_INTERPOLATION_TABLES='hmf-interpolation'");
    const double redshift = redshift_f, prev_redshift = prev_redshift_f;
    IonConsts c;
    set_ionbox_constants(redshift, prev_redshift, &c);
    const int nx = so->HII_DIM, ny = so->HII_DIM, nz = hii_d_para();
    const long long N = (long long)nx * ny * nz;
    Fft3D *plan = fft_plan(nx, ny, nz);
    std::vector<RadiusSpec> radii = setup_radii(c);
    const int n_radii = (int)radii.size();

    box->log10_Mturnover_ave = io.log10_Mcrit_ACG_ave;
    box->log10_Mturnover_MINI_ave = io.log10_Mcrit_MCG_ave;
    const double Mturn_avg = pow(10., io.log10_Mcrit_ACG_ave);
    if (ao->INTEGRATION_METHOD_ATOMIC == INTEG_GL) initialise_GL(c.lnMmin, c.lnMmax_gl);
    box->mean_f_coll = Nion_General(redshift, c.lnMmin, c.lnMmax_gl, Mturn_avg, &c.sc);
    const double f_limit = Nion_General(so->Z_HEAT_MAX, c.lnMmin, c.lnMmax_gl, Mturn_avg, &c.sc);
    box->mean_f_coll_MINI = 0.;
    if (!std::isfinite(box->mean_f_coll) || box->mean_f_coll < 0)
        b200_throw(B200_InfinityorNaNError, "Mean collapse fraction is invalid");
    if (io.wait_slot >= 0) main_wait_copy_event(io.wait_slot);
    if (box->mean_f_coll * c.ion_eff_factor_gl < HII_ROUND_ERR) {
        { FillArgs f = {N, io.z_reion, -1.0f}; B200_LAUNCH(fill_kernel, grid_for(N, 1024), 256, 0, f); }
        NeutralArgs na = {N, io.density, io.xH, io.Tk, (float)(1. - xion_RECFAST(redshift)), c.TK_nofluct, c.adia_TK_term, nullptr, nullptr};
        B200_LAUNCH(neutral_box_kernel, grid_for(N, 1024), 256, 0, na);
        return;
    }
    std::vector<int> todo;
    for (int R_ct = n_radii; R_ct--;) {
        if (c.M_min > RtoM(radii[R_ct].R)) break;
        todo.push_back(R_ct);
    }
    const int n_todo = (int)todo.size();
    const bool complete = n_todo > 0 && todo.back() == 0; /* the loop reached R_index 0 */

    const size_t kn = plan->n_cplx();
    DevBuf<float2> k_dens(kn), k_stars(kn), k_sfr(recomb ? kn : 0), k_nrec;
    DevBuf<float2> w_dens(kn), w_stars(kn), w_sfr(recomb ? kn : 0), w_nrec;
    DevBuf<unsigned char> d_mask((size_t)N);
    dev_zero(d_mask, (size_t)N);
    DevBuf<int> d_flag(1);
    dev_zero(d_flag, sizeof(int));
    /* prepare_box_for_filtering (IonisationBox.c:1478-1488,1515-1518) */
    ZPrologue pro;
    pro.src = io.density; pro.src_row_stride = nz; pro.premul = 1.f;
    pro.clip = 1; pro.clip_lo = -1.f; pro.clip_hi = 1e6f;
    pro.post_scale = 1.f / (float)N;
    fft_r2c(plan, k_dens, pro);
    ZPrologue ps = pro;
    ps.src = io.halo_nion; ps.clip_lo = 0.f; ps.clip_hi = 1e20f;
    fft_r2c(plan, k_stars, ps);
    if (recomb) { ZPrologue pf = ps; pf.src = io.halo_wsfr; fft_r2c(plan, k_sfr, pf); }
    const bool rec_grid_filtered = filter_rec && io.prev_rec;
    if (rec_grid_filtered) {
        k_nrec.alloc(kn); w_nrec.alloc(kn);
        ZPrologue pr = ps; pr.src = io.prev_rec;
        fft_r2c(plan, k_nrec, pr);
    }
    const double dk0 = 2.0 * M_PI / so->BOX_LEN, dkz = 2.0 * M_PI / (so->BOX_LEN * so->NON_CUBIC_FACTOR);
    const int filter_hf = ao->USE_EXP_FILTER ? 3 : c.hii_filter;
    /* top-hat / gaussian windows over |n|^2 as in the Eulerian ladder (cubic boxes): one table per radius, shared by
       every grid that is filtered with HII_FILTER; the exponential filter keeps the per-mode double evaluation */
    const bool cubic = nx == ny && ny == nz && so->NON_CUBIC_FACTOR == 1.0f;
    const bool use_wtab = cubic && (c.hii_filter == 0 || c.hii_filter == 2);
    const int wtab_n = use_wtab ? window_table_size(plan) : 0;
    const size_t wtab3_n = use_wtab ? window_table3_size(plan) : 0;
    const bool use_wtab3 = use_wtab && (nx & (nx - 1)) == 0 && nx >= 16 && wtab3_n * sizeof(float) <= ((size_t)1 << 30);
    DevBuf<float> d_wtab((size_t)wtab_n), d_wtab3(use_wtab3 ? wtab3_n : 0);
    for (int k = 0; k < n_todo; k++) {
        const RadiusSpec &rs = radii[todo[k]];
        KMul km, kh; /* density / N_rec window, halo-field window */
        if (rs.R_index > 0) {
            km.kind = KMUL_FILTER; km.filter_type = c.hii_filter; km.R = (float)rs.R;
            km.dk[0] = dk0; km.dk[1] = dk0; km.dk[2] = dkz;
            kh = km; /* the exact window unless it is the density's own */
            if (use_wtab) { /* stream-ordered reuse of the one slot: the passes of the previous radius are enqueued before */
                window_table_build(plan, c.hii_filter, km.R, dk0, d_wtab);
                km.fast = 1; km.wtab = d_wtab; km.wtab_n = wtab_n;
                if (use_wtab3) { window_table_expand(plan, d_wtab, d_wtab3); km.wtab3 = d_wtab3; }
                if (filter_hf == c.hii_filter) kh = km;
            }
            kh.filter_type = filter_hf;
            if (filter_hf == 3) { /* filter_box(.., 3, R, mfp, 0.) with float arguments (filtering.c:308,330) */
                kh.R_param = (double)(float)c.mfp_meandens;
                const float q = -(float)rs.R / (float)c.mfp_meandens; /* the quotient of two floats: the window's
                                                                        cancellation amplifies its rounding to 1e-4 */
                kh.r_const = exp((double)q);
            }
        }
        ZEpilogue e0; /* the clips of calculate_fcoll_grid are applied where the sweeps read the grids */
        e0.scale = 1.f;
        fft_c2r(plan, k_dens, w_dens, km, e0);
        fft_c2r(plan, k_stars, w_stars, kh, e0);
        if (recomb) fft_c2r(plan, k_sfr, w_sfr, kh, e0);
        if (rec_grid_filtered) fft_c2r(plan, k_nrec, w_nrec, km, e0);
        CritArgs ca;
        memset(&ca, 0, sizeof(ca));
        ca.n = N; ca.density = io.density; ca.prev_zre = io.prev_zre;
        ca.mask = d_mask; ca.xH = io.xH; ca.z_reion = io.z_reion; ca.Tk = io.Tk;
        ca.f_limit = f_limit; ca.ion_eff_factor = 1.; ca.mass_dep_zeta = 1;
        ca.R_index = rs.R_index; ca.redshift = c.redshift;
        ca.TK_nofluct = c.TK_nofluct; ca.adia_TK_term = c.adia_TK_term; ca.T_re = c.T_re;
        ca.filtered = reinterpret_cast<const float *>(w_dens.p); ca.nz = nz; ca.nzc = plan->pitch;
        ca.stars_grid = reinterpret_cast<const float *>(w_stars.p);
        ca.rho_baryon = rho_crit() * cosmo_params_global->OMb;
        if (recomb) {
            ca.recomb = rec_grid_filtered ? 1 : (recomb == 2 && io.prev_rec) ? 2 : 3;
            ca.rec_grid = rec_grid_filtered ? reinterpret_cast<const float *>(w_nrec.p) : io.prev_rec;
            ca.rec_scalar = recomb == 1 ? io.prev_rec_scalar : 0.;
            ca.sfr_grid = reinterpret_cast<const float *>(w_sfr.p);
            ca.G12 = io.G12; ca.mfp = io.mfp;
            ca.R = rs.R; ca.gamma_prefactor = c.gamma_prefactor;
        }
        B200_LAUNCH(ionise_lagrangian_kernel, grid_for(N, 1024), 256, 0, ca);
    }
    /* the output's mean_f_coll is the grid mean of the last radius (no mean fix to report), floored like the
       reference's (IonisationBox.c:1566-1570,1623-1628); a ladder cut short leaves the global value */
    if (complete) {
        const int nb = grid_for((long long)nx * ny, 1);
        DevBuf<double> d_p((size_t)nb);
        RowSumArgs sa = {(long long)nx * ny, nz, plan->pitch, reinterpret_cast<const float *>(w_stars.p), d_p};
        B200_LAUNCH(clipped_sum_kernel, nb, 256, 0, sa);
        std::vector<double> hp((size_t)nb);
        d2h(hp.data(), d_p, hp.size() * sizeof(double));
        g_stats.d2h -= (long long)(hp.size() * sizeof(double));
        double sum = 0.;
        for (int i = 0; i < nb; i++) sum += hp[i];
        double grid_mean = sum / (double)N;
        if (grid_mean <= f_limit) grid_mean = f_limit;
        box->mean_f_coll = grid_mean;
    }
    {
        const float zf = (float)c.stored_redshift, Tref = (float)c.T_re;
        FinalArgs fa = {N, d_mask, io.density, io.prev_zre, io.xH, io.z_reion, io.Tk, d_flag,
                        c.redshift, c.stored_redshift, c.T_re, c.TK_nofluct, c.adia_TK_term,
                        pow((double)Tref, 1.7), pow(1e4 * ((1. + zf) / 4.), 1.7), nullptr, nullptr};
        B200_LAUNCH(finalize_kernel, grid_for(N, 1024), 256, 0, fa);
        int flag = 0;
        d2h(&flag, d_flag, sizeof(int));
        g_stats.d2h -= (long long)sizeof(int);
        if (flag) b200_throw(B200_InfinityorNaNError, "Tk after full ionisation is infinite or NaN");
    }
    recomb_update(recomb, io, c, N, d_flag);
    if (recomb && io.rec_written) *io.rec_written = true;
}

static void reset_stats() { g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0; g_stats.ms = 0; }

/* fills a host array with one value on a few helper threads while the caller goes on (the radius ladder
   leaves the host cores idle); joined when the object leaves scope, on the error path as well.  The
   single-threaded loop it replaces cost 55 ms per call at 512^3 -- more than the whole ladder. */
struct HostFill {
    std::vector<std::thread> workers;
    void start(float *p, long long n, float value, int nthreads = 4) {
        for (int t = 0; t < nthreads; t++) {
            const long long lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            workers.emplace_back([=] { std::fill(p + lo, p + hi, value); });
        }
    }
    ~HostFill() {
        for (auto &w : workers) w.join();
    }
};

extern "C" int ComputeIonizedBox(float redshift, float prev_redshift, PerturbedField *perturbed_field,
                                 PerturbedField *previous_perturbed_field, IonizedBox *previous_ionize_box,
                                 TsBox *spin_temp, HaloBox *halos, InitialConditions *ini_boxes, IonizedBox *box) {
    (void)previous_perturbed_field; (void)ini_boxes;
    try {
        require_params(true);
        rt_init();
        reset_stats();
        DevTimer timer;
        timer.start();
        const SimulationOptions *so = simulation_options_global;
        const long long N = (long long)so->HII_DIM * so->HII_DIM * hii_d_para();
        if (!perturbed_field || !perturbed_field->density || !box || !box->neutral_fraction || !box->z_reion)
            b200_throw(B200_ValueError, "ComputeIonizedBox: required arrays are NULL");
        const bool lagrangian = matter_options_global->SOURCE_MODEL == SRC_L_INTEGRAL;
        if (lagrangian && (!halos || !halos->n_ion))
            b200_throw(B200_ValueError, "ComputeIonizedBox: SOURCE_MODEL = L-INTEGRAL needs a computed HaloBox");


        /* first snapshot: the reference writes z_reion = -1 into the *previous* box
           (setup_first_z_prevbox, IonisationBox.c:365-386) */
        const bool first = prev_redshift < 1;
        HostFill prev_fill;
        if (first && previous_ionize_box && previous_ionize_box->z_reion)
            prev_fill.start(previous_ionize_box->z_reion, N, -1.0f);
        const bool ts = astro_options_global->USE_TS_FLUCT;
        if (ts && (!spin_temp || !spin_temp->xray_ionised_fraction || !spin_temp->kinetic_temp_neutral))
            b200_throw(B200_ValueError, "ComputeIonizedBox: USE_TS_FLUCT needs a computed TsBox");

        DevBuf<float> d_density_own, d_xH(N), d_zre(N), d_prev, d_Tk, d_nion;
        /* the perturbed density this box is computed from: the copy ComputePerturbedField left on the device
           (opt-in residency, rt.h) or an upload of the caller's array */
        const float *d_density = resident_get(perturbed_field->density, (size_t)N);
        if (!d_density) {
            d_density_own.alloc(N);
            h2d(d_density_own, perturbed_field->density, N * sizeof(float));
            d_density = d_density_own;
        }
        /* neutral_fraction / kinetic_temperature / previous z_reion are first read at the last
           radius: their upload rides on the copy stream behind the radius ladder */
        const int slot = 60;
        copy_wait_main();
        h2d_copy_stream(d_xH, box->neutral_fraction, N * sizeof(float));
        const bool want_Tk = !matter_options_global->MINIMIZE_MEMORY && box->kinetic_temperature;
        if (want_Tk) { d_Tk.alloc(N); h2d_copy_stream(d_Tk, box->kinetic_temperature, N * sizeof(float)); }
        if (box->unnormalised_nion && !lagrangian) d_nion.alloc(N);
        if (!first && previous_ionize_box && previous_ionize_box->z_reion) {
            d_prev.alloc(N);
            h2d_copy_stream(d_prev, previous_ionize_box->z_reion, N * sizeof(float));
        }
        /* recombinations: Gamma12 / mean free path keep the caller's values where no cell crosses the
           barrier; the previous snapshot's cumulative recombinations are an input */
        const int recomb = astro_options_global->RECOMB_MODEL;
        DevBuf<float> d_G12, d_mfp, d_cum, d_prev_rec;
        double cum_scalar = 0.;
        bool rec_written = false;
        if (recomb) {
            if (!box->ionisation_rate_G12 || !box->cumulative_recombinations || !previous_ionize_box ||
                !previous_ionize_box->cumulative_recombinations)
                b200_throw(B200_ValueError, "ComputeIonizedBox: RECOMB_MODEL != none needs ionisation_rate_G12 and the "
                                            "cumulative_recombinations of this and the previous box");
            d_G12.alloc(N);
            h2d_copy_stream(d_G12, box->ionisation_rate_G12, N * sizeof(float));
            if (!matter_options_global->MINIMIZE_MEMORY && box->mean_free_path) {
                d_mfp.alloc(N);
                h2d_copy_stream(d_mfp, box->mean_free_path, N * sizeof(float));
            }
            if (recomb == 2) {
                d_cum.alloc(N);
                d_prev_rec.alloc(N);
                h2d_copy_stream(d_prev_rec, previous_ionize_box->cumulative_recombinations, N * sizeof(float));
            }
        }
        DevBuf<float> d_xe, d_Tkn;
        if (ts) {
            d_xe.alloc(N); d_Tkn.alloc(N);
            h2d_copy_stream(d_xe, spin_temp->xray_ionised_fraction, N * sizeof(float));
            h2d_copy_stream(d_Tkn, spin_temp->kinetic_temp_neutral, N * sizeof(float));
        }
        copy_event_record(slot);
        bool nion_written = false;
        IonDeviceIO io = {d_density, d_prev.p, d_xH, d_zre, d_Tk.p, d_nion.p, slot, &nion_written};
        io.xe = d_xe.p; io.Tk_neutral = d_Tkn.p;
        if (recomb) {
            io.prev_rec = d_prev_rec.p;
            io.prev_rec_scalar = recomb == 1 ? (double)previous_ionize_box->cumulative_recombinations[0] : 0.;
            io.G12 = d_G12; io.mfp = d_mfp.p; io.cum_rec = d_cum.p;
            io.cum_rec_scalar_out = &cum_scalar; io.rec_written = &rec_written;
        }
        DevBuf<float> d_hnion, d_hwsfr;
        if (lagrangian) { /* HaloBox grids: read by the first transforms of the ladder */
            d_hnion.alloc(N);
            h2d(d_hnion, halos->n_ion, N * sizeof(float));
            io.halo_nion = d_hnion;
            if (recomb) {
                if (!halos->whalo_sfr) b200_throw(B200_ValueError, "ComputeIonizedBox: RECOMB_MODEL != none needs the HaloBox's whalo_sfr");
                d_hwsfr.alloc(N);
                h2d(d_hwsfr, halos->whalo_sfr, N * sizeof(float));
                io.halo_wsfr = d_hwsfr;
            }
            io.log10_Mcrit_ACG_ave = halos->log10_Mcrit_ACG_ave;
            io.log10_Mcrit_MCG_ave = halos->log10_Mcrit_MCG_ave;
            ionize_core_lagrangian(redshift, prev_redshift, io, box);
        } else {
            ionize_core(redshift, prev_redshift, io, box);
        }
        if (rec_written) {
            d2h(box->ionisation_rate_G12, d_G12, N * sizeof(float));
            if (d_mfp.p) d2h(box->mean_free_path, d_mfp, N * sizeof(float));
            if (recomb == 2) d2h(box->cumulative_recombinations, d_cum, N * sizeof(float));
            else box->cumulative_recombinations[0] = (float)cum_scalar;
        }

        d2h(box->neutral_fraction, d_xH, N * sizeof(float));
        d2h(box->z_reion, d_zre, N * sizeof(float));
        if (want_Tk) d2h(box->kinetic_temperature, d_Tk, N * sizeof(float));
        if (d_nion.p && nion_written) d2h(box->unnormalised_nion, d_nion, N * sizeof(float));
        if (resident_enabled()) { /* ComputeBrightnessTemp reads the neutral fraction next */
            resident_put(box->neutral_fraction, d_xH.p, (size_t)N);
            d_xH.p = nullptr; d_xH.n = 0;
        }
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        try { copy_stream_sync(); } catch (B200Error &) {} /* no copy may outlive the buffers released above */
        if (getenv("B200_VERBOSE") || e.code == B200_CUDAError)
            fprintf(stderr, "[21cmfast_b200] ComputeIonizedBox: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

/* Device-resident variant (bench.py `value` leg): d_* structs hold DEVICE pointers. */
extern "C" int b200_ComputeIonizedBox_device(float redshift, float prev_redshift, PerturbedField *d_pf, IonizedBox *d_box) {
    try {
        require_params(true);
        rt_init();
        reset_stats();
        DevTimer timer;
        timer.start();
        IonDeviceIO io = {d_pf->density, nullptr, d_box->neutral_fraction, d_box->z_reion, d_box->kinetic_temperature,
                          d_box->unnormalised_nion};
        ionize_core(redshift, prev_redshift, io, d_box);
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ComputeIonizedBox_device: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

/* Radius-parallel variant of the device-resident entry point (see IonPartition): d_mask is a device
   buffer of HII_DIM^2 * HII_D_PARA bytes owned by the caller, who all-reduces it (MAX) between the
   phase-0 and the phase-1 call. */
extern "C" int b200_ComputeIonizedBox_device_part(float redshift, float prev_redshift, PerturbedField *d_pf,
                                                  IonizedBox *d_box, unsigned char *d_mask, int part, int nparts,
                                                  int phase) {
    try {
        require_params(true);
        rt_init();
        reset_stats();
        if (!d_mask || nparts < 1 || part < 0 || part >= nparts || (phase != 0 && phase != 1))
            b200_throw(B200_ValueError, "b200_ComputeIonizedBox_device_part: bad partition arguments");
        DevTimer timer;
        timer.start();
        IonDeviceIO io = {d_pf->density, nullptr, d_box->neutral_fraction, d_box->z_reion, d_box->kinetic_temperature,
                          d_box->unnormalised_nion};
        IonPartition pt;
        pt.part = part; pt.nparts = nparts; pt.phase = phase; pt.mask = d_mask;
        ionize_core(redshift, prev_redshift, io, d_box, pt);
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ComputeIonizedBox_device_part: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

/* Slab-decomposed variant ("slab FFT"): ONE box over the ranks connected with
   b200_dist_init / b200_dist_connect.  The structs hold DEVICE pointers to this rank's x-slab
   [HII_DIM / P][HII_DIM][HII_D_PARA] of every array; the result is this rank's slab of the outputs,
   bit-identical to the same planes of the single-GPU box. */
extern "C" int b200_ComputeIonizedBox_slab(float redshift, float prev_redshift, PerturbedField *d_pf, IonizedBox *d_box) {
    try {
        require_params(true);
        rt_init();
        reset_stats();
        dist_require();
        DevTimer timer;
        timer.start();
        IonDeviceIO io = {d_pf->density, nullptr, d_box->neutral_fraction, d_box->z_reion, d_box->kinetic_temperature,
                          d_box->unnormalised_nion};
        IonPartition pt;
        pt.slab = true;
        ionize_core(redshift, prev_redshift, io, d_box, pt);
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ComputeIonizedBox_slab: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

/* Host-only part of ComputeIonizedBox (no device work): lets the CPU test tier check the host
   scalar chain of the shipped library against the golden fixtures.
   out = [mean_f_coll, f_limit, M_min, sigma_minmass, TK_nofluct, growth, ion_eff_factor, n_radii,
          R_0, sigma_max_0, R_1, sigma_max_1, ...] truncated to n entries. */
extern "C" int b200_ionize_host_scalars(float redshift, double *out, int n) {
    try {
        require_params(true);
        IonConsts c;
        set_ionbox_constants(redshift, -1.0, &c);
        std::vector<RadiusSpec> radii = setup_radii(c);
        double mean_f_coll, f_limit;
        if (c.mass_dep_zeta) {
            mean_f_coll = Nion_General(redshift, c.lnMmin, c.lnMmax_gl, astro_params_global->M_TURN, &c.sc);
            f_limit = Nion_General(simulation_options_global->Z_HEAT_MAX, c.lnMmin, c.lnMmax_gl,
                                   astro_params_global->M_TURN, &c.sc);
        } else {
            mean_f_coll = Fcoll_General(redshift, c.lnMmin, c.lnMmax_gl);
            f_limit = Fcoll_General(simulation_options_global->Z_HEAT_MAX, c.lnMmin, c.lnMmax_gl);
        }
        std::vector<double> v = {mean_f_coll, f_limit, c.M_min, c.sigma_minmass, c.TK_nofluct,
                                 c.growth_factor, c.ion_eff_factor, (double)radii.size()};
        for (const RadiusSpec &r : radii) { v.push_back(r.R); v.push_back(r.sigma_maxmass); }
        for (int i = 0; i < n; i++) out[i] = i < (int)v.size() ? v[i] : 0.0;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ionize_host_scalars: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
