/*
 * perturb.cu -- ComputePerturbedField: move the initial-condition mass with ZA / 2LPT, deposit it
 * with cloud-in-cell onto the low-resolution grid, clip, and derive the velocity field
 * (reference PerturbedField.c:24-496 and move_grid_masses, map_mass.c:23-60,146-208).
 *
 * Scope: LINEAR, ZELDOVICH and 2LPT algorithms; PERTURB_ON_HIGH_RES off (default: the optimised path
 * below) or on (perturb_core_hires).
 *
 * Deposit design: the reference adds 8 weighted contributions per hi-res particle into a double
 * grid with `omp atomic` (order, hence rounding, varies run to run).  Here every contribution is
 * converted to 2^-40 fixed point and accumulated in unsigned 64-bit integers: integer addition is
 * associative, so the deposit is bit-reproducible (and can be split over GPUs and merged by an
 * integer all-reduce, PerturbPartition), and its quantisation (4.5e-13 per contribution) is far
 * below the float the sum is finally rounded to.
 *   integer DIM / HII_DIM ratio F (the default 3): move_cic_grouped_kernel -- one thread per low-res
 *     velocity cell contracts its F^3 particles into a 3x3x3 block: 27 reductions instead of 8 F^3;
 *   otherwise: move_cic_kernel -- a CTA owns a brick of particles, accumulates into a shared-memory
 *     tile of the low-res grid (brick + halo) and flushes it once; contributions outside the tile go
 *     straight to global atomics, so correctness never depends on the displacement size.
 * Host interface: with a cold IC cache the hi-res density travels in x-slabs on a copy stream and
 * the deposit of slab k starts as soon as its copy has landed (PendingUpload).
 */
#include "fft.h"
#include "dist.h"
#include "host_physics.h"

#include <vector>

#define FIXED_SCALE 1099511627776.0 /* 2^40 */

/* ------------------------------------------------------------------ device-resident IC cache */
struct IcsCache {
    bool valid = false;
    const void *host[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long sig[7] = {0, 0, 0, 0, 0, 0, 0};
    int dim = 0, hii = 0, dpara = 0, hpara = 0;
    DevBuf<float> hires, v[3], v2[3];
};
static IcsCache g_ics;
void ics_cache_drop() {
    g_ics.valid = false;
    g_ics.hires.release();
    for (int a = 0; a < 3; a++) { g_ics.v[a].release(); g_ics.v2[a].release(); }
}
/* The cache is OPT-IN (b200_ics_cache(1) or B200_ICS_CACHE=1): the reference reads the caller's
   arrays on every call, so by default so does this library.  When enabled, the caller promises not
   to modify the IC arrays in place between calls (or to call b200_ics_cache_invalidate() after
   doing so): a hit is decided by the host pointers, the grid sizes and a signature of ~4096
   sampled words per array, which guards against a reused allocation, not against a sparse edit.
   A full-content hash would cost as much host time as the upload it saves (14.5 GB at DIM=1536). */
/* b200_ics_share: the ranks of a connected group (b200_dist_connect) call ComputePerturbedField TOGETHER, each
   with its own redshift but the SAME initial conditions in host memory (the redshifts of a coeval run or a
   lightcone spread over the GPUs).  Off by default. */
static bool g_ics_share = false;
extern "C" void b200_ics_share(int enable) { g_ics_share = enable != 0; }
static int g_ics_cache_on = -1; /* -1: follow the environment */
extern "C" void b200_ics_cache(int enable) {
    g_ics_cache_on = enable ? 1 : 0;
    if (!enable) ics_cache_drop();
}
extern "C" void b200_ics_cache_invalidate(void) { g_ics.valid = false; }
static bool ics_cache_enabled() {
    if (g_ics_cache_on >= 0) return g_ics_cache_on == 1;
    const char *ce = getenv("B200_ICS_CACHE");
    return ce && ce[0] == '1';
}
static unsigned long long sample_signature(const float *p, size_t n) {
    if (!p) return 0;
    unsigned long long h = 1469598103934665603ULL ^ n;
    const size_t step = n / 4096 ? n / 4096 : 1;
    for (size_t i = 0; i < n; i += step) {
        unsigned int w;
        memcpy(&w, p + i, 4);
        h = (h ^ w) * 1099511628211ULL;
    }
    return h;
}

/* ------------------------------------------------------------------ kernels */
struct MoveArgs {
    int dn[3];   /* hi-res (particle) grid */
    int vn[3];   /* velocity grid */
    int on[3];   /* output grid */
    const float *dens;
    const float *v[3];
    const float *v2[3];  /* null for Zel'dovich */
    unsigned long long *acc;  /* output accumulator, fixed point */
    double ratio_vel, ratio_out;
    double vdf[3], vdf2[3];
    double init_growth;
    int brick[3];   /* hi-res particles per CTA along each axis */
    int tile0[3];   /* low-res cells covered by a brick (without halo) */
    int halo;
    int tiles[3];   /* bricks per axis */
    /* slab-decomposed deposit (move_cic_grouped_kernel<.., SLAB = true>): this rank holds the velocity
       cells of x-planes [vel_x0, vel_x0 + vn[0]) (vn[0] = local planes), the hi-res planes starting at the
       (unwrapped) global plane dens_x0, and an accumulator window of out_nxl planes whose first one is the
       (unwrapped) global plane out_x0 = x0 - halo; mass leaving the window raises *overflow */
    int vel_x0, dens_x0, out_x0, out_nxl;
    int *overflow;
};

/* single periodic wrap without a branch, for indices known to lie in [-n, 2n) */
DEV int wrap_once(int i, int n) {
    i += (i < 0) ? n : 0;
    i -= (i >= n) ? n : 0;
    return i;
}
DEV int wrap_index(int i, int n) {
    while (i >= n) i -= n;
    while (i < 0) i += n;
    return i;
}

__global__ void __launch_bounds__(256) move_cic_kernel(MoveArgs a) {
    DYN_SMEM(unsigned long long, tile);
    const int tx = a.tile0[0] + 2 * a.halo, ty = a.tile0[1] + 2 * a.halo, tz = a.tile0[2] + 2 * a.halo;
    const int tcells = tx * ty * tz;
    for (int i = threadIdx.x; i < tcells; i += blockDim.x) tile[i] = 0ULL;
    __syncthreads();
    /* brick coordinates */
    const int bz = blockIdx.x % a.tiles[2];
    const int by = (blockIdx.x / a.tiles[2]) % a.tiles[1];
    const int bx = blockIdx.x / (a.tiles[2] * a.tiles[1]);
    const int i0 = bx * a.brick[0], j0 = by * a.brick[1], k0 = bz * a.brick[2];
    /* low-res origin of the tile: cell containing the brick origin, minus the halo */
    const int ox = (int)floor(i0 * a.ratio_out) - a.halo;
    const int oy = (int)floor(j0 * a.ratio_out) - a.halo;
    const int oz = (int)floor(k0 * a.ratio_out) - a.halo;
    const int np = a.brick[0] * a.brick[1] * a.brick[2];
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        const int k = k0 + p % a.brick[2];
        const int j = j0 + (p / a.brick[2]) % a.brick[1];
        const int i = i0 + p / (a.brick[2] * a.brick[1]);
        if (i >= a.dn[0] || j >= a.dn[1] || k >= a.dn[2]) continue;
        /* nearest velocity cell (resample_index + wrap_coord, indexing.h:110-114) */
        const int vi = wrap_index((int)(i * a.ratio_vel + 0.5), a.vn[0]);
        const int vj = wrap_index((int)(j * a.ratio_vel + 0.5), a.vn[1]);
        const int vk = wrap_index((int)(k * a.ratio_vel + 0.5), a.vn[2]);
        const long long vidx = (long long)vk + (long long)a.vn[2] * ((long long)vj + (long long)a.vn[1] * vi);
        double pos[3] = {(double)i, (double)j, (double)k};
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            pos[ax] += (double)a.v[ax][vidx] * a.vdf[ax];
            if (a.v2[0]) pos[ax] -= (double)a.v2[ax][vidx] * a.vdf2[ax];
            pos[ax] *= a.ratio_out;
        }
        const long long didx = (long long)k + (long long)a.dn[2] * ((long long)j + (long long)a.dn[1] * i);
        const double mass = 1.0 + (double)a.dens[didx] * a.init_growth;
        int ip[3];
        double d[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            ip[ax] = (int)floor(pos[ax]);
            d[ax] = pos[ax] - ip[ax];
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int cx = c & 1, cy = (c >> 1) & 1, cz = (c >> 2) & 1;
            const double w = (cx ? d[0] : 1. - d[0]) * (cy ? d[1] : 1. - d[1]) * (cz ? d[2] : 1. - d[2]);
            const long long q = llrint(mass * w * FIXED_SCALE);
            const int gx = ip[0] + cx, gy = ip[1] + cy, gz = ip[2] + cz;
            const int lx = gx - ox, ly = gy - oy, lz = gz - oz;
            if (lx >= 0 && lx < tx && ly >= 0 && ly < ty && lz >= 0 && lz < tz) {
                atomic_add_u64(&tile[(lx * ty + ly) * tz + lz], (unsigned long long)q);
            } else {
                const int wx = wrap_index(gx, a.on[0]), wy = wrap_index(gy, a.on[1]), wz = wrap_index(gz, a.on[2]);
                atomic_add_u64(&a.acc[(long long)wz + (long long)a.on[2] * ((long long)wy + (long long)a.on[1] * wx)],
                               (unsigned long long)q);
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tcells; t += blockDim.x) {
        const unsigned long long q = tile[t];
        if (q == 0ULL) continue;
        const int lz = t % tz, ly = (t / tz) % ty, lx = t / (tz * ty);
        const int wx = wrap_index(ox + lx, a.on[0]), wy = wrap_index(oy + ly, a.on[1]), wz = wrap_index(oz + lz, a.on[2]);
        atomic_add_u64(&a.acc[(long long)wz + (long long)a.on[2] * ((long long)wy + (long long)a.on[1] * wx)], q);
    }
}

/* Integer DIM/HII_DIM ratio F: the F^3 particles whose nearest velocity cell is the same low-res
   cell share one displacement, so along each axis their F positions span less than one output
   cell and touch at most 3 cells.  One thread owns such a group: it contracts the F^3 masses with
   the three per-axis F x 3 CIC weight matrices (z, then y, then x) into a 3x3x3 block and issues
   27 fixed-point adds instead of 8 F^3.  Same positions and weights as the generic kernel. */
struct GroupArgs {
    MoveArgs m;
    long long p_begin, p_end; /* range of groups (x-major flat index) this launch deposits */
    int gbrick[3];  /* source (low-res) cells per CTA */
    int gtiles[3];
};

/* per-axis CIC data of one group: position of sub-particle t -> (cell offset o_t in {0,1} relative
   to the base cell, weight w1_t of the upper cell).  The F x 3 weight matrix row is
   W[t][a] = (a == o_t) ? 1 - w1_t : (a == o_t + 1) ? w1_t : 0. */
template <int F, typename T> DEV void axis_cic(int c, double disp, double r, int &base, T (&w1)[F], int (&o)[F]) {
    const int istart = (int)ceil((double)F * c - 0.5 * F);
    int b0 = 0;
#pragma unroll
    for (int t = 0; t < F; t++) {
        double pos = (double)(istart + t);
        pos += disp;
        pos *= r;
        const double fl = floor(pos);
        const int b = (int)fl;
        if (t == 0) b0 = b;
        o[t] = b - b0;
        w1[t] = (T)(pos - fl);
    }
    base = b0;
}
template <typename T> DEV T cic_w(int a, int o, T w1) { return a == o ? (T)1 - w1 : (a == o + 1 ? w1 : (T)0); }
DEV long long to_fixed(double v) { return llrint(v * FIXED_SCALE); }
DEV long long to_fixed(float v) { return llrintf(v * (float)FIXED_SCALE); }

/* T = arithmetic of the F^3 -> 3x3x3 contraction.  Positions, cell indices and the CIC weights
   are always derived in double (the weight is a difference of O(DIM) numbers); with T = float
   (default) the weights and the masses 1 + delta D(z_i) are then rounded to float and contracted
   in single precision: each of the 27 sums carries ~1e-7 relative error, which averages to
   < 6e-8 of the cell total (itself rounded to float32 afterwards), at a third of the instruction
   cost of the double contraction (B200_CIC_DOUBLE=1 selects T = double). */
template <int F, typename T, bool SLAB> __global__ void __launch_bounds__(128, 4) move_cic_grouped_kernel(GroupArgs g) {
    const MoveArgs &a = g.m;
    const int nzg = a.vn[2], nyg = a.vn[1], nxg = a.vn[0];
    const long long ngroups = (long long)nxg * nyg * nzg;
    const long long out_sx = (long long)a.on[1] * a.on[2];
    const T growth = (T)a.init_growth;
    (void)ngroups;
    for (long long p = g.p_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < g.p_end;
         p += (long long)gridDim.x * blockDim.x) {
        const int cz = (int)(p % nzg);
        const int cy = (int)((p / nzg) % nyg);
        const int cx = (int)(p / ((long long)nzg * nyg)) + (SLAB ? a.vel_x0 : 0); /* global x of the velocity cell */
        double disp[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            disp[ax] = (double)a.v[ax][p] * a.vdf[ax];
            if (a.v2[0]) disp[ax] -= (double)a.v2[ax][p] * a.vdf2[ax];
        }
        T w1x[F], w1y[F], w1z[F];
        int ox_[F], oy_[F], oz_[F];
        int Bx, By, Bz;
        axis_cic<F, T>(cx, disp[0], a.ratio_out, Bx, w1x, ox_);
        axis_cic<F, T>(cy, disp[1], a.ratio_out, By, w1y, oy_);
        axis_cic<F, T>(cz, disp[2], a.ratio_out, Bz, w1z, oz_);
        const int isx = (int)ceil((double)F * cx - 0.5 * F), isy = (int)ceil((double)F * cy - 0.5 * F),
                  isz = (int)ceil((double)F * cz - 0.5 * F);
        /* all F^3 densities first: ncu's source view showed 36 % of the stall samples on the first
           use of each row's loads when they were issued row by row (nine exposed memory latencies
           per group); issued together they overlap */
        float dens[F][F][F];
        {
            /* the group's cells start at most F/2 below 0 and end at most F/2 above the grid: one
               branch-free wrap per index keeps the 27 loads unconditional, so they can be issued
               back to back */
            int hx[F], hy[F], hz[F];
#pragma unroll
            for (int t = 0; t < F; t++) {
                hx[t] = SLAB ? isx + t - a.dens_x0 : wrap_once(isx + t, a.dn[0]); /* slab: index into the rank's own planes */
                hy[t] = wrap_once(isy + t, a.dn[1]);
                hz[t] = wrap_once(isz + t, a.dn[2]);
            }
#pragma unroll
            for (int t0 = 0; t0 < F; t0++)
#pragma unroll
                for (int t1 = 0; t1 < F; t1++) {
                    const float *row = a.dens + (long long)a.dn[2] * ((long long)hy[t1] + (long long)a.dn[1] * hx[t0]);
#pragma unroll
                    for (int t2 = 0; t2 < F; t2++) dens[t0][t1][t2] = ldg(&row[hz[t2]]);
                }
        }
        /* contract z then y for each x-slice of the group: Cy[t0][b][c] */
        T Cy[F][3][3];
#pragma unroll
        for (int t0 = 0; t0 < F; t0++) {
#pragma unroll
            for (int i = 0; i < 9; i++) (&Cy[t0][0][0])[i] = (T)0;
#pragma unroll
            for (int t1 = 0; t1 < F; t1++) {
                T Bzv[3] = {(T)0, (T)0, (T)0};
#pragma unroll
                for (int t2 = 0; t2 < F; t2++) {
                    const T mass = (T)1 + (T)dens[t0][t1][t2] * growth;
#pragma unroll
                    for (int c = 0; c < 3; c++) Bzv[c] += mass * cic_w<T>(c, oz_[t2], w1z[t2]);
                }
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const T wy = cic_w<T>(b, oy_[t1], w1y[t1]);
#pragma unroll
                    for (int c = 0; c < 3; c++) Cy[t0][b][c] += wy * Bzv[c];
                }
            }
        }
        /* x contraction one output plane at a time (rolled loop keeps the kernel inside the
           instruction cache), 9 fixed-point 64-bit reductions per plane straight into L2 */
        const bool inside = (SLAB || (Bx >= 0 && Bx + 2 < a.on[0])) && By >= 0 && Bz >= 0 && By + 2 < a.on[1] && Bz + 2 < a.on[2];
#pragma unroll 1
        for (int aa = 0; aa < 3; aa++) {
            T A[3][3];
#pragma unroll
            for (int i = 0; i < 9; i++) (&A[0][0])[i] = (T)0;
#pragma unroll
            for (int t0 = 0; t0 < F; t0++) {
                const T wx = cic_w<T>(aa, ox_[t0], w1x[t0]);
#pragma unroll
                for (int i = 0; i < 9; i++) (&A[0][0])[i] += wx * (&Cy[t0][0][0])[i];
            }
            int gx;
            if (SLAB) { /* plane of the rank's accumulator window (interior + halo), no wrap */
                gx = Bx + aa - a.out_x0;
                if (gx < 0 || gx >= a.out_nxl) { *a.overflow = 1; continue; }
            } else {
                gx = inside ? Bx + aa : wrap_index(Bx + aa, a.on[0]);
            }
            unsigned long long *plane = a.acc + (long long)gx * out_sx;
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const int gy = inside ? By + b : wrap_index(By + b, a.on[1]);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const long long q = to_fixed(A[b][c]);
                    if (q == 0) continue;
                    const int gz = inside ? Bz + c : wrap_index(Bz + c, a.on[2]);
                    atomic_add_u64(&plane[(long long)gy * a.on[2] + gz], (unsigned long long)q);
                }
            }
        }
    }
}

template <int F> static void launch_grouped(const MoveArgs &a, long long p_begin, long long p_end) {
    GroupArgs g;
    g.m = a;
    g.p_begin = p_begin; g.p_end = p_end;
    for (int ax = 0; ax < 3; ax++) { g.gbrick[ax] = 0; g.gtiles[ax] = 0; }
    const long long ngroups = p_end - p_begin;
    if (ngroups <= 0) return;
    long long blocks = (ngroups + 127) / 128;
    const long long cap = (long long)dev_num_sms() * 64;
    if (blocks > cap) blocks = cap;
    static int dbl = -1;
    if (dbl < 0) { const char *e = getenv("B200_CIC_DOUBLE"); dbl = (e && e[0] == '1') ? 1 : 0; }
    if (a.overflow) { /* slab-decomposed deposit */
        if (dbl) {
            auto kp = &move_cic_grouped_kernel<F, double, true>;
            B200_LAUNCH_T("move_cic_grouped_kernel", kp, (int)blocks, 128, 0, g);
        } else {
            auto kp = &move_cic_grouped_kernel<F, float, true>;
            B200_LAUNCH_T("move_cic_grouped_kernel", kp, (int)blocks, 128, 0, g);
        }
    } else if (dbl) {
        auto kp = &move_cic_grouped_kernel<F, double, false>;
        B200_LAUNCH_T("move_cic_grouped_kernel", kp, (int)blocks, 128, 0, g);
    } else {
        auto kp = &move_cic_grouped_kernel<F, float, false>;
        B200_LAUNCH_T("move_cic_grouped_kernel", kp, (int)blocks, 128, 0, g);
    }
}

/* largest |x displacement| of the rank's velocity cells, in output cells, as an order key */
struct MaxDispArgs {
    long long n;
    const float *vx, *vx2;
    double vdf, vdf2, ratio_out;
    int *keys; /* [1] = max key */
};
__global__ void max_disp_kernel(MaxDispArgs a) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        double d = (double)a.vx[i] * a.vdf;
        if (a.vx2) d -= (double)a.vx2[i] * a.vdf2;
        const float f = (float)fabs(d * a.ratio_out);
        m = f > m ? f : m;
    }
    atomic_max_i32(&a.keys[1], float_order_key(float_as_int_bits(m)));
}

struct AccSlabArgs {
    int nxl, halo;
    long long plane; /* ny * nz */
    int ny, nz, nzc;
    const unsigned long long *acc, *left, *right; /* own window, and the windows of the ranks below / above (peer memory) */
    float *padded;
    double mass_factor;
};
/* halo exchange + normalise_delta_grid: the mass that this rank's neighbours deposited into their halo
   planes is pulled over NVLink and added to the rank's own planes -- integer sums, so the total is
   bit-identical to the single-GPU accumulator */
__global__ void acc_to_delta_slab_kernel(AccSlabArgs a) {
    const long long nrows = (long long)a.nxl * a.ny;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int xl = (int)(row / a.ny);
        const long long yz0 = (row - (long long)xl * a.ny) * a.nz;
        for (int z = threadIdx.x; z < a.nz; z += blockDim.x) {
            unsigned long long q = a.acc[(long long)(a.halo + xl) * a.plane + yz0 + z];
            if (xl < a.halo) q += a.left[(long long)(a.nxl + a.halo + xl) * a.plane + yz0 + z];
            if (xl >= a.nxl - a.halo) q += a.right[(long long)(xl - (a.nxl - a.halo)) * a.plane + yz0 + z];
            const double m = (double)(long long)q * (1.0 / FIXED_SCALE);
            float v = (float)m;
            v = (float)((double)v * a.mass_factor);
            v = v - 1.0f;
            a.padded[row * 2 * a.nzc + z] = v;
        }
    }
}

struct AccToDeltaArgs {
    long long nrows;
    int nz, nzc;
    const unsigned long long *acc;
    float *padded;
    double mass_factor;
    int to_delta; /* 1: (m * mass_factor) - 1 (normalise_delta_grid); 0: plain double -> float copy */
};
/* double -> float copy into the FFT layout (PerturbedField.c:115-128) + normalise_delta_grid (:180-210) */
__global__ void acc_to_delta_kernel(AccToDeltaArgs a) {
    for (long long row = blockIdx.x; row < a.nrows; row += gridDim.x)
        for (int z = threadIdx.x; z < a.nz; z += blockDim.x) {
            const double m = (double)(long long)a.acc[row * a.nz + z] * (1.0 / FIXED_SCALE);
            float v = (float)m;
            if (a.to_delta) {
                v = (float)((double)v * a.mass_factor);
                v = v - 1.0f;
            }
            a.padded[row * 2 * a.nzc + z] = v;
        }
}

struct LinearArgs {
    long long nrows;
    int nz, nzc;
    const float *dens;
    float *padded;
    double growth;
};
/* PERTURB_ALGORITHM = LINEAR (PerturbedField.c:66-82) */
__global__ void linear_density_kernel(LinearArgs a) {
    for (long long row = blockIdx.x; row < a.nrows; row += gridDim.x)
        for (int z = threadIdx.x; z < a.nz; z += blockDim.x)
            a.padded[row * 2 * a.nzc + z] = (float)(a.growth * (double)a.dens[row * a.nz + z]);
}

/* ------------------------------------------------------------------ orchestration */
struct ResampleArgs {
    int ln[3], hn[3], h_pitch, l_pitch;
    double ratio;
    const float *hi_padded;
    float *lo;          /* padded (l_pitch > 0) or unpadded (l_pitch == 0) low-res destination */
    float scale, offset;
};
/* nearest-cell sub-sampling hi -> lo through resample_index (indexing.h:110-114):
   lo = hi * scale + offset */
__global__ void resample_kernel(ResampleArgs a) {
    const long long n = (long long)a.ln[0] * a.ln[1] * a.ln[2];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % a.ln[2]), j = (int)((t / a.ln[2]) % a.ln[1]), i = (int)(t / ((long long)a.ln[2] * a.ln[1]));
        const int hi = (int)(i * a.ratio + 0.5), hj = (int)(j * a.ratio + 0.5), hk = (int)(k * a.ratio + 0.5);
        float v = a.hi_padded[(long long)hk + 2LL * a.h_pitch * ((long long)hj + (long long)a.hn[1] * hi)] * a.scale;
        v = v + a.offset;
        const long long dst = a.l_pitch ? (long long)k + 2LL * a.l_pitch * ((long long)j + (long long)a.ln[1] * i) : t;
        a.lo[dst] = v;
    }
}

/* host arrays still to be uploaded when perturb_core starts: the deposit then runs slab by slab
   behind the copies (ComputePerturbedField with a cold IC cache) */
struct PendingUpload {
    const float *h_hires = nullptr;          /* host DIM^3 */
    const float *h_v[3] = {nullptr, nullptr, nullptr}, *h_v2[3] = {nullptr, nullptr, nullptr}; /* host HII^3 */
    float *d_hires = nullptr, *d_v[3] = {nullptr, nullptr, nullptr}, *d_v2[3] = {nullptr, nullptr, nullptr};
};
static void upload_all(const PendingUpload &u, long long M, long long N) {
    h2d(u.d_hires, u.h_hires, M * sizeof(float));
    for (int a = 0; a < 3; a++) {
        if (u.h_v[a]) h2d(u.d_v[a], u.h_v[a], N * sizeof(float));
        if (u.h_v2[a]) h2d(u.d_v2[a], u.h_v2[a], N * sizeof(float));
    }
}

struct PerturbDeviceIO {
    const float *hires_density; /* device, DIM^3 (2LPT / ZA) */
    const float *lowres_density; /* device, HII^3 (LINEAR only) */
    const float *v[3], *v2[3];   /* device low-res velocity boxes */
    float *density, *vel[3];     /* device outputs (vel[a] may be null) */
};

/* PERTURB_ON_HIGH_RES = True (PerturbedField.c:24-178, 284-387): the particles are deposited on the
   hi-res grid with the hi-res velocity boxes, the evolved field is low-pass filtered (real-space
   top-hat at the low-res cell scale) and sub-sampled, and the velocities are derived from the
   unfiltered hi-res k-space field, filtered and sub-sampled.  io.v / io.v2 are DIM^3 boxes here. */
static void perturb_core_hires(float redshift_f, const PerturbDeviceIO &io) {
    const SimulationOptions *so = simulation_options_global;
    const MatterOptions *mo = matter_options_global;
    const double redshift = redshift_f;
    const int hn[3] = {so->HII_DIM, so->HII_DIM, hii_d_para()};
    const int dn[3] = {so->DIM, so->DIM, d_para()};
    const long long N = (long long)hn[0] * hn[1] * hn[2];
    const long long M = (long long)dn[0] * dn[1] * dn[2];
    Fft3D *plan = fft_plan(hn[0], hn[1], hn[2]);
    Fft3D *hplan = fft_plan(dn[0], dn[1], dn[2]);
    DevBuf<float2> kbox(plan->n_cplx()), work(plan->n_cplx());
    DevBuf<float2> hk(hplan->n_cplx()), hsaved(hplan->n_cplx()), hwork(hplan->n_cplx());
    float *padded = reinterpret_cast<float *>(kbox.p);
    float *hpadded = reinterpret_cast<float *>(hk.p);
    const int hrow_blocks = (int)((long long)dn[0] * dn[1] < 4096 ? (long long)dn[0] * dn[1] : 4096);
    const int flat_blocks = dev_num_sms() * 8;
    const double growth = dicke(redshift);
    const double dk0 = 2.0 * M_PI / so->BOX_LEN, dkz = 2.0 * M_PI / (so->BOX_LEN * so->NON_CUBIC_FACTOR);

    if (mo->PERTURB_ALGORITHM == PERTURB_LINEAR) {
        LinearArgs la = {(long long)dn[0] * dn[1], dn[2], hplan->pitch, io.hires_density, hpadded, growth};
        B200_LAUNCH(linear_density_kernel, hrow_blocks, 256, 0, la);
    } else {
        MoveArgs a;
        memset(&a, 0, sizeof(a));
        const double box_size[3] = {so->BOX_LEN, so->BOX_LEN, so->BOX_LEN * so->NON_CUBIC_FACTOR};
        const double init_growth = dicke(so->INITIAL_REDSHIFT);
        const double d2 = -(3.0 / 7.0) * growth * growth, d2i = -(3.0 / 7.0) * init_growth * init_growth;
        for (int ax = 0; ax < 3; ax++) {
            a.dn[ax] = dn[ax]; a.vn[ax] = dn[ax]; a.on[ax] = dn[ax];
            a.v[ax] = io.v[ax];
            a.v2[ax] = (mo->PERTURB_ALGORITHM == PERTURB_2LPT) ? io.v2[ax] : nullptr;
            a.vdf[ax] = (growth - init_growth) / box_size[ax] * dn[ax];
            a.vdf2[ax] = (d2 - d2i) / box_size[ax] * dn[ax];
        }
        a.dens = io.hires_density;
        a.ratio_vel = 1.0; a.ratio_out = 1.0;
        a.init_growth = init_growth;
        DevBuf<unsigned long long> acc(M);
        dev_zero(acc, M * sizeof(unsigned long long));
        a.acc = acc;
        launch_grouped<1>(a, 0, M); /* every particle is its own group */
        AccToDeltaArgs ca = {(long long)dn[0] * dn[1], dn[2], hplan->pitch, acc, hpadded, 1.0, 0};
        B200_LAUNCH(acc_to_delta_kernel, hrow_blocks, 256, 0, ca);
    }

    /* assign_to_lowres_grid (PerturbedField.c:137-178) */
    ZPrologue pro;
    fft_r2c(hplan, hk, pro);
    d2d(hsaved, hk, hplan->n_cplx() * sizeof(float2));
    KMul lowpass;
    lowpass.kind = KMUL_FILTER; lowpass.filter_type = 0;
    lowpass.R = (float)(pc::l_factor * so->BOX_LEN / (hn[0] + 0.0));
    lowpass.dk[0] = dk0; lowpass.dk[1] = dk0; lowpass.dk[2] = dkz;
    ZEpilogue plain;
    fft_c2r(hplan, hk, hwork, lowpass, plain);
    {
        /* normalise_delta_grid with mass_factor = 1 (:180-210) rides on the sub-sampling */
        const bool to_delta = mo->PERTURB_ALGORITHM != PERTURB_LINEAR;
        ResampleArgs ra = {{hn[0], hn[1], hn[2]}, {dn[0], dn[1], dn[2]}, hplan->pitch, plan->pitch,
                           dn[0] / (double)hn[0], reinterpret_cast<const float *>(hwork.p), padded,
                           1.0f / (float)M, 0.f};
        B200_LAUNCH(resample_kernel, flat_blocks, 256, 0, ra);
        (void)to_delta;
    }
    if (mo->PERTURB_ALGORITHM != PERTURB_LINEAR) {
        /* "*cell *= 1.0; *cell -= 1": a second, in-place pass keeps the reference's rounding sequence */
        ResampleArgs rb = {{hn[0], hn[1], hn[2]}, {hn[0], hn[1], hn[2]}, plan->pitch, plan->pitch, 1.0,
                           padded, padded, 1.0f, -1.0f};
        B200_LAUNCH(resample_kernel, flat_blocks, 256, 0, rb);
    }

    /* smooth_and_clip_density (:212-282); the saved k-space box is the hi-res one */
    fft_r2c(plan, kbox, pro);
    if (mo->SMOOTH_EVOLVED_DENSITY_FIELD) {
        KMul ks;
        ks.dk[0] = dk0; ks.dk[1] = dk0; ks.dk[2] = dkz;
        ks.kind = KMUL_FILTER; ks.filter_type = 2;
        ks.R = (float)(so->DENSITY_SMOOTH_RADIUS * so->BOX_LEN / (float)so->HII_DIM);
        fft_apply_window(plan, kbox, ks);
    }
    ZEpilogue epi;
    epi.scale = 1.f / (float)N;
    epi.clip = 1; epi.clip_lo = (float)(-1.0 + pc::FRACT_FLOAT_ERR); epi.clip_hi = 3.0e38f;
    epi.dst = io.density; epi.dst_row_stride = hn[2];
    fft_c2r(plan, kbox, work, KMul(), epi);

    /* compute_perturbed_velocities on the hi-res grid (:284-387) */
    if (so->HII_DIM > 1) {
        const double dDdt_over_D = ddickedt(redshift) / dicke(redshift);
        for (int ax = 0; ax < 3; ax++) {
            if (!io.vel[ax]) continue;
            KMul kv;
            kv.dk[0] = dk0; kv.dk[1] = dk0; kv.dk[2] = dkz;
            kv.op = KOP_VELOCITY_F; kv.axis_a = ax; kv.op_factor = dDdt_over_D / (double)M;
            if (so->DIM != so->HII_DIM) { kv.kind = KMUL_FILTER; kv.filter_type = 0; kv.R = lowpass.R; }
            fft_c2r(hplan, hsaved, hwork, kv, plain);
            ResampleArgs rv = {{hn[0], hn[1], hn[2]}, {dn[0], dn[1], dn[2]}, hplan->pitch, 0,
                               dn[0] / (double)hn[0], reinterpret_cast<const float *>(hwork.p), io.vel[ax], 1.0f, 0.f};
            B200_LAUNCH(resample_kernel, flat_blocks, 256, 0, rv);
        }
    }
    dev_sync(); /* the hi-res work boxes are released on return */
}

/* Slab-parallel deposit of ONE box on several GPUs ("PerturbField move+CIC"):
   phase 0: this rank deposits the groups of its x-slab [part, part+1) * HII_DIM / nparts into the
   caller's fixed-point accumulator `acc` (N x u64, zeroed here) and returns; the caller sums the
   accumulators over the ranks (all-reduce SUM on int64 -- integer addition, so the result is
   bit-identical to the single-GPU deposit whatever the number of ranks); phase 1: every rank turns
   the merged accumulator into delta and runs the (small) FFT chain.  phase -1: everything. */
struct PerturbPartition {
    int part = 0, nparts = 1, phase = -1;
    unsigned long long *acc = nullptr;
};

static void perturb_core(float redshift_f, const PerturbDeviceIO &io, const PendingUpload *pending = nullptr,
                         const PerturbPartition &pt = PerturbPartition()) {
    const SimulationOptions *so = simulation_options_global;
    const MatterOptions *mo = matter_options_global;
    if (mo->PERTURB_ON_HIGH_RES) {
        if (pt.phase >= 0) b200_throw(B200_ValueError, "the slab-parallel deposit is not built for PERTURB_ON_HIGH_RES");
        if (pending) {
            const long long Mh = (long long)so->DIM * so->DIM * d_para();
            upload_all(*pending, Mh, Mh); /* density and velocity boxes are all hi-res here */
        }
        perturb_core_hires(redshift_f, io);
        return;
    }
    const double redshift = redshift_f;
    const int hn[3] = {so->HII_DIM, so->HII_DIM, hii_d_para()};
    const int dn[3] = {so->DIM, so->DIM, d_para()};
    const long long N = (long long)hn[0] * hn[1] * hn[2];
    const long long M = (long long)dn[0] * dn[1] * dn[2];
    Fft3D *plan = fft_plan(hn[0], hn[1], hn[2]);
    DevBuf<float2> kbox(plan->n_cplx()), work(plan->n_cplx());
    float *padded = reinterpret_cast<float *>(kbox.p);
    const int row_blocks = (int)((long long)hn[0] * hn[1] < 4096 ? (long long)hn[0] * hn[1] : 4096);

    const double growth = dicke(redshift);
    if (pt.phase >= 0 && mo->PERTURB_ALGORITHM == PERTURB_LINEAR)
        b200_throw(B200_ValueError, "the slab-parallel deposit needs PERTURB_ALGORITHM = ZELDOVICH or 2LPT");
    if (mo->PERTURB_ALGORITHM == PERTURB_LINEAR) {
        if (pending) h2d(pending->d_hires, pending->h_hires, N * sizeof(float));
        LinearArgs la = {(long long)hn[0] * hn[1], hn[2], plan->pitch, io.lowres_density, padded, growth};
        B200_LAUNCH(linear_density_kernel, row_blocks, 256, 0, la);
    } else {
        /* move_grid_masses, map_mass.c:146-208 */
        MoveArgs a;
        memset(&a, 0, sizeof(a));
        const double boxlen = so->BOX_LEN, boxlen_z = boxlen * so->NON_CUBIC_FACTOR;
        const double box_size[3] = {boxlen, boxlen, boxlen_z};
        const double init_growth = dicke(so->INITIAL_REDSHIFT);
        const double d2 = -(3.0 / 7.0) * growth * growth, d2i = -(3.0 / 7.0) * init_growth * init_growth;
        for (int ax = 0; ax < 3; ax++) {
            a.dn[ax] = dn[ax]; a.vn[ax] = hn[ax]; a.on[ax] = hn[ax];
            a.v[ax] = io.v[ax];
            a.v2[ax] = (mo->PERTURB_ALGORITHM == PERTURB_2LPT) ? io.v2[ax] : nullptr;
            a.vdf[ax] = (growth - init_growth) / box_size[ax] * dn[ax];
            a.vdf2[ax] = (d2 - d2i) / box_size[ax] * dn[ax];
        }
        a.dens = io.hires_density;
        a.ratio_vel = (double)hn[0] / (double)dn[0];
        a.ratio_out = (double)hn[0] / (double)dn[0];
        a.init_growth = init_growth;
        DevBuf<unsigned long long> own_acc;
        unsigned long long *acc = pt.acc;
        if (pt.phase < 0) { own_acc.alloc(N); acc = own_acc; }
        if (pt.phase <= 0) dev_zero(acc, N * sizeof(unsigned long long));
        a.acc = acc;
        /* bricks of ~8 low-res cells per axis (fewer if the grid is small), halo of 4 cells */
        a.halo = 4;
        for (int ax = 0; ax < 3; ax++) {
            int cells = hn[ax] < 8 ? hn[ax] : 8;
            a.brick[ax] = (int)ceil(cells / a.ratio_out);
            if (a.brick[ax] > dn[ax]) a.brick[ax] = dn[ax];
            a.tiles[ax] = (dn[ax] + a.brick[ax] - 1) / a.brick[ax];
            a.tile0[ax] = (int)ceil(a.brick[ax] * a.ratio_out) + 1;
        }
        /* integer hi/lo ratio (the default DIM = 3 HII_DIM): grouped kernel, else the generic one */
        const int F = dn[0] / hn[0];
        const bool integer_ratio = F * hn[0] == dn[0] && F * hn[1] == dn[1] && F * hn[2] == dn[2] && F >= 1 && F <= 4;
        const char *force = getenv("B200_CIC_GENERIC");
        auto deposit = [&](long long p0, long long p1) {
            if (F == 1) launch_grouped<1>(a, p0, p1);
            else if (F == 2) launch_grouped<2>(a, p0, p1);
            else if (F == 3) launch_grouped<3>(a, p0, p1);
            else launch_grouped<4>(a, p0, p1);
        };
        const long long plane_groups = (long long)hn[1] * hn[2];
        if (pt.phase >= 0 && !integer_ratio)
            b200_throw(B200_ValueError, "the slab-parallel deposit needs an integer DIM / HII_DIM ratio <= 4");
        if (pt.phase == 0) {
            /* this rank's x-slab of groups only; the stream is drained before the caller's collective */
            const int cx0 = (int)((long long)hn[0] * pt.part / pt.nparts), cx1 = (int)((long long)hn[0] * (pt.part + 1) / pt.nparts);
            deposit((long long)cx0 * plane_groups, (long long)cx1 * plane_groups);
            dev_sync();
            return;
        }
        if (pt.phase == 1) {
            /* merged accumulator supplied by the caller: nothing to deposit */
        } else if (integer_ratio && !(force && force[0] == '1')) {
            if (pending && hn[0] >= 8) {
                /* Pipeline the upload with the deposit: x-slabs of the hi-res density (and of the
                   low-res velocity boxes) travel on the copy stream; the groups of slab k are
                   deposited as soon as its copy has landed.  Group cx reads hi-res planes
                   ceil(F cx - F/2) .. + F - 1, i.e. it starts at most one plane below F cx: inside
                   slab k or the (earlier) slab k - 1.  Only cx = 0 reaches back to plane DIM - 1; it
                   is deposited last. */
                const int nslab = hn[0] >= 64 ? 16 : (hn[0] >= 16 ? 4 : 2);
                const long long hplane = (long long)dn[1] * dn[2];
                copy_wait_main(); /* the destination buffers may still be in use by earlier work */
                for (int k = 0; k < nslab; k++) {
                    const int cx0 = (int)((long long)hn[0] * k / nslab), cx1 = (int)((long long)hn[0] * (k + 1) / nslab);
                    const long long lo = (long long)cx0 * plane_groups, cnt = (long long)(cx1 - cx0) * plane_groups;
                    for (int ax = 0; ax < 3; ax++) {
                        if (pending->h_v[ax]) h2d_copy_stream(pending->d_v[ax] + lo, pending->h_v[ax] + lo, cnt * sizeof(float));
                        if (pending->h_v2[ax]) h2d_copy_stream(pending->d_v2[ax] + lo, pending->h_v2[ax] + lo, cnt * sizeof(float));
                    }
                    const long long hlo = (long long)F * cx0 * hplane, hcnt = (long long)F * (cx1 - cx0) * hplane;
                    h2d_copy_stream(pending->d_hires + hlo, pending->h_hires + hlo, hcnt * sizeof(float));
                    copy_event_record(k);
                    main_wait_copy_event(k);
                    deposit((k == 0 ? 1 : cx0) * plane_groups, (long long)cx1 * plane_groups);
                }
                deposit(0, plane_groups);
            } else {
                if (pending) upload_all(*pending, M, N);
                deposit(0, (long long)hn[0] * plane_groups);
            }
        } else {
            if (pending) upload_all(*pending, M, N);
            const size_t smem = sizeof(unsigned long long) * (size_t)(a.tile0[0] + 2 * a.halo) *
                                (a.tile0[1] + 2 * a.halo) * (a.tile0[2] + 2 * a.halo);
#ifndef B200_EMU
            if (smem > 48 * 1024)
                CUDA_CHECK(cudaFuncSetAttribute(move_cic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
            B200_LAUNCH(move_cic_kernel, a.tiles[0] * a.tiles[1] * a.tiles[2], 256, smem, a);
        }
        AccToDeltaArgs ca = {(long long)hn[0] * hn[1], hn[2], plan->pitch, acc, padded, (double)N / (double)M, 1};
        B200_LAUNCH(acc_to_delta_kernel, row_blocks, 256, 0, ca);
        /* acc returns to the pool at scope exit; reuse is stream-ordered (single stream) */
    }

    /* smooth_and_clip_density, PerturbedField.c:212-282 */
    ZPrologue pro;
    fft_r2c(plan, kbox, pro);
    KMul km;
    const double dk0 = 2.0 * M_PI / so->BOX_LEN, dkz = 2.0 * M_PI / (so->BOX_LEN * so->NON_CUBIC_FACTOR);
    km.dk[0] = dk0; km.dk[1] = dk0; km.dk[2] = dkz;
    if (mo->SMOOTH_EVOLVED_DENSITY_FIELD) {
        /* gaussian smoothing of the evolved field (PerturbedField.c:221-227); the smoothed k-space
           box is also what the velocities are derived from, so it is smoothed in place */
        KMul ks = km;
        ks.kind = KMUL_FILTER; ks.filter_type = 2;
        ks.R = (float)(so->DENSITY_SMOOTH_RADIUS * so->BOX_LEN / (float)so->HII_DIM);
        fft_apply_window(plan, kbox, ks);
    }
    ZEpilogue epi;
    epi.scale = 1.f / (float)N;
    epi.clip = 1; epi.clip_lo = (float)(-1.0 + pc::FRACT_FLOAT_ERR); epi.clip_hi = 3.0e38f;
    epi.dst = io.density; epi.dst_row_stride = hn[2];
    fft_c2r(plan, kbox, work, KMul(), epi);

    /* compute_perturbed_velocities, PerturbedField.c:284-387 */
    if (so->HII_DIM > 1) {
        const double dDdt_over_D = ddickedt(redshift) / dicke(redshift);
        for (int ax = 0; ax < 3; ax++) {
            if (!io.vel[ax]) continue;
            KMul kv = km;
            kv.op = KOP_VELOCITY_F; kv.axis_a = ax; kv.op_factor = dDdt_over_D / (double)N;
            ZEpilogue ev;
            ev.dst = io.vel[ax]; ev.dst_row_stride = hn[2];
            fft_c2r(plan, kbox, work, kv, ev);
        }
    }
}

/* ONE box over the ranks of dist.h, every stage on x-slabs (SURVEY.md section 8e, rows "PerturbField
   move+CIC" and "slab FFT"): the rank deposits the particles of its own velocity cells into an
   accumulator window of its planes plus `halo` planes on either side (the halo is sized from the largest
   x displacement of the box), pulls its neighbours' halo planes over NVLink while normalising
   (acc_to_delta_slab_kernel), and runs the density / velocity transforms slab-decomposed.  The IC
   arrays of io are the rank's slabs: velocity boxes [nxl][ny][nz], hi-res density F nxl planes starting
   at global plane F x0 - F/2 (periodic).  Bit-identical to perturb_core on the whole box. */
static void perturb_core_slab(float redshift_f, const PerturbDeviceIO &io) {
    const SimulationOptions *so = simulation_options_global;
    const MatterOptions *mo = matter_options_global;
    dist_require();
    if (mo->PERTURB_ON_HIGH_RES)
        b200_throw(B200_ValueError, "the slab-decomposed perturbed field works on the low-res grid (PERTURB_ON_HIGH_RES = False)");
    const bool linear = mo->PERTURB_ALGORITHM == PERTURB_LINEAR;
    const double redshift = redshift_f;
    const int hn[3] = {so->HII_DIM, so->HII_DIM, hii_d_para()};
    const int dn[3] = {so->DIM, so->DIM, d_para()};
    const long long N = (long long)hn[0] * hn[1] * hn[2];
    const long long M = (long long)dn[0] * dn[1] * dn[2];
    const int F = dn[0] / hn[0];
    if (!(F * hn[0] == dn[0] && F * hn[1] == dn[1] && F * hn[2] == dn[2] && F >= 1 && F <= 4))
        b200_throw(B200_ValueError, "the slab-decomposed deposit needs an integer DIM / HII_DIM ratio <= 4");
    Fft3D *plan = fft_plan(hn[0], hn[1], hn[2]);
    dist_reset();
    FftSlab slab = fft_slab_setup(plan);
    const int nxl = slab.nxl, x0 = slab.x0, P = slab.P;
    const long long plane = (long long)hn[1] * hn[2];
    DevBuf<float2> kT(slab.n_cplx()), work(slab.n_cplx());
    float *padded = reinterpret_cast<float *>(work.p);

    const double growth = dicke(redshift);
    DevBuf<int> d_over(1);
    dev_zero(d_over, sizeof(int));
    int halo = 0;
    if (linear) { /* PerturbedField.c:66-82: the linear field of the rank's own planes, nothing moves */
        if (!io.lowres_density) b200_throw(B200_ValueError, "slab perturb (LINEAR): lowres_density slab is NULL");
        const int row_blocks = (int)((long long)nxl * hn[1] < 4096 ? (long long)nxl * hn[1] : 4096);
        LinearArgs la = {(long long)nxl * hn[1], hn[2], plan->pitch, io.lowres_density, padded, growth};
        B200_LAUNCH(linear_density_kernel, row_blocks, 256, 0, la);
    } else {
        MoveArgs a;
        memset(&a, 0, sizeof(a));
        const double boxlen = so->BOX_LEN, boxlen_z = boxlen * so->NON_CUBIC_FACTOR;
        const double box_size[3] = {boxlen, boxlen, boxlen_z};
        const double init_growth = dicke(so->INITIAL_REDSHIFT);
        const double d2 = -(3.0 / 7.0) * growth * growth, d2i = -(3.0 / 7.0) * init_growth * init_growth;
        for (int ax = 0; ax < 3; ax++) {
            a.dn[ax] = dn[ax]; a.vn[ax] = hn[ax]; a.on[ax] = hn[ax];
            a.v[ax] = io.v[ax];
            a.v2[ax] = (mo->PERTURB_ALGORITHM == PERTURB_2LPT) ? io.v2[ax] : nullptr;
            a.vdf[ax] = (growth - init_growth) / box_size[ax] * dn[ax];
            a.vdf2[ax] = (d2 - d2i) / box_size[ax] * dn[ax];
        }
        a.vn[0] = nxl; /* local velocity planes */
        a.dens = io.hires_density;
        a.ratio_vel = (double)hn[0] / (double)dn[0];
        a.ratio_out = (double)hn[0] / (double)dn[0];
        a.init_growth = init_growth;

        /* halo width from the largest x displacement of the whole box */
        int *keys_sym = (int *)dist_alloc(2 * sizeof(int));
        DevBuf<int> d_keys(2);
        {
            const int init[2] = {2147483647, -2147483647 - 1};
            h2d(keys_sym, init, sizeof(init));
            g_stats.h2d -= (long long)sizeof(init);
            MaxDispArgs ma = {(long long)nxl * plane, a.v[0], a.v2[0], a.vdf[0], a.vdf2[0], a.ratio_out, keys_sym};
            B200_LAUNCH(max_disp_kernel, dev_num_sms() * 4, 256, 0, ma);
            dist_barrier_minmax(keys_sym, d_keys);
        }
        int hkeys[2];
        d2h(hkeys, d_keys, sizeof(hkeys));
        g_stats.d2h -= (long long)sizeof(hkeys);
        const double max_disp = (double)float_from_order_key(hkeys[1]);
        halo = (int)ceil(max_disp) + 2;
        if (const char *e = getenv("B200_SLAB_HALO")) halo = atoi(e);
        if (halo < 1) halo = 1;
        if (halo > nxl)
            b200_throw(B200_ValueError, "slab deposit: displacements of %.1f cells exceed the slab thickness %d (use fewer ranks)", max_disp, nxl);

        a.vel_x0 = x0;
        a.dens_x0 = F * x0 - F / 2;
        a.out_x0 = x0 - halo;
        a.out_nxl = nxl + 2 * halo;
        a.overflow = d_over;
        const size_t acc_n = (size_t)a.out_nxl * plane;
        unsigned long long *acc = (unsigned long long *)dist_alloc(acc_n * sizeof(unsigned long long));
        dev_zero(acc, acc_n * sizeof(unsigned long long));
        a.acc = acc;
        const long long ngroups = (long long)nxl * plane;
        if (F == 1) launch_grouped<1>(a, 0, ngroups);
        else if (F == 2) launch_grouped<2>(a, 0, ngroups);
        else if (F == 3) launch_grouped<3>(a, 0, ngroups);
        else launch_grouped<4>(a, 0, ngroups);
        dist_barrier(); /* every rank's deposit has landed before the halos are pulled */
        {
            const int row_blocks = (int)((long long)nxl * hn[1] < 4096 ? (long long)nxl * hn[1] : 4096);
            AccSlabArgs ca = {nxl, halo, plane, hn[1], hn[2], plan->pitch, acc,
                              dist_peer(acc, (slab.rank + P - 1) % P), dist_peer(acc, (slab.rank + 1) % P), padded,
                              (double)N / (double)M};
            B200_LAUNCH(acc_to_delta_slab_kernel, row_blocks, 256, 0, ca);
        }

    }

    /* smooth_and_clip_density, PerturbedField.c:212-282 */
    ZPrologue pro;
    fft_r2c_slab(&slab, kT, work, pro);
    KMul km;
    const double dk0 = 2.0 * M_PI / so->BOX_LEN, dkz = 2.0 * M_PI / (so->BOX_LEN * so->NON_CUBIC_FACTOR);
    km.dk[0] = dk0; km.dk[1] = dk0; km.dk[2] = dkz;
    if (mo->SMOOTH_EVOLVED_DENSITY_FIELD) {
        KMul ks = km;
        ks.kind = KMUL_FILTER; ks.filter_type = 2;
        ks.R = (float)(so->DENSITY_SMOOTH_RADIUS * so->BOX_LEN / (float)so->HII_DIM);
        fft_apply_window(plan, kT, ks, slab.nyl, slab.y0);
    }
    ZEpilogue epi;
    epi.scale = 1.f / (float)N;
    epi.clip = 1; epi.clip_lo = (float)(-1.0 + pc::FRACT_FLOAT_ERR); epi.clip_hi = 3.0e38f;
    epi.dst = io.density; epi.dst_row_stride = hn[2];
    fft_c2r_slab(&slab, kT, work, KMul(), epi);

    /* compute_perturbed_velocities, PerturbedField.c:284-387 */
    if (so->HII_DIM > 1) {
        const double dDdt_over_D = ddickedt(redshift) / dicke(redshift);
        for (int ax = 0; ax < 3; ax++) {
            if (!io.vel[ax]) continue;
            KMul kv = km;
            kv.op = KOP_VELOCITY_F; kv.axis_a = ax; kv.op_factor = dDdt_over_D / (double)N;
            ZEpilogue ev;
            ev.dst = io.vel[ax]; ev.dst_row_stride = hn[2];
            fft_c2r_slab(&slab, kT, work, kv, ev);
        }
    }
    dist_barrier(); /* no rank leaves (and reuses the symmetric heap) before every rank is done */
    int over = 0;
    d2h(&over, d_over, sizeof(int));
    g_stats.d2h -= (long long)sizeof(int);
    dist_check();
    if (over) b200_throw(B200_ValueError, "slab deposit: mass left the halo of %d planes (B200_SLAB_HALO too small)", halo);
}

static void reset_stats() { g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0; g_stats.ms = 0; }

extern "C" int ComputePerturbedField(float redshift, InitialConditions *boxes, PerturbedField *pf) {
    try {
        require_params(false);
        rt_init();
        reset_stats();
        DevTimer timer;
        timer.start();
        const SimulationOptions *so = simulation_options_global;
        const MatterOptions *mo = matter_options_global;
        if (!boxes || !pf || !pf->density) b200_throw(B200_ValueError, "ComputePerturbedField: NULL struct/array");
        const long long N = (long long)so->HII_DIM * so->HII_DIM * hii_d_para();
        const long long M = (long long)so->DIM * so->DIM * d_para();
        const bool linear = mo->PERTURB_ALGORITHM == PERTURB_LINEAR;
        const bool lpt2 = mo->PERTURB_ALGORITHM == PERTURB_2LPT;
        const bool on_hires = mo->PERTURB_ON_HIGH_RES;
        /* PERTURB_ON_HIGH_RES: hi-res density also for LINEAR, hi-res velocity boxes (make_density_grid,
           PerturbedField.c:32-54) */
        const float *hv[7] = {(linear && !on_hires) ? boxes->lowres_density : boxes->hires_density,
                              on_hires ? boxes->hires_vx : boxes->lowres_vx, on_hires ? boxes->hires_vy : boxes->lowres_vy,
                              on_hires ? boxes->hires_vz : boxes->lowres_vz,
                              on_hires ? boxes->hires_vx_2LPT : boxes->lowres_vx_2LPT,
                              on_hires ? boxes->hires_vy_2LPT : boxes->lowres_vy_2LPT,
                              on_hires ? boxes->hires_vz_2LPT : boxes->lowres_vz_2LPT};
        const size_t NV = (size_t)(on_hires ? M : N);
        const size_t hn[7] = {(size_t)((linear && !on_hires) ? N : M), NV, NV, NV, NV, NV, NV};
        const int nuse = linear ? 1 : (lpt2 ? 7 : 4);
        for (int i = 0; i < nuse; i++)
            if (!hv[i]) b200_throw(B200_ValueError, "ComputePerturbedField: a required IC array is NULL");
        /* opt-in, several GPUs working on the SAME initial conditions (one redshift each): every rank uploads
           1 / world of every IC array over its own PCIe link and sends that share to the peers over NVLink */
        if (g_ics_share && g_dist.ready && g_dist.world > 1) {
            if (on_hires || linear) b200_throw(B200_ValueError, "shared IC upload: ZELDOVICH / 2LPT on the low-res grid only");
            const int P = g_dist.world, me = g_dist.rank;
            dist_reset();
            float *sym[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            for (int i = 0; i < nuse; i++) sym[i] = (float *)dist_alloc(hn[i] * sizeof(float));
            for (int i = 0; i < nuse; i++) {
                const size_t lo = hn[i] * (size_t)me / P, hi = hn[i] * (size_t)(me + 1) / P;
                h2d(sym[i] + lo, hv[i] + lo, (hi - lo) * sizeof(float));
                for (int r = 1; r < P; r++) { /* staggered: at any time every rank writes to a different peer */
                    const int peer = (me + r) % P;
                    d2d(dist_peer(sym[i], peer) + lo, sym[i] + lo, (hi - lo) * sizeof(float));
                }
            }
            dist_barrier(); /* every share of every array has landed everywhere */
            DevBuf<float> d_density(N), d_v[3];
            float *host_v[3] = {mo->KEEP_3D_VELOCITIES ? pf->velocity_x : nullptr,
                                mo->KEEP_3D_VELOCITIES ? pf->velocity_y : nullptr, pf->velocity_z};
            PerturbDeviceIO io;
            memset(&io, 0, sizeof(io));
            io.hires_density = sym[0]; io.lowres_density = sym[0];
            for (int a = 0; a < 3; a++) {
                io.v[a] = sym[1 + a]; io.v2[a] = lpt2 ? sym[4 + a] : nullptr;
                if (host_v[a]) { d_v[a].alloc(N); io.vel[a] = d_v[a]; }
            }
            io.density = d_density;
            perturb_core(redshift, io, nullptr);
            dist_barrier(); /* no rank reuses the heap (next call's dist_reset) while a peer still reads its ICs */
            d2h(pf->density, d_density, N * sizeof(float));
            for (int a = 0; a < 3; a++)
                if (host_v[a]) d2h(host_v[a], d_v[a], N * sizeof(float));
            dist_check();
            if (resident_enabled()) {
                resident_put(pf->density, d_density.p, (size_t)N);
                d_density.p = nullptr; d_density.n = 0;
            }
            g_stats.ms = timer.stop_ms();
            return 0;
        }
        /* opt-in: keep the initial conditions resident between calls (perturb_field is called once
           per redshift on the same ICs), see b200_ics_cache() above; default = upload every call */
        const bool use_cache = ics_cache_enabled();
        bool hit = use_cache && g_ics.valid && g_ics.dim == so->DIM && g_ics.hii == so->HII_DIM;
        unsigned long long sig[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < nuse && use_cache; i++) {
            sig[i] = sample_signature(hv[i], hn[i]);
            if (hit && (g_ics.host[i] != hv[i] || g_ics.sig[i] != sig[i])) hit = false;
        }
        PendingUpload pending;
        if (!hit) {
            ics_cache_drop();
            g_ics.hires.alloc(hn[0]);
            pending.h_hires = hv[0]; pending.d_hires = g_ics.hires;
            for (int a = 0; a < 3 && nuse > 1; a++) {
                g_ics.v[a].alloc(NV);
                pending.h_v[a] = hv[1 + a]; pending.d_v[a] = g_ics.v[a];
                if (lpt2) { g_ics.v2[a].alloc(NV); pending.h_v2[a] = hv[4 + a]; pending.d_v2[a] = g_ics.v2[a]; }
            }
            for (int i = 0; i < 7; i++) { g_ics.host[i] = i < nuse ? hv[i] : nullptr; g_ics.sig[i] = sig[i]; }
            g_ics.dim = so->DIM; g_ics.hii = so->HII_DIM;
            g_ics.valid = false; /* becomes valid once the (pipelined) upload has been issued in full */
        }
        DevBuf<float> d_density(N), d_v[3];
        float *host_v[3] = {mo->KEEP_3D_VELOCITIES ? pf->velocity_x : nullptr,
                            mo->KEEP_3D_VELOCITIES ? pf->velocity_y : nullptr, pf->velocity_z};
        PerturbDeviceIO io;
        memset(&io, 0, sizeof(io));
        io.hires_density = g_ics.hires; io.lowres_density = g_ics.hires;
        for (int a = 0; a < 3; a++) {
            io.v[a] = g_ics.v[a].p; io.v2[a] = g_ics.v2[a].p;
            if (host_v[a]) { d_v[a].alloc(N); io.vel[a] = d_v[a]; }
        }
        io.density = d_density;
        perturb_core(redshift, io, hit ? nullptr : &pending);
        g_ics.valid = true;
        d2h(pf->density, d_density, N * sizeof(float));
        for (int a = 0; a < 3; a++)
            if (host_v[a]) d2h(host_v[a], d_v[a], N * sizeof(float));
        if (resident_enabled()) { /* ComputeIonizedBox / ComputeBrightnessTemp read this box next (rt.h) */
            resident_put(pf->density, d_density.p, (size_t)N);
            d_density.p = nullptr; d_density.n = 0;
        }
        if (!use_cache) ics_cache_drop();
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        try { copy_stream_sync(); } catch (B200Error &) {}
        if (getenv("B200_VERBOSE") || e.code == B200_CUDAError)
            fprintf(stderr, "[21cmfast_b200] ComputePerturbedField: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

static void fill_device_io(PerturbDeviceIO &io, InitialConditions *d_boxes, PerturbedField *d_pf) {
    memset(&io, 0, sizeof(io));
    io.hires_density = d_boxes->hires_density; io.lowres_density = d_boxes->lowres_density;
    const bool on_hires = matter_options_global->PERTURB_ON_HIGH_RES;
    io.v[0] = on_hires ? d_boxes->hires_vx : d_boxes->lowres_vx;
    io.v[1] = on_hires ? d_boxes->hires_vy : d_boxes->lowres_vy;
    io.v[2] = on_hires ? d_boxes->hires_vz : d_boxes->lowres_vz;
    io.v2[0] = on_hires ? d_boxes->hires_vx_2LPT : d_boxes->lowres_vx_2LPT;
    io.v2[1] = on_hires ? d_boxes->hires_vy_2LPT : d_boxes->lowres_vy_2LPT;
    io.v2[2] = on_hires ? d_boxes->hires_vz_2LPT : d_boxes->lowres_vz_2LPT;
    io.density = d_pf->density;
    io.vel[0] = matter_options_global->KEEP_3D_VELOCITIES ? d_pf->velocity_x : nullptr;
    io.vel[1] = matter_options_global->KEEP_3D_VELOCITIES ? d_pf->velocity_y : nullptr;
    io.vel[2] = d_pf->velocity_z;
}

/* Slab-parallel variant of the device-resident entry point (see PerturbPartition): d_acc is a device
   buffer of HII_DIM^2 * HII_D_PARA 64-bit integers owned by the caller, who all-reduces it (SUM)
   between the phase-0 and the phase-1 call.  In phase 0 only the hi-res planes of the rank's own
   slab (plus one plane below it) are read. */
extern "C" int b200_ComputePerturbedField_device_part(float redshift, InitialConditions *d_boxes, PerturbedField *d_pf,
                                                      unsigned long long *d_acc, int part, int nparts, int phase) {
    try {
        require_params(false);
        rt_init();
        reset_stats();
        if (!d_acc || nparts < 1 || part < 0 || part >= nparts || (phase != 0 && phase != 1))
            b200_throw(B200_ValueError, "b200_ComputePerturbedField_device_part: bad partition arguments");
        DevTimer timer;
        timer.start();
        PerturbDeviceIO io;
        fill_device_io(io, d_boxes, d_pf);
        PerturbPartition pt;
        pt.part = part; pt.nparts = nparts; pt.phase = phase; pt.acc = d_acc;
        perturb_core(redshift, io, nullptr, pt);
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ComputePerturbedField_device_part: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

extern "C" int b200_ComputePerturbedField_device(float redshift, InitialConditions *d_boxes, PerturbedField *d_pf) {
    try {
        require_params(false);
        rt_init();
        reset_stats();
        DevTimer timer;
        timer.start();
        PerturbDeviceIO io;
        fill_device_io(io, d_boxes, d_pf);
        perturb_core(redshift, io);
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ComputePerturbedField_device: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

/* Slab-decomposed variant: DEVICE pointers to this rank's slabs (see perturb_core_slab); ranks connected
   with b200_dist_init / b200_dist_connect. */
extern "C" int b200_ComputePerturbedField_slab(float redshift, InitialConditions *d_boxes, PerturbedField *d_pf) {
    try {
        require_params(false);
        rt_init();
        reset_stats();
        DevTimer timer;
        timer.start();
        PerturbDeviceIO io;
        fill_device_io(io, d_boxes, d_pf);
        perturb_core_slab(redshift, io);
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_ComputePerturbedField_slab: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
