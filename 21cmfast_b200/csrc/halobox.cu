/*
 * halobox.cu -- Lagrangian source grids: ComputeHaloBox for SOURCE_MODEL = L-INTEGRAL.
 *
 * Replaces set_fixed_grids (HaloBox.c:296-436) + move_grid_galprops (map_mass.c:214-331): every cell of
 * the initial-condition grid (low-res, or hi-res with PERTURB_ON_HIGH_RES) carries the conditional
 * mass-function integrals of its linear density -- ionising photons N_ion and star formation -- and is
 * moved to its Eulerian position with the same ZA / 2LPT displacement as the perturbed field, where the
 * two quantities are deposited cloud-in-cell on the low-res grid.  The integrals come from two 400-point
 * tables over the grid's density range (initialise_Nion_Conditional_spline /
 * initialise_SFRD_Conditional_table, built on the host) and are evaluated per cell with the reference's
 * arithmetic (EvaluateRGTable1D_f in double on a float table, then exp).
 *
 * One kernel does table lookup, displacement and deposit; the deposit accumulates in DOUBLE with native
 * L2 reductions (the reference adds floats under `omp atomic`, in thread order: neither is ordered, ours
 * carries 29 more bits), and a second kernel rounds the two accumulators to the float outputs
 * (whalo_sfr = n_ion / (t_h t_star) with recombinations, map_mass.c:326-331).
 *
 * In scope: no mini-halos, no spin temperature (halo_xray), no extra fields -- anything else returns ValueError.
 * For the mass functions without a conditional form (WATSON, WATSON-Z, REED07, YUNG24) the grids are rescaled to the
 * unconditional expectation (mean_fix_grids, HaloBox.c:209-243).
 */
#include "rt.h"
#include "host_physics.h"
#include "../../include/py21cmfast_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" double dicke(double z);
extern "C" double minimum_source_mass(double redshift, bool xray);

struct DevCondTable {
    double x_min, x_width, inv_width;
    float y[N_DENS_INTERP];
};

struct GalpropsArgs {
    int dn[3];                 /* source (Lagrangian) grid */
    int on[3];                 /* output grid */
    const float *dens;         /* linear density of the source grid at z = 0 */
    const float *v[3], *v2[3]; /* displacement fields of the source grid (v2 null unless 2LPT) */
    double growth;             /* dicke(z) */
    double vdf[3], vdf2[3];    /* displacement factors in source-cell units (map_mass.c:243-250) */
    double ratio_out;          /* out_dim / dens_dim */
    const DevCondTable *nion, *sfrd;
    double pref_nion, pref_sfr;
    double *acc_nion, *acc_sfr;
};

/* EvaluateRGTable1D_f (interpolation.c:123-131) then exp */
DEV double cond_table_eval(const DevCondTable *t, double x) {
    const int idx = (int)floor((x - t->x_min) * t->inv_width);
    const double table_val = t->x_min + t->x_width * (float)idx;
    const double f = (x - table_val) * t->inv_width;
    return exp((double)t->y[idx] * (1 - f) + (double)t->y[idx + 1] * f);
}

DEV int wrap_cell(int i, int n) {
    while (i >= n) i -= n;
    while (i < 0) i += n;
    return i;
}

__global__ void __launch_bounds__(256) galprops_cic_kernel(GalpropsArgs a) {
    __shared__ DevCondTable t_nion, t_sfrd;
    for (int i = threadIdx.x; i < N_DENS_INTERP; i += blockDim.x) {
        t_nion.y[i] = a.nion->y[i];
        t_sfrd.y[i] = a.sfrd->y[i];
    }
    if (threadIdx.x == 0) {
        t_nion.x_min = a.nion->x_min; t_nion.x_width = a.nion->x_width; t_nion.inv_width = a.nion->inv_width;
        t_sfrd.x_min = a.sfrd->x_min; t_sfrd.x_width = a.sfrd->x_width; t_sfrd.inv_width = a.sfrd->inv_width;
    }
    __syncthreads();
    const long long np = (long long)a.dn[0] * a.dn[1] * a.dn[2];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(p % a.dn[2]);
        const int j = (int)((p / a.dn[2]) % a.dn[1]);
        const int i = (int)(p / ((long long)a.dn[2] * a.dn[1]));
        /* the displacement grid is the source grid itself (set_fixed_grids passes grid_dim twice) */
        double pos[3] = {(double)i, (double)j, (double)k};
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            pos[ax] += (double)a.v[ax][p] * a.vdf[ax];
            if (a.v2[0]) pos[ax] -= (double)a.v2[ax][p] * a.vdf2[ax];
            pos[ax] *= a.ratio_out;
        }
        const double curr_dens = (double)a.dens[p] * a.growth;
        const double w_nion = cond_table_eval(&t_nion, curr_dens) * a.pref_nion;
        const double w_sfr = cond_table_eval(&t_sfrd, curr_dens) * a.pref_sfr;
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
            const int wx = wrap_cell(ip[0] + cx, a.on[0]), wy = wrap_cell(ip[1] + cy, a.on[1]), wz = wrap_cell(ip[2] + cz, a.on[2]);
            const long long q = (long long)wz + (long long)a.on[2] * ((long long)wy + (long long)a.on[1] * wx);
            atomic_add_f64(&a.acc_nion[q], w_nion * w);
            atomic_add_f64(&a.acc_sfr[q], w_sfr * w);
        }
    }
}

struct GalpropsOutArgs {
    long long n;
    const double *acc_nion, *acc_sfr;
    float *n_ion, *halo_sfr, *whalo_sfr; /* whalo_sfr may be null */
    double pref_wsfr;
};
__global__ void galprops_out_kernel(GalpropsOutArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        const float nion = (float)a.acc_nion[i];
        a.n_ion[i] = nion;
        a.halo_sfr[i] = (float)a.acc_sfr[i];
        if (a.whalo_sfr) a.whalo_sfr[i] = (float)((double)nion * a.pref_wsfr);
    }
}

/* grid extrema of dens * growth, seeded with 0 like the reference's reductions (HaloBox.c:299-300,358-365) */
struct DensRangeArgs {
    long long n;
    const float *dens;
    int *keys; /* {min key, max key} of the raw densities */
};
__global__ void dens_range_kernel(DensRangeArgs a) {
    float lo = 0.f, hi = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a.dens[i];
        lo = fminf(lo, d);
        hi = fmaxf(hi, d);
    }
    atomic_min_i32(&a.keys[0], float_order_key(float_as_int_bits(lo)));
    atomic_max_i32(&a.keys[1], float_order_key(float_as_int_bits(hi)));
}

/* box mean of a float grid (deterministic block sums, added in order on the host) and its rescaling */
struct GridSumArgs {
    long long n;
    const float *grid;
    double *partial; /* [gridDim.x] */
};
__global__ void __launch_bounds__(256) grid_sum_kernel(GridSumArgs a) {
    __shared__ double red[256];
    double acc = 0.;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)a.grid[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partial[blockIdx.x] = red[0];
}
struct GridScaleArgs {
    long long n;
    float *grid;
    double ratio;
};
__global__ void grid_scale_kernel(GridScaleArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
        a.grid[i] = (float)((double)a.grid[i] * a.ratio);
}

static int grid_of(long long n) {
    long long want = (n + 255) / 256, cap = (long long)dev_num_sms() * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

extern "C" int ComputeHaloBox(double redshift, InitialConditions *ini_boxes, HaloCatalog *halos, TsBox *previous_spin_temp,
                              IonizedBox *previous_ionize_box, HaloBox *grids) {
    (void)halos; (void)previous_spin_temp; (void)previous_ionize_box;
    try {
        require_params(true);
        rt_init();
        const SimulationOptions *so = simulation_options_global;
        const MatterOptions *mo = matter_options_global;
        const AstroOptions *ao = astro_options_global;
        if (mo->SOURCE_MODEL != SRC_L_INTEGRAL)
            b200_throw(B200_ValueError, "ComputeHaloBox: only SOURCE_MODEL = L-INTEGRAL is built (the halo samplers are out of scope)");
        /* USE_UPPER_STELLAR_TURNOVER (the reference's default) is accepted: it acts on sampled halos (get_halo_stellarmass,
           scaling_relations.c:326-370) and on L_X / SFR (:314-324, spin temperature), not on the fixed grids' integrals */
        if (ao->USE_MINI_HALOS || ao->USE_TS_FLUCT || config_settings.EXTRA_HALOBOX_FIELDS || ao->HALO_SCALING_RELATIONS_MEDIAN ||
            ao->PHOTON_CONS_TYPE != 0)
            b200_throw(B200_ValueError, "ComputeHaloBox: mini-halos, spin temperature, extra fields, median scaling relations "
                                        "and photon conservation are not built");
        if (mo->USE_INTERPOLATION_TABLES != 2)
            b200_throw(B200_ValueError, "this build needs USE_INTERPOLATION_TABLES='hmf-interpolation'");
        if (!ini_boxes || !grids || !grids->n_ion || !grids->halo_sfr)
            b200_throw(B200_ValueError, "ComputeHaloBox: NULL struct/array");
        const bool recomb = ao->RECOMB_MODEL != 0;
        if (recomb && !grids->whalo_sfr) b200_throw(B200_ValueError, "ComputeHaloBox: RECOMB_MODEL != none needs whalo_sfr");
        const bool hires = mo->PERTURB_ON_HIGH_RES;
        /* PERTURB_ALGORITHM = LINEAR moves the sources by the first-order velocities, like ZELDOVICH (map_mass.c:269-283) */
        const bool lpt2 = mo->PERTURB_ALGORITHM == PERTURB_2LPT;
        const int on[3] = {so->HII_DIM, so->HII_DIM, hii_d_para()};
        const int dn[3] = {hires ? so->DIM : so->HII_DIM, hires ? so->DIM : so->HII_DIM, hires ? d_para() : hii_d_para()};
        const long long N = (long long)on[0] * on[1] * on[2], NS = (long long)dn[0] * dn[1] * dn[2];
        const float *h_dens = hires ? ini_boxes->hires_density : ini_boxes->lowres_density;
        const float *h_v[3] = {hires ? ini_boxes->hires_vx : ini_boxes->lowres_vx, hires ? ini_boxes->hires_vy : ini_boxes->lowres_vy,
                               hires ? ini_boxes->hires_vz : ini_boxes->lowres_vz};
        const float *h_v2[3] = {hires ? ini_boxes->hires_vx_2LPT : ini_boxes->lowres_vx_2LPT,
                                hires ? ini_boxes->hires_vy_2LPT : ini_boxes->lowres_vy_2LPT,
                                hires ? ini_boxes->hires_vz_2LPT : ini_boxes->lowres_vz_2LPT};
        if (!h_dens || !h_v[0] || !h_v[1] || !h_v[2] || (lpt2 && (!h_v2[0] || !h_v2[1] || !h_v2[2])))
            b200_throw(B200_ValueError, "ComputeHaloBox: a required IC array is NULL");

        /* scalars (ComputeHaloBox, HaloBox.c:603-640; get_log10_turnovers without mini-halos, :456-463) */
        ScalingConstants sc;
        set_scaling_constants(redshift, &sc);
        grids->log10_Mcrit_ACG_ave = log10(sc.mturn_a_nofb);
        grids->log10_Mcrit_MCG_ave = log10(0.); /* mturn_m_nofb is 0 without mini-halos: -inf, as the reference stores */
        const double M_min = minimum_source_mass(redshift, false), M_max = pc::M_MAX_INTEGRAL;

        DevBuf<float> d_dens((size_t)NS), d_v[3], d_v2[3];
        h2d(d_dens, h_dens, NS * sizeof(float));
        for (int ax = 0; ax < 3; ax++) {
            d_v[ax].alloc((size_t)NS);
            h2d(d_v[ax], h_v[ax], NS * sizeof(float));
            if (lpt2) { d_v2[ax].alloc((size_t)NS); h2d(d_v2[ax], h_v2[ax], NS * sizeof(float)); }
        }
        DevBuf<double> acc_nion((size_t)N), acc_sfr((size_t)N);
        dev_zero(acc_nion, N * sizeof(double));
        dev_zero(acc_sfr, N * sizeof(double));
        DevBuf<float> d_nion((size_t)N), d_sfr((size_t)N), d_wsfr(recomb ? (size_t)N : 0);

        if (M_min < M_max) { /* set_fixed_grids */
            const double growth = dicke(redshift);
            const double M_cell = rho_crit() * cosmo_params_global->OMm * box_volume() / (double)NS;
            /* table range: extrema of dens * growth seeded with 0, widened by 0.1 % (HaloBox.c:358-380) */
            DevBuf<int> d_keys(2);
            const int init[2] = {float_order_key(0), float_order_key(0)}; /* the bits of 0.0f */
            h2d(d_keys, init, sizeof(init));
            DensRangeArgs ra = {NS, d_dens, d_keys};
            B200_LAUNCH(dens_range_kernel, grid_of(NS), 256, 0, ra);
            int keys[2];
            d2h(keys, d_keys, sizeof(keys));
            /* the reference multiplies in double inside the reduction: min over (double)dens * growth */
            const double min_density = (double)float_from_order_key(keys[0]) * growth * 1.001;
            const double max_density = (double)float_from_order_key(keys[1]) * growth * 1.001;

            const int method = ao->INTEGRATION_METHOD_ATOMIC;
            if (method == INTEG_GL) initialise_GL(log(M_min), log(M_max));
            FcollTable t_sfrd, t_nion;
            const ScalingConstants sc_sfrd = evolve_scaling_constants_sfr(&sc);
            build_cond_table(&t_sfrd, redshift, min_density, max_density, M_min, M_max, M_cell, &sc_sfrd, method, -50., so->N_THREADS);
            build_cond_table(&t_nion, redshift, min_density, max_density, M_min, M_max, M_cell, &sc, method, -40., so->N_THREADS);
            DevCondTable ht[2];
            const FcollTable *src[2] = {&t_nion, &t_sfrd};
            for (int i = 0; i < 2; i++) {
                ht[i].x_min = src[i]->x_min; ht[i].x_width = src[i]->x_width; ht[i].inv_width = 1.0 / src[i]->x_width;
                memcpy(ht[i].y, src[i]->y, sizeof(ht[i].y));
            }
            DevBuf<DevCondTable> d_tab(2);
            h2d(d_tab, ht, sizeof(ht));

            GalpropsArgs ga;
            memset(&ga, 0, sizeof(ga));
            const double boxlen = so->BOX_LEN, boxlen_z = boxlen * so->NON_CUBIC_FACTOR;
            const double box_size[3] = {boxlen, boxlen, boxlen_z};
            const double init_growth = dicke(so->INITIAL_REDSHIFT);
            const double disp2 = -(3.0 / 7.0) * growth * growth, init_disp2 = -(3.0 / 7.0) * init_growth * init_growth;
            for (int ax = 0; ax < 3; ax++) {
                ga.dn[ax] = dn[ax]; ga.on[ax] = on[ax];
                ga.v[ax] = d_v[ax]; ga.v2[ax] = lpt2 ? d_v2[ax].p : nullptr;
                ga.vdf[ax] = (growth - init_growth) / box_size[ax] * dn[ax];
                ga.vdf2[ax] = (disp2 - init_disp2) / box_size[ax] * dn[ax];
            }
            ga.dens = d_dens; ga.growth = growth;
            ga.ratio_out = (double)on[0] / (double)dn[0];
            ga.nion = d_tab.p; ga.sfrd = d_tab.p + 1;
            const double vol_ratio_out = (double)N / (double)NS;
            const double pref_stars = rho_crit() * cosmo_params_global->OMb * sc.fstar_10 * vol_ratio_out;
            ga.pref_sfr = pref_stars / sc.t_star / sc.t_h;
            ga.pref_nion = pref_stars * sc.fesc_10 * sc.pop2_ion;
            ga.acc_nion = acc_nion; ga.acc_sfr = acc_sfr;
            B200_LAUNCH(galprops_cic_kernel, grid_of(NS), 256, 0, ga);
        }
        GalpropsOutArgs oa = {N, acc_nion, acc_sfr, d_nion, d_sfr, recomb ? d_wsfr.p : nullptr, 1. / sc.t_h / sc.t_star};
        B200_LAUNCH(galprops_out_kernel, grid_of(N), 256, 0, oa);
        /* mean_fix_grids (HaloBox.c:209-243; set_scaling_constants sets fix_mean for the mass functions that have no
           conditional form, scaling_relations.c:40-42): every grid is rescaled so that its box mean equals the
           unconditional integral (get_uhmf_averages, :105-163) */
        const bool fix_mean = mo->HMF == HMF_WATSON || mo->HMF == HMF_WATSON_Z || mo->HMF == HMF_REED07 || mo->HMF == HMF_YUNG24;
        if (fix_mean && M_min < M_max) {
            const double lnMmin = log(M_min), lnMmax = log(M_max);
            const double Mturn_a = pow(10., grids->log10_Mcrit_ACG_ave);
            const ScalingConstants sc_sfrd = evolve_scaling_constants_sfr(&sc);
            const double pref_stars = rho_crit() * cosmo_params_global->OMb * sc.fstar_10;
            const double pref_sfr = pref_stars / sc.t_star / sc.t_h;
            const double i_fesc = Nion_General(redshift, lnMmin, lnMmax, Mturn_a, &sc);
            const double i_stars = Nion_General(redshift, lnMmin, lnMmax, Mturn_a, &sc_sfrd);
            const double want[3] = {i_fesc * pref_stars * sc.fesc_10 * sc.pop2_ion, i_stars * pref_sfr,
                                    i_fesc * pref_sfr * sc.fesc_10 * sc.pop2_ion};
            float *grid[3] = {d_nion.p, d_sfr.p, recomb ? d_wsfr.p : nullptr};
            const int nb = grid_of(N);
            DevBuf<double> d_part((size_t)nb);
            std::vector<double> hp((size_t)nb);
            for (int g = 0; g < 3; g++) {
                if (!grid[g]) continue;
                GridSumArgs sa = {N, grid[g], d_part};
                B200_LAUNCH(grid_sum_kernel, nb, 256, 0, sa);
                d2h(hp.data(), d_part, hp.size() * sizeof(double));
                double sum = 0.;
                for (int i = 0; i < nb; i++) sum += hp[i];
                GridScaleArgs ga2 = {N, grid[g], want[g] / (sum / (double)N)};
                B200_LAUNCH(grid_scale_kernel, nb, 256, 0, ga2);
            }
        }
        d2h(grids->n_ion, d_nion, N * sizeof(float));
        d2h(grids->halo_sfr, d_sfr, N * sizeof(float));
        if (recomb) d2h(grids->whalo_sfr, d_wsfr, N * sizeof(float));
    } catch (B200Error &e) {
        if (getenv("B200_VERBOSE") || e.code == B200_CUDAError) fprintf(stderr, "[21cmfast_b200] ComputeHaloBox: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
