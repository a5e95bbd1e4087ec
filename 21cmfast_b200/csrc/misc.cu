/*
 * misc.cu -- test_filter (filtering.c:397-445) and ComputeBrightnessTemp
 * (BrightnessTemperatureBox.c:22-105; with USE_TS_FLUCT the caller's TsBox supplies the spin temperature).
 */
#include "fft.h"
#include "host_physics.h"

struct Cast64Args {
    long long n;
    const float *src;
    double *dst;
};
__global__ void float_to_double_kernel(Cast64Args a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x)
        a.dst[i] = (double)a.src[i];
}

/* r2c, divide by N, one filter_box, c2r: the reference's known-answer hook for its filter tests */
extern "C" int test_filter(float *input_box, double R, double R_param, double R_star, int filter_flag,
                           double *result) {
    (void)R_star;
    try {
        require_params(false);
        rt_init();
        g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0; g_stats.ms = 0;
        if (filter_flag < 0 || filter_flag > 4)
            b200_throw(B200_ValueError, "filter type %d is outside the scoped path (0-4 supported)", filter_flag);
        const SimulationOptions *so = simulation_options_global;
        const int nx = so->HII_DIM, ny = so->HII_DIM, nz = (int)(so->NON_CUBIC_FACTOR * so->HII_DIM);
        const long long N = (long long)nx * ny * nz;
        Fft3D *plan = fft_plan(nx, ny, nz);
        DevBuf<float> d_in(N), d_out(N);
        DevBuf<double> d_res(N);
        DevBuf<float2> kbox(plan->n_cplx()), work(plan->n_cplx());
        h2d(d_in, input_box, N * sizeof(float));
        ZPrologue pro;
        pro.src = d_in; pro.src_row_stride = nz;
        pro.post_scale = (float)(1.0 / (double)N);
        fft_r2c(plan, kbox, pro);
        KMul km;
        km.kind = KMUL_FILTER; km.filter_type = filter_flag; km.R = (float)R; km.R_param = (float)R_param;
        if (filter_flag == 3) { /* filter_box: exp(-R / R_param) with the quotient of two floats (filtering.c:320-322) */
            const float q = -(float)R / (float)R_param;
            km.r_const = exp((double)q);
        }
        km.dk[0] = 2.0 * M_PI / so->BOX_LEN; km.dk[1] = km.dk[0];
        km.dk[2] = 2.0 * M_PI / (so->BOX_LEN * so->NON_CUBIC_FACTOR);
        ZEpilogue epi;
        epi.dst = d_out; epi.dst_row_stride = nz;
        fft_c2r(plan, kbox, work, km, epi);
        Cast64Args ca = {N, d_out, d_res};
        B200_LAUNCH(float_to_double_kernel, dev_num_sms() * 4, 256, 0, ca);
        d2h(result, d_res, N * sizeof(double));
    } catch (B200Error &e) {
        if (getenv("B200_VERBOSE") || e.code == B200_CUDAError) fprintf(stderr, "[21cmfast_b200] test_filter: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

/* Measurement hook for bench.py's library-baseline leg: the library's own
   3-D transforms of an n^3 box timed alone with CUDA events on the library's stream -- r2c, plain
   c2r, and c2r with the per-radius window on load exactly as the ionisation ladder runs it (window
   table + expansion included).  Mean milliseconds per transform over `iters` repetitions; the box
   holds whatever the allocator returned (the FFT's time does not depend on the data). */
extern "C" int b200_fft_probe(int n, int iters, double box_len, double *ms_r2c, double *ms_c2r, double *ms_c2r_window) {
    try {
        rt_init();
        if (n < 16 || iters < 1) b200_throw(B200_ValueError, "b200_fft_probe: bad arguments");
        Fft3D *plan = fft_plan(n, n, n);
        DevBuf<float2> kbox(plan->n_cplx()), work(plan->n_cplx());
        dev_zero(kbox, plan->n_cplx() * sizeof(float2));
        dev_zero(work, plan->n_cplx() * sizeof(float2));
        const bool pow2 = (n & (n - 1)) == 0;
        DevBuf<float> wtab((size_t)window_table_size(plan)), wtab3(pow2 ? window_table3_size(plan) : 0);
        const double dk = 2.0 * M_PI / box_len;
        KMul km;
        km.kind = KMUL_FILTER; km.filter_type = 0; km.R = (float)(box_len / 40.0); km.fast = 1;
        km.dk[0] = km.dk[1] = km.dk[2] = dk;
        ZPrologue pro;
        pro.post_scale = 1.f / ((float)n * n * n);
        ZEpilogue epi;
        epi.clip = 1; epi.clip_lo = -1.f; epi.clip_hi = 1e6f;
        DevTimer t;
        for (int w = 0; w < 2; w++) { fft_r2c(plan, work, pro); fft_c2r(plan, kbox, work, KMul(), epi); } /* warm-up */
        t.start();
        for (int i = 0; i < iters; i++) fft_r2c(plan, work, pro);
        if (ms_r2c) *ms_r2c = t.stop_ms() / iters;
        t.start();
        for (int i = 0; i < iters; i++) fft_c2r(plan, kbox, work, KMul(), epi);
        if (ms_c2r) *ms_c2r = t.stop_ms() / iters;
        t.start();
        for (int i = 0; i < iters; i++) {
            window_table_build(plan, 0, km.R, dk, wtab);
            km.wtab = wtab; km.wtab_n = window_table_size(plan);
            if (pow2) { window_table_expand(plan, wtab, wtab3); km.wtab3 = wtab3; }
            fft_c2r(plan, kbox, work, km, epi);
        }
        if (ms_c2r_window) *ms_c2r_window = t.stop_ms() / iters;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_fft_probe: %s\n", e.msg);
        return e.code;
    }
    return 0;
}

struct TbArgs {
    long long n;
    const float *density, *xH;
    float *tb;
    float const_factor;
    const float *Ts; /* spin temperature (USE_TS_FLUCT) or null: saturated limit */
    float *tau;      /* 21-cm optical depth output, with Ts only */
    float redshift, T_rad;
    int *nonfinite;
};
__global__ void brightness_kernel(TbArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        float tb = a.const_factor * a.xH[i] * (1 + a.density[i]);
        if (a.Ts) {
            /* prefactors -> optical depth (1000: K -> mK), then the full radiative-transfer form */
            const float Ts = a.Ts[i];
            tb *= (1. + a.redshift) / (1000. * Ts);
            a.tau[i] = tb;
            tb = (1. - exp(-(double)tb)) * 1000. * (Ts - a.T_rad) / (1. + a.redshift);
        }
        a.tb[i] = tb;
        if (!isfinite(tb)) *a.nonfinite = 1;
    }
}

extern "C" int ComputeBrightnessTemp(float redshift, TsBox *spin_temp, IonizedBox *ionized_box,
                                     PerturbedField *perturb_field, BrightnessTemp *box) {
    try {
        require_params(true);
        rt_init();
        g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0; g_stats.ms = 0;
        const bool ts = astro_options_global->USE_TS_FLUCT;
        if (ts && (!spin_temp || !spin_temp->spin_temperature || !box->tau_21))
            b200_throw(B200_ValueError, "ComputeBrightnessTemp: USE_TS_FLUCT needs the TsBox's spin_temperature and tau_21");
        const SimulationOptions *so = simulation_options_global;
        const CosmoParams *cp = cosmo_params_global;
        const long long N = (long long)so->HII_DIM * so->HII_DIM * hii_d_para();
        const float const_factor =
            27 * (cp->OMb * cp->hlittle * cp->hlittle / 0.023) *
            sqrt((0.15 / (cp->OMm) / (cp->hlittle) / (cp->hlittle)) * (1. + redshift) / 10.0);
        DevBuf<float> d_d_own, d_x_own, d_t(N);
        /* inputs: the copies the two producing calls left on the device (opt-in residency, rt.h), else uploads */
        const float *d_d = resident_get(perturb_field->density, (size_t)N);
        const float *d_x = resident_get(ionized_box->neutral_fraction, (size_t)N);
        if (!d_d) { d_d_own.alloc(N); h2d(d_d_own, perturb_field->density, N * sizeof(float)); d_d = d_d_own; }
        if (!d_x) { d_x_own.alloc(N); h2d(d_x_own, ionized_box->neutral_fraction, N * sizeof(float)); d_x = d_x_own; }
        DevBuf<float> d_ts, d_tau;
        DevBuf<int> d_flag(1);
        dev_zero(d_flag, sizeof(int));
        if (ts) {
            d_ts.alloc(N); d_tau.alloc(N);
            h2d(d_ts, spin_temp->spin_temperature, N * sizeof(float));
        }
        const float T_rad = pc::T_cmb * (1 + redshift);
        TbArgs a = {N, d_d, d_x, d_t, const_factor, d_ts.p, d_tau.p, redshift, T_rad, d_flag};
        B200_LAUNCH(brightness_kernel, dev_num_sms() * 4, 256, 0, a);
        d2h(box->brightness_temp, d_t, N * sizeof(float));
        if (ts) d2h(box->tau_21, d_tau, N * sizeof(float));
        int flag = 0;
        d2h(&flag, d_flag, sizeof(int));
        if (flag) b200_throw(B200_InfinityorNaNError, "brightness temperature is infinite or NaN");
    } catch (B200Error &e) {
        if (getenv("B200_VERBOSE") || e.code == B200_CUDAError)
            fprintf(stderr, "[21cmfast_b200] ComputeBrightnessTemp: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
