/*
 * ics.cu -- ComputeInitialConditions: Gaussian random field + Zel'dovich / 2LPT displacement
 * fields (reference InitialConditions.c:26-772, rng.c:31-90).
 *
 * Scope: V_CB_MODEL without fluctuations, analytic power spectra; velocities on the low-res grid
 * (default) or, with PERTURB_ON_HIGH_RES, unfiltered on the hi-res grid.
 *
 * Random numbers: with the default B200_IC_RNG=gsl the Gaussian stream is the reference's own for
 * N_THREADS = 1 -- GSL-seeded MT19937 feeding the polar Box-Muller, two deviates per mode in
 * x-major order (InitialConditions.c:111-131) -- generated on the host (the generator is
 * sequential) and streamed to the device in x-slabs while the device computes sqrt(V P(k) / 2).
 * B200_IC_RNG=device uses a counter-based generator on the device instead (no seed parity with
 * the reference; used for the large benchmark boxes where ICs are outside the timed region).
 * Everything else (Hermitian fix-up, 15 FFTs at DIM^3, gradients, Laplacians, 2LPT source,
 * sub-sampling) runs on the device.
 */
#include "fft.h"
#include "host_numerics.h"
#include "host_physics.h"

/* gslrng.cu */
struct GslStream;
GslStream *gsl_stream_create(unsigned long mt_seed);
void gsl_stream_destroy(GslStream *s);
void gsl_stream_gaussians(GslStream *s, double *d_out, long long want);

#include <vector>

/* ------------------------------------------------------------------ power spectrum on device */
struct PsConsts {
    int which;
    double sound_horizon, alpha_nu, beta_c, omhh, f_nu, theta_cmb, sigma_norm;
    double ps_norm, n_s, h, OMm, OMb;
};
void ps_export_consts(PsConsts *out); /* host_physics_export below */

DEV double ps_transfer(double k, const PsConsts &c) {
    if (c.which == 0) { /* EH99, cosmology.c:52-71 */
        const double q = k * pow(c.theta_cmb, 2) / c.omhh;
        const double sa = sqrt(c.alpha_nu);
        const double gamma_eff = sa + (1.0 - sa) / (1.0 + pow(0.43 * k * c.sound_horizon, 4));
        const double q_eff = q / gamma_eff;
        double TF = log(M_E + 1.84 * c.beta_c * sa * q_eff);
        TF /= TF + pow(q_eff, 2) * (14.4 + 325.0 / (1.0 + 60.5 * pow(q_eff, 1.11)));
        const double q_nu = 3.92 * q / sqrt(c.f_nu / 1.0);
        TF *= 1.0 + (1.2 * pow(c.f_nu, 0.64) * pow(1.0, 0.3 + 0.6 * c.f_nu)) / (pow(q_nu, -1.6) + pow(q_nu, 0.8));
        return TF;
    }
    if (c.which == 1) {
        const double gamma = c.OMm * c.h * exp(-(c.OMb) - (c.OMb / c.OMm));
        const double q = k / (c.h * gamma);
        return (log(1.0 + 2.34 * q) / (2.34 * q)) *
               pow(1.0 + 3.89 * q + pow(16.1 * q, 2) + pow(5.46 * q, 3) + pow(6.71 * q, 4), -0.25);
    }
    if (c.which == 2) {
        const double gamma = c.OMm * c.h * c.h;
        const double aa = 6.4 / gamma, bb = 3.0 / gamma, c2 = 1.7 / gamma, nu = 1.13;
        return pow(1 + pow(aa * k + pow(bb * k, 1.5) + pow(c2 * k, 2), nu), -1. / nu);
    }
    if (c.which == 3) {
        const double gamma = c.OMm * c.h * exp(-(c.OMb) - (c.OMb / c.OMm));
        const double aa = 8.0 / (c.h * gamma), bb = 4.7 / pow(c.h * gamma, 2);
        return 1 + aa * k + bb * k * k;
    }
    const double gamma = c.OMm * c.h * c.h * exp(-(c.OMb) - (c.OMb / c.OMm));
    const double aa = 1.7 / gamma, bb = 9.0 / pow(gamma, 1.5), c2 = 1.0 / pow(gamma, 2);
    return 139.284 / (1 + aa * k + bb * pow(k, 1.5) + c2 * k * k);
}
DEV double ps_power(double k, const PsConsts &c) { /* power_in_k, cosmology.c:278-303 */
    if (k == 0.) return 0.;
    double T = ps_transfer(k, c);
    T *= k * k;
    const double primordial = c.ps_norm * pow(k / 0.05, c.n_s - 1.);
    return c.sigma_norm * primordial * T * T / pow(k, 3);
}

/* ------------------------------------------------------------------ counter-based RNG (device mode) */
DEV unsigned int mulhi32(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
DEV void philox4x32(unsigned int ctr[4], unsigned int k0, unsigned int k1) {
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = mulhi32(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const unsigned int hi1 = mulhi32(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const unsigned int n0 = hi1 ^ ctr[1] ^ k0, n2 = hi0 ^ ctr[3] ^ k1;
        ctr[0] = n0; ctr[1] = lo1; ctr[2] = n2; ctr[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct ModeArgs {
    int nx, ny, nz, nzc, pitch;
    int x0, nxs;              /* x-slab handled by this launch */
    double dk[3], volume;
    PsConsts ps;
    const double *gauss;      /* host-stream mode: 2 doubles per mode of the slab, or null */
    unsigned long long seed;  /* device mode */
    float2 *box;
};
/* sample_ic_modes, InitialConditions.c:103-139 */
__global__ void __launch_bounds__(256) ic_modes_kernel(ModeArgs a) {
    const long long per_slab = (long long)a.nxs * a.ny * a.nzc;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < per_slab;
         m += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(m % a.nzc);
        const int iy = (int)((m / a.nzc) % a.ny);
        const int ix = a.x0 + (int)(m / ((long long)a.nzc * a.ny));
        const double kx = ((ix <= a.nx / 2) ? ix : ix - a.nx) * a.dk[0];
        const double ky = ((iy <= a.ny / 2) ? iy : iy - a.ny) * a.dk[1];
        const double kz = iz * a.dk[2];
        const double kmag = sqrt(kx * kx + ky * ky + kz * kz);
        const double amp = sqrt(a.volume * ps_power(kmag, a.ps) / 2.0);
        double ga, gb;
        if (a.gauss) {
            ga = a.gauss[2 * m];
            gb = a.gauss[2 * m + 1];
        } else {
            const long long gm = ((long long)ix * a.ny + iy) * a.nzc + iz;
            unsigned int c[4] = {(unsigned int)gm, (unsigned int)(gm >> 32), 0x21c3fa57u, 0u};
            philox4x32(c, (unsigned int)a.seed, (unsigned int)(a.seed >> 32));
            const double u1 = ((double)c[0] + 0.5) / 4294967296.0, u2 = ((double)c[1] + 0.5) / 4294967296.0;
            const double r = sqrt(-2.0 * log(u1));
            double s, co;
            sincospi(2.0 * u2, &s, &co);
            ga = r * co;
            gb = r * s;
        }
        a.box[((long long)ix * a.ny + iy) * a.pitch + iz] = make_float2((float)(amp * ga), (float)(amp * gb));
    }
}

struct ConjArgs {
    int nx, ny, nz, nzc, pitch;
    float2 *box;
};
/* adj_complex_conj, InitialConditions.c:26-101: Hermitian symmetry on the kz = 0 and kz = Nyquist
   planes, real corners, zero DC.  One thread per (i, j) pair the reference's loops visit. */
__global__ void ic_hermitian_kernel(ConjArgs a) {
    const int mx = a.nx / 2, my = a.ny / 2, mz = a.nz / 2;
    const long long sy = a.pitch, sx = (long long)a.ny * a.pitch;
    const int kplanes[2] = {0, mz};
    /* loop A: i in 1..mx-1, all j handled by the reference's two inner loops */
    const long long nA = (long long)(mx - 1) * (my + 1);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nA;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = 1 + (int)(t / (my + 1)), j = (int)(t % (my + 1));
        for (int kk = 0; kk < 2; kk++) {
            const int k = kplanes[kk];
            if (kk == 1 && mz == 0) break;
            if (j == 0 || j == my) {
                /* "just j corners": box[i,j,k] = conj(box[nx-i,j,k]) */
                const float2 s = a.box[(long long)(a.nx - i) * sx + j * sy + k];
                a.box[(long long)i * sx + j * sy + k] = make_float2(s.x, -s.y);
            } else {
                /* "all of j": box[i,j] = conj(box[nx-i,ny-j]); box[i,ny-j] = conj(box[nx-i,j]) */
                const float2 s1 = a.box[(long long)(a.nx - i) * sx + (long long)(a.ny - j) * sy + k];
                const float2 s2 = a.box[(long long)(a.nx - i) * sx + j * sy + k];
                a.box[(long long)i * sx + j * sy + k] = make_float2(s1.x, -s1.y);
                a.box[(long long)i * sx + (long long)(a.ny - j) * sy + k] = make_float2(s2.x, -s2.y);
            }
            if (mz == 0) break;
        }
    }
    /* loop B ("i corners"): i in {0, mx}, j in 1..my-1: box[i,j,k] = conj(box[i,ny-j,k]) */
    const long long nB = 2LL * (my - 1 > 0 ? my - 1 : 0);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nB;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (t / (my - 1)) ? mx : 0, j = 1 + (int)(t % (my - 1));
        if (i == mx && mx == 0) continue;
        for (int kk = 0; kk < 2; kk++) {
            const int k = kplanes[kk];
            const float2 s = a.box[(long long)i * sx + (long long)(a.ny - j) * sy + k];
            a.box[(long long)i * sx + j * sy + k] = make_float2(s.x, -s.y);
            if (mz == 0) break;
        }
    }
    /* seven real corners and the zero mode (InitialConditions.c:38-53) */
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int ci[7][3] = {{0, 0, mz}, {0, my, 0}, {0, my, mz}, {mx, 0, 0}, {mx, 0, mz}, {mx, my, 0}, {mx, my, mz}};
        for (int c = 0; c < 7; c++) a.box[(long long)ci[c][0] * sx + ci[c][1] * sy + ci[c][2]].y = 0.f;
        a.box[0] = make_float2(0.f, 0.f);
    }
}

struct SubsampleArgs {
    int ln[3], hn[3], hnzc;
    double ratio;
    const float *hi_padded;
    float *lo;
    float scale;
};
/* nearest-cell sub-sampling hi -> lo through resample_index (indexing.h:110-114) */
__global__ void subsample_kernel(SubsampleArgs a) {
    const long long n = (long long)a.ln[0] * a.ln[1] * a.ln[2];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % a.ln[2]), j = (int)((t / a.ln[2]) % a.ln[1]), i = (int)(t / ((long long)a.ln[2] * a.ln[1]));
        const int hi = (int)(i * a.ratio + 0.5), hj = (int)(j * a.ratio + 0.5), hk = (int)(k * a.ratio + 0.5);
        a.lo[t] = a.hi_padded[(long long)hk + 2LL * a.hnzc * ((long long)hj + (long long)a.hn[1] * hi)] * a.scale;
    }
}

struct Lpt2Args {
    long long nrows;
    int nz, nzc;
    const float *dii, *djj; /* unpadded diagonal terms */
    const float *pij;       /* padded */
    float *src;             /* padded accumulator */
    int first;
};
/* 2LPT source accumulation, InitialConditions.c:452-476 (same float/double sequence) */
__global__ void lpt2_source_kernel(Lpt2Args a) {
    for (long long row = blockIdx.x; row < a.nrows; row += gridDim.x)
        for (int z = threadIdx.x; z < a.nz; z += blockDim.x) {
            const long long r = row * a.nz + z, f = row * 2 * a.nzc + z;
            const double cii = a.dii[r], cjj = a.djj[r], cij = a.pij[f];
            float s = a.first ? 0.f : a.src[f];
            s = (float)((double)s + cii * cjj);
            s = (float)((double)s - cij * cij);
            a.src[f] = s;
        }
}

struct ScaleArgs {
    long long nrows;
    int nz, nzc;
    float *padded;
    float div;
    const float *unpadded_src; /* optional: padded[...] = (src * mul) / div (reverse branch) */
    float mul;
};
__global__ void scale_rows_kernel(ScaleArgs a) {
    for (long long row = blockIdx.x; row < a.nrows; row += gridDim.x)
        for (int z = threadIdx.x; z < a.nz; z += blockDim.x) {
            const long long f = row * 2 * a.nzc + z;
            if (a.unpadded_src)
                a.padded[f] = (a.unpadded_src[row * a.nz + z] * a.mul) / a.div;
            else
                a.padded[f] = a.padded[f] / a.div;
        }
}

/* ------------------------------------------------------------------ orchestration */
extern "C" int ComputeInitialConditions(unsigned long long random_seed, InitialConditions *boxes) {
    try {
        require_params(false);
        rt_init();
        g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0; g_stats.ms = 0;
        DevTimer timer;
        timer.start();
        const SimulationOptions *so = simulation_options_global;
        const MatterOptions *mo = matter_options_global;
        if (!boxes || !boxes->hires_density || !boxes->lowres_density)
            b200_throw(B200_ValueError, "ComputeInitialConditions: NULL struct/array");
        if (mo->V_CB_MODEL == 2) b200_throw(B200_ValueError, "V_CB_MODEL=FLUCTS (CLASS tables) is outside the scoped path");
        const int hn[3] = {so->DIM, so->DIM, d_para()};
        const int ln[3] = {so->HII_DIM, so->HII_DIM, hii_d_para()};
        const long long M = (long long)hn[0] * hn[1] * hn[2], N = (long long)ln[0] * ln[1] * ln[2];
        const float VOLUME = so->BOX_LEN * so->BOX_LEN * so->NON_CUBIC_FACTOR * so->BOX_LEN;
        const double ratio = hn[0] / (double)ln[0];
        Fft3D *plan = fft_plan(hn[0], hn[1], hn[2]);
        const int nzc = plan->pitch; /* row pitch of the padded / complex layouts */
        const int nzc_modes = plan->nzc;
        const long long Mk = (long long)plan->n_cplx();
        const double box_len[3] = {so->BOX_LEN, so->BOX_LEN, so->NON_CUBIC_FACTOR * so->BOX_LEN};
        const double dk[3] = {2. * M_PI / box_len[0], 2. * M_PI / box_len[1], 2. * M_PI / box_len[2]};
        const int row_blocks = (int)((long long)hn[0] * hn[1] < 8192 ? (long long)hn[0] * hn[1] : 8192);
        const int flat_blocks = dev_num_sms() * 8;

        DevBuf<float2> K0(Mk), W(Mk);
        DevBuf<float> d_lo(N);

        /* all-zero check of the caller's hires_density decides the direction (InitialConditions.c:619-631) */
        bool non_zero_input = false;
#pragma omp parallel for reduction(|| : non_zero_input)
        for (long long i = 0; i < M; i++)
            if (boxes->hires_density[i] != 0.f) non_zero_input = true;

        if (non_zero_input) {
            DevBuf<float> d_in(M);
            h2d(d_in, boxes->hires_density, M * sizeof(float));
            ScaleArgs sa = {(long long)hn[0] * hn[1], hn[2], nzc, reinterpret_cast<float *>(K0.p),
                            (float)(unsigned long long)M, d_in, VOLUME};
            B200_LAUNCH(scale_rows_kernel, row_blocks, 256, 0, sa);
            ZPrologue pro;
            fft_r2c(plan, K0, pro);
        } else {
            PsConsts ps;
            ps_export_consts(&ps);
            /* B200_IC_RNG: unset / "stream" = the reference's GSL mt19937 polar-Gaussian stream generated
               on the device in stream order (gslrng.cu; seed parity with N_THREADS = 1);
               "host" = the same stream from the sequential host generator (slow: ~26 ns per Gaussian);
               "device" = a counter-based Philox field with the same P(k) (no seed parity; benchmarks) */
            const char *mode = getenv("B200_IC_RNG");
            const bool device_rng = mode && strcmp(mode, "device") == 0;
            const bool host_rng = mode && strcmp(mode, "host") == 0;
            ModeArgs ma;
            memset(&ma, 0, sizeof(ma));
            ma.nx = hn[0]; ma.ny = hn[1]; ma.nz = hn[2]; ma.nzc = nzc_modes; ma.pitch = nzc;
            ma.dk[0] = dk[0]; ma.dk[1] = dk[1]; ma.dk[2] = dk[2];
            ma.volume = VOLUME; ma.ps = ps; ma.box = K0; ma.seed = random_seed;
            if (device_rng) {
                ma.x0 = 0; ma.nxs = hn[0]; ma.gauss = nullptr;
                B200_LAUNCH(ic_modes_kernel, flat_blocks, 256, 0, ma);
            } else {
                unsigned int seeds[1];
                hostnum::derive_thread_seeds(random_seed, 1, seeds);
                const long long plane = (long long)hn[1] * nzc_modes;
                /* x-slabs of up to 2^25 modes (2^26 Gaussians, 512 MB of doubles) */
                int slab = (int)(((long long)1 << (host_rng ? 24 : 25)) / plane);
                if (slab < 1) slab = 1;
                if (slab > hn[0]) slab = hn[0];
                DevBuf<double> d_g((size_t)slab * plane * 2);
                if (host_rng) {
                    hostnum::Mt19937 rng(seeds[0]);
                    std::vector<double> host_g((size_t)slab * plane * 2);
                    for (int x0 = 0; x0 < hn[0]; x0 += slab) {
                        const int nxs = (x0 + slab <= hn[0]) ? slab : hn[0] - x0;
                        const long long cnt = (long long)nxs * plane * 2;
                        for (long long i = 0; i < cnt; i++) host_g[i] = rng.ugaussian();
                        h2d(d_g, host_g.data(), cnt * sizeof(double));
                        ma.x0 = x0; ma.nxs = nxs; ma.gauss = d_g;
                        B200_LAUNCH(ic_modes_kernel, flat_blocks, 256, 0, ma);
                        dev_sync(); /* host_g is refilled next iteration */
                    }
                } else {
                    GslStream *gs = gsl_stream_create(seeds[0]);
                    try {
                        for (int x0 = 0; x0 < hn[0]; x0 += slab) {
                            const int nxs = (x0 + slab <= hn[0]) ? slab : hn[0] - x0;
                            gsl_stream_gaussians(gs, d_g, (long long)nxs * plane * 2);
                            ma.x0 = x0; ma.nxs = nxs; ma.gauss = d_g;
                            B200_LAUNCH(ic_modes_kernel, flat_blocks, 256, 0, ma);
                        }
                    } catch (...) {
                        gsl_stream_destroy(gs);
                        throw;
                    }
                    gsl_stream_destroy(gs);
                }
            }
            ConjArgs ca = {hn[0], hn[1], hn[2], nzc_modes, nzc, K0};
            B200_LAUNCH(ic_hermitian_kernel, 64, 256, 0, ca);
            /* hires_density = c2r(K0) / VOLUME  (InitialConditions.c:667-692) */
            DevBuf<float> d_hi(M);
            ZEpilogue epi;
            epi.scale = 1.f / VOLUME; epi.dst = d_hi; epi.dst_row_stride = hn[2];
            fft_c2r(plan, K0, W, KMul(), epi);
            d2h(boxes->hires_density, d_hi, M * sizeof(float));
        }

        KMul lowpass; /* top-hat at the low-res cell scale, only if the grids differ */
        lowpass.dk[0] = dk[0]; lowpass.dk[1] = dk[1]; lowpass.dk[2] = dk[2];
        if (so->DIM != so->HII_DIM) {
            lowpass.kind = KMUL_FILTER; lowpass.filter_type = 0;
            lowpass.R = (float)(pc::l_factor * so->BOX_LEN / (so->HII_DIM + 0.0));
        }
        auto to_lowres = [&](float *host_dst, float scale) {
            SubsampleArgs sa = {{ln[0], ln[1], ln[2]}, {hn[0], hn[1], hn[2]}, nzc, ratio,
                                reinterpret_cast<const float *>(W.p), d_lo, scale};
            B200_LAUNCH(subsample_kernel, flat_blocks, 256, 0, sa);
            d2h(host_dst, d_lo, N * sizeof(float));
        };
        ZEpilogue plain;

        /* lowres_density (InitialConditions.c:694-730) */
        fft_c2r(plan, K0, W, lowpass, plain);
        to_lowres(boxes->lowres_density, 1.f / VOLUME);

        /* Zel'dovich velocities (compute_velocity_fields, :299-364): low-pass filtered and sub-sampled
           onto the low-res grid, or (PERTURB_ON_HIGH_RES) unfiltered on the hi-res grid */
        const bool on_hires = mo->PERTURB_ON_HIGH_RES;
        DevBuf<float> d_hv(on_hires ? (size_t)M : 0);
        auto to_hires = [&](float *host_dst, const KMul &km, const float2 *kbox_src, float scale) {
            ZEpilogue e;
            e.scale = scale; e.dst = d_hv; e.dst_row_stride = hn[2];
            fft_c2r(plan, kbox_src, W, km, e);
            d2h(host_dst, d_hv, M * sizeof(float));
        };
        float *vel[3] = {on_hires ? boxes->hires_vx : boxes->lowres_vx, on_hires ? boxes->hires_vy : boxes->lowres_vy,
                         on_hires ? boxes->hires_vz : boxes->lowres_vz};
        for (int ax = 0; ax < 3; ax++) {
            if (!vel[ax]) b200_throw(B200_ValueError, "velocity array of the initial conditions is NULL");
            KMul km = on_hires ? KMul() : lowpass;
            km.dk[0] = dk[0]; km.dk[1] = dk[1]; km.dk[2] = dk[2];
            km.op = KOP_GRADIENT_D; km.axis_a = ax;
            if (on_hires) {
                to_hires(vel[ax], km, K0, 1.f / VOLUME);
            } else {
                fft_c2r(plan, K0, W, km, plain);
                to_lowres(vel[ax], 1.f / VOLUME);
            }
        }

        /* 2LPT (compute_velocity_fields_2LPT, :366-544) */
        if (mo->PERTURB_ALGORITHM == PERTURB_2LPT) {
            float *vel2[3] = {on_hires ? boxes->hires_vx_2LPT : boxes->lowres_vx_2LPT,
                              on_hires ? boxes->hires_vy_2LPT : boxes->lowres_vy_2LPT,
                              on_hires ? boxes->hires_vz_2LPT : boxes->lowres_vz_2LPT};
            float *scratch_out[3] = {boxes->hires_vx_2LPT, boxes->hires_vy_2LPT, boxes->hires_vz_2LPT};
            DevBuf<float> diag[3];
            DevBuf<float2> S(Mk);
            for (int c = 0; c < 3; c++) {
                diag[c].alloc(M);
                KMul km;
                km.dk[0] = dk[0]; km.dk[1] = dk[1]; km.dk[2] = dk[2];
                km.op = KOP_LAPLACIAN_D; km.axis_a = c; km.axis_b = c;
                ZEpilogue e;
                e.dst = diag[c]; e.dst_row_stride = hn[2];
                fft_c2r(plan, K0, W, km, e);
                /* the reference leaves phi_ii in the hires_v*_2LPT arrays it used as scratch */
                const char *skip = getenv("B200_SKIP_SCRATCH_OUTPUTS");
                if (scratch_out[c] && !on_hires && !(skip && skip[0] == '1')) d2h(scratch_out[c], diag[c], M * sizeof(float));
            }
            const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
            for (int c = 0; c < 3; c++) {
                KMul km;
                km.dk[0] = dk[0]; km.dk[1] = dk[1]; km.dk[2] = dk[2];
                km.op = KOP_LAPLACIAN_D; km.axis_a = pairs[c][0]; km.axis_b = pairs[c][1];
                fft_c2r(plan, K0, W, km, plain);
                Lpt2Args la = {(long long)hn[0] * hn[1], hn[2], nzc, diag[pairs[c][0]], diag[pairs[c][1]],
                               reinterpret_cast<const float *>(W.p), reinterpret_cast<float *>(S.p), c == 0};
                B200_LAUNCH(lpt2_source_kernel, row_blocks, 256, 0, la);
            }
            ScaleArgs sc = {(long long)hn[0] * hn[1], hn[2], nzc, reinterpret_cast<float *>(S.p),
                            VOLUME * VOLUME * (float)(unsigned long long)M, nullptr, 0.f};
            B200_LAUNCH(scale_rows_kernel, row_blocks, 256, 0, sc);
            ZPrologue pro;
            fft_r2c(plan, S, pro);
            for (int ax = 0; ax < 3; ax++) {
                if (!vel2[ax]) b200_throw(B200_ValueError, "2LPT velocity array of the initial conditions is NULL");
                KMul km = on_hires ? KMul() : lowpass;
                km.dk[0] = dk[0]; km.dk[1] = dk[1]; km.dk[2] = dk[2];
                km.op = KOP_GRADIENT_D; km.axis_a = ax;
                if (on_hires) {
                    to_hires(vel2[ax], km, S, 1.f);
                } else {
                    fft_c2r(plan, S, W, km, plain);
                    to_lowres(vel2[ax], 1.f);
                }
            }
        }
        g_stats.ms = timer.stop_ms();
    } catch (B200Error &e) {
        if (getenv("B200_VERBOSE") || e.code == B200_CUDAError)
            fprintf(stderr, "[21cmfast_b200] ComputeInitialConditions: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
