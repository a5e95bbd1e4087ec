/*
 * host_physics.cpp -- host scalar helpers (see host_physics.h).  Formulae and arithmetic types
 * follow the reference function cited above each block; the numerical methods come from
 * host_numerics.h.  Everything here is O(10^2..10^5) flops per Compute* call.
 */
#include "host_physics.h"

#include <omp.h>

#include <string>
#include <map>
#include <mutex>
#include <vector>

#include "host_numerics.h"

/* ------------------------------------------------------------------ global parameter pointers
 * InputParameters.c:11-89.  The structs are owned by Python; only cosmo_tables is deep-copied. */
extern "C" {
SimulationOptions *simulation_options_global = nullptr;
MatterOptions *matter_options_global = nullptr;
CosmoParams *cosmo_params_global = nullptr;
AstroParams *astro_params_global = nullptr;
AstroOptions *astro_options_global = nullptr;
CosmoTables *cosmo_tables_global = nullptr;
ConfigSettings config_settings = {2.0, false, nullptr, nullptr};
}
static bool g_tables_allocated = false;

static Table1D *copy_table(const Table1D *t) {
    if (!t || t->size <= 0 || !t->x_values || !t->y_values) return nullptr;
    Table1D *c = (Table1D *)malloc(sizeof(Table1D));
    c->size = t->size;
    c->x_values = (double *)malloc(sizeof(double) * t->size);
    c->y_values = (double *)malloc(sizeof(double) * t->size);
    memcpy(c->x_values, t->x_values, sizeof(double) * t->size);
    memcpy(c->y_values, t->y_values, sizeof(double) * t->size);
    return c;
}
static void free_table(Table1D *t) {
    if (!t) return;
    free(t->x_values);
    free(t->y_values);
    free(t);
}

extern "C" void Broadcast_struct_global_all(SimulationOptions *so, MatterOptions *mo,
                                            CosmoParams *cp, AstroParams *ap, AstroOptions *ao,
                                            CosmoTables *ct) {
    simulation_options_global = so;
    matter_options_global = mo;
    cosmo_params_global = cp;
    astro_params_global = ap;
    astro_options_global = ao;
    if (!g_tables_allocated && ct) {
        cosmo_tables_global = (CosmoTables *)calloc(1, sizeof(CosmoTables));
        cosmo_tables_global->ps_norm = ct->ps_norm;
        cosmo_tables_global->USE_SIGMA_8 = ct->USE_SIGMA_8;
        cosmo_tables_global->V_CB_AVG = ct->V_CB_AVG;
        if (mo && mo->POWER_SPECTRUM == 5) {
            cosmo_tables_global->transfer_density = copy_table(ct->transfer_density);
            cosmo_tables_global->transfer_vcb = copy_table(ct->transfer_vcb);
        }
        g_tables_allocated = true;
    }
}
extern "C" void Broadcast_struct_global_noastro(SimulationOptions *so, MatterOptions *mo,
                                                CosmoParams *cp) {
    simulation_options_global = so;
    matter_options_global = mo;
    cosmo_params_global = cp;
}
extern "C" void Free_cosmo_tables_global(void) {
    if (!g_tables_allocated) return;
    free_table(cosmo_tables_global->transfer_density);
    free_table(cosmo_tables_global->transfer_vcb);
    free(cosmo_tables_global);
    cosmo_tables_global = nullptr;
    g_tables_allocated = false;
}

void require_params(bool need_astro) {
    if (!simulation_options_global || !matter_options_global || !cosmo_params_global)
        b200_throw(B200_ValueError, "input structs were never broadcast");
    if (need_astro && (!astro_params_global || !astro_options_global || !cosmo_tables_global))
        b200_throw(B200_ValueError, "astro structs / cosmo tables were never broadcast");
}
int hii_d_para() {
    return (int)(simulation_options_global->NON_CUBIC_FACTOR * simulation_options_global->HII_DIM);
}
int d_para() {
    return (int)(simulation_options_global->NON_CUBIC_FACTOR * simulation_options_global->DIM);
}
double box_volume() { /* VOLUME macro, indexing.h:39-43 (float products, as in C) */
    const SimulationOptions *s = simulation_options_global;
    return s->BOX_LEN * s->BOX_LEN * s->NON_CUBIC_FACTOR * s->BOX_LEN;
}

#define CP cosmo_params_global
#define MO matter_options_global

/* ------------------------------------------------------------------ memo cache for host scalars
 * sigma_z0 / Nion_General / Fcoll_General are pure functions of the broadcast input structs and
 * their arguments; a coeval run asks for the same ~70 values at every call, each one an adaptive
 * quadrature.  Results are memoised under a key that hashes the *contents* of the input structs
 * (Python may re-use a pointer for new values), the function id and the argument bits. */
#include <mutex>
#include <unordered_map>
static std::unordered_map<unsigned long long, double> g_memo;
static std::mutex g_memo_lock;             /* two host threads may call the library at once */
static unsigned long long g_sigma_epoch = 0; /* bumped whenever the sigma(M) table is (re)built or freed: the
                                                memoised integrals read it when USE_INTERPOLATION_TABLES is on */
static unsigned long long fnv(const void *p, size_t n, unsigned long long h) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ULL;
    return h;
}
static unsigned long long params_signature(bool astro) {
    unsigned long long h = 1469598103934665603ULL;
    h = fnv(CP, sizeof(CosmoParams), h);
    h = fnv(MO, sizeof(MatterOptions), h);
    if (cosmo_tables_global) {
        h = fnv(&cosmo_tables_global->ps_norm, sizeof(double), h);
        h = fnv(&cosmo_tables_global->USE_SIGMA_8, sizeof(bool), h);
    }
    if (astro && astro_params_global) h = fnv(astro_params_global, sizeof(AstroParams), h);
    if (astro && astro_options_global) h = fnv(astro_options_global, sizeof(AstroOptions), h);
    return h;
}
template <typename F> static double memoised(int fn_id, bool astro, const double *args, int nargs, F &&compute) {
    unsigned long long h = params_signature(astro);
    h = fnv(&fn_id, sizeof(int), h);
    h = fnv(args, sizeof(double) * nargs, h);
    h = fnv(&g_sigma_epoch, sizeof(g_sigma_epoch), h);
    {
        std::lock_guard<std::mutex> guard(g_memo_lock);
        auto it = g_memo.find(h);
        if (it != g_memo.end()) return it->second;
    }
    const double v = compute(); /* outside the lock: it may recurse into other memoised scalars */
    std::lock_guard<std::mutex> guard(g_memo_lock);
    if (g_memo.size() > 100000) g_memo.clear();
    g_memo[h] = v;
    return v;
}

/* ------------------------------------------------------------------ background cosmology */
double hubble_H0() { return (double)(CP->hlittle * 3.2407e-18); } /* Constants.h:92 */
double rho_crit() {                                                /* Constants.h:94-98 */
    const double Ho = hubble_H0();
    return (3.0 * Ho * Ho / (8.0 * M_PI * pc::G)) *
           (pc::cm_per_Mpc * pc::cm_per_Mpc * pc::cm_per_Mpc) / pc::Msun;
}
double n_b0() { /* Constants.h:99-109 */
    const double Ho = hubble_H0();
    const double rc = 3.0 * Ho * Ho / (8.0 * M_PI * pc::G);
    const double No = rc * CP->OMb * (1 - CP->Y_He) / pc::m_p;
    const double He = rc * CP->OMb * CP->Y_He / (4.0 * pc::m_p);
    return No + He;
}
double MtoR(double M) { /* cosmology.c:612-622 */
    if (MO->FILTER == FILTER_TOPHAT)
        return pow(3 * M / (4 * M_PI * CP->OMm * rho_crit()), 1.0 / 3.0);
    if (MO->FILTER == FILTER_GAUSSIAN)
        return pow(M / (pow(2 * M_PI, 1.5) * CP->OMm * rho_crit()), 1.0 / 3.0);
    b200_throw(B200_ValueError, "No such filter = %d", MO->FILTER);
}
double RtoM(double R) { /* cosmology.c:625-635 */
    if (MO->FILTER == FILTER_TOPHAT)
        return (4.0 / 3.0) * M_PI * pow(R, 3) * (CP->OMm * rho_crit());
    if (MO->FILTER == FILTER_GAUSSIAN)
        return pow(2 * M_PI, 1.5) * CP->OMm * rho_crit() * pow(R, 3);
    b200_throw(B200_ValueError, "No such filter = %d", MO->FILTER);
}
double omega_mz(float z) { /* cosmology.c:638-642 */
    return CP->OMm * pow(1 + z, 3) /
           (CP->OMm * pow(1 + z, 3) + CP->OMl + CP->OMr * pow(1 + z, 4) + CP->OMk * pow(1 + z, 2));
}
static double deltac_nonlinear(float z) { /* Bryan & Norman 1998, cosmology.c:648-652 */
    const double d = omega_mz(z) - 1.0;
    return 18 * M_PI * M_PI + 82 * d - 39 * d * d;
}
double TtoM(double z, double T, double mu) { /* Barkana & Loeb 2001, cosmology.c:661-665 */
    return 7030.97 / (CP->hlittle) * sqrt(omega_mz(z) / (CP->OMm * deltac_nonlinear(z))) *
           pow(T / (mu * (1 + z)), 1.5);
}
extern "C" double atomic_cooling_threshold(float z) { return TtoM(z, 1e4, 0.59); }

extern "C" double dicke(double z) { /* cosmology.c:692-735 */
    const double tiny = 1e-4;
    if (fabs(CP->OMm - 1.0) < tiny) return 1.0 / (1.0 + z);
    if ((CP->OMl > (-tiny)) && (fabs(CP->OMl + CP->OMm + CP->OMr - 1.0) < 0.01) &&
        (fabs(CP->wl + 1.0) < tiny)) {
        /* flat LCDM + radiation, Liddle et al. (astro-ph/9512102) */
        const double omegaM_z = CP->OMm * pow(1 + z, 3) /
                                (CP->OMl + CP->OMm * pow(1 + z, 3) + CP->OMr * pow(1 + z, 4));
        const double dick_z =
            2.5 * omegaM_z / (1.0 / 70.0 + omegaM_z * (209 - omegaM_z) / 140.0 + pow(omegaM_z, 4.0 / 7.0));
        const double dick_0 =
            2.5 * CP->OMm / (1.0 / 70.0 + CP->OMm * (209 - CP->OMm) / 140.0 + pow(CP->OMm, 4.0 / 7.0));
        return dick_z / (dick_0 * (1.0 + z));
    }
    if ((CP->OMtot < (1 + tiny)) && (fabs(CP->OMl) < tiny)) { /* open, no Lambda (Peebles) */
        const double x_0 = 1.0 / (CP->OMm + 0.0) - 1.0;
        const double dick_0 = 1 + 3.0 / x_0 + 3 * log(sqrt(1 + x_0) - sqrt(x_0)) * sqrt(1 + x_0) / pow(x_0, 1.5);
        const double x = fabs(1.0 / (CP->OMm + 0.0) - 1.0) / (1 + z);
        const double dick_z = 1 + 3.0 / x + 3 * log(sqrt(1 + x) - sqrt(x)) * sqrt(1 + x) / pow(x, 1.5);
        return dick_z / dick_0;
    }
    b200_throw(B200_ValueError, "No growth function for these cosmological parameters");
}
double dtdz(float z) { /* cosmology.c:738-748 */
    const double x = sqrt(CP->OMl / CP->OMm) * pow(1 + z, -3.0 / 2.0);
    const double dxdz = sqrt(CP->OMl / CP->OMm) * pow(1 + z, -5.0 / 2.0) * (-3.0 / 2.0);
    const double const1 = 2 * sqrt(1 + CP->OMm / CP->OMl) / (3.0 * hubble_H0());
    const double numer = dxdz * (1 + x * pow(pow(x, 2) + 1, -0.5));
    const double denom = x + sqrt(pow(x, 2) + 1);
    return const1 * numer / denom;
}
double ddickedt(double z) { /* cosmology.c:751-757: one-sided difference with a *float* step */
    const float dz = 1e-10;
    return (dicke(z + dz) - dicke(z)) / dz / dtdz(z);
}
double hubble(float z) { /* cosmology.c:796-799 */
    return hubble_H0() * sqrt(CP->OMm * pow(1 + z, 3) + CP->OMr * pow(1 + z, 4) + CP->OMl);
}
double t_hubble(float z) { return 1.0 / hubble(z); }

/* ------------------------------------------------------------------ matter power spectrum */
static struct {
    double sound_horizon, alpha_nu, beta_c, omhh, f_nu, f_baryon, theta_cmb, sigma_norm;
    bool ready;
} cc = {0, 0, 0, 0, 0, 0, 0, 0, false};

static double tf_EH(double k) { /* Eisenstein & Hu 1999 (TFmdm), cosmology.c:52-71; N_nu = 1 */
    const double N_nu = 1.0;
    const double q = k * pow(cc.theta_cmb, 2) / cc.omhh;
    const double gamma_eff =
        sqrt(cc.alpha_nu) + (1.0 - sqrt(cc.alpha_nu)) / (1.0 + pow(0.43 * k * cc.sound_horizon, 4));
    const double q_eff = q / gamma_eff;
    double TF_m = log(M_E + 1.84 * cc.beta_c * sqrt(cc.alpha_nu) * q_eff);
    TF_m /= TF_m + pow(q_eff, 2) * (14.4 + 325.0 / (1.0 + 60.5 * pow(q_eff, 1.11)));
    const double q_nu = 3.92 * q / sqrt(cc.f_nu / N_nu);
    TF_m *= 1.0 + (1.2 * pow(cc.f_nu, 0.64) * pow(N_nu, 0.3 + 0.6 * cc.f_nu)) /
                      (pow(q_nu, -1.6) + pow(q_nu, 0.8));
    return TF_m;
}
static double tf_other(double k, int which) { /* cosmology.c:75-127 */
    if (which == 1) { /* BBKS + Sugiyama */
        const double gamma = CP->OMm * CP->hlittle * exp(-(CP->OMb) - (CP->OMb / CP->OMm));
        const double q = k / (CP->hlittle * gamma);
        return (log(1.0 + 2.34 * q) / (2.34 * q)) *
               pow(1.0 + 3.89 * q + pow(16.1 * q, 2) + pow(5.46 * q, 3) + pow(6.71 * q, 4), -0.25);
    }
    if (which == 2) { /* Efstathiou, Bond & White 1992 */
        const double gamma = CP->OMm * CP->hlittle * CP->hlittle;
        const double aa = 6.4 / gamma, bb = 3.0 / gamma, c2 = 1.7 / gamma, nu = 1.13;
        return pow(1 + pow(aa * k + pow(bb * k, 1.5) + pow(c2 * k, 2), nu), -1. / nu);
    }
    if (which == 3) { /* Peebles 1980 */
        const double gamma = CP->OMm * CP->hlittle * exp(-(CP->OMb) - (CP->OMb / CP->OMm));
        const double aa = 8.0 / (CP->hlittle * gamma), bb = 4.7 / pow(CP->hlittle * gamma, 2);
        return 1 + aa * k + bb * k * k;
    }
    if (which == 4) { /* White / DEFW 1985 */
        const double gamma =
            CP->OMm * CP->hlittle * CP->hlittle * exp(-(CP->OMb) - (CP->OMb / CP->OMm));
        const double aa = 1.7 / gamma, bb = 9.0 / pow(gamma, 1.5), c2 = 1.0 / pow(gamma, 2);
        return 139.284 / (1 + aa * k + bb * pow(k, 1.5) + c2 * k * k);
    }
    b200_throw(B200_ValueError,
               "POWER_SPECTRUM=%d (CLASS tables) is outside the scoped path; use EH", which);
}

/* CLASS transfer tables (transfer_function_CLASS, cosmology.c:130-213): natural cubic splines (gsl_interp_cspline)
   through the caller's T(k) samples of the matter density at z = 0 and, with V_CB_MODEL = FLUCTS, of the
   dark-matter/baryon relative velocity at kinematic decoupling; above the last sample the density follows
   Eisenstein & Hu scaled to the last sample, the velocity a log-log line through the last two. */
static struct {
    bool ready = false, has_vcb = false;
    std::vector<double> k, Tm, Tv;
    hostnum::CubicSpline dens, vcb;
    double eh_ratio_at_kmax = 0.;
    double *d_nodes = nullptr; /* device copy for the IC kernels: k, Tm, c_m, Tv, c_v (n each) */
} g_class;
static void class_free() {
    if (g_class.d_nodes) dev_free(g_class.d_nodes);
    g_class.d_nodes = nullptr;
    g_class.ready = false;
}
/* called when the device-side caches are released (b200_release_device_cache, also on a device switch): the spline
   nodes are uploaded again by the next ps_export_consts */
void ps_device_tables_drop() {
    if (g_class.d_nodes) dev_free(g_class.d_nodes);
    g_class.d_nodes = nullptr;
}
static void class_init() {
    const Table1D *td = cosmo_tables_global->transfer_density;
    if (!td || td->size < 3 || !td->x_values || !td->y_values)
        b200_throw(B200_ValueError, "POWER_SPECTRUM=CLASS needs cosmo_tables.transfer_density (k, T) samples");
    class_free();
    g_class.k.assign(td->x_values, td->x_values + td->size);
    g_class.Tm.assign(td->y_values, td->y_values + td->size);
    g_class.dens.init(g_class.k, g_class.Tm);
    const double kmax = g_class.k.back();
    g_class.eh_ratio_at_kmax = g_class.Tm.back() / kmax / kmax / tf_EH(kmax);
    g_class.has_vcb = MO->V_CB_MODEL == 2;
    if (g_class.has_vcb) {
        const Table1D *tv = cosmo_tables_global->transfer_vcb;
        if (!tv || tv->size != td->size || !tv->y_values)
            b200_throw(B200_ValueError, "V_CB_MODEL=FLUCTS needs cosmo_tables.transfer_vcb on the k samples of transfer_density");
        g_class.Tv.assign(tv->y_values, tv->y_values + tv->size);
        g_class.vcb.init(g_class.k, g_class.Tv);
    }
    g_class.ready = true;
}
static double tf_class(double k, int dv) {
    if (!g_class.ready || (dv == 1 && !g_class.has_vcb)) {
        /* tables were never broadcast / splined (Broadcast_struct_global_all copies them once, init_ps splines
           them): a number, not a crash, for the scalar entry points of the C surface */
        static bool said = false;
        if (!said) fprintf(stderr, "[21cmfast_b200] POWER_SPECTRUM=CLASS: the transfer tables are not initialised (Free_cosmo_tables_global, Broadcast_struct_global_all, init_ps)\n");
        said = true;
        return std::nan("");
    }
    const size_t n = g_class.k.size();
    if (k > g_class.k[n - 1]) {
        if (dv == 0) return g_class.eh_ratio_at_kmax * tf_EH(k) * k * k;
        const std::vector<double> &Tv = g_class.Tv, &kc = g_class.k;
        return exp(log(Tv[n - 1]) + (log(Tv[n - 1]) - log(Tv[n - 2])) / (log(kc[n - 1]) - log(kc[n - 2])) * (log(k) - log(kc[n - 1])));
    }
    return dv == 0 ? g_class.dens.eval(k) : g_class.vcb.eval(k);
}
/* average relative-velocity suppression of the matter power (cosmology.c:27-29,295-300) */
static const double KP_VCB_PM = 300.0, A_VCB_PM = 0.24, SIGMAK_VCB_PM = 0.9;

extern "C" double power_in_k(double k) { /* cosmology.c:278-303 */
    if (k == 0.) return 0.;
    double T;
    if (MO->POWER_SPECTRUM == 5) {
        T = tf_class(k, 0); /* CLASS convention: T = delta / zeta, no k^2 */
    } else {
        T = (MO->POWER_SPECTRUM == 0) ? tf_EH(k) : tf_other(k, MO->POWER_SPECTRUM);
        T *= k * k; /* non-CLASS transfer functions tend to 1 at k->0 */
    }
    const double primordial = cosmo_tables_global->ps_norm * pow(k / 0.05, CP->POWER_INDEX - 1.);
    double p = cc.sigma_norm * primordial * T * T / pow(k, 3);
    if (MO->POWER_SPECTRUM == 5 && MO->V_CB_MODEL != 0)
        p *= 1.0 - A_VCB_PM * exp(-pow(log(k / KP_VCB_PM), 2.0) / (2.0 * SIGMAK_VCB_PM * SIGMAK_VCB_PM));
    return p;
}
extern "C" double power_in_vcb(double k) { /* cosmology.c:310-332 */
    if (MO->POWER_SPECTRUM != 5 || !g_class.ready || !g_class.has_vcb) {
        fprintf(stderr, "[21cmfast_b200] power_in_vcb needs POWER_SPECTRUM=CLASS and V_CB_MODEL=FLUCTS\n");
        return std::nan("");
    }
    if (k == 0.) return 0.;
    const double T = tf_class(k, 1);
    const double primordial = cosmo_tables_global->ps_norm * pow(k / 0.05, CP->POWER_INDEX - 1.);
    return cc.sigma_norm * primordial * T * T / pow(k, 3);
}

static double window_of_kR(double kR, int filter) { /* filtering.c:18-46 */
    if (filter == 0) {
        if (kR < 1e-4) return 1 - kR * kR / 10;
        return 3.0 * pow(kR, -3) * (sin(kR) - cos(kR) * kR);
    }
    if (filter == 1) return (kR * 0.413566994 > 1) ? 0. : 1.;
    if (filter == 2) return exp(-0.643 * 0.643 * (kR * kR) / 2.);
    b200_throw(B200_ValueError, "No such filter: %d", filter);
}
static double dw2dm_of_k(double k, double R, int filter) { /* filtering.c:49-78 */
    const double kR = k * R;
    double w, dwdr, drdm;
    if (filter == 0) {
        w = (kR < 1.0e-4) ? 1.0 : 3.0 * (sin(kR) / pow(kR, 3) - cos(kR) / pow(kR, 2));
        dwdr = (kR < 1.0e-10) ? 0
                              : 9 * cos(kR) * k / pow(kR, 3) + 3 * sin(kR) * (1 - 3 / (kR * kR)) / (kR * R);
        drdm = 1.0 / (4.0 * M_PI * CP->OMm * rho_crit() * R * R);
    } else if (filter == 2) {
        w = exp(-kR * kR / 2.0);
        dwdr = -k * kR * w;
        drdm = 1.0 / (pow(2 * M_PI, 1.5) * CP->OMm * rho_crit() * 3 * R * R);
    } else {
        b200_throw(B200_ValueError, "No such filter for dWdM: %d", filter);
    }
    return 2 * w * dwdr * drdm;
}

static double sigma_z0_compute(double M);
extern "C" double sigma_z0(double M) {
    if (omp_in_parallel()) return sigma_z0_compute(M); /* the memo map is not thread-safe */
    return memoised(1, false, &M, 1, [&] { return sigma_z0_compute(M); });
}
static double sigma_z0_compute(double M) { /* cosmology.c:369-404 */
    const double R = MtoR(M);
    const int filt = MO->FILTER;
    double res, err;
    auto f = [&](double k) {
        const double w = window_of_kR(k * R, filt);
        return k * k * power_in_k(k) * w * w / (2.0 * M_PI * M_PI);
    };
    const int st = hostnum::qag61(f, 1.0e-99 / R, 350.0 / R, 0, pc::FRACT_FLOAT_ERR * 10, 1000, &res, &err);
    if (st != 0) b200_throw(B200_GSLError, "sigma_z0(M=%g): quadrature status %d", M, st);
    return sqrt(res);
}
extern "C" double dsigmasqdm_z0(double M) { /* cosmology.c:421-456 */
    const double R = MtoR(M);
    const int filt = MO->FILTER;
    double res, err;
    auto f = [&](double k) { return k * k * power_in_k(k) * dw2dm_of_k(k, R, filt) / (2.0 * M_PI * M_PI); };
    const int st = hostnum::qag61(f, 1.0e-99 / R, 350.0 / R, 0, pc::FRACT_FLOAT_ERR * 10, 1000, &res, &err);
    if (st != 0) b200_throw(B200_GSLError, "dsigmasqdm_z0(M=%g): quadrature status %d", M, st);
    return res;
}

static void init_ps_impl() { /* cosmology.c:459-557 */
    require_params(false);
    if (!cosmo_tables_global) b200_throw(B200_ValueError, "cosmo tables were never broadcast");
    cc.omhh = CP->OMm * CP->hlittle * CP->hlittle;
    cc.theta_cmb = pc::T_cmb / 2.7;
    cc.f_nu = fmax(CP->OMn / CP->OMm, 1e-10);
    cc.f_baryon = fmax(CP->OMb / CP->OMm, 1e-10);
    { /* TFset_parameters */
        const double f_nu = cc.f_nu, f_b = cc.f_baryon, omhh = cc.omhh, theta_cmb = cc.theta_cmb;
        const double obhh = CP->OMb * CP->hlittle * CP->hlittle;
        const double z_equality = 25000 * omhh * pow(theta_cmb, -4) - 1.0;
        const double k_equality = 0.0746 * omhh / (theta_cmb * theta_cmb);
        double z_drag = 0.313 * pow(omhh, -0.419) * (1 + 0.607 * pow(omhh, 0.674));
        z_drag = 1 + z_drag * pow(obhh, 0.238 * pow(omhh, 0.223));
        z_drag *= 1291 * pow(omhh, 0.251) / (1 + 0.659 * pow(omhh, 0.828));
        const double y_d = (1 + z_equality) / (1.0 + z_drag);
        const double R_drag = 31.5 * obhh * pow(theta_cmb, -4) * 1000 / (1.0 + z_drag);
        const double R_equality = 31.5 * obhh * pow(theta_cmb, -4) * 1000 / (1.0 + z_equality);
        cc.sound_horizon = 2.0 / 3.0 / k_equality * sqrt(6.0 / R_equality) *
                           log((sqrt(1 + R_drag) + sqrt(R_drag + R_equality)) / (1.0 + sqrt(R_equality)));
        const double p_c = -(5 - sqrt(1 + 24 * (1 - f_nu - f_b))) / 4.0;
        const double p_cb = -(5 - sqrt(1 + 24 * (1 - f_nu))) / 4.0;
        const double f_c = 1 - f_nu - f_b, f_cb = 1 - f_nu, f_nub = f_nu + f_b;
        double alpha_nu = (f_c / f_cb) * (2 * (p_c + p_cb) + 5) / (4 * p_cb + 5.0);
        alpha_nu *= 1 - 0.553 * f_nub + 0.126 * pow(f_nub, 3);
        alpha_nu /= 1 - 0.193 * sqrt(f_nu) + 0.169 * f_nu;
        alpha_nu *= pow(1 + y_d, p_c - p_cb);
        alpha_nu *= 1 + (p_cb - p_c) / 2.0 * (1.0 + 1.0 / (4.0 * p_c + 3.0) / (4.0 * p_cb + 7.0)) / (1.0 + y_d);
        cc.alpha_nu = alpha_nu;
        cc.beta_c = 1.0 / (1.0 - 0.949 * f_nub);
    }
    if (MO->POWER_SPECTRUM == 5) class_init(); /* after TFset_parameters: E&H extrapolates the tables (cosmology.c:515-524) */
    if (cosmo_tables_global->USE_SIGMA_8) {
        const double Radius_8 = 8.0 / CP->hlittle;
        cc.sigma_norm = 1;
        cc.sigma_norm = pow(cosmo_tables_global->ps_norm / sigma_z0_compute(RtoM(Radius_8)), 2);
    } else {
        cc.sigma_norm = 2.0 * M_PI * M_PI;
    }
    cc.ready = true;
    g_memo.clear();
}
extern "C" void init_ps(void) {
    try { init_ps_impl(); } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] init_ps failed: %s\n", e.msg);
    }
}
extern "C" void free_ps(void) { cc.ready = false; class_free(); }

/* constants the device-side power spectrum (ics.cu) needs */
struct PsConsts {
    int which;
    double sound_horizon, alpha_nu, beta_c, omhh, f_nu, theta_cmb, sigma_norm;
    double ps_norm, n_s, h, OMm, OMb;
    /* CLASS tables: n spline nodes each (device pointers), see g_class */
    int n_class, vcb_suppression;
    const double *ck, *cTm, *cCm, *cTv, *cCv;
    double eh_ratio_at_kmax;
};
void ps_export_consts(PsConsts *o) {
    if (!cc.ready) init_ps_impl();
    o->n_class = 0; o->vcb_suppression = 0;
    o->ck = o->cTm = o->cCm = o->cTv = o->cCv = nullptr;
    o->eh_ratio_at_kmax = 0.;
    if (MO->POWER_SPECTRUM == 5) {
        if (!g_class.ready) b200_throw(B200_ValueError, "POWER_SPECTRUM=CLASS: the transfer tables are not initialised");
        const size_t n = g_class.k.size();
        if (!g_class.d_nodes) {
            std::vector<double> h(5 * n, 0.0);
            for (size_t i = 0; i < n; i++) {
                h[i] = g_class.k[i]; h[n + i] = g_class.Tm[i]; h[2 * n + i] = g_class.dens.coeffs()[i];
                if (g_class.has_vcb) { h[3 * n + i] = g_class.Tv[i]; h[4 * n + i] = g_class.vcb.coeffs()[i]; }
            }
            g_class.d_nodes = (double *)dev_alloc(5 * n * sizeof(double));
            h2d(g_class.d_nodes, h.data(), 5 * n * sizeof(double));
            dev_sync();
        }
        o->n_class = (int)n;
        o->ck = g_class.d_nodes; o->cTm = o->ck + n; o->cCm = o->ck + 2 * n;
        o->cTv = g_class.has_vcb ? o->ck + 3 * n : nullptr; o->cCv = g_class.has_vcb ? o->ck + 4 * n : nullptr;
        o->eh_ratio_at_kmax = g_class.eh_ratio_at_kmax;
        o->vcb_suppression = MO->V_CB_MODEL != 0;
    }
    o->which = MO->POWER_SPECTRUM;
    o->sound_horizon = cc.sound_horizon; o->alpha_nu = cc.alpha_nu; o->beta_c = cc.beta_c;
    o->omhh = cc.omhh; o->f_nu = cc.f_nu; o->theta_cmb = cc.theta_cmb; o->sigma_norm = cc.sigma_norm;
    o->ps_norm = cosmo_tables_global->ps_norm; o->n_s = CP->POWER_INDEX;
    o->h = CP->hlittle; o->OMm = CP->OMm; o->OMb = CP->OMb;
}

/* ------------------------------------------------------------------ sigma(M) tables
 * interp_tables.c:1135-1185: 300 log-spaced masses, float tables, float mass argument. */
#define N_MASS_INTERP 300
static struct {
    bool ready;
    double x_min, x_width;
    float sigma[N_MASS_INTERP], dsig[N_MASS_INTERP];
} st = {false, 0, 0, {0}, {0}};

static double eval_table_f(double x, double x_min, double x_width, const float *y) {
    /* EvaluateRGTable1D_f, interpolation.c:123-131 (bin edge through a float cast of idx) */
    const int idx = (int)floor((x - x_min) / x_width);
    const double table_val = x_min + x_width * (float)idx;
    const double t = (x - table_val) / x_width;
    return y[idx] * (1 - t) + y[idx + 1] * t;
}

extern "C" void initialiseSigmaMInterpTable(float M_min, float M_max) {
    try {
        require_params(false);
        if (!cc.ready) init_ps_impl();
        st.x_min = log(M_min);
        st.x_width = (log(M_max) - log(M_min)) / (N_MASS_INTERP - 1.);
        int fail = 0;
        int nt = simulation_options_global->N_THREADS;
        if (nt < 1) nt = 1;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 4)
        for (int i = 0; i < N_MASS_INTERP; i++) {
            try {
                const float Mass = exp(st.x_min + i * st.x_width);
                st.sigma[i] = sigma_z0(Mass);
                st.dsig[i] = log10(-dsigmasqdm_z0(Mass));
                if (!std::isfinite(st.sigma[i]) || !std::isfinite(st.dsig[i])) fail = 1;
            } catch (B200Error &) { fail = 1; }
        }
        if (fail) b200_throw(B200_TableGenerationError, "sigma(M) table has non-finite entries");
        st.ready = true;
        g_sigma_epoch++;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] initialiseSigmaMInterpTable failed: %s\n", e.msg);
    }
}
extern "C" void freeSigmaMInterpTable(void) { st.ready = false; g_sigma_epoch++; }

double EvaluateSigma(double lnM) { /* interp_tables.c:1171-1177 */
    if (MO->USE_INTERPOLATION_TABLES != 0) {
        if (!st.ready) b200_throw(B200_TableEvaluationError, "sigma table not initialised");
        return eval_table_f(lnM, st.x_min, st.x_width, st.sigma);
    }
    return sigma_z0(exp(lnM));
}
double EvaluatedSigmasqdm(double lnM) { /* interp_tables.c:1179-1185 */
    if (MO->USE_INTERPOLATION_TABLES != 0) {
        if (!st.ready) b200_throw(B200_TableEvaluationError, "sigma table not initialised");
        return -pow(10., eval_table_f(lnM, st.x_min, st.x_width, st.dsig));
    }
    return dsigmasqdm_z0(exp(lnM));
}

/* ------------------------------------------------------------------ scaling relations */
static float mass_limit(float logM, float PL, float FRAC) { /* hmf.c:1268 */
    return FRAC * pow(pow(10., logM) / 1e10, PL);
}
static float mass_limit_bisection(float Mmin, float Mmax, float PL, float FRAC) { /* hmf.c:1274-1315 */
    int iter = 0;
    const int max_iter = 200;
    const float rel_tol = 0.001;
    float logMlow = log10(Mmin), logMupper = log10(Mmax), x, x1;
    if (PL < 0.) {
        if (mass_limit(logMlow, PL, FRAC) <= 1.) return Mmin;
    } else if (PL > 0.) {
        if (mass_limit(logMupper, PL, FRAC) <= 1.) return Mmax;
    } else
        return 0;
    x = (logMlow + logMupper) / 2.; ++iter;
    do {
        if ((mass_limit(logMlow, PL, FRAC) - 1.) * (mass_limit(x, PL, FRAC) - 1.) < 0.)
            logMupper = x;
        else
            logMlow = x;
        x1 = (logMlow + logMupper) / 2.; ++iter;
        if (fabs(x1 - x) < rel_tol) return pow(10., x1);
        x = x1;
    } while (iter < max_iter);
    b200_throw(B200_MassDepZetaError, "no mass limit keeps the stellar/escape fraction below 1");
}

void set_scaling_constants(double redshift, ScalingConstants *sc) { /* scaling_relations.c:36-115 */
    const AstroParams *ap = astro_params_global;
    sc->redshift = redshift;
    sc->fstar_10 = ap->F_STAR10;
    sc->alpha_star = ap->ALPHA_STAR;
    sc->t_h = t_hubble(redshift);
    sc->t_star = ap->t_STAR;
    sc->alpha_esc = ap->ALPHA_ESC;
    sc->fesc_10 = ap->F_ESC10;
    sc->pop2_ion = ap->POP2_ION;
    sc->mturn_a_nofb = ap->M_TURN;
    sc->Mlim_Fstar = mass_limit_bisection(pc::M_MIN_INTEGRAL, pc::M_MAX_INTEGRAL, sc->alpha_star, sc->fstar_10);
    sc->Mlim_Fesc = mass_limit_bisection(pc::M_MIN_INTEGRAL, pc::M_MAX_INTEGRAL, sc->alpha_esc, sc->fesc_10);
}

extern "C" double minimum_source_mass(double redshift, bool xray) { /* hmf.c:1319-1348 */
    const double min_factor =
        (MO->SOURCE_MODEL != SRC_CONST_ION_EFF && !astro_options_global->USE_MINI_HALOS) ? 50. : 1.;
    double Mmin;
    if (astro_options_global->USE_MINI_HALOS) {
        Mmin = pc::M_MIN_INTEGRAL;
    } else if (astro_options_global->M_MIN_in_Mass) {
        Mmin = astro_params_global->M_TURN;
    } else {
        const double t_vir_min = xray ? astro_params_global->X_RAY_Tvir_MIN : astro_params_global->ION_Tvir_MIN;
        const double mu_factor = t_vir_min < 9.99999e3 ? 1.22 : 0.6;
        Mmin = TtoM(redshift, t_vir_min, mu_factor);
    }
    return Mmin / min_factor;
}

/* ------------------------------------------------------------------ mass functions (1/rho_bar units) */
static double sheth_delc_fixed(double del, double sig) { /* hmf.c:151-154 (Jenkins a,b,c) */
    return sqrt(0.73) * del * (1. + 0.34 * pow(sig * sig / (0.73 * del * del), 0.81));
}
extern "C" double get_delta_crit(int HMF, double sigma, double growthf) { /* hmf.c:166-171 */
    if (HMF == HMF_DELOS) return pc::delta_c_delos;
    if (HMF == HMF_ST) return sheth_delc_fixed(pc::delta_c_sph / growthf, sigma) * growthf;
    return pc::delta_c_sph;
}

static double umf(double growthf, double lnM, int HMF, double z) { /* hmf.c:553-580 */
    double sigma = EvaluateSigma(lnM), dsigmadm = EvaluatedSigmasqdm(lnM);
    if (HMF == HMF_PS) { /* hmf.c:345-355 */
        sigma = sigma * growthf;
        dsigmadm = dsigmadm * (growthf * growthf / (2. * sigma));
        return -sqrt(2 / M_PI) * (pc::delta_c_sph / (sigma * sigma)) * dsigmadm *
               exp(-(pc::delta_c_sph * pc::delta_c_sph) / (2 * sigma * sigma));
    }
    if (HMF == HMF_ST) { /* hmf.c:301-313, Jenkins-fit a=0.73 p=0.175 A=0.353 */
        sigma = sigma * growthf;
        dsigmadm = dsigmadm * (growthf * growthf / (2. * sigma));
        const double nuhat = sqrt(0.73) * pc::delta_c_sph / sigma;
        return -(dsigmadm / sigma) * sqrt(2. / M_PI) * 0.353 * (1 + pow(nuhat, -2 * 0.175)) * nuhat *
               exp(-nuhat * nuhat / 2.0);
    }
    if (HMF == HMF_DELOS) { /* hmf.c:188-207 */
        const double sigma_inv = 1 / sigma;
        const double dsdm = dsigmadm * (0.5 * sigma_inv);
        const double nu = pc::delta_c_delos * sigma_inv / growthf;
        const double dfdnu = 0.519 * pow(nu, 0.582) * exp(-0.469 * nu * nu);
        return dfdnu * fabs(dsdm) * sigma_inv;
    }
    if (HMF == HMF_WATSON || HMF == HMF_WATSON_Z) { /* Watson et al. 2013 FOF fits, hmf.c:370-417 */
        sigma = sigma * growthf;
        dsigmadm = dsigmadm * (growthf * growthf / (2. * sigma));
        double A = 0.282, alpha = 2.163, beta = 1.406, gamma = 1.210;
        if (HMF == HMF_WATSON_Z) { /* their eqs. 12-15: the fit parameters follow Omega_m(z) */
            const double Om_z = CP->OMm * pow(1. + z, 3.) / (CP->OMl + CP->OMm * pow(1. + z, 3.) + CP->OMr * pow(1. + z, 4.));
            A = Om_z * (0.990 * pow(1. + z, -3.216) + 0.074);
            alpha = Om_z * (5.907 * pow(1. + z, -3.058) + 2.349);
            beta = Om_z * (3.136 * pow(1. + z, -3.599) + 2.344);
            gamma = 1.318;
        }
        const double f_sigma = A * (pow(beta / sigma, alpha) + 1.) * exp(-gamma / (sigma * sigma));
        return -(dsigmadm / sigma) * f_sigma;
    }
    if (HMF == HMF_REED07) { /* Reed et al. 2007 (astro-ph/0607150) fit with its n_eff term, hmf.c:156-163,419-439 */
        const double neff = -3. * (2. * (-exp(lnM) * dsigmadm / (2. * sigma * sigma)) + 1.);
        const double sigma_z = sigma * growthf;
        dsigmadm = dsigmadm * (growthf * growthf / (2. * sigma_z));
        const double nu = pc::delta_c_sph / sigma_z, lnsigma = -log(sigma_z);
        const double G1 = exp(-pow(lnsigma - 0.4, 2) / (2. * 0.6 * 0.6)), G2 = exp(-pow(lnsigma - 0.75, 2) / (2. * 0.2 * 0.2));
        const double ac = 0.764 / 1.08;
        const double f_sigma = 0.3222 * sqrt(2. * ac / M_PI) * (1. + pow(1. / (ac * nu * nu), 0.3) + 0.6 * G1 + 0.4 * G2) * nu *
                               exp(-1.08 * ac * nu * nu / 2. - 0.03 * pow(nu, 0.6) / pow(neff + 3., 2));
        return -(dsigmadm / sigma_z) * f_sigma;
    }
    if (HMF == HMF_YUNG24) { /* Yung et al. 2024 (arXiv:2304.04348): Watson form, parameters quadratic in z; hmf.c:441-459 */
        const double sigma_z = sigma * growthf;
        dsigmadm = dsigmadm * (growthf * growthf / (2. * sigma_z));
        const double A_z = 0.13765772 + -0.01003821 * z + 0.00102964 * z * z;
        const double a_z = 1.06641384 + 0.02475576 * z + -0.00283342 * z * z;
        const double b_z = 4.86693806 + 0.09212356 * z + -0.01426283 * z * z;
        const double c_z = 1.19837952 + -0.00142967 * z + -0.00033074 * z * z;
        const double f_sigma = A_z * (pow(sigma_z / b_z, -a_z) + 1.) * exp(-c_z / (sigma_z * sigma_z));
        return -(dsigmadm / sigma_z) * f_sigma;
    }
    b200_throw(B200_ValueError, "invalid HMF %d", HMF);
}

static double st_taylor_factor(double sig, double sig_cond, double growthf, double *zeroth) {
    /* Sheth & Tormen 2002 moving-barrier Taylor series, hmf.c:234-267 */
    const double a = 0.73, alpha = 0.81, beta = 0.34;
    const double del = pc::delta_c_sph / growthf;
    const double sigsq = sig * sig, sigsq_inv = 1. / sigsq, sigcsq = sig_cond * sig_cond;
    const double sigdiff = sig == sig_cond ? 1e-6 : sigsq - sigcsq;
    double t[6];
    t[0] = 1.;
    for (int i = 1; i < 6; i++) t[i] = t[i - 1] * (-sigdiff) / i * (alpha - i + 1) * sigsq_inv;
    double result = 0.;
    for (int i = 5; i >= 0; i--) result += t[i];
    const double pre1 = sqrt(a) * del;
    const double pre2 = beta * pow(sigsq_inv * (a * del * del), -alpha);
    result = pre1 * (1 + pre2 * result);
    *zeroth = pre1 * (1 + pre2);
    return result;
}
static double cmf(double growthf, double lnM, double delta_cond, double sigma_cond, int HMF) {
    /* conditional_hmf, hmf.c:511-525 */
    const double sigma1 = EvaluateSigma(lnM);
    const double dsigmasqdm = EvaluatedSigmasqdm(lnM);
    if (HMF == HMF_ST) { /* hmf.c:270-285 */
        if (sigma1 < sigma_cond) return 0.;
        const double delta_0 = delta_cond / growthf;
        double Barrier;
        const double factor = st_taylor_factor(sigma1, sigma_cond, growthf, &Barrier) - delta_0;
        const double sdi = sigma1 == sigma_cond ? 1e6 : 1 / (sigma1 * sigma1 - sigma_cond * sigma_cond);
        return -dsigmasqdm * factor * pow(sdi, 1.5) *
               exp(-(Barrier - delta_0) * (Barrier - delta_0) * 0.5 * (sdi)) / sqrt(2. * M_PI);
    }
    if (HMF == HMF_DELOS) { /* hmf.c:209-230 */
        if (sigma1 < sigma_cond) return 0.;
        const double dsdm = dsigmasqdm * 0.5;
        const double sdi = sigma1 == sigma_cond ? 1e6 : 1 / (sigma1 * sigma1 - sigma_cond * sigma_cond);
        const double nu = (pc::delta_c_delos - delta_cond) * sqrt(sdi) / growthf;
        const double dfdnu = 0.519 * pow(nu, 0.582) * exp(-0.469 * nu * nu);
        return dfdnu * fabs(dsdm) * sdi;
    }
    /* EPS, hmf.c:317-331 (also the fallback for HMFs without a CMF) */
    if (sigma1 < sigma_cond) return 0.;
    const double sdi = sigma1 == sigma_cond ? 1e6 : 1 / (sigma1 * sigma1 - sigma_cond * sigma_cond);
    const double del = (pc::delta_c_sph - delta_cond) / growthf;
    return -del * dsigmasqdm * pow(sdi, 1.5) * exp(-del * del * 0.5 * sdi) / sqrt(2. * M_PI);
}

struct MFParams {
    double redshift; /* read by the z-dependent fits (WATSON-Z, YUNG24) */
    double growthf;
    int HMF;
    double sigma_cond, delta;
    double Mturn, f_star_norm, alpha_star, Mlim_star, f_esc_norm, alpha_esc, Mlim_esc;
};
static double log_pl_limit(double lnM, double ln_norm, double alpha, double ln_pivot, double ln_limit) {
    /* scaling_relations.c:211-231 */
    if ((alpha > 0. && lnM > ln_limit) || (alpha < 0. && lnM < ln_limit)) return -ln_norm;
    return alpha * (lnM - ln_pivot);
}
static double nion_fraction(double lnM, const MFParams &p) { /* hmf.c:462-468 */
    const double Fstar = log_pl_limit(lnM, p.f_star_norm, p.alpha_star, 10 * M_LN10, p.Mlim_star);
    const double Fesc = log_pl_limit(lnM, p.f_esc_norm, p.alpha_esc, 10 * M_LN10, p.Mlim_esc);
    return exp(Fstar + Fesc - p.Mturn / exp(lnM) + lnM);
}

static double integrate_qag(double lo, double hi, const MFParams &p, int which) {
    /* IntegratedNdM_QAG, hmf.c:612-653: rel tol 1e-3, 61-point rule, 1000 panels */
    double res, err;
    auto f = [&](double lnM) {
        switch (which) {
            case 0: return exp(lnM) * umf(p.growthf, lnM, p.HMF, p.redshift);                /* u_fcoll */
            case 1: return nion_fraction(lnM, p) * umf(p.growthf, lnM, p.HMF, p.redshift);  /* u_nion  */
            default:
                return nion_fraction(lnM, p) * cmf(p.growthf, lnM, p.delta, p.sigma_cond, p.HMF);
        }
    };
    const int st_ = hostnum::qag61(f, lo, hi, 0, 1e-3, 1000, &res, &err);
    if (st_ != 0) b200_throw(B200_GSLError, "mass-function quadrature status %d", st_);
    return res;
}

#define NGL_INT 100
static double xi_GL[NGL_INT + 1], wi_GL[NGL_INT + 1], GL_limit[2] = {0, 0};
extern "C" void initialise_GL(double lnM_Min, double lnM_Max) { /* hmf.c:699-706 */
    if (lnM_Min == GL_limit[0] && lnM_Max == GL_limit[1]) return;
    hostnum::gauss_legendre(lnM_Min, lnM_Max, NGL_INT, xi_GL, wi_GL);
    GL_limit[0] = lnM_Min;
    GL_limit[1] = lnM_Max;
}
static double integrate_gl(double lo, double hi, const MFParams &p) { /* hmf.c:710-725 */
    if ((float)lo != (float)GL_limit[0] || (float)hi != (float)GL_limit[1])
        b200_throw(B200_TableGenerationError, "integral limits do not match the Gauss-Legendre nodes");
    double integral = 0;
    for (int i = 1; i < NGL_INT + 1; i++)
        integral += wi_GL[i] * (nion_fraction(xi_GL[i], p) * cmf(p.growthf, xi_GL[i], p.delta, p.sigma_cond, p.HMF));
    return integral;
}

static MFParams mf_params(double growthf, double Mturn, const ScalingConstants *sc) {
    MFParams p;
    memset(&p, 0, sizeof(p));
    p.growthf = growthf;
    p.HMF = MO->HMF;
    p.Mturn = Mturn;
    if (sc) {
        p.alpha_star = sc->alpha_star;
        p.alpha_esc = sc->alpha_esc;
        p.f_star_norm = log(sc->fstar_10);
        p.f_esc_norm = log(sc->fesc_10);
        p.Mlim_star = log(sc->Mlim_Fstar);
        p.Mlim_esc = log(sc->Mlim_Fesc);
    }
    return p;
}

double Fcoll_General(double z, double lnMmin, double lnMmax) { /* hmf.c:945-953 */
    const double args[3] = {z, lnMmin, lnMmax};
    return memoised(2, true, args, 3, [&] {
        MFParams p = mf_params(dicke(z), 0., nullptr);
        p.redshift = z;
        return integrate_qag(lnMmin, lnMmax, p, 0);
    });
}
double Nion_General(double z, double lnMmin, double lnMmax, double Mturn, const ScalingConstants *sc) {
    /* the scaling constants are part of the key: callers also pass modified copies (evolve_scaling_constants_sfr) */
    const double args[10] = {z, lnMmin, lnMmax, Mturn, sc->fstar_10, sc->alpha_star, sc->Mlim_Fstar,
                             sc->fesc_10, sc->alpha_esc, sc->Mlim_Fesc};
    return memoised(3, true, args, 10, [&] {
        MFParams p = mf_params(dicke(z), Mturn, sc); /* hmf.c:955-971 */
        p.redshift = z;
        return integrate_qag(lnMmin, lnMmax, p, 1);
    });
}
/* ---- INTEGRATION_METHOD_ATOMIC = GAMMA-APPROX (Munoz et al. 2022, appendix B; hmf.c:728-892) ----
   sigma(M) is taken as a triple power law around two pivot masses and the turnover as a sharp cut, which
   turns the conditional Press-Schechter integral of M^(alpha_star + alpha_esc) into differences of
   upper incomplete gamma functions Gamma(1/2 + beta, nu/2). */
static double expint_e1(double x) { /* E1(x) = Gamma(0, x) */
    if (x <= 1.0) {
        double sum = 0, term = 1;
        for (int k = 1; k < 60; k++) { term *= -x / k; sum -= term / k; }
        return -0.5772156649015328606 - log(x) + sum;
    }
    double b = x + 1.0, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 200; i++) { /* modified Lentz */
        const double an = -1.0 * i * i;
        b += 2.0; d = 1.0 / (an * d + b); c = b + an / c;
        const double del = c * d; h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h * exp(-x);
}
static double upper_gamma_pos(double a, double x) { /* Gamma(a, x), a > 0 */
    if (x <= 0) return tgamma(a);
    if (x < a + 1.0) { /* series of the lower function */
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 500; n++) { ap += 1; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-16) break; }
        return tgamma(a) - sum * exp(-x + a * log(x));
    }
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d; /* continued fraction */
    for (int i = 1; i < 500; i++) {
        const double an = -i * (i - a);
        b += 2.0; d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d; const double del = d * c; h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return exp(-x + a * log(x)) * h;
}
/* Legendre continued fraction of Gamma(a, x) by modified Lentz: converges for every real a when x > 0
   (slowly below x ~ 1/4, where the recurrence is used instead, as gsl_sf_gamma_inc does) */
static double upper_gamma_cf(double a, double x) {
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 5000; i++) {
        const double an = -i * (i - a);
        b += 2.0; d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d; const double del = d * c; h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return exp(-x + a * log(x)) * h;
}
static double upper_gamma(double a, double x) { /* any real a (gsl_sf_gamma_inc at hmf.c:733) */
    if (a > 0) return upper_gamma_pos(a, x);
    if (x <= 0) return INFINITY;
    /* the downward recurrence from (0, 1] cancels catastrophically for large x (rel. error 3.5e-4 at
       a = -4.75, x = 300): the continued fraction is exact there */
    if (x > 0.25) return a == 0.0 ? expint_e1(x) : upper_gamma_cf(a, x);
    /* start in (0, 1] (E1 for an integer a) and recur down: Gamma(a, x) = (Gamma(a + 1, x) - x^a e^-x) / a */
    const double fa = a - floor(a);
    double g, acur;
    if (fa == 0.0) { g = expint_e1(x); acur = 0.0; }
    else { g = upper_gamma_pos(fa, x); acur = fa; }
    while (acur > a + 0.5) {
        acur -= 1.0;
        g = (g - pow(x, acur) * exp(-x)) / acur;
    }
    return g;
}
extern "C" double b200_upper_gamma(double a, double x) { return upper_gamma(a, x); } /* test hook */
static double fcoll_approx(double numin, double beta) { /* int nu^beta exp(-nu/2) / sqrt(nu) dnu from numin */
    return upper_gamma(0.5 + beta, 0.5 * numin) * pow(2, 0.5 + beta) * pow(2.0 * M_PI, -0.5);
}
static double fcoll_approx_condition(double numin, double nucondition, double beta) {
    return (fcoll_approx(numin, beta) - fcoll_approx(nucondition, beta)) + fcoll_approx(nucondition, 0.) * pow(nucondition, beta);
}
static double nion_conditional_gamma_approx(double lnM_lo, double lnM_hi, const MFParams &p) {
    /* MFIntegral_Approx for the conditional N_ion integral (gamma_type = -3) */
    const double M_PIVOT1 = 1.5e9, M_PIVOT2 = 5.3e5, A1 = 9.0, A2 = 13.6, A3 = 21.0; /* hmf.c:97-101 */
    const double delta = p.delta, sigma_c = p.sigma_cond;
    double lo = lnM_lo;
    const double lnMturn = log(p.Mturn);
    if (lnMturn > lo) lo = lnMturn; /* sharp lower cut at the turnover */
    if (lo >= lnM_hi || EvaluateSigma(lo) <= sigma_c) return 0.;
    const double index_base = p.alpha_star + p.alpha_esc;
    const double delta_arg = pow((pc::delta_c_sph - delta) / p.growthf, 2);
    const double beta1 = index_base * A1 * 0.5, beta2 = index_base * A2 * 0.5, beta3 = index_base * A3 * 0.5;
    const double s1 = EvaluateSigma(log(M_PIVOT1)), s2 = EvaluateSigma(log(M_PIVOT2)), slo = EvaluateSigma(lo);
    const double nu_pivot1_umf = delta_arg / (s1 * s1), nu_pivot2_umf = delta_arg / (s2 * s2);
    const double nu_condition = delta_arg / (sigma_c * sigma_c);
    const double nu_pivot1 = delta_arg / (s1 * s1 - sigma_c * sigma_c), nu_pivot2 = delta_arg / (s2 * s2 - sigma_c * sigma_c);
    const double nu_lo = delta_arg / (slo * slo - sigma_c * sigma_c);
    if (nu_lo >= nu_condition) return fcoll_approx(nu_lo, 0.); /* flat part of sigma(nu): an erfc */
    double fcoll = 0.;
    if (nu_lo >= nu_pivot1) {
        fcoll += fcoll_approx_condition(nu_lo, nu_condition, beta1) * pow(nu_pivot1_umf, -beta1);
    } else {
        fcoll += fcoll_approx_condition(nu_pivot1, nu_condition, beta1) * pow(nu_pivot1_umf, -beta1);
        if (nu_lo > nu_pivot2) {
            fcoll += (fcoll_approx(nu_lo, beta2) - fcoll_approx(nu_pivot1, beta2)) * pow(nu_pivot1_umf, -beta2);
        } else {
            fcoll += (fcoll_approx(nu_pivot2, beta2) - fcoll_approx(nu_pivot1, beta2)) * pow(nu_pivot1_umf, -beta2);
            fcoll += (fcoll_approx(nu_lo, beta3) - fcoll_approx(nu_pivot2, beta3)) * pow(nu_pivot2_umf, -beta3);
        }
    }
    if (fcoll <= 0.0) fcoll = 1e-40;
    return fcoll;
}

double Nion_ConditionalM(double growthf, double lnM1, double lnM2, double lnM_cond, double sigma2,
                         double delta2, double Mturn, const ScalingConstants *sc, int method) {
    /* hmf.c:1106-1140 */
    MFParams p = mf_params(growthf, Mturn, sc);
    p.sigma_cond = sigma2;
    p.delta = delta2;
    if (lnM1 >= lnM_cond) return 0.;
    if (delta2 > (float)0.99 * get_delta_crit(p.HMF, sigma2, growthf)) {
        if (lnM_cond * (1 - pc::FRACT_FLOAT_ERR) <= lnM2) return nion_fraction(lnM_cond, p) / exp(lnM_cond);
        return 0.;
    }
    if (p.HMF != HMF_PS && p.HMF != HMF_ST && p.HMF != HMF_DELOS) p.HMF = HMF_PS;
    /* IntegratedNdM, hmf.c:896-905: Gauss-Legendre degrades near the barrier -> QAG above 1.2 */
    if (method == INTEG_QAG || (method == INTEG_GL && delta2 > 1.2)) return integrate_qag(lnM1, lnM2, p, 2);
    if (method == INTEG_GL) return integrate_gl(lnM1, lnM2, p);
    if (method == INTEG_GAMMA) return nion_conditional_gamma_approx(lnM1, lnM2, p);
    b200_throw(B200_ValueError, "invalid integration method %d", method);
}

/* ------------------------------------------------------------------ constant-zeta collapse fraction */
static float erfcc(float x) { /* Numerical Recipes erfc fit, hmf.c:1187-1203 (float in/out) */
    const double q = fabs(x), t = 1.0 / (1.0 + 0.5 * q);
    const double ans =
        t * exp(-q * q - 1.2655122 +
                t * (1.0000237 +
                     t * (0.374092 +
                          t * (0.0967842 +
                               t * (-0.1862881 +
                                    t * (0.2788681 +
                                         t * (-1.13520398 + t * (1.4885159 + t * (-0.82215223 + t * 0.17087277)))))))));
    return x >= 0.0 ? ans : 2.0 - ans;
}
double FgtrM_bias_fast(float growthf, float del_bias, float sig_small, float sig_large) { /* hmf.c:1221-1241 */
    if (sig_large > sig_small) b200_throw(B200_ValueError, "FgtrM in a region where M_min > M_max");
    if (sig_large == sig_small) return 0.;
    const double sig = sqrt(sig_small * sig_small - sig_large * sig_large);
    const double del = (pc::delta_c_sph - del_bias) / growthf;
    const double x = del / (sqrt(2) * sig);
    if (x < 0) return 1.0;
    return erfcc(x);
}

void build_fgtrm_table(FcollTable *t, double min_dens, double max_dens, double growthf,
                       double sigma_min, double sigma_max) { /* interp_tables.c:226-250 */
    t->x_min = min_dens;
    t->x_width = (max_dens - min_dens) / (N_DENS_INTERP - 1.);
    t->log_valued = 0;
    for (int i = 0; i < N_DENS_INTERP; i++) {
        const double dens = t->x_min + i * t->x_width;
        t->y[i] = FgtrM_bias_fast(growthf, dens, sigma_min, sigma_max);
    }
}

/* Everything in the Gauss-Legendre integrand that does not depend on the cell density: the same
   factors, evaluated once per node instead of once per (density, node).  The per-density
   expression below multiplies them in the order of cmf() / integrate_gl(), so the table values
   are the same doubles. */
struct GLNode {
    double w, nf, dsig, taylor, barrier, sdi, sdi15, sqrt_sdi;
    bool below; /* sigma(M) < sigma_cond: the conditional mass function vanishes */
};
static void gl_prepare_nodes(GLNode *nd, const MFParams &p) {
    for (int i = 1; i < NGL_INT + 1; i++) {
        GLNode &n = nd[i];
        const double lnM = xi_GL[i];
        const double sigma1 = EvaluateSigma(lnM);
        n.w = wi_GL[i];
        n.nf = nion_fraction(lnM, p);
        n.dsig = EvaluatedSigmasqdm(lnM);
        n.below = sigma1 < p.sigma_cond;
        n.sdi = sigma1 == p.sigma_cond ? 1e6 : 1 / (sigma1 * sigma1 - p.sigma_cond * p.sigma_cond);
        n.sdi15 = pow(n.sdi, 1.5);
        n.sqrt_sdi = sqrt(n.sdi);
        n.taylor = n.barrier = 0.;
        if (p.HMF == HMF_ST && !n.below) n.taylor = st_taylor_factor(sigma1, p.sigma_cond, p.growthf, &n.barrier);
    }
}
static double gl_integral(const GLNode *nd, const MFParams &p, double delta) {
    double integral = 0;
    const double delta_0 = delta / p.growthf;
    for (int i = 1; i < NGL_INT + 1; i++) {
        const GLNode &n = nd[i];
        double c;
        if (n.below) {
            c = 0.;
        } else if (p.HMF == HMF_ST) {
            const double factor = n.taylor - delta_0;
            c = -n.dsig * factor * n.sdi15 * exp(-(n.barrier - delta_0) * (n.barrier - delta_0) * 0.5 * (n.sdi)) /
                sqrt(2. * M_PI);
        } else if (p.HMF == HMF_DELOS) {
            const double nu = (pc::delta_c_delos - delta) * n.sqrt_sdi / p.growthf;
            const double dfdnu = 0.519 * pow(nu, 0.582) * exp(-0.469 * nu * nu);
            c = dfdnu * fabs(n.dsig * 0.5) * n.sdi;
        } else {
            const double del = (pc::delta_c_sph - delta) / p.growthf;
            c = -del * n.dsig * n.sdi15 * exp(-del * del * 0.5 * n.sdi) / sqrt(2. * M_PI);
        }
        integral += n.w * (n.nf * c);
    }
    return integral;
}

/* The table entries above delta = 1.2 are adaptive QAG integrals (hmf.c:896-905) of the same
   integrand at different delta.  Everything the integrand needs at an abscissa except one
   exponential is independent of delta (sigma, d sigma^2/dM, the moving-barrier series, the
   escape/stellar fractions), and QAG's abscissae depend only on the panel: the per-panel node data
   is computed once per table and shared by all entries and threads.  Same arithmetic per node as
   cmf() * nion_fraction(), so the integrals are bit-identical to the uncached evaluation. */
struct PanelNodes { GLNode n[61]; };
struct CondNodeCache {
    std::map<std::pair<double, double>, PanelNodes> panels;
    std::mutex mu;
};
static void cond_node(GLNode &n, double lnM, const MFParams &p) {
    const double sigma1 = EvaluateSigma(lnM);
    n.w = 1.;
    n.nf = nion_fraction(lnM, p);
    n.dsig = EvaluatedSigmasqdm(lnM);
    n.below = sigma1 < p.sigma_cond;
    n.sdi = sigma1 == p.sigma_cond ? 1e6 : 1 / (sigma1 * sigma1 - p.sigma_cond * p.sigma_cond);
    n.sdi15 = pow(n.sdi, 1.5);
    n.sqrt_sdi = sqrt(n.sdi);
    n.taylor = n.barrier = 0.;
    if (p.HMF == HMF_ST && !n.below) n.taylor = st_taylor_factor(sigma1, p.sigma_cond, p.growthf, &n.barrier);
}
static double cond_node_value(const GLNode &n, const MFParams &p, double delta) { /* nion_fraction * cmf */
    double c;
    if (n.below) {
        c = 0.;
    } else if (p.HMF == HMF_ST) {
        const double delta_0 = delta / p.growthf;
        const double factor = n.taylor - delta_0;
        c = -n.dsig * factor * n.sdi15 * exp(-(n.barrier - delta_0) * (n.barrier - delta_0) * 0.5 * (n.sdi)) /
            sqrt(2. * M_PI);
    } else if (p.HMF == HMF_DELOS) {
        const double nu = (pc::delta_c_delos - delta) * n.sqrt_sdi / p.growthf;
        const double dfdnu = 0.519 * pow(nu, 0.582) * exp(-0.469 * nu * nu);
        c = dfdnu * fabs(n.dsig * 0.5) * n.sdi;
    } else {
        const double del = (pc::delta_c_sph - delta) / p.growthf;
        c = -del * n.dsig * n.sdi15 * exp(-del * del * 0.5 * n.sdi) / sqrt(2. * M_PI);
    }
    return n.nf * c;
}
static double integrate_qag_cond_cached(double lo, double hi, const MFParams &p, CondNodeCache &cache) {
    double res, err;
    auto panel = [&](double a, double b, double *e, double *ra, double *rs) {
        const PanelNodes *pn = nullptr;
        {
            std::lock_guard<std::mutex> lk(cache.mu);
            auto it = cache.panels.find({a, b});
            if (it != cache.panels.end()) pn = &it->second; /* std::map nodes are stable */
        }
        if (!pn) {
            /* computed outside the lock (another thread may do the same work; the first insert wins) */
            PanelNodes fresh;
            double x[61];
            hostnum::gk61_abscissae(a, b, x);
            for (int i = 0; i < 61; i++) cond_node(fresh.n[i], x[i], p);
            std::lock_guard<std::mutex> lk(cache.mu);
            pn = &cache.panels.emplace(std::make_pair(a, b), fresh).first->second;
        }
        int k = 0;
        return hostnum::gk61_panel([&](double) { return cond_node_value(pn->n[k++], p, p.delta); }, a, b, e, ra, rs);
    };
    const int st_ = hostnum::qag61_panels(panel, lo, hi, 0, 1e-3, 1000, &res, &err);
    if (st_ != 0) b200_throw(B200_GSLError, "mass-function quadrature status %d", st_);
    return res;
}

void build_nion_table(FcollTable *t, double redshift, double min_dens, double max_dens, double Mmin,
                      double Mmax, const ScalingConstants *sc, int method, int n_threads, int part, int nparts) {
    /* initialise_Nion_Conditional_spline without mini-halos, interp_tables.c:291-408 */
    const double growthf = dicke(redshift);
    const double lnMmin = log(Mmin), lnMmax = log(Mmax), lnMcond = log(Mmax);
    const double sigma2 = EvaluateSigma(log(Mmax));
    t->x_min = min_dens;
    t->x_width = (max_dens - min_dens) / (N_DENS_INTERP - 1.);
    t->log_valued = 1;
    int err_code = 0;
    if (n_threads < 1) n_threads = 1;

    MFParams p = mf_params(growthf, sc->mturn_a_nofb, sc);
    p.sigma_cond = sigma2;
    if (p.HMF != HMF_PS && p.HMF != HMF_ST && p.HMF != HMF_DELOS) p.HMF = HMF_PS;
    GLNode nodes[NGL_INT + 1];
    const bool gl_fast = method == INTEG_GL && lnMmin < lnMcond;
    if (gl_fast) {
        if ((float)lnMmin != (float)GL_limit[0] || (float)lnMmax != (float)GL_limit[1])
            b200_throw(B200_TableGenerationError, "integral limits do not match the Gauss-Legendre nodes");
        gl_prepare_nodes(nodes, p);
    }
    const double dcrit_lim = (float)0.99 * get_delta_crit(matter_options_global->HMF, sigma2, growthf);
    CondNodeCache qag_cache;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 1)
    for (int i = part; i < N_DENS_INTERP; i += nparts) {
        try {
            const double dens = min_dens + (float)i / ((float)N_DENS_INTERP - 1.) * (max_dens - min_dens);
            double v;
            if (gl_fast && !(dens > dcrit_lim) && !(dens > 1.2)) {
                v = gl_integral(nodes, p, dens);
            } else if (lnMmin < lnMcond && !(dens > dcrit_lim) && (method == INTEG_QAG || (method == INTEG_GL && dens > 1.2))) {
                /* Nion_ConditionalM's QAG branch (hmf.c:1106-1140, 896-905) with shared node data */
                MFParams pd = p;
                pd.delta = dens;
                v = integrate_qag_cond_cached(lnMmin, lnMmax, pd, qag_cache);
            } else {
                v = Nion_ConditionalM(growthf, lnMmin, lnMmax, lnMcond, sigma2, dens, sc->mturn_a_nofb, sc, method);
            }
            float y = log(v);
            if (y < -40.) y = -40.;
            t->y[i] = y;
            if (!std::isfinite(y)) err_code = B200_TableGenerationError;
        } catch (B200Error &e) { err_code = e.code; }
    }
    if (err_code) b200_throw(err_code, "conditional Nion table generation failed");
}

void build_cond_table(FcollTable *t, double redshift, double min_dens, double max_dens, double Mmin,
                      double Mmax, double Mcond, const ScalingConstants *sc, int method, double log_floor,
                      int n_threads) {
    const double growthf = dicke(redshift);
    const double lnMmin = log(Mmin), lnMmax = log(Mmax), lnMcond = log(Mcond);
    const double sigma2 = EvaluateSigma(lnMcond);
    t->x_min = min_dens;
    t->x_width = (max_dens - min_dens) / (N_DENS_INTERP - 1.);
    t->log_valued = 1;
    int err_code = 0;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 1)
    for (int i = 0; i < N_DENS_INTERP; i++) {
        try {
            const double dens = min_dens + (float)i / ((float)N_DENS_INTERP - 1.) * (max_dens - min_dens);
            float y = log(Nion_ConditionalM(growthf, lnMmin, lnMmax, lnMcond, sigma2, dens, sc->mturn_a_nofb, sc, method));
            if (y < log_floor) y = log_floor;
            t->y[i] = y;
            if (!std::isfinite(y)) err_code = B200_TableGenerationError;
        } catch (B200Error &e) { err_code = e.code; }
    }
    if (err_code) b200_throw(err_code, "conditional table generation failed");
}

ScalingConstants evolve_scaling_constants_sfr(const ScalingConstants *sc) {
    ScalingConstants s = *sc;
    s.fesc_10 = 1.;
    s.alpha_esc = 0.;
    s.Mlim_Fesc = 0.;
    return s;
}

/* ------------------------------------------------------------------ IonizeBox host constants */
void set_ionbox_constants(double redshift, double prev_redshift, IonConsts *c) {
    /* IonisationBox.c:125-227 (photon conservation is out of scope) */
    c->redshift = redshift;
    c->prev_redshift = prev_redshift;
    c->stored_redshift = redshift;
    set_scaling_constants(redshift, &c->sc);
    c->growth_factor = dicke(redshift);
    c->mass_dep_zeta = matter_options_global->SOURCE_MODEL != SRC_CONST_ION_EFF;
    c->hii_filter = astro_options_global->HII_FILTER;
    c->T_re = astro_params_global->T_RE;
    if (c->mass_dep_zeta)
        c->ion_eff_factor_gl = c->sc.pop2_ion * c->sc.fstar_10 * c->sc.fesc_10;
    else
        c->ion_eff_factor_gl = astro_params_global->HII_EFF_FACTOR;
    c->ion_eff_factor = c->ion_eff_factor_gl;
    /* the halo fields already carry f_star, f_esc and the photon yield (IonisationBox.c:170-178) */
    c->lagrangian = matter_options_global->SOURCE_MODEL >= SRC_L_INTEGRAL;
    if (c->lagrangian) c->ion_eff_factor = 1.;
    c->mfp_meandens = 25.483241248322766 / cosmo_params_global->hlittle;
    c->M_min = minimum_source_mass(redshift, false);
    c->lnMmin = log(c->M_min);
    c->lnMmax_gl = log(pc::M_MAX_INTEGRAL);
    c->sigma_minmass = sigma_z0(c->M_min);
    c->TK_nofluct = T_RECFAST(redshift);
    c->adia_TK_term = cT_approx(redshift);
    c->pixel_length = simulation_options_global->BOX_LEN / (double)simulation_options_global->HII_DIM;
    /* the first snapshot takes its step from ZPRIME_STEP_FACTOR */
    if (prev_redshift < 1) c->dz = (1. + redshift) * (simulation_options_global->ZPRIME_STEP_FACTOR - 1.);
    else c->dz = prev_redshift - redshift;
    c->fabs_dtdz = fabs(dtdz(redshift)) / 1e15; /* the rate table is in (1e15 s)^-1 */
    c->gamma_prefactor = pow(1 + redshift, 2) * pc::cm_per_Mpc * pc::sigma_HI * astro_params_global->ALPHA_UVB /
                         (astro_params_global->ALPHA_UVB + 2.75) * n_b0() * c->ion_eff_factor / 1.0e-12;
    if (c->lagrangian) c->gamma_prefactor /= rho_crit() * cosmo_params_global->OMb;
    else c->gamma_prefactor = c->gamma_prefactor / (c->sc.t_h * c->sc.t_star);
}

std::vector<RadiusSpec> setup_radii(const IonConsts &c) { /* IonisationBox.c:964-1006 */
    const AstroParams *ap = astro_params_global;
    const double maximum_radius = fmin(ap->R_BUBBLE_MAX, pc::l_factor * simulation_options_global->BOX_LEN);
    double cell_length_factor = pc::l_factor;
    if (c.lagrangian && !astro_options_global->IONISE_ENTIRE_SPHERE && c.pixel_length < 1) cell_length_factor = 1.;
    const double minimum_radius = fmax(ap->R_BUBBLE_MIN, cell_length_factor * c.pixel_length);
    int n_radii = (int)(log(maximum_radius / minimum_radius) / log(ap->DELTA_R_HII_FACTOR) + 1);
    std::vector<RadiusSpec> r;
    for (int i = 0; i < n_radii; i++) {
        RadiusSpec s;
        s.R_index = i;
        s.R = minimum_radius * pow(ap->DELTA_R_HII_FACTOR, i);
        if (s.R > maximum_radius - pc::FRACT_FLOAT_ERR) {
            s.R = maximum_radius;
            n_radii = i + 1;
        }
        s.M_max_R = RtoM(s.R);
        s.ln_M_max_R = log(s.M_max_R);
        s.sigma_maxmass = sigma_z0(s.M_max_R);
        r.push_back(s);
    }
    return r;
}


/* ------------------------------------------------------------------ RECFAST boundary values
 * heating_helper_progs.c:94-197: z, x_e, -, T_k columns, 501 rows from z=500 down to 0. */
static hostnum::CubicSpline g_T_spline, g_x_spline;
extern "C" int init_heat(void) {
    try {
        if (!config_settings.external_table_path) b200_throw(B200_IOError, "external_table_path is not set");
        std::string fn = std::string(config_settings.external_table_path) + "/recfast_LCDM.dat";
        FILE *F = fopen(fn.c_str(), "r");
        if (!F) b200_throw(B200_IOError, "unable to open %s", fn.c_str());
        const int npts = 501;
        std::vector<double> z(npts), T(npts), xe(npts);
        for (int i = npts - 1; i >= 0; i--) {
            float cz, cx, tr, ct;
            if (fscanf(F, "%f %E %E %E", &cz, &cx, &tr, &ct) != 4) {
                /* the reference ignores short reads (last rows keep stale values); a 500-row
                   file therefore duplicates the last parsed row */
                if (i + 1 < npts) { cz = (float)z[i + 1]; cx = (float)xe[i + 1]; ct = (float)T[i + 1]; }
            }
            z[i] = cz; xe[i] = cx; T[i] = ct;
        }
        fclose(F);
        /* keep the knots strictly increasing for the spline */
        for (int i = 1; i < npts; i++)
            if (!(z[i] > z[i - 1])) { z.erase(z.begin() + i - 1); T.erase(T.begin() + i - 1); xe.erase(xe.begin() + i - 1); i = 0; }
        g_T_spline.init(z, T);
        g_x_spline.init(z, xe);
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] init_heat failed: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
extern "C" void destruct_heat(void) { g_T_spline.clear(); g_x_spline.clear(); }
bool heat_ready() { return g_T_spline.ready(); }
double T_RECFAST(float z) {
    if (!heat_ready()) b200_throw(B200_ValueError, "init_heat() was never called");
    if (z > g_T_spline.xmax()) b200_throw(B200_ValueError, "T_RECFAST called with z=%f", z);
    return g_T_spline.eval(z);
}
double xion_RECFAST(float z) {
    if (!heat_ready()) b200_throw(B200_ValueError, "init_heat() was never called");
    if (z > g_x_spline.xmax()) b200_throw(B200_ValueError, "xion_RECFAST called with z=%f", z);
    return g_x_spline.eval(z);
}
float cT_approx(float z) { return 0.58 - 0.006 * (z - 10.0); }

extern "C" int CreateFFTWWisdoms(void) { return 0; }

/* integral_wrappers.c:18-24: sigma(M) and d sigma^2/dM from the interpolation table for an array of masses */
extern "C" void get_sigma(int n_masses, double *mass_values, double *sigma_out, double *dsigmasqdm_out) {
    for (int i = 0; i < n_masses; i++) {
        try {
            sigma_out[i] = EvaluateSigma(log(mass_values[i]));
            dsigmasqdm_out[i] = EvaluatedSigmasqdm(log(mass_values[i]));
        } catch (B200Error &e) { /* table not initialised / out of range: NaN, never an exception across the C boundary */
            fprintf(stderr, "[21cmfast_b200] get_sigma: %s\n", e.msg);
            sigma_out[i] = dsigmasqdm_out[i] = std::nan("");
        }
    }
}
/* read by the reference's coeval driver (coeval.py:686); photon conservation is not built */
extern "C" { bool photon_cons_allocated = false; }
