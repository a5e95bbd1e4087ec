/*
 * host_physics.h -- host-side scalar cosmology / mass-function helpers of the hot path.
 *
 * These are the "(host)" rows of SURVEY.md section 2a: they produce a few hundred doubles per
 * call (growth factors, the radius ladder, sigma(M), mean collapse fractions, the 400-point
 * per-radius tables) that parameterise the grid kernels.  They follow the same formulae and the
 * same float/double types as the reference so that GPU-vs-reference differences isolate the
 * grid kernels (citations per function in host_physics.cpp).
 */
#pragma once
#include "rt.h"

/* physical constants (Constants.c:4-47) */
namespace pc {
constexpr double c_kms = 2.99792458e5, G = 6.6743e-8, m_p = 1.67262192369e-24;
constexpr double Msun = 1.989e33, cm_per_Mpc = 3.08567758e24, sigma_HI = 6.3e-18;
constexpr double T_cmb = 2.7255, l_factor = 0.620350491, delta_c_sph = 1.686, delta_c_delos = 1.5;
constexpr double FRACT_FLOAT_ERR = 1e-7, TINY = 1e-30;
constexpr double M_MIN_INTEGRAL = 1e5, M_MAX_INTEGRAL = 1e16;
}  // namespace pc

/* enum values of InputParameters.h:9-57 */
enum { HMF_PS = 0, HMF_ST = 1, HMF_WATSON = 2, HMF_WATSON_Z = 3, HMF_DELOS = 4, HMF_REED07 = 5, HMF_YUNG24 = 6 };
enum { FILTER_TOPHAT = 0, FILTER_SHARP_K = 1, FILTER_GAUSSIAN = 2 };
enum { PERTURB_LINEAR = 0, PERTURB_ZELDOVICH = 1, PERTURB_2LPT = 2 };
enum { SRC_CONST_ION_EFF = 0, SRC_E_INTEGRAL = 1, SRC_L_INTEGRAL = 2 };
enum { INTEG_QAG = 0, INTEG_GL = 1, INTEG_GAMMA = 2 };

/* parameter access with a loud failure if Broadcast_struct_global_* was never called */
void require_params(bool need_astro);
int hii_d_para();
int d_para();
double box_volume();

/* cosmology.c */
double hubble_H0();   /* s^-1 */
double rho_crit();    /* Msun Mpc^-3 */
double n_b0();        /* baryon number density today, cm^-3 */
double MtoR(double M);
double RtoM(double R);
double omega_mz(float z);
double TtoM(double z, double T, double mu);
double dtdz(float z);
double hubble(float z);
double t_hubble(float z);
double ddickedt(double z);

/* scaling_relations.c:36-115 */
struct ScalingConstants {
    double redshift;
    double fstar_10, alpha_star, fesc_10, alpha_esc, pop2_ion, t_h, t_star;
    double mturn_a_nofb, Mlim_Fstar, Mlim_Fesc;
};
void set_scaling_constants(double redshift, ScalingConstants *sc);

/* hmf.c */
double EvaluateSigma(double lnM);
double EvaluatedSigmasqdm(double lnM);
double Nion_General(double z, double lnMmin, double lnMmax, double Mturn, const ScalingConstants *sc);
double Fcoll_General(double z, double lnMmin, double lnMmax);
double Nion_ConditionalM(double growthf, double lnM1, double lnM2, double lnM_cond, double sigma2,
                         double delta2, double Mturn, const ScalingConstants *sc, int method);
double FgtrM_bias_fast(float growthf, float del_bias, float sig_small, float sig_large);
extern "C" void initialise_GL(double lnM_Min, double lnM_Max); /* also part of the reference's cffi surface */

/* heating_helper_progs.c:94-197 */
double T_RECFAST(float z);
double xion_RECFAST(float z);
float cT_approx(float z);
bool heat_ready();

/* IonisationBox.c:125-227 / :964-1006 host constants (kept in the g++-compiled unit: under nvcc's
   host pass, math calls on float arguments silently bind to float overloads) */
#include <vector>
struct IonConsts {
    double redshift, stored_redshift, prev_redshift, growth_factor;
    bool mass_dep_zeta;
    int hii_filter;
    ScalingConstants sc;
    double T_re, ion_eff_factor, ion_eff_factor_gl;
    double TK_nofluct, adia_TK_term;
    double M_min, lnMmin, lnMmax_gl, sigma_minmass, pixel_length;
    double dz, fabs_dtdz, gamma_prefactor; /* recombination bookkeeping (IonisationBox.c:132-137,144,211-218) */
    bool lagrangian;      /* source grids come from a HaloBox (IonisationBox.c:151-153,172-178) */
    double mfp_meandens;  /* mean free path of the exponential filter, Mpc (IonisationBox.c:191) */
};
struct RadiusSpec {
    double R, M_max_R, ln_M_max_R, sigma_maxmass;
    int R_index;
};

void set_ionbox_constants(double redshift, double prev_redshift, IonConsts *c);
std::vector<RadiusSpec> setup_radii(const IonConsts &c);

/* per-radius 400-point table of f_coll(delta) (interp_tables.c:226-250 / :291-408) */
#define N_DENS_INTERP 400
struct FcollTable {
    double x_min, x_width;
    float y[N_DENS_INTERP];
    int log_valued; /* 1: y = ln(Nion), evaluate exp(interp) (E-INTEGRAL); 0: linear (CONST) */
};
void build_fgtrm_table(FcollTable *t, double min_dens, double max_dens, double growthf,
                       double sigma_min, double sigma_max);
/* part / nparts: only the entries i = part (mod nparts) are computed (one table built by several ranks) */
void build_nion_table(FcollTable *t, double redshift, double min_dens, double max_dens,
                      double Mmin, double Mmax, const ScalingConstants *sc, int method,
                      int n_threads, int part = 0, int nparts = 1);
/* the same table for a condition mass that is not the upper limit (the Lagrangian cell of the halo
   boxes): initialise_Nion_Conditional_spline / initialise_SFRD_Conditional_table without mini-halos
   (interp_tables.c:291-408, :415-495); log_floor = -40 resp. -50 */
void build_cond_table(FcollTable *t, double redshift, double min_dens, double max_dens, double Mmin,
                      double Mmax, double Mcond, const ScalingConstants *sc, int method, double log_floor,
                      int n_threads);
/* scaling_relations.c:122-131: the star-formation integrals are the N_ion integrals without f_esc */
ScalingConstants evolve_scaling_constants_sfr(const ScalingConstants *sc);
