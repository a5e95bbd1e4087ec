/*
 * py21cmfast_b200.h -- C ABI of lib21cmfast_b200.so (B200 / sm_100a implementation of the
 * 21cmFAST 3-D grid hot path).
 *
 * Every entry point below replaces the function of the same name in the reference's cffi
 * extension `py21cmfast.c_21cmfast`; names, argument order, struct layouts and the integer
 * status convention are the reference's, so a cffi/ctypes binding written for the reference
 * binds this library unchanged (INTEGRATION.md shows the binding).  Citations are
 * reference file:line under /root/reference/src/py21cmfast/src/.
 *
 * Conventions:
 *   - all array pointers are HOST memory owned by the caller (numpy); the library copies
 *     host->device, runs hand-written CUDA kernels and copies results back before returning;
 *   - real boxes are unpadded C-order [x][y][z] float32 (indexing.h:84-86);
 *   - parameters are read at call time through the global pointers set by
 *     Broadcast_struct_global_all (InputParameters.c:11-54); the library never owns them;
 *   - return 0 on success, else a code of exceptions.h:12-21 (10 = CUDA runtime failure);
 *   - not re-entrant (global parameter pointers, static tables), like the reference.
 */
#ifndef PY21CMFAST_B200_H
#define PY21CMFAST_B200_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- input structs: _inputparams_wrapper.h:11-182 (field order and types are ABI) ---------- */
typedef struct CosmoParams {
    float hlittle, OMm, OMl, OMb, POWER_INDEX;
    float OMn, OMk, OMr, OMtot, Y_He, wl;
} CosmoParams;

typedef struct SimulationOptions {
    int HII_DIM;
    int DIM;
    float BOX_LEN;
    float NON_CUBIC_FACTOR;
    int N_THREADS;
    double Z_HEAT_MAX;
    double ZPRIME_STEP_FACTOR;
    float SAMPLER_MIN_MASS;
    double SAMPLER_BUFFER_FACTOR;
    int N_COND_INTERP;
    int N_PROB_INTERP;
    double MIN_LOGPROB;
    double HALOMASS_CORRECTION;
    double PARKINSON_G0;
    double PARKINSON_y1;
    double PARKINSON_y2;
    float INITIAL_REDSHIFT;
    double DELTA_R_FACTOR;
    double DENSITY_SMOOTH_RADIUS;
    double DEXM_OPTIMIZE_MINMASS;
    double DEXM_R_OVERLAP;
    double CORR_STAR;
    double CORR_SFR;
    double CORR_LX;
    double MIN_XE_FOR_FCOLL_IN_TAUX;
} SimulationOptions;

typedef struct MatterOptions {
    bool USE_FFTW_WISDOM;
    int HMF;
    int V_CB_MODEL;
    int POWER_SPECTRUM;
    int USE_INTERPOLATION_TABLES;
    bool PERTURB_ON_HIGH_RES;
    int PERTURB_ALGORITHM;
    bool MINIMIZE_MEMORY;
    bool KEEP_3D_VELOCITIES;
    bool DEXM_OPTIMIZE;
    int FILTER;
    int HALO_FILTER;
    bool SMOOTH_EVOLVED_DENSITY_FIELD;
    int SOURCE_MODEL;
    int SAMPLE_METHOD;
} MatterOptions;

typedef struct AstroParams {
    float HII_EFF_FACTOR;
    float F_STAR10;
    float ALPHA_STAR;
    float ALPHA_STAR_MINI;
    float SIGMA_STAR;
    double UPPER_STELLAR_TURNOVER_MASS;
    double UPPER_STELLAR_TURNOVER_INDEX;
    float F_STAR7_MINI;
    float t_STAR;
    double SIGMA_SFR_INDEX;
    double SIGMA_SFR_LIM;
    double L_X;
    double L_X_MINI;
    double SIGMA_LX;
    float F_ESC10;
    float ALPHA_ESC;
    float F_ESC7_MINI;
    float T_RE;
    float M_TURN;
    float R_BUBBLE_MAX;
    float ION_Tvir_MIN;
    double F_H2_SHIELD;
    float NU_X_THRESH;
    float X_RAY_SPEC_INDEX;
    float X_RAY_Tvir_MIN;
    double A_LW;
    double BETA_LW;
    double A_VCB;
    double BETA_VCB;
    double V_CB_AVG_DEBUG;
    double POP2_ION;
    double POP3_ION;
    double PHOTONCONS_CALIBRATION_END;
    double CLUMPING_FACTOR;
    double ALPHA_UVB;
    float R_MAX_TS;
    int N_STEP_TS;
    double DELTA_R_HII_FACTOR;
    float R_BUBBLE_MIN;
    double MAX_DVDR;
    double NU_X_MAX;
    double NU_X_BAND_MAX;
} AstroParams;

typedef struct AstroOptions {
    bool USE_MINI_HALOS;
    bool USE_X_RAY_HEATING;
    bool USE_CMB_HEATING;
    bool USE_LYA_HEATING;
    int RECOMB_MODEL;
    bool USE_TS_FLUCT;
    bool M_MIN_in_Mass;
    bool USE_EXP_FILTER;
    bool CELL_RECOMB;
    bool LYA_MULTIPLE_SCATTERING;
    bool USE_ADIABATIC_FLUCTUATIONS;
    int PHOTON_CONS_TYPE;
    bool USE_UPPER_STELLAR_TURNOVER;
    bool HALO_SCALING_RELATIONS_MEDIAN;
    int HII_FILTER;
    int HEAT_FILTER;
    bool IONISE_ENTIRE_SPHERE;
    int INTEGRATION_METHOD_ATOMIC;
    int INTEGRATION_METHOD_MINI;
} AstroOptions;

typedef struct Table1D {
    int size;
    double *x_values;
    double *y_values;
} Table1D;

typedef struct CosmoTables {
    Table1D *transfer_density;
    Table1D *transfer_vcb;
    double ps_norm;
    bool USE_SIGMA_8;
    double V_CB_AVG;
} CosmoTables;

typedef struct ConfigSettings {
    double HALO_CATALOG_MEM_FACTOR;
    bool EXTRA_HALOBOX_FIELDS;
    char *external_table_path;
    char *wisdoms_path;
} ConfigSettings;

/* ---- output structs: _outputstructs_wrapper.h:6-105 -------------------------------------- */
typedef struct InitialConditions {
    float *lowres_density, *lowres_vx, *lowres_vy, *lowres_vz;
    float *lowres_vx_2LPT, *lowres_vy_2LPT, *lowres_vz_2LPT;
    float *hires_density, *hires_vx, *hires_vy, *hires_vz;
    float *hires_vx_2LPT, *hires_vy_2LPT, *hires_vz_2LPT;
    float *lowres_vcb;
} InitialConditions;

typedef struct PerturbedField {
    float *density, *velocity_x, *velocity_y, *velocity_z;
} PerturbedField;

typedef struct HaloBox {
    float *halo_mass, *halo_stars, *halo_stars_mini, *count;
    float *n_ion, *halo_sfr, *halo_xray, *halo_sfr_mini, *whalo_sfr;
    double log10_Mcrit_ACG_ave;
    double log10_Mcrit_MCG_ave;
} HaloBox;

typedef struct TsBox {
    float *spin_temperature, *xray_ionised_fraction, *kinetic_temp_neutral, *J_21_LW;
    double Q_HI;
} TsBox;

typedef struct IonizedBox {
    double mean_f_coll;
    double mean_f_coll_MINI;
    double log10_Mturnover_ave;
    double log10_Mturnover_MINI_ave;
    float *neutral_fraction;
    float *ionisation_rate_G12;
    float *mean_free_path;
    float *z_reion;
    float *cumulative_recombinations;
    float *kinetic_temperature;
    float *unnormalised_nion;
    float *unnormalised_nion_mini;
} IonizedBox;

typedef struct BrightnessTemp {
    float *brightness_temp;
    float *tau_21;
} BrightnessTemp;

/* ---- global parameter pointers: _inputparams_wrapper.h:195-202, InputParameters.c:80-89 --- */
extern SimulationOptions *simulation_options_global;
extern MatterOptions *matter_options_global;
extern CosmoParams *cosmo_params_global;
extern AstroParams *astro_params_global;
extern AstroOptions *astro_options_global;
extern CosmoTables *cosmo_tables_global;
extern ConfigSettings config_settings;

/* ---- hot-path compute functions: _functionprototypes_wrapper.h:6-9,23-26,28-29 ----------- */
int ComputeInitialConditions(unsigned long long int random_seed, InitialConditions *boxes);
int ComputePerturbedField(float redshift, InitialConditions *boxes, PerturbedField *perturbed_field);
int ComputeIonizedBox(float redshift, float prev_redshift, PerturbedField *perturbed_field,
                      PerturbedField *previous_perturbed_field, IonizedBox *previous_ionize_box,
                      TsBox *spin_temp, HaloBox *halos, InitialConditions *ini_boxes,
                      IonizedBox *box);
int ComputeBrightnessTemp(float redshift, TsBox *spin_temp, IonizedBox *ionized_box,
                          PerturbedField *perturb_field, BrightnessTemp *box);
/* filter known-answer hook used by the reference's tests (_functionprototypes_wrapper.h:130-131) */
int test_filter(float *input_box, double R, double R_param, double R_star, int filter_flag,
                double *result);

/* ---- initialisation / teardown: _functionprototypes_wrapper.h:67-87 ----------------------- */
void Broadcast_struct_global_all(SimulationOptions *simulation_options,
                                 MatterOptions *matter_options, CosmoParams *cosmo_params,
                                 AstroParams *astro_params, AstroOptions *astro_options,
                                 CosmoTables *cosmo_tables);
void Broadcast_struct_global_noastro(SimulationOptions *simulation_options,
                                     MatterOptions *matter_options, CosmoParams *cosmo_params);
void Free_cosmo_tables_global(void);
void init_ps(void);
void free_ps(void);
void initialiseSigmaMInterpTable(float M_Min, float M_Max);
void freeSigmaMInterpTable(void);
int init_heat(void);
void destruct_heat(void);
void init_MHR(void);          /* recombination-rate tables (recombinations.c:92-122); needed when RECOMB_MODEL != none */
void free_MHR(void);
/* PDF-integrated recombination rate per baryon in (1e15 s)^-1 (recombinations.c:66-90); NaN before init_MHR() */
double splined_recombination_rate(double z_eff, double gamma12_bg);
int CreateFFTWWisdoms(void);  /* no-op: the FFT is the library's own sm_100a kernels */

/* ---- the rest of the reference's cffi surface (_functionprototypes_wrapper.h), OUTSIDE the scoped hot
   path.  Exported so that the library is a complete link target for the reference's API-mode cffi build
   (SURVEY.md section 8b); each one only reports "not built": status 3 (ValueError) / NaN / a message on
   stderr.  Generated by tools/gen_unscoped_stubs.py into csrc/unscoped.cpp. ------------------------------ */
typedef struct HaloCatalog HaloCatalog;                   /* opaque here: pointers are passed through only */
typedef struct PerturbedHaloCatalog PerturbedHaloCatalog;
typedef struct XraySourceBox XraySourceBox;
extern bool photon_cons_allocated;                        /* always false: photon conservation is not built */
int ComputeHaloCatalog(float redshift_desc, float redshift, InitialConditions *boxes, unsigned long long int random_seed, HaloCatalog *halos_desc, HaloCatalog *halos);
int ComputePerturbedHaloCatalog(float redshift, InitialConditions *boxes, TsBox *prev_ts, IonizedBox *prev_ion, HaloCatalog *halos, PerturbedHaloCatalog *halos_perturbed);
int ComputeTsBox(float redshift, float prev_redshift, float perturbed_field_redshift, short cleanup, PerturbedField *perturbed_field, XraySourceBox *source_box, TsBox *previous_spin_temp, InitialConditions *ini_boxes, TsBox *this_spin_temp);
int ComputeHaloBox(double redshift, InitialConditions *ini_boxes, HaloCatalog *halos, TsBox *previous_spin_temp, IonizedBox *previous_ionize_box, HaloBox *grids);
int UpdateXraySourceBox(HaloBox *halobox, double R_inner, double R_outer, int R_ct, double R_star, XraySourceBox *source_box);
int InitialisePhotonCons(void);
int PhotonCons_Calibration(double *z_estimate, double *xH_estimate, int NSpline);
int ComputeZstart_PhotonCons(double *zstart);
void adjust_redshifts_for_photoncons(double z_step_factor, float *redshift, float *stored_redshift, float *absolute_delta_z);
void determine_deltaz_for_photoncons(void);
int ObtainPhotonConsData(double *z_at_Q_data, double *Q_data, int *Ndata_analytic, double *z_cal_data, double *nf_cal_data, int *Ndata_calibration, double *PhotonCons_NFdata, double *PhotonCons_deltaz, int *Ndata_PhotonCons);
void FreePhotonConsMemory(void);
void set_alphacons_params(double norm, double slope);
int ComputeLF(int nbins, int component, int NUM_OF_REDSHIFT_FOR_LF, float *z_LF, float *M_TURNs, double *M_uv_z, double *M_h_z, double *log10phi);
float ComputeTau(int NPoints, float *redshifts, float *global_xHI, float z_re_HeII);
void get_condition_integrals(double redshift, double z_prev, int n_conditions, double *cond_values, double *out_n_exp, double *out_m_exp);
void get_halo_chmf_interval(double redshift, double z_prev, int n_conditions, double *cond_values, int n_masslim, double *lnM_lo, double *lnM_hi, double *out_n);
void get_halomass_at_probability(double redshift, double z_prev, int n_conditions, double *cond_values, double *probabilities, double *out_mass);
void get_global_SFRD_z(int n_redshift, double *redshifts, double *log10_turnovers_mcg, double *out_sfrd, double *out_sfrd_mini);
void get_global_Nion_z(int n_redshift, double *redshifts, double *log10_turnovers_mcg, double *out_nion, double *out_nion_mini);
void get_conditional_FgtrM(double redshift, double R, int n_densities, double *densities, double *out_fcoll, double *out_dfcoll);
void get_conditional_SFRD(double redshift, double R, int n_densities, double *densities, double log10_mturns_mini, double *out_sfrd, double *out_sfrd_mini);
void get_conditional_Nion(double redshift, double R, int n_densities, double *densities, double log10_mturn_acg, double log10_mturn_mcg, double *out_nion, double *out_nion_mini);
void get_conditional_Xray(double redshift, double R, int n_densities, double *densities, double log10_mturns_mini, double *out_xray);
int SomethingThatCatches(bool sub_func);
int FunctionThatCatches(bool sub_func, bool pass, double *result);
void FunctionThatThrows(void);
int single_test_sample(unsigned long long int seed, int n_condition, float *conditions, float *cond_crd, double z_out, double z_in, int *out_n_tot, int *out_n_cell, double *out_n_exp, double *out_m_cell, double *out_m_exp, float *out_halo_masses, float *out_halo_coords);
int test_halo_props(double redshift, float *vcb_grid, float *J21_LW_grid, float *z_re_grid, float *Gamma12_ion_grid, int n_halos, float *halo_masses, float *halo_coords, float *star_rng, float *sfr_rnd, float *xray_rng, float *halo_props_out);
double compute_mu_for_multiple_scattering(double x_em);
double compute_eta_for_multiple_scattering(double x_em);
double hyper_2F3(double kR, double alpha, double beta);
double power_in_vcb(double k);
double unconditional_hmf(double growthf, double lnM, double z, int HMF);
double conditional_hmf(double growthf, double lnM, double delta, double sigma, int HMF);
double expected_nhalo(double redshift);
void compute_mturns(float z, float J_21_LW, float vcb, float Gamma12, float z_reion, double *M_turn_a, double *M_turn_m);
/* real implementations (integral_wrappers.c:18-24, hmf.c:699-706) */
void get_sigma(int n_masses, double *mass_values, double *sigma_out, double *dsigmasqdm_out);
void initialise_GL(double lnM_Min, double lnM_Max);

/* ---- scalar cosmology helpers the Python layer and tests call directly (:136-150) -------- */
double dicke(double z);
double sigma_z0(double M);
double dsigmasqdm_z0(double M);
double power_in_k(double k);
double get_delta_crit(int HMF, double sigma, double growthf);
double atomic_cooling_threshold(float z);
double minimum_source_mass(double redshift, bool xray);

/* ---- B200-specific additions (not in the reference; optional for a drop-in) --------------- */
/* Device selection for one-process-per-GPU launches (default: device 0 / CUDA_VISIBLE_DEVICES). */
int b200_set_device(int device);
/* Drop every cached device buffer (initial conditions kept resident between calls, FFT plans). */
void b200_release_device_cache(void);
/* Device-resident initial conditions across ComputePerturbedField calls.  OFF by default: like the
   reference (PerturbedField.c:389-496 reads boxes->hires_density on every call) the library uploads
   the caller's arrays each time.  b200_ics_cache(1) (or B200_ICS_CACHE=1) keeps them in HBM between
   calls on the same InitialConditions; the caller then promises not to modify those arrays in place
   without calling b200_ics_cache_invalidate().  b200_ics_cache(0) switches it off and frees the copy. */
void b200_ics_cache(int enable);
void b200_ics_cache_invalidate(void);
/* Opt-in residency of the boxes handed from one of the reference's entry points to the next
   (drivers/coeval.py:835-853: PerturbedField -> IonizedBox -> BrightnessTemp): with enable != 0,
   ComputePerturbedField leaves its density and ComputeIonizedBox its neutral fraction on the device under the host
   pointer the caller received them in, and ComputeIonizedBox / ComputeBrightnessTemp use those copies instead of
   uploading the arrays again.  The caller must not modify or free those host arrays while it is on;
   enable = 0 drops every copy.  Off by default (the reference reads the caller's arrays on every call). */
void b200_residency(int enable);
/* Counters for the last Compute* call: kernels launched, H2D and D2H bytes, device milliseconds
   (CUDA events on the library's stream). Any pointer may be NULL. */
void b200_last_call_stats(long long *kernel_launches, long long *h2d_bytes, long long *d2h_bytes,
                          double *device_ms);
/* Device-resident variants used by bench.py's `value` leg (inputs already in HBM): identical
   computation, but the structs hold DEVICE pointers and no host<->device copies are made. */
int b200_ComputePerturbedField_device(float redshift, InitialConditions *d_boxes,
                                      PerturbedField *d_perturbed_field);
int b200_ComputeIonizedBox_device(float redshift, float prev_redshift,
                                  PerturbedField *d_perturbed_field, IonizedBox *d_box);
/* Slab-parallel particle deposit of ONE box across `nparts` GPUs (move_grid_masses,
   map_mass.c:146-208: the deposit is a sum over particles).  phase 0: rank `part` deposits the
   particles of its x-slab into d_acc (N 64-bit fixed-point integers, device, zeroed by the call);
   the caller all-reduces d_acc with SUM over the ranks (integer addition: bit-identical to the
   single-GPU deposit); phase 1: every rank normalises the merged accumulator and runs the
   density / velocity FFT chain (PerturbedField.c:212-387).  ZELDOVICH / 2LPT with an integer
   DIM / HII_DIM ratio only. */
int b200_ComputePerturbedField_device_part(float redshift, InitialConditions *d_boxes,
                                           PerturbedField *d_perturbed_field,
                                           unsigned long long *d_acc, int part, int nparts, int phase);
/* Radius-parallel ionisation of ONE box across `nparts` GPUs (one process per GPU).  The filter
   radii of find_HII_bubbles (IonisationBox.c:1531-1630) are independent given the k-space density
   and their ionised flags combine by OR, so rank `part` runs the radii k = part (mod nparts):
     phase 0: all assigned radii but the last one of the ladder -> d_mask (N bytes, 1 = ionised);
              the caller then all-reduces d_mask with MAX over the ranks (the only collective);
     phase 1: every rank runs the last radius (partial ionisations) on the merged mask and
              finalises: d_box is complete on every rank.
   Same device-pointer convention as b200_ComputeIonizedBox_device. */
int b200_ComputeIonizedBox_device_part(float redshift, float prev_redshift,
                                       PerturbedField *d_perturbed_field, IonizedBox *d_box,
                                       unsigned char *d_mask, int part, int nparts, int phase);

/* ---- ONE box over the GPUs of a node, slab-decomposed (replaces dft_r2c_cube /
   dft_c2r_cube, dft.c:18-72, and move_grid_masses, map_mass.c:146-208, for a box split over GPUs) ------
   One process per GPU.  b200_dist_init allocates this rank's symmetric heap (peer-mapped device memory:
   the transposes of the slab FFT and the halo exchange of the deposit are kernel stores / loads on peer
   pointers over NVLink, ordered by a barrier kernel on the library's stream -- no NCCL call) and returns
   a 64-byte handle; the caller exchanges the handles of all ranks by any means (torch.distributed in
   21cmfast_b200/distributed.py) and passes the concatenation, in rank order, to b200_dist_connect.
   HII_DIM must be a multiple of `world` (<= 8). */
int b200_dist_init(int rank, int world, unsigned long long heap_bytes, void *handle_out_64_bytes);
int b200_dist_connect(const void *handles_world_times_64_bytes);
int b200_dist_shutdown(void);
int b200_dist_rank(void);
int b200_dist_world(void);
int b200_dist_barrier(void);
/* Redshift-parallel runs on shared initial conditions: with enable != 0 the ranks of a
   connected group call ComputePerturbedField TOGETHER, each with its own redshift and output arrays but the same
   IC arrays in host memory; every rank then uploads 1 / world of each IC array over its own PCIe link and sends
   that share to the peers' heaps over NVLink (the heap must hold the ICs: 4 (DIM^3 + 6 HII_DIM^3) bytes).  On a
   node whose host-to-device rate is shared by the GPUs this divides the upload time by the number of ranks. */
void b200_ics_share(int enable);
/* Slab entry points: the structs hold DEVICE pointers to this rank's x-slab of every array, planes
   [rank, rank + 1) * HII_DIM / world: low-res boxes [HII_DIM / world][HII_DIM][HII_D_PARA]; hires_density
   the F * HII_DIM / world hi-res planes that start at global plane F * x0 - F / 2 (periodic), F = DIM /
   HII_DIM integer <= 4.  Outputs are the rank's slab of the result, bit-identical to the same planes of the
   single-GPU box (same per-line FFT arithmetic, integer deposit sums, fixed reduction tree of the grid
   sum).  LINEAR / ZELDOVICH / 2LPT on the low-res grid (LINEAR reads the lowres_density slab); ionisation without recombinations / spin temperature. */
int b200_ComputePerturbedField_slab(float redshift, InitialConditions *d_boxes_slab, PerturbedField *d_pf_slab);
int b200_ComputeIonizedBox_slab(float redshift, float prev_redshift, PerturbedField *d_pf_slab, IonizedBox *d_box_slab);
/* ComputeInitialConditions (InitialConditions.c:547-772) with the hi-res box split into x-slabs, so that DIM is
   bounded by the node's HBM instead of one GPU's: hires_density holds the DIM / world hi-res planes that start at
   DIM / world * rank (no half-cell shift), the low-res density / velocity boxes the rank's HII_DIM / world
   planes.  Bit-identical to the same planes of the single-GPU result, with the N_THREADS = 1 random stream of the
   reference (every rank walks the one stream and keeps the modes of its k-space slab) or B200_IC_RNG=device.
   Velocities on the low-res grid, integer DIM / HII_DIM, no relative velocities. */
int b200_ComputeInitialConditions_slab(unsigned long long random_seed, InitialConditions *d_boxes_slab);

/* Test hooks for the random stream of sample_ic_modes (InitialConditions.c:103-139; rng.c:31-90):
   n1 then n2 values of gsl_ran_ugaussian on gsl_rng_mt19937 seeded with mt_seed, produced by the
   device pipeline (csrc/gslrng.cu) resp. by the sequential host generator; and the sequential
   conversion rule on a given array of raw 32-bit words (returns the words consumed, -1 if short). */
int b200_gsl_gaussian_stream(unsigned long mt_seed, long long n1, long long n2, double *host_out);
int b200_host_gaussian_stream(unsigned long mt_seed, long long n, double *host_out);
long long b200_gaussians_from_raw_host(const unsigned int *raw, long long n_raw, long long want, double *out);
/* Measurement hook (bench.py's cuFFT yardstick leg): mean device milliseconds of
   the library's own n^3 r2c, c2r and c2r-with-window transforms over `iters` repetitions. */
int b200_fft_probe(int n, int iters, double box_len, double *ms_r2c, double *ms_c2r, double *ms_c2r_window);
/* upper incomplete gamma function Gamma(a, x) for any real a as used by the GAMMA-APPROX integrals
   (gsl_sf_gamma_inc at hmf.c:733); test hook for the known-answer test against mpmath */
double b200_upper_gamma(double a, double x);
/* x-plane range [begin, end) of thread t of n_threads under the static schedule of sample_ic_modes
   (InitialConditions.c:103-134); test hook for the N_THREADS > 1 seed-parity path */
void b200_omp_static_range(int n, int n_threads, int t, int *begin, int *end);

#ifdef __cplusplus
}
#endif
#endif /* PY21CMFAST_B200_H */
