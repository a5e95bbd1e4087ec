"""ctypes mirror of the reference's C ABI for the grid hot path.

Field order and C types follow the reference's cffi headers and are an ABI contract
(byte-identical layout is required for the library to be a drop-in):

* input structs  -- src/py21cmfast/src/_inputparams_wrapper.h:11-182
* output structs -- src/py21cmfast/src/_outputstructs_wrapper.h:6-105
* enum integers  -- src/py21cmfast/src/InputParameters.h:9-57

The same layouts are declared for C callers in ``include/py21cmfast_b200.h``; the test
``tests/test_abi_layout.py`` checks that ctypes and the C header agree on every offset.
"""
import ctypes as C

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class CosmoParamsStruct(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "hlittle", "OMm", "OMl", "OMb", "POWER_INDEX", "OMn", "OMk", "OMr", "OMtot", "Y_He", "wl")]


class SimulationOptionsStruct(C.Structure):
    _fields_ = [
        ("HII_DIM", C.c_int), ("DIM", C.c_int), ("BOX_LEN", C.c_float),
        ("NON_CUBIC_FACTOR", C.c_float), ("N_THREADS", C.c_int),
        ("Z_HEAT_MAX", C.c_double), ("ZPRIME_STEP_FACTOR", C.c_double),
        ("SAMPLER_MIN_MASS", C.c_float), ("SAMPLER_BUFFER_FACTOR", C.c_double),
        ("N_COND_INTERP", C.c_int), ("N_PROB_INTERP", C.c_int), ("MIN_LOGPROB", C.c_double),
        ("HALOMASS_CORRECTION", C.c_double), ("PARKINSON_G0", C.c_double),
        ("PARKINSON_y1", C.c_double), ("PARKINSON_y2", C.c_double),
        ("INITIAL_REDSHIFT", C.c_float), ("DELTA_R_FACTOR", C.c_double),
        ("DENSITY_SMOOTH_RADIUS", C.c_double), ("DEXM_OPTIMIZE_MINMASS", C.c_double),
        ("DEXM_R_OVERLAP", C.c_double), ("CORR_STAR", C.c_double), ("CORR_SFR", C.c_double),
        ("CORR_LX", C.c_double), ("MIN_XE_FOR_FCOLL_IN_TAUX", C.c_double),
    ]


class MatterOptionsStruct(C.Structure):
    _fields_ = [
        ("USE_FFTW_WISDOM", C.c_bool), ("HMF", C.c_int), ("V_CB_MODEL", C.c_int),
        ("POWER_SPECTRUM", C.c_int), ("USE_INTERPOLATION_TABLES", C.c_int),
        ("PERTURB_ON_HIGH_RES", C.c_bool), ("PERTURB_ALGORITHM", C.c_int),
        ("MINIMIZE_MEMORY", C.c_bool), ("KEEP_3D_VELOCITIES", C.c_bool),
        ("DEXM_OPTIMIZE", C.c_bool), ("FILTER", C.c_int), ("HALO_FILTER", C.c_int),
        ("SMOOTH_EVOLVED_DENSITY_FIELD", C.c_bool), ("SOURCE_MODEL", C.c_int),
        ("SAMPLE_METHOD", C.c_int),
    ]


class AstroParamsStruct(C.Structure):
    _fields_ = [
        ("HII_EFF_FACTOR", C.c_float), ("F_STAR10", C.c_float), ("ALPHA_STAR", C.c_float),
        ("ALPHA_STAR_MINI", C.c_float), ("SIGMA_STAR", C.c_float),
        ("UPPER_STELLAR_TURNOVER_MASS", C.c_double), ("UPPER_STELLAR_TURNOVER_INDEX", C.c_double),
        ("F_STAR7_MINI", C.c_float), ("t_STAR", C.c_float), ("SIGMA_SFR_INDEX", C.c_double),
        ("SIGMA_SFR_LIM", C.c_double), ("L_X", C.c_double), ("L_X_MINI", C.c_double),
        ("SIGMA_LX", C.c_double), ("F_ESC10", C.c_float), ("ALPHA_ESC", C.c_float),
        ("F_ESC7_MINI", C.c_float), ("T_RE", C.c_float), ("M_TURN", C.c_float),
        ("R_BUBBLE_MAX", C.c_float), ("ION_Tvir_MIN", C.c_float), ("F_H2_SHIELD", C.c_double),
        ("NU_X_THRESH", C.c_float), ("X_RAY_SPEC_INDEX", C.c_float), ("X_RAY_Tvir_MIN", C.c_float),
        ("A_LW", C.c_double), ("BETA_LW", C.c_double), ("A_VCB", C.c_double),
        ("BETA_VCB", C.c_double), ("V_CB_AVG_DEBUG", C.c_double), ("POP2_ION", C.c_double),
        ("POP3_ION", C.c_double), ("PHOTONCONS_CALIBRATION_END", C.c_double),
        ("CLUMPING_FACTOR", C.c_double), ("ALPHA_UVB", C.c_double), ("R_MAX_TS", C.c_float),
        ("N_STEP_TS", C.c_int), ("DELTA_R_HII_FACTOR", C.c_double), ("R_BUBBLE_MIN", C.c_float),
        ("MAX_DVDR", C.c_double), ("NU_X_MAX", C.c_double), ("NU_X_BAND_MAX", C.c_double),
    ]


class AstroOptionsStruct(C.Structure):
    _fields_ = [
        ("USE_MINI_HALOS", C.c_bool), ("USE_X_RAY_HEATING", C.c_bool),
        ("USE_CMB_HEATING", C.c_bool), ("USE_LYA_HEATING", C.c_bool), ("RECOMB_MODEL", C.c_int),
        ("USE_TS_FLUCT", C.c_bool), ("M_MIN_in_Mass", C.c_bool), ("USE_EXP_FILTER", C.c_bool),
        ("CELL_RECOMB", C.c_bool), ("LYA_MULTIPLE_SCATTERING", C.c_bool),
        ("USE_ADIABATIC_FLUCTUATIONS", C.c_bool), ("PHOTON_CONS_TYPE", C.c_int),
        ("USE_UPPER_STELLAR_TURNOVER", C.c_bool), ("HALO_SCALING_RELATIONS_MEDIAN", C.c_bool),
        ("HII_FILTER", C.c_int), ("HEAT_FILTER", C.c_int), ("IONISE_ENTIRE_SPHERE", C.c_bool),
        ("INTEGRATION_METHOD_ATOMIC", C.c_int), ("INTEGRATION_METHOD_MINI", C.c_int),
    ]


class Table1DStruct(C.Structure):
    _fields_ = [("size", C.c_int), ("x_values", c_double_p), ("y_values", c_double_p)]


class CosmoTablesStruct(C.Structure):
    _fields_ = [
        ("transfer_density", C.POINTER(Table1DStruct)), ("transfer_vcb", C.POINTER(Table1DStruct)),
        ("ps_norm", C.c_double), ("USE_SIGMA_8", C.c_bool), ("V_CB_AVG", C.c_double),
    ]


class ConfigSettingsStruct(C.Structure):
    _fields_ = [
        ("HALO_CATALOG_MEM_FACTOR", C.c_double), ("EXTRA_HALOBOX_FIELDS", C.c_bool),
        ("external_table_path", C.c_char_p), ("wisdoms_path", C.c_char_p),
    ]


class InitialConditionsStruct(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "lowres_density", "lowres_vx", "lowres_vy", "lowres_vz", "lowres_vx_2LPT",
        "lowres_vy_2LPT", "lowres_vz_2LPT", "hires_density", "hires_vx", "hires_vy", "hires_vz",
        "hires_vx_2LPT", "hires_vy_2LPT", "hires_vz_2LPT", "lowres_vcb")]


class PerturbedFieldStruct(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("density", "velocity_x", "velocity_y", "velocity_z")]


class HaloBoxStruct(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "halo_mass", "halo_stars", "halo_stars_mini", "count", "n_ion", "halo_sfr", "halo_xray",
        "halo_sfr_mini", "whalo_sfr")] + [
        ("log10_Mcrit_ACG_ave", C.c_double), ("log10_Mcrit_MCG_ave", C.c_double)]


class TsBoxStruct(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "spin_temperature", "xray_ionised_fraction", "kinetic_temp_neutral", "J_21_LW")] + [
        ("Q_HI", C.c_double)]


class IonizedBoxStruct(C.Structure):
    _fields_ = [
        ("mean_f_coll", C.c_double), ("mean_f_coll_MINI", C.c_double),
        ("log10_Mturnover_ave", C.c_double), ("log10_Mturnover_MINI_ave", C.c_double),
    ] + [(n, c_float_p) for n in (
        "neutral_fraction", "ionisation_rate_G12", "mean_free_path", "z_reion",
        "cumulative_recombinations", "kinetic_temperature", "unnormalised_nion",
        "unnormalised_nion_mini")]


class BrightnessTempStruct(C.Structure):
    _fields_ = [("brightness_temp", c_float_p), ("tau_21", c_float_p)]


# enum integer values (InputParameters.h:9-57)
HMF = {"PS": 0, "ST": 1, "WATSON": 2, "WATSON-Z": 3, "DELOS": 4, "REED07": 5, "YUNG24": 6}
POWER_SPECTRUM = {"EH": 0, "BBKS": 1, "EFSTATHIOU": 2, "PEEBLES": 3, "WHITE": 4, "CLASS": 5}
INTERPOLATION = {"no-interpolation": 0, "sigma-interpolation": 1, "hmf-interpolation": 2}
SAMPLE_METHOD = {"MASS-LIMITED": 0, "NUMBER-LIMITED": 1, "PARTITION": 2, "BINARY-SPLIT": 3}
FILTER = {"spherical-tophat": 0, "sharp-k": 1, "gaussian": 2}
PERTURB_ALGORITHM = {"LINEAR": 0, "ZELDOVICH": 1, "2LPT": 2}
SOURCE_MODEL = {"CONST-ION-EFF": 0, "E-INTEGRAL": 1, "L-INTEGRAL": 2, "DEXM-ESF": 3,
                "CHMF-SAMPLER": 4}
PHOTON_CONS = {"no-photoncons": 0, "z-photoncons": 1, "alpha-photoncons": 2, "f-photoncons": 3}
INTEGRATION_METHOD = {"GSL-QAG": 0, "GAUSS-LEGENDRE": 1, "GAMMA-APPROX": 2}
RECOMB_MODEL = {"none": 0, "homogeneous": 1, "inhomogeneous": 2}
V_CB_MODEL = {"NONE": 0, "AVG-AUTO": 1, "FLUCTS": 2, "AVG-DEBUG": 3}

# error codes returned by the Compute* functions (exceptions.h:12-21)
ERROR_CODES = {
    1: "IOError", 2: "GSLError", 3: "ValueError", 4: "PhotonConsError",
    5: "TableGenerationError", 6: "TableEvaluationError", 7: "InfinityorNaNError",
    8: "MassDepZetaError", 9: "MemoryAllocError", 10: "CUDAError",
}
