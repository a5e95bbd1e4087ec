/* Oracle shim for <fftw3.h> -- TEST INFRASTRUCTURE ONLY.
 * FFTW3 (single precision) is a third-party dependency of the reference (build_cffi.py:109,
 * call sites dft.c:38-40,66-68) that is absent from /root/reference and from this image.  This
 * header declares the dozen entry points the reference uses; fftw_shim.c implements the two
 * in-place 3-D plans (r2c / c2r, unnormalised, FFTW's padded real layout) on top of either an
 * own mixed-radix FFT or Intel MKL's DFTI as exported by torch's libtorch_cpu.so. */
#ifndef ORACLE_FFTW3_H
#define ORACLE_FFTW3_H
#include <complex.h>
#include <stddef.h>
typedef float _Complex fftwf_complex;
typedef struct oracle_fftwf_plan_s *fftwf_plan;
#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_WISDOM_ONLY (1U << 21)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftwf_cleanup(void);
void fftwf_cleanup_threads(void);
void fftwf_forget_wisdom(void);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int nthreads);
int fftwf_import_wisdom_from_filename(const char *filename);
int fftwf_export_wisdom_to_filename(const char *filename);
#endif
