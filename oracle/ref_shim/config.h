/* Hand-written config.h for the serial oracle build of the reference (test infrastructure only).
 * Mirrors what the reference's autoconf would define (configure.ac:81-205) for:
 * built-in FFT, single-rank MPI shim, OpenMP, no ScaLAPACK, BLAS from scipy's OpenBLAS. */
#ifndef QB200_ORACLE_CONFIG_H
#define QB200_ORACLE_CONFIG_H
#define FFT_NOLIB 1
#define USE_MPI 1
#define HAVE_OPENMP 1
#define PACKAGE_STRING "qball-oracle-serial"
#define PACKAGE_VERSION "oracle"
#define FC_FUNC(name,NAME) scipy_##name##_
#define FC_FUNC_(name,NAME) name##_
#endif
