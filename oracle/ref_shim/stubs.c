/* Stubs for the Fortran DFT-D3 library (only reached with `set vdw D3`); oracle build only. */
#include <stdio.h>
#include <stdlib.h>
void f90_dftd3_init_(const int* f) { (void)f; }
void f90_dftd3_end_(void) {}
void f90_dftd3_pbc_dispersion_(const int* n, const double* c, const int* z, const double* l, double* d, double* g, double* s)
{ (void)n; (void)c; (void)z; (void)l; (void)d; (void)g; (void)s; fprintf(stderr, "dftd3 stub called\n"); abort(); }
