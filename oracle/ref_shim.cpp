/*
 * ref_shim.cpp -- builds the UNMODIFIED reference translation unit into a
 * shared library (oracle/_ref/libmgref.so) so tests and bench.py can call the
 * reference's own operators and its own main().
 *
 * TEST INFRASTRUCTURE ONLY.  No reference source is copied into this repo: the
 * reference file is #included from where it lies (the Makefile passes
 * -I$(REF)/src) and is compiled with the reference's own flags (src/Makefile:8:
 * g++ -fopenmp, no -O).  Its main() is renamed so it can be called as a function.
 */
#define main mg_reference_main
#include "MG_solver_CPU.cpp"
#undef main

extern "C" {
void ref_getSource(int N, double L, double *F, double mx, double my) { getSource(N, L, F, mx, my); }
void ref_getBoundary(int N, double L, double *F, double mx, double my) { getBoundary(N, L, F, mx, my); }
void ref_getAnalytic(int N, double L, double *U, double mx, double my) { getAnalytic(N, L, U, mx, my); }
void ref_getResidual(int N, double L, double *U, double *F, double *D) { getResidual(N, L, U, F, D); }
void ref_doGridAddition(int N, double *U1, double *U2) { doGridAddition(N, U1, U2); }
void ref_doSmoothing(int N, double L, double *U, double *F, int step, double *err) { doSmoothing(N, L, U, F, step, err); }
void ref_doExactSolver(int N, double L, double *U, double *F, double tol, int opt) { doExactSolver(N, L, U, F, tol, opt); }
void ref_doRestriction(int N, double *Uf, int M, double *Uc) { doRestriction(N, Uf, M, Uc); }
void ref_doProlongation(int N, double *Uc, int M, double *Uf) { doProlongation(N, Uc, M, Uf); }
void ref_set_threads(int n) { omp_set_num_threads(n); }
/* ./MG_CPU <threads> <cycle file> as a function call (writes Sol_CPU_<file> into the cwd) */
int ref_main(int threads, const char *cycle_file)
{
    char a0[] = "MG_CPU", a1[16];
    snprintf(a1, sizeof a1, "%d", threads);
    char *argv[] = {a0, a1, const_cast<char *>(cycle_file), nullptr};
    return mg_reference_main(3, argv);
}
}
