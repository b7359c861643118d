// mg_kernels.h -- launchers of the sm_100a kernels (internal C++ API behind include/mg_abi.h).
#pragma once
#include <cuda_runtime.h>

#include "mg_context.h"

namespace mg {

// Spacing-derived constants, computed on the HOST with the same libm call the reference
// makes (pow(dx,2) at -O0 is a real call and differs from dx*dx by 1 ulp for some N;
// SURVEY.md 0.3).  MG_solver_CPU.cpp:555,560,574,590,611.
struct Spacing {
    double dx, h2, inv_h2;
};
Spacing spacing(int N, double L);

// ---- baseline kernels: one reference operator each
// rows [row0, row0+rows) of the grid into an array that starts at row0 (rows < 0: the whole grid)
void launch_source(int N, double L, double *F, double min_x, double min_y, bool analytic, int row0 = 0, int rows = -1);
void launch_residual(int N, double inv_h2, const double *U, const double *F, double *D);
void launch_add(int N, double *U1, const double *U2);
void launch_negate(int N, double *D);
void launch_sweep(int N, double h2, const double *U_in, const double *F, double *U_out, bool in_is_zero);
// S = sum over interior red points |inv_h2*(sum4-4U)-F|; result (S+S)/N/N -> *out_dev and/or pinned slot
void launch_smooth_error(int N, double inv_h2, const double *U, const double *F, double *out_dev, double *out_slot_dev);
void launch_restrict(int N, const double *U_f, int M, double *U_c);
// U_f = P(U_c)            (add_to == nullptr)
// U_f = add_to + P(U_c)   (otherwise; add_to may alias U_f)
void launch_prolong(int N, const double *U_c, int M, double *U_f, const double *add_to);
void launch_mean_abs_diff(size_t n, const double *A, const double *B, double denom, double *out_dev);

// ---- exact solvers
// iters_slot: optional device alias of a pinned scalar slot that receives the iteration count
void launch_gauss_seidel(int N, double L, double *U, const double *F, double target, double *iters_slot);
void launch_inverse_matrix(int N, double L, double *U, const double *F);

const RestrictTable &restrict_table(int N, int M);
const ProlongTable &prolong_table(int N, int M);

}  // namespace mg
