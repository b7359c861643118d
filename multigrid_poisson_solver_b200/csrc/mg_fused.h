// mg_fused.h -- smoothing passes and the fused cycle legs (internal C++ API).
#pragma once
#include "mg_context.h"

namespace mg {

void fused_init();

// `step` Jacobi sweeps (MG_solver_CPU.cpp:578-601) followed by the smoothing error (:607-622).
// `a` holds the input (or is treated as all zeros when in_is_zero) and `b` is its ping-pong
// partner of the same size; both may be overwritten.  Returns the buffer holding the result.
// The error goes to err_dev (device double) and, if non-null, to err_slot (device alias of a
// pinned scalar slot).
double *smooth_out_of_place(int N, double L, double *a, double *b, const double *F, int step, bool in_is_zero,
                            double *err_dev, double *err_slot);

// -1 node: [U = 0]; step sweeps; error; F_c = restrict(-(residual(U, F)))   (:246-287)
double *down_leg(int N, double L, double *U, double *U_work, const double *F, int step, bool zero_init, int M,
                 double *F_c, double *err_slot);

// 1 node: U_f += prolong(U_c); step sweeps; error                            (:350-416)
double *up_leg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int step,
               double *err_slot);

}  // namespace mg
