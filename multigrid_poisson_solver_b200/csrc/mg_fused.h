// mg_fused.h -- smoothing passes and the fused cycle legs (internal C++ API).
#pragma once
#include "mg_context.h"

namespace mg {

void fused_init();

// `step` Jacobi sweeps (MG_solver_CPU.cpp:578-601) followed by the smoothing error (:607-622).
// The first pass reads `in` (never written; treated as all zeros when in_is_zero) and the
// passes write alternately to `a`, `b`, `a`, ... (`in` may alias `b`).  Returns the buffer
// holding the result (`in` itself when step == 0).  The error goes to err_dev (device double)
// and, if non-null, to err_slot (device alias of a pinned scalar slot).
double *smooth_out_of_place(int N, double L, const double *in, double *a, double *b, const double *F, int step,
                            bool in_is_zero, double *err_dev, double *err_slot);
// number of out-of-place passes smooth_out_of_place will make
int smooth_pass_count(int N, int step);

// -1 node: [U = 0]; step sweeps; error; F_c = restrict(-(residual(U, F)))   (:246-287)
double *down_leg(int N, double L, double *U, double *U_work, const double *F, int step, bool zero_init, int M,
                 double *F_c, double *err_slot);

// 1 node: U_f += prolong(U_c); step sweeps; error                            (:350-416)
double *up_leg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int step,
               double *err_slot);

}  // namespace mg
