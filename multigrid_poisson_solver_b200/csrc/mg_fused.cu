// mg_fused.cu -- smoothing passes and the fused -1 / 1 cycle legs.
#include "mg_fused.h"

#include "mg_device.cuh"
#include "mg_kernels.h"

namespace mg {

void fused_init() {}

double *smooth_out_of_place(int N, double L, double *a, double *b, const double *F, int step, bool in_is_zero,
                            double *err_dev, double *err_slot)
{
    const Spacing sp = spacing(N, L);
    double *cur = a, *other = b;
    for (int s = 0; s < step; ++s) {
        if (s == 0 && in_is_zero) {
            launch_sweep(N, sp.h2, cur, F, cur, true);  // the input is implied zeros: safe in place
        } else {
            launch_sweep(N, sp.h2, cur, F, other, false);
            double *t = cur; cur = other; other = t;
        }
    }
    if (step == 0 && in_is_zero)
        check(cudaMemsetAsync(cur, 0, (size_t)N * N * sizeof(double), ctx().stream), "cudaMemsetAsync");
    if (err_dev || err_slot) launch_smooth_error(N, sp.inv_h2, cur, F, err_dev, err_slot);
    return cur;
}

double *down_leg(int N, double L, double *U, double *U_work, const double *F, int step, bool zero_init, int M,
                 double *F_c, double *err_slot)
{
    const Spacing sp = spacing(N, L);
    double *res = smooth_out_of_place(N, L, U, U_work, F, step, zero_init, step > 0 ? ctx().dev_scalar : nullptr,
                                      step > 0 ? err_slot : nullptr);
    double *D = scratch_grid((size_t)N * N);
    if (!D) return res;
    launch_residual(N, sp.inv_h2, res, F, D);
    launch_negate(N, D);
    launch_restrict(N, D, M, F_c);
    return res;
}

double *up_leg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int step,
               double *err_slot)
{
    launch_prolong(Nc, U_c, N, U_f, U_f);
    if (step <= 0) return U_f;
    return smooth_out_of_place(N, L, U_f, U_work, F, step, false, ctx().dev_scalar, err_slot);
}

}  // namespace mg
