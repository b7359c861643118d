// mg_device.cuh -- per-point arithmetic of the reference operators, spelled with explicitly
// rounded intrinsics so the association (and therefore every bit) matches the -O0 CPU
// reference.  Shared by the baseline and the fused kernels.
#pragma once
#include <cuda_runtime.h>

namespace mg {

// ((up + down) + right) + left        MG_solver_CPU.cpp:560,590,611
__device__ __forceinline__ double sum4(double up, double down, double right, double left)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(up, down), right), left);
}

// U + 0.25*((sum4 - 4*U) - h2*F)      MG_solver_CPU.cpp:590 / :597   (h2f = pow(dx,2)*F, rounded once)
__device__ __forceinline__ double jacobi_at(double u, double s4, double h2f)
{
    const double bracket = __dsub_rn(__dsub_rn(s4, __dmul_rn(4.0, u)), h2f);
    return __dadd_rn(u, __dmul_rn(0.25, bracket));
}

// (1.0/pow(dx,2))*(sum4 - 4*U) - F    MG_solver_CPU.cpp:560 / :611
__device__ __forceinline__ double residual_at(double u, double s4, double f, double inv_h2)
{
    return __dsub_rn(__dmul_rn(inv_h2, __dsub_rn(s4, __dmul_rn(4.0, u))), f);
}

// 0.25*((((l + r) + t) + b) - h2*F)   MG_solver_CPU.cpp:1020 / :1043
__device__ __forceinline__ double gauss_seidel_at(double l, double r, double t, double b, double h2f)
{
    return __dmul_rn(0.25, __dsub_rn(__dadd_rn(__dadd_rn(__dadd_rn(l, r), t), b), h2f));
}

// b*d*U00 + a*d*U10 + c*b*U01 + a*c*U11, b = 1-a, d = 1-c      MG_solver_CPU.cpp:665-676
__device__ __forceinline__ double restrict_at(double u00, double u10, double u01, double u11, double a, double c)
{
    const double b = __dsub_rn(1.0, a), d = __dsub_rn(1.0, c);
    double v = __dmul_rn(__dmul_rn(b, d), u00);
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(a, d), u10));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(c, b), u01));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(a, c), u11));
    return v;
}

// ((c1*(c2x-fx) + c2*(fx-c1x))*(c3y-fy) + (c3*(c4x-fx) + c4*(fx-c3x))*(fy-c1y))/c_dx/c_dx   MG_solver_CPU.cpp:700
// wx = {c2x-fx, fx-c1x}, wy = {c3y-fy, fy-c1y}
__device__ __forceinline__ double prolong_at(double c1, double c2, double c3, double c4, double2 wx, double2 wy, double c_dx)
{
    const double bottom = __dadd_rn(__dmul_rn(c1, wx.x), __dmul_rn(c2, wx.y));
    const double top = __dadd_rn(__dmul_rn(c3, wx.x), __dmul_rn(c4, wx.y));
    const double v = __dadd_rn(__dmul_rn(bottom, wy.x), __dmul_rn(top, wy.y));
    return __ddiv_rn(__ddiv_rn(v, c_dx), c_dx);
}

// Deterministic CTA-wide sum (fixed shuffle tree, fixed warp order).  Result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *smem32)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
    __syncthreads();  // protect smem32 against a previous use
    if (lane == 0) smem32[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < THREADS / 32 ? smem32[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
    }
    return v;
}

}  // namespace mg
