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

// ---- 1-D transfer maps (shared by the table kernels and the coarse-tail kernel)
// doRestriction's map of coarse index t (MG_solver_CPU.cpp:661-664): lower fine index and weight.
__device__ __forceinline__ void restrict_table_entry(int t, int N, int M, int &lo, double &w)
{
    const double h_f = __ddiv_rn(1.0, (double)(N - 1)), h_c = __ddiv_rn(1.0, (double)(M - 1));
    const double pos = __dmul_rn((double)t, h_c);
    lo = (int)floor(__ddiv_rn(pos, h_f));
    w = __ddiv_rn(fmod(pos, h_f), h_f);
}

// doProlongation in gather form (MG_solver_CPU.cpp:688-718): the coarse cell q
// that owns fine index t is the one with ceil(q*ratio) <= t < ceil((q+1)*ratio), or -1.
__device__ __forceinline__ int prolong_cell_of(int t, int N, double ratio)
{
    int q = (int)floor(__ddiv_rn((double)t, ratio));
    q = max(0, min(q, N - 2));
    while (q > 0 && ceil(__dmul_rn((double)q, ratio)) > (double)t) --q;
    while (q < N - 2 && ceil(__dmul_rn((double)(q + 1), ratio)) <= (double)t) ++q;
    const bool inside = ceil(__dmul_rn((double)q, ratio)) <= (double)t && (double)t < ceil(__dmul_rn((double)(q + 1), ratio));
    return inside ? q : -1;
}

// Row and column entry of fine index t (N coarse, M fine): cell and weights {c_hi - f, f - c_lo}.
// Rows and columns differ only in the patched last line (:701-718).
__device__ __forceinline__ void prolong_table_entry(int t, int N, int M, int &row_cell, int &col_cell, double2 &row_w,
                                                    double2 &col_w)
{
    const double c_dx = __ddiv_rn(1.0, (double)(N - 1)), f_dx = __ddiv_rn(1.0, (double)(M - 1));
    const double ratio = __ddiv_rn(c_dx, f_dx);
    int rq, cq;
    double rf, cf;
    if (t < M - 1) {
        rq = cq = prolong_cell_of(t, N, ratio);
        if (rq < 0) rq = cq = max(0, min((int)floor(__ddiv_rn((double)t, ratio)), N - 2));  // never written by the reference
        rf = cf = __dmul_rn((double)t, f_dx);
    } else {
        int q1 = prolong_cell_of(M - 2, N, ratio);
        if (q1 < 0) q1 = N - 2;
        const int q2 = prolong_cell_of(M - 1, N, ratio);
        const double f_last = __dmul_rn((double)(M - 1), f_dx);
        rq = (q2 >= 0 && q2 != q1) ? q2 : q1;  // :706-718 (the patch ends the row loop of its own cell only)
        rf = f_last;
        if (q2 >= 0) { cq = q2; cf = f_last; }  // regular pass rewrites the patched column (last writer wins)
        else         { cq = q1; cf = 1.0; }     // :701-704
    }
    const double r_lo = __dmul_rn((double)rq, c_dx), c_lo = __dmul_rn((double)cq, c_dx);
    row_cell = rq;
    col_cell = cq;
    row_w = make_double2(__dsub_rn(__dadd_rn(r_lo, c_dx), rf), __dsub_rn(rf, r_lo));
    col_w = make_double2(__dsub_rn(__dadd_rn(c_lo, c_dx), cf), __dsub_rn(cf, c_lo));
}

// Red/black Gauss-Seidel (MG_solver_CPU.cpp:952-1066) on a grid held in shared memory, run by the
// first `warps` warps of a CTA (all of their threads must call; other warps must not).  u must be
// zero on entry (:993).  Three named barriers per iteration, convergence test every iteration as
// in the reference; the mean |residual| is folded in a fixed order (warp tree, then warp order).
// Returns the iteration count (same value in every calling thread).
__device__ __forceinline__ void named_barrier(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int MAX_PTS>   // points per thread: N*N <= MAX_PTS * 32 * warps
__device__ int gauss_seidel_shared(int N, double h2, double inv_h2, double target, double *u, const double *f, double *partials,
                                   int warps, int bar_id, int max_iters)
{
    const int T = warps * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = N * N;
    int cell[MAX_PTS];
    unsigned red_mask = 0, in_mask = 0;
#pragma unroll
    for (int k = 0; k < MAX_PTS; ++k) {
        const int c = tid + k * T;
        cell[k] = c < n ? c : 0;
        const int i = cell[k] / N, j = cell[k] - i * N;
        if (c < n && i > 0 && i < N - 1 && j > 0 && j < N - 1) {
            in_mask |= 1u << k;
            if (((i + j) & 1) == 0) red_mask |= 1u << k;   // the ieven table (:972-980) enumerates (ix+iy) even
        }
    }
    const double denom = (double)((N - 2) * (N - 2));
    int it = 0;
    double e;
    do {
#pragma unroll
        for (int colour = 0; colour < 2; ++colour) {
            const unsigned m = colour == 0 ? red_mask : (in_mask & ~red_mask);
#pragma unroll
            for (int k = 0; k < MAX_PTS; ++k)
                if (m >> k & 1u) {
                    const int c = cell[k];
                    u[c] = gauss_seidel_at(u[c - 1], u[c + 1], u[c + N], u[c - N], __dmul_rn(h2, f[c]));
                }
            if (warps == 1) __syncwarp(); else named_barrier(bar_id, T);
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < MAX_PTS; ++k)
            if (in_mask >> k & 1u) {
                const int c = cell[k];
                acc = __dadd_rn(acc, fabs(residual_at(u[c], sum4(u[c + N], u[c - N], u[c + 1], u[c - 1]), f[c], inv_h2)));
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        if (warps > 1) {
            // double-buffered by iteration parity: no barrier is needed before the next write
            double *slot = partials + (it & 1) * 32;
            if (lane == 0) slot[warp] = acc;
            named_barrier(bar_id, T);
            acc = 0.0;
            for (int w = 0; w < warps; ++w) acc = __dadd_rn(acc, slot[w]);
        }
        e = __ddiv_rn(acc, denom);                                    // :1059
        ++it;
    } while (e > target && it < max_iters);
    return it;
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
