// mg_kernels.cu -- baseline sm_100a kernels: one per reference operator, generic in N
// (odd sizes, 64-bit indexing).  The fused / temporally blocked kernels live in
// mg_fused.cu; these remain the fallback for shapes the fused kernels do not cover
// and the implementation of the eight stand-alone ABI operators.
//
// Bit-exactness contract: every expression uses the explicitly rounded
// intrinsics (__dadd_rn, __dmul_rn, ...) in the reference's association, so nvcc can
// neither contract to FMA nor reassociate; the file is also compiled with -fmad=false.
#include "mg_kernels.h"

#include <cmath>
#include <cstdio>

#include "mg_device.cuh"

namespace mg {

Spacing spacing(int N, double L)
{
    volatile double two = 2.0;  // keep pow() a libm call (never dx*dx)
    Spacing s;
    s.dx = L / (double)(N - 1);
    s.h2 = pow(s.dx, two);
    s.inv_h2 = 1.0 / pow(s.dx, two);
    return s;
}

namespace {

constexpr int TX = 256;   // threads along the fast index
constexpr int ROWS = 16;  // rows streamed by one CTA

struct Tiling {
    int col_blocks, row_blocks;
    unsigned blocks() const { return (unsigned)col_blocks * (unsigned)row_blocks; }
};
Tiling tiling(int N, int rows = ROWS) { return {(N + TX - 1) / TX, (N + rows - 1) / rows}; }

#define MG_LAUNCH(kernel, grid, block, smem, ...)                                   \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), ctx().stream>>>(__VA_ARGS__);             \
        ctx().launches++;                                                           \
        check(cudaGetLastError(), #kernel);                                         \
    } while (0)

// ------------------------------------------------------------------ source / analytic
// MG_solver_CPU.cpp:468-493 (source), :525-548 (analytic).  exp() is CUDA's (<= 1 ulp from glibc).
template <bool ANALYTIC>
__global__ void __launch_bounds__(TX) k_problem(int N, int col_blocks, double h, double min_x, double min_y,
                                                double *__restrict__ out, int row0)
{
    const int j = (blockIdx.x % col_blocks) * TX + threadIdx.x;  // ix
    const int i = row0 + blockIdx.x / col_blocks;                // iy (global); the array starts at row0
    if (j >= N) return;
    double v = 0.0;
    if (i > 0 && i < N - 1 && j > 0 && j < N - 1) {
        const double x = __dadd_rn(__dmul_rn((double)j, h), min_x);
        const double y = __dadd_rn(__dmul_rn((double)i, h), min_y);
        const double e = exp(__dsub_rn(x, y));
        if (ANALYTIC) {
            // exp(x-y)*x*(1.0-x)*y*(1.0-y)
            v = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(e, x), __dsub_rn(1.0, x)), y), __dsub_rn(1.0, y));
        } else {
            // 2.0*x*(y-1)*(y - 2.0*x + x*y + 2.0)*exp(x-y)
            const double poly = __dadd_rn(__dadd_rn(__dsub_rn(y, __dmul_rn(2.0, x)), __dmul_rn(x, y)), 2.0);
            v = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, x), __dsub_rn(y, 1.0)), poly), e);
        }
    }
    out[(size_t)(i - row0) * N + j] = v;
}

// ------------------------------------------------------------------ residual
// MG_solver_CPU.cpp:554-564.  One thread per column, ROWS rows streamed with a 3-row register window.
__global__ void __launch_bounds__(TX) k_residual(int N, int col_blocks, double inv_h2, const double *__restrict__ U,
                                                 const double *__restrict__ F, double *__restrict__ D)
{
    const int j = (blockIdx.x % col_blocks) * TX + threadIdx.x;
    const int i0 = (blockIdx.x / col_blocks) * ROWS;
    if (j >= N) return;
    const int i1 = min(i0 + ROWS, N);
    const bool col_in = j > 0 && j < N - 1;
    size_t c = (size_t)i0 * N + j;
    double below = i0 > 0 ? U[c - N] : 0.0, here = U[c];
    for (int i = i0; i < i1; ++i, c += N) {
        const double above = (i + 1 < N) ? U[c + N] : 0.0;
        double out = 0.0;
        if (col_in && i > 0 && i < N - 1) out = residual_at(here, sum4(above, below, U[c + 1], U[c - 1]), F[c], inv_h2);
        D[c] = out;
        below = here;
        here = above;
    }
}

// ------------------------------------------------------------------ one Jacobi sweep, out of place
// MG_solver_CPU.cpp:578-601 (both half sweeps read U_old => Jacobi).  Boundary values are carried over.
template <bool IN_IS_ZERO>
__global__ void __launch_bounds__(TX) k_sweep(int N, int col_blocks, double h2, const double *__restrict__ Uin,
                                              const double *__restrict__ F, double *__restrict__ Uout)
{
    const int j = (blockIdx.x % col_blocks) * TX + threadIdx.x;
    const int i0 = (blockIdx.x / col_blocks) * ROWS;
    if (j >= N) return;
    const int i1 = min(i0 + ROWS, N);
    const bool col_in = j > 0 && j < N - 1;
    size_t c = (size_t)i0 * N + j;
    if (IN_IS_ZERO) {
        for (int i = i0; i < i1; ++i, c += N) {
            double out = 0.0;
            if (col_in && i > 0 && i < N - 1) out = jacobi_at(0.0, sum4(0.0, 0.0, 0.0, 0.0), __dmul_rn(h2, F[c]));
            Uout[c] = out;
        }
        return;
    }
    double below = i0 > 0 ? Uin[c - N] : 0.0, here = Uin[c];
    for (int i = i0; i < i1; ++i, c += N) {
        const double above = (i + 1 < N) ? Uin[c + N] : 0.0;
        double out = here;
        if (col_in && i > 0 && i < N - 1)
            out = jacobi_at(here, sum4(above, below, Uin[c + 1], Uin[c - 1]), __dmul_rn(h2, F[c]));
        Uout[c] = out;
        below = here;
        here = above;
    }
}

// ------------------------------------------------------------------ smoothing error
// MG_solver_CPU.cpp:607-622: the SAME red-parity sum twice.  Deterministic: fixed per-thread
// order, fixed tree per CTA, last CTA folds the per-CTA partials in a fixed order.
__global__ void __launch_bounds__(TX) k_smooth_error(int N, int col_blocks, double inv_h2, const double *__restrict__ U,
                                                     const double *__restrict__ F, double *__restrict__ partials,
                                                     unsigned int *counter, double *out_dev, double *out_slot)
{
    __shared__ double red_smem[32];
    __shared__ bool is_last;
    const int j = (blockIdx.x % col_blocks) * TX + threadIdx.x;
    const int i0 = (blockIdx.x / col_blocks) * ROWS;
    double acc = 0.0;
    if (j > 0 && j < N - 1) {
        const int lo = max(i0, 1), hi = min(i0 + ROWS, N - 1);
        int i = lo + ((lo + j) & 1);  // first row with (i+j) even
        size_t c = (size_t)i * N + j;
        for (; i < hi; i += 2, c += 2 * (size_t)N)
            acc = __dadd_rn(acc, fabs(residual_at(U[c], sum4(U[c + N], U[c - N], U[c + 1], U[c - 1]), F[c], inv_h2)));
    }
    const double total = block_sum<TX>(acc, red_smem);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = total;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += TX) s = __dadd_rn(s, __ldcg(&partials[k]));
    s = block_sum<TX>(s, red_smem);
    if (threadIdx.x == 0) {
        double e = __dadd_rn(s, s);  // sum1 + sum2
        e = __ddiv_rn(e, (double)N);
        e = __ddiv_rn(e, (double)N);
        if (out_dev) *out_dev = e;
        if (out_slot) { *out_slot = e; __threadfence_system(); }
        *counter = 0u;
    }
}

// ------------------------------------------------------------------ elementwise
__global__ void __launch_bounds__(256) k_add(size_t n, double *__restrict__ a, const double *__restrict__ b)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) a[k] = __dadd_rn(a[k], b[k]);
}
__global__ void __launch_bounds__(256) k_add2(size_t n2, double2 *__restrict__ a, const double2 *__restrict__ b)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        double2 x = a[k];
        const double2 y = b[k];
        x.x = __dadd_rn(x.x, y.x);
        x.y = __dadd_rn(x.y, y.y);
        a[k] = x;
    }
}
__global__ void __launch_bounds__(256) k_negate(size_t n, double *__restrict__ a)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) a[k] = -a[k];
}

// ------------------------------------------------------------------ restriction
// MG_solver_CPU.cpp:661-666: the x and y maps are the same 1-D function of the coarse index.
__global__ void k_restrict_table(int N, int M, int *lo, double *w)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    restrict_table_entry(t, N, M, lo[t], w[t]);
}

// MG_solver_CPU.cpp:673-676
__global__ void __launch_bounds__(TX) k_restrict(int N, int M, int col_blocks, const double *__restrict__ Uf,
                                                 double *__restrict__ Uc, const int *__restrict__ lo,
                                                 const double *__restrict__ w)
{
    const int ix = (blockIdx.x % col_blocks) * TX + threadIdx.x;
    const int iy = blockIdx.x / col_blocks;
    if (ix >= M) return;
    double v = 0.0;
    if (ix > 0 && ix < M - 1 && iy > 0 && iy < M - 1) {
        const double a = w[ix], cw = w[iy];
        const size_t f = (size_t)lo[ix] + (size_t)lo[iy] * N;
        v = restrict_at(Uf[f], Uf[f + 1], Uf[f + N], Uf[f + N + 1], a, cw);
    }
    Uc[(size_t)iy * M + ix] = v;
}

// ------------------------------------------------------------------ prolongation
__global__ void k_prolong_table(int N, int M, int *row_cell, int *col_cell, double2 *row_w, double2 *col_w, double4 *row_info)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    prolong_table_entry(t, N, M, row_cell[t], col_cell[t], row_w[t], col_w[t]);
    row_info[t] = make_double4(row_w[t].x, row_w[t].y, (double)row_cell[t], 0.0);
}

template <bool ADD>
__global__ void __launch_bounds__(TX) k_prolong(int N, int M, int col_blocks, double c_dx, const double *__restrict__ Uc,
                                                double *Uf, const double *add_to, const int *__restrict__ row_cell,
                                                const int *__restrict__ col_cell, const double2 *__restrict__ row_w,
                                                const double2 *__restrict__ col_w)
{
    const int l = (blockIdx.x % col_blocks) * TX + threadIdx.x;
    const int k = blockIdx.x / col_blocks;
    if (l >= M) return;
    const int q = col_cell[l];
    const double2 wx = col_w[l], wy = row_w[k];
    const double *lo_row = Uc + (size_t)row_cell[k] * N + q;
    double v = prolong_at(lo_row[0], lo_row[1], lo_row[N], lo_row[N + 1], wx, wy, c_dx);
    const size_t o = (size_t)k * M + l;
    if (ADD) v = __dadd_rn(add_to[o], v);
    Uf[o] = v;
}

// ------------------------------------------------------------------ mean |A-B| (final report :441-445)
__global__ void __launch_bounds__(TX) k_abs_diff(size_t n, const double *__restrict__ A, const double *__restrict__ B,
                                                 double denom, double *__restrict__ partials, unsigned int *counter,
                                                 double *out_dev)
{
    __shared__ double red_smem[32];
    __shared__ bool is_last;
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * TX;
    for (size_t k = (size_t)blockIdx.x * TX + threadIdx.x; k < n; k += stride) acc = __dadd_rn(acc, fabs(__dsub_rn(A[k], B[k])));
    const double total = block_sum<TX>(acc, red_smem);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = total;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += TX) s = __dadd_rn(s, __ldcg(&partials[k]));
    s = block_sum<TX>(s, red_smem);
    if (threadIdx.x == 0) {
        *out_dev = __ddiv_rn(s, denom);
        *counter = 0u;
    }
}

}  // namespace

// =============================================================================== launchers

void launch_source(int N, double L, double *F, double min_x, double min_y, bool analytic, int row0, int rows)
{
    const double h = L / (double)(N - 1);
    const int cb = (N + TX - 1) / TX;
    if (rows < 0) rows = N;
    const unsigned blocks = (unsigned)cb * (unsigned)rows;
    if (analytic) MG_LAUNCH(k_problem<true>, blocks, TX, 0, N, cb, h, min_x, min_y, F, row0);
    else          MG_LAUNCH(k_problem<false>, blocks, TX, 0, N, cb, h, min_x, min_y, F, row0);
}

void launch_residual(int N, double inv_h2, const double *U, const double *F, double *D)
{
    const Tiling t = tiling(N);
    MG_LAUNCH(k_residual, t.blocks(), TX, 0, N, t.col_blocks, inv_h2, U, F, D);
}

void launch_sweep(int N, double h2, const double *U_in, const double *F, double *U_out, bool in_is_zero)
{
    const Tiling t = tiling(N);
    if (in_is_zero) MG_LAUNCH(k_sweep<true>, t.blocks(), TX, 0, N, t.col_blocks, h2, U_in, F, U_out);
    else            MG_LAUNCH(k_sweep<false>, t.blocks(), TX, 0, N, t.col_blocks, h2, U_in, F, U_out);
}

void launch_smooth_error(int N, double inv_h2, const double *U, const double *F, double *out_dev, double *out_slot_dev)
{
    const Tiling t = tiling(N);
    double *partials = partials_buf(t.blocks());
    MG_LAUNCH(k_smooth_error, t.blocks(), TX, 0, N, t.col_blocks, inv_h2, U, F, partials, ctx().counters, out_dev,
              out_slot_dev);
}

static unsigned elementwise_blocks(size_t n)
{
    const size_t want = (n + 255) / 256;
    const size_t cap = (size_t)ctx().sm_count * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

void launch_add(int N, double *U1, const double *U2)
{
    const size_t n = (size_t)N * N;
    const bool vec = (n % 2 == 0) && (((uintptr_t)U1 | (uintptr_t)U2) % 16 == 0);
    if (vec) MG_LAUNCH(k_add2, elementwise_blocks(n / 2), 256, 0, n / 2, (double2 *)U1, (const double2 *)U2);
    else     MG_LAUNCH(k_add, elementwise_blocks(n), 256, 0, n, U1, U2);
}

void launch_negate(int N, double *D)
{
    const size_t n = (size_t)N * N;
    MG_LAUNCH(k_negate, elementwise_blocks(n), 256, 0, n, D);
}

const RestrictTable &restrict_table(int N, int M)
{
    auto &cache = ctx().restrict_tables;
    auto it = cache.find({N, M});
    if (it != cache.end()) return it->second;
    RestrictTable t;
    check(cudaMalloc(&t.lo, (size_t)M * sizeof(int)), "cudaMalloc restrict table");
    check(cudaMalloc(&t.w, (size_t)M * sizeof(double)), "cudaMalloc restrict table");
    MG_LAUNCH(k_restrict_table, (M + 255) / 256, 256, 0, N, M, t.lo, t.w);
    return cache.emplace(std::make_pair(N, M), t).first->second;
}

void launch_restrict(int N, const double *U_f, int M, double *U_c)
{
    const RestrictTable &t = restrict_table(N, M);
    const int cb = (M + TX - 1) / TX;
    MG_LAUNCH(k_restrict, (unsigned)cb * (unsigned)M, TX, 0, N, M, cb, U_f, U_c, t.lo, t.w);
}

const ProlongTable &prolong_table(int N, int M)
{
    auto &cache = ctx().prolong_tables;
    auto it = cache.find({N, M});
    if (it != cache.end()) return it->second;
    ProlongTable t;
    check(cudaMalloc(&t.row_cell, (size_t)M * sizeof(int)), "cudaMalloc prolong table");
    check(cudaMalloc(&t.col_cell, (size_t)M * sizeof(int)), "cudaMalloc prolong table");
    check(cudaMalloc(&t.row_w, (size_t)M * sizeof(double2)), "cudaMalloc prolong table");
    check(cudaMalloc(&t.col_w, (size_t)M * sizeof(double2)), "cudaMalloc prolong table");
    check(cudaMalloc(&t.row_info, (size_t)M * sizeof(double4)), "cudaMalloc prolong table");
    MG_LAUNCH(k_prolong_table, (M + 255) / 256, 256, 0, N, M, t.row_cell, t.col_cell, t.row_w, t.col_w, t.row_info);
    return cache.emplace(std::make_pair(N, M), t).first->second;
}

void launch_prolong(int N, const double *U_c, int M, double *U_f, const double *add_to)
{
    const ProlongTable &t = prolong_table(N, M);
    const int cb = (M + TX - 1) / TX;
    const double c_dx = 1.0 / (double)(N - 1);
    if (add_to)
        MG_LAUNCH(k_prolong<true>, (unsigned)cb * (unsigned)M, TX, 0, N, M, cb, c_dx, U_c, U_f, add_to, t.row_cell,
                  t.col_cell, t.row_w, t.col_w);
    else
        MG_LAUNCH(k_prolong<false>, (unsigned)cb * (unsigned)M, TX, 0, N, M, cb, c_dx, U_c, U_f, add_to, t.row_cell,
                  t.col_cell, t.row_w, t.col_w);
}

void launch_mean_abs_diff(size_t n, const double *A, const double *B, double denom, double *out_dev)
{
    const unsigned blocks = elementwise_blocks(n);
    double *partials = partials_buf(blocks);
    MG_LAUNCH(k_abs_diff, blocks, TX, 0, n, A, B, denom, partials, ctx().counters, out_dev);
}

}  // namespace mg
