// mg_exact.cu -- coarse-level exact solvers, entirely on the device (no host round trip per
// iteration, unlike the reference GPU program's cudaMemcpy of error partials every
// iteration, MG_solver_GPU.cu:1513-1521).
//
//   option 1  GaussSeidel   MG_solver_CPU.cpp:952-1066  red/black in place from U = 0 until
//                           mean |residual| over the interior <= target
//   option 0  InverseMatrix MG_solver_CPU.cpp:758-950   dense LU of the N^2 x N^2 operator
#include <climits>
#include <cstdio>

#include "mg_device.cuh"
#include "mg_kernels.h"

namespace mg {
namespace {

#define MG_LAUNCH(kernel, grid, block, smem, ...)                                   \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), ctx().stream>>>(__VA_ARGS__);             \
        ctx().launches++;                                                           \
        check(cudaGetLastError(), #kernel);                                         \
    } while (0)

constexpr int GS_MAX_ITERS = 200 * 1000 * 1000;

// ---------------------------------------------------------------- GS, whole grid in shared memory
// One CTA of `warps` warps (mg_device.cuh: gauss_seidel_shared): three named barriers per
// iteration, each thread owns up to PTS interior points.
template <int PTS>
__global__ void __launch_bounds__(1024) k_gs_smem(int N, int warps, double h2, double inv_h2, double target, double *__restrict__ U,
                          const double *__restrict__ F, int *__restrict__ iters_out, double *iters_slot)
{
    extern __shared__ double sm[];
    const int n = N * N, T = warps * 32;
    double *u = sm, *f = sm + n, *partials = sm + 2 * n;
    for (int k = threadIdx.x; k < n; k += T) {
        u[k] = 0.0;  // :993
        f[k] = F[k];
    }
    __syncthreads();
    const int it = gauss_seidel_shared<PTS>(N, h2, inv_h2, target, u, f, partials, warps, 1, GS_MAX_ITERS);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += T) U[k] = u[k];
    if (threadIdx.x == 0) {
        *iters_out = it;
        if (iters_slot) { *iters_slot = (double)it; __threadfence_system(); }
    }
}

// ---------------------------------------------------------------- GS, grid in global memory (large N)
// state[0] = done flag, state[1] = iteration count.  Every kernel is a no-op once done is set,
// so the host may enqueue batches of iterations and poll the flag between batches.
__global__ void __launch_bounds__(256) k_gs_colour(int N, int col_blocks, int colour, double h2, double *U,
                                                   const double *__restrict__ F, const int *state)
{
    if (state[0]) return;
    const int half = (blockIdx.x % col_blocks) * 256 + threadIdx.x;
    const int iy = 1 + blockIdx.x / col_blocks;
    const int ix = 1 + ((iy + 1 + colour) & 1) + 2 * half;
    if (ix >= N - 1) return;
    const size_t c = (size_t)ix + (size_t)iy * N;
    U[c] = gauss_seidel_at(U[c - 1], U[c + 1], U[c + N], U[c - N], __dmul_rn(h2, F[c]));
}

__global__ void __launch_bounds__(256) k_gs_check(int N, int col_blocks, double inv_h2, double target,
                                                  const double *__restrict__ U, const double *__restrict__ F,
                                                  double *partials, unsigned int *counter, int *state)
{
    __shared__ double red_smem[32];
    __shared__ bool is_last;
    if (state[0]) return;
    const int ix = 1 + (blockIdx.x % col_blocks) * 256 + threadIdx.x;
    const int iy = 1 + blockIdx.x / col_blocks;
    double acc = 0.0;
    if (ix < N - 1) {
        const size_t c = (size_t)ix + (size_t)iy * N;
        acc = fabs(residual_at(U[c], sum4(U[c + N], U[c - N], U[c + 1], U[c - 1]), F[c], inv_h2));
    }
    const double total = block_sum<256>(acc, red_smem);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = total;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += 256) s = __dadd_rn(s, __ldcg(&partials[k]));
    s = block_sum<256>(s, red_smem);
    if (threadIdx.x == 0) {
        const double err = __ddiv_rn(s, (double)(N - 2) * (double)(N - 2));
        state[1] += 1;
        if (!(err > target)) state[0] = 1;
        *counter = 0u;
    }
}

// ---------------------------------------------------------------- InverseMatrix (dense LU), one CTA
// A is never materialised: row r of the operator is the identity on the boundary and the
// 5-point stencil (-4/h2, 1/h2) elsewhere (:807-832).
__device__ __forceinline__ double lap_entry(int r, int c, int N, double diag, double off)
{
    const int i = r % N, j = r / N;
    if (i == 0 || i == N - 1 || j == 0 || j == N - 1) return r == c ? 1.0 : 0.0;
    if (c == r) return diag;
    if (c == r - 1 || c == r + 1 || c == r + N || c == r - N) return off;
    return 0.0;
}

__global__ void __launch_bounds__(1024) k_inverse_matrix(int N, double diag, double off, double *Lo, double *Up, int *P,
                                                         double *Z, double *X, const double *__restrict__ F)
{
    const int n = N * N;
    __shared__ int s_restart, s_checked, s_swap;
    const int T = blockDim.x, tid = threadIdx.x;
#define AT(m, r, c) m[(size_t)(r) * n + (c)]
    for (size_t k = tid; k < (size_t)n * n; k += T) { Lo[k] = 0.0; Up[k] = 0.0; }
    __syncthreads();
    for (int r = tid; r < n; r += T) { AT(Up, r, r) = 1.0; P[r] = r; X[r] = 0.0; Z[r] = 0.0; }
    if (tid == 0) { s_checked = 0; s_swap = 1; }
    __syncthreads();

    // :842-896 -- column sweep with the reference's restart-on-zero-pivot
    bool again;
    do {
        if (tid == 0) s_restart = 0;
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            for (int i = k + tid; i < n; i += T) {
                double acc = 0.0;
                for (int j = 0; j < k; ++j) acc = __dadd_rn(acc, __dmul_rn(AT(Lo, i, j), AT(Up, j, k)));
                AT(Lo, i, k) = __dsub_rn(lap_entry(P[i], k, N, diag, off), acc);
            }
            __syncthreads();
            if (tid == 0) {
                if (AT(Lo, k, k) == 0.0) {
                    if (s_swap >= n) {
                        printf("Having Zero Pivote ! det(A) = 0\n");
                    } else {
                        P[k] = s_swap; P[s_swap] = k; s_restart = 1;
                    }
                    s_swap += 1;
                } else {
                    s_checked += 1; s_swap = s_checked + 1;
                }
            }
            __syncthreads();
            if (s_restart) break;
            const double inv_piv = __ddiv_rn(1.0, AT(Lo, k, k));
            for (int j = k + tid; j < n; j += T) {
                double acc = 0.0;
                for (int i = 0; i < k; ++i) acc = __dadd_rn(acc, __dmul_rn(AT(Lo, k, i), AT(Up, i, j)));
                AT(Up, k, j) = __dmul_rn(inv_piv, __dsub_rn(lap_entry(P[k], j, N, diag, off), acc));
            }
            __syncthreads();
        }
        again = s_restart != 0;
        __syncthreads();
    } while (again);

    // :916-922 forward substitution; each row keeps its own running sum in column order
    {
        // the running sum of row i lives in Z[i] until the row is finalised
        for (int k = 0; k < n; ++k) {
            if (tid == 0) Z[k] = __dmul_rn(__ddiv_rn(1.0, AT(Lo, k, k)), __dsub_rn(F[k], Z[k]));
            __syncthreads();
            const double zk = Z[k];
            for (int i = k + 1 + tid; i < n; i += T) Z[i] = __dadd_rn(Z[i], __dmul_rn(AT(Lo, i, k), zk));
            __syncthreads();
        }
    }
    // :925-931 back substitution: the row sum runs over ascending columns, which only become
    // available last-to-first, so rows are serial; entries beyond the band (k > i+N) are exact
    // zeros of the factor and add nothing.
    if (tid == 0) {
        for (int i = n - 1; i >= 0; --i) {
            double acc = 0.0;
            const int kend = min(n, i + N + 1);
            for (int k = i + 1; k < kend; ++k) acc = __dadd_rn(acc, __dmul_rn(AT(Up, i, k), X[k]));
            X[i] = __dsub_rn(Z[i], acc);
        }
    }
#undef AT
}

}  // namespace

void launch_gauss_seidel(int N, double L, double *U, const double *F, double target, double *iters_slot)
{
    const Spacing sp = spacing(N, L);
    Context &c = ctx();
    const size_t smem = ((size_t)2 * N * N + 64) * sizeof(double);
    const int n = N * N;
    if (N >= 3 && smem <= 220 * 1024) {
        // as few warps as keep <= 4 (16 for the largest grids) points per thread: barriers dominate an iteration
        const int warps = n <= 128 ? 1 : n <= 512 ? 4 : n <= 1024 ? 8 : n <= 4096 ? 32 : 32;
#define GS_CASE(PTS)                                                                                            \
    do {                                                                                                        \
        static size_t opted = 0;                                                                                \
        if (smem > opted) {                                                                                     \
            check(cudaFuncSetAttribute(k_gs_smem<PTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                  "cudaFuncSetAttribute");                                                                      \
            opted = smem;                                                                                       \
        }                                                                                                       \
        MG_LAUNCH((k_gs_smem<PTS>), 1, warps * 32, smem, N, warps, sp.h2, sp.inv_h2, target, U, F, c.gs_iters, iters_slot); \
    } while (0)
        if (n <= 4096) GS_CASE(4);
        else GS_CASE(16);
#undef GS_CASE
        return;
    }
    // large grids: global-memory red/black with a device-side convergence flag
    int *state = c.gs_iters + 2;  // [done, iterations]
    check(cudaMemsetAsync(state, 0, 2 * sizeof(int), c.stream), "cudaMemsetAsync");
    check(cudaMemsetAsync(U, 0, (size_t)N * N * sizeof(double), c.stream), "cudaMemsetAsync");
    const int half_cols = (N - 2 + 1) / 2, cb_half = (half_cols + 255) / 256, cb_full = (N - 2 + 255) / 256;
    const unsigned rows = (unsigned)(N - 2);
    double *partials = partials_buf((size_t)cb_full * rows);
    int host_state[2] = {0, 0};
    while (!host_state[0] && host_state[1] < GS_MAX_ITERS && c.err_code == 0) {
        for (int b = 0; b < 32; ++b) {
            MG_LAUNCH(k_gs_colour, cb_half * rows, 256, 0, N, cb_half, 0, sp.h2, U, F, state);
            MG_LAUNCH(k_gs_colour, cb_half * rows, 256, 0, N, cb_half, 1, sp.h2, U, F, state);
            MG_LAUNCH(k_gs_check, cb_full * rows, 256, 0, N, cb_full, sp.inv_h2, target, U, F, partials, c.counters, state);
        }
        check(cudaMemcpyAsync(host_state, state, sizeof host_state, cudaMemcpyDeviceToHost, c.stream), "cudaMemcpyAsync");
        check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
    }
    check(cudaMemcpyAsync(c.gs_iters, state + 1, sizeof(int), cudaMemcpyDeviceToDevice, c.stream), "cudaMemcpyAsync");
    if (iters_slot) *iters_slot = (double)host_state[1];  // pinned alias: host and device address are the same memory
}

void launch_inverse_matrix(int N, double L, double *U, const double *F)
{
    const Spacing sp = spacing(N, L);
    const size_t n = (size_t)N * N;
    // scratch: Lo (n*n) | Up (n*n) | Z (n) | P (n ints, padded to n doubles)
    double *buf = scratch_grid(2 * n * n + 2 * n);
    if (!buf) return;
    double *Lo = buf, *Up = buf + n * n, *Z = Up + n * n;
    int *P = (int *)(Z + n);
    const int threads = n >= 1024 ? 1024 : (int)((n + 31) / 32 * 32);
    MG_LAUNCH(k_inverse_matrix, 1, threads, 0, N, -4.0 / sp.h2, 1.0 / sp.h2, Lo, Up, P, Z, U, F);
}

}  // namespace mg
