// mg_tail.cu -- the coarse tail of a cycle in ONE kernel.
//
// Once a cycle is down at a level of at most TAIL_MAX_N points per side, every node it executes
// until it prolongs back above that level (the -1 / 0 / 1 nodes of the coarse sub-cycle: 7 nodes
// in Vcycle.txt's 64..8 tail, 2^k exact solves in a W-cycle) works on grids that fit in shared
// memory together.  Running them as separate launches costs 17-60 us per node (a single warp's
// dependency chain per node); this kernel interprets the whole node sub-stream in
// one CTA with all levels resident in shared memory: one thread per grid point, a CTA barrier per
// sweep, no global traffic except the entry level's F and U.
//
// Arithmetic: the same mg_device.cuh expressions as the big kernels, so grids are bit-identical;
// the error sums are reduced in a different (fixed) order.
//   -1 node  MG_solver_CPU.cpp:246-287   0 node :305-313 (GaussSeidel :952-1066)   1 node :350-416
#include <cmath>
#include <vector>

#include "../../include/mg_abi.h"
#include "mg_device.cuh"
#include "mg_kernels.h"

namespace mg {
namespace {

constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_MAX_OPS = 56;
constexpr int TAIL_MAX_DEPTH = 8;
constexpr double TRIGGER = 0.01;   // MG_solver_CPU.cpp:99

struct TailOp {
    int kind;        // -1, 0, 1
    int step;        // sweeps; -1 = error trigger; (kind 0: unused)
    int zero_init;   // -1 node: U = 0 first
    int lev;         // depth inside the tail the node works on (0 = entry level); kind 1: the FINE level
    double h2, inv_h2;   // spacing constants of that level (host libm)
    double target;   // kind 0: Gauss-Seidel target error
};

struct TailProgram {
    int n_ops, n_levels;
    int N[TAIL_MAX_DEPTH];          // size of each depth
    int off_U[TAIL_MAX_DEPTH];      // shared-memory offsets (doubles)
    int off_F[TAIL_MAX_DEPTH];
    int off_scratch, off_tab;
    double *U_entry;                // global, in/out
    const double *F_entry;          // global, in
    double *out;                    // device alias of pinned slots: out[2*i] = error, out[2*i+1] = sweeps / iterations
    TailOp ops[TAIL_MAX_OPS];
};

// Thread (ty, tx) = (tid / 64, tid % 64) owns the points (i = ty + 16 k, j = tx) of every level
// (levels are at most 64 wide), so no integer division is needed to find a point's row and column.
#define TAIL_FOR_POINTS(N)                                  \
    for (int i = (int)(threadIdx.x >> 6); i < (N); i += 16) \
        if (const int j = (int)(threadIdx.x & 63); j < (N))

__device__ __forceinline__ bool interior(int i, int j, int N) { return i > 0 && i < N - 1 && j > 0 && j < N - 1; }

__device__ __forceinline__ double tail_block_sum(double v, double *red)
{
    return block_sum<TAIL_THREADS>(v, red);   // valid in thread 0
}

// broadcast thread 0's value to the CTA
__device__ __forceinline__ double tail_bcast(double v, double *red)
{
    if (threadIdx.x == 0) red[32] = v;
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
}

// one Jacobi sweep src -> dst (boundary carried over)                    :578-601
__device__ __forceinline__ void tail_sweep(int N, double h2, const double *src, const double *f, double *dst)
{
    TAIL_FOR_POINTS(N) {
        const int c = i * N + j;
        double v = src[c];
        if (interior(i, j, N)) v = jacobi_at(v, sum4(src[c + N], src[c - N], src[c + 1], src[c - 1]), __dmul_rn(h2, f[c]));
        dst[c] = v;
    }
    __syncthreads();
}

// (sum1 + sum2)/N/N over the red interior points                          :607-622
__device__ double tail_error(int N, double inv_h2, const double *u, const double *f, double *red)
{
    double acc = 0.0;
    TAIL_FOR_POINTS(N) {
        const int c = i * N + j;
        if (interior(i, j, N) && ((i + j) & 1) == 0)
            acc = __dadd_rn(acc, fabs(residual_at(u[c], sum4(u[c + N], u[c - N], u[c + 1], u[c - 1]), f[c], inv_h2)));
    }
    double s = tail_block_sum(acc, red);
    if (threadIdx.x == 0) {
        s = __dadd_rn(s, s);
        s = __ddiv_rn(s, (double)N);
        s = __ddiv_rn(s, (double)N);
    }
    return tail_bcast(s, red);
}

// `step` sweeps or the error-trigger loop (:216-230), ping-ponging between u and w; the result is
// copied back into u if it ends in w.  Returns sweeps done, error in *err.
__device__ int tail_smooth(int N, int step, double h2, double inv_h2, double *u, const double *f, double *w, double *red, double *err)
{
    double *cur = u, *other = w;
    int done = 0;
    double e = 0.0;
    if (step > 0) {
        for (int s = 0; s < step; ++s) {
            tail_sweep(N, h2, cur, f, other);
            double *t = cur; cur = other; other = t;
        }
        done = step;
        e = tail_error(N, inv_h2, cur, f, red);
    } else {
        double slope = TRIGGER + 1.0, prev = 0.0;
        while (slope > TRIGGER) {
            tail_sweep(N, h2, cur, f, other);
            double *t = cur; cur = other; other = t;
            e = tail_error(N, inv_h2, cur, f, red);
            ++done;
            if (done > 1) slope = fabs(e - prev);
            prev = e;
        }
    }
    if (cur != u) {
        TAIL_FOR_POINTS(N) u[i * N + j] = cur[i * N + j];
        __syncthreads();
    }
    *err = e;
    return done;
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) k_coarse_tail(const TailProgram P)
{
    extern __shared__ double sm[];
    double *red = sm + P.off_tab;            // [0..32] reduction scratch
    int *itab = (int *)(red + 112);          // 2 x 64 ints   (red[48..111]: Gauss-Seidel warp partials)
    double *dtab = red + 112 + 64;           // 6 x 64 doubles
    double *scratch = sm + P.off_scratch;
    const int tid = threadIdx.x;

    {   // entry level from global memory
        const int n = P.N[0] * P.N[0];
        double *u = sm + P.off_U[0], *f = sm + P.off_F[0];
        for (int c = tid; c < n; c += TAIL_THREADS) { u[c] = P.U_entry[c]; f[c] = P.F_entry[c]; }
        __syncthreads();
    }

    for (int o = 0; o < P.n_ops; ++o) {
        const TailOp op = P.ops[o];
        const int N = P.N[op.lev];
        double *u = sm + P.off_U[op.lev], *f = sm + P.off_F[op.lev];
        double err = 0.0;
        int done = 0;

        if (op.kind == -1) {
            if (op.zero_init) {
                TAIL_FOR_POINTS(N) u[i * N + j] = 0.0;
                __syncthreads();
            }
            done = tail_smooth(N, op.step, op.h2, op.inv_h2, u, f, scratch, red, &err);
            // D = -(getResidual)                                             :268, :277-280
            TAIL_FOR_POINTS(N) {
                const int c = i * N + j;
                double r = 0.0;
                if (interior(i, j, N)) r = residual_at(u[c], sum4(u[c + N], u[c - N], u[c + 1], u[c - 1]), f[c], op.inv_h2);
                scratch[c] = -r;
            }
            // doRestriction(N, D, M, F_next)                                  :640-680
            const int M = P.N[op.lev + 1];
            double *fc = sm + P.off_F[op.lev + 1];
            if (tid < M) restrict_table_entry(tid, N, M, itab[tid], dtab[tid]);
            __syncthreads();
            TAIL_FOR_POINTS(M) {
                double v = 0.0;
                if (interior(i, j, M)) {
                    const int fidx = itab[j] + itab[i] * N;
                    v = restrict_at(scratch[fidx], scratch[fidx + 1], scratch[fidx + N], scratch[fidx + N + 1], dtab[j], dtab[i]);
                }
                fc[i * M + j] = v;
            }
            __syncthreads();
        } else if (op.kind == 0) {
            // GaussSeidel                                                     :952-1066
            TAIL_FOR_POINTS(N) u[i * N + j] = 0.0;
            __syncthreads();
            // few warps with named barriers: a 1024-thread barrier costs more than a whole iteration
            const int warps = N * N <= 128 ? 1 : N * N <= 512 ? 4 : N * N <= 1024 ? 8 : 32;   // <= 4 points per thread
            if (tid < warps * 32) done = gauss_seidel_shared<4>(N, op.h2, op.inv_h2, op.target, u, f, red + 48, warps, 1, 100000000);
            __syncthreads();
        } else {
            // 1 node: U_f += doProlongation(U_c), then smoothing                :350-416
            const int Nc = P.N[op.lev + 1];
            const double *uc = sm + P.off_U[op.lev + 1];
            int *row_cell = itab, *col_cell = itab + 64;
            double2 *row_w = (double2 *)dtab, *col_w = (double2 *)(dtab + 128);
            if (tid < N) prolong_table_entry(tid, Nc, N, row_cell[tid], col_cell[tid], row_w[tid], col_w[tid]);
            __syncthreads();
            const double c_dx = __ddiv_rn(1.0, (double)(Nc - 1));
            TAIL_FOR_POINTS(N) {
                const int c = i * N + j;
                const double *lo_row = uc + row_cell[i] * Nc + col_cell[j];
                const double v = prolong_at(lo_row[0], lo_row[1], lo_row[Nc], lo_row[Nc + 1], col_w[j], row_w[i], c_dx);
                u[c] = __dadd_rn(u[c], v);
            }
            __syncthreads();
            if (op.step != 0) done = tail_smooth(N, op.step, op.h2, op.inv_h2, u, f, scratch, red, &err);
        }
        if (tid == 0) {
            P.out[2 * o] = err;
            P.out[2 * o + 1] = (double)done;
        }
    }
    {
        const int n = P.N[0] * P.N[0];
        const double *u = sm + P.off_U[0];
        for (int c = tid; c < n; c += TAIL_THREADS) P.U_entry[c] = u[c];
    }
    if (tid == 0) __threadfence_system();
}

}  // namespace
}  // namespace mg

using namespace mg;

extern "C" int mgCoarseTailMaxN(void) { return 64; }
extern "C" int mgCoarseTailMaxOps(void) { return TAIL_MAX_OPS; }

// ops: n_ops x {kind, step, zero_init, N, next_N(kind -1) , option} as ints + targets as doubles.
// Returns 0 if the tail ran, > 0 if it cannot (caller falls back to node-by-node execution).
extern "C" int mgCoarseTail(double L, double *U_entry, double *F_entry, int n_ops, const int *kind, const int *step,
                            const int *zero_init, const int *N_of_op, const int *next_N, const double *target,
                            const int *option, double *out_slots)
{
    if (!ensure_ready()) return 10;
    if (n_ops < 1 || n_ops > TAIL_MAX_OPS) return 1;
    TailProgram P{};
    P.n_ops = n_ops;
    // simulate the level stack to assign depths; every depth must keep one size
    int depth = 0;
    P.N[0] = N_of_op[0];
    P.n_levels = 1;
    for (int d = 1; d < TAIL_MAX_DEPTH; ++d) P.N[d] = 0;
    for (int i = 0; i < n_ops; ++i) {
        TailOp &op = P.ops[i];
        op.kind = kind[i];
        op.step = step[i];
        op.zero_init = zero_init[i];
        op.target = target[i];
        if (kind[i] == -1) {
            if (N_of_op[i] != P.N[depth] || depth + 1 >= TAIL_MAX_DEPTH) return 2;
            if (P.N[depth + 1] != 0 && P.N[depth + 1] != next_N[i]) return 2;
            if (next_N[i] < 3 || next_N[i] > P.N[depth] || step[i] == 0) return 2;
            P.N[depth + 1] = next_N[i];
            op.lev = depth;
            ++depth;
            P.n_levels = std::max(P.n_levels, depth + 1);
        } else if (kind[i] == 0) {
            if (option[i] != 1 || N_of_op[i] != P.N[depth]) return 3;   // InverseMatrix stays on the stand-alone path
            op.lev = depth;
        } else if (kind[i] == 1) {
            if (depth < 1 || N_of_op[i] != P.N[depth - 1]) return 4;
            --depth;
            op.lev = depth;
        } else return 5;
        const Spacing sp = spacing(P.N[op.lev], L);
        op.h2 = sp.h2;
        op.inv_h2 = sp.inv_h2;
    }
    if (depth != 0) return 6;                      // the tail must come back to its entry level
    if (P.N[0] > 64 || P.N[0] < 3) return 7;
    int off = 0;
    for (int d = 0; d < P.n_levels; ++d) {
        P.off_U[d] = off; off += P.N[d] * P.N[d];
        P.off_F[d] = off; off += P.N[d] * P.N[d];
    }
    P.off_scratch = off; off += P.N[0] * P.N[0];
    off = (off + 1) & ~1;                          // 16-byte alignment of the double2 tables
    P.off_tab = off; off += 112 + 64 + 6 * 64;
    const size_t smem = (size_t)off * sizeof(double);
    if (smem > 200 * 1024) return 8;
    P.U_entry = U_entry;
    P.F_entry = F_entry;
    P.out = slot_device_ptr(out_slots);
    if (!P.out) return 9;
    static size_t opted = 0;
    if (smem > opted) {
        check(cudaFuncSetAttribute(k_coarse_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(k_coarse_tail)");
        opted = smem;
    }
    k_coarse_tail<<<1, TAIL_THREADS, smem, ctx().stream>>>(P);
    ctx().launches++;
    check(cudaGetLastError(), "k_coarse_tail");
    return 0;
}
