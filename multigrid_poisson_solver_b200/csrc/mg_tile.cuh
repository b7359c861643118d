// mg_tile.cuh -- the fused passes of a cycle node for SMALL grids: one CTA per 32 x 32 tile,
// all levels of the pass in shared memory.
//
// The streaming kernel (mg_stream.cuh) needs even N (16-byte aligned rows).  Here the same pass
// -- [U = 0 | U | U_f + P(U_c)] + S Jacobi sweeps + error + residual + negate + restrict -- is
// tiled in two dimensions with no alignment requirement: a CTA loads its tile with a halo of
// S + 2 points, runs the sweeps sweep by sweep with a barrier in between (the valid region
// shrinks by one ring per sweep; the halo is recomputed, not exchanged), and emits the owned
// 32 x 32 points, its share of the restricted grid and one error partial.  It serves the ODD
// level sizes up to 1024 that would otherwise fall back to one kernel per operator (e.g. 181 in
// the ladder 23168 / 2^k: V-cycle 181 -> 8 0.31 -> 0.26 ms).  On small EVEN grids it ties with the
// streaming kernel (about 14 us per node at N = 256 either way: launch latency and the pass's
// dependency chain, not throughput), so those keep the streaming kernel; mgSetTileMaxN routes
// every size through it for tests.  Whole grids only (single GPU / agglomerated levels).
//
// Per-point arithmetic = the functions of mg_device.cuh, called with the operand order of the
// one-kernel-per-operator path (mg_kernels.cu), so every bit of U and F_c matches the reference:
//   sweep      MG_solver_CPU.cpp:578-601   jacobi_at(u, sum4(above, below, right, left), h2*F), boundary carried over
//   error      :607-622                    2 * sum over (row+col) even interior points of |residual| / N / N
//   residual   :554-564, negate :277-280   0 on the boundary, then D = -D
//   restrict   :661-676                    through the inverse floor map f2c (coarse boundary = 0)
//   prolong    :688-724 + :569             U_f + prolong_at(...) on ALL points (tables of mg_kernels.cu)
#pragma once
#include "mg_device.cuh"
#include "mg_stream.cuh"

namespace mg {

constexpr int TILE = 32;           // owned points per CTA and dimension
constexpr int TILE_THREADS = 256;

template <int S, int IN, bool ERR, bool RES>
__global__ void __launch_bounds__(TILE_THREADS) k_tile(const StreamParams p)
{
    constexpr bool NEED_R = ERR || RES;
    constexpr int G = S + (RES ? 2 : NEED_R ? 1 : 0);     // halo of level 0
    constexpr int E = TILE + 2 * G;                        // extent of the shared arrays
    constexpr int NT = TILE_THREADS;
    __shared__ double bufA[E * E], bufB[E * E], Fs[E * E];
    __shared__ double red_smem[32];
    __shared__ bool is_last;

    const int N = p.N, tid = threadIdx.x;
    const int tiles_x = (N + TILE - 1) / TILE;
    const int r0 = (blockIdx.x / tiles_x) * TILE, c0 = (blockIdx.x % tiles_x) * TILE;
    const int gr = r0 - G, gc = c0 - G;                    // grid coordinates of shared element (0, 0)
    const double *__restrict__ Fg = p.F;
    const double *__restrict__ Ug = p.Uin;

    // ---- level 0 and F on the whole extent (points outside the grid stay 0 and are never read)
    for (int idx = tid; idx < E * E; idx += NT) {
        const int a = idx / E, b = idx - a * E, i = gr + a, j = gc + b;
        double u = 0.0, f = 0.0;
        if (i >= 0 && i < N && j >= 0 && j < N) {
            const size_t o = (size_t)i * N + j;
            f = Fg[o];
            if (IN == IN_LOAD) u = Ug[o];
            if (IN == IN_PROLONG) {
                const int q = p.col_cell[j];
                const double *lo_row = p.Uc + (size_t)p.row_cell[i] * p.Nc + q;
                const double v = prolong_at(lo_row[0], lo_row[1], lo_row[p.Nc], lo_row[p.Nc + 1], p.col_w[j], p.row_w[i], p.c_dx);
                u = __dadd_rn(Ug[o], v);
            }
        }
        bufA[idx] = u;
        Fs[idx] = f;
    }
    __syncthreads();

    // ---- S sweeps; level k+1 is valid on the extent minus k+1 rings
    double *cur = bufA, *nxt = bufB;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        const int lo = k + 1, w = E - 2 * (k + 1);
        for (int idx = tid; idx < w * w; idx += NT) {
            const int a = lo + idx / w, b = lo + idx % w, i = gr + a, j = gc + b;
            if (i < 0 || i >= N || j < 0 || j >= N) continue;
            const int s = a * E + b;
            double out = cur[s];
            if (i > 0 && i < N - 1 && j > 0 && j < N - 1)
                out = jacobi_at(out, sum4(cur[s + E], cur[s - E], cur[s + 1], cur[s - 1]), __dmul_rn(p.h2, Fs[s]));
            nxt[s] = out;
        }
        __syncthreads();
        double *t = cur; cur = nxt; nxt = t;
    }

    // ---- owned points of level S
    if (p.Uout) {
        for (int idx = tid; idx < TILE * TILE; idx += NT) {
            const int a = idx / TILE, b = idx % TILE, i = r0 + a, j = c0 + b;
            if (i < N && j < N) p.Uout[(size_t)i * N + j] = cur[(a + G) * E + b + G];
        }
    }

    if (NEED_R) {
        // ---- residual of level S on the owned points (one more row and column for the restriction pairs)
        constexpr int RW = TILE + (RES ? 1 : 0);
        double acc = 0.0;
        for (int idx = tid; idx < RW * RW; idx += NT) {
            const int a = idx / RW, b = idx % RW, i = r0 + a, j = c0 + b;
            if (i >= N || j >= N) continue;
            const int s = (a + G) * E + b + G;
            double res = 0.0;
            if (i > 0 && i < N - 1 && j > 0 && j < N - 1)
                res = residual_at(cur[s], sum4(cur[s + E], cur[s - E], cur[s + 1], cur[s - 1]), Fs[s], p.inv_h2);
            if (ERR && a < TILE && b < TILE && ((i + j) & 1) == 0) acc = __dadd_rn(acc, fabs(res));   // boundary adds |0|
            if (RES) nxt[s] = -res;                                                                     // D = -D
        }
        if (RES) {
            __syncthreads();
            for (int idx = tid; idx < TILE * TILE; idx += NT) {
                const int a = idx / TILE, b = idx % TILE, i = r0 + a, j = c0 + b;
                if (i >= N || j >= N) continue;
                const int ci = p.f2c[i], cj = p.f2c[j];
                if (ci < 0 || cj < 0) continue;
                const int s = (a + G) * E + b + G;
                const bool edge = ci == 0 || ci == p.M - 1 || cj == 0 || cj == p.M - 1;
                p.Fc[(size_t)ci * p.M + cj] = edge ? 0.0 : restrict_at(nxt[s], nxt[s + 1], nxt[s + E], nxt[s + E + 1], p.rw[cj], p.rw[ci]);
            }
        }
        if (ERR) {
            // deterministic: fixed per-thread order, fixed tree per CTA, the last CTA folds the partials in order
            const double total = block_sum<NT>(acc, red_smem);
            if (tid == 0) {
                p.partials[blockIdx.x] = total;
                __threadfence();
                is_last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
            }
            __syncthreads();
            if (!is_last) return;
            __threadfence();
            double s = 0.0;
            for (unsigned k = tid; k < gridDim.x; k += NT) s = __dadd_rn(s, __ldcg(&p.partials[k]));
            s = block_sum<NT>(s, red_smem);
            if (tid == 0) {
                double e = __dadd_rn(s, s);                    // sum1 + sum2 over the same parity (:621)
                e = __ddiv_rn(e, (double)N);
                e = __ddiv_rn(e, (double)N);
                if (p.err_dev) *p.err_dev = e;
                if (p.err_slot) *p.err_slot = e;
                *p.counter = 0u;
            }
        }
    }
}

}  // namespace mg
