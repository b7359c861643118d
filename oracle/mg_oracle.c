/*
 * mg_oracle.c -- CPU restatement of the reference grid operators (fp64).
 *
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h).  Each function cites the
 * reference lines it restates (file = /root/reference/src/MG_solver_CPU.cpp).
 *
 * Bit-level contract: the reference is built with no -O flag (src/Makefile:8),
 * i.e. every fp operation is an individually rounded IEEE double operation in
 * source order, pow(dx,2) is a real libm call, and there is no FMA.  This file
 * is compiled with -O2 -ffp-contract=off -fno-builtin (oracle/Makefile) and
 * spells out the association of every sum so that the same rounded operations
 * happen in the same order.  Grids are N x N doubles, index = fast + N*slow,
 * boundary included.
 */
#include "mg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int g_last_gs_iterations = 0;

int orc_last_gs_iterations(void) { return g_last_gs_iterations; }

/* pow(x, 2) exactly as glibc evaluates it: the exponent is hidden from the
 * optimiser so the call can never become x*x (they differ by 1 ulp for some
 * grid spacings). */
static double libm_square(double x)
{
    volatile double two = 2.0;
    return pow(x, two);
}

static size_t cells(int N) { return (size_t)N * (size_t)N; }

/* ------------------------------------------------------------------ problem */

/* :497-523 -- homogeneous Dirichlet problem: the whole array becomes 0. */
void orc_getBoundary(int N, double L, double *F, double min_x, double min_y)
{
    (void)L; (void)min_x; (void)min_y;
    memset(F, 0, cells(N) * sizeof(double));
}

/* :468-493 -- interior source term, x = ix*h + min_x, y = iy*h + min_y. */
void orc_getSource(int N, double L, double *F, double min_x, double min_y)
{
    const double h = L / (double)(N - 1);
    orc_getBoundary(N, L, F, min_x, min_y);
#pragma omp parallel for
    for (int iy = 1; iy < N - 1; ++iy) {
        const double y = (double)iy * h + min_y;
        for (int ix = 1; ix < N - 1; ++ix) {
            const double x = (double)ix * h + min_x;
            /* :488  2.0*x*(y-1)*(y - 2.0*x + x*y + 2.0)*exp(x-y), left to right */
            const double poly = ((y - 2.0 * x) + x * y) + 2.0;
            F[(size_t)ix + (size_t)N * iy] = (((2.0 * x) * (y - 1)) * poly) * exp(x - y);
        }
    }
}

/* :525-548 -- analytic solution used for the final error report. */
void orc_getAnalytic(int N, double L, double *U, double min_x, double min_y)
{
    const double h = L / (double)(N - 1);
    orc_getBoundary(N, L, U, min_x, min_y);
#pragma omp parallel for
    for (int iy = 1; iy < N - 1; ++iy) {
        const double y = (double)iy * h + min_y;
        for (int ix = 1; ix < N - 1; ++ix) {
            const double x = (double)ix * h + min_x;
            /* :544  exp(x-y)*x*(1.0-x)*y*(1.0-y) */
            U[(size_t)ix + (size_t)N * iy] = (((exp(x - y) * x) * (1.0 - x)) * y) * (1.0 - y);
        }
    }
}

/* ---------------------------------------------------------------- operators */

/* the five-point sum in the reference's order: ((up + down) + right) + left,
 * "up" = slow index + 1 (:560, :590, :611). */
static inline double nb_sum(const double *g, size_t c, int N)
{
    return ((g[c + N] + g[c - N]) + g[c + 1]) + g[c - 1];
}

/* :554-564 -- D = (1/h^2) * (sum4 - 4U) - F in the interior, 0 on the edge. */
void orc_getResidual(int N, double L, double *U, double *F, double *D)
{
    const double dx = L / (double)(N - 1);
    const double inv_h2 = 1.0 / libm_square(dx);
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) {
            const size_t c = (size_t)i * N + j;
            if (i == 0 || j == 0 || i == N - 1 || j == N - 1)
                D[c] = 0.0;
            else
                D[c] = inv_h2 * (nb_sum(U, c, N) - 4 * U[c]) - F[c];
        }
    }
}

/* :566-571 */
void orc_doGridAddition(int N, double *U1, double *U2)
{
    const size_t n = cells(N);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) U1[i] = U1[i] + U2[i];
}

/* :573-625 -- "Gauss-Seidel" that is really Jacobi: both half sweeps read the
 * snapshot U_old (:581-599), so one step is
 *     U <- U_old + 0.25*((sum4(U_old) - 4*U_old) - pow(dx,2)*F)
 * for every interior point.  The error (:607-622) adds the same red-parity sum
 * twice: error = (S + S)/N/N with S = sum over interior (i+j) even of
 * |inv_h2*(sum4(U) - 4U) - F|, accumulated row-major (== reference, 1 thread). */
void orc_doSmoothing(int N, double L, double *U, double *F, int step, double *error)
{
    const double dx = L / (double)(N - 1);
    const double h2 = libm_square(dx);
    const double inv_h2 = 1.0 / libm_square(dx);
    double *snap = (double *)malloc(cells(N) * sizeof(double));

    for (int s = 0; s < step; ++s) {
        memcpy(snap, U, cells(N) * sizeof(double));
#pragma omp parallel for
        for (int i = 1; i < N - 1; ++i) {
            for (int j = 1; j < N - 1; ++j) {
                const size_t c = (size_t)i * N + j;
                const double bracket = (nb_sum(snap, c, N) - 4 * snap[c]) - h2 * F[c];
                U[c] = snap[c] + 0.25 * bracket;
            }
        }
    }
    free(snap);

    double red = 0.0;
    for (int i = 1; i < N - 1; ++i) {
        for (int j = (i % 2 == 0) ? 2 : 1; j < N - 1; j += 2) {
            const size_t c = (size_t)i * N + j;
            red += fabs(inv_h2 * (nb_sum(U, c, N) - 4 * U[c]) - F[c]);
        }
    }
    double e = red + red;
    e = e / N;
    e = e / N;
    *error = e;
}

/* :640-680 -- bilinear point sampling of the fine grid at the coarse points.
 * The index (floor of a quotient) and the weight (fmod of the un-divided
 * product) are computed exactly as written there; they can disagree for
 * non-nested ladders and that disagreement is part of the contract. */
void orc_doRestriction(int N, double *U_f, int M, double *U_c)
{
    const double h_f = 1.0 / (double)(N - 1);
    const double h_c = 1.0 / (double)(M - 1);
    memset(U_c, 0, cells(M) * sizeof(double));

    /* the x and y maps are the same 1-D function of the coarse index */
    int *lo = (int *)malloc((size_t)M * sizeof(int));
    double *w = (double *)malloc((size_t)M * sizeof(double));
    for (int t = 0; t < M; ++t) {
        const double pos = (double)t * h_c;
        lo[t] = (int)floor(pos / h_f);
        w[t] = fmod(pos, h_f) / h_f;
    }
#pragma omp parallel for
    for (int iy = 1; iy < M - 1; ++iy) {
        const double c = w[iy], d = 1.0 - c;
        for (int ix = 1; ix < M - 1; ++ix) {
            const double a = w[ix], b = 1.0 - a;
            const size_t f = (size_t)lo[ix] + (size_t)lo[iy] * N;
            /* :676  b*d*U00 + a*d*U10 + c*b*U01 + a*c*U11, left to right */
            U_c[(size_t)ix + (size_t)iy * M] =
                (((b * d) * U_f[f] + (a * d) * U_f[f + 1]) + (c * b) * U_f[f + N]) + (a * c) * U_f[f + N + 1];
        }
    }
    free(lo);
    free(w);
}

/* :682-724 -- bilinear prolongation.  The reference scatters from coarse
 * cells with ceil() ranges and patches the last row/column; this restatement
 * is the equivalent gather:
 *   fine index t <= M-2 belongs to the coarse cell q with ceil(q*r) <= t < ceil((q+1)*r),
 *   and sits at coordinate t*f_dx;
 *   fine row   M-1 uses the cell of row M-2 at coordinate (M-1)*f_dx      (:706-712);
 *   fine column M-1 uses the cell of column M-2 at coordinate 1.0         (:701-704),
 *   unless the last cell's range reaches M, in which case the regular pass
 *   rewrites it at coordinate (M-1)*f_dx (last writer wins). */
typedef struct { int cell; double lo_w, hi_w; } prol_map; /* lo_w = (cell+1)*c_dx - f,  hi_w = f - cell*c_dx */

static void prol_build(int N, int M, int is_column, prol_map *map)
{
    const double c_dx = 1.0 / (double)(N - 1), f_dx = 1.0 / (double)(M - 1);
    const double ratio = c_dx / f_dx;
    for (int t = 0; t < M; ++t) map[t].cell = -1;
    for (int q = 0; q < N - 1; ++q) {
        const double q_lo = q * c_dx, q_hi = q_lo + c_dx;
        const double end = ceil((q + 1) * ratio);
        for (int t = (int)ceil(q * ratio); t < end && t < M; ++t) {
            double f = t * f_dx;
            map[t].cell = q; map[t].lo_w = q_hi - f; map[t].hi_w = f - q_lo;
            if (t == M - 2) {
                /* patch of the last line */
                f = is_column ? 1.0 : (M - 1) * f_dx;
                map[M - 1].cell = q; map[M - 1].lo_w = q_hi - f; map[M - 1].hi_w = f - q_lo;
                if (!is_column) break; /* :707 sets k = M-1, ending the row loop */
            }
        }
    }
}

void orc_doProlongation(int N, double *U_c, int M, double *U_f)
{
    const double c_dx = 1.0 / (double)(N - 1);
    prol_map *col = (prol_map *)malloc((size_t)M * sizeof(prol_map));
    prol_map *row = (prol_map *)malloc((size_t)M * sizeof(prol_map));
    prol_build(N, M, 1, col);
    prol_build(N, M, 0, row);
#pragma omp parallel for
    for (int k = 0; k < M; ++k) {
        if (row[k].cell < 0) continue; /* never written by the reference either */
        const double *lo_row = U_c + (size_t)row[k].cell * N;
        const double *hi_row = lo_row + N;
        for (int l = 0; l < M; ++l) {
            const int j = col[l].cell;
            if (j < 0) continue;
            /* :700  ((c1*(c2x-fx) + c2*(fx-c1x))*(c3y-fy) + (c3*(c4x-fx) + c4*(fx-c3x))*(fy-c1y))/c_dx/c_dx */
            const double bottom = lo_row[j] * col[l].lo_w + lo_row[j + 1] * col[l].hi_w;
            const double top = hi_row[j] * col[l].lo_w + hi_row[j + 1] * col[l].hi_w;
            U_f[(size_t)k * M + l] = ((bottom * row[k].lo_w + top * row[k].hi_w) / c_dx) / c_dx;
        }
    }
    free(col);
    free(row);
}

/* ------------------------------------------------------------- exact solver */

/* :952-1066 -- red-black Gauss-Seidel, in place, from U = 0, until the mean
 * absolute interior residual (divided by (N-2)^2, :1059) is <= target.
 * Red = (ix+iy) even (the ieven table :972-980 enumerates exactly those). */
static void gauss_seidel(int N, double L, double *U, double *F, double target)
{
    const double h = L / (double)(N - 1);
    const double h2 = libm_square(h);
    double err = target + 1.0;
    double *R = (double *)malloc(cells(N) * sizeof(double));
    int iters = 0;

    memset(U, 0, cells(N) * sizeof(double));
    while (err > target) {
        for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for
            for (int iy = 1; iy < N - 1; ++iy) {
                for (int ix = 1 + ((iy + 1 + colour) & 1); ix < N - 1; ix += 2) {
                    const size_t c = (size_t)ix + (size_t)iy * N;
                    /* :1020  0.25*(U[l] + U[r] + U[t] + U[b] - pow(h,2)*F) */
                    U[c] = 0.25 * ((((U[c - 1] + U[c + 1]) + U[c + N]) + U[c - N]) - h2 * F[c]);
                }
            }
        }
        ++iters;
        orc_getResidual(N, L, U, F, R);
        err = 0.0;
        for (int j = 1; j < N - 1; ++j)
            for (int i = 1; i < N - 1; ++i) err = err + fabs(R[(size_t)i + (size_t)N * j]);
        err = err / (double)((N - 2) * (N - 2));
    }
    g_last_gs_iterations = iters;
    free(R);
}

/* :758-950 -- dense LU of the N^2 x N^2 Laplacian (unit-diagonal upper factor),
 * forward/back substitution.  Only feasible for tiny N. */
static void inverse_matrix(int N, double Length, double *X, double *F)
{
    const double h = Length / (double)(N - 1);
    const double h2 = libm_square(h);
    const int n = N * N;
    const size_t nn = (size_t)n * n;
    double *A = (double *)calloc(nn, sizeof(double));
    double *Lo = (double *)calloc(nn, sizeof(double));
    double *Up = (double *)calloc(nn, sizeof(double));
    double *Z = (double *)calloc((size_t)n, sizeof(double));
    int *P = (int *)malloc((size_t)n * sizeof(int));
#define AT(m, r, c) m[(size_t)(r) * n + (c)]

    for (int r = 0; r < n; ++r) { AT(Up, r, r) = 1.0; P[r] = r; }
    memset(X, 0, (size_t)n * sizeof(double));

    /* :807-832 -- one matrix row per grid point, identity rows on the boundary */
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
            const int p = i + N * j;
            if (i == 0 || i == N - 1 || j == 0 || j == N - 1) {
                AT(A, p, p) = 1.0;
            } else {
                AT(A, p, p) = -4.0 / h2;
                AT(A, p, p - 1) = 1.0 / h2;
                AT(A, p, p + 1) = 1.0 / h2;
                AT(A, p, p + N) = 1.0 / h2;
                AT(A, p, p - N) = 1.0 / h2;
            }
        }

    /* :842-896 -- Crout-style sweep with the reference's restart-on-zero-pivot */
    int restart, checked = 0, swap_with = 1;
    do {
        restart = 0;
        for (int k = 0; k < n && !restart; ++k) {
            for (int i = k; i < n; ++i) {
                double acc = 0.0;
                for (int j = 0; j < k; ++j) acc = acc + AT(Lo, i, j) * AT(Up, j, k);
                AT(Lo, i, k) = AT(A, P[i], k) - acc;
                if (i == k) {
                    if (AT(Lo, i, k) == 0.0) {
                        if (swap_with >= n) {
                            printf("Having Zero Pivote ! det(A) = 0\n");
                        } else {
                            P[i] = swap_with; P[swap_with] = i; restart = 1;
                        }
                        ++swap_with;
                        break;
                    }
                    ++checked; swap_with = checked + 1;
                }
            }
            if (restart) break;
            for (int j = k; j < n; ++j) {
                double acc = 0.0;
                for (int i = 0; i < k; ++i) acc = acc + AT(Lo, k, i) * AT(Up, i, j);
                AT(Up, k, j) = (1.0 / AT(Lo, k, k)) * (AT(A, P[k], j) - acc);
            }
        }
    } while (restart);

    /* :916-931 */
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int k = 0; k < i; ++k) acc = acc + AT(Lo, i, k) * Z[k];
        Z[i] = (1.0 / AT(Lo, i, i)) * (F[i] - acc);
    }
    for (int i = n - 1; i >= 0; --i) {
        double acc = 0.0;
        for (int k = i + 1; k < n; ++k) acc = acc + AT(Up, i, k) * X[k];
        X[i] = Z[i] - acc;
    }
#undef AT
    free(A); free(Lo); free(Up); free(Z); free(P);
}

/* :627-638 */
void orc_doExactSolver(int N, double L, double *U, double *F, double target_error, int option)
{
    if (option == 0) inverse_matrix(N, L, U, F);
    if (option == 1) gauss_seidel(N, L, U, F, target_error);
}

void orc_default_ops(mg_ops *ops)
{
    ops->getSource = orc_getSource;
    ops->getAnalytic = orc_getAnalytic;
    ops->getResidual = orc_getResidual;
    ops->doGridAddition = orc_doGridAddition;
    ops->doSmoothing = orc_doSmoothing;
    ops->doExactSolver = orc_doExactSolver;
    ops->doRestriction = orc_doRestriction;
    ops->doProlongation = orc_doProlongation;
}

void orc_free(void *p) { free(p); }
