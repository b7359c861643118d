/*
 * mg_oracle_cycle.c -- CPU restatement of the reference cycle driver
 * (main(), /root/reference/src/MG_solver_CPU.cpp:36-462, and the level stack
 * /root/reference/src/linkedlist.cpp:7-124).
 *
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h).  The operators are called
 * through an mg_ops table so the same driver can be run with the oracle's
 * operators or with the unmodified reference operators from
 * oracle/_ref/libmgref.so; it records one orc_trace_rec per executed node with
 * full-precision errors (the reference prints %lf only).
 *
 * Deliberate deviation: a node stream that does not end with the code 2 stops
 * at end of file (the reference re-executes the last node, :158-160, and then
 * usually dereferences a null prevNode).
 */
#include "mg_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TRIGGER_SLOPE 0.01 /* :99 */

typedef struct level {
    int N;
    double *U, *F, *D;
    int step;
    double smoothing_error;
} level;

typedef struct stack {
    level *lv;
    int depth, cap;
    int init; /* linkedlist.h:41-44: 1 until the stack has returned to a single node once */
} stack;

static size_t cells(int N) { return (size_t)N * (size_t)N; }

static void push_level(stack *s, int N) /* linkedlist.cpp:7-44: three uninitialised N*N arrays */
{
    if (s->depth == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 16;
        s->lv = (level *)realloc(s->lv, (size_t)s->cap * sizeof(level));
    }
    level *l = &s->lv[s->depth++];
    l->N = N;
    l->U = (double *)malloc(cells(N) * sizeof(double));
    l->F = (double *)malloc(cells(N) * sizeof(double));
    l->D = (double *)malloc(cells(N) * sizeof(double));
    l->step = 0;
    l->smoothing_error = 0;
}

static void pop_level(stack *s) /* linkedlist.cpp:46-69 */
{
    level *l = &s->lv[--s->depth];
    free(l->U); free(l->F); free(l->D);
    if (s->depth == 1) s->init = 0;
}

static level *top(stack *s) { return &s->lv[s->depth - 1]; }

static void fingerprint(const double *U, int N, double *sum, double *maxabs)
{
    double a = 0.0, m = 0.0;
    const size_t n = cells(N);
    for (size_t i = 0; i < n; ++i) { a += U[i]; if (fabs(U[i]) > m) m = fabs(U[i]); }
    *sum = a; *maxabs = m;
}

typedef struct recorder {
    orc_trace_rec *recs; int max, n;
    orc_snap_fn snap; void *ctx;
    double seconds;   /* time spent recording: NOT part of the reference's timer span, subtracted from it */
} recorder;

static void record(recorder *r, int node, int N, int steps, double err, const double *U)
{
    const double t_in = omp_get_wtime();
    if (r->recs && r->n < r->max) {
        orc_trace_rec *t = &r->recs[r->n];
        t->node = node; t->N = N; t->steps = steps; t->err = err;
        fingerprint(U, N, &t->sumU, &t->maxabsU);
    }
    if (r->snap) r->snap(r->ctx, r->n, node, N, U);
    r->n++;
    r->seconds += omp_get_wtime() - t_in;
}

/* Smoothing part shared by the -1 node (:194-269) and the 1 node (:376-422).
 * step == -1: sweep one at a time until two successive errors differ by <= TRIGGER
 * (minimum two sweeps); any other non-zero step goes to doSmoothing as it is (:244-259,
 * :410-421 -- a negative count sweeps zero times and only evaluates the error).
 * Returns the sweep count the reference prints. */
static int smooth_level(const mg_ops *ops, level *l, double L, int step)
{
    if (step != -1) {
        ops->doSmoothing(l->N, L, l->U, l->F, step, &l->smoothing_error);
        l->step = step;
        return step;
    }
    double slope = TRIGGER_SLOPE + 1.0, previous = 0.0;
    l->step = 0;
    while (slope > TRIGGER_SLOPE) {
        ops->doSmoothing(l->N, L, l->U, l->F, 1, &l->smoothing_error);
        l->step += 1;
        if (l->step > 1) slope = fabs(l->smoothing_error - previous);
        previous = l->smoothing_error;
    }
    return l->step;
}

int orc_run_cycle(const char *cycle_path, const mg_ops *ops_in, int n_threads,
                  orc_trace_rec *recs, int max_recs, orc_snap_fn snap, void *snap_ctx,
                  double **U_out, orc_cycle_result *res)
{
    mg_ops own;
    if (!ops_in) { orc_default_ops(&own); ops_in = &own; }
    const mg_ops *ops = ops_in;
    if (n_threads > 0) omp_set_num_threads(n_threads);

    FILE *fp = fopen(cycle_path, "r");
    if (!fp) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", cycle_path); return 1; }

    double L, min_x, min_y;
    int con_step, con_N, N_max, N_min;
    if (fscanf(fp, "%lf %lf %lf %d %d %d %d", &L, &min_x, &min_y, &con_step, &con_N, &N_max, &N_min) != 7) {
        fclose(fp); return 2;
    }

    /* :111-146 -- ladder of grid sizes */
    int *ladder = NULL, ladder_len = 0, pos = 0;
    if (con_N == 1) {
        for (int n = N_max; n >= N_min; n /= 2) ++ladder_len;
        ladder = (int *)malloc((size_t)(ladder_len + 1) * sizeof(int));
        for (int i = 0, n = N_max; i < ladder_len; ++i, n /= 2) ladder[i] = n;
    } else if (con_N == 2) {
        ladder_len = N_max - N_min + 1;
        ladder = (int *)malloc((size_t)(ladder_len + 1) * sizeof(int));
        for (int i = 0; i < ladder_len; ++i) ladder[i] = N_max - i;
    }

    stack st = {0};
    st.init = 1;
    recorder rec = {recs, max_recs, 0, snap, snap_ctx, 0.0};
    int rc = 0;

    push_level(&st, N_max);                                           /* :149 */
    ops->getSource(N_max, L, top(&st)->F, min_x, min_y);              /* :153, outside the timer */

    const double t0 = omp_get_wtime();                                /* :156 */
    for (;;) {
        int node;
        if (fscanf(fp, "%d", &node) != 1) break;                      /* deviation, see header */
        if (node == 2) break;                                         /* :162 */

        if (node == -1) {                                             /* :169-301 */
            int step, next_N;
            if (con_step == 0) { if (fscanf(fp, "%d", &step) != 1) { rc = 3; break; } }
            else step = con_step;
            if (con_N == 0) { if (fscanf(fp, "%d", &next_N) != 1) { rc = 3; break; } }
            else {
                if (pos + 1 >= ladder_len) { rc = 4; break; }         /* reference reads past N_array */
                next_N = ladder[++pos];
            }
            level *l = top(&st);
            if (step == 0) continue;                                  /* :241-243,:296-299 FMG placeholder */

            if (!(st.init == 0 && st.depth == 1))                     /* :209-214 / :252-257 */
                memset(l->U, 0, cells(l->N) * sizeof(double));
            const int done = smooth_level(ops, l, L, step);
            ops->getResidual(l->N, L, l->U, l->F, l->D);              /* :239 / :268 */
            record(&rec, -1, l->N, done, l->smoothing_error, l->U);

            const long long n = (long long)cells(l->N);               /* :277-280 (omp parallel for there too) */
            double *Dn = l->D;
#pragma omp parallel for
            for (long long i = 0; i < n; ++i) Dn[i] = -Dn[i];
            const int fine_N = l->N;
            double *fine_D = l->D;
            push_level(&st, next_N);                                  /* :283 (may move st.lv) */
            ops->doRestriction(fine_N, fine_D, next_N, top(&st)->F);  /* :287 */
        } else if (node == 0) {                                       /* :305-325 */
            double target; int option;
            if (fscanf(fp, "%lf %d", &target, &option) != 2) { rc = 3; break; }
            level *l = top(&st);
            ops->doExactSolver(l->N, L, l->U, l->F, target, option);
            record(&rec, 0, l->N, -1, 0.0, l->U);
        } else if (node == 1) {                                       /* :329-424 */
            int step;
            if (con_step == 0) { if (fscanf(fp, "%d", &step) != 1) { rc = 3; break; } }
            else step = con_step;
            if (con_N != 0) --pos;
            if (st.depth < 2) { rc = 5; break; }                      /* reference: null prevNode */
            level *coarse = top(&st);
            const int fine_N = st.lv[st.depth - 2].N;                 /* :350 */
            double *tmp = (double *)malloc(cells(fine_N) * sizeof(double));
            ops->doProlongation(coarse->N, coarse->U, fine_N, tmp);   /* :354 */
            pop_level(&st);                                           /* :363 */
            level *l = top(&st);
            ops->doGridAddition(l->N, l->U, tmp);                     /* :368 */
            free(tmp);
            int done = 0;
            if (step != 0) done = smooth_level(ops, l, L, step);      /* :376-422 */
            record(&rec, 1, l->N, done, l->smoothing_error, l->U);
        } else {
            rc = 6; break;
        }
    }
    const double t1 = omp_get_wtime();                                /* :429 */
    fclose(fp);

    if (rc == 0 && res) {
        level *l = top(&st);
        double *ana = (double *)malloc(cells(l->N) * sizeof(double));
        ops->getAnalytic(l->N, L, ana, min_x, min_y);                 /* :439 */
        double acc = 0.0;
        const size_t n = cells(l->N);
        for (size_t i = 0; i < n; ++i) acc = acc + fabs(ana[i] - l->U[i]);   /* :442-444 */
        free(ana);
        res->mg_error = acc / (double)(l->N * l->N);
        res->time_ms = 1000.0 * (t1 - t0 - rec.seconds);   /* the reference's span :156-:429, without our recording */
        res->N = l->N;
        res->n_recs = rec.n < max_recs ? rec.n : max_recs;
        fingerprint(l->U, l->N, &res->sumU, &res->maxabsU);
        if (U_out) {
            *U_out = (double *)malloc(n * sizeof(double));
            memcpy(*U_out, l->U, n * sizeof(double));
        }
    }
    while (st.depth > 0) { level *l = &st.lv[--st.depth]; free(l->U); free(l->F); free(l->D); }
    free(st.lv);
    free(ladder);
    return rc;
}
