/*
 * mg_oracle.h -- CPU restatement ("oracle") of the reference multigrid hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under multigrid_poisson_solver_b200/ may
 * include, link or dlopen this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or the timed CPU baseline.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_ref.py)
 * bit-for-bit against the unmodified reference operators compiled from
 * /root/reference/src into oracle/_ref/libmgref.so, and against the golden
 * per-node values captured from the unmodified reference in SURVEY.md 8(c)
 * (tests/golden/).
 *
 * All citations are file:line in /root/reference/src/MG_solver_CPU.cpp unless
 * another file is named.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the eight operators + getAnalytic (argument lists = reference prototypes :16-34) */
void orc_getSource(int N, double L, double *F, double min_x, double min_y);
void orc_getBoundary(int N, double L, double *F, double min_x, double min_y);
void orc_getAnalytic(int N, double L, double *U, double min_x, double min_y);
void orc_getResidual(int N, double L, double *U, double *F, double *D);
void orc_doGridAddition(int N, double *U1, double *U2);
void orc_doSmoothing(int N, double L, double *U, double *F, int step, double *error);
void orc_doExactSolver(int N, double L, double *U, double *F, double target_error, int option);
void orc_doRestriction(int N, double *U_f, int M, double *U_c);
void orc_doProlongation(int N, double *U_c, int M, double *U_f);

/* number of Gauss-Seidel iterations taken by the last orc_doExactSolver(option 1) call */
int orc_last_gs_iterations(void);

/* ---- operator table so the cycle driver below can run either the oracle's
 *      operators or the unmodified reference's (oracle/_ref/libmgref.so). */
typedef struct mg_ops {
    void (*getSource)(int, double, double *, double, double);
    void (*getAnalytic)(int, double, double *, double, double);
    void (*getResidual)(int, double, double *, double *, double *);
    void (*doGridAddition)(int, double *, double *);
    void (*doSmoothing)(int, double, double *, double *, int, double *);
    void (*doExactSolver)(int, double, double *, double *, double, int);
    void (*doRestriction)(int, double *, int, double *);
    void (*doProlongation)(int, double *, int, double *);
} mg_ops;

void orc_default_ops(mg_ops *ops);

/* ---- one record per executed node of the Cycle.txt stream */
typedef struct orc_trace_rec {
    int node;      /* -1 pre-smooth+restrict, 0 exact solve, 1 prolong+post-smooth */
    int N;         /* grid size the node worked on (for node 1: the finer grid) */
    int steps;     /* smoothing sweeps actually done (trigger mode: counted) ; node 0: GS iterations or -1 */
    double err;    /* *ptrError after the node's smoothing; node 0: 0 */
    double sumU;   /* sum of the level's U after the node (row-major sequential) */
    double maxabsU;
} orc_trace_rec;

/* optional snapshot hook: called after every node with the level's U */
typedef void (*orc_snap_fn)(void *ctx, int rec_index, int node, int N, const double *U);

typedef struct orc_cycle_result {
    int n_recs;          /* records written (<= max_recs) */
    int N;               /* size of the final (first-node) grid */
    double mg_error;     /* mean |analytic - U| as printed by the reference (:441-445) */
    double time_ms;      /* wall time around the node loop (:156,:429-430) */
    double sumU, maxabsU;/* fingerprint of final U */
} orc_cycle_result;

/* Runs a Cycle.txt the way main() does (:36-462, linkedlist.cpp).  ops==NULL -> oracle ops.
 * If U_out != NULL it receives a malloc'd copy of the final U (caller frees with orc_free).
 * Returns 0 on success, non-zero on parse/usage errors. */
int orc_run_cycle(const char *cycle_path, const mg_ops *ops, int n_threads,
                  orc_trace_rec *recs, int max_recs, orc_snap_fn snap, void *snap_ctx,
                  double **U_out, orc_cycle_result *res);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
