/*
 * mg_abi.h -- C ABI of the B200-native multigrid Poisson hot path (libmgb200.so).
 *
 * This is the drop-in boundary for the reference's grid operators.  The reference
 * has no FFI: its boundary is the set of free functions declared at
 * /root/reference/src/MG_solver_CPU.cpp:16-34 (GPU twins MG_solver_GPU.cu:31-45),
 * called only by main().  The eight operators below keep those names, argument
 * order and meaning; the only change is that every grid pointer is a DEVICE
 * pointer to a contiguous N x N fp64 array (index = ix + N*iy, boundary
 * included).  Scalars returned through pointers (`error`) are HOST pointers and
 * are valid when the call returns.
 *
 * All calls are issued by one host thread and are ordered on one CUDA stream per
 * context.  There is no CPU fallback: every entry point aborts through the error
 * channel (mgLastError) if the CUDA device is missing.
 *
 * Plain C, no C++/torch types in any signature.
 */
#ifndef MG_ABI_H
#define MG_ABI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ lifecycle
 * The reference folds these into malloc/memset/free and cudaSetDevice(0)
 * (MG_solver_GPU.cu:58; linkedlist.cpp:9-11,48-50). */
int mgInit(int device);                 /* 0 on success; creates the context on `device` */
void mgShutdown(void);
int mgLastErrorCode(void);              /* 0 = no error since the last mgClearError */
const char *mgLastError(void);
void mgClearError(void);
void mgSync(void);                      /* wait for everything queued so far */
void *mgStream(void);                   /* the context's cudaStream_t (for event timing by callers) */
int mgKernelLaunches(void);             /* kernels launched by this library since mgInit */
/* Task geometry of a fused pass (host-only, no GPU needed): the row segments {first, past-last},
 * relative to the first owned row, in the order the persistent warps take them.  subset: 0 the
 * whole owned range, 1 / 2 the edge / interior launch of a split slab pass.  Returns the number
 * of segments written to out[2*k], out[2*k+1], or < 0. */
int mgSegmentPlan(int rows, int n_strips, int resident_warps, int lead_rows, int subset, int *out, int max_out);
/* Which whole grids the shared-memory tile kernel serves instead of the streaming kernel.  Default
 * (n < 0): odd sizes up to 1024 -- the streaming kernel needs even N and is as fast on small even
 * grids.  n >= 0: every size up to n (0: never); also MG_TILE_MAX_N.  Returns the old setting (-1 =
 * default).  All paths produce the same bits; the switch exists for tuning and for testing each
 * kernel on every size. */
int mgSetTileMaxN(int n);

double *mgGridAlloc(int N);             /* pooled device array of N*N doubles (uninitialised, like malloc) */
void mgGridFree(double *grid);
void mgGridZero(int N, double *grid);   /* memset(U, 0, ...)           MG_solver_CPU.cpp:213,256 */
void mgGridNegate(int N, double *grid); /* D = -D                      MG_solver_CPU.cpp:277-280 */
void mgGridUpload(int N, double *dev, const double *host);
void mgGridDownload(int N, const double *dev, double *host);
void mgGridCopy(int N, double *dst, const double *src);

/* ------------------------------------------------------- the eight operators
 * Same names / argument lists as the reference prototypes. */
/* MG_solver_CPU.cpp:468-493 */
void getSource(int N, double L, double *F, double min_x, double min_y);
/* MG_solver_CPU.cpp:497-523 */
void getBoundary(int N, double L, double *F, double min_x, double min_y);
/* MG_solver_CPU.cpp:554-564 */
void getResidual(int N, double L, double *U, double *F, double *D);
/* MG_solver_CPU.cpp:573-625 -- `error` is a host pointer */
void doSmoothing(int N, double L, double *U, double *F, int step, double *error);
/* MG_solver_CPU.cpp:640-680 -- N fine, M coarse */
void doRestriction(int N, double *U_f, int M, double *U_c);
/* MG_solver_CPU.cpp:682-724 -- N coarse, M fine */
void doProlongation(int N, double *U_c, int M, double *U_f);
/* MG_solver_CPU.cpp:566-571 */
void doGridAddition(int N, double *U1, double *U2);
/* MG_solver_CPU.cpp:627-638 -- option 0 InverseMatrix (:758-950), 1 GaussSeidel (:952-1066) */
void doExactSolver(int N, double L, double *U, double *F, double target_error, int option);

/* MG_solver_CPU.cpp:525-548 */
void getAnalytic(int N, double L, double *U, double min_x, double min_y);
/* mean |analytic - U| of the final report, MG_solver_CPU.cpp:434-445 (synchronous) */
double mgAnalyticError(int N, double L, const double *U, double min_x, double min_y);
/* mean |A - B| over N^2 points (the final report's reduction with a caller-supplied reference grid; synchronous) */
double mgMeanAbsDiff(int N, const double *A, const double *B);
/* Gauss-Seidel iterations taken by the last doExactSolver(option 1) (synchronous) */
int mgLastExactSolverIterations(void);

/* ----------------------------------------------------------- fused operators
 * What the cycle driver uses where the cycle permits.  Results are identical to
 * the corresponding sequence of the eight operators.
 *
 * `error` arguments of the fused calls are HOST pointers obtained from
 * mgScalarSlot(): pinned, device-visible doubles that the kernel itself writes.
 * The value is valid after mgSync() (no per-call synchronisation). */
double *mgScalarSlot(int index);        /* index in [0, MG_SCALAR_SLOTS) */
#define MG_SCALAR_SLOTS 4096

/* step Jacobi sweeps out of place (U_in untouched, result in U_out) + the
 * reference's smoothing error.  == doSmoothing on a copy.  error_slot == NULL skips the
 * error pass (then step == 1 is exactly one sweep kernel). */
void mgSmooth(int N, double L, const double *U_in, double *F, int step, double *U_out, double *error_slot);

/* the whole -1 node (MG_solver_CPU.cpp:246-287) without materialising D:
 *   if zero_init: U = 0;  U <- step sweeps;  *error;  F_c <- doRestriction(N, -(getResidual(U,F)), M)
 * U_work is scratch of N*N doubles (ping-pong partner); on return the smoothed
 * grid is in the buffer the call returns (either U or U_work). */
double *mgDownLeg(int N, double L, double *U, double *U_work, double *F, int step, int zero_init,
                  int M, double *F_c, double *error_slot);

/* doExactSolver that also reports its Gauss-Seidel iteration count (as a double, -1 for
 * option 0) into a scalar slot, without synchronising. */
void mgExactSolve(int N, double L, double *U, double *F, double target_error, int option, double *iters_slot);

/* the whole 1 node (MG_solver_CPU.cpp:350-416) without materialising tempU:
 *   U_f <- U_f + doProlongation(Nc, U_c, N);  U_f <- step sweeps;  *error
 * returns the buffer holding the result (U_f or U_work). */
double *mgUpLeg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, double *F, int step,
                double *error_slot);

/* The error-trigger forms of the two nodes (con_step = -1, MG_solver_CPU.cpp:194-240 and :376-408): sweep, evaluate the
 * smoothing error, stop when two successive errors differ by <= 0.01 (at least two sweeps).  Two sweeps per launch, the
 * launch reporting the error after each of them; every launch of the -1 node also restricts its result, so a node that
 * stops at the minimum of two sweeps is one launch and one synchronisation.  If the loop ends after an odd number of
 * sweeps the last launch is repeated with a single sweep from the same input.  *steps = sweeps done, *error = the last
 * smoothing error (host pointers, valid on return).  Returns the buffer holding the result (U / U_f or U_work), or NULL if
 * the size is not served by the streaming kernel (odd N, ...): the caller then sweeps one at a time with mgSmooth. */
double *mgDownLegTrigger(int N, double L, double *U, double *U_work, double *F, int zero_init, int M, double *F_c,
                         int *steps, double *error);
double *mgUpLegTrigger(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, double *F, int *steps,
                       double *error);

/* The coarse tail of a cycle in one kernel: a node sub-stream that starts at a level of at most
 * mgCoarseTailMaxN() points per side and returns to it (parallel arrays, one entry per node:
 * kind -1/0/1, step (-1 = error trigger), zero_init, the size the node works on, next_N for -1
 * nodes, target/option for 0 nodes).  All levels live in shared memory.  out_slots (from
 * mgScalarSlot) receives per node {smoothing error, sweeps or GS iterations}.  Returns 0 if it
 * ran, > 0 if the stream is not representable (caller falls back to one call per node). */
int mgCoarseTailMaxN(void);
int mgCoarseTailMaxOps(void);
int mgCoarseTail(double L, double *U_entry, double *F_entry, int n_ops, const int *kind, const int *step,
                 const int *zero_init, const int *N_of_op, const int *next_N, const double *target, const int *option,
                 double *out_slots);

/* ---------------------------------------------------------------- the driver
 * Cycle.txt interpreter = the reference's main() loop (MG_solver_CPU.cpp:36-462)
 * over a device-resident level stack (linkedlist.cpp). */
typedef struct mgTraceRec {
    int node;       /* -1, 0, 1 */
    int N;          /* grid the node smoothed / solved on */
    int steps;      /* sweeps done (trigger mode: counted); node 0: GS iterations or -1 */
    double err;     /* smoothing error after the node (0 for node 0) */
} mgTraceRec;

typedef struct mgCycleResult {
    int n_recs;
    int N;              /* top-level grid size */
    double mg_error;    /* mean |analytic - U| */
    double time_ms;     /* device time of the node loop (CUDA events) */
    double wall_ms;     /* host wall time of the node loop incl. final sync */
    int launches;       /* kernels launched inside the node loop */
} mgCycleResult;

#define MG_RUN_FUSED 1          /* use mgDownLeg/mgUpLeg where the cycle permits (default path) */
#define MG_RUN_UNFUSED 0        /* one of the eight operators per reference call */
#define MG_RUN_QUIET 2          /* do not print the reference's stdout log */
#define MG_RUN_SKIP_SOURCE 4    /* F of the top level is supplied by the caller (F_top, used in place, never written) */
#define MG_RUN_NO_FINAL_ERROR 8 /* skip the analytic-error report after the node loop (mg_error = 0) */
#define MG_RUN_DEFER_HARVEST 16 /* mgRunSubcycle only: do not synchronise at the end; the error scalars reach
                                 * `recs` at the next mgSubcycleHarvest (or when the slot ring fills) */

/* Runs a cycle file.  F_top (device, N_max^2) is used instead of getSource when
 * MG_RUN_SKIP_SOURCE is set.  U_top, if non-NULL (device, N_max^2), receives the
 * final solution.  Returns 0 on success. */
int mgRunCycleFile(const char *path, int flags, const double *F_top, double *U_top,
                   mgTraceRec *recs, int max_recs, mgCycleResult *res);

/* The problem plug point (the reference compiles the problem in: source MG_solver_CPU.cpp:488, boundary :509-519,
 * analytic solution :544).  Any member may be NULL (= the reference's built-in).  All grids are device arrays of N_max^2.
 *   F_top         the source grid (used in place, never written)
 *   U0_top        an initial grid INCLUDING its boundary values: non-zero Dirichlet data.  The top level then starts the
 *                 way the reference restarts (:209-211): its first -1 node keeps U instead of zeroing it; sweeps touch
 *                 interior points only (:587-599), so the boundary data are carried and enter every residual.
 *   analytic_top  the reference solution of the final error report (:434-445) */
typedef struct mgProblem {
    const double *F_top;
    const double *U0_top;
    const double *analytic_top;
} mgProblem;
int mgRunCycleFileEx(const char *path, int flags, const mgProblem *prob, double *U_top, mgTraceRec *recs, int max_recs,
                     mgCycleResult *res);
/* the same with HOST grids (each may be NULL) */
int mgRunCycleFileHostEx(const char *path, int flags, const double *F_host, const double *U0_host, const double *analytic_host,
                         double *U_host, mgTraceRec *recs, int max_recs, mgCycleResult *res);

/* The same interpreter on ONE level owned by the caller: runs the node sub-stream that starts at
 * tok[*cur] (tok = the cycle file as numeric tokens) and returns to that level; stops before the
 * 1 node that would prolong above it, before the code 2, or at the end.  Used by the slab driver
 * for the levels agglomerated on rank 0 (execute == 0: parse only, keeps the other ranks in step). */
int mgRunSubcycle(const double *tok, int n_tok, int *cur, int *pos, const int *ladder, int n_ladder, int con_step,
                  int con_N, double L, int N, double **U, double **W, double *F, int depth_offset, int *init_io,
                  int flags, mgTraceRec *recs, int max_recs, int *n_recs_io, int execute);
/* Synchronises and moves the scalars of sub-cycles run with MG_RUN_DEFER_HARVEST into their records. */
void mgSubcycleHarvest(void);

/* Same with HOST buffers: F_host (N_max^2, may be NULL -> getSource on device) is
 * copied to the device, the cycle runs, the solution is copied back into U_host. */
int mgRunCycleFileHost(const char *path, int flags, const double *F_host, double *U_host,
                       mgTraceRec *recs, int max_recs, mgCycleResult *res);
/* n independent problems through the same cycle file, HOST buffers (pinned for full speed),
 * double-buffered: upload of problem i+1 and download of problem i-1 overlap the cycle of problem i.
 * F_hosts[i] / U_hosts[i]: N_max^2 doubles each, all non-NULL; res: n results or NULL.
 * Bit-identical to n mgRunCycleFileHost calls (which is what the reference's main() would do per problem,
 * MG_solver_CPU.cpp:149-462). */
int mgRunCycleFileHostBatch(const char *path, int flags, int n, const double *const *F_hosts, double *const *U_hosts,
                            mgCycleResult *res);

/* ------------------------------------------------------------ multi-GPU (row slabs)
 * One process per GPU.  Levels with at least `threshold` rows are partitioned into row slabs (one per
 * rank).  The halo rows of every array are stored straight into the neighbours' slabs by the kernel that
 * produces them (peer memory, CUDA IPC over NVLink) and announced through flag words the neighbours'
 * streams wait on; smaller levels are agglomerated: their source is broadcast to every rank and the coarse
 * sub-cycle runs redundantly everywhere.  Rendezvous: rank 0 calls mgDistUniqueId, the host program
 * broadcasts the 128 bytes (e.g. with torch.distributed), every rank calls mgDistInit (collective).
 * NCCL (bound with dlopen at run time) carries the rendezvous, the IPC handles and the scalar all-reduces. */
/* Host-only planning of the slab geometry for a ladder of level sizes: per level
 * [N, distributed?, bound[0..world]] (world+3 ints).  Needs no GPU.  Returns the number of
 * levels, or < 0 if a distributed level cannot be served (odd size / non-fusable pair). */
int mgDistPlan(const int *ladder, int n_levels, int world, int threshold, int *out, int max_out);
int mgDistUniqueId(void *out128);
int mgDistInit(int rank, int world, const void *id128);
void mgDistShutdown(void);
/* Runs a cycle file on all ranks collectively (fused driver, MG_RUN_QUIET / MG_RUN_NO_FINAL_ERROR
 * honoured).  U_own_host, if non-NULL, receives this rank's owned rows [*own_lo, *own_hi) of the
 * final solution (host buffer of at least N_max*N_max doubles is always enough).  Every rank gets the
 * complete trace records. */
int mgDistRunCycleFile(const char *path, int threshold, int flags, double *U_own_host, int *own_lo, int *own_hi,
                       mgTraceRec *recs, int max_recs, mgCycleResult *res);
/* Host-buffer path of the slab driver: geometry of this rank's top-level slab (rows
 * [row0, row0+rows) held incl. halo, [own_lo, own_hi) owned), upload of its source rows from a
 * host buffer (used by the next mgDistRunCycleFile calls that pass MG_RUN_SKIP_SOURCE), and
 * download of the cached source slab. */
int mgDistSourceSlab(int N, int threshold, int *row0, int *rows, int *own_lo, int *own_hi);
int mgDistUploadSource(int N, int threshold, const double *F_slab_host);
int mgDistDownloadSource(int N, double *F_slab_host);
/* n independent problems through the same cycle file, HOST buffers per rank (pinned for full speed):
 * F_slab_hosts[i] = this rank's source rows [row0, row0+rows) of problem i, U_own_hosts[i] receives its owned
 * rows.  Double-buffered like mgRunCycleFileHostBatch (upload i+1 / cycle i / download i-1 on two copy
 * streams per rank).  Collective; bit-identical to n (mgDistUploadSource, mgDistRunCycleFile) pairs. */
int mgDistRunCycleFileHostBatch(const char *path, int threshold, int flags, int n, const double *const *F_slab_hosts,
                                double *const *U_own_hosts, mgCycleResult *res);
/* doSmoothing on row slabs, repeated (smoothing-only stress of BASELINE config 5): `reps` x `step`
 * Jacobi sweeps from U = 0 on the analytic source, passes of <= 3 fused sweeps whose halo rows go
 * straight into the neighbours' slabs, error all-reduced per repetition.  N even, up to 65536.  Collective
 * over the ranks of mgDistInit; on one GPU it works without it.  U_own_host (optional): the owned rows. */
int mgDistSmoothStress(int N, double L, int step, int reps, double *ms_per_rep, double *error_out, double *U_own_host,
                       int *own_lo, int *own_hi);
/* The same slab algorithm with all `world` ranks emulated inside this process on the current GPU (the same
 * arenas, peer stores, flag words and stream waits; every rank's passes queued in rank order on one stream):
 * lets the slab logic be verified on one GPU.  F_host (N_max^2, may be NULL -> getSource on the device) is the
 * source grid, U_host (N_max^2) receives the assembled solution. */
int mgDistEmuRunCycleFile(const char *path, int world, int threshold, int flags, const double *F_host, double *U_host,
                          mgTraceRec *recs, int max_recs, mgCycleResult *res);

/* CSV dump in the reference's format (doPrint2File, MG_solver_CPU.cpp:735-754) from a host array */
int mgPrint2File(int N, const double *U_host, const char *file_name);

#ifdef __cplusplus
}
#endif
#endif
