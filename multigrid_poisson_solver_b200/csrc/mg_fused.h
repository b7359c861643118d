// mg_fused.h -- smoothing passes and the fused cycle legs (internal C++ API).
#pragma once
#include <vector>
#include "mg_context.h"

namespace mg {

void fused_init();
int set_tile_max_n(int n);   // n >= 0: whole grids up to n take the tile kernel (0: never); n < 0: default (odd sizes up to 1024); returns the old setting

// `step` Jacobi sweeps (MG_solver_CPU.cpp:578-601) followed by the smoothing error (:607-622).
// The first pass reads `in` (never written; treated as all zeros when in_is_zero) and the
// passes write alternately to `a`, `b`, `a`, ... (`in` may alias `b`).  Returns the buffer
// holding the result (`in` itself when step == 0).  The error goes to err_dev (device double)
// and, if non-null, to err_slot (device alias of a pinned scalar slot).
double *smooth_out_of_place(int N, double L, const double *in, double *a, double *b, const double *F, int step,
                            bool in_is_zero, double *err_dev, double *err_slot);
// number of out-of-place passes smooth_out_of_place will make
int smooth_pass_count(int N, int step);

// -1 node: [U = 0]; step sweeps; error; F_c = restrict(-(residual(U, F)))   (:246-287)
double *down_leg(int N, double L, double *U, double *U_work, const double *F, int step, bool zero_init, int M,
                 double *F_c, double *err_slot);

// 1 node: U_f += prolong(U_c); step sweeps; error                            (:350-416)
double *up_leg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int step,
               double *err_slot);

// Error-trigger loops (con_step = -1, :194-240 / :376-408) two sweeps per launch; *steps = sweeps done, *error = the last
// smoothing error (both valid on return: the loop synchronises once per launch).  Return the buffer holding the result.
bool trigger_fusable_down(int N, int M);
bool trigger_fusable_up(int Nc, int N);
double *down_leg_trigger(int N, double L, double *U, double *U_work, const double *F, bool zero_init, int M, double *F_c, int *steps,
                         double *error);
double *up_leg_trigger(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int *steps, double *error);

// ------------------------------------------------------------------ row slabs (multi-GPU, mg_dist.cu)
// A slab holds global rows [row0, row0+rows) of an N-column grid and owns [own_lo, own_hi).
struct Slab {
    int row0 = 0, rows = 0, own_lo = 0, own_hi = 0;
};
// even N, injective restriction map N -> M, prolongation ratio the staged coarse row can hold
bool slab_pair_fusable(int N, int M);
// smallest coarse row whose lower fine row (floor map of doRestriction) is >= fine_row; M if none
int restrict_first_coarse_at_or_after(int N, int M, int fine_row);
// Peer memory of a slab pass (k_strip): where the rows the neighbours keep as halos are ALSO stored, and the flag words
// that tell them the pass has drained.  Pointers are the neighbours' arrays pre-offset to GLOBAL rows; null = no neighbour.
struct PeerLinks {
    double *U_lo = nullptr, *U_hi = nullptr;       // output rows < u_lo_end -> lower neighbour, rows >= u_hi_begin -> upper neighbour
    int u_lo_end = 0, u_hi_begin = 0;
    double *Fc_lo = nullptr, *Fc_hi = nullptr;     // the same for the rows of the restricted grid
    int fc_lo_end = 0, fc_hi_begin = 0;
    unsigned int *flag_lo = nullptr, *flag_hi = nullptr;
    unsigned int flag_val = 0;
};
// One streaming pass (S <= 3 sweeps) on a slab.  in_mode: 0 load, 1 zero, 2 prolong (+add).
// want_err: the slab's plain red-parity sum goes to *raw_err_dev.  coarse_out != null: restrict
// the negated residual into the local F_c array of that coarse slab.  coarse_in: the local U_c.
void slab_pass(int N, double L, int S, int in_mode, const double *Uin, const double *F, double *Uout, const Slab &fine,
               bool want_err, double *raw_err_dev, int M, double *Fc, const Slab *coarse_out, int Nc, const double *Uc,
               const Slab *coarse_in, const PeerLinks &peers);
// one pass of a trigger loop on a slab: S = 2 (raw sums of both errors -> raw_err_dev2[0] last, [1] first sweep) or S = 1
void slab_trigger_pass(int N, double L, int S, int in_mode, const double *Uin, const double *F, double *Uout, const Slab &fine, double *raw_err_dev2,
                       int M, double *Fc, const Slab *coarse_out, int Nc, const double *Uc, const Slab *coarse_in, const PeerLinks &peers);
void dist_release_on_shutdown();   // mgShutdown: the slab driver's arenas and cached source go with the context

// Row segments {first, past-last} (relative to the first owned row) a fused pass over `rows` owned rows
// hands to its warps, in queue order; host-only (no device state).  subset as in slab_pass.
std::vector<int> segment_plan(int rows, int n_strips, int resident_warps, int lead_rows, int subset);

// ---- launch plumbing shared by mg_fused.cu (round-1 kernels) and mg_legs.cu (k_strip)
struct StreamParams;
// fills the task fields of `p` (strips, row segments, queue, partials), shifts the array bases to global rows;
// returns the CTA count of the persistent grid, 0 if there is nothing to launch
int stream_launch_prepare(StreamParams &p, int W, int warps, int min_ctas, bool err, int lead_rows);
// one fused pass with the 4-column bulk-copy kernel; in: 0 load, 1 zero, 2 prolong; mode: 0 plain, 1 ERR, 2 ERR+RES
void launch_strip(int S, int in, int mode, StreamParams &p);
// the -1 node (in 0 / 1, mode 2) or the 1 node (in 2, mode 0 / 1) on a slab with peer memory (mg_peer.cu)
void launch_stream_peer(int S, int in, int mode, StreamParams &p);
// two sweeps with both smoothing errors (error-trigger loops); in: 0 load, 1 zero (res only), 2 prolong (no res); peer stores
// when the parameters carry flag words (mg_peer.cu)
void launch_stream_mid(int in, bool res, StreamParams &p);

}  // namespace mg
