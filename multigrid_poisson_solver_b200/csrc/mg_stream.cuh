// mg_stream.cuh -- the temporally blocked, register-streaming stencil kernel.
//
// One kernel template covers every fused pass of the cycle:
//
//   level 0 (input)   IN_LOAD     U_in                                   (smoothing pass)
//                     IN_ZERO     0                                      (-1 node: U = 0, :252-257)
//                     IN_PROLONG  U_f + doProlongation(U_c)              (1 node: :354 + :368)
//   levels 1..S       S Jacobi sweeps (doSmoothing, :578-601), all in registers
//   residual stage    ERR: the red-parity error sum of doSmoothing (:607-622)
//                     RES: F_c = doRestriction(-(getResidual(U_S, F)))   (:268, :277-280, :287)
//
// Work decomposition.  A WARP owns a strip of W columns x H rows.  It streams down the rows of
// a 64-column window (2 adjacent columns per lane, one 16-byte load per lane and row, 512
// contiguous bytes per warp and row) that overlaps the neighbouring strips by the halo the S
// sweeps (+ residual, + restriction) consume; at step r it loads row r of level 0 and produces
// row r-1 of level 1, row r-2 of level 2, ... row r-S of level S and residual row r-S-1, so
// every level keeps only two rows per lane in registers and nothing but the final level is ever
// written.  Left/right neighbours come from the adjacent lanes by warp shuffle; there is no
// shared memory and no CTA barrier in the streaming loop.  Warps are independent, so the halo is
// recomputed rather than exchanged: 64/W in x and (H+2S+3)/H in y.
//
// Compulsory HBM traffic per fine point (fp64): smoothing pass 24 B (U in, F in, U out) for S
// sweeps instead of 24*S; -1 node 16 B + 8 B per coarse point (F in, U out, F_c out) instead of
// 146 B unfused; 1 node 24 B + 8 B per coarse point instead of 122 B.
//
// Bit parity: every expression uses the reference's association through mg_device.cuh.  The only
// fused-multiply-adds are (a) s4 - 4*u, exact because 4*u is exact, and (b) the quotient
// refinement of x / c_dx in the prolongation, which is correctly rounded (Markstein) and falls
// back to an IEEE division outside a safe exponent range.
#pragma once
#include <cuda_runtime.h>

#include "mg_device.cuh"

namespace mg {

enum { IN_LOAD = 0, IN_ZERO = 1, IN_PROLONG = 2 };

#ifndef MG_STREAM_DEPTH
#define MG_STREAM_DEPTH 4
#endif
#ifndef MG_ZERO_SHORTCUT
#define MG_ZERO_SHORTCUT 1       // IN_ZERO: first sweep without stencil
#endif
#ifndef MG_RES_EVEN
#define MG_RES_EVEN 1            // restriction: nested ladders skip the odd column
#endif
// CTA shape per variant (measured on B200, N = 16384): passes without restriction run best as
// 4-warp CTAs, three per SM (12 warps, up to 168 registers, no spills); the -1 node (RES) as
// 8-warp CTAs, two per SM (16 warps, 128 registers).
#ifndef MG_RES_WARPS
#define MG_RES_WARPS 8
#endif
#ifndef MG_RES_CTAS
#define MG_RES_CTAS 2
#endif
#ifndef MG_PLAIN_WARPS
#define MG_PLAIN_WARPS 4
#endif
#ifndef MG_PLAIN_CTAS
#define MG_PLAIN_CTAS 3
#endif
struct StreamShape { int warps, min_ctas; };
__host__ __device__ constexpr StreamShape stream_shape(bool res)
{
    return res ? StreamShape{MG_RES_WARPS, MG_RES_CTAS} : StreamShape{MG_PLAIN_WARPS, MG_PLAIN_CTAS};
}
constexpr int STREAM_SMAX = 3;          // sweeps fused per pass
constexpr int STREAM_DEPTH = MG_STREAM_DEPTH;   // rows in flight per warp (cp.async ring in shared memory), power of two
// shared memory: [warp][slot][U | F (| coarse row, 1 node only)][lane] x 16 B
// the 1 node adds the staged coarse row (512 B) and the row's {row_w, row_cell} entry (32 B)
// RES: 16 more bytes hold the restriction table entry of the fine row the step pairs up, copied by lane 0 and read by all
// lanes (a __syncwarp on each side).  It used to be a global load one step ahead: a quarter of all stall samples of the
// -1 node sat on the F2I that consumes it (ncu, long scoreboard).  (32 lanes copying the same 16 bytes each is not an
// option: same-address LDGSTS serialise -- measured 2.7x slower.)
__host__ __device__ constexpr int stream_slot_bytes(int in, bool res = false) { return in == 2 ? 3 * 512 + 32 : res ? 2 * 512 + 32 : 2 * 512; }
__host__ __device__ constexpr int stream_smem_bytes(int in, int warps, bool res = false)
{
    return warps * STREAM_DEPTH * stream_slot_bytes(in, res);
}

struct StreamParams {
    int N;                  // grid size (even): columns, and rows of the GLOBAL grid
    // Row slab (multi-GPU): the arrays hold global rows [row0, row0+rows); this call owns (writes,
    // sums, restricts) global rows [own_lo, own_hi).  Single GPU: row0 = 0, rows = N, own = [0, N).
    // ALL grid pointers below are pre-offset by the launcher so that they are indexed with GLOBAL
    // rows (Uin/F/Uout by -row0*N, Fc by -fc_row0*M, Uc by -uc_row0*Nc): the kernel never subtracts.
    int row0, rows, own_lo, own_hi;
    int fc_row0;            // global index of local row 0 of the F_c array
    int uc_row0, uc_rows;   // global index of local row 0 of the U_c array, and its local row count
    int raw_sum;            // ERR: write the plain red-parity sum (the host combines slabs) instead of (S+S)/N/N
    int subset;             // host only: 0 whole owned range; 1 / 2 edge / interior launch of a split pass (slab driver)
    int err_add;            // ERR: add this launch's sum to *err_dev (the other part of a split pass wrote it)
    unsigned long long *trace;   // debug (MG_TASK_TRACE): per task {start ns, end ns, SM, warp}; nullptr normally
    const int2 *segs;       // [n_segs] row segments {first, past-last}, relative to own_lo; one task = (strip, segment)
    int n_strips, n_segs, n_tasks;   // tasks (strip, row segment of the chosen subset) are handed to warps through an atomic queue
    double h2, inv_h2;
    const double *F_valid;  // any dereferenceable address (source operand of zero-fill copies)
    const double *Uin;      // IN_LOAD: U ; IN_PROLONG: U_f
    const double *F;
    double *Uout;
    // ERR
    double *partials;       // [n_tasks] per-task error partials (fixed order => deterministic sum)
    unsigned int *counter;  // [0] task queue head, [1] warps finished; both return to 0 at kernel end
    double *err_dev, *err_slot;
    // RES (restriction of the negated residual)
    int M;
    double *Fc;
    const int *f2c;         // [N]  fine index -> coarse index whose lower-left fine point it is, or -1
    const double *rw;       // [M]  fmod weight of the coarse index
    const double2 *rrow;    // [N]  per fine row: {coarse row (as a double) or -1, its fmod weight}
    // IN_PROLONG
    int Nc;
    const double *Uc;
    const int *row_cell, *col_cell;
    const double2 *row_w, *col_w;
    const double4 *row_info;   // [N] {row_w.x, row_w.y, row_cell, 0} (k_strip: one bulk copy per row)
    double c_dx, inv_c_dx;
    // Row slabs with peer memory (k_strip only; all null / unused on one GPU).  The rows of the output a neighbour
    // keeps as its halo are ALSO stored straight into the neighbour's array (pointers pre-offset to global rows like
    // Uout / Fc): owned rows < u_lo_end go to the lower neighbour, owned rows >= u_hi_begin to the upper one; the same
    // for the rows of the restricted grid.  When the launch has drained, its last CTA publishes `flag_val` in the
    // neighbours' flag words (release at system scope) -- the consumer's stream waits for it before its next pass.
    double *peer_U_lo, *peer_U_hi;
    int u_lo_end, u_hi_begin;
    double *peer_Fc_lo, *peer_Fc_hi;
    int fc_lo_end, fc_hi_begin;
    unsigned int *flag_lo, *flag_hi;
    unsigned int flag_val;
};

template <int S, bool NEED_R, bool RES>
struct StreamGeo {
    static constexpr int HL = (S + (NEED_R ? 1 : 0) + 1) / 2 * 2;                 // left halo, even (16-byte loads)
    static constexpr int HR = S + (RES ? 2 : NEED_R ? 1 : 0);                    // right halo
    static constexpr int W = (64 - HL - HR) / 2 * 2;                             // owned columns per strip, even
    static constexpr int ROW_LEAD = S + (NEED_R ? 1 : 0);                        // rows streamed before the first owned row
    static constexpr int ROW_TAIL = S + (RES ? 2 : NEED_R ? 1 : 0);              // steps after the last owned row
};

// x / d, correctly rounded, with y = 1/d (correctly rounded, host side).  Two Newton
// corrections make the quotient faithful, then Markstein's final step rounds it correctly.
__device__ __forceinline__ double div_by_invariant(double x, double d, double y)
{
    const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
    if (e - 200u > 1600u) return __ddiv_rn(x, d);   // zero, subnormal, huge, inf, nan: IEEE path
    const double q0 = __dmul_rn(x, y);
    const double r0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, x);
    return __fma_rn(r1, y, q1);
}

// The same quotient without the range test (callers test div_unsafe and redo the rare cases).
__device__ __forceinline__ double div_fast(double x, double d, double y)
{
    const double q0 = __dmul_rn(x, y);
    const double r0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, x);
    return __fma_rn(r1, y, q1);
}
__device__ __forceinline__ bool div_unsafe(double x)
{
    const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
    return e - 200u > 1600u;                       // zero, subnormal, huge, inf, nan
}
// Guard for TWO chained divisions (x / d) / d with 2^-20 <= d <= 1: if x passes, so does the first
// quotient (its exponent grows by at most 20), so one test covers both.  +0 passes too -- div_fast
// maps it to +0 exactly, and it is what every boundary column and off-grid lane feeds in (flagging
// it made the two edge strips of every row segment take the IEEE path on every row).
__device__ __forceinline__ bool div2_unsafe(double x)
{
    const unsigned hi = (unsigned)__double2hiint(x), e = (hi >> 20) & 0x7ffu;
    return (e - 220u > 1560u) && ((hi | (unsigned)__double2loint(x)) != 0u);
}

// Cheap superset of div2_unsafe(x) | div2_unsafe(y) for the common path (no zero exemption, one compare for both):
// the caller evaluates the precise test only when some lane reports a suspect.
__device__ __forceinline__ bool div2_suspect(double x, double y)
{
    const unsigned ax = ((unsigned)__double2hiint(x) & 0x7fffffffu) - (220u << 20);
    const unsigned ay = ((unsigned)__double2hiint(y) & 0x7fffffffu) - (220u << 20);
    return max(ax, ay) > (1561u << 20) - 1u;       // biased exponent outside [220, 1780]
}

// ---- end of a launch: the last CTA to drain the queue resets it and folds the per-task error
// partials.  Thread t adds partials t, t+T, ... in that order, then the fixed shuffle tree and the
// warps in order: the summation tree depends only on the task geometry and the CTA shape, never on
// which warp ran which task.  A whole CTA keeps 16 loads per thread in flight (a lone warp needed
// ~1 us per 1024 partials: 14-20 us of tail at N >= 8192).
template <int WARPS, bool ERR, bool MID = false>
__device__ __forceinline__ void finish_launch(const StreamParams &p)
{
    __shared__ double fold_s[WARPS];
    __shared__ unsigned int fold_last;
    constexpr int T = WARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool peers = p.flag_lo || p.flag_hi;
    if (peers) __threadfence_system();            // every thread's halo rows in the neighbours' memory, before the ticket
    else if (lane == 0) __threadfence();          // this warp's partials, before the CTA's ticket
    __syncthreads();                              // every warp of this CTA has drained the queue
    if (tid == 0) fold_last = atomicAdd(p.counter + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!fold_last) return;
    __threadfence();
    if (peers && tid == 0) {                      // all CTAs have fenced their peer stores: publish the pass number
        __threadfence_system();
        if (p.flag_lo) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag_lo), "r"(p.flag_val) : "memory");
        if (p.flag_hi) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag_hi), "r"(p.flag_val) : "memory");
    }
    if (ERR) {
        // which: 0 the error after the last sweep -> err[0]; 1 (MID) the error after the first of two sweeps -> err[1]
#pragma unroll
        for (int which = 0; which < (MID ? 2 : 1); ++which) {
            const double *part = p.partials + (size_t)which * p.n_tasks;
            double s = 0.0;
            for (int k0 = tid; k0 < p.n_tasks; k0 += T * 16) {
                double v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = k0 + T * j < p.n_tasks ? __ldcg(&part[k0 + T * j]) : 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) s = __dadd_rn(s, v[j]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, off));
            if (which) __syncthreads();           // fold_s is reused
            if (lane == 0) fold_s[warp] = s;
            __syncthreads();
            if (tid == 0) {
                s = fold_s[0];
#pragma unroll
                for (int w = 1; w < WARPS; ++w) s = __dadd_rn(s, fold_s[w]);
                double e = s;
                if (p.err_add && p.err_dev) e = __dadd_rn(p.err_dev[which], s);   // second launch of a split pass
                if (!p.raw_sum) {
                    e = __dadd_rn(s, s);                          // sum1 + sum2 over the same parity (:621)
                    e = __ddiv_rn(e, (double)p.N);
                    e = __ddiv_rn(e, (double)p.N);
                }
                if (p.err_dev) p.err_dev[which] = e;
                if (p.err_slot) p.err_slot[which] = e;            // host-mapped; the host reads it after a stream sync
            }
        }
    }
    if (tid == 0) {
        p.counter[0] = 0u;
        p.counter[1] = 0u;
        if (p.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.trace[4 * p.n_tasks]));   // after the fold
    }
}

// s4 - 4*u with one rounding == the reference's (s4 - RN(4*u)) because 4*u is exact.
__device__ __forceinline__ double sub4(double s4, double u) { return __fma_rn(-4.0, u, s4); }

__device__ __forceinline__ double jacobi_fast(double u, double s4, double h2f)
{
    return __dadd_rn(u, __dmul_rn(0.25, __dsub_rn(sub4(s4, u), h2f)));
}
__device__ __forceinline__ double residual_fast(double u, double s4, double f, double inv_h2)
{
    return __dsub_rn(__dmul_rn(inv_h2, sub4(s4, u)), f);
}

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }

// 16-byte asynchronous global->shared copy (LDGSTS, L2 only: streamed rows must not evict the
// few local-memory lines from L1); !valid zero-fills the destination.
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void *gsrc, bool valid)
{
    const int src_bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
// 8-byte variant (coarse rows of odd length are only 8-byte aligned)
__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void *gsrc, bool valid)
{
    const int src_bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_addr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const void *gsrc, bool valid)
{
    const int src_bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ int lds_int(unsigned smem_addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds1(unsigned smem_addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_addr) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }
__device__ __forceinline__ double2 lds2(unsigned smem_addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts2(unsigned smem_addr, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(smem_addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <bool B>
struct BoolTag { static constexpr bool value = B; };

// PEER: row slabs with peer memory -- the rows a neighbouring GPU keeps as its halo are also stored straight into its
// array (a separate instantiation, mg_peer.cu: the single-GPU kernels sit exactly at their register limits).
// MID (S == 2 only): also the smoothing error after the FIRST of the two sweeps (partials [n_tasks, 2 n_tasks), result in
// err[1]) -- the error-trigger loop (:216-230) takes two sweeps per launch and still sees every sweep's error.
template <int S, int IN, bool ERR, bool RES, bool PEER = false, bool MID = false>
__global__ void __launch_bounds__(stream_shape(RES).warps * 32, stream_shape(RES).min_ctas) k_stream(const StreamParams p)
{
    static_assert(!MID || (S == 2 && ERR), "MID: two sweeps with both errors");
    constexpr int STREAM_WARPS = stream_shape(RES).warps;
    constexpr bool NEED_R = ERR || RES;
    using G = StreamGeo<S, NEED_R, RES>;
    constexpr int NLV = S + (NEED_R ? 1 : 0);    // levels that keep a two-row window (level t feeds stage t)
    constexpr int NF = NLV;                      // F rows alive at once: rows r-1 ... r-NF
    constexpr int NR = NF <= 2 ? 2 : 4;          // F ring size; also the unroll factor (even: the row windows have period 2)
    constexpr int U = NR;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;
    const double h2 = p.h2, inv_h2 = p.inv_h2;
    const double *__restrict__ Fp = p.F;
    const double *__restrict__ Up = p.Uin;
    double *__restrict__ Op = p.Uout;
    const ptrdiff_t ldn = N;

    // Streamed rows are staged through a per-warp ring in shared memory filled by cp.async:
    // STREAM_DEPTH rows of U and F in flight per warp at no register cost.  Each lane copies
    // and later reads back only its own 16 bytes, so no barrier is involved, only wait_group.
    extern __shared__ __align__(16) unsigned char stream_smem[];
    constexpr int SLOT_BYTES = stream_slot_bytes(IN, RES);   // [U | F | coarse row or restriction table entry][lane] x 16 B
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(stream_smem);
    const unsigned warp_ring = smem0 + warp * (STREAM_DEPTH * SLOT_BYTES);
    const unsigned ring_base = warp_ring + lane * 16;
    const unsigned coarse_dst = warp_ring + 1024 + lane * 8;    // IN_PROLONG: this lane's piece of the staged coarse row

  // Persistent warps: every warp pulls (strip, row segment) tasks from an atomic queue until it is
  // empty, so there is no wave quantisation and no CTA waits for its slowest warp.
  for (;;) {
    int task = 0;
    if (lane == 0) task = (int)atomicAdd(p.counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= p.n_tasks) break;
    unsigned long long trace_t0 = 0;
    if (p.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
    const int seg_idx = task / p.n_strips;
    const int strip = task - seg_idx * p.n_strips;
    const int seg = seg_idx;                                    // consecutive tasks = adjacent strips of one row segment
    constexpr bool active = true;

    const int own_c_lo = strip * G::W, own_c_hi = min(own_c_lo + G::W, N);
    const int c_first = own_c_lo - G::HL;                       // first column of the 64-wide window (even)
    const int cx = c_first + 2 * lane;                          // this lane's columns: cx, cx+1
    const bool col_ok = active && cx >= 0 && cx < N;            // N even => cx+1 < N too
    const bool col_own = col_ok && cx >= own_c_lo && cx < own_c_hi;
    const bool x_in = cx > 0, y_in = cx + 1 < N - 1;            // interior columns
    const bool strip_fast = c_first >= 1 && c_first + 63 <= N - 2;   // every column of the window is interior
    // all row indices below are GLOBAL; the arrays are addressed with (row - row0)
    const int2 seg_rows = __ldg(p.segs + seg);                  // rows of the segment relative to own_lo
    const int own_r_lo = p.own_lo + seg_rows.x, own_r_hi = p.own_lo + seg_rows.y;
    // even, so that the parity of every row index of an unrolled chunk is a compile-time fact (the red-parity error
    // then needs the residual of ONE of the lane's two columns); the extra leading row is one more warm-up row
    const int r_first = max(0, own_r_lo - G::ROW_LEAD) & ~1;
    const int r_last = min(own_r_hi - 1 + G::ROW_TAIL, N - 1 + G::ROW_LEAD);
    // rows of this task whose output is also a neighbour's halo (slabs with peer memory)
    const bool peer_rows = PEER && ((p.peer_U_lo && own_r_lo < p.u_lo_end) || (p.peer_U_hi && own_r_hi > p.u_hi_begin));

    // Register state.  Every index below is a compile-time constant after unrolling, so the
    // arrays live in registers and rotate by renaming, not by moves.
    double2 w[NLV > 0 ? NLV : 1][2];             // w[t][slot]: the two newest rows of level t
    double2 fr[NR];                              // F ring: the row that arrives at in-chunk step k sits in fr[k % NR]
#pragma unroll
    for (int t = 0; t < (NLV > 0 ? NLV : 1); ++t) w[t][0] = w[t][1] = make_double2(0.0, 0.0);
#pragma unroll
    for (int t = 0; t < NR; ++t) fr[t] = make_double2(0.0, 0.0);

    unsigned slot_off = 0;                       // byte offset of the slot that holds the row of the current step

    // ---- restriction state
    double2 d_prev = make_double2(0.0, 0.0);
    int ccx = -1, ccy = -1;
    double ax = 0.0, ay = 0.0;
    bool zx = true, zy = true;                   // coarse column on the coarse boundary: value forced to 0

    if (RES && col_own) {
        ccx = p.f2c[cx];
        ccy = p.f2c[cx + 1];
        if (ccx >= 0) ax = p.rw[ccx];
        if (ccy >= 0) ay = p.rw[ccy];
        zx = ccx == 0 || ccx == p.M - 1;
        zy = ccy == 0 || ccy == p.M - 1;
    }
    const bool res_even = MG_RES_EVEN && RES && __all_sync(0xffffffffu, ccy < 0);   // nested ladder: coarse points at even fine columns only
    // ---- prolongation state
    // Coarse rows are staged through the third part of each ring slot: the slot of fine row r
    // holds doubles [cbase, cbase+64) of coarse row row_cell[r]+1 (the upper row of its cell).
    // row_cell[] / row_w[] of fine row r travel in the ring slot of row r; the coarse row index needed
    // when a copy is ISSUED (DEPTH rows ahead) comes from a lane-distributed table: lane l holds
    // row_cell[tab_base + l], looked up by shuffle and refilled every 32 rows (double buffered).
    int cqx = 0, cqy = 0, cbase = 0, tab = 0, tab_nxt = 0, tab_base = 0;
    int prev_rq = -4;
    double2 bot = make_double2(0.0, 0.0), top = bot;
    unsigned ox = 0, oy = 0;                     // byte offsets of this lane's two cells inside a staged coarse row
    double2 wcx, wcy;                            // this lane's column weights
    if (IN == IN_PROLONG) {
        wcx = wcy = make_double2(0.0, 0.0);
        if (col_ok) {
            cqx = p.col_cell[cx];
            cqy = p.col_cell[cx + 1];
            wcx = p.col_w[cx];
            wcy = p.col_w[cx + 1];
        }
        tab_base = r_first;
        tab = p.row_cell[min(r_first + lane, N - 1)];
        tab_nxt = p.row_cell[min(r_first + 32 + lane, N - 1)];
        cbase = __reduce_min_sync(0xffffffffu, col_ok ? cqx : 0x7fffffff) & ~1;
        ox = col_ok ? (unsigned)(cqx - cbase) * 8u : 0u;     // lanes outside the grid read slot element 0 (unused)
        oy = col_ok ? (unsigned)(cqy - cbase) * 8u : 0u;
    }
    double err_acc = 0.0, mid_acc = 0.0;

    auto cell_of_row = [&](int r) {              // row_cell[min(r, N-1)] for the row being issued (non-decreasing r)
        if (r - tab_base >= 32) {
            tab = tab_nxt;
            tab_base += 32;
            tab_nxt = p.row_cell[min(tab_base + 32 + lane, N - 1)];
        }
        return __shfl_sync(0xffffffffu, tab, r - tab_base);
    };
    // guarded issue of level-0 row r and of F row r-1 (first used at step r) into ring slot `off`;
    // for the 1 node also the upper coarse row of fine row r's cell (`rq` = row_cell[r])
    auto issue = [&](int r, unsigned off, int rq) {
        if (IN != IN_ZERO) {
            const bool ok = col_ok && r >= p.row0 && r < p.row0 + p.rows;
            cp_async16(ring_base + off, ok ? (const void *)(Up + (ptrdiff_t)r * ldn + cx) : (const void *)p.F_valid, ok);
        }
        if (NF > 0) {
            const bool ok = col_ok && r - 1 >= p.row0 && r - 1 < p.row0 + p.rows;
            cp_async16(ring_base + off + 512, ok ? (const void *)(Fp + (ptrdiff_t)(r - 1) * ldn + cx) : (const void *)p.F_valid, ok);
        }
        if (RES && lane == 0) {                          // {coarse row or -1, weight} of fine row r-S-2: what step r pairs with the row above
            const int fr_ = r - S - 2;
            const bool ok = fr_ >= 0 && fr_ <= N - 1;
            cp_async16(warp_ring + off + 1024, ok ? (const void *)(p.rrow + fr_) : (const void *)p.F_valid, ok);
        }
        if (IN == IN_PROLONG) {
            // lane l copies elements l and 32 + l of the 64 staged doubles: each copy instruction fills 256 contiguous
            // bytes (interleaved 8-byte pieces at a 16-byte stride cost two extra shared-memory wavefronts per instruction)
            const int c0 = cbase + lane;
            const bool row_ok = active && r <= N - 1;
            // rows prefetched beyond the ones a task needs may map outside the local coarse slab: clamp (never used)
            const double *src = p.Uc + (ptrdiff_t)min(max(rq + 1, p.uc_row0), p.uc_row0 + p.uc_rows - 1) * p.Nc + c0;
            const bool ok0 = row_ok && c0 < p.Nc, ok1 = row_ok && c0 + 32 < p.Nc;
            cp_async8(coarse_dst + off, ok0 ? (const void *)src : (const void *)p.F_valid, ok0);
            cp_async8(coarse_dst + off + 256, ok1 ? (const void *)(src + 32) : (const void *)p.F_valid, ok1);
            if (lane == 0) {
                const int rr = row_ok ? r : 0;
                cp_async16(warp_ring + off + 1536, p.row_w + rr, row_ok);
                cp_async4(warp_ring + off + 1552, p.row_cell + rr, row_ok);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < STREAM_DEPTH; ++d) {
        int rq = 0;
        if (IN == IN_PROLONG) rq = cell_of_row(r_first + d);
        issue(r_first + d, d * SLOT_BYTES, rq);
    }
    if (IN == IN_PROLONG && active) {
        const int rq_first = __shfl_sync(0xffffffffu, tab, 0);
        // lower coarse row of the first cell: the only one that is not staged
        const double *c_lo = p.Uc + (ptrdiff_t)max(rq_first, p.uc_row0) * p.Nc;   // (clamped: only a warm-up row can ask for less)
        double2 low = make_double2(0.0, 0.0);
        if (col_ok) {
            low.x = __dadd_rn(__dmul_rn(c_lo[cqx], wcx.x), __dmul_rn(c_lo[cqx + 1], wcx.y));
            low.y = __dadd_rn(__dmul_rn(c_lo[cqy], wcy.x), __dmul_rn(c_lo[cqy + 1], wcy.y));
        }
        top = low;
        prev_rq = rq_first - 1;                   // so that the first step shifts `top` down and takes the staged upper row
    }
    // Level 0 of the 1 node, U_f + P(U_c) (MG_solver_CPU.cpp:700 + :569), for the fine row staged in
    // ring slot `so`; lanes whose quotient left the safe range of the FMA division report `bad`.
    // The cell's two coarse rows, interpolated in x, are register rows that move under a warp-uniform branch when the
    // cell changes (every other fine row of a nested ladder); the lane's column weights stay in registers.  (Measured
    // against selects + weights in shared memory: -7 % -- the kernel is bound by issue slots and shared-memory bandwidth
    // together, and under sustained load by the 1000 W power limit.)
    double2 x_next = make_double2(0.0, 0.0);
    double pvx = 0.0, pvy = 0.0, puf_x = 0.0, puf_y = 0.0;   // operands kept for the rare IEEE redo
    auto prolong_from_slot = [&](unsigned so) -> bool {
        const double2 uf = lds2(ring_base + so);
        const double2 wr = lds2(warp_ring + so + 1536);          // row_w[r]
        const int rq = lds_int(warp_ring + so + 1552);           // row_cell[r]
        const unsigned cs = warp_ring + so + 1024;
        if (rq != prev_rq) {                                     // warp-uniform: the cell moved up one coarse row
            bot = top;
            top.x = __dadd_rn(__dmul_rn(lds1(cs + ox), wcx.x), __dmul_rn(lds1(cs + ox + 8), wcx.y));
            top.y = __dadd_rn(__dmul_rn(lds1(cs + oy), wcy.x), __dmul_rn(lds1(cs + oy + 8), wcy.y));
            prev_rq = rq;
        }
        const double vx = __dadd_rn(__dmul_rn(bot.x, wr.x), __dmul_rn(top.x, wr.y));
        const double vy = __dadd_rn(__dmul_rn(bot.y, wr.x), __dmul_rn(top.y, wr.y));
        const double d = p.c_dx, y = p.inv_c_dx;
        const double qx = div_fast(vx, d, y), qy = div_fast(vy, d, y);
        x_next.x = __dadd_rn(uf.x, div_fast(qx, d, y));
        x_next.y = __dadd_rn(uf.y, div_fast(qy, d, y));
        pvx = vx; pvy = vy; puf_x = uf.x; puf_y = uf.y;
        return div2_suspect(vx, vy);                             // (a superset of the unsafe lanes: zeros, off-grid lanes)
    };
    auto prolong_redo = [&](bool suspect) {                      // rare: IEEE divisions for the whole warp
        if (__any_sync(0xffffffffu, suspect)) {
            // the precise test: +0 is safe (every boundary column feeds it in), off-grid lanes compute -0 = c * 0 (never stored)
            const bool bad = col_ok && (div2_unsafe(pvx) | div2_unsafe(pvy));
            if (__any_sync(0xffffffffu, bad)) {
                const double d = p.c_dx;
                x_next.x = __dadd_rn(puf_x, __ddiv_rn(__ddiv_rn(pvx, d), d));
                x_next.y = __dadd_rn(puf_y, __ddiv_rn(__ddiv_rn(pvy, d), d));
            }
        }
    };
    if (IN == IN_PROLONG) {
        cp_async_wait<STREAM_DEPTH - 1>();        // the group of row r_first has landed
        __syncwarp();
        prolong_redo(prolong_from_slot(0));
    }

    // One chunk = U consecutive steps.  FAST: every row touched by every stage is an interior
    // row, every column of the window is an interior column and all loads are in range, so the
    // body carries no boundary selects at all.
    auto chunk = [&](auto fast_tag, const int rb) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = rb + k;
            bool bad = false;
            double2 x = make_double2(0.0, 0.0), f_new = x, ri_cur = make_double2(-1.0, 0.0);
            if (IN == IN_PROLONG) {
                // level 0 of row r was produced one step ago; produce row r+1 now, next to this row's sweeps
                cp_async_wait<STREAM_DEPTH - 2>();                // the groups of rows r and r+1 have landed
                __syncwarp();                                     // slots are filled by all lanes (and lane 0's table entry)
                x = x_next;
                if (NF > 0) f_new = lds2(ring_base + slot_off + 512);
                unsigned slot_n = slot_off + SLOT_BYTES;
                if (slot_n == STREAM_DEPTH * SLOT_BYTES) slot_n = 0;
                bad = prolong_from_slot(slot_n);
            } else {
                cp_async_wait<STREAM_DEPTH - 1>();                // the group of row r has landed
                if (RES) __syncwarp();                            // ... lane 0's table entry for every lane
                if (IN != IN_ZERO) x = lds2(ring_base + slot_off);
                if (NF > 0) f_new = lds2(ring_base + slot_off + 512);
                if (RES) {
                    ri_cur = lds2(warp_ring + slot_off + 1024);
                    __syncwarp();                                 // every lane has read it: lane 0 may refill the slot
                }
            }

            // refill the slot with row r + DEPTH (the values above are in registers by now)
            if (FAST) {
                if (IN != IN_ZERO) cp_async16(ring_base + slot_off, Up + (ptrdiff_t)(r + STREAM_DEPTH) * ldn + cx);
                if (NF > 0) cp_async16(ring_base + slot_off + 512, Fp + (ptrdiff_t)(r + STREAM_DEPTH - 1) * ldn + cx);
                if (RES && lane == 0) cp_async16(warp_ring + slot_off + 1024, p.rrow + (r + STREAM_DEPTH - S - 2));
                if (IN == IN_PROLONG) {
                    const int c0 = cbase + lane;
                    const int rq_iss = cell_of_row(r + STREAM_DEPTH);
                    const double *src = p.Uc + (ptrdiff_t)min(max(rq_iss + 1, p.uc_row0), p.uc_row0 + p.uc_rows - 1) * p.Nc + c0;
                    const bool ok0 = c0 < p.Nc, ok1 = c0 + 32 < p.Nc;
                    cp_async8(coarse_dst + slot_off, ok0 ? (const void *)src : (const void *)p.F_valid, ok0);
                    cp_async8(coarse_dst + slot_off + 256, ok1 ? (const void *)(src + 32) : (const void *)p.F_valid, ok1);
                    if (lane == 0) {
                        cp_async16(warp_ring + slot_off + 1536, p.row_w + r + STREAM_DEPTH, true);
                        cp_async4(warp_ring + slot_off + 1552, p.row_cell + r + STREAM_DEPTH, true);
                    }
                }
                cp_async_commit();
            } else {
                issue(r + STREAM_DEPTH, slot_off, IN == IN_PROLONG ? cell_of_row(r + STREAM_DEPTH) : 0);
            }

            slot_off += SLOT_BYTES;
            if (slot_off == STREAM_DEPTH * SLOT_BYTES) slot_off = 0;

            if (NF > 0) fr[k % NR] = f_new;

            // ---- S sweeps: stage t turns level t row (r-t-1) into level t+1
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const int i = r - t - 1;
                const double2 f = fr[(k - t + 4 * NR) % NR];
                double2 nx;
                if (MG_ZERO_SHORTCUT && IN == IN_ZERO && t == 0) {
                    // level 0 is all zeros: jacobi_at(0, 0, h2 f) = 0 + 0.25*((0 - 0) - h2 f) with the same roundings -- no
                    // stencil, no shuffles, no level-0 window
                    nx.x = __dadd_rn(0.0, __dmul_rn(0.25, __dsub_rn(0.0, __dmul_rn(h2, f.x))));
                    nx.y = __dadd_rn(0.0, __dmul_rn(0.25, __dsub_rn(0.0, __dmul_rn(h2, f.y))));
                    if (!FAST) {
                        const bool row_in = i > 0 && i < N - 1;
                        nx.x = (row_in && x_in) ? nx.x : 0.0;
                        nx.y = (row_in && y_in) ? nx.y : 0.0;
                    }
                } else {
                    const double2 below = w[t][k & 1], c = w[t][(k & 1) ^ 1];
                    const double left = shfl_up1(c.y), right = shfl_dn1(c.x);
                    const double s4x = sum4(x.x, below.x, c.y, left), s4y = sum4(x.y, below.y, right, c.x);
                    nx.x = jacobi_fast(c.x, s4x, __dmul_rn(h2, f.x));
                    nx.y = jacobi_fast(c.y, s4y, __dmul_rn(h2, f.y));
                    if (!FAST) {                                  // boundary rows / columns are carried over
                        const bool row_in = i > 0 && i < N - 1;
                        nx.x = (row_in && x_in) ? nx.x : c.x;
                        nx.y = (row_in && y_in) ? nx.y : c.y;
                    }
                    if (MID && t == 1) {
                        // the smoothing error of level 1 (after the first sweep): its row i is the centre of this very stencil
                        const bool i_odd = (k & 1) != 0;              // i = rb + k - 2, rb even
                        double v = i_odd ? residual_fast(c.y, s4y, f.y, inv_h2) : residual_fast(c.x, s4x, f.x, inv_h2);   // red point (:609-611)
                        if (!FAST) {
                            const bool row_in = i > 0 && i < N - 1;
                            v = (row_in && (i_odd ? y_in : x_in)) ? v : 0.0;
                        }
                        const bool take = col_own && i >= own_r_lo && i < own_r_hi;
                        mid_acc = __dadd_rn(mid_acc, take ? fabs(v) : 0.0);
                    }
                    w[t][k & 1] = x;
                }
                x = nx;
            }

            // ---- x is now level S, row r-S
            {
                const int i = r - S;
                if (Op && i >= own_r_lo && i < own_r_hi && col_own) {
                    const ptrdiff_t o = (ptrdiff_t)i * ldn + cx;
                    *reinterpret_cast<double2 *>(Op + o) = x;
                    if (PEER && peer_rows) {             // slabs with peer memory: the same row into the neighbour's halo
                        if (p.peer_U_lo && i < p.u_lo_end) *reinterpret_cast<double2 *>(p.peer_U_lo + o) = x;
                        if (p.peer_U_hi && i >= p.u_hi_begin) *reinterpret_cast<double2 *>(p.peer_U_hi + o) = x;
                    }
                }
            }

            if (NEED_R) {
                const int rho = r - S - 1;                        // residual row
                const double2 below = w[S][k & 1], c = w[S][(k & 1) ^ 1];
                const double2 f = fr[(k - S + 4 * NR) % NR];
                const bool rho_odd = ((k + S + 1) & 1) != 0;      // rb is even: known after unrolling
                double2 res = make_double2(0.0, 0.0);
                if (RES || !rho_odd) {                            // (without RES only the red column's residual is needed)
                    const double left = shfl_up1(c.y);
                    res.x = residual_fast(c.x, sum4(x.x, below.x, c.y, left), f.x, inv_h2);
                }
                if (RES || rho_odd) {
                    const double right = shfl_dn1(c.x);
                    res.y = residual_fast(c.y, sum4(x.y, below.y, right, c.x), f.y, inv_h2);
                }
                if (!FAST) {                                      // 0 on the boundary (:559)
                    const bool row_in = rho > 0 && rho < N - 1;
                    res.x = (row_in && x_in) ? res.x : 0.0;
                    res.y = (row_in && y_in) ? res.y : 0.0;
                }
                w[S][k & 1] = x;
                if (ERR) {
                    // red = (row + column) even: exactly one of the lane's two columns (:609-611)
                    const double v = rho_odd ? res.y : res.x;     // cx is even
                    const bool take = col_own && rho >= own_r_lo && rho < own_r_hi;
                    err_acc = __dadd_rn(err_acc, take ? fabs(v) : 0.0);   // + 0.0 is exact
                }
                if (RES) {
                    const double2 d_cur = make_double2(-res.x, -res.y);   // D = -D (:277-280)
                    const int f_row = rho - 1;                            // lower fine row of the pair (f_row, rho)
                    const double2 ri = ri_cur;                            // {coarse row of f_row or -1, its weight}, staged with the row
                    const int crow = (FAST || (f_row >= 0 && f_row <= N - 1)) ? (int)ri.x : -1;
                    if (crow >= 0 && f_row >= own_r_lo && f_row < own_r_hi) {   // warp-uniform
                        const double cw = ri.y;
                        const bool row_edge = crow == 0 || crow == p.M - 1;
                        const ptrdiff_t ro = (ptrdiff_t)crow * p.M;
                        double *out = p.Fc + ro;
                        // slabs with peer memory: a coarse row a neighbour keeps as its halo also goes straight into its array
                        double *peer_lo = (PEER && p.peer_Fc_lo && crow < p.fc_lo_end) ? p.peer_Fc_lo + ro : nullptr;
                        double *peer_hi = (PEER && p.peer_Fc_hi && crow >= p.fc_hi_begin) ? p.peer_Fc_hi + ro : nullptr;
                        if (ccx >= 0) {
                            const double val = (row_edge || zx) ? 0.0 : restrict_at(d_prev.x, d_prev.y, d_cur.x, d_cur.y, ax, cw);
                            out[ccx] = val;
                            if (PEER && peer_lo) peer_lo[ccx] = val;
                            if (PEER && peer_hi) peer_hi[ccx] = val;
                        }
                        if (!res_even) {                 // (nested ladders: no lane has a coarse point at its odd column)
                            const double np = shfl_dn1(d_prev.x), nc = shfl_dn1(d_cur.x);
                            if (ccy >= 0) {
                                const double val = (row_edge || zy) ? 0.0 : restrict_at(d_prev.y, np, d_cur.y, nc, ay, cw);
                                out[ccy] = val;
                                if (PEER && peer_lo) peer_lo[ccy] = val;
                                if (PEER && peer_hi) peer_hi[ccy] = val;
                            }
                        }
                    }
                    d_prev = d_cur;
                }
            }
            if (IN == IN_PROLONG) prolong_redo(bad);
        }
    };

    for (int rb = r_first; rb <= r_last; rb += U) {
        // interior rows only, and every row the chunk loads (up to DEPTH ahead) is present locally
        const bool fast = strip_fast && rb - NLV >= 1 && rb + U + STREAM_DEPTH <= N - 1 && rb - 1 >= p.row0 &&
                          rb + U + STREAM_DEPTH < p.row0 + p.rows;
        if (fast) chunk(BoolTag<true>(), rb);
        else chunk(BoolTag<false>(), rb);
    }

    cp_async_wait<0>();
    __syncwarp();                                 // the ring is reused by the next task
    if (ERR) {
        double v = err_acc;                       // fixed shuffle tree => the task's partial is deterministic
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
        if (lane == 0) p.partials[task] = v;
        if (MID) {
            double m = mid_acc;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m = __dadd_rn(m, __shfl_down_sync(0xffffffffu, m, off));
            if (lane == 0) p.partials[p.n_tasks + task] = m;
        }
    }
    if (p.trace && lane == 0) {
        unsigned long long t1;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[4 * task + 0] = trace_t0;
        p.trace[4 * task + 1] = t1;
        p.trace[4 * task + 2] = smid;
        p.trace[4 * task + 3] = blockIdx.x * STREAM_WARPS + warp;
    }
  }  // task loop

    finish_launch<STREAM_WARPS, ERR, MID>(p);
}

}  // namespace mg
