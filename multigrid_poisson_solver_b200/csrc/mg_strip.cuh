// mg_strip.cuh -- the fused passes of a cycle on even grids: 4 columns per lane, rows staged by
// per-warp bulk copies (cp.async.bulk + mbarrier), all sweeps in registers.
//
//   level 0 (input)   IN_LOAD     U_in                                   (smoothing pass)
//                     IN_ZERO     0                                      (-1 node: U = 0, MG_solver_CPU.cpp:252-257)
//                     IN_PROLONG  U_f + doProlongation(U_c)              (1 node: :354 + :368)
//   levels 1..S       S Jacobi sweeps (doSmoothing, :578-601), all in registers
//   residual stage    ERR: the red-parity error sum of doSmoothing (:607-622)
//                     RES: F_c = doRestriction(-(getResidual(U_S, F)))   (:268, :277-280, :287)
//
// Work decomposition (as in mg_stream.cuh): a WARP owns a strip of W columns x H rows and streams
// down the rows of a 128-column window, 4 ADJACENT columns per lane and TWO rows per step; a step takes
// in rows r, r+1 of level 0 and produces rows r-1, r of level 1 ... rows r-S, r-S+1 of level S and residual
// rows r-S-1, r-S, keeping two rows per level in registers (8 independent fp64 chains per stage: the warp
// hides its own arithmetic latency, which the register budget of 4 columns -- 8 warps per SM -- requires).  Left/right neighbours come from the adjacent lanes by shuffle; warps
// are independent (halo recomputed: 128/W in x, (H+2S+3)/H in y); tasks (strip, row segment) are
// pulled from an atomic queue by persistent warps.
//
// What is new against the 2-column kernel (ncu, round 1: its -1 / 1 nodes were ISSUE bound, 59-69 %
// issue-slot utilisation at 0.65 of the HBM peak, with 2.6-3.1 non-fp64 instructions per point):
//  * 4 columns per lane for EVERY pass type, so shuffles, shared-memory loads, predicates and
//    address arithmetic are paid once per 4 points instead of once per 2.
//  * Rows travel global -> shared as ONE bulk copy per array and row, issued by one elected lane and
//    tracked by an mbarrier per ring slot (UBLKCP + SYNCS in SASS): no per-lane LDGSTS, no per-lane
//    64-bit address arithmetic, no commit / wait groups.  Rows of an even-sized grid are 16-byte
//    aligned, which is all a 1-D bulk copy needs (no tensor map).
//  * IN_ZERO: the first sweep of a zeroed grid needs no stencil -- level 1 is 0 + 0.25*((0 - 0) - h2 F)
//    evaluated with the same roundings -- so its four additions, its shuffles and the level-0 window
//    disappear.
//  * IN_PROLONG: coarse rows are staged ONCE per coarse row (not once per fine row), interpolated in
//    x when they arrive and kept as two register rows (bottom / top of the current cell); the row's
//    table entry {weights, cell} rides in the row's ring slot as a third bulk copy.
//  * RES: nested ladders (every lane's coarse points sit at its columns 0 and 2) take a path without
//    shuffles and with half the predicated bilinear evaluations.
//  * Row slabs: the rows a neighbouring GPU keeps as its halo are stored straight into its memory
//    (peer pointers) by the warps that produce them; the launch's last CTA publishes a flag.
//
// Bit parity: every expression goes through mg_device.cuh / the *_fast forms of mg_stream.cuh, which
// are value-identical to the reference's association (see mg_stream.cuh).
#pragma once
#include "mg_stream.cuh"

namespace mg {

template <int S, bool NEED_R, bool RES>
struct StripGeo {
    static constexpr int HL = 4;                                                 // left halo >= S+1, a multiple of 4
    static constexpr int HR_NEED = S + (RES ? 2 : NEED_R ? 1 : 0);               // right halo
    static constexpr int W = (124 - HR_NEED) / 4 * 4;                            // owned columns per strip
    static constexpr int ROW_LEAD = S + (NEED_R ? 1 : 0);                        // rows streamed before the first owned row
    static constexpr int ROW_TAIL = S + (RES ? 2 : NEED_R ? 1 : 0);              // steps after the last owned row
};

constexpr int SP_WARPS = 4;            // warps per CTA
constexpr int SP_SLOT = 2048;          // one row of the ring: [U row segment 1 KiB | F row segment 1 KiB]; 4 rows = 2 pair slots in flight
// CTAs per SM (register budget 65536 / (128 * CTAs)): tuned on B200, see DESIGN.md
#ifndef MG_SP_CTAS_PLAIN
#define MG_SP_CTAS_PLAIN 2
#endif
#ifndef MG_SP_CTAS_RES
#define MG_SP_CTAS_RES 2
#endif
#ifndef MG_SP_CTAS_PROLONG
#define MG_SP_CTAS_PROLONG 2
#endif
__host__ __device__ constexpr int strip_min_ctas(int in, bool res)
{
    return in == 2 ? MG_SP_CTAS_PROLONG : res ? MG_SP_CTAS_RES : MG_SP_CTAS_PLAIN;
}
// Shared memory of one warp:
//   [0, 8192)          ring of SP_DEPTH slots
//   [8192, 8224)       one mbarrier per slot              (+32 pad)
//   [8256, 8384)       IN_PROLONG: row-table entry of the row in each slot (32 B each)
//   IN_PROLONG:        4 raw coarse row segments of 1 KiB, then per lane {4 x col_w, 4 x byte offset} (96 B)
//   RES:               per lane {4 x coarse column, 4 x weight} (48 B)
__host__ __device__ constexpr int strip_warp_bytes(int in, bool res)
{
    return 8192 + 64 + 128 + (in == 2 ? 4096 + 32 * 96 : 0) + (res ? 32 * 48 : 0);
}
__host__ __device__ constexpr int strip_smem_bytes(int in, bool res) { return SP_WARPS * strip_warp_bytes(in, res); }

struct d4 { double v[4]; };
#ifdef MG_SP_DEBUG
__device__ unsigned long long g_strip_dbg[64];
#endif

__device__ __forceinline__ d4 lds4(unsigned addr)
{
    d4 r;
    const double2 a = lds2(addr), b = lds2(addr + 16);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    return r;
}

// ---- mbarrier + bulk copy (PTX ISA: mbarrier.*, cp.async.bulk)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// one arrival that also announces `bytes` of asynchronous copies (bytes == 0: a plain arrival)
__device__ __forceinline__ void mbar_arrive_expect(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// `bytes` (a multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory; completes on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

template <int S, int IN, bool ERR, bool RES>
__global__ void __launch_bounds__(SP_WARPS * 32, strip_min_ctas(IN, RES)) k_strip(const StreamParams p)
{
    constexpr bool NEED_R = ERR || RES;
    using G = StripGeo<S, NEED_R, RES>;
    constexpr int NLV = S + (NEED_R ? 1 : 0);    // levels that keep a two-row window (level t feeds stage t)
    constexpr int NF = NLV;                      // F rows a step needs: rows r0 ... r0-NF (r0 = the older of the two new rows)
    constexpr int WB = strip_warp_bytes(IN, RES);
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;
    const double h2 = p.h2, inv_h2 = p.inv_h2;
    double *__restrict__ Op = p.Uout;
    const ptrdiff_t ldn = N;

    extern __shared__ __align__(128) unsigned char strip_smem[];
    const unsigned wbase = (unsigned)__cvta_generic_to_shared(strip_smem) + warp * WB;
    const unsigned rd = wbase + lane * 32;               // this lane's 4 doubles inside a 1 KiB row segment
    const unsigned mbar = wbase + 8192;
    const unsigned rinfo = wbase + 8192 + 64;            // IN_PROLONG: table entries of the 4 rows in the ring (32 B each)
    const unsigned raw = wbase + 8192 + 64 + 128;        // IN_PROLONG: raw coarse rows, slot = coarse row & 3
    const unsigned ctab = raw + 4096 + lane * 96;        // IN_PROLONG: this lane's column weights / offsets
    const unsigned rtab = wbase + 8192 + 64 + 128 + (IN == IN_PROLONG ? 4096 + 32 * 96 : 0) + lane * 48;   // RES

    if (lane == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;                                  // parity of the phase the next chunk waits for (both pair slots alike)

  for (;;) {
    int task = 0;
    if (lane == 0) task = (int)atomicAdd(p.counter, 1u);
    task = __shfl_sync(FULL, task, 0);
    if (task >= p.n_tasks) break;
    // the previous task's last reads of the ring (generic proxy) precede this task's copies (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int seg = task / p.n_strips;                   // consecutive tasks = adjacent strips of one row segment
    const int strip = task - seg * p.n_strips;

    const int own_c_lo = strip * G::W, own_c_hi = min(own_c_lo + G::W, N);
    const int c_first = own_c_lo - G::HL;                // first column of the 128-wide window (a multiple of 4)
    const int cx = c_first + 4 * lane;                   // this lane's columns cx .. cx+3
    bool in_col[4], ok_col[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        ok_col[q] = cx + q >= 0 && cx + q < N;
        in_col[q] = cx + q > 0 && cx + q < N - 1;
    }
    const bool own01 = ok_col[0] && cx >= own_c_lo && cx < own_c_hi;            // ownership per 16-byte pair (N is even)
    const bool own23 = ok_col[2] && cx + 2 >= own_c_lo && cx + 2 < own_c_hi;
    const bool strip_fast = c_first >= 1 && c_first + 127 <= N - 2;             // every column of the window is interior
    const int2 seg_rows = __ldg(p.segs + seg);           // rows of the segment relative to own_lo
    const int own_r_lo = p.own_lo + seg_rows.x, own_r_hi = p.own_lo + seg_rows.y;   // all row indices are GLOBAL
    const int r_first = max(0, own_r_lo - G::ROW_LEAD);
    const int r_last = min(own_r_hi - 1 + G::ROW_TAIL, N - 1 + G::ROW_LEAD);
    const int r_end = r_first + ((r_last - r_first) / 4 + 1) * 4;               // rows [r_first, r_end) are streamed, 4 per chunk
    // the part of a row this strip copies: columns [cs, ce) (even bounds: 16-byte aligned, size a multiple of 16)
    const int cs = max(c_first, 0), ce = min(c_first + 128, N);
    const unsigned cbytes = (unsigned)(ce - cs) * 8u, cdst = (unsigned)(cs - c_first) * 8u;
    // rows of this task whose output is also a neighbour's halo (slabs with peer memory)
    const bool peer_rows = (p.peer_U_lo && own_r_lo < p.u_lo_end) || (p.peer_U_hi && own_r_hi > p.u_hi_begin);

    // Register state: two rows per level, the last two F row pairs.
    d4 w[NLV > 0 ? NLV : 1][2], fp[2][2];
#pragma unroll
    for (int t = 0; t < (NLV > 0 ? NLV : 1); ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) w[t][0].v[q] = w[t][1].v[q] = 0.0;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) fp[t][0].v[q] = fp[t][1].v[q] = 0.0;

    // ---- restriction state: per column {coarse column or -1, weight} in shared memory, previous D row in registers
    d4 d_prev;
#pragma unroll
    for (int q = 0; q < 4; ++q) d_prev.v[q] = 0.0;
    double2 ri_a = make_double2(-1.0, 0.0), ri_b = ri_a;   // {coarse row or -1, weight} of the two fine rows the NEXT step pairs up
    bool res_even = false;                               // every lane's coarse points sit at its columns 0 and 2 (nested ladders)
    auto row_info_of = [&](int f) { return (f >= 0 && f <= N - 1) ? p.rrow[f] : make_double2(-1.0, 0.0); };
    if (RES) {
        bool even = true;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int cc = -1;
            double a = 0.0;
            const bool own = q < 2 ? own01 : own23;
            if (own) {
                cc = p.f2c[cx + q];
                if (cc >= 0) a = p.rw[cc];
            }
            if ((q & 1) && cc >= 0) even = false;
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(rtab + 4 * q), "r"(cc) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(rtab + 16 + 8 * q), "d"(a) : "memory");
        }
        res_even = __all_sync(FULL, even);
        ri_a = row_info_of(r_first - S - 2);             // first step: residual rows r_first-S-1, r_first-S pair with the rows below them
        ri_b = row_info_of(r_first - S - 1);
    }
    double err_acc = 0.0;

    // ---- row copies: one elected lane, one mbarrier phase per PAIR of rows and chunk.  Rows [r_first, r_end) are issued
    // exactly once and in order (two pairs in the prologue, then one pair per step), and every pair is waited for once.
    int r_issue = r_first;
    const double *gU = (IN != IN_ZERO ? p.Uin : p.F) + (ptrdiff_t)r_first * ldn + cs;
    const double *gF = p.F + (ptrdiff_t)(r_first - 1) * ldn + cs;
    auto issue = [&](int d, bool guarded) {              // pair slot d <- rows r_issue, r_issue + 1 (and F rows r_issue - 1, r_issue)
        if (r_issue < r_end && lane == 0) {
            const unsigned bar = mbar + 8 * d, dst = wbase + d * (2 * SP_SLOT) + cdst;
            bool u_ok[2], f_ok[2], i_ok[2];
            unsigned bytes = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = r_issue + j;
                u_ok[j] = IN != IN_ZERO;
                f_ok[j] = NF > 0;
                i_ok[j] = IN == IN_PROLONG;
                if (guarded) {                           // rows outside the grid / the local slab are not copied (never used)
                    u_ok[j] = u_ok[j] && r >= p.row0 && r < p.row0 + p.rows;
                    f_ok[j] = f_ok[j] && r - 1 >= p.row0 && r - 1 < p.row0 + p.rows;
                    i_ok[j] = i_ok[j] && r <= N - 1;
                }
                bytes += (u_ok[j] ? cbytes : 0u) + (f_ok[j] ? cbytes : 0u) + (i_ok[j] ? 32u : 0u);
            }
            mbar_arrive_expect(bar, bytes);              // bytes == 0: a plain arrival
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (u_ok[j]) bulk_g2s(dst + j * SP_SLOT, gU + j * ldn, cbytes, bar);
                if (f_ok[j]) bulk_g2s(dst + j * SP_SLOT + 1024, gF + j * ldn, cbytes, bar);
            }
            if (IN == IN_PROLONG) {
                if (i_ok[0] && i_ok[1]) bulk_g2s(rinfo + 64 * d, p.row_info + r_issue, 64u, bar);
                else if (i_ok[0]) bulk_g2s(rinfo + 64 * d, p.row_info + r_issue, 32u, bar);
            }
        }
        r_issue += 2;
        gU += 2 * ldn;
        gF += 2 * ldn;
    };
    issue(0, true);
    issue(1, true);

    // ---- prolongation state: the current cell's bottom / top coarse rows, interpolated in x for this lane's 4 columns
    d4 bot, top;
#pragma unroll
    for (int q = 0; q < 4; ++q) bot.v[q] = top.v[q] = 0.0;
    int cell = 0, cbase = 0, cw = 0;
    // coarse row q of the window (doubles [cbase, cbase+cw)) -> raw slot q & 3; zero-filled outside the local coarse rows
    auto request_raw = [&](int q) {
        const bool row_ok = q >= p.uc_row0 && q < p.uc_row0 + p.uc_rows;
        const double *src = p.Uc + (ptrdiff_t)q * p.Nc + cbase + lane;
        const unsigned dst = raw + ((unsigned)(q & 3) << 10) + lane * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (32 * j < cw) {                           // warp-uniform
                const bool ok = row_ok && lane + 32 * j < cw;
                cp_async8(dst + 256 * j, ok ? (const void *)(src + 32 * j) : (const void *)p.F_valid, ok);
            }
        cp_async_commit();
    };
    // x-interpolation of raw coarse row q for this lane's columns: c[cell] * w.x + c[cell+1] * w.y (:700, bottom / top)
    auto interp_raw = [&](int q) {
        const unsigned slot = raw + ((unsigned)(q & 3) << 10);
        int4 off;
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(off.x), "=r"(off.y), "=r"(off.z), "=r"(off.w) : "r"(ctab + 64) : "memory");
        const int o[4] = {off.x, off.y, off.z, off.w};
        d4 out;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const double2 wc = lds2(ctab + 16 * q4);
            out.v[q4] = __dadd_rn(__dmul_rn(lds1(slot + o[q4]), wc.x), __dmul_rn(lds1(slot + o[q4] + 8), wc.y));
        }
        return out;
    };
    if (IN == IN_PROLONG) {
        int cq[4];
        int lo = 0x7fffffff, hi = -1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cq[q] = ok_col[q] ? p.col_cell[cx + q] : -1;
            if (ok_col[q]) { lo = min(lo, cq[q]); hi = max(hi, cq[q] + 2); }
        }
        cbase = __reduce_min_sync(FULL, lo);
        cw = min(__reduce_max_sync(FULL, hi), p.Nc) - cbase;               // doubles of a coarse row this window touches (<= 128)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double2 wc = ok_col[q] ? p.col_w[cx + q] : make_double2(0.0, 0.0);
            sts2(ctab + 16 * q, wc);
            const int off = ok_col[q] ? (cq[q] - cbase) * 8 : 0;          // off-grid lanes read element 0 (unused)
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(ctab + 64 + 4 * q), "r"(off) : "memory");
        }
        cell = __ldg(p.row_cell + min(r_first, N - 1));
#pragma unroll
        for (int j = 0; j < 4; ++j) request_raw(cell + j);
        cp_async_wait<2>();                              // rows cell and cell+1 have landed
        __syncwarp();
        bot = interp_raw(cell);
        top = interp_raw(cell + 1);
        __syncwarp();                                    // every lane has read raw row `cell`: its slot takes row cell+4
        request_raw(cell + 4);                           // in flight from here on: rows cell+2, cell+3, cell+4
    }
    // the cell of fine row (table entry at `ri`) becomes current: bot / top move up one coarse row when it changed
    auto enter_cell = [&](int rq) {
        if (rq != cell) {                                // warp-uniform
            cp_async_wait<2>();                          // raw row cell+2 (the oldest of the three in flight) has landed
            __syncwarp();
            bot = top;
            top = interp_raw(cell + 2);
            request_raw(cell + 5);                       // into the slot of row cell+1 (read one move ago)
            cell = rq;
        }
    };

    // One chunk = 4 rows = 2 steps of 2 rows.  A step takes in level-0 rows r0, r0+1 and turns, stage by stage, the two new
    // rows of level t (plus the two rows of its window) into two new rows of level t+1: 8 independent fp64 chains per stage.
    // FAST: every row touched by every stage is an interior row present in the local arrays, every column of the window is an
    // interior column: the body carries no boundary selects.
    auto chunk = [&](auto fast_tag, const int rb) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int r0 = rb + 2 * d;                   // new level-0 rows r0, r0+1
            const unsigned slot = rd + d * (2 * SP_SLOT);
            mbar_wait(mbar + 8 * d, phase);              // rows r0, r0+1 (and F rows r0-1, r0, and the rows' table entries) have landed
            d4 X0, X1, Fa, Fb;
#pragma unroll
            for (int q = 0; q < 4; ++q) X0.v[q] = X1.v[q] = Fa.v[q] = Fb.v[q] = 0.0;
            if (IN == IN_LOAD) { X0 = lds4(slot); X1 = lds4(slot + SP_SLOT); }
            if (NF > 0) { Fa = lds4(slot + 1024); Fb = lds4(slot + SP_SLOT + 1024); }
            if (IN == IN_PROLONG) {
                // level 0 of the 1 node: U_f + P(U_c) (:700 + :569), rows r0 and r0+1
                d4 uf[2], v[2];
                bool bad = false;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) uf[j].v[q] = v[j].v[q] = 0.0;
                    if (FAST || r0 + j <= N - 1) {       // (general body: rows streamed past the grid carry no data)
                        uf[j] = lds4(slot + j * SP_SLOT);
                        const double2 wr = lds2(rinfo + 64 * d + 32 * j);           // {c3y - f_y, f_y - c1y}
                        enter_cell((int)lds1(rinfo + 64 * d + 32 * j + 16));        // the row's cell (exact in a double)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            v[j].v[q] = __dadd_rn(__dmul_rn(bot.v[q], wr.x), __dmul_rn(top.v[q], wr.y));
                            bad = bad || ((FAST || ok_col[q]) && div2_unsafe(v[j].v[q]));
                        }
                    }
                }
                const double dd = p.c_dx, y = p.inv_c_dx;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    X0.v[q] = __dadd_rn(uf[0].v[q], div_fast(div_fast(v[0].v[q], dd, y), dd, y));
                    X1.v[q] = __dadd_rn(uf[1].v[q], div_fast(div_fast(v[1].v[q], dd, y), dd, y));
                }
                if (__any_sync(FULL, bad)) {             // rare: IEEE divisions for the whole warp (same values wherever the fast path is valid)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        X0.v[q] = __dadd_rn(uf[0].v[q], __ddiv_rn(__ddiv_rn(v[0].v[q], dd), dd));
                        X1.v[q] = __dadd_rn(uf[1].v[q], __ddiv_rn(__ddiv_rn(v[1].v[q], dd), dd));
                    }
                }
            }
            // Every lane has read pair slot d: refill it with rows r0+4, r0+5.  The reads went through the generic proxy, the
            // copy writes through the async proxy: without the proxy fence the copy may overtake them (measured on B200: a
            // plain S = 1 pass at N = 2048 handed out rows of the NEXT occupant in 20 of 25 runs).
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            issue(d, !FAST);

            // F rows in registers: the pair that just arrived, the pair before it, and the last row of the pair before that
            const d4 fold = fp[d][1];                    // F row r0-4
            fp[d][0] = Fa;                               // F row r0-1
            fp[d][1] = Fb;                               // F row r0
            auto Frow = [&](int back) -> const d4 & {    // F row r0 - back, back = 0 .. 4 (compile-time)
                return back == 0 ? fp[d][1] : back == 1 ? fp[d][0] : back == 2 ? fp[d ^ 1][1] : back == 3 ? fp[d ^ 1][0] : fold;
            };

            // ---- S sweeps: stage t turns level t rows (r0-t-1, r0-t) into level t+1
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const int i0 = r0 - t - 1;               // output rows i0, i0+1
                const d4 &f0 = Frow(t + 1), &f1 = Frow(t);
                d4 Y0, Y1;
                if (IN == IN_ZERO && t == 0) {
                    // level 0 is all zeros: jacobi_at(0, 0, h2 f) = 0 + 0.25*((0 - 0) - h2 f), same roundings, no stencil
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        Y0.v[q] = __dadd_rn(0.0, __dmul_rn(0.25, __dsub_rn(0.0, __dmul_rn(h2, f0.v[q]))));
                        Y1.v[q] = __dadd_rn(0.0, __dmul_rn(0.25, __dsub_rn(0.0, __dmul_rn(h2, f1.v[q]))));
                    }
                    if (!FAST) {
                        const bool in0 = i0 > 0 && i0 < N - 1, in1 = i0 + 1 > 0 && i0 + 1 < N - 1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            Y0.v[q] = (in0 && in_col[q]) ? Y0.v[q] : 0.0;
                            Y1.v[q] = (in1 && in_col[q]) ? Y1.v[q] : 0.0;
                        }
                    }
                } else {
                    const d4 a = w[t][0], b = w[t][1];   // rows i0-1, i0; the new rows X0 = i0+1, X1 = i0+2
                    const double lb = shfl_up1(b.v[3]), rgb = shfl_dn1(b.v[0]), lc = shfl_up1(X0.v[3]), rgc = shfl_dn1(X0.v[0]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double l0 = q == 0 ? lb : b.v[q - 1], g0 = q == 3 ? rgb : b.v[q + 1];
                        const double l1 = q == 0 ? lc : X0.v[q - 1], g1 = q == 3 ? rgc : X0.v[q + 1];
                        Y0.v[q] = jacobi_fast(b.v[q], sum4(X0.v[q], a.v[q], g0, l0), __dmul_rn(h2, f0.v[q]));
                        Y1.v[q] = jacobi_fast(X0.v[q], sum4(X1.v[q], b.v[q], g1, l1), __dmul_rn(h2, f1.v[q]));
                    }
                    if (!FAST) {                         // boundary rows / columns are carried over
                        const bool in0 = i0 > 0 && i0 < N - 1, in1 = i0 + 1 > 0 && i0 + 1 < N - 1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            Y0.v[q] = (in0 && in_col[q]) ? Y0.v[q] : b.v[q];
                            Y1.v[q] = (in1 && in_col[q]) ? Y1.v[q] : X0.v[q];
                        }
                    }
                    w[t][0] = X0;
                    w[t][1] = X1;
                }
                X0 = Y0;
                X1 = Y1;
            }

            // ---- X0, X1 are now level S, rows r0-S, r0-S+1
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = r0 - S + j;
                const d4 &x = j == 0 ? X0 : X1;
                if (Op && i >= own_r_lo && i < own_r_hi) {
                    const ptrdiff_t o = (ptrdiff_t)i * ldn + cx;
                    if (own01) *reinterpret_cast<double2 *>(Op + o) = make_double2(x.v[0], x.v[1]);
                    if (own23) *reinterpret_cast<double2 *>(Op + o + 2) = make_double2(x.v[2], x.v[3]);
                    if (peer_rows) {                     // the same row into the neighbour's halo
                        if (p.peer_U_lo && i < p.u_lo_end) {
                            if (own01) *reinterpret_cast<double2 *>(p.peer_U_lo + o) = make_double2(x.v[0], x.v[1]);
                            if (own23) *reinterpret_cast<double2 *>(p.peer_U_lo + o + 2) = make_double2(x.v[2], x.v[3]);
                        }
                        if (p.peer_U_hi && i >= p.u_hi_begin) {
                            if (own01) *reinterpret_cast<double2 *>(p.peer_U_hi + o) = make_double2(x.v[0], x.v[1]);
                            if (own23) *reinterpret_cast<double2 *>(p.peer_U_hi + o + 2) = make_double2(x.v[2], x.v[3]);
                        }
                    }
                }
            }

            if (NEED_R) {
                const int rho0 = r0 - S - 1;             // residual rows rho0, rho0+1
                const d4 a = w[S][0], b = w[S][1];       // rows rho0-1, rho0; the new rows X0 = rho0+1, X1 = rho0+2
                const d4 &f0 = Frow(S + 1), &f1 = Frow(S);
                const double lb = shfl_up1(b.v[3]), rgb = shfl_dn1(b.v[0]), lc = shfl_up1(X0.v[3]), rgc = shfl_dn1(X0.v[0]);
                d4 R0, R1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double l0 = q == 0 ? lb : b.v[q - 1], g0 = q == 3 ? rgb : b.v[q + 1];
                    const double l1 = q == 0 ? lc : X0.v[q - 1], g1 = q == 3 ? rgc : X0.v[q + 1];
                    R0.v[q] = residual_fast(b.v[q], sum4(X0.v[q], a.v[q], g0, l0), f0.v[q], inv_h2);
                    R1.v[q] = residual_fast(X0.v[q], sum4(X1.v[q], b.v[q], g1, l1), f1.v[q], inv_h2);
                }
                if (!FAST) {                             // 0 on the boundary (:559)
                    const bool in0 = rho0 > 0 && rho0 < N - 1, in1 = rho0 + 1 > 0 && rho0 + 1 < N - 1;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        R0.v[q] = (in0 && in_col[q]) ? R0.v[q] : 0.0;
                        R1.v[q] = (in1 && in_col[q]) ? R1.v[q] : 0.0;
                    }
                }
                w[S][0] = X0;
                w[S][1] = X1;
                if (ERR) {
                    // red = (row + column) even; cx is even: columns 0, 2 on even rows, 1, 3 on odd rows (:609-611)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int rho = rho0 + j;
                        const d4 &res = j == 0 ? R0 : R1;
                        const bool odd = rho & 1;
                        const double v0 = odd ? res.v[1] : res.v[0], v1 = odd ? res.v[3] : res.v[2];
                        const bool row_own = rho >= own_r_lo && rho < own_r_hi;
                        err_acc = __dadd_rn(err_acc, (row_own && own01) ? fabs(v0) : 0.0);   // + 0.0 is exact
                        err_acc = __dadd_rn(err_acc, (row_own && own23) ? fabs(v1) : 0.0);
                    }
                }
                if (RES) {
                    d4 D0, D1;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { D0.v[q] = -R0.v[q]; D1.v[q] = -R1.v[q]; }   // D = -D (:277-280)
                    // coarse row of the fine row pair (f_row, f_row + 1) = (dp, dc), if the floor map has one there
                    auto emit_pair = [&](const d4 &dp, const d4 &dc, const double2 ri, const int f_row) {
                        const int crow = (int)ri.x;
                        if (crow >= 0 && f_row >= own_r_lo && f_row < own_r_hi) {   // warp-uniform
                            const double cwt = ri.y;
                            const bool row_edge = crow == 0 || crow == p.M - 1;
                            const ptrdiff_t ro = (ptrdiff_t)crow * p.M;
                            double *peer_lo = (p.peer_Fc_lo && crow < p.fc_lo_end) ? p.peer_Fc_lo + ro : nullptr;
                            double *peer_hi = (p.peer_Fc_hi && crow >= p.fc_hi_begin) ? p.peer_Fc_hi + ro : nullptr;
                            double *out = p.Fc + ro;
                            int4 cc;
                            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(cc.x), "=r"(cc.y), "=r"(cc.z), "=r"(cc.w) : "r"(rtab) : "memory");
                            const d4 aw = lds4(rtab + 16);
                            const int ccq[4] = {cc.x, cc.y, cc.z, cc.w};
                            auto emit = [&](int q, double p1, double c1) {
                                if (ccq[q] >= 0) {
                                    const bool edge = row_edge || ccq[q] == 0 || ccq[q] == p.M - 1;
                                    const double val = edge ? 0.0 : restrict_at(dp.v[q], p1, dc.v[q], c1, aw.v[q], cwt);
                                    out[ccq[q]] = val;
                                    if (peer_lo) peer_lo[ccq[q]] = val;
                                    if (peer_hi) peer_hi[ccq[q]] = val;
                                }
                            };
                            if (res_even) {              // nested ladder: coarse points at this lane's columns 0 and 2 only
                                emit(0, dp.v[1], dc.v[1]);
                                emit(2, dp.v[3], dc.v[3]);
                            } else {
                                const double np = shfl_dn1(dp.v[0]), nc = shfl_dn1(dc.v[0]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) emit(q, q == 3 ? np : dp.v[q + 1], q == 3 ? nc : dc.v[q + 1]);
                            }
                        }
                    };
                    const double2 ra = ri_a, rb2 = ri_b;
                    if (FAST) { ri_a = p.rrow[rho0 + 1]; ri_b = p.rrow[rho0 + 2]; }
                    else { ri_a = row_info_of(rho0 + 1); ri_b = row_info_of(rho0 + 2); }
                    emit_pair(d_prev, D0, ra, rho0 - 1);
                    emit_pair(D0, D1, rb2, rho0);
                    d_prev = D1;
                }
            }
        }
        phase ^= 1u;
    };

    for (int rb = r_first; rb <= r_last; rb += 4) {
        // interior rows only, and every row the chunk copies (4 ahead) is present locally
        const bool fast = strip_fast && rb - NLV >= 1 && rb + 8 <= N - 1 && rb - 1 >= p.row0 && rb + 8 < p.row0 + p.rows;
        if (fast) chunk(BoolTag<true>(), rb);
        else chunk(BoolTag<false>(), rb);
    }

    if (IN == IN_PROLONG) cp_async_wait<0>();            // the raw rows requested ahead of the last cell
    if (ERR) {
        double v = err_acc;                              // fixed shuffle tree => the task's partial is deterministic
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(FULL, v, off));
        if (lane == 0) p.partials[task] = v;
    }
  }  // task loop

    finish_launch<SP_WARPS, ERR>(p);
}

}  // namespace mg
