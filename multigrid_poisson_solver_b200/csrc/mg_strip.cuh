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
// down the rows of a 128-column window, 4 ADJACENT columns per lane; at step r it takes in row r of
// level 0 and produces row r-1 of level 1 ... row r-S of level S and residual row r-S-1, keeping two
// rows per level in registers.  Left/right neighbours come from the adjacent lanes by shuffle; warps
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
constexpr int SP_DEPTH = 4;            // rows in flight per warp = ring slots = unroll factor
constexpr int SP_SLOT = 2048;          // ring slot: [U row segment 1 KiB | F row segment 1 KiB]
// CTAs per SM (register budget 65536 / (128 * CTAs)): tuned on B200, see DESIGN.md
#ifndef MG_SP_CTAS_PLAIN
#define MG_SP_CTAS_PLAIN 3
#endif
#ifndef MG_SP_CTAS_RES
#define MG_SP_CTAS_RES 3
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
    constexpr int NF = NLV;                      // F rows alive at once: rows r-1 ... r-NF
    constexpr int NR = 4;                        // F ring (registers); also the unroll factor
    constexpr int U = SP_DEPTH;
    constexpr int WB = strip_warp_bytes(IN, RES);
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(U == NR && U == 4, "slot, F ring and row-window indices are compile-time constants of the unrolled chunk");

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;
    const double h2 = p.h2, inv_h2 = p.inv_h2;
    double *__restrict__ Op = p.Uout;
    const ptrdiff_t ldn = N;

    extern __shared__ __align__(128) unsigned char strip_smem[];
    const unsigned wbase = (unsigned)__cvta_generic_to_shared(strip_smem) + warp * WB;
    const unsigned rd = wbase + lane * 32;               // this lane's 4 doubles inside a 1 KiB row segment
    const unsigned mbar = wbase + 8192;
    const unsigned rinfo = wbase + 8192 + 64;            // IN_PROLONG
    const unsigned raw = wbase + 8192 + 64 + 128;        // IN_PROLONG: raw coarse rows, slot = coarse row & 3
    const unsigned ctab = raw + 4096 + lane * 96;        // IN_PROLONG: this lane's column weights / offsets
    const unsigned rtab = wbase + 8192 + 64 + 128 + (IN == IN_PROLONG ? 4096 + 32 * 96 : 0) + lane * 48;   // RES

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < SP_DEPTH; ++k) mbar_init(mbar + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;                                  // parity of the phase the next chunk waits for (all slots alike)

  for (;;) {
    int task = 0;
    if (lane == 0) task = (int)atomicAdd(p.counter, 1u);
    task = __shfl_sync(FULL, task, 0);
    if (task >= p.n_tasks) break;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();                                        // the previous task's last reads of the ring precede this task's copies
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
    const int r_end = r_first + ((r_last - r_first) / U + 1) * U;               // rows [r_first, r_end) are streamed
    // the part of a row this strip copies: columns [cs, ce) (even bounds: 16-byte aligned, size a multiple of 16)
    const int cs = max(c_first, 0), ce = min(c_first + 128, N);
    const unsigned cbytes = (unsigned)(ce - cs) * 8u, cdst = (unsigned)(cs - c_first) * 8u;
    // rows of this task whose output is also a neighbour's halo (slabs with peer memory)
    const bool peer_rows = (p.peer_U_lo && own_r_lo < p.u_lo_end) || (p.peer_U_hi && own_r_hi > p.u_hi_begin);

    d4 w[NLV > 0 ? NLV : 1][2], fr[NR];
#pragma unroll
    for (int t = 0; t < (NLV > 0 ? NLV : 1); ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) w[t][0].v[q] = w[t][1].v[q] = 0.0;
#pragma unroll
    for (int t = 0; t < NR; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) fr[t].v[q] = 0.0;

    // ---- restriction state: per column {coarse column or -1, weight} in shared memory, previous D row in registers
    d4 d_prev;
#pragma unroll
    for (int q = 0; q < 4; ++q) d_prev.v[q] = 0.0;
    double2 rinfo_next = make_double2(-1.0, 0.0);
    bool res_even = false;                               // every lane's coarse points sit at its columns 0 and 2 (nested ladders)
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
        const int f0 = r_first - S - 2;
        if (f0 >= 0 && f0 <= N - 1) rinfo_next = p.rrow[f0];
    }
    double err_acc = 0.0;

    // ---- row copies: one elected lane, one mbarrier phase per slot and chunk.  Rows [r_first, r_end) are issued
    // exactly once and in order (4 in the prologue, then one per step), and every one is waited for exactly once.
    int r_issue = r_first;
    const double *gU = (IN != IN_ZERO ? p.Uin : p.F) + (ptrdiff_t)r_first * ldn + cs;
    const double *gF = p.F + (ptrdiff_t)(r_first - 1) * ldn + cs;
    auto issue = [&](int k, bool guarded) {
        if (r_issue < r_end && lane == 0) {
            const unsigned bar = mbar + 8 * k, dst = wbase + k * SP_SLOT + cdst;
            bool u_ok = IN != IN_ZERO, f_ok = NF > 0, i_ok = IN == IN_PROLONG;
            if (guarded) {                               // rows outside the grid / the local slab are not copied (never used)
                u_ok = u_ok && r_issue >= p.row0 && r_issue < p.row0 + p.rows;
                f_ok = f_ok && r_issue - 1 >= p.row0 && r_issue - 1 < p.row0 + p.rows;
                i_ok = i_ok && r_issue <= N - 1;
            }
            mbar_arrive_expect(bar, (u_ok ? cbytes : 0u) + (f_ok ? cbytes : 0u) + (i_ok ? 32u : 0u));
            if (u_ok) bulk_g2s(dst, gU, cbytes, bar);
            if (f_ok) bulk_g2s(dst + 1024, gF, cbytes, bar);
            if (IN == IN_PROLONG && i_ok) bulk_g2s(rinfo + 32 * k, p.row_info + r_issue, 32u, bar);
        }
        ++r_issue;
        gU += ldn;
        gF += ldn;
    };
#pragma unroll
    for (int k = 0; k < SP_DEPTH; ++k) issue(k, true);

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

    // One chunk = U consecutive steps.  FAST: every row touched by every stage is an interior row present in the
    // local arrays, every column of the window is an interior column: the body carries no boundary selects.
    auto chunk = [&](auto fast_tag, const int rb) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = rb + k;
            mbar_wait(mbar + 8 * k, phase);              // row r (and F row r-1, and the row's table entry) have landed
            d4 x, f_new;
#pragma unroll
            for (int q = 0; q < 4; ++q) x.v[q] = f_new.v[q] = 0.0;
            if (IN == IN_LOAD) x = lds4(rd + k * SP_SLOT);
            if (NF > 0) f_new = lds4(rd + k * SP_SLOT + 1024);
#ifdef MG_SP_DEBUG
            if (IN == IN_LOAD && FAST) {                 // what did the ring hand out?  compare with global memory
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double g = __ldcg(p.Uin + (ptrdiff_t)r * ldn + cx + q);
                    if (g != x.v[q]) {
                        const double prev = __ldcg(p.Uin + (ptrdiff_t)(r - 4) * ldn + cx + q), next = __ldcg(p.Uin + (ptrdiff_t)(r + 4) * ldn + cx + q);
                        const int kind = x.v[q] == prev ? 1 : x.v[q] == next ? 2 : 3;
                        atomicAdd(&g_strip_dbg[kind], 1ull);
                        __nanosleep(2000);
                        const double again = lds1(rd + k * SP_SLOT + 8 * q);
                        atomicAdd(&g_strip_dbg[again == g ? 4 : again == x.v[q] ? 5 : 6], 1ull);
                        atomicAdd(&g_strip_dbg[8 + (rb == r_first ? 0 : 1)], 1ull);          // first chunk of a task or later
                        atomicAdd(&g_strip_dbg[10 + k], 1ull);
                        if (atomicAdd(&g_strip_dbg[0], 1ull) < 8) {
                            g_strip_dbg[16 + 0] = (unsigned long long)task; g_strip_dbg[16 + 1] = r; g_strip_dbg[16 + 2] = lane; g_strip_dbg[16 + 3] = r_first;
                            g_strip_dbg[16 + 4] = r_end; g_strip_dbg[16 + 5] = phase;
                        }
                    }
                    const double gf = __ldcg(p.F + (ptrdiff_t)(r - 1) * ldn + cx + q);
                    if (gf != f_new.v[q]) atomicAdd(&g_strip_dbg[7], 1ull);
                }
            }
#endif
            if (IN == IN_PROLONG && (FAST || r <= N - 1)) {
                // level 0 of the 1 node: U_f + P(U_c) (:700 + :569)
                const d4 uf = lds4(rd + k * SP_SLOT);
                const double2 wr = lds2(rinfo + 32 * k);                  // {c3y - f_y, f_y - c1y}
                const int rq = (int)lds1(rinfo + 32 * k + 16);            // the row's cell (exact in a double)
                if (rq != cell) {                        // warp-uniform: the cell moved up one coarse row
                    cp_async_wait<2>();                  // raw row cell+2 (the oldest of the three in flight) has landed
                    __syncwarp();
                    bot = top;
                    top = interp_raw(cell + 2);
                    request_raw(cell + 5);               // into the slot of row cell+1 (read one move ago)
                    cell = rq;
                }
                d4 v;
                bool bad = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v.v[q] = __dadd_rn(__dmul_rn(bot.v[q], wr.x), __dmul_rn(top.v[q], wr.y));
                    bad = bad || ((FAST || ok_col[q]) && div2_unsafe(v.v[q]));
                }
                const double d = p.c_dx, y = p.inv_c_dx;
#pragma unroll
                for (int q = 0; q < 4; ++q) x.v[q] = __dadd_rn(uf.v[q], div_fast(div_fast(v.v[q], d, y), d, y));
                if (__any_sync(FULL, bad)) {             // rare: IEEE divisions for the whole warp (same values wherever the fast path is valid)
#pragma unroll
                    for (int q = 0; q < 4; ++q) x.v[q] = __dadd_rn(uf.v[q], __ddiv_rn(__ddiv_rn(v.v[q], d), d));
                }
            }
            // Every lane has read slot k: refill it with row r + DEPTH.  The reads went through the generic proxy, the
            // copy writes through the async proxy: without the proxy fence the copy may overtake them (measured on
            // B200: a plain S = 1 pass at N = 2048 handed out rows of the NEXT occupant in 20 of 25 runs).
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            issue(k, !FAST);

            if (NF > 0) fr[k % NR] = f_new;

            // ---- S sweeps: stage t turns level t row (r-t-1) into level t+1
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const int i = r - t - 1;
                const d4 f = fr[(k - t + 4 * NR) % NR];
                d4 nx;
                if (IN == IN_ZERO && t == 0) {
                    // level 0 is all zeros: jacobi_at(0, 0, h2 f) = 0 + 0.25*((0 - 0) - h2 f), same roundings, no stencil
#pragma unroll
                    for (int q = 0; q < 4; ++q) nx.v[q] = __dadd_rn(0.0, __dmul_rn(0.25, __dsub_rn(0.0, __dmul_rn(h2, f.v[q]))));
                    if (!FAST) {
                        const bool row_in = i > 0 && i < N - 1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) nx.v[q] = (row_in && in_col[q]) ? nx.v[q] : 0.0;
                    }
                } else {
                    const d4 below = w[t][k & 1], c = w[t][(k & 1) ^ 1];
                    const double left = shfl_up1(c.v[3]), right = shfl_dn1(c.v[0]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double l = q == 0 ? left : c.v[q - 1], rr = q == 3 ? right : c.v[q + 1];
                        nx.v[q] = jacobi_fast(c.v[q], sum4(x.v[q], below.v[q], rr, l), __dmul_rn(h2, f.v[q]));
                    }
                    if (!FAST) {                         // boundary rows / columns are carried over
                        const bool row_in = i > 0 && i < N - 1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) nx.v[q] = (row_in && in_col[q]) ? nx.v[q] : c.v[q];
                    }
                    w[t][k & 1] = x;
                }
                x = nx;
            }

            // ---- x is now level S, row r-S
            {
                const int i = r - S;
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
                const int rho = r - S - 1;               // residual row
                const d4 below = w[S][k & 1], c = w[S][(k & 1) ^ 1];
                const d4 f = fr[(k - S + 4 * NR) % NR];
                const double left = shfl_up1(c.v[3]), right = shfl_dn1(c.v[0]);
                d4 res;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double l = q == 0 ? left : c.v[q - 1], rr = q == 3 ? right : c.v[q + 1];
                    res.v[q] = residual_fast(c.v[q], sum4(x.v[q], below.v[q], rr, l), f.v[q], inv_h2);
                }
                if (!FAST) {                             // 0 on the boundary (:559)
                    const bool row_in = rho > 0 && rho < N - 1;
#pragma unroll
                    for (int q = 0; q < 4; ++q) res.v[q] = (row_in && in_col[q]) ? res.v[q] : 0.0;
                }
                w[S][k & 1] = x;
                if (ERR) {
                    // red = (row + column) even; cx is even: columns 0, 2 on even rows, 1, 3 on odd rows (:609-611)
                    const bool odd = rho & 1;
                    const double v0 = odd ? res.v[1] : res.v[0], v1 = odd ? res.v[3] : res.v[2];
                    const bool row_own = rho >= own_r_lo && rho < own_r_hi;
                    err_acc = __dadd_rn(err_acc, (row_own && own01) ? fabs(v0) : 0.0);   // + 0.0 is exact
                    err_acc = __dadd_rn(err_acc, (row_own && own23) ? fabs(v1) : 0.0);
                }
                if (RES) {
                    d4 d_cur;
#pragma unroll
                    for (int q = 0; q < 4; ++q) d_cur.v[q] = -res.v[q];       // D = -D (:277-280)
                    const int f_row = rho - 1;                                // lower fine row of the pair (f_row, rho)
                    const double2 ri = rinfo_next;                            // {coarse row of f_row or -1, its weight}
                    if (FAST || (f_row + 1 >= 0 && f_row + 1 <= N - 1)) rinfo_next = p.rrow[f_row + 1];
                    else rinfo_next = make_double2(-1.0, 0.0);
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
                        const d4 a = lds4(rtab + 16);
                        const int ccq[4] = {cc.x, cc.y, cc.z, cc.w};
                        auto emit = [&](int q, double p1, double c1) {
                            if (ccq[q] >= 0) {
                                const bool edge = row_edge || ccq[q] == 0 || ccq[q] == p.M - 1;
                                const double val = edge ? 0.0 : restrict_at(d_prev.v[q], p1, d_cur.v[q], c1, a.v[q], cwt);
                                out[ccq[q]] = val;
                                if (peer_lo) peer_lo[ccq[q]] = val;
                                if (peer_hi) peer_hi[ccq[q]] = val;
                            }
                        };
                        if (res_even) {                  // nested ladder: coarse points at this lane's columns 0 and 2 only
                            emit(0, d_prev.v[1], d_cur.v[1]);
                            emit(2, d_prev.v[3], d_cur.v[3]);
                        } else {
                            const double np = shfl_dn1(d_prev.v[0]), nc = shfl_dn1(d_cur.v[0]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) emit(q, q == 3 ? np : d_prev.v[q + 1], q == 3 ? nc : d_cur.v[q + 1]);
                        }
                    }
                    d_prev = d_cur;
                }
            }
        }
        phase ^= 1u;
    };

    for (int rb = r_first; rb <= r_last; rb += U) {
        // interior rows only, and every row the chunk copies (up to DEPTH ahead) is present locally
        const bool fast = strip_fast && rb - NLV >= 1 && rb + U + SP_DEPTH <= N - 1 && rb - 1 >= p.row0 &&
                          rb + U + SP_DEPTH < p.row0 + p.rows;
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
