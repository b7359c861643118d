// mg_stream4.cuh -- the 4-columns-per-lane variant of the register-streaming kernel (mg_stream.cuh).
//
// Same pipeline, same arithmetic, same task queue; a warp now owns a 128-column window and every
// lane 4 ADJACENT columns.  Per grid point this halves the warp shuffles (two per stage and lane
// instead of two per stage and column pair), halves the per-step integer / address / predicate
// overhead, doubles the independent fp64 chains per warp (4 instead of 2) and shrinks the
// recomputed column halo from 64/W = 1.14-1.19 to 128/W = 1.07-1.10.  The kernels it replaces
// were issue-bound (ncu: issue 65-68 %, fp64 pipe 47 %), not HBM-bound.
//
// Rows are copied global -> shared with fully coalesced 16-byte cp.async chunks (lane l copies
// chunks l and l+32 of the 1 KiB row segment) and each lane then reads ITS 32 bytes, so a
// __syncwarp separates the copy's completion from the reads and the reads from the slot's refill.
// Used for the smoothing pass (IN_LOAD / IN_ZERO without restriction): 94 % of the measured HBM
// peak for 3 fused sweeps + error.  The -1 node compiles too (and is bit-identical) but needs
// 230-254 registers, which leaves 8 warps per SM and makes it slower than the 2-column kernel;
// the launcher keeps the -1 and 1 nodes on k_stream.
#pragma once
#include "mg_stream.cuh"

namespace mg {

template <int S, bool NEED_R, bool RES>
struct Stream4Geo {
    static constexpr int HL = 4;                                                 // >= S+1, multiple of 4
    static constexpr int HR_NEED = S + (RES ? 2 : NEED_R ? 1 : 0);
    static constexpr int W = (124 - HR_NEED) / 4 * 4;                            // owned columns per strip
    static constexpr int ROW_LEAD = S + (NEED_R ? 1 : 0);
    static constexpr int ROW_TAIL = S + (RES ? 2 : NEED_R ? 1 : 0);
};

constexpr int S4_WARPS = 4, S4_MIN_CTAS = 2;
constexpr int S4_SLOT = 2048;                                  // [U 1 KiB | F 1 KiB]
__host__ __device__ constexpr int stream4_smem_bytes() { return S4_WARPS * (STREAM_DEPTH * S4_SLOT + 32 * 48); }

struct d4 { double v[4]; };

__device__ __forceinline__ d4 lds4(unsigned addr)
{
    d4 r;
    const double2 a = lds2(addr), b = lds2(addr + 16);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    return r;
}

template <int S, int IN, bool ERR, bool RES>
__global__ void __launch_bounds__(S4_WARPS * 32, S4_MIN_CTAS) k_stream4(const StreamParams p)
{
    constexpr bool NEED_R = ERR || RES;
    using G = Stream4Geo<S, NEED_R, RES>;
    constexpr int NLV = S + (NEED_R ? 1 : 0);
    constexpr int NF = NLV;
    constexpr int NR = NF <= 2 ? 2 : 4;
    constexpr int U = NR;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;
    const double h2 = p.h2, inv_h2 = p.inv_h2;
    const double *__restrict__ Fp = p.F;
    const double *__restrict__ Up = p.Uin;
    double *__restrict__ Op = p.Uout;
    const ptrdiff_t ldn = N;

    extern __shared__ __align__(16) unsigned char stream_smem[];
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(stream_smem);
    const unsigned warp_ring = smem0 + warp * (STREAM_DEPTH * S4_SLOT);
    const unsigned rd_base = warp_ring + lane * 32;                 // this lane's 4 doubles inside a 1 KiB part
    const unsigned cc_addr = smem0 + S4_WARPS * (STREAM_DEPTH * S4_SLOT) + warp * (32 * 48) + lane * 48;   // RES: 4 ints + 4 doubles

  for (;;) {
    int task = 0;
    if (lane == 0) task = (int)atomicAdd(p.counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= p.n_tasks) break;
    const int seg_idx = task / p.n_strips;
    const int strip = task - seg_idx * p.n_strips;
    const int seg = seg_idx;                                    // consecutive tasks = adjacent strips of one row segment

    const int own_c_lo = strip * G::W, own_c_hi = min(own_c_lo + G::W, N);
    const int c_first = own_c_lo - G::HL;                           // first column of the 128-wide window (multiple of 4)
    const int cx = c_first + 4 * lane;                              // this lane's columns cx .. cx+3
    bool in_col[4], ok_col[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        ok_col[q] = cx + q >= 0 && cx + q < N;
        in_col[q] = cx + q > 0 && cx + q < N - 1;
    }
    const bool own01 = ok_col[0] && cx >= own_c_lo && cx < own_c_hi;            // ownership per 16-byte pair
    const bool own23 = ok_col[2] && cx + 2 >= own_c_lo && cx + 2 < own_c_hi;
    const bool strip_fast = c_first >= 1 && c_first + 127 <= N - 2;
    const int2 seg_rows = __ldg(p.segs + seg);                  // rows of the segment relative to own_lo
    const int own_r_lo = p.own_lo + seg_rows.x, own_r_hi = p.own_lo + seg_rows.y;
    const int r_first = max(0, own_r_lo - G::ROW_LEAD);
    const int r_last = min(own_r_hi - 1 + G::ROW_TAIL, N - 1 + G::ROW_LEAD);

    d4 w[NLV > 0 ? NLV : 1][2], fr[NR];
#pragma unroll
    for (int t = 0; t < (NLV > 0 ? NLV : 1); ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) w[t][0].v[q] = w[t][1].v[q] = 0.0;
#pragma unroll
    for (int t = 0; t < NR; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) fr[t].v[q] = 0.0;
    unsigned slot_off = 0;

    // ---- restriction state: per column {coarse column or -1, weight} in shared memory, previous D row in registers
    d4 d_prev;
#pragma unroll
    for (int q = 0; q < 4; ++q) d_prev.v[q] = 0.0;
    double2 rinfo_next = make_double2(-1.0, 0.0);
    if (RES) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int cc = -1;
            double a = 0.0;
            const bool own = q < 2 ? own01 : own23;
            if (own) {
                cc = p.f2c[cx + q];
                if (cc >= 0) a = p.rw[cc];
            }
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(cc_addr + 4 * q), "r"(cc) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(cc_addr + 16 + 8 * q), "d"(a) : "memory");
        }
        const int f0 = r_first - S - 2;
        if (f0 >= 0 && f0 <= N - 1) rinfo_next = p.rrow[f0];
    }
    double err_acc = 0.0;

    // cp.async of level-0 row r and F row r-1: lane copies the 16-byte chunks `lane` and `lane + 32`
    auto issue = [&](int r, unsigned off, bool guarded) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ccol = c_first + 2 * (lane + 32 * h);        // first column of the chunk (even)
            const unsigned dst = warp_ring + off + (lane + 32 * h) * 16;
            if (IN != IN_ZERO) {
                const bool ok = !guarded || (ccol >= 0 && ccol < N && r >= p.row0 && r < p.row0 + p.rows);
                cp_async16(dst, ok ? (const void *)(Up + (ptrdiff_t)r * ldn + ccol) : (const void *)p.F_valid, ok);
            }
            if (NF > 0) {
                const bool ok = !guarded || (ccol >= 0 && ccol < N && r - 1 >= p.row0 && r - 1 < p.row0 + p.rows);
                cp_async16(dst + 1024, ok ? (const void *)(Fp + (ptrdiff_t)(r - 1) * ldn + ccol) : (const void *)p.F_valid, ok);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < STREAM_DEPTH; ++d) issue(r_first + d, d * S4_SLOT, true);

    auto chunk = [&](auto fast_tag, const int rb) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = rb + k;
            cp_async_wait<STREAM_DEPTH - 1>();
            __syncwarp();                                           // every lane's chunks of row r have landed
            d4 x, f_new;
#pragma unroll
            for (int q = 0; q < 4; ++q) x.v[q] = f_new.v[q] = 0.0;
            if (IN != IN_ZERO) x = lds4(rd_base + slot_off);
            if (NF > 0) f_new = lds4(rd_base + slot_off + 1024);
            __syncwarp();                                           // all lanes have read the slot: refill it
            issue(r + STREAM_DEPTH, slot_off, !FAST);
            slot_off = (slot_off + S4_SLOT) & (STREAM_DEPTH * S4_SLOT - 1);

            if (NF > 0) fr[k % NR] = f_new;

            // ---- S sweeps: stage t turns level t row (r-t-1) into level t+1
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const int i = r - t - 1;
                const d4 below = w[t][k & 1], c = w[t][(k & 1) ^ 1];
                const d4 f = fr[(k - t + 4 * NR) % NR];
                const double left = shfl_up1(c.v[3]), right = shfl_dn1(c.v[0]);
                d4 nx;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double l = q == 0 ? left : c.v[q - 1], rr = q == 3 ? right : c.v[q + 1];
                    nx.v[q] = jacobi_fast(c.v[q], sum4(x.v[q], below.v[q], rr, l), __dmul_rn(h2, f.v[q]));
                }
                if (!FAST) {
                    const bool row_in = i > 0 && i < N - 1;
#pragma unroll
                    for (int q = 0; q < 4; ++q) nx.v[q] = (row_in && in_col[q]) ? nx.v[q] : c.v[q];
                }
                w[t][k & 1] = x;
                x = nx;
            }

            // ---- x is now level S, row r-S
            {
                const int i = r - S;
                if (Op && i >= own_r_lo && i < own_r_hi) {
                    double *dst = Op + (ptrdiff_t)i * ldn + cx;
                    if (own01) *reinterpret_cast<double2 *>(dst) = make_double2(x.v[0], x.v[1]);
                    if (own23) *reinterpret_cast<double2 *>(dst + 2) = make_double2(x.v[2], x.v[3]);
                }
            }

            if (NEED_R) {
                const int rho = r - S - 1;
                const d4 below = w[S][k & 1], c = w[S][(k & 1) ^ 1];
                const d4 f = fr[(k - S + 4 * NR) % NR];
                const double left = shfl_up1(c.v[3]), right = shfl_dn1(c.v[0]);
                d4 res;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double l = q == 0 ? left : c.v[q - 1], rr = q == 3 ? right : c.v[q + 1];
                    res.v[q] = residual_fast(c.v[q], sum4(x.v[q], below.v[q], rr, l), f.v[q], inv_h2);
                }
                if (!FAST) {
                    const bool row_in = rho > 0 && rho < N - 1;
#pragma unroll
                    for (int q = 0; q < 4; ++q) res.v[q] = (row_in && in_col[q]) ? res.v[q] : 0.0;
                }
                w[S][k & 1] = x;
                if (ERR) {
                    // red = (row + column) even; cx is even: columns 0,2 on even rows, 1,3 on odd rows (:609-611)
                    const bool odd = rho & 1;
                    const double v0 = odd ? res.v[1] : res.v[0], v1 = odd ? res.v[3] : res.v[2];
                    const bool row_own = rho >= own_r_lo && rho < own_r_hi;
                    err_acc = __dadd_rn(err_acc, (row_own && own01) ? fabs(v0) : 0.0);
                    err_acc = __dadd_rn(err_acc, (row_own && own23) ? fabs(v1) : 0.0);
                }
                if (RES) {
                    d4 d_cur;
#pragma unroll
                    for (int q = 0; q < 4; ++q) d_cur.v[q] = -res.v[q];       // D = -D (:277-280)
                    const int f_row = rho - 1;
                    const double2 ri = rinfo_next;
                    if (FAST || (f_row + 1 >= 0 && f_row + 1 <= N - 1)) rinfo_next = p.rrow[f_row + 1];
                    else rinfo_next = make_double2(-1.0, 0.0);
                    const int crow = (int)ri.x;
                    if (crow >= 0 && f_row >= own_r_lo && f_row < own_r_hi) {
                        const double cw = ri.y;
                        const double np = shfl_dn1(d_prev.v[0]), nc = shfl_dn1(d_cur.v[0]);
                        const bool row_edge = crow == 0 || crow == p.M - 1;
                        double *out = p.Fc + (ptrdiff_t)crow * p.M;
                        int4 cc;
                        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(cc.x), "=r"(cc.y), "=r"(cc.z), "=r"(cc.w) : "r"(cc_addr) : "memory");
                        const d4 a = lds4(cc_addr + 16);
                        const int ccq[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (ccq[q] >= 0) {
                                const double p1 = q == 3 ? np : d_prev.v[q + 1], c1 = q == 3 ? nc : d_cur.v[q + 1];
                                const bool edge = row_edge || ccq[q] == 0 || ccq[q] == p.M - 1;
                                out[ccq[q]] = edge ? 0.0 : restrict_at(d_prev.v[q], p1, d_cur.v[q], c1, a.v[q], cw);
                            }
                        }
                    }
                    d_prev = d_cur;
                }
            }
        }
    };

    for (int rb = r_first; rb <= r_last; rb += U) {
        const bool fast = strip_fast && rb - NLV >= 1 && rb + U + STREAM_DEPTH <= N - 1 && rb - 1 >= p.row0 &&
                          rb + U + STREAM_DEPTH < p.row0 + p.rows;
        if (fast) chunk(BoolTag<true>(), rb);
        else chunk(BoolTag<false>(), rb);
    }

    cp_async_wait<0>();
    __syncwarp();
    if (ERR) {
        double v = err_acc;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
        if (lane == 0) p.partials[task] = v;
    }
  }  // task loop

    finish_launch<S4_WARPS, ERR>(p);
}

}  // namespace mg
