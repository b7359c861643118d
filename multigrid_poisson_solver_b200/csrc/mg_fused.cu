// mg_fused.cu -- smoothing passes and the fused -1 / 1 cycle legs, built on the
// register-streaming kernel of mg_stream.cuh.  Odd grid sizes (rows not 16-byte aligned) and
// transfer pairs whose restriction map is not injective take the baseline kernels of
// mg_kernels.cu instead; both paths produce identical bits.
#include "mg_fused.h"

#include "../../include/mg_abi.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mg_kernels.h"
#include "mg_stream.cuh"
#include "mg_stream4.cuh"
#include "mg_tile.cuh"

namespace mg {
namespace {

struct FusedRestrictTable {
    bool usable = false;
    int *f2c = nullptr;     // device [N]
    double *rw = nullptr;   // device [M]
    double2 *rrow = nullptr;  // device [N]: {coarse row or -1, its weight}
    std::vector<int> fine_of_coarse;   // host [M]: the fine row whose pair (f, f+1) produces coarse row c
};
std::map<std::pair<int, int>, FusedRestrictTable> g_restrict_tables;
int g_force_H = 0;
bool g_disable = false;

// Host construction of the restriction map with the very libm calls the reference makes
// (MG_solver_CPU.cpp:661-664), inverted: f2c[f] = coarse index t with floor(t*h_c/h_f) == f.
// Coarse boundary indices are mapped too (their value is forced to 0 by the kernel): 0 -> 0
// and M-1 -> N-2.  Unusable (-> baseline kernels) if two coarse indices share a fine index.
const FusedRestrictTable &fused_restrict_table(int N, int M)
{
    auto it = g_restrict_tables.find({N, M});
    if (it != g_restrict_tables.end()) return it->second;
    FusedRestrictTable t;
    if (M >= 3 && M < N && N >= 4) {
        const double h_f = 1.0 / (double)(N - 1), h_c = 1.0 / (double)(M - 1);
        std::vector<int> f2c((size_t)N, -1);
        std::vector<double> rw((size_t)M, 0.0);
        bool ok = true;
        for (int c = 1; c <= M - 2 && ok; ++c) {
            const double pos = (double)c * h_c;
            const int f = (int)floor(pos / h_f);
            rw[c] = fmod(pos, h_f) / h_f;
            if (f < 1 || f > N - 3 || f2c[f] != -1) ok = false;   // f+1 must stay inside, map must be injective
            else f2c[f] = c;
        }
        if (ok && f2c[0] == -1 && f2c[N - 2] == -1) {
            f2c[0] = 0;
            f2c[N - 2] = M - 1;
            std::vector<double2> rrow((size_t)N);
            for (int f = 0; f < N; ++f) rrow[f] = make_double2((double)f2c[f], f2c[f] >= 0 ? rw[f2c[f]] : 0.0);
            bool good = check(cudaMalloc(&t.f2c, (size_t)N * sizeof(int)), "cudaMalloc f2c");
            good = good && check(cudaMalloc(&t.rrow, (size_t)N * sizeof(double2)), "cudaMalloc rrow");
            good = good && check(cudaMemcpy(t.rrow, rrow.data(), (size_t)N * sizeof(double2), cudaMemcpyHostToDevice), "H2D rrow");
            good = good && check(cudaMalloc(&t.rw, (size_t)M * sizeof(double)), "cudaMalloc rw");
            // synchronous copies from pageable memory: the vectors die at the end of this scope
            good = good && check(cudaMemcpy(t.f2c, f2c.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice), "H2D f2c");
            good = good && check(cudaMemcpy(t.rw, rw.data(), (size_t)M * sizeof(double), cudaMemcpyHostToDevice), "H2D rw");
            t.usable = good;
            t.fine_of_coarse.assign((size_t)M, 0);
            for (int f = 0; f < N; ++f)
                if (f2c[f] >= 0) t.fine_of_coarse[f2c[f]] = f;
        }
    }
    return g_restrict_tables.emplace(std::make_pair(N, M), t).first->second;
}

// ---- row segments of a pass (one task = one strip x one segment)
// Measured on B200 for every pass type, N = 2048..16384, uniform height H: the time behaves like
// x (1 + warm-up/H) + c H with x = rows per resident warp -- redundant warm-up rows against a tail
// in which warps finish up to one task apart (best uniform H ~ 3.5 sqrt(x)).  The table therefore
// hands out tall segments first and shrinks them towards the end of the queue (guided
// self-scheduling): height = remaining work / (g x resident warps), clamped to [h_min, h_max].
// Grids with fewer tasks than resident warps are latency bound instead (a lone warp needs about
// 1 us per row): they get the uniform height that minimises rounds x (rows + warm-up + fixed cost).
// Split passes of the slab driver: the edge launch (subset 1) runs two EDGE_H-row segments at each
// end of the owned range -- they produce every row a neighbour's halo needs, also of the restricted
// grid (8 coarse rows <= 22 fine rows up to ratio 2.5; the last segment may be ragged, hence two) --
// and the interior launch (subset 2) the rows in between.
struct SegmentTable {
    int2 *dev = nullptr;
    int n = 0;
};
int g_sched_hmax = 256, g_sched_hmin = 16;
double g_sched_g = 1.5;

std::vector<int2> build_segments(int rows, int n_strips, int resident_warps, int lead_rows, int subset)
{
    constexpr int EDGE_H = 24, EDGE_E = 2;
    std::vector<int2> out;
    int lo = 0, hi = rows;
    if (subset != 0) {
        const int n_edge = (rows + EDGE_H - 1) / EDGE_H;
        if (n_edge < 2 * EDGE_E + 1) {               // too thin to split: the edge launch does it all
            if (subset == 2) return out;
        } else if (subset == 1) {
            for (int k = 0; k < EDGE_E; ++k) out.push_back(make_int2(k * EDGE_H, (k + 1) * EDGE_H));
            for (int k = n_edge - EDGE_E; k < n_edge; ++k) out.push_back(make_int2(k * EDGE_H, std::min(rows, (k + 1) * EDGE_H)));
            return out;
        } else {
            lo = EDGE_H * EDGE_E;
            hi = EDGE_H * (n_edge - EDGE_E);
        }
    }
    const int span = hi - lo;
    auto uniform = [&](int H) {
        for (int r = lo; r < hi; r += H) out.push_back(make_int2(r, std::min(hi, r + H)));
    };
    if (g_force_H > 0) { uniform(g_force_H); return out; }
    const double x = (double)span * n_strips / (double)resident_warps;
    int H = (int)(3.5 * std::sqrt(x) + 0.5);
    H = std::max(8, std::min(256, H >= 32 ? (H + 4) / 8 * 8 : (H + 2) / 4 * 4));
    if ((long long)((span + H - 1) / H) * n_strips < resident_warps) {      // latency-bound small grid
        long long best = -1;
        for (int h = 4; h <= 32; ++h) {
            const long long tasks = (long long)((span + h - 1) / h) * n_strips;
            const long long rounds = (tasks + resident_warps - 1) / resident_warps;
            const long long cost = rounds * (h + lead_rows + 6);
            if (best < 0 || cost <= best) { best = cost; H = h; }
        }
        uniform(H);
        return out;
    }
    if (g_sched_g <= 0.0) { uniform(H); return out; }
    int r = lo;
    while (r < hi) {
        const int left = hi - r;
        int h = (int)((double)left * n_strips / (g_sched_g * resident_warps));
        h = std::max(g_sched_hmin, std::min(g_sched_hmax, h / 4 * 4));
        if (left - h < g_sched_hmin / 2) h = left;   // no sliver at the end
        h = std::min(h, left);
        out.push_back(make_int2(r, r + h));
        r += h;
    }
    return out;
}

// a pass over `rows` owned rows can be launched as edge segments + interior (build_segments: subset 1 / 2)
bool segments_splittable(int rows) { return (rows + 23) / 24 >= 5; }

const SegmentTable &segment_table(int rows, int n_strips, int resident_warps, int lead_rows, int subset)
{
    static std::map<std::vector<int>, SegmentTable> cache;
    const std::vector<int> key = {rows, n_strips, resident_warps, lead_rows, subset};
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    SegmentTable t;
    const std::vector<int2> segs = build_segments(rows, n_strips, resident_warps, lead_rows, subset);
    t.n = (int)segs.size();
    if (t.n > 0) {
        bool ok = check(cudaMalloc(&t.dev, segs.size() * sizeof(int2)), "cudaMalloc segment table");
        ok = ok && check(cudaMemcpy(t.dev, segs.data(), segs.size() * sizeof(int2), cudaMemcpyHostToDevice), "H2D segment table");
        if (!ok) t.n = 0;
    }
    return cache.emplace(key, t).first->second;
}

// Whole grids up to g_tile_max_N take the shared-memory tile kernel (mg_tile.cuh): by default only ODD sizes (the
// streaming kernel needs even N and is as fast on small even grids: 14 us per node either way at N = 256, measured);
// MG_TILE_MAX_N / mgSetTileMaxN(n >= 0) route every size up to n through it (0 disables it).
int g_tile_max_N = 1024;
bool g_tile_even = false;
int g_tile_odd_max_N = 1 << 30;   // ODD sizes of any extent take the tile kernel (the streaming kernels need 16-byte aligned rows); the
                                  // alternative is one kernel per operator (~12 launches and ~6x the traffic per node)
int g_cols4 = 1;   // 4 columns per lane (mg_stream4.cuh) for the passes that have a variant; MG_COLS4=0 disables
int g_strip = 1;   // the bulk-copy 4-column kernel (mg_strip.cuh, instantiated in mg_legs.cu) for smoothing passes; MG_STRIP=0: never
int g_strip_min_N = 8192;   // ... from this grid size on (MG_STRIP_MIN_N)
long long g_split_min_points = 32ll << 20;   // slab passes are split into edge + interior launches from this many owned points on (MG_SPLIT_MIN_POINTS)

// Task geometry and persistent grid shared by the streaming kernels: fills the task fields of `p`, shifts the array
// bases to global rows and returns the CTA count (0: nothing to launch).
}  // namespace

int stream_launch_prepare(StreamParams &p, int W, int warps, int min_ctas, bool err, int lead_rows)
{
    Context &c = ctx();
    const int N = p.N;
    p.n_strips = (N + W - 1) / W;
    const int resident_warps = min_ctas * c.sm_count * warps;
    const SegmentTable &st = segment_table(p.own_hi - p.own_lo, p.n_strips, resident_warps, lead_rows, p.subset);
    if (st.n == 0) return 0;                         // (an interior launch with nothing left to do)
    p.segs = st.dev;
    p.n_segs = st.n;
    p.n_tasks = p.n_strips * p.n_segs;
    // the kernel indexes every grid with GLOBAL rows: shift the bases of the local arrays
    p.F_valid = p.F;
    const ptrdiff_t fine_shift = (ptrdiff_t)p.row0 * N;
    if (p.Uin) p.Uin -= fine_shift;
    if (p.F) p.F -= fine_shift;
    if (p.Uout) p.Uout -= fine_shift;
    if (p.Fc) p.Fc -= (ptrdiff_t)p.fc_row0 * p.M;
    if (p.Uc) p.Uc -= (ptrdiff_t)p.uc_row0 * p.Nc;
    if (err) p.partials = partials_buf(2 * (size_t)p.n_tasks);   // (MID passes keep a second partial per task)
    p.counter = c.counters + 8;   // [8] queue head, [9] finished warps (self-resetting)
    return std::max(1, std::min(min_ctas * c.sm_count, (p.n_tasks + warps - 1) / warps));
}

namespace {

template <typename Kernel>
void launch_stream_kernel(Kernel kernel, StreamParams &p, int W, int warps, int min_ctas, int smem_bytes, bool err, int lead_rows,
                          bool &opted_in)
{
    Context &c = ctx();
    const int N = p.N;
    const int blocks = stream_launch_prepare(p, W, warps, min_ctas, err, lead_rows);
    if (blocks == 0) return;
    if (!opted_in) {
        check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), "cudaFuncSetAttribute(k_stream)");
        opted_in = true;
    }
    static const char *trace_path = getenv("MG_TASK_TRACE");       // debug: dump the task time line of every launch
    if (trace_path) check(cudaMalloc(&p.trace, (size_t)(4 * p.n_tasks + 1) * sizeof(unsigned long long)), "cudaMalloc trace");
    kernel<<<blocks, warps * 32, smem_bytes, c.stream>>>(p);
    c.launches++;
    check(cudaGetLastError(), "k_stream");
    if (trace_path && p.trace) {
        std::vector<unsigned long long> h((size_t)4 * p.n_tasks + 1);
        cudaStreamSynchronize(c.stream);
        cudaMemcpy(h.data(), p.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaFree(p.trace);
        if (FILE *f = fopen(trace_path, "a")) {
            fprintf(f, "launch N=%d strips=%d segs=%d tasks=%d warps=%d end=%llu\n", N, p.n_strips, p.n_segs, p.n_tasks, blocks * warps, h.back());
            for (int t = 0; t < p.n_tasks; ++t)
                fprintf(f, "%d %llu %llu %llu %llu\n", t, h[4 * t], h[4 * t + 1], h[4 * t + 2], h[4 * t + 3]);
            fclose(f);
        }
    }
}

// Whole grids of odd size (any extent) or, when asked for, small even ones: one CTA per 32 x 32 tile instead of one warp per strip (mg_tile.cuh).
bool tile_ok_size(int N) { return (N % 2 != 0) ? N <= g_tile_odd_max_N : (g_tile_even && N <= g_tile_max_N); }
bool tile_ok(const StreamParams &p)
{
    return tile_ok_size(p.N) && p.row0 == 0 && p.rows == p.N && p.own_lo == 0 && p.own_hi == p.N &&
           !p.raw_sum && !p.subset && !p.err_add;
}


template <int S, int IN, bool ERR, bool RES>
void launch_tile(StreamParams &p)
{
    Context &c = ctx();
    const int tiles = (p.N + TILE - 1) / TILE, blocks = tiles * tiles;
    if (ERR) {
        p.partials = partials_buf((size_t)blocks);
        p.counter = c.counters + 10;   // last-CTA ticket (self-resetting)
    }
    k_tile<S, IN, ERR, RES><<<blocks, TILE_THREADS, 0, c.stream>>>(p);
    c.launches++;
    check(cudaGetLastError(), "k_tile");
}

template <int S, int IN, bool ERR, bool RES>
void launch_stream(StreamParams &p)
{
    if (tile_ok(p)) { launch_tile<S, IN, ERR, RES>(p); return; }
    // Measured on B200 (N = 16384): 4 columns per lane win for the passes without restriction (smoothing
    // pass 1.05 vs 1.09 ms); with restriction the 230-254 registers leave 8 warps per SM and lose (1.40 vs 1.22 ms).
    if (IN != IN_PROLONG && !RES && g_cols4) {
        // instantiated only for IN_LOAD / IN_ZERO (the branch is dead for IN_PROLONG)
        constexpr int IN4 = IN == IN_PROLONG ? IN_LOAD : IN;
        using G4 = Stream4Geo<S, ERR || RES, RES>;
        static bool opted4 = false;
        launch_stream_kernel(k_stream4<S, IN4, ERR, RES>, p, G4::W, S4_WARPS, S4_MIN_CTAS, stream4_smem_bytes(), ERR, 2 * S + 3, opted4);
        return;
    }
    using G = StreamGeo<S, ERR || RES, RES>;
    constexpr int STREAM_WARPS = stream_shape(RES).warps, STREAM_MIN_CTAS = stream_shape(RES).min_ctas;
    static bool opted_in = false;   // one flag per instantiation
    launch_stream_kernel(k_stream<S, IN, ERR, RES>, p, G::W, STREAM_WARPS, STREAM_MIN_CTAS, stream_smem_bytes(IN, STREAM_WARPS, RES), ERR, 2 * S + 3,
                         opted_in);
}

template <int IN, bool ERR, bool RES>
void launch_stream_s(int S, StreamParams &p)
{
    switch (S) {
        case 0: launch_stream<0, IN, ERR, RES>(p); break;
        case 1: launch_stream<1, IN, ERR, RES>(p); break;
        case 2: launch_stream<2, IN, ERR, RES>(p); break;
        default: launch_stream<3, IN, ERR, RES>(p); break;
    }
}

// mode: 0 plain, 1 ERR, 2 ERR+RES
void launch_stream_any(int S, int in, int mode, StreamParams &p)
{
    // Which streaming kernel (all bit-identical; measured on B200): the 4-column bulk-copy kernel (k_strip) wins
    // for passes WITHOUT restriction / prolongation on large grids (HBM bound: 0.93 of the measured peak against 0.87); the
    // -1 and 1 nodes are issue bound and need the 12-16 warps per SM only the 2-column kernel (k_stream) leaves room for.
    // Slab passes with peer memory: k_strip or k_stream (k_stream4 has no peer stores).
    const bool peers = p.peer_U_lo || p.peer_U_hi || p.peer_Fc_lo || p.peer_Fc_hi;
    const bool plain_pass = in != IN_PROLONG && mode != 2;
    if (g_strip && !tile_ok(p) && plain_pass && (p.N >= g_strip_min_N || peers)) { launch_strip(S, in, mode, p); return; }
    if (peers) { launch_stream_peer(S, in, mode, p); return; }
    if (in == IN_LOAD) {
        if (mode == 0) launch_stream_s<IN_LOAD, false, false>(S, p);
        else if (mode == 1) launch_stream_s<IN_LOAD, true, false>(S, p);
        else launch_stream_s<IN_LOAD, true, true>(S, p);
    } else if (in == IN_ZERO) {
        if (mode == 0) launch_stream_s<IN_ZERO, false, false>(S, p);
        else if (mode == 1) launch_stream_s<IN_ZERO, true, false>(S, p);
        else launch_stream_s<IN_ZERO, true, true>(S, p);
    } else {
        if (mode == 0) launch_stream_s<IN_PROLONG, false, false>(S, p);
        else launch_stream_s<IN_PROLONG, true, false>(S, p);
    }
}

bool streamable(int N) { return !g_disable && N >= 4 && (N % 2 == 0); }
// a whole grid the fused passes can serve: even (streaming kernel) or small (tile kernel, any parity)
bool fusable(int N) { return !g_disable && N >= 4 && (N % 2 == 0 || N <= g_tile_odd_max_N); }

// Split `step` sweeps into passes of at most STREAM_SMAX, as evenly as possible.
std::vector<int> split_passes(int step)
{
    std::vector<int> out;
    if (step <= 0) return out;
    const int n = (step + STREAM_SMAX - 1) / STREAM_SMAX;
    for (int k = 0; k < n; ++k) out.push_back(step / n + (k < step % n ? 1 : 0));
    return out;
}

struct LegSpec {
    int in = IN_LOAD;          // level 0 of the FIRST pass
    bool want_err = false;     // error after the LAST pass
    bool want_res = false;     // restriction after the LAST pass
    // prolongation (first pass)
    int Nc = 0;
    const double *Uc = nullptr;
    // restriction (last pass)
    int M = 0;
    double *Fc = nullptr;
    double *err_dev = nullptr, *err_slot = nullptr;
    // row slab (defaults = the whole grid, filled in by run_leg when rows == 0)
    int row0 = 0, rows = 0, own_lo = 0, own_hi = 0, fc_row0 = 0, uc_row0 = 0, uc_rows = 0;
    bool raw_sum = false;
};

// Runs the passes of one leg on even N.  The first pass reads `in` (never written, unused for
// IN_ZERO); the passes write alternately to `a`, `b`, `a`, ...  Returns the buffer holding the
// final level (`in` itself if no pass produced a new level).
double *run_leg(int N, double L, const double *in, double *a, double *b, const double *F, int step, const LegSpec &spec)
{
    const Spacing sp = spacing(N, L);
    std::vector<int> passes = split_passes(step);
    if (passes.empty()) passes.push_back(0);   // a pass without sweeps (prolong-add only / residual+restrict only)
    const double *src = in;
    double *bufs[2] = {a, b};
    int next = 0;
    for (size_t k = 0; k < passes.size(); ++k) {
        const bool first = k == 0, last = k + 1 == passes.size();
        StreamParams p{};
        p.N = N;
        p.row0 = spec.rows ? spec.row0 : 0;
        p.rows = spec.rows ? spec.rows : N;
        p.own_lo = spec.rows ? spec.own_lo : 0;
        p.own_hi = spec.rows ? spec.own_hi : N;
        p.fc_row0 = spec.fc_row0;
        p.uc_row0 = spec.uc_row0;
        p.uc_rows = spec.uc_rows ? spec.uc_rows : spec.Nc;
        p.raw_sum = spec.raw_sum ? 1 : 0;
        p.h2 = sp.h2;
        p.inv_h2 = sp.inv_h2;
        p.F = F;
        const int in_mode = first ? spec.in : IN_LOAD;
        p.Uin = src;
        double *out = bufs[next];
        p.Uout = out;
        int mode = 0;
        if (last && spec.want_res) mode = 2;
        else if (last && spec.want_err) mode = 1;
        if (mode >= 1) {
            p.err_dev = spec.err_dev;
            p.err_slot = spec.err_slot;
        }
        if (mode == 2) {
            const FusedRestrictTable &t = fused_restrict_table(N, spec.M);
            p.M = spec.M;
            p.Fc = spec.Fc;
            p.f2c = t.f2c;
            p.rw = t.rw;
            p.rrow = t.rrow;
        }
        if (in_mode == IN_PROLONG) {
            const ProlongTable &t = prolong_table(spec.Nc, N);
            p.Nc = spec.Nc;
            p.Uc = spec.Uc;
            p.row_cell = t.row_cell;
            p.col_cell = t.col_cell;
            p.row_w = t.row_w;
            p.col_w = t.col_w;
            p.row_info = t.row_info;
            p.c_dx = 1.0 / (double)(spec.Nc - 1);
            p.inv_c_dx = 1.0 / p.c_dx;
        }
        if (passes[k] == 0 && in_mode == IN_LOAD && mode == 0) break;   // nothing to do
        if (passes[k] == 0 && in_mode == IN_LOAD) {
            // residual / restriction of the input itself: no new level is produced, keep `src`
            p.Uout = nullptr;   // level S is the input itself: nothing to write
            launch_stream_any(0, in_mode, mode, p);
            continue;
        }
        launch_stream_any(passes[k], in_mode, mode, p);
        src = out;
        next ^= 1;
    }
    return const_cast<double *>(src);
}

}  // namespace

std::vector<int> segment_plan(int rows, int n_strips, int resident_warps, int lead_rows, int subset)
{
    std::vector<int> out;
    for (const int2 &sg : build_segments(rows, n_strips, resident_warps, lead_rows, subset)) { out.push_back(sg.x); out.push_back(sg.y); }
    return out;
}


// ------------------------------------------------------------------ row-slab passes (multi-GPU)
bool slab_pair_fusable(int N, int M)
{
    return streamable(N) && M >= 3 && fused_restrict_table(N, M).usable && (double)(N - 1) >= 1.2 * (double)(M - 1);
}

int restrict_first_coarse_at_or_after(int N, int M, int fine_row)
{
    const FusedRestrictTable &t = fused_restrict_table(N, M);
    int c = 0;
    while (c < M && t.fine_of_coarse[c] < fine_row) ++c;   // fine_of_coarse is strictly increasing
    return c;
}

// A slab pass with peer memory is launched in two parts: the thin row segments at both ends of the owned range -- they
// produce every row a neighbour keeps as a halo, also of the restricted grid (8 coarse rows <= 22 fine rows up to ratio
// 2.5) -- with the peer-store instantiation, then the interior with the plain kernel.  The halo rows cross NVLink while
// the interior is swept, and the bulk of the pass runs the very kernel of the single-GPU path (the peer-store variants
// cost 5-10 %: measured).  Slabs too thin to split: one launch.
template <class Launch>
void launch_slab(const StreamParams &base, const PeerLinks &peers, Launch launch)
{
    auto with_peers = [&](StreamParams &q) {
        q.peer_U_lo = peers.U_lo; q.peer_U_hi = peers.U_hi; q.u_lo_end = peers.u_lo_end; q.u_hi_begin = peers.u_hi_begin;
        q.peer_Fc_lo = peers.Fc_lo; q.peer_Fc_hi = peers.Fc_hi; q.fc_lo_end = peers.fc_lo_end; q.fc_hi_begin = peers.fc_hi_begin;
    };
    auto with_flags = [&](StreamParams &q) { q.flag_lo = peers.flag_lo; q.flag_hi = peers.flag_hi; q.flag_val = peers.flag_val; };
    const bool stores = peers.U_lo || peers.U_hi || peers.Fc_lo || peers.Fc_hi;
    // (small slabs are latency bound: a second launch costs more than the peer-store variant does)
    if (stores && segments_splittable(base.own_hi - base.own_lo) && (long long)(base.own_hi - base.own_lo) * base.N >= g_split_min_points) {
        // The pass number is published by the EDGE launch: once it has drained, every row a neighbour needs has been stored
        // into its slab, and this rank's own halo rows -- read by the edge segments only (the interior keeps >= 48 rows away
        // from both ends) -- are free to be overwritten by the neighbours' next pass.  The interior launches of different
        // ranks therefore never wait for one another: skew between ranks is absorbed instead of amplified.
        StreamParams a = base, b = base;
        a.subset = 1;
        with_peers(a);
        with_flags(a);
        launch(a);
        b.subset = 2;
        b.err_add = b.err_dev ? 1 : 0;           // the interior adds its error sum(s) to the edge launch's
        launch(b);
    } else {
        StreamParams q = base;
        with_peers(q);
        with_flags(q);
        launch(q);
    }
}

void slab_pass(int N, double L, int S, int in_mode, const double *Uin, const double *F, double *Uout, const Slab &fine,
               bool want_err, double *raw_err_dev, int M, double *Fc, const Slab *coarse_out, int Nc, const double *Uc,
               const Slab *coarse_in, const PeerLinks &peers)
{
    const Spacing sp = spacing(N, L);
    StreamParams p{};
    p.N = N;
    p.row0 = fine.row0;
    p.rows = fine.rows;
    p.own_lo = fine.own_lo;
    p.own_hi = fine.own_hi;
    p.raw_sum = 1;
    p.h2 = sp.h2;
    p.inv_h2 = sp.inv_h2;
    p.F = F;
    p.Uin = Uin;
    p.Uout = Uout;
    int mode = want_err ? 1 : 0;
    if (want_err) p.err_dev = raw_err_dev;
    if (coarse_out) {
        const FusedRestrictTable &t = fused_restrict_table(N, M);
        mode = 2;
        p.M = M;
        p.Fc = Fc;
        p.fc_row0 = coarse_out->row0;
        p.f2c = t.f2c;
        p.rw = t.rw;
        p.rrow = t.rrow;
        p.err_dev = want_err ? raw_err_dev : nullptr;
    }
    if (in_mode == IN_PROLONG) {
        const ProlongTable &t = prolong_table(Nc, N);
        p.Nc = Nc;
        p.Uc = Uc;
        p.uc_row0 = coarse_in->row0;
        p.uc_rows = coarse_in->rows;
        p.row_cell = t.row_cell;
        p.col_cell = t.col_cell;
        p.row_w = t.row_w;
        p.col_w = t.col_w;
        p.row_info = t.row_info;
        p.c_dx = 1.0 / (double)(Nc - 1);
        p.inv_c_dx = 1.0 / p.c_dx;
    }
    launch_slab(p, peers, [&](StreamParams &q) { launch_stream_any(S, in_mode, mode, q); });
}

// ------------------------------------------------------------------ error-trigger loops (con_step = -1)
namespace {

// One pass of the trigger loop on a whole grid or a slab: S = 2 sweeps with BOTH errors (err[0] after the second, err[1]
// after the first sweep), or S = 1 with err[0]; optional prolongation first (first pass of a 1 node), optional restriction.
struct TriggerPass {
    int N = 0, S = 2, in_mode = IN_LOAD;
    double L = 1.0;
    const double *in = nullptr, *F = nullptr;
    double *out = nullptr;
    int M = 0;                     // > 0: restrict into Fc
    double *Fc = nullptr;
    int Nc = 0;                    // IN_PROLONG
    const double *Uc = nullptr;
    double *err_dev = nullptr, *err_slot = nullptr;   // two consecutive doubles each
    // row slab (rows == 0: the whole grid)
    Slab fine, coarse_out, coarse_in;
    bool slab = false;
    PeerLinks peers;
};

void run_trigger_pass(const TriggerPass &t)
{
    const Spacing sp = spacing(t.N, t.L);
    StreamParams p{};
    p.N = t.N;
    p.row0 = t.slab ? t.fine.row0 : 0;
    p.rows = t.slab ? t.fine.rows : t.N;
    p.own_lo = t.slab ? t.fine.own_lo : 0;
    p.own_hi = t.slab ? t.fine.own_hi : t.N;
    p.raw_sum = t.slab ? 1 : 0;
    p.h2 = sp.h2;
    p.inv_h2 = sp.inv_h2;
    p.F = t.F;
    p.Uin = t.in;
    p.Uout = t.out;
    p.err_dev = t.err_dev;
    p.err_slot = t.err_slot;
    if (t.M > 0) {
        const FusedRestrictTable &rt = fused_restrict_table(t.N, t.M);
        p.M = t.M;
        p.Fc = t.Fc;
        p.fc_row0 = t.slab ? t.coarse_out.row0 : 0;
        p.f2c = rt.f2c;
        p.rw = rt.rw;
        p.rrow = rt.rrow;
    }
    if (t.in_mode == IN_PROLONG) {
        const ProlongTable &pt = prolong_table(t.Nc, t.N);
        p.Nc = t.Nc;
        p.Uc = t.Uc;
        p.uc_row0 = t.slab ? t.coarse_in.row0 : 0;
        p.uc_rows = t.slab ? t.coarse_in.rows : t.Nc;
        p.row_cell = pt.row_cell;
        p.col_cell = pt.col_cell;
        p.row_w = pt.row_w;
        p.col_w = pt.col_w;
        p.row_info = pt.row_info;
        p.c_dx = 1.0 / (double)(t.Nc - 1);
        p.inv_c_dx = 1.0 / p.c_dx;
    }
    launch_slab(p, t.peers, [&](StreamParams &q) {
        if (t.S == 2) launch_stream_mid(t.in_mode, t.M > 0, q);
        else launch_stream_any(t.S, t.in_mode, t.M > 0 ? 2 : 1, q);
    });
}

constexpr double TRIGGER = 0.01;   // MG_solver_CPU.cpp:99

// The loop of :216-230 / :388-402 -- sweep, error, stop when two successive errors differ by <= TRIGGER (at least two
// sweeps) -- taken two sweeps per launch.  `first_mode` is the level 0 of the first pass (zero / load / prolong-add).  If the
// loop ends after an odd number of sweeps the last launch went one sweep too far: it is repeated with one sweep from the
// same input (passes are out of place, the input is still there).  With M > 0 every launch also restricts its result (the
// last one is the one that counts), so a -1 node whose trigger fires at the minimum of two sweeps is ONE launch.
double *trigger_loop(int N, double L, int first_mode, double *U, double *U_work, const double *F, int M, double *Fc, int Nc,
                     const double *Uc, int *steps_out, double *error_out)
{
    Context &c = ctx();
    double *slot = c.slots_host + (MG_SCALAR_SLOTS - 4), *slot_dev = c.slots_dev + (MG_SCALAR_SLOTS - 4);
    double *cur = U, *other = U_work;
    int done = 0, mode = first_mode;
    double prev = 0.0, err = 0.0;
    for (;;) {
        TriggerPass t;
        t.N = N; t.L = L; t.S = 2; t.in_mode = mode; t.in = cur; t.out = other; t.F = F; t.M = M; t.Fc = Fc; t.Nc = Nc; t.Uc = Uc;
        t.err_dev = c.dev_scalar + 2;
        t.err_slot = slot_dev;
        run_trigger_pass(t);
        check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
        if (c.err_code) break;
        const double e2 = slot[0], e1 = slot[1];
        if (done + 1 > 1 && std::fabs(e1 - prev) <= TRIGGER) {        // the loop ends after the first of these two sweeps
            t.S = 1;
            run_trigger_pass(t);                                         // same input, one sweep (and its restriction)
            check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
            done += 1;
            err = slot[0];
            cur = other;
            break;
        }
        done += 2;
        cur = other;
        other = (cur == U) ? U_work : U;
        mode = IN_LOAD;
        err = e2;
        if (std::fabs(e2 - e1) <= TRIGGER) break;
        prev = e2;
    }
    if (steps_out) *steps_out = done;
    if (error_out) *error_out = err;
    return cur;
}

}  // namespace

bool trigger_fusable_down(int N, int M) { return streamable(N) && !tile_ok_size(N) && fused_restrict_table(N, M).usable; }
bool trigger_fusable_up(int Nc, int N) { return streamable(N) && !tile_ok_size(N) && Nc >= 2 && (double)(N - 1) >= 1.2 * (double)(Nc - 1); }

double *down_leg_trigger(int N, double L, double *U, double *U_work, const double *F, bool zero_init, int M, double *F_c, int *steps,
                         double *error)
{
    return trigger_loop(N, L, zero_init ? IN_ZERO : IN_LOAD, U, U_work, F, M, F_c, 0, nullptr, steps, error);
}

double *up_leg_trigger(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int *steps, double *error)
{
    return trigger_loop(N, L, IN_PROLONG, U_f, U_work, F, 0, nullptr, Nc, U_c, steps, error);
}

void slab_trigger_pass(int N, double L, int S, int in_mode, const double *Uin, const double *F, double *Uout, const Slab &fine, double *raw_err_dev2,
                       int M, double *Fc, const Slab *coarse_out, int Nc, const double *Uc, const Slab *coarse_in, const PeerLinks &peers)
{
    TriggerPass t;
    t.N = N; t.L = L; t.S = S; t.in_mode = in_mode; t.in = Uin; t.out = Uout; t.F = F;
    t.M = coarse_out ? M : 0; t.Fc = Fc; t.Nc = Nc; t.Uc = Uc;
    t.err_dev = raw_err_dev2;
    t.slab = true;
    t.fine = fine;
    if (coarse_out) t.coarse_out = *coarse_out;
    if (coarse_in) t.coarse_in = *coarse_in;
    t.peers = peers;
    run_trigger_pass(t);
}

void fused_init()
{
    if (const char *h = getenv("MG_STREAM_H")) g_force_H = atoi(h);
    if (const char *h = getenv("MG_SCHED_HMAX")) g_sched_hmax = std::max(4, atoi(h));
    if (const char *h = getenv("MG_SCHED_HMIN")) g_sched_hmin = std::max(4, atoi(h));
    if (const char *h = getenv("MG_SCHED_G")) g_sched_g = atof(h);     // <= 0: uniform segments
    if (const char *d = getenv("MG_NO_STREAM")) g_disable = atoi(d) != 0;
    if (const char *d = getenv("MG_COLS4")) g_cols4 = atoi(d);
    if (const char *d = getenv("MG_STRIP")) g_strip = atoi(d);
    if (const char *d = getenv("MG_STRIP_MIN_N")) g_strip_min_N = atoi(d);
    if (const char *d = getenv("MG_SPLIT_MIN_POINTS")) g_split_min_points = atoll(d);
    if (const char *d = getenv("MG_TILE_MAX_N")) { g_tile_max_N = g_tile_odd_max_N = std::max(0, atoi(d)); g_tile_even = true; }
}

int set_tile_max_n(int n)
{
    const int old = g_tile_even ? g_tile_max_N : -1;
    if (n >= 0) { g_tile_max_N = g_tile_odd_max_N = n; g_tile_even = true; }      // every size up to n (0: never)
    else        { g_tile_max_N = 1024; g_tile_odd_max_N = 1 << 30; g_tile_even = false; }   // default: odd sizes only, any extent
    return old;
}

int smooth_pass_count(int N, int step)
{
    if (step <= 0) return 0;
    return fusable(N) ? (step + STREAM_SMAX - 1) / STREAM_SMAX : step;
}

double *smooth_out_of_place(int N, double L, const double *in, double *a, double *b, const double *F, int step,
                            bool in_is_zero, double *err_dev, double *err_slot)
{
    const bool want_err = err_dev || err_slot;
    if (step == 0 && in_is_zero) {
        check(cudaMemsetAsync(a, 0, (size_t)N * N * sizeof(double), ctx().stream), "cudaMemsetAsync");
        in = a;
        std::swap(a, b);
    }
    if (fusable(N) && (step > 0 || want_err)) {
        LegSpec spec;
        spec.in = (in_is_zero && step > 0) ? IN_ZERO : IN_LOAD;
        spec.want_err = want_err;
        spec.err_dev = err_dev;
        spec.err_slot = err_slot;
        return run_leg(N, L, in, a, b, F, step, spec);
    }
    const Spacing sp = spacing(N, L);
    const double *cur = in;
    double *bufs[2] = {a, b};
    for (int s = 0; s < step; ++s) {
        double *out = bufs[s & 1];
        launch_sweep(N, sp.h2, cur, F, out, s == 0 && in_is_zero);
        cur = out;
    }
    if (want_err) launch_smooth_error(N, sp.inv_h2, cur, F, err_dev, err_slot);
    return const_cast<double *>(cur);
}

double *down_leg(int N, double L, double *U, double *U_work, const double *F, int step, bool zero_init, int M,
                 double *F_c, double *err_slot)
{
    if (fusable(N) && fused_restrict_table(N, M).usable) {
        if (step == 0 && zero_init) check(cudaMemsetAsync(U, 0, (size_t)N * N * sizeof(double), ctx().stream), "cudaMemsetAsync");
        LegSpec spec;
        spec.in = (zero_init && step > 0) ? IN_ZERO : IN_LOAD;
        spec.want_err = true;
        spec.want_res = true;
        spec.M = M;
        spec.Fc = F_c;
        spec.err_dev = step > 0 ? ctx().dev_scalar : nullptr;
        spec.err_slot = step > 0 ? err_slot : nullptr;
        return run_leg(N, L, U, U_work, U, F, step, spec);
    }
    const Spacing sp = spacing(N, L);
    double *res = smooth_out_of_place(N, L, U, U_work, U, F, step, zero_init, step > 0 ? ctx().dev_scalar : nullptr,
                                      step > 0 ? err_slot : nullptr);
    double *D = scratch_grid((size_t)N * N);
    if (!D) return res;
    launch_residual(N, sp.inv_h2, res, F, D);
    launch_negate(N, D);
    launch_restrict(N, D, M, F_c);
    return res;
}

double *up_leg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, const double *F, int step,
               double *err_slot)
{
    // a 64-column window must map into at most 62 coarse cells (the staged part of a coarse row)
    if (fusable(N) && Nc >= 2 && (double)(N - 1) >= 1.2 * (double)(Nc - 1)) {
        LegSpec spec;
        spec.in = IN_PROLONG;
        spec.Nc = Nc;
        spec.Uc = U_c;
        spec.want_err = step > 0;
        spec.err_dev = step > 0 ? ctx().dev_scalar : nullptr;
        spec.err_slot = step > 0 ? err_slot : nullptr;
        return run_leg(N, L, U_f, U_work, U_f, F, step > 0 ? step : 0, spec);
    }
    launch_prolong(Nc, U_c, N, U_f, U_f);
    if (step <= 0) return U_f;
    return smooth_out_of_place(N, L, U_f, U_work, U_f, F, step, false, ctx().dev_scalar, err_slot);
}

}  // namespace mg
