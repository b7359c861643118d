// mg_driver.cpp -- Cycle.txt interpreter over a device-resident level stack.
//
// Host-side C++ that calls only the C ABI of include/mg_abi.h.  It reproduces the control
// flow of the reference's main() (MG_solver_CPU.cpp:36-462) and of its LinkedList level
// stack (linkedlist.cpp:7-124): the same node stream, option parsing for all six
// (con_step, con_N) cases, the U-zeroing rule, the init/restart flag, the error-trigger
// loops, and the same stdout blocks.  Two execution modes:
//   MG_RUN_UNFUSED  one ABI operator per reference call (literal drop-in)
//   MG_RUN_FUSED    mgDownLeg / mgUpLeg per node, no D / tempU grids, no host sync per node
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/mg_abi.h"

namespace {

constexpr double TRIGGER = 0.01;  // MG_solver_CPU.cpp:99

// One NVTX range per node of the cycle ("node -1 N=16384"): shows up in nsys / ncu time lines; a no-op without a profiler.
struct NodeRange {
    NodeRange(int node, int N)
    {
        char label[48];
        snprintf(label, sizeof label, "node %d N=%d", node, N);
        nvtxRangePushA(label);
    }
    ~NodeRange() { nvtxRangePop(); }
};

struct Level {
    int N = 0;
    double *U = nullptr;     // current solution buffer
    double *W = nullptr;     // ping-pong partner of U (fused mode) / unused
    double *F = nullptr;
    double *D = nullptr;     // only materialised in unfused mode
    int step = 0;
    double smoothing_error = 0;
    bool owns_F = true;      // false when F is the caller's top-level source used in place
    bool borrowed = false;   // the whole level belongs to the caller (mgRunSubcycle)
};

struct Pending {  // a trace record whose scalar lives in a pinned slot until the next sync
    int rec;
    int slot;
    bool is_iters;   // the slot holds a sweep / iteration count (-> steps) instead of an error
};

class Cycle {
public:
    Cycle(int flags, mgTraceRec *recs, int max_recs) : flags_(flags), recs_(recs), max_recs_(max_recs) {}
    ~Cycle() { while (!stack_.empty()) pop(); }

    bool fused() const { return (flags_ & MG_RUN_FUSED) != 0; }
    bool quiet() const { return (flags_ & MG_RUN_QUIET) != 0; }

    void push(int N, double *borrowed_F = nullptr)  // linkedlist.cpp:7-44: three uninitialised arrays per node
    {
        Level l;
        l.N = N;
        if (dry_) { stack_.push_back(l); return; }     // parse-only run: sizes are all that is tracked
        l.U = mgGridAlloc(N);
        l.F = borrowed_F ? borrowed_F : mgGridAlloc(N);
        l.owns_F = borrowed_F == nullptr;
        if (fused()) l.W = mgGridAlloc(N);
        else l.D = mgGridAlloc(N);
        stack_.push_back(l);
    }
    void pop()  // linkedlist.cpp:46-69
    {
        Level &l = stack_.back();
        if (!dry_ && !l.borrowed) {
            mgGridFree(l.U); mgGridFree(l.W); mgGridFree(l.D);
            if (l.owns_F) mgGridFree(l.F);
        }
        stack_.pop_back();
        if (stack_.size() + depth_offset_ == 1) init_ = 0;
    }
    Level &top() { return stack_.back(); }
    size_t depth() const { return stack_.size(); }
    bool restart_top() const { return init_ == 0 && stack_.size() + depth_offset_ == 1; }  // :209, :252

    int next_slot()
    {
        if (slot_ == MG_SCALAR_SLOTS) harvest();
        return slot_++;
    }
    // after a sync, move slot values into the trace records
    void harvest()
    {
        mgSync();
        for (const Pending &p : pending_) {
            if (p.rec >= max_recs_ || !recs_) continue;
            const double v = *mgScalarSlot(p.slot);
            if (p.is_iters) recs_[p.rec].steps = (int)v;
            else recs_[p.rec].err = v;
        }
        pending_.clear();
        slot_ = 0;
    }
    int record(int node, int N, int steps, double err)
    {
        const int r = n_recs_++;
        if (recs_ && r < max_recs_) { recs_[r].node = node; recs_[r].N = N; recs_[r].steps = steps; recs_[r].err = err; }
        return r;
    }
    void defer(int rec, int slot, bool is_iters) { pending_.push_back({rec, slot, is_iters}); }
    int n_recs() const { return n_recs_; }

    // Error-trigger smoothing (:216-230 / :388-402): one sweep at a time, stop when two
    // successive errors differ by <= TRIGGER; the host needs the scalar after every sweep.
    int trigger_smooth(Level &l, double L)
    {
        double slope = TRIGGER + 1.0, previous = 0.0;
        l.step = 0;
        while (slope > TRIGGER) {
            if (fused()) {
                double *slot = mgScalarSlot(MG_SCALAR_SLOTS - 1);
                mgSmooth(l.N, L, l.U, l.F, 1, l.W, slot);
                std::swap(l.U, l.W);
                mgSync();
                l.smoothing_error = *slot;
            } else {
                doSmoothing(l.N, L, l.U, l.F, 1, &l.smoothing_error);
            }
            l.step += 1;
            if (l.step > 1) slope = std::fabs(l.smoothing_error - previous);
            previous = l.smoothing_error;
            if (mgLastErrorCode()) break;
        }
        return l.step;
    }

    void log_smoothing(const Level &l, int steps)
    {
        if (quiet()) return;
        printf("          ~Smoothing~\n");
        printf("Current Grid Size N = %d\n", l.N);
        printf("    Smoothing Steps = %d\n", steps);
        printf("              Error = %lf\n", l.smoothing_error);
    }

    // Collects the node sub-stream that stays at or below the current (small) level and runs it
    // as one kernel.  Returns the number of nodes consumed (0 = not applicable, < 0 = error).
    int try_tail(const std::vector<double> &tok, size_t &cur, size_t &pos, const std::vector<int> &ladder, int con_step,
                 int con_N, double L, bool quiet)
    {
        std::vector<int> kind, step, zero, Nop, nextN, option;
        std::vector<double> target;
        std::vector<int> sizes{top().N};                  // simulated stack below the entry level
        size_t c = cur, p = pos;
        int sim_init = init_;
        const size_t base_depth = stack_.size() + depth_offset_;
        const int max_ops = mgCoarseTailMaxOps();
        for (;;) {
            if (c >= tok.size()) break;
            const int node = (int)tok[c];
            if (node == 2) break;
            if (node == 1 && sizes.size() == 1) break;    // would prolong above the entry level
            if ((int)kind.size() >= max_ops) return 0;
            ++c;
            if (node == -1) {
                int st, nn;
                if (con_step == 0) { if (c >= tok.size()) return 0; st = (int)tok[c++]; } else st = con_step;
                if (con_N == 0) { if (c >= tok.size()) return 0; nn = (int)tok[c++]; }
                else { if (p + 1 >= ladder.size()) return 0; nn = ladder[++p]; }
                if (st == 0 || st < -1) return 0;         // FMG placeholder / zero-sweep quirk: leave it to the node-by-node path
                const bool restart = sim_init == 0 && base_depth + sizes.size() - 1 == 1;
                kind.push_back(-1); step.push_back(st); zero.push_back(restart ? 0 : 1); Nop.push_back(sizes.back());
                nextN.push_back(nn); option.push_back(0); target.push_back(0.0);
                sizes.push_back(nn);
            } else if (node == 0) {
                if (c + 2 > tok.size()) return 0;
                const double tg = tok[c++];
                const int opt = (int)tok[c++];
                kind.push_back(0); step.push_back(0); zero.push_back(0); Nop.push_back(sizes.back());
                nextN.push_back(0); option.push_back(opt); target.push_back(tg);
            } else if (node == 1) {
                int st;
                if (con_step == 0) { if (c >= tok.size()) return 0; st = (int)tok[c++]; } else st = con_step;
                if (st < -1) return 0;
                if (con_N != 0 && p > 0) --p;
                sizes.pop_back();
                if (base_depth + sizes.size() - 1 == 1) sim_init = 0;
                kind.push_back(1); step.push_back(st); zero.push_back(0); Nop.push_back(sizes.back());
                nextN.push_back(0); option.push_back(0); target.push_back(0.0);
            } else return 0;
        }
        if (sizes.size() != 1 || kind.size() < 2 || dry_) return 0;
        const int n = (int)kind.size();
        if (slot_ + 2 * n > MG_SCALAR_SLOTS - 1) harvest();
        const int slot0 = slot_;
        Level &l = top();
        const int rc = mgCoarseTail(L, l.U, l.F, n, kind.data(), step.data(), zero.data(), Nop.data(), nextN.data(), target.data(),
                                    option.data(), mgScalarSlot(slot0));
        if (rc == 10) return -10;
        if (rc != 0) return 0;                           // not representable: node-by-node
        slot_ += 2 * n;
        const int first_rec = n_recs_;
        for (int i = 0; i < n; ++i) {
            const int r = record(kind[i], Nop[i], kind[i] == 0 ? -1 : step[i], 0.0);
            if (kind[i] == 0) defer(r, slot0 + 2 * i + 1, true);
            else {
                if (step[i] != 0) defer(r, slot0 + 2 * i, false);
                if (step[i] < 0) defer(r, slot0 + 2 * i + 1, true);
            }
        }
        init_ = sim_init;
        if (!quiet) {                                    // the log needs the values now
            harvest();
            for (int i = 0; i < n; ++i) {
                const mgTraceRec *t = (recs_ && first_rec + i < max_recs_) ? &recs_[first_rec + i] : nullptr;
                if (kind[i] == 0) {
                    printf("          ~Exact Solver~\nCurrent Grid Size N = %d\n", Nop[i]);
                    if (option[i] == 1) printf("   Use Exact Solver = GaussSeidel Even / Odd\n");
                    printf("       Target Error = %.3e\n", target[i]);
                    continue;
                }
                if (kind[i] == 1) fputs("             *\n             |\nProlongation |\n             |\n             *\n", stdout);
                if (step[i] != 0) {
                    const double e = t ? t->err : *mgScalarSlot(slot0 + 2 * i);
                    const int st = t ? t->steps : (int)*mgScalarSlot(slot0 + 2 * i + 1);
                    printf("          ~Smoothing~\nCurrent Grid Size N = %d\n    Smoothing Steps = %d\n              Error = %lf\n", Nop[i], st, e);
                }
                if (kind[i] == -1) fputs("             *\n             |\n Restriction |\n             |\n             *\n", stdout);
            }
        }
        const int consumed = (int)(c - cur);
        cur = c;
        pos = p;
        return consumed > 0 ? consumed : 0;
    }

    bool dry_ = false;            // parse only (ranks that do not hold an agglomerated sub-cycle)
    size_t depth_offset_ = 0;     // levels of the caller's stack above stack_[0] (mgRunSubcycle)
    int flags_;
    mgTraceRec *recs_;
    int max_recs_;
    int n_recs_ = 0;
    int init_ = 1;  // linkedlist.h:41-44
    int slot_ = 0;
    std::vector<Level> stack_;
    std::vector<Pending> pending_;
};

const char *kRestrictArt = "             *\n             |\n Restriction |\n             |\n             *\n";
const char *kProlongArt = "             *\n             |\nProlongation |\n             |\n             *\n";

struct NodeStream {           // the token stream after the three header lines, with the ladder position
    const std::vector<double> &tok;
    size_t cur, pos;
    const std::vector<int> &ladder;
    int con_step, con_N;
    double L;
};

// The node loop of main() (MG_solver_CPU.cpp:158-426).  stop_depth == 0: until the code 2 / end
// of the stream.  stop_depth > 0 (sub-cycle of a larger driver): also stops, without reading the
// node, when a 1 node would pop the stack below stop_depth levels or when the code 2 comes up.
int interpret(Cycle &cy, NodeStream &s, size_t stop_depth)
{
    int rc = 0;
    int node = 0;
    const bool fused = cy.fused(), quiet = cy.quiet(), dry = cy.dry_;
    const bool sync_each_node = fused && !quiet;
    const std::vector<double> &tok = s.tok;
    size_t &cur = s.cur, &pos = s.pos;
    const std::vector<int> &ladder = s.ladder;
    const int con_step = s.con_step, con_N = s.con_N;
    const double L = s.L;
    mgTraceRec *recs = cy.recs_;
    const int max_recs = cy.max_recs_;
    (void)recs; (void)max_recs;
    auto have = [&](size_t n) { return cur + n <= tok.size(); };
    auto next_int = [&]() { return (int)tok[cur++]; };
    const int tail_max_N = (fused && !dry && !getenv("MG_NO_TAIL")) ? mgCoarseTailMaxN() : 0;
    // MG_TRACE=1: per-node time line of the cycle on stderr (device time between the nodes' last launches)
    struct Mark { int node, N; cudaEvent_t ev; };
    std::vector<Mark> marks;
    static const bool tracing = getenv("MG_TRACE") && atoi(getenv("MG_TRACE")) != 0;
    auto mark = [&](int node_, int N_) {
        if (!tracing || dry) return;
        Mark m{node_, N_, nullptr};
        cudaEventCreate(&m.ev);
        cudaEventRecord(m.ev, (cudaStream_t)mgStream());
        marks.push_back(m);
    };
    struct MarkDump {
        std::vector<Mark> &m;
        ~MarkDump()
        {
            if (m.size() > 1) cudaEventSynchronize(m.back().ev);
            for (size_t i = 1; i < m.size(); ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, m[i - 1].ev, m[i].ev);
                fprintf(stderr, "[mg trace] node %2d N=%-6d device %8.3f ms\n", m[i].node, m[i].N, ms);
            }
            for (Mark &k : m) cudaEventDestroy(k.ev);
        }
    } mark_dump{marks};
    mark(9, cy.top().N);
    while (have(1)) {                                                 // :158-160 (stops at EOF instead of re-running)
        if (tracing && !marks.empty() && !cy.stack_.empty()) mark(node, cy.top().N);
        // ---- coarse tail: the whole sub-cycle below a small level in one kernel (mg_tail.cu)
        if (tail_max_N && cy.top().N <= tail_max_N && ((int)tok[cur] == -1 || (int)tok[cur] == 0)) {
            const int took = cy.try_tail(tok, cur, pos, ladder, con_step, con_N, L, quiet);
            if (took < 0) { rc = -took; break; }
            if (took > 0) continue;
        }
        if (stop_depth > 0 && (int)tok[cur] == 1 && cy.depth() <= stop_depth) break;   // would prolong above the sub-cycle
        if (stop_depth > 0 && (int)tok[cur] == 2) break;             // the caller consumes the end marker
        node = next_int();
        if (node == 2) break;                                         // :162
        if (mgLastErrorCode()) { rc = 10; break; }
        const NodeRange nvtx_range(node, cy.stack_.empty() ? 0 : cy.top().N);

        if (node == -1) {                                             // :169-301
            int step, next_N;
            if (con_step == 0) { if (!have(1)) { rc = 3; break; } step = next_int(); } else step = con_step;
            if (con_N == 0) { if (!have(1)) { rc = 3; break; } next_N = next_int(); }
            else {
                if (pos + 1 >= ladder.size()) { rc = 4; break; }
                next_N = ladder[++pos];
            }
            if (step == 0) continue;                                  // :241-243, :296-299 (FMG placeholder)
            Level *l = &cy.top();
            const bool zero_init = !cy.restart_top();                 // :209-214 / :252-257
            const int fine_N = l->N;

            if (dry) {                                                // parse only: same stack moves, same record count
                cy.record(-1, fine_N, step > 0 ? step : 0, 0.0);
                cy.push(next_N);
                continue;
            }
            if (fused && step > 0) {
                const int slot = cy.next_slot();
                cy.push(next_N);
                l = &cy.stack_[cy.depth() - 2];
                Level &c = cy.top();
                double *resbuf = mgDownLeg(fine_N, L, l->U, l->W, l->F, step, zero_init, next_N, c.F, mgScalarSlot(slot));
                if (resbuf != l->U) std::swap(l->U, l->W);
                l->step = step;
                const int r = cy.record(-1, fine_N, step, 0.0);
                cy.defer(r, slot, false);
                if (sync_each_node) { cy.harvest(); l->smoothing_error = *mgScalarSlot(slot); }
                if (!quiet) { cy.log_smoothing(*l, step); fputs(kRestrictArt, stdout); }
                continue;
            }

            if (fused && step == -1) {
                // error trigger (:194-240): two sweeps per launch with both errors; the launch that ends the loop has also
                // restricted its residual, so the common two-sweep node is ONE launch and one synchronisation
                cy.push(next_N);
                l = &cy.stack_[cy.depth() - 2];
                Level &c = cy.top();
                int done = 0;
                double err = 0.0;
                double *resbuf = mgDownLegTrigger(fine_N, L, l->U, l->W, l->F, zero_init, next_N, c.F, &done, &err);
                if (resbuf) {
                    if (resbuf != l->U) std::swap(l->U, l->W);
                    l->step = done;
                    l->smoothing_error = err;
                } else {                                              // a size the streaming kernel does not serve: one sweep per call
                    if (zero_init) mgGridZero(l->N, l->U);
                    done = cy.trigger_smooth(*l, L);
                    mgDownLeg(fine_N, L, l->U, l->W, l->F, 0, 0, next_N, c.F, nullptr);
                }
                if (!quiet) { cy.log_smoothing(*l, done); fputs(kRestrictArt, stdout); }
                cy.record(-1, fine_N, done, l->smoothing_error);
                continue;
            }

            if (zero_init) mgGridZero(l->N, l->U);
            int done;
            if (step == -1) done = cy.trigger_smooth(*l, L);
            else { doSmoothing(l->N, L, l->U, l->F, step, &l->smoothing_error); l->step = done = step; }
            if (!quiet) cy.log_smoothing(*l, done);
            cy.record(-1, l->N, done, l->smoothing_error);

            if (fused) {
                // trigger mode in the fused driver: residual + negate + restrict without sweeps
                cy.push(next_N);
                l = &cy.stack_[cy.depth() - 2];
                mgDownLeg(fine_N, L, l->U, l->W, l->F, 0, 0, next_N, cy.top().F, nullptr);
            } else {
                getResidual(l->N, L, l->U, l->F, l->D);               // :239 / :268
                mgGridNegate(l->N, l->D);                             // :277-280
                cy.push(next_N);                                      // :283
                l = &cy.stack_[cy.depth() - 2];
                doRestriction(fine_N, l->D, next_N, cy.top().F);      // :287
            }
            if (!quiet) fputs(kRestrictArt, stdout);
        } else if (node == 0) {                                       // :305-325
            double target; int option;
            if (!have(2)) { rc = 3; break; }
            target = tok[cur++];
            option = next_int();
            Level &l = cy.top();
            if (dry) { cy.record(0, l.N, -1, 0.0); continue; }
            const int slot = cy.next_slot();
            mgExactSolve(l.N, L, l.U, l.F, target, option, mgScalarSlot(slot));
            const int r = cy.record(0, l.N, -1, 0.0);
            cy.defer(r, slot, true);
            if (!quiet) {
                printf("          ~Exact Solver~\n");
                printf("Current Grid Size N = %d\n", l.N);
                if (option == 0) printf("   Use Exact Solver = Inverse Matrix\n");
                if (option == 1) printf("   Use Exact Solver = GaussSeidel Even / Odd\n");
                printf("       Target Error = %.3e\n", target);
            }
        } else if (node == 1) {                                       // :329-424
            int step;
            if (con_step == 0) { if (!have(1)) { rc = 3; break; } step = next_int(); } else step = con_step;
            if (con_N != 0 && pos > 0) --pos;
            if (cy.depth() < 2) { rc = 5; break; }                    // reference: null prevNode
            Level coarse = cy.top();
            Level *l = &cy.stack_[cy.depth() - 2];
            if (dry) {
                cy.pop();
                cy.record(1, cy.top().N, step > 0 ? step : 0, 0.0);
                continue;
            }

            if (fused && step == -1) {
                int done = 0;
                double err = 0.0;
                double *resbuf = mgUpLegTrigger(coarse.N, coarse.U, l->N, L, l->U, l->W, l->F, &done, &err);   // :350-408, two sweeps per launch
                if (resbuf) {
                    if (resbuf != l->U) std::swap(l->U, l->W);
                    if (!quiet) fputs(kProlongArt, stdout);
                    cy.pop();
                    l = &cy.top();
                    l->step = done;
                    l->smoothing_error = err;
                    cy.record(1, l->N, done, err);
                    if (!quiet) cy.log_smoothing(*l, done);
                    continue;
                }
            }
            if (fused) {
                const int slot = step > 0 ? cy.next_slot() : -1;
                double *resbuf = mgUpLeg(coarse.N, coarse.U, l->N, L, l->U, l->W, l->F, step > 0 ? step : 0,
                                         slot >= 0 ? mgScalarSlot(slot) : nullptr);
                if (resbuf != l->U) std::swap(l->U, l->W);
                if (!quiet) fputs(kProlongArt, stdout);
                cy.pop();                                             // stream-ordered: safe to recycle the coarse grids
                l = &cy.top();
                int done = 0;
                if (step == -1) done = cy.trigger_smooth(*l, L);
                else if (step > 0) { l->step = done = step; }
                else if (step < -1) {                                 // :410-421: doSmoothing with a negative count = 0 sweeps + error
                    doSmoothing(l->N, L, l->U, l->F, step, &l->smoothing_error);
                    l->step = done = step;
                }
                const int r = cy.record(1, l->N, done, step < 0 ? l->smoothing_error : 0.0);
                if (slot >= 0) {
                    cy.defer(r, slot, false);
                    if (sync_each_node) { cy.harvest(); l->smoothing_error = *mgScalarSlot(slot); }
                }
                if (!quiet && step != 0) cy.log_smoothing(*l, done);
                continue;
            }

            double *tmp = mgGridAlloc(l->N);                          // :353
            doProlongation(coarse.N, coarse.U, l->N, tmp);            // :354
            if (!quiet) fputs(kProlongArt, stdout);
            cy.pop();                                                 // :363
            l = &cy.top();
            doGridAddition(l->N, l->U, tmp);                          // :368
            mgGridFree(tmp);                                          // :371
            int done = 0;
            if (step == -1) done = cy.trigger_smooth(*l, L);
            else if (step != 0) { doSmoothing(l->N, L, l->U, l->F, step, &l->smoothing_error); l->step = done = step; }
            if (!quiet && step != 0) cy.log_smoothing(*l, done);
            cy.record(1, l->N, done, l->smoothing_error);
        } else {
            rc = 6;
            break;
        }
    }
    return rc;
}

// U0_top / analytic_top: the problem plug point.  U0_top (device, N_max^2) is an initial grid INCLUDING its
// boundary values -- non-zero Dirichlet data; the cycle then starts the way the reference restarts (:209-211): the first -1
// node of the top level keeps U instead of zeroing it, the boundary is carried through every sweep (:587-599 touch interior
// points only) and enters the residual.  analytic_top replaces getAnalytic in the final report (:434-445).
int run(const char *path, int flags, const double *F_top, double *U_top, mgTraceRec *recs, int max_recs,
        mgCycleResult *res, const double *U0_top = nullptr, const double *analytic_top = nullptr)
{
    std::ifstream f(path);
    if (!f.is_open()) {
        fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path);
        return 1;
    }
    // The whole file as whitespace-delimited numeric tokens (what `ifstream >>` sees), so that the
    // interpreter can look ahead for a coarse tail.
    std::vector<double> tok;
    for (double d; f >> d;) tok.push_back(d);
    size_t cur = 0;
    auto have = [&](size_t n) { return cur + n <= tok.size(); };
    auto next_int = [&]() { return (int)tok[cur++]; };
    if (!have(7)) return 2;
    double L, min_x, min_y;
    int con_step, con_N, N_max, N_min;
    L = tok[cur++]; min_x = tok[cur++]; min_y = tok[cur++];   // :103
    con_step = next_int(); con_N = next_int();                // :106
    N_max = next_int(); N_min = next_int();                   // :109

    std::vector<int> ladder;    // :111-146
    if (con_N == 1) for (int n = N_max; n >= N_min; n /= 2) ladder.push_back(n);
    if (con_N == 2) for (int n = N_max; n >= N_min; --n) ladder.push_back(n);
    size_t pos = 0;

    Cycle cy(flags, recs, max_recs);
    const bool quiet = cy.quiet();

    if ((flags & MG_RUN_SKIP_SOURCE) && F_top) {
        cy.push(N_max, const_cast<double *>(F_top));                  // the operators never write F
    } else {
        cy.push(N_max);                                               // :149
        getSource(N_max, L, cy.top().F, min_x, min_y);                // :153 (outside the timer)
    }
    if (U0_top) {
        mgGridCopy(N_max, cy.top().U, U0_top);
        cy.init_ = 0;                                                 // "more like a restart method" (:210)
    }
    mgSync();

    cudaStream_t stream = (cudaStream_t)mgStream();
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    const int launches0 = mgKernelLaunches();
    const auto wall0 = std::chrono::steady_clock::now();
    cudaEventRecord(ev0, stream);                                     // :156

    NodeStream ns{tok, cur, pos, ladder, con_step, con_N, L};
    int rc = interpret(cy, ns, 0);
    cudaEventRecord(ev1, stream);                                     // :429
    cy.harvest();
    const auto wall1 = std::chrono::steady_clock::now();
    if (mgLastErrorCode() && rc == 0) rc = 10;

    if (res) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        res->time_ms = ms;
        res->wall_ms = std::chrono::duration<double, std::milli>(wall1 - wall0).count();
        res->launches = mgKernelLaunches() - launches0;
        res->n_recs = cy.n_recs() < max_recs ? cy.n_recs() : max_recs;
        res->N = cy.stack_.empty() ? 0 : cy.stack_.front().N;
        res->mg_error = 0.0;
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);

    if (rc == 0 && !cy.stack_.empty()) {
        Level &l = cy.top();
        if (res && !(flags & MG_RUN_NO_FINAL_ERROR))                  // :434-445
            res->mg_error = analytic_top ? mgMeanAbsDiff(l.N, analytic_top, l.U) : mgAnalyticError(l.N, L, l.U, min_x, min_y);
        if (U_top) { mgGridCopy(l.N, U_top, l.U); mgSync(); }
        if (!quiet && res) {
            printf("\n\n");
            printf("===== Final Result =====\n");
            printf("    Error = %lf\n", res->mg_error);
            printf("Time Used = %lf (ms)\n", res->wall_ms);
        }
    }
    return rc;
}

}  // namespace

// Runs, on ONE level owned by the caller (N, *U, *W, F: device grids), the node sub-stream that
// starts at tok[*cur] and returns to that level (stops before the 1 node that would prolong above
// it, before the code 2, or at the end).  Used by the slab driver for the levels agglomerated on
// rank 0; execute == 0 parses only (ranks that hold no data) so that all ranks stay in step.
// depth_offset = number of levels of the caller's stack above this one; *init_io = the level
// stack's restart flag (linkedlist.h:41-44).
// Scalars of sub-cycles run with MG_RUN_DEFER_HARVEST: still in the pinned slot ring, owed to g_carry_recs.
namespace {
std::vector<Pending> g_carry;
int g_carry_slot = 0, g_carry_max = 0;
mgTraceRec *g_carry_recs = nullptr;
}  // namespace

extern "C" void mgSubcycleHarvest(void)
{
    if (g_carry.empty()) { g_carry_slot = 0; return; }
    Cycle cy(MG_RUN_FUSED | MG_RUN_QUIET, g_carry_recs, g_carry_max);
    cy.pending_.swap(g_carry);
    cy.harvest();
    g_carry_slot = 0;
}

extern "C" int mgRunSubcycle(const double *tok, int n_tok, int *cur, int *pos, const int *ladder, int n_ladder, int con_step,
                             int con_N, double L, int N, double **U, double **W, double *F, int depth_offset, int *init_io,
                             int flags, mgTraceRec *recs, int max_recs, int *n_recs_io, int execute)
{
    const bool defer = (flags & MG_RUN_DEFER_HARVEST) && execute;
    if (execute && !g_carry.empty() && (!defer || recs != g_carry_recs)) mgSubcycleHarvest();   // another caller's leftovers
    const std::vector<double> tokens(tok, tok + n_tok);
    const std::vector<int> lad(ladder, ladder + n_ladder);
    Cycle cy((flags & ~MG_RUN_DEFER_HARVEST) | MG_RUN_FUSED | MG_RUN_QUIET, recs, max_recs);
    cy.dry_ = !execute;
    cy.depth_offset_ = (size_t)depth_offset;
    cy.init_ = *init_io;
    cy.n_recs_ = *n_recs_io;
    if (defer) {                // continue the slot ring where the previous deferred sub-cycle stopped
        cy.pending_.swap(g_carry);
        cy.slot_ = g_carry_slot;
    }
    Level l;
    l.N = N;
    l.borrowed = true;
    if (execute) { l.U = *U; l.W = *W; l.F = F; }
    cy.stack_.push_back(l);
    NodeStream ns{tokens, (size_t)*cur, (size_t)*pos, lad, con_step, con_N, L};
    const int rc = interpret(cy, ns, 1);
    if (defer) {
        g_carry.swap(cy.pending_);
        g_carry_slot = cy.slot_;
        g_carry_recs = recs;
        g_carry_max = max_recs;
    } else cy.harvest();
    if (execute && !cy.stack_.empty()) { *U = cy.stack_[0].U; *W = cy.stack_[0].W; }
    *cur = (int)ns.cur;
    *pos = (int)ns.pos;
    *init_io = cy.init_;
    *n_recs_io = cy.n_recs_;
    cy.stack_.clear();          // nothing here is owned by the Cycle
    return rc;
}

extern "C" int mgRunCycleFile(const char *path, int flags, const double *F_top, double *U_top, mgTraceRec *recs,
                              int max_recs, mgCycleResult *res)
{
    if (mgLastErrorCode()) return 10;
    return run(path, flags, F_top, U_top, recs, max_recs, res);
}

extern "C" int mgRunCycleFileEx(const char *path, int flags, const mgProblem *prob, double *U_top, mgTraceRec *recs, int max_recs,
                                mgCycleResult *res)
{
    if (mgLastErrorCode()) return 10;
    if (!prob) return run(path, flags, nullptr, U_top, recs, max_recs, res);
    if (prob->F_top) flags |= MG_RUN_SKIP_SOURCE;
    return run(path, flags, prob->F_top, U_top, recs, max_recs, res, prob->U0_top, prob->analytic_top);
}

extern "C" int mgRunCycleFileHostEx(const char *path, int flags, const double *F_host, const double *U0_host, const double *analytic_host,
                                    double *U_host, mgTraceRec *recs, int max_recs, mgCycleResult *res)
{
    if (mgLastErrorCode()) return 10;
    std::ifstream f(path);
    if (!f.is_open()) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
    double L, mx, my; int cs, cn, N_max, N_min;
    f >> L >> mx >> my >> cs >> cn >> N_max >> N_min;
    if (!f) return 2;
    f.close();
    double *dF = nullptr, *dU0 = nullptr, *dA = nullptr, *dU = nullptr;
    if (F_host) { dF = mgGridAlloc(N_max); mgGridUpload(N_max, dF, F_host); }
    if (U0_host) { dU0 = mgGridAlloc(N_max); mgGridUpload(N_max, dU0, U0_host); }
    if (analytic_host) { dA = mgGridAlloc(N_max); mgGridUpload(N_max, dA, analytic_host); }
    if (U_host) dU = mgGridAlloc(N_max);
    mgProblem prob{dF, dU0, dA};
    const int rc = mgRunCycleFileEx(path, flags, &prob, dU, recs, max_recs, res);
    if (rc == 0 && U_host) mgGridDownload(N_max, dU, U_host);
    mgGridFree(dF); mgGridFree(dU0); mgGridFree(dA); mgGridFree(dU);
    return rc;
}

extern "C" int mgRunCycleFileHost(const char *path, int flags, const double *F_host, double *U_host, mgTraceRec *recs,
                                  int max_recs, mgCycleResult *res)
{
    if (mgLastErrorCode()) return 10;
    // peek N_max for the staging grids
    std::ifstream f(path);
    if (!f.is_open()) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
    double L, mx, my; int cs, cn, N_max, N_min;
    f >> L >> mx >> my >> cs >> cn >> N_max >> N_min;
    if (!f) return 2;
    f.close();
    double *dF = nullptr, *dU = nullptr;
    if (F_host) { dF = mgGridAlloc(N_max); mgGridUpload(N_max, dF, F_host); flags |= MG_RUN_SKIP_SOURCE; }
    if (U_host) dU = mgGridAlloc(N_max);
    const int rc = run(path, flags, dF, dU, recs, max_recs, res);
    if (rc == 0 && U_host) mgGridDownload(N_max, dU, U_host);
    mgGridFree(dF);
    mgGridFree(dU);
    return rc;
}

// n independent problems (same cycle file, different sources) with HOST buffers, double-buffered:
// while the cycle of problem i runs, the source of problem i+1 is uploaded and the solution of
// problem i-1 is downloaded on two copy streams (PCIe is full duplex, and a 2 GiB copy takes ten
// times longer than the cycle at N = 16384).  Results are bit-identical to n mgRunCycleFileHost calls.
extern "C" int mgRunCycleFileHostBatch(const char *path, int flags, int n, const double *const *F_hosts, double *const *U_hosts,
                                       mgCycleResult *res)
{
    if (mgLastErrorCode()) return 10;
    if (n < 1 || !F_hosts || !U_hosts) return 11;
    for (int i = 0; i < n; ++i)
        if (!F_hosts[i] || !U_hosts[i]) return 11;
    std::ifstream f(path);
    if (!f.is_open()) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
    double L, mx, my; int cs, cn, N_max, N_min;
    f >> L >> mx >> my >> cs >> cn >> N_max >> N_min;
    if (!f) return 2;
    f.close();
    const size_t bytes = (size_t)N_max * N_max * sizeof(double);
    cudaStream_t compute = (cudaStream_t)mgStream(), up = nullptr, down = nullptr;
    cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking);
    double *dF[2] = {mgGridAlloc(N_max), n > 1 ? mgGridAlloc(N_max) : nullptr};
    double *dU[2] = {mgGridAlloc(N_max), n > 1 ? mgGridAlloc(N_max) : nullptr};
    std::vector<cudaEvent_t> uploaded((size_t)n), downloaded((size_t)n);
    for (int i = 0; i < n; ++i) {
        cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&downloaded[i], cudaEventDisableTiming);
    }
    int rc = (dF[0] && dU[0] && (n == 1 || (dF[1] && dU[1]))) ? 0 : 12;
    auto upload = [&](int i) {      // the previous reader of dF[i % 2], cycle i-2, has been waited for on the host
        cudaMemcpyAsync(dF[i % 2], F_hosts[i], bytes, cudaMemcpyHostToDevice, up);
        cudaEventRecord(uploaded[i], up);
    };
    mgSync();                       // the staging grids may be recycled blocks with work still queued on them
    if (rc == 0) upload(0);
    for (int i = 0; i < n && rc == 0; ++i) {
        if (i + 1 < n) upload(i + 1);
        cudaStreamWaitEvent(compute, uploaded[i], 0);
        if (i >= 2) cudaStreamWaitEvent(compute, downloaded[i - 2], 0);      // dU[i % 2] is free again
        rc = run(path, flags | MG_RUN_SKIP_SOURCE, dF[i % 2], dU[i % 2], nullptr, 0, res ? res + i : nullptr);   // returns synchronised
        if (rc) break;
        cudaMemcpyAsync(U_hosts[i], dU[i % 2], bytes, cudaMemcpyDeviceToHost, down);
        cudaEventRecord(downloaded[i], down);
    }
    cudaStreamSynchronize(up);
    cudaStreamSynchronize(down);
    if (rc == 0 && cudaGetLastError() != cudaSuccess) rc = 13;
    for (int i = 0; i < n; ++i) { cudaEventDestroy(uploaded[i]); cudaEventDestroy(downloaded[i]); }
    cudaStreamDestroy(up);
    cudaStreamDestroy(down);
    for (int k = 0; k < 2; ++k) { mgGridFree(dF[k]); mgGridFree(dU[k]); }
    return rc;
}
