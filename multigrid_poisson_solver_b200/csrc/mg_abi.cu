// mg_abi.cu -- extern "C" surface of libmgb200 (include/mg_abi.h): context, pooled grids and
// the reference's eight operators over device-resident grids.  No CPU fallback anywhere:
// without a usable CUDA device every entry point reports through mgLastError and does nothing.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mg_abi.h"
#include "mg_fused.h"
#include "mg_kernels.h"

namespace mg {

Context &ctx()
{
    static Context c;
    return c;
}

void fail(int code, const std::string &msg)
{
    Context &c = ctx();
    if (c.err_code == 0) {
        c.err_code = code;
        c.err_msg = msg;
    }
    fprintf(stderr, "[ ERROR ]: libmgb200: %s\n", msg.c_str());
}

bool check(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    fail((int)e, std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

bool ensure_ready()
{
    if (ctx().ready) return true;
    fail(-1, "no CUDA context: call mgInit(device) first (there is no CPU fallback)");
    return false;
}

double *scratch_grid(size_t elems)
{
    Context &c = ctx();
    if (elems > c.scratch_elems) {
        if (c.scratch) {
            check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
            check(cudaFree(c.scratch), "cudaFree scratch");
            c.scratch = nullptr;
            c.scratch_elems = 0;
        }
        if (!check(cudaMalloc(&c.scratch, elems * sizeof(double)), "cudaMalloc scratch")) return nullptr;
        c.scratch_elems = elems;
    }
    return c.scratch;
}

double *partials_buf(size_t elems)
{
    Context &c = ctx();
    if (elems > c.partials_elems) {
        if (c.partials) {
            check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
            check(cudaFree(c.partials), "cudaFree partials");
        }
        const size_t want = elems < 65536 ? 65536 : elems;
        if (!check(cudaMalloc(&c.partials, want * sizeof(double)), "cudaMalloc partials")) return nullptr;
        c.partials_elems = want;
    }
    return c.partials;
}

double *slot_device_ptr(double *host_slot)
{
    Context &c = ctx();
    if (!host_slot) return nullptr;
    if (host_slot < c.slots_host || host_slot >= c.slots_host + MG_SCALAR_SLOTS) {
        fail(-2, "error_slot must come from mgScalarSlot()");
        return nullptr;
    }
    return c.slots_dev + (host_slot - c.slots_host);
}

static size_t grid_bytes(int N)
{
    return (size_t)N * (size_t)N * sizeof(double);
}

void *pool_alloc(size_t bytes)
{
    Context &c = ctx();
    bytes = (bytes + 255) / 256 * 256;
    void *p = nullptr;
    auto it = c.free_lists.find(bytes);
    if (it != c.free_lists.end() && !it->second.empty()) {
        p = it->second.back();
        it->second.pop_back();
        c.pooled_bytes -= bytes;
    } else if (!check(cudaMalloc(&p, bytes), "cudaMalloc grid")) {
        return nullptr;
    }
    c.live[p] = bytes;
    return p;
}

void pool_free(void *ptr)
{
    if (!ptr) return;
    Context &c = ctx();
    auto it = c.live.find(ptr);
    if (it == c.live.end()) { fail(-7, "pool_free: pointer was not allocated by the pool"); return; }
    // stream-ordered reuse: every consumer is queued on the same stream, so the block can be handed out again at once
    c.free_lists[it->second].push_back(ptr);
    c.pooled_bytes += it->second;
    c.live.erase(it);
}

}  // namespace mg

using namespace mg;

extern "C" {

// ------------------------------------------------------------------ lifecycle
int mgInit(int device)
{
    Context &c = ctx();
    if (c.ready) {
        if (c.device == device) return 0;
        fail(-3, "mgInit: context already bound to another device");
        return -3;
    }
    c.err_code = 0;
    c.err_msg.clear();
    int count = 0;
    if (!check(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count == 0) {
        fail(-4, "mgInit: no CUDA device (libmgb200 has no CPU fallback)");
        return -4;
    }
    if (!check(cudaSetDevice(device), "cudaSetDevice")) return -5;
    cudaDeviceProp prop;
    if (!check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return -5;
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    bool ok = check(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    ok = ok && check(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate (comm)");
    ok = ok && check(cudaMalloc(&c.counters, 64 * sizeof(unsigned int)), "cudaMalloc counters");
    ok = ok && check(cudaMemset(c.counters, 0, 64 * sizeof(unsigned int)), "cudaMemset counters");
    ok = ok && check(cudaMalloc(&c.gs_iters, 16 * sizeof(int)), "cudaMalloc gs_iters");
    ok = ok && check(cudaMemset(c.gs_iters, 0, 16 * sizeof(int)), "cudaMemset gs_iters");
    ok = ok && check(cudaMalloc(&c.dev_scalar, 16 * sizeof(double)), "cudaMalloc dev_scalar");
    ok = ok && check(cudaHostAlloc(&c.slots_host, MG_SCALAR_SLOTS * sizeof(double), cudaHostAllocMapped), "cudaHostAlloc slots");
    ok = ok && check(cudaHostGetDevicePointer(&c.slots_dev, c.slots_host, 0), "cudaHostGetDevicePointer");
    if (!ok) return c.err_code ? c.err_code : -6;
    memset(c.slots_host, 0, MG_SCALAR_SLOTS * sizeof(double));
    partials_buf(65536);
    fused_init();
    c.ready = c.err_code == 0;
    return c.err_code;
}

void mgShutdown(void)
{
    Context &c = ctx();
    if (!c.ready) return;
    cudaStreamSynchronize(c.stream);
    cudaStreamSynchronize(c.comm_stream);
    dist_release_on_shutdown();
    for (auto &kv : c.free_lists)
        for (void *p : kv.second) cudaFree(p);
    for (auto &kv : c.live) cudaFree(kv.first);
    c.free_lists.clear();
    c.live.clear();
    for (auto &kv : c.restrict_tables) { cudaFree(kv.second.lo); cudaFree(kv.second.w); }
    for (auto &kv : c.prolong_tables) {
        cudaFree(kv.second.row_cell); cudaFree(kv.second.col_cell); cudaFree(kv.second.row_w); cudaFree(kv.second.col_w);
        cudaFree(kv.second.row_info);
    }
    c.restrict_tables.clear();
    c.prolong_tables.clear();
    cudaFree(c.scratch); cudaFree(c.partials); cudaFree(c.counters); cudaFree(c.gs_iters); cudaFree(c.dev_scalar);
    cudaFreeHost(c.slots_host);
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.comm_stream);
    c = Context();
}

int mgSetTileMaxN(int n) { return set_tile_max_n(n); }

int mgSegmentPlan(int rows, int n_strips, int resident_warps, int lead_rows, int subset, int *out, int max_out)
{
    // host-only: usable without a GPU (tests of the task geometry)
    if (rows < 1 || n_strips < 1 || resident_warps < 1 || lead_rows < 0 || subset < 0 || subset > 2) return -1;
    const std::vector<int> plan = segment_plan(rows, n_strips, resident_warps, lead_rows, subset);
    if ((int)plan.size() > max_out) return -2;
    for (size_t i = 0; i < plan.size(); ++i) out[i] = plan[i];
    return (int)plan.size() / 2;
}

int mgLastErrorCode(void) { return ctx().err_code; }
const char *mgLastError(void) { return ctx().err_msg.c_str(); }
void mgClearError(void) { ctx().err_code = 0; ctx().err_msg.clear(); }
void mgSync(void) { if (ensure_ready()) check(cudaStreamSynchronize(ctx().stream), "cudaStreamSynchronize"); }
void *mgStream(void) { return (void *)ctx().stream; }
int mgKernelLaunches(void) { return (int)ctx().launches; }
double *mgScalarSlot(int index)
{
    if (!ensure_ready() || index < 0 || index >= MG_SCALAR_SLOTS) return nullptr;
    return ctx().slots_host + index;
}

double *mgGridAlloc(int N)
{
    if (!ensure_ready()) return nullptr;
    return (double *)pool_alloc(grid_bytes(N));
}

void mgGridFree(double *grid)
{
    if (!grid || !ensure_ready()) return;
    pool_free(grid);
}

void mgGridZero(int N, double *grid)
{
    if (ensure_ready()) check(cudaMemsetAsync(grid, 0, (size_t)N * N * sizeof(double), ctx().stream), "cudaMemsetAsync");
}
void mgGridNegate(int N, double *grid) { if (ensure_ready()) launch_negate(N, grid); }
void mgGridUpload(int N, double *dev, const double *host)
{
    if (ensure_ready()) check(cudaMemcpyAsync(dev, host, (size_t)N * N * sizeof(double), cudaMemcpyHostToDevice, ctx().stream), "H2D");
}
void mgGridDownload(int N, const double *dev, double *host)
{
    if (!ensure_ready()) return;
    check(cudaMemcpyAsync(host, dev, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream), "D2H");
    check(cudaStreamSynchronize(ctx().stream), "cudaStreamSynchronize");
}
void mgGridCopy(int N, double *dst, const double *src)
{
    if (ensure_ready()) check(cudaMemcpyAsync(dst, src, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream), "D2D");
}

// ------------------------------------------------------------------ the eight operators
void getSource(int N, double L, double *F, double min_x, double min_y)
{
    if (ensure_ready()) launch_source(N, L, F, min_x, min_y, false);
}

void getBoundary(int N, double L, double *F, double min_x, double min_y)
{
    (void)L; (void)min_x; (void)min_y;
    mgGridZero(N, F);  // homogeneous Dirichlet: memset + zero edges (:502-519)
}

void getAnalytic(int N, double L, double *U, double min_x, double min_y)
{
    if (ensure_ready()) launch_source(N, L, U, min_x, min_y, true);
}

void getResidual(int N, double L, double *U, double *F, double *D)
{
    if (!ensure_ready()) return;
    launch_residual(N, spacing(N, L).inv_h2, U, F, D);
}

void doSmoothing(int N, double L, double *U, double *F, int step, double *error)
{
    if (!ensure_ready()) return;
    Context &c = ctx();
    double *work = scratch_grid((size_t)N * N);
    if (!work) return;
    double *res = smooth_out_of_place(N, L, U, work, U, F, step, false, c.dev_scalar, nullptr);
    if (res != U) mgGridCopy(N, U, res);
    if (error) {
        check(cudaMemcpyAsync(error, c.dev_scalar, sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H error");
        check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
    }
}

void doRestriction(int N, double *U_f, int M, double *U_c)
{
    if (ensure_ready()) launch_restrict(N, U_f, M, U_c);
}

void doProlongation(int N, double *U_c, int M, double *U_f)
{
    if (ensure_ready()) launch_prolong(N, U_c, M, U_f, nullptr);
}

void doGridAddition(int N, double *U1, double *U2)
{
    if (ensure_ready()) launch_add(N, U1, U2);
}

void mgExactSolve(int N, double L, double *U, double *F, double target_error, int option, double *iters_slot)
{
    if (!ensure_ready()) return;
    double *slot = slot_device_ptr(iters_slot);
    if (option == 0) {
        launch_inverse_matrix(N, L, U, F);
        if (iters_slot) *iters_slot = -1.0;
    } else if (option == 1) {
        launch_gauss_seidel(N, L, U, F, target_error, slot);
    }
    // any other option: the reference does nothing (:627-638)
}

void doExactSolver(int N, double L, double *U, double *F, double target_error, int option)
{
    mgExactSolve(N, L, U, F, target_error, option, nullptr);
}

int mgLastExactSolverIterations(void)
{
    if (!ensure_ready()) return -1;
    int it = -1;
    check(cudaMemcpyAsync(&it, ctx().gs_iters, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream), "D2H iters");
    check(cudaStreamSynchronize(ctx().stream), "cudaStreamSynchronize");
    return it;
}

double mgAnalyticError(int N, double L, const double *U, double min_x, double min_y)
{
    if (!ensure_ready()) return -1.0;
    Context &c = ctx();
    double *ana = scratch_grid((size_t)N * N);
    if (!ana) return -1.0;
    launch_source(N, L, ana, min_x, min_y, true);
    launch_mean_abs_diff((size_t)N * N, ana, U, (double)N * (double)N, c.dev_scalar + 1);
    double out = -1.0;
    check(cudaMemcpyAsync(&out, c.dev_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H");
    check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
    return out;
}

double mgMeanAbsDiff(int N, const double *A, const double *B)
{
    if (!ensure_ready()) return -1.0;
    Context &c = ctx();
    launch_mean_abs_diff((size_t)N * N, A, B, (double)N * (double)N, c.dev_scalar + 1);
    double out = -1.0;
    check(cudaMemcpyAsync(&out, c.dev_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H");
    check(cudaStreamSynchronize(c.stream), "cudaStreamSynchronize");
    return out;
}

// ------------------------------------------------------------------ fused operators
void mgSmooth(int N, double L, const double *U_in, double *F, int step, double *U_out, double *error_slot)
{
    if (!ensure_ready()) return;
    // out-of-place contract: U_in is never written; the passes alternate between U_out and a
    // scratch grid, ordered so that the last one lands in U_out
    Context &c = ctx();
    double *err_dev = error_slot ? c.dev_scalar : nullptr;
    const int passes = smooth_pass_count(N, step);
    if (passes == 0) mgGridCopy(N, U_out, U_in);
    double *work = passes > 1 ? scratch_grid((size_t)N * N) : nullptr;
    if (passes > 1 && !work) return;
    const double *in = passes == 0 ? U_out : U_in;
    double *a = (passes % 2) ? U_out : work, *b = (passes % 2) ? work : U_out;
    double *res = smooth_out_of_place(N, L, in, a, b, F, step, false, err_dev, slot_device_ptr(error_slot));
    if (passes > 0 && res != U_out) fail(-8, "mgSmooth: internal pass parity error");
}

double *mgDownLeg(int N, double L, double *U, double *U_work, double *F, int step, int zero_init, int M, double *F_c,
                  double *error_slot)
{
    if (!ensure_ready()) return U;
    return down_leg(N, L, U, U_work, F, step, zero_init != 0, M, F_c, slot_device_ptr(error_slot));
}

double *mgUpLeg(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, double *F, int step,
                double *error_slot)
{
    if (!ensure_ready()) return U_f;
    return up_leg(Nc, U_c, N, L, U_f, U_work, F, step, slot_device_ptr(error_slot));
}

double *mgDownLegTrigger(int N, double L, double *U, double *U_work, double *F, int zero_init, int M, double *F_c, int *steps,
                         double *error)
{
    if (!ensure_ready() || !trigger_fusable_down(N, M)) return nullptr;
    return down_leg_trigger(N, L, U, U_work, F, zero_init != 0, M, F_c, steps, error);
}

double *mgUpLegTrigger(int Nc, const double *U_c, int N, double L, double *U_f, double *U_work, double *F, int *steps, double *error)
{
    if (!ensure_ready() || !trigger_fusable_up(Nc, N)) return nullptr;
    return up_leg_trigger(Nc, U_c, N, L, U_f, U_work, F, steps, error);
}

int mgPrint2File(int N, const double *U_host, const char *file_name)
{
    FILE *out = fopen(file_name, "w");
    if (!out) return 1;
    // rows from the top (j = N-1) down, "%lf" fields (doPrint2File, MG_solver_CPU.cpp:735-754)
    for (int j = N - 1; j >= 0; --j) {
        const double *row = U_host + (size_t)N * j;
        for (int i = 0; i < N; ++i) fprintf(out, i == N - 1 ? "%lf\n" : "%lf,", row[i]);
    }
    fclose(out);
    return 0;
}

}  // extern "C"
