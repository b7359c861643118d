// mg_context.h -- process-wide context of libmgb200: stream, pooled device grids,
// scratch arena, pinned scalar slots, cached 1-D transfer tables.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace mg {

// 1-D interpolation map of doRestriction (MG_solver_CPU.cpp:661-666): for coarse index t,
// lo[t] = (int)floor(t*h_c/h_f), w[t] = fmod(t*h_c, h_f)/h_f.
struct RestrictTable {
    int *lo = nullptr;
    double *w = nullptr;
};

// 1-D ownership map of doProlongation (MG_solver_CPU.cpp:688-718) in gather form: fine index t
// takes coarse cell `cell[t]` with weights lo_w = (cell+1)*c_dx - f, hi_w = f - cell*c_dx.
// Rows and columns differ only in the patched last line.
struct ProlongTable {
    int *row_cell = nullptr, *col_cell = nullptr;
    double2 *row_w = nullptr, *col_w = nullptr;  // {lo_w, hi_w}
    double4 *row_info = nullptr;                 // per fine row {lo_w, hi_w, cell (exact as a double), 0}: one 32-byte bulk copy per row
};

struct Context {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;   // highest priority; the slab driver issues all communication here
    bool ready = false;

    // error channel
    int err_code = 0;
    std::string err_msg;

    // pooled grids: bytes -> free list ; live pointer -> bytes
    std::map<size_t, std::vector<void *>> free_lists;
    std::unordered_map<void *, size_t> live;
    size_t pooled_bytes = 0;

    // scratch arena (grown on demand, never shrunk)
    double *scratch = nullptr;
    size_t scratch_elems = 0;
    double *partials = nullptr;       // per-CTA partial sums of reductions
    size_t partials_elems = 0;
    unsigned int *counters = nullptr; // last-block-done tickets (zeroed, self-resetting)
    int *gs_iters = nullptr;          // device int: iterations of the last exact solve
    double *dev_scalar = nullptr;     // device double: reduction result when no slot is given

    // pinned, device-visible scalars written by kernels (error slots)
    double *slots_host = nullptr;
    double *slots_dev = nullptr;

    std::map<std::pair<int, int>, RestrictTable> restrict_tables;  // key (N fine, M coarse)
    std::map<std::pair<int, int>, ProlongTable> prolong_tables;    // key (N coarse, M fine)

    long long launches = 0;
};

Context &ctx();
bool ensure_ready();
void fail(int code, const std::string &msg);
bool check(cudaError_t e, const char *what);

void *pool_alloc(size_t bytes);   // size-keyed, stream-ordered pool behind mgGridAlloc
void pool_free(void *ptr);
double *scratch_grid(size_t elems);
double *partials_buf(size_t elems);
double *slot_device_ptr(double *host_slot);  // pinned host slot -> device alias, or nullptr

}  // namespace mg
