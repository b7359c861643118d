// mg_dist.cu -- row-slab multi-GPU cycle driver.
//
// Fine levels are partitioned into contiguous row slabs, one per rank (one process per GPU); every
// slab carries HALO extra rows on each side.  The halo rows of an array are refreshed WHEN THE ARRAY
// IS PRODUCED and BY THE KERNEL THAT PRODUCES IT: the fused pass (k_stream<PEER> / k_strip) stores the
// rows a neighbour keeps as its halo -- of the smoothed grid and of the restricted source -- straight
// into the neighbour's slab through peer pointers (CUDA IPC over NVLink).  Large slabs are launched in
// two parts: the thin row segments at both ends (peer stores, flags) first, then the interior with the
// single-GPU kernel, so the halo rows cross NVLink while the interior is swept.  When the (edge) launch
// has drained, its last CTA publishes the pass number
// in a flag word of each neighbour (st.release.sys); the neighbour's stream waits for that value
// (cuStreamWaitValue32, no SM, no host) in front of its next pass.  There is no communication kernel,
// no send/recv group and no extra stream on the data path.
//
// The coarse partition is INDUCED by the fine one through the reference's floor map (a rank owns the
// coarse rows whose lower fine row it owns), so restriction needs no communication beyond those
// halo rows and prolongation only the coarse halo.  Levels below `threshold` rows are agglomerated:
// at the boundary every rank broadcasts its share of the restricted grid to ALL ranks (peer stores
// by a small copy kernel, flag per source rank), and every rank then runs the whole coarse sub-cycle
// REDUNDANTLY with the single-GPU interpreter (fused nodes + coarse tail kernel).  All ranks end up
// with the same coarse correction bit for bit, so the way up needs no scatter and nobody idles.
//
// Slabs, gather buffers and flags live in one ARENA per rank, cut by a deterministic bump allocator
// that mirrors the level stack: every rank can compute every other rank's offsets, so one IPC
// handle per rank (exchanged when the arena is created or grown, outside the timed region after the
// first cycle) gives all peer pointers.
//
// The smoothing error is the sum of the slabs' red-parity sums.  The sums stay on the device per batch
// of nodes; one exchange over peer memory (every rank stores its partials into every rank's scalar
// table and waits for the others' flag, the host adds in rank order: identical on all ranks) replaces
// an all-reduce; error-trigger loops exchange two scalars per two-sweep pass.  NCCL is used for the
// rendezvous, for exchanging the IPC handles and for host-side agreement on allocation success only; it
// is bound at run time (dlopen of the libnccl.so.2 the host program already loaded), so the library
// has no link-time NCCL dependency.  A rank that fails locally publishes an abort word: its peers come
// back with an error instead of waiting (gather waits also time out).
//
//   EmuComm   all ranks live in this process on the current GPU: the same arenas, peer stores, flags
//             and stream waits, with every rank's passes queued in rank order on one stream.  It
//             exists so the whole slab logic is testable bit-for-bit on one GPU.
//   NcclComm  one rank per process.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include <sys/stat.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mg_abi.h"
#include "mg_fused.h"
#include "mg_kernels.h"

namespace mg {
namespace {

constexpr int HALO = 8;           // >= S+2 for S <= 3 sweeps per pass, on the fine and (ratio >= 1.2) the coarse side
constexpr double TRIGGER = 0.01;  // MG_solver_CPU.cpp:99
constexpr int MAX_WORLD = 16;

// ---- driver entry points for stream memory operations
struct DriverApi {
    CUresult (*WaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*WriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    bool load()
    {
        if (WaitValue32 && WriteValue32) return true;
        cudaDriverEntryPointQueryResult q;
        void *f = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
        WaitValue32 = (decltype(WaitValue32))f;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
        WriteValue32 = (decltype(WriteValue32))f;
        return true;
    }
};
DriverApi g_drv;

// ------------------------------------------------------------------ communicators
// What is left to a communicator: the scalar all-reduce and the exchange of the arenas' IPC handles.
class Comm {
public:
    virtual ~Comm() {}
    int world = 1;
    std::vector<int> local;       // ranks living in this process
    bool is_local(int r) const { return std::find(local.begin(), local.end(), r) != local.end(); }
    // vals[i] = device pointer of local rank i; on return every one holds the sum over all ranks (queued on `stream`)
    virtual void allreduce_sum(const std::vector<double *> &vals, int n, cudaStream_t stream) = 0;
    // every rank contributes `bytes` bytes; out = world x bytes in rank order (collective, synchronous); false on failure
    virtual bool allgather_host(const void *mine, void *out, size_t bytes) = 0;
    // min over all ranks of a host integer (collective, synchronous): agreement on success / failure
    virtual int allreduce_min_host(int v) = 0;
};

class EmuComm : public Comm {
public:
    explicit EmuComm(int g)
    {
        world = g;
        for (int r = 0; r < g; ++r) local.push_back(r);
    }
    ~EmuComm() override { cudaDeviceSynchronize(); }
    void allreduce_sum(const std::vector<double *> &vals, int n, cudaStream_t stream) override
    {
        std::vector<double> acc((size_t)n, 0.0), tmp((size_t)n);
        check(cudaStreamSynchronize(stream), "sync");
        for (double *v : vals) {   // rank order: the same association every run
            check(cudaMemcpy(tmp.data(), v, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost), "emu allreduce D2H");
            for (int i = 0; i < n; ++i) acc[i] += tmp[i];
        }
        for (double *v : vals) check(cudaMemcpy(v, acc.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice), "emu allreduce H2D");
    }
    bool allgather_host(const void *, void *, size_t) override { return true; }   // (all ranks are local: nothing to exchange)
    int allreduce_min_host(int v) override { return v; }
};

// ---- NCCL bound at run time
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the host program already uses, if any
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW);
        if (!handle) { fail(-30, std::string("cannot load libnccl.so.2: ") + dlerror()); return false; }
#define MG_SYM(field, name) field = (decltype(field))dlsym(handle, name); if (!field) { fail(-31, "libnccl.so.2 lacks " name); return false; }
        MG_SYM(GetUniqueId, "ncclGetUniqueId") MG_SYM(CommInitRank, "ncclCommInitRank") MG_SYM(CommDestroy, "ncclCommDestroy")
        MG_SYM(AllReduce, "ncclAllReduce") MG_SYM(AllGather, "ncclAllGather") MG_SYM(GetErrorString, "ncclGetErrorString")
#undef MG_SYM
        return true;
    }
};
NcclApi g_nccl;

class NcclComm : public Comm {
public:
    ncclComm_t comm = nullptr;
    int rank = 0;
    unsigned char *stage = nullptr;   // device staging of the host collectives
    size_t stage_bytes = 0;
    bool ok(ncclResult_t r, const char *what)
    {
        if (r == ncclSuccess) return true;
        fail(-32, std::string(what) + ": " + g_nccl.GetErrorString(r));
        return false;
    }
    ~NcclComm() override { cudaFree(stage); }
    bool reserve(size_t bytes)               // staging of allgather_host; called once at init so that no collective can fail on memory later
    {
        if (bytes <= stage_bytes) return true;
        cudaFree(stage);
        stage = nullptr;
        stage_bytes = 0;
        if (cudaMalloc(&stage, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
        stage_bytes = bytes;
        return true;
    }
    void allreduce_sum(const std::vector<double *> &vals, int n, cudaStream_t stream) override
    {
        ok(g_nccl.AllReduce(vals[0], vals[0], (size_t)n, ncclDouble, ncclSum, comm, stream), "ncclAllReduce");
    }
    // NOTE: a host collective is always entered, whatever failed locally before (a rank that returned early would leave
    // the others blocked inside NCCL): local failure travels as data through allreduce_min_host.
    bool allgather_host(const void *mine, void *out, size_t bytes) override
    {
        cudaStream_t st = ctx().stream;
        if ((size_t)(world + 1) * bytes > stage_bytes) { fail(-39, "slab driver: host collective larger than its staging buffer"); return false; }
        cudaMemcpyAsync(stage, mine, bytes, cudaMemcpyHostToDevice, st);
        const bool sent = ok(g_nccl.AllGather(stage, stage + bytes, bytes, ncclChar, comm, st), "ncclAllGather");
        cudaStreamSynchronize(st);
        if (sent) cudaMemcpy(out, stage + bytes, (size_t)world * bytes, cudaMemcpyDeviceToHost);
        return sent;
    }
    int allreduce_min_host(int v) override
    {
        cudaStream_t st = ctx().stream;
        int *d = (int *)(ctx().dev_scalar + 8);
        cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, st);
        ok(g_nccl.AllReduce(d, d, 1, ncclInt, ncclMin, comm, st), "ncclAllReduce (min)");
        cudaStreamSynchronize(st);
        int out = 0;
        cudaMemcpy(&out, d, sizeof(int), cudaMemcpyDeviceToHost);
        return out;
    }
};
std::unique_ptr<NcclComm> g_nccl_comm;

// ------------------------------------------------------------------ geometry (host only, no device state)
// fine row whose pair (f, f+1) produces coarse row c: the floor map of doRestriction
// (MG_solver_CPU.cpp:661-662), with the two coarse boundary rows pinned to 0 and N-2 as in
// mg_fused.cu.  Empty vector = the map is not injective / leaves the grid (pair not fusable).
std::vector<int> fine_of_coarse_uncached(int N, int M)
{
    std::vector<int> out;
    if (!(M >= 3 && M < N && N >= 4)) return out;
    const double h_f = 1.0 / (double)(N - 1), h_c = 1.0 / (double)(M - 1);
    out.assign((size_t)M, 0);
    int prev = 0;
    for (int c = 1; c <= M - 2; ++c) {
        const int f = (int)floor((double)c * h_c / h_f);
        if (f < 1 || f > N - 3 || f <= prev) return std::vector<int>();
        out[c] = prev = f;
    }
    if (prev >= N - 2) return std::vector<int>();
    out[M - 1] = N - 2;
    return out;
}

// every node of every cycle asks for the same few pairs: keep them (the map costs ~10 ns per coarse row)
const std::vector<int> &fine_of_coarse_host(int N, int M)
{
    static std::map<std::pair<int, int>, std::vector<int>> cache;
    auto it = cache.find({N, M});
    if (it == cache.end()) it = cache.emplace(std::make_pair(N, M), fine_of_coarse_uncached(N, M)).first;
    return it->second;
}

bool pair_fusable_host(int N, int M)
{
    return N >= 4 && N % 2 == 0 && M >= 3 && (double)(N - 1) >= 1.2 * (double)(M - 1) && !fine_of_coarse_host(N, M).empty();
}

struct LevelGeom {
    int N = 0;
    bool dist = false;
    std::vector<int> bound;   // dist: owned rows of rank k = [bound[k], bound[k+1])
};

// rank k owns the coarse rows whose lower fine row it owns
std::vector<int> coarse_bounds(const LevelGeom &fine, int M, int world)
{
    const std::vector<int> &foc = fine_of_coarse_host(fine.N, M);
    std::vector<int> b((size_t)world + 1, M);
    int c = 0;
    for (int k = 0; k < world; ++k) {
        while (c < M && foc[c] < fine.bound[k]) ++c;
        b[k] = c;
    }
    return b;
}

LevelGeom top_geometry(int N, int world, int threshold)
{
    LevelGeom g;
    g.N = N;
    g.dist = world > 1 && N >= threshold && N % 2 == 0 && N / world >= 2 * HALO;
    if (g.dist)
        for (int k = 0; k <= world; ++k) g.bound.push_back((int)((long long)N * k / world));
    return g;
}

// coarse geometry induced by the fine one; false (+why) if a distributed fine level cannot be served
bool induce_geometry(const LevelGeom &fine, int M, int world, int threshold, LevelGeom &coarse, std::string &why)
{
    coarse.N = M;
    coarse.dist = false;
    coarse.bound.clear();
    if (!fine.dist) return true;
    if (!pair_fusable_host(fine.N, M)) { why = "a distributed level needs an even size and a fusable transfer pair"; return false; }
    if (!(world > 1 && M >= threshold && M % 2 == 0)) return true;
    const std::vector<int> b = coarse_bounds(fine, M, world);
    for (int k = 0; k < world; ++k)
        if (b[k + 1] - b[k] < 2 * HALO) return true;   // slabs too thin for single-neighbour halos: agglomerate
    coarse.dist = true;
    coarse.bound = b;
    return true;
}

Slab slab_of(const LevelGeom &g, int rank)
{
    Slab s;
    if (!g.dist) { s.row0 = 0; s.rows = g.N; s.own_lo = 0; s.own_hi = g.N; return s; }
    s.own_lo = g.bound[rank];
    s.own_hi = g.bound[rank + 1];
    s.row0 = std::max(0, s.own_lo - HALO);
    s.rows = std::min(g.N, s.own_hi + HALO) - s.row0;
    return s;
}

size_t slab_bytes(const LevelGeom &g, int rank)
{
    const Slab s = slab_of(g, rank);
    return ((size_t)s.rows * g.N * sizeof(double) + 255) / 256 * 256;
}

// ------------------------------------------------------------------ the fabric: arenas + flag words of all ranks
// Flag words of a rank (unsigned int): [0] pass number published by the rank below, [1] by the rank above,
// [2 + s] gather number published by source rank s, [2 + MAX_WORLD + s] scalar-exchange number of source rank s,
// [2 + 2 MAX_WORLD] set when a peer gave up (error).
constexpr int W_FROM_LO = 0, W_FROM_HI = 1, W_GATHER = 2, W_SCAL = 2 + MAX_WORLD, W_ABORT = 2 + 2 * MAX_WORLD, N_WORDS = 64;
constexpr int SCAL_N = 64;                                            // doubles per rank in the scalar exchange
constexpr size_t SCAL_TABLE_BYTES = 2 * MAX_WORLD * SCAL_N * sizeof(double);   // [parity][source rank][SCAL_N], at the bottom of every arena

__global__ void k_peer_bcast(const double *src, size_t count, int n_dst, const double *const *dst_table, unsigned int *const *flag_table,
                             unsigned int flag_val, unsigned int *ticket)
{
    // every destination gets the same `count` doubles at its own address; 16-byte vectors when everything is aligned
    __shared__ bool last;
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int d = 0; d < n_dst; ++d) {
        double *dst = const_cast<double *>(dst_table[d]);
        if ((((size_t)src | (size_t)dst) & 15) == 0) {
            const double2 *s2 = reinterpret_cast<const double2 *>(src);
            double2 *d2 = reinterpret_cast<double2 *>(dst);
            for (size_t k = t0; k < count / 2; k += stride) d2[k] = s2[k];
            if ((count & 1) && t0 == 0) dst[count - 1] = src[count - 1];
        } else {
            for (size_t k = t0; k < count; k += stride) dst[k] = src[k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        for (int d = 0; d < n_dst; ++d) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_table[d]), "r"(flag_val) : "memory");
        *ticket = 0u;
    }
}

// one warp: lane s waits until source rank s has published gather number >= want (or a peer aborted)
__global__ void k_wait_gather(const unsigned int *words, int first_word, int world, int self, unsigned int want)
{
    const int s = threadIdx.x;
    if (s >= world || s == self) return;
    const unsigned int *w = words + first_word + s;
    unsigned int *abort_w = const_cast<unsigned int *>(words) + W_ABORT;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned int v, a;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
        if ((int)(v - want) >= 0) break;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(a) : "l"(abort_w) : "memory");
        if (a) break;
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) { *abort_w = 2u; break; }   // 20 s without the peer's rows: give up instead of holding the GPU (reported as an error)
    }
}

struct PeerView {             // one rank's arena and flag words as seen from this process
    unsigned char *arena = nullptr;
    unsigned int *words = nullptr;
    bool arena_ipc = false, words_ipc = false;
};

class Fabric {
public:
    Comm &comm;
    int world;
    std::vector<PeerView> peer;            // [world]
    size_t cap = 0;                        // bytes of every rank's arena
    size_t gather_bytes = 0;               // size of ONE gather buffer (two of them sit at the bottom of every arena)
    std::vector<size_t> top;               // [world] bump pointers (identical bookkeeping in every process)
    unsigned int seq = 0;                  // distributed passes launched so far (the same on every rank)
    unsigned int gathers = 0;              // agglomeration gathers so far
    bool ok = false;
    // device tables of the broadcast kernel, one set per local rank
    struct Tables {                        // per gather parity: device arrays + the host copy they were filled from
        const double **dst[2] = {nullptr, nullptr};
        unsigned int **flag[2] = {nullptr, nullptr};
        const double *dst_h[2][MAX_WORLD] = {};
        unsigned int *flag_h[2][MAX_WORLD] = {};
    };
    std::vector<Tables> tables;            // [local index]: the gather broadcast
    std::vector<Tables> stables;           // [local index]: the scalar exchange
    unsigned int exchanges = 0;            // scalar exchanges so far

    explicit Fabric(Comm &c) : comm(c), world(c.world), peer((size_t)c.world), top((size_t)c.world, 0)
    {
        ok = g_drv.load() && world <= MAX_WORLD;
        if (!ok) { fail(-36, "slab driver: cuStreamWaitValue32 is unavailable (or more than 16 ranks)"); return; }
        // flag words: allocated once, zeroed, exchanged once
        for (int r : comm.local) {
            unsigned int *w = nullptr;
            if (cudaMalloc(&w, N_WORDS * sizeof(unsigned int)) != cudaSuccess || cudaMemset(w, 0, N_WORDS * sizeof(unsigned int)) != cudaSuccess) ok = false;
            peer[r].words = w;
        }
        tables.resize(comm.local.size());
        stables.resize(comm.local.size());
        for (auto *set : {&tables, &stables})
            for (auto &t : *set)
                for (int k = 0; k < 2; ++k)
                    if (cudaMalloc(&t.dst[k], MAX_WORLD * sizeof(double *)) != cudaSuccess || cudaMalloc(&t.flag[k], MAX_WORLD * sizeof(unsigned int *)) != cudaSuccess) ok = false;
        cudaDeviceSynchronize();
        ok = exchange(false) && ok;
    }
    ~Fabric() { release(true); }
    Fabric(const Fabric &) = delete;
    Fabric &operator=(const Fabric &) = delete;

    void release(bool words_too)
    {
        cudaDeviceSynchronize();
        for (int r = 0; r < world; ++r) {
            PeerView &p = peer[r];
            if (p.arena) { if (p.arena_ipc) cudaIpcCloseMemHandle(p.arena); else if (comm.is_local(r)) cudaFree(p.arena); }
            p.arena = nullptr;
            p.arena_ipc = false;
            if (words_too && p.words) {
                if (p.words_ipc) cudaIpcCloseMemHandle(p.words); else if (comm.is_local(r)) cudaFree(p.words);
                p.words = nullptr;
                p.words_ipc = false;
            }
        }
        if (words_too)
            for (auto *set : {&tables, &stables})
                for (auto &t : *set)
                    for (int k = 0; k < 2; ++k) { cudaFree(t.dst[k]); cudaFree(t.flag[k]); t.dst[k] = nullptr; t.flag[k] = nullptr; }
        for (auto *set : {&tables, &stables})        // the arenas move: cached destination tables are stale
            for (auto &t : *set) { memset(t.dst_h, 0, sizeof t.dst_h); memset(t.flag_h, 0, sizeof t.flag_h); }
        cap = 0;
    }

    // Collective: publish the local arena (arenas == true) or flag words to the other processes and map theirs.
    bool exchange(bool arenas)
    {
        if (comm.local.size() == (size_t)world) return true;            // all ranks in this process: the pointers are shared already
        const int me = comm.local[0];
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof mine);
        void *ptr = arenas ? (void *)peer[me].arena : (void *)peer[me].words;
        int good = ptr && cudaIpcGetMemHandle(&mine, ptr) == cudaSuccess ? 1 : 0;
        std::vector<cudaIpcMemHandle_t> all((size_t)world);
        good = comm.allgather_host(&mine, all.data(), sizeof mine) && good;
        good = comm.allreduce_min_host(good);                            // every rank has a handle to offer
        if (good) {
            for (int q = 0; q < world && good; ++q) {
                if (q == me) continue;
                void *m = nullptr;
                if (cudaIpcOpenMemHandle(&m, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = 0; cudaGetLastError(); break; }
                if (arenas) { peer[q].arena = (unsigned char *)m; peer[q].arena_ipc = true; }
                else { peer[q].words = (unsigned int *)m; peer[q].words_ipc = true; }
            }
        }
        good = comm.allreduce_min_host(good);                            // ... and every rank mapped all of them
        if (!good) fail(-37, "slab driver: CUDA IPC mapping of the peer arenas failed (peer access over NVLink is required)");
        return good != 0;
    }

    // Collective: arenas of at least `need` bytes (the same number on every rank).  Keeps the current ones when large enough.
    bool reserve(size_t need, size_t gather_need)
    {
        if (!ok) return false;
        if (need <= cap && gather_need <= gather_bytes) return true;
        release(false);
        gather_bytes = std::max(gather_bytes, gather_need);
        const size_t want = std::max(need, (size_t)1 << 20);
        int good = 1;
        for (int r : comm.local) {
            void *a = nullptr;
            if (cudaMalloc(&a, want) != cudaSuccess) { good = 0; cudaGetLastError(); }
            peer[r].arena = (unsigned char *)a;
        }
        good = comm.allreduce_min_host(good);
        if (!good) { fail(-38, "slab driver: cannot allocate the slab arena"); ok = false; return false; }
        cap = want;
        ok = exchange(true);
        return ok;
    }

    // ---- bump allocation, the same in every process: `bytes[r]` from rank r's arena; returns the offsets
    std::vector<size_t> alloc(const std::vector<size_t> &bytes)
    {
        std::vector<size_t> off((size_t)world);
        for (int r = 0; r < world; ++r) { off[r] = top[r]; top[r] += bytes[r]; }
        return off;
    }
    template <class T>
    T *at(int rank, size_t off) const { return reinterpret_cast<T *>(peer[rank].arena + off); }
    // bottom of every arena: the scalar-exchange table, then two gather buffers (double-buffered by gather parity), then the level stack
    size_t scal_off(unsigned int x, int src) const { return ((size_t)(x & 1u) * MAX_WORLD + (size_t)src) * SCAL_N * sizeof(double); }
    size_t gather_off(unsigned int g) const { return SCAL_TABLE_BYTES + (size_t)(g & 1u) * gather_bytes; }
    size_t stack_base() const { return SCAL_TABLE_BYTES + 2 * gather_bytes; }

    void wait32(cudaStream_t st, unsigned int *addr, unsigned int value)
    {
        if (g_drv.WaitValue32((CUstream)st, (CUdeviceptr)addr, value, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) fail(-36, "cuStreamWaitValue32");
    }
    // Tell every peer that this rank gives up (their flag waits pass, their gather waits end): called on a local error so
    // that the other ranks come back with an error instead of hanging.
    void abort_peers(cudaStream_t st)
    {
        for (int me : comm.local)
            for (int q = 0; q < world; ++q) {
                if (q == me || !peer[q].words) continue;
                g_drv.WriteValue32((CUstream)st, (CUdeviceptr)(peer[q].words + W_ABORT), 1u, 0);
                if (q == me - 1) g_drv.WriteValue32((CUstream)st, (CUdeviceptr)(peer[q].words + W_FROM_HI), 0x7fffffffu, 0);
                if (q == me + 1) g_drv.WriteValue32((CUstream)st, (CUdeviceptr)(peer[q].words + W_FROM_LO), 0x7fffffffu, 0);
            }
    }
};

struct RankLevel {            // one local rank's share of one level
    Slab slab;
    double *U = nullptr, *W = nullptr, *F = nullptr;
    bool owns_F = true;       // agglomerated levels: F from the pool (false: the gather buffer / the caller's source)
    bool pooled = false;      // agglomerated level: U, W (and F if owns_F) come from the pool
};

struct LevelAlloc {           // arena offsets of a distributed level, for EVERY rank
    std::vector<size_t> U, W, F;
    std::vector<size_t> saved_top;
};

struct RankState {
    int rank = 0;
    std::vector<RankLevel> lv;
    double *scal = nullptr;   // device scalars: error partials of the batch, [60] final abs-diff partial
};

class DistCycle {
public:
    DistCycle(Comm &c, Fabric &f, int threshold) : comm(c), fab(f), threshold_(threshold)
    {
        for (int r : comm.local) {
            RankState st;
            st.rank = r;
            st.scal = (double *)pool_alloc(64 * sizeof(double));
            ranks.push_back(st);
        }
        std::fill(fab.top.begin(), fab.top.end(), fab.stack_base());        // the exchange table and the gather buffers sit at the bottom
    }
    ~DistCycle()
    {
        cudaStreamSynchronize(ctx().stream);     // nothing queued may outlive the buffers
        while (!geom.empty()) pop();
        for (auto &st : ranks) pool_free(st.scal);
    }

    Comm &comm;
    Fabric &fab;
    int threshold_;
    std::vector<LevelGeom> geom;
    std::vector<LevelAlloc> alloc;
    std::vector<RankState> ranks;
    int init_ = 1;
    int scal_used_ = 0;         // slots of the ranks' scal arrays holding error sums of this batch

    // borrowed_F: the top level's source slab of local rank 0 (single local rank only); gathered: F is the gather buffer
    void push(const LevelGeom &g, double *borrowed_F = nullptr, bool gathered = false, unsigned int gather_no = 0)
    {
        geom.push_back(g);
        LevelAlloc a;
        a.saved_top = fab.top;
        if (g.dist) {
            std::vector<size_t> bytes((size_t)comm.world);
            for (int r = 0; r < comm.world; ++r) bytes[r] = slab_bytes(g, r);
            a.U = fab.alloc(bytes);
            a.W = fab.alloc(bytes);
            if (!borrowed_F) a.F = fab.alloc(bytes);
        }
        alloc.push_back(a);
        for (auto &st : ranks) {
            RankLevel l;
            l.slab = slab_of(g, st.rank);
            if (g.dist) {
                l.U = fab.at<double>(st.rank, a.U[st.rank]);
                l.W = fab.at<double>(st.rank, a.W[st.rank]);
                l.F = borrowed_F ? borrowed_F : fab.at<double>(st.rank, a.F[st.rank]);
                l.owns_F = false;
            } else {
                const size_t bytes = (size_t)g.N * g.N * sizeof(double);
                l.pooled = true;
                l.U = (double *)pool_alloc(bytes);
                l.W = (double *)pool_alloc(bytes);
                if (gathered) { l.F = fab.at<double>(st.rank, fab.gather_off(gather_no)); l.owns_F = false; }
                else if (borrowed_F) { l.F = borrowed_F; l.owns_F = false; }
                else l.F = (double *)pool_alloc(bytes);
            }
            st.lv.push_back(l);
        }
    }
    void pop()
    {
        for (auto &st : ranks) {
            RankLevel &l = st.lv.back();
            if (l.pooled) {
                pool_free(l.U); pool_free(l.W);
                if (l.owns_F) pool_free(l.F);
            }
            st.lv.pop_back();
        }
        fab.top = alloc.back().saved_top;
        alloc.pop_back();
        geom.pop_back();
        if (geom.size() == 1) init_ = 0;
    }
    bool restart_top() const { return init_ == 0 && geom.size() == 1; }

    bool induce(const LevelGeom &fine, int M, LevelGeom &coarse, std::string &why)
    {
        return induce_geometry(fine, M, comm.world, threshold_, coarse, why);
    }

    // Scalar exchange without a collective library: every rank stores its SCAL_N partials into every rank's table (peer
    // stores by the small broadcast kernel, flag per source rank), waits for everybody else's, and copies the table of this
    // exchange to `host` ([world][SCAL_N], queued on the stream).  The caller adds the ranks' values IN RANK ORDER after a
    // synchronisation: the same association on every rank and in every run, so all ranks take the same decisions.
    void exchange(double *host)
    {
        Context &c = ctx();
        const unsigned int x = ++fab.exchanges;
        const int par = (int)(x & 1u);
        for (size_t i = 0; i < ranks.size(); ++i) {
            const int r = ranks[i].rank;
            const double *dst_h[MAX_WORLD] = {};
            unsigned int *flag_h[MAX_WORLD] = {};
            for (int q = 0; q < comm.world; ++q) {          // (its own table too: a local store)
                dst_h[q] = fab.at<double>(q, fab.scal_off(x, r));
                flag_h[q] = fab.peer[q].words + W_SCAL + r;
            }
            Fabric::Tables &t = fab.stables[i];
            if (memcmp(t.dst_h[par], dst_h, sizeof dst_h) || memcmp(t.flag_h[par], flag_h, sizeof flag_h)) {
                memcpy(t.dst_h[par], dst_h, sizeof dst_h);
                memcpy(t.flag_h[par], flag_h, sizeof flag_h);
                check(cudaMemcpyAsync(t.dst[par], dst_h, sizeof dst_h, cudaMemcpyHostToDevice, c.stream), "H2D exchange table");
                check(cudaMemcpyAsync(t.flag[par], flag_h, sizeof flag_h, cudaMemcpyHostToDevice, c.stream), "H2D exchange table");
            }
            k_peer_bcast<<<1, 64, 0, c.stream>>>(ranks[i].scal, SCAL_N, comm.world, t.dst[par], t.flag[par], x, c.counters + 12);
            c.launches++;
            check(cudaGetLastError(), "k_peer_bcast (scalars)");
        }
        for (size_t i = 0; i < ranks.size(); ++i) {
            const int r = ranks[i].rank;
            if (comm.world > 1) {
                k_wait_gather<<<1, 32, 0, c.stream>>>(fab.peer[r].words, W_SCAL, comm.world, r, x);
                c.launches++;
                check(cudaGetLastError(), "k_wait_gather (scalars)");
            }
        }
        const int r0 = ranks[0].rank;
        check(cudaMemcpyAsync(host, fab.at<double>(r0, fab.scal_off(x, 0)), (size_t)comm.world * SCAL_N * sizeof(double), cudaMemcpyDeviceToHost, c.stream),
              "D2H exchange table");
    }
    static double rank_sum(const double *table, int world, int idx)
    {
        double s = 0.0;
        for (int q = 0; q < world; ++q) s += table[(size_t)q * SCAL_N + idx];
        return s;
    }

    // One fused pass over every local slab of distributed level `li`: S sweeps from U (in_mode 0), from zero (1) or
    // from U + prolongation of the coarse grid `uc` (2); optional red-parity error sum into scal[scal_idx]; optional
    // restriction of the negated residual into `fc` (the coarse level li+1: its slabs when fc_dist, else every rank's
    // full gather buffer).  The rows the neighbours keep as halos -- of the output W and of a distributed F_c -- go
    // straight into their arenas; the pass number is published when the launch has drained.
    void pass(int li, double L, int S, int in_mode, bool want_err, int scal_idx, int M, bool with_fc, bool fc_dist,
              const std::vector<int> &cb, const std::vector<double *> &fc_full, int Nc, const std::vector<const double *> &uc,
              const std::vector<Slab> &uc_slab, bool mid = false)
    {
        const LevelGeom &g = geom[li];
        const bool writes_U = !(S == 0 && in_mode == 0);
        const unsigned int prev = fab.seq, mine = ++fab.seq;
        cudaStream_t st_ = ctx().stream;
        for (size_t i = 0; i < ranks.size(); ++i) {
            const int r = ranks[i].rank;
            RankLevel &l = ranks[i].lv[li];
            PeerLinks pl;
            // the previous pass of both neighbours has drained: my halos are current, and nobody still reads what I overwrite
            if (r > 0) fab.wait32(st_, fab.peer[r].words + W_FROM_LO, prev);
            if (r + 1 < comm.world) fab.wait32(st_, fab.peer[r].words + W_FROM_HI, prev);
            pl.flag_val = mine;
            if (r > 0) pl.flag_lo = fab.peer[r - 1].words + W_FROM_HI;
            if (r + 1 < comm.world) pl.flag_hi = fab.peer[r + 1].words + W_FROM_LO;
            if (writes_U) {
                if (r > 0) pl.U_lo = fab.at<double>(r - 1, alloc[li].W[r - 1]) - (ptrdiff_t)slab_of(g, r - 1).row0 * g.N;
                if (r + 1 < comm.world) pl.U_hi = fab.at<double>(r + 1, alloc[li].W[r + 1]) - (ptrdiff_t)slab_of(g, r + 1).row0 * g.N;
                pl.u_lo_end = l.slab.own_lo + HALO;
                pl.u_hi_begin = l.slab.own_hi - HALO;
            }
            Slab fcs;
            double *fc = nullptr;
            if (with_fc) {
                if (fc_dist) {
                    const LevelGeom &gc = geom[li + 1];
                    fcs = ranks[i].lv[li + 1].slab;
                    fc = ranks[i].lv[li + 1].F;
                    if (r > 0) pl.Fc_lo = fab.at<double>(r - 1, alloc[li + 1].F[r - 1]) - (ptrdiff_t)slab_of(gc, r - 1).row0 * M;
                    if (r + 1 < comm.world) pl.Fc_hi = fab.at<double>(r + 1, alloc[li + 1].F[r + 1]) - (ptrdiff_t)slab_of(gc, r + 1).row0 * M;
                    pl.fc_lo_end = cb[r] + HALO;
                    pl.fc_hi_begin = cb[r + 1] - HALO;
                } else {       // the rank's coarse rows land in its own full-size gather buffer; the broadcast follows the pass
                    fcs.row0 = 0; fcs.rows = M; fcs.own_lo = cb[r]; fcs.own_hi = cb[r + 1];
                    fc = fc_full[i];
                }
            }
            if (mid)     // a pass of an error-trigger loop: S = 2 with both errors (scal[idx] last, scal[idx+1] first sweep) or S = 1
                slab_trigger_pass(g.N, L, S, in_mode, l.U, l.F, l.W, l.slab, ranks[i].scal + scal_idx, M, fc, with_fc ? &fcs : nullptr, Nc,
                                  uc.empty() ? nullptr : uc[i], uc.empty() ? nullptr : &uc_slab[i], pl);
            else
                slab_pass(g.N, L, S, in_mode, l.U, l.F, writes_U ? l.W : nullptr, l.slab, want_err, ranks[i].scal + scal_idx, M, fc,
                          with_fc ? &fcs : nullptr, Nc, uc.empty() ? nullptr : uc[i], uc.empty() ? nullptr : &uc_slab[i], pl);
        }
        if (writes_U) {
            for (auto &st : ranks) std::swap(st.lv[li].U, st.lv[li].W);
            std::swap(alloc[li].U, alloc[li].W);
        }
    }

    // the last pass went one sweep too far (trigger loop): its input becomes the current grid again
    void unswap(int li)
    {
        for (auto &st : ranks) std::swap(st.lv[li].U, st.lv[li].W);
        std::swap(alloc[li].U, alloc[li].W);
    }

    // Agglomeration boundary: every rank sends its rows [cb[r], cb[r+1]) of the M x M restricted grid (already in its own
    // gather buffer) to the same place in every other rank's gather buffer, then waits for everybody else's rows.
    void gather_all(int M, const std::vector<int> &cb, unsigned int g)
    {
        Context &c = ctx();
        for (size_t i = 0; i < ranks.size(); ++i) {        // all broadcasts are queued before any wait (one stream when emulated)
            const int r = ranks[i].rank;
            const size_t off = fab.gather_off(g) + (size_t)cb[r] * M * sizeof(double), count = (size_t)(cb[r + 1] - cb[r]) * M;
            const double *dst_h[MAX_WORLD] = {};
            unsigned int *flag_h[MAX_WORLD] = {};
            int n = 0;
            for (int q = 0; q < comm.world; ++q) {
                if (q == r) continue;
                dst_h[n] = fab.at<double>(q, off);
                flag_h[n] = fab.peer[q].words + W_GATHER + r;
                ++n;
            }
            Fabric::Tables &t = fab.tables[i];
            const int par = (int)(g & 1u);
            if (memcmp(t.dst_h[par], dst_h, sizeof dst_h) || memcmp(t.flag_h[par], flag_h, sizeof flag_h)) {   // same every cycle: uploaded once
                memcpy(t.dst_h[par], dst_h, sizeof dst_h);
                memcpy(t.flag_h[par], flag_h, sizeof flag_h);
                check(cudaMemcpyAsync(t.dst[par], dst_h, sizeof dst_h, cudaMemcpyHostToDevice, c.stream), "H2D gather table");
                check(cudaMemcpyAsync(t.flag[par], flag_h, sizeof flag_h, cudaMemcpyHostToDevice, c.stream), "H2D gather table");
            }
            const int blocks = (int)std::max<size_t>(1, std::min<size_t>(2 * c.sm_count, (count / 2 + 255) / 256));
            k_peer_bcast<<<blocks, 256, 0, c.stream>>>(fab.at<double>(r, off), count, n, t.dst[par], t.flag[par], g, c.counters + 12);
            c.launches++;
            check(cudaGetLastError(), "k_peer_bcast");
        }
        for (size_t i = 0; i < ranks.size(); ++i) {
            const int r = ranks[i].rank;
            k_wait_gather<<<1, 32, 0, c.stream>>>(fab.peer[r].words, W_GATHER, comm.world, r, g);
            c.launches++;
            check(cudaGetLastError(), "k_wait_gather");
        }
    }
};

// Top-level source slab kept between calls (MG_RUN_SKIP_SOURCE): benchmarks and host-buffer callers
// generate / upload F once instead of once per cycle.
struct SourceCache {
    int N = 0, world = 0, rank = -1, row0 = 0, rows = 0;
    double L = 0, min_x = 0, min_y = 0;
    double *F = nullptr;
    bool from_host = false;
};
SourceCache g_src;
std::unique_ptr<Fabric> g_fabric;   // of the NCCL communicator (persistent: the arenas and their IPC mappings survive between calls)

void drop_source_cache()
{
    if (g_src.F && ctx().ready) pool_free(g_src.F);
    g_src = SourceCache();
}

const char *kRestrictArt = "             *\n             |\n Restriction |\n             |\n             *\n";
const char *kProlongArt = "             *\n             |\nProlongation |\n             |\n             *\n";

struct Header {
    double L = 0, min_x = 0, min_y = 0;
    int con_step = 0, con_N = 0, N_max = 0, N_min = 0;
    std::vector<int> ladder;
    size_t first_node = 7;
};

bool parse_header(const std::vector<double> &tok, Header &h)
{
    if (tok.size() < 7) return false;
    h.L = tok[0]; h.min_x = tok[1]; h.min_y = tok[2];
    h.con_step = (int)tok[3]; h.con_N = (int)tok[4]; h.N_max = (int)tok[5]; h.N_min = (int)tok[6];
    if (h.con_N == 1) for (int n = h.N_max; n >= h.N_min; n /= 2) h.ladder.push_back(n);
    if (h.con_N == 2) for (int n = h.N_max; n >= h.N_min; --n) h.ladder.push_back(n);
    return true;
}

// Walks the node stream with the geometry only: the arena every rank needs (max over ranks, in bytes, above the two
// gather buffers) and the largest gathered grid.  Mirrors the stack moves of run_dist exactly; agglomerated sub-streams
// are skipped with the parse-only mode of the single-GPU interpreter.  Returns 0, or the error code run_dist would hit.
int plan_arena(const std::vector<double> &tok, const Header &h, int world, int threshold, size_t &arena_need, size_t &gather_need)
{
    std::vector<LevelGeom> geom;
    std::vector<std::vector<size_t>> tops;       // per level: tops before its allocation
    std::vector<size_t> top((size_t)world, 0), peak((size_t)world, 0);
    gather_need = 0;
    auto push = [&](const LevelGeom &g) {          // U, W and F slabs (the top level's F may be the caller's: planned anyway)
        tops.push_back(top);
        geom.push_back(g);
        if (g.dist)
            for (int r = 0; r < world; ++r) {
                top[r] += 3 * slab_bytes(g, r);
                peak[r] = std::max(peak[r], top[r]);
            }
    };
    auto pop = [&]() { top = tops.back(); tops.pop_back(); geom.pop_back(); };
    push(top_geometry(h.N_max, world, threshold));
    size_t cur = h.first_node, pos = 0;
    int init = 1, n_recs = 0;
    std::string why;
    while (cur < tok.size()) {
        if (!geom.back().dist && ((int)tok[cur] == -1 || (int)tok[cur] == 0)) {
            int c2 = (int)cur, p2 = (int)pos;
            double *u = nullptr, *w = nullptr;
            const int rc = mgRunSubcycle(tok.data(), (int)tok.size(), &c2, &p2, h.ladder.data(), (int)h.ladder.size(), h.con_step, h.con_N, h.L,
                                         geom.back().N, &u, &w, nullptr, (int)geom.size() - 1, &init, MG_RUN_FUSED | MG_RUN_QUIET, nullptr, 0,
                                         &n_recs, 0);
            if (rc) return rc;
            if ((size_t)c2 == cur) return 7;     // no progress: malformed stream
            cur = (size_t)c2; pos = (size_t)p2;
            continue;
        }
        const int node = (int)tok[cur++];
        if (node == 2) break;
        if (node == -1) {
            int step, next_N;
            if (h.con_step == 0) { if (cur >= tok.size()) return 3; step = (int)tok[cur++]; } else step = h.con_step;
            if (h.con_N == 0) { if (cur >= tok.size()) return 3; next_N = (int)tok[cur++]; }
            else { if (pos + 1 >= h.ladder.size()) return 4; next_N = h.ladder[++pos]; }
            if (step == 0) continue;
            LevelGeom coarse;
            if (!induce_geometry(geom.back(), next_N, world, threshold, coarse, why)) return 20;
            if (!coarse.dist) gather_need = std::max(gather_need, ((size_t)next_N * next_N * sizeof(double) + 255) / 256 * 256);
            push(coarse);
        } else if (node == 0) {
            return 21;                           // an exact solve on a distributed level
        } else if (node == 1) {
            if (h.con_step == 0) { if (cur >= tok.size()) return 3; ++cur; }
            if (h.con_N != 0 && pos > 0) --pos;
            if (geom.size() < 2) return 5;
            pop();
            if (geom.size() == 1) init = 0;
        } else return 6;
    }
    arena_need = 0;
    for (int r = 0; r < world; ++r) arena_need = std::max(arena_need, peak[r]);
    return 0;
}

struct RunIo {                 // where the top-level source comes from and where the solution goes
    const double *F_host_full = nullptr;   // emulation: the whole N_max^2 source on the host (each rank uploads its slab)
    const double *F_slab_dev = nullptr;    // one rank per process: its source slab, already on the device (rows [row0, row0+rows))
    double *U_host_full = nullptr;         // emulation: assembled solution
    double *U_host_own = nullptr;          // the rank's owned rows -> host (synchronous)
    double *U_dev_own = nullptr;           // the rank's owned rows -> device buffer (stream-ordered copy; batch call)
    int *own_lo = nullptr, *own_hi = nullptr;
};

int run_dist(Comm &comm, Fabric &fab, const char *path, int threshold, int flags, mgTraceRec *recs, int max_recs, mgCycleResult *res,
             const RunIo &io)
{
    // The token stream and the arena plan of a cycle file are kept between calls (throughput loops run the same file again
    // and again; every microsecond of host work here is a microsecond the GPUs of ALL ranks idle, through the neighbour
    // dependencies): keyed by path, size and modification time.
    struct CachedFile {
        std::string path;
        long long size = -1, mtime_ns = 0;
        int world = 0, threshold = 0;
        std::vector<double> tok;
        Header h;
        int plan_rc = 0;
        size_t arena_need = 0, gather_need = 0;
    };
    static CachedFile cached;
    {
        struct stat sb;
        if (stat(path, &sb) != 0) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
        const long long mt = (long long)sb.st_mtim.tv_sec * 1000000000ll + sb.st_mtim.tv_nsec;
        if (cached.path != path || cached.size != (long long)sb.st_size || cached.mtime_ns != mt || cached.world != comm.world ||
            cached.threshold != threshold) {
            std::ifstream f(path);
            if (!f.is_open()) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
            CachedFile nf;
            for (double d; f >> d;) nf.tok.push_back(d);
            if (!parse_header(nf.tok, nf.h)) return 2;
            nf.path = path; nf.size = (long long)sb.st_size; nf.mtime_ns = mt; nf.world = comm.world; nf.threshold = threshold;
            nf.plan_rc = plan_arena(nf.tok, nf.h, comm.world, threshold, nf.arena_need, nf.gather_need);
            cached = std::move(nf);
        }
    }
    const std::vector<double> &tok = cached.tok;
    const Header &h = cached.h;
    const std::vector<int> &ladder = h.ladder;
    const double L = h.L, min_x = h.min_x, min_y = h.min_y;
    const int con_step = h.con_step, con_N = h.con_N, N_max = h.N_max;
    size_t cur = h.first_node, pos = 0;
    auto have = [&](size_t n) { return cur + n <= tok.size(); };
    auto next_int = [&]() { return (int)tok[cur++]; };
    const bool quiet = (flags & MG_RUN_QUIET) != 0 || !comm.is_local(0);
    Context &c = ctx();

    // ---- arenas: sized from a dry walk of the node stream (every rank computes the same numbers), created or grown collectively
    {
        const size_t arena_need = cached.arena_need, gather_need = cached.gather_need;
        const int prc = cached.plan_rc;
        if (prc == 20) fail(-40, "a distributed level needs an even size and a fusable transfer pair: raise the threshold");
        if (prc == 21) fail(-41, "the exact solver runs on an agglomerated level: lower the coarsest size or raise the threshold");
        if (prc) return prc;
        const size_t g_need = std::max(gather_need, fab.gather_bytes);
        if (!fab.reserve(SCAL_TABLE_BYTES + 2 * g_need + arena_need, g_need)) return 30;
    }

    DistCycle cy(comm, fab, threshold);
    {
        const LevelGeom g = top_geometry(N_max, comm.world, threshold);
        // one-rank-per-process runs may keep the top-level source slab between calls
        const bool cacheable = (flags & MG_RUN_SKIP_SOURCE) && cy.ranks.size() == 1 && !io.F_slab_dev && !io.F_host_full;
        double *borrowed = const_cast<double *>(io.F_slab_dev);
        if (cacheable) {
            const Slab s0 = slab_of(g, cy.ranks[0].rank);
            const bool hit = g_src.F && g_src.N == N_max && g_src.world == comm.world && g_src.rank == cy.ranks[0].rank &&
                             g_src.row0 == s0.row0 && g_src.rows == s0.rows &&
                             (g_src.from_host || (g_src.L == L && g_src.min_x == min_x && g_src.min_y == min_y));
            if (!hit) {
                drop_source_cache();
                g_src.F = (double *)pool_alloc((size_t)s0.rows * N_max * sizeof(double));
                g_src.N = N_max; g_src.world = comm.world; g_src.rank = cy.ranks[0].rank; g_src.row0 = s0.row0; g_src.rows = s0.rows;
                g_src.L = L; g_src.min_x = min_x; g_src.min_y = min_y;
                launch_source(N_max, L, g_src.F, min_x, min_y, false, s0.row0, s0.rows);
            }
            borrowed = g_src.F;
        }
        cy.push(g, borrowed);
        if (!borrowed)
            for (auto &st : cy.ranks) {
                RankLevel &l = st.lv[0];
                if (io.F_host_full)      // emulation with a host source: every rank takes its rows, halo included
                    check(cudaMemcpyAsync(l.F, io.F_host_full + (size_t)l.slab.row0 * N_max, (size_t)l.slab.rows * N_max * sizeof(double),
                                          cudaMemcpyHostToDevice, c.stream), "H2D source slab");
                else launch_source(N_max, L, l.F, min_x, min_y, false, l.slab.row0, l.slab.rows);   // halo rows computed locally
            }
    }
    check(cudaStreamSynchronize(c.stream), "sync");

    // Error sums of fixed-step nodes stay on the device as per-rank partials in consecutive scal slots;
    // ONE all-reduce per batch (at the end, or when the slots run out) turns them into the trace records.
    int n_recs = 0;
    auto record = [&](int node, int N, int steps, double err) {
        const int r = n_recs++;
        if (recs && r < max_recs) { recs[r].node = node; recs[r].N = N; recs[r].steps = steps; recs[r].err = err; }
        return r;
    };
    constexpr int SCAL_BATCH = 48;
    static double *xtable = nullptr;               // pinned: [world][SCAL_N] table of the last scalar exchange
    if (!xtable && !check(cudaHostAlloc(&xtable, MAX_WORLD * SCAL_N * sizeof(double), cudaHostAllocDefault), "cudaHostAlloc exchange table")) return 30;
    std::vector<std::pair<int, int>> deferred;     // (trace record, scal slot) holding a raw red-parity sum
    auto harvest = [&]() {
        if (cy.scal_used_ > 0) {
            cy.exchange(xtable);
            check(cudaStreamSynchronize(c.stream), "sync");
            for (const auto &d : deferred) {
                if (!recs || d.first >= max_recs) continue;
                const double v = DistCycle::rank_sum(xtable, comm.world, d.second), N = recs[d.first].N;
                recs[d.first].err = (v + v) / N / N;                      // (sum1+sum2)/N/N, :621-622
            }
            deferred.clear();
            cy.scal_used_ = 0;
        }
        check(cudaStreamSynchronize(c.stream), "sync");
        mgSubcycleHarvest();                     // scalars of the agglomerated sub-cycles
    };
    auto next_scal = [&]() {
        if (cy.scal_used_ == SCAL_BATCH) harvest();
        return cy.scal_used_++;
    };
    // all-reduce scal[idx] and wait for the value on the host (trigger loops)
    auto reduced_now = [&](int idx) {
        cy.exchange(xtable);
        check(cudaStreamSynchronize(c.stream), "sync");
        return DistCycle::rank_sum(xtable, comm.world, idx);
    };
    const std::vector<const double *> no_uc;
    const std::vector<Slab> no_slab;
    const std::vector<double *> no_fc;
    const std::vector<int> no_cb;
    // The error-trigger loop on a distributed level (:216-230 / :388-402), two sweeps per pass: the pass reports the error after
    // each of its sweeps, both are all-reduced, every rank takes the same decision.  If the loop ends after an odd number of
    // sweeps the last pass is repeated with one sweep from the same input.  With a coarse level (M > 0) every pass restricts
    // its result speculatively, so a -1 node whose trigger fires at the minimum of two sweeps costs one pass.
    auto trigger_loop = [&](int li, int first_mode, int N, int M, bool with_fc, bool fc_dist, const std::vector<int> &cb,
                            const std::vector<double *> &fc_full, int Nc, const std::vector<const double *> &uc, const std::vector<Slab> &uc_slab,
                            int &done, double &err_host) {
        double prev = 0.0;
        int mode = first_mode;
        done = 0;
        for (;;) {
            if (cy.scal_used_ + 2 > SCAL_BATCH) harvest();
            const int idx = cy.scal_used_;
            cy.scal_used_ += 2;
            cy.pass(li, L, 2, mode, true, idx, M, with_fc, fc_dist, cb, fc_full, mode == 2 ? Nc : 0, mode == 2 ? uc : no_uc, mode == 2 ? uc_slab : no_slab, true);
            cy.exchange(xtable);
            check(cudaStreamSynchronize(c.stream), "sync");
            if (c.err_code) break;
            const double s[2] = {DistCycle::rank_sum(xtable, comm.world, idx), DistCycle::rank_sum(xtable, comm.world, idx + 1)};
            const double e2 = (s[0] + s[0]) / N / N, e1 = (s[1] + s[1]) / N / N;
            if (done + 1 > 1 && std::fabs(e1 - prev) <= TRIGGER) {       // the loop ends after the first of these two sweeps
                cy.unswap(li);
                const int j = next_scal();
                cy.pass(li, L, 1, mode, true, j, M, with_fc, fc_dist, cb, fc_full, mode == 2 ? Nc : 0, mode == 2 ? uc : no_uc, mode == 2 ? uc_slab : no_slab, true);
                const double r = reduced_now(j);
                done += 1;
                err_host = (r + r) / N / N;
                break;
            }
            done += 2;
            mode = 0;
            err_host = e2;
            if (std::fabs(e2 - e1) <= TRIGGER) break;
            prev = e2;
        }
    };

    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    const long long launches0 = c.launches;
    const auto wall0 = std::chrono::steady_clock::now();
    cudaEventRecord(ev0, c.stream);

    // MG_DIST_TRACE=1: per-node time line (device time on the stream, host enqueue time)
    struct Mark { const char *what; int N; cudaEvent_t ev; std::chrono::steady_clock::time_point host; };
    std::vector<Mark> marks;
    const bool tracing = getenv("MG_DIST_TRACE") && atoi(getenv("MG_DIST_TRACE")) != 0;
    auto mark = [&](const char *what, int N) {
        if (!tracing) return;
        Mark m{what, N, nullptr, std::chrono::steady_clock::now()};
        cudaEventCreate(&m.ev);
        cudaEventRecord(m.ev, c.stream);
        marks.push_back(m);
    };
    mark("start", N_max);

    int rc = 0, node = 0;
    std::string why;
    while (have(1)) {
        // ---- agglomerated sub-cycle: everything that happens at or below a level every rank holds in full is run by the
        // single-GPU interpreter (fused nodes + coarse tail kernel), redundantly on every rank.
        if (!cy.geom.back().dist && ((int)tok[cur] == -1 || (int)tok[cur] == 0)) {
            const int li = (int)cy.geom.size() - 1;
            int icur = (int)cur, ipos = (int)pos, n_rec_io = n_recs, init_io = cy.init_;
            int sub_rc = 0;
            for (size_t i = 0; i < cy.ranks.size(); ++i) {
                RankLevel &l = cy.ranks[i].lv[li];
                int c2 = (int)cur, p2 = (int)pos, r2 = n_recs, i2 = cy.init_;
                const bool lead = i == 0;    // emulation: only the first local rank reports into the caller's records
                sub_rc = mgRunSubcycle(tok.data(), (int)tok.size(), &c2, &p2, ladder.data(), (int)ladder.size(), con_step, con_N, L,
                                       cy.geom[li].N, &l.U, &l.W, l.F, li, &i2, lead ? (flags | MG_RUN_DEFER_HARVEST) : flags,
                                       lead ? recs : nullptr, lead ? max_recs : 0, &r2, 1);
                icur = c2; ipos = p2; n_rec_io = r2; init_io = i2;
                if (sub_rc) break;
            }
            if (sub_rc) { rc = sub_rc; break; }
            cur = (size_t)icur; pos = (size_t)ipos; n_recs = n_rec_io; cy.init_ = init_io;
            mark("sub-cycle", cy.geom[li].N);
            continue;
        }
        node = next_int();
        if (node == 2) break;
        if (c.err_code) { rc = 10; break; }
        struct NodeRange {                       // one NVTX range per distributed node (a no-op without a profiler)
            NodeRange(int node_, int N_)
            {
                char label[56];
                snprintf(label, sizeof label, "slab node %d N=%d", node_, N_);
                nvtxRangePushA(label);
            }
            ~NodeRange() { nvtxRangePop(); }
        } nvtx_range(node, cy.geom.back().N);
        if (node == -1) {
            int step, next_N;
            if (con_step == 0) { if (!have(1)) { rc = 3; break; } step = next_int(); } else step = con_step;
            if (con_N == 0) { if (!have(1)) { rc = 3; break; } next_N = next_int(); }
            else { if (pos + 1 >= ladder.size()) { rc = 4; break; } next_N = ladder[++pos]; }
            if (step == 0) continue;
            // (agglomerated levels never get here: the sub-cycle interpreter above took the node)
            const int li = (int)cy.geom.size() - 1;
            const LevelGeom fine = cy.geom[li];
            const bool zero_init = !cy.restart_top();
            LevelGeom coarse;
            if (!cy.induce(fine, next_N, coarse, why)) { fail(-40, why); rc = 20; break; }
            const unsigned int gather_no = coarse.dist ? 0u : ++fab.gathers;
            cy.push(coarse, nullptr, !coarse.dist, gather_no);

            const std::vector<int> cb = coarse_bounds(fine, next_N, comm.world);
            std::vector<double *> fc_full;
            if (!coarse.dist)
                for (auto &st : cy.ranks) fc_full.push_back(st.lv[li + 1].F);
            int done = step;
            double err_host = 0.0;
            bool host_err = false;
            if (step == -1) {           // :194-240: trigger loop; the pass that ends it has also restricted its residual
                trigger_loop(li, zero_init ? 1 : 0, fine.N, next_N, true, coarse.dist, cb, fc_full, 0, no_uc, no_slab, done, err_host);
                host_err = true;
            } else {
                // step > 0: passes of at most 3 sweeps, the last one also restricts.  step < -1 (:244-259): doSmoothing with a
                // negative count = no sweep; the grid is still zeroed, the error evaluated, the residual restricted.
                const int sweeps = step > 0 ? step : 0;
                const int n_pass = std::max(1, (sweeps + 2) / 3), idx = next_scal();
                if (sweeps == 0 && zero_init)
                    for (auto &st : cy.ranks)
                        check(cudaMemsetAsync(st.lv[li].U, 0, (size_t)st.lv[li].slab.rows * fine.N * sizeof(double), c.stream), "memset");
                for (int k = 0; k < n_pass; ++k) {
                    const int S = sweeps / n_pass + (k < sweeps % n_pass ? 1 : 0);
                    const bool first = k == 0, last = k + 1 == n_pass;
                    const int in_mode = (first && zero_init && sweeps > 0) ? 1 : 0;
                    if (last) cy.pass(li, L, S, in_mode, true, idx, next_N, true, coarse.dist, cb, fc_full, 0, no_uc, no_slab);
                    else      cy.pass(li, L, S, in_mode, false, idx, 0, false, false, no_cb, no_fc, 0, no_uc, no_slab);
                }
                deferred.push_back({record(-1, fine.N, done, 0.0), idx});
            }
            if (host_err) record(-1, fine.N, done, err_host);
            if (!coarse.dist) cy.gather_all(next_N, cb, gather_no);
            if (!quiet) fputs(kRestrictArt, stdout);
            mark(coarse.dist ? "down" : "down+gather", fine.N);
        } else if (node == 0) {
            if (!have(2)) { rc = 3; break; }
            // an exact solve on an agglomerated level is part of the sub-cycle taken above
            fail(-41, "the exact solver runs on an agglomerated level: lower the coarsest size or raise the threshold");
            rc = 21;
            break;
        } else if (node == 1) {
            int step;
            if (con_step == 0) { if (!have(1)) { rc = 3; break; } step = next_int(); } else step = con_step;
            if (con_N != 0 && pos > 0) --pos;
            if (cy.geom.size() < 2) { rc = 5; break; }
            const int lc = (int)cy.geom.size() - 1, lf = lc - 1;
            const LevelGeom coarse = cy.geom[lc], fine = cy.geom[lf];
            if (!fine.dist) { rc = 7; break; }   // cannot happen: levels below an agglomerated one belong to the sub-cycle

            // U_c: the rank's coarse slab (halos current by construction) or its own full copy of the agglomerated level
            std::vector<const double *> uc(cy.ranks.size(), nullptr);
            std::vector<Slab> uc_slab(cy.ranks.size());
            for (size_t i = 0; i < cy.ranks.size(); ++i) { uc[i] = cy.ranks[i].lv[lc].U; uc_slab[i] = cy.ranks[i].lv[lc].slab; }

            if (step == -1) {            // prolongation + the trigger loop, two sweeps per pass
                int done = 0;
                double err_host = 0.0;
                trigger_loop(lf, 2, fine.N, 0, false, false, no_cb, no_fc, coarse.N, uc, uc_slab, done, err_host);
                record(1, fine.N, done, err_host);
                if (!quiet) fputs(kProlongArt, stdout);
                cy.pop();
                mark(coarse.dist ? "up" : "up (from the sub-cycle)", fine.N);
                continue;
            }
            const int fixed = step > 0 ? step : 0;
            const int n_pass = std::max(1, (fixed + 2) / 3), idx = next_scal();
            for (int k = 0; k < n_pass; ++k) {
                const int S = fixed / n_pass + (k < fixed % n_pass ? 1 : 0);
                const bool first = k == 0, last = k + 1 == n_pass;
                if (first) cy.pass(lf, L, S, 2, last && step > 0, idx, 0, false, false, no_cb, no_fc, coarse.N, uc, uc_slab);
                else       cy.pass(lf, L, S, 0, last && step > 0, idx, 0, false, false, no_cb, no_fc, 0, no_uc, no_slab);
            }
            if (step > 0) {
                deferred.push_back({record(1, fine.N, step, 0.0), idx});
            } else if (step < -1) {              // :410-421: no sweep, the error is still evaluated
                const int j = next_scal();
                cy.pass(lf, L, 0, 0, true, j, 0, false, false, no_cb, no_fc, 0, no_uc, no_slab);
                deferred.push_back({record(1, fine.N, step, 0.0), j});
            } else record(1, fine.N, 0, 0.0);
            if (!quiet) fputs(kProlongArt, stdout);
            cy.pop();
            mark(coarse.dist ? "up" : "up (from the sub-cycle)", fine.N);
        } else { rc = 6; break; }
    }
    cudaEventRecord(ev1, c.stream);
    if (c.err_code && rc == 0) rc = 10;
    if (rc) fab.abort_peers(c.stream);           // the other ranks must not wait for passes that will never come
    unsigned int *aborted = (unsigned int *)(c.slots_host + (MG_SCALAR_SLOTS - 8));   // pinned: read back with the batch's sums, no extra sync
    *aborted = 0;
    if (comm.world > 1)                          // did a peer give up (or a gather wait time out)?
        check(cudaMemcpyAsync(aborted, fab.peer[cy.ranks[0].rank].words + W_ABORT, sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream), "D2H abort word");
    harvest();
    if (*aborted && rc == 0) { fail(-42, *aborted == 2 ? "slab driver: timed out waiting for a peer's rows" : "slab driver: a peer rank reported an error"); rc = 31; }
    const auto wall1 = std::chrono::steady_clock::now();
    for (size_t i = 1; i < marks.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, marks[i - 1].ev, marks[i].ev);
        if (comm.is_local(0))
            fprintf(stderr, "[mg trace] %-24s N=%-6d device %8.3f ms   host enqueue %8.3f ms\n", marks[i].what, marks[i].N, ms,
                    std::chrono::duration<double, std::milli>(marks[i].host - marks[i - 1].host).count());
    }
    for (Mark &m : marks) cudaEventDestroy(m.ev);
    if (c.err_code && rc == 0) rc = 10;

    if (res) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        res->time_ms = ms;
        res->wall_ms = std::chrono::duration<double, std::milli>(wall1 - wall0).count();
        res->launches = (int)(c.launches - launches0);
        res->n_recs = n_recs < max_recs ? n_recs : max_recs;
        res->N = N_max;
        res->mg_error = 0.0;
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);

    if (rc == 0) {
        const LevelGeom &g = cy.geom[0];
        // ---- final report: mean |analytic - U| (:434-445), slab partial sums all-reduced
        if (res && !(flags & MG_RUN_NO_FINAL_ERROR)) {
            for (auto &st : cy.ranks) {
                RankLevel &l = st.lv[0];
                check(cudaMemsetAsync(st.scal + 60, 0, sizeof(double), c.stream), "memset");
                if (!g.dist && st.rank != 0) continue;     // every rank holds the whole grid: count it once
                launch_source(g.N, L, l.W, min_x, min_y, true, l.slab.row0, l.slab.rows);
                const size_t off = (size_t)(l.slab.own_lo - l.slab.row0) * g.N, cnt = (size_t)(l.slab.own_hi - l.slab.own_lo) * g.N;
                launch_mean_abs_diff(cnt, l.W + off, l.U + off, 1.0, st.scal + 60);
            }
            const double s = reduced_now(60);
            res->mg_error = s / ((double)g.N * (double)g.N);
        }
        // ---- solution out
        for (auto &st : cy.ranks) {
            RankLevel &l = st.lv[0];
            const size_t off = (size_t)(l.slab.own_lo - l.slab.row0) * g.N, cnt = (size_t)(l.slab.own_hi - l.slab.own_lo) * g.N;
            if (io.U_host_full) check(cudaMemcpyAsync(io.U_host_full + (size_t)l.slab.own_lo * g.N, l.U + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H U");
            if (io.U_host_own) check(cudaMemcpyAsync(io.U_host_own, l.U + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H U");
            if (io.U_dev_own) check(cudaMemcpyAsync(io.U_dev_own, l.U + off, cnt * sizeof(double), cudaMemcpyDeviceToDevice, c.stream), "D2D U");
            if (io.own_lo) *io.own_lo = l.slab.own_lo;
            if (io.own_hi) *io.own_hi = l.slab.own_hi;
        }
        check(cudaStreamSynchronize(c.stream), "sync");
        if (!quiet && res) {
            printf("\n\n===== Final Result =====\n    Error = %lf\nTime Used = %lf (ms)\n", res->mg_error, res->wall_ms);
        }
    }
    return rc;
}

Fabric *nccl_fabric()
{
    if (!g_nccl_comm) return nullptr;
    if (!g_fabric) g_fabric.reset(new Fabric(*g_nccl_comm));
    return g_fabric.get();
}

}  // namespace

void dist_release_on_shutdown()
{
    g_fabric.reset();
    g_src = SourceCache();                      // the pool that held the slab is being destroyed by the caller
}

}  // namespace mg

using namespace mg;

extern "C" {

int mgDistEmuRunCycleFile(const char *path, int world, int threshold, int flags, const double *F_host, double *U_host, mgTraceRec *recs,
                          int max_recs, mgCycleResult *res)
{
    if (!ensure_ready()) return 10;
    if (world < 1 || world > MAX_WORLD) return 11;
    EmuComm comm(world);
    Fabric fab(comm);
    if (!fab.ok) return 30;
    RunIo io;
    io.F_host_full = F_host;
    io.U_host_full = U_host;
    return run_dist(comm, fab, path, threshold, flags, recs, max_recs, res, io);
}

int mgDistPlan(const int *ladder, int n_levels, int world, int threshold, int *out, int max_out)
{
    // out: per level [N, dist, bound[0..world]] = world + 3 ints; pure host computation (no GPU needed)
    if (n_levels < 1 || world < 1) return -1;
    const int stride = world + 3;
    if (max_out < n_levels * stride) return -2;
    LevelGeom g = top_geometry(ladder[0], world, threshold);
    for (int l = 0; l < n_levels; ++l) {
        if (l > 0) {
            LevelGeom c;
            std::string why;
            if (!induce_geometry(g, ladder[l], world, threshold, c, why)) return -(10 + l);
            g = c;
        }
        int *o = out + (size_t)l * stride;
        o[0] = g.N;
        o[1] = g.dist ? 1 : 0;
        for (int k = 0; k <= world; ++k) o[2 + k] = g.dist ? g.bound[k] : (k == 0 ? 0 : g.N);
    }
    return n_levels;
}

int mgDistUniqueId(void *out128)
{
    if (!g_nccl.load()) return 1;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { fail(-33, "ncclGetUniqueId failed"); return 2; }
    memcpy(out128, &id, sizeof id);
    return 0;
}

int mgDistInit(int rank, int world, const void *id128)
{
    if (!ensure_ready()) return 10;
    if (!g_nccl.load()) return 1;
    if (g_nccl_comm) return 0;
    if (world > MAX_WORLD) { fail(-35, "mgDistInit: at most 16 ranks"); return 3; }
    std::unique_ptr<NcclComm> cm(new NcclComm());
    cm->world = world;
    cm->rank = rank;
    cm->local = {rank};
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    if (!cm->ok(g_nccl.CommInitRank(&cm->comm, world, id, rank), "ncclCommInitRank")) return 2;
    if (!cm->reserve((size_t)(MAX_WORLD + 1) * 256)) { fail(-38, "mgDistInit: cudaMalloc"); return 4; }
    g_nccl_comm = std::move(cm);
    return nccl_fabric()->ok ? 0 : 30;          // flag words allocated and mapped on every rank (collective)
}

int mgDistSourceSlab(int N, int threshold, int *row0, int *rows, int *own_lo, int *own_hi)
{
    if (!g_nccl_comm) { fail(-34, "mgDistSourceSlab: call mgDistInit first"); return 12; }
    const LevelGeom g = top_geometry(N, g_nccl_comm->world, threshold);
    const Slab s = slab_of(g, g_nccl_comm->rank);
    *row0 = s.row0; *rows = s.rows; *own_lo = s.own_lo; *own_hi = s.own_hi;
    return 0;
}

int mgDistUploadSource(int N, int threshold, const double *F_slab_host)
{
    if (!ensure_ready()) return 10;
    int row0, rows, lo, hi;
    if (mgDistSourceSlab(N, threshold, &row0, &rows, &lo, &hi)) return 12;
    if (rows == 0) return 0;
    const bool match = g_src.F && g_src.N == N && g_src.world == g_nccl_comm->world && g_src.rank == g_nccl_comm->rank &&
                       g_src.row0 == row0 && g_src.rows == rows;
    if (!match) {
        drop_source_cache();
        g_src.F = (double *)pool_alloc((size_t)rows * N * sizeof(double));
        g_src.N = N; g_src.world = g_nccl_comm->world; g_src.rank = g_nccl_comm->rank; g_src.row0 = row0; g_src.rows = rows;
    }
    g_src.from_host = true;
    check(cudaMemcpyAsync(g_src.F, F_slab_host, (size_t)rows * N * sizeof(double), cudaMemcpyHostToDevice, ctx().stream), "H2D source slab");
    return 0;
}

int mgDistDownloadSource(int N, double *F_slab_host)
{
    if (!ensure_ready() || !g_src.F || g_src.N != N) return 1;
    check(cudaMemcpyAsync(F_slab_host, g_src.F, (size_t)g_src.rows * N * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream), "D2H source slab");
    check(cudaStreamSynchronize(ctx().stream), "sync");
    return 0;
}

// doSmoothing on row slabs, repeated: `reps` times `step` Jacobi sweeps (passes of <= 3 fused sweeps, halo rows stored
// into the neighbours' slabs by every pass, error all-reduced once per repetition) on the analytic source grid,
// starting from U = 0.  BASELINE config 5's smoothing-only stress; N up to 65536 (64-bit indexing).
// Works without mgDistInit on one GPU.  U_own_host (optional) receives the rank's owned rows.
int mgDistSmoothStress(int N, double L, int step, int reps, double *ms_per_rep, double *error_out, double *U_own_host, int *own_lo,
                       int *own_hi)
{
    if (!ensure_ready()) return 10;
    if (N % 2 || N < 64 || step < 1 || reps < 1) return 1;
    EmuComm solo(1);
    std::unique_ptr<Fabric> solo_fab;
    Comm &comm = g_nccl_comm ? static_cast<Comm &>(*g_nccl_comm) : static_cast<Comm &>(solo);
    if (!g_nccl_comm) solo_fab.reset(new Fabric(solo));
    Fabric &fab = g_nccl_comm ? *nccl_fabric() : *solo_fab;
    const int rank = g_nccl_comm ? g_nccl_comm->rank : 0;
    Context &c = ctx();
    const LevelGeom g = top_geometry(N, comm.world, 0);
    if (comm.world > 1 && !g.dist) return 2;
    // arena: U and W slabs of every rank (F is local); a local allocation failure travels through the collective reserve
    size_t need = 0;
    for (int r = 0; r < comm.world; ++r) need = std::max(need, 2 * slab_bytes(g, r));
    if (!fab.reserve(fab.stack_base() + need, fab.gather_bytes)) return 3;
    std::fill(fab.top.begin(), fab.top.end(), fab.stack_base());
    std::vector<size_t> bytes((size_t)comm.world);
    for (int r = 0; r < comm.world; ++r) bytes[r] = slab_bytes(g, r);
    std::vector<size_t> offU = fab.alloc(bytes), offW = fab.alloc(bytes);
    const Slab sl = slab_of(g, rank);
    const size_t sbytes = (size_t)sl.rows * N * sizeof(double);
    double *U = fab.at<double>(rank, offU[rank]), *W = fab.at<double>(rank, offW[rank]);
    double *F = (double *)pool_alloc(sbytes), *scal = (double *)pool_alloc(64 * sizeof(double));
    const int mem_ok = comm.allreduce_min_host(F && scal ? 1 : 0);
    if (!mem_ok) { pool_free(F); pool_free(scal); return 3; }
    launch_source(N, L, F, 0.0, 0.0, false, sl.row0, sl.rows);
    check(cudaMemsetAsync(U, 0, sbytes, c.stream), "memset");
    check(cudaMemsetAsync(W, 0, sbytes, c.stream), "memset");
    if (comm.world > 1) { check(cudaStreamSynchronize(c.stream), "sync"); comm.allreduce_min_host(1); }   // every rank's arrays are zeroed before any peer store

    const int n_pass = (step + 2) / 3;
    int rot = 0;
    auto one_rep = [&]() {
        if (++rot == 48) { check(cudaStreamSynchronize(c.stream), "sync"); rot = 1; }   // scal slots still queued for an all-reduce
        for (int k = 0; k < n_pass; ++k) {
            const int S = step / n_pass + (k < step % n_pass ? 1 : 0);
            const bool last = k + 1 == n_pass;
            const unsigned int prev = fab.seq, mine = ++fab.seq;
            PeerLinks pl;
            if (rank > 0) fab.wait32(c.stream, fab.peer[rank].words + W_FROM_LO, prev);
            if (rank + 1 < comm.world) fab.wait32(c.stream, fab.peer[rank].words + W_FROM_HI, prev);
            pl.flag_val = mine;
            if (rank > 0) {
                pl.flag_lo = fab.peer[rank - 1].words + W_FROM_HI;
                pl.U_lo = fab.at<double>(rank - 1, offW[rank - 1]) - (ptrdiff_t)slab_of(g, rank - 1).row0 * N;
            }
            if (rank + 1 < comm.world) {
                pl.flag_hi = fab.peer[rank + 1].words + W_FROM_LO;
                pl.U_hi = fab.at<double>(rank + 1, offW[rank + 1]) - (ptrdiff_t)slab_of(g, rank + 1).row0 * N;
            }
            pl.u_lo_end = sl.own_lo + HALO;
            pl.u_hi_begin = sl.own_hi - HALO;
            slab_pass(N, L, S, 0, U, F, W, sl, last, scal + rot, 0, nullptr, nullptr, 0, nullptr, nullptr, pl);
            std::swap(U, W);
            std::swap(offU, offW);
        }
        if (comm.world > 1) comm.allreduce_sum({scal + rot}, 1, c.stream);
    };
    one_rep();                                   // warm-up (also NCCL connection set-up)
    check(cudaStreamSynchronize(c.stream), "sync");
    if (comm.world > 1) comm.allreduce_min_host(1);   // the warm-up's last halo rows have landed everywhere before the reset
    check(cudaMemsetAsync(U, 0, sbytes, c.stream), "memset");
    check(cudaStreamSynchronize(c.stream), "sync");
    if (comm.world > 1) comm.allreduce_min_host(1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, c.stream);
    for (int r = 0; r < reps; ++r) one_rep();
    cudaEventRecord(e1, c.stream);
    double s = 0.0;
    check(cudaMemcpyAsync(&s, scal + rot, sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H");
    check(cudaStreamSynchronize(c.stream), "sync");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_rep) *ms_per_rep = ms / reps;
    if (error_out) *error_out = (s + s) / (double)N / (double)N;
    if (U_own_host) {
        const size_t off = (size_t)(sl.own_lo - sl.row0) * N, cnt = (size_t)(sl.own_hi - sl.own_lo) * N;
        check(cudaMemcpyAsync(U_own_host, U + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H U");
        check(cudaStreamSynchronize(c.stream), "sync");
    }
    if (own_lo) *own_lo = sl.own_lo;
    if (own_hi) *own_hi = sl.own_hi;
    if (comm.world > 1) comm.allreduce_min_host(1);   // nobody frees while a neighbour may still store into its halos
    pool_free(F); pool_free(scal);
    return c.err_code ? 10 : 0;
}

void mgDistShutdown(void)
{
    if (ctx().ready) {
        cudaStreamSynchronize(ctx().stream);
        drop_source_cache();
    }
    g_fabric.reset();
    if (g_nccl_comm) {
        g_nccl.CommDestroy(g_nccl_comm->comm);
        g_nccl_comm.reset();
    }
}

int mgDistRunCycleFile(const char *path, int threshold, int flags, double *U_own_host, int *own_lo, int *own_hi,
                       mgTraceRec *recs, int max_recs, mgCycleResult *res)
{
    if (!ensure_ready()) return 10;
    if (!g_nccl_comm) { fail(-34, "mgDistRunCycleFile: call mgDistInit first"); return 12; }
    RunIo io;
    io.U_host_own = U_own_host;
    io.own_lo = own_lo;
    io.own_hi = own_hi;
    return run_dist(*g_nccl_comm, *nccl_fabric(), path, threshold, flags, recs, max_recs, res, io);
}

// n independent problems through the same cycle file on the slabs of mgDistInit, HOST buffers per rank: F_slab_hosts[i] =
// this rank's source rows [row0, row0+rows) of problem i (mgDistSourceSlab), U_own_hosts[i] receives its owned rows.
// Double-buffered like mgRunCycleFileHostBatch: the upload of problem i+1 and the download of problem i-1 run on two copy
// streams while the cycle of problem i computes.  Collective; bit-identical to n (mgDistUploadSource, mgDistRunCycleFile) pairs.
int mgDistRunCycleFileHostBatch(const char *path, int threshold, int flags, int n, const double *const *F_slab_hosts,
                                double *const *U_own_hosts, mgCycleResult *res)
{
    if (!ensure_ready()) return 10;
    if (!g_nccl_comm) { fail(-34, "mgDistRunCycleFileHostBatch: call mgDistInit first"); return 12; }
    if (n < 1 || !F_slab_hosts || !U_own_hosts) return 11;
    std::ifstream f(path);
    if (!f.is_open()) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
    double L, mx, my; int cs, cn, N_max, N_min;
    f >> L >> mx >> my >> cs >> cn >> N_max >> N_min;
    if (!f) return 2;
    f.close();
    int row0, rows, lo, hi;
    if (mgDistSourceSlab(N_max, threshold, &row0, &rows, &lo, &hi)) return 12;
    const size_t f_bytes = (size_t)rows * N_max * sizeof(double), u_bytes = (size_t)(hi - lo) * N_max * sizeof(double);
    cudaStream_t compute = ctx().stream, up = nullptr, down = nullptr;
    cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking);
    double *dF[2] = {(double *)pool_alloc(f_bytes), n > 1 ? (double *)pool_alloc(f_bytes) : nullptr};
    double *dU[2] = {(double *)pool_alloc(u_bytes), n > 1 ? (double *)pool_alloc(u_bytes) : nullptr};
    std::vector<cudaEvent_t> uploaded((size_t)n), downloaded((size_t)n);
    for (int i = 0; i < n; ++i) {
        cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&downloaded[i], cudaEventDisableTiming);
    }
    // a local allocation failure must not leave the other ranks alone in the collective cycles: agree first
    int rc = g_nccl_comm->allreduce_min_host((dF[0] && dU[0] && (n == 1 || (dF[1] && dU[1]))) ? 1 : 0) ? 0 : 12;
    auto upload = [&](int i) {      // the previous reader of dF[i % 2], cycle i-2, has been waited for on the host
        cudaMemcpyAsync(dF[i % 2], F_slab_hosts[i], f_bytes, cudaMemcpyHostToDevice, up);
        cudaEventRecord(uploaded[i], up);
    };
    cudaStreamSynchronize(compute);
    if (rc == 0) upload(0);
    for (int i = 0; i < n && rc == 0; ++i) {
        if (i + 1 < n) upload(i + 1);
        cudaStreamWaitEvent(compute, uploaded[i], 0);
        if (i >= 2) cudaStreamWaitEvent(compute, downloaded[i - 2], 0);      // dU[i % 2] is free again
        RunIo io;
        io.F_slab_dev = dF[i % 2];
        io.U_dev_own = dU[i % 2];
        rc = run_dist(*g_nccl_comm, *nccl_fabric(), path, threshold, flags | MG_RUN_SKIP_SOURCE, nullptr, 0, res ? res + i : nullptr, io);   // returns synchronised
        if (rc) break;
        cudaMemcpyAsync(U_own_hosts[i], dU[i % 2], u_bytes, cudaMemcpyDeviceToHost, down);
        cudaEventRecord(downloaded[i], down);
    }
    cudaStreamSynchronize(up);
    cudaStreamSynchronize(down);
    if (rc == 0 && cudaGetLastError() != cudaSuccess) rc = 13;
    for (int i = 0; i < n; ++i) { cudaEventDestroy(uploaded[i]); cudaEventDestroy(downloaded[i]); }
    cudaStreamDestroy(up);
    cudaStreamDestroy(down);
    for (int k = 0; k < 2; ++k) { pool_free(dF[k]); pool_free(dU[k]); }
    return rc;
}

}  // extern "C"
