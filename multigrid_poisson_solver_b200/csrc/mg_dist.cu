// mg_dist.cu -- row-slab multi-GPU cycle driver.
//
// Fine levels are partitioned into contiguous row slabs, one per rank (one process per GPU);
// every slab carries HALO extra rows on each side.  A fused pass needs the halo rows of its
// input to be current; they are refreshed WHEN AN ARRAY IS PRODUCED: right behind every pass one
// grouped send/recv carries the edge rows of its output (and of the restricted F_c) to both
// neighbours, so no pass exchanges before it reads and a node costs one communication group.
// The coarse partition is INDUCED by the fine one through the reference's floor map (a rank owns
// the coarse rows whose lower fine row it owns), so restriction needs no communication at all and
// prolongation only the coarse halo.  Levels below `threshold` rows are agglomerated on rank 0
// (gather of F_c on the way down, scatter of U_c with halos on the way up) and handed as one
// sub-cycle to the single-GPU interpreter; the other ranks idle there.
// The smoothing error is the all-reduced sum of the slabs' red-parity sums (one collective per
// batch of nodes; per sweep only in the trigger mode).
//
// Transports of the row transfers: NCCL send/recv groups (default), or -- MG_DIST_TRANSPORT=staged --
// copy engines writing into fixed staging buffers of the destination rank (opened once through CUDA
// IPC) with sequence flags handled by stream memory operations: no SM is needed, so the exchange
// really overlaps a persistent compute kernel (StagedTransport below).
//
// Two communicators implement the same three primitives (point-to-point row transfers, scalar
// all-reduce):
//   EmuComm   all ranks live in this process on the current GPU (device-to-device copies).  It
//             exists so the whole slab logic is testable bit-for-bit on one GPU.
//   NcclComm  one rank per process; ncclSend/ncclRecv groups and ncclAllReduce on the library's
//             stream.  NCCL is bound at run time (dlopen of the libnccl.so.2 the host program --
//             e.g. torch -- already loaded), so the library has no link-time NCCL dependency.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
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

struct Xfer {                     // `count` doubles from src (valid on src_rank) to dst (valid on dst_rank)
    int src_rank, dst_rank;
    const double *src;
    double *dst;
    size_t count;
};

class Comm {
public:
    virtual ~Comm() {}
    int world = 1;
    std::vector<int> local;       // ranks living in this process
    bool is_local(int r) const { return std::find(local.begin(), local.end(), r) != local.end(); }
    // both primitives are queued on `stream`; one communicator must only ever be driven from ONE
    // stream at a time (NCCL operations of a communicator may not run concurrently)
    virtual void transfer(const std::vector<Xfer> &xs, cudaStream_t stream) = 0;
    // vals[i] = device pointer of local rank i; on return every one holds the sum over all ranks
    virtual void allreduce_sum(const std::vector<double *> &vals, int n, cudaStream_t stream) = 0;
};


// ------------------------------------------------------------------ staged peer-to-peer transport
// Row transfers without communication kernels.  Every rank owns a staging area with one slot per
// (source rank, message parity) and a few 32-bit words: flag[src] (sequence number of the last
// message src has delivered here), ack[dst] (last message of mine dst has consumed).  A message
// from s to d is
//     s:  wait  ack[d] >= seq - 2            (the slot of this parity is free again)
//         copy  rows -> d.stage[s][seq & 1]   (peer write by a copy engine, over NVLink)
//         write d.flag[s] = seq               (4-byte copy behind the data, same stream: ordered)
//     d:  wait  flag[s] >= seq                (stream memory operation: no SM, no host)
//         copy  stage[s][seq & 1] -> rows     (local)
//         write s.ack[d] = seq
// Sequence numbers are counted per ordered pair on both sides; every rank walks the same transfer
// lists in the same order, so they agree without any handshake.  All sends of a call are queued
// before its receives, hence no cycle of waits.  The staging areas of the other ranks are mapped
// once (CUDA IPC when they live in other processes).
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

class StagedTransport {
public:
    static constexpr size_t CAP = 8u << 20;      // bytes per (source rank, parity) slot
    struct Local {                                // one per rank living in this process
        int rank = -1;
        unsigned char *stage = nullptr;           // [world][2][CAP]
        unsigned int *words = nullptr;            // flag[world] | ack[world] | scratch[2]
        std::vector<unsigned char *> peer_stage;  // every rank's staging area as seen from here
        std::vector<unsigned int *> peer_words;
        std::vector<unsigned int> seq_out, seq_in;
    };
    int world = 0;
    std::vector<Local> locals;
    bool ready = false;

    ~StagedTransport()
    {
        for (Local &l : locals) {
            for (int q = 0; q < world; ++q) {
                if (q == l.rank || ipc_opened.empty()) continue;
                if (ipc_opened[q]) { cudaIpcCloseMemHandle(l.peer_stage[q]); cudaIpcCloseMemHandle(l.peer_words[q]); }
            }
            cudaFree(l.stage);
            cudaFree(l.words);
        }
    }
    std::vector<char> ipc_opened;                 // per rank: its areas were opened through IPC (one local rank only)

    size_t words_count() const { return (size_t)2 * world + 2; }
    bool add_local(int rank, int world_)
    {
        world = world_;
        Local l;
        l.rank = rank;
        if (cudaMalloc(&l.stage, (size_t)world * 2 * CAP) != cudaSuccess) return false;
        if (cudaMalloc(&l.words, words_count() * sizeof(unsigned int)) != cudaSuccess) return false;
        if (cudaMemset(l.words, 0, words_count() * sizeof(unsigned int)) != cudaSuccess) return false;
        l.peer_stage.assign((size_t)world, nullptr);
        l.peer_words.assign((size_t)world, nullptr);
        l.seq_out.assign((size_t)world, 0u);
        l.seq_in.assign((size_t)world, 0u);
        locals.push_back(l);
        return true;
    }
    // all ranks in this process: the "peer" views are the other ranks' own pointers
    bool link_in_process()
    {
        for (Local &a : locals)
            for (Local &b : locals) { a.peer_stage[b.rank] = b.stage; a.peer_words[b.rank] = b.words; }
        ready = g_drv.load();
        return ready;
    }
    Local *local_of(int rank)
    {
        for (Local &l : locals) if (l.rank == rank) return &l;
        return nullptr;
    }
    // same answer on every rank: it depends on the transfer sizes only
    bool applicable(const std::vector<Xfer> &xs) const
    {
        if (!ready) return false;
        std::map<std::pair<int, int>, size_t> bytes;
        for (const Xfer &x : xs)
            if (x.count && x.src_rank != x.dst_rank) bytes[{x.src_rank, x.dst_rank}] += x.count * sizeof(double);
        for (const auto &kv : bytes) if (kv.second > CAP) return false;
        return true;
    }
    void wait32(cudaStream_t st, unsigned int *addr, unsigned int value)
    {
        if (g_drv.WaitValue32((CUstream)st, (CUdeviceptr)addr, value, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) fail(-36, "cuStreamWaitValue32");
    }
    // *remote = value, ordered behind everything queued on st so far (local scratch word, then a 4-byte copy)
    void post32(cudaStream_t st, Local &l, int scratch, unsigned int *remote, unsigned int value)
    {
        unsigned int *w = l.words + 2 * world + scratch;
        if (g_drv.WriteValue32((CUstream)st, (CUdeviceptr)w, value, 0) != CUDA_SUCCESS) fail(-36, "cuStreamWriteValue32");
        check(cudaMemcpyAsync(remote, w, sizeof(unsigned int), cudaMemcpyDefault, st), "staged flag copy");
    }
    void transfer(const std::vector<Xfer> &xs, cudaStream_t st)
    {
        struct Item { const Xfer *x; size_t off; };
        std::map<std::pair<int, int>, std::vector<Item>> groups;    // ordered: the same walk on every rank
        std::map<std::pair<int, int>, size_t> fill;
        for (const Xfer &x : xs) {
            if (!x.count) continue;
            if (x.src_rank == x.dst_rank) {
                if (local_of(x.src_rank)) check(cudaMemcpyAsync(x.dst, x.src, x.count * sizeof(double), cudaMemcpyDeviceToDevice, st), "self transfer");
                continue;
            }
            size_t &off = fill[{x.src_rank, x.dst_rank}];
            groups[{x.src_rank, x.dst_rank}].push_back({&x, off});
            off += x.count * sizeof(double);
        }
        for (auto &kv : groups) {                  // ---- sends
            const int s = kv.first.first, d = kv.first.second;
            Local *l = local_of(s);
            if (!l) continue;
            const unsigned int seq = ++l->seq_out[d];
            if (seq > 2) wait32(st, l->words + world + d, seq - 2);
            unsigned char *slot = l->peer_stage[d] + ((size_t)s * 2 + (seq & 1u)) * CAP;
            for (const Item &it : kv.second)
                check(cudaMemcpyAsync(slot + it.off, it.x->src, it.x->count * sizeof(double), cudaMemcpyDefault, st), "staged send");
            post32(st, *l, 0, l->peer_words[d] + s, seq);
        }
        for (auto &kv : groups) {                  // ---- receives
            const int s = kv.first.first, d = kv.first.second;
            Local *l = local_of(d);
            if (!l) continue;
            const unsigned int seq = ++l->seq_in[s];
            wait32(st, l->words + s, seq);
            const unsigned char *slot = l->stage + ((size_t)s * 2 + (seq & 1u)) * CAP;
            for (const Item &it : kv.second)
                check(cudaMemcpyAsync(it.x->dst, slot + it.off, it.x->count * sizeof(double), cudaMemcpyDeviceToDevice, st), "staged receive");
            post32(st, *l, 1, l->peer_words[s] + world + d, seq);
        }
    }
};

bool staged_requested()
{
    const char *e = getenv("MG_DIST_TRANSPORT");
    return e && strcmp(e, "staged") == 0;
}

class EmuComm : public Comm {
public:
    std::unique_ptr<StagedTransport> staged;      // MG_DIST_TRANSPORT=staged: the protocol of the real transport, in one process
    explicit EmuComm(int g)
    {
        world = g;
        for (int r = 0; r < g; ++r) local.push_back(r);
        if (g > 1 && staged_requested()) {
            staged.reset(new StagedTransport());
            bool good = true;
            for (int r = 0; r < g && good; ++r) good = staged->add_local(r, g);
            if (!good || !staged->link_in_process()) { staged.reset(); fail(-37, "staged transport: set-up failed"); }
        }
    }
    ~EmuComm() override { cudaDeviceSynchronize(); }
    void transfer(const std::vector<Xfer> &xs, cudaStream_t stream) override
    {
        if (staged && staged->applicable(xs)) { staged->transfer(xs, stream); return; }
        for (const Xfer &x : xs)
            if (x.count) check(cudaMemcpyAsync(x.dst, x.src, x.count * sizeof(double), cudaMemcpyDeviceToDevice, stream), "emu transfer");
    }
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
};

// ---- NCCL bound at run time
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the host program already uses, if any
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW);
        if (!handle) { fail(-30, std::string("cannot load libnccl.so.2: ") + dlerror()); return false; }
#define MG_SYM(field, name) field = (decltype(field))dlsym(handle, name); if (!field) { fail(-31, "libnccl.so.2 lacks " name); return false; }
        MG_SYM(GetUniqueId, "ncclGetUniqueId") MG_SYM(CommInitRank, "ncclCommInitRank") MG_SYM(CommDestroy, "ncclCommDestroy")
        MG_SYM(Send, "ncclSend") MG_SYM(Recv, "ncclRecv") MG_SYM(AllReduce, "ncclAllReduce") MG_SYM(AllGather, "ncclAllGather") MG_SYM(GroupStart, "ncclGroupStart")
        MG_SYM(GroupEnd, "ncclGroupEnd") MG_SYM(GetErrorString, "ncclGetErrorString")
#undef MG_SYM
        return true;
    }
};
NcclApi g_nccl;

class NcclComm : public Comm {
public:
    ncclComm_t comm = nullptr;
    int rank = 0;
    bool ok(ncclResult_t r, const char *what)
    {
        if (r == ncclSuccess) return true;
        fail(-32, std::string(what) + ": " + g_nccl.GetErrorString(r));
        return false;
    }
    std::unique_ptr<StagedTransport> staged;      // MG_DIST_TRANSPORT=staged

    // Collective: every rank allocates its staging area, the IPC handles are all-gathered, every rank
    // maps the others' areas, and the transport is switched on only if ALL ranks succeeded.
    bool setup_staged()
    {
        struct Pack { cudaIpcMemHandle_t stage, words; };
        cudaStream_t st = ctx().stream;
        std::unique_ptr<StagedTransport> t(new StagedTransport());
        int good = (t->add_local(rank, world) && g_drv.load()) ? 1 : 0;
        Pack mine;
        memset(&mine, 0, sizeof mine);
        if (good) {
            StagedTransport::Local &l = t->locals[0];
            good = cudaIpcGetMemHandle(&mine.stage, l.stage) == cudaSuccess && cudaIpcGetMemHandle(&mine.words, l.words) == cudaSuccess;
        }
        unsigned char *dsend = nullptr, *drecv = nullptr;
        int *dflag = nullptr;
        std::vector<Pack> packs((size_t)world);
        bool coll = cudaMalloc(&dsend, sizeof(Pack)) == cudaSuccess && cudaMalloc(&drecv, (size_t)world * sizeof(Pack)) == cudaSuccess &&
                    cudaMalloc(&dflag, sizeof(int)) == cudaSuccess;
        if (!coll) { fail(-37, "staged transport: cudaMalloc"); return false; }
        cudaMemcpy(dsend, &mine, sizeof(Pack), cudaMemcpyHostToDevice);
        coll = ok(g_nccl.AllGather(dsend, drecv, sizeof(Pack), ncclChar, comm, st), "ncclAllGather (IPC handles)");
        cudaStreamSynchronize(st);
        cudaMemcpy(packs.data(), drecv, (size_t)world * sizeof(Pack), cudaMemcpyDeviceToHost);
        if (good && coll) {
            StagedTransport::Local &l = t->locals[0];
            t->ipc_opened.assign((size_t)world, 0);
            l.peer_stage[rank] = l.stage;
            l.peer_words[rank] = l.words;
            for (int q = 0; q < world && good; ++q) {
                if (q == rank) continue;
                void *a = nullptr, *b = nullptr;
                good = cudaIpcOpenMemHandle(&a, packs[q].stage, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
                       cudaIpcOpenMemHandle(&b, packs[q].words, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                if (good) { l.peer_stage[q] = (unsigned char *)a; l.peer_words[q] = (unsigned int *)b; t->ipc_opened[q] = 1; }
            }
            if (!good) cudaGetLastError();
        }
        int all = (good && coll) ? 1 : 0;
        cudaMemcpy(dflag, &all, sizeof(int), cudaMemcpyHostToDevice);
        ok(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMin, comm, st), "ncclAllReduce (staged transport)");
        cudaStreamSynchronize(st);
        cudaMemcpy(&all, dflag, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(dsend); cudaFree(drecv); cudaFree(dflag);
        if (all) {
            t->ready = true;
            staged = std::move(t);
            if (rank == 0 && getenv("MG_DIST_TRACE")) fprintf(stderr, "[mg trace] staged transport ready on %d ranks (CUDA IPC)\n", world);
        }
        else if (rank == 0) fprintf(stderr, "[ WARNING ]: staged transport unavailable on some rank; using NCCL send/recv\n");
        return all != 0;
    }

    void transfer(const std::vector<Xfer> &xs, cudaStream_t stream) override
    {
        if (staged && staged->applicable(xs)) { staged->transfer(xs, stream); return; }
        ok(g_nccl.GroupStart(), "ncclGroupStart");
        for (const Xfer &x : xs) {
            if (!x.count) continue;
            if (x.src_rank == rank && x.dst_rank == rank) {
                check(cudaMemcpyAsync(x.dst, x.src, x.count * sizeof(double), cudaMemcpyDeviceToDevice, stream), "self transfer");
                continue;
            }
            if (x.src_rank == rank) ok(g_nccl.Send(x.src, x.count, ncclDouble, x.dst_rank, comm, stream), "ncclSend");
            if (x.dst_rank == rank) ok(g_nccl.Recv(x.dst, x.count, ncclDouble, x.src_rank, comm, stream), "ncclRecv");
        }
        ok(g_nccl.GroupEnd(), "ncclGroupEnd");
    }
    void allreduce_sum(const std::vector<double *> &vals, int n, cudaStream_t stream) override
    {
        ok(g_nccl.AllReduce(vals[0], vals[0], (size_t)n, ncclDouble, ncclSum, comm, stream), "ncclAllReduce");
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
    if (!(world > 1 && M >= threshold)) return true;
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

struct RankLevel {            // one rank's share of one level
    Slab slab;
    bool present = false;     // dist: every rank; agglomerated: rank 0 only
    double *U = nullptr, *W = nullptr, *F = nullptr;
    bool owns_F = true;
};

struct RankState {
    int rank = 0;
    std::vector<RankLevel> lv;
    double *scal = nullptr;   // device scalars: [0] error partial, [1] final abs-diff partial
};

// Halo rows of one array of a distributed level: rank k's last HALO owned rows go to rank k+1's
// lower halo, rank k+1's first HALO owned rows to rank k's upper halo.  ptr(rank) = base of that
// rank's local array (nullptr when the rank lives in another process).
template <class Ptr>
void halo_xfers(const LevelGeom &g, const Comm &comm, Ptr ptr, std::vector<Xfer> &xs)
{
    if (!g.dist) return;
    const size_t N = g.N;
    for (int k = 0; k + 1 < comm.world; ++k) {
        if (!comm.is_local(k) && !comm.is_local(k + 1)) continue;
        const Slab a = slab_of(g, k), b = slab_of(g, k + 1);
        double *pa = ptr(k), *pb = ptr(k + 1);
        const int up_lo = std::max(a.own_hi - HALO, b.row0);
        xs.push_back({k, k + 1, pa ? pa + (size_t)(up_lo - a.row0) * N : nullptr, pb ? pb + (size_t)(up_lo - b.row0) * N : nullptr,
                      (size_t)(a.own_hi - up_lo) * N});
        const int dn_hi = std::min(b.own_lo + HALO, a.row0 + a.rows);
        xs.push_back({k + 1, k, pb ? pb + (size_t)(b.own_lo - b.row0) * N : nullptr, pa ? pa + (size_t)(b.own_lo - a.row0) * N : nullptr,
                      (size_t)(dn_hi - b.own_lo) * N});
    }
}

// Stream protocol of the slab driver.  Kernels run on the compute stream `ms`; EVERY communication
// primitive (halo exchange, gather/scatter, all-reduce, the scalar read-backs that follow an
// all-reduce) is queued on the communication stream `cs`, so the communicator sees one ordered
// stream.  A pass is launched in two parts: the edge row segments first -- they produce the rows
// the neighbours' halos need -- then the interior; the halo exchange of the pass OUTPUT is queued
// on cs behind the edge launch and runs while the interior launch computes.  The next pass waits
// for that exchange (wait_halo) before it starts.  Because halos are refreshed when an array is
// produced, no pass ever has to exchange before it reads.
// Opt-in (MG_DIST_OVERLAP=1): measured on 2 x B200 it does not pay with NCCL transfers -- NCCL's
// send/recv kernels need SMs, and the persistent interior launch holds every SM until it ends, so
// the exchange starts late anyway.  By default both streams are the compute stream
// and passes are launched whole, the exchange of the output right behind them.
struct TwoStream {
    cudaStream_t ms = nullptr, cs = nullptr, cs_hi = nullptr;   // cs: where communication goes right now (cs_hi or ms)
    cudaEvent_t ev_join = nullptr, ev_edge = nullptr, ev_halo = nullptr;
    bool overlap = true, halo_pending = false;
    long long min_points = 0;   // slabs smaller than this run their communication in line on ms (MG_DIST_OVERLAP_MIN_POINTS)
    TwoStream()
    {
        ms = ctx().stream;
        cs_hi = ctx().comm_stream;
        const char *e = getenv("MG_DIST_OVERLAP");
        overlap = e && atoi(e) != 0;
        if ((e = getenv("MG_DIST_OVERLAP_MIN_POINTS"))) min_points = atoll(e);
        cs = overlap ? cs_hi : ms;
        cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_edge, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_halo, cudaEventDisableTiming);
    }
    ~TwoStream()
    {
        sync();
        cudaEventDestroy(ev_join); cudaEventDestroy(ev_edge); cudaEventDestroy(ev_halo);
    }
    TwoStream(const TwoStream &) = delete;
    TwoStream &operator=(const TwoStream &) = delete;
    // communication queued after this sees everything the compute stream has been given so far
    void to_comm()
    {
        if (cs == ms) return;
        cudaEventRecord(ev_join, ms);
        cudaStreamWaitEvent(cs, ev_join, 0);
    }
    // kernels queued after this see every communication queued so far
    void to_compute()
    {
        halo_pending = false;
        if (cs == ms) return;
        cudaEventRecord(ev_join, cs);
        cudaStreamWaitEvent(ms, ev_join, 0);
    }
    void halo_posted() { cudaEventRecord(ev_halo, cs); halo_pending = true; }
    void wait_halo()
    {
        if (halo_pending && cs != ms) cudaStreamWaitEvent(ms, ev_halo, 0);
        halo_pending = false;
    }
    void edge_done()   // the exchange queued next may start as soon as the edge launch has finished
    {
        if (cs == ms) return;
        cudaEventRecord(ev_edge, ms);
        cudaStreamWaitEvent(cs, ev_edge, 0);
    }
    void sync()
    {
        check(cudaStreamSynchronize(ms), "sync");
        check(cudaStreamSynchronize(cs_hi), "sync (comm)");
        halo_pending = false;
    }
    // choose where the communication of the next pass goes: overlapped on cs_hi, or in line on ms
    // (small levels: the cross-stream hand-offs cost more than the exchange they would hide)
    void select(bool overlapped)
    {
        cudaStream_t want = (overlap && overlapped) ? cs_hi : ms;
        if (want == cs) return;
        if (cs != ms) to_compute();   // leaving cs_hi: ms continues behind everything queued there
        cs = want;                    // entering cs_hi: every primitive there starts with to_comm()/edge_done()
    }
};

class DistCycle {
public:
    DistCycle(Comm &c, int threshold) : comm(c), threshold_(threshold)
    {
        for (int r : comm.local) {
            RankState st;
            st.rank = r;
            st.scal = (double *)pool_alloc(64 * sizeof(double));
            ranks.push_back(st);
        }
    }
    ~DistCycle()
    {
        ts.sync();                               // nothing queued may outlive the buffers
        while (!geom.empty()) pop();
        for (auto &st : ranks) pool_free(st.scal);
    }

    Comm &comm;
    int threshold_;
    std::vector<LevelGeom> geom;
    std::vector<RankState> ranks;
    int init_ = 1;
    TwoStream ts;
    int scal_used_ = 0;         // slots of the ranks' scal arrays holding error sums of this batch

    void push(const LevelGeom &g, double *borrowed_F = nullptr)
    {
        geom.push_back(g);
        for (auto &st : ranks) {
            RankLevel l;
            l.slab = slab_of(g, st.rank);
            l.present = g.dist || st.rank == 0;
            if (l.present) {
                const size_t bytes = (size_t)l.slab.rows * g.N * sizeof(double);
                l.U = (double *)pool_alloc(bytes);
                l.W = (double *)pool_alloc(bytes);
                l.F = borrowed_F ? borrowed_F : (double *)pool_alloc(bytes);
                l.owns_F = borrowed_F == nullptr;
            }
            st.lv.push_back(l);
        }
    }
    void pop()
    {
        for (auto &st : ranks) {
            RankLevel &l = st.lv.back();
            pool_free(l.U); pool_free(l.W);
            if (l.owns_F) pool_free(l.F);
            st.lv.pop_back();
        }
        geom.pop_back();
        if (geom.size() == 1) init_ = 0;
    }
    bool restart_top() const { return init_ == 0 && geom.size() == 1; }

    // coarse geometry induced by the fine one (see the header comment)
    bool induce(const LevelGeom &fine, int M, LevelGeom &coarse, std::string &why)
    {
        return induce_geometry(fine, M, comm.world, threshold_, coarse, why);
    }

    // ---- halo transfers of one array of level `li` (which: 0 U, 1 F, 2 W)
    void add_halo(int li, int which, std::vector<Xfer> &xs)
    {
        halo_xfers(geom[li], comm, [&](int rank) -> double * {
            for (auto &st : ranks)
                if (st.rank == rank) return which == 0 ? st.lv[li].U : which == 1 ? st.lv[li].F : st.lv[li].W;
            return nullptr;
        }, xs);
    }

    // sum the ranks' partials at scal[idx .. idx+n); every local rank ends with the global sums (on ts.cs)
    void allreduce(int idx, int n = 1)
    {
        ts.to_comm();
        if (comm.world == 1) return;
        std::vector<double *> v;
        for (auto &st : ranks) v.push_back(st.scal + idx);
        comm.allreduce_sum(v, n, ts.cs);
    }

    // gather / scatter between rank 0 and the slabs: compute -> comm -> compute
    void transfer_now(const std::vector<Xfer> &xs)
    {
        ts.to_comm();
        comm.transfer(xs, ts.cs);
        ts.to_compute();
    }

    // One fused pass over every local slab of distributed level `li`: S sweeps from U (in_mode 0),
    // from zero (1) or from U + prolongation of the coarse slabs `uc` (2); optional red-parity error
    // sum into scal[scal_idx]; optional restriction of the negated residual into `fc`.  The output
    // (W, then swapped into U) and, with fc_halo, the coarse source get their halos refreshed by an
    // exchange that overlaps the interior part of the pass.
    void pass(int li, double L, int S, int in_mode, bool want_err, int scal_idx, int M, const std::vector<double *> &fc,
              const std::vector<Slab> &fc_slab, bool fc_halo, int Nc, const std::vector<const double *> &uc,
              const std::vector<Slab> &uc_slab)
    {
        const LevelGeom &g = geom[li];
        const bool writes_U = !(S == 0 && in_mode == 0);
        // the edge segments hold >= 24 fine rows per side: enough for 8 coarse halo rows up to ratio 2.5
        ts.select((long long)g.N * (g.N / comm.world) >= ts.min_points);
        const bool split = ts.cs != ts.ms && (!fc_halo || (double)(g.N - 1) <= 2.5 * (double)(M - 1));
        auto launch = [&](int subset) {
            for (size_t i = 0; i < ranks.size(); ++i) {
                RankLevel &l = ranks[i].lv[li];
                slab_pass(g.N, L, S, in_mode, l.U, l.F, writes_U ? l.W : nullptr, l.slab, want_err, ranks[i].scal + scal_idx, M,
                          fc.empty() ? nullptr : fc[i], fc.empty() ? nullptr : &fc_slab[i], Nc, uc.empty() ? nullptr : uc[i],
                          uc.empty() ? nullptr : &uc_slab[i], subset);
            }
        };
        auto post_halo = [&]() {
            std::vector<Xfer> xs;
            if (writes_U) add_halo(li, 2, xs);
            if (fc_halo) add_halo(li + 1, 1, xs);
            if (xs.empty()) return;
            ts.edge_done();
            comm.transfer(xs, ts.cs);
            ts.halo_posted();
        };
        ts.wait_halo();
        if (split) { launch(1); post_halo(); launch(2); }
        else       { launch(0); post_halo(); }
        if (writes_U)
            for (auto &st : ranks) std::swap(st.lv[li].U, st.lv[li].W);
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

const char *kRestrictArt = "             *\n             |\n Restriction |\n             |\n             *\n";
const char *kProlongArt = "             *\n             |\nProlongation |\n             |\n             *\n";

int run_dist(Comm &comm, const char *path, int threshold, int flags, mgTraceRec *recs, int max_recs, mgCycleResult *res,
             double *U_host_full, double *U_host_own, int *own_lo_out, int *own_hi_out)
{
    std::ifstream f(path);
    if (!f.is_open()) { fprintf(stderr, "[ ERROR ]: Cannot open file %s\n", path); return 1; }
    std::vector<double> tok;
    for (double d; f >> d;) tok.push_back(d);
    if (tok.size() < 7) return 2;
    size_t cur = 0;
    auto have = [&](size_t n) { return cur + n <= tok.size(); };
    auto next_int = [&]() { return (int)tok[cur++]; };
    double L, min_x, min_y;
    int con_step, con_N, N_max, N_min;
    L = tok[cur++]; min_x = tok[cur++]; min_y = tok[cur++];
    con_step = next_int(); con_N = next_int(); N_max = next_int(); N_min = next_int();
    std::vector<int> ladder;
    if (con_N == 1) for (int n = N_max; n >= N_min; n /= 2) ladder.push_back(n);
    if (con_N == 2) for (int n = N_max; n >= N_min; --n) ladder.push_back(n);
    size_t pos = 0;
    const bool quiet = (flags & MG_RUN_QUIET) != 0 || !comm.is_local(0);

    DistCycle cy(comm, threshold);
    Context &c = ctx();
    {
        const LevelGeom g = top_geometry(N_max, comm.world, threshold);
        // one-rank-per-process runs may keep the top-level source slab between calls
        const bool cacheable = (flags & MG_RUN_SKIP_SOURCE) && cy.ranks.size() == 1;
        double *borrowed = nullptr;
        if (cacheable) {
            const Slab s0 = slab_of(g, cy.ranks[0].rank);
            const bool hit = g_src.F && g_src.N == N_max && g_src.world == comm.world && g_src.rank == cy.ranks[0].rank &&
                             g_src.row0 == s0.row0 && g_src.rows == s0.rows &&
                             (g_src.from_host || (g_src.L == L && g_src.min_x == min_x && g_src.min_y == min_y));
            if (!hit && (g.dist || cy.ranks[0].rank == 0)) {
                if (g_src.F) pool_free(g_src.F);
                g_src = SourceCache();
                g_src.F = (double *)pool_alloc((size_t)s0.rows * N_max * sizeof(double));
                g_src.N = N_max; g_src.world = comm.world; g_src.rank = cy.ranks[0].rank; g_src.row0 = s0.row0; g_src.rows = s0.rows;
                g_src.L = L; g_src.min_x = min_x; g_src.min_y = min_y;
                launch_source(N_max, L, g_src.F, min_x, min_y, false, s0.row0, s0.rows);
            }
            if (g.dist || cy.ranks[0].rank == 0) borrowed = g_src.F;
        }
        cy.push(g, borrowed);
        if (!borrowed)
            for (auto &st : cy.ranks) {
                RankLevel &l = st.lv[0];
                if (l.present) launch_source(N_max, L, l.F, min_x, min_y, false, l.slab.row0, l.slab.rows);   // halo rows computed locally
            }
    }
    check(cudaStreamSynchronize(c.stream), "sync");

    // Error sums of fixed-step nodes stay on the device as per-rank partials in consecutive scal slots;
    // ONE all-reduce per batch (before an agglomerated sub-cycle, at the end, or when the slots run
    // out) turns them into the trace records -- one collective instead of one per node.
    int n_recs = 0;
    auto record = [&](int node, int N, int steps, double err) {
        const int r = n_recs++;
        if (recs && r < max_recs) { recs[r].node = node; recs[r].N = N; recs[r].steps = steps; recs[r].err = err; }
        return r;
    };
    constexpr int SCAL_BATCH = 48;
    std::vector<std::pair<int, int>> deferred;     // (trace record, scal slot) holding a raw red-parity sum
    auto harvest = [&]() {
        if (cy.scal_used_ > 0) {
            double sums[SCAL_BATCH];
            cy.allreduce(0, cy.scal_used_);
            check(cudaMemcpyAsync(sums, cy.ranks[0].scal, cy.scal_used_ * sizeof(double), cudaMemcpyDeviceToHost, cy.ts.cs), "D2H sums");
            cy.ts.sync();
            for (const auto &d : deferred) {
                if (!recs || d.first >= max_recs) continue;
                const double v = sums[d.second], N = recs[d.first].N;
                recs[d.first].err = (v + v) / N / N;                      // (sum1+sum2)/N/N, :621-622
            }
            deferred.clear();
            cy.scal_used_ = 0;
        }
        cy.ts.sync();
        mgSubcycleHarvest();                     // scalars of the agglomerated sub-cycles (rank 0)
    };
    auto next_scal = [&]() {
        if (cy.scal_used_ == SCAL_BATCH) harvest();
        return cy.scal_used_++;
    };
    // all-reduce scal[idx] and wait for the value on the host (trigger loops)
    auto reduced_now = [&](int idx) {
        cy.allreduce(idx);
        double s = 0.0;
        check(cudaMemcpyAsync(&s, cy.ranks[0].scal + idx, sizeof(double), cudaMemcpyDeviceToHost, cy.ts.cs), "D2H");
        check(cudaStreamSynchronize(cy.ts.cs), "sync");
        return s;
    };
    const std::vector<double *> no_fc;
    const std::vector<const double *> no_uc;
    const std::vector<Slab> no_slab;

    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    const long long launches0 = c.launches;
    const auto wall0 = std::chrono::steady_clock::now();
    cudaEventRecord(ev0, c.stream);

    // MG_DIST_TRACE=1: per-node time line (device time on the compute stream, host enqueue time)
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
        // ---- agglomerated sub-cycle: everything that happens at or below a level held by rank 0
        // alone is run by the single-GPU interpreter (fused nodes + coarse tail kernel) on rank 0;
        // the other ranks parse the same nodes without executing them.
        if (!cy.geom.back().dist && ((int)tok[cur] == -1 || (int)tok[cur] == 0)) {
            const int li = (int)cy.geom.size() - 1;
            int icur = (int)cur, ipos = (int)pos, n_rec_io = n_recs, init_io = cy.init_;
            int sub_rc = 0;
            for (auto &st : cy.ranks) {
                RankLevel &l = st.lv[li];
                int c2 = (int)cur, p2 = (int)pos, r2 = n_recs, i2 = cy.init_;
                sub_rc = mgRunSubcycle(tok.data(), (int)tok.size(), &c2, &p2, ladder.data(), (int)ladder.size(), con_step, con_N, L,
                                       cy.geom[li].N, &l.U, &l.W, l.F, li, &i2, flags | MG_RUN_DEFER_HARVEST,
                                       (st.rank == 0 || !comm.is_local(0)) ? recs : nullptr, max_recs, &r2, st.rank == 0 ? 1 : 0);
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
            cy.push(coarse);

            // F_c target per rank: the coarse slab's F if the coarse level is distributed, else a
            // temporary holding exactly the rank's coarse rows (rank 0 writes into the full array)
            const std::vector<int> cb = coarse_bounds(fine, next_N, comm.world);
            std::vector<double *> fc_tmp(cy.ranks.size(), nullptr), fc(cy.ranks.size(), nullptr);
            std::vector<Slab> fc_slab(cy.ranks.size());
            for (size_t i = 0; i < cy.ranks.size(); ++i) {
                RankState &st = cy.ranks[i];
                fc[i] = st.lv[li + 1].F;
                fc_slab[i] = st.lv[li + 1].slab;
                if (coarse.dist || st.rank == 0) continue;
                fc_slab[i].row0 = fc_slab[i].own_lo = cb[st.rank];
                fc_slab[i].own_hi = cb[st.rank + 1];
                fc_slab[i].rows = cb[st.rank + 1] - cb[st.rank];
                fc[i] = fc_tmp[i] = (double *)pool_alloc((size_t)std::max(1, fc_slab[i].rows) * next_N * sizeof(double));
            }

            if (step > 0) {
                // passes of at most 3 sweeps; the last one also restricts
                const int n_pass = (step + 2) / 3, idx = next_scal();
                for (int k = 0; k < n_pass; ++k) {
                    const int S = step / n_pass + (k < step % n_pass ? 1 : 0);
                    const bool first = k == 0, last = k + 1 == n_pass;
                    if (last) cy.pass(li, L, S, (first && zero_init) ? 1 : 0, true, idx, next_N, fc, fc_slab, coarse.dist, 0, no_uc, no_slab);
                    else      cy.pass(li, L, S, (first && zero_init) ? 1 : 0, false, idx, 0, no_fc, no_slab, false, 0, no_uc, no_slab);
                }
                deferred.push_back({record(-1, fine.N, step, 0.0), idx});
            } else {   // trigger loop: the scalar is needed after every sweep
                double slope = TRIGGER + 1.0, prev = 0.0, err_host = 0.0;
                int done = 0;
                while (slope > TRIGGER) {
                    const int idx = next_scal();
                    cy.pass(li, L, 1, (done == 0 && zero_init) ? 1 : 0, true, idx, 0, no_fc, no_slab, false, 0, no_uc, no_slab);
                    const double s = reduced_now(idx);
                    err_host = (s + s) / fine.N / fine.N;
                    ++done;
                    if (done > 1) slope = std::fabs(err_host - prev);
                    prev = err_host;
                    if (c.err_code) break;
                }
                cy.pass(li, L, 0, 0, false, 0, next_N, fc, fc_slab, coarse.dist, 0, no_uc, no_slab);   // residual + restriction only
                record(-1, fine.N, done, err_host);
            }
            if (!coarse.dist) {           // gather the slabs' coarse rows into rank 0's full grid
                std::vector<Xfer> xs;
                for (int k = 1; k < comm.world; ++k) {
                    const double *src = nullptr;
                    double *dst = nullptr;
                    for (size_t i = 0; i < cy.ranks.size(); ++i) {
                        if (cy.ranks[i].rank == k) src = fc_tmp[i];
                        if (cy.ranks[i].rank == 0) dst = cy.ranks[i].lv[li + 1].F + (size_t)cb[k] * next_N;
                    }
                    if (!comm.is_local(k) && !comm.is_local(0)) continue;
                    xs.push_back({k, 0, src, dst, (size_t)(cb[k + 1] - cb[k]) * next_N});
                }
                cy.transfer_now(xs);
            }
            for (double *p : fc_tmp) pool_free(p);
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

            // ---- U_c rows with halo on every rank (halos of distributed arrays are current by construction)
            std::vector<double *> uc_tmp(cy.ranks.size(), nullptr);
            std::vector<const double *> uc(cy.ranks.size(), nullptr);
            std::vector<Slab> uc_slab(cy.ranks.size());
            for (size_t i = 0; i < cy.ranks.size(); ++i) { uc[i] = cy.ranks[i].lv[lc].U; uc_slab[i] = cy.ranks[i].lv[lc].slab; }
            if (!coarse.dist) {          // scatter rank 0's full coarse grid: each rank gets its rows plus halo
                const std::vector<int> cb = coarse_bounds(fine, coarse.N, comm.world);
                std::vector<Xfer> xs;
                const double *full = nullptr;
                for (auto &st : cy.ranks) if (st.rank == 0) full = st.lv[lc].U;
                for (int k = 1; k < comm.world; ++k) {
                    Slab s;
                    s.own_lo = cb[k]; s.own_hi = cb[k + 1];
                    s.row0 = std::max(0, cb[k] - HALO);
                    s.rows = std::min(coarse.N, cb[k + 1] + HALO) - s.row0;
                    double *dst = nullptr;
                    for (size_t i = 0; i < cy.ranks.size(); ++i) {
                        if (cy.ranks[i].rank != k) continue;
                        uc_slab[i] = s;
                        uc[i] = dst = uc_tmp[i] = (double *)pool_alloc((size_t)s.rows * coarse.N * sizeof(double));
                    }
                    if (!comm.is_local(k) && !comm.is_local(0)) continue;
                    xs.push_back({0, k, full ? full + (size_t)s.row0 * coarse.N : nullptr, dst, (size_t)s.rows * coarse.N});
                }
                cy.transfer_now(xs);
            }

            const int fixed = step > 0 ? step : 0;
            const int n_pass = std::max(1, (fixed + 2) / 3), idx = next_scal();
            for (int k = 0; k < n_pass; ++k) {
                const int S = fixed / n_pass + (k < fixed % n_pass ? 1 : 0);
                const bool first = k == 0, last = k + 1 == n_pass;
                if (first) cy.pass(lf, L, S, 2, last && step > 0, idx, 0, no_fc, no_slab, false, coarse.N, uc, uc_slab);
                else       cy.pass(lf, L, S, 0, last && step > 0, idx, 0, no_fc, no_slab, false, 0, no_uc, no_slab);
            }
            if (step > 0) {
                deferred.push_back({record(1, fine.N, step, 0.0), idx});
            } else if (step < 0) {
                double slope = TRIGGER + 1.0, prev = 0.0, err_host = 0.0;
                int done = 0;
                while (slope > TRIGGER) {
                    const int j = next_scal();
                    cy.pass(lf, L, 1, 0, true, j, 0, no_fc, no_slab, false, 0, no_uc, no_slab);
                    const double s = reduced_now(j);
                    err_host = (s + s) / fine.N / fine.N;
                    ++done;
                    if (done > 1) slope = std::fabs(err_host - prev);
                    prev = err_host;
                    if (c.err_code) break;
                }
                record(1, fine.N, done, err_host);
            } else record(1, fine.N, 0, 0.0);
            for (double *p : uc_tmp) pool_free(p);
            if (!quiet) fputs(kProlongArt, stdout);
            cy.pop();
            mark(coarse.dist ? "up" : "scatter+up", fine.N);
        } else { rc = 6; break; }
    }
    cy.ts.to_compute();                          // the timed span ends when both streams have drained
    cudaEventRecord(ev1, c.stream);
    harvest();
    const auto wall1 = std::chrono::steady_clock::now();
    for (size_t i = 1; i < marks.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, marks[i - 1].ev, marks[i].ev);
        if (comm.is_local(0))
            fprintf(stderr, "[mg trace] %-12s N=%-6d device %8.3f ms   host enqueue %8.3f ms\n", marks[i].what, marks[i].N, ms,
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
                if (!l.present) continue;
                launch_source(g.N, L, l.W, min_x, min_y, true, l.slab.row0, l.slab.rows);
                const size_t off = (size_t)(l.slab.own_lo - l.slab.row0) * g.N, cnt = (size_t)(l.slab.own_hi - l.slab.own_lo) * g.N;
                launch_mean_abs_diff(cnt, l.W + off, l.U + off, 1.0, st.scal + 60);
            }
            const double s = reduced_now(60);   // ranks without a share contribute 0
            res->mg_error = s / ((double)g.N * (double)g.N);
        }
        // ---- solution out
        for (auto &st : cy.ranks) {
            RankLevel &l = st.lv[0];
            if (!l.present) continue;
            const size_t off = (size_t)(l.slab.own_lo - l.slab.row0) * g.N, cnt = (size_t)(l.slab.own_hi - l.slab.own_lo) * g.N;
            if (U_host_full) check(cudaMemcpyAsync(U_host_full + (size_t)l.slab.own_lo * g.N, l.U + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H U");
            if (U_host_own) check(cudaMemcpyAsync(U_host_own, l.U + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream), "D2H U");
            if (own_lo_out) *own_lo_out = l.slab.own_lo;
            if (own_hi_out) *own_hi_out = l.slab.own_hi;
        }
        check(cudaStreamSynchronize(c.stream), "sync");
        if (!quiet && res) {
            printf("\n\n===== Final Result =====\n    Error = %lf\nTime Used = %lf (ms)\n", res->mg_error, res->wall_ms);
        }
    }
    return rc;
}

}  // namespace
}  // namespace mg

using namespace mg;

extern "C" {

int mgDistEmuRunCycleFile(const char *path, int world, int threshold, int flags, double *U_host, mgTraceRec *recs,
                          int max_recs, mgCycleResult *res)
{
    if (!ensure_ready()) return 10;
    if (world < 1) return 11;
    EmuComm comm(world);
    return run_dist(comm, path, threshold, flags, recs, max_recs, res, U_host, nullptr, nullptr, nullptr);
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
    std::unique_ptr<NcclComm> cm(new NcclComm());
    cm->world = world;
    cm->rank = rank;
    cm->local = {rank};
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    if (!cm->ok(g_nccl.CommInitRank(&cm->comm, world, id, rank), "ncclCommInitRank")) return 2;
    g_nccl_comm = std::move(cm);
    if (world > 1 && staged_requested()) g_nccl_comm->setup_staged();   // collective; falls back to NCCL transfers if any rank fails
    return 0;
}

int mgDistSourceSlab(int N, int threshold, int *row0, int *rows, int *own_lo, int *own_hi)
{
    if (!g_nccl_comm) { fail(-34, "mgDistSourceSlab: call mgDistInit first"); return 12; }
    const LevelGeom g = top_geometry(N, g_nccl_comm->world, threshold);
    const Slab s = slab_of(g, g_nccl_comm->rank);
    const bool present = g.dist || g_nccl_comm->rank == 0;
    *row0 = s.row0; *rows = present ? s.rows : 0; *own_lo = s.own_lo; *own_hi = present ? s.own_hi : s.own_lo;
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
        if (g_src.F) pool_free(g_src.F);
        g_src = SourceCache();
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

// doSmoothing on row slabs, repeated: `reps` times `step` Jacobi sweeps (passes of <= 3 fused sweeps,
// one halo exchange per pass, error all-reduced once per repetition) on the analytic source grid,
// starting from U = 0.  BASELINE config 5's smoothing-only stress; N up to 65536 (64-bit indexing).
// Works without mgDistInit on one GPU.  U_own_host (optional) receives the rank's owned rows.
int mgDistSmoothStress(int N, double L, int step, int reps, double *ms_per_rep, double *error_out, double *U_own_host, int *own_lo,
                       int *own_hi)
{
    if (!ensure_ready()) return 10;
    if (N % 2 || N < 64 || step < 1 || reps < 1) return 1;
    EmuComm solo(1);
    Comm &comm = g_nccl_comm ? static_cast<Comm &>(*g_nccl_comm) : static_cast<Comm &>(solo);
    const int rank = g_nccl_comm ? g_nccl_comm->rank : 0;
    Context &c = ctx();
    const LevelGeom g = top_geometry(N, comm.world, 0);
    if (comm.world > 1 && !g.dist) return 2;
    const Slab sl = slab_of(g, rank);
    const size_t bytes = (size_t)sl.rows * N * sizeof(double);
    double *U = (double *)pool_alloc(bytes), *W = (double *)pool_alloc(bytes), *F = (double *)pool_alloc(bytes);
    double *scal = (double *)pool_alloc(64 * sizeof(double));
    if (!U || !W || !F || !scal) return 3;
    launch_source(N, L, F, 0.0, 0.0, false, sl.row0, sl.rows);
    check(cudaMemsetAsync(U, 0, bytes, c.stream), "memset");
    check(cudaMemsetAsync(W, 0, bytes, c.stream), "memset");

    TwoStream ts;
    const int n_pass = (step + 2) / 3;
    int rot = 0;
    auto one_rep = [&]() {
        if (++rot == 48) { ts.sync(); rot = 1; }          // scal slots still queued for an all-reduce
        for (int k = 0; k < n_pass; ++k) {
            const int S = step / n_pass + (k < step % n_pass ? 1 : 0);
            const bool last = k + 1 == n_pass;
            std::vector<Xfer> xs;
            halo_xfers(g, comm, [&](int r) -> double * { return r == rank ? W : nullptr; }, xs);
            ts.wait_halo();
            if (ts.cs != ts.ms && !xs.empty()) {
                slab_pass(N, L, S, 0, U, F, W, sl, last, scal + rot, 0, nullptr, nullptr, 0, nullptr, nullptr, 1);
                ts.edge_done();
                comm.transfer(xs, ts.cs);
                ts.halo_posted();
                slab_pass(N, L, S, 0, U, F, W, sl, last, scal + rot, 0, nullptr, nullptr, 0, nullptr, nullptr, 2);
            } else {
                slab_pass(N, L, S, 0, U, F, W, sl, last, scal + rot, 0, nullptr, nullptr, 0, nullptr, nullptr, 0);
                if (!xs.empty()) { comm.transfer(xs, ts.cs); ts.halo_posted(); }
            }
            std::swap(U, W);
        }
        if (comm.world > 1) { ts.to_comm(); comm.allreduce_sum({scal + rot}, 1, ts.cs); }
    };
    one_rep();                                   // warm-up (also NCCL connection set-up)
    ts.to_compute();                             // the warm-up's last halo exchange lands before the reset
    check(cudaMemsetAsync(U, 0, bytes, c.stream), "memset");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, c.stream);
    for (int r = 0; r < reps; ++r) one_rep();
    ts.to_compute();
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
    pool_free(U); pool_free(W); pool_free(F); pool_free(scal);
    return c.err_code ? 10 : 0;
}

void mgDistShutdown(void)
{
    if (g_nccl_comm) {
        cudaStreamSynchronize(ctx().stream);
        cudaStreamSynchronize(ctx().comm_stream);
        g_nccl_comm->staged.reset();
        g_nccl.CommDestroy(g_nccl_comm->comm);
        g_nccl_comm.reset();
    }
}

int mgDistRunCycleFile(const char *path, int threshold, int flags, double *U_own_host, int *own_lo, int *own_hi,
                       mgTraceRec *recs, int max_recs, mgCycleResult *res)
{
    if (!ensure_ready()) return 10;
    if (!g_nccl_comm) { fail(-34, "mgDistRunCycleFile: call mgDistInit first"); return 12; }
    return run_dist(*g_nccl_comm, path, threshold, flags, recs, max_recs, res, nullptr, U_own_host, own_lo, own_hi);
}

}  // extern "C"
