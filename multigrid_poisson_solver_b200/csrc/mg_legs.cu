// mg_legs.cu -- instantiations and launcher of k_strip (mg_strip.cuh), the 4-column bulk-copy kernel that runs
// every fused pass on even-sized grids: smoothing passes, the -1 node and the 1 node, whole grids and row slabs.
#include "mg_fused.h"
#include "mg_kernels.h"
#include "mg_strip.cuh"

namespace mg {
namespace {

template <int S, int IN, bool ERR, bool RES>
void launch_one(StreamParams &p)
{
    using G = StripGeo<S, ERR || RES, RES>;
    constexpr int ctas = strip_min_ctas(IN, RES), smem = strip_smem_bytes(IN, RES);
    const int blocks = stream_launch_prepare(p, G::W, SP_WARPS, ctas, ERR, 2 * S + 3);
    if (blocks == 0) return;
    static bool opted_in = false;   // one flag per instantiation
    if (!opted_in) {
        check(cudaFuncSetAttribute(k_strip<S, IN, ERR, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute(k_strip)");
        opted_in = true;
    }
    Context &c = ctx();
    k_strip<S, IN, ERR, RES><<<blocks, SP_WARPS * 32, smem, c.stream>>>(p);
    c.launches++;
    check(cudaGetLastError(), "k_strip");
}

template <int IN, bool ERR, bool RES>
void launch_s(int S, StreamParams &p)
{
    switch (S) {
        case 0: launch_one<0, IN, ERR, RES>(p); break;
        case 1: launch_one<1, IN, ERR, RES>(p); break;
        case 2: launch_one<2, IN, ERR, RES>(p); break;
        default: launch_one<3, IN, ERR, RES>(p); break;
    }
}

}  // namespace

void launch_strip(int S, int in, int mode, StreamParams &p)
{
    if (in == IN_LOAD) {
        if (mode == 0) launch_s<IN_LOAD, false, false>(S, p);
        else if (mode == 1) launch_s<IN_LOAD, true, false>(S, p);
        else launch_s<IN_LOAD, true, true>(S, p);
    } else if (in == IN_ZERO) {
        if (mode == 0) launch_s<IN_ZERO, false, false>(S, p);
        else if (mode == 1) launch_s<IN_ZERO, true, false>(S, p);
        else launch_s<IN_ZERO, true, true>(S, p);
    } else {
        if (mode == 0) launch_s<IN_PROLONG, false, false>(S, p);
        else launch_s<IN_PROLONG, true, false>(S, p);
    }
}

}  // namespace mg

#ifdef MG_SP_DEBUG
extern "C" void mgStripDebug(unsigned long long *out64, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out64, mg::g_strip_dbg, 64 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[64] = {}; cudaMemcpyToSymbol(mg::g_strip_dbg, z, sizeof z); }
}
#endif
