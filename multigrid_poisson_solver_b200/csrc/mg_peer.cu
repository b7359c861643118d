// mg_peer.cu -- the -1 and 1 nodes on row slabs with peer memory: instantiations of the 2-column streaming kernel
// (mg_stream.cuh) that also store the rows a neighbouring GPU keeps as its halo straight into its array.  Separate
// from the single-GPU instantiations (mg_fused.cu), which sit exactly at their register limits.
#include "mg_fused.h"
#include "mg_kernels.h"
#include "mg_stream.cuh"

namespace mg {
namespace {

template <int S, int IN, bool ERR, bool RES>
void launch_one(StreamParams &p)
{
    using G = StreamGeo<S, ERR || RES, RES>;
    constexpr int warps = stream_shape(RES).warps, ctas = stream_shape(RES).min_ctas, smem = stream_smem_bytes(IN, warps, RES);
    const int blocks = stream_launch_prepare(p, G::W, warps, ctas, ERR, 2 * S + 3);
    if (blocks == 0) return;
    static bool opted_in = false;   // one flag per instantiation
    if (!opted_in) {
        check(cudaFuncSetAttribute(k_stream<S, IN, ERR, RES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute(k_stream, peer)");
        opted_in = true;
    }
    Context &c = ctx();
    k_stream<S, IN, ERR, RES, true><<<blocks, warps * 32, smem, c.stream>>>(p);
    c.launches++;
    check(cudaGetLastError(), "k_stream (peer)");
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

template <int IN, bool RES, bool PEER>
void launch_mid(StreamParams &p)
{
    using G = StreamGeo<2, true, RES>;
    constexpr int warps = stream_shape(RES).warps, ctas = stream_shape(RES).min_ctas, smem = stream_smem_bytes(IN, warps, RES);
    const int blocks = stream_launch_prepare(p, G::W, warps, ctas, true, 2 * 2 + 3);
    if (blocks == 0) return;
    static bool opted_in = false;
    if (!opted_in) {
        check(cudaFuncSetAttribute(k_stream<2, IN, true, RES, PEER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute(k_stream, mid)");
        opted_in = true;
    }
    Context &c = ctx();
    k_stream<2, IN, true, RES, PEER, true><<<blocks, warps * 32, smem, c.stream>>>(p);
    c.launches++;
    check(cudaGetLastError(), "k_stream (mid)");
}

template <bool PEER>
void launch_mid_any(int in, bool res, StreamParams &p)
{
    if (in == IN_PROLONG) launch_mid<IN_PROLONG, false, PEER>(p);
    else if (res) { if (in == IN_ZERO) launch_mid<IN_ZERO, true, PEER>(p); else launch_mid<IN_LOAD, true, PEER>(p); }
    else { if (in == IN_ZERO) launch_mid<IN_ZERO, false, PEER>(p); else launch_mid<IN_LOAD, false, PEER>(p); }
}

}  // namespace

void launch_stream_mid(int in, bool res, StreamParams &p)
{
    if (p.peer_U_lo || p.peer_U_hi || p.peer_Fc_lo || p.peer_Fc_hi) launch_mid_any<true>(in, res, p);
    else launch_mid_any<false>(in, res, p);
}

// in: 0 load, 1 zero, 2 prolong; mode: 2 (ERR + RES) for in 0 / 1, 0 or 1 for in 2
void launch_stream_peer(int S, int in, int mode, StreamParams &p)
{
    if (in == IN_PROLONG) {
        if (mode == 0) launch_s<IN_PROLONG, false, false>(S, p);
        else launch_s<IN_PROLONG, true, false>(S, p);
    } else if (in == IN_ZERO) launch_s<IN_ZERO, true, true>(S, p);
    else launch_s<IN_LOAD, true, true>(S, p);
}

}  // namespace mg
