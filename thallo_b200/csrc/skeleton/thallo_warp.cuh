// thallo_b200 solver skeleton: warp-level primitives of the scatter path.
//
// Reference roles replaced (API/src/cuda_util.t): `ballot` / `laneid` / `shfl` wrappers :164-285
// (pre-Volta non-_sync PTX there), `get_peers` :329-350, `reduce_peersf/d` :352-418, and the
// by-key reduction the residualwise operators use before their float atomics
// (`warp_aggregated_atomic_reduction_by_key`, thallo.t:3361-3402).  The reference's own
// known-answer tests for these (tests/cuda_unit_tests/{ballot,get_peers,reduce_peers}.t) are
// re-run against this file through ThalloB200_WarpSelfTest (kernels at the bottom).
//
// sm_100a forms: peer discovery is one MATCH instruction (__match_any_sync) instead of the
// reference's claim-and-ballot loop; all shuffles are the _sync variants over the mask of lanes
// that entered the call together.
#pragma once

__device__ __forceinline__ unsigned th_laneid() {
    unsigned l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

// lanes (of the lanes in `active`) whose predicate is non-zero
__device__ __forceinline__ unsigned th_ballot(unsigned active, int pred) { return __ballot_sync(active, pred); }

// lanes of `active` that hold the same key as the caller (the caller included)
__device__ __forceinline__ unsigned th_get_peers(unsigned active, int key) { return __match_any_sync(active, key); }
__device__ __forceinline__ unsigned th_get_peers(unsigned active, long long key) {
    return __match_any_sync(active, (unsigned long long)key);
}

__device__ __forceinline__ float th_shfl(unsigned active, float v, int src) { return __shfl_sync(active, v, src); }
__device__ __forceinline__ double th_shfl(unsigned active, double v, int src) { return __shfl_sync(active, v, src); }

// Sum of x over every peer set, delivered in the lowest lane of the set (other lanes hold partial
// sums).  Pairwise tree over the caller's rank among its peers: in every round the lanes at an
// even rank take the value of the next remaining peer above them, then the odd ranks retire; a
// set of n peers finishes in ceil(log2 n) rounds and the loop runs until the largest set is done.
template <class T> __device__ __forceinline__ T th_reduce_peers(unsigned active, T x, unsigned peers) {
    const unsigned lane = th_laneid();
    unsigned rank = __popc(peers & ((1u << lane) - 1u));
    unsigned above = peers & (0xfffffffeu << lane);
    while (__any_sync(active, above != 0u)) {
        const int src = __ffs(above) - 1;                       // nearest remaining peer above (-1: none)
        const T t = th_shfl(active, x, src & 31);               // every lane shuffles; only takers add
        if (src >= 0) x += t;
        const unsigned retired = th_ballot(active, (int)(rank & 1u));
        above &= ~retired;
        rank >>= 1;
    }
    return x;
}

// dest += sum of x over the peer set, one atomic per set (dest may be null: reduce only)
template <class T> __device__ __forceinline__ void th_reduce_peers_atomic(unsigned active, T* dest, T x, unsigned peers) {
    x = th_reduce_peers(active, x, peers);
    if (dest != nullptr && th_laneid() == (unsigned)(__ffs(peers) - 1)) atomicAdd(dest, x);
}

#ifdef TH_WARP_KAT
// Known-answer kernels, one warp each, written after the reference's unit tests:
//   ballot.t       every lane votes with its lane id; max over lanes of the ballot = 0xfffffffe
//   get_peers.t    key = lane % 4; sum over lanes of (peers & 0xff) = 255 * 32 / 4
//   reduce_peers.t key = lane % 4; dest[key] += lane  ->  dest[i] = 112 + 8 i
extern "C" __global__ void th_kat_ballot(unsigned* result) {
    const unsigned t = th_ballot(0xffffffffu, (int)threadIdx.x);
    atomicMax(result, t);
}
extern "C" __global__ void th_kat_get_peers(unsigned* result) {
    const unsigned t = th_get_peers(0xffffffffu, (int)(threadIdx.x % 4)) & 0xffu;
    atomicAdd(result, t);
}
extern "C" __global__ void th_kat_reduce_peers(float* resultf, double* resultd, int nkeys) {
    const int key = (int)threadIdx.x % nkeys;
    const unsigned peers = th_get_peers(0xffffffffu, key);
    th_reduce_peers_atomic(0xffffffffu, resultf + key, (float)threadIdx.x, peers);
    th_reduce_peers_atomic(0xffffffffu, resultd + key, (double)threadIdx.x, peers);
}
#endif
