// NVLink peer mailboxes shared by the exchange kernels (ree_peer.cu) and the persistent sigma / beta search (ree_opt.cu).
#pragma once
#include "common.cuh"

#define REE_PEER_MAX 16
#define REE_PEER_SPIN_LIMIT (1ull << 26)  /* polls of a remote-written flag in L2: a few seconds */

struct PeerSlot {  // 64 bytes
    double v[4];
    unsigned long long seq;
    unsigned long long pad[3];
};
struct PeerBox {
    PeerSlot slot[2][REE_PEER_MAX];
    unsigned long long seq;     // exchanges completed by the owner of this mailbox (only the owner reads / writes it)
    unsigned long long ar_seq;  // all-reduces completed by the owner (parity of the staging area below)
    unsigned long long pad[6];
};

// Behind every mailbox (same allocation, same IPC handle): two staging areas of REE_PEER_STAGE doubles for the packed
// d x d sums of ree_peer_allreduce_sum (2 (d^2 + d + 2) doubles, d <= 256).
#define REE_PEER_STAGE (2 * (256 * 256 + 256 + 2) + 4)
#define REE_PEER_BYTES (sizeof(PeerBox) + 2 * (size_t)REE_PEER_STAGE * sizeof(double))
__host__ __device__ __forceinline__ double* peer_stage(PeerBox* box, int parity) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(box) + sizeof(PeerBox)) + (size_t)parity * REE_PEER_STAGE;
}

// What a kernel needs to talk to the peers.  `box[r]` is rank r's mailbox as mapped into this process (CUDA IPC, or plain
// peer access when one process drives several devices); box[rank] is this rank's own.
struct PeerArgs {
    PeerBox* box[REE_PEER_MAX];
    int rank, world;
    int* error;  // device flag: a spin timed out
};

// Exchange one 4-tuple with every rank and merge in rank order (identical bits everywhere).  Called by one full warp;
// `t` (shared memory, REE_PEER_MAX x 4 doubles) is scratch; lane 0 returns the merge.  `seq` must be the same number on
// every rank for the same exchange and grow by one per exchange (two slot parities make reuse safe, see ree_peer.cu).
__device__ __forceinline__ ExpSums peer_exchange_warp(const PeerArgs& a, unsigned long long seq, const ExpSums& mine,
                                                      double (*t)[4]) {
    const int r = threadIdx.x & 31;
    const int par = (int)(seq & 1ull);
    const double m0 = __shfl_sync(0xffffffffu, mine.m, 0), m1 = __shfl_sync(0xffffffffu, mine.s1, 0);
    const double m2 = __shfl_sync(0xffffffffu, mine.s2, 0), m3 = __shfl_sync(0xffffffffu, mine.cnt, 0);
    if (r < a.world) {
        volatile PeerSlot* dst = &a.box[r]->slot[par][a.rank];
        dst->v[0] = m0; dst->v[1] = m1; dst->v[2] = m2; dst->v[3] = m3;
        __threadfence_system();
        dst->seq = seq;
        __threadfence_system();
        volatile PeerSlot* src = &a.box[a.rank]->slot[par][r];
        unsigned long long spins = 0;
        while (src->seq != seq) {
            if (++spins > REE_PEER_SPIN_LIMIT) {  // a peer died or the call order diverged
                *a.error = 1;
                break;
            }
        }
        __threadfence_system();
#pragma unroll
        for (int k = 0; k < 4; ++k) t[r][k] = src->v[k];
    }
    __syncwarp();
    ExpSums acc{t[0][0], t[0][1], t[0][2], t[0][3]};
    if (r == 0)
        for (int q = 1; q < a.world; ++q) acc = expsums_merge(acc, ExpSums{t[q][0], t[q][1], t[q][2], t[q][3]});
    __syncwarp();
    return acc;
}

// host side (ree_peer.cu): arguments of the connected peer group of the current device, or nullptr
const PeerArgs* ree_peer_args_current();
