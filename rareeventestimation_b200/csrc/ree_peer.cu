// Exchange of the per-rank 4-tuples [M, S1, S2, cnt] over NVLink peer memory (one box, one process per GPU).
//
// Every objective evaluation of the sigma / beta optimisers ends in a cross-rank merge of 32 bytes per rank
// (DESIGN.md section 6).  Through NCCL that is an all_gather launch + a merge kernel: ~40 us of latency for a
// reduction that itself takes ~20-40 us, ~30 times per iteration.  Here each rank owns a small mailbox in its own HBM,
// exported with CUDA IPC and mapped by every peer.  One kernel
//   1. stores this rank's tuple and then a sequence number into slot [seq & 1][rank] of EVERY rank's mailbox (plain
//      stores through the peer mapping: NVLink writes land in the owner's L2),
//   2. spins on its OWN mailbox until all slots of parity seq & 1 carry `seq`,
//   3. merges the tuples in rank order (same arithmetic as ree_merge_expsums => identical bits on every rank).
// Two slot parities: rank A can only overwrite a slot of parity p after it finished exchange seq + 1, which needs
// everybody's seq + 1 tuple, which a rank sends after it has read all seq tuples.  Every rank calls the exchange in the
// same order (the host control flow is identical on all ranks), so `seq` is a plain per-process counter.
#include <string.h>

#include "peer.cuh"

struct PeerState {
    int rank, world;
    PeerBox* local;                 // this rank's mailbox (cudaMalloc)
    PeerBox* box[REE_PEER_MAX];     // every rank's mailbox as mapped into this process (box[rank] == local)
    int* d_error;                   // device flag: a spin timed out
    unsigned char handle[REE_PEER_MAX][64];  // IPC handles already mapped (a second process group reuses the mappings)
    PeerArgs args;
};
static PeerState g_peer[64];
static bool g_peer_ready[64] = {false};

const PeerArgs* ree_peer_args_current() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_peer_ready[dev]) return nullptr;
    return &g_peer[dev].args;
}

// The exchange counter lives in the owner's mailbox (device memory): the persistent sigma / beta search performs a
// data-dependent number of exchanges inside one kernel, so the host cannot number them.
__global__ void __launch_bounds__(32) k_peer_exchange(double* out4, const PeerArgs a) {
    __shared__ double t[REE_PEER_MAX][4];
    PeerBox* own = a.box[a.rank];
    const unsigned long long seq = own->seq + 1;
    const ExpSums mine{out4[0], out4[1], out4[2], out4[3]};
    const ExpSums acc = peer_exchange_warp(a, seq, mine, t);
    if (threadIdx.x == 0) {
        out4[0] = acc.m; out4[1] = acc.s1; out4[2] = acc.s2; out4[3] = acc.cnt;
        own->seq = seq;
    }
}

// Sum of a small fp64 vector over all ranks, in place, as ONE kernel over peer memory (the packed d x d sums after the
// Gram pass: 162 KB at d = 100) -- instead of an NCCL all_reduce launch next to the kernels that produce and consume it.
//   1. copy the local vector into this rank's staging area [parity] (behind its mailbox), system fence;
//   2. one mailbox exchange as the barrier: a rank publishes its sequence number only after its staging area is complete;
//   3. every rank adds the staging areas of all ranks in rank order (NVLink reads, L1 bypassed) -> identical bits.
// Two parities: a rank overwrites parity p again two all-reduces later, i.e. after a barrier that every rank only
// enters once it has finished reading p.
__global__ void __launch_bounds__(1024) k_peer_allreduce(double* buf, int n, const PeerArgs a) {
    __shared__ double t[REE_PEER_MAX][4];
    PeerBox* own = a.box[a.rank];
    const int par = (int)(own->ar_seq & 1ull);
    double* mine = peer_stage(own, par);
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = buf[i];
    __syncthreads();  // (the publishing lanes fence at system scope before they store the sequence number: cumulative)
    if (threadIdx.x < 32) {
        const unsigned long long seq = own->seq + 1;
        (void)peer_exchange_warp(a, seq, ExpSums{0.0, 0.0, 0.0, 0.0}, t);
        if (threadIdx.x == 0) {
            own->seq = seq;
            own->ar_seq += 1;
        }
    }
    __syncthreads();
    // batches of AR_B elements per thread: AR_B x world independent NVLink reads in flight, then the rank-order sums
    constexpr int AR_B = 8;
    const double* stage[REE_PEER_MAX];
    for (int r = 0; r < a.world; ++r) stage[r] = peer_stage(a.box[r], par);
    for (int i0 = threadIdx.x; i0 < n; i0 += AR_B * blockDim.x) {
        double acc[AR_B];
#pragma unroll
        for (int j = 0; j < AR_B; ++j) acc[j] = 0.0;
        for (int r = 0; r < a.world; ++r) {
            double v[AR_B];
#pragma unroll
            for (int j = 0; j < AR_B; ++j) {
                const int i = i0 + j * (int)blockDim.x;
                v[j] = 0.0;
                if (i < n) asm("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(stage[r] + i));
            }
#pragma unroll
            for (int j = 0; j < AR_B; ++j) acc[j] += v[j];
        }
#pragma unroll
        for (int j = 0; j < AR_B; ++j) {
            const int i = i0 + j * (int)blockDim.x;
            if (i < n) buf[i] = acc[j];
        }
    }
}

extern "C" int ree_peer_allreduce_sum(double* buf, int64_t n, void* stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_peer_ready[dev] || !buf)
        return ree_set_error(REE_E_BADARG, "ree_peer_allreduce_sum: peers are not connected");
    if (n < 1 || n > REE_PEER_STAGE) return ree_set_error(REE_E_UNSUPPORTED, "ree_peer_allreduce_sum: vector too long");
    k_peer_allreduce<<<1, 1024, 0, ree_stream(stream)>>>(buf, (int)n, g_peer[dev].args);
    return ree_check_launch("ree_peer_allreduce_sum");
}

extern "C" int64_t ree_peer_allreduce_max(void) { return REE_PEER_STAGE; }

// ---- setup -------------------------------------------------------------------------------------
// 1. every rank: ree_peer_create(rank, world, handle_out[64])      -> allocates the mailbox, returns its IPC handle
// 2. exchange the handles (any transport; engine.py uses torch.distributed.all_gather_object)
// 3. every rank: ree_peer_connect(handles[world][64])               -> maps the peers' mailboxes
extern "C" int ree_peer_create(int rank, int world, unsigned char* handle_out) {
    if (rank < 0 || world < 2 || world > REE_PEER_MAX || rank >= world || !handle_out)
        return ree_set_error(REE_E_BADARG, "ree_peer_create: bad argument");
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess || dev < 0 || dev >= 64) return ree_set_error(REE_E_BADARG, "ree_peer_create: no device");
    PeerState& s = g_peer[dev];
    if (!s.local) {
        err = cudaMalloc((void**)&s.local, REE_PEER_BYTES);
        if (err == cudaSuccess) err = cudaMemset(s.local, 0, REE_PEER_BYTES);
        if (err == cudaSuccess) err = cudaMalloc((void**)&s.d_error, sizeof(int));
        if (err == cudaSuccess) err = cudaMemset(s.d_error, 0, sizeof(int));
        if (err == cudaSuccess) err = cudaDeviceSynchronize();  // the mailbox is zero before its handle leaves the process
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
    } else {
        // a new process group in the same process: empty the mailbox again (the handle exchange that follows is the
        // barrier before any peer writes into it) and forget the old error
        err = cudaMemset(s.local, 0, REE_PEER_BYTES);
        if (err == cudaSuccess) err = cudaMemset(s.d_error, 0, sizeof(int));
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
    }
    g_peer_ready[dev] = false;
    s.rank = rank; s.world = world;
    cudaIpcMemHandle_t h;
    err = cudaIpcGetMemHandle(&h, s.local);
    if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    return 0;
}

extern "C" int ree_peer_connect(const unsigned char* handles) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess || dev < 0 || dev >= 64 || !g_peer[dev].local || !handles)
        return ree_set_error(REE_E_BADARG, "ree_peer_connect: ree_peer_create first");
    PeerState& s = g_peer[dev];
    for (int r = 0; r < s.world; ++r) {
        if (r == s.rank) {
            s.box[r] = s.local;
            continue;
        }
        if (s.box[r] && memcmp(s.handle[r], handles + 64 * (size_t)r, 64) == 0) continue;  // already mapped
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
        s.box[r] = (PeerBox*)p;
        memcpy(s.handle[r], handles + 64 * (size_t)r, 64);
    }
    memset(&s.args, 0, sizeof(s.args));
    for (int r = 0; r < s.world; ++r) s.args.box[r] = s.box[r];
    s.args.rank = s.rank; s.args.world = s.world; s.args.error = s.d_error;
    g_peer_ready[dev] = true;
    return 0;
}

// One process driving several devices (one host thread per device): the peers' mailboxes are plain device pointers of
// this process, reached through peer access -- no IPC.  Every device's thread: ree_peer_create (the handle is not
// used), a host barrier, ree_peer_connect_local(devices[world]), a host barrier.
extern "C" int ree_peer_connect_local(const int* devices) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess || dev < 0 || dev >= 64 || !g_peer[dev].local || !devices)
        return ree_set_error(REE_E_BADARG, "ree_peer_connect_local: ree_peer_create first");
    PeerState& s = g_peer[dev];
    for (int r = 0; r < s.world; ++r) {
        const int od = devices[r];
        if (od < 0 || od >= 64 || !g_peer[od].local)
            return ree_set_error(REE_E_BADARG, "ree_peer_connect_local: a device of the group has no mailbox yet");
        if (od != dev) {
            int can = 0;
            err = cudaDeviceCanAccessPeer(&can, dev, od);
            if (err != cudaSuccess || !can) return ree_set_error(REE_E_UNSUPPORTED, "ree_peer_connect_local: no peer access");
            err = cudaDeviceEnablePeerAccess(od, 0);
            if (err == cudaErrorPeerAccessAlreadyEnabled) {
                (void)cudaGetLastError();
            } else if (err != cudaSuccess) {
                return ree_set_error((int)err, cudaGetErrorString(err));
            }
        }
        s.box[r] = g_peer[od].local;
        memset(s.handle[r], 0, 64);
    }
    memset(&s.args, 0, sizeof(s.args));
    for (int r = 0; r < s.world; ++r) s.args.box[r] = s.box[r];
    s.args.rank = s.rank; s.args.world = s.world; s.args.error = s.d_error;
    g_peer_ready[dev] = true;
    return 0;
}

extern "C" int ree_peer_ready(void) {
    int dev = 0;
    return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && g_peer_ready[dev]) ? 1 : 0;
}

// In place: out4 (DEVICE, this rank's tuple) becomes the merge over all ranks.  Stream ordered; all ranks must call it
// in the same order.  A timed-out wait is reported by the NEXT call (the flag is read back lazily) or by ree_peer_check.
extern "C" int ree_peer_combine_expsums(double* out4, void* stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_peer_ready[dev] || !out4)
        return ree_set_error(REE_E_BADARG, "ree_peer_combine_expsums: peers are not connected");
    k_peer_exchange<<<1, 32, 0, ree_stream(stream)>>>(out4, g_peer[dev].args);
    return ree_check_launch("ree_peer_combine_expsums");
}

extern "C" int ree_peer_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_peer_ready[dev]) return 0;
    int flag = 0;
    cudaError_t err = cudaMemcpy(&flag, g_peer[dev].d_error, sizeof(int), cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
    return flag ? ree_set_error(REE_E_BADARG, "peer exchange timed out (a rank died or the call order diverged)") : 0;
}
