// Library plumbing of libree_b200.so: error state, device properties, ABI version.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
static int g_ctas[64] = {0};
static long long g_launches = 0;

int ree_set_error(int code, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s", what ? what : "unknown error");
    return code;
}

int ree_check_launch(const char* what) {
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(err));
    return (int)err;
}

int ree_grid_ctas() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 296;
    if (g_ctas[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        g_ctas[dev] = 2 * sms;  // two resident CTAs per SM: 296 on a B200
    }
    return g_ctas[dev];
}

extern "C" int ree_abi_version(void) { return REE_ABI_VERSION; }
extern "C" const char* ree_last_error(void) { return g_err; }
extern "C" int ree_num_ctas(void) { return ree_grid_ctas(); }
extern "C" int64_t ree_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------------------------------
// Low-latency readback of a handful of doubles (the 4-tuples the scalar optimisers consume ~30 times per iteration).
// cudaMemcpy D2H + the tensor plumbing around it costs 20-30 us per evaluation -- more than the reduction kernel itself
// (80-160 MB at HBM speed).  Here a one-warp kernel, stream-ordered behind the reduction, copies the values into mapped
// pinned host memory and then publishes a sequence number; the host spins on the sequence number.
// ---------------------------------------------------------------------------------------
#define REE_FETCH_MAX 64

struct FetchSlot {
    volatile unsigned long long seq;
    unsigned long long pad[7];
    volatile double val[REE_FETCH_MAX];
};
static FetchSlot* g_fetch_host[64] = {nullptr};
static FetchSlot* g_fetch_dev[64] = {nullptr};
static unsigned long long g_fetch_seq[64] = {0};

__global__ void k_publish(const double* __restrict__ src, int n, FetchSlot* slot, unsigned long long seq) {
    if (threadIdx.x < n) slot->val[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        slot->seq = seq;
        __threadfence_system();
    }
}

extern "C" int ree_fetch_small(const double* dev_src, int n, double* host_out, void* stream) {
    if (!dev_src || !host_out || n < 1 || n > REE_FETCH_MAX) return ree_set_error(REE_E_BADARG, "ree_fetch_small: bad argument");
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess || dev < 0 || dev >= 64) return ree_set_error(REE_E_BADARG, "ree_fetch_small: no device");
    if (!g_fetch_host[dev]) {
        void* h = nullptr;
        err = cudaHostAlloc(&h, sizeof(FetchSlot), cudaHostAllocMapped);
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
        memset(h, 0, sizeof(FetchSlot));
        void* d = nullptr;
        err = cudaHostGetDevicePointer(&d, h, 0);
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
        g_fetch_host[dev] = (FetchSlot*)h;
        g_fetch_dev[dev] = (FetchSlot*)d;
    }
    cudaStream_t st = ree_stream(stream);
    const unsigned long long seq = ++g_fetch_seq[dev];
    k_publish<<<1, REE_FETCH_MAX, 0, st>>>(dev_src, n, g_fetch_dev[dev], seq);
    int rc = ree_check_launch("ree_fetch_small");
    if (rc) return rc;
    FetchSlot* slot = g_fetch_host[dev];
    unsigned long long spins = 0;
    while (slot->seq != seq) {
        if ((++spins & 0xfff) == 0) {  // every 4096 polls: has the stream died or drained without publishing?
            err = cudaStreamQuery(st);
            if (err != cudaSuccess && err != cudaErrorNotReady) return ree_set_error((int)err, cudaGetErrorString(err));
            if (err == cudaSuccess) {  // the publishing kernel has finished: its write is on its way
                for (int k = 0; k < 1000000 && slot->seq != seq; ++k) {
                }
                if (slot->seq != seq) return ree_set_error(REE_E_BADARG, "ree_fetch_small: stream drained without publishing");
            }
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    for (int i = 0; i < n; ++i) host_out[i] = slot->val[i];
    return 0;
}

// ---------------------------------------------------------------------------------------
// Device -> pageable host copy of an N-vector (Solution.lsf_eval_hist rows, cache.lsf_evals): cudaMemcpy into pageable
// memory runs at ~2 GB/s for a fresh 80 MB array (staging inside the driver + first-touch page faults on the critical
// path).  Here two pinned 8 MB buffers alternate: the DMA engine fills one while the host copies the other out.
// ---------------------------------------------------------------------------------------
#define REE_DL_CHUNK (1 << 20)  // doubles per pinned buffer (8 MB)
static double* g_dl_pinned[64][2] = {{nullptr}};
static cudaEvent_t g_dl_event[64][2];

extern "C" int ree_download_vector(const double* dev_src, double* host_dst, int64_t n, void* stream) {
    if (!dev_src || !host_dst || n < 0) return ree_set_error(REE_E_BADARG, "ree_download_vector: bad argument");
    if (n == 0) return 0;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess || dev < 0 || dev >= 64) return ree_set_error(REE_E_BADARG, "ree_download_vector: no device");
    if (!g_dl_pinned[dev][0]) {
        for (int b = 0; b < 2; ++b) {
            err = cudaHostAlloc((void**)&g_dl_pinned[dev][b], (size_t)REE_DL_CHUNK * sizeof(double), cudaHostAllocDefault);
            if (err == cudaSuccess) err = cudaEventCreateWithFlags(&g_dl_event[dev][b], cudaEventDisableTiming);
            if (err != cudaSuccess) {
                g_dl_pinned[dev][0] = nullptr;
                return ree_set_error((int)err, cudaGetErrorString(err));
            }
        }
    }
    cudaStream_t st = ree_stream(stream);
    const int64_t nchunks = (n + REE_DL_CHUNK - 1) / REE_DL_CHUNK;
    auto issue = [&](int64_t c) -> cudaError_t {
        const int64_t i0 = c * REE_DL_CHUNK, m = (n - i0 < REE_DL_CHUNK) ? n - i0 : REE_DL_CHUNK;
        cudaError_t e = cudaMemcpyAsync(g_dl_pinned[dev][c & 1], dev_src + i0, (size_t)m * sizeof(double),
                                        cudaMemcpyDeviceToHost, st);
        return e == cudaSuccess ? cudaEventRecord(g_dl_event[dev][c & 1], st) : e;
    };
    err = issue(0);
    for (int64_t c = 0; c < nchunks && err == cudaSuccess; ++c) {
        if (c + 1 < nchunks) err = issue(c + 1);  // the other buffer was drained in the previous round
        if (err == cudaSuccess) err = cudaEventSynchronize(g_dl_event[dev][c & 1]);
        if (err != cudaSuccess) break;
        const int64_t i0 = c * REE_DL_CHUNK, m = (n - i0 < REE_DL_CHUNK) ? n - i0 : REE_DL_CHUNK;
        memcpy(host_dst + i0, g_dl_pinned[dev][c & 1], (size_t)m * sizeof(double));
    }
    return err == cudaSuccess ? 0 : ree_set_error((int)err, cudaGetErrorString(err));
}
