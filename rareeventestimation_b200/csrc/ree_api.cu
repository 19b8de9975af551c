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
