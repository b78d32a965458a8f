// Error reporting and device queries for the C ABI (include/morpheus_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// The reference never checks launches (gridencoder.cu has no cudaGetLastError); we surface launch
// errors synchronously and leave execution errors to the caller's next sync, like any CUDA library.
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return MB_ECUDA;
    }
    return MB_OK;
}

}  // namespace mb

extern "C" int mb_version(void) { return 100; }
extern "C" const char* mb_last_error(void) { return mb::g_err; }
extern "C" int mb_sm_count(void) {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;  // B200
    }
    cached = n;
    return n;
}
