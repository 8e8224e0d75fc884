// Error string, version, device queries.
#include <stdarg.h>
#include <stdlib.h>

#include "clb_common.cuh"

namespace clb {
static thread_local char g_err[512] = "";
// default: tensor-core parity mode (bf16 hi/lo split for the conv kernels, 3-pass TF32 split elsewhere); CLB_MM_MODE=0/1/2/3
// overrides at first use
static int g_mm_mode = -1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;  // B200
    }
    return cached;
}
int mm_mode() {
    if (g_mm_mode < 0) {
        const char* e = getenv("CLB_MM_MODE");
        g_mm_mode = (e && e[0] >= '0' && e[0] <= '3' && e[1] == 0) ? (e[0] - '0') : CLB_MM_BF16X3;
    }
    return g_mm_mode;
}
int pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("CLB_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }
unsigned long long launches() { return g_launches; }
}  // namespace clb

extern "C" {
const char* clb_last_error(void) { return clb::g_err; }
int clb_version(void) { return 100; }
int clb_sm_count(int* out) {
    CLB_CHECK_ARG(out != nullptr);
    int dev = 0, n = 0;
    CLB_CUDA(cudaGetDevice(&dev));
    CLB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    *out = n;
    return CLB_OK;
}
int clb_set_matmul_mode(int mode) {
    CLB_CHECK_ARG(mode == CLB_MM_FP32_SIMT || mode == CLB_MM_TF32X3 || mode == CLB_MM_TF32X1 || mode == CLB_MM_BF16X3);
    clb::g_mm_mode = mode;
    return CLB_OK;
}
int clb_get_matmul_mode(void) { return clb::mm_mode(); }
unsigned long long clb_launch_count(void) { return clb::launches(); }
int clb_stream_create(void** out) {
    CLB_CHECK_ARG(out != nullptr);
    cudaStream_t st = nullptr;
    CLB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *out = (void*)st;
    return CLB_OK;
}
int clb_memset_zero(void* p, size_t bytes, void* stream) {
    CLB_CHECK_ARG(p != nullptr || bytes == 0);
    if (bytes) CLB_CUDA(cudaMemsetAsync(p, 0, bytes, clb::as_stream(stream)));
    return CLB_OK;
}
}
