// Shared helpers for the clb CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/clb.h"

namespace clb {

void set_error(const char* fmt, ...);
int sm_count();
int mm_mode();
// modes that keep fp32 parity through an operand split (the lo planes are needed)
inline bool mm_split() { return mm_mode() == CLB_MM_TF32X3 || mm_mode() == CLB_MM_BF16X3; }
void count_launch();      // every kernel launch of this library is counted (bench.py's gpu_launches)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define CLB_CHECK_ARG(cond)                                                        \
    do {                                                                           \
        if (!(cond)) {                                                             \
            clb::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
            return CLB_EINVAL;                                                     \
        }                                                                          \
    } while (0)

#define CLB_CHECK_LAUNCH()                                                         \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            clb::set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return CLB_ECUDA;                                                      \
        }                                                                          \
    } while (0)

#define CLB_CUDA(call)                                                             \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            clb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CLB_ECUDA;                                                      \
        }                                                                          \
    } while (0)

// Programmatic dependent launch.  Kernels of the per-step chain are launched with the programmatic-stream-serialization
// attribute: a kernel may become resident while its predecessor in the stream is still draining (CTA by CTA for the persistent
// GEMMs), runs its prologue, and blocks in pdl_wait() until the predecessor has completed and flushed its memory.  Every kernel
// launched through launch_pdl() MUST call pdl_wait() before its first global-memory access; pdl_trigger() lets the successor in.
// Captured into CUDA graphs as programmatic edges.  CLB_PDL=0 launches plainly.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// tensor-core (tcgen05) path, clb_gemm_tc.cu
bool tc_fwd_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
bool tc_dgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
bool tc_wgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
int tc_conv_fwd(const float* x, const float* w2, const float* bias, float* y, int N, int C, int H, int W, int K, int R,
                int S, int pad, int relu, bool with_lo, cudaStream_t s);
size_t tc_wgrad_ws_floats(int N, int C, int H, int W, int K, int R, int S);
int tc_conv_wgrad(const float* x, const float* dy, float* dw, float* ws, float* bias_part, bool* bias_partials_done, int N,
                  int C, int H, int W, int K, int R, int S, int pad, bool with_lo, cudaStream_t s);
// bias-gradient partial sums fused with the lo-plane split of dY (clb_gemm_simt.cu)
void conv_bias_partials_and_lo(const float* dy, float* dy_lo, float* scratch, int N, int K, int PQ, cudaStream_t s);
void conv_bias_partials_and_bf16(const float* dy, uint16_t* dy_hi, uint16_t* dy_lo, float* scratch, int N, int K, int PQ,
                                 cudaStream_t s);
bool tc4_wgrad_supported(int H, int W, int R, int S);
int tc4_conv_wgrad(const float* x, const float* dy, float* ws_partials, float* bias_part, float* dy_planes, int N, int C, int H,
                   int W, int K, int R, int S, int pad, int splits32, int kb_per_split32, int* splits_out, cudaStream_t s);
bool tc3_linear_supported(int M, int in, int out);
size_t tc3_linear_ws_floats(int M, int in, int out);
int tc3_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* ws, int M, int in, int out, int relu,
                   bool with_lo, cudaStream_t s);
int tc3_linear_dgrad(const float* dy, const float* w, float* dx, float* ws, int M, int in, int out, bool with_lo, cudaStream_t s);
int tc3_linear_wgrad(const float* x, const float* dy, float* dw, float* ws, int M, int in, int out, bool with_lo, cudaStream_t s);
void tc_permute_w_fwd(const float* w, float* w2, int K, int C, int RS, cudaStream_t s, bool bf16 = false);
void tc_permute_w_dgrad(const float* w, float* wd, int K, int C, int R, int S, cudaStream_t s, bool bf16 = false);
// bf16 hi/lo split kernels for conv fwd / dgrad (clb_gemm_tc4.cu); the bf16 planes sit where the fp32 planes would
bool tc4_fwd_supported(int C);
bool tc_bf16_route(int reduction_channels);      // mode == CLB_MM_BF16X3 and the bf16 kernel can take this layer
int tc_conv_fwd_bf16(const float* x, const float* w, float* w_ws, const float* bias, float* y, int N, int C, int H, int W, int K,
                     int R, int S, int pad, int relu, cudaStream_t s);
int tc_conv_dgrad_bf16(const float* dy, const float* w, float* wt_ws, float* dx, int N, int C, int P, int Q, int K, int R, int S,
                       int pad, cudaStream_t s);
int tc4_conv_fwd(const float* x, const void* w_hi, const void* w_lo, const float* bias, float* y, int N, int C, int H, int W,
                 int K, int R, int S, int pad, int relu, cudaStream_t s);

}  // namespace clb
