// Internal interface of the "planes" pipeline (NHWC bf16 hi/lo activation planes, TMA-fed tcgen05 conv kernels).
#pragma once
#include "clb_tc_ptx.cuh"

namespace clb {
namespace pl {

// clb_planes_conv.cu
bool conv_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
bool linear_supported(int in, int out);
// taps = 9: 3x3 / pad 1 convolution; taps = 1: nn.Linear (H = W = 1, N = rows)
int conv_fwd(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, int relu,
             const uint16_t* mask_hi, uint16_t* y_hi, uint16_t* y_lo, int N, int H, int W, int Cred, int Cout, int taps, cudaStream_t s);
void wgrad_plan(int N, int H, int W, int Cred, int Cout, int taps, int* splits, int* kb_per_split, int* n_kb);
size_t wgrad_ws_floats(int N, int H, int W, int Cred, int Cout, int taps);
int conv_wgrad_partials(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dy_hi, const uint16_t* dy_lo, float* ws,
                        int* splits_out, int N, int H, int W, int Cred, int Cout, int taps, cudaStream_t s);

// bf16 helpers shared by the layer kernels
__device__ __forceinline__ uint32_t bf16_bits_rn(float x) {          // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u = __float_as_uint(x);
    if ((u & 0x7F800000u) == 0x7F800000u) return u >> 16;             // inf / nan: truncate
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u >> 16;
}
__device__ __forceinline__ void split1(float v, uint32_t& hi, uint32_t& lo) {
    hi = bf16_bits_rn(v);
    lo = bf16_bits_rn(v - __uint_as_float(hi << 16));
}
__device__ __forceinline__ float join1(uint32_t hi, uint32_t lo) { return __uint_as_float(hi << 16) + __uint_as_float(lo << 16); }

}  // namespace pl
}  // namespace clb
