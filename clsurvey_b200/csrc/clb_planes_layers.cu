// HBM-bound kernels of the "planes" pipeline (NHWC bf16 hi/lo planes): weight re-ordering, 2x2 max-pooling forward /
// backward (with the layout changes at both ends of the conv stack), bias gradients and the split-K reduction.
// Replaces nn.MaxPool2d / nn.ReLU backward of src/models/VGGSlim.py:27-40 for the layers that run on the planes kernels.
#include <float.h>

#include "clb_planes.cuh"

namespace clb {
namespace pl {

static inline int ew_grid(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads, cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------------ weights -> planes
// w [K][C][3][3] fp32  ->  wf [K][tap][C] (forward B operand) and wt [C][8 - tap][K] (dgrad B operand: flipped taps,
// transposed channels), each as bf16 hi / lo planes.  blockIdx.y selects the layout so that the writes of both are coalesced.
__global__ void __launch_bounds__(256) weights_to_planes_kernel(const float* __restrict__ w, uint16_t* __restrict__ wf_hi,
                                                                uint16_t* __restrict__ wf_lo, uint16_t* __restrict__ wt_hi,
                                                                uint16_t* __restrict__ wt_lo, int K, int C, int taps) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = (int64_t)K * C, gs = (int64_t)gridDim.x * blockDim.x;
    const bool dgrad = blockIdx.y == 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        int k, c;
        if (!dgrad) { k = (int)(i / C); c = (int)(i - (int64_t)k * C); }          // c fastest
        else { c = (int)(i / K); k = (int)(i - (int64_t)c * K); }                  // k fastest
        const float* src = w + ((int64_t)k * C + c) * taps;
        for (int t = 0; t < taps; ++t) {
            uint32_t hi, lo;
            split1(__ldg(src + t), hi, lo);
            if (!dgrad) {
                const int64_t o = ((int64_t)k * taps + t) * C + c;
                wf_hi[o] = (uint16_t)hi; wf_lo[o] = (uint16_t)lo;
            } else {
                const int64_t o = ((int64_t)c * taps + (taps - 1 - t)) * K + k;
                wt_hi[o] = (uint16_t)hi; wt_lo[o] = (uint16_t)lo;
            }
        }
    }
}
int weights_to_planes(const float* w, uint16_t* wf_hi, uint16_t* wf_lo, uint16_t* wt_hi, uint16_t* wt_lo, int K, int C, int taps,
                      cudaStream_t s) {
    dim3 grid(ew_grid((int64_t)K * C, 256), wt_hi ? 2 : 1);
    launch_pdl(weights_to_planes_kernel, dim3(grid), dim3(256), 0, s, w, wf_hi, wf_lo, wt_hi, wt_lo, K, C, taps); clb::count_launch();
    return CLB_OK;
}

// all planes convs of a model in ONE launch (blockIdx.y = 2 * layer + layout): a VGG-11 step otherwise pays seven ~12 us
// latency-bound launches for 37 MB of weights
struct WeightBatch {
    static constexpr int kMax = 24;
    int n;
    const float* w[kMax];
    uint16_t *wf_hi[kMax], *wf_lo[kMax], *wt_hi[kMax], *wt_lo[kMax];
    int K[kMax], C[kMax], taps[kMax];                // taps = 9 (3x3 conv weight [K][C][3][3]) or 1 (nn.Linear weight [K][C])
};
__global__ void __launch_bounds__(256) weights_to_planes_batch_kernel(const __grid_constant__ WeightBatch b) {
    pdl_trigger();
    pdl_wait();
    // thread = two adjacent elements of the contiguous output dimension (c for the forward layout, k for the dgrad layout):
    // 4-byte stores, coalesced along that dimension; C and K are multiples of 64
    const int layer = blockIdx.y >> 1;
    const bool dgrad = blockIdx.y & 1;
    const int K = b.K[layer], C = b.C[layer], taps = b.taps[layer];
    const float* __restrict__ w = b.w[layer];
    uint32_t* __restrict__ o_hi = reinterpret_cast<uint32_t*>(dgrad ? b.wt_hi[layer] : b.wf_hi[layer]);
    uint32_t* __restrict__ o_lo = reinterpret_cast<uint32_t*>(dgrad ? b.wt_lo[layer] : b.wf_lo[layer]);
    const int64_t total = (int64_t)K * C / 2, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        int k, c;
        const float *s0, *s1;
        if (!dgrad) {
            k = (int)(i / (C / 2)); c = 2 * (int)(i - (int64_t)k * (C / 2));
            s0 = w + ((int64_t)k * C + c) * taps; s1 = s0 + taps;
        } else {
            c = (int)(i / (K / 2)); k = 2 * (int)(i - (int64_t)c * (K / 2));
            s0 = w + ((int64_t)k * C + c) * taps; s1 = s0 + (int64_t)C * taps;
        }
        for (int t = 0; t < taps; ++t) {
            uint32_t h0, l0, h1, l1;
            split1(__ldg(s0 + t), h0, l0);
            split1(__ldg(s1 + t), h1, l1);
            const int64_t o = dgrad ? (((int64_t)c * taps + (taps - 1 - t)) * K + k) >> 1 : (((int64_t)k * taps + t) * C + c) >> 1;
            o_hi[o] = h0 | (h1 << 16); o_lo[o] = l0 | (l1 << 16);
        }
    }
}
int weights_to_planes_batch(int n, const float* const* w, void* const* wf_hi, void* const* wf_lo, void* const* wt_hi, void* const* wt_lo,
                            const int* K, const int* C, const int* taps, cudaStream_t s) {
    WeightBatch b;
    b.n = n;
    int64_t mx = 1;
    for (int i = 0; i < n; ++i) {
        b.w[i] = w[i]; b.wf_hi[i] = (uint16_t*)wf_hi[i]; b.wf_lo[i] = (uint16_t*)wf_lo[i];
        b.wt_hi[i] = (uint16_t*)wt_hi[i]; b.wt_lo[i] = (uint16_t*)wt_lo[i];
        b.K[i] = K[i]; b.C[i] = C[i]; b.taps[i] = taps ? taps[i] : 9;
        if ((int64_t)K[i] * C[i] > mx) mx = (int64_t)K[i] * C[i];
    }
    int gx = (int)((mx / 2 + 255) / 256);
    if (gx > 256) gx = 256;
    launch_pdl(weights_to_planes_batch_kernel, dim3(gx, 2 * n), dim3(256), 0, s, b); clb::count_launch();
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------------ max-pool 2x2 / 2
// Window scanned row-major with strict '>' so the FIRST maximum wins (ATen semantics, like clb_maxpool_fwd).
struct U8x8 { uint32_t a, b; };

// planes [N][H][W][C] -> planes [N][H/2][W/2][C] (OUT_NCHW: fp32 [N][C][H/2][W/2], the classifier's flatten order) + argmax
// [N][H/2][W/2][C] u8.  One thread = 8 channels of one output pixel.
// OUT: 0 = planes NHWC, 1 = fp32 [N][C][PH][PW], 2 = planes in that same flatten order (input of a planes nn.Linear)
template <int OUT>
__global__ void __launch_bounds__(256) pool_planes_fwd_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                              uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                                                              float* __restrict__ y_f32, uint8_t* __restrict__ am, int N, int H,
                                                              int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int PH = H >> 1, PW = W >> 1, C8 = C >> 3;
    const int64_t total = (int64_t)N * PH * PW * C8, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int c8 = (int)(i % C8);
        const int64_t op = i / C8;                                   // output pixel (n, ph, pw)
        const int pw = (int)(op % PW), ph = (int)((op / PW) % PH), n = (int)(op / ((int64_t)PW * PH));
        float best[8];
        uint32_t bh[8], bl[8], bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -FLT_MAX; bh[j] = bl[j] = bi[j] = 0; }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int64_t ip = (((int64_t)n * H + 2 * ph + (t >> 1)) * W + 2 * pw + (t & 1)) * C + c8 * 8;
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(x_hi + ip)), l = __ldg(reinterpret_cast<const uint4*>(x_lo + ip));
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t hb = (j & 1) ? hw[j >> 1] >> 16 : hw[j >> 1] & 0xFFFFu;
                const uint32_t lb = (j & 1) ? lw[j >> 1] >> 16 : lw[j >> 1] & 0xFFFFu;
                const float v = join1(hb, lb);
                if (v > best[j] || v != v) { best[j] = v; bh[j] = hb; bl[j] = lb; bi[j] = t; }
            }
        }
        const int64_t o = op * C + c8 * 8;
        if (OUT == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y_f32[(((int64_t)n * C + c8 * 8 + j) * PH + ph) * PW + pw] = best[j];
        } else if (OUT == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t q = (((int64_t)n * C + c8 * 8 + j) * PH + ph) * PW + pw;
                y_hi[q] = (uint16_t)bh[j]; y_lo[q] = (uint16_t)bl[j];
            }
        } else {
            *reinterpret_cast<uint4*>(y_hi + o) = make_uint4(bh[0] | (bh[1] << 16), bh[2] | (bh[3] << 16), bh[4] | (bh[5] << 16), bh[6] | (bh[7] << 16));
            *reinterpret_cast<uint4*>(y_lo + o) = make_uint4(bl[0] | (bl[1] << 16), bl[2] | (bl[3] << 16), bl[4] | (bl[5] << 16), bl[6] | (bl[7] << 16));
        }
        *reinterpret_cast<uint2*>(am + o) = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24),
                                                        bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
    }
}

// fp32 NCHW [N][C][H][W] -> planes [N][H/2][W/2][C] + argmax: one CTA = one output row (n, ph) x 64 channels, transposed
// through shared memory so that reads (along w) and writes (along c) are both coalesced.  First conv layer's output.
__global__ void __launch_bounds__(256) pool_nchw_to_planes_kernel(const float* __restrict__ x, uint16_t* __restrict__ y_hi,
                                                                  uint16_t* __restrict__ y_lo, uint8_t* __restrict__ am, int N,
                                                                  int C, int H, int W) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];                                   // [PW][65] values, then [PW][65] arg-max as float bits
    const int PH = H >> 1, PW = W >> 1;
    const int cb = blockIdx.y * 64, ph = blockIdx.x % PH, n = blockIdx.x / PH;
    float* sv = sm;
    uint32_t* si = reinterpret_cast<uint32_t*>(sm + PW * 65);
    for (int i = threadIdx.x; i < 64 * PW; i += blockDim.x) {
        const int pw = i % PW, c = i / PW;
        const float* src = x + (((int64_t)n * C + cb + c) * H + 2 * ph) * W + 2 * pw;
        const float2 a = __ldg(reinterpret_cast<const float2*>(src)), b = __ldg(reinterpret_cast<const float2*>(src + W));
        float best = -FLT_MAX;
        uint32_t bi = 0;
        const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (v[t] > best || v[t] != v[t]) { best = v[t]; bi = t; }
        sv[pw * 65 + c] = best;
        si[pw * 65 + c] = bi;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * PW; i += blockDim.x) {
        const int c = i & 63, pw = i >> 6;
        uint32_t hi, lo;
        split1(sv[pw * 65 + c], hi, lo);
        const int64_t o = (((int64_t)n * PH + ph) * PW + pw) * C + cb + c;
        y_hi[o] = (uint16_t)hi; y_lo[o] = (uint16_t)lo; am[o] = (uint8_t)si[pw * 65 + c];
    }
}

// backward, planes out: dx[n][2ph+r][2pw+s][c] = (argmax == 2r+s && pooled > 0) ? dy[n][ph][pw][c] : 0   (max-pool backward
// fused with the ReLU mask of the conv in front: pooled > 0 <=> the selected pre-pool activation was > 0).
// IN: 0 = dy / pooled as planes NHWC, 1 = fp32 [N][C][PH][PW] (classifier side), 2 = planes in that flatten order.
template <int IN>
__global__ void __launch_bounds__(256) pool_planes_bwd_kernel(const uint16_t* __restrict__ dy_hi, const uint16_t* __restrict__ dy_lo,
                                                              const float* __restrict__ dy_f32, const uint16_t* __restrict__ pooled_hi,
                                                              const float* __restrict__ pooled_f32, const uint8_t* __restrict__ am,
                                                              uint16_t* __restrict__ dx_hi, uint16_t* __restrict__ dx_lo, int N, int H,
                                                              int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int PH = H >> 1, PW = W >> 1, C8 = C >> 3;
    const int64_t total = (int64_t)N * PH * PW * C8, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int c8 = (int)(i % C8);
        const int64_t op = i / C8;
        const int pw = (int)(op % PW), ph = (int)((op / PW) % PH), n = (int)(op / ((int64_t)PW * PH));
        const int64_t o = op * C + c8 * 8;
        uint32_t gh[8], gl[8], idx[8];
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(am + o));
#pragma unroll
        for (int j = 0; j < 8; ++j) idx[j] = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xFFu;
        if (IN == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t q = (((int64_t)n * C + c8 * 8 + j) * PH + ph) * PW + pw;
                const float g = __ldg(pooled_f32 + q) > 0.f ? __ldg(dy_f32 + q) : 0.f;
                split1(g, gh[j], gl[j]);
            }
        } else if (IN == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t q = (((int64_t)n * C + c8 * 8 + j) * PH + ph) * PW + pw;
                const uint32_t mb = pooled_hi[q];
                const bool on = mb != 0 && mb < 0x8000u;
                gh[j] = on ? dy_hi[q] : 0u;
                gl[j] = on ? dy_lo[q] : 0u;
            }
        } else {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(dy_hi + o)), l = __ldg(reinterpret_cast<const uint4*>(dy_lo + o));
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(pooled_hi + o));
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w}, mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t mb = (j & 1) ? mw[j >> 1] >> 16 : mw[j >> 1] & 0xFFFFu;
                const bool on = mb != 0 && mb < 0x8000u;
                gh[j] = on ? ((j & 1) ? hw[j >> 1] >> 16 : hw[j >> 1] & 0xFFFFu) : 0u;
                gl[j] = on ? ((j & 1) ? lw[j >> 1] >> 16 : lw[j >> 1] & 0xFFFFu) : 0u;
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t oh[8], ol[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { oh[j] = idx[j] == (uint32_t)t ? gh[j] : 0u; ol[j] = idx[j] == (uint32_t)t ? gl[j] : 0u; }
            const int64_t ip = (((int64_t)n * H + 2 * ph + (t >> 1)) * W + 2 * pw + (t & 1)) * C + c8 * 8;
            *reinterpret_cast<uint4*>(dx_hi + ip) = make_uint4(oh[0] | (oh[1] << 16), oh[2] | (oh[3] << 16), oh[4] | (oh[5] << 16), oh[6] | (oh[7] << 16));
            *reinterpret_cast<uint4*>(dx_lo + ip) = make_uint4(ol[0] | (ol[1] << 16), ol[2] | (ol[3] << 16), ol[4] | (ol[5] << 16), ol[6] | (ol[7] << 16));
        }
    }
}

// backward, fp32 NCHW out (the first conv layer's dY): one CTA = one pooled row (n, ph) x 64 channels; reads planes along c,
// writes the two full-resolution rows of every channel along w.
__global__ void __launch_bounds__(256) pool_planes_bwd_to_nchw_kernel(const uint16_t* __restrict__ dy_hi, const uint16_t* __restrict__ dy_lo,
                                                                      const uint16_t* __restrict__ pooled_hi, const uint8_t* __restrict__ am,
                                                                      float* __restrict__ dx, int N, int C, int H, int W) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];                                   // [64][2][W + 1]
    const int PH = H >> 1, PW = W >> 1, ldw = W + 1;
    const int cb = blockIdx.y * 64, ph = blockIdx.x % PH, n = blockIdx.x / PH;
    for (int i = threadIdx.x; i < 64 * PW; i += blockDim.x) {
        const int c = i & 63, pw = i >> 6;
        const int64_t o = (((int64_t)n * PH + ph) * PW + pw) * C + cb + c;
        const uint32_t mb = pooled_hi[o];
        const float g = (mb != 0 && mb < 0x8000u) ? join1(dy_hi[o], dy_lo[o]) : 0.f;
        const uint32_t idx = am[o];
#pragma unroll
        for (int t = 0; t < 4; ++t) sm[(c * 2 + (t >> 1)) * ldw + 2 * pw + (t & 1)] = idx == (uint32_t)t ? g : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 2 * W; i += blockDim.x) {
        const int w = i % W, r = (i / W) & 1, c = i / (2 * W);
        dx[(((int64_t)n * C + cb + c) * H + 2 * ph + r) * W + w] = sm[(c * 2 + r) * ldw + w];
    }
}

int pool_fwd(const uint16_t* x_hi, const uint16_t* x_lo, uint16_t* y_hi, uint16_t* y_lo, float* y_f32, uint8_t* am, int N, int H,
             int W, int C, int flat, cudaStream_t s) {
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
    if (y_f32) launch_pdl(pool_planes_fwd_kernel<1>, dim3(ew_grid(total, 256)), dim3(256), 0, s, x_hi, x_lo, nullptr, nullptr, y_f32, am, N, H, W, C);
    else if (flat) launch_pdl(pool_planes_fwd_kernel<2>, dim3(ew_grid(total, 256)), dim3(256), 0, s, x_hi, x_lo, y_hi, y_lo, nullptr, am, N, H, W, C);
    else launch_pdl(pool_planes_fwd_kernel<0>, dim3(ew_grid(total, 256)), dim3(256), 0, s, x_hi, x_lo, y_hi, y_lo, nullptr, am, N, H, W, C);
    clb::count_launch();
    return CLB_OK;
}
int pool_fwd_from_nchw(const float* x, uint16_t* y_hi, uint16_t* y_lo, uint8_t* am, int N, int C, int H, int W, cudaStream_t s) {
    const size_t smem = (size_t)(W / 2) * 65 * 8;
    launch_pdl(pool_nchw_to_planes_kernel, dim3(N * (H / 2), C / 64), dim3(256), smem, s, x, y_hi, y_lo, am, N, C, H, W); clb::count_launch();
    return CLB_OK;
}
int pool_bwd(const uint16_t* dy_hi, const uint16_t* dy_lo, const float* dy_f32, const uint16_t* pooled_hi, const float* pooled_f32,
             const uint8_t* am, uint16_t* dx_hi, uint16_t* dx_lo, int N, int H, int W, int C, int flat, cudaStream_t s) {
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
    if (dy_f32) launch_pdl(pool_planes_bwd_kernel<1>, dim3(ew_grid(total, 256)), dim3(256), 0, s, nullptr, nullptr, dy_f32, nullptr, pooled_f32, am, dx_hi, dx_lo, N, H, W, C);
    else if (flat) launch_pdl(pool_planes_bwd_kernel<2>, dim3(ew_grid(total, 256)), dim3(256), 0, s, dy_hi, dy_lo, nullptr, pooled_hi, nullptr, am, dx_hi, dx_lo, N, H, W, C);
    else launch_pdl(pool_planes_bwd_kernel<0>, dim3(ew_grid(total, 256)), dim3(256), 0, s, dy_hi, dy_lo, nullptr, pooled_hi, nullptr, am, dx_hi, dx_lo, N, H, W, C);
    clb::count_launch();
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------------ fp32 <-> planes
__global__ void __launch_bounds__(256) planes_to_f32_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                                            float* __restrict__ out, int64_t n) {
    pdl_trigger();
    pdl_wait();
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) out[i] = join1(hi[i], lo[i]);
}
// hi/lo planes of x, zeroed where mask_hi (a post-ReLU activation's hi plane, may be NULL) is <= 0: the ReLU backward of a
// planes nn.Linear applied to the fp32 gradient that arrives from a layer outside the planes pipeline
__global__ void __launch_bounds__(256) f32_to_planes_kernel(const float* __restrict__ x, const uint16_t* __restrict__ mask_hi,
                                                            uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t n) {
    pdl_trigger();
    pdl_wait();
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        float v = x[i];
        if (mask_hi) { const uint32_t mb = mask_hi[i]; if (!(mb != 0 && mb < 0x8000u)) v = 0.f; }
        uint32_t h, l;
        split1(v, h, l);
        hi[i] = (uint16_t)h; lo[i] = (uint16_t)l;
    }
}
int planes_to_f32(const uint16_t* hi, const uint16_t* lo, float* out, int64_t n, cudaStream_t s) {
    launch_pdl(planes_to_f32_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, s, hi, lo, out, n); clb::count_launch();
    return CLB_OK;
}
int f32_to_planes(const float* x, const uint16_t* mask_hi, uint16_t* hi, uint16_t* lo, int64_t n, cudaStream_t s) {
    launch_pdl(f32_to_planes_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, s, x, mask_hi, hi, lo, n); clb::count_launch();
    return CLB_OK;
}
int pool_bwd_to_nchw(const uint16_t* dy_hi, const uint16_t* dy_lo, const uint16_t* pooled_hi, const uint8_t* am, float* dx, int N,
                     int C, int H, int W, cudaStream_t s) {
    const size_t smem = (size_t)64 * 2 * (W + 1) * 4;
    launch_pdl(pool_planes_bwd_to_nchw_kernel, dim3(N * (H / 2), C / 64), dim3(256), smem, s, dy_hi, dy_lo, pooled_hi, am, dx, N, C, H, W); clb::count_launch();
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------------ bias gradient
// db[k] = sum over pixels of dY[pix][k].  Two deterministic stages: kBiasChunks CTAs sum a contiguous pixel range each (thread =
// channel pair, fixed order), then one pass adds the chunk partials in order.  No atomics: bit-reproducible.
constexpr int kBiasChunks = 592;
__global__ void __launch_bounds__(256, 6) bias_partials_kernel(const uint16_t* __restrict__ dy_hi, const uint16_t* __restrict__ dy_lo,
                                                            float* __restrict__ part, int64_t npix, int K, int64_t pix_per_chunk) {
    pdl_trigger();
    pdl_wait();
    // thread = 8 channels (one 16-byte load per plane) of every `rows`-th pixel of the chunk; channel groups beyond the CTA's
    // 256 threads are handled in further passes
    const int K8 = K >> 3;
    const int tpr = K8 < 256 ? K8 : 256;                            // threads per pixel row
    const int rows = 256 / tpr;                                     // pixel rows in flight
    const int sub = threadIdx.x / tpr, lc = threadIdx.x - sub * tpr;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_chunk, p1 = min(npix, p0 + pix_per_chunk);
    __shared__ float red[256 * 9];
    for (int c0 = 0; c0 < K8; c0 += tpr) {                          // same trip count for every thread
        const int c8 = c0 + lc;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (sub < rows && c8 < K8) {
            const uint4* hp = reinterpret_cast<const uint4*>(dy_hi) + c8;
            const uint4* lp = reinterpret_cast<const uint4*>(dy_lo) + c8;
#pragma unroll 4
            for (int64_t p = p0 + sub; p < p1; p += rows) {
                const uint4 h = __ldg(hp + p * K8), l = __ldg(lp + p * K8);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[2 * j] += __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
                    acc[2 * j + 1] += __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 9 + j] = acc[j];
        __syncthreads();
        const int kc = (tpr * 8 < K - c0 * 8) ? tpr * 8 : K - c0 * 8;      // channels of this pass
        for (int i = threadIdx.x; i < kc; i += 256) {                    // fixed-order sum over the rows
            const int cc8 = i >> 3, j = i & 7;
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += red[(r * tpr + cc8) * 9 + j];
            part[(int64_t)blockIdx.x * K + c0 * 8 + i] = s;
        }
        __syncthreads();
    }
}
// one warp per channel: lane l adds chunks l, l + 32, ... in order, then a fixed xor-shuffle tree (deterministic)
__global__ void __launch_bounds__(256) bias_final_kernel(const float* __restrict__ part, float* __restrict__ db, int K, int chunks) {
    pdl_trigger();
    pdl_wait();
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= K) return;
    float s = 0.f;
    for (int c = lane; c < chunks; c += 32) s += part[(int64_t)c * K + k];
    s = warp_sum(s);
    if (lane == 0) db[k] = s;
}
size_t bias_ws_floats(int K) { return (size_t)kBiasChunks * K; }
int bias_grad(const uint16_t* dy_hi, const uint16_t* dy_lo, float* db, float* part, int64_t npix, int K, cudaStream_t s) {
    const int64_t per = (npix + kBiasChunks - 1) / kBiasChunks;
    const int chunks = (int)((npix + per - 1) / per);
    launch_pdl(bias_partials_kernel, dim3(chunks), dim3(256), 0, s, dy_hi, dy_lo, part, npix, K, per); clb::count_launch();
    launch_pdl(bias_final_kernel, dim3((K + 7) / 8), dim3(256), 0, s, part, db, K, chunks); clb::count_launch();
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------------ split-K reduction
// dw[k][c][tap] = sum_z ws[z][k][tap][c]   (fixed summation order -> bit-reproducible), optionally followed in the same pass
// by the importance update of the reference's passes over the batch gradient (J-1: "fused into the backward pass"):
//   imp_mode 1 (EWC, main_EWC.py:151-156):  omega += dw * dw / imp_a
//   imp_mode 2 (MAS, train_MAS.py:163-177): omega = (omega * imp_a + |dw|) / imp_b
template <int TAPS>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, float* __restrict__ omega,
                                                           int K, int C, int splits, int imp_mode, float imp_a, float imp_b) {
    pdl_trigger();
    pdl_wait();
    // warp = 32 consecutive channels of one filter k: reads are coalesced along c (ws is [z][k][tap][c]), the 288 results are
    // transposed through shared memory and written as one contiguous run of dw[k][c0..c0+31][9]
    __shared__ float buf[8][32 * TAPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = (int64_t)K * C * TAPS;
    const int cg = C >> 5;                                           // C % 64 == 0
    const int64_t n_groups = (int64_t)K * cg;
    for (int64_t g = (int64_t)blockIdx.x * 8 + warp; g < n_groups; g += (int64_t)gridDim.x * 8) {
        const int k = (int)(g / cg), c0 = (int)(g - (int64_t)k * cg) * 32;
        const float* src = ws + ((int64_t)k * TAPS) * C + c0 + lane;
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {                 // four interleaved partial sums: loads in flight, order still fixed
            const float* q = src + (int64_t)t * C;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int z = 0;
            for (; z + 4 <= splits; z += 4) {
                a0 += q[(int64_t)z * total]; a1 += q[(int64_t)(z + 1) * total];
                a2 += q[(int64_t)(z + 2) * total]; a3 += q[(int64_t)(z + 3) * total];
            }
            for (; z < splits; ++z) a0 += q[(int64_t)z * total];
            buf[warp][lane * TAPS + t] = (a0 + a1) + (a2 + a3);
        }
        __syncwarp();
        const int64_t o0 = ((int64_t)k * C + c0) * TAPS;
#pragma unroll
        for (int it = 0; it < TAPS; ++it) {
            const float s = buf[warp][it * 32 + lane];
            const int64_t o = o0 + it * 32 + lane;
            dw[o] = s;
            if (imp_mode == 1) omega[o] = __fadd_rn(omega[o], __fdiv_rn(__fmul_rn(s, s), imp_a));
            else if (imp_mode == 2) omega[o] = __fdiv_rn(__fadd_rn(__fmul_rn(omega[o], imp_a), fabsf(s)), imp_b);
        }
        __syncwarp();
    }
}
int wgrad_reduce(const float* ws, float* dw, float* omega, int K, int C, int taps, int splits, int imp_mode, float imp_a, float imp_b,
                 cudaStream_t s) {
    if (taps == 9) launch_pdl(wgrad_reduce_kernel<9>, dim3(ew_grid((int64_t)K * C, 32)), dim3(256), 0, s, ws, dw, omega, K, C, splits, imp_mode, imp_a, imp_b);
    else launch_pdl(wgrad_reduce_kernel<1>, dim3(ew_grid((int64_t)K * C, 32)), dim3(256), 0, s, ws, dw, omega, K, C, splits, imp_mode, imp_a, imp_b);
    clb::count_launch();
    return CLB_OK;
}

}  // namespace pl
}  // namespace clb
