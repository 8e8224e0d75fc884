// First conv layer of the VGG stacks (3 input channels, 3x3 / stride 1 / pad 1, 64 filters) fused with what follows it:
//   forward : conv + bias + ReLU + 2x2 max-pool  ->  bf16 hi/lo planes [N][H/2][W/2][64] + arg-max
//   backward: max-pool backward + ReLU backward + weight / bias gradient (no data gradient: it is the first layer)
// Replaces features[0..2] of src/models/VGGSlim.py:27-40 (nn.Conv2d(3, 64, 3, padding=1), nn.ReLU, nn.MaxPool2d(2, 2)) and
// their autograd backward.  K = 27 is no tensor-core shape and the layer is bound by its 210 MB fp32 output (N = 200,
// 64x64): fused, that tensor is never written -- the forward writes 65 MB, the backward reads 92 MB -- and the arithmetic
// (2.8 GFLOP each way, exact fp32 FFMA) hides under it.
#include <float.h>

#include "clb_planes.cuh"

namespace clb {
namespace pl {

constexpr int kF1K = 64, kF1Taps = 27;

// tile[c][i][1 + w]: input rows 2ph-1 .. 2ph+2 of the three channels, zero padded; row pitch W + 4 (16-byte multiple).
// Loaded with cp.async (zero fill outside the image) into one of two buffers, so that the tile of the CTA's next row
// streams in underneath the arithmetic of the current one.
__device__ __forceinline__ void load_tile_async(const float* __restrict__ x, float* tile, int n, int ph, int H, int W, int pitch) {
    for (int i = threadIdx.x; i < 3 * 4 * pitch; i += blockDim.x) {
        const int col = i % pitch, row = (i / pitch) & 3, c = i / (4 * pitch);
        const int h = 2 * ph - 1 + row, w = col - 1;
        const bool ok = (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
        const float* src = ok ? x + (((int64_t)n * 3 + c) * H + h) * W + w : x;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void wait_tile() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// thread = 8 filters (kq = lane & 7) of one pooled pixel (pw = 4 * warp + (lane >> 3), + 32 per pass): the 8 x 16-byte
// stores of a pixel are one contiguous 128-byte line per plane
__global__ void __launch_bounds__(256, 2) conv1_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, uint16_t* __restrict__ y_hi,
                                                             uint16_t* __restrict__ y_lo, uint8_t* __restrict__ am, int N, int H, int W) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    const int PH = H >> 1, PW = W >> 1, pitch = W + 4;
    float* w_s = sm;                                  // [27][64]
    float* tiles = sm + kF1Taps * kF1K;               // 2 x [3][4][pitch]
    const int tile_elems = 12 * pitch;
    for (int i = threadIdx.x; i < kF1Taps * kF1K; i += blockDim.x) {
        const int k = i & 63, t = i >> 6;             // w[k][c][r][s], t = c * 9 + r * 3 + s
        w_s[i] = __ldg(w + k * kF1Taps + t);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kq = lane & 7, pwl = 4 * warp + (lane >> 3);
    float b8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b8[j] = bias ? __ldg(bias + kq * 8 + j) : 0.f;
    int buf = 0;
    if ((int)blockIdx.x < N * PH) load_tile_async(x, tiles, blockIdx.x / PH, blockIdx.x % PH, H, W, pitch);
    for (int row = blockIdx.x; row < N * PH; row += gridDim.x, buf ^= 1) {
        const int n = row / PH, ph = row - n * PH;
        const float* tile = tiles + buf * tile_elems;
        wait_tile();
        __syncthreads();                              // this row's tile is complete; everyone is done with the other buffer
        const int nrow = row + gridDim.x;
        if (nrow < N * PH) load_tile_async(x, tiles + (buf ^ 1) * tile_elems, nrow / PH, nrow % PH, H, W, pitch);
        for (int pw = pwl; pw < PW; pw += 32) {
            float acc[4][8];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                float p[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a = *reinterpret_cast<const float2*>(tile + (c * 4 + i) * pitch + 2 * pw);
                    const float2 b = *reinterpret_cast<const float2*>(tile + (c * 4 + i) * pitch + 2 * pw + 2);
                    p[i][0] = a.x; p[i][1] = a.y; p[i][2] = b.x; p[i][3] = b.y;
                }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const float4 w0 = *reinterpret_cast<const float4*>(w_s + (c * 9 + r * 3 + s) * kF1K + kq * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(w_s + (c * 9 + r * 3 + s) * kF1K + kq * 8 + 4);
                        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float xv = p[(t >> 1) + r][(t & 1) + s];
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(xv, wv[j], acc[t][j]);
                        }
                    }
            }
            uint32_t hh[8], ll[8], bi[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float best = -FLT_MAX;
                uint32_t idx = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float v = fmaxf(acc[t][j] + b8[j], 0.f);
                    if (v > best || v != v) { best = v; idx = t; }
                }
                split1(best, hh[j], ll[j]);
                bi[j] = idx;
            }
            const int64_t o = (((int64_t)n * PH + ph) * PW + pw) * kF1K + kq * 8;
            *reinterpret_cast<uint4*>(y_hi + o) = make_uint4(hh[0] | (hh[1] << 16), hh[2] | (hh[3] << 16), hh[4] | (hh[5] << 16), hh[6] | (hh[7] << 16));
            *reinterpret_cast<uint4*>(y_lo + o) = make_uint4(ll[0] | (ll[1] << 16), ll[2] | (ll[3] << 16), ll[4] | (ll[5] << 16), ll[6] | (ll[7] << 16));
            *reinterpret_cast<uint2*>(am + o) = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24),
                                                            bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
        }
    }
}

// backward: thread = 2 filters (lane = filter pair) x the pooled pixels pw = warp, warp + 8, ... of the CTA's rows; the input
// patch of a pixel is the same for all lanes (shared-memory broadcast).  g = (pooled > 0) ? d_pooled : 0 is routed to the
// arg-max position by predication (4 x 27 FMAs with three of the four factors zero: no dynamic register indexing).
// part[CTA][64][28]: 27 weight taps + bias; the CTA's eight warps are added in warp order, the CTAs in CTA order.
constexpr int kF1Ctas = 592;
__global__ void __launch_bounds__(256) conv1_pool_bwd_kernel(const float* __restrict__ x, const uint16_t* __restrict__ dp_hi,
                                                             const uint16_t* __restrict__ dp_lo, const uint16_t* __restrict__ pooled_hi,
                                                             const uint8_t* __restrict__ am, float* __restrict__ part, int N, int H, int W,
                                                             int rows_per_cta) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    const int PH = H >> 1, PW = W >> 1, pitch = W + 4;
    float* tiles = sm;                                // 2 x [3][4][pitch]
    const int tile_elems = 12 * pitch;
    float* red = sm + 2 * tile_elems;                 // [64][28]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[2][28];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int t = 0; t < 28; ++t) acc[q][t] = 0.f;
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(N * PH, row0 + rows_per_cta);
    int buf = 0;
    if (row0 < row1) load_tile_async(x, tiles, row0 / PH, row0 % PH, H, W, pitch);
    for (int row = row0; row < row1; ++row, buf ^= 1) {
        const int n = row / PH, ph = row - n * PH;
        const float* tile = tiles + buf * tile_elems;
        wait_tile();
        __syncthreads();
        if (row + 1 < row1) load_tile_async(x, tiles + (buf ^ 1) * tile_elems, (row + 1) / PH, (row + 1) % PH, H, W, pitch);
        for (int pw = warp; pw < PW; pw += 8) {
            const int64_t o = (((int64_t)n * PH + ph) * PW + pw) * kF1K + 2 * lane;
            const uint32_t h2 = __ldg(reinterpret_cast<const uint32_t*>(dp_hi + o)), l2 = __ldg(reinterpret_cast<const uint32_t*>(dp_lo + o));
            const uint32_t m2 = __ldg(reinterpret_cast<const uint32_t*>(pooled_hi + o));
            const uint32_t a2 = *reinterpret_cast<const unsigned short*>(am + o);
            float g[2];
            uint32_t idx[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t mb = q ? m2 >> 16 : m2 & 0xFFFFu;
                const float v = q ? __uint_as_float(h2 & 0xFFFF0000u) + __uint_as_float(l2 & 0xFFFF0000u)
                                  : __uint_as_float(h2 << 16) + __uint_as_float(l2 << 16);
                g[q] = (mb != 0 && mb < 0x8000u) ? v : 0.f;
                idx[q] = (a2 >> (8 * q)) & 0xFFu;
                acc[q][27] += g[q];
            }
            float gt[2][4];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int t = 0; t < 4; ++t) gt[q][t] = idx[q] == (uint32_t)t ? g[q] : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float p[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a = *reinterpret_cast<const float2*>(tile + (c * 4 + i) * pitch + 2 * pw);
                    const float2 b = *reinterpret_cast<const float2*>(tile + (c * 4 + i) * pitch + 2 * pw + 2);
                    p[i][0] = a.x; p[i][1] = a.y; p[i][2] = b.x; p[i][3] = b.y;
                }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int s = 0; s < 3; ++s)
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float xv = p[(t >> 1) + r][(t & 1) + s];
                            acc[0][c * 9 + r * 3 + s] = fmaf(gt[0][t], xv, acc[0][c * 9 + r * 3 + s]);
                            acc[1][c * 9 + r * 3 + s] = fmaf(gt[1][t], xv, acc[1][c * 9 + r * 3 + s]);
                        }
            }
        }
    }
    __syncthreads();
    for (int wv = 0; wv < 8; ++wv) {                  // fixed warp order: deterministic
        if (warp == wv) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int t = 0; t < 28; ++t) {
                    float* dst = red + (2 * lane + q) * 28 + t;
                    *dst = wv == 0 ? acc[q][t] : *dst + acc[q][t];
                }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kF1K * 28; i += blockDim.x) part[(int64_t)blockIdx.x * (kF1K * 28) + i] = red[i];
}
// dw[k][27], db[k] = sum over CTAs (warp per output, lanes stride the CTAs in order, fixed shuffle tree)
__global__ void __launch_bounds__(256) conv1_bwd_final_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ db,
                                                              int chunks) {
    pdl_trigger();
    pdl_wait();
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= kF1K * 28) return;
    float s = 0.f;
    for (int c = lane; c < chunks; c += 32) s += part[(int64_t)c * (kF1K * 28) + o];
    s = warp_sum(s);
    if (lane == 0) {
        const int k = o / 28, t = o - k * 28;
        if (t < 27) dw[k * 27 + t] = s;
        else if (db) db[k] = s;
    }
}

bool first_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    return C == 3 && K == kF1K && R == 3 && S == 3 && stride == 1 && pad == 1 && (H % 2) == 0 && (W % 2) == 0 && W >= 4 && W <= 256;
}
int conv1_pool_fwd(const float* x, const float* w, const float* bias, uint16_t* y_hi, uint16_t* y_lo, uint8_t* am, int N, int H, int W,
                   cudaStream_t s) {
    const size_t smem = (size_t)(kF1Taps * kF1K + 2 * 12 * (W + 4)) * 4;
    int grid = N * (H / 2);
    if (grid > sm_count() * 6) grid = sm_count() * 6;
    launch_pdl(conv1_pool_fwd_kernel, dim3(grid), dim3(256), smem, s, x, w, bias, y_hi, y_lo, am, N, H, W); clb::count_launch();
    return CLB_OK;
}
size_t conv1_bwd_ws_floats() { return (size_t)kF1Ctas * kF1K * 28; }
int conv1_pool_bwd(const float* x, const uint16_t* dp_hi, const uint16_t* dp_lo, const uint16_t* pooled_hi, const uint8_t* am, float* dw,
                   float* db, float* part, int N, int H, int W, cudaStream_t s) {
    const int rows = N * (H / 2);
    const int per = (rows + kF1Ctas - 1) / kF1Ctas, ctas = (rows + per - 1) / per;
    const size_t smem = (size_t)(2 * 12 * (W + 4) + kF1K * 28) * 4;
    launch_pdl(conv1_pool_bwd_kernel, dim3(ctas), dim3(256), smem, s, x, dp_hi, dp_lo, pooled_hi, am, part, N, H, W, per); clb::count_launch();
    launch_pdl(conv1_bwd_final_kernel, dim3((kF1K * 28 + 7) / 8), dim3(256), 0, s, part, dw, db, ctas); clb::count_launch();
    return CLB_OK;
}

}  // namespace pl
}  // namespace clb

using namespace clb;
extern "C" {

int clb_planes_conv1_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    return pl::first_supported(C, H, W, K, R, S, stride, pad) ? 1 : 0;
}
int clb_planes_conv1_pool_fwd(const float* x, const float* w, const float* bias, void* y_hi, void* y_lo, uint8_t* argmax, int N, int C,
                              int H, int W, int K, void* stream) {
    CLB_CHECK_ARG(x && w && y_hi && y_lo && argmax && N > 0 && pl::first_supported(C, H, W, K, 3, 3, 1, 1));
    int rc = pl::conv1_pool_fwd(x, w, bias, (uint16_t*)y_hi, (uint16_t*)y_lo, argmax, N, H, W, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}
size_t clb_planes_conv1_ws(void) { return pl::conv1_bwd_ws_floats() * sizeof(float); }
int clb_planes_conv1_pool_bwd(const float* x, const void* dp_hi, const void* dp_lo, const void* pooled_hi, const uint8_t* argmax, float* dw,
                              float* dbias, float* ws, size_t ws_bytes, int N, int C, int H, int W, int K, void* stream) {
    CLB_CHECK_ARG(x && dp_hi && dp_lo && pooled_hi && argmax && dw && ws && N > 0 && pl::first_supported(C, H, W, K, 3, 3, 1, 1));
    if (ws_bytes < clb_planes_conv1_ws()) { set_error("clb_planes_conv1_pool_bwd: workspace too small"); return CLB_EWORKSPACE; }
    int rc = pl::conv1_pool_bwd(x, (const uint16_t*)dp_hi, (const uint16_t*)dp_lo, (const uint16_t*)pooled_hi, argmax, dw, dbias, ws, N, H, W,
                                as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

}  // extern "C"
