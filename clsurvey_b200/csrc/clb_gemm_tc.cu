// tcgen05 (5th-gen tensor core) implicit-GEMM path for the 3x3 stride-1 convolutions of the hot path:
// conv2d fwd, dgrad (as a forward conv of dY over flipped weights) and wgrad (split-K, deterministic 2-pass).
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] (smem, K-major, 128B swizzle) * B[BN x 32]^T (smem, K-major, 128B swizzle)
//
// Precision: the reference computes in fp32 (cuDNN/MKL-DNN fp32; north_star: loss and importance weights within 1e-4
// of the reference).  tcgen05 has no fp32 MMA, so the parity mode is a 3-pass TF32 split: x = hi + lo with
// hi = x & 0xFFFFE000 (exact 10-bit-mantissa TF32), lo = x - hi (exact), and
//     a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b      (dropped lo*lo term <= 2^-20 relative), fp32 accumulation in TMEM.
// CLB_MM_TF32X1 issues only the hi*hi pass (fast, NOT parity mode).
//
// This file holds the host-side dispatch of the NCHW tensor-core kernels (clb_gemm_tc2.cu: A through TMEM; clb_gemm_tc3.cu:
// TMA-fed weights / dY; clb_gemm_tc4.cu: bf16 hi/lo split) and their helper kernels (weight re-ordering, split-K reduce).
// They serve the layers the planes pipeline (clb_planes_*.cu) does not take: nets with C % 64 != 0, AlexNet, Linear.
// GEMM-K order is (r, s, c) so that one 32-wide K block has a single (r, s): one bounds test per block per pixel.
#include <stdlib.h>

#include "clb_tc_ptx.cuh"

namespace clb {
namespace tc {

// ---------------------------------------------------------------------------------------------- helper kernels
// w[K][C][RS] -> w2[K][ld] with w2[k][rs*C + c]   (forward GEMM-B, K order (r,s,c)); ld > RS*C zero-pads the row
// bf16 planes (mode CLB_MM_BF16X3): hi = bf16_rn(v), lo = bf16_rn(v - hi), stored as 16-bit arrays in the same workspace
__device__ __forceinline__ void store_bf16_split(float v, float* hi_plane, float* lo_plane, int64_t i) {
    uint32_t u = __float_as_uint(v);
    u += 0x7FFFu + ((u >> 16) & 1u);
    const uint32_t h = u >> 16;
    uint32_t r = __float_as_uint(v - __uint_as_float(h << 16));
    r += 0x7FFFu + ((r >> 16) & 1u);
    reinterpret_cast<uint16_t*>(hi_plane)[i] = (uint16_t)h;
    reinterpret_cast<uint16_t*>(lo_plane)[i] = (uint16_t)(r >> 16);
}
__global__ void permute_w_fwd_kernel(const float* __restrict__ w, float* __restrict__ w2, float* __restrict__ w2_lo, int K,
                                     int C, int RS, int ld, int bf16) {
    const int64_t total = (int64_t)K * ld, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int j = (int)(i % ld), k = (int)(i / ld);
        const int c = j % C, rs = j / C;
        const float v = (j < RS * C) ? w[((int64_t)k * C + c) * RS + rs] : 0.f;
        if (bf16) { store_bf16_split(v, w2, w2_lo, i); continue; }
        w2[i] = v;
        if (w2_lo) w2_lo[i] = v - __uint_as_float(__float_as_uint(v) & kHiMask);     // TMA-fed kernels read the lo plane
    }
}
// w[K][C][R][S] -> wd[C][(R-1-r, S-1-s)][K]   (dgrad = forward conv of dY: rows = c, K order (r', s', kout))
__global__ void permute_w_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wd, float* __restrict__ wd_lo, int K,
                                       int C, int R, int S, int bf16) {
    const int64_t total = (int64_t)K * C * R * S, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int k = (int)(i % K), rs = (int)((i / K) % (R * S)), c = (int)(i / ((int64_t)K * R * S));
        const int r = R - 1 - rs / S, s = S - 1 - rs % S;
        const float v = w[(((int64_t)k * C + c) * R + r) * S + s];
        if (bf16) { store_bf16_split(v, wd, wd_lo, i); continue; }
        wd[i] = v;
        if (wd_lo) wd_lo[i] = v - __uint_as_float(__float_as_uint(v) & kHiMask);
    }
}
// dw[K][C][RS] = sum_z ws[z][K][RS][C]   (fixed summation order -> bit-reproducible)
__global__ void splitk_reduce_permute_kernel(const float* __restrict__ ws, float* __restrict__ dw, int K, int C, int RS,
                                             int splits) {
    const int64_t total = (int64_t)K * C * RS, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int c = (int)(i % C), rs = (int)((i / C) % RS), k = (int)(i / ((int64_t)C * RS));
        float s = ws[i];
        for (int zz = 1; zz < splits; ++zz) s += ws[(int64_t)zz * total + i];
        dw[((int64_t)k * C + c) * RS + rs] = s;
    }
}

static inline int ew_blocks(int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ entry points
// (called from clb_gemm_simt.cu's C-ABI functions when the matmul mode selects tensor cores and the shape qualifies)

int tc2_conv_fwd(const float* x, const float* w2, const float* bias, float* y, int N, int C, int H, int W, int K, int R,
                 int S, int pad, int relu, bool with_lo, cudaStream_t s);
int tc2_conv_wgrad(const float* x, const float* dy, float* ws, int N, int C, int H, int W, int K, int R, int S, int pad,
                   bool with_lo, cudaStream_t s);

size_t tc_w_plane_floats(int K, int C, int R, int S);
bool tc3_wgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
size_t tc3_wgrad_extra_floats(int N, int C, int H, int W, int K);
int tc3_conv_fwd(const float* x, const float* w2, const float* w2_lo, const float* bias, float* y, int N, int C, int H, int W,
                 int K, int R, int S, int pad, int relu, bool with_lo, cudaStream_t s);
int tc3_conv_wgrad(const float* x, const float* dy, float* ws_partials, float* bias_part, float* dy_lo, int N, int C, int H, int W,
                   int K, int R, int S, int pad, bool with_lo, int splits, int kb_per_split, cudaStream_t s);

// CLB_TC_IMPL: 2 = A through TMEM (clb_gemm_tc2.cu), 3 (default) = gen-2 plus TMA-fed kernels where the memory layout allows
// it (wgrad on >= 8x8 maps).  The first generation (both operands gathered into smem) was removed in round 2.
static int tc_impl() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("CLB_TC_IMPL");
        v = (e && e[0] >= '2' && e[0] <= '3') ? (e[0] - '0') : 3;
    }
    return v;
}

static void tc3_wgrad_plan(int N, int C, int H, int W, int K, int R, int S, int* splits, int* kb_per_split) {
    const int n_rows = R * S * C, nkb = N * H * W / 32;
    const int64_t tiles = (int64_t)((K + 127) / 128) * ((n_rows + 127) / 128);
    int64_t want = (2LL * 148 + tiles - 1) / tiles;
    const int64_t max_splits = (nkb + 15) / 16;
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    const int per = (int)((nkb + want - 1) / want);
    *kb_per_split = per;
    *splits = (nkb + per - 1) / per;
}

// shapes the tensor-core path takes (everything else stays on the exact-fp32 SIMT kernels):
//   fwd  : stride 1, "same" padding, C % 32 == 0 (one filter tap per K block) or the whole C*R*S <= 32 (first layer; gen-2 only)
//   dgrad: a forward conv of dY, i.e. the fwd rule with C := K
//   wgrad: stride 1, "same" padding, Q % 4 == 0 (16-byte pixel chunks)
static bool same_conv(int H, int W, int R, int S, int stride, int pad) {
    return stride == 1 && R == S && 2 * pad == R - 1 && H > 0 && W > 0;
}
bool tc_fwd_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    if (!same_conv(H, W, R, S, stride, pad)) return false;
    if ((C % 32) == 0) return true;
    // first layer (C*R*S <= 32): one K block per CTA -- the per-CTA prologue (TMEM alloc, barrier init) dominates and the
    // exact-fp32 SIMT kernel is faster (158 us vs 256 us at batch 200); opt-in only
    static int small_c = -1;
    if (small_c < 0) { const char* e = getenv("CLB_TC_SMALLC"); small_c = (e && e[0] == '1') ? 1 : 0; }
    return small_c && C * R * S <= 32;
}
bool tc_dgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    return same_conv(H, W, R, S, stride, pad) && (K % 32) == 0;
}
bool tc_wgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    if (!same_conv(H, W, R, S, stride, pad) || (W % 4) != 0) return false;
    return true;
}

size_t tc_weight_ws_floats(int C, int K, int R, int S) { return (size_t)K * C * R * S; }

int tc_conv_fwd(const float* x, const float* w2 /*[K][RS][C]*/, const float* bias, float* y, int N, int C, int H, int W,
                int K, int R, int S, int pad, int relu, bool with_lo, cudaStream_t s) {
    using namespace tc;
    if (tc_impl() >= 3) {
        // w2 holds [hi plane][lo plane]; the plane size was fixed by whoever re-ordered the weights:
        // fwd: (K outputs, C inputs); dgrad calls us with (C_in := K_out, K_out := C_in) -> the same product K*C*R*S
        const size_t plane = tc_w_plane_floats(K, C, R, S);
        return tc3_conv_fwd(x, w2, w2 + plane, bias, y, N, C, H, W, K, R, S, pad, relu, with_lo, s);
    }
    return tc2_conv_fwd(x, w2, bias, y, N, C, H, W, K, R, S, pad, relu, with_lo, s);
}

void tc_wgrad_plan(int N, int C, int H, int W, int K, int R, int S, int* bn, int* splits, int* kb_per_split) {
    const int n_rows = R * S * C, npix = N * H * W;
    const int BN = (n_rows % 128 == 0) ? 128 : 64;
    const int64_t tiles = (int64_t)((K + 127) / 128) * ((n_rows + BN - 1) / BN);
    const int nkb = (npix + tc::BK - 1) / tc::BK;
    int64_t want = (2LL * 148 + tiles - 1) / tiles;            // ~2 CTAs per SM (fixed => device independent results)
    int64_t max_splits = (nkb + 15) / 16;                      // >= 16 K blocks per split
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    int per = (int)((nkb + want - 1) / want);
    *bn = BN;
    *kb_per_split = per;
    *splits = (nkb + per - 1) / per;
}

size_t tc_wgrad_ws_floats(int N, int C, int H, int W, int K, int R, int S) {
    int bn, splits, per;
    tc_wgrad_plan(N, C, H, W, K, R, S, &bn, &splits, &per);
    size_t need = (size_t)splits * K * C * R * S;
    if (tc_impl() >= 3 && tc3_wgrad_supported(C, H, W, K, R, S, 1, (R - 1) / 2)) {
        int s3, p3;
        tc3_wgrad_plan(N, C, H, W, K, R, S, &s3, &p3);
        const size_t n3 = (size_t)s3 * K * C * R * S + 8 + tc3_wgrad_extra_floats(N, C, H, W, K);
        if (n3 > need) need = n3;
    }
    return need;
}

int tc_conv_wgrad(const float* x, const float* dy, float* dw, float* ws, float* bias_part, bool* bias_partials_done, int N,
                  int C, int H, int W, int K, int R, int S, int pad, bool with_lo, cudaStream_t s) {
    *bias_partials_done = false;
    using namespace tc;
    const int P = H, Q = W, npix = N * P * Q, n_rows = R * S * C;
    int bn, splits, per;
    if (tc_impl() >= 3 && tc3_wgrad_supported(C, H, W, K, R, S, 1, pad)) {
        tc3_wgrad_plan(N, C, H, W, K, R, S, &splits, &per);
        size_t off = ((size_t)splits * K * n_rows + 7) & ~(size_t)3;              // keep the lo plane 16-byte aligned
        float* dy_lo = ws + off;
        if (mm_mode() == CLB_MM_BF16X3 && with_lo && bias_part && tc4_wgrad_supported(H, W, R, S)) {
            int used = splits;
            int rc4 = tc4_conv_wgrad(x, dy, ws, bias_part, dy_lo, N, C, H, W, K, R, S, pad, splits, per, &used, s);
            if (rc4) return rc4;
            *bias_partials_done = true;
            splitk_reduce_permute_kernel<<<ew_blocks((int64_t)K * n_rows), 256, 0, s>>>(ws, dw, K, C, R * S, used); clb::count_launch();
            return CLB_OK;
        }
        int rc3 = tc3_conv_wgrad(x, dy, ws, with_lo ? bias_part : nullptr, dy_lo, N, C, H, W, K, R, S, pad, with_lo, splits, per, s);
        if (rc3) return rc3;
        *bias_partials_done = with_lo && bias_part != nullptr;
        splitk_reduce_permute_kernel<<<ew_blocks((int64_t)K * n_rows), 256, 0, s>>>(ws, dw, K, C, R * S, splits); clb::count_launch();
        return CLB_OK;
    }
    tc_wgrad_plan(N, C, H, W, K, R, S, &bn, &splits, &per);
    const int rc = tc2_conv_wgrad(x, dy, ws, N, C, H, W, K, R, S, pad, with_lo, s);
    if (rc) return rc;
    splitk_reduce_permute_kernel<<<ew_blocks((int64_t)K * n_rows), 256, 0, s>>>(ws, dw, K, C, R * S, splits); clb::count_launch();
    return CLB_OK;
}

// workspace layout for the re-ordered weights: [hi plane: plane floats][lo plane: plane floats], plane = tc_w_plane_floats()
size_t tc_w_plane_floats(int K, int C, int R, int S) {
    const size_t a = (size_t)K * C * R * S, b = (size_t)K * 32, c = (size_t)C * 32;
    size_t m = a > b ? a : b;
    m = m > c ? m : c;
    return (m + 3) & ~(size_t)3;
}
bool tc_bf16_route(int reduction_channels) {
    return mm_mode() == CLB_MM_BF16X3 && tc_impl() >= 3 && tc4_fwd_supported(reduction_channels);
}
int tc_conv_fwd_bf16(const float* x, const float* w, float* w_ws, const float* bias, float* y, int N, int C, int H, int W, int K,
                     int R, int S, int pad, int relu, cudaStream_t s) {
    tc_permute_w_fwd(w, w_ws, K, C, R * S, s, true);                           // [K][C][RS] -> bf16 hi / lo planes [K][RS][C]
    return tc4_conv_fwd(x, w_ws, w_ws + tc_w_plane_floats(K, C, R, S), bias, y, N, C, H, W, K, R, S, pad, relu, s);
}
// dgrad = forward conv of dY [N, K, P, Q] with the flipped / transposed filters, output [N, C, P, Q] (stride 1, same size)
int tc_conv_dgrad_bf16(const float* dy, const float* w, float* wt_ws, float* dx, int N, int C, int P, int Q, int K, int R, int S,
                       int pad, cudaStream_t s) {
    tc_permute_w_dgrad(w, wt_ws, K, C, R, S, s, true);                         // -> bf16 planes [C][flipped RS][K]
    return tc4_conv_fwd(dy, wt_ws, wt_ws + tc_w_plane_floats(K, C, R, S), nullptr, dx, N, K, P, Q, C, R, S, R - 1 - pad, 0, s);
}
void tc_permute_w_fwd(const float* w, float* w2, int K, int C, int RS, cudaStream_t s, bool bf16) {
    const int ld = (C % 32 == 0) ? RS * C : 32;
    float* lo = (tc_impl() >= 3 || bf16) ? w2 + tc_w_plane_floats(K, C, RS, 1) : nullptr;
    tc::permute_w_fwd_kernel<<<tc::ew_blocks((int64_t)K * ld), 256, 0, s>>>(w, w2, lo, K, C, RS, ld, bf16 ? 1 : 0); clb::count_launch();
}
void tc_permute_w_dgrad(const float* w, float* wd, int K, int C, int R, int S, cudaStream_t s, bool bf16) {
    float* lo = (tc_impl() >= 3 || bf16) ? wd + tc_w_plane_floats(K, C, R, S) : nullptr;
    tc::permute_w_dgrad_kernel<<<tc::ew_blocks((int64_t)K * C * R * S), 256, 0, s>>>(w, wd, lo, K, C, R, S, bf16 ? 1 : 0); clb::count_launch();
}

}  // namespace clb
