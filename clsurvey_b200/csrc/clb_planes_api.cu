// C ABI of the "planes" pipeline (include/clb.h, section "NHWC bf16 hi/lo planes").
#include "clb_planes.cuh"

namespace clb {
namespace pl {
int weights_to_planes(const float* w, uint16_t* wf_hi, uint16_t* wf_lo, uint16_t* wt_hi, uint16_t* wt_lo, int K, int C, int taps,
                      cudaStream_t s);
int weights_to_planes_batch(int n, const float* const* w, void* const* wf_hi, void* const* wf_lo, void* const* wt_hi, void* const* wt_lo,
                            const int* K, const int* C, const int* taps, cudaStream_t s);
int pool_fwd(const uint16_t* x_hi, const uint16_t* x_lo, uint16_t* y_hi, uint16_t* y_lo, float* y_f32, uint8_t* am, int N, int H, int W,
             int C, int flat, cudaStream_t s);
int planes_to_f32(const uint16_t* hi, const uint16_t* lo, float* out, int64_t n, cudaStream_t s);
int f32_to_planes(const float* x, const uint16_t* mask_hi, uint16_t* hi, uint16_t* lo, int64_t n, cudaStream_t s);
int pool_fwd_from_nchw(const float* x, uint16_t* y_hi, uint16_t* y_lo, uint8_t* am, int N, int C, int H, int W, cudaStream_t s);
int pool_bwd(const uint16_t* dy_hi, const uint16_t* dy_lo, const float* dy_f32, const uint16_t* pooled_hi, const float* pooled_f32,
             const uint8_t* am, uint16_t* dx_hi, uint16_t* dx_lo, int N, int H, int W, int C, int flat, cudaStream_t s);
int pool_bwd_to_nchw(const uint16_t* dy_hi, const uint16_t* dy_lo, const uint16_t* pooled_hi, const uint8_t* am, float* dx, int N, int C,
                     int H, int W, cudaStream_t s);
size_t bias_ws_floats(int K);
int bias_grad(const uint16_t* dy_hi, const uint16_t* dy_lo, float* db, float* part, int64_t npix, int K, cudaStream_t s);
int wgrad_reduce(const float* ws, float* dw, float* omega, int K, int C, int taps, int splits, int imp_mode, float imp_a, float imp_b,
                 cudaStream_t s);
}  // namespace pl
}  // namespace clb

using namespace clb;
typedef uint16_t u16;

extern "C" {

int clb_planes_conv_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    return pl::conv_supported(C, H, W, K, R, S, stride, pad) ? 1 : 0;
}

int clb_planes_weights(const float* w, void* wf_hi, void* wf_lo, void* wt_hi, void* wt_lo, int K, int C, void* stream) {
    CLB_CHECK_ARG(w && wf_hi && wf_lo && K > 0 && C > 0 && ((wt_hi == nullptr) == (wt_lo == nullptr)));
    int rc = pl::weights_to_planes(w, (u16*)wf_hi, (u16*)wf_lo, (u16*)wt_hi, (u16*)wt_lo, K, C, 9, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_weights_batch(int n, const float* const* w, void* const* wf_hi, void* const* wf_lo, void* const* wt_hi, void* const* wt_lo,
                              const int* K, const int* C, const int* taps, void* stream) {
    CLB_CHECK_ARG(n > 0 && n <= 24 && w && wf_hi && wf_lo && wt_hi && wt_lo && K && C);
    for (int i = 0; i < n; ++i) {
        CLB_CHECK_ARG(w[i] && wf_hi[i] && wf_lo[i] && wt_hi[i] && wt_lo[i] && K[i] > 0 && C[i] > 0 && (K[i] % 2) == 0 && (C[i] % 2) == 0);
        CLB_CHECK_ARG(taps == nullptr || taps[i] == 1 || taps[i] == 9);
    }
    int rc = pl::weights_to_planes_batch(n, w, wf_hi, wf_lo, wt_hi, wt_lo, K, C, taps, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_conv_fwd(const void* x_hi, const void* x_lo, const void* wf_hi, const void* wf_lo, const float* bias, void* y_hi,
                        void* y_lo, int N, int H, int W, int C, int K, int relu, void* stream) {
    CLB_CHECK_ARG(x_hi && x_lo && wf_hi && wf_lo && y_hi && y_lo && N > 0);
    CLB_CHECK_ARG(pl::conv_supported(C, H, W, K, 3, 3, 1, 1));
    int rc = pl::conv_fwd((const u16*)x_hi, (const u16*)x_lo, (const u16*)wf_hi, (const u16*)wf_lo, bias, relu, nullptr, (u16*)y_hi,
                          (u16*)y_lo, N, H, W, C, K, 9, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_conv_dgrad(const void* dy_hi, const void* dy_lo, const void* wt_hi, const void* wt_lo, const void* mask_hi, void* dx_hi,
                          void* dx_lo, int N, int H, int W, int C, int K, void* stream) {
    CLB_CHECK_ARG(dy_hi && dy_lo && wt_hi && wt_lo && dx_hi && dx_lo && N > 0);
    CLB_CHECK_ARG(pl::conv_supported(C, H, W, K, 3, 3, 1, 1));
    // forward conv of dY [N,H,W,K] with the flipped / transposed filters [C][9][K]: reduction over K, output channels C
    int rc = pl::conv_fwd((const u16*)dy_hi, (const u16*)dy_lo, (const u16*)wt_hi, (const u16*)wt_lo, nullptr, 0, (const u16*)mask_hi,
                          (u16*)dx_hi, (u16*)dx_lo, N, H, W, K, C, 9, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

size_t clb_planes_conv_wgrad_ws(int N, int H, int W, int C, int K) {
    return (pl::wgrad_ws_floats(N, H, W, C, K, 9) + pl::bias_ws_floats(K) + 8) * sizeof(float);
}

int clb_planes_conv_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw, float* dbias, float* ws,
                          size_t ws_bytes, int N, int H, int W, int C, int K, int imp_mode, float* omega, float imp_a, float imp_b,
                          void* stream) {
    CLB_CHECK_ARG(x_hi && x_lo && dy_hi && dy_lo && dw && ws && N > 0);
    CLB_CHECK_ARG(pl::conv_supported(C, H, W, K, 3, 3, 1, 1));
    CLB_CHECK_ARG(imp_mode >= 0 && imp_mode <= 2 && (imp_mode == 0 || omega != nullptr));
    if (ws_bytes < clb_planes_conv_wgrad_ws(N, H, W, C, K)) {
        set_error("clb_planes_conv_wgrad: workspace %zu bytes < required %zu", ws_bytes, clb_planes_conv_wgrad_ws(N, H, W, C, K));
        return CLB_EWORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    int splits = 0;
    int rc = pl::conv_wgrad_partials((const u16*)x_hi, (const u16*)x_lo, (const u16*)dy_hi, (const u16*)dy_lo, ws, &splits, N, H, W, C, K, 9, s);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    rc = pl::wgrad_reduce(ws, dw, omega, K, C, 9, splits, imp_mode, imp_a, imp_b, s);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    if (dbias) {
        float* part = ws + ((pl::wgrad_ws_floats(N, H, W, C, K, 9) + 3) & ~(size_t)3);
        rc = pl::bias_grad((const u16*)dy_hi, (const u16*)dy_lo, dbias, part, (int64_t)N * H * W, K, s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
    }
    return CLB_OK;
}

int clb_planes_pool_fwd(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, float* y_f32, uint8_t* argmax, int N, int H, int W,
                        int C, void* stream) {
    CLB_CHECK_ARG(x_hi && x_lo && argmax && ((y_hi && y_lo) || y_f32) && N > 0 && (H % 2) == 0 && (W % 2) == 0 && (C % 8) == 0);
    int rc = pl::pool_fwd((const u16*)x_hi, (const u16*)x_lo, (u16*)y_hi, (u16*)y_lo, y_f32, argmax, N, H, W, C, 0, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_pool_fwd_nchw(const float* x, void* y_hi, void* y_lo, uint8_t* argmax, int N, int C, int H, int W, void* stream) {
    CLB_CHECK_ARG(x && y_hi && y_lo && argmax && N > 0 && (H % 2) == 0 && (W % 2) == 0 && (C % 64) == 0 && W <= 256);
    int rc = pl::pool_fwd_from_nchw(x, (u16*)y_hi, (u16*)y_lo, argmax, N, C, H, W, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_pool_bwd(const void* dy_hi, const void* dy_lo, const float* dy_f32, const void* pooled_hi, const float* pooled_f32,
                        const uint8_t* argmax, void* dx_hi, void* dx_lo, int N, int H, int W, int C, void* stream) {
    CLB_CHECK_ARG(argmax && dx_hi && dx_lo && N > 0 && (H % 2) == 0 && (W % 2) == 0 && (C % 8) == 0);
    CLB_CHECK_ARG((dy_f32 && pooled_f32) || (dy_hi && dy_lo && pooled_hi));
    int rc = pl::pool_bwd((const u16*)dy_hi, (const u16*)dy_lo, dy_f32, (const u16*)pooled_hi, pooled_f32, argmax, (u16*)dx_hi, (u16*)dx_lo,
                          N, H, W, C, 0, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_pool_bwd_nchw(const void* dy_hi, const void* dy_lo, const void* pooled_hi, const uint8_t* argmax, float* dx, int N, int C,
                             int H, int W, void* stream) {
    CLB_CHECK_ARG(dy_hi && dy_lo && pooled_hi && argmax && dx && N > 0 && (H % 2) == 0 && (W % 2) == 0 && (C % 64) == 0 && W <= 128);
    int rc = pl::pool_bwd_to_nchw((const u16*)dy_hi, (const u16*)dy_lo, (const u16*)pooled_hi, argmax, dx, N, C, H, W, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

/* ---- nn.Linear on the planes kernels: a 1x1 "conv" over a 1x1 map, rows = samples ------------------------------------ */
int clb_planes_linear_supported(int in, int out) { return pl::linear_supported(in, out) ? 1 : 0; }

int clb_planes_linear_fwd(const void* x_hi, const void* x_lo, const void* wf_hi, const void* wf_lo, const float* bias, void* y_hi,
                          void* y_lo, int M, int in, int out, int relu, void* stream) {
    CLB_CHECK_ARG(x_hi && x_lo && wf_hi && wf_lo && y_hi && y_lo && M > 0 && pl::linear_supported(in, out));
    int rc = pl::conv_fwd((const u16*)x_hi, (const u16*)x_lo, (const u16*)wf_hi, (const u16*)wf_lo, bias, relu, nullptr, (u16*)y_hi,
                          (u16*)y_lo, M, 1, 1, in, out, 1, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_linear_dgrad(const void* dy_hi, const void* dy_lo, const void* wt_hi, const void* wt_lo, const void* mask_hi, void* dx_hi,
                            void* dx_lo, int M, int in, int out, void* stream) {
    CLB_CHECK_ARG(dy_hi && dy_lo && wt_hi && wt_lo && dx_hi && dx_lo && M > 0 && pl::linear_supported(in, out));
    int rc = pl::conv_fwd((const u16*)dy_hi, (const u16*)dy_lo, (const u16*)wt_hi, (const u16*)wt_lo, nullptr, 0, (const u16*)mask_hi,
                          (u16*)dx_hi, (u16*)dx_lo, M, 1, 1, out, in, 1, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

size_t clb_planes_linear_wgrad_ws(int M, int in, int out) {
    return (pl::wgrad_ws_floats(M, 1, 1, in, out, 1) + pl::bias_ws_floats(out) + 8) * sizeof(float);
}

int clb_planes_linear_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw, float* dbias, float* ws,
                            size_t ws_bytes, int M, int in, int out, int imp_mode, float* omega, float imp_a, float imp_b, void* stream) {
    CLB_CHECK_ARG(x_hi && x_lo && dy_hi && dy_lo && dw && ws && M > 0 && pl::linear_supported(in, out));
    CLB_CHECK_ARG(imp_mode >= 0 && imp_mode <= 2 && (imp_mode == 0 || omega != nullptr));
    if (ws_bytes < clb_planes_linear_wgrad_ws(M, in, out)) { set_error("clb_planes_linear_wgrad: workspace too small"); return CLB_EWORKSPACE; }
    cudaStream_t s = as_stream(stream);
    int splits = 0;
    int rc = pl::conv_wgrad_partials((const u16*)x_hi, (const u16*)x_lo, (const u16*)dy_hi, (const u16*)dy_lo, ws, &splits, M, 1, 1, in, out, 1, s);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    rc = pl::wgrad_reduce(ws, dw, omega, out, in, 1, splits, imp_mode, imp_a, imp_b, s);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    if (dbias) {
        float* part = ws + ((pl::wgrad_ws_floats(M, 1, 1, in, out, 1) + 3) & ~(size_t)3);
        rc = pl::bias_grad((const u16*)dy_hi, (const u16*)dy_lo, dbias, part, (int64_t)M, out, s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
    }
    return CLB_OK;
}

/* planes in the classifier's flatten order [N][C][H/2][W/2] at the conv / classifier boundary */
int clb_planes_pool_fwd_flat(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, uint8_t* argmax, int N, int H, int W, int C,
                             void* stream) {
    CLB_CHECK_ARG(x_hi && x_lo && y_hi && y_lo && argmax && N > 0 && (H % 2) == 0 && (W % 2) == 0 && (C % 8) == 0);
    int rc = pl::pool_fwd((const u16*)x_hi, (const u16*)x_lo, (u16*)y_hi, (u16*)y_lo, nullptr, argmax, N, H, W, C, 1, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}
int clb_planes_pool_bwd_flat(const void* dy_hi, const void* dy_lo, const void* pooled_hi, const uint8_t* argmax, void* dx_hi, void* dx_lo,
                             int N, int H, int W, int C, void* stream) {
    CLB_CHECK_ARG(dy_hi && dy_lo && pooled_hi && argmax && dx_hi && dx_lo && N > 0 && (H % 2) == 0 && (W % 2) == 0 && (C % 8) == 0);
    int rc = pl::pool_bwd((const u16*)dy_hi, (const u16*)dy_lo, nullptr, (const u16*)pooled_hi, nullptr, argmax, (u16*)dx_hi, (u16*)dx_lo, N,
                          H, W, C, 1, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}
int clb_planes_to_f32(const void* hi, const void* lo, float* out, int64_t n, void* stream) {
    CLB_CHECK_ARG(hi && lo && out && n >= 0);
    if (n == 0) return CLB_OK;
    int rc = pl::planes_to_f32((const u16*)hi, (const u16*)lo, out, n, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}
int clb_planes_from_f32(const float* x, const void* mask_hi, void* hi, void* lo, int64_t n, void* stream) {
    CLB_CHECK_ARG(x && hi && lo && n >= 0);
    if (n == 0) return CLB_OK;
    int rc = pl::f32_to_planes(x, (const u16*)mask_hi, (u16*)hi, (u16*)lo, n, as_stream(stream));
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

}  // extern "C"
