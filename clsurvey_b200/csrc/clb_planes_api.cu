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

// Side stream.  The memory-bound helpers of a layer -- the bias gradient (column sums of dY), the split-K reduction of dW and the
// weight -> planes conversion -- need a few KB of shared memory and <= 40 registers per thread, so their CTAs fit on the SMs next
// to the one persistent CTA of a tcgen05 GEMM (shared-memory / tensor bound).  They are forked off the caller's stream onto one
// side stream per device and joined back
//   - before the call returns (default): callers, and a CUDA graph being captured on their stream, see one ordered call;
//   - or, after clb_planes_defer_join(1), at the caller's next clb_planes_join(stream): the engine issues the dgrad GEMM of the
//     same layer (resp. the fp32 first-layer kernel) in between, which touches neither the workspace nor dW.
// CLB_BIAS_SIDE=0 keeps everything on the caller's stream.
namespace {
struct Side { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, fork2 = nullptr, join = nullptr; int pending = 0; };
int g_defer_join = 0;
Side* side_of_current_device() {
    static Side sides[32];
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("CLB_BIAS_SIDE"); enabled = (e && e[0] == '0') ? 0 : 1; }
    int dev = 0;
    if (!enabled || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
    Side& sd = sides[dev];
    if (!sd.st) {
        if (cudaStreamCreateWithFlags(&sd.st, cudaStreamNonBlocking) != cudaSuccess) { sd.st = nullptr; return nullptr; }
        cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&sd.fork2, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming);
    }
    return &sd;
}
void side_join_pending(Side* sd, cudaStream_t s) {      // a deferred join nobody collected yet: collect it now
    if (sd && sd->pending) { cudaStreamWaitEvent(s, sd->join, 0); sd->pending = 0; }
}
Side* side_fork(cudaStream_t s) {                       // the side stream now waits for everything queued on s so far
    Side* sd = side_of_current_device();
    side_join_pending(sd, s);
    if (sd && cudaEventRecord(sd->fork, s) == cudaSuccess && cudaStreamWaitEvent(sd->st, sd->fork, 0) == cudaSuccess) return sd;
    return nullptr;
}
void side_fork_again(Side* sd, cudaStream_t s) {        // ... and for what was queued on s since
    if (sd && cudaEventRecord(sd->fork2, s) == cudaSuccess) cudaStreamWaitEvent(sd->st, sd->fork2, 0);
}
void side_join(Side* sd, cudaStream_t s, bool may_defer) {
    if (!sd || cudaEventRecord(sd->join, sd->st) != cudaSuccess) return;
    if (may_defer && g_defer_join) sd->pending = 1;
    else cudaStreamWaitEvent(s, sd->join, 0);
}
// dW = sum of the split-K partials (+ importance), db = column sums of dY: on the side stream when there is one
int wgrad_tail(Side* sd, cudaStream_t s, const float* ws, float* dw, float* omega, int K, int C, int taps, int splits, int imp_mode,
               float imp_a, float imp_b, const u16* dy_hi, const u16* dy_lo, float* dbias, float* part, int64_t rows) {
    cudaStream_t t = sd ? sd->st : s;
    int rc = CLB_OK;
    if (dbias) rc = pl::bias_grad(dy_hi, dy_lo, dbias, part, rows, K, t);       // independent of the GEMM: runs next to it
    side_fork_again(sd, s);                                                      // the reduction needs the partials
    if (!rc) rc = pl::wgrad_reduce(ws, dw, omega, K, C, taps, splits, imp_mode, imp_a, imp_b, t);
    side_join(sd, s, true);
    return rc;
}
}  // namespace

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
    cudaStream_t s = as_stream(stream);
    Side* sd = g_defer_join ? side_fork(s) : nullptr;   // worth a fork only when the caller has work to put next to it
    int rc = pl::weights_to_planes_batch(n, w, wf_hi, wf_lo, wt_hi, wt_lo, K, C, taps, sd ? sd->st : s);
    side_join(sd, s, true);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_planes_defer_join(int on) {
    g_defer_join = on ? 1 : 0;
    return CLB_OK;
}

int clb_planes_join(void* stream) {
    side_join_pending(side_of_current_device(), as_stream(stream));
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
    Side* sd = side_fork(s);
    int rc = pl::conv_wgrad_partials((const u16*)x_hi, (const u16*)x_lo, (const u16*)dy_hi, (const u16*)dy_lo, ws, &splits, N, H, W, C, K, 9, s);
    float* part = ws + ((pl::wgrad_ws_floats(N, H, W, C, K, 9) + 3) & ~(size_t)3);
    if (rc) side_join(sd, s, false);
    else rc = wgrad_tail(sd, s, ws, dw, omega, K, C, 9, splits, imp_mode, imp_a, imp_b, (const u16*)dy_hi, (const u16*)dy_lo, dbias, part,
                         (int64_t)N * H * W);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
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
    Side* sd = side_fork(s);
    int rc = pl::conv_wgrad_partials((const u16*)x_hi, (const u16*)x_lo, (const u16*)dy_hi, (const u16*)dy_lo, ws, &splits, M, 1, 1, in, out, 1, s);
    float* part = ws + ((pl::wgrad_ws_floats(M, 1, 1, in, out, 1) + 3) & ~(size_t)3);
    if (rc) side_join(sd, s, false);
    else rc = wgrad_tail(sd, s, ws, dw, omega, out, in, 1, splits, imp_mode, imp_a, imp_b, (const u16*)dy_hi, (const u16*)dy_lo, dbias, part,
                         (int64_t)M);
    if (rc) return rc;
    CLB_CHECK_LAUNCH();
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
