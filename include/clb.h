/*
 * clb.h -- C ABI of the B200-native continual-learning hot-path engine ("clb").
 *
 * This is the drop-in boundary of the repo.  The reference (Mattdl/CLsurvey) is
 * pure Python on PyTorch and has no FFI: its operator-level boundary is torch itself
 * (`model(inputs)`, `criterion(outputs, labels)`, `loss.backward()`, `optimizer.step(reg_params)`,
 * src/methods/EWC/train_EWC.py:178-189).  Each entry point below replaces one group of those
 * implicit ATen/cuDNN/cuBLAS calls; the reference call site it replaces is cited per function.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 = CLB_E* on failure; clb_last_error() gives the text
 *   - all tensor pointers are caller-owned DEVICE pointers, fp32 contiguous (labels int64, argmax u8)
 *   - activations NCHW, conv weights [K,C,R,S], linear weights [out,in]  (the reference's layouts)
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and stream-ordered,
 *     allocation-free and CUDA-graph capturable (workspaces are passed in by the caller)
 *   - one host thread per handle/rank; no global mutable state besides the last-error string
 */
#ifndef CLB_H_
#define CLB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLB_OK 0
#define CLB_EINVAL (-1)   /* bad argument / unsupported shape */
#define CLB_ECUDA (-2)    /* CUDA runtime error (text in clb_last_error) */
#define CLB_EWORKSPACE (-3) /* workspace too small */
#define CLB_ENCCL (-4)    /* NCCL error / library not loadable */

/* loss modes of clb_softmax_loss */
#define CLB_LOSS_MEAN_CE 0 /* nn.CrossEntropyLoss()                       main_EWC.py:55 */
#define CLB_LOSS_SUM_NLL 1 /* nll_loss(log_softmax, size_average=False)   main_EWC.py:148 */
#define CLB_LOSS_SUM_SQ 2  /* MSELoss(size_average=False) vs zeros         train_MAS.py:552-560 */

/* GEMM precision modes (clb_set_matmul_mode; process default = CLB_MM_BF16X3, or the CLB_MM_MODE environment variable) */
#define CLB_MM_FP32_SIMT 0 /* exact fp32 FFMA path */
#define CLB_MM_TF32X3 1    /* tcgen05 kind::tf32, 3-pass split (hi*hi + hi*lo + lo*hi), fp32 accum in TMEM */
#define CLB_MM_TF32X1 2    /* tcgen05 kind::tf32 single pass (fast, NOT parity mode) */
#define CLB_MM_BF16X3 3    /* default.  conv fwd/dgrad/wgrad on tcgen05 kind::f16 with a bf16 hi/lo 3-pass split (half the
                              tensor work and operand bytes of TF32X3, same measured error: 3e-6..7e-6 per layer);
                              layers the bf16 kernels do not take (C % 64, 4x4 maps, Linear) run as CLB_MM_TF32X3 */

const char* clb_last_error(void);
int clb_version(void);
int clb_sm_count(int* out);
int clb_set_matmul_mode(int mode);
int clb_get_matmul_mode(void);
unsigned long long clb_launch_count(void); /* kernels launched by this library so far (process-wide) */
/* stream-ordered zero fill (cudaMemsetAsync; a memset node under graph capture): the per-step resets of the loss /
 * #correct accumulators and of the gradient buffer -- optimizer.zero_grad() of train_EWC.py:177 */
int clb_memset_zero(void* p, size_t bytes, void* stream);
/* A new non-blocking stream of the current device, owned by the caller for the life of the process.  The host side wraps
 * these for its side work (H2D staging of the next batch, checkpoint D2H, gradient all-reduce): torch.cuda.Stream() hands out
 * pooled streams round-robin, which after 32 requests alias the stream a CUDA graph is being captured on. */
int clb_stream_create(void** out);

/* ------------------------------------------------------------------------------------------
 * Layer kernels (a2, a3).  Replace model(inputs) / loss.backward() of train_EWC.py:181-187.
 * ---------------------------------------------------------------------------------------- */

/* y = conv2d(x, w) + bias, optional fused ReLU.  nn.Conv2d + nn.ReLU  (models/VGGSlim.py:34-38)
 * w_ws: scratch of 2*max(K*C*R*S, 32*K, 32*C)+8 floats for the tensor-core path's re-ordered weight planes (hi, lo);
 *       NULL forces the fp32 path */
int clb_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, float* w_ws, int N, int C, int H,
                   int W, int K, int R, int S, int stride, int pad, int relu, void* stream);
/* dx = conv2d_backward_input(dy, w).  wt_ws: scratch sized like clb_conv2d_fwd's w_ws (transposed/flipped weight planes). */
int clb_conv2d_dgrad(const float* dy, const float* w, float* dx, float* wt_ws, int N, int C, int H, int W, int K,
                     int R, int S, int stride, int pad, void* stream);
/* dw = conv2d_backward_weight(x, dy), dbias = sum dy.  ws: split-K partials, ws_bytes >= clb_conv2d_wgrad_ws(...) */
size_t clb_conv2d_wgrad_ws(int N, int C, int H, int W, int K, int R, int S, int stride, int pad);
int clb_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, float* ws, size_t ws_bytes, int N,
                     int C, int H, int W, int K, int R, int S, int stride, int pad, void* stream);

/* y[M,out] = x[M,in] @ w[out,in]^T + bias, optional ReLU.  nn.Linear (+ReLU)  (VGGSlim.py:58-73) */
/* ws / ws_bytes: scratch for the tensor-core path (lo planes, transposed copies), >= clb_linear_ws(M,in,out) bytes;
 * NULL / too small selects the exact-fp32 path */
size_t clb_linear_ws(int M, int in, int out);
int clb_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* ws, size_t ws_bytes, int M, int in,
                   int out, int relu, void* stream);
int clb_linear_dgrad(const float* dy, const float* w, float* dx, float* ws, size_t ws_bytes, int M, int in, int out,
                     void* stream);
int clb_linear_wgrad(const float* x, const float* dy, float* dw, float* dbias, float* ws, size_t ws_bytes, int M, int in,
                     int out, void* stream);

/* dx = dy * (y > 0)   (ReLU backward from the saved OUTPUT; in place allowed: dx == dy) */
int clb_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);

/* MaxPool2d(k, stride) fwd (argmax saved as window-local index, first max wins like ATen) / bwd (gather, no atomics).
 * bwd optionally fuses the ReLU mask of the layer in front: dx *= (x_relu_out > 0) when x_relu_out != NULL. */
int clb_maxpool_fwd(const float* x, float* y, uint8_t* argmax, int N, int C, int H, int W, int k, int stride,
                    void* stream);
int clb_maxpool_bwd(const float* dy, const uint8_t* argmax, const float* x_relu_out, float* dx, int N, int C, int H,
                    int W, int k, int stride, void* stream);

/* AdaptiveAvgPool2d((OH,OW))  (torchvision AlexNet.avgpool) */
int clb_adaptive_avgpool_fwd(const float* x, float* y, int N, int C, int H, int W, int OH, int OW, void* stream);
int clb_adaptive_avgpool_bwd(const float* dy, float* dx, int N, int C, int H, int W, int OH, int OW, void* stream);

/* y[r, c] = x[r, c] * mask[(mask_rows == 1 ? 0 : r), c]   dropout with a host-drawn, pre-scaled mask
 * (F.dropout element mask: mask_rows == rows; GEM per-unit mask shared over the batch, gem.py:183-192: mask_rows == 1) */
int clb_mask_mul(const float* x, const float* mask, float* y, int rows, int cols, int mask_rows, void* stream);

/* Loss head on logits[B, ld] restricted to columns [col_off, col_off+ncols)  (GEM head masking, gem.py:199-203,242).
 * mode: CLB_LOSS_*.  dlogits (B x ld, may be NULL) receives d loss / d logits * grad_scale (zeros outside the slice).
 * loss_out[0] += loss contribution (caller zeroes), correct_out[0] += #argmax==label.
 * For MEAN_CE the mean is over `mean_denominator` rows (global batch under data parallelism). */
int clb_softmax_loss(const float* logits, int ld, int col_off, int ncols, const int64_t* labels, int B, int mode,
                     float mean_denominator, float* loss_out, int* correct_out, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------
 * NHWC bf16 hi/lo "planes" pipeline (a2).  The VGG conv stacks (models/VGGSlim.py:27-40: 3x3 / stride 1 / pad 1 convs,
 * ReLU, 2x2 max-pools) keep every activation x as two bf16 planes x = hi + lo in NHWC so that both operands of every
 * conv GEMM are TMA box loads (csrc/clb_planes_conv.cu).  Plane pointers are `void*` to device arrays of uint16 (bf16
 * bit patterns), `[N][H][W][C]`; weights planes are `[K][tap][C]` (forward) and `[C][8 - tap][K]` (dgrad).
 * Same replaced call sites as clb_conv2d_* / clb_maxpool_*: model(inputs) / loss.backward() of train_EWC.py:181-187.
 * ---------------------------------------------------------------------------------------- */

/* 1 when a Conv2d(C -> K, RxS, stride, pad) on an HxW map can run on the planes kernels (3x3/1/1, C % 64 == K % 64 == 0,
 * power-of-two maps 4 <= W <= 64) */
int clb_planes_conv_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
/* w [K][C][3][3] fp32 -> forward planes wf_{hi,lo} [K][9][C] and (when non-NULL) dgrad planes wt_{hi,lo} [C][9][K] */
int clb_planes_weights(const float* w, void* wf_hi, void* wf_lo, void* wt_hi, void* wt_lo, int K, int C, void* stream);
/* the same for n layers in ONE launch; every argument is a HOST array of n entries (device pointers / sizes), n <= 24 */
int clb_planes_weights_batch(int n, const float* const* w, void* const* wf_hi, void* const* wf_lo, void* const* wt_hi, void* const* wt_lo,
                              const int* K, const int* C, const int* taps, void* stream);   /* taps[i]: 9 = conv [K][C][3][3], 1 = nn.Linear
                                                                                               [K][C]; NULL = all 9 */
/* The memory-bound helpers of the planes calls (weight conversion above; split-K reduction and bias gradient of the wgrad calls
 * below) run on a per-device side stream next to the tcgen05 GEMMs.  By default every call joins that stream before it returns.
 * After clb_planes_defer_join(1) those calls return with the join pending and the caller collects it with clb_planes_join(stream)
 * -- after queueing work that depends neither on their results nor on the workspace (the engine: the dgrad GEMM of the same layer,
 * train_EWC.py:171 loss.backward()).  A pending join is also collected by the next call that forks.  Process-wide switch. */
int clb_planes_defer_join(int on);
int clb_planes_join(void* stream);
/* y = conv3x3(x, w) + bias, optional fused ReLU; x planes [N][H][W][C], y planes [N][H][W][K]   (nn.Conv2d + nn.ReLU) */
int clb_planes_conv_fwd(const void* x_hi, const void* x_lo, const void* wf_hi, const void* wf_lo, const float* bias, void* y_hi,
                        void* y_lo, int N, int H, int W, int C, int K, int relu, void* stream);
/* dx = conv2d_backward_input(dy, w); mask_hi (may be NULL): hi plane of this conv's (post-ReLU) INPUT activation -- dx is
 * zeroed where it is <= 0, i.e. the ReLU backward of the layer in front is fused into the epilogue */
int clb_planes_conv_dgrad(const void* dy_hi, const void* dy_lo, const void* wt_hi, const void* wt_lo, const void* mask_hi, void* dx_hi,
                          void* dx_lo, int N, int H, int W, int C, int K, void* stream);
/* dw [K][C][3][3] fp32 = conv2d_backward_weight(x, dy), dbias = sum dy (may be NULL); deterministic split-K.
 * imp_mode != 0 fuses the importance update of the batch gradient into the split-K reduction (north_star "fused into the
 * backward pass"):  1: omega += dw*dw / imp_a   (diag_fisher, EWC/main_EWC.py:151-156)
 *                   2: omega = (omega*imp_a + |dw|) / imp_b   (Objective_After_SGD.step, MAS/train_MAS.py:163-177)
 * omega: the [K][C][3][3] slot of the flat importance buffer. */
size_t clb_planes_conv_wgrad_ws(int N, int H, int W, int C, int K);
int clb_planes_conv_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw, float* dbias, float* ws,
                          size_t ws_bytes, int N, int H, int W, int C, int K, int imp_mode, float* omega, float imp_a, float imp_b,
                          void* stream);
/* First layer of the VGG stacks fused with its ReLU and 2x2 max-pool (features[0..2] of VGGSlim.py:27-40: Conv2d(3, 64, 3,
 * padding=1), ReLU, MaxPool2d(2, 2)): x fp32 NCHW [N][3][H][W] -> planes [N][H/2][W/2][64] + argmax, exact fp32 FFMA.
 * The backward call fuses max-pool backward, ReLU backward and the weight / bias gradient (dp = gradient w.r.t. the pooled
 * output, pooled_hi = hi plane of the pooled output).  ws_bytes >= clb_planes_conv1_ws(). */
int clb_planes_conv1_supported(int C, int H, int W, int K, int R, int S, int stride, int pad);
int clb_planes_conv1_pool_fwd(const float* x, const float* w, const float* bias, void* y_hi, void* y_lo, uint8_t* argmax, int N, int C,
                              int H, int W, int K, void* stream);
size_t clb_planes_conv1_ws(void);
int clb_planes_conv1_pool_bwd(const float* x, const void* dp_hi, const void* dp_lo, const void* pooled_hi, const uint8_t* argmax, float* dw,
                              float* dbias, float* ws, size_t ws_bytes, int N, int C, int H, int W, int K, void* stream);
/* MaxPool2d(2, 2) on planes (first maximum wins, like ATen).  Output: planes [N][H/2][W/2][C], or -- y_f32 != NULL -- fp32
 * [N][C][H/2][W/2] (the classifier's flatten order).  argmax: [N][H/2][W/2][C] window-local index. */
int clb_planes_pool_fwd(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, float* y_f32, uint8_t* argmax, int N, int H, int W,
                        int C, void* stream);
/* same, from an fp32 NCHW input (output of a conv that does not run on the planes kernels, e.g. the C = 3 first layer) */
int clb_planes_pool_fwd_nchw(const float* x, void* y_hi, void* y_lo, uint8_t* argmax, int N, int C, int H, int W, void* stream);
/* max-pool backward fused with the ReLU backward of the conv in front: dx[window] = (argmax && pooled > 0) ? dy : 0.
 * dy / pooled activation either as planes or (dy_f32, pooled_f32 != NULL) as fp32 [N][C][H/2][W/2]; H, W = INPUT size */
int clb_planes_pool_bwd(const void* dy_hi, const void* dy_lo, const float* dy_f32, const void* pooled_hi, const float* pooled_f32,
                        const uint8_t* argmax, void* dx_hi, void* dx_lo, int N, int H, int W, int C, void* stream);
/* same with an fp32 NCHW result [N][C][H][W] (dY of a conv that does not run on the planes kernels) */
int clb_planes_pool_bwd_nchw(const void* dy_hi, const void* dy_lo, const void* pooled_hi, const uint8_t* argmax, float* dx, int N, int C,
                             int H, int W, void* stream);

/* nn.Linear (+ReLU) of the VGG classifiers (VGGSlim.py:58-73) on the same kernels: a 1x1 "conv" over a 1x1 map whose rows
 * are the M samples.  in % 64 == out % 64 == 0.  x planes [M][in], y planes [M][out], weight planes from
 * clb_planes_weights_batch with taps = 1 ([out][in] forward, [in][out] dgrad).  mask_hi / imp_* as for the convs. */
int clb_planes_linear_supported(int in, int out);
int clb_planes_linear_fwd(const void* x_hi, const void* x_lo, const void* wf_hi, const void* wf_lo, const float* bias, void* y_hi,
                          void* y_lo, int M, int in, int out, int relu, void* stream);
int clb_planes_linear_dgrad(const void* dy_hi, const void* dy_lo, const void* wt_hi, const void* wt_lo, const void* mask_hi, void* dx_hi,
                            void* dx_lo, int M, int in, int out, void* stream);
size_t clb_planes_linear_wgrad_ws(int M, int in, int out);
int clb_planes_linear_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw, float* dbias, float* ws,
                            size_t ws_bytes, int M, int in, int out, int imp_mode, float* omega, float imp_a, float imp_b, void* stream);
/* the conv / classifier boundary: max-pool whose output (backward: whose incoming gradient and pooled activation) are planes
 * in the classifier's flatten order [N][C][H/2][W/2]  (x.view(x.size(0), -1), VGGSlim.py:66) */
int clb_planes_pool_fwd_flat(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, uint8_t* argmax, int N, int H, int W, int C,
                             void* stream);
int clb_planes_pool_bwd_flat(const void* dy_hi, const void* dy_lo, const void* pooled_hi, const uint8_t* argmax, void* dx_hi, void* dx_lo,
                             int N, int H, int W, int C, void* stream);
/* element-wise layout changes at the edges of the planes pipeline: out = hi + lo;  (hi, lo) = split(x), zeroed where the
 * activation whose hi plane is mask_hi (may be NULL) is <= 0 (ReLU backward) */
int clb_planes_to_f32(const void* hi, const void* lo, float* out, int64_t n, void* stream);
int clb_planes_from_f32(const float* x, const void* mask_hi, void* hi, void* lo, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser / importance streaming kernels (a4-a11).  One launch over the flat parameter buffer.
 * ---------------------------------------------------------------------------------------- */

/* Weight_Regularized_SGD.step (EWC/train_EWC.py:46-84, MAS/train_MAS.py:45-93) and plain optim.SGD (main_SGD.py:74):
 *   d = g*grad_scale [+ (theta-theta*) * (omega * two_lambda) for i < n_penalised] [+ wd*theta]
 *   buf = first_step ? d : mu*buf + d ;  theta -= lr*buf
 * omega/theta_star may be NULL when n_penalised == 0. */
int clb_sgd_penalty_step(float* theta, const float* g, const float* omega, const float* theta_star, float* buf,
                         int64_t n, int64_t n_penalised, float two_lambda, float lr, float momentum,
                         float weight_decay, float grad_scale, int first_step, void* stream);

/* Elastic_SGD.step (SI/train_SI.py:48-125): as above for ALL n elements plus  w += -(theta_new - theta_old) * g0 */
int clb_si_step(float* theta, const float* g, const float* omega, const float* theta_star, float* buf, float* w,
                int64_t n, float two_lambda, float lr, float momentum, float weight_decay, float grad_scale,
                int first_step, void* stream);

/* diag_fisher accumulate (EWC/main_EWC.py:151-156):  omega += g*g / data_len */
int clb_fisher_accum(float* omega, const float* g, float data_len, int64_t n, void* stream);

/* Objective_After_SGD.step (MAS/train_MAS.py:163-177):  omega = (omega*prev_size + |g|) / curr_size */
int clb_mas_accum(float* omega, const float* g, float prev_size, float curr_size, int64_t n, void* stream);

/* update_reg_params (SI/train_SI.py:390-417):  omega += max(w / ((theta-theta*)^2 + slack), 0); w = 0; theta* = theta */
int clb_si_consolidate(float* omega, float* w, const float* theta, float* theta_star, float slack, int64_t n,
                       void* stream);

/* mode-IMM merge (IMM/merge.py:228-231): acc (+)= (prec / sum_prec) * theta, element-wise, each op rounded like the
 * reference's tensor ops; first != 0 overwrites acc (the reference starts from torch.zeros).  No alignment requirement. */
int clb_imm_merge_accum(float* acc, const float* prec, const float* sum_prec, const float* theta, int64_t n, int first,
                        void* stream);

/* accumelate_reg_params (EWC/main_EWC.py:205-232):  dst = a + b*scale_b   (also used to finish a sharded pass) */
int clb_axpby(float* dst, const float* a, const float* b, float scale_b, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * GEM (a15, a16).  Gradient memory is TASK-MAJOR G[n_tasks][ld] (reference: [P, n_tasks], gem.py:131-132).
 * ---------------------------------------------------------------------------------------- */

/* dots[i] = <g, G[idx[i]]>, gram[i][j] = <G[idx[i]], G[idx[j]]>  (fp64 accumulation of exact fp32 products);
 * replaces torch.mm(grads[:,t], grads.index_select(1,indx)) (gem.py:275-276) and the host MM^T of gem.py:73.
 * dots/gram must be zeroed by the caller (they are accumulated with one atomic per CTA). k <= 16. */
int clb_gem_dots_gram(const float* g, const float* G, int64_t ld, int64_t P, const int* idx_dev, int k, double* dots,
                      double* gram, void* stream);

/* Device QP of project2cone2 (gem.py:73-78): P = 0.5(gram+gram^T)+eps I, q = -dots, min 1/2 v'Pv + dots'v s.t. v >= margin.
 * Writes v[k] and viol[0] = #(dots < 0).  If viol == 0, v is set to 0 (no projection). */
int clb_gem_solve_qp(const double* dots, const double* gram, int k, double margin, double eps, double* v, int* viol,
                     void* stream);
/* same solver on the host (tests / INTEGRATION.md); pointers are HOST pointers */
int clb_gem_solve_qp_host(const double* dots, const double* gram, int k, double margin, double eps, double* v,
                          int* viol);

/* g[i] = float( sum_j v[j]*G[idx[j]][i] + g[i] )  computed in fp64 when viol[0] != 0 (gem.py:79-80) */
int clb_gem_project(float* g, const float* G, int64_t ld, int64_t P, const int* idx_dev, int k, const double* v,
                    const int* viol, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data parallel (8e): thin wrapper over ncclAllReduce(sum, fp32) on a communicator created here.
 * ---------------------------------------------------------------------------------------- */
int clb_nccl_unique_id(void* out128);                              /* rank 0: 128-byte ncclUniqueId */
int clb_nccl_init(const void* id128, int rank, int world, void** comm_out);
int clb_nccl_allreduce_f32(void* comm, float* buf, int64_t n, void* stream);
int clb_nccl_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* CLB_H_ */
