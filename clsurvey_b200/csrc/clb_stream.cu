// HBM-streaming kernels over the flat parameter buffers: penalised SGD, SI step, Fisher / MAS accumulators,
// SI consolidation.  One launch per buffer, 128-bit accesses, grid = one full wave of 148 SMs x 8 CTAs.
//
// Arithmetic follows the reference op-by-op (each torch elementwise op is its own rounding step; the
// `a + alpha*b` forms are fused multiply-adds exactly like ATen's add kernel):
//   Weight_Regularized_SGD.step  src/methods/EWC/train_EWC.py:46-84, src/methods/MAS/train_MAS.py:45-93
//   Elastic_SGD.step             src/methods/SI/train_SI.py:48-125
//   diag_fisher                  src/methods/EWC/main_EWC.py:151-156
//   Objective_After_SGD.step     src/methods/MAS/train_MAS.py:163-177
//   update_reg_params            src/methods/SI/train_SI.py:390-417
#include "clb_common.cuh"

namespace clb {

constexpr int kThreads = 256;

static inline int stream_grid(int64_t n_vec) {
    int64_t blocks = (n_vec + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count() * 8;  // 8 x 256 threads = 2048 resident threads per SM: exactly one wave
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ float4 ld4_stream(const float* p, int64_t i) {
    return __ldcs(reinterpret_cast<const float4*>(p) + i);  // read-once data: evict-first
}
__device__ __forceinline__ void st4(float* p, int64_t i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }

// ---- penalised SGD ------------------------------------------------------------------------------
struct SgdArgs {
    float two_lambda, lr, mu, wd, gscale;
    int first;
};

__device__ __forceinline__ void sgd_elem(float& th, float g, float om, float ts, float& bf, bool pen, const SgdArgs& a) {
    float d = (a.gscale == 1.0f) ? g : __fmul_rn(g, a.gscale);
    if (pen) {
        float diff = __fsub_rn(th, ts);                // curr.add(-1, init_val)
        float t2 = __fmul_rn(om, a.two_lambda);        // (2*lambda) * omega
        d = __fadd_rn(d, __fmul_rn(diff, t2));         // d_p.add_(weight_dif.mul(...))
    }
    if (a.wd != 0.0f) d = __fmaf_rn(a.wd, th, d);       // d_p.add_(weight_decay, p.data)
    float b = a.first ? d : __fadd_rn(__fmul_rn(bf, a.mu), d);  // buf.mul_(mu).add_(d)
    bf = b;
    th = __fmaf_rn(-a.lr, b, th);                      // p.data.add_(-lr, buf)
}

__global__ void __launch_bounds__(kThreads)
sgd_penalty_kernel(float* __restrict__ theta, const float* __restrict__ g, const float* __restrict__ omega,
                   const float* __restrict__ tstar, float* __restrict__ buf, int64_t n, int64_t n_pen, SgdArgs a) {
    pdl_trigger();
    pdl_wait();
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 th = ld4(theta, i);
        float4 gg = ld4_stream(g, i);
        float4 bf = a.first ? make_float4(0, 0, 0, 0) : ld4(buf, i);
        float4 om = make_float4(0, 0, 0, 0), ts = make_float4(0, 0, 0, 0);
        const bool any_pen = e < n_pen;
        if (any_pen) {
            om = ld4(omega, i);
            ts = ld4(tstar, i);
        }
        sgd_elem(th.x, gg.x, om.x, ts.x, bf.x, e + 0 < n_pen, a);
        sgd_elem(th.y, gg.y, om.y, ts.y, bf.y, e + 1 < n_pen, a);
        sgd_elem(th.z, gg.z, om.z, ts.z, bf.z, e + 2 < n_pen, a);
        sgd_elem(th.w, gg.w, om.w, ts.w, bf.w, e + 3 < n_pen, a);
        st4(theta, i, th);
        st4(buf, i, bf);
    }
    // scalar tail (n % 4 elements), handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        float th = theta[e], bf = a.first ? 0.f : buf[e];
        const bool pen = e < n_pen;
        sgd_elem(th, g[e], pen ? omega[e] : 0.f, pen ? tstar[e] : 0.f, bf, pen, a);
        theta[e] = th;
        buf[e] = bf;
    }
}

// ---- SI step ------------------------------------------------------------------------------------
__device__ __forceinline__ void si_elem(float& th, float g, float om, float ts, float& bf, float& w, const SgdArgs& a) {
    const float g0 = (a.gscale == 1.0f) ? g : __fmul_rn(g, a.gscale);   // unreg_dp
    const float th_old = th;                                             // curr_wegiht_val (clone)
    float d = __fadd_rn(g0, __fmul_rn(__fsub_rn(th, ts), __fmul_rn(om, a.two_lambda)));
    if (a.wd != 0.0f) d = __fmaf_rn(a.wd, th, d);
    float b = a.first ? d : __fadd_rn(__fmul_rn(bf, a.mu), d);
    bf = b;
    th = __fmaf_rn(-a.lr, b, th);
    const float w_diff = __fsub_rn(th, th_old);                          // p.data.add(-1, curr)
    const float change = __fmul_rn(__fmul_rn(w_diff, g0), -1.0f);        // w_diff.mul(unreg_dp) * -1
    w = __fadd_rn(w, change);
}

__global__ void __launch_bounds__(kThreads)
si_step_kernel(float* __restrict__ theta, const float* __restrict__ g, const float* __restrict__ omega,
               const float* __restrict__ tstar, float* __restrict__ buf, float* __restrict__ w, int64_t n, SgdArgs a) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 th = ld4(theta, i), gg = ld4_stream(g, i), om = ld4(omega, i), ts = ld4(tstar, i), ww = ld4(w, i);
        float4 bf = a.first ? make_float4(0, 0, 0, 0) : ld4(buf, i);
        si_elem(th.x, gg.x, om.x, ts.x, bf.x, ww.x, a);
        si_elem(th.y, gg.y, om.y, ts.y, bf.y, ww.y, a);
        si_elem(th.z, gg.z, om.z, ts.z, bf.z, ww.z, a);
        si_elem(th.w, gg.w, om.w, ts.w, bf.w, ww.w, a);
        st4(theta, i, th);
        st4(buf, i, bf);
        st4(w, i, ww);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        float th = theta[e], bf = a.first ? 0.f : buf[e], ww = w[e];
        si_elem(th, g[e], omega[e], tstar[e], bf, ww, a);
        theta[e] = th;
        buf[e] = bf;
        w[e] = ww;
    }
}

// ---- importance accumulators --------------------------------------------------------------------
__device__ __forceinline__ float fisher_elem(float om, float g, float len) {
    return __fadd_rn(om, __fdiv_rn(__fmul_rn(g, g), len));  // omega += grad**2 / data_len
}
__global__ void __launch_bounds__(kThreads)
fisher_kernel(float* __restrict__ omega, const float* __restrict__ g, float len, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 om = ld4(omega, i), gg = ld4_stream(g, i);
        om.x = fisher_elem(om.x, gg.x, len);
        om.y = fisher_elem(om.y, gg.y, len);
        om.z = fisher_elem(om.z, gg.z, len);
        om.w = fisher_elem(om.w, gg.w, len);
        st4(omega, i, om);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        omega[e] = fisher_elem(omega[e], g[e], len);
    }
}

__device__ __forceinline__ float mas_elem(float om, float g, float prev, float curr) {
    // omega = omega.mul(prev_size); omega = omega.add(|g|); omega = omega.div(curr_size)
    return __fdiv_rn(__fadd_rn(__fmul_rn(om, prev), fabsf(g)), curr);
}
__global__ void __launch_bounds__(kThreads)
mas_kernel(float* __restrict__ omega, const float* __restrict__ g, float prev, float curr, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 om = ld4(omega, i), gg = ld4_stream(g, i);
        om.x = mas_elem(om.x, gg.x, prev, curr);
        om.y = mas_elem(om.y, gg.y, prev, curr);
        om.z = mas_elem(om.z, gg.z, prev, curr);
        om.w = mas_elem(om.w, gg.w, prev, curr);
        st4(omega, i, om);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        omega[e] = mas_elem(omega[e], g[e], prev, curr);
    }
}

// mode-IMM merge (IMM/merge.py:228-231): mean_param += (precision_k / sum_precision) * theta_k, op by op like the reference
// (div, mul, add each rounded); first != 0 starts from the reference's torch.zeros.  Any alignment (per-tensor calls).
__global__ void __launch_bounds__(kThreads)
imm_merge_kernel(float* __restrict__ acc, const float* __restrict__ prec, const float* __restrict__ sum_prec,
                 const float* __restrict__ theta, int64_t n, int first) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = __fmul_rn(__fdiv_rn(prec[i], sum_prec[i]), theta[i]);
        acc[i] = __fadd_rn(first ? 0.f : acc[i], d);
    }
}

__device__ __forceinline__ void consolidate_elem(float& om, float& w, float th, float& ts, float slack) {
    const float pd = __fsub_rn(th, ts);
    const float dom = __fadd_rn(__fmul_rn(pd, pd), slack);        // path_diff.pow(2).add_(slak)
    const float t = fmaxf(__fdiv_rn(w, dom), 0.0f);                // max(w / dominator, 0)
    om = __fadd_rn(om, t);
    w = 0.0f;
    ts = th;
}
__global__ void __launch_bounds__(kThreads)
si_consolidate_kernel(float* __restrict__ omega, float* __restrict__ w, const float* __restrict__ theta,
                      float* __restrict__ tstar, float slack, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 om = ld4(omega, i), ww = ld4(w, i), th = ld4(theta, i), ts = ld4(tstar, i);
        consolidate_elem(om.x, ww.x, th.x, ts.x, slack);
        consolidate_elem(om.y, ww.y, th.y, ts.y, slack);
        consolidate_elem(om.z, ww.z, th.z, ts.z, slack);
        consolidate_elem(om.w, ww.w, th.w, ts.w, slack);
        st4(omega, i, om);
        st4(w, i, ww);
        st4(tstar, i, ts);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        float om = omega[e], ww = w[e], ts = tstar[e];
        consolidate_elem(om, ww, theta[e], ts, slack);
        omega[e] = om;
        w[e] = ww;
        tstar[e] = ts;
    }
}

__global__ void __launch_bounds__(kThreads)
axpby_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b, float sb, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 x = ld4(a, i), y = ld4(b, i);
        if (sb == 1.0f) {
            x.x = __fadd_rn(x.x, y.x); x.y = __fadd_rn(x.y, y.y); x.z = __fadd_rn(x.z, y.z); x.w = __fadd_rn(x.w, y.w);
        } else {
            x.x = __fadd_rn(x.x, __fmul_rn(y.x, sb)); x.y = __fadd_rn(x.y, __fmul_rn(y.y, sb));
            x.z = __fadd_rn(x.z, __fmul_rn(y.z, sb)); x.w = __fadd_rn(x.w, __fmul_rn(y.w, sb));
        }
        st4(dst, i, x);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        dst[e] = (sb == 1.0f) ? __fadd_rn(a[e], b[e]) : __fadd_rn(a[e], __fmul_rn(b[e], sb));
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace clb

using namespace clb;

extern "C" {

int clb_sgd_penalty_step(float* theta, const float* g, const float* omega, const float* theta_star, float* buf,
                         int64_t n, int64_t n_penalised, float two_lambda, float lr, float momentum,
                         float weight_decay, float grad_scale, int first_step, void* stream) {
    CLB_CHECK_ARG(theta && g && buf && n >= 0 && n_penalised >= 0 && n_penalised <= n);
    CLB_CHECK_ARG(n_penalised == 0 || (omega && theta_star));
    CLB_CHECK_ARG(aligned16(theta) && aligned16(g) && aligned16(buf) && aligned16(omega) && aligned16(theta_star));
    if (n == 0) return CLB_OK;
    SgdArgs a{two_lambda, lr, momentum, weight_decay, grad_scale, first_step ? 1 : 0};
    launch_pdl(sgd_penalty_kernel, dim3(stream_grid(n >> 2)), dim3(kThreads), 0, as_stream(stream), theta, g, omega, theta_star, buf, n,
               n_penalised, a); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_si_step(float* theta, const float* g, const float* omega, const float* theta_star, float* buf, float* w,
                int64_t n, float two_lambda, float lr, float momentum, float weight_decay, float grad_scale,
                int first_step, void* stream) {
    CLB_CHECK_ARG(theta && g && omega && theta_star && buf && w && n >= 0);
    CLB_CHECK_ARG(aligned16(theta) && aligned16(g) && aligned16(buf) && aligned16(omega) && aligned16(theta_star) &&
                  aligned16(w));
    if (n == 0) return CLB_OK;
    SgdArgs a{two_lambda, lr, momentum, weight_decay, grad_scale, first_step ? 1 : 0};
    si_step_kernel<<<stream_grid(n >> 2), kThreads, 0, as_stream(stream)>>>(theta, g, omega, theta_star, buf, w, n, a); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_fisher_accum(float* omega, const float* g, float data_len, int64_t n, void* stream) {
    CLB_CHECK_ARG(omega && g && n >= 0 && data_len > 0 && aligned16(omega) && aligned16(g));
    if (n == 0) return CLB_OK;
    fisher_kernel<<<stream_grid(n >> 2), kThreads, 0, as_stream(stream)>>>(omega, g, data_len, n); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_mas_accum(float* omega, const float* g, float prev_size, float curr_size, int64_t n, void* stream) {
    CLB_CHECK_ARG(omega && g && n >= 0 && curr_size > 0 && aligned16(omega) && aligned16(g));
    if (n == 0) return CLB_OK;
    mas_kernel<<<stream_grid(n >> 2), kThreads, 0, as_stream(stream)>>>(omega, g, prev_size, curr_size, n); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_si_consolidate(float* omega, float* w, const float* theta, float* theta_star, float slack, int64_t n,
                       void* stream) {
    CLB_CHECK_ARG(omega && w && theta && theta_star && n >= 0);
    CLB_CHECK_ARG(aligned16(omega) && aligned16(w) && aligned16(theta) && aligned16(theta_star));
    if (n == 0) return CLB_OK;
    si_consolidate_kernel<<<stream_grid(n >> 2), kThreads, 0, as_stream(stream)>>>(omega, w, theta, theta_star, slack,
                                                                                    n); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_imm_merge_accum(float* acc, const float* prec, const float* sum_prec, const float* theta, int64_t n, int first,
                        void* stream) {
    CLB_CHECK_ARG(acc && prec && sum_prec && theta && n >= 0);
    if (n == 0) return CLB_OK;
    imm_merge_kernel<<<stream_grid((n + 3) >> 2), kThreads, 0, as_stream(stream)>>>(acc, prec, sum_prec, theta, n, first); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_axpby(float* dst, const float* a, const float* b, float scale_b, int64_t n, void* stream) {
    CLB_CHECK_ARG(dst && a && b && n >= 0 && aligned16(dst) && aligned16(a) && aligned16(b));
    if (n == 0) return CLB_OK;
    axpby_kernel<<<stream_grid(n >> 2), kThreads, 0, as_stream(stream)>>>(dst, a, b, scale_b, n); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

}  // extern "C"
