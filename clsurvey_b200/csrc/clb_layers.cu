// Non-GEMM layer kernels: ReLU backward, MaxPool2d fwd/bwd, AdaptiveAvgPool2d fwd/bwd, dropout-mask multiply,
// fused softmax / loss / argmax head.  All NCHW fp32 (the reference's layout).
//
// Reference call sites: nn.ReLU / nn.MaxPool2d in src/models/VGGSlim.py:27-40 and torchvision AlexNet,
// nn.CrossEntropyLoss (src/methods/EWC/main_EWC.py:55), sum-NLL (main_EWC.py:148), sum-of-squares (train_MAS.py:552-560),
// torch.max(outputs,1)/sum(preds==labels) (train_EWC.py:182,197), GEM head slice (rehearsal/model/gem.py:199-203,242).
#include <float.h>

#include "clb_common.cuh"

namespace clb {

static inline int ew_grid(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                int64_t n) {
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 d = reinterpret_cast<const float4*>(dy)[i];
        const float4 v = reinterpret_cast<const float4*>(y)[i];
        d.x = v.x > 0.f ? d.x : 0.f; d.y = v.y > 0.f ? d.y : 0.f;
        d.z = v.z > 0.f ? d.z : 0.f; d.w = v.w > 0.f ? d.w : 0.f;
        reinterpret_cast<float4*>(dx)[i] = d;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        dx[e] = y[e] > 0.f ? dy[e] : 0.f;
    }
}

// one thread per output element; window scanned row-major, strict '>' so the FIRST maximum wins (ATen semantics)
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ am,
                                   int64_t total, int H, int W, int PH, int PW, int k, int stride) {
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gs) {
        const int pw = (int)(o % PW);
        const int ph = (int)((o / PW) % PH);
        const int64_t nc = o / ((int64_t)PW * PH);
        const float* src = x + nc * H * W + (int64_t)(ph * stride) * W + pw * stride;
        float best = -FLT_MAX;
        int bi = 0;
        for (int r = 0; r < k; ++r)
            for (int s = 0; s < k; ++s) {
                const float v = src[r * W + s];
                if (v > best || v != v) { best = v; bi = r * k + s; }
            }
        y[o] = best;
        am[o] = (uint8_t)bi;
    }
}

// gather form (deterministic, no atomics): each INPUT element sums dy of the windows whose argmax it is
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ am,
                                   const float* __restrict__ relu_out, float* __restrict__ dx, int64_t total, int H,
                                   int W, int PH, int PW, int k, int stride) {
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const int64_t nc = i / ((int64_t)W * H);
        float acc = 0.f;
        if (relu_out == nullptr || relu_out[i] > 0.f) {
            int ph0 = h - k + 1;
            ph0 = ph0 <= 0 ? 0 : (ph0 + stride - 1) / stride;
            int pw0 = w - k + 1;
            pw0 = pw0 <= 0 ? 0 : (pw0 + stride - 1) / stride;
            const int ph1 = min(h / stride, PH - 1), pw1 = min(w / stride, PW - 1);
            for (int ph = ph0; ph <= ph1; ++ph)
                for (int pw = pw0; pw <= pw1; ++pw) {
                    const int64_t o = (nc * PH + ph) * PW + pw;
                    const int local = (h - ph * stride) * k + (w - pw * stride);
                    if ((int)am[o] == local) acc += dy[o];
                }
        }
        dx[i] = acc;
    }
}

// 2x2 / stride-2 specialisation (every VGG pool): one thread per PAIR of adjacent outputs -> float4 input rows,
// float2 output; H, W even and W % 4 == 0.  Same first-max-wins rule.
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ am,
                                    int64_t total_pairs, int H, int W) {
    const int PW2 = W >> 2, PH = H >> 1;
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_pairs; t += gs) {
        const int pw2 = (int)(t % PW2), ph = (int)((t / PW2) % PH);
        const int64_t nc = t / ((int64_t)PW2 * PH);
        const float* src = x + nc * H * W + (int64_t)(2 * ph) * W + 4 * pw2;
        const float4 r0 = *reinterpret_cast<const float4*>(src), r1 = *reinterpret_cast<const float4*>(src + W);
        float b0 = r0.x; int i0 = 0;
        if (r0.y > b0 || r0.y != r0.y) { b0 = r0.y; i0 = 1; }
        if (r1.x > b0 || r1.x != r1.x) { b0 = r1.x; i0 = 2; }
        if (r1.y > b0 || r1.y != r1.y) { b0 = r1.y; i0 = 3; }
        float b1 = r0.z; int i1 = 0;
        if (r0.w > b1 || r0.w != r0.w) { b1 = r0.w; i1 = 1; }
        if (r1.z > b1 || r1.z != r1.z) { b1 = r1.z; i1 = 2; }
        if (r1.w > b1 || r1.w != r1.w) { b1 = r1.w; i1 = 3; }
        const int64_t o = (nc * PH + ph) * (W >> 1) + 2 * pw2;
        *reinterpret_cast<float2*>(y + o) = make_float2(b0, b1);
        *reinterpret_cast<uchar2*>(am + o) = make_uchar2((unsigned char)i0, (unsigned char)i1);
    }
}
__global__ void maxpool2_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ am,
                                    const float* __restrict__ relu_out, float* __restrict__ dx, int64_t total_pairs,
                                    int H, int W) {
    const int PW2 = W >> 2, PH = H >> 1;
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_pairs; t += gs) {
        const int pw2 = (int)(t % PW2), ph = (int)((t / PW2) % PH);
        const int64_t nc = t / ((int64_t)PW2 * PH);
        const int64_t o = (nc * PH + ph) * (W >> 1) + 2 * pw2;
        const float2 d = *reinterpret_cast<const float2*>(dy + o);
        const uchar2 a = *reinterpret_cast<const uchar2*>(am + o);
        const int64_t ioff = nc * H * W + (int64_t)(2 * ph) * W + 4 * pw2;
        float4 r0 = make_float4(a.x == 0 ? d.x : 0.f, a.x == 1 ? d.x : 0.f, a.y == 0 ? d.y : 0.f, a.y == 1 ? d.y : 0.f);
        float4 r1 = make_float4(a.x == 2 ? d.x : 0.f, a.x == 3 ? d.x : 0.f, a.y == 2 ? d.y : 0.f, a.y == 3 ? d.y : 0.f);
        if (relu_out) {
            const float4 m0 = *reinterpret_cast<const float4*>(relu_out + ioff);
            const float4 m1 = *reinterpret_cast<const float4*>(relu_out + ioff + W);
            r0.x = m0.x > 0.f ? r0.x : 0.f; r0.y = m0.y > 0.f ? r0.y : 0.f; r0.z = m0.z > 0.f ? r0.z : 0.f; r0.w = m0.w > 0.f ? r0.w : 0.f;
            r1.x = m1.x > 0.f ? r1.x : 0.f; r1.y = m1.y > 0.f ? r1.y : 0.f; r1.z = m1.z > 0.f ? r1.z : 0.f; r1.w = m1.w > 0.f ? r1.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + ioff) = r0;
        *reinterpret_cast<float4*>(dx + ioff + W) = r1;
    }
}

// AdaptiveAvgPool2d: window of output (oh,ow) = [floor(oh*H/OH), ceil((oh+1)*H/OH))
__global__ void aavgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int H, int W,
                                    int OH, int OW) {
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gs) {
        const int ow = (int)(o % OW), oh = (int)((o / OW) % OH);
        const int64_t nc = o / ((int64_t)OW * OH);
        const int h0 = (oh * H) / OH, h1 = ((oh + 1) * H + OH - 1) / OH;
        const int w0 = (ow * W) / OW, w1 = ((ow + 1) * W + OW - 1) / OW;
        float s = 0.f;
        for (int h = h0; h < h1; ++h)
            for (int w = w0; w < w1; ++w) s += x[nc * H * W + (int64_t)h * W + w];
        y[o] = s / (float)((h1 - h0) * (w1 - w0));
    }
}
__global__ void aavgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t total, int H, int W,
                                    int OH, int OW) {
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const int64_t nc = i / ((int64_t)W * H);
        float acc = 0.f;
        for (int oh = 0; oh < OH; ++oh) {
            const int h0 = (oh * H) / OH, h1 = ((oh + 1) * H + OH - 1) / OH;
            if (h < h0 || h >= h1) continue;
            for (int ow = 0; ow < OW; ++ow) {
                const int w0 = (ow * W) / OW, w1 = ((ow + 1) * W + OW - 1) / OW;
                if (w < w0 || w >= w1) continue;
                acc += dy[(nc * OH + oh) * OW + ow] / (float)((h1 - h0) * (w1 - w0));
            }
        }
        dx[i] = acc;
    }
}

__global__ void mask_mul_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y,
                                int64_t total, int cols, int bcast) {
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs)
        y[i] = x[i] * mask[bcast ? (i % cols) : i];
}

// Loss head. One CTA; thread r handles rows r, r+blockDim, ... ; deterministic fixed-order reduction.
constexpr int kLossThreads = 256;
__global__ void __launch_bounds__(kLossThreads)
softmax_loss_kernel(const float* __restrict__ logits, int ld, int col_off, int ncols, const int64_t* __restrict__ labels,
                    int B, int mode, float denom, float* __restrict__ loss_out, int* __restrict__ correct_out,
                    float* __restrict__ dlogits) {
    pdl_trigger();
    pdl_wait();
    float my_loss = 0.f;
    int my_corr = 0;
    for (int r = threadIdx.x; r < B; r += blockDim.x) {
        const float* z = logits + (int64_t)r * ld + col_off;
        float* dz = dlogits ? dlogits + (int64_t)r * ld : nullptr;
        const int y = labels ? (int)labels[r] : -1;
        float mx = -FLT_MAX;
        int am = 0;
        for (int c = 0; c < ncols; ++c) {
            const float v = z[c];
            if (v > mx) { mx = v; am = c; }
        }
        if (labels && am == y) ++my_corr;
        if (dz) {
            for (int c = 0; c < col_off; ++c) dz[c] = 0.f;
            for (int c = col_off + ncols; c < ld; ++c) dz[c] = 0.f;
        }
        if (mode == CLB_LOSS_SUM_SQ) {
            float s = 0.f;
            for (int c = 0; c < ncols; ++c) {
                const float v = z[c];
                s += v * v;
                if (dz) dz[col_off + c] = 2.0f * v;
            }
            my_loss += s;
        } else {
            float se = 0.f;
            for (int c = 0; c < ncols; ++c) se += expf(z[c] - mx);
            const float lse = logf(se);
            const float scale = (mode == CLB_LOSS_MEAN_CE) ? 1.0f / denom : 1.0f;
            const bool y_ok = (y >= 0 && y < ncols);          // out-of-range label: poison the loss, never read OOB
            my_loss += y_ok ? -((z[y] - mx) - lse) : __int_as_float(0x7fc00000);
            if (dz) {
                for (int c = 0; c < ncols; ++c) {
                    const float p = expf((z[c] - mx) - lse);
                    dz[col_off + c] = (p - (c == y ? 1.0f : 0.0f)) * scale;
                }
            }
        }
    }
    __shared__ float s_loss[kLossThreads];
    __shared__ int s_corr[kLossThreads];
    s_loss[threadIdx.x] = my_loss;
    s_corr[threadIdx.x] = my_corr;
    __syncthreads();
    for (int o = kLossThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
            s_corr[threadIdx.x] += s_corr[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float l = s_loss[0];
        if (mode == CLB_LOSS_MEAN_CE) l = l / denom;
        loss_out[0] += l;
        if (correct_out) correct_out[0] += s_corr[0];
    }
}

}  // namespace clb

using namespace clb;

extern "C" {

int clb_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
    CLB_CHECK_ARG(dy && y && dx && n >= 0);
    CLB_CHECK_ARG((((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx) & 15) == 0);
    if (n == 0) return CLB_OK;
    relu_bwd_kernel<<<ew_grid(n >> 2, 256), 256, 0, as_stream(stream)>>>(dy, y, dx, n); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_maxpool_fwd(const float* x, float* y, uint8_t* argmax, int N, int C, int H, int W, int k, int stride,
                    void* stream) {
    CLB_CHECK_ARG(x && y && argmax && N > 0 && C > 0 && k >= 1 && k <= 15 && stride >= 1 && H >= k && W >= k);
    const int PH = (H - k) / stride + 1, PW = (W - k) / stride + 1;
    const int64_t total = (int64_t)N * C * PH * PW;
    if (k == 2 && stride == 2 && (W & 3) == 0 && (H & 1) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
        const int64_t pairs = total >> 1;
        maxpool2_fwd_kernel<<<ew_grid(pairs, 256), 256, 0, as_stream(stream)>>>(x, y, argmax, pairs, H, W); clb::count_launch();
        CLB_CHECK_LAUNCH();
        return CLB_OK;
    }
    maxpool_fwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(x, y, argmax, total, H, W, PH, PW, k, stride); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_maxpool_bwd(const float* dy, const uint8_t* argmax, const float* x_relu_out, float* dx, int N, int C, int H,
                    int W, int k, int stride, void* stream) {
    CLB_CHECK_ARG(dy && argmax && dx && N > 0 && C > 0 && k >= 1 && k <= 15 && stride >= 1 && H >= k && W >= k);
    const int PH = (H - k) / stride + 1, PW = (W - k) / stride + 1;
    const int64_t total = (int64_t)N * C * H * W;
    if (k == 2 && stride == 2 && (W & 3) == 0 && (H & 1) == 0 && (((uintptr_t)dy | (uintptr_t)dx | (uintptr_t)x_relu_out) & 15) == 0) {
        const int64_t pairs = total >> 3;
        maxpool2_bwd_kernel<<<ew_grid(pairs, 256), 256, 0, as_stream(stream)>>>(dy, argmax, x_relu_out, dx, pairs, H, W); clb::count_launch();
        CLB_CHECK_LAUNCH();
        return CLB_OK;
    }
    maxpool_bwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(dy, argmax, x_relu_out, dx, total, H, W, PH,
                                                                            PW, k, stride); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_adaptive_avgpool_fwd(const float* x, float* y, int N, int C, int H, int W, int OH, int OW, void* stream) {
    CLB_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0);
    const int64_t total = (int64_t)N * C * OH * OW;
    aavgpool_fwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(x, y, total, H, W, OH, OW); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}
int clb_adaptive_avgpool_bwd(const float* dy, float* dx, int N, int C, int H, int W, int OH, int OW, void* stream) {
    CLB_CHECK_ARG(dy && dx && N > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0);
    const int64_t total = (int64_t)N * C * H * W;
    aavgpool_bwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(dy, dx, total, H, W, OH, OW); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_mask_mul(const float* x, const float* mask, float* y, int rows, int cols, int mask_rows, void* stream) {
    CLB_CHECK_ARG(x && mask && y && rows > 0 && cols > 0 && (mask_rows == 1 || mask_rows == rows));
    const int64_t total = (int64_t)rows * cols;
    mask_mul_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(x, mask, y, total, cols, mask_rows == 1 && rows != 1); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_softmax_loss(const float* logits, int ld, int col_off, int ncols, const int64_t* labels, int B, int mode,
                     float mean_denominator, float* loss_out, int* correct_out, float* dlogits, void* stream) {
    CLB_CHECK_ARG(logits && loss_out && B > 0 && ncols > 0 && col_off >= 0 && col_off + ncols <= ld);
    CLB_CHECK_ARG(mode == CLB_LOSS_MEAN_CE || mode == CLB_LOSS_SUM_NLL || mode == CLB_LOSS_SUM_SQ);
    CLB_CHECK_ARG(mode == CLB_LOSS_SUM_SQ || labels != nullptr);
    CLB_CHECK_ARG(mode != CLB_LOSS_MEAN_CE || mean_denominator > 0.f);
    launch_pdl(softmax_loss_kernel, dim3(1), dim3(kLossThreads), 0, as_stream(stream), logits, ld, col_off, ncols, labels, B, mode,
               mean_denominator, loss_out, correct_out, dlogits); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

}  // extern "C"
