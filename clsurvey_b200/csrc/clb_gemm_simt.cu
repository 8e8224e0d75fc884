// Exact-fp32 implicit-GEMM path (FFMA, CUDA cores) for conv2d fwd/dgrad/wgrad and linear fwd/dgrad/wgrad.
//
// This is the bit-faithful fp32 path (CLB_MM_FP32_SIMT): every product and sum is an fp32 FMA, like the reference's
// cuDNN/cuBLAS fp32 kernels (nn.Conv2d / nn.Linear in src/models/VGGSlim.py:27-76, torchvision AlexNet).  It also
// serves the shapes the tcgen05 path does not take (C=3 first layers, 11x11 stride-4, 20-wide heads).
// One generic register-tiled kernel (16x16 threads, TMxTN micro-tile, BK=16, double-buffered smem) is specialised
// by operand accessors -- im2col gathers are generated on the fly, nothing is materialised in HBM:
//     conv fwd :  Y[pix, kout]   = sum_crs  im2col(X)[pix, crs] * W[kout, crs]
//     conv wgrad: dW[kout, crs]  = sum_pix  dY[kout, pix]      * im2col(X)[pix, crs]      (split-K, 2-pass, deterministic)
//     conv dgrad (stride 1) = conv fwd of dY with flipped/transposed weights
#include <stdlib.h>

#include "clb_common.cuh"

namespace clb {

struct FastDiv {
    uint32_t d, magic, shift;
    FastDiv() : d(1), magic(0), shift(0) {}
    explicit FastDiv(uint32_t dd) : d(dd) {
        if (dd <= 1) { d = 1; magic = 0; shift = 0; return; }
        shift = 0;
        while ((1ull << shift) < dd) ++shift;
        magic = (uint32_t)(((1ull << 32) * ((1ull << shift) - dd)) / dd + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        return d == 1 ? n : (uint32_t)(((uint64_t)__umulhi(n, magic) + n) >> shift);
    }
};

// ---- operand accessors: element(outer, inner); the GEMM reduction index is always `inner` ----------------------
struct Idx { int off; bool ok; };

struct StridedMat {               // element = p[outer*so + inner*si]
    const float* p; int64_t so, si; int n_outer, n_inner;
    struct Outer { int64_t off; bool ok; };
    struct Inner { int64_t off; bool ok; };
    __device__ __forceinline__ Outer outer(int o) const { return {o * so, o < n_outer}; }
    __device__ __forceinline__ Inner inner(int i) const { return {i * si, i < n_inner}; }
    __device__ __forceinline__ float load(const Outer& o, const Inner& i) const { return (o.ok && i.ok) ? p[o.off + i.off] : 0.f; }
};

struct ConvGeom {
    int N, C, H, W, K, R, S, stride, pad, P, Q;
    FastDiv dPQ, dQ, dRS, dS;
};

struct PixInfo { int base; int ih0, iw0; bool ok; };   // base = img*C*H*W + ih0*W + iw0 (may point before the image: guarded)
struct CrsInfo { int off; int r, s; bool ok; };        // off = c*H*W + r*W + s

struct Im2colBase {
    const float* x; ConvGeom g; int n_pix, n_crs;
    __device__ __forceinline__ PixInfo pix(int m) const {
        PixInfo o;
        o.ok = m < n_pix;
        const uint32_t img = g.dPQ.div(m), pq = m - img * (g.P * g.Q);
        const uint32_t p = g.dQ.div(pq), q = pq - p * g.Q;
        o.ih0 = (int)p * g.stride - g.pad;
        o.iw0 = (int)q * g.stride - g.pad;
        o.base = (int)img * g.C * g.H * g.W + o.ih0 * g.W + o.iw0;
        return o;
    }
    __device__ __forceinline__ CrsInfo crs(int k) const {
        CrsInfo o;
        o.ok = k < n_crs;
        const uint32_t c = g.dRS.div(k), rs = k - c * (g.R * g.S);
        const uint32_t r = g.dS.div(rs), s = rs - r * g.S;
        o.r = (int)r; o.s = (int)s;
        o.off = (int)c * g.H * g.W + (int)r * g.W + (int)s;
        return o;
    }
    __device__ __forceinline__ float at(const PixInfo& p, const CrsInfo& c) const {
        const int ih = p.ih0 + c.r, iw = p.iw0 + c.s;
        return (p.ok && c.ok && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W) ? x[p.base + c.off] : 0.f;
    }
};
struct Im2colPixOuter : Im2colBase {   // A of conv fwd: outer = pixel, inner = (c,r,s)
    using Outer = PixInfo; using Inner = CrsInfo;
    __device__ __forceinline__ Outer outer(int o) const { return pix(o); }
    __device__ __forceinline__ Inner inner(int i) const { return crs(i); }
    __device__ __forceinline__ float load(const Outer& o, const Inner& i) const { return at(o, i); }
};
struct Im2colCrsOuter : Im2colBase {   // B of conv wgrad: outer = (c,r,s), inner = pixel
    using Outer = CrsInfo; using Inner = PixInfo;
    __device__ __forceinline__ Outer outer(int o) const { return crs(o); }
    __device__ __forceinline__ Inner inner(int i) const { return pix(i); }
    __device__ __forceinline__ float load(const Outer& o, const Inner& i) const { return at(i, o); }
};
struct DyKoutOuter {                   // A of conv wgrad: element(kout, pix) = dy[img][kout][pq]
    const float* dy; int K, PQ, n_pix; FastDiv dPQ;
    struct Outer { int off; bool ok; };
    struct Inner { int off; bool ok; };
    __device__ __forceinline__ Outer outer(int k) const { return {k * PQ, k < K}; }
    __device__ __forceinline__ Inner inner(int m) const {
        const uint32_t img = dPQ.div(m), pq = m - img * PQ;
        return {(int)img * K * PQ + (int)pq, m < n_pix};
    }
    __device__ __forceinline__ float load(const Outer& o, const Inner& i) const { return (o.ok && i.ok) ? dy[o.off + i.off] : 0.f; }
};

// ---- epilogues ------------------------------------------------------------------------------------------------
struct EpiRowMajor {                   // C[m*ldc + n] = act(acc + bias[n]); contiguous along n
    float* c; int64_t ldc; const float* bias; int relu; int M, N;
    __device__ __forceinline__ void store(int m, int n, const float* v, int cnt) const {
        if (m >= M) return;
        float* dst = c + (int64_t)m * ldc + n;
        if (cnt == 4 && n + 3 < N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (bias) { o.x += bias[n]; o.y += bias[n + 1]; o.z += bias[n + 2]; o.w += bias[n + 3]; }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(dst) = o;
            return;
        }
        for (int j = 0; j < cnt; ++j)
            if (n + j < N) {
                float o = v[j] + (bias ? bias[n + j] : 0.f);
                dst[j] = relu ? fmaxf(o, 0.f) : o;
            }
    }
};
struct EpiConvNCHW {                   // y[img][n][pq] = act(acc + bias[n]), m = img*PQ + pq; contiguous along m
    float* y; const float* bias; int relu; int M, N, PQ; FastDiv dPQ;
    __device__ __forceinline__ void store(int m, int n, const float* v, int cnt) const {
        if (n >= N || m >= M) return;
        const float b = bias ? bias[n] : 0.f;
        const uint32_t img = dPQ.div(m), pq = m - img * PQ;
        float* dst = y + ((int64_t)img * N + n) * PQ + pq;
        if (cnt == 4 && (PQ & 3) == 0 && m + 3 < M && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            float4 o = make_float4(v[0] + b, v[1] + b, v[2] + b, v[3] + b);
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(dst) = o;
            return;
        }
        for (int j = 0; j < cnt; ++j) {
            const int mm = m + j;
            if (mm >= M) break;
            const uint32_t im = dPQ.div(mm), pp = mm - im * PQ;
            const float o = v[j] + b;
            y[((int64_t)im * N + n) * PQ + pp] = relu ? fmaxf(o, 0.f) : o;
        }
    }
};

template <int T> __device__ __forceinline__ int chunk_index(int t, int i, int B) {
    if (T == 8) return (i < 4) ? (t * 4 + i) : (B / 2 + t * 4 + (i - 4));
    if (T == 4) return t * 4 + i;
    return t * T + i;
}

template <int TM, int TN, class AAcc, class BAcc, bool A_OC, bool B_OC, class Epi, bool C_MMAJOR>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(AAcc A, BAcc B, Epi epi, int K, int k_chunk, int64_t split_stride) {
    constexpr int BM = 16 * TM, BN = 16 * TN, BK = 16, PAD = 4;
    constexpr int LA = BM * BK / 256, LB = BN * BK / 256;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * k_chunk;
    const int kend = min(K, kbeg + k_chunk);
    const int tx = tid & 15, ty = tid >> 4;
    const int tm = C_MMAJOR ? tx : ty, tn = C_MMAJOR ? ty : tx;

    typename AAcc::Outer ao[A_OC ? 1 : LA];
    typename BAcc::Outer bo[B_OC ? 1 : LB];
    if (A_OC) ao[0] = A.outer(m0 + tid % BM);
    else {
#pragma unroll
        for (int j = 0; j < LA; ++j) ao[A_OC ? 0 : j] = A.outer(m0 + (tid >> 4) + 16 * j);
    }
    if (B_OC) bo[0] = B.outer(n0 + tid % BN);
    else {
#pragma unroll
        for (int j = 0; j < LB; ++j) bo[B_OC ? 0 : j] = B.outer(n0 + (tid >> 4) + 16 * j);
    }

    float ra[LA], rb[LB];
    auto gload = [&](int k0) {
        if (A_OC) {
#pragma unroll
            for (int j = 0; j < LA; ++j) {
                const int k = k0 + tid / BM + (256 / BM) * j;
                ra[j] = (k < kend) ? A.load(ao[0], A.inner(k)) : 0.f;
            }
        } else {
            const int k = k0 + (tid & 15);
            const typename AAcc::Inner in = A.inner(k);
#pragma unroll
            for (int j = 0; j < LA; ++j) ra[j] = (k < kend) ? A.load(ao[A_OC ? 0 : j], in) : 0.f;
        }
        if (B_OC) {
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int k = k0 + tid / BN + (256 / BN) * j;
                rb[j] = (k < kend) ? B.load(bo[0], B.inner(k)) : 0.f;
            }
        } else {
            const int k = k0 + (tid & 15);
            const typename BAcc::Inner in = B.inner(k);
#pragma unroll
            for (int j = 0; j < LB; ++j) rb[j] = (k < kend) ? B.load(bo[B_OC ? 0 : j], in) : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < LA; ++j) {
            if (A_OC) As[buf][tid / BM + (256 / BM) * j][tid % BM] = ra[j];
            else As[buf][tid & 15][(tid >> 4) + 16 * j] = ra[j];
        }
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            if (B_OC) Bs[buf][tid / BN + (256 / BN) * j][tid % BN] = rb[j];
            else Bs[buf][tid & 15][(tid >> 4) + 16 * j] = rb[j];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (kbeg < kend) {
        gload(kbeg);
        sstore(0);
    }
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK, buf ^= 1) {
        const bool has_next = k0 + BK < kend;
        if (has_next) gload(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][chunk_index<TM>(tm, i, BM)];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][chunk_index<TN>(tn, j, BN)];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) sstore(buf ^ 1);
        __syncthreads();
    }

    Epi e = epi;
    e.shift(blockIdx.z * split_stride);
    if (C_MMAJOR) {  // contiguous along m: chunks of consecutive m for each n
        constexpr int CH = TM >= 4 ? 4 : TM;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + chunk_index<TN>(tn, j, BN);
#pragma unroll
            for (int i = 0; i < TM; i += CH) {
                float v[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) v[c] = acc[i + c][j];
                e.store(m0 + chunk_index<TM>(tm, i, BM), n, v, CH);
            }
        }
    } else {
        constexpr int CH = TN >= 4 ? 4 : TN;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + chunk_index<TM>(tm, i, BM);
#pragma unroll
            for (int j = 0; j < TN; j += CH) {
                float v[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) v[c] = acc[i][j + c];
                e.store(m, n0 + chunk_index<TN>(tn, j, BN), v, CH);
            }
        }
    }
}

// epilogue wrappers that add the split-K shift hook
struct EpiRM : EpiRowMajor { __device__ __forceinline__ void shift(int64_t s) { c += s; } };
struct EpiNCHW : EpiConvNCHW { __device__ __forceinline__ void shift(int64_t) {} };

// ---- small helper kernels -----------------------------------------------------------------------------------
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int64_t n, int splits) {
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        float s = ws[i];
        for (int z = 1; z < splits; ++z) s += ws[(int64_t)z * n + i];   // fixed order: deterministic
        out[i] = s;
    }
}

// dbias[k] = sum over (img, pq) of dy[img][k][pq].  Two deterministic stages: grid (K, nsplit) CTAs each reduce a
// contiguous range of images of one channel (float4 loads, fixed-order tree), then the first nsplit partials are summed
// in order by the last stage.  No atomics: bit-reproducible.
// WRITE_LO: the same pass also writes the lo plane (x - trunc19(x)) of dY that the TMA-fed wgrad kernel consumes, so dY
// is read once for both (the summation order of the bias partials is unchanged).
// WRITE_LO == 2: instead of the fp32 lo plane, the bf16 hi plane (into dy_lo) and bf16 lo plane (into dy_lo2) of dY for
// the kind::f16 wgrad kernel (hi = bf16_rn(x), lo = bf16_rn(x - hi)).
__device__ __forceinline__ void bf16_split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    uint32_t a = __float_as_uint(v0), b = __float_as_uint(v1);
    a += 0x7FFFu + ((a >> 16) & 1u);
    b += 0x7FFFu + ((b >> 16) & 1u);
    hi = (a >> 16) | (b & 0xFFFF0000u);
    uint32_t c = __float_as_uint(v0 - __uint_as_float(a & 0xFFFF0000u)), d = __float_as_uint(v1 - __uint_as_float(b & 0xFFFF0000u));
    c += 0x7FFFu + ((c >> 16) & 1u);
    d += 0x7FFFu + ((d >> 16) & 1u);
    lo = (c >> 16) | (d & 0xFFFF0000u);
}
template <int WRITE_LO>
__global__ void __launch_bounds__(256) conv_bias_grad_partial_kernel(const float* __restrict__ dy, float* __restrict__ part,
                                                                     float* __restrict__ dy_lo, uint16_t* __restrict__ dy_lo2,
                                                                     int N, int K, int PQ, int imgs_per_split) {
    const int k = blockIdx.x, sp = blockIdx.y;
    const int n0 = sp * imgs_per_split, n1 = min(N, n0 + imgs_per_split);
    float s = 0.f;
    if ((PQ & 3) == 0) {
        const int pq4 = PQ >> 2;
        const int total = (n1 - n0) * pq4;
        for (int i = threadIdx.x; i < total; i += 256) {
            const int img = n0 + i / pq4, q = i - (i / pq4) * pq4;
            const int64_t row = ((int64_t)img * K + k) * PQ;
            const float4 v = __ldg(reinterpret_cast<const float4*>(dy + row) + q);
            s += (v.x + v.y) + (v.z + v.w);
            if (WRITE_LO == 2) {
                uint32_t h0, h1, l0, l1;
                bf16_split_pair(v.x, v.y, h0, l0);
                bf16_split_pair(v.z, v.w, h1, l1);
                reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dy_lo) + row)[q] = make_uint2(h0, h1);
                reinterpret_cast<uint2*>(dy_lo2 + row)[q] = make_uint2(l0, l1);
            } else if (WRITE_LO == 1) {
                float4 o;
                o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                reinterpret_cast<float4*>(dy_lo + row)[q] = o;
            }
        }
    } else {
        const int total = (n1 - n0) * PQ;
        for (int i = threadIdx.x; i < total; i += 256) {
            const int img = n0 + i / PQ, pq = i - (i / PQ) * PQ;
            s += dy[((int64_t)img * K + k) * PQ + pq];
        }
    }
    __shared__ float red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[(int64_t)sp * K + k] = red[0];
}
__global__ void conv_bias_grad_final_kernel(const float* __restrict__ part, float* __restrict__ db, int K, int nsplit) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float s = part[k];
    for (int sp = 1; sp < nsplit; ++sp) s += part[(int64_t)sp * K + k];
    db[k] = s;
}
static inline int bias_grad_splits(int N, int K) {
    int want = (4 * 148 + K - 1) / K;            // ~4 CTAs per SM (fixed 148 => device-independent summation order)
    if (want > N) want = N;
    if (want > 64) want = 64;
    return want < 1 ? 1 : want;
}
// scratch: bias_grad_splits(N,K) * K floats, taken from the tail of the wgrad workspace
static void conv_bias_grad(const float* dy, float* db, float* scratch, int N, int K, int PQ, cudaStream_t s) {
    const int want = bias_grad_splits(N, K);
    const int per = (N + want - 1) / want, nsplit = (N + per - 1) / per;
    conv_bias_grad_partial_kernel<0><<<dim3(K, nsplit), 256, 0, s>>>(dy, scratch, nullptr, nullptr, N, K, PQ, per); clb::count_launch();
    conv_bias_grad_final_kernel<<<(K + 127) / 128, 128, 0, s>>>(scratch, db, K, nsplit); clb::count_launch();
}
// First half of conv_bias_grad fused with the lo-plane split of dY (PQ % 4 == 0); conv_bias_grad_finish() is the rest.
void conv_bias_partials_and_lo(const float* dy, float* dy_lo, float* scratch, int N, int K, int PQ, cudaStream_t s) {
    const int want = bias_grad_splits(N, K);
    const int per = (N + want - 1) / want, nsplit = (N + per - 1) / per;
    conv_bias_grad_partial_kernel<1><<<dim3(K, nsplit), 256, 0, s>>>(dy, scratch, dy_lo, nullptr, N, K, PQ, per); clb::count_launch();
}
void conv_bias_partials_and_bf16(const float* dy, uint16_t* dy_hi, uint16_t* dy_lo, float* scratch, int N, int K, int PQ,
                                 cudaStream_t s) {
    const int want = bias_grad_splits(N, K);
    const int per = (N + want - 1) / want, nsplit = (N + per - 1) / per;
    conv_bias_grad_partial_kernel<2><<<dim3(K, nsplit), 256, 0, s>>>(dy, scratch, reinterpret_cast<float*>(dy_hi), dy_lo, N, K, PQ,
                                                                     per); clb::count_launch();
}
static void conv_bias_grad_finish(const float* scratch, float* db, int N, int K, cudaStream_t s) {
    const int want = bias_grad_splits(N, K);
    const int per = (N + want - 1) / want, nsplit = (N + per - 1) / per;
    conv_bias_grad_final_kernel<<<(K + 127) / 128, 128, 0, s>>>(scratch, db, K, nsplit); clb::count_launch();
}

// db[o] = sum_m dy[m][o]: one CTA of 32 x 8 threads per 32 outputs (lanes along o: coalesced rows; the 8 row groups are added
// in a fixed order: deterministic)
__global__ void __launch_bounds__(256) linear_bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, int M, int out) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (o < out)
        for (int m = grp; m < M; m += 8) s += dy[(int64_t)m * out + o];
    red[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && o < out) {
        float t = red[0][lane];
        for (int g = 1; g < 8; ++g) t += red[g][lane];
        db[o] = t;
    }
}

// wt[c][k][R-1-r][S-1-s] = w[k][c][r][s]   (dgrad as a forward conv of dY)
__global__ void flip_transpose_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int C, int R,
                                              int S) {
    const int64_t total = (int64_t)K * C * R * S, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int s = (int)(i % S), r = (int)((i / S) % R);
        const int c = (int)((i / ((int64_t)S * R)) % C), k = (int)(i / ((int64_t)S * R * C));
        wt[(((int64_t)c * K + k) * R + (R - 1 - r)) * S + (S - 1 - s)] = w[i];
    }
}

// generic (any stride) dgrad, gather form: used only when stride > 1 (no such layer needs dgrad in AlexNet/VGG)
__global__ void conv_dgrad_naive_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                        ConvGeom g) {
    const int64_t total = (int64_t)g.N * g.C * g.H * g.W, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
        const int iw = (int)(i % g.W), ih = (int)((i / g.W) % g.H);
        const int c = (int)((i / ((int64_t)g.W * g.H)) % g.C), n = (int)(i / ((int64_t)g.W * g.H * g.C));
        float s = 0.f;
        for (int k = 0; k < g.K; ++k)
            for (int r = 0; r < g.R; ++r) {
                const int hh = ih + g.pad - r;
                if (hh < 0 || hh % g.stride) continue;
                const int p = hh / g.stride;
                if (p >= g.P) continue;
                for (int q0 = 0; q0 < g.S; ++q0) {
                    const int ww = iw + g.pad - q0;
                    if (ww < 0 || ww % g.stride) continue;
                    const int q = ww / g.stride;
                    if (q >= g.Q) continue;
                    s = fmaf(dy[(((int64_t)n * g.K + k) * g.P + p) * g.Q + q], w[(((int64_t)k * g.C + c) * g.R + r) * g.S + q0], s);
                }
            }
        dx[i] = s;
    }
}

static ConvGeom make_geom(int N, int C, int H, int W, int K, int R, int S, int stride, int pad) {
    ConvGeom g;
    g.N = N; g.C = C; g.H = H; g.W = W; g.K = K; g.R = R; g.S = S; g.stride = stride; g.pad = pad;
    g.P = (H + 2 * pad - R) / stride + 1;
    g.Q = (W + 2 * pad - S) / stride + 1;
    g.dPQ = FastDiv(g.P * g.Q); g.dQ = FastDiv(g.Q); g.dRS = FastDiv(R * S); g.dS = FastDiv(S);
    return g;
}

static int conv_fwd_simt(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, int relu,
                         cudaStream_t s) {
    const int M = g.N * g.P * g.Q, N = g.K, Kg = g.C * g.R * g.S;
    Im2colPixOuter A; A.x = x; A.g = g; A.n_pix = M; A.n_crs = Kg;
    StridedMat B{w, Kg, 1, N, Kg};
    EpiNCHW e; e.y = y; e.bias = bias; e.relu = relu; e.M = M; e.N = N; e.PQ = g.P * g.Q; e.dPQ = g.dPQ;
    const int64_t tiles128 = (int64_t)((M + 127) / 128) * ((N + 127) / 128);
    if (N <= 32) {
        dim3 grid((M + 127) / 128, (N + 31) / 32, 1);
        gemm_simt_kernel<8, 2, Im2colPixOuter, StridedMat, true, false, EpiNCHW, true><<<grid, 256, 0, s>>>(A, B, e, Kg, Kg, 0); clb::count_launch();
    } else if (tiles128 < 2 * sm_count() || N <= 64) {
        dim3 grid((M + 63) / 64, (N + 63) / 64, 1);
        gemm_simt_kernel<4, 4, Im2colPixOuter, StridedMat, true, false, EpiNCHW, true><<<grid, 256, 0, s>>>(A, B, e, Kg, Kg, 0); clb::count_launch();
    } else {
        dim3 grid((M + 127) / 128, (N + 127) / 128, 1);
        gemm_simt_kernel<8, 8, Im2colPixOuter, StridedMat, true, false, EpiNCHW, true><<<grid, 256, 0, s>>>(A, B, e, Kg, Kg, 0); clb::count_launch();
    }
    return 0;
}

static void wgrad_plan(const ConvGeom& g, int* splits, int* k_chunk, int* bm) {
    const int M = g.K, N = g.C * g.R * g.S, Kg = g.N * g.P * g.Q;
    const int b = (M >= 128 && N >= 128) ? 128 : 64;
    const int64_t tiles = (int64_t)((M + b - 1) / b) * ((N + b - 1) / b);
    int64_t want = (4LL * 148 + tiles - 1) / tiles;           // ~4 waves of CTAs on 148 SMs (fixed: results independent of device)
    int64_t max_splits = (Kg + 255) / 256;                    // at least 256 reduction elements per split
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    if (want > 512) want = 512;
    int chunk = (int)((Kg + want - 1) / want);
    chunk = (chunk + 15) / 16 * 16;
    *splits = (Kg + chunk - 1) / chunk;
    *k_chunk = chunk;
    *bm = b;
}

}  // namespace clb

using namespace clb;

extern "C" {

int clb_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, float* w_ws, int N, int C, int H, int W,
                   int K, int R, int S, int stride, int pad, int relu, void* stream) {
    CLB_CHECK_ARG(x && w && y && N > 0 && C > 0 && H > 0 && W > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0);
    CLB_CHECK_ARG(H + 2 * pad >= R && W + 2 * pad >= S);
    CLB_CHECK_ARG((int64_t)N * C * H * W < (1LL << 31) && (int64_t)N * K * H * W < (1LL << 31));
    ConvGeom g = make_geom(N, C, H, W, K, R, S, stride, pad);
    if (mm_mode() != CLB_MM_FP32_SIMT && w_ws != nullptr && tc_fwd_supported(C, H, W, K, R, S, stride, pad)) {
        if (tc_bf16_route(C)) {
            int rc4 = tc_conv_fwd_bf16(x, w, w_ws, bias, y, N, C, H, W, K, R, S, pad, relu, as_stream(stream));
            if (rc4) return rc4;
            CLB_CHECK_LAUNCH();
            return CLB_OK;
        }
        tc_permute_w_fwd(w, w_ws, K, C, R * S, as_stream(stream));           // [K][C][RS] -> [K][RS][C]
        int rc = tc_conv_fwd(x, w_ws, bias, y, N, C, H, W, K, R, S, pad, relu, mm_split(), as_stream(stream));
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
        return CLB_OK;
    }
    conv_fwd_simt(x, w, bias, y, g, relu, as_stream(stream));
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_conv2d_dgrad(const float* dy, const float* w, float* dx, float* wt_ws, int N, int C, int H, int W, int K, int R,
                     int S, int stride, int pad, void* stream) {
    CLB_CHECK_ARG(dy && w && dx && N > 0 && C > 0 && H > 0 && W > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0);
    cudaStream_t s = as_stream(stream);
    ConvGeom g = make_geom(N, C, H, W, K, R, S, stride, pad);
    if (mm_mode() != CLB_MM_FP32_SIMT && wt_ws != nullptr && tc_dgrad_supported(C, H, W, K, R, S, stride, pad)) {
        if (tc_bf16_route(K)) {
            int rc4 = tc_conv_dgrad_bf16(dy, w, wt_ws, dx, N, C, g.P, g.Q, K, R, S, pad, s);
            if (rc4) return rc4;
            CLB_CHECK_LAUNCH();
            return CLB_OK;
        }
        tc_permute_w_dgrad(w, wt_ws, K, C, R, S, s);                          // [K][C][R][S] -> [C][flipped RS][K]
        int rc = tc_conv_fwd(dy, wt_ws, nullptr, dx, N, K, g.P, g.Q, C, R, S, R - 1 - pad, 0, mm_split(), s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
        return CLB_OK;
    }
    if (stride == 1 && R - 1 - pad >= 0 && S - 1 - pad >= 0 && R == S) {
        CLB_CHECK_ARG(wt_ws != nullptr);
        const int64_t total = (int64_t)K * C * R * S;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sm_count() * 8) blocks = sm_count() * 8;
        flip_transpose_weights_kernel<<<blocks, 256, 0, s>>>(w, wt_ws, K, C, R, S); clb::count_launch();
        // forward conv over dY: input [N, K, P, Q], output [N, C, H, W], pad' = R-1-pad
        ConvGeom gd = make_geom(N, K, g.P, g.Q, C, R, S, 1, R - 1 - pad);
        CLB_CHECK_ARG(gd.P == H && gd.Q == W);
        conv_fwd_simt(dy, wt_ws, nullptr, dx, gd, 0, s);
    } else {
        const int64_t total = (int64_t)N * C * H * W;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sm_count() * 16) blocks = sm_count() * 16;
        conv_dgrad_naive_kernel<<<blocks, 256, 0, s>>>(dy, w, dx, g); clb::count_launch();
    }
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

size_t clb_conv2d_wgrad_ws(int N, int C, int H, int W, int K, int R, int S, int stride, int pad) {
    ConvGeom g = make_geom(N, C, H, W, K, R, S, stride, pad);
    int splits, chunk, bm;
    wgrad_plan(g, &splits, &chunk, &bm);
    size_t need = (size_t)splits * K * C * R * S * sizeof(float);
    if (tc_wgrad_supported(C, H, W, K, R, S, stride, pad)) {
        const size_t t = tc_wgrad_ws_floats(N, C, H, W, K, R, S) * sizeof(float);
        if (t > need) need = t;
    }
    return need + (size_t)64 * K * sizeof(float);           // + bias-gradient partials
}

int clb_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, float* ws, size_t ws_bytes, int N, int C,
                     int H, int W, int K, int R, int S, int stride, int pad, void* stream) {
    CLB_CHECK_ARG(x && dy && dw && N > 0 && C > 0 && H > 0 && W > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0);
    CLB_CHECK_ARG((int64_t)N * C * H * W < (1LL << 31) && (int64_t)N * K * H * W < (1LL << 31));
    cudaStream_t s = as_stream(stream);
    ConvGeom g = make_geom(N, C, H, W, K, R, S, stride, pad);
    const int M = K, Ng = C * R * S, Kg = N * g.P * g.Q;
    if (mm_mode() != CLB_MM_FP32_SIMT && tc_wgrad_supported(C, H, W, K, R, S, stride, pad)) {
        const size_t tneed = tc_wgrad_ws_floats(N, C, H, W, K, R, S) * sizeof(float);
        if (ws == nullptr || ws_bytes < tneed) {
            set_error("clb_conv2d_wgrad: workspace %zu bytes < required %zu", ws_bytes, tneed);
            return CLB_EWORKSPACE;
        }
        if (dbias && ws_bytes < tneed + (size_t)64 * K * sizeof(float)) { set_error("clb_conv2d_wgrad: workspace too small for bias partials"); return CLB_EWORKSPACE; }
        float* bias_part = dbias ? ws + tneed / sizeof(float) : nullptr;
        bool partials_done = false;
        int rc = tc_conv_wgrad(x, dy, dw, ws, bias_part, &partials_done, N, C, H, W, K, R, S, pad, mm_split(), s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
        if (dbias) {
            if (partials_done) conv_bias_grad_finish(bias_part, dbias, N, K, s);
            else conv_bias_grad(dy, dbias, bias_part, N, K, g.P * g.Q, s);
            CLB_CHECK_LAUNCH();
        }
        return CLB_OK;
    }
    int splits, chunk, bm;
    wgrad_plan(g, &splits, &chunk, &bm);
    const size_t need = (size_t)splits * M * Ng * sizeof(float);
    if (splits > 1 && (ws == nullptr || ws_bytes < need)) {
        set_error("clb_conv2d_wgrad: workspace %zu bytes < required %zu", ws_bytes, need);
        return CLB_EWORKSPACE;
    }
    DyKoutOuter A; A.dy = dy; A.K = K; A.PQ = g.P * g.Q; A.n_pix = Kg; A.dPQ = g.dPQ;
    Im2colCrsOuter B; B.x = x; B.g = g; B.n_pix = Kg; B.n_crs = Ng;
    EpiRM e; e.c = splits > 1 ? ws : dw; e.ldc = Ng; e.bias = nullptr; e.relu = 0; e.M = M; e.N = Ng;
    if (bm == 128) {
        dim3 grid((M + 127) / 128, (Ng + 127) / 128, splits);
        gemm_simt_kernel<8, 8, DyKoutOuter, Im2colCrsOuter, false, false, EpiRM, false><<<grid, 256, 0, s>>>(
            A, B, e, Kg, chunk, (int64_t)M * Ng); clb::count_launch();
    } else {
        dim3 grid((M + 63) / 64, (Ng + 63) / 64, splits);
        gemm_simt_kernel<4, 4, DyKoutOuter, Im2colCrsOuter, false, false, EpiRM, false><<<grid, 256, 0, s>>>(
            A, B, e, Kg, chunk, (int64_t)M * Ng); clb::count_launch();
    }
    CLB_CHECK_LAUNCH();
    if (splits > 1) {
        const int64_t n = (int64_t)M * Ng;
        int blocks = (int)((n + 255) / 256);
        if (blocks > sm_count() * 8) blocks = sm_count() * 8;
        splitk_reduce_kernel<<<blocks, 256, 0, s>>>(ws, dw, n, splits); clb::count_launch();
        CLB_CHECK_LAUNCH();
    }
    if (dbias) {
        const size_t off = splits > 1 ? need : 0;
        if (ws == nullptr || ws_bytes < off + (size_t)64 * K * sizeof(float)) { set_error("clb_conv2d_wgrad: workspace too small for bias partials"); return CLB_EWORKSPACE; }
        conv_bias_grad(dy, dbias, ws + off / sizeof(float), N, K, g.P * g.Q, s);
        CLB_CHECK_LAUNCH();
    }
    return CLB_OK;
}

size_t clb_linear_ws(int M, int in, int out) { return tc3_linear_ws_floats(M, in, out) * sizeof(float); }

static bool linear_on_tc(const float* ws, size_t ws_bytes, int M, int in, int out) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("CLB_TC_LINEAR"); enabled = (e && e[0] == '0') ? 0 : 1; }
    return enabled && mm_mode() != CLB_MM_FP32_SIMT && ws != nullptr && tc3_linear_supported(M, in, out) &&
           ws_bytes >= tc3_linear_ws_floats(M, in, out) * sizeof(float);
}

int clb_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* ws, size_t ws_bytes, int M, int in,
                   int out, int relu, void* stream) {
    CLB_CHECK_ARG(x && w && y && M > 0 && in > 0 && out > 0);
    cudaStream_t s = as_stream(stream);
    if (linear_on_tc(ws, ws_bytes, M, in, out)) {
        int rc = tc3_linear_fwd(x, w, bias, y, ws, M, in, out, relu, mm_split(), s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
        return CLB_OK;
    }
    StridedMat A{x, in, 1, M, in};
    StridedMat B{w, in, 1, out, in};
    EpiRM e; e.c = y; e.ldc = out; e.bias = bias; e.relu = relu; e.M = M; e.N = out;
    if (out <= 32) {
        dim3 grid((M + 127) / 128, (out + 31) / 32, 1);
        gemm_simt_kernel<8, 2, StridedMat, StridedMat, false, false, EpiRM, false><<<grid, 256, 0, s>>>(A, B, e, in, in, 0); clb::count_launch();
    } else {
        dim3 grid((M + 63) / 64, (out + 63) / 64, 1);
        gemm_simt_kernel<4, 4, StridedMat, StridedMat, false, false, EpiRM, false><<<grid, 256, 0, s>>>(A, B, e, in, in, 0); clb::count_launch();
    }
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_linear_dgrad(const float* dy, const float* w, float* dx, float* ws, size_t ws_bytes, int M, int in, int out,
                     void* stream) {
    CLB_CHECK_ARG(dy && w && dx && M > 0 && in > 0 && out > 0);
    cudaStream_t s = as_stream(stream);
    if (linear_on_tc(ws, ws_bytes, M, in, out)) {
        int rc = tc3_linear_dgrad(dy, w, dx, ws, M, in, out, mm_split(), s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
        return CLB_OK;
    }
    StridedMat A{dy, out, 1, M, out};        // element(m, o)
    StridedMat B{w, 1, in, in, out};         // element(n = i, k = o) = w[o*in + i]
    EpiRM e; e.c = dx; e.ldc = in; e.bias = nullptr; e.relu = 0; e.M = M; e.N = in;
    dim3 grid((M + 63) / 64, (in + 63) / 64, 1);
    gemm_simt_kernel<4, 4, StridedMat, StridedMat, false, true, EpiRM, false><<<grid, 256, 0, s>>>(A, B, e, out, out, 0); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_linear_wgrad(const float* x, const float* dy, float* dw, float* dbias, float* ws, size_t ws_bytes, int M, int in,
                     int out, void* stream) {
    CLB_CHECK_ARG(x && dy && dw && M > 0 && in > 0 && out > 0);
    cudaStream_t s = as_stream(stream);
    if (linear_on_tc(ws, ws_bytes, M, in, out)) {
        int rc = tc3_linear_wgrad(x, dy, dw, ws, M, in, out, mm_split(), s);
        if (rc) return rc;
        CLB_CHECK_LAUNCH();
        if (dbias) {
            linear_bias_grad_kernel<<<(out + 31) / 32, 256, 0, s>>>(dy, dbias, M, out); clb::count_launch();
            CLB_CHECK_LAUNCH();
        }
        return CLB_OK;
    }
    StridedMat A{dy, 1, out, out, M};        // element(m' = o, k = m) = dy[m*out + o]
    StridedMat B{x, 1, in, in, M};           // element(n = i, k = m) = x[m*in + i]
    EpiRM e; e.c = dw; e.ldc = in; e.bias = nullptr; e.relu = 0; e.M = out; e.N = in;
    if (out <= 32) {
        dim3 grid((out + 31) / 32, (in + 127) / 128, 1);
        gemm_simt_kernel<2, 8, StridedMat, StridedMat, true, true, EpiRM, false><<<grid, 256, 0, s>>>(A, B, e, M, M, 0); clb::count_launch();
    } else {
        dim3 grid((out + 63) / 64, (in + 63) / 64, 1);
        gemm_simt_kernel<4, 4, StridedMat, StridedMat, true, true, EpiRM, false><<<grid, 256, 0, s>>>(A, B, e, M, M, 0); clb::count_launch();
    }
    CLB_CHECK_LAUNCH();
    if (dbias) {
        linear_bias_grad_kernel<<<(out + 31) / 32, 256, 0, s>>>(dy, dbias, M, out); clb::count_launch();
        CLB_CHECK_LAUNCH();
    }
    return CLB_OK;
}

}  // extern "C"
