// GEM gradient-memory kernels (a15, a16): fused multi-dot + Gram in one HBM pass, device-side QP, fused projection.
//
// Reference: src/methods/rehearsal/model/gem.py
//   dotp = torch.mm(grads[:, t].unsqueeze(0), grads.index_select(1, indx))   :275-276
//   (dotp < 0).sum()                                                          :277
//   project2cone2: P = M M^T (fp64, host), +eps I, quadprog, x = v M + g      :58-80
// The reference keeps grads as [P, n_tasks] (column stride 40 B) and ships the whole [P,k] matrix to the host in fp64
// for every violation.  Here the memory is task-major G[n_tasks][ld]; one pass reads g and the k rows once
// (4(k+1) B/param) and produces both the k dots and the k x k Gram matrix in fp64; the k x k QP is solved by one CTA
// (2^k active sets, one per thread) and the projection is a second streaming pass (4(k+2) B/param).
#include "clb_common.cuh"

namespace clb {

constexpr int kMaxK = 10;
constexpr int kGemThreads = 256;

template <int K>
__global__ void __launch_bounds__(kGemThreads)
gem_dots_gram_kernel(const float* __restrict__ g, const float* __restrict__ G, int64_t ld, int64_t P,
                     const int* __restrict__ idx, double* __restrict__ dots, double* __restrict__ gram) {
    constexpr int NPAIR = K * (K + 1) / 2;
    double acc_d[K], acc_g[NPAIR];
#pragma unroll
    for (int i = 0; i < K; ++i) acc_d[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NPAIR; ++i) acc_g[i] = 0.0;
    const float* rows[K];
#pragma unroll
    for (int i = 0; i < K; ++i) rows[i] = G + (int64_t)idx[i] * ld;

    // Products and short partial sums in fp32 (the reference's torch.mm accumulates ALL of P in fp32, gem.py:275-276), flushed
    // into the fp64 accumulators every U float4 groups (<= 16 terms per partial sum: the fp32 rounding of such a partial is
    // ~1e-7 of its own magnitude and unbiased).  U independent groups of loads are in flight per thread; with the fp64 FMAs
    // of the first version the kernel was bound by the DFMA pipe at every k (58-61 % of HBM, profiles/r1_stream_kernels.jsonl).
    constexpr int U = K <= 3 ? 4 : 2;
    const int64_t n4 = P >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
        float4 gv[U], m[U][K];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * stride;
            const bool ok = i < n4;
            gv[u] = ok ? __ldcs(reinterpret_cast<const float4*>(g) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int a = 0; a < K; ++a)
                m[u][a] = ok ? __ldcs(reinterpret_cast<const float4*>(rows[a]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float s_d[K], s_g[NPAIR];
#pragma unroll
        for (int a = 0; a < K; ++a) s_d[a] = 0.f;
#pragma unroll
        for (int a = 0; a < NPAIR; ++a) s_g[a] = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int a = 0; a < K; ++a)
                s_d[a] = fmaf(gv[u].x, m[u][a].x, fmaf(gv[u].y, m[u][a].y, fmaf(gv[u].z, m[u][a].z, fmaf(gv[u].w, m[u][a].w, s_d[a]))));
            int p = 0;
#pragma unroll
            for (int a = 0; a < K; ++a)
#pragma unroll
                for (int b = a; b < K; ++b) {
                    s_g[p] = fmaf(m[u][a].x, m[u][b].x, fmaf(m[u][a].y, m[u][b].y, fmaf(m[u][a].z, m[u][b].z, fmaf(m[u][a].w, m[u][b].w, s_g[p]))));
                    ++p;
                }
        }
#pragma unroll
        for (int a = 0; a < K; ++a) acc_d[a] += (double)s_d[a];
#pragma unroll
        for (int a = 0; a < NPAIR; ++a) acc_g[a] += (double)s_g[a];
    }
    if (blockIdx.x == 0 && threadIdx.x < (P & 3)) {  // scalar tail
        const int64_t e = (n4 << 2) + threadIdx.x;
        const float gv = g[e];
        float m[K];
#pragma unroll
        for (int a = 0; a < K; ++a) m[a] = rows[a][e];
#pragma unroll
        for (int a = 0; a < K; ++a) acc_d[a] += (double)gv * m[a];
        int p = 0;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = a; b < K; ++b) acc_g[p++] += (double)m[a] * m[b];
    }
    // block reduction: warp shuffle, then one smem hop, then ONE atomic per CTA per output
    __shared__ double red[kGemThreads / 32][K + NPAIR];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        double v = warp_sum(acc_d[a]);
        if (lane == 0) red[wid][a] = v;
    }
#pragma unroll
    for (int a = 0; a < NPAIR; ++a) {
        double v = warp_sum(acc_g[a]);
        if (lane == 0) red[wid][K + a] = v;
    }
    __syncthreads();
    if (threadIdx.x < K + NPAIR) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kGemThreads / 32; ++w) v += red[w][threadIdx.x];
        if (threadIdx.x < K) {
            atomicAdd(&dots[threadIdx.x], v);
        } else {
            // unpack pair index -> (a,b), write both triangles
            int p = threadIdx.x - K, a = 0, rowlen = K;
            while (p >= rowlen) { p -= rowlen; ++a; --rowlen; }
            const int b = a + p;
            atomicAdd(&gram[a * K + b], v);
            if (a != b) atomicAdd(&gram[b * K + a], v);
        }
    }
}

// ---- QP: min 1/2 v'Pv + dots'v  s.t. v >= margin, P = 0.5(gram+gram^T) + eps I -------------------
struct QpCandidate {
    double viol, obj;
};

// Solve the KKT system of active set `mask` (bit i set => v_i fixed at margin). Returns violation + objective.
__host__ __device__ inline QpCandidate qp_solve_subset(const double* Pm, const double* dots, int k, double margin,
                                                       unsigned mask, double* v /*[kMaxK]*/) {
    int F[kMaxK], nf = 0;
    for (int i = 0; i < k; ++i) {
        v[i] = margin;
        if (!((mask >> i) & 1u)) F[nf++] = i;
    }
    double A[kMaxK][kMaxK + 1];
    for (int r = 0; r < nf; ++r) {
        double rhs = -dots[F[r]];
        for (int j = 0; j < k; ++j)
            if ((mask >> j) & 1u) rhs -= Pm[F[r] * k + j] * margin;
        for (int c = 0; c < nf; ++c) A[r][c] = Pm[F[r] * k + F[c]];
        A[r][nf] = rhs;
    }
    // Gaussian elimination with partial pivoting (SPD principal minor; pivoting only for robustness)
    for (int c = 0; c < nf; ++c) {
        int piv = c;
        double best = A[c][c] < 0 ? -A[c][c] : A[c][c];
        for (int r = c + 1; r < nf; ++r) {
            double t = A[r][c] < 0 ? -A[r][c] : A[r][c];
            if (t > best) { best = t; piv = r; }
        }
        if (piv != c)
            for (int j = c; j <= nf; ++j) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        const double inv = 1.0 / A[c][c];
        for (int r = c + 1; r < nf; ++r) {
            const double f = A[r][c] * inv;
            for (int j = c; j <= nf; ++j) A[r][j] -= f * A[c][j];
        }
    }
    for (int r = nf - 1; r >= 0; --r) {
        double s = A[r][nf];
        for (int c = r + 1; c < nf; ++c) s -= A[r][c] * v[F[c]];
        v[F[r]] = s / A[r][r];
    }
    QpCandidate out;
    double viol = 0.0, obj = 0.0;
    for (int i = 0; i < k; ++i) {
        double lam = dots[i];
        for (int j = 0; j < k; ++j) lam += Pm[i * k + j] * v[j];
        obj += v[i] * (0.5 * (lam - dots[i]) + dots[i]);
        if ((mask >> i) & 1u) {
            if (-lam > viol) viol = -lam;           // dual feasibility on active bounds
        } else {
            if (margin - v[i] > viol) viol = margin - v[i];   // primal feasibility on free variables
        }
    }
    out.viol = viol;
    out.obj = obj;
    return out;
}

__host__ __device__ inline void qp_build(const double* dots, const double* gram, int k, double eps, double* Pm,
                                         double* scale_out, int* nviol_out) {
    double scale = 1.0;
    int nv = 0;
    for (int i = 0; i < k; ++i) {
        for (int j = 0; j < k; ++j) {
            double p = 0.5 * (gram[i * k + j] + gram[j * k + i]) + (i == j ? eps : 0.0);
            Pm[i * k + j] = p;
            double ap = p < 0 ? -p : p;
            if (ap > scale) scale = ap;
        }
        double ad = dots[i] < 0 ? -dots[i] : dots[i];
        if (ad > scale) scale = ad;
        if (dots[i] < 0) ++nv;
    }
    *scale_out = scale;
    *nviol_out = nv;
}

__global__ void gem_qp_kernel(const double* __restrict__ dots, const double* __restrict__ gram, int k, double margin,
                              double eps, double* __restrict__ v_out, int* __restrict__ viol_out) {
    __shared__ double Pm[kMaxK * kMaxK];
    __shared__ double sd[kMaxK];
    __shared__ double s_scale;
    __shared__ int s_nviol;
    __shared__ double s_viol[1 << kMaxK];
    __shared__ double s_obj[1 << kMaxK];
    __shared__ int s_win;
    if (threadIdx.x == 0) {
        for (int i = 0; i < k; ++i) sd[i] = dots[i];
        qp_build(sd, gram, k, eps, Pm, &s_scale, &s_nviol);
    }
    __syncthreads();
    if (s_nviol == 0) {
        if (threadIdx.x < k) v_out[threadIdx.x] = 0.0;
        if (threadIdx.x == 0) viol_out[0] = 0;
        return;
    }
    const unsigned nsub = 1u << k;
    double v[kMaxK];
    if (threadIdx.x < nsub) {
        QpCandidate c = qp_solve_subset(Pm, sd, k, margin, threadIdx.x, v);
        s_viol[threadIdx.x] = c.viol;
        s_obj[threadIdx.x] = c.obj;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double tol = 1e-9 * s_scale;
        int win = -1, fb = 0;
        double best_obj = 0.0, best_viol = s_viol[0];
        for (unsigned s = 0; s < nsub; ++s) {
            const double vi = s_viol[s];
            if (vi == vi && vi <= tol && (win < 0 || s_obj[s] < best_obj)) { win = (int)s; best_obj = s_obj[s]; }
            if (vi == vi && vi < best_viol) { best_viol = vi; fb = (int)s; }
        }
        s_win = win >= 0 ? win : fb;
        viol_out[0] = s_nviol;
    }
    __syncthreads();
    if ((int)threadIdx.x == s_win)
        for (int i = 0; i < k; ++i) v_out[i] = v[i];
}

template <int K>
__global__ void __launch_bounds__(kGemThreads)
gem_project_kernel(float* __restrict__ g, const float* __restrict__ G, int64_t ld, int64_t P,
                   const int* __restrict__ idx, const double* __restrict__ v, const int* __restrict__ viol) {
    if (viol[0] == 0) return;
    const float* rows[K];
    double vv[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { rows[i] = G + (int64_t)idx[i] * ld; vv[i] = v[i]; }
    const int64_t n4 = P >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gv = reinterpret_cast<const float4*>(g)[i];
        double x = 0, y = 0, z = 0, w = 0;
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const float4 m = __ldcs(reinterpret_cast<const float4*>(rows[a]) + i);
            x += vv[a] * m.x; y += vv[a] * m.y; z += vv[a] * m.z; w += vv[a] * m.w;
        }
        gv.x = (float)(x + (double)gv.x); gv.y = (float)(y + (double)gv.y);
        gv.z = (float)(z + (double)gv.z); gv.w = (float)(w + (double)gv.w);
        reinterpret_cast<float4*>(g)[i] = gv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (P & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        double x = 0;
#pragma unroll
        for (int a = 0; a < K; ++a) x += vv[a] * rows[a][e];
        g[e] = (float)(x + (double)g[e]);
    }
}

static inline int gem_grid(int64_t n_vec, int per_sm) {
    int64_t blocks = (n_vec + kGemThreads - 1) / kGemThreads;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

template <int K>
static int launch_dots_gram(const float* g, const float* G, int64_t ld, int64_t P, const int* idx, double* dots,
                            double* gram, cudaStream_t s) {
    gem_dots_gram_kernel<K><<<gem_grid(P >> 2, K <= 2 ? 8 : (K <= 4 ? 4 : 2)), kGemThreads, 0, s>>>(g, G, ld, P, idx, dots, gram); clb::count_launch();
    return 0;
}
template <int K>
static int launch_project(float* g, const float* G, int64_t ld, int64_t P, const int* idx, const double* v,
                          const int* viol, cudaStream_t s) {
    gem_project_kernel<K><<<gem_grid(P >> 2, 4), kGemThreads, 0, s>>>(g, G, ld, P, idx, v, viol); clb::count_launch();
    return 0;
}

}  // namespace clb

using namespace clb;

#define CLB_K_SWITCH(k, CALL)                                       \
    switch (k) {                                                    \
        case 1: CALL(1); break;  case 2: CALL(2); break;            \
        case 3: CALL(3); break;  case 4: CALL(4); break;            \
        case 5: CALL(5); break;  case 6: CALL(6); break;            \
        case 7: CALL(7); break;  case 8: CALL(8); break;            \
        case 9: CALL(9); break;  case 10: CALL(10); break;          \
        default: break;                                             \
    }

extern "C" {

int clb_gem_dots_gram(const float* g, const float* G, int64_t ld, int64_t P, const int* idx_dev, int k, double* dots,
                      double* gram, void* stream) {
    CLB_CHECK_ARG(g && G && idx_dev && dots && gram && k >= 1 && k <= kMaxK && P >= 0 && ld >= P);
    CLB_CHECK_ARG((ld & 3) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)G & 15) == 0);
    cudaStream_t s = as_stream(stream);
#define CALL(KK) launch_dots_gram<KK>(g, G, ld, P, idx_dev, dots, gram, s)
    CLB_K_SWITCH(k, CALL)
#undef CALL
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_gem_solve_qp(const double* dots, const double* gram, int k, double margin, double eps, double* v, int* viol,
                     void* stream) {
    CLB_CHECK_ARG(dots && gram && v && viol && k >= 1 && k <= kMaxK);
    const int threads = (1 << k) < 32 ? 32 : (1 << k);
    gem_qp_kernel<<<1, threads, 0, as_stream(stream)>>>(dots, gram, k, margin, eps, v, viol); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

int clb_gem_solve_qp_host(const double* dots, const double* gram, int k, double margin, double eps, double* v,
                          int* viol) {
    CLB_CHECK_ARG(dots && gram && v && viol && k >= 1 && k <= kMaxK);
    double Pm[kMaxK * kMaxK], scale;
    int nv;
    qp_build(dots, gram, k, eps, Pm, &scale, &nv);
    viol[0] = nv;
    if (nv == 0) {
        for (int i = 0; i < k; ++i) v[i] = 0.0;
        return CLB_OK;
    }
    const double tol = 1e-9 * scale;
    double best_obj = 0.0, best_viol = 0.0, cand[kMaxK], fbv[kMaxK];
    bool have = false, have_fb = false;
    for (unsigned s = 0; s < (1u << k); ++s) {
        QpCandidate c = qp_solve_subset(Pm, dots, k, margin, s, cand);
        if (c.viol != c.viol) continue;
        if (c.viol <= tol && (!have || c.obj < best_obj)) {
            have = true;
            best_obj = c.obj;
            for (int i = 0; i < k; ++i) v[i] = cand[i];
        }
        if (!have_fb || c.viol < best_viol) {
            have_fb = true;
            best_viol = c.viol;
            for (int i = 0; i < k; ++i) fbv[i] = cand[i];
        }
    }
    if (!have)
        for (int i = 0; i < k; ++i) v[i] = fbv[i];
    return CLB_OK;
}

int clb_gem_project(float* g, const float* G, int64_t ld, int64_t P, const int* idx_dev, int k, const double* v,
                    const int* viol, void* stream) {
    CLB_CHECK_ARG(g && G && idx_dev && v && viol && k >= 1 && k <= kMaxK && P >= 0 && ld >= P);
    CLB_CHECK_ARG((ld & 3) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)G & 15) == 0);
    cudaStream_t s = as_stream(stream);
#define CALL(KK) launch_project<KK>(g, G, ld, P, idx_dev, v, viol, s)
    CLB_K_SWITCH(k, CALL)
#undef CALL
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

}  // extern "C"
