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
// Operands are gathered by loader warps straight from the reference's NCHW / [K,C,R,S]-derived layouts (LDG ->
// split -> swizzled STS.128, nothing is materialised in HBM); a single elected thread issues tcgen05.mma; smem stages
// are recycled through mbarriers signalled by tcgen05.commit; the epilogue reads the accumulator with tcgen05.ld.
// GEMM-K order is (r, s, c) so that one 32-wide K block has a single (r, s): one bounds test per block per pixel.
//
// Warp roles (288 threads): warps 0-3 load even K blocks, warps 4-7 load odd K blocks (two groups keep two blocks of
// global loads in flight), warp 8 allocates TMEM and issues the MMAs; warps 0-7 then run the epilogue.
#include <stdlib.h>

#include "clb_tc_ptx.cuh"

namespace clb {
namespace tc {

constexpr int kThreads = 288;

// store one 4-float chunk (row r, 16-byte chunk c) of a K block into the swizzled hi / lo tiles
template <bool WITH_LO>
__device__ __forceinline__ void store_chunk(uint32_t tile_hi, uint32_t tile_lo, int r, int c, float v0, float v1, float v2,
                                            float v3) {
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    const uint32_t h0 = __float_as_uint(v0) & kHiMask, h1 = __float_as_uint(v1) & kHiMask;
    const uint32_t h2 = __float_as_uint(v2) & kHiMask, h3 = __float_as_uint(v3) & kHiMask;
    st_shared_v4(tile_hi + off, h0, h1, h2, h3);
    if (WITH_LO) {
        st_shared_v4(tile_lo + off, __float_as_uint(v0 - __uint_as_float(h0)), __float_as_uint(v1 - __uint_as_float(h1)),
                     __float_as_uint(v2 - __uint_as_float(h2)), __float_as_uint(v3 - __uint_as_float(h3)));
    }
}

// ---------------------------------------------------------------------------------------------- operand loaders
// Each loader fills a [ROWS x 32] K block: fill<WITH_LO>(kb, tg, row0, tile_hi, tile_lo), tg = thread in group (0..127).

// (1) pixel rows gathered from an NCHW tensor: row = output pixel, K = (r, s, c); used for A of fwd / dgrad.
struct PixelGather {
    const float* x; int C, H, W, R, S, pad, P, Q, M;    // M = N_img * P * Q rows in total; stride 1
    FastDiv32 dPQ, dQ, dC, dS;
    template <bool WITH_LO>
    __device__ __forceinline__ void fill(int kb, int tg, int row0, uint32_t tile_hi, uint32_t tile_lo) const {
        const int m = row0 + tg;                         // one full 128-byte row per thread
        const uint32_t k0 = (uint32_t)kb * BK;
        const uint32_t rs = dC.div(k0), c0 = k0 - rs * C;
        const uint32_t r = dS.div(rs), s = rs - r * S;
        const uint32_t img = dPQ.div(m), pq = m - img * (P * Q);
        const uint32_t p = dQ.div(pq), q = pq - p * Q;
        const int ih = (int)p + (int)r - pad, iw = (int)q + (int)s - pad;
        const bool ok = m < M && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
        const float* src = x + ((size_t)img * C + c0) * H * W + (ok ? ih * W + iw : 0);
        const int HW = H * W;
        float v[BK];
#pragma unroll
        for (int j = 0; j < BK; ++j) v[j] = ok ? __ldg(src + (size_t)j * HW) : 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) store_chunk<WITH_LO>(tile_hi, tile_lo, tg, c, v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
};

// (2) rows that are contiguous along K in memory (16-byte aligned chunks): weights [rows][ldk] (B of fwd / dgrad) and
//     dY viewed as [kout][pixel] with an image stride (A of wgrad: addr = img*img_stride + row*ld + pq).
template <int ROWS>
struct RowsKContig {
    const float* p; int n_rows; int64_t ld; int k_total; int pq; int64_t img_stride; FastDiv32 dPQ;   // pq == 0: plain matrix
    template <bool WITH_LO>
    __device__ __forceinline__ void fill(int kb, int tg, int row0, uint32_t tile_hi, uint32_t tile_lo) const {
        const int chunk = tg & 7;
        const int kk = kb * BK + chunk * 4;
        int64_t koff = kk;
        if (pq != 0) {
            const uint32_t img = dPQ.div(kk);
            koff = (int64_t)img * img_stride + (kk - (int)img * pq);
        }
        const bool kok = kk < k_total;
        float4 v[ROWS / 16];
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i) {
            const int r = (tg >> 3) + 16 * i;
            const bool ok = kok && (row0 + r) < n_rows;
            v[i] = ok ? __ldg(reinterpret_cast<const float4*>(p + (int64_t)(row0 + r) * ld + koff)) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i)
            store_chunk<WITH_LO>(tile_hi, tile_lo, (tg >> 3) + 16 * i, chunk, v[i].x, v[i].y, v[i].z, v[i].w);
    }
};

// (3) im2col rows over pixels: row = (r, s, c) filter tap, K = pixel; B of wgrad.
template <int ROWS>
struct TapRowsOverPixels {
    const float* x; int C, H, W, R, S, pad, P, Q, n_rows, k_total;     // n_rows = R*S*C, k_total = N_img*P*Q
    FastDiv32 dPQ, dQ, dC, dS;
    template <bool WITH_LO>
    __device__ __forceinline__ void fill(int kb, int tg, int row0, uint32_t tile_hi, uint32_t tile_lo) const {
        const int chunk = tg & 7;
        const int pix = kb * BK + chunk * 4;             // 4 consecutive pixels of one image row (Q % 4 == 0)
        const uint32_t img = dPQ.div(pix), pq = pix - img * (P * Q);
        const uint32_t p = dQ.div(pq), q0 = pq - p * Q;
        const bool kok = pix < k_total;
        const float* img_base = x + (size_t)img * C * H * W;
        float v[ROWS / 16][4];
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i) {
            const int n = row0 + (tg >> 3) + 16 * i;
            const uint32_t rs = dC.div(n), c = n - rs * C;
            const uint32_t r = dS.div(rs), s = rs - r * S;
            const int ih = (int)p + (int)r - pad;
            const int iw0 = (int)q0 + (int)s - pad;
            const bool rok = kok && n < n_rows && (unsigned)ih < (unsigned)H;
            const float* src = img_base + ((size_t)c * H + (rok ? ih : 0)) * W;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int iw = iw0 + j;
                v[i][j] = (rok && (unsigned)iw < (unsigned)W) ? __ldg(src + iw) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i)
            store_chunk<WITH_LO>(tile_hi, tile_lo, (tg >> 3) + 16 * i, chunk, v[i][0], v[i][1], v[i][2], v[i][3]);
    }
};

// ---------------------------------------------------------------------------------------------- epilogues
struct EpiNCHW {     // y[img][n][pq] = act(acc + bias[n]);  row m = img*PQ + pq  (lanes = consecutive pixels: coalesced)
    float* y; const float* bias; int relu, M, N, PQ; FastDiv32 dPQ;
    __device__ __forceinline__ void store16(int m, int n0, const uint32_t (&r)[16], int /*z*/) const {
        if (m >= M) return;
        const uint32_t img = dPQ.div(m), pq = m - img * PQ;
        float* dst = y + ((size_t)img * N + n0) * PQ + pq;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (n0 + j < N) {
                float v = __uint_as_float(r[j]) + (bias ? __ldg(bias + n0 + j) : 0.f);
                dst[(size_t)j * PQ] = relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
};
struct EpiSplitK {   // ws[z][m][n] row-major partial sums (wgrad); reduced in fixed order by a second kernel
    float* ws; int M, N; int64_t split_stride;
    __device__ __forceinline__ void store16(int m, int n0, const uint32_t (&r)[16], int z) const {
        if (m >= M) return;
        float* dst = ws + (int64_t)z * split_stride + (int64_t)m * N + n0;
        if (n0 + 15 < N && (N & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + j < N) dst[j] = __uint_as_float(r[j]);
        }
    }
};

// ---------------------------------------------------------------------------------------------- the kernel
template <int BN, int STAGES, bool WITH_LO> struct SmemLayout {
    static constexpr int kATile = BM * 128, kBTile = BN * 128;
    static constexpr int kStage = (kATile + kBTile) * (WITH_LO ? 2 : 1);
    static constexpr int kBarOff = kStage * STAGES;
    static constexpr int kTotal = kBarOff + 256 + 1024;     // barriers + tmem slot, + slack for 1024-B alignment
};

template <int BN, int STAGES, bool WITH_LO, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(ALoad A, BLoad B, Epi epi, int num_kb_total, int kb_per_split) {
    using L = SmemLayout<BN, STAGES, WITH_LO>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + L::kBarOff, bar_empty = bar_full + 8 * STAGES, bar_tmem = bar_empty + 8 * STAGES;
    const uint32_t tmem_slot = bar_tmem + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, z = blockIdx.z;
    const int kb_begin = z * kb_per_split;
    const int kb_end = min(num_kb_total, kb_begin + kb_per_split);
    const int nkb = max(kb_end - kb_begin, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 4);       // one arrive per loader warp of the owning group
            mbar_init(bar_empty + 8 * s, 1);      // tcgen05.commit
        }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
    }
    // two accumulators when splitting: columns [0,BN) take hi*hi, columns [BN,2BN) take the ~2^-11 smaller cross terms.
    // The tensor core's fp32 accumulate truncates; keeping the small terms apart cuts the number of (biased) roundings
    // applied to the large accumulator by 3x.
    constexpr uint32_t kTmemCols = WITH_LO ? 2 * BN : BN;
    if (warp == 8) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 8) {
        // ---------------- loaders: group g handles K blocks with (i % 2) == g
        const int group = warp >> 2, tg = threadIdx.x & 127;
        for (int i = group; i < nkb; i += 2) {
            const int s = i % STAGES;
            const uint32_t it = (uint32_t)(i / STAGES);
            mbar_wait(bar_empty + 8 * s, (it & 1u) ^ 1u);
            const uint32_t st = base + (uint32_t)s * L::kStage;
            const uint32_t a_hi = st, b_hi = st + L::kATile;
            const uint32_t a_lo = st + L::kATile + L::kBTile, b_lo = a_lo + L::kATile;
            A.template fill<WITH_LO>(kb_begin + i, tg, m0, a_hi, a_lo);
            B.template fill<WITH_LO>(kb_begin + i, tg, n0, b_hi, b_lo);
            fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer (one thread)
        constexpr uint32_t idesc = make_idesc(BN);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            const uint32_t it = (uint32_t)(i / STAGES);
            mbar_wait(bar_full + 8 * s, it & 1u);
            tc_fence_after();
            const uint32_t st = base + (uint32_t)s * L::kStage;
            const uint64_t a_hi = make_desc(st), b_hi = make_desc(st + L::kATile);
            const uint64_t a_lo = make_desc(st + L::kATile + L::kBTile), b_lo = make_desc(st + 2 * L::kATile + L::kBTile);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {     // +32 bytes (>>4 = 2) per K=8 step inside the swizzle row
                if (WITH_LO) {
                    umma_tf32(tmem_base + BN, a_lo + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                    umma_tf32(tmem_base + BN, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
                    umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                } else {
                    umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                }
            }
            umma_commit(bar_empty + 8 * s);        // frees the smem stage once these MMAs have read it
        }
        umma_commit(bar_tmem);                     // accumulator complete
    }

    if (warp < 8) {
        // ---------------- epilogue: warp w owns TMEM lanes 32*(w%4).., column half (w/4)
        if (nkb > 0) {
            mbar_wait(bar_tmem, 0);
            tc_fence_after();
        }
        const int lane_grp = warp & 3, col_half = warp >> 2;
        const int m = m0 + lane_grp * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 16) {
            const int col = col_half * (BN / 2) + c;
            uint32_t r[16];
            if (nkb > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)col, r);
                if (WITH_LO) {
                    uint32_t r2[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(BN + col), r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            epi.store16(m, n0 + col, r, z);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, kTmemCols);
}

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

template <int BN, int STAGES, bool WITH_LO, class ALoad, class BLoad, class Epi>
static int launch(const ALoad& A, const BLoad& B, const Epi& e, dim3 grid, int nkb_total, int kb_per_split, cudaStream_t s) {
    using L = SmemLayout<BN, STAGES, WITH_LO>;
    auto kern = gemm_tc_kernel<BN, STAGES, WITH_LO, ALoad, BLoad, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (err != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%d): %s", L::kTotal, cudaGetErrorString(err));
            return CLB_ECUDA;
        }
        configured = true;
    }
    kern<<<grid, kThreads, L::kTotal, s>>>(A, B, e, nkb_total, kb_per_split); clb::count_launch();
    return CLB_OK;
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

// CLB_TC_IMPL: 1 = first generation (both operands gathered into smem), 2 = A through TMEM, 3 (default) = gen-2 plus
// TMA-fed kernels where the memory layout allows it (wgrad on >= 8x8 maps)
static int tc_impl() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("CLB_TC_IMPL");
        v = (e && e[0] >= '1' && e[0] <= '3') ? (e[0] - '0') : 3;
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
    return small_c && tc_impl() >= 2 && C * R * S <= 32;
}
bool tc_dgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    return same_conv(H, W, R, S, stride, pad) && (K % 32) == 0;
}
bool tc_wgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    if (!same_conv(H, W, R, S, stride, pad) || (W % 4) != 0) return false;
    return tc_impl() >= 2 || (C % 32) == 0;
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
    if (tc_impl() >= 2) return tc2_conv_fwd(x, w2, bias, y, N, C, H, W, K, R, S, pad, relu, with_lo, s);
    const int P = H, Q = W, M = N * P * Q;
    PixelGather A{x, C, H, W, R, S, pad, P, Q, M, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
    EpiNCHW e{y, bias, relu, M, K, P * Q, FastDiv32(P * Q)};
    const int nkb = R * S * C / BK;
    if (K % 128 == 0 || K > 64) {
        RowsKContig<128> B{w2, K, (int64_t)R * S * C, R * S * C, 0, 0, FastDiv32(1)};
        dim3 grid((M + BM - 1) / BM, (K + 127) / 128, 1);
        return with_lo ? launch<128, 3, true>(A, B, e, grid, nkb, nkb, s) : launch<128, 4, false>(A, B, e, grid, nkb, nkb, s);
    }
    RowsKContig<64> B{w2, K, (int64_t)R * S * C, R * S * C, 0, 0, FastDiv32(1)};
    dim3 grid((M + BM - 1) / BM, (K + 63) / 64, 1);
    return with_lo ? launch<64, 4, true>(A, B, e, grid, nkb, nkb, s) : launch<64, 4, false>(A, B, e, grid, nkb, nkb, s);
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
    const int nkb = (npix + BK - 1) / BK;
    RowsKContig<128> A{dy, K, (int64_t)P * Q, npix, P * Q, (int64_t)K * P * Q, FastDiv32(P * Q)};
    EpiSplitK e{ws, K, n_rows, (int64_t)K * n_rows};
    int rc;
    if (tc_impl() >= 2) {
        rc = tc2_conv_wgrad(x, dy, ws, N, C, H, W, K, R, S, pad, with_lo, s);
    } else if (bn == 128) {
        TapRowsOverPixels<128> B{x, C, H, W, R, S, pad, P, Q, n_rows, npix, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
        dim3 grid((K + BM - 1) / BM, (n_rows + 127) / 128, splits);
        rc = with_lo ? launch<128, 3, true>(A, B, e, grid, nkb, per, s) : launch<128, 4, false>(A, B, e, grid, nkb, per, s);
    } else {
        TapRowsOverPixels<64> B{x, C, H, W, R, S, pad, P, Q, n_rows, npix, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
        dim3 grid((K + BM - 1) / BM, (n_rows + 63) / 64, splits);
        rc = with_lo ? launch<64, 4, true>(A, B, e, grid, nkb, per, s) : launch<64, 4, false>(A, B, e, grid, nkb, per, s);
    }
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
