// Operand loaders (registers -> TMEM / swizzled smem) and epilogues shared by the gen-2 and gen-3 tcgen05 kernels.
#pragma once
#include <limits.h>

#include "clb_tc_ptx.cuh"

namespace clb {
namespace tcl {
using namespace clb::tc;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, M=128, K=8
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---------------------------------------------------------------------------------------------- A row loaders
// row(kb, m, v): the 32 K-block values of GEMM row m (zeros where out of range)

struct PixelRows {          // conv fwd / dgrad: row = output pixel, K = (r, s, c) over an NCHW tensor, stride 1
    const float* x; int C, H, W, R, S, pad, P, Q, M;
    FastDiv32 dPQ, dQ, dC, dS;
    struct Ctx { const float* pix; int p, q; bool ok; };          // per-thread, constant over the K loop
    __device__ __forceinline__ Ctx prep(int m) const {
        const uint32_t img = dPQ.div(m), pq = m - img * (P * Q);
        const uint32_t p = dQ.div(pq), q = pq - p * Q;
        return {x + (size_t)img * C * H * W + (int)p * W + (int)q, (int)p, (int)q, m < M};
    }
    __device__ __forceinline__ void row(int kb, const Ctx& t, float (&v)[BK]) const {
        const uint32_t k0 = (uint32_t)kb * BK;                     // warp-uniform tap decomposition
        const uint32_t rs = dC.div(k0), c0 = k0 - rs * C;
        const uint32_t r = dS.div(rs), s = rs - r * S;
        const int dr = (int)r - pad, ds = (int)s - pad;
        const bool ok = t.ok && (unsigned)(t.p + dr) < (unsigned)H && (unsigned)(t.q + ds) < (unsigned)W;
        const int HW = H * W;
        const float* src = t.pix + ((int)c0 * HW + dr * W + ds);
#pragma unroll
        for (int j = 0; j < BK; ++j) v[j] = ok ? __ldg(src + j * HW) : 0.f;
    }
};

struct PixelRowsSmallC {    // first layer (C*R*S <= 32, e.g. 3x3x3 = 27): the whole reduction is ONE zero-padded K block
    const float* x; int C, H, W, R, S, pad, P, Q, M, ktot;
    FastDiv32 dPQ, dQ, dC, dS;
    struct Ctx { int m; };
    __device__ __forceinline__ Ctx prep(int m) const { return {m}; }
    __device__ __forceinline__ void row(int /*kb*/, const Ctx& t, float (&v)[BK]) const {
        const int m = t.m;
        const uint32_t img = dPQ.div(m), pq = m - img * (P * Q);
        const uint32_t p = dQ.div(pq), q = pq - p * Q;
        const float* base = x + (size_t)img * C * H * W;
        const bool mok = m < M;
#pragma unroll
        for (int j = 0; j < BK; ++j) {
            const uint32_t rs = dC.div(j), c = j - rs * C;
            const uint32_t r = dS.div(rs), s = rs - r * S;
            const int ih = (int)p + (int)r - pad, iw = (int)q + (int)s - pad;
            const bool ok = mok && j < ktot && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
            v[j] = ok ? __ldg(base + ((size_t)c * H + ih) * W + iw) : 0.f;
        }
    }
};

struct DyRows {             // conv wgrad: row = output channel kout, K = pixel: dy[img][kout][pq], 16-byte chunks
    const float* dy; int K, PQ, npix; FastDiv32 dPQ;
    struct Ctx { const float* rowp; bool ok; };
    __device__ __forceinline__ Ctx prep(int m) const { return {dy + (size_t)m * PQ, m < K}; }
    __device__ __forceinline__ void row(int kb, const Ctx& t, float (&v)[BK]) const {
        const size_t img_stride = (size_t)K * PQ;
        if ((PQ & 31) == 0) {                                      // the whole K block lies in one image (uniform)
            const int pix0 = kb * BK;
            const uint32_t img = dPQ.div(pix0), pq0 = pix0 - img * PQ;
            const bool ok = t.ok && pix0 < npix;
            const float4* src = reinterpret_cast<const float4*>(t.rowp + img * img_stride + pq0);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = ok ? __ldg(src + c) : make_float4(0, 0, 0, 0);
                v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int pix = kb * BK + c * 4;
                const uint32_t img = dPQ.div(pix), pq = pix - img * PQ;
                const bool ok = t.ok && pix < npix;
                const float4 q = ok ? __ldg(reinterpret_cast<const float4*>(t.rowp + img * img_stride + pq))
                                    : make_float4(0, 0, 0, 0);
                v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------- B tile loaders (smem)
template <int ROWS> struct BRegs { float4 v[ROWS / 16]; };

template <int ROWS>
struct WeightRows {         // rows contiguous along K: w2[rows][ld]
    const float* p; int n_rows; int64_t ld; int k_total;
    struct Ctx { const float* rp[ROWS / 16]; };                   // per-thread row pointers (+ chunk offset), NULL = masked
    __device__ __forceinline__ Ctx prep(int tg, int row0) const {
        Ctx t;
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i) {
            const int r = row0 + (tg >> 3) + 16 * i;
            t.rp[i] = r < n_rows ? p + (int64_t)r * ld + (tg & 7) * 4 : nullptr;
        }
        return t;
    }
    __device__ __forceinline__ void load(int kb, int tg, const Ctx& t, BRegs<ROWS>& g) const {
        const int kk = kb * BK;
        const bool kok = kk + (tg & 7) * 4 < k_total;
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i)
            g.v[i] = (kok && t.rp[i]) ? __ldg(reinterpret_cast<const float4*>(t.rp[i] + kk)) : make_float4(0, 0, 0, 0);
    }
};

template <int ROWS>
struct TapRows {            // conv wgrad B: row = (r, s, c) tap, K = pixel (4 consecutive pixels of an image row per chunk)
    const float* x; int C, H, W, R, S, pad, P, Q, n_rows, k_total;
    FastDiv32 dPQ, dQ, dC, dS;
    struct Ctx { int off[ROWS / 16]; short dr[ROWS / 16], ds[ROWS / 16]; };   // off = c*H*W + dr*W + ds, or -1 = masked row
    __device__ __forceinline__ Ctx prep(int tg, int row0) const {
        Ctx t;
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i) {
            const int n = row0 + (tg >> 3) + 16 * i;
            const uint32_t rs = dC.div(n), c = n - rs * C;
            const uint32_t r = dS.div(rs), s = rs - r * S;
            t.dr[i] = (short)((int)r - pad);
            t.ds[i] = (short)((int)s - pad);
            t.off[i] = n < n_rows ? (int)c * H * W + t.dr[i] * W + t.ds[i] : INT_MIN;
        }
        return t;
    }
    __device__ __forceinline__ void load(int kb, int tg, const Ctx& t, BRegs<ROWS>& g) const {
        const int pix = kb * BK + (tg & 7) * 4;
        const uint32_t img = dPQ.div(pix), pq = pix - img * (P * Q);
        const uint32_t p = dQ.div(pq), q0 = pq - p * Q;
        const bool kok = pix < k_total;
        const float* base = x + (size_t)img * C * H * W + (int)p * W + (int)q0;
#pragma unroll
        for (int i = 0; i < ROWS / 16; ++i) {
            const int ih = (int)p + t.dr[i], iw0 = (int)q0 + t.ds[i];
            const bool rok = kok && t.off[i] != INT_MIN && (unsigned)ih < (unsigned)H;
            const float* src = base + (rok ? t.off[i] : 0);
            float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = (rok && (unsigned)(iw0 + j) < (unsigned)W) ? __ldg(src + j) : 0.f;
            g.v[i] = make_float4(u[0], u[1], u[2], u[3]);
        }
    }
};

template <int ROWS, bool WITH_LO>
__device__ __forceinline__ void store_b(const BRegs<ROWS>& g, int tg, uint32_t tile_hi, uint32_t tile_lo) {
    const int c = tg & 7;
    const int r0 = tg >> 3;                       // rows r0 + 16 i: (r & 7) == (r0 & 7) for every i -> one swizzled offset
    const uint32_t off0 = (uint32_t)r0 * 128u + (uint32_t)((c ^ (r0 & 7)) << 4);
#pragma unroll
    for (int i = 0; i < ROWS / 16; ++i) {
        const uint32_t off = off0 + (uint32_t)i * 2048u;
        const float4 v = g.v[i];
        const uint32_t h0 = __float_as_uint(v.x) & kHiMask, h1 = __float_as_uint(v.y) & kHiMask;
        const uint32_t h2 = __float_as_uint(v.z) & kHiMask, h3 = __float_as_uint(v.w) & kHiMask;
#ifdef CLB_TC_RAW_HI
        st_shared_v4(tile_hi + off, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
#else
        st_shared_v4(tile_hi + off, h0, h1, h2, h3);
#endif
        if (WITH_LO)
            st_shared_v4(tile_lo + off, __float_as_uint(v.x - __uint_as_float(h0)), __float_as_uint(v.y - __uint_as_float(h1)),
                         __float_as_uint(v.z - __uint_as_float(h2)), __float_as_uint(v.w - __uint_as_float(h3)));
    }
}

// ---------------------------------------------------------------------------------------------- epilogues
struct EpiNCHW {
    float* y; const float* bias; int relu, M, N, PQ; FastDiv32 dPQ;
    __device__ __forceinline__ void store16(int m, int n0, const uint32_t (&r)[16], int) const {
        if (m >= M) return;
        const uint32_t img = dPQ.div(m), pq = m - img * PQ;
        float* dst = y + ((size_t)img * N + n0) * PQ + pq;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < N) {
                float v = __uint_as_float(r[j]) + (bias ? __ldg(bias + n0 + j) : 0.f);
                dst[(size_t)j * PQ] = relu ? fmaxf(v, 0.f) : v;
            }
    }
};
struct EpiSplitK {
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


}  // namespace tcl
}  // namespace clb
