// tcgen05 implicit-GEMM, second generation: the A operand never touches shared memory.
//
// ncu on the first generation (profiles/r1_ncu_full_gemm_tc_v1.csv) showed the tensor pipe 44-48 % active: with both
// operands staged in smem the 3-pass TF32 split writes 64 KB and the MMAs read 96 KB of smem per K block -- 1280 clk of
// the 128 B/clk smem port against 768 clk of MMA -- and the loader warps sit on LDG latency.  Here:
//   * A rows (one 128-byte K block row per loader thread) go registers -> TMEM with tcgen05.st (TMEM write port,
//     256 B/clk) as hi / lo TF32 planes, and the MMA takes A from TMEM (tcgen05.mma [d], [a_tmem], b_desc ...);
//   * only B (BN x 32, hi / lo) is staged in swizzled smem: 32 KB written + 48 KB read per K block = 640 clk < 768 clk;
//   * loaders issue their global loads BEFORE waiting for a free stage, two A groups alternate K blocks, and the B
//     loader warp group prefetches one K block ahead.
// Warp roles (default 544 threads): 2 A loader groups (warps 0-7, warp%4 = TMEM lane quarter), 2 B loader groups (warps
// 8-15), warp 16 = TMEM allocator + MMA issuer; warps 0-7 run the epilogue.  Group counts are compile-time (-DCLB_TC2_GROUPS*).
// TMEM columns: [0,BN) hi*hi accumulator, [BN,2BN) cross-term accumulator, then 4 A stages x (32 hi + 32 lo) columns = 512.
#include <limits.h>

#include "clb_tc_ptx.cuh"

namespace clb {
namespace tc2 {
using namespace clb::tc;

#ifndef CLB_TC2_GROUPS
#define CLB_TC2_GROUPS 2
#endif
constexpr int kGroupsA = CLB_TC2_GROUPS;                   // A loader groups (4 warps each); K block i belongs to group i % kGroupsA
#ifndef CLB_TC2_GROUPS_B
#define CLB_TC2_GROUPS_B 2
#endif
constexpr int kGroupsB = CLB_TC2_GROUPS_B;                   // B loader groups (4 warps each); K block i belongs to group i % kGroupsB
constexpr int kWarpsA = 4 * kGroupsA;
constexpr int kWarpMma = kWarpsA + 4 * kGroupsB;             // first warp after the loaders issues the MMAs
constexpr int kThreads = (kWarpMma + 1) * 32;                // default 2+2 groups: 17 warps = 544 threads (<= 96 regs)
constexpr int kStagesA = 4;                                  // TMEM A stages (4 x 64 columns + 256 accumulator columns = 512)

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
        st_shared_v4(tile_hi + off, h0, h1, h2, h3);
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

// ---------------------------------------------------------------------------------------------- kernel
template <int BN, int STAGES_B, bool WITH_LO> struct Layout {
    static constexpr int kBTile = BN * 128;
    static constexpr int kStageB = kBTile * (WITH_LO ? 2 : 1);
    static constexpr int kBarOff = kStageB * STAGES_B;
    static constexpr int kTotal = kBarOff + 256 + 1024;
    static constexpr int kAccCols = WITH_LO ? 2 * BN : BN;
    static constexpr int kAStageCols = WITH_LO ? 64 : 32;
    static constexpr int kColsNeeded = kAccCols + kStagesA * kAStageCols;
    static constexpr int kTmemCols = kColsNeeded <= 128 ? 128 : (kColsNeeded <= 256 ? 256 : 512);
    static_assert(kColsNeeded <= 512, "TMEM budget");
};

template <int BN, int STAGES_B, bool WITH_LO, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(ALoad A, BLoad B, Epi epi, int num_kb_total, int kb_per_split) {
    using L = Layout<BN, STAGES_B, WITH_LO>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + L::kBarOff;
    const uint32_t a_full = bar, a_empty = bar + 8 * kStagesA;
    const uint32_t b_full = bar + 16 * kStagesA, b_empty = b_full + 8 * STAGES_B;
    const uint32_t bar_tmem = b_empty + 8 * STAGES_B, tmem_slot = bar_tmem + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, z = blockIdx.z;
    const int kb_begin = z * kb_per_split;
    const int nkb = max(min(num_kb_total, kb_begin + kb_per_split) - kb_begin, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesA; ++s) { mbar_init(a_full + 8 * s, 4); mbar_init(a_empty + 8 * s, 1); }
        for (int s = 0; s < STAGES_B; ++s) { mbar_init(b_full + 8 * s, 4); mbar_init(b_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
    }
    if (warp == kWarpMma) tmem_alloc(tmem_slot, L::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tmem_a0 = tmem_base + L::kAccCols;

    if (warp < kWarpsA) {
        // ---------------- A loaders: registers -> TMEM; group g owns K blocks with i % kGroupsA == g; thread = one GEMM row.
        // ncu (r1_ncu_full_gemm_tc_v1): with 2 groups the loader warps ran at ~0.25 IPC on dependent address math and
        // LDG latency (~3300 clk per group per K block vs 768 clk of MMA): more groups in flight, not more bandwidth.
        const int group = warp >> 2, tg = threadIdx.x & 127;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const typename ALoad::Ctx actx = A.prep(m0 + tg);
        for (int i = group; i < nkb; i += kGroupsA) {
            float v[BK];
            A.row(kb_begin + i, actx, v);                      // global loads in flight before we block on the stage
            const int s = i % kStagesA;
            const uint32_t it = (uint32_t)(i / kStagesA);
            mbar_wait(a_empty + 8 * s, (it & 1u) ^ 1u);
            tc_fence_after();
            uint32_t hi[BK];
#pragma unroll
            for (int j = 0; j < BK; ++j) hi[j] = __float_as_uint(v[j]) & kHiMask;
            const uint32_t col = tmem_a0 + (uint32_t)s * L::kAStageCols;
            tmem_st32(lane_field + col, hi);
            if (WITH_LO) {
                uint32_t lo[BK];
#pragma unroll
                for (int j = 0; j < BK; ++j) lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                tmem_st32(lane_field + col + 32, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * s);
        }
    } else if (warp < kWarpMma) {
        // ---------------- B loaders: gmem -> registers -> swizzled smem hi / lo; group g owns K blocks i % kGroupsB == g
        // (a single B group was the gen-2 bottleneck: ~300 dependent instructions per thread per K block at < 0.25 IPC)
        const int tg = threadIdx.x & 127, group = (warp - kWarpsA) >> 2;
        const typename BLoad::Ctx bctx = B.prep(tg, n0);
        for (int i = group; i < nkb; i += kGroupsB) {
            BRegs<BN> cur;
            B.load(kb_begin + i, tg, bctx, cur);                // loads in flight before blocking on the stage
            const int s = i % STAGES_B;
            const uint32_t it = (uint32_t)(i / STAGES_B);
            mbar_wait(b_empty + 8 * s, (it & 1u) ^ 1u);
            const uint32_t st = base + (uint32_t)s * L::kStageB;
            store_b<BN, WITH_LO>(cur, tg, st, st + L::kBTile);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_full + 8 * s);
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer
        constexpr uint32_t idesc = make_idesc(BN);
        for (int i = 0; i < nkb; ++i) {
            const int sa = i % kStagesA, sb = i % STAGES_B;
            mbar_wait(a_full + 8 * sa, (uint32_t)(i / kStagesA) & 1u);
            mbar_wait(b_full + 8 * sb, (uint32_t)(i / STAGES_B) & 1u);
            tc_fence_after();
            const uint32_t st = base + (uint32_t)sb * L::kStageB;
            const uint64_t b_hi = make_desc(st), b_lo = make_desc(st + L::kBTile);
            const uint32_t a_hi = tmem_a0 + (uint32_t)sa * L::kAStageCols, a_lo = a_hi + 32;
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
                if (WITH_LO) {
                    umma_tf32_ts(tmem_base + BN, a_lo + 8 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                    umma_tf32_ts(tmem_base + BN, a_hi + 8 * k, b_lo + 2 * k, idesc, 1);
                }
                umma_tf32_ts(tmem_base, a_hi + 8 * k, b_hi + 2 * k, idesc, (i | k) != 0);
            }
            umma_commit(a_empty + 8 * sa);
            umma_commit(b_empty + 8 * sb);
        }
        umma_commit(bar_tmem);
    }

    if (warp < 8) {
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
    if (warp == kWarpMma) tmem_dealloc(tmem_base, L::kTmemCols);
}

template <int BN, int STAGES_B, bool WITH_LO, class ALoad, class BLoad, class Epi>
static int launch(const ALoad& A, const BLoad& B, const Epi& e, dim3 grid, int nkb_total, int kb_per_split, cudaStream_t s) {
    using L = Layout<BN, STAGES_B, WITH_LO>;
    auto kern = gemm_tc2_kernel<BN, STAGES_B, WITH_LO, ALoad, BLoad, Epi>;
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

}  // namespace tc2

int tc2_conv_fwd(const float* x, const float* w2, const float* bias, float* y, int N, int C, int H, int W, int K, int R,
                 int S, int pad, int relu, bool with_lo, cudaStream_t s) {
    using namespace tc2;
    const int P = H, Q = W, M = N * P * Q;
    EpiNCHW e{y, bias, relu, M, K, P * Q, FastDiv32(P * Q)};
    if (C % 32 != 0) {          // small-C first layer: w2 is [K][32] zero-padded, one K block
        PixelRowsSmallC A{x, C, H, W, R, S, pad, P, Q, M, R * S * C, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
        if (K > 64) {
            WeightRows<128> B{w2, K, 32, 32};
            dim3 grid((M + BM - 1) / BM, (K + 127) / 128, 1);
            return with_lo ? launch<128, 4, true>(A, B, e, grid, 1, 1, s) : launch<128, 4, false>(A, B, e, grid, 1, 1, s);
        }
        WeightRows<64> B{w2, K, 32, 32};
        dim3 grid((M + BM - 1) / BM, (K + 63) / 64, 1);
        return with_lo ? launch<64, 4, true>(A, B, e, grid, 1, 1, s) : launch<64, 4, false>(A, B, e, grid, 1, 1, s);
    }
    PixelRows A{x, C, H, W, R, S, pad, P, Q, M, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
    const int nkb = R * S * C / BK;
    if (K % 128 == 0 || K > 64) {
        WeightRows<128> B{w2, K, (int64_t)R * S * C, R * S * C};
        dim3 grid((M + BM - 1) / BM, (K + 127) / 128, 1);
        return with_lo ? launch<128, 4, true>(A, B, e, grid, nkb, nkb, s) : launch<128, 4, false>(A, B, e, grid, nkb, nkb, s);
    }
    WeightRows<64> B{w2, K, (int64_t)R * S * C, R * S * C};
    dim3 grid((M + BM - 1) / BM, (K + 63) / 64, 1);
    return with_lo ? launch<64, 4, true>(A, B, e, grid, nkb, nkb, s) : launch<64, 4, false>(A, B, e, grid, nkb, nkb, s);
}

void tc_wgrad_plan(int N, int C, int H, int W, int K, int R, int S, int* bn, int* splits, int* kb_per_split);

int tc2_conv_wgrad(const float* x, const float* dy, float* ws, int N, int C, int H, int W, int K, int R, int S, int pad,
                   bool with_lo, cudaStream_t s) {
    using namespace tc2;
    const int P = H, Q = W, npix = N * P * Q, n_rows = R * S * C;
    int bn, splits, per;
    tc_wgrad_plan(N, C, H, W, K, R, S, &bn, &splits, &per);
    const int nkb = (npix + BK - 1) / BK;
    DyRows A{dy, K, P * Q, npix, FastDiv32(P * Q)};
    EpiSplitK e{ws, K, n_rows, (int64_t)K * n_rows};
    if (bn == 128) {
        TapRows<128> B{x, C, H, W, R, S, pad, P, Q, n_rows, npix, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
        dim3 grid((K + BM - 1) / BM, (n_rows + 127) / 128, splits);
        return with_lo ? launch<128, 4, true>(A, B, e, grid, nkb, per, s) : launch<128, 4, false>(A, B, e, grid, nkb, per, s);
    }
    TapRows<64> B{x, C, H, W, R, S, pad, P, Q, n_rows, npix, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
    dim3 grid((K + BM - 1) / BM, (n_rows + 63) / 64, splits);
    return with_lo ? launch<64, 4, true>(A, B, e, grid, nkb, per, s) : launch<64, 4, false>(A, B, e, grid, nkb, per, s);
}

}  // namespace clb
