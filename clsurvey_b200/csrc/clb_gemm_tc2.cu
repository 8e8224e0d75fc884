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
#include "clb_tc_loaders.cuh"

namespace clb {
namespace tc2 {
using namespace clb::tc;
using namespace clb::tcl;

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
#ifdef CLB_TC_RAW_HI            // experiment: does the tensor core truncate the low 13 mantissa bits itself?
            {
                uint32_t raw[BK];
#pragma unroll
                for (int j = 0; j < BK; ++j) raw[j] = __float_as_uint(v[j]);
                tmem_st32(lane_field + col, raw);
            }
#else
            tmem_st32(lane_field + col, hi);
#endif
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
